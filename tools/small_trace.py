#!/usr/bin/env python
"""Timeline of the small-batch decode step (CAPDEC_TRACE=1): where a 16-image beam-search step spends its time.
    CAPDEC_TRACE=1 python tools/small_trace.py [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CAPDEC_TRACE"] = "1"
from simpleimagecaptionzoo_b200 import capdec, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dims = synth.DIMS["BUTD"]
sd = synth.make_state_dict("BUTD", seed=0, **dims)
dec = capdec.CaptionDecoder("BUTD", sd, hidden_dim=dims["hidden_dim"], embed_dim=dims["embed_dim"], vocab_size=dims["vocab_size"],
                            atten_dim=dims["atten_dim"], enc_dim=dims["enc_dim"], max_batch=B, max_regions=36, max_rows=3, max_seq=20)
f = torch.from_numpy(synth.make_region_feats(B, 36, 2048, 1)).cuda()
for _ in range(5):
    dec.prepare(f)
    dec.beam_search(3, 20)
torch.cuda.synchronize()
dec.debug_trace()
dec.prepare(f)
dec.beam_search(3, 20)
tr = dec.debug_trace()
tr = tr[tr[:, 0] > 0]
tr = tr[np.argsort(tr[:, 0])]
print("launches traced", len(tr))
names = ["entry", "setup", "q0 data", "q0 mma", "q0 tfull", "q0 ctr", "q0 epi", "q0 bar", "q1 data", "q1 mma", "q1 tfull", "q1 ctr", "q1 epi",
         "q1 bar", "exit"]
two = tr[tr[:, 8] > 0]
rel = (two[:, :15] - two[:, :1]) / 1e3
for kind, label in ((0, "even launches (top-down gates -> dec_att)"), (1, "odd launches (language gates -> logits)")):
    med = np.median(rel[4 + kind::2], axis=0)
    print(label)
    for n, v in zip(names, med):
        if n != "q1 bar":
            print(f"  {n:10s} {v:8.2f} us")
gaps = (tr[1:, 0] - tr[:-1, 14]) / 1e3
dur = (tr[:, 14] - tr[:, 0]) / 1e3
print("kernel durations (us): median", np.median(dur[4:]), " gaps between small kernels (attention / bookkeeping in between): median",
      np.median(gaps[4:]), "alternating:", np.round(gaps[4:12], 1))
print("step period (us):", np.median(tr[6::2, 0][1:] - tr[6::2, 0][:-1]) / 1e3)
