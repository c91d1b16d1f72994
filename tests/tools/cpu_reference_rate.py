#!/usr/bin/env python
"""CPU rate of the REFERENCE's own beam search next to the numpy port's (oracle), same images, same host, all cores -- runs
only where /root/reference exists (the build container).  bench.py's CPU arm on the GPU box can only time the port
(`cpu_baseline.kind: "port"`); this records how the two relate.

    python tests/tools/cpu_reference_rate.py [images]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_golden as mg  # noqa: E402
from oracle import capdec_oracle as orc  # noqa: E402
from simpleimagecaptionzoo_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.set_num_threads(os.cpu_count())
dims = dict(synth.DIMS["BUTD"])
sd = synth.make_state_dict("BUTD", seed=0, **dims)
ref = mg.build_decoder("BUTD", dims, sd)
ref.max_step_limit = 20
feats = synth.make_region_feats(n, 36, 2048, 1000)
tf = torch.from_numpy(feats)
with torch.no_grad():
    ref.beam_search_sample(tf[:1], beam_size=3)  # warm-up
    t0 = time.perf_counter()
    toks = []
    for b in range(n):
        seq, _ = ref.beam_search_sample(tf[b:b + 1], beam_size=3)
        toks.append([int(x) for x in seq[0].tolist()])
    t_ref = time.perf_counter() - t0
o = orc.make_decoder("BUTD", sd)
o.prepare(feats[:1])
orc.beam_search_reference_form(o, 3, 20)
t0 = time.perf_counter()
o.prepare(feats)
res = orc.beam_search_reference_form(o, 3, 20)
t_port = time.perf_counter() - t0
same = sum(list(res.tokens[b][:len(toks[b])]) == toks[b] for b in range(n))
out = {"workload": "BUTDDetection beam=3 max_seq=20, 36x2048 feats, V=9487, one image per call", "images": n, "cores": os.cpu_count(),
       "reference_captions_per_s": n / t_ref, "port_captions_per_s": n / t_port, "port_over_reference": t_ref / t_port,
       "identical_captions": same, "torch": torch.__version__,
       "note": "the port folds weight-norm once and hoists enc_att out of the step loop; the reference redoes both every step"}
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_cpu_reference_vs_port.json"), "w"), indent=1)
