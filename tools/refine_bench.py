#!/usr/bin/env python
"""Per-kernel-category device time of capdec_prepare_bottom_up (AoA projection + refiner + decoder prepare) alone.
Usage (GPU box): python tools/refine_bench.py [batch] [regions] [math]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpleimagecaptionzoo_b200 import capdec, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
R = int(sys.argv[2]) if len(sys.argv) > 2 else 36
math = sys.argv[3] if len(sys.argv) > 3 else "f16"
dims = synth.DIMS["AOA"]
sd = synth.make_state_dict("AOA", seed=0, **dims)
sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
dec = capdec.CaptionDecoder("AOA", sd, hidden_dim=dims["hidden_dim"], embed_dim=dims["embed_dim"], vocab_size=dims["vocab_size"],
                            enc_dim=2048, num_heads=dims["num_heads"], max_batch=B, max_regions=R, max_rows=3, max_seq=20, math=math)
bu = torch.from_numpy(synth.make_region_feats(B, R, 2048, 1)).cuda()
for _ in range(3):
    dec.prepare_bottom_up(bu)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dec.prepare_bottom_up(bu)
e1.record()
torch.cuda.synchronize()
print(f"prepare_bottom_up B={B} R={R} {math}: {e0.elapsed_time(e1) / 10:.3f} ms")
dec.profile(True)
for _ in range(3):
    dec.prepare_bottom_up(bu)
torch.cuda.synchronize()
for cat, (ms, fl, cnt) in dec.profile_read().items():
    if cnt:
        print(f"  {cat:12s} {ms / 3:8.3f} ms/call  {cnt // 3:3d} launches  " + (f"{fl / (ms * 1e-3) / 1e12:7.1f} TFLOP/s" if fl else ""))
