#!/usr/bin/env python
"""Where does the end-to-end step lose time against the device-resident one?  Streams the same batch through
B200Captioner.beam_search_stream with (a) device-resident input, (b) pinned fp16 host input, (c) pinned fp32 host input."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpleimagecaptionzoo_b200 import engine, synth  # noqa: E402

B, K, T, steps = 1536, 3, 20, 30
dims = synth.DIMS["BUTD"]
sd = synth.make_state_dict("BUTD", seed=0, **dims)
settings = dict(model_type="BUTDDetection", embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"], atten_dim=dims["atten_dim"])
cap = engine.B200Captioner("BUTDDetection", settings, dims["vocab_size"], sd, max_batch=B, max_regions=36, max_rows=K, max_seq=T)
host32 = torch.from_numpy(synth.make_region_feats(B, 36, 2048, 1)).pin_memory()
host16 = host32.half().pin_memory()
dev32 = host32.cuda()
dev16 = host16.cuda()


def run(x, n):
    for _ in cap.beam_search_stream(({"bu_feats": x} for _ in range(n)), beam_size=K, max_seq=T):
        pass
    torch.cuda.synchronize()


def direct(x, n):
    for _ in range(n):
        cap.decoder.prepare(x)
        cap.decoder.beam_search(K, T)
    torch.cuda.synchronize()


for name, fn, x in (("direct dev fp32", direct, dev32), ("direct dev fp16", direct, dev16), ("stream dev fp32", run, dev32),
                    ("stream dev fp16", run, dev16), ("stream host fp16", run, host16), ("stream host fp32", run, host32)):
    fn(x, 3)
    t0 = time.perf_counter()
    fn(x, steps)
    dt = time.perf_counter() - t0
    print(f"{name:18s} {1e3 * dt / steps:7.3f} ms/step  {B * steps / dt:9.0f} captions/s", flush=True)
