#!/usr/bin/env python
"""GEMM shape sweep on the GPU box (tuning aid): TFLOP/s of the decode-step GEMMs in isolation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simpleimagecaptionzoo_b200 import capdec
torch.cuda.init()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4608
shapes = [("td_lstm", M, 4096, 2048, 1), ("lm_lstm", M, 4096, 4096, 1), ("logits", M, 9487, 1024, 3),
          ("dec_att", M, 1024, 1024, 0), ("proj", M * 12, 1024, 2048, 0), ("td_store", M, 4096, 2048, 0), ("lm_store", M, 4096, 4096, 0),
          ("big_store", 8192, 8192, 4096, 0)]
for name, m, n, k, epi in shapes:
    us = capdec.gemm_time_us(m, n, k, epi, "f16", 30)
    print(f"{name:10s} M={m} N={n} K={k} epi={epi}: {us:8.1f} us  {2.0*m*n*k/us/1e6:7.1f} TFLOP/s", flush=True)
