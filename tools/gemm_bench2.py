#!/usr/bin/env python
"""Isolated GEMM timings (capdec_test_gemm_time): python tools/gemm_bench2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpleimagecaptionzoo_b200 import capdec
for (m, n, k, epi, name) in [(4608, 9487, 1024, 3, "logits topk K=1024"), (4608, 9487, 512, 3, "logits topk K=512"),
                             (4608, 9487, 2048, 3, "logits topk K=2048"), (4608, 9472, 1024, 0, "store fp32 N=9472 K=1024"),
                             (4608, 4096, 2048, 1, "lstm K=2048"), (4608, 4096, 4096, 1, "lstm K=4096"),
                             (4608, 1024, 1024, 0, "dec_att"), (55296, 1024, 2048, 0, "projection")]:
    us = capdec.gemm_time_us(m, n, k, epi, "f16", 30)
    print(f"{name:28s} {us:8.1f} us  {2.0 * m * n * k / us / 1e6:8.1f} TFLOP/s")
