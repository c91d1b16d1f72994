#!/usr/bin/env python
"""Bring-up diagnostics for the GPU box: runs each check in isolation (own try/except) and prints a summary.
Usage: python tests/tools/gpu_check.py [gemm] [tiny] [full] [sample]   (default: all).  Output is also useful as a log
under gpurun_out/.  The numpy oracle is used here as the CHECKER only."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import capdec_oracle as orc  # noqa: E402
from simpleimagecaptionzoo_b200 import capdec  # noqa: E402
from tests.golden_util import case_names, load_case, rebuild  # noqa: E402

RESULTS = []


def record(name, ok, msg=""):
    RESULTS.append((name, ok, msg))
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {msg}", flush=True)


def check_gemm():
    g = torch.Generator(device="cpu").manual_seed(0)
    for (m, n, k) in [(128, 256, 64), (128, 256, 256), (200, 300, 128), (48, 32, 192), (1000, 4096, 1024), (3072, 1024, 1024)]:
        for math in ("f16", "f16x3"):
            try:
                a = torch.randn(m, k, generator=g).cuda()
                b = (torch.randn(n, k, generator=g) * 0.05).cuda()
                bias = torch.randn(n, generator=g).cuda()
                d = capdec.test_gemm(a, b, bias, math)
                torch.cuda.synchronize()
                if math == "f16":
                    ref = a.half().double() @ b.half().double().T + bias.double()
                    tol = 2e-5
                else:
                    ref = a.double() @ b.double().T + bias.double()
                    tol = 2e-5
                err = (d.double() - ref).abs().max().item()
                scale = ref.abs().max().item()
                record(f"gemm {m}x{n}x{k} {math}", err <= tol * max(scale, 1.0) * 10, f"max_abs_err={err:.3e} scale={scale:.2f}")
            except Exception as e:  # noqa: BLE001
                record(f"gemm {m}x{n}x{k} {math}", False, repr(e))
                traceback.print_exc()
                return False
    return True


def make_decoder(meta, math, max_rows=None):
    sd, feats, mask = rebuild(meta)
    dims = meta["dims"]
    arch = meta["arch"]
    rows = max_rows or max(meta["K"], meta.get("n_samples", 1))
    dec = capdec.CaptionDecoder(arch, sd, hidden_dim=dims["hidden_dim"], embed_dim=dims["embed_dim"],
                                vocab_size=dims["vocab_size"], atten_dim=dims.get("atten_dim", 0),
                                enc_dim=dims.get("enc_dim", 2048), num_heads=dims.get("num_heads", 8),
                                max_batch=meta["B"], max_regions=max(meta["R"], 1), max_rows=rows, max_seq=meta["T"], math=math)
    return dec, sd, feats, mask


def check_beam(name, math):
    meta, gold = load_case(name)
    dec, sd, feats, mask = make_decoder(meta, math)
    dec.prepare(torch.from_numpy(feats).cuda(), None if mask is None else torch.from_numpy(mask).cuda())
    tok, sc, ln = dec.beam_search(meta["K"], meta["T"])
    torch.cuda.synchronize()
    tok, sc, ln = tok.cpu().numpy(), sc.cpu().numpy(), ln.cpu().numpy()
    o = orc.make_decoder(meta["arch"], sd, num_heads=meta["dims"].get("num_heads", 8))
    o.prepare(feats, mask)
    res = orc.beam_search_batched(o, meta["K"], meta["T"])
    verdict = orc.agreement(tok, gold["tokens"], res.min_gap, tol=1e-4)
    exact = np.array([v == "exact" for v in verdict])
    n_exact, n_tie, n_diff = sum(v == "exact" for v in verdict), sum(v == "tie" for v in verdict), sum(v == "diff" for v in verdict)
    sc_err = float(np.abs(sc[exact] - gold["scores"][exact]).max()) if exact.any() else float("nan")
    len_ok = bool(np.array_equal(ln[exact], gold["lengths"][exact]))
    ok = n_diff == 0 and len_ok and (not exact.any() or sc_err < 1e-3) if math == "f16x3" else (n_exact + n_tie) >= min(0.9 * len(verdict), len(verdict) - 1)
    record(f"beam {name} {math}", ok, f"exact={n_exact} tie={n_tie} diff={n_diff} of {len(verdict)} score_err={sc_err:.2e} len_ok={len_ok}")
    if n_diff and math == "f16x3":
        bad = [i for i, v in enumerate(verdict) if v == "diff"][:3]
        for i in bad:
            print("   img", i, "cuda", tok[i].tolist(), "\n        gold", gold["tokens"][i].tolist(), "gaps", res.min_gap[i][:6])
    dec.close()


def check_sample(name, math):
    meta, gold = load_case(name)
    n = meta["n_samples"]
    dec, sd, feats, mask = make_decoder(meta, math, max_rows=max(n, 1))
    dec.prepare(torch.from_numpy(feats).cuda(), None if mask is None else torch.from_numpy(mask).cuda())
    gtok, _ = dec.sample(capdec.SAMPLE_GREEDY, 1, 0, meta["T"])
    stok, slp = dec.sample(capdec.SAMPLE_MULTINOMIAL, n, meta["sample_seed"], meta["T"])
    torch.cuda.synchronize()
    gtok = gtok.cpu().numpy()
    stok = stok.cpu().numpy().reshape(meta["B"], n, meta["T"])
    slp = slp.cpu().numpy().reshape(meta["B"], n, meta["T"])
    g_same = (gtok == gold["greedy"]).all(1)
    s_same = (stok == gold["sample_seq"]).all(-1)
    lp_err = float(np.abs(slp[s_same] - gold["sample_logprobs"][s_same]).max()) if s_same.any() else float("nan")
    ok = g_same.mean() >= 0.85 and s_same.mean() >= 0.85 and (lp_err < 2e-3 or math == "f16")
    record(f"sample {name} {math}", ok, f"greedy_same={g_same.mean():.3f} sample_same={s_same.mean():.3f} logprob_err={lp_err:.2e}")
    dec.close()


def main():
    what = set(sys.argv[1:]) or {"gemm", "tiny", "full", "sample"}
    print("device:", torch.cuda.get_device_name(0), "capability", torch.cuda.get_device_capability(0))
    t0 = time.time()
    if "gemm" in what:
        if not check_gemm():
            print("GEMM bring-up failed; skipping decode checks")
            return 1
    for kind in ("tiny", "full"):
        if kind not in what:
            continue
        for name in case_names(kind):
            for math in ("f16x3", "f16"):
                try:
                    check_beam(name, math)
                except Exception as e:  # noqa: BLE001
                    record(f"beam {name} {math}", False, repr(e))
                    traceback.print_exc()
    if "sample" in what:
        for name in case_names("tiny") + case_names("full"):
            for math in ("f16x3",):
                try:
                    check_sample(name, math)
                except Exception as e:  # noqa: BLE001
                    record(f"sample {name} {math}", False, repr(e))
                    traceback.print_exc()
    nfail = sum(1 for _, ok, _ in RESULTS if not ok)
    print(f"\n{len(RESULTS) - nfail}/{len(RESULTS)} checks passed in {time.time() - t0:.1f}s")
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
