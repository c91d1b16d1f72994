#!/usr/bin/env python
"""Caption agreement of both math modes with the REFERENCE's own tokens on the committed agreement sets (north_star:
>= 90 % exact-or-tie-justified on 5k synthetic images) -- a direct comparison: tests/golden/agree_*.npz hold the captions
the reference's code produced for these images (tests/golden/make_agreement_set.py) and the tie mask of the 1e-4 rule.

    python tests/tools/agreement.py [--set butd --images 5000] [--out gpurun_out/agreement.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import agreement_util as au  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="butd", choices=sorted(au.SETS))
    ap.add_argument("--images", type=int, default=5000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    out = {}
    for math in ("f16", "f16x3"):
        t0 = time.time()
        r = au.evaluate(args.set, args.images, math)
        r["seconds"] = time.time() - t0
        gaps = sorted(d["min_gap_up_to_step"] for d in r["diffs"])
        r["diff_gap_quantiles"] = {"min": gaps[0], "median": gaps[len(gaps) // 2], "max": gaps[-1]} if gaps else None
        out[math] = r
        print(f"{args.set} {math}: exact {r['exact']} tie {r['tie_justified']} diff {r['diff']} ({r['diff_with_later_sub_tol_gap']} with a later sub-tol gap) of {r['images']}"
              f" -> {r['exact_or_tie_frac']:.4f}  diff gaps {r['diff_gap_quantiles']}", flush=True)
    path = args.out or os.path.join(ROOT, "gpurun_out", f"agreement_{args.set}_{args.images}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
