#!/usr/bin/env python
"""Reference captions for the agreement sets (north_star: >= 90 % exact-or-tie-justified on 5k synthetic images).

Runs only in the build container (needs /root/reference).  For every image of a set the REFERENCE's own
``beam_search_sample`` / ``beam_search_sampler`` (source-exec'd with the two shims of make_golden.py, one image per
call as Utils.py:72-73 forces) decodes the caption; the numpy oracle's batched form supplies the per-step
top-(k+1) gaps needed by the tie-justified rule.  The committed ``agree_*.npz`` hold

    tokens   int16  [N, 1+T]   reference tokens (<sta> first, <pad>=0 after <end>)
    tie_bits uint32 [N]        bit t-1 set <=> the oracle's top-(k+1) gap at step t is below 1e-4
    gaps     f16    [N, T]     the oracle's top-(k+1) gap at every step (how close a differing caption's decision was)
    min_gap  f32    [N]        smallest gap of the image (diagnostic)
    oracle_equal bool [N]      the oracle's caption == the reference's caption

and the inputs are regenerated from the seeds (``feats_for`` below == tests/tools/agreement.py).

    python tests/golden/make_agreement_set.py butd 5000
    python tests/golden/make_agreement_set.py aoa_bu 1000
    python tests/golden/make_agreement_set.py aoa 1000
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

K, T, R, CHUNK, TOL = 3, 20, 36, 1000, 1e-4
SETS = {"butd": "BUTD", "aoa": "AOA", "aoa_bu": "AOA"}


def feats_for(name, lo, n):
    """Inputs of images [lo, lo+n) of a set; lo is a multiple of CHUNK (the seed is per 1000-image chunk)."""
    from simpleimagecaptionzoo_b200 import synth
    dims = synth.DIMS[SETS[name]]
    if name in ("butd", "aoa_bu"):
        return synth.make_region_feats(n, R, dims.get("enc_dim", 2048), 7000 + lo)
    return synth.make_refined_feats(n, R, dims["hidden_dim"], 7000 + lo)


_W = {}


def _init(name):
    import torch
    torch.set_num_threads(1)
    import make_golden as mg
    from oracle import capdec_oracle as orc
    from simpleimagecaptionzoo_b200 import synth
    arch = SETS[name]
    dims = dict(synth.DIMS[arch])
    sd = synth.make_state_dict(arch, seed=0, **dims)
    if name == "aoa_bu":
        sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
        m = mg.load_reference("AoA_Model")
        ref = m.AoADetection_Captioner(vocab_size=dims["vocab_size"], num_heads=dims["num_heads"], hidden_dim=dims["hidden_dim"],
                                       embed_dim=dims["embed_dim"])
        ref.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()}, strict=True)
        ref.eval()
        ref.decoder.max_step_limit = T
    else:
        ref = mg.build_decoder(arch, dims, sd)
        ref.max_step_limit = T
    _W.update(name=name, arch=arch, sd=sd, ref=ref, orc=orc, oracle=orc.make_decoder(arch, sd), torch=torch)


def _work(job):
    lo, off, n = job  # images [lo+off, lo+off+n) of the chunk seeded 7000+lo
    torch, orc, name, ref = _W["torch"], _W["orc"], _W["name"], _W["ref"]
    f = feats_for(name, lo, CHUNK)[off:off + n]
    tokens = np.zeros((n, 1 + T), np.int16)
    with torch.no_grad():
        for b in range(0 if os.environ.get("AGREE_GAPS_ONLY") == "1" else n):
            fb = torch.from_numpy(f[b:b + 1])
            if name == "aoa_bu":
                seq_t = ref.beam_search_sampler({"bu_feats": fb, "bu_masks": None}, beam_size=K)
            elif name == "aoa":
                seq_t, _ = ref.beam_search_sample(fb, beam_size=K, bu_masks=None)
            else:
                seq_t, _ = ref.beam_search_sample(fb, beam_size=K)
            seq = [int(x) for x in seq_t[0].tolist()]
            tokens[b, :len(seq)] = seq
    o = _W["oracle"]
    o.prepare(orc.aoa_project_refine(_W["sd"], f) if name == "aoa_bu" else f)
    res = orc.beam_search_batched(o, K, T)
    bits = np.zeros(n, np.uint32)
    for t in range(T):
        bits |= (res.min_gap[:, t] < TOL).astype(np.uint32) << np.uint32(t)
    return lo + off, tokens, bits, res.min_gap.min(1).astype(np.float32), (res.tokens == tokens).all(1), np.minimum(res.min_gap, 1e3).astype(np.float16)


def main():
    name, n_img = sys.argv[1], int(sys.argv[2])
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, (os.cpu_count() or 2) - 2)
    per = 25
    jobs = [(lo, off, min(per, min(CHUNK, n_img - lo) - off)) for lo in range(0, n_img, CHUNK)
            for off in range(0, min(CHUNK, n_img - lo), per)]
    tokens = np.zeros((n_img, 1 + T), np.int16)
    bits = np.zeros(n_img, np.uint32)
    gap = np.zeros(n_img, np.float32)
    gaps = np.zeros((n_img, T), np.float16)
    same = np.zeros(n_img, bool)
    t0 = time.time()
    done = 0
    with mp.get_context("fork").Pool(procs, initializer=_init, initargs=(name,)) as pool:
        for start, tk, bt, g, eq, gs in pool.imap_unordered(_work, jobs):
            n = tk.shape[0]
            tokens[start:start + n], bits[start:start + n], gap[start:start + n], same[start:start + n] = tk, bt, g, eq
            gaps[start:start + n] = gs
            done += n
            print(f"[{done}/{n_img}] oracle==reference {int(same.sum())}  ({time.time() - t0:.0f}s)", flush=True)
    import torch
    meta = dict(set=name, arch=SETS[name], images=n_img, beam=K, max_seq=T, regions=R, tol=TOL, chunk=CHUNK, seed_base=7000,
                torch=torch.__version__, oracle_equal=int(same.sum()))
    path = os.path.join(ROOT, "tests", "golden", f"agree_{name}_{n_img}.npz")
    if os.environ.get("AGREE_GAPS_ONLY") == "1":  # add the per-step gaps to a set whose reference tokens exist already
        old = np.load(path)
        assert (old["tie_bits"] == bits).all()
        tokens, same, meta = old["tokens"], old["oracle_equal"], json.loads(str(old["meta"]))
    np.savez_compressed(path, tokens=tokens, tie_bits=bits, min_gap=gap, gaps=gaps, oracle_equal=same, meta=np.array(json.dumps(meta)))
    print(json.dumps(meta), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
