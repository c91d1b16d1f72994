"""Direct agreement of the CUDA decoder with the REFERENCE's own captions on the committed agreement sets
(tests/golden/agree_*.npz, written by tests/golden/make_agreement_set.py from /root/reference's code).

north_star: emitted token sequences must match the reference beam search exactly, except where the reference's top-k gap
is below 1e-4 ("tie-justified"); target >= 90 % exact-or-tie on 5k synthetic images."""
import json
import os

import numpy as np

from simpleimagecaptionzoo_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = {"butd": "BUTD", "aoa": "AOA", "aoa_bu": "AOA"}


def set_path(name, images):
    return os.path.join(GOLDEN_DIR, f"agree_{name}_{images}.npz")


def load_set(name, images):
    z = np.load(set_path(name, images))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def feats_for(name, lo, n, regions=36):
    """Inputs of images [lo, lo+n): the generator seeds per 1000-image chunk (== make_agreement_set.feats_for)."""
    dims = synth.DIMS[SETS[name]]
    if name in ("butd", "aoa_bu"):
        return synth.make_region_feats(n, regions, dims.get("enc_dim", 2048), 7000 + lo)
    return synth.make_refined_feats(n, regions, dims["hidden_dim"], 7000 + lo)


def make_decoder(name, math, chunk, beam, max_seq, regions=36):
    from simpleimagecaptionzoo_b200 import capdec
    arch = SETS[name]
    dims = synth.DIMS[arch]
    sd = synth.make_state_dict(arch, seed=0, **dims)
    if name == "aoa_bu":
        sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
    return capdec.CaptionDecoder(arch, sd, math=math, hidden_dim=dims["hidden_dim"], embed_dim=dims["embed_dim"],
                                 vocab_size=dims["vocab_size"], atten_dim=dims.get("atten_dim", 0), enc_dim=dims.get("enc_dim", 2048),
                                 num_heads=dims.get("num_heads", 8), max_batch=chunk, max_regions=regions, max_rows=beam, max_seq=max_seq)


def evaluate(name, images, math, limit=None):
    """Decode the set with the CUDA path in ``math`` mode and classify every caption against the reference's:
    exact / tie (a reference top-(k+1) gap < 1e-4 at or before the first divergence) / diff.  For every diff the smallest
    reference gap at or before the divergence step is recorded (a flip inside the operand-rounding error has a small one)."""
    import torch
    meta, gold = load_set(name, images)
    K, T, R, chunk = meta["beam"], meta["max_seq"], meta["regions"], meta["chunk"]
    n_img = min(limit or meta["images"], meta["images"])
    dec = make_decoder(name, math, chunk, K, T, R)
    exact = tie = late = 0
    diffs = []
    for lo in range(0, n_img, chunk):
        n = min(chunk, meta["images"] - lo)
        ft = torch.from_numpy(feats_for(name, lo, n, R)).cuda()
        (dec.prepare_bottom_up if name == "aoa_bu" else dec.prepare)(ft)
        tok, _, _ = dec.beam_search(K, T)
        tok = tok.cpu().numpy()
        n_use = min(n, n_img - lo)
        for i in range(n_use):
            want = gold["tokens"][lo + i].astype(np.int32)
            if np.array_equal(tok[i], want):
                exact += 1
                continue
            t = int(np.argmax(tok[i] != want))  # position p is produced at step p
            steps = max(t, 1)
            if int(gold["tie_bits"][lo + i]) & ((1 << steps) - 1):
                tie += 1
                continue
            # not justified by north_star's rule as this repo reads it (a sub-1e-4 gap at or before the first differing
            # position).  Reported beside it: whether a sub-1e-4 gap occurs LATER -- the returned hypothesis is chosen among
            # beams that already differ at that position, so a late exact tie flips the output too (image 1038 of the BUTD set:
            # its two best live beams reach gap 0.0 at steps 19-20 and the numpy oracle itself returns the other one) -- but
            # with random-init models late near-ties are common, so this class is NOT counted as justified.
            late_tie = bool(int(gold["tie_bits"][lo + i]) >> steps)
            late += late_tie
            g = float(gold["gaps"][lo + i, :steps].min()) if "gaps" in gold else float(gold["min_gap"][lo + i])
            diffs.append({"image": lo + i, "step": t, "min_gap_up_to_step": g, "later_sub_tol_gap": late_tie})
    dec.close()
    return {"set": name, "arch": meta["arch"], "images": n_img, "math": math, "beam": K, "max_seq": T, "exact": exact,
            "tie_justified": tie, "diff": len(diffs), "diff_with_later_sub_tol_gap": late,
            "exact_or_tie_frac": (exact + tie) / n_img, "exact_frac": exact / n_img,
            "tie_tolerance": meta["tol"], "diffs": diffs,
            "against": "the reference's own beam_search_sample tokens (tests/golden/make_agreement_set.py)"}
