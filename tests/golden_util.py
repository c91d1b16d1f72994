"""Shared helpers: load a golden case and rebuild its weights / inputs from (arch, dims, seed)."""
import glob
import json
import os

import numpy as np

from simpleimagecaptionzoo_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names(kind=None):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    refiner = kind == "aoaref"
    names = [n for n in names if n.startswith("aoaref") == refiner]  # aoaref_*: full-captioner cases (rebuild_refiner)
    if kind and not refiner:
        names = [n for n in names if kind in n]
    return names


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def rebuild_refiner(meta):
    """Full AoADetection checkpoint (decoder + img_feats_porjection + aoa_refine), bottom-up features and mask of an
    ``aoaref_*`` case, exactly as make_golden.run_refiner_case built them."""
    dims = meta["dims"]
    sd = synth.make_state_dict("AOA", seed=meta["seed"], chaotic=meta["chaotic"], end_boost=meta.get("end_boost", 0.0), **dims)
    sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=meta["enc_dim"], seed=meta["seed"],
                                            chaotic=meta["chaotic"]))
    B, R, seed = meta["B"], meta["R"], meta["seed"]
    bu = synth.make_region_feats(B, R, meta["enc_dim"], seed)
    mask = None
    if meta["masked"]:
        mask = synth.make_region_mask(B, R, max(1, R // 3), seed)
        bu = bu * mask[:, :, None]
    return sd, bu, mask


def rebuild(meta):
    """state_dict (numpy), feats, mask for a golden case, exactly as make_golden.py built them."""
    arch, dims = meta["arch"], meta["dims"]
    sd = synth.make_state_dict(arch, seed=meta["seed"], chaotic=meta["chaotic"],
                               end_boost=meta.get("end_boost", 0.0), **dims)
    B, R, seed = meta["B"], meta["R"], meta["seed"]
    mask = None
    if arch == "BUTD":
        feats = synth.make_region_feats(B, R, dims["enc_dim"], seed)
    elif arch == "NIC":
        feats = synth.make_image_embed(B, dims["embed_dim"], seed)
    else:
        feats = synth.make_refined_feats(B, R, dims["hidden_dim"], seed)
        if meta["masked"]:
            mask = synth.make_region_mask(B, R, max(1, R // 3), seed)
            feats = feats * mask[:, :, None]
    return sd, feats, mask
