#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REFERENCE's own code.

Runs only in the build container (needs /root/reference); the GPU box and the test-suite use
the committed ``*.npz`` files.  Usage:  python tests/golden/make_golden.py

The reference model files are imported as source text from /root/reference/Models, never
copied into this repo, with the minimal shims of SURVEY.md section 8c applied textually:

1. ``prev_word_inds = top_k_words / self.vocab_size`` -> ``//``  (true division on torch>=1.5
   yields a float index and an IndexError; floor division is the torch<=1.4 meaning).
2. ``max_step_limit = 50`` -> read from ``self.max_step_limit`` (north_star decodes max_seq=20).
3. ``torch.multinomial`` is replaced, for the ``sample_rl`` vectors only, by the Gumbel-max draw
   with the counter-based noise of ``oracle.capdec_oracle.gumbel_noise`` (same distribution,
   reproducible stream) -- the reference's own RNG stream cannot be shared with a CUDA kernel.

Weights and inputs come from ``simpleimagecaptionzoo_b200.synth`` (numpy PCG64), loaded into
the reference modules with ``load_state_dict`` under their own key names, so every vector can
be regenerated from (arch, dims, seed) without storing tensors.
Each case stores the reference's tokens, the reference's teacher-forced sequence log-prob
(its own ``forward``), greedy ids, and the shimmed ``sample_rl`` outputs.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/Models"

from simpleimagecaptionzoo_b200 import synth  # noqa: E402
from oracle import capdec_oracle as orc  # noqa: E402

_STATE = {"seed": 0, "t": 0}


def load_reference(name):
    path = os.path.join(REF, name + ".py")
    src = open(path).read()
    n1 = src.count("prev_word_inds = top_k_words / self.vocab_size")
    src = src.replace("prev_word_inds = top_k_words / self.vocab_size",
                      "prev_word_inds = top_k_words // self.vocab_size")
    n2 = src.count("max_step_limit = 50")
    src = src.replace("max_step_limit = 50", "max_step_limit = getattr(self, 'max_step_limit', 50)")
    assert n1 == 1 and n2 == 1, (name, n1, n2)
    mod = types.ModuleType("ref_" + name)
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def _gumbel_multinomial(prob, num_samples=1, **kw):
    assert num_samples == 1
    rows = np.arange(prob.shape[0])
    g = orc.gumbel_noise(_STATE["seed"], rows, _STATE["t"], prob.shape[1])
    _STATE["t"] += 1
    it = (torch.log(prob) + torch.from_numpy(g)).argmax(dim=1, keepdim=True)
    return it


def t_sd(sd, strip="decoder."):
    return {k[len(strip):]: torch.from_numpy(v.copy()) for k, v in sd.items() if k.startswith(strip)}


def build_decoder(arch, dims, sd):
    if arch == "BUTD":
        m = load_reference("BUTD_Model")
        dec = m.DecoderRNN(atten_dim=dims["atten_dim"], embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"],
                           vocab_size=dims["vocab_size"], enc_dim=dims["enc_dim"])
    elif arch == "NIC":
        m = load_reference("NIC_Model")
        dec = m.DecoderRNN(embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"], vocab_size=dims["vocab_size"])
    else:
        m = load_reference("AoA_Model")
        dec = m.AoA_Decoder(hidden_dim=dims["hidden_dim"], num_heads=dims["num_heads"], embed_dim=dims["embed_dim"],
                            vocab_size=dims["vocab_size"], d_model=dims["hidden_dim"])
    missing = dec.load_state_dict(t_sd(sd), strict=True)
    dec.eval()
    return dec


def make_inputs(arch, dims, B, R, seed, masked):
    if arch == "BUTD":
        return synth.make_region_feats(B, R, dims["enc_dim"], seed), None
    if arch == "NIC":
        return synth.make_image_embed(B, dims["embed_dim"], seed), None
    feats = synth.make_refined_feats(B, R, dims["hidden_dim"], seed)
    mask = synth.make_region_mask(B, R, max(1, R // 3), seed) if masked else None
    if mask is not None:
        feats = feats * mask[:, :, None]  # padded regions are zero (AoA_Engine.py:36-41)
    return feats, mask


def ref_seq_logprob(arch, dec, feats_b, mask_b, seq):
    """Reference's own teacher-forced forward -> sum of log-probs of seq[1:]."""
    cap = torch.tensor([seq], dtype=torch.long)
    lengths = [len(seq) - 1]
    with torch.no_grad():
        if arch == "BUTD":
            packed, _ = dec(feats_b, cap, lengths)
        elif arch == "NIC":
            packed = dec(feats_b, cap, lengths)
        else:
            packed = dec(feats_b, cap, lengths, mask_b)
    logits = packed.data  # (L-1, V), batch of one -> time-major == sequence order
    lp = torch.log_softmax(logits, dim=1)
    return float(lp[torch.arange(len(seq) - 1), cap[0, 1:]].sum())


def run_case(name, arch, dims_key, B, R, K, T, seed, chaotic, masked=False, end_boost=0.0, n_samples=2):
    dims = dict((synth.TINY_DIMS if dims_key == "tiny" else synth.DIMS)[arch])
    sd = synth.make_state_dict(arch, seed=seed, chaotic=chaotic, end_boost=end_boost, **dims)
    dec = build_decoder(arch, dims, sd)
    dec.max_step_limit = T
    feats, mask = make_inputs(arch, dims, B, R, seed, masked)
    tf = torch.from_numpy(feats)
    tm = None if mask is None else torch.from_numpy(mask)

    tokens = np.zeros((B, 1 + T), np.int32)
    lengths = np.zeros(B, np.int32)
    scores = np.zeros(B, np.float32)
    out_is_float = np.zeros(B, bool)
    with torch.no_grad():
        for b in range(B):  # the reference decodes one image per call (Utils.py:72-73)
            fb = tf[b:b + 1]
            mb = None if tm is None else tm[b:b + 1]
            if arch == "BUTD":
                seq_t, _ = dec.beam_search_sample(fb, beam_size=K)
            elif arch == "NIC":
                seq_t = dec.beam_search_sample(fb, beam_size=K)
            else:
                seq_t, _ = dec.beam_search_sample(fb, beam_size=K, bu_masks=mb)
            out_is_float[b] = seq_t.dtype == torch.float32  # BUTD_Model.py:309 vs :314
            seq = [int(x) for x in seq_t[0].tolist()]
            tokens[b, :len(seq)] = seq
            lengths[b] = len(seq)
            scores[b] = ref_seq_logprob(arch, dec, fb, mb, seq)
        # greedy (BUTD_Model.py:153-189)
        if arch == "BUTD":
            gids, _ = dec.sample(tf, max_len=T)
        elif arch == "NIC":
            gids = dec.sample(tf, max_len=T)
        else:
            gids, _ = dec.sample(tf, max_len=T, bu_masks=tm)
        # sampling rollout with the shimmed multinomial, n_samples rows per image
        rep = lambda a: None if a is None else a.repeat_interleave(n_samples, dim=0)
        _STATE["seed"], _STATE["t"] = 1234 + seed, 0
        real = torch.multinomial
        torch.multinomial = _gumbel_multinomial
        try:
            if arch == "AOA":
                sseq, slp = dec.sample_rl(rep(tf), max_len=T, bu_masks=rep(tm))
            else:
                sseq, slp = dec.sample_rl(rep(tf), max_len=T)
        finally:
            torch.multinomial = real
    out = dict(
        tokens=tokens, lengths=lengths, scores=scores, out_is_float=out_is_float,
        greedy=gids.numpy().astype(np.int32),
        sample_seq=sseq.numpy().astype(np.int32).reshape(B, n_samples, T),
        sample_logprobs=slp.numpy().astype(np.float32).reshape(B, n_samples, T),
        meta=np.array(json.dumps(dict(name=name, arch=arch, dims_key=dims_key, dims=dims, B=B, R=R, K=K, T=T,
                                      seed=seed, chaotic=chaotic, masked=masked, end_boost=end_boost, n_samples=n_samples,
                                      sample_seed=1234 + seed, torch=torch.__version__))),
    )
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    ncomp = int((tokens == orc.END).any(1).sum())
    print(f"{name}: B={B} completed={ncomp} lens {lengths.min()}..{lengths.max()} -> {os.path.relpath(path, ROOT)}")


CASES = [
    # name, arch, dims, B, R, K, T, seed, chaotic, masked, end_boost
    ("butd_tiny_k3", "BUTD", "tiny", 64, 6, 3, 20, 1, True, False, 1.0),
    ("butd_tiny_k5", "BUTD", "tiny", 32, 9, 5, 20, 3, True, False, 1.5),
    ("butd_tiny_k1", "BUTD", "tiny", 16, 6, 1, 12, 2, True, False, 1.0),
    ("nic_tiny_k3", "NIC", "tiny", 64, 0, 3, 20, 3, True, False, 1.0),
    ("nic_tiny_k5", "NIC", "tiny", 32, 0, 5, 20, 3, True, False, 1.0),
    # AoA: U(-1,1) weights amplify fp32 re-association noise to O(1) within ~8 steps (LayerNorm
    # on a saturated LSTM), so the reference disagrees with ITSELF between batch sizes; U(-0.3,0.3) is well conditioned
    ("aoa_tiny_k3", "AOA", "tiny", 64, 9, 3, 20, 0, 0.3, False, 0.8),
    ("aoa_tiny_k3_masked", "AOA", "tiny", 64, 9, 3, 20, 1, 0.3, True, 0.6),
    ("butd_full_k3", "BUTD", "full", 16, 36, 3, 20, 0, False, False),   # BASELINE configs[0]
    ("butd_full_k5_r196", "BUTD", "full", 4, 196, 5, 20, 1, False, False),  # configs[2] shape
    ("nic_full_k3", "NIC", "full", 16, 0, 3, 20, 0, False, False),      # configs[1] decoder
    ("aoa_full_k3", "AOA", "full", 8, 36, 3, 20, 0, False, False),      # configs[3] decoder
]


def run_refiner_case(name, dims_key, B, R, K, T, seed, chaotic, masked, end_boost=0.0, store_stride=1):
    """AoADetection_Captioner end to end from the bottom-up features (AoA_Model.py:657-753): the refined features of
    ``aoa_refine(pack_wrapper(img_feats_porjection, bu_feats, bu_masks), bu_masks)``, greedy ids of the batched
    ``sampler`` and beam tokens of ``beam_search_sampler``.  Beam search is called the way the reference's evaluation
    does: one image per call, its features cut to its own number of regions and no mask
    (AoA_Engine.modify_visual_inputs builds the mask per batch and drops it when every region is valid)."""
    dims = dict((synth.TINY_DIMS if dims_key == "tiny" else synth.DIMS)["AOA"])
    sd = synth.make_state_dict("AOA", seed=seed, chaotic=chaotic, end_boost=end_boost, **dims)
    sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=seed, chaotic=chaotic))
    m = load_reference("AoA_Model")
    cap = m.AoADetection_Captioner(vocab_size=dims["vocab_size"], num_heads=dims["num_heads"], hidden_dim=dims["hidden_dim"],
                                   embed_dim=dims["embed_dim"])
    cap.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()}, strict=True)
    cap.eval()
    cap.decoder.max_step_limit = T
    bu = synth.make_region_feats(B, R, 2048, seed)
    mask = synth.make_region_mask(B, R, max(1, R // 3), seed) if masked else None
    if mask is not None:
        bu = bu * mask[:, :, None]  # AoA_Engine.py:36-41: padded regions are zero
    tb = torch.from_numpy(bu)
    tm = None if mask is None else torch.from_numpy(mask)
    tokens = np.zeros((B, 1 + T), np.int32)
    lengths = np.zeros(B, np.int32)
    with torch.no_grad():
        refined = cap.aoa_refine(x=m.pack_wrapper(module=cap.img_feats_porjection, bu_feats=tb, bu_masks=tm), bu_mask=tm)
        gids = cap.sampler({"bu_feats": tb, "bu_masks": tm}, max_len=T)
        for b in range(B):
            n = R if mask is None else int(mask[b].sum())
            seq_t = cap.beam_search_sampler({"bu_feats": tb[b:b + 1, :n], "bu_masks": None}, beam_size=K)
            seq = [int(x) for x in seq_t[0].tolist()]
            tokens[b, :len(seq)] = seq
            lengths[b] = len(seq)
    out = dict(
        tokens=tokens, lengths=lengths, greedy=gids.numpy().astype(np.int32),
        refined=refined.numpy().astype(np.float32)[:, :, ::store_stride],
        meta=np.array(json.dumps(dict(name=name, arch="AOA", dims_key=dims_key, dims=dims, B=B, R=R, K=K, T=T, seed=seed,
                                      chaotic=chaotic, masked=masked, end_boost=end_boost, refiner=True, enc_dim=2048,
                                      store_stride=store_stride, torch=torch.__version__))),
    )
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    ncomp = int((tokens == orc.END).any(1).sum())
    print(f"{name}: B={B} completed={ncomp} lens {lengths.min()}..{lengths.max()} -> {os.path.relpath(path, ROOT)}")


REFINER_CASES = [
    # name, dims, B, R, K, T, seed, chaotic, masked, end_boost, store_stride
    ("aoaref_tiny_k3", "tiny", 16, 9, 3, 20, 4, 0.3, False, 0.6, 1),
    ("aoaref_tiny_k3_masked", "tiny", 16, 9, 3, 20, 5, 0.3, True, 0.9, 1),
    ("aoaref_full_k3", "full", 4, 36, 3, 20, 0, False, False, 0.0, 8),     # BASELINE configs[3] from bu_feats
    ("aoaref_full_k3_masked", "full", 4, 36, 3, 20, 1, False, True, 0.0, 8),
]

if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    for c in CASES:
        if only and c[0] not in only:
            continue
        run_case(*c)
    for c in REFINER_CASES:
        if only and c[0] not in only:
            continue
        run_refiner_case(*c)
