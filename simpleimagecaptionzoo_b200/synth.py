"""Synthetic weights and inputs for the caption decoders (numpy only, no torch RNG).

The reference ships no checkpoints, so parity fixtures, GPU tests and
bench.py all decode with random-init weights.  Weights are drawn here with numpy's PCG64 so
they are bit-identical in the build container (where the reference is imported to produce
the golden vectors) and on the GPU box (where the reference does not exist).

The dictionaries returned use the reference's own ``state_dict`` key names, including the
legacy weight-norm naming (``weight_g`` (out,1) / ``weight_v`` (out,in)), so they can be fed
to ``load_state_dict`` of the reference captioners unchanged:

* BUTD  - Models/BUTD_Model.py:40-95   (SoftAttention, DecoderRNN.__init__/init_weights)
* NIC   - Models/NIC_Model.py:39-50    (DecoderRNN.__init__)
* AoA   - Models/AoA_Model.py:71-89,197-221 (AoABlock, AoA_Decoder.__init__/init_weights)

Distributions follow PyTorch's defaults for the same modules (Linear / LSTMCell:
U(-1/sqrt(fan), 1/sqrt(fan)); Embedding N(0,1) unless the reference overrides it with
U(-0.1,0.1)), which is what "random-init weights of each named architecture" means in
BASELINE.json.  ``chaotic=s`` (s=1.0 or True is the SURVEY.md section 4 recipe: all parameters
U(-s,s), predict.weight_g doubled), together with ``end_boost`` on the ``<end>`` bias, makes small
models emit ``<end>`` at varied lengths so that the beam bookkeeping is exercised.
"""
from __future__ import annotations

import numpy as np

ARCHS = ("NIC", "BUTD", "AOA")

PAD, STA, END, UNK = 0, 1, 2, 3


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _u(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _linear(rng, sd, name, out_f, in_f, chaotic, weight_norm=False, w_bound=None, zero_bias=False):
    bound = float(chaotic) if chaotic else (w_bound if w_bound is not None else 1.0 / np.sqrt(in_f))
    w = _u(rng, (out_f, in_f), bound)
    b_bound = float(chaotic) if chaotic else 1.0 / np.sqrt(in_f)
    b = np.zeros(out_f, np.float32) if (zero_bias and not chaotic) else _u(rng, (out_f,), b_bound)
    if weight_norm:
        # torch.nn.utils.weight_norm at construction: v = w, g = ||w||_row  (legacy naming)
        sd[name + ".weight_v"] = w
        if chaotic:
            sd[name + ".weight_g"] = _u(rng, (out_f, 1), float(chaotic))
        else:
            sd[name + ".weight_g"] = np.sqrt((w.astype(np.float64) ** 2).sum(1, keepdims=True)).astype(np.float32)
    else:
        sd[name + ".weight"] = w
    sd[name + ".bias"] = b


def _lstm(rng, sd, name, in_f, hid, chaotic):
    bound = float(chaotic) if chaotic else 1.0 / np.sqrt(hid)
    sd[name + ".weight_ih"] = _u(rng, (4 * hid, in_f), bound)
    sd[name + ".weight_hh"] = _u(rng, (4 * hid, hid), bound)
    sd[name + ".bias_ih"] = _u(rng, (4 * hid,), bound)
    sd[name + ".bias_hh"] = _u(rng, (4 * hid,), bound)


def make_state_dict(arch: str, *, vocab_size: int, hidden_dim: int, embed_dim: int,
                    atten_dim: int = 0, enc_dim: int = 2048, num_heads: int = 8,
                    seed: int = 0, chaotic: float = 0.0, end_boost: float = 0.0) -> dict:
    """Decoder ``state_dict`` (numpy fp32) with the reference's key names, prefix ``decoder.``."""
    arch = arch.upper()
    assert arch in ARCHS
    rng = _rng(seed)
    sd: dict = {}
    V, H, E = vocab_size, hidden_dim, embed_dim
    if arch == "BUTD":
        A, D = atten_dim, enc_dim
        _linear(rng, sd, "decoder.atten.enc_att", A, D, chaotic, weight_norm=True)
        _linear(rng, sd, "decoder.atten.dec_att", A, H, chaotic, weight_norm=True)
        _linear(rng, sd, "decoder.atten.affine", 1, A, chaotic, weight_norm=True)
        sd["decoder.embed.0.weight"] = _u(rng, (V, E), float(chaotic) if chaotic else 0.1)
        _lstm(rng, sd, "decoder.TD_atten", E + D + H, H, chaotic)
        _lstm(rng, sd, "decoder.language_model", D + H, H, chaotic)
        _linear(rng, sd, "decoder.predict", V, H, chaotic, weight_norm=True, w_bound=0.1, zero_bias=True)
    elif arch == "NIC":
        if chaotic:
            sd["decoder.embed.weight"] = _u(rng, (V, E), float(chaotic))
        else:
            sd["decoder.embed.weight"] = rng.standard_normal((V, E)).astype(np.float32)
        _lstm(rng, sd, "decoder.lstm", E, H, chaotic)
        _linear(rng, sd, "decoder.predict", V, H, chaotic, weight_norm=True)
    else:  # AOA
        assert E > 0 and H % num_heads == 0
        _lstm(rng, sd, "decoder.lstm", E + H, H, chaotic)
        for nm in ("linear_Q", "linear_K", "linear_V"):
            _linear(rng, sd, "decoder.aoa_block." + nm, H, H, chaotic)
        _linear(rng, sd, "decoder.aoa_block.aoa_module.0", 2 * H, 2 * H, chaotic)
        sd["decoder.embed.0.weight"] = _u(rng, (V, E), float(chaotic) if chaotic else 0.1)
        if chaotic:
            sd["decoder.h_norm.gain"] = _u(rng, (H,), 1.0)
            sd["decoder.h_norm.bias"] = _u(rng, (H,), 1.0)
        else:
            sd["decoder.h_norm.gain"] = np.ones(H, np.float32)
            sd["decoder.h_norm.bias"] = np.zeros(H, np.float32)
        _linear(rng, sd, "decoder.predict", V, H, chaotic, weight_norm=True, w_bound=0.1, zero_bias=True)
    if chaotic:
        sd["decoder.predict.weight_g"] = sd["decoder.predict.weight_g"] * 2.0
    if end_boost:
        # raise the <end> logit so that small models finish at varied lengths (test recipe only)
        sd["decoder.predict.bias"] = sd["decoder.predict.bias"].copy()
        sd["decoder.predict.bias"][END] += np.float32(end_boost)
    return sd


def make_refiner_state_dict(*, hidden_dim: int, enc_dim: int = 2048, num_layers: int = 6, seed: int = 0,
                            chaotic: float = 0.0) -> dict:
    """Encoder-side ``state_dict`` entries of ``AoADetection_Captioner`` / ``AoASpatial_Captioner`` (numpy fp32, the
    reference's key names): ``img_feats_porjection`` (Models/AoA_Model.py:661-665, Linear 2048->H + ReLU) and the
    ``aoa_refine`` stack (:122-162: 6 x [pre-LayerNorm, AoABlock self-attention, residual] + final LayerNorm).
    PyTorch default Linear init; LayerNorm gain 1 / bias 0 (:18-19) unless ``chaotic``."""
    rng = _rng(7_000_003 + seed)
    sd: dict = {}
    H = hidden_dim
    _linear(rng, sd, "img_feats_porjection.0", H, enc_dim, chaotic)

    def norm(name):
        if chaotic:
            sd[name + ".gain"] = (1.0 + _u(rng, (H,), 0.5)).astype(np.float32)
            sd[name + ".bias"] = _u(rng, (H,), 0.5)
        else:
            sd[name + ".gain"] = np.ones(H, np.float32)
            sd[name + ".bias"] = np.zeros(H, np.float32)

    for i in range(num_layers):
        p = f"aoa_refine.aoa_layers.{i}."
        for nm in ("linear_Q", "linear_K", "linear_V"):
            _linear(rng, sd, p + "aoa_block." + nm, H, H, chaotic)
        _linear(rng, sd, p + "aoa_block.aoa_module.0", 2 * H, 2 * H, chaotic)
        norm(p + "sublayer.norm")
    norm("aoa_refine.norm")
    return sd


def make_region_feats(batch: int, regions: int, dim: int, seed: int = 0) -> np.ndarray:
    """|N(0,1)| region features, (B,R,D) fp32: post-ReLU-like bottom-up / CNN-grid features."""
    return np.abs(_rng(1_000_003 + seed).standard_normal((batch, regions, dim))).astype(np.float32)


def make_refined_feats(batch: int, regions: int, dim: int, seed: int = 0) -> np.ndarray:
    """N(0,1) features, (B,R,H) fp32: stand-in for the LayerNorm-ed AoA refiner output."""
    return _rng(2_000_003 + seed).standard_normal((batch, regions, dim)).astype(np.float32)


def make_image_embed(batch: int, dim: int, seed: int = 0) -> np.ndarray:
    """N(0,1)*0.5 image embeddings, (B,E) fp32: stand-in for NIC's EncoderCNN output."""
    return (0.5 * _rng(3_000_003 + seed).standard_normal((batch, dim))).astype(np.float32)


def make_region_mask(batch: int, regions: int, min_regions: int, seed: int = 0) -> np.ndarray:
    """Float {0,1} mask (B,R) with a prefix of ones per image, like AoA 'adaptive' features
    (ModelEngines/AoA_Engine.py:35-44).  At least one image keeps all regions."""
    rng = _rng(4_000_003 + seed)
    lens = rng.integers(min_regions, regions + 1, size=batch)
    lens[0] = regions
    return (np.arange(regions)[None, :] < lens[:, None]).astype(np.float32)


def make_uniforms(batch: int, steps: int, seed: int = 0) -> np.ndarray:
    return _rng(5_000_003 + seed).random((batch, steps), dtype=np.float32)


def make_caption_corpus(n_images: int = 60, vocab_size: int = 64, refs_per_image: int = 5, seed: int = 0):
    """Synthetic tokenised caption corpus for the CIDEr-D reward tests: ``ix2word`` (ids 0-3 = <pad>,<sta>,<end>,<unk> as
    PreProcess/Build_caption_vocab.py:37-40 assigns them) and, per image, ``refs_per_image`` (+0..2) reference strings of
    5-16 Zipf-distributed words; ~3 % of the reference words are outside the vocabulary."""
    rng = _rng(8_000_003 + seed)
    ix2word = ["<pad>", "<sta>", "<end>", "<unk>"] + [f"w{i}" for i in range(4, vocab_size)]
    p = 1.0 / np.arange(1, vocab_size - 3)
    p /= p.sum()
    refs = []
    for _ in range(n_images):
        topic = rng.choice(np.arange(4, vocab_size), size=6, p=p)  # words an image's captions tend to share
        caps = []
        for _ in range(refs_per_image + int(rng.integers(0, 3))):
            L = int(rng.integers(5, 17))
            ids = np.where(rng.random(L) < 0.5, rng.choice(topic, size=L), rng.choice(np.arange(4, vocab_size), size=L, p=p))
            words = [ix2word[i] for i in ids]
            for j in range(L):
                if rng.random() < 0.03:
                    words[j] = f"oov{int(rng.integers(0, 20))}"
            caps.append(" ".join(words))
        refs.append(caps)
    return ix2word, refs


def make_rollouts(ix2word, refs, img_index, n_per_image: int, max_len: int = 20, seed: int = 0):
    """Synthetic SCST rollouts for images ``img_index`` of a corpus: ``gen`` (B*n, T) in ``sample_rl``'s storage (<end> and
    everything after it = 0) and ``greedy`` (B, T) in ``sample``'s (an <end>=2 somewhere, words keep coming after it).
    Rows are noisy copies of the image's references so that scores are non-trivial; a few rows hit the edge cases (empty
    caption, full-length caption without <end>)."""
    rng = _rng(9_000_003 + seed)
    word2ix = {w: i for i, w in enumerate(ix2word)}
    V, T = len(ix2word), max_len

    def noisy(img):
        ref = refs[img][int(rng.integers(0, len(refs[img])))].split()
        ids = [word2ix.get(w, UNK) for w in ref][:int(rng.integers(1, T))]
        return [int(rng.integers(4, V)) if rng.random() < 0.25 else i for i in ids]

    B = len(img_index)
    gen = np.zeros((B * n_per_image, T), np.int32)
    greedy = np.zeros((B, T), np.int32)
    for b, img in enumerate(img_index):
        for j in range(n_per_image):
            ids = noisy(img)
            r = rng.random()
            if r < 0.05:
                ids = []                                  # <end> drawn first: stored as all zeros
            elif r < 0.10:
                ids = [int(x) for x in rng.integers(4, V, size=T)]  # never finished
            gen[b * n_per_image + j, :len(ids)] = ids
        ids = noisy(img)
        if rng.random() < 0.05:
            ids = []
        row = [int(x) for x in rng.integers(4, V, size=T)]   # greedy keeps generating after <end> (BUTD_Model.py:153-189)
        row[:len(ids)] = ids
        if len(ids) < T and rng.random() < 0.9:
            row[len(ids)] = END
        greedy[b] = row
    return gen, greedy


# Named dimension sets ------------------------------------------------------------------------------------

DIMS = {
    # Configs/Models/BUTDDetection.json:4-6, V from BASELINE.json configs[0]
    "BUTD": dict(vocab_size=9487, hidden_dim=1024, embed_dim=1024, atten_dim=1024, enc_dim=2048),
    # Configs/Models/NIC.json:3-4
    "NIC": dict(vocab_size=9487, hidden_dim=512, embed_dim=512),
    # Configs/Models/AoADetection.json:3-4, Models/AoA_Model.py:658
    "AOA": dict(vocab_size=9487, hidden_dim=1024, embed_dim=1024, num_heads=8),
}

TINY_DIMS = {
    "BUTD": dict(vocab_size=32, hidden_dim=64, embed_dim=64, atten_dim=64, enc_dim=128),
    "NIC": dict(vocab_size=32, hidden_dim=64, embed_dim=64),
    "AOA": dict(vocab_size=32, hidden_dim=64, embed_dim=64, num_heads=8),
}
