"""Pin the numpy oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import capdec_oracle as orc
from tests.golden_util import case_names, load_case, rebuild

TINY = case_names("tiny")
FULL = case_names("full")


def _decoder(meta):
    sd, feats, mask = rebuild(meta)
    dec = orc.make_decoder(meta["arch"], sd, num_heads=meta["dims"].get("num_heads", 8))
    dec.prepare(feats, mask)
    return dec


def _check_beam(meta, gold, res, tol_tie=1e-4, tol_score=1e-3):
    verdict = orc.agreement(res.tokens, gold["tokens"], res.min_gap, tol=tol_tie)
    assert "diff" not in verdict, [i for i, v in enumerate(verdict) if v == "diff"]
    exact = np.array([v == "exact" for v in verdict])
    # reference quirk kept: the best COMPLETED hypothesis wins; output is float32 then (BUTD_Model.py:309)
    assert np.array_equal(res.completed[exact], gold["out_is_float"][exact])
    assert np.array_equal(res.lengths[exact], gold["lengths"][exact])
    # sequence log-prob: ours is the running beam score, the golden is the reference's teacher-forced forward
    assert np.allclose(res.scores[exact], gold["scores"][exact], atol=tol_score, rtol=1e-5)
    return exact.mean()


@pytest.mark.parametrize("name", TINY)
def test_tiny_beam_reference_form(name):
    meta, gold = load_case(name)
    dec = _decoder(meta)
    res = orc.beam_search_reference_form(dec, meta["K"], meta["T"])
    frac = _check_beam(meta, gold, res)
    assert frac >= 0.95


@pytest.mark.parametrize("name", TINY + FULL)
def test_beam_batched_matches_golden(name):
    meta, gold = load_case(name)
    dec = _decoder(meta)
    res = orc.beam_search_batched(dec, meta["K"], meta["T"])
    frac = _check_beam(meta, gold, res)
    assert frac >= 0.95


@pytest.mark.parametrize("name", TINY)
def test_batched_equals_reference_form(name):
    meta, _ = load_case(name)
    dec = _decoder(meta)
    a = orc.beam_search_reference_form(dec, meta["K"], meta["T"])
    b = orc.beam_search_batched(dec, meta["K"], meta["T"])
    same = (a.tokens == b.tokens).all(1)
    # the two forms batch the matmuls differently; only fp32 re-association near-ties may differ
    assert same.mean() >= 0.95
    assert np.allclose(a.scores[same], b.scores[same], atol=1e-4)
    assert np.array_equal(a.completed[same], b.completed[same])


@pytest.mark.parametrize("name", TINY + FULL)
def test_greedy_matches_golden(name):
    meta, gold = load_case(name)
    dec = _decoder(meta)
    ids, gaps, _ = orc.greedy_sample(dec, meta["T"])
    for b in range(meta["B"]):
        if not np.array_equal(ids[b], gold["greedy"][b]):
            t = int(np.argmax(ids[b] != gold["greedy"][b]))
            assert gaps[b, t] < 1e-4, (b, t, gaps[b, t])


@pytest.mark.parametrize("name", TINY + FULL)
def test_sampling_matches_golden(name):
    meta, gold = load_case(name)
    dec = _decoder(meta)
    n = meta["n_samples"]
    seq, lps, gaps = orc.multinomial_sample(dec, meta["T"], n, meta["sample_seed"])
    same = (seq == gold["sample_seq"]).all(-1)
    assert same.mean() >= 0.9, same.mean()
    assert np.allclose(lps[same], gold["sample_logprobs"][same], atol=1e-3)
    for b, j in zip(*np.nonzero(~same)):
        t = int(np.argmax(seq[b, j] != gold["sample_seq"][b, j]))
        assert gaps[b, j, t] < 1e-3, (b, j, t, gaps[b, j, t])


def test_gumbel_noise_is_gumbel():
    g = orc.gumbel_noise(7, np.arange(64), 3, 4096).ravel()
    assert abs(g.mean() - 0.5772) < 0.01 and abs(g.var() - np.pi ** 2 / 6) < 0.03
    # counter based: independent of batch composition
    g2 = orc.gumbel_noise(7, np.array([5]), 3, 4096)
    assert np.array_equal(g2[0], orc.gumbel_noise(7, np.arange(64), 3, 4096)[5])


def test_weight_norm_fold():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((7, 5)).astype(np.float32)
    g = rng.standard_normal((7, 1)).astype(np.float32)
    w = orc.fold_weight_norm(g, v)
    assert np.allclose(np.linalg.norm(w, axis=1), np.abs(g[:, 0]), rtol=1e-5)


# ---------------------------------------------------------------------------------------------------
# AoA encoder side (img_feats_porjection + aoa_refine) in front of the decoder: SURVEY.md section 8f row 1
# ---------------------------------------------------------------------------------------------------
from tests.golden_util import rebuild_refiner  # noqa: E402

REFINER = case_names("aoaref")


@pytest.mark.parametrize("name", REFINER)
def test_refiner_matches_golden(name):
    """The oracle's projection + 6-layer AoA refiner against the reference's refined features (fp32 re-association
    through six residual layers: 2e-4 absolute on LayerNorm-ed, O(1) outputs)."""
    meta, gold = load_case(name)
    sd, bu, mask = rebuild_refiner(meta)
    ref = orc.aoa_project_refine(sd, bu, mask, num_heads=meta["dims"]["num_heads"])
    got = ref[:, :, ::meta["store_stride"]]
    valid = np.ones(got.shape[:2], bool) if mask is None else mask.astype(bool)
    assert np.abs(got - gold["refined"])[valid].max() < 2e-4
    # padded regions: the reference runs them through the layers too (they only see the valid keys); same here
    assert np.abs(got - gold["refined"]).max() < 2e-4


@pytest.mark.parametrize("name", REFINER)
def test_refiner_then_decode_matches_golden(name):
    """Whole AoADetection path from bottom-up features: batched masked decode == the reference's one-image-per-call
    beam search on the image's own regions, and the reference's batched greedy ``sampler``."""
    meta, gold = load_case(name)
    sd, bu, mask = rebuild_refiner(meta)
    dec = orc.make_decoder("AOA", sd, num_heads=meta["dims"]["num_heads"])
    dec.prepare(orc.aoa_project_refine(sd, bu, mask, num_heads=meta["dims"]["num_heads"]), mask)
    res = orc.beam_search_batched(dec, meta["K"], meta["T"])
    verdict = orc.agreement(res.tokens, gold["tokens"], res.min_gap, tol=1e-4)
    assert "diff" not in verdict, verdict
    assert np.mean([v == "exact" for v in verdict]) >= 0.9
    ids, gaps, _ = orc.greedy_sample(dec, meta["T"])
    for b in range(meta["B"]):
        if not np.array_equal(ids[b], gold["greedy"][b]):
            t = int(np.argmax(ids[b] != gold["greedy"][b]))
            assert gaps[b, t] < 1e-4, (b, t, gaps[b, t])


@pytest.mark.parametrize("name,images,lo", [("butd", 5000, 1025), ("aoa_bu", 1000, 450)])
def test_oracle_reproduces_the_reference_captions_of_the_agreement_sets(name, images, lo):
    """A slice of the agreement sets (captions decoded by the REFERENCE's own code, tests/golden/make_agreement_set.py),
    chosen to contain the images where oracle and reference differ (1038; 460): the oracle returns the reference's caption,
    or the first difference is preceded / followed by a sub-1e-4 gap (an arbitrary choice between tied beams)."""
    from tests import agreement_util as au
    if not os.path.exists(au.set_path(name, images)):
        pytest.skip("agreement set not generated")
    meta, gold = au.load_set(name, images)
    n = 16
    chunk_lo = (lo // meta["chunk"]) * meta["chunk"]
    f = au.feats_for(name, chunk_lo, meta["chunk"], meta["regions"])[lo - chunk_lo:lo - chunk_lo + n]
    from simpleimagecaptionzoo_b200 import synth
    dims = synth.DIMS[au.SETS[name]]
    sd = synth.make_state_dict(au.SETS[name], seed=0, **dims)
    if name == "aoa_bu":
        sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
        f = orc.aoa_project_refine(sd, f)
    o = orc.make_decoder(au.SETS[name], sd)
    o.prepare(f)
    res = orc.beam_search_batched(o, meta["beam"], meta["max_seq"])
    want = gold["tokens"][lo:lo + n].astype(np.int32)
    same = (res.tokens == want).all(1)
    assert np.array_equal(same, gold["oracle_equal"][lo:lo + n])
    assert same.sum() >= n - 1
    for b in np.nonzero(~same)[0]:
        assert int(gold["tie_bits"][lo + b]) != 0  # a sub-1e-4 gap somewhere in the decode
    # the stored per-step gaps are this oracle's
    assert np.allclose(np.minimum(res.min_gap, 1e3).astype(np.float16).astype(np.float32), gold["gaps"][lo:lo + n].astype(np.float32), atol=2e-3)
