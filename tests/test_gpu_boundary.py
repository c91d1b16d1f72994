"""GPU tests of the drop-in boundary proper: ``engine.install()`` on a live (stand-in) reference Engine, the captioner's
``sampler_rl`` wrapper, the captured-decode cache, and the direct agreement with the reference's captions on the committed
agreement sets.  Everything goes through the C ABI."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import capdec_oracle as orc  # noqa: E402
from tests import agreement_util as au  # noqa: E402
from tests.golden_util import load_case, rebuild  # noqa: E402

pytestmark = pytest.mark.gpu


class _Vocab:
    def __init__(self, n):
        self.ix2word = {0: "<pad>", 1: "<sta>", 2: "<end>", 3: "<unk>", **{i: f"w{i}" for i in range(4, n)}}

    def __len__(self):
        return len(self.ix2word)


class _RefModel(torch.nn.Module):
    """Stands in for the reference's ``BUTDDetection_Captioner``: the same ``state_dict()`` a loaded checkpoint gives
    (Engine.py:43-70) and the four methods ``install`` may rebind (here they only record that they were called)."""

    def __init__(self, sd):
        super().__init__()
        self._sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
        self.called = []

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def sampler(self, visual_inputs, max_len=20):
        self.called.append("sampler")

    def sampler_rl(self, visual_inputs, max_len=20):
        self.called.append("sampler_rl")
        return "reference rollout"

    def beam_search_sampler(self, visual_inputs, beam_size=5):
        self.called.append("beam_search_sampler")

    def eval_test_image(self, *a, **k):
        self.called.append("eval_test_image")


def _live_engine(name):
    meta, gold = load_case(name)
    sd, feats, mask = rebuild(meta)
    d = meta["dims"]
    eng = types.SimpleNamespace(
        model=_RefModel(sd), device="cuda:0", caption_vocab=_Vocab(d["vocab_size"]),
        settings=dict(model_type="BUTDDetection", embed_dim=d["embed_dim"], hidden_dim=d["hidden_dim"], atten_dim=d["atten_dim"]))
    return eng, meta, gold, feats, d


def test_install_rebinds_the_eval_methods_of_a_live_engine():
    """INTEGRATION.md section 3 / Main.py:111-129: after ``install(engine)`` the model's eval-time decode methods give the
    reference's golden outputs; ``sampler_rl`` (the SCST training rollout) is left alone."""
    from simpleimagecaptionzoo_b200 import engine
    eng, meta, gold, feats, d = _live_engine("butd_tiny_k3")
    fast = engine.install(eng, max_seq=meta["T"], max_batch=meta["B"], max_regions=meta["R"], max_rows=meta["K"],
                          enc_dim=d["enc_dim"], math="f16x3")
    vi = {"bu_feats": torch.from_numpy(feats).cuda(), "bu_bboxes": None, "bu_masks": None}
    tok = eng.model.beam_search_sampler(vi, beam_size=meta["K"])
    assert tok.dtype == torch.int64 and tuple(tok.shape) == (meta["B"], 1 + meta["T"])
    o = orc.make_decoder("BUTD", rebuild(meta)[0])
    o.prepare(feats)
    res = orc.beam_search_batched(o, meta["K"], meta["T"])
    verdict = orc.agreement(tok.cpu().numpy(), gold["tokens"], res.min_gap, tol=1e-4)
    assert "diff" not in verdict and np.mean([v == "exact" for v in verdict]) >= 0.95, verdict
    greedy = eng.model.sampler(vi, max_len=meta["T"])
    assert greedy.dtype == torch.int64 and (greedy.cpu().numpy() == gold["greedy"]).all(1).mean() >= 0.95
    assert eng.model.sampler_rl(vi, max_len=meta["T"]) == "reference rollout"  # not rebound by default
    assert eng.model.called == ["sampler_rl"]  # ... and none of the rebound methods reached the reference's code
    # single-image test path: caption words + attention maps of the generated words
    words, alphas = eng.model.eval_test_image({"bu_feats": vi["bu_feats"][:1], "bu_masks": None}, eng.caption_vocab,
                                              max_len=meta["T"], eval_beam_size=meta["K"])
    want = orc.ids_to_caption(gold["tokens"][0], eng.caption_vocab.ix2word).split()
    if verdict[0] == "exact":
        assert words == want
    assert len(alphas) == 1 and alphas[0].shape[0] == 1 and alphas[0].shape[2] == meta["R"]
    assert torch.allclose(alphas[0].sum(-1), torch.ones_like(alphas[0].sum(-1)), atol=1e-3)
    fast.decoder.close()


def test_install_with_rebind_rl_keeps_scst_training_on_the_reference_path():
    """an installed engine must still train.  With ``rebind_rl=True`` the fast rollout serves no-grad / eval
    calls; with autograd recording and the model in training mode the reference's own ``sampler_rl`` runs."""
    from simpleimagecaptionzoo_b200 import engine
    eng, meta, gold, feats, d = _live_engine("butd_tiny_k3")
    fast = engine.install(eng, rebind_rl=True, max_seq=meta["T"], max_batch=meta["B"], max_regions=meta["R"], max_rows=meta["K"],
                          enc_dim=d["enc_dim"], math="f16x3")
    vi = {"bu_feats": torch.from_numpy(feats).cuda(), "bu_bboxes": None, "bu_masks": None}
    eng.model.train()
    with torch.enable_grad():
        assert eng.model.sampler_rl(vi, max_len=meta["T"]) == "reference rollout"
    assert eng.model.called == ["sampler_rl"]
    eng.model.eval()
    with torch.no_grad():
        seq, logp = eng.model.sampler_rl(vi, max_len=meta["T"], n_per_image=meta["n_samples"], seed=meta["sample_seed"])
    assert eng.model.called == ["sampler_rl"]
    n = meta["n_samples"]
    assert seq.dtype == torch.int64 and tuple(seq.shape) == (meta["B"] * n, meta["T"]) and tuple(logp.shape) == tuple(seq.shape)
    same = (seq.cpu().numpy().reshape(meta["B"], n, -1) == gold["sample_seq"]).all(2)
    assert same.mean() >= 0.95
    err = np.abs(logp.cpu().numpy().reshape(meta["B"], n, -1) - gold["sample_logprobs"])[same]
    assert err.max() < 1e-3
    fast.decoder.close()


def test_captioner_sampler_rl_and_scst_rollouts_wrappers():
    """``B200Captioner.sampler_rl`` / ``scst_rollouts`` == the decoder calls they wrap == the reference's golden rollout."""
    from simpleimagecaptionzoo_b200 import capdec, engine
    meta, gold = load_case("butd_tiny_k3")
    sd, feats, _ = rebuild(meta)
    d = meta["dims"]
    n = meta["n_samples"]
    cap = engine.B200Captioner("BUTDDetection", dict(embed_dim=d["embed_dim"], hidden_dim=d["hidden_dim"], atten_dim=d["atten_dim"]),
                               d["vocab_size"], sd, max_batch=meta["B"], max_regions=meta["R"], max_rows=n + 1, max_seq=meta["T"],
                               math="f16x3", enc_dim=d["enc_dim"])
    vi = {"bu_feats": torch.from_numpy(feats).cuda()}
    seq, logp = cap.sampler_rl(vi, max_len=meta["T"], n_per_image=n, seed=meta["sample_seed"])
    tok, lp = cap.decoder.sample(capdec.SAMPLE_MULTINOMIAL, n, meta["sample_seed"], meta["T"])
    assert torch.equal(seq, tok.long()) and torch.equal(logp, lp)
    assert (seq.cpu().numpy().reshape(meta["B"], n, -1) == gold["sample_seq"]).all(2).mean() >= 0.95
    greedy, seq2, logp2 = cap.scst_rollouts(vi, max_len=meta["T"], n_per_image=n, seed=meta["sample_seed"])
    assert torch.equal(seq2, seq) and torch.allclose(logp2, logp, atol=2e-4)  # other row count -> other GEMM tiling / summation order
    assert torch.equal(greedy, cap.sampler(vi, max_len=meta["T"]))
    # consecutive calls without a seed draw from different streams
    a, _ = cap.sampler_rl(vi, max_len=meta["T"])
    b, _ = cap.sampler_rl(vi, max_len=meta["T"])
    assert not torch.equal(a, b)
    cap.decoder.close()


def test_decode_graph_cache_serves_alternating_shapes(monkeypatch):
    """A ragged last batch, alternating beam sizes and the rollouts each capture ONCE (8-entry cache); replayed results are
    identical to eager launches (CAPDEC_NO_GRAPH=1)."""
    from simpleimagecaptionzoo_b200 import capdec
    meta, gold = load_case("butd_tiny_k3")
    sd, feats, _ = rebuild(meta)
    d = meta["dims"]
    kw = dict(hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"], atten_dim=d["atten_dim"],
              enc_dim=d["enc_dim"], max_batch=meta["B"], max_regions=meta["R"], max_rows=5, max_seq=meta["T"], math="f16x3")
    ft = torch.from_numpy(feats).cuda()

    def run(dec):
        out = []
        for B, K in ((meta["B"], 3), (5, 3), (meta["B"], 2), (meta["B"], 3), (5, 3), (meta["B"], 2)):
            dec.prepare(ft[:B])
            out.append(dec.beam_search(K, meta["T"])[0].clone())
        for seed in (11, 12, 11):
            out += [t.clone() for t in dec.sample(capdec.SAMPLE_MULTINOMIAL, 2, seed, meta["T"])]
        out += [t.clone() for t in dec.sample(capdec.SAMPLE_GREEDY, 1, 0, meta["T"])]
        out += [t.clone() for t in dec.scst_rollout(2, 11, meta["T"])]
        out.append(dec.score(out[-3], 2).clone())  # re-score the sampled rows of the one-pass rollout
        torch.cuda.synchronize()
        return out

    dec = capdec.CaptionDecoder("BUTD", sd, **kw)
    got = run(dec)
    # beam: (B,3) (5,3) (B,2); multinomial n=2; greedy; scst; score
    assert dec.graph_captures == 7, dec.graph_captures
    got2 = run(dec)
    assert dec.graph_captures == 7
    dec.close()
    monkeypatch.setenv("CAPDEC_NO_GRAPH", "1")
    eager = capdec.CaptionDecoder("BUTD", sd, **kw)
    want = run(eager)
    assert eager.graph_captures == 0
    eager.close()
    for a, b, c in zip(got, got2, want):
        assert torch.equal(a, c) and torch.equal(b, c)
    assert torch.equal(got[6], got[10]) and not torch.equal(got[6], got[8])  # same seed same rollout, other seed another one
    # the re-scored log-probs are the rollout's own up to the first stored 0 (sample_rl stores <end> and everything after it
    # as 0 while its log-probs stay those of the words it drew, BUTD_Model.py:226-232)
    live = (got[-4] != 0).long().cumprod(1).bool()
    assert live.any() and torch.allclose(got[-1][live], got[-3][live], atol=1e-4)


@pytest.mark.parametrize("B", [5, 70])  # 15 rows: the small-batch GEMM path; 210 rows: the 256-row pair tiles
def test_aoa_detection_adaptive_features_default_region_limit(B):
    """AoADetection takes adaptive bottom-up features (10-100 boxes per image, padded + bu_masks, AoA_Engine.py:23-47): the
    captioner's default region limit covers them without the caller passing max_regions, and the masked decode
    from 100-box features matches the oracle."""
    from simpleimagecaptionzoo_b200 import engine, synth
    d = dict(synth.TINY_DIMS["AOA"])
    sd = synth.make_state_dict("AOA", seed=3, chaotic=0.3, end_boost=0.2, **d)  # a mix of immediate <end> and full-length captions
    sd.update(synth.make_refiner_state_dict(hidden_dim=d["hidden_dim"], enc_dim=2048, seed=3, chaotic=0.3))
    R, K, T = 100, 3, 20
    mask = synth.make_region_mask(B, R, 10, 3)
    mask[0] = 1  # one image with all 100 boxes
    bu = synth.make_region_feats(B, R, 2048, 3) * mask[:, :, None]
    cap = engine.B200Captioner("AoADetection", dict(embed_dim=d["embed_dim"], hidden_dim=d["hidden_dim"]), d["vocab_size"], sd,
                               max_batch=B, max_rows=K, max_seq=T, math="f16x3", num_heads=d["num_heads"])
    assert cap.decoder.max_regions == 100 and cap.native_refiner
    tok = cap.beam_search_sampler({"bu_feats": torch.from_numpy(bu).cuda(), "bu_masks": torch.from_numpy(mask).cuda()}, beam_size=K)
    o = orc.make_decoder("AOA", sd, num_heads=d["num_heads"])
    o.prepare(orc.aoa_project_refine(sd, bu, mask, num_heads=d["num_heads"]), mask)
    res = orc.beam_search_batched(o, K, T)
    verdict = orc.agreement(tok.cpu().numpy(), res.tokens, res.min_gap, tol=1e-4)
    assert "diff" not in verdict and np.mean([v == "exact" for v in verdict]) >= 0.9, verdict
    with pytest.raises(RuntimeError, match="regions"):
        cap.beam_search_sampler({"bu_feats": torch.zeros(1, 101, 2048).cuda(), "bu_masks": None}, beam_size=K)
    cap.decoder.close()


@pytest.mark.parametrize("name", ["butd_full_k3", "butd_full_k5_r196", "nic_full_k3", "aoa_full_k3", "butd_tiny_k1", "nic_tiny_k5"])
def test_small_batch_path_matches_large_tile_path(name, monkeypatch):
    """<= 128 activation rows take the swap-AB split-K kernel (smallm.cuh), with the dependent GEMMs of a step fused into one
    launch; the same decode through separate launches (CAPDEC_NO_FUSE=1) and through the large-tile kernels
    (CAPDEC_NO_SMALLM=1) must give the same captions and (up to fp32 summation order) the same scores / log-probs."""
    from simpleimagecaptionzoo_b200 import capdec
    meta, gold = load_case(name)
    sd, feats, mask = rebuild(meta)
    d = meta["dims"]
    B = min(meta["B"], 128 // max(meta["K"], 2))
    assert B * meta["K"] <= 128
    kw = dict(hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"], atten_dim=d.get("atten_dim", 0),
              enc_dim=d.get("enc_dim", 2048), num_heads=d.get("num_heads", 8), max_batch=B, max_regions=max(meta["R"], 1),
              max_rows=max(meta["K"], 2), max_seq=meta["T"], math="f16x3")
    ft = torch.from_numpy(feats[:B]).cuda()
    mk = None if mask is None else torch.from_numpy(mask[:B]).cuda()

    def run():
        dec = capdec.CaptionDecoder(meta["arch"], sd, **kw)
        dec.prepare(ft, mk)
        out = list(dec.beam_search(meta["K"], meta["T"]))
        out += list(dec.sample(capdec.SAMPLE_MULTINOMIAL, 2, 5, meta["T"]))
        out += list(dec.sample(capdec.SAMPLE_GREEDY, 1, 0, meta["T"]))
        out = [t.clone() for t in out]
        torch.cuda.synchronize()
        dec.close()
        return out

    fused = run()
    monkeypatch.setenv("CAPDEC_NO_FUSE", "1")
    unfused = run()
    monkeypatch.setenv("CAPDEC_NO_SMALLM", "1")
    large = run()
    for a, b in zip(fused, unfused):
        assert torch.equal(a, b)  # same kernels, same split-K order: bit-identical
    tok, score, length, seq, logp, gseq, glogp = fused
    ltok, lscore, llength, lseq, llogp, lgseq, lglogp = large
    same = (tok == ltok).all(1)
    assert same.float().mean() >= 0.9 and torch.allclose(score[same], lscore[same], atol=2e-4)
    assert (tok.cpu().numpy() == gold["tokens"][:B]).all(1).mean() >= 0.9
    srow = (seq == lseq).all(1)
    assert srow.float().mean() >= 0.9 and torch.allclose(logp[srow], llogp[srow], atol=2e-4)
    grow = (gseq == lgseq).all(1)
    assert grow.float().mean() >= 0.9 and torch.allclose(glogp[grow], lglogp[grow], atol=2e-4)


@pytest.mark.parametrize("name", ["butd_full_k3", "nic_full_k3"])
def test_sampling_epilogue_draws_do_not_depend_on_its_warp_count(name, monkeypatch):
    """The large-tile logit GEMM's sampling epilogue runs on 12 warps (three column shares per tile and accumulator row,
    CAPDEC_LOGIT_EW=8: two).  The noise is a function of (seed, row, step, word) only and the shares' partial records meet in
    the bookkeeping kernel, so the drawn words are the same and the log-probs agree to fp32 summation order; likewise SCST's
    one-pass rollouts and teacher-forced re-scoring."""
    from simpleimagecaptionzoo_b200 import capdec
    meta, gold = load_case(name)
    sd, feats, mask = rebuild(meta)
    d = meta["dims"]
    B, n = meta["B"], 5
    monkeypatch.setenv("CAPDEC_NO_SMALLM", "1")  # the large-tile kernels at the goldens' 96 rows
    kw = dict(hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"], atten_dim=d.get("atten_dim", 0),
              enc_dim=d.get("enc_dim", 2048), num_heads=d.get("num_heads", 8), max_batch=B, max_regions=max(meta["R"], 1),
              max_rows=n + 1, max_seq=meta["T"], math="f16x3")
    ft = torch.from_numpy(feats).cuda()

    def run():
        dec = capdec.CaptionDecoder(meta["arch"], sd, **kw)
        dec.prepare(ft, None)
        seq, logp = dec.sample(capdec.SAMPLE_MULTINOMIAL, n, 17, meta["T"])
        out = [seq.clone(), logp.clone()] + [t.clone() for t in dec.scst_rollout(n, 23, meta["T"])]
        out.append(dec.score(seq, n).clone())
        torch.cuda.synchronize()
        dec.close()
        return out

    twelve = run()
    monkeypatch.setenv("CAPDEC_LOGIT_EW", "8")
    eight = run()
    for a, b in zip(twelve, eight):
        if a.dtype in (torch.int32, torch.int64):
            assert torch.equal(a, b)
        else:
            assert torch.allclose(a, b, atol=2e-4)


@pytest.mark.parametrize("arch,B,R,K,T", [("BUTD", 3, 1, 8, 20), ("BUTD", 1, 36, 8, 1), ("BUTD", 40, 6, 2, 5), ("NIC", 7, 0, 8, 20),
                                         ("AOA", 2, 1, 8, 3), ("AOA", 33, 12, 4, 20), ("BUTD", 300, 4, 7, 9)])
def test_extreme_shapes_against_the_oracle(arch, B, R, K, T):
    """Limits of the interface -- the widest beam the library takes (8), one region, one step, one image, beams that are not
    a template size (2, 4, 7), row counts on both sides of the small-batch limit -- against the numpy oracle (tiny chaotic
    models that emit <end> at varied lengths, fp32-grade mode: exact or tie-justified)."""
    from simpleimagecaptionzoo_b200 import capdec, synth
    d = dict(synth.TINY_DIMS[arch])
    # <end> boosts chosen so that, even with the widest beam, some hypotheses complete early and others run to the limit
    sd = synth.make_state_dict(arch, seed=11, chaotic=0.3 if arch == "AOA" else True,
                               end_boost={"BUTD": 0.3, "NIC": -0.25, "AOA": 0.0}[arch], **d)
    if arch == "BUTD":
        feats = synth.make_region_feats(B, R, d["enc_dim"], 11)
    elif arch == "NIC":
        feats = synth.make_image_embed(B, d["embed_dim"], 11)
    else:
        feats = synth.make_refined_feats(B, R, d["hidden_dim"], 11)
    dec = capdec.CaptionDecoder(arch, sd, hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"],
                                atten_dim=d.get("atten_dim", 0), enc_dim=d.get("enc_dim", 2048), num_heads=d.get("num_heads", 8),
                                max_batch=B, max_regions=max(R, 1), max_rows=8, max_seq=20, math="f16x3")
    dec.prepare(torch.from_numpy(feats).cuda())
    tok, score, length = dec.beam_search(K, T)
    greedy, glp = dec.sample(capdec.SAMPLE_GREEDY, 1, 0, T)
    torch.cuda.synchronize()
    o = orc.make_decoder(arch, sd, num_heads=d.get("num_heads", 8))
    o.prepare(feats)
    res = orc.beam_search_batched(o, K, T)
    verdict = orc.agreement(tok.cpu().numpy(), res.tokens, res.min_gap, tol=1e-4)
    assert "diff" not in verdict, verdict
    exact = np.array([v == "exact" for v in verdict])
    assert exact.mean() >= 0.8
    assert np.array_equal(length.cpu().numpy()[exact], res.lengths[exact])
    assert np.allclose(score.cpu().numpy()[exact], res.scores[exact], atol=1e-3)
    ids, gaps, _ = orc.greedy_sample(o, T)
    g = greedy.cpu().numpy()
    for b in range(B):
        if not np.array_equal(g[b], ids[b]):
            t = int(np.argmax(g[b] != ids[b]))
            assert gaps[b, t] < 1e-4, (b, t, gaps[b, t])
    with pytest.raises(RuntimeError):
        dec.beam_search(9, T)  # wider than max_rows
    dec.close()


@pytest.mark.parametrize("B,reps", [(1, 60), (16, 100), (42, 40)])
def test_small_batch_path_is_bit_stable_over_many_replays(B, reps):
    """The split-K exchange sums in split order and the CTAs synchronise through counters: a race there would show up as
    run-to-run differences.  The same decode replayed many times (beam search + multinomial rollout, both math modes' default)
    must return bit-identical tokens AND scores every time."""
    from simpleimagecaptionzoo_b200 import capdec, synth
    d = synth.DIMS["BUTD"]
    sd = synth.make_state_dict("BUTD", seed=0, **d)
    dec = capdec.CaptionDecoder("BUTD", sd, hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"],
                                atten_dim=d["atten_dim"], enc_dim=d["enc_dim"], max_batch=B, max_regions=36, max_rows=3, max_seq=20)
    dec.prepare(torch.from_numpy(synth.make_region_feats(B, 36, d["enc_dim"], 5)).cuda())
    tok0, score0, _ = (t.clone() for t in dec.beam_search(3, 20))
    seq0, lp0 = (t.clone() for t in dec.sample(capdec.SAMPLE_MULTINOMIAL, 3, 9, 20))
    for _ in range(reps):
        tok, score, _ = dec.beam_search(3, 20)
        assert torch.equal(tok, tok0) and torch.equal(score, score0)
    for _ in range(reps // 4):
        seq, lp = dec.sample(capdec.SAMPLE_MULTINOMIAL, 3, 9, 20)
        assert torch.equal(seq, seq0) and torch.equal(lp, lp0)
    dec.close()


def test_small_batch_launches_of_any_count_and_kind_follow_each_other(monkeypatch):
    """The fused small-batch launches synchronise through global counters that must be back at zero whenever a launch ends:
    an ODD number of fused launches per decode (NIC: one per step, 15 steps), graph replays back to back, and eager launches
    (attention maps requested) between two replays all give the captions of the unfused path."""
    from simpleimagecaptionzoo_b200 import capdec
    out = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("CAPDEC_NO_FUSE", "0" if fuse == "1" else "1")
        meta, gold = load_case("nic_full_k3")
        sd, feats, _ = rebuild(meta)
        d = meta["dims"]
        nic = capdec.CaptionDecoder("NIC", sd, hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"],
                                    max_batch=meta["B"], max_rows=3, max_seq=20, math="f16x3")
        nic.prepare(torch.from_numpy(feats).cuda())
        res = [nic.beam_search(3, 15)[0].clone() for _ in range(3)]  # 15 fused launches per decode, replayed
        res += [nic.beam_search(3, 20)[0].clone(), nic.beam_search(3, 15)[0].clone()]
        nic.close()
        meta, gold = load_case("butd_full_k3")
        sd, feats, _ = rebuild(meta)
        d = meta["dims"]
        butd = capdec.CaptionDecoder("BUTD", sd, hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"],
                                     atten_dim=d["atten_dim"], enc_dim=d["enc_dim"], max_batch=meta["B"], max_regions=meta["R"], max_rows=3,
                                     max_seq=meta["T"], math="f16x3")
        butd.prepare(torch.from_numpy(feats).cuda())
        res.append(butd.beam_search(3, meta["T"])[0].clone())                     # captured
        res.append(butd.beam_search(3, meta["T"], return_alphas=True)[0].clone())  # eager (attention maps)
        res.append(butd.beam_search(3, 7)[0].clone())                              # another graph, odd step count
        res.append(butd.beam_search(3, meta["T"])[0].clone())                      # replay of the first
        torch.cuda.synchronize()
        assert torch.equal(res[5], res[6]) and torch.equal(res[5], res[8])
        assert (res[5].cpu().numpy() == gold["tokens"]).all(1).mean() >= 0.9
        butd.close()
        out[fuse] = res
    for a, b in zip(out["1"], out["0"]):
        assert torch.equal(a, b)
    assert torch.equal(out["1"][0], out["1"][1]) and torch.equal(out["1"][0], out["1"][2]) and torch.equal(out["1"][0], out["1"][4])


@pytest.mark.parametrize("shape", [(1, 9487, 1024), (3, 4096, 2048), (48, 1000, 64), (64, 130, 192), (65, 4096, 4096), (128, 96, 640)])
@pytest.mark.parametrize("math", ["f16", "f16x3"])
def test_small_batch_gemm_against_torch_fp64(shape, math):
    """The swap-AB split-K GEMM (1..128 rows; partial weight tiles, every K split) against a plain fp64 matmul."""
    from simpleimagecaptionzoo_b200 import capdec
    m, n, k = shape
    g = torch.Generator().manual_seed(m * 131 + n)
    a = torch.randn(m, k, generator=g).cuda()
    b = (torch.randn(n, k, generator=g) * 0.05).cuda()
    bias = torch.randn(n, generator=g).cuda()
    d = capdec.test_gemm(a, b, bias, math)
    ref = (a.half().double() @ b.half().double().T if math == "f16" else a.double() @ b.double().T) + bias.double()
    err = (d.double() - ref).abs().max().item()
    assert err <= 4e-6 * max(ref.abs().max().item(), 1.0) * max(1.0, (k / 64) ** 0.5), err


@pytest.mark.parametrize("name", ["butd_tiny_k3", "nic_tiny_k3", "aoa_tiny_k3_masked", "butd_full_k3"])
def test_rescore_with_autograd_through_the_vocabulary_layer(name):
    """seqLogprobs of a fused rollout WITH a graph -- capdec_score_states exports the rows `predict` saw,
    scst.differentiable_logprobs rebuilds predict + log_softmax + gather on the live weight-normed parameters."""
    from simpleimagecaptionzoo_b200 import capdec, scst
    meta, gold = load_case(name)
    sd, feats, mask = rebuild(meta)
    d = meta["dims"]
    dec = capdec.CaptionDecoder(meta["arch"], sd, hidden_dim=d["hidden_dim"], embed_dim=d["embed_dim"], vocab_size=d["vocab_size"],
                                atten_dim=d.get("atten_dim", 0), enc_dim=d.get("enc_dim", 2048), num_heads=d.get("num_heads", 8),
                                max_batch=meta["B"], max_regions=max(meta["R"], 1), max_rows=2, max_seq=meta["T"], math="f16x3")
    dec.prepare(torch.from_numpy(feats).cuda(), None if mask is None else torch.from_numpy(mask).cuda())
    seq, logp = dec.sample(capdec.SAMPLE_MULTINOMIAL, 2, 7, meta["T"])
    want = dec.score(seq, 2)
    got, states = dec.score(seq, 2, return_states=True)
    assert torch.equal(got, want) and tuple(states.shape) == (meta["B"] * 2, meta["T"], d["hidden_dim"])
    pre = "decoder." if "decoder.predict.weight_g" in sd else ""
    g, v, b = (torch.from_numpy(np.array(sd[pre + "predict." + k])).cuda().requires_grad_(True) for k in ("weight_g", "weight_v", "bias"))
    lp = scst.differentiable_logprobs(states, seq, g, v, b)
    assert lp.requires_grad and torch.allclose(lp, want, atol=1e-3)
    # the policy-gradient loss of RewardCriterion: -(logprob * reward * mask).sum() / mask.sum()  (Utils.py:305-316)
    reward = torch.linspace(-1, 1, lp.numel(), device=lp.device).view_as(lp)
    live = (seq != 0).long().cumprod(1).float()
    loss = -(lp * reward * live).sum() / live.sum().clamp(min=1)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in (g, v, b))
    # d loss / d bias[w] = -sum_pos reward*live*(1[word == w] - p_w) / n: the entries sum to zero
    assert abs(float(b.grad.sum())) < 1e-3 * float(b.grad.abs().sum() + 1e-6) + 1e-5
    dec.close()


AGREEMENT = [
    # set, images, math, required exact-or-tie fraction vs the reference's own tokens
    ("butd", 5000, "f16", 0.99),
    ("butd", 5000, "f16x3", 0.999),
    ("aoa_bu", 1000, "f16", 0.95),
    ("aoa_bu", 1000, "f16x3", 0.99),
]


@pytest.mark.parametrize("name,images,math,need", AGREEMENT)
def test_agreement_with_the_reference_tokens(name, images, math, need):
    """north_star's agreement target, measured DIRECTLY against captions the reference's own code produced."""
    if not os.path.exists(au.set_path(name, images)):
        pytest.skip(f"{au.set_path(name, images)} not generated (tests/golden/make_agreement_set.py)")
    r = au.evaluate(name, images, math)
    assert r["exact_or_tie_frac"] >= need, {k: v for k, v in r.items() if k != "diffs"}
    # every remaining difference sits where the reference's own top-(k+1) gap is inside the mode's operand-rounding error
    bound = 2e-2 if math == "f16" else 1e-3
    wide = [x for x in r["diffs"] if x["min_gap_up_to_step"] >= bound]
    assert len(wide) <= max(1, images // 1000), wide
