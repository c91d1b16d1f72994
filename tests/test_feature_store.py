"""Packed fp16 feature shards + loader + batch detokeniser.  CPU tests: file format round
trip, ragged (adaptive) features -> masks, batching; GPU tests (-m gpu): fp16 features through capdec_prepare_f16 give the
captions of the fp32 path on the same rounded values, and a shard streamed through the pipelined captioner API."""
import numpy as np
import pytest

from simpleimagecaptionzoo_b200 import feature_store as fs
from simpleimagecaptionzoo_b200 import synth


def _write(tmp_path, n=10, R=6, D=64, ragged=False, seed=0):
    rng = np.random.default_rng(seed)
    feats, ids = [], []
    path = str(tmp_path / "feats.shard")
    with fs.FeatureShardWriter(path, R, D, with_bboxes=True) as w:
        for i in range(n):
            m = int(rng.integers(1, R + 1)) if ragged else R
            f = np.abs(rng.standard_normal((m, D))).astype(np.float32)
            w.append(1000 + i, f, rng.random((m, 4)).astype(np.float32))
            feats.append(f)
            ids.append(1000 + i)
    return path, feats, ids


def test_round_trip_is_fp16_rounding(tmp_path):
    path, feats, ids = _write(tmp_path)
    sh = fs.FeatureShard(path)
    assert len(sh) == 10 and list(sh.image_ids) == ids and sh.R == 6 and sh.D == 64
    for i, f in enumerate(feats):
        assert np.array_equal(sh.features(i), f.astype(np.float16))


def test_ragged_features_give_masks_and_batches(tmp_path):
    torch = pytest.importorskip("torch")
    path, feats, ids = _write(tmp_path, n=11, ragged=True, seed=1)
    sh = fs.FeatureShard(path)
    seen = []
    for image_ids, vi in sh.batches(4, pinned=False):
        B = len(image_ids)
        assert vi["bu_feats"].shape == (B, 6, 64) and vi["bu_feats"].dtype == torch.float16
        assert vi["bu_masks"] is not None and vi["bu_masks"].shape == (B, 6)
        for b in range(B):
            i = ids.index(int(image_ids[b]))
            n = feats[i].shape[0]
            assert vi["bu_masks"][b].sum() == n
            assert np.array_equal(vi["bu_feats"][b, :n].numpy(), feats[i].astype(np.float16))
            assert (vi["bu_feats"][b, n:] == 0).all()
            assert vi["bu_bboxes"][b].shape == (n, 4)
        seen.extend(int(x) for x in image_ids)
    assert seen == ids  # file order, last batch short


@pytest.mark.parametrize("ring", [2, 3, 5])
def test_ring_slots_are_never_rewritten_under_the_consumer(tmp_path, ring):
    """The slot of batch k goes back to the fill thread only when batch k+2 is requested (and after the consumer's
    registered copy event): a consumer that holds the last two batches always sees their own rows, for every ring size."""
    pytest.importorskip("torch")
    path, feats, ids = _write(tmp_path, n=23, seed=2)
    sh = fs.FeatureShard(path)

    class Event:  # stands in for the CUDA event beam_search_stream registers
        def __init__(self):
            self.synced = False

        def synchronize(self):
            self.synced = True

    held, events = [], []
    for image_ids, vi in sh.batches(2, pinned=False, ring=ring):
        ev = Event()
        vi["_on_copied"](ev)
        events.append(ev)
        held.append((list(image_ids), vi["bu_feats"]))
        for hid, buf in held[-2:]:  # the two batches a consumer may still be using
            for b, image_id in enumerate(hid):
                assert np.array_equal(buf[b].numpy(), feats[ids.index(int(image_id))].astype(np.float16))
    assert [i for h, _ in held for i in h] == ids
    assert all(e.synced for e in events[:max(0, len(events) - ring - 1)])  # refilled slots waited for their copy
    with pytest.raises(ValueError):
        next(sh.batches(2, pinned=False, ring=1))


def test_truncated_shard_is_rejected(tmp_path):
    path, _, _ = _write(tmp_path, n=6)
    size = __import__("os").path.getsize(path)
    with open(path, "r+b") as f:
        f.truncate(size - 8)  # cuts into the last block of the file
    with pytest.raises(ValueError):
        fs.FeatureShard(path)


def test_writer_rejects_bad_rows(tmp_path):
    w = fs.FeatureShardWriter(str(tmp_path / "x.shard"), 4, 8)
    with pytest.raises(ValueError):
        w.append(1, np.zeros((5, 8), np.float32))
    with pytest.raises(ValueError):
        w.append(1, np.full((2, 8), 1e6, np.float32))  # leaves the fp16 range
    with pytest.raises(ValueError):
        fs.FeatureShard(__file__)


def test_ids_to_captions_matches_engine_loop():
    from simpleimagecaptionzoo_b200.engine import ids_to_caption
    rng = np.random.default_rng(0)
    ix2word = ["<pad>", "<sta>", "<end>", "<unk>"] + [f"w{i}" for i in range(4, 50)]
    caps = rng.integers(0, 50, size=(64, 21))
    caps[:, 0] = 1
    caps[5] = 1  # only <sta>
    caps[6, 1] = 2  # empty caption
    got = fs.ids_to_captions(caps, ix2word)
    assert got == [ids_to_caption(c, ix2word) for c in caps]
    assert fs.ids_to_captions(caps, dict(enumerate(ix2word))) == got


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["BUTD", "AOA"])
def test_fp16_features_decode_like_rounded_fp32(arch):
    torch = pytest.importorskip("torch")
    from simpleimagecaptionzoo_b200 import capdec
    dims = synth.DIMS[arch]
    sd = synth.make_state_dict(arch, seed=0, **dims)
    if arch == "AOA":
        sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
    B, R = 24, 36
    dec = capdec.CaptionDecoder(arch, sd, hidden_dim=dims["hidden_dim"], embed_dim=dims["embed_dim"], vocab_size=dims["vocab_size"],
                                atten_dim=dims.get("atten_dim", 0), enc_dim=2048, max_batch=B, max_regions=R, max_rows=3, max_seq=20)
    f16 = torch.from_numpy(synth.make_region_feats(B, R, 2048, 7)).half().cuda()
    prep = dec.prepare_bottom_up if arch == "AOA" else dec.prepare
    prep(f16)
    a, sa, _ = dec.beam_search(3, 20)
    prep(f16.float())
    b, sb, _ = dec.beam_search(3, 20)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.allclose(sa, sb, atol=1e-5)
    with pytest.raises(RuntimeError, match="f16x3|fp32"):
        d2 = capdec.CaptionDecoder("BUTD", synth.make_state_dict("BUTD", seed=0, **synth.TINY_DIMS["BUTD"]), math="f16x3",
                                   max_batch=4, max_regions=6, **{k: v for k, v in synth.TINY_DIMS["BUTD"].items()})
        d2.prepare(torch.zeros(4, 6, 128, dtype=torch.float16).cuda())
    dec.close()


@pytest.mark.gpu
def test_shard_streams_through_the_captioner(tmp_path):
    torch = pytest.importorskip("torch")
    from simpleimagecaptionzoo_b200 import engine
    dims = synth.DIMS["BUTD"]
    sd = synth.make_state_dict("BUTD", seed=0, **dims)
    N, R = 40, 36
    feats = synth.make_region_feats(N, R, 2048, 3)
    path = str(tmp_path / "coco.shard")
    with fs.FeatureShardWriter(path, R, 2048) as w:
        for i in range(N):
            w.append(i, feats[i])
    settings = dict(model_type="BUTDDetection", embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"], atten_dim=dims["atten_dim"])
    cap = engine.B200Captioner("BUTDDetection", settings, dims["vocab_size"], sd, max_batch=16, max_regions=R, max_rows=3)
    shard = fs.FeatureShard(path)
    ids, toks = [], []
    stream = shard.batches(16)
    order = []

    def inputs():
        for image_ids, vi in stream:
            order.append(image_ids)
            yield vi

    for t in cap.beam_search_stream(inputs(), beam_size=3):
        toks.append(t)
    got = np.concatenate(toks)
    assert np.array_equal(np.concatenate(order), np.arange(N)) and got.shape == (N, 21)
    ref = cap.beam_search_sampler({"bu_feats": torch.from_numpy(feats[:16]).half().float()}, beam_size=3).cpu().numpy()
    assert np.array_equal(got[:16], ref)


@pytest.mark.gpu
def test_fp16_refined_features_are_rejected_cleanly():
    """AoA without the refiner entries: fp16 input has no meaning (the decoder takes refined features in fp32)."""
    torch = pytest.importorskip("torch")
    from simpleimagecaptionzoo_b200 import capdec
    dims = synth.TINY_DIMS["AOA"]
    dec = capdec.CaptionDecoder("AOA", synth.make_state_dict("AOA", seed=0, **dims), max_batch=4, max_regions=6, **dims)
    with pytest.raises(RuntimeError, match="fp32"):
        dec.prepare(torch.zeros(4, 6, dims["hidden_dim"], dtype=torch.float16).cuda())
    dec.prepare(torch.zeros(4, 6, dims["hidden_dim"]).cuda())  # the handle is still usable
    tok, _, _ = dec.beam_search(3, 5)
    assert tok.shape == (4, 6)
    dec.close()
