"""GPU tests of the CNN encoder feed: the fp16 channels-last, CUDA-graph-replayed ResNet-101
feed against the reference's own encoder arithmetic -- the same torchvision modules in fp32 (NIC_Model.py:8-37,
BUTD_Model.py:8-38) -- and the captioner API on raw images."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from simpleimagecaptionzoo_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu

# stated bound: fp16 storage of activations through 101 convolution layers (fp32 accumulation inside cuDNN)
REL_L2_BOUND = 2e-2


def _fp32_reference(sd, images, model_type, grid=7):
    from simpleimagecaptionzoo_b200 import cnn_feed
    fx = cnn_feed.build_feature_extractor()
    fx.load_state_dict({k[len("encoder.feature_extractor."):]: v for k, v in sd.items() if k.startswith("encoder.feature_extractor.")})
    fx = fx.eval().cuda()
    with torch.no_grad():
        f = fx(images.cuda())
        if model_type == "NIC":
            v, g = sd["encoder.img_embedding.weight_v"].double(), sd["encoder.img_embedding.weight_g"].double()
            W = (v * (g / v.norm(dim=1, keepdim=True))).float().cuda()
            return torch.addmm(sd["encoder.img_embedding.bias"].float().cuda(), f.mean(dim=(2, 3)), W.t())
        f = torch.nn.functional.adaptive_avg_pool2d(f, (grid, grid))
        return f.permute(0, 2, 3, 1).reshape(f.shape[0], -1, f.shape[1])


@pytest.mark.parametrize("model_type,grid", [("NIC", 7), ("BUTDSpatial", 7), ("AoASpatial", 7)])
def test_feed_matches_fp32_torchvision(model_type, grid):
    from simpleimagecaptionzoo_b200 import cnn_feed
    sd = cnn_feed.make_encoder_state_dict(embed_dim=512 if model_type == "NIC" else None, seed=1)
    images = torch.randn(6, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    feed = cnn_feed.CnnFeed(model_type, sd, enc_img_size=grid)
    got = feed({"img_tensors": images.pin_memory()})
    again = feed({"img_tensors": images.pin_memory()})  # second call replays the captured graph
    torch.cuda.synchronize()
    ref = _fp32_reference(sd, images, model_type, grid)
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < REL_L2_BOUND, rel
    assert torch.equal(got, again)
    if model_type != "NIC":
        assert got.shape == (6, grid * grid, 2048)


def test_nic_captioner_on_raw_images():
    """B200Captioner('NIC') + attached feed: captions from images == captions decoded from the feed's own embedding."""
    from simpleimagecaptionzoo_b200 import cnn_feed, engine
    dims = synth.DIMS["NIC"]
    sd = synth.make_state_dict("NIC", seed=0, **dims)
    sd.update(cnn_feed.make_encoder_state_dict(embed_dim=dims["embed_dim"], seed=0))
    settings = dict(model_type="NIC", embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"])
    cap = engine.B200Captioner("NIC", settings, dims["vocab_size"], sd, max_batch=8, max_rows=3, max_seq=20)
    feed = cnn_feed.attach(cap, sd)
    images = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    tok = cap.beam_search_sampler({"img_tensors": images}, beam_size=3)
    emb = feed({"img_tensors": images})
    cap.decoder.prepare(emb)
    tok2, _, _ = cap.decoder.beam_search(3, 20)
    assert tok.shape == (8, 21) and tok.dtype == torch.int64
    assert torch.equal(tok.int(), tok2)
    streamed = list(cap.beam_search_stream(({"img_tensors": images.pin_memory()} for _ in range(3)), beam_size=3))
    for s in streamed:
        assert np.array_equal(s, tok2.cpu().numpy())


def test_aoa_spatial_from_images_runs_the_native_refiner():
    """AoASpatial: images -> CNN grid (feed) -> img_feats_porjection + aoa_refine + decoder inside the library; refined
    features equal the oracle's refiner applied to the feed's own grid."""
    from oracle import capdec_oracle as orc
    from simpleimagecaptionzoo_b200 import cnn_feed, engine
    dims = synth.DIMS["AOA"]
    sd = synth.make_state_dict("AOA", seed=0, **dims)
    sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
    sd.update(cnn_feed.make_encoder_state_dict(seed=0))
    settings = dict(model_type="AoASpatial", embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"], enc_img_size=7)
    cap = engine.B200Captioner("AoASpatial", settings, dims["vocab_size"], sd, max_batch=4, max_rows=3, max_seq=20)
    feed = cnn_feed.attach(cap, sd)
    assert cap.native_refiner
    images = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(9))
    tok = cap.beam_search_sampler({"img_tensors": images}, beam_size=3)
    assert tok.shape == (4, 21)
    grid = feed({"img_tensors": images})
    assert grid.shape == (4, 49, 2048)
    got = cap.decoder.refined_features().cpu().numpy()
    ref = orc.aoa_project_refine({k: v for k, v in sd.items() if isinstance(v, np.ndarray)}, grid.cpu().numpy(), None, num_heads=8)
    assert np.abs(got - ref).max() < 6e-2


def test_fused_trunk_is_used_and_matches_modules():
    """The folded-BN / fused-epilogue trunk is what runs (the constructor's self-check accepted it) and agrees with the
    plain module forward of the same weights."""
    from simpleimagecaptionzoo_b200 import cnn_feed
    sd = cnn_feed.make_encoder_state_dict(embed_dim=512, seed=2)
    images = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    fused = cnn_feed.CnnFeed("NIC", sd)
    plain = cnn_feed.CnnFeed("NIC", sd, fuse=False)
    assert fused.trunk is not None and plain.trunk is None
    a, b = fused({"img_tensors": images}), plain({"img_tensors": images})
    assert ((a - b).norm() / b.norm()).item() < REL_L2_BOUND
