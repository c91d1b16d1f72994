"""CIDEr-D self-critical reward.  CPU: the oracle's restatement against the golden rewards the
reference's own ``get_self_critical_reward`` + ``CiderD`` produced (tests/golden/make_golden_cider.py) and the host-side
n-gram hashing.  GPU (-m gpu): ``capdec_cider_reward`` through the C ABI against both."""
import json
import os

import numpy as np
import pytest

from oracle import capdec_oracle as orc
from simpleimagecaptionzoo_b200 import scst, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["cider_n1", "cider_n5"]


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    ix2word, refs = synth.make_caption_corpus(meta["n_images"], meta["vocab"], seed=meta["seed"])
    df, ref_len = scst.document_frequency_from_corpus(refs)
    assert len(df) == meta["df_entries"] and ref_len == meta["n_images"]
    return meta, z, ix2word, refs, df, ref_len


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_reward(name):
    meta, z, ix2word, refs, df, ref_len = _load(name)
    n = meta["n_per_image"]
    gts = dict(enumerate(refs))
    ids = [int(i) for i in np.repeat(z["img_index"], n)]
    rewards, _ = orc.self_critical_reward(z["gen"], np.repeat(z["greedy"], n, axis=0), gts, ids, ix2word, df, ref_len)
    assert rewards.shape == z["rewards"].shape
    assert np.abs(rewards - z["rewards"]).max() < 1e-6
    # the synthetic rollouts regenerate from the seed (fixtures hold the copies the reference scored)
    gen, greedy = synth.make_rollouts(ix2word, refs, list(z["img_index"]), n, meta["max_len"], meta["seed"])
    assert np.array_equal(gen, z["gen"]) and np.array_equal(greedy, z["greedy"])


def test_ngram_keys_are_distinct_and_stable():
    rng = np.random.default_rng(0)
    for k in (1, 2, 3, 4):
        ids = rng.integers(0, 12000, size=(20000, k))
        keys = scst.ngram_keys(ids)
        uniq_rows = len(np.unique(ids, axis=0))
        assert len(np.unique(keys)) == uniq_rows
    # the n-gram length is part of the key: (5,) != (5, 5)
    assert scst.ngram_keys(np.array([[5]]))[0] != scst.ngram_keys(np.array([[5, 5]]))[0]
    assert scst.ngram_keys(np.array([[1, 2, 3]]))[0] == np.uint64(scst.ngram_keys(np.array([[1, 2, 3]]))[0])


def test_word_ids_extend_the_vocabulary():
    w = scst.WordIds({"<pad>": 0, "a": 4, "b": 5})
    assert list(w.sentence("a zebra b zebra yak")) == [4, 6, 5, 6, 7]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_reward_matches_reference(name):
    """capdec_cider_reward == the reference's rewards (fp64 on both sides; stated bound 1e-5 on scores in [0, 10])."""
    torch = pytest.importorskip("torch")
    from simpleimagecaptionzoo_b200 import capdec
    meta, z, ix2word, refs, df, ref_len = _load(name)
    n = meta["n_per_image"]
    lib = capdec.load_library()
    # host hash == device hash
    for ids in ([7], [7, 9], [1, 2, 3], [60, 61, 62, 63]):
        arr = (np.asarray(ids, np.int32))
        assert int(scst.ngram_keys(arr[None, :])[0]) == int(lib.capdec_cider_ngram_key(arr.ctypes.data, len(ids)))
    word2ix = {w: i for i, w in enumerate(ix2word)}
    scorer = scst.CiderDReward(word2ix, df, ref_len)
    gts = dict(enumerate(refs))
    img_ids = [int(i) for i in z["img_index"]]
    gen, greedy = torch.from_numpy(z["gen"]).cuda(), torch.from_numpy(z["greedy"]).cuda()
    rewards = scorer(gen, greedy, gts, img_ids, n_per_image=n)
    scores, flat = scorer.scores_and_rewards(gen, greedy, gts, img_ids, n_per_image=n)
    torch.cuda.synchronize()
    assert rewards.shape == z["rewards"].shape and rewards.dtype == torch.float32
    assert np.abs(rewards.cpu().numpy() - z["rewards"]).max() < 1e-5
    # absolute scores against the oracle (samples then greedy per image)
    ids_rep = [int(i) for i in np.repeat(z["img_index"], n)]
    _, ref_scores = orc.self_critical_reward(z["gen"], np.repeat(z["greedy"], n, axis=0), gts, ids_rep, ix2word, df, ref_len)
    B = len(img_ids)
    assert np.abs(scores[:, :n].reshape(-1).cpu().numpy() - ref_scores[:B * n]).max() < 1e-5
    assert np.abs(scores[:, n].cpu().numpy() - ref_scores[B * n::n]).max() < 1e-5
    assert float(scores.max()) > 1.0  # the fixture is not degenerate
    scorer.close()


@pytest.mark.gpu
def test_device_reward_rejects_bad_input():
    torch = pytest.importorskip("torch")
    ix2word, refs = synth.make_caption_corpus(8, 32, seed=3)
    df, ref_len = scst.document_frequency_from_corpus(refs)
    scorer = scst.CiderDReward({w: i for i, w in enumerate(ix2word)}, df, ref_len)
    gen = torch.zeros((4, 20), dtype=torch.int32).cuda()
    with pytest.raises(ValueError):
        scorer(gen, gen[:3], dict(enumerate(refs)), [0, 1, 2, 3])
    with pytest.raises(ValueError, match="tokens"):
        scorer(gen, gen, {i: [" ".join(["w5"] * 70)] for i in range(4)}, [0, 1, 2, 3])
    scorer.close()


@pytest.mark.gpu
def test_engine_scst_forward_epoch():
    """CaptionEngine.scst_forward_epoch: the forward half of an SCST epoch (prefetch, one-pass rollouts, device reward)
    yields what RewardCriterion consumes, and its rewards equal the oracle's on the rollouts it produced."""
    torch = pytest.importorskip("torch")
    from simpleimagecaptionzoo_b200 import engine
    dims = dict(synth.TINY_DIMS["BUTD"])
    ix2word, refs = synth.make_caption_corpus(12, dims["vocab_size"], seed=4)
    df, ref_len = scst.document_frequency_from_corpus(refs)
    sd = synth.make_state_dict("BUTD", seed=1, chaotic=True, end_boost=1.0, **dims)

    class Vocab:
        def __init__(self, words):
            self.ix2word = dict(enumerate(words))
            self.word2ix = {w: i for i, w in enumerate(words)}

        def __len__(self):
            return len(self.ix2word)

    settings = dict(model_type="BUTDDetection", embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"], atten_dim=dims["atten_dim"])
    eng = engine.BUTDDetection_Eng(settings, "synthetic", Vocab(ix2word), state_dict=sd, max_batch=4, max_regions=6, max_rows=3,
                                   max_seq=20, enc_dim=dims["enc_dim"])
    feats = synth.make_region_feats(12, 6, dims["enc_dim"], 4)
    gts = dict(enumerate(refs))
    loader = [([4 * b + i for i in range(4)], None, gts, [{"bu_feat": feats[4 * b + i], "bu_bbox": None} for i in range(4)]) for b in range(3)]
    reward = scst.CiderDReward(eng.caption_vocab.word2ix, df, ref_len)
    seen = 0
    for img_ids, seq, logprobs, greedy, rewards in eng.scst_forward_epoch(loader, reward, n_per_image=2, max_len=20):
        assert seq.shape == (8, 20) and logprobs.shape == (8, 20) and greedy.shape == (4, 20) and rewards.shape == (8, 20)
        ids_rep = [i for i in img_ids for _ in range(2)]
        want, _ = orc.self_critical_reward(seq.cpu().numpy(), np.repeat(greedy.cpu().numpy(), 2, axis=0), gts, ids_rep, ix2word, df, ref_len)
        assert np.abs(rewards.cpu().numpy() - want).max() < 1e-5
        seen += len(img_ids)
    assert seen == 12
    reward.close()
