"""Host side of the CIDEr-D self-critical reward -- the mirror of
``Utils.get_self_critical_reward`` (Utils.py:319-367), which in the reference turns the two rollouts of an SCST step
into word strings and scores them with the pure-Python ``CiderD`` (cider/pyciderevalcap/ciderD/ciderD_scorer.py) on
the CPU.  Here the rollouts stay on the device as word ids and one kernel of libcapdec.so scores them
(``capdec_cider_reward``, csrc/cider.cuh); the host only

* hashes the training corpus' document-frequency table once (the reference's ``cider/data/<dataset>-train.p``:
  ``{'document_frequency': {ngram tuple of words: df}, 'ref_len': number of training images}``, written by
  PreProcess/CIDEr_idf_preproccess.py) with the same 64-bit n-gram key the kernel uses, and
* maps each image's reference captions to word ids once (out-of-vocabulary words get private ids >= len(vocab), so an
  n-gram of ids is the same n-gram of words) and ships them per batch.

There is no CPU scoring path in this module: without libcapdec.so / a B200 the constructor raises.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import capdec

MAX_REF_TOKENS = 65  # csrc/cider.cuh CIDER_MAX_TOKENS

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(x: np.ndarray) -> np.ndarray:
    x = x ^ (x >> np.uint64(30))
    x = x * _M1
    x = x ^ (x >> np.uint64(27))
    x = x * _M2
    return x ^ (x >> np.uint64(31))


def ngram_keys(ids: np.ndarray) -> np.ndarray:
    """Keys of a batch of n-grams of one length: ids (N, k) int -> (N,) uint64.  Same function as
    ``cider_ngram_key`` in csrc/cider.cuh (checked against ``capdec_cider_ngram_key`` in the tests)."""
    ids = np.ascontiguousarray(ids)
    k = ids.shape[1]
    with np.errstate(over="ignore"):
        h = np.full(ids.shape[0], np.uint64(k) * _GOLD, dtype=np.uint64)
        for i in range(k):
            h = _mix64(h ^ ((ids[:, i].astype(np.uint32).astype(np.uint64) + np.uint64(1)) * _M1))
    h[h == 0] = np.uint64(1)
    return h


class WordIds:
    """word -> id with the caption vocabulary's ids for known words and private ids (>= len(vocab)) for the rest."""

    def __init__(self, word2ix: Mapping[str, int]):
        self.word2ix = dict(word2ix)
        self.extra: Dict[str, int] = {}
        self.base = max(self.word2ix.values()) + 1 if self.word2ix else 0

    def __call__(self, word: str) -> int:
        i = self.word2ix.get(word)
        if i is None:
            i = self.extra.get(word)
            if i is None:
                i = self.extra[word] = self.base + len(self.extra)
        return i

    def sentence(self, s: str) -> np.ndarray:
        return np.fromiter((self(w) for w in s.split()), dtype=np.int32)


def document_frequency_from_corpus(refs_by_image: Iterable[Sequence[str]], n: int = 4) -> Tuple[dict, int]:
    """PreProcess/CIDEr_idf_preproccess.py:41-66: df[ngram] = number of images whose references contain it."""
    df: Dict[tuple, float] = {}
    count = 0
    for refs in refs_by_image:
        seen = set()
        for ref in refs:
            words = ref.split()
            for k in range(1, n + 1):
                for i in range(len(words) - k + 1):
                    seen.add(tuple(words[i:i + k]))
        for g in seen:
            df[g] = df.get(g, 0.0) + 1.0
        count += 1
    return df, count


class CiderDReward:
    """``rewards = CiderDReward(...)(gen_result, greedy_res, ground_truth, img_ids)`` -- the call
    ``get_self_critical_reward(gen_result, greedy_res, ground_truth, img_ids, caption_vocab, dataset_name)`` makes, with
    device tensors in and a device tensor out."""

    def __init__(self, word2ix: Mapping[str, int], document_frequency: Mapping[tuple, float], ref_len: float, *,
                 sigma: float = 6.0, cider_weight: float = 1.0, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("CiderDReward needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = capdec.load_library()
        self.device = torch.device("cuda", device)
        self.sigma, self.weight = float(sigma), float(cider_weight)
        self.ids = WordIds(word2ix)
        self._ref_cache: Dict[object, Tuple[np.ndarray, ...]] = {}
        h = ctypes.c_void_p()
        rc = self.lib.capdec_cider_create(device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"capdec_cider_create failed ({rc}): {self.lib.capdec_cider_last_error(None).decode()}")
        self._h = h
        self._set_df(document_frequency, ref_len)

    @classmethod
    def from_pickle(cls, path: str, word2ix, **kw):
        """The reference's ``cider/data/<dataset>-train.p`` (ciderD_scorer.py:79-83)."""
        import pickle
        with open(path, "rb") as f:
            p = pickle.load(f, encoding="latin1")
        return cls(word2ix, p["document_frequency"], p["ref_len"], **kw)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.capdec_cider_last_error(self._h).decode()}")

    def _set_df(self, document_frequency, ref_len):
        by_len: Dict[int, list] = {}
        vals: Dict[int, list] = {}
        for gram, v in document_frequency.items():
            k = len(gram)
            by_len.setdefault(k, []).append([self.ids(w) for w in gram])
            vals.setdefault(k, []).append(float(v))
        keys = [ngram_keys(np.asarray(by_len[k], dtype=np.int64)) for k in sorted(by_len)]
        dfs = [np.asarray(vals[k], dtype=np.float32) for k in sorted(by_len)]
        keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
        dfs = np.concatenate(dfs) if dfs else np.zeros(0, np.float32)
        if len(np.unique(keys)) != len(keys):
            raise RuntimeError("64-bit n-gram key collision in the document-frequency table")
        self._check(self.lib.capdec_cider_set_df(self._h, keys.ctypes.data_as(ctypes.c_void_p), dfs.ctypes.data_as(ctypes.c_void_p),
                                                 len(keys), math.log(float(ref_len))), "capdec_cider_set_df")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.capdec_cider_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ references
    def encode_refs(self, ground_truth: Mapping[object, Sequence[str]], img_ids: Sequence[object]):
        """-> (ref_tokens [n_refs, 65] int32, ref_lens [n_refs] int32, ref_offsets [B+1] int32) for the batch.  Each image's
        references are mapped to ids once and cached as a padded block; a batch is a concatenation of blocks."""
        blocks, lens, counts = [], [], []
        for i in img_ids:
            enc = self._ref_cache.get(i)
            if enc is None:
                sents = [self.ids.sentence(s) for s in ground_truth[i]]
                for e in sents:
                    if len(e) > MAX_REF_TOKENS:
                        raise ValueError(f"reference caption of image {i} has {len(e)} tokens (> {MAX_REF_TOKENS})")
                blk = np.zeros((len(sents), MAX_REF_TOKENS), np.int32)
                for r, e in enumerate(sents):
                    blk[r, :len(e)] = e
                enc = self._ref_cache[i] = (blk, np.asarray([len(e) for e in sents], np.int32))
            blocks.append(enc[0])
            lens.append(enc[1])
            counts.append(len(enc[1]))
        offs = np.zeros(len(counts) + 1, np.int32)
        np.cumsum(counts, out=offs[1:])
        return np.concatenate(blocks), np.concatenate(lens), offs

    # ------------------------------------------------------------------ reward
    def scores_and_rewards(self, gen_result, greedy_res, ground_truth, img_ids, n_per_image: Optional[int] = None):
        import torch
        B = greedy_res.shape[0]
        n = n_per_image or gen_result.shape[0] // B
        if gen_result.shape[0] != B * n or len(img_ids) != B:
            raise ValueError("gen_result must hold n_per_image rows per row of greedy_res / entry of img_ids")
        T = gen_result.shape[1]
        gen = gen_result.to(self.device, torch.int32).contiguous()
        greedy = greedy_res.to(self.device, torch.int32).contiguous()
        if greedy.shape[1] != T:
            raise ValueError("gen_result and greedy_res must have the same max_len")
        tok, lens, offs = self.encode_refs(ground_truth, img_ids)
        # pinned staging: the upload is asynchronous and does not drain the stream the rollouts were enqueued on
        d_tok, d_lens, d_offs = (torch.from_numpy(a).pin_memory().to(self.device, non_blocking=True) for a in (tok, lens, offs))
        rewards = torch.empty((B * n,), dtype=torch.float32, device=self.device)
        scores = torch.empty((B, n + 1), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.capdec_cider_reward(
                self._h, gen.data_ptr(), n, greedy.data_ptr(), B, T, d_tok.data_ptr(), d_lens.data_ptr(), d_offs.data_ptr(),
                tok.shape[1], self.sigma, self.weight, rewards.data_ptr(), scores.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream), "capdec_cider_reward")
        self._keep = (gen, greedy, d_tok, d_lens, d_offs)
        return scores, rewards

    def __call__(self, gen_result, greedy_res, ground_truth, img_ids, n_per_image: Optional[int] = None):
        """-> rewards (B*n, max_len) float32 on the device: the per-sequence reward repeated over the time steps, as
        ``get_self_critical_reward`` returns it (Utils.py:364-365) for ``RewardCriterion`` (Utils.py:290-317)."""
        _, r = self.scores_and_rewards(gen_result, greedy_res, ground_truth, img_ids, n_per_image)
        return r[:, None].expand(-1, gen_result.shape[1]).contiguous()


def differentiable_logprobs(states, tokens, weight_g, weight_v, bias):
    """``seqLogprobs`` WITH an autograd graph through the vocabulary layer, for ``RewardCriterion`` (Utils.py:290-317) on
    rollouts produced by the fused decoder: ``states`` [M, T, H] are the rows ``predict`` saw at every step
    (``CaptionDecoder.score(tokens, n, return_states=True)``), ``tokens`` [M, T] the rollout, ``weight_g / weight_v / bias``
    the LIVE parameters of the reference's weight-normed ``decoder.predict`` (BUTD_Model.py:84).  Returns [M, T] fp32
    ``log_softmax(predict(state))[token]`` -- forward values equal to ``score``'s, gradients flow into the three parameters.
    Gradients of the recurrent weights need back-propagation through the decode loop: that is the reference's training
    path, outside this library."""
    import torch
    W = weight_v * (weight_g / weight_v.norm(dim=1, keepdim=True))  # torch.nn.utils.weight_norm, dim=0
    out = []
    for t in range(states.shape[1]):  # one step at a time: [M, V] logits live only for the step's backward
        logp = torch.log_softmax(torch.addmm(bias, states[:, t], W.t()), dim=1)
        out.append(logp.gather(1, tokens[:, t:t + 1].long()))
    return torch.cat(out, dim=1)
