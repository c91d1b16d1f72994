"""ctypes binding of libcapdec.so (include/capdec.h) and the tensor-level decoder object.

PyTorch is plumbing here: device memory, streams, dtype conversion.  Every decode step runs inside the C
library as hand-written sm_100a kernels; there is NO CPU or eager-PyTorch fallback -- if the shared library
is missing or no B200 is present the constructors raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Mapping, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CAPDEC_LIB selects another build of the same library (A/B timing of two kernel versions on one GPU box)
LIB_PATH = os.environ.get("CAPDEC_LIB") or os.path.join(_HERE, "libcapdec.so")

ARCH = {"NIC": 0, "BUTD": 1, "AOA": 2}
MATH = {"f16": 0, "f16x3": 1}
SAMPLE_GREEDY, SAMPLE_MULTINOMIAL = 0, 1

# every symbol include/capdec.h declares (tests check that the built library exports all of them)
SYMBOLS = (
    "capdec_abi_version", "capdec_create", "capdec_destroy", "capdec_last_error", "capdec_load_weight",
    "capdec_finalize_weights", "capdec_prepare", "capdec_beam_search", "capdec_sample", "capdec_launch_count",
    "capdec_test_gemm", "capdec_profile", "capdec_profile_read", "capdec_test_gemm_time",
    "capdec_prepare_bottom_up", "capdec_get_refined", "capdec_score", "capdec_prepare_f16", "capdec_scst_rollout",
    "capdec_cider_create", "capdec_cider_destroy", "capdec_cider_last_error", "capdec_cider_ngram_key", "capdec_cider_set_df",
    "capdec_cider_reward", "capdec_graph_captures", "capdec_debug_trace", "capdec_score_states",
)
CATEGORIES = ("gemm_lstm", "gemm_store", "gemm_glu", "gemm_logits", "attention", "bookkeeping", "other")


class CapdecConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "arch", "hidden_dim", "embed_dim", "atten_dim", "enc_dim", "vocab_size", "num_heads", "max_batch",
        "max_regions", "max_rows", "max_seq", "math_mode", "device")]


_lib = None


def load_library(path: str = LIB_PATH) -> ctypes.CDLL:
    """dlopen libcapdec.so and declare the prototypes.  Raises if the library has not been built
    (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build the CUDA extension first (__graft_entry__.build()); "
                           "there is no CPU fallback for the caption decoder")
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    lib.capdec_abi_version.restype = ctypes.c_int
    lib.capdec_create.argtypes = [ctypes.POINTER(CapdecConfig), ctypes.POINTER(vp)]
    lib.capdec_destroy.argtypes = [vp]
    lib.capdec_destroy.restype = None
    lib.capdec_last_error.argtypes = [vp]
    lib.capdec_last_error.restype = ctypes.c_char_p
    lib.capdec_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32, vp]
    lib.capdec_finalize_weights.argtypes = [vp, vp]
    lib.capdec_prepare.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.capdec_prepare_bottom_up.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.capdec_prepare_f16.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.capdec_get_refined.argtypes = [vp, vp, vp]
    lib.capdec_beam_search.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.capdec_sample.argtypes = [vp, i32, i32, ctypes.c_uint64, i32, vp, vp, vp, vp]
    lib.capdec_score.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.capdec_score_states.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.capdec_scst_rollout.argtypes = [vp, i32, ctypes.c_uint64, i32, vp, vp, vp, vp]
    lib.capdec_cider_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.capdec_cider_destroy.argtypes = [vp]
    lib.capdec_cider_destroy.restype = None
    lib.capdec_cider_last_error.argtypes = [vp]
    lib.capdec_cider_last_error.restype = ctypes.c_char_p
    lib.capdec_cider_ngram_key.argtypes = [vp, i32]
    lib.capdec_cider_ngram_key.restype = ctypes.c_uint64
    lib.capdec_cider_set_df.argtypes = [vp, vp, vp, i64, ctypes.c_double]
    lib.capdec_cider_reward.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp, vp, i32, ctypes.c_double, ctypes.c_double, vp, vp, vp]
    lib.capdec_launch_count.argtypes = [vp]
    lib.capdec_launch_count.restype = i64
    lib.capdec_graph_captures.argtypes = [vp]
    lib.capdec_graph_captures.restype = i64
    lib.capdec_debug_trace.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), i64]
    lib.capdec_test_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.capdec_profile.argtypes = [vp, i32]
    lib.capdec_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(i64)]
    lib.capdec_test_gemm_time.argtypes = [i32, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_float)]
    if lib.capdec_abi_version() != 1:
        raise RuntimeError("libcapdec.so ABI version mismatch")
    if path == LIB_PATH:
        _lib = lib
    return lib


def _torch():
    import torch
    return torch


def _stream_ptr(device) -> int:
    torch = _torch()
    return torch.cuda.current_stream(device).cuda_stream


def test_gemm(a, b, bias=None, math: str = "f16"):
    """D = A @ B.T (+ bias) through the library's tcgen05 GEMM (test hook).  a [M,K], b [N,K] fp32 CUDA tensors."""
    torch = _torch()
    lib = load_library()
    a = a.contiguous().float()
    b = b.contiguous().float()
    m, k = a.shape
    n = b.shape[0]
    d = torch.empty((m, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.capdec_test_gemm(a.data_ptr(), b.data_ptr(), bias.data_ptr() if bias is not None else None, d.data_ptr(),
                                  m, n, k, MATH[math], _stream_ptr(a.device))
    if rc != 0:
        raise RuntimeError(f"capdec_test_gemm failed ({rc}): {lib.capdec_last_error(None).decode()}")
    return d


def gemm_time_us(m: int, n: int, k: int, epi: int = 0, math: str = "f16", iters: int = 20) -> float:
    """Mean device time of one GEMM launch of the given shape / epilogue (test + tuning hook)."""
    lib = load_library()
    us = ctypes.c_float()
    rc = lib.capdec_test_gemm_time(m, n, k, epi, MATH[math], iters, ctypes.byref(us))
    if rc != 0:
        raise RuntimeError(f"capdec_test_gemm_time failed ({rc}): {lib.capdec_last_error(None).decode()}")
    return float(us.value)


class CaptionDecoder:
    """Batched beam / greedy / multinomial caption decoder for one architecture on one GPU.

    ``state_dict`` is the reference checkpoint's mapping (torch tensors or numpy arrays); entries whose key
    starts with ``decoder.`` (or that carry no prefix at all) are handed to the library unchanged -- weight-norm
    folding, gate interleaving and fp16 packing happen inside (capdec_finalize_weights).
    """

    def __init__(self, arch: str, state_dict: Mapping[str, object], *, hidden_dim: int, embed_dim: int, vocab_size: int,
                 atten_dim: int = 0, enc_dim: int = 2048, num_heads: int = 8, max_batch: int = 64, max_regions: int = 36,
                 max_rows: int = 3, max_seq: int = 20, math: str = "f16", device: int = 0):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("CaptionDecoder needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.arch = arch.upper()
        self.device = torch.device("cuda", device)
        self.V, self.H, self.E = vocab_size, hidden_dim, embed_dim
        self.max_batch, self.max_rows, self.max_seq, self.max_regions = max_batch, max_rows, max_seq, max_regions
        self.math = math
        cfg = CapdecConfig(ARCH[self.arch], hidden_dim, embed_dim, atten_dim, enc_dim if self.arch != "NIC" else 0,
                           vocab_size, num_heads if self.arch == "AOA" else 0, max_batch,
                           0 if self.arch == "NIC" else max_regions, max_rows, max_seq, MATH[math], device)
        handle = ctypes.c_void_p()
        rc = self.lib.capdec_create(ctypes.byref(cfg), ctypes.byref(handle))
        if rc != 0:
            raise RuntimeError(f"capdec_create failed ({rc}): {self.lib.capdec_last_error(None).decode()}")
        self._h = handle
        # decode work runs on an own (capturable) stream, ordered after / before the caller's current stream
        self.stream = torch.cuda.Stream(self.device)
        self._keep = None
        self.B, self.R = 0, 0
        self.has_refiner = False
        self._load(state_dict)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.capdec_last_error(self._h).decode()}")

    def _load(self, state_dict):
        torch = _torch()
        stream = _stream_ptr(self.device)
        n = 0
        with torch.cuda.device(self.device):
            for key, val in state_dict.items():
                if key.startswith("decoder."):
                    name = key[len("decoder."):]
                elif self.arch == "AOA" and key.split(".")[0] in ("img_feats_porjection", "aoa_refine"):
                    name = key  # AoA encoder side (projection + refiner): runs in the library too (prepare_bottom_up)
                    self.has_refiner = True
                elif "." in key and key.split(".")[0] in ("encoder", "img_feats_porjection", "aoa_refine"):
                    continue  # CNN encoder entries of the full captioner checkpoint are not part of this path
                else:
                    name = key
                if isinstance(val, np.ndarray):
                    t = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32))
                else:
                    t = val.detach().to(torch.float32).contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                self._check(self.lib.capdec_load_weight(self._h, name.encode(), t.data_ptr(), shape, t.dim(), stream),
                            f"capdec_load_weight({name})")
                n += 1
            self._check(self.lib.capdec_finalize_weights(self._h, stream), "capdec_finalize_weights")
        return n

    def close(self):
        if getattr(self, "_h", None):
            self.lib.capdec_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.capdec_launch_count(self._h))

    @property
    def graph_captures(self) -> int:
        return int(self.lib.capdec_graph_captures(self._h))

    def debug_trace(self, max_launches: int = 4000):
        """CAPDEC_TRACE=1: -> int64 array [launches, 16] of %globaltimer stamps (ns) of the small-batch kernel's CTA 0."""
        n = 1 + 16 * max_launches
        buf = (ctypes.c_uint64 * n)()
        self._check(self.lib.capdec_debug_trace(self._h, buf, n), "capdec_debug_trace")
        a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
        k = min(int(a[0]), max_launches)
        return a[1:1 + 16 * k].reshape(k, 16)

    def profile(self, enable: bool):
        self._check(self.lib.capdec_profile(self._h, 1 if enable else 0), "capdec_profile")

    def profile_read(self):
        """-> {category: (device ms, algorithmic GEMM flops, launches)} since the last read (waits for the events)."""
        n = len(CATEGORIES)
        ms, fl, cnt = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
        self._check(self.lib.capdec_profile_read(self._h, ms, fl, cnt), "capdec_profile_read")
        return {c: (ms[i], fl[i], int(cnt[i])) for i, c in enumerate(CATEGORIES)}

    # ------------------------------------------------------------------ decode API (device tensors)
    def prepare(self, feats, mask=None):
        """feats: CUDA fp32 [B,R,D] (BUTD), [B,R,H] refined (AoA) or [B,E] (NIC); mask: [B,R] float or None."""
        torch = _torch()
        f16 = feats.dtype == torch.float16 and feats.dim() == 3  # packed fp16 feature shards (feature_store.py)
        feats = feats.to(self.device, torch.float16 if f16 else torch.float32).contiguous()
        if mask is not None:
            mask = mask.to(self.device, torch.float32).contiguous()
        B = feats.shape[0]
        R = feats.shape[1] if feats.dim() == 3 else 0
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            if f16:
                self._check(self.lib.capdec_prepare_f16(self._h, feats.data_ptr(), None if mask is None else mask.data_ptr(), B, R,
                                                        self.stream.cuda_stream), "capdec_prepare_f16")
            else:
                self._check(self.lib.capdec_prepare(self._h, feats.data_ptr(), None if mask is None else mask.data_ptr(), B, R,
                                                    self.stream.cuda_stream), "capdec_prepare")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self._keep = (feats, mask)  # the library reads them during decode
        self.B, self.R = B, R

    def prepare_bottom_up(self, bu_feats, mask=None):
        """AoA only: bu_feats CUDA fp32 [B,R,enc_dim] (bottom-up / CNN grid features), mask [B,R] float prefix mask or
        None -> img_feats_porjection + aoa_refine + prepare, all inside the library (AoA_Model.py:748-751)."""
        torch = _torch()
        if not self.has_refiner:
            raise RuntimeError("prepare_bottom_up needs the checkpoint's img_feats_porjection.* / aoa_refine.* entries")
        if bu_feats.dtype == torch.float16:  # packed fp16 shard rows: same path, no conversion pass
            return self.prepare(bu_feats, mask)
        bu_feats = bu_feats.to(self.device, torch.float32).contiguous()
        if mask is not None:
            mask = mask.to(self.device, torch.float32).contiguous()
        B, R = bu_feats.shape[0], bu_feats.shape[1]
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self._check(self.lib.capdec_prepare_bottom_up(self._h, bu_feats.data_ptr(), None if mask is None else mask.data_ptr(),
                                                          B, R, self.stream.cuda_stream), "capdec_prepare_bottom_up")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self._keep = (bu_feats, mask)
        self.B, self.R = B, R

    def refined_features(self):
        """Refined features [B,R,H] fp32 of the batch prepared by ``prepare_bottom_up`` (what aoa_refine returns)."""
        torch = _torch()
        out = torch.empty((self.B, self.R, self.H), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self._check(self.lib.capdec_get_refined(self._h, out.data_ptr(), self.stream.cuda_stream), "capdec_get_refined")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return out

    def beam_search(self, beam: int, max_seq: int = 20, return_alphas: bool = False):
        """-> tokens [B,1+max_seq] int32 (<sta> first), seq_logprob [B] fp32, lengths [B] int32 (CUDA tensors)
        [, alphas [B,max_seq,R] fp32 attention maps of the returned hypotheses]."""
        torch = _torch()
        B = self.B
        if return_alphas and self.arch == "NIC":
            raise RuntimeError("NIC has no attention maps (its decoder has no attention); call without return_alphas")
        tokens = torch.empty((B, 1 + max_seq), dtype=torch.int32, device=self.device)
        scores = torch.empty((B,), dtype=torch.float32, device=self.device)
        lengths = torch.empty((B,), dtype=torch.int32, device=self.device)
        alphas = torch.empty((B, max_seq, self.R), dtype=torch.float32, device=self.device) if return_alphas else None
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self._check(self.lib.capdec_beam_search(self._h, beam, max_seq, tokens.data_ptr(), scores.data_ptr(),
                                                    lengths.data_ptr(), alphas.data_ptr() if return_alphas else None,
                                                    self.stream.cuda_stream), "capdec_beam_search")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return (tokens, scores, lengths, alphas) if return_alphas else (tokens, scores, lengths)

    def scst_rollout(self, n_per_image: int = 1, seed: int = 0, max_seq: int = 20):
        """Both rollouts of an SCST step in one pass -> (sample tokens [B*n,T] int32, sample logprobs [B*n,T] fp32, greedy
        tokens [B,T] int32); row for row what ``sample(MULTINOMIAL, n, seed)`` and ``sample(GREEDY, 1)`` return."""
        torch = _torch()
        M = self.B * n_per_image
        tokens = torch.empty((M, max_seq), dtype=torch.int32, device=self.device)
        logprobs = torch.empty((M, max_seq), dtype=torch.float32, device=self.device)
        greedy = torch.empty((self.B, max_seq), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self._check(self.lib.capdec_scst_rollout(self._h, n_per_image, seed, max_seq, tokens.data_ptr(), logprobs.data_ptr(),
                                                     greedy.data_ptr(), self.stream.cuda_stream), "capdec_scst_rollout")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return tokens, logprobs, greedy

    def score(self, tokens, n_per_image: int = 1, return_states: bool = False):
        """Teacher-forced log-probs of given words: tokens [B*n, T] int (no <sta>, the layout ``sample`` returns) ->
        logprobs [B*n, T] fp32 CUDA tensor, log p(word t | image, <sta>, words < t)
        [, states [B*n, T, H] fp32: the rows ``predict`` was applied to -- see ``scst.differentiable_logprobs``]."""
        torch = _torch()
        tokens = tokens.to(self.device, torch.int32).contiguous()
        M, T = tokens.shape
        if M != self.B * n_per_image:
            raise ValueError(f"tokens has {M} rows, the prepared batch has {self.B} images x {n_per_image}")
        logprobs = torch.empty((M, T), dtype=torch.float32, device=self.device)
        states = torch.empty((M, T, self.H), dtype=torch.float32, device=self.device) if return_states else None
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            if return_states:
                self._check(self.lib.capdec_score_states(self._h, tokens.data_ptr(), n_per_image, T, logprobs.data_ptr(),
                                                         states.data_ptr(), self.stream.cuda_stream), "capdec_score_states")
            else:
                self._check(self.lib.capdec_score(self._h, tokens.data_ptr(), n_per_image, T, logprobs.data_ptr(),
                                                  self.stream.cuda_stream), "capdec_score")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self._keep_tokens = tokens
        return (logprobs, states) if return_states else logprobs

    def sample(self, mode: int, n_per_image: int = 1, seed: int = 0, max_seq: int = 20, return_alphas: bool = False):
        """-> tokens [B*n,max_seq] int32, logprobs [B*n,max_seq] fp32 (CUDA tensors) [, alphas [B*n,max_seq,R] fp32]."""
        torch = _torch()
        M = self.B * n_per_image
        if return_alphas and self.arch == "NIC":
            raise RuntimeError("NIC has no attention maps (its decoder has no attention); call without return_alphas")
        tokens = torch.empty((M, max_seq), dtype=torch.int32, device=self.device)
        logprobs = torch.empty((M, max_seq), dtype=torch.float32, device=self.device)
        alphas = torch.zeros((M, max_seq, self.R), dtype=torch.float32, device=self.device) if return_alphas else None
        with torch.cuda.device(self.device):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self._check(self.lib.capdec_sample(self._h, mode, n_per_image, seed, max_seq, tokens.data_ptr(), logprobs.data_ptr(),
                                               alphas.data_ptr() if return_alphas else None, self.stream.cuda_stream),
                        "capdec_sample")
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return (tokens, logprobs, alphas) if return_alphas else (tokens, logprobs)
