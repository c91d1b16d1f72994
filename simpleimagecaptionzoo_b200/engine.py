"""Host-side mirror of the reference's Engine / captioner API for the decode path.

What a user of zyj0021200/simpleImageCaptionZoo calls, with the same names, argument meaning and outputs:

* ``B200Captioner.sampler(visual_inputs, max_len=20)``          -> (B, max_len) long          BUTD_Model.py:478-489
* ``B200Captioner.sampler_rl(visual_inputs, max_len=20)``       -> ((B,T) long, (B,T) float)  BUTD_Model.py:491-503
* ``B200Captioner.beam_search_sampler(visual_inputs, beam_size)`` -> (B, 1+max_seq) long      BUTD_Model.py:505-516
  (same for Models/NIC_Model.py:266-304 and Models/AoA_Model.py:700-753)
* ``CaptionEngine.eval_captions_json_generation(dataloader, eval_beam_size)``                  Engine.py:274-300
* ``*_Eng.modify_visual_inputs``                                 ModelEngines/{NIC,BUTD,AoA}_Engine.py

Differences, all deliberate: beam search is BATCHED (the reference forces one image per call,
Utils.py:72-73); the beam step limit is ``max_seq`` (default 20, BASELINE.json) instead of the hard-coded 50;
``beam_search_sampler`` always returns an int64 device tensor padded with <pad>=0 after <end> (the reference returns a
float32 CPU tensor of ragged length when a hypothesis completed, BUTD_Model.py:309) -- ``Engine.py:288-296`` stops at
<end> and skips <sta>, so the id->word loop downstream is unchanged.

The decode loop itself runs in libcapdec.so (``capdec.CaptionDecoder``), and so does the AoA captioners' encoder-side
half (``img_feats_porjection`` + ``aoa_refine``, AoA_Model.py:748-751) whenever the checkpoint carries its entries.  The
CNN encoders (ResNet-101) stay outside this path and are taken as a plain callable ``feature_fn(visual_inputs)``.
"""
from __future__ import annotations

import json
from typing import Callable, Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import capdec

MODEL_ARCH = {  # Main.py:48-62 model_type -> decoder family
    "NIC": "NIC",
    "BUTDSpatial": "BUTD",
    "BUTDDetection": "BUTD",
    "AoASpatial": "AOA",
    "AoADetection": "AOA",
}
END_WORD, STA_WORD = "<end>", "<sta>"


def _torch():
    import torch
    return torch


class B200Captioner:
    """Extension-backed stand-in for the reference's ``*_Captioner`` modules on the decode path."""

    def __init__(self, model_type: str, settings: Mapping[str, object], vocab_size: int, state_dict: Mapping[str, object], *,
                 feature_fn: Optional[Callable] = None, feature_fn_returns: str = "decoder_input", max_batch: int = 64,
                 max_regions: Optional[int] = None,
                 max_rows: int = 5, max_seq: int = 20, math: str = "f16", device: int = 0, enc_dim: int = 2048,
                 num_heads: int = 8, sample_seed: int = 0):
        if model_type not in MODEL_ARCH:
            raise ValueError(f"unknown model_type {model_type!r}")
        self.model_type = model_type
        self.arch = MODEL_ARCH[model_type]
        self.settings = dict(settings)
        self.feature_fn = feature_fn
        self.max_seq = max_seq
        self._seed = sample_seed
        self._calls = 0
        if max_regions is None:
            # Spatial: the CNN grid; AoADetection: adaptive bottom-up features carry 10-100 boxes per image
            # (AoA_Engine.py:23-47 pads them and builds bu_masks); BUTDDetection: the fixed 36 boxes
            s = int(self.settings.get("enc_img_size", 0) or 0)
            max_regions = s * s if model_type.endswith("Spatial") and s else (100 if model_type == "AoADetection" else 36)
        H, E = int(self.settings["hidden_dim"]), int(self.settings["embed_dim"])
        self.decoder = capdec.CaptionDecoder(
            self.arch, state_dict, hidden_dim=H, embed_dim=E, vocab_size=vocab_size,
            atten_dim=int(self.settings.get("atten_dim", 0) or 0), enc_dim=enc_dim, num_heads=num_heads,
            max_batch=max_batch, max_regions=max_regions, max_rows=max_rows, max_seq=max_seq, math=math, device=device)
        self.device = self.decoder.device
        # AoA: run img_feats_porjection + aoa_refine in the library when the checkpoint has them and the features handed
        # over are the bottom-up / CNN-grid ones (``feature_fn_returns="bottom_up"``, or no feature_fn for AoADetection)
        if feature_fn_returns not in ("decoder_input", "bottom_up"):
            raise ValueError("feature_fn_returns must be 'decoder_input' or 'bottom_up'")
        self.native_refiner = (self.arch == "AOA" and self.decoder.has_refiner
                               and (feature_fn is None or feature_fn_returns == "bottom_up"))
        if self.arch == "AOA" and feature_fn is None and not self.decoder.has_refiner and model_type == "AoADetection":
            raise RuntimeError("AoADetection: the checkpoint has no img_feats_porjection / aoa_refine entries; pass a "
                               "feature_fn that returns the refined features")

    def _prepare(self, feats, mask):
        if self.native_refiner:
            self.decoder.prepare_bottom_up(feats, mask)
        else:
            self.decoder.prepare(feats, mask)

    # nn.Module surface the Engine touches
    def eval(self):
        return self

    def train(self, mode: bool = True):
        return self

    def to(self, device):
        return self

    # ------------------------------------------------------------------ inputs
    def _features(self, visual_inputs, to_device: bool = True):
        """Pull the decoder's input out of ``visual_inputs`` the way the reference wrappers do
        (BUTD_Model.py:486,513; AoA_Model.py:748-751; NIC_Model.py:277-279)."""
        torch = _torch()
        mask = None
        if self.feature_fn is not None:
            out = self.feature_fn(visual_inputs)
            feats, mask = out if isinstance(out, tuple) else (out, None)
        elif self.model_type in ("BUTDDetection", "AoADetection"):
            feats = visual_inputs["bu_feats"]
            if self.arch == "AOA":
                mask = visual_inputs.get("bu_masks")
        else:
            raise RuntimeError(f"{self.model_type} needs feature_fn (its CNN encoder / refiner is outside the decode path)")
        if isinstance(feats, np.ndarray):
            feats = torch.from_numpy(feats)
        if isinstance(mask, np.ndarray):
            mask = torch.from_numpy(mask)
        if to_device:
            feats = feats.to(self.device, non_blocking=True)
            if mask is not None:
                mask = mask.to(self.device, non_blocking=True)
        return feats, mask

    # ------------------------------------------------------------------ the three decode methods
    def beam_search_sampler(self, visual_inputs, beam_size: int = 5, max_seq: Optional[int] = None):
        feats, mask = self._features(visual_inputs)
        self._prepare(feats, mask)
        tokens, self.last_scores, self.last_lengths = self.decoder.beam_search(beam_size, max_seq or self.max_seq)
        return tokens.long()

    def beam_search_stream(self, batches, beam_size: int = 5, max_seq: Optional[int] = None, slots: int = 3,
                           on_device_tokens: Optional[Callable] = None):
        """Pipelined form of ``beam_search_sampler`` for a sequence of batches (what an evaluation loop feeds):
        yields one HOST int32 array [B, 1+max_seq] per input batch, in order.  The host->device copy of batch i+1
        runs on a copy stream while batch i decodes, and the captions of batch i are read back (pinned buffer,
        asynchronous) while batch i+1 is already enqueued -- the GPU never waits for the host.
        ``batches`` yields ``visual_inputs`` dicts whose feature tensors may live on the host (ideally pinned).
        ``slots`` device staging buffers are used in turn (>= 2); ``on_device_tokens(tokens)`` is called with every
        batch's DEVICE caption tensor right after its decode was enqueued (multi-GPU: ``CaptionGather.add``)."""
        import collections
        torch = _torch()
        T = max_seq or self.max_seq
        main = torch.cuda.current_stream(self.device)
        slots = max(2, int(slots))
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._slots = []
        while len(self._slots) < slots:
            self._slots.append(dict(buf=None, mask=None, free=None, ready=None))
        copy = self._copy_stream

        def stage(i, visual_inputs):
            slot = self._slots[i % slots]
            feats, mask = self._features(visual_inputs, to_device=False)
            with torch.cuda.stream(copy):
                if slot["free"] is not None:
                    copy.wait_event(slot["free"])  # the decode that last read this buffer has finished
                if feats.is_cuda:
                    slot["buf"] = feats
                else:
                    if slot["buf"] is None or slot["buf"].shape != feats.shape or slot["buf"].dtype != feats.dtype:
                        slot["buf"] = torch.empty(feats.shape, dtype=feats.dtype, device=self.device)
                    slot["buf"].copy_(feats, non_blocking=True)
                slot["mask"] = None if mask is None else mask.to(self.device, non_blocking=True)
                slot["ready"] = torch.cuda.Event()
                slot["ready"].record(copy)
                if isinstance(visual_inputs, dict) and callable(visual_inputs.get("_on_copied")):
                    visual_inputs["_on_copied"](slot["ready"])  # the producer may reuse its host buffer after this event
            return slot

        it = iter(batches)
        pending = collections.deque()
        try:
            nxt = stage(0, next(it))
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            cur = nxt
            try:
                nxt = stage(i + 1, next(it))
            except StopIteration:
                nxt = None
            main.wait_event(cur["ready"])
            self._prepare(cur["buf"], cur["mask"])
            tokens, _, _ = self.decoder.beam_search(beam_size, T)
            cur["free"] = torch.cuda.Event()
            cur["free"].record(main)
            if on_device_tokens is not None:
                on_device_tokens(tokens)
            host = torch.empty(tokens.shape, dtype=tokens.dtype, pin_memory=True)
            host.copy_(tokens, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            pending.append((host, done))
            if len(pending) > 1:
                h, e = pending.popleft()
                e.synchronize()
                yield h.numpy()
            i += 1
        while pending:
            h, e = pending.popleft()
            e.synchronize()
            yield h.numpy()

    def prefetch_to_device(self, batches):
        """Generator over ``visual_inputs`` dicts whose host tensors (ideally pinned) are replaced by device tensors, the copy
        of batch i+1 running on a copy stream while the caller works on batch i -- what an SCST training loop needs around
        its two rollouts (``Engine.SCST_training_epoch`` moves every batch synchronously, Engine.py:254-255).  The device
        buffers are two alternating slots: use a batch before asking for the one after the next."""
        torch = _torch()
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._slots = []
        copy = self._copy_stream
        bufs, free = [dict(), dict()], [None, None]

        def stage(i, vi):
            out = dict(vi)
            with torch.cuda.stream(copy):
                if free[i % 2] is not None:
                    copy.wait_event(free[i % 2])
                for k, v in vi.items():
                    if isinstance(v, np.ndarray) and v.dtype.kind == "f":
                        v = torch.from_numpy(v)
                    if torch.is_tensor(v) and not v.is_cuda:
                        b = bufs[i % 2].get(k)
                        if b is None or b.shape != v.shape or b.dtype != v.dtype:
                            b = bufs[i % 2][k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                        b.copy_(v, non_blocking=True)
                        out[k] = b
                ev = torch.cuda.Event()
                ev.record(copy)
            return out, ev

        it = iter(batches)
        try:
            nxt = stage(0, next(it))
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = stage(i + 1, next(it))
            except StopIteration:
                nxt = None
            main.wait_event(ev)
            yield cur
            free[i % 2] = torch.cuda.Event()
            free[i % 2].record(main)  # everything the caller enqueued on this batch precedes the slot's next overwrite
            i += 1

    def sampler(self, visual_inputs, max_len: int = 20):
        feats, mask = self._features(visual_inputs)
        self._prepare(feats, mask)
        tokens, _ = self.decoder.sample(capdec.SAMPLE_GREEDY, 1, 0, max_len)
        return tokens.long()

    def eval_test_image(self, visual_inputs, caption_vocab, max_len: int = 20, eval_beam_size: int = -1):
        """Single-image test path (``eval_test_image``, BUTD_Model.py:518-544 / AoA_Model.py:755-786 / NIC_Model.py:306-332):
        -> (caption words, [alphas]) with alphas (1, n_words, R) -- the attention map per generated word that
        ``show_additional_rlt`` plots -- or (caption, []) for NIC."""
        feats, mask = self._features(visual_inputs)
        assert feats.shape[0] == 1
        self._prepare(feats, mask)
        want = self.arch != "NIC"
        if eval_beam_size != -1:
            out = self.decoder.beam_search(eval_beam_size, self.max_seq, return_alphas=want)
            tokens, lengths = out[0], out[2]
            n_words = int(lengths[0]) - 1
        else:
            out = self.decoder.sample(capdec.SAMPLE_GREEDY, 1, 0, max_len, return_alphas=want)
            tokens, n_words = out[0], max_len
        caption = []
        for word_id in tokens[0].cpu().numpy():
            word = caption_vocab.ix2word[int(word_id)]
            if word == END_WORD:
                break
            if word != STA_WORD:
                caption.append(word)
        return caption, ([out[-1][:, :n_words]] if want else [])

    def sampler_rl(self, visual_inputs, max_len: int = 20, n_per_image: int = 1, seed: Optional[int] = None):
        """Multinomial rollout (eval-mode numerics; forward only -- the SCST backward pass is out of scope).
        ``n_per_image`` > 1 draws several samples per image while reading its features once (BASELINE config 5)."""
        feats, mask = self._features(visual_inputs)
        self._prepare(feats, mask)
        if seed is None:
            seed = self._seed + self._calls
            self._calls += 1
        tokens, logprobs = self.decoder.sample(capdec.SAMPLE_MULTINOMIAL, n_per_image, seed, max_len)
        return tokens.long(), logprobs


    def scst_rollouts(self, visual_inputs, max_len: int = 20, n_per_image: int = 1, seed: Optional[int] = None):
        """``greedy_res = sampler(...)`` and ``seq_gen, seqLogprobs = sampler_rl(...)`` of one SCST step
        (Engine.py:258-262) in ONE pass over the batch -> (greedy (B,T) long, seq (B*n,T) long, logprobs (B*n,T) float);
        row for row the same values as the two separate calls."""
        feats, mask = self._features(visual_inputs)
        self._prepare(feats, mask)
        if seed is None:
            seed = self._seed + self._calls
            self._calls += 1
        tokens, logprobs, greedy = self.decoder.scst_rollout(n_per_image, seed, max_len)
        return greedy.long(), tokens.long(), logprobs

    # not part of the reference's captioner surface: forward values of its teacher-forced ``forward`` for a rollout
    def score(self, visual_inputs, tokens, n_per_image: int = 1):
        """log p(word t | image, previous words) of given rollouts: tokens (B*n_per_image, T) as ``sampler_rl`` returns
        them -> (B*n_per_image, T) float.  What ``decoder.forward`` + log_softmax + gather give in the reference
        (BUTD_Model.py:97-151); lets a training loop re-score sequences sampled by the fused rollout."""
        feats, mask = self._features(visual_inputs)
        self._prepare(feats, mask)
        return self.decoder.score(tokens, n_per_image)


def ids_to_caption(ids: Sequence[int], ix2word) -> str:
    """Engine.py:288-297: words until '<end>', skipping '<sta>'."""
    words: List[str] = []
    for i in ids:
        w = ix2word[int(i)]
        if w == END_WORD:
            break
        if w != STA_WORD:
            words.append(w)
    return " ".join(words)


class CaptionEngine:
    """Mirror of ``Engine`` (Engine.py:16-41, 274-300) for evaluation-time caption generation."""

    model_type: Optional[str] = None

    def __init__(self, model_settings_json, dataset_name, caption_vocab, data_dir=None, use_bu="unused", device="cuda:0",
                 state_dict: Optional[Mapping[str, object]] = None, **decoder_kwargs):
        if isinstance(model_settings_json, (str, bytes)):
            with open(model_settings_json) as f:
                self.settings = json.load(f)
        else:
            self.settings = dict(model_settings_json)
        if self.model_type is not None and self.settings.get("model_type", self.model_type) != self.model_type:
            raise ValueError(f"{type(self).__name__} drives model_type {self.model_type}")
        self.device = device
        self.data_dir = data_dir
        self.dataset_name = dataset_name
        self.use_bu = use_bu
        self.caption_vocab = caption_vocab
        self.tag = "Model_" + self.settings["model_type"] + "_Dataset_" + dataset_name
        self._decoder_kwargs = decoder_kwargs
        self.model = None
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def _device_index(self) -> int:
        d = str(self.device)
        return int(d.split(":")[1]) if ":" in d else 0

    def load_state_dict(self, state_dict):
        """The checkpoint is the reference's ``model.state_dict()`` (Engine.py:81-88), loaded unchanged."""
        self.model = B200Captioner(self.settings["model_type"], self.settings, len(self.caption_vocab), state_dict,
                                   device=self._device_index(), **self._decoder_kwargs)

    def load_from_checkpoint(self, path):
        torch = _torch()
        self.load_state_dict(torch.load(path, map_location="cpu"))

    def modify_visual_inputs(self, img_tensors, supp_info_datas=(), device=None):
        torch = _torch()
        dev = device or self.device
        return {"img_tensors": img_tensors.to(dev) if torch.is_tensor(img_tensors) else img_tensors}

    def eval_captions_json_generation(self, dataloader, eval_beam_size=-1, tqdm_visible=False):
        self.model.eval()
        ix2word = self.caption_vocab.ix2word
        result = []
        fn = getattr(self.model, "feature_fn", None)
        host_ok = fn is None or getattr(fn, "accepts_host_inputs", False)  # a CUDA module as feature_fn needs device inputs
        if eval_beam_size != -1 and hasattr(self.model, "beam_search_stream") and host_ok:
            # pipelined: copy of batch i+1 and caption read-back of batch i overlap the decode
            ids_q = []

            def inputs():
                for image_ids, img_tensors, supp_info_datas in dataloader:
                    ids_q.append(image_ids)
                    yield self.modify_visual_inputs(img_tensors=img_tensors, supp_info_datas=supp_info_datas, device="cpu")

            for captions in self.model.beam_search_stream(inputs(), beam_size=eval_beam_size):
                image_ids = ids_q.pop(0)
                for i in range(captions.shape[0]):
                    result.append({"image_id": int(image_ids[i]), "caption": ids_to_caption(captions[i], ix2word)})
            return result
        for image_ids, img_tensors, supp_info_datas in dataloader:
            visual_inputs = self.modify_visual_inputs(img_tensors=img_tensors, supp_info_datas=supp_info_datas)
            if eval_beam_size != -1:
                generated = self.model.beam_search_sampler(visual_inputs=visual_inputs, beam_size=eval_beam_size)
            else:
                generated = self.model.sampler(visual_inputs=visual_inputs, max_len=20)
            captions = generated.cpu().numpy()
            for i in range(captions.shape[0]):
                result.append({"image_id": int(image_ids[i]), "caption": ids_to_caption(captions[i], ix2word)})
        return result


    def scst_forward_epoch(self, dataloader, reward, n_per_image: int = 1, max_len: int = 20):
        """Forward half of ``Engine.SCST_training_epoch`` (Engine.py:251-272) with everything on the device: for every
        ``(img_ids, img_tensors, img_gts, supp_info_datas)`` batch of the SCST dataloader yields
        ``(img_ids, seq_gen (B*n,T) long, seqLogprobs (B*n,T) float, greedy_res (B,T) long, rewards (B*n,T) float)`` --
        the inputs of ``RewardCriterion`` (Utils.py:290-317).  The host->device copy of batch i+1 overlaps the rollouts of
        batch i, both rollouts run in one pass, ``reward`` is a :class:`scst.CiderDReward`.  The backward pass and the
        optimiser step are the caller's (out of scope here)."""
        meta = []

        def inputs():
            for img_ids, img_tensors, img_gts, supp_info_datas in dataloader:
                meta.append((img_ids, img_gts))
                yield self.modify_visual_inputs(img_tensors=img_tensors, supp_info_datas=supp_info_datas, device="cpu")

        for vi in self.model.prefetch_to_device(inputs()):
            img_ids, img_gts = meta.pop(0)
            greedy, seq, logprobs = self.model.scst_rollouts(vi, max_len=max_len, n_per_image=n_per_image)
            ids = [int(i) if not isinstance(i, (str, bytes)) else i for i in img_ids]
            rewards = reward(seq, greedy, img_gts, ids, n_per_image=n_per_image)
            yield img_ids, seq, logprobs, greedy, rewards


class NIC_Eng(CaptionEngine):
    model_type = "NIC"


class BUTDSpatial_Eng(CaptionEngine):
    model_type = "BUTDSpatial"


class _BottomUpMixin:
    def modify_visual_inputs(self, img_tensors, supp_info_datas=None, device=None):
        """ModelEngines/BUTD_Engine.py:23-47 / AoA_Engine.py:23-47: pad per-image (n_i, 2048) bottom-up features to a
        batch tensor plus a {0,1} mask (None when every image has the same number of boxes).  ``device="cpu"`` keeps
        the batch in (pinned) host memory for the pipelined eval loop, which copies it on its own stream."""
        torch = _torch()
        dev = device or self.device
        bu_feats = [np.asarray(s["bu_feat"], dtype=np.float32) for s in supp_info_datas]
        bu_bboxes = [s.get("bu_bbox") for s in supp_info_datas]
        max_len = max(f.shape[0] for f in bu_feats)
        feats = np.zeros((len(bu_feats), max_len, bu_feats[0].shape[1]), np.float32)
        masks = np.zeros(feats.shape[:2], np.float32)
        for i, f in enumerate(bu_feats):
            feats[i, :f.shape[0]] = f
            masks[i, :f.shape[0]] = 1
        bu_masks = None if masks.sum() == masks.size else torch.from_numpy(masks).to(dev)
        t = torch.from_numpy(feats)
        if str(dev) == "cpu" and torch.cuda.is_available():
            t = t.pin_memory()
        return {"bu_feats": t.to(dev), "bu_bboxes": bu_bboxes, "bu_masks": bu_masks}


class BUTDDetection_Eng(_BottomUpMixin, CaptionEngine):
    model_type = "BUTDDetection"


class AoADetection_Eng(_BottomUpMixin, CaptionEngine):
    model_type = "AoADetection"


class AoASpatial_Eng(CaptionEngine):
    model_type = "AoASpatial"


ENGINES = {c.model_type: c for c in (NIC_Eng, BUTDSpatial_Eng, BUTDDetection_Eng, AoADetection_Eng, AoASpatial_Eng)}


def install(engine, state_dict=None, rebind_rl: bool = False, max_seq: int = 50, **decoder_kwargs):
    """Drop the extension into a LIVE reference ``Engine`` instance: the EVALUATION-time decode methods of
    ``engine.model`` -- ``sampler``, ``beam_search_sampler``, ``eval_test_image`` -- are rebound to the B200 decoder built
    from the model's own ``state_dict`` (checkpoint loaded unchanged).  See INTEGRATION.md.

    ``sampler_rl`` is the SCST TRAINING rollout (Engine.py:262): the reference calls it in train mode and back-propagates
    through the log-probs it returns, which the extension's forward-only rollout cannot provide.  It is therefore left
    alone unless ``rebind_rl=True``; even then the original method runs whenever autograd is recording and the model is
    in training mode, so ``Engine.SCST_training_epoch`` keeps working on an installed engine.

    ``max_seq`` is the beam-search step limit; the default is the reference's hard-coded 50 (BUTD_Model.py:260,
    NIC_Model.py:169, AoA_Model.py:435).  ``max_regions`` (decoder keyword) bounds the boxes per image: default 36 for
    BUTDDetection, 100 for AoADetection's adaptive features, the CNN grid for the Spatial models."""
    torch = _torch()
    sd = state_dict if state_dict is not None else engine.model.state_dict()
    model_type = engine.settings["model_type"]
    feature_fn = None
    ref = engine.model
    if model_type == "NIC":
        feature_fn = lambda vi: ref.encoder(vi["img_tensors"])  # noqa: E731  NIC_Model.py:277
    elif model_type == "BUTDSpatial":
        feature_fn = lambda vi: ref.encoder(vi["img_tensors"])  # noqa: E731  BUTD_Model.py:382
    elif model_type == "AoASpatial":
        # the CNN grid features; img_feats_porjection + aoa_refine (AoA_Model.py:598-601) run inside the library
        feature_fn = lambda vi: ref.encoder(vi["img_tensors"])  # noqa: E731
    # AoADetection: no feature_fn -- bu_feats / bu_masks go straight to the library (AoA_Model.py:748-751)
    dev = str(engine.device)
    fast = B200Captioner(model_type, engine.settings, len(engine.caption_vocab), sd, feature_fn=feature_fn,
                         feature_fn_returns="bottom_up" if model_type == "AoASpatial" else "decoder_input",
                         device=int(dev.split(":")[1]) if ":" in dev else 0, max_seq=max_seq, **decoder_kwargs)
    ref.sampler, ref.beam_search_sampler = fast.sampler, fast.beam_search_sampler
    ref.eval_test_image = fast.eval_test_image
    if rebind_rl:
        original_rl = ref.sampler_rl

        def sampler_rl(visual_inputs, max_len=20, **kw):
            if torch.is_grad_enabled() and getattr(ref, "training", False):
                return original_rl(visual_inputs, max_len=max_len)  # SCST training: needs dropout + an autograd graph
            return fast.sampler_rl(visual_inputs, max_len=max_len, **kw)

        ref.sampler_rl = sampler_rl
    return fast


# ---------------------------------------------------------------------------------------------------- host placement
def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (and therefore the pinned host buffers it allocates afterwards, first-touch) to the NUMA node
    the GPU hangs off, so that the per-batch host->device copies of the 8 ranks of a box do not cross sockets.
    Returns the node, or None when the topology cannot be read (the call is then a no-op)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001  (topology files missing, NVML absent, ...)
        return None


# ---------------------------------------------------------------------------------------------------- multi-GPU
def shard_bounds(n_images: int, rank: int, world: int):
    """Contiguous image shard of ``rank``: images are independent, K rows of an image stay together."""
    per = (n_images + world - 1) // world
    lo = min(rank * per, n_images)
    return lo, min(lo + per, n_images)


class CaptionGather:
    """One collective per EPOCH (north_star: "a single NCCL all-gather collects the captions"): every rank appends the
    [B_local, L] device caption block of each batch it decodes (``add``: a stream-ordered device copy, no communication,
    so the ranks never run in lock-step), and ``finish`` gathers all ranks' blocks with ONE ``all_gather_into_tensor``
    and returns them in original image order: ``[steps, world * B_local, L]`` where batch s is the concatenation of the
    ranks' shards of batch s in rank order (``shard_bounds``).  NCCL on GPUs, gloo on CPU for the tests."""

    def __init__(self, steps: int, b_local: int, length: int, device, dtype=None, group=None):
        torch = _torch()
        self.group = group
        self.buf = torch.zeros((steps, b_local, length), dtype=dtype or torch.int32, device=device)
        self.n = 0

    def reset(self):
        self.n = 0

    def add(self, tokens_local):
        if self.n >= self.buf.shape[0]:
            raise RuntimeError("CaptionGather: more batches than the epoch was sized for")
        rows = tokens_local.shape[0]
        self.buf[self.n, :rows].copy_(tokens_local, non_blocking=True)
        if rows < self.buf.shape[1]:
            self.buf[self.n, rows:].zero_()  # ragged last batch: <pad> rows
        self.n += 1

    def finish(self):
        torch = _torch()
        import torch.distributed as dist
        steps, b, L = self.n, self.buf.shape[1], self.buf.shape[2]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return self.buf[:steps]
        world = dist.get_world_size(self.group)
        mine = self.buf[:steps].contiguous()
        out = torch.empty((world * steps, b, L), dtype=mine.dtype, device=mine.device)  # rank-major concatenation
        dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.view(world, steps, b, L).permute(1, 0, 2, 3).reshape(steps, world * b, L)


def all_gather_captions(tokens_local, n_images: int, group=None):
    """The path's only collective: gather every rank's [B_local, L] int32 caption block in rank order
    (= original image order).  NCCL on GPUs (one all-gather over NVLink), gloo on CPU for the tests."""
    torch = _torch()
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (n_images + world - 1) // world
    L = tokens_local.shape[1]
    padded = torch.zeros((per, L), dtype=tokens_local.dtype, device=tokens_local.device)
    padded[:tokens_local.shape[0]] = tokens_local
    out = torch.empty((world * per, L), dtype=tokens_local.dtype, device=tokens_local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_images]
