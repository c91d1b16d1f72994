"""Packed feature shards and their loader.

The reference keeps one zlib-compressed ``.npz`` per image -- ``np.savez_compressed(path, feat=(36,2048) fp32, bbox=...)``
written by PreProcess/Generate_coco14_bottom_up_features_data.py:56-58 -- and its ``Dataset.__getitem__`` opens and inflates
them one by one (Datasets.py:138-145); ``*_Eng.modify_visual_inputs`` then pads a batch in numpy and copies it to the
device (BUTD_Engine.py:36-45).  At >10^5 captions/s that loader is the wall, and the fp32 host->device copy (453 MB per
1536 images) is what bounds this repo's end-to-end number.  A shard here is ONE flat file:

    [ 4096-byte JSON header | feats fp16[N,R,D] | image_ids int64[N] | lengths int32[N] | bboxes fp32[N,R,4] (optional) ]

memory-mapped, features stored as the fp16 values the decoder's tensor-core operands use anyway, rows zero-padded to R
regions with the true count in ``lengths`` (adaptive bottom-up features -> ``bu_masks``).  ``batches()`` fills a small ring
of pinned host buffers from the map on a background thread and yields ``visual_inputs`` dicts whose ``bu_feats`` is a
pinned fp16 tensor: ``B200Captioner.beam_search_stream`` copies it asynchronously and the library consumes it without a
conversion pass (``capdec_prepare_f16``).
"""
from __future__ import annotations

import json
import os
import queue
import threading
from typing import Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np

MAGIC = "b200-caption-feature-shard"
HEADER_BYTES = 4096


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class FeatureShardWriter:
    """``with FeatureShardWriter(path, regions=36, dim=2048) as w: w.append(image_id, feat[, bbox])``.  Feature rows are
    streamed to the file as they arrive (a COCO-sized shard is ~18 GB); the small per-image index follows the feature block
    when the writer is closed and the header records where everything is."""

    def __init__(self, path: str, regions: int, dim: int = 2048, with_bboxes: bool = False):
        self.path, self.R, self.D, self.with_bboxes = path, regions, dim, with_bboxes
        self._ids, self._lens, self._boxes = [], [], []
        self._f = open(path, "wb")
        self._f.write(b" " * HEADER_BYTES)
        self._closed = False

    def append(self, image_id: int, feat: np.ndarray, bbox: Optional[np.ndarray] = None):
        feat = np.asarray(feat)
        n = feat.shape[0]
        if feat.ndim != 2 or feat.shape[1] != self.D or n < 1 or n > self.R:
            raise ValueError(f"feature of image {image_id} has shape {feat.shape}; expected (1..{self.R}, {self.D})")
        row = np.zeros((self.R, self.D), np.float16)
        with np.errstate(over="ignore"):
            row[:n] = feat  # fp32 -> fp16, round to nearest even: the rounding the decoder applies to its operands
        if not np.isfinite(row).all():
            raise ValueError(f"feature of image {image_id} leaves the fp16 range")
        self._f.write(row.tobytes())
        self._ids.append(int(image_id))
        self._lens.append(n)
        if self.with_bboxes:
            b = np.zeros((self.R, 4), np.float32)
            if bbox is not None:
                b[:n] = bbox
            self._boxes.append(b)

    def close(self):
        if self._closed:
            return
        self._closed = True
        N = len(self._ids)
        off_feats = HEADER_BYTES
        off_ids = _align(off_feats + 2 * N * self.R * self.D)
        off_lens = _align(off_ids + 8 * N)
        off_box = _align(off_lens + 4 * N)
        header = dict(magic=MAGIC, version=1, n=N, regions=self.R, dim=self.D, dtype="float16", with_bboxes=self.with_bboxes,
                      off_ids=off_ids, off_lens=off_lens, off_bboxes=off_box, off_feats=off_feats)
        blob = json.dumps(header).encode()
        assert len(blob) < HEADER_BYTES
        f = self._f
        f.seek(off_ids)
        f.write(np.asarray(self._ids, np.int64).tobytes())
        f.seek(off_lens)
        f.write(np.asarray(self._lens, np.int32).tobytes())
        f.seek(off_box)
        f.write(np.stack(self._boxes).tobytes() if (self.with_bboxes and N) else b"\0")
        f.seek(0)
        f.write(blob.ljust(HEADER_BYTES, b" "))
        f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        else:
            self._f.close()


def convert_npz_directory(npz_paths: Iterable[Tuple[int, str]], shard_path: str, regions: int, dim: int = 2048):
    """One-time conversion of the reference's per-image files (``np.load(p)['feat']``, ``['bbox']``) into a shard."""
    with FeatureShardWriter(shard_path, regions, dim, with_bboxes=True) as w:
        for image_id, p in npz_paths:
            z = np.load(p)
            w.append(image_id, z["feat"], z["bbox"] if "bbox" in z.files else None)


class FeatureShard:
    """Read side: ``len(shard)``, ``shard.image_ids``, ``shard.features(i)`` (fp16 view), ``shard.batches(...)``."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            header = json.loads(f.read(HEADER_BYTES).decode().strip())
        if header.get("magic") != MAGIC:
            raise ValueError(f"{path} is not a feature shard")
        self.header = header
        self.N, self.R, self.D = header["n"], header["regions"], header["dim"]
        self.image_ids = np.memmap(path, np.int64, "r", header["off_ids"], (self.N,))
        self.lengths = np.memmap(path, np.int32, "r", header["off_lens"], (self.N,))
        self.bboxes = (np.memmap(path, np.float32, "r", header["off_bboxes"], (self.N, self.R, 4))
                       if header["with_bboxes"] else None)
        self.feats = np.memmap(path, np.float16, "r", header["off_feats"], (self.N, self.R, self.D))
        expect = max(header["off_feats"] + 2 * self.N * self.R * self.D, header["off_lens"] + 4 * self.N,
                     header["off_ids"] + 8 * self.N,
                     header["off_bboxes"] + 16 * self.N * self.R if header["with_bboxes"] else 0)
        if os.path.getsize(path) < expect:
            raise ValueError(f"{path} is truncated: {os.path.getsize(path)} < {expect} bytes")

    def __len__(self):
        return self.N

    def features(self, i: int) -> np.ndarray:
        return self.feats[i, :self.lengths[i]]

    def batches(self, batch_size: int, *, start: int = 0, stop: Optional[int] = None, pinned: bool = True,
                ring: int = 3) -> Iterator[Tuple[np.ndarray, dict]]:
        """Yields ``(image_ids, visual_inputs)`` in file order; ``visual_inputs = {'bu_feats': fp16 (B,R,D) host tensor
        (pinned when a CUDA device is present), 'bu_bboxes': ..., 'bu_masks': float (B,R) tensor or None}`` -- the dict
        ``BUTDDetection_Eng.modify_visual_inputs`` builds (BUTD_Engine.py:23-47).  The buffers belong to a ring of ``ring``
        slots refilled by a background thread.  A consumer that copies a batch asynchronously hands the copy's CUDA event
        to ``visual_inputs['_on_copied'](event)`` (``beam_search_stream`` does): the slot is not refilled before that event
        has completed.  Other consumers must be done with a batch before they ask for the one after the next.  ``ring`` >= 2;
        3 or more slots let the refill of the next batch overlap the consumer's work on the current two."""
        import torch
        stop = self.N if stop is None else min(stop, self.N)
        pin = pinned and torch.cuda.is_available()
        if ring < 2:
            raise ValueError("ring must be >= 2 (one slot with the consumer, one being filled)")
        slots = [torch.empty((batch_size, self.R, self.D), dtype=torch.float16, pin_memory=pin) for _ in range(ring)]
        q: "queue.Queue" = queue.Queue()
        # Explicit hand-back of slots: the fill thread takes a slot from ``free`` (blocking), the consumer side returns
        # the slot of batch k when batch k+2 is requested, together with the CUDA event of its asynchronous copy if one
        # was registered -- so a slot is never rewritten under a pending or not-yet-issued copy, whatever ``ring`` is.
        free: "queue.Queue" = queue.Queue()
        for i in range(ring):
            free.put((i, None))
        copied = [None] * ring  # CUDA event of the last asynchronous copy out of each slot

        def on_copied(slot):
            def register(event):
                copied[slot] = event
            return register

        def fill():
            try:
                for lo in range(start, stop, batch_size):
                    hi = min(lo + batch_size, stop)
                    item = free.get()
                    if item is None:  # the consumer went away
                        return
                    slot, event = item
                    if event is not None:
                        event.synchronize()
                    buf = slots[slot][:hi - lo]
                    np.copyto(buf.numpy(), self.feats[lo:hi])  # page-cache -> pinned memory, no decompression
                    lens = np.asarray(self.lengths[lo:hi])
                    mask = None
                    if (lens != self.R).any():
                        mask = torch.from_numpy((np.arange(self.R)[None, :] < lens[:, None]).astype(np.float32))
                    boxes = None if self.bboxes is None else [np.asarray(self.bboxes[i, :lens[i - lo]]) for i in range(lo, hi)]
                    q.put((slot, (np.asarray(self.image_ids[lo:hi]),
                                  {"bu_feats": buf, "bu_bboxes": boxes, "bu_masks": mask, "_on_copied": on_copied(slot)})))
                q.put(None)
            except BaseException as e:  # noqa: BLE001  (surface loader errors in the consumer)
                q.put(e)

        t = threading.Thread(target=fill, daemon=True)
        t.start()
        held = []  # slots of the batches the consumer may still be using (the last two yielded)
        try:
            while True:
                if len(held) == 2:  # asking for batch k+2 releases batch k
                    s0 = held.pop(0)
                    free.put((s0, copied[s0]))
                    copied[s0] = None
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                held.append(item[0])
                yield item[1]
        finally:
            free.put(None)
        t.join()


def ids_to_captions(captions: np.ndarray, ix2word, end_id: int = 2, sta_id: int = 1) -> list:
    """Vectorised form of the id -> words loop of ``Engine.eval_captions_json_generation`` (Engine.py:288-297) for a whole
    batch: words until ``<end>``, ``<sta>`` skipped.  ``captions`` (B, L) ints, ``ix2word`` a sequence or dict."""
    captions = np.asarray(captions)
    if isinstance(ix2word, dict):
        table = np.empty(max(ix2word) + 1, dtype=object)
        for i, w in ix2word.items():
            table[i] = w
    else:
        table = np.asarray(ix2word, dtype=object)
    is_end = captions == end_id
    first_end = np.where(is_end.any(1), is_end.argmax(1), captions.shape[1])
    words = table[captions]
    keep = (np.arange(captions.shape[1])[None, :] < first_end[:, None]) & (captions != sta_id)
    return [" ".join(words[b, keep[b]]) for b in range(captions.shape[0])]
