"""CNN encoder feed in front of the decoders: the reference's ``EncoderCNN`` modules --

* NIC          ``Models/NIC_Model.py:8-37``     images -> ResNet-101 -> avgpool -> weight-normed Linear(2048 -> E)  => (B, E)
  (``self.bn`` is constructed but never applied in ``forward``)
* BUTDSpatial  ``Models/BUTD_Model.py:8-38``    images -> ResNet-101 -> AdaptiveAvgPool(s, s) -> permute       => (B, s*s, 2048)
* AoASpatial   ``Models/AoA_Model.py:164-195``  same grid output

The convolutions stay LIBRARY code (cuDNN through torchvision's ResNet-101, the very module the reference wraps); what
this file adds is how they are fed on a B200: channels-last fp16 tensors (tensor-core convolution kernels), the whole
forward captured once per batch shape in a CUDA graph (ResNet-101 is ~350 small launches), a static device input buffer
filled by an asynchronous copy from pinned host memory, and the fp32 hand-off the decoder library expects.  The checkpoint
is the reference's: ``encoder.feature_extractor.<i>.*`` (``nn.Sequential`` of conv1, bn1, relu, maxpool, layer1-4) and, for
NIC, ``encoder.img_embedding.{weight_g,weight_v,bias}``.  There is no CPU path: constructing a feed without CUDA raises.
"""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np


def _torch():
    import torch
    return torch


def build_feature_extractor():
    """The reference's ``feature_extractor`` Sequential (same child order => same state_dict keys), random init,
    without the ``pretrained=True`` download (no network here; BASELINE asks for random-init weights anyway)."""
    torch = _torch()
    import torchvision
    r = torchvision.models.resnet101(weights=None)
    return torch.nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool, r.layer1, r.layer2, r.layer3, r.layer4)


def make_encoder_state_dict(embed_dim: Optional[int] = None, seed: int = 0, residual_gain: float = 0.25) -> dict:
    """Synthetic ``encoder.*`` checkpoint entries (torch tensors): torchvision's default ResNet-101 init under
    ``torch.manual_seed(seed)``.  The last BatchNorm gain of every bottleneck is scaled by ``residual_gain`` -- with
    identity running statistics (mean 0 / var 1) an untrained residual stack otherwise doubles its variance per block
    and leaves the fp16 range, which a trained network never does.  NIC adds ``img_embedding`` (weight-normed Linear)."""
    torch = _torch()
    g = torch.Generator().manual_seed(seed)
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        fx = build_feature_extractor()
    sd = {}
    for k, v in fx.state_dict().items():
        v = v.clone()
        if k.endswith("bn3.weight"):
            v.mul_(residual_gain)
        sd["encoder.feature_extractor." + k] = v
    if embed_dim:
        bound = 1.0 / np.sqrt(2048)
        w = (torch.rand(embed_dim, 2048, generator=g) * 2 - 1) * bound
        sd["encoder.img_embedding.weight_v"] = w
        sd["encoder.img_embedding.weight_g"] = w.norm(dim=1, keepdim=True)
        sd["encoder.img_embedding.bias"] = (torch.rand(embed_dim, generator=g) * 2 - 1) * bound
    return sd


def _fold(conv, bn, dtype, device):
    """conv + eval-mode BatchNorm -> (weight, bias) of ONE convolution, folded in fp64: w' = w * g / sqrt(var + eps),
    b' = beta - mean * g / sqrt(var + eps).  Weight in channels-last memory format for the NHWC tensor-core kernels."""
    torch = _torch()
    w = conv.weight.detach().double()
    scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
    w = (w * scale[:, None, None, None]).to(device=device, dtype=dtype).contiguous(memory_format=torch.channels_last)
    return w, b.to(device=device, dtype=dtype), tuple(conv.stride), tuple(conv.padding)


class FusedResNetTrunk:
    """The reference's ``feature_extractor`` (ResNet-101 trunk, eval mode) as 105 fused cuDNN calls instead of ~350
    kernels: every BatchNorm is folded into its convolution, every conv + bias + ReLU is one ``cudnn_convolution_relu`` and
    every bottleneck's last conv + bias + residual add + ReLU one ``cudnn_convolution_add_relu`` (cuDNN's fused epilogues:
    the activation tensor is written once per convolution instead of three times)."""

    def __init__(self, fx, dtype, device):
        conv1, bn1, _, _, l1, l2, l3, l4 = list(fx.children())
        self.stem = _fold(conv1, bn1, dtype, device)
        self.blocks = []
        for layer in (l1, l2, l3, l4):
            for blk in layer.children():
                down = _fold(blk.downsample[0], blk.downsample[1], dtype, device) if blk.downsample is not None else None
                self.blocks.append((_fold(blk.conv1, blk.bn1, dtype, device), _fold(blk.conv2, blk.bn2, dtype, device),
                                    _fold(blk.conv3, blk.bn3, dtype, device), down))

    def __call__(self, x):
        torch = _torch()
        F = torch.nn.functional
        one = (1, 1)

        def conv_relu(t, p):
            return torch.cudnn_convolution_relu(t, p[0], p[1], p[2], p[3], one, 1)

        x = conv_relu(x, self.stem)
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        for c1, c2, c3, down in self.blocks:
            identity = x if down is None else F.conv2d(x, down[0], down[1], down[2], down[3])
            out = conv_relu(conv_relu(x, c1), c2)
            x = torch.cudnn_convolution_add_relu(out, c3[0], identity, 1.0, c3[1], c3[2], c3[3], one, 1)
        return x


class CnnFeed:
    """``feature_fn`` for :class:`engine.B200Captioner`: ``visual_inputs['img_tensors']`` (B,3,224,224) fp32, host
    (ideally pinned) or device -> the decoder's input as a CUDA fp32 tensor."""

    accepts_host_inputs = True  # the pipelined eval loop may hand it (pinned) host images

    def __init__(self, model_type: str, state_dict: Mapping[str, object], *, enc_img_size: int = 7, device: int = 0,
                 dtype: str = "fp16", use_graph: bool = True, fuse: bool = True):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("CnnFeed needs a CUDA device; there is no CPU path")
        if model_type not in ("NIC", "BUTDSpatial", "AoASpatial"):
            raise ValueError(f"{model_type} has no CNN encoder")
        self.model_type = model_type
        self.device = torch.device("cuda", device)
        self.dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[dtype]
        self.grid = enc_img_size
        self.use_graph = use_graph
        fx = build_feature_extractor()
        prefix = "encoder.feature_extractor."
        sub = {k[len(prefix):]: (torch.as_tensor(v) if not torch.is_tensor(v) else v) for k, v in state_dict.items()
               if k.startswith(prefix)}
        fx.load_state_dict(sub, strict=True)
        fx = fx.eval()
        self.trunk = None
        if fuse:  # folded-BN, fused-epilogue cuDNN calls; checked once against the module forward, else the modules are kept
            try:
                trunk = FusedResNetTrunk(fx, self.dtype, self.device)
                probe = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(self.device, self.dtype)
                probe = probe.contiguous(memory_format=torch.channels_last)
                with torch.no_grad():
                    got = trunk(probe).float()
                    ref = fx.to(self.device)(probe.float())
                if torch.isfinite(got).all() and ((got - ref).norm() / ref.norm()).item() < 2e-2:
                    self.trunk = trunk
            except Exception:  # noqa: BLE001  (a cuDNN build without the fused conv-bias-relu engines)
                self.trunk = None
        self.fx = fx.to(self.device, self.dtype).to(memory_format=torch.channels_last)
        for p in self.fx.parameters():
            p.requires_grad_(False)
        self.W = self.b = None
        if model_type == "NIC":  # weight-normed Linear folded once: w = g * v / ||v|| (NIC_Model.py:24)
            v = torch.as_tensor(state_dict["encoder.img_embedding.weight_v"]).double()
            g = torch.as_tensor(state_dict["encoder.img_embedding.weight_g"]).double()
            self.W = (v * (g / v.norm(dim=1, keepdim=True))).float().to(self.device)
            self.b = torch.as_tensor(state_dict["encoder.img_embedding.bias"]).float().to(self.device)
        self._graphs = {}
        self._turn = {}
        self.stream = torch.cuda.Stream(self.device)       # capture stream
        self.copy_stream = torch.cuda.Stream(self.device)  # host -> device copies of the next batch

    # ------------------------------------------------------------------ forward
    def _forward(self, x):
        torch = _torch()
        f = self.trunk(x) if self.trunk is not None else self.fx(x)  # (B, 2048, h, w) channels-last
        if self.model_type == "NIC":
            pooled = f.float().mean(dim=(2, 3))  # resnet.avgpool + view (NIC_Model.py:34-35), accumulated in fp32
            return torch.addmm(self.b, pooled, self.W.t())  # img_embedding (:36)
        if f.shape[2] != self.grid or f.shape[3] != self.grid:
            f = torch.nn.functional.adaptive_avg_pool2d(f.float(), (self.grid, self.grid))
        # (B, 2048, s, s) -> (B, s*s, 2048): channels-last storage already is that layout, so this is a view + cast
        return f.permute(0, 2, 3, 1).reshape(f.shape[0], -1, f.shape[1]).float().contiguous()

    def _capture(self, key, images):
        """One replayable instance of the forward for this input shape: static input / output buffers + CUDA graph."""
        torch = _torch()
        cur = torch.cuda.current_stream(self.device)
        static_in = torch.empty(key, dtype=torch.float32, device=self.device)
        static_in.copy_(images, non_blocking=True)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for _ in range(2):  # warm-up outside capture (cuDNN algorithm selection, workspace allocation)
                self._forward(static_in.to(self.dtype).contiguous(memory_format=torch.channels_last))
        cur.wait_stream(self.stream)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=self.stream):
            static_out = self._forward(static_in.to(self.dtype).contiguous(memory_format=torch.channels_last))
        return dict(graph=graph, static_in=static_in, static_out=static_out, free=None)

    def __call__(self, visual_inputs):
        torch = _torch()
        images = visual_inputs["img_tensors"] if isinstance(visual_inputs, dict) else visual_inputs
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(images)
        key = tuple(images.shape)
        cur = torch.cuda.current_stream(self.device)
        with torch.no_grad():
            if not self.use_graph:
                x = images.to(self.device, non_blocking=True).to(self.dtype).contiguous(memory_format=torch.channels_last)
                return self._forward(x)
            # two instances per shape, used in turn: the host -> device copy of batch i+1 (copy stream) overlaps the
            # convolutions of batch i (caller's stream)
            slots = self._graphs.setdefault(key, [])
            turn = self._turn.get(key, 0)
            self._turn[key] = turn + 1
            if len(slots) < 2:
                slots.append(self._capture(key, images))
            slot = slots[turn % len(slots)] if len(slots) == 2 else slots[-1]
            if images.is_cuda:
                slot["static_in"].copy_(images, non_blocking=True)
            else:
                with torch.cuda.stream(self.copy_stream):
                    if slot["free"] is not None:
                        self.copy_stream.wait_event(slot["free"])  # the replay that last read this buffer has finished
                    slot["static_in"].copy_(images, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(self.copy_stream)
                cur.wait_event(ready)
            slot["graph"].replay()
            # a pipelined caller stages batch i+1 before batch i is decoded: hand out a copy, not the graph's own buffer
            out = slot["static_out"].clone()
            slot["free"] = torch.cuda.Event()
            slot["free"].record(cur)
            return out


def attach(captioner, state_dict, **kw):
    """Give a :class:`engine.B200Captioner` of a CNN-fed model type its encoder feed (sets ``feature_fn``)."""
    s = int(captioner.settings.get("enc_img_size", 7) or 7)
    feed = CnnFeed(captioner.model_type, state_dict, enc_img_size=s, device=captioner.device.index or 0, **kw)
    captioner.feature_fn = feed
    if captioner.model_type == "AoASpatial" and captioner.decoder.has_refiner:
        captioner.native_refiner = True  # grid features -> img_feats_porjection + aoa_refine inside the library
    return feed
