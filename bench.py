#!/usr/bin/env python
"""Benchmark of the caption decode hot path: captions/sec at beam=3, max_seq=20 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on the box's host cores

A "step" is one pass of the hot path over one batch of synthetic images: step-invariant preparation
(projection GEMMs) + max_seq beam-search steps + result selection (+ the caption all-gather when N > 1).
``value`` is timed with the inputs resident in HBM; ``e2e`` is the same step through the reference-facing
captioner API (``beam_search_sampler(visual_inputs)``) with HOST inputs: pinned host -> device copy of the step's
features and device -> host read of the captions inside the timed region.

The oracle (oracle/capdec_oracle.py, a numpy port of the reference's decode loop) is executed here ONLY for the
``cpu_baseline`` object / the ``--impl reference`` arm, and as a checker of a handful of GPU captions.
"""
import argparse
import json
import os
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1; the reference arm is a CPU measurement "with all the host threads it can use", so the
# BLAS thread count is restored BEFORE numpy is imported.  NCCL's version banner is kept off stdout (one JSON line).
if "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"  # both VERSION and WARN make NCCL printf its version banner to stdout

# The contract is ONE JSON line on stdout.  Native libraries (NCCL, cuDNN, ...) printf to file descriptor 1 whatever Python's
# sys.stdout is, so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from simpleimagecaptionzoo_b200 import synth  # noqa: E402

METRIC = "captions/sec (beam=3, max_seq=20)"
# /root/reference does not travel to the GPU box, so the CPU arm times the numpy port.  Where both run (the build container,
# 8 cores, tests/tools/cpu_reference_rate.py -> profiles/r02_cpu_reference_vs_port.json) the reference's OWN PyTorch code
# decodes 3.85 captions/s and the port 2.54: divide a GPU/port ratio by ~1.5 to compare with the real reference.
PORT_VS_REFERENCE = ("build-container measurement on the same 16 images and 8 cores: reference's own code 3.85 captions/s, this "
                     "port 2.54 (0.66x) -- profiles/r02_cpu_reference_vs_port.json")
UNIT = "captions/s"

WORKLOADS = {
    # name: arch, model_type, regions, beam, images per GPU, description
    # 1536 images x beam 3 = 4608 rows = 18 row blocks of 256 (one CTA pair each): the gate GEMMs (16 column tiles) are
    # 288 pair tiles = 3.9 waves of the 74 SM pairs
    "butd_det": dict(arch="BUTD", model_type="BUTDDetection", R=36, beam=3, batch=1536,
                     desc="BUTDDetection beam=3 eval, synthetic 36x2048 region feats, vocab 9487, random init "
                          "(BASELINE configs[0] model at a GPU-sized batch)"),
    "butd_spatial": dict(arch="BUTD", model_type="BUTDSpatial", R=196, beam=5, batch=921,
                         desc="BUTDSpatial beam=5 over a 14x14x2048 feature grid (configs[2], decoder only)"),
    "nic": dict(arch="NIC", model_type="NIC", R=0, beam=3, batch=1536,
                desc="NIC LSTM decoder beam=3 on synthetic image embeddings (configs[1] without the ResNet-101 encoder)"),
    # configs[4]: SCST rollout -- 5 multinomial samples + 1 greedy baseline per image; a "caption" here is one image's rollout set
    "scst": dict(arch="BUTD", model_type="BUTDDetection", R=36, beam=5, batch=921, scst=True,
                 desc="BUTDDetection SCST sampling rollout (5 multinomial samples/image + greedy baseline), forward values only "
                      "(configs[4] per-GPU shard)"),
    "aoa": dict(arch="AOA", model_type="AoADetection", R=36, beam=3, batch=1536,
                desc="AoADetection 8-head AoA decoder beam=3 (configs[3] per-GPU shard, refined feats synthetic)"),
    # configs[1] as BASELINE.json states it: 224x224 images -> ResNet-101 (cuDNN via torchvision, channels-last fp16, CUDA
    # graph: library code) -> weight-normed Linear -> this repo's NIC decoder; batch 256
    "nic_images": dict(arch="NIC", model_type="NIC", R=0, beam=3, batch=256, images=True,
                       desc="NIC (ResNet-101 encoder + LSTM) beam=3 on synthetic 224x224 images, batch 256 (configs[1]); encoder = "
                            "torchvision/cuDNN fp16 channels-last under a CUDA graph, decoder = libcapdec"),
    # configs[3] from the bottom-up features: img_feats_porjection + 6-layer aoa_refine + decoder, all in the library
    "aoa_bu": dict(arch="AOA", model_type="AoADetection", R=36, beam=3, batch=1536, refiner=True,
                   desc="AoADetection beam=3 from synthetic 36x2048 bottom-up feats: projection + 6-layer AoA refiner + 8-head "
                        "AoA decoder (configs[3] per-GPU shard)"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def make_inputs(w, batch, seed):
    dims = synth.DIMS[w["arch"]]
    if w.get("images"):
        return np.random.Generator(np.random.PCG64(6_000_003 + seed)).standard_normal((batch, 3, 224, 224), dtype=np.float32)
    if w["arch"] == "BUTD" or w.get("refiner"):
        return synth.make_region_feats(batch, w["R"], dims.get("enc_dim", 2048), seed)
    if w["arch"] == "NIC":
        return synth.make_image_embed(batch, dims["embed_dim"], seed)
    return synth.make_refined_feats(batch, w["R"], dims["hidden_dim"], seed)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML; nvidia-smi fallback)."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nvml = None

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
             0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._nvml.nvmlDeviceGetClockInfo(self._h, self._nvml.NVML_CLOCK_SM))
                    mask = self._nvml.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(
                        self._nvml, "nvmlDeviceGetCurrentClocksEventReasons") else self._nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.NAMES.items():
                        if mask & bit and name != "gpu_idle":
                            self.reasons.add(name)
                else:
                    import subprocess
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    self.samples.append(int(f[0]))
                    self.max_mhz = int(f[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------- CPU arms
def oracle_decoder(w, sd):
    from oracle import capdec_oracle as orc
    return orc, orc.make_decoder(w["arch"], sd)


_ORACLE_CACHE = {}


def time_reference_form(w, sd, feats, beam, max_seq):
    """The reference's own driver shape -- one image per beam-search call (Utils.py:72-73) -- on all host cores.
    Weights are folded once, outside the timed region (a checkpoint load, not decode work).
    Returns (seconds, BeamResult)."""
    if _ORACLE_CACHE.get("sd") is not sd:  # one workload at a time: keyed on the state_dict object itself
        _ORACLE_CACHE.clear()
        _ORACLE_CACHE["sd"] = sd
        _ORACLE_CACHE["dec"] = oracle_decoder(w, sd)
    orc, dec = _ORACLE_CACHE["dec"]
    t0 = time.perf_counter()
    if w.get("images"):  # the reference's own encoder on the CPU: torchvision ResNet-101 fp32 + img_embedding (NIC_Model.py:27-37)
        import torch
        from simpleimagecaptionzoo_b200 import cnn_feed
        if "cnn" not in _ORACLE_CACHE:
            fx = cnn_feed.build_feature_extractor()
            fx.load_state_dict({k[len("encoder.feature_extractor."):]: v for k, v in sd.items() if k.startswith("encoder.feature_extractor.")})
            v, g = sd["encoder.img_embedding.weight_v"].double(), sd["encoder.img_embedding.weight_g"].double()
            _ORACLE_CACHE["cnn"] = (fx.eval(), (v * (g / v.norm(dim=1, keepdim=True))).float(), sd["encoder.img_embedding.bias"].float())
        fx, W, b = _ORACLE_CACHE["cnn"]
        with torch.no_grad():
            feats = torch.addmm(b, fx(torch.from_numpy(feats)).mean(dim=(2, 3)), W.t()).numpy()
    if w.get("refiner"):  # AoA_Model.py:748-751: projection + refiner run on every sampler call
        feats = orc.aoa_project_refine(sd, feats, None)
    dec.prepare(feats)
    if w.get("scst"):  # the SCST step forward as the reference runs it (Engine.py:258-263): greedy, 1..n samples, CIDEr-D
        n_img = feats.shape[0]
        if "cider" not in _ORACLE_CACHE:
            from simpleimagecaptionzoo_b200 import scst as scst_mod
            ix2word, refs = synth.make_caption_corpus(64, synth.DIMS[w["arch"]]["vocab_size"], seed=0)
            _ORACLE_CACHE["cider"] = (ix2word, refs) + scst_mod.document_frequency_from_corpus(refs)
        ix2word, refs, df, ref_len = _ORACLE_CACHE["cider"]
        greedy, _, _ = orc.greedy_sample(dec, max_seq)
        seq, _, _ = orc.multinomial_sample(dec, max_seq, beam, 0)
        ids = [i % len(refs) for i in range(n_img) for _ in range(beam)]
        orc.self_critical_reward(seq.reshape(n_img * beam, max_seq), np.repeat(greedy, beam, axis=0), dict(enumerate(refs)), ids,
                                 ix2word, df, ref_len)
        return time.perf_counter() - t0, None
    res = orc.beam_search_reference_form(dec, beam, max_seq)
    return time.perf_counter() - t0, res


def make_weights(w):
    dims = synth.DIMS[w["arch"]]
    sd = synth.make_state_dict(w["arch"], seed=0, **dims)
    if w.get("refiner"):
        sd.update(synth.make_refiner_state_dict(hidden_dim=dims["hidden_dim"], enc_dim=2048, seed=0))
    if w.get("images"):
        from simpleimagecaptionzoo_b200 import cnn_feed
        sd.update(cnn_feed.make_encoder_state_dict(embed_dim=dims["embed_dim"], seed=0))
    return sd


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dims = synth.DIMS[w["arch"]]
    sd = make_weights(w)
    n_img = args.ref_images
    for i in range(args.warmup):
        time_reference_form(w, sd, make_inputs(w, n_img, 100 + i), w["beam"], args.max_seq)
    total = 0.0
    for i in range(args.steps):
        dt, _ = time_reference_form(w, sd, make_inputs(w, n_img, 200 + i), w["beam"], args.max_seq)
        total += dt
    value = args.steps * n_img / total
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "images_per_step": n_img, "beam": w["beam"], "max_seq": args.max_seq,
                   "note": "reference algorithm (numpy port, weight-norm folded once, enc_att hoisted), one image per "
                           "beam-search call as the reference forces; CPU only, rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_img} images per step x {args.steps} steps, numpy/BLAS on {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------- GPU arm
# BASELINE.json configs at their STATED batches (per GPU where the config says "sharded 8 GPUs" / "8xB200") and the
# fp32-grade math mode, run as short passes after the headline and reported under "workloads" in the same JSON line.
EXTRAS = [
    # key, workload, images per GPU, math
    ("cfg1_butd_det_batch16", "butd_det", 16, "f16"),          # configs[0] as written (the reference's CPU-runnable case)
    ("cfg2_nic_images_batch256", "nic_images", 256, "f16"),    # configs[1]
    ("cfg3_butd_spatial_batch1024", "butd_spatial", 1024, "f16"),  # configs[2]
    ("cfg4_aoa_bu_256_per_gpu", "aoa_bu", 256, "f16"),         # configs[3]: batch 2048 over 8 GPUs
    ("cfg5_scst_512_per_gpu", "scst", 512, "f16"),             # configs[4]: batch 4096 over 8 GPUs
    ("headline_f16x3", "butd_det", None, "f16x3"),             # the headline workload in the fp32-grade math mode
]


def kernel_sources_sha():
    """sha256 over the CUDA sources: stamps profiles/ncu_traffic.json (tools/ncu_traffic.py) so that a DRAM-traffic figure
    captured for other kernels than the ones being timed is not reported."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "simpleimagecaptionzoo_b200", "csrc")
    for n in sorted(os.listdir(d)):
        if n.endswith((".cu", ".cuh")):
            h.update(n.encode())
            h.update(open(os.path.join(d, n), "rb").read())
    return h.hexdigest()[:16]


class GpuContext:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from simpleimagecaptionzoo_b200 import engine
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the caption decoder has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa_node = engine.bind_to_gpu_numa_node(self.local) if self.world > 1 else None  # NUMA-local pinned buffers per rank
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == args.gpus or self.world == 1, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure_h2d_ceiling(ctx, host, iters=8):
    """Pinned host -> device copy rate with ALL ranks copying at once and nothing else running: the ceiling of any
    end-to-end number whose inputs start on the host.  -> (GB/s of this job, GB/s per GPU)."""
    torch = ctx.torch
    dst = torch.empty_like(host, device=ctx.dev)
    st = torch.cuda.Stream(ctx.dev)
    with torch.cuda.stream(st):
        for _ in range(2):
            dst.copy_(host, non_blocking=True)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(iters):
            dst.copy_(host, non_blocking=True)
        e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    nbytes = host.numel() * host.element_size()
    per_gpu = nbytes * iters / (ms * 1e-3) / 1e9
    del dst
    return per_gpu * ctx.world, per_gpu


def measure(ctx, args, wname, batch, math, steps, warmup, cpu_images, headline):
    """One workload on this rank's GPU (all ranks run it together): device-resident throughput, end-to-end throughput
    through the captioner API with host buffers, per-kernel event timing, a CPU-oracle parity sample (rank 0, N=1)."""
    import gc

    torch, dist = ctx.torch, ctx.dist
    from simpleimagecaptionzoo_b200 import capdec, engine

    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    w = dict(WORKLOADS[wname])
    dims = synth.DIMS[w["arch"]]
    B, K, T, R = batch, w["beam"], args.max_seq, w["R"]
    sd = make_weights(w)
    settings = dict(model_type=w["model_type"], embed_dim=dims["embed_dim"], hidden_dim=dims["hidden_dim"],
                    atten_dim=dims.get("atten_dim", 0))
    feature_fn = None
    raw_bu = w["model_type"] == "BUTDDetection" or bool(w.get("refiner"))
    if not raw_bu:  # CNN encoder (or a synthetic stand-in for the refined features): the decoder's input is the input
        feature_fn = lambda vi: vi["feats"]  # noqa: E731
    cap = engine.B200Captioner(w["model_type"], settings, dims["vocab_size"], sd, feature_fn=feature_fn, max_batch=B,
                               max_regions=max(R, 1), max_rows=K + 1 if w.get("scst") else K, max_seq=T, math=math, device=local,
                               enc_dim=dims.get("enc_dim", 2048), num_heads=dims.get("num_heads", 8))
    dec = cap.decoder
    key = "bu_feats" if raw_bu else "feats"
    if w.get("images"):
        from simpleimagecaptionzoo_b200 import cnn_feed
        cnn_feed.attach(cap, sd)
        key = "img_tensors"

    # rank-local shard of the global batch (weak scaling: B images per GPU), distinct per rank
    host_feats = torch.from_numpy(make_inputs(w, B, 1000 + rank)).pin_memory()
    dev_feats = host_feats.to(dev)
    n_total = B * world
    scst = bool(w.get("scst"))
    L_out = K * T if scst else T + 1
    # multi-GPU: every rank appends its caption block per batch; ONE all-gather per timed run collects them
    gather = engine.CaptionGather(max(steps, warmup, 3), B, L_out, dev)

    def step_device():
        if w.get("images"):
            cap._prepare(cap.feature_fn({key: dev_feats}), None)
        else:
            cap._prepare(dev_feats, None)
        tok, _, _ = dec.beam_search(K, T)
        gather.add(tok)
        return tok

    reward = None
    if scst:  # the SCST step forward: both rollouts in one pass (Engine.py:258-262) + the CIDEr-D reward on the device (Utils.py:319-367)
        from simpleimagecaptionzoo_b200 import scst as scst_mod
        ix2word, refs = synth.make_caption_corpus(B, dims["vocab_size"], seed=rank)
        df, ref_len = scst_mod.document_frequency_from_corpus(refs)
        reward = scst_mod.CiderDReward({wd: i for i, wd in enumerate(ix2word)}, df, ref_len, device=local)
        gts, img_ids = dict(enumerate(refs)), list(range(B))
        calls = [0]
        last_greedy = [None]

        def step_device():  # noqa: F811
            dec.prepare(dev_feats)
            tok, _, greedy = dec.scst_rollout(K, calls[0], T)
            calls[0] += 1
            last_greedy[0] = greedy
            step_device.rewards = reward(tok, greedy, gts, img_ids, n_per_image=K)
            tok = tok.view(B, K * T)
            gather.add(tok)
            return tok

    def run_e2e(n_steps, host=host_feats):
        """n_steps batches through the pipelined captioner API: every step's features start in pinned HOST memory and
        every step's captions end in host memory (H2D of step i+1 overlaps the decode of step i); with several GPUs the
        device caption blocks are gathered once at the end of the run and read back too."""
        last = None
        gather.reset()
        if scst:  # prefetched H2D, greedy + multinomial rollouts, CIDEr-D reward, read-back
            for vi in cap.prefetch_to_device(({key: host} for _ in range(n_steps))):
                greedy, seq, _ = cap.scst_rollouts(vi, max_len=T, n_per_image=K)
                rew = reward(seq, greedy, gts, img_ids, n_per_image=K)
                gather.add(seq.view(B, K * T).to(torch.int32))
                last = seq.view(B, K * T).to(torch.int32).cpu().numpy()
                rew[:, 0].cpu()
        else:
            for tok in cap.beam_search_stream(({key: host} for _ in range(n_steps)), beam_size=K, max_seq=T,
                                              on_device_tokens=gather.add):
                last = tok
        if world > 1:  # ONE all-gather for the run, then every caption of every rank's batches to the host
            last = gather.finish().cpu().numpy()[-1]
        return last

    def timed_e2e(host):
        run_e2e(3, host)
        ctx.barrier()
        t0 = time.perf_counter()
        tok_host = run_e2e(steps, host)
        torch.cuda.synchronize()
        sec = ctx.max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            dist.barrier()
        return sec, tok_host

    # ---- device-resident timing (value)
    for _ in range(warmup):
        step_device()
    gather.reset()
    ctx.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = dec.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tokens = step_device()
    gathered = gather.finish()  # the path's only collective (no-op on one GPU)
    e1.record()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = dec.launch_count - launches0
    ms_per_step = ms_total / steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- end to end through the captioner API with host buffers.  Host format: the packed fp16 feature-shard rows of
    # feature_store.py where the decoder takes them (BUTD / AoA bottom-up features in the f16 math mode) --
    # half the bytes of the reference's fp32 arrays, no conversion pass; the fp32-host number is reported beside it.
    fp16_host = (w["arch"] == "BUTD" or bool(w.get("refiner"))) and math == "f16" and not scst
    e2e_s, tok_host = timed_e2e(host_feats)
    h2d = host_feats.numel() * host_feats.element_size()
    e2e_fp32 = {"value": n_total * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "ms_per_step": 1e3 * e2e_s / steps,
                "host_format": "fp32 arrays (the reference's format)"}
    d2h = B * L_out * 4 * (1 + (world if world > 1 else 0))  # the rank's own captions per batch (+ its copy of the gathered set)
    e2e = dict(e2e_fp32)
    host_f16 = None
    if fp16_host:
        host_f16 = host_feats.half().pin_memory()
        f16_s, tok_host = timed_e2e(host_f16)
        e2e = {"value": n_total * steps / f16_s, "unit": UNIT, "h2d_bytes_per_step": host_f16.numel() * 2,
               "ms_per_step": 1e3 * f16_s / steps, "host_format": "fp16 feature-shard rows (feature_store.py)"}
    e2e["d2h_bytes_per_step"] = d2h

    # ---- per-kernel timing (CUDA events on the launching stream) for the roofline object
    peaks = load_peaks()
    gather.reset()
    dec.profile(True)
    prof_steps = 2
    for _ in range(prof_steps):
        step_device()
    torch.cuda.synchronize()
    prof = dec.profile_read()
    dec.profile(False)
    gather.reset()
    kern = {}
    for cat, (ms, fl, cnt) in prof.items():
        if cnt:
            kern[cat] = {"ms_per_step": ms / prof_steps, "launches_per_step": cnt // prof_steps,
                         "tflops": (fl / (ms * 1e-3) / 1e12) if (fl and ms) else None}
    gemm_ms = {c: v["ms_per_step"] for c, v in kern.items() if c.startswith("gemm")}
    dom = max(gemm_ms, key=gemm_ms.get)
    dms, dfl, dcnt = prof[dom]
    achieved = dfl / (dms * 1e-3) / 1e12
    traffic, traffic_note = None, None  # DRAM bytes per launch of the dominant kernel from the round's ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and headline and wname == "butd_det" and B == WORKLOADS["butd_det"]["batch"] and math == "f16":
        tj = json.load(open(tpath))
        if tj.get("kernel_sources_sha") == kernel_sources_sha():
            traffic = tj.get(dom, {}).get("mean_per_launch")
            traffic_note = f"ncu --set full capture {tj.get('capture', '')} (tools/ncu_traffic.py), same kernel sources"
        else:
            traffic_note = "profiles/ncu_traffic.json was captured for other kernel sources (sha mismatch): not reported"
    roofline = {"kernel": f"capdec::gemm2_kernel<{dom}> (tcgen05 cta_group::2, 256x256 tiles)", "bound": "tensor", "achieved": achieved,
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peaks["source"] + ", sustained bf16 cuBLAS",
                "flops_per_launch": dfl / dcnt, "us_per_launch": 1e3 * dms / dcnt,
                # event-timed kernel time per step over the graph-replayed step time of the timed region
                "share_of_step": dms / prof_steps / ms_per_step}
    if "attention" in kern:  # HBM-bound companion kernel: algorithmic bytes = R*(A+D)*4 per image-step
        esz = 2.0 if (math == "f16" and w["arch"] == "BUTD") else 4.0  # fp16 mode reads the fp16 feature copies
        a_bytes = B * R * (dims.get("atten_dim", 0) + dims.get("enc_dim", dims["hidden_dim"] * 2)) * esz * T if w["arch"] == "BUTD" \
            else B * R * 2 * dims["hidden_dim"] * 4.0 * T
        gbs = a_bytes / (kern["attention"]["ms_per_step"] * 1e-3) / 1e9
        kern["attention"].update({"algorithmic_gbs": gbs, "hbm_frac": gbs / peaks["hbm_gbs"]})

    small = None
    if B * (K + 1 if scst else K) <= 128 and w["arch"] == "BUTD":
        # small-batch regime: a decode step streams the packed fp16 weights once -- top-down gates [4H, 2H], dec_att [A, H],
        # language gates [4H, D + 2H], vocabulary [V, H] -- and little else; bound = HBM (the weights do not stay in L2)
        H_, A_, D_, V_ = dims["hidden_dim"], dims["atten_dim"], dims["enc_dim"], dims["vocab_size"]
        wbytes = 2.0 * (4 * H_ * 2 * H_ + A_ * H_ + 4 * H_ * (D_ + 2 * H_) + V_ * H_)
        gbs = wbytes * T / (ms_per_step * 1e-3) / 1e9
        small = {"bound": "hbm", "weight_bytes_per_decode_step": wbytes, "decode_steps": T, "achieved": gbs, "peak": load_peaks()["hbm_gbs"],
                 "unit": "GB/s", "frac": gbs / load_peaks()["hbm_gbs"], "us_per_decode_step": 1e3 * ms_per_step / T,
                 "note": "swap-AB split-K kernel (csrc/smallm.cuh); a step is 4 dependent launches / 7 dependent phases"}
    out = {
        "value": value, "ms_per_step": ms_per_step, "steps": steps, "launches": launches, "clocks": clocks, "small_batch_roofline": small,
        "dtype": "f16 operands, f32 accumulate" if math == "f16" else "f16x3 split (fp32-grade), f32 accumulate",
        "config": {"workload": w["desc"], "images_per_gpu": B, "global_batch": n_total, "beam": K, "max_seq": T, "regions": R,
                   "vocab": dims["vocab_size"], "math": math,
                   "parallelism": f"dp{world} (images sharded, no data-path collective, one caption all-gather per run)",
                   "host_numa_node": ctx.numa_node,
                   "l2": f"inputs larger than L2: {h2d / 1e6:.0f} MB of features + {sum(int(np.prod(v.shape)) for v in sd.values()) * 2 / 1e6:.0f} MB "
                         "of fp16 weights are re-read every step" if h2d > 126e6 else
                         f"small batch: {h2d / 1e6:.1f} MB of features + fp16 weights stay L2-resident between steps (the reference's "
                         "own evaluation shape; launch / latency bound)"},
        "e2e": e2e, "e2e_fp32_host": e2e_fp32 if fp16_host else None, "roofline": roofline, "kernels": kern,
        "graph_captures": dec.graph_captures,
    }

    if headline:  # isolated host->device ceiling (all ranks at once, no decode) next to the end-to-end number
        agg, per = measure_h2d_ceiling(ctx, host_f16 if host_f16 is not None else host_feats)
        e2e_gbs = e2e["h2d_bytes_per_step"] * world / (e2e["ms_per_step"] * 1e-3) / 1e9
        out["h2d_ceiling"] = {"aggregate_gbs": agg, "per_gpu_gbs": per, "e2e_h2d_gbs": e2e_gbs, "e2e_frac_of_ceiling": e2e_gbs / agg,
                              "note": "pinned host -> device copies of one batch's features on all ranks concurrently, nothing else running"}

    if world > 1 and headline and not scst:
        # the N-GPU result must be bit-identical per image to a 1-GPU run, rows in original image order.
        # Rank 0 alone re-decodes every rank's shard (inputs regenerated from the seeds) and compares with the gathered tensor.
        ident = ident_e2e = None
        if rank == 0:
            ident = ident_e2e = True
            e2e_all = torch.from_numpy(tok_host).to(dev)  # [world * B, L]: last batch of the end-to-end run, gathered
            for r in range(world):
                f = torch.from_numpy(make_inputs(w, B, 1000 + r)).to(dev)
                cap._prepare(cap.feature_fn({key: f}) if w.get("images") else f, None)
                tok_r, _, _ = dec.beam_search(K, T)
                for s_ in (0, steps - 1):
                    ident = ident and bool((gathered[s_, r * B:(r + 1) * B] == tok_r).all().item())
                if fp16_host:  # the end-to-end run fed fp16 rows: compare with a 1-GPU decode of the same rows
                    cap._prepare(f.half(), None)
                    tok_r, _, _ = dec.beam_search(K, T)
                ident_e2e = ident_e2e and bool((e2e_all[r * B:(r + 1) * B] == tok_r).all().item())
                del f
        out["identity"] = {"device_run": ident, "e2e_run": ident_e2e,
                           "what": "rank 0 alone re-decodes every rank's shard; the gathered captions must be bit-identical, rows in image order"}
        dist.barrier()

    if rank == 0 and world == 1 and cpu_images > 0:
        from oracle import capdec_oracle as orc
        n_cpu = min(cpu_images, B)
        feats_np = host_feats[:n_cpu].numpy()
        if scst:  # parity of the greedy rows of the one-pass rollout with the oracle's greedy rollout
            _, o = oracle_decoder(w, sd)
            t0 = time.perf_counter()
            o.prepare(feats_np)
            want, _, _ = orc.greedy_sample(o, T)
            dt = time.perf_counter() - t0
            same = (last_greedy[0][:n_cpu].cpu().numpy() == want).all(1)
            out["parity_sample"] = {"images": n_cpu, "exact": int(same.sum()), "tie_justified": 0, "diff": int((~same).sum()),
                                    "what": "greedy rows of capdec_scst_rollout vs the oracle's greedy rollout"}
        else:
            dt, res = time_reference_form(w, sd, feats_np, K, T)
            gpu_tok = tokens[:n_cpu].cpu().numpy()
            verdict = orc.agreement(gpu_tok, res.tokens, res.min_gap, tol=1e-4)
            out["cpu_baseline"] = {"value": n_cpu / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"first {n_cpu} images of the same batch, one image per call (reference form), "
                                             f"numpy/BLAS on {os.cpu_count()} threads, {dt:.1f} s",
                                   "port_vs_reference": PORT_VS_REFERENCE}
            out["parity_sample"] = {"images": n_cpu, "exact": sum(v == "exact" for v in verdict),
                                    "tie_justified": sum(v == "tie" for v in verdict), "diff": sum(v == "diff" for v in verdict)}
    # release this workload's device memory before the next one
    dec.close()
    if reward is not None:
        reward.close()
    del cap, dec, dev_feats, host_feats, host_f16, gather, gathered
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args, w):
    ctx = GpuContext(args)
    head = measure(ctx, args, args.workload, args.batch, args.math, args.steps, args.warmup,
                   0 if args.no_cpu_baseline else args.cpu_images, headline=True)
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": head["dtype"], "data": "synthetic", "config": head["config"], "clocks": head["clocks"], "e2e": head["e2e"],
        "e2e_fp32_host": head["e2e_fp32_host"], "gpu_launches": head["launches"], "roofline": head["roofline"],
        "kernels": head["kernels"], "graph_captures": head["graph_captures"],
    }
    for k in ("h2d_ceiling", "identity", "cpu_baseline", "parity_sample", "small_batch_roofline"):
        if head.get(k) is not None:
            line[k] = head[k]
    default_run = args.workload == "butd_det" and args.batch == WORKLOADS["butd_det"]["batch"] and args.math == "f16"
    try:  # whatever happens in the extra passes, the headline line is printed
        if default_run and not args.no_extras:
            line["workloads"] = {}
            for name, wname, batch, math in EXTRAS:
                b = batch or WORKLOADS[wname]["batch"]
                n_steps = args.extra_steps if b >= 128 else 10 * args.extra_steps  # a 16-image batch takes ~1 ms
                try:
                    r = measure(ctx, args, wname, b, math, n_steps, 3, 0 if args.no_cpu_baseline else args.extra_cpu_images, headline=False)
                except Exception as exc:  # noqa: BLE001  (one workload must not take the headline line down with it)
                    line["workloads"][name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                    continue
                dom = r["roofline"]
                line["workloads"][name] = {
                    "value": r["value"], "unit": UNIT if wname != "scst" else "image rollout sets/s (5 samples + greedy + CIDEr-D reward)",
                    "ms_per_step": r["ms_per_step"], "steps": n_steps, "images_per_gpu": b, "global_batch": b * ctx.world, "math": math,
                    "workload": r["config"]["workload"], "e2e": r["e2e"], "clocks": r["clocks"], "gpu_launches": r["launches"],
                    "dominant_kernel": {"kernel": dom["kernel"], "frac": dom["frac"], "achieved_tflops": dom["achieved"],
                                        "us_per_launch": dom["us_per_launch"], "share_of_step": dom["share_of_step"]},
                    "kernels": {c: {"ms_per_step": round(v["ms_per_step"], 4), "tflops": v["tflops"] and round(v["tflops"], 1),
                                    **({"hbm_frac": round(v["hbm_frac"], 3)} if "hbm_frac" in v else {})} for c, v in r["kernels"].items()},
                    "parity_sample": r.get("parity_sample"), "cpu_baseline": r.get("cpu_baseline"),
                    **({"small_batch_roofline": r["small_batch_roofline"]} if r.get("small_batch_roofline") else {}),
                }
    except BaseException as exc:  # noqa: BLE001
        line["workloads_aborted"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        if ctx.rank == 0:
            emit(line)
    if ctx.world > 1:
        try:
            ctx.dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="butd_det", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU")
    ap.add_argument("--max-seq", type=int, default=20)
    ap.add_argument("--math", default="f16", choices=["f16", "f16x3"])
    ap.add_argument("--cpu-images", type=int, default=64, help="images of the same batch the CPU port decodes (about 10 s)")
    ap.add_argument("--ref-images", type=int, default=4, help="images per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short passes over the other BASELINE configs / the f16x3 mode")
    ap.add_argument("--extra-steps", type=int, default=8, help="timed steps of each extra workload (x10 for batches below 128)")
    ap.add_argument("--extra-cpu-images", type=int, default=8, help="images of each extra workload the CPU port decodes as a parity sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = dict(WORKLOADS[args.workload])
    if args.batch is None:
        args.batch = w["batch"]
    if args.impl == "reference":
        return run_reference_arm(args, w)
    return run_gpu_arm(args, w)


if __name__ == "__main__":
    sys.exit(main())
