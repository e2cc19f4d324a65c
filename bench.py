#!/usr/bin/env python
"""bench.py - audio samples/sec of the mel -> wav generator forward (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload hifigan_b64] [--impl reference]

One process per GPU (torchrun for N > 1: RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the env).  A "step" is one
forward of the generator over one batch of synthetic mel.  Rank 0 prints ONE JSON line (see README / DESIGN.md).
Multi-GPU = embarrassingly parallel batch split (weak scaling, no data-path collective; NCCL only for the
barrier and the max-over-ranks reduction of the device time).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, batch, n_mels, T, hop, sample_rate, description)
    "hifigan_b64": ("hifigan", 64, 80, 94, 256, 24000,
                    "hifigan generator (80 mel, hop 256, rates 8-8-2-2, ch 512), batch 64 x 1 s @ 24 kHz"),
    "hifigan_b1": ("hifigan", 1, 80, 94, 256, 24000, "hifigan baseline generator, batch 1 x 1 s @ 24 kHz"),
    "bigvgan_b32": ("bigvgan", 32, 100, 87, 512, 44100,
                    "bigvgan generator with anti-aliased Snake (100 mel, hop 512), batch 32 x 1 s @ 44.1 kHz"),
    "hifigan_yaml_b32": ("hifigan5", 32, 128, 87, 512, 44100,
                         "hifigan.yaml as written (128 mel, hop 512, rates 8-8-2-2-2, ch 512), batch 32 x 1 s @ 44.1 kHz"),
    "firefly_b32": ("firefly", 32, 128, 87, 512, 44100,
                    "firefly-gan-base (ConvNeXt [3,3,9,3]x[128,256,384,512] + HiFiGAN head, rates 8-8-2-2-2, k13 pre/post), "
                    "batch 32 x 1 s @ 44.1 kHz"),
    "vocos_huge_b128": ("vocos", 128, 100, 94, 256, 24000,
                        "vocos_huge ConvNeXt [3,3,27,3]x[352,704,1408,2816] + ISTFT(1024/256), batch 128 x 1 s @ 24 kHz"),
}


def build_model(kind: str):
    from vocoder_b200.encoders import ConvNeXtEncoder
    from vocoder_b200.generators import BigVGANGenerator, HiFiGANGenerator, ISTFTHead, UnifyGenerator
    torch.manual_seed(0)  # "ref-init": the reference constructor's own initialisation, no checkpoint offline
    if kind == "hifigan":
        return HiFiGANGenerator(hop_length=256, upsample_rates=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4),
                                resblock_kernel_sizes=(3, 7, 11), resblock_dilation_sizes=((1, 3, 5),) * 3,
                                num_mels=80, upsample_initial_channel=512, use_template=False,
                                pre_conv_kernel_size=7, post_conv_kernel_size=7)
    if kind == "hifigan5":  # configs/model/generator/hifigan.yaml + resolution/44100_512_2048.yaml (SURVEY 8d variant A')
        return HiFiGANGenerator(hop_length=512, upsample_rates=(8, 8, 2, 2, 2), upsample_kernel_sizes=(16, 16, 8, 2, 2),
                                resblock_kernel_sizes=(3, 7, 11), resblock_dilation_sizes=((1, 3, 5),) * 3,
                                num_mels=128, upsample_initial_channel=512, use_template=False,
                                pre_conv_kernel_size=7, post_conv_kernel_size=7)
    if kind == "bigvgan":
        return BigVGANGenerator(hop_length=512, num_mels=100, use_template=False)
    if kind == "firefly":  # configs/model/generator/firefly-gan-base.yaml + resolution/44100_512_2048.yaml
        return UnifyGenerator(
            backbone=ConvNeXtEncoder(input_channels=128, depths=[3, 3, 9, 3], dims=[128, 256, 384, 512],
                                     drop_path_rate=0.2, kernel_size=7),
            head=HiFiGANGenerator(hop_length=512, upsample_rates=(8, 8, 2, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4, 4),
                                  resblock_kernel_sizes=(3, 7, 11), resblock_dilation_sizes=((1, 3, 5),) * 3,
                                  num_mels=512, upsample_initial_channel=512, use_template=False,
                                  pre_conv_kernel_size=13, post_conv_kernel_size=13))
    if kind == "vocos":
        return UnifyGenerator(
            backbone=ConvNeXtEncoder(input_channels=100, depths=[3, 3, 27, 3], dims=[352, 704, 1408, 2816],
                                     drop_path_rate=0.4, kernel_size=7),
            head=ISTFTHead(dim=2816, n_fft=1024, hop_length=256, win_length=1024, padding="same"))
    raise KeyError(kind)


def synthetic_mel(B, n_mels, T, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.empty(B, n_mels, T).uniform_(-11.5129, 2.0, generator=g)


def oracle_forward(kind, sd, mel, model):
    from oracle import generators as G
    if kind in ("hifigan", "hifigan5"):
        return G.hifigan_forward(sd, mel, model.upsample_rates)
    if kind == "bigvgan":
        return G.bigvgan_forward(sd, mel, model.upsample_rates)
    if kind == "firefly":
        return G.unify_hifigan_forward(sd, mel, model.head.upsample_rates)
    return G.unify_vocos_forward(sd, mel, 1024, 256, 1024)


def time_cpu_port(kind, model, n_mels, T, hop, sample_B, steps=1, warmup=1):
    """The reference's CPU PyTorch path (oracle port: same torch ops the reference modules call) on host cores."""
    sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
    mel = synthetic_mel(sample_B, n_mels, T, 1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle_forward(kind, sd, mel, model)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return sample_B * T * hop / min(times), sum(times) / len(times), min(times)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tensor_tflops": p["bf16_tflops"],
                "tensor_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1590.0, "tensor_tflops_sustained": 1400.0, "source": "fallback"}


def _alg_conv1d(a16, pc, L_out=None, **kw):
    B, L_in, _ = a16.shape
    rows = B * (L_in if L_out is None else L_out)
    flops = 2.0 * rows * pc.c_out * pc.c_in * pc.n_taps
    byts = 2.0 * B * L_in * pc.c_in + 2.0 * pc.n_phase * pc.n_taps * pc.c_out * pc.c_in
    if kw.get("residual") is not None:
        byts += 4.0 * rows * pc.c_out
    if kw.get("out32") is not None:
        byts += 4.0 * rows * pc.c_out * (2 if kw.get("accumulate") else 1)
    if kw.get("out16") is not None:
        byts += 2.0 * rows * pc.c_out
    return flops, byts


def _alg_mrf_fused(x32, pm, out32, **kw):
    # whole MRF stage: x read once, mean written once (+ fp16 operand of the next layer), every tap tile read once
    B, L, _ = x32.shape
    taps = sum(k * 2 * len(pm.dil1[j]) for j, k in enumerate(pm.ksize))
    flops = 2.0 * B * L * pm.C * pm.C * taps
    byts = 8.0 * B * L * pm.C + 2.0 * taps * pm.C * pm.C + (2.0 * B * L * pm.C if kw.get("out16") is not None else 0.0)
    return flops, byts


def _alg_snake(x32, out16, alpha, beta, fu, fd, C, *a, **kw):
    B, L, _ = x32.shape
    return 0.0, 6.0 * B * L * C


def _alg_dwln(x32, C, *a, out16=None, out32=None, **kw):
    B, T, _ = x32.shape
    return 0.0, B * T * C * (4.0 + (2.0 if out16 is not None else 0.0) + (4.0 if out32 is not None else 0.0))


def _alg_post(a16, w32, bias, C, *a, **kw):
    B, L, _ = a16.shape
    return 2.0 * B * L * C * w32.shape[0], 2.0 * B * L * C + 4.0 * B * L


def _alg_pack(x, *a, **kw):
    return 0.0, 6.0 * x.numel()


def _alg_ola(frames, window, n_fft, hop, *a, **kw):
    B, T, _ = frames.shape
    return 0.0, 4.0 * B * T * (n_fft + hop)


# C-ABI wrapper -> (kernel family, algorithmic (flops, bytes) of one launch); formulas in DESIGN.md section 4
FAMILIES = {"conv1d": ("conv_tc_kernel", _alg_conv1d), "mrf_fused": ("mrf_fused_kernel", _alg_mrf_fused),
            "snake_aa": ("snake_aa_kernel", _alg_snake), "dwconv_layernorm": ("dwconv_ln_kernel", _alg_dwln),
            "conv_post_tanh": ("conv_post_kernel", _alg_post), "pack_input": ("pack_input_kernel", _alg_pack),
            "istft_ola": ("istft_ola_kernel", _alg_ola)}


def profile_launches(model, mel):
    """One instrumented eager forward: CUDA-event duration (events recorded on the launching stream) and algorithmic
    flops/bytes of every launch of the kernel families above."""
    from vocoder_b200 import cabi
    recs, saved = [], {}
    stream = torch.cuda.current_stream()

    def make(name, fam, alg, orig):
        def wrapped(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = orig(*a, **kw)
            e1.record(stream)
            fl, by = alg(*a, **kw)
            recs.append((fam, e0, e1, fl, by))
            return out
        return wrapped

    for name, (fam, alg) in FAMILIES.items():
        saved[name] = getattr(cabi, name)
        setattr(cabi, name, make(name, fam, alg, saved[name]))
    try:
        with torch.no_grad():
            model(mel)
        torch.cuda.synchronize()
    finally:
        for name, fn in saved.items():
            setattr(cabi, name, fn)
    return [{"kernel": fam, "ms": e0.elapsed_time(e1), "flops": fl, "bytes": by} for fam, e0, e1, fl, by in recs]


def roofline_of(recs, peaks, step_ms):
    """Per-family totals; the roofline object describes the family with the largest share of the step."""
    fams = {}
    for r in recs:
        f = fams.setdefault(r["kernel"], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "t_roof": 0.0})
        f["launches"] += 1
        f["ms"] += r["ms"]
        f["flops"] += r["flops"]
        f["bytes"] += r["bytes"]
        f["t_roof"] += max(r["flops"] / (peaks["tensor_tflops"] * 1e12), r["bytes"] / (peaks["hbm_gbs"] * 1e9))
    table = {}
    for k, f in fams.items():
        t = f["ms"] * 1e-3
        table[k] = {"launches": f["launches"], "ms_per_step": f["ms"], "share_of_step": f["ms"] / step_ms,
                    "alg_tflop": f["flops"] / 1e12, "alg_gbytes": f["bytes"] / 1e9,
                    "tflops": f["flops"] / t / 1e12, "gbs": f["bytes"] / t / 1e9,
                    "roofline_frac": f["t_roof"] / t}
    top = max(fams, key=lambda k: fams[k]["ms"])
    f = fams[top]
    t = f["ms"] * 1e-3
    t_tensor = f["flops"] / (peaks["tensor_tflops"] * 1e12)
    t_hbm = f["bytes"] / (peaks["hbm_gbs"] * 1e9)
    if t_hbm >= t_tensor:
        roof = {"bound": "hbm", "achieved": f["bytes"] / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": f["flops"] / t / 1e12, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None  # filled by main() from profiles/r01_traffic.json (ncu --set full capture of a launch of this family)
    roof.update({"kernel": f"{top} (all {f['launches']} launches of one step)", "launches": f["launches"],
                 "kernel_ms_per_step": f["ms"], "share_of_step": f["ms"] / step_ms,
                 "alg_tflop_per_step": f["flops"] / 1e12, "alg_gbytes_per_step": f["bytes"] / 1e9,
                 "tensor_frac": f["flops"] / t / 1e12 / peaks["tensor_tflops"],
                 "hbm_frac": f["bytes"] / t / 1e9 / peaks["hbm_gbs"],
                 "per_launch_roofline_frac": f["t_roof"] / t, "peak_source": peaks["source"] + " (burst)",
                 "families": table})
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hifigan_b64", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fuse-mrf", action="store_true", help="layer-wise fv_conv1d launches instead of fv_mrf_fused")
    ap.add_argument("--tc-tuning", default="", help="block_n,m_sub,epilogue,mainloop overrides for fv_conv1d (0 = auto)")
    ap.add_argument("--mrf-silu-exact", action="store_true", help="ex2+rcp SiLU inside fv_mrf_fused instead of tanh.approx")
    ap.add_argument("--cpu-sample", type=int, default=0, help="utterances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch conv timing table (json) here")
    ap.add_argument("--micro-batch", type=int, default=0,
                    help="utterances per residual-block pass (0 = whole batch, the measured optimum)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    kind, B, n_mels, T, hop, sr, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    samples_per_step = B * T * hop
    config = {"workload": f"{args.workload}: {desc}", "batch_per_gpu": B, "mel_shape": [B, n_mels, T],
              "samples_per_step_per_gpu": samples_per_step, "weights": "ref-init (seed 0), no checkpoint offline",
              "parallelism": f"batch-shard x{world}" if world > 1 else "single GPU"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        model = build_model(kind).eval()
        sample_B = args.cpu_sample or max(1, min(B, 8))
        sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
        mel = synthetic_mel(sample_B, n_mels, T, 1234)
        with torch.no_grad():
            for _ in range(max(1, min(args.warmup, 1))):
                oracle_forward(kind, sd, mel, model)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                oracle_forward(kind, sd, mel, model)
            dt = time.perf_counter() - t0
        val = sample_B * T * hop * args.steps / dt
        cores = torch.get_num_threads()
        line = {"impl": "reference", "metric": "audio samples/sec, mel->wav generator forward", "value": val,
                "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample_B} of {B} utterances per step (oracle port of the reference's "
                                           f"CPU PyTorch path, {cores} threads)"},
                "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "rtf": val / sr}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from vocoder_b200 import cabi

    if args.tc_tuning:
        cabi.set_tc_tuning(*[int(v) for v in args.tc_tuning.split(",")])
    model = build_model(kind).eval().to(dev)
    for m in model.modules():
        if hasattr(m, "use_cuda_graph"):
            m.use_cuda_graph = False
        if args.micro_batch > 0 and hasattr(m, "micro_batch"):
            m.micro_batch = args.micro_batch
        if args.no_fuse_mrf and hasattr(m, "fuse_mrf"):
            m.fuse_mrf = False
        if args.mrf_silu_exact and hasattr(m, "mrf_silu_tanh"):
            m.mrf_silu_tanh = False
    mel_host = synthetic_mel(B, n_mels, T, 1234 + rank).pin_memory()
    mel = mel_host.to(dev)
    wav_host = torch.empty(B, 1, T * hop).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        cabi.reset_launch_count()
        y = model(mel)
        torch.cuda.synchronize()
        launches_per_step = cabi.launch_count()
        assert y.shape == (B, 1, T * hop) and bool(torch.isfinite(y).all())
        if not args.no_graph:
            model.use_cuda_graph = True
        for _ in range(args.warmup):
            model(mel)
        torch.cuda.synchronize()

        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        # ---- device-resident timing: K steps, L2 flushed between steps, device time summed per step
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        wall0 = time.perf_counter()
        for e0, e1 in evs:
            flush.zero_()
            e0.record()
            model(mel)
            e1.record()
        barrier()
        wall = time.perf_counter() - wall0
        dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)

        # ---- end to end: pinned host mel -> H2D -> forward -> D2H wav, every step
        for _ in range(2):
            wav_host.copy_(model(mel_host.to(dev, non_blocking=True)), non_blocking=True)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            x = mel_host.to(dev, non_blocking=True)
            wav_host.copy_(model(x), non_blocking=True)
        s1.record()
        barrier()
        e2e_ms = s0.elapsed_time(s1)
        clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_samples = samples_per_step * args.steps * world
    value = total_samples / (dev_ms * 1e-3)
    e2e_value = total_samples / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = load_peaks()
        model.use_cuda_graph = False
        with torch.no_grad():
            for _ in range(2):
                recs = profile_launches(model, mel)
        if args.dump_launches:
            with open(args.dump_launches, "w") as f:
                json.dump(recs, f)
        roofline = roofline_of(recs, peaks, dev_ms / args.steps)
        try:  # DRAM bytes of a profiled launch of the dominant family (ncu --set full, committed under profiles/)
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(args.workload)
            if tr and tr.get("traffic_bytes") and roofline["kernel"].startswith(tr["kernel"]):
                roofline["traffic"] = tr["traffic_bytes"]
                roofline["traffic_ref"] = {k: tr[k] for k in ("launch", "algorithmic_bytes", "source")}
        except (OSError, ValueError):
            pass

        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(max(1, os.cpu_count() or 1))
            sample_B = args.cpu_sample or max(1, min(B, 8))
            v, mean_s, best_s = time_cpu_port(kind, model, n_mels, T, hop, sample_B, steps=2, warmup=1)
            cpu_baseline = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{sample_B} of {B} utterances, best of 2 after 1 warm-up "
                                      f"({best_s:.2f} s per forward; oracle port of the reference CPU PyTorch path)"}
        config["l2"] = ("L2 flushed between timed steps (256 MiB memset outside the per-step event pairs); per-step "
                        f"working set {sum(m._ws.nbytes() for m in model.modules() if hasattr(m, '_ws')) / 1e9:.2f} GB > 126 MB L2")
        config["cuda_graph"] = not args.no_graph
        config["fuse_mrf"] = not args.no_fuse_mrf
        config["mrf_silu"] = "ex2+rcp" if args.mrf_silu_exact else "tanh.approx"
        config["micro_batch"] = args.micro_batch if args.micro_batch > 0 else "whole batch"
        line = {"metric": "audio samples/sec, mel->wav generator forward", "value": value, "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 operands (TF32-grade mantissa), f32 accumulate/residual stream", "data": "synthetic",
                "config": config, "rtf_per_gpu": value / world / sr,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": mel_host.numel() * 4,
                        "d2h_bytes_per_step": wav_host.numel() * 4, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
                "wall_s_timed_region": wall, "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
