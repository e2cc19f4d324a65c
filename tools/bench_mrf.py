#!/usr/bin/env python
"""Micro-benchmark of fv_mrf_fused at the HiFiGAN cfg-B stage shapes: CUDA-event time per launch (L2 flushed between
launches), tensor throughput, and max error of each inner-activation variant against the fp64 contract reference
(on a small slice).  Usage: python tools/bench_mrf.py [--acts silu,tanh] [--shapes 64x12032x64,32x24064x64]"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vocoder_b200 import cabi  # noqa: E402


def make_blocks(C, ks=(3, 7, 11), seed=0, dils=(1, 3, 5)):
    torch.manual_seed(seed)
    blocks = []
    for k in ks:
        mk = lambda d: torch.nn.Conv1d(C, C, k, dilation=d, padding=(k * d - d) // 2)
        c1s, c2s = [mk(d) for d in dils], [mk(1) for _ in dils]
        for c in c1s + c2s:
            c.weight.data.normal_(0, 0.5 / math.sqrt(C * k))
            c.bias.data.normal_(0, 0.05)
        blocks.append((c1s, c2s))
    return blocks


def reference(x, blocks):
    r = lambda t: t.float().half().double()
    xs = x.double().permute(0, 2, 1)
    total = torch.zeros_like(xs)
    for c1s, c2s in blocks:
        xk = xs.clone()
        for c1, c2 in zip(c1s, c2s):
            xt = r(F.silu(xk))
            xt = F.conv1d(xt, r(c1.weight), c1.bias.double(), padding=c1.padding[0], dilation=c1.dilation[0])
            xt = r(F.silu(xt))
            xt = F.conv1d(xt, r(c2.weight), c2.bias.double(), padding=c2.padding[0], dilation=c2.dilation[0])
            xk = xk + xt
        total += xk
    return (total / len(blocks)).permute(0, 2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--acts", default="silu,tanh,h2")
    ap.add_argument("--shapes", default="64x12032x64,32x24064x64,16x44544x32")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    acts = {"silu": cabi.ACT_SILU, "tanh": cabi.ACT_SILU_TANH, "leaky": cabi.ACT_LEAKY, "h2": cabi.ACT_SILU_H2}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for shp in args.shapes.split(","):
        C, L, B = (int(v) for v in shp.split("x"))
        blocks = make_blocks(C)
        pm = cabi.pack_mrf(C, blocks)
        pm.w, pm.bias = pm.w.cuda(), pm.bias.cuda()
        x = torch.randn(B, L, C, device="cuda")
        out32 = torch.empty(B, L, C, device="cuda")
        out16 = torch.empty(B, L, C, device="cuda", dtype=torch.float16)
        with torch.no_grad():
            want = reference(x[:1, :1500].cpu(), blocks)
        flops = 2.0 * B * L * C * C * sum(k * 6 for k in pm.ksize)
        for name in args.acts.split(","):
            a = acts[name]
            for _ in range(2):
                cabi.mrf_fused(x, pm, out32, out16=out16, act=a, act_param=0.1, out_act=cabi.ACT_SILU)
            ts = []
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                cabi.mrf_fused(x, pm, out32, out16=out16, act=a, act_param=0.1, out_act=cabi.ACT_SILU)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            err = float("nan")
            if name != "leaky":
                o1 = torch.empty(1, 1500, C, device="cuda")
                cabi.mrf_fused(x[:1, :1500].contiguous(), pm, o1, act=a)
                err = float((o1.cpu().double() - want).abs().max())
            print(f"mrf_fused C={C} L={L} B={B} act={name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  "
                  f"max|err| vs fp64 contract {err:.2e} (scale {float(want.abs().max()):.2f})", flush=True)


def main_pairs(iters=10, C=128, L=6016, B=64):
    """A stage of HiFiGAN cfg B (C = 128: L = 6016; C = 64: L = 12032; B = 64), pair by pair: time per (k, d) launch and for
    the whole stage."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    x = torch.randn(B, L, C, device="cuda")
    out32 = torch.empty(B, L, C, device="cuda")
    total = 0.0
    for k in (3, 7, 11):
        for d in (1, 3, 5):
            blocks = make_blocks(C, ks=(k,), dils=(d,))
            pm = cabi.pack_mrf(C, blocks)
            pm.w, pm.bias = pm.w.cuda(), pm.bias.cuda()
            for _ in range(2):
                cabi.mrf_fused(x, pm, out32, act=cabi.ACT_SILU_TANH, out_scale=1.0)
            ts = []
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                cabi.mrf_fused(x, pm, out32, act=cabi.ACT_SILU_TANH, out_scale=1.0)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            total += ms
            flops = 2.0 * B * L * C * C * 2 * k
            print(f"mrf_fused pair C={C} k={k} d={d}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s  "
                  f"{8.0 * B * L * C / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    print(f"C={C} stage, 9 pair launches: {total:.3f} ms")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pairs":
        main_pairs()
    elif len(sys.argv) > 1 and sys.argv[1] == "pairs64":
        main_pairs(C=64, L=12032)
    else:
        main()
