#!/usr/bin/env python
"""Micro-benchmark of fv_conv1d at one layer shape over the tuning space (N tile, accumulators per CTA, epilogue, mainloop):
python tools/bench_conv.py [--C 128 --L 6016 --B 64]   (CUDA events, L2 flushed between launches, median of 6)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vocoder_b200 import cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=128)
    ap.add_argument("--L", type=int, default=6016)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--ks", default="3,7,11")
    ap.add_argument("--configs", default="0:0:0:0,0:0:0:2,128:1:0:0,64:2:0:0,64:4:0:0")
    args = ap.parse_args()
    B, L, C = args.B, args.L, args.C
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    a = torch.randn(B, L, C, device="cuda").half()
    res = torch.randn(B, L, C, device="cuda")
    o16 = torch.empty(B, L, C, device="cuda", dtype=torch.float16)
    o32 = torch.empty(B, L, C, device="cuda")
    for k in (int(v) for v in args.ks.split(",")):
        conv = torch.nn.Conv1d(C, C, k, padding=(k - 1) // 2).cuda()
        pc = cabi.pack_conv(conv.weight, conv.bias, 1)
        gf = 2.0 * B * L * C * C * k / 1e9
        for cfg in args.configs.split(","):
            cabi.set_tc_tuning(*[int(v) for v in cfg.split(":")])
            row = []
            for kind in ("convs1", "convs2"):
                ts = []
                try:
                    for _ in range(6):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        if kind == "convs1":
                            cabi.conv1d(a, pc, out16=o16, act=cabi.ACT_SILU_TANH)
                        else:
                            cabi.conv1d(a, pc, residual=res, out32=o32, out16=o16, act=cabi.ACT_SILU_TANH)
                        e1.record()
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    t = sorted(ts)[len(ts) // 2]
                    row.append(f"{kind} {t * 1e3:7.1f} us {gf / t:7.1f} TF/s")
                except cabi.FvError as e:
                    row.append(f"{kind} unsupported")
            print(f"C={C} L={L} B={B} k={k:2d} cfg={cfg:12s} " + " | ".join(row), flush=True)
        cabi.set_tc_tuning(0, 0, 0, 0)


if __name__ == "__main__":
    main()
