#!/usr/bin/env python
"""Run fv_debug_umma_rate over (mode, N, background traffic) and print SM cycles per tcgen05.mma (M = 128 per CTA, K = 16,
fp16) next to the tensor-pipe floor N/2 (128 x N x 16 MACs at 4096 MACs/clk/SM).  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vocoder_b200 import cabi  # noqa: E402


def fma_rates(L):
    """fp32 FMA throughput of the CUDA cores (roofline of the anti-aliased Snake kernel), CUDA-event timed."""
    sink = torch.zeros(4, device="cuda")
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    iters = 20000
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    print("\n| fp32 FMA form | TFLOP/s (2 flops per FMA) | FMA lanes / clk / SM (at the measured kernel clock) |")
    print("|---|---:|---:|")
    for variant, name in ((0, "scalar fma, 3 register operands"), (1, "scalar fma, constant multiplier"),
                          (2, "packed fma.rn.f32x2, broadcast constant multiplier")):
        for _ in range(2):
            L.fv_debug_fma_rate(variant, iters, sink.data_ptr(), out.data_ptr(), st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.fv_debug_fma_rate(variant, iters, sink.data_ptr(), out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fmas = sms * 8 * 256 * iters * 16.0   # 16 fp32 FMAs per thread and iteration in every variant (8 f32x2 = 16 lanes)
        cycles = float(out[1])
        print(f"| {name} | {2 * fmas / ms / 1e9:.1f} | {fmas / sms / max(cycles, 1.0):.1f} |", flush=True)


def main():
    L = cabi.lib()
    out = torch.zeros(256, dtype=torch.int32, device="cuda")
    fma_rates(L)
    print()
    names = {0: "SS cta_group::1", 1: "SS cta_group::2 (pair)", 2: "A in TMEM", 3: "SS weight-stationary (.ws)"}
    print("| operands | N | background smem traffic | cycles / UMMA (median over CTAs) | floor N/2 | ratio |")
    print("|---|---:|---|---:|---:|---:|")
    for mode in (0, 2, 3, 1):
        for n in ((64, 128, 256) if mode == 3 else (32, 64, 128, 256)):
            for bg in (0, 1, 2):
                out.zero_()
                rc = L.fv_debug_umma_rate(mode, n, 512, bg, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    print(f"| {names[mode]} | {n} | {bg} | error {rc}: {L.fv_last_error().decode()} | | |")
                    continue
                torch.cuda.synchronize()
                v = out[out > 0].float() / 1000.0
                med = float(v.median()) if v.numel() else float("nan")
                print(f"| {names[mode]} | {n} | {('none', 'st.shared', 'ld.shared')[bg]} | {med:.1f} | {n / 2:.0f} | "
                      f"{med / (n / 2):.2f} |", flush=True)


if __name__ == "__main__":
    main()
