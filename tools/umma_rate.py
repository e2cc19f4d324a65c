#!/usr/bin/env python
"""Run fv_debug_umma_rate over (mode, N, background traffic) and print SM cycles per tcgen05.mma (M = 128 per CTA, K = 16,
fp16) next to the tensor-pipe floor N/2 (128 x N x 16 MACs at 4096 MACs/clk/SM).  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vocoder_b200 import cabi  # noqa: E402


def main():
    L = cabi.lib()
    out = torch.zeros(256, dtype=torch.int32, device="cuda")
    names = {0: "SS cta_group::1", 1: "SS cta_group::2 (pair)", 2: "A in TMEM", 3: "SS weight-stationary (.ws)"}
    print("| operands | N | background smem traffic | cycles / UMMA (median over CTAs) | floor N/2 | ratio |")
    print("|---|---:|---|---:|---:|---:|")
    for mode in (0, 1, 2, 3):
        for n in (32, 64, 128, 256):
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
