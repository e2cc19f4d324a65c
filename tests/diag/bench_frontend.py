#!/usr/bin/env python
"""Throughput of the mel front-end (audio -> log-mel) on the GPU path next to the oracle (torch.stft + matmul) on the host
cores: python tests/diag/bench_frontend.py [--batch 64] [--seconds 1.0]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import frontend as O  # noqa: E402  (checker / CPU baseline only)
from vocoder_b200 import cabi  # noqa: E402
from vocoder_b200.transforms import LogMelSpectrogram  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    for kw in (dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=80),
               dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=100)):
        hop = kw["hop_length"]
        L = int(round(args.seconds * kw["sample_rate"] / hop)) * hop
        g = torch.Generator().manual_seed(1)
        y = (torch.rand(args.batch, L, generator=g) * 1.6 - 0.8)
        m = LogMelSpectrogram(**kw).cuda()
        yc = y.cuda()
        with torch.no_grad():
            for _ in range(3):
                out = m(yc)
            torch.cuda.synchronize()
            cabi.reset_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                out = m(yc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            torch.set_num_threads(os.cpu_count() or 1)
            sd = {k: v.cpu() for k, v in m.state_dict().items()}
            t0 = time.perf_counter()
            for _ in range(3):
                ref = O.log_mel_spectrogram(y, sd["spectrogram.window"], sd["mel_scale.fb"], kw["n_fft"], hop,
                                            kw["win_length"])
            cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
        err = float((out.cpu() - ref).abs().max())
        print(json.dumps({"front_end": kw, "batch": args.batch, "samples_in": args.batch * L, "gpu_ms": ms,
                          "gpu_samples_per_s": args.batch * L / ms * 1e3, "launches_per_call": cabi.launch_count() // args.iters,
                          "cpu_ms": cpu_ms, "cpu_samples_per_s": args.batch * L / cpu_ms * 1e3,
                          "cpu_threads": torch.get_num_threads(), "max_abs_err_vs_oracle": err}))


if __name__ == "__main__":
    main()
