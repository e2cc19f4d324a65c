"""Generate tests/golden/frontend_*.npz by running the UNMODIFIED reference mel front-end
(fish_vocoder/data/transforms/spectrogram.py with this image's torchaudio MelScale) on CPU, fp32.

Run only in the build container (needs /root/reference):  python oracle/make_golden_frontend.py
Each fixture stores the constructor kwargs (json), the state_dict (window, mel filterbank), the audio input and both
outputs (linear magnitude spectrogram and log-mel).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = ["/root/reference"]
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from fish_vocoder.data.transforms.spectrogram import LogMelSpectrogram  # noqa: E402


def test_signal(B, L, sr, seed):
    """Audio-like signal: a few harmonics with vibrato, a decaying noise burst and a quiet noise floor (peak < 0.9)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float64) / sr
    y = torch.zeros(B, L, dtype=torch.float64)
    for b in range(B):
        f0 = 110.0 * (1 + b) * (1.0 + 0.01 * torch.sin(2 * torch.pi * 5.0 * t))
        ph = 2 * torch.pi * torch.cumsum(f0, 0) / sr
        for h in range(1, 9):
            y[b] += (0.5 / h) * torch.sin(h * ph + 0.3 * h)
        y[b] += 0.2 * torch.randn(L, generator=g, dtype=torch.float64) * torch.exp(-t * 6.0)
        y[b] += 1e-3 * torch.randn(L, generator=g, dtype=torch.float64)
    y = 0.85 * y / y.abs().max()
    return y.float()


@torch.no_grad()
def main():
    torch.set_num_threads(4)
    for name, kw, L in [
        ("frontend_24k", dict(sample_rate=24000, n_fft=1024, win_length=1024, hop_length=256, n_mels=100), 6144),
        ("frontend_44k", dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=128), 8192),
    ]:
        m = LogMelSpectrogram(**kw).eval()
        y = test_signal(2, L, kw["sample_rate"], seed=7)
        lin = m.spectrogram(y)
        mel = m(y)
        arrs = {"sd::" + k: v.numpy() for k, v in m.state_dict().items()}
        arrs["in::audio"] = y.numpy()
        arrs["out"] = mel.numpy()
        arrs["out_linear"] = lin.numpy()
        arrs["kwargs"] = np.frombuffer(json.dumps(kw).encode(), dtype=np.uint8)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrs)
        print(f"{name}: mel {tuple(mel.shape)} range [{mel.min():.3f}, {mel.max():.3f}] linear {tuple(lin.shape)} "
              f"{os.path.getsize(path) / 1e6:.2f} MB; state_dict keys {list(m.state_dict().keys())}")


if __name__ == "__main__":
    main()
