"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product package).

Functional, state-dict driven restatement (plain torch fp32 on CPU) of the reference's mel front-end,
fish_vocoder/data/transforms/spectrogram.py.  Pinned by tests/test_oracle_cpu.py against golden vectors produced by
running the *unmodified* reference classes (torchaudio's MelScale included) in the build container
(oracle/make_golden_frontend.py -> tests/golden/frontend_*.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def linear_spectrogram(y: Tensor, window: Tensor, n_fft: int, hop_length: int, win_length: int) -> Tensor:
    """LinearSpectrogram.forward (spectrogram.py:25-57), center=False, mode="pow2_sqrt": [B, L] -> [B, n_fft/2+1, T]."""
    if y.ndim == 3:
        y = y.squeeze(1)
    y = F.pad(y.unsqueeze(1), ((win_length - hop_length) // 2, (win_length - hop_length + 1) // 2),
              mode="reflect").squeeze(1)                                                   # spectrogram.py:29-37
    spec = torch.stft(y, n_fft, hop_length=hop_length, win_length=win_length, window=window, center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)  # :39-50
    spec = torch.view_as_real(spec)
    return torch.sqrt(spec.pow(2).sum(-1) + 1e-6)                                           # :54-55


def log_mel_spectrogram(y: Tensor, window: Tensor, fb: Tensor, n_fft: int, hop_length: int, win_length: int) -> Tensor:
    """LogMelSpectrogram.forward (spectrogram.py:99-104): MelScale is `fb^T @ spec` over the frequency axis
    (torchaudio.transforms.MelScale.forward: matmul(spec.transpose(-1,-2), fb).transpose(-1,-2)), then log(clamp(1e-5))."""
    spec = linear_spectrogram(y, window, n_fft, hop_length, win_length)
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
    return torch.log(torch.clamp(mel, min=1e-5))                                            # :93-94
