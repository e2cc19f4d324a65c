"""Mel front-end on the B200 path: the reference's `LinearSpectrogram` / `LogMelSpectrogram`
(fish_vocoder/data/transforms/spectrogram.py:6-104, used at test.py:71 and models/gan.py:284) with the same constructor
kwargs, buffers (`window`; `spectrogram.window`, `mel_scale.fb`) and output, so audio -> mel -> generator stays on the GPU.

    reflect pad ((win-hop)//2, (win-hop+1)//2) -> torch.stft(center=False, onesided) -> sqrt(re^2 + im^2 + 1e-6)
    -> slaney mel matmul -> log(clamp(., 1e-5))

No cuFFT: with the padded signal laid out as rows of `hop` samples the framed, windowed DFT is a conv over rows with
n_fft / hop taps whose weights are window * {cos, -sin} - the same tcgen05 implicit-GEMM kernel as every generator layer.
Both contractions run in the strict operand mode (fp16 hi + lo pairs, fp32-grade): a log-mel must resolve bins 100 dB below
the frame's peak, which the 11-bit fp16 operand grade of the generator path does not.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import cabi
from ..runtime import Workspace, forward_signature, params_key, require_cuda


def slaney_mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(..., norm="slaney", mel_scale="slaney") restated (what MelScale builds at
    spectrogram.py:78-86): [n_freqs, n_mels] triangles on the Slaney (linear below 1 kHz, log above) mel scale, each
    normalised by 2 / (f_right - f_left).  Only used when no checkpoint supplies `mel_scale.fb`."""
    def hz_to_mel(f):
        f = torch.as_tensor(f, dtype=torch.float64)
        f_sp, min_log_hz = 200.0 / 3, 1000.0
        min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
        return torch.where(f >= min_log_hz, min_log_mel + torch.log(torch.clamp(f, min=1e-10) / min_log_hz) / logstep,
                           f / f_sp)

    def mel_to_hz(m):
        f_sp, min_log_hz = 200.0 / 3, 1000.0
        min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
        return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f_sp * m)

    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64)
    m_pts = torch.linspace(float(hz_to_mel(f_min)), float(hz_to_mel(f_max)), n_mels + 2, dtype=torch.float64)
    f_pts = mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)           # [n_freqs, n_mels + 2]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0.0)
    fb = fb * (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)
    return fb.float()


class LinearSpectrogram(nn.Module):
    """spectrogram.py:6-57.  forward(y [B, L] | [B, 1, L]) -> [B, n_fft/2 + 1, T] fp32 magnitudes."""

    def __init__(self, n_fft=2048, win_length=2048, hop_length=512, center=False, mode="pow2_sqrt"):
        super().__init__()
        if center:
            raise NotImplementedError("center=True is not used by the reference configs (spectrogram.py:10)")
        if mode != "pow2_sqrt":
            raise NotImplementedError("only mode='pow2_sqrt' (magnitude) has a kernel")
        if n_fft % hop_length != 0 or win_length > n_fft:
            raise NotImplementedError("the framed-DFT conv needs n_fft % hop_length == 0 and win_length <= n_fft")
        self.n_fft, self.win_length, self.hop_length, self.center, self.mode = n_fft, win_length, hop_length, center, mode
        self.register_buffer("window", torch.hann_window(win_length))
        self._ws = Workspace()
        self._packed = None
        self._packed_key = None

    @property
    def n_freqs(self) -> int:
        return self.n_fft // 2 + 1

    def _ensure_packed(self):
        key = params_key([self.window])
        if self._packed is not None and self._packed_key == key:
            return self._packed
        n_fft, hop, F = self.n_fft, self.hop_length, self.n_freqs
        dev = self.window.device
        with torch.no_grad():
            win = torch.zeros(n_fft, dtype=torch.float64, device=dev)
            left = (n_fft - self.win_length) // 2          # torch.stft centres a short window inside n_fft
            win[left:left + self.win_length] = self.window.double()
            n = torch.arange(n_fft, dtype=torch.float64, device=dev)
            f = torch.arange(F, dtype=torch.float64, device=dev)
            ang = 2.0 * math.pi * torch.outer(f, n) / n_fft  # [F, n_fft], exact in fp64 before the fp16 hi/lo split
            basis = torch.cat([torch.cos(ang) * win, -torch.sin(ang) * win], dim=0).float()  # [2F, n_fft]
            taps = [basis[:, j * hop:(j + 1) * hop].contiguous() for j in range(n_fft // hop)]
            self._packed = cabi.pack_taps(taps, list(range(n_fft // hop)))
        self._packed_key = key
        return self._packed

    def _spectrum_cl(self, y: torch.Tensor):
        """-> (re|im fp32 [B, T, >=2F], T); runs inside cabi.precision('strict')."""
        if y.ndim == 3:
            y = y.squeeze(1)
        require_cuda(y, type(self).__name__)
        y = y.contiguous().float()
        B, L = y.shape
        n_fft, hop = self.n_fft, self.hop_length
        pad_l, pad_r = (self.win_length - hop) // 2, (self.win_length - hop + 1) // 2
        Lp = L + pad_l + pad_r
        if Lp < n_fft:
            raise ValueError(f"signal too short: {L} samples for n_fft={n_fft}")
        T = (Lp - n_fft) // hop + 1
        k = n_fft // hop
        pc = self._ensure_packed()
        rows = T + k - 1
        self._ws.enter(forward_signature(y))
        a16 = cabi.frame_audio(y, hop, pad_l, pad_r, rows)
        spec = self._ws.f32("spec", B, T, pc.c_out, y.device)
        cabi.conv1d(a16, pc, T, out32=spec)
        return spec, T

    def forward(self, y: torch.Tensor) -> torch.Tensor:
        require_cuda(y, type(self).__name__)
        with cabi.precision("strict"), torch.cuda.device(y.device):
            spec, T = self._spectrum_cl(y)
            B = spec.shape[0]
            mag = self._ws.f32("mag", B, T, self.n_freqs, spec.device)
            cabi.spec_mag(spec, self.n_freqs, 1e-6, out32=mag)
            return cabi.unpack_output(mag, self.n_freqs)


class _MelScale(nn.Module):
    """Holder of the `mel_scale.fb` buffer [n_freqs, n_mels] (torchaudio.transforms.MelScale's state_dict layout)."""

    def __init__(self, n_mels, sample_rate, f_min, f_max, n_stft):
        super().__init__()
        self.register_buffer("fb", slaney_mel_filterbank(n_stft, f_min, f_max, n_mels, sample_rate))


class LogMelSpectrogram(nn.Module):
    """spectrogram.py:60-104.  forward(x [B, L] | [B, 1, L]) -> log-mel [B, n_mels, T] fp32."""

    #: mel weights are scaled by 2^10 before the fp16 hi/lo split (their tails are ~1e-5, near fp16's subnormal grid) and
    #: the scale is undone in fp32 by the GEMM epilogue
    FB_SCALE = 1024.0

    def __init__(self, sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, n_mels=128, center=False,
                 f_min=0.0, f_max=None):
        super().__init__()
        self.sample_rate, self.n_fft, self.win_length, self.hop_length = sample_rate, n_fft, win_length, hop_length
        self.center, self.n_mels, self.f_min, self.f_max = center, n_mels, f_min, f_max or sample_rate // 2
        self.spectrogram = LinearSpectrogram(n_fft, win_length, hop_length, center)
        self.mel_scale = _MelScale(n_mels, sample_rate, self.f_min, self.f_max, n_fft // 2 + 1)
        self._ws = Workspace()
        self._packed = None
        self._packed_key = None

    def compress(self, x: torch.Tensor) -> torch.Tensor:
        return torch.log(torch.clamp(x, min=1e-5))

    def decompress(self, x: torch.Tensor) -> torch.Tensor:
        return torch.exp(x)

    def _ensure_packed(self):
        key = params_key([self.mel_scale.fb])
        if self._packed is None or self._packed_key != key:
            with torch.no_grad():
                self._packed = cabi.pack_linear(self.mel_scale.fb.t().contiguous() * self.FB_SCALE, None)
            self._packed_key = key
        return self._packed

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        require_cuda(x, type(self).__name__)
        with cabi.precision("strict"), torch.cuda.device(x.device):
            spec, T = self.spectrogram._spectrum_cl(x)
            B, F, dev = spec.shape[0], self.spectrogram.n_freqs, spec.device
            pc = self._ensure_packed()
            self._ws.enter(forward_signature(x))
            mag16 = self._ws.f16("mag16", B, T, F, dev)
            cabi.spec_mag(spec, F, 1e-6, out16=mag16)
            mel = self._ws.f32("mel", B, T, self.n_mels, dev)
            cabi.conv1d(mag16, pc, out32=mel, out_scale=1.0 / self.FB_SCALE)
            return cabi.log_mel_out(mel, self.n_mels, 1e-5)
