"""ISTFT head - drop-in mirror of ``fish_vocoder.modules.generators.vocos.ISTFTHead`` (reference file
fish_vocoder/modules/generators/vocos.py:6-69; state_dict: ``out.weight [2*n_fft, dim, 1]``, ``out.bias``,
``istft.window [win]``).

Also here: ``VocosBackbone``, the backbone of the upstream vocos==0.0.2 model that scripts/vocos_gen.py runs.

Launch sequence, features channels-last:
    fv_conv1d (k=1, POLAR epilogue): only the n_fft/2+1 live (log-mag, phase) row pairs of `out` are computed
                                     (irfft ignores the rest) -> S16 [B][T][2K] = (Re, Im) interleaved
    fv_conv1d (k=1)                : S16 x windowed inverse-real-DFT basis [n_fft, 2K]  -> frames32 [B][T][n_fft]
    fv_istft_ola                   : overlap-add, trim (win-hop)/2, divide by the hann^2 envelope -> wav [B][T*hop]
No cuFFT: the inverse FFT of vocos.spectral_ops.ISTFT is a dense tensor-core contraction here.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import cabi
from ..runtime import (GraphedForward, Workspace, forward_signature, module_params_key, params_key, require_channels, require_cuda,
                       with_precision)


class ISTFT(nn.Module):
    """Holder of the ``window`` buffer of vocos.spectral_ops.ISTFT (vocos==0.0.2)."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int, padding: str = "same"):
        super().__init__()
        if padding not in ("center", "same"):
            raise ValueError("Padding must be 'center' or 'same'.")
        self.padding = padding
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.register_buffer("window", torch.hann_window(win_length))


class ISTFTHead(nn.Module):
    """``upstream_layout=False``: the reference head (vocos.py:19-41): ``out`` = Conv1d(dim, 2*n_fft, 1).
    ``upstream_layout=True``: the upstream vocos==0.0.2 head that scripts/vocos_gen.py:5-16 runs: ``out`` =
    Linear(dim, n_fft + 2), win_length = n_fft (state_dict keys ``out.weight [n_fft+2, dim]``, ``out.bias``,
    ``istft.window``).  Both paddings of vocos.spectral_ops.ISTFT are implemented:
        "same"   wav [B, T*hop]      (every fish-vocoder yaml)
        "center" wav [B, (T-1)*hop]  (torch.istft(center=True); the reference's two-sided [B, n_fft, T] spectrum is cut to
                                      its first n_fft/2+1 rows by ATen before the c2r transform, so the same rows are live)
    """

    def __init__(self, dim: int, n_fft: int, hop_length: int, win_length: Optional[int] = None, padding: str = "same",
                 upstream_layout: bool = False):
        super().__init__()
        win_length = n_fft if win_length is None else win_length
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.upstream_layout = bool(upstream_layout)
        self.istft = ISTFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length, padding=padding)
        if self.upstream_layout:
            self.out = nn.Linear(dim, n_fft + 2)
        else:
            self.out = nn.Conv1d(dim, n_fft * 2, 1)  # vocos.py:40-41: out_dim = 2*n_fft (half of it is dead compute)
        self.dim = dim
        self._ws = Workspace()
        self._packed = None
        self._packed_key = None
        self._pack_gen = 0
        self._graphed: Optional[GraphedForward] = None
        self._ws.add_listener(self._drop_graphs)
        self.use_cuda_graph = False
        self.clone_graph_output = True
        self.engine = cabi.ENGINE_TC

    def _drop_graphs(self):
        if self._graphed is not None:
            self._graphed.invalidate()

    @staticmethod
    def _sens_ctx():
        """ "mixed" precision: both contractions of the head run strict.  exp() turns the absolute error of a log-magnitude
        into a relative error of a magnitude of up to 100, and the inverse-DFT operand holds those magnitudes: a single
        fp16 word there costs up to 5e-2 per bin."""
        return cabi.strict_layer(cabi.is_mixed())

    def _ensure_packed(self, device):
        key = module_params_key(self)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        center = self.istft.padding == "center"
        N = self.n_fft
        if not center and (self.win_length != N or (self.win_length - self.hop_length) % 2):
            raise NotImplementedError("ISTFTHead(padding='same') needs win_length == n_fft and an even (win - hop), as "
                                      "vocos.spectral_ops.ISTFT itself does")
        if self.win_length > N:
            raise ValueError("win_length must not exceed n_fft")
        with torch.no_grad():
            w = self.out.weight.detach().float().reshape(self.out.weight.shape[0], -1)  # [rows, dim]
            b = self.out.bias.detach().float()
            half = w.shape[0] // 2                      # rows per chunk: mag = rows[:half], phase = rows[half:]
            # irfft (padding="same") and torch.istft with a real output (padding="center") both consume bins 0..N/2 of
            # each chunk and ignore Im(DC), Im(Nyquist); the upstream layout has exactly those rows
            K = N // 2 + 1
            ck = torch.full((K,), 2.0, dtype=torch.float64, device=w.device)
            ck[0] = 1.0
            ck[-1] = 1.0
            assert K <= half
            # interleave live rows: 2k -> log-magnitude k, 2k+1 -> phase k
            w_live = torch.stack([w[:K], w[half:half + K]], dim=1).reshape(2 * K, -1)
            b_live = torch.stack([b[:K], b[half:half + K]], dim=1).reshape(2 * K)
            win = self.istft.window.detach().double()
            if self.win_length < N:  # torch.istft centres a short window inside n_fft
                left = (N - self.win_length) // 2
                win = torch.nn.functional.pad(win, (left, N - self.win_length - left))
            # windowed inverse DFT basis, scaled by N (the 1/N goes into the epilogue's out_scale so the fp16 basis
            # entries stay O(1)):  frame[n] = (1/N) sum_k c_k (Re_k cos(2 pi k n/N) - Im_k sin(...)).  With c_k = 2 the
            # sine column of DC / Nyquist is identically zero = "irfft ignores their imaginary parts".
            n = torch.arange(N, dtype=torch.float64, device=win.device)
            k = torch.arange(K, dtype=torch.float64, device=win.device)
            ang = 2.0 * math.pi * n[:, None] * k[None, :] / N
            re = ck[None, :] * torch.cos(ang) * win[:, None]
            im = -ck[None, :] * torch.sin(ang) * win[:, None]
            im[:, 0] = 0.0
            im[:, -1] = 0.0
            basis = torch.stack([re, im], dim=2).reshape(N, 2 * K).float()
            with self._sens_ctx():
                head = cabi.pack_linear(w_live, b_live)
                idft = cabi.pack_linear(basis, None)
            P = dict(head=head, idft=idft, window=win.float().contiguous(), K=K, center=center)
        self._packed, self._packed_key = P, key
        self._pack_gen += 1
        self._drop_graphs()
        return P

    def _forward_cl(self, h16: torch.Tensor) -> torch.Tensor:
        """h16 fp16 [B, T, pitch(dim)] ([hi | lo] in a strict layer) -> wav fp32 [B, T*hop] | [B, (T-1)*hop]."""
        P = self._ensure_packed(h16.device)
        ws, dev = self._ws, h16.device
        ws.enter(forward_signature(h16))
        B, T, _ = h16.shape
        with self._sens_ctx():
            S16 = ws.f16("S16", B, T, 2 * P["K"], dev)
            cabi.conv1d(h16, P["head"], out16=S16, act=cabi.ACT_POLAR, engine=self.engine)
            frames = ws.f32("frames", B, T, self.n_fft, dev)
            cabi.conv1d(S16, P["idft"], out32=frames, out_scale=1.0 / self.n_fft, engine=self.engine)
        return cabi.istft_ola(frames, P["window"], self.n_fft, self.hop_length, center=P["center"])

    def _forward_eager(self, x):
        with self._sens_ctx():
            a0 = cabi.pack_input(x)
        return self._forward_cl(a0)

    @with_precision
    def forward(self, x: torch.Tensor, template=None) -> torch.Tensor:
        """[B, dim, T] -> [B, T*hop]  (vocos.py:43-69).  ``template`` is accepted and ignored: the reference's
        UnifyGenerator passes it (unify.py:25) although the reference head cannot take it.
        With ``upstream_layout`` the input is [B, T, dim], as upstream vocos hands it over."""
        require_cuda(x, "ISTFTHead")
        if self.upstream_layout:
            x = x.transpose(1, 2)
        require_channels(x, self.dim, "ISTFTHead")
        x = x.contiguous().float()
        if self.use_cuda_graph and not torch.is_grad_enabled():
            self._ensure_packed(x.device)
            if self._graphed is None:
                self._graphed = GraphedForward(self._forward_eager)
            y = self._graphed(x, tag=self._pack_gen)
            return y.clone() if self.clone_graph_output else y
        return self._forward_eager(x)
