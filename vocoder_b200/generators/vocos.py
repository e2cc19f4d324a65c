"""ISTFT head - drop-in mirror of ``fish_vocoder.modules.generators.vocos.ISTFTHead`` (reference file
fish_vocoder/modules/generators/vocos.py:6-69; state_dict: ``out.weight [2*n_fft, dim, 1]``, ``out.bias``,
``istft.window [win]``).

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
from ..runtime import GraphedForward, Workspace, params_key, require_cuda, with_precision


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
    def __init__(self, dim: int, n_fft: int, hop_length: int, win_length: int, padding: str = "same"):
        super().__init__()
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.istft = ISTFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length, padding=padding)
        self.out = nn.Conv1d(dim, n_fft * 2, 1)  # vocos.py:40-41: out_dim = 2*n_fft (half of it is dead compute)
        self.dim = dim
        self._ws = Workspace()
        self._packed = None
        self._packed_key = None
        self._graphed: Optional[GraphedForward] = None
        self.use_cuda_graph = False
        self.engine = cabi.ENGINE_TC

    def _ensure_packed(self, device):
        key = params_key(list(self.parameters()) + list(self.buffers()))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        if self.istft.padding != "same":
            raise NotImplementedError("ISTFTHead: only padding='same' (every fish-vocoder yaml) has a CUDA path")
        if self.win_length != self.n_fft or (self.win_length - self.hop_length) % 2:
            raise NotImplementedError("ISTFTHead: needs win_length == n_fft and an even (win - hop)")
        N, K = self.n_fft, self.n_fft // 2 + 1
        with torch.no_grad():
            w = self.out.weight.detach().float()[:, :, 0]  # [2N, dim]
            b = self.out.bias.detach().float()
            # interleave live rows: 2k -> log-magnitude k, 2k+1 -> phase k
            w_live = torch.stack([w[:K], w[N:N + K]], dim=1).reshape(2 * K, -1)
            b_live = torch.stack([b[:K], b[N:N + K]], dim=1).reshape(2 * K)
            head = cabi.pack_linear(w_live, b_live)
            # windowed inverse real DFT basis, scaled by N (the 1/N goes into the epilogue's out_scale so the
            # fp16 basis entries stay O(1)):  frame[n] = (1/N) sum_k c_k (Re_k cos(2 pi k n/N) - Im_k sin(...))
            win = self.istft.window.detach().double()
            n = torch.arange(N, dtype=torch.float64, device=win.device)
            k = torch.arange(K, dtype=torch.float64, device=win.device)
            ck = torch.full((K,), 2.0, dtype=torch.float64, device=win.device)
            ck[0] = 1.0
            ck[-1] = 1.0
            ang = 2.0 * math.pi * n[:, None] * k[None, :] / N
            re = ck[None, :] * torch.cos(ang) * win[:, None]
            im = -ck[None, :] * torch.sin(ang) * win[:, None]
            basis = torch.stack([re, im], dim=2).reshape(N, 2 * K).float()
            idft = cabi.pack_linear(basis, None)
            P = dict(head=head, idft=idft, window=self.istft.window.detach().float().contiguous(), K=K)
        self._packed, self._packed_key = P, key
        if self._graphed is not None:
            self._graphed.invalidate()
        return P

    def _forward_cl(self, h16: torch.Tensor) -> torch.Tensor:
        """h16 fp16 [B, T, pitch(dim)] -> wav fp32 [B, T*hop]."""
        P = self._ensure_packed(h16.device)
        ws, dev = self._ws, h16.device
        B, T, _ = h16.shape
        S16 = ws.f16("S16", B, T, 2 * P["K"], dev)
        cabi.conv1d(h16, P["head"], out16=S16, act=cabi.ACT_POLAR, engine=self.engine)
        frames = ws.f32("frames", B, T, self.n_fft, dev)
        cabi.conv1d(S16, P["idft"], out32=frames, out_scale=1.0 / self.n_fft, engine=self.engine)
        return cabi.istft_ola(frames, P["window"], self.n_fft, self.hop_length)

    def _forward_eager(self, x):
        return self._forward_cl(cabi.pack_input(x))

    @with_precision
    def forward(self, x: torch.Tensor, template=None) -> torch.Tensor:
        """[B, dim, T] -> [B, T*hop]  (vocos.py:43-69).  ``template`` is accepted and ignored: the reference's
        UnifyGenerator passes it (unify.py:25) although the reference head cannot take it."""
        require_cuda(x, "ISTFTHead")
        x = x.contiguous().float()
        if self.use_cuda_graph and not torch.is_grad_enabled():
            self._ensure_packed(x.device)
            if self._graphed is None:
                self._graphed = GraphedForward(self._forward_eager)
            return self._graphed(x).clone()
        return self._forward_eager(x)
