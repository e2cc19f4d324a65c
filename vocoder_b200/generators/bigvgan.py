"""BigVGAN generator - drop-in mirror of ``fish_vocoder.modules.generators.bigvgan`` (reference file
fish_vocoder/modules/generators/bigvgan.py): constructor kwargs of bigvgan.py:256-270, state_dict keys
``resblocks.N.convs{1,2}.N`` (flat index stage*num_kernels+kernel), ``resblocks.N.activations.N.{act.alpha,
act.beta, upsample.filter, downsample.lowpass.filter}``, ``activation_post.*``.  The anti-aliased Snake
(alias_free_torch.Activation1d, bigvgan.py:9,226-233) is one fused CUDA kernel (fv_snake_aa).
"""
from __future__ import annotations

import math
from typing import Callable

import torch
from torch import nn

from .. import cabi
from ._mrf import MRFGeneratorBase, init_normal, strip_weight_norm, wn_conv


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """Kaiser-windowed sinc taps of alias-free-torch 0.0.6, shape [1, 1, kernel_size]."""
    half = kernel_size // 2
    atten = 2.285 * (half - 1) * math.pi * 4.0 * half_width + 7.95
    beta = 0.1102 * (atten - 8.7) if atten > 50.0 else (
        0.5842 * (atten - 21.0) ** 0.4 + 0.07886 * (atten - 21.0) if atten >= 21.0 else 0.0)
    window = torch.kaiser_window(kernel_size, periodic=False, beta=beta, dtype=torch.float32)
    t = torch.arange(-half, half, dtype=torch.float32) + 0.5 if kernel_size % 2 == 0 else \
        torch.arange(kernel_size, dtype=torch.float32) - half
    taps = 2.0 * cutoff * window * torch.sinc(2.0 * cutoff * t)
    return (taps / taps.sum()).view(1, 1, kernel_size)


class _ParamOnly(nn.Module):
    """Module with parameters but no CPU math: the arithmetic lives in libfv_b200.so."""

    def forward(self, *a, **k):
        raise cabi.FvError(f"{type(self).__name__} is evaluated by the fused CUDA kernels of its generator")


class Snake(_ParamOnly):
    """x + sin^2(alpha x)/alpha parameters (bigvgan.py:18-71)."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        base = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(base * alpha, requires_grad=alpha_trainable)


class SnakeBeta(_ParamOnly):
    """x + sin^2(alpha x)/beta parameters (bigvgan.py:74-135)."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        base = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(base * alpha, requires_grad=alpha_trainable)
        self.beta = nn.Parameter(base.clone() * alpha, requires_grad=alpha_trainable)


class _Filter(_ParamOnly):
    def __init__(self, taps):
        super().__init__()
        self.register_buffer("filter", taps)


class _Down(_ParamOnly):
    def __init__(self, taps):
        super().__init__()
        self.lowpass = _Filter(taps)


class Activation1d(_ParamOnly):
    """State-dict compatible holder for alias_free_torch.Activation1d (up 2x / act / down 2x, 12 taps):
    buffers ``upsample.filter`` and ``downsample.lowpass.filter`` [1,1,12], submodule ``act``."""

    #: how both anti-alias filters pad their input: "replicate" (the BigVGAN-flavoured alias_free_torch the state_dict
    #: layout matches), "reflect" or "zero" (the PyPI 0.0.6 wheel could not be inspected offline)
    edge_mode = "replicate"

    def __init__(self, activation, up_ratio=2, down_ratio=2, up_kernel_size=12, down_kernel_size=12):
        super().__init__()
        if (up_ratio, down_ratio, up_kernel_size, down_kernel_size) != (2, 2, 12, 12):
            raise NotImplementedError("the fused kernel implements the 2x / 12-tap configuration used by BigVGAN")
        self.act = activation
        self.upsample = _Filter(kaiser_sinc_filter1d(0.5 / up_ratio, 0.6 / up_ratio, up_kernel_size))
        self.downsample = _Down(kaiser_sinc_filter1d(0.5 / down_ratio, 0.6 / down_ratio, down_kernel_size))


class AMPBlock(nn.Module):
    """Parameter holder for bigvgan.py:138-233: 6 convs + 6 anti-aliased SnakeBeta activations."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5), activation=SnakeBeta, snake_logscale=True):
        super().__init__()
        self.convs1 = nn.ModuleList([wn_conv(channels, channels, kernel_size, d) for d in dilation])
        self.convs2 = nn.ModuleList([wn_conv(channels, channels, kernel_size, 1) for _ in dilation])
        init_normal(self)
        self.num_layers = len(self.convs1) + len(self.convs2)
        self.activations = nn.ModuleList(
            [Activation1d(activation=activation(channels, alpha_logscale=snake_logscale))
             for _ in range(self.num_layers)])

    def remove_parametrizations(self):
        strip_weight_norm(self)


class BigVGANGenerator(MRFGeneratorBase):
    snake_blocks = True
    #: overrides Activation1d.edge_mode of every anti-aliased activation when set ("replicate" | "reflect" | "zero")
    aa_edge_mode = None

    def __init__(
        self,
        *,
        hop_length: int = 512,
        upsample_rates=(8, 8, 2, 2, 2),
        upsample_kernel_sizes=(16, 16, 8, 2, 2),
        resblock_kernel_sizes=(3, 7, 11),
        resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
        num_mels: int = 128,
        upsample_initial_channel: int = 512,
        activation: Callable = SnakeBeta,
        use_template: bool = True,
        pre_conv_kernel_size: int = 7,
        post_conv_kernel_size: int = 7,
    ):
        super().__init__()
        self._build_trunk(hop_length=hop_length, upsample_rates=upsample_rates,
                          upsample_kernel_sizes=upsample_kernel_sizes, num_mels=num_mels,
                          upsample_initial_channel=upsample_initial_channel, use_template=use_template,
                          pre_conv_kernel_size=pre_conv_kernel_size, post_conv_kernel_size=post_conv_kernel_size)
        self.num_kernels = len(resblock_kernel_sizes)
        self.resblocks = nn.ModuleList()
        for ch in self.stage_channels:
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(AMPBlock(ch, k, tuple(d)))
        self.activation_post = Activation1d(activation=activation(self.stage_channels[-1], alpha_logscale=True))
        self._finish_trunk()
        self._filt_cache = {}

    def _block_modules(self, stage):
        nk = self.num_kernels
        return [self.resblocks[stage * nk + j] for j in range(nk)]

    def _pre_act(self):  # no activation before ups (bigvgan.py:355-356)
        return cabi.ACT_NONE, 0.0

    def _stage_out_act(self, last_stage):
        if last_stage:
            return None, 0.0  # activation_post is the anti-aliased Snake: separate fused kernel
        return cabi.ACT_NONE, 0.0

    def _taps(self, a1d: Activation1d):
        key = id(a1d)
        v = self._filt_cache.get(key)
        ver = (a1d.upsample.filter._version, a1d.downsample.lowpass.filter._version,
               a1d.upsample.filter.data_ptr())
        if v is None or v[0] != ver:
            up = a1d.upsample.filter.detach().reshape(-1).float().cpu().tolist()
            dn = a1d.downsample.lowpass.filter.detach().reshape(-1).float().cpu().tolist()
            v = (ver, up, dn)
            self._filt_cache[key] = v
        return v[1], v[2]

    def _snake(self, a1d: Activation1d, x32, out16, C, split=None):
        act = a1d.act
        up, dn = self._taps(a1d)
        beta = act.beta.detach() if isinstance(act, SnakeBeta) else None
        cabi.snake_aa(x32, out16, act.alpha.detach(), beta, up, dn, C, logscale=bool(act.alpha_logscale), split=split,
                      edge_mode=self.aa_edge_mode or a1d.edge_mode)

    def _final_activation(self, acc, h16, C, split):
        self._snake(self.activation_post, acc, h16, C, split)
