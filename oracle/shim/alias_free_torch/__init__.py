"""TEST INFRASTRUCTURE ONLY - stand-in for the third-party package ``alias-free-torch==0.0.6``.

The reference imports ``alias_free_torch.Activation1d`` (fish_vocoder/modules/generators/bigvgan.py:9,
pinned in pdm.lock:35-36). That package is not vendored in /root/reference and is not installed in
this image, so its published arithmetic is restated here purely so that the
*unmodified* reference BigVGAN module can be imported by ``oracle/make_golden.py``.

Parity status: UNPINNED against the real wheel (it cannot be fetched offline).  Cross-checked
bit-exactly against the independent copy shipped in this image's ``transformers``
(models/qwen2_5_omni/modeling_qwen2_5_omni.py: kaiser_sinc_filter1d / UpSample1d / DownSample1d /
TorchActivation1d) by tests/test_oracle_cpu.py.  Edge handling = ``replicate``.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """Kaiser-windowed sinc low-pass, normalised to unit DC gain; returns [1, 1, kernel_size]."""
    half = kernel_size // 2
    delta_f = 4.0 * half_width
    atten = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    if atten > 50.0:
        beta = 0.1102 * (atten - 8.7)
    elif atten >= 21.0:
        beta = 0.5842 * (atten - 21.0) ** 0.4 + 0.07886 * (atten - 21.0)
    else:
        beta = 0.0
    win = torch.kaiser_window(kernel_size, periodic=False, beta=beta, dtype=torch.float32)
    if kernel_size % 2 == 0:
        t = torch.arange(-half, half, dtype=torch.float32) + 0.5
    else:
        t = torch.arange(kernel_size, dtype=torch.float32) - half
    if cutoff == 0:
        return torch.zeros(1, 1, kernel_size)
    taps = 2.0 * cutoff * win * torch.sinc(2.0 * cutoff * t)
    taps = taps / taps.sum()
    return taps.view(1, 1, kernel_size)


class LowPassFilter1d(nn.Module):
    def __init__(self, cutoff=0.5, half_width=0.6, stride=1, padding=True,
                 padding_mode="replicate", kernel_size=12):
        super().__init__()
        self.kernel_size = kernel_size
        self.even = kernel_size % 2 == 0
        self.pad_left = kernel_size // 2 - int(self.even)
        self.pad_right = kernel_size // 2
        self.stride = stride
        self.padding = padding
        self.padding_mode = padding_mode
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff, half_width, kernel_size))

    def forward(self, x):
        c = x.shape[1]
        if self.padding:
            x = F.pad(x, (self.pad_left, self.pad_right), mode=self.padding_mode)
        return F.conv1d(x, self.filter.expand(c, -1, -1), stride=self.stride, groups=c)


class UpSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.stride = ratio
        self.pad = self.kernel_size // ratio - 1
        self.pad_left = self.pad * self.stride + (self.kernel_size - self.stride) // 2
        self.pad_right = self.pad * self.stride + (self.kernel_size - self.stride + 1) // 2
        self.register_buffer(
            "filter", kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, self.kernel_size))

    def forward(self, x):
        c = x.shape[1]
        x = F.pad(x, (self.pad, self.pad), mode="replicate")
        x = self.ratio * F.conv_transpose1d(x, self.filter.expand(c, -1, -1),
                                            stride=self.stride, groups=c)
        return x[..., self.pad_left:-self.pad_right]


class DownSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.lowpass = LowPassFilter1d(cutoff=0.5 / ratio, half_width=0.6 / ratio,
                                       stride=ratio, kernel_size=self.kernel_size)

    def forward(self, x):
        return self.lowpass(x)


class Activation1d(nn.Module):
    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2,
                 up_kernel_size: int = 12, down_kernel_size: int = 12):
        super().__init__()
        self.up_ratio = up_ratio
        self.down_ratio = down_ratio
        self.act = activation
        self.upsample = UpSample1d(up_ratio, up_kernel_size)
        self.downsample = DownSample1d(down_ratio, down_kernel_size)

    def forward(self, x):
        return self.downsample(self.act(self.upsample(x)))
