"""HiFiGAN generator - drop-in mirror of ``fish_vocoder.modules.generators.hifigan`` (reference file
fish_vocoder/modules/generators/hifigan.py): same constructor kwargs (hifigan.py:137-151), same state_dict key
layout (``conv_pre``, ``ups.N``, ``resblocks.N.blocks.N.convs{1,2}.N``, ``conv_post`` with weight-norm
``parametrizations.weight.original{0,1}``), same ``forward(x, template=None)`` / ``remove_parametrizations()``.
The forward pass runs entirely in libfv_b200.so (see generators/_mrf.py for the launch sequence).
"""
from __future__ import annotations

from functools import partial
from typing import Callable

from torch import nn

from .. import cabi
from ._mrf import MRFGeneratorBase, act_of_module, init_normal, strip_weight_norm, wn_conv


class ResBlock1(nn.Module):
    """Parameter holder for hifigan.py:25-99: three (dilated conv, conv) pairs, weight-normed, N(0, .01) init."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([wn_conv(channels, channels, kernel_size, d) for d in dilation])
        self.convs2 = nn.ModuleList([wn_conv(channels, channels, kernel_size, 1) for _ in dilation])
        init_normal(self)

    def remove_parametrizations(self):
        strip_weight_norm(self)


class ParralelBlock(nn.Module):
    """Parameter holder for hifigan.py:117-133 (name kept as spelled in the reference: it is a state_dict key)."""

    def __init__(self, channels, kernel_sizes=(3, 7, 11), dilation_sizes=((1, 3, 5),) * 3):
        super().__init__()
        assert len(kernel_sizes) == len(dilation_sizes)
        self.blocks = nn.ModuleList([ResBlock1(channels, k, d) for k, d in zip(kernel_sizes, dilation_sizes)])


class HiFiGANGenerator(MRFGeneratorBase):
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
        use_template: bool = True,
        pre_conv_kernel_size: int = 7,
        post_conv_kernel_size: int = 7,
        post_activation: Callable = partial(nn.SiLU, inplace=True),
    ):
        super().__init__()
        self._build_trunk(hop_length=hop_length, upsample_rates=upsample_rates,
                          upsample_kernel_sizes=upsample_kernel_sizes, num_mels=num_mels,
                          upsample_initial_channel=upsample_initial_channel, use_template=use_template,
                          pre_conv_kernel_size=pre_conv_kernel_size, post_conv_kernel_size=post_conv_kernel_size)
        self.num_kernels = len(resblock_kernel_sizes)
        self.resblocks = nn.ModuleList(
            [ParralelBlock(ch, tuple(resblock_kernel_sizes), tuple(tuple(d) for d in resblock_dilation_sizes))
             for ch in self.stage_channels])
        self.activation_post = post_activation()
        self._finish_trunk()

    def _block_modules(self, stage):
        return list(self.resblocks[stage].blocks)

    def _pre_act(self):  # F.silu before ups[0] (hifigan.py:230)
        return self._silu(), 0.0

    def _stage_out_act(self, last_stage):
        if last_stage:  # activation_post (hifigan.py:245)
            act, param = act_of_module(self.activation_post)
            return (self._silu() if act == cabi.ACT_SILU else act), param
        return self._silu(), 0.0  # F.silu before the next ups (hifigan.py:230)
