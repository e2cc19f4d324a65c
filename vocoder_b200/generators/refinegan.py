"""RefineGAN generator - drop-in mirror of ``fish_vocoder.modules.generators.refinegan`` (reference file
fish_vocoder/modules/generators/refinegan.py): constructor kwargs of refinegan.py:183-193, state_dict keys
``template_conv.*``, ``downsample_blocks.N.1.convs{1,2}.M.*``, ``mel_conv.*``, ``upsample_conv_blocks.N.input_conv.*``,
``upsample_conv_blocks.N.blocks.M.{0,2}.weight`` (AdaIN), ``upsample_conv_blocks.N.blocks.M.1.convs{1,2}.K.*``,
``output_conv.*``; ``forward(mel, template) -> [B, 1, T*hop]``.

U-Net launch sequence (channels-last, all arithmetic in libfv_b200.so):
    template -> conv k7 -> 4 x [leaky -> linear resample down -> ResBlock(C -> 2C)] -> cat(mel conv k7)
             -> 4 x [leaky -> linear resample up -> cat(skip) -> input conv k7 -> mean_k(AdaIN -> ResBlock -> AdaIN)]
             -> leaky -> conv k7 -> tanh
``torch.cat`` never materialises: producers write straight into channel slices of the fp16 operand buffer of the
consuming conv.  AdaIN's gaussian noise (drawn by the reference even in eval mode, refinegan.py:124-127) comes from
``self.noise_fn(B, L, C, pitch, device) -> fp32 [B, L, pitch]``; tests inject it for parity.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm

from .. import cabi
from ..runtime import (GraphedForward, Workspace, forward_signature, module_params_key, params_key, require_channels, require_cuda,
                       with_precision)
from ._mrf import same_padding, strip_weight_norm


class ResBlock(nn.Module):
    """Parameter holder for refinegan.py:37-110: three (conv, conv) pairs, BOTH convs of a pair dilated."""

    def __init__(self, *, in_channels, out_channels, kernel_size=7, dilation=(1, 3, 5), leaky_relu_slope=0.2):
        super().__init__()
        self.leaky_relu_slope = leaky_relu_slope
        self.in_channels, self.out_channels = in_channels, out_channels

        def conv(ci, d):
            c = nn.Conv1d(ci, out_channels, kernel_size, stride=1, dilation=d, padding=same_padding(kernel_size, d))
            c.weight.data.normal_(0, 0.01)
            c.bias.data.fill_(0.0)
            return weight_norm(c)

        self.convs1 = nn.ModuleList([conv(in_channels if i == 0 else out_channels, d) for i, d in enumerate(dilation)])
        self.convs2 = nn.ModuleList([conv(out_channels, d) for d in dilation])

    def remove_parametrizations(self):
        strip_weight_norm(self)


class AdaIN(nn.Module):
    """Per-channel noise gain of refinegan.py:113-127 (the noise itself is an input of the fused kernel)."""

    def __init__(self, *, channels, leaky_relu_slope=0.2):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.activation = nn.LeakyReLU(leaky_relu_slope)


class ParallelResBlock(nn.Module):
    """Parameter holder for refinegan.py:130-179."""

    def __init__(self, *, in_channels, out_channels, kernel_sizes=(3, 7, 11), dilation=(1, 3, 5), leaky_relu_slope=0.2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.input_conv = nn.Conv1d(in_channels, out_channels, kernel_size=7, stride=1, padding=3)
        self.blocks = nn.ModuleList([
            nn.Sequential(AdaIN(channels=out_channels),
                          ResBlock(in_channels=out_channels, out_channels=out_channels, kernel_size=k,
                                   dilation=dilation, leaky_relu_slope=leaky_relu_slope),
                          AdaIN(channels=out_channels))
            for k in kernel_sizes])

    def remove_parametrizations(self):
        for blk in self.blocks:
            blk[1].remove_parametrizations()


def _default_noise(B, L, C, pitch, device):
    return torch.randn(B, L, pitch, device=device, dtype=torch.float32)


class RefineGANGenerator(nn.Module):
    def __init__(
        self,
        *,
        sampling_rate: int = 44100,
        hop_length: int = 256,
        downsample_rates=(2, 2, 8, 8),
        upsample_rates=(8, 8, 2, 2),
        leaky_relu_slope: float = 0.2,
        num_mels: int = 128,
        start_channels: int = 16,
    ):
        super().__init__()
        self.sampling_rate, self.hop_length = sampling_rate, hop_length
        self.downsample_rates = tuple(int(r) for r in downsample_rates)
        self.upsample_rates = tuple(int(r) for r in upsample_rates)
        self.leaky_relu_slope = leaky_relu_slope
        assert np.prod(downsample_rates) == np.prod(upsample_rates) == hop_length
        self.template_conv = weight_norm(nn.Conv1d(1, start_channels, kernel_size=7, stride=1, padding=3))
        ch = start_channels
        self.downsample_blocks = nn.ModuleList()
        for rate in self.downsample_rates:
            self.downsample_blocks.append(nn.Sequential(
                nn.Upsample(scale_factor=1 / rate, mode="linear"),
                ResBlock(in_channels=ch, out_channels=ch * 2, kernel_size=7, dilation=(1, 3, 5),
                         leaky_relu_slope=leaky_relu_slope)))
            ch *= 2
        self.mel_conv = weight_norm(nn.Conv1d(num_mels, ch, kernel_size=7, stride=1, padding=3))
        ch *= 2
        self.upsample_blocks = nn.ModuleList()
        self.upsample_conv_blocks = nn.ModuleList()
        for rate in self.upsample_rates:
            self.upsample_blocks.append(nn.Upsample(scale_factor=rate, mode="linear"))
            self.upsample_conv_blocks.append(ParallelResBlock(
                in_channels=ch + ch // 4, out_channels=ch // 2, kernel_sizes=(3, 7, 11), dilation=(1, 3, 5),
                leaky_relu_slope=leaky_relu_slope))
            ch //= 2
        self.output_conv = weight_norm(nn.Conv1d(ch, 1, kernel_size=7, stride=1, padding=3))
        self.noise_fn: Callable = _default_noise
        self._ws = Workspace()
        self._packed = None
        self._packed_key = None
        self._graphed: Optional[GraphedForward] = None
        self.use_cuda_graph = False  # noise is drawn inside the forward: keep eager unless noise_fn is graph-safe
        self.engine = cabi.ENGINE_TC

    def remove_parametrizations(self) -> None:
        strip_weight_norm(self)
        self._packed = None

    # ---- packing --------------------------------------------------------------------------------------
    @staticmethod
    def _pack_resblock(rb: ResBlock):
        c1 = [cabi.pack_conv(c.weight, c.bias, c.dilation[0]) for c in rb.convs1]
        c2 = [cabi.pack_conv(c.weight, c.bias, c.dilation[0]) for c in rb.convs2]
        return c1, c2

    def _ensure_packed(self, device):
        key = module_params_key(self, buffers=False)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous()
        with torch.no_grad():
            P = {"tpl": cabi.pack_conv(self.template_conv.weight, self.template_conv.bias),
                 "mel": cabi.pack_conv(self.mel_conv.weight, self.mel_conv.bias),
                 "down": [self._pack_resblock(blk[1]) for blk in self.downsample_blocks], "up": []}
            for cb in self.upsample_conv_blocks:
                blocks = [(f32(b[0].weight), self._pack_resblock(b[1]), f32(b[2].weight)) for b in cb.blocks]
                P["up"].append((cabi.pack_conv(cb.input_conv.weight, cb.input_conv.bias), blocks))
            P["post_w"] = f32(self.output_conv.weight)[0].t().contiguous()
            P["post_b"] = f32(self.output_conv.bias)
        self._packed, self._packed_key = P, key
        return P

    # ---- launch sequence ----------------------------------------------------------------------------------
    def _resblock(self, tag, x16_in, res32, c1s, c2s, B, L, C, dev):
        """refinegan.py:87-99.  x16_in = leaky(x) operand; res32 = x (fp32) or None when C_in != C_out (first
        pair then has no residual).  Returns the fp32 block output (workspace buffer)."""
        ws, eng, sl = self._ws, self.engine, self.leaky_relu_slope
        xr = ws.f32(f"{tag}_xr", B, L, C, dev)
        xa = ws.f16(f"{tag}_xa", B, L, C, dev)
        ta = ws.f16(f"{tag}_ta", B, L, C, dev)
        a_in, res = x16_in, res32
        for c1, c2 in zip(c1s, c2s):
            cabi.conv1d(a_in, c1, out16=ta, act=cabi.ACT_LEAKY, act_param=sl, engine=eng)
            cabi.conv1d(ta, c2, residual=res, out32=xr, out16=xa, act=cabi.ACT_LEAKY, act_param=sl, engine=eng)
            a_in, res = xa, xr
        return xr

    def _forward_eager(self, mel, tpl):
        P = self._ensure_packed(mel.device)
        ws, dev, eng, sl = self._ws, mel.device, self.engine, self.leaky_relu_slope
        ws.enter(forward_signature(mel, tpl))
        LK = cabi.ACT_LEAKY
        B, _, T = mel.shape
        L = tpl.shape[-1]
        a_t = cabi.pack_input(tpl.reshape(B, 1, L))
        C = P["tpl"].c_out
        x = ws.f32("x_tpl", B, L, C, dev)
        cabi.conv1d(a_t, P["tpl"], out32=x, engine=eng)
        # channel plan of the up path: cat buffer i holds [upsampled x (Cx) | skip (Cd)]
        n_down = len(self.downsample_rates)
        down_C = [C * (2 ** i) for i in range(n_down)]           # channels of the skip taken before down block i
        down_L = []
        Lc = L
        for r in self.downsample_rates:
            down_L.append(Lc)
            Lc = Lc // r
        cats, cat_x = [], []
        Cx = down_C[-1] * 2 * 2                                   # channels entering the first up stage
        for i in range(len(self.upsample_rates)):
            skip_i = n_down - 1 - i
            # zero-initialised once: the channel padding of the operand buffer is never written afterwards
            cats.append(ws.get(f"cat_{i}", (B, down_L[skip_i], cabi.f16_width(Cx + down_C[skip_i])), torch.float16, dev,
                               zero=True))
            cat_x.append(Cx)
            Cx //= 2
        # ---- down path
        Lc = L
        for i, r in enumerate(self.downsample_rates):
            skip_up = n_down - 1 - i                              # up stage that consumes this skip
            d32 = ws.f32(f"d32_{i}", B, Lc, C, dev)
            # x = leaky(x); downs.append(x): fp32 copy feeds the resampler, fp16 copy lands in the cat operand
            cabi.act_cast(x, C, LK, sl, out32=d32, out16=cats[skip_up], out16_coff=cat_x[skip_up])
            Ln = int(Lc * (1.0 / r))
            r16 = ws.f16(f"r16_{i}", B, Ln, C, dev)
            cabi.resample_linear(d32, C, Ln, float(r), act=LK, act_param=sl, out16=r16)
            c1s, c2s = P["down"][i]
            x = self._resblock(f"dn{i}", r16, None, c1s, c2s, B, Ln, 2 * C, dev)
            C, Lc = 2 * C, Ln
        # ---- bottleneck: cat([x, mel_conv(mel)])
        assert Lc == T, f"template length {L} does not reduce to the {T} mel frames"
        a0 = cabi.pack_input(mel)
        m32 = ws.f32("m32", B, T, P["mel"].c_out, dev)
        cabi.conv1d(a0, P["mel"], out32=m32, engine=eng)
        srcs = [(x, C), (m32, P["mel"].c_out)]
        # ---- up path
        for i, r in enumerate(self.upsample_rates):
            cat = cats[i]
            Lu = Lc * r
            off = 0
            for src, cs in srcs:   # x = leaky(x); x = upsample(x)  -> channel slice of the cat operand
                cabi.resample_linear(src, cs, Lu, 1.0 / r, pre_act=LK, pre_param=sl, out16=cat, out_coff=off)
                off += cs
            pc_in, blocks = P["up"][i]
            Co = pc_in.c_out
            y = ws.f32(f"y_{i}", B, Lu, Co, dev)
            cabi.conv1d(cat, pc_in, out32=y, engine=eng)
            acc = ws.f32(f"uacc_{i}", B, Lu, Co, dev)
            a32 = ws.f32(f"a32_{i}", B, Lu, Co, dev)
            a16 = ws.f16(f"a16_{i}", B, Lu, Co, dev)
            nk = len(blocks)
            pitch = cabi.pitch_of(Co)
            for j, (w0, (c1s, c2s), w2) in enumerate(blocks):
                n0 = self.noise_fn(B, Lu, Co, pitch, dev)
                cabi.act_cast(y, Co, LK, sl, noise=n0, noise_w=w0, out32=a32, out16=a16, act16=LK, act16_param=sl)
                xr = self._resblock(f"up{i}", a16, a32, c1s, c2s, B, Lu, Co, dev)
                n2 = self.noise_fn(B, Lu, Co, pitch, dev)
                cabi.act_cast(xr, Co, LK, sl, noise=n2, noise_w=w2, out32=acc, out_scale=1.0 / nk, accumulate=j > 0)
            srcs = [(acc, Co)]
            Lc = Lu
        x, C = srcs[0]
        h16 = ws.f16("h_post", B, Lc, C, dev)
        cabi.act_cast(x, C, LK, sl, out16=h16)
        return cabi.conv_post_tanh(h16, P["post_w"], P["post_b"], C, apply_tanh=True)

    @with_precision
    def forward(self, mel: torch.Tensor, template: torch.Tensor) -> torch.Tensor:
        require_cuda(mel, "RefineGANGenerator")
        return self._forward_eager(mel.contiguous().float(), template.contiguous().float())
