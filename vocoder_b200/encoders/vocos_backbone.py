"""VocosBackbone - the backbone of the upstream ``vocos==0.0.2`` model that the reference's ``scripts/vocos_gen.py:5-16``
runs (``vocos.Vocos.from_pretrained("charactr/vocos-mel-24khz")``: input_channels 100, dim 512, intermediate_dim 1536,
8 layers).  Third-party code, not part of /root/reference: restated from its published source,
parity pinned only against the CPU restatement in ``oracle/generators.py::vocos_backbone_forward``.

State-dict layout of upstream: ``embed.{weight [dim, C_in, 7], bias}``, ``norm.{weight,bias}``,
``convnext.N.{dwconv.*, norm.*, pwconv1.*, pwconv2.*, gamma}``, ``final_layer_norm.{weight,bias}``.

Same kernels as the ConvNeXt encoder, different glue:
    fv_conv1d (k=7)          embed                       (a strict layer in "mixed" precision, like the ConvNeXt stem)
    fv_dwconv_layernorm(k=0) norm
    per block                fv_dwconv_layernorm -> fv_conv1d + GELU -> fv_conv1d * gamma + residual
    fv_dwconv_layernorm(k=0) final_layer_norm -> fp16 operand of the head
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import cabi
from ..runtime import (GraphedForward, Workspace, forward_signature, module_params_key, params_key, require_channels, require_cuda,
                       with_precision)
from .convnext import ConvNeXtBlock


class VocosBackbone(nn.Module):
    def __init__(self, input_channels: int, dim: int, intermediate_dim: int, num_layers: int,
                 layer_scale_init_value: Optional[float] = None, adanorm_num_embeddings: Optional[int] = None):
        super().__init__()
        if adanorm_num_embeddings is not None:
            raise NotImplementedError("AdaLayerNorm conditioning (multi-bandwidth EnCodec Vocos) has no CUDA path")
        self.input_channels, self.dim, self.intermediate_dim = input_channels, dim, intermediate_dim
        self.embed = nn.Conv1d(input_channels, dim, kernel_size=7, padding=3)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        ls = layer_scale_init_value or 1.0 / num_layers
        self.convnext = nn.ModuleList([
            ConvNeXtBlock(dim=dim, layer_scale_init_value=ls, mlp_ratio=intermediate_dim / dim, kernel_size=7)
            for _ in range(num_layers)])
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-6)
        for m in self.modules():  # upstream _init_weights
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.constant_(m.bias, 0)
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
        return cabi.strict_layer(cabi.is_mixed())

    def _ensure_packed(self, device):
        key = module_params_key(self, buffers=False)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous()
        with torch.no_grad():
            with self._sens_ctx():
                embed = cabi.pack_conv(self.embed.weight, self.embed.bias)
            blocks = []
            for blk in self.convnext:
                C = blk.dwconv.weight.shape[0]
                blocks.append(dict(
                    k=blk.dwconv.kernel_size[0], dw_w=f32(blk.dwconv.weight).reshape(C, -1).t().contiguous(),
                    dw_b=f32(blk.dwconv.bias), ln_w=f32(blk.norm.weight), ln_b=f32(blk.norm.bias), eps=blk.norm.eps,
                    pw1=cabi.pack_linear(blk.pwconv1.weight, blk.pwconv1.bias),
                    pw2=cabi.pack_linear(blk.pwconv2.weight, blk.pwconv2.bias),
                    gamma=None if blk.gamma is None else f32(blk.gamma)))
            P = dict(embed=embed, blocks=blocks, norm=(f32(self.norm.weight), f32(self.norm.bias), self.norm.eps),
                     final=(f32(self.final_layer_norm.weight), f32(self.final_layer_norm.bias),
                            self.final_layer_norm.eps))
        self._packed, self._packed_key = P, key
        self._pack_gen += 1
        self._drop_graphs()
        return P

    def _pack_input(self, x: torch.Tensor) -> torch.Tensor:
        with self._sens_ctx():
            return cabi.pack_input(x)

    def _encode_cl(self, a0: torch.Tensor, want32: bool = False, out_strict: bool = False):
        P = self._ensure_packed(a0.device)
        ws, dev, eng = self._ws, a0.device, self.engine
        ws.enter(forward_signature(a0) + (want32, out_strict))
        B, T, _ = a0.shape
        mixed, C, Ci = cabi.is_mixed(), self.dim, self.intermediate_dim

        def pit(C_):
            with cabi.strict_layer(mixed):
                return cabi.pitch_of(C_)

        f32 = lambda name, C_: ws.get(name, (B, T, pit(C_)), torch.float32, dev)
        f16 = lambda name, C_, hilo=False: ws.get(
            name, (B, T, pit(C_) * (2 if (hilo or cabi.is_strict()) else 1)), torch.float16, dev)
        y, x = f32("embed", C), f32("x", C)
        with self._sens_ctx():
            cabi.conv1d(a0, P["embed"], out32=y, engine=eng)
        cabi.dwconv_layernorm(y, C, None, None, *P["norm"], 0, out32=x)
        h16, z16 = f16("h", C), f16("z", Ci)
        for blk in P["blocks"]:
            cabi.dwconv_layernorm(x, C, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"], blk["eps"], blk["k"], out16=h16)
            cabi.conv1d(h16, blk["pw1"], out16=z16, act=cabi.ACT_GELU, engine=eng)
            cabi.conv1d(z16, blk["pw2"], gamma=blk["gamma"], residual=x, out32=x, engine=eng)
        out16 = f16("out16", C, hilo=out_strict)
        out32 = f32("out32", C) if want32 else None
        with cabi.strict_layer(out_strict):
            cabi.dwconv_layernorm(x, C, None, None, *P["final"], 0, out16=out16, out32=out32)
        return out16, out32

    def _forward_eager(self, x):
        _, out32 = self._encode_cl(self._pack_input(x), want32=True)
        return out32[..., :self.dim]

    @with_precision
    def forward(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """[B, input_channels, T] -> [B, T, dim] fp32 (upstream hands the head a channels-last feature map)."""
        require_cuda(x, "VocosBackbone")
        require_channels(x, self.input_channels, "VocosBackbone")
        return self._forward_eager(x.contiguous().float()).clone()
