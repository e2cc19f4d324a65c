"""ConvNeXt backbone - drop-in mirror of ``fish_vocoder.modules.encoders.convnext`` (reference file
fish_vocoder/modules/encoders/convnext.py): constructor kwargs of convnext.py:147-155, state_dict keys
``downsample_layers.N.N.{weight,bias}``, ``stages.N.N.{gamma, dwconv.*, norm.*, pwconv1.*, pwconv2.*}``, ``norm.*``.

Launch sequence per ConvNeXtBlock (convnext.py:124-143), activations channels-last [B][T][C]:
    fv_dwconv_layernorm : depthwise k7 + LayerNorm(C, eps 1e-6)            x32 -> h16
    fv_conv1d (k=1)     : pwconv1 + bias + exact-erf GELU epilogue          h16 -> z16 [B][T][4C]
    fv_conv1d (k=1)     : pwconv2 + bias, * gamma, + residual epilogue      z16 -> x32 (in place)
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from .. import cabi
from ..runtime import (GraphedForward, Workspace, forward_signature, module_params_key, params_key, require_channels, require_cuda,
                       with_precision)


class LayerNorm(nn.Module):
    """weight/bias holder for convnext.py:48-74 (both data formats normalise over C; eps 1e-6)."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)


class ConvNeXtBlock(nn.Module):
    """Parameter holder for convnext.py:78-143."""

    def __init__(self, dim, drop_path=0.0, layer_scale_init_value=1e-6, mlp_ratio=4.0, kernel_size=7, dilation=1):
        super().__init__()
        # the reference's `dilation` only changes the padding, never the conv's dilation (convnext.py:100-108)
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=kernel_size, padding=int(dilation * (kernel_size - 1) / 2),
                                groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, int(mlp_ratio * dim))
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(int(mlp_ratio * dim), dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones(dim), requires_grad=True)
                      if layer_scale_init_value > 0 else None)
        self.drop_path = nn.Identity()  # stochastic depth is the identity in eval mode (convnext.py:20-21)
        self.drop_prob = drop_path


class ConvNeXtEncoder(nn.Module):
    def __init__(
        self,
        input_channels: int = 3,
        depths: List[int] = [3, 3, 9, 3],
        dims: List[int] = [96, 192, 384, 768],
        drop_path_rate: float = 0.0,
        layer_scale_init_value: float = 1e-6,
        kernel_size: int = 7,
        kernel_sizes: Optional[List[int]] = None,
    ):
        super().__init__()
        # vocos-huge.yaml:9 spells the kwarg `kernel_sizes: [7]` (a TypeError in the reference);
        # accept and validate both spellings.
        if kernel_sizes is not None:
            ks = list(kernel_sizes) if isinstance(kernel_sizes, (list, tuple)) else [kernel_sizes]
            if len(set(ks)) != 1:
                raise ValueError("kernel_sizes must name a single depthwise kernel size")
            kernel_size = int(ks[0])
        assert len(depths) == len(dims)
        self.kernel_size = kernel_size
        self.dims = list(dims)
        self.depths = list(depths)
        self.input_channels = input_channels
        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(
            nn.Conv1d(input_channels, dims[0], kernel_size=kernel_size, padding=kernel_size // 2),
            LayerNorm(dims[0], eps=1e-6, data_format="channels_first")))
        for i in range(len(depths) - 1):
            self.downsample_layers.append(nn.Sequential(
                LayerNorm(dims[i], eps=1e-6, data_format="channels_first"),
                nn.Conv1d(dims[i], dims[i + 1], kernel_size=1)))
        rates = [r.item() for r in torch.linspace(0, drop_path_rate, sum(depths))]
        self.stages = nn.ModuleList()
        cur = 0
        for i in range(len(depths)):
            self.stages.append(nn.Sequential(*[
                ConvNeXtBlock(dim=dims[i], drop_path=rates[cur + j], layer_scale_init_value=layer_scale_init_value,
                              kernel_size=kernel_size) for j in range(depths[i])]))
            cur += depths[i]
        self.norm = LayerNorm(dims[-1], eps=1e-6, data_format="channels_first")
        self.apply(self._init_weights)
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
        """ "mixed" precision: the stem conv and the three 1x1 downsample convs (~1% of the tensor work, but every later
        block sees their error) carry [hi | lo] operands."""
        return cabi.strict_layer(cabi.is_mixed())

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv1d, nn.Linear)):  # convnext.py:201-204
            nn.init.trunc_normal_(m.weight, std=0.02)
            nn.init.constant_(m.bias, 0)

    # ---- packing ----------------------------------------------------------------------------------
    def _ensure_packed(self, device):
        key = module_params_key(self, buffers=False)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous()
        with torch.no_grad():
            P = {"down": [], "stages": []}
            for i, layer in enumerate(self.downsample_layers):
                with self._sens_ctx():
                    if i == 0:
                        conv, ln = layer[0], layer[1]
                        P["down"].append((cabi.pack_conv(conv.weight, conv.bias), f32(ln.weight), f32(ln.bias), ln.eps))
                    else:
                        ln, conv = layer[0], layer[1]
                        P["down"].append((cabi.pack_linear(conv.weight[:, :, 0], conv.bias), f32(ln.weight),
                                          f32(ln.bias), ln.eps))
            for stage in self.stages:
                blocks = []
                for blk in stage:
                    C = blk.dwconv.weight.shape[0]
                    blocks.append(dict(
                        C=C, k=blk.dwconv.kernel_size[0], dw_w=f32(blk.dwconv.weight).reshape(C, -1).t().contiguous(),
                        dw_b=f32(blk.dwconv.bias), ln_w=f32(blk.norm.weight), ln_b=f32(blk.norm.bias),
                        eps=blk.norm.eps, pw1=cabi.pack_linear(blk.pwconv1.weight, blk.pwconv1.bias),
                        pw2=cabi.pack_linear(blk.pwconv2.weight, blk.pwconv2.bias),
                        gamma=None if blk.gamma is None else f32(blk.gamma)))
                P["stages"].append(blocks)
            P["norm"] = (f32(self.norm.weight), f32(self.norm.bias), self.norm.eps)
        self._packed, self._packed_key = P, key
        self._pack_gen += 1
        self._drop_graphs()
        return P

    # ---- launch sequence ----------------------------------------------------------------------------
    def _pack_input(self, x: torch.Tensor) -> torch.Tensor:
        with self._sens_ctx():
            return cabi.pack_input(x)

    def _encode_cl(self, a0: torch.Tensor, want32: bool = False, out_strict: bool = False):
        """a0 fp16 [B,T,pitch(input_channels)] (from _pack_input) -> final-LayerNorm output as fp16 operand (and fp32 if
        asked).  out_strict: the consumer of the operand is a strict layer of a "mixed" forward (ISTFT head)."""
        P = self._ensure_packed(a0.device)
        ws, dev, eng = self._ws, a0.device, self.engine
        ws.enter(forward_signature(a0) + (want32, out_strict))
        B, T, _ = a0.shape
        mixed = cabi.is_mixed()

        def pit(C_):  # channel pitch of every buffer of this forward: the strict-layer pitch in "mixed" mode, so the plain
            with cabi.strict_layer(mixed):  # and the [hi | lo] operand rows derived from one fp32 stream share it
                return cabi.pitch_of(C_)

        def f32(name, C_):
            return ws.get(name, (B, T, pit(C_)), torch.float32, dev)

        def f16(name, C_, hilo=False):
            return ws.get(name, (B, T, pit(C_) * (2 if (hilo or cabi.is_strict()) else 1)), torch.float16, dev)

        x = None
        for i, blocks in enumerate(P["stages"]):
            pc, ln_w, ln_b, eps = P["down"][i]
            C = pc.c_out
            xn = f32(f"x_{i}", C)
            if i == 0:   # stem: conv k + LayerNorm(channels_first)            convnext.py:160-169
                y = f32("stem", C)
                with self._sens_ctx():
                    cabi.conv1d(a0, pc, out32=y, engine=eng)
                cabi.dwconv_layernorm(y, C, None, None, ln_w, ln_b, eps, 0, out32=xn)
            else:        # LayerNorm(channels_first) + 1x1 conv                 convnext.py:172-177
                Cp = self.dims[i - 1]
                h = f16(f"dn_{i}", Cp, hilo=mixed)
                with self._sens_ctx():
                    cabi.dwconv_layernorm(x, Cp, None, None, ln_w, ln_b, eps, 0, out16=h)
                    cabi.conv1d(h, pc, out32=xn, engine=eng)
            x = xn
            h16 = f16(f"h_{i}", C)
            z16 = f16(f"z_{i}", blocks[0]["pw1"].c_out if blocks else C)
            for blk in blocks:
                cabi.dwconv_layernorm(x, C, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"], blk["eps"], blk["k"],
                                      out16=h16)
                cabi.conv1d(h16, blk["pw1"], out16=z16, act=cabi.ACT_GELU, engine=eng)
                cabi.conv1d(z16, blk["pw2"], gamma=blk["gamma"], residual=x, out32=x, engine=eng)
        ln_w, ln_b, eps = P["norm"]
        C = self.dims[-1]
        out16 = f16("enc_out16", C, hilo=out_strict)
        out32 = f32("enc_out32", C) if want32 else None
        with cabi.strict_layer(out_strict):
            cabi.dwconv_layernorm(x, C, None, None, ln_w, ln_b, eps, 0, out16=out16, out32=out32)
        return out16, out32

    def _forward_eager(self, x):
        a0 = self._pack_input(x)
        _, out32 = self._encode_cl(a0, want32=True)
        return cabi.unpack_output(out32, self.dims[-1])

    @with_precision
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B, input_channels, T] -> [B, dims[-1], T] fp32 (convnext.py:206-214)."""
        require_cuda(x, "ConvNeXtEncoder")
        require_channels(x, self.input_channels, "ConvNeXtEncoder")
        x = x.contiguous().float()
        if self.use_cuda_graph and not torch.is_grad_enabled():
            self._ensure_packed(x.device)
            if self._graphed is None:
                self._graphed = GraphedForward(self._forward_eager)
            y = self._graphed(x, tag=self._pack_gen)
            return y.clone() if self.clone_graph_output else y
        return self._forward_eager(x)
