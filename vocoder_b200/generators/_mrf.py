"""Shared host-side driver for the two MRF ("multi-receptive-field") generators, HiFiGAN and BigVGAN.

conv_pre -> N x [ConvTranspose1d upsample -> mean of num_kernels residual blocks] -> activation -> conv_post -> tanh
(reference: fish_vocoder/modules/generators/hifigan.py:226-249 and bigvgan.py:352-371).  This file only sequences
kernel launches of libfv_b200.so over channels-last workspace buffers; all arithmetic happens in the kernels.

HBM layout of one stage (B utterances, L time steps, C channels; pitch = C rounded up to 8):
    x0   fp32 [B][L][pitch]  stage input (ups output)        - residual for the first pair of each block
    xr   fp32 [B][L][pitch]  running residual stream of the block being evaluated
    xa   fp16 [B][L][pitch]  act(x): tensor-core operand of convs1
    ta   fp16 [B][L][pitch]  act(convs1 out): operand of convs2   (BigVGAN: + t32 fp32 pre-activation)
    acc  fp32 [B][L][pitch]  sum_j block_j(x0) / num_kernels (MRF mean, accumulated by the last epilogue)
    h16  fp16 [B][L][pitch]  activated stage output: operand of the next ups / conv_post
"""
from __future__ import annotations

from math import prod
from typing import List, Optional

import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm
from torch.nn.utils.parametrize import remove_parametrizations as _torch_remove_parametrizations

from .. import cabi
import contextlib

from ..runtime import (GraphedForward, Workspace, forward_signature, module_params_key, params_key, require_channels, require_cuda,
                       with_precision)


def same_padding(kernel_size: int, dilation: int = 1) -> int:
    return (kernel_size * dilation - dilation) // 2


def wn_conv(c_in: int, c_out: int, k: int, dilation: int = 1) -> nn.Module:
    """weight-normed "same" Conv1d parameter holder (state_dict: bias, parametrizations.weight.original{0,1})."""
    conv = nn.Conv1d(c_in, c_out, k, 1, dilation=dilation, padding=same_padding(k, dilation))
    return weight_norm(conv)


def init_normal(module: nn.Module, std: float = 0.01) -> None:
    """N(0, std) on every conv weight below `module` (the reference's init_weights, hifigan.py:15-18)."""
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
            m.weight.data.normal_(0.0, std)


def strip_weight_norm(module: nn.Module) -> None:
    for m in module.modules():
        if hasattr(m, "parametrizations") and "weight" in getattr(m, "parametrizations", {}):
            _torch_remove_parametrizations(m, "weight")


_ACT_OF_MODULE = {nn.SiLU: (cabi.ACT_SILU, 0.0), nn.Identity: (cabi.ACT_NONE, 0.0), nn.Tanh: (cabi.ACT_TANH, 0.0),
                  nn.GELU: (cabi.ACT_GELU, 0.0)}


def act_of_module(m: nn.Module):
    if isinstance(m, nn.LeakyReLU):
        return cabi.ACT_LEAKY, float(m.negative_slope)
    for cls, v in _ACT_OF_MODULE.items():
        if type(m) is cls:
            return v
    raise NotImplementedError(f"post activation {type(m).__name__} has no fused kernel epilogue")


class MRFGeneratorBase(nn.Module):
    """Common constructor pieces + launch sequence.  Subclasses provide the residual blocks."""

    snake_blocks = False  # BigVGAN: anti-aliased Snake between convs instead of SiLU epilogues

    def _build_trunk(self, *, hop_length, upsample_rates, upsample_kernel_sizes, num_mels,
                     upsample_initial_channel, use_template, pre_conv_kernel_size, post_conv_kernel_size):
        assert prod(upsample_rates) == hop_length, f"hop_length must be {prod(upsample_rates)}"
        self.hop_length = hop_length
        self.upsample_rates = tuple(int(u) for u in upsample_rates)
        self.upsample_kernel_sizes = tuple(int(k) for k in upsample_kernel_sizes)
        self.num_mels = num_mels
        self.num_upsamples = len(upsample_rates)
        self.use_template = use_template
        ch0 = upsample_initial_channel
        self.conv_pre = wn_conv(num_mels, ch0, pre_conv_kernel_size)
        self.noise_convs = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.stage_channels: List[int] = []
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            c_in, c_out = ch0 // (2 ** i), ch0 // (2 ** (i + 1))
            self.stage_channels.append(c_out)
            self.ups.append(weight_norm(nn.ConvTranspose1d(c_in, c_out, k, u, padding=(k - u) // 2)))
            if use_template:
                if i + 1 < len(self.upsample_rates):
                    s = int(prod(self.upsample_rates[i + 1:]))
                    self.noise_convs.append(nn.Conv1d(1, c_out, kernel_size=s * 2, stride=s, padding=s // 2))
                else:
                    self.noise_convs.append(nn.Conv1d(1, c_out, kernel_size=1))
        self._post_k = post_conv_kernel_size
        self._ws = Workspace()
        self._packed = None
        self._packed_key = None
        self._pack_gen = 0  # bumped by every re-pack: part of the key of every CUDA graph that replays packed pointers
        self._side_streams, self._ones_cache = {}, {}
        self._graphed: Optional[GraphedForward] = None
        self._ws.add_listener(self._drop_graphs)
        self.use_cuda_graph = False
        self.clone_graph_output = True  # False: forward returns the graph's static output buffer (valid until the next call)
        self.engine = cabi.ENGINE_TC

    def _drop_graphs(self):
        if self._graphed is not None:
            self._graphed.invalidate()

    def _finish_trunk(self):
        self.conv_post = wn_conv(self.stage_channels[-1], 1, self._post_k)
        init_normal(self.ups)
        init_normal(self.conv_post)

    def saturation_report(self):
        """fp16 operands clamp at +-65504 instead of overflowing: how many elements of the last forward's operand buffers
        sit on the clamp (0 for any sane checkpoint).  Debugging aid (synchronises)."""
        from ..runtime import saturation_report
        return saturation_report(self)

    # ---- reference surface -------------------------------------------------------------------
    def remove_parametrizations(self):
        """Fold weight-norm into plain weights (hifigan.py:251-257).  The kernels always consume folded,
        pre-packed weights, so this only changes the state_dict layout, exactly as in the reference."""
        strip_weight_norm(self)
        self._packed = None

    @with_precision
    def forward(self, x: torch.Tensor, template: Optional[torch.Tensor] = None) -> torch.Tensor:
        require_cuda(x, type(self).__name__)
        require_channels(x, self.num_mels, type(self).__name__)
        if self.use_template and template is None:
            raise ValueError("use_template=True requires a template [B, 1, T*hop]")
        x = x.contiguous().float()
        tpl = None
        if self.use_template:
            tpl = template.contiguous().float()
            if tpl.shape[0] != x.shape[0] or tpl.shape[-1] != x.shape[-1] * self.hop_length:
                raise ValueError(f"template must be [B, 1, T*hop] = [{x.shape[0]}, 1, {x.shape[-1] * self.hop_length}], "
                                 f"got {tuple(tpl.shape)}")
        if self.use_cuda_graph and not torch.is_grad_enabled():
            self._ensure_packed(x.device)
            if self._graphed is None:
                self._graphed = GraphedForward(self._forward_eager)
            y = self._graphed(x, tpl, tag=self._pack_gen)
            return y.clone() if self.clone_graph_output else y
        return self._forward_eager(x, tpl)

    def _trunk_strict(self) -> bool:
        """ "mixed" precision: conv_pre / ups / conv_post of the Snake generators carry [hi | lo] operands.  Measured on the
        reference goldens (tests/diag/precision_report.py): the trunk alone is ~90% of the fp16-operand waveform error, at
        < 5% of the tensor work."""
        return cabi.is_mixed() and self.snake_blocks

    def _trunk_ctx(self):
        return cabi.strict_layer() if self._trunk_strict() else contextlib.nullcontext()

    def _forward_eager(self, x, tpl):
        with self._trunk_ctx():
            a0 = cabi.pack_input(x)
        return self._forward_cl(a0, tpl)

    # ---- weight packing ------------------------------------------------------------------------
    def _block_modules(self, stage: int):  # -> list of blocks, each with .convs1/.convs2 (+ .activations)
        raise NotImplementedError

    def _ensure_packed(self, device):
        key = module_params_key(self) + (
            self.fuse_mrf, self.fuse_mrf_pairs, tuple(self.mrf_pairwise_channels), self.engine, self.conv_row_pairs, self.row_pairs_max_taps, self.chain_streams, self.chain_streams_max_rows)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        with torch.no_grad():
            with self._trunk_ctx():
                P = {"pre": cabi.pack_conv(self.conv_pre.weight, self.conv_pre.bias), "ups": [], "blocks": [],
                     "noise": [], "fused": [], "fused_split": [], "pairs": []}
            for i, up in enumerate(self.ups):
                with self._trunk_ctx():
                    P["ups"].append(cabi.pack_conv_transpose(up.weight, up.bias, self.upsample_rates[i]))
                blocks = []
                mods = self._block_modules(i)
                pairs = [(list(blk.convs1), list(blk.convs2)) for blk in mods]
                fused, fused_pairs, fused_split = None, None, None
                can_fuse = self.fuse_mrf and not self.snake_blocks and self.engine == cabi.ENGINE_TC
                Ci = self.stage_channels[i]
                if (can_fuse and Ci in self.mrf_pairwise_channels and Ci != 128
                        and cabi.mrf_fusable(Ci, pairs, pairwise=True)):
                    # opt-in: a C = 64 stage pair by pair as well (two co-resident CTAs per SM overlap MMA and epilogue)
                    fused_pairs = [[cabi.pack_mrf(Ci, [([c1], [c2])]) for c1, c2 in zip(c1s, c2s)] for c1s, c2s in pairs]
                elif can_fuse and cabi.mrf_fusable(self.stage_channels[i], pairs):
                    fused = cabi.pack_mrf(self.stage_channels[i], pairs)   # whole stage = one fv_mrf_fused launch
                    if len(pairs) > 1:   # short sequences: the last (largest) kernel size as its own concurrent launch
                        fused_split = (cabi.pack_mrf(Ci, pairs[:-1]), cabi.pack_mrf(Ci, pairs[-1:]))
                else:
                    if can_fuse and self.fuse_mrf_pairs and cabi.mrf_fusable(Ci, pairs, pairwise=True):
                        # C = 128: one fv_mrf_fused launch per (conv, conv) pair (256-row tiles, pair halo <= 30 rows);
                        # "auto" keeps the layer-wise weights as well and chooses per forward by the number of rows
                        fused_pairs = [[cabi.pack_mrf(Ci, [([c1], [c2])]) for c1, c2 in zip(c1s, c2s)] for c1s, c2s in pairs]
                    if fused_pairs is None or self.fuse_mrf_pairs == "auto":
                        # Snake stages with C <= 16 (BigVGAN's last): the same convs on pairs of time steps, see
                        # cabi.pack_conv_row_pairs; chosen per forward (even length, plain fp16 operands)
                        pair_rows = (self.conv_row_pairs and self.snake_blocks and Ci <= 16 and not cabi.is_strict()
                                     and cabi.pitch_of(Ci) == Ci)
                        for blk in mods:
                            c1 = [cabi.pack_conv(c.weight, c.bias, c.dilation[0]) for c in blk.convs1]
                            c2 = [cabi.pack_conv(c.weight, c.bias, c.dilation[0]) for c in blk.convs2]
                            if pair_rows:
                                # measured on B200 (C = 16, L = 44544, B = 32; us per launch, plain -> pairs): k=3 48 -> 30,
                                # k=7 d=3,5 (11 pair taps) 48 -> 45, k=11 d=1 (7) 59 -> 35, k=11 d=3,5 (17) 58 -> 61: beyond
                                # `row_pairs_max_taps` pair taps the MMA issue time of the longer tap list outweighs the epilogue
                                def rp_pack(c):
                                    pc = cabi.pack_conv_row_pairs(c.weight, c.bias, c.dilation[0])
                                    return pc if pc.n_taps <= self.row_pairs_max_taps else None
                                blk_pairs = ([rp_pack(c) for c in blk.convs1], [rp_pack(c) for c in blk.convs2])
                            else:
                                blk_pairs = None
                            blocks.append((c1, c2, blk, blk_pairs))
                P["blocks"].append(blocks)
                P["fused"].append(fused)
                P["fused_split"].append(fused_split)
                P["pairs"].append(fused_pairs)
            for nc in self.noise_convs:
                P["noise"].append((nc.weight.detach().float().reshape(nc.out_channels, -1).contiguous(),
                                   nc.bias.detach().float().contiguous(), nc.kernel_size[0], nc.stride[0],
                                   nc.padding[0]))
            wp = self.conv_post.weight.detach().float()  # [1, C, k]
            P["post_w"] = wp[0].t().contiguous()          # [k, C]
            P["post_b"] = self.conv_post.bias.detach().float().contiguous()
        self._packed, self._packed_key = P, key
        self._pack_gen += 1
        self._drop_graphs()
        return P

    # ---- hooks for the activation flavour -------------------------------------------------------
    def _pre_act(self):        # activation fused into conv_pre's epilogue (what ups[0] consumes)
        raise NotImplementedError

    def _stage_out_act(self, last_stage: bool):   # activation fused into the stage's final epilogue
        raise NotImplementedError

    def _final_activation(self, acc, h16, C, split):     # un-fusable activation_post (BigVGAN AA-Snake) else no-op
        return None

    # ---- the launch sequence ----------------------------------------------------------------------
    def _forward_cl(self, a0: torch.Tensor, tpl: Optional[torch.Tensor]) -> torch.Tensor:
        """a0: fp16 [B, T, pitch(num_mels)] channels-last ([hi | lo] when the trunk is strict).  Returns wav fp32 [B, 1, T*hop]."""
        P = self._ensure_packed(a0.device)
        ws, dev, eng = self._ws, a0.device, self.engine
        ws.enter(forward_signature(a0, tpl))
        B, T, _ = a0.shape
        pre = P["pre"]
        trunk = self._trunk_strict()   # "mixed": h16 (what ups / conv_post consume) is [hi | lo] between stages

        def h_buf(name, L_, C_):       # operand of a trunk layer + the `split` its producer must write
            with self._trunk_ctx():
                t = ws.f16(name, B, L_, C_, dev)
            return t, ((t.shape[-1] // 2) if (trunk or cabi.is_strict()) else 0)

        act_pre, act_pre_p = self._pre_act()
        h16, h_split = h_buf("h_pre", T, pre.c_out)
        with self._trunk_ctx():
            cabi.conv1d(a0, pre, out16=h16, act=act_pre, act_param=act_pre_p, engine=eng)
        L = T
        n_stage = self.num_upsamples
        for i in range(n_stage):
            up = P["ups"][i]
            C = up.c_out
            Lo = cabi.conv_transpose_out_len(L, self.upsample_kernel_sizes[i], self.upsample_rates[i])
            x0 = ws.f32(f"x0_{i}", B, Lo, C, dev)
            nz = None
            if self.use_template:
                w, b, k, s, p = P["noise"][i]
                nz = ws.f32(f"nz_{i}", B, Lo, C, dev)
                cabi.noise_conv(tpl.reshape(B, -1), w, b, nz, C, k, s, p)
            last_stage = i == n_stage - 1
            fused = P["fused"][i]
            if fused is not None:
                # C <= 64 SiLU stages: ups writes only the fp32 stage input; the whole MRF (3 kernel sizes x 3 pairs)
                # runs on chip per time tile and leaves the mean (+ the activated fp16 operand of the next layer)
                cabi.conv1d(h16, up, Lo, residual=nz, out32=x0, engine=eng)
                acc = ws.f32(f"acc_{i}", B, Lo, C, dev)
                h_next = ws.f16(f"h_{i}", B, Lo, C, dev)
                out_act, out_act_p = self._stage_out_act(last_stage)
                inner = cabi.ACT_SILU_H2 if (self.mrf_silu_h2 and self.mrf_silu_tanh) else self._silu()
                split_packs = P["fused_split"][i]
                if split_packs is not None and self._chains_concurrent(B, Lo):
                    # few tiles (B = 1 ... 2): the largest kernel size (more taps than the others together) runs as its own
                    # launch on the side stream next to the launch of the other chains; the same three scaled terms are
                    # summed in the same order as in the single launch
                    nk = len(fused.ksize)
                    a_side = ws.f32(f"acc_side_{i}", B, Lo, C, dev)
                    side, cur = self._side_stream(dev), torch.cuda.current_stream()
                    side.wait_stream(cur)
                    with torch.cuda.stream(side):
                        cabi.mrf_fused(x0, split_packs[1], a_side, act=inner, out_scale=1.0 / nk)
                    cabi.mrf_fused(x0, split_packs[0], acc, act=inner, out_scale=1.0 / nk)
                    cur.wait_stream(side)
                    cabi.act_cast(acc, C, cabi.ACT_NONE, noise=a_side, noise_w=self._ones(C, dev), out32=acc,
                                  out16=h_next if out_act is not None else None, act16=out_act or cabi.ACT_NONE,
                                  act16_param=out_act_p)
                else:
                    cabi.mrf_fused(x0, fused, acc, out16=h_next if out_act is not None else None, act=inner,
                                   out_act=out_act or cabi.ACT_NONE, out_act_param=out_act_p)
                if last_stage:
                    self._final_activation(acc, h_next, C, 0)
                h16, L = h_next, Lo
                continue
            fpairs = P["pairs"][i]
            if fpairs is not None and self.fuse_mrf_pairs == "auto" and B * Lo > self.mrf_pairs_max_rows:
                fpairs = None   # enough rows to fill the layer-wise kernels: they are faster there (see fuse_mrf_pairs)
            if fpairs is not None:
                # C = 128 SiLU stage, pair by pair on chip: x -> x + conv2(silu(conv1(silu(x)))) per launch, fp32 in / out;
                # the last pair of every chain adds its chain's share of the MRF mean onto acc
                cabi.conv1d(h16, up, Lo, residual=nz, out32=x0, engine=eng)
                acc = ws.f32(f"acc_{i}", B, Lo, C, dev)
                h_next = ws.f16(f"h_{i}", B, Lo, C, dev)
                out_act, out_act_p = self._stage_out_act(last_stage)
                inner = cabi.ACT_SILU_H2 if (self.mrf_silu_h2 and self.mrf_silu_tanh) else self._silu()
                nk = len(fpairs)

                def pair_chain(j, tag, before_final=None):
                    xp = (ws.f32(f"xp0{tag}_{i}", B, Lo, C, dev), ws.f32(f"xp1{tag}_{i}", B, Lo, C, dev))
                    src, chain = x0, fpairs[j]
                    for p_i, pm in enumerate(chain):
                        if p_i + 1 < len(chain):
                            dst = xp[0] if src is not xp[0] else xp[1]
                            cabi.mrf_fused(src, pm, dst, act=inner, out_scale=1.0)
                            src = dst
                        else:
                            if before_final is not None:
                                before_final()
                            final = j == nk - 1
                            cabi.mrf_fused(src, pm, acc, act=inner, accumulate=j > 0, out_scale=1.0 / nk,
                                           out16=h_next if (final and out_act is not None) else None,
                                           out_act=(out_act or cabi.ACT_NONE) if final else cabi.ACT_NONE,
                                           out_act_param=out_act_p)

                self._run_chains(nk, pair_chain, B, Lo, dev)
                if last_stage:
                    self._final_activation(acc, h_next, C, 0)
                h16, L = h_next, Lo
                continue
            with self._trunk_ctx():
                if self.snake_blocks:
                    cabi.conv1d(h16, up, Lo, residual=nz, out32=x0, engine=eng)
                    xa0 = None
                else:
                    xa0 = ws.f16(f"xa0_{i}", B, Lo, C, dev)
                    cabi.conv1d(h16, up, Lo, residual=nz, out32=x0, out16=xa0, act=self._silu(), engine=eng)
            with self._trunk_ctx():   # same channel pitch as the [hi | lo] operand the final activation derives from it
                acc = ws.f32(f"acc_{i}", B, Lo, C, dev)
            h_next, h_split = h_buf(f"h_{i}", Lo, C)
            # The residual blocks can run per micro-batch of utterances (working set resident in the 126 MB L2).
            # Measured on B200 (HiFiGAN cfg B): 64 -> 8.6 ms, 32 -> 9.4 ms, 16 -> 10.6 ms, 8 -> 13.7 ms per forward:
            # tail and launch effects outweigh the L2 hits, so the default is the whole batch.
            mb = self._micro_batch(B, Lo, C)
            xr_w = ws.f32(f"xr_{i}", mb, Lo, C, dev)
            xa_w = ws.f16(f"xa_{i}", mb, Lo, C, dev)
            ta = ws.f16(f"ta_{i}", mb, Lo, C, dev)
            t32 = ws.f32(f"t32_{i}", mb, Lo, C, dev) if self.snake_blocks else None
            blocks = P["blocks"][i]
            nk = len(blocks)
            out_act, out_act_p = self._stage_out_act(last_stage)
            for b0 in range(0, B, mb):
                b1 = min(B, b0 + mb)
                n = b1 - b0
                x0_m, acc_m, h_m = x0[b0:b1], acc[b0:b1], h_next[b0:b1]
                xa0_m = None if xa0 is None else xa0[b0:b1]

                def layer_chain(j, tag, before_final=None):
                    if tag:   # a chain on the side stream works in its own buffers
                        xr_m = ws.f32(f"xr{tag}_{i}", mb, Lo, C, dev)[:n]
                        xa_m = ws.f16(f"xa{tag}_{i}", mb, Lo, C, dev)[:n]
                        ta_m = ws.f16(f"ta{tag}_{i}", mb, Lo, C, dev)[:n]
                        t32_m = ws.f32(f"t32{tag}_{i}", mb, Lo, C, dev)[:n] if self.snake_blocks else None
                    else:
                        xr_m, xa_m, ta_m = xr_w[:n], xa_w[:n], ta[:n]
                        t32_m = None if t32 is None else t32[:n]
                    c1s, c2s, blk, blk_pairs = blocks[j]
                    xr, xa = x0_m, xa0_m
                    n_pairs = len(c1s)
                    rp = (blk_pairs is not None and cabi.row_pairs_ok(C, Lo)
                          and all(t.shape[-1] == C for t in (x0_m, xr_m, xa_m, ta_m, t32_m, acc_m)))
                    for p_i in range(n_pairs):
                        last_pair = p_i == n_pairs - 1
                        if rp:   # snake stage on [n, Lo / 2, 2 C] views: full 32-column tiles for C = 16
                            self._snake(blk.activations[2 * p_i], xr, xa_m, C)
                            if blk_pairs[0][p_i] is not None:
                                cabi.conv1d_row_pairs(xa_m, blk_pairs[0][p_i], out32=t32_m, engine=eng)
                            else:
                                cabi.conv1d(xa_m, c1s[p_i], out32=t32_m, engine=eng)
                            self._snake(blk.activations[2 * p_i + 1], t32_m, ta_m, C)
                            if blk_pairs[1][p_i] is None:
                                pass   # plain view below
                            elif not last_pair:
                                cabi.conv1d_row_pairs(ta_m, blk_pairs[1][p_i], residual=xr, out32=xr_m, engine=eng)
                                xr = xr_m
                                continue
                            elif not ((j == nk - 1) and out_act is not None):
                                if before_final is not None:
                                    before_final()
                                cabi.conv1d_row_pairs(ta_m, blk_pairs[1][p_i], residual=xr, out32=acc_m, accumulate=j > 0,
                                                      out_scale=1.0 / nk, engine=eng)
                                continue
                            # the stage's very last conv also writes the next ups operand ([hi | lo] rows in "mixed"): plain view
                        elif self.snake_blocks:
                            self._snake(blk.activations[2 * p_i], xr, xa_m, C)
                            cabi.conv1d(xa_m, c1s[p_i], out32=t32_m, engine=eng)
                            self._snake(blk.activations[2 * p_i + 1], t32_m, ta_m, C)
                        else:
                            cabi.conv1d(xa, c1s[p_i], out16=ta_m, act=self._silu(), engine=eng)
                        if not last_pair:
                            cabi.conv1d(ta_m, c2s[p_i], residual=xr, out32=xr_m,
                                        out16=None if self.snake_blocks else xa_m, act=self._silu(), engine=eng)
                            xr, xa = xr_m, xa_m
                        else:
                            if before_final is not None:
                                before_final()
                            want16 = (j == nk - 1) and out_act is not None
                            cabi.conv1d(ta_m, c2s[p_i], residual=xr, out32=acc_m, accumulate=j > 0,
                                        out_scale=1.0 / nk, out16=h_m if want16 else None,
                                        act=out_act if want16 else cabi.ACT_NONE, act_param=out_act_p, engine=eng,
                                        out16_split=h_split if want16 else None)

                self._run_chains(nk, layer_chain, n, Lo, dev, allow=(mb == B))
            if last_stage:
                self._final_activation(acc, h_next, C, h_split)
            h16, L = h_next, Lo
        wav = cabi.conv_post_tanh(h16, P["post_w"], P["post_b"], self.stage_channels[-1], apply_tanh=True,
                                  split=h_split if (trunk or cabi.is_strict()) else 0)
        return wav

    def _snake(self, act_module, x32, out16, C, split=None):
        raise NotImplementedError

    # ---- short sequences: the kernel-size chains of a stage are independent until the MRF mean -------------------------
    #: Run the kernel-size chains of a stage on their own streams when the stage is too small to fill the GPU (test.py
    #: runs B = 1 ... 2: a 752-row C = 256 stage is 3 tiles on 148 SMs, and every launch is a serial chain of MMAs on
    #: one tile).  Layer-wise and pair-wise stages: one stream per chain; whole-stage fused stages: the largest kernel size
    #: (more taps than the others together) as its own launch.  The chains meet at their final conv (running-mean
    #: accumulate, ordered by events in chain order), i.e. results are bit-identical to the serial order.  "auto" = stages with at most
    #: `chain_streams_max_rows` rows (batch x length); True / False force it.  Works under CUDA-graph capture (fork / join).
    chain_streams = "auto"
    chain_streams_max_rows = 148 * 256

    def _chains_concurrent(self, B: int, L: int) -> bool:
        if self.chain_streams == "auto":
            return B * L <= self.chain_streams_max_rows
        return bool(self.chain_streams)

    def _side_stream(self, device, idx: int = 0) -> "torch.cuda.Stream":
        key = (str(device), idx)
        st = self._side_streams.get(key)
        if st is None:
            st = self._side_streams[key] = torch.cuda.Stream(device=device)
        return st

    def _ones(self, C: int, device) -> torch.Tensor:
        key = (C, str(device))
        t = self._ones_cache.get(key)
        if t is None:
            t = self._ones_cache[key] = torch.ones(C, dtype=torch.float32, device=device)
        return t

    def _run_chains(self, nk: int, chain_fn, B: int, L: int, device, allow: bool = True) -> None:
        """chain_fn(j, tag, before_final): issue chain j; `tag` names its private work buffers; `before_final` must be called
        right before the chain's final launch (the one that accumulates onto the stage's running mean)."""
        if not (allow and nk > 1 and self._chains_concurrent(B, L)):
            for j in range(nk):
                chain_fn(j, "")
            return
        # one stream per chain (chain 0 stays on the forward's stream): at these sizes a launch costs ~10 us whatever its
        # taps, so the k = 3 and k = 7 chains together (12 launches) would outlast the k = 11 chain (6 launches) on one stream
        cur = torch.cuda.current_stream()
        sides = [self._side_stream(device, j) for j in range(1, nk)]
        for st in sides:
            st.wait_stream(cur)                     # fork: the stage input is complete
        chain_fn(0, "")
        prev = torch.cuda.Event()
        prev.record(cur)                            # running mean holds chain 0
        for j in range(1, nk):
            st = sides[j - 1]
            with torch.cuda.stream(st):
                chain_fn(j, f"s{j}", before_final=lambda st=st, ev=prev: st.wait_event(ev))
                prev = torch.cuda.Event()
                prev.record(st)                     # ... and chains 0 .. j
        for st in sides:
            cur.wait_stream(st)                     # join

    #: C <= 64 SiLU stages as one on-chip kernel per stage (fv_mrf_fused); False = layer-wise fv_conv1d launches
    fuse_mrf = True
    #: C = 128 SiLU stages as one on-chip kernel per (conv, conv) pair (needs fuse_mrf); False = layer-wise launches.
    #: Measured on B200 (HiFiGAN cfg B, C = 128, L = 6016, B = 64): 2.42 ms pair-wise against 2.0 ms layer-wise - an N = 128
    #: UMMA costs ~97 cycles whatever feeds it, the 256-row tiles recompute 25% halo, and with TMEM full (X + T) one CTA per
    #: SM cannot overlap its epilogue / entry / exit with MMAs the way the layer-wise kernel's double-buffered
    #: accumulators do.  At B = 1 (6016 rows: 32 tiles on 148 SMs either way) the pair-wise stage saves nine launches and
    #: measured 0.776 against 0.832 ms per forward.  "auto" (default) = pair-wise when the stage has at most
    #: `mrf_pairs_max_rows` rows (batch x length), layer-wise above; True / False force one path.
    fuse_mrf_pairs = "auto"
    #: C <= 16 Snake stages: convs on the [B, L/2, 2C] view of the channels-last buffers (cabi.pack_conv_row_pairs);
    #: False = one 16-channel row per GEMM row (half-empty 32-column tiles)
    conv_row_pairs = True
    row_pairs_max_taps = 12
    mrf_pairs_max_rows = 148 * 256
    #: channel counts (besides 128) whose stage runs pair by pair instead of as one whole-stage launch; (64,) trades ~7x the
    #: stage's HBM traffic for MMA / epilogue overlap between two co-resident CTAs.  Measured: see DESIGN.md section 4.2.
    mrf_pairwise_channels = ()
    def _silu(self) -> int:
        return cabi.ACT_SILU_TANH if self.mrf_silu_tanh else cabi.ACT_SILU

    #: SiLU epilogues (fused stages and layer-wise convs) as x/2 + x/2 tanh(x/2) with tanh.approx (one SFU op instead of two; |error| <=
    #: 2.4e-4 |x| before the fp16 rounding of the operand, see FV_ACT_SILU_TANH).  Measured on B200: stage-level error
    #: 5.8e-5 vs 4.3e-5 (ex2 + rcp) against the fp64 contract, waveform error of the full-width stress model unchanged
    #: (8.1e-5 both); False selects the ex2 + rcp form.
    mrf_silu_tanh = True
    #: inner SiLU of the fused stages on packed fp16 pairs (FV_ACT_SILU_H2: one SFU op per two channels, ~3 fp16 roundings in
    #: the operand instead of 1).  Opt-in until measured: see DESIGN.md section 4.2.
    mrf_silu_h2 = False

    #: utterances per residual-block pass; None = whole batch; 0 = size the block working set for L2 (_micro_batch)
    micro_batch = None
    l2_budget_bytes = 96 * 1024 * 1024

    def _micro_batch(self, B: int, L: int, C: int) -> int:
        if self.micro_batch is None:
            return B
        if self.micro_batch > 0:
            return max(1, min(B, int(self.micro_batch)))
        # live set of one (c1, c2) pair per utterance: xr fp32 (read + written in place), xa + ta fp16 (+ t32 fp32)
        per_utt = L * cabi.pitch_of(C) * (4 + 2 + 2 + (4 if self.snake_blocks else 0))
        mb = max(1, self.l2_budget_bytes // max(1, per_utt))
        # keep at least ~2 waves of 256-row tiles on 148 SMs when the batch allows it
        tiles_per_utt = -(-L // 256)
        mb = min(B, max(mb, -(-296 // tiles_per_utt)))
        n_chunks = -(-B // mb)
        return int(-(-B // n_chunks))  # equal-sized chunks
