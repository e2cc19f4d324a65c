"""UnifyGenerator - drop-in mirror of ``fish_vocoder.modules.generators.unify.UnifyGenerator`` (reference file
fish_vocoder/modules/generators/unify.py:5-60): backbone -> (vq) -> head(x, template=...) -> [B, 1, L].
When backbone and head are both ours, the feature map stays channels-last fp16 between them (no layout round trip).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import cabi
from ..runtime import GraphedForward, require_cuda, with_precision


class UnifyGenerator(nn.Module):
    def __init__(self, backbone: nn.Module, head: nn.Module, vq: Optional[nn.Module] = None):
        super().__init__()
        self.backbone = backbone
        self.head = head
        self.vq = vq
        self.use_cuda_graph = False
        self.clone_graph_output = True
        self._graphed: Optional[GraphedForward] = None
        # the sub-modules' workspaces hold the buffers a graph captured here replays: dropping them drops the graph
        for m in (backbone, head):
            ws = getattr(m, "_ws", None)
            if ws is not None:
                ws.add_listener(self._drop_graphs)

    def _drop_graphs(self):
        if self._graphed is not None:
            self._graphed.invalidate()

    def _fused_ok(self) -> bool:
        return self.vq is None and hasattr(self.backbone, "_encode_cl") and hasattr(self.head, "_forward_cl")

    def _forward_fused(self, x, template):
        a0 = self.backbone._pack_input(x)
        mrf_head = getattr(self.head, "use_template", None) is not None   # MRF head (firefly-gan-base.yaml)
        # "mixed" precision: the ISTFT head's contractions are strict layers, so the encoder hands over [hi | lo]
        out_strict = cabi.is_mixed() and (self.head._trunk_strict() if mrf_head else True)
        h16, _ = self.backbone._encode_cl(a0, out_strict=out_strict)
        if mrf_head:
            y = self.head._forward_cl(h16, template)
        else:
            y = self.head._forward_cl(h16)
        return y[:, None, :] if y.ndim == 2 else y

    @with_precision
    def forward(self, x: torch.Tensor, template=None):
        require_cuda(x, "UnifyGenerator")
        if self._fused_ok():
            x = x.contiguous().float()
            if template is not None:
                template = template.contiguous().float()
            if self.use_cuda_graph and not torch.is_grad_enabled():
                self.backbone._ensure_packed(x.device)
                self.head._ensure_packed(x.device)
                if self._graphed is None:
                    self._graphed = GraphedForward(self._forward_fused)
                # the graph replays pointers into the sub-modules' packed weights: key it on their pack generations, so
                # load_state_dict / remove_parametrizations / an in-place weight edit / a precision or engine change
                # after a graphed forward re-captures instead of replaying freed memory
                tag = (self.backbone._pack_gen, self.head._pack_gen)
                y = self._graphed(x, template, tag=tag)
                return y.clone() if self.clone_graph_output else y
            return self._forward_fused(x, template)
        x = self.backbone(x)
        vq_result = None
        if self.vq is not None:
            vq_result = self.vq(x)
            x = vq_result.z
        x = self.head(x, template=template)
        if x.ndim == 2:
            x = x[:, None, :]
        if self.vq is not None:
            return x, vq_result
        return x

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        if self.vq is None:
            raise ValueError("VQ module is not present in the model.")
        return self.vq(self.backbone(x)).codes

    def decode(self, codes: torch.Tensor, template=None) -> torch.Tensor:
        if self.vq is None:
            raise ValueError("VQ module is not present in the model.")
        x = self.head(self.vq.from_codes(codes)[0], template=template)
        return x[:, None, :] if x.ndim == 2 else x

    def remove_parametrizations(self):
        for m in (self.backbone, self.head):
            if hasattr(m, "remove_parametrizations"):
                m.remove_parametrizations()
