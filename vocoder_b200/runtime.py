"""Host-side runtime helpers: workspace buffers, weight-pack caching, CUDA-graph replay.

Nothing here computes: torch is used for device memory, streams and graph capture only.
"""
from __future__ import annotations

import functools
from collections import OrderedDict
from typing import Callable, Dict, Iterable, List, Tuple

import operator

import torch

from . import cabi


class Workspace:
    """Named device buffers reused across forward calls (stable addresses => CUDA-graph friendly).

    Buffers are grouped by the *signature* of the forward that asked for them (``enter(sig)``: input shape, precision
    mode, device).  At most ``max_signatures`` signatures stay resident; the least recently used group is dropped when a
    new one arrives, so synthesising a directory of files with ever-changing lengths does not grow device memory
    without bound.  Dropping buffers invalidates every CUDA graph that captured their addresses: listeners
    (``GraphedForward.invalidate``) are called on eviction.
    """

    def __init__(self, max_signatures: int = 4):
        self.max_signatures = max_signatures
        self._groups: "OrderedDict[Tuple, Dict[Tuple, torch.Tensor]]" = OrderedDict()
        self._sig: Tuple = ()
        self._listeners: List[Callable[[], None]] = []

    def add_listener(self, fn: Callable[[], None]) -> None:
        self._listeners.append(fn)

    def enter(self, sig: Tuple) -> None:
        """Declare the signature of the forward that is about to request buffers."""
        self._sig = sig
        if sig in self._groups:
            self._groups.move_to_end(sig)
            return
        self._groups[sig] = {}
        evicted = False
        while len(self._groups) > self.max_signatures:
            self._groups.popitem(last=False)
            evicted = True
        if evicted:
            for fn in self._listeners:
                fn()

    def get(self, name: str, shape: Tuple[int, ...], dtype: torch.dtype, device, zero: bool = False) -> torch.Tensor:
        group = self._groups.setdefault(self._sig, {})
        key = (name, tuple(shape), dtype, str(device))
        t = group.get(key)
        if t is None:
            # always zero-filled at creation: kernels that write only c < C (act_cast, resample_linear, snake_aa) rely
            # on the channel padding of operand buffers being zero (0-weights x NaN garbage would still be NaN)
            t = torch.zeros(shape, dtype=dtype, device=device)
            group[key] = t
        return t

    def f32(self, name, B, L, C, device):
        return self.get(name, (B, L, cabi.pitch_of(C)), torch.float32, device)

    def f16(self, name, B, L, C, device):
        return self.get(name, (B, L, cabi.f16_width(C)), torch.float16, device)  # strict layer: [hi | lo]

    def clear(self):
        self._groups.clear()
        for fn in self._listeners:
            fn()

    def saturation_report(self) -> Dict[str, int]:
        """fp16 operand buffers of the most recent forward -> number of elements clamped to +-65504 (operands saturate
        instead of overflowing: cvt.rn.satfinite).  Synchronises; a debugging aid, not part of the forward."""
        group = self._groups.get(self._sig, {})
        rep = {}
        for (name, _, dtype, _), t in group.items():
            if dtype == torch.float16 and t.is_cuda:
                n = int((t.abs() >= 65504.0).sum())
                if n:
                    rep[name] = n
        return rep

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for g in self._groups.values() for t in g.values())

    def n_signatures(self) -> int:
        return len(self._groups)


def forward_signature(*tensors) -> Tuple:
    """Workspace / graph signature of a forward: precision mode + shape, dtype and device of every input."""
    return (cabi.mode(), cabi.is_strict()) + tuple(
        (tuple(t.shape), t.dtype, t.device.index) if t is not None else None for t in tensors)


def params_key(tensors: Iterable[torch.Tensor]) -> Tuple:
    """Cheap fingerprint of a parameter set: changes when any tensor is updated in place, replaced or moved."""
    return (cabi.mode(), cabi.is_strict()) + tuple((t.data_ptr(), t._version, t.device.index) for t in tensors)


class _ModuleSlots:
    """Where a module tree keeps its parameters and buffers: [(dict, name)] slots + per-module structure counts.

    `module.parameters()` walks ~150 sub-modules of a weight-normed generator on every call (1.1 ms for HiFiGAN, 2.7 ms for
    BigVGAN on the build host) - more than a whole B = 1 forward takes on the GPU.  The slots are collected once; every
    forward only reads them (dict lookups), so tensors that are updated in place, moved (`.to()` swaps `.data`) or replaced
    in their slot (`load_state_dict(assign=True)`, optimizers that re-assign) still change the fingerprint, and a change of
    the tree itself (parametrizations removed, parameters / sub-modules added or deleted) changes the per-module counts and
    rebuilds the slot list."""

    def __init__(self, module: torch.nn.Module, buffers: bool):
        mods = list(module.modules())
        self.dicts = [d for m in mods for d in (m._parameters, m._buffers, m._modules)]
        self.counts = tuple(map(len, self.dicts))
        self.slots = [(m._parameters, n) for m in mods for n in m._parameters]
        if buffers:
            self.slots += [(m._buffers, n) for m in mods for n in m._buffers]
        # empty slots (bias=None) stay in the list: one that is filled later must change the fingerprint
        self.has_none = any(d[n] is None for d, n in self.slots)

    def valid(self) -> bool:
        return tuple(map(len, self.dicts)) == self.counts

    def tensors(self):
        ts = [d[n] for d, n in self.slots]          # (a slot that disappeared changes the counts first: valid() is called before)
        return [t for t in ts if t is not None] if self.has_none else ts


_data_ptr = operator.methodcaller("data_ptr")
_version = operator.attrgetter("_version")


def module_params_key(module: torch.nn.Module, buffers: bool = True) -> Tuple:
    """`params_key` over the parameters (and buffers) of `module`, without walking the module tree on every call."""
    st = module.__dict__.get("_fv_slots")
    if st is None or not st.valid():
        st = _ModuleSlots(module, buffers)
        module.__dict__["_fv_slots"] = st
    ts = st.tensors()
    return (cabi.mode(), cabi.is_strict(), ts[0].device if ts else None,
            tuple(map(_data_ptr, ts)), tuple(map(_version, ts)))


def require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise cabi.FvError(
            f"{who}: the generator forward runs only in the sm_100a CUDA kernels of libfv_b200.so; got a "
            f"{x.device} tensor. Move the module and its input to a CUDA device (there is no CPU fallback).")
    cabi.lib()  # raises loudly when the extension has not been built


def require_channels(x: torch.Tensor, channels: int, who: str) -> None:
    """A mismatched channel count would run silently (TMA zero-fills the missing weight columns): refuse it."""
    if x.ndim != 3 or x.shape[1] != channels:
        raise ValueError(f"{who}: expected input [B, {channels}, T], got {tuple(x.shape)}")


def with_precision(fn):
    """Run a module's forward under its ``precision`` attribute (see cabi.precision) and on the device of its input:
    launches use the CURRENT device's stream, so a module living on cuda:1 must make cuda:1 current for the call."""

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        x = args[0] if args else None
        with cabi.precision(getattr(self, "precision", None) or cabi.DEFAULT_PRECISION):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                with torch.cuda.device(x.device):
                    return fn(self, *args, **kwargs)
            return fn(self, *args, **kwargs)

    return wrapper


class GraphedForward:
    """Capture ``fn(static_inputs...) -> static_output`` once per input signature and replay it.

    The captured region contains only kernels of libfv_b200.so (launched on torch's capturing stream);
    replay removes the per-launch host cost of the ~80-150 kernels of one generator forward.  At most ``max_graphs``
    signatures are kept (least recently used first out).  ``tag`` is an extra key component: callers pass the pack
    generation of the weights the graph was captured with, so a re-pack (load_state_dict, remove_parametrizations,
    precision / engine change) can never replay pointers of freed packed weights.
    """

    def __init__(self, fn: Callable[..., torch.Tensor], warmup: int = 2, max_graphs: int = 4):
        self.fn = fn
        self.warmup = warmup
        self.max_graphs = max_graphs
        self._graphs: "OrderedDict[Tuple, Tuple]" = OrderedDict()
        self._tag = None

    def invalidate(self):
        self._graphs.clear()

    def __call__(self, *inputs: torch.Tensor, tag=None) -> torch.Tensor:
        if tag != self._tag:
            self.invalidate()
            self._tag = tag
        sig = forward_signature(*inputs)
        entry = self._graphs.get(sig)
        if entry is None:
            static_in = [None if t is None else t.clone() for t in inputs]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self.fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.fn(*static_in)
            entry = (graph, static_in, static_out)
            # the warm-up may have evicted a workspace group and with it every older graph: insert after it
            self._graphs[sig] = entry
            while len(self._graphs) > self.max_graphs:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(sig)
        graph, static_in, static_out = entry
        for s, t in zip(static_in, inputs):
            if s is not None:
                s.copy_(t, non_blocking=True)
        graph.replay()
        return static_out


def saturation_report(module: torch.nn.Module) -> Dict:
    """{"saturated": total, "buffers": {name: count}} over every workspace below `module` (see Workspace.saturation_report).
    Intermediates of the on-chip fused stages (fv_mrf_fused) never reach HBM and are not visible here."""
    bufs = {}
    for name, sub in module.named_modules():
        ws = getattr(sub, "_ws", None)
        if isinstance(ws, Workspace):
            for k, v in ws.saturation_report().items():
                bufs[f"{name + '.' if name else ''}{k}"] = v
    return {"saturated": sum(bufs.values()), "buffers": bufs}
