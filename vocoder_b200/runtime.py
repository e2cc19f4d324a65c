"""Host-side runtime helpers: workspace buffers, weight-pack caching, CUDA-graph replay.

Nothing here computes: torch is used for device memory, streams and graph capture only.
"""
from __future__ import annotations

import functools
from typing import Callable, Dict, Iterable, Tuple

import torch

from . import cabi


class Workspace:
    """Named device buffers reused across forward calls (stable addresses => CUDA-graph friendly)."""

    def __init__(self):
        self._bufs: Dict[Tuple, torch.Tensor] = {}

    def get(self, name: str, shape: Tuple[int, ...], dtype: torch.dtype, device, zero: bool = False) -> torch.Tensor:
        key = (name, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            # always zero-filled at creation: kernels that write only c < C (act_cast, resample_linear, snake_aa) rely
            # on the channel padding of operand buffers being zero (0-weights x NaN garbage would still be NaN)
            t = torch.zeros(shape, dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    def f32(self, name, B, L, C, device):
        return self.get(name, (B, L, cabi.pitch_of(C)), torch.float32, device)

    def f16(self, name, B, L, C, device):
        return self.get(name, (B, L, cabi.f16_width(C)), torch.float16, device)  # strict mode: [hi | lo]

    def clear(self):
        self._bufs.clear()

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._bufs.values())


def params_key(tensors: Iterable[torch.Tensor]) -> Tuple:
    """Cheap fingerprint of a parameter set: changes when any tensor is updated in place, replaced or moved."""
    return (cabi.is_strict(),) + tuple((t.data_ptr(), t._version, t.device.index) for t in tensors)


def require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise cabi.FvError(
            f"{who}: the generator forward runs only in the sm_100a CUDA kernels of libfv_b200.so; got a "
            f"{x.device} tensor. Move the module and its input to a CUDA device (there is no CPU fallback).")
    cabi.lib()  # raises loudly when the extension has not been built


def with_precision(fn):
    """Run a module's forward under its ``precision`` attribute ("fp16" default | "strict", see cabi.precision)."""

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with cabi.precision(getattr(self, "precision", "fp16")):
            return fn(self, *args, **kwargs)

    return wrapper


class GraphedForward:
    """Capture ``fn(static_inputs...) -> static_output`` once per input signature and replay it.

    The captured region contains only kernels of libfv_b200.so (launched on torch's capturing stream);
    replay removes the per-launch host cost of the ~80-150 kernels of one generator forward.
    """

    def __init__(self, fn: Callable[..., torch.Tensor], warmup: int = 2):
        self.fn = fn
        self.warmup = warmup
        self._graphs: Dict[Tuple, Tuple] = {}

    def invalidate(self):
        self._graphs.clear()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        sig = (cabi.is_strict(),) + tuple((tuple(t.shape), t.dtype, t.device.index) if t is not None else None
                                          for t in inputs)
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
            self._graphs[sig] = entry
        graph, static_in, static_out = entry
        for s, t in zip(static_in, inputs):
            if s is not None:
                s.copy_(t, non_blocking=True)
        graph.replay()
        return static_out
