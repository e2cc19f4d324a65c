"""ctypes binding of ``libfv_b200.so`` (C ABI declared in ``include/fv_vocoder.h``).

PyTorch is used for device memory and streams only: every function here passes raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream to the library.  There is NO CPU fallback: if the
shared library is missing or a tensor is not on a CUDA device the call raises.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
import threading
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfv_b200.so")

ACT_NONE, ACT_SILU, ACT_LEAKY, ACT_GELU, ACT_TANH, ACT_POLAR, ACT_SILU_TANH, ACT_SILU_H2 = range(8)
ENGINE_TC, ENGINE_SIMT = 0, 1
EDGE_MODES = {"replicate": 0, "reflect": 1, "zero": 2}  # enum fv_edge_mode
MAX_TAPS = 64

# every symbol include/fv_vocoder.h declares (tests check the library exports exactly these)
EXPORTS = [
    "fv_last_error", "fv_abi_version", "fv_launch_count", "fv_reset_launch_count", "fv_conv1d",
    "fv_set_tc_tuning", "fv_pack_input", "fv_unpack_output", "fv_conv_post_tanh", "fv_snake_aa",
    "fv_dwconv_layernorm", "fv_istft_ola", "fv_noise_conv", "fv_act_cast", "fv_resample_linear",
    "fv_mrf_fused", "fv_frame_audio", "fv_spec_mag", "fv_log_mel_out", "fv_debug_rowshift_probe",
    "fv_debug_umma_rate", "fv_debug_fma_rate",
]
MRF_MAX_BLOCKS, MRF_MAX_PAIRS, MRF_MAX_REACH = 4, 4, 32


class FvError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """Mirror of ``struct fv_conv_desc``."""
    _fields_ = [
        ("a", ctypes.c_void_p), ("B", ctypes.c_int32), ("L_in", ctypes.c_int32), ("a_pitch", ctypes.c_int32),
        ("w", ctypes.c_void_p), ("n_phase", ctypes.c_int32), ("n_taps", ctypes.c_int32),
        ("C_out", ctypes.c_int32), ("C_out_pad", ctypes.c_int32), ("w_pitch", ctypes.c_int32),
        ("tap_off", ctypes.POINTER(ctypes.c_int32)),
        ("L_out", ctypes.c_int32),
        ("bias", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("residual", ctypes.c_void_p),
        ("res_pitch", ctypes.c_int32),
        ("out32", ctypes.c_void_p), ("out32_pitch", ctypes.c_int32), ("accumulate", ctypes.c_int32),
        ("out_scale", ctypes.c_float),
        ("out16", ctypes.c_void_p), ("out16_pitch", ctypes.c_int32), ("act", ctypes.c_int32),
        ("act_param", ctypes.c_float),
        ("a_split", ctypes.c_int32), ("out16_split", ctypes.c_int32),
    ]


class MrfDesc(ctypes.Structure):
    """Mirror of ``struct fv_mrf_desc``."""
    _fields_ = [
        ("x", ctypes.c_void_p), ("B", ctypes.c_int32), ("L", ctypes.c_int32), ("C", ctypes.c_int32),
        ("x_pitch", ctypes.c_int32),
        ("w", ctypes.c_void_p), ("w_rows", ctypes.c_int32),
        ("bias", ctypes.c_void_p), ("n_blocks", ctypes.c_int32), ("n_pairs", ctypes.c_int32),
        ("ksize", ctypes.c_int32 * MRF_MAX_BLOCKS),
        ("dil1", (ctypes.c_int32 * MRF_MAX_PAIRS) * MRF_MAX_BLOCKS),
        ("dil2", (ctypes.c_int32 * MRF_MAX_PAIRS) * MRF_MAX_BLOCKS),
        ("w_row0", ((ctypes.c_int32 * 2) * MRF_MAX_PAIRS) * MRF_MAX_BLOCKS),
        ("act", ctypes.c_int32), ("act_param", ctypes.c_float),
        ("out32", ctypes.c_void_p), ("out32_pitch", ctypes.c_int32),
        ("out16", ctypes.c_void_p), ("out16_pitch", ctypes.c_int32),
        ("out_act", ctypes.c_int32), ("out_act_param", ctypes.c_float),
        ("accumulate", ctypes.c_int32), ("out_scale", ctypes.c_float),
    ]


_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FvError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(vocoder_b200/csrc/build.sh). There is no CPU fallback for the generator forward.")
    L = ctypes.CDLL(LIB_PATH)
    L.fv_last_error.restype = ctypes.c_char_p
    L.fv_abi_version.restype = ctypes.c_int
    L.fv_launch_count.restype = ctypes.c_int64
    L.fv_reset_launch_count.restype = None
    L.fv_set_tc_tuning.restype = None
    L.fv_set_tc_tuning.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.fv_conv1d.argtypes = [ctypes.POINTER(ConvDesc), ci, vp]
    L.fv_pack_input.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
    L.fv_unpack_output.argtypes = [vp, vp, ci, ci, ci, ci, vp]
    L.fv_conv_post_tanh.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
    L.fv_snake_aa.argtypes = [vp, vp, vp, vp, ctypes.POINTER(cf), ctypes.POINTER(cf), ci, ci, ci, ci, ci, ci, ci, vp]
    L.fv_dwconv_layernorm.argtypes = [vp, vp, vp, vp, vp, vp, vp, cf, ci, ci, ci, ci, ci, ci, vp]
    L.fv_istft_ola.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    L.fv_noise_conv.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]
    L.fv_act_cast.argtypes = [vp, vp, vp, vp, vp, ci, cf, ci, cf, cf, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
    L.fv_resample_linear.argtypes = [vp, vp, vp, ci, cf, ci, cf, ci, ci, ci, ci, ci, ci, ci, cf, ci, vp]
    L.fv_mrf_fused.argtypes = [ctypes.POINTER(MrfDesc), vp]
    L.fv_frame_audio.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]
    L.fv_spec_mag.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, cf, vp]
    L.fv_log_mel_out.argtypes = [vp, vp, ci, ci, ci, ci, cf, vp]
    L.fv_debug_rowshift_probe.argtypes = [vp, vp, vp, vp]
    L.fv_debug_umma_rate.argtypes = [ci, ci, ci, ci, vp, vp]
    L.fv_debug_fma_rate.argtypes = [ci, ci, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("fv_last_error", "fv_launch_count", "fv_reset_launch_count", "fv_set_tc_tuning"):
            fn.restype = ctypes.c_int
    if L.fv_abi_version() != 3:
        raise FvError("libfv_b200.so ABI version mismatch")
    _lib = L
    return L


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise FvError(f"{what} failed (rc={rc}): {lib().fv_last_error().decode()}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise FvError("vocoder_b200 kernels need CUDA tensors (no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's stream, tensor maps and kernel attributes are per device: a pointer of
        # another GPU would fault or silently use peer access.  Module forwards enter torch.cuda.device(x.device).
        raise FvError(f"tensor lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}: "
                      "wrap the call in torch.cuda.device(tensor.device)")
    if not t.is_contiguous():
        raise FvError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise FvError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ----------------------------------------------------------------------------------------------
# operand precision
#   "fp16"   (default) tensor-core operands are IEEE fp16 (10-bit mantissa = TF32 grade, the grade the reference
#            itself runs at on GPU: test.py:15 enables TF32), fp32 everywhere else.
#   "strict" operands carry a second fp16 word (value = hi + lo, 21+ mantissa bits) and every contraction is
#            evaluated as hi*Whi + lo*Whi + hi*Wlo in the same kernel (fv_conv_desc.a_split): fp32-grade results
#            at 3x the tensor work and 2x the operand bytes.  Activation buffers are then [B][L][2*pitch] = [hi | lo].
# The mode is thread-local state entered by the module's forward (``module.precision``).
# ----------------------------------------------------------------------------------------------
#   "mixed"  fp16 operands, except on the few contractions that dominate the waveform error (measured per layer on the
#            reference goldens, tests/diag/precision_report.py): the trunk of BigVGAN (conv_pre, ups, conv_post) and the stem /
#            downsample / head / inverse-DFT contractions of the Vocos path, which run strict.  The residual-block convs
#            and the ConvNeXt pointwise GEMMs (>= 95% of the tensor work) stay single fp16.
PRECISIONS = ("fp16", "mixed", "strict")
DEFAULT_PRECISION = "mixed"  # what a module without an explicit ``precision`` attribute runs in
_tls = threading.local()


def is_strict() -> bool:
    """True while the CURRENT layer carries [hi | lo] operands (whole forward in "strict", selected layers in "mixed")."""
    return getattr(_tls, "strict", False)


def mode() -> str:
    return getattr(_tls, "mode", "fp16")


def is_mixed() -> bool:
    return mode() == "mixed"


@contextlib.contextmanager
def precision(mode_: str):
    if mode_ not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}, got {mode_!r}")
    old = (is_strict(), mode())
    _tls.strict, _tls.mode = mode_ == "strict", mode_
    try:
        yield
    finally:
        _tls.strict, _tls.mode = old


@contextlib.contextmanager
def strict_layer(on: bool = True):
    """Operand layout of the enclosed pack / launch calls: [hi | lo] when `on` (a "mixed"-mode strict layer)."""
    old = is_strict()
    _tls.strict = bool(on) or old
    try:
        yield
    finally:
        _tls.strict = old


def pitch_of(channels: int) -> int:
    """Channel pitch of a channels-last activation buffer (strict mode: K chunks must not straddle hi | lo)."""
    if is_strict():
        return 16 if channels <= 16 else (32 if channels <= 32 else round_up(channels, 64))
    return round_up(channels, 8)


def f16_width(channels: int) -> int:
    """Halfs per row of an fp16 operand buffer: pitch, or [hi | lo] = 2 * pitch in strict mode."""
    p = pitch_of(channels)
    return 2 * p if is_strict() else p


def split_of(t16: Optional[torch.Tensor]) -> int:
    """The `split` argument of the fp16-producing entry points for operand buffer `t16`."""
    return t16.shape[-1] // 2 if (is_strict() and t16 is not None) else 0


def _hi_lo(w32: torch.Tensor):
    hi = w32.to(torch.float16)
    lo = (w32 - hi.float()).to(torch.float16)
    return hi, lo


def c_out_pad_of(c_out: int) -> int:
    """Weight rows per tap: a multiple of every N tile the kernel may pick for this C_out."""
    if c_out <= 128:
        p = 32  # the smallest N tile the tcgen05 engine picks (TMA epilogue boxes are 32 columns wide)
        while p < c_out:
            p *= 2
        return p
    return round_up(c_out, 256)


# ----------------------------------------------------------------------------------------------
# packed weights
# ----------------------------------------------------------------------------------------------
@dataclass
class PackedConv:
    """Weights of one fv_conv1d call, laid out [n_phase][n_taps][C_out_pad][w_pitch] fp16."""
    w: torch.Tensor
    bias: Optional[torch.Tensor]
    tap_off: Sequence[int]
    n_phase: int
    n_taps: int
    c_in: int
    c_out: int
    c_out_pad: int
    w_pitch: int
    split: int = 0  # strict mode: operand pitch P; the K axis of `w` is [Whi | Whi | Wlo] (w_pitch = 3P)
    # fp16 range guard: rows of `w` whose magnitudes sit outside fp16's comfortable range were multiplied by a power of two
    # s[o] at pack time (exact); w_scale = 1 / s is applied by the fp32 epilogue (through its layer-scale slot) and `bias`
    # already holds bias * s, so (acc * s + bias * s) / s == acc + bias.  None = no row needed it (the usual case).
    w_scale: Optional[torch.Tensor] = None

    def __post_init__(self):
        assert self.n_phase * self.n_taps <= MAX_TAPS, "too many taps for one call"
        self._tap_arr = (ctypes.c_int32 * len(self.tap_off))(*[int(v) for v in self.tap_off])
        self._gamma_cache = {}

    def to(self, device) -> "PackedConv":
        return PackedConv(self.w.to(device), None if self.bias is None else self.bias.to(device),
                          list(self.tap_off), self.n_phase, self.n_taps, self.c_in, self.c_out,
                          self.c_out_pad, self.w_pitch, self.split,
                          None if self.w_scale is None else self.w_scale.to(device))

    def gamma_for(self, gamma: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """The per-channel factor the epilogue applies: the caller's layer scale times the weight de-scaling."""
        if self.w_scale is None:
            return gamma
        if gamma is None:
            return self.w_scale
        key = (gamma.data_ptr(), gamma._version)
        g = self._gamma_cache.get(key)
        if g is None:
            self._gamma_cache.clear()
            g = (gamma.detach().float() * self.w_scale).contiguous()
            self._gamma_cache[key] = g
        return g


# rows whose largest |w| falls outside [2^-9, 2^13] are rescaled: fp16 is normal down to 6.1e-5 (2^-14) and the row's small
# entries need ~5 more bits below its maximum to keep the contraction's relative error at the 2^-11 of a well-scaled row
W_SCALE_LO, W_SCALE_HI = 2.0 ** -9, 2.0 ** 13


def weight_row_scale(amax: torch.Tensor) -> Optional[torch.Tensor]:
    """amax [C_out] = max |w| of every output channel.  None when every (non-zero) row is inside the comfortable fp16
    range; otherwise the power-of-two scale s[o] that brings the row maximum into [0.5, 1).  One device sync per call."""
    amax = amax.detach().float()
    inf = torch.full_like(amax, float("inf"))
    lo, hi = torch.stack([torch.where(amax > 0, amax, inf).min(), amax.max()]).tolist()
    if hi <= 0.0 or (lo >= W_SCALE_LO and hi <= W_SCALE_HI):
        return None
    nz = amax > 0
    e = torch.ceil(torch.log2(torch.where(nz, amax, torch.ones_like(amax))))
    return torch.where(nz, torch.exp2(-e), torch.ones_like(amax))


def _scaled_bias(bias: Optional[torch.Tensor], s: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if bias is None:
        return None
    b = bias.detach().float()
    return (b if s is None else b * s).contiguous()


def _alloc_w(n_phase: int, n_taps: int, cp: int, wp: int, device):
    """Zeroed packed-weight tensor and the `put(phase, tap, W[c_out, c_in])` writer for the active precision."""
    strict = is_strict()
    w = torch.zeros(n_phase, n_taps, cp, 3 * wp if strict else wp, dtype=torch.float16, device=device)

    def put(r: int, i: int, m32: torch.Tensor):
        co, ci = m32.shape
        if strict:
            hi, lo = _hi_lo(m32)
            w[r, i, :co, :ci] = hi
            w[r, i, :co, wp:wp + ci] = hi
            w[r, i, :co, 2 * wp:2 * wp + ci] = lo
        else:
            w[r, i, :co, :ci] = m32.to(torch.float16)

    return w, put, (wp if strict else 0)


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], dilation: int = 1) -> PackedConv:
    """nn.Conv1d weight [C_out, C_in, k] with "same" padding (k*d-d)//2 (hifigan.py:21-22)."""
    c_out, c_in, k = weight.shape
    assert k % 2 == 1, "same-padded convs on this path have odd kernels"
    cp, wp = c_out_pad_of(c_out), pitch_of(c_in)
    w, put, split = _alloc_w(1, k, cp, wp, weight.device)
    wf = weight.detach().float()
    s = weight_row_scale(wf.abs().amax(dim=(1, 2)))
    if s is not None:
        wf = wf * s[:, None, None]
    for j in range(k):
        put(0, j, wf[:, :, j])
    offs = [(j - (k - 1) // 2) * dilation for j in range(k)]
    return PackedConv(w.contiguous(), _scaled_bias(bias, s), offs, 1, k, c_in, c_out, cp, w.shape[-1], split,
                      None if s is None else (1.0 / s).contiguous())


def pack_taps(mats: Sequence[torch.Tensor], offsets: Sequence[int], bias: Optional[torch.Tensor] = None) -> PackedConv:
    """General single-phase packer: out[q] = sum_i mats[i] @ a[q + offsets[i]], mats[i] = [C_out, C_in]."""
    c_out, c_in = mats[0].shape
    cp, wp = c_out_pad_of(c_out), pitch_of(c_in)
    w, put, split = _alloc_w(1, len(mats), cp, wp, mats[0].device)
    s = weight_row_scale(torch.stack([m.detach().float().abs().amax(dim=1) for m in mats]).amax(dim=0))
    for i, m in enumerate(mats):
        mf = m.detach().float()
        put(0, i, mf if s is None else mf * s[:, None])
    return PackedConv(w.contiguous(), _scaled_bias(bias, s), [int(o) for o in offsets], 1, len(mats), c_in, c_out, cp,
                      w.shape[-1], split, None if s is None else (1.0 / s).contiguous())


def pack_conv_row_pairs(weight: torch.Tensor, bias: Optional[torch.Tensor], dilation: int = 1) -> PackedConv:
    """The "same"-padded Conv1d of `pack_conv`, evaluated on PAIRS of time steps.  A channels-last buffer [B, L, C] whose rows
    are exactly C values apart (pitch == C) and whose L is even is also [B, L/2, 2C]; on that view the conv is a conv with
    2C x 2C tap matrices: out[2m + p] = sum_j W_j a[2m + p + o_j] with p + o_j = 2q + r puts W_j into block (p, r) of the
    tap at pair offset q.  For C = 16 this turns half-empty 32-column tiles (16 real channels) into full ones and halves
    the rows every epilogue touches (BigVGAN's last stage: 48 / 67 -> ~30 / 42 us per launch); a dilated tap list gets
    longer (k = 11, d = 5: 17 pair taps) but stays far below the launch's memory time.  Call through `conv1d_row_pairs`."""
    c_out, c_in, k = weight.shape
    assert k % 2 == 1, "same-padded convs on this path have odd kernels"
    wf = weight.detach().float()
    half = (k - 1) // 2
    mats = {}
    for p in (0, 1):
        for j in range(k):
            q, r = divmod(p + (j - half) * dilation, 2)
            m = mats.get(q)
            if m is None:
                m = mats[q] = torch.zeros(2 * c_out, 2 * c_in, dtype=torch.float32, device=weight.device)
            m[p * c_out:(p + 1) * c_out, r * c_in:(r + 1) * c_in] += wf[:, :, j]
    offs = sorted(mats)
    b2 = None if bias is None else torch.cat([bias.detach().float(), bias.detach().float()])
    return pack_taps([mats[q] for q in offs], offs, b2)


def row_pairs_ok(channels: int, L: int) -> bool:
    """True when [B, L, C] buffers of this layer can be viewed as [B, L/2, 2C] (no channel padding, even length)."""
    return L % 2 == 0 and pitch_of(channels) == channels and f16_width(channels) == channels


def _pair_view(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    B, L, P = t.shape
    if L % 2 or not t.is_contiguous():
        raise FvError(f"row-pair view needs a contiguous [B, even L, C] buffer, got {tuple(t.shape)} / strides {t.stride()}")
    return t.view(B, L // 2, 2 * P)


def conv1d_row_pairs(a16: torch.Tensor, pc: PackedConv, *, residual=None, out32=None, accumulate=False, out_scale=1.0,
                     out16=None, act=ACT_NONE, act_param=0.0, engine=ENGINE_TC) -> None:
    """`conv1d` for weights packed by `pack_conv_row_pairs`: every [B, L, C] buffer is passed as its [B, L/2, 2C] view."""
    conv1d(_pair_view(a16), pc, residual=_pair_view(residual), out32=_pair_view(out32), accumulate=accumulate,
           out_scale=out_scale, out16=_pair_view(out16), act=act, act_param=act_param, engine=engine, out16_split=0)


def pack_linear(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> PackedConv:
    """nn.Linear / 1x1 conv weight [C_out, C_in]."""
    return pack_conv(weight.reshape(weight.shape[0], weight.shape[1], 1), bias)


def pack_conv_transpose(weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int) -> PackedConv:
    """nn.ConvTranspose1d weight [C_in, C_out, k], padding (k-u)//2, as u polyphase convs:
    y[q*u + r] = sum_m x[q - m] * w[:, :, m*u + r + p]  over m with 0 <= m*u + r + p < k."""
    c_in, c_out, k = weight.shape
    u = stride
    p = (k - u) // 2
    phases = []
    for r in range(u):
        taps = []
        m = -((r + p) // u)
        while m * u + r + p < k:
            j = m * u + r + p
            if j >= 0:
                taps.append((-m, j))
            m += 1
        phases.append(taps)
    n_taps = max(1, max(len(t) for t in phases))
    cp, wp = c_out_pad_of(c_out), pitch_of(c_in)
    w, put, split = _alloc_w(u, n_taps, cp, wp, weight.device)
    offs = []
    wf = weight.detach().float()
    s = weight_row_scale(wf.abs().amax(dim=(0, 2)))
    if s is not None:
        wf = wf * s[None, :, None]
    for r, taps in enumerate(phases):
        for i in range(n_taps):
            if i < len(taps):
                off, j = taps[i]
                put(r, i, wf[:, :, j].t())
                offs.append(off)
            else:
                offs.append(0)
    return PackedConv(w.contiguous(), _scaled_bias(bias, s), offs, u, n_taps, c_in, c_out, cp, w.shape[-1], split,
                      None if s is None else (1.0 / s).contiguous())


def conv_transpose_out_len(L: int, k: int, u: int) -> int:
    return (L - 1) * u - 2 * ((k - u) // 2) + k


# ----------------------------------------------------------------------------------------------
# op wrappers
# ----------------------------------------------------------------------------------------------
def conv1d(a16: torch.Tensor, pc: PackedConv, L_out: Optional[int] = None, *, gamma=None, residual=None,
           out32=None, accumulate=False, out_scale=1.0, out16=None, act=ACT_NONE, act_param=0.0,
           engine=ENGINE_TC, use_bias=True, out16_split: Optional[int] = None) -> None:
    """a16 [B, L_in, a_pitch] fp16 -> out32 [B, L_out, >=C_out] fp32 and/or out16 fp16 (see fv_conv_desc)."""
    B, L_in, a_pitch = a16.shape
    if L_out is None:
        L_out = L_in
    for name, t in (("residual", residual), ("out32", out32), ("out16", out16)):
        if t is not None and (t.shape[0] != B or t.shape[1] != L_out):
            raise FvError(f"{name} has shape {tuple(t.shape)}, expected [{B}, {L_out}, pitch]")
    if (pc.n_phase == 1 and pc.n_taps == 1 and int(pc.tap_off[0]) == 0 and L_out == L_in and B > 1
            and all(t is None or t.is_contiguous() for t in (a16, residual, out32, out16))):
        # pointwise conv / linear layer: rows of different utterances never mix, so [B][L] is one GEMM M axis of B*L rows
        # (tiles of 128/256 rows then fill completely: Vocos has L = 94 rows per utterance, 73% of a 128-row tile)
        B, L_in, L_out = 1, B * L_in, B * L_in
    d = ConvDesc()
    d.a, d.B, d.L_in, d.a_pitch = _ptr(a16, torch.float16), B, L_in, a_pitch
    d.w, d.n_phase, d.n_taps = _ptr(pc.w, torch.float16), pc.n_phase, pc.n_taps
    d.C_out, d.C_out_pad, d.w_pitch = pc.c_out, pc.c_out_pad, pc.w_pitch
    d.tap_off = pc._tap_arr
    d.L_out = L_out
    d.bias = _ptr(pc.bias, torch.float32) if (use_bias and pc.bias is not None) else None
    d.gamma = _ptr(pc.gamma_for(gamma), torch.float32)
    d.residual, d.res_pitch = _ptr(residual, torch.float32), (0 if residual is None else residual.shape[2])
    d.out32, d.out32_pitch = _ptr(out32, torch.float32), (0 if out32 is None else out32.shape[2])
    d.accumulate, d.out_scale = int(bool(accumulate)), float(out_scale)
    d.out16, d.out16_pitch = _ptr(out16, torch.float16), (0 if out16 is None else out16.shape[2])
    d.act, d.act_param = int(act), float(act_param)
    if pc.split and a_pitch != 2 * pc.split:
        raise FvError(f"strict-precision weights (operand pitch {pc.split}) need a [hi | lo] operand of "
                      f"{2 * pc.split} halfs per row, got {a_pitch}")
    if not pc.split and is_strict():
        raise FvError("weights were packed in fp16 mode but the call runs in strict mode: repack")
    # out16_split: None = layout of the current layer context; an int = the consumer's layout ("mixed": a plain layer
    # feeding a strict one writes [hi | lo], a strict layer feeding a plain one writes plain fp16)
    d.a_split, d.out16_split = pc.split, (split_of(out16) if out16_split is None else int(out16_split))
    _check(lib().fv_conv1d(ctypes.byref(d), int(engine), _stream()), "fv_conv1d")


def pack_input(x: torch.Tensor, pitch: Optional[int] = None) -> torch.Tensor:
    """[B, C, T] fp32 channels-first -> [B, T, pitch] fp16 channels-last."""
    B, C, T = x.shape
    pitch = pitch or pitch_of(C)
    split = pitch if is_strict() else 0
    out = torch.empty(B, T, pitch + split, dtype=torch.float16, device=x.device)
    _check(lib().fv_pack_input(_ptr(x, torch.float32), _ptr(out), B, C, T, pitch, split, _stream()),
           "fv_pack_input")
    return out


def unpack_output(x32: torch.Tensor, C: int) -> torch.Tensor:
    B, L, pitch = x32.shape
    out = torch.empty(B, C, L, dtype=torch.float32, device=x32.device)
    _check(lib().fv_unpack_output(_ptr(x32, torch.float32), _ptr(out), B, C, L, pitch, _stream()),
           "fv_unpack_output")
    return out


def conv_post_tanh(a16: torch.Tensor, w32: torch.Tensor, bias: Optional[torch.Tensor], C: int,
                   apply_tanh: bool = True, out: Optional[torch.Tensor] = None,
                   split: Optional[int] = None) -> torch.Tensor:
    """a16 [B, L, pitch] fp16 ([hi | lo] when split > 0), w32 [k, C] fp32 -> wav [B, 1, L] fp32."""
    B, L, width = a16.shape
    split = split_of(a16) if split is None else int(split)
    k = w32.shape[0]
    if out is None:
        out = torch.empty(B, 1, L, dtype=torch.float32, device=a16.device)
    _check(lib().fv_conv_post_tanh(_ptr(a16, torch.float16), _ptr(w32, torch.float32), _ptr(bias, torch.float32),
                                   _ptr(out, torch.float32), B, L, C, width - split, k, int(apply_tanh), split,
                                   _stream()), "fv_conv_post_tanh")
    return out


def snake_aa(x32: torch.Tensor, out16: torch.Tensor, alpha: torch.Tensor, beta: Optional[torch.Tensor],
             filt_up: Sequence[float], filt_down: Sequence[float], C: int, logscale: bool = True,
             split: Optional[int] = None, edge_mode: str = "replicate") -> None:
    if edge_mode not in EDGE_MODES:
        raise ValueError(f"edge_mode must be one of {tuple(EDGE_MODES)}, got {edge_mode!r}")
    B, L, pitch = x32.shape
    fu = (ctypes.c_float * 12)(*[float(v) for v in filt_up])
    fd = (ctypes.c_float * 12)(*[float(v) for v in filt_down])
    split = split_of(out16) if split is None else int(split)
    _check(lib().fv_snake_aa(_ptr(x32, torch.float32), _ptr(out16, torch.float16), _ptr(alpha, torch.float32),
                             _ptr(beta, torch.float32), fu, fd, int(logscale), B, L, C, pitch, split,
                             EDGE_MODES[edge_mode], _stream()),
           "fv_snake_aa")


def dwconv_layernorm(x32: torch.Tensor, C: int, dw_w, dw_b, ln_w, ln_b, eps: float, k: int,
                     out16: Optional[torch.Tensor] = None, out32: Optional[torch.Tensor] = None,
                     split: Optional[int] = None) -> None:
    B, T, pitch = x32.shape
    split = split_of(out16) if split is None else int(split)
    _check(lib().fv_dwconv_layernorm(_ptr(x32, torch.float32), _ptr(out16, torch.float16),
                                     _ptr(out32, torch.float32), _ptr(dw_w, torch.float32),
                                     _ptr(dw_b, torch.float32), _ptr(ln_w, torch.float32),
                                     _ptr(ln_b, torch.float32), float(eps), B, T, C, pitch, int(k),
                                     split, _stream()),
           "fv_dwconv_layernorm")


def istft_ola(frames: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int,
              out: Optional[torch.Tensor] = None, center: bool = False) -> torch.Tensor:
    """frames [B, T, >=n_fft] fp32 -> wav [B, T*hop] (padding="same") or [B, (T-1)*hop] (padding="center")."""
    B, T, fp = frames.shape
    if out is None:
        out = torch.empty(B, (T - 1) * hop if center else T * hop, dtype=torch.float32, device=frames.device)
    _check(lib().fv_istft_ola(_ptr(frames, torch.float32), _ptr(window, torch.float32), _ptr(out), B, T, n_fft,
                              hop, fp, int(bool(center)), _stream()), "fv_istft_ola")
    return out


def noise_conv(tpl: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out32: torch.Tensor, C: int, k: int,
               stride: int, pad: int) -> None:
    B, L_out, pitch = out32.shape
    L_audio = tpl.shape[-1]
    _check(lib().fv_noise_conv(_ptr(tpl, torch.float32), _ptr(w, torch.float32), _ptr(bias, torch.float32),
                               _ptr(out32, torch.float32), B, L_audio, L_out, C, pitch, k, stride, pad, _stream()),
           "fv_noise_conv")


def act_cast(x32: torch.Tensor, C: int, act: int, act_param: float = 0.0, *, noise=None, noise_w=None,
             out16=None, out16_coff: int = 0, act16: int = ACT_NONE, act16_param: float = 0.0, out32=None,
             out_scale: float = 1.0, accumulate: bool = False) -> None:
    B, L, in_pitch = x32.shape
    _check(lib().fv_act_cast(_ptr(x32, torch.float32), _ptr(noise, torch.float32), _ptr(noise_w, torch.float32),
                             _ptr(out16, torch.float16), _ptr(out32, torch.float32), int(act), float(act_param),
                             int(act16), float(act16_param), float(out_scale), int(bool(accumulate)), B, L, C,
                             in_pitch, 0 if out16 is None else out16.shape[2], out16_coff,
                             0 if out32 is None else out32.shape[2], split_of(out16), _stream()), "fv_act_cast")


def resample_linear(x32: torch.Tensor, C: int, L_out: int, scale: float, *, pre_act=ACT_NONE, pre_param=0.0,
                    act=ACT_NONE, act_param=0.0, out16=None, out32=None, out_coff: int = 0) -> None:
    B, L_in, in_pitch = x32.shape
    o = out16 if out16 is not None else out32
    _check(lib().fv_resample_linear(_ptr(x32, torch.float32), _ptr(out32, torch.float32),
                                    _ptr(out16, torch.float16), int(pre_act), float(pre_param), int(act),
                                    float(act_param), B, L_in, L_out, C, in_pitch, o.shape[2], out_coff, float(scale),
                                    split_of(out16), _stream()), "fv_resample_linear")


# ----------------------------------------------------------------------------------------------
# mel front-end helpers (fv_frontend.cu)
# ----------------------------------------------------------------------------------------------
def frame_audio(y: torch.Tensor, hop: int, pad_left: int, pad_right: int, rows: int) -> torch.Tensor:
    """y [B, L] fp32 -> fp16 operand [B, rows, pitch(hop)] of the reflect-padded signal, `hop` samples per row."""
    B, L = y.shape
    pitch = pitch_of(hop)
    split = pitch if is_strict() else 0
    out = torch.empty(B, rows, pitch + split, dtype=torch.float16, device=y.device)
    _check(lib().fv_frame_audio(_ptr(y, torch.float32), _ptr(out), B, L, hop, pad_left, pad_right, rows, pitch, split,
                                _stream()), "fv_frame_audio")
    return out


def spec_mag(spec: torch.Tensor, F: int, eps: float, *, out16: Optional[torch.Tensor] = None,
             out32: Optional[torch.Tensor] = None) -> None:
    """spec [B, T, >=2F] fp32 ([re | im]) -> sqrt(re^2 + im^2 + eps) as fp16 operand and/or fp32 [B, T, >=F]."""
    B, T, sp = spec.shape
    pitch = (out16.shape[2] // (2 if is_strict() else 1)) if out16 is not None else out32.shape[2]
    _check(lib().fv_spec_mag(_ptr(spec, torch.float32), _ptr(out16, torch.float16), _ptr(out32, torch.float32), B, T, F,
                             sp, pitch, split_of(out16), 0 if out32 is None else out32.shape[2], float(eps),
                             _stream()), "fv_spec_mag")


def log_mel_out(x32: torch.Tensor, C: int, floor: float) -> torch.Tensor:
    """x32 [B, T, pitch] fp32 channels-last -> log(max(x, floor)) as [B, C, T] fp32."""
    B, T, pitch = x32.shape
    out = torch.empty(B, C, T, dtype=torch.float32, device=x32.device)
    _check(lib().fv_log_mel_out(_ptr(x32, torch.float32), _ptr(out), B, C, T, pitch, float(floor), _stream()),
           "fv_log_mel_out")
    return out


# ----------------------------------------------------------------------------------------------
# fused MRF stage (fv_mrf_fused)
# ----------------------------------------------------------------------------------------------
@dataclass
class PackedMrf:
    """Weights of one fused MRF stage: every [C_out, C_in] tap tile of every conv, concatenated row-wise."""
    w: torch.Tensor        # fp16 [rows, C]
    bias: torch.Tensor     # fp32 [n_blocks, n_pairs, 2, C]
    C: int
    ksize: Sequence[int]
    dil1: Sequence[Sequence[int]]
    dil2: Sequence[Sequence[int]]
    w_row0: Sequence[Sequence[Sequence[int]]]


def mrf_fusable(C: int, blocks, pairwise: bool = False) -> bool:
    """blocks: [(convs1, convs2)] of nn.Conv1d-like modules.  True when fv_mrf_fused covers the stage: as one launch for
    C in {16, 32, 64}; pairwise=True asks whether every (conv, conv) pair can be its own launch (C = 128)."""
    if pairwise:
        return C in (64, 128) and len(blocks) >= 1 and all(
            _mrf_fusable(C, [([c1], [c2])], _tile=256, _chans=(64, 128)) for c1s, c2s in blocks for c1, c2 in zip(c1s, c2s))
    return _mrf_fusable(C, blocks)


def _mrf_fusable(C: int, blocks, _tile: int = 512, _chans=(16, 32, 64)) -> bool:
    if is_strict() or C not in _chans or not 1 <= len(blocks) <= MRF_MAX_BLOCKS:
        return False
    n_pairs = len(blocks[0][0])
    halo = 0
    for c1s, c2s in blocks:
        if len(c1s) != n_pairs or len(c2s) != n_pairs or not 1 <= n_pairs <= MRF_MAX_PAIRS:
            return False
        k = c1s[0].kernel_size[0]
        h = 0
        for c in list(c1s) + list(c2s):
            if c.kernel_size[0] != k or k % 2 == 0 or c.in_channels != C or c.out_channels != C or c.bias is None:
                return False
            if weight_row_scale(c.weight.detach().abs().amax(dim=(1, 2))) is not None:
                return False  # fp16 range guard: the layer-wise path rescales such rows, the fused kernel cannot
            reach = (k - 1) // 2 * c.dilation[0]
            if reach > MRF_MAX_REACH:
                return False
            h += reach
        halo = max(halo, h)
    return _tile - round_up(halo, 32) - halo >= 32


def pack_mrf(C: int, blocks) -> PackedMrf:
    """blocks: [(convs1, convs2)] with folded weights [C, C, k] (hifigan.py:29-98)."""
    tiles, biases, ksize, dil1, dil2, row0 = [], [], [], [], [], []
    rows = 0
    for c1s, c2s in blocks:
        k = c1s[0].kernel_size[0]
        ksize.append(k)
        dil1.append([c.dilation[0] for c in c1s])
        dil2.append([c.dilation[0] for c in c2s])
        r_blk, b_blk = [], []
        for c1, c2 in zip(c1s, c2s):
            r_pair = []
            for c in (c1, c2):
                w = c.weight.detach().float()                         # [C_out, C_in, k]
                tiles.append(w.permute(2, 0, 1).reshape(k * C, C).to(torch.float16))
                r_pair.append(rows)
                rows += k * C
            r_blk.append(r_pair)
            b_blk.append(torch.stack([c1.bias.detach().float(), c2.bias.detach().float()]))
        row0.append(r_blk)
        biases.append(torch.stack(b_blk))
    return PackedMrf(torch.cat(tiles).contiguous(), torch.stack(biases).contiguous(), C, ksize, dil1, dil2, row0)


def mrf_fused(x32: torch.Tensor, pm: PackedMrf, out32: torch.Tensor, *, out16: Optional[torch.Tensor] = None,
              act: int = ACT_SILU, act_param: float = 0.0, out_act: int = ACT_NONE, out_act_param: float = 0.0,
              accumulate: bool = False, out_scale: float = 0.0) -> None:
    """x32 [B, L, pitch] fp32 -> out32 = mean over residual blocks (and out16 = fp16(out_act(out32))).
    accumulate / out_scale: out32 += out_scale * block(x) (pair-wise evaluation of a C = 128 stage; out_scale 0 = 1/n_blocks).
    out32 must not alias x32: tiles read their neighbours' rows of x as halo."""
    if out32.data_ptr() == x32.data_ptr():
        raise FvError("fv_mrf_fused: out32 must not alias x (tiles read their neighbours' rows as halo)")
    B, L, pitch = x32.shape
    d = MrfDesc()
    d.x, d.B, d.L, d.C, d.x_pitch = _ptr(x32, torch.float32), B, L, pm.C, pitch
    d.w, d.w_rows, d.bias = _ptr(pm.w, torch.float16), pm.w.shape[0], _ptr(pm.bias, torch.float32)
    d.n_blocks, d.n_pairs = len(pm.ksize), len(pm.dil1[0])
    for j, k in enumerate(pm.ksize):
        d.ksize[j] = k
        for i in range(d.n_pairs):
            d.dil1[j][i], d.dil2[j][i] = pm.dil1[j][i], pm.dil2[j][i]
            d.w_row0[j][i][0], d.w_row0[j][i][1] = pm.w_row0[j][i]
    d.act, d.act_param = int(act), float(act_param)
    for name, t in (("out32", out32), ("out16", out16)):
        if t is not None and (t.shape[0] != B or t.shape[1] != L):
            raise FvError(f"{name} has shape {tuple(t.shape)}, expected [{B}, {L}, pitch]")
    d.out32, d.out32_pitch = _ptr(out32, torch.float32), out32.shape[2]
    d.out16, d.out16_pitch = _ptr(out16, torch.float16), (0 if out16 is None else out16.shape[2])
    d.out_act, d.out_act_param = int(out_act), float(out_act_param)
    d.accumulate, d.out_scale = int(bool(accumulate)), float(out_scale)
    _check(lib().fv_mrf_fused(ctypes.byref(d), _stream()), "fv_mrf_fused")


def launch_count() -> int:
    return int(lib().fv_launch_count())


def reset_launch_count() -> None:
    lib().fv_reset_launch_count()


def set_tc_tuning(block_n: int = 0, m_sub: int = 0, epilogue: int = 0, mainloop: int = 0) -> None:
    """epilogue: 0 auto, 1 LSU (smem transpose + coalesced ld/st.global), 2 TMA (bulk tensor load/store);
    mainloop: 0 auto, 1 per-tap stages, 2 operand slab + row-shifted descriptors."""
    lib().fv_set_tc_tuning(int(block_n), int(m_sub), int(epilogue), int(mainloop))
