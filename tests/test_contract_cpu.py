"""CPU-side checks of the drop-in boundary: checkpoint layout, constructor surface, C-ABI exports, weight packing,
and that the product path refuses to run without CUDA (no silent CPU fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch
import torch.nn.functional as F

from tests.util import GOLDEN, load_golden
from vocoder_b200 import cabi
from vocoder_b200.encoders import ConvNeXtEncoder
from vocoder_b200.generators import (BigVGANGenerator, HiFiGANGenerator, ISTFTHead, RefineGANGenerator,
                                     UnifyGenerator)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _contract():
    with open(os.path.join(GOLDEN, "state_dict_contract.json")) as f:
        return json.load(f)


def _build(name):
    if name == "hifigan_cfgA":
        return HiFiGANGenerator(hop_length=256, upsample_rates=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4),
                                num_mels=80, use_template=False)
    if name == "hifigan_yaml_44k":
        return HiFiGANGenerator(hop_length=512, num_mels=128, use_template=False)
    if name == "bigvgan_cfgC":
        return BigVGANGenerator(hop_length=512, num_mels=100, use_template=False)
    if name == "vocos_yaml":
        return UnifyGenerator(
            backbone=ConvNeXtEncoder(input_channels=128, depths=[3, 3, 27, 3], dims=[128, 256, 512, 1024],
                                     drop_path_rate=0.4, kernel_size=7),
            head=ISTFTHead(dim=1024, n_fft=2048, hop_length=512, win_length=2048, padding="same"))
    if name == "refinegan_default":
        return RefineGANGenerator()
    raise KeyError(name)


@pytest.mark.parametrize("name", ["hifigan_cfgA", "hifigan_yaml_44k", "bigvgan_cfgC", "vocos_yaml", "refinegan_default"])
def test_state_dict_layout_matches_reference(name):
    want = _contract()[name]
    got = {k: list(v.shape) for k, v in _build(name).state_dict().items()}
    assert got == want


@pytest.mark.parametrize("name,cls", [("hifigan_small_ref", HiFiGANGenerator),
                                      ("hifigan_template_stress", HiFiGANGenerator),
                                      ("bigvgan_small_stress", BigVGANGenerator),
                                      ("refinegan_small_stress", RefineGANGenerator)])
def test_reference_state_dict_loads_strict(name, cls):
    kwargs, sd, _, _, _ = load_golden(name)
    m = cls(**kwargs)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_lightning_checkpoint_prefix_roundtrip():
    # test.py:32-37 loads ckpt["state_dict"] with keys "generator.<...>" into GANModel
    kwargs, sd, _, _, _ = load_golden("hifigan_small_ref")
    ckpt = {"state_dict": {"generator." + k: v for k, v in sd.items()}}
    m = HiFiGANGenerator(**kwargs)
    stripped = {k[len("generator."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("generator.")}
    m.load_state_dict(stripped, strict=True)


def test_remove_parametrizations_changes_layout_like_reference():
    kwargs, sd, _, _, _ = load_golden("hifigan_small_ref")
    m = HiFiGANGenerator(**kwargs)
    m.load_state_dict(sd)
    w_before = m.conv_pre.weight.detach().clone()
    m.remove_parametrizations()
    keys = set(m.state_dict())
    assert "conv_pre.weight" in keys and not any("parametrizations" in k for k in keys)
    assert torch.allclose(m.conv_pre.weight, w_before, atol=1e-7)


def test_vocos_huge_yaml_kwarg_spelling_is_accepted():
    # vocos-huge.yaml:9 passes kernel_sizes=[7] (TypeError in the reference, SURVEY 8b defect 2)
    enc = ConvNeXtEncoder(input_channels=8, depths=[1], dims=[16], kernel_sizes=[7])
    assert enc.kernel_size == 7
    with pytest.raises(ValueError):
        ConvNeXtEncoder(input_channels=8, depths=[1], dims=[16], kernel_sizes=[3, 7])


def test_hop_length_assertion_matches_reference():
    with pytest.raises(AssertionError):
        HiFiGANGenerator(hop_length=256)  # default rates multiply to 512 (hifigan.py:154-156)


def test_forward_refuses_cpu_tensors():
    kwargs, sd, ins, _, _ = load_golden("hifigan_small_ref")
    m = HiFiGANGenerator(**kwargs).eval()
    with pytest.raises(cabi.FvError):
        m(ins["mel"])
    head = ISTFTHead(dim=8, n_fft=16, hop_length=4, win_length=16)
    with pytest.raises(cabi.FvError):
        head(torch.zeros(1, 8, 3))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fv_vocoder.h")).read()
    declared = sorted(set(re.findall(r"^FV_API [^\n(]*?(fv_\w+)\(", header, flags=re.M)))
    assert declared == sorted(cabi.EXPORTS)
    assert os.path.exists(cabi.LIB_PATH), "libfv_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert cabi.lib().fv_abi_version() == 3


def test_conv_desc_struct_matches_header_layout():
    # field order of the ctypes mirror == field order in the header
    header = open(os.path.join(ROOT, "include", "fv_vocoder.h")).read()
    start = header.index("typedef struct fv_conv_desc {") + len("typedef struct fv_conv_desc {")
    body = header[start:header.index("} fv_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            tok = re.findall(r"[\w]+", part)
            if tok:
                names.append(tok[-1])
    assert names == [f[0] for f in cabi.ConvDesc._fields_]


def test_argument_validation_returns_error_not_crash():
    lib = cabi.lib()
    assert lib.fv_conv1d(None, 0, None) == -1
    assert b"null" in lib.fv_last_error()
    assert lib.fv_pack_input(None, None, 1, 1, 1, 8, 0, None) == -1


def _ref_from_pack(a, pc, L_out):
    """fp64 evaluation of the fv_conv1d contract straight from a PackedConv (what the kernels compute)."""
    B, L_in, ap = a.shape
    kmax = min(ap, pc.w_pitch)
    out = torch.zeros(B, L_out, pc.c_out, dtype=torch.float64)
    W = pc.w.double()
    for ph in range(pc.n_phase):
        rows = torch.arange(ph, L_out, pc.n_phase)
        q = rows // pc.n_phase
        for tp in range(pc.n_taps):
            src = q + pc.tap_off[ph * pc.n_taps + tp]
            ok = (src >= 0) & (src < L_in)
            out[:, rows[ok]] += a[:, src[ok], :kmax].double() @ W[ph, tp, :pc.c_out, :kmax].t()
    return out + pc.bias.double()


@pytest.mark.parametrize("k,u", [(16, 8), (4, 2), (8, 2), (2, 2), (11, 5), (10, 5), (8, 4)])
def test_pack_conv_transpose_polyphase(k, u):
    torch.manual_seed(0)
    ci, co, L = 12, 20, 9
    a = torch.zeros(2, L, cabi.pitch_of(ci))
    a[..., :ci] = torch.randn(2, L, ci)
    w, b = torch.randn(ci, co, k), torch.randn(co)
    pc = cabi.pack_conv_transpose(w, b, u)
    L_out = cabi.conv_transpose_out_len(L, k, u)
    ref = F.conv_transpose1d(a[..., :ci].permute(0, 2, 1).double(), w.half().double(), b.double(), stride=u,
                             padding=(k - u) // 2).permute(0, 2, 1)
    assert ref.shape[1] == L_out
    assert float((_ref_from_pack(a, pc, L_out) - ref).abs().max()) < 1e-9


@pytest.mark.parametrize("k,d", [(1, 1), (3, 1), (7, 3), (11, 5), (13, 1)])
def test_pack_conv_same_padding(k, d):
    torch.manual_seed(0)
    ci, co, L = 12, 20, 40
    a = torch.zeros(2, L, cabi.pitch_of(ci))
    a[..., :ci] = torch.randn(2, L, ci)
    w, b = torch.randn(co, ci, k), torch.randn(co)
    pc = cabi.pack_conv(w, b, d)
    ref = F.conv1d(a[..., :ci].permute(0, 2, 1).double(), w.half().double(), b.double(), dilation=d,
                   padding=(k * d - d) // 2).permute(0, 2, 1)
    assert float((_ref_from_pack(a, pc, L) - ref).abs().max()) < 1e-9
    assert pc.c_out_pad % 16 == 0 and pc.w_pitch % 8 == 0


def test_compat_package_provides_reference_dotted_paths():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for mod, name in [("fish_vocoder.modules.generators.hifigan", "HiFiGANGenerator"),
                          ("fish_vocoder.modules.generators.bigvgan", "BigVGANGenerator"),
                          ("fish_vocoder.modules.generators.unify", "UnifyGenerator"),
                          ("fish_vocoder.modules.generators.vocos", "ISTFTHead"),
                          ("fish_vocoder.modules.encoders.convnext", "ConvNeXtEncoder")]:
            cls = getattr(importlib.import_module(mod), name)
            assert cls.__module__.startswith("vocoder_b200.")
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for k in [k for k in sys.modules if k == "fish_vocoder" or k.startswith("fish_vocoder.")]:
            del sys.modules[k]
