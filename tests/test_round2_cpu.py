"""CPU tests added in round 2: oracle known-answer tests for the new paths (ISTFT "center", anti-alias edge modes,
upstream-Vocos restatement) and the host-side runtime logic (workspace LRU, graph tags, precision modes, Hydra stand-in)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import generators as G
from vocoder_b200 import cabi
from vocoder_b200.runtime import Workspace


# ------------------------------------------------------------------------------------------------
# oracle KATs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_fft,hop,win", [(64, 16, 64), (64, 16, 48), (1024, 256, 1024)])
def test_istft_center_restatement_matches_torch_istft(n_fft, hop, win):
    """oracle.istft_center == torch.istft(center=True) for a one-sided spectrum, a two-sided one (what the reference head
    hands over: ATen keeps rows 0..n_fft/2) and a window shorter than n_fft (centred zero padding)."""
    torch.manual_seed(0)
    T = 9
    window = torch.hann_window(win)
    one = torch.randn(2, n_fft // 2 + 1, T, dtype=torch.complex64)
    two = torch.randn(2, n_fft, T, dtype=torch.complex64)
    for spec in (one, two):
        want = torch.istft(spec, n_fft, hop, win, window, center=True)
        got = G.istft_center(spec, n_fft, hop, win, window)
        assert got.shape == want.shape == (2, (T - 1) * hop)
        assert float((got - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))


def test_istft_center_inverts_torch_stft():
    torch.manual_seed(1)
    n_fft, hop = 64, 16
    y = torch.randn(2, 20 * hop)
    window = torch.hann_window(n_fft)
    spec = torch.stft(y, n_fft, hop, n_fft, window, center=True, return_complex=True)
    rec = G.istft_center(spec, n_fft, hop, n_fft, window)
    assert rec.shape == y.shape and float((rec - y).abs().max()) < 2e-5


@pytest.mark.parametrize("mode", ["replicate", "reflect", "zero"])
def test_aa_edge_modes_explicit_formula(mode):
    """The direct formula fv_snake_aa's edge-mode kernel evaluates (v~ = padded activated 2x signal, x~ = padded input)
    == the conv formulation with F.pad(mode) on both filters; interior samples do not depend on the mode."""
    torch.manual_seed(0)
    f = G.kaiser_sinc_taps().double()
    a, b = 1.3, 0.7
    act = lambda u: u + torch.sin(a * u) ** 2 / (b + 1e-9)

    def edge(i, n):
        if 0 <= i < n:
            return i
        if mode == "zero":
            return None
        if mode == "reflect":
            return -i if i < 0 else 2 * (n - 1) - i
        return 0 if i < 0 else n - 1

    for L in (6, 9, 23):
        x = torch.randn(L).double()

        def xv(i):
            j = edge(i, L)
            return 0.0 if j is None else x[j]

        v = torch.zeros(2 * L).double()
        for n in range(2 * L):
            even = 1 - (n & 1)
            base = (n + 5 - even) // 2
            v[n] = act(2 * sum(f[2 * q + even] * xv(base - q) for q in range(6)))

        def vv(n):
            j = edge(n, 2 * L)
            return 0.0 if j is None else v[j]

        out = torch.tensor([float(sum(f[j] * vv(2 * t - 5 + j) for j in range(12))) for t in range(L)])
        ref = G.aa_activation(x.float().view(1, 1, L), act, f.float(), f.float(), edge_mode=mode).view(-1)
        assert float((ref.double() - out.double()).abs().max()) < 2e-6
        rep = G.aa_activation(x.float().view(1, 1, L), act, f.float(), f.float()).view(-1)
        if L > 22:  # +-5 input samples of support per filter stage: samples >= 11 from either end see no padding
            assert float((ref[11:-11] - rep[11:-11]).abs().max()) == 0.0


def test_upstream_vocos_restatement_shapes_and_head_layouts():
    """vocos==0.0.2 (third-party, unpinned): VocosBackbone -> ISTFTHead(n_fft + 2 outputs).  The one-sided upstream head
    and the reference head restricted to its live rows are the same function."""
    from vocoder_b200.encoders import VocosBackbone
    from vocoder_b200.generators import ISTFTHead
    torch.manual_seed(0)
    bb = VocosBackbone(input_channels=20, dim=32, intermediate_dim=96, num_layers=2)
    head = ISTFTHead(dim=32, n_fft=64, hop_length=16, padding="center", upstream_layout=True)
    sd = {"backbone." + k: v for k, v in bb.state_dict().items()}
    sd.update({"head." + k: v for k, v in head.state_dict().items()})
    assert sd["head.out.weight"].shape == (66, 32) and sd["backbone.embed.weight"].shape == (32, 20, 7)
    mel = torch.randn(2, 20, 12)
    with torch.no_grad():
        y = G.upstream_vocos_forward(sd, mel, 64, 16, "center")
        # same weights in the reference layout [2*n_fft, dim, 1]: rows beyond n_fft/2 of each chunk are dead
        w, b = sd["head.out.weight"], sd["head.out.bias"]
        w2 = torch.randn(128, 32, 1)
        b2 = torch.randn(128)
        w2[:33, :, 0], w2[64:97, :, 0], b2[:33], b2[64:97] = w[:33], w[33:], b[:33], b[33:]
        sd2 = dict(sd)
        sd2["head.out.weight"], sd2["head.out.bias"] = w2, b2
        x = G.vocos_backbone_forward(sd2, mel, "backbone.")
        y2 = G.istft_head_forward(sd2, x, 64, 16, 64, "head.", "center")[:, None, :]
    assert y.shape == (2, 1, 11 * 16)
    assert float((y - y2).abs().max()) < 1e-5


# ------------------------------------------------------------------------------------------------
# host runtime logic
# ------------------------------------------------------------------------------------------------
def test_workspace_keeps_a_bounded_number_of_signatures_and_notifies():
    ws = Workspace(max_signatures=2)
    dropped = []
    ws.add_listener(lambda: dropped.append(1))
    for T in (10, 20, 30, 40):
        ws.enter(("sig", T))
        a = ws.get("x", (2, T, 8), torch.float32, "cpu")
        assert ws.get("x", (2, T, 8), torch.float32, "cpu") is a      # stable address inside a signature
    assert ws.n_signatures() == 2 and len(dropped) == 2
    assert ws.nbytes() == (30 + 40) * 2 * 8 * 4
    ws.enter(("sig", 30))                                             # touching keeps it resident
    ws.enter(("sig", 50))
    ws.get("x", (2, 50, 8), torch.float32, "cpu")
    ws.enter(("sig", 30))
    assert ws.get("x", (2, 30, 8), torch.float32, "cpu").shape == (2, 30, 8) and len(dropped) == 3


def test_precision_modes_and_strict_layer_contexts():
    assert cabi.mode() == "fp16" and not cabi.is_strict()
    with cabi.precision("mixed"):
        assert cabi.is_mixed() and not cabi.is_strict() and cabi.pitch_of(352) == 352
        with cabi.strict_layer():
            assert cabi.is_strict() and cabi.is_mixed() and cabi.pitch_of(352) == 384 and cabi.f16_width(352) == 768
        with cabi.strict_layer(False):
            assert not cabi.is_strict()
        assert not cabi.is_strict()
    with cabi.precision("strict"):
        assert cabi.is_strict() and not cabi.is_mixed()
        with cabi.strict_layer(False):   # a plain layer cannot be carved out of a strict forward
            assert cabi.is_strict()
    with pytest.raises(ValueError):
        with cabi.precision("fp8"):
            pass
    assert cabi.DEFAULT_PRECISION in cabi.PRECISIONS


def test_mixed_mode_packs_only_the_sensitive_layers_strict():
    """BigVGAN trunk (conv_pre / ups) and the Vocos stem / downsample / head contractions carry [Whi | Whi | Wlo] weights
    in "mixed"; residual-block convs and the ConvNeXt pointwise GEMMs stay single fp16; HiFiGAN is untouched."""
    from vocoder_b200.encoders import ConvNeXtEncoder
    from vocoder_b200.generators import BigVGANGenerator, HiFiGANGenerator, ISTFTHead
    kw = dict(hop_length=8, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], num_mels=12, upsample_initial_channel=32,
              use_template=False)
    with cabi.precision("mixed"):
        P = BigVGANGenerator(**kw)._ensure_packed("cpu")
        assert P["pre"].split > 0 and all(u.split > 0 for u in P["ups"])
        assert all(c.split == 0 for blocks in P["blocks"] for c1, c2, _, _ in blocks for c in c1 + c2)
        H = HiFiGANGenerator(**kw)
        H.fuse_mrf = False
        PH = H._ensure_packed("cpu")
        assert PH["pre"].split == 0 and all(u.split == 0 for u in PH["ups"])
        E = ConvNeXtEncoder(input_channels=12, depths=[1, 1], dims=[16, 32])._ensure_packed("cpu")
        assert all(d[0].split > 0 for d in E["down"])
        assert all(b["pw1"].split == 0 and b["pw2"].split == 0 for st in E["stages"] for b in st)
        Ph = ISTFTHead(dim=32, n_fft=64, hop_length=16, win_length=64)._ensure_packed("cpu")
        assert Ph["head"].split > 0 and Ph["idft"].split > 0
    with cabi.precision("fp16"):
        P = BigVGANGenerator(**kw)._ensure_packed("cpu")
        assert P["pre"].split == 0


def test_repack_bumps_the_pack_generation():
    from vocoder_b200.generators import HiFiGANGenerator
    m = HiFiGANGenerator(hop_length=8, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], num_mels=12,
                         upsample_initial_channel=32, use_template=False)
    with cabi.precision("fp16"):
        m._ensure_packed("cpu")
        g0 = m._pack_gen
        m._ensure_packed("cpu")
        assert m._pack_gen == g0                  # cached
        m.load_state_dict(m.state_dict())          # in-place copy bumps tensor versions
        m._ensure_packed("cpu")
        assert m._pack_gen == g0 + 1
        m.remove_parametrizations()
        m._ensure_packed("cpu")
        assert m._pack_gen == g0 + 2


def test_forward_validates_channel_count_and_template_shape():
    from vocoder_b200.generators import HiFiGANGenerator
    from vocoder_b200.runtime import require_channels
    with pytest.raises(ValueError):
        require_channels(torch.zeros(1, 7, 5), 8, "x")
    require_channels(torch.zeros(1, 8, 5), 8, "x")


def test_instantiate_handles_partial_and_torch_targets():
    import functools

    from vocoder_b200 import inference as inf
    cfg = {"_target_": "fish_vocoder.modules.generators.hifigan.HiFiGANGenerator", "hop_length": 8,
           "upsample_rates": [4, 2], "upsample_kernel_sizes": [8, 4], "num_mels": 12, "upsample_initial_channel": 32,
           "use_template": False, "post_activation": {"_target_": "torch.nn.SiLU", "_partial_": True, "inplace": True}}
    gen = inf.instantiate(cfg)
    assert isinstance(gen.activation_post, torch.nn.SiLU)
    part = inf.instantiate({"_target_": "torch.nn.LeakyReLU", "_partial_": True, "negative_slope": 0.2})
    assert isinstance(part, functools.partial) and part().negative_slope == 0.2
    with pytest.raises(KeyError):
        inf.instantiate({"_target_": "os.system", "command": "true"})


def test_load_mel_refuses_pickled_objects(tmp_path):
    from vocoder_b200 import inference as inf

    class Evil:
        def __reduce__(self):
            return (print, ("pwned",))

    torch.save(Evil(), tmp_path / "x.pt")
    with pytest.raises(Exception):
        inf.load_mel(str(tmp_path / "x.pt"))
    torch.save(torch.zeros(3, 4), tmp_path / "ok.pt")
    assert inf.load_mel(str(tmp_path / "ok.pt")).shape == (1, 3, 4)


def test_abi_header_declares_version_3_and_edge_modes():
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "fv_vocoder.h")).read()
    assert "#define FV_ABI_VERSION 3" in hdr
    assert cabi.lib().fv_abi_version() == 3
    for name, val in cabi.EDGE_MODES.items():
        assert re.search(rf"FV_EDGE_{name.upper()} = {val}\b", hdr)
    # every FV_API declaration is in EXPORTS and vice versa
    declared = set(re.findall(r"FV_API\s+[\w\s\*]+?\b(fv_\w+)\s*\(", hdr))
    assert declared == set(cabi.EXPORTS)


@pytest.mark.parametrize("k,d", [(3, 1), (3, 3), (3, 5), (7, 1), (7, 3), (7, 5), (11, 1), (11, 3), (11, 5)])
def test_row_pair_packing_is_the_same_conv(k, d):
    """cabi.pack_conv_row_pairs: the [B, L/2, 2C] view of a channels-last buffer convolved with the 2C x 2C pair taps equals
    the original "same"-padded Conv1d (hifigan.py:21-22 padding rule), for every (kernel, dilation) of an AMP block."""
    torch.manual_seed(k * 10 + d)
    C, L, B = 16, 46, 2
    w = torch.randn(C, C, k) * 0.2
    b = torch.randn(C)
    x = torch.randn(B, C, L)
    with cabi.precision("fp16"):
        pc = cabi.pack_conv_row_pairs(w, b, d)
    assert pc.c_in == 2 * C and pc.c_out == 2 * C and pc.n_phase == 1 and pc.n_taps <= cabi.MAX_TAPS
    assert pc.n_taps == len(set((p + (j - (k - 1) // 2) * d) // 2 for p in (0, 1) for j in range(k)))
    ref = F.conv1d(x, w.half().float(), b, padding=(k * d - d) // 2, dilation=d)           # [B, C, L]
    a2 = x.permute(0, 2, 1).reshape(B, L // 2, 2 * C)                                          # the row-pair view
    out2 = torch.zeros(B, L // 2, 2 * C)
    W = pc.w.float()                                                                           # [1, taps, C_out_pad, w_pitch]
    for i, off in enumerate(pc.tap_off):
        sh = torch.zeros_like(a2)
        lo, hi = max(0, -off), min(L // 2, L // 2 - off)
        if hi > lo:
            sh[:, lo:hi] = a2[:, lo + off:hi + off]                                            # rows outside [0, L/2) are zero
        out2 += sh @ W[0, i, :2 * C, :2 * C].t()
    scale = 1.0 if pc.w_scale is None else pc.w_scale
    out2 = (out2 + pc.bias) * scale
    got = out2.reshape(B, L, C).permute(0, 2, 1)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())


def test_snake_stage_packs_row_pairs_only_for_narrow_unpadded_channels():
    from vocoder_b200.generators import BigVGANGenerator
    kw = dict(hop_length=8, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], num_mels=12, upsample_initial_channel=64,
              use_template=False)
    with cabi.precision("mixed"):
        g = BigVGANGenerator(**kw)                       # stages of 32 and 16 channels
        P = g._ensure_packed("cpu")
        assert all(bp is None for _, _, _, bp in P["blocks"][0])
        assert all(bp is not None and bp[0][0].c_out == 32 for _, _, _, bp in P["blocks"][1])
        g.conv_row_pairs = False
        assert all(bp is None for blocks in g._ensure_packed("cpu")["blocks"] for _, _, _, bp in blocks)
    with cabi.precision("strict"):
        assert all(bp is None for blocks in BigVGANGenerator(**kw)._ensure_packed("cpu")["blocks"] for _, _, _, bp in blocks)
    assert cabi.row_pairs_ok(16, 10) and not cabi.row_pairs_ok(16, 11) and not cabi.row_pairs_ok(12, 10)


def test_chain_streams_auto_threshold():
    from vocoder_b200.generators import HiFiGANGenerator
    m = HiFiGANGenerator()
    assert m._chains_concurrent(1, 24064) and m._chains_concurrent(2, 12032) and not m._chains_concurrent(64, 752 * 8)
    m.chain_streams = False
    assert not m._chains_concurrent(1, 10)


def test_module_params_key_sees_every_kind_of_weight_change():
    """runtime.module_params_key (the per-forward fingerprint that keeps packed weights and CUDA graphs valid) reads cached
    parameter slots instead of walking the module tree: it must still change on in-place updates, dtype / device style
    `.data` swaps, slot replacement (load_state_dict(assign=True)), removed parametrizations and added parameters."""
    from vocoder_b200.generators import HiFiGANGenerator
    from vocoder_b200.runtime import module_params_key
    m = HiFiGANGenerator(hop_length=8, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], num_mels=12,
                         upsample_initial_channel=32, use_template=False).eval()
    k0 = module_params_key(m)
    assert module_params_key(m) == k0                                   # stable while nothing changes
    with torch.no_grad():
        m.conv_pre.bias.add_(1.0)
    k1 = module_params_key(m)
    assert k1 != k0                                                     # in-place update (_version)
    m.conv_post.bias.data = m.conv_post.bias.data.clone()
    k2 = module_params_key(m)
    assert k2 != k1                                                     # .data swapped (what .to() / .half() do)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict(sd, assign=True)
    k3 = module_params_key(m)
    assert k3 != k2                                                     # tensors replaced in their slots
    torch.nn.utils.parametrize.remove_parametrizations(m.ups[0], "weight")   # directly on a sub-module, not our method
    k4 = module_params_key(m)
    assert k4 != k3 and len(k4[3]) == len(k3[3]) - 1                    # tree changed: slots rebuilt (g, v -> weight)
    m.register_buffer("extra", torch.zeros(1))
    assert len(module_params_key(m)[3]) == len(k4[3]) + 1               # new slot
    with cabi.precision("strict"):
        assert module_params_key(m)[:2] != k4[:2]                       # precision mode is part of the key
