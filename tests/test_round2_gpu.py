"""GPU parity tests added in round 2 (run with -m gpu on the B200 box; every call goes through the C ABI).

Tolerances: waveforms max|delta| <= 1e-3 * max(1, peak) against the fp32 CPU oracle (north_star); single kernels with
fp16 outputs 2^-10 relative to the output scale.  No trained checkpoint is available offline: "ref-init" = the reference
constructor's own initialisation, "stress" = SURVEY 8d.
"""
import math

import pytest
import torch

from oracle import generators as G
from tests.util import build_module, load_golden, oracle_forward, stress_init
from vocoder_b200 import cabi

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _set_precision(m, mode):
    for sub in m.modules():
        if hasattr(sub, "_ws"):
            sub.precision = mode
    m.precision = mode


# ------------------------------------------------------------------------------------------------
# the exact benched shapes (BASELINE cfg B / C / D) with stress weights, B = 2
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("workload,weights", [("hifigan_b64", "stress"), ("hifigan_b64", "ref-init"),
                                              ("bigvgan_b32", "stress"), ("bigvgan_b32", "ref-init"),
                                              ("vocos_huge_b128", "stress"), ("vocos_huge_b128", "ref-init")])
def test_benched_shape_parity_vs_oracle(workload, weights):
    """The modules bench.py times (full width, full depth: Vocos [3,3,27,3]), at the benched T, in their default precision,
    with stress weights so that every residual branch carries signal."""
    import bench
    kind, _, n_mels, T, hop, sr, _ = bench.WORKLOADS[workload]
    model = bench.build_model(kind).eval()
    if weights == "stress":
        stress_init(model, seed=1)
    mel = bench.synthetic_mel(2, n_mels, T, 1234)
    sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = bench.oracle_forward(kind, sd, mel, model)
        y = model.cuda()(mel.cuda()).cpu()
    assert y.shape == want.shape == (2, 1, T * hop)
    peak = max(1.0, float(want.abs().max()))
    err = float((y - want).abs().max())
    print(f"{workload} [{weights}, default precision]: max|delta| vs fp32 oracle {err:.3e} (peak {peak:.3f})")
    assert err <= TOL * peak, f"{workload}/{weights}: max|delta|={err:.3e} (peak {peak:.3f})"


# ------------------------------------------------------------------------------------------------
# anti-aliased activation: edge modes, plain Snake, non-logscale parameters (bigvgan.py:60-71,122-135)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["replicate", "reflect", "zero"])
@pytest.mark.parametrize("kind,logscale", [("snakebeta", True), ("snake", True), ("snake", False), ("snakebeta", False)])
def test_snake_edge_modes(mode, kind, logscale):
    torch.manual_seed(hash((mode, kind, logscale)) % 1000)
    f = G.kaiser_sinc_taps()
    for B, L, C in ((2, 6, 8), (1, 17, 24), (2, 301, 16)):
        x = torch.randn(B, C, L)
        alpha = torch.randn(C) * 0.5 if logscale else 0.5 + torch.rand(C)
        beta = (torch.randn(C) * 0.5 if logscale else 0.5 + torch.rand(C)) if kind == "snakebeta" else None
        act = (lambda v: G.snake_beta(v, alpha, beta, logscale)) if kind == "snakebeta" else (
            lambda v: G.snake(v, alpha, logscale))
        want = G.aa_activation(x, act, f, f, edge_mode=mode).permute(0, 2, 1)       # [B, L, C]
        x_cl = x.permute(0, 2, 1).contiguous().cuda()
        out16 = torch.full((B, L, C), float("nan"), dtype=torch.float16, device="cuda")
        cabi.snake_aa(x_cl, out16, alpha.cuda(), None if beta is None else beta.cuda(), f.tolist(), f.tolist(), C,
                      logscale=logscale, split=0, edge_mode=mode)
        got = out16.float().cpu()
        scale = max(1.0, float(want.abs().max()))
        err = float((got - want).abs().max())
        assert err <= 2.0 ** -10 * scale, f"{mode}/{kind}/logscale={logscale} L={L}: {err:.3e} (scale {scale:.2f})"


def test_bigvgan_edge_mode_generator_level():
    """BigVGANGenerator.aa_edge_mode reaches every anti-aliased activation: the waveform follows the oracle evaluated with
    the same mode and differs from the default mode only within the receptive field of the two sequence ends."""
    kwargs, sd, ins, out, extra = load_golden("bigvgan_small_stress")
    m = build_module("bigvgan_small_stress", kwargs)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    _set_precision(m, "strict")
    mel = torch.empty(1, kwargs["num_mels"], 160).uniform_(-11.5129, 2.0, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        base = m(mel.cuda()).cpu()
        for mode in ("reflect", "zero"):
            m.aa_edge_mode = mode
            y = m(mel.cuda()).cpu()
            want = G.bigvgan_forward(sd, mel, kwargs["upsample_rates"], kwargs["resblock_dilation_sizes"], edge_mode=mode)
            assert float((y - want).abs().max()) <= 2e-4, mode
            assert float((y - base).abs().max()) > 1e-4          # the mode matters at the edges ...
            mid = y.shape[-1] // 2
            assert float((y - base)[..., mid - 256:mid + 256].abs().max()) <= 1e-5   # ... and only there


# ------------------------------------------------------------------------------------------------
# upstream-Vocos layout (scripts/vocos_gen.py:5-16)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("padding", ["center", "same"])
def test_upstream_vocos_backbone_and_head(padding):
    from vocoder_b200.encoders import VocosBackbone
    from vocoder_b200.generators import ISTFTHead, UnifyGenerator
    torch.manual_seed(0)
    bb = VocosBackbone(input_channels=20, dim=64, intermediate_dim=192, num_layers=3)
    head = ISTFTHead(dim=64, n_fft=64, hop_length=16, padding=padding, upstream_layout=True)
    m = UnifyGenerator(backbone=bb, head=head).eval()
    sd = m.state_dict()
    g = torch.Generator().manual_seed(5)
    for k in sd:  # trunc_normal(0.02) leaves the head silent: scale weights so magnitudes and phases move
        if k.endswith("weight") and sd[k].ndim >= 2:
            sd[k] = sd[k] * 6.0
        if k.endswith(".gamma"):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.45 + 0.05
    m.load_state_dict(sd)
    mel = torch.randn(2, 20, 14)
    with torch.no_grad():
        want = G.upstream_vocos_forward({k: v.float() for k, v in sd.items()}, mel, 64, 16, padding)
        y = m.cuda()(mel.cuda()).cpu()
        feats = bb(mel.cuda()).cpu()                                 # [B, T, dim], as upstream hands it to the head
        feats_want = G.vocos_backbone_forward({k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")},
                                              mel).transpose(1, 2)
        y_head = head(feats.cuda()).cpu()[:, None, :]
    L = (14 - 1) * 16 if padding == "center" else 14 * 16
    assert y.shape == want.shape == (2, 1, L)
    peak = max(1.0, float(want.abs().max()))
    assert float((y - want).abs().max()) <= TOL * peak
    assert float((feats - feats_want).abs().max()) <= 2e-3 * max(1.0, float(feats_want.abs().max()))
    assert float((y_head - want).abs().max()) <= 2 * TOL * peak


# ------------------------------------------------------------------------------------------------
# fp16 range robustness (fp16 has a 5-bit exponent; operands saturate at 65504 instead of overflowing)
# ------------------------------------------------------------------------------------------------
def _small_hifigan(seed=0):
    from vocoder_b200.generators import HiFiGANGenerator
    torch.manual_seed(seed)
    m = HiFiGANGenerator(hop_length=32, upsample_rates=[4, 4, 2], upsample_kernel_sizes=[8, 8, 4], num_mels=20,
                         upsample_initial_channel=64, use_template=False)
    stress_init(m, seed=2)
    return m.eval()


def _scale_conv(m, name, factor):
    sd = m.state_dict()
    k0 = f"{name}.parametrizations.weight.original0"
    sd[k0] = sd[k0] * factor                          # weight-norm gain g: scales the effective weight
    sd[f"{name}.bias"] = sd[f"{name}.bias"] * factor
    m.load_state_dict(sd)


@pytest.mark.parametrize("fuse", [True, False])
def test_large_activations_up_to_the_fp16_range_stay_within_tolerance(fuse):
    """Intermediate activations of 1e3 .. 2e4 in every stage (conv_pre gain x 1000, undone by conv_post / 1000 so the
    waveform is not simply tanh-saturated): fp16 keeps its 11-bit relative precision up to 65504, so the waveform still
    matches the fp32 oracle within 1e-3."""
    m = _small_hifigan()
    _scale_conv(m, "conv_pre", 1000.0)
    _scale_conv(m, "conv_post", 1.0e-3)
    mel = torch.empty(2, 20, 13).uniform_(-11.5129, 2.0, generator=torch.Generator().manual_seed(1))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        pre = G.conv_same(sd, "conv_pre", mel)
        want = G.hifigan_forward(sd, mel, m.upsample_rates)
    assert 5e3 < float(pre.abs().max()) < 6e4 and 0.05 < float(want.abs().max()) < 0.999
    m = m.cuda()
    m.fuse_mrf = fuse
    with torch.no_grad():
        y = m(mel.cuda()).cpu()
    rep = m.saturation_report()
    assert bool(torch.isfinite(y).all())
    assert rep["saturated"] == 0, rep
    assert float((y - want).abs().max()) <= TOL


def test_activations_beyond_the_fp16_range_saturate_finite_and_are_reported():
    m = _small_hifigan()
    _scale_conv(m, "conv_pre", 1.0e5)
    mel = torch.empty(1, 20, 13).uniform_(-11.5129, 2.0, generator=torch.Generator().manual_seed(1))
    m = m.cuda()
    m.fuse_mrf = False
    with torch.no_grad():
        y = m(mel.cuda())
    rep = m.saturation_report()
    assert bool(torch.isfinite(y).all()) and float(y.abs().max()) <= 1.0     # never inf / NaN: operands clamp at 65504
    assert rep["saturated"] > 0 and rep["buffers"]                            # ... and the clamp is visible to the caller


@pytest.mark.parametrize("fuse", [True, False])
def test_tiny_weights_are_rescaled_per_output_channel(fuse):
    """Residual-block weights of ~1e-6 sit in fp16's subnormal range (spacing 6e-8 = 6% of the value).  The packer folds a
    power-of-two per-output-channel scale into the fp16 weights and undoes it in the fp32 epilogue, so the result is as
    accurate as for O(1) weights."""
    m = _small_hifigan()
    sd = m.state_dict()
    for k in list(sd):
        if ".convs2." in k and k.endswith("original0"):
            sd[k] = sd[k] * 1e-6
        if k.startswith("ups.") and k.endswith("original0"):
            sd[k] = sd[k] * 2e-5
    m.load_state_dict(sd)
    mel = torch.empty(2, 20, 13).uniform_(-11.5129, 2.0, generator=torch.Generator().manual_seed(3))
    sdc = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        want = G.hifigan_forward(sdc, mel, m.upsample_rates)
    m = m.cuda()
    m.fuse_mrf = fuse
    with torch.no_grad():
        y = m(mel.cuda()).cpu()
    peak = float(want.abs().max())
    err = float((y - want).abs().max())
    print(f"tiny weights [fuse={fuse}]: max|delta| {err:.3e} at waveform peak {peak:.3e}")
    assert err <= 2e-3 * peak + 1e-7, f"{err:.3e} vs peak {peak:.3e}"


# ------------------------------------------------------------------------------------------------
# runtime: graph invalidation, bounded workspace, device guard
# ------------------------------------------------------------------------------------------------
def test_unify_graph_is_recaptured_after_a_repack():
    """a UnifyGenerator graph must not replay pointers into packed weights its sub-modules have since replaced."""
    name = "vocos_small_stress"
    kwargs, sd, ins, out, extra = load_golden(name)
    m = build_module(name, kwargs)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    mel = ins["mel"].cuda()
    with torch.no_grad():
        m.use_cuda_graph = True
        y0 = m(mel).clone()
        sd2 = {k: (v * 1.25 if k.endswith("pwconv2.weight") else v) for k, v in sd.items()}
        m.load_state_dict(sd2)                                   # in-place copy: sub-modules repack on the next forward
        y1 = m(mel).clone()
        m.remove_parametrizations()
        y2 = m(mel).clone()
        m.use_cuda_graph = False
        e1 = m(mel).clone()
    assert float((y0 - y1).abs().max()) > 1e-3                   # the new weights took effect under graph replay
    assert torch.equal(y1, e1) and torch.equal(y2, e1)


def test_workspace_and_graphs_stay_bounded_over_many_shapes():
    from vocoder_b200 import inference as inf
    m = _small_hifigan().cuda()
    m.use_cuda_graph = True
    sizes = []
    with torch.no_grad():
        for T in range(9, 33):
            m(torch.randn(1, 20, T, device="cuda"))
            sizes.append(m._ws.nbytes())
    assert m._ws.n_signatures() <= m._ws.max_signatures
    assert len(m._graphed._graphs) <= m._graphed.max_graphs
    assert max(sizes[8:]) <= 4 * max(sizes[:4]) * (32 / 9)       # bounded by the few largest shapes, not by their number
    # eager result after all the evictions still equals a fresh module's
    x = torch.randn(1, 20, 21, device="cuda")
    with torch.no_grad():
        y = m(x).clone()
        m.use_cuda_graph = False
        assert torch.equal(y, m(x))
        # chunked synthesis with a template: three chunk shapes at most
        kwargs, sd, ins, out, extra = load_golden("hifigan_template_stress")
        g = build_module("hifigan_template_stress", kwargs)
        g.load_state_dict(sd)
        g = g.eval().cuda()
        mel = torch.empty(1, 12, 160, device="cuda").uniform_(-11.5, 2.0)
        tpl = torch.randn(1, 1, 160 * 20, device="cuda") * 0.3
        full = g(mel, tpl).clone()
        got = inf.chunked_forward(g, mel, chunk_frames=32, template=tpl)
    assert float((got - full).abs().max()) <= 1e-6
    assert g._ws.n_signatures() <= g._ws.max_signatures


def test_pointer_of_another_device_is_refused():
    if torch.cuda.device_count() < 2:
        x = torch.zeros(1, 8, 8, device="cuda")
        assert cabi._ptr(x) == x.data_ptr()
        pytest.skip("single-GPU box: the cross-device refusal needs two devices")
    x = torch.zeros(1, 8, 8, device="cuda:1")
    with torch.cuda.device(0):
        with pytest.raises(cabi.FvError):
            cabi._ptr(x)
    m = _small_hifigan().to("cuda:1")
    with torch.no_grad():
        y = m(torch.randn(1, 20, 9, device="cuda:1"))           # forward enters the tensor's device itself
    assert y.device.index == 1 and bool(torch.isfinite(y).all())


# ------------------------------------------------------------------------------------------------
# row-pair convs of the narrow Snake stages (cabi.pack_conv_row_pairs)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [16, 8])
@pytest.mark.parametrize("k,d", [(3, 1), (3, 5), (7, 3), (11, 1), (11, 5)])
def test_row_pair_conv_matches_the_plain_conv(C, k, d):
    """Same weights, same buffers: fv_conv1d on the [B, L/2, 2C] view with pair taps == fv_conv1d on [B, L, C] (residual,
    running-sum accumulate and out_scale included), and both match an fp32 torch conv on the fp16-rounded operands."""
    import torch.nn.functional as F
    torch.manual_seed(100 * C + 10 * k + d)
    B, L = 3, 2 * 1187                      # ragged against every tile size
    w = (torch.randn(C, C, k) * 0.3).cuda()
    bias = torch.randn(C).cuda()
    a16 = torch.randn(B, L, C, device="cuda").half()
    res = torch.randn(B, L, C, device="cuda")
    acc0 = torch.randn(B, L, C, device="cuda")
    with cabi.precision("fp16"):
        plain, pairs = cabi.pack_conv(w, bias, d), cabi.pack_conv_row_pairs(w, bias, d)
        o_plain, o_pairs = acc0.clone(), acc0.clone()
        cabi.conv1d(a16, plain, residual=res, out32=o_plain, accumulate=True, out_scale=0.5)
        cabi.conv1d_row_pairs(a16, pairs, residual=res, out32=o_pairs, accumulate=True, out_scale=0.5)
        t_plain, t_pairs = torch.empty_like(acc0), torch.empty_like(acc0)
        cabi.conv1d(a16, plain, out32=t_plain)
        cabi.conv1d_row_pairs(a16, pairs, out32=t_pairs)
    torch.cuda.synchronize()
    ref = F.conv1d(a16.float().permute(0, 2, 1), w.half().float(), bias, padding=(k * d - d) // 2, dilation=d).permute(0, 2, 1)
    scale = float(ref.abs().max())
    assert float((t_pairs - ref).abs().max()) <= 2e-5 * scale + 1e-5
    assert float((t_pairs - t_plain).abs().max()) <= 2e-5 * scale + 1e-5      # fp32 sums in a different order
    want = (ref + res) * 0.5 + acc0
    assert float((o_pairs - want).abs().max()) <= 2e-5 * scale + 1e-5
    assert float((o_pairs - o_plain).abs().max()) <= 2e-5 * scale + 1e-5


def test_bigvgan_row_pair_stage_equals_plain_stage():
    """Generator level: BigVGAN with a 16-channel last stage, conv_row_pairs on / off, same weights: identical up to fp32
    summation order, and within tolerance of the oracle."""
    from vocoder_b200.generators import BigVGANGenerator
    torch.manual_seed(5)
    kw = dict(hop_length=16, upsample_rates=[4, 2, 2], upsample_kernel_sizes=[8, 4, 4], num_mels=20,
              upsample_initial_channel=128, use_template=False)
    m = BigVGANGenerator(**kw).eval()
    stress_init(m, seed=3)
    mel = torch.randn(2, 20, 37)
    sd = {k_: v.detach().cpu().float() for k_, v in m.state_dict().items()}
    with torch.no_grad():
        want = G.bigvgan_forward(sd, mel, kw["upsample_rates"])
        m = m.cuda()
        launches0 = cabi.launch_count()
        y_pairs = m(mel.cuda()).cpu()
        n_pairs = cabi.launch_count() - launches0
        m.conv_row_pairs = False
        y_plain = m(mel.cuda()).cpu()
    assert m._packed["blocks"][-1][0][3] is None        # repacked without the pair taps
    peak = max(1.0, float(want.abs().max()))
    assert float((y_pairs - want).abs().max()) <= TOL * peak
    assert float((y_pairs - y_plain).abs().max()) <= 1e-4 * peak
    assert n_pairs > 0


# ------------------------------------------------------------------------------------------------
# short sequences: kernel-size chains on two streams (MRFGeneratorBase.chain_streams)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["hifigan", "bigvgan"])
@pytest.mark.parametrize("graph", [False, True])
def test_chain_streams_are_bit_identical_to_the_serial_order(kind, graph):
    """B = 1, T = 94 (test.py's shape): every stage type (layer-wise C = 256, pair-wise C = 128, whole-stage fused C = 64 / 32;
    BigVGAN: Snake stages incl. row-pair convs) with the largest kernel-size chain on the side stream == the serial schedule,
    eagerly and under CUDA-graph replay (fork / join captured), repeated to catch ordering races."""
    import bench
    model = bench.build_model(kind).eval()
    stress_init(model, seed=2)
    n_mels = model.num_mels
    mel = bench.synthetic_mel(1, n_mels, 94, 77).cuda()
    model = model.cuda()
    with torch.no_grad():
        model.chain_streams, model.use_cuda_graph = False, False
        serial = model(mel).clone()
        model.chain_streams, model.use_cuda_graph = True, graph
        n0 = cabi.launch_count()
        outs = [model(mel).clone() for _ in range(5)]
        assert cabi.launch_count() > n0 or graph
    torch.cuda.synchronize()
    for y in outs:
        assert torch.equal(y, serial), float((y - serial).abs().max())
    sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = bench.oracle_forward(kind, sd, mel.cpu(), model)
    assert float((outs[-1].cpu() - want).abs().max()) <= TOL * max(1.0, float(want.abs().max()))
