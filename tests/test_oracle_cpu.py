"""Pin the CPU oracle (oracle/generators.py) against the reference-generated golden vectors and
the self-derived known-answer tests of SURVEY.md section 8c."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import ALL_GOLDEN, load_golden, oracle_forward
from oracle import generators as G


@pytest.mark.parametrize("name", ALL_GOLDEN)
def test_oracle_matches_reference_golden(name):
    kwargs, sd, ins, out, extra = load_golden(name)
    with torch.no_grad():
        y = oracle_forward(name, kwargs, sd, ins, extra)
    assert y.shape == out.shape
    scale = max(1.0, float(out.abs().max()))
    assert float((y - out).abs().max()) <= 2e-6 * scale


def test_kaiser_sinc_taps_known_answer():
    # 12 taps (symmetric), sum 1
    f = G.kaiser_sinc_taps()
    want = [0.0020289647, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832, 0.4432097971]
    assert torch.allclose(f[:6], torch.tensor(want), atol=2e-8)
    assert torch.allclose(f, f.flip(0), atol=1e-9)
    assert abs(float(f.sum()) - 1.0) < 2e-7


def test_aa_activation_identity_impulse():
    # KAT 3: Activation1d(identity) on a unit impulse has support exactly [-5, +5] and sum 1
    f = G.kaiser_sinc_taps()
    x = torch.zeros(1, 1, 41)
    x[0, 0, 20] = 1.0
    y = G.aa_activation(x, lambda v: v, f, f)[0, 0]
    nz = torch.nonzero(y.abs() > 0).flatten()
    assert int(nz.min()) == 15 and int(nz.max()) == 25
    assert abs(float(y.sum()) - 1.0) < 5e-7


def test_aa_activation_matches_independent_copy():
    # cross-check against the copy shipped in transformers (qwen2_5_omni), if importable
    mod = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    act = mod.TorchActivation1d(torch.sin)
    f = G.kaiser_sinc_taps()
    x = torch.randn(2, 5, 33)
    assert float((act(x) - G.aa_activation(x, torch.sin, f, f)).abs().max()) < 1e-6


def test_aa_snake_explicit_formula_B4():
    """Appendix B4 polyphase/clamp formula (what the CUDA kernel implements) == conv formulation."""
    torch.manual_seed(0)
    f = G.kaiser_sinc_taps().double()
    for L in (1, 2, 5, 6, 17):
        x = torch.randn(L).double()
        a, b = 1.3, 0.7
        xc = lambda i: x[min(max(i, 0), L - 1)]
        up = torch.zeros(2 * L).double()
        for t in range(L):
            up[2 * t] = 2 * sum(f[2 * q + 1] * xc(t + 2 - q) for q in range(6))
            up[2 * t + 1] = 2 * sum(f[2 * q] * xc(t + 3 - q) for q in range(6))
        v = up + torch.sin(a * up) ** 2 / (b + 1e-9)
        vc = lambda n: v[min(max(n, 0), 2 * L - 1)]
        out = torch.tensor([sum(f[j] * vc(2 * t + j - 5) for j in range(12)) for t in range(L)])
        ref = G.aa_activation(x.float().view(1, 1, L),
                              lambda u: u + torch.sin(a * u) ** 2 / (b + 1e-9), f.float(), f.float())
        assert float((ref.view(-1).double() - out).abs().max()) < 2e-6


def test_istft_same_inverts_reference_stft_framing():
    # KAT 1: ISTFT("same") inverts reflect-pad + stft(center=False) framing of
    # fish_vocoder/data/transforms/spectrogram.py:29-49
    for n_fft, hop in ((1024, 256), (2048, 512), (64, 16)):
        win = n_fft
        torch.manual_seed(0)
        T = 12
        y = torch.randn(2, T * hop)
        window = torch.hann_window(win)
        p = (win - hop) // 2
        yp = F.pad(y[:, None], (p, p), mode="reflect")[:, 0]
        spec = torch.stft(yp, n_fft, hop_length=hop, win_length=win, window=window, center=False,
                          return_complex=True)
        rec = G.istft_same(spec, n_fft, hop, win, window)
        assert rec.shape == y.shape
        assert float((rec - y).abs().max()) < 2e-5


def test_irfft_ignores_dead_head_channels():
    # KAT 2: irfft(S[:, :n_fft]) only consumes bins 0..n_fft/2 and ignores Im(DC), Im(Nyquist)
    n_fft = 64
    S = torch.randn(1, n_fft, 5, dtype=torch.complex64)
    a = torch.fft.irfft(S, n_fft, dim=1)
    S2 = S[:, : n_fft // 2 + 1].clone()
    S2[:, 0] = S2[:, 0].real + 0j
    S2[:, -1] = S2[:, -1].real + 0j
    b = torch.fft.irfft(S2, n_fft, dim=1)
    assert float((a - b).abs().max()) < 1e-6


def test_istft_as_basis_gemm_B6():
    """Appendix B6: windowed inverse-DFT basis GEMM + overlap-add == oracle ISTFT (kernel math)."""
    n_fft, hop = 64, 16
    K = n_fft // 2 + 1
    T = 9
    torch.manual_seed(1)
    S = torch.randn(2, K, T, dtype=torch.complex64)
    window = torch.hann_window(n_fft)
    want = G.istft_same(S, n_fft, hop, n_fft, window)
    n = torch.arange(n_fft).double()
    k = torch.arange(K).double()
    ck = torch.full((K,), 2.0).double()
    ck[0] = ck[-1] = 1.0
    ang = 2 * math.pi * k[:, None] * n[None, :] / n_fft
    basis_re = (ck[:, None] * torch.cos(ang) * window.double()[None, :] / n_fft)
    basis_im = (-ck[:, None] * torch.sin(ang) * window.double()[None, :] / n_fft)
    frames = torch.einsum("bkt,kn->btn", S.real.double(), basis_re) + torch.einsum(
        "bkt,kn->btn", S.imag.double(), basis_im)
    pad = (n_fft - hop) // 2
    out = torch.zeros(2, T * hop).double()
    env = torch.zeros(T * hop).double()
    for s in range(T * hop):
        for fr in range(T):
            idx = s + pad - fr * hop
            if 0 <= idx < n_fft:
                out[:, s] += frames[:, fr, idx]
                env[s] += float(window[idx]) ** 2
    got = (out / env).float()
    assert float((got - want).abs().max()) < 1e-5


def test_weight_norm_fold_identity():
    # KAT 4: w == g * v / ||v|| for Conv1d (dim 0 = out) and ConvTranspose1d (dim 0 = in)
    from torch.nn.utils.parametrizations import weight_norm
    torch.manual_seed(0)
    for mod in (torch.nn.Conv1d(6, 4, 3), torch.nn.ConvTranspose1d(6, 4, 4, 2)):
        m = weight_norm(mod)
        m.parametrizations.weight.original0.data.uniform_(0.5, 2.0)
        sd = {"c." + k: v for k, v in m.state_dict().items()}
        assert float((G.wn_weight(sd, "c") - m.weight.detach()).abs().max()) < 1e-6


def test_conv_transpose_polyphase_formula_B3():
    """Appendix B3 polyphase index formula (what the CUDA path implements) == F.conv_transpose1d."""
    torch.manual_seed(0)
    for (k, u) in ((16, 8), (4, 2), (8, 2), (2, 2), (11, 5), (10, 5), (8, 4)):
        p = (k - u) // 2
        Ci, Co, L = 3, 2, 7
        x = torch.randn(1, Ci, L).double()
        w = torch.randn(Ci, Co, k).double()
        ref = F.conv_transpose1d(x, w, stride=u, padding=p)[0]
        L_out = (L - 1) * u - 2 * p + k          # == L*u when k-u is even (all generator yamls)
        assert ref.shape[-1] == L_out
        out = torch.zeros(Co, L_out).double()
        for q in range(-(-L_out // u)):
            for r in range(u):
                if q * u + r >= L_out:
                    continue
                m_lo = -((r + p) // u)
                m = m_lo
                while m * u + r + p < k:
                    j = m * u + r + p
                    if j >= 0 and 0 <= q - m < L:
                        out[:, q * u + r] += x[0, :, q - m] @ w[:, :, j]
                    m += 1
        assert float((out - ref).abs().max()) < 1e-10


def test_linear_resample_pairs_B7():
    # KAT 6: nn.Upsample(linear) x1/2 -> mean(x[2i], x[2i+1]);  x1/8 -> mean(x[8i+3], x[8i+4])
    x = torch.randn(1, 2, 32)
    a = G.linear_resample(x, 0.5)
    assert torch.allclose(a, 0.5 * (x[..., 0::2] + x[..., 1::2]), atol=1e-6)
    b = G.linear_resample(x, 0.125)
    assert torch.allclose(b, 0.5 * (x[..., 3::8] + x[..., 4::8]), atol=1e-6)


def test_batch_independence_and_length():
    kwargs, sd, ins, out, _ = load_golden("hifigan_small_stress")
    with torch.no_grad():
        y0 = G.hifigan_forward(sd, ins["mel"][:1], kwargs["upsample_rates"], kwargs["resblock_dilation_sizes"])
    assert y0.shape[-1] == ins["mel"].shape[-1] * kwargs["hop_length"]
    assert float((y0 - out[:1]).abs().max()) < 1e-6
