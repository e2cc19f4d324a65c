"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product package).

A functional, state-dict driven restatement (plain torch fp32 on CPU) of the reference's
mel -> waveform generator forward passes.  Each function cites the reference file:line it follows
(paths relative to /root/reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.

Pinning: ``tests/test_oracle_cpu.py`` checks every function here against golden vectors produced by
running the *unmodified* reference modules in the build container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``).  The two third-party ops the reference imports (alias_free_torch.Activation1d,
vocos.spectral_ops.ISTFT) are not shipped with the reference; their restatement (``oracle/shim``)
is UNPINNED against the real wheels and pinned only by known-answer tests + an independent local copy.
"""
from __future__ import annotations

import math
from typing import Mapping, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Mapping[str, Tensor]


# --------------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------------
def wn_weight(sd: SD, prefix: str) -> Tensor:
    """Effective weight of a (possibly weight-normed) conv.

    torch.nn.utils.parametrizations.weight_norm, as applied at hifigan.py:31-97,158,178,214:
    w = g * v / ||v||, norm over every dim except dim 0 (out-channels for Conv1d, in-channels for
    ConvTranspose1d).  Keys: ``parametrizations.weight.original0`` = g, ``original1`` = v.
    """
    k0 = prefix + ".parametrizations.weight.original0"
    if k0 in sd:
        g = sd[k0]
        v = sd[prefix + ".parametrizations.weight.original1"]
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
        return g * v / nrm
    return sd[prefix + ".weight"]


def same_pad(k: int, d: int = 1) -> int:
    """hifigan.py:21-22 get_padding."""
    return (k * d - d) // 2


def conv_same(sd: SD, prefix: str, x: Tensor, dilation: int = 1) -> Tensor:
    w = wn_weight(sd, prefix)
    return F.conv1d(x, w, sd[prefix + ".bias"], dilation=dilation,
                    padding=same_pad(w.shape[-1], dilation))


def conv_transpose(sd: SD, prefix: str, x: Tensor, stride: int) -> Tensor:
    """hifigan.py:177-187: ConvTranspose1d(C, C/2, k, u, padding=(k-u)//2)."""
    w = wn_weight(sd, prefix)
    k = w.shape[-1]
    return F.conv_transpose1d(x, w, sd[prefix + ".bias"], stride=stride, padding=(k - stride) // 2)


def _count(sd: SD, fmt: str) -> int:
    n = 0
    while any(key.startswith(fmt.format(n)) for key in sd):
        n += 1
    return n


def kaiser_sinc_taps(cutoff: float = 0.25, half_width: float = 0.3, k: int = 12) -> Tensor:
    """alias-free-torch 0.0.6 kaiser_sinc_filter1d; float32 taps of length k."""
    half = k // 2
    atten = 2.285 * (half - 1) * math.pi * (4.0 * half_width) + 7.95
    if atten > 50.0:
        beta = 0.1102 * (atten - 8.7)
    elif atten >= 21.0:
        beta = 0.5842 * (atten - 21.0) ** 0.4 + 0.07886 * (atten - 21.0)
    else:
        beta = 0.0
    win = torch.kaiser_window(k, periodic=False, beta=beta, dtype=torch.float32)
    t = (torch.arange(-half, half, dtype=torch.float32) + 0.5) if k % 2 == 0 else (
        torch.arange(k, dtype=torch.float32) - half)
    f = 2.0 * cutoff * win * torch.sinc(2.0 * cutoff * t)
    return f / f.sum()


def aa_activation(x: Tensor, act, f_up: Tensor, f_down: Tensor, edge_mode: str = "replicate") -> Tensor:
    """alias_free_torch.Activation1d (up 2x -> act -> down 2x), replicate edges.

    up  : pad 5|5 replicate, 2 * conv_transpose1d(stride 2, depthwise f), crop 15|15
    down: pad 5|6 replicate, conv1d(stride 2, depthwise f)
    edge_mode: "replicate" (BigVGAN-flavoured copy, default) | "reflect" | "zero" - the padding mode of BOTH filters
    (SURVEY 8c unresolved ambiguity: the PyPI 0.0.6 wheel may differ and cannot be inspected offline).
    """
    C = x.shape[1]
    k = f_up.numel()
    pad = k // 2 - 1
    crop_l = pad * 2 + (k - 2) // 2
    crop_r = pad * 2 + (k - 2 + 1) // 2
    pmode = {"replicate": "replicate", "reflect": "reflect", "zero": "constant"}[edge_mode]
    u = F.pad(x, (pad, pad), mode=pmode)
    u = 2.0 * F.conv_transpose1d(u, f_up.reshape(1, 1, -1).expand(C, 1, -1), stride=2, groups=C)
    u = u[..., crop_l:-crop_r]
    u = act(u)
    kd = f_down.numel()
    u = F.pad(u, (kd // 2 - 1, kd // 2), mode=pmode)
    return F.conv1d(u, f_down.reshape(1, 1, -1).expand(C, 1, -1), stride=2, groups=C)


def snake_beta(x: Tensor, alpha: Tensor, beta: Tensor, logscale: bool = True) -> Tensor:
    """bigvgan.py:122-135 SnakeBeta.forward."""
    a = alpha[None, :, None]
    b = beta[None, :, None]
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a).pow(2)


def snake(x: Tensor, alpha: Tensor, logscale: bool = True) -> Tensor:
    """bigvgan.py:60-71 Snake.forward (beta == alpha)."""
    a = alpha[None, :, None]
    if logscale:
        a = torch.exp(a)
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a).pow(2)


# --------------------------------------------------------------------------------------------
# HiFiGAN  (fish_vocoder/modules/generators/hifigan.py)
# --------------------------------------------------------------------------------------------
def hifigan_resblock1(sd: SD, prefix: str, x: Tensor, dilations: Sequence[int]) -> Tensor:
    """hifigan.py:101-108 ResBlock1.forward."""
    for i, d in enumerate(dilations):
        xt = F.silu(x)
        xt = conv_same(sd, f"{prefix}.convs1.{i}", xt, d)
        xt = F.silu(xt)
        xt = conv_same(sd, f"{prefix}.convs2.{i}", xt, 1)
        x = xt + x
    return x


def hifigan_forward(sd: SD, mel: Tensor, upsample_rates: Sequence[int],
                    resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5),) * 3,
                    template: Tensor | None = None, post_activation=F.silu) -> Tensor:
    """hifigan.py:226-249 HiFiGANGenerator.forward.  mel [B, n_mels, T] -> wav [B, 1, T*hop]."""
    x = conv_same(sd, "conv_pre", mel)
    use_template = any(k.startswith("noise_convs.") for k in sd)
    for i, u in enumerate(upsample_rates):
        x = F.silu(x)
        x = conv_transpose(sd, f"ups.{i}", x, u)
        if use_template:  # hifigan.py:233-234
            x = x + noise_conv(sd, i, template, upsample_rates)
        nk = _count(sd, f"resblocks.{i}.blocks." + "{}.")
        ys = [hifigan_resblock1(sd, f"resblocks.{i}.blocks.{j}", x, resblock_dilation_sizes[j])
              for j in range(nk)]
        x = torch.stack(ys, 0).mean(0)  # hifigan.py:132-133 ParralelBlock
    x = post_activation(x)
    x = conv_same(sd, "conv_post", x)
    return torch.tanh(x)


def noise_conv(sd: SD, i: int, template: Tensor, upsample_rates: Sequence[int]) -> Tensor:
    """hifigan.py:191-204: Conv1d(1, C_i, 2s, stride s, pad s//2), last stage kernel 1."""
    w, b = sd[f"noise_convs.{i}.weight"], sd[f"noise_convs.{i}.bias"]
    if i + 1 < len(upsample_rates):
        s = int(math.prod(upsample_rates[i + 1:]))
        return F.conv1d(template, w, b, stride=s, padding=s // 2)
    return F.conv1d(template, w, b)


# --------------------------------------------------------------------------------------------
# BigVGAN  (fish_vocoder/modules/generators/bigvgan.py)
# --------------------------------------------------------------------------------------------
def _act1d(sd: SD, prefix: str, x: Tensor, kind: str | None = None, edge_mode: str = "replicate",
           logscale: bool = True) -> Tensor:
    """Activation1d(SnakeBeta | Snake) (bigvgan.py:226-233).  kind None = by the keys present (SnakeBeta has `act.beta`);
    `logscale` is a constructor flag of the activation (bigvgan.py:34,90), not visible in the state dict."""
    f_up = sd[prefix + ".upsample.filter"].reshape(-1)
    f_dn = sd[prefix + ".downsample.lowpass.filter"].reshape(-1)
    if kind is None:
        kind = "snakebeta" if (prefix + ".act.beta") in sd else "snake"
    if kind == "snakebeta":
        fn = lambda v: snake_beta(v, sd[prefix + ".act.alpha"], sd[prefix + ".act.beta"], logscale)
    else:
        fn = lambda v: snake(v, sd[prefix + ".act.alpha"], logscale)
    return aa_activation(x, fn, f_up, f_dn, edge_mode)


def bigvgan_ampblock(sd: SD, prefix: str, x: Tensor, dilations: Sequence[int], edge_mode: str = "replicate",
                     logscale: bool = True) -> Tensor:
    """bigvgan.py:235-245 AMPBlock.forward (activations[::2] before convs1, [1::2] before convs2)."""
    for i, d in enumerate(dilations):
        xt = _act1d(sd, f"{prefix}.activations.{2 * i}", x, edge_mode=edge_mode, logscale=logscale)
        xt = conv_same(sd, f"{prefix}.convs1.{i}", xt, d)
        xt = _act1d(sd, f"{prefix}.activations.{2 * i + 1}", xt, edge_mode=edge_mode, logscale=logscale)
        xt = conv_same(sd, f"{prefix}.convs2.{i}", xt, 1)
        x = xt + x
    return x


def bigvgan_forward(sd: SD, mel: Tensor, upsample_rates: Sequence[int],
                    resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5),) * 3,
                    template: Tensor | None = None, post_kind: str | None = None,
                    edge_mode: str = "replicate", block_logscale=None) -> Tensor:
    """bigvgan.py:352-371 BigVGANGenerator.forward (no activation before ups, bigvgan.py:355-356).
    block_logscale: optional {resblock index: bool} for AMPBlocks built with snake_logscale=False (bigvgan.py:145)."""
    x = conv_same(sd, "conv_pre", mel)
    nk = len(resblock_dilation_sizes)
    use_template = any(k.startswith("noise_convs.") for k in sd)
    for i, u in enumerate(upsample_rates):
        x = conv_transpose(sd, f"ups.{i}", x, u)
        if use_template:
            x = x + noise_conv(sd, i, template, upsample_rates)
        ys = [bigvgan_ampblock(sd, f"resblocks.{i * nk + j}", x, resblock_dilation_sizes[j], edge_mode,
                               (block_logscale or {}).get(i * nk + j, True))
              for j in range(nk)]
        x = torch.stack(ys, 0).mean(0)
    x = _act1d(sd, "activation_post", x, post_kind, edge_mode)
    x = conv_same(sd, "conv_post", x)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------------
# ConvNeXt backbone + ISTFT head  (encoders/convnext.py, generators/vocos.py, generators/unify.py)
# --------------------------------------------------------------------------------------------
def ln_channels_first(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """convnext.py:69-74 LayerNorm(data_format="channels_first")."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None] * x + b[:, None]


def convnext_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """convnext.py:124-143 ConvNeXtBlock.forward (eval: DropPath is identity)."""
    C = x.shape[1]
    k = sd[p + ".dwconv.weight"].shape[-1]
    h = F.conv1d(x, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=(k - 1) // 2, groups=C)
    h = h.permute(0, 2, 1)
    h = F.layer_norm(h, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = F.linear(h, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])
    if p + ".gamma" in sd:
        h = sd[p + ".gamma"] * h
    return x + h.permute(0, 2, 1)


def convnext_forward(sd: SD, x: Tensor, prefix: str = "") -> Tensor:
    """convnext.py:206-214 ConvNeXtEncoder.forward."""
    n_stage = _count(sd, prefix + "downsample_layers.{}.")
    for i in range(n_stage):
        d = f"{prefix}downsample_layers.{i}"
        if i == 0:  # stem: Conv1d(k, pad k//2) + LN(channels_first)   convnext.py:160-169
            w = sd[d + ".0.weight"]
            x = F.conv1d(x, w, sd[d + ".0.bias"], padding=w.shape[-1] // 2)
            x = ln_channels_first(x, sd[d + ".1.weight"], sd[d + ".1.bias"])
        else:       # LN(channels_first) + Conv1d 1x1                   convnext.py:172-177
            x = ln_channels_first(x, sd[d + ".0.weight"], sd[d + ".0.bias"])
            x = F.conv1d(x, sd[d + ".1.weight"], sd[d + ".1.bias"])
        depth = _count(sd, f"{prefix}stages.{i}." + "{}.")
        for j in range(depth):
            x = convnext_block(sd, f"{prefix}stages.{i}.{j}", x)
    return ln_channels_first(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])


def istft_same(spec: Tensor, n_fft: int, hop: int, win: int, window: Tensor) -> Tensor:
    """vocos==0.0.2 ISTFT(padding="same"): irfft -> window -> overlap-add -> trim -> /env."""
    pad = (win - hop) // 2
    B, N, T = spec.shape
    frames = torch.fft.irfft(spec, n_fft, dim=1, norm="backward") * window[None, :, None]
    out_len = (T - 1) * hop + win
    y = frames.new_zeros(B, out_len)
    env = torch.zeros(out_len)
    wsq = window.square()
    for t in range(T):
        y[:, t * hop:t * hop + win] += frames[:, :, t]
        env[t * hop:t * hop + win] += wsq
    y, env = y[:, pad:out_len - pad], env[pad:out_len - pad]
    assert (env > 1e-11).all()
    return y / env


def istft_center(spec: Tensor, n_fft: int, hop: int, win: int, window: Tensor) -> Tensor:
    """vocos==0.0.2 ISTFT(padding="center") = torch.istft(spec, n_fft, hop, win, window, center=True), restated:
    the window is zero-padded (centred) to n_fft; a two-sided spectrum (n_fft rows - what the reference head produces
    with its 2*n_fft outputs, vocos.py:40-41,57-69) is cut to its first n_fft/2+1 rows (ATen istft with a real output:
    `input.slice(-1, 0, n_fft/2+1)` before the c2r transform), then irfft; overlap-add, divide by the squared-window
    envelope, trim n_fft/2 either side -> (T-1)*hop samples.  Checked against torch.istft in tests/test_oracle_cpu.py."""
    B, N, T = spec.shape
    if win < n_fft:
        left = (n_fft - win) // 2
        window = F.pad(window, (left, n_fft - win - left))
    frames = torch.fft.irfft(spec[:, :n_fft // 2 + 1], n_fft, dim=1, norm="backward")
    frames = frames * window[None, :, None]
    out_len = (T - 1) * hop + n_fft
    y = frames.new_zeros(B, out_len)
    env = torch.zeros(out_len)
    wsq = window.square()
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += frames[:, :, t]
        env[t * hop:t * hop + n_fft] += wsq
    lo, hi = n_fft // 2, out_len - n_fft // 2
    y, env = y[:, lo:hi], env[lo:hi]
    assert (env.abs() > 1e-11).all()
    return y / env


def istft_head_forward(sd: SD, x: Tensor, n_fft: int, hop: int, win: int, prefix: str = "",
                       padding: str = "same") -> Tensor:
    """vocos.py:43-69 ISTFTHead.forward.  `out.weight` [2*n_fft, dim, 1] = the reference layout; [n_fft+2, dim] = the
    upstream-Vocos layout (vocos==0.0.2 heads.ISTFTHead: nn.Linear(dim, n_fft + 2), scripts/vocos_gen.py:5-16)."""
    w = sd[prefix + "out.weight"]
    if w.ndim == 2:
        w = w[:, :, None]
    x = F.conv1d(x, w, sd[prefix + "out.bias"])
    mag, p = x.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(p) + 1j * torch.sin(p))
    if padding == "center":
        return istft_center(S, n_fft, hop, win, sd[prefix + "istft.window"])
    return istft_same(S, n_fft, hop, win, sd[prefix + "istft.window"])


def vocos_backbone_forward(sd: SD, x: Tensor, prefix: str = "") -> Tensor:
    """Upstream vocos==0.0.2 models.VocosBackbone.forward (third-party, NOT in /root/reference; restated from its
    published source - parity unpinned): embed Conv1d(k=7, pad 3) -> LayerNorm over C -> n x ConvNeXtBlock (dwconv k7,
    LayerNorm, Linear, GELU, Linear, gamma, residual) -> final LayerNorm.  Returns [B, dim, T] (channels-first; upstream
    hands [B, T, dim] to the head, whose Linear is the same contraction)."""
    x = F.conv1d(x, sd[prefix + "embed.weight"], sd[prefix + "embed.bias"], padding=sd[prefix + "embed.weight"].shape[-1] // 2)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], 1e-6)
    x = x.transpose(1, 2)
    n = _count(sd, prefix + "convnext.{}.")
    for j in range(n):
        x = convnext_block(sd, f"{prefix}convnext.{j}", x)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[prefix + "final_layer_norm.weight"],
                     sd[prefix + "final_layer_norm.bias"], 1e-6)
    return x.transpose(1, 2)


def upstream_vocos_forward(sd: SD, mel: Tensor, n_fft: int, hop: int, padding: str = "center") -> Tensor:
    """vocos.Vocos.decode (scripts/vocos_gen.py:12-16): VocosBackbone -> ISTFTHead(out_dim = n_fft + 2, win = n_fft)."""
    x = vocos_backbone_forward(sd, mel, "backbone.")
    return istft_head_forward(sd, x, n_fft, hop, n_fft, "head.", padding)[:, None, :]


def unify_vocos_forward(sd: SD, mel: Tensor, n_fft: int, hop: int, win: int, padding: str = "same") -> Tensor:
    """unify.py:18-33 UnifyGenerator.forward with ConvNeXtEncoder backbone + ISTFTHead (no vq)."""
    x = convnext_forward(sd, mel, "backbone.")
    x = istft_head_forward(sd, x, n_fft, hop, win, "head.", padding)
    return x[:, None, :]


def unify_hifigan_forward(sd: SD, mel: Tensor, upsample_rates: Sequence[int],
                          resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5),) * 3) -> Tensor:
    """unify.py:18-33 with ConvNeXtEncoder backbone + HiFiGANGenerator head = configs/model/generator/firefly-gan-base.yaml
    (the head consumes the backbone's [B, dims[-1], T] output as its "mel", pre/post kernels 13)."""
    x = convnext_forward(sd, mel, "backbone.")
    head = {k[len("head."):]: v for k, v in sd.items() if k.startswith("head.")}
    return hifigan_forward(head, x, upsample_rates, resblock_dilation_sizes)


# --------------------------------------------------------------------------------------------
# RefineGAN  (fish_vocoder/modules/generators/refinegan.py)
# --------------------------------------------------------------------------------------------
def refinegan_resblock(sd: SD, p: str, x: Tensor, dilations=(1, 3, 5), slope: float = 0.2) -> Tensor:
    """refinegan.py:87-99 ResBlock.forward: both convs of a pair are dilated (refinegan.py:72-83)."""
    for i, d in enumerate(dilations):
        xt = F.leaky_relu(x, slope)
        xt = conv_same(sd, f"{p}.convs1.{i}", xt, d)
        xt = F.leaky_relu(xt, slope)
        xt = conv_same(sd, f"{p}.convs2.{i}", xt, d)
        if i != 0 or xt.shape[1] == x.shape[1]:
            x = xt + x
        else:
            x = xt
    return x


def linear_resample(x: Tensor, scale: float) -> Tensor:
    """nn.Upsample(scale_factor=scale, mode="linear") (refinegan.py:222,252)."""
    return F.interpolate(x, scale_factor=scale, mode="linear")


def refinegan_forward(sd: SD, mel: Tensor, template: Tensor, noise_fn,
                      downsample_rates=(2, 2, 8, 8), upsample_rates=(8, 8, 2, 2),
                      slope: float = 0.2) -> Tensor:
    """refinegan.py:287-323.  ``noise_fn(shape) -> Tensor`` supplies AdaIN's gaussian (refinegan.py:125)
    in call order (the reference draws torch.randn_like even in eval mode)."""
    def adain(p, v):  # refinegan.py:124-127
        return F.leaky_relu(v + noise_fn(v.shape) * sd[p + ".weight"][None, :, None], slope)

    x = conv_same(sd, "template_conv", template)
    downs = []
    for i, r in enumerate(downsample_rates):
        x = F.leaky_relu(x, slope)
        downs.append(x)
        x = linear_resample(x, 1.0 / r)
        x = refinegan_resblock(sd, f"downsample_blocks.{i}.1", x, slope=slope)
    x = torch.cat([x, conv_same(sd, "mel_conv", mel)], dim=1)
    for i, (r, down) in enumerate(zip(upsample_rates, reversed(downs))):
        x = F.leaky_relu(x, slope)
        x = linear_resample(x, float(r))
        x = torch.cat([x, down], dim=1)
        p = f"upsample_conv_blocks.{i}"
        x = F.conv1d(x, sd[p + ".input_conv.weight"], sd[p + ".input_conv.bias"], padding=3)
        ys = []
        for j in range(_count(sd, p + ".blocks.{}.")):
            y = adain(f"{p}.blocks.{j}.0", x)
            y = refinegan_resblock(sd, f"{p}.blocks.{j}.1", y, slope=slope)
            y = adain(f"{p}.blocks.{j}.2", y)
            ys.append(y)
        x = torch.stack(ys).mean(0)
    x = F.leaky_relu(x, slope)
    x = conv_same(sd, "output_conv", x)
    return torch.tanh(x)
