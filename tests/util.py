"""Shared helpers for the test-suite (golden loading, oracle dispatch)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    ins = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in::")}
    kwargs = json.loads(bytes(z["kwargs"]).decode())
    extra = {k: z[k] for k in z.files if not (k.startswith("sd::") or k.startswith("in::") or k in ("out", "kwargs"))}
    return kwargs, sd, ins, torch.from_numpy(z["out"]), extra


def numpy_noise_fn(seed):
    rs = np.random.RandomState(int(seed))
    return lambda shape: torch.from_numpy(rs.standard_normal(tuple(shape)).astype(np.float32))


def oracle_forward(name, kwargs, sd, ins, extra=None):
    """Run the CPU oracle for a golden fixture family."""
    from oracle import generators as G

    if name.startswith("hifigan"):
        return G.hifigan_forward(sd, ins["mel"], kwargs["upsample_rates"], kwargs["resblock_dilation_sizes"],
                                 template=ins.get("template"))
    if name.startswith("bigvgan"):
        # bigvgan_snake_mix_stress: the four AMPBlocks of the first two stages were built with snake_logscale=False
        block_logscale = {i: False for i in range(4)} if name.startswith("bigvgan_snake_mix") else None
        return G.bigvgan_forward(sd, ins["mel"], kwargs["upsample_rates"], kwargs["resblock_dilation_sizes"],
                                 template=ins.get("template"), block_logscale=block_logscale)
    if name.startswith("vocos"):
        h = kwargs["head"]
        return G.unify_vocos_forward(sd, ins["mel"], h["n_fft"], h["hop_length"], h["win_length"],
                                     h.get("padding", "same"))
    if name.startswith("firefly"):
        return G.unify_hifigan_forward(sd, ins["mel"], kwargs["head"]["upsample_rates"],
                                       kwargs["head"]["resblock_dilation_sizes"])
    if name.startswith("refinegan"):
        return G.refinegan_forward(sd, ins["mel"], ins["template"], numpy_noise_fn(extra["noise_seed"][0]),
                                   kwargs["downsample_rates"], kwargs["upsample_rates"], kwargs["leaky_relu_slope"])
    raise KeyError(name)


ALL_GOLDEN = ["hifigan_small_ref", "hifigan_small_stress", "hifigan_template_stress", "bigvgan_small_ref",
              "bigvgan_small_stress", "vocos_small_ref", "vocos_small_stress", "refinegan_small_stress", "firefly_small_stress",
              "vocos_center_stress", "bigvgan_snake_mix_stress", "bigvgan_template_stress"]


def build_module(name, kwargs):
    """Our module for a golden fixture (same constructor calls oracle/make_golden*.py made on the reference classes)."""
    from vocoder_b200.encoders import ConvNeXtEncoder
    from vocoder_b200.generators import BigVGANGenerator, HiFiGANGenerator, ISTFTHead, UnifyGenerator
    if name.startswith("hifigan"):
        return HiFiGANGenerator(**kwargs)
    if name.startswith("bigvgan_snake_mix"):
        from vocoder_b200.generators.bigvgan import AMPBlock, Snake, SnakeBeta
        m = BigVGANGenerator(activation=Snake, **kwargs)
        nk = len(kwargs["resblock_kernel_sizes"])
        for j, (k, d) in enumerate(zip(kwargs["resblock_kernel_sizes"], kwargs["resblock_dilation_sizes"])):
            m.resblocks[j] = AMPBlock(m.stage_channels[0], k, tuple(d), activation=Snake, snake_logscale=False)
            m.resblocks[nk + j] = AMPBlock(m.stage_channels[1], k, tuple(d), activation=SnakeBeta, snake_logscale=False)
        return m
    if name.startswith("bigvgan"):
        return BigVGANGenerator(**kwargs)
    if name.startswith("vocos"):
        return UnifyGenerator(backbone=ConvNeXtEncoder(**kwargs["backbone"]), head=ISTFTHead(**kwargs["head"]))
    if name.startswith("firefly"):  # configs/model/generator/firefly-gan-base.yaml: ConvNeXt backbone + HiFiGAN head
        return UnifyGenerator(backbone=ConvNeXtEncoder(**kwargs["backbone"]), head=HiFiGANGenerator(**kwargs["head"]))
    if name.startswith("refinegan"):
        from vocoder_b200.generators.refinegan import RefineGANGenerator
        return RefineGANGenerator(**kwargs)
    raise KeyError(name)


def channels_last_noise(seed):
    """AdaIN noise for the RefineGAN parity test: the same numpy stream the golden was generated with
    (drawn channels-first [B, C, L] like torch.randn_like in refinegan.py:125), handed over channels-last."""
    rs = np.random.RandomState(int(seed))

    def fn(B, L, C, pitch, device):
        z = torch.from_numpy(rs.standard_normal((B, C, L)).astype(np.float32))
        out = torch.zeros(B, L, pitch)
        out[..., :C] = z.permute(0, 2, 1)
        return out.to(device)

    return fn


class emulate_f16_operands:
    """Context manager: run the CPU oracle with the operands of every dense conv / conv-transpose / linear rounded
    to fp16 (round-to-nearest), fp32 accumulation - the arithmetic of the default "f16" tensor-core mode
    (depthwise convs, anti-alias filters, LayerNorm, activations and residual adds stay fp32, as in the kernels)."""

    def __enter__(self):
        import torch.nn.functional as F
        self.F = F
        self.saved = (F.conv1d, F.conv_transpose1d, F.linear)
        c1, ct, li = self.saved
        r = lambda t: t.half().float()

        def conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
            if groups == 1:
                x, w = r(x), r(w)
            return c1(x, w, b, stride, padding, dilation, groups)

        def conv_transpose1d(x, w, b=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
            if groups == 1:
                x, w = r(x), r(w)
            return ct(x, w, b, stride, padding, output_padding, groups, dilation)

        def linear(x, w, b=None):
            return li(r(x), r(w), b)

        F.conv1d, F.conv_transpose1d, F.linear = conv1d, conv_transpose1d, linear
        return self

    def __exit__(self, *exc):
        self.F.conv1d, self.F.conv_transpose1d, self.F.linear = self.saved
        return False


def stress_init(model: torch.nn.Module, seed: int = 1) -> None:
    """SURVEY 8d "stress-init" (same distributions as oracle/make_golden.py): weight-norm directions ~ N(0, (s/sqrt(fan_in))^2)
    with s = 0.5 for the residual-block convs and 1 elsewhere, g = ||v||, biases ~ N(0, 0.02^2), Snake alpha/beta ~
    N(0, 0.5^2), ConvNeXt gamma ~ U(0.05, 0.5): residual branches carry signal like a trained network's."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith("parametrizations.weight.original1"):
            fan_in = v.shape[0] * v.shape[2] if "ups." in k else v[0].numel()
            s = 0.5 if (".convs1." in k or ".convs2." in k) else 1.0
            new[k] = torch.randn(v.shape, generator=g) * (s / fan_in ** 0.5)
        elif k.endswith(".bias") and v.ndim == 1:
            new[k] = torch.randn(v.shape, generator=g) * 0.02
        elif k.endswith(".act.alpha") or k.endswith(".act.beta"):
            new[k] = torch.randn(v.shape, generator=g) * 0.5
        elif k.endswith(".gamma"):
            new[k] = torch.rand(v.shape, generator=g) * 0.45 + 0.05
    for k, v in list(new.items()):
        if k.endswith("original1"):
            k0 = k[:-1] + "0"
            new[k0] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(sd[k0].shape)
    sd.update(new)
    model.load_state_dict(sd)
