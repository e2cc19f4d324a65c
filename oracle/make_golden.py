"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (CPU, fp32).

Run only in the build container (needs /root/reference):  python oracle/make_golden.py
The reference is imported from where it lies (sys.path injection; nothing is copied); the two
third-party ops it needs come from oracle/shim (restated, see the headers there).
Each fixture stores: the constructor kwargs (json), the full state_dict, the inputs and the
reference output.  Also writes tests/golden/state_dict_contract.json = key -> shape for the
full-size configs (checkpoint layout contract).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = ["/root/reference", os.path.join(HERE, "shim")]
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from fish_vocoder.modules.encoders.convnext import ConvNeXtEncoder  # noqa: E402
from fish_vocoder.modules.generators.bigvgan import BigVGANGenerator  # noqa: E402
from fish_vocoder.modules.generators.hifigan import HiFiGANGenerator  # noqa: E402
from fish_vocoder.modules.generators.refinegan import RefineGANGenerator  # noqa: E402
from fish_vocoder.modules.generators.unify import UnifyGenerator  # noqa: E402
from fish_vocoder.modules.generators.vocos import ISTFTHead  # noqa: E402


def stress_init(model: torch.nn.Module, seed: int = 1) -> None:
    """SURVEY 8d "stress-init": make residual branches non-trivial like a trained net."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        if k.endswith("parametrizations.weight.original1"):
            fan_in = v[0].numel() if "ups." not in k else v.shape[0] * v.shape[2] / max(1, 1)
            if "ups." in k:  # ConvTranspose: [C_in, C_out, k]; each output sees C_in*k/u taps
                fan_in = v.shape[0] * v.shape[2]
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


def mel_input(B, n_mels, T, seed=1234):
    torch.manual_seed(seed)
    return torch.empty(B, n_mels, T).uniform_(-11.5129, 2.0)


def save(name, kwargs, model, inputs, out, extra=None):
    arrs = {"sd::" + k: v.detach().numpy() for k, v in model.state_dict().items()}
    for k, v in inputs.items():
        arrs["in::" + k] = v.detach().numpy()
    arrs["out"] = out.detach().numpy()
    arrs["kwargs"] = np.frombuffer(json.dumps(kwargs).encode(), dtype=np.uint8)
    for k, v in (extra or {}).items():
        arrs[k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: out {tuple(out.shape)} absmax {out.abs().max():.4f}  {os.path.getsize(path)/1e6:.2f} MB")


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    # ---------------- HiFiGAN (small, both inits) ----------------
    hk = dict(hop_length=32, upsample_rates=[4, 4, 2], upsample_kernel_sizes=[8, 8, 4],
              resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
              num_mels=20, upsample_initial_channel=64, use_template=False,
              pre_conv_kernel_size=7, post_conv_kernel_size=7)
    for tag, stress in (("ref", False), ("stress", True)):
        torch.manual_seed(0)
        m = HiFiGANGenerator(**hk).eval()
        if stress:
            stress_init(m)
        x = mel_input(2, 20, 13)
        save(f"hifigan_small_{tag}", hk, m, {"mel": x}, m(x))

    # HiFiGAN with template (noise_convs) and firefly-style k=13 pre/post, odd upsample mix (8,2 w/ k=8? no: k-u even)
    hk2 = dict(hop_length=20, upsample_rates=[5, 2, 2], upsample_kernel_sizes=[11, 8, 2],
               resblock_kernel_sizes=[3, 5], resblock_dilation_sizes=[[1, 3, 5], [1, 2, 3]],
               num_mels=12, upsample_initial_channel=32, use_template=True,
               pre_conv_kernel_size=13, post_conv_kernel_size=13)
    torch.manual_seed(0)
    m = HiFiGANGenerator(**hk2).eval()
    stress_init(m)
    x = mel_input(1, 12, 9)
    torch.manual_seed(7)
    tpl = torch.randn(1, 1, 9 * 20) * 0.3
    save("hifigan_template_stress", hk2, m, {"mel": x, "template": tpl}, m(x, tpl))

    # ---------------- BigVGAN (small) ----------------
    bk = dict(hop_length=32, upsample_rates=[4, 2, 2, 2], upsample_kernel_sizes=[8, 8, 2, 4],
              resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
              num_mels=20, upsample_initial_channel=64, use_template=False,
              pre_conv_kernel_size=7, post_conv_kernel_size=7)
    for tag, stress in (("ref", False), ("stress", True)):
        torch.manual_seed(0)
        m = BigVGANGenerator(**bk).eval()
        if stress:
            stress_init(m)
        x = mel_input(2, 20, 11)
        save(f"bigvgan_small_{tag}", bk, m, {"mel": x}, m(x))

    # ---------------- Vocos = ConvNeXt + ISTFT head (small) ----------------
    vk = dict(backbone=dict(input_channels=20, depths=[1, 2], dims=[32, 48], drop_path_rate=0.2,
                            kernel_size=7),
              head=dict(dim=48, n_fft=64, hop_length=16, win_length=64, padding="same"))
    for tag, stress in (("ref", False), ("stress", True)):
        torch.manual_seed(0)
        m = UnifyGenerator(backbone=ConvNeXtEncoder(**vk["backbone"]), head=ISTFTHead(**vk["head"])).eval()
        if stress:
            stress_init(m)
            sd = m.state_dict()
            g = torch.Generator().manual_seed(3)
            for k in sd:
                if k.endswith("weight") and sd[k].ndim >= 2:
                    sd[k] = sd[k] * 4.0
                if k.endswith("norm.weight") or (k.endswith(".weight") and sd[k].ndim == 1):
                    sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=g)
            m.load_state_dict(sd)
        x = mel_input(2, 20, 10)
        # the reference UnifyGenerator passes template= to ISTFTHead.forward(x) -> TypeError
        # (unify.py:25 vs vocos.py:43, SURVEY 8b defect 1); call backbone/head exactly as forward does.
        y = m.head(m.backbone(x))[:, None, :]
        save(f"vocos_small_{tag}", vk, m, {"mel": x}, y)

    # ---------------- RefineGAN (small), AdaIN noise injected deterministically ----------------
    rk = dict(sampling_rate=16000, hop_length=16, downsample_rates=[2, 8], upsample_rates=[8, 2],
              leaky_relu_slope=0.2, num_mels=12, start_channels=4)
    torch.manual_seed(0)
    m = RefineGANGenerator(**rk).eval()
    stress_init(m)
    x = mel_input(1, 12, 7)
    torch.manual_seed(9)
    tpl = torch.randn(1, 1, 7 * 16) * 0.3
    rs = np.random.RandomState(4321)
    orig = torch.randn_like
    torch.randn_like = lambda t, **kw: torch.from_numpy(rs.standard_normal(tuple(t.shape)).astype(np.float32))
    try:
        y = m(x, tpl)
    finally:
        torch.randn_like = orig
    save("refinegan_small_stress", rk, m, {"mel": x, "template": tpl}, y,
         extra={"noise_seed": np.array([4321])})

    # ---------------- checkpoint-layout contract for the full-size configs ----------------
    contract = {}
    torch.manual_seed(0)
    full = {
        "hifigan_cfgA": HiFiGANGenerator(hop_length=256, upsample_rates=(8, 8, 2, 2),
                                         upsample_kernel_sizes=(16, 16, 4, 4), num_mels=80,
                                         use_template=False),
        "hifigan_yaml_44k": HiFiGANGenerator(hop_length=512, num_mels=128, use_template=False),
        "bigvgan_cfgC": BigVGANGenerator(hop_length=512, num_mels=100, use_template=False),
        "vocos_yaml": UnifyGenerator(
            backbone=ConvNeXtEncoder(input_channels=128, depths=[3, 3, 27, 3], dims=[128, 256, 512, 1024],
                                     drop_path_rate=0.4, kernel_size=7),
            head=ISTFTHead(dim=1024, n_fft=2048, hop_length=512, win_length=2048, padding="same")),
        "refinegan_default": RefineGANGenerator(),
    }
    for name, mod in full.items():
        contract[name] = {k: list(v.shape) for k, v in mod.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_contract.json"), "w") as f:
        json.dump(contract, f, indent=0, sort_keys=True)
    print("contract:", {k: len(v) for k, v in contract.items()})


if __name__ == "__main__":
    main()
