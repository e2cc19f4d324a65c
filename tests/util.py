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
        return G.bigvgan_forward(sd, ins["mel"], kwargs["upsample_rates"], kwargs["resblock_dilation_sizes"],
                                 template=ins.get("template"))
    if name.startswith("vocos"):
        h = kwargs["head"]
        return G.unify_vocos_forward(sd, ins["mel"], h["n_fft"], h["hop_length"], h["win_length"])
    if name.startswith("refinegan"):
        return G.refinegan_forward(sd, ins["mel"], ins["template"], numpy_noise_fn(extra["noise_seed"][0]),
                                   kwargs["downsample_rates"], kwargs["upsample_rates"], kwargs["leaky_relu_slope"])
    raise KeyError(name)


ALL_GOLDEN = ["hifigan_small_ref", "hifigan_small_stress", "hifigan_template_stress", "bigvgan_small_ref",
              "bigvgan_small_stress", "vocos_small_ref", "vocos_small_stress", "refinegan_small_stress"]


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
