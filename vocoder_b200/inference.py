"""Inference helpers around the generator forward: Lightning-checkpoint ingest without
Hydra/Lightning, bounded-memory chunked synthesis of long utterances, and a small CLI.

    python -m vocoder_b200.inference --config gen.yaml --ckpt step_001235000.ckpt --input mels/ --output-dir out/

`gen.yaml` is the reference's generator yaml (fish_vocoder/configs/model/generator/*.yaml) with concrete numbers
instead of ${...} interpolations; `_target_` may name either the reference class or ours.  Inputs are `.pt` / `.npy`
mel tensors `[n_mels, T]` or `[B, n_mels, T]` (transposed automatically when the last dim is n_mels, as
fish_vocoder/test.py:73-84 does); outputs are 16-bit PCM `.wav` files.
"""
from __future__ import annotations

import argparse
import math
import os
import wave
from typing import Dict, Iterable, Optional

import numpy as np
import torch

import functools
import importlib

_TARGETS = {
    "VocosBackbone": ("vocoder_b200.encoders.vocos_backbone", "VocosBackbone"),
    "HiFiGANGenerator": ("vocoder_b200.generators.hifigan", "HiFiGANGenerator"),
    "BigVGANGenerator": ("vocoder_b200.generators.bigvgan", "BigVGANGenerator"),
    "RefineGANGenerator": ("vocoder_b200.generators.refinegan", "RefineGANGenerator"),
    "UnifyGenerator": ("vocoder_b200.generators.unify", "UnifyGenerator"),
    "ISTFTHead": ("vocoder_b200.generators.vocos", "ISTFTHead"),
    "ConvNeXtEncoder": ("vocoder_b200.encoders.convnext", "ConvNeXtEncoder"),
}


def instantiate(cfg):
    """Minimal stand-in for hydra.utils.instantiate (fish_vocoder/test.py:31): builds the object a `_target_` names,
    recursively, mapping the reference's dotted paths onto the vocoder_b200 classes."""
    if isinstance(cfg, dict) and "_target_" in cfg:
        target = cfg["_target_"]
        cls_name = target.split(".")[-1]
        if cls_name in _TARGETS:
            mod, name = _TARGETS[cls_name]
            cls = getattr(importlib.import_module(mod), name)
        elif target.startswith(("torch.nn.", "functools.")):
            # plain torch callables the reference yamls pass by name, e.g.
            # post_activation: {_target_: torch.nn.SiLU, _partial_: true}
            mod, name = target.rsplit(".", 1)
            cls = getattr(importlib.import_module(mod), name)
        else:
            raise KeyError(f"no B200 implementation for _target_ {target}")
        kwargs = {k: instantiate(v) for k, v in cfg.items() if k not in ("_target_", "_partial_", "_recursive_",
                                                                           "_convert_")}
        if cfg.get("_partial_", False):   # hydra: functools.partial(target, **kwargs)
            return functools.partial(cls, **kwargs)
        return cls(**kwargs)
    if isinstance(cfg, dict):
        return {k: instantiate(v) for k, v in cfg.items()}
    if isinstance(cfg, list):
        return [instantiate(v) for v in cfg]
    return cfg


def generator_state_dict(ckpt: Dict, prefix: str = "generator.") -> Dict[str, torch.Tensor]:
    """Generator tensors of a Lightning checkpoint (`ckpt["state_dict"]`, keys `generator.<...>`, test.py:32-37);
    also accepts a bare state dict with or without the prefix."""
    sd = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if any(k.startswith(prefix) for k in sd):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    return dict(sd)


def load_generator(cfg, ckpt_path: Optional[str], device="cuda") -> torch.nn.Module:
    gen = instantiate(cfg)
    if ckpt_path:
        try:  # tensors-only first; a Lightning checkpoint with pickled hyper-parameters needs the full loader
            ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=True)
        except Exception:  # noqa: BLE001
            ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        gen.load_state_dict(generator_state_dict(ckpt), strict=True)
    return gen.eval().to(device)


# ------------------------------------------------------------------------------------------------
# chunked synthesis
# ------------------------------------------------------------------------------------------------
def context_frames(gen) -> int:
    """Conservative one-sided receptive field of the generator in mel frames (input context a chunk needs so that its
    interior equals the un-chunked forward)."""
    from .encoders.convnext import ConvNeXtEncoder
    from .encoders.vocos_backbone import VocosBackbone
    from .generators._mrf import MRFGeneratorBase
    from .generators.unify import UnifyGenerator
    from .generators.vocos import ISTFTHead

    if isinstance(gen, MRFGeneratorBase):
        P_blocks = [gen._block_modules(i) for i in range(gen.num_upsamples)]
        reach = (gen.conv_post.kernel_size[0] - 1) // 2 + (5 if gen.snake_blocks else 0)  # output-rate samples
        for i in reversed(range(gen.num_upsamples)):
            blk_reach = 0
            for blk in P_blocks[i]:
                r = 0
                for c1, c2 in zip(blk.convs1, blk.convs2):
                    k = c1.kernel_size[0]
                    r += (k - 1) // 2 * c1.dilation[0] + (k - 1) // 2 * c2.dilation[0]
                    if gen.snake_blocks:
                        r += 10  # two anti-aliased activations, +-5 samples each
                blk_reach = max(blk_reach, r)
            reach += blk_reach
            u, k = gen.upsample_rates[i], gen.upsample_kernel_sizes[i]
            reach = math.ceil((reach + k) / u) + 1  # through the transposed conv, in input-rate samples
        reach += (gen.conv_pre.kernel_size[0] - 1) // 2
        return int(reach) + 1
    if isinstance(gen, ConvNeXtEncoder):
        return sum(gen.depths) * (gen.kernel_size // 2) + gen.kernel_size // 2 + 1
    if isinstance(gen, ISTFTHead):
        return gen.win_length // gen.hop_length + 1
    if isinstance(gen, VocosBackbone):
        return (len(gen.convnext) + 1) * 3 + 1
    if isinstance(gen, UnifyGenerator):
        return context_frames(gen.backbone) + context_frames(gen.head)
    raise TypeError(f"no receptive-field model for {type(gen).__name__}: synthesise it with a single forward "
                    "(RefineGAN's U-Net mixes time scales; chunk it upstream of the mel if memory is short)")


def hop_of(gen) -> int:
    if hasattr(gen, "hop_length"):
        return int(gen.hop_length)
    return hop_of(gen.head)


def _call(gen, mel, template):
    return gen(mel) if template is None else gen(mel, template)


@torch.no_grad()
def chunked_forward(gen, mel: torch.Tensor, chunk_frames: int = 2048, context: Optional[int] = None,
                    template: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mel [B, n_mels, T] -> wav [B, 1, T*hop] in chunks of `chunk_frames` frames (+ `context` frames of overlap on each
    side, recomputed and discarded), so device memory is bounded by the chunk, not by the utterance length: the
    generators keep workspace buffers and CUDA graphs for the few most recent input shapes only (runtime.Workspace), and
    a long file presents at most three (first / interior / last chunk).  `template` [B, 1, T*hop] (generators built with
    use_template=True, hifigan.py:233-234) is sliced alongside the mel."""
    B, _, T = mel.shape
    hop = hop_of(gen)
    if template is not None and template.shape[-1] != T * hop:
        raise ValueError(f"template must hold T*hop = {T * hop} samples, got {template.shape[-1]}")
    try:
        ctx = context_frames(gen) if context is None else int(context)
    except TypeError:
        return _call(gen, mel, template)  # no receptive-field model (RefineGAN): one forward
    if T <= chunk_frames + 2 * ctx:
        return _call(gen, mel, template)
    out = torch.empty(B, 1, T * hop, dtype=torch.float32, device=mel.device)
    for start in range(0, T, chunk_frames):
        end = min(T, start + chunk_frames)
        lo, hi = max(0, start - ctx), min(T, end + ctx)
        tpl = None if template is None else template[:, :, lo * hop:hi * hop].contiguous()
        y = _call(gen, mel[:, :, lo:hi].contiguous(), tpl)
        out[:, :, start * hop:end * hop] = y[:, :, (start - lo) * hop:(end - lo) * hop]
    return out


# ------------------------------------------------------------------------------------------------
# I/O + CLI
# ------------------------------------------------------------------------------------------------
def load_mel(path: str, n_mels: Optional[int] = None) -> torch.Tensor:
    # mel inputs are plain tensors: never unpickle arbitrary objects from a data directory
    x = torch.from_numpy(np.load(path, allow_pickle=False)) if path.endswith(".npy") else torch.load(
        path, map_location="cpu", weights_only=True)
    x = x.to(torch.float32)
    if x.ndim == 2:
        x = x[None]
    if n_mels is not None and x.shape[-1] == n_mels and x.shape[1] != n_mels:
        x = x.transpose(1, 2)  # test.py:81-82
    return x.contiguous()


def write_wav(path: str, wav: torch.Tensor, sample_rate: int) -> None:
    """wav [C, L] float in [-1, 1] -> 16-bit PCM (channels = batch entries, as test.py:98-99 writes them)."""
    pcm = (wav.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16).cpu().numpy().T.copy()
    with wave.open(path, "wb") as f:
        f.setnchannels(pcm.shape[1])
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.tobytes())


def iter_inputs(path: str) -> Iterable[str]:
    if os.path.isfile(path):
        yield path
        return
    for root, _, files in os.walk(path):
        for name in sorted(files):
            if name.endswith((".pt", ".pth", ".npy")):
                yield os.path.join(root, name)


def main(argv=None) -> int:
    import yaml

    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", required=True, help="generator yaml/json (the reference's model/generator/*.yaml)")
    ap.add_argument("--ckpt", default=None, help="Lightning .ckpt or state dict; omitted = random init")
    ap.add_argument("--input", required=True, help=".pt/.npy mel file or a directory of them")
    ap.add_argument("--output-dir", required=True)
    ap.add_argument("--sample-rate", type=int, default=44100)
    ap.add_argument("--chunk-frames", type=int, default=4096)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--template-dir", default=None,
                    help="directory of <name>.pt/.npy templates [1, T*hop] for generators built with use_template=True")
    args = ap.parse_args(argv)
    dev = torch.device(args.device)
    if dev.type == "cuda" and dev.index is not None:
        torch.cuda.set_device(dev)  # kernels launch on the current device's stream
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    gen = load_generator(cfg, args.ckpt, args.device)
    n_mels = getattr(gen, "num_mels", None) or getattr(getattr(gen, "backbone", None), "input_channels", None)
    os.makedirs(args.output_dir, exist_ok=True)
    for path in iter_inputs(args.input):
        mel = load_mel(path, n_mels).to(args.device)
        tpl = None
        if args.template_dir:
            stem = os.path.splitext(os.path.basename(path))[0]
            for ext in (".pt", ".pth", ".npy"):
                cand = os.path.join(args.template_dir, stem + ext)
                if os.path.exists(cand):
                    tpl = load_mel(cand).reshape(mel.shape[0], 1, -1).to(args.device)
                    break
            if tpl is None:
                raise FileNotFoundError(f"no template for {path} in {args.template_dir}")
        wav = chunked_forward(gen, mel, args.chunk_frames, template=tpl)
        out = os.path.join(args.output_dir, os.path.splitext(os.path.basename(path))[0] + ".wav")
        write_wav(out, wav[:, 0], args.sample_rate)
        print(f"{path} -> {out}  ({wav.shape[-1] / args.sample_rate:.2f} s)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
