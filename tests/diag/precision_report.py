#!/usr/bin/env python
"""Waveform error of every golden fixture and of full-width models at the benched shapes against the fp32 CPU oracle, per
operand precision mode (fp16 | mixed | strict).  GPU only; prints a markdown table (committed under profiles/)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import ALL_GOLDEN, build_module, channels_last_noise, load_golden, stress_init  # noqa: E402


def set_precision(m, mode):
    for sub in m.modules():
        if hasattr(sub, "_ws"):
            sub.precision = mode
    m.precision = mode


def run_fixture(name, mode):
    kwargs, sd, ins, out, extra = load_golden(name)
    m = build_module(name, kwargs)
    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()
    set_precision(m, mode)
    mel = ins["mel"].cuda()
    tpl = ins["template"].cuda() if "template" in ins else None
    with torch.no_grad():
        if name.startswith("refinegan"):
            m.noise_fn = channels_last_noise(extra["noise_seed"][0])
            y = m(mel, tpl)
        else:
            y = m(mel, tpl) if tpl is not None else m(mel)
    return float((y.cpu() - out).abs().max()), float(out.abs().max())


def main():
    modes = ("fp16", "mixed", "strict")
    print("## golden fixtures (max |delta| vs the reference's fp32 output)\n")
    print("| fixture | peak | " + " | ".join(modes) + " |")
    print("|---|---:|" + "---:|" * len(modes))
    for name in ALL_GOLDEN:
        errs, peak = [], 0.0
        for mode in modes:
            try:
                e, peak = run_fixture(name, mode)
                errs.append(f"{e:.2e}")
            except Exception as ex:  # noqa: BLE001
                errs.append(f"ERR {type(ex).__name__}: {str(ex)[:80]}")
        print(f"| {name} | {peak:.3f} | " + " | ".join(errs) + " |", flush=True)

    print("\n## full-width models at the benched shapes, B = 2 (max |delta| vs the fp32 CPU oracle)\n")
    print("| model | weights | peak | " + " | ".join(modes) + " |")
    print("|---|---|---:|" + "---:|" * len(modes))
    import bench
    torch.set_num_threads(os.cpu_count() or 8)
    for kind, wl in (("hifigan", "hifigan_b64"), ("bigvgan", "bigvgan_b32"), ("vocos", "vocos_huge_b128")):
        _, _, n_mels, T, hop, sr, _ = bench.WORKLOADS[wl]
        for weights in ("ref-init", "stress"):
            model = bench.build_model(kind).eval()
            if weights == "stress":
                stress_init(model, seed=1)
            mel = bench.synthetic_mel(2, n_mels, T, 1234)
            sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
            t0 = time.time()
            with torch.no_grad():
                want = bench.oracle_forward(kind, sd, mel, model)
            model = model.cuda()
            errs = []
            for mode in modes:
                set_precision(model, mode)
                try:
                    with torch.no_grad():
                        y = model(mel.cuda()).cpu()
                    errs.append(f"{float((y - want).abs().max()):.2e}")
                except Exception as ex:  # noqa: BLE001
                    errs.append(f"ERR {type(ex).__name__}: {str(ex)[:80]}")
            print(f"| {kind} ({wl} shape) | {weights} | {float(want.abs().max()):.3f} | " + " | ".join(errs) +
                  f" |  <!-- oracle {time.time() - t0:.1f}s -->", flush=True)
            del model
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
