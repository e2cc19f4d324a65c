"""Design evidence kept as a test: what fp16 operands would cost if the anti-aliased Snake's two FIR filters ran as banded
tensor-core GEMMs ("why the filters stay on the CUDA cores").

The shipped kernel filters the fp32 residual stream in fp32 and rounds only its OUTPUT (the next conv's operand) to fp16.  A
tcgen05 formulation (Toeplitz up-filter, activation on the accumulator, Toeplitz down-filter) needs fp16 operands, i.e. it
also rounds the kernel's INPUT (and the activated 2x signal).  This test replays the CPU oracle with those extra roundings on
the reference-generated BigVGAN fixtures: rounding the input lifts the waveform error above the 1e-3 parity bar on the stress
fixtures, which is why such a kernel would need [hi | lo] inputs (two passes) and loses its tensor-core advantage.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import generators as G
from tests.util import load_golden, oracle_forward


def _h(x):
    return x.half().float()


def _forward_with_roundings(name, mode, monkeypatch):
    orig_aa, orig_act1d = G.aa_activation, G._act1d
    state = {"post": False}

    def aa(x, act, f_up, f_down, edge_mode="replicate"):
        if state["post"] or mode == "fp32":      # activation_post feeds a strict layer: stays on the fp32 CUDA-core kernel
            return orig_aa(x, act, f_up, f_down, edge_mode)
        C, k = x.shape[1], f_up.numel()
        pad = k // 2 - 1
        xin = _h(x) if "in" in mode else x
        u = F.pad(xin, (pad, pad), mode="replicate")
        u = 2.0 * F.conv_transpose1d(u, f_up.reshape(1, 1, -1).expand(C, 1, -1), stride=2, groups=C)
        u = act(u[..., pad * 2 + (k - 2) // 2:-(pad * 2 + (k - 1) // 2)])
        if "mid" in mode:
            u = _h(u)
        kd = f_down.numel()
        u = F.pad(u, (kd // 2 - 1, kd // 2), mode="replicate")
        return _h(F.conv1d(u, f_down.reshape(1, 1, -1).expand(C, 1, -1), stride=2, groups=C))   # "out": what ships today

    def act1d(sd, prefix, x, *a, **k):
        state["post"] = prefix == "activation_post"
        try:
            return orig_act1d(sd, prefix, x, *a, **k)
        finally:
            state["post"] = False

    monkeypatch.setattr(G, "aa_activation", aa)
    monkeypatch.setattr(G, "_act1d", act1d)
    kwargs, sd, ins, out, extra = load_golden(name)
    y = oracle_forward(name, kwargs, sd, ins, extra)
    return float((y - out).abs().max()), max(1.0, float(out.abs().max()))


@pytest.mark.parametrize("name", ["bigvgan_small_stress", "bigvgan_snake_mix_stress"])
def test_fp16_input_of_the_snake_filters_breaks_the_parity_bar(name, monkeypatch):
    e_fp32, _ = _forward_with_roundings(name, "fp32", monkeypatch)
    e_out, peak = _forward_with_roundings(name, "out", monkeypatch)
    e_in_out, _ = _forward_with_roundings(name, "in+out", monkeypatch)
    e_mid_out, _ = _forward_with_roundings(name, "mid+out", monkeypatch)
    print(f"{name}: fp32 {e_fp32:.2e} | out {e_out:.2e} | in+out {e_in_out:.2e} | mid+out {e_mid_out:.2e}")
    assert e_fp32 < 1e-5                              # the patched oracle is the oracle when nothing is rounded
    assert e_out < 1e-3 * peak                        # today's operand rounding alone fits the bar ...
    assert e_in_out > 1e-3 * peak                     # ... an fp16 kernel INPUT does not
    assert e_in_out > 1.25 * e_out
    assert e_mid_out < 1.35 * e_out                   # the activated 2x signal in fp16 would have been affordable
