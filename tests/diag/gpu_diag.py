"""Kernel bring-up diagnostics (run on the B200 box):  python tests/diag/gpu_diag.py [--group NAME]

Without --group, every group runs in its own subprocess (a trapped kernel poisons the CUDA context) under a
timeout; results are merged into gpurun_out/diag.json.  References are computed on the CPU with torch.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vocoder_b200 import cabi  # noqa: E402

OUT_DIR = os.path.join(ROOT, "gpurun_out")


def ref_conv(a16, pc, L_out, bias=None, gamma=None, residual=None, scale=1.0, old=None, act=0, act_param=0.0):
    A = a16.float().cpu()
    W = pc.w.float().cpu()
    B, L_in, ap = A.shape
    kmax = min(ap, pc.w_pitch)
    r8 = cabi.round_up(pc.c_out, 8)
    out = torch.zeros(B, L_out, r8, dtype=torch.float64)
    P = pc.n_phase
    for ph in range(P):
        rows = torch.arange(ph, L_out, P)
        q = rows // P
        for tp in range(pc.n_taps):
            src = q + pc.tap_off[ph * pc.n_taps + tp]
            ok = (src >= 0) & (src < L_in)
            if ok.sum() == 0:
                continue
            contrib = A[:, src[ok], :kmax].double() @ W[ph, tp, :pc.c_out, :kmax].double().t()
            out[:, rows[ok], :pc.c_out] += contrib
    if bias is not None:
        out[..., :pc.c_out] += bias.cpu().double()
    if gamma is not None:
        out[..., :pc.c_out] *= gamma.cpu().double()
    if residual is not None:
        out += residual.cpu().double()[..., :r8]
    out = out * scale
    if old is not None:
        out += old.cpu().double()[..., :r8]
    o32 = out.float()
    if act == cabi.ACT_SILU:
        o16 = torch.nn.functional.silu(o32)
    elif act == cabi.ACT_LEAKY:
        o16 = torch.nn.functional.leaky_relu(o32, act_param)
    elif act == cabi.ACT_GELU:
        o16 = torch.nn.functional.gelu(o32)
    elif act == cabi.ACT_TANH:
        o16 = torch.tanh(o32)
    elif act == cabi.ACT_POLAR:
        m = torch.clip(torch.exp(o32[..., 0::2]), max=100.0)
        p = o32[..., 1::2]
        o16 = torch.stack([m * torch.cos(p), m * torch.sin(p)], dim=-1).reshape(o32.shape)
    else:
        o16 = o32
    return o32, o16


EPILOGUE = 0
MAINLOOP = 0


def run_conv_case(name, B, L, c_in, c_out, k=3, d=1, convT=None, act=0, use_res=False, use_gamma=False,
                  accumulate=False, scale=1.0, engines=("tc", "simt"), seed=0, time_it=False, check=True,
                  want32=True, want16=True):
    dev = "cuda"
    g = torch.Generator().manual_seed(seed)
    ap = cabi.pitch_of(c_in)
    a = torch.zeros(B, L, ap)
    a[..., :c_in] = torch.randn(B, L, c_in, generator=g)
    a16 = a.half().to(dev)
    if convT:
        kk, u = convT
        w = torch.randn(c_in, c_out, kk, generator=g) / (c_in * kk / u) ** 0.5
        bias = torch.randn(c_out, generator=g) * 0.1
        pc = cabi.pack_conv_transpose(w, bias, u).to(dev)
        L_out = cabi.conv_transpose_out_len(L, kk, u)
    else:
        w = torch.randn(c_out, c_in, k, generator=g) / (c_in * k) ** 0.5
        bias = torch.randn(c_out, generator=g) * 0.1
        pc = cabi.pack_conv(w, bias, d).to(dev)
        L_out = L
    if act == cabi.ACT_POLAR:
        pc.bias.mul_(3.0)
    r8 = cabi.round_up(c_out, 8)
    gamma = (torch.rand(c_out, generator=g) + 0.5).to(dev) if use_gamma else None
    residual = torch.randn(B, L_out, r8, generator=g).to(dev) if use_res else None
    old = torch.randn(B, L_out, r8, generator=g).to(dev) if accumulate else None
    res = {"name": name, "shape": [B, L, c_in, c_out, k, d, convT], "L_out": L_out}
    outs = {}
    for eng in engines:
        out32 = old.clone() if accumulate else (torch.full((B, L_out, r8), float("nan"), device=dev)
                                                if want32 else None)
        out16 = torch.full((B, L_out, r8), float("nan"), dtype=torch.float16, device=dev) if want16 else None
        e = cabi.ENGINE_TC if eng == "tc" else cabi.ENGINE_SIMT
        kw = dict(gamma=gamma, residual=residual, out32=out32, accumulate=accumulate, out_scale=scale,
                  out16=out16, act=act, act_param=0.2, engine=e)
        cabi.conv1d(a16, pc, L_out, **kw)
        torch.cuda.synchronize()
        outs[eng] = (None if out32 is None else out32.cpu(), None if out16 is None else out16.float().cpu())
        if time_it and eng == "tc":
            kw["accumulate"] = False
            for _ in range(3):
                cabi.conv1d(a16, pc, L_out, **kw)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            ev0.record()
            for _ in range(reps):
                cabi.conv1d(a16, pc, L_out, **kw)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / reps
            flops = 2.0 * B * L_out * c_out * c_in * (pc.n_taps if not convT else pc.n_taps)
            res["ms"] = ms
            res["tflops"] = flops / ms / 1e9
    if check:
        r32, r16 = ref_conv(a16, pc, L_out, pc.bias, gamma, residual, scale, old, act, 0.2)
        for eng in engines:
            o32, o16 = outs[eng]
            res[f"{eng}_err32"] = float((o32 - r32).abs().max())
            res[f"{eng}_err16"] = float((o16 - r16).abs().max())
            res[f"{eng}_nan"] = bool(torch.isnan(o32).any() or torch.isnan(o16).any())
        res["ref_absmax"] = float(r32.abs().max())
    elif len(engines) == 2:
        res["tc_vs_simt32"] = float((outs["tc"][0] - outs["simt"][0]).abs().max())
    return res


def group_conv_small():
    A = cabi
    cases = [
        dict(name="c64_k3", B=2, L=300, c_in=64, c_out=64, k=3, d=1, act=A.ACT_SILU),
        dict(name="c128_k7_d3_res", B=2, L=517, c_in=128, c_out=128, k=7, d=3, act=A.ACT_SILU, use_res=True),
        dict(name="c256_k11_d5", B=1, L=200, c_in=256, c_out=256, k=11, d=5, act=A.ACT_NONE, use_res=True),
        dict(name="c32_k3", B=3, L=1000, c_in=32, c_out=32, k=3, d=1, act=A.ACT_SILU),
        dict(name="c16_k7", B=2, L=700, c_in=16, c_out=16, k=7, d=1, act=A.ACT_LEAKY),
        dict(name="pre_80_512_k7", B=2, L=94, c_in=80, c_out=512, k=7, act=A.ACT_SILU),
        dict(name="pre_100_512_k7", B=2, L=87, c_in=100, c_out=512, k=7, act=A.ACT_NONE),
        dict(name="tiny_L13", B=2, L=13, c_in=20, c_out=64, k=7, act=A.ACT_SILU),
        dict(name="msub1_L128", B=2, L=128, c_in=64, c_out=128, k=3, act=A.ACT_NONE),
        dict(name="acc_scale", B=2, L=300, c_in=64, c_out=64, k=3, d=5, act=A.ACT_SILU, use_res=True,
             accumulate=True, scale=1.0 / 3.0),
    ]
    return [run_conv_case(**c) for c in cases]


def group_convT():
    A = cabi
    cases = [
        dict(name="T_512_256_k16u8", B=2, L=94, c_in=512, c_out=256, convT=(16, 8), act=A.ACT_SILU),
        dict(name="T_256_128_k16u8", B=1, L=300, c_in=256, c_out=128, convT=(16, 8), act=A.ACT_NONE),
        dict(name="T_128_64_k4u2", B=2, L=500, c_in=128, c_out=64, convT=(4, 2), act=A.ACT_SILU),
        dict(name="T_64_32_k8u2", B=2, L=300, c_in=64, c_out=32, convT=(8, 2), act=A.ACT_SILU, use_res=True),
        dict(name="T_32_16_k2u2", B=2, L=300, c_in=32, c_out=16, convT=(2, 2), act=A.ACT_NONE),
        dict(name="T_64_32_k11u5", B=1, L=77, c_in=64, c_out=32, convT=(11, 5), act=A.ACT_SILU),
        dict(name="T_64_32_k10u5_odd", B=1, L=77, c_in=64, c_out=32, convT=(10, 5), act=A.ACT_SILU),
    ]
    return [run_conv_case(**c) for c in cases]


def group_gemm():
    A = cabi
    cases = [
        dict(name="lin_352_1408_gelu", B=2, L=94, c_in=352, c_out=1408, k=1, act=A.ACT_GELU),
        dict(name="lin_1408_352_gamma_res", B=2, L=94, c_in=1408, c_out=352, k=1, act=A.ACT_NONE, use_res=True,
             use_gamma=True),
        dict(name="lin_512_1026_polar", B=2, L=50, c_in=512, c_out=1026, k=1, act=A.ACT_POLAR),
        dict(name="lin_1032_1024", B=2, L=50, c_in=1026, c_out=1024, k=1, act=A.ACT_NONE),
        dict(name="lin_2816_11264_gelu", B=1, L=300, c_in=2816, c_out=11264, k=1, act=A.ACT_GELU,
             engines=("tc",)),
        # 256-wide tiles = the CTA-pair (cta_group::2) kernel: ragged row counts (second CTA of the last pair partly or
        # fully past the end), every epilogue input, 10%-padded C_out, dilated taps crossing utterance boundaries
        dict(name="pair_lin_704_1408_res_gamma_acc", B=7, L=111, c_in=704, c_out=1408, k=1, act=A.ACT_NONE, use_res=True,
             use_gamma=True, accumulate=True, scale=0.5),
        dict(name="pair_lin_256_512_silu_rows257", B=1, L=257, c_in=256, c_out=512, k=1, act=A.ACT_SILU),
        dict(name="pair_conv_256_256_k7_d3_res", B=3, L=700, c_in=256, c_out=256, k=7, d=3, act=A.ACT_SILU, use_res=True),
        dict(name="pair_conv_512_256_k3_L384", B=2, L=384, c_in=512, c_out=256, k=3, d=1, act=A.ACT_LEAKY),
    ]
    return [run_conv_case(**c) for c in cases]


def group_conv_big():
    """many tiles per CTA: exercises the persistent loop, smem ring wrap and TMEM double buffering"""
    A = cabi
    cases = [
        dict(name="big_c128_k7", B=4, L=6016, c_in=128, c_out=128, k=7, d=3, act=A.ACT_SILU, use_res=True,
             engines=("tc", "simt"), check=False),
        dict(name="big_c32_k11", B=4, L=24064, c_in=32, c_out=32, k=11, d=5, act=A.ACT_SILU, use_res=True,
             engines=("tc", "simt"), check=False),
        dict(name="big_c256_k3", B=8, L=752, c_in=256, c_out=256, k=3, d=1, act=A.ACT_SILU, use_res=True,
             engines=("tc", "simt"), check=False),
    ]
    return [run_conv_case(**c) for c in cases]


def group_perf():
    """cfg-B (HiFiGAN, B=64) layer shapes, tensor-core engine only, timed"""
    A = cabi
    out = []
    for (C, L) in ((256, 752), (128, 6016), (64, 12032), (32, 24064)):
        for k in (3, 7, 11):
            out.append(run_conv_case(name=f"perf_c{C}_k{k}", B=64, L=L, c_in=C, c_out=C, k=k, d=1, act=A.ACT_SILU,
                                     use_res=True, engines=("tc",), check=False, time_it=True))
    out.append(run_conv_case(name="perf_T_512_256", B=64, L=94, c_in=512, c_out=256, convT=(16, 8), act=A.ACT_SILU,
                             engines=("tc",), check=False, time_it=True))
    out.append(run_conv_case(name="perf_lin_1408_5632", B=128, L=94, c_in=1408, c_out=5632, k=1, act=A.ACT_GELU,
                             engines=("tc",), check=False, time_it=True))
    out.append(run_conv_case(name="perf_lin_5632_1408", B=128, L=94, c_in=5632, c_out=1408, k=1, act=A.ACT_NONE,
                             use_res=True, use_gamma=True, engines=("tc",), check=False, time_it=True))
    out.append(run_conv_case(name="perf_lin_2816_11264", B=128, L=94, c_in=2816, c_out=11264, k=1, act=A.ACT_GELU,
                             engines=("tc",), check=False, time_it=True))
    out.append(run_conv_case(name="perf_c128_k7_c1", B=64, L=6016, c_in=128, c_out=128, k=7, d=1, act=A.ACT_SILU,
                             use_res=False, engines=("tc",), check=False, time_it=True, want32=False))
    for msub in (1, 2):
        cabi.set_tc_tuning(0, msub, EPILOGUE, MAINLOOP)
        out.append(run_conv_case(name=f"perf_c256_k7_msub{msub}", B=64, L=752, c_in=256, c_out=256, k=7, d=1,
                                 act=A.ACT_SILU, use_res=True, engines=("tc",), check=False, time_it=True))
        out.append(run_conv_case(name=f"perf_c128_k7_msub{msub}", B=64, L=6016, c_in=128, c_out=128, k=7, d=1,
                                 act=A.ACT_SILU, use_res=True, engines=("tc",), check=False, time_it=True))
        out.append(run_conv_case(name=f"perf_lin_1408_5632_msub{msub}", B=128, L=94, c_in=1408, c_out=5632, k=1,
                                 act=A.ACT_GELU, engines=("tc",), check=False, time_it=True))
    cabi.set_tc_tuning(0, 0, EPILOGUE, MAINLOOP)
    return out


def group_simt():
    """CUDA-core kernels vs CPU torch references"""
    import torch.nn.functional as F
    sys.path.insert(0, ROOT)
    from oracle import generators as G
    dev = "cuda"
    res = []
    g = torch.Generator().manual_seed(0)
    # pack / unpack
    x = torch.randn(3, 100, 87, generator=g)
    p = cabi.pack_input(x.to(dev))
    want = torch.zeros(3, 87, 104)
    want[..., :100] = x.permute(0, 2, 1)
    res.append({"name": "pack_input", "err": float((p.float().cpu() - want.half().float()).abs().max())})
    y = torch.randn(2, 301, 40, generator=g)
    u = cabi.unpack_output(y.to(dev), 33)
    res.append({"name": "unpack_output", "err": float((u.cpu() - y[..., :33].permute(0, 2, 1)).abs().max())})
    # conv_post
    for C, k in ((32, 7), (16, 7), (64, 13), (20, 7)):
        pitch = cabi.pitch_of(C)
        a = torch.zeros(2, 500, pitch)
        a[..., :C] = torch.randn(2, 500, C, generator=g)
        a16 = a.half()
        w = torch.randn(1, C, k, generator=g) * 0.2
        b = torch.randn(1, generator=g)
        ref = torch.tanh(F.conv1d(a16.float()[..., :C].permute(0, 2, 1), w, b, padding=(k - 1) // 2))
        got = cabi.conv_post_tanh(a16.to(dev), w[0].t().contiguous().to(dev), b.to(dev), C)
        res.append({"name": f"conv_post_C{C}_k{k}", "err": float((got.cpu() - ref).abs().max())})
    # snake
    f = G.kaiser_sinc_taps()
    for (B, L, C) in ((2, 700, 32), (1, 5, 16), (2, 1, 8), (1, 16, 20), (2, 33, 256), (1, 17, 8), (1, 31, 8)):
        pitch = cabi.pitch_of(C)
        xx = torch.zeros(B, L, pitch)
        xx[..., :C] = torch.randn(B, L, C, generator=g) * 2
        al, be = torch.randn(C, generator=g) * 0.5, torch.randn(C, generator=g) * 0.5
        ref = G.aa_activation(xx[..., :C].permute(0, 2, 1), lambda v: G.snake_beta(v, al, be), f, f).permute(0, 2, 1)
        out16 = torch.full((B, L, pitch), float("nan"), dtype=torch.float16, device=dev)
        cabi.snake_aa(xx.to(dev), out16, al.to(dev), be.to(dev), f.tolist(), f.tolist(), C)
        got = out16.float().cpu()
        res.append({"name": f"snake_B{B}_L{L}_C{C}", "err": float((got[..., :C] - ref).abs().max()),
                    "pad_ok": bool((got[..., C:] == 0).all()), "absmax": float(ref.abs().max())})
    # dwconv + LN, plain LN
    # (C = 20 / 36: padded channel columns; T = 21 / 9 / 30: ragged last tiles; 1408: the 704-thread pipelined configuration)
    for (B, T, C, k) in ((2, 94, 352, 7), (1, 10, 32, 7), (2, 50, 2816, 7), (2, 50, 704, 0), (2, 13, 20, 5), (1, 7, 48, 7),
                         (3, 5, 2816, 0), (2, 21, 20, 7), (1, 30, 1408, 7), (4, 9, 36, 0), (150, 17, 64, 7)):
        pitch = cabi.pitch_of(C)
        xx = torch.zeros(B, T, pitch)
        xx[..., :C] = torch.randn(B, T, C, generator=g)
        lw, lb = torch.randn(C, generator=g), torch.randn(C, generator=g)
        if k > 0:
            dw, db = torch.randn(C, 1, k, generator=g) * 0.3, torch.randn(C, generator=g) * 0.1
            h = F.conv1d(xx[..., :C].permute(0, 2, 1), dw, db, padding=k // 2, groups=C).permute(0, 2, 1)
        else:
            dw = db = None
            h = xx[..., :C]
        ref = F.layer_norm(h, (C,), lw, lb, 1e-6)
        o16 = torch.full((B, T, pitch), float("nan"), dtype=torch.float16, device=dev)
        o32 = torch.full((B, T, pitch), float("nan"), device=dev)
        cabi.dwconv_layernorm(xx.to(dev), C, None if dw is None else dw.reshape(C, k).t().contiguous().to(dev),
                              None if db is None else db.to(dev), lw.to(dev), lb.to(dev), 1e-6, k, out16=o16, out32=o32)
        res.append({"name": f"dwconv_ln_B{B}_T{T}_C{C}_k{k}", "err32": float((o32.cpu()[..., :C] - ref).abs().max()),
                    "err16": float((o16.float().cpu()[..., :C] - ref).abs().max()),
                    "pad_ok": bool((o16.float().cpu()[..., C:] == 0).all()) and bool((o32.cpu()[..., C:] == 0).all())})
    # istft ola
    for (n_fft, hop, T) in ((64, 16, 9), (1024, 256, 20)):
        K = n_fft // 2 + 1
        S = torch.randn(2, K, T, dtype=torch.complex64)
        win = torch.hann_window(n_fft)
        want = G.istft_same(S, n_fft, hop, n_fft, win)
        frames = (torch.fft.irfft(S, n_fft, dim=1) * win[None, :, None]).permute(0, 2, 1).contiguous()
        got = cabi.istft_ola(frames.to(dev), win.to(dev), n_fft, hop)
        res.append({"name": f"istft_ola_{n_fft}", "err": float((got.cpu() - want).abs().max())})
    # noise conv
    tpl = torch.randn(2, 1, 640, generator=g)
    w, b = torch.randn(24, 1, 16, generator=g), torch.randn(24, generator=g)
    ref = F.conv1d(tpl, w, b, stride=8, padding=4).permute(0, 2, 1)
    out32 = torch.empty(2, ref.shape[1], 24, device=dev)
    cabi.noise_conv(tpl[:, 0].contiguous().to(dev), w.reshape(24, 16).contiguous().to(dev), b.to(dev), out32, 24, 16, 8, 4)
    res.append({"name": "noise_conv", "err": float((out32.cpu() - ref).abs().max())})
    # resample + act_cast
    xx = torch.randn(2, 64, 16, generator=g)
    for sf in (0.5, 0.125, 2.0, 8.0):
        ref = F.interpolate(xx.permute(0, 2, 1), scale_factor=sf, mode="linear").permute(0, 2, 1)
        o = torch.empty(2, ref.shape[1], 16, device=dev)
        cabi.resample_linear(xx.to(dev), 16, ref.shape[1], 1.0 / sf, out32=o)
        res.append({"name": f"resample_{sf}", "err": float((o.cpu() - ref).abs().max())})
    nz = torch.randn(2, 64, 16, generator=g)
    nw = torch.randn(16, generator=g)
    ref = F.leaky_relu(xx + nz * nw, 0.2)
    o = torch.empty(2, 64, 16, device=dev)
    o16 = torch.zeros(2, 64, 24, dtype=torch.float16, device=dev)
    cabi.act_cast(xx.to(dev), 16, cabi.ACT_LEAKY, 0.2, noise=nz.to(dev), noise_w=nw.to(dev), out32=o, out16=o16,
                  out16_coff=8, act16=cabi.ACT_LEAKY, act16_param=0.2)
    res.append({"name": "act_cast_adain", "err": float((o.cpu() - ref).abs().max()),
                "err16": float((o16.float().cpu()[..., 8:24] - F.leaky_relu(ref, 0.2)).abs().max())})
    o2 = o.clone()
    cabi.act_cast(xx.to(dev), 16, cabi.ACT_NONE, out32=o2, out_scale=0.5, accumulate=True)
    res.append({"name": "act_cast_accumulate", "err": float((o2.cpu() - (ref + 0.5 * xx)).abs().max())})
    ref = F.interpolate(F.leaky_relu(xx, 0.2).permute(0, 2, 1), scale_factor=2.0, mode="linear").permute(0, 2, 1)
    o = torch.empty(2, 128, 16, device=dev)
    cabi.resample_linear(xx.to(dev), 16, 128, 0.5, pre_act=cabi.ACT_LEAKY, pre_param=0.2, out32=o)
    res.append({"name": "resample_preact", "err": float((o.cpu() - ref).abs().max())})
    torch.cuda.synchronize()
    return res


def group_probe():
    """row-shifted UMMA descriptor probe (see csrc/fv_debug.cu)"""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(144, 64, generator=g).half()
    w = torch.randn(64, 64, generator=g).half()
    out = torch.full((12, 2, 128, 64), float("nan"), device="cuda")
    a_d, w_d = a.cuda(), w.cuda()  # keep the device copies alive across the launch
    rc = cabi.lib().fv_debug_rowshift_probe(a_d.data_ptr(), w_d.data_ptr(), out.data_ptr(), None)
    torch.cuda.synchronize()
    res = [{"rc": rc}]
    o = out.cpu()
    for r in range(12):
        ref = a[r:r + 128].float() @ w.float().t()
        res.append({"shift": r, "err_base0": float((o[r, 0] - ref).abs().max()),
                    "err_base_r": float((o[r, 1] - ref).abs().max())})
    return res


def group_perf2():
    """epilogue / pipeline floor experiments on the HiFiGAN stage-2 shape (B=64, L=6016, C=128)"""
    A = cabi
    out = []
    kw = dict(B=64, L=6016, c_in=128, c_out=128, engines=("tc",), check=False, time_it=True)
    out.append(run_conv_case(name="c1_k1_none", k=1, act=A.ACT_NONE, want32=False, **kw))
    out.append(run_conv_case(name="c1_k1_silu", k=1, act=A.ACT_SILU, want32=False, **kw))
    out.append(run_conv_case(name="c1_k3_none", k=3, act=A.ACT_NONE, want32=False, **kw))
    out.append(run_conv_case(name="c1_k3_silu", k=3, act=A.ACT_SILU, want32=False, **kw))
    out.append(run_conv_case(name="c1_k7_none", k=7, act=A.ACT_NONE, want32=False, **kw))
    out.append(run_conv_case(name="c1_k7_silu", k=7, act=A.ACT_SILU, want32=False, **kw))
    out.append(run_conv_case(name="o32only_k3", k=3, act=A.ACT_NONE, want32=True, want16=False, **kw))
    out.append(run_conv_case(name="c2_k3_res_silu", k=3, act=A.ACT_SILU, use_res=True, **kw))
    for msub in (1, 2):
        cabi.set_tc_tuning(0, msub, EPILOGUE, MAINLOOP)
        out.append(run_conv_case(name=f"c1_k3_silu_msub{msub}", k=3, act=A.ACT_SILU, want32=False, **kw))
    cabi.set_tc_tuning(0, 0, EPILOGUE, MAINLOOP)
    return out


GROUPS = {"perf2": group_perf2, "probe": group_probe, "simt": group_simt, "conv_small": group_conv_small, "convT": group_convT, "gemm": group_gemm,
          "conv_big": group_conv_big, "perf": group_perf}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--only", default=None, help="comma separated group names (name@E runs with epilogue E)")
    args = ap.parse_args()
    os.makedirs(OUT_DIR, exist_ok=True)
    if args.group:
        global EPILOGUE, MAINLOOP
        gname = args.group
        if "@" in gname:  # name@E or name@E.M  (epilogue flavour, mainloop flavour)
            gname, e = gname.split("@")
            if "." in e:
                e, m = e.split(".")
                MAINLOOP = int(m)
            EPILOGUE = int(e)
            cabi.set_tc_tuning(0, 0, EPILOGUE, MAINLOOP)
        t0 = time.time()
        try:
            out = {"ok": True, "results": GROUPS[gname]()}
        except Exception as e:  # noqa: BLE001
            import traceback
            out = {"ok": False, "error": repr(e), "trace": traceback.format_exc()}
        out["seconds"] = time.time() - t0
        with open(os.path.join(OUT_DIR, f"diag_{args.group.replace('@', '_e').replace('.', '_m')}.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out, indent=1))
        return
    names = args.only.split(",") if args.only else list(GROUPS)
    summary = {}
    for name in names:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", name], timeout=args.timeout,
                               capture_output=True, text=True)
            path = os.path.join(OUT_DIR, f"diag_{name.replace('@', '_e').replace('.', '_m')}.json")
            if os.path.exists(path):
                summary[name] = json.load(open(path))
            else:
                summary[name] = {"ok": False, "error": "no result file"}
            summary[name]["rc"] = p.returncode
            summary[name]["stderr_tail"] = p.stderr[-1500:]
            if not summary[name].get("ok"):
                summary[name]["stdout_tail"] = p.stdout[-1500:]
        except subprocess.TimeoutExpired:
            summary[name] = {"ok": False, "error": "timeout"}
    with open(os.path.join(OUT_DIR, "diag.json"), "w") as f:
        json.dump(summary, f, indent=1)
    for name, s in summary.items():
        print("=====", name, "ok" if s.get("ok") else "FAILED", s.get("error", ""))
        for r in s.get("results", []):
            print("   ", json.dumps(r))
        if not s.get("ok"):
            print(s.get("trace", ""), s.get("stderr_tail", ""))


if __name__ == "__main__":
    main()
