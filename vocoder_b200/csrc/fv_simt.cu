// fv_simt.cu - CUDA-core kernels of the generator forward: layout entry/exit, conv_post(+tanh), the fused
// anti-aliased Snake, depthwise-conv + LayerNorm, ISTFT overlap-add, the template (noise_convs) path, RefineGAN
// glue, and a CUDA-core implementation of fv_conv1d with semantics identical to the tcgen05 kernel (bring-up /
// cross-check engine).  All HBM-bound: coalesced along the channel axis of the channels-last layout, 16-byte
// vector accesses where the layout allows, grids sized from the problem (>= several waves of 148 SMs at the
// BASELINE shapes).
#include <cstdlib>
#include <mutex>

#include "fv_common.cuh"

namespace fv {

// ------------------------------------------------------------------------------------------------
// fv_conv1d, CUDA-core engine (one thread = one output row x one column pair)
// ------------------------------------------------------------------------------------------------
struct ConvSimtParams {
  const __half* a;
  const __half* w;
  int B, L_in, a_pitch, n_phase, n_taps, C_out, C_out_r8, C_out_pad, w_pitch, L_out;
  const float* bias;
  const float* gamma;
  const float* residual;
  float* out32;
  __half* out16;
  int res_pitch, out32_pitch, out16_pitch, accumulate, act;
  float out_scale, act_param;
  int a_split, out16_split;
  int16_t tap_off[FV_MAX_TAPS];
};

__global__ void conv_simt_kernel(const __grid_constant__ ConvSimtParams p) {
  const int half_cols = p.C_out_r8 / 2;
  const long long total = (long long)p.B * p.L_out * half_cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % half_cols) * 2;
    const long long grow = idx / half_cols;
    const int b = (int)(grow / p.L_out);
    const int orow = (int)(grow % p.L_out);
    const int phase = orow % p.n_phase;
    const int q = orow / p.n_phase;
    float acc[2] = {0.f, 0.f};
    // strict precision: weight columns [Whi | Whi | Wlo] (3P) against operand columns [hi | lo | hi again]
    const int kmax = p.a_split > 0 ? 3 * p.a_split : (p.a_pitch < p.w_pitch ? p.a_pitch : p.w_pitch);
    for (int tap = 0; tap < p.n_taps; ++tap) {
      const int r = q + p.tap_off[phase * p.n_taps + tap];
      if (r < 0 || r >= p.L_in) continue;
      const __half* arow = p.a + ((size_t)b * p.L_in + r) * p.a_pitch;
      for (int e = 0; e < 2; ++e) {
        if (col + e >= p.C_out) continue;
        const __half* wrow = p.w + ((size_t)(phase * p.n_taps + tap) * p.C_out_pad + col + e) * p.w_pitch;
        float s = 0.f;
        for (int c = 0; c < kmax; ++c) {
          const int ca = (p.a_split > 0 && c >= 2 * p.a_split) ? c - 2 * p.a_split : c;
          s = fmaf(__half2float(arow[ca]), __half2float(wrow[c]), s);
        }
        acc[e] += s;
      }
    }
    float o[2];
    for (int e = 0; e < 2; ++e) {
      const int c = col + e;
      float v = acc[e] + ((p.bias && c < p.C_out) ? p.bias[c] : 0.f);
      v *= (p.gamma && c < p.C_out) ? p.gamma[c] : 1.f;
      if (p.residual) v += p.residual[grow * p.res_pitch + c];
      v *= p.out_scale;
      if (p.accumulate && p.out32) v += p.out32[grow * p.out32_pitch + c];
      o[e] = v;
    }
    if (p.out32) {
      p.out32[grow * p.out32_pitch + col] = o[0];
      p.out32[grow * p.out32_pitch + col + 1] = o[1];
    }
    if (p.out16) {
      if (p.act == FV_ACT_POLAR) {
        const float m = fminf(expf(o[0]), 100.f);
        float sn, cs;
        sincosf(o[1], &sn, &cs);
        o[0] = m * cs;
        o[1] = m * sn;
      } else {
        o[0] = act_apply(o[0], p.act, p.act_param);
        o[1] = act_apply(o[1], p.act, p.act_param);
      }
      store_half_split(p.out16 + grow * p.out16_pitch + col, o[0], p.out16_split);
      store_half_split(p.out16 + grow * p.out16_pitch + col + 1, o[1], p.out16_split);
    }
  }
}

int conv1d_simt(const fv_conv_desc* d, cudaStream_t stream) {
  ConvSimtParams p;
  memset(&p, 0, sizeof(p));
  p.a = reinterpret_cast<const __half*>(d->a);
  p.w = reinterpret_cast<const __half*>(d->w);
  p.B = d->B; p.L_in = d->L_in; p.a_pitch = d->a_pitch; p.n_phase = d->n_phase; p.n_taps = d->n_taps;
  p.C_out = d->C_out; p.C_out_r8 = round_up(d->C_out, 8); p.C_out_pad = d->C_out_pad; p.w_pitch = d->w_pitch;
  p.L_out = d->L_out; p.bias = d->bias; p.gamma = d->gamma; p.residual = d->residual; p.out32 = d->out32;
  p.out16 = reinterpret_cast<__half*>(d->out16); p.res_pitch = d->res_pitch; p.out32_pitch = d->out32_pitch;
  p.out16_pitch = d->out16_pitch; p.accumulate = d->accumulate; p.act = d->act; p.out_scale = d->out_scale;
  p.act_param = d->act_param;
  p.a_split = d->a_split;
  p.out16_split = d->out16_split;
  for (int i = 0; i < d->n_phase * d->n_taps; ++i) p.tap_off[i] = (int16_t)d->tap_off[i];
  const long long total = (long long)p.B * p.L_out * (p.C_out_r8 / 2);
  const int threads = 128;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 64) blocks = 148 * 64;
  conv_simt_kernel<<<(int)blocks, threads, 0, stream>>>(p);
  FV_CHECK_LAUNCH("conv_simt_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// layout entry / exit: [B][C][T] fp32 <-> channels-last
// ------------------------------------------------------------------------------------------------
__global__ void pack_input_kernel(const float* __restrict__ x, __half* __restrict__ out, int C, int T, int pitch,
                                  int split) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // read: t fastest
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? x[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // write: c fastest
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < pitch) store_half_split(out + ((size_t)b * T + t) * (pitch + split) + c, tile[threadIdx.x][i], split);
  }
}

__global__ void unpack_output_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int L, int pitch) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // read: c fastest
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < L && c < C) ? x[((size_t)b * L + t) * pitch + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // write: t fastest
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < L) out[((size_t)b * C + c) * L + t] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------
// conv_post (+tanh): C_out == 1.  G lanes cooperate on one output sample (channel axis split across lanes,
// 16-byte fp16 loads), partial dot products combined with warp shuffles.
// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void conv_post_kernel(const __half* __restrict__ a, const float* __restrict__ w, const float* bias,
                                 float* __restrict__ wav, int B, int L, int C, int pitch, int k, int apply_tanh,
                                 int split) {
  extern __shared__ float s_w[];  // [k][pitch]
  const int rpitch = pitch + split;  // strict precision: the row holds [hi | lo], the operand value is hi + lo
  for (int i = threadIdx.x; i < k * pitch; i += blockDim.x) {
    const int j = i / pitch, c = i % pitch;
    s_w[i] = c < C ? w[j * C + c] : 0.f;
  }
  __syncthreads();
  const int g = threadIdx.x % G;
  const int half_k = (k - 1) / 2;
  const int chunks = pitch / 8;
  const long long total = (long long)B * L;
  const int groups = blockDim.x / G;
  const long long per_iter = (long long)gridDim.x * groups;
  const long long n_iter = (total + per_iter - 1) / per_iter;
  for (long long it = 0; it < n_iter; ++it) {  // uniform trip count: every lane reaches the shuffles
    const long long s = (it * gridDim.x + blockIdx.x) * groups + threadIdx.x / G;
    const bool live = s < total;
    float acc = 0.f;
    if (live) {
      const int b = (int)(s / L), t = (int)(s % L);
      for (int j = 0; j < k; ++j) {
        const int r = t + j - half_k;
        if (r < 0 || r >= L) continue;
        const __half* row = a + ((size_t)b * L + r) * rpitch;
        for (int ci = g; ci < chunks; ci += G) {
          const uint4 pk = *reinterpret_cast<const uint4*>(row + ci * 8);
          const __half2* h2 = reinterpret_cast<const __half2*>(&pk);
          const float* wj = s_w + j * pitch + ci * 8;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            acc = fmaf(f.x, wj[2 * e], acc);
            acc = fmaf(f.y, wj[2 * e + 1], acc);
          }
          if (split > 0) {
            const uint4 pl = *reinterpret_cast<const uint4*>(row + split + ci * 8);
            const __half2* l2 = reinterpret_cast<const __half2*>(&pl);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(l2[e]);
              acc = fmaf(f.x, wj[2 * e], acc);
              acc = fmaf(f.y, wj[2 * e + 1], acc);
            }
          }
        }
      }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (live && g == 0) {
      const float v = acc + (bias ? bias[0] : 0.f);
      wav[s] = apply_tanh ? tanhf(v) : v;
    }
  }
}

// Sliding-window variant for the common shapes (k = 7 taps): a group of G lanes (one 16-byte chunk of 8 channels each)
// produces PT consecutive samples from PT + K - 1 rows read ONCE (the per-sample kernel above re-reads every row K times
// through L1), with the lane's K x 8 weights held in registers.  Interior segments load their 14 rows in two batches of
// seven unconditional 16-byte loads (all in flight before the first use: the per-row bounds test of the edge path kept the
// loads behind their branches, one latency per row: 76 us for 98 MB); SPLIT: rows hold [hi | lo] and the operand is hi + lo.
constexpr int POST_T = 8;
template <int G, int K, bool SPLIT>
__global__ void __launch_bounds__(256) conv_post_sliding_kernel(const __half* __restrict__ a, const float* __restrict__ w,
                                                                const float* bias, float* __restrict__ wav, int B, int L,
                                                                int C, int pitch, int apply_tanh) {
  constexpr int HALF = (K - 1) / 2;
  constexpr int WIN = POST_T + K - 1;
  constexpr int BATCH = (WIN + 1) / 2;
  const int rpitch = SPLIT ? 2 * pitch : pitch;
  const int g = threadIdx.x % G;
  float wr[K][8];
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      wr[j][e] = c < C ? w[j * C + c] : 0.f;
    }
  const int segs_per_b = (L + POST_T - 1) / POST_T;
  const long long total = (long long)B * segs_per_b;
  const int groups = blockDim.x / G;
  const long long per_iter = (long long)gridDim.x * groups;
  const long long n_iter = (total + per_iter - 1) / per_iter;
  const float b0 = bias ? bias[0] : 0.f;
  for (long long it = 0; it < n_iter; ++it) {  // uniform trip count: every lane reaches the shuffles
    const long long sidx = (it * gridDim.x + blockIdx.x) * groups + threadIdx.x / G;
    const bool live = sidx < total;
    float acc[POST_T];
#pragma unroll
    for (int i = 0; i < POST_T; ++i) acc[i] = 0.f;
    int b = 0, t0 = 0;
    auto consume = [&](int i, const uint4& pk, const uint4& pl) {  // window row i is tap j of local output i - j
      const __half2* h2 = reinterpret_cast<const __half2*>(&pk);
      const __half2* l2 = reinterpret_cast<const __half2*>(&pl);
      float x[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = __half22float2(h2[e]);
        if (SPLIT) {
          const float2 fl = __half22float2(l2[e]);
          f.x += fl.x;
          f.y += fl.y;
        }
        x[2 * e] = f.x;
        x[2 * e + 1] = f.y;
      }
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int o = i - j;
        if (o >= 0 && o < POST_T) {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[o] = fmaf(x[e], wr[j][e], acc[o]);
        }
      }
    };
    if (live) {
      b = (int)(sidx / segs_per_b);
      t0 = (int)(sidx % segs_per_b) * POST_T;
      const __half* base = a + (size_t)b * L * rpitch + g * 8;
      if (t0 - HALF >= 0 && t0 - HALF + WIN <= L) {
        const __half* row0 = base + (size_t)(t0 - HALF) * rpitch;
#pragma unroll
        for (int i0 = 0; i0 < WIN; i0 += BATCH) {
          uint4 pk[BATCH], pl[SPLIT ? BATCH : 1];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            if (i0 + u < WIN) {
              pk[u] = *reinterpret_cast<const uint4*>(row0 + (size_t)(i0 + u) * rpitch);
              if (SPLIT) pl[u] = *reinterpret_cast<const uint4*>(row0 + (size_t)(i0 + u) * rpitch + pitch);
            }
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u)
            if (i0 + u < WIN) consume(i0 + u, pk[u], pl[SPLIT ? u : 0]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          const int r = t0 - HALF + i;
          if (r < 0 || r >= L) continue;
          const uint4 pk = *reinterpret_cast<const uint4*>(base + (size_t)r * rpitch);
          uint4 pl = make_uint4(0u, 0u, 0u, 0u);
          if (SPLIT) pl = *reinterpret_cast<const uint4*>(base + (size_t)r * rpitch + pitch);
          consume(i, pk, pl);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < POST_T; ++i) {
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    }
    if (live && g == 0) {
#pragma unroll
      for (int i = 0; i < POST_T; ++i) {
        if (t0 + i < L) {
          const float v = acc[i] + b0;
          wav[(size_t)b * L + t0 + i] = apply_tanh ? tanhf(v) : v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// anti-aliased Snake: up 2x (polyphase 6+6 taps) -> x + sin^2(a x)/(b+1e-9) -> down 2x (12 taps), one pass.
// One thread = one channel x SN_T consecutive time steps; the activated 2x signal lives only in registers.
// ------------------------------------------------------------------------------------------------
constexpr int SN_BLOCKS = 2;  // resident 256-thread blocks per SM (118 registers; 3 blocks = 80 registers spill: 3.9 -> 4.7 ms)
constexpr int SN_SEG_MIN = 36, SN_SEG_MAX = 96;  // time steps per thread: a multiple of 6 (period of the register rings)
struct SnakeFilt {
  float up[12];
  float dn[12];
};

// Packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2): one issue slot for two channels.  ptxas encodes a {f, f} pair built
// from a kernel parameter as a scalar-broadcast uniform-register operand (FFMA2 R, R.F32x2, UR.F32, R.F32x2), so the 24
// filter taps cost no vector registers.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 bc2(float f) { return make_float2(f, f); }
// sin.approx.ftz: the non-ftz form pays an FSEL + ISETP pair per sine for subnormal arguments, whose sine is the argument
// itself to 2^-126 either way
__device__ __forceinline__ float sin_fast(float t) {
  float r;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}

// x + sin^2(a x) / b for two channels.  sin.approx = one multiply by 1/(2 pi) + MUFU.SIN, which reduces the argument
// itself: absolute error ~2^-21 inside [-2 pi, 2 pi] and ~1.2e-7 |t| beyond (rounding of the product), i.e. <= 2.5e-5 for
// the |alpha * u| <= 200 a Snake sees - two orders below the parity budget and below the fp16 rounding of the output.
__device__ __forceinline__ float2 snake_fn2(float2 u, float2 a, float2 inv_b) {
  const float2 t = fmul2(u, a);
  const float2 sn = make_float2(sin_fast(t.x), sin_fast(t.y));
  return ffma2(fmul2(inv_b, sn), sn, u);
}

// fp16 pair store; in strict mode (split > 0) also the residual pair fp16(v - hi) `split` halfs further on
__device__ __forceinline__ void store_half2_split(__half* dst, float2 v, int split) {
  const uint32_t h = pack_half2_sat(v.x, v.y);
  *reinterpret_cast<uint32_t*>(dst) = h;
  if (split > 0) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    *reinterpret_cast<uint32_t*>(dst + split) = pack_half2_sat(v.x - f.x, v.y - f.y);
  }
}

// Streaming formulation (Appendix B4 of SURVEY.md).  With x~ the edge-replicated input and v~ the edge-replicated
// activated 2x signal:   out[t] = sum_j f[j] * v~[2t - 5 + j],  and both new values a step needs,
//   v[2t+7] = act(2 * sum_q f[2q]   * x~[t+6-q]),   v[2t+8] = act(2 * sum_q f[2q+1] * x~[t+6-q]),
// read the same 6-sample window.  A thread owns TWO adjacent channels (packed fp32x2 math) and marches over seg_len steps
// keeping x~[t+1..t+6] in a 6-slot ring and v~[2t-5..2t+6] in a 12-slot ring (unrolled by 6 so ring slots are
// compile-time registers): per sample pair one 8-byte load, 24 FFMA2, 4 SFU sines and one 4-byte store; six "pre-roll"
// steps fill the rings.  Loads run one 6-step window ahead of the math (software pipeline).
// EDGE = false is the interior fast path (no index clamps, no replicate logic, pointer increments only).
// RING = true: the look-ahead loads go through a per-thread shared-memory ring filled by cp.async (18 slots of 8 bytes, two
// 6-step windows in flight) instead of 12 prefetch registers: deeper latency cover with fewer registers, so three blocks
// fit an SM.  `ring` = this thread's slot 0 (slot stride = blockDim.x float2, i.e. every warp access is 256 contiguous bytes).
template <bool EDGE, bool SPLIT, int RING = 0>
__device__ __forceinline__ void snake_segment(const float* __restrict__ xc, __half* __restrict__ orow,
                                              const SnakeFilt& f, float2 a, float2 inv_b, int t0, int t_end, int L,
                                              int pitch, int opitch, int split, float2* ring = nullptr,
                                              int ring_stride = 0) {
  const int nl = 2 * L - 1;
  auto ldx = [&](int t) {
    if (EDGE) t = t < 0 ? 0 : (t > L - 1 ? L - 1 : t);  // replicate edge of the INPUT (first filter pads its own input)
    return *reinterpret_cast<const float2*>(xc + (size_t)t * pitch);
  };
  // ring: window w (steps t0 + 6 w + 6 .. + 11) lives in slots (w % (RING + 1)) * 6 + k; RING windows are in flight
  constexpr int RW = RING > 0 ? RING + 1 : 1;
  auto ring_issue = [&](int w) {
    if constexpr (RING > 0) {
      const int tw = t0 + 6 * w + 6;
      if (tw < t_end + 6) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          int t = tw + k;
          if (EDGE) t = t < 0 ? 0 : (t > L - 1 ? L - 1 : t);  // (interior segments: only needed windows are issued)
          const uint32_t dst = smem_u32(ring + (size_t)((w % RW) * 6 + k) * ring_stride);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(xc + (size_t)t * pitch) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  const float2 zero2 = make_float2(0.f, 0.f);
  float2 X[6];   // ring: x~[t+6-q] lives in X[(k - q) mod 6] at unrolled position k
  float2 V[12];  // ring: v~[2t-5+j] lives in V[(2k + j) mod 12]
  float2 vlast = zero2;
  // every load of the pre-roll is issued before the first use:
  // window before the first pre-roll step (t = t0 - 6): x~[t+1 .. t+5] = x~[t0-5 .. t0-1] in X[1..5], then x~[t0 .. t0+5]
  float2 xp[6];
#pragma unroll
  for (int i = 1; i < 6; ++i) X[i] = ldx(t0 - 6 + i);
#pragma unroll
  for (int k = 0; k < 6; ++k) xp[k] = ldx(t0 + k);
  // ... and so is the first main-loop window x~[t0+6 .. t0+11]
  float2 xn[6];
  const float* px = xc + (size_t)(t0 + 6) * pitch;  // interior path: running pointers instead of index math
  if constexpr (RING > 0) {
#pragma unroll
    for (int w0 = 0; w0 < RING; ++w0) ring_issue(w0);
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (EDGE) xn[k] = ldx(t0 + k + 6);
      else xn[k] = *reinterpret_cast<const float2*>(px + (size_t)k * pitch);
    }
    px += (size_t)6 * pitch;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {  // pre-roll: t = t0 - 6 + k produces v[2t0-5+2k], v[2t0-4+2k]
    const int t = t0 - 6 + k;
    X[k] = xp[k];
    float2 uo = zero2, ue = zero2;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const float2 xv = X[(k - q + 6) % 6];
      uo = ffma2(xv, bc2(f.up[2 * q]), uo);
      ue = ffma2(xv, bc2(f.up[2 * q + 1]), ue);
    }
    float2 va = snake_fn2(uo, a, inv_b), vb = snake_fn2(ue, a, inv_b);  // f.up carries the 2x gain
    if (EDGE) {
      if (2 * t + 7 <= nl) vlast = va; else va = vlast;
      if (2 * t + 8 <= nl) vlast = vb; else vb = vlast;
    }
    V[(2 * k) % 12] = va;
    V[(2 * k + 1) % 12] = vb;
  }
  if (EDGE && t0 == 0) {  // v~[n < 0] = v[0]: replicate edge of the ACTIVATED signal (second filter pads its own input)
#pragma unroll
    for (int j = 0; j < 5; ++j) V[j] = V[5];
  }
  __half* po = orow + (size_t)t0 * opitch;
  int w = 0;
  for (int tb = t0; tb < t_end; tb += 6, ++w) {
    float2 xw[6];
    if constexpr (RING > 0) {
      asm volatile("cp.async.wait_group %0;" ::"n"(RING - 1) : "memory");  // window w has landed (later ones may be in flight)
#pragma unroll
      for (int k = 0; k < 6; ++k) xw[k] = ring[(size_t)((w % RW) * 6 + k) * ring_stride];
      ring_issue(w + RING);  // into the slots window w - 1 was read from in the previous iteration
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) xw[k] = xn[k];
      if (tb + 6 < t_end) {  // next window: six independent loads in flight under this window's math
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          if (EDGE) xn[k] = ldx(tb + k + 12);
          else xn[k] = *reinterpret_cast<const float2*>(px + (size_t)k * pitch);
        }
        px += (size_t)6 * pitch;
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int t = tb + k;
      // two independent partial sums (even / odd taps): dependent FFMA2 chains of 6 instead of 12
      float2 oe = fmul2(V[(2 * k) % 12], bc2(f.dn[0])), oo = fmul2(V[(2 * k + 1) % 12], bc2(f.dn[1]));
#pragma unroll
      for (int j = 2; j < 12; j += 2) {
        oe = ffma2(V[(2 * k + j) % 12], bc2(f.dn[j]), oe);
        oo = ffma2(V[(2 * k + j + 1) % 12], bc2(f.dn[j + 1]), oo);
      }
      const float2 o = fadd2(oe, oo);
      if (!EDGE || t < t_end) store_half2_split(po + (size_t)k * opitch, o, SPLIT ? split : 0);
      X[k] = xw[k];
      float2 uo = zero2, ue = zero2;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const float2 xv = X[(k - q + 6) % 6];
        uo = ffma2(xv, bc2(f.up[2 * q]), uo);
        ue = ffma2(xv, bc2(f.up[2 * q + 1]), ue);
      }
      float2 va = snake_fn2(uo, a, inv_b), vb = snake_fn2(ue, a, inv_b);  // f.up carries the 2x gain
      if (EDGE) {
        if (2 * t + 7 <= nl) vlast = va; else va = vlast;
        if (2 * t + 8 <= nl) vlast = vb; else vb = vlast;
      }
      V[(2 * k) % 12] = va;
      V[(2 * k + 1) % 12] = vb;
    }
    po += (size_t)6 * opitch;
  }
}

// one thread = one channel PAIR x seg_len time steps (pitch is a multiple of 8, so pairs never straddle a row)
template <bool SPLIT>
__global__ void __launch_bounds__(256, SN_BLOCKS) snake_aa_kernel(const float* __restrict__ x, __half* __restrict__ out,
                                                          const float* __restrict__ alpha,
                                                          const float* __restrict__ beta, const SnakeFilt f,
                                                          int logscale, int B, int L, int C, int pitch, int n_seg,
                                                          int seg_len, int split, int reverse) {
  pdl_launch_dependents();
  pdl_wait();
  const int hp = pitch >> 1;
  const long long total = (long long)B * n_seg * hp;
  long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (item >= total) return;
  if (reverse) item = total - 1 - item;   // start with what the producer wrote last (next_tile_direction)
  const int c = 2 * (int)(item % hp);
  const int seg = (int)((item / hp) % n_seg);
  const int b = (int)(item / ((long long)hp * n_seg));
  const int t0 = seg * seg_len;
  const int t_end = min(t0 + seg_len, L);
  const int opitch = pitch + split;
  __half* orow = out + ((size_t)b * L) * opitch + c;
  if (c >= C) {  // padded channels stay zero
    for (int t = t0; t < t_end; ++t) store_half2_split(orow + (size_t)t * opitch, make_float2(0.f, 0.f), split);
    return;
  }
  // a padded odd channel (c + 1 == C) reads x = 0 and therefore writes 0 whatever its parameters
  const int c1 = c + 1 < C ? c + 1 : c;
  float2 a = make_float2(alpha[c], alpha[c1]);
  float2 bb = beta ? make_float2(beta[c], beta[c1]) : a;
  if (logscale) {
    a = make_float2(expf(a.x), expf(a.y));
    bb = make_float2(expf(bb.x), expf(bb.y));
  }
  const float2 inv_b = make_float2(1.0f / (bb.x + 1e-9f), 1.0f / (bb.y + 1e-9f));
  const float* xc = x + ((size_t)b * L) * pitch + c;
  // interior segment: every x index in [t0-5, t0+seg_len+5] and every v index up to 2(t0+seg_len)+6 is in range
  const bool interior = (t0 >= 6) && (t0 + seg_len + 6 <= L - 1);
  if (interior) snake_segment<false, SPLIT>(xc, orow, f, a, inv_b, t0, t_end, L, pitch, opitch, split);
  else snake_segment<true, SPLIT>(xc, orow, f, a, inv_b, t0, t_end, L, pitch, opitch, split);
}

// The same kernel with the look-ahead in a shared-memory ring (see snake_segment<..., RING = true>): RB = 3 blocks per SM
// (80 registers) or RB = 2 (the register budget of the default kernel, only the deeper look-ahead differs).
template <int RB, int RW>   // RW = windows in flight: (RW + 1) x 6 ring slots of 8 bytes per thread
__global__ void __launch_bounds__(256, RB)
    snake_aa_ring_kernel(const float* __restrict__ x, __half* __restrict__ out, const float* __restrict__ alpha,
                         const float* __restrict__ beta, const SnakeFilt f, int logscale, int B, int L, int C, int pitch,
                         int n_seg, int seg_len, int reverse) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float2 s_ring[];  // [(RW + 1) * 6][256]
  const int hp = pitch >> 1;
  const long long total = (long long)B * n_seg * hp;
  long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (item >= total) return;
  if (reverse) item = total - 1 - item;   // start with what the producer wrote last (next_tile_direction)
  const int c = 2 * (int)(item % hp);
  const int seg = (int)((item / hp) % n_seg);
  const int b = (int)(item / ((long long)hp * n_seg));
  const int t0 = seg * seg_len;
  const int t_end = min(t0 + seg_len, L);
  __half* orow = out + ((size_t)b * L) * pitch + c;
  if (c >= C) {  // padded channels stay zero
    for (int t = t0; t < t_end; ++t) store_half2_split(orow + (size_t)t * pitch, make_float2(0.f, 0.f), 0);
    return;
  }
  const int c1 = c + 1 < C ? c + 1 : c;
  float2 a = make_float2(alpha[c], alpha[c1]);
  float2 bb = beta ? make_float2(beta[c], beta[c1]) : a;
  if (logscale) {
    a = make_float2(expf(a.x), expf(a.y));
    bb = make_float2(expf(bb.x), expf(bb.y));
  }
  const float2 inv_b = make_float2(1.0f / (bb.x + 1e-9f), 1.0f / (bb.y + 1e-9f));
  const float* xc = x + ((size_t)b * L) * pitch + c;
  float2* ring = s_ring + threadIdx.x;
  const bool interior = (t0 >= 6) && (t0 + seg_len + 6 <= L - 1);
  if (interior) snake_segment<false, false, RW>(xc, orow, f, a, inv_b, t0, t_end, L, pitch, pitch, 0, ring, 256);
  else snake_segment<true, false, RW>(xc, orow, f, a, inv_b, t0, t_end, L, pitch, pitch, 0, ring, 256);
}

// Edge modes other than `replicate` (FV_EDGE_REFLECT / FV_EDGE_ZERO: what another release of alias_free_torch may pad
// its two filters with, SURVEY 8c "unresolved ambiguity"): a direct, non-streaming evaluation, one thread per output
// sample pair.  3x the arithmetic of the streaming kernel; never on the default path.
//   out[t] = sum_j dn[j] * v~[2t - 5 + j],   v~ = v padded by the edge mode over [-5, 2L + 5]
//   v[n]   = act(sum_q up[2q + (n even)] * x~[(n + 5 - (n even)) / 2 - q]),   x~ = x padded over [-5, L + 4]
__device__ __forceinline__ int edge_index(int i, int n, int mode, bool& zero) {
  zero = false;
  if (i >= 0 && i < n) return i;
  if (mode == FV_EDGE_ZERO) {
    zero = true;
    return 0;
  }
  if (mode == FV_EDGE_REFLECT) {  // F.pad(mode="reflect"): the edge sample itself is not repeated
    i = i < 0 ? -i : 2 * (n - 1) - i;
    return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
  }
  return i < 0 ? 0 : n - 1;
}

__global__ void __launch_bounds__(256) snake_aa_direct_kernel(const float* __restrict__ x, __half* __restrict__ out,
                                                              const float* __restrict__ alpha,
                                                              const float* __restrict__ beta, const SnakeFilt f,
                                                              int logscale, int B, int L, int C, int pitch, int split,
                                                              int mode) {
  const int hp = pitch >> 1;
  const long long total = (long long)B * L * hp;
  const long long item = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (item >= total) return;
  const int c = 2 * (int)(item % hp);
  const int t = (int)((item / hp) % L);
  const int b = (int)(item / ((long long)hp * L));
  const int opitch = pitch + split;
  __half* o = out + ((size_t)b * L + t) * opitch + c;
  if (c >= C) {
    store_half2_split(o, make_float2(0.f, 0.f), split);
    return;
  }
  const int c1 = c + 1 < C ? c + 1 : c;
  float2 a = make_float2(alpha[c], alpha[c1]);
  float2 bb = beta ? make_float2(beta[c], beta[c1]) : a;
  if (logscale) {
    a = make_float2(expf(a.x), expf(a.y));
    bb = make_float2(expf(bb.x), expf(bb.y));
  }
  const float2 inv_b = make_float2(1.0f / (bb.x + 1e-9f), 1.0f / (bb.y + 1e-9f));
  const float* xc = x + ((size_t)b * L) * pitch + c;
  const int nl = 2 * L;
  float2 acc = make_float2(0.f, 0.f);
  for (int j = 0; j < 12; ++j) {
    bool vz;
    const int n = edge_index(2 * t - 5 + j, nl, mode, vz);
    if (vz) continue;
    const int even = (n & 1) ? 0 : 1;
    const int base = (n + 5 - even) / 2;
    float2 u = make_float2(0.f, 0.f);
    for (int q = 0; q < 6; ++q) {
      bool xz;
      const int xi = edge_index(base - q, L, mode, xz);
      if (xz) continue;
      const float2 xv = *reinterpret_cast<const float2*>(xc + (size_t)xi * pitch);
      const float w = f.up[2 * q + even];
      u.x = fmaf(xv.x, w, u.x);
      u.y = fmaf(xv.y, w, u.y);
    }
    const float2 v = snake_fn2(u, a, inv_b);
    acc.x = fmaf(v.x, f.dn[j], acc.x);
    acc.y = fmaf(v.y, f.dn[j], acc.y);
  }
  store_half2_split(o, acc, split);
}

// ------------------------------------------------------------------------------------------------
// depthwise conv (k taps, zero pad) + LayerNorm over C, one warp per (b, t) row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dwconv_ln_kernel(const float* __restrict__ x, __half* __restrict__ out16,
                                                        float* __restrict__ out32, const float* __restrict__ dw_w,
                                                        const float* __restrict__ dw_b, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, float eps, int B, int T, int C,
                                                        int pitch, int k, int split) {
  extern __shared__ float s_h[];  // [warps][pitch]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + warp;
  if (row >= (long long)B * T) return;
  const int b = (int)(row / T), t = (int)(row % T);
  float* h = s_h + warp * pitch;
  const float* xb = x + (size_t)b * T * pitch;
  const int half_k = (k - 1) / 2;
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) {
    float acc;
    if (k > 0) {
      acc = dw_b[c];
      for (int j = 0; j < k; ++j) {
        const int r = t + j - half_k;
        if (r >= 0 && r < T) acc = fmaf(dw_w[(size_t)j * C + c], xb[(size_t)r * pitch + c], acc);
      }
    } else {
      acc = xb[(size_t)t * pitch + c];
    }
    h[c] = acc;
    sum += acc;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum / C;
  float var = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float dlt = h[c] - mean;
    var = fmaf(dlt, dlt, var);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
  const float rstd = 1.0f / sqrtf(var / C + eps);
  for (int c = lane; c < pitch; c += 32) {
    const float y = c < C ? fmaf((h[c] - mean) * rstd, ln_w[c], ln_b[c]) : 0.f;
    if (out16) store_half_split(out16 + (size_t)row * (pitch + split) + c, y, split);
    if (out32) out32[(size_t)row * pitch + c] = y;
  }
}

// Tiled version for the ConvNeXt shapes: one CTA = DW_R consecutive time steps x all channels.  Each thread walks
// channels c = tid, tid + 256, ...: the DW_R + K - 1 input rows and the K taps (weights stored [k][C]) are read once
// per channel with coalesced 128-byte warp accesses (2.5 loads per output instead of 7), the depthwise results stay in
// shared memory for the two LayerNorm reductions and the normalised write.  K = 0: plain LayerNorm over C.
constexpr int DW_R = 4;

template <int K>
__global__ void __launch_bounds__(256) dwconv_ln_tiled_kernel(const float* __restrict__ x, __half* __restrict__ out16,
                                                              float* __restrict__ out32,
                                                              const float* __restrict__ dw_wT,
                                                              const float* __restrict__ dw_b,
                                                              const float* __restrict__ ln_w,
                                                              const float* __restrict__ ln_b, float eps, int T, int C,
                                                              int pitch, int tiles_per_b, int split) {
  extern __shared__ float s_h[];  // [DW_R][pitch]
  __shared__ float s_red[8][DW_R];
  __shared__ float s_mean[DW_R], s_rstd[DW_R];
  const int b = blockIdx.x / tiles_per_b;
  const int t0 = (blockIdx.x % tiles_per_b) * DW_R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = x + (size_t)b * T * pitch;
  constexpr int HALF = K > 0 ? (K - 1) / 2 : 0;
  constexpr int WIN = DW_R + (K > 0 ? K - 1 : 0);
  float sum[DW_R];
#pragma unroll
  for (int r = 0; r < DW_R; ++r) sum[r] = 0.f;
  for (int c = tid; c < C; c += 256) {
    float xw[WIN];
#pragma unroll
    for (int i = 0; i < WIN; ++i) {
      const int t = t0 - HALF + i;
      xw[i] = (t >= 0 && t < T) ? xb[(size_t)t * pitch + c] : 0.f;
    }
    float acc[DW_R];
    if constexpr (K > 0) {
      const float bias = dw_b[c];
#pragma unroll
      for (int r = 0; r < DW_R; ++r) acc[r] = bias;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float w = dw_wT[(size_t)j * C + c];
#pragma unroll
        for (int r = 0; r < DW_R; ++r) acc[r] = fmaf(w, xw[r + j], acc[r]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < DW_R; ++r) acc[r] = xw[r];
    }
#pragma unroll
    for (int r = 0; r < DW_R; ++r) {
      s_h[r * pitch + c] = acc[r];
      sum[r] += acc[r];
    }
  }
  // block reduction of the DW_R row sums -> means
#pragma unroll
  for (int r = 0; r < DW_R; ++r) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], off);
    if (lane == 0) s_red[warp][r] = sum[r];
  }
  __syncthreads();
  if (tid < DW_R) {
    float m = 0.f;
    for (int w = 0; w < 8; ++w) m += s_red[w][tid];
    s_mean[tid] = m / C;
  }
  __syncthreads();
  float var[DW_R];
#pragma unroll
  for (int r = 0; r < DW_R; ++r) var[r] = 0.f;
  for (int c = tid; c < C; c += 256) {
#pragma unroll
    for (int r = 0; r < DW_R; ++r) {
      const float dlt = s_h[r * pitch + c] - s_mean[r];
      var[r] = fmaf(dlt, dlt, var[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < DW_R; ++r) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) var[r] += __shfl_xor_sync(0xffffffffu, var[r], off);
    if (lane == 0) s_red[warp][r] = var[r];
  }
  __syncthreads();
  if (tid < DW_R) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += s_red[w][tid];
    s_rstd[tid] = 1.0f / sqrtf(v / C + eps);
  }
  __syncthreads();
  for (int c = tid; c < pitch; c += 256) {
    const float g = c < C ? ln_w[c] : 0.f, be = c < C ? ln_b[c] : 0.f;
#pragma unroll
    for (int r = 0; r < DW_R; ++r) {
      const int t = t0 + r;
      if (t < T) {
        const float y = c < C ? fmaf((s_h[r * pitch + c] - s_mean[r]) * s_rstd[r], g, be) : 0.f;
        const size_t o = ((size_t)b * T + t) * pitch + c;
        if (out16) store_half_split(out16 + ((size_t)b * T + t) * (pitch + split) + c, y, split);
        if (out32) out32[o] = y;
      }
    }
  }
}

// Vectorised variant: one CTA = DW8_R consecutive time steps x all channels, a thread owns groups of FOUR adjacent channels
// (16-byte loads / stores: a warp moves 512 contiguous bytes per instruction), 14 input rows per 8 outputs (1.75 loads per
// output).  The depthwise results stay in shared memory for the two LayerNorm reductions (mean, then centred variance).
constexpr int DW8_R = 8;

template <int K>
__global__ void __launch_bounds__(256, 2) dwconv_ln_vec_kernel(const float* __restrict__ x, __half* __restrict__ out16,
                                                            float* __restrict__ out32, const float* __restrict__ dw_wT,
                                                            const float* __restrict__ dw_b,
                                                            const float* __restrict__ ln_w,
                                                            const float* __restrict__ ln_b, float eps, int T, int C,
                                                            int pitch, int tiles_per_b, int split) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_h[];  // [DW8_R][pitch]
  __shared__ float s_red[8][DW8_R];
  __shared__ float s_mean[DW8_R], s_rstd[DW8_R];
  const int b = blockIdx.x / tiles_per_b;
  const int t0 = (blockIdx.x % tiles_per_b) * DW8_R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x, nwarp = nthr >> 5;
  const float* xb = x + (size_t)b * T * pitch;
  constexpr int HALF = K > 0 ? (K - 1) / 2 : 0;
  constexpr int WIN = DW8_R + (K > 0 ? K - 1 : 0);
  const int C4 = C >> 2;  // C % 4 == 0 (checked by the launcher)
  float sum[DW8_R];
#pragma unroll
  for (int r = 0; r < DW8_R; ++r) sum[r] = 0.f;
  for (int g = tid; g < C4; g += nthr) {
    const int c = g * 4;
    float4 acc[DW8_R];
    if constexpr (K > 0) {
      const float4 bias = *reinterpret_cast<const float4*>(dw_b + c);
#pragma unroll
      for (int r = 0; r < DW8_R; ++r) acc[r] = bias;
      float4 w[K];
#pragma unroll
      for (int j = 0; j < K; ++j) w[j] = *reinterpret_cast<const float4*>(dw_wT + (size_t)j * C + c);
      auto tap_row = [&](int i, const float4& xv) {  // input row i is tap j of output row r = i - j
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int r = i - j;
          if (r >= 0 && r < DW8_R) {
            acc[r].x = fmaf(w[j].x, xv.x, acc[r].x);
            acc[r].y = fmaf(w[j].y, xv.y, acc[r].y);
            acc[r].z = fmaf(w[j].z, xv.z, acc[r].z);
            acc[r].w = fmaf(w[j].w, xv.w, acc[r].w);
          }
        }
      };
      if (t0 - HALF >= 0 && t0 - HALF + WIN <= T) {
        // interior tile: two batches of seven unconditional loads, all in flight before their first use (behind the
        // per-row bounds test of the edge path every load waits for its own branch: one DRAM latency per row)
        constexpr int BATCH = (WIN + 1) / 2;
        const float* x0 = xb + (size_t)(t0 - HALF) * pitch + c;
#pragma unroll
        for (int i0 = 0; i0 < WIN; i0 += BATCH) {
          float4 xin[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u)
            if (i0 + u < WIN) xin[u] = *reinterpret_cast<const float4*>(x0 + (size_t)(i0 + u) * pitch);
#pragma unroll
          for (int u = 0; u < BATCH; ++u)
            if (i0 + u < WIN) tap_row(i0 + u, xin[u]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          const int t = t0 - HALF + i;
          float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t >= 0 && t < T) xv = *reinterpret_cast<const float4*>(xb + (size_t)t * pitch + c);
          tap_row(i, xv);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < DW8_R; ++r) {
        const int t = t0 + r;
        acc[r] = (t < T) ? *reinterpret_cast<const float4*>(xb + (size_t)t * pitch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < DW8_R; ++r) {
      *reinterpret_cast<float4*>(s_h + r * pitch + c) = acc[r];
      sum[r] += (acc[r].x + acc[r].y) + (acc[r].z + acc[r].w);
    }
  }
#pragma unroll
  for (int r = 0; r < DW8_R; ++r) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], off);
    if (lane == 0) s_red[warp][r] = sum[r];
  }
  __syncthreads();
  if (tid < DW8_R) {
    float m = 0.f;
    for (int w = 0; w < nwarp; ++w) m += s_red[w][tid];
    s_mean[tid] = m / C;
  }
  __syncthreads();
  float var[DW8_R];
#pragma unroll
  for (int r = 0; r < DW8_R; ++r) var[r] = 0.f;
  for (int g = tid; g < C4; g += nthr) {
#pragma unroll
    for (int r = 0; r < DW8_R; ++r) {
      const float4 h = *reinterpret_cast<const float4*>(s_h + r * pitch + g * 4);
      const float m = s_mean[r];
      const float d0 = h.x - m, d1 = h.y - m, d2 = h.z - m, d3 = h.w - m;
      var[r] = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, var[r]))));
    }
  }
#pragma unroll
  for (int r = 0; r < DW8_R; ++r) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) var[r] += __shfl_xor_sync(0xffffffffu, var[r], off);
    if (lane == 0) s_red[warp][r] = var[r];
  }
  __syncthreads();
  if (tid < DW8_R) {
    float v = 0.f;
    for (int w = 0; w < nwarp; ++w) v += s_red[w][tid];
    s_rstd[tid] = 1.0f / sqrtf(v / C + eps);
  }
  __syncthreads();
  const int P4 = pitch >> 2;
  for (int g = tid; g < P4; g += nthr) {
    const int c = g * 4;
    float4 gw = make_float4(0.f, 0.f, 0.f, 0.f), gb = gw;
    if (c < C) {
      gw = *reinterpret_cast<const float4*>(ln_w + c);
      gb = *reinterpret_cast<const float4*>(ln_b + c);
    }
#pragma unroll
    for (int r = 0; r < DW8_R; ++r) {
      const int t = t0 + r;
      if (t >= T) break;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        const float4 h = *reinterpret_cast<const float4*>(s_h + r * pitch + c);
        const float m = s_mean[r], rs = s_rstd[r];
        y.x = fmaf((h.x - m) * rs, gw.x, gb.x);
        y.y = fmaf((h.y - m) * rs, gw.y, gb.y);
        y.z = fmaf((h.z - m) * rs, gw.z, gb.z);
        y.w = fmaf((h.w - m) * rs, gw.w, gb.w);
      }
      const size_t row = (size_t)b * T + t;
      if (out16) {
        __half* dst = out16 + row * (pitch + split) + c;
        const uint32_t h0 = pack_half2_sat(y.x, y.y), h1 = pack_half2_sat(y.z, y.w);
        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
        if (split > 0) {
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
          *reinterpret_cast<uint2*>(dst + split) =
              make_uint2(pack_half2_sat(y.x - f0.x, y.y - f0.y), pack_half2_sat(y.z - f1.x, y.w - f1.y));
        }
      }
      if (out32) *reinterpret_cast<float4*>(out32 + row * pitch + c) = y;
    }
  }
}

// Persistent pipelined variant (round 2): one CTA per SM walks over tiles of DW8_R time steps with the input rows of the next
// two tiles always in flight.  A thread owns one PAIR of channels for the whole launch and copies exactly the columns it later
// reads - 14 rows x 8 bytes per tile with cp.async into one of two shared-memory stages, no registers held, completion by
// the thread's own cp.async group - so no barrier or mbarrier orders the stages.  (First version: ONE 79 KB cp.async.bulk per
// tile, then one per row pair: the data arrived later than two tiles of compute, ~11 spins of the mbarrier wait per tile at
// 18% DRAM utilisation - the bulk-copy engine keeps too few bytes in flight per SM for DRAM-latency streams.)
// Depthwise taps, bias and LayerNorm affine live in registers across tiles, the conv results stay in registers for both
// LayerNorm passes, all math is packed fp32x2.  Rows outside [0, T) (the conv's zero padding) are zero-filled in place.
__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// edge tiles of the pipelined kernel: rows outside [0, T) are the conv's zero padding (kept out of line: two tiles in twelve)
__device__ __noinline__ void dwln_fill_edge(uint32_t dst, const float* src, int t_first, int T, int win, uint32_t row_bytes,
                                            int pitch) {
  for (int i = 0; i < win; ++i) {
    const int t = t_first + i;
    if (t >= 0 && t < T) cp_async8(dst, src);
    else asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(dst), "f"(0.f) : "memory");
    dst += row_bytes;
    src += pitch;
  }
}

// warp-wide sums of eight per-thread values in 12 shuffles instead of 40: every step halves the number of values a lane
// still carries (lanes keep the half selected by one of their lane-id bits and send the other half to their partner);
// returns the sum of v[row] over the warp with row = 4*bit4 + 2*bit3 + bit2 of the lane id
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
  float a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = b4 ? v[4 + i] : v[i], send = b4 ? v[i] : v[4 + i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = b3 ? a[2 + i] : a[i], send = b3 ? a[i] : a[2 + i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float keep = b2 ? b[1] : b[0], send = b2 ? b[0] : b[1];
  float c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
}

// A thread owns one PAIR of channels (packed fp32x2 math, 704 threads = 22 warps for C = 1408) for the whole launch.
// MAXT = block-size bound of the instantiation (register budget).  ncu of the first version of this kernel (4 channels per
// thread, scalar math, 11 warps): 1250 instructions per thread and tile, a quarter of them integer address / bounds work,
// 0.39 IPC per scheduler - issue-bound with too few warps, not memory-bound (long-scoreboard 0.18 per issue).
template <int K, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
    dwconv_ln_pipe_kernel(const float* __restrict__ x, __half* __restrict__ out16, float* __restrict__ out32,
                          const float* __restrict__ dw_wT, const float* __restrict__ dw_b,
                          const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int T, int C,
                          int pitch, int tiles_per_b, int total_tiles, int split) {
  pdl_launch_dependents();
  constexpr int R = DW8_R;
  constexpr int HALF = K > 0 ? (K - 1) / 2 : 0;
  constexpr int WIN = R + (K > 0 ? K - 1 : 0);
  extern __shared__ __align__(128) float s_in[];  // [2][WIN][pitch]
  __shared__ float s_red[MAXT / 32][R], s_red2[MAXT / 32][R];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid * 2;                 // this thread's channel pair (c >= pitch: idle, only joins the barriers)
  const bool live = c < C, padcol = c >= C && c < pitch;
  const size_t stage_elems = (size_t)WIN * pitch;
  const uint32_t row_bytes = (uint32_t)pitch * 4u;
  const int red_row = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  for (int i = tid; i < (MAXT / 32) * R; i += (int)blockDim.x) {
    (&s_red[0][0])[i] = 0.f;
    (&s_red2[0][0])[i] = 0.f;
  }
  __syncthreads();
  pdl_wait();

  // stage fill (this thread's two columns of the tile's WIN input rows): async copies of the rows inside the sequence,
  // zeros for the rows outside it; always commits a group so that "all but the newest group" is the tile being waited for.
  // (b, tb) = batch item and tile-in-item of the tile to fetch; `ok` = that tile exists.
  const uint32_t st_base0 = smem_u32(s_in + c), st_stage = (uint32_t)(stage_elems * sizeof(float));
  auto issue = [&](bool ok, int b, int tb, int s) {
    if (live && ok) {
      const int t_first = tb * R - HALF;
      uint32_t dst = st_base0 + (uint32_t)s * st_stage;
      const float* src = x + ((size_t)b * T + t_first) * pitch + c;   // (row t_first may lie outside: never dereferenced)
      if (t_first >= 0 && t_first + WIN <= T) {
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          cp_async8(dst, src);
          dst += row_bytes;
          src += pitch;
        }
      } else {
        dwln_fill_edge(dst, src, t_first, T, WIN, row_bytes, pitch);   // first / last tile of an utterance (out of line)
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // every warp sums the per-warp partials itself (lane = (part, row): four strided partial sums per row, two shuffles), so
  // no single warp computes while the others wait at a barrier; returns the total of row (lane & 7) in every lane
  auto cross_warp = [&](const float (*part)[R]) {
    float m = 0.f;
    const float* pl = &part[lane >> 3][lane & 7];
#pragma unroll
    for (int u = 0; u < MAXT / 128; ++u) m += pl[4 * u * R];   // rows of warps that do not exist stay zero
    m += __shfl_xor_sync(0xffffffffu, m, 8);
    m += __shfl_xor_sync(0xffffffffu, m, 16);
    return m;
  };

  float2 w[K > 0 ? K : 1], bias = make_float2(0.f, 0.f), gw = bias, gb = bias;
  if (live) {
    if constexpr (K > 0) {
      bias = *reinterpret_cast<const float2*>(dw_b + c);
#pragma unroll
      for (int j = 0; j < K; ++j) w[j] = *reinterpret_cast<const float2*>(dw_wT + (size_t)j * C + c);
    }
    gw = *reinterpret_cast<const float2*>(ln_w + c);
    gb = *reinterpret_cast<const float2*>(ln_b + c);
  }
  // tiles of this CTA: blockIdx.x + k * gridDim.x; (b, tb) advance by (q, rem) with a carry instead of a division per tile
  const int stride = (int)gridDim.x;
  const int q = stride / tiles_per_b, rem = stride - q * tiles_per_b;
  int b = (int)blockIdx.x / tiles_per_b, tb = (int)blockIdx.x - b * tiles_per_b;   // the tile being computed
  int bi = b, tbi = tb, tile_i = (int)blockIdx.x;                                   // the tile being fetched
  auto advance = [&](int& bb, int& tt) {
    bb += q;
    tt += rem;
    if (tt >= tiles_per_b) {
      tt -= tiles_per_b;
      ++bb;
    }
  };
  issue(tile_i < total_tiles, bi, tbi, 0);
  advance(bi, tbi);
  tile_i += stride;
  issue(tile_i < total_tiles, bi, tbi, 1);
  const float inv_c = 1.0f / (float)C;
  const int opitch = pitch + split;
  uint32_t it = 0;
  for (int tile = (int)blockIdx.x; tile < total_tiles; tile += stride, ++it) {
    const int s = (int)(it & 1u);
    const int t0 = tb * R;
    const uint32_t st = st_base0 + (uint32_t)s * st_stage;
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this tile's copies (own columns) have landed
    float2 acc[R];
    float red[R];
    if (live) {
      uint32_t la = st;
      auto lds2 = [&]() {   // next input row of this thread's columns
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(la));
        la += row_bytes;
        return v;
      };
      if constexpr (K > 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = bias;
#pragma unroll
        for (int i = 0; i < WIN; ++i) {
          const float2 xv = lds2();
#pragma unroll
          for (int j = 0; j < K; ++j) {  // input row i is tap j of output row r = i - j
            const int r = i - j;
            if (r >= 0 && r < R) acc[r] = ffma2(w[j], xv, acc[r]);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = lds2();
      }
#pragma unroll
      for (int r = 0; r < R; ++r) red[r] = acc[r].x + acc[r].y;
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r] = make_float2(0.f, 0.f);
        red[r] = 0.f;
      }
    }
    {
      const float v = warp_reduce8(red, lane);
      if ((lane & 3) == 0) s_red[warp][red_row] = v;
    }
    advance(bi, tbi);
    tile_i += stride;
    issue(tile_i < total_tiles, bi, tbi, s);  // stage s is in registers: refill this thread's columns with the tile after next
    __syncthreads();
    const float mean_l = cross_warp(s_red) * inv_c;      // of row (lane & 7)
    float mean[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] = __shfl_sync(0xffffffffu, mean_l, r);
      const float d0 = acc[r].x - mean[r], d1 = acc[r].y - mean[r];
      red[r] = live ? fmaf(d0, d0, d1 * d1) : 0.f;
    }
    {
      const float v = warp_reduce8(red, lane);
      if ((lane & 3) == 0) s_red2[warp][red_row] = v;
    }
    __syncthreads();
    const float rstd_l = rsqrtf(cross_warp(s_red2) * inv_c + eps);
    float rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = __shfl_sync(0xffffffffu, rstd_l, r);   // (whole warp: before any lane drops out)
    if (live || padcol) {
      const int rows = T - t0 < R ? T - t0 : R;
      const size_t row0 = (size_t)b * T + t0;
      __half* o16 = out16 ? out16 + row0 * opitch + c : nullptr;
      float* o32 = out32 ? out32 + row0 * pitch + c : nullptr;
      auto put_row = [&](int r) {
        float2 y = make_float2(0.f, 0.f);
        if (live) y = ffma2(ffma2(acc[r], bc2(rstd[r]), bc2(-mean[r] * rstd[r])), gw, gb);
        if (o16) {
          const uint32_t h = pack_half2_sat(y.x, y.y);
          *reinterpret_cast<uint32_t*>(o16) = h;
          if (split > 0) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
            *reinterpret_cast<uint32_t*>(o16 + split) = pack_half2_sat(y.x - f.x, y.y - f.y);
          }
          o16 += opitch;
        }
        if (o32) {
          *reinterpret_cast<float2*>(o32) = y;
          o32 += pitch;
        }
      };
      if (rows == R) {
#pragma unroll
        for (int r = 0; r < R; ++r) put_row(r);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r < rows) put_row(r);
      }
    }
    advance(b, tb);
  }
}

// ------------------------------------------------------------------------------------------------
// ISTFT("same") overlap-add + envelope normalisation
// ------------------------------------------------------------------------------------------------
__global__ void istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                                 float* __restrict__ wav, int B, int T, int n_fft, int hop, int frame_pitch, int trim,
                                 int Lw) {
  const long long total = (long long)B * Lw;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = (int)(idx / Lw), s = (int)(idx % Lw);
  const int pos = s + trim;
  int f_hi = pos / hop;
  if (f_hi > T - 1) f_hi = T - 1;
  int f_lo = (pos - n_fft + hop) / hop;  // ceil((pos - n_fft + 1) / hop) for pos - n_fft + 1 > 0
  if (pos - n_fft + 1 <= 0) f_lo = 0;
  float acc = 0.f, env = 0.f;
  for (int f = f_lo; f <= f_hi; ++f) {
    const int i = pos - f * hop;
    if (i < 0 || i >= n_fft) continue;
    acc += frames[((size_t)b * T + f) * frame_pitch + i];
    const float wv = window[i];
    env = fmaf(wv, wv, env);
  }
  wav[idx] = acc / env;
}

// ------------------------------------------------------------------------------------------------
// template path: noise_convs[i](template), C_in == 1
// ------------------------------------------------------------------------------------------------
__global__ void noise_conv_kernel(const float* __restrict__ tpl, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ out, int B, int L_audio, int L_out,
                                  int C, int pitch, int k, int stride, int pad) {
  const long long total = (long long)B * L_out * pitch;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % pitch);
  const long long grow = idx / pitch;
  const int b = (int)(grow / L_out), t = (int)(grow % L_out);
  float acc = 0.f;
  if (c < C) {
    acc = bias[c];
    const float* tb = tpl + (size_t)b * L_audio;
    for (int j = 0; j < k; ++j) {
      const int r = t * stride - pad + j;
      if (r >= 0 && r < L_audio) acc = fmaf(w[c * k + j], tb[r], acc);
    }
  }
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------------------
// RefineGAN glue
// ------------------------------------------------------------------------------------------------
__global__ void act_cast_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                const float* __restrict__ noise_w, __half* __restrict__ out16,
                                float* __restrict__ out32, int act, float param, int act16, float param16,
                                float out_scale, int accumulate, long long rows, int C, int in_pitch,
                                int out16_pitch, int out16_coff, int out32_pitch, int split) {
  const long long total = rows * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long r = idx / C;
  float v = x[r * in_pitch + c];
  if (noise) v = fmaf(noise[r * in_pitch + c], noise_w[c], v);
  v = act_apply(v, act, param);
  if (out32) {
    float o = v * out_scale;
    if (accumulate) o += out32[r * out32_pitch + c];
    out32[r * out32_pitch + c] = o;
  }
  if (out16) store_half_split(out16 + r * out16_pitch + out16_coff + c, act_apply(v, act16, param16), split);
}

__global__ void resample_linear_kernel(const float* __restrict__ x, float* __restrict__ out32,
                                       __half* __restrict__ out16, int pre_act, float pre_param, int act, float param,
                                       int B, int L_in, int L_out, int C, int in_pitch, int out_pitch, int out_coff,
                                       float scale, int split) {
  const long long total = (long long)B * L_out * C;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const long long grow = idx / C;
  const int b = (int)(grow / L_out), t = (int)(grow % L_out);
  float src = (t + 0.5f) * scale - 0.5f;  // align_corners=False
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > L_in - 1) i0 = L_in - 1;
  const int i1 = i0 + 1 < L_in ? i0 + 1 : L_in - 1;
  const float lam = src - (float)i0;
  const float* xb = x + (size_t)b * L_in * in_pitch + c;
  const float v0 = act_apply(xb[(size_t)i0 * in_pitch], pre_act, pre_param);
  const float v1 = act_apply(xb[(size_t)i1 * in_pitch], pre_act, pre_param);
  float v = (1.0f - lam) * v0 + lam * v1;
  v = act_apply(v, act, param);
  if (out32) out32[grow * out_pitch + out_coff + c] = v;
  if (out16) store_half_split(out16 + grow * out_pitch + out_coff + c, v, split);
}

}  // namespace fv

// ================================================================================================
// C ABI
// ================================================================================================
using namespace fv;

static inline int grid1d(long long total, int threads) { return (int)((total + threads - 1) / threads); }

extern "C" int fv_pack_input(const float* x, void* out16, int B, int C, int T, int pitch, int split, void* stream) {
  FV_REQUIRE(x && out16 && B > 0 && C > 0 && T > 0 && pitch >= C && pitch % 8 == 0 && (split == 0 || split == pitch),
             FV_E_BADARG, "fv_pack_input: bad arguments (B=%d C=%d T=%d pitch=%d split=%d)", B, C, T, pitch, split);
  dim3 grid(ceil_div(T, 32), ceil_div(pitch, 32), B), block(32, 8);
  pack_input_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, (__half*)out16, C, T, pitch, split);
  FV_CHECK_LAUNCH("pack_input_kernel");
  return 0;
}

extern "C" int fv_unpack_output(const float* x32, float* out, int B, int C, int L, int pitch, void* stream) {
  FV_REQUIRE(x32 && out && B > 0 && C > 0 && L > 0 && pitch >= C, FV_E_BADARG, "fv_unpack_output: bad arguments");
  dim3 grid(ceil_div(L, 32), ceil_div(C, 32), B), block(32, 8);
  unpack_output_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x32, out, C, L, pitch);
  FV_CHECK_LAUNCH("unpack_output_kernel");
  return 0;
}

extern "C" int fv_conv_post_tanh(const void* a16, const float* w32, const float* bias, float* wav, int B, int L, int C,
                                 int pitch, int k, int apply_tanh, int split, void* stream) {
  FV_REQUIRE(a16 && w32 && wav && B > 0 && L > 0 && C > 0 && pitch >= C && pitch % 8 == 0 && k > 0 && (k & 1) &&
                 (split == 0 || split == pitch),
             FV_E_BADARG, "fv_conv_post_tanh: bad arguments (C=%d pitch=%d k=%d split=%d)", C, pitch, k, split);
  const int smem = k * pitch * (int)sizeof(float);
  FV_REQUIRE(smem <= 48 * 1024, FV_E_UNSUPPORTED, "fv_conv_post_tanh: k*pitch too large (%d bytes)", smem);
  const int chunks = pitch / 8;
  const int threads = 256;
  const long long total = (long long)B * L;
  if (k == 7 && (chunks == 1 || chunks == 2 || chunks == 4 || chunks == 8)) {
    const long long groups_total = (long long)B * ceil_div(L, POST_T);
#define FV_POST_S(G)                                                                                        \
  {                                                                                                         \
    long long blocks = (groups_total + (threads / G) - 1) / (threads / G);                                  \
    if (blocks > 148 * 8) blocks = 148 * 8;                                                                 \
    if (split)                                                                                              \
      conv_post_sliding_kernel<G, 7, true><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(              \
          (const __half*)a16, w32, bias, wav, B, L, C, pitch, apply_tanh);                                  \
    else                                                                                                    \
      conv_post_sliding_kernel<G, 7, false><<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(             \
          (const __half*)a16, w32, bias, wav, B, L, C, pitch, apply_tanh);                                  \
  }
    if (chunks == 1) FV_POST_S(1)
    else if (chunks == 2) FV_POST_S(2)
    else if (chunks == 4) FV_POST_S(4)
    else FV_POST_S(8)
#undef FV_POST_S
    FV_CHECK_LAUNCH("conv_post_sliding_kernel");
    return 0;
  }
#define FV_POST(G)                                                                                          \
  {                                                                                                         \
    long long blocks = (total + (threads / G) - 1) / (threads / G);                                         \
    if (blocks > 148 * 16) blocks = 148 * 16;                                                               \
    conv_post_kernel<G><<<(int)blocks, threads, smem, (cudaStream_t)stream>>>(                              \
        (const __half*)a16, w32, bias, wav, B, L, C, pitch, k, apply_tanh, split);                          \
  }
  if (chunks <= 1) FV_POST(1)
  else if (chunks <= 2) FV_POST(2)
  else if (chunks <= 4) FV_POST(4)
  else if (chunks <= 8) FV_POST(8)
  else if (chunks <= 16) FV_POST(16)
  else FV_POST(32)
#undef FV_POST
  FV_CHECK_LAUNCH("conv_post_kernel");
  return 0;
}

extern "C" int fv_snake_aa(const float* x32, void* out16, const float* alpha, const float* beta, const float* filt_up,
                           const float* filt_down, int logscale, int B, int L, int C, int pitch, int split,
                           int edge_mode, void* stream) {
  FV_REQUIRE(x32 && out16 && alpha && filt_up && filt_down && B > 0 && L > 0 && C > 0 && pitch >= C && pitch % 2 == 0 &&
                 (reinterpret_cast<uintptr_t>(x32) & 7) == 0 && (reinterpret_cast<uintptr_t>(out16) & 3) == 0 &&
                 (split == 0 || split == pitch),
             FV_E_BADARG, "fv_snake_aa: bad arguments");
  FV_REQUIRE(edge_mode == FV_EDGE_REPLICATE || edge_mode == FV_EDGE_REFLECT || edge_mode == FV_EDGE_ZERO, FV_E_BADARG,
             "fv_snake_aa: unknown edge mode %d", edge_mode);
  FV_REQUIRE(edge_mode != FV_EDGE_REFLECT || L >= 6, FV_E_UNSUPPORTED,
             "fv_snake_aa: reflect padding of 5 samples needs L >= 6 (got %d)", L);
  SnakeFilt f;
  // the 12 taps are tiny, deterministic buffers of the module; fetch them once per call (async, stream ordered
  // copies would need a staging buffer - the module passes HOST copies of the taps instead, see python side)
  for (int i = 0; i < 12; ++i) {
    f.up[i] = 2.0f * filt_up[i];  // the up-sampler's gain of `ratio` (= 2) is folded into its taps
    f.dn[i] = filt_down[i];
  }
  if (edge_mode != FV_EDGE_REPLICATE) {
    const long long items = (long long)B * L * (pitch / 2);
    snake_aa_direct_kernel<<<grid1d(items, 256), 256, 0, (cudaStream_t)stream>>>(x32, (__half*)out16, alpha, beta, f,
                                                                                logscale, B, L, C, pitch, split, edge_mode);
    FV_CHECK_LAUNCH("snake_aa_direct_kernel");
    return 0;
  }
  // Segment length: every thread does the same amount of work, so the grid runs in whole waves of SN_BLOCKS x 256 threads per
  // SM and a launch of 3.1 waves costs 4.  Pick the multiple of 6 in [36, 96] with the best (work / waves) ratio,
  // counting the 6 pre-roll steps every segment pays.
  // Look-ahead of the plain (non-split) launches: a shared-memory ring with two 6-step windows in flight and two blocks per SM
  // (default; measured on B200, BigVGAN cfg C: 4.12 -> 3.96 ms of Snake time per forward against the one-window register
  // look-ahead at the same occupancy; three blocks per SM at 80 registers: 4.27 ms).  FV_SNAKE_RING=0 / 1 / 2: register
  // look-ahead / ring with three blocks / ring with two blocks.
  static const int ring_blocks = [] {
    const char* e = getenv("FV_SNAKE_RING");
    return (e && e[0] == '0') ? 0 : ((e && e[0] == '1') ? 3 : 2);
  }();
  static const int ring_windows = [] {   // FV_SNAKE_RING=3: two blocks per SM, three windows in flight
    const char* e = getenv("FV_SNAKE_RING");
    return (e && e[0] == '3') ? 3 : 2;
  }();
  const bool use_ring = ring_blocks != 0 && split == 0;
  int seg_len = 48;
  {
    const long long per_wave = (long long)num_sms() * (use_ring ? ring_blocks : SN_BLOCKS) * 256;
    double best = -1.0;
    for (int sl = SN_SEG_MIN; sl <= SN_SEG_MAX; sl += 6) {
      const long long thr = (long long)B * ceil_div(L, sl) * (pitch / 2);
      const long long waves = (thr + per_wave - 1) / per_wave;
      const double score = (double)B * L * (pitch / 2) / ((double)waves * per_wave * (sl + 6));  // useful steps per slot-step
      if (score > best) {
        best = score;
        seg_len = sl;
      }
    }
  }
  const int n_seg = ceil_div(L, seg_len);
  const long long total = (long long)B * n_seg * (pitch / 2);
  const int reverse = next_tile_direction();
  if (use_ring) {
    const int smem = (ring_windows + 1) * 6 * 256 * (int)sizeof(float2);
    cudaError_t le;
    if (ring_blocks == 3)
      le = launch_kernel(snake_aa_ring_kernel<3, 2>, dim3(grid1d(total, 256)), dim3(256), smem, (cudaStream_t)stream, 1, x32,
                         (__half*)out16, alpha, beta, f, logscale, B, L, C, pitch, n_seg, seg_len, reverse);
    else if (ring_windows == 3)
      le = launch_kernel(snake_aa_ring_kernel<2, 3>, dim3(grid1d(total, 256)), dim3(256), smem, (cudaStream_t)stream, 1, x32,
                         (__half*)out16, alpha, beta, f, logscale, B, L, C, pitch, n_seg, seg_len, reverse);
    else
      le = launch_kernel(snake_aa_ring_kernel<2, 2>, dim3(grid1d(total, 256)), dim3(256), smem, (cudaStream_t)stream, 1, x32,
                         (__half*)out16, alpha, beta, f, logscale, B, L, C, pitch, n_seg, seg_len, reverse);
    FV_REQUIRE(le == cudaSuccess, FV_E_DRIVER, "launch of snake_aa_ring_kernel failed");
    FV_CHECK_LAUNCH("snake_aa_ring_kernel");
    return 0;
  }
  if (split)  // strict precision: [hi | lo] fp16 pairs (the extra stores stay out of the default instantiation)
    FV_REQUIRE(launch_kernel(snake_aa_kernel<true>, dim3(grid1d(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, x32,
                             (__half*)out16, alpha, beta, f, logscale, B, L, C, pitch, n_seg, seg_len,
                             split, reverse) == cudaSuccess,
               FV_E_DRIVER, "launch of snake_aa_kernel failed");
  else
    FV_REQUIRE(launch_kernel(snake_aa_kernel<false>, dim3(grid1d(total, 256)), dim3(256), 0, (cudaStream_t)stream, 1, x32,
                             (__half*)out16, alpha, beta, f, logscale, B, L, C, pitch, n_seg, seg_len,
                             split, reverse) == cudaSuccess,
               FV_E_DRIVER, "launch of snake_aa_kernel failed");
  FV_CHECK_LAUNCH("snake_aa_kernel");
  return 0;
}

extern "C" int fv_dwconv_layernorm(const float* x32, void* out16, float* out32, const float* dw_w, const float* dw_b,
                                   const float* ln_w, const float* ln_b, float eps, int B, int T, int C, int pitch,
                                   int k, int split, void* stream) {
  FV_REQUIRE(x32 && (out16 || out32) && ln_w && ln_b && B > 0 && T > 0 && C > 0 && pitch >= C &&
                 (split == 0 || split == pitch),
             FV_E_BADARG, "fv_dwconv_layernorm: bad arguments");
  FV_REQUIRE(k <= 0 || (dw_w && dw_b && (k & 1)), FV_E_BADARG, "fv_dwconv_layernorm: bad depthwise kernel");
  const int vec_smem = DW8_R * pitch * (int)sizeof(float);
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(x32) | reinterpret_cast<uintptr_t>(out16) |
                           reinterpret_cast<uintptr_t>(out32) | reinterpret_cast<uintptr_t>(dw_w) |
                           reinterpret_cast<uintptr_t>(dw_b) | reinterpret_cast<uintptr_t>(ln_w) |
                           reinterpret_cast<uintptr_t>(ln_b)) & 15) == 0;
  static const bool vec_on = [] {
    const char* e = getenv("FV_DWLN_VEC");  // FV_DWLN_VEC=0: the scalar 4-row kernel (A/B measurements)
    return !(e && e[0] == '0');
  }();
  static const bool pipe_on = [] {
    const char* e = getenv("FV_DWLN_PIPE");  // FV_DWLN_PIPE=0: the per-tile CTA kernels (A/B measurements)
    return !(e && e[0] == '0');
  }();
  {
    const int win = DW8_R + (k > 0 ? k - 1 : 0);
    const int pipe_smem = 2 * win * pitch * (int)sizeof(float);
    const int threads = round_up(pitch / 2, 32);
    if (pipe_on && (k <= 0 || k == 7) && C % 2 == 0 && pitch % 4 == 0 && aligned16 && threads <= 1024 &&
        pipe_smem <= 200 * 1024) {
      const int tiles_per_b = ceil_div(T, DW8_R);
      const int total_tiles = B * tiles_per_b;
      // narrow layers leave room for several resident CTAs per SM (each with its own two stages in flight)
      int per_sm = (200 * 1024) / (pipe_smem + 4096);
      per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
      {  // ... and the register file: ~112 / 80 / 64 registers per thread in the 384 / 768 / 1024-thread instantiations
        const int regs = threads <= 384 ? 112 : (threads <= 768 ? 80 : 64);
        const int by_regs = 65536 / (threads * regs);
        per_sm = per_sm > by_regs ? (by_regs > 0 ? by_regs : 1) : per_sm;
      }
      const int slots = num_sms() * per_sm;
      const int grid = total_tiles < slots ? total_tiles : slots;
#define FV_DWLN_PIPE_LAUNCH(KK, MAXT)                                                                              \
  do {                                                                                                             \
    static std::atomic<unsigned long long> done{0};                                                                \
    int rc = check_cuda(ensure_dyn_smem(dwconv_ln_pipe_kernel<KK, MAXT>, 200 * 1024, done),                        \
                        "cudaFuncSetAttribute(dwconv_ln_pipe_kernel)");                                            \
    if (rc) return rc;                                                                                             \
    cudaError_t le = launch_kernel(dwconv_ln_pipe_kernel<KK, MAXT>, dim3(grid), dim3(threads), pipe_smem,          \
                                   (cudaStream_t)stream, 1, x32, (__half*)out16, out32, dw_w, dw_b, ln_w, ln_b,    \
                                   eps, T, C, pitch, tiles_per_b, total_tiles, split);                             \
    FV_REQUIRE(le == cudaSuccess, FV_E_DRIVER, "launch of dwconv_ln_pipe_kernel failed: %s",                       \
               cudaGetErrorString(le));                                                                            \
  } while (0)
      if (k == 7) {
        if (threads <= 384) FV_DWLN_PIPE_LAUNCH(7, 384);
        else if (threads <= 768) FV_DWLN_PIPE_LAUNCH(7, 768);
        else FV_DWLN_PIPE_LAUNCH(7, 1024);
      } else {
        if (threads <= 384) FV_DWLN_PIPE_LAUNCH(0, 384);
        else if (threads <= 768) FV_DWLN_PIPE_LAUNCH(0, 768);
        else FV_DWLN_PIPE_LAUNCH(0, 1024);
      }
#undef FV_DWLN_PIPE_LAUNCH
      FV_CHECK_LAUNCH("dwconv_ln_pipe_kernel");
      return 0;
    }
  }
  if (vec_on && (k <= 0 || k == 7) && C % 4 == 0 && pitch % 4 == 0 && aligned16 && vec_smem <= 200 * 1024) {
    static std::atomic<unsigned long long> attr_done7{0}, attr_done0{0};
    int rc = check_cuda(ensure_dyn_smem(dwconv_ln_vec_kernel<7>, 200 * 1024, attr_done7),
                        "cudaFuncSetAttribute(dwconv_ln_vec_kernel)");
    if (!rc) rc = check_cuda(ensure_dyn_smem(dwconv_ln_vec_kernel<0>, 200 * 1024, attr_done0),
                             "cudaFuncSetAttribute(dwconv_ln_vec_kernel)");
    if (rc) return rc;
    const int tiles_per_b = ceil_div(T, DW8_R);
    const int grid = B * tiles_per_b;
    // threads: every thread walks ceil(C4 / threads) channel groups; pick the count that leaves the fewest idle slots
    // (C = 1408: 352 groups -> 192 threads x 2 = 92% instead of 256 x 2 = 69%)
    const int c4 = C / 4;
    const int iters = ceil_div(c4, 256);
    int threads = round_up(ceil_div(c4, iters), 32);
    threads = threads < 64 ? 64 : (threads > 256 ? 256 : threads);
    cudaError_t le;
    if (k == 7)
      le = launch_kernel(dwconv_ln_vec_kernel<7>, dim3(grid), dim3(threads), vec_smem, (cudaStream_t)stream, 1, x32,
                         (__half*)out16, out32, dw_w, dw_b, ln_w, ln_b, eps, T, C, pitch, tiles_per_b, split);
    else
      le = launch_kernel(dwconv_ln_vec_kernel<0>, dim3(grid), dim3(threads), vec_smem, (cudaStream_t)stream, 1, x32,
                         (__half*)out16, out32, dw_w, dw_b, ln_w, ln_b, eps, T, C, pitch, tiles_per_b, split);
    FV_REQUIRE(le == cudaSuccess, FV_E_DRIVER, "launch of dwconv_ln_vec_kernel failed: %s", cudaGetErrorString(le));
    FV_CHECK_LAUNCH("dwconv_ln_vec_kernel");
    return 0;
  }
  const int tiled_smem = DW_R * pitch * (int)sizeof(float);
  if ((k <= 0 || k == 7) && tiled_smem <= 48 * 1024) {
    const int tiles_per_b = ceil_div(T, DW_R);
    const int grid = B * tiles_per_b;
    if (k == 7)
      dwconv_ln_tiled_kernel<7><<<grid, 256, tiled_smem, (cudaStream_t)stream>>>(
          x32, (__half*)out16, out32, dw_w, dw_b, ln_w, ln_b, eps, T, C, pitch, tiles_per_b, split);
    else
      dwconv_ln_tiled_kernel<0><<<grid, 256, tiled_smem, (cudaStream_t)stream>>>(
          x32, (__half*)out16, out32, dw_w, dw_b, ln_w, ln_b, eps, T, C, pitch, tiles_per_b, split);
    FV_CHECK_LAUNCH("dwconv_ln_tiled_kernel");
    return 0;
  }
  const int warps = 4;
  const int smem = warps * pitch * (int)sizeof(float);
  FV_REQUIRE(smem <= 48 * 1024, FV_E_UNSUPPORTED, "fv_dwconv_layernorm: C too large (%d)", C);
  const long long rows = (long long)B * T;
  dwconv_ln_kernel<<<(int)((rows + warps - 1) / warps), warps * 32, smem, (cudaStream_t)stream>>>(
      x32, (__half*)out16, out32, dw_w, dw_b, ln_w, ln_b, eps, B, T, C, pitch, k, split);
  FV_CHECK_LAUNCH("dwconv_ln_kernel");
  return 0;
}

extern "C" int fv_istft_ola(const float* frames, const float* window, float* wav, int B, int T, int n_fft, int hop,
                            int frame_pitch, int center, void* stream) {
  FV_REQUIRE(frames && window && wav && B > 0 && T > 0 && n_fft > 0 && hop > 0 && hop <= n_fft && frame_pitch >= n_fft,
             FV_E_BADARG, "fv_istft_ola: bad arguments");
  // "same" (vocos.spectral_ops.ISTFT): trim (win - hop) / 2 either side of the (T-1) hop + win long overlap-add -> T hop
  // "center" (torch.istft(center=True)): trim n_fft / 2 either side -> (T - 1) hop samples
  FV_REQUIRE(center || (n_fft - hop) % 2 == 0, FV_E_BADARG, "fv_istft_ola: padding='same' needs an even (win - hop)");
  FV_REQUIRE(!center || T >= 2, FV_E_BADARG, "fv_istft_ola: padding='center' needs at least two frames");
  const int trim = center ? n_fft / 2 : (n_fft - hop) / 2;
  const int Lw = center ? (T - 1) * hop : T * hop;
  const long long total = (long long)B * Lw;
  istft_ola_kernel<<<grid1d(total, 256), 256, 0, (cudaStream_t)stream>>>(frames, window, wav, B, T, n_fft, hop,
                                                                        frame_pitch, trim, Lw);
  FV_CHECK_LAUNCH("istft_ola_kernel");
  return 0;
}

extern "C" int fv_noise_conv(const float* tpl, const float* w, const float* bias, float* out32, int B, int L_audio,
                             int L_out, int C, int pitch, int k, int stride, int pad, void* stream) {
  FV_REQUIRE(tpl && w && bias && out32 && B > 0 && L_audio > 0 && L_out > 0 && C > 0 && pitch >= C && k > 0 &&
                 stride > 0,
             FV_E_BADARG, "fv_noise_conv: bad arguments");
  const long long total = (long long)B * L_out * pitch;
  noise_conv_kernel<<<grid1d(total, 256), 256, 0, (cudaStream_t)stream>>>(tpl, w, bias, out32, B, L_audio, L_out, C,
                                                                         pitch, k, stride, pad);
  FV_CHECK_LAUNCH("noise_conv_kernel");
  return 0;
}

extern "C" int fv_act_cast(const float* x32, const float* noise, const float* noise_w, void* out16, float* out32,
                           int act, float act_param, int act16, float act16_param, float out_scale, int accumulate,
                           int B, int L, int C, int in_pitch, int out16_pitch, int out16_coff, int out32_pitch,
                           int split, void* stream) {
  FV_REQUIRE(x32 && (out16 || out32) && B > 0 && L > 0 && C > 0 && in_pitch >= C &&
                 (noise == nullptr || noise_w != nullptr),
             FV_E_BADARG, "fv_act_cast: bad arguments");
  FV_REQUIRE(!out16 || out16_pitch >= out16_coff + C + split, FV_E_BADARG,
             "fv_act_cast: out16 slice exceeds its pitch");
  FV_REQUIRE(!out32 || out32_pitch >= C, FV_E_BADARG, "fv_act_cast: out32 pitch too small");
  const long long rows = (long long)B * L;
  act_cast_kernel<<<grid1d(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(
      x32, noise, noise_w, (__half*)out16, out32, act, act_param, act16, act16_param, out_scale, accumulate, rows, C,
      in_pitch, out16_pitch, out16_coff, out32_pitch, split);
  FV_CHECK_LAUNCH("act_cast_kernel");
  return 0;
}

extern "C" int fv_resample_linear(const float* x32, float* out32, void* out16, int pre_act, float pre_param, int act,
                                  float act_param, int B, int L_in, int L_out, int C, int in_pitch, int out_pitch,
                                  int out_coff, float scale, int split, void* stream) {
  FV_REQUIRE(x32 && (out16 || out32) && B > 0 && L_in > 0 && L_out > 0 && C > 0 && in_pitch >= C &&
                 out_pitch >= out_coff + C + split && scale > 0.f && !(split > 0 && out32),
             FV_E_BADARG, "fv_resample_linear: bad arguments");
  const long long total = (long long)B * L_out * C;
  resample_linear_kernel<<<grid1d(total, 256), 256, 0, (cudaStream_t)stream>>>(
      x32, out32, (__half*)out16, pre_act, pre_param, act, act_param, B, L_in, L_out, C, in_pitch, out_pitch,
      out_coff, scale, split);
  FV_CHECK_LAUNCH("resample_linear_kernel");
  return 0;
}
