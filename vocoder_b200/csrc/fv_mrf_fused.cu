// fv_mrf_fused.cu - one MRF stage (mean over kernel sizes of ResBlock1 chains) of the HiFiGAN generator as ONE
// kernel whose intermediates never leave the SM (entry: fv_mrf_fused, contract in include/fv_vocoder.h).
//
// Reference: ParralelBlock.forward / ResBlock1.forward, fish_vocoder/modules/generators/hifigan.py:101-108,117-133
//     for each kernel size k:  x_k = x;  3 x { xt = silu(x_k); xt = conv(k, d_i)(xt); xt = silu(xt);
//                                             xt = conv(k, 1)(xt); x_k = xt + x_k }
//     out = mean_k x_k
//
// Per CTA (one per SM, persistent over 512-row time tiles; rows = time steps, columns = channels):
//     TMEM   X[4][128 x C] fp32   residual stream of the tile (the accumulator the second conv of a pair adds onto)
//            T[4][128 x C] fp32   accumulator of the first conv of a pair
//     smem   XA[576][C] fp16      act(x)   operand of conv1  (K-major, swizzled; 32 guard rows either side)
//            TA[576][C] fp16      act(t)   operand of conv2
//            ring of [C_out x C_in] fp16 weight tap tiles (TMA, all CTAs stream the same tiles out of L2)
//     a tap / dilation shift is a UMMA descriptor that starts `off` rows into the slab (hardware-verified by
//     fv_debug_rowshift_probe); rows outside the sequence are written as zeros (= the reference's zero padding);
//     the tile's halo (sum of all tap reaches, 60 rows for k=11, d=(1,3,5)) is recomputed by neighbouring tiles.
// Warp roles: warp 0 = weight-tile TMA producer, warp 1 = tcgen05.mma issuer + TMEM owner, warps 4-19 = epilogue
// (one thread per tile row: TMEM -> bias/activation -> fp16 operand row in smem; tile entry/exit through TMA boxes).
// HBM traffic per element of the stage: 4 B in (x) + 2..6 B out, against 16 B per conv pair for the layer-wise path.
#include <cstdlib>
#include <mutex>

#include "fv_common.cuh"

namespace fv {

constexpr int kFrGuard = 32;     // guard rows either side of an operand slab = largest tap reach supported
constexpr int kFrSmemSm = 233472;  // shared memory of one SM; every resident CTA also pays 1 KB of driver reservation
constexpr int kFrMaxConvs = FV_MRF_MAX_BLOCKS * FV_MRF_MAX_PAIRS * 2;

struct MrfParams {
  CUtensorMap tmX;    // x     fp32 {pitch, L, B}, box {32, 32, 1}, SWIZZLE_128B
  CUtensorMap tmW;    // w     fp16 {C, w_rows},   box {C, C},      swizzle = row bytes
  CUtensorMap tmO32;  // out32 fp32 {pitch, L, B}, box {32, 32, 1}, SWIZZLE_128B (store + running-sum reload)
  CUtensorMap tmO16;  // out16 fp16 {pitch, L, B}, box {32, 32, 1}, SWIZZLE_64B
  int B, L, tiles_per_b, total_tiles;
  int V, h0;  // valid rows per tile (multiple of 32) and their first local row (multiple of 32, >= halo)
  int n_blocks, n_pairs;
  int ksize[FV_MRF_MAX_BLOCKS];
  int dil[FV_MRF_MAX_BLOCKS][FV_MRF_MAX_PAIRS][2];
  int w_row0[FV_MRF_MAX_BLOCKS][FV_MRF_MAX_PAIRS][2];
  const float* bias;
  int act, out_act, has_o16;
  int accumulate;  // the first block's exit adds onto what out32 already holds (pair-wise launches building the MRF mean)
  int use_ws;  // weight-stationary MMAs (tcgen05.mma.ws): the tap tile's K chunk is latched once per four 128-row blocks
  float act_param, out_act_param, out_scale;
};

// EW epilogue warps (a multiple of 4, at most 16: warp w owns TMEM lane quarter w % 4), NCTA co-resident CTAs per SM (with
// C = 32 two CTAs fit - 256 TMEM columns and < 113 KB each - so the tensor pipe works on one CTA's conv while the other
// CTA's epilogue warps turn accumulators into the next operand), MB 128-row blocks per tile (4 = 512 rows; 2 for C = 128,
// whose residual stream + conv1 accumulator fill TMEM with 256 rows), MAXCONV convs whose biases are staged.
// C = 128: an operand row (256 B) spans two 128-byte swizzle atoms, so every operand slab is KH = 2 sub-slabs of
// [rows x 64 channels] and every tap tile streams through the ring as KH K-halves [C_out x 64].
template <int C, int EW, int NCTA, int MB = 4, int MAXCONV = kFrMaxConvs>
struct FrCfg {
  static constexpr int THREADS = 64 + EW * 32;                // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
  static constexpr int SMEM_LIMIT = kFrSmemSm / NCTA - 1024;
  static constexpr int KH = C > 64 ? C / 64 : 1;              // K halves per operand row / tap tile
  static constexpr int KC = C / KH;                           // channels per K half
  static constexpr int ROWB = KC * 2;                         // bytes per sub-slab row = swizzle span
  static constexpr int TILE = MB * 128;                       // rows per tile
  static constexpr int SLAB_ROWS = TILE + 2 * kFrGuard;
  static constexpr int SUB = SLAB_ROWS * ROWB;                // one sub-slab (multiple of 1024)
  static constexpr int SLAB = KH * SUB;
  static constexpr int W_BYTES = C * ROWB;                    // one [C_out x KC] K-half of a tap tile
  static constexpr int W_TILE = W_BYTES < 1024 ? 1024 : W_BYTES;  // ring slot (swizzled tiles stay 1024-byte aligned)
  static constexpr int CW = C < 32 ? C : 32;                  // columns per epilogue patch (one TMEM load / TMA box)
  static constexpr int NCH = C / CW;                          // patches per row
  static constexpr int P32 = 32 * CW * 4;                     // per-warp 32 x CW fp32 patch, swizzle span = CW * 4 bytes
  static constexpr int P16 = 32 * CW * 2;                     // per-warp 32 x CW fp16 patch, swizzle span = CW * 2 bytes
  static constexpr int STG32 = EW * P32;
  static constexpr int STG16 = EW * P16;
  // when the staging patches do not fit next to the slabs they alias them (slabs are dead at tile entry / exit)
  static constexpr bool ALIAS = 2 * SLAB + STG32 + STG16 + 8 * W_TILE + 16384 > SMEM_LIMIT;
  static constexpr int XA_OFF = 0;
  static constexpr int TA_OFF = SLAB;
  static constexpr int STG32_OFF = ALIAS ? TA_OFF : 2 * SLAB;
  static constexpr int STG16_OFF = ALIAS ? XA_OFF : 2 * SLAB + STG32;
  static constexpr int RING_OFF = ALIAS ? 2 * SLAB : 2 * SLAB + STG32 + STG16;
  static constexpr int BIAS_BYTES = MAXCONV * C * 4;
  static constexpr int TAIL = 1024 + BIAS_BYTES;
  static constexpr int NS_RAW = (SMEM_LIMIT - RING_OFF - TAIL) / W_TILE;
  static constexpr int NS = NS_RAW > 16 ? 16 : NS_RAW;
  static constexpr int SMEM = RING_OFF + NS * W_TILE + TAIL;
  static constexpr int TMEM_COLS = 2 * MB * C;                // 512 (C = 64, 128) / 256 (C = 32) / 128 (C = 16)
  static constexpr int GROUPS = EW / 4;                       // epilogue warps per TMEM lane quarter
  static constexpr int ITEMS = MB * NCH;                      // (128-row block, column patch) items per lane quarter
  static constexpr int IPW = ITEMS / GROUPS;                  // items each epilogue warp walks
  static_assert(EW % 4 == 0 && EW >= 4 && EW <= 16 && ITEMS % GROUPS == 0, "bad epilogue warp count");
  static_assert(TMEM_COLS * NCTA <= 512, "co-resident CTAs exceed TMEM");
  static_assert(!ALIAS || (STG32 <= SLAB && STG16 <= SLAB), "staging does not fit in the slabs it aliases");
  static_assert(NS >= 3, "weight ring too shallow");
  static_assert(SUB % 1024 == 0 && W_TILE % 1024 == 0 && (STG32 % 1024 == 0) && (STG16 % 1024 == 0),
                "swizzled tiles must stay 1024-byte aligned");
  static_assert(C == 16 || C == 32 || C == 64 || C == 128, "supported channel counts");
};

// byte offset of 16-byte chunk `chunk` of row `row` inside a swizzled tile whose rows are ROWB bytes = the swizzle span
// (base 1024-aligned): K-major operand slabs and the TMA staging patches use the same function
template <int ROWB>
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {
  if constexpr (ROWB == 128) return row * 128 + ((chunk ^ (row & 7)) << 4);
  else if constexpr (ROWB == 64) return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
  else return row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4);
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// CW = 32 or 16 accumulator columns of this thread's TMEM lane <-> registers
template <int CW>
__device__ __forceinline__ void tmem_ld_cw(uint32_t taddr, uint32_t (&r)[32]) {
  if constexpr (CW == 32) tmem_ld_32x32b_x32(taddr, r);
  else tmem_ld_32x32b_x16(taddr, r);
}
template <int CW>
__device__ __forceinline__ void tmem_st_cw(uint32_t taddr, const uint32_t (&r)[32]) {
  if constexpr (CW == 32) tmem_st_32x32b_x32(taddr, r);
  else tmem_st_32x32b_x16(taddr, r);
}

// Inner activation of the residual blocks, resolved at compile time: the epilogue body must be straight-line code
// (a per-element switch costs a uniform branch + BSSY/BSYNC per element and serialises the SFU latencies).
//   kActSilu     x / (1 + 2^(-x log2 e)): ex2.approx.ftz + rcp.approx.ftz (2 SFU ops, ~1e-7 relative)
//   kActSiluTanh x/2 + x/2 * tanh(x/2):   tanh.approx (1 SFU op, |error| <= 2.4e-4 |x|) - opt-in, see fv_set_mrf_tuning
//   kActLeaky    x > 0 ? x : slope * x
//   kActSiluH2   the same formula evaluated on packed fp16 pairs: cvt.rn.f16x2 of (x0/2, x1/2), ONE tanh.approx.f16x2 and
//                one fma.rn.f16x2 produce the operand word of two channels (5 instructions / 1 SFU op per pair instead of
//                9 / 2); the result carries ~3 fp16 roundings instead of 1 - opt-in (FV_ACT_SILU_H2)
constexpr int kActSilu = 0, kActSiluTanh = 1, kActLeaky = 2, kActSiluH2 = 3;
template <int ACT>
__device__ __forceinline__ float fr_act(float v, float param) {
  if constexpr (ACT == kActSilu) {
    return silu_fast(v);
  } else if constexpr (ACT == kActSiluTanh || ACT == kActSiluH2) {
    return silu_tanh(v);
  } else {
    return v > 0.f ? v : v * param;
  }
}
// packed operand word of two channels: fp16x2(act(a), act(b)), a in the low half
template <int ACT>
__device__ __forceinline__ uint32_t fr_act_pack(float a, float b, float param) {
  if constexpr (ACT == kActSiluH2) {
    const uint32_t h = pack_half2_sat(0.5f * a, 0.5f * b);
    uint32_t t, r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(h), "r"(t));
    return r;
  } else {
    return pack_half2_sat(fr_act<ACT>(a, param), fr_act<ACT>(b, param));
  }
}

template <int C, int ACT, int EW, int NCTA, int MB, int MAXCONV>
__global__ void __launch_bounds__(FrCfg<C, EW, NCTA, MB, MAXCONV>::THREADS, NCTA)
    mrf_fused_kernel(const __grid_constant__ MrfParams p) {
  using Cfg = FrCfg<C, EW, NCTA, MB, MAXCONV>;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int KH = Cfg::KH, KC = Cfg::KC;
  constexpr int CW = Cfg::CW;      // columns per epilogue patch
  constexpr int NCH = Cfg::NCH;    // patches per row
  constexpr int R32 = CW * 4, R16 = CW * 2;  // row bytes of the fp32 / fp16 staging patches
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("fv: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* tail = smem + Cfg::RING_OFF + Cfg::NS * Cfg::W_TILE;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* w_empty = w_full + Cfg::NS;
  uint64_t* a_ready = w_empty + Cfg::NS;   // [2] epilogue warps -> MMA: operand of the next conv is in smem / X is in TMEM
  uint64_t* acc_full = a_ready + 2;        // [2] MMA -> epilogue: accumulators of the current conv are complete
  uint64_t* stg_bar = acc_full + 2;        // one per epilogue warp: TMA loads into its staging patch
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_bar + EW);
  float* s_bias = reinterpret_cast<float*>(tail + 1024);  // [block][pair][2][C]: b1, cumulative b2

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmO32);
    if (p.has_o16) tma_prefetch_desc(&p.tmO16);
    for (int i = 0; i < Cfg::NS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(&a_ready[0], EW);
    mbar_init(&a_ready[1], EW);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    for (int i = 0; i < EW; ++i) mbar_init(&stg_bar[i], 1);
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  pdl_wait();  // global memory written by the previous kernel is read from here on
  // biases -> smem (conv2 biases as running sums over the pairs: the TMEM residual stream never sees them)
  for (int i = threadIdx.x; i < p.n_blocks * C; i += blockDim.x) {
    const int j = i / C, c = i % C;
    float cum = 0.f;
    for (int pi = 0; pi < p.n_pairs; ++pi) {
      const float* src = p.bias + ((size_t)(j * p.n_pairs + pi) * 2) * C;
      s_bias[((j * p.n_pairs + pi) * 2 + 0) * C + c] = src[c];
      cum += src[C + c];
      s_bias[((j * p.n_pairs + pi) * 2 + 1) * C + c] = cum;
    }
  }
  // guard rows of every sub-slab of both slabs: never written afterwards, only ever feed rows outside a tile's valid window
  for (int i = threadIdx.x; i < 2 * KH * 2 * kFrGuard * ROWB / 16; i += blockDim.x) {
    const int per = kFrGuard * ROWB / 16;
    const int which = i / per, o = i % per;   // which = ((slab * KH + sub) * 2 + end)
    const int end = which & 1, sub = (which >> 1) % KH, slab = (which >> 1) / KH;
    uint8_t* base = smem + (slab ? Cfg::TA_OFF : Cfg::XA_OFF) + sub * Cfg::SUB + (end ? (kFrGuard + Cfg::TILE) * ROWB : 0);
    reinterpret_cast<uint4*>(base)[o] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t X_COL = 0, T_COL = MB * C;

  if (warp == 0) {
    // ---------------------------------------------------------------- weight-tile producer
    const bool leader = elect_one();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int j = 0; j < p.n_blocks; ++j)
        for (int pi = 0; pi < p.n_pairs; ++pi)
          for (int cv = 0; cv < 2; ++cv) {
            const int row0 = p.w_row0[j][pi][cv];
            for (int tap = 0; tap < p.ksize[j]; ++tap)
              for (int kh = 0; kh < KH; ++kh, ++it) {   // ring item = one K-half [C_out x KC] of a tap tile
                const int s = it % Cfg::NS;
                mbar_wait(&w_empty[s], ((it / Cfg::NS) & 1) ^ 1);
                if (leader) {
                  mbar_arrive_expect_tx(&w_full[s], Cfg::W_BYTES);
                  tma_load_2d(smem + Cfg::RING_OFF + s * Cfg::W_TILE, &p.tmW, &w_full[s], kh * KC, row0 + tap * C);
                }
              }
          }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_f16(128, C);
    uint32_t it = 0, n = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int j = 0; j < p.n_blocks; ++j) {
        const int k = p.ksize[j], half_k = (k - 1) / 2;
        for (int pi = 0; pi < p.n_pairs; ++pi)
          for (int cv = 0; cv < 2; ++cv, ++n) {
            const uint32_t slab = smem_u32(smem + (cv == 0 ? Cfg::XA_OFF : Cfg::TA_OFF)) + kFrGuard * ROWB;
            const uint32_t d_col = tmem_base + (cv == 0 ? T_COL : X_COL);
            const int dil = p.dil[j][pi][cv];
            mbar_wait(&a_ready[0], n & 1);
            tc_fence_after();
            for (int tap = 0; tap < k; ++tap)
              for (int kh = 0; kh < KH; ++kh, ++it) {
                const int s = it % Cfg::NS;
                mbar_wait(&w_full[s], (it / Cfg::NS) & 1);
                tc_fence_after();
                if (leader) {
                  // a tap / dilation shift = a descriptor that starts `off` rows into the (sub-)slab
                  const uint64_t da0 = make_kmajor_desc(slab + kh * Cfg::SUB + ((tap - half_k) * dil) * ROWB, ROWB);
                  const uint64_t db0 = make_kmajor_desc(smem_u32(smem + Cfg::RING_OFF + s * Cfg::W_TILE), ROWB);
                  const uint32_t acc0 = (cv == 1 || tap > 0 || kh > 0) ? 1u : 0u;
                  bool done = false;
                  if constexpr (C == 64 && MB == 4) {
                    if (p.use_ws) {
                      // K step outermost: B chunk kk is read from shared memory ONCE (collector fill) for the four blocks
                      // instead of four times: 4 x 4 KB (A) + 2 KB (B) per four UMMAs instead of 4 x 6 KB
#define FV_WS_STEP(KK)                                                                                                   \
  {                                                                                                                      \
    const uint64_t db = desc_advance(db0, KK * 32);                                                                      \
    const uint32_t acc = (acc0 || KK > 0) ? 1u : 0u;                                                                     \
    umma_f16_ws<KK, 0>(d_col + 0 * C, desc_advance(da0, 0 * 128 * ROWB + KK * 32), db, idesc, acc);                     \
    umma_f16_ws<KK, 1>(d_col + 1 * C, desc_advance(da0, 1 * 128 * ROWB + KK * 32), db, idesc, acc);                     \
    umma_f16_ws<KK, 1>(d_col + 2 * C, desc_advance(da0, 2 * 128 * ROWB + KK * 32), db, idesc, acc);                     \
    umma_f16_ws<KK, 2>(d_col + 3 * C, desc_advance(da0, 3 * 128 * ROWB + KK * 32), db, idesc, acc);                     \
  }
                      FV_WS_STEP(0) FV_WS_STEP(1) FV_WS_STEP(2) FV_WS_STEP(3)
#undef FV_WS_STEP
                      done = true;
                    }
                  }
                  if (!done) {
#pragma unroll
                    for (int m = 0; m < MB; ++m) {
#pragma unroll
                      for (int kk = 0; kk < KC / 16; ++kk)
                        umma_f16_ss(d_col + m * C, desc_advance(da0, m * 128 * ROWB + kk * 32), desc_advance(db0, kk * 32),
                                    idesc, (acc0 || kk > 0) ? 1u : 0u);
                    }
                  }
                  umma_commit(&w_empty[s]);
                }
                __syncwarp();
              }
            if (leader) umma_commit(&acc_full[0]);
            __syncwarp();
          }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = tile row
    // warp w may only touch TMEM lanes 32 (w % 4) .. +31: it owns that lane quarter; the (128-row block m, column patch cc)
    // items of a quarter are dealt to its GROUPS warps: item = g + i * GROUPS, m = item / NCH, cc = item % NCH
    // (g = the warp's index among the warps of the quarter); local row = 128 m + 32 q + lane
    const int e = warp - 2;
    const int q = warp & 3;
    const int g = e >> 2;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* patch32 = smem + Cfg::STG32_OFF + e * Cfg::P32;
    uint8_t* patch16 = smem + Cfg::STG16_OFF + e * Cfg::P16;
    const uint32_t p32_base = smem_u32(patch32), p16_base = smem_u32(patch16);  // this thread's row = lane
    const uint32_t xa_base = smem_u32(smem + Cfg::XA_OFF), ta_base = smem_u32(smem + Cfg::TA_OFF);
    uint64_t* my_bar = &stg_bar[e];
    uint32_t stg_phase = 0, n = 0;

    // activation -> fp16 -> this thread's CW columns (CW / 8 16-byte chunks) of operand row `srow`; rows outside the
    // sequence (live == false) are written as zeros = the reference's zero padding
    auto put_operand = [&](uint32_t slab_base, int srow, int cc, const float (&v)[32], bool live) {
#pragma unroll
      for (int qq = 0; qq < CW / 8; ++qq) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t pk = fr_act_pack<ACT>(v[8 * qq + 2 * u], v[8 * qq + 2 * u + 1], p.act_param);
          w[u] = live ? pk : 0u;
        }
        // channel cc * CW + 8 qq lives in sub-slab (K half) ch / KC, 16-byte chunk (ch % KC) / 8 of its row
        const int ch = cc * CW + 8 * qq;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_base + (ch / KC) * Cfg::SUB +
                                                                     swz_off<ROWB>(srow, (ch % KC) / 8)),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
    };
    // v = accumulator + bias (CW consecutive channels; the bias row is 16-byte aligned: broadcast LDS.128)
    auto add_bias = [&](float (&v)[32], const uint32_t (&r)[32], const float* bias) {
      const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
      for (int qq = 0; qq < CW / 4; ++qq) {
        const float4 bb = b4[qq];
        v[4 * qq] = __uint_as_float(r[4 * qq]) + bb.x;
        v[4 * qq + 1] = __uint_as_float(r[4 * qq + 1]) + bb.y;
        v[4 * qq + 2] = __uint_as_float(r[4 * qq + 2]) + bb.z;
        v[4 * qq + 3] = __uint_as_float(r[4 * qq + 3]) + bb.w;
      }
    };
    // one 32-row x 32-column accumulator patch (128-row block m, column chunk cc): + bias -> activation -> operand slab
    auto acc_item32 = [&](uint32_t col, int m, int cc, const float* bias, uint32_t slab_base, int g0) {
      const int row_l = m * 128 + q * 32 + lane;
      const int gr = g0 + row_l;
      const bool in_seq = gr >= 0 && gr < p.L;
      uint32_t r[32];
      tmem_ld_cw<CW>(t_lane + col + m * C + cc * CW, r);
      tmem_ld_wait();
      float v[32];
      add_bias(v, r, bias + cc * CW);
      put_operand(slab_base, kFrGuard + row_l, cc, v, in_seq);
    };
    // tile entry of one patch: x -> X (TMEM, fp32) and act(x) -> XA (fp16); rows outside the sequence arrive as zeros
    auto entry_item = [&](int m, int cc, int b, int g0) {
      const int blk0 = m * 128 + q * 32;  // first tile row of this warp's 32-row group
      __syncwarp();
      if (lane == 0) {
        tma_store_wait_read();
        mbar_arrive_expect_tx(my_bar, Cfg::P32);
        tma_load_3d(patch32, &p.tmX, my_bar, cc * CW, g0 + blk0, b);
      }
      __syncwarp();
      mbar_wait(my_bar, stg_phase);
      stg_phase ^= 1;
      uint32_t r[32];
#pragma unroll
      for (int qq = 0; qq < CW / 4; ++qq)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[4 * qq]), "=r"(r[4 * qq + 1]), "=r"(r[4 * qq + 2]), "=r"(r[4 * qq + 3])
                     : "r"(p32_base + swz_off<R32>(lane, qq)));
      tmem_st_cw<CW>(t_lane + X_COL + m * C + cc * CW, r);
      float v[32];
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
      put_operand(xa_base, kFrGuard + blk0 + lane, cc, v, true);   // act(0) == 0: out-of-sequence rows stay zero
    };
    // tile exit of one patch: block output -> running mean in out32 (-> activated fp16 after the last block)
    auto exit_item = [&](int m, int cc, int b, int g0, int j, const float* b2c) {
      const bool last = j == p.n_blocks - 1;
      const bool add_old = j > 0 || p.accumulate;   // out32 already holds a partial mean (earlier block / earlier launch)
      const int blk0 = m * 128 + q * 32;
      const int grow = g0 + blk0;
      if (!(blk0 >= p.h0 && blk0 < p.h0 + p.V && grow < p.L)) return;  // halo rows / past the sequence
      __syncwarp();
      if (lane == 0) {
        if (add_old) {
          tma_store_wait_all();  // the partial sums this warp stored for block j-1 are visible
          mbar_arrive_expect_tx(my_bar, Cfg::P32);
          tma_load_3d(patch32, &p.tmO32, my_bar, cc * CW, grow, b);
        } else {
          tma_store_wait_read();
        }
      }
      __syncwarp();
      uint32_t r[32];
      tmem_ld_cw<CW>(t_lane + X_COL + m * C + cc * CW, r);
      tmem_ld_wait();
      float v[32];
      add_bias(v, r, b2c + cc * CW);
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] *= p.out_scale;
      if (add_old) {
        mbar_wait(my_bar, stg_phase);
        stg_phase ^= 1;
#pragma unroll
        for (int qq = 0; qq < CW / 4; ++qq) {
          float4 a;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                       : "r"(p32_base + swz_off<R32>(lane, qq)));
          v[4 * qq] += a.x;
          v[4 * qq + 1] += a.y;
          v[4 * qq + 2] += a.z;
          v[4 * qq + 3] += a.w;
        }
      }
#pragma unroll
      for (int qq = 0; qq < CW / 4; ++qq)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(p32_base + swz_off<R32>(lane, qq)),
                     "f"(v[4 * qq]), "f"(v[4 * qq + 1]), "f"(v[4 * qq + 2]), "f"(v[4 * qq + 3])
                     : "memory");
      if (last && p.has_o16) {
        if (p.out_act == FV_ACT_SILU) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fr_act<kActSilu>(v[i], 0.f);
        } else if (p.out_act == FV_ACT_SILU_TANH || p.out_act == FV_ACT_SILU_H2) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fr_act<kActSiluTanh>(v[i], 0.f);
        } else if (p.out_act == FV_ACT_LEAKY) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fr_act<kActLeaky>(v[i], p.out_act_param);
        } else if (p.out_act == FV_ACT_TANH) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = tanhf(v[i]);
        } else if (p.out_act == FV_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = gelu_erf_fast(v[i]);
        }
#pragma unroll
        for (int qq = 0; qq < CW / 8; ++qq) {
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) w[u] = pack_half2_sat(v[8 * qq + 2 * u], v[8 * qq + 2 * u + 1]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p16_base + swz_off<R16>(lane, qq)), "r"(w[0]),
                       "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&p.tmO32, patch32, cc * CW, grow, b);
        if (last && p.has_o16) tma_store_3d(&p.tmO16, patch16, cc * CW, grow, b);
        tma_store_commit();
      }
    };
    auto publish = [&](uint64_t* bar) {  // operand rows / TMEM stores of this warp are done -> MMA warp
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_b;
      const int g0 = (tile % p.tiles_per_b) * p.V - p.h0;  // global row of tile row 0
      for (int j = 0; j < p.n_blocks; ++j) {
#pragma unroll 1
        for (int i = 0; i < Cfg::IPW; ++i) {
          const int item = g + i * Cfg::GROUPS;
          entry_item(item / NCH, item % NCH, b, g0);
        }
        // the staging patches alias operand-slab rows that OTHER warps write in the conv epilogues that follow: make "every
        // warp is done with its entry patches" an explicit barrier, not only a consequence of the mbarrier chain
        if constexpr (Cfg::ALIAS) named_bar_sync(1, EW * 32);
        tmem_st_wait();
        publish(&a_ready[0]);
        for (int pi = 0; pi < p.n_pairs; ++pi) {
          const float* b1 = s_bias + ((j * p.n_pairs + pi) * 2) * C;
          const float* b2c = b1 + C;
#pragma unroll 1
          for (int cv = 0; cv < 2; ++cv, ++n) {
            // cv = 0: conv1 done, T + b1 -> act -> TA;  cv = 1: conv2 done, X (+ cumulative b2) -> act -> XA, or exit
            const uint32_t col = cv == 0 ? T_COL : X_COL;
            const float* bias = cv == 0 ? b1 : b2c;
            const uint32_t dst = cv == 0 ? ta_base : xa_base;
            const bool is_exit = cv == 1 && pi + 1 == p.n_pairs;
            mbar_wait(&acc_full[0], n & 1);
            tc_fence_after();
            if (!is_exit) {
#pragma unroll 1
              for (int i = 0; i < Cfg::IPW; ++i) {
                const int item = g + i * Cfg::GROUPS;
                acc_item32(col, item / NCH, item % NCH, bias, dst, g0);
              }
              publish(&a_ready[0]);
            } else {
              // same on the way out: no warp touches its (aliased) exit patches before every warp has finished writing
              // operand rows of the previous conv epilogue
              if constexpr (Cfg::ALIAS) named_bar_sync(1, EW * 32);
#pragma unroll 1
              for (int i = 0; i < Cfg::IPW; ++i) {
                const int item = g + i * Cfg::GROUPS;
                exit_item(item / NCH, item % NCH, b, g0, j, b2c);
              }
            }
            if (is_exit && Cfg::ALIAS && j == p.n_blocks - 1 && p.has_o16) {
              // the fp16 exit patches alias XA: nobody may start the next tile's entry before every store has read them
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
              named_bar_sync(1, EW * 32);
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int encode_rows_map(EncodeTiledFn enc, CUtensorMap* tm, const void* base, bool fp16, int C, int pitch, int L,
                           int B, int cw) {
  const cuuint64_t es = fp16 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * (cuuint64_t)L};
  cuuint32_t box[3] = {(cuuint32_t)cw, 32, 1};  // one epilogue patch: 32 rows x cw columns, swizzle span = row bytes
  const int row_bytes = cw * (int)es;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(mrf rows) failed: %d", (int)r);
  return 0;
}

// FV_MRF_C32_CTAS=1 selects the one-CTA-per-SM configuration for C = 32 (A/B measurements); default 2
static const int g_mrf_c32_ctas = [] {
  const char* e = getenv("FV_MRF_C32_CTAS");
  return (e && e[0] == '1') ? 1 : 2;
}();

// Weight-stationary MMAs for the C = 64 whole-stage configuration (see MrfParams::use_ws): measured on B200 (HiFiGAN cfg B,
// C = 64, L = 12032, B = 64) 1.640 -> 1.556 ms with bit-identical results.  FV_MRF_WS=0 disables (A/B measurements).
static const bool g_mrf_ws = [] {
  const char* e = getenv("FV_MRF_WS");
  return !(e && e[0] == '0');
}();

// FV_MRF_C64_PAIR=0: single-pair C = 64 launches use the whole-stage configuration (512-row tiles, one CTA per SM) instead
// of the pair configuration (256-row tiles, two co-resident CTAs); A/B measurement switch
static const bool g_mrf_c64_pair = [] {
  const char* e = getenv("FV_MRF_C64_PAIR");
  return !(e && e[0] == '0');
}();

// pair configuration: 256-row tiles (MB = 2), at most one (conv, conv) pair per launch.  Always for C = 128 (X + T fill
// TMEM with 256 rows); for C = 64 when the launch IS a single pair: 256 TMEM columns and < 113 KB per CTA, so two CTAs
// share an SM and the MMAs of one overlap the epilogue / entry / exit of the other.
static bool pair_config(const fv_mrf_desc* d) {
  return d->C == 128 || (d->C == 64 && d->n_blocks == 1 && d->n_pairs == 1 && g_mrf_c64_pair);
}
// (Measured and dropped, round 2: 256-row whole-stage tiles with two co-resident CTAs for short launches - C = 64, L = 12032,
// B = 1: 94 tiles instead of 32 - hifigan_b1 0.546 ms against 0.537 ms without while the forward was still host-bound, 0.429 against
// 0.426 ms after that was fixed: the fused stages' launches are shorter (0.47 against 0.50 ms summed) but not on the critical path.)
static int tile_rows_for(const fv_mrf_desc* d) { return pair_config(d) ? 256 : 512; }

template <int C, int ACT, int EW, int NCTA, int MB, int MAXCONV>
static int launch_mrf(const fv_mrf_desc* d, MrfParams& p, cudaStream_t stream) {
  using Cfg = FrCfg<C, EW, NCTA, MB, MAXCONV>;
  FV_REQUIRE(d->n_blocks * d->n_pairs * 2 <= MAXCONV, FV_E_UNSUPPORTED,
             "fv_mrf_fused: C = %d supports at most %d convs per launch (got %d blocks x %d pairs)", C, MAXCONV, d->n_blocks,
             d->n_pairs);
  EncodeTiledFn enc = get_encode_fn();
  FV_REQUIRE(enc != nullptr, FV_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  int rc = encode_rows_map(enc, &p.tmX, d->x, false, C, d->x_pitch, d->L, d->B, Cfg::CW);
  if (!rc) rc = encode_rows_map(enc, &p.tmO32, d->out32, false, C, d->out32_pitch, d->L, d->B, Cfg::CW);
  if (!rc && d->out16) rc = encode_rows_map(enc, &p.tmO16, d->out16, true, C, d->out16_pitch, d->L, d->B, Cfg::CW);
  if (rc) return rc;
  {
    // tap tiles [C_out x C_in] row-major; one ring item = a [C_out x KC] K-half (the whole tile for C <= 64)
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)d->w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)Cfg::KC, (cuuint32_t)C};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     Cfg::ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (Cfg::ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(mrf W) failed: %d", (int)r);
  }
  static std::atomic<unsigned long long> attr_done{0};
  rc = check_cuda(ensure_dyn_smem(mrf_fused_kernel<C, ACT, EW, NCTA, MB, MAXCONV>, Cfg::SMEM, attr_done),
                  "cudaFuncSetAttribute(mrf_fused_kernel)");
  if (rc) return rc;
  const int slots = num_sms() * NCTA;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  rc = check_cuda(launch_kernel(mrf_fused_kernel<C, ACT, EW, NCTA, MB, MAXCONV>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM,
                                stream, 1, p),
                  "cudaLaunchKernelEx(mrf_fused_kernel)");
  if (rc) return rc;
  FV_CHECK_LAUNCH("mrf_fused_kernel");
  return 0;
}

}  // namespace fv

using namespace fv;

extern "C" int fv_mrf_fused(const fv_mrf_desc* d, void* stream) {
  FV_REQUIRE(d != nullptr, FV_E_BADARG, "fv_mrf_fused: null descriptor");
  FV_REQUIRE(d->x && d->w && d->bias && d->out32, FV_E_BADARG, "fv_mrf_fused: null pointer (x, w, bias, out32 required)");
  FV_REQUIRE(d->B > 0 && d->L > 0, FV_E_BADARG, "fv_mrf_fused: bad sizes B=%d L=%d", d->B, d->L);
  FV_REQUIRE(d->C == 16 || d->C == 32 || d->C == 64 || d->C == 128, FV_E_UNSUPPORTED,
             "fv_mrf_fused: C must be 16, 32, 64 or 128 (got %d)", d->C);
  FV_REQUIRE(d->n_blocks >= 1 && d->n_blocks <= FV_MRF_MAX_BLOCKS && d->n_pairs >= 1 && d->n_pairs <= FV_MRF_MAX_PAIRS,
             FV_E_BADARG, "fv_mrf_fused: n_blocks=%d n_pairs=%d out of range", d->n_blocks, d->n_pairs);
  FV_REQUIRE(d->act == FV_ACT_SILU || d->act == FV_ACT_LEAKY || d->act == FV_ACT_SILU_TANH || d->act == FV_ACT_SILU_H2,
             FV_E_UNSUPPORTED, "fv_mrf_fused: inner activation must be SiLU or leaky ReLU");
  FV_REQUIRE((d->out_act >= FV_ACT_NONE && d->out_act <= FV_ACT_TANH) || d->out_act == FV_ACT_SILU_TANH ||
                 d->out_act == FV_ACT_SILU_H2,
             FV_E_BADARG, "fv_mrf_fused: bad out_act");
  FV_REQUIRE(d->x_pitch >= d->C && d->x_pitch % 4 == 0 && d->out32_pitch >= d->C && d->out32_pitch % 4 == 0 &&
                 (!d->out16 || (d->out16_pitch >= d->C && d->out16_pitch % 8 == 0)),
             FV_E_ALIGN, "fv_mrf_fused: bad pitches");
  FV_REQUIRE(((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->w) |
               reinterpret_cast<uintptr_t>(d->out32) | reinterpret_cast<uintptr_t>(d->out16)) & 15) == 0,
             FV_E_ALIGN, "fv_mrf_fused: pointers must be 16-byte aligned");
  MrfParams p;
  memset(&p, 0, sizeof(p));
  int halo = 0;
  for (int j = 0; j < d->n_blocks; ++j) {
    const int k = d->ksize[j];
    FV_REQUIRE(k >= 1 && (k & 1), FV_E_BADARG, "fv_mrf_fused: kernel size %d must be odd", k);
    int h = 0;
    for (int i = 0; i < d->n_pairs; ++i) {
      const int d1 = d->dil1[j][i], d2 = d->dil2[j][i];
      FV_REQUIRE(d1 >= 1 && d2 >= 1 && (k - 1) / 2 * d1 <= kFrGuard && (k - 1) / 2 * d2 <= kFrGuard, FV_E_UNSUPPORTED,
                 "fv_mrf_fused: tap reach (k=%d, dilations %d/%d) exceeds %d rows", k, d1, d2, kFrGuard);
      h += (k - 1) / 2 * (d1 + d2);
      p.dil[j][i][0] = d1;
      p.dil[j][i][1] = d2;
      for (int c = 0; c < 2; ++c) {
        const int r0 = d->w_row0[j][i][c];
        FV_REQUIRE(r0 >= 0 && r0 + k * d->C <= d->w_rows, FV_E_BADARG, "fv_mrf_fused: weight rows out of range");
        p.w_row0[j][i][c] = r0;
      }
    }
    halo = h > halo ? h : halo;
    p.ksize[j] = k;
  }
  const int tile_rows = tile_rows_for(d);
  p.h0 = round_up(halo, 32);
  p.V = (tile_rows - p.h0 - halo) / 32 * 32;
  FV_REQUIRE(p.V >= 32, FV_E_UNSUPPORTED, "fv_mrf_fused: receptive field (%d rows per side) too large for a %d-row tile",
             halo, tile_rows);
  p.B = d->B;
  p.L = d->L;
  p.tiles_per_b = ceil_div(d->L, p.V);
  const long long total = (long long)p.tiles_per_b * d->B;
  FV_REQUIRE(total < (1ll << 30), FV_E_BADARG, "fv_mrf_fused: too many tiles");
  p.total_tiles = (int)total;
  p.n_blocks = d->n_blocks;
  p.n_pairs = d->n_pairs;
  p.bias = d->bias;
  p.act = d->act;
  p.act_param = d->act_param;
  p.out_act = d->out_act;
  p.out_act_param = d->out_act_param;
  p.has_o16 = d->out16 != nullptr;
  p.out_scale = d->out_scale != 0.f ? d->out_scale : 1.0f / (float)d->n_blocks;
  p.accumulate = d->accumulate ? 1 : 0;
  p.use_ws = g_mrf_ws ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  // C = 64: the residual stream + conv1 accumulator of a 512-row tile fill TMEM -> one CTA per SM, 16 epilogue warps;
  // C = 32 / 16: two co-resident CTAs with 8 epilogue warps each (MMA of one overlaps the epilogue of the other);
  // C = 128: 256-row tiles (X + T = 512 TMEM columns), operand rows in two K halves, at most one pair (2 convs) per launch:
  //          the host runs the stage pair by pair (halo <= 30 rows instead of 60 for the whole chain)
#define FV_MRF_DISPATCH(CC, EW, NCTA, MB, MAXCONV)                                                            \
  do {                                                                                                         \
    if (d->act == FV_ACT_SILU) return launch_mrf<CC, kActSilu, EW, NCTA, MB, MAXCONV>(d, p, st);              \
    if (d->act == FV_ACT_SILU_TANH) return launch_mrf<CC, kActSiluTanh, EW, NCTA, MB, MAXCONV>(d, p, st);     \
    if (d->act == FV_ACT_SILU_H2) return launch_mrf<CC, kActSiluH2, EW, NCTA, MB, MAXCONV>(d, p, st);         \
    return launch_mrf<CC, kActLeaky, EW, NCTA, MB, MAXCONV>(d, p, st);                                        \
  } while (0)
  if (d->C == 128) FV_MRF_DISPATCH(128, 16, 1, 2, 2);
  if (d->C == 64 && pair_config(d)) FV_MRF_DISPATCH(64, 8, 2, 2, 2);
  if (d->C == 64) FV_MRF_DISPATCH(64, 16, 1, 4, kFrMaxConvs);
  if (d->C == 16) FV_MRF_DISPATCH(16, 8, 2, 4, kFrMaxConvs);  // 32-byte operand rows, 16-column patches, two CTAs per SM
  if (g_mrf_c32_ctas == 1) FV_MRF_DISPATCH(32, 16, 1, 4, kFrMaxConvs);
  FV_MRF_DISPATCH(32, 8, 2, 4, kFrMaxConvs);
#undef FV_MRF_DISPATCH
}
