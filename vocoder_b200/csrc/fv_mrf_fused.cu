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
#include <mutex>

#include "fv_common.cuh"

namespace fv {

constexpr int kFrTile = 512;     // rows per tile = 4 UMMA M blocks of 128
constexpr int kFrMBlocks = 4;
constexpr int kFrGuard = 32;     // guard rows either side of an operand slab = largest tap reach supported
constexpr int kFrEpiWarps = 16;  // one epilogue thread per tile row
constexpr int kFrThreads = 128 + kFrEpiWarps * 32;
constexpr int kFrSmemLimit = 232448;
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
  float act_param, out_act_param, out_scale;
};

template <int C>
struct FrCfg {
  static constexpr int ROWB = C * 2;                          // bytes per operand row = swizzle span
  static constexpr int SLAB_ROWS = kFrTile + 2 * kFrGuard;
  static constexpr int SLAB = SLAB_ROWS * ROWB;               // multiple of 1024
  static constexpr int W_TILE = C * ROWB;
  static constexpr int STG32 = kFrEpiWarps * 4096;            // per-warp 32 x 32 fp32 patch, SWIZZLE_128B
  static constexpr int STG16 = kFrEpiWarps * 2048;            // per-warp 32 x 32 fp16 patch, SWIZZLE_64B
  // C = 64: the staging patches alias the operand slabs (dead at tile entry / exit); C = 32: own region
  static constexpr bool ALIAS = 2 * SLAB + STG32 + STG16 + 8 * W_TILE + 16384 > kFrSmemLimit;
  static constexpr int XA_OFF = 0;
  static constexpr int TA_OFF = SLAB;
  static constexpr int STG32_OFF = ALIAS ? TA_OFF : 2 * SLAB;
  static constexpr int STG16_OFF = ALIAS ? XA_OFF : 2 * SLAB + STG32;
  static constexpr int RING_OFF = ALIAS ? 2 * SLAB : 2 * SLAB + STG32 + STG16;
  static constexpr int BIAS_BYTES = kFrMaxConvs * C * 4;
  static constexpr int TAIL = 1024 + BIAS_BYTES;
  static constexpr int NS_RAW = (kFrSmemLimit - RING_OFF - TAIL) / W_TILE;
  static constexpr int NS = NS_RAW > 16 ? 16 : NS_RAW;
  static constexpr int SMEM = RING_OFF + NS * W_TILE + TAIL;
  static constexpr int TMEM_COLS = 2 * kFrMBlocks * C;        // 512 (C = 64) / 256 (C = 32)
  static_assert(!ALIAS || (STG32 <= SLAB && STG16 <= SLAB), "staging does not fit in the slabs it aliases");
  static_assert(NS >= 4, "weight ring too shallow");
  static_assert(SLAB % 1024 == 0 && W_TILE % 1024 == 0, "swizzled tiles must stay 1024-byte aligned");
};

// byte offset of 16-byte chunk `chunk` of operand row `row` inside a K-major swizzled slab (base 1024-aligned)
template <int ROWB>
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {
  if constexpr (ROWB == 128) return row * 128 + ((chunk ^ (row & 7)) << 4);
  else return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

template <int C>
__global__ void __launch_bounds__(kFrThreads, 1) mrf_fused_kernel(const __grid_constant__ MrfParams p) {
  using Cfg = FrCfg<C>;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int NCH = C / 32;  // 32-column chunks per row
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("fv: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* tail = smem + Cfg::RING_OFF + Cfg::NS * Cfg::W_TILE;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* w_empty = w_full + Cfg::NS;
  uint64_t* a_ready = w_empty + Cfg::NS;   // epilogue warps -> MMA: operand of the next conv is in smem / X is in TMEM
  uint64_t* acc_full = a_ready + 1;        // MMA -> epilogue: accumulators of the current conv are complete
  uint64_t* stg_bar = acc_full + 1;        // one per epilogue warp: TMA loads into its staging patch
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_bar + kFrEpiWarps);
  float* s_bias = reinterpret_cast<float*>(tail + 1024);  // [block][pair][2][C]: b1, cumulative b2

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmO32);
    if (p.has_o16) tma_prefetch_desc(&p.tmO16);
    for (int i = 0; i < Cfg::NS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(a_ready, kFrEpiWarps);
    mbar_init(acc_full, 1);
    for (int i = 0; i < kFrEpiWarps; ++i) mbar_init(&stg_bar[i], 1);
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
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
  // guard rows of both slabs: never written afterwards, only ever feed rows outside a tile's valid window
  for (int i = threadIdx.x; i < 2 * 2 * kFrGuard * ROWB / 16; i += blockDim.x) {
    const int per = kFrGuard * ROWB / 16;
    const int which = i / per, o = i % per;
    uint8_t* base = smem + ((which & 1) ? Cfg::TA_OFF : Cfg::XA_OFF) + ((which & 2) ? (kFrGuard + kFrTile) * ROWB : 0);
    reinterpret_cast<uint4*>(base)[o] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t X_COL = 0, T_COL = kFrMBlocks * C;

  if (warp == 0) {
    // ---------------------------------------------------------------- weight-tile producer
    const bool leader = elect_one();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int j = 0; j < p.n_blocks; ++j)
        for (int pi = 0; pi < p.n_pairs; ++pi)
          for (int cv = 0; cv < 2; ++cv) {
            const int row0 = p.w_row0[j][pi][cv];
            for (int tap = 0; tap < p.ksize[j]; ++tap, ++it) {
              const int s = it % Cfg::NS;
              mbar_wait(&w_empty[s], ((it / Cfg::NS) & 1) ^ 1);
              if (leader) {
                mbar_arrive_expect_tx(&w_full[s], Cfg::W_TILE);
                tma_load_2d(smem + Cfg::RING_OFF + s * Cfg::W_TILE, &p.tmW, &w_full[s], 0, row0 + tap * C);
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
            mbar_wait(a_ready, n & 1);
            tc_fence_after();
            const uint32_t slab = smem_u32(smem + (cv == 0 ? Cfg::XA_OFF : Cfg::TA_OFF)) + kFrGuard * ROWB;
            const uint32_t d_col = tmem_base + (cv == 0 ? T_COL : X_COL);
            const int dil = p.dil[j][pi][cv];
            for (int tap = 0; tap < k; ++tap, ++it) {
              const int s = it % Cfg::NS;
              mbar_wait(&w_full[s], (it / Cfg::NS) & 1);
              tc_fence_after();
              if (leader) {
                const uint64_t da0 = make_kmajor_desc(slab + (tap - half_k) * dil * ROWB, ROWB);
                const uint64_t db0 = make_kmajor_desc(smem_u32(smem + Cfg::RING_OFF + s * Cfg::W_TILE), ROWB);
#pragma unroll
                for (int m = 0; m < kFrMBlocks; ++m) {
#pragma unroll
                  for (int kk = 0; kk < C / 16; ++kk)
                    umma_f16_ss(d_col + m * C, desc_advance(da0, m * 128 * ROWB + kk * 32), desc_advance(db0, kk * 32),
                                idesc, (cv == 1 || tap > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&w_empty[s]);
              }
              __syncwarp();
            }
            if (leader) umma_commit(acc_full);
            __syncwarp();
          }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: thread = tile row
    const int e = warp - 4;             // rows 32e .. 32e+31 of the tile; TMEM lane quarter = warp % 4 = e % 4
    const int m = e >> 2;
    const int row_l = e * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((e & 3) * 32) << 16);
    uint8_t* patch32 = smem + Cfg::STG32_OFF + e * 4096;
    uint8_t* patch16 = smem + Cfg::STG16_OFF + e * 2048;
    const uint32_t p32_row = smem_u32(patch32) + lane * 128;
    const uint32_t p16_row = smem_u32(patch16) + lane * 64;
    const uint32_t r_xor = lane & 7, h_xor = (lane >> 1) & 3;
    const uint32_t xa_base = smem_u32(smem + Cfg::XA_OFF), ta_base = smem_u32(smem + Cfg::TA_OFF);
    const int srow = kFrGuard + row_l;  // this thread's row inside the operand slabs
    uint64_t* my_bar = &stg_bar[e];
    uint32_t stg_phase = 0, n = 0;
    const bool in_window = (e * 32 >= p.h0) && (e * 32 < p.h0 + p.V);

    // activation -> fp16 -> this thread's 64 bytes (32 columns) of an operand row
    auto put_operand = [&](uint32_t slab_base, int cc, const float (&v)[32]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w0 = pack_half2_sat(v[8 * q + 0], v[8 * q + 1]);
        const uint32_t w1 = pack_half2_sat(v[8 * q + 2], v[8 * q + 3]);
        const uint32_t w2 = pack_half2_sat(v[8 * q + 4], v[8 * q + 5]);
        const uint32_t w3 = pack_half2_sat(v[8 * q + 6], v[8 * q + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_base + swz_off<ROWB>(srow, cc * 4 + q)),
                     "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                     : "memory");
      }
    };
    auto publish = [&]() {  // operand rows / TMEM stores of this warp are done -> MMA warp
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    };

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_b;
      const int g0 = (tile % p.tiles_per_b) * p.V - p.h0;  // global row of tile row 0
      const int g = g0 + row_l;
      const bool in_seq = g >= 0 && g < p.L;
      const bool live = in_window && (g0 + e * 32 < p.L);
      for (int j = 0; j < p.n_blocks; ++j) {
        // ---- tile entry: x -> X (TMEM, fp32) and act(x) -> XA (fp16); rows outside the sequence arrive as zeros
        for (int cc = 0; cc < NCH; ++cc) {
          __syncwarp();
          if (lane == 0) {
            tma_store_wait_read();
            mbar_arrive_expect_tx(my_bar, 4096);
            tma_load_3d(patch32, &p.tmX, my_bar, cc * 32, g0 + e * 32, b);
          }
          __syncwarp();
          mbar_wait(my_bar, stg_phase);
          stg_phase ^= 1;
          uint32_t r[32];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[4 * q]), "=r"(r[4 * q + 1]), "=r"(r[4 * q + 2]), "=r"(r[4 * q + 3])
                         : "r"(p32_row + ((static_cast<uint32_t>(q) ^ r_xor) << 4)));
          tmem_st_32x32b_x32(t_lane + X_COL + m * C + cc * 32, r);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = act_apply(__uint_as_float(r[i]), p.act, p.act_param);
          put_operand(xa_base, cc, v);
        }
        tmem_st_wait();
        publish();
        for (int pi = 0; pi < p.n_pairs; ++pi) {
          const float* b1 = s_bias + ((j * p.n_pairs + pi) * 2) * C;
          const float* b2c = b1 + C;
          // ---- conv1 done: T + b1 -> act -> TA
          mbar_wait(acc_full, n & 1);
          ++n;
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < NCH; ++cc) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_lane + T_COL + m * C + cc * 32, r);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float t = __uint_as_float(r[i]) + b1[cc * 32 + i];
              v[i] = in_seq ? act_apply(t, p.act, p.act_param) : 0.f;
            }
            put_operand(ta_base, cc, v);
          }
          publish();
          // ---- conv2 done: X (+ cumulative b2) is the residual stream after this pair
          mbar_wait(acc_full, n & 1);
          ++n;
          tc_fence_after();
          if (pi + 1 < p.n_pairs) {
#pragma unroll 1
            for (int cc = 0; cc < NCH; ++cc) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(t_lane + X_COL + m * C + cc * 32, r);
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float t = __uint_as_float(r[i]) + b2c[cc * 32 + i];
                v[i] = in_seq ? act_apply(t, p.act, p.act_param) : 0.f;
              }
              put_operand(xa_base, cc, v);
            }
            publish();
          } else if (live) {
            // ---- tile exit: block output -> running mean in out32 (-> activated fp16 after the last block)
            const bool last = j == p.n_blocks - 1;
            const int grow = g0 + e * 32;
#pragma unroll 1
            for (int cc = 0; cc < NCH; ++cc) {
              __syncwarp();
              if (lane == 0) {
                if (j > 0) {
                  tma_store_wait_all();  // the partial sums this warp stored for block j-1 are visible
                  mbar_arrive_expect_tx(my_bar, 4096);
                  tma_load_3d(patch32, &p.tmO32, my_bar, cc * 32, grow, b);
                } else {
                  tma_store_wait_read();
                }
              }
              __syncwarp();
              uint32_t r[32];
              tmem_ld_32x32b_x32(t_lane + X_COL + m * C + cc * 32, r);
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = (__uint_as_float(r[i]) + b2c[cc * 32 + i]) * p.out_scale;
              if (j > 0) {
                mbar_wait(my_bar, stg_phase);
                stg_phase ^= 1;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  float4 a;
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                               : "r"(p32_row + ((static_cast<uint32_t>(q) ^ r_xor) << 4)));
                  v[4 * q] += a.x;
                  v[4 * q + 1] += a.y;
                  v[4 * q + 2] += a.z;
                  v[4 * q + 3] += a.w;
                }
              }
#pragma unroll
              for (int q = 0; q < 8; ++q)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(p32_row +
                                                                              ((static_cast<uint32_t>(q) ^ r_xor) << 4)),
                             "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                             : "memory");
              if (last && p.has_o16) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t w[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    w[u] = pack_half2_sat(act_apply(v[8 * q + 2 * u], p.out_act, p.out_act_param),
                                          act_apply(v[8 * q + 2 * u + 1], p.out_act, p.out_act_param));
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p16_row +
                                                                                ((static_cast<uint32_t>(q) ^ h_xor) << 4)),
                               "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                               : "memory");
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&p.tmO32, patch32, cc * 32, grow, b);
                if (last && p.has_o16) tma_store_3d(&p.tmO16, patch16, cc * 32, grow, b);
                tma_store_commit();
              }
            }
          }
          if (Cfg::ALIAS && pi + 1 == p.n_pairs && j == p.n_blocks - 1 && p.has_o16) {
            // the fp16 exit patches alias XA: nobody may start the next tile's entry before every store has read them
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
            named_bar_sync(1, kFrEpiWarps * 32);
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
                           int B) {
  const cuuint64_t es = fp16 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * (cuuint64_t)L};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   fp16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(mrf rows) failed: %d", (int)r);
  return 0;
}

template <int C>
static int launch_mrf(const fv_mrf_desc* d, MrfParams& p, cudaStream_t stream) {
  using Cfg = FrCfg<C>;
  EncodeTiledFn enc = get_encode_fn();
  FV_REQUIRE(enc != nullptr, FV_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  int rc = encode_rows_map(enc, &p.tmX, d->x, false, C, d->x_pitch, d->L, d->B);
  if (!rc) rc = encode_rows_map(enc, &p.tmO32, d->out32, false, C, d->out32_pitch, d->L, d->B);
  if (!rc && d->out16) rc = encode_rows_map(enc, &p.tmO16, d->out16, true, C, d->out16_pitch, d->L, d->B);
  if (rc) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)d->w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)C};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     Cfg::ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(mrf W) failed: %d", (int)r);
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(mrf_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  });
  rc = check_cuda(attr_err, "cudaFuncSetAttribute(mrf_fused_kernel)");
  if (rc) return rc;
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  mrf_fused_kernel<C><<<grid, kFrThreads, Cfg::SMEM, stream>>>(p);
  FV_CHECK_LAUNCH("mrf_fused_kernel");
  return 0;
}

}  // namespace fv

using namespace fv;

extern "C" int fv_mrf_fused(const fv_mrf_desc* d, void* stream) {
  FV_REQUIRE(d != nullptr, FV_E_BADARG, "fv_mrf_fused: null descriptor");
  FV_REQUIRE(d->x && d->w && d->bias && d->out32, FV_E_BADARG, "fv_mrf_fused: null pointer (x, w, bias, out32 required)");
  FV_REQUIRE(d->B > 0 && d->L > 0, FV_E_BADARG, "fv_mrf_fused: bad sizes B=%d L=%d", d->B, d->L);
  FV_REQUIRE(d->C == 32 || d->C == 64, FV_E_UNSUPPORTED, "fv_mrf_fused: C must be 32 or 64 (got %d)", d->C);
  FV_REQUIRE(d->n_blocks >= 1 && d->n_blocks <= FV_MRF_MAX_BLOCKS && d->n_pairs >= 1 && d->n_pairs <= FV_MRF_MAX_PAIRS,
             FV_E_BADARG, "fv_mrf_fused: n_blocks=%d n_pairs=%d out of range", d->n_blocks, d->n_pairs);
  FV_REQUIRE(d->act == FV_ACT_SILU || d->act == FV_ACT_LEAKY, FV_E_UNSUPPORTED,
             "fv_mrf_fused: inner activation must be SiLU or leaky ReLU");
  FV_REQUIRE(d->out_act >= FV_ACT_NONE && d->out_act <= FV_ACT_TANH, FV_E_BADARG, "fv_mrf_fused: bad out_act");
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
  p.h0 = round_up(halo, 32);
  p.V = (kFrTile - p.h0 - halo) / 32 * 32;
  FV_REQUIRE(p.V >= 32, FV_E_UNSUPPORTED, "fv_mrf_fused: receptive field (%d rows per side) too large for a %d-row tile",
             halo, kFrTile);
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
  p.out_scale = 1.0f / (float)d->n_blocks;
  if (d->C == 64) return launch_mrf<64>(d, p, (cudaStream_t)stream);
  return launch_mrf<32>(d, p, (cudaStream_t)stream);
}
