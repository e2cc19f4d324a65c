// fv_conv_tc.cu - the hot op of the generator forward: (dilated | polyphase-transposed | pointwise) Conv1d as an
// implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM), operands
// staged by TMA into 128B-swizzled shared memory, fused bias / layer-scale / residual / MRF-mean / activation
// epilogue.  One persistent CTA per SM, warp-specialised:
//     warp 0    : TMA producer  (one elected lane)
//     warp 1    : tcgen05.mma issuer (one elected lane) + TMEM allocator
//     warps 2-9 : epilogue, two warps per TMEM lane quarter: TMEM -> registers -> per-warp smem transpose ->
//                 coalesced global I/O (every warp instruction touches whole 128-byte rows of the channels-last
//                 tensors: residual / MRF-accumulator reads, fp32 + fp16 writes)
//
// GEMM view (time on M, C_out on N, K = taps x C_in):
//     D[q, o] = sum_tap sum_c A[b, q + off[phase][tap], c] * W[phase][tap][o][c]
// A is the channels-last fp16 activation [B][L][pitch]: a tap / dilation shift is a different TMA box origin along
// L, and rows outside [0, L) are zero-filled by TMA = the reference's "same" zero padding for free.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv1d / nn.ConvTranspose1d / nn.Linear in
// fish_vocoder/modules/generators/hifigan.py:29-98,158-187,214-222, bigvgan.py:149-218,
// encoders/convnext.py:104-116,160-177 and generators/vocos.py:41,55.
#include <cstdlib>
#include <mutex>

#include "fv_common.cuh"

namespace fv {

constexpr int kEpiWarps = 8;            // a multiple of 4: kEpiWarps / 4 warps share each TMEM lane quarter
constexpr int kEpiStride = kEpiWarps / 4;  // work items (sub-tile, column chunk) are dealt round-robin to them
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTcThreads = 64 + kEpiThreads;
constexpr int kSmemLimit = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct ConvTcParams {
  CUtensorMap tmA;  // 3D {pitch, L_in, B} fp16, box {BLOCK_K, rows, 1}
  CUtensorMap tmW;  // 2D {w_pitch, n_phase*n_taps*C_out_pad} fp16, box {BLOCK_K, BLOCK_N}
  CUtensorMap tmA2; // slab mainloop: same tensor as tmA, box {BLOCK_K, 64 rows, 1}
  // epilogue maps, 4D {C_out_r8, n_phase, L_out / n_phase, B}, box {32, 1, 32, 1} (TMA epilogue only)
  CUtensorMap tmR;    // residual fp32 (SWIZZLE_128B)
  CUtensorMap tmO32;  // out32 fp32    (SWIZZLE_128B): accumulate-load and store
  CUtensorMap tmO16;  // out16 fp16    (SWIZZLE_64B)
  int B, n_phase, n_taps, k_chunks, C_out, C_out_r8, C_out_pad, q_rows, L_out;
  int m_tiles, n_tiles, total_tiles;
  int a_split, out16_split;  // strict precision: [hi | lo] operand / output layout (0 = plain fp16), see fv_conv_desc
  int reverse;   // walk the tiles from the last to the first (see next_tile_direction)
  int use_pair;  // host only: launch the cta_group::2 variant when the tile shape has one
  int use_slab, off_min, slab_boxes, w_resident;  // slab mainloop: smallest tap offset, 64-row boxes per slab, weights stay in smem
  const float* bias;
  const float* gamma;
  const float* residual;
  float* out32;
  __half* out16;
  int res_pitch, out32_pitch, out16_pitch, accumulate, act;
  float out_scale, act_param;
  int16_t tap_off[FV_MAX_TAPS];
};

template <int BLOCK_N, int M_SUB, int BLOCK_K, bool EPI_TMA, bool SLAB = false, bool PAIR = false, bool FAT = false>
struct TcCfg {
  static constexpr int EW = FAT ? 16 : kEpiWarps;  // epilogue warps (FAT: the fp16-only epilogue with 4 warps per scheduler)
  static constexpr int ROW_BYTES = BLOCK_K * 2;
  static constexpr int A_SUB_BYTES = 128 * ROW_BYTES;
  static constexpr int A_STAGE = M_SUB * A_SUB_BYTES;
  static constexpr int A_BOX_ROWS = (M_SUB * 128 <= 256) ? M_SUB * 128 : 256;
  static constexpr int N_A_BOX = M_SUB * 128 / A_BOX_ROWS;
  static constexpr int B_STAGE = (BLOCK_N / (PAIR ? 2 : 1)) * ROW_BYTES;  // CTA pair: each CTA stages half of the N rows
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int ACC_COLS = M_SUB * BLOCK_N;
  static constexpr int ACC_BUFS = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS_RAW = ACC_BUFS * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int CH = BLOCK_N < 32 ? BLOCK_N : 32;  // epilogue column chunk
  static constexpr int STG_STRIDE = CH + 1;               // padded row stride (words) of the transpose buffer
  // epilogue staging.  LSU flavour: per-warp padded transpose patch.  TMA flavour: per-warp swizzled boxes, two
  // 32x32 fp32 patches (4 KB each, SWIZZLE_128B: residual-in / out32, or ping-pong outputs when there is no
  // residual) and a 32x32 fp16 patch (2 KB, SWIZZLE_64B).
  // FAT: two ping-pong 32x32 fp16 patches (2 KB each, SWIZZLE_64B) per warp
  static constexpr int EPI_WARP_BYTES = FAT ? 4096 : (EPI_TMA ? (2 * 4096 + 2048) : (32 * STG_STRIDE * 4));
  static constexpr int STG_BYTES = ((EW * EPI_WARP_BYTES + 1023) / 1024) * 1024;
  static constexpr int TAIL_BYTES = 1024 + 2 * BLOCK_N * 4;  // barriers, tmem ptr, bias/gamma
  // per-tap mainloop: ring of (operand tile, weight tile) stages.  Slab mainloop: two operand slabs of
  // M_SUB*128 + 64 rows (one per K chunk, every tap is a row-shifted UMMA descriptor into it) + a ring of weight tiles.
  static constexpr int SLAB_ROWS = M_SUB * 128 + 64;
  static constexpr int A_SLAB = SLAB_ROWS * ROW_BYTES;
  static constexpr int NB_RAW = (kSmemLimit - TAIL_BYTES - STG_BYTES - 2 * A_SLAB) / B_STAGE;
  static constexpr int NB = NB_RAW > 16 ? 16 : NB_RAW;
  static constexpr int STAGES_RAW = (kSmemLimit - TAIL_BYTES - STG_BYTES) / STAGE;
  static constexpr int STAGES = SLAB ? NB : (STAGES_RAW > 8 ? 8 : STAGES_RAW);
  static constexpr int MAIN_BYTES = SLAB ? (2 * A_SLAB + NB * B_STAGE) : STAGES * STAGE;
  static constexpr int SMEM_BYTES = MAIN_BYTES + STG_BYTES + TAIL_BYTES;
  static constexpr int NCH = BLOCK_N / CH;                // column chunks per 128-row accumulator
  static constexpr int LPR = CH / 4;                      // lanes per row in the coalesced phase (float4 each)
  static constexpr int RPI = 32 / LPR;                    // rows per warp instruction
  static constexpr int ITERS = 32 / RPI;                  // warp instructions per 32-row chunk
  static constexpr bool VALID = STAGES >= 2;  // configurations whose pipeline does not fit are never instantiated
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert(TMEM_COLS_RAW <= 512, "accumulators exceed TMEM");
};

// HEAVY_ACT = false keeps only the cheap activations (none / SiLU / leaky / GELU) in the epilogue body; tanh and the
// polar map (expf + sincosf with its Payne-Hanek slow path) live in the HEAVY_ACT = true instantiations, so the hot
// kernels stay small enough for the instruction cache.
// PAIR = true: the kernel is launched in clusters of two CTAs (the two SMs of a TPC) that share every tile: 2 x M_SUB x 128
// rows by BLOCK_N columns, `tcgen05.mma.cta_group::2` (M = 256) issued by the leader CTA.  Each CTA stages its own rows
// of A and HALF of the weight tile, so a 128x256x16 step reads 4 + 4 KB of shared memory per SM instead of 4 + 8 KB
// (a 256-wide SS-mode UMMA on one SM sits at 96 B/clk of the 128 B/clk port before the TMA writes are counted).
// FAT = true: launches whose only output is the activated fp16 operand (convs1, conv_pre, ConvNeXt pwconv1: no residual, no
// fp32 output, no layer scale) are bound by the epilogue when it runs on 8 warps = 2 per scheduler (ncu: ~5 stall cycles per
// issued instruction, a 12032 x 5632 GELU GEMM tile takes 18k cycles against 11k of MMA).  This variant runs 16 epilogue
// warps (576 threads) over a lean fp16-only path with two ping-pong 2 KB staging patches per warp.
template <int BLOCK_N, int M_SUB, int BLOCK_K, bool EPI_TMA, bool HEAVY_ACT, bool SLAB, bool PAIR = false, bool FAT = false>
__global__ void __launch_bounds__(64 + (FAT ? 16 : kEpiWarps) * 32, 1)
    conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using Cfg = TcCfg<BLOCK_N, M_SUB, BLOCK_K, EPI_TMA, SLAB, PAIR, FAT>;
  static_assert(!FAT || (EPI_TMA && !HEAVY_ACT), "the 16-warp epilogue is the fp16-only TMA flavour");
  constexpr int EW = Cfg::EW, ESTRIDE = EW / 4, ETHREADS = EW * 32;
  static_assert(!PAIR || (!SLAB && EPI_TMA), "the CTA-pair variant exists for the per-tap mainloop with the TMA epilogue");
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;   // tiles are dealt to CTAs, or to CTA pairs
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  constexpr int TILE_ROWS = (PAIR ? 2 : 1) * M_SUB * 128;                 // rows of one tile (of the pair)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) {  // swizzled TMA/UMMA tiles need a 1024-byte aligned window
    if (threadIdx.x == 0) printf("fv: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* s_stage_raw = smem + Cfg::MAIN_BYTES;  // 1024-aligned (tile sizes are multiples of the swizzle atom)
  uint8_t* tail = s_stage_raw + Cfg::STG_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);   // per-tap: stage ring; slab: weight-tile ring
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* afull_bar = empty_bar + Cfg::STAGES;            // slab mainloop: the two operand slabs
  uint64_t* aempty_bar = afull_bar + 2;
  uint64_t* tfull_bar = aempty_bar + 2;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* epi_bar = tempty_bar + 2;  // two per epilogue warp (TMA epilogue: residual prefetch, running-sum load)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_bar + 2 * EW);
  float* s_bias = reinterpret_cast<float*>(tail + 1024);
  float* s_gamma = s_bias + BLOCK_N;
  float* s_stage = reinterpret_cast<float*>(s_stage_raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();  // the next kernel may start its prologue under this kernel's tail

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmW);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], (PAIR ? 2 : 1) * ETHREADS);  // pair: both CTAs' epilogues release the leader
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    if (SLAB) tma_prefetch_desc(&p.tmA2);
    for (int i = 0; i < 2 * EW; ++i) mbar_init(&epi_bar[i], 1);
    if (EPI_TMA) {
      if (p.residual) tma_prefetch_desc(&p.tmR);
      if (p.out32) tma_prefetch_desc(&p.tmO32);
      if (p.out16) tma_prefetch_desc(&p.tmO16);
    }
    fence_barrier_init();
  } else if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // barrier inits of BOTH CTAs are visible before any remote arrive / TMA
  __syncthreads();                         // (pair: the cluster barrier already orders this; the CTA barrier is what tools see)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above touched only this CTA's smem / TMEM; global memory of the previous kernel from here on

  const int k_steps = p.n_taps * p.k_chunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp runs the loops and the barrier waits; one elected lane issues the bulk copies)
    {
      const bool leader = elect_one();
      uint32_t it = 0, ita = 0;
      bool w_loaded = false;
      // strict precision: the weight K axis is [Whi | Whi | Wlo] (3P), the operand holds [hi | lo] (2P): the third
      // segment re-reads the hi block
      auto a_col = [&](int kc) {
        const int col = kc * BLOCK_K;
        return (p.a_split > 0 && col >= 2 * p.a_split) ? col - 2 * p.a_split : col;
      };
      for (int tile = worker; tile < p.total_tiles; tile += n_workers) {
        int r = p.reverse ? p.total_tiles - 1 - tile : tile;   // phase fastest: the phases of a polyphase ConvTranspose share their input rows and interleave their
        const int phase = r % p.n_phase; r /= p.n_phase;   // output rows, so they run side by side (L2 hits, merged lines)
        const int n_t = r % p.n_tiles; r /= p.n_tiles;
        const int m_t = r % p.m_tiles; r /= p.m_tiles;
        const int b = r;
        const int q0 = m_t * TILE_ROWS + (int)cta_rank * (M_SUB * 128);
        const int n0 = n_t * BLOCK_N;
        if constexpr (SLAB) {
          uint8_t* b_ring = smem + 2 * Cfg::A_SLAB;
          for (int kc = 0; kc < p.k_chunks; ++kc, ++ita) {
            const int sa = ita & 1;
            mbar_wait(&aempty_bar[sa], ((ita >> 1) & 1) ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&afull_bar[sa], p.slab_boxes * 64 * Cfg::ROW_BYTES);
              for (int bx = 0; bx < p.slab_boxes; ++bx)
                tma_load_3d(smem + sa * Cfg::A_SLAB + bx * 64 * Cfg::ROW_BYTES, &p.tmA2, &afull_bar[sa], a_col(kc),
                            q0 + p.off_min + bx * 64, b);
            }
            if (!(p.w_resident && w_loaded)) {
              for (int tap = 0; tap < p.n_taps; ++tap, ++it) {
                const int s = it % Cfg::NB;
                mbar_wait(&empty_bar[s], ((it / Cfg::NB) & 1) ^ 1);
                if (leader) {
                  mbar_arrive_expect_tx(&full_bar[s], Cfg::B_STAGE);
                  tma_load_2d(b_ring + s * Cfg::B_STAGE, &p.tmW, &full_bar[s], kc * BLOCK_K,
                              (phase * p.n_taps + tap) * p.C_out_pad + n0);
                }
              }
            }
          }
          w_loaded = true;
        } else {
          for (int tap = 0; tap < p.n_taps; ++tap) {
            const int row0 = q0 + p.tap_off[phase * p.n_taps + tap];
            const int wrow = (phase * p.n_taps + tap) * p.C_out_pad + n0;
            for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
              const int s = it % Cfg::STAGES;
              mbar_wait(&empty_bar[s], ((it / Cfg::STAGES) & 1) ^ 1);
              if (leader) {
                uint8_t* sa = smem + s * Cfg::STAGE;
                if constexpr (PAIR) {
                  // both CTAs' bytes complete on the leader's barrier; the leader arms it for the sum
                  if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE);
#pragma unroll
                  for (int bx = 0; bx < Cfg::N_A_BOX; ++bx)
                    tma_load_3d_pair(sa + bx * Cfg::A_BOX_ROWS * Cfg::ROW_BYTES, &p.tmA, &full_bar[s], a_col(kc),
                                     row0 + bx * Cfg::A_BOX_ROWS, b);
                  tma_load_2d_pair(sa + Cfg::A_STAGE, &p.tmW, &full_bar[s], kc * BLOCK_K,
                                   wrow + (int)cta_rank * (BLOCK_N / 2));
                } else {
                  mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE);
#pragma unroll
                  for (int bx = 0; bx < Cfg::N_A_BOX; ++bx)
                    tma_load_3d(sa + bx * Cfg::A_BOX_ROWS * Cfg::ROW_BYTES, &p.tmA, &full_bar[s], a_col(kc),
                                row0 + bx * Cfg::A_BOX_ROWS, b);
                  tma_load_2d(sa + Cfg::A_STAGE, &p.tmW, &full_bar[s], kc * BLOCK_K, wrow);
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // (whole warp runs the loops and the barrier waits; one elected lane issues tcgen05.mma / tcgen05.commit;
    //  CTA pair: only the leader CTA issues, its commits arrive on both CTAs' barriers)
    if (!PAIR || cta_rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, BLOCK_N);
      uint32_t it = 0, ita = 0, tile_i = 0;
      for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++tile_i) {
        const uint32_t buf = tile_i % Cfg::ACC_BUFS;
        mbar_wait(&tempty_bar[buf], ((tile_i / Cfg::ACC_BUFS) & 1) ^ 1);
        tc_fence_after();
        if constexpr (SLAB) {
          const int phase = (p.reverse ? p.total_tiles - 1 - tile : tile) % p.n_phase;
          const uint32_t b_ring = smem_u32(smem + 2 * Cfg::A_SLAB);
          for (int kc = 0; kc < p.k_chunks; ++kc, ++ita) {
            const int sa = ita & 1;
            mbar_wait(&afull_bar[sa], (ita >> 1) & 1);
            tc_fence_after();
            const uint32_t slab = smem_u32(smem + sa * Cfg::A_SLAB);
            for (int tap = 0; tap < p.n_taps; ++tap) {
              int s;
              if (p.w_resident) {
                s = kc * p.n_taps + tap;
                if (tile_i == 0) mbar_wait(&full_bar[s], 0);
              } else {
                s = it % Cfg::NB;
                mbar_wait(&full_bar[s], (it / Cfg::NB) & 1);
              }
              tc_fence_after();
              // the tap is a row shift into the slab: descriptor start += rows * row_bytes (swizzle is a function
              // of the absolute smem address, verified by fv_debug_rowshift_probe)
              const uint32_t a_tap = slab + (p.tap_off[phase * p.n_taps + tap] - p.off_min) * Cfg::ROW_BYTES;
              const uint32_t b_base = b_ring + s * Cfg::B_STAGE;
              if (leader) {
                const uint64_t da0 = make_kmajor_desc(a_tap, Cfg::ROW_BYTES);
                const uint64_t db0 = make_kmajor_desc(b_base, Cfg::ROW_BYTES);
#pragma unroll
                for (int sub = 0; sub < M_SUB; ++sub) {
#pragma unroll
                  for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
                    umma_f16_ss(tmem_base + (buf * M_SUB + sub) * BLOCK_N,
                                desc_advance(da0, sub * Cfg::A_SUB_BYTES + kk * 32), desc_advance(db0, kk * 32), idesc,
                                (kc > 0 || tap > 0 || kk > 0) ? 1u : 0u);
                  }
                }
                if (!p.w_resident) umma_commit(&empty_bar[s]);  // weight tile consumed
              }
              __syncwarp();
              if (!p.w_resident) ++it;
            }
            if (leader) umma_commit(&aempty_bar[sa]);  // slab consumed by every tap of this K chunk
            __syncwarp();
          }
        } else {
          for (int st = 0; st < k_steps; ++st, ++it) {
            const int s = it % Cfg::STAGES;
            mbar_wait(&full_bar[s], (it / Cfg::STAGES) & 1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + s * Cfg::STAGE);
            const uint32_t b_base = a_base + Cfg::A_STAGE;
            if (leader) {
              const uint64_t da0 = make_kmajor_desc(a_base, Cfg::ROW_BYTES);
              const uint64_t db0 = make_kmajor_desc(b_base, Cfg::ROW_BYTES);
#pragma unroll
              for (int sub = 0; sub < M_SUB; ++sub) {
#pragma unroll
                for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
                  if constexpr (PAIR)
                    umma_f16_ss_pair(tmem_base + (buf * M_SUB + sub) * BLOCK_N,
                                     desc_advance(da0, sub * Cfg::A_SUB_BYTES + kk * 32), desc_advance(db0, kk * 32),
                                     idesc, (st > 0 || kk > 0) ? 1u : 0u);
                  else
                    umma_f16_ss(tmem_base + (buf * M_SUB + sub) * BLOCK_N,
                                desc_advance(da0, sub * Cfg::A_SUB_BYTES + kk * 32), desc_advance(db0, kk * 32), idesc,
                                (st > 0 || kk > 0) ? 1u : 0u);
                }
              }
              // frees the smem stage (of both CTAs) once these MMAs have read it
              if constexpr (PAIR) umma_commit_pair(&empty_bar[s]);
              else umma_commit(&empty_bar[s]);
            }
            __syncwarp();
          }
        }
        if (leader) {  // accumulator complete -> epilogue (of both CTAs)
          if constexpr (PAIR) umma_commit_pair(&tfull_bar[buf]);
          else umma_commit(&tfull_bar[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    if constexpr (FAT) {
      // fp16-only epilogue on 16 warps: accumulator + bias -> * out_scale -> activation -> fp16 -> swizzled 32x32 patch ->
      // bulk store; thread = accumulator row, items (128-row block, 32-column chunk) dealt to the 4 warps of a lane quarter
      const int ew = warp - 2;
      const int quarter = warp & 3;
      const int cgrp = ew >> 2;
      const int tid_e = threadIdx.x - 64;
      uint8_t* H0 = s_stage_raw + ew * 4096;  // 32 rows x 64 B, SWIZZLE_64B, two of them
      uint32_t parity = 0;
      int staged_n_t = -1;
      const uint32_t h_xor = static_cast<uint32_t>((lane >> 1) & 3);
      uint32_t tile_i = 0;
      for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++tile_i) {
        int r = p.reverse ? p.total_tiles - 1 - tile : tile;   // phase fastest: the phases of a polyphase ConvTranspose share their input rows and interleave their
        const int phase = r % p.n_phase; r /= p.n_phase;   // output rows, so they run side by side (L2 hits, merged lines)
        const int n_t = r % p.n_tiles; r /= p.n_tiles;
        const int m_t = r % p.m_tiles; r /= p.m_tiles;
        const int b = r;
        const int q0 = m_t * TILE_ROWS + (int)cta_rank * (M_SUB * 128);
        const int n0 = n_t * BLOCK_N;
        const uint32_t buf = tile_i % Cfg::ACC_BUFS;
        if (n_t != staged_n_t) {  // bias of this N tile -> smem
          named_bar_sync(1, ETHREADS);
          for (int i = tid_e; i < BLOCK_N; i += ETHREADS) {
            const int col = n0 + i;
            s_bias[i] = (p.bias != nullptr && col < p.C_out) ? p.bias[col] : 0.f;
          }
          named_bar_sync(1, ETHREADS);
          staged_n_t = n_t;
        }
        int n_ch = (p.C_out_r8 - n0 + 31) / 32;  // column chunks of this tile that hold real channels
        n_ch = n_ch < Cfg::NCH ? n_ch : Cfg::NCH;
        const int n_items = M_SUB * n_ch;
        mbar_wait(&tfull_bar[buf], (tile_i / Cfg::ACC_BUFS) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int item = cgrp; item < n_items; item += ESTRIDE) {
          const int sub = item / n_ch, ch = item % n_ch;
          const int qb = q0 + sub * 128 + quarter * 32;
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (buf * M_SUB + sub) * BLOCK_N +
                                 ch * 32, acc);
          tmem_ld_wait();
          if (item + ESTRIDE >= n_items) {  // this warp's last TMEM read of the tile: hand the accumulator back
            tc_fence_before();
            if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[buf]); else mbar_arrive(&tempty_bar[buf]);
          }
          float o[32];
          const float4* bp = reinterpret_cast<const float4*>(s_bias + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = bp[j];
            o[4 * j + 0] = (__uint_as_float(acc[4 * j + 0]) + bb.x) * p.out_scale;
            o[4 * j + 1] = (__uint_as_float(acc[4 * j + 1]) + bb.y) * p.out_scale;
            o[4 * j + 2] = (__uint_as_float(acc[4 * j + 2]) + bb.z) * p.out_scale;
            o[4 * j + 3] = (__uint_as_float(acc[4 * j + 3]) + bb.w) * p.out_scale;
          }
          if (p.act == FV_ACT_SILU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = silu_fast(o[i]);
          } else if (p.act == FV_ACT_SILU_TANH) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = silu_tanh(o[i]);
          } else if (p.act == FV_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = gelu_erf_fast(o[i]);
          } else if (p.act == FV_ACT_LEAKY) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = o[i] > 0.f ? o[i] : o[i] * p.act_param;
          }
          uint8_t* Hout = H0 + parity * 2048;
          parity ^= 1;
          if (lane == 0) tma_store_wait_read_keep1();  // the store issued two items ago has drained this patch
          __syncwarp();
          const uint32_t h_row = smem_u32(Hout) + lane * 64;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t w0 = pack_half2_sat(o[8 * jj + 0], o[8 * jj + 1]);
            const uint32_t w1 = pack_half2_sat(o[8 * jj + 2], o[8 * jj + 3]);
            const uint32_t w2 = pack_half2_sat(o[8 * jj + 4], o[8 * jj + 5]);
            const uint32_t w3 = pack_half2_sat(o[8 * jj + 6], o[8 * jj + 7]);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h_row + ((static_cast<uint32_t>(jj) ^ h_xor) << 4)),
                         "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                         : "memory");
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the bulk-copy engine
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.tmO16, Hout, n0 + ch * 32, phase, qb, b);
            tma_store_commit();
          }
        }
        if (cgrp >= n_items) {  // idle warp of this tile shape still owes its TMEM-release arrivals
          tc_fence_before();
          if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[buf]); else mbar_arrive(&tempty_bar[buf]);
        }
      }
      if (lane == 0) tma_store_wait_all();
    } else if constexpr (EPI_TMA) {
      // thread = accumulator row throughout (the native TMEM layout); all global traffic of the epilogue is TMA:
      // the residual patch is bulk-loaded into a swizzled smem box, combined in place, and bulk-stored as out32;
      // the activated fp16 patch goes out through a second box.  No per-element address math, no LSU global ops.
      const int ew = warp - 2;
      const int quarter = warp & 3;
      const int cgrp = ew >> 2;
      const int tid_e = threadIdx.x - 64;
      uint8_t* R0 = s_stage_raw + ew * 8192;                       // 32 rows x 128 B, SWIZZLE_128B
      uint8_t* R1 = R0 + 4096;
      uint8_t* Hp = s_stage_raw + kEpiWarps * 8192 + ew * 2048;    // 32 rows x  64 B, SWIZZLE_64B
      uint64_t* my_bar = &epi_bar[ew];
      uint64_t* acc_bar = &epi_bar[kEpiWarps + ew];  // separate barrier: a residual prefetch may be in flight
      uint32_t ld_phase = 0, acc_phase = 0;
      uint32_t parity = 0;  // ping-pong of the output staging when no residual is loaded
      int staged_n_t = -1;
      const uint32_t r_xor = static_cast<uint32_t>(lane & 7);
      const uint32_t h_xor = static_cast<uint32_t>((lane >> 1) & 3);
      const bool has_res = p.residual != nullptr, has_o32 = p.out32 != nullptr, has_o16 = p.out16 != nullptr;
      uint32_t tile_i = 0;
      for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++tile_i) {
        int r = p.reverse ? p.total_tiles - 1 - tile : tile;   // phase fastest: the phases of a polyphase ConvTranspose share their input rows and interleave their
        const int phase = r % p.n_phase; r /= p.n_phase;   // output rows, so they run side by side (L2 hits, merged lines)
        const int n_t = r % p.n_tiles; r /= p.n_tiles;
        const int m_t = r % p.m_tiles; r /= p.m_tiles;
        const int b = r;
        const int q0 = m_t * TILE_ROWS + (int)cta_rank * (M_SUB * 128);
        const int n0 = n_t * BLOCK_N;
        const uint32_t buf = tile_i % Cfg::ACC_BUFS;

        if (n_t != staged_n_t) {  // bias / layer-scale of this N tile -> smem (once per CTA when there is one N tile)
          named_bar_sync(1, kEpiThreads);
          for (int i = tid_e; i < BLOCK_N; i += kEpiThreads) {
            const int col = n0 + i;
            s_bias[i] = (p.bias != nullptr && col < p.C_out) ? p.bias[col] : 0.f;
            s_gamma[i] = (p.gamma != nullptr && col < p.C_out) ? p.gamma[col] : 1.f;
          }
          named_bar_sync(1, kEpiThreads);
          staged_n_t = n_t;
        }

        int n_ch = (p.C_out_r8 - n0 + 31) / 32;  // column chunks of this tile that hold real channels
        n_ch = n_ch < Cfg::NCH ? n_ch : Cfg::NCH;
        const int n_items = M_SUB * n_ch;

        // residual patches are bulk-loaded into R0 one item ahead; nothing but these loads ever writes R0 in that
        // mode, so the next load is issued as soon as the current patch has been read into registers.
        auto issue_res_load = [&](int item) {  // lane 0 only
          const int sub = item / n_ch, ch = item % n_ch;
          mbar_arrive_expect_tx(my_bar, 4096);
          tma_load_4d(R0, &p.tmR, my_bar, n0 + ch * 32, phase, q0 + sub * 128 + quarter * 32, b);
        };
        if (has_res && cgrp < n_items && lane == 0) issue_res_load(cgrp);

        mbar_wait(&tfull_bar[buf], (tile_i / Cfg::ACC_BUFS) & 1);
        tc_fence_after();

#pragma unroll 1
        for (int item = cgrp; item < n_items; item += kEpiStride) {
          const int sub = item / n_ch, ch = item % n_ch;
          const int qb = q0 + sub * 128 + quarter * 32;
          const int col0 = n0 + ch * 32;
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (buf * M_SUB + sub) * BLOCK_N +
                                 ch * 32, acc);
          tmem_ld_wait();
          if (item + kEpiStride >= n_items) {
            tc_fence_before();
            if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[buf]); else mbar_arrive(&tempty_bar[buf]);
          }
          float o[32];
          const float4* bp = reinterpret_cast<const float4*>(s_bias + ch * 32);
          const float4* gp = reinterpret_cast<const float4*>(s_gamma + ch * 32);
          if (has_res) {
            mbar_wait(my_bar, ld_phase);
            ld_phase ^= 1;
          }
          const uint32_t r0_row = smem_u32(R0) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = bp[j];
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res) {
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rr.x), "=f"(rr.y), "=f"(rr.z), "=f"(rr.w)
                           : "r"(r0_row + ((static_cast<uint32_t>(j) ^ r_xor) << 4)));
            }
            o[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + bb.x;
            o[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + bb.y;
            o[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + bb.z;
            o[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + bb.w;
            if (p.gamma != nullptr) {
              const float4 gg = gp[j];
              o[4 * j + 0] *= gg.x; o[4 * j + 1] *= gg.y; o[4 * j + 2] *= gg.z; o[4 * j + 3] *= gg.w;
            }
            o[4 * j + 0] = (o[4 * j + 0] + rr.x) * p.out_scale;
            o[4 * j + 1] = (o[4 * j + 1] + rr.y) * p.out_scale;
            o[4 * j + 2] = (o[4 * j + 2] + rr.z) * p.out_scale;
            o[4 * j + 3] = (o[4 * j + 3] + rr.w) * p.out_scale;
          }
          if (has_res && item + kEpiStride < n_items) {  // R0 has been consumed: prefetch the next residual patch right away
            __syncwarp();
            if (lane == 0) issue_res_load(item + kEpiStride);
          }
          // output staging: with a residual R1 + Hp (single); without, ping-pong R0/R1 (and the fp16 patch in the
          // other half of the same buffer when there is no fp32 output) so stores of the previous item may still drain
          uint8_t* Rout = has_res ? R1 : (parity ? R1 : R0);
          uint8_t* Hout = (has_res || has_o32) ? Hp : Rout;
          const bool pingpong = !has_res && !(has_o32 && has_o16) && !p.accumulate && p.out16_split == 0;
          parity ^= 1;
          if (lane == 0) {
            if (pingpong) tma_store_wait_read_keep1();
            else tma_store_wait_read();
          }
          __syncwarp();
          const uint32_t r_row = smem_u32(Rout) + lane * 128;
          const uint32_t h_row = smem_u32(Hout) + lane * 64;
          if (p.accumulate) {  // MRF mean: add the running sum already in out32 (bulk load into the output patch)
            if (lane == 0) {
              mbar_arrive_expect_tx(acc_bar, 4096);
              tma_load_4d(Rout, &p.tmO32, acc_bar, col0, phase, qb, b);
            }
            mbar_wait(acc_bar, acc_phase);
            acc_phase ^= 1;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 oo;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(oo.x), "=f"(oo.y), "=f"(oo.z), "=f"(oo.w)
                           : "r"(r_row + ((static_cast<uint32_t>(j) ^ r_xor) << 4)));
              o[4 * j + 0] += oo.x; o[4 * j + 1] += oo.y; o[4 * j + 2] += oo.z; o[4 * j + 3] += oo.w;
            }
            __syncwarp();
          }
          if (has_o32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(r_row + ((static_cast<uint32_t>(j) ^ r_xor) << 4)),
                           "f"(o[4 * j + 0]), "f"(o[4 * j + 1]), "f"(o[4 * j + 2]), "f"(o[4 * j + 3])
                           : "memory");
          }
          if (has_o16) {
            if (p.act == FV_ACT_SILU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = silu_fast(o[i]);
            } else if (p.act == FV_ACT_SILU_TANH) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = silu_tanh(o[i]);
            } else if (p.act == FV_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = gelu_erf_fast(o[i]);
            } else if (p.act == FV_ACT_LEAKY) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = o[i] > 0.f ? o[i] : o[i] * p.act_param;
            }
            if constexpr (HEAVY_ACT) {
              if (p.act == FV_ACT_POLAR) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float m = fminf(expf(o[i]), 100.f);
                  float sn, cs;
                  sincosf(o[i + 1], &sn, &cs);
                  o[i] = m * cs;
                  o[i + 1] = m * sn;
                }
              } else if (p.act == FV_ACT_TANH) {
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = tanhf(o[i]);
              }
            }
            uint32_t hw[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) hw[i] = pack_half2_sat(o[2 * i], o[2 * i + 1]);
            if (p.out16_split > 0) {  // strict precision: o := v - fp16(v), stored as the "lo" word after the hi patch
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
                o[2 * i] -= f.x;
                o[2 * i + 1] -= f.y;
              }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h_row + ((static_cast<uint32_t>(jj) ^ h_xor) << 4)),
                           "r"(hw[4 * jj + 0]), "r"(hw[4 * jj + 1]), "r"(hw[4 * jj + 2]), "r"(hw[4 * jj + 3])
                           : "memory");
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the bulk-copy engine
          __syncwarp();
          if (lane == 0) {
            if (has_o32) tma_store_4d(&p.tmO32, Rout, col0, phase, qb, b);
            if (has_o16) tma_store_4d(&p.tmO16, Hout, col0, phase, qb, b);
            tma_store_commit();
          }
          if (has_o16 && p.out16_split > 0) {  // second pass through the same fp16 staging patch for the lo words
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint32_t w0 = pack_half2_sat(o[8 * jj + 0], o[8 * jj + 1]);
              const uint32_t w1 = pack_half2_sat(o[8 * jj + 2], o[8 * jj + 3]);
              const uint32_t w2 = pack_half2_sat(o[8 * jj + 4], o[8 * jj + 5]);
              const uint32_t w3 = pack_half2_sat(o[8 * jj + 6], o[8 * jj + 7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h_row + ((static_cast<uint32_t>(jj) ^ h_xor) << 4)),
                           "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&p.tmO16, Hout, p.out16_split + col0, phase, qb, b);
              tma_store_commit();
            }
          }
        }
        if (cgrp >= n_items) {
          tc_fence_before();
          if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[buf]); else mbar_arrive(&tempty_bar[buf]);
        }
      }
      if (lane == 0) tma_store_wait_all();  // global writes complete before the CTA retires
    } else {
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int cgrp = ew >> 2;                // work items (sub, chunk) are split between the two warps of a quarter
    const int tid_e = threadIdx.x - 64;      // 0..255
    float* stg = s_stage + ew * (32 * Cfg::STG_STRIDE);
    const int r_in = lane / Cfg::LPR;        // row within a warp instruction (coalesced phase)
    const int c4 = (lane % Cfg::LPR) * 4;    // first of this lane's 4 consecutive columns within the chunk
    constexpr int N_ITEMS = M_SUB * Cfg::NCH;
    uint32_t tile_i = 0;
    for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++tile_i) {
      int r = p.reverse ? p.total_tiles - 1 - tile : tile;   // phase fastest: the phases of a polyphase ConvTranspose share their input rows and interleave their
      const int phase = r % p.n_phase; r /= p.n_phase;   // output rows, so they run side by side (L2 hits, merged lines)
      const int n_t = r % p.n_tiles; r /= p.n_tiles;
      const int m_t = r % p.m_tiles; r /= p.m_tiles;
      const int b = r;
      const int q0 = m_t * TILE_ROWS + (int)cta_rank * (M_SUB * 128);
      const int n0 = n_t * BLOCK_N;
      const uint32_t buf = tile_i % Cfg::ACC_BUFS;
      const size_t brow0 = static_cast<size_t>(b) * p.L_out;

      named_bar_sync(1, kEpiThreads);  // previous tile's readers of s_bias/s_gamma are done
      for (int i = tid_e; i < BLOCK_N; i += kEpiThreads) {
        const int col = n0 + i;
        s_bias[i] = (p.bias != nullptr && col < p.C_out) ? p.bias[col] : 0.f;
        s_gamma[i] = (p.gamma != nullptr && col < p.C_out) ? p.gamma[col] : 1.f;
      }
      named_bar_sync(1, kEpiThreads);

      // residual prefetch for one work item: the rows/columns this lane handles in the coalesced phase
      auto load_res = [&](int item, float4 (&res)[Cfg::ITERS]) {
        const int sub = item / Cfg::NCH, ch = item % Cfg::NCH;
        const int col = n0 + ch * Cfg::CH + c4;
#pragma unroll
        for (int it = 0; it < Cfg::ITERS; ++it) {
          res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int q = q0 + sub * 128 + quarter * 32 + it * Cfg::RPI + r_in;
          const int orow = q * p.n_phase + phase;
          if (p.residual != nullptr && q < p.q_rows && orow < p.L_out && col < p.C_out_r8)
            res[it] = *reinterpret_cast<const float4*>(p.residual + (brow0 + orow) * p.res_pitch + col);
        }
      };

      float4 res_cur[Cfg::ITERS];
      if (cgrp < N_ITEMS) load_res(cgrp, res_cur);

      mbar_wait(&tfull_bar[buf], (tile_i / Cfg::ACC_BUFS) & 1);
      tc_fence_after();

#pragma unroll 1
      for (int item = cgrp; item < N_ITEMS; item += kEpiStride) {
        const int sub = item / Cfg::NCH, ch = item % Cfg::NCH;
        uint32_t acc[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               (buf * M_SUB + sub) * BLOCK_N + ch * Cfg::CH;
        if constexpr (Cfg::CH == 32) tmem_ld_32x32b_x32(taddr, acc);
        else tmem_ld_32x32b_x16(taddr, acc);
        tmem_ld_wait();
        if (item + kEpiStride >= N_ITEMS) {  // this warp's last TMEM read of the tile: hand the accumulator back
          tc_fence_before();
          if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[buf]); else mbar_arrive(&tempty_bar[buf]);
        }
        // transpose through this warp's private smem patch: thread = row  ->  lanes along the channel axis
        __syncwarp();
#pragma unroll
        for (int i = 0; i < Cfg::CH; ++i) stg[lane * Cfg::STG_STRIDE + i] = __uint_as_float(acc[i]);
        __syncwarp();

        // prefetch the next item's residual while this one is processed
        float4 res_nxt[Cfg::ITERS];
        const bool has_next = item + kEpiStride < N_ITEMS;
        if (has_next) load_res(item + kEpiStride, res_nxt);

        const int ccol = ch * Cfg::CH + c4;  // column within the N tile
        const int col = n0 + ccol;
        const bool col_ok = col < p.C_out_r8;
        const float4 bia = *reinterpret_cast<const float4*>(s_bias + ccol);
        const float4 gam = *reinterpret_cast<const float4*>(s_gamma + ccol);
#pragma unroll
        for (int it = 0; it < Cfg::ITERS; ++it) {
          const int rl = it * Cfg::RPI + r_in;
          const int q = q0 + sub * 128 + quarter * 32 + rl;
          const int orow = q * p.n_phase + phase;
          const bool ok = col_ok && (q < p.q_rows) && (orow < p.L_out);
          const size_t grow = brow0 + orow;
          const float* sp = stg + rl * Cfg::STG_STRIDE + c4;
          float o[4];
          o[0] = (sp[0] + bia.x) * gam.x + res_cur[it].x;
          o[1] = (sp[1] + bia.y) * gam.y + res_cur[it].y;
          o[2] = (sp[2] + bia.z) * gam.z + res_cur[it].z;
          o[3] = (sp[3] + bia.w) * gam.w + res_cur[it].w;
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] *= p.out_scale;
          if (ok) {
            if (p.out32 != nullptr) {
              float* dst = p.out32 + grow * p.out32_pitch + col;
              if (p.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(dst);
                o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
              }
              *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            }
            if (p.out16 != nullptr) {
              if (p.act == FV_ACT_SILU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = silu_fast(o[e]);
              } else if (p.act == FV_ACT_SILU_TANH) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = silu_tanh(o[e]);
              } else if (p.act == FV_ACT_GELU) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = gelu_erf_fast(o[e]);
              } else if (p.act == FV_ACT_LEAKY) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = o[e] > 0.f ? o[e] : o[e] * p.act_param;
              }
              if constexpr (HEAVY_ACT) {
                if (p.act == FV_ACT_POLAR) {
#pragma unroll
                  for (int e = 0; e < 4; e += 2) {
                    const float m = fminf(expf(o[e]), 100.f);
                    float sn, cs;
                    sincosf(o[e + 1], &sn, &cs);
                    o[e] = m * cs;
                    o[e + 1] = m * sn;
                  }
                } else if (p.act == FV_ACT_TANH) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) o[e] = tanhf(o[e]);
                }
              }
              uint2 pk;
              pk.x = pack_half2_sat(o[0], o[1]);
              pk.y = pack_half2_sat(o[2], o[3]);
              *reinterpret_cast<uint2*>(p.out16 + grow * p.out16_pitch + col) = pk;
              if (p.out16_split > 0) {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&pk.x));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&pk.y));
                uint2 lo;
                lo.x = pack_half2_sat(o[0] - f0.x, o[1] - f0.y);
                lo.y = pack_half2_sat(o[2] - f1.x, o[3] - f1.y);
                *reinterpret_cast<uint2*>(p.out16 + grow * p.out16_pitch + p.out16_split + col) = lo;
              }
            }
          }
        }
        if (has_next) {
#pragma unroll
          for (int it = 0; it < Cfg::ITERS; ++it) res_cur[it] = res_nxt[it];
        }
      }
      if (cgrp >= N_ITEMS) {  // idle warp of this tile shape still owes its TMEM-release arrivals
        tc_fence_before();
        if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[buf]); else mbar_arrive(&tempty_bar[buf]);
      }
    }
    }  // LSU epilogue
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may leave (or free TMEM) while the pair's MMAs / arrivals are in flight
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + dispatch
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 4D view of an output-side tensor [B][L_out][pitch]: {channel, phase, q, batch} with row = q * n_phase + phase
static int encode_epi_map(EncodeTiledFn enc, CUtensorMap* tm, const void* base, bool fp16, const fv_conv_desc* d,
                          int pitch, int c_out_r8) {
  const cuuint64_t es = fp16 ? 2 : 4;
  if (fp16 && d->out16_split > 0) c_out_r8 += d->out16_split;  // [hi | lo]: the lo words live out16_split columns on
  cuuint64_t dims[4] = {(cuuint64_t)c_out_r8, (cuuint64_t)d->n_phase, (cuuint64_t)(d->L_out / d->n_phase),
                        (cuuint64_t)d->B};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * d->n_phase,
                           (cuuint64_t)pitch * es * d->L_out};
  cuuint32_t box[4] = {32, 1, 32, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   fp16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(epilogue) failed: %d", (int)r);
  return 0;
}

template <int BLOCK_N, int M_SUB, int BLOCK_K, bool EPI_TMA, bool HEAVY_ACT, bool SLAB, bool PAIR = false, bool FAT = false>
static int launch_tc_impl(const fv_conv_desc* d, ConvTcParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<BLOCK_N, M_SUB, BLOCK_K, EPI_TMA, SLAB, PAIR, FAT>;
  constexpr int THREADS = 64 + Cfg::EW * 32;
  if constexpr (!Cfg::VALID) {
    return set_error(FV_E_UNSUPPORTED, "tile configuration N=%d M_SUB=%d K=%d does not fit in shared memory", BLOCK_N,
                     M_SUB, BLOCK_K);
  } else {
  EncodeTiledFn enc = get_encode_fn();
  FV_REQUIRE(enc != nullptr, FV_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");

  {  // activations: {pitch, L_in, B}
    cuuint64_t dims[3] = {(cuuint64_t)d->a_pitch, (cuuint64_t)d->L_in, (cuuint64_t)d->B};
    cuuint64_t strides[2] = {(cuuint64_t)d->a_pitch * 2, (cuuint64_t)d->a_pitch * 2 * (cuuint64_t)d->L_in};
    cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)Cfg::A_BOX_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(d->a), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(Cfg::ROW_BYTES), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  if constexpr (SLAB) {  // same activations, 64-row boxes (slab pieces)
    cuuint64_t dims[3] = {(cuuint64_t)d->a_pitch, (cuuint64_t)d->L_in, (cuuint64_t)d->B};
    cuuint64_t strides[2] = {(cuuint64_t)d->a_pitch * 2, (cuuint64_t)d->a_pitch * 2 * (cuuint64_t)d->L_in};
    cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, 64, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&p.tmA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(d->a), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(Cfg::ROW_BYTES), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(A slab) failed: %d", (int)r);
    int omin = d->tap_off[0], omax = d->tap_off[0];
    for (int i = 0; i < d->n_phase * d->n_taps; ++i) {
      omin = d->tap_off[i] < omin ? d->tap_off[i] : omin;
      omax = d->tap_off[i] > omax ? d->tap_off[i] : omax;
    }
    p.off_min = omin;
    p.slab_boxes = ceil_div(M_SUB * 128 + (omax - omin), 64);
  }
  {  // weights: {w_pitch, n_phase*n_taps*C_out_pad}
    cuuint64_t dims[2] = {(cuuint64_t)d->w_pitch, (cuuint64_t)d->n_phase * d->n_taps * d->C_out_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->w_pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(BLOCK_N / (PAIR ? 2 : 1))};  // pair: half the N rows per CTA
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(Cfg::ROW_BYTES), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
  }
  if constexpr (EPI_TMA) {
    int rc = 0;
    if (d->residual) rc = encode_epi_map(enc, &p.tmR, d->residual, false, d, d->res_pitch, p.C_out_r8);
    if (!rc && d->out32) rc = encode_epi_map(enc, &p.tmO32, d->out32, false, d, d->out32_pitch, p.C_out_r8);
    if (!rc && d->out16) rc = encode_epi_map(enc, &p.tmO16, d->out16, true, d, d->out16_pitch, p.C_out_r8);
    if (rc) return rc;
  }
  p.k_chunks = d->a_split > 0 ? (3 * d->a_split) / BLOCK_K : ceil_div(d->a_pitch, BLOCK_K);
  p.m_tiles = ceil_div(p.q_rows, (PAIR ? 2 : 1) * M_SUB * 128);
  p.n_tiles = ceil_div(d->C_out, BLOCK_N);
  FV_REQUIRE(p.n_tiles * BLOCK_N <= d->C_out_pad, FV_E_BADARG, "C_out_pad %d too small for %d tiles of %d",
             d->C_out_pad, p.n_tiles, BLOCK_N);
  const long long total = (long long)p.m_tiles * p.n_tiles * d->B * d->n_phase;
  FV_REQUIRE(total > 0 && total < (1ll << 30), FV_E_BADARG, "bad tile count %lld", total);
  p.total_tiles = (int)total;
  if constexpr (SLAB) {
    // all weight tiles of the conv fit in the ring and every tile of a CTA uses the same ones: load them once
    p.w_resident = (d->n_taps * p.k_chunks <= Cfg::NB && p.n_tiles == 1 && d->n_phase == 1) ? 1 : 0;
  }

  static std::atomic<unsigned long long> attr_done{0};
  int rc = check_cuda(ensure_dyn_smem(conv_tc_kernel<BLOCK_N, M_SUB, BLOCK_K, EPI_TMA, HEAVY_ACT, SLAB, PAIR, FAT>,
                                      Cfg::SMEM_BYTES, attr_done),
                      "cudaFuncSetAttribute(conv_tc_kernel)");
  if (rc) return rc;
  {
    // PAIR: clusters of two CTAs = the two SMs of a TPC, one tile per cluster at a time
    const int workers = PAIR ? num_sms() / 2 : num_sms();
    const int n_work = p.total_tiles < workers ? p.total_tiles : workers;
    rc = check_cuda(launch_kernel(conv_tc_kernel<BLOCK_N, M_SUB, BLOCK_K, EPI_TMA, HEAVY_ACT, SLAB, PAIR, FAT>,
                                  dim3((PAIR ? 2 : 1) * n_work), dim3(THREADS), Cfg::SMEM_BYTES, stream, PAIR ? 2 : 1, p),
                    "cudaLaunchKernelEx(conv_tc_kernel)");
    if (rc) return rc;
  }
  FV_CHECK_LAUNCH("conv_tc_kernel");
  return 0;
  }
}

template <int BLOCK_N, int M_SUB, int BLOCK_K, bool EPI_TMA>
static int launch_tc(const fv_conv_desc* d, ConvTcParams& p, cudaStream_t stream) {
  const bool heavy = d->out16 != nullptr && (d->act == FV_ACT_POLAR || d->act == FV_ACT_TANH);
  if (heavy) return launch_tc_impl<BLOCK_N, M_SUB, BLOCK_K, EPI_TMA, true, false>(d, p, stream);
  if constexpr (EPI_TMA) {
    // slab mainloop: every tap offset must fall inside the 64 extra slab rows
    int omin = d->tap_off[0], omax = d->tap_off[0];
    for (int i = 0; i < d->n_phase * d->n_taps; ++i) {
      omin = d->tap_off[i] < omin ? d->tap_off[i] : omin;
      omax = d->tap_off[i] > omax ? d->tap_off[i] : omax;
    }
    const bool slab_ok = (omax - omin) <= 64 && d->n_taps > 1;
    if constexpr (TcCfg<BLOCK_N, M_SUB, BLOCK_K, true, true>::VALID) {
      if (slab_ok && p.use_slab) return launch_tc_impl<BLOCK_N, M_SUB, BLOCK_K, true, false, true>(d, p, stream);
    }
  }
  if constexpr (EPI_TMA && BLOCK_K == 64 && ((BLOCK_N == 256 && M_SUB == 1) || (BLOCK_N == 128 && M_SUB == 2))) {
    // fp16-only output (convs1, conv_pre, pwconv1): the 16-warp epilogue (FV_TC_FAT=0 disables)
    static const bool fat_on = [] {
      const char* e = getenv("FV_TC_FAT");
      return !(e && e[0] == '0');
    }();
    const bool o16_only = fat_on && d->out16 && !d->out32 && !d->residual && !d->gamma && !d->accumulate &&
                          d->out16_split == 0 && d->a_split == 0 && !p.use_slab;
    if (o16_only) {
      if constexpr (BLOCK_N == 256) {
        if ((p.use_pair & 1) && p.q_rows >= 256)
          return launch_tc_impl<BLOCK_N, M_SUB, BLOCK_K, true, false, false, true, true>(d, p, stream);
      }
      return launch_tc_impl<BLOCK_N, M_SUB, BLOCK_K, true, false, false, false, true>(d, p, stream);
    }
    // CTA pair (cta_group::2) for the wide layers: 256 x 256 (GEMMs, C = 256 convs) or 512 x 128 (C = 128 convs) tiles
    // over two SMs; FV_TC_PAIR bit 0 enables the N = 256 shape, bit 1 the N = 128 shape
    if ((p.use_pair & (BLOCK_N == 256 ? 1 : 2)) && p.q_rows >= 2 * M_SUB * 128)
      return launch_tc_impl<BLOCK_N, M_SUB, BLOCK_K, true, false, false, true>(d, p, stream);
  }
  return launch_tc_impl<BLOCK_N, M_SUB, BLOCK_K, EPI_TMA, false, false>(d, p, stream);
}

template <int BLOCK_N, int M_SUB, bool EPI_TMA>
static int dispatch_k(const fv_conv_desc* d, ConvTcParams& p, cudaStream_t s) {
  // K chunk: 64 channels when the operand is wide enough; in strict mode a chunk must not straddle the hi/lo blocks
  int bk = d->a_pitch > 32 ? 64 : (d->a_pitch > 16 ? 32 : 16);
  if (d->a_split > 0) bk = (d->a_split % 64 == 0) ? 64 : ((d->a_split % 32 == 0) ? 32 : 16);
  if (bk == 64) return launch_tc<BLOCK_N, M_SUB, 64, EPI_TMA>(d, p, s);
  if constexpr (M_SUB >= 2) {  // small-K variants exist for the multi-accumulator tiles only
    if (bk == 32) return launch_tc<BLOCK_N, M_SUB, 32, EPI_TMA>(d, p, s);
    return launch_tc<BLOCK_N, M_SUB, 16, EPI_TMA>(d, p, s);
  } else {
    return set_error(FV_E_UNSUPPORTED, "no single-accumulator kernel for K chunk %d", bk);
  }
}

template <int BLOCK_N>
static int dispatch_m(const fv_conv_desc* d, ConvTcParams& p, cudaStream_t s, int m_sub, bool epi_tma) {
  if constexpr (BLOCK_N >= 32) {
    if (epi_tma) {
      if (m_sub == 1) return dispatch_k<BLOCK_N, 1, true>(d, p, s);
      if constexpr (BLOCK_N <= 64) {  // four 128-row accumulators: 512-row tiles amortise the per-tile cost at small C
        if (m_sub == 4) return dispatch_k<BLOCK_N, 4, true>(d, p, s);
        if constexpr (BLOCK_N == 32) {  // eight accumulators (1024-row tiles, 2 x 256 TMEM columns) for the narrowest layers
          if (m_sub == 8) return dispatch_k<BLOCK_N, 8, true>(d, p, s);
        }
      }
      return dispatch_k<BLOCK_N, 2, true>(d, p, s);
    }
  }
  if (m_sub == 1) return dispatch_k<BLOCK_N, 1, false>(d, p, s);
  return dispatch_k<BLOCK_N, 2, false>(d, p, s);
}

int pick_block_n(int C_out, int C_out_pad) {
  if (C_out <= 16 && C_out_pad < 32) return 16;  // weights packed for 16-row tiles only
  if (C_out <= 32) return 32;                    // N = 32 keeps the TMA epilogue (32-column boxes) even for C <= 16
  if (C_out <= 64) return 64;
  if (C_out <= 128) return 128;
  // 256-wide tiles (the CTA-pair kernel) unless they would pad C_out by more than 10%
  const int pad128 = round_up(C_out, 128), pad256 = round_up(C_out, 256);
  static const bool wide = [] {
    const char* e = getenv("FV_TC_WIDE");  // FV_TC_WIDE=0: 256-wide tiles only when they pad no more than 128-wide ones
    return !(e && e[0] == '0');
  }();
  return (pad256 <= pad128 || (wide && pad256 * 10 <= pad128 * 11)) ? 256 : 128;
}

// Called by fv_conv1d (fv_api.cu) after argument validation.  m_sub_override / block_n_override: 0 = heuristic.
int conv1d_tc(const fv_conv_desc* d, cudaStream_t stream, int block_n_override, int m_sub_override, int epilogue,
              int mainloop) {
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.n_phase = d->n_phase;
  p.n_taps = d->n_taps;
  p.C_out = d->C_out;
  p.C_out_r8 = round_up(d->C_out, 8);
  p.C_out_pad = d->C_out_pad;
  p.L_out = d->L_out;
  p.q_rows = ceil_div(d->L_out, d->n_phase);
  p.bias = d->bias;
  p.gamma = d->gamma;
  p.residual = d->residual;
  p.out32 = d->out32;
  p.out16 = reinterpret_cast<__half*>(d->out16);
  p.res_pitch = d->res_pitch;
  p.out32_pitch = d->out32_pitch;
  p.out16_pitch = d->out16_pitch;
  p.accumulate = d->accumulate;
  p.act = d->act;
  p.out_scale = d->out_scale;
  p.act_param = d->act_param;
  p.a_split = d->a_split;
  p.out16_split = d->out16_split;
  p.reverse = next_tile_direction();
  // mainloop: 0 = auto, 1 = per-tap stages, 2 = slab.  Measured on B200 (BigVGAN cfg C, per launch): for C_in >= 64 the two
  // are within 3% (C = 256: slab up to 20% slower), for C_in <= 32 with k >= 7 the slab saves 8-25% (one operand load
  // per tile instead of one small TMA stage per tap: 96.7 -> 76.2 us for C = 16, k = 11).  Auto picks accordingly.
  p.use_slab = mainloop == 2 || (mainloop == 0 && d->a_pitch <= 32 && d->n_taps >= 5 && d->n_phase == 1);
  for (int i = 0; i < d->n_phase * d->n_taps; ++i) p.tap_off[i] = (int16_t)d->tap_off[i];
  {
    static const int pair_mask = [] {
      const char* e = getenv("FV_TC_PAIR");  // FV_TC_PAIR=0 keeps every launch on single-CTA tiles (A/B measurements)
      return e ? atoi(e) : 1;
    }();
    p.use_pair = (d->a_split == 0 && !p.use_slab) ? pair_mask : 0;
  }

  int bn = block_n_override ? block_n_override : pick_block_n(d->C_out, d->C_out_pad);
  if (!block_n_override && bn >= 128 && d->a_split == 0) {
    // short sequences at small batch (test.py runs B = 1-2): a 256-wide tile leaves most SMs idle (C = 256, L = 752, B = 1:
    // 3 pair tiles on 148 SMs).  FV_TC_SPLITN=1 narrows the N tile until the launch has at least a quarter wave of tiles.
    // Measured on B200 (hifigan_b1, CUDA-graph replay): 0.832 ms without, 0.890 ms with - every narrower tile re-reads the
    // operand and a 64-wide UMMA costs 69 cycles against 161 for four times the columns - so it stays opt-in.
    static const bool split_n = [] {
      const char* e = getenv("FV_TC_SPLITN");
      return e && e[0] == '1';
    }();
    auto tiles_for = [&](int n) {
      const long long rows = (n == 256) ? 256 : (p.q_rows > 128 ? 256 : 128);
      return (long long)ceil_div(p.q_rows, (int)rows) * ceil_div(d->C_out, n) * d->B * d->n_phase;
    };
    while (split_n && bn > 64 && tiles_for(bn) * 4 < num_sms() && (d->C_out % (bn / 2)) == 0) bn /= 2;
  }
  // two 128-row accumulators per CTA share every weight tile; a single one when the sequence is short
  int m_sub = m_sub_override ? m_sub_override : ((p.q_rows > 128 && bn < 256) ? 2 : 1);
  if (bn == 256) m_sub = 1;  // two 256-column accumulators do not leave room for a pipelined smem ring
  if (!m_sub_override && bn <= 64 && p.q_rows >= 2048) m_sub = 4;
  {
    static const bool m8 = [] {
      // 1024-row tiles (eight accumulators) for N = 32, K <= 32 layers with long sequences: twice the epilogue items per
      // tile (all 8 epilogue warps busy) and half the per-tile latencies; BigVGAN cfg C 8.01 -> 7.83 ms.  FV_TC_M8=0 disables.
      const char* e = getenv("FV_TC_M8");
      return !(e && e[0] == '0');
    }();
    if (m8 && !m_sub_override && bn == 32 && d->a_pitch <= 32 && d->a_split == 0 && p.q_rows >= 8192 && (d->L_out % d->n_phase) == 0 && epilogue != 1) m_sub = 8;
  }
  if (d->a_pitch <= 32 && m_sub < 2) m_sub = 2;
  if (d->a_split > 0 && d->a_split % 64 != 0 && m_sub < 2) m_sub = 2;
  // TMA epilogue needs a rectangular {phase, q} view of the output rows; epilogue: 0 = auto, 1 = LSU, 2 = TMA
  const bool tma_ok = (d->L_out % d->n_phase) == 0 && bn >= 32;
  const bool epi_tma = tma_ok && epilogue != 1;
  switch (bn) {
    case 16: return dispatch_m<16>(d, p, stream, m_sub, epi_tma);
    case 32: return dispatch_m<32>(d, p, stream, m_sub, epi_tma);
    case 64: return dispatch_m<64>(d, p, stream, m_sub, epi_tma);
    case 128: return dispatch_m<128>(d, p, stream, m_sub, epi_tma);
    case 256: return dispatch_m<256>(d, p, stream, m_sub, epi_tma);
    default: return set_error(FV_E_BADARG, "unsupported BLOCK_N %d", bn);
  }
}

}  // namespace fv
