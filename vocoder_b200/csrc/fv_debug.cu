// fv_debug.cu - hardware probes used during bring-up (not on the product path).
//
// fv_debug_rowshift_probe: can a K-major SWIZZLE_128B shared-memory matrix descriptor start at an arbitrary ROW of
// a TMA-written slab (start address += r * 128 B), so that the taps of a convolution become descriptor offsets
// into ONE staged slab instead of one TMA box per tap?  For every shift r it runs D = A[r : r+128, :] * W^T twice:
// variant 0 leaves the descriptor's base-offset field 0, variant 1 sets it to (r & 7).  The host compares both with
// the exact answer.
#include "fv_common.cuh"

namespace fv {

constexpr int PR_ROWS = 144, PR_K = 64, PR_N = 64, PR_SHIFTS = 12;

struct ProbeParams {
  CUtensorMap tmA;  // 2D {64, rows} fp16, box {64, 144}
  CUtensorMap tmW;  // 2D {64, 64} fp16, box {64, 64}
  float* out;       // [PR_SHIFTS][2][128][64]
};

__global__ void __launch_bounds__(128, 1) rowshift_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                         // 144 rows * 128 B = 18432
  uint8_t* sW = smem + 18432;                 // 64 rows * 128 B = 8192   (18432 is a multiple of 1024)
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 18432 + 8192);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, 18432 + 8192);
    tma_load_2d(sA, &p.tmA, bar_load, 0, 0);
    tma_load_2d(sW, &p.tmW, bar_load, 0, 0);
  }
  mbar_wait(bar_load, 0);
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc_f16(128, PR_N);
  uint32_t mma_phase = 0;
  for (int r = 0; r < PR_SHIFTS; ++r) {
    for (int variant = 0; variant < 2; ++variant) {
      if (threadIdx.x == 0) {
        for (int kk = 0; kk < PR_K / 16; ++kk) {
          const uint64_t da = make_kmajor_desc(smem_u32(sA) + r * 128 + kk * 32, 128, variant ? (r & 7) : 0);
          const uint64_t db = make_kmajor_desc(smem_u32(sW) + kk * 32, 128);
          umma_f16_ss(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, mma_phase);
      mma_phase ^= 1;
      tc_fence_after();
      uint32_t acc[32];
      for (int ch = 0; ch < 2; ++ch) {
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, acc);
        tmem_ld_wait();
        float* dst = p.out + ((static_cast<size_t>(r) * 2 + variant) * 128 + warp * 32 + lane) * PR_N + ch * 32;
        for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(acc[i]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 64);
  }
}


// ------------------------------------------------------------------------------------------------
// fv_debug_umma_rate: how many SM cycles does one tcgen05.mma (M = 128 per CTA, K = 16, fp16) take for a given N and operand
// source, with the operands already resident?  One CTA per SM (or one CTA pair per TPC) issues `reps` x 16 UMMAs
// (4 accumulator blocks x 4 K steps, cycling through 4 A tiles and 2 B tiles so the shared-memory reads are real) and
// times the span from the first issue to the completion of the last one with clock64.
//   mode 0: SS, cta_group::1                      mode 1: SS, cta_group::2 (M = 256 over a CTA pair, B split)
//   mode 2: A from TMEM (TS), cta_group::1        mode 3: SS weight-stationary (tcgen05.mma.ws, B latched in a collector
//                                                         buffer per K step and reused by the 4 accumulator blocks)
//   bg    : 0 = idle epilogue warps; 1 = 8 warps stream st.shared.v4 into a scratch slab (an epilogue writing its
//           operand rows); 2 = 8 warps stream ld.shared.v4
// out[cta] = cycles per UMMA * 1000 (int).
// ------------------------------------------------------------------------------------------------
struct RateParams {
  int* out;
  int n, reps, mode, bg;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(320, 1) umma_rate_kernel(const RateParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // 4 A tiles of 128 x 64 fp16 (16 KB each), 2 B tiles of up to 256 x 64 (32 KB each), 32 KB scratch, barriers
  uint8_t* sA = smem;
  uint8_t* sB = smem + 4 * 16384;
  uint8_t* sS = sB + 2 * 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sS + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (4 * 16384 + 2 * 32768 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u | ((i * 2654435761u) & 0x03ff03ffu);  // fp16 values in [1, 2)
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    *stop = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int N = p.n;
  const int blocks = (4 * N <= 448) ? 4 : (2 * N <= 448 ? 2 : 1);
  if (warp == 1) {
    if (p.mode == 2) {  // park two A tiles (128 lanes x 32 packed columns) at TMEM columns 448..511
      uint32_t r[32];
      for (int i = 0; i < 32; ++i) r[i] = 0x3c003c00u | ((lane * 40503u + i * 9973u) & 0x03ff03ffu);
      // warp 1 owns lanes 32..63 only; the probe times the pipe, the values are irrelevant
      tmem_st_32x32b_x32(tmem_base + (32u << 16) + 448, r);
      tmem_st_32x32b_x32(tmem_base + (32u << 16) + 480, r);
      tmem_st_wait();
      tc_fence_before();
    }
    __syncwarp();
    if ((!PAIR || rank == 0) && elect_one()) {
      const uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, N);
      const long long t0 = clock64();
      if (p.mode == 3 && !PAIR) {
        // weight-stationary: K step outermost, the B chunk of a K step is filled once and used by all accumulator blocks
        for (int rep = 0; rep < p.reps; ++rep) {
          const uint64_t db0 = make_kmajor_desc(smem_u32(sB) + (rep & 1) * 32768, 128);
#define FV_WS_STEP(KK)                                                                                               \
  {                                                                                                                  \
    const uint64_t db = desc_advance(db0, KK * 32);                                                                  \
    for (int m = 0; m < blocks; ++m) {                                                                               \
      const uint64_t da = desc_advance(make_kmajor_desc(smem_u32(sA) + ((m + rep) & 3) * 16384, 128), KK * 32);      \
      const uint32_t d = tmem_base + m * N;                                                                          \
      if (m == 0) umma_f16_ws<KK, 0>(d, da, db, idesc, 1u);                                                          \
      else if (m == blocks - 1) umma_f16_ws<KK, 2>(d, da, db, idesc, 1u);                                            \
      else umma_f16_ws<KK, 1>(d, da, db, idesc, 1u);                                                                 \
    }                                                                                                                \
  }
          FV_WS_STEP(0) FV_WS_STEP(1) FV_WS_STEP(2) FV_WS_STEP(3)
#undef FV_WS_STEP
        }
      } else
      for (int rep = 0; rep < p.reps; ++rep) {
#pragma unroll 1
        for (int m = 0; m < blocks; ++m) {
          const uint64_t da0 = make_kmajor_desc(smem_u32(sA) + ((m + rep) & 3) * 16384, 128);
          const uint64_t db0 = make_kmajor_desc(smem_u32(sB) + (rep & 1) * 32768, 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t d = tmem_base + m * N;
            if (p.mode == 2) umma_f16_ts(d, tmem_base + 448 + (rep & 1) * 32 + kk * 8, desc_advance(db0, kk * 32), idesc, 1u);
            else if (PAIR) umma_f16_ss_pair(d, desc_advance(da0, kk * 32), desc_advance(db0, kk * 32), idesc, 1u);
            else umma_f16_ss(d, desc_advance(da0, kk * 32), desc_advance(db0, kk * 32), idesc, 1u);
          }
        }
      }
      if constexpr (PAIR) umma_commit_pair(&bar[0]);
      else umma_commit(&bar[0]);
      while (!mbar_try_wait(&bar[0], 0)) {
      }
      const long long t1 = clock64();
      p.out[blockIdx.x] = (int)((t1 - t0) * 1000 / ((long long)p.reps * blocks * 4));
      *stop = 1;
    } else if (PAIR && rank != 0 && lane == 0) {
      // the leader's commit is multicast to this CTA's barrier as well: release this CTA's background warps
      while (!mbar_try_wait(&bar[0], 0)) {
      }
      *stop = 1;
    }
    __syncwarp();
  } else if (warp >= 2 && p.bg != 0) {
    // background shared-memory traffic from 8 "epilogue" warps until the MMA thread is done
    const uint32_t base = smem_u32(sS) + (warp - 2) * 4096 + lane * 16;
    uint32_t acc = 0;
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (p.bg == 1) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 512), "r"(acc) : "memory");
        } else {
          uint32_t a, b, c, d;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + i * 512));
          acc += a + b + c + d;
        }
      }
    }
    if (acc == 0x12345678u) p.out[0] = (int)acc;
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// fv_debug_fma_rate: fp32 FMA throughput of the CUDA cores, the roofline of the anti-aliased Snake kernel (26 filter FMAs +
// 4 activation FMAs per sample, no tensor-core form: SURVEY B4).  Every thread runs 8 independent accumulator chains:
//   variant 0: scalar fma.rn.f32, all three operands in registers        variant 1: scalar, multiplier a kernel constant
//   variant 2: packed fma.rn.f32x2 (two fp32 lanes per instruction), multiplier a broadcast kernel constant
// out[0] = FMAs executed (lanes x 2 for the packed form), out[1] = elapsed SM cycles of block 0.
// ------------------------------------------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(256) fma_rate_kernel(float* sink, long long* out, int iters, float m0, float m1) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-3f + i;
    b[i] = 1.0f + i * 1e-3f;
  }
  float mv = m0 + threadIdx.x * 0.f;  // a register copy of the multiplier for the 3-register form
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (VARIANT == 0) {
        a[i] = fmaf(a[i], mv, b[i]);
        b[i] = fmaf(b[i], mv, a[i]);
      } else if constexpr (VARIANT == 1) {
        a[i] = fmaf(a[i], m0, b[i]);
        b[i] = fmaf(b[i], m1, a[i]);
      } else {
        unsigned long long ra, rb, rm;
        asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a[i]), "f"(b[i]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b[i]), "f"(a[i]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(rm) : "f"(m0));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(ra) : "l"(rm), "l"(rb));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(b[i]) : "l"(ra));
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i] + b[i];
  if (acc == 1.2345e-30f) sink[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = t1 - t0;
}

}  // namespace fv

using namespace fv;

extern "C" int fv_debug_rowshift_probe(const void* a16, const void* w16, float* out, void* stream) {
  FV_REQUIRE(a16 && w16 && out, FV_E_BADARG, "fv_debug_rowshift_probe: null pointer");
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FV_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
                 qres == cudaDriverEntryPointSuccess,
             FV_E_DRIVER, "cuTensorMapEncodeTiled unavailable");
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  p.out = out;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {PR_K, PR_ROWS};
    cuuint64_t strides[1] = {PR_K * 2};
    cuuint32_t box[2] = {PR_K, PR_ROWS};
    CUresult r = enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(a16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "encode A failed %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {PR_K, PR_N};
    cuuint64_t strides[1] = {PR_K * 2};
    cuuint32_t box[2] = {PR_K, PR_N};
    CUresult r = enc(&p.tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FV_REQUIRE(r == CUDA_SUCCESS, FV_E_DRIVER, "encode W failed %d", (int)r);
  }
  const int smem = 1024 + 18432 + 8192 + 64;
  rowshift_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  FV_CHECK_LAUNCH("rowshift_probe_kernel");
  return 0;
}

extern "C" int fv_debug_umma_rate(int mode, int n, int reps, int bg, int* out, void* stream) {
  FV_REQUIRE(out && (mode >= 0 && mode <= 3) && n >= 16 && n <= 256 && n % 16 == 0 && reps > 0 && bg >= 0 && bg <= 2,
             FV_E_BADARG, "fv_debug_umma_rate: bad arguments");
  RateParams p = {out, n, reps, mode, bg};
  const int smem = 4 * 16384 + 2 * 32768 + 32768 + 64;
  const int ctas = num_sms();
  cudaError_t e;
  if (mode == 1) {
    static std::atomic<unsigned long long> done{0};
    e = ensure_dyn_smem(umma_rate_kernel<true>, smem, done);
    if (e == cudaSuccess)
      e = launch_kernel(umma_rate_kernel<true>, dim3(ctas / 2 * 2), dim3(320), smem, (cudaStream_t)stream, 2, p);
  } else {
    static std::atomic<unsigned long long> done{0};
    e = ensure_dyn_smem(umma_rate_kernel<false>, smem, done);
    if (e == cudaSuccess)
      e = launch_kernel(umma_rate_kernel<false>, dim3(ctas), dim3(320), smem, (cudaStream_t)stream, 1, p);
  }
  int rc = check_cuda(e, "umma_rate_kernel");
  if (rc) return rc;
  FV_CHECK_LAUNCH("umma_rate_kernel");
  return 0;
}

extern "C" int fv_debug_fma_rate(int variant, int iters, float* sink, long long* out, void* stream) {
  FV_REQUIRE(sink && out && variant >= 0 && variant <= 2 && iters > 0, FV_E_BADARG, "fv_debug_fma_rate: bad arguments");
  const int blocks = num_sms() * 8;  // 8 x 256 threads per SM = full occupancy at <= 32 registers
  const float m0 = 0.999f, m1 = 1.001f;
  if (variant == 0) fma_rate_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, out, iters, m0, m1);
  else if (variant == 1) fma_rate_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, out, iters, m0, m1);
  else fma_rate_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, out, iters, m0, m1);
  FV_CHECK_LAUNCH("fma_rate_kernel");
  return 0;
}
