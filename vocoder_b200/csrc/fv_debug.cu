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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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
