// fv_api.cu - C ABI plumbing: error reporting, launch accounting, argument validation + engine dispatch for
// fv_conv1d.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include <cstdlib>

#include "fv_common.cuh"

namespace fv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_tune_block_n{0};
static std::atomic<int> g_tune_m_sub{0};
static std::atomic<int> g_tune_epilogue{0};
static std::atomic<int> g_tune_mainloop{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Tile-order direction of the streaming kernels (fv_conv1d on tensor cores, fv_snake_aa): it alternates launch by launch, so a
// consumer starts with the tiles its producer wrote LAST - the part of a 90-200 MB intermediate that is still in the 126 MB L2
// - instead of the tiles written first, which have already been evicted.  FV_SERPENTINE=0 disables (A/B measurements).
static std::atomic<unsigned> g_direction{0};
int next_tile_direction() {
  static const bool on = [] {
    const char* e = getenv("FV_SERPENTINE");
    return !(e && e[0] == '0');
  }();
  return on ? (int)(g_direction.fetch_add(1, std::memory_order_relaxed) & 1u) : 0;
}

int conv1d_tc(const fv_conv_desc* d, cudaStream_t stream, int block_n_override, int m_sub_override, int epilogue,
              int mainloop);
int conv1d_simt(const fv_conv_desc* d, cudaStream_t stream);

}  // namespace fv

namespace fv {
bool pdl_enabled() {
  static const bool on = [] {
    // Measured on B200 under CUDA-graph replay (HiFiGAN b64 / b1, BigVGAN b32, Vocos b128): within run-to-run noise
    // (b1: 0.873 -> 0.852 ms), so the attribute is opt-in: FV_PDL=1
    const char* e = getenv("FV_PDL");
    return e && e[0] == '1';
  }();
  return on;
}
}  // namespace fv

using namespace fv;

extern "C" const char* fv_last_error(void) { return g_err; }
extern "C" int fv_abi_version(void) { return FV_ABI_VERSION; }
extern "C" int64_t fv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void fv_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }
extern "C" void fv_set_tc_tuning(int block_n, int m_sub, int epilogue, int mainloop) {
  g_tune_block_n.store(block_n);
  g_tune_m_sub.store(m_sub);
  g_tune_epilogue.store(epilogue);
  g_tune_mainloop.store(mainloop);
}

extern "C" int fv_conv1d(const fv_conv_desc* d, int engine, void* stream) {
  FV_REQUIRE(d != nullptr, FV_E_BADARG, "fv_conv1d: null descriptor");
  FV_REQUIRE(d->a && d->w && d->tap_off, FV_E_BADARG, "fv_conv1d: null operand pointer");
  FV_REQUIRE(d->B > 0 && d->L_in > 0 && d->L_out > 0 && d->C_out > 0, FV_E_BADARG,
             "fv_conv1d: bad sizes B=%d L_in=%d L_out=%d C_out=%d", d->B, d->L_in, d->L_out, d->C_out);
  FV_REQUIRE(d->n_phase >= 1 && d->n_taps >= 1 && d->n_phase * d->n_taps <= FV_MAX_TAPS, FV_E_BADARG,
             "fv_conv1d: n_phase*n_taps = %d*%d exceeds %d", d->n_phase, d->n_taps, FV_MAX_TAPS);
  FV_REQUIRE(d->a_pitch > 0 && d->a_pitch % 8 == 0 && d->w_pitch > 0 && d->w_pitch % 8 == 0, FV_E_ALIGN,
             "fv_conv1d: operand pitches must be multiples of 8 halfs (a=%d w=%d)", d->a_pitch, d->w_pitch);
  FV_REQUIRE(d->C_out_pad >= d->C_out && d->C_out_pad % 16 == 0, FV_E_BADARG, "fv_conv1d: bad C_out_pad %d",
             d->C_out_pad);
  FV_REQUIRE((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0,
             FV_E_ALIGN, "fv_conv1d: operand base pointers must be 16-byte aligned");
  const int r8 = round_up(d->C_out, 8);
  FV_REQUIRE(d->out32 || d->out16, FV_E_BADARG, "fv_conv1d: no output requested");
  if (d->out32)
    FV_REQUIRE(d->out32_pitch >= r8 && d->out32_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(d->out32) & 15) == 0,
               FV_E_ALIGN, "fv_conv1d: out32 pitch/alignment (pitch %d, need >= %d)", d->out32_pitch, r8);
  if (d->out16)
    FV_REQUIRE(d->out16_pitch >= r8 && d->out16_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out16) & 15) == 0,
               FV_E_ALIGN, "fv_conv1d: out16 pitch/alignment (pitch %d, need >= %d)", d->out16_pitch, r8);
  if (d->residual)
    FV_REQUIRE(d->res_pitch >= r8 && d->res_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(d->residual) & 15) == 0,
               FV_E_ALIGN, "fv_conv1d: residual pitch/alignment (pitch %d, need >= %d)", d->res_pitch, r8);
  if (d->a_split > 0)
    FV_REQUIRE(d->a_split % 16 == 0 && d->a_pitch == 2 * d->a_split && d->w_pitch >= 3 * d->a_split, FV_E_BADARG,
               "fv_conv1d: strict operand layout needs a_pitch == 2*a_split (16 | a_split) and w_pitch >= 3*a_split "
               "(a_split=%d a_pitch=%d w_pitch=%d)", d->a_split, d->a_pitch, d->w_pitch);
  if (d->out16_split > 0)
    FV_REQUIRE(d->out16 && d->out16_split % 8 == 0 && d->out16_split >= r8 && d->out16_pitch >= d->out16_split + r8,
               FV_E_BADARG, "fv_conv1d: strict output layout needs out16_pitch >= out16_split + C_out (split=%d pitch=%d)",
               d->out16_split, d->out16_pitch);
  FV_REQUIRE(d->act >= FV_ACT_NONE && d->act <= FV_ACT_SILU_TANH, FV_E_BADARG, "fv_conv1d: bad activation %d", d->act);
  FV_REQUIRE(!(d->accumulate && !d->out32), FV_E_BADARG, "fv_conv1d: accumulate needs out32");
  for (int i = 0; i < d->n_phase * d->n_taps; ++i)
    FV_REQUIRE(d->tap_off[i] > -30000 && d->tap_off[i] < 30000, FV_E_BADARG, "fv_conv1d: tap offset out of range");
  if (engine == FV_ENGINE_SIMT) return conv1d_simt(d, (cudaStream_t)stream);
  FV_REQUIRE(engine == FV_ENGINE_TC, FV_E_BADARG, "fv_conv1d: unknown engine %d", engine);
  return conv1d_tc(d, (cudaStream_t)stream, g_tune_block_n.load(), g_tune_m_sub.load(), g_tune_epilogue.load(),
                   g_tune_mainloop.load());
}
