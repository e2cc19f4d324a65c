// fv_frontend.cu - CUDA-core pieces of the mel front-end (the step immediately before the generator forward):
// LinearSpectrogram / LogMelSpectrogram, fish_vocoder/data/transforms/spectrogram.py:6-104.
//
//   y [B][L]  --fv_frame_audio-->  rows of `hop` samples of the reflect-padded signal, fp16 [hi | lo]
//             --fv_conv1d (n_fft/hop taps, strict precision)-->  windowed DFT  [B][T][re_0..re_F-1 | im_0..im_F-1]
//             --fv_spec_mag-->  sqrt(re^2 + im^2 + 1e-6)  fp16 [hi | lo] (+ fp32)
//             --fv_conv1d (pointwise, strict)-->  mel energies  --fv_log_mel_out-->  log(max(., 1e-5)) [B][n_mels][T]
//
// The framed DFT is a strided conv (kernel n_fft, stride hop); with the signal re-laid as rows of `hop` samples it is
// an ordinary conv over rows with n_fft / hop taps - the same tcgen05 implicit GEMM as every other layer (no cuFFT).
#include "fv_common.cuh"

namespace fv {

// out[b][r][c] = y_pad[r * hop + c], y_pad = F.pad(y, (pad_left, pad_right), mode="reflect") (spectrogram.py:29-37);
// rows past the padded signal are zero.  One thread per 2 samples.
__global__ void frame_audio_kernel(const float* __restrict__ y, __half* __restrict__ out, int L, int hop, int pad_left,
                                   int Lp, int R, int pitch, int split, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int hp = pitch >> 1;
  const int c = 2 * (int)(i % hp);
  const long long br = i / hp;
  const int r = (int)(br % R);
  const int b = (int)(br / R);
  float v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int p = r * hop + c + e;  // index into the padded signal
    float s = 0.f;
    if (c + e < hop && p < Lp) {
      int src = p - pad_left;
      if (src < 0) src = -src;
      if (src >= L) src = 2 * (L - 1) - src;
      s = y[(size_t)b * L + src];
    }
    v[e] = s;
  }
  __half* dst = out + (size_t)br * (pitch + split) + c;
  const uint32_t h = pack_half2_sat(v[0], v[1]);
  *reinterpret_cast<uint32_t*>(dst) = h;
  if (split > 0) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    *reinterpret_cast<uint32_t*>(dst + split) = pack_half2_sat(v[0] - f.x, v[1] - f.y);
  }
}

// mag[b][t][f] = sqrt(re^2 + im^2 + eps) (spectrogram.py:54-55), re at column f, im at column F + f of `spec`
__global__ void spec_mag_kernel(const float* __restrict__ spec, __half* __restrict__ out16, float* __restrict__ out32,
                                int F, int spec_pitch, int pitch, int split, int out32_pitch, float eps,
                                long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % pitch);
  const long long row = i / pitch;
  float m = 0.f;
  if (f < F) {
    const float re = spec[(size_t)row * spec_pitch + f];
    const float im = spec[(size_t)row * spec_pitch + F + f];
    m = sqrtf(fmaf(re, re, fmaf(im, im, eps)));
  }
  if (out16) store_half_split(out16 + (size_t)row * (pitch + split) + f, m, split);
  if (out32 && f < out32_pitch) out32[(size_t)row * out32_pitch + f] = m;
}

// out[b][c][t] = log(max(x[b][t][c], floor)) (spectrogram.py:93-94): leaves the channels-last layout
__global__ void log_mel_out_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int T, int pitch,
                                   float floor_v) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // read: c fastest
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && c < C) ? logf(fmaxf(x[((size_t)b * T + t) * pitch + c], floor_v)) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {  // write: t fastest
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) out[((size_t)b * C + c) * T + t] = tile[threadIdx.x][i];
  }
}

static inline int grid1d_ll(long long total, int threads) { return (int)((total + threads - 1) / threads); }

}  // namespace fv

using namespace fv;

extern "C" int fv_frame_audio(const float* y, void* out16, int B, int L, int hop, int pad_left, int pad_right, int R,
                              int pitch, int split, void* stream) {
  FV_REQUIRE(y && out16 && B > 0 && L > 1 && hop > 0 && R > 0 && pitch >= hop && pitch % 2 == 0 && pad_left >= 0 &&
                 pad_right >= 0 && pad_left < L && pad_right < L && (split == 0 || split == pitch),
             FV_E_BADARG, "fv_frame_audio: bad arguments (L=%d hop=%d pads=%d/%d pitch=%d split=%d)", L, hop, pad_left,
             pad_right, pitch, split);
  const long long total = (long long)B * R * (pitch / 2);
  frame_audio_kernel<<<grid1d_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(y, (__half*)out16, L, hop, pad_left,
                                                                             L + pad_left + pad_right, R, pitch, split,
                                                                             total);
  FV_CHECK_LAUNCH("frame_audio_kernel");
  return 0;
}

extern "C" int fv_spec_mag(const float* spec, void* out16, float* out32, int B, int T, int F, int spec_pitch, int pitch,
                           int split, int out32_pitch, float eps, void* stream) {
  FV_REQUIRE(spec && (out16 || out32) && B > 0 && T > 0 && F > 0 && spec_pitch >= 2 * F && pitch >= F &&
                 (split == 0 || split == pitch) && (!out32 || out32_pitch >= F),
             FV_E_BADARG, "fv_spec_mag: bad arguments");
  const long long total = (long long)B * T * pitch;
  spec_mag_kernel<<<grid1d_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(spec, (__half*)out16, out32, F, spec_pitch,
                                                                          pitch, split, out32 ? out32_pitch : 0, eps,
                                                                          total);
  FV_CHECK_LAUNCH("spec_mag_kernel");
  return 0;
}

extern "C" int fv_log_mel_out(const float* x32, float* out, int B, int C, int T, int pitch, float floor_v,
                              void* stream) {
  FV_REQUIRE(x32 && out && B > 0 && C > 0 && T > 0 && pitch >= C && floor_v > 0.f, FV_E_BADARG,
             "fv_log_mel_out: bad arguments");
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B), block(32, 8);
  log_mel_out_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x32, out, C, T, pitch, floor_v);
  FV_CHECK_LAUNCH("log_mel_out_kernel");
  return 0;
}
