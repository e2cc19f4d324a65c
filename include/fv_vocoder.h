/*
 * fv_vocoder.h - C ABI of libfv_b200.so: B200 (sm_100a) kernels for the fish-vocoder generator forward.
 *
 * The reference (fishaudio/vocoder) has NO native/FFI layer of its own: its boundary is the
 * nn.Module contract `Generator.forward(mel[, template]) -> wav`.  This header is the
 * thin C boundary underneath our Python mirror of that contract; every entry point names the
 * reference call it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *  - All data pointers are CUDA device pointers, caller-owned, never retained past the call.
 *  - `stream` is a cudaStream_t passed as void*; every call is stream-ordered, no hidden sync.
 *  - Activations are channels-last inside the path:  [B][L][pitch]  (pitch = channel count rounded
 *    up to a multiple of 8; padded channels hold zeros).  "a16" tensors are IEEE fp16 tensor-core
 *    operands, "x32" tensors are fp32 (residual stream / accumulators).
 *  - Return value: 0 ok; <0 argument/shape error (FV_E_*); >0 a cudaError_t.  fv_last_error() returns a
 *    thread-local human readable message for the last non-zero return.
 *  - Only mel-in and wav-out are channels-first, as in the reference ([B, n_mels, T] / [B, 1, T*hop]).
 */
#ifndef FV_VOCODER_H_
#define FV_VOCODER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FV_ABI_VERSION 3

#if defined(__GNUC__)
#define FV_API __attribute__((visibility("default")))
#else
#define FV_API
#endif

/* error codes (negative) */
#define FV_E_BADARG (-1)
#define FV_E_ALIGN (-2)
#define FV_E_UNSUPPORTED (-3)
#define FV_E_DRIVER (-4)

/* fused epilogue activations */
enum fv_act {
  FV_ACT_NONE = 0,  /* y = o                                                                      */
  FV_ACT_SILU = 1,  /* F.silu: hifigan.py:103,105,230,245                                          */
  FV_ACT_LEAKY = 2, /* F.leaky_relu(o, act_param): refinegan.py:89,91,302,313,319                  */
  FV_ACT_GELU = 3,  /* nn.GELU() exact erf: encoders/convnext.py:115                               */
  FV_ACT_TANH = 4,  /* torch.tanh: hifigan.py:247                                                  */
  FV_ACT_POLAR = 5, /* column pairs (2k,2k+1) = (log-mag, phase) -> (min(exp(m),100)cos p, ..sin p):
                       generators/vocos.py:57-67                                                  */
  FV_ACT_SILU_TANH = 6, /* F.silu as x/2 + x/2 tanh(x/2) with tanh.approx (one SFU op, |err| <= 2.4e-4 |x|):
                          inner activation of fv_mrf_fused and the SiLU epilogues of fv_conv1d     */
  FV_ACT_SILU_H2 = 7    /* the same formula on packed fp16 pairs (cvt.f16x2, tanh.approx.f16x2, fma.f16x2: one SFU op per
                          TWO channels, the operand word carries ~3 fp16 roundings instead of 1): opt-in inner activation
                          of fv_mrf_fused only; as an out_act / in fv_conv1d it is evaluated like FV_ACT_SILU_TANH */
};

/* how the two anti-alias filters of fv_snake_aa pad their inputs (the BigVGAN-flavoured Activation1d
 * replicates the edge sample; other releases of alias_free_torch may reflect or zero-pad - unverifiable offline) */
enum fv_edge_mode {
  FV_EDGE_REPLICATE = 0, /* F.pad(mode="replicate"): default, streaming kernel */
  FV_EDGE_REFLECT = 1,   /* F.pad(mode="reflect"): mirror without repeating the edge sample */
  FV_EDGE_ZERO = 2       /* F.pad(mode="constant", value=0) */
};

/* which kernel family executes fv_conv1d */
enum fv_engine {
  FV_ENGINE_TC = 0,  /* tcgen05.mma + TMA implicit GEMM (product path)                             */
  FV_ENGINE_SIMT = 1 /* plain CUDA-core kernel with identical semantics (bring-up / cross-check)   */
};

#define FV_MAX_TAPS 64 /* n_phase * n_taps <= FV_MAX_TAPS */

/*
 * One (dilated | transposed-polyphase | pointwise) convolution = one implicit GEMM with a fused epilogue.
 *
 *   acc[b, q, o] = sum_{tap} sum_{c} a[b, q + tap_off[phase][tap], c] * w[phase][tap][o][c]      (fp16 x fp16 -> fp32)
 *   row  = q * n_phase + phase                     (n_phase == 1: plain conv;  == stride u: ConvTranspose1d)
 *   v    = acc + bias[o];  if gamma: v *= gamma[o];  if residual: v += residual[b, row, o]
 *   o32  = v * out_scale (+ out32[b,row,o] if accumulate)          -> out32 (fp32), optional
 *   o16  = fp16(act(o32))                                           -> out16 (fp16), optional
 *   rows of `a` outside [0, L_in) read as zero ("same" zero padding of the reference convs).
 *
 * Replaces: nn.Conv1d in ResBlock1/AMPBlock (hifigan.py:29-98, bigvgan.py:149-218), conv_pre/conv_post
 * (hifigan.py:158-166), nn.ConvTranspose1d ups (hifigan.py:177-187, polyphase form SURVEY B3),
 * nn.Linear pwconv1/2 and 1x1 convs (encoders/convnext.py:112-116,172-177), ISTFTHead.out (vocos.py:41,55)
 * and the inverse-DFT basis GEMM that stands in for torch.fft.irfft.
 */
typedef struct fv_conv_desc {
  /* A operand: fp16 [B][L_in][a_pitch] */
  const void* a;
  int32_t B, L_in, a_pitch;
  /* weights: fp16 [n_phase][n_taps][C_out_pad][w_pitch]; rows >= C_out and cols >= C_in are zero */
  const void* w;
  int32_t n_phase, n_taps, C_out, C_out_pad, w_pitch;
  const int32_t* tap_off; /* HOST pointer, n_phase * n_taps input-row offsets */
  /* output rows per batch element (L_out = L_in for a conv, L_in*u (+1) for ConvTranspose1d) */
  int32_t L_out;
  /* epilogue */
  const float* bias;     /* [C_out] or NULL */
  const float* gamma;    /* [C_out] or NULL: ConvNeXt layer scale (convnext.py:137-138) */
  const float* residual; /* fp32 [B][L_out][res_pitch] or NULL */
  int32_t res_pitch;
  float* out32; /* fp32 [B][L_out][out32_pitch] or NULL */
  int32_t out32_pitch;
  int32_t accumulate; /* 1: out32 += ... (MRF mean over resblocks, hifigan.py:132-133) */
  float out_scale;
  void* out16; /* fp16 [B][L_out][out16_pitch] or NULL */
  int32_t out16_pitch;
  int32_t act; /* enum fv_act applied to the value written to out16 */
  float act_param;
  /* strict ("f16x3") precision: operands carry a second fp16 word so that hi + lo is fp32-grade (22-bit mantissa).
   * a_split = P > 0: `a` is [B][L_in][2P] = [hi | lo]; `w` rows are [Whi | Whi | Wlo] (w_pitch = 3P) and the K loop
   *   reads A columns {0..2P, 0..P}: acc = hi*Whi + lo*Whi + hi*Wlo  (the lo*Wlo term, ~2^-22 relative, is dropped).
   * out16_split = Po > 0: out16 is [B][L_out][2 Po]; the epilogue writes hi at [o] and lo = fp16(v - hi) at [Po + o]. */
  int32_t a_split;
  int32_t out16_split;
} fv_conv_desc;

FV_API const char* fv_last_error(void);
FV_API int fv_abi_version(void);
/* number of kernels launched by this library since load / since the last reset (bench.py "gpu_launches") */
FV_API int64_t fv_launch_count(void);
FV_API void fv_reset_launch_count(void);

FV_API int fv_conv1d(const fv_conv_desc* d, int engine, void* stream);
/* tuning overrides for the tcgen05 engine (0 = built-in heuristic): N tile in {16,32,64,128,256}, 128-row
 * accumulators per CTA in {1,2}, epilogue flavour (1 = LSU + smem transpose, 2 = TMA bulk load/store), mainloop
 * (1 = one smem stage per tap, 2 = one operand slab per K chunk with row-shifted UMMA descriptors per tap).
 * Process-global; meant for benchmarking sweeps. */
FV_API void fv_set_tc_tuning(int block_n, int m_sub, int epilogue, int mainloop);

/* mel [B][C][T] fp32 channels-first -> fp16 channels-last [B][T][pitch] (zero padded channels).
 * Entry of the path: the tensor handed to Generator.forward (hifigan.py:226, convnext.py:206).
 * Every fp16-producing entry point takes `split`: 0 = plain fp16; P > 0 = strict mode, the row holds 2P halfs, hi at [c]
 * and lo = fp16(v - hi) at [P + c] (see fv_conv_desc.a_split). */
FV_API int fv_pack_input(const float* x, void* out16, int B, int C, int T, int pitch, int split, void* stream);

/* fp32 channels-last [B][L][pitch] -> fp32 channels-first [B][C][L]  (leaving the path; debugging) */
FV_API int fv_unpack_output(const float* x32, float* out, int B, int C, int L, int pitch, void* stream);

/* conv_post + tanh (hifigan.py:214-222,246-247): a16 [B][L][pitch] (already activated) * w32 [k][C] + bias
 * -> wav fp32 [B][L] (== [B,1,L]).  C_out == 1, so this is a CUDA-core dot product with a warp-shuffle
 * reduction along the channel axis.  split = pitch in strict mode (rows are [hi | lo], the operand is hi + lo). */
FV_API int fv_conv_post_tanh(const void* a16, const float* w32, const float* bias, float* wav, int B, int L, int C,
                      int pitch, int k, int apply_tanh, int split, void* stream);

/* Anti-aliased Snake / SnakeBeta = alias_free_torch.Activation1d(SnakeBeta) (bigvgan.py:226-233,335-337;
 * SURVEY B4): 2x Kaiser-sinc up (12 taps, replicate edges) -> x + sin^2(a x)/(b+1e-9) -> 2x down, ONE kernel.
 * x32 [B][L][pitch] fp32 -> out16 [B][L][pitch] fp16.  alpha/beta are the raw (log-scale) parameters [C];
 * beta == NULL selects Snake (beta := alpha).  filt_up / filt_down = HOST pointers to the 12 fp32 taps (the
 * module's `upsample.filter` / `downsample.lowpass.filter` buffers; passed by value to the kernel).
 * edge_mode: enum fv_edge_mode, applied to the padding of both filters. */
FV_API int fv_snake_aa(const float* x32, void* out16, const float* alpha, const float* beta, const float* filt_up,
                const float* filt_down, int logscale, int B, int L, int C, int pitch, int split, int edge_mode,
                void* stream);

/* ConvNeXt block front half (convnext.py:127-129): depthwise conv k (zero pad) + LayerNorm over C (eps) -> fp16.
 * x32 [B][T][pitch] -> out16 [B][T][pitch].  dw_w [k][C] (tap-major, i.e. the module's [C,1,k] weight transposed so
 * that channel-parallel threads read it coalesced), dw_b [C], ln_w/ln_b [C]. k <= 0 means "no conv"
 * (plain LayerNorm over C: convnext.py:64-74).  out32 (optional) receives the fp32 result as well. */
FV_API int fv_dwconv_layernorm(const float* x32, void* out16, float* out32, const float* dw_w, const float* dw_b,
                        const float* ln_w, const float* ln_b, float eps, int B, int T, int C, int pitch, int k,
                        int split, void* stream);

/* ISTFT tail (vocos==0.0.2 spectral_ops.ISTFT): windowed frames [B][T][n_fft] fp32 (window already
 * folded into the inverse-DFT basis; `window` [n_fft] is only used for the squared-window envelope) -> overlap-add,
 * trim, divide by the envelope.
 *   center == 0: padding="same"   (generators/vocos.py:33-38): trim (win-hop)/2 either side, wav [B][T*hop]
 *   center != 0: padding="center" (torch.istft(center=True), the upstream-Vocos head of scripts/vocos_gen.py:5-16):
 *                trim n_fft/2 either side, wav [B][(T-1)*hop] */
FV_API int fv_istft_ola(const float* frames, const float* window, float* wav, int B, int T, int n_fft, int hop,
                 int frame_pitch, int center, void* stream);

/* template path (hifigan.py:191-204,233-234): noise_convs[i](template) as fp32 [B][L_i][pitch]; C_in == 1.
 * template [B][L_audio], w [C][k], bias [C]; out row t = sum_j w[c][j] * template[t*stride - pad + j]. */
FV_API int fv_noise_conv(const float* tpl, const float* w, const float* bias, float* out32, int B, int L_audio, int L_out,
                  int C, int pitch, int k, int stride, int pad, void* stream);

/* RefineGAN helpers (refinegan.py:124-127,222,252,302-313): elementwise / linear-resample glue, channels-last.
 * fv_act_cast:  v = act(x32 [+ noise * noise_w[c]]);   (noise != NULL: AdaIN, refinegan.py:124-127)
 *               out32[.., c]              = v * out_scale (+ out32 if accumulate)        fp32 [B][L][out32_pitch]
 *               out16[.., out16_coff + c] = fp16(act16(v))    (channel slice of a wider buffer = torch.cat for free)
 * fv_resample_linear: F.interpolate(mode="linear", align_corners=False) along L of pre_act(x32), then act;
 *               scale = 1 / scale_factor (what torch uses when scale_factor is given). */
FV_API int fv_act_cast(const float* x32, const float* noise, const float* noise_w, void* out16, float* out32, int act,
                       float act_param, int act16, float act16_param, float out_scale, int accumulate, int B, int L,
                       int C, int in_pitch, int out16_pitch, int out16_coff, int out32_pitch, int split, void* stream);
FV_API int fv_resample_linear(const float* x32, float* out32, void* out16, int pre_act, float pre_param, int act,
                              float act_param, int B, int L_in, int L_out, int C, int in_pitch, int out_pitch,
                              int out_coff, float scale, int split, void* stream);

/* Mel front-end, the step immediately before the path (LinearSpectrogram / LogMelSpectrogram,
 * fish_vocoder/data/transforms/spectrogram.py:6-104; used at test.py:71, models/gan.py:284):
 * fv_frame_audio: y [B][L] fp32 -> out16 [B][R][pitch] fp16, row r = samples r*hop .. r*hop+hop-1 of
 *                 F.pad(y, (pad_left, pad_right), "reflect") (spectrogram.py:29-37); with the signal laid out like this the
 *                 framed, windowed DFT (torch.stft, center=False) is an fv_conv1d with n_fft/hop taps (offsets 0..k-1) whose
 *                 weights are window * {cos, -sin}: output columns [re_0..re_{F-1} | im_0..im_{F-1}], F = n_fft/2 + 1.
 * fv_spec_mag:    sqrt(re^2 + im^2 + eps) (spectrogram.py:54-55) -> fp16 operand of the mel matmul and/or fp32.
 * fv_log_mel_out: log(max(x, floor)) (spectrogram.py:93-94), channels-last [B][T][pitch] -> channels-first [B][C][T]. */
FV_API int fv_frame_audio(const float* y, void* out16, int B, int L, int hop, int pad_left, int pad_right, int R,
                          int pitch, int split, void* stream);
FV_API int fv_spec_mag(const float* spec, void* out16, float* out32, int B, int T, int F, int spec_pitch, int pitch,
                       int split, int out32_pitch, float eps, void* stream);
FV_API int fv_log_mel_out(const float* x32, float* out, int B, int C, int T, int pitch, float floor_v, void* stream);

/*
 * Fused MRF stage: mean over `n_blocks` residual blocks, each a chain of `n_pairs` pairs
 *     xt = act(x); xt = conv_{k, dil1}(xt) + b1; xt = act(xt); xt = conv_{k, dil2}(xt) + b2; x = x + xt
 * evaluated ENTIRELY ON CHIP per 512-row time tile (halo recomputed): the fp32 residual stream lives in TMEM as the
 * accumulator the second conv of every pair adds onto, the fp16 operands of the 2*n_pairs convs live in swizzled shared
 * memory (taps = row-shifted UMMA descriptors), weights stream through a TMA ring.  HBM traffic per stage: x read,
 * result written - instead of 16 bytes per element per conv pair for the layer-wise path.
 *     out32[b, t, c] = (1/n_blocks) * sum_j block_j(x)[b, t, c]        (also the running-sum scratch: required)
 *     out16          = fp16(out_act(out32))                            (optional)
 * Replaces ParralelBlock.forward / ResBlock1.forward (hifigan.py:101-108,117-133): the stack([...]).mean(0) over
 * kernel sizes (3,7,11) of 3 x {silu, conv(k,d), silu, conv(k,1), +x}.  C in {16, 32, 64}: whole stages on 512-row
 * tiles; C = 128: 256-row tiles and at most ONE pair per launch (n_blocks = n_pairs = 1; x -> x + pair(x), chained by the
 * host with `accumulate` / `out_scale` building the mean).  Tap reach (k-1)/2*dil <= 32.
 */
#define FV_MRF_MAX_BLOCKS 4
#define FV_MRF_MAX_PAIRS 4
typedef struct fv_mrf_desc {
  const float* x; /* [B][L][x_pitch] fp32 channels-last stage input */
  int32_t B, L, C, x_pitch;
  const void* w;  /* fp16 [w_rows][C]: the [C_out][C_in] tap tiles of every conv, concatenated */
  int32_t w_rows;
  const float* bias; /* fp32 [n_blocks][n_pairs][2][C]: b1, b2 of every pair */
  int32_t n_blocks, n_pairs;
  int32_t ksize[FV_MRF_MAX_BLOCKS];
  int32_t dil1[FV_MRF_MAX_BLOCKS][FV_MRF_MAX_PAIRS];
  int32_t dil2[FV_MRF_MAX_BLOCKS][FV_MRF_MAX_PAIRS];
  int32_t w_row0[FV_MRF_MAX_BLOCKS][FV_MRF_MAX_PAIRS][2]; /* first row in `w` of tap 0 of (block, pair, conv) */
  int32_t act; /* FV_ACT_SILU | FV_ACT_LEAKY | FV_ACT_SILU_TANH */
  float act_param;
  float* out32;
  int32_t out32_pitch;
  void* out16;
  int32_t out16_pitch;
  int32_t out_act;
  float out_act_param;
  /* pair-wise evaluation of a stage (C = 128: one launch per (conv, conv) pair, the host chains them): */
  int32_t accumulate; /* 1: the result is ADDED to what out32 already holds (out32 += out_scale * block(x)) */
  float out_scale;    /* 0 = 1 / n_blocks (the MRF mean of a whole-stage launch) */
} fv_mrf_desc;
FV_API int fv_mrf_fused(const fv_mrf_desc* d, void* stream);

/* bring-up probe (not on the product path): 12 row shifts x {base_offset 0, base_offset r&7} of a 128x64x64 UMMA whose
 * A descriptor starts r rows into a TMA-written 144x64 fp16 slab.  a16 [144][64], w16 [64][64], out [12][2][128][64]. */
FV_API int fv_debug_rowshift_probe(const void* a16, const void* w16, float* out, void* stream);

/* bring-up probe (not on the product path): SM cycles per tcgen05.mma (M = 128 per CTA, K = 16, fp16) with resident
 * operands.  mode 0 = SS cta_group::1, 1 = SS cta_group::2 (CTA pair, M = 256), 2 = A from TMEM, 3 = SS weight-stationary
 * (tcgen05.mma.ws); n = UMMA N; bg = background
 * shared-memory traffic of 8 other warps (0 none, 1 stores, 2 loads).  out[cta] = cycles per UMMA x 1000. */
FV_API int fv_debug_umma_rate(int mode, int n, int reps, int bg, int* out, void* stream);
/* bring-up probe: fp32 FMA throughput of the CUDA cores (the roofline of fv_snake_aa).  variant 0 = scalar fma, three register
 * operands; 1 = scalar fma with a kernel-constant multiplier; 2 = packed fma.rn.f32x2.  Every thread of num_sms x 8 x 256
 * runs `iters` x 16 dependent-chain FMAs over 8 independent chains; out[1] = elapsed SM cycles of block 0. */
FV_API int fv_debug_fma_rate(int variant, int iters, float* sink, long long* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FV_VOCODER_H_ */
