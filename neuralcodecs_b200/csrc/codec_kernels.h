// Non-GEMM kernels of the codec hot path (HBM-bound): first conv (Cin = 1), layout
// transposes at the API boundary, fused residual vector quantisation, code lookup.
#pragma once
#include <cstdint>
#include <vector>

#include "runtime.h"

namespace nc {

// out[b, t, co] = bias[co] + sum_j w[co, j] * f(in[b, t + j*dil - pad])   (Cin = 1, stride 1)
// in: [B][in_stride] with `in_len` valid samples per clip (reads beyond in_len give 0: this is
// DAC.Preprocess' right zero padding, Models/DAC.cs:151-153); out: [B][t_out][cout] channels-last.
void launch_conv_cin1(const float* in, long long in_stride, int in_len, float* out, int t_out, int cout,
                      const float* w /*[cout][k]*/, const float* bias, int k, int dil, int pad, int batch,
                      const LaunchCtx& ctx, int reflect = 0 /*1: reflect padding instead of zeros*/,
                      long long out_clip_stride = 0 /*floats between output clips; 0 = t_out*cout*/);

// out[b, t] = act(bias + sum_j sum_c w[j][c] * in[b, t + j - pad, c])   (Cout = 1, stride 1, dilation 1; act 1 = tanh)
// in: [B][T][C] channels-last (already activated); out: [B][T]; w_kc: [k][C].
void launch_conv_cout1(const float* in, float* out, int T, int C, const float* w_kc, const float* bias, int k, int pad,
                       int act, int batch, const LaunchCtx& ctx, int reflect = 0, long long in_clip_stride = 0);

// out16[i] = fp16(in[i]) (round to nearest even): the latent handed to the fp16-operand decoder layers
void launch_f32_to_f16(const float* in, void* out16, long long n, const LaunchCtx& ctx);

// [B][C][T] <-> [B][T][C]
void launch_transpose_ct_to_tc(const float* in, float* out, int batch, int C, int T, const LaunchCtx& ctx);
void launch_transpose_tc_to_ct(const float* in, float* out, int batch, int C, int T, const LaunchCtx& ctx);

// Device-resident parameters of a DAC/SNAC-style factorised RVQ
// (Modules/DAC/VectorQuantizer.cs, ResidualVectorQuantizer.cs): per stage in_proj [D][Dz] + bias [D],
// codebook [K][D] + squared norms [K], out_proj [Dz][D] + bias [Dz].
struct RvqWeights {
  int n_stages = 0, Dz = 0, D = 0, K = 0;
  const float* in_w = nullptr;   // [stage][D][Dz]
  const float* in_b = nullptr;   // [stage][D]
  const float* cb = nullptr;     // [stage][K][D]
  const float* cb_sq = nullptr;  // [stage][K]   fp32 sum of squares (d ascending)
  const float* out_w = nullptr;  // [stage][Dz][D]
  const float* out_b = nullptr;  // [stage][Dz]
  // per-stage shared-memory image for the block kernel (rvq_stage_blob): [stage][blob_floats], nullptr = warp kernel only
  const float* blob = nullptr;
  int blob_floats = 0;
};
// One stage's weights in the layout rvq_encode_block_kernel reads from shared memory (all fp32):
//   in_proj  [D][Dz/128][32 lanes][4]   element (d, i4, lane, j) = in_w[d][lane + 32 * (4 * i4 + j)]
//   in_b [D] | codebook dims 0-3 [K][4] | dims 4-7 [K][4] | cb_sq [K] | out_proj dims 0-3 [Dz][4] | dims 4-7 [Dz][4] | out_b [Dz]
// Requires D == 8 and Dz % 128 == 0; returns an empty vector otherwise.
std::vector<float> rvq_stage_blob(int Dz, int D, int K, const float* in_w, const float* in_b, const float* cb, const float* cb_sq,
                                  const float* out_w, const float* out_b);

// Fused DAC RVQ encode over frames z[b, t, :] (channels-last rows of Dz floats):
//   per stage: zE = in_proj(res); idx = argmin_k (|zE|^2 + |c_k|^2) - 2 zE.c_k (lowest index wins);
//   zQ = out_proj(zE + (c_idx - zE)); zq_sum += zQ; res -= zQ
// Outputs: zq [B][T][Dz] (nullable), codes [B][n_q][T] int64 (nullable), latents [B][n_q*D][T] (nullable).
void launch_rvq_encode(const RvqWeights& w, const float* z, float* zq, int64_t* codes, float* latents,
                       int batch, int T, int n_q, const LaunchCtx& ctx);

// ResidualVectorQuantizer.FromCodes (Modules/DAC/ResidualVectorQuantizer.cs:211-238):
// zq[b, t, :] = sum_i out_proj_i(codebook_i[codes[b, i, t]]).  codes: [B][n_q][T] int64.
// Returns false through *bad (device flag, nullable) if a code is out of range (clamped to 0..K-1).
void launch_rvq_from_codes(const RvqWeights& w, const int64_t* codes, float* zq, int batch, int T, int n_q,
                           const LaunchCtx& ctx);

// Dia's delayed codes [B][T][C] -> DAC codes [n_items][C][len] for the listed items (revert delay, clamp to 0).
void launch_dia_revert(const int64_t* gen, const int* items_dev, const int* delay_dev, int64_t* codes, int n_items, int T, int C,
                       int len, int K, const LaunchCtx& ctx);

// Input conditioning (the step before the codec path).
// Linear resampler of SNAC.ResampleAudio (Models/SNAC.cs:284-308) = AudioUtils.ResampleLinear
// (NeuralCodecs.Core/Utils/AudioUtils.cs:329-352): position = i / ratio in double, two-point interpolation in double,
// last sample held.  in [batch][n_in] (stride in_stride) -> out [batch][n_out].  n_out = (long)(n_in * ratio).
void launch_resample_linear(const float* in, long long n_in, long long in_stride, float* out, long long n_out, long long out_stride,
                            double ratio, int batch, const LaunchCtx& ctx);
// AudioUtils.ConvertToMono (AudioUtils.cs:45-62): out[i] = (sum over channels in order, float) / channels
void launch_to_mono(const float* interleaved, float* out, long long frames, int channels, const LaunchCtx& ctx);

}  // namespace nc
