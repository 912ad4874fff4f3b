// SNAC-specific HBM-bound kernels (snac_kernels.cu).
#pragma once
#include <cstdint>

#include "runtime.h"

namespace nc {

// Depthwise Conv1d, kernel 7, stride 1, padding 3*dil (groups = channels), channels-last [B][T][C]:
//   out[b,t,c] = post(bias[c] + sum_j w[j][c] * pro(in[b, t + (j-3)*dil, c]))
// pro / post = Snake with the given alpha vectors (null = none).  w_kc: [7][C].
void launch_dwconv7(const float* in, float* out, int T, int C, const float* w_kc, const float* bias, int dil,
                    const float* pro_alpha, const float* post_alpha, int batch, const LaunchCtx& ctx, const char* layer,
                    bool fast_sin = false /*MUFU sin in the Snakes (as the tensor-core layers do)*/);

// One stage of SNAC's residual VQ (Modules/SNAC/VectorQuantizer.cs:82-103) on channels-last rows.
struct SnacVqStage {
  int stride = 1;
  const float* in_w = nullptr;   // [8][Dz]
  const float* in_b = nullptr;   // [8]
  const float* cb = nullptr;     // [K][8]
  const float* cb_sq = nullptr;  // [K]
  const float* out_w = nullptr;  // [Dz][8]
  const float* out_b = nullptr;  // [Dz]
};
// residual, zq: [B][T][Dz] updated in place (zq += zQ_i, residual -= zQ_i); codes: [B][T/stride] int64 (nullable)
void launch_snac_vq_stage(const SnacVqStage& s, float* residual, float* zq, int64_t* codes, int batch, int T, int Dz, int K,
                          const LaunchCtx& ctx);

constexpr int kSnacMaxStages = 8;
struct SnacFromCodes {
  int n_stages = 0;
  int stride[kSnacMaxStages];
  const int64_t* codes[kSnacMaxStages];  // [B][T/stride]
  const float* cb[kSnacMaxStages];
  const float* out_w[kSnacMaxStages];
  const float* out_b[kSnacMaxStages];
};
// zq[b,t,:] = sum_i out_proj_i(codebook_i[codes_i[b, t / stride_i]])   (ResidualVectorQuantizer.cs:91-131)
void launch_snac_from_codes(const SnacFromCodes& a, float* zq, int batch, int T, int Dz, int K, const LaunchCtx& ctx);

// y[r, :] = LayerNorm(x[r, :]) * gamma + beta over C channels (eps 1e-5), rows of a channels-last tensor
void launch_layernorm_rows(const float* x, float* y, const float* gamma, const float* beta, long long rows, int C,
                           const LaunchCtx& ctx);
// Non-causal attention inside windows of 32 frames with rotary embeddings (Modules/SNAC/LocalMHA.cs:93-110):
// qkv [B][T][3C] -> out [B][T][C]; heads = C / 64.
void launch_local_attn(const float* qkv, float* out, const float* inv_freq, int batch, int T, int C, int heads, int window,
                       const LaunchCtx& ctx);

// out[b, 0:out_len] = in[b*in_stride + 0:out_len]
void launch_trim_rows(const float* in, float* out, int batch, long long in_stride, long long out_len, const LaunchCtx& ctx);
// N(0,1) samples from a counter-based generator keyed by (seed, stream_id, index)
void launch_randn(float* out, int clips, int T, long long clip0, uint64_t seed, uint32_t stream_id, const LaunchCtx& ctx);

}  // namespace nc
