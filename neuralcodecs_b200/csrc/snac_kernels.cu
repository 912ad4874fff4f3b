// HBM-bound kernels specific to SNAC: depthwise k=7 conv (+ Snake prologue / post-activation), strided
// residual VQ stages (avg-pool -> in_proj -> argmin -> lookup -> out_proj -> repeat), codes -> latent.
// Reference: Modules/SNAC/ResidualUnit.cs:25-60 (groups = channels), Modules/SNAC/VectorQuantizer.cs:82-157,
// Modules/SNAC/ResidualVectorQuantizer.cs:69-131.
#include "snac_kernels.h"

#include <cfloat>

namespace nc {

// |sin(t)| to ~1.2 ulp for |t| < 1e5: 3-term Cody-Waite reduction by pi/2 + minimax polynomials (branch-free;
// the sign is dropped because Snake only uses sin^2).  Same routine as the conv kernel's prologue.
__device__ __forceinline__ float sin_abs_cw(float x) {
  const int q = __float2int_rn(x * 0.636619772f);
  const float j = __int2float_rn(q);
  float r = fmaf(j, -1.57079601e+00f, x);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  const float s = r * r;
  const bool odd = (q & 1) != 0;
  float p = odd ? 2.44331571e-5f : -1.95152959e-4f;
  p = fmaf(p, s, odd ? -1.38873163e-3f : 8.33216087e-3f);
  p = fmaf(p, s, odd ? 4.16666457e-2f : -1.66666546e-1f);
  const float a = odd ? fmaf(p, s, -0.5f) : p;
  const float m = odd ? s : r * s;
  const float b = odd ? 1.0f : r;
  return fmaf(a, m, b);
}
// x + sin^2(a x) / a with ia = 1/a (0 where a == 0 -> identity): where(alpha == 0, x, addcdiv(x, sin(ax)^2, alpha))
template <bool kFast>
__device__ __forceinline__ float snake_dev(float x, float a, float ia) {
  const float t = a * x;
  float s;
  if (kFast) {
    s = __sinf(t);           // MUFU: abs error ~4e-7, below the tensor-core layers' own noise
  } else {
    s = sin_abs_cw(t);
    if (fabsf(t) > 9.0e4f) s = sinf(t);
  }
  return fmaf(s * s, ia, x);
}
__device__ __forceinline__ float4 inv4(float4 a) {
  return make_float4(a.x == 0.f ? 0.f : 1.0f / a.x, a.y == 0.f ? 0.f : 1.0f / a.y, a.z == 0.f ? 0.f : 1.0f / a.z,
                     a.w == 0.f ? 0.f : 1.0f / a.w);
}

// ------------------------------------------------------------------------------ depthwise conv, k = 7
constexpr int kDwTT = 128;  // output time steps per block
constexpr int kDwCC = 64;   // channels per block

template <bool kFast>
__global__ void __launch_bounds__(256)
dwconv7_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int C, const float* __restrict__ w_kc,
               const float* __restrict__ bias, int dil, const float* __restrict__ pro_alpha,
               const float* __restrict__ post_alpha, int tiles_per_clip) {
  extern __shared__ float tile[];  // [(kDwTT + 6*dil)][kDwCC]
  const int b = blockIdx.x / tiles_per_clip;
  const int t0 = (blockIdx.x - b * tiles_per_clip) * kDwTT;
  const int c0 = blockIdx.y * kDwCC;
  const int c4 = threadIdx.x & 15, rl = threadIdx.x >> 4;   // 16 float4 per row, 16 rows per pass
  const int c = c0 + 4 * c4;
  const bool c_ok = c < C;
  const float* x = in + (long long)b * T * C;
  float* y = out + (long long)b * T * C;
  const int rows = kDwTT + 6 * dil;
  float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pro_alpha && c_ok) pa = __ldg(reinterpret_cast<const float4*>(pro_alpha + c));
  const float4 pia = inv4(pa);
  for (int r = rl; r < rows; r += 16) {
    const int t = t0 - 3 * dil + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c_ok && t >= 0 && t < T) {
      v = __ldg(reinterpret_cast<const float4*>(x + (long long)t * C + c));
      if (pro_alpha) {
        v.x = snake_dev<kFast>(v.x, pa.x, pia.x); v.y = snake_dev<kFast>(v.y, pa.y, pia.y);
        v.z = snake_dev<kFast>(v.z, pa.z, pia.z); v.w = snake_dev<kFast>(v.w, pa.w, pia.w);
      }
    }
    *reinterpret_cast<float4*>(tile + r * kDwCC + 4 * c4) = v;
  }
  __syncthreads();
  if (!c_ok) return;
  float4 wj[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) wj[j] = __ldg(reinterpret_cast<const float4*>(w_kc + (long long)j * C + c));
  const float4 bb = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 qa = make_float4(0.f, 0.f, 0.f, 0.f);
  if (post_alpha) qa = __ldg(reinterpret_cast<const float4*>(post_alpha + c));
  const float4 qia = inv4(qa);
  for (int r = rl; r < kDwTT; r += 16) {
    const int t = t0 + r;
    if (t >= T) break;
    // conv1d accumulates the taps in order j = 0..6 and adds the bias last (ATen: output = conv + bias)
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(tile + (r + j * dil) * kDwCC + 4 * c4);
      a.x = fmaf(wj[j].x, v.x, a.x); a.y = fmaf(wj[j].y, v.y, a.y);
      a.z = fmaf(wj[j].z, v.z, a.z); a.w = fmaf(wj[j].w, v.w, a.w);
    }
    a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
    if (post_alpha) {
      a.x = snake_dev<kFast>(a.x, qa.x, qia.x); a.y = snake_dev<kFast>(a.y, qa.y, qia.y);
      a.z = snake_dev<kFast>(a.z, qa.z, qia.z); a.w = snake_dev<kFast>(a.w, qa.w, qia.w);
    }
    *reinterpret_cast<float4*>(y + (long long)t * C + c) = a;
  }
}

void launch_dwconv7(const float* in, float* out, int T, int C, const float* w_kc, const float* bias, int dil,
                    const float* pro_alpha, const float* post_alpha, int batch, const LaunchCtx& ctx, const char* layer,
                    bool fast_sin) {
  if (C % 4 != 0) throw Error(NC_UNSUPPORTED, "dwconv7: channel count must be a multiple of 4");
  if ((long long)batch * T == 0) return;
  const int tiles = (T + kDwTT - 1) / kDwTT;
  const size_t smem = (size_t)(kDwTT + 6 * dil) * kDwCC * sizeof(float);
  auto kern = fast_sin ? dwconv7_kernel<true> : dwconv7_kernel<false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  dim3 grid((unsigned)(batch * tiles), (unsigned)((C + kDwCC - 1) / kDwCC));
  const int ev = ctx.begin();
  kern<<<grid, 256, smem, ctx.stream>>>(in, out, T, C, w_kc, bias, dil, pro_alpha, post_alpha, tiles);
  check_launch((int)cudaGetLastError(), "dwconv7");
  ctx.end(ev, "dwconv7", 2.0 * 7 * C * (double)T * batch, 8.0 * batch * (double)T * C, layer ? layer : "");
}

// ------------------------------------------------------------------------------ strided VQ stage
// One warp per pooled frame (stride s input frames): residual rows are channels-last [B][T][Dz].
//   ze = in_proj(mean_s(residual)); idx = argmin_k (|ze|^2 + |c_k|^2) - 2 ze.c_k; zq = out_proj(ze + (c_idx - ze));
//   for each of the s frames: zq_sum += zq; residual -= zq.
template <int CPL>
__global__ void __launch_bounds__(256)
snac_vq_stage_kernel(const float* __restrict__ in_w, const float* __restrict__ in_b, const float* __restrict__ cb,
                     const float* __restrict__ cb_sq, const float* __restrict__ out_w, const float* __restrict__ out_b,
                     float* __restrict__ residual, float* __restrict__ zq, int64_t* __restrict__ codes, int batch, int T,
                     int stride, int Dz, int K) {
  constexpr int D = 8;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int Tp = T / stride;
  const long long frames = (long long)batch * Tp;
  for (long long f = warp0; f < frames; f += nwarps) {
    const int b = (int)(f / Tp), tp = (int)(f % Tp);
    float* rbase = residual + ((long long)b * T + (long long)tp * stride) * Dz;
    float* qbase = zq + ((long long)b * T + (long long)tp * stride) * Dz;
    float r[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      float s = rbase[lane + 32 * i];
      for (int u = 1; u < stride; ++u) s += rbase[(long long)u * Dz + lane + 32 * i];   // avg_pool1d: sum / s
      r[i] = stride > 1 ? s / (float)stride : s;
    }
    float ze[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) p = fmaf(__ldg(in_w + (size_t)d * Dz + lane + 32 * i), r[i], p);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      ze[d] = p + __ldg(in_b + d);
    }
    float e2 = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) e2 = fmaf(ze[d], ze[d], e2);
    float best = FLT_MAX;
    int best_k = 0;
    for (int k = lane; k < K; k += 32) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * D));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * D) + 1);
      float dot = ze[0] * c0.x;
      dot = fmaf(ze[1], c0.y, dot); dot = fmaf(ze[2], c0.z, dot); dot = fmaf(ze[3], c0.w, dot);
      dot = fmaf(ze[4], c1.x, dot); dot = fmaf(ze[5], c1.y, dot); dot = fmaf(ze[6], c1.z, dot);
      dot = fmaf(ze[7], c1.w, dot);
      const float dist = (e2 + __ldg(cb_sq + k)) - 2.0f * dot;
      if (dist < best) { best = dist; best_k = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (od < best || (od == best && ok < best_k)) { best = od; best_k = ok; }
    }
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)best_k * D));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)best_k * D) + 1);
    float q[D] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = ze[d] + (q[d] - ze[d]);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(out_w + (size_t)c * D));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(out_w + (size_t)c * D) + 1);
      float v = w0.x * q[0];
      v = fmaf(w0.y, q[1], v); v = fmaf(w0.z, q[2], v); v = fmaf(w0.w, q[3], v);
      v = fmaf(w1.x, q[4], v); v = fmaf(w1.y, q[5], v); v = fmaf(w1.z, q[6], v); v = fmaf(w1.w, q[7], v);
      v += __ldg(out_b + c);
      for (int u = 0; u < stride; ++u) {   // repeat_interleave(stride)
        qbase[(long long)u * Dz + c] += v;
        rbase[(long long)u * Dz + c] -= v;
      }
    }
    if (codes && lane == 0) codes[(long long)b * Tp + tp] = best_k;
  }
}

void launch_snac_vq_stage(const SnacVqStage& s, float* residual, float* zq, int64_t* codes, int batch, int T, int Dz, int K,
                          const LaunchCtx& ctx) {
  if (Dz % 32 != 0 || K % 32 != 0) throw Error(NC_UNSUPPORTED, "snac vq: latent dim / codebook size must be multiples of 32");
  if (T % s.stride != 0) throw Error(NC_INVALID_ARGUMENT, "snac vq: frames not a multiple of the stage stride");
  const long long frames = (long long)batch * (T / s.stride);
  if (frames == 0) return;
  long long blocks = (frames + 7) / 8;
  const long long cap = (long long)ctx.num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int ev = ctx.begin();
#define NC_VQ_CASE(n)                                                                                                   \
  case n:                                                                                                               \
    snac_vq_stage_kernel<n><<<(unsigned)blocks, 256, 0, ctx.stream>>>(s.in_w, s.in_b, s.cb, s.cb_sq, s.out_w, s.out_b,   \
                                                                      residual, zq, codes, batch, T, s.stride, Dz, K);  \
    break;
  switch (Dz / 32) {
    NC_VQ_CASE(1) NC_VQ_CASE(2) NC_VQ_CASE(3) NC_VQ_CASE(4) NC_VQ_CASE(6) NC_VQ_CASE(8) NC_VQ_CASE(12) NC_VQ_CASE(16) NC_VQ_CASE(24) NC_VQ_CASE(32)
    default: throw Error(NC_UNSUPPORTED, "snac vq: unsupported latent dim " + std::to_string(Dz));
  }
#undef NC_VQ_CASE
  check_launch((int)cudaGetLastError(), "snac_vq_stage");
  ctx.end(ev, "snac_vq_stage", (double)frames * 2.0 * (2.0 * 8 * Dz + 8.0 * K), (double)batch * T * Dz * 4.0 * 4);
}

// ------------------------------------------------------------------------------ codes -> latent
template <int CPL>
__global__ void __launch_bounds__(256)
snac_from_codes_kernel(SnacFromCodes a, float* __restrict__ zq, int batch, int T, int Dz, int K) {
  constexpr int D = 8;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long frames = (long long)batch * T;
  for (long long f = warp0; f < frames; f += nwarps) {
    const int b = (int)(f / T), t = (int)(f % T);
    float acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
    for (int s = 0; s < a.n_stages; ++s) {
      const int st = a.stride[s];
      long long code = a.codes[s][(long long)b * (T / st) + t / st];
      code = code < 0 ? 0 : (code >= K ? K - 1 : code);
      const float* e = a.cb[s] + (size_t)code * D;
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(e));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(e) + 1);
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.out_w[s] + (size_t)c * D));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.out_w[s] + (size_t)c * D) + 1);
        float v = w0.x * q0.x;
        v = fmaf(w0.y, q0.y, v); v = fmaf(w0.z, q0.z, v); v = fmaf(w0.w, q0.w, v);
        v = fmaf(w1.x, q1.x, v); v = fmaf(w1.y, q1.y, v); v = fmaf(w1.z, q1.z, v); v = fmaf(w1.w, q1.w, v);
        v += __ldg(a.out_b[s] + c);
        acc[i] = s == 0 ? v : acc[i] + v;   // first stage assigns, later stages add (ResidualVectorQuantizer.cs:118-126)
      }
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) zq[f * Dz + lane + 32 * i] = acc[i];
  }
}

void launch_snac_from_codes(const SnacFromCodes& a, float* zq, int batch, int T, int Dz, int K, const LaunchCtx& ctx) {
  if (Dz % 32 != 0) throw Error(NC_UNSUPPORTED, "snac from_codes: latent dim must be a multiple of 32");
  const long long frames = (long long)batch * T;
  if (frames == 0) return;
  long long blocks = (frames + 7) / 8;
  const long long cap = (long long)ctx.num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int ev = ctx.begin();
#define NC_FC_CASE(n)                                                                                        \
  case n:                                                                                                    \
    snac_from_codes_kernel<n><<<(unsigned)blocks, 256, 0, ctx.stream>>>(a, zq, batch, T, Dz, K);             \
    break;
  switch (Dz / 32) {
    NC_FC_CASE(1) NC_FC_CASE(2) NC_FC_CASE(3) NC_FC_CASE(4) NC_FC_CASE(6) NC_FC_CASE(8) NC_FC_CASE(12) NC_FC_CASE(16) NC_FC_CASE(24) NC_FC_CASE(32)
    default: throw Error(NC_UNSUPPORTED, "snac from_codes: unsupported latent dim " + std::to_string(Dz));
  }
#undef NC_FC_CASE
  check_launch((int)cudaGetLastError(), "snac_from_codes");
  ctx.end(ev, "snac_from_codes", (double)frames * a.n_stages * 2.0 * 8 * Dz, (double)frames * Dz * 4.0);
}

// ------------------------------------------------------------------------------ LocalMHA pieces
// LayerNorm over channels of a channels-last row (Modules/SNAC/LocalMHA.cs:88: nn.LayerNorm(dim), eps 1e-5).
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, long long rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)C + eps);
    for (int c = lane; c < C; c += 32) y[r * C + c] = (xr[c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
  }
}

void launch_layernorm_rows(const float* x, float* y, const float* gamma, const float* beta, long long rows, int C,
                           const LaunchCtx& ctx) {
  if (rows == 0) return;
  const int blocks = (int)std::min<long long>((rows + 7) / 8, (long long)ctx.num_sms * 8);
  const int ev = ctx.begin();
  layernorm_rows_kernel<<<blocks, 256, 0, ctx.stream>>>(x, y, gamma, beta, rows, C, 1e-5f);
  check_launch((int)cudaGetLastError(), "layernorm_rows");
  ctx.end(ev, "layernorm_rows", 0, 8.0 * rows * C);
}

// Windowed multi-head attention with rotary position embedding, window = 32, head dim = 64
// (LocalMHA.cs:93-110, SinusoidalEmbedding.cs:60-72, RotaryEmbedding.cs:33-52; xpos scale = 1).
// qkv: [B][T][3C] (q | k | v); out: [B][T][C].  One warp per (clip, window, head); lane = query position.
constexpr int kAttnW = 32, kAttnD = 64;
__global__ void __launch_bounds__(64)
local_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, const float* __restrict__ inv_freq, int batch,
                  int T, int C, int heads) {
  __shared__ float ks[2][kAttnW][kAttnD];
  __shared__ float vs[2][kAttnW][kAttnD];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int windows = T / kAttnW;
  const long long item = (long long)blockIdx.x * 2 + wib;
  const long long total = (long long)batch * windows * heads;
  if (item >= total) return;
  const int h = (int)(item % heads);
  const int w = (int)((item / heads) % windows);
  const int b = (int)(item / ((long long)heads * windows));
  const float* row = qkv + ((long long)b * T + (long long)w * kAttnW + lane) * 3 * C + h * kAttnD;
  float q[kAttnD], kk[kAttnD];
#pragma unroll
  for (int d4 = 0; d4 < kAttnD / 4; ++d4) {
    const float4 a = *reinterpret_cast<const float4*>(row + 4 * d4);
    const float4 c = *reinterpret_cast<const float4*>(row + C + 4 * d4);
    const float4 v = *reinterpret_cast<const float4*>(row + 2 * C + 4 * d4);
    q[4 * d4] = a.x; q[4 * d4 + 1] = a.y; q[4 * d4 + 2] = a.z; q[4 * d4 + 3] = a.w;
    kk[4 * d4] = c.x; kk[4 * d4 + 1] = c.y; kk[4 * d4 + 2] = c.z; kk[4 * d4 + 3] = c.w;
    *reinterpret_cast<float4*>(&vs[wib][lane][4 * d4]) = v;
  }
  // rotary: x * cos(f) + rotate_half(x) * sin(f), f[d] = pos * inv_freq[d % 32], rotate_half = cat(-x2, x1)
#pragma unroll
  for (int d = 0; d < kAttnD / 2; ++d) {
    const float f = (float)lane * __ldg(inv_freq + d);
    const float cs = cosf(f), sn = sinf(f);
    const float q1 = q[d], q2 = q[d + 32], k1 = kk[d], k2 = kk[d + 32];
    q[d] = q1 * cs + (-q2) * sn;
    q[d + 32] = q2 * cs + q1 * sn;
    kk[d] = k1 * cs + (-k2) * sn;
    kk[d + 32] = k2 * cs + k1 * sn;
  }
#pragma unroll
  for (int d = 0; d < kAttnD; ++d) ks[wib][lane][d] = kk[d];
  __syncwarp();
  float sc[kAttnW];
  float mx = -3.4e38f;
#pragma unroll
  for (int j = 0; j < kAttnW; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < kAttnD; ++d) s = fmaf(q[d], ks[wib][j][d], s);
    s *= 0.125f;   // 1 / sqrt(64)
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kAttnW; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.0f / sum;
  float* orow = out + ((long long)b * T + (long long)w * kAttnW + lane) * C + h * kAttnD;
#pragma unroll
  for (int d4 = 0; d4 < kAttnD / 4; ++d4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kAttnW; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(&vs[wib][j][4 * d4]);
      const float p = sc[j] * inv;
      a.x = fmaf(p, v.x, a.x); a.y = fmaf(p, v.y, a.y); a.z = fmaf(p, v.z, a.z); a.w = fmaf(p, v.w, a.w);
    }
    *reinterpret_cast<float4*>(orow + 4 * d4) = a;
  }
}

// Any other window size (Modules/SNAC/LocalMHA.cs:46-70 takes any windowSize; the presets use 32): one CTA per
// (clip, window, head), one thread per query row, K / V and the score rows in shared memory.  Same arithmetic order as the
// 32-wide kernel (sequential dot products over d, max, exp, sum, weighted sum over j).
__global__ void local_attn_generic_kernel(const float* __restrict__ qkv, float* __restrict__ out, const float* __restrict__ inv_freq,
                                          int T, int C, int heads, int W) {
  extern __shared__ float attn_smem[];
  float* ks = attn_smem;                       // [W][64]
  float* vs = ks + (size_t)W * kAttnD;         // [W][64]
  float* sc = vs + (size_t)W * kAttnD;         // [W][W + 1]
  const int i = threadIdx.x;
  const int windows = T / W;
  const long long item = blockIdx.x;
  const int h = (int)(item % heads);
  const int w = (int)((item / heads) % windows);
  const int b = (int)(item / ((long long)heads * windows));
  float q[kAttnD];
  if (i < W) {
    const float* row = qkv + ((long long)b * T + (long long)w * W + i) * 3 * C + h * kAttnD;
    float kk[kAttnD];
#pragma unroll
    for (int d = 0; d < kAttnD; ++d) { q[d] = row[d]; kk[d] = row[C + d]; vs[(size_t)i * kAttnD + d] = row[2 * C + d]; }
#pragma unroll
    for (int d = 0; d < kAttnD / 2; ++d) {     // rotary with the position inside the window
      const float f = (float)i * __ldg(inv_freq + d);
      const float cs = cosf(f), sn = sinf(f);
      const float q1 = q[d], q2 = q[d + 32], k1 = kk[d], k2 = kk[d + 32];
      q[d] = q1 * cs + (-q2) * sn;
      q[d + 32] = q2 * cs + q1 * sn;
      kk[d] = k1 * cs + (-k2) * sn;
      kk[d + 32] = k2 * cs + k1 * sn;
    }
#pragma unroll
    for (int d = 0; d < kAttnD; ++d) ks[(size_t)i * kAttnD + d] = kk[d];
  }
  __syncthreads();
  if (i >= W) return;
  float* my = sc + (size_t)i * (W + 1);
  float mx = -3.4e38f;
  for (int j = 0; j < W; ++j) {
    float sv = 0.f;
#pragma unroll
    for (int d = 0; d < kAttnD; ++d) sv = fmaf(q[d], ks[(size_t)j * kAttnD + d], sv);
    sv *= 0.125f;   // 1 / sqrt(64)
    my[j] = sv;
    mx = fmaxf(mx, sv);
  }
  float sum = 0.f;
  for (int j = 0; j < W; ++j) { const float e = expf(my[j] - mx); my[j] = e; sum += e; }
  const float inv = 1.0f / sum;
  float* orow = out + ((long long)b * T + (long long)w * W + i) * C + h * kAttnD;
  for (int d4 = 0; d4 < kAttnD / 4; ++d4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < W; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(&vs[(size_t)j * kAttnD + 4 * d4]);
      const float pj = my[j] * inv;
      a.x = fmaf(pj, v.x, a.x); a.y = fmaf(pj, v.y, a.y); a.z = fmaf(pj, v.z, a.z); a.w = fmaf(pj, v.w, a.w);
    }
    *reinterpret_cast<float4*>(orow + 4 * d4) = a;
  }
}

void launch_local_attn(const float* qkv, float* out, const float* inv_freq, int batch, int T, int C, int heads, int window,
                       const LaunchCtx& ctx) {
  if (C != heads * kAttnD) throw Error(NC_UNSUPPORTED, "local attention: head dim must be 64");
  if (window < 1 || window > 256) throw Error(NC_UNSUPPORTED, "local attention: window size must be in 1..256");
  if (T % window != 0) throw Error(NC_INVALID_ARGUMENT, "local attention: frames not a multiple of the window");
  const long long total = (long long)batch * (T / window) * heads;
  if (total == 0) return;
  const int ev = ctx.begin();
  if (window == kAttnW) {
    local_attn_kernel<<<(unsigned)((total + 1) / 2), 64, 0, ctx.stream>>>(qkv, out, inv_freq, batch, T, C, heads);
  } else {
    const size_t smem = ((size_t)2 * window * kAttnD + (size_t)window * (window + 1)) * sizeof(float);
    NC_CUDA(cudaFuncSetAttribute(local_attn_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    local_attn_generic_kernel<<<(unsigned)total, (window + 31) / 32 * 32, smem, ctx.stream>>>(qkv, out, inv_freq, T, C, heads, window);
  }
  check_launch((int)cudaGetLastError(), "local_attn");
  ctx.end(ev, "local_attn", (double)total * 2.0 * 2 * window * window * kAttnD, 16.0 * batch * (double)T * C);
}

// ------------------------------------------------------------------------------ misc
__global__ void trim_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long in_stride, long long out_len,
                                 long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / out_len, t = i - b * out_len;
    out[i] = in[b * in_stride + t];
  }
}

void launch_trim_rows(const float* in, float* out, int batch, long long in_stride, long long out_len, const LaunchCtx& ctx) {
  const long long total = (long long)batch * out_len;
  if (total == 0) return;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)ctx.num_sms * 16;
  if (blocks > cap) blocks = cap;
  const int ev = ctx.begin();
  trim_rows_kernel<<<(unsigned)blocks, 256, 0, ctx.stream>>>(in, out, in_stride, out_len, total);
  check_launch((int)cudaGetLastError(), "trim_rows");
  ctx.end(ev, "trim_rows", 0, 8.0 * total);
}

// Philox-free counter RNG for production-mode decoder noise: N(0,1) via Box-Muller on a hashed counter.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// out[b][t], b in [0, clips): the counter is (global clip index clip0 + b, t), so a clip's noise does not depend on how
// the batch was split into micro-batches or across ranks.
__global__ void randn_kernel(float* __restrict__ out, long long n, int T, long long clip0, uint64_t seed, uint32_t stream_id) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long clip = clip0 + i / T;
    const uint32_t lo = (uint32_t)(i % T), hi = hash32((uint32_t)clip) ^ hash32((uint32_t)(clip >> 32) + 0x51ed270bu);
    const uint32_t a = hash32(lo ^ hash32(hi + 0x9e3779b9u) ^ hash32((uint32_t)seed + stream_id * 0x85ebca6bu));
    const uint32_t b = hash32(a ^ (uint32_t)(seed >> 32) ^ 0xc2b2ae35u);
    const float u1 = ((a >> 8) + 1) * (1.0f / 16777216.0f), u2 = (b >> 8) * (1.0f / 16777216.0f);
    out[i] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
}
void launch_randn(float* out, int clips, int T, long long clip0, uint64_t seed, uint32_t stream_id, const LaunchCtx& ctx) {
  const long long n = (long long)clips * T;
  if (n == 0) return;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)ctx.num_sms * 16;
  if (blocks > cap) blocks = cap;
  const int ev = ctx.begin();
  randn_kernel<<<(unsigned)blocks, 256, 0, ctx.stream>>>(out, n, T, clip0, seed, stream_id);
  check_launch((int)cudaGetLastError(), "randn");
  ctx.end(ev, "randn", 0, 4.0 * n);
}

}  // namespace nc
