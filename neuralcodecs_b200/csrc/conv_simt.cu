// fp32 CUDA-core executor of the multi-tap row-shifted GEMM (conv_plan.h).  Used when a layer
// runs in "fp32" precision mode and for shapes the tcgen05 kernel does not take (channel counts
// that are not multiples of 32, tiny test configurations).  Same math, true fp32 FMA.
#include "conv_plan.h"
#include <cuda_runtime.h>

namespace nc {

constexpr int SBM = 128, SBN = 64, SBK = 16, STM = 8, STN = 4;
constexpr int kSimtThreads = (SBM / STM) * (SBN / STN);  // 256

__device__ __forceinline__ float simt_prologue(float x, float a, int kind) {
  if (kind == PRO_SNAKE) {
    if (a == 0.f) return x;
    float s = sinf(a * x);
    return x + (s * s) / a;  // addcdiv(x, sin(ax)^2, a)  (Modules/DAC/Snake1d.cs:52)
  }
  if (kind == PRO_ELU) return x > 0.f ? x : expm1f(x);
  return x;
}

__global__ void __launch_bounds__(kSimtThreads)
conv_simt_kernel(const __grid_constant__ ConvSimtParams p) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Ws[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tile_m = blockIdx.x;
  const int b = tile_m / p.m_tiles_per_clip;
  const int mt = tile_m - b * p.m_tiles_per_clip;
  const int m0 = mt * SBM;
  const int n0 = blockIdx.y * SBN;
  const int tx = tid % (SBN / STN);  // n direction
  const int ty = tid / (SBN / STN);  // m direction
  const float* Ab = p.A + (long long)b * p.a_clip_stride;

  float acc[STM][STN];
#pragma unroll
  for (int i = 0; i < STM; ++i)
#pragma unroll
    for (int j = 0; j < STN; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < p.n_taps; ++t) {
    const SimtTap tap = p.taps[t];
    if (p.mask_bn > 0) {
      // skip taps that contribute nothing to this column block (transposed conv phases)
      bool any = false;
      for (int n = n0; n < n0 + SBN && n < p.n_pad; n += p.mask_bn)
        any |= ((p.tap_mask[n / p.mask_bn] >> t) & 1) != 0;
      if (!any) continue;
    }
    const float* Wt = p.W + tap.w_off;
    for (int k0 = 0; k0 < tap.klen; k0 += SBK) {
      // A tile: SBM rows x SBK
      for (int i = tid; i < SBM * SBK; i += kSimtThreads) {
        const int kk = i % SBK, mm = i / SBK;
        const int r = m0 + mm + tap.shift;
        const int k = k0 + kk;
        float v = 0.f;
        if (k < tap.klen && r >= 0 && r < p.a_rows) {
          const long long e = (long long)r * p.a_pitch + tap.koff + k;
          if (e < p.a_valid) {
            v = __ldg(Ab + e);
            if (p.prologue != PRO_NONE)
              v = simt_prologue(v, p.prologue == PRO_SNAKE ? __ldg(p.alpha + (tap.koff + k) % p.alpha_period) : 0.f,
                                p.prologue);
          }
        }
        As[kk][mm] = v;
      }
      for (int i = tid; i < SBN * SBK; i += kSimtThreads) {
        const int kk = i % SBK, nn = i / SBK;
        const int n = n0 + nn, k = k0 + kk;
        Ws[kk][nn] = (n < p.n_pad && k < tap.klen) ? __ldg(Wt + (long long)n * tap.klen + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SBK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * STM]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * STM + 4]);
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * STN]);
        const float a[STM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float w[STN] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
        for (int i = 0; i < STM; ++i)
#pragma unroll
          for (int j = 0; j < STN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < STM; ++i) {
    const int row = m0 + ty * STM + i;
    if (row >= p.m_rows) continue;
    const long long row_off = (long long)row * p.n_total;
    float* Drow = p.D + (long long)b * p.d_clip_stride + row_off;
    const float* Rrow = p.R ? p.R + (long long)b * p.d_clip_stride + row_off : nullptr;
    const float nz = p.noise ? __ldg(p.noise + (long long)b * p.m_rows + row) : 0.f;
#pragma unroll
    for (int j = 0; j < STN; ++j) {
      const int n = n0 + tx * STN + j;
      if (n >= p.n_valid || row_off + n >= p.d_valid) continue;
      float x = acc[i][j];
      if (p.bias) x += __ldg(p.bias + n % p.bias_period);
      if (Rrow) x = p.noise ? fmaf(nz, x, Rrow[n]) : x + Rrow[n];
      if (p.post != PRO_NONE)
        x = simt_prologue(x, p.post == PRO_SNAKE ? __ldg(p.post_alpha + n % p.post_period) : 0.f, p.post);
      if (p.act == ACT_TANH) x = tanhf(x);
      Drow[n] = x;
    }
  }
}

int launch_conv_simt(const ConvSimtParams& p, cudaStream_t stream) {
  dim3 grid(p.batch * p.m_tiles_per_clip, (p.n_valid + SBN - 1) / SBN);
  if (grid.x == 0 || grid.y == 0) return 0;
  conv_simt_kernel<<<grid, kSimtThreads, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace nc
