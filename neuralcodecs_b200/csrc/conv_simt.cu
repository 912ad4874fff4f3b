// fp32 CUDA-core executor of the multi-tap row-shifted GEMM (conv_plan.h).  Used when a layer runs in "fp32"
// precision mode (the Encodec encoder by default: its 128-d residual VQ needs true fp32 accumulation to keep the
// codes) and for shapes the tcgen05 kernel does not take (channel counts that are not multiples of 32, tiny test
// configurations).  Same math, true fp32 FMA, one accumulator per output, products added in (tap, k) order.
//
// 128 x BN output tile per CTA (BN = 32 / 64 / 128 picked per layer), 256 threads, 8 x BN/16 outputs per thread,
// K step 16.  Global -> register -> shared double buffering: the next stage's A and W fragments are fetched as
// float4 along K (channels-last: contiguous) while the current stage is multiplied; one __syncthreads per stage.
#include "conv_plan.h"
#include <cuda_runtime.h>

namespace nc {

constexpr int SBM = 128, SBK = 16, STM = 8;
constexpr int kSimtThreads = 256;

__device__ __forceinline__ float simt_prologue(float x, float a, int kind) {
  if (kind == PRO_SNAKE) {
    if (a == 0.f) return x;
    float s = sinf(a * x);
    return x + (s * s) / a;  // addcdiv(x, sin(ax)^2, a)  (Modules/DAC/Snake1d.cs:52)
  }
  if (kind == PRO_ELU) return x > 0.f ? x : expm1f(x);
  return x;
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(kSimtThreads)
conv_simt_kernel(const __grid_constant__ ConvSimtParams p) {
  constexpr int TN = BN / 16;                  // outputs per thread along N
  constexpr int WQ = (BN * 4 + kSimtThreads - 1) / kSimtThreads;   // float4 weight fetches per thread per stage
  __shared__ __align__(16) float As[2][SBK][SBM + 4];
  __shared__ __align__(16) float Ws[2][SBK][BN + 4];
  const int tid = threadIdx.x;
  const int tile_m = blockIdx.x;
  const int b = tile_m / p.m_tiles_per_clip;
  const int mt = tile_m - b * p.m_tiles_per_clip;
  const int m0 = mt * SBM;
  const int n0 = blockIdx.y * BN;
  const int tx = tid & 15;   // n direction
  const int ty = tid >> 4;   // m direction
  const float* Ab = p.A + (long long)b * p.a_clip_stride;

  float acc[STM][TN];
#pragma unroll
  for (int i = 0; i < STM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // taps that contribute to this column block (transposed-conv phases own disjoint column ranges)
  unsigned live = 0;
  for (int t = 0; t < p.n_taps; ++t) {
    bool any = p.mask_bn <= 0;
    if (!any)
      for (int n = n0; n < n0 + BN && n < p.n_pad; n += p.mask_bn) any |= ((p.tap_mask[n / p.mask_bn] >> t) & 1) != 0;
    if (any) live |= 1u << t;
  }

  float4 ra[2], rw[WQ];
  // fetch stage (tap t, k0) into registers, activation applied on the way
  auto fetch = [&](int t, int k0) {
    const SimtTap tap = p.taps[t];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kSimtThreads;
      const int mm = idx >> 2, k = k0 + (idx & 3) * 4;
      const int r = m0 + mm + tap.shift;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (r >= 0 && r < p.a_rows && k < tap.klen) {
        const long long e = (long long)r * p.a_pitch + tap.koff + k;
        if (VEC) {
          if (e < p.a_valid) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(Ab + e));
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k + j < tap.klen && e + j < p.a_valid) v[j] = __ldg(Ab + e + j);
        }
        if (p.prologue != PRO_NONE) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool in = VEC ? (e < p.a_valid) : (k + j < tap.klen && e + j < p.a_valid);
            if (in)
              v[j] = simt_prologue(v[j], p.prologue == PRO_SNAKE ? __ldg(p.alpha + (tap.koff + k + j) % p.alpha_period) : 0.f,
                                   p.prologue);
          }
        }
      }
      ra[q] = make_float4(v[0], v[1], v[2], v[3]);
    }
    const float* Wt = p.W + tap.w_off;
#pragma unroll
    for (int q = 0; q < WQ; ++q) {
      const int idx = tid + q * kSimtThreads;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BN * 4) {
        const int nn = idx >> 2, k = k0 + (idx & 3) * 4;
        const int n = n0 + nn;
        if (n < p.n_pad && k < tap.klen) {
          const float* src = Wt + (long long)n * tap.klen + k;
          if (VEC) {
            w = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            w.x = __ldg(src);
            if (k + 1 < tap.klen) w.y = __ldg(src + 1);
            if (k + 2 < tap.klen) w.z = __ldg(src + 2);
            if (k + 3 < tap.klen) w.w = __ldg(src + 3);
          }
        }
      }
      rw[q] = w;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kSimtThreads;
      const int mm = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][mm] = ra[q].x; As[buf][kq + 1][mm] = ra[q].y; As[buf][kq + 2][mm] = ra[q].z; As[buf][kq + 3][mm] = ra[q].w;
    }
#pragma unroll
    for (int q = 0; q < WQ; ++q) {
      const int idx = tid + q * kSimtThreads;
      if (idx < BN * 4) {
        const int nn = idx >> 2, kq = (idx & 3) * 4;
        Ws[buf][kq + 0][nn] = rw[q].x; Ws[buf][kq + 1][nn] = rw[q].y; Ws[buf][kq + 2][nn] = rw[q].z; Ws[buf][kq + 3][nn] = rw[q].w;
      }
    }
  };
  // stage iterator over (live tap, k0)
  int t_cur = 0, k_cur = 0;
  auto seek = [&]() {   // move (t_cur, k_cur) to the next valid stage at or after the current position
    while (t_cur < p.n_taps && (!((live >> t_cur) & 1u) || k_cur >= p.taps[t_cur].klen)) { ++t_cur; k_cur = 0; }
    return t_cur < p.n_taps;
  };

  bool have = seek();
  if (have) {
    fetch(t_cur, k_cur);
    stash(0);
    k_cur += SBK;
  }
  __syncthreads();
  int buf = 0;
  while (have) {
    const bool next = seek();
    if (next) {
      fetch(t_cur, k_cur);
      k_cur += SBK;
    }
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * STM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * STM + 4]);
      const float a[STM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float w[TN];
      if constexpr (TN == 8) {   // columns tx*4..+3 and 64 + tx*4..+3: each half-warp reads 256 contiguous bytes
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
        const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      } else if constexpr (TN == 4) {
        const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
      } else {
        const float2 w0 = *reinterpret_cast<const float2*>(&Ws[buf][kk][tx * 2]);
        w[0] = w0.x; w[1] = w0.y;
      }
#pragma unroll
      for (int i = 0; i < STM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (next) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
    have = next;
  }

  // per-column constants once per thread (the integer modulo is ~20 instructions; it used to run per output element)
  float bj[TN], pj[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + (TN == 8 ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx * TN + j);
    bj[j] = (p.bias && n < p.n_valid) ? __ldg(p.bias + n % p.bias_period) : 0.f;
    pj[j] = (p.post == PRO_SNAKE && n < p.n_valid) ? __ldg(p.post_alpha + n % p.post_period) : 0.f;
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
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (TN == 8 ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx * TN + j);
      if (n >= p.n_valid || row_off + n >= p.d_valid) continue;
      float x = acc[i][j] + bj[j];
      if (Rrow) x = p.noise ? fmaf(nz, x, Rrow[n]) : x + Rrow[n];
      if (p.post != PRO_NONE) x = simt_prologue(x, pj[j], p.post);
      if (p.act == ACT_TANH) x = tanhf(x);
      Drow[n] = x;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Dense k3 variant (stride 1, dilation 1, three taps over the same channel range -- Encodec's resnet k3 convs on the
// narrow layers): the 128 + 2 input rows of a K chunk are fetched, activated and staged ONCE and the three taps read
// them at row offsets 0, 1, 2 (the general kernel re-fetches and re-activates the tile once per tap), and the
// per-thread 8-row window is held in registers: 3 + 3 LDS.128 per 96 FMAs.  K chunks outer, taps inner.
template <int BN>
__global__ void __launch_bounds__(kSimtThreads)
conv_simt_k3_kernel(const __grid_constant__ ConvSimtParams p) {
  constexpr int TN = BN / 16, NT = 3, ROWS = SBM + NT - 1;
  constexpr int AQ = (ROWS * 4 + kSimtThreads - 1) / kSimtThreads;       // float4 A fetches per thread per stage
  constexpr int WQ = (NT * BN * 4 + kSimtThreads - 1) / kSimtThreads;    // float4 W fetches per thread per stage
  __shared__ __align__(16) float As[2][SBK][SBM + 12];                   // rows 0..ROWS-1 used; reads reach ty*8 + 11
  __shared__ __align__(16) float Ws[2][NT][SBK][BN + 4];
  const int tid = threadIdx.x;
  const int tile_m = blockIdx.x;
  const int b = tile_m / p.m_tiles_per_clip;
  const int m0 = (tile_m - b * p.m_tiles_per_clip) * SBM;
  const int n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;
  const float* Ab = p.A + (long long)b * p.a_clip_stride;
  const int klen = p.taps[0].klen, koff = p.taps[0].koff, smin = p.taps[0].shift;

  float acc[STM][TN];
#pragma unroll
  for (int i = 0; i < STM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[AQ], rw[WQ];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < AQ; ++q) {
      const int idx = tid + q * kSimtThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < ROWS * 4) {
        const int mm = idx >> 2, k = k0 + (idx & 3) * 4;
        const int r = m0 + mm + smin;
        if (r >= 0 && r < p.a_rows && k < klen) {
          const long long e = (long long)r * p.a_pitch + koff + k;
          if (e < p.a_valid) {
            v = __ldg(reinterpret_cast<const float4*>(Ab + e));
            if (p.prologue != PRO_NONE) {
              const int ai = (koff + k) % p.alpha_period;
              v.x = simt_prologue(v.x, p.prologue == PRO_SNAKE ? __ldg(p.alpha + ai) : 0.f, p.prologue);
              v.y = simt_prologue(v.y, p.prologue == PRO_SNAKE ? __ldg(p.alpha + (ai + 1) % p.alpha_period) : 0.f, p.prologue);
              v.z = simt_prologue(v.z, p.prologue == PRO_SNAKE ? __ldg(p.alpha + (ai + 2) % p.alpha_period) : 0.f, p.prologue);
              v.w = simt_prologue(v.w, p.prologue == PRO_SNAKE ? __ldg(p.alpha + (ai + 3) % p.alpha_period) : 0.f, p.prologue);
            }
          }
        }
      }
      ra[q] = v;
    }
#pragma unroll
    for (int q = 0; q < WQ; ++q) {
      const int idx = tid + q * kSimtThreads;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < NT * BN * 4) {
        const int t = idx / (BN * 4), rem = idx - t * (BN * 4);
        const int nn = rem >> 2, k = k0 + (rem & 3) * 4;
        const int n = n0 + nn;
        if (n < p.n_pad && k < klen) w = __ldg(reinterpret_cast<const float4*>(p.W + p.taps[t].w_off + (long long)n * klen + k));
      }
      rw[q] = w;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < AQ; ++q) {
      const int idx = tid + q * kSimtThreads;
      if (idx < ROWS * 4) {
        const int mm = idx >> 2, kq = (idx & 3) * 4;
        As[buf][kq + 0][mm] = ra[q].x; As[buf][kq + 1][mm] = ra[q].y; As[buf][kq + 2][mm] = ra[q].z; As[buf][kq + 3][mm] = ra[q].w;
      }
    }
#pragma unroll
    for (int q = 0; q < WQ; ++q) {
      const int idx = tid + q * kSimtThreads;
      if (idx < NT * BN * 4) {
        const int t = idx / (BN * 4), rem = idx - t * (BN * 4);
        const int nn = rem >> 2, kq = (rem & 3) * 4;
        Ws[buf][t][kq + 0][nn] = rw[q].x; Ws[buf][t][kq + 1][nn] = rw[q].y; Ws[buf][t][kq + 2][nn] = rw[q].z; Ws[buf][t][kq + 3][nn] = rw[q].w;
      }
    }
  };

  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < klen; k0 += SBK) {
    const bool next = k0 + SBK < klen;
    if (next) fetch(k0 + SBK);
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      // rows ty*8 .. ty*8 + 11 of this K column: the 8-row output window plus the two extra rows the taps reach
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * STM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * STM + 4]);
      const float4 a2 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * STM + 8]);
      const float a[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float w[TN];
        if constexpr (TN == 4) {
          const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][t][kk][tx * 4]);
          w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
        } else {
          const float2 w0 = *reinterpret_cast<const float2*>(&Ws[buf][t][kk][tx * 2]);
          w[0] = w0.x; w[1] = w0.y;
        }
#pragma unroll
        for (int i = 0; i < STM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i + t], w[j], acc[i][j]);
      }
    }
    if (next) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  float bj[TN], pj[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + tx * TN + j;
    bj[j] = (p.bias && n < p.n_valid) ? __ldg(p.bias + n % p.bias_period) : 0.f;
    pj[j] = (p.post == PRO_SNAKE && n < p.n_valid) ? __ldg(p.post_alpha + n % p.post_period) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < STM; ++i) {
    const int row = m0 + ty * STM + i;
    if (row >= p.m_rows) continue;
    const long long row_off = (long long)row * p.n_total;
    float* Drow = p.D + (long long)b * p.d_clip_stride + row_off;
    const float* Rrow = p.R ? p.R + (long long)b * p.d_clip_stride + row_off : nullptr;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.n_valid || row_off + n >= p.d_valid) continue;
      float x = acc[i][j] + bj[j];
      if (Rrow) x += Rrow[n];
      if (p.post != PRO_NONE) x = simt_prologue(x, pj[j], p.post);
      if (p.act == ACT_TANH) x = tanhf(x);
      Drow[n] = x;
    }
  }
}

// the k3 variant applies to: three taps, same channel window, consecutive shifts, no masks / noise, float4-aligned views
static bool dense_k3_ok(const ConvSimtParams& p, bool vec, int bn) {
  if (!vec || bn > 64 || p.n_taps != 3 || p.mask_bn > 0 || p.noise) return false;
  for (int t = 1; t < 3; ++t)
    if (p.taps[t].koff != p.taps[0].koff || p.taps[t].klen != p.taps[0].klen || p.taps[t].shift != p.taps[0].shift + t) return false;
  return true;
}

template <int BN>
static int launch_bn(const ConvSimtParams& p, bool vec, cudaStream_t stream) {
  dim3 grid(p.batch * p.m_tiles_per_clip, (p.n_valid + BN - 1) / BN);
  if (grid.x == 0 || grid.y == 0) return 0;
  if (vec)
    conv_simt_kernel<BN, true><<<grid, kSimtThreads, 0, stream>>>(p);
  else
    conv_simt_kernel<BN, false><<<grid, kSimtThreads, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

int launch_conv_simt(const ConvSimtParams& p, cudaStream_t stream) {
  // float4 fetches along K need every tap's window, the row pitch, the clip stride and the valid extent 16-byte aligned
  bool vec = (p.a_pitch % 4 == 0) && (p.a_clip_stride % 4 == 0) && (p.a_valid % 4 == 0) &&
             (reinterpret_cast<uintptr_t>(p.A) % 16 == 0) && (reinterpret_cast<uintptr_t>(p.W) % 16 == 0);
  for (int t = 0; t < p.n_taps; ++t)
    vec = vec && (p.taps[t].koff % 4 == 0) && (p.taps[t].klen % 4 == 0) && (p.taps[t].w_off % 4 == 0);
  // column tile: least padded columns, larger tile on ties
  const int n = p.n_valid;
  int best = 128;
  long best_cols = (long)((n + 127) / 128) * 128;
  for (int bn : {64, 32}) {
    const long cols = (long)((n + bn - 1) / bn) * bn;
    if (cols < best_cols) { best = bn; best_cols = cols; }
  }
  if (dense_k3_ok(p, vec, best)) {
    dim3 grid(p.batch * p.m_tiles_per_clip, (p.n_valid + best - 1) / best);
    if (grid.x == 0 || grid.y == 0) return 0;
    if (best == 64) conv_simt_k3_kernel<64><<<grid, kSimtThreads, 0, stream>>>(p);
    else conv_simt_k3_kernel<32><<<grid, kSimtThreads, 0, stream>>>(p);
    return (int)cudaGetLastError();
  }
  switch (best) {
    case 128: return launch_bn<128>(p, vec, stream);
    case 64: return launch_bn<64>(p, vec, stream);
    default: return launch_bn<32>(p, vec, stream);
  }
}

}  // namespace nc
