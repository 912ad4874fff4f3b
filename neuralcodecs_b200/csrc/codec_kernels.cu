// HBM-bound kernels of the codec hot path.  See codec_kernels.h for the contracts.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include "codec_kernels.h"
#include "umma.cuh"

#include <cfloat>

namespace nc {

// ------------------------------------------------------------------------------ first conv (Cin = 1)
// Each thread owns 4 output channels for the whole launch (weights + bias in registers) and walks time;
// a warp writes 2 consecutive time steps x 64 channels = 512 contiguous bytes per store instruction.
template <int K>
__global__ void __launch_bounds__(256)
conv_cin1_kernel(const float* __restrict__ in, long long in_stride, int in_len, float* __restrict__ out, int t_out,
                 int cout, const float* __restrict__ w, const float* __restrict__ bias, int dil, int pad, int batch,
                 int reflect, long long out_clip_stride) {
  const int c4n = cout / 4;
  const int lanes_t = blockDim.x / c4n;              // time steps covered by one block iteration
  const int c4 = threadIdx.x % c4n, tl = threadIdx.x / c4n;
  if (tl >= lanes_t) return;
  float wr[4][K], br[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    br[i] = bias ? __ldg(bias + 4 * c4 + i) : 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) wr[i][j] = __ldg(w + (4 * c4 + i) * K + j);
  }
  // (clip, time) walked incrementally: one 64-bit division per thread instead of one per output step
  const long long bt0 = (long long)blockIdx.x * lanes_t + tl, step = (long long)gridDim.x * lanes_t;
  int b = (int)(bt0 / t_out), t = (int)(bt0 - (long long)b * t_out);
  const int step_b = (int)(step / t_out), step_t = (int)(step - (long long)step_b * t_out);
  for (; b < batch; b += step_b, t += step_t) {
    if (t >= t_out) { t -= t_out; ++b; if (b >= batch) break; }
    const float* x = in + (long long)b * in_stride;
    float a0 = br[0], a1 = br[1], a2 = br[2], a3 = br[3];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      int ti = t + j * dil - pad;
      if (reflect) {   // F.pad(mode="reflect"): x[-i] = x[i], x[n-1+i] = x[n-1-i]
        if (ti < 0) ti = -ti;
        if (ti >= in_len) ti = 2 * (in_len - 1) - ti;
      }
      const float v = (ti >= 0 && ti < in_len) ? __ldg(x + ti) : 0.f;
      a0 = fmaf(wr[0][j], v, a0); a1 = fmaf(wr[1][j], v, a1); a2 = fmaf(wr[2][j], v, a2); a3 = fmaf(wr[3][j], v, a3);
    }
    reinterpret_cast<float4*>(out + (long long)b * out_clip_stride)[(long long)t * c4n + c4] = make_float4(a0, a1, a2, a3);
  }
}

void launch_conv_cin1(const float* in, long long in_stride, int in_len, float* out, int t_out, int cout,
                      const float* w, const float* bias, int k, int dil, int pad, int batch, const LaunchCtx& ctx,
                      int reflect, long long out_clip_stride) {
  if (out_clip_stride == 0) out_clip_stride = (long long)t_out * cout;
  if (cout % 4 != 0 || cout > 1024) throw Error(NC_UNSUPPORTED, "conv_cin1: Cout must be a multiple of 4, at most 1024");
  if (k != 7 && k != 3) throw Error(NC_UNSUPPORTED, "conv_cin1: kernel size must be 3 or 7");
  const long long total_t = (long long)batch * t_out;
  if (total_t == 0) return;
  const int lanes_t = 256 / (cout / 4);
  long long blocks = (total_t + lanes_t - 1) / lanes_t;
  const long long cap = (long long)ctx.num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int ev = ctx.begin();
  if (k == 7)
    conv_cin1_kernel<7><<<(unsigned)blocks, 256, 0, ctx.stream>>>(in, in_stride, in_len, out, t_out, cout, w, bias, dil, pad, batch, reflect, out_clip_stride);
  else
    conv_cin1_kernel<3><<<(unsigned)blocks, 256, 0, ctx.stream>>>(in, in_stride, in_len, out, t_out, cout, w, bias, dil, pad, batch, reflect, out_clip_stride);
  check_launch((int)cudaGetLastError(), "conv_cin1");
  ctx.end(ev, "conv_cin1", 2.0 * k * cout * (double)t_out * batch, 4.0 * batch * ((double)in_len + (double)t_out * cout));
}

// ------------------------------------------------------------------------------ last conv (Cout = 1)
// out[b, t] = act(bias + sum_j sum_c w[j, c] * x[b, t + j - pad, c]); x channels-last, already carrying the
// preceding Snake.  One block produces kCoutTile consecutive samples: every input row is read ONCE (8 lanes
// per row, 384 contiguous bytes for C = 96), its K partial dot products go to shared memory, then thread t
// gathers its K diagonal entries.  HBM-bound: C*4 bytes in, 4 bytes out per sample.
constexpr int kCoutTile = 256;
template <int K>
__global__ void __launch_bounds__(256)
conv_cout1_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int C, const float* __restrict__ w /*[K][C]*/,
                  const float* __restrict__ bias, int pad, int act, int tiles_per_clip, int reflect, long long in_clip_stride) {
  extern __shared__ float sm[];
  float* sw = sm;                    // [K][C]
  float* sp = sm + K * C;            // [kCoutTile + K - 1][K]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = w[i];
  const int b = blockIdx.x / tiles_per_clip;
  const int t0 = (blockIdx.x - b * tiles_per_clip) * kCoutTile;
  const float* x = in + (long long)b * in_clip_stride;
  __syncthreads();
  const int grp = threadIdx.x >> 3, l8 = threadIdx.x & 7;   // 32 row groups of 8 lanes
  const int rows = kCoutTile + K - 1;
  const int c4_per_lane = C / 32;                            // float4 per lane
  for (int r0 = 0; r0 < rows; r0 += 32) {   // uniform trip count: the shuffles below need the whole warp
    const int r = r0 + grp;
    int tr = t0 + r - pad;
    if (reflect) {
      if (tr < 0) tr = -tr;
      if (tr >= T) tr = 2 * (T - 1) - tr;
    }
    float acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.f;
    if (r < rows && tr >= 0 && tr < T) {
      const float4* xr = reinterpret_cast<const float4*>(x + (long long)tr * C);
      for (int i = 0; i < c4_per_lane; ++i) {
        const int c4 = i * 8 + l8;
        const float4 v = __ldg(xr + c4);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float4 ww = *reinterpret_cast<const float4*>(sw + j * C + 4 * c4);
          acc[j] = fmaf(v.x, ww.x, acc[j]); acc[j] = fmaf(v.y, ww.y, acc[j]);
          acc[j] = fmaf(v.z, ww.z, acc[j]); acc[j] = fmaf(v.w, ww.w, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
      acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    }
    if (l8 == 0 && r < rows) {
#pragma unroll
      for (int j = 0; j < K; ++j) sp[r * K + j] = acc[j];
    }
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t < T) {
    float v = bias ? __ldg(bias) : 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) v += sp[(threadIdx.x + j) * K + j];
    if (act == 1) v = tanhf(v);
    out[(long long)b * T + t] = v;
  }
}

void launch_conv_cout1(const float* in, float* out, int T, int C, const float* w_kc, const float* bias, int k, int pad,
                       int act, int batch, const LaunchCtx& ctx, int reflect, long long in_clip_stride) {
  if (in_clip_stride == 0) in_clip_stride = (long long)T * C;
  if (C % 32 != 0 || k != 7) throw Error(NC_UNSUPPORTED, "conv_cout1: needs C % 32 == 0 and kernel size 7");
  if ((long long)batch * T == 0) return;
  const int tiles = (T + kCoutTile - 1) / kCoutTile;
  const size_t smem = (size_t)(7 * C + (kCoutTile + 6) * 7) * sizeof(float);
  const int ev = ctx.begin();
  conv_cout1_kernel<7><<<(unsigned)(batch * tiles), 256, smem, ctx.stream>>>(in, out, T, C, w_kc, bias, pad, act, tiles, reflect, in_clip_stride);
  check_launch((int)cudaGetLastError(), "conv_cout1");
  ctx.end(ev, "conv_cout1", 2.0 * k * C * (double)T * batch, 4.0 * batch * ((double)T * C + T));
}

// ------------------------------------------------------------------------------ transposes
__global__ void f32_to_f16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    out[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}
void launch_f32_to_f16(const float* in, void* out16, long long n, const LaunchCtx& ctx) {
  if (n == 0) return;
  if (n % 4 != 0) throw Error(NC_INTERNAL, "f32_to_f16: element count must be a multiple of 4");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)ctx.num_sms * 8;
  if (blocks > cap) blocks = cap;
  const int ev = ctx.begin();
  f32_to_f16_kernel<<<(unsigned)blocks, 256, 0, ctx.stream>>>(reinterpret_cast<const float4*>(in), static_cast<uint2*>(out16), n4);
  check_launch((int)cudaGetLastError(), "f32_to_f16");
  ctx.end(ev, "f32_to_f16", 0, 6.0 * n);
}

// in: [B][R][Cn] -> out: [B][Cn][R]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const float* src = in + (long long)b * R * Cn;
  float* dst = out + (long long)b * R * Cn;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cn) ? src[(long long)r * Cn + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cn) dst[(long long)c * R + r] = tile[threadIdx.x][i];
  }
}

static void launch_transpose(const float* in, float* out, int batch, int R, int Cn, const LaunchCtx& ctx,
                             const char* name) {
  if (batch == 0 || R == 0 || Cn == 0) return;
  dim3 grid((Cn + 31) / 32, (R + 31) / 32, batch), block(32, 8);
  const int ev = ctx.begin();
  transpose_kernel<<<grid, block, 0, ctx.stream>>>(in, out, R, Cn);
  check_launch((int)cudaGetLastError(), name);
  ctx.end(ev, name, 0, 8.0 * batch * (double)R * Cn);
}
void launch_transpose_ct_to_tc(const float* in, float* out, int batch, int C, int T, const LaunchCtx& ctx) {
  launch_transpose(in, out, batch, C, T, ctx, "transpose");
}
void launch_transpose_tc_to_ct(const float* in, float* out, int batch, int C, int T, const LaunchCtx& ctx) {
  launch_transpose(in, out, batch, T, C, ctx, "transpose");
}

// ------------------------------------------------------------------------------ fused RVQ encode
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kVqD = 8;

template <int CPL>
__global__ void __launch_bounds__(256)
rvq_encode_kernel(const RvqWeights w, const float* __restrict__ z, float* __restrict__ zq, int64_t* __restrict__ codes,
                  float* __restrict__ latents, int batch, int T, int n_q) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long frames = (long long)batch * T;
  const int Dz = w.Dz, K = w.K;
  for (long long f = warp0; f < frames; f += nwarps) {
    const int b = (int)(f / T), t = (int)(f % T);
    const float* zr = z + f * Dz;
    float r[CPL], acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      r[i] = __ldg(zr + lane + 32 * i);
      acc[i] = 0.f;
    }
    for (int s = 0; s < n_q; ++s) {
      // ---- in_proj (WNConv1d Dz -> 8, k = 1): Modules/DAC/VectorQuantizer.cs:70
      const float* Win = w.in_w + (size_t)s * kVqD * Dz;
      float ze[kVqD];
#pragma unroll
      for (int d = 0; d < kVqD; ++d) {
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) p = fmaf(__ldg(Win + (size_t)d * Dz + lane + 32 * i), r[i], p);
        ze[d] = warp_sum(p) + __ldg(w.in_b + s * kVqD + d);
      }
      // ---- nearest codebook entry, un-normalised expanded form (VectorQuantizer.cs:110-121)
      float e2 = 0.f;
#pragma unroll
      for (int d = 0; d < kVqD; ++d) e2 = fmaf(ze[d], ze[d], e2);
      const float* cb = w.cb + (size_t)s * K * kVqD;
      const float* csq = w.cb_sq + (size_t)s * K;
      float best = FLT_MAX;
      int best_k = 0;
      for (int k = lane; k < K; k += 32) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * kVqD));
        const float4 c1 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * kVqD) + 1);
        float dot = ze[0] * c0.x;
        dot = fmaf(ze[1], c0.y, dot); dot = fmaf(ze[2], c0.z, dot); dot = fmaf(ze[3], c0.w, dot);
        dot = fmaf(ze[4], c1.x, dot); dot = fmaf(ze[5], c1.y, dot); dot = fmaf(ze[6], c1.z, dot);
        dot = fmaf(ze[7], c1.w, dot);
        const float dist = (e2 + __ldg(csq + k)) - 2.0f * dot;
        if (dist < best) { best = dist; best_k = k; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        if (od < best || (od == best && ok < best_k)) { best = od; best_k = ok; }  // argmin: lowest index wins
      }
      // ---- lookup + straight-through arithmetic (VectorQuantizer.cs:81) + out_proj (:82)
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)best_k * kVqD));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(cb + (size_t)best_k * kVqD) + 1);
      float q[kVqD] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int d = 0; d < kVqD; ++d) q[d] = ze[d] + (q[d] - ze[d]);
      const float* Wout = w.out_w + (size_t)s * Dz * kVqD;
      const float* bout = w.out_b + (size_t)s * Dz;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wout + (size_t)c * kVqD));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wout + (size_t)c * kVqD) + 1);
        float v = w0.x * q[0];
        v = fmaf(w0.y, q[1], v); v = fmaf(w0.z, q[2], v); v = fmaf(w0.w, q[3], v);
        v = fmaf(w1.x, q[4], v); v = fmaf(w1.y, q[5], v); v = fmaf(w1.z, q[6], v); v = fmaf(w1.w, q[7], v);
        v += __ldg(bout + c);
        acc[i] += v;   // zQ.add_(zQi)        ResidualVectorQuantizer.cs:68
        r[i] -= v;     // residual.sub_(zQi)  ResidualVectorQuantizer.cs:69
      }
      if (codes && lane == 0) codes[((long long)b * n_q + s) * T + t] = best_k;
      if (latents && lane < kVqD) {
        float mine = ze[0];
#pragma unroll
        for (int d = 1; d < kVqD; ++d) mine = lane == d ? ze[d] : mine;
        latents[((long long)b * n_q * kVqD + s * kVqD + lane) * T + t] = mine;
      }
    }
    if (zq) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) zq[f * Dz + lane + 32 * i] = acc[i];
    }
  }
}

// ------------------------------------------------------------------------------ fused RVQ encode, block version
// The warp kernel above re-reads every stage's in_proj / codebook / out_proj (100 KB) from L1/L2 for every frame: 98
// GB/s of algorithmic traffic (1.5 % of HBM peak), latency-bound.  Here one persistent CTA per SM stages a stage's weights
// in shared memory ONCE per pass of 16 frames (1-D bulk async copies, double-buffered so stage s+1 streams in while
// stage s computes) and each warp carries TWO frames in registers, so every weight fetched from shared memory feeds
// two FMAs.  Arithmetic and evaluation order are those of the warp kernel (lane-strided in_proj partial sums in
// ascending channel order + butterfly, expanded-form distance, lowest-index argmin): codes are bit-identical to it.
// The kernel's own bound is shared-memory bandwidth / fp32 FMA issue, not HBM: 442 KFLOP per 8.3 KB frame is 53 FLOP/B,
// five times the fp32 ridge of the machine.
std::vector<float> rvq_stage_blob(int Dz, int D, int K, const float* in_w, const float* in_b, const float* cb, const float* cb_sq,
                                  const float* out_w, const float* out_b) {
  std::vector<float> v;
  if (D != kVqD || Dz % 128 != 0) return v;
  const int cpl = Dz / 32;
  v.reserve((size_t)17 * Dz + 9 * K + 8);
  for (int d = 0; d < D; ++d)
    for (int i4 = 0; i4 < cpl / 4; ++i4)
      for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < 4; ++j) v.push_back(in_w[(size_t)d * Dz + lane + 32 * (4 * i4 + j)]);
  for (int d = 0; d < D; ++d) v.push_back(in_b[d]);
  for (int k = 0; k < K; ++k) for (int d = 0; d < 4; ++d) v.push_back(cb[(size_t)k * D + d]);
  for (int k = 0; k < K; ++k) for (int d = 4; d < 8; ++d) v.push_back(cb[(size_t)k * D + d]);
  for (int k = 0; k < K; ++k) v.push_back(cb_sq[k]);
  for (int c = 0; c < Dz; ++c) for (int d = 0; d < 4; ++d) v.push_back(out_w[(size_t)c * D + d]);
  for (int c = 0; c < Dz; ++c) for (int d = 4; d < 8; ++d) v.push_back(out_w[(size_t)c * D + d]);
  for (int c = 0; c < Dz; ++c) v.push_back(out_b[c]);
  return v;
}

constexpr int kRvqWarps = 8, kRvqF = 2;

template <int CPL>
__global__ void __launch_bounds__(kRvqWarps * 32, 1)
rvq_encode_block_kernel(const RvqWeights w, const float* __restrict__ z, float* __restrict__ zq, int64_t* __restrict__ codes,
                        float* __restrict__ latents, int batch, int T, int n_q) {
  extern __shared__ __align__(128) uint8_t rvq_smem[];
  __shared__ uint64_t full[2];
  const int Dz = w.Dz, K = w.K;
  const uint32_t blob_bytes = (uint32_t)w.blob_floats * 4u;
  const uint32_t buf_stride = (blob_bytes + 127u) & ~127u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long frames = (long long)batch * T;
  const long long passes = (frames + kRvqWarps * kRvqF - 1) / (kRvqWarps * kRvqF);
  const long long my_passes = passes > (long long)blockIdx.x ? (passes - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long n_total = my_passes * n_q;
  if (tid == 0) {
    ptx::mbar_init(&full[0], 1);
    ptx::mbar_init(&full[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  auto fetch = [&](long long n) {   // stage (n % n_q)'s image -> buffer (n & 1); thread 0 only
    uint8_t* dst = rvq_smem + (size_t)(n & 1) * buf_stride;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(w.blob + (size_t)(n % n_q) * w.blob_floats);
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive_expect_tx(&full[n & 1], blob_bytes);
    for (uint32_t off = 0; off < blob_bytes; off += 32768u) {
      const uint32_t len = blob_bytes - off < 32768u ? blob_bytes - off : 32768u;
      ptx::bulk_g2s(dst + off, src + off, len, &full[n & 1]);
    }
  };
  if (tid == 0 && n_total > 0) fetch(0);

  float r[kRvqF][CPL], acc[kRvqF][CPL];
  long long fidx[kRvqF];
  bool valid[kRvqF];
  for (long long n = 0; n < n_total; ++n) {
    const int s = (int)(n % n_q);
    if (s == 0) {
      const long long pass = (long long)blockIdx.x + (n / n_q) * gridDim.x;
#pragma unroll
      for (int f = 0; f < kRvqF; ++f) {
        const long long fi = pass * (kRvqWarps * kRvqF) + warp * kRvqF + f;
        valid[f] = fi < frames;
        fidx[f] = valid[f] ? fi : frames - 1;
        const float* zr = z + fidx[f] * Dz;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          r[f][i] = __ldg(zr + lane + 32 * i);
          acc[f][i] = 0.f;
        }
      }
    }
    if (tid == 0 && n + 1 < n_total) fetch(n + 1);   // buffer (n+1)&1 was released by the barrier that ended step n-1
    ptx::mbar_wait(&full[n & 1], (uint32_t)(n >> 1) & 1u);
    const float* sb = reinterpret_cast<const float*>(rvq_smem + (size_t)(n & 1) * buf_stride);
    const float4* win4 = reinterpret_cast<const float4*>(sb);
    const float* in_b = sb + (size_t)kVqD * Dz;
    const float4* cb_lo = reinterpret_cast<const float4*>(in_b + kVqD);
    const float4* cb_hi = cb_lo + K;
    const float* csq = reinterpret_cast<const float*>(cb_hi + K);
    const float4* wo_lo = reinterpret_cast<const float4*>(csq + K);
    const float4* wo_hi = wo_lo + Dz;
    const float* out_b = reinterpret_cast<const float*>(wo_hi + Dz);

    // ---- in_proj (Modules/DAC/VectorQuantizer.cs:70)
    float ze[kRvqF][kVqD];
#pragma unroll
    for (int d = 0; d < kVqD; ++d) {
      float p[kRvqF];
#pragma unroll
      for (int f = 0; f < kRvqF; ++f) p[f] = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < CPL / 4; ++i4) {
        const float4 wv = win4[(d * (CPL / 4) + i4) * 32 + lane];
#pragma unroll
        for (int f = 0; f < kRvqF; ++f) {
          p[f] = fmaf(wv.x, r[f][4 * i4 + 0], p[f]);
          p[f] = fmaf(wv.y, r[f][4 * i4 + 1], p[f]);
          p[f] = fmaf(wv.z, r[f][4 * i4 + 2], p[f]);
          p[f] = fmaf(wv.w, r[f][4 * i4 + 3], p[f]);
        }
      }
      const float bd = in_b[d];
#pragma unroll
      for (int f = 0; f < kRvqF; ++f) ze[f][d] = warp_sum(p[f]) + bd;
    }
    // ---- nearest codebook entry, un-normalised expanded form (VectorQuantizer.cs:110-121)
    float e2[kRvqF], best[kRvqF];
    int best_k[kRvqF];
#pragma unroll
    for (int f = 0; f < kRvqF; ++f) {
      e2[f] = 0.f;
#pragma unroll
      for (int d = 0; d < kVqD; ++d) e2[f] = fmaf(ze[f][d], ze[f][d], e2[f]);
      best[f] = FLT_MAX;
      best_k[f] = 0;
    }
    for (int k = lane; k < K; k += 32) {
      const float4 c0 = cb_lo[k], c1 = cb_hi[k];
      const float cs = csq[k];
#pragma unroll
      for (int f = 0; f < kRvqF; ++f) {
        float dot = ze[f][0] * c0.x;
        dot = fmaf(ze[f][1], c0.y, dot); dot = fmaf(ze[f][2], c0.z, dot); dot = fmaf(ze[f][3], c0.w, dot);
        dot = fmaf(ze[f][4], c1.x, dot); dot = fmaf(ze[f][5], c1.y, dot); dot = fmaf(ze[f][6], c1.z, dot);
        dot = fmaf(ze[f][7], c1.w, dot);
        const float dist = (e2[f] + cs) - 2.0f * dot;
        if (dist < best[f]) { best[f] = dist; best_k[f] = k; }
      }
    }
#pragma unroll
    for (int f = 0; f < kRvqF; ++f) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best[f], o);
        const int ok = __shfl_xor_sync(0xffffffffu, best_k[f], o);
        if (od < best[f] || (od == best[f] && ok < best_k[f])) { best[f] = od; best_k[f] = ok; }   // lowest index wins
      }
    }
    // ---- lookup + straight-through arithmetic (VectorQuantizer.cs:81) + out_proj (:82)
    float q[kRvqF][kVqD];
#pragma unroll
    for (int f = 0; f < kRvqF; ++f) {
      const float4 q0 = cb_lo[best_k[f]], q1 = cb_hi[best_k[f]];
      const float qq[kVqD] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int d = 0; d < kVqD; ++d) q[f][d] = ze[f][d] + (qq[d] - ze[f][d]);
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      const float4 w0 = wo_lo[c], w1 = wo_hi[c];
      const float bo = out_b[c];
#pragma unroll
      for (int f = 0; f < kRvqF; ++f) {
        float v = w0.x * q[f][0];
        v = fmaf(w0.y, q[f][1], v); v = fmaf(w0.z, q[f][2], v); v = fmaf(w0.w, q[f][3], v);
        v = fmaf(w1.x, q[f][4], v); v = fmaf(w1.y, q[f][5], v); v = fmaf(w1.z, q[f][6], v); v = fmaf(w1.w, q[f][7], v);
        v += bo;
        acc[f][i] += v;   // zQ.add_(zQi)        ResidualVectorQuantizer.cs:68
        r[f][i] -= v;     // residual.sub_(zQi)  ResidualVectorQuantizer.cs:69
      }
    }
#pragma unroll
    for (int f = 0; f < kRvqF; ++f) {
      if (!valid[f]) continue;
      const int b = (int)(fidx[f] / T), t = (int)(fidx[f] % T);
      if (codes && lane == 0) codes[((long long)b * n_q + s) * T + t] = best_k[f];
      if (latents && lane < kVqD) {
        float mine = ze[f][0];
#pragma unroll
        for (int d = 1; d < kVqD; ++d) mine = lane == d ? ze[f][d] : mine;
        latents[((long long)b * n_q * kVqD + s * kVqD + lane) * T + t] = mine;
      }
      if (zq && s == n_q - 1) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) zq[fidx[f] * Dz + lane + 32 * i] = acc[f][i];
      }
    }
    __syncthreads();   // every warp is done with buffer n&1 before step n+1 prefetches into it
  }
}

template <int CPL>
static bool rvq_encode_block_launch(const RvqWeights& w, const float* z, float* zq, int64_t* codes, float* latents,
                                    int batch, int T, int n_q, const LaunchCtx& ctx) {
  const size_t buf_stride = ((size_t)w.blob_floats * 4 + 127) & ~(size_t)127;
  const size_t smem = 2 * buf_stride;
  if (smem > 226 * 1024) return false;
  auto k = rvq_encode_block_kernel<CPL>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const long long frames = (long long)batch * T;
  long long passes = (frames + kRvqWarps * kRvqF - 1) / (kRvqWarps * kRvqF);
  const int grid = (int)std::min<long long>(passes, ctx.num_sms);
  k<<<grid, kRvqWarps * 32, smem, ctx.stream>>>(w, z, zq, codes, latents, batch, T, n_q);
  return true;
}

template <int CPL>
static void rvq_encode_launch(const RvqWeights& w, const float* z, float* zq, int64_t* codes, float* latents,
                              int batch, int T, int n_q, const LaunchCtx& ctx) {
  const long long frames = (long long)batch * T;
  long long blocks = (frames + 7) / 8;
  const long long cap = (long long)ctx.num_sms * 8;
  if (blocks > cap) blocks = cap;
  rvq_encode_kernel<CPL><<<(unsigned)blocks, 256, 0, ctx.stream>>>(w, z, zq, codes, latents, batch, T, n_q);
}

void launch_rvq_encode(const RvqWeights& w, const float* z, float* zq, int64_t* codes, float* latents, int batch,
                       int T, int n_q, const LaunchCtx& ctx) {
  if (w.D != kVqD) throw Error(NC_UNSUPPORTED, "rvq: codebook_dim must be 8");
  if (w.Dz % 32 != 0 || w.K % 32 != 0) throw Error(NC_UNSUPPORTED, "rvq: latent dim / codebook size must be multiples of 32");
  if ((long long)batch * T == 0) return;
  const int ev = ctx.begin();
  static const int use_block = getenv("NC_RVQ_BLOCK") ? atoi(getenv("NC_RVQ_BLOCK")) : 1;
  bool done = false;
  if (use_block && w.blob && w.Dz % 128 == 0) {
    switch (w.Dz / 32) {
      case 4: done = rvq_encode_block_launch<4>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
      case 8: done = rvq_encode_block_launch<8>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
      case 16: done = rvq_encode_block_launch<16>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
      case 32: done = rvq_encode_block_launch<32>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
      default: break;
    }
  }
  if (!done) switch (w.Dz / 32) {
    case 1: rvq_encode_launch<1>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    case 2: rvq_encode_launch<2>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    case 4: rvq_encode_launch<4>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    case 8: rvq_encode_launch<8>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    case 16: rvq_encode_launch<16>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    case 24: rvq_encode_launch<24>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    case 32: rvq_encode_launch<32>(w, z, zq, codes, latents, batch, T, n_q, ctx); break;
    default: throw Error(NC_UNSUPPORTED, "rvq: unsupported latent dim " + std::to_string(w.Dz));
  }
  check_launch((int)cudaGetLastError(), "rvq_encode");
  const double fr = (double)batch * T;
  ctx.end(ev, "rvq_encode", fr * n_q * 2.0 * (2.0 * w.D * w.Dz + (double)w.D * w.K),
          fr * (2.0 * w.Dz * 4 + (codes ? n_q * 8.0 : 0) + (latents ? n_q * w.D * 4.0 : 0)));
}

// ------------------------------------------------------------------------------ codes -> latent
template <int CPL>
__global__ void __launch_bounds__(256)
rvq_from_codes_kernel(const RvqWeights w, const int64_t* __restrict__ codes, float* __restrict__ zq, int batch, int T,
                      int n_q) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long frames = (long long)batch * T;
  const int Dz = w.Dz, K = w.K;
  for (long long f = warp0; f < frames; f += nwarps) {
    const int b = (int)(f / T), t = (int)(f % T);
    float acc[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
    for (int s = 0; s < n_q; ++s) {
      long long code = codes[((long long)b * n_q + s) * T + t];
      code = code < 0 ? 0 : (code >= K ? K - 1 : code);
      const float* cb = w.cb + ((size_t)s * K + (size_t)code) * kVqD;
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(cb));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(cb) + 1);
      const float* Wout = w.out_w + (size_t)s * Dz * kVqD;
      const float* bout = w.out_b + (size_t)s * Dz;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int c = lane + 32 * i;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(Wout + (size_t)c * kVqD));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(Wout + (size_t)c * kVqD) + 1);
        float v = w0.x * q0.x;
        v = fmaf(w0.y, q0.y, v); v = fmaf(w0.z, q0.z, v); v = fmaf(w0.w, q0.w, v);
        v = fmaf(w1.x, q1.x, v); v = fmaf(w1.y, q1.y, v); v = fmaf(w1.z, q1.z, v); v = fmaf(w1.w, q1.w, v);
        v += __ldg(bout + c);
        acc[i] += v;
      }
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) zq[f * Dz + lane + 32 * i] = acc[i];
  }
}

template <int CPL>
static void rvq_from_codes_launch(const RvqWeights& w, const int64_t* codes, float* zq, int batch, int T, int n_q,
                                  const LaunchCtx& ctx) {
  const long long frames = (long long)batch * T;
  long long blocks = (frames + 7) / 8;
  const long long cap = (long long)ctx.num_sms * 8;
  if (blocks > cap) blocks = cap;
  rvq_from_codes_kernel<CPL><<<(unsigned)blocks, 256, 0, ctx.stream>>>(w, codes, zq, batch, T, n_q);
}

void launch_rvq_from_codes(const RvqWeights& w, const int64_t* codes, float* zq, int batch, int T, int n_q,
                           const LaunchCtx& ctx) {
  if (w.D != kVqD) throw Error(NC_UNSUPPORTED, "rvq: codebook_dim must be 8");
  if (w.Dz % 32 != 0) throw Error(NC_UNSUPPORTED, "rvq: latent dim must be a multiple of 32");
  if ((long long)batch * T == 0) return;
  const int ev = ctx.begin();
  switch (w.Dz / 32) {
    case 1: rvq_from_codes_launch<1>(w, codes, zq, batch, T, n_q, ctx); break;
    case 2: rvq_from_codes_launch<2>(w, codes, zq, batch, T, n_q, ctx); break;
    case 4: rvq_from_codes_launch<4>(w, codes, zq, batch, T, n_q, ctx); break;
    case 8: rvq_from_codes_launch<8>(w, codes, zq, batch, T, n_q, ctx); break;
    case 16: rvq_from_codes_launch<16>(w, codes, zq, batch, T, n_q, ctx); break;
    case 24: rvq_from_codes_launch<24>(w, codes, zq, batch, T, n_q, ctx); break;
    case 32: rvq_from_codes_launch<32>(w, codes, zq, batch, T, n_q, ctx); break;
    default: throw Error(NC_UNSUPPORTED, "rvq: unsupported latent dim " + std::to_string(w.Dz));
  }
  check_launch((int)cudaGetLastError(), "rvq_from_codes");
  const double fr = (double)batch * T;
  ctx.end(ev, "rvq_from_codes", fr * n_q * 2.0 * w.D * w.Dz, fr * (w.Dz * 4.0 + n_q * 8.0));
}

// ------------------------------------------------------------------------------ Dia hand-off
// Dia.GenerateOutput (Models/Dia.cs:1010-1044) + Decode's layout change (:973-981): undo the per-channel delay
// (gather at min(t + delay[c], T-1); the reference's pad mask can never fire because the index is clamped first),
// replace values outside [0, K) by 0, and write the DAC layout [n][C][len] for the selected items.
__global__ void dia_revert_kernel(const int64_t* __restrict__ gen, const int* __restrict__ items, const int* __restrict__ delay,
                                  int64_t* __restrict__ codes, int n_items, int T, int C, int len, int K) {
  const long long total = (long long)n_items * C * len;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % len);
    const int c = (int)((i / len) % C);
    const int n = (int)(i / ((long long)len * C));
    int ts = t + delay[c];
    if (ts > T - 1) ts = T - 1;
    long long v = gen[((long long)items[n] * T + ts) * C + c];
    if (v < 0 || v >= K) v = 0;
    codes[i] = v;
  }
}

void launch_dia_revert(const int64_t* gen, const int* items_dev, const int* delay_dev, int64_t* codes, int n_items, int T, int C,
                       int len, int K, const LaunchCtx& ctx) {
  const long long total = (long long)n_items * C * len;
  if (total == 0) return;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx.num_sms * 16);
  const int ev = ctx.begin();
  dia_revert_kernel<<<blocks, 256, 0, ctx.stream>>>(gen, items_dev, delay_dev, codes, n_items, T, C, len, K);
  check_launch((int)cudaGetLastError(), "dia_revert");
  ctx.end(ev, "dia_revert", 0, 16.0 * total);
}

// ------------------------------------------------------------------------------ input conditioning
__global__ void resample_linear_kernel(const float* __restrict__ in, long long n_in, long long in_stride, float* __restrict__ out,
                                       long long n_out, long long out_stride, double ratio) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const float* x = in + (long long)blockIdx.y * in_stride;
  // same double arithmetic as the reference loop, no fused multiply-add
  const double position = __ddiv_rn((double)i, ratio);
  const long long index = (long long)position;
  const double fraction = __dsub_rn(position, (double)index);
  float y;
  if (index >= n_in - 1) {
    y = x[n_in - 1];
  } else {
    const double a = __dmul_rn(__dsub_rn(1.0, fraction), (double)x[index]);
    const double b = __dmul_rn(fraction, (double)x[index + 1]);
    y = (float)__dadd_rn(a, b);
  }
  out[(long long)blockIdx.y * out_stride + i] = y;
}

void launch_resample_linear(const float* in, long long n_in, long long in_stride, float* out, long long n_out, long long out_stride,
                            double ratio, int batch, const LaunchCtx& ctx) {
  if (batch <= 0 || n_out <= 0) return;
  const int ev = ctx.begin();
  dim3 grid((unsigned)((n_out + 255) / 256), batch);
  resample_linear_kernel<<<grid, 256, 0, ctx.stream>>>(in, n_in, in_stride, out, n_out, out_stride, ratio);
  check_launch((int)cudaGetLastError(), "resample_linear");
  ctx.end(ev, "resample_linear", 0.0, 4.0 * batch * (double)(n_in + n_out));
}

__global__ void to_mono_kernel(const float* __restrict__ in, float* __restrict__ out, long long frames, int channels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames) return;
  float sum = 0.f;
  for (int ch = 0; ch < channels; ++ch) sum += in[i * channels + ch];
  out[i] = sum / (float)channels;
}

void launch_to_mono(const float* interleaved, float* out, long long frames, int channels, const LaunchCtx& ctx) {
  if (frames <= 0) return;
  const int ev = ctx.begin();
  to_mono_kernel<<<(unsigned)((frames + 255) / 256), 256, 0, ctx.stream>>>(interleaved, out, frames, channels);
  check_launch((int)cudaGetLastError(), "to_mono");
  ctx.end(ev, "to_mono", 0.0, 4.0 * (double)frames * (channels + 1));
}

}  // namespace nc
