// Encodec-specific kernels: reflect-padding fix-up of margin rows, residual VQ with 128-d Euclidean
// codebooks (fp32 register-tiled distance GEMM + running argmin), codes -> embedding sum, and the SEANet
// LSTM as a persistent cooperative kernel.
// Reference: Modules/Encodec/SConv1d.cs:144-173,252-274 (Pad1d), EuclideanCodebook.cs:155-182 (Quantize),
// ResidualVectorQuantizer.cs:107-157, SLSTM.cs:40-57.
#include "encodec_kernels.h"

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cooperative_groups.h>

namespace nc {

// ------------------------------------------------------------------------------ reflect padding into margins
// x points at row 0 of clip 0; rows -left..-1 and T..T+right-1 of every clip receive x[-j] = x[j],
// x[T-1+j] = x[T-1-j] (F.pad mode="reflect").
__global__ void reflect_pad_kernel(float* __restrict__ x, int T, int C, long long clip_stride, int left, int right, int batch) {
  const int c4n = C / 4;
  const int per_clip = (left + right) * c4n;
  const long long total = (long long)batch * per_clip;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_clip);
    const int rem = (int)(i - (long long)b * per_clip);
    const int j = rem / c4n, c4 = rem - j * c4n;
    int dst, src;
    if (j < left) { dst = -(j + 1); src = j + 1; } else { const int k = j - left + 1; dst = T - 1 + k; src = T - 1 - k; }
    float4* base = reinterpret_cast<float4*>(x + (long long)b * clip_stride);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0 && src < T) v = base[(long long)src * c4n + c4];
    base[(long long)dst * c4n + c4] = v;
  }
}

void launch_reflect_pad(float* x, int T, int C, long long clip_stride, int left, int right, int batch, const LaunchCtx& ctx) {
  if (left + right == 0 || batch == 0) return;
  if (C % 4 != 0) throw Error(NC_UNSUPPORTED, "reflect_pad: channel count must be a multiple of 4");
  if (T <= left || T <= right) throw Error(NC_UNSUPPORTED, "reflect_pad: clip shorter than its padding");
  const long long total = (long long)batch * (left + right) * (C / 4);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 1024);
  const int ev = ctx.begin();
  reflect_pad_kernel<<<blocks, 256, 0, ctx.stream>>>(x, T, C, clip_stride, left, right, batch);
  check_launch((int)cudaGetLastError(), "reflect_pad");
  ctx.end(ev, "reflect_pad", 0, 32.0 * total);
}

__global__ void pad1d_dense_kernel(const float* __restrict__ in, long long in_clip_stride, int T, int C, int extra_zero, int left,
                                   int Tp, float* __restrict__ out, long long total) {
  const int Te = T + extra_zero;   // length after the zero extension
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int t = (int)(r % Tp), b = (int)(r / Tp);
    int j = t - left;
    if (j < 0) j = -j;                       // F.pad(mode = "reflect") on the extended signal
    if (j >= Te) j = 2 * (Te - 1) - j;
    out[i] = (j >= 0 && j < T) ? in[(long long)b * in_clip_stride + (long long)j * C + c] : 0.f;
  }
}
void launch_pad1d_dense(const float* in, long long in_clip_stride, int T, int C, int extra_zero, int left, int right, float* out,
                        int batch, const LaunchCtx& ctx) {
  const int Tp = T + extra_zero + left + right;
  const long long total = (long long)batch * Tp * C;
  if (total == 0) return;
  if (left >= T + extra_zero || right >= T + extra_zero) throw Error(NC_INTERNAL, "pad1d: reflect padding longer than the extended input");
  const int blocks = (int)std::min<long long>((total + 255) / 256, 4096);
  const int ev = ctx.begin();
  pad1d_dense_kernel<<<blocks, 256, 0, ctx.stream>>>(in, in_clip_stride, T, C, extra_zero, left, Tp, out, total);
  check_launch((int)cudaGetLastError(), "pad1d_dense");
  ctx.end(ev, "pad1d_dense", 0, 8.0 * total);
}

// ------------------------------------------------------------------------------ VQ stage (D = 128)
// One block = 64 frames.  dist[f][k] = (|x_f|^2 + |e_k|^2) + (-2 * x_f.e_k), dot accumulated sequentially over d
// in fp32 (4x4 register tile per thread, x and e tiles staged in shared memory); running argmin with lowest-index
// tie-break; then residual[f] -= embed[argmin].
// The codebook chunk is staged ROW-MAJOR ([entry][d], 16-byte stores, conflict-free and coalesced); round 1 transposed it with
// scalar stores that hit 2 banks per warp (16-way conflicts), which cost as much as the FMAs.  The compute loop then reads its
// four entries with broadcast scalar loads (a warp holds only two distinct entry quads).
constexpr int kVqFrames = 64, kVqEntries = 64, kVqD = 128, kVqLd = kVqD + 4;

__global__ void __launch_bounds__(256)
encodec_vq_stage_kernel(float* __restrict__ residual, long long frames, const float* __restrict__ embed,
                        const float* __restrict__ embed_sq, int K, int64_t* __restrict__ codes, int T, int nq, int stage) {
  extern __shared__ __align__(16) float vq_smem[];
  float (*xs)[kVqFrames + 4] = reinterpret_cast<float (*)[kVqFrames + 4]>(vq_smem);                               // [d][frame]
  float (*es)[kVqLd] = reinterpret_cast<float (*)[kVqLd]>(vq_smem + kVqD * (kVqFrames + 4));                      // [entry][d]
  __shared__ float best_d[kVqFrames][16];
  __shared__ int best_k[kVqFrames][16];
  __shared__ int win[kVqFrames];
  const long long f0 = (long long)blockIdx.x * kVqFrames;
  const int tid = threadIdx.x;
  const int tf = tid & 15, te = tid >> 4;      // thread tile: frames 4*tf..+3, entries 4*te..+3
  for (int i = tid; i < kVqFrames * (kVqD / 4); i += 256) {
    const int fr = i / (kVqD / 4), d4 = i % (kVqD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f0 + fr < frames) v = *reinterpret_cast<const float4*>(residual + (f0 + fr) * kVqD + 4 * d4);
    xs[4 * d4 + 0][fr] = v.x; xs[4 * d4 + 1][fr] = v.y; xs[4 * d4 + 2][fr] = v.z; xs[4 * d4 + 3][fr] = v.w;
  }
  __syncthreads();
  // |x|^2 per frame: pow(2).sum(1) -- sequential over d
  float x2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
    for (int d = 0; d < kVqD; ++d) { const float v = xs[d][4 * tf + i]; s += v * v; }
    x2[i] = s;
  }
  float bd[4];
  int bk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { bd[i] = FLT_MAX; bk[i] = 0; }
  for (int k0 = 0; k0 < K; k0 += kVqEntries) {
    __syncthreads();
    for (int i = tid; i < kVqEntries * (kVqD / 4); i += 256) {
      const int en = i / (kVqD / 4), d4 = i % (kVqD / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + en < K) v = __ldg(reinterpret_cast<const float4*>(embed + (size_t)(k0 + en) * kVqD) + d4);
      *reinterpret_cast<float4*>(&es[en][4 * d4]) = v;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* e0 = es[4 * te + 0];
    const float* e1 = es[4 * te + 1];
    const float* e2r = es[4 * te + 2];
    const float* e3 = es[4 * te + 3];
#pragma unroll 8
    for (int d = 0; d < kVqD; ++d) {
      const float4 xv = *reinterpret_cast<const float4*>(&xs[d][4 * tf]);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ea[4] = {e0[d], e1[d], e2r[d], e3[d]};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ea[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * te + j;
      if (k >= K) continue;
      const float e2 = __ldg(embed_sq + k);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dist = (x2[i] + e2) + (-2.0f * acc[i][j]);
        if (dist < bd[i]) { bd[i] = dist; bk[i] = k; }   // k ascending within a thread: first minimum wins
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { best_d[4 * tf + i][te] = bd[i]; best_k[4 * tf + i][te] = bk[i]; }
  __syncthreads();
  if (tid < kVqFrames) {
    float d = best_d[tid][0];
    int k = best_k[tid][0];
    for (int j = 1; j < 16; ++j) {
      const float dj = best_d[tid][j];
      const int kj = best_k[tid][j];
      if (dj < d || (dj == d && kj < k)) { d = dj; k = kj; }
    }
    win[tid] = k;
    const long long f = f0 + tid;
    if (f < frames && codes) codes[((f / T) * nq + stage) * T + (f % T)] = k;
  }
  __syncthreads();
  // residual -= quantized
  for (int i = tid; i < kVqFrames * (kVqD / 4); i += 256) {
    const int fr = i / (kVqD / 4), d4 = i % (kVqD / 4);
    if (f0 + fr >= frames) continue;
    const float4 q = __ldg(reinterpret_cast<const float4*>(embed + (size_t)win[fr] * kVqD) + d4);
    float4* r = reinterpret_cast<float4*>(residual + (f0 + fr) * kVqD) + d4;
    float4 v = *r;
    v.x -= q.x; v.y -= q.y; v.z -= q.z; v.w -= q.w;
    *r = v;
  }
}

void launch_encodec_vq_stage(float* residual, long long frames, const float* embed, const float* embed_sq, int K, int D,
                             int64_t* codes, int T, int nq, int stage, const LaunchCtx& ctx) {
  if (D != kVqD) throw Error(NC_UNSUPPORTED, "encodec vq: codebook dimension must be 128");
  if (frames == 0) return;
  const unsigned blocks = (unsigned)((frames + kVqFrames - 1) / kVqFrames);
  const size_t smem = ((size_t)kVqD * (kVqFrames + 4) + (size_t)kVqEntries * kVqLd) * sizeof(float);
  cudaFuncSetAttribute(encodec_vq_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int ev = ctx.begin();
  encodec_vq_stage_kernel<<<blocks, 256, smem, ctx.stream>>>(residual, frames, embed, embed_sq, K, codes, T, nq, stage);
  check_launch((int)cudaGetLastError(), "encodec_vq_stage");
  ctx.end(ev, "encodec_vq_stage", 2.0 * frames * K * D, (double)frames * D * 8.0);
}

// out[b, t, :] = sum_i embed_i[codes[b, i, t]]   (ResidualVectorQuantizer.Decode); out rows may be strided per clip
__global__ void encodec_decode_codes_kernel(const int64_t* __restrict__ codes, const float* const* __restrict__ embeds,
                                            float* __restrict__ out, long long out_clip_stride, int batch, int T, int nq,
                                            int K) {
  const long long total = (long long)batch * T * (kVqD / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d4 = (int)(i % (kVqD / 4));
    const long long f = i / (kVqD / 4);
    const int b = (int)(f / T), t = (int)(f % T);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nq; ++s) {
      long long c = codes[((long long)b * nq + s) * T + t];
      c = c < 0 ? 0 : (c >= K ? K - 1 : c);
      const float4 e = __ldg(reinterpret_cast<const float4*>(embeds[s] + (size_t)c * kVqD) + d4);
      a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
    }
    reinterpret_cast<float4*>(out + (long long)b * out_clip_stride + (long long)t * kVqD)[d4] = a;
  }
}

void launch_encodec_decode_codes(const int64_t* codes, const float* const* embeds_dev, float* out, long long out_clip_stride,
                                 int batch, int T, int nq, int K, int D, const LaunchCtx& ctx) {
  if (D != kVqD) throw Error(NC_UNSUPPORTED, "encodec vq: codebook dimension must be 128");
  const long long total = (long long)batch * T * (kVqD / 4);
  if (total == 0) return;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx.num_sms * 16);
  const int ev = ctx.begin();
  encodec_decode_codes_kernel<<<blocks, 256, 0, ctx.stream>>>(codes, embeds_dev, out, out_clip_stride, batch, T, nq, K);
  check_launch((int)cudaGetLastError(), "encodec_decode_codes");
  ctx.end(ev, "encodec_decode_codes", 0, (double)batch * T * (D * 4.0 + nq * 8.0));
}

// ------------------------------------------------------------------------------ LSTM layer (persistent)
// One launch = one layer over all T steps.  Grid: (H/16 unit slices) x (batch slices of 16); all CTAs are
// co-resident (cooperative launch).  A CTA owns the 64 gate rows (i,f,g,o x 16 units) of W_hh for its units, held
// in REGISTERS (thread = 4 gate rows x 32-wide K slice), and the cell state of its 16 units x 16 clips.
// Per step: stage h_{t-1} of the 16 clips in shared memory, partial dots per K slice, shuffle-reduce over the 16
// K slices, add the hoisted input projection (W_ih x_t + b_ih + b_hh, computed by one tensor-core GEMM),
// gate math in fp32, write h_t; CTAs of the same batch slice meet at a global-memory barrier.
constexpr int kLU = 16;   // units per CTA
constexpr int kLB = 16;   // clips per CTA

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256, 1)
lstm_layer_kernel(const float* __restrict__ xproj, long long xproj_clip_stride,   // [B][T][4H] (+clip stride)
                  const float* __restrict__ w_hh,                                  // [4H][H]
                  float* __restrict__ hbuf,                                        // [2][Bpad][H] ping-pong, zero-initialised
                  float* __restrict__ out, long long out_clip_stride,              // [B][T][H]
                  const float* __restrict__ skip, long long skip_clip_stride,      // nullable: out = h + skip (SLSTM skip)
                  int post_elu, unsigned int* __restrict__ barriers,               // [batch slices], zero-initialised
                  int batch, int T, int H) {
  __shared__ __align__(16) float hs[kLB][16 * 36];   // [clip][K slice][32 + 4 pad]: conflict-free 128-bit reads
  __shared__ float gates[4 * kLU][kLB + 1];
  const int tid = threadIdx.x;
  const int ks = tid & 15;          // K slice: columns 32*ks .. +31
  const int rg = tid >> 4;          // row group: gate rows 4*rg .. +3 of this CTA's 64
  const int u0 = blockIdx.x * kLU, b0 = blockIdx.y * kLB;
  const int n_unit_ctas = gridDim.x;
  const int bpad = gridDim.y * kLB;
  // local gate row r (0..63): gate = r / 16, unit = r % 16  -> global row gate*H + u0 + unit
  float w[4][32];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 4 * rg + i;
    const float* wr = w_hh + ((size_t)(r / kLU) * H + u0 + (r % kLU)) * H + 32 * ks;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(wr) + k4);
      w[i][4 * k4] = v.x; w[i][4 * k4 + 1] = v.y; w[i][4 * k4 + 2] = v.z; w[i][4 * k4 + 3] = v.w;
    }
  }
  // cell state: thread tid owns (unit = tid / 16, clip = tid % 16)
  const int my_u = tid >> 4, my_b = tid & 15;
  float c_state = 0.f;
  const bool clip_ok = b0 + my_b < batch;
  // The hoisted input projection and the skip value of a step do not depend on other CTAs: their HBM loads are
  // issued one step ahead, between arriving at the grid barrier and waiting on it.
  float nxi = 0.f, nxf = 0.f, nxg = 0.f, nxo = 0.f, nsk = 0.f;
  auto prefetch = [&](int t) {
    if (clip_ok && t < T) {
      const float* xp = xproj + (long long)(b0 + my_b) * xproj_clip_stride + (long long)t * 4 * H + (u0 + my_u);
      nxi = __ldg(xp); nxf = __ldg(xp + H); nxg = __ldg(xp + 2 * H); nxo = __ldg(xp + 3 * H);
      if (skip) nsk = __ldg(skip + (long long)(b0 + my_b) * skip_clip_stride + (long long)t * H + (u0 + my_u));
    }
  };
  prefetch(0);
  for (int t = 0; t < T; ++t) {
    const float* hprev = hbuf + (size_t)((t + 1) & 1) * bpad * H;   // written at step t-1 (zeros at t = 0)
    float* hnext = hbuf + (size_t)(t & 1) * bpad * H;
    const float xi = nxi, xf = nxf, xg = nxg, xo = nxo, sk = nsk;
    // H == 512 (checked at launch): 128 float4 per clip, 8 fully unrolled loads per thread issued back to back
#pragma unroll
    for (int i = tid; i < kLB * 128; i += 256) {
      const int bb = i >> 7, k4 = i & 127;
      // written by other CTAs during this launch: bypass L1 (ld.global.cg)
      const float4 v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)(b0 + bb) * H) + k4);
      *reinterpret_cast<float4*>(&hs[bb][(k4 >> 3) * 36 + 4 * (k4 & 7)]) = v;
    }
    __syncthreads();
    for (int bg = 0; bg < 4; ++bg) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        float4 hv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) hv[j] = *reinterpret_cast<const float4*>(&hs[4 * bg + j][36 * ks + 4 * k4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i][j] = fmaf(w[i][4 * k4 + 0], hv[j].x, acc[i][j]);
            acc[i][j] = fmaf(w[i][4 * k4 + 1], hv[j].y, acc[i][j]);
            acc[i][j] = fmaf(w[i][4 * k4 + 2], hv[j].z, acc[i][j]);
            acc[i][j] = fmaf(w[i][4 * k4 + 3], hv[j].w, acc[i][j]);
          }
      }
      // reduce-scatter over the 16 K slices (lanes ks = tid & 15 of each half warp): a butterfly that halves the
      // number of live values per lane at every step (8 + 4 + 2 + 1 = 15 shuffles instead of 16 x 4); lane ks ends
      // up with the complete sum of accumulator (i, j) = (ks >> 2, ks & 3).
      float v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[4 * i + j] = acc[i][j];
#pragma unroll
      for (int m = 8; m >= 1; m >>= 1) {
        const bool up = (ks & m) != 0;
#pragma unroll
        for (int q = 0; q < m; ++q) {
          const float send = up ? v[q] : v[q + m];
          const float keep = up ? v[q + m] : v[q];
          v[q] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      }
      gates[4 * rg + (ks >> 2)][4 * bg + (ks & 3)] = v[0];
    }
    __syncthreads();
    if (clip_ok) {
      const int b = b0 + my_b, u = u0 + my_u;
      const float gi = gates[0 * kLU + my_u][my_b] + xi;
      const float gf = gates[1 * kLU + my_u][my_b] + xf;
      const float gg = gates[2 * kLU + my_u][my_b] + xg;
      const float go = gates[3 * kLU + my_u][my_b] + xo;
      c_state = sigmoid_f(gf) * c_state + sigmoid_f(gi) * tanhf(gg);
      const float h = sigmoid_f(go) * tanhf(c_state);
      __stcg(hnext + (size_t)b * H + u, h);
      float y = h;
      if (skip) y += sk;
      if (post_elu) y = y > 0.f ? y : expm1f(y);
      out[(long long)b * out_clip_stride + (long long)t * H + u] = y;
    }
    // barrier among the CTAs of this batch slice.  One device-scope fence by the signalling thread AFTER the CTA barrier
    // (fences are cumulative over what the barrier ordered before them) instead of one MEMBAR.GPU per thread before it.
    __syncthreads();
    if (tid == 0)   // release at device scope: cumulative over the h_t stores the CTA barrier ordered before it
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(barriers + blockIdx.y), "r"(1u) : "memory");
    prefetch(t + 1);
    if (tid == 0) {
      const unsigned int target = (unsigned int)n_unit_ctas * (unsigned int)(t + 1);
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(barriers + blockIdx.y) : "memory");
      } while (seen < target);
    }
    __syncthreads();
  }
}

// Tried in round 2 and dropped (profiles/r02_encodec_lstm_two_slices.txt): a CTA owning 8 units of TWO batch slices and alternating
// between them, to hide one slice's exchange + barrier behind the other's FMAs.  Bit-identical results but 28.9 ms instead of
// 20.3 ms per 64 x 10 s: the exchange latency (~4.6 us: device-scope fence, L2 atomic, polling) is four times one slice's FMA
// phase, so two slices cannot cover it and every half step still waits for a barrier signalled one half step earlier.

void launch_lstm_layer(const float* xproj, long long xproj_clip_stride, const float* w_hh, float* hbuf, float* out,
                       long long out_clip_stride, const float* skip, long long skip_clip_stride, int post_elu,
                       unsigned int* barriers, int batch, int T, int H, const LaunchCtx& ctx) {
  if (H != 512) throw Error(NC_UNSUPPORTED, "lstm: hidden size must be 512");
  if (batch == 0 || T == 0) return;
  const int by = (batch + kLB - 1) / kLB;
  dim3 grid(H / kLU, by);
  if ((int)(grid.x * grid.y) > ctx.num_sms) throw Error(NC_INTERNAL, "lstm: batch slice too large for a co-resident grid");
  NC_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)2 * by * kLB * H * sizeof(float), ctx.stream));
  NC_CUDA(cudaMemsetAsync(barriers, 0, (size_t)by * sizeof(unsigned int), ctx.stream));
  const int ev = ctx.begin();
  void* args[] = {(void*)&xproj, (void*)&xproj_clip_stride, (void*)&w_hh, (void*)&hbuf, (void*)&out, (void*)&out_clip_stride,
                  (void*)&skip, (void*)&skip_clip_stride, (void*)&post_elu, (void*)&barriers, (void*)&batch, (void*)&T, (void*)&H};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)lstm_layer_kernel, grid, dim3(256), args, 0, ctx.stream);
  check_launch((int)e, "lstm_layer");
  ctx.end(ev, "lstm_layer", 2.0 * 4 * H * (double)H * T * batch, (double)batch * T * H * 4.0 * 6);
}

// ------------------------------------------------------------------ .ecdc bit packing (BitPacker.cs / BitUnpacker.cs)
// The reference pushes code (t, k) -- t outer, k inner -- into an LSB-first accumulator of `bits` bits per value and
// emits the low byte whenever 8 bits are available; Flush() emits the last partial byte.  Equivalent closed form:
// value v = t*nq + k occupies stream bits [v*bits, (v+1)*bits); byte j is bits [8j, 8j+8).  One thread per output byte.
__global__ void ecdc_pack_kernel(const int64_t* __restrict__ codes, uint8_t* __restrict__ out, long long out_stride, int T, int nq,
                                 int bits, long long nbytes, long long row_stride, long long clip_stride) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= nbytes) return;
  const long long nvals = (long long)T * nq;
  const long long bit0 = j * 8;
  long long v = bit0 / bits;
  const long long v_last = min((bit0 + 7) / bits, nvals - 1);
  const int64_t* cb = codes + (long long)b * clip_stride;
  unsigned int byte = 0;
  for (; v <= v_last; ++v) {
    const int t = (int)(v / nq), k = (int)(v % nq);
    const unsigned long long val = (unsigned long long)cb[(long long)k * row_stride + t] & ((1ull << bits) - 1ull);
    const long long sh = v * bits - bit0;   // position of the value's bit 0 relative to this byte
    byte |= (unsigned int)((sh >= 0 ? (val << sh) : (val >> (-sh))) & 0xFFull);
  }
  out[(long long)b * out_stride + j] = (uint8_t)byte;
}

// One thread per value: gather the (at most 5) bytes it straddles, shift, mask (BitUnpacker.Pull, BitUnpacker.cs:60-95).
__global__ void ecdc_unpack_kernel(const uint8_t* __restrict__ in, long long in_stride, int64_t* __restrict__ codes, int T, int nq,
                                   int bits, long long nbytes, long long row_stride, long long clip_stride) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= (long long)T * nq) return;
  const long long bit0 = v * bits;
  const long long j0 = bit0 >> 3;
  const int off = (int)(bit0 & 7);
  const uint8_t* src = in + (long long)b * in_stride;
  unsigned long long w = 0;
  const int need = (off + bits + 7) >> 3;   // <= 5 for bits <= 32
  for (int i = 0; i < need; ++i)
    if (j0 + i < nbytes) w |= (unsigned long long)src[j0 + i] << (8 * i);
  const int t = (int)(v / nq), k = (int)(v % nq);
  codes[(long long)b * clip_stride + (long long)k * row_stride + t] = (int64_t)((w >> off) & ((1ull << bits) - 1ull));
}

void launch_ecdc_pack(const int64_t* codes, uint8_t* out, long long out_stride, int batch, int T, int nq, int bits,
                      const LaunchCtx& ctx, long long row_stride, long long clip_stride) {
  if (row_stride == 0) row_stride = T;
  if (clip_stride == 0) clip_stride = (long long)nq * T;
  if (bits <= 0 || bits > 32) throw Error(NC_INVALID_ARGUMENT, "Bits must be between 1 and 32");   // BitPacker.cs:120-130
  const long long nbytes = ((long long)T * nq * bits + 7) / 8;
  if (batch == 0 || nbytes == 0) return;
  const int ev = ctx.begin();
  dim3 grid((unsigned)((nbytes + 255) / 256), batch);
  ecdc_pack_kernel<<<grid, 256, 0, ctx.stream>>>(codes, out, out_stride, T, nq, bits, nbytes, row_stride, clip_stride);
  check_launch((int)cudaGetLastError(), "ecdc_pack");
  ctx.end(ev, "ecdc_pack", 0.0, (double)batch * ((double)T * nq * 8 + nbytes));
}

void launch_ecdc_unpack(const uint8_t* in, long long in_stride, int64_t* codes, int batch, int T, int nq, int bits,
                        const LaunchCtx& ctx, long long row_stride, long long clip_stride) {
  if (row_stride == 0) row_stride = T;
  if (clip_stride == 0) clip_stride = (long long)nq * T;
  if (bits <= 0 || bits > 32) throw Error(NC_INVALID_ARGUMENT, "Bits must be between 1 and 32");
  const long long nvals = (long long)T * nq;
  if (batch == 0 || nvals == 0) return;
  const long long nbytes = (nvals * bits + 7) / 8;
  const int ev = ctx.begin();
  dim3 grid((unsigned)((nvals + 255) / 256), batch);
  ecdc_unpack_kernel<<<grid, 256, 0, ctx.stream>>>(in, in_stride, codes, T, nq, bits, nbytes, row_stride, clip_stride);
  check_launch((int)cudaGetLastError(), "ecdc_unpack");
  ctx.end(ev, "ecdc_unpack", 0.0, (double)batch * ((double)nvals * 8 + nbytes));
}

// ------------------------------------------------------------------------------ 48 kHz preset: GroupNorm(1, C), segments
// time_group_norm (NormConv1d.cs:136-160): one mean / variance per clip over (C, T).  Pass 1: per-thread fp32 partials over
// at most kGnPerThread elements, block reduction and the cross-block sum in fp64.
constexpr int kGnChunk = 16384;   // floats per block
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, long long clip_stride, long long n, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const float4* p = reinterpret_cast<const float4*>(x + (long long)b * clip_stride);
  const long long n4 = n >> 2;
  const long long i0 = (long long)blockIdx.x * (kGnChunk / 4), i1 = min(i0 + kGnChunk / 4, n4);
  float s = 0.f, ss = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float4 v = __ldg(p + i);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  double ds = s, dss = ss;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dss += __shfl_xor_sync(0xffffffffu, dss, o);
  }
  __shared__ double sh[2][8];
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ds; sh[1][threadIdx.x >> 5] = dss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; c += sh[1][i]; }
    atomicAdd(stats + 2 * b, a);
    atomicAdd(stats + 2 * b + 1, c);
  }
}

void launch_gn_stats(const float* x, long long clip_stride, long long n_per_clip, double* stats, int batch, const LaunchCtx& ctx) {
  if (batch == 0 || n_per_clip == 0) return;
  if (n_per_clip % 4 != 0) throw Error(NC_INTERNAL, "gn_stats: row width must be a multiple of 4");
  NC_CUDA(cudaMemsetAsync(stats, 0, (size_t)batch * 2 * sizeof(double), ctx.stream));
  const int ev = ctx.begin();
  dim3 grid((unsigned)((n_per_clip + kGnChunk - 1) / kGnChunk), batch);
  gn_stats_kernel<<<grid, 256, 0, ctx.stream>>>(x, clip_stride, n_per_clip, stats);
  check_launch((int)cudaGetLastError(), "gn_stats");
  ctx.end(ev, "gn_stats", 0.0, 4.0 * batch * (double)n_per_clip);
}

__device__ __forceinline__ void gn_coeffs(const double* stats, int b, double count, float eps, float& mean, float& rstd) {
  const double m = stats[2 * b] / count;
  const double var = fmax(stats[2 * b + 1] / count - m * m, 0.0);
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}
__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }   // HBM-bound pass: the precise form is free here

// Pass 2, in place over rows [0, T): y = (x - mean) * rstd * gamma[c] + beta[c] (+ residual) (-> ELU).  The residual is either
// final values or -- stats2 != nullptr -- another conv's RAW output normalised on the fly with its own statistics and affine
// (the resnet shortcut: saves that tensor's own apply pass).
__global__ void __launch_bounds__(256)
gn_apply_kernel(float* __restrict__ y, long long clip_stride, int T, int C, const double* __restrict__ stats, double count, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ residual,
                long long res_clip_stride, int elu, const double* __restrict__ stats2, double count2, const float* __restrict__ gamma2,
                const float* __restrict__ beta2) {
  const int b = blockIdx.y, c4n = C >> 2;
  float mean, rstd, mean2 = 0.f, rstd2 = 1.f;
  gn_coeffs(stats, b, count, eps, mean, rstd);
  if (stats2) gn_coeffs(stats2, b, count2, eps, mean2, rstd2);
  float4* p = reinterpret_cast<float4*>(y + (long long)b * clip_stride);
  const float4* r = residual ? reinterpret_cast<const float4*>(residual + (long long)b * res_clip_stride) : nullptr;
  const long long n4 = (long long)T * c4n;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const int c4 = (int)(i % c4n);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 v = p[i];
    v.x = (v.x - mean) * rstd * g.x + be.x; v.y = (v.y - mean) * rstd * g.y + be.y;
    v.z = (v.z - mean) * rstd * g.z + be.z; v.w = (v.w - mean) * rstd * g.w + be.w;
    if (r) {
      float4 q = __ldg(r + i);
      if (stats2) {
        const float4 g2 = __ldg(reinterpret_cast<const float4*>(gamma2) + c4), b2 = __ldg(reinterpret_cast<const float4*>(beta2) + c4);
        q.x = (q.x - mean2) * rstd2 * g2.x + b2.x; q.y = (q.y - mean2) * rstd2 * g2.y + b2.y;
        q.z = (q.z - mean2) * rstd2 * g2.z + b2.z; q.w = (q.w - mean2) * rstd2 * g2.w + b2.w;
      }
      v.x = q.x + v.x; v.y = q.y + v.y; v.z = q.z + v.z; v.w = q.w + v.w;    // shortcut + block, the reference's operand order
    }
    if (elu) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
    p[i] = v;
  }
}

void launch_gn_apply(float* y, long long clip_stride, int T, int C, const double* stats, double count, float eps, const float* gamma,
                     const float* beta, const float* residual, long long res_clip_stride, int elu, int batch, const LaunchCtx& ctx,
                     const double* stats2, double count2, const float* gamma2, const float* beta2) {
  if (batch == 0 || T == 0) return;
  if (C % 4 != 0) throw Error(NC_INTERNAL, "gn_apply: channel count must be a multiple of 4");
  const long long n4 = (long long)T * (C / 4);
  const int ev = ctx.begin();
  dim3 grid((unsigned)std::min<long long>((n4 + 1023) / 1024, 8LL * ctx.num_sms), batch);
  gn_apply_kernel<<<grid, 256, 0, ctx.stream>>>(y, clip_stride, T, C, stats, count, eps, gamma, beta, residual, res_clip_stride, elu,
                                                stats2, count2, gamma2, beta2);
  check_launch((int)cudaGetLastError(), "gn_apply");
  ctx.end(ev, "gn_apply", 0.0, (residual ? 12.0 : 8.0) * batch * (double)T * C);
}

// Segment item i of a launch = global item item0 + i -> clip b = item / segs, segment s = s0 + item % segs, samples
// [s * seg_stride, s * seg_stride + seg_len) of audio [B][C][L].
// Loudness scale (Encodec.cs:469-480): sqrt(mean_t (mean_c x)^2) + 1e-8, one block per item, fp64 accumulation.
__global__ void __launch_bounds__(256)
segment_scale_kernel(const float* __restrict__ audio, int C, long long L, int segs, int s0, long long seg_stride, int seg_len,
                     int item0, int n_seg_total, float* __restrict__ scales) {
  const int item = item0 + blockIdx.x, b = item / segs, s = s0 + item % segs;
  const float* x = audio + (long long)b * C * L + (long long)s * seg_stride;
  double acc = 0;
  for (int t = threadIdx.x; t < seg_len; t += 256) {
    float m = 0.f;
    for (int c = 0; c < C; ++c) m += x[(long long)c * L + t];
    m /= (float)C;
    acc += (double)(m * m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += sh[i];
    scales[(long long)b * n_seg_total + s] = sqrtf((float)(a / seg_len)) + 1e-8f;
  }
}

// planar [B][C][L] segment -> channels-last rows [seg_len][Cpad] (channels >= C zero), divided by the item's scale
__global__ void __launch_bounds__(256)
segment_prep_kernel(const float* __restrict__ audio, int C, long long L, int segs, int s0, long long seg_stride, int seg_len, int item0,
                    int n_seg_total, const float* __restrict__ scales, float* __restrict__ out, long long out_clip_stride, int Cpad) {
  const int i = blockIdx.y, item = item0 + i, b = item / segs, s = s0 + item % segs;
  const float* x = audio + (long long)b * C * L + (long long)s * seg_stride;
  const float sc = scales ? scales[(long long)b * n_seg_total + s] : 1.f;
  float4* o = reinterpret_cast<float4*>(out + (long long)i * out_clip_stride);
  const int c4n = Cpad >> 2;                       // C <= 4 real channels live in the first float4 of a row
  const long long n = (long long)seg_len * c4n;
  for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < n; j += (long long)gridDim.x * 256) {
    const int c4 = (int)(j % c4n);
    const long long t = j / c4n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 == 0) {
      v.x = x[t];
      if (C > 1) v.y = x[L + t];
      if (scales) { v.x = v.x / sc; v.y = v.y / sc; }
    }
    o[j] = v;
  }
}

void launch_encodec_segment_prep(const float* audio, int C, long long L, int segs, int s0, long long seg_stride, int seg_len, int item0,
                                 int n_seg_total, float* scales, float* out, long long out_clip_stride, int Cpad, int items,
                                 const LaunchCtx& ctx) {
  if (items == 0 || seg_len == 0) return;
  const int ev = ctx.begin();
  if (scales) {
    segment_scale_kernel<<<items, 256, 0, ctx.stream>>>(audio, C, L, segs, s0, seg_stride, seg_len, item0, n_seg_total, scales);
    check_launch((int)cudaGetLastError(), "segment_scale");
  }
  if (C > 2 || Cpad % 4 != 0) throw Error(NC_INTERNAL, "segment_prep: at most 2 channels");
  const long long n = (long long)seg_len * (Cpad / 4);
  dim3 grid((unsigned)std::min<long long>((n + 1023) / 1024, 4LL * ctx.num_sms), items);
  segment_prep_kernel<<<grid, 256, 0, ctx.stream>>>(audio, C, L, segs, s0, seg_stride, seg_len, item0, n_seg_total, scales, out,
                                                    out_clip_stride, Cpad);
  check_launch((int)cudaGetLastError(), "segment_prep");
  if (ctx.launches && scales) ++*ctx.launches;
  ctx.end(ev, "segment_prep", 0.0, 4.0 * items * (double)seg_len * (C * (scales ? 2 : 1) + Cpad));
}

// last decoder conv's raw rows [T][Cpad] -> (GroupNorm over the C real channels) * scale -> planar frame [C][frame_ld]
__global__ void __launch_bounds__(256)
frame_out_kernel(const float* __restrict__ raw, long long clip_stride, int T, int Cpad, int C, const double* __restrict__ stats, float eps,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int segs, int s0, int item0, int n_seg_total,
                 const float* __restrict__ scales, float* __restrict__ frames, long long frame_ld) {
  const int i = blockIdx.y, item = item0 + i, b = item / segs, s = s0 + item % segs;
  float mean = 0.f, rstd = 1.f;
  if (stats) gn_coeffs(stats, i, (double)T * C, eps, mean, rstd);
  const float sc = scales ? scales[(long long)b * n_seg_total + s] : 1.f;
  const float* x = raw + (long long)i * clip_stride;
  float* o = frames + ((long long)b * n_seg_total + s) * C * frame_ld;
  for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < (long long)T * C; j += (long long)gridDim.x * 256) {
    const int c = (int)(j / T);
    const int t = (int)(j - (long long)c * T);
    float v = x[(long long)t * Cpad + c];
    if (stats) v = (v - mean) * rstd * gamma[c] + beta[c];
    if (scales) v = v * sc;
    o[(long long)c * frame_ld + t] = v;
  }
}

void launch_encodec_frame_out(const float* raw, long long clip_stride, int T, int Cpad, int C, const double* stats, float eps,
                              const float* gamma, const float* beta, int segs, int s0, int item0, int n_seg_total, const float* scales,
                              float* frames, long long frame_ld, int items, const LaunchCtx& ctx) {
  if (items == 0 || T == 0) return;
  const long long n = (long long)T * C;
  const int ev = ctx.begin();
  dim3 grid((unsigned)std::min<long long>((n + 1023) / 1024, 4LL * ctx.num_sms), items);
  frame_out_kernel<<<grid, 256, 0, ctx.stream>>>(raw, clip_stride, T, Cpad, C, stats, eps, gamma, beta, segs, s0, item0, n_seg_total,
                                                 scales, frames, frame_ld);
  check_launch((int)cudaGetLastError(), "frame_out");
  ctx.end(ev, "frame_out", 0.0, 4.0 * items * (double)T * (Cpad + C));
}

// DSP.LinearOverlapAdd (AudioTensorDSP.cs:161-261): out[t] = (sum_s w[t - s*stride] * frame_s[t - s*stride]) / sum_s w[..],
// w = 0.5 - |linspace(0, 1, len0 + 2)[1:-1] - 0.5| (torch's two-sided linspace), frames summed in ascending order from 0.
// frames [B][n_seg][C][frame_ld]; frame s has lens[s] samples (device array).  segmented == 0: plain copy.
__device__ __forceinline__ float ola_weight(int j, int len0) {
  const int steps = len0 + 2, idx = j + 1;
  const float step = 1.0f / (float)(steps - 1);
  const float t = idx < steps / 2 ? step * (float)idx : 1.0f - step * (float)(steps - idx - 1);
  return 0.5f - fabsf(t - 0.5f);
}
__global__ void __launch_bounds__(256)
overlap_add_kernel(const float* __restrict__ frames, int n_seg, int C, long long frame_ld, const int* __restrict__ lens, int len_max,
                   long long stride, int segmented, float* __restrict__ out, long long out_len) {
  const int bc = blockIdx.y, b = bc / C, c = bc % C;
  const int len0 = lens[0];
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < out_len; t += (long long)gridDim.x * 256) {
    float v;
    if (!segmented) {
      v = frames[((long long)b * n_seg * C + c) * frame_ld + t];
    } else {
      long long s_lo = t - (len_max - 1);
      s_lo = s_lo <= 0 ? 0 : (s_lo + stride - 1) / stride;
      const long long s_hi = min((long long)n_seg - 1, t / stride);
      float acc = 0.f, wsum = 0.f;
      for (long long s = s_lo; s <= s_hi; ++s) {
        const int j = (int)(t - s * stride);
        if (j >= lens[s]) continue;
        const float w = ola_weight(j, len0);
        acc += frames[(((long long)b * n_seg + s) * C + c) * frame_ld + j] * w;
        wsum += w;
      }
      v = acc / wsum;
    }
    out[(long long)bc * out_len + t] = v;
  }
}

void launch_encodec_overlap_add(const float* frames, int batch, int n_seg, int C, long long frame_ld, const int* lens_dev, int len_max,
                                long long stride, int segmented, float* out, long long out_len, const LaunchCtx& ctx) {
  if (batch == 0 || out_len == 0) return;
  const int ev = ctx.begin();
  dim3 grid((unsigned)std::min<long long>((out_len + 1023) / 1024, 8LL * ctx.num_sms), batch * C);
  overlap_add_kernel<<<grid, 256, 0, ctx.stream>>>(frames, n_seg, C, frame_ld, lens_dev, len_max, stride, segmented, out, out_len);
  check_launch((int)cudaGetLastError(), "overlap_add");
  ctx.end(ev, "overlap_add", 0.0, 8.0 * batch * C * (double)out_len);
}

// codes of `items` segment items (item = b*segs + j), dense [items][nq][T]  <->  the caller's [B][nq][T_total] with the
// group's j-th segment at column col0 + j*T
__global__ void __launch_bounds__(256)
codes_segment_copy_kernel(int64_t* __restrict__ dense, int64_t* __restrict__ user, int segs, int item0, int nq, int T,
                          long long T_total, long long col0, int to_user, long long n) {
  for (long long j = (long long)blockIdx.x * 256 + threadIdx.x; j < n; j += (long long)gridDim.x * 256) {
    const int t = (int)(j % T);
    const long long r = j / T;
    const int q = (int)(r % nq), i = (int)(r / nq);
    const int item = item0 + i, b = item / segs, j_seg = item % segs;
    const long long u = ((long long)b * nq + q) * T_total + col0 + (long long)j_seg * T + t;
    if (to_user) user[u] = dense[j]; else dense[j] = user[u];
  }
}

void launch_encodec_codes_segment_copy(int64_t* dense, int64_t* user, int segs, int item0, int nq, int T, long long T_total,
                                       long long col0, int to_user, int items, const LaunchCtx& ctx) {
  const long long n = (long long)items * nq * T;
  if (n == 0) return;
  const int ev = ctx.begin();
  codes_segment_copy_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 2048), 256, 0, ctx.stream>>>(dense, user, segs, item0, nq, T,
                                                                                                         T_total, col0, to_user, n);
  check_launch((int)cudaGetLastError(), "codes_segment_copy");
  ctx.end(ev, "codes_segment_copy", 0.0, 16.0 * n);
}

int lstm_max_batch(int num_sms, int H) { return (num_sms / (H / kLU)) * kLB; }

}  // namespace nc
