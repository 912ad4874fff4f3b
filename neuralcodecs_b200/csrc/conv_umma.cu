// tcgen05 / TMEM implicit-GEMM executor of the multi-tap row-shifted GEMM (conv_plan.h).
//
// One persistent CTA per SM, 448 threads, warp-specialised:
//   warp 0      weight producer : one thread streams pre-swizzled [BN x 32] fp32 weight tiles
//                                 HBM/L2 -> smem with 1-D bulk async copies (UBLKCP) on an mbarrier ring
//   warp 1      MMA issuer      : one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8); a conv
//                                 tap is a ROW SHIFT of the A-tile smem descriptor (the 128B swizzle is a
//                                 function of absolute smem address bits, so start addresses that are not
//                                 multiples of the 8-row period are legal; verified by tools/probe_umma.cu)
//   warps 2-5   epilogue        : TMEM -> registers (tcgen05.ld 32x32b), + bias, + residual, tanh, store
//   warps 6-13  A producers     : coalesced 128-bit loads of the channels-last input tile (+halo), the
//                                 consumer's Snake/ELU applied ONCE per staged element, tf32 rounding
//                                 (or hi/lo split for 3xTF32), swizzled st.shared, fence.proxy.async
// Accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include "conv_plan.h"
#include "umma.cuh"

namespace nc {

using namespace ptx;

constexpr int kBM = 128;
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kFirstEpilogueWarp = 2;
constexpr int kFirstProducerWarp = 6;
constexpr int kUmmaThreads = 32 * (kFirstProducerWarp + kProducerWarps);  // 448
constexpr int kAStages = 2;
constexpr int kMaxWStages = 8;
constexpr int kMaxARowIters = 6;  // (128 + span) <= 192 rows
// 227 KB opt-in limit minus the kernel's static shared memory (barriers), rounded up to 1 KB
constexpr size_t kUmmaMaxDynSmem = 227 * 1024 - 1024;

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <bool kFast>
__device__ __forceinline__ float snake_f(float x, float a, float ia) {
  float s = kFast ? __sinf(a * x) : sinf(a * x);
  return fmaf(s * s, ia, x);  // a == 0 -> ia == 0 -> x
}
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }

__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ ConvGemmParams p, const UmmaLaunch L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_tile_bytes = (uint32_t)L.a_rows_alloc * 128u;
  const uint32_t a_stage_bytes = a_tile_bytes * (p.passes == 3 ? 2u : 1u);
  const uint32_t w_tile_bytes = (uint32_t)p.BN * 128u;
  const uint32_t w_stage_bytes = w_tile_bytes * (p.passes == 3 ? 2u : 1u);
  uint8_t* sA = smem;
  uint8_t* sW = smem + kAStages * a_stage_bytes;

  __shared__ uint64_t a_full[kAStages], a_empty[kAStages];
  __shared__ uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kAStages; ++i) {
      mbar_init(&a_full[i], kProducerWarps);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(&tmem_base_s);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int tiles_per_n = p.batch * p.m_tiles_per_clip;
  const int total_tiles = p.n_tiles * tiles_per_n;

  if (warp == 0) {
    // ===================================================================== weight producer
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile / tiles_per_n;
        const unsigned mask = p.tap_mask[nt];
        const float* whi = p.W_hi + (size_t)nt * p.tiles_per_ntile * (size_t)(p.BN * 32);
        const float* wlo = p.passes == 3 ? p.W_lo + (size_t)nt * p.tiles_per_ntile * (size_t)(p.BN * 32) : nullptr;
        for (int kci = 0; kci < p.n_kc; ++kci) {
          const int kc = p.kc_begin + kci;
          for (int j = 0; j < p.n_taps; ++j) {
            if (!((mask >> j) & 1u) || kc < p.taps[j].kc_lo || kc >= p.taps[j].kc_hi) continue;
            mbar_wait(&w_empty[ws], wph ^ 1u);
            const size_t toff = (size_t)(p.taps[j].tile_base + (kc - p.taps[j].kc_lo)) * (size_t)(p.BN * 32);
            mbar_arrive_expect_tx(&w_full[ws], w_stage_bytes);
            bulk_g2s(sW + (size_t)ws * w_stage_bytes, whi + toff, w_tile_bytes, &w_full[ws]);
            if (p.passes == 3)
              bulk_g2s(sW + (size_t)ws * w_stage_bytes + w_tile_bytes, wlo + toff, w_tile_bytes, &w_full[ws]);
            if (++ws == L.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(kBM, p.BN);
      int ws = 0, as = 0, it = 0;
      uint32_t wph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int nt = tile / tiles_per_n;
        const unsigned mask = p.tap_mask[nt];
        const int buf = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&acc_empty[buf], acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
        uint32_t acc = 0;
        for (int kci = 0; kci < p.n_kc; ++kci) {
          const int kc = p.kc_begin + kci;
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_stage = smem_u32(sA) + (uint32_t)as * a_stage_bytes;
          for (int j = 0; j < p.n_taps; ++j) {
            if (!((mask >> j) & 1u) || kc < p.taps[j].kc_lo || kc >= p.taps[j].kc_hi) continue;
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            const uint32_t a_hi = a_stage + (uint32_t)(p.taps[j].shift - p.smin) * 128u;
            const uint32_t b_hi = smem_u32(sW) + (uint32_t)ws * w_stage_bytes;
            if (p.passes == 1) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_tf32(d_tmem, smem_desc_sw128(a_hi + k * 32, 1024), smem_desc_sw128(b_hi + k * 32, 1024),
                          idesc, acc);
                acc = 1;
              }
            } else {
              const uint32_t a_lo = a_hi + a_tile_bytes;
              const uint32_t b_lo = b_hi + w_tile_bytes;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_tf32(d_tmem, smem_desc_sw128(a_lo + k * 32, 1024), smem_desc_sw128(b_hi + k * 32, 1024),
                          idesc, acc);
                acc = 1;
                umma_tf32(d_tmem, smem_desc_sw128(a_hi + k * 32, 1024), smem_desc_sw128(b_lo + k * 32, 1024),
                          idesc, 1);
                umma_tf32(d_tmem, smem_desc_sw128(a_hi + k * 32, 1024), smem_desc_sw128(b_hi + k * 32, 1024),
                          idesc, 1);
              }
            }
            tc_commit(&w_empty[ws]);
            if (++ws == L.w_stages) { ws = 0; wph ^= 1u; }
          }
          tc_commit(&a_empty[as]);
          if (++as == kAStages) { as = 0; aph ^= 1u; }
        }
        tc_commit(&acc_full[buf]);
      }
    }
  } else if (warp < kFirstProducerWarp) {
    // ===================================================================== epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    int it = 0;
    const bool vec_ok = (p.n_total & 3) == 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile / tiles_per_n;
      const int rem = tile - nt * tiles_per_n;
      const int b = rem / p.m_tiles_per_clip;
      const int mt = rem - b * p.m_tiles_per_clip;
      const int buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const int row = mt * kBM + q * 32 + lane;
      const bool row_ok = row < p.m_rows;
      const long long row_off = (long long)row * p.n_total;
      float* Drow = p.D + (long long)b * p.d_clip_stride + row_off;
      const float* Rrow = p.R ? p.R + (long long)b * p.d_clip_stride + row_off : nullptr;
      const float nz = (p.noise && row_ok) ? __ldg(p.noise + (long long)b * p.m_rows + row) : 0.f;
      mbar_wait(&acc_full[buf], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)buf * 256u + ((uint32_t)(q * 32) << 16);
      for (int cc = 0; cc < p.BN / 16; ++cc) {
        float v[16];
        __syncwarp();
        tmem_ld16(t_addr + cc * 16, v);
        tmem_ld_wait();
        const int n0 = nt * p.BN + cc * 16;
        if (!row_ok || n0 >= p.n_valid) continue;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + (n0 + i) % p.bias_period);
        }
        const bool full = (n0 + 16 <= p.n_valid) && (row_off + n0 + 16 <= p.d_valid);
        if (full && vec_ok) {
          if (Rrow) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 r = __ldg(reinterpret_cast<const float4*>(Rrow + n0) + i);
              if (p.noise) {
                v[4 * i + 0] = fmaf(nz, v[4 * i + 0], r.x); v[4 * i + 1] = fmaf(nz, v[4 * i + 1], r.y);
                v[4 * i + 2] = fmaf(nz, v[4 * i + 2], r.z); v[4 * i + 3] = fmaf(nz, v[4 * i + 3], r.w);
              } else {
                v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
              }
            }
          }
          if (p.act == ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(Drow + n0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + i;
            if (n < p.n_valid && row_off + n < p.d_valid) {
              float x = v[i];
              if (Rrow) x = p.noise ? fmaf(nz, x, Rrow[n]) : x + Rrow[n];
              if (p.act == ACT_TANH) x = tanhf(x);
              Drow[n] = x;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else {
    // ===================================================================== A producers
    const int ptid = tid - kFirstProducerWarp * 32;
    const int c = ptid & 7;      // 16-byte chunk of the 128-byte row
    const int rho0 = ptid >> 3;  // 0..31
    const int rows_needed = kBM + p.span;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile / tiles_per_n;
      const int rem = tile - nt * tiles_per_n;
      const int b = rem / p.m_tiles_per_clip;
      const int mt = rem - b * p.m_tiles_per_clip;
      const float* Ab = p.A + (long long)b * p.a_clip_stride;
      const int r_base = mt * kBM + p.smin;
      for (int kci = 0; kci < p.n_kc; ++kci) {
        const int kcol = (p.kc_begin + kci) * 32 + c * 4;
        float4 al = make_float4(0, 0, 0, 0), ia = al;
        if (p.prologue == PRO_SNAKE) {
          const int ai = kcol % p.alpha_period;
          al = __ldg(reinterpret_cast<const float4*>(p.alpha + ai));
          ia = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + ai));
        }
        float4 v[kMaxARowIters];
#pragma unroll
        for (int i = 0; i < kMaxARowIters; ++i) {
          const int rho = rho0 + 32 * i;
          const int r = r_base + rho;
          const long long e = (long long)r * p.a_pitch + kcol;
          v[i] = make_float4(0, 0, 0, 0);
          if (rho < rows_needed && r >= 0 && r < p.a_rows && e < p.a_valid)
            v[i] = __ldg(reinterpret_cast<const float4*>(Ab + e));
        }
        mbar_wait(&a_empty[as], aph ^ 1u);
        uint8_t* stage = sA + (size_t)as * a_stage_bytes;
#pragma unroll
        for (int i = 0; i < kMaxARowIters; ++i) {
          const int rho = rho0 + 32 * i;
          if (rho >= rows_needed) continue;
          float4 x = v[i];
          if (p.prologue == PRO_SNAKE) {
            if (p.fast_sin) {
              x.x = snake_f<true>(x.x, al.x, ia.x); x.y = snake_f<true>(x.y, al.y, ia.y);
              x.z = snake_f<true>(x.z, al.z, ia.z); x.w = snake_f<true>(x.w, al.w, ia.w);
            } else {
              x.x = snake_f<false>(x.x, al.x, ia.x); x.y = snake_f<false>(x.y, al.y, ia.y);
              x.z = snake_f<false>(x.z, al.z, ia.z); x.w = snake_f<false>(x.w, al.w, ia.w);
            }
          } else if (p.prologue == PRO_ELU) {
            x.x = elu_f(x.x); x.y = elu_f(x.y); x.z = elu_f(x.z); x.w = elu_f(x.w);
          }
          float4 hi = make_float4(rna_tf32(x.x), rna_tf32(x.y), rna_tf32(x.z), rna_tf32(x.w));
          const uint32_t off = sw128_offset((uint32_t)rho, (uint32_t)c);
          *reinterpret_cast<float4*>(stage + off) = hi;
          if (p.passes == 3) {
            float4 lo = make_float4(rna_tf32(x.x - hi.x), rna_tf32(x.y - hi.y), rna_tf32(x.z - hi.z),
                                    rna_tf32(x.w - hi.w));
            *reinterpret_cast<float4*>(stage + a_tile_bytes + off) = lo;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (++as == kAStages) { as = 0; aph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------- host launcher
static bool g_umma_attr_set[64] = {};

size_t umma_smem_bytes(const ConvGemmParams& p, UmmaLaunch* L) {
  const int rows = ((kBM + p.span) + 7) / 8 * 8;
  const size_t a_stage = (size_t)rows * 128 * (p.passes == 3 ? 2 : 1);
  const size_t w_stage = (size_t)p.BN * 128 * (p.passes == 3 ? 2 : 1);
  const size_t budget = kUmmaMaxDynSmem - 1024 /*alignment slack*/;
  long avail = (long)budget - (long)(kAStages * a_stage);
  int ws = (int)(avail / (long)w_stage);
  if (ws > kMaxWStages) ws = kMaxWStages;
  L->w_stages = ws;
  L->a_rows_alloc = rows;
  if (ws < 2 || p.span > 64) return 0;
  return 1024 + kAStages * a_stage + (size_t)ws * w_stage;
}

// returns cudaError_t as int; 0 on success; -1 if the shape does not fit this kernel
int launch_conv_umma(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  UmmaLaunch L;
  const size_t smem = umma_smem_bytes(p, &L);
  if (smem == 0) return -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !g_umma_attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaMaxDynSmem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) g_umma_attr_set[dev] = true;
  }
  const int total_tiles = p.n_tiles * p.batch * p.m_tiles_per_clip;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (grid <= 0) return 0;
  conv_umma_kernel<<<grid, kUmmaThreads, smem, stream>>>(p, L);
  return (int)cudaGetLastError();
}

}  // namespace nc
