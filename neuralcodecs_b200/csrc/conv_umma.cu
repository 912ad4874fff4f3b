// tcgen05 / TMEM implicit-GEMM executor of the multi-tap row-shifted GEMM (conv_plan.h).
//
// One persistent CTA per SM, 576 threads, warp-specialised:
//   warp 0       weight producer : one thread streams pre-swizzled [BN x 128 B] weight tiles
//                                  HBM/L2 -> smem with 1-D bulk async copies (UBLKCP) on an mbarrier ring
//   warp 1       MMA issuer      : one thread issues tcgen05.mma (M=128, N=BN); a conv tap is a ROW SHIFT
//                                  of the A-tile smem descriptor (the 128B swizzle is a function of absolute
//                                  smem address bits, so start addresses that are not multiples of the 8-row
//                                  period are legal; verified by tools/probe_umma.cu)
//   warps 2-9    epilogue        : TMEM -> registers (tcgen05.ld 32x32b), + bias, + residual, the NEXT layer's
//                                  Snake/ELU ("post"), tanh, store; two warps per TMEM lane quarter
//   warps 10-17  A transformers  : the raw fp32 input tile (+halo rows) of one 32-channel K chunk arrives by
//                                  TMA (warp 18); these warps apply this conv's Snake/ELU ONCE per staged
//                                  element and the operand rounding / hi-lo split IN PLACE in smem, then
//                                  fence.proxy.async and hand the stage to the MMA issuer
//   warp 18      A loader        : one thread issues cp.async.bulk.tensor (3-D map [clip][row][channel],
//                                  128B swizzle, out-of-range rows zero-filled = the conv's zero padding) into
//                                  a ring of up to 6 stages, so tens of KB are in flight per SM
// Accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Operand modes: see MmaMode in conv_plan.h.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <mutex>

#include "conv_plan.h"
#include "umma.cuh"

namespace nc {

using namespace ptx;

constexpr int kBM = 128;
constexpr int kEpilogueWarps = 8;
constexpr int kProducerWarps = 8;
constexpr int kFirstEpilogueWarp = 2;
constexpr int kFirstProducerWarp = kFirstEpilogueWarp + kEpilogueWarps;           // 10
constexpr int kLoaderWarp = kFirstProducerWarp + kProducerWarps;                  // 18
constexpr int kResidualWarp = kLoaderWarp + 1;                                    // 19
constexpr int kUmmaThreads = 32 * (kResidualWarp + 1);                            // 640
constexpr int kEpiStages = 3;
constexpr int kMaxEpiStages = 5;   // conv_umma_kernel picks L.epi_stages = 3 or 5 (two-team epilogue with residual prefetch)
constexpr uint32_t kEpiStageBytes = kBM * 128;                                    // [128 rows][32 fp32]
constexpr int kMaxAStages = 6;
constexpr int kMaxWStages = 16;
// measurement only: role timeline of CTA 0 (env NC_TRACE_RU / NC_TRACE_UMMA), see scripts/ru_trace_analyze.py
constexpr int kTraceTiles = 96, kTraceEvents = 32;
#define RU_TRACE(it_, ev_) do { if (L.trace && blockIdx.x == 0 && (it_) < kTraceTiles) L.trace[(it_) * kTraceEvents + (ev_)] = (unsigned long long)clock64(); } while (0)

// 227 KB opt-in limit minus the kernel's static shared memory (barriers), rounded up to 1 KB
constexpr size_t kUmmaMaxDynSmem = 227 * 1024 - 1024 - 3072;   // 3 KB of static shared memory hold the epilogue coefficients

// producer template codes
constexpr int P_NONE = 0, P_SNAKE_FAST = 1, P_SNAKE_PRECISE = 2, P_ELU = 3;
// operand template codes (BF16X3 and F16X3 share one instantiation)
constexpr int O_TF32 = 0, O_TF32X3 = 1, O_H16X3 = 2;

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// |sin(x)| to ~1.2 ulp for |x| < 1e5 (3-term Cody-Waite reduction by pi/2 + degree-7/8 minimax
// polynomials; the sign is dropped because Snake only uses sin^2).  Branch-free.
__device__ __forceinline__ float sin_abs_precise(float x) {
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

template <bool kPrecise>
__device__ __forceinline__ float snake_f(float x, float a, float ia) {
  const float t = a * x;
  float s;
  if (kPrecise) {
    s = sin_abs_precise(t);
    if (fabsf(t) > 9.0e4f) s = sinf(t);  // Payne-Hanek territory: never reached by real activations
  } else {
    s = __sinf(t);
  }
  return fmaf(s * s, ia, x);  // a == 0 -> ia == 0 -> x   (where(alpha == 0, x, x + sin^2(alpha x)/alpha))
}
// ELU on the tensor-core path: expm1 for x <= 0 in ~12 instructions instead of expm1f's ~40 (the role timelines showed the ELU of
// the operand prologue and of the epilogue bounding Encodec's narrow layers): Taylor series to degree 8 on (-0.5, 0] (relative
// truncation error 1.1e-8) and MUFU ex2 - 1 below (relative error of the difference <= 2^-22.5 * e^x / (1 - e^x) < 2.7e-7), i.e.
// within 3 ulp of expm1f everywhere; NaN propagates, -inf -> -1.
__device__ __forceinline__ float elu_f(float x) {   // branch-free: both forms are evaluated and selected (FSEL)
  float p = fmaf(x, 1.0f / 40320.0f, 1.0f / 5040.0f);
  p = fmaf(x, p, 1.0f / 720.0f);
  p = fmaf(x, p, 1.0f / 120.0f);
  p = fmaf(x, p, 1.0f / 24.0f);
  p = fmaf(x, p, 1.0f / 6.0f);
  p = fmaf(x, p, 0.5f);
  p = fmaf(x, p, 1.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  const float neg = x > -0.5f ? x * p : e - 1.0f;
  return x > 0.f ? x : neg;
}

// Fast Snake on four values with packed fp32x2 multiplies / FMAs (sm_100 FMUL2 / FFMA2): the same round-to-nearest operations as
// snake_f<false> per element (bit-identical), 7 instead of 10 FP-pipe instructions per pair.
__device__ __forceinline__ void snake4_fast(float& x0, float& x1, float& x2, float& x3, const float4& al, const float4& ia) {
  const float2 xa = make_float2(x0, x1), xb = make_float2(x2, x3);
  const float2 ta = __fmul2_rn(make_float2(al.x, al.y), xa), tb = __fmul2_rn(make_float2(al.z, al.w), xb);
  const float2 sa = make_float2(__sinf(ta.x), __sinf(ta.y)), sb = make_float2(__sinf(tb.x), __sinf(tb.y));
  const float2 ra = __ffma2_rn(__fmul2_rn(sa, sa), make_float2(ia.x, ia.y), xa);
  const float2 rb = __ffma2_rn(__fmul2_rn(sb, sb), make_float2(ia.z, ia.w), xb);
  x0 = ra.x; x1 = ra.y; x2 = rb.x; x3 = rb.y;
}

template <int PRO>
__device__ __forceinline__ float4 prologue4(float4 x, const float4& al, const float4& ia) {
  if (PRO == P_SNAKE_FAST) {
    snake4_fast(x.x, x.y, x.z, x.w, al, ia);
  } else if (PRO == P_SNAKE_PRECISE) {
    x.x = snake_f<true>(x.x, al.x, ia.x); x.y = snake_f<true>(x.y, al.y, ia.y);
    x.z = snake_f<true>(x.z, al.z, ia.z); x.w = snake_f<true>(x.w, al.w, ia.w);
  } else if (PRO == P_ELU) {
    x.x = elu_f(x.x); x.y = elu_f(x.y); x.z = elu_f(x.z); x.w = elu_f(x.w);
  }
  return x;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 3-D tiled TMA load: box [1][rows][32 floats] at (channel c0, row r0, clip b) -> smem, completion on bar
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int r0, int b, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(b), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int r0, int b) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(r0), "r"(b)
               : "memory");
}

// v[16] += bias[n0 .. n0+16)  (index modulo bias_period: transposed convs repeat the bias per phase)
__device__ __forceinline__ void epi_bias(const ConvGemmParams& p, float (&v)[16], int n0) {
  if (!p.bias) return;
  const int bi = n0 % p.bias_period;
  if ((p.bias_period & 3) == 0 && bi + 16 <= p.bias_period) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + bi) + i);
      v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + (n0 + i) % p.bias_period);
  }
}

// Epilogue coefficients cached in shared memory (conv_umma_kernel, one N tile): with the whole carve-out given to the rings the
// L1 holds nothing, so the per-group bias / Snake-parameter loads were L2 round trips (~250 clk each) on the epilogue warps'
// critical path (role timeline).  sc: [bias | post_alpha | post_inv_alpha][256], index = output column.
__device__ __forceinline__ void epi_bias_cached(const float* sc, float (&v)[16], int n0) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 bb = *reinterpret_cast<const float4*>(sc + n0 + 4 * i);
    v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
  }
}
__device__ __forceinline__ void epi_snake_cached(const ConvGemmParams& p, const float* sc, float (&v)[16], int n0) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 al = *reinterpret_cast<const float4*>(sc + 256 + n0 + 4 * i);
    const float4 ia = *reinterpret_cast<const float4*>(sc + 512 + n0 + 4 * i);
    if (p.precise_sin) {
      v[4 * i + 0] = snake_f<true>(v[4 * i + 0], al.x, ia.x); v[4 * i + 1] = snake_f<true>(v[4 * i + 1], al.y, ia.y);
      v[4 * i + 2] = snake_f<true>(v[4 * i + 2], al.z, ia.z); v[4 * i + 3] = snake_f<true>(v[4 * i + 3], al.w, ia.w);
    } else {
      snake4_fast(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], al, ia);
    }
  }
}

__device__ __forceinline__ void epi_snake_smem(bool precise, const float* al_s, const float* ia_s, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 al = *reinterpret_cast<const float4*>(al_s + 4 * i);
    const float4 ia = *reinterpret_cast<const float4*>(ia_s + 4 * i);
    if (precise) {
      v[4 * i + 0] = snake_f<true>(v[4 * i + 0], al.x, ia.x); v[4 * i + 1] = snake_f<true>(v[4 * i + 1], al.y, ia.y);
      v[4 * i + 2] = snake_f<true>(v[4 * i + 2], al.z, ia.z); v[4 * i + 3] = snake_f<true>(v[4 * i + 3], al.w, ia.w);
    } else {
      snake4_fast(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], al, ia);
    }
  }
}

// the consumer's activation (applied once per element, here) followed by this layer's own activation
__device__ __forceinline__ void epi_post(const ConvGemmParams& p, float (&v)[16], int n0) {
  if (p.post == PRO_SNAKE) {
    const int pi = n0 % p.post_period;
    if ((p.post_period & 3) == 0 && pi + 16 <= p.post_period) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 al = __ldg(reinterpret_cast<const float4*>(p.post_alpha + pi) + i);
        const float4 ia = __ldg(reinterpret_cast<const float4*>(p.post_inv_alpha + pi) + i);
        if (p.precise_sin) {
          v[4 * i + 0] = snake_f<true>(v[4 * i + 0], al.x, ia.x); v[4 * i + 1] = snake_f<true>(v[4 * i + 1], al.y, ia.y);
          v[4 * i + 2] = snake_f<true>(v[4 * i + 2], al.z, ia.z); v[4 * i + 3] = snake_f<true>(v[4 * i + 3], al.w, ia.w);
        } else {
          v[4 * i + 0] = snake_f<false>(v[4 * i + 0], al.x, ia.x); v[4 * i + 1] = snake_f<false>(v[4 * i + 1], al.y, ia.y);
          v[4 * i + 2] = snake_f<false>(v[4 * i + 2], al.z, ia.z); v[4 * i + 3] = snake_f<false>(v[4 * i + 3], al.w, ia.w);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ai = (n0 + i) % p.post_period;
        const float al = __ldg(p.post_alpha + ai), ia = __ldg(p.post_inv_alpha + ai);
        v[i] = p.precise_sin ? snake_f<true>(v[i], al, ia) : snake_f<false>(v[i], al, ia);
      }
    }
  } else if (p.post == PRO_ELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = elu_f(v[i]);
  }
  if (p.act == ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
  }
}

// Encodec time_group_norm: a warp's partial {sum, sum of squares} of its rows of one tile -> the clip's fp64 totals
__device__ __forceinline__ void gn_stats_add(double* dst, float s, float ss, int lane) {
  double ds = (double)s, dss = (double)ss;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dss += __shfl_xor_sync(0xffffffffu, dss, o);
  }
  if (lane == 0) { atomicAdd(dst, ds); atomicAdd(dst + 1, dss); }
}

template <int OPS, int PRO, int RIT>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_umma_kernel(const __grid_constant__ ConvGemmParams p, const UmmaLaunch L, const __grid_constant__ CUtensorMap tmapA,
                 const __grid_constant__ CUtensorMap tmapD, const __grid_constant__ CUtensorMap tmapR) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_tile_bytes = (uint32_t)L.a_rows_alloc * 128u;
  const uint32_t a_stage_bytes = a_tile_bytes * (OPS == O_TF32X3 ? 2u : 1u);
  const uint32_t w_tile_bytes = (uint32_t)p.BN * ((OPS == O_H16X3 && p.w_hi_only) ? 64u : 128u);
  const uint32_t w_stage_bytes = w_tile_bytes * (OPS == O_TF32X3 ? 2u : 1u);
  uint8_t* sA = smem;
  uint8_t* sW = smem + (uint32_t)L.a_stages * a_stage_bytes;
  uint8_t* sE = sW + (((uint32_t)L.w_stages * w_stage_bytes + 1023u) & ~1023u);   // epilogue ring (TMA epilogue only)

  __shared__ uint64_t raw_full[kMaxAStages], a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  __shared__ uint64_t acc_full[2], acc_empty[2], lo_empty;
  __shared__ uint64_t r_full[kMaxEpiStages], e_free[kMaxEpiStages];
  __shared__ __align__(16) float s_coef[3 * 256];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // short accumulation chains (conv_plan.h: acc_split): TMEM columns [0,BN) / [BN,2BN) = main partials,
  // [2BN,3BN) = lo-term accumulator, [3BN,4BN) = running sum of the folded partials
  // acc_split = 2 (N tile up to 256): main at [0,BN), lo at [256,256+BN), one tile in flight, no folding.
  const bool s3 = (OPS == O_H16X3 || OPS == O_TF32X3) && p.acc_split != 0;
  const bool s3_wide = s3 && p.acc_split == 2;
  const int fold_kc = (s3 && !s3_wide && p.fold_kc > 0) ? p.fold_kc : p.n_kc;
  const uint32_t acc_stride = (s3 && !s3_wide) ? (uint32_t)p.BN : 256u;
  const uint32_t lo_col = s3_wide ? 256u : 2u * acc_stride;
  auto part_buf = [&](uint32_t pc) { return s3_wide ? 0 : (int)(pc & 1u); };
  auto part_phase = [&](uint32_t pc) { return s3_wide ? (pc & 1u) : ((pc >> 1) & 1u); };

  if (tid == 0) {
    mbar_init(&lo_empty, kEpilogueWarps);
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&a_full[i], kProducerWarps);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      // two-team epilogue, one 32-column group per tile: only the team that owns the tile's group reads the accumulator
      mbar_init(&acc_empty[i], (L.epi_teams && p.BN == 32) ? kEpilogueWarps / 2 : kEpilogueWarps);
    }
    for (int i = 0; i < kMaxEpiStages; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&e_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(&tmem_base_s);
    tmem_relinquish();
  }
  const bool coef_cached = L.tma_epilogue && p.n_tiles == 1 && p.BN <= 256;
  if (coef_cached) {
    for (int i = threadIdx.x; i < p.BN; i += blockDim.x) {
      s_coef[i] = p.bias ? __ldg(p.bias + i % p.bias_period) : 0.f;
      if (p.post == PRO_SNAKE) {
        s_coef[256 + i] = __ldg(p.post_alpha + i % p.post_period);
        s_coef[512 + i] = __ldg(p.post_inv_alpha + i % p.post_period);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int tiles_per_n = p.batch * p.m_tiles_per_clip;
  const int total_tiles = p.n_tiles * tiles_per_n;

  // ===================================================================== MMA issuer role
  // One thread sustains one tcgen05.mma per ~80 clk whatever N is (profiles/r02_mma_issue_rate_probe.txt), which bounds every layer
  // with N <= 128; two issuing warps double the rate.  With L.dual_issue (no residual tile to load, dense taps, one N tile, no
  // folded partials) the otherwise idle residual-loader warp issues the odd tiles into accumulator 1 while warp 1 issues the even
  // tiles into accumulator 0: both walk the same A / weight rings, their positions computed from the tile index.  Measured on the
  // layers that qualify (Encodec's convs without a residual; job AM): results identical, time neutral to -2 % -- those layers are
  // bound by their transform / epilogue warps, not by MMA issue -- so it is off by default (NC_DUAL_ISSUE=1); the layers that ARE
  // issue-bound (the fused residual units) need their weight ring deepened or CTA pairs before a second issuer can help.
  auto mma_role = [&](const int it0, const int it_step) {
      const uint32_t idesc = OPS == O_H16X3 ? idesc_f16(kBM, p.BN, p.mode == MODE_BF16X3 ? 1 : 0) : idesc_tf32(kBM, p.BN);
      const uint64_t a_desc0 = desc_at(smem_u32(sA));
      const uint64_t w_desc0 = (OPS == O_H16X3 && p.w_hi_only) ? (kDescSw64Base | (uint64_t)((smem_u32(sW) & 0x3FFFFu) >> 4))
                                                              : desc_at(smem_u32(sW));
      const uint32_t a_stage_u = a_stage_bytes >> 4, w_stage_u = w_stage_bytes >> 4;   // descriptor units (16 B)
      const uint32_t a_lo_u = a_tile_bytes >> 4, w_lo_u = w_tile_bytes >> 4;
      const uint32_t tap_u = (uint32_t)p.dense_step * 8u;                             // rows * 128 B / 16
      int ws = 0, as = 0, it = it0;
      uint32_t wph = 0, aph = 0;
      uint32_t pc = 0;   // accumulator partials issued so far (== tiles when nothing is folded)
      uint64_t a_desc = a_desc0, w_desc = w_desc0;
      const uint32_t lo_tmem = tmem_base + lo_col;
      const int n_w = p.n_kc * p.n_taps;   // weight stages per tile (dual issue only: dense taps, one N tile)
      for (int tile = (int)blockIdx.x + it0 * (int)gridDim.x; tile < total_tiles; tile += it_step * (int)gridDim.x, it += it_step) {
        if (it_step == 2) {
          // dual issue: this thread takes every other tile, so the ring positions follow from the tile index
          const long long sw = (long long)it * n_w, sa = (long long)it * p.n_kc;
          ws = (int)(sw % L.w_stages); wph = (uint32_t)((sw / L.w_stages) & 1);
          as = (int)(sa % L.a_stages); aph = (uint32_t)((sa / L.a_stages) & 1);
          w_desc = w_desc0 + (uint64_t)ws * w_stage_u;
          a_desc = a_desc0 + (uint64_t)as * a_stage_u;
          pc = (uint32_t)it;
        }
        if (s3) {   // the previous tile's epilogue has read the lo accumulator
          mbar_wait(&lo_empty, ((uint32_t)it & 1u) ^ 1u);
          tc_fence_after();
        }
        int buf = 0, in_part = 0;
        bool have_buf = false;
        uint32_t d_tmem = 0;
        uint32_t acc = 0, acc_lo = 0;
        RU_TRACE(it, 3);
        const unsigned mask = p.dense_step >= 0 ? 0u : (unsigned)p.tap_mask[tile % p.n_tiles];
        for (int kci = 0; kci < p.n_kc; ++kci) {
          if (!have_buf) {
            buf = part_buf(pc);
            mbar_wait(&acc_empty[buf], part_phase(pc) ^ 1u);
            tc_fence_after();
            d_tmem = tmem_base + (uint32_t)buf * acc_stride;
            acc = 0;
            have_buf = true;
            if (kci == 0) RU_TRACE(it, 4);
          }
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          if (kci == 0) RU_TRACE(it, 5);
          uint64_t a_tap = a_desc;
          for (int j = 0; j < p.n_taps; ++j) {
            if (p.dense_step >= 0) {
              // stride-1 conv: every tap reads every chunk; tap j is the A tile shifted down by j*dilation rows
              if (j) a_tap += tap_u;
            } else {
              const int kc = p.kc_begin + kci;
              if (!((mask >> j) & 1u) || kc < p.taps[j].kc_lo || kc >= p.taps[j].kc_hi) continue;
              a_tap = a_desc + (uint32_t)(p.taps[j].shift - p.smin) * 8u;
            }
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            const uint64_t a0 = a_tap, b0 = w_desc;
            if ((L.knock & 8) && OPS == O_H16X3) {
              if (!acc) umma_f16(d_tmem, a0, b0, idesc, 0);
            } else if (OPS == O_TF32) {
#pragma unroll
              for (int k = 0; k < 4; ++k)   // K step = 8 tf32 = 32 B = +2 descriptor units
                umma_tf32(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);
            } else if (OPS == O_TF32X3) {
              const uint64_t a1 = a0 + a_lo_u, b1 = b0 + w_lo_u;
              if (s3) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_tf32(lo_tmem, a1 + 2 * k, b0 + 2 * k, idesc, acc_lo | (uint32_t)k);  // lo * hi -> lo accumulator
                  umma_tf32(lo_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1);                     // hi * lo -> lo accumulator
                  umma_tf32(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);      // hi * hi -> main partial
                }
                acc_lo = 1;
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_tf32(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);  // lo * hi
                  umma_tf32(d_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1);                  // hi * lo
                  umma_tf32(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1);                  // hi * hi
                }
              }
            } else {
              // row = [32 hi halves (64 B) | 32 lo halves (64 B)]; K step = 16 halves = 32 B
              if (s3) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  umma_f16(lo_tmem, a0 + 4 + 2 * k, b0 + 2 * k, idesc, acc_lo | (uint32_t)k);  // lo * hi -> lo accumulator
                  umma_f16(lo_tmem, a0 + 2 * k, b0 + 4 + 2 * k, idesc, 1);                     // hi * lo -> lo accumulator
                  umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);          // hi * hi -> main partial
                }
                acc_lo = 1;
              } else if (p.passes == 3) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  umma_f16(d_tmem, a0 + 4 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);  // lo * hi
                  umma_f16(d_tmem, a0 + 2 * k, b0 + 4 + 2 * k, idesc, 1);                  // hi * lo
                  umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1);                      // hi * hi
                }
              } else if (p.passes == 2) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  umma_f16(d_tmem, a0 + 4 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);  // lo * hi
                  umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1);                      // hi * hi
                }
              } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);
              }
            }
            tc_commit(&w_empty[ws]);
            acc = 1;
            w_desc += w_stage_u;
            if (++ws == L.w_stages) { ws = 0; wph ^= 1u; w_desc = w_desc0; }
          }
          tc_commit(&a_empty[as]);
          a_desc += a_stage_u;
          if (++as == L.a_stages) { as = 0; aph ^= 1u; a_desc = a_desc0; }
          if (++in_part == fold_kc || kci == p.n_kc - 1) {   // this partial is complete: hand it to the epilogue warps
            if (kci == p.n_kc - 1) RU_TRACE(it, 6);
            tc_commit(&acc_full[buf]);
            ++pc;
            in_part = 0;
            have_buf = false;
          }
        }
      }
  };

  if (warp == 0) {
    // ===================================================================== weight producer
    if (elect_one()) {
      int ws = 0, wcount = 0, it = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        RU_TRACE(it, 15);
        const int nt = tile % p.n_tiles;   // N tile fastest: the N tiles of one M tile run side by side, A re-reads hit L2
        const unsigned mask = p.tap_mask[nt];
        const float* wbase = p.W + (size_t)nt * p.tiles_per_ntile * (size_t)p.w_tile_floats;
        for (int kci = 0; kci < p.n_kc; ++kci) {
          const int kc = p.kc_begin + kci;
          for (int j = 0; j < p.n_taps; ++j) {
            if (!((mask >> j) & 1u) || kc < p.taps[j].kc_lo || kc >= p.taps[j].kc_hi) continue;
            mbar_wait(&w_empty[ws], wph ^ 1u);
            const size_t toff = (size_t)(p.taps[j].tile_base + (kc - p.taps[j].kc_lo)) * (size_t)p.w_tile_floats;
            if ((L.knock & 1) && wcount >= L.w_stages) {
              mbar_arrive(&w_full[ws]);
            } else {
              mbar_arrive_expect_tx(&w_full[ws], w_stage_bytes);
              bulk_g2s(sW + (size_t)ws * w_stage_bytes, wbase + toff, w_stage_bytes, &w_full[ws]);
            }
            ++wcount;
            if (++ws == L.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // ONE elected lane runs the whole persistent loop (elect.sync tells ptxas the region is single-threaded,
    // so descriptors go straight to uniform registers and each tcgen05.mma is a single UTCHMMA).
    if (elect_one()) mma_role(0, L.dual_issue ? 2 : 1);
  } else if (warp == kLoaderWarp) {
    // ===================================================================== A loader (TMA)
    if (elect_one()) {
      int as = 0, acount = 0, it = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int nt = tile % p.n_tiles;   // N tile fastest: the N tiles of one M tile run side by side, A re-reads hit L2
        const int rem = tile / p.n_tiles;
        const int b = rem / p.m_tiles_per_clip;
        const int mt = rem - b * p.m_tiles_per_clip;
        const int r_base = mt * kBM + p.smin;
        for (int kci = 0; kci < p.n_kc; ++kci) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (kci == 0) RU_TRACE(it, 0);
          if (kci == p.n_kc - 1) RU_TRACE(it, 1);
          if ((L.knock & 4) && acount >= L.a_stages) {
            mbar_arrive(&raw_full[as]);
          } else {
            mbar_arrive_expect_tx(&raw_full[as], a_tile_bytes);
            tma_load_3d(sA + (size_t)as * a_stage_bytes, &tmapA, (p.kc_begin + kci) * 32, r_base, b, &raw_full[as]);
          }
          ++acount;
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == kResidualWarp) {
    // ===================================================================== second MMA issuer (dual issue) / residual loader (TMA)
    if (L.dual_issue) {
      if (elect_one()) mma_role(1, 2);
    } else if (L.tma_epilogue && p.R && elect_one()) {
      const int groups = p.BN / 32;
      int es = 0;
      uint32_t eph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;   // N tile fastest: the N tiles of one M tile run side by side, A re-reads hit L2
        const int rem = tile / p.n_tiles;
        const int b = rem / p.m_tiles_per_clip;
        const int mt = rem - b * p.m_tiles_per_clip;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(&e_free[es], eph ^ 1u);
          mbar_arrive_expect_tx(&r_full[es], kEpiStageBytes);
          tma_load_3d(sE + (size_t)es * kEpiStageBytes, &tmapR, nt * p.BN + g * 32, mt * kBM, b, &r_full[es]);
          if (++es == L.epi_stages) { es = 0; eph ^= 1u; }
        }
      }
    }
  } else if (warp < kFirstProducerWarp) {
    // ===================================================================== epilogue
    const int q = warp & 3;                                  // TMEM lane quarter this warp may touch
    const int half = (warp - kFirstEpilogueWarp) >> 2;       // which 16-column half of a 32-column group
    int it = 0;
    if (L.tma_epilogue) {
      // Output (and residual) tiles move as [128 rows x 32 cols] fp32 boxes through a 128B-swizzled smem
      // ring: residual arrives by TMA load, the result leaves by TMA store; every thread touches only its own
      // row with conflict-free 128-bit smem accesses, so no uncoalesced global traffic is issued.
      const int groups = p.BN / 32;
      const bool leader = warp == kFirstEpilogueWarp && lane == 0;
      if (L.epi_teams) {
        // Two teams of 4 warps (each covers the 4 TMEM lane quarters) take alternate 32-column groups, so two groups'
        // latency chains (tcgen05.ld -> accumulator hand-back -> stage wait -> math -> smem -> barrier -> TMA store) overlap:
        // the role timeline showed narrow layers bound by that chain (~3k clk per group), not by any throughput.
        // Group k (counted over the whole launch) uses stage k % E; a team frees a stage one of ITS stores later.
        const int team = half;
        const bool tleader = ((warp - kFirstEpilogueWarp) & 3) == 0 && lane == 0;
        const int E = L.epi_stages;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        int prev = -1;
        long long k0 = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it, k0 += groups) {
          const int nt = tile % p.n_tiles;
          const int rem = tile / p.n_tiles;
          const int b = rem / p.m_tiles_per_clip;
          const int mt = rem - b * p.m_tiles_per_clip;
          const int rloc = q * 32 + lane;
          const int row = mt * kBM + rloc;
          float gn_s = 0.f, gn_ss = 0.f;
          const int g_first = (int)((k0 ^ (long long)team) & 1);
          if (g_first < groups) {
            const int buf = it & 1;
            if (leader) RU_TRACE(it, 19);
            mbar_wait(&acc_full[buf], ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            if (leader) RU_TRACE(it, 12);
            const uint32_t t_addr = tmem_base + (uint32_t)buf * acc_stride + lane_bits;
            const int g_last = g_first + ((groups - 1 - g_first) & ~1);
            for (int g = g_first; g < groups; g += 2) {
              const long long k = k0 + g;
              const int es = (int)(k % E);
              const uint32_t eph = (uint32_t)((k / E) & 1);
              uint8_t* stage = sE + (size_t)es * kEpiStageBytes;
#pragma unroll 1
              for (int hh = 0; hh < 2; ++hh) {
                float v[16];
                __syncwarp();
                tmem_ld16(t_addr + g * 32 + hh * 16, v);
                tmem_ld_wait();
                if (g == g_last && hh == 1) {   // this warp is done with the accumulator
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                const int n0 = nt * p.BN + g * 32 + hh * 16;
                if (coef_cached) epi_bias_cached(s_coef, v, n0); else epi_bias(p, v, n0);
                if (p.gn_stats && row < p.m_rows) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) { gn_s += v[i]; gn_ss = fmaf(v[i], v[i], gn_ss); }
                }
                if (hh == 0) {
                  if (p.R) mbar_wait(&r_full[es], eph); else mbar_wait(&e_free[es], eph ^ 1u);
                }
                if (p.R && !(L.knock & 16)) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 r = *reinterpret_cast<const float4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(hh * 4 + i)));
                    v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
                  }
                }
                if (!(L.knock & 16)) {
                  if (coef_cached && p.post == PRO_SNAKE && p.act != ACT_TANH) epi_snake_cached(p, s_coef, v, n0); else epi_post(p, v, n0);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  *reinterpret_cast<float4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(hh * 4 + i))) =
                      make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              }
              fence_proxy_async_smem();
              if (team == 0) asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueWarps * 16) : "memory");
              else asm volatile("bar.sync 3, %0;" ::"n"(kEpilogueWarps * 16) : "memory");
              if (tleader) {
                if (!(L.knock & 32)) tma_store_3d(&tmapD, stage, nt * p.BN + g * 32, mt * kBM, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this team's previous store has left smem
                if (prev >= 0) mbar_arrive(&e_free[prev]);
                prev = es;
                if (leader && g == g_last) RU_TRACE(it, 14);
              }
            }
          }
          if (p.gn_stats) gn_stats_add(p.gn_stats + 2 * b, gn_s, gn_ss, lane);
        }
        if (tleader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      } else {
      int es = 0, prev = -1;
      uint32_t eph = 0;
      uint32_t pc = 0;   // accumulator partials consumed so far (mirrors the MMA issuer's counter)
      const int n_parts = (p.n_kc + fold_kc - 1) / fold_kc;
      const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
      const uint32_t t_lo = tmem_base + lo_col + lane_bits, t_run = tmem_base + 3u * acc_stride + lane_bits;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int nt = tile % p.n_tiles;   // N tile fastest: the N tiles of one M tile run side by side, A re-reads hit L2
        const int rem = tile / p.n_tiles;
        const int b = rem / p.m_tiles_per_clip;
        const int mt = rem - b * p.m_tiles_per_clip;
        const int rloc = q * 32 + lane;
        const int row = mt * kBM + rloc;
        const float nz = (p.noise && row < p.m_rows) ? __ldg(p.noise + (long long)b * p.m_rows + row) : 0.f;
        float gn_s = 0.f, gn_ss = 0.f;   // GroupNorm statistics of this thread's row (p.gn_stats)
        if (leader) RU_TRACE(it, 19);
        // fold every main partial but the last into the running sum (fp32 round-to-nearest adds, kept in TMEM; each
        // thread only ever touches its own lane and columns of it)
        for (int part = 0; part + 1 < n_parts; ++part, ++pc) {
          const int pbuf = part_buf(pc);
          mbar_wait(&acc_full[pbuf], part_phase(pc));
          tc_fence_after();
          const uint32_t t_part = tmem_base + (uint32_t)pbuf * acc_stride + lane_bits;
          for (int g = 0; g < groups; ++g) {
            float v[16], r[16];
            __syncwarp();
            tmem_ld16(t_part + g * 32 + half * 16, v);
            if (part > 0) tmem_ld16(t_run + g * 32 + half * 16, r);
            tmem_ld_wait();
            if (part > 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += r[i];
            }
            tmem_st16(t_run + g * 32 + half * 16, v);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[pbuf]);
        }
        const int buf = part_buf(pc);
        const uint32_t acc_ph = part_phase(pc);
        ++pc;
        mbar_wait(&acc_full[buf], acc_ph);
        tc_fence_after();
        if (leader) RU_TRACE(it, 12);
        const uint32_t t_addr = tmem_base + (uint32_t)buf * acc_stride + lane_bits;
        for (int g = 0; g < groups; ++g) {
          float v[16];
          __syncwarp();
          tmem_ld16(t_addr + g * 32 + half * 16, v);
          if (s3) {
            float r[16], l[16];
            if (n_parts > 1) tmem_ld16(t_run + g * 32 + half * 16, r);
            tmem_ld16(t_lo + g * 32 + half * 16, l);
            tmem_ld_wait();
            if (n_parts > 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += r[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += l[i];
          } else {
            tmem_ld_wait();
          }
          if (leader && g == 0) RU_TRACE(it, 22);
          if (g == groups - 1) {   // accumulator fully read by this warp: hand the TMEM buffers back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&acc_empty[buf]);
              if (s3) mbar_arrive(&lo_empty);
            }
          }
          const int n0 = nt * p.BN + g * 32 + half * 16;
          if (coef_cached) epi_bias_cached(s_coef, v, n0); else epi_bias(p, v, n0);
          if (p.gn_stats && row < p.m_rows) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { gn_s += v[i]; gn_ss = fmaf(v[i], v[i], gn_ss); }
          }
          uint8_t* stage = sE + (size_t)es * kEpiStageBytes;
          if (leader && g == 0) RU_TRACE(it, 23);
          if (p.R) mbar_wait(&r_full[es], eph); else mbar_wait(&e_free[es], eph ^ 1u);
          if (leader && g == 0) RU_TRACE(it, 13);
          if (p.R && !(L.knock & 16)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 r = *reinterpret_cast<const float4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(half * 4 + i)));
              if (p.noise) {
                v[4 * i + 0] = fmaf(nz, v[4 * i + 0], r.x); v[4 * i + 1] = fmaf(nz, v[4 * i + 1], r.y);
                v[4 * i + 2] = fmaf(nz, v[4 * i + 2], r.z); v[4 * i + 3] = fmaf(nz, v[4 * i + 3], r.w);
              } else {
                v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
              }
            }
          }
          if (!(L.knock & 16)) {
            if (coef_cached && p.post == PRO_SNAKE && p.act != ACT_TANH) epi_snake_cached(p, s_coef, v, n0); else epi_post(p, v, n0);
          }
          if (leader && g == 0) RU_TRACE(it, 24);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(half * 4 + i))) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          fence_proxy_async_smem();
          if (leader && g == 0) RU_TRACE(it, 25);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueWarps * 32) : "memory");
          if (leader && g == 0) RU_TRACE(it, 26);
          if (leader) {
            if (!(L.knock & 32)) tma_store_3d(&tmapD, stage, nt * p.BN + g * 32, mt * kBM, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the previous store has left smem
            if (prev >= 0) mbar_arrive(&e_free[prev]);
            prev = es;
            if (g == groups - 1) RU_TRACE(it, 14);
          }
          if (++es == L.epi_stages) { es = 0; eph ^= 1u; }
        }
        if (p.gn_stats) gn_stats_add(p.gn_stats + 2 * b, gn_s, gn_ss, lane);
      }
      if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }   // one-team epilogue
    } else {
      // direct path (ragged outputs, Cout not a multiple of 32): each thread stores its own row
      const bool vec_ok = (p.n_total & 3) == 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int nt = tile % p.n_tiles;   // N tile fastest: the N tiles of one M tile run side by side, A re-reads hit L2
        const int rem = tile / p.n_tiles;
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
        float gn_s = 0.f, gn_ss = 0.f;
        for (int cc = half; cc < p.BN / 16; cc += 2) {
          float v[16];
          __syncwarp();
          tmem_ld16(t_addr + cc * 16, v);
          tmem_ld_wait();
          const int n0 = nt * p.BN + cc * 16;
          if (!row_ok || n0 >= p.n_valid) continue;
          const bool full = (n0 + 16 <= p.n_valid) && (row_off + n0 + 16 <= p.d_valid);
          epi_bias(p, v, n0);
          if (p.gn_stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n0 + i < p.n_valid && row_off + n0 + i < p.d_valid) { gn_s += v[i]; gn_ss = fmaf(v[i], v[i], gn_ss); }
          }
          if (Rrow) {
            if (full && vec_ok) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(Rrow + n0) + i);
                if (p.noise) {
                  v[4 * i + 0] = fmaf(nz, v[4 * i + 0], r.x); v[4 * i + 1] = fmaf(nz, v[4 * i + 1], r.y);
                  v[4 * i + 2] = fmaf(nz, v[4 * i + 2], r.z); v[4 * i + 3] = fmaf(nz, v[4 * i + 3], r.w);
                } else {
                  v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int n = n0 + i;
                if (n < p.n_valid && row_off + n < p.d_valid) v[i] = p.noise ? fmaf(nz, v[i], Rrow[n]) : v[i] + Rrow[n];
              }
            }
          }
          epi_post(p, v, n0);
          if (full && vec_ok) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(Drow + n0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = n0 + i;
              if (n < p.n_valid && row_off + n < p.d_valid) Drow[n] = v[i];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        if (p.gn_stats) gn_stats_add(p.gn_stats + 2 * b, gn_s, gn_ss, lane);
      }
    }
  } else {
    // ===================================================================== A transformers
    const int ptid = tid - kFirstProducerWarp * 32;
    const int c = ptid & 7;      // 16-byte (4-float) chunk of the 128-byte input row
    // rows of one warp: R, R+2, R+4, R+6.  The hi halves of a row occupy 16-byte chunks {0..3} ^ (row & 7), i.e. the low or
    // the high 64 B of its 128-byte line depending on bit 2 of the row: four CONSECUTIVE rows put all their 8-byte hi
    // (and lo) stores into the same 16 banks (4 wavefronts per store instead of 2; ncu counted 52 M bank conflicts per
    // launch in round 1); rows two apart split them evenly.
    const int rho0 = 8 * (ptid >> 6) + ((ptid >> 5) & 1) + 2 * ((ptid & 31) >> 3);  // 0..31, each once
    const int rows_needed = kBM + p.span;
    int as = 0, tt = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tt) {
      for (int kci = 0; kci < p.n_kc; ++kci) {
        if (kci == 0 && ptid == 0) RU_TRACE(tt, 20);
        float4 al = make_float4(0.f, 0.f, 0.f, 0.f), ia = al;
        if (PRO == P_SNAKE_FAST || PRO == P_SNAKE_PRECISE) {
          const int ai = ((p.kc_begin + kci) * 32 + c * 4) % p.alpha_period;
          al = __ldg(reinterpret_cast<const float4*>(p.alpha + ai));
          ia = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + ai));
        }
        uint8_t* stage = sA + (size_t)as * a_stage_bytes;
        mbar_wait(&raw_full[as], aph);
        if (kci == 0 && ptid == 0) RU_TRACE(tt, 2);
        if (L.knock & 2) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[as]);
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
          continue;
        }
        if (OPS == O_H16X3 && p.dw_w != nullptr) {
          // ---- depthwise k7 conv folded into the operand prologue (SNAC ResidualUnit; conv_plan.h)
          // phase 1: Snake on every row of the tile (halo included), fp32 in place.  OOB rows are TMA zero fill and
          // snake(0) = 0, which is exactly the conv's zero padding.
#pragma unroll
          for (int i = 0; i < RIT; ++i) {
            const int rho = rho0 + 32 * i;
            if (rho < rows_needed) {
              float4* ptr = reinterpret_cast<float4*>(stage + sw128_offset((uint32_t)rho, (uint32_t)c));
              *ptr = prologue4<PRO>(*ptr, al, ia);
            }
          }
          asm volatile("bar.sync 2, %0;" ::"n"(kProducerWarps * 32) : "memory");
          // phase 2: 7 taps per output; output row o reads stage rows o + j*dil
          const int ch = (p.kc_begin + kci) * 32 + c * 4;
          const int d = p.dw_dil;
          float4 acc[kBM / 32];
          const float4 bb = p.dw_b ? __ldg(reinterpret_cast<const float4*>(p.dw_b + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
          // packed fp32x2 FMAs (FFMA2: two independent round-to-nearest FMAs per instruction, identical results): this prologue is
          // instruction-issue bound (profiles/r02_narrow_layer_timelines.txt)
          float2 a01[kBM / 32], a23[kBM / 32];
#pragma unroll
          for (int i = 0; i < kBM / 32; ++i) { a01[i] = make_float2(bb.x, bb.y); a23[i] = make_float2(bb.z, bb.w); }
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(p.dw_w + (size_t)j * p.alpha_period + ch));
            const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
            for (int i = 0; i < kBM / 32; ++i) {
              const float4 x = *reinterpret_cast<const float4*>(stage + sw128_offset((uint32_t)(rho0 + 32 * i + j * d), (uint32_t)c));
              a01[i] = __ffma2_rn(w01, make_float2(x.x, x.y), a01[i]);
              a23[i] = __ffma2_rn(w23, make_float2(x.z, x.w), a23[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < kBM / 32; ++i) acc[i] = make_float4(a01[i].x, a01[i].y, a23[i].x, a23[i].y);
          if (p.dw_post_alpha) {
            const float4 pa = __ldg(reinterpret_cast<const float4*>(p.dw_post_alpha + ch));
            const float4 pi = __ldg(reinterpret_cast<const float4*>(p.dw_post_inv_alpha + ch));
#pragma unroll
            for (int i = 0; i < kBM / 32; ++i)
              acc[i] = p.precise_sin ? prologue4<P_SNAKE_PRECISE>(acc[i], pa, pi) : prologue4<P_SNAKE_FAST>(acc[i], pa, pi);
          }
          asm volatile("bar.sync 2, %0;" ::"n"(kProducerWarps * 32) : "memory");
          // phase 3: hi | lo operand halves into stage rows [0, 128)
#pragma unroll
          for (int i = 0; i < kBM / 32; ++i) {
            const int rho = rho0 + 32 * i;
            const float4 x = acc[i];
            uint2 hi, lo;
            if (p.mode == MODE_BF16X3) {
              hi.x = pack_bf16(x.x, x.y); hi.y = pack_bf16(x.z, x.w);
              const __nv_bfloat162 h0 = *reinterpret_cast<__nv_bfloat162*>(&hi.x), h1 = *reinterpret_cast<__nv_bfloat162*>(&hi.y);
              lo.x = pack_bf16(x.x - __low2float(h0), x.y - __high2float(h0));
              lo.y = pack_bf16(x.z - __low2float(h1), x.w - __high2float(h1));
            } else {
              hi.x = pack_f16(x.x, x.y); hi.y = pack_f16(x.z, x.w);
              const __half2 h0 = *reinterpret_cast<__half2*>(&hi.x), h1 = *reinterpret_cast<__half2*>(&hi.y);
              lo.x = pack_f16(x.x - __low2float(h0), x.y - __high2float(h0));
              lo.y = pack_f16(x.z - __low2float(h1), x.w - __high2float(h1));
            }
            const uint32_t sub = (uint32_t)(c & 1) * 8u;
            *reinterpret_cast<uint2*>(stage + sw128_offset((uint32_t)rho, (uint32_t)(c >> 1)) + sub) = hi;
            *reinterpret_cast<uint2*>(stage + sw128_offset((uint32_t)rho, 4u + (uint32_t)(c >> 1)) + sub) = lo;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[as]);
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
          continue;
        }
        float4 cur[RIT];
#pragma unroll
        for (int i = 0; i < RIT; ++i) {
          const int rho = rho0 + 32 * i;
          cur[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rho < rows_needed) cur[i] = *reinterpret_cast<const float4*>(stage + sw128_offset((uint32_t)rho, (uint32_t)c));
        }
        if (OPS == O_H16X3) __syncwarp();  // the 8 lanes of a row exchange 16-byte chunks in place
#pragma unroll
        for (int i = 0; i < RIT; ++i) {
          const int rho = rho0 + 32 * i;
          const float4 x = prologue4<PRO>(cur[i], al, ia);
          if (rho < rows_needed) {
            if (OPS == O_H16X3) {
              // hi halves of channels 4c..4c+3 -> bytes [8c, 8c+8) of the row; lo halves -> 64 + [8c, 8c+8)
              uint2 hi, lo;
              if (p.mode == MODE_BF16X3) {
                hi.x = pack_bf16(x.x, x.y); hi.y = pack_bf16(x.z, x.w);
                const __nv_bfloat162 h0 = *reinterpret_cast<__nv_bfloat162*>(&hi.x), h1 = *reinterpret_cast<__nv_bfloat162*>(&hi.y);
                lo.x = pack_bf16(x.x - __low2float(h0), x.y - __high2float(h0));
                lo.y = pack_bf16(x.z - __low2float(h1), x.w - __high2float(h1));
              } else {
                hi.x = pack_f16(x.x, x.y); hi.y = pack_f16(x.z, x.w);
                const __half2 h0 = *reinterpret_cast<__half2*>(&hi.x), h1 = *reinterpret_cast<__half2*>(&hi.y);
                lo.x = pack_f16(x.x - __low2float(h0), x.y - __high2float(h0));
                lo.y = pack_f16(x.z - __low2float(h1), x.w - __high2float(h1));
              }
              const uint32_t sub = (uint32_t)(c & 1) * 8u;
              *reinterpret_cast<uint2*>(stage + sw128_offset((uint32_t)rho, (uint32_t)(c >> 1)) + sub) = hi;
              if (p.passes > 1) *reinterpret_cast<uint2*>(stage + sw128_offset((uint32_t)rho, 4u + (uint32_t)(c >> 1)) + sub) = lo;
            } else {
              const float4 hi = make_float4(rna_tf32(x.x), rna_tf32(x.y), rna_tf32(x.z), rna_tf32(x.w));
              const uint32_t off = sw128_offset((uint32_t)rho, (uint32_t)c);
              *reinterpret_cast<float4*>(stage + off) = hi;
              if (OPS == O_TF32X3) {
                const float4 lo = make_float4(rna_tf32(x.x - hi.x), rna_tf32(x.y - hi.y), rna_tf32(x.z - hi.z),
                                              rna_tf32(x.w - hi.w));
                *reinterpret_cast<float4*>(stage + a_tile_bytes + off) = lo;
              }
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (kci == p.n_kc - 1 && ptid == 0) RU_TRACE(tt, 21);
        if (++as == L.a_stages) { as = 0; aph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------- fused ResidualUnit
// y = conv_k1(post1(conv_k7(f(x)) + b1)) + b2 + x  [-> post2]   in ONE kernel (DAC ResidualUnit.cs:24-59):
// the k7 accumulator (TMEM acc1) is drained by the epilogue warps 32 channels at a time, activated, split into
// bf16 hi|lo and written as K-chunk operand tiles ("H stages") in shared memory; the MMA thread multiplies them
// with the 1x1 weights into a second accumulator (acc2); the usual TMA epilogue adds bias, the residual tile (TMA
// load of x) and the following Snake and stores y.  The intermediate never touches HBM.
// C <= 128: both accumulators double-buffered (4 x C <= 512 TMEM columns) and the k7 MMAs of tile i+1 are issued
// BEFORE the 1x1 MMAs of tile i, so draining acc1 overlaps tensor-core work.  128 < C <= 256: single-buffered,
// issue order k7(i), 1x1(i).  Operand mode: bf16x3 / f16x3 only.
constexpr int kHStages = 3;   // most; the launch picks L.h_stages / L.epi_stages (the weight ring gets the rest)

template <int PRO, int RIT>
__global__ void __launch_bounds__(kUmmaThreads, 1)
conv_ru_fused_kernel(const __grid_constant__ ConvGemmParams p, const __grid_constant__ ConvGemmParams p2, const UmmaLaunch L,
                     const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapD,
                     const __grid_constant__ CUtensorMap tmapR) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_stage_bytes = (uint32_t)L.a_rows_alloc * 128u;
  const uint32_t w_stage_bytes = (uint32_t)p.BN * ((p.w_hi_only && p2.w_hi_only) ? 64u : 128u);
  uint8_t* sA = smem;
  uint8_t* sW = sA + (uint32_t)L.a_stages * a_stage_bytes;
  uint8_t* sH = sW + (uint32_t)L.w_stages * w_stage_bytes;
  uint8_t* sE = sH + (uint32_t)L.h_stages * kEpiStageBytes;

  __shared__ uint64_t raw_full[kMaxAStages], a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  __shared__ uint64_t acc1_full[2], acc1_empty[2], acc2_full[2], acc2_empty[2];
  __shared__ uint64_t h_full[kHStages], h_free[kHStages];
  __shared__ uint64_t r_full[kEpiStages], e_free[kEpiStages];
  __shared__ uint32_t tmem_base_s;
  // bias / Snake parameters of both convs (C <= 128): [b1 | a1 | 1/a1 | b2 | a2 | 1/a2][128]; global loads would be L2 round trips
  // on the drain warps' critical path (no L1 under a full shared-memory carve-out)
  __shared__ __align__(16) float s_cf[6 * 128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool cf = p.BN <= 128 && p.act != ACT_TANH && p2.act != ACT_TANH;
  if (cf) {
    for (int i = tid; i < p.BN; i += blockDim.x) {
      s_cf[i] = p.bias ? __ldg(p.bias + i % p.bias_period) : 0.f;
      s_cf[128 + i] = p.post == PRO_SNAKE ? __ldg(p.post_alpha + i % p.post_period) : 0.f;
      s_cf[256 + i] = p.post == PRO_SNAKE ? __ldg(p.post_inv_alpha + i % p.post_period) : 0.f;
      s_cf[384 + i] = p2.bias ? __ldg(p2.bias + i % p2.bias_period) : 0.f;
      s_cf[512 + i] = p2.post == PRO_SNAKE ? __ldg(p2.post_alpha + i % p2.post_period) : 0.f;
      s_cf[640 + i] = p2.post == PRO_SNAKE ? __ldg(p2.post_inv_alpha + i % p2.post_period) : 0.f;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < kMaxAStages; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&a_full[i], kProducerWarps); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], kEpilogueWarps);   // one warp set drains acc1 (see split_a)
      mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], kEpilogueWarps);
    }
    for (int i = 0; i < kHStages; ++i) { mbar_init(&h_full[i], kEpilogueWarps); mbar_init(&h_free[i], 1); }
    for (int i = 0; i < kEpiStages; ++i) { mbar_init(&r_full[i], 1); mbar_init(&e_free[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc<512>(&tmem_base_s); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int total_tiles = p.batch * p.m_tiles_per_clip;      // one N tile (C <= 256)
  const int my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const bool dbl = p.BN <= 128;
  const int n_kc = p.n_kc;                                   // K chunks of the k7 conv == K chunks of the 1x1 conv
  auto acc1_col = [&](int it) { return (uint32_t)(dbl ? (it & 1) * 128 : 0); };
  auto acc2_col = [&](int it) { return (uint32_t)(256 + (dbl ? (it & 1) * 128 : 0)); };
  auto buf_of = [&](int it) { return dbl ? (it & 1) : 0; };
  auto use_of = [&](int it) { return (uint32_t)(dbl ? (it >> 1) : it); };   // how often that buffer was used before

  // E1 (acc1 -> activation -> hi|lo operand tiles of the 1x1 conv, "H stages") is written once and run by whichever
  // warps own it: the 8 transform warps when both accumulators are double-buffered (they are otherwise idle half the
  // time, and the epilogue warps then only do E2), the 8 epilogue warps when single-buffered.  Both sets cover every
  // (TMEM lane quarter, 16-column half) pair exactly once.
  const int q = warp & 3;
  const int half = ((warp >= kFirstProducerWarp ? warp - kFirstProducerWarp : warp - kFirstEpilogueWarp) >> 2) & 1;
  const int rloc = q * 32 + lane;
  const int groups = p.BN / 32;
  int hs = 0;
  uint32_t hph = 0;
  // 32-column groups [0, split_a) are drained by the transform warps, [split_a, groups) by the epilogue warps.
  // Measured (B200, C = 64 / 96 / 128): everything on the transform warps when double-buffered beats both the
  // original all-on-epilogue-warps split and a half/half split (7.1 vs 7.6 vs 7.6 ms at C = 64): the units are bound
  // by the MMA thread's issue interval (~80-100 clk per tcgen05.mma whatever N is, profiles/r02_mma_issue_rate_probe.txt),
  // not by either warp set's throughput; a 4-deep accumulator ring (C <= 64) measured no faster than 2-deep and was dropped.
  const int split_a = (p.BN <= 128) ? groups : 0;
  auto e1 = [&](int it, int g_begin, int g_end) {
    const int b = buf_of(it);
    if (tid == kFirstProducerWarp * 32) RU_TRACE(it, 18);
    mbar_wait(&acc1_full[b], use_of(it) & 1u);
    tc_fence_after();
    if (tid == kFirstProducerWarp * 32) RU_TRACE(it, 10);
    const uint32_t t_addr = tmem_base + acc1_col(it) + ((uint32_t)(q * 32) << 16);
    for (int g = 0; g < groups; ++g) {
      if (g < g_begin || g >= g_end) {   // the other warp set's group: only advance the ring position
        if (++hs == L.h_stages) { hs = 0; hph ^= 1u; }
        continue;
      }
      float v[16];
      __syncwarp();
      tmem_ld16(t_addr + g * 32 + half * 16, v);
      tmem_ld_wait();
      if (g == g_end - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);
      }
      const int n0 = g * 32 + half * 16;
      if (!(L.knock & 64)) {
        if (cf && p.post == PRO_SNAKE) {
          epi_bias_cached(s_cf, v, n0);
          epi_snake_smem(p.precise_sin != 0, s_cf + 128 + n0, s_cf + 256 + n0, v);
        } else {
          epi_bias(p, v, n0);
          epi_post(p, v, n0);
        }
      }
      // 16 channels -> bf16/f16 hi (32 B) and lo (32 B) halves of this row of the K-chunk operand tile
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x0 = v[2 * i], x1 = v[2 * i + 1];
        if (p2.mode == MODE_BF16X3) {   // H is the 1x1 conv's A operand: its format
          hi[i] = pack_bf16(x0, x1);
          const __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&hi[i]);
          lo[i] = pack_bf16(x0 - __low2float(h), x1 - __high2float(h));
        } else {
          hi[i] = pack_f16(x0, x1);
          const __half2 h = *reinterpret_cast<__half2*>(&hi[i]);
          lo[i] = pack_f16(x0 - __low2float(h), x1 - __high2float(h));
        }
      }
      mbar_wait(&h_free[hs], hph ^ 1u);
      uint8_t* stage = sH + (size_t)hs * kEpiStageBytes;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        *reinterpret_cast<uint4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(half * 2 + c))) =
            make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        *reinterpret_cast<uint4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(4 + half * 2 + c))) =
            make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_full[hs]);
      if (g == groups - 1 && tid == kFirstProducerWarp * 32) RU_TRACE(it, 11);
      if (++hs == L.h_stages) { hs = 0; hph ^= 1u; }
    }
  };

  if (warp == 0) {
    // ===================================================================== weight producer (W1 and W2 tiles, MMA order)
    if (elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      const uint32_t w1_bytes = (uint32_t)p.BN * (p.w_hi_only ? 64u : 128u), w2_bytes = (uint32_t)p.BN * (p2.w_hi_only ? 64u : 128u);
      auto push = [&](const float* src, uint32_t bytes) {
        mbar_wait(&w_empty[ws], wph ^ 1u);
        if (L.knock & 128) {   // measurement only: no weight stream from L2
          mbar_arrive(&w_full[ws]);
        } else {
          mbar_arrive_expect_tx(&w_full[ws], bytes);
          bulk_g2s(sW + (size_t)ws * w_stage_bytes, src, bytes, &w_full[ws]);
        }
        if (++ws == L.w_stages) { ws = 0; wph ^= 1u; }
      };
      auto w1 = [&]() {
        for (int kci = 0; kci < n_kc; ++kci)
          for (int j = 0; j < p.n_taps; ++j)
            push(p.W + (size_t)(p.taps[j].tile_base + kci) * (size_t)p.w_tile_floats, w1_bytes);
      };
      auto w2 = [&]() {
        for (int g = 0; g < n_kc; ++g) push(p2.W + (size_t)g * (size_t)p2.w_tile_floats, w2_bytes);
      };
      if (dbl) {
        if (my_tiles > 0) w1();
        for (int it = 0; it < my_tiles; ++it) {
          RU_TRACE(it, 15);
          if (it + 1 < my_tiles) w1();
          RU_TRACE(it, 16);
          w2();
          RU_TRACE(it, 17);
        }
      } else {
        for (int it = 0; it < my_tiles; ++it) { w1(); w2(); }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      // the two convs may use different operand modes (k7: one fp16 product, 1x1: bf16 three-pass split)
      const uint32_t idesc1 = idesc_f16(kBM, p.BN, p.mode == MODE_BF16X3 ? 1 : 0);
      const uint32_t idesc2 = idesc_f16(kBM, p.BN, p2.mode == MODE_BF16X3 ? 1 : 0);
      const uint64_t a_desc0 = desc_at(smem_u32(sA)), w_desc0 = desc_at(smem_u32(sW)), h_desc0 = desc_at(smem_u32(sH));
      // hi-only weight tiles are SWIZZLE_64B images: same address bits, different layout / SBO fields
      const uint64_t w1_fix = p.w_hi_only ? (kDescSw64Base ^ kDescSw128Base) : 0ull;
      const uint64_t w2_fix = p2.w_hi_only ? (kDescSw64Base ^ kDescSw128Base) : 0ull;
      const uint32_t a_stage_u = a_stage_bytes >> 4, w_stage_u = w_stage_bytes >> 4, h_stage_u = kEpiStageBytes >> 4;
      const uint32_t tap_u = (uint32_t)p.dense_step * 8u;
      int ws = 0, as = 0, hs = 0;
      uint32_t wph = 0, aph = 0, hph = 0;
      uint64_t a_desc = a_desc0, w_desc = w_desc0, h_desc = h_desc0;
      auto mma6 = [&](uint32_t d_tmem, uint64_t a0, uint64_t b0, uint32_t acc, int passes, uint32_t idesc) {
        if (L.knock & 8) {   // measurement only: one MMA per accumulator
          if (!acc) umma_f16(d_tmem, a0, b0, idesc, 0);
        } else if (passes == 3) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_f16(d_tmem, a0 + 4 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);  // lo * hi
            umma_f16(d_tmem, a0 + 2 * k, b0 + 4 + 2 * k, idesc, 1);                  // hi * lo
            umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1);                      // hi * hi
          }
        } else if (passes == 2) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_f16(d_tmem, a0 + 4 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);  // lo * hi
            umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, 1);                      // hi * hi
          }
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_f16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, acc | (uint32_t)k);
        }
      };
      auto next_w = [&]() {
        tc_commit(&w_empty[ws]);
        w_desc += w_stage_u;
        if (++ws == L.w_stages) { ws = 0; wph ^= 1u; w_desc = w_desc0; }
      };
      auto m1 = [&](int it) {
        const int b = buf_of(it);
        RU_TRACE(it, 3);
        mbar_wait(&acc1_empty[b], (use_of(it) & 1u) ^ 1u);
        tc_fence_after();
        RU_TRACE(it, 4);
        const uint32_t d_tmem = tmem_base + acc1_col(it);
        uint32_t acc = 0;
        for (int kci = 0; kci < n_kc; ++kci) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          if (kci == 0) RU_TRACE(it, 5);
          uint64_t a_tap = a_desc;
          for (int j = 0; j < p.n_taps; ++j) {
            if (j) a_tap += tap_u;
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            mma6(d_tmem, a_tap, w_desc ^ w1_fix, acc, p.passes, idesc1);
            acc = 1;
            next_w();
          }
          tc_commit(&a_empty[as]);
          a_desc += a_stage_u;
          if (++as == L.a_stages) { as = 0; aph ^= 1u; a_desc = a_desc0; }
        }
        RU_TRACE(it, 6);
        tc_commit(&acc1_full[b]);
      };
      auto m2 = [&](int it) {
        const int b = buf_of(it);
        mbar_wait(&acc2_empty[b], (use_of(it) & 1u) ^ 1u);
        tc_fence_after();
        RU_TRACE(it, 7);
        const uint32_t d_tmem = tmem_base + acc2_col(it);
        for (int g = 0; g < n_kc; ++g) {
          mbar_wait(&h_full[hs], hph);
          if (g == 0) RU_TRACE(it, 8);
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          mma6(d_tmem, h_desc, w_desc ^ w2_fix, g ? 1u : 0u, p2.passes, idesc2);
          next_w();
          tc_commit(&h_free[hs]);
          h_desc += h_stage_u;
          if (++hs == L.h_stages) { hs = 0; hph ^= 1u; h_desc = h_desc0; }
        }
        RU_TRACE(it, 9);
        tc_commit(&acc2_full[b]);
      };
      if (dbl) {
        if (my_tiles > 0) m1(0);
        for (int it = 0; it < my_tiles; ++it) { if (it + 1 < my_tiles) m1(it + 1); m2(it); }
      } else {
        for (int it = 0; it < my_tiles; ++it) { m1(it); m2(it); }
      }
    }
  } else if (warp == kLoaderWarp) {
    // ===================================================================== A loader (TMA)
    if (elect_one()) {
      int as = 0;
      uint32_t aph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int b = tile / p.m_tiles_per_clip;
        const int mt = tile - b * p.m_tiles_per_clip;
        const int r_base = mt * kBM + p.smin;
        for (int kci = 0; kci < n_kc; ++kci) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (kci == 0) RU_TRACE(it, 0);
          if (kci == n_kc - 1) RU_TRACE(it, 1);
          mbar_arrive_expect_tx(&raw_full[as], a_stage_bytes);
          tma_load_3d(sA + (size_t)as * a_stage_bytes, &tmapA, (p.kc_begin + kci) * 32, r_base, b, &raw_full[as]);
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == kResidualWarp) {
    // ===================================================================== residual loader (TMA): tiles of x for E2
    if (elect_one()) {
      const int groups = p.BN / 32;
      int es = 0;
      uint32_t eph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / p.m_tiles_per_clip;
        const int mt = tile - b * p.m_tiles_per_clip;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(&e_free[es], eph ^ 1u);
          if (L.knock & 1024) {   // measurement only: no residual read
            mbar_arrive(&r_full[es]);
          } else {
            mbar_arrive_expect_tx(&r_full[es], kEpiStageBytes);
            tma_load_3d(sE + (size_t)es * kEpiStageBytes, &tmapR, g * 32, mt * kBM, b, &r_full[es]);
          }
          if (++es == L.epi_stages) { es = 0; eph ^= 1u; }
        }
      }
    }
  } else if (warp < kFirstProducerWarp) {
    // ===================================================================== epilogue warps: E2 (acc2 -> y); E1 too when single-buffered
    const bool leader = warp == kFirstEpilogueWarp && lane == 0;
    int es = 0, prev = -1;
    uint32_t eph = 0;
    auto e2 = [&](int it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int bb = tile / p.m_tiles_per_clip;
      const int mt = tile - bb * p.m_tiles_per_clip;
      const int b = buf_of(it);
      if (leader) RU_TRACE(it, 19);
      mbar_wait(&acc2_full[b], use_of(it) & 1u);
      tc_fence_after();
      if (leader) RU_TRACE(it, 12);
      const uint32_t t_addr = tmem_base + acc2_col(it) + ((uint32_t)(q * 32) << 16);
      for (int g = 0; g < groups; ++g) {
        float v[16];
        __syncwarp();
        tmem_ld16(t_addr + g * 32 + half * 16, v);
        tmem_ld_wait();
        if (g == groups - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc2_empty[b]);
        }
        const int n0 = g * 32 + half * 16;
        uint8_t* stage = sE + (size_t)es * kEpiStageBytes;
        mbar_wait(&r_full[es], eph);
        if (leader && g == 0) RU_TRACE(it, 13);
        if (!(L.knock & 256)) {   // (knock 256, measurement only: no bias / residual / activation math)
          if (cf) epi_bias_cached(s_cf + 384, v, n0); else epi_bias(p2, v, n0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 r = *reinterpret_cast<const float4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(half * 4 + i)));
            v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
          }
          if (cf && p2.post == PRO_SNAKE) epi_snake_smem(p2.precise_sin != 0, s_cf + 512 + n0, s_cf + 640 + n0, v);
          else epi_post(p2, v, n0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<float4*>(stage + sw128_offset((uint32_t)rloc, (uint32_t)(half * 4 + i))) =
              make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueWarps * 32) : "memory");
        if (leader) {
          if (!(L.knock & 512)) tma_store_3d(&tmapD, stage, g * 32, mt * kBM, bb);   // (knock 512, measurement only: no store)
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          if (prev >= 0) mbar_arrive(&e_free[prev]);
          prev = es;
          if (g == groups - 1) RU_TRACE(it, 14);
        }
        if (++es == L.epi_stages) { es = 0; eph ^= 1u; }
      }
    };
    if (dbl) {
      for (int it = 0; it < my_tiles; ++it) e2(it);          // E1 runs on the transform warps (split_a == groups)
    } else {
      for (int it = 0; it < my_tiles; ++it) { e1(it, 0, groups); e2(it); }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ===================================================================== A transformers (as conv_umma_kernel, H16X3)
    const int ptid = tid - kFirstProducerWarp * 32;
    const int c = ptid & 7;
    const int rho0 = 8 * (ptid >> 6) + ((ptid >> 5) & 1) + 2 * ((ptid & 31) >> 3);   // see conv_umma_kernel: bank-conflict-free hi / lo stores
    const int rows_needed = kBM + p.span;
    int as = 0, tt = 0;   // tt: tiles transformed so far
    uint32_t aph = 0;
    auto transform_tile = [&]() {
      for (int kci = 0; kci < n_kc; ++kci) {
        if (kci == 0 && ptid == 0) RU_TRACE(tt, 20);
        const int ai = ((p.kc_begin + kci) * 32 + c * 4) % p.alpha_period;
        const float4 al = __ldg(reinterpret_cast<const float4*>(p.alpha + ai));
        const float4 ia = __ldg(reinterpret_cast<const float4*>(p.inv_alpha + ai));
        uint8_t* stage = sA + (size_t)as * a_stage_bytes;
        mbar_wait(&raw_full[as], aph);
        if (kci == 0 && ptid == 0) RU_TRACE(tt, 2);
        if (L.knock & 2) {   // measurement only: operands left as they arrived
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[as]);
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
          continue;
        }
        float4 cur[RIT];
#pragma unroll
        for (int i = 0; i < RIT; ++i) {
          const int rho = rho0 + 32 * i;
          cur[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rho < rows_needed) cur[i] = *reinterpret_cast<const float4*>(stage + sw128_offset((uint32_t)rho, (uint32_t)c));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RIT; ++i) {
          const int rho = rho0 + 32 * i;
          const float4 x = prologue4<PRO>(cur[i], al, ia);
          if (rho < rows_needed) {
            uint2 hi, lo;
            if (p.mode == MODE_BF16X3) {
              hi.x = pack_bf16(x.x, x.y); hi.y = pack_bf16(x.z, x.w);
              const __nv_bfloat162 h0 = *reinterpret_cast<__nv_bfloat162*>(&hi.x), h1 = *reinterpret_cast<__nv_bfloat162*>(&hi.y);
              lo.x = pack_bf16(x.x - __low2float(h0), x.y - __high2float(h0));
              lo.y = pack_bf16(x.z - __low2float(h1), x.w - __high2float(h1));
            } else {
              hi.x = pack_f16(x.x, x.y); hi.y = pack_f16(x.z, x.w);
              const __half2 h0 = *reinterpret_cast<__half2*>(&hi.x), h1 = *reinterpret_cast<__half2*>(&hi.y);
              lo.x = pack_f16(x.x - __low2float(h0), x.y - __high2float(h0));
              lo.y = pack_f16(x.z - __low2float(h1), x.w - __high2float(h1));
            }
            const uint32_t sub = (uint32_t)(c & 1) * 8u;
            *reinterpret_cast<uint2*>(stage + sw128_offset((uint32_t)rho, (uint32_t)(c >> 1)) + sub) = hi;
            *reinterpret_cast<uint2*>(stage + sw128_offset((uint32_t)rho, 4u + (uint32_t)(c >> 1)) + sub) = lo;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[as]);
        if (kci == n_kc - 1 && ptid == 0) RU_TRACE(tt, 21);
        if (++as == L.a_stages) { as = 0; aph ^= 1u; }
      }
      ++tt;
    };
    if (dbl) {
      // operands of tile it+1 first (the MMA thread issues k7(it+1) before 1x1(it)), then drain acc1 of tile it
      if (my_tiles > 0) transform_tile();
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) transform_tile();
        if (split_a > 0) e1(it, 0, split_a);
      }
    } else {
      for (int it = 0; it < my_tiles; ++it) transform_tile();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------- host launcher
size_t umma_smem_bytes(const ConvGemmParams& p, UmmaLaunch* L) {
  const int rows = ((kBM + p.span) + 7) / 8 * 8;
  const int mul = p.mode == MODE_TF32X3 ? 2 : 1;
  const long a_stage = (long)rows * 128 * mul;
  const bool w_hi_only = p.w_hi_only && (p.mode == MODE_BF16X3 || p.mode == MODE_F16X3);
  const long w_stage = (long)p.BN * (w_hi_only ? 64 : 128) * mul;
  // TMA epilogue: whole [rows x n_total] output per clip, 16-byte strides, N tile a multiple of 32 columns
  L->tma_epilogue = (p.BN % 32 == 0 && p.n_total % 4 == 0 && p.d_valid == (long long)p.m_rows * p.n_total &&
                     p.d_clip_stride % 4 == 0 && p.n_valid == p.n_total &&
                     (reinterpret_cast<uintptr_t>(p.D) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.R) & 15) == 0)
                        ? 1 : 0;
  L->a_rows_alloc = rows;
  L->a_stages = L->w_stages = 0;
  L->epi_stages = kEpiStages;
  if (p.span > 64) return 0;
  static const int w_first = std::min(std::max(getenv("NC_W_FIRST") ? atoi(getenv("NC_W_FIRST")) : 4, 2), kMaxWStages);   // tuning knob, clamped
  // at least 2 A stages and 2 W stages; then alternate while both fit (A up to 6, W up to 16):
  // A stages buy bytes in flight from HBM, W stages hide L2 latency of the weight stream
  auto size_rings = [&](long budget, int* as_out, int* ws_out) {
    int as = 2, ws = 2;
    if (as * a_stage + ws * w_stage > budget) return false;
    for (;;) {
      bool grew = false;
      if (ws < w_first && as * a_stage + (ws + 1) * w_stage <= budget) { ++ws; grew = true; }
      if (as < kMaxAStages && (as + 1) * a_stage + ws * w_stage <= budget && (ws >= w_first || as < 3)) { ++as; grew = true; }
      if (!grew) break;
    }
    while (ws < kMaxWStages && as * a_stage + (ws + 1) * w_stage <= budget) ++ws;
    *as_out = as; *ws_out = ws;
    return true;
  };
  const long base = (long)kUmmaMaxDynSmem - 1024 /*alignment slack*/;
  int as = 0, ws = 0;
  // a 5-stage epilogue ring (two-team epilogue with both teams' residual tiles prefetched) for the narrow layers (N <= 64) when
  // the A / W rings still get 4 stages each: their rings would otherwise only soak up the unused shared memory (one-box A/B:
  // Encodec -4.4 %, SNAC -3.5 %; the same preference on the wider DAC layers cost 3.5 %)
  bool ok = false;
  if (L->tma_epilogue && p.BN <= 64 && size_rings(base - (long)kMaxEpiStages * kEpiStageBytes, &as, &ws) && as >= 4 && ws >= 4) {
    L->epi_stages = kMaxEpiStages;
    ok = true;
  }
  if (!ok) {
    const long budget = base - (L->tma_epilogue ? (long)kEpiStages * kEpiStageBytes : 0);
    if (!size_rings(budget, &as, &ws)) return 0;
    // wider layers: only shared memory the rings cannot use anyway (both at their caps)
    const long used = (long)as * a_stage + (((long)ws * w_stage + 1023) & ~1023L);
    if (L->tma_epilogue && used + (long)(kMaxEpiStages - kEpiStages) * kEpiStageBytes <= budget) L->epi_stages = kMaxEpiStages;
  }
  L->a_stages = as;
  L->w_stages = ws;
  return 1024 + (size_t)as * a_stage + (((size_t)ws * w_stage + 1023) & ~(size_t)1023) +
         (L->tma_epilogue ? (size_t)L->epi_stages * kEpiStageBytes : 0);
}

typedef void (*UmmaKernel)(const ConvGemmParams, const UmmaLaunch, const CUtensorMap, const CUtensorMap, const CUtensorMap);

template <int OPS, int PRO>
static UmmaKernel pick_rit(int rit) {
  switch (rit) {
    case 4: return conv_umma_kernel<OPS, PRO, 4>;
    case 5: return conv_umma_kernel<OPS, PRO, 5>;
    default: return conv_umma_kernel<OPS, PRO, 6>;
  }
}
template <int OPS>
static UmmaKernel pick_pro(int pro, int rit) {
  switch (pro) {
    case P_NONE: return pick_rit<OPS, P_NONE>(rit);
    case P_SNAKE_FAST: return pick_rit<OPS, P_SNAKE_FAST>(rit);
    case P_SNAKE_PRECISE: return pick_rit<OPS, P_SNAKE_PRECISE>(rit);
    default: return pick_rit<OPS, P_ELU>(rit);
  }
}
static UmmaKernel pick_kernel(int mode, int pro, int rit) {
  switch (mode) {
    case MODE_TF32: return pick_pro<O_TF32>(pro, rit);
    case MODE_TF32X3: return pick_pro<O_TF32X3>(pro, rit);
    default: return pick_pro<O_H16X3>(pro, rit);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// The A view can be fetched by TMA when every clip is a whole number of rows and strides are 16-byte multiples.
bool umma_view_ok(const ConvGemmParams& p) {
  return p.a_pitch % 4 == 0 && p.a_valid == (long long)p.a_rows * p.a_pitch && p.a_clip_stride % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && p.batch >= 1;
}

// returns cudaError_t as int; 0 on success; -1 if the shape does not fit this kernel
int launch_conv_umma(const ConvGemmParams& p_in, int num_sms, cudaStream_t stream) {
  UmmaLaunch L{};
  const size_t smem = umma_smem_bytes(p_in, &L);
  if (smem == 0 || !umma_view_ok(p_in)) return -1;
  ConvGemmParams p = p_in;
  if (p.acc_split && (!L.tma_epilogue || p.BN > (p.acc_split == 2 ? 256 : 128) ||
                      (p.mode != MODE_BF16X3 && p.mode != MODE_F16X3 && p.mode != MODE_TF32X3) || p.passes != 3 || p.dw_w)) {
    p.acc_split = 0;   // ragged outputs / other operand modes: one accumulator per tile
    p.fold_kc = 0;
  }
  static const int knock = getenv("NC_KNOCK") ? atoi(getenv("NC_KNOCK")) : 0;
  L.knock = knock;
  static const int teams_env = getenv("NC_EPI_TEAMS") ? atoi(getenv("NC_EPI_TEAMS")) : 1;
  L.epi_teams = (teams_env && L.tma_epilogue && !p.acc_split && !p.noise && (!p.R || L.epi_stages == kMaxEpiStages)) ? 1 : 0;
  static const int dual_env = getenv("NC_DUAL_ISSUE") ? atoi(getenv("NC_DUAL_ISSUE")) : 0;   // opt-in: measured neutral to -2 % (job AM)
  L.dual_issue = (dual_env && L.tma_epilogue && !p.R && !p.acc_split && p.dense_step >= 0 && p.n_tiles == 1 && !p.dw_w && p.BN <= 128 &&
                  L.w_stages >= p.n_kc * p.n_taps + 2 && L.a_stages >= p.n_kc + 1) ? 1 : 0;
  const int rows_needed = kBM + p.span;
  const int rit = (rows_needed + 31) / 32;
  if (rit > 6) return -1;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  alignas(64) CUtensorMap tmap, tmapD, tmapR;
  std::memset(&tmapD, 0, sizeof tmapD);
  std::memset(&tmapR, 0, sizeof tmapR);
  if (L.tma_epilogue) {
    const cuuint64_t dd[3] = {(cuuint64_t)p.n_total, (cuuint64_t)p.m_rows, (cuuint64_t)p.batch};
    const cuuint64_t ds[2] = {(cuuint64_t)p.n_total * 4, (cuuint64_t)p.d_clip_stride * 4};
    const cuuint32_t db[3] = {32, (cuuint32_t)kBM, 1};
    const cuuint32_t de[3] = {1, 1, 1};
    if (enc(&tmapD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.D, dd, ds, db, de, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
    if (p.R && enc(&tmapR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.R), dd, ds, db, de,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  const cuuint64_t gdim[3] = {(cuuint64_t)p.a_pitch, (cuuint64_t)p.a_rows, (cuuint64_t)p.batch};
  const cuuint64_t gstr[2] = {(cuuint64_t)p.a_pitch * 4, (cuuint64_t)p.a_clip_stride * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)L.a_rows_alloc, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.A), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  int pro = P_NONE;
  if (p.prologue == PRO_SNAKE) pro = p.precise_sin ? P_SNAKE_PRECISE : P_SNAKE_FAST;
  else if (p.prologue == PRO_ELU) pro = P_ELU;
  UmmaKernel k = pick_kernel(p.mode, pro, rit < 4 ? 4 : rit);
  // opt-in shared memory: idempotent, set on every launch (handles may live on any device)
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  const int total_tiles = p.n_tiles * p.batch * p.m_tiles_per_clip;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (grid <= 0) return 0;
  // measurement only: NC_TRACE_UMMA=<file> NC_TRACE_UMMA_MATCH=<BN>,<a_pitch>,<n_taps> [NC_TRACE_UMMA_SKIP=n] dumps CTA 0's role
  // timeline of one matching launch (scripts/ru_trace_analyze.py)
  static const char* trace_path = getenv("NC_TRACE_UMMA");
  static int t_bn = 0, t_pitch = 0, t_taps = 0, t_skip = 3;
  static const bool trace_cfg = trace_path && getenv("NC_TRACE_UMMA_MATCH") &&
                                std::sscanf(getenv("NC_TRACE_UMMA_MATCH"), "%d,%d,%d", &t_bn, &t_pitch, &t_taps) == 3 &&
                                ((t_skip = getenv("NC_TRACE_UMMA_SKIP") ? atoi(getenv("NC_TRACE_UMMA_SKIP")) : 3), true);
  const bool tracing = trace_cfg && p.BN == t_bn && p.a_pitch == t_pitch && p.n_taps == t_taps && t_skip-- == 0;
  if (tracing) {
    if (cudaMalloc(&L.trace, sizeof(unsigned long long) * kTraceTiles * kTraceEvents) != cudaSuccess) return (int)cudaErrorMemoryAllocation;
    cudaMemsetAsync(L.trace, 0, sizeof(unsigned long long) * kTraceTiles * kTraceEvents, stream);
  }
  k<<<grid, kUmmaThreads, smem, stream>>>(p, L, tmap, tmapD, tmapR);
  if (tracing) {
    std::vector<unsigned long long> h((size_t)kTraceTiles * kTraceEvents);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), L.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(L.trace);
    if (FILE* f = std::fopen(trace_path, "w")) {
      std::fprintf(f, "# conv_umma BN=%d n_tiles=%d n_kc=%d span=%d taps=%d mode=%d passes=%d acc_split=%d dw=%d residual=%d a_stages=%d w_stages=%d grid=%d tiles=%d\n",
                   p.BN, p.n_tiles, p.n_kc, p.span, p.n_taps, p.mode, p.passes, p.acc_split, p.dw_w ? 1 : 0, p.R ? 1 : 0, L.a_stages, L.w_stages, grid,
                   total_tiles);
      for (int t = 0; t < kTraceTiles; ++t) {
        for (int e = 0; e < 28; ++e) std::fprintf(f, "%llu ", h[(size_t)t * kTraceEvents + e]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
    }
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------- fused RU launcher
typedef void (*RuKernel)(const ConvGemmParams, const ConvGemmParams, const UmmaLaunch, const CUtensorMap, const CUtensorMap,
                         const CUtensorMap);

// Measured (profiles/r01_layers_dac_b16x30s_fusedru.txt): fusion wins while both accumulators can be double-buffered in
// TMEM (C <= 128); the single-buffered variant (128 < C <= 256) loses to two launches, so it is off by default.
static int ru_fuse_wide_max_c() {
  static const int v = getenv("NC_RU_FUSE_WIDE_MAX_C") ? atoi(getenv("NC_RU_FUSE_WIDE_MAX_C")) : 256;
  return v > 256 ? 256 : v;
}
static int ru_fuse_max_c() {
  static const int v = getenv("NC_RU_FUSE_MAX_C") ? atoi(getenv("NC_RU_FUSE_MAX_C")) : 128;
  return v;
}

bool ru_fused_supported(const ConvGemmParams& p, const ConvGemmParams& p2) {
  const bool h16 = (p.mode == MODE_BF16X3 || p.mode == MODE_F16X3) && (p2.mode == MODE_BF16X3 || p2.mode == MODE_F16X3);
  // C <= 128: both accumulators double-buffered.  128 < C <= 256: one tile in flight (acc1 + acc2 fill TMEM); the k7
  // MMAs are then exposed to the acc1 drain, which only pays when they are short: one-pass fp16 k7 convs only.
  const bool width_ok = p.BN <= ru_fuse_max_c() || (p.BN <= ru_fuse_wide_max_c() && p.passes == 1);
  return h16 && width_ok && !p.acc_split && !p2.acc_split && p.n_tiles == 1 && p2.n_tiles == 1 &&
         p.BN == p2.BN && p.BN % 32 == 0 && p.n_total == p.BN && p.n_valid == p.BN && p.dense_step >= 0 &&
         p.prologue == PRO_SNAKE && p2.n_taps == 1 && p2.n_kc == p.n_kc && p.kc_begin == 0 && p.span <= 64 &&
         p.a_pitch == p.BN && umma_view_ok(p) && p.d_valid == (long long)p.m_rows * p.n_total && p.d_clip_stride % 4 == 0 &&
         !p.noise && (reinterpret_cast<uintptr_t>(p.D) & 15) == 0;
}

// p: the k7 conv's plan with p.D = final output y, p.R = x (the unit's input), p.post = Snake2;
// p2: the 1x1 conv's plan (weights, bias, post = the Snake that follows the unit or none).
int launch_ru_fused(const ConvGemmParams& p, const ConvGemmParams& p2, int num_sms, cudaStream_t stream) {
  if (!ru_fused_supported(p, p2)) return -1;
  UmmaLaunch L{};
  const int rows = ((kBM + p.span) + 7) / 8 * 8;
  const long a_stage = (long)rows * 128, w_stage = (long)p.BN * ((p.w_hi_only && p2.w_hi_only) ? 64 : 128);
  // Shared-memory split (sweep: profiles/r02_ru_fused_smem_split.txt).  These units are bound by the issue interval of their single
  // MMA thread (~80-100 clk per tcgen05.mma whatever N is: profiles/r02_mma_issue_rate_probe.txt), so no ring depth changes much;
  // the best measured split keeps 3 H stages, 2 E stages, an A ring of one tile (+1 chunk at C = 64) and gives the rest to the
  // weight ring (-6 % over the round-1 split of 4 / 3 / 3).
  static const int env_as = getenv("NC_RU_AS") ? atoi(getenv("NC_RU_AS")) : 0;
  static const int env_hs = getenv("NC_RU_HS") ? atoi(getenv("NC_RU_HS")) : 0;
  static const int env_es = getenv("NC_RU_ES") ? atoi(getenv("NC_RU_ES")) : 0;
  const int hs = std::min(std::max(env_hs ? env_hs : 3, 2), kHStages), es = std::min(std::max(env_es ? env_es : 2, 2), kEpiStages);
  const long budget = (long)kUmmaMaxDynSmem - 1024 - (long)(hs + es) * kEpiStageBytes;
  int as = env_as ? std::min(std::max(env_as, 2), kMaxAStages) : std::min(std::max(p.n_kc + 1, 3), 4), ws = 2;
  while (as > 2 && as * a_stage + ws * w_stage > budget) --as;
  if (as * a_stage + ws * w_stage > budget) return -1;
  while (ws < kMaxWStages && as * a_stage + (ws + 1) * w_stage <= budget) ++ws;
  static const int knock_ru = getenv("NC_KNOCK_RU") ? atoi(getenv("NC_KNOCK_RU")) : 0;
  L.a_stages = as; L.w_stages = ws; L.a_rows_alloc = rows; L.tma_epilogue = 1; L.knock = knock_ru; L.epi_stages = es; L.h_stages = hs;
  const size_t smem = 1024 + (size_t)as * a_stage + (size_t)ws * w_stage + (size_t)(hs + es) * kEpiStageBytes;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  alignas(64) CUtensorMap tA, tD, tR;
  const cuuint32_t estr[3] = {1, 1, 1};
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)p.a_pitch, (cuuint64_t)p.a_rows, (cuuint64_t)p.batch};
    const cuuint64_t gstr[2] = {(cuuint64_t)p.a_pitch * 4, (cuuint64_t)p.a_clip_stride * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)rows, 1};
    if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.A), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  {
    const cuuint64_t dd[3] = {(cuuint64_t)p.n_total, (cuuint64_t)p.m_rows, (cuuint64_t)p.batch};
    const cuuint64_t ds[2] = {(cuuint64_t)p.n_total * 4, (cuuint64_t)p.d_clip_stride * 4};
    const cuuint32_t db[3] = {32, (cuuint32_t)kBM, 1};
    if (enc(&tD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.D, dd, ds, db, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
    const cuuint64_t rs[2] = {(cuuint64_t)p.a_pitch * 4, (cuuint64_t)p.a_clip_stride * 4};
    if (enc(&tR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.A), dd, rs, db, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  const int rit = (kBM + p.span + 31) / 32;
  RuKernel k;
  if (p.precise_sin) k = rit <= 5 ? conv_ru_fused_kernel<P_SNAKE_PRECISE, 5> : conv_ru_fused_kernel<P_SNAKE_PRECISE, 6>;
  else k = rit <= 5 ? conv_ru_fused_kernel<P_SNAKE_FAST, 5> : conv_ru_fused_kernel<P_SNAKE_FAST, 6>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaMaxDynSmem);
  if (e != cudaSuccess) return (int)e;
  const int total_tiles = p.batch * p.m_tiles_per_clip;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (grid <= 0) return 0;
  // measurement only: NC_TRACE_RU=<file> [NC_TRACE_RU_C=<channels>] dumps CTA 0's role timeline of one launch
  static const char* trace_path = getenv("NC_TRACE_RU");
  static const int trace_c = getenv("NC_TRACE_RU_C") ? atoi(getenv("NC_TRACE_RU_C")) : 64;
  static int trace_skip = getenv("NC_TRACE_RU_SKIP") ? atoi(getenv("NC_TRACE_RU_SKIP")) : 3;   // warm launches first
  L.trace = nullptr;
  const bool tracing = trace_path && p.BN == trace_c && trace_skip-- == 0;
  if (tracing) {
    if (cudaMalloc(&L.trace, sizeof(unsigned long long) * kTraceTiles * kTraceEvents) != cudaSuccess) return (int)cudaErrorMemoryAllocation;
    cudaMemsetAsync(L.trace, 0, sizeof(unsigned long long) * kTraceTiles * kTraceEvents, stream);
  }
  k<<<grid, kUmmaThreads, smem, stream>>>(p, p2, L, tA, tD, tR);
  if (tracing) {
    std::vector<unsigned long long> h((size_t)kTraceTiles * kTraceEvents);
    cudaStreamSynchronize(stream);
    cudaMemcpy(h.data(), L.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(L.trace);
    if (FILE* f = std::fopen(trace_path, "w")) {
      std::fprintf(f, "# BN=%d n_kc=%d span=%d taps=%d a_stages=%d w_stages=%d h_stages=%d e_stages=%d grid=%d tiles=%d\n", p.BN, p.n_kc, p.span,
                   p.n_taps, L.a_stages, L.w_stages, L.h_stages, L.epi_stages, grid, total_tiles);
      for (int t = 0; t < kTraceTiles; ++t) {
        for (int e = 0; e < 28; ++e) std::fprintf(f, "%llu ", h[(size_t)t * kTraceEvents + e]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
    }
  }
  return (int)cudaGetLastError();
}

}  // namespace nc
