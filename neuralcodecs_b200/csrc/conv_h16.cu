// fp16-operand tcgen05 executor of the multi-tap row-shifted GEMM (conv_plan.h) for the layers that take ONE
// fp16 product per term (the DAC decoder's wide layers: decoder.conv1, the transposed convs and the residual units
// with more than 128 channels).
//
// Why a second kernel (measured on B200, profiles/r02_knockout_*.txt): with one MMA per product the fp32-in /
// transform-in-smem kernel (conv_umma.cu) is no longer bound by the tensor pipe but by (i) the per-N-tile fp32->fp16
// operand transform, (ii) mbarrier round trips -- one weight-stage handshake (~270 clk) per two 64-clk MMAs -- and
// (iii) the L2->smem weight stream (a 128-row tile re-streams the layer's whole weight set: 64 B/clk/SM at full MMA
// rate against ~42 B/clk/SM of L2 bandwidth).  Here
//   * activations between these layers live in HBM as fp16 in the consumer's operand format (the producing layer's
//     epilogue applies the consumer's Snake and rounds once), so an A tile goes TMA -> smem -> tcgen05.mma with no
//     transform warps, half the bytes and 64 channels (4 MMAs per tap) per stage;
//   * a residual reader gets the raw fp32 value from a second TMA store of the same epilogue ("dual output");
//   * two CTAs of a cluster pair up on a 256-row tile (tcgen05.mma.cta_group::2): each CTA stages only half of every
//     weight tile, which halves both the L2 weight stream and the B-operand shared-memory reads per MMA.
//
// Warp roles (384 threads): 0 weight producer (TMA 2-D boxes of the pre-tiled weights), 1 MMA issuer (leader CTA of a
// pair only), 2-9 epilogue (tcgen05.ld -> bias, fp32 residual, raw fp32 store and/or Snake -> fp16 store, all through
// swizzled smem rings and TMA), 10 A loader (TMA 3-D, zero fill = conv padding), 11 residual loader.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "conv_plan.h"
#include "umma.cuh"

namespace nc {

using namespace ptx;

namespace {

constexpr int kBM = 128;
constexpr int kEpiWarps = 8;
constexpr int kFirstEpiWarp = 2;
constexpr int kLoaderWarp = kFirstEpiWarp + kEpiWarps;   // 10
constexpr int kResidualWarp = kLoaderWarp + 1;           // 11
constexpr int kThreads = 32 * (kResidualWarp + 1);       // 384
constexpr int kMaxA = 6, kMaxW = 8, kMaxEpi = 6;
// one epilogue stage = 64 output columns = two TMA boxes of 32 columns (a swizzled box is at most 128 B wide in fp32)
constexpr uint32_t kBox32 = kBM * 128;                   // [128 rows][32 fp32]
constexpr uint32_t kBox16 = kBM * 64;                    // [128 rows][32 fp16]
constexpr uint32_t kStage32 = 2 * kBox32;
constexpr uint32_t kStage16 = 2 * kBox16;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
constexpr size_t kMaxDynSmem = 227 * 1024 - 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Arrivals on a barrier given by its shared::cluster address (own CTA or the pair leader).  Default semantics (release at
// CTA scope), as CUTLASS' ClusterBarrier does: an explicit .release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR, ~500
// clocks per pipeline step (measured with ncu: the weight producer spent its time there, profiles/r02_h16_membar.txt).
// What these arrivals order is tensor-core / TMA traffic, which tcgen05.fence / the mbarrier's complete_tx already cover.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}

template <bool kPair>
__device__ __forceinline__ void tmem_alloc512(uint32_t* dst) {
  if (kPair)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(dst)) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(dst)) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void tmem_relinquish_p() {
  if (kPair) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <bool kPair>
__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr) {
  if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (kPair)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit: arrive on the barrier at this smem offset in this CTA (single) or in both CTAs of the pair
template <bool kPair>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (kPair)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMA loads whose completion lands on the barrier `bar_addr` (a shared::cluster address: own CTA, or the pair leader)
template <bool kPair>
__device__ __forceinline__ void tma_load3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar_addr) {
  if (kPair)
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_addr)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_addr)
        : "memory");
}
template <bool kPair>
__device__ __forceinline__ void tma_load2(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_addr) {
  if (kPair)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_addr)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_addr)
        : "memory");
}
__device__ __forceinline__ void tma_store3a(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ float sin_abs_precise16(float x) {   // as conv_umma.cu: |sin| by Cody-Waite + minimax polynomials
  const int q = __float2int_rn(x * 0.636619772f);
  const float j = __int2float_rn(q);
  float r = fmaf(j, -1.57079601e+00f, x);
  r = fmaf(j, -3.13916473e-07f, r);
  r = fmaf(j, -5.39030253e-15f, r);
  const float s = r * r;
  const bool odd = (q & 1) != 0;
  float pl = odd ? 2.44331571e-5f : -1.95152959e-4f;
  pl = fmaf(pl, s, odd ? -1.38873163e-3f : 8.33216087e-3f);
  pl = fmaf(pl, s, odd ? 4.16666457e-2f : -1.66666546e-1f);
  const float a = odd ? fmaf(pl, s, -0.5f) : pl;
  const float m = odd ? s : r * s;
  const float b = odd ? 1.0f : r;
  return fmaf(a, m, b);
}
__device__ __forceinline__ float snake16(float x, float a, float ia, bool precise) {
  const float t = a * x;
  const float s = precise ? sin_abs_precise16(t) : __sinf(t);
  return fmaf(s * s, ia, x);   // alpha == 0 -> ia == 0 -> x  (Modules/DAC/Snake1d.cs:49-58)
}

}  // namespace

// p.A = fp16 activations (as const float*), p.a_pitch / a_clip_stride in HALVES; taps' kc_* and n_kc / kc_begin count
// 64-channel chunks; p.D = raw fp32 output (nullable), p.D16 = activated fp16 output (nullable); p.R fp32 residual.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
conv_h16_kernel(const __grid_constant__ ConvGemmParams p, const UmmaLaunch L, const __grid_constant__ CUtensorMap tmapA,
                const __grid_constant__ CUtensorMap tmapW, const __grid_constant__ CUtensorMap tmapD32,
                const __grid_constant__ CUtensorMap tmapD16, const __grid_constant__ CUtensorMap tmapR) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t kCtas = kPair ? 2u : 1u;
  const uint32_t a_stage_bytes = (uint32_t)L.a_rows_alloc * 128u;        // [rows][64 fp16]
  const uint32_t w_rows = (uint32_t)p.BN / kCtas;                        // weight rows this CTA stages per tile
  const uint32_t w_stage_bytes = w_rows * 128u;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (uint32_t)L.a_stages * a_stage_bytes;
  // epilogue ring: E stages of an optional fp32 part (residual in / raw value out) and an optional fp16 part
  const int E = L.epi_stages;
  uint8_t* sE32 = sW + (((uint32_t)L.w_stages * w_stage_bytes + 1023u) & ~1023u);
  uint8_t* sE16 = sE32 + ((p.R || p.D) ? (uint32_t)E * kStage32 : 0u);

  __shared__ uint64_t a_full[kMaxA], a_empty[kMaxA], w_full[kMaxW], w_empty[kMaxW];
  __shared__ uint64_t acc_full[2], acc_empty[2], r_full[kMaxEpi], e_free[kMaxEpi];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (tid == 0) {
    for (int i = 0; i < kMaxA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kMaxW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps * kCtas); }
    for (int i = 0; i < kMaxEpi; ++i) { mbar_init(&r_full[i], 1); mbar_init(&e_free[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc512<kPair>(&tmem_base_s);
    tmem_relinquish_p<kPair>();
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // tiles: N tile fastest, then (pair-)M tile, then clip.  A pair covers rows [256*mp, 256*mp + 256) of a clip.
  const int m_units = kPair ? (p.m_tiles_per_clip + 1) / 2 : p.m_tiles_per_clip;
  const int total_tiles = p.n_tiles * p.batch * m_units;
  const int first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto decode = [&](int tile, int* nt, int* b, int* mt) {
    *nt = tile % p.n_tiles;
    const int rem = tile / p.n_tiles;
    *b = rem / m_units;
    const int mu = rem - *b * m_units;
    *mt = kPair ? 2 * mu + (int)rank : mu;      // this CTA's 128-row tile (may lie past the clip's end: all rows OOB)
  };
  // barriers the loaders signal: the leader's (the MMA issuer waits there)
  auto leader_bar = [&](uint64_t* bar) { return kPair ? map_to_cta(smem_u32(bar), 0u) : smem_u32(bar); };

  if (warp == 0) {
    // ===================================================================== weight producer
    if (elect_one()) {
      int ws = 0, wcount = 0;
      uint32_t wph = 0;
      for (int tile = first; tile < total_tiles; tile += step) {
        int nt, b, mt;
        decode(tile, &nt, &b, &mt);
        const unsigned mask = p.tap_mask[nt];
        const int row0 = nt * p.tiles_per_ntile;          // tile index base inside the weight array
        for (int kci = 0; kci < p.n_kc; ++kci) {
          const int kc = p.kc_begin + kci;
          for (int j = 0; j < p.n_taps; ++j) {
            if (!((mask >> j) & 1u) || kc < p.taps[j].kc_lo || kc >= p.taps[j].kc_hi) continue;
            mbar_wait(&w_empty[ws], wph ^ 1u);
            const int t_idx = row0 + p.taps[j].tile_base + (kc - p.taps[j].kc_lo);
            const uint32_t bar = leader_bar(&w_full[ws]);
            if ((L.knock & 1) && wcount >= L.w_stages) {   // measurement only: no weight traffic after the ring filled once
              if (leader) mbar_arrive_cluster(bar);
            } else {
              if (leader) mbar_expect_tx_cluster(bar, w_stage_bytes * kCtas);
              tma_load2<kPair>(smem_u32(sW + (size_t)ws * w_stage_bytes), &tmapW, 0, t_idx * p.BN + (int)(rank * w_rows), bar);
            }
            ++wcount;
            if (++ws == L.w_stages) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (pair: leader CTA only)
    if (leader && elect_one()) {
      const uint32_t idesc = idesc_f16(kPair ? 256 : 128, p.BN, 0);
      const uint64_t a_desc0 = desc_at(smem_u32(sA)), w_desc0 = desc_at(smem_u32(sW));
      const uint32_t a_stage_u = a_stage_bytes >> 4, w_stage_u = w_stage_bytes >> 4;
      const uint32_t tap_u = (uint32_t)p.dense_step * 8u;
      int ws = 0, as = 0, it = 0;
      uint32_t wph = 0, aph = 0;
      uint64_t a_desc = a_desc0, w_desc = w_desc0;
      for (int tile = first; tile < total_tiles; tile += step, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256u;
        uint32_t acc = 0;
        const unsigned mask = p.dense_step >= 0 ? 0u : (unsigned)p.tap_mask[tile % p.n_tiles];
        for (int kci = 0; kci < p.n_kc; ++kci) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          uint64_t a_tap = a_desc;
          for (int j = 0; j < p.n_taps; ++j) {
            if (p.dense_step >= 0) {
              if (j) a_tap += tap_u;
            } else {
              const int kc = p.kc_begin + kci;
              if (!((mask >> j) & 1u) || kc < p.taps[j].kc_lo || kc >= p.taps[j].kc_hi) continue;
              a_tap = a_desc + (uint32_t)(p.taps[j].shift - p.smin) * 8u;
            }
            mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            if (L.knock & 8) {   // measurement only: one MMA per tile
              if (!acc) mma_f16<kPair>(d_tmem, a_tap, w_desc, idesc, 0);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)   // K step = 16 halves = 32 B = +2 descriptor units
                mma_f16<kPair>(d_tmem, a_tap + 2 * k, w_desc + 2 * k, idesc, acc | (uint32_t)k);
            }
            commit<kPair>(&w_empty[ws]);
            acc = 1;
            w_desc += w_stage_u;
            if (++ws == L.w_stages) { ws = 0; wph ^= 1u; w_desc = w_desc0; }
          }
          commit<kPair>(&a_empty[as]);
          a_desc += a_stage_u;
          if (++as == L.a_stages) { as = 0; aph ^= 1u; a_desc = a_desc0; }
        }
        commit<kPair>(&acc_full[buf]);
      }
    }
  } else if (warp == kLoaderWarp) {
    // ===================================================================== A loader (TMA, fp16 rows + halo)
    if (elect_one()) {
      int as = 0, acount = 0;
      uint32_t aph = 0;
      for (int tile = first; tile < total_tiles; tile += step) {
        int nt, b, mt;
        decode(tile, &nt, &b, &mt);
        const int r_base = mt * kBM + p.smin;
        for (int kci = 0; kci < p.n_kc; ++kci) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          const uint32_t bar = leader_bar(&a_full[as]);
          if ((L.knock & 4) && acount >= L.a_stages) {
            if (leader) mbar_arrive_cluster(bar);
          } else {
            if (leader) mbar_expect_tx_cluster(bar, a_stage_bytes * kCtas);
            tma_load3<kPair>(smem_u32(sA + (size_t)as * a_stage_bytes), &tmapA, (p.kc_begin + kci) * 64, r_base, b, bar);
          }
          ++acount;
          if (++as == L.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == kResidualWarp) {
    // ===================================================================== residual loader (fp32 tiles of R)
    if (p.R && !(L.knock & 16) && elect_one()) {
      const int groups = p.BN / 64;
      int es = 0;
      uint32_t eph = 0;
      for (int tile = first; tile < total_tiles; tile += step) {
        int nt, b, mt;
        decode(tile, &nt, &b, &mt);
        for (int g = 0; g < groups; ++g) {
          mbar_wait(&e_free[es], eph ^ 1u);
          if (L.knock & 64) {   // measurement only: no residual traffic
            mbar_arrive(&r_full[es]);
          } else {
            mbar_arrive_expect_tx(&r_full[es], kStage32);
            const uint32_t dst = smem_u32(sE32 + (size_t)es * kStage32);
            tma_load3<false>(dst, &tmapR, nt * p.BN + g * 64, mt * kBM, b, smem_u32(&r_full[es]));
            tma_load3<false>(dst + kBox32, &tmapR, nt * p.BN + g * 64 + 32, mt * kBM, b, smem_u32(&r_full[es]));
          }
          if (++es == E) { es = 0; eph ^= 1u; }
        }
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kLoaderWarp) {
    // ===================================================================== epilogue
    const int q = warp & 3;                              // TMEM lane quarter this warp may touch
    const int hb = (warp - kFirstEpiWarp) >> 2;          // which 32-column box of a 64-column stage this warp handles
    const uint32_t rloc = (uint32_t)(q * 32 + lane);
    const int groups = p.BN / 64;
    const bool store_leader = warp == kFirstEpiWarp && lane == 0;
    const bool precise = p.precise_sin != 0;
    const uint32_t acc_empty_leader = leader_bar(&acc_empty[0]);
    const uint32_t e32 = smem_u32(sE32) + (uint32_t)hb * kBox32, e16 = smem_u32(sE16) + (uint32_t)hb * kBox16;
    int es = 0, it = 0;
    long long gcount = 0;   // stages stored so far
    uint32_t eph = 0;
    for (int tile = first; tile < total_tiles; tile += step, ++it) {
      int nt, b, mt;
      decode(tile, &nt, &b, &mt);
      const int buf = it & 1;
      mbar_wait(&acc_full[buf], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)buf * 256u + ((uint32_t)(q * 32) << 16);
      if (L.knock & 16) {   // measurement only: no epilogue work at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)buf * 8u);
        continue;
      }
      for (int g = 0; g < groups; ++g) {
        // 32 accumulator columns of this thread's row: two 16-column TMEM loads in flight together
        float v[32];
        __syncwarp();
        tmem_ld16(t_addr + g * 64 + hb * 32, v);
        tmem_ld16(t_addr + g * 64 + hb * 32 + 16, v + 16);
        const int n0 = nt * p.BN + g * 64 + hb * 32;
        float4 bb[8];
        if (p.bias) {
          const int bi = n0 % p.bias_period;   // bias_period % 32 == 0 (checked by the launcher)
#pragma unroll
          for (int i = 0; i < 8; ++i) bb[i] = __ldg(reinterpret_cast<const float4*>(p.bias + bi) + i);
        }
        tmem_ld_wait();
        if (g == groups - 1) {   // accumulator fully read by this warp: hand the TMEM buffer back (to the leader's MMA thread)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)buf * 8u);
        }
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[4 * i] += bb[i].x; v[4 * i + 1] += bb[i].y; v[4 * i + 2] += bb[i].z; v[4 * i + 3] += bb[i].w; }
        }
        const uint32_t st32 = e32 + (uint32_t)es * kStage32, st16 = e16 + (uint32_t)es * kStage16;
        if (p.R) mbar_wait(&r_full[es], eph); else mbar_wait(&e_free[es], eph ^ 1u);
        if (p.R) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 r = lds128(st32 + sw128_offset(rloc, (uint32_t)i));
            v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
          }
        }
        if (p.D) {   // raw fp32 value (bias + residual) for a later residual reader
#pragma unroll
          for (int i = 0; i < 8; ++i)
            sts128(st32 + sw128_offset(rloc, (uint32_t)i), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
        }
        if (p.D16) {   // the consumer's activation, rounded once to its fp16 operand format
          if (p.post == PRO_SNAKE) {
            const int pi = n0 % p.post_period;   // post_period % 32 == 0
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 al = __ldg(reinterpret_cast<const float4*>(p.post_alpha + pi) + i);
              const float4 ia = __ldg(reinterpret_cast<const float4*>(p.post_inv_alpha + pi) + i);
              v[4 * i + 0] = snake16(v[4 * i + 0], al.x, ia.x, precise); v[4 * i + 1] = snake16(v[4 * i + 1], al.y, ia.y, precise);
              v[4 * i + 2] = snake16(v[4 * i + 2], al.z, ia.z, precise); v[4 * i + 3] = snake16(v[4 * i + 3], al.w, ia.w, precise);
            }
          }
          uint32_t h[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __half2 hv = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            h[i] = *reinterpret_cast<uint32_t*>(&hv);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)   // the row's 64-byte line of this box: four 16-byte chunks
            sts128u(st16 + sw64_offset(rloc, (uint32_t)c), h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (store_leader) {
          const int c0 = nt * p.BN + g * 64, r0 = mt * kBM;
          const uint32_t s32 = smem_u32(sE32) + (uint32_t)es * kStage32, s16 = smem_u32(sE16) + (uint32_t)es * kStage16;
          if (p.D && !(L.knock & 32)) { tma_store3a(&tmapD32, s32, c0, r0, b); tma_store3a(&tmapD32, s32 + kBox32, c0 + 32, r0, b); }
          if (p.D16 && !(L.knock & 32)) { tma_store3a(&tmapD16, s16, c0, r0, b); tma_store3a(&tmapD16, s16 + kBox16, c0 + 32, r0, b); }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          // every store group but the newest has finished reading smem: hand the PREVIOUS stage back right away, so the
          // residual loader can prefetch E-1 stages ahead (freeing a stage only when the next one needs it would cut the
          // prefetch distance to one stage: measured 2.5 -> 1.5 ms on the C = 384 1x1 convs)
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          if (gcount > 0) mbar_arrive(&e_free[es == 0 ? E - 1 : es - 1]);
        }
        ++gcount;
        if (++es == E) { es = 0; eph ^= 1u; }
      }
    }
    if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc512<kPair>(tmem_base);
}

// ------------------------------------------------------------------------------- host launcher
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

bool h16_supported(const ConvGemmParams& p) {
  return p.a16 && p.BN % 64 == 0 && p.BN <= 256 && p.n_total % 8 == 0 && p.n_valid == p.n_total &&
         p.d_valid == (long long)p.m_rows * p.n_total && p.d_clip_stride % 8 == 0 && p.a_pitch % 64 == 0 &&
         p.a_valid == (long long)p.a_rows * p.a_pitch && p.a_clip_stride % 8 == 0 && p.span <= 64 &&
         (!p.bias || p.bias_period % 32 == 0) && (p.post != PRO_SNAKE || p.post_period % 32 == 0) && p.post != PRO_ELU &&
         p.act == ACT_NONE && !p.noise && (p.D || p.D16) && (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.D) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.D16) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.R) & 15) == 0;
}

// returns 0 on success, a cudaError_t (> 0), or -1 when the plan does not fit this kernel
int launch_conv_h16(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (!h16_supported(p)) return -1;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  static const int pair_env = getenv("NC_H16_PAIR") ? atoi(getenv("NC_H16_PAIR")) : 1;
  // a pair needs an even N split whose halves keep the 8-row swizzle period and the UMMA N granularity of 16
  const bool pair = pair_env != 0 && p.m_tiles_per_clip * p.batch >= 2;
  UmmaLaunch L{};
  const int rows = ((kBM + p.span) + 7) / 8 * 8;
  const long a_stage = (long)rows * 128;
  const long w_stage = (long)(p.BN / (pair ? 2 : 1)) * 128;
  const long e_stage = ((p.R || p.D) ? (long)kStage32 : 0) + (p.D16 ? (long)kStage16 : 0);
  const long budget = (long)kMaxDynSmem - 2048;
  // Shared-memory split.  Layers with little MMA work per output element (1x1 convs: one tap) are bound by their
  // epilogue streams (fp32 residual in, fp32 + fp16 out), so the epilogue ring gets the depth (bytes in flight);
  // multi-tap layers put the memory into A / weight stages and keep a short ring.
  const bool light = p.n_taps == 1;
  int E = light ? 3 : 2;
  int as = 2, ws = 2;
  while (E > 2 && as * a_stage + ws * w_stage + E * e_stage > budget) --E;
  if (as * a_stage + ws * w_stage + E * e_stage > budget) return -1;
  const int as_cap = light ? 3 : 4, ws_cap = light ? 3 : 4;
  for (;;) {
    bool grew = false;
    if (ws < ws_cap && as * a_stage + (ws + 1) * w_stage + E * e_stage <= budget) { ++ws; grew = true; }
    if (as < as_cap && (as + 1) * a_stage + ws * w_stage + E * e_stage <= budget) { ++as; grew = true; }
    if (!grew) break;
  }
  if (light) {
    while (E < kMaxEpi && as * a_stage + ws * w_stage + (E + 1) * e_stage <= budget) ++E;
  }
  while (ws < kMaxW && as * a_stage + (ws + 1) * w_stage + E * e_stage <= budget) ++ws;
  while (as < kMaxA && (as + 1) * a_stage + ws * w_stage + E * e_stage <= budget) ++as;
  while (E < kMaxEpi && as * a_stage + ws * w_stage + (E + 1) * e_stage <= budget) ++E;
  static const int knock = getenv("NC_KNOCK") ? atoi(getenv("NC_KNOCK")) : 0;
  L.a_stages = as; L.w_stages = ws; L.a_rows_alloc = rows; L.tma_epilogue = 1; L.knock = knock; L.epi_stages = E;
  const size_t smem = 1024 + (size_t)as * a_stage + (((size_t)ws * w_stage + 1023) & ~(size_t)1023) + (size_t)E * e_stage;

  alignas(64) CUtensorMap tA, tW, tD32, tD16, tR;
  std::memset(&tD32, 0, sizeof tD32); std::memset(&tD16, 0, sizeof tD16); std::memset(&tR, 0, sizeof tR);
  const cuuint32_t one3[3] = {1, 1, 1};
  {
    const cuuint64_t gd[3] = {(cuuint64_t)p.a_pitch, (cuuint64_t)p.a_rows, (cuuint64_t)p.batch};
    const cuuint64_t gs[2] = {(cuuint64_t)p.a_pitch * 2, (cuuint64_t)p.a_clip_stride * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
    if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<float*>(p.A), gd, gs, box, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  {   // weights: [n_tiles * tiles_per_ntile * BN rows][64 halves], plain row-major; the TMA applies the 128B swizzle
    const cuuint64_t gd[2] = {64, (cuuint64_t)p.n_tiles * p.tiles_per_ntile * p.BN};
    const cuuint64_t gs[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)(p.BN / (pair ? 2 : 1))};
    const cuuint32_t one2[2] = {1, 1};
    if (enc(&tW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<float*>(p.W), gd, gs, box, one2, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  const cuuint64_t dd[3] = {(cuuint64_t)p.n_total, (cuuint64_t)p.m_rows, (cuuint64_t)p.batch};
  const cuuint32_t db[3] = {32, (cuuint32_t)kBM, 1};
  if (p.D || p.R) {
    const cuuint64_t ds[2] = {(cuuint64_t)p.n_total * 4, (cuuint64_t)p.d_clip_stride * 4};
    if (p.D && enc(&tD32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.D, dd, ds, db, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
    if (p.R && enc(&tR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.R), dd, ds, db, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  if (p.D16) {
    const cuuint64_t ds[2] = {(cuuint64_t)p.n_total * 2, (cuuint64_t)p.d_clip_stride * 2};
    if (enc(&tD16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, p.D16, dd, ds, db, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }

  const int m_units = pair ? (p.m_tiles_per_clip + 1) / 2 : p.m_tiles_per_clip;
  const int total = p.n_tiles * p.batch * m_units;
  if (total <= 0) return 0;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaError_t e;
  if (pair) {
    auto k = conv_h16_kernel<true>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem)) != cudaSuccess) return (int)e;
    int clusters = total < num_sms / 2 ? total : num_sms / 2;
    cfg.gridDim = dim3(2 * clusters);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    static int max_clusters = -1;   // co-resident 2-CTA clusters at this smem footprint (one CTA per SM)
    if (max_clusters < 0) {
      int n = 0;
      cudaLaunchConfig_t probe = cfg;
      probe.gridDim = dim3(num_sms / 2 * 2);
      if (cudaOccupancyMaxActiveClusters(&n, k, &probe) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 2; }
      max_clusters = n;
    }
    if (clusters > max_clusters) cfg.gridDim = dim3(2 * max_clusters);
    e = cudaLaunchKernelEx(&cfg, k, p, L, tA, tW, tD32, tD16, tR);
  } else {
    auto k = conv_h16_kernel<false>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem)) != cudaSuccess) return (int)e;
    cfg.gridDim = dim3(total < num_sms ? total : num_sms);
    e = cudaLaunchKernelEx(&cfg, k, p, L, tA, tW, tD32, tD16, tR);
  }
  return (int)e;
}

}  // namespace nc
