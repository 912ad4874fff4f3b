// Thin inline-PTX wrappers for the sm_100a tensor-core path (tcgen05 / TMEM / mbarrier /
// bulk async copy).  Hand-written for this engine; bit layouts follow the PTX ISA
// "tcgen05 matrix descriptors" / "instruction descriptor" tables.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nc {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / bulk copy)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------- bulk copy (UBLKCP)
// 1-D bulk async copy global -> shared, completion on an mbarrier (complete_tx bytes).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ----------------------------------------------------------------------------- TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// commit all prior tcgen05.mma of this thread; arrive(1) on the mbarrier when they finish
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// registers -> TMEM, same shape as tmem_ld16 (thread i of the warp writes lane base_lane + i, 16 consecutive columns)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major, 128-byte swizzle, rows of 128 B, 8-row groups
// SBO bytes apart.  base_off = 3-bit "matrix base offset" field.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t sbo_bytes,
                                                    uint32_t base_off = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);                 // [0,14)  start address
  d |= static_cast<uint64_t>(1) << 16;                                // [16,30) LBO (unused, =1)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;        // [32,46) SBO
  d |= static_cast<uint64_t>(1) << 46;                                // [46,48) version = 1
  d |= static_cast<uint64_t>(base_off & 7) << 49;                     // [49,52) base offset
  d |= static_cast<uint64_t>(2) << 61;                                // [61,64) SWIZZLE_128B
  return d;
}

// Instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, dense.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (2u << 7)                    // a_format = TF32
         | (2u << 10)                   // b_format = TF32
         | (0u << 15) | (0u << 16)      // a_major, b_major = K
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::f16 with BF16 (fmt = 1) or FP16 (fmt = 0) operands, fp32 accumulate, K-major, dense.
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int fmt) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// constant part of a K-major SWIZZLE_128B descriptor with SBO = 1024 (8-row groups contiguous);
// OR in ((smem byte address & 0x3FFFF) >> 4) to get the descriptor of a tile / K step / row shift.
constexpr uint64_t kDescSw128Base = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                    (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
// K-major SWIZZLE_64B (64-byte rows, 8-row groups of 512 B contiguous): the hi-only weight tiles of the one- and
// two-pass fp16 modes.
constexpr uint64_t kDescSw64Base = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(512 >> 4) << 32) |
                                   (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(4) << 61);
__device__ __forceinline__ uint64_t desc_at(uint32_t saddr) {
  return kDescSw128Base | static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major SWIZZLE_128B tile whose
// base is 1024-byte aligned and whose 8-row groups are contiguous (SBO = 1024)
__host__ __device__ constexpr uint32_t sw128_offset(uint32_t r, uint32_t c) {
  return r * 128u + ((c ^ (r & 7u)) << 4);
}

// same for a SWIZZLE_64B tile (64-byte rows, chunk c in 0..3; the XOR uses address bits 7-8 = (r >> 1) & 3)
__host__ __device__ constexpr uint32_t sw64_offset(uint32_t r, uint32_t c) {
  return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4);
}

}  // namespace ptx
}  // namespace nc
