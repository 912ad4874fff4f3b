// Hardware probe (run on a B200 via gpurun): how many clocks does one tcgen05.mma (kind::f16, bf16, M = 128, K = 16, operands in
// shared memory) take when the MMAs are issued back to back by one thread
//   (a) all into ONE accumulator (a dependent chain, what a conv tile's k loop is), or
//   (b) alternating between TWO accumulators (independent chains),
// for N = 32 .. 256?  Answers whether the ~100 clk per MMA seen in the role timelines of the narrow fused units
// (profiles/r02_ru_fused_timeline.txt) is shared-memory operand bandwidth or a minimum issue interval of dependent MMAs.
// Operand contents are irrelevant (zeros).
#include <cstdio>
#include <cstdlib>
#include "../umma.cuh"

using namespace nc::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int n_mma, int accs, int same_operands, int issuers, int unrolled, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 4 A tiles of [128 rows][128 B]
  uint8_t* sB = smem + 4 * 16384;     // 4 B tiles of [256 rows][128 B]
  __shared__ uint64_t bar, bar2;
  __shared__ long long t_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4 * 16384 + 4 * 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) { tmem_alloc<512>(&tmem_base_s); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // issuer threads: lane 0 of warp 0 (and of warp 1 when issuers == 2); issuer w owns accumulators starting at column 256*w... (accs per issuer)
  if ((tid & 31) == 0 && warp < issuers) {
    uint64_t* my_bar = warp == 0 ? &bar : &bar2;
    const uint32_t idesc = idesc_f16(128, N, 1);
    const uint64_t a0 = desc_at(smem_u32(sA)), b0 = desc_at(smem_u32(sB));
    const uint32_t acc_base = tmem_base + (uint32_t)warp * 256u;
    const uint32_t acc_step = issuers == 2 ? 128u : 256u;   // two accumulators per issuer fit 128 columns each when N <= 128
    for (int i = 0; i < 16; ++i) umma_f16(acc_base, a0, b0, idesc, i ? 1u : 0u);
    tc_commit(my_bar);
    mbar_wait(my_bar, 0);
    tc_fence_after();
    const int n = n_mma / issuers;
    const long long t0 = clock64();
    if (unrolled) {   // the production pattern: six MMAs per "stage" with descriptor offsets known at compile time
      for (int i = 0; i < n; i += 6) {
        const uint32_t d = acc_base + (uint32_t)((i / 6) % accs) * acc_step;
        const uint64_t a = same_operands ? a0 : a0 + (uint64_t)(((i / 6) & 3) * (16384 >> 4));
        const uint64_t b = same_operands ? b0 : b0 + (uint64_t)(((i / 6) & 3) * (32768 >> 4));
        const uint32_t first = i >= 6 * accs ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_f16(d, a + 4 + 2 * k, b + 2 * k, idesc, first | (uint32_t)k);
          umma_f16(d, a + 2 * k, b + 4 + 2 * k, idesc, 1);
          umma_f16(d, a + 2 * k, b + 2 * k, idesc, 1);
        }
      }
    } else {
      for (int i = 0; i < n; ++i) {
        const uint32_t d = acc_base + (uint32_t)(i % accs) * acc_step;
        const uint64_t a = same_operands ? a0 : a0 + (uint64_t)(((i >> 2) & 3) * (16384 >> 4) + 2 * (i & 3));
        const uint64_t b = same_operands ? b0 : b0 + (uint64_t)(((i >> 2) & 3) * (32768 >> 4) + 2 * (i & 3));
        umma_f16(d, a, b, idesc, i >= accs ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    tc_commit(my_bar);
    mbar_wait(my_bar, 1);
    const long long t2 = clock64();
    t_done[warp] = t2 - t0;
    if (warp == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (tid == 0) out[1] = issuers == 2 ? (t_done[0] > t_done[1] ? t_done[0] : t_done[1]) : t_done[0];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  const size_t smem = 1024 + 4 * 16384 + 4 * 32768;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_mma = 2048;
  printf("# tcgen05.mma kind::f16 (bf16), M=128, K=16, operands in smem, %d MMAs issued back to back by one thread; issue clk = per MMA of one issuer; total clk = (last completion - first issue) / all MMAs\n", n_mma);
  printf("%5s %6s %9s %8s %9s %12s %12s %10s\n", "N", "accs", "operands", "issuers", "unrolled", "issue clk", "total clk", "floor N/2");
  for (int N : {64, 128, 256}) {
    for (int unrolled : {0, 1}) {
      for (int issuers : {1, 2}) {
        if (issuers == 2 && N > 128) continue;   // 4 accumulators of N columns must fit 512
        for (int accs : {1, 2}) {
          const int n = 6 * 2 * 170;   // divisible by 6 and by the issuer count
          rate_kernel<<<1, 128, smem>>>(N, n, accs, 0, issuers, unrolled, d_out);
          CK(cudaDeviceSynchronize());
          long long h[2];
          CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
          printf("%5d %6d %9s %8d %9d %12.1f %12.1f %10d\n", N, accs, "rotating", issuers, unrolled, (double)h[0] / (n / issuers), (double)h[1] / n, N / 2);
        }
      }
    }
  }
  return 0;
}
