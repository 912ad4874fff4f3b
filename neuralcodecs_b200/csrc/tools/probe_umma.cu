// Hardware probe for the multi-tap implicit-GEMM conv design (run on a B200 via gpurun):
//   D[m, n] = sum_j sum_k A[m + j*dil, k] * W_j[n, k]      (M=128, K=32 per tap, tf32)
// Questions answered:
//   1. do hand-built SWIZZLE_128B K-major descriptors + idesc_tf32 give a correct GEMM?
//   2. can a conv tap be expressed as a ROW SHIFT of the A descriptor start address
//      (shift not a multiple of the 8-row swizzle period), and which value of the
//      descriptor "base offset" field does that need (0, or (addr>>7)&7)?
//   3. does a pre-swizzled weight tile land correctly through a 1-D bulk async copy?
// Inputs are multiples of 1/4 in [-1,1] so every product/sum is exact in tf32/fp32.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../umma.cuh"

using namespace nc::ptx;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A_g, const float* __restrict__ Wsw_g, float* __restrict__ D_g,
             int rowsA, int taps, int dil, int N, int bo_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A tile][W tiles]; both 1024-aligned
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = ((rowsA + 7) / 8) * 1024;
  uint8_t* sA = smem;
  uint8_t* sW = smem + a_bytes;
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    tmem_alloc<256>(&tmem_base_s);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  // A: plain [rowsA][32] -> swizzled smem (generic proxy writes)
  for (int i = tid; i < rowsA * 8; i += blockDim.x) {
    int r = i >> 3, c = i & 7;
    float4 v = reinterpret_cast<const float4*>(A_g)[r * 8 + c];
    *reinterpret_cast<float4*>(sA + sw128_offset(r, c)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t tile_bytes = N * 128;
    mbar_arrive_expect_tx(&bar_w, tile_bytes * taps);
    for (int j = 0; j < taps; ++j)
      bulk_g2s(sW + j * tile_bytes, Wsw_g + (size_t)j * N * 32, tile_bytes, &bar_w);
    mbar_wait(&bar_w, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, N);
    uint32_t acc = 0;
    for (int j = 0; j < taps; ++j) {
      const uint32_t a_row_addr = smem_u32(sA) + (uint32_t)(j * dil) * 128u;
      const uint32_t bo = bo_mode ? ((a_row_addr >> 7) & 7u) : 0u;
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = smem_desc_sw128(a_row_addr + k * 32, 1024, bo);
        uint64_t bd = smem_desc_sw128(smem_u32(sW) + j * tile_bytes + k * 32, 1024, 0);
        umma_tf32(tmem_base, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    tc_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < N / 16; ++c) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 16, v);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
    for (int i = 0; i < 16; ++i) D_g[(size_t)row * N + c * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

static float qrand() { return (float)((rand() % 9) - 4) / 4.0f; }

static int run_case(int taps, int dil, int N, int bo_mode) {
  const int rowsA = 128 + (taps - 1) * dil;
  std::vector<float> A((size_t)rowsA * 32), W((size_t)taps * N * 32), Wsw(W.size()), D((size_t)128 * N), Dref(D.size());
  for (auto& x : A) x = qrand();
  for (auto& x : W) x = qrand();
  // pre-swizzle W tiles: element (n, k) of tap j -> byte sw128_offset(n, k/4) + (k%4)*4
  for (int j = 0; j < taps; ++j)
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < 32; ++k) {
        uint32_t off = nc::ptx::sw128_offset(n, k / 4) + (k % 4) * 4;
        Wsw[(size_t)j * N * 32 + off / 4] = W[((size_t)j * N + n) * 32 + k];
      }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int j = 0; j < taps; ++j)
        for (int k = 0; k < 32; ++k) s += A[(size_t)(m + j * dil) * 32 + k] * W[((size_t)j * N + n) * 32 + k];
      Dref[(size_t)m * N + n] = s;
    }
  float *dA, *dW, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dW, Wsw.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, Wsw.data(), Wsw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, D.size() * 4));
  size_t smem = 1024 + ((rowsA + 7) / 8) * 1024 + (size_t)taps * N * 128;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dA, dW, dD, rowsA, taps, dil, N, bo_mode);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) {
    double e = fabs((double)D[i] - (double)Dref[i]);
    if (!(e <= 1e-5)) ++bad;
    if (e > maxerr || e != e) maxerr = e;
  }
  printf("case taps=%d dil=%d N=%d bo_mode=%d : %s (bad=%d/%zu maxerr=%g)\n", taps, dil, N, bo_mode,
         bad ? "FAIL" : "PASS", bad, D.size(), maxerr);
  cudaFree(dA); cudaFree(dW); cudaFree(dD);
  return bad;
}

int main() {
  srand(7);
  int fails = 0;
  fails += run_case(1, 0, 64, 0) != 0;    // plain GEMM K=32
  fails += run_case(1, 0, 256, 0) != 0;   // N=256
  fails += run_case(2, 8, 64, 0) != 0;    // shift by a whole swizzle period
  for (int bo = 0; bo < 2; ++bo) {
    run_case(2, 1, 64, bo);
    run_case(7, 1, 64, bo);
    run_case(7, 3, 64, bo);
    run_case(7, 9, 64, bo);
    run_case(3, 5, 128, bo);
  }
  printf("base cases failed: %d\n", fails);
  return 0;
}
