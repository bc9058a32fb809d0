// Probe (B200 only): does tcgen05.mma accept a K-major SWIZZLE_128B operand whose start address is
// offset by r rows (r * 128 B) from the 1024-byte swizzle atom, and does the descriptor's
// base_offset field (bits [49,52)) have to carry r % 8?  The halo-reuse convolution kernel relies
// on the answer: one halo'd activation tile in shared memory serves all 9 taps of a 3x3 filter by
// shifting the descriptor start address.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe scripts/probe_umma_offset.cu && /tmp/probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../ideas_b200/csrc/umma_ptx.cuh"

using namespace ideas;

constexpr int kRows = 192;   // rows of A held in smem
constexpr int kN = 32;

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* A, const float* B, float* D, int shift, int use_bo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  float* sa = reinterpret_cast<float*>(gen);
  float* sb = reinterpret_cast<float*>(gen + kRows * 128);
  // manual 128B swizzle: element (row i, col c) -> i*128 + ((c/4) ^ (i%8))*16 + (c%4)*4
  for (int e = threadIdx.x; e < kRows * 32; e += blockDim.x) {
    const int i = e / 32, c = e % 32;
    sa[i * 32 + (((c >> 2) ^ (i & 7)) << 2) + (c & 3)] = A[e];
  }
  for (int e = threadIdx.x; e < kN * 32; e += blockDim.x) {
    const int i = e / 32, c = e % 32;
    sb[i * 32 + (((c >> 2) ^ (i & 7)) << 2) + (c & 3)] = B[e];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 32);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = ptx::idesc_tf32(128, kN, 0, 0);
    uint64_t adesc = ptx::smem_desc_sw128(base + shift * 128, 16, 1024);
    if (use_bo) adesc |= (uint64_t)(shift & 7) << 49;
    const uint64_t bdesc = ptx::smem_desc_sw128(base + kRows * 128, 16, 1024);
    for (int k = 0; k < 4; ++k) ptx::mma_tf32(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    ptx::mma_commit(ptx::smem_u32(&bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  float v[32];
  ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int j = 0; j < kN; ++j) D[row * kN + j] = v[j];
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 32);
  }
}

int main() {
  std::vector<float> A(kRows * 32), B(kN * 32), D(128 * kN);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 17 - 8);
  for (auto& v : B) v = (float)(rand() % 9 - 4);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = kRows * 128 + kN * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 33, 34, 35, 63};
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int s : shifts) {
      cudaMemset(dD, 0, D.size() * 4);
      probe_kernel<<<1, 128, smem>>>(dA, dB, dD, s, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("shift %2d base_offset=%d : CUDA error %s\n", s, use_bo, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double worst = 0;
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < 32; ++k) ref += (double)A[(m + s) * 32 + k] * B[n * 32 + k];
          const double err = fabs(ref - D[m * kN + n]);
          if (err > worst) worst = err;
          bad += err > 1e-3;
        }
      printf("shift %2d base_offset=%d : max abs err %.3f, mismatches %d / %d -> %s\n", s, use_bo, worst, bad, 128 * kN,
             bad ? "WRONG" : "ok");
    }
  return 0;
}
