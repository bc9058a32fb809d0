// Probe (B200 only): MN-major fp32 operands (TMA SWIZZLE_128B_ATOM_32B, descriptor layout SWIZZLE_128B_BASE32B)
// whose start address is shifted by s rows (s * 128 B, one row = one K index = one pixel).  The multi-tap
// weight-gradient kernel relies on it: one halo'd x tile serves several filter taps.
//   D[k][c] = sum_{p < 32} dy[p][k] * x[p + s][c],   k < 128, c < 32
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -o build/probe_mn scripts/probe_umma_mn_offset.cu && build/probe_mn
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../ideas_b200/csrc/umma_ptx.cuh"

using namespace ideas;

constexpr int kXRows = 96;   // pixels of x held in smem
constexpr int kPix = 32;     // reduction length

struct alignas(64) Params {
  CUtensorMap dy;   // (32 ch, P px, 4 groups) box (32, 32, 4)
  CUtensorMap x;    // (32 ch, P px) box (32, 96)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p, float* D, int shift) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, full;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sdy = base, sx = base + 4 * kPix * 128;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::mbar_init(ptx::smem_u32(&full), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 32);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full), 4 * kPix * 128 + kXRows * 128);
    ptx::tma_load_3d(sdy, &p.dy, ptx::smem_u32(&full), 0, 0, 0);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sx), "l"(&p.x), "r"(ptx::smem_u32(&full)), "r"(0), "r"(0) : "memory");
    ptx::mbar_wait(ptx::smem_u32(&full), 0);
    ptx::tc_fence_after();
    constexpr uint32_t idesc = ptx::idesc_tf32(128, 32, 1, 1);
    for (int k = 0; k < kPix / 8; ++k) {
      const uint64_t adesc = ptx::smem_desc_sw128_base32(sdy + k * 1024, kPix * 128, 512);
      const uint64_t bdesc = ptx::smem_desc_sw128_base32(sx + k * 1024 + shift * 128, kXRows * 128, 512);
      ptx::mma_tf32(tmem, adesc, bdesc, idesc, k != 0);
    }
    ptx::mma_commit(ptx::smem_u32(&bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  float v[32];
  ptx::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int j = 0; j < 32; ++j) D[row * 32 + j] = v[j];
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 32);
  }
}

int main() {
  const int P = kXRows;
  std::vector<float> dy(P * 128), x(P * 32), D(128 * 32);
  srand(2);
  for (auto& v : dy) v = (float)(rand() % 9 - 4);
  for (auto& v : x) v = (float)(rand() % 17 - 8);
  float *d_dy, *d_x, *d_D;
  cudaMalloc(&d_dy, dy.size() * 4);
  cudaMalloc(&d_x, x.size() * 4);
  cudaMalloc(&d_D, D.size() * 4);
  cudaMemcpy(d_dy, dy.data(), dy.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_x, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  Params p;
  {
    cuuint64_t gd[3] = {32, (cuuint64_t)P, 4}, gs[2] = {128 * 4, 128};
    cuuint32_t bx[3] = {32, kPix, 4}, es[3] = {1, 1, 1};
    CUresult r = enc(&p.dy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_dy, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode dy failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t gd[2] = {32, (cuuint64_t)P}, gs[1] = {32 * 4};
    cuuint32_t bx[2] = {32, kXRows}, es[2] = {1, 1};
    CUresult r = enc(&p.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_x, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode x failed %d\n", (int)r); return 1; }
  }
  const int smem = 4 * kPix * 128 + kXRows * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 19, 33, 42, 63};
  for (int s : shifts) {
    cudaMemset(d_D, 0, D.size() * 4);
    probe_kernel<<<1, 128, smem>>>(p, d_D, s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %2d: CUDA error %s\n", s, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), d_D, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    double worst = 0;
    for (int k = 0; k < 128; ++k)
      for (int c = 0; c < 32; ++c) {
        double ref = 0;
        for (int pp = 0; pp < kPix; ++pp) ref += (double)dy[pp * 128 + k] * x[(pp + s) * 32 + c];
        const double err = fabs(ref - D[k * 32 + c]);
        worst = err > worst ? err : worst;
        bad += err > 1e-3;
      }
    printf("MN-major shift %2d : max abs err %.3f, mismatches %d / %d -> %s\n", s, worst, bad, 128 * 32, bad ? "WRONG" : "ok");
  }
  return 0;
}
