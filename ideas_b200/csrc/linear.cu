// A6: EqualLinear GEMMs (reference: F.linear -> cuBLAS sgemm, stylegan2/model.py:151-161) as a hand-written exact-fp32
// FFMA kernel.  The reference never enables TF32 for matmuls (torch default cuda.matmul.allow_tf32 = False), so its
// linears are true fp32; so are these.  Sizes are tiny next to the convolutions (< 0.05 GFLOP per image): the
// modulation linears of a whole Generator call (16 layers, one shared texture vector) run as ONE launch on the
// concatenated weights, the discriminator heads as one launch each.
//
//   C[i, j] = alpha * sum_r A(i, r) * B(j, r)          A(i,r) = a[i*sa_i + r*sa_r],  B(j,r) = b[j*sb_j + r*sb_r]
//
// Both operands are addressed through explicit element strides, so the forward (x W^T), the input gradient
// (dY W) and the weight gradient (dY^T X) are the same kernel on transposed views -- and the autograd Function
// above it (op/linear.py) is closed under differentiation (R1 needs second order through the discriminator heads).
// Long reductions with few output tiles (Dreal's 8192 -> 512 head at batch <= 128: 64 tiles) are split over
// gridDim.z and merged with fp32 atomics into the zero-initialised output.
#include "common.cuh"

#include <atomic>

namespace ideas {

static std::atomic<int> g_gemm_split{16};
void set_gemm_split(int v) { g_gemm_split.store(v); }

namespace {

constexpr int kT = 32;       // C tile is kT x kT, reduction step kT
constexpr int kLinThreads = 256;

template <bool A_R_FAST, bool B_R_FAST>
__global__ void __launch_bounds__(kLinThreads) gemm_nt_kernel(float* __restrict__ c, const float* __restrict__ a,
                                                              const float* __restrict__ b, int M, int N, int R,
                                                              int64_t sa_i, int64_t sa_r, int64_t sb_j, int64_t sb_r,
                                                              int64_t ldc, float alpha, int r_per_split, int atomic) {
  __shared__ float As[kT][kT + 1];   // [r][i]
  __shared__ float Bs[kT][kT + 1];   // [r][j]
  const int i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
  const int r_begin = blockIdx.z * r_per_split;
  const int r_end = min(R, r_begin + r_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // thread computes rows ty*2+{0,1}, cols tx*2+{0,1}
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int r0 = r_begin; r0 < r_end; r0 += kT) {
#pragma unroll
    for (int e = threadIdx.x; e < kT * kT; e += kLinThreads) {
      // the index that is contiguous in memory runs fastest across the threads of a warp
      const int ra = A_R_FAST ? (e & (kT - 1)) : (e / kT), ia = A_R_FAST ? (e / kT) : (e & (kT - 1));
      const int rb = B_R_FAST ? (e & (kT - 1)) : (e / kT), jb = B_R_FAST ? (e / kT) : (e & (kT - 1));
      As[ra][ia] = (i0 + ia < M && r0 + ra < r_end) ? __ldg(a + (int64_t)(i0 + ia) * sa_i + (int64_t)(r0 + ra) * sa_r) : 0.f;
      Bs[rb][jb] = (j0 + jb < N && r0 + rb < r_end) ? __ldg(b + (int64_t)(j0 + jb) * sb_j + (int64_t)(r0 + rb) * sb_r) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kT; ++r) {
      const float a0 = As[r][ty * 2], a1 = As[r][ty * 2 + 1];
      const float b0 = Bs[r][tx * 2], b1 = Bs[r][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int i = i0 + ty * 2 + u, j = j0 + tx * 2 + v;
      if (i < M && j < N) {
        float* p = c + (int64_t)i * ldc + j;
        if (atomic) atomicAdd(p, alpha * acc[u][v]);
        else *p = alpha * acc[u][v];
      }
    }
}

}  // namespace
}  // namespace ideas

// C (M x N, row stride ldc) = alpha * A B^T with strided operands (see the header comment).  When the kernel decides
// to split the reduction it accumulates atomically: C must then be zero-initialised -- the caller always passes a
// zeroed C (`c_is_zero` = 1) or forbids the split (0).
extern "C" int ideas_gemm_nt(float* c, const float* a, const float* b, int M, int N, int R, int64_t sa_i, int64_t sa_r,
                             int64_t sb_j, int64_t sb_r, int64_t ldc, float alpha, int c_is_zero, void* stream) {
  using namespace ideas;
  IDEAS_REQUIRE(M >= 0 && N >= 0 && R >= 0, "gemm_nt: negative extent");
  if (M == 0 || N == 0) return IDEAS_OK;
  IDEAS_REQUIRE(c && (R == 0 || (a && b)), "gemm_nt: null pointer");
  const int tiles = ceil_div(M, kT) * ceil_div(N, kT);
  // The maps on this path are skinny (M = batch): each CTA walks its share of the reduction as a serial chain of
  // 32-wide steps, so the launch is latency bound unless there are several CTAs per SM.  Split the reduction until
  // about `target` CTAs per SM exist (option "gemm_split": 16; 0 = the earlier rule, split only below one wave), every
  // split keeping at least 4 steps.  The 16-linear modulation GEMM of a Generator call (32 x 5 000 x 2 048: 156 tiles
  // of 64 steps) runs 16-way split: 149 -> 51 us (scripts/bench_gemm.py).
  int splits = 1;
  const int target = g_gemm_split.load();
  if (c_is_zero && R >= 8 * kT) {
    if (target <= 0) {
      if (tiles < kNumSMs) splits = (2 * kNumSMs) / tiles;
    } else {
      splits = ceil_div(target * kNumSMs, tiles);
    }
    const int max_splits = R / (4 * kT);
    splits = splits > max_splits ? max_splits : splits;
    splits = splits < 1 ? 1 : splits;
  }
  int r_per_split = ceil_div(ceil_div(R, splits), kT) * kT;
  if (r_per_split < kT) r_per_split = kT;
  splits = R > 0 ? ceil_div(R, r_per_split) : 1;
  dim3 grid(ceil_div(N, kT), ceil_div(M, kT), splits);
  IDEAS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_nt: M too large");
  const int atomic = splits > 1 ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool af = sa_r == 1, bf = sb_r == 1;
  if (af && bf) gemm_nt_kernel<true, true><<<grid, kLinThreads, 0, st>>>(c, a, b, M, N, R, sa_i, sa_r, sb_j, sb_r, ldc, alpha, r_per_split, atomic);
  else if (af) gemm_nt_kernel<true, false><<<grid, kLinThreads, 0, st>>>(c, a, b, M, N, R, sa_i, sa_r, sb_j, sb_r, ldc, alpha, r_per_split, atomic);
  else if (bf) gemm_nt_kernel<false, true><<<grid, kLinThreads, 0, st>>>(c, a, b, M, N, R, sa_i, sa_r, sb_j, sb_r, ldc, alpha, r_per_split, atomic);
  else gemm_nt_kernel<false, false><<<grid, kLinThreads, 0, st>>>(c, a, b, M, N, R, sa_i, sa_r, sb_j, sb_r, ldc, alpha, r_per_split, atomic);
  IDEAS_CHECK_LAUNCH("gemm_nt");
  return IDEAS_OK;
}
