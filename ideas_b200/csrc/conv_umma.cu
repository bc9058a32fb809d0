// A1/A5: implicit-GEMM convolution on the Blackwell tensor path (tcgen05.mma kind::tf32,
// TMEM accumulators, TMA-fed 128-byte-swizzled shared memory).  The reference has no native code
// here: it calls cuDNN through F.conv2d / F.conv_transpose2d (stylegan2/model.py:115,258,267,273;
// models.py:32).
//
// Forward-form kernel (forward conv, data gradient, transposed conv -- every ConvGeom):
//   GEMM  D[pixel, k] = sum_{tap, c} X[pixel + tap, c] * Wp[tap][k][c]
//   * no im2col buffer: for each (tap, 32-channel slab) ONE 4-D TMA box (32 ch x BW x BH x BN pixels,
//     BW*BH*BN = 128) lands the shifted pixel patch in shared memory as 128 rows x 128 B -- exactly the
//     K-major SWIZZLE_128B operand layout tcgen05 wants; zero padding is the TMA out-of-bounds fill.
//     Stride-2 sources are addressed through per-parity tensor maps (no elementStrides).
//   * CTA tile = 256 pixels (two 128-row sub-tiles, two accumulators) x BLOCK_N output channels, so
//     each weight slab fetched from L2 feeds two MMAs (the kernel is L2->SM-feed bound: fp32 operands).
//   * warp 0 = TMA producer, warp 1 = MMA issuer (single thread) + TMEM owner, warps 2-5 = epilogue:
//     tcgen05.ld -> demodulation d[n,k] -> +bias -> leaky ReLU*gain -> 128-bit NHWC stores.
// Weight-gradient kernel: see conv_umma_wgrad below.
#include "common.cuh"
#include "conv_geom.h"
#include "umma_ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>

namespace ideas {

void set_blur_variant(int v);   // upfirdn2d.cu
void set_resample_variant(int v);
void set_gemm_split(int v);      // linear.cu

namespace {

// option "tma_tf32" (default 1): tensor maps declare CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, so the TMA unit
// rounds fp32 to tf32 (to nearest) while filling shared memory.  With plain FLOAT32 maps tcgen05 simply
// drops the low 13 mantissa bits, a truncation that biases every product low: measured on B200 at
// K = 4608, max rel error vs fp64 7.5e-4 (signed bias -7.1e-4) truncated vs 3.0e-4 (bias -1e-5) rounded.
std::atomic<int> g_tma_tf32{1};
std::atomic<int> g_tail_split{1};     // option "tail_split": balance the last round of the pixel-major kernel (launch_fwd)

constexpr int kThreads = 192;
constexpr int kBlockM = 128;   // pixels per sub-tile == UMMA M
constexpr int kSub = 2;        // sub-tiles per CTA
constexpr int kBlockK = 32;    // fp32 channels per k-block (= 128 B swizzle span)
constexpr int kUmmaK = 8;      // tf32 K per instruction (32 B)
constexpr uint32_t kABytes = kBlockM * 128;

struct TapU {
  int16_t oy, ox;   // offset added to the box origin (in units of the addressed map's pixels)
  int16_t widx;     // weight tap index
  int16_t map;      // which source tensor map (parity)
};

struct alignas(64) FwdParams {
  CUtensorMap src[4];
  CUtensorMap w;
  float* dst;
  const float* out_scale;
  const float* bias;
  const float* residual;
  float res_scale;
  int N, QH, QW;
  int OH, OW, OC;
  int o_s, o_py, o_px;
  int bw, bh, bn;
  int tiles_x, tiles_y, subtiles;
  int full_pairs;   // CTAs with blockIdx.y < full_pairs own two sub-tiles, the rest one (tail balancing, launch_fwd)
  int ntaps, csteps;
  int act;
  float alpha, gain;
  TapU taps[kMaxTaps];
};

template <int BN> struct FwdCfg {
  static constexpr uint32_t kBBytes = BN * 128;
  static constexpr uint32_t kStageBytes = kSub * kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 6 ? 6 : (200 * 1024 / kStageBytes);
  static constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024;
};

// Epilogue of one 128-pixel sub-tile (accumulator columns [j*BN, (j+1)*BN)): one thread = one pixel (TMEM lane),
// tcgen05.ld of 32 channels at a time -> demodulation -> +bias -> leaky ReLU*gain -> (+residual)*res_scale -> 128-bit
// NHWC stores.  Shared by the single-CTA and the CTA-pair kernel.
template <int BN>
__device__ __forceinline__ void fwd_epilogue_subtile(const FwdParams& p, uint32_t tmem_base, int quarter, int lane, int j,
                                                     int qx0, int qy0, int n0, bool sub_ok, int k0) {
  const int row = quarter * 32 + lane;
  const int wq = row % p.bw;
  const int t = row / p.bw;
  const int qx = qx0 + wq, qy = qy0 + (t % p.bh), n = n0 + t / p.bh;
  const bool valid = sub_ok && n < p.N && qy < p.QH && qx < p.QW;
  float* dp = nullptr;
  const float* os = nullptr;
  const float* rp = nullptr;
  if (valid) {
    const int64_t off = (((int64_t)n * p.OH + (qy * p.o_s + p.o_py)) * p.OW + (qx * p.o_s + p.o_px)) * p.OC + k0;
    dp = p.dst + off;
    if (p.residual) rp = p.residual + off;
    if (p.out_scale) os = p.out_scale + (int64_t)n * p.OC + k0;
  }
  for (int ch = 0; ch < BN / 32; ++ch) {
    float v[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * BN + ch * 32), v);
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        if (os) {
          const float4 s = __ldg(reinterpret_cast<const float4*>(os + ch * 32 + i));
          r.x *= s.x; r.y *= s.y; r.z *= s.z; r.w *= s.w;
        }
        if (p.bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + k0 + ch * 32 + i));
          r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
        }
        if (p.act == IDEAS_ACT_LRELU) {
          r.x = lrelu(r.x, p.alpha) * p.gain; r.y = lrelu(r.y, p.alpha) * p.gain;
          r.z = lrelu(r.z, p.alpha) * p.gain; r.w = lrelu(r.w, p.alpha) * p.gain;
        }
        if (rp) {
          const float4 q = ld_stream4(rp + ch * 32 + i);
          const float rs = p.res_scale;
          r.x = (r.x + q.x) * rs; r.y = (r.y + q.y) * rs; r.z = (r.z + q.z) * rs; r.w = (r.w + q.w) * rs;
        }
        *reinterpret_cast<float4*>(dp + ch * 32 + i) = r;
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_fwd_kernel(const __grid_constant__ FwdParams p) {
  using Cfg = FwdCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::kStages];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::kStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int k0 = blockIdx.x * BN;
  const int nkb = p.ntaps * p.csteps;

  // sub-tile origins
  int qx0[kSub], qy0[kSub], n0[kSub];
  bool sub_ok[kSub];
  const bool pair = (int)blockIdx.y < p.full_pairs;
  const int nsub = pair ? kSub : 1;
  const int first = pair ? (int)blockIdx.y * kSub : p.full_pairs * kSub + ((int)blockIdx.y - p.full_pairs);
#pragma unroll
  for (int j = 0; j < kSub; ++j) {
    const int id = first + j;
    sub_ok[j] = j < nsub && id < p.subtiles;
    const int bx = id % p.tiles_x;
    const int t = id / p.tiles_x;
    qx0[j] = bx * p.bw;
    qy0[j] = (t % p.tiles_y) * p.bh;
    n0[j] = (t / p.tiles_y) * p.bn;
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&accum_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), Cfg::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      ptx::tma_prefetch_desc(&p.w);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int t = kb / p.csteps;
        const int c0 = (kb - t * p.csteps) * kBlockK;
        const TapU tp = p.taps[t];
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
        ptx::mbar_arrive_expect_tx(fb, (uint32_t)nsub * kABytes + Cfg::kBBytes);
        const uint32_t sa = tiles + stage * Cfg::kStageBytes;
#pragma unroll
        for (int j = 0; j < kSub; ++j)
          if (j < nsub) ptx::tma_load_4d(sa + j * kABytes, &p.src[tp.map], fb, c0, qx0[j] + tp.ox, qy0[j] + tp.oy, n0[j]);
        ptx::tma_load_3d(sa + kSub * kABytes, &p.w, fb, c0, k0, tp.widx);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = ptx::idesc_tf32(kBlockM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
        ptx::tc_fence_after();
        const uint32_t sa = tiles + stage * Cfg::kStageBytes;
        const uint64_t bdesc = ptx::smem_desc_sw128(sa + kSub * kABytes, 16, 1024);
#pragma unroll
        for (int j = 0; j < kSub; ++j) {
          if (j >= nsub) break;
          const uint64_t adesc = ptx::smem_desc_sw128(sa + j * kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::mma_tf32(tmem_base + j * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((kb | k) != 0));
        }
        ptx::mma_commit(ptx::smem_u32(&empty_bar[stage]));   // frees the stage when these MMAs retire
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
      }
      ptx::mma_commit(ptx::smem_u32(&accum_bar));
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    ptx::mbar_wait(ptx::smem_u32(&accum_bar), 0);
    ptx::tc_fence_after();
#pragma unroll
    for (int j = 0; j < kSub; ++j) {
      if (j >= nsub) break;
      fwd_epilogue_subtile<BN>(p, tmem_base, quarter, lane, j, qx0[j], qy0[j], n0[j], sub_ok[j], k0);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------
// CTA-pair, persistent variant of the forward-form kernel for 256-channel output tiles (option "pair").
// The kernel above leaves the tensor pipe idle for about a quarter of each CTA's life (ncu: pipe active 70 % forward,
// 80 % data gradient at cfg 3): its two 256-column accumulators fill TMEM, so the epilogue -- 256 KB through four
// warps -- cannot overlap the next tile, and with one CTA per SM nothing else runs meanwhile.  Here two CTAs on the
// SMs of one TPC execute ONE tcgen05.mma.cta_group::2 of M = 256 pixels (128 per CTA, each CTA's own accumulator
// lanes) x N = 256 output channels, each CTA staging its 128 pixels and HALF of the weight tile: the same bytes per
// flop as above with half the accumulator columns per CTA, which leaves room for TWO accumulator buffers.  The
// cluster is persistent over a static list of (channel tile, pixel-pair tile) items, so the epilogue of item i runs
// under the MMAs of item i+1, the TMA producers run ahead across items, and the work granule is 128 pixels per SM
// instead of 256 (a finer tail).
//   warp 0  TMA producer (both CTAs): own pixels + own half of the weight rows; all bytes are counted on the LEADER's
//           full barrier (cp.async.bulk.tensor ... cta_group::2, barrier address via mapa)
//   warp 1  MMA issuer (leader only) + TMEM owner (tcgen05.alloc.cta_group::2 in both); tcgen05.commit multicasts the
//           stage-free and accumulator-ready arrivals to both CTAs
//   warps 2-5 epilogue of this CTA's 128 lanes; "buffer drained" arrives on the leader's barrier from both CTAs
// ------------------------------------------------------------------------------------------
template <int BN> struct PairCfg {
  static constexpr uint32_t kBBytes = (BN / 2) * 128;             // this CTA's half of the BN-channel weight tile
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;      // 32 KB (BN = 256) / 24 KB (BN = 128)
  static constexpr uint32_t kEpiBytes = 4 * 2 * 4096;             // two 4 KB TMA-store staging tiles per epilogue warp
  static constexpr int kMaxStages = (226 * 1024 - kEpiBytes) / kStageBytes;   // 6 / 8: all of the 227 KB a CTA can have
  static constexpr int kStages = BN == 256 ? 6 : 8;               // default ring depth (option "pair_stages")
  static constexpr uint32_t kTmemCols = 2 * BN;                   // two accumulator buffers
};

struct alignas(64) PairParams {
  FwdParams f;
  CUtensorMap dst;      // output as {OC, QW, QH, N} with the phase offset / stride folded in (TMA-store epilogue)
  int items, ktiles, stages;
  int epi;              // 1 = epilogue through shared memory + TMA store, 0 = direct 128-bit stores
  int wy, wn;           // a warp's 32 accumulator rows as a box: bw x wy x wn pixels
};

// Epilogue of one 128-pixel sub-tile through shared memory and TMA stores (pair kernel).  A warp owns 32 accumulator
// rows = a (bw x wy x wn)-pixel box.  Per 32-channel chunk: tcgen05.ld -> demodulation, bias, leaky ReLU, residual in
// registers -> the thread's 128-byte row into a 4 KB staging tile in the 128-byte-swizzle pattern (conflict-free
// 16-byte stores) -> fence.proxy.async -> one cp.async.bulk.tensor store of the box.  Pixels outside the image are
// clipped by the TMA unit, so there are no store predicates and no per-pixel address arithmetic; two staging tiles per
// warp let the store of chunk i drain under the arithmetic of chunk i+1.
template <int BN>
__device__ __forceinline__ void pair_epilogue_tma(const PairParams& pp, uint32_t tmem_base, uint32_t stage_smem, int quarter,
                                                  int lane, int buf, int qx0, int qy0, int n0, bool sub_ok, int k0,
                                                  uint32_t& nstores) {
  const FwdParams& p = pp.f;
  const int row = quarter * 32 + lane;
  const int wq = row % p.bw;
  const int t = row / p.bw;
  const int qx = qx0 + wq, qy = qy0 + (t % p.bh), n = n0 + t / p.bh;
  const bool valid = sub_ok && n < p.N && qy < p.QH && qx < p.QW;      // guards the per-pixel LOADS only
  const float* os = nullptr;
  const float* rp = nullptr;
  if (valid) {
    if (p.residual)
      rp = p.residual + (((int64_t)n * p.OH + (qy * p.o_s + p.o_py)) * p.OW + (qx * p.o_s + p.o_px)) * p.OC + k0;
    if (p.out_scale) os = p.out_scale + (int64_t)n * p.OC + k0;
  }
  // box origin of this warp's 32 rows
  const int t0 = (quarter * 32) / p.bw;
  const int by = qy0 + (t0 % p.bh), bn = n0 + t0 / p.bh;
  const uint32_t rowoff = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
  for (int ch = 0; ch < BN / 32; ++ch) {
    float v[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + ch * 32), v);
    const uint32_t tile = stage_smem + (nstores & 1u) * 4096u;
    if (lane == 0) ptx::tma_store_wait_read<1>();        // the store that last read this tile has drained it
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 r = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      if (os) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(os + ch * 32 + i));
        r.x *= s.x; r.y *= s.y; r.z *= s.z; r.w *= s.w;
      }
      if (p.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + k0 + ch * 32 + i));
        r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
      }
      if (p.act == IDEAS_ACT_LRELU) {
        r.x = lrelu(r.x, p.alpha) * p.gain; r.y = lrelu(r.y, p.alpha) * p.gain;
        r.z = lrelu(r.z, p.alpha) * p.gain; r.w = lrelu(r.w, p.alpha) * p.gain;
      }
      if (rp) {
        const float4 q = ld_stream4(rp + ch * 32 + i);
        const float rs = p.res_scale;
        r.x = (r.x + q.x) * rs; r.y = (r.y + q.y) * rs; r.z = (r.z + q.z) * rs; r.w = (r.w + q.w) * rs;
      }
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                   ::"r"(tile + rowoff + ((((uint32_t)i >> 2) ^ sw) << 4)), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w)
                   : "memory");
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_4d(&pp.dst, tile, k0 + ch * 32, qx0, by, bn);
      ptx::tma_store_commit();
    }
    ++nstores;
  }
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_umma_fwd_pair_kernel(const __grid_constant__ PairParams pp) {
  using Cfg = PairCfg<BN>;
  constexpr int kPairMaxStages = Cfg::kMaxStages;
  const int kPairStages = pp.stages;
  constexpr uint32_t kPairStageBytes = Cfg::kStageBytes;
  const FwdParams& p = pp.f;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPairMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kPairMaxStages];
  __shared__ __align__(8) uint64_t tfull[2];
  __shared__ __align__(8) uint64_t tempty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1, nclusters = (int)gridDim.x >> 1;
  const int nkb = p.ntaps * p.csteps;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty[s]), 8);      // four epilogue warps in each of the two CTAs
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair(ptx::smem_u32(&tmem_slot), Cfg::kTmemCols);
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync();               // barriers of both CTAs initialised before any remote arrival
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // item -> (channel tile, this CTA's 128-pixel sub-tile)
#define IDEAS_PAIR_DECODE(item)                         \
  const int k0 = ((item) % pp.ktiles) * BN;             \
  const int sid = ((item) / pp.ktiles) * 2 + (int)rank; \
  const int bx = sid % p.tiles_x;                       \
  const int tt = sid / p.tiles_x;                       \
  const int qx0 = bx * p.bw;                            \
  const int qy0 = (tt % p.tiles_y) * p.bh;              \
  const int n0 = (tt / p.tiles_y) * p.bn;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      ptx::tma_prefetch_desc(&p.w);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < pp.items; item += nclusters) {
        IDEAS_PAIR_DECODE(item)
        for (int kb = 0; kb < nkb; ++kb) {
          const int t = kb / p.csteps;
          const int c0 = (kb - t * p.csteps) * kBlockK;
          const TapU tp = p.taps[t];
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
          if (rank == 0) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full_bar[stage]), 2u * kPairStageBytes);
          const uint32_t fb = ptx::mapa_shared(ptx::smem_u32(&full_bar[stage]), 0);
          const uint32_t sa = tiles + stage * kPairStageBytes;
          // a sub-tile past the end (odd count) has n0 >= N: the box is all out-of-bounds zero fill
          ptx::tma_load_4d_pair(sa, &p.src[tp.map], fb, c0, qx0 + tp.ox, qy0 + tp.oy, n0);
          ptx::tma_load_3d_pair(sa + kABytes, &p.w, fb, c0, k0 + (int)rank * (BN / 2), tp.widx);
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader only) =====
      constexpr uint32_t idesc = ptx::idesc_tf32(256, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = cluster_id; item < pp.items; item += nclusters, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty[buf]), ((uint32_t)(it >> 1) & 1u) ^ 1u);   // both epilogues drained it
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t sa = tiles + stage * kPairStageBytes;
          const uint64_t adesc = ptx::smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = ptx::smem_desc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            ptx::mma_tf32_pair(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          ptx::mma_commit_pair(ptx::smem_u32(&empty_bar[stage]), 3);   // frees the stage in both CTAs
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
        ptx::mma_commit_pair(ptx::smem_u32(&tfull[buf]), 3);           // accumulators ready in both CTAs
      }
    }
  } else {
    // ===== epilogue: warps 2..5 of each CTA drain that CTA's 128 accumulator lanes =====
    const int quarter = warp & 3;
    const uint32_t epi_smem = tiles + (uint32_t)kPairStages * kPairStageBytes + (uint32_t)(warp - 2) * 8192u;
    uint32_t nstores = 0;
    int it = 0;
    for (int item = cluster_id; item < pp.items; item += nclusters, ++it) {
      IDEAS_PAIR_DECODE(item)
      const int buf = it & 1;
      ptx::mbar_wait(ptx::smem_u32(&tfull[buf]), (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      if (pp.epi)
        pair_epilogue_tma<BN>(pp, tmem_base, epi_smem, quarter, lane, buf, qx0, qy0, n0, sid < p.subtiles, k0, nstores);
      else
        fwd_epilogue_subtile<BN>(p, tmem_base, quarter, lane, buf, qx0, qy0, n0, sid < p.subtiles, k0);
      // every tcgen05.ld of this warp has completed (wait::ld): hand the buffer back to the leader's MMA thread
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&tempty[buf]), 0));
    }
    if (lane == 0) ptx::tma_store_wait<0>();             // every bulk store of this warp has completed
  }
#undef IDEAS_PAIR_DECODE
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync();               // the leader's MMAs read the peer's shared memory: leave together
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

// fp32 tensor viewed as rank-`rank` (innermost first), 128-byte swizzle, zero OOB fill
int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, bool plain_f32 = false) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return IDEAS_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  // loads round fp32 to tf32 in the TMA unit (option tma_tf32); stores must keep the fp32 bits
  const CUtensorMapDataType dt =
      (g_tma_tf32.load() && !plain_f32) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return IDEAS_ERR_CUDA;
  }
  return IDEAS_OK;
}

int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// choose the 128-pixel box (bw x bh x bn) that covers the q grid with the fewest sub-tiles
void choose_box(int N, int QH, int QW, int* bw, int* bh, int* bn, int* tx, int* ty, int* subtiles) {
  long best = -1;
  for (int w = 16; w >= 4; w >>= 1) {
    if (w > pow2_ceil(QW) && w > 4) continue;
    for (int h = 128 / w; h >= 1; h >>= 1) {
      if (h > pow2_ceil(QH) && h > 1) continue;
      const int n = 128 / (w * h);
      const long cnt = (long)ceil_div(QW, w) * ceil_div(QH, h) * ceil_div(N, n);
      if (best < 0 || cnt < best) {
        best = cnt; *bw = w; *bh = h; *bn = n; *tx = ceil_div(QW, w); *ty = ceil_div(QH, h);
      }
    }
  }
  *subtiles = (int)best;
}

template <int BN>
int launch_fwd(const FwdParams& p, int ntiles_n, cudaStream_t st) {
  using Cfg = FwdCfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_umma_fwd_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "conv_umma: cudaFuncSetAttribute");
  // Tail balancing: one CTA per SM (the tiles fill shared memory), so a grid of T CTAs runs in ceil(T / 148)
  // rounds.  When the last round would be less than half full, its CTAs are given ONE 128-pixel sub-tile instead
  // of two -- twice as many CTAs of half the length, i.e. the tail costs ~half a round (cfg 3 at batch 16: 512
  // CTAs = 3.46 rounds -> 444 pairs + 136 singles = 3.5 instead of 4).
  const int pf = p.subtiles / kSub;                       // complete pairs; an odd last sub-tile is a single anyway
  int full_pairs = pf;
  const int total = (pf + (p.subtiles - pf * kSub)) * ntiles_n;
  const int tail = total % kNumSMs;
  if (total > kNumSMs && tail > 0 && tail <= kNumSMs / 2 && g_tail_split.load()) {
    int cp = ceil_div(tail, ntiles_n);
    cp = cp > pf ? pf : cp;
    full_pairs = pf - cp;
  }
  FwdParams q = p;
  q.full_pairs = full_pairs;
  dim3 grid(ntiles_n, full_pairs + (p.subtiles - full_pairs * kSub));
  conv_umma_fwd_kernel<BN><<<grid, kThreads, Cfg::kSmemBytes, st>>>(q);
  IDEAS_CHECK_LAUNCH("conv_umma_fwd");
  return IDEAS_OK;
}


std::atomic<int> g_pair_epi{1};        // option "pair_epi": pair kernel epilogue through shared memory + TMA store (1) or direct stores (0)
std::atomic<int> g_pair_stages{0};     // option "pair_stages": TMA ring depth of the pair kernel (0 = default)
std::atomic<int> g_pair{1};            // option "pair": 256-channel output tiles on persistent CTA pairs (cta_group::2)

template <int BN>
int launch_fwd_pair(const FwdParams& p, int ntiles_n, cudaStream_t st) {
  using Cfg = PairCfg<BN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_umma_fwd_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kMaxStages * Cfg::kStageBytes + Cfg::kEpiBytes + 1024);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "conv_umma_pair: cudaFuncSetAttribute");
  PairParams q;
  q.f = p;
  q.ktiles = ntiles_n;
  const int64_t items = (int64_t)ceil_div(p.subtiles, 2) * ntiles_n;
  if (items > (1ll << 30)) return IDEAS_ERR_UNSUPPORTED;
  q.items = (int)items;
  const int want = g_pair_stages.load();
  q.stages = want >= 2 && want <= Cfg::kMaxStages ? want : Cfg::kStages;
  // TMA-store epilogue: the output seen as {OC, QW, QH, N} -- phase offset in the base, destination stride in the
  // strides -- so a warp's 32 accumulator rows (bw x wy x wn pixels) are one box and the edges are clipped by the unit
  q.epi = g_pair_epi.load() ? 1 : 0;
  q.wy = p.bh < 32 / p.bw ? p.bh : 32 / p.bw;
  q.wn = 32 / (p.bw * q.wy);
  if (q.epi) {
    const uint64_t dims[4] = {(uint64_t)p.OC, (uint64_t)p.QW, (uint64_t)p.QH, (uint64_t)p.N};
    const uint64_t strides[3] = {(uint64_t)p.o_s * p.OC * 4, (uint64_t)p.o_s * p.OW * p.OC * 4, (uint64_t)p.OH * p.OW * p.OC * 4};
    const uint32_t box[4] = {32u, (uint32_t)p.bw, (uint32_t)q.wy, (uint32_t)q.wn};
    float* base = p.dst + ((int64_t)p.o_py * p.OW + p.o_px) * p.OC;
    int rc = encode_map(&q.dst, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, true);
    if (rc) return rc;
  } else {
    memset(&q.dst, 0, sizeof(q.dst));
  }
  const int clusters = q.items < kNumSMs / 2 ? q.items : kNumSMs / 2;
  conv_umma_fwd_pair_kernel<BN><<<2 * clusters, kThreads, q.stages * Cfg::kStageBytes + Cfg::kEpiBytes + 1024, st>>>(q);
  IDEAS_CHECK_LAUNCH("conv_umma_fwd_pair");
  return IDEAS_OK;
}

// ==========================================================================================
// Halo-reuse kernel (source stride 1, several taps): the implicit GEMM above fetches every shifted
// 128-pixel patch again for each filter tap -- 9x the activation bytes of a 3x3 filter through L2 --
// and its prologue/epilogue are exposed once per CTA.  Here the roles are swapped,
//   D[k, pixel] = sum_{tap, c} Wp[tap][k][c] * X[pixel + tap, c]        (M = 128 output channels,
//                                                                        N = pixels of the tile)
// and ONE halo'd activation slab per 32-channel step -- a TMA box of (bh + halo_y) image rows x
// (bw + halo_x) columns, rows of 128 B in K-major SWIZZLE_128B -- serves every tap: the B descriptor
// of tap (dy,dx) simply starts (dy*bwp + dx) rows further (tcgen05 computes the swizzle from the
// address bits, so a start address that is not 1024-byte aligned is legal with base_offset 0;
// scripts/probe_umma_offset.cu is the measurement).  The N extent therefore runs over the padded
// row width bwp: the halo_x columns of every row produce outputs that are dropped (bw/bwp of the
// MMA work is useful) in exchange for ~9x fewer activation bytes.
// The kernel is persistent (one CTA per SM walks a static list of (k-tile, pixel-tile) items) with
// two TMEM accumulator buffers, so the epilogue of item i overlaps the main loop of item i+1 and
// the TMA producers run ahead across items:
//   warp 0      weight-tile TMA producer (ring of up to kHaloMaxWStages x 16 KB)
//   warp 1      activation-slab TMA producer (ring of 2-3 slabs)
//   warp 2      MMA issuer + TMEM owner
//   warps 3..10 epilogue: tcgen05.ld (lane = output channel, column = pixel) -> d[n,k] -> +bias ->
//               lrelu*gain -> NHWC stores (a warp writes 32 consecutive channels of one pixel: 128 B)
// ==========================================================================================
constexpr int kHaloMaxWStages = 12;         // weight ring: as deep as shared memory allows (TMA latency >> one tap)
constexpr int kHaloThreads = 352;
constexpr int kHaloEpiWarps = 8;
constexpr uint32_t kHaloWBytes = 128 * 128;   // 128 output channels x 32 fp32 input channels
constexpr int kHaloSlack = 24;                // rows past the box that rounding N up to 16 may touch
constexpr uint32_t kHaloSmemBudget = 224 * 1024;
constexpr uint32_t kHaloStageBytes = kHaloEpiWarps * 32 * 128;   // one 32 x 32 fp32 transpose tile per epilogue warp
std::atomic<int> g_dgrad_phases{1};          // option "dgrad_phases": strided data gradients: all phases in one halo launch
std::atomic<int> g_halo_epi{1};               // option "halo_epi": 0 = per-element stores, 1 = TMA-store epilogue on 32-wide tiles and the
                                              // transposed 128-bit-store epilogue elsewhere, 3 = transposed epilogue everywhere

std::atomic<int> g_halo_mode{1};              // 0 = never, 1 = heuristic, 2 = whenever eligible

struct TapH {
  int32_t rowoff;   // (dy - min_dy) * bwp + (dx - min_dx)
  int32_t widx;
};

struct alignas(64) HaloParams {
  CUtensorMap src;
  CUtensorMap w;
  float* dst;
  const float* out_scale;
  const float* bias;
  int N, QH, QW;
  int OH, OW, OC;
  int o_s, o_py, o_px;
  int bw, bwp, bh;
  int nmma, box_rows;
  int x_org, y_org;
  int tiles_x, tiles_y, ktiles, items;
  int ntaps, csteps;
  int x_stages, w_stages;
  uint32_t x_slot_bytes, x_tx_bytes;
  int act;
  int epi;
  float alpha, gain;
  // Output phases served by ONE launch (nphase > 1: the stride^2 phases of a strided data gradient / transposed conv).
  // Every phase reads the same source tensor, so running the phases of a pixel tile on neighbouring CTAs at the same
  // time turns three of the four reads of each activation slab into L2 hits (separate launches re-read the whole
  // tensor from HBM once per phase: it is several times the size of the L2).  ph[0] mirrors the scalar fields above.
  int nphase;
  struct Phase { int ntaps, tap0, o_py, o_px, QH, QW, x_org, y_org; } ph[4];
  TapH taps[kMaxTaps];
  CUtensorMap dmap[4];    // epi == 2: the output of each phase as {OC, QW, QH, N} for the TMA-store epilogue
};

__device__ __forceinline__ void st_global_pred(float* p, float v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q st.global.f32 [%0], %1;\n\t}"
      ::"l"(p), "f"(v), "r"((int)pred)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(kHaloThreads, 1) conv_umma_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t wfull[kHaloMaxWStages];
  __shared__ __align__(8) uint64_t wempty[kHaloMaxWStages];
  __shared__ __align__(8) uint64_t xfull[3];
  __shared__ __align__(8) uint64_t xempty[3];
  __shared__ __align__(8) uint64_t tfull[2];
  __shared__ __align__(8) uint64_t tempty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wring = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xring = wring + p.w_stages * kHaloWBytes;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kHaloMaxWStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&wfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&wempty[s]), 1);
    }
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(ptx::smem_u32(&xfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&xempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty[s]), kHaloEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // item -> (k-tile, pixel tile); the k-tile runs fastest so that concurrently running CTAs share activations
  // item -> (phase, k-tile, pixel tile).  The nphase items of a (k-tile, pixel tile) are consecutive, i.e. run on
  // neighbouring CTAs in the same round; the phase is rotated by the round so that every CTA sees all phases in turn
  // (they differ in cost: 1, 2, 2 and 4 taps for a 3x3 stride-2 filter).  gridDim.x is a multiple of nphase.
#define IDEAS_HALO_DECODE(item)                                                       \
  const int phs = ((item) % p.nphase + (item) / (int)gridDim.x) % p.nphase;           \
  const int tix = (item) / p.nphase;                                                  \
  const int kt = tix % p.ktiles;                                                      \
  const int pt = tix / p.ktiles;                                                      \
  const int bx = pt % p.tiles_x;                                                      \
  const int tt = pt / p.tiles_x;                                                      \
  const int qx0 = bx * p.bw;                                                          \
  const int qy0 = (tt % p.tiles_y) * p.bh;                                            \
  const int n = tt / p.tiles_y;                                                       \
  const int k0 = kt * 128;

  if (warp == 0) {
    if (lane == 0) {
      // ===== weight tiles: one (tap, 32-channel step) per ring slot =====
      ptx::tma_prefetch_desc(&p.w);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        IDEAS_HALO_DECODE(item)
        (void)qx0; (void)qy0; (void)n;
        const int nt = p.ph[phs].ntaps, t0 = p.ph[phs].tap0;
        for (int cs = 0; cs < p.csteps; ++cs)
          for (int t = 0; t < nt; ++t) {
            ptx::mbar_wait(ptx::smem_u32(&wempty[stage]), phase ^ 1u);
            const uint32_t fb = ptx::smem_u32(&wfull[stage]);
            ptx::mbar_arrive_expect_tx(fb, kHaloWBytes);
            ptx::tma_load_3d(wring + stage * kHaloWBytes, &p.w, fb, cs * kBlockK, k0, p.taps[t0 + t].widx);
            if (++stage == p.w_stages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== activation slabs: one halo'd box per 32-channel step =====
      ptx::tma_prefetch_desc(&p.src);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        IDEAS_HALO_DECODE(item)
        (void)k0;
        for (int cs = 0; cs < p.csteps; ++cs) {
          ptx::mbar_wait(ptx::smem_u32(&xempty[stage]), phase ^ 1u);
          const uint32_t fb = ptx::smem_u32(&xfull[stage]);
          ptx::mbar_arrive_expect_tx(fb, p.x_tx_bytes);
          ptx::tma_load_4d(xring + stage * p.x_slot_bytes, &p.src, fb, cs * kBlockK, qx0 + p.ph[phs].x_org,
                           qy0 + p.ph[phs].y_org, n);
          if (++stage == p.x_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = ptx::idesc_tf32(128, p.nmma, 0, 0);
      int ws = 0, xs = 0;
      uint32_t wphase = 0, xphase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const int phs = (item % p.nphase + item / (int)gridDim.x) % p.nphase;
        const int nt = p.ph[phs].ntaps, t0 = p.ph[phs].tap0;
        ptx::mbar_wait(ptx::smem_u32(&tempty[buf]), ((uint32_t)(it >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * 256;
        for (int cs = 0; cs < p.csteps; ++cs) {
          ptx::mbar_wait(ptx::smem_u32(&xfull[xs]), xphase);
          const uint32_t xa = xring + xs * p.x_slot_bytes;
          for (int t = 0; t < nt; ++t) {
            ptx::mbar_wait(ptx::smem_u32(&wfull[ws]), wphase);
            ptx::tc_fence_after();
            const uint64_t adesc = ptx::smem_desc_sw128(wring + ws * kHaloWBytes, 16, 1024);
            const uint64_t bdesc = ptx::smem_desc_sw128(xa + (uint32_t)p.taps[t0 + t].rowoff * 128u, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              ptx::mma_tf32(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((cs | t | k) != 0));
            ptx::mma_commit(ptx::smem_u32(&wempty[ws]));
            if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
          }
          ptx::mma_commit(ptx::smem_u32(&xempty[xs]));   // the slab is free once every tap has consumed it
          if (++xs == p.x_stages) { xs = 0; xphase ^= 1u; }
        }
        ptx::mma_commit(ptx::smem_u32(&tfull[buf]));
      }
    }
  } else {
    // ===== epilogue: warps 3..10; TMEM lane quarter = warp % 4 (lane = output channel); the two warps of a
    // quarter take alternate 16-column chunks.  Within a chunk the 16 pixels span at most two tile rows
    // (bwp >= 16), so every address is an independent function of the chunk origin: no serial chain. =====
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int npix = p.bh * p.bwp;
    const int chunks = (npix + 15) >> 4;
    const int pix_step = p.o_s * p.OC;
    const int row_step = p.o_s * p.OW * p.OC;
    const bool lrelu_on = p.act == IDEAS_ACT_LRELU;
    const float alpha = p.alpha, gain = lrelu_on ? p.gain : 1.f;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      IDEAS_HALO_DECODE(item)
      const int buf = it & 1;
      const int k = k0 + quarter * 32 + lane;
      const bool kok = k < p.OC;
      float os = 1.f, bs = 0.f;
      if (kok) {
        if (p.out_scale) os = __ldg(p.out_scale + (int64_t)n * p.OC + k);
        if (p.bias) bs = __ldg(p.bias + k);
      }
      const int o_py = p.ph[phs].o_py, o_px = p.ph[phs].o_px;
      const int xlim = min(p.bw, p.ph[phs].QW - qx0);
      const int ylim = min(p.bh, p.ph[phs].QH - qy0);
      float* base = p.dst + (((int64_t)n * p.OH + ((int64_t)qy0 * p.o_s + o_py)) * p.OW + ((int64_t)qx0 * p.o_s + o_px)) * p.OC + k;
      ptx::mbar_wait(ptx::smem_u32(&tfull[buf]), (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256;
      if (p.epi == 2) {
        // TMA-store epilogue (tile width a multiple of 32): segment sx of tile row y is the 32 accumulator columns
        // from y*bwp + sx on, i.e. 32 consecutive pixels of one image row.  The thread (= one output channel) finishes them and writes
        // them as a column of the warp's 32 pixel x 32 channel staging tile in the 128-byte-swizzle pattern (a warp
        // store is one permuted 128-byte row: conflict-free); one cp.async.bulk.tensor store then writes the
        // {32 channels, 32 pixels} box.  The unit clips the box at the image edges and at OC, and the tensor map of
        // the phase carries the destination stride, so there is no address arithmetic, no predicate and no read-back.
        const uint32_t stage = xring + p.x_stages * p.x_slot_bytes + (uint32_t)(warp - 3) * 4096u;
        const uint32_t colw = (uint32_t)(lane & 3) * 4u, chunk = (uint32_t)(lane >> 2);
        const int nseg = p.bw >> 5;                              // 32-pixel segments per tile row (bw % 32 == 0)
        const int nunits = xlim > 0 ? ylim * nseg : 0;           // rows / tiles outside this phase's grid: nothing to store
        for (int u = half; u < nunits; u += 2) {
          const int y = u / nseg, sx = (u - y * nseg) << 5;
          float v[32];
          ptx::tmem_ld_32x32(acc + (uint32_t)(y * p.bwp + sx), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float r = fmaf(v[j], os, bs);
            r = (lrelu_on && r < 0.f) ? r * alpha : r;
            v[j] = r * gain;
          }
          if (lane == 0) ptx::tma_store_wait_read<0>();        // the previous store has drained the staging tile
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            asm volatile("st.shared.f32 [%0], %1;"
                         ::"r"(stage + (uint32_t)(j * 128) + ((chunk ^ (uint32_t)(j & 7)) << 4) + colw), "f"(v[j])
                         : "memory");
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_4d(&p.dmap[phs], stage, k0 + quarter * 32, qx0 + sx, qy0 + y, n);
            ptx::tma_store_commit();
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ptx::smem_u32(&tempty[buf]));
        continue;
      }
      if (p.epi) {
        // Transposed epilogue: tcgen05.ld hands a thread (= one output channel) 32 consecutive pixels.  The thread
        // finishes them (demodulation, bias, leaky ReLU: per-channel scalars), writes them as one COLUMN of a private
        // 32 pixel x 32 channel shared-memory tile (a warp store = one 128-byte row segment, conflict-free), and the
        // tile comes back row-wise: lane l owns the 16-byte chunk l%8 of pixels l/8 + 4j, so one global store
        // instruction writes four complete 128-byte lines instead of 32 four-byte pieces, and the pixel -> address
        // arithmetic runs once per pixel group instead of once per element.
        const uint32_t stage = xring + p.x_stages * p.x_slot_bytes + (uint32_t)(warp - 3) * 4096u;
        const int cj = lane & 7, pr = lane >> 3;
        const int kc = k0 + quarter * 32 + cj * 4;               // first of this lane's four channels on the way out
        const bool kcok = kc < p.OC;
        const int units = (npix + 31) >> 5;
        const int64_t rowwrap = (int64_t)row_step - (int64_t)p.bwp * pix_step;
        float* const obase = p.dst + (((int64_t)n * p.OH + ((int64_t)qy0 * p.o_s + o_py)) * p.OW + ((int64_t)qx0 * p.o_s + o_px)) * p.OC + kc;
        for (int u = half; u < units; u += 2) {
          float v[32];
          ptx::tmem_ld_32x32(acc + (uint32_t)(u << 5), v);
          __syncwarp();                                   // the previous unit's reads of the tile are done
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float r = fmaf(v[j], os, bs);
            r = (lrelu_on && r < 0.f) ? r * alpha : r;
            r *= gain;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(stage + (uint32_t)(j * 128 + lane * 4)), "f"(r) : "memory");
          }
          __syncwarp();
          const int m = (u << 5) + pr;
          int y = m / p.bwp;
          int x = m - y * p.bwp;
          float* op = obase + (int64_t)y * row_step + (int64_t)x * pix_step;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                         : "r"(stage + (uint32_t)((4 * j + pr) * 128 + cj * 16)));
            if (kcok && x < xlim && y < ylim) *reinterpret_cast<float4*>(op) = r;
            x += 4;
            op += 4 * pix_step;
            if (x >= p.bwp) { x -= p.bwp; ++y; op += rowwrap; }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ptx::smem_u32(&tempty[buf]));
        continue;
      }
      for (int c = half; c < chunks; c += 2) {
        const int m0 = c << 4;
        const int y0 = m0 / p.bwp;
        const int x0 = m0 - y0 * p.bwp;
        float v[16];
        ptx::tmem_ld_32x16(acc + (uint32_t)m0, v);
        float* r0 = base + (int64_t)y0 * row_step;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int xi = x0 + i;
          const int wrap = xi >= p.bwp ? 1 : 0;
          xi -= wrap ? p.bwp : 0;
          const bool ok = kok && xi < xlim && (y0 + wrap) < ylim;
          float r = fmaf(v[i], os, bs);
          r = (lrelu_on && r < 0.f) ? r * alpha : r;
          r *= gain;
          // predicated store, no branch: with two epilogue warps per scheduler a branch per element showed up as
          // branch-resolving stalls on every store (ncu source view), and the epilogue warps were busy 70 % of the
          // kernel at 128 input channels -- the bottleneck below that
          st_global_pred(r0 + (wrap * row_step + xi * pix_step), r, ok);
        }
      }
      // every tcgen05.ld of this warp has completed (wait::ld): hand the accumulator back
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ptx::smem_u32(&tempty[buf]));
    }
    if (p.epi == 2 && lane == 0) ptx::tma_store_wait<0>();     // every bulk store of this warp has completed
  }
#undef IDEAS_HALO_DECODE
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

struct HaloTile {
  int bw, bwp, bh, nmma, tiles_x, tiles_y, x_stages, w_stages;
  uint32_t slot_bytes;
  double cycles;   // modelled cycles per 32-channel step over the q grid of one image
};

// Pick (bw, bh) minimising modelled time = tiles * max(MMA cycles, L2-feed cycles) per channel step.
bool choose_halo_tile(int QW, int QH, int hx, int hy, int ntaps, HaloTile* out) {
  struct Key { int a, b, c, d, e; bool operator<(const Key& o) const { return memcmp(this, &o, sizeof(Key)) < 0; } };
  static std::mutex mu;
  static std::map<Key, HaloTile> cache;
  const Key key{QW, QH, hx, hy, ntaps};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return it->second.bw > 0; }
  }
  HaloTile best{};
  best.bw = 0; best.cycles = 1e300;
  for (int bw = 4; bw <= QW && bw + hx <= 256; ++bw) {
    const int bwp = bw + hx;
    if (bwp < 16) continue;   // the epilogue assumes a 16-pixel chunk spans at most two rows
    for (int bh = 1; bh * bwp <= 256 && bh - 1 < QH; ++bh) {
      const int nmma = ((bh * bwp + 15) / 16) * 16;
      const int rows = bh + hy;
      const uint32_t slot = (uint32_t)((rows * bwp + hx + kHaloSlack) * 128 + 1023) & ~1023u;
      if (2 * slot + 4 * kHaloWBytes + kHaloStageBytes > kHaloSmemBudget) break;
      const int tx = ceil_div(QW, bw), ty = ceil_div(QH, bh);
      const double mma = (double)ntaps * 4 * (nmma / 2.0);
      const double l2 = ((double)ntaps * kHaloWBytes + (double)rows * bwp * 128) / 48.0;
      const double cyc = (double)tx * ty * ((mma > l2 ? mma : l2) + 24.0);
      if (cyc < best.cycles) {
        best.bw = bw; best.bwp = bwp; best.bh = bh; best.nmma = nmma; best.tiles_x = tx; best.tiles_y = ty;
        best.slot_bytes = slot; best.cycles = cyc; best.x_stages = 2;
        const int ws = (int)((kHaloSmemBudget - kHaloStageBytes - 2 * slot) / kHaloWBytes);
        best.w_stages = ws > kHaloMaxWStages ? kHaloMaxWStages : ws;
      }
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = best;
  }
  *out = best;
  return best.bw > 0;
}

// nph == 1: one geometry.  nph > 1: the output phases of one strided data gradient (same tensors, same channel
// counts; each phase has its own taps, q grid and destination offset), served by a single launch.
int halo_conv_launch_phases(const ConvGeom* gs, int nph, float* dst, const float* src, const float* w,
                            const float* out_scale, const float* bias, int act, float alpha, float gain,
                            cudaStream_t st) {
  const int mode = g_halo_mode.load();
  const ConvGeom& g = gs[0];
  if (mode == 0 || nph < 1 || nph > 4) return IDEAS_ERR_UNSUPPORTED;
  if (nph == 1 && g.ntaps < 2) return IDEAS_ERR_UNSUPPORTED;
  int hx = 0, hy = 0, max_widx = 0, QWm = 0, QHm = 0, total_taps = 0, max_taps = 0;
  int min_dx[4], min_dy[4];
  for (int f = 0; f < nph; ++f) {
    const ConvGeom& q = gs[f];
    if (q.i_s != 1 || q.ntaps < 1 || q.QW < 1 || q.QH < 1) return IDEAS_ERR_UNSUPPORTED;
    if (q.N != g.N || q.IC != g.IC || q.OC != g.OC || q.IH != g.IH || q.IW != g.IW || q.OH != g.OH ||
        q.OW != g.OW || q.o_s != g.o_s)
      return IDEAS_ERR_UNSUPPORTED;
    int lo_x = 1 << 30, hi_x = -(1 << 30), lo_y = 1 << 30, hi_y = -(1 << 30);
    for (int t = 0; t < q.ntaps; ++t) {
      lo_x = q.taps[t].dx < lo_x ? q.taps[t].dx : lo_x; hi_x = q.taps[t].dx > hi_x ? q.taps[t].dx : hi_x;
      lo_y = q.taps[t].dy < lo_y ? q.taps[t].dy : lo_y; hi_y = q.taps[t].dy > hi_y ? q.taps[t].dy : hi_y;
      max_widx = q.taps[t].widx > max_widx ? q.taps[t].widx : max_widx;
    }
    min_dx[f] = lo_x; min_dy[f] = lo_y;
    hx = hi_x - lo_x > hx ? hi_x - lo_x : hx;
    hy = hi_y - lo_y > hy ? hi_y - lo_y : hy;
    QWm = q.QW > QWm ? q.QW : QWm; QHm = q.QH > QHm ? q.QH : QHm;
    total_taps += q.ntaps;
    max_taps = q.ntaps > max_taps ? q.ntaps : max_taps;
  }
  if (QWm < 8 || total_taps > kMaxTaps) return IDEAS_ERR_UNSUPPORTED;
  if (hx > 8 || hy > 8) return IDEAS_ERR_UNSUPPORTED;
  HaloTile tile;
  if (!choose_halo_tile(QWm, QHm, hx, hy, max_taps, &tile)) return IDEAS_ERR_UNSUPPORTED;
  if (mode == 1) {
    // heuristic from scripts/kernels_microbench.py (IDEAS_HALO=0 vs 2): with <= 128 input channels the
    // pixel-major kernel is bound by activation re-reads and by its exposed epilogue (halo kernel 1.2-1.7x);
    // from 256 input channels on, weight tiles dominate the L2 traffic of both kernels and the halo columns
    // only cost tensor work (0.8-0.95x).  Strided destinations (data gradients of stride-2 convs, 1-4 taps
    // per phase) have short reductions, where the overlapped epilogue wins (1.25x).
    if (g.OC < 64 || !(g.IC <= 128 || g.o_s == 2)) return IDEAS_ERR_UNSUPPORTED;
  }

  HaloParams p;
  p.dst = dst; p.out_scale = out_scale; p.bias = bias;
  p.N = g.N; p.QH = QHm; p.QW = QWm; p.OH = g.OH; p.OW = g.OW; p.OC = g.OC;
  p.o_s = g.o_s; p.o_py = g.o_py; p.o_px = g.o_px;
  p.bw = tile.bw; p.bwp = tile.bwp; p.bh = tile.bh; p.nmma = tile.nmma;
  p.box_rows = tile.bh + hy;
  p.x_org = min_dx[0]; p.y_org = min_dy[0];
  p.tiles_x = tile.tiles_x; p.tiles_y = tile.tiles_y;
  p.ktiles = ceil_div(g.OC, 128);
  const int64_t items = (int64_t)g.N * p.tiles_y * p.tiles_x * p.ktiles * nph;
  if (items > (1ll << 30)) return IDEAS_ERR_UNSUPPORTED;
  p.items = (int)items;
  p.ntaps = g.ntaps; p.csteps = g.IC / kBlockK;
  p.x_stages = tile.x_stages;
  p.w_stages = tile.w_stages;
  p.x_slot_bytes = tile.slot_bytes;
  p.x_tx_bytes = (uint32_t)(p.box_rows * p.bwp * 128);
  p.act = act; p.alpha = alpha; p.gain = gain;
  p.nphase = nph;
  int tap0 = 0;
  for (int f = 0; f < 4; ++f) {
    const ConvGeom& q = gs[f < nph ? f : 0];
    const int ff = f < nph ? f : 0;
    p.ph[f].ntaps = q.ntaps; p.ph[f].tap0 = f < nph ? tap0 : 0;
    p.ph[f].o_py = q.o_py; p.ph[f].o_px = q.o_px; p.ph[f].QH = q.QH; p.ph[f].QW = q.QW;
    p.ph[f].x_org = min_dx[ff]; p.ph[f].y_org = min_dy[ff];
    if (f >= nph) continue;
    for (int t = 0; t < q.ntaps; ++t) {
      p.taps[tap0 + t].rowoff = (q.taps[t].dy - min_dy[f]) * p.bwp + (q.taps[t].dx - min_dx[f]);
      p.taps[tap0 + t].widx = q.taps[t].widx;
    }
    tap0 += q.ntaps;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g.IC, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)g.IC * 4, (uint64_t)g.IW * g.IC * 4, (uint64_t)g.IH * g.IW * g.IC * 4};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bwp, (uint32_t)p.box_rows, 1u};
    int rc = encode_map(&p.src, src, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.IC, (uint64_t)g.OC, (uint64_t)(max_widx + 1)};
    const uint64_t strides[2] = {(uint64_t)g.IC * 4, (uint64_t)g.OC * g.IC * 4};
    const uint32_t wbox[3] = {(uint32_t)kBlockK, 128u, 1u};
    int rc = encode_map(&p.w, w, 3, dims, strides, wbox);
    if (rc) return rc;
  }
  const uint32_t smem = p.w_stages * kHaloWBytes + p.x_stages * p.x_slot_bytes + kHaloStageBytes + 1024;
  p.epi = g_halo_epi.load();
  if (p.epi >= 1 && g_halo_epi.load() != 3 && p.bw % 32 == 0) {
    // tile rows made of 32-pixel segments (the tile search picks 32-wide tiles for 256, 64 and 32 pixel rows):
    // TMA-store epilogue
    p.epi = 2;
    for (int f = 0; f < 4; ++f) {
      const ConvGeom& q = gs[f < nph ? f : 0];
      const uint64_t dims[4] = {(uint64_t)q.OC, (uint64_t)q.QW, (uint64_t)q.QH, (uint64_t)q.N};
      const uint64_t strides[3] = {(uint64_t)q.o_s * q.OC * 4, (uint64_t)q.o_s * q.OW * q.OC * 4,
                                   (uint64_t)q.OH * q.OW * q.OC * 4};
      const uint32_t box[4] = {32u, 32u, 1u, 1u};
      float* base = dst + ((int64_t)q.o_py * q.OW + q.o_px) * q.OC;
      int rc = encode_map(&p.dmap[f], base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, true);
      if (rc) return rc;
    }
  } else {
    if (p.epi > 1) p.epi = 1;
    memset(p.dmap, 0, sizeof(p.dmap));
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_umma_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kHaloSmemBudget + 1024);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "conv_umma_halo: cudaFuncSetAttribute");
  int grid = p.items < kNumSMs ? p.items : kNumSMs;
  grid -= grid % nph;            // the phase rotation in the kernel needs a whole number of phase groups per round
  conv_umma_halo_kernel<<<grid, kHaloThreads, smem, st>>>(p);
  IDEAS_CHECK_LAUNCH("conv_umma_halo");
  return IDEAS_OK;
}

int halo_conv_launch(const ConvGeom& g, float* dst, const float* src, const float* w, const float* out_scale,
                     const float* bias, int act, float alpha, float gain, cudaStream_t st) {
  return halo_conv_launch_phases(&g, 1, dst, src, w, out_scale, bias, act, alpha, gain, st);
}


// ==========================================================================================
// Pixel-major halo kernel ("pmh"): the halo-reuse idea with the GEMM roles of the forward-form kernel,
//   D[pixel, k] = sum_{tap, c} X[pixel + tap, c] * Wp[tap][k][c]     (M = 128 slab positions, N = OCT output channels)
// The halo kernel above puts the output channels on the 128 TMEM lanes: with 32 or 64 of them half or three
// quarters of every MMA is spent on zero rows (tcgen05 costs max(M,128)*N/256 clk whatever M is), and its
// epilogue (lane = channel) stores 4 bytes per thread per instruction.  Here the A operand is the halo'd
// activation slab itself: sub-tile s of tap (dy,dx) is the 128 CONSECUTIVE slab rows starting at
// s*128 + dy*bwp + dx (address-based swizzle, as above), so accumulator row m is slab position m = y*bwp + x,
// N = OCT in {32, 64, 128} is exactly the layer's width, and the epilogue is pixel-major: one thread = one pixel,
// 32 consecutive channels per tcgen05.ld, 128-bit NHWC stores.  Slab positions in the halo columns (x >= bw)
// and past the tile produce rows that are dropped; rows read past the TMA box only feed such rows (row m of D
// depends on row m of A alone).  Persistent, two TMEM buffers of nsub accumulators (nsub*OCT <= 256 columns
// each), weight tiles and slabs in TMA rings, epilogue of item i under the MMAs of item i+1 -- as above.
// The epilogue also takes a residual operand: out = (act(d*acc + bias) + residual) * res_scale, the merge
// (out + skip)/sqrt(2) of the residual blocks (models.py:178,227) without its own pass over HBM.
// ==========================================================================================
constexpr int kPmhMaxWStages = 16;
constexpr int kPmhThreads = 352;
constexpr int kPmhEpiWarps = 8;
constexpr uint32_t kPmhSmemBudget = 224 * 1024;

std::atomic<int> g_pmh_mode{1};               // 0 = never, 1 = heuristic, 2 = whenever eligible
std::atomic<int> g_pmh_resident{1};           // keep small filters in shared memory for the CTA's lifetime
std::atomic<int> g_pmh_xstages{2};            // slab ring depth (2 or 3)
std::atomic<int> g_pmh_minbw{0};              // tile search: smallest tile width considered (0 = no bound)
std::atomic<int> g_pmh_align{0};              // tile search: padded row width must be a multiple of this (0 = any)

struct alignas(64) PmhParams {
  CUtensorMap src;
  CUtensorMap w;
  float* dst;
  const float* out_scale;
  const float* bias;
  const float* residual;
  float res_scale;
  int N, QH, QW;
  int OH, OW, OC;
  int o_s, o_py, o_px;
  int bw, bwp, bh, nsub, oct;
  int box_rows;
  int x_org, y_org;
  int tiles_x, tiles_y, ktiles, items;
  int ntaps, csteps;
  int x_stages, w_stages;
  int w_resident;   // all ntaps*csteps weight tiles stay in shared memory for the CTA's lifetime (one k-tile, <= 80 KB)
  uint32_t x_slot_bytes, x_tx_bytes, w_bytes;
  int act;
  float alpha, gain;
  TapH taps[kMaxTaps];
};

constexpr uint32_t kPmhStageBytes = kPmhEpiWarps * 32 * 128;   // one 32 x 32 fp32 transpose tile per epilogue warp

__global__ void __launch_bounds__(kPmhThreads, 1) conv_umma_pmh_kernel(const __grid_constant__ PmhParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t wfull[kPmhMaxWStages];
  __shared__ __align__(8) uint64_t wempty[kPmhMaxWStages];
  __shared__ __align__(8) uint64_t xfull[3];
  __shared__ __align__(8) uint64_t xempty[3];
  __shared__ __align__(8) uint64_t tfull[2];
  __shared__ __align__(8) uint64_t tempty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wring = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xring = wring + p.w_stages * p.w_bytes;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kPmhMaxWStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&wfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&wempty[s]), 1);
    }
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(ptx::smem_u32(&xfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&xempty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty[s]), kPmhEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

#define IDEAS_PMH_DECODE(item)                          \
  const int kt = (item) % p.ktiles;                     \
  const int pt = (item) / p.ktiles;                     \
  const int bx = pt % p.tiles_x;                        \
  const int tt = pt / p.tiles_x;                        \
  const int qx0 = bx * p.bw;                            \
  const int qy0 = (tt % p.tiles_y) * p.bh;              \
  const int n = tt / p.tiles_y;                         \
  const int k0 = kt * p.oct;

  if (warp == 0) {
    if (lane == 0) {
      // ===== weight tiles: one (32-channel step, tap) per ring slot, OCT rows of 128 B =====
      ptx::tma_prefetch_desc(&p.w);
      if (p.w_resident) {
        // the whole filter fits: fetch it once, every item reuses it (32 -> 64 / 64 -> 32 channel layers, where an
        // item's weight tiles would otherwise cost as much L2 traffic as its activations)
        const uint32_t fb = ptx::smem_u32(&wfull[0]);
        ptx::mbar_arrive_expect_tx(fb, (uint32_t)(p.csteps * p.ntaps) * p.w_bytes);
        for (int cs = 0; cs < p.csteps; ++cs)
          for (int t = 0; t < p.ntaps; ++t)
            ptx::tma_load_3d(wring + (uint32_t)(cs * p.ntaps + t) * p.w_bytes, &p.w, fb, cs * kBlockK, 0, p.taps[t].widx);
      } else {
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
          const int k0 = (item % p.ktiles) * p.oct;
          for (int cs = 0; cs < p.csteps; ++cs)
            for (int t = 0; t < p.ntaps; ++t) {
              ptx::mbar_wait(ptx::smem_u32(&wempty[stage]), phase ^ 1u);
              const uint32_t fb = ptx::smem_u32(&wfull[stage]);
              ptx::mbar_arrive_expect_tx(fb, p.w_bytes);
              ptx::tma_load_3d(wring + stage * p.w_bytes, &p.w, fb, cs * kBlockK, k0, p.taps[t].widx);
              if (++stage == p.w_stages) { stage = 0; phase ^= 1u; }
            }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== activation slabs: one halo'd box per 32-channel step =====
      ptx::tma_prefetch_desc(&p.src);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        IDEAS_PMH_DECODE(item)
        (void)k0;
        for (int cs = 0; cs < p.csteps; ++cs) {
          ptx::mbar_wait(ptx::smem_u32(&xempty[stage]), phase ^ 1u);
          const uint32_t fb = ptx::smem_u32(&xfull[stage]);
          ptx::mbar_arrive_expect_tx(fb, p.x_tx_bytes);
          ptx::tma_load_4d(xring + stage * p.x_slot_bytes, &p.src, fb, cs * kBlockK, qx0 + p.x_org, qy0 + p.y_org, n);
          if (++stage == p.x_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = ptx::idesc_tf32(128, p.oct, 0, 0);
      const bool resident = p.w_resident != 0;
      int ws = 0, xs = 0;
      uint32_t wphase = 0, xphase = 0;
      int it = 0;
      if (resident) ptx::mbar_wait(ptx::smem_u32(&wfull[0]), 0);
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty[buf]), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + buf * 256;
        for (int cs = 0; cs < p.csteps; ++cs) {
          ptx::mbar_wait(ptx::smem_u32(&xfull[xs]), xphase);
          const uint32_t xa = xring + xs * p.x_slot_bytes;
          for (int t = 0; t < p.ntaps; ++t) {
            if (resident) ws = cs * p.ntaps + t;
            else ptx::mbar_wait(ptx::smem_u32(&wfull[ws]), wphase);
            ptx::tc_fence_after();
            const uint64_t bdesc = ptx::smem_desc_sw128(wring + ws * p.w_bytes, 16, 1024);
            const uint32_t a0 = xa + (uint32_t)p.taps[t].rowoff * 128u;
            for (int s = 0; s < p.nsub; ++s) {
              const uint64_t adesc = ptx::smem_desc_sw128(a0 + (uint32_t)s * (128u * 128u), 16, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                ptx::mma_tf32(acc + s * p.oct, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((cs | t | k) != 0));
            }
            if (!resident) {
              ptx::mma_commit(ptx::smem_u32(&wempty[ws]));
              if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
            }
          }
          ptx::mma_commit(ptx::smem_u32(&xempty[xs]));
          if (++xs == p.x_stages) { xs = 0; xphase ^= 1u; }
        }
        ptx::mma_commit(ptx::smem_u32(&tfull[buf]));
      }
    }
  } else {
    // ===== epilogue: warps 3..10; TMEM lane quarter = warp % 4 (lane = slab position inside the sub-tile); the two
    // warps of a quarter take alternate (sub-tile, 32-channel chunk) units =====
    // A unit is a 32-position x 32-channel block.  tcgen05.ld hands every thread one position's 32 channels; stored
    // like that, a warp instruction would touch 32 separate 16-byte pieces 4*OC bytes apart.  The block therefore
    // goes through a private 4 KB shared-memory tile (16-byte chunks XOR-swizzled by the row, conflict-free both
    // ways) and comes back transposed: lane l owns chunk l%8 of positions l/8 + 4j, so one instruction writes four
    // complete 128-byte lines -- and the per-channel operands (demodulation, bias) are loaded once per unit.
    const int quarter = warp & 3;
    const int half = (warp - 3) >> 2;
    const int cpt = p.oct >> 5;
    const int units = p.nsub * cpt;
    const bool lrelu_on = p.act == IDEAS_ACT_LRELU;
    const float alpha = p.alpha, gain = lrelu_on ? p.gain : 1.f;
    const float rs = p.res_scale;
    const uint32_t stage = xring + p.x_stages * p.x_slot_bytes + (uint32_t)(warp - 3) * 4096u;
    const int cj = lane & 7;          // 16-byte chunk (4 channels) this lane stores
    const int pr = lane >> 3;         // position within each group of four
    // every address below is a base register plus a compile-time term: the epilogue warps share the issue slots with
    // nothing but each other, and at 32 input channels an item's MMAs last only ~3.5k clk
    const uint32_t st_base = stage + (uint32_t)lane * 128u;
    const uint32_t st_sw = (uint32_t)(lane & 7);
    const uint32_t ld_base = stage + (uint32_t)pr * 128u;
    const uint32_t ld_sw0 = (uint32_t)((cj ^ pr) << 4), ld_sw1 = (uint32_t)((cj ^ (pr | 4)) << 4);
    const int pixstride = p.o_s * p.OC;
    const int64_t rowstride = (int64_t)p.o_s * p.OW * p.OC;
    const int64_t rowwrap = rowstride - (int64_t)p.bwp * pixstride;
    const bool has_res = p.residual != nullptr;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      IDEAS_PMH_DECODE(item)
      const int buf = it & 1;
      const int xlim = min(p.bw, p.QW - qx0);
      const int ylim = min(p.bh, p.QH - qy0);
      const int64_t base = (((int64_t)n * p.OH + ((int64_t)qy0 * p.o_s + p.o_py)) * p.OW + ((int64_t)qx0 * p.o_s + p.o_px)) * p.OC + k0 + cj * 4;
      const float* osn = p.out_scale ? p.out_scale + (int64_t)n * p.OC + k0 + cj * 4 : nullptr;
      const float* bsn = p.bias ? p.bias + k0 + cj * 4 : nullptr;
      ptx::mbar_wait(ptx::smem_u32(&tfull[buf]), (uint32_t)(it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 256;
      for (int u = half; u < units; u += 2) {
        const int sub = u / cpt;
        const int ch = u - sub * cpt;
        float v[32];
        ptx::tmem_ld_32x32(acc + (uint32_t)(sub * p.oct + ch * 32), v);
        __syncwarp();                                   // the previous unit's reads of the tile are done
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_base + ((st_sw ^ (uint32_t)j) << 4)), "f"(v[4 * j]),
                       "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
        __syncwarp();
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (osn) s4 = __ldg(reinterpret_cast<const float4*>(osn + ch * 32));
        if (bsn) b4 = __ldg(reinterpret_cast<const float4*>(bsn + ch * 32));
        const int m = sub * 128 + quarter * 32 + pr;
        int y = m / p.bwp;
        int x = m - y * p.bwp;
        int64_t off = base + (int64_t)y * rowstride + (int64_t)x * pixstride + ch * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                       : "r"(ld_base + (uint32_t)(j * 512) + ((j & 1) ? ld_sw1 : ld_sw0)));
          if (x < xlim && y < ylim) {
            r.x = fmaf(r.x, s4.x, b4.x); r.y = fmaf(r.y, s4.y, b4.y); r.z = fmaf(r.z, s4.z, b4.z); r.w = fmaf(r.w, s4.w, b4.w);
            if (lrelu_on) {
              r.x = lrelu(r.x, alpha); r.y = lrelu(r.y, alpha); r.z = lrelu(r.z, alpha); r.w = lrelu(r.w, alpha);
            }
            r.x *= gain; r.y *= gain; r.z *= gain; r.w *= gain;
            if (has_res) {
              const float4 q = ld_stream4(p.residual + off);
              r.x = (r.x + q.x) * rs; r.y = (r.y + q.y) * rs; r.z = (r.z + q.z) * rs; r.w = (r.w + q.w) * rs;
            }
            *reinterpret_cast<float4*>(p.dst + off) = r;
          }
          x += 4;
          off += 4 * pixstride;
          if (x >= p.bwp) { x -= p.bwp; ++y; off += rowwrap; }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ptx::smem_u32(&tempty[buf]));
    }
  }
#undef IDEAS_PMH_DECODE
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

struct PmhTile {
  int bw, bwp, bh, nsub, tiles_x, tiles_y, w_stages, x_stages;
  uint32_t slot_bytes;
  double cycles;
};

// (bw, bh) minimising modelled time per image = tiles * max(MMA cycles, L2-feed cycles) per 32-channel step
bool choose_pmh_tile(int QW, int QH, int hx, int hy, int ntaps, int oct, int resident_tiles, PmhTile* out) {
  struct Key { int a, b, c, d, e, f; bool operator<(const Key& o) const { return memcmp(this, &o, sizeof(Key)) < 0; } };
  static std::mutex mu;
  static std::map<Key, PmhTile> cache;
  const int xst = g_pmh_xstages.load() == 3 ? 3 : 2;
  const int minbw = g_pmh_minbw.load();
  const int align = g_pmh_align.load();
  const Key key{QW, QH, hx * 16 + hy, ntaps + 64 * resident_tiles, oct + 4096 * align, xst * 1024 + minbw};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return it->second.bw > 0; }
  }
  PmhTile best{};
  best.bw = 0; best.cycles = 1e300;
  const uint32_t w_bytes = (uint32_t)oct * 128u;
  const int max_sub = 256 / oct;
  int bw0 = QW < 4 ? QW : 4;
  if (minbw > bw0) bw0 = minbw < QW ? minbw : QW;
  for (int bw = bw0; bw <= QW && bw + hx <= 256; ++bw) {
    const int bwp = bw + hx;
    if (align > 1 && bwp % align != 0) continue;
    for (int bh = 1; bh <= QH && bh + hy <= 256; ++bh) {
      const int nsub = ceil_div(bh * bwp, 128);
      if (nsub > max_sub) break;
      const uint32_t slot = (uint32_t)((nsub * 128 + hy * bwp + hx) * 128 + 1023) & ~1023u;
      const uint32_t wmin = resident_tiles ? (uint32_t)resident_tiles * w_bytes : 4 * w_bytes;
      if (xst * slot + wmin + kPmhStageBytes > kPmhSmemBudget) break;
      const int tx = ceil_div(QW, bw), ty = ceil_div(QH, bh);
      const double mma = (double)ntaps * nsub * 4 * (oct / 2.0);
      const double l2 = ((resident_tiles ? 0.0 : (double)ntaps * w_bytes) + (double)(bh + hy) * bwp * 128) / 48.0;
      const double cyc = (double)tx * ty * ((mma > l2 ? mma : l2) + 24.0);
      if (cyc < best.cycles) {
        best.bw = bw; best.bwp = bwp; best.bh = bh; best.nsub = nsub; best.tiles_x = tx; best.tiles_y = ty;
        best.slot_bytes = slot; best.cycles = cyc;
        const int ws = (int)((kPmhSmemBudget - kPmhStageBytes - xst * slot) / w_bytes);
        best.w_stages = resident_tiles ? resident_tiles : (ws > kPmhMaxWStages ? kPmhMaxWStages : ws);
        best.x_stages = xst;
      }
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = best;
  }
  *out = best;
  return best.bw > 0;
}

int pmh_conv_launch(const ConvGeom& g, float* dst, const float* src, const float* w, const float* out_scale,
                    const float* bias, const float* residual, float res_scale, int act, float alpha, float gain,
                    cudaStream_t st) {
  const int mode = g_pmh_mode.load();
  if (mode == 0 || g.i_s != 1 || g.ntaps < 1) return IDEAS_ERR_UNSUPPORTED;
  if (g.QW < 4 || g.QH < 1) return IDEAS_ERR_UNSUPPORTED;
  int min_dx = 1 << 30, max_dx = -(1 << 30), min_dy = 1 << 30, max_dy = -(1 << 30), max_widx = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    min_dx = g.taps[t].dx < min_dx ? g.taps[t].dx : min_dx; max_dx = g.taps[t].dx > max_dx ? g.taps[t].dx : max_dx;
    min_dy = g.taps[t].dy < min_dy ? g.taps[t].dy : min_dy; max_dy = g.taps[t].dy > max_dy ? g.taps[t].dy : max_dy;
    max_widx = g.taps[t].widx > max_widx ? g.taps[t].widx : max_widx;
  }
  const int hx = max_dx - min_dx, hy = max_dy - min_dy;
  if (hx > 8 || hy > 8) return IDEAS_ERR_UNSUPPORTED;
  const int oct = (g.OC % 128 == 0) ? 128 : (g.OC % 64 == 0) ? 64 : 32;
  if (mode == 1) {
    // where it wins (scripts/kernels_microbench.py, IDEAS_OPTS pmh=0 vs 2): layers whose output width is not a
    // multiple of 128 -- 32 and 64 output channels, i.e. the first blocks of E / Dco / Dreal and the data gradients
    // of their 64 -> 128 / 32 -> 64 convolutions -- and every residual-merging epilogue
    if (oct == 128 && !(residual && g.IC <= 128)) return IDEAS_ERR_UNSUPPORTED;
    if (g.IC > 128) return IDEAS_ERR_UNSUPPORTED;
  }
  // resident weights: one k-tile and the whole filter within 80 KB (a single barrier covers the one-time load)
  const int wtiles = g.ntaps * (g.IC / kBlockK);
  const bool resident = g_pmh_resident.load() && g.OC == oct && (uint32_t)wtiles * (uint32_t)oct * 128u <= 80u * 1024u;
  PmhTile tile;
  if (!choose_pmh_tile(g.QW, g.QH, hx, hy, g.ntaps, oct, resident ? wtiles : 0, &tile)) return IDEAS_ERR_UNSUPPORTED;

  PmhParams p;
  p.dst = dst; p.out_scale = out_scale; p.bias = bias; p.residual = residual; p.res_scale = res_scale;
  p.N = g.N; p.QH = g.QH; p.QW = g.QW; p.OH = g.OH; p.OW = g.OW; p.OC = g.OC;
  p.o_s = g.o_s; p.o_py = g.o_py; p.o_px = g.o_px;
  p.bw = tile.bw; p.bwp = tile.bwp; p.bh = tile.bh; p.nsub = tile.nsub; p.oct = oct;
  p.box_rows = tile.bh + hy;
  p.x_org = min_dx; p.y_org = min_dy;
  p.tiles_x = tile.tiles_x; p.tiles_y = tile.tiles_y;
  p.ktiles = g.OC / oct;
  const int64_t items = (int64_t)g.N * p.tiles_y * p.tiles_x * p.ktiles;
  if (items > (1ll << 30)) return IDEAS_ERR_UNSUPPORTED;
  p.items = (int)items;
  p.ntaps = g.ntaps; p.csteps = g.IC / kBlockK;
  p.x_stages = tile.x_stages;
  p.w_stages = tile.w_stages;
  p.w_resident = resident ? 1 : 0;
  p.x_slot_bytes = tile.slot_bytes;
  p.x_tx_bytes = (uint32_t)(p.box_rows * p.bwp * 128);
  p.w_bytes = (uint32_t)oct * 128u;
  p.act = act; p.alpha = alpha; p.gain = gain;
  for (int t = 0; t < g.ntaps; ++t) {
    p.taps[t].rowoff = (g.taps[t].dy - min_dy) * p.bwp + (g.taps[t].dx - min_dx);
    p.taps[t].widx = g.taps[t].widx;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g.IC, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)g.IC * 4, (uint64_t)g.IW * g.IC * 4, (uint64_t)g.IH * g.IW * g.IC * 4};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bwp, (uint32_t)p.box_rows, 1u};
    int rc = encode_map(&p.src, src, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.IC, (uint64_t)g.OC, (uint64_t)(max_widx + 1)};
    const uint64_t strides[2] = {(uint64_t)g.IC * 4, (uint64_t)g.OC * g.IC * 4};
    const uint32_t wbox[3] = {(uint32_t)kBlockK, (uint32_t)oct, 1u};
    int rc = encode_map(&p.w, w, 3, dims, strides, wbox);
    if (rc) return rc;
  }
  const uint32_t smem = p.w_stages * p.w_bytes + p.x_stages * p.x_slot_bytes + kPmhStageBytes + 1024;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_umma_pmh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kPmhSmemBudget + 1024);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "conv_umma_pmh: cudaFuncSetAttribute");
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  conv_umma_pmh_kernel<<<grid, kPmhThreads, smem, st>>>(p);
  IDEAS_CHECK_LAUNCH("conv_umma_pmh");
  return IDEAS_OK;
}

}  // namespace

// All output phases of a strided data gradient in one launch of the channel-major halo kernel (option
// "dgrad_phases"): every phase reads the whole dy tensor, and one launch schedules the phases of a pixel tile next
// to each other so the re-reads hit L2.  IDEAS_ERR_UNSUPPORTED = caller falls back to one launch per phase.
int umma_dgrad_phases_launch(const ConvGeom* gs, int nph, float* dst, const float* src, const float* w,
                             const float* out_scale, const float* bias, int act, float alpha, float gain,
                             cudaStream_t st) {
  const int mode = g_dgrad_phases.load();
  if (mode == 0 || nph < 2 || nph > 4) return IDEAS_ERR_UNSUPPORTED;
  const ConvGeom& g = gs[0];
  if (g.IC % 32 || g.OC % 32) return IDEAS_ERR_UNSUPPORTED;
  for (int f = 0; f < nph; ++f)
    if (gs[f].ntaps < 1 || (int64_t)gs[f].N * gs[f].QH * gs[f].QW < 256 || gs[f].QW < 2) return IDEAS_ERR_UNSUPPORTED;
  if (!aligned16(dst) || !aligned16(src) || !aligned16(w) || (out_scale && !aligned16(out_scale)) ||
      (bias && !aligned16(bias)))
    return IDEAS_ERR_UNSUPPORTED;
  if (g.IH > 32767 || g.IW > 32767) return IDEAS_ERR_UNSUPPORTED;
  if (mode == 1 && g_pmh_mode.load() != 0 && !(g.OC % 128 == 0 || g.IC > 128))
    return IDEAS_ERR_UNSUPPORTED;     // narrow outputs: the pixel-major halo kernel serves the phases one by one
  return halo_conv_launch_phases(gs, nph, dst, src, w, out_scale, bias, act, alpha, gain, st);
}

int umma_conv_launch(const ConvGeom& g, float* dst, const float* src, const float* w, const float* out_scale,
                     const float* bias, int act, float alpha, float gain, cudaStream_t st, bool dry_run,
                     const float* residual, float res_scale) {
  // ---- eligibility
  if (g.IC % 32 || g.OC % 32 || g.ntaps < 1) return IDEAS_ERR_UNSUPPORTED;
  if (g.i_s != 1 && g.i_s != 2) return IDEAS_ERR_UNSUPPORTED;
  if ((int64_t)g.N * g.QH * g.QW < 256 || g.QW < 2) return IDEAS_ERR_UNSUPPORTED;   // tiny problems: SIMT
  if (!aligned16(dst) || !aligned16(src) || !aligned16(w) || (out_scale && !aligned16(out_scale)) ||
      (bias && !aligned16(bias)))
    return IDEAS_ERR_UNSUPPORTED;
  if (g.IH > 32767 || g.IW > 32767) return IDEAS_ERR_UNSUPPORTED;
  if (residual && !aligned16(residual)) return IDEAS_ERR_UNSUPPORTED;
  if (dry_run) return IDEAS_OK;
  {
    const int rc = pmh_conv_launch(g, dst, src, w, out_scale, bias, residual, res_scale, act, alpha, gain, st);
    if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
  }
  if (!residual) {      // the channel-major halo kernel has no residual operand; both pixel-major kernels do
    const int rc = halo_conv_launch(g, dst, src, w, out_scale, bias, act, alpha, gain, st);
    if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
  }

  FwdParams p;
  p.dst = dst; p.out_scale = out_scale; p.bias = bias; p.residual = residual; p.res_scale = res_scale;
  p.N = g.N; p.QH = g.QH; p.QW = g.QW; p.OH = g.OH; p.OW = g.OW; p.OC = g.OC;
  p.o_s = g.o_s; p.o_py = g.o_py; p.o_px = g.o_px;
  p.act = act; p.alpha = alpha; p.gain = gain;
  p.ntaps = g.ntaps; p.csteps = g.IC / kBlockK;
  choose_box(g.N, g.QH, g.QW, &p.bw, &p.bh, &p.bn, &p.tiles_x, &p.tiles_y, &p.subtiles);

  // ---- taps -> (map, offset); which parity maps are needed
  bool need[4] = {false, false, false, false};
  int max_widx = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    const ConvTap& tp = g.taps[t];
    int py = 0, px = 0, oy = tp.dy, ox = tp.dx;
    if (g.i_s == 2) {
      py = ((tp.dy % 2) + 2) % 2; px = ((tp.dx % 2) + 2) % 2;
      oy = (tp.dy - py) / 2; ox = (tp.dx - px) / 2;
    }
    p.taps[t].oy = (int16_t)oy; p.taps[t].ox = (int16_t)ox; p.taps[t].widx = (int16_t)tp.widx;
    p.taps[t].map = (int16_t)(py * 2 + px);
    need[py * 2 + px] = true;
    if (tp.widx > max_widx) max_widx = tp.widx;
  }
  const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
  for (int m = 0; m < 4; ++m) {
    if (!need[m]) continue;
    const int py = m >> 1, px = m & 1;
    const int s = g.i_s;
    const int64_t wd = (g.IW - px + s - 1) / s, hd = (g.IH - py + s - 1) / s;
    if (wd < 1 || hd < 1) return IDEAS_ERR_UNSUPPORTED;
    const uint64_t dims[4] = {(uint64_t)g.IC, (uint64_t)wd, (uint64_t)hd, (uint64_t)g.N};
    const uint64_t strides[3] = {(uint64_t)s * g.IC * 4, (uint64_t)s * g.IW * g.IC * 4, (uint64_t)g.IH * g.IW * g.IC * 4};
    const float* base = src + ((int64_t)py * g.IW + px) * g.IC;
    int rc = encode_map(&p.src[m], base, 4, dims, strides, box);
    if (rc) return rc;
  }
  for (int m = 0; m < 4; ++m)
    if (!need[m]) p.src[m] = p.src[need[0] ? 0 : (need[1] ? 1 : (need[2] ? 2 : 3))];

  const int BN = (g.OC % 256 == 0) ? 256 : (g.OC % 128 == 0) ? 128 : (g.OC % 64 == 0) ? 64 : 32;
  // option "pair": 0 = never, 1 = heuristic, 2 = every 256- and 128-channel tile, 3 = 256-channel tiles only.
  // 256-channel tiles: always (1.0-1.65x, scripts/bench_pair.py).  128-channel tiles gain where the overlapped
  // epilogue and the finer granule matter (fused activation or stride-2 source, several rounds of items: 1.16-1.32x)
  // and lose on short launches (384-channel layers at 16x16: 0.7x) and plain data gradients (0.93x).
  const int pair_mode = g_pair.load();
  bool use_pair = false;
  if (p.subtiles >= 4) {
    if (BN == 256) use_pair = pair_mode >= 1;
    if (BN == 128) {
      const int64_t items = (int64_t)ceil_div(p.subtiles, 2) * (g.OC / 128);
      use_pair = pair_mode == 2 ||
                 (pair_mode == 1 && items >= 3 * (kNumSMs / 2) && (g.i_s == 2 || act != IDEAS_ACT_NONE));
    }
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.IC, (uint64_t)g.OC, (uint64_t)(max_widx + 1)};
    const uint64_t strides[2] = {(uint64_t)g.IC * 4, (uint64_t)g.OC * g.IC * 4};
    const uint32_t wbox[3] = {(uint32_t)kBlockK, (uint32_t)(use_pair ? BN / 2 : BN), 1u};
    int rc = encode_map(&p.w, w, 3, dims, strides, wbox);
    if (rc) return rc;
  }
  if (use_pair) return BN == 256 ? launch_fwd_pair<256>(p, g.OC / 256, st) : launch_fwd_pair<128>(p, g.OC / 128, st);
  switch (BN) {
    case 256: return launch_fwd<256>(p, g.OC / 256, st);
    case 128: return launch_fwd<128>(p, g.OC / 128, st);
    case 64: return launch_fwd<64>(p, g.OC / 64, st);
    default: return launch_fwd<32>(p, g.OC / 32, st);
  }
}

// ==========================================================================================
// weight gradient:  dWp[tap][k][c] += sum_pixels dY[pixel, k] * X[pixel*stride + tap, c]
//   GEMM per tap with M = 128 output channels, N = BNC input channels, reduction over pixels.
//   Both operands are "MN-major": the reduction index (pixel) is the row of the NHWC tensors and the
//   GEMM M/N index (channel) is contiguous.  One 5-D TMA box (32 ch x pixels x channel groups) per
//   operand lands [channel group][pixel][32 ch] in shared memory, which is the MN-major canonical
//   layout.  For 32-bit MN-major operands tcgen05 accepts only the "128 B swizzle with 32 B atom"
//   pattern (descriptor layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): LBO = stride between
//   32-channel groups, SBO = stride between 4-pixel groups (512 B); each tcgen05.mma consumes 8
//   pixels (K = 8 for tf32).
//   Tap reuse (stride-1 convs): a CTA owns a GROUP of taps -- all of them when ntaps*BNC <= 512 TMEM
//   columns, else one filter row -- with one accumulator per tap.  The dY tile of a 32-pixel box is
//   loaded once per group, and ONE halo'd X box (bw+halo_x columns, bh+halo_y rows) serves every
//   tap of the group: tap (dy,dx) starts its B descriptor (dy*xbw + dx) rows further (address-based
//   swizzle: scripts/probe_umma_mn_offset.cu).  The 8 pixels of one MMA are 8 consecutive box
//   columns (bw is a multiple of 8), so no reduction index is wasted.  Without it every tap re-read
//   both operands through L2 (9x for a 3x3 filter), which bounded the kernel at 20-40 % of the tensor peak
//   for <= 128 channels.  Stride-2 convs keep one tap per group (per-parity tensor maps).
//   grid = (k-tiles * c-tiles, tap groups, pixel splits); partial sums are merged with vector red.add.
// ==========================================================================================
namespace {

constexpr int kWgPix = 32;                       // pixels per k-block
constexpr uint32_t kWgABytes = 4 * kWgPix * 128; // 128 output channels
constexpr int kWgMaxStages = 8;
constexpr int kWgMaxGroupTaps = 9;

struct TapGroup {
  int16_t oy, ox;        // box origin offset in the addressed x map
  int16_t map;           // which x tensor map (parity)
  int16_t ntaps;
  int16_t drow[kWgMaxGroupTaps];   // row offset of the tap inside the x box
  int16_t widx[kWgMaxGroupTaps];
};

struct alignas(64) WgradParams {
  CUtensorMap dy;        // (32, OW, OH, N, OC/32)
  CUtensorMap x[4];      // (32, Wd, Hd, N, IC/32) per parity
  float* dwp;
  int OC, IC, bnc;
  int bw, bh, bn;        // pixel box of dY, bw*bh*bn == 32
  int tiles_x, tiles_y, nboxes;
  int boxes_per_split;
  int ctiles;
  int stages;
  uint32_t x_group_bytes;   // bytes of one 32-channel group of the x box
  uint32_t stage_bytes, tmem_cols;
  int rowbase[kWgPix / 8];  // x-box row of the first pixel of each 8-pixel MMA step
  TapGroup groups[kMaxTaps];
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int T>   // taps per group (upper bound; groups may hold fewer)
__global__ void __launch_bounds__(kThreads, 1) conv_umma_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int ktile = blockIdx.x / p.ctiles, ctile = blockIdx.x - ktile * p.ctiles;
  const int k0 = ktile * 128, c0 = ctile * p.bnc;
  const TapGroup& tg = p.groups[blockIdx.y];
  const int b_begin = blockIdx.z * p.boxes_per_split;
  const int b_end = min(b_begin + p.boxes_per_split, p.nboxes);
  const int nkb = b_end - b_begin;
  if (nkb <= 0) return;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWgMaxStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&accum_bar), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), p.tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = kWgABytes + (uint32_t)(p.bnc / 32) * p.x_group_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int id = b_begin + kb;
        const int bx = id % p.tiles_x;
        const int t = id / p.tiles_x;
        const int qx0 = bx * p.bw, qy0 = (t % p.tiles_y) * p.bh, n0 = (t / p.tiles_y) * p.bn;
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
        ptx::mbar_arrive_expect_tx(fb, tx_bytes);
        const uint32_t sa = tiles + stage * p.stage_bytes;
        ptx::tma_load_5d(sa, &p.dy, fb, 0, qx0, qy0, n0, k0 / 32);
        ptx::tma_load_5d(sa + kWgABytes, &p.x[tg.map], fb, 0, qx0 + tg.ox, qy0 + tg.oy, n0, c0 / 32);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // The issuing thread is the critical path (an N = 64 MMA retires in ~32 clk): every descriptor is
      // base + a precomputed 32-bit offset held in registers, the tap loop is unrolled at compile time.
      const uint32_t idesc = ptx::idesc_tf32(128, p.bnc, 1, 1);
      uint32_t boff[T][kWgPix / kUmmaK];
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int k = 0; k < kWgPix / kUmmaK; ++k)
          boff[t][k] = (uint32_t)((p.rowbase[k] + (t < tg.ntaps ? tg.drow[t] : 0)) * 128) >> 4;
      const int ntaps = tg.ntaps;
      const uint32_t bnc = p.bnc;
      const uint64_t adesc0 = ptx::smem_desc_sw128_base32(tiles, kWgPix * 128, 512);
      const uint64_t bdesc0 = ptx::smem_desc_sw128_base32(tiles + kWgABytes, p.x_group_bytes, 512);
      const uint32_t stage_enc = p.stage_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t senc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
        ptx::tc_fence_after();
        const uint64_t ad = adesc0 + senc, bd = bdesc0 + senc;
        const uint32_t acc_flag = kb != 0;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          if (t < ntaps) {
#pragma unroll
            for (int k = 0; k < kWgPix / kUmmaK; ++k)
              ptx::mma_tf32(tmem_base + t * bnc, ad + (uint32_t)(k * 64), bd + boff[t][k], idesc, k ? 1u : acc_flag);
          }
        }
        ptx::mma_commit(ptx::smem_u32(&empty_bar[stage]));
        senc += stage_enc;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; senc = 0; }
      }
      ptx::mma_commit(ptx::smem_u32(&accum_bar));
    }
  } else {
    const int quarter = warp & 3;
    const int k = k0 + quarter * 32 + lane;
    ptx::mbar_wait(ptx::smem_u32(&accum_bar), 0);
    ptx::tc_fence_after();
    for (int t = 0; t < tg.ntaps; ++t) {
      float* dp = p.dwp + ((int64_t)tg.widx[t] * p.OC + k) * p.IC + c0;
      for (int ch = 0; ch < p.bnc / 32; ++ch) {
        float v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * p.bnc + ch * 32), v);
        if (k < p.OC && c0 + ch * 32 < p.IC) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) red_add_v4(dp + ch * 32 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// 32-pixel box (bw x bh x bn) covering the q grid with the fewest boxes weighted by the x halo overhead;
// with tap reuse bw must be a multiple of 8 (one MMA = 8 consecutive box columns).
void choose_box32(int N, int QH, int QW, int hx, int hy, bool reuse, int* bw, int* bh, int* bn, int* tx, int* ty, int* nboxes) {
  double best = -1;
  for (int w = 32; w >= (reuse ? 8 : 1); w >>= 1) {
    if (w > pow2_ceil(QW) && w > (reuse ? 8 : 1)) continue;
    for (int h = 32 / w; h >= 1; h >>= 1) {
      if (h > pow2_ceil(QH) && h > 1) continue;
      const int n = 32 / (w * h);
      const long cnt = (long)ceil_div(QW, w) * ceil_div(QH, h) * ceil_div(N, n);
      // cost per box: the dY tile (128 channels x 32 pixels) plus the halo'd x tile
      const double cost = (double)cnt * (32.0 + (double)(w + hx) * (h + hy) * n);
      if (best < 0 || cost < best) {
        best = cost; *bw = w; *bh = h; *bn = n; *tx = ceil_div(QW, w); *ty = ceil_div(QH, h);
        *nboxes = (int)cnt;
      }
    }
  }
}

std::atomic<int> g_wgrad_reuse{1};

template <int T>
int launch_wgrad(const WgradParams& p, dim3 grid, uint32_t smem, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_umma_wgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "conv_umma_wgrad: cudaFuncSetAttribute");
  conv_umma_wgrad_kernel<T><<<grid, kThreads, smem, st>>>(p);
  IDEAS_CHECK_LAUNCH("conv_umma_wgrad");
  return IDEAS_OK;
}

}  // namespace

// g is the FORWARD geometry (src = x, dst = dy); dwp is zero-initialised / accumulated into.
int umma_wgrad_launch(const ConvGeom& g, float* dwp, const float* x, const float* dy, cudaStream_t st, bool dry_run) {
  if (g.IC % 32 || g.OC % 32 || g.ntaps < 1 || g.o_s != 1) return IDEAS_ERR_UNSUPPORTED;
  if (g.i_s != 1 && g.i_s != 2) return IDEAS_ERR_UNSUPPORTED;
  if ((int64_t)g.N * g.QH * g.QW < 512) return IDEAS_ERR_UNSUPPORTED;
  if (!aligned16(dwp) || !aligned16(x) || !aligned16(dy)) return IDEAS_ERR_UNSUPPORTED;
  if (g.IH > 32767 || g.IW > 32767) return IDEAS_ERR_UNSUPPORTED;
  if (dry_run) return IDEAS_OK;

  WgradParams p;
  p.dwp = dwp; p.OC = g.OC; p.IC = g.IC;
  // ---- tap groups
  int min_dx = 1 << 30, max_dx = -(1 << 30), min_dy = 1 << 30, max_dy = -(1 << 30);
  for (int t = 0; t < g.ntaps; ++t) {
    min_dx = g.taps[t].dx < min_dx ? g.taps[t].dx : min_dx; max_dx = g.taps[t].dx > max_dx ? g.taps[t].dx : max_dx;
    min_dy = g.taps[t].dy < min_dy ? g.taps[t].dy : min_dy; max_dy = g.taps[t].dy > max_dy ? g.taps[t].dy : max_dy;
  }
  // tap reuse pays where activation/gradient re-reads bound the kernel (<= 128 input channels); wider layers
  // keep one tap per CTA with a 256-channel N tile (measured: scripts/kernels_microbench.py)
  const int reuse_mode = g_wgrad_reuse.load();
  const bool reuse = reuse_mode && g.i_s == 1 && g.ntaps > 1 && g.QW >= 8 && (max_dx - min_dx) <= 4 &&
                     (max_dy - min_dy) <= 4 && (g.IC <= 128 || reuse_mode == 2);
  int BNC;
  if (reuse) BNC = g.IC >= 128 && g.IC % 128 == 0 ? 128 : (g.IC % 64 == 0 ? 64 : 32);
  else BNC = g.IC >= 256 && g.IC % 256 == 0 ? 256 : (g.IC > 64 ? 128 : (g.IC > 32 ? 64 : 32));
  p.bnc = BNC;
  p.ctiles = ceil_div(g.IC, BNC);
  const int ktiles = ceil_div(g.OC, 128);
  int ngroups = 0, hx = 0, hy = 0;
  bool need[4] = {false, false, false, false};
  if (reuse) {
    const bool all = g.ntaps * BNC <= 512 && g.ntaps <= kWgMaxGroupTaps;
    hx = max_dx - min_dx;
    hy = all ? max_dy - min_dy : 0;
    for (int t = 0; t < g.ntaps; ++t) {
      const ConvTap& tp = g.taps[t];
      int gi = -1;
      if (all) gi = ngroups ? 0 : -1;
      else
        for (int q = 0; q < ngroups; ++q)
          if (p.groups[q].oy == tp.dy) gi = q;      // one group per filter row
      if (gi < 0) {
        gi = ngroups++;
        p.groups[gi].oy = (int16_t)(all ? min_dy : tp.dy);
        p.groups[gi].ox = (int16_t)min_dx;
        p.groups[gi].map = 0;
        p.groups[gi].ntaps = 0;
      }
      TapGroup& G = p.groups[gi];
      if (G.ntaps >= kWgMaxGroupTaps || (G.ntaps + 1) * BNC > 512) return IDEAS_ERR_UNSUPPORTED;
      G.drow[G.ntaps] = 0;   // filled below once the box is known
      G.widx[G.ntaps] = (int16_t)t;   // index into g.taps, resolved below
      ++G.ntaps;
    }
    need[0] = true;
  } else {
    for (int t = 0; t < g.ntaps; ++t) {
      const ConvTap& tp = g.taps[t];
      int py = 0, px = 0, oy = tp.dy, ox = tp.dx;
      if (g.i_s == 2) {
        py = ((tp.dy % 2) + 2) % 2; px = ((tp.dx % 2) + 2) % 2;
        oy = (tp.dy - py) / 2; ox = (tp.dx - px) / 2;
      }
      TapGroup& G = p.groups[ngroups++];
      G.oy = (int16_t)oy; G.ox = (int16_t)ox; G.map = (int16_t)(py * 2 + px); G.ntaps = 1;
      G.drow[0] = 0; G.widx[0] = (int16_t)tp.widx;
      need[py * 2 + px] = true;
    }
  }
  choose_box32(g.N, g.QH, g.QW, hx, hy, reuse, &p.bw, &p.bh, &p.bn, &p.tiles_x, &p.tiles_y, &p.nboxes);
  const int xbw = p.bw + hx, xbh = p.bh + hy;
  if (reuse) {
    for (int q = 0; q < ngroups; ++q) {
      TapGroup& G = p.groups[q];
      for (int j = 0; j < G.ntaps; ++j) {
        const ConvTap& tp = g.taps[G.widx[j]];
        G.drow[j] = (int16_t)((tp.dy - G.oy) * xbw + (tp.dx - G.ox));
        G.widx[j] = (int16_t)tp.widx;
      }
    }
  }
  for (int k = 0; k < kWgPix / 8; ++k) {
    const int q = k * 8;
    const int xq = q % p.bw, yq = (q / p.bw) % p.bh, nq = q / (p.bw * p.bh);
    p.rowbase[k] = (nq * xbh + yq) * xbw + xq;
  }
  p.x_group_bytes = (uint32_t)(p.bn * xbh * xbw) * 128u;
  const uint32_t xbytes = (uint32_t)(BNC / 32) * p.x_group_bytes;
  p.stage_bytes = (kWgABytes + xbytes + 1023u) & ~1023u;
  p.stages = (int)(200 * 1024 / p.stage_bytes);
  if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  if (p.stages < 2) return IDEAS_ERR_UNSUPPORTED;
  int max_group = 1;
  for (int q = 0; q < ngroups; ++q) max_group = p.groups[q].ntaps > max_group ? p.groups[q].ntaps : max_group;
  uint32_t cols = 32;
  while (cols < (uint32_t)(max_group * BNC)) cols <<= 1;
  p.tmem_cols = cols;
  {
    const uint64_t dims[5] = {32, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N, (uint64_t)(g.OC / 32)};
    const uint64_t strides[4] = {(uint64_t)g.OC * 4, (uint64_t)g.OW * g.OC * 4, (uint64_t)g.OH * g.OW * g.OC * 4, 128};
    const uint32_t box[5] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn, 4};
    int rc = encode_map(&p.dy, dy, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  for (int m = 0; m < 4; ++m) {
    if (!need[m]) continue;
    const int py = m >> 1, px = m & 1, s = g.i_s;
    const int64_t wd = (g.IW - px + s - 1) / s, hd = (g.IH - py + s - 1) / s;
    if (wd < 1 || hd < 1) return IDEAS_ERR_UNSUPPORTED;
    const uint64_t dims[5] = {32, (uint64_t)wd, (uint64_t)hd, (uint64_t)g.N, (uint64_t)(g.IC / 32)};
    const uint64_t strides[4] = {(uint64_t)s * g.IC * 4, (uint64_t)s * g.IW * g.IC * 4, (uint64_t)g.IH * g.IW * g.IC * 4, 128};
    const uint32_t box[5] = {32, (uint32_t)xbw, (uint32_t)xbh, (uint32_t)p.bn, (uint32_t)(BNC / 32)};
    int rc = encode_map(&p.x[m], x + ((int64_t)py * g.IW + px) * g.IC, 5, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  for (int m = 0; m < 4; ++m)
    if (!need[m]) p.x[m] = p.x[need[0] ? 0 : (need[1] ? 1 : (need[2] ? 2 : 3))];

  // one CTA per SM (200 KB of shared memory each): size the pixel splits so the grid is ONE wave that fills
  // the 148 SMs without spilling into a second, nearly empty one (ncu on the 128->128 @256^2 shape: a grid of
  // 297 CTAs left the SMs idle a third of the time)
  const int base = ktiles * p.ctiles * ngroups;
  int splits = kNumSMs / base;
  if (splits > p.nboxes) splits = p.nboxes;
  if (splits < 1) splits = 1;
  p.boxes_per_split = ceil_div(p.nboxes, splits);
  splits = ceil_div(p.nboxes, p.boxes_per_split);
  dim3 grid(ktiles * p.ctiles, ngroups, splits);
  const uint32_t smem = p.stages * p.stage_bytes + 1024;
  switch (max_group) {
    case 1: return launch_wgrad<1>(p, grid, smem, st);
    case 2: return launch_wgrad<2>(p, grid, smem, st);
    case 3: return launch_wgrad<3>(p, grid, smem, st);
    case 4: return launch_wgrad<4>(p, grid, smem, st);
    case 9: return launch_wgrad<9>(p, grid, smem, st);
    default: break;
  }
  set_error("conv_umma_wgrad: unsupported tap group size %d", max_group);
  return IDEAS_ERR_UNSUPPORTED;
}

}  // namespace ideas

extern "C" int ideas_umma_available(void) { return 1; }

extern "C" int ideas_set_option(const char* name, int value) {
  if (name && !strcmp(name, "tma_tf32")) {
    ideas::g_tma_tf32.store(value ? 1 : 0);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "blur_variant")) {
    ideas::set_blur_variant(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "wgrad_reuse")) {
    ideas::g_wgrad_reuse.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "halo")) {
    ideas::g_halo_mode.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "dgrad_phases")) {
    if (value < 0 || value > 2) { ideas::set_error("ideas_set_option: dgrad_phases must be 0, 1 or 2"); return IDEAS_ERR_INVALID; }
    ideas::g_dgrad_phases.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "gemm_split")) {
    ideas::set_gemm_split(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "resample_variant")) {
    ideas::set_resample_variant(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pair_epi")) {
    ideas::g_pair_epi.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pair_stages")) {
    ideas::g_pair_stages.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pair")) {
    ideas::g_pair.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "halo_epi")) {
    ideas::g_halo_epi.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "tail_split")) {
    ideas::g_tail_split.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pmh")) {
    ideas::g_pmh_mode.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pmh_resident")) {
    ideas::g_pmh_resident.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pmh_xstages")) {
    ideas::g_pmh_xstages.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pmh_align")) {
    ideas::g_pmh_align.store(value);
    return IDEAS_OK;
  }
  if (name && !strcmp(name, "pmh_minbw")) {
    ideas::g_pmh_minbw.store(value);
    return IDEAS_OK;
  }
  ideas::set_error("ideas_set_option: unknown option '%s'", name ? name : "(null)");
  return IDEAS_ERR_INVALID;
}
