// A1/A5: implicit-GEMM convolution, fp32 FFMA path ("SIMT"), plus the C entry points of the
// convolution family.  The reference has no native code here: it calls cuDNN through
// F.conv2d / F.conv_transpose2d (stylegan2/model.py:115,258,267,273; models.py:32) on
// per-sample materialised weights.
//
// This file is (a) the exact-fp32 differential reference for the tcgen05 kernels in
// conv_umma.cu, and (b) the production path for the shapes tensor cores cannot take
// (C or K not a multiple of 32: the 3-, 1- and 8-channel ends of the networks).
// Tiling: 64 pixels x 64 output channels per 256-thread CTA, 16 reduction channels per
// step, register-prefetched double-buffered shared memory, 4x4 outputs per thread.
// The style modulation s[n,c] is applied while staging the source tile, the demodulation
// d[n,k], bias and leaky ReLU in the epilogue -- per-sample weights are never built.
#include "common.cuh"
#include "conv_geom.h"

namespace ideas {

// implemented in conv_umma.cu; return IDEAS_ERR_UNSUPPORTED when the shape does not qualify
int umma_dgrad_phases_launch(const ConvGeom* gs, int nph, float* dst, const float* src, const float* w,
                             const float* out_scale, const float* bias, int act, float alpha, float gain,
                             cudaStream_t st);
int umma_conv_launch(const ConvGeom& g, float* dst, const float* src, const float* w, const float* out_scale,
                     const float* bias, int act, float alpha, float gain, cudaStream_t st, bool dry_run,
                     const float* residual = nullptr, float res_scale = 1.f);
int umma_wgrad_launch(const ConvGeom& g, float* dwp, const float* src, const float* dy, cudaStream_t st, bool dry_run);
// implemented in conv_pointwise.cu: 1x1 convs with <= 4 channels on one side (HBM-bound streaming kernels)
int pointwise_conv_launch(const ConvGeom& g, float* dst, const float* src, const float* w, const float* bias, int act,
                          float alpha, float gain, cudaStream_t st);
int pointwise_wgrad_launch(const ConvGeom& g, float* dwp, const float* x, const float* dy, cudaStream_t st);

struct ConvArgs {
  ConvGeom g;
  const float* src;
  const float* w;
  float* dst;
  const float* in_scale;
  const float* out_scale;
  const float* bias;
  const float* residual;   // dst-shaped, merged after the activation: (act(..) + residual) * res_scale
  float res_scale;
  int act;
  float alpha, gain;
  int64_t M;
};

constexpr int BM = 64, BN = 64, BK = 16, LD = 68;

template <bool VEC>
__global__ void __launch_bounds__(256) conv_igemm_simt_kernel(const __grid_constant__ ConvArgs a) {
  __shared__ __align__(16) float As[2][BK][LD];
  __shared__ __align__(16) float Bs[2][BK][LD];
  const ConvGeom& g = a.g;
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int k0 = blockIdx.y * BN;

  // ---- loader role: one source pixel row / one weight row per thread, 4 channels each
  const int lrow = tid >> 2;
  const int lq = (tid & 3) * 4;
  const int64_t m = m0 + lrow;
  const bool mvalid = m < a.M;
  int n = 0, qy = 0, qx = 0;
  if (mvalid) {
    qx = (int)(m % g.QW);
    int64_t t = m / g.QW;
    qy = (int)(t % g.QH);
    n = (int)(t / g.QH);
  }
  const float* srcn = a.src + (int64_t)n * g.IH * g.IW * g.IC;
  const float* isc = a.in_scale ? a.in_scale + (int64_t)n * g.IC : nullptr;
  const int kb = k0 + lrow;
  const bool kvalid = kb < g.OC;
  const int csteps = (g.IC + BK - 1) / BK;
  const int total = g.ntaps * csteps;

  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
  auto load = [&](int step) {
    const int t = step / csteps;
    const int c0 = (step - t * csteps) * BK + lq;
    const ConvTap tp = g.taps[t];
    const int iy = qy * g.i_s + tp.dy, ix = qx * g.i_s + tp.dx;
    const bool ok = mvalid && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    if (VEC) {
      if (c0 < g.IC) {
        if (ok) {
          ra = __ldg(reinterpret_cast<const float4*>(srcn + ((int64_t)iy * g.IW + ix) * g.IC + c0));
          if (isc) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(isc + c0));
            ra.x *= s.x; ra.y *= s.y; ra.z *= s.z; ra.w *= s.w;
          }
        }
        if (kvalid) rb = __ldg(reinterpret_cast<const float4*>(a.w + ((int64_t)tp.widx * g.OC + kb) * g.IC + c0));
      }
    } else {
      float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        if (c < g.IC) {
          if (ok) {
            float v = __ldg(srcn + ((int64_t)iy * g.IW + ix) * g.IC + c);
            if (isc) v *= __ldg(isc + c);
            va[j] = v;
          }
          if (kvalid) vb[j] = __ldg(a.w + ((int64_t)tp.widx * g.OC + kb) * g.IC + c);
        }
      }
      ra = make_float4(va[0], va[1], va[2], va[3]);
      rb = make_float4(vb[0], vb[1], vb[2], vb[3]);
    }
  };
  auto stage = [&](int buf) {
    As[buf][lq + 0][lrow] = ra.x; As[buf][lq + 1][lrow] = ra.y; As[buf][lq + 2][lrow] = ra.z; As[buf][lq + 3][lrow] = ra.w;
    Bs[buf][lq + 0][lrow] = rb.x; Bs[buf][lq + 1][lrow] = rb.y; Bs[buf][lq + 2][lrow] = rb.z; Bs[buf][lq + 3][lrow] = rb.w;
  };

  // ---- compute role: 4 pixels x 4 channels per thread
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (total > 0) {
    load(0);
    stage(0);
    __syncthreads();
    for (int step = 0; step < total; ++step) {
      const int cur = step & 1;
      if (step + 1 < total) load(step + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w};
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
      if (step + 1 < total) stage(cur ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: demodulation, bias, activation, strided (phase) store
  const int kc = k0 + tx * 4;
  const bool vec_store = (g.OC % 4 == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t mm = m0 + ty * 4 + i;
    if (mm >= a.M) continue;
    const int ox = (int)(mm % g.QW);
    const int64_t t = mm / g.QW;
    const int oy = (int)(t % g.QH);
    const int nn = (int)(t / g.QH);
    const int64_t poff = (((int64_t)nn * g.OH + (oy * g.o_s + g.o_py)) * g.OW + (ox * g.o_s + g.o_px)) * g.OC;
    float* dp = a.dst + poff;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc + j;
      float r = acc[i][j];
      if (k < g.OC) {
        if (a.out_scale) r *= __ldg(a.out_scale + (int64_t)nn * g.OC + k);
        if (a.bias) r += __ldg(a.bias + k);
        if (a.act == IDEAS_ACT_LRELU) r = lrelu(r, a.alpha) * a.gain;
        if (a.residual) r = (r + __ldg(a.residual + poff + k)) * a.res_scale;
      }
      v[j] = r;
    }
    if (vec_store && kc + 3 < g.OC) {
      *reinterpret_cast<float4*>(dp + kc) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (kc + j < g.OC) dp[kc + j] = v[j];
    }
  }
}

// ---------------------------------------------------------------------------------------
// weight gradient: dwp[widx][k][c] += sum_pixels (out_scale*dy)[pix_o, k] * (in_scale*x)[pix_i, c]
// 64 (k) x 64 (c) tile per CTA for one tap and one slice of the pixel range.
// ---------------------------------------------------------------------------------------
struct WgradArgs {
  ConvGeom g;          // forward geometry: src = x, dst = dy
  const float* x;
  const float* dy;
  float* dwp;
  const float* in_scale;
  const float* out_scale;
  int64_t M;           // N*QH*QW
  int64_t per_split;   // pixels per blockIdx.z, multiple of 16
  int ctiles;
};

template <bool VEC>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const __grid_constant__ WgradArgs a) {
  __shared__ __align__(16) float Ds[2][16][64];
  __shared__ __align__(16) float Xs[2][16][64];
  const ConvGeom& g = a.g;
  const int tid = threadIdx.x;
  const int ktile = blockIdx.x / a.ctiles, ctile = blockIdx.x - ktile * a.ctiles;
  const int k0 = ktile * 64, c0 = ctile * 64;
  const ConvTap tp = g.taps[blockIdx.y];
  const int64_t p_begin = (int64_t)blockIdx.z * a.per_split;
  const int64_t p_end = min(p_begin + a.per_split, a.M);
  if (p_begin >= p_end) return;
  const int steps = (int)((p_end - p_begin + 15) / 16);

  const int lp = tid >> 4;          // pixel within the 16-pixel step
  const int lq = (tid & 15) * 4;    // channel quad
  float4 rd, rx;
  auto load = [&](int step) {
    rd = make_float4(0.f, 0.f, 0.f, 0.f);
    rx = rd;
    const int64_t m = p_begin + (int64_t)step * 16 + lp;
    if (m >= p_end) return;
    const int qx = (int)(m % g.QW);
    const int64_t t = m / g.QW;
    const int qy = (int)(t % g.QH);
    const int n = (int)(t / g.QH);
    const int oy = qy * g.o_s + g.o_py, ox = qx * g.o_s + g.o_px;
    const int iy = qy * g.i_s + tp.dy, ix = qx * g.i_s + tp.dx;
    const float* dp = a.dy + (((int64_t)n * g.OH + oy) * g.OW + ox) * g.OC;
    const bool ok = iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
    const float* xp = a.x + (((int64_t)n * g.IH + iy) * g.IW + ix) * g.IC;
    float vd[4] = {0.f, 0.f, 0.f, 0.f}, vx[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {
      if (k0 + lq < g.OC) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dp + k0 + lq));
        vd[0] = v.x; vd[1] = v.y; vd[2] = v.z; vd[3] = v.w;
      }
      if (ok && c0 + lq < g.IC) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xp + c0 + lq));
        vx[0] = v.x; vx[1] = v.y; vx[2] = v.z; vx[3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + lq + j < g.OC) vd[j] = __ldg(dp + k0 + lq + j);
        if (ok && c0 + lq + j < g.IC) vx[j] = __ldg(xp + c0 + lq + j);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (a.out_scale && k0 + lq + j < g.OC) vd[j] *= __ldg(a.out_scale + (int64_t)n * g.OC + k0 + lq + j);
      if (a.in_scale && c0 + lq + j < g.IC) vx[j] *= __ldg(a.in_scale + (int64_t)n * g.IC + c0 + lq + j);
    }
    rd = make_float4(vd[0], vd[1], vd[2], vd[3]);
    rx = make_float4(vx[0], vx[1], vx[2], vx[3]);
  };
  auto stage = [&](int buf) {
    *reinterpret_cast<float4*>(&Ds[buf][lp][lq]) = rd;
    *reinterpret_cast<float4*>(&Xs[buf][lp][lq]) = rx;
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load(0);
  stage(0);
  __syncthreads();
  for (int step = 0; step < steps; ++step) {
    const int cur = step & 1;
    if (step + 1 < steps) load(step + 1);
#pragma unroll
    for (int pp = 0; pp < 16; ++pp) {
      const float4 av = *reinterpret_cast<const float4*>(&Ds[cur][pp][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Xs[cur][pp][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    if (step + 1 < steps) stage(cur ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= g.OC) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < g.IC) atomicAdd(a.dwp + ((int64_t)tp.widx * g.OC + k) * g.IC + c, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// weight (un)packing
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weight_kernel(float* __restrict__ dst, const float* __restrict__ src, int O,
                                                          int I, int taps, int transpose, int flip, float scale) {
  const int64_t total = (int64_t)O * I * taps;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates dst: [tap'][A][B] with (A,B) = (o,i) or (i,o)
    const int A = transpose ? I : O, B = transpose ? O : I;
    const int b = (int)(idx % B);
    const int64_t t = idx / B;
    const int aidx = (int)(t % A);
    const int tapd = (int)(t / A);
    const int o = transpose ? b : aidx, i = transpose ? aidx : b;
    const int tap = flip ? (taps - 1 - tapd) : tapd;
    dst[idx] = scale * __ldg(src + ((int64_t)o * I + i) * taps + tap);
  }
}

__global__ void __launch_bounds__(256) unpack_wgrad_kernel(float* __restrict__ dst, const float* __restrict__ src, int O,
                                                           int I, int taps, int transpose, float scale,
                                                           int accumulate) {
  const int64_t total = (int64_t)O * I * taps;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates dst OIHW
    const int tap = (int)(idx % taps);
    const int64_t t = idx / taps;
    const int i = (int)(t % I);
    const int o = (int)(t / I);
    const float v = scale * (transpose ? __ldg(src + ((int64_t)tap * I + i) * O + o)
                                       : __ldg(src + ((int64_t)tap * O + o) * I + i));
    dst[idx] = accumulate ? dst[idx] + v : v;
  }
}

// wpt[taps-1-t][c][k] = wp[t][k][c]: the data-gradient operand derived from the packed forward weight
__global__ void __launch_bounds__(256) repack_dgrad_kernel(float* __restrict__ dst, const float* __restrict__ src, int K,
                                                           int C, int taps) {
  const int64_t total = (int64_t)taps * K * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const int64_t t = idx / K;
    const int c = (int)(t % C);
    const int td = (int)(t / C);
    dst[idx] = __ldg(src + ((int64_t)(taps - 1 - td) * K + k) * C + c);
  }
}

static int launch_simt(const ConvGeom& g, float* dst, const float* src, const float* w, const float* in_scale,
                       const float* out_scale, const float* bias, int act, float alpha, float gain, cudaStream_t st,
                       const float* residual = nullptr, float res_scale = 1.f) {
  ConvArgs a;
  a.g = g; a.src = src; a.w = w; a.dst = dst; a.in_scale = in_scale; a.out_scale = out_scale; a.bias = bias;
  a.residual = residual; a.res_scale = res_scale;
  a.act = act; a.alpha = alpha; a.gain = gain;
  a.M = (int64_t)g.N * g.QH * g.QW;
  if (a.M == 0 || g.OC == 0) return IDEAS_OK;
  const int64_t mt = ceil_div64(a.M, BM);
  IDEAS_REQUIRE(mt <= 0x7fffffff, "conv: too many pixel tiles");
  dim3 grid((unsigned)mt, ceil_div(g.OC, BN));
  const bool vec = g.IC % 4 == 0 && aligned16(src) && aligned16(w) && (!in_scale || aligned16(in_scale)) &&
                   aligned16(dst);
  if (vec) conv_igemm_simt_kernel<true><<<grid, 256, 0, st>>>(a);
  else conv_igemm_simt_kernel<false><<<grid, 256, 0, st>>>(a);
  IDEAS_CHECK_LAUNCH("conv_igemm_simt");
  return IDEAS_OK;
}

static bool umma_wanted(int impl) { return impl == IDEAS_IMPL_AUTO || impl == IDEAS_IMPL_UMMA; }

}  // namespace ideas

using namespace ideas;

static int check_conv_args(const char* who, int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad) {
  IDEAS_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 1 && K >= 1, "%s: bad tensor shape N=%d H=%d W=%d C=%d K=%d", who, N,
                H, W, C, K);
  IDEAS_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= kMaxTaps, "%s: kernel %dx%d outside the supported 1..16 taps", who, kh,
                kw);
  IDEAS_REQUIRE(stride >= 1 && stride <= 4 && pad >= 0, "%s: bad stride/pad", who);
  return IDEAS_OK;
}

extern "C" int ideas_pack_weight(float* dst, const float* src, int O, int I, int kh, int kw, int transpose, int flip,
                                 float scale, void* stream) {
  IDEAS_REQUIRE(dst && src && O >= 1 && I >= 1 && kh >= 1 && kw >= 1, "pack_weight: bad arguments");
  const int64_t total = (int64_t)O * I * kh * kw;
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_weight_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, O, I, kh * kw, transpose, flip, scale);
  IDEAS_CHECK_LAUNCH("pack_weight");
  return IDEAS_OK;
}

extern "C" int ideas_repack_dgrad(float* dst, const float* wp, int K, int C, int taps, void* stream) {
  IDEAS_REQUIRE(dst && wp && K >= 1 && C >= 1 && taps >= 1, "repack_dgrad: bad arguments");
  const int64_t total = (int64_t)taps * K * C;
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  repack_dgrad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dst, wp, K, C, taps);
  IDEAS_CHECK_LAUNCH("repack_dgrad");
  return IDEAS_OK;
}

extern "C" int ideas_unpack_weight_grad(float* dst, const float* src, int O, int I, int kh, int kw, int transpose,
                                        float scale, int accumulate, void* stream) {
  IDEAS_REQUIRE(dst && src && O >= 1 && I >= 1 && kh >= 1 && kw >= 1, "unpack_weight_grad: bad arguments");
  const int64_t total = (int64_t)O * I * kh * kw;
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  unpack_wgrad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, O, I, kh * kw, transpose, scale, accumulate);
  IDEAS_CHECK_LAUNCH("unpack_weight_grad");
  return IDEAS_OK;
}

extern "C" int ideas_conv2d_forward(float* y, const float* x, const float* wp, const float* in_scale,
                                    const float* out_scale, const float* bias, int N, int H, int W, int C, int K,
                                    int kh, int kw, int stride, int pad, int act, float alpha, float gain, int impl,
                                    void* stream) {
  return ideas_conv2d_forward_res(y, x, wp, in_scale, out_scale, bias, nullptr, 1.f, N, H, W, C, K, kh, kw, stride, pad,
                                  act, alpha, gain, impl, stream);
}

extern "C" int ideas_conv2d_forward_res(float* y, const float* x, const float* wp, const float* in_scale,
                                        const float* out_scale, const float* bias, const float* residual,
                                        float res_scale, int N, int H, int W, int C, int K, int kh, int kw, int stride,
                                        int pad, int act, float alpha, float gain, int impl, void* stream) {
  int rc = check_conv_args("conv2d_forward", N, H, W, C, K, kh, kw, stride, pad);
  if (rc) return rc;
  IDEAS_REQUIRE(H + 2 * pad >= kh && W + 2 * pad >= kw, "conv2d_forward: kernel larger than padded input");
  const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
  if (N == 0) return IDEAS_OK;
  IDEAS_REQUIRE(y && x && wp, "conv2d_forward: null pointer");
  const ConvGeom g = geom_forward(N, H, W, C, K, kh, kw, stride, pad, OH, OW);
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == IDEAS_IMPL_AUTO && !in_scale && !out_scale && !residual) {
    rc = pointwise_conv_launch(g, y, x, wp, bias, act, alpha, gain, st);
    if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
  }
  if (umma_wanted(impl) && !in_scale) {
    rc = umma_conv_launch(g, y, x, wp, out_scale, bias, act, alpha, gain, st, false, residual, res_scale);
    if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
  }
  if (impl == IDEAS_IMPL_UMMA) {
    set_error("conv2d_forward: shape not eligible for the tcgen05 path (or in_scale given)");
    return IDEAS_ERR_UNSUPPORTED;
  }
  return launch_simt(g, y, x, wp, in_scale, out_scale, bias, act, alpha, gain, st, residual, res_scale);
}

extern "C" int ideas_conv2d_dgrad(float* dx, const float* dy, const float* wpt, const float* in_scale,
                                  const float* out_scale, const float* bias, int N, int H, int W, int C, int K, int kh,
                                  int kw, int stride, int pad, int OH, int OW, int act, float alpha, float gain,
                                  int impl, void* stream) {
  int rc = check_conv_args("conv2d_dgrad", N, H, W, C, K, kh, kw, stride, pad);
  if (rc) return rc;
  IDEAS_REQUIRE(OH >= 1 && OW >= 1 && (OH - 1) * stride + kh - 2 * pad <= H && (OW - 1) * stride + kw - 2 * pad <= W,
                "conv2d_dgrad: dy %dx%d does not fit dx %dx%d", OH, OW, H, W);
  if (N == 0) return IDEAS_OK;
  IDEAS_REQUIRE(dx && dy && wpt, "conv2d_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (stride == 2 && umma_wanted(impl) && !out_scale && H >= 2 && W >= 2) {
    ConvGeom gs[4];
    int nph = 0;
    bool ok = true;
    for (int py = 0; py < 2 && ok; ++py)
      for (int px = 0; px < 2 && ok; ++px) {
        gs[nph] = geom_dgrad_phase(N, H, W, C, K, kh, kw, stride, pad, OH, OW, py, px);
        ok = gs[nph].ntaps >= 1;      // a phase without taps is bias/zero only: leave it to the per-phase path
        ++nph;
      }
    if (ok) {
      rc = umma_dgrad_phases_launch(gs, nph, dx, dy, wpt, in_scale, bias, act, alpha, gain, st);
      if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
    }
  }
  for (int py = 0; py < stride; ++py)
    for (int px = 0; px < stride; ++px) {
      if (py >= H || px >= W) continue;
      const ConvGeom g = geom_dgrad_phase(N, H, W, C, K, kh, kw, stride, pad, OH, OW, py, px);
      rc = IDEAS_ERR_UNSUPPORTED;
      if (impl == IDEAS_IMPL_AUTO && !in_scale && !out_scale)
        rc = pointwise_conv_launch(g, dx, dy, wpt, bias, act, alpha, gain, st);
      // in this geometry the GEMM reduces over K: `out_scale` (per dy channel) plays the input role
      if (rc == IDEAS_ERR_UNSUPPORTED && umma_wanted(impl) && !out_scale)
        rc = umma_conv_launch(g, dx, dy, wpt, in_scale, bias, act, alpha, gain, st, false);
      if (rc == IDEAS_ERR_UNSUPPORTED) {
        if (impl == IDEAS_IMPL_UMMA) {
          set_error("conv2d_dgrad: shape not eligible for the tcgen05 path (or out_scale given)");
          return rc;
        }
        rc = launch_simt(g, dx, dy, wpt, out_scale, in_scale, bias, act, alpha, gain, st);
      }
      if (rc) return rc;
    }
  return IDEAS_OK;
}

extern "C" int ideas_conv2d_wgrad(float* dwp, const float* x, const float* dy, const float* in_scale,
                                  const float* out_scale, int N, int H, int W, int C, int K, int kh, int kw, int stride,
                                  int pad, int OH, int OW, int impl, void* stream) {
  int rc = check_conv_args("conv2d_wgrad", N, H, W, C, K, kh, kw, stride, pad);
  if (rc) return rc;
  IDEAS_REQUIRE(OH >= 1 && OW >= 1, "conv2d_wgrad: empty dy");
  if (N == 0) return IDEAS_OK;
  IDEAS_REQUIRE(dwp && x && dy, "conv2d_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const ConvGeom g = geom_forward(N, H, W, C, K, kh, kw, stride, pad, OH, OW);
  if (impl == IDEAS_IMPL_AUTO && !in_scale && !out_scale) {
    rc = pointwise_wgrad_launch(g, dwp, x, dy, st);
    if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
  }
  if (umma_wanted(impl) && !in_scale && !out_scale) {
    rc = umma_wgrad_launch(g, dwp, x, dy, st, false);
    if (rc != IDEAS_ERR_UNSUPPORTED) return rc;
  }
  if (impl == IDEAS_IMPL_UMMA) {
    set_error("conv2d_wgrad: shape not eligible for the tcgen05 path (or scales given)");
    return IDEAS_ERR_UNSUPPORTED;
  }
  WgradArgs a;
  a.g = g; a.x = x; a.dy = dy; a.dwp = dwp; a.in_scale = in_scale; a.out_scale = out_scale;
  a.M = (int64_t)N * g.QH * g.QW;
  a.ctiles = ceil_div(C, 64);
  const int tiles = ceil_div(K, 64) * a.ctiles * g.ntaps;
  int64_t splits = ceil_div64((int64_t)kNumSMs * 4, tiles);
  const int64_t max_splits = ceil_div64(a.M, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  a.per_split = ceil_div64(ceil_div64(a.M, splits), 16) * 16;
  splits = ceil_div64(a.M, a.per_split);
  dim3 grid(ceil_div(K, 64) * a.ctiles, g.ntaps, (unsigned)splits);
  const bool vec = C % 4 == 0 && K % 4 == 0 && aligned16(x) && aligned16(dy);
  if (vec) conv_wgrad_simt_kernel<true><<<grid, 256, 0, st>>>(a);
  else conv_wgrad_simt_kernel<false><<<grid, 256, 0, st>>>(a);
  IDEAS_CHECK_LAUNCH("conv_wgrad_simt");
  return IDEAS_OK;
}
