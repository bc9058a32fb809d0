// A5: 1x1 convolutions with a tiny channel count on one side (3-channel image ends of the
// networks: E stem 3->32, Dreal 3->64, Dco 3->32, G to_rgb 128->3 -- models.py:240,294,336,383 of the
// reference, all served by cuDNN there).  With <= 4 channels on one side there is no GEMM worth a
// tensor core: these are HBM-bound streaming kernels (the roofline is the large tensor's bytes).
//
//   expand : out[p, k] = act(sum_{c<=4} in[p, c] * w[k][c] + b[k])      bytes ~ 4*P*K   (write-bound)
//   reduce : out[p, j] = sum_c in[p, c] * w[j][c] (+ b[j]),  j <= 4      bytes ~ 4*P*C   (read-bound)
//   wgrad  : dw[k][c] += sum_p dy[p, k] * x[p, c], one of K, C <= 4      bytes ~ 4*P*max(K,C)
//
// Forward of a small-C conv and the data gradient of a small-K conv are "expand"; forward of a
// small-K conv and the data gradient of a small-C conv are "reduce" (the packed weight layouts
// wp[k][c] / wpt[c][k] already have the reduction index contiguous in both cases).
#include "common.cuh"
#include "conv_geom.h"

namespace ideas {

namespace {

constexpr int kSmall = 4;

// one thread = 4 consecutive output channels (float4 store) of a strided set of pixels
template <int CS>
__global__ void __launch_bounds__(256) pointwise_expand_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               int64_t P, int K, int act, float alpha, float gain) {
  const int k4n = K >> 2;
  const int kq = threadIdx.x % k4n;                 // blockDim.x is a multiple of k4n
  const int lanes_p = blockDim.x / k4n;             // pixels handled per CTA per iteration
  const int pl = threadIdx.x / k4n;
  float wr[4][CS];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < CS; ++c) wr[j][c] = __ldg(w + (int64_t)(kq * 4 + j) * CS + c);
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias) + kq);
  for (int64_t p = (int64_t)blockIdx.x * lanes_p + pl; p < P; p += (int64_t)gridDim.x * lanes_p) {
    float xv[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) xv[c] = __ldg(in + p * CS + c);
    float r[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < CS; ++c) r[j] = fmaf(xv[c], wr[j][c], r[j]);
    if (act == IDEAS_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = lrelu(r[j], alpha) * gain;
    }
    st_stream4(out + p * K + kq * 4, make_float4(r[0], r[1], r[2], r[3]));
  }
}

// one warp = PB pixels at a time: lanes stride over the C input channels in float4, JS <= 4 dot products per
// pixel.  With C <= 128 every lane owns one float4 of the pixel, so the weights live in registers and PB
// independent 512-byte loads are in flight per warp before the shuffle reductions start (one pixel at a time
// left the kernel latency-bound at 2.6 TB/s).
template <int JS>
__global__ void __launch_bounds__(256) pointwise_reduce_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               int64_t P, int C, int act, float alpha, float gain) {
  constexpr int PB = 8;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int c4n = C >> 2;
  float bj[JS];
#pragma unroll
  for (int j = 0; j < JS; ++j) bj[j] = bias ? __ldg(bias + j) : 0.f;
  if (c4n <= 32) {
    const bool live = lane < c4n;
    float4 ww[JS];
#pragma unroll
    for (int j = 0; j < JS; ++j)
      ww[j] = live ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)j * C) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t p0 = (int64_t)wid * PB; p0 < P; p0 += (int64_t)warps * PB) {
      float4 x[PB];
#pragma unroll
      for (int i = 0; i < PB; ++i)
        x[i] = (live && p0 + i < P) ? ld_stream4(in + (p0 + i) * C + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float acc[PB][JS];
#pragma unroll
      for (int i = 0; i < PB; ++i)
#pragma unroll
        for (int j = 0; j < JS; ++j)
          acc[i][j] = fmaf(x[i].x, ww[j].x, fmaf(x[i].y, ww[j].y, fmaf(x[i].z, ww[j].z, x[i].w * ww[j].w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < PB; ++i)
#pragma unroll
          for (int j = 0; j < JS; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], o);
      // lane i writes pixel p0 + i (every lane holds every sum after the butterfly)
#pragma unroll
      for (int i = 0; i < PB; ++i) {
        if (lane == i && p0 + i < P) {
#pragma unroll
          for (int j = 0; j < JS; ++j) {
            float r = acc[i][j] + bj[j];
            if (act == IDEAS_ACT_LRELU) r = lrelu(r, alpha) * gain;
            out[(p0 + i) * JS + j] = r;
          }
        }
      }
    }
    return;
  }
  for (int64_t p = wid; p < P; p += warps) {
    float acc[JS];
#pragma unroll
    for (int j = 0; j < JS; ++j) acc[j] = 0.f;
    for (int q = lane; q < c4n; q += 32) {
      const float4 x = ld_stream4(in + p * C + q * 4);
#pragma unroll
      for (int j = 0; j < JS; ++j) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (int64_t)j * C) + q);
        acc[j] = fmaf(x.x, ww.x, fmaf(x.y, ww.y, fmaf(x.z, ww.z, fmaf(x.w, ww.w, acc[j]))));
      }
    }
#pragma unroll
    for (int j = 0; j < JS; ++j)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < JS; ++j) {
        float r = acc[j] + bj[j];
        if (act == IDEAS_ACT_LRELU) r = lrelu(r, alpha) * gain;
        out[p * JS + j] = r;
      }
    }
  }
}

// dw[big][small] (BIG_IS_K) or dw[small][big]: thread = 4 channels of the big side, SS of the small side
template <int SS, bool BIG_IS_K>
__global__ void __launch_bounds__(512) pointwise_wgrad_kernel(float* __restrict__ dw, const float* __restrict__ big,
                                                              const float* __restrict__ small, int64_t P, int NB) {
  extern __shared__ float red[];                    // blockDim.x * 4 * SS floats
  const int b4n = NB >> 2;
  const int bq = threadIdx.x % b4n;
  const int lanes_p = blockDim.x / b4n;
  const int pl = threadIdx.x / b4n;
  float acc[SS][4];
#pragma unroll
  for (int s = 0; s < SS; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[s][j] = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * lanes_p + pl; p < P; p += (int64_t)gridDim.x * lanes_p) {
    const float4 v = ld_stream4(big + p * NB + bq * 4);
#pragma unroll
    for (int s = 0; s < SS; ++s) {
      const float sv = __ldg(small + p * SS + s);
      acc[s][0] = fmaf(sv, v.x, acc[s][0]); acc[s][1] = fmaf(sv, v.y, acc[s][1]);
      acc[s][2] = fmaf(sv, v.z, acc[s][2]); acc[s][3] = fmaf(sv, v.w, acc[s][3]);
    }
  }
  float* mine = red + (size_t)threadIdx.x * 4 * SS;
#pragma unroll
  for (int s = 0; s < SS; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) mine[s * 4 + j] = acc[s][j];
  __syncthreads();
  if ((int)threadIdx.x < b4n) {
    for (int t = threadIdx.x + b4n; t < (int)blockDim.x; t += b4n) {
      const float* o = red + (size_t)t * 4 * SS;
#pragma unroll
      for (int i = 0; i < 4 * SS; ++i) mine[i] += o[i];
    }
#pragma unroll
    for (int s = 0; s < SS; ++s)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int b = threadIdx.x * 4 + j;
        float* dst = BIG_IS_K ? dw + (int64_t)b * SS + s : dw + (int64_t)s * NB + b;
        atomicAdd(dst, mine[s * 4 + j]);
      }
  }
}

int grid_cap(int64_t items, int per_block, int per_sm) {
  int64_t blocks = ceil_div64(items, per_block);
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

bool is_pointwise(const ConvGeom& g) {
  return g.ntaps == 1 && g.taps[0].dy == 0 && g.taps[0].dx == 0 && g.i_s == 1 && g.o_s == 1 && g.IH == g.OH &&
         g.IW == g.OW && g.QH == g.OH && g.QW == g.OW;
}

}  // namespace

// forward-form launch; returns IDEAS_ERR_UNSUPPORTED when the geometry is not a small pointwise conv
int pointwise_conv_launch(const ConvGeom& g, float* dst, const float* src, const float* w, const float* bias, int act,
                          float alpha, float gain, cudaStream_t st) {
  if (!is_pointwise(g)) return IDEAS_ERR_UNSUPPORTED;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const float* wt = w + (int64_t)g.taps[0].widx * g.OC * g.IC;
  if (g.IC <= kSmall && g.OC % 4 == 0 && g.OC / 4 <= 256 && aligned16(dst) && (!bias || aligned16(bias))) {
    const int k4n = g.OC / 4;
    const int threads = k4n * (256 / k4n);
    const int grid = grid_cap(P, threads / k4n * 8, 8);
    switch (g.IC) {
      case 1: pointwise_expand_kernel<1><<<grid, threads, 0, st>>>(dst, src, wt, bias, P, g.OC, act, alpha, gain); break;
      case 2: pointwise_expand_kernel<2><<<grid, threads, 0, st>>>(dst, src, wt, bias, P, g.OC, act, alpha, gain); break;
      case 3: pointwise_expand_kernel<3><<<grid, threads, 0, st>>>(dst, src, wt, bias, P, g.OC, act, alpha, gain); break;
      default: pointwise_expand_kernel<4><<<grid, threads, 0, st>>>(dst, src, wt, bias, P, g.OC, act, alpha, gain); break;
    }
    IDEAS_CHECK_LAUNCH("pointwise_expand");
    return IDEAS_OK;
  }
  if (g.OC <= kSmall && g.IC % 4 == 0 && g.IC >= 32 && aligned16(src) && aligned16(wt)) {
    const int grid = grid_cap(P, 8 * 4, 8);
    switch (g.OC) {
      case 1: pointwise_reduce_kernel<1><<<grid, 256, 0, st>>>(dst, src, wt, bias, P, g.IC, act, alpha, gain); break;
      case 2: pointwise_reduce_kernel<2><<<grid, 256, 0, st>>>(dst, src, wt, bias, P, g.IC, act, alpha, gain); break;
      case 3: pointwise_reduce_kernel<3><<<grid, 256, 0, st>>>(dst, src, wt, bias, P, g.IC, act, alpha, gain); break;
      default: pointwise_reduce_kernel<4><<<grid, 256, 0, st>>>(dst, src, wt, bias, P, g.IC, act, alpha, gain); break;
    }
    IDEAS_CHECK_LAUNCH("pointwise_reduce");
    return IDEAS_OK;
  }
  return IDEAS_ERR_UNSUPPORTED;
}

// g = forward geometry (src = x with IC channels, dst = dy with OC channels); dwp[widx][k][c] accumulated
int pointwise_wgrad_launch(const ConvGeom& g, float* dwp, const float* x, const float* dy, cudaStream_t st) {
  if (!is_pointwise(g)) return IDEAS_ERR_UNSUPPORTED;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  float* dw = dwp + (int64_t)g.taps[0].widx * g.OC * g.IC;
  const bool small_c = g.IC <= kSmall && g.OC % 4 == 0 && g.OC / 4 <= 512 && aligned16(dy);
  const bool small_k = g.OC <= kSmall && g.IC % 4 == 0 && g.IC / 4 <= 512 && aligned16(x);
  if (!small_c && !small_k) return IDEAS_ERR_UNSUPPORTED;
  const int NB = small_c ? g.OC : g.IC, SS = small_c ? g.IC : g.OC;
  const float* big = small_c ? dy : x;
  const float* small = small_c ? x : dy;
  const int b4n = NB / 4;
  const int threads = b4n * (512 / b4n);
  const int grid = grid_cap(P, threads / b4n * 16, 4);
  const size_t smem = (size_t)threads * 4 * SS * sizeof(float);
#define IDEAS_PW_WGRAD(S)                                                                                        \
  if (small_c) pointwise_wgrad_kernel<S, true><<<grid, threads, smem, st>>>(dw, big, small, P, NB);             \
  else pointwise_wgrad_kernel<S, false><<<grid, threads, smem, st>>>(dw, big, small, P, NB)
  switch (SS) {
    case 1: IDEAS_PW_WGRAD(1); break;
    case 2: IDEAS_PW_WGRAD(2); break;
    case 3: IDEAS_PW_WGRAD(3); break;
    default: IDEAS_PW_WGRAD(4); break;
  }
#undef IDEAS_PW_WGRAD
  IDEAS_CHECK_LAUNCH("pointwise_wgrad");
  return IDEAS_OK;
}

}  // namespace ideas
