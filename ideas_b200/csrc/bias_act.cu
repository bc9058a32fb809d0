// A3: fused bias + leaky ReLU, forward / first derivative / fused backward with bias-gradient.
// Replaces stylegan2/op/fused_bias_act.cpp:11-20 + fused_bias_act_kernel.cu:18-99 of the
// reference (one scalar 4-byte load per thread-iteration, integer div/mod per element,
// separate .sum() pass for the bias gradient, 32-bit indexing).
//
// B200 design: HBM-bound, 8 B/element forward, 12 B/element backward.  128-bit streaming
// loads/stores, 64-bit indexing, grid = k x 148 SMs with a grid-stride loop, and the bias
// gradient accumulated in registers per thread (each thread owns a fixed group of four
// channels), reduced through shared memory and flushed with one atomic per channel per CTA.
#include "common.cuh"

namespace ideas {

enum ChMode { CH_NONE = 0, CH_INNER = 1, CH_OUTER = 2 };

template <int CODE>  // act*10+grad, fused_bias_act_kernel.cu:36-47
__device__ __forceinline__ float act_apply(float v, float ref, float alpha, float scale) {
  float y;
  if (CODE == 30) y = (v > 0.f) ? v : v * alpha;
  else if (CODE == 31) y = (ref > 0.f) ? v : v * alpha;
  else if (CODE == 12 || CODE == 32) y = 0.f;
  else y = v;
  return y * scale;
}

template <int CODE, int CH>
__global__ void __launch_bounds__(256) bias_act_vec_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                           const float* __restrict__ bias, const float* __restrict__ ref,
                                                           float alpha, float scale, int64_t nvec, int64_t step_b,
                                                           int size_b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float4 a = ld_stream4(x + v * 4);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (CH == CH_INNER) {
      b = *reinterpret_cast<const float4*>(bias + (v * 4) % size_b);
    } else if (CH == CH_OUTER) {
      float s = __ldg(bias + ((v * 4) / step_b) % size_b);
      b = make_float4(s, s, s, s);
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (CODE == 31) r = ld_stream4(ref + v * 4);
    float4 y;
    y.x = act_apply<CODE>(a.x + b.x, r.x, alpha, scale);
    y.y = act_apply<CODE>(a.y + b.y, r.y, alpha, scale);
    y.z = act_apply<CODE>(a.z + b.z, r.z, alpha, scale);
    y.w = act_apply<CODE>(a.w + b.w, r.w, alpha, scale);
    st_stream4(out + v * 4, y);
  }
}

template <int CODE>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ ref, float alpha, float scale,
                                                              int64_t n, int64_t step_b, int size_b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = x[i];
    if (bias) v += __ldg(bias + (i / step_b) % size_b);
    float r = (CODE == 31) ? ref[i] : 0.f;
    out[i] = act_apply<CODE>(v, r, alpha, scale);
  }
}

// Fused backward for channel-innermost data (NHWC or (B,C)), C % 4 == 0, C/4 <= blockDim.
// blockDim.x is a multiple of C/4, so thread t always sees channel group t % (C/4).
__global__ void __launch_bounds__(512) bias_act_bwd_inner_kernel(float* __restrict__ gx, float* __restrict__ gbias,
                                                                 const float* __restrict__ gy,
                                                                 const float* __restrict__ outp, float alpha,
                                                                 float scale, int64_t nvec, int cvec) {
  extern __shared__ float4 red[];  // blockDim.x entries
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;  // multiple of cvec
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    float4 g = ld_stream4(gy + v * 4);
    float4 o = ld_stream4(outp + v * 4);
    float4 r;
    r.x = ((o.x > 0.f) ? g.x : g.x * alpha) * scale;
    r.y = ((o.y > 0.f) ? g.y : g.y * alpha) * scale;
    r.z = ((o.z > 0.f) ? g.z : g.z * alpha) * scale;
    r.w = ((o.w > 0.f) ? g.w : g.w * alpha) * scale;
    st_stream4(gx + v * 4, r);
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  if (gbias == nullptr) return;
  red[threadIdx.x] = acc;
  __syncthreads();
  // threads 0..cvec-1 fold the replicas of their channel group
  if ((int)threadIdx.x < cvec) {
    float4 s = red[threadIdx.x];
    for (int t = threadIdx.x + cvec; t < (int)blockDim.x; t += cvec) {
      float4 o = red[t];
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    float* dst = gbias + threadIdx.x * 4;
    atomicAdd(dst + 0, s.x);
    atomicAdd(dst + 1, s.y);
    atomicAdd(dst + 2, s.z);
    atomicAdd(dst + 3, s.w);
  }
}

// Generic per-channel sum: x viewed (outer, C, inner); one CTA per (channel, slice of outer).
__global__ void __launch_bounds__(256) channel_sum_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                          int64_t outer, int C, int64_t inner) {
  const int c = blockIdx.x;
  float acc = 0.f;
  const int64_t per = outer * inner;
  for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < per; j += (int64_t)gridDim.y * blockDim.x) {
    int64_t o = j / inner, i = j - o * inner;
    acc += x[(o * C + c) * inner + i];
  }
  __shared__ float sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out + c, sm[0]);
}


// Backward of  out = lrelu(d[n,c]*u + bias[c]) * gain  for NHWC data, in one pass:
//   g1 = gy * (out > 0 ? 1 : alpha) * gain            gradient w.r.t. the pre-activation z
//   g1d[n,p,c] = g1 * d[n,c]                          gradient w.r.t. the un-demodulated conv output u
//   gsum[n,c] += sum_p g1                             (bias gradient = sum over n)
//   dotz[n,c] += sum_p g1 * z,  z = lrelu^-1(out)     (gives dL/dd = (dotz - bias*gsum)/d)
// grid (slices, N); blockDim multiple of C/4.
__global__ void __launch_bounds__(512) modconv_act_bwd_kernel(float* __restrict__ g1d, float* __restrict__ gsum,
                                                              float* __restrict__ dotz, const float* __restrict__ gy,
                                                              const float* __restrict__ outp,
                                                              const float* __restrict__ d, float alpha, float gain,
                                                              float inv_pos, float inv_neg, int64_t P, int c4n) {
  extern __shared__ float4 red[];  // 2 * blockDim.x
  const int n = blockIdx.y;
  const int64_t pc4 = P * c4n;
  const int64_t base = (int64_t)n * pc4 * 4;
  const int cg = threadIdx.x % c4n;
  float4 dd = make_float4(1.f, 1.f, 1.f, 1.f);
  if (d) dd = __ldg(reinterpret_cast<const float4*>(d + (int64_t)n * c4n * 4) + cg);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < pc4; v += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = ld_stream4(gy + base + v * 4);
    const float4 o = ld_stream4(outp + base + v * 4);
    float4 r, z;
    r.x = ((o.x > 0.f) ? g.x : g.x * alpha) * gain; z.x = o.x * ((o.x > 0.f) ? inv_pos : inv_neg);
    r.y = ((o.y > 0.f) ? g.y : g.y * alpha) * gain; z.y = o.y * ((o.y > 0.f) ? inv_pos : inv_neg);
    r.z = ((o.z > 0.f) ? g.z : g.z * alpha) * gain; z.z = o.z * ((o.z > 0.f) ? inv_pos : inv_neg);
    r.w = ((o.w > 0.f) ? g.w : g.w * alpha) * gain; z.w = o.w * ((o.w > 0.f) ? inv_pos : inv_neg);
    st_stream4(g1d + base + v * 4, make_float4(r.x * dd.x, r.y * dd.y, r.z * dd.z, r.w * dd.w));
    s1.x += r.x; s1.y += r.y; s1.z += r.z; s1.w += r.w;
    s2.x = fmaf(r.x, z.x, s2.x); s2.y = fmaf(r.y, z.y, s2.y); s2.z = fmaf(r.z, z.z, s2.z); s2.w = fmaf(r.w, z.w, s2.w);
  }
  red[threadIdx.x] = s1;
  red[blockDim.x + threadIdx.x] = s2;
  __syncthreads();
  if ((int)threadIdx.x < c4n) {
    float4 a = red[threadIdx.x], b = red[blockDim.x + threadIdx.x];
    for (int t = threadIdx.x + c4n; t < (int)blockDim.x; t += c4n) {
      const float4 oa = red[t], ob = red[blockDim.x + t];
      a.x += oa.x; a.y += oa.y; a.z += oa.z; a.w += oa.w;
      b.x += ob.x; b.y += ob.y; b.z += ob.z; b.w += ob.w;
    }
    float* p1 = gsum + ((int64_t)n * c4n + threadIdx.x) * 4;
    float* p2 = dotz + ((int64_t)n * c4n + threadIdx.x) * 4;
    atomicAdd(p1 + 0, a.x); atomicAdd(p1 + 1, a.y); atomicAdd(p1 + 2, a.z); atomicAdd(p1 + 3, a.w);
    atomicAdd(p2 + 0, b.x); atomicAdd(p2 + 1, b.y); atomicAdd(p2 + 2, b.z); atomicAdd(p2 + 3, b.w);
  }
}

static int grid_for(int64_t work_items, int threads, int per_sm) {
  int64_t blocks = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)kNumSMs * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int CODE>
static int launch_bias_act(float* out, const float* x, const float* bias, const float* ref, float alpha, float scale,
                           int64_t n, int64_t step_b, int size_b, cudaStream_t st) {
  const bool vec_ok = (n % 4 == 0) && aligned16(out) && aligned16(x) && (CODE != 31 || aligned16(ref));
  int mode = -1;
  if (vec_ok) {
    if (!bias) mode = CH_NONE;
    else if (step_b == 1 && size_b % 4 == 0 && aligned16(bias)) mode = CH_INNER;
    else if (step_b % 4 == 0) mode = CH_OUTER;
  }
  if (mode >= 0) {
    const int64_t nvec = n / 4;
    const int grid = grid_for(nvec, 256, 8);
    if (mode == CH_NONE)
      bias_act_vec_kernel<CODE, CH_NONE><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, nvec, step_b, size_b);
    else if (mode == CH_INNER)
      bias_act_vec_kernel<CODE, CH_INNER><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, nvec, step_b, size_b);
    else
      bias_act_vec_kernel<CODE, CH_OUTER><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, nvec, step_b, size_b);
  } else {
    const int grid = grid_for(n, 256, 8);
    bias_act_scalar_kernel<CODE><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, n, step_b, size_b);
  }
  IDEAS_CHECK_LAUNCH("fused_bias_act");
  return IDEAS_OK;
}

}  // namespace ideas

using namespace ideas;

extern "C" int ideas_fused_bias_act(float* out, const float* x, const float* bias, const float* ref, int act, int grad,
                                    float alpha, float scale, int64_t n, int64_t step_b, int size_b, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IDEAS_REQUIRE(n >= 0, "fused_bias_act: negative element count");
  if (n == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && x, "fused_bias_act: null input/output");
  IDEAS_REQUIRE(!bias || (step_b >= 1 && size_b >= 1), "fused_bias_act: bad bias geometry step_b=%lld size_b=%d",
                (long long)step_b, size_b);
  if (!bias) { step_b = 1; size_b = 1; }
  const int code = act * 10 + grad;
  IDEAS_REQUIRE(code != 31 || ref, "fused_bias_act: grad=1 needs the reference (forward output) tensor");
  switch (code) {
    case 30: return launch_bias_act<30>(out, x, bias, ref, alpha, scale, n, step_b, size_b, st);
    case 31: return launch_bias_act<31>(out, x, bias, ref, alpha, scale, n, step_b, size_b, st);
    case 12:
    case 32: return launch_bias_act<32>(out, x, bias, ref, alpha, scale, n, step_b, size_b, st);
    default: return launch_bias_act<10>(out, x, bias, ref, alpha, scale, n, step_b, size_b, st);  // kernel.cu:37-40
  }
}

extern "C" int ideas_bias_act_backward(float* gx, float* gbias, const float* gy, const float* out, float alpha,
                                       float scale, int64_t n, int64_t step_b, int size_b, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IDEAS_REQUIRE(n >= 0 && step_b >= 1 && size_b >= 1, "bias_act_backward: bad geometry");
  if (n == 0) return IDEAS_OK;
  IDEAS_REQUIRE(gx && gy && out, "bias_act_backward: null pointer");
  const int cvec = size_b / 4;
  const bool fused = step_b == 1 && size_b % 4 == 0 && cvec <= 512 && n % size_b == 0 && aligned16(gx) &&
                     aligned16(gy) && aligned16(out);
  if (fused) {
    const int threads = cvec * (512 / cvec);
    const int64_t nvec = n / 4;
    const int grid = grid_for(nvec, threads, 4);
    bias_act_bwd_inner_kernel<<<grid, threads, threads * sizeof(float4), st>>>(gx, gbias, gy, out, alpha, scale, nvec,
                                                                               cvec);
    IDEAS_CHECK_LAUNCH("bias_act_backward");
    return IDEAS_OK;
  }
  int rc = launch_bias_act<31>(gx, gy, nullptr, out, alpha, scale, n, 1, 1, st);
  if (rc != IDEAS_OK || !gbias) return rc;
  IDEAS_REQUIRE(n % (step_b * size_b) == 0, "bias_act_backward: n not a multiple of C*inner");
  const int64_t outer = n / (step_b * size_b);
  int slices = (int)ceil_div64(outer * step_b, 256 * 64);
  if (slices < 1) slices = 1;
  if (slices > 64) slices = 64;
  channel_sum_kernel<<<dim3(size_b, slices), 256, 0, st>>>(gbias, gx, outer, size_b, step_b);
  IDEAS_CHECK_LAUNCH("bias_act_backward/channel_sum");
  return IDEAS_OK;
}

extern "C" int ideas_modconv_act_backward(float* g1d, float* gsum, float* dotz, const float* gy, const float* out,
                                          const float* d, float alpha, float gain, int N, int64_t P, int C,
                                          void* stream) {
  IDEAS_REQUIRE(N >= 0 && P >= 0 && C >= 1, "modconv_act_backward: bad shape");
  if ((int64_t)N * P == 0) return IDEAS_OK;
  IDEAS_REQUIRE(g1d && gsum && dotz && gy && out, "modconv_act_backward: null pointer");
  IDEAS_REQUIRE(alpha != 0.f && gain != 0.f, "modconv_act_backward: activation not invertible (alpha or gain is 0)");
  const int c4n = C / 4;
  if (!(C % 4 == 0 && c4n <= 512 && N <= 65535 && aligned16(g1d) && aligned16(gy) && aligned16(out) &&
        (!d || aligned16(d)))) {
    set_error("modconv_act_backward: needs C %% 4 == 0, C <= 2048 and 16-byte aligned tensors (C=%d)", C);
    return IDEAS_ERR_UNSUPPORTED;
  }
  const int threads = c4n * (512 / c4n);
  int slices = (int)ceil_div64(P * c4n, (int64_t)threads * 8);
  const int cap = ceil_div(kNumSMs * 4, N);
  if (slices > cap) slices = cap;
  if (slices < 1) slices = 1;
  modconv_act_bwd_kernel<<<dim3(slices, N), threads, 2 * threads * sizeof(float4), (cudaStream_t)stream>>>(
      g1d, gsum, dotz, gy, out, d, alpha, gain, 1.0f / gain, 1.0f / (gain * alpha), P, c4n);
  IDEAS_CHECK_LAUNCH("modconv_act_backward");
  return IDEAS_OK;
}
