// Library plumbing (error string, version) and the NHWC helpers of the modulation path:
// channel scaling x*s[n,c] (stylegan2/model.py:239-240 applied to activations instead of
// weights), the per-(sample,channel) dot products its backward needs, and the residual merge
// (out+skip)/sqrt(2) of models.py:178,227.  All HBM-bound, 128-bit vectorised.
#include "common.cuh"

#include <atomic>

namespace ideas {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

__global__ void __launch_bounds__(256) scale_channels_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                             const float* __restrict__ s, int64_t pc4, int c4n,
                                                             int64_t total4) {
  // element v (float4 index) -> sample n = v / pc4, channel group = v % c4n
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total4; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = v / pc4;
    const int cg = (int)(v % c4n);
    const float4 a = ld_stream4(x + v * 4);
    const float4 m = __ldg(reinterpret_cast<const float4*>(s + n * c4n * 4) + cg);
    st_stream4(out + v * 4, make_float4(a.x * m.x, a.y * m.y, a.z * m.z, a.w * m.w));
  }
}

__global__ void __launch_bounds__(256) scale_channels_scalar_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                                    const float* __restrict__ s, int64_t pc, int C,
                                                                    int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] * __ldg(s + (i / pc) * C + (i % C));
}

// grid (slices, N); blockDim multiple of C/4 so each thread owns one channel group
__global__ void __launch_bounds__(512) channel_dot_kernel(float* __restrict__ dot, float* __restrict__ out,
                                                          const float* __restrict__ a, const float* __restrict__ b,
                                                          const float* __restrict__ s, int64_t P, int c4n) {
  extern __shared__ float4 red[];
  const int n = blockIdx.y;
  const int64_t pc4 = P * c4n;
  const float* an = a + (int64_t)n * pc4 * 4;
  const float* bn = b + (int64_t)n * pc4 * 4;
  float* on = out ? out + (int64_t)n * pc4 * 4 : nullptr;
  const int cg = threadIdx.x % c4n;
  float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
  if (s) m = __ldg(reinterpret_cast<const float4*>(s + (int64_t)n * c4n * 4) + cg);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < pc4; v += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = ld_stream4(an + v * 4);
    const float4 y = ld_stream4(bn + v * 4);
    acc.x = fmaf(x.x, y.x, acc.x); acc.y = fmaf(x.y, y.y, acc.y);
    acc.z = fmaf(x.z, y.z, acc.z); acc.w = fmaf(x.w, y.w, acc.w);
    if (on) st_stream4(on + v * 4, make_float4(y.x * m.x, y.y * m.y, y.z * m.z, y.w * m.w));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < c4n) {
    float4 t = red[threadIdx.x];
    for (int j = threadIdx.x + c4n; j < (int)blockDim.x; j += c4n) {
      const float4 o = red[j];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    float* d = dot + ((int64_t)n * c4n + threadIdx.x) * 4;
    atomicAdd(d + 0, t.x); atomicAdd(d + 1, t.y); atomicAdd(d + 2, t.z); atomicAdd(d + 3, t.w);
  }
}

// generic fallback (any C): one CTA per (n, c)
__global__ void __launch_bounds__(256) channel_dot_scalar_kernel(float* __restrict__ dot, float* __restrict__ out,
                                                                 const float* __restrict__ a,
                                                                 const float* __restrict__ b,
                                                                 const float* __restrict__ s, int64_t P, int C) {
  const int n = blockIdx.y, c = blockIdx.x;
  const float m = s ? s[(int64_t)n * C + c] : 1.f;
  float acc = 0.f;
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const int64_t i = ((int64_t)n * P + p) * C + c;
    acc = fmaf(a[i], b[i], acc);
    if (out) out[i] = b[i] * m;
  }
  __shared__ float sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(dot + (int64_t)n * C + c, sm[0]);
}

__global__ void __launch_bounds__(256) add_scale_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                        const float* __restrict__ b, float gain, int64_t n4) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = ld_stream4(a + v * 4), y = ld_stream4(b + v * 4);
    st_stream4(out + v * 4, make_float4((x.x + y.x) * gain, (x.y + y.y) * gain, (x.z + y.z) * gain, (x.w + y.w) * gain));
  }
}
__global__ void __launch_bounds__(256) add_scale_scalar_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                               const float* __restrict__ b, float gain, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (a[i] + b[i]) * gain;
}

static int blocks_for(int64_t items, int threads, int per_sm) {
  int64_t b = ceil_div64(items, threads);
  if (b > (int64_t)kNumSMs * per_sm) b = (int64_t)kNumSMs * per_sm;
  return b < 1 ? 1 : (int)b;
}

}  // namespace ideas

using namespace ideas;

extern "C" int ideas_abi_version(void) { return 1; }
extern "C" const char* ideas_last_error(void) { return g_err; }
extern "C" unsigned long long ideas_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int ideas_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "device_cc");
  return major * 10 + minor;
}

extern "C" int ideas_scale_channels(float* out, const float* x, const float* s, int N, int64_t P, int C, void* stream) {
  IDEAS_REQUIRE(N >= 0 && P >= 0 && C >= 1, "scale_channels: bad shape");
  const int64_t total = (int64_t)N * P * C;
  if (total == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && x && s, "scale_channels: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 4 == 0 && aligned16(out) && aligned16(x) && aligned16(s)) {
    scale_channels_kernel<<<blocks_for(total / 4, 256, 8), 256, 0, st>>>(out, x, s, P * (C / 4), C / 4, total / 4);
  } else {
    scale_channels_scalar_kernel<<<blocks_for(total, 256, 8), 256, 0, st>>>(out, x, s, P * C, C, total);
  }
  IDEAS_CHECK_LAUNCH("scale_channels");
  return IDEAS_OK;
}

extern "C" int ideas_channel_dot(float* dot, float* out, const float* a, const float* b, const float* s, int N,
                                 int64_t P, int C, void* stream) {
  IDEAS_REQUIRE(N >= 0 && P >= 0 && C >= 1, "channel_dot: bad shape");
  if ((int64_t)N * P == 0) return IDEAS_OK;
  IDEAS_REQUIRE(dot && a && b && (!out || s), "channel_dot: null pointer");
  IDEAS_REQUIRE(N <= 65535, "channel_dot: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int c4n = C / 4;
  if (C % 4 == 0 && c4n <= 512 && aligned16(a) && aligned16(b) && (!out || aligned16(out)) && (!s || aligned16(s))) {
    const int threads = c4n * (512 / c4n);
    int slices = (int)ceil_div64(P * c4n, (int64_t)threads * 8);
    const int cap = ceil_div(kNumSMs * 4, N);
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    channel_dot_kernel<<<dim3(slices, N), threads, threads * sizeof(float4), st>>>(dot, out, a, b, s, P, c4n);
  } else {
    channel_dot_scalar_kernel<<<dim3(C, N), 256, 0, st>>>(dot, out, a, b, s, P, C);
  }
  IDEAS_CHECK_LAUNCH("channel_dot");
  return IDEAS_OK;
}

extern "C" int ideas_add_scale(float* out, const float* a, const float* b, float gain, int64_t n, void* stream) {
  IDEAS_REQUIRE(n >= 0, "add_scale: negative size");
  if (n == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && a && b, "add_scale: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (n % 4 == 0 && aligned16(out) && aligned16(a) && aligned16(b))
    add_scale_kernel<<<blocks_for(n / 4, 256, 8), 256, 0, st>>>(out, a, b, gain, n / 4);
  else
    add_scale_scalar_kernel<<<blocks_for(n, 256, 8), 256, 0, st>>>(out, a, b, gain, n);
  IDEAS_CHECK_LAUNCH("add_scale");
  return IDEAS_OK;
}

// ---------------------------------------------------------------------------------------
// patchify: the reference cuts n_crop random boxes out of every image and
// bilinearly resizes each to (th, tw) with a Python loop of F.interpolate calls (utils.py:127-149).
// Here the boxes are DATA (int32 (n_crop, 4) = y, x, h, w on the device, shared by the whole batch as in
// the reference), so one launch serves any crop geometry -- which also makes the step CUDA-graph safe.
// Sampling follows F.interpolate(mode="bilinear", align_corners=False): src = (dst + 0.5) * (h / th) - 0.5,
// clamped at 0, neighbours clamped at the crop edge.  NHWC, one thread per output pixel.
// ---------------------------------------------------------------------------------------
namespace ideas {

struct BilinearTap { int i0, i1; float l0, l1; };

__device__ __forceinline__ BilinearTap bilinear_tap(int dst, int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  BilinearTap t;
  t.i0 = (int)src;
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}

// Boxes are data written by the host (or a device RNG): clamp them into the image so that a malformed row can
// never turn into an out-of-bounds access (a well-formed box is returned unchanged).
struct CropBox { int y, x, h, w; };

__device__ __forceinline__ CropBox load_box(const int* b, int H, int W) {
  CropBox c;
  c.h = min(max(__ldg(b + 2), 1), H);
  c.w = min(max(__ldg(b + 3), 1), W);
  c.y = min(max(__ldg(b), 0), H - c.h);
  c.x = min(max(__ldg(b + 1), 0), W - c.w);
  return c;
}

__global__ void __launch_bounds__(256) patchify_fwd_kernel(float* __restrict__ out, const float* __restrict__ img,
                                                           const int* __restrict__ boxes, int B, int H, int W, int C,
                                                           int n_crop, int th, int tw) {
  const int64_t total = (int64_t)B * n_crop * th * tw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int tx = (int)(idx % tw);
    int64_t t = idx / tw;
    const int ty = (int)(t % th);
    t /= th;
    const int j = (int)(t % n_crop);
    const int b = (int)(t / n_crop);
    const CropBox bx = load_box(boxes + j * 4, H, W);
    const int y0 = bx.y, x0 = bx.x, h = bx.h, w = bx.w;
    const BilinearTap ty_ = bilinear_tap(ty, h, th), tx_ = bilinear_tap(tx, w, tw);
    const float* base = img + (int64_t)b * H * W * C;
    const float* p00 = base + ((int64_t)(y0 + ty_.i0) * W + (x0 + tx_.i0)) * C;
    const float* p01 = base + ((int64_t)(y0 + ty_.i0) * W + (x0 + tx_.i1)) * C;
    const float* p10 = base + ((int64_t)(y0 + ty_.i1) * W + (x0 + tx_.i0)) * C;
    const float* p11 = base + ((int64_t)(y0 + ty_.i1) * W + (x0 + tx_.i1)) * C;
    float* o = out + idx * C;
    for (int c = 0; c < C; ++c)
      o[c] = ty_.l0 * (tx_.l0 * __ldg(p00 + c) + tx_.l1 * __ldg(p01 + c)) +
             ty_.l1 * (tx_.l0 * __ldg(p10 + c) + tx_.l1 * __ldg(p11 + c));
  }
}

__global__ void __launch_bounds__(256) patchify_bwd_kernel(float* __restrict__ gimg, const float* __restrict__ gout,
                                                           const int* __restrict__ boxes, int B, int H, int W, int C,
                                                           int n_crop, int th, int tw) {
  const int64_t total = (int64_t)B * n_crop * th * tw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int tx = (int)(idx % tw);
    int64_t t = idx / tw;
    const int ty = (int)(t % th);
    t /= th;
    const int j = (int)(t % n_crop);
    const int b = (int)(t / n_crop);
    const CropBox bx = load_box(boxes + j * 4, H, W);
    const int y0 = bx.y, x0 = bx.x, h = bx.h, w = bx.w;
    const BilinearTap ty_ = bilinear_tap(ty, h, th), tx_ = bilinear_tap(tx, w, tw);
    float* base = gimg + (int64_t)b * H * W * C;
    float* p00 = base + ((int64_t)(y0 + ty_.i0) * W + (x0 + tx_.i0)) * C;
    float* p01 = base + ((int64_t)(y0 + ty_.i0) * W + (x0 + tx_.i1)) * C;
    float* p10 = base + ((int64_t)(y0 + ty_.i1) * W + (x0 + tx_.i0)) * C;
    float* p11 = base + ((int64_t)(y0 + ty_.i1) * W + (x0 + tx_.i1)) * C;
    const float* g = gout + idx * C;
    for (int c = 0; c < C; ++c) {
      const float v = g[c];
      atomicAdd(p00 + c, ty_.l0 * tx_.l0 * v);
      atomicAdd(p01 + c, ty_.l0 * tx_.l1 * v);
      atomicAdd(p10 + c, ty_.l1 * tx_.l0 * v);
      atomicAdd(p11 + c, ty_.l1 * tx_.l1 * v);
    }
  }
}

}  // namespace ideas

static int check_patchify(const char* who, int B, int H, int W, int C, int n_crop, int th, int tw) {
  IDEAS_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 1 && n_crop >= 0 && th >= 1 && tw >= 1, "%s: bad shape", who);
  return IDEAS_OK;
}

extern "C" int ideas_patchify_forward(float* out, const float* img, const int* boxes, int B, int H, int W, int C,
                                      int n_crop, int th, int tw, void* stream) {
  int rc = check_patchify("patchify_forward", B, H, W, C, n_crop, th, tw);
  if (rc) return rc;
  const int64_t total = (int64_t)B * n_crop * th * tw;
  if (total == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && img && boxes, "patchify_forward: null pointer");
  int64_t blocks = ideas::ceil_div64(total, 256);
  if (blocks > ideas::kNumSMs * 8) blocks = ideas::kNumSMs * 8;
  ideas::patchify_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, img, boxes, B, H, W, C, n_crop, th, tw);
  IDEAS_CHECK_LAUNCH("patchify_forward");
  return IDEAS_OK;
}

extern "C" int ideas_patchify_backward(float* gimg, const float* gout, const int* boxes, int B, int H, int W, int C,
                                       int n_crop, int th, int tw, void* stream) {
  int rc = check_patchify("patchify_backward", B, H, W, C, n_crop, th, tw);
  if (rc) return rc;
  const int64_t total = (int64_t)B * n_crop * th * tw;
  if (total == 0) return IDEAS_OK;
  IDEAS_REQUIRE(gimg && gout && boxes, "patchify_backward: null pointer");
  int64_t blocks = ideas::ceil_div64(total, 256);
  if (blocks > ideas::kNumSMs * 8) blocks = ideas::kNumSMs * 8;
  ideas::patchify_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(gimg, gout, boxes, B, H, W, C, n_crop, th, tw);
  IDEAS_CHECK_LAUNCH("patchify_backward");
  return IDEAS_OK;
}

// ---------------------------------------------------------------------------------------
// reflection padding on NHWC data (nn.ReflectionPad2d of the reference's ResBlocks, models.py:102-108).
// torch's CUDA kernel returns an NCHW-contiguous tensor for a channels_last input (and an NCHW gradient), which
// costs a layout-conversion copy on each side of every pad; these two gather kernels stay in NHWC.
//   forward : out[n, y, x, :] = in[n, r(y - p, H), r(x - p, W), :],  r(i, L) = i < 0 ? -i : (i >= L ? 2(L-1) - i : i)
//   backward: gin[n, y, x, :] = sum of gout over the (<= 3 x 3) padded positions that read (y, x)
// ---------------------------------------------------------------------------------------
namespace ideas {

__device__ __forceinline__ int reflect_index(int i, int L) { return i < 0 ? -i : (i >= L ? 2 * (L - 1) - i : i); }

template <typename V>
__global__ void __launch_bounds__(256) reflect_pad_kernel(V* __restrict__ out, const V* __restrict__ in, int H, int W, int cv,
                                                          int p, int64_t total) {
  const int OH = H + 2 * p, OW = W + 2 * p;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % cv);
    int64_t t = i / cv;
    const int x = (int)(t % OW);
    t /= OW;
    const int y = (int)(t % OH);
    const int64_t n = t / OH;
    out[i] = in[((n * H + reflect_index(y - p, H)) * W + reflect_index(x - p, W)) * cv + c];
  }
}

__device__ __forceinline__ void acc_add(float& a, const float& b) { a += b; }
__device__ __forceinline__ void acc_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void acc_zero(float& a) { a = 0.f; }
__device__ __forceinline__ void acc_zero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }

// padded coordinates that read source index i (axis length L, pad p): i + p, and the two mirror images
__device__ __forceinline__ int reflect_sources(int i, int L, int p, int (&o)[3]) {
  int n = 0;
  o[n++] = i + p;
  if (i >= 1 && i <= p) o[n++] = p - i;
  if (i >= L - 1 - p && i <= L - 2) o[n++] = p + 2 * (L - 1) - i;
  return n;
}

template <typename V>
__global__ void __launch_bounds__(256) reflect_pad_bwd_kernel(V* __restrict__ gin, const V* __restrict__ gout, int H, int W,
                                                              int cv, int p, int64_t total) {
  const int OH = H + 2 * p, OW = W + 2 * p;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % cv);
    int64_t t = i / cv;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int64_t n = t / H;
    int ys[3], xs[3];
    const int ny = reflect_sources(y, H, p, ys), nx = reflect_sources(x, W, p, xs);
    V acc;
    acc_zero(acc);
    for (int a = 0; a < ny; ++a)
      for (int b = 0; b < nx; ++b) acc_add(acc, gout[((n * OH + ys[a]) * OW + xs[b]) * cv + c]);
    gin[i] = acc;
  }
}

}  // namespace ideas

extern "C" int ideas_reflect_pad2d(float* out, const float* x, int N, int H, int W, int C, int pad, int backward,
                                   void* stream) {
  IDEAS_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 1, "reflect_pad2d: bad shape");
  IDEAS_REQUIRE(pad >= 0 && pad < H && pad < W, "reflect_pad2d: padding %d must be smaller than the input (%dx%d)", pad, H, W);
  if (N == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && x, "reflect_pad2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = C % 4 == 0 && ideas::aligned16(out) && ideas::aligned16(x);
  const int cv = vec ? C / 4 : C;
  // `out` is the padded tensor in the forward and the un-padded gradient in the backward
  const int64_t total = backward ? (int64_t)N * H * W * cv : (int64_t)N * (H + 2 * pad) * (W + 2 * pad) * cv;
  const int blocks = blocks_for(total, 256, 8);
  if (vec) {
    if (backward) ideas::reflect_pad_bwd_kernel<float4><<<blocks, 256, 0, st>>>((float4*)out, (const float4*)x, H, W, cv, pad, total);
    else ideas::reflect_pad_kernel<float4><<<blocks, 256, 0, st>>>((float4*)out, (const float4*)x, H, W, cv, pad, total);
  } else {
    if (backward) ideas::reflect_pad_bwd_kernel<float><<<blocks, 256, 0, st>>>(out, x, H, W, cv, pad, total);
    else ideas::reflect_pad_kernel<float><<<blocks, 256, 0, st>>>(out, x, H, W, cv, pad, total);
  }
  IDEAS_CHECK_LAUNCH("reflect_pad2d");
  return IDEAS_OK;
}
