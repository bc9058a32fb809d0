// A4: upfirdn2d (zero-insert upsample -> pad/crop -> 2-D FIR -> decimate).
// Replaces stylegan2/op/upfirdn2d.cpp:12-22 + upfirdn2d_kernel.cu:17-369 of the reference
// (16x64 shared-memory tiles per 256-thread CTA, `volatile` smem, scalar loads, one (b,c)
// plane per blockIdx.z).
//
// B200 design.  IDEAS only ever runs up = down = 1 with the 4x4 binomial kernel and
// asymmetric zero padding: an HBM-bound stencil, 4 B read + 4 B written
// per element.  Activations are NHWC, so the fast path assigns one thread to a group of
// four channels (128-bit, fully coalesced across the warp) of one output column and lets
// it slide down a strip of rows: every input row is loaded once per thread (4 taps along
// x, neighbours hit L1), multiplied into the four kernel rows and accumulated into three
// pending output rows held in registers -- no shared memory, no barriers.  An optional
// epilogue applies bias + leaky ReLU (the Blur -> FusedLeakyReLU pair of an upsampling
// StyledConv) so that tensor makes one HBM round trip instead of two.
// Every other (up, down, kernel size, minor) combination -- used by the reference only in
// Upsample/Downsample/ADA -- goes through a direct-form kernel with the same semantics.
#include "common.cuh"

#include <atomic>
#include <type_traits>

namespace ideas {

static std::atomic<int> g_blur_variant{0};
int blur_variant() { return g_blur_variant.load(); }
void set_blur_variant(int v) { g_blur_variant.store(v); }
static std::atomic<int> g_resample_variant{0};      // option "resample_variant": strip shapes of the up2 / down2 kernels
void set_resample_variant(int v) { g_resample_variant.store(v); }

struct UpfirdnParams {
  int major, in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
  int out_h, out_w;
  float alpha, gain;
  const float* residual;   // out-shaped; merged as (result + residual) * res_scale (up2 fast path only)
  float res_scale;
  float* out2;             // EPI 1 only: second output out2 = out * post[n,c] (the next layer's pre-modulated input)
  const float* post;
};

// ---------------------------------------------------------------------------------------
// generic direct form, any parameters
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                                const float* __restrict__ kernel,
                                                                const float* __restrict__ bias, UpfirdnParams p,
                                                                int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int c = (int)(idx % p.minor);
    int64_t t = idx / p.minor;
    int ox = (int)(t % p.out_w);
    t /= p.out_w;
    int oy = (int)(t % p.out_h);
    int m = (int)(t / p.out_h);
    float v = 0.f;
    for (int ky = 0; ky < p.kh; ++ky) {
      int uy = oy * p.down_y + ky - p.pad_y0;  // position in the zero-inserted signal
      if (uy < 0 || uy % p.up_y) continue;
      int iy = uy / p.up_y;
      if (iy >= p.in_h) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        int ux = ox * p.down_x + kx - p.pad_x0;
        if (ux < 0 || ux % p.up_x) continue;
        int ix = ux / p.up_x;
        if (ix >= p.in_w) continue;
        // true convolution: tap (ky,kx) of the sliding window meets the flipped kernel
        v += x[(((int64_t)m * p.in_h + iy) * p.in_w + ix) * p.minor + c] *
             __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx));
      }
    }
    if (bias) v = lrelu(v + __ldg(bias + c), p.alpha) * p.gain;
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------
// fast path: up = down = 1, kernel <= 4x4, minor % 4 == 0 (NHWC)
// ---------------------------------------------------------------------------------------
struct Taps4 { float k[4][4]; };

// Output epilogues of the blur kernels:
//   EPI 0: none
//   EPI 1: (v + bias[c]) -> lrelu(alpha) -> * gain                      (Blur -> FusedLeakyReLU, forward)
//   EPI 2: v * (ref > 0 ? 1 : alpha) * gain, per-thread channel sums    (backward of act -> Blur: the blurred
//          gradient meets the mask of the activation that fed the blur; the sums are the bias gradient)
template <int EPI>
__device__ __forceinline__ void blur_epilogue(float4& v, const float4& b4, const UpfirdnParams& p, const float4& r,
                                              float4& bsum) {
  if (EPI == 1) {
    v.x = lrelu(v.x + b4.x, p.alpha) * p.gain;
    v.y = lrelu(v.y + b4.y, p.alpha) * p.gain;
    v.z = lrelu(v.z + b4.z, p.alpha) * p.gain;
    v.w = lrelu(v.w + b4.w, p.alpha) * p.gain;
  } else if (EPI == 2) {
    v.x = (r.x > 0.f ? v.x : v.x * p.alpha) * p.gain;
    v.y = (r.y > 0.f ? v.y : v.y * p.alpha) * p.gain;
    v.z = (r.z > 0.f ? v.z : v.z * p.alpha) * p.gain;
    v.w = (r.w > 0.f ? v.w : v.w * p.alpha) * p.gain;
    bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
  } else if (EPI == 3) {
    // backward of  u -> d[n,c] * u -> Blur: v = blur^T(g); dot[n,c] += sum_p v * u (r = u), result = v * d (b4 = d)
    bsum.x = fmaf(v.x, r.x, bsum.x); bsum.y = fmaf(v.y, r.y, bsum.y);
    bsum.z = fmaf(v.z, r.z, bsum.z); bsum.w = fmaf(v.w, r.w, bsum.w);
    v.x *= b4.x; v.y *= b4.y; v.z *= b4.z; v.w *= b4.w;
  }
}  // already flipped and zero-extended: k[ky][kx] meets in[oy+ky-pad][ox+kx-pad]

__device__ __forceinline__ float4 fma4(const float4& a, float s, const float4& acc) {
  return make_float4(fmaf(a.x, s, acc.x), fmaf(a.y, s, acc.y), fmaf(a.z, s, acc.z), fmaf(a.w, s, acc.w));
}
__device__ __forceinline__ float4 mul4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// XT adjacent output columns per thread: a row of XT+3 input vectors feeds XT outputs, so the
// L1/LSU traffic per output drops from 4 loads to (XT+3)/XT (the kernel is L1-wavefront bound
// otherwise: at HBM rate 4 loads + 1 store per element would need ~115 B/clk/SM of the 128 available).
// general (non-separable) 4x4 taps: every input row is folded into three pending output rows
template <int ROWS, int XT, int EPI>
__device__ __forceinline__ void blur4_general(float* __restrict__ out, const float* __restrict__ x,
                                              const float* __restrict__ bias, const UpfirdnParams& p, const Taps4& tp,
                                              const float* __restrict__ ref, float4& bsum) {
  const bool sep = false;
  float kx[4], kyv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) kx[i] = kyv[i] = 0.f;
  const int c4n = p.minor >> 2;
  const int groups = (p.out_w + XT - 1) / XT;
  const int lin = blockIdx.x * blockDim.x + threadIdx.x;  // (column group, c4), c4 fastest
  if (lin >= groups * c4n) return;
  const int og = lin / c4n;
  const int c = (lin - og * c4n) << 2;
  const int ox0 = og * XT;
  const int m = blockIdx.z;
  const int oy0 = blockIdx.y * ROWS;
  const int oy1 = min(oy0 + ROWS, p.out_h);
  const int ix0 = ox0 - p.pad_x0;

  const float* xin = x + (int64_t)m * p.in_h * p.in_w * p.minor + c;
  float* o = out + (((int64_t)m * p.out_h) * p.out_w + ox0) * p.minor + c;
  const int64_t orow = (int64_t)p.out_w * p.minor;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI == 1) b4 = *reinterpret_cast<const float4*>(bias + c);
  if (EPI == 3) b4 = bias ? *reinterpret_cast<const float4*>(bias + (int64_t)m * p.minor + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float* refp = EPI >= 2 ? ref + (((int64_t)m * p.out_h) * p.out_w + ox0) * p.minor + c : nullptr;

  bool colok[XT + 3];
#pragma unroll
  for (int i = 0; i < XT + 3; ++i) colok[i] = (ix0 + i) >= 0 && (ix0 + i) < p.in_w;

  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s0[XT], s1[XT], s2[XT];
#pragma unroll
  for (int j = 0; j < XT; ++j) s0[j] = s1[j] = s2[j] = zero;
  // input rows oy0-pad .. oy1-1-pad+3 ; output row (iy + pad - 3) completes at input row iy
  const int iy_begin = oy0 - p.pad_y0;
  const int iy_end = oy1 - 1 - p.pad_y0 + 3;
  for (int iy = iy_begin; iy <= iy_end; ++iy) {
    float4 r[XT + 3];
    const bool rowok = iy >= 0 && iy < p.in_h;
    const float* rp = xin + ((int64_t)iy * p.in_w + ix0) * p.minor;
#pragma unroll
    for (int i = 0; i < XT + 3; ++i)
      r[i] = (rowok && colok[i]) ? __ldg(reinterpret_cast<const float4*>(rp + (int64_t)i * p.minor)) : zero;
    const int oy = iy + p.pad_y0 - 3;
#pragma unroll
    for (int j = 0; j < XT; ++j) {
      float4 h[4];
      if (sep) {
        // rank-1 kernel (every IDEAS blur: outer([1,3,3,1])): one horizontal pass, scaled per kernel row
        float4 a = mul4(r[j], kx[0]);
        a = fma4(r[j + 1], kx[1], a);
        a = fma4(r[j + 2], kx[2], a);
        a = fma4(r[j + 3], kx[3], a);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) h[ky] = mul4(a, kyv[ky]);
      } else {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          float4 a = mul4(r[j], tp.k[ky][0]);
          a = fma4(r[j + 1], tp.k[ky][1], a);
          a = fma4(r[j + 2], tp.k[ky][2], a);
          a = fma4(r[j + 3], tp.k[ky][3], a);
          h[ky] = a;
        }
      }
      float4 done = make_float4(s2[j].x + h[3].x, s2[j].y + h[3].y, s2[j].z + h[3].z, s2[j].w + h[3].w);
      s2[j] = make_float4(s1[j].x + h[2].x, s1[j].y + h[2].y, s1[j].z + h[2].z, s1[j].w + h[2].w);
      s1[j] = make_float4(s0[j].x + h[1].x, s0[j].y + h[1].y, s0[j].z + h[1].z, s0[j].w + h[1].w);
      s0[j] = h[0];
      if (oy >= oy0 && ox0 + j < p.out_w) {
        float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI >= 2) rv = ld_stream4(refp + (int64_t)oy * orow + (int64_t)j * p.minor);
        blur_epilogue<EPI>(done, b4, p, rv, bsum);
        st_stream4(o + (int64_t)oy * orow + (int64_t)j * p.minor, done);
        if (EPI == 1 && p.out2 != nullptr) {
          const float4 q = *reinterpret_cast<const float4*>(p.post + (int64_t)m * p.minor + c);
          st_stream4(p.out2 + (o - out) + (int64_t)oy * orow + (int64_t)j * p.minor,
                     make_float4(done.x * q.x, done.y * q.y, done.z * q.z, done.w * q.w));
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------
// separable fast path (every blur IDEAS runs: outer([1,3,3,1]) * gain).  Same thread mapping as
// above, but written for instruction count -- at HBM rate the kernel above is issue-bound (ncu:
// 67 % issue-active, profiles/prof_blur_r1.raw.csv):
//   * packed fp32x2 FMAs (fma.rn.f32x2): a float4 of channels is two instructions per tap;
//   * the vertical pass is a 4-row sliding window of horizontal results kept in registers
//     (slot rotation by a x4-unrolled row loop, no moves): 4 + 4 packed taps per output instead
//     of scattering every row into three pending sums;
//   * tiles that do not touch a border (all but the outer ring) skip every predicate.
// ---------------------------------------------------------------------------------------
struct f2x2 { unsigned long long lo, hi; };   // four channels as two packed fp32 pairs

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2x2 ldg_f2x2(const float* p) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return f2x2{pack2(v.x, v.y), pack2(v.z, v.w)};
}

template <int XT, int EPI, bool INTERIOR>
__device__ __forceinline__ void blur4sep_strip(float* __restrict__ o, const float* __restrict__ xin, const UpfirdnParams& p,
                                               const unsigned long long (&kx)[4], const unsigned long long (&ky)[4],
                                               int ox0, int ix0, int oy0, int oy1, const float4& b4,
                                               const float* __restrict__ refp, float4& bsum,
                                               float* __restrict__ o2 = nullptr, const float4& post4 = float4{1.f, 1.f, 1.f, 1.f}) {
  const int64_t orow = (int64_t)p.out_w * p.minor;
  const int64_t irow = (int64_t)p.in_w * p.minor;
  bool colok[XT + 3];
#pragma unroll
  for (int i = 0; i < XT + 3; ++i) colok[i] = INTERIOR || ((ix0 + i) >= 0 && (ix0 + i) < p.in_w);
  f2x2 w[4][XT];   // horizontal results of the last four input rows (slot = row & 3 after unrolling)
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < XT; ++j) w[s][j] = f2x2{0ull, 0ull};
  const int iy_begin = oy0 - p.pad_y0;
  const int iy_end = oy1 - 1 - p.pad_y0 + 3;
  const float* rp = xin + (int64_t)iy_begin * irow + (int64_t)ix0 * p.minor;
  float* op = o + (int64_t)(oy0 - 3) * orow;   // output row completed by input row iy is iy + pad - 3
  int oy = oy0 - 3;
  for (int iy = iy_begin; iy <= iy_end; iy += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (iy + u <= iy_end) {
        const bool rowok = INTERIOR || ((iy + u) >= 0 && (iy + u) < p.in_h);
        f2x2 r[XT + 3];
#pragma unroll
        for (int i = 0; i < XT + 3; ++i)
          r[i] = (rowok && colok[i]) ? ldg_f2x2(rp + (int64_t)i * p.minor) : f2x2{0ull, 0ull};
        // EPI 2: the mask operand of the row this iteration completes is fetched together with the input row, so it
        // is in flight during the arithmetic instead of being a dependent load in the epilogue
        float4 refv[XT];
#pragma unroll
        for (int j = 0; j < XT; ++j) {
          refv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI >= 2 && oy >= oy0 && (INTERIOR || ox0 + j < p.out_w)) refv[j] = ld_stream4(refp + (op - o) + (int64_t)j * p.minor);
        }
#pragma unroll
        for (int j = 0; j < XT; ++j) {
          f2x2 a;
          a.lo = fma2(r[j + 3].lo, kx[3], fma2(r[j + 2].lo, kx[2], fma2(r[j + 1].lo, kx[1], mul2(r[j].lo, kx[0]))));
          a.hi = fma2(r[j + 3].hi, kx[3], fma2(r[j + 2].hi, kx[2], fma2(r[j + 1].hi, kx[1], mul2(r[j].hi, kx[0]))));
          w[u][j] = a;   // slot u holds row iy+u; slots (u+1)&3, (u+2)&3, (u+3)&3 hold rows -3, -2, -1
          if (oy >= oy0 && (INTERIOR || ox0 + j < p.out_w)) {
            const f2x2& w0 = w[(u + 1) & 3][j];
            const f2x2& w1 = w[(u + 2) & 3][j];
            const f2x2& w2 = w[(u + 3) & 3][j];
            const unsigned long long dl = fma2(a.lo, ky[3], fma2(w2.lo, ky[2], fma2(w1.lo, ky[1], mul2(w0.lo, ky[0]))));
            const unsigned long long dh = fma2(a.hi, ky[3], fma2(w2.hi, ky[2], fma2(w1.hi, ky[1], mul2(w0.hi, ky[0]))));
            float4 done;
            unpack2(dl, done.x, done.y);
            unpack2(dh, done.z, done.w);
            blur_epilogue<EPI>(done, b4, p, refv[j], bsum);
            st_stream4(op + (int64_t)j * p.minor, done);
            if (EPI == 1 && o2 != nullptr)
              st_stream4(o2 + (op - o) + (int64_t)j * p.minor,
                         make_float4(done.x * post4.x, done.y * post4.y, done.z * post4.z, done.w * post4.w));
          }
        }
        rp += irow;
        op += orow;
        ++oy;
      }
    }
  }
}

template <int ROWS, int XT, int EPI>
__global__ void __launch_bounds__(128) blur4_nhwc_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                         const float* __restrict__ kernel,
                                                         const float* __restrict__ bias, UpfirdnParams p,
                                                         const float* __restrict__ ref, float* __restrict__ gbias) {
  __shared__ float4 red[EPI >= 2 ? 128 : 1];
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  // taps: flipped (true convolution) and zero-extended to 4x4; 16 uniform loads per thread
  Taps4 tp;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      tp.k[a][b] = (a < p.kh && b < p.kw) ? __ldg(kernel + (p.kh - 1 - a) * p.kw + (p.kw - 1 - b)) : 0.f;
  // separable? (exact test: the binomial taps are small integers / 2^k, so rank 1 holds bit-exactly)
  bool sep = tp.k[0][0] != 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) sep = sep && (tp.k[a][b] * tp.k[0][0] == tp.k[a][0] * tp.k[0][b]);
  const int c4n = p.minor >> 2;
  const int groups = (p.out_w + XT - 1) / XT;
  const int lin = blockIdx.x * blockDim.x + threadIdx.x;  // (column group, c4), c4 fastest
  const bool active = lin < groups * c4n;
  if (!sep) {
    blur4_general<ROWS, XT, EPI>(out, x, bias, p, tp, ref, bsum);
  } else if (active) {
    unsigned long long kx[4], ky[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float cv = tp.k[i][0] / tp.k[0][0];
      kx[i] = pack2(tp.k[0][i], tp.k[0][i]);
      ky[i] = pack2(cv, cv);
    }
    const int og = lin / c4n;
    const int c = (lin - og * c4n) << 2;
    const int ox0 = og * XT;
    const int m = blockIdx.z;
    const int oy0 = blockIdx.y * ROWS;
    const int oy1 = min(oy0 + ROWS, p.out_h);
    const int ix0 = ox0 - p.pad_x0;
    const float* xin = x + (int64_t)m * p.in_h * p.in_w * p.minor + c;
    const int64_t obase = (((int64_t)m * p.out_h) * p.out_w + ox0) * p.minor + c;
    float* o = out + obase;
    const float* refp = EPI >= 2 ? ref + obase : nullptr;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == 1) b4 = *reinterpret_cast<const float4*>(bias + c);
    if (EPI == 3) b4 = bias ? *reinterpret_cast<const float4*>(bias + (int64_t)m * p.minor + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const bool interior = ix0 >= 0 && ix0 + XT + 3 <= p.in_w && ox0 + XT <= p.out_w && oy0 - p.pad_y0 >= 0 &&
                          oy1 - 1 - p.pad_y0 + 3 < p.in_h;
    float* o2 = (EPI == 1 && p.out2) ? p.out2 + obase : nullptr;
    float4 post4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (o2) post4 = *reinterpret_cast<const float4*>(p.post + (int64_t)m * p.minor + c);
    if (interior) blur4sep_strip<XT, EPI, true>(o, xin, p, kx, ky, ox0, ix0, oy0, oy1, b4, refp, bsum, o2, post4);
    else blur4sep_strip<XT, EPI, false>(o, xin, p, kx, ky, ox0, ix0, oy0, oy1, b4, refp, bsum, o2, post4);
  }
  if (EPI >= 2 && gbias != nullptr) {
    if (EPI == 3) gbias += (int64_t)blockIdx.z * p.minor;      // per-(image, channel) sums
    // bias gradient: fold the replicas of each channel group inside the CTA (blockDim is a multiple of c4n or
    // c4n a multiple of blockDim: thread t always owns channel group (blockIdx.x*128 + t) % c4n), then one
    // atomic per channel per CTA
    red[threadIdx.x] = bsum;
    __syncthreads();
    const int first = (int)((int64_t)blockIdx.x * blockDim.x % c4n);
    const int span = c4n < (int)blockDim.x ? c4n : (int)blockDim.x;   // distinct channel groups in this CTA
    if ((int)threadIdx.x < span) {
      float4 sacc = red[threadIdx.x];
      for (int t = threadIdx.x + c4n; t < (int)blockDim.x; t += c4n) {
        const float4 q = red[t];
        sacc.x += q.x; sacc.y += q.y; sacc.z += q.z; sacc.w += q.w;
      }
      float* dst = gbias + (((first + threadIdx.x) % c4n) << 2);
      atomicAdd(dst + 0, sacc.x); atomicAdd(dst + 1, sacc.y); atomicAdd(dst + 2, sacc.z); atomicAdd(dst + 3, sacc.w);
    }
  }
}

// ---------------------------------------------------------------------------------------
// resampling fast paths (separable taps, <= 4x4, NHWC, minor % 4 == 0):
//   down = 2: the skip branch of a down-sampling ResBlock is Blur -> 1x1 stride-2 conv (reference
//     models.py:68-74,213-227); only every other blurred pixel is ever read, so the host side runs
//     upfirdn2d(down = 2) -> 1x1 stride-1 conv instead (identical arithmetic): the blur writes a quarter
//     of the pixels and the conv reads a quarter.
//   up = 2: its adjoint, and the forward of the up-sampling skip (1x1 transposed conv stride 2 -> Blur
//     == 1x1 conv at low resolution -> upfirdn2d(up = 2), models.py:78-95): each output pixel meets only
//     the 2x2 taps whose parity matches the zero-inserted grid, and the zeros are never materialised.
// Both keep the thread mapping of the blur kernel (4 channels x a few columns, sliding down a strip).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool taps_separable(const float* __restrict__ kernel, const UpfirdnParams& p, float (&rowk)[4],
                                               float (&colk)[4]) {
  Taps4 tp;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      tp.k[a][b] = (a < p.kh && b < p.kw) ? __ldg(kernel + (p.kh - 1 - a) * p.kw + (p.kw - 1 - b)) : 0.f;
  bool sep = tp.k[0][0] != 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) sep = sep && (tp.k[a][b] * tp.k[0][0] == tp.k[a][0] * tp.k[0][b]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    rowk[i] = tp.k[0][i];
    colk[i] = tp.k[i][0] / tp.k[0][0];
  }
  return sep;
}

// direct form for one (pixel, 4 channels): used only when the taps are not rank 1
__device__ float4 upfirdn_direct4(const float* __restrict__ x, const float* __restrict__ kernel, const UpfirdnParams& p, int m,
                                  int oy, int ox, int c) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ky = 0; ky < p.kh; ++ky) {
    const int uy = oy * p.down_y + ky - p.pad_y0;
    if (uy < 0 || uy % p.up_y) continue;
    const int iy = uy / p.up_y;
    if (iy >= p.in_h) continue;
    for (int kx = 0; kx < p.kw; ++kx) {
      const int ux = ox * p.down_x + kx - p.pad_x0;
      if (ux < 0 || ux % p.up_x) continue;
      const int ix = ux / p.up_x;
      if (ix >= p.in_w) continue;
      const float w = __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx));
      const float4 a = __ldg(reinterpret_cast<const float4*>(x + (((int64_t)m * p.in_h + iy) * p.in_w + ix) * p.minor + c));
      v.x = fmaf(a.x, w, v.x); v.y = fmaf(a.y, w, v.y); v.z = fmaf(a.z, w, v.z); v.w = fmaf(a.w, w, v.w);
    }
  }
  return v;
}

__device__ __forceinline__ f2x2 ldg_f2x2_guard(const float* p, bool ok) {
  return ok ? ldg_f2x2(p) : f2x2{0ull, 0ull};
}
__device__ __forceinline__ void st_f2x2(float* p, const f2x2& v) {
  float4 o;
  unpack2(v.lo, o.x, o.y);
  unpack2(v.hi, o.z, o.w);
  st_stream4(p, o);
}

// out[oy][ox] = sum_{ky,kx} k[ky][kx] * in[2*oy + ky - pad_y0][2*ox + kx - pad_x0]
template <int ROWS, int XT>
__global__ void __launch_bounds__(128) blur4_down2_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                          const float* __restrict__ kernel, UpfirdnParams p) {
  float rowk[4], colk[4];
  const bool sep = taps_separable(kernel, p, rowk, colk);
  const int c4n = p.minor >> 2;
  const int groups = (p.out_w + XT - 1) / XT;
  const int lin = blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= groups * c4n) return;
  const int og = lin / c4n;
  const int c = (lin - og * c4n) << 2;
  const int ox0 = og * XT;
  const int m = blockIdx.z;
  const int oy0 = blockIdx.y * ROWS;
  const int oy1 = min(oy0 + ROWS, p.out_h);
  float* o = out + (((int64_t)m * p.out_h) * p.out_w + ox0) * p.minor + c;
  const int64_t orow = (int64_t)p.out_w * p.minor;
  if (!sep) {
    for (int oy = oy0; oy < oy1; ++oy)
      for (int j = 0; j < XT && ox0 + j < p.out_w; ++j)
        st_stream4(o + (int64_t)oy * orow + (int64_t)j * p.minor, upfirdn_direct4(x, kernel, p, m, oy, ox0 + j, c));
    return;
  }
  unsigned long long kx[4], ky[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    kx[i] = pack2(rowk[i], rowk[i]);
    ky[i] = pack2(colk[i], colk[i]);
  }
  constexpr int NL = 2 * XT + 2;          // input columns feeding XT outputs
  const int ix0 = 2 * ox0 - p.pad_x0;
  bool colok[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) colok[i] = (ix0 + i) >= 0 && (ix0 + i) < p.in_w;
  const float* xin = x + (int64_t)m * p.in_h * p.in_w * p.minor + c;
  const int64_t irow = (int64_t)p.in_w * p.minor;
  f2x2 w[4][XT];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < XT; ++j) w[s][j] = f2x2{0ull, 0ull};
  const int iy_begin = 2 * oy0 - p.pad_y0;
  const int iy_end = 2 * (oy1 - 1) - p.pad_y0 + 3;
  const float* rp = xin + (int64_t)iy_begin * irow + (int64_t)ix0 * p.minor;
  float* op = o + (int64_t)oy0 * orow;
  for (int iy = iy_begin; iy <= iy_end; iy += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (iy + u <= iy_end) {
        const bool rowok = (iy + u) >= 0 && (iy + u) < p.in_h;
        f2x2 r[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) r[i] = ldg_f2x2_guard(rp + (int64_t)i * p.minor, rowok && colok[i]);
        // relative row (iy + u - iy_begin) is odd and >= 3  <=>  an output row completes here
        const bool emit = (u & 1) && (iy + u - iy_begin) >= 3;
#pragma unroll
        for (int j = 0; j < XT; ++j) {
          f2x2 a;
          a.lo = fma2(r[2 * j + 3].lo, kx[3], fma2(r[2 * j + 2].lo, kx[2], fma2(r[2 * j + 1].lo, kx[1], mul2(r[2 * j].lo, kx[0]))));
          a.hi = fma2(r[2 * j + 3].hi, kx[3], fma2(r[2 * j + 2].hi, kx[2], fma2(r[2 * j + 1].hi, kx[1], mul2(r[2 * j].hi, kx[0]))));
          w[u][j] = a;
          if (emit && ox0 + j < p.out_w) {
            const f2x2& w0 = w[(u + 1) & 3][j];
            const f2x2& w1 = w[(u + 2) & 3][j];
            const f2x2& w2 = w[(u + 3) & 3][j];
            f2x2 d;
            d.lo = fma2(a.lo, ky[3], fma2(w2.lo, ky[2], fma2(w1.lo, ky[1], mul2(w0.lo, ky[0]))));
            d.hi = fma2(a.hi, ky[3], fma2(w2.hi, ky[2], fma2(w1.hi, ky[1], mul2(w0.hi, ky[0]))));
            st_f2x2(op + (int64_t)j * p.minor, d);
          }
        }
        rp += irow;
        if (emit) op += orow;
      }
    }
  }
}

// out[oy][ox] = sum_{ky,kx} k[ky][kx] * z[oy + ky - pad_y0][ox + kx - pad_x0],  z[2i][2j] = in[i][j], 0 elsewhere.
// With s = ox - pad_x0: the two live taps are kx = (s & 1), (s & 1) + 2 and meet in[ja], in[ja + 1], ja = (s + 1) >> 1
// (same along y).  PX = parity of pad_x0 (compile time, so the per-thread register arrays are indexed statically).
template <int ROWS, int XT, int PX>
__global__ void __launch_bounds__(128) blur4_up2_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                        const float* __restrict__ kernel, UpfirdnParams p) {
  float rowk[4], colk[4];
  const bool sep = taps_separable(kernel, p, rowk, colk);
  constexpr int NO = 2 * XT;              // output columns per thread
  const int c4n = p.minor >> 2;
  const int groups = (p.out_w + NO - 1) / NO;
  const int lin = blockIdx.x * blockDim.x + threadIdx.x;
  if (lin >= groups * c4n) return;
  const int og = lin / c4n;
  const int c = (lin - og * c4n) << 2;
  const int ox0 = og * NO;
  const int m = blockIdx.z;
  const int oy0 = blockIdx.y * ROWS;
  const int oy1 = min(oy0 + ROWS, p.out_h);
  float* o = out + (((int64_t)m * p.out_h) * p.out_w + ox0) * p.minor + c;
  const int64_t orow = (int64_t)p.out_w * p.minor;
  if (!sep) {
    for (int oy = oy0; oy < oy1; ++oy)
      for (int j = 0; j < NO && ox0 + j < p.out_w; ++j) {
        float4 v = upfirdn_direct4(x, kernel, p, m, oy, ox0 + j, c);
        float* dp = o + (int64_t)oy * orow + (int64_t)j * p.minor;
        if (p.residual) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(p.residual + (dp - out)));
          v.x = (v.x + q.x) * p.res_scale; v.y = (v.y + q.y) * p.res_scale;
          v.z = (v.z + q.z) * p.res_scale; v.w = (v.w + q.w) * p.res_scale;
        }
        st_stream4(dp, v);
      }
    return;
  }
  unsigned long long kx[4], ky[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    kx[i] = pack2(rowk[i], rowk[i]);
    ky[i] = pack2(colk[i], colk[i]);
  }
  // columns: s0 = ox0 - pad_x0 has parity PX (ox0 is even); first input column ja0 = (s0 + 1) >> 1
  const int s0 = ox0 - p.pad_x0;
  const int ja0 = (s0 + 1) >> 1;
  constexpr int NL = XT + 2;
  bool colok[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) colok[i] = (ja0 + i) >= 0 && (ja0 + i) < p.in_w;
  const float* xin = x + (int64_t)m * p.in_h * p.in_w * p.minor + c;
  const int64_t irow = (int64_t)p.in_w * p.minor;
  // rows: output row oy (t = oy - pad_y0) meets input rows ia, ia + 1 with ia = (t + 1) >> 1 and taps (t & 1), (t & 1) + 2
  const int ia_begin = (oy0 - p.pad_y0 + 1) >> 1;
  const int ia_end = ((oy1 - 1 - p.pad_y0 + 1) >> 1) + 1;
  f2x2 hprev[NO], hcur[NO];
#pragma unroll
  for (int j = 0; j < NO; ++j) hprev[j] = hcur[j] = f2x2{0ull, 0ull};
  const float* rp = xin + (int64_t)ia_begin * irow + (int64_t)ja0 * p.minor;
  for (int ia = ia_begin; ia <= ia_end; ++ia) {
    const bool rowok = ia >= 0 && ia < p.in_h;
    f2x2 r[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) r[i] = ldg_f2x2_guard(rp + (int64_t)i * p.minor, rowok && colok[i]);
#pragma unroll
    for (int j = 0; j < NO; ++j) {
      // s = s0 + j: parity (PX + j) & 1; input offset ((s + 1) >> 1) - ja0
      const int par = (PX + j) & 1;
      const int off = PX ? j / 2 : j / 2 + (j & 1);   // ((s0 + j + 1) >> 1) - ((s0 + 1) >> 1)
      hcur[j].lo = fma2(r[off + 1].lo, kx[par + 2], mul2(r[off].lo, kx[par]));
      hcur[j].hi = fma2(r[off + 1].hi, kx[par + 2], mul2(r[off].hi, kx[par]));
    }
    // rows (ia - 1, ia) feed output rows t = 2*(ia - 1) - 1 + ... : t with (t + 1) >> 1 == ia - 1  ->  t = 2*ia - 3, 2*ia - 2
    if (ia > ia_begin) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int t = 2 * ia - 3 + q;
        const int oy = t + p.pad_y0;
        if (oy >= oy0 && oy < oy1) {
          const int par = t & 1;
          const unsigned long long k0 = par ? ky[1] : ky[0], k2 = par ? ky[3] : ky[2];
          float* op = o + (int64_t)oy * orow;
#pragma unroll
          for (int j = 0; j < NO; ++j) {
            if (ox0 + j < p.out_w) {
              f2x2 d;
              d.lo = fma2(hcur[j].lo, k2, mul2(hprev[j].lo, k0));
              d.hi = fma2(hcur[j].hi, k2, mul2(hprev[j].hi, k0));
              if (p.residual) {
                const unsigned long long rs2 = pack2(p.res_scale, p.res_scale);
                const f2x2 q = ldg_f2x2(p.residual + (op - out) + (int64_t)j * p.minor);
                d.lo = mul2(fma2(q.lo, pack2(1.f, 1.f), d.lo), rs2);
                d.hi = mul2(fma2(q.hi, pack2(1.f, 1.f), d.hi), rs2);
              }
              st_f2x2(op + (int64_t)j * p.minor, d);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NO; ++j) hprev[j] = hcur[j];
    rp += irow;
  }
}

}  // namespace ideas

using namespace ideas;

extern "C" int ideas_upfirdn2d(float* out, const float* x, const float* kernel, int major, int in_h, int in_w,
                               int minor, int kernel_h, int kernel_w, int up_x, int up_y, int down_x, int down_y,
                               int pad_x0, int pad_x1, int pad_y0, int pad_y1, const float* bias, float alpha,
                               float gain, void* stream) {
  return ideas_upfirdn2d_res(out, x, kernel, major, in_h, in_w, minor, kernel_h, kernel_w, up_x, up_y, down_x, down_y, pad_x0,
                             pad_x1, pad_y0, pad_y1, bias, alpha, gain, nullptr, 1.f, stream);
}

extern "C" int ideas_upfirdn2d_res(float* out, const float* x, const float* kernel, int major, int in_h, int in_w,
                                   int minor, int kernel_h, int kernel_w, int up_x, int up_y, int down_x, int down_y,
                                   int pad_x0, int pad_x1, int pad_y0, int pad_y1, const float* bias, float alpha,
                                   float gain, const float* residual, float res_scale, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IDEAS_REQUIRE(major >= 0 && in_h >= 0 && in_w >= 0 && minor >= 1, "upfirdn2d: bad input shape");
  IDEAS_REQUIRE(kernel_h >= 1 && kernel_w >= 1, "upfirdn2d: empty FIR kernel");
  IDEAS_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down factors must be >= 1");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kernel_h; p.kw = kernel_w;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  // upfirdn2d_kernel.cu:236-239
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h + down_y) / down_y;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w + down_x) / down_x;
  p.alpha = alpha; p.gain = gain;
  p.residual = residual; p.res_scale = res_scale;
  p.out2 = nullptr; p.post = nullptr;
  const bool empty = major == 0 || in_h == 0 || in_w == 0;
  IDEAS_REQUIRE(empty || ((in_h * up_y + pad_y0 + pad_y1 - kernel_h) >= 0 && (in_w * up_x + pad_x0 + pad_x1 - kernel_w) >= 0),
                "upfirdn2d: padding/cropping leaves no output (in %dx%d, kernel %dx%d)", in_h, in_w, kernel_h, kernel_w);
  if (empty) return IDEAS_OK;
  const int64_t total = (int64_t)major * p.out_h * p.out_w * minor;
  if (total == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && x && kernel, "upfirdn2d: null pointer");

  if (residual) {      // only the fused up-sampling path merges a residual; callers fall back to a separate merge otherwise
    const bool ok = up_x == 2 && up_y == 2 && down_x == 1 && down_y == 1 && kernel_h <= 4 && kernel_w <= 4 && minor % 4 == 0 &&
                    aligned16(out) && aligned16(x) && aligned16(residual) && !bias && major <= 65535;
    if (!ok) {
      set_error("upfirdn2d: a residual operand needs the up=2 fast path (<= 4x4 kernel, channels %% 4 == 0, no bias)");
      return IDEAS_ERR_UNSUPPORTED;
    }
  }
  const bool fast = up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kernel_h <= 4 && kernel_w <= 4 &&
                    minor % 4 == 0 && aligned16(out) && aligned16(x) && (!bias || aligned16(bias)) && major <= 65535;
  if (fast) {
    const int c4n = minor / 4;
    const int variant = ideas::blur_variant();
    if (!bias && variant != 0) {                    // tuning variants (option "blur_variant"), plain blur only
      auto launch = [&](auto rows, auto xt) {
        constexpr int R = decltype(rows)::value, X = decltype(xt)::value;
        dim3 grid(ceil_div(ceil_div(p.out_w, X) * c4n, 128), ceil_div(p.out_h, R), major);
        blur4_nhwc_kernel<R, X, 0><<<grid, 128, 0, st>>>(out, x, kernel, bias, p, nullptr, nullptr);
      };
      if (variant == 1) launch(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
      else if (variant == 2) launch(std::integral_constant<int, 64>{}, std::integral_constant<int, 4>{});
      else if (variant == 4) launch(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
      else if (variant == 5) launch(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
      else if (variant == 6) launch(std::integral_constant<int, 16>{}, std::integral_constant<int, 2>{});
      else if (variant == 7) launch(std::integral_constant<int, 64>{}, std::integral_constant<int, 1>{});
      else launch(std::integral_constant<int, 64>{}, std::integral_constant<int, 2>{});
      IDEAS_CHECK_LAUNCH("upfirdn2d(fast)");
      return IDEAS_OK;
    }
    // 32 rows x 2 columns per thread: scripts/blur_variants.py (x4 columns: 1-17 % slower, 127 registers)
    constexpr int ROWS = 32, XT = 2;
    dim3 grid(ceil_div(ceil_div(p.out_w, XT) * c4n, 128), ceil_div(p.out_h, ROWS), major);
    if (bias) blur4_nhwc_kernel<ROWS, XT, 1><<<grid, 128, 0, st>>>(out, x, kernel, bias, p, nullptr, nullptr);
    else blur4_nhwc_kernel<ROWS, XT, 0><<<grid, 128, 0, st>>>(out, x, kernel, bias, p, nullptr, nullptr);
    IDEAS_CHECK_LAUNCH("upfirdn2d(fast)");
    return IDEAS_OK;
  }
  const bool resample_ok = kernel_h <= 4 && kernel_w <= 4 && minor % 4 == 0 && aligned16(out) && aligned16(x) && !bias &&
                           major <= 65535;
  if (resample_ok && up_x == 1 && up_y == 1 && down_x == 2 && down_y == 2) {
#define IDEAS_DOWN2(ROWS, XT)                                                                              \
  {                                                                                                       \
    dim3 grid(ceil_div(ceil_div(p.out_w, XT) * (minor / 4), 128), ceil_div(p.out_h, ROWS), major);        \
    blur4_down2_kernel<ROWS, XT><<<grid, 128, 0, st>>>(out, x, kernel, p);                                \
  }
    switch (g_resample_variant.load()) {
      case 1: IDEAS_DOWN2(32, 2) break;
      case 2: IDEAS_DOWN2(8, 2) break;
      case 3: IDEAS_DOWN2(16, 1) break;
      case 4: IDEAS_DOWN2(32, 1) break;
      case 5: IDEAS_DOWN2(16, 2) break;        // the default until scripts/bench_resample.py: 1.3-2.6x slower
      case 6: IDEAS_DOWN2(4, 1) break;
      default: IDEAS_DOWN2(8, 1) break;         // 8 output rows x 1 column per thread: most threads in flight
    }
#undef IDEAS_DOWN2
    IDEAS_CHECK_LAUNCH("upfirdn2d(down2)");
    return IDEAS_OK;
  }
  if (resample_ok && up_x == 2 && up_y == 2 && down_x == 1 && down_y == 1) {
#define IDEAS_UP2(ROWS, XT)                                                                                \
  {                                                                                                       \
    dim3 grid(ceil_div(ceil_div(p.out_w, 2 * XT) * (minor / 4), 128), ceil_div(p.out_h, ROWS), major);    \
    if (pad_x0 & 1) blur4_up2_kernel<ROWS, XT, 1><<<grid, 128, 0, st>>>(out, x, kernel, p);               \
    else blur4_up2_kernel<ROWS, XT, 0><<<grid, 128, 0, st>>>(out, x, kernel, p);                          \
  }
    switch (g_resample_variant.load()) {
      case 1: IDEAS_UP2(16, 2) break;
      case 2: IDEAS_UP2(64, 2) break;
      case 3: IDEAS_UP2(32, 1) break;
      case 4: IDEAS_UP2(32, 2) break;          // the default until scripts/bench_resample.py: 6-12 % slower
      case 5: IDEAS_UP2(8, 1) break;
      case 6: IDEAS_UP2(64, 1) break;
      default: IDEAS_UP2(16, 1) break;          // 16 output rows x 2 columns per thread
    }
#undef IDEAS_UP2
    IDEAS_CHECK_LAUNCH("upfirdn2d(up2)");
    return IDEAS_OK;
  }
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  upfirdn2d_generic_kernel<<<(int)blocks, 256, 0, st>>>(out, x, kernel, bias, p, total);
  IDEAS_CHECK_LAUNCH("upfirdn2d(generic)");
  return IDEAS_OK;
}

// Backward of  y = act(u) -> z = blur(y)  in one pass:  gx = blur^T(gz) * (ref > 0 ? 1 : alpha) * gain, where blur^T is
// upfirdn2d with the caller-supplied (already flipped) kernel and gradient pads, `ref` = y (same shape as gx) and
// gbias[c] += sum over everything but the channel of gx (may be NULL).  Replaces UpFirDn2dBackward followed by
// FusedLeakyReLUFunctionBackward (upfirdn2d.py:19-42 + fused_act.py:20-49): the blurred gradient is never written.
extern "C" int ideas_blur_act_backward(float* gx, float* gbias, const float* g, const float* ref, const float* kernel,
                                       int major, int in_h, int in_w, int minor, int kernel_h, int kernel_w,
                                       int pad_x0, int pad_x1, int pad_y0, int pad_y1, float alpha, float gain,
                                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IDEAS_REQUIRE(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1, "blur_act_backward: bad input shape");
  IDEAS_REQUIRE(kernel_h >= 1 && kernel_w >= 1, "blur_act_backward: empty FIR kernel");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kernel_h; p.kw = kernel_w;
  p.up_x = p.up_y = p.down_x = p.down_y = 1; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = in_h + pad_y0 + pad_y1 - kernel_h + 1;
  p.out_w = in_w + pad_x0 + pad_x1 - kernel_w + 1;
  p.alpha = alpha; p.gain = gain;
  p.residual = nullptr; p.res_scale = 1.f;
  p.out2 = nullptr; p.post = nullptr;
  IDEAS_REQUIRE(p.out_h >= 1 && p.out_w >= 1, "blur_act_backward: padding/cropping leaves no output");
  if (major == 0) return IDEAS_OK;
  IDEAS_REQUIRE(gx && g && ref && kernel, "blur_act_backward: null pointer");
  const int c4n = minor / 4;
  const bool fast = kernel_h <= 4 && kernel_w <= 4 && minor % 4 == 0 && aligned16(gx) && aligned16(g) && aligned16(ref) &&
                    major <= 65535 && (c4n <= 128 ? 128 % c4n == 0 : c4n % 128 == 0);
  if (!fast) {
    set_error("blur_act_backward: needs a <= 4x4 kernel, channels %% 4 == 0 and a power-of-two channel count (C=%d)", minor);
    return IDEAS_ERR_UNSUPPORTED;
  }
  constexpr int ROWS = 32, XT = 2;
  dim3 grid(ceil_div(ceil_div(p.out_w, XT) * c4n, 128), ceil_div(p.out_h, ROWS), major);
  blur4_nhwc_kernel<ROWS, XT, 2><<<grid, 128, 0, st>>>(gx, g, kernel, nullptr, p, ref, gbias);
  IDEAS_CHECK_LAUNCH("blur_act_backward");
  return IDEAS_OK;
}

// Backward of  u -> d[n,c] * u -> Blur  (the demodulated, up-sampled branch of ModulatedConv2d, stylegan2/model.py:250-261)
// in one pass over the incoming gradient g:  v = blur^T(g)  (caller-supplied flipped kernel and gradient pads),
//   gx[n,p,c] = v * scale[n,c]        (scale = d, may be NULL = 1)
//   dot[n,c] += sum_p v[n,p,c] * ref[n,p,c]        (ref = u, shape of gx; dot zero-initialised by the caller)
// so that dL/dd = dot / d.  Replaces UpFirDn2dBackward followed by a channel_dot pass: the blurred gradient is never written.
extern "C" int ideas_blur_scale_dot_backward(float* gx, float* dot, const float* g, const float* ref, const float* scale,
                                             const float* kernel, int major, int in_h, int in_w, int minor,
                                             int kernel_h, int kernel_w, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                                             void* stream) {
  const float alpha = 0.f, gain = 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  IDEAS_REQUIRE(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1, "blur_scale_dot_backward: bad input shape");
  IDEAS_REQUIRE(kernel_h >= 1 && kernel_w >= 1, "blur_scale_dot_backward: empty FIR kernel");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kernel_h; p.kw = kernel_w;
  p.up_x = p.up_y = p.down_x = p.down_y = 1; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = in_h + pad_y0 + pad_y1 - kernel_h + 1;
  p.out_w = in_w + pad_x0 + pad_x1 - kernel_w + 1;
  p.alpha = alpha; p.gain = gain;
  p.residual = nullptr; p.res_scale = 1.f;
  p.out2 = nullptr; p.post = nullptr;
  IDEAS_REQUIRE(p.out_h >= 1 && p.out_w >= 1, "blur_scale_dot_backward: padding/cropping leaves no output");
  if (major == 0) return IDEAS_OK;
  IDEAS_REQUIRE(gx && g && ref && kernel, "blur_scale_dot_backward: null pointer");
  const int c4n = minor / 4;
  const bool fast = kernel_h <= 4 && kernel_w <= 4 && minor % 4 == 0 && aligned16(gx) && aligned16(g) && aligned16(ref) && (!scale || aligned16(scale)) &&
                    major <= 65535 && (c4n <= 128 ? 128 % c4n == 0 : c4n % 128 == 0);
  if (!fast) {
    set_error("blur_scale_dot_backward: needs a <= 4x4 kernel, channels %% 4 == 0 and a power-of-two channel count (C=%d)", minor);
    return IDEAS_ERR_UNSUPPORTED;
  }
  constexpr int ROWS = 32, XT = 2;
  dim3 grid(ceil_div(ceil_div(p.out_w, XT) * c4n, 128), ceil_div(p.out_h, ROWS), major);
  blur4_nhwc_kernel<ROWS, XT, 3><<<grid, 128, 0, st>>>(gx, g, kernel, scale, p, ref, dot);
  IDEAS_CHECK_LAUNCH("blur_scale_dot_backward");
  return IDEAS_OK;
}

// Blur -> FusedLeakyReLU of an up-sampling StyledConv (stylegan2/model.py:261,375) with a second output for the NEXT
// layer: out = gain * lrelu(blur(x) + bias[c]),  out2 = out * post[n,c]  (that layer's style modulation, so it needs
// no modulation pass of its own).  up = down = 1 fast path only (<= 4x4 kernel, channels % 4 == 0).
extern "C" int ideas_blur_bias_act_post(float* out, float* out2, const float* x, const float* kernel, const float* bias,
                                        const float* post, int major, int in_h, int in_w, int minor, int kernel_h,
                                        int kernel_w, int pad_x0, int pad_x1, int pad_y0, int pad_y1, float alpha,
                                        float gain, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IDEAS_REQUIRE(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1, "blur_bias_act_post: bad input shape");
  IDEAS_REQUIRE(kernel_h >= 1 && kernel_w >= 1, "blur_bias_act_post: empty FIR kernel");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kernel_h; p.kw = kernel_w;
  p.up_x = p.up_y = p.down_x = p.down_y = 1; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = in_h + pad_y0 + pad_y1 - kernel_h + 1;
  p.out_w = in_w + pad_x0 + pad_x1 - kernel_w + 1;
  p.alpha = alpha; p.gain = gain;
  p.residual = nullptr; p.res_scale = 1.f;
  p.out2 = out2; p.post = post;
  IDEAS_REQUIRE(p.out_h >= 1 && p.out_w >= 1, "blur_bias_act_post: padding/cropping leaves no output");
  if (major == 0) return IDEAS_OK;
  IDEAS_REQUIRE(out && out2 && x && kernel && bias && post, "blur_bias_act_post: null pointer");
  const bool fast = kernel_h <= 4 && kernel_w <= 4 && minor % 4 == 0 && aligned16(out) && aligned16(out2) && aligned16(x) &&
                    aligned16(bias) && aligned16(post) && major <= 65535;
  if (!fast) {
    set_error("blur_bias_act_post: needs a <= 4x4 kernel and channels %% 4 == 0 (C=%d)", minor);
    return IDEAS_ERR_UNSUPPORTED;
  }
  constexpr int ROWS = 32, XT = 2;
  dim3 grid(ceil_div(ceil_div(p.out_w, XT) * (minor / 4), 128), ceil_div(p.out_h, ROWS), major);
  blur4_nhwc_kernel<ROWS, XT, 1><<<grid, 128, 0, st>>>(out, x, kernel, bias, p, nullptr, nullptr);
  IDEAS_CHECK_LAUNCH("blur_bias_act_post");
  return IDEAS_OK;
}
