// Geometry of one implicit-GEMM convolution launch, shared by the SIMT and tcgen05 paths.
//
// Every convolution variant of the IDEAS path (same-resolution 3x3, 1x1, stride-2 "valid"
// convs after a Blur, 2x2 valid, stride-2 transposed convs, and all of their data/weight
// gradients) reduces to at most stride^2 launches of ONE form:
//
//   dst[n, qy*o_s + o_py, qx*o_s + o_px, k] =
//       epilogue( sum_{t < ntaps} sum_c  W[widx_t][k][c] * src[n, qy*i_s + dy_t, qx*i_s + dx_t, c] )
//
// for q in a QH x QW grid, out-of-range source pixels reading as zero.  A forward conv has
// i_s = stride, o_s = 1; a data gradient / transposed conv has i_s = 1, o_s = stride and one
// launch per output phase (o_py, o_px), each with the subset of taps whose parity matches.
#pragma once
#include <stdint.h>

namespace ideas {

constexpr int kMaxTaps = 16;

struct ConvTap {
  int dy, dx, widx;
};

struct ConvGeom {
  int N;
  int IH, IW, IC;        // source tensor (N, IH, IW, IC)
  int OH, OW, OC;        // destination tensor (N, OH, OW, OC)
  int QH, QW;            // extent of the q grid of this launch
  int o_s, o_py, o_px;   // destination pixel = q*o_s + o_p
  int i_s;               // source pixel      = q*i_s + tap.d
  int ntaps;
  ConvTap taps[kMaxTaps];
};

// forward conv: x (N,H,W,C) -> y (N,OH,OW,K), taps in (kh,kw) order, widx = kh*KW+kw
inline ConvGeom geom_forward(int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad, int OH, int OW) {
  ConvGeom g{};
  g.N = N; g.IH = H; g.IW = W; g.IC = C; g.OH = OH; g.OW = OW; g.OC = K;
  g.QH = OH; g.QW = OW; g.o_s = 1; g.o_py = 0; g.o_px = 0; g.i_s = stride;
  g.ntaps = 0;
  for (int a = 0; a < kh; ++a)
    for (int b = 0; b < kw; ++b) g.taps[g.ntaps++] = ConvTap{a - pad, b - pad, a * kw + b};
  return g;
}

// one output phase (py,px) of the data gradient: dy (N,OH,OW,K) -> dx (N,H,W,C).
// dx[iy] = sum_{kh : (iy + pad - kh) % stride == 0} dy[(iy + pad - kh)/stride] * w[kh]; with
// iy = stride*q + py the source row is q + (py + pad - kh)/stride.  The packed dgrad weights
// are flipped, so tap (kh,kw) lives at widx = (KH-1-kh)*KW + (KW-1-kw).
inline ConvGeom geom_dgrad_phase(int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad, int OH,
                                 int OW, int py, int px) {
  ConvGeom g{};
  g.N = N; g.IH = OH; g.IW = OW; g.IC = K; g.OH = H; g.OW = W; g.OC = C;
  g.QH = (H - py + stride - 1) / stride; g.QW = (W - px + stride - 1) / stride;
  g.o_s = stride; g.o_py = py; g.o_px = px; g.i_s = 1;
  g.ntaps = 0;
  for (int a = 0; a < kh; ++a) {
    if (((py + pad - a) % stride + stride) % stride) continue;
    for (int b = 0; b < kw; ++b) {
      if (((px + pad - b) % stride + stride) % stride) continue;
      // floor division is exact here (remainder 0)
      g.taps[g.ntaps++] = ConvTap{(py + pad - a) / stride, (px + pad - b) / stride, (kh - 1 - a) * kw + (kw - 1 - b)};
    }
  }
  return g;
}

}  // namespace ideas
