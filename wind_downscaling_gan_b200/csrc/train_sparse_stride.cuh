// Direct fp32 kernels for convolutions whose stride exceeds the kernel size: the critic's shortcut convolution of the
// shipped checkpoint topology (tf_utils.py:15-32 at 96 px: 6x6 kernel, stride 11, padding 4, 9x9x128 -> 2x2x256).
// The windows do not overlap and most of the zero-padded window lies outside the image (4 of 36 taps hit a pixel), so the
// layer is a handful of small dot products per output: as a strided implicit GEMM its backward-data would be stride^2 = 121
// launches of mostly-empty residue classes.  One thread per output element, exact fp32, fixed summation order
// (deterministic: data-parallel replicas stay bit-identical), no scratch.
#pragma once
#include <cuda_runtime.h>

#include "train_geo.cuh"

namespace wdg_sparse {

// y[n,oy,ox,co] = leaky_alpha((y +) bias[co] + sum_{ky,kx,ci} x[n,oy*s-pt+ky,ox*s-pl+kx,ci] * w[ky][kx][ci][co])
__global__ void __launch_bounds__(256) fwd_kernel(ConvGeo g, const float* __restrict__ x, const float* __restrict__ w,
                                                  const float* __restrict__ bias, float* __restrict__ y, int accumulate, float alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.Ho * g.Wo * g.Co;
  if (i >= total) return;
  const int co = (int)(i % g.Co);
  const long long p = i / g.Co;
  const int ox = (int)(p % g.Wo), oy = (int)((p / g.Wo) % g.Ho);
  const long long n = p / ((long long)g.Wo * g.Ho);
  float acc = bias ? bias[co] : 0.f;
  for (int ky = 0; ky < g.kh; ++ky) {
    const int iy = oy * g.stride - g.pad_t + ky;
    if (iy < 0 || iy >= g.H) continue;
    for (int kx = 0; kx < g.kw; ++kx) {
      const int ix = ox * g.stride - g.pad_l + kx;
      if (ix < 0 || ix >= g.W) continue;
      const float* xp = x + ((n * g.H + iy) * g.W + ix) * g.x_cs + g.x_co;
      const float* wp = w + (long long)(ky * g.kw + kx) * g.Ci * g.Co + co;
      for (int ci = 0; ci < g.Ci; ++ci) acc = fmaf(__ldg(xp + ci), __ldg(wp + (long long)ci * g.Co), acc);
    }
  }
  float* yp = y + p * g.y_cs + g.y_co + co;
  const float v = accumulate ? *yp + acc : acc;
  *yp = v >= 0.f ? v : alpha * v;
}

// dx[n,iy,ix,ci] (+)= sum_co dy[n,oy,ox,co] * w[ky][kx][ci][co] for the ONE (oy,ky) / (ox,kx) with oy*s - pt + ky = iy
// (stride > kernel: at most one window covers a pixel); pixels no window covers receive 0.
__global__ void __launch_bounds__(256) bwd_data_kernel(ConvGeo g, const float* __restrict__ dy, const float* __restrict__ w,
                                                       float* __restrict__ dx, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.H * g.W * g.Ci;
  if (i >= total) return;
  const int ci = (int)(i % g.Ci);
  const long long p = i / g.Ci;
  const int ix = (int)(p % g.W), iy = (int)((p / g.W) % g.H);
  const long long n = p / ((long long)g.W * g.H);
  const int ty = iy + g.pad_t, tx = ix + g.pad_l;
  const int oy = ty / g.stride, ky = ty - oy * g.stride, ox = tx / g.stride, kx = tx - ox * g.stride;
  float acc = 0.f;
  if (ky < g.kh && oy < g.Ho && kx < g.kw && ox < g.Wo) {
    const float* dp = dy + ((n * g.Ho + oy) * g.Wo + ox) * g.y_cs + g.y_co;
    const float* wp = w + ((long long)(ky * g.kw + kx) * g.Ci + ci) * g.Co;
    for (int co = 0; co < g.Co; ++co) acc = fmaf(__ldg(dp + co), __ldg(wp + co), acc);
  }
  float* xp = dx + p * g.x_cs + g.x_co + ci;
  *xp = accumulate ? *xp + acc : acc;
}

// dw[ky][kx][ci][co] (+)= sum_{n,oy,ox} x[n,oy*s-pt+ky,ox*s-pl+kx,ci] * dy[n,oy,ox,co], images in order
__global__ void __launch_bounds__(256) bwd_weight_kernel(ConvGeo g, const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.kh * g.kw * g.Ci * g.Co;
  if (i >= total) return;
  const int co = (int)(i % g.Co), ci = (int)((i / g.Co) % g.Ci);
  const int t = (int)(i / ((long long)g.Co * g.Ci)), kx = t % g.kw, ky = t / g.kw;
  float acc = 0.f;
  for (int oy = 0; oy < g.Ho; ++oy) {
    const int iy = oy * g.stride - g.pad_t + ky;
    if (iy < 0 || iy >= g.H) continue;
    for (int ox = 0; ox < g.Wo; ++ox) {
      const int ix = ox * g.stride - g.pad_l + kx;
      if (ix < 0 || ix >= g.W) continue;
      for (long long n = 0; n < g.N; ++n)
        acc = fmaf(__ldg(x + ((n * g.H + iy) * g.W + ix) * g.x_cs + g.x_co + ci),
                   __ldg(dy + ((n * g.Ho + oy) * g.Wo + ox) * g.y_cs + g.y_co + co), acc);
    }
  }
  dw[i] = accumulate ? dw[i] + acc : acc;
}

static inline bool applies(const ConvGeo& g) { return g.stride > g.kh && g.stride > g.kw; }

}  // namespace wdg_sparse
