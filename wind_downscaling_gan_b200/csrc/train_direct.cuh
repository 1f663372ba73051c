// Direct fp32 kernels for the narrow 3x3 / stride-1 / 'same' convolutions of the training path (the critic's
// full-resolution layers with 2..16 channels, the generator's 16 -> 2 output conv): with K = 9 Ci <= 144 and
// N = Co <= 64 they are bandwidth-bound stencils, and as 128 x 16 GEMM tiles of a single K block they spent their time
// in per-tile overheads (13 824 CTAs of one K block each; measured 260-480 us against 15-60 us of memory time).
// One thread per pixel (forward, backward-data) / per weight element (backward-weight); weights live in shared
// memory and are read as broadcast float4s; every loop is unrolled at compile time for the (Ci, Co) pair.
#pragma once
#include <cuda_runtime.h>

namespace wdg_direct {

// y[n,y,x,co] = leaky_alpha((y +) bias[co] + sum_{ky,kx,ci} x[n,y+ky-1,x+kx-1,ci] * w[ky][kx][ci][co])
template <int CI, int CO>
__global__ void __launch_bounds__(256) conv3x3_fwd_kernel(const float* __restrict__ x, int x_cs, int x_co, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ y, int y_cs, int y_co,
                                                          long long npix, int H, int W, int accumulate, float alpha) {
  __shared__ __align__(16) float sw[9 * CI * CO];
  for (int i = threadIdx.x; i < 9 * CI * CO; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int px = (int)(p % W), py = (int)((p / W) % H);
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = bias ? bias[o] : 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const float* xp = x + (p + (long long)(ky - 1) * W + (kx - 1)) * x_cs + x_co;
      float xv[CI];
#pragma unroll
      for (int c = 0; c < CI; ++c) xv[c] = xp[c];
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const float* wr = sw + ((ky * 3 + kx) * CI + c) * CO;
        if constexpr (CO % 4 == 0) {
#pragma unroll
          for (int o = 0; o < CO; o += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + o);
            acc[o] = fmaf(xv[c], w4.x, acc[o]); acc[o + 1] = fmaf(xv[c], w4.y, acc[o + 1]);
            acc[o + 2] = fmaf(xv[c], w4.z, acc[o + 2]); acc[o + 3] = fmaf(xv[c], w4.w, acc[o + 3]);
          }
        } else {
#pragma unroll
          for (int o = 0; o < CO; ++o) acc[o] = fmaf(xv[c], wr[o], acc[o]);
        }
      }
    }
  }
  float* yp = y + p * y_cs + y_co;
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    const float v = accumulate ? yp[o] + acc[o] : acc[o];
    yp[o] = v >= 0.f ? v : alpha * v;
  }
}

// dx[n,iy,ix,ci] (+)= sum_{ky,kx,co} dy[n,iy+1-ky,ix+1-kx,co] * w[ky][kx][ci][co]
template <int CI, int CO>
__global__ void __launch_bounds__(256) conv3x3_bwd_data_kernel(const float* __restrict__ dy, int y_cs, int y_co, const float* __restrict__ w,
                                                               float* __restrict__ dx, int x_cs, int x_co, long long npix, int H, int W,
                                                               int accumulate) {
  __shared__ __align__(16) float sw[9 * CI * CO];
  for (int i = threadIdx.x; i < 9 * CI * CO; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int px = (int)(p % W), py = (int)((p / W) % H);
  float acc[CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) acc[c] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int oy = py + 1 - ky;
    if (oy < 0 || oy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ox = px + 1 - kx;
      if (ox < 0 || ox >= W) continue;
      const float* dp = dy + (p + (long long)(1 - ky) * W + (1 - kx)) * y_cs + y_co;
#pragma unroll
      for (int o0 = 0; o0 < CO; o0 += 4) {
        float dv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) dv[e] = o0 + e < CO ? dp[o0 + e] : 0.f;
#pragma unroll
        for (int c = 0; c < CI; ++c) {
          const float* wr = sw + ((ky * 3 + kx) * CI + c) * CO + o0;
          if constexpr (CO % 4 == 0) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr);
            acc[c] = fmaf(dv[0], w4.x, acc[c]); acc[c] = fmaf(dv[1], w4.y, acc[c]);
            acc[c] = fmaf(dv[2], w4.z, acc[c]); acc[c] = fmaf(dv[3], w4.w, acc[c]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (o0 + e < CO) acc[c] = fmaf(dv[e], wr[e], acc[c]);
          }
        }
      }
    }
  }
  float* xp = dx + p * x_cs + x_co;
#pragma unroll
  for (int c = 0; c < CI; ++c) xp[c] = accumulate ? xp[c] + acc[c] : acc[c];
}

// part[slab][(ky,kx,ci,co)] = sum over the slab's pixels of x[n,y+ky-1,x+kx-1,ci] * dy[n,y,x,co]
// Block = one slab of pixels, processed in chunks of PCH pixels staged in shared memory (the 9 shifted copies of x with
// zeros outside the image, and dy); thread t owns the weight elements t, t + 256, ...
template <int CI, int CO>
__global__ void __launch_bounds__(256) conv3x3_bwd_weight_kernel(const float* __restrict__ x, int x_cs, int x_co,
                                                                 const float* __restrict__ dy, int y_cs, int y_co,
                                                                 float* __restrict__ part, long long npix, int H, int W,
                                                                 long long px_per_slab) {
  constexpr int O = 9 * CI * CO;
  constexpr int PCH = 64;
  constexpr int NOUT = (O + 255) / 256;
  __shared__ float xs[9 * CI][PCH + 1];      // [(tap, ci)][pixel]
  __shared__ float ds[CO][PCH + 1];          // [co][pixel]
  __shared__ int spx[PCH], spy[PCH];         // image coordinates of the chunk's pixels
  const long long p0 = blockIdx.x * px_per_slab, p1 = min(npix, p0 + px_per_slab);
  float acc[NOUT];
  int row_x[NOUT], row_d[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    acc[j] = 0.f;
    const int e = threadIdx.x + 256 * j;        // (tap, ci, co) flattened, co fastest
    row_x[j] = e < O ? e / CO : 0;              // (tap, ci)
    row_d[j] = e < O ? e % CO : 0;
  }
  for (long long c0 = p0; c0 < p1; c0 += PCH) {
    __syncthreads();
    if (threadIdx.x < PCH) {
      const long long p = c0 + threadIdx.x;
      spx[threadIdx.x] = (int)(p % W); spy[threadIdx.x] = (int)((p / W) % H);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PCH * 9 * CI; i += blockDim.x) {
      const int ci = i % CI, tap = (i / CI) % 9, q = i / (9 * CI);
      const long long p = c0 + q;
      float v = 0.f;
      if (p < p1) {
        const int px = spx[q], py = spy[q];
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[(p + (long long)(tap / 3 - 1) * W + (tap % 3 - 1)) * x_cs + x_co + ci];
      }
      xs[tap * CI + ci][q] = v;
    }
    for (int i = threadIdx.x; i < PCH * CO; i += blockDim.x) {
      const int o = i % CO, q = i / CO;
      const long long p = c0 + q;
      ds[o][q] = p < p1 ? dy[p * y_cs + y_co + o] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const float* xr = xs[row_x[j]];
      const float* dr = ds[row_d[j]];
      float a = acc[j];
#pragma unroll 8
      for (int q = 0; q < PCH; ++q) a = fmaf(xr[q], dr[q], a);
      acc[j] = a;
    }
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    const int e = threadIdx.x + 256 * j;
    if (e < O) part[(long long)blockIdx.x * O + e] = acc[j];
  }
}

}  // namespace wdg_direct
