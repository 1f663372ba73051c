// Bandwidth-bound kernels of the generator forward: input packing (concat + pad + cast),
// bilinear x2 upsample of the concatenated skip tensor, and the final 3x3 16->2 convolution.
#pragma once
#include "ptx.cuh"

namespace wdg {

// K1: Concatenate([image, noise]) + ZeroPadding2D(3) + cast (models.py:28,32), written as the
// zero-padded bf16 image [n][S+6][S+6][CP] that the 8x8 stride-2 implicit GEMM reads through an
// overlapping-stride tensor map.  One thread per pixel; border and pad channels stay zero.
__global__ void pack_input_kernel(const float* __restrict__ image, const float* __restrict__ noise,
                                  __nv_bfloat16* __restrict__ xpad, long long npix, int S, int cin, int cnoise,
                                  int CP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int x = (int)(i % S);
  const int y = (int)((i / S) % S);
  const long long n = i / ((long long)S * S);
  const int SP = S + 6;
  __nv_bfloat16* dst = xpad + (((n * SP) + (y + 3)) * SP + (x + 3)) * CP;
  const float* ip = image + i * cin;
  const float* np = noise + i * cnoise;
  for (int c = 0; c < CP; c += 2) {
    float a = 0.f, b = 0.f;
    if (c < cin) a = ip[c]; else if (c < cin + cnoise) a = np[c - cin];
    const int c1 = c + 1;
    if (c1 < cin) b = ip[c1]; else if (c1 < cin + cnoise) b = np[c1 - cin];
    *reinterpret_cast<uint32_t*>(dst + c) = pack_bf16x2(a, b);
  }
}

// K7: Concatenate([g7, res_2]) + UpSampling2D(2, 'bilinear') (models.py:60-62), half-pixel centres
// with edge clamp.  One thread per (output pixel, 8-channel group); fp32 blend, bf16 store.
// g7: [n][h][w][C0] plain; res2: zero-padded [n][h+2][w+2][C1] (interior at +1,+1).
__global__ void upsample_concat_kernel(const __nv_bfloat16* __restrict__ g7, const __nv_bfloat16* __restrict__ res2,
                                       __nv_bfloat16* __restrict__ up, long long total, int h, int w, int C0, int C1) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C = C0 + C1;
  const int groups = C / 8;
  const int g = (int)(i % groups);
  const long long pix = i / groups;
  const int X = (int)(pix % (2 * w));
  const int Y = (int)((pix / (2 * w)) % (2 * h));
  const long long n = pix / ((long long)4 * w * h);
  // source taps: out[2k] = .25 in[k-1] + .75 in[k]; out[2k+1] = .75 in[k] + .25 in[k+1] (clamped)
  const int ky = Y >> 1, kx = X >> 1;
  const int y0 = (Y & 1) ? ky : max(ky - 1, 0);
  const int y1 = (Y & 1) ? min(ky + 1, h - 1) : ky;
  const float wy0 = (Y & 1) ? 0.75f : 0.25f;
  const int x0 = (X & 1) ? kx : max(kx - 1, 0);
  const int x1 = (X & 1) ? min(kx + 1, w - 1) : kx;
  const float wx0 = (X & 1) ? 0.75f : 0.25f;
  const int c = g * 8;
  const __nv_bfloat16* src;
  long long sy, sx, base;
  if (c < C0) {
    src = g7 + c; sx = C0; sy = (long long)w * C0; base = n * h * sy;
  } else {
    src = res2 + (c - C0); sx = C1; sy = (long long)(w + 2) * C1; base = n * (h + 2) * sy + sy + sx;
  }
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int ys[2] = {y0, y1};
  const int xs[2] = {x0, x1};
  const float wys[2] = {wy0, 1.f - wy0};
  const float wxs[2] = {wx0, 1.f - wx0};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + base + ys[a] * sy + xs[b] * sx);
      const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
      const float wgt = wys[a] * wxs[b];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(hv[k]);
        acc[2 * k] += wgt * f.x;
        acc[2 * k + 1] += wgt * f.y;
      }
    }
  uint4 o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                       pack_bf16x2(acc[6], acc[7]));
  *reinterpret_cast<uint4*>(up + pix * C + c) = o;
}

// K9: Conv2D(out_channels, 3x3, 'same', linear) on the post-BatchNorm 16-channel tensor
// (models.py:70-71).  CUDA cores: K = 144, N = 2 is a bandwidth-bound stencil.  One thread per pixel.
template <int CIN, int COUT>
__global__ void final_conv3x3_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ wgt /*[3][3][CIN][COUT]*/,
                                     const float* __restrict__ bias, float* __restrict__ out, long long npix, int S) {
  __shared__ float ws[9 * CIN * COUT];
  for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) ws[i] = wgt[i];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int x = (int)(i % S);
  const int y = (int)((i / S) % S);
  const long long n = i / ((long long)S * S);
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = bias[o];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y + dy - 1;
    if (yy < 0 || yy >= S) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int xx = x + dx - 1;
      if (xx < 0 || xx >= S) continue;
      const __nv_bfloat16* p = in + ((n * S + yy) * S + xx) * CIN;
      const float* wk = ws + (dy * 3 + dx) * CIN * COUT;
#pragma unroll
      for (int c8 = 0; c8 < CIN; c8 += 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(p + c8);
        const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(hv[k]);
#pragma unroll
          for (int o = 0; o < COUT; ++o) {
            acc[o] += f.x * wk[(c8 + 2 * k) * COUT + o];
            acc[o] += f.y * wk[(c8 + 2 * k + 1) * COUT + o];
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < COUT; ++o) out[i * COUT + o] = acc[o];
}

}  // namespace wdg
