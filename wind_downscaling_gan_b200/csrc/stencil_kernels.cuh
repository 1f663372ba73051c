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

// K7 border lines.  The fused upsample + 5x5 transposed convolution (conv_umma EPI_UPCONV) works on the
// zero-padded low-res concat image `catp` [n][h+4][w+4][C] (interior at +2,+2), i.e. it sees the bilinear
// interpolation continued with zeros beyond the border.  The reference instead clamps the interpolation at the
// border and zero-pads the UPSAMPLED image (models.py:62-63).  The difference is confined to high-res rows/cols
// {-1, 0, 2h-1, 2h}: +-0.25 x the edge row/column upsampled along the edge ("border dipole").
// This kernel writes those four upsampled edge lines, E[n][edge][P = 2h+8][C], position p = j + 4:
//   edge 0/1 (top/bottom): clamped x-upsample of low-res row 0 / h-1,       j in [0, 2w)
//   edge 2/3 (left/right): zero-extended y-upsample of low-res col 0 / w-1, j in [-1, 2h]
// One thread per (n, edge, p, 8-channel group).
__global__ void edge_lines_kernel(const __nv_bfloat16* __restrict__ catp, __nv_bfloat16* __restrict__ E,
                                  long long total, int h, int C, int pitch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int groups = C / 8;
  const int P = 2 * h + 8;
  const int g = (int)(i % groups);
  const int p = (int)((i / groups) % P);
  const int edge = (int)((i / ((long long)groups * P)) % 4);
  const long long n = i / ((long long)groups * P * 4);
  const int j = p - 4;
  const int PW = h + 4;
  const __nv_bfloat16* img = catp + n * (long long)PW * PW * pitch + g * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const bool clamped = edge < 2;
  const bool in_range = clamped ? (j >= 0 && j < 2 * h) : (j >= -1 && j <= 2 * h);
  if (in_range) {
    const int odd = j & 1;
    const int k0 = (j - odd) / 2;                 // floor(j / 2)
    int ka = odd ? k0 : k0 - 1, kb = odd ? k0 + 1 : k0;
    const float wa = odd ? 0.75f : 0.25f;
    if (clamped) { ka = max(ka, 0); kb = min(kb, h - 1); }
    const int fixed = (edge == 0 || edge == 2) ? 0 : h - 1;
    const int ks[2] = {ka, kb};
    const float ws[2] = {wa, 1.f - wa};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      // padded coordinates (+2): positions outside [0, h) read the physical zero ring
      const int y = (edge < 2) ? fixed : ks[t];
      const int x = (edge < 2) ? ks[t] : fixed;
      const uint4 v = *reinterpret_cast<const uint4*>(img + ((long long)(y + 2) * PW + (x + 2)) * pitch);
      const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(hv[k]);
        acc[2 * k] += ws[t] * f.x;
        acc[2 * k + 1] += ws[t] * f.y;
      }
    }
  }
  *reinterpret_cast<uint4*>(E + ((n * 4 + edge) * P + p) * C + g * 8) =
      make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                 pack_bf16x2(acc[6], acc[7]));
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
