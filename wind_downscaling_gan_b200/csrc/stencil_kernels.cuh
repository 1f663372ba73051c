// Bandwidth-bound kernels of the generator forward: input packing (concat + pad + cast),
// bilinear x2 upsample of the concatenated skip tensor, and the final 3x3 16->2 convolution.
#pragma once
#include "philox.cuh"
#include "ptx.cuh"

namespace wdg {

// K1: Concatenate([image, noise]) + ZeroPadding2D(3) + cast (models.py:28,32), written as the space-to-depth
// act_t image X2[n][Q][Q][(p, q, c)], Q = (S+6)/2: padded pixel (y+3, x+3) = (2Y+p, 2X+q), CP channels per pixel
// (pad channels and the zero ring are never written).  One block per image row: the row's fp32 image (and noise) is
// staged in shared memory with coalesced float4 loads, then each thread emits one CP-channel pixel.
// pack_input_gen_noise_kernel: the noise is not read but DRAWN here, in registers (api.py:136 draws it per group with the TF generator):
// channel c of pixel i is element (i * cnoise + c) of the Philox stream of wdg_noise_normal(seed, offset) scaled by
// noise_std -- bit-identical to generating the (B,T,S,S,cnoise) tensor first, without its 2 x 737 KB per field of HBM
// traffic.  Requires cnoise % 4 == 0 (one counter block = 4 channels of one pixel).
struct NoiseSpec {
  float stddev;
  uint32_t k0, k1;
  unsigned long long offset;   // counter block of element 0
};

template <int PREC>
__global__ void __launch_bounds__(128)
pack_input_s2d_kernel(const float* __restrict__ image, const float* __restrict__ noise,
                      typename Prec<PREC>::act_t* __restrict__ x2, int S, int cin, int cnoise, int CP) {
  using P = Prec<PREC>;
  extern __shared__ float row[];          // [S*cin image | S*cnoise noise]
  const long long r = blockIdx.x;         // n * S + y
  const int y = (int)(r % S);
  const long long n = r / S;
  const float4* img4 = reinterpret_cast<const float4*>(image + r * S * cin);
  const float4* noi4 = reinterpret_cast<const float4*>(noise + r * S * cnoise);
  const int n_img4 = S * cin / 4, n_noi4 = S * cnoise / 4;
  float4* row4 = reinterpret_cast<float4*>(row);
  for (int i = threadIdx.x; i < n_img4; i += blockDim.x) row4[i] = __ldg(img4 + i);
  for (int i = threadIdx.x; i < n_noi4; i += blockDim.x) row4[n_img4 + i] = __ldg(noi4 + i);
  __syncthreads();
  const int Q = (S + 6) / 2;
  const int py = y + 3, Y = py >> 1, p = py & 1;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    const int px = x + 3, X = px >> 1, q = px & 1;
    typename P::act_t* dst = x2 + ((((n * Q + Y) * Q + X) * 2 + p) * 2 + q) * CP;
    const float* ip = row + x * cin;
    const float* np = row + S * cin + x * cnoise;
    for (int c = 0; c < CP; c += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int cc = c + k;
        v[k] = cc < cin ? ip[cc] : (cc < cin + cnoise ? np[cc - cin] : 0.f);
      }
      P::store8(dst + c, v);
    }
  }
}

// Same with the channel counts fixed at compile time (the reference's 3 + 20 -> 24, api.py:25-27) and the noise drawn
// in registers; one thread per pixel, everything unrolled (no local-memory arrays).
template <int PREC, int CIN, int CNOISE, int CP>
__global__ void __launch_bounds__(128)
pack_input_gen_noise_kernel(const float* __restrict__ image, NoiseSpec ns, typename Prec<PREC>::act_t* __restrict__ x2, int S) {
  using P = Prec<PREC>;
  static_assert(CNOISE % 4 == 0 && CP % 8 == 0 && CIN + CNOISE <= CP, "channel layout");
  extern __shared__ float row[];          // [S*CIN image]
  const long long r = blockIdx.x;         // n * S + y
  const int y = (int)(r % S);
  const long long n = r / S;
  const float4* img4 = reinterpret_cast<const float4*>(image + r * S * CIN);
  float4* row4 = reinterpret_cast<float4*>(row);
  for (int i = threadIdx.x; i < S * CIN / 4; i += blockDim.x) row4[i] = __ldg(img4 + i);
  __syncthreads();
  const int Q = (S + 6) / 2;
  const int py = y + 3, Y = py >> 1, p = py & 1;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    const int px = x + 3, X = px >> 1, q = px & 1;
    typename P::act_t* dst = x2 + ((((n * Q + Y) * Q + X) * 2 + p) * 2 + q) * CP;
    float v[CP];
#pragma unroll
    for (int c = 0; c < CIN; ++c) v[c] = row[x * CIN + c];
    const unsigned long long blk0 = ns.offset + (unsigned long long)((r * S + x) * (long long)(CNOISE / 4));
#pragma unroll
    for (int b = 0; b < CNOISE / 4; ++b) {
      float nz[4];
      philox_normal4(blk0 + b, ns.k0, ns.k1, nz);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[CIN + 4 * b + k] = nz[k] * ns.stddev;
    }
#pragma unroll
    for (int c = CIN + CNOISE; c < CP; ++c) v[c] = 0.f;
#pragma unroll
    for (int c = 0; c < CP; c += 8) {
      const float w[8] = {v[c], v[c + 1], v[c + 2], v[c + 3], v[c + 4], v[c + 5], v[c + 6], v[c + 7]};
      P::store8(dst + c, w);
    }
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
template <int PREC>
__global__ void edge_lines_kernel(const typename Prec<PREC>::act_t* __restrict__ catp, typename Prec<PREC>::act_t* __restrict__ E,
                                  long long total, int h, int C, int pitch) {
  using P = Prec<PREC>;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int groups = C / 8;
  const int PP = 2 * h + 8;
  // one 64-bit division for the image index, 32-bit arithmetic below it (the kernel is index-math bound)
  const unsigned per_img = (unsigned)groups * PP * 4;
  const long long n = i / per_img;
  const unsigned r = (unsigned)(i - n * per_img);
  const int g = (int)(r % groups);
  const unsigned pe = r / groups;
  const int p = (int)(pe % PP);
  const int edge = (int)(pe / PP);
  const int j = p - 4;
  const int PW = h + 4;
  const typename P::act_t* img = catp + n * (long long)PW * PW * pitch + g * 8;
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
      float f[8];
      P::load8(img + ((long long)(y + 2) * PW + (x + 2)) * pitch, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += ws[t] * f[k];
    }
  }
  P::store8(E + ((n * 4 + edge) * PP + p) * C + g * 8, acc);
}

// K9: Conv2D(out_channels, 3x3, 'same', linear) on the post-BatchNorm 16-channel tensor (models.py:70-71).
// CUDA cores: K = 144, N = 2 is a bandwidth-bound stencil.  The 288 weights travel as a kernel parameter, i.e. in
// the constant bank, so every FFMA takes its weight as a constant operand (no shared-memory or global loads).
// One thread per output pixel.
template <int CIN, int COUT>
struct FinalConvW {
  float w[9 * CIN * COUT];  // [3][3][CIN][COUT]
  float b[COUT];
};

template <int CIN, int COUT, int PREC>
__global__ void __launch_bounds__(128)
final_conv3x3_kernel(const typename Prec<PREC>::act_t* __restrict__ in, long long in_sn, long long in_sy,
                     const __grid_constant__ FinalConvW<CIN, COUT> W, float* __restrict__ out, long long npix, int S) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int x = (int)(i % S);
  const int y = (int)((i / S) % S);
  const long long n = i / ((long long)S * S);
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = W.b[o];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y + dy - 1;
    if (yy < 0 || yy >= S) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int xx = x + dx - 1;
      if (xx < 0 || xx >= S) continue;
      const typename Prec<PREC>::act_t* p = in + n * in_sn + yy * in_sy + (long long)xx * CIN;
#pragma unroll
      for (int c8 = 0; c8 < CIN; c8 += 8) {
        float f[8];
        Prec<PREC>::load8(p + c8, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int o = 0; o < COUT; ++o) acc[o] = fmaf(f[k], W.w[((dy * 3 + dx) * CIN + c8 + k) * COUT + o], acc[o]);
        }
      }
    }
  }
  if (COUT == 2) {
    *reinterpret_cast<float2*>(out + i * 2) = make_float2(acc[0], acc[1]);
  } else {
#pragma unroll
    for (int o = 0; o < COUT; ++o) out[i * COUT + o] = acc[o];
  }
}

}  // namespace wdg
