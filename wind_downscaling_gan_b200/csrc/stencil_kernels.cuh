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

// Shared-memory tiled, register-blocked form of the same layer (both precisions; the tf32 path keeps this 0.14 % of the
// MACs in full fp32).  One block = FT_ROWS output rows x S columns of one image.  The (FT_ROWS + 2) x (S + 2) x 16
// input tile (zero ring of the padded g9 image included) is staged once, converted to fp32, with coalesced 16-byte
// loads at a pixel pitch of 20 floats (conflict-free LDS.128 for threads on consecutive columns).  Every thread then
// computes FT_PER vertically adjacent output pixels: each staged value is read from shared memory once per thread and
// used for up to 3 output rows (18 LDS.128 per output pixel instead of 36 -- the unblocked form is bound by the
// shared-memory port, not by HBM), with the weights in the constant bank.
constexpr int FT_ROWS = 8;
constexpr int FT_PER = 4;
constexpr int FT_PITCH = 20;
template <int PREC>
__global__ void __launch_bounds__(192, 2)
final_conv3x3_tiled_kernel(const typename Prec<PREC>::act_t* __restrict__ in_padded, long long in_sn, long long in_sy,
                           const __grid_constant__ FinalConvW<16, 2> W, float* __restrict__ out, int S) {
  // in_padded points at padded pixel (row 0, col 3) of image 0: interior pixel (y, x) is at row y + 1, col x + 1 of
  // this view; rows have in_sy elements, 16 channels per pixel.
  extern __shared__ float tile[];      // [(FT_ROWS + 2)][(S + 2)][FT_PITCH]
  const int tiles_per_img = (S + FT_ROWS - 1) / FT_ROWS;
  const long long n = blockIdx.x / tiles_per_img;
  const int y0 = (blockIdx.x % tiles_per_img) * FT_ROWS;
  const int TW = S + 2;
  const typename Prec<PREC>::act_t* src = in_padded + n * in_sn + (long long)y0 * in_sy;
  const int out_rows = min(FT_ROWS, S - y0);
  for (int i = threadIdx.x; i < (out_rows + 2) * TW * 2; i += blockDim.x) {        // 8 channels per item
    const int h = i & 1, px = (i >> 1) % TW, r = (i >> 1) / TW;
    float v[8];
    Prec<PREC>::load8(src + (long long)r * in_sy + px * 16 + h * 8, v);
    float4* d = reinterpret_cast<float4*>(tile + (r * TW + px) * FT_PITCH + h * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
  // Each thread walks down FT_PER output rows of one column with three rotating accumulators: input row iy of its
  // strip feeds output rows iy (dy = 0), iy - 1 (dy = 1) and iy - 2 (dy = 2).  The row loop is NOT unrolled (a fully
  // unrolled body made ptxas hoist all 72 LDS.128 and spill); every weight is a compile-time constant-bank operand.
  const int groups = (out_rows + FT_PER - 1) / FT_PER;
  for (int i = threadIdx.x; i < groups * S; i += blockDim.x) {
    const int x = i % S, r0 = (i / S) * FT_PER;           // output rows r0 .. r0 + FT_PER - 1 of the tile
    const int nrow = min(FT_PER, out_rows - r0);
    float a0 = W.b[0], a1 = W.b[1], b0 = a0, b1 = a1, c0 = a0, c1 = a1;   // output rows iy - 2, iy - 1, iy
    const float* col = tile + (r0 * TW + x) * FT_PITCH;
    float* dst = out + ((n * S + y0 + r0) * S + x) * 2;
#pragma unroll 1
    for (int iy = 0; iy < nrow + 2; ++iy) {
      const bool ta = iy >= 2, tb = iy >= 1 && iy <= nrow, tc = iy < nrow;     // warp-uniform
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float4 f[3];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) f[dx] = *reinterpret_cast<const float4*>(col + dx * FT_PITCH + c4 * 4);
#define WDG_FC_ROW(DY, A0, A1)                                                                        \
  _Pragma("unroll") for (int dx = 0; dx < 3; ++dx) {                                                  \
    constexpr int wi = 0;                                                                             \
    A0 = fmaf(f[dx].x, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 0], A0);                        \
    A1 = fmaf(f[dx].x, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 1], A1);                        \
    A0 = fmaf(f[dx].y, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 2], A0);                        \
    A1 = fmaf(f[dx].y, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 3], A1);                        \
    A0 = fmaf(f[dx].z, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 4], A0);                        \
    A1 = fmaf(f[dx].z, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 5], A1);                        \
    A0 = fmaf(f[dx].w, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 6], A0);                        \
    A1 = fmaf(f[dx].w, W.w[(((DY) * 3 + dx) * 16 + c4 * 4) * 2 + wi + 7], A1);                        \
  }
        if (tc) { WDG_FC_ROW(0, c0, c1) }
        if (tb) { WDG_FC_ROW(1, b0, b1) }
        if (ta) { WDG_FC_ROW(2, a0, a1) }
#undef WDG_FC_ROW
      }
      if (ta) *reinterpret_cast<float2*>(dst + (long long)(iy - 2) * S * 2) = make_float2(a0, a1);
      a0 = b0; a1 = b1; b0 = c0; b1 = c1; c0 = W.b[0]; c1 = W.b[1];
      col += TW * FT_PITCH;
    }
  }
}

}  // namespace wdg
