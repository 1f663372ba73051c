// Building blocks of the WGAN training step (critic forward/backward, training-mode
// generator, optimiser): the C ABI of the training path, the exact fp32 convolution GEMMs on CUDA cores (forward /
// backward-data / backward-weight; the tcgen05 tf32 / bf16 versions live in train_gemm_tc.cu and are selected by
// wdg_train_set_precision), normalisation layers, ConvLSTM gate math and the fused small-filter cell, bilinear resize
// and its adjoint, overlap-add, column sums, Adam, spectral normalisation.  All tensors are fp32 channels-last;
// `cs`/`co` arguments are the channel stride / offset of a tensor inside a wider (concatenated) buffer.
#include <cuda_runtime.h>

#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "../../include/wdg.h"
#include "train_direct.cuh"
#include "train_sparse_stride.cuh"
#include "train_geo.cuh"

namespace {

// ------------------------------------------------------------------ implicit GEMM on CUDA cores
// C[M][N] (+)= sum_k A(m,k) * B(k,n); 64x64 block tile, 16-deep K tiles, 256 threads x (4x4) outputs.
// Each problem supplies tile loaders that (a) walk global memory along its contiguous axis and (b) decompose the
// GEMM indices into (image, y, x) / (tap, channel) ONCE per thread instead of once per element.
// Two tile shapes: 64x64 (4x4 outputs per thread) and, for GEMMs whose N axis is <= 16 wide (tiny channel counts of
// the critic's ConvLSTMs, the 16-channel transposed conv, 2-channel output conv), 128x16 (2x4 per thread).
constexpr int TK = 16;
template <int TM, int TN>
struct Tiles {
  using A = float[TK][TM + 4];
  using B = float[TK][TN + 4];
  static constexpr int RA = TM / 16;            // A elements per thread (K-fastest: rows (tid>>4) + 16 i)
  static constexpr int RBK = TN / 16;           // B elements per thread, K-fastest loaders
  static constexpr int RBN = TN * TK / 256;     // B elements per thread, N-fastest loaders
  static constexpr int KSTEP_N = 256 / TN;      // k rows covered per pass by N-fastest loaders
  static constexpr int RAM = TM * TK / 256;     // A elements per thread, M-fastest loader
  static constexpr int KSTEP_M = 256 / TM;
};

struct FwdProblem {       // y = conv(x, w) + bias
  ConvGeo g; const float* x; const float* w; const float* bias; float* y; int accumulate; float alpha;
  __device__ int M() const { return g.N * g.Ho * g.Wo; }
  __device__ int Nn() const { return g.Co; }
  __device__ int K() const { return g.kh * g.kw * g.Ci; }
  template <int TM, int TN>
  struct ALoader {        // K-fastest: thread owns column kk = tid & 15 and rows (tid >> 4) + 16 i
    static constexpr int R = Tiles<TM, TN>::RA;
    const FwdProblem& p; int kk; int iy0[R], ix0[R]; const float* base[R];
    __device__ ALoader(const FwdProblem& p_, int m0, int tid) : p(p_), kk(tid & 15) {
      const ConvGeo& g = p.g;
      const int M = p.M();
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int m = m0 + (tid >> 4) + 16 * i;
        if (m < M) {
          const int ox = m % g.Wo, oy = (m / g.Wo) % g.Ho, n = m / (g.Wo * g.Ho);
          iy0[i] = oy * g.stride - g.pad_t; ix0[i] = ox * g.stride - g.pad_l;
          base[i] = p.x + (long long)n * g.H * g.W * g.x_cs + g.x_co;
        } else { iy0[i] = -(1 << 28); ix0[i] = 0; base[i] = p.x; }
      }
    }
    __device__ void load(typename Tiles<TM, TN>::A& sA, int tid, int k0, int k_end) const {
      const ConvGeo& g = p.g;
      const int k = k0 + kk;
      const bool kv = k < k_end;
      const int ci = k % g.Ci, tap = k / g.Ci, kx = tap % g.kw, ky = tap / g.kw;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int iy = iy0[i] + ky, ix = ix0[i] + kx;
        float v = 0.f;
        if (kv && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = base[i][((long long)iy * g.W + ix) * g.x_cs + ci];
        sA[kk][(tid >> 4) + 16 * i] = v;
      }
    }
  };
  template <int TM, int TN>
  struct BLoader {        // N-fastest: w[k][n]
    const FwdProblem& p; int n; bool nv;
    __device__ BLoader(const FwdProblem& p_, int n0, int tid) : p(p_), n(n0 + tid % TN), nv(n0 + tid % TN < p_.g.Co) {}
    __device__ void load(typename Tiles<TM, TN>::B& sB, int tid, int k0, int k_end) const {
#pragma unroll
      for (int i = 0; i < Tiles<TM, TN>::RBN; ++i) {
        const int kk = tid / TN + Tiles<TM, TN>::KSTEP_N * i, k = k0 + kk;
        sB[kk][tid % TN] = (nv && k < k_end) ? p.w[(long long)k * p.g.Co + n] : 0.f;
      }
    }
  };
  __device__ void store(int m, int n, float v) const {
    float* q = y + (long long)m * g.y_cs + g.y_co + n;
    if (bias) v += bias[n];
    if (accumulate) v += *q;
    *q = v >= 0.f ? v : alpha * v;
  }
};

struct BwdDataProblem {   // dx = conv_bwd_data(dy, w)
  ConvGeo g; const float* dy; const float* w; float* dx; int accumulate;
  __device__ int M() const { return g.N * g.H * g.W; }
  __device__ int Nn() const { return g.Ci; }
  __device__ int K() const { return g.kh * g.kw * g.Co; }
  template <int TM, int TN>
  struct ALoader {        // K-fastest (co is the inner K index and contiguous in dy)
    static constexpr int R = Tiles<TM, TN>::RA;
    const BwdDataProblem& p; int kk; int ty0[R], tx0[R]; const float* base[R];
    __device__ ALoader(const BwdDataProblem& p_, int m0, int tid) : p(p_), kk(tid & 15) {
      const ConvGeo& g = p.g;
      const int M = p.M();
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int m = m0 + (tid >> 4) + 16 * i;
        if (m < M) {
          const int ix = m % g.W, iy = (m / g.W) % g.H, n = m / (g.W * g.H);
          ty0[i] = iy + g.pad_t; tx0[i] = ix + g.pad_l;
          base[i] = p.dy + (long long)n * g.Ho * g.Wo * g.y_cs + g.y_co;
        } else { ty0[i] = -(1 << 28); tx0[i] = 0; base[i] = p.dy; }
      }
    }
    __device__ void load(typename Tiles<TM, TN>::A& sA, int tid, int k0, int k_end) const {
      const ConvGeo& g = p.g;
      const int k = k0 + kk;
      const bool kv = k < k_end;
      const int co = k % g.Co, tap = k / g.Co, kx = tap % g.kw, ky = tap / g.kw;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int ty = ty0[i] - ky, tx = tx0[i] - kx;
        float v = 0.f;
        if (kv && ty >= 0 && tx >= 0) {
          int oy = ty, ox = tx;
          bool ok = true;
          if (g.stride > 1) { oy = ty / g.stride; ox = tx / g.stride; ok = (oy * g.stride == ty) && (ox * g.stride == tx); }
          if (ok && oy < g.Ho && ox < g.Wo) v = base[i][((long long)oy * g.Wo + ox) * g.y_cs + co];
        }
        sA[kk][(tid >> 4) + 16 * i] = v;
      }
    }
  };
  template <int TM, int TN>
  struct BLoader {        // K-fastest: w[tap][ci = n][co], co contiguous
    const BwdDataProblem& p; int kk; int n0;
    __device__ BLoader(const BwdDataProblem& p_, int n0_, int tid) : p(p_), kk(tid & 15), n0(n0_) {}
    __device__ void load(typename Tiles<TM, TN>::B& sB, int tid, int k0, int k_end) const {
      const ConvGeo& g = p.g;
      const int k = k0 + kk;
      const bool kv = k < k_end;
      const int co = k % g.Co, tap = k / g.Co;
#pragma unroll
      for (int i = 0; i < Tiles<TM, TN>::RBK; ++i) {
        const int nn = (tid >> 4) + 16 * i, n = n0 + nn;
        sB[kk][nn] = (kv && n < g.Ci) ? p.w[((long long)tap * g.Ci + n) * g.Co + co] : 0.f;
      }
    }
  };
  __device__ void store(int m, int n, float v) const {
    float* q = dx + (long long)m * g.x_cs + g.x_co + n;
    *q = accumulate ? *q + v : v;
  }
};

// Backward-data of a stride-s convolution, one residue class (ry, rx) of (iy+pad_t, ix+pad_l) mod s at a time: inside
// a class every input pixel sees the same taps ky = ry + s*j, so the GEMM has no structural zeros (a dense
// formulation multiplies zeros for (s*s-1)/(s*s) of its K axis: 8/9 for the critic's 7x7 stride-3 convolutions).
struct BwdDataClassProblem {
  ConvGeo g; const float* dy; const float* w; float* dx; int accumulate;
  int ry, rx, fy, fx, Hc, Wc, Jy, Jx;   // class residues, first pixel of the class, class grid, taps per axis
  __device__ int M() const { return g.N * Hc * Wc; }
  __device__ int Nn() const { return g.Ci; }
  __device__ int K() const { return Jy * Jx * g.Co; }
  template <int TM, int TN>
  struct ALoader {
    static constexpr int R = Tiles<TM, TN>::RA;
    const BwdDataClassProblem& p; int kk; int ay[R], ax[R]; const float* base[R];
    __device__ ALoader(const BwdDataClassProblem& p_, int m0, int tid) : p(p_), kk(tid & 15) {
      const ConvGeo& g = p.g;
      const int M = p.M();
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int m = m0 + (tid >> 4) + 16 * i;
        if (m < M) {
          const int b = m % p.Wc, a = (m / p.Wc) % p.Hc, n = m / (p.Wc * p.Hc);
          ay[i] = (p.fy + g.stride * a + g.pad_t - p.ry) / g.stride;   // = oy + j
          ax[i] = (p.fx + g.stride * b + g.pad_l - p.rx) / g.stride;
          base[i] = p.dy + (long long)n * g.Ho * g.Wo * g.y_cs + g.y_co;
        } else { ay[i] = -(1 << 28); ax[i] = 0; base[i] = p.dy; }
      }
    }
    __device__ void load(typename Tiles<TM, TN>::A& sA, int tid, int k0, int k_end) const {
      const ConvGeo& g = p.g;
      const int k = k0 + kk;
      const bool kv = k < k_end;
      const int co = k % g.Co, tap = k / g.Co, jx = tap % p.Jx, jy = tap / p.Jx;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int oy = ay[i] - jy, ox = ax[i] - jx;
        float v = 0.f;
        if (kv && oy >= 0 && oy < g.Ho && ox >= 0 && ox < g.Wo) v = base[i][((long long)oy * g.Wo + ox) * g.y_cs + co];
        sA[kk][(tid >> 4) + 16 * i] = v;
      }
    }
  };
  template <int TM, int TN>
  struct BLoader {
    const BwdDataClassProblem& p; int kk; int n0;
    __device__ BLoader(const BwdDataClassProblem& p_, int n0_, int tid) : p(p_), kk(tid & 15), n0(n0_) {}
    __device__ void load(typename Tiles<TM, TN>::B& sB, int tid, int k0, int k_end) const {
      const ConvGeo& g = p.g;
      const int k = k0 + kk;
      const bool kv = k < k_end;
      const int co = k % g.Co, tap = k / g.Co, jx = tap % p.Jx, jy = tap / p.Jx;
      const int ky = p.ry + g.stride * jy, kx = p.rx + g.stride * jx;
#pragma unroll
      for (int i = 0; i < Tiles<TM, TN>::RBK; ++i) {
        const int nn = (tid >> 4) + 16 * i, n = n0 + nn;
        sB[kk][nn] = (kv && n < g.Ci) ? p.w[(((long long)ky * g.kw + kx) * g.Ci + n) * g.Co + co] : 0.f;
      }
    }
  };
  __device__ void store(int m, int n, float v) const {
    const int b = m % Wc, a = (m / Wc) % Hc, img = m / (Wc * Hc);
    const int iy = fy + g.stride * a, ix = fx + g.stride * b;
    float* q = dx + (((long long)img * g.H + iy) * g.W + ix) * g.x_cs + g.x_co + n;
    *q = accumulate ? *q + v : v;
  }
};

struct BwdWeightProblem {  // dw[tap][ci][co] = sum over pixels x * dy; split over K (pixels) into partial buffers
  ConvGeo g; const float* x; const float* dy; float* part; int k_per_split;
  __device__ int M() const { return g.kh * g.kw * g.Ci; }
  __device__ int Nn() const { return g.Co; }
  __device__ int K() const { return g.N * g.Ho * g.Wo; }
  template <int TM, int TN>
  struct ALoader {        // M-fastest: m = (tap, ci), ci contiguous in x
    const BwdWeightProblem& p; int ky, kx, ci; bool mv;
    __device__ ALoader(const BwdWeightProblem& p_, int m0, int tid) : p(p_) {
      const ConvGeo& g = p.g;
      const int m = m0 + tid % TM;
      mv = m < p.M();
      ci = m % g.Ci;
      const int tap = m / g.Ci;
      kx = tap % g.kw; ky = tap / g.kw;
    }
    __device__ void load(typename Tiles<TM, TN>::A& sA, int tid, int k0, int k_end) const {
      const ConvGeo& g = p.g;
#pragma unroll
      for (int i = 0; i < Tiles<TM, TN>::RAM; ++i) {
        const int kk = tid / TM + Tiles<TM, TN>::KSTEP_M * i, k = k0 + kk;
        float v = 0.f;
        if (mv && k < k_end) {
          const int ox = k % g.Wo, oy = (k / g.Wo) % g.Ho, n = k / (g.Wo * g.Ho);
          const int iy = oy * g.stride - g.pad_t + ky, ix = ox * g.stride - g.pad_l + kx;
          if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) v = p.x[(((long long)n * g.H + iy) * g.W + ix) * g.x_cs + g.x_co + ci];
        }
        sA[kk][tid % TM] = v;
      }
    }
  };
  template <int TM, int TN>
  struct BLoader {        // N-fastest: dy[k][co]
    const BwdWeightProblem& p; int n; bool nv;
    __device__ BLoader(const BwdWeightProblem& p_, int n0, int tid) : p(p_), n(n0 + tid % TN), nv(n0 + tid % TN < p_.g.Co) {}
    __device__ void load(typename Tiles<TM, TN>::B& sB, int tid, int k0, int k_end) const {
#pragma unroll
      for (int i = 0; i < Tiles<TM, TN>::RBN; ++i) {
        const int kk = tid / TN + Tiles<TM, TN>::KSTEP_N * i, k = k0 + kk;
        sB[kk][tid % TN] = (nv && k < k_end) ? p.dy[(long long)k * p.g.y_cs + p.g.y_co + n] : 0.f;
      }
    }
  };
};

template <class P, int TM, int TN>
__device__ __forceinline__ void gemm_tile(const P& p, int m0, int n0, int k_begin, int k_end, float (&acc)[TM * TN / 1024][4]) {
  constexpr int RM = TM * TN / 1024;            // rows per thread (columns per thread = 4)
  __shared__ __align__(16) typename Tiles<TM, TN>::A sA;
  __shared__ __align__(16) typename Tiles<TM, TN>::B sB;
  const int tid = threadIdx.x;
  const int tm = (tid / (TN / 4)) * RM, tn = (tid % (TN / 4)) * 4;
  const typename P::template ALoader<TM, TN> la(p, m0, tid);
  const typename P::template BLoader<TM, TN> lb(p, n0, tid);
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    la.load(sA, tid, k0, k_end);
    lb.load(sB, tid, k0, k_end);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[RM];
      if constexpr (RM == 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sA[kk][tm]);
        a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
      } else {
        const float2 a2 = *reinterpret_cast<const float2*>(&sA[kk][tm]);
        a[0] = a2.x; a[1] = a2.y;
      }
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[kk][tn]);
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

template <class P, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_kernel(P p) {
  constexpr int RM = TM * TN / 1024;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  float acc[RM][4] = {};
  gemm_tile<P, TM, TN>(p, m0, n0, 0, p.K(), acc);
  const int tm = (threadIdx.x / (TN / 4)) * RM, tn = (threadIdx.x % (TN / 4)) * 4;
  const int M = p.M(), N = p.Nn();
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (m0 + tm + i < M && n0 + tn + j < N) p.store(m0 + tm + i, n0 + tn + j, acc[i][j]);
}

template <int TM, int TN>
__global__ void __launch_bounds__(256) gemm_wgrad_kernel(BwdWeightProblem p) {
  constexpr int RM = TM * TN / 1024;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN, split = blockIdx.z;
  const int K = p.K();
  const int kb = split * p.k_per_split, ke = min(K, kb + p.k_per_split);
  float acc[RM][4] = {};
  gemm_tile<BwdWeightProblem, TM, TN>(p, m0, n0, kb, ke, acc);
  const int tm = (threadIdx.x / (TN / 4)) * RM, tn = (threadIdx.x % (TN / 4)) * 4;
  const int M = p.M(), N = p.Nn();
  float* out = p.part + (long long)split * M * N;
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (m0 + tm + i < M && n0 + tn + j < N) out[(long long)(m0 + tm + i) * N + n0 + tn + j] = acc[i][j];
}

// N <= 16 wide problems use the 128x16 tile
template <class P>
static cudaError_t launch_gemm(const P& p, long long M, int N, cudaStream_t stream) {
  if (N <= 16) {
    dim3 grid((unsigned)((M + 127) / 128), (N + 15) / 16);
    gemm_kernel<P, 128, 16><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid((unsigned)((M + 63) / 64), (N + 63) / 64);
    gemm_kernel<P, 64, 64><<<grid, 256, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

// dst[i] (+)= sum_s part[s * stride + i]   (fixed order: deterministic)
__global__ void reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ dst, long long n, int splits,
                                     int accumulate, long long stride = 0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (stride == 0) stride = n;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(long long)k * stride + i];
  dst[i] = accumulate ? dst[i] + s : s;
}

// ------------------------------------------------------------------ column reductions over rows of [R][cs] (+co)
// One block per 32 channels x row-slab (8 row lanes, 4 rows in flight per lane); partials part[slab][2][C] are then
// reduced in slab order (deterministic).  MODE 0: s1 = sum a   1: s1 = sum a*b   2: s1 = sum a*a
// 3: s1 = sum a, s2 = sum a*a (BatchNorm statistics)   4: s1 = sum a*(b - mean)*invstd, s2 = sum a (BatchNorm backward)
constexpr int CS_SLABS = 256;
template <int MODE>
__global__ void colsum_partial_kernel(const float* __restrict__ a, int a_cs, int a_co, const float* __restrict__ b, int b_cs,
                                      int b_co, const float* __restrict__ mean, const float* __restrict__ invstd, long long R,
                                      int C, float* __restrict__ part) {
  __shared__ float sm[2][8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long rows_per = (R + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  float s1 = 0.f, s2 = 0.f;
  if (c < C) {
    float mu = 0.f, is = 0.f;
    if (MODE == 4) { mu = mean[c]; is = invstd[c]; }
    for (long long r = r0 + threadIdx.y; r < r1; r += 32) {
      float av[4], bv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long rr = r + 8 * j;
        av[j] = rr < r1 ? a[rr * a_cs + a_co + c] : 0.f;
        if (MODE == 1 || MODE == 4) bv[j] = rr < r1 ? b[rr * b_cs + b_co + c] : mu;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 0) s1 += av[j];
        else if (MODE == 1) s1 += av[j] * bv[j];
        else if (MODE == 2) s1 += av[j] * av[j];
        else if (MODE == 3) { s1 += av[j]; s2 += av[j] * av[j]; }
        else { s1 += av[j] * (bv[j] - mu) * is; s2 += av[j]; }
      }
    }
  }
  sm[0][threadIdx.y][threadIdx.x] = s1;
  sm[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y < 2 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[threadIdx.y][i][threadIdx.x];
    part[((long long)blockIdx.y * 2 + threadIdx.y) * C + c] = t;
  }
}

// ------------------------------------------------------------------ elementwise helpers
__global__ void leaky_fwd_kernel(float* __restrict__ x, long long n, float alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; x[i] = v >= 0.f ? v : alpha * v; }
}
// dx = dy * (y >= 0 ? 1 : alpha), in place on dy; y is the activation OUTPUT (same sign as its input)
__global__ void leaky_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long long n, float alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && y[i] < 0.f) dy[i] *= alpha;
}
// out = a*x + b*y (y may be null), with channel stride/offset on every operand: generic strided copy / axpby
__global__ void axpby_kernel(float* __restrict__ out, int o_cs, int o_co, const float* __restrict__ x, int x_cs, int x_co,
                             float a, const float* __restrict__ y, int y_cs, int y_co, float b, long long rows, int C,
                             int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i % C);
  float v = a * x[r * x_cs + x_co + c];
  if (y) v += b * y[r * y_cs + y_co + c];
  float* o = out + r * o_cs + o_co + c;
  *o = accumulate ? *o + v : v;
}
// x = act(x + bias[c]) in place on channels [co, co+C) of a buffer with pixel pitch cs; alpha = 1 -> linear
__global__ void bias_act_kernel(float* __restrict__ x, int cs, int co, const float* __restrict__ bias, long long rows, int C,
                                float alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int c = (int)(i % C);
  float* p = x + (i / C) * cs + co + c;
  const float v = *p + (bias ? bias[c] : 0.f);
  *p = v >= 0.f ? v : alpha * v;
}
// out[b][a][inner] = in[a][b][inner]
__global__ void transpose01_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int B, long long inner) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A * B * inner) return;
  const long long k = i % inner;
  const long long ab = i / inner;
  const int b = (int)(ab % B), a = (int)(ab / B);
  out[((long long)b * A + a) * inner + k] = in[i];
}
// Overlap-add of per-pixel tap columns: the stride-1 transposed convolution y[n,iy,ix,o] = sum_{ky,kx,c} x[n, iy+p-ky,
// ix+p-kx, c] w[ky,kx,o,c] computed as ONE 1x1 GEMM cols[pix][(ky,kx,o)] = x[pix][:] . w[ky,kx,o,:] (no tap
// re-reads of the wide input: a 5x5 gather over 160 channels costs 25x its input in L2 traffic) followed by
// y[n,iy,ix,o] = sum_{ky,kx} cols[(n, iy+p-ky, ix+p-kx)][(ky,kx,o)].  One thread per (pixel, 4 channels).
__global__ void col2im_kernel(const float* __restrict__ cols, float* __restrict__ out, long long total, int H, int W, int kh,
                              int kw, int Co, int pad, int o_cs, int o_co) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C4 = (Co + 3) / 4;
  const int c0 = (int)(i % C4) * 4;
  const long long pix = i / C4;
  const int ix = (int)(pix % W), iy = (int)((pix / W) % H);
  const long long img = pix / ((long long)W * H);
  const int ld = kh * kw * Co;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool v4 = (Co & 3) == 0;
  for (int ky = 0; ky < kh; ++ky) {
    const int y = iy + pad - ky;
    if (y < 0 || y >= H) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int x = ix + pad - kx;
      if (x < 0 || x >= W) continue;
      const float* q = cols + ((img * H + y) * W + x) * ld + (ky * kw + kx) * Co + c0;
      if (v4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(q));
        acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
      } else {
        for (int e = 0; e < 4 && c0 + e < Co; ++e) acc[e] += q[e];
      }
    }
  }
  float* o = out + pix * o_cs + o_co + c0;
  for (int e = 0; e < 4 && c0 + e < Co; ++e) o[e] = acc[e];
}
// combined[b,...] = eps[b]*real + (1-eps[b])*fake   (ganbase.py:31)
__global__ void lerp_batch_kernel(float* __restrict__ out, const float* __restrict__ real, const float* __restrict__ fake,
                                  const float* __restrict__ eps, long long per_sample, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = eps[i / per_sample];
  out[i] = e * real[i] + (1.f - e) * fake[i];
}

// ------------------------------------------------------------------ BatchNorm (training) / LayerNorm
// y = (x - mean[c]) * invstd[c] * gamma[c] + beta[c]   (mean/invstd supplied: batch or moving statistics)
__global__ void bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, long long rows, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int c = (int)(i % C);
  y[i] = (x[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
}
// stats from sums: mean = s1/R, var = s2/R - mean^2 (biased); moving stats update (momentum, Bessel-corrected var)
__global__ void bn_finalize_kernel(const float* __restrict__ s1, const float* __restrict__ s2, long long R, int C, float eps,
                                   float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = (double)s1[c] / (double)R;
  double v = (double)s2[c] / (double)R - m * m;
  if (v < 0) v = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(v + (double)eps));
  if (moving_mean) {
    moving_mean[c] = moving_mean[c] * momentum + (float)m * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + (float)(v * ((double)R / (double)(R - 1))) * (1.f - momentum);
  }
}
// dx = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat)), with sums dbeta = sum dy, dgamma = sum dy*xhat supplied
__global__ void bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                              const float* __restrict__ dbeta, const float* __restrict__ dgamma, float* __restrict__ dx,
                              long long rows, int C, long long rows_global, float act_alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int c = (int)(i % C);
  const float xv = x[i];
  const float xhat = (xv - mean[c]) * invstd[c];
  const float inv_r = 1.f / (float)rows_global;
  const float d = gamma[c] * invstd[c] * (dy[i] - dbeta[c] * inv_r - xhat * dgamma[c] * inv_r);
  // act_alpha != 1: x is the output of a LeakyReLU(act_alpha); its backward is folded in here
  dx[i] = xv >= 0.f ? d : act_alpha * d;
}

// LayerNorm over the channel axis: LPR lanes per pixel (LPR = 4 for the critic's 16-channel full-resolution layers, so a
// warp handles 8 pixels and no lane idles); saves mean and invstd per pixel.
template <int LPR>
__global__ void ln_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ invstd,
                              long long rows, int C, float eps, int y_cs, int y_co) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  const bool ok = r < rows;
  const float* xr = x + (ok ? r : 0) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += LPR) s += xr[c];
#pragma unroll
  for (int o = LPR / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
  const float m = s / C;
  float v = 0.f;
  for (int c = lane; c < C; c += LPR) { const float d = xr[c] - m; v += d * d; }
#pragma unroll
  for (int o = LPR / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  const float is = rsqrtf(v / C + eps);
  if (!ok) return;
  if (lane == 0 && mean) { mean[r] = m; invstd[r] = is; }
  for (int c = lane; c < C; c += LPR) y[r * y_cs + y_co + c] = (xr[c] - m) * is * gamma[c] + beta[c];
}
// dx = invstd * (g - mean_c(g) - xhat * mean_c(g * xhat)), g = dy * gamma; also emits dy*xhat for dgamma
template <int LPR>
__global__ void ln_bwd_kernel(const float* __restrict__ dy, int dy_cs, int dy_co, const float* __restrict__ x,
                              const float* __restrict__ gamma, const float* __restrict__ mean,
                              const float* __restrict__ invstd, float* __restrict__ dx, float* __restrict__ dy_xhat,
                              long long rows, int C, float act_alpha) {
  const long long r0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  const bool ok = r0 < rows;
  const long long r = ok ? r0 : 0;
  const float m = mean[r], is = invstd[r];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += LPR) {
    const float g = dy[r * dy_cs + dy_co + c] * gamma[c];
    const float xh = (x[r * C + c] - m) * is;
    s1 += g; s2 += g * xh;
  }
#pragma unroll
  for (int o = LPR / 2; o; o >>= 1) { s1 += __shfl_xor_sync(0xffffffff, s1, o); s2 += __shfl_xor_sync(0xffffffff, s2, o); }
  if (!ok) return;
  s1 /= C; s2 /= C;
  for (int c = lane; c < C; c += LPR) {
    const float d = dy[r * dy_cs + dy_co + c];
    const float xv = x[r * C + c];
    const float xh = (xv - m) * is;
    const float dd = is * (d * gamma[c] - s1 - xh * s2);
    dx[r * C + c] = xv >= 0.f ? dd : act_alpha * dd;       // LeakyReLU backward of the layer that produced x, folded in
    dy_xhat[r * C + c] = d * xh;
  }
}
static inline int ln_lanes(int C) { return C <= 16 ? 4 : (C <= 32 ? 8 : (C <= 64 ? 16 : 32)); }

// ------------------------------------------------------------------ ConvLSTM gate math (gates i, f, c~, o on the last axis)
// z: [rows][4F] pre-activations (x-conv + bias + h-conv).  Saves the activated gates in place of z for the backward.
__global__ void lstm_gates_fwd_kernel(float* __restrict__ z, const float* __restrict__ c_prev, float* __restrict__ c_out,
                                      float* __restrict__ h_out, long long rows, int F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const long long r = i / F;
  const int c = (int)(i % F);
  float* zr = z + r * 4 * F;
  const float gi = fminf(fmaxf(0.2f * zr[c] + 0.5f, 0.f), 1.f);
  const float gf = fminf(fmaxf(0.2f * zr[F + c] + 0.5f, 0.f), 1.f);
  const float gc = tanhf(zr[2 * F + c]);
  const float go = fminf(fmaxf(0.2f * zr[3 * F + c] + 0.5f, 0.f), 1.f);
  const float cp = c_prev ? c_prev[i] : 0.f;
  const float cn = gf * cp + gi * gc;
  zr[c] = gi; zr[F + c] = gf; zr[2 * F + c] = gc; zr[3 * F + c] = go;
  c_out[i] = cn;
  h_out[i] = go * tanhf(cn);
}
// Backward of one step.  gates: saved activated gates; dh: dL/dh_t (from the output and from step t+1);
// dc: in = dL/dc_t carried from step t+1, out = dL/dc_{t-1}.  Writes dz (pre-activation gradient) over `gates`.
// dh_rec (optional): the recurrent contribution dL/dh_t carried from step t+1, added here instead of by a separate pass.
__global__ void lstm_gates_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_cur,
                                      const float* __restrict__ dh, const float* __restrict__ dh_rec, float* __restrict__ dc,
                                      long long rows, int F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const long long r = i / F;
  const int c = (int)(i % F);
  float* g = gates + r * 4 * F;
  const float gi = g[c], gf = g[F + c], gc = g[2 * F + c], go = g[3 * F + c];
  const float tc = tanhf(c_cur[i]);
  const float dhv = dh_rec ? dh[i] + dh_rec[i] : dh[i];
  const float dct = dc[i] + dhv * go * (1.f - tc * tc);
  const float cp = c_prev ? c_prev[i] : 0.f;
  // hard_sigmoid'(z) = 0.2 strictly inside (0, 1)
  const float di = dct * gc * ((gi > 0.f && gi < 1.f) ? 0.2f : 0.f);
  const float df = dct * cp * ((gf > 0.f && gf < 1.f) ? 0.2f : 0.f);
  const float dg = dct * gi * (1.f - gc * gc);
  const float dov = dhv * tc * ((go > 0.f && go < 1.f) ? 0.2f : 0.f);
  g[c] = di; g[F + c] = df; g[2 * F + c] = dg; g[3 * F + c] = dov;
  dc[i] = dct * gf;
}

// ConvLSTM cells with very few filters (the critic's high-resolution branch: ConvLSTM2D(2) on 96x96 images): the
// recurrent 3x3 convolution has K = 9 F <= 36 and N = 4 F <= 16, a bandwidth-bound stencil rather than a GEMM, so one
// thread per pixel does the recurrent convolution AND the gate math in one pass (fp32, whatever the GEMM precision).
// z in: x-conv + bias of this step; out: activated gates (as lstm_gates_fwd_kernel leaves them).  R: [3][3][F][4F].
template <int F>
__global__ void lstm_small_fwd_kernel(float* __restrict__ z, const float* __restrict__ h_prev, const float* __restrict__ R,
                                      const float* __restrict__ c_prev, float* __restrict__ c_out, float* __restrict__ h_out,
                                      int N, int H, int W) {
  __shared__ float sR[9 * F * 4 * F];
  for (int i = threadIdx.x; i < 9 * F * 4 * F; i += blockDim.x) sR[i] = R[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * H * W) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  const long long img = idx / ((long long)W * H);
  float acc[4 * F];
#pragma unroll
  for (int g = 0; g < 4 * F; ++g) acc[g] = z[idx * 4 * F + g];
  if (h_prev) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float* hp = h_prev + ((img * H + yy) * W + xx) * F;
#pragma unroll
        for (int ci = 0; ci < F; ++ci) {
          const float hv = hp[ci];
#pragma unroll
          for (int g = 0; g < 4 * F; ++g) acc[g] = fmaf(hv, sR[((ky * 3 + kx) * F + ci) * 4 * F + g], acc[g]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < F; ++c) {
    const float gi = fminf(fmaxf(0.2f * acc[c] + 0.5f, 0.f), 1.f);
    const float gf = fminf(fmaxf(0.2f * acc[F + c] + 0.5f, 0.f), 1.f);
    const float gc = tanhf(acc[2 * F + c]);
    const float go = fminf(fmaxf(0.2f * acc[3 * F + c] + 0.5f, 0.f), 1.f);
    const float cp = c_prev ? c_prev[idx * F + c] : 0.f;
    const float cn = gf * cp + gi * gc;
    z[idx * 4 * F + c] = gi; z[idx * 4 * F + F + c] = gf; z[idx * 4 * F + 2 * F + c] = gc; z[idx * 4 * F + 3 * F + c] = go;
    c_out[idx * F + c] = cn;
    h_out[idx * F + c] = go * tanhf(cn);
  }
}
// dh_rec[n, iy, ix, ci] = sum_{ky, kx, g} dz[n, iy + 1 - ky, ix + 1 - kx, g] * R[ky][kx][ci][g]
template <int F>
__global__ void lstm_small_bwd_data_kernel(const float* __restrict__ dz, const float* __restrict__ R, float* __restrict__ dh_rec,
                                           int N, int H, int W) {
  __shared__ float sR[9 * F * 4 * F];
  for (int i = threadIdx.x; i < 9 * F * 4 * F; i += blockDim.x) sR[i] = R[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * H * W) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  const long long img = idx / ((long long)W * H);
  float acc[F];
#pragma unroll
  for (int c = 0; c < F; ++c) acc[c] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int oy = y + 1 - ky;
    if (oy < 0 || oy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ox = x + 1 - kx;
      if (ox < 0 || ox >= W) continue;
      const float* d = dz + ((img * H + oy) * W + ox) * 4 * F;
#pragma unroll
      for (int g = 0; g < 4 * F; ++g) {
        const float dv = d[g];
#pragma unroll
        for (int ci = 0; ci < F; ++ci) acc[ci] = fmaf(dv, sR[((ky * 3 + kx) * F + ci) * 4 * F + g], acc[ci]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < F; ++c) dh_rec[idx * F + c] = acc[c];
}

// ------------------------------------------------------------------ bilinear x2 (half-pixel centres, edge clamp) and adjoint
__device__ __forceinline__ void up_taps(int j, int n, int& a, int& b, float& wa) {
  const int k = j >> 1;
  if (j & 1) { a = k; b = min(k + 1, n - 1); wa = 0.75f; }
  else { a = max(k - 1, 0); b = k; wa = 0.25f; }
}
__global__ void upsample2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n_img, int h, int w, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = n_img * 4 * h * w * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pix = i / C;
  const int X = (int)(pix % (2 * w)), Y = (int)((pix / (2 * w)) % (2 * h));
  const long long n = pix / ((long long)4 * w * h);
  int ya, yb, xa, xb; float wy, wx;
  up_taps(Y, h, ya, yb, wy); up_taps(X, w, xa, xb, wx);
  const float* p = x + n * h * w * C + c;
  y[i] = wy * (wx * p[((long long)ya * w + xa) * C] + (1.f - wx) * p[((long long)ya * w + xb) * C]) +
         (1.f - wy) * (wx * p[((long long)yb * w + xa) * C] + (1.f - wx) * p[((long long)yb * w + xb) * C]);
}
// adjoint: dx[k] gathers from the (at most 3x3) hi-res positions that read it
__global__ void upsample2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long n_img, int h, int w, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = n_img * h * w * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pix = i / C;
  const int kx = (int)(pix % w), ky = (int)((pix / w) % h);
  const long long n = pix / ((long long)w * h);
  const float* p = dy + n * 4 * h * w * C + c;
  float acc = 0.f;
  for (int Y = max(2 * ky - 2, 0); Y <= min(2 * ky + 3, 2 * h - 1); ++Y) {
    int ya, yb; float wy;
    up_taps(Y, h, ya, yb, wy);
    const float cy = (ya == ky ? wy : 0.f) + (yb == ky ? 1.f - wy : 0.f);
    if (cy == 0.f) continue;
    for (int X = max(2 * kx - 2, 0); X <= min(2 * kx + 3, 2 * w - 1); ++X) {
      int xa, xb; float wx;
      up_taps(X, w, xa, xb, wx);
      const float cx = (xa == kx ? wx : 0.f) + (xb == kx ? 1.f - wx : 0.f);
      if (cx != 0.f) acc += cy * cx * p[((long long)Y * 2 * w + X) * C];
    }
  }
  dx[i] = acc;
}

// float4 versions (C % 4 == 0): one thread per (output pixel, 4 channels), 32-bit index math
__global__ void upsample2x_fwd_v4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int n_img, int h, int w, int C4) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned total = (unsigned)n_img * 4u * h * w * C4;
  if (i >= total) return;
  const int c = i % C4;
  const unsigned pix = i / C4;
  const int X = pix % (2 * w), Y = (pix / (2 * w)) % (2 * h);
  const int n = pix / (4u * w * h);
  int ya, yb, xa, xb; float wy, wx;
  up_taps(Y, h, ya, yb, wy); up_taps(X, w, xa, xb, wx);
  const float4* p = x + (long long)n * h * w * C4 + c;
  const float4 aa = __ldg(p + (ya * w + xa) * C4), ab = __ldg(p + (ya * w + xb) * C4);
  const float4 ba = __ldg(p + (yb * w + xa) * C4), bb = __ldg(p + (yb * w + xb) * C4);
  float4 o;
  o.x = wy * (wx * aa.x + (1.f - wx) * ab.x) + (1.f - wy) * (wx * ba.x + (1.f - wx) * bb.x);
  o.y = wy * (wx * aa.y + (1.f - wx) * ab.y) + (1.f - wy) * (wx * ba.y + (1.f - wx) * bb.y);
  o.z = wy * (wx * aa.z + (1.f - wx) * ab.z) + (1.f - wy) * (wx * ba.z + (1.f - wx) * bb.z);
  o.w = wy * (wx * aa.w + (1.f - wx) * ab.w) + (1.f - wy) * (wx * ba.w + (1.f - wx) * bb.w);
    y[i] = o;
}
// adjoint per axis: dx[k] = .25 dy[2k-1] (k >= 1) + (.75 + .25 [k == 0]) dy[2k] + (.75 + .25 [k == n-1]) dy[2k+1] + .25 dy[2k+2] (k <= n-2)
__device__ __forceinline__ void up_adj_taps(int k, int n, float (&wt)[4]) {
  wt[0] = k >= 1 ? 0.25f : 0.f;
  wt[1] = k == 0 ? 1.f : 0.75f;
  wt[2] = k == n - 1 ? 1.f : 0.75f;
  wt[3] = k <= n - 2 ? 0.25f : 0.f;
}
__global__ void upsample2x_bwd_v4_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, int n_img, int h, int w, int C4) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned total = (unsigned)n_img * h * w * C4;
  if (i >= total) return;
  const int c = i % C4;
  const unsigned pix = i / C4;
  const int kx = pix % w, ky = (pix / w) % h;
  const int n = pix / ((unsigned)w * h);
  const float4* p = dy + (long long)n * 4 * h * w * C4 + c;
  float wy[4], wx[4];
  up_adj_taps(ky, h, wy); up_adj_taps(kx, w, wx);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (wy[a] == 0.f) continue;
    const int Y = 2 * ky - 1 + a;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wx[b] == 0.f) continue;
      const int X = 2 * kx - 1 + b;
      const float4 v = __ldg(p + (Y * 2 * w + X) * C4);
      const float cw = wy[a] * wx[b];
      acc.x += cw * v.x; acc.y += cw * v.y; acc.z += cw * v.z; acc.w += cw * v.w;
    }
  }
  dx[i] = acc;
}

// ------------------------------------------------------------------ Dense(1) + temporal mean, and their backward
// score[b] = mean_t (flat[b,t,:] . w + bias)
__global__ void dense_mean_fwd_kernel(const float* __restrict__ flat, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ score, int T, int D) {
  __shared__ float sm[256];
  const int b = blockIdx.x;
  float s = 0.f;
  for (long long i = threadIdx.x; i < (long long)T * D; i += blockDim.x) s += flat[(long long)b * T * D + i] * w[i % D];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) score[b] = sm[0] / T + bias[0];
}
// dflat[b,t,d] = dscore[b] * w[d] / T ;  (dw, dbias by separate reductions on the host side of the ABI)
__global__ void dense_mean_bwd_kernel(const float* __restrict__ dscore, const float* __restrict__ w, float* __restrict__ dflat,
                                      long long B, int T, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T * D) return;
  dflat[i] = dscore[i / ((long long)T * D)] * w[i % D] / T;
}
// dw[d] = sum_{b,t} dscore[b]/T * flat[b,t,d]
__global__ void dense_mean_wgrad_kernel(const float* __restrict__ dscore, const float* __restrict__ flat, float* __restrict__ dw,
                                        float* __restrict__ dbias, int B, int T, int D) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) {
    float s = 0.f;
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < T; ++t) s += dscore[b] / T * flat[((long long)b * T + t) * D + d];
    dw[d] = s;
  }
  if (d == 0) { float s = 0.f; for (int b = 0; b < B; ++b) s += dscore[b]; dbias[0] = s; }
}

// ------------------------------------------------------------------ generic full reduction: out[0] = sum f(a[,b])
template <int MODE>   // 0 sum a, 1 sum a*b, 2 sum a*a
__global__ void reduce_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ part) {
  __shared__ double sm[256];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double av = a[i];
    s += MODE == 0 ? av : (MODE == 1 ? av * (double)b[i] : av * av);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
}
__global__ void reduce_final_kernel(const double* __restrict__ part, int n, float* __restrict__ out, double scale) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { double s = 0; for (int i = 0; i < n; ++i) s += part[i]; out[0] = (float)(s * scale); }
}
// gradient-penalty norms (ganbase.py:36): out[b*C + c] = sqrt(sum_{t,h,w} g[b,t,h,w,c]^2)
__global__ void gp_norm_kernel(const float* __restrict__ g, float* __restrict__ out, long long per_sample_px, int C) {
  __shared__ double sm[256];
  const int b = blockIdx.x, c = blockIdx.y;
  double s = 0.0;
  for (long long i = threadIdx.x; i < per_sample_px; i += blockDim.x) {
    const double v = g[((long long)b * per_sample_px + i) * C + c];
    s += v * v;
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[b * C + c] = (float)sqrt(sm[0]);
}

// ------------------------------------------------------------------ Keras Adam, spectral normalisation
// m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= lr_t * m / (sqrt(v) + eps)       (train.py:35,58; eps outside the sqrt)
__global__ void adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                            long long n, float lr_t, float b1, float b2, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.f - b1);
  const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
  m[i] = mi; v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}
// Graph-capturable form: the step counter and the bias-corrected learning rate live in device memory.
__global__ void adam_lr_kernel(float* lr_t, int* step, float lr, float b1, float b2) {
  const int t = ++(*step);
  *lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
}
__global__ void adam_dev_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                                long long n, const float* __restrict__ lr_t_dev, float b1, float b2, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr_t = *lr_t_dev;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.f - b1);
  const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
  m[i] = mi; v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}
// One power iteration of TFA SpectralNormalization on W = reshape(w, (R, C)):  v = l2n(u W^T); u' = l2n(v W);
// sigma = v W u'^T; w /= sigma; u = u'.  Four small grid-wide kernels, every reduction in a fixed order
// (deterministic, so data-parallel replicas stay bit-identical).  scratch: R + SN_SLABS*C + 4 floats.
constexpr int SN_SLABS = 64;
// v_raw[r] = sum_c u[c] W[r][c]: one warp per row
__global__ void sn_rowdot_kernel(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ vbuf, int R, int C) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += u[c] * w[(long long)r * C + c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) vbuf[r] = s;
}
// part[slab][c] = sum_{r in slab} v_raw[r] W[r][c]   (32 channels x 8 row lanes per block)
__global__ void sn_coldot_kernel(const float* __restrict__ w, const float* __restrict__ vbuf, float* __restrict__ part, int R, int C) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  float s = 0.f;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += vbuf[r] * w[(long long)r * C + c];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    part[(long long)blockIdx.y * C + c] = t;
  }
}
// single block: |v_raw|, u2_raw = (sum of slabs) / |v_raw|, |u2_raw|, u = l2n(u2_raw), scal[0] = 1 / sigma
__global__ void sn_finish_kernel(const float* __restrict__ vbuf, const float* __restrict__ part, float* __restrict__ u,
                                 float* __restrict__ scal, int R, int C, int slabs) {
  __shared__ float red[256];
  __shared__ float bc;
  const int tid = threadIdx.x;
  float nv = 0.f;
  for (int r = tid; r < R; r += 256) nv += vbuf[r] * vbuf[r];
  red[tid] = nv; __syncthreads();
  for (int o = 128; o; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  if (tid == 0) bc = rsqrtf(fmaxf(red[0], 1e-12f));
  __syncthreads();
  const float inv_nv = bc;
  __syncthreads();
  float nu = 0.f;
  for (int c = tid; c < C; c += 256) {
    float s = 0.f;
    for (int k = 0; k < slabs; ++k) s += part[(long long)k * C + c];
    s *= inv_nv;
    u[c] = s;
    nu += s * s;
  }
  red[tid] = nu; __syncthreads();
  for (int o = 128; o; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  const float inv_nu = rsqrtf(fmaxf(red[0], 1e-12f));
  // sigma = v W u'^T = sum_c u2_raw[c] u'[c] = |u2_raw|^2 * inv_nu
  const float sigma = red[0] * inv_nu;
  for (int c = tid; c < C; c += 256) u[c] *= inv_nu;
  if (tid == 0) scal[0] = 1.f / sigma;
}
__global__ void sn_scale_kernel(float* __restrict__ w, const float* __restrict__ scal, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] *= scal[0];
}

__global__ void rsqrt_eps_kernel(const float* __restrict__ v, float* __restrict__ o, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) o[c] = 1.f / sqrtf(v[c] + eps);
}

inline unsigned blocks_for(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace

// ====================================================================== C ABI
// Arithmetic of the convolution GEMMs: 0 = fp32 on CUDA cores (this file), 1 = tf32 / 2 = bf16 operands on tcgen05
// with fp32 accumulation (train_gemm_tc.cu).  Process-wide, like the error string; set between steps.
static int g_train_precision = 0;
extern "C" int wdg_train_set_precision(int mode) {
  if (mode < 0 || mode > 2) return wdg_set_error("wdg_train_set_precision: mode must be 0 (fp32), 1 (tf32) or 2 (bf16)");
  g_train_precision = mode;
  return 0;
}
extern "C" int wdg_train_get_precision(void) { return g_train_precision; }

// ---- direct kernels for narrow 3x3 / stride-1 / same convolutions (train_direct.cuh); WDG_NO_DIRECT=1 disables them
// Which (Ci, Co) pairs use them was decided per GEMM kind from measurements at 192 x 96 x 96 (tools/bench_tc_conv.py,
// ms direct vs tcgen05 tf32): forward 2->8 0.046 / 0.270, 2->16 0.128 / 0.264 (16->2 0.484 / 0.363: not used);
// backward-data 2->8 0.136 / 0.328, 16->2 0.123 / 0.298 (2->16 0.483 / 0.396: not used); backward-weight 2->8
// 0.239 / 0.469, 2->16 0.354 / 0.558 (16->2 0.991 / 0.613: not used).  5->64 and 16->16 lose everywhere (FMA-bound).
#define WDG_DIRECT_PAIRS(X) X(2, 8) X(2, 16) X(16, 2)
enum { DIRECT_FWD = 0, DIRECT_BWD_DATA = 1, DIRECT_BWD_WEIGHT = 2 };
static bool direct_ok(const ConvGeo& g, int kind) {
  static int off = -1;
  if (off < 0) off = getenv("WDG_NO_DIRECT") ? 1 : 0;
  if (off || g.kh != 3 || g.kw != 3 || g.stride != 1 || g.pad_t != 1 || g.pad_l != 1 || g.Ho != g.H || g.Wo != g.W) return false;
  if (g.Ci == 2 && g.Co == 8) return true;
  if (g.Ci == 2 && g.Co == 16) return kind != DIRECT_BWD_DATA;
  if (g.Ci == 16 && g.Co == 2) return kind == DIRECT_BWD_DATA;
  return false;
}
static int direct_fwd(const ConvGeo& g, const float* x, const float* w, const float* bias, float* y, int accumulate, float alpha,
                      cudaStream_t st) {
  const long long npix = (long long)g.N * g.H * g.W;
#define X(ci, co)                                                                                                            \
  if (g.Ci == ci && g.Co == co)                                                                                              \
    wdg_direct::conv3x3_fwd_kernel<ci, co><<<blocks_for(npix), 256, 0, st>>>(x, g.x_cs, g.x_co, w, bias, y, g.y_cs, g.y_co, npix, \
                                                                            g.H, g.W, accumulate, alpha);
  WDG_DIRECT_PAIRS(X)
#undef X
  CKT(cudaGetLastError());
  return 0;
}
static int direct_bwd_data(const ConvGeo& g, const float* dy, const float* w, float* dx, int accumulate, cudaStream_t st) {
  const long long npix = (long long)g.N * g.H * g.W;
#define X(ci, co)                                                                                                              \
  if (g.Ci == ci && g.Co == co)                                                                                                \
    wdg_direct::conv3x3_bwd_data_kernel<ci, co><<<blocks_for(npix), 256, 0, st>>>(dy, g.y_cs, g.y_co, w, dx, g.x_cs, g.x_co, npix, \
                                                                                 g.H, g.W, accumulate);
  WDG_DIRECT_PAIRS(X)
#undef X
  CKT(cudaGetLastError());
  return 0;
}
static long long direct_slabs(const ConvGeo& g) {
  const long long npix = (long long)g.N * g.H * g.W;
  long long slabs = (npix + 1023) / 1024;
  if (slabs > 592) slabs = 592;
  return slabs < 1 ? 1 : slabs;
}
static int direct_bwd_weight(const ConvGeo& g, const float* x, const float* dy, float* dw, float* part, int accumulate, cudaStream_t st) {
  const long long npix = (long long)g.N * g.H * g.W, slabs = direct_slabs(g), per = (npix + slabs - 1) / slabs;
#define X(ci, co)                                                                                                            \
  if (g.Ci == ci && g.Co == co)                                                                                              \
    wdg_direct::conv3x3_bwd_weight_kernel<ci, co><<<(unsigned)slabs, 256, 0, st>>>(x, g.x_cs, g.x_co, dy, g.y_cs, g.y_co, part, \
                                                                                  npix, g.H, g.W, per);
  WDG_DIRECT_PAIRS(X)
#undef X
  CKT(cudaGetLastError());
  const long long n = 9ll * g.Ci * g.Co;
  reduce_splits_kernel<<<blocks_for(n), 256, 0, st>>>(part, dw, n, (int)slabs, accumulate);
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int wdg_conv2d_fwd_act(const float* x, const float* w, const float* bias, float* y, const int* geo, int accumulate,
                                  float alpha, void* stream) {
  if (wdg_sparse::applies(make_geo(geo))) {   // stride > kernel (critic shortcut conv): exact fp32 direct kernel, every mode
    const ConvGeo g = make_geo(geo);
    wdg_sparse::fwd_kernel<<<blocks_for((long long)g.N * g.Ho * g.Wo * g.Co), 256, 0, (cudaStream_t)stream>>>(g, x, w, bias, y, accumulate, alpha);
    CKT(cudaGetLastError());
    return 0;
  }
  if (direct_ok(make_geo(geo), DIRECT_FWD)) return direct_fwd(make_geo(geo), x, w, bias, y, accumulate, alpha, (cudaStream_t)stream);
  if (g_train_precision)
    return wdg_tc_conv2d_fwd(make_geo(geo), x, w, bias, y, accumulate, alpha, g_train_precision, (cudaStream_t)stream);
  FwdProblem p{make_geo(geo), x, w, bias, y, accumulate, alpha};
  const long long M = (long long)p.g.N * p.g.Ho * p.g.Wo;
  CKT(launch_gemm(p, M, p.g.Co, (cudaStream_t)stream));
  return 0;
}
extern "C" int wdg_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, const int* geo, int accumulate,
                              void* stream) {
  return wdg_conv2d_fwd_act(x, w, bias, y, geo, accumulate, 1.f, stream);
}

extern "C" int wdg_conv2d_bwd_data(const float* dy, const float* w, float* dx, const int* geo, int accumulate, void* stream) {
  const ConvGeo gg = make_geo(geo);
  if (wdg_sparse::applies(gg)) {
    wdg_sparse::bwd_data_kernel<<<blocks_for((long long)gg.N * gg.H * gg.W * gg.Ci), 256, 0, (cudaStream_t)stream>>>(gg, dy, w, dx, accumulate);
    CKT(cudaGetLastError());
    return 0;
  }
  if (direct_ok(gg, DIRECT_BWD_DATA)) return direct_bwd_data(gg, dy, w, dx, accumulate, (cudaStream_t)stream);
  if (g_train_precision) return wdg_tc_conv2d_bwd_data(gg, dy, w, dx, accumulate, g_train_precision, (cudaStream_t)stream);
  if (gg.stride > 1 && gg.kh >= gg.stride && gg.kw >= gg.stride) {
    const int s = gg.stride;
    for (int ry = 0; ry < s; ++ry)
      for (int rx = 0; rx < s; ++rx) {
        BwdDataClassProblem p{gg, dy, w, dx, accumulate};
        p.ry = ry; p.rx = rx;
        p.fy = ((ry - gg.pad_t) % s + s) % s; p.fx = ((rx - gg.pad_l) % s + s) % s;
        p.Hc = gg.H > p.fy ? (gg.H - p.fy + s - 1) / s : 0; p.Wc = gg.W > p.fx ? (gg.W - p.fx + s - 1) / s : 0;
        p.Jy = (gg.kh - ry + s - 1) / s; p.Jx = (gg.kw - rx + s - 1) / s;
        const long long M = (long long)gg.N * p.Hc * p.Wc;
        if (M == 0) continue;
        CKT(launch_gemm(p, M, gg.Ci, (cudaStream_t)stream));
      }
    return 0;
  }
  BwdDataProblem p{gg, dy, w, dx, accumulate};
  const long long M = (long long)p.g.N * p.g.H * p.g.W;
  CKT(launch_gemm(p, M, p.g.Ci, (cudaStream_t)stream));
  return 0;
}

static int sm_count_now() {      // split-K heuristic of the fp32 backward-weight GEMM: about four CTAs per SM
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}
// Split-K plan / scratch of the backward-weight GEMM in arithmetic `mode` (0 fp32, 1 tf32, 2 bf16).  Also called by
// wdg_critic.cu, which sizes its scratch for every mode without touching the process-wide setting.
int wdg_wgrad_scratch_for_mode(const int* geo, int mode, size_t* bytes, int* splits_out) {
  const ConvGeo g = make_geo(geo);
  const long long M = (long long)g.kh * g.kw * g.Ci, K = (long long)g.N * g.Ho * g.Wo;
  if (wdg_sparse::applies(g)) {   // no split-K partials
    if (splits_out) *splits_out = 1;
    if (bytes) *bytes = 0;
    return 0;
  }
  if (direct_ok(g, DIRECT_BWD_WEIGHT)) {
    const long long sl = direct_slabs(g);
    if (splits_out) *splits_out = (int)sl;
    if (bytes) *bytes = (size_t)(sl * M * g.Co * sizeof(float));
    return 0;
  }
  if (mode) {
    int sp; long long kps;
    wdg_tc_wgrad_plan(g, mode, &sp, &kps);
    if (splits_out) *splits_out = sp;
    if (bytes) *bytes = (size_t)((long long)sp * M * g.Co * sizeof(float));
    return 0;
  }
  const long long tiles = g.Co <= 16 ? ((M + 127) / 128) * ((g.Co + 15) / 16) : ((M + 63) / 64) * ((g.Co + 63) / 64);
  long long splits = (4ll * sm_count_now() + tiles - 1) / tiles;
  const long long max_splits = (K + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits_out) *splits_out = (int)splits;
  if (bytes) *bytes = (size_t)(splits * M * g.Co * sizeof(float));
  return 0;
}

extern "C" int wdg_conv2d_bwd_weight_scratch(const int* geo, size_t* bytes, int* splits_out) {
  return wdg_wgrad_scratch_for_mode(geo, g_train_precision, bytes, splits_out);
}

extern "C" int wdg_conv2d_bwd_weight(const float* x, const float* dy, float* dw, const int* geo, void* scratch,
                                     int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int splits = 1;
  if (wdg_sparse::applies(make_geo(geo))) {
    const ConvGeo g = make_geo(geo);
    wdg_sparse::bwd_weight_kernel<<<blocks_for((long long)g.kh * g.kw * g.Ci * g.Co), 256, 0, stream>>>(g, x, dy, dw, accumulate);
    CKT(cudaGetLastError());
    return 0;
  }
  wdg_conv2d_bwd_weight_scratch(geo, nullptr, &splits);
  if (direct_ok(make_geo(geo), DIRECT_BWD_WEIGHT)) return direct_bwd_weight(make_geo(geo), x, dy, dw, (float*)scratch, accumulate, stream);
  BwdWeightProblem p{make_geo(geo), x, dy, (float*)scratch, 0};
  const long long M = (long long)p.g.kh * p.g.kw * p.g.Ci, K = (long long)p.g.N * p.g.Ho * p.g.Wo;
  p.k_per_split = (int)((K + splits - 1) / splits);
  if (g_train_precision) {
    long long kps;
    wdg_tc_wgrad_plan(p.g, g_train_precision, &splits, &kps);
    if (int rc = wdg_tc_conv2d_bwd_weight(p.g, x, dy, (float*)scratch, splits, kps, g_train_precision, stream)) return rc;
  } else if (p.g.Co <= 16) {
    dim3 grid((unsigned)((M + 127) / 128), (p.g.Co + 15) / 16, splits);
    gemm_wgrad_kernel<128, 16><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid((unsigned)((M + 63) / 64), (p.g.Co + 63) / 64, splits);
    gemm_wgrad_kernel<64, 64><<<grid, 256, 0, stream>>>(p);
  }
  CKT(cudaGetLastError());
  const long long n = M * p.g.Co;
  reduce_splits_kernel<<<blocks_for(n), 256, 0, stream>>>((const float*)scratch, dw, n, splits, accumulate);
  CKT(cudaGetLastError());
  return 0;
}

// Column sums of a DENSE [R][32] fp32 matrix (the folded view of a narrow tensor): float4 loads, 8 lanes per row, 32
// rows per block pass, 4 passes in flight per thread.  Same partial layout as colsum_partial_kernel (part[slab][2][32]).
template <int MODE>   // 0: sum a   1: sum a*b   2: sum a*a   3: sum a and sum a*a
__global__ void __launch_bounds__(256) colsum32_dense_kernel(const float4* __restrict__ a, const float4* __restrict__ b, long long R,
                                                             float* __restrict__ part) {
  __shared__ float4 sm[2][32][9];
  const int q = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const long long rows_per = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * rows_per, r1 = min(R, r0 + rows_per);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (long long r = r0 + rl; r < r1; r += 128) {
    float4 av[4], bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long rr = r + 32 * j;
      av[j] = rr < r1 ? __ldg(a + rr * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 1) bv[j] = rr < r1 ? __ldg(b + rr * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE == 0 || MODE == 3) { s1.x += av[j].x; s1.y += av[j].y; s1.z += av[j].z; s1.w += av[j].w; }
      if (MODE == 1) { s1.x += av[j].x * bv[j].x; s1.y += av[j].y * bv[j].y; s1.z += av[j].z * bv[j].z; s1.w += av[j].w * bv[j].w; }
      if (MODE == 2) { s1.x += av[j].x * av[j].x; s1.y += av[j].y * av[j].y; s1.z += av[j].z * av[j].z; s1.w += av[j].w * av[j].w; }
      if (MODE == 3) { s2.x += av[j].x * av[j].x; s2.y += av[j].y * av[j].y; s2.z += av[j].z * av[j].z; s2.w += av[j].w * av[j].w; }
    }
  }
  sm[0][rl][q] = s1; sm[1][rl][q] = s2;
  __syncthreads();
  if (threadIdx.x < 16) {
    const int which = threadIdx.x >> 3, qq = threadIdx.x & 7;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < 32; ++i) { const float4 v = sm[which][i][qq]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    reinterpret_cast<float4*>(part + ((long long)blockIdx.x * 2 + which) * 32)[qq] = t;
  }
}

// dst[c] (+)= sum_s sum_j part[s * stride + j * C + c]: final step of a column sum computed on the [R/k][k*C] view of a
// dense narrow matrix (k = 32 / C pixels per 32-lane row, so no lane idles for the critic's 2..16-channel tensors)
__global__ void reduce_fold_kernel(const float* __restrict__ part, float* __restrict__ dst, int C, int k, int splits,
                                   long long stride, int accumulate) {
  const int c = blockIdx.x, lane = threadIdx.x;          // one warp per output channel, fixed reduction tree
  float s = 0.f;
  for (int i = lane; i < splits * k; i += 32) s += part[(long long)(i / k) * stride + (i % k) * C + c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) dst[c] = accumulate ? dst[c] + s : s;
}
// Two column sums in one pass (modes 3 / 4 of colsum_partial_kernel); scratch >= 2*CS_SLABS*max(C, 32) floats.
static int colsum_dual(int mode, const float* a, int a_cs, int a_co, const float* b, int b_cs, int b_co, const float* mean,
                       const float* invstd, long long R, int C, float* out1, float* out2, void* scratch, int accumulate,
                       cudaStream_t stream) {
  int fold = 1;
  const int C_out = C;
  if (mode != 4 && C < 32 && 32 % C == 0 && a_cs == C && a_co == 0 && (mode != 1 || (b_cs == C && b_co == 0)) &&
      R % (32 / C) == 0) {
    fold = 32 / C; R /= fold; C = 32; a_cs = 32; b_cs = 32;
  }
  long long slabs = (R + 255) / 256;
  if (slabs > CS_SLABS) slabs = CS_SLABS;
  if (slabs < 1) slabs = 1;
  dim3 grid((C + 31) / 32, (unsigned)slabs), block(32, 8);
  float* part = (float*)scratch;
  const bool dense32 = fold > 1 && ((uintptr_t)a & 15) == 0 && (mode != 1 || ((uintptr_t)b & 15) == 0);
  if (dense32) {
    const float4* a4 = (const float4*)a;
    const float4* b4 = (const float4*)b;
    if (mode == 0) colsum32_dense_kernel<0><<<(unsigned)slabs, 256, 0, stream>>>(a4, b4, R, part);
    else if (mode == 1) colsum32_dense_kernel<1><<<(unsigned)slabs, 256, 0, stream>>>(a4, b4, R, part);
    else if (mode == 2) colsum32_dense_kernel<2><<<(unsigned)slabs, 256, 0, stream>>>(a4, b4, R, part);
    else colsum32_dense_kernel<3><<<(unsigned)slabs, 256, 0, stream>>>(a4, b4, R, part);
  } else switch (mode) {
    case 0: colsum_partial_kernel<0><<<grid, block, 0, stream>>>(a, a_cs, a_co, b, b_cs, b_co, mean, invstd, R, C, part); break;
    case 1: colsum_partial_kernel<1><<<grid, block, 0, stream>>>(a, a_cs, a_co, b, b_cs, b_co, mean, invstd, R, C, part); break;
    case 2: colsum_partial_kernel<2><<<grid, block, 0, stream>>>(a, a_cs, a_co, b, b_cs, b_co, mean, invstd, R, C, part); break;
    case 3: colsum_partial_kernel<3><<<grid, block, 0, stream>>>(a, a_cs, a_co, b, b_cs, b_co, mean, invstd, R, C, part); break;
    default: colsum_partial_kernel<4><<<grid, block, 0, stream>>>(a, a_cs, a_co, b, b_cs, b_co, mean, invstd, R, C, part); break;
  }
  CKT(cudaGetLastError());
  if (fold > 1) {
    reduce_fold_kernel<<<C_out, 32, 0, stream>>>(part, out1, C_out, fold, (int)slabs, 2ll * C, accumulate);
    if (out2) reduce_fold_kernel<<<C_out, 32, 0, stream>>>(part + C, out2, C_out, fold, (int)slabs, 2ll * C, accumulate);
  } else {
    reduce_splits_kernel<<<blocks_for(C), 256, 0, stream>>>(part, out1, C, (int)slabs, accumulate, 2ll * C);
    if (out2) reduce_splits_kernel<<<blocks_for(C), 256, 0, stream>>>(part + C, out2, C, (int)slabs, accumulate, 2ll * C);
  }
  CKT(cudaGetLastError());
  return 0;
}
// column sums: mode 0 sum a, 1 sum a*b, 2 sum a*a over R rows; out[C]; scratch >= 512*C floats
extern "C" int wdg_colsum(int mode, const float* a, int a_cs, int a_co, const float* b, int b_cs, int b_co, long long R, int C,
                          float* out, void* scratch, int accumulate, void* stream_) {
  if (mode < 0 || mode > 2) return wdg_set_error("wdg_colsum: mode must be 0, 1 or 2");
  return colsum_dual(mode, a, a_cs, a_co, b, b_cs, b_co, nullptr, nullptr, R, C, out, nullptr, scratch, accumulate,
                     (cudaStream_t)stream_);
}

extern "C" int wdg_leaky_relu_fwd(float* x, long long n, float alpha, void* stream) {
  leaky_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(x, n, alpha);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_leaky_relu_bwd(float* dy, const float* y, long long n, float alpha, void* stream) {
  leaky_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(dy, y, n, alpha);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_axpby(float* out, int o_cs, int o_co, const float* x, int x_cs, int x_co, float a, const float* y, int y_cs,
                         int y_co, float b, long long rows, int C, int accumulate, void* stream) {
  axpby_kernel<<<blocks_for(rows * C), 256, 0, (cudaStream_t)stream>>>(out, o_cs, o_co, x, x_cs, x_co, a, y, y_cs, y_co, b, rows,
                                                                      C, accumulate);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_bias_act(float* x, int cs, int co, const float* bias, long long rows, int C, float alpha, void* stream) {
  bias_act_kernel<<<blocks_for(rows * C), 256, 0, (cudaStream_t)stream>>>(x, cs, co, bias, rows, C, alpha);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_transpose01(const float* in, float* out, int A, int B, long long inner, void* stream) {
  transpose01_kernel<<<blocks_for((long long)A * B * inner), 256, 0, (cudaStream_t)stream>>>(in, out, A, B, inner);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_col2im(const float* cols, float* out, int N, int H, int W, int kh, int kw, int Co, int pad, int o_cs, int o_co,
                          void* stream) {
  if (!cols || !out || N <= 0 || Co <= 0) return wdg_set_error("wdg_col2im: bad argument");
  if ((Co & 3) == 0 && ((uintptr_t)cols & 15) != 0) return wdg_set_error("wdg_col2im: cols must be 16-byte aligned");
  const long long total = (long long)N * H * W * ((Co + 3) / 4);
  col2im_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(cols, out, total, H, W, kh, kw, Co, pad, o_cs, o_co);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_lerp_batch(float* out, const float* real, const float* fake, const float* eps, long long per_sample,
                              long long n, void* stream) {
  lerp_batch_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(out, real, fake, eps, per_sample, n);
  CKT(cudaGetLastError());
  return 0;
}

// BatchNorm, training mode: batch statistics (biased variance), moving statistics updated in place.
// scratch >= (512*max(C,32) + 2*C) floats.  Saves mean / invstd ([C] each) for the backward.
extern "C" int wdg_bn_train_fwd(const float* x, float* y, const float* gamma, const float* beta, float* moving_mean,
                                float* moving_var, float* save_mean, float* save_invstd, long long rows, int C, float eps,
                                float momentum, void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  float* s1 = (float*)scratch + 2 * CS_SLABS * (size_t)(C < 32 ? 32 : C);
  float* s2 = s1 + C;
  if (colsum_dual(3, x, C, 0, nullptr, 0, 0, nullptr, nullptr, rows, C, s1, s2, scratch, 0, stream)) return 1;
  bn_finalize_kernel<<<blocks_for(C), 256, 0, stream>>>(s1, s2, rows, C, eps, momentum, save_mean, save_invstd, moving_mean, moving_var);
  CKT(cudaGetLastError());
  bn_apply_kernel<<<blocks_for(rows * C), 256, 0, stream>>>(x, y, save_mean, save_invstd, gamma, beta, rows, C);
  CKT(cudaGetLastError());
  return 0;
}
// Pieces of the above for data-parallel (synchronised) BatchNorm: the caller all-reduces the per-channel sums
// between the two calls.  s1 = sum x, s2 = sum x^2 over `rows_global` rows.
extern "C" int wdg_bn_finalize_apply(const float* x, float* y, const float* gamma, const float* beta, float* moving_mean,
                                     float* moving_var, const float* s1, const float* s2, float* save_mean, float* save_invstd,
                                     long long rows_local, long long rows_global, int C, float eps, float momentum,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  bn_finalize_kernel<<<blocks_for(C), 256, 0, stream>>>(s1, s2, rows_global, C, eps, momentum, save_mean, save_invstd, moving_mean, moving_var);
  CKT(cudaGetLastError());
  bn_apply_kernel<<<blocks_for(rows_local * C), 256, 0, stream>>>(x, y, save_mean, save_invstd, gamma, beta, rows_local, C);
  CKT(cudaGetLastError());
  return 0;
}
// dgamma_part = sum dy*xhat, dbeta_part = sum dy over the local rows (to be all-reduced), then dx with global sums.
extern "C" int wdg_bn_bwd_sums(const float* dy, const float* x, const float* save_mean, const float* save_invstd, float* dgamma,
                               float* dbeta, long long rows, int C, void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (colsum_dual(4, dy, C, 0, x, C, 0, save_mean, save_invstd, rows, C, dgamma, dbeta, scratch, 0, stream)) return 1;
  return 0;
}
extern "C" int wdg_bn_bwd_dx(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                             const float* dgamma, const float* dbeta, float* dx, long long rows_local, long long rows_global,
                             int C, float act_alpha, void* stream_) {
  bn_bwd_kernel<<<blocks_for(rows_local * C), 256, 0, (cudaStream_t)stream_>>>(dy, x, save_mean, save_invstd, gamma, dbeta, dgamma, dx,
                                                                              rows_local, C, rows_global, act_alpha);
  CKT(cudaGetLastError());
  return 0;
}

// Inference-mode BN with explicit statistics (moving mean / variance): invstd computed on the fly into scratch[C].
extern "C" int wdg_bn_infer(const float* x, float* y, const float* gamma, const float* beta, const float* mean,
                            const float* var, long long rows, int C, float eps, void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  float* invstd = (float*)scratch;
  rsqrt_eps_kernel<<<blocks_for(C), 256, 0, stream>>>(var, invstd, C, eps);
  CKT(cudaGetLastError());
  bn_apply_kernel<<<blocks_for(rows * C), 256, 0, stream>>>(x, y, mean, invstd, gamma, beta, rows, C);
  CKT(cudaGetLastError());
  return 0;
}
// dgamma, dbeta ([C]) and dx.  scratch >= 512*C floats.
extern "C" int wdg_bn_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean,
                                const float* save_invstd, float* dx, float* dgamma, float* dbeta, long long rows, int C,
                                float act_alpha, void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (colsum_dual(4, dy, C, 0, x, C, 0, save_mean, save_invstd, rows, C, dgamma, dbeta, scratch, 0, stream)) return 1;
  bn_bwd_kernel<<<blocks_for(rows * C), 256, 0, stream>>>(dy, x, save_mean, save_invstd, gamma, dbeta, dgamma, dx, rows, C, rows, act_alpha);
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int wdg_ln_fwd(const float* x, float* y, int y_cs, int y_co, const float* gamma, const float* beta, float* save_mean,
                          float* save_invstd, long long rows, int C, float eps, void* stream) {
  const int lpr = ln_lanes(C);
  const unsigned nb = blocks_for(rows * lpr);
  cudaStream_t st = (cudaStream_t)stream;
  if (lpr == 4) ln_fwd_kernel<4><<<nb, 256, 0, st>>>(x, y, gamma, beta, save_mean, save_invstd, rows, C, eps, y_cs, y_co);
  else if (lpr == 8) ln_fwd_kernel<8><<<nb, 256, 0, st>>>(x, y, gamma, beta, save_mean, save_invstd, rows, C, eps, y_cs, y_co);
  else if (lpr == 16) ln_fwd_kernel<16><<<nb, 256, 0, st>>>(x, y, gamma, beta, save_mean, save_invstd, rows, C, eps, y_cs, y_co);
  else ln_fwd_kernel<32><<<nb, 256, 0, st>>>(x, y, gamma, beta, save_mean, save_invstd, rows, C, eps, y_cs, y_co);
  CKT(cudaGetLastError());
  return 0;
}
// dx, dgamma, dbeta.  scratch >= rows*C + 512*C floats.
extern "C" int wdg_ln_bwd(const float* dy, int dy_cs, int dy_co, const float* x, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* dx, float* dgamma, float* dbeta, long long rows, int C,
                          float act_alpha, void* scratch, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  float* tmp = (float*)scratch;
  float* part = tmp + rows * C;
  const int lpr = ln_lanes(C);
  const unsigned nb = blocks_for(rows * lpr);
  if (lpr == 4) ln_bwd_kernel<4><<<nb, 256, 0, stream>>>(dy, dy_cs, dy_co, x, gamma, save_mean, save_invstd, dx, tmp, rows, C, act_alpha);
  else if (lpr == 8) ln_bwd_kernel<8><<<nb, 256, 0, stream>>>(dy, dy_cs, dy_co, x, gamma, save_mean, save_invstd, dx, tmp, rows, C, act_alpha);
  else if (lpr == 16) ln_bwd_kernel<16><<<nb, 256, 0, stream>>>(dy, dy_cs, dy_co, x, gamma, save_mean, save_invstd, dx, tmp, rows, C, act_alpha);
  else ln_bwd_kernel<32><<<nb, 256, 0, stream>>>(dy, dy_cs, dy_co, x, gamma, save_mean, save_invstd, dx, tmp, rows, C, act_alpha);
  CKT(cudaGetLastError());
  if (wdg_colsum(0, tmp, C, 0, nullptr, 0, 0, rows, C, dgamma, part, 0, stream_)) return 1;
  if (wdg_colsum(0, dy, dy_cs, dy_co, nullptr, 0, 0, rows, C, dbeta, part, 0, stream_)) return 1;
  return 0;
}

extern "C" int wdg_lstm_gates_fwd(float* z, const float* c_prev, float* c_out, float* h_out, long long rows, int F, void* stream) {
  lstm_gates_fwd_kernel<<<blocks_for(rows * F), 256, 0, (cudaStream_t)stream>>>(z, c_prev, c_out, h_out, rows, F);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_lstm_gates_bwd(float* gates, const float* c_prev, const float* c_cur, const float* dh, const float* dh_rec,
                                  float* dc, long long rows, int F, void* stream) {
  lstm_gates_bwd_kernel<<<blocks_for(rows * F), 256, 0, (cudaStream_t)stream>>>(gates, c_prev, c_cur, dh, dh_rec, dc, rows, F);
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int wdg_lstm_small_fwd(float* z, const float* h_prev, const float* R, const float* c_prev, float* c_out, float* h_out,
                                  int N, int H, int W, int F, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const unsigned nb = blocks_for((long long)N * H * W);
  if (F == 1) lstm_small_fwd_kernel<1><<<nb, 256, 0, stream>>>(z, h_prev, R, c_prev, c_out, h_out, N, H, W);
  else if (F == 2) lstm_small_fwd_kernel<2><<<nb, 256, 0, stream>>>(z, h_prev, R, c_prev, c_out, h_out, N, H, W);
  else if (F == 4) lstm_small_fwd_kernel<4><<<nb, 256, 0, stream>>>(z, h_prev, R, c_prev, c_out, h_out, N, H, W);
  else return wdg_set_error("wdg_lstm_small_fwd: F must be 1, 2 or 4");
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_lstm_small_bwd_data(const float* dz, const float* R, float* dh_rec, int N, int H, int W, int F, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const unsigned nb = blocks_for((long long)N * H * W);
  if (F == 1) lstm_small_bwd_data_kernel<1><<<nb, 256, 0, stream>>>(dz, R, dh_rec, N, H, W);
  else if (F == 2) lstm_small_bwd_data_kernel<2><<<nb, 256, 0, stream>>>(dz, R, dh_rec, N, H, W);
  else if (F == 4) lstm_small_bwd_data_kernel<4><<<nb, 256, 0, stream>>>(dz, R, dh_rec, N, H, W);
  else return wdg_set_error("wdg_lstm_small_bwd_data: F must be 1, 2 or 4");
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int wdg_upsample2x_fwd(const float* x, float* y, long long n_img, int h, int w, int C, void* stream) {
  if (C % 4 == 0 && n_img * 4 * h * w * (C / 4) < (1ll << 31) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0)
    upsample2x_fwd_v4_kernel<<<blocks_for(n_img * 4 * h * w * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)x, (float4*)y, (int)n_img, h, w, C / 4);
  else
    upsample2x_fwd_kernel<<<blocks_for(n_img * 4 * h * w * C), 256, 0, (cudaStream_t)stream>>>(x, y, n_img, h, w, C);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_upsample2x_bwd(const float* dy, float* dx, long long n_img, int h, int w, int C, void* stream) {
  if (C % 4 == 0 && n_img * 4 * h * w * (C / 4) < (1ll << 31) && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)dy & 15) == 0)
    upsample2x_bwd_v4_kernel<<<blocks_for(n_img * h * w * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, (float4*)dx, (int)n_img, h, w, C / 4);
  else
    upsample2x_bwd_kernel<<<blocks_for(n_img * h * w * C), 256, 0, (cudaStream_t)stream>>>(dy, dx, n_img, h, w, C);
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int wdg_dense_mean_fwd(const float* flat, const float* w, const float* bias, float* score, int B, int T, int D,
                                  void* stream) {
  dense_mean_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(flat, w, bias, score, T, D);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_dense_mean_bwd(const float* dscore, const float* flat, const float* w, float* dflat, float* dw, float* dbias,
                                  int B, int T, int D, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  dense_mean_bwd_kernel<<<blocks_for((long long)B * T * D), 256, 0, stream>>>(dscore, w, dflat, B, T, D);
  CKT(cudaGetLastError());
  if (dw) {
    dense_mean_wgrad_kernel<<<blocks_for(D), 256, 0, stream>>>(dscore, flat, dw, dbias, B, T, D);
    CKT(cudaGetLastError());
  }
  return 0;
}

// out[0] = scale * sum f;  mode 0 sum a, 1 sum a*b, 2 sum a*a.  scratch >= 1024 doubles.
extern "C" int wdg_reduce(int mode, const float* a, const float* b, long long n, double scale, float* out, void* scratch,
                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  double* part = (double*)scratch;
  if (mode == 0) reduce_partial_kernel<0><<<blocks, 256, 0, stream>>>(a, b, n, part);
  else if (mode == 1) reduce_partial_kernel<1><<<blocks, 256, 0, stream>>>(a, b, n, part);
  else reduce_partial_kernel<2><<<blocks, 256, 0, stream>>>(a, b, n, part);
  CKT(cudaGetLastError());
  reduce_final_kernel<<<1, 32, 0, stream>>>(part, blocks, out, scale);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_gp_norm(const float* g, float* out, int B, long long per_sample_px, int C, void* stream) {
  gp_norm_kernel<<<dim3(B, C), 256, 0, (cudaStream_t)stream>>>(g, out, per_sample_px, C);
  CKT(cudaGetLastError());
  return 0;
}

extern "C" int wdg_adam(float* w, float* m, float* v, const float* g, long long n, float lr_t, float b1, float b2, float eps,
                        void* stream) {
  adam_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(w, m, v, g, n, lr_t, b1, b2, eps);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_adam_lr(float* lr_t_dev, int* step_dev, float lr, float b1, float b2, void* stream) {
  if (!lr_t_dev || !step_dev) return wdg_set_error("null argument");
  adam_lr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(lr_t_dev, step_dev, lr, b1, b2);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_adam_dev(float* w, float* m, float* v, const float* g, long long n, const float* lr_t_dev, float b1, float b2,
                            float eps, void* stream) {
  adam_dev_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(w, m, v, g, n, lr_t_dev, b1, b2, eps);
  CKT(cudaGetLastError());
  return 0;
}
extern "C" int wdg_sn_update(float* w, float* u, int R, int C, void* scratch /* >= R + 64*C + 4 floats */, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  float* vbuf = (float*)scratch;
  float* part = vbuf + R;
  float* scal = part + (long long)SN_SLABS * C;
  int slabs = (R + 31) / 32;
  if (slabs > SN_SLABS) slabs = SN_SLABS;
  sn_rowdot_kernel<<<(R + 7) / 8, 256, 0, stream>>>(w, u, vbuf, R, C);
  sn_coldot_kernel<<<dim3((C + 31) / 32, slabs), dim3(32, 8), 0, stream>>>(w, vbuf, part, R, C);
  sn_finish_kernel<<<1, 256, 0, stream>>>(vbuf, part, u, scal, R, C, slabs);
  sn_scale_kernel<<<blocks_for((long long)R * C), 256, 0, stream>>>(w, scal, (long long)R * C);
  CKT(cudaGetLastError());
  return 0;
}
