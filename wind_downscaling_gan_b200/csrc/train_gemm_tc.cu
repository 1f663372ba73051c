// Training convolutions on the 5th-generation tensor cores (critic and training-mode
// generator forward / backward-data / backward-weight).
//
// One kernel template, `tc_gemm_kernel<P, BN, OP>`: C[128 x BN] tiles of an implicit GEMM.  The A operand (and, for
// backward-weight, B) is GATHERED by 8 loader warps straight from the fp32 channels-last activations (no im2col
// buffer), rounded to bf16 or tf32 in registers and written into shared memory in the canonical K-major
// SWIZZLE_128B UMMA layout (one 16-byte chunk = 8 bf16 / 4 tf32 per st.shared.v4).  For forward / backward-data the B
// operand is the weight matrix: a small pre-pass rounds and packs it K-major ([N][K]) into a library arena and a
// producer warp streams its tiles with TMA (cp.async.bulk.tensor, SWIZZLE_128B), so the loader warps make ONE global
// round trip per K block.  A ninth warp issues `tcgen05.mma` (kind::f16 or kind::tf32) with the fp32 accumulator in
// TMEM; after the K loop the loader warps read the accumulator back (`tcgen05.ld`) and run the problem's epilogue
// (bias / accumulate / strided channel views / split-K partials).  Two CTAs are co-resident per SM so one tile's
// epilogue overlaps the other's gathers.
//
// The three problems are the same ones train_ops.cu defines for the CUDA-core path:
//   forward        C[m = (n,oy,ox)][co]      = sum_{k=(ky,kx,ci)} x[n, oy*s-p+ky, ox*s-p+kx, ci] * w[k][co]
//   backward-data  C[m = (n,a,b) in class][ci] = sum_{k=(jy,jx,co)} dy[n, ay-jy, ax-jx, co] * w[ry+s*jy][rx+s*jx][ci][co]
//                  (one residue class of the stride at a time: no structural zeros; stride 1 is the single class)
//   backward-weight C[m = (ky,kx,ci)][co]    = sum_{k=(n,oy,ox)} x[n, oy*s-p+ky, ox*s-p+kx, ci] * dy[k][co]   (split-K)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/wdg.h"
#include "ptx.cuh"
#include "train_geo.cuh"

namespace {
using namespace wdg;

constexpr int OP_TF32 = 1, OP_BF16 = 2;
template <int OP> struct OpT { static constexpr int G = (OP == OP_BF16) ? 8 : 4; };   // elements per 16-byte chunk
constexpr int LOADER_THREADS = 256;
constexpr int TC_THREADS = LOADER_THREADS + 64;   // + MMA warp + TMA producer warp
constexpr int TILE_M = 128;
constexpr uint32_t A_STAGE_BYTES = TILE_M * 128;

// 16 bytes of one K-major SWIZZLE_128B row: tile rows are 128 bytes, chunk c of row r lives at chunk c ^ (r & 7).
template <int OP>
__device__ __forceinline__ void store_chunk(uint32_t tile, int row, int chunk, const float (&v)[OpT<OP>::G]) {
  const uint32_t addr = tile + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
  uint32_t w0, w1, w2, w3;
  if constexpr (OP == OP_BF16) {
    w0 = pack_bf16x2(v[0], v[1]); w1 = pack_bf16x2(v[2], v[3]); w2 = pack_bf16x2(v[4], v[5]); w3 = pack_bf16x2(v[6], v[7]);
  } else {
    w0 = to_tf32(v[0]); w1 = to_tf32(v[1]); w2 = to_tf32(v[2]); w3 = to_tf32(v[3]);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
}

// ------------------------------------------------------------------------------------------------ K walkers
// k = (ty * ntx + tx) * C + c, advanced incrementally (no division in the K loop).
struct TapK {
  int c, tx, ty;
  __device__ void init(long long k, int C, int ntx) {
    c = (int)(k % C);
    const int t = (int)(k / C);
    tx = t % ntx; ty = t / ntx;
  }
  __device__ __forceinline__ void advance(int d, int C, int ntx) {
    c += d;
    while (c >= C) { c -= C; if (++tx == ntx) { tx = 0; ++ty; } }
  }
};

// ------------------------------------------------------------------------------------------------ A: tap gather
// K-fastest gather of 4 rows x one 16-byte chunk per thread (rows (tid >> 3) + 32 i, chunk tid & 7): the
// forward A operand (sgn = +1) and the backward-data A operand (sgn = -1).
struct GatherSrc {
  int cs, C, ntx, nty, sgn, Hs, Ws, vec;
};
template <int OP>
struct TapGatherA {
  static constexpr int G = OpT<OP>::G;
  GatherSrc s;
  const float* base[4];   // pixel (y0, x0) of the row's image, channel offset applied (may point outside; never read then)
  int y0[4], x0[4];
  __device__ __forceinline__ void set_row(int i, bool valid, const float* img, int y, int x) {
    y0[i] = valid ? y : -(1 << 28);
    x0[i] = x;
    base[i] = valid ? img + ((long long)y * s.Ws + x) * s.cs : img;
  }
  __device__ __forceinline__ void fetch(const TapK& k, float (&v)[4][G]) const {
    if (s.vec && k.c + G <= s.C) {
      const int dy = s.sgn * k.ty, dx = s.sgn * k.tx;
      const long long off = ((long long)dy * s.Ws + dx) * s.cs + k.c;
      const bool kin = k.ty < s.nty;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int y = y0[i] + dy, x = x0[i] + dx;
        if (kin && (unsigned)y < (unsigned)s.Hs && (unsigned)x < (unsigned)s.Ws) {
          const float4* q = reinterpret_cast<const float4*>(base[i] + off);
#pragma unroll
          for (int h = 0; h < G / 4; ++h) {
            const float4 t = __ldg(q + h);
            v[i][4 * h] = t.x; v[i][4 * h + 1] = t.y; v[i][4 * h + 2] = t.z; v[i][4 * h + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < G; ++e) v[i][e] = 0.f;
        }
      }
    } else {
      TapK w = k;
#pragma unroll
      for (int e = 0; e < G; ++e) {
        const int dy = s.sgn * w.ty, dx = s.sgn * w.tx;
        const long long off = ((long long)dy * s.Ws + dx) * s.cs + w.c;
        const bool kin = w.ty < s.nty;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int y = y0[i] + dy, x = x0[i] + dx;
          v[i][e] = (kin && (unsigned)y < (unsigned)s.Hs && (unsigned)x < (unsigned)s.Ws) ? __ldg(base[i] + off) : 0.f;
        }
        w.advance(1, s.C, s.ntx);
      }
    }
  }
  __device__ __forceinline__ void store(uint32_t tile, int tid, const float (&v)[4][G]) const {
    const int j = tid & 7, r0 = tid >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) store_chunk<OP>(tile, r0 + 32 * i, j, v[i]);
  }
  __device__ __forceinline__ void load(uint32_t tile, int tid, const TapK& k) const {
    float v[4][G];
    fetch(k, v);
    store(tile, tid, v);
  }
};

// ------------------------------------------------------------------------------------------------ B: linear rows
// Row-fastest loader of B(k, n) = base[k * ld + n] (forward: the HWIO weights; backward-weight: dy): thread owns
// row tid % BN and CPT consecutive chunks; lanes walk n, so every load instruction is coalesced.
template <int OP, int BN>
struct LinearRowFastB {
  static constexpr int G = OpT<OP>::G;
  static constexpr int CPT = BN >= 32 ? BN / 32 : 1;
  const float* q;          // element (k of this thread's first chunk, n)
  long long ld;
  long long k, k_end;      // k of this thread's first chunk in the current K block
  bool active, nv;
  int row, chunk0;
  __device__ void init(const float* base, long long ld_, int n0, int n_valid, long long k_begin, long long k_end_, int tid) {
    row = tid % BN;
    const int group = tid / BN;
    chunk0 = group * CPT;
    active = chunk0 < 8;
    nv = n0 + row < n_valid;
    ld = ld_; k_end = k_end_;
    k = k_begin + chunk0 * G;
    q = base + k * ld + n0 + row;
  }
  __device__ __forceinline__ void load(uint32_t tile) {
    if (active) {
#pragma unroll 1
      for (int c0 = 0; c0 < CPT; c0 += 4) {
        constexpr int NB = CPT < 4 ? CPT : 4;
        float v[NB][G];
#pragma unroll
        for (int c = 0; c < NB; ++c)
#pragma unroll
          for (int e = 0; e < G; ++e) {
            const int kk = (c0 + c) * G + e;
            v[c][e] = (nv && k + kk < k_end) ? __ldg(q + kk * ld) : 0.f;
          }
#pragma unroll
        for (int c = 0; c < NB; ++c) store_chunk<OP>(tile, row, chunk0 + c0 + c, v[c]);
      }
    }
    k += 8 * G;
    q += 8 * G * ld;
  }
  __device__ __forceinline__ void load_skip() {
    k += 8 * G;
    q += 8 * G * ld;
  }
  // Split form: fetch the first (up to 4-chunk) batch into registers so that the caller can issue it together with
  // its own global loads (ONE round trip per K block), store it later; the remaining batches (BN = 256) follow.
  static constexpr int NB0 = CPT < 4 ? CPT : 4;
  __device__ __forceinline__ void fetch0(float (&v)[NB0][G]) const {
#pragma unroll
    for (int c = 0; c < NB0; ++c)
#pragma unroll
      for (int e = 0; e < G; ++e) {
        const int kk = c * G + e;
        v[c][e] = (active && nv && k + kk < k_end) ? __ldg(q + kk * ld) : 0.f;
      }
  }
  __device__ __forceinline__ void store0_and_rest(uint32_t tile, const float (&v0)[NB0][G]) {
    if (active) {
#pragma unroll
      for (int c = 0; c < NB0; ++c) store_chunk<OP>(tile, row, chunk0 + c, v0[c]);
#pragma unroll 1
      for (int c0 = 4; c0 < CPT; c0 += 4) {
        float v[NB0][G];
#pragma unroll
        for (int c = 0; c < NB0; ++c)
#pragma unroll
          for (int e = 0; e < G; ++e) {
            const int kk = (c0 + c) * G + e;
            v[c][e] = (nv && k + kk < k_end) ? __ldg(q + kk * ld) : 0.f;
          }
#pragma unroll
        for (int c = 0; c < NB0; ++c) store_chunk<OP>(tile, row, chunk0 + c0 + c, v[c]);
      }
    }
    k += 8 * G;
    q += 8 * G * ld;
  }
};

// ------------------------------------------------------------------------------------------------ problems
struct TcFwd {
  ConvGeo g; const float* x; const float* w; const float* bias; float* y; int accumulate, vec_in, vec_out; float alpha;
  __device__ long long M() const { return (long long)g.N * g.Ho * g.Wo; }
  __device__ int Nn() const { return g.Co; }
  __device__ long long K() const { return (long long)g.kh * g.kw * g.Ci; }

  static constexpr bool TMA_B = true;
  static constexpr bool PAIR = true;    // loaders offer fetch / store (two K blocks in flight)
  __device__ float B(long long k, int n) const { return w[k * g.Co + n]; }
  template <int OP, int BN>
  struct Loaders {
    TapGatherA<OP> a; TapK k;
    __device__ Loaders(const TcFwd& p, long long m0, int n0, long long k_begin, long long k_end, int tid) {
      const ConvGeo& g = p.g;
      a.s = GatherSrc{g.x_cs, g.Ci, g.kw, g.kh, +1, g.H, g.W, p.vec_in};
      const long long M = p.M();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long m = m0 + (tid >> 3) + 32 * i;
        const bool valid = m < M;
        const long long mm = valid ? m : 0;
        const int ox = (int)(mm % g.Wo), oy = (int)((mm / g.Wo) % g.Ho), n = (int)(mm / ((long long)g.Wo * g.Ho));
        a.set_row(i, valid, p.x + g.x_co + (long long)n * g.H * g.W * g.x_cs, oy * g.stride - g.pad_t, ox * g.stride - g.pad_l);
      }
      k.init(k_begin + (tid & 7) * OpT<OP>::G, g.Ci, g.kw);
    }
    __device__ __forceinline__ void load(uint32_t a_tile, uint32_t, int tid) {
      a.load(a_tile, tid, k);
      k.advance(8 * OpT<OP>::G, a.s.C, a.s.ntx);
    }
    // split form: the kernel keeps the gathers of TWO K blocks in flight before the first shared-memory store
    __device__ __forceinline__ void fetch(float (&v)[4][OpT<OP>::G]) {
      a.fetch(k, v);
      k.advance(8 * OpT<OP>::G, a.s.C, a.s.ntx);
    }
    __device__ __forceinline__ void store(uint32_t a_tile, int tid, const float (&v)[4][OpT<OP>::G]) const { a.store(a_tile, tid, v); }
  };
  __device__ __forceinline__ void store16(long long m, int n, const uint32_t (&r)[16], int) const {
    float* q = y + m * g.y_cs + g.y_co + n;
    if (vec_out && n + 16 <= g.Co) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                               __uint_as_float(r[4 * i + 3]));
        if (bias) { v.x += __ldg(bias + n + 4 * i); v.y += __ldg(bias + n + 4 * i + 1); v.z += __ldg(bias + n + 4 * i + 2); v.w += __ldg(bias + n + 4 * i + 3); }
        float4* d = reinterpret_cast<float4*>(q) + i;
        if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        v.x = v.x >= 0.f ? v.x : alpha * v.x; v.y = v.y >= 0.f ? v.y : alpha * v.y;
        v.z = v.z >= 0.f ? v.z : alpha * v.z; v.w = v.w >= 0.f ? v.w : alpha * v.w;
        *d = v;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (n + e < g.Co) {
          float v = __uint_as_float(r[e]);
          if (bias) v += __ldg(bias + n + e);
          if (accumulate) v += q[e];
          q[e] = v >= 0.f ? v : alpha * v;
        }
    }
  }
};

struct TcBwdData {
  ConvGeo g; BwdClass c; const float* dy; const float* w; float* dx; int accumulate, vec_in, vec_w, vec_out;
  __device__ long long M() const { return (long long)g.N * c.Hc * c.Wc; }
  __device__ int Nn() const { return g.Ci; }
  __device__ long long K() const { return (long long)c.Jy * c.Jx * g.Co; }

  static constexpr bool TMA_B = true;
  static constexpr bool PAIR = true;    // loaders offer fetch / store (two K blocks in flight)
  // B(k = (jy, jx, co), n = ci) = w[((ry + s jy) * kw + rx + s jx) * Ci + ci][co]
  __device__ float B(long long k, int n) const {
    const int co = (int)(k % g.Co), t = (int)(k / g.Co), jx = t % c.Jx, jy = t / c.Jx;
    return w[(((long long)(c.ry + g.stride * jy) * g.kw + c.rx + g.stride * jx) * g.Ci + n) * g.Co + co];
  }
  template <int OP, int BN>
  struct Loaders {
    static constexpr int G = OpT<OP>::G;
    TapGatherA<OP> a; TapK k;
    __device__ Loaders(const TcBwdData& p, long long m0, int, long long k_begin, long long, int tid) {
      const ConvGeo& g = p.g;
      a.s = GatherSrc{g.y_cs, g.Co, p.c.Jx, p.c.Jy, -1, g.Ho, g.Wo, p.vec_in};
      const long long M = p.M();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long m = m0 + (tid >> 3) + 32 * i;
        const bool valid = m < M;
        const long long mm = valid ? m : 0;
        const int b = (int)(mm % p.c.Wc), aa = (int)((mm / p.c.Wc) % p.c.Hc), n = (int)(mm / ((long long)p.c.Wc * p.c.Hc));
        const int ay = (p.c.fy + g.stride * aa + g.pad_t - p.c.ry) / g.stride;   // = oy + jy
        const int ax = (p.c.fx + g.stride * b + g.pad_l - p.c.rx) / g.stride;
        a.set_row(i, valid, p.dy + g.y_co + (long long)n * g.Ho * g.Wo * g.y_cs, ay, ax);
      }
      k.init(k_begin + (tid & 7) * G, g.Co, p.c.Jx);
    }
    __device__ __forceinline__ void load(uint32_t a_tile, uint32_t, int tid) {
      a.load(a_tile, tid, k);
      k.advance(8 * G, a.s.C, a.s.ntx);
    }
    __device__ __forceinline__ void fetch(float (&v)[4][G]) {
      a.fetch(k, v);
      k.advance(8 * G, a.s.C, a.s.ntx);
    }
    __device__ __forceinline__ void store(uint32_t a_tile, int tid, const float (&v)[4][G]) const { a.store(a_tile, tid, v); }
  };
  __device__ __forceinline__ void store16(long long m, int n, const uint32_t (&r)[16], int) const {
    const int b = (int)(m % c.Wc), a = (int)((m / c.Wc) % c.Hc), img = (int)(m / ((long long)c.Wc * c.Hc));
    const int iy = c.fy + g.stride * a, ix = c.fx + g.stride * b;
    float* q = dx + (((long long)img * g.H + iy) * g.W + ix) * g.x_cs + g.x_co + n;
    if (vec_out && n + 16 <= g.Ci) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                               __uint_as_float(r[4 * i + 3]));
        float4* d = reinterpret_cast<float4*>(q) + i;
        if (accumulate) { const float4 o = *d; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *d = v;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (n + e < g.Ci) q[e] = accumulate ? q[e] + __uint_as_float(r[e]) : __uint_as_float(r[e]);
    }
  }
};

struct TcWgrad {
  ConvGeo g; const float* x; const float* dy; float* part; long long k_per_split; int vec_a, vec_b;
  __device__ long long M() const { return (long long)g.kh * g.kw * g.Ci; }
  __device__ int Nn() const { return g.Co; }
  __device__ long long K() const { return (long long)g.N * g.Ho * g.Wo; }
  static constexpr bool TMA_B = false;
  static constexpr bool PAIR = false;

  // Both operands are contiguous along their ROW index in memory (A: ci, B: co) while a 16-byte smem chunk holds G
  // consecutive k (pixels) of ONE row, so the loaders transpose in registers.  Vector form (channel counts and views
  // multiples of 4): a thread loads float4 = 4 consecutive rows for each of the G pixels of its chunk and writes 4
  // chunks (G LDG.128 per 4 chunks); scalar form: 4 G LDG.32 per 4 chunks, lanes walking the rows.  Which one runs
  // is decided per layer from measurements (wdg_tc_conv2d_bwd_weight below).
  template <int OP, int BN>
  struct Loaders {
    static constexpr int G = OpT<OP>::G;
    static constexpr int KB = 8 * G;
    static constexpr int RUN = 4 * G;      // scalar form: consecutive k per thread and K block (4 chunks); two thread halves interleave
    LinearRowFastB<OP, BN> b;
    // A(m = (ky, kx, ci), k = (n, oy, ox)) = x[n, oy s - p + ky, ox s - p + kx, ci]
    const float* col;   // x + x_co + ci of this thread's (first) row
    const float* dyb;   // dy + y_co + n0
    int ky, kx, row, chunk0;
    bool mv, vecA, vecB;
    int ox, oy, n;
    long long k, k_end, kb0;
    int H, W, Ho, Wo, s, pad_t, pad_l, cs, y_cs, n_left;
    __device__ Loaders(const TcWgrad& p, long long m0, int n0, long long k_begin, long long k_end_, int tid) {
      const ConvGeo& g = p.g;
      vecA = p.vec_a; vecB = p.vec_b;
      if (vecA) { row = 4 * (tid >> 3); chunk0 = tid & 7; }
      else { row = tid & 127; chunk0 = (tid >> 7) * 4; }
      const long long m = m0 + row;
      mv = m < p.M();
      const int mm = mv ? (int)m : 0;
      const int ci = mm % g.Ci, tap = mm / g.Ci;
      kx = tap % g.kw; ky = tap / g.kw;
      col = p.x + g.x_co + ci;
      H = g.H; W = g.W; Ho = g.Ho; Wo = g.Wo; s = g.stride; pad_t = g.pad_t; pad_l = g.pad_l; cs = g.x_cs;
      k = k_begin + chunk0 * G; k_end = k_end_; kb0 = k_begin;
      ox = (int)(k % Wo); oy = (int)((k / Wo) % Ho); n = (int)(k / ((long long)Wo * Ho));
      b.init(p.dy + g.y_co, g.y_cs, n0, g.Co, k_begin, k_end_, tid);
      dyb = p.dy + g.y_co + n0; y_cs = g.y_cs; n_left = g.Co - n0;
    }
    __device__ __forceinline__ void advance(int d) {
      k += d; ox += d;
      while (ox >= Wo) { ox -= Wo; if (++oy == Ho) { oy = 0; ++n; } }
    }
    __device__ __forceinline__ void step() {
      ++k;
      if (++ox == Wo) { ox = 0; if (++oy == Ho) { oy = 0; ++n; } }
    }
    __device__ __forceinline__ void load(uint32_t a_tile, uint32_t b_tile, int tid) {
      // incremental addressing along an image row: pointer += s * cs per pixel, full recomputation only at a row wrap
      int iy = oy * s - pad_t + ky, ix = ox * s - pad_l + kx;
      bool rowok = mv && (unsigned)iy < (unsigned)H;
      const float* p = col + (((long long)n * H + iy) * W + ix) * cs;
      const int dp = s * cs;
      auto next_px = [&]() {
        ++k; ix += s; p += dp;
        if (++ox == Wo) {
          ox = 0;
          if (++oy == Ho) { oy = 0; ++n; }
          iy = oy * s - pad_t + ky; ix = kx - pad_l;
          rowok = mv && (unsigned)iy < (unsigned)H;
          p = col + (((long long)n * H + iy) * W + ix) * cs;
        }
      };
      if (vecA) {
        float4 t[G];
#pragma unroll
        for (int e = 0; e < G; ++e) {
          const bool ok = rowok && k < k_end && (unsigned)ix < (unsigned)W;
          t[e] = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
          next_px();
        }
        float v[4][G];
#pragma unroll
        for (int e = 0; e < G; ++e) { v[0][e] = t[e].x; v[1][e] = t[e].y; v[2][e] = t[e].z; v[3][e] = t[e].w; }
#pragma unroll
        for (int i = 0; i < 4; ++i) store_chunk<OP>(a_tile, row + i, chunk0, v[i]);
        advance(KB - G);
      } else {
        float v[4][G];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < G; ++e) {
            const bool ok = rowok && k < k_end && (unsigned)ix < (unsigned)W;
            v[c][e] = ok ? __ldg(p) : 0.f;
            next_px();
          }
        if (!vecB) {       // issue the B gathers before the first shared-memory store: one global round trip per K block
          float vb[LinearRowFastB<OP, BN>::NB0][G];
          b.fetch0(vb);
#pragma unroll
          for (int c = 0; c < 4; ++c) store_chunk<OP>(a_tile, row, chunk0 + c, v[c]);
          advance(RUN);
          b.store0_and_rest(b_tile, vb);
          kb0 += KB;
          return;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) store_chunk<OP>(a_tile, row, chunk0 + c, v[c]);
        advance(RUN);
      }
      if (vecB) {
        // B(k, n) = dy[k][n]: row groups of 4 consecutive n, chunk = tid & 7
        const int j = tid & 7;
        const long long kk = kb0 + j * G;
#pragma unroll
        for (int i = 0; i < (BN + 127) / 128; ++i) {
          const int r4 = 4 * ((tid >> 3) + 32 * i);
          if (r4 < BN) {
            const bool nv = r4 < n_left;
            float4 t[G];
#pragma unroll
            for (int e = 0; e < G; ++e)
              t[e] = (nv && kk + e < k_end) ? __ldg(reinterpret_cast<const float4*>(dyb + (kk + e) * y_cs + r4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float v[4][G];
#pragma unroll
            for (int e = 0; e < G; ++e) { v[0][e] = t[e].x; v[1][e] = t[e].y; v[2][e] = t[e].z; v[3][e] = t[e].w; }
#pragma unroll
            for (int q = 0; q < 4; ++q) store_chunk<OP>(b_tile, r4 + q, j, v[q]);
          }
        }
        b.load_skip();
      } else {
        b.load(b_tile);
      }
      kb0 += KB;
    }
  };
  __device__ __forceinline__ void store16(long long m, int n, const uint32_t (&r)[16], int split) const {
    float* q = part + ((long long)split * M() + m) * g.Co + n;
    if ((g.Co & 3) == 0 && n + 16 <= g.Co) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(q)[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                      __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (n + e < g.Co) q[e] = __uint_as_float(r[e]);
    }
  }
};

// ------------------------------------------------------------------------------------------------ kernel
template <int BN> struct TcCfg {
  static constexpr int STAGES = BN >= 256 ? 2 : (BN >= 128 ? 3 : 4);
  static constexpr uint32_t B_STAGE_BYTES = BN * 128;
  static constexpr uint32_t SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// Rounds B(k, n) to the operand type and packs it K-major: out[n][K_pad] (K_pad = K rounded up to whole K blocks, zero
// filled).  N_FAST: consecutive threads walk n (forward weights: n contiguous in memory), else k.
template <class P, int OP, bool N_FAST>
__global__ void pack_b_kernel(const P p, void* __restrict__ out, int N, long long K, long long K_pad) {
  constexpr int G = OpT<OP>::G;
  const long long chunks = K_pad / G;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= chunks * N) return;
  const int n = N_FAST ? (int)(i % N) : (int)(i / chunks);
  const long long kc = N_FAST ? i / N : i % chunks;
  float v[G];
#pragma unroll
  for (int e = 0; e < G; ++e) {
    const long long k = kc * G + e;
    v[e] = k < K ? p.B(k, n) : 0.f;
  }
  uint4 o;
  if constexpr (OP == OP_BF16) {
    o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  } else {
    o = make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
  }
  reinterpret_cast<uint4*>(out)[(long long)n * chunks + kc] = o;
}

template <class P, int BN, int OP>
__global__ void __launch_bounds__(TC_THREADS, 2) tc_gemm_kernel(const P p, int n_tiles_n, long long k_per_split,
                                                                const __grid_constant__ CUtensorMap tmB) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int KB = 8 * OpT<OP>::G;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = sm;
  uint8_t* smB = sm + STAGES * A_STAGE_BYTES;
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], LOADER_THREADS + (P::TMA_B ? 1 : 0)); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
    if (P::TMA_B) prefetch_tmap(&tmB);
  }
  if (warp == 8) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const int n_tile = blockIdx.x % n_tiles_n;
  const long long m0 = (long long)(blockIdx.x / n_tiles_n) * TILE_M;
  const int n0 = n_tile * BN;
  const int split = blockIdx.y;
  const long long K = p.K();
  const long long k_begin = (long long)split * k_per_split;
  const long long k_end = (k_begin + k_per_split < K) ? k_begin + k_per_split : K;
  const int num_kb = k_end > k_begin ? (int)((k_end - k_begin + KB - 1) / KB) : 0;

  if (warp < 8) {
    // ===================================================== gather warps, then epilogue
    {
      typename P::template Loaders<OP, BN> ld(p, m0, n0, k_begin, k_end, tid);
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (P::PAIR && OP == OP_TF32 && STAGES >= 3) {
        // tf32 chunks are 16 source bytes: one K block keeps only 4 x 16 B per thread in flight and the loop is bound by
        // the L2 round trip, so two K blocks are gathered back to back (8 x 16 B in flight) before the first store
        for (int kb = 0; kb < num_kb; kb += 2) {
          const bool two = kb + 1 < num_kb;
          int stage2 = stage + 1;
          uint32_t phase2 = phase;
          if (stage2 == STAGES) { stage2 = 0; phase2 ^= 1; }
          float v0[4][OpT<OP>::G], v1[4][OpT<OP>::G];
          ld.fetch(v0);
          if (two) ld.fetch(v1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          ld.store(smem_u32(smA + stage * A_STAGE_BYTES), tid, v0);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&full_bar[stage]);
          if (two) {
            mbar_wait(&empty_bar[stage2], phase2 ^ 1);
            ld.store(smem_u32(smA + stage2 * A_STAGE_BYTES), tid, v1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full_bar[stage2]);
            stage = stage2; phase = phase2;
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      } else {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          ld.load(smem_u32(smA + stage * A_STAGE_BYTES), smem_u32(smB + stage * Cfg::B_STAGE_BYTES), tid);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
          mbar_arrive(&full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (num_kb > 0) {
      mbar_wait(&acc_bar, 0);
      tc_fence_after();
    }
    const int quad = warp & 3, half = warp >> 2;
    const long long m = m0 + quad * 32 + lane;
    const bool mvalid = m < p.M();
    const int N = p.Nn();
    constexpr int COLS_PER_HALF = BN >= 32 ? BN / 2 : BN;
    if (BN >= 32 || half == 0) {
#pragma unroll 1
      for (int c0 = half * COLS_PER_HALF; c0 < (half + 1) * COLS_PER_HALF; c0 += 16) {
        uint32_t r[16];
        if (num_kb > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = 0u;
        }
        if (mvalid && n0 + c0 < N) p.store16(m, n0 + c0, r, split);
      }
    }
  } else if (warp == 9) {
    // ===================================================== TMA producer of the packed B tiles
    if (P::TMA_B) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_STAGE_BYTES);
          tma_load_2d(smB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], (int)(k_begin + (long long)kb * KB), n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = (OP == OP_BF16) ? umma_idesc_bf16(TILE_M, BN) : umma_idesc_tf32(TILE_M, BN);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_kmajor(smem_u32(smA + stage * A_STAGE_BYTES), 128u);
        const uint64_t db = umma_desc_kmajor(smem_u32(smB + stage * Cfg::B_STAGE_BYTES), 128u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // 4 x 32 bytes of K per 128-byte row
          if constexpr (OP == OP_BF16) umma_bf16(tmem_base, da + 2 * q, db + 2 * q, idesc, (kb | q) ? 1u : 0u);
          else umma_tf32(tmem_base, da + 2 * q, db + 2 * q, idesc, (kb | q) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(&acc_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---- library arena for the packed B operands (weights): a 64 MB ring, allocated on first use.  Calls are ordered
// by the single stream the training ops run on; a packed copy is only read by the GEMM launched right after it.
constexpr size_t ARENA_BYTES = 64ull << 20;
constexpr int MAX_DEVICES = 64;
char* g_arena[MAX_DEVICES] = {};       // one ring per device (allocated on that device at first use)
size_t g_arena_off[MAX_DEVICES] = {};
int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < MAX_DEVICES ? d : 0;
}
void* arena_get(size_t bytes) {
  const int d = current_device();
  if (!g_arena[d] && cudaMalloc(&g_arena[d], ARENA_BYTES) != cudaSuccess) return nullptr;
  bytes = (bytes + 1023) & ~(size_t)1023;
  if (bytes > ARENA_BYTES) return nullptr;
  if (g_arena_off[d] + bytes > ARENA_BYTES) g_arena_off[d] = 0;
  void* p = g_arena[d] + g_arena_off[d];
  g_arena_off[d] += bytes;
  return p;
}
// SM count of the current device (the grid / split-K heuristics size their waves with it)
int sm_count_now() {
  static int cached[MAX_DEVICES] = {};
  const int d = current_device();
  if (!cached[d]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
    cached[d] = n;
  }
  return cached[d];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <class P, int BN, int OP>
cudaError_t launch_one(const P& p, long long M, int N, long long K, int splits, long long k_per_split, cudaStream_t stream) {
  static bool attr_set[MAX_DEVICES] = {};      // the attribute is per device: set it once on each one used
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<P, BN, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)TcCfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  if constexpr (P::TMA_B) if (K > 0) {
    constexpr int G = OpT<OP>::G, KB = 8 * G;
    constexpr size_t ES = OP == OP_BF16 ? 2 : 4;
    const long long K_pad = (K + KB - 1) / KB * KB;
    void* packed = arena_get((size_t)N * K_pad * ES);
    EncodeTiledFn enc = get_encode_fn();
    if (!packed || !enc) return cudaErrorMemoryAllocation;
    const long long items = (long long)N * (K_pad / G);
    pack_b_kernel<P, OP, std::is_same<P, TcFwd>::value><<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(p, packed, N, K, K_pad);
    cuuint64_t gdim[2] = {(cuuint64_t)K_pad, (cuuint64_t)N}, gstr[1] = {(cuuint64_t)K_pad * ES};
    cuuint32_t gbox[2] = {(cuuint32_t)KB, (cuuint32_t)BN}, estr[2] = {1, 1};
    if (enc(&tm, OP == OP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, packed, gdim, gstr, gbox,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int ntn = (N + BN - 1) / BN;
  const long long ntm = (M + TILE_M - 1) / TILE_M;
  dim3 grid((unsigned)(ntm * ntn), (unsigned)splits);
  tc_gemm_kernel<P, BN, OP><<<grid, TC_THREADS, TcCfg<BN>::SMEM_BYTES, stream>>>(p, ntn, k_per_split, tm);
  return cudaGetLastError();
}

int pick_bn(int N) {
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return (N % 256 == 0 || N > 384) ? 256 : 128;
}
// Without split-K (forward, backward-data) a short M axis leaves SMs idle: the generator's recurrent convolution is 36 M
// tiles x 2 N tiles of 256 = 72 CTAs for 148 SMs, each walking 36-144 K blocks.  Narrower N tiles re-gather the (small) A
// operand but fill the machine.
int pick_bn_grid(int N, long long M, int splits) {
  int bn = pick_bn(N);
  const long long mt = (M + TILE_M - 1) / TILE_M;
  while (bn > 32 && mt * ((N + bn - 1) / bn) * splits < sm_count_now()) bn /= 2;
  return bn;
}

template <class P, int OP>
cudaError_t launch_bn(const P& p, long long M, int N, long long K, int splits, long long kps, cudaStream_t stream) {
  switch (P::TMA_B ? pick_bn_grid(N, M, splits) : pick_bn(N)) {
    case 16: return launch_one<P, 16, OP>(p, M, N, K, splits, kps, stream);
    case 32: return launch_one<P, 32, OP>(p, M, N, K, splits, kps, stream);
    case 64: return launch_one<P, 64, OP>(p, M, N, K, splits, kps, stream);
    case 128: return launch_one<P, 128, OP>(p, M, N, K, splits, kps, stream);
    default: return launch_one<P, 256, OP>(p, M, N, K, splits, kps, stream);
  }
}
template <class P>
cudaError_t launch_tc(const P& p, long long M, int N, long long K, int splits, long long kps, int op, cudaStream_t stream) {
  return op == OP_BF16 ? launch_bn<P, OP_BF16>(p, M, N, K, splits, kps, stream) : launch_bn<P, OP_TF32>(p, M, N, K, splits, kps, stream);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int wdg_tc_conv2d_fwd(const ConvGeo& g, const float* x, const float* w, const float* bias, float* y, int accumulate, float alpha,
                      int op, cudaStream_t stream) {
  TcFwd p{g, x, w, bias, y, accumulate, 0, 0, alpha};
  p.vec_in = (g.x_cs % 4 == 0) && (g.x_co % 4 == 0) && (g.Ci % 4 == 0) && al16(x);
  p.vec_out = (g.y_cs % 4 == 0) && (g.y_co % 4 == 0) && al16(y);
  const long long M = (long long)g.N * g.Ho * g.Wo, K = (long long)g.kh * g.kw * g.Ci;
  if (M == 0) return 0;
  CKT(launch_tc(p, M, g.Co, K, 1, K, op, stream));
  return 0;
}

int wdg_tc_conv2d_bwd_data(const ConvGeo& g, const float* dy, const float* w, float* dx, int accumulate, int op,
                           cudaStream_t stream) {
  const int s = g.stride;
  for (int ry = 0; ry < s; ++ry)
    for (int rx = 0; rx < s; ++rx) {
      TcBwdData p{g, make_bwd_class(g, ry, rx), dy, w, dx, accumulate, 0, 0, 0};
      p.vec_in = (g.y_cs % 4 == 0) && (g.y_co % 4 == 0) && (g.Co % 4 == 0) && al16(dy);
      p.vec_w = (g.Co % 4 == 0) && al16(w);
      p.vec_out = (g.x_cs % 4 == 0) && (g.x_co % 4 == 0) && al16(dx);
      const long long M = (long long)g.N * p.c.Hc * p.c.Wc, K = (long long)p.c.Jy * p.c.Jx * g.Co;
      if (M == 0) continue;
      CKT(launch_tc(p, M, g.Ci, K, 1, K > 0 ? K : 1, op, stream));
    }
  return 0;
}

// Split-K plan of the backward-weight GEMM: about two waves of CTAs (2 CTAs per SM), at least 4 K blocks per split.
void wdg_tc_wgrad_plan(const ConvGeo& g, int op, int* splits_out, long long* kps_out) {
  const long long M = (long long)g.kh * g.kw * g.Ci, K = (long long)g.N * g.Ho * g.Wo;
  const int BN = pick_bn(g.Co), KB = op == OP_BF16 ? 64 : 32;
  const long long tiles = ((M + TILE_M - 1) / TILE_M) * ((g.Co + BN - 1) / BN);
  long long splits = (4ll * sm_count_now() + tiles - 1) / tiles;
  const long long max_splits = (K + 4 * KB - 1) / (4 * KB);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long kps = (K + splits - 1) / splits;
  kps = (kps + KB - 1) / KB * KB;
  splits = (K + kps - 1) / kps;
  if (splits < 1) splits = 1;
  *splits_out = (int)splits;
  *kps_out = kps;
}

int wdg_tc_conv2d_bwd_weight(const ConvGeo& g, const float* x, const float* dy, float* part, int splits, long long kps, int op,
                             cudaStream_t stream) {
  TcWgrad p{g, x, dy, part, kps, 0, 0};
  // Measured on B200 (profiles/r1_train_conv_table.md): with incremental row addressing and both gathers issued in one
  // round trip the lane-per-row scalar form beats the float4 register-transposing form for every layer wider than 16
  // output channels (-10..35 %); the float4 form keeps a 3-15 % edge for the narrow ones.  WDG_WGRAD_VEC=0/1 forces it.
  static int force = -2;
  if (force == -2) { const char* e = getenv("WDG_WGRAD_VEC"); force = e ? atoi(e) : -1; }
  const bool want = force < 0 ? g.Co <= 16 : force != 0;
  p.vec_a = want && (g.Ci % 4 == 0) && (g.x_cs % 4 == 0) && (g.x_co % 4 == 0) && al16(x);
  p.vec_b = want && (g.Co % 4 == 0) && (g.y_cs % 4 == 0) && (g.y_co % 4 == 0) && al16(dy);
  const long long M = (long long)g.kh * g.kw * g.Ci, K = (long long)g.N * g.Ho * g.Wo;
  CKT(launch_tc(p, M, g.Co, K, splits, kps, op, stream));
  return 0;
}
