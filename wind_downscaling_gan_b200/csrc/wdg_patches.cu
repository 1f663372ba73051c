// Device-side tiling driver of the reference's predict() (api.py:117-129,140-151): patch gather with
// reversed latitude, per-(column, channel) normalisation statistics, and the cropped overlap-mean stitch.
// Integer semantics follow oracle/patches.py, which is pinned against the reference's own lines.
#include <cuda_runtime.h>

#include <string>

#include "../../include/wdg.h"

extern int wdg_set_error(const std::string& m);  // wdg_generator.cu

#define CKP(call)                                                                                 \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess) return wdg_set_error(std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace {

struct Geo {
  const float* u10;   // (T_total, uh, uw): hi-res (uh = H, uw = W, maps null) or coarse + nearest-neighbour maps
  const float* v10;
  const float* elev;  // (eh, ew)
  const int* sx;      // device, nx patch column starts
  const int* sy;      // device, ny patch row starts
  int T_total, H, W, nx, ny, seq, img, ntimeseq;
  // optional nearest-neighbour regrid folded into the gather (api.py:31-43): hi-res (row, col) -> source (row, col)
  const int* uv_rmap; const int* uv_cmap; int uh, uw;
  const int* e_rmap; const int* e_cmap; int eh, ew;
  float elev_div;     // elevation is divided by this (1 when already in km; 1e3 on the fused path, api.py:96)
};

// api.py:119: patch row p -> domain row; rows sy+img-1 .. sy, or img .. 1 when sy == 0 (F10)
__device__ __forceinline__ int domain_row(int sy, int p, int img) { return (sy != 0 ? sy + img - 1 : img) - p; }

__device__ __forceinline__ float load_var(const Geo& g, int var, int t, int row, int col) {
  if (var == 2) {
    const int r = g.e_rmap ? g.e_rmap[row] : row, c = g.e_cmap ? g.e_cmap[col] : col;
    const float e = g.elev[(long long)r * g.ew + c];
    return g.elev_div == 1.f ? e : e / g.elev_div;   // fp32 division, as numpy does for a float32 DEM (api.py:96)
  }
  const float* src = var == 0 ? g.u10 : g.v10;
  const int r = g.uv_rmap ? g.uv_rmap[row] : row, c = g.uv_cmap ? g.uv_cmap[col] : col;
  return src[((long long)t * g.uh + r) * g.uw + c];
}

// One block per (var, ix, iy, k) x timestep (blockIdx.y: a block per whole sequence left 60 blocks walking 2304 dependent
// map -> field loads each); thread j owns patch column j (coalesced along the domain's lon axis).
// pass 0: partial[blk][t][j] = (sum x, count of non-NaN); pass 1: sum (x - mean)^2.
__global__ void stats_partial_kernel(Geo g, int pass, const double* __restrict__ mean, double* __restrict__ psum,
                                     double* __restrict__ pcnt) {
  const int j = threadIdx.x;
  int b = blockIdx.x;
  const int t = blockIdx.y;
  const int k = b % g.ntimeseq; b /= g.ntimeseq;
  const int iy = b % g.ny; b /= g.ny;
  const int ix = b % g.nx; b /= g.nx;
  const int var = b;
  if (j >= g.img) return;
  const int col = g.sx[ix] + j, sy = g.sy[iy];
  const double mu = pass ? mean[j * 3 + var] : 0.0;
  double s = 0.0, c = 0.0;
  for (int p = 0; p < g.img; ++p) {
    const float x = load_var(g, var, k * g.seq + t, domain_row(sy, p, g.img), col);
    if (x == x) {
      const double d = (double)x - mu;
      s += pass ? d * d : d;
      c += 1.0;
    }
  }
  psum[((long long)blockIdx.x * g.seq + t) * g.img + j] = s;
  pcnt[((long long)blockIdx.x * g.seq + t) * g.img + j] = c;
}

// Fixed-order final reduction: one WARP per (j, var): lane l sums partials l, l+32, ... in order, then a butterfly
// (the same tree on every run: deterministic) -> mean (pass 0) or std (pass 1).  A single thread per output walked the
// 480 partials of a 20-patch window serially: 174 us per pass, more than the gather itself.
__global__ void stats_final_kernel(int img, int blocks_per_var, int pass, const double* __restrict__ psum,
                                   const double* __restrict__ pcnt, double* __restrict__ out) {
  const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= img * 3) return;                      // whole warps leave together
  const int j = w / 3, var = w % 3;
  double s = 0.0, c = 0.0;
  for (int b = lane; b < blocks_per_var; b += 32) {
    s += psum[((long long)var * blocks_per_var + b) * img + j];
    c += pcnt[((long long)var * blocks_per_var + b) * img + j];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) out[j * 3 + var] = pass ? sqrt(s / c) : s / c;
}

// out (N, seq, img, img, 3) fp32 = (x - mean[col, var]) / std[col, var], computed in fp64 (api.py:128-129).
__global__ void gather_normalise_kernel(Geo g, const double* __restrict__ mean, const double* __restrict__ stdv,
                                        float* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % g.img);
  const int p = (int)((i / g.img) % g.img);
  const int t = (int)((i / ((long long)g.img * g.img)) % g.seq);
  const long long n = i / ((long long)g.img * g.img * g.seq);
  const int k = (int)(n % g.ntimeseq);
  const int iy = (int)((n / g.ntimeseq) % g.ny);
  const int ix = (int)(n / ((long long)g.ntimeseq * g.ny));
  const int row = domain_row(g.sy[iy], p, g.img), col = g.sx[ix] + j;
#pragma unroll
  for (int var = 0; var < 3; ++var) {
    const double x = (double)load_var(g, var, k * g.seq + t, row, col);
    out[i * 3 + var] = (float)((x - mean[j * 3 + var]) / stdv[j * 3 + var]);
  }
}

// Overlap mean of the cropped patches (api.py:148-150) as a gather: one thread per output value.  The mean is
// pandas' Cython `group_mean` (what `groupby(level=...).mean()` runs): a Kahan-compensated running sum of the
// contributions in order of appearance (= patch order: sx-major, then sy), divided by their count, in the template's
// floating type ACC -- double for pandas 1.3.3 (the reference's pin; float32 columns are upcast first and the result
// is cast back), float for pandas >= 1.5.  The three Kahan statements must not be re-associated or contracted:
// they contain no multiply, and this file is built without fast-math.
// out (C, T_total', nrows, ncols) with T_total' = ntimeseq * seq.
template <typename ACC>
__global__ void stitch_kernel(const float* __restrict__ pred, const int* __restrict__ sx, const int* __restrict__ sy,
                              int nx, int ny, int ntimeseq, int seq, int img, int crop, int C,
                              const int* __restrict__ rows, int nrows, const int* __restrict__ cols, int ncols,
                              float* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % ncols);
  const int ri = (int)((i / ncols) % nrows);
  const int tg = (int)((i / ((long long)ncols * nrows)) % (ntimeseq * seq));
  const int ch = (int)(i / ((long long)ncols * nrows * ntimeseq * seq));
  const int r = rows[ri], c = cols[ci];
  const int k = tg / seq, t = tg % seq;
  ACC sumx = 0, comp = 0;
  int cnt = 0;
  for (int ix = 0; ix < nx; ++ix) {
    const int pc = c - sx[ix];
    if (pc < crop || pc >= img - crop) continue;
    for (int iy = 0; iy < ny; ++iy) {
      const int s = sy[iy];
      const int pr = (s != 0 ? s + img - 1 : img) - r;
      if (pr < crop || pr >= img - crop) continue;
      const long long n = ((long long)ix * ny + iy) * ntimeseq + k;
      const ACC val = (ACC)pred[(((n * seq + t) * img + pr) * img + pc) * C + ch];
      const ACC y = val - comp;
      const ACC tt = sumx + y;
      comp = (tt - sumx) - y;
      sumx = tt;
      ++cnt;
    }
  }
  out[i] = (float)(sumx / (ACC)cnt);
}

}  // namespace

extern "C" int wdg_patch_scratch_bytes(int nx, int ny, int ntimeseq, int img, size_t* bytes) {
  if (!bytes || nx <= 0 || ny <= 0 || ntimeseq <= 0 || img <= 0) return wdg_set_error("bad argument");
  *bytes = (size_t)2 * 3 * nx * ny * ntimeseq * WDG_MAX_SEQ * img * sizeof(double);   // one partial per (var, patch, timestep)
  return 0;
}

static int gather_normalise_impl(Geo g, double* mean_dev, double* std_dev, float* out_dev, void* scratch_dev, void* stream_);

extern "C" int wdg_gather_normalise(const float* u10_dev, const float* v10_dev, const float* elev_km_dev, int T_total,
                                    int H, int W, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny,
                                    int seq, int img, double* mean_dev, double* std_dev, float* out_dev,
                                    void* scratch_dev, void* stream_) {
  if (!u10_dev || !v10_dev || !elev_km_dev || !starts_x_dev || !starts_y_dev || !mean_dev || !std_dev || !out_dev ||
      !scratch_dev)
    return wdg_set_error("null argument");
  Geo g{u10_dev, v10_dev, elev_km_dev, starts_x_dev, starts_y_dev, T_total, H, W, nx, ny, seq, img, seq > 0 ? T_total / seq : 0,
        nullptr, nullptr, H, W, nullptr, nullptr, H, W, 1.f};
  return gather_normalise_impl(g, mean_dev, std_dev, out_dev, scratch_dev, stream_);
}

extern "C" int wdg_gather_normalise_regrid(const float* u10_coarse_dev, const float* v10_coarse_dev, int T_total, int uh, int uw,
                                           const int* uv_row_map_dev, const int* uv_col_map_dev, const float* dem_dev, int eh,
                                           int ew, const int* dem_row_map_dev, const int* dem_col_map_dev, float dem_divisor,
                                           int H, int W, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny,
                                           int seq, int img, double* mean_dev, double* std_dev, float* out_dev,
                                           void* scratch_dev, void* stream_) {
  if (!u10_coarse_dev || !v10_coarse_dev || !uv_row_map_dev || !uv_col_map_dev || !dem_dev || !dem_row_map_dev ||
      !dem_col_map_dev || !starts_x_dev || !starts_y_dev || !mean_dev || !std_dev || !out_dev || !scratch_dev)
    return wdg_set_error("null argument");
  Geo g{u10_coarse_dev, v10_coarse_dev, dem_dev, starts_x_dev, starts_y_dev, T_total, H, W, nx, ny, seq, img,
        seq > 0 ? T_total / seq : 0, uv_row_map_dev, uv_col_map_dev, uh, uw, dem_row_map_dev, dem_col_map_dev, eh, ew, dem_divisor};
  return gather_normalise_impl(g, mean_dev, std_dev, out_dev, scratch_dev, stream_);
}

static int gather_normalise_impl(Geo g, double* mean_dev, double* std_dev, float* out_dev, void* scratch_dev, void* stream_) {
  const int nx = g.nx, ny = g.ny, img = g.img, seq = g.seq;
  if (img > 1024) return wdg_set_error("img too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (g.ntimeseq <= 0) return wdg_set_error("time window shorter than one sequence");
  if (seq > WDG_MAX_SEQ) return wdg_set_error("sequence length above WDG_MAX_SEQ");
  const int bpv = nx * ny * g.ntimeseq;
  double* psum = (double*)scratch_dev;
  double* pcnt = psum + (size_t)3 * bpv * seq * img;
  const int threads = (img + 31) / 32 * 32;
  for (int pass = 0; pass < 2; ++pass) {
    stats_partial_kernel<<<dim3(3 * bpv, seq), threads, 0, stream>>>(g, pass, mean_dev, psum, pcnt);
    CKP(cudaGetLastError());
    stats_final_kernel<<<(img * 3 * 32 + 127) / 128, 128, 0, stream>>>(img, bpv * seq, pass, psum, pcnt, pass ? std_dev : mean_dev);
    CKP(cudaGetLastError());
  }
  const long long total = (long long)bpv * seq * img * img;
  gather_normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(g, mean_dev, std_dev, out_dev, total);
  CKP(cudaGetLastError());
  return 0;
}

extern "C" int wdg_stitch_accum(const float* pred_dev, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny,
                                int ntimeseq, int seq, int img, int crop, int channels, const int* rows_dev, int nrows,
                                const int* cols_dev, int ncols, float* out_dev, int accum, void* stream_) {
  if (!pred_dev || !starts_x_dev || !starts_y_dev || !rows_dev || !cols_dev || !out_dev)
    return wdg_set_error("null argument");
  if (accum != WDG_STITCH_F64 && accum != WDG_STITCH_F32) return wdg_set_error("accum must be WDG_STITCH_F64 or WDG_STITCH_F32");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long total = (long long)channels * ntimeseq * seq * nrows * ncols;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (accum == WDG_STITCH_F64)
    stitch_kernel<double><<<grid, 256, 0, stream>>>(pred_dev, starts_x_dev, starts_y_dev, nx, ny, ntimeseq, seq, img, crop,
                                                    channels, rows_dev, nrows, cols_dev, ncols, out_dev, total);
  else
    stitch_kernel<float><<<grid, 256, 0, stream>>>(pred_dev, starts_x_dev, starts_y_dev, nx, ny, ntimeseq, seq, img, crop,
                                                   channels, rows_dev, nrows, cols_dev, ncols, out_dev, total);
  CKP(cudaGetLastError());
  return 0;
}

extern "C" int wdg_stitch(const float* pred_dev, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny,
                          int ntimeseq, int seq, int img, int crop, int channels, const int* rows_dev, int nrows,
                          const int* cols_dev, int ncols, float* out_dev, void* stream_) {
  return wdg_stitch_accum(pred_dev, starts_x_dev, nx, starts_y_dev, ny, ntimeseq, seq, img, crop, channels, rows_dev, nrows,
                          cols_dev, ncols, out_dev, WDG_STITCH_F64, stream_);
}
