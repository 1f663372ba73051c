// On-device evaluation metrics of the reference's gan/metrics.py: fused per-sample
// reductions (wind-speed-weighted RMSE, wind-speed RMSE, extreme-weighted RMSE, angular cosine distance, opposite
// cosine similarity), the log spectral distance and the spatially convolved Kolmogorov-Smirnov statistic.
// Bandwidth-bound kernels; every reduction runs in a fixed order (partials per block, combined in double).
// Tensors are fp32 channels-last [B, T, H, W, C].
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <string>

#include "../../include/wdg.h"

extern int wdg_set_error(const std::string& m);

#define CKM(call)                                                                                    \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) return wdg_set_error(std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace {

constexpr int PW_TERMS = 6;      // per-sample sums kept by the pointwise kernel
constexpr int PW_BLOCKS = 64;    // blocks per sample

__device__ __forceinline__ float nan0(float v) { return isnan(v) ? 0.f : v; }

// metrics.py:32-45, 81-91, 97-112, 66-73: one pass over (real, fake) producing, per sample b and block slab,
//   s[0] = sum tau * ((u^ - beta u)^2 + (v^ - beta v)^2)      (ws_weighted_rmse numerator, over T*H*W pixels)
//   s[1] = sum (|w| - |w^|)^2                                   (ws_rmse numerator)
//   s[2] = sum acos(clip(cos, -1, 1)) / pi                      (acd numerator)
//   s[3] = sum .5 * (1 - cos)                                   (opposite_cosine_similarity numerator)
//   s[4] = sum_c r^2 (r - f)^2                                  (extreme_weighted_rmse numerator, over T*H*W*C)
//   s[5] = sum_c r^2                                            (its batch-global denominator)
// cos follows Keras cosine_similarity: l2_normalize(x) = x * rsqrt(max(sum x^2, 1e-12)) along the channel axis.
__global__ void metrics_pointwise_kernel(const float* __restrict__ real, const float* __restrict__ fake, long long px, int C,
                                         double* __restrict__ part) {
  const int b = blockIdx.y;
  const long long per = (px + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(px, p0 + per);
  const float* r = real + (long long)b * px * C;
  const float* f = fake + (long long)b * px * C;
  double s[PW_TERMS] = {0, 0, 0, 0, 0, 0};
  for (long long p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    float rr = 0.f, ff = 0.f, rf = 0.f, ext = 0.f;
    for (int c = 0; c < C; ++c) {
      const float a = r[p * C + c], h = f[p * C + c];
      rr += a * a; ff += h * h; rf += a * h;
      ext += nan0(a * a * (a - h) * (a - h));
    }
    s[4] += ext; s[5] += rr;
    const float inv = rsqrtf(fmaxf(rr, 1e-12f)) * rsqrtf(fmaxf(ff, 1e-12f));
    const float cs = rf * inv;
    s[2] += acosf(fminf(fmaxf(cs, -1.f), 1.f)) / CUDART_PI_F;
    s[3] += 0.5f * (1.f - cs);
    if (C >= 2) {
      const float u = r[p * C], v = r[p * C + 1], uh = f[p * C], vh = f[p * C + 1];
      const float est = sqrtf(uh * uh + vh * vh), rea = sqrtf(u * u + v * v);
      const float beta = (4.f + rea) / (4.f + est);
      const float tau = est >= rea ? 0.425f : 1.f - 0.425f;
      s[0] += nan0(tau * ((uh - beta * u) * (uh - beta * u) + (vh - beta * v) * (vh - beta * v)));
      s[1] += nan0((rea - est) * (rea - est));
    }
  }
  __shared__ double sm[256];
  for (int t = 0; t < PW_TERMS; ++t) {
    sm[threadIdx.x] = s[t];
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) part[((long long)b * gridDim.x + blockIdx.x) * PW_TERMS + t] = sm[0];
    __syncthreads();
  }
}
// out[m][b], m = ws_weighted_rmse, ws_rmse, acd, opposite_cosine_similarity, extreme_rmse
__global__ void metrics_pointwise_finish_kernel(const double* __restrict__ part, int B, int blocks, long long px,
                                                float* __restrict__ out) {
  const int b = threadIdx.x;
  if (b >= B) return;
  double s[PW_TERMS] = {0, 0, 0, 0, 0, 0}, den = 0;
  for (int k = 0; k < blocks; ++k)
    for (int t = 0; t < PW_TERMS; ++t) s[t] += part[((long long)b * blocks + k) * PW_TERMS + t];
  for (int bb = 0; bb < B; ++bb)
    for (int k = 0; k < blocks; ++k) den += part[((long long)bb * blocks + k) * PW_TERMS + 5];
  out[0 * B + b] = (float)sqrt(s[0] / (double)px);
  out[1 * B + b] = (float)sqrt(s[1] / (double)px);
  out[2 * B + b] = (float)(s[2] / (double)px);
  out[3 * B + b] = (float)(s[3] / (double)px);
  out[4 * B + b] = den != 0 ? (float)sqrt(s[4] / den) : 0.f;     // divide_no_nan
}

// metrics.py:121-137.  tf.signal.rfft2d acts on the two INNERMOST axes of the [B,T,H,W,C] tensor, i.e. (W, C): a real
// FFT of length C along the channels followed by a complex FFT of length W along the image rows (the reference
// transposes only afterwards).  With C = 2 the channel transform is (u + v, u - v).  One block per image row
// (b, t, h): direct O(W^2) DFT of the C/2+1 channel-frequency sequences, twiddles from a shared table.
__global__ void metrics_lsd_kernel(const float* __restrict__ real, const float* __restrict__ fake, int W, int C, long long rows,
                                   long long rows_per_sample, double* __restrict__ part) {
  extern __shared__ float sh[];
  float* cs = sh;              // cos(2 pi k / W)
  float* sn = sh + W;          // sin(2 pi k / W)
  float* xr = sh + 2 * W;      // [2 tensors][CF][W] real parts of the channel transform
  float* xi = xr + 2 * (C / 2 + 1) * W;
  const int CF = C / 2 + 1;
  const long long row = blockIdx.x;
  for (int k = threadIdx.x; k < W; k += blockDim.x) {
    float s, c;
    sincospif(2.f * (float)k / (float)W, &s, &c);
    cs[k] = c; sn[k] = s;
  }
  // channel rfft (length C) per pixel: X_q = sum_c x_c exp(-2 pi i q c / C)
  for (int i = threadIdx.x; i < 2 * CF * W; i += blockDim.x) {
    const int w = i % W, q = (i / W) % CF, which = i / (W * CF);
    const float* x = (which ? fake : real) + (row * W + w) * C;
    float re = 0.f, im = 0.f;
    for (int c = 0; c < C; ++c) {
      float s, co;
      sincospif(-2.f * (float)((q * c) % C) / (float)C, &s, &co);
      re += x[c] * co; im += x[c] * s;
    }
    xr[i] = re; xi[i] = im;
  }
  __syncthreads();
  const float eps = 1e-7f;     // keras.backend.epsilon()
  double acc = 0;
  for (int i = threadIdx.x; i < CF * W; i += blockDim.x) {
    const int k = i % W, q = i / W;
    float p[2];
    for (int which = 0; which < 2; ++which) {
      const float* ar = xr + (which * CF + q) * W;
      const float* ai = xi + (which * CF + q) * W;
      float re = 0.f, im = 0.f;
      int idx = 0;
      for (int w = 0; w < W; ++w) {            // exp(-2 pi i k w / W) = cs[idx] - i sn[idx], idx = k w mod W
        re += ar[w] * cs[idx] + ai[w] * sn[idx];
        im += ai[w] * cs[idx] - ar[w] * sn[idx];
        idx += k; if (idx >= W) idx -= W;
      }
      p[which] = re * re + im * im;
    }
    const float den = p[1] + eps;
    const float ratio = den != 0.f ? (p[0] + eps) / den : 0.f;
    const float l = 10.f * (logf(ratio) / logf(10.f));
    acc += (double)(l * l);
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[row] = red[0];
}
__global__ void metrics_lsd_finish_kernel(const double* __restrict__ part, int B, long long rows_per_sample, long long count,
                                          float* __restrict__ out) {
  const int b = blockIdx.x;
  __shared__ double red[256];
  double s = 0;
  for (long long r = threadIdx.x; r < rows_per_sample; r += 256) s += part[(long long)b * rows_per_sample + r];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) {
    const float v = (float)sqrt(red[0] / (double)count);
    out[b] = isnan(v) ? 0.f : v;
  }
}

// metrics.py:155-187.  For every (time, channel, sample) image and every PxP window (stride 1, VALID) the two-sample
// KS statistic of the window's real vs fake values evaluated at the 100 points linspace(-30, 30, 100); the result
// is the mean KS image over (time, channel, sample).  Each thread owns one window: bin index k0(x) = first point
// with x <= point (exact float comparisons against the table), a signed 101-bin difference histogram, and its
// running sum gives cdf_real - cdf_fake at every point.  Images are visited in a fixed order (deterministic mean).
constexpr int KS_POINTS = 100;
constexpr int KS_TILE = 16;
__constant__ float c_ks_points[KS_POINTS];

__device__ __forceinline__ int ks_bin(float x) {
  // first k with x <= points[k]; KS_POINTS if none (also for NaN)
  if (!(x == x)) return KS_POINTS;
  int k = (int)fminf(fmaxf(ceilf((x + 30.f) * (99.f / 60.f)), 0.f), (float)KS_POINTS);
  while (k > 0 && x <= c_ks_points[k - 1]) --k;
  while (k < KS_POINTS && !(x <= c_ks_points[k])) ++k;
  return k;
}

__global__ void metrics_ks_kernel(const float* __restrict__ real, const float* __restrict__ fake, int B, int T, int H, int W,
                                  int C, int P, float* __restrict__ out) {
  extern __shared__ unsigned char ks_sm[];
  const int TW = KS_TILE + P - 1;
  unsigned char* bins_r = ks_sm;                       // [TW*TW] bin index of every pixel of the tile, real
  unsigned char* bins_f = bins_r + TW * TW;            // fake
  signed char* hist = reinterpret_cast<signed char*>(bins_f + TW * TW);   // [256 threads][KS_POINTS + 1]
  const int Ho = H - P + 1, Wo = W - P + 1;
  const int tx = threadIdx.x % KS_TILE, ty = threadIdx.x / KS_TILE;
  const int ox = blockIdx.x * KS_TILE + tx, oy = blockIdx.y * KS_TILE + ty;
  const bool valid = ox < Wo && oy < Ho;
  signed char* h = hist + threadIdx.x * (KS_POINTS + 1);
  double acc = 0;
  const int n_img = T * C * B;
  for (int img = 0; img < n_img; ++img) {
    // order of the reference's list: time outer, channel inner, then the batch axis of each entry
    const int b = img % B, ch = (img / B) % C, t = img / (B * C);
    const long long base = ((long long)b * T + t) * H * W;
    for (int i = threadIdx.x; i < TW * TW; i += blockDim.x) {
      const int yy = blockIdx.y * KS_TILE + i / TW, xx = blockIdx.x * KS_TILE + i % TW;
      int kr = KS_POINTS, kf = KS_POINTS;
      if (yy < H && xx < W) {
        kr = ks_bin(real[(base + (long long)yy * W + xx) * C + ch]);
        kf = ks_bin(fake[(base + (long long)yy * W + xx) * C + ch]);
      }
      bins_r[i] = (unsigned char)kr; bins_f[i] = (unsigned char)kf;
    }
    __syncthreads();
    if (valid) {
      for (int k = 0; k <= KS_POINTS; ++k) h[k] = 0;
      for (int dy = 0; dy < P; ++dy)
        for (int dx = 0; dx < P; ++dx) {
          const int i = (ty + dy) * TW + tx + dx;
          ++h[bins_r[i]]; --h[bins_f[i]];
        }
      int run = 0, best = 0;
      for (int k = 0; k < KS_POINTS; ++k) { run += h[k]; best = max(best, abs(run)); }
      acc += (double)((float)best / (float)(P * P));
    }
    __syncthreads();
  }
  if (valid) out[oy * Wo + ox] = (float)(acc / (double)n_img);
}

}  // namespace

extern "C" int wdg_metrics_pointwise_scratch(int B, size_t* bytes) {
  if (!bytes || B <= 0) return wdg_set_error("wdg_metrics_pointwise_scratch: bad argument");
  *bytes = (size_t)B * PW_BLOCKS * PW_TERMS * sizeof(double);
  return 0;
}

extern "C" int wdg_metrics_pointwise(const float* real, const float* fake, int B, long long pixels_per_sample, int C, float* out,
                                     void* scratch, void* stream_) {
  if (!real || !fake || !out || !scratch || B <= 0 || B > 1024 || pixels_per_sample <= 0 || C <= 0)
    return wdg_set_error("wdg_metrics_pointwise: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  metrics_pointwise_kernel<<<dim3(PW_BLOCKS, B), 256, 0, stream>>>(real, fake, pixels_per_sample, C, (double*)scratch);
  CKM(cudaGetLastError());
  metrics_pointwise_finish_kernel<<<1, 1024, 0, stream>>>((const double*)scratch, B, PW_BLOCKS, pixels_per_sample, out);
  CKM(cudaGetLastError());
  return 0;
}

extern "C" int wdg_metric_lsd(const float* real, const float* fake, int B, int T, int H, int W, int C, float* out, void* scratch,
                              void* stream_) {
  if (!real || !fake || !out || !scratch || B <= 0 || T <= 0 || H <= 0 || W <= 0 || C <= 0)
    return wdg_set_error("wdg_metric_lsd: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int CF = C / 2 + 1;
  const size_t smem = (size_t)(2 * W + 4 * CF * W) * sizeof(float);
  if (smem > 200 * 1024) return wdg_set_error("wdg_metric_lsd: row too wide for the shared-memory DFT");
  if (smem > 48 * 1024) CKM(cudaFuncSetAttribute(metrics_lsd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long rows = (long long)B * T * H, rows_per_sample = (long long)T * H;
  metrics_lsd_kernel<<<(unsigned)rows, 256, smem, stream>>>(real, fake, W, C, rows, rows_per_sample, (double*)scratch);
  CKM(cudaGetLastError());
  metrics_lsd_finish_kernel<<<B, 256, 0, stream>>>((const double*)scratch, B, rows_per_sample, rows_per_sample * W * CF, out);
  CKM(cudaGetLastError());
  return 0;
}

extern "C" int wdg_metric_spatial_ks(const float* real, const float* fake, int B, int T, int H, int W, int C, int patch, float* out,
                                     void* stream_) {
  if (!real || !fake || !out || B <= 0 || T <= 0 || C <= 0 || patch <= 0 || patch > H || patch > W || patch * patch > 127)
    return wdg_set_error("wdg_metric_spatial_ks: bad argument (patch*patch must be <= 127)");
  cudaStream_t stream = (cudaStream_t)stream_;
  static bool table = false;
  if (!table) {
    float pts[KS_POINTS];
    for (int k = 0; k < KS_POINTS; ++k) pts[k] = (float)(-30.0 + (60.0 / 99.0) * k);   // np.linspace(-30., 30., 100)
    pts[KS_POINTS - 1] = 30.f;
    CKM(cudaMemcpyToSymbol(c_ks_points, pts, sizeof pts));
    table = true;
  }
  const int TW = KS_TILE + patch - 1;
  const size_t smem = (size_t)2 * TW * TW + (size_t)KS_TILE * KS_TILE * (KS_POINTS + 1);
  if (smem > 48 * 1024) CKM(cudaFuncSetAttribute(metrics_ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int Ho = H - patch + 1, Wo = W - patch + 1;
  metrics_ks_kernel<<<dim3((Wo + KS_TILE - 1) / KS_TILE, (Ho + KS_TILE - 1) / KS_TILE), KS_TILE * KS_TILE, smem, stream>>>(
      real, fake, B, T, H, W, C, patch, out);
  CKM(cudaGetLastError());
  return 0;
}
