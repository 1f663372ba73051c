// Host side of libwdg.so: weight packing, TMA descriptor / launch-plan construction and the
// C ABI declared in include/wdg.h for the generator forward (reference models.py:9-73).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wdg.h"
#include "conv_umma.cuh"
#include "conv_umma2.cuh"
#include "halo_conv.cuh"
#include "stencil_kernels.cuh"

using namespace wdg;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) {
  g_err = m;
  return 1;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));        \
  } while (0)

int wdg_set_error(const std::string& m) { return fail(m); }  // shared with wdg_patches.cu

extern "C" const char* wdg_last_error(void) { return g_err.c_str(); }

extern "C" int wdg_device_info(int device, int* sm_major, int* sm_minor, int* sm_count) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (sm_major) *sm_major = prop.major;
  if (sm_minor) *sm_minor = prop.minor;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return 0;
}

// --------------------------------------------------------- TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
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

// Tensor map of rank 2..5 over bf16 (esz 2) or fp32 (esz 4: tf32 operands) elements.  dims/box in elements (dim 0
// innermost), strides in ELEMENTS for dims 1..rank-1.
int wdg_make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                  const uint32_t* box, int inner_bytes, int esz) {   // shared with train_lstm16.cu
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_el[i - 1] * (uint64_t)esz;
  }
  CUtensorMapSwizzle sw = inner_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : inner_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                              : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(tm, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gdim, gstr, gbox, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
             box[0], box[1], rank > 2 ? box[2] : 0);
    return fail(buf);
  }
  return 0;
}

// ------------------------------------------------------------- the handle
struct ConvLaunch {
  CUtensorMap tmA[3];
  CUtensorMap tmB;
  ConvParams p;
  int bn;
  int epi;
  int grid;
  int shallow = 0;   // 1: 3-stage ring, two CTAs per SM (short-K layers)
};

struct WeightDesc {
  std::string name;
  std::vector<int64_t> dims;
  int64_t count() const {
    int64_t c = 1;
    for (auto d : dims) c *= d;
    return c;
  }
};

struct Plan {
  int B = 0, T = 0;
  uint8_t *xpad, *res4, *hseq, *g5, *catp, *edgeE, *g9;   // act_t buffers (bf16 or tf32-in-fp32), addressed in bytes
  float *cstate, *deltaD;
  ConvLaunch L0, L2, L5, L7, LE, L9;
  std::vector<ConvLaunch> LS;  // ConvLSTM, one launch per timestep (fallback)
  ConvLaunch LSP;              // ConvLSTM, all timesteps in ONE persistent cooperative launch (conv_umma.cuh: sync_flags)
  bool lstm_persist = false;
  unsigned long long* lstm_flags = nullptr;   // [T] step counters in the workspace
  // the same launch on CTA pairs (conv_umma2.cuh: tcgen05.mma.cta_group::2, B tile split over the two SMs of a TPC)
  bool lstm_pair = false;
  bool l2_pair = false;        // 4x4 stride-2 convolution on CTA pairs (conv_pair_kernel<128, EPI_AFFINE>)
  CUtensorMap l2B_half;
  int l2_pair_grid = 0;
  CUtensorMap tmB_half;        // B boxes of 128 weight rows
  ConvParams pair_p;
  int pair_grid = 0;
  bool use_halo = false;       // halo-reuse kernels (halo_conv.cuh) for the 8x8 s2 conv and the fused upsample conv
  CUtensorMap hA, hB, h0A, h0B, h11A, h11B, h5A, h5B;
  HaloParams hp, h0p, h11p, h5p;
  int hgrid = 0, h0grid = 0, h11grid = 0, h5grid = 0;
  // CTA-pair form (halo_conv.cuh PAIR) of the two big halo kernels: B boxes of BN/2 rows, clusters of 2
  bool halo_pair = false;
  CUtensorMap hB_half, h0B_half, h5B_half;
  int hgrid_pair = 0, h0grid_pair = 0, h5grid_pair = 0;
  bool use_halo11 = false;     // final 3x3 conv on the tensor cores (super-pixel form)
  bool use_halo5 = false;      // 3x3 conv on the (zero-ring padded) hidden-state sequence with halo reuse
  int launches = 0;
};

struct wdg_generator {
  int S, cin, cnoise, cout, T_default, F;
  int CP;                 // padded input channels of the packed image
  int prec = PREC_BF16;   // operand precision of the GEMM stages (wdg_generator_set_precision)
  int esz() const { return prec == PREC_TF32 ? 4 : 2; }     // bytes per activation / packed-weight element
  int kbe() const { return prec == PREC_TF32 ? 32 : 64; }   // elements per 128-byte K-block row
  int catp_pitch() const { return prec == PREC_TF32 ? 160 : 192; }   // channels per pixel of the concat image
  int device = 0;
  int sm_count = 148;
  std::vector<WeightDesc> descs;
  std::map<std::string, std::vector<float>> w;
  std::map<std::string, bool> set_;
  bool finalized = false;
  // packed device weights
  void *B0 = nullptr, *B2 = nullptr, *BL = nullptr, *B5 = nullptr, *B7 = nullptr, *B9 = nullptr, *B9h = nullptr, *BE = nullptr, *B11 = nullptr, *B5h = nullptr;
  float* fparams = nullptr;  // all fp32 per-column vectors, see offsets
  float *bias0, *sc0, *sh0, *bias2, *sc2, *sh2, *biasL, *bias5, *sc5, *sh5, *bias7, *sc7, *sh7, *bias9, *sc9, *sh9,
      *w11, *b11;
  // plans: [0] = the bound (B, T); [1], [2] = chunk / tail plans used by the pipelined host entry point.
  // Every plan owns a disjoint region of the caller's workspace, so the zero rings of its padded buffers stay zero.
  Plan plans[1];   // [0] = the bound (B, T)
  // host pipeline (predict_host*): the batch is cut into pieces, each with its own plan (one per distinct size)
  std::map<int, Plan> piece_plans;
  std::vector<int> sched_host, sched_dev;   // piece sizes: host-supplied noise (H2D-bound) / noise drawn on the device
  int max_piece = 0;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_h2d[2] = {}, ev_fwd[2] = {}, ev_d2h[2] = {};
  FinalConvW<16, 2> w11h;      // final 3x3 conv weights, passed by value (constant bank)
  float *zero48 = nullptr, *one48 = nullptr;
  // optional per-stage CUDA-event timing (bench.py roofline)
  bool profiling = false;
  cudaEvent_t ev[WDG_NUM_STAGES + 1] = {};
};

static const char* LWF = "layer_with_weights-%d/%s";
static std::string wname(int i, const char* leaf) {
  char b[96];
  snprintf(b, sizeof b, LWF, i, leaf);
  return b;
}

extern "C" int wdg_generator_create(wdg_generator** out, int image_size, int in_channels, int noise_channels,
                                    int out_channels, int n_timesteps, int feature_channels) {
  if (!out) return fail("null out");
  if (image_size % 4 != 0) return fail("image_size % 4 != 0 (models.py:19)");
  if (feature_channels % 8 != 0) return fail("feature_channels % 8 != 0 (models.py:20)");
  const int cin_total = in_channels + noise_channels;
  const int f0 = cin_total * 8 <= feature_channels ? cin_total * 8 : feature_channels;  // models.py:31
  if (feature_channels != 128 || f0 != 128)
    return fail("sm_100a kernels are built for feature_channels == 128 and 8*(in+noise) >= 128 (api.py:22-28)");
  if (feature_channels / 8 < out_channels) return fail("feature_channels/8 < out_channels: reference branch models.py:66-68 is broken; not built");
  if (out_channels != 2) return fail("final convolution kernel is instantiated for out_channels == 2 (api.py:28)");
  if (image_size % 32 != 0) return fail("image_size must be a multiple of 32 for the tile shapes used");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("no CUDA device: libwdg has no CPU fallback");
  auto* g = new wdg_generator();
  g->S = image_size; g->cin = in_channels; g->cnoise = noise_channels; g->cout = out_channels;
  g->T_default = n_timesteps; g->F = feature_channels;
  g->CP = (cin_total + 7) / 8 * 8;
  cudaGetDevice(&g->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, g->device) != cudaSuccess) { delete g; return fail("cudaGetDeviceProperties failed"); }
  if (prop.major != 10) { delete g; return fail("libwdg requires an sm_100 (B200) device"); }
  g->sm_count = prop.multiProcessorCount;
  const int64_t F = feature_channels, C = cin_total, O = out_channels;
  auto add = [&](int i, const char* leaf, std::vector<int64_t> d) { g->descs.push_back({wname(i, leaf), d}); };
  auto bn = [&](int i, int64_t c) {
    add(i, "gamma", {c}); add(i, "beta", {c}); add(i, "moving_mean", {c}); add(i, "moving_variance", {c});
  };
  add(0, "layer/w", {8, 8, C, f0}); add(0, "layer/layer/bias", {f0}); add(0, "layer/sn_u", {1, f0}); bn(1, f0);
  add(2, "layer/w", {4, 4, f0, F}); add(2, "layer/layer/bias", {F}); add(2, "layer/sn_u", {1, F}); bn(3, F);
  add(4, "cell/kernel", {3, 3, F, 4 * F}); add(4, "cell/recurrent_kernel", {3, 3, F, 4 * F}); add(4, "cell/bias", {4 * F});
  add(5, "layer/w", {3, 3, F, F / 2}); add(5, "layer/layer/bias", {F / 2}); add(5, "layer/sn_u", {1, F / 2}); bn(6, F / 2);
  add(7, "layer/w", {2, 2, F / 4, F / 2 + F}); add(7, "layer/layer/bias", {F / 4}); add(7, "layer/sn_u", {1, F / 2 + F}); bn(8, F / 4);
  add(9, "layer/kernel", {5, 5, F / 8, F / 4 + f0}); add(9, "layer/bias", {F / 8}); bn(10, F / 8);
  add(11, "layer/kernel", {3, 3, F / 8, O}); add(11, "layer/bias", {O});
  for (auto& d : g->descs) g->w[d.name] = std::vector<float>((size_t)d.count(), 0.f);
  *out = g;
  return 0;
}

extern "C" void wdg_generator_destroy(wdg_generator* g) {
  if (!g) return;
  for (auto& e : g->ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (g->ev_h2d[i]) cudaEventDestroy(g->ev_h2d[i]);
    if (g->ev_fwd[i]) cudaEventDestroy(g->ev_fwd[i]);
    if (g->ev_d2h[i]) cudaEventDestroy(g->ev_d2h[i]);
  }
  if (g->copy_in) cudaStreamDestroy(g->copy_in);
  if (g->copy_out) cudaStreamDestroy(g->copy_out);
  cudaFree(g->B0); cudaFree(g->B2); cudaFree(g->BL); cudaFree(g->B5); cudaFree(g->B7); cudaFree(g->B9); cudaFree(g->B9h); cudaFree(g->BE); cudaFree(g->B11); cudaFree(g->B5h);
  cudaFree(g->fparams);
  delete g;
}

extern "C" int wdg_generator_num_weights(const wdg_generator* g) { return g ? (int)g->descs.size() : 0; }

extern "C" int wdg_generator_weight_info(const wdg_generator* g, int index, const char** name, int64_t* dims, int* ndim) {
  if (!g || index < 0 || index >= (int)g->descs.size()) return fail("weight index out of range");
  const WeightDesc& d = g->descs[index];
  if (name) *name = d.name.c_str();
  if (ndim) *ndim = (int)d.dims.size();
  if (dims) for (size_t i = 0; i < d.dims.size(); ++i) dims[i] = d.dims[i];
  return 0;
}

extern "C" int wdg_generator_set_weight(wdg_generator* g, const char* name, const float* host_data, const int64_t* dims, int ndim) {
  if (!g || !name || !host_data) return fail("null argument");
  for (auto& d : g->descs) {
    if (d.name != name) continue;
    if ((int)d.dims.size() != ndim) return fail(std::string("rank mismatch for ") + name);
    for (int i = 0; i < ndim; ++i)
      if (d.dims[i] != dims[i]) return fail(std::string("shape mismatch for ") + name);
    std::memcpy(g->w[d.name].data(), host_data, sizeof(float) * (size_t)d.count());
    g->set_[d.name] = true;
    g->finalized = false;
    return 0;
  }
  return fail(std::string("unknown weight name: ") + name);
}

extern "C" int wdg_generator_get_weight(const wdg_generator* g, const char* name, float* host_data, int64_t count) {
  if (!g || !name || !host_data) return fail("null argument");
  auto it = g->w.find(name);
  if (it == g->w.end()) return fail(std::string("unknown weight name: ") + name);
  if ((int64_t)it->second.size() != count) return fail(std::string("size mismatch for ") + name);
  std::memcpy(host_data, it->second.data(), sizeof(float) * (size_t)count);
  return 0;
}

// -------------------------------------------------------- weight packing
static inline __nv_bfloat16 tobf(float v) { return __float2bfloat16_rn(v); }
// cvt.rna.tf32.f32 on the host: round to nearest, ties away, onto 10 mantissa bits
static inline float totf32(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x1000u) & ~0x1FFFu;
  std::memcpy(&v, &u, 4);
  return v;
}

// Packed K-major weight matrix [n_rows][num_kb * kbe]: bf16 (kbe 64) or tf32-rounded fp32 (kbe 32).
template <class Fn>
static int upload_B(const wdg_generator* g, void** dev, int n_rows, int num_kb, Fn value /* (row, kb, j) -> float */) {
  const int kbe = g->kbe();
  const size_t count = (size_t)n_rows * num_kb * kbe;
  std::vector<uint8_t> h(count * g->esz());
  for (int r = 0; r < n_rows; ++r)
    for (int kb = 0; kb < num_kb; ++kb)
      for (int j = 0; j < kbe; ++j) {
        const size_t idx = ((size_t)r * num_kb + kb) * kbe + j;
        const float v = value(r, kb, j);
        if (g->prec == PREC_TF32) reinterpret_cast<float*>(h.data())[idx] = totf32(v);
        else reinterpret_cast<__nv_bfloat16*>(h.data())[idx] = tobf(v);
      }
  if (*dev) cudaFree(*dev);
  *dev = nullptr;
  CK(cudaMalloc(dev, h.size()));
  CK(cudaMemcpy(*dev, h.data(), h.size(), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int wdg_generator_set_precision(wdg_generator* g, int mode) {
  if (!g) return fail("null handle");
  if (mode != WDG_PREC_BF16 && mode != WDG_PREC_TF32) return fail("precision must be WDG_PREC_BF16 (0) or WDG_PREC_TF32 (1)");
  if (mode != g->prec) {
    g->prec = mode;
    g->finalized = false;
    for (auto& pl : g->plans) pl.B = 0;
  }
  return 0;
}
extern "C" int wdg_generator_get_precision(const wdg_generator* g) { return g ? g->prec : -1; }

extern "C" int wdg_generator_finalize(wdg_generator* g) {
  if (!g) return fail("null handle");
  const int F = g->F, C = g->cin + g->cnoise, CP = g->CP;
  const int kbe = g->kbe();          // elements per K-block (one 128-byte row)
  const int cpk = F / kbe;           // K-blocks per 128 channels
  auto W = [&](int i, const char* leaf) -> const std::vector<float>& { return g->w[wname(i, leaf)]; };
  // ---- L0: 8x8 s2 conv on the space-to-depth image X2[n][Y][X][(p,q,c)] (padded pixel (2Y+p, 2X+q), CP channels).
  //      It becomes a 4x4 stride-1 conv over 4*CP channels; two horizontally adjacent s2d pixels are contiguous
  //      (8*CP = 192 elements = 3 (bf16) / 6 (tf32) chunks), so tap (a, w) = rows +a, window starting at pixel +2w, and
  //      K-block = chunk*8 + (a*2 + w).  Window element e: pixel w' = e / (4*CP), parity (p,q), channel c.
  {
    const auto& w = W(0, "layer/w");  // [8][8][C][128]
    if (8 * CP != 192) return fail("8x8 s2 conv kernel expects 8*CP == 192 (CP == 24)");
    if (upload_B(g, &g->B0, 128, 192 / kbe * 8, [&](int n, int kb, int j) {
          const int chunk = kb / 8, tap = kb % 8, a = tap / 2, ww = tap % 2;
          const int e = chunk * kbe + j, wp = e / (4 * CP), rem = e % (4 * CP), pq = rem / CP, c = rem % CP;
          const int ky = 2 * a + pq / 2, kx = 2 * (2 * ww + wp) + pq % 2;
          return c < C ? w[(((size_t)ky * 8 + kx) * C + c) * 128 + n] : 0.f;
        })) return 1;
  }
  // ---- L2: 4x4 s2 conv on the padded res_2.  K within tap row = kx*128 + c.
  {
    const auto& w = W(2, "layer/w");  // [4][4][128][128]
    const int per = 4 * 128 / kbe;    // K-blocks per filter row
    if (upload_B(g, &g->B2, 128, 4 * per, [&](int n, int kb, int j) {
          const int ky = kb / per, e = (kb % per) * kbe + j, kx = e / 128, c = e % 128;
          return w[(((size_t)ky * 4 + kx) * 128 + c) * 128 + n];
        })) return 1;
  }
  // ---- ConvLSTM: packed column = ntile*256 + gate*64 + cc  <->  original gate*F + ntile*64 + cc.
  {
    const auto& k = W(4, "cell/kernel");
    const auto& r = W(4, "cell/recurrent_kernel");
    const int nk = 9 * cpk;           // K-blocks of one operand (x-conv or h-conv)
    if (upload_B(g, &g->BL, 4 * F, 2 * nk, [&](int n, int kb, int j) {
          const int ntile = n / 256, gate = (n % 256) / 64, cc = n % 64;
          const int orig = gate * F + ntile * 64 + cc;
          const int kk = kb % nk, tap = kk / cpk, c = (kk % cpk) * kbe + j;
          const auto& src = kb < nk ? k : r;
          return src[((size_t)tap * F + c) * (4 * F) + orig];
        })) return 1;
  }
  // ---- L5: 3x3 same conv 128 -> 64
  {
    const auto& w = W(5, "layer/w");  // [3][3][128][64]
    if (upload_B(g, &g->B5, F / 2, 9 * cpk, [&](int n, int kb, int j) {
          const int tap = kb / cpk, c = (kb % cpk) * kbe + j;
          return w[((size_t)tap * F + c) * (F / 2) + n];
        })) return 1;
    // halo kernel: K-block = chunk * 9 + tap
    if (upload_B(g, &g->B5h, F / 2, 9 * cpk, [&](int n, int kb, int j) {
          const int tap = kb % 9, c = (kb / 9) * kbe + j;
          return w[((size_t)tap * F + c) * (F / 2) + n];
        })) return 1;
  }
  // ---- L7: ConvT 2x2 s2, kernel (kh,kw,out,in); column = (ky*2+kx)*32 + o; K = in channel (g5 first, res_4 second)
  {
    const auto& w = W(7, "layer/w");  // [2][2][32][192]
    const int O = F / 4, I = F / 2 + F;
    if (upload_B(g, &g->B7, 4 * O, I / kbe, [&](int n, int kb, int j) {
          const int grp = n / O, o = n % O, i = kb * kbe + j;
          return w[((size_t)grp * O + o) * I + i];
        })) return 1;
  }
  // ---- L9: UpSampling2D(2, bilinear) + ConvT 5x5 same s1, fused.  The ConvT is a SAME correlation with
  //      Wf[ty][tx][c][o] = W[4-ty][4-tx][o][c].  GEMM row = anchor (r, s) of a 4x4 low-res window (rows r..r+3),
  //      column = (py, px, o): high-res pixel (2r+3+py, 2s+3+px).  High-res row 2r+1+m (m = py+ty) is the bilinear
  //      blend U[m][dy] of window rows, so  comp[py,px,o][dy,dx,c] = sum_{ty,tx} U[py+ty][dy] U[px+tx][dx] Wf[ty][tx][c][o].
  {
    const auto& w = W(9, "layer/kernel");  // [5][5][16][160]
    const int O = F / 8, I = F / 4 + 128;
    if (O != 16 || I != 160) return fail("fused upsample conv expects 160 -> 16 channels");
    const int nch = (I + kbe - 1) / kbe;   // channel chunks: 3 (bf16; the last one half full) / 5 (tf32)
    static const double U[6][4] = {{.75, .25, 0, 0}, {.25, .75, 0, 0}, {0, .75, .25, 0},
                                   {0, .25, .75, 0}, {0, 0, .75, .25}, {0, 0, .25, .75}};
    auto Wf = [&](int ty, int tx, int c, int o) { return (double)w[(((size_t)(4 - ty) * 5 + (4 - tx)) * O + o) * I + c]; };
    auto comp = [&](int n, int tap, int chunk, int j) {
      const int py = n / (2 * O), px = (n / O) % 2, o = n % O;
      const int dy = tap / 4, dx = tap % 4;
      const int c = chunk * kbe + j;
      if (c >= I) return 0.f;
      double acc = 0;
      for (int ty = 0; ty < 5; ++ty) {
        const double uy = U[py + ty][dy];
        if (uy == 0) continue;
        for (int tx = 0; tx < 5; ++tx) acc += uy * U[px + tx][dx] * Wf(ty, tx, c, o);
      }
      return (float)acc;
    };
    // generic kernel: K-block = tap*nch + chunk;  halo kernel: K-block = chunk*16 + tap
    if (upload_B(g, &g->B9, 4 * O, 16 * nch, [&](int n, int kb, int j) { return comp(n, kb / nch, kb % nch, j); })) return 1;
    if (upload_B(g, &g->B9h, 4 * O, 16 * nch, [&](int n, int kb, int j) { return comp(n, kb % 16, kb / 16, j); })) return 1;
    // Border "dipole" corrections (see stencil_kernels.cuh: edge_lines_kernel): 1-D 5-tap convolutions of the four
    // upsampled edge lines; row = edge*48 + e*16 + o, where e is the distance of the output row/col from that edge.
    if (upload_B(g, &g->BE, 4 * 48, 5 * nch, [&](int n, int kb, int j) {
          const int edge = n / 48, e = (n % 48) / 16, o = n % 16;
          const int t = kb / nch, chunk = kb % nch;
          const int c = chunk * kbe + j;
          if (c >= I) return 0.f;
          double v = 0;
          switch (edge) {
            case 0: v = Wf(2 - e, t, c, o) - (e <= 1 ? Wf(1 - e, t, c, o) : 0.0); break;
            case 1: v = Wf(2 + e, t, c, o) - (e <= 1 ? Wf(3 + e, t, c, o) : 0.0); break;
            case 2: v = Wf(t, 2 - e, c, o) - (e <= 1 ? Wf(t, 1 - e, c, o) : 0.0); break;
            default: v = Wf(t, 2 + e, c, o) - (e <= 1 ? Wf(t, 3 + e, c, o) : 0.0); break;
          }
          return (float)(0.25 * v);
        })) return 1;
    // Final 3x3 conv (16 -> 2) in super-pixel form (bf16 only; the tf32 path keeps this 0.14 % of the MACs in full fp32 on
    // the CUDA cores): one GEMM row = 4 horizontally adjacent pixels x 16 channels (one
    // 128-byte row of the padded g9 image), K-block = super-tap (dy, dsx) with dsx in {-1, 0, +1} super-pixels, GEMM
    // column n = (output pixel po, output channel o) for n < 8; weight of input pixel pi / channel c:
    // w[dy][kx][c][o] with kx = 4 (dsx - 1) + pi - po + 1 when that lies in 0..2, else 0.
    if (g->prec == PREC_BF16) {
      const auto& w11 = W(11, "layer/kernel");   // [3][3][16][2]
      if (F / 8 != 16 || g->cout != 2) return fail("final conv kernel expects 16 -> 2 channels");
      // columns 8..15 carry the bf16 residual of the same weights (w - bf16(w)); the epilogue adds the two halves, so the
      // output layer sees its weights to ~16 bits at no extra MMA (N = 16 is the minimum tile width anyway)
      if (upload_B(g, &g->B11, 16, 9, [&](int n, int kb, int j) {
            const int po = (n % 8) / 2, o = n % 2, dy = kb / 3, dsx = kb % 3, pi = j / 16, c = j % 16;
            const int kx = 4 * (dsx - 1) + pi - po + 1;
            if (kx < 0 || kx > 2) return 0.f;
            const float w = w11[((size_t)(dy * 3 + kx) * 16 + c) * 2 + o];
            return n < 8 ? w : w - __bfloat162float(tobf(w));
          })) return 1;
    }
  }
  // ---- fp32 per-column vectors
  std::vector<float> fp;
  auto push = [&](const std::vector<float>& v) {  // 16-byte aligned so the kernels can use float4 loads
    while (fp.size() % 4) fp.push_back(0.f);
    size_t o = fp.size(); fp.insert(fp.end(), v.begin(), v.end()); return o; };
  auto bn_fold = [&](int i, int c, std::vector<float>& sc, std::vector<float>& sh) {
    const auto &ga = W(i, "gamma"), &be = W(i, "beta"), &mu = W(i, "moving_mean"), &va = W(i, "moving_variance");
    sc.resize(c); sh.resize(c);
    for (int k = 0; k < c; ++k) {
      const double s = (double)ga[k] / std::sqrt((double)va[k] + 1e-3);
      sc[k] = (float)s;
      sh[k] = (float)((double)be[k] - (double)mu[k] * s);
    }
  };
  std::vector<float> sc, sh;
  size_t o_b0 = push(W(0, "layer/layer/bias")); bn_fold(1, 128, sc, sh); size_t o_sc0 = push(sc), o_sh0 = push(sh);
  size_t o_b2 = push(W(2, "layer/layer/bias")); bn_fold(3, F, sc, sh); size_t o_sc2 = push(sc), o_sh2 = push(sh);
  std::vector<float> bl(4 * F);
  {
    const auto& b = W(4, "cell/bias");
    for (int n = 0; n < 4 * F; ++n) {
      const int ntile = n / 256, gate = (n % 256) / 64, cc = n % 64;
      bl[n] = b[gate * F + ntile * 64 + cc];
    }
  }
  size_t o_bl = push(bl);
  size_t o_b5 = push(W(5, "layer/layer/bias")); bn_fold(6, F / 2, sc, sh); size_t o_sc5 = push(sc), o_sh5 = push(sh);
  {
    const auto& b = W(7, "layer/layer/bias");
    bn_fold(8, F / 4, sc, sh);
    std::vector<float> b4, sc4, sh4;
    for (int grp = 0; grp < 4; ++grp) {
      b4.insert(b4.end(), b.begin(), b.end()); sc4.insert(sc4.end(), sc.begin(), sc.end()); sh4.insert(sh4.end(), sh.begin(), sh.end());
    }
    size_t o_b7 = push(b4), o_sc7 = push(sc4), o_sh7 = push(sh4);
    size_t o_b9 = push(W(9, "layer/bias")); bn_fold(10, F / 8, sc, sh); size_t o_sc9 = push(sc), o_sh9 = push(sh);
    size_t o_w11 = push(W(11, "layer/kernel")), o_b11 = push(W(11, "layer/bias"));
    size_t o_zero = push(std::vector<float>(192, 0.f)), o_one = push(std::vector<float>(192, 1.f));
    if (g->fparams) cudaFree(g->fparams);
    g->fparams = nullptr;
    CK(cudaMalloc(&g->fparams, fp.size() * sizeof(float)));
    CK(cudaMemcpy(g->fparams, fp.data(), fp.size() * sizeof(float), cudaMemcpyHostToDevice));
    float* f = g->fparams;
    g->bias0 = f + o_b0; g->sc0 = f + o_sc0; g->sh0 = f + o_sh0;
    g->bias2 = f + o_b2; g->sc2 = f + o_sc2; g->sh2 = f + o_sh2;
    g->biasL = f + o_bl;
    g->bias5 = f + o_b5; g->sc5 = f + o_sc5; g->sh5 = f + o_sh5;
    g->bias7 = f + o_b7; g->sc7 = f + o_sc7; g->sh7 = f + o_sh7;
    g->bias9 = f + o_b9; g->sc9 = f + o_sc9; g->sh9 = f + o_sh9;
    g->w11 = f + o_w11; g->b11 = f + o_b11;
    std::memcpy(g->w11h.w, W(11, "layer/kernel").data(), sizeof g->w11h.w);
    std::memcpy(g->w11h.b, W(11, "layer/bias").data(), sizeof g->w11h.b);
    g->zero48 = f + o_zero; g->one48 = f + o_one;
  }
  (void)C;
  g->finalized = true;
  for (auto& pl : g->plans) pl.B = 0;  // any previous plan referenced old buffers
  return 0;
}

// ------------------------------------------------------------- workspace
struct WsLayout {
  size_t xpad, res4, hseq, cstate, g5, catp, edgeE, deltaD, g9, flags, total;
};
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static WsLayout ws_layout(const wdg_generator* g, int B, int T) {
  const size_t N = (size_t)B * T, S = g->S, S2 = S / 2, S4 = S / 4, F = g->F, E = g->esz();
  WsLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  L.xpad = take(N * (S + 6) * (S + 6) * g->CP * E + 1024);   // s2d image [N][(S+6)/2][(S+6)/2][4*CP] (+ window slack)
  L.res4 = take(N * S4 * S4 * F * E);
  L.hseq = take(N * (S4 + 2) * (S4 + 2) * F * E);   // zero ring 1 (halo form of the 3x3 conv)
  L.cstate = take((size_t)B * S4 * S4 * F * 4);
  L.g5 = take(N * S4 * S4 * (F / 2) * E);
  L.catp = take(N * (S2 + 4) * (S2 + 4) * g->catp_pitch() * E + 8192);
  L.edgeE = take(N * 4 * (S + 8) * (F / 4 + 128) * E);
  L.deltaD = take(N * S * 192 * 4);
  L.g9 = take(N * (S + 2) * (S + 8) * (F / 8) * E);   // zero ring: 1 row above/below, one 4-pixel super-pixel left/right
  L.flags = take((size_t)T * ((B + 1) / 2) * sizeof(unsigned long long));   // (step, image pair) counters of the persistent ConvLSTM launch
  L.total = o;
  return L;
}

// The pipelined host entry point cuts B into pieces; every distinct piece size gets its own plan and workspace region.
// H2D copy of piece i+1 | forward of piece i | D2H copy of piece i-1 run on three streams.
//  * host-supplied noise (434 MB per 512 fields): the pipeline is bound by the H2D copies; what it adds to them is the
//    forward + D2H of the LAST piece, so: uniform chunks of CHUNK_B and a short last piece (measured 8.72 -> 8.4 ms / 64 seq);
//  * noise drawn on the device: compute-bound.  step = H2D(first piece) + sum of forwards + D2H(last piece), and a forward
//    costs ~0.2 ms + 0.027 ms per sequence (profiles/r2_batch_sweep.json), so: a small head (its copy is all that is exposed),
//    pieces that double while the copy engine stays ahead of the SMs (a piece's H2D must finish before its predecessor's
//    forward does: 17.7 us per sequence copied vs >= 27 us computed), large middle pieces, a small tail.
//    B = 64: 8 | 16 | 32 | 8.
// WDG_CHUNK_B=n forces uniform chunks of n for both (tuning / tests).
static const int CHUNK_B = 16, SHORT_B = 4, HEAD_B = 8, BIG_B = 32;
struct Schedules { std::vector<int> host, dev; };
static Schedules make_schedules(int B) {
  Schedules S;
  int cb = CHUNK_B;
  bool forced = false;
  if (const char* e = getenv("WDG_CHUNK_B")) {
    const int v = atoi(e);
    if (v > 0) { cb = v; forced = true; }
  }
  if (B < 2 * cb) return S;                       // small batch: one piece, no pipeline
  for (int b = 0; b < B; b += cb) S.host.push_back(B - b < cb ? B - b : cb);
  if (B % cb == 0 && cb >= 3 * SHORT_B) { S.host.back() = cb - SHORT_B; S.host.push_back(SHORT_B); }
  if (forced) {
    for (int b = 0; b < B; b += cb) S.dev.push_back(B - b < cb ? B - b : cb);
    return S;
  }
  S.dev.push_back(HEAD_B);
  int rem = B - 2 * HEAD_B, next = 2 * HEAD_B;
  while (rem > 0) {
    int take = next < rem ? next : rem;
    if (rem - take < HEAD_B) take = rem;          // no sliver before the tail
    S.dev.push_back(take);
    rem -= take;
    if (next < BIG_B) next *= 2;
  }
  S.dev.push_back(HEAD_B);
  return S;
}
static std::vector<int> distinct_sizes(const Schedules& S) {
  std::vector<int> v;
  for (const std::vector<int>* l : {&S.host, &S.dev})
    for (int n : *l) {
      bool seen = false;
      for (int m : v) seen = seen || m == n;
      if (!seen) v.push_back(n);
    }
  return v;
}

// Introspection (host arithmetic only): the piece sizes predict_host* cuts a batch of B sequences into.
extern "C" int wdg_generator_pipeline_schedule(int B, int host_noise, int* pieces, int capacity) {
  if (B <= 0 || (!pieces && capacity > 0)) return -1;
  const Schedules S = make_schedules(B);
  const std::vector<int>& v = host_noise ? S.host : S.dev;
  for (int i = 0; i < (int)v.size() && i < capacity; ++i) pieces[i] = v[i];
  return (int)v.size();
}

extern "C" int wdg_generator_workspace_bytes(const wdg_generator* g, int B, int T, size_t* bytes) {
  if (!g || !bytes || B <= 0 || T <= 0) return fail("bad argument");
  *bytes = ws_layout(g, B, T).total;
  for (int n : distinct_sizes(make_schedules(B))) *bytes += ws_layout(g, n, T).total;
  return 0;
}

extern "C" int wdg_generator_io_bytes(const wdg_generator* g, int B, int T, size_t* bytes) {
  if (!g || !bytes || B <= 0 || T <= 0) return fail("bad argument");
  int mp = 0;
  for (int n : distinct_sizes(make_schedules(B))) mp = n > mp ? n : mp;
  const int seqs = 2 * mp > B ? 2 * mp : B;        // double-buffered staging of the largest piece, or the whole batch
  const size_t px = (size_t)seqs * T * g->S * g->S;
  *bytes = align_up(px * g->cin * 4, 256) + align_up(px * g->cnoise * 4, 256) + align_up(px * g->cout * 4, 256) + 1024;
  return 0;
}

static void set_tiles(ConvParams& p, int H, int W, int N, int tw, int th, int tn, int n_tiles_N, int n_coord) {
  p.H = H; p.W = W; p.N = N;
  p.tile_w = tw; p.tile_h = th; p.tile_n = tn;
  p.tiles_x = (W + tw - 1) / tw; p.tiles_y = (H + th - 1) / th; p.tiles_n = (N + tn - 1) / tn;
  p.n_tiles_N = n_tiles_N;
  p.n_coord = n_coord;
  p.ntile_coord = -1;
}
static void affine_epi(EpiParams& e, const float* bias, const float* sc, const float* sh, void* out, long long sn,
                       long long sy, long long sx, int c0, int lrelu) {
  std::memset(&e, 0, sizeof e);
  e.bias = bias; e.scale = sc; e.shift = sh; e.out = out;
  e.out_sn = sn; e.out_sy = sy; e.out_sx = sx; e.out_c0 = c0;
  e.out_mul = 1; e.group_cols = 1 << 30; e.lrelu = lrelu; e.out_f32 = 0;
}

// ---- CTA-pair ConvLSTM launch (conv_umma2.cuh): clusters of 2, cooperative (the step flags need every CTA resident)
template <int PREC>
static int lstm_pair_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attrs, int grid, cudaStream_t stream) {
  auto kern = conv_pair_kernel<256, EPI_LSTM, PREC>;
  static bool done[64] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<256>::SMEM_BYTES));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  std::memset(cfg, 0, sizeof *cfg);
  cfg->gridDim = dim3(grid);
  cfg->blockDim = dim3(192);
  cfg->dynamicSmemBytes = PairCfg<256>::SMEM_BYTES;
  cfg->stream = stream;
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg->attrs = attrs;
  // WDG_LSTM_COOP=0 (profiling only): Nsight Compute refuses the cooperative + cluster launch (LaunchFailed); without the
  // attribute the grid (<= one CTA per SM) is still fully resident on an otherwise idle GPU, which is what ncu's serialised
  // replay provides.  The bounded flag wait traps instead of hanging if that ever does not hold.
  static const int coop = getenv("WDG_LSTM_COOP") ? atoi(getenv("WDG_LSTM_COOP")) : 1;
  cfg->numAttrs = coop ? 2 : 1;
  return 0;
}
static int lstm_pair_max_clusters(const wdg_generator* g, int* n) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  *n = 0;
  if (g->prec == PREC_TF32) {
    if (lstm_pair_config<PREC_TF32>(&cfg, attrs, 2 * g->sm_count, nullptr)) return 1;
    cfg.numAttrs = 1;       // the occupancy query takes the cluster shape only
    CK(cudaOccupancyMaxActiveClusters(n, conv_pair_kernel<256, EPI_LSTM, PREC_TF32>, &cfg));
  } else {
    if (lstm_pair_config<PREC_BF16>(&cfg, attrs, 2 * g->sm_count, nullptr)) return 1;
    cfg.numAttrs = 1;
    CK(cudaOccupancyMaxActiveClusters(n, conv_pair_kernel<256, EPI_LSTM, PREC_BF16>, &cfg));
  }
  return 0;
}
template <int PREC>
static int launch_lstm_pair(const Plan& pl, cudaStream_t stream);

static int build_plan(wdg_generator* g, Plan& pl, int B, int T, uint8_t* ws) {
  const WsLayout L = ws_layout(g, B, T);
  pl.xpad = ws + L.xpad;
  pl.res4 = ws + L.res4; pl.hseq = ws + L.hseq;
  pl.cstate = (float*)(ws + L.cstate); pl.g5 = ws + L.g5; pl.catp = ws + L.catp;
  pl.edgeE = ws + L.edgeE; pl.deltaD = (float*)(ws + L.deltaD); pl.g9 = ws + L.g9;
  pl.lstm_flags = (unsigned long long*)(ws + L.flags);
  const uint64_t N = (uint64_t)B * T, S = g->S, S2 = S / 2, S4 = S / 4, F = g->F, CP = g->CP;
  const int esz = g->esz();
  const uint32_t kbe = (uint32_t)g->kbe();       // elements per K-block
  const int cpk = (int)(F / kbe);                // K-blocks per 128 channels
  const uint64_t CI = g->catp_pitch();           // channel pitch of the concat image
  const int sms = g->sm_count;
  auto at = [&](uint8_t* base, long long elems) { return base + elems * esz; };   // element offset into an act_t buffer
  auto tmap = [&](CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* str, const uint32_t* box,
                  int inner_bytes = 128) { return wdg_make_tmap(tm, base, rank, dims, str, box, inner_bytes, esz); };
  auto grid_for = [&](const ConvParams& p, int ctas_per_sm = 1) {
    const int total = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_N;
    return total < sms * ctas_per_sm ? total : sms * ctas_per_sm;
  };
  static const int shallow_mask = getenv("WDG_SHALLOW") ? atoi(getenv("WDG_SHALLOW")) : 7;   // bit 0: convT 2x2, 1: border GEMM, 2: conv 4x4
  // ---------------- L0: 8x8 s2 as 4x4 s1 on the s2d image X2 [N][Q][Q][4*CP], Q = (S+6)/2; dims (window 8*CP, X, Y, n)
  {
    ConvLaunch& c = pl.L0;
    std::memset(&c.p, 0, sizeof c.p);
    const uint64_t Q = (S + 6) / 2, PC = 4 * CP;
    const int nch = (int)(2 * PC / kbe);          // chunks of the 2-pixel window: 3 (bf16) / 6 (tf32)
    uint64_t dims[5] = {2 * PC, Q, Q, N, 1};
    uint64_t str[4] = {PC, Q * PC, Q * Q * PC, N * Q * Q * PC};
    uint32_t box[5] = {kbe, 16, 8, 1, 1};
    if (tmap(&c.tmA[0], pl.xpad, 5, dims, str, box)) return 1;
    c.tmA[1] = c.tmA[0]; c.tmA[2] = c.tmA[0];
    uint64_t bd[2] = {(uint64_t)nch * 8 * kbe, 128};
    uint64_t bs[1] = {(uint64_t)nch * 8 * kbe};
    uint32_t bb[2] = {kbe, 128};
    if (tmap(&c.tmB, g->B0, 2, bd, bs, bb)) return 1;
    set_tiles(c.p, (int)S2, (int)S2, (int)N, 16, 8, 1, 1, 3);
    c.p.num_kb = nch * 8;
    for (int ch = 0; ch < nch; ++ch)
      for (int tap = 0; tap < 8; ++tap) {
        KBlock& k = c.p.kb[ch * 8 + tap];
        k.src = 0; k.half = 0; k.o0 = (int16_t)(ch * kbe); k.o1 = (int16_t)(2 * (tap % 2)); k.o2 = (int16_t)(tap / 2); k.o3 = 0;
      }
    // res_2 is written once, as channels 32..159 of the zero-padded concat image `catp` (ring 2): it is read there
    // by the 4x4 s2 conv (through an overlapping-stride window map) and by the fused upsample conv.
    const long long PW = S2 + 4, ci = (long long)CI;
    affine_epi(c.p.ep, g->bias0, g->sc0, g->sh0, at(pl.catp, (2 * PW + 2) * ci), PW * PW * ci, PW * ci, ci, (int)(F / 4), 1);
    c.bn = 128; c.epi = EPI_AFFINE; c.grid = grid_for(c.p);
    // halo-reuse variant: flat positions of X2 (pitch Q), taps shift by a*Q + 2w rows
    const bool fits = (3 * Q + 2 + H_TILES * TILE_M) <= (uint64_t)H_ROWS;
    if (fits) {
      const uint64_t flat = N * Q * Q;
      uint64_t hd[2] = {2 * PC, flat};
      uint64_t hs[1] = {PC};
      uint32_t hb[2] = {kbe, H_BOX_ROWS};
      if (tmap(&pl.h0A, pl.xpad, 2, hd, hs, hb)) return 1;
      pl.h0B = c.tmB;
      HaloParams& h = pl.h0p;
      std::memset(&h, 0, sizeof h);
      h.num_passes = (int)((flat + H_TILES * TILE_M - 1) / (H_TILES * TILE_M));
      h.n_img = (int)N; h.pw = (int)Q; h.ph = (int)Q; h.box_rows = H_BOX_ROWS;
      for (int tap = 0; tap < 8; ++tap) { h.tap_shift[tap] = (tap / 2) * (int)Q + 2 * (tap % 2); h.kmask[tap] = 0xF; }
      h.bias = g->bias0; h.scale = g->sc0; h.shift = g->sh0;
      h.vw = (int)S2; h.vh = (int)S2;
      h.out1 = at(pl.catp, (2 * PW + 2) * ci + F / 4); h.o1_sn = PW * PW * ci; h.o1_sy = PW * ci; h.o1_sx = ci;
      h.out2 = nullptr;
      pl.h0grid = h.num_passes < sms ? h.num_passes : sms;
      uint32_t bbh[2] = {kbe, 64};
      if (tmap(&pl.h0B_half, g->B0, 2, bd, bs, bbh)) return 1;
      const int pairs = (h.num_passes + 1) / 2;
      pl.h0grid_pair = 2 * (pairs < sms / 2 ? pairs : sms / 2);
    }
    pl.use_halo = fits;
  }
  // ---------------- L2: ZeroPadding2D(1) + 4x4 s2 on res_2 = channels 32..159 of catp [N][S2+4][S2+4][CI] (ring 2,
  //                  so the pad-1 image starts at (1,1)); dims (window of 4 pixels, ox, oy, row parity, n)
  {
    ConvLaunch& c = pl.L2;
    std::memset(&c.p, 0, sizeof c.p);
    const uint64_t PW = S2 + 4;
    const int per = (int)(4 * 128 / kbe);         // K-blocks per filter row
    uint64_t dims[5] = {3 * CI + 128, S4, PW / 2 - 1, 2, N};
    uint64_t str[4] = {2 * CI, 2 * PW * CI, PW * CI, PW * PW * CI};
    uint32_t box[5] = {kbe, 8, 8, 1, 2};
    if (tmap(&c.tmA[0], at(pl.catp, (PW + 1) * CI + F / 4), 5, dims, str, box)) return 1;
    c.tmA[1] = c.tmA[0]; c.tmA[2] = c.tmA[0];
    uint64_t bd[2] = {(uint64_t)4 * per * kbe, 128};
    uint64_t bs[1] = {(uint64_t)4 * per * kbe};
    uint32_t bb[2] = {kbe, 128};
    if (tmap(&c.tmB, g->B2, 2, bd, bs, bb)) return 1;
    set_tiles(c.p, (int)S4, (int)S4, (int)N, 8, 8, 2, 1, 4);
    c.p.num_kb = 4 * per;
    for (int ky = 0; ky < 4; ++ky)
      for (int q = 0; q < per; ++q) {   // element q*kbe of the 4-pixel x 128-channel filter row: pixel kx, channel c0
        KBlock& k = c.p.kb[ky * per + q];
        const int e0 = q * (int)kbe, kx = e0 / 128, c0 = e0 % 128;
        k.src = 0; k.half = 0; k.o0 = (int16_t)(kx * CI + c0); k.o1 = 0; k.o2 = (int16_t)(ky / 2); k.o3 = (int16_t)(ky % 2);
      }
    affine_epi(c.p.ep, g->bias2, g->sc2, g->sh2, pl.res4, (long long)S4 * S4 * F, (long long)S4 * F, F, 0, 1);
    c.bn = 128; c.epi = EPI_AFFINE; c.shallow = (shallow_mask >> 2) & 1; c.grid = grid_for(c.p, c.shallow ? 2 : 1);
    static const int l2_pair_env = getenv("WDG_CONV_PAIR") ? atoi(getenv("WDG_CONV_PAIR")) : 1;
    uint32_t bbh[2] = {kbe, 64};
    if (tmap(&pl.l2B_half, g->B2, 2, bd, bs, bbh)) return 1;
    const int m_tiles = c.p.tiles_x * c.p.tiles_y * c.p.tiles_n, pair_tiles = ((m_tiles + 1) / 2) * c.p.n_tiles_N;
    pl.l2_pair_grid = 2 * (pair_tiles < sms / 2 ? pair_tiles : sms / 2);
    pl.l2_pair = l2_pair_env != 0 && pl.l2_pair_grid > 0;
  }
  // ---------------- ConvLSTM steps: A maps over (c, x, y, t, b) of res4 (x_t) and hseq (h_{t-1})
  {
    CUtensorMap tmX, tmH, tmB;
    const int nk = 9 * cpk;                       // K-blocks of one operand
    uint64_t dims[5] = {F, S4, S4, (uint64_t)T, (uint64_t)B};
    uint64_t str[4] = {F, S4 * F, S4 * S4 * F, (uint64_t)T * S4 * S4 * F};
    uint32_t box[5] = {kbe, 8, 8, 1, 2};
    if (tmap(&tmX, pl.res4, 5, dims, str, box)) return 1;
    // hseq images carry a zero ring of 1 pixel: same logical dims, padded strides, base at the interior origin
    const uint64_t SP = S4 + 2;
    uint64_t strh[4] = {F, SP * F, SP * SP * F, (uint64_t)T * SP * SP * F};
    if (tmap(&tmH, at(pl.hseq, (SP + 1) * F), 5, dims, strh, box)) return 1;
    uint64_t bd[2] = {(uint64_t)2 * nk * kbe, 4 * F};
    uint64_t bs[1] = {(uint64_t)2 * nk * kbe};
    uint32_t bb[2] = {kbe, 256};
    if (tmap(&tmB, g->BL, 2, bd, bs, bb)) return 1;
    if (2 * nk > MAX_KB) return fail("ConvLSTM K-blocks exceed MAX_KB");
    pl.LS.assign(T, ConvLaunch());
    for (int t = 0; t < T; ++t) {
      ConvLaunch& c = pl.LS[t];
      std::memset(&c.p, 0, sizeof c.p);
      c.tmA[0] = tmX; c.tmA[1] = tmH; c.tmA[2] = tmX; c.tmB = tmB;
      set_tiles(c.p, (int)S4, (int)S4, B, 8, 8, 2, (int)(4 * F / 256), 4);
      // K axis: input half (x_t: time offset 0) then recurrent half (h_{t-1}: time offset -1); step 0 runs the first only
      c.p.num_kb = 2 * nk; c.p.num_kb_first = nk;
      c.p.t_begin = t; c.p.t_end = t + 1;
      for (int kb = 0; kb < c.p.num_kb; ++kb) {
        KBlock& k = c.p.kb[kb];
        const int kk = kb % nk, tap = kk / cpk;
        k.src = kb < nk ? 0 : 1; k.half = 0;
        k.o0 = (int16_t)((kk % cpk) * kbe); k.o1 = (int16_t)(tap % 3 - 1); k.o2 = (int16_t)(tap / 3 - 1);
        k.o3 = (int16_t)(kb < nk ? 0 : -1);
      }
      EpiParams& e = c.p.ep;
      std::memset(&e, 0, sizeof e);
      e.bias = g->biasL; e.c_state = pl.cstate; e.h_out = at(pl.hseq, (SP + 1) * F);
      e.h_sn = (long long)T * SP * SP * F; e.h_off = 0; e.h_step = (long long)SP * SP * F; e.h_pitch = (int)SP;
      e.F = (int)F;
      c.bn = 256; c.epi = EPI_LSTM; c.grid = grid_for(c.p);
    }
    // All T steps in one cooperative launch: no launch / prologue / drain between steps, and the input half of step
    // t+1 overlaps the tail of step t (north_star: the recurrence stays inside one resident kernel).
    pl.LSP = pl.LS[0];
    pl.LSP.p.t_begin = 0; pl.LSP.p.t_end = T;
    pl.LSP.p.ep.sync_flags = pl.lstm_flags;
    pl.LSP.p.ep.sync_total = 4u * (unsigned)(pl.LSP.p.tiles_x * pl.LSP.p.tiles_y * pl.LSP.p.n_tiles_N);   // per image group: 4 epilogue warps per tile
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, g->device);
    pl.lstm_persist = T > 1 && coop && pl.LSP.grid <= sms && !getenv("WDG_NO_LSTM_PERSIST");
    // CTA-pair form of the same launch (WDG_LSTM_PAIR=0 keeps the single-CTA kernel)
    static const int pair_env = getenv("WDG_LSTM_PAIR") ? atoi(getenv("WDG_LSTM_PAIR")) : 1;
    pl.lstm_pair = false;
    if (pl.lstm_persist && pair_env) {
      uint32_t bb2[2] = {kbe, 128};
      if (tmap(&pl.tmB_half, g->BL, 2, bd, bs, bb2)) return 1;
      pl.pair_p = pl.LSP.p;
      const int m_tiles = pl.pair_p.tiles_x * pl.pair_p.tiles_y * pl.pair_p.tiles_n;
      const int pair_tiles = ((m_tiles + 1) / 2) * pl.pair_p.n_tiles_N;
      int max_clusters = 0;
      if (lstm_pair_max_clusters(g, &max_clusters)) return 1;
      const int pairs = pair_tiles < max_clusters ? pair_tiles : max_clusters;
      pl.pair_grid = 2 * pairs;
      pl.lstm_pair = pairs > 0;
    }
  }
  // ---------------- L5: 3x3 same 128 -> 64 on hseq (c, x, y, n)
  {
    ConvLaunch& c = pl.L5;
    std::memset(&c.p, 0, sizeof c.p);
    const uint64_t SP = S4 + 2;
    uint64_t dims[5] = {F, S4, S4, N, 1};
    uint64_t str[4] = {F, SP * F, SP * SP * F, N * SP * SP * F};
    uint32_t box[5] = {kbe, 8, 8, 2, 1};
    if (tmap(&c.tmA[0], at(pl.hseq, (SP + 1) * F), 5, dims, str, box)) return 1;
    c.tmA[1] = c.tmA[0]; c.tmA[2] = c.tmA[0];
    uint64_t bd[2] = {(uint64_t)9 * cpk * kbe, F / 2};
    uint64_t bs[1] = {(uint64_t)9 * cpk * kbe};
    uint32_t bb[2] = {kbe, 64};
    if (tmap(&c.tmB, g->B5, 2, bd, bs, bb)) return 1;
    set_tiles(c.p, (int)S4, (int)S4, (int)N, 8, 8, 2, 1, 3);
    c.p.num_kb = 9 * cpk;
    for (int kb = 0; kb < 9 * cpk; ++kb) {
      KBlock& k = c.p.kb[kb];
      const int tap = kb / cpk;
      k.src = 0; k.half = 0; k.o0 = (int16_t)((kb % cpk) * kbe); k.o1 = (int16_t)(tap % 3 - 1); k.o2 = (int16_t)(tap / 3 - 1); k.o3 = 0;
    }
    affine_epi(c.p.ep, g->bias5, g->sc5, g->sh5, pl.g5, (long long)S4 * S4 * (F / 2), (long long)S4 * (F / 2), F / 2, 0, 1);
    c.bn = 64; c.epi = EPI_AFFINE; c.grid = grid_for(c.p);
    // halo-reuse variant: flat positions of the padded hseq images (pitch SP), tap (dy, dx) = rows + dy*SP + dx
    pl.use_halo5 = F == 128 && (2 * SP + 2 + H_TILES * TILE_M) <= (uint64_t)H_ROWS && !getenv("WDG_NO_HALO5");
    if (pl.use_halo5) {
      const uint64_t rows = N * SP * SP;
      const uint32_t box5 = (uint32_t)(((H_TILES * TILE_M + 2 * SP + 2 + 1) / 2 + 7) / 8 * 8);
      uint64_t ad[2] = {F, rows};
      uint64_t as_[1] = {F};
      uint32_t ab[2] = {kbe, box5};
      if (tmap(&pl.h5A, pl.hseq, 2, ad, as_, ab)) return 1;
      if (tmap(&pl.h5B, g->B5h, 2, bd, bs, bb)) return 1;
      HaloParams& h = pl.h5p;
      std::memset(&h, 0, sizeof h);
      h.num_passes = (int)((rows + H_TILES * TILE_M - 1) / (H_TILES * TILE_M));
      h.n_img = (int)N; h.pw = (int)SP; h.ph = (int)SP; h.box_rows = (int)box5;
      for (int tap = 0; tap < 9; ++tap) { h.tap_shift[tap] = (tap / 3) * (int)SP + tap % 3; h.kmask[tap] = 0xF; }
      h.bias = g->bias5; h.scale = g->sc5; h.shift = g->sh5;
      h.vw = (int)S4; h.vh = (int)S4;
      h.out1 = pl.g5; h.o1_sn = (long long)S4 * S4 * (F / 2); h.o1_sy = (long long)S4 * (F / 2); h.o1_sx = F / 2;
      h.out2 = nullptr;
      pl.h5grid = h.num_passes < sms ? h.num_passes : sms;
      uint32_t bbh[2] = {kbe, 32};
      if (tmap(&pl.h5B_half, g->B5h, 2, bd, bs, bbh)) return 1;
      const int pairs5 = (h.num_passes + 1) / 2;
      pl.h5grid_pair = 2 * (pairs5 < sms / 2 ? pairs5 : sms / 2);
    }
  }
  // ---------------- L7: ConvT 2x2 s2 on concat(g5, res4) -> g7 [N][S2][S2][32] (pixel shuffle)
  {
    ConvLaunch& c = pl.L7;
    std::memset(&c.p, 0, sizeof c.p);
    uint64_t d0[5] = {F / 2, S4, S4, N, 1};
    uint64_t s0[4] = {F / 2, S4 * (F / 2), S4 * S4 * (F / 2), N * S4 * S4 * (F / 2)};
    uint64_t d1[5] = {F, S4, S4, N, 1};
    uint64_t s1[4] = {F, S4 * F, S4 * S4 * F, N * S4 * S4 * F};
    uint32_t box[5] = {kbe, 8, 8, 2, 1};
    if (tmap(&c.tmA[0], pl.g5, 5, d0, s0, box)) return 1;
    if (tmap(&c.tmA[1], pl.res4, 5, d1, s1, box)) return 1;
    c.tmA[2] = c.tmA[0];
    const int n5 = (int)(F / 2 / kbe), n4 = (int)(F / kbe);   // K-blocks from g5 / res_4
    uint64_t bd[2] = {(uint64_t)(n5 + n4) * kbe, 128};
    uint64_t bs[1] = {(uint64_t)(n5 + n4) * kbe};
    uint32_t bb[2] = {kbe, 128};
    if (tmap(&c.tmB, g->B7, 2, bd, bs, bb)) return 1;
    set_tiles(c.p, (int)S4, (int)S4, (int)N, 8, 8, 2, 1, 3);
    c.p.num_kb = n5 + n4;
    for (int kb = 0; kb < n5 + n4; ++kb) {
      KBlock& k = c.p.kb[kb];
      k.src = kb < n5 ? 0 : 1; k.half = 0; k.o0 = (int16_t)((kb < n5 ? kb : kb - n5) * kbe); k.o1 = 0; k.o2 = 0; k.o3 = 0;
    }
    const long long O = F / 4, ci = (long long)CI, PW = S2 + 4;
    affine_epi(c.p.ep, g->bias7, g->sc7, g->sh7, at(pl.catp, (2 * PW + 2) * ci), PW * PW * ci, PW * ci, ci, 0, 1);
    c.p.ep.out_mul = 2; c.p.ep.group_cols = (int)O;
    c.bn = 128; c.epi = EPI_AFFINE; c.shallow = shallow_mask & 1; c.grid = grid_for(c.p, c.shallow ? 2 : 1);
  }
  const uint64_t I9 = F / 4 + 128;                          // channels of concat(g7, res_2)
  const int nch9 = (int)((I9 + kbe - 1) / kbe);             // 3 (bf16, the last chunk half full) / 5 (tf32)
  const bool half9 = I9 % kbe != 0;
  // ---------------- LE: border corrections, 1-D 5-tap conv over the 4 edge lines E[N][4][S+8][160] -> D[N][S][4*48] fp32
  {
    ConvLaunch& c = pl.LE;
    std::memset(&c.p, 0, sizeof c.p);
    const uint64_t I = I9, P = S + 8;
    uint64_t dims[5] = {I, P, 4, N, 1};
    uint64_t str[4] = {I, P * I, 4 * P * I, N * 4 * P * I};
    uint32_t box[5] = {kbe, 16, 1, 8, 1};
    uint32_t boxh[5] = {kbe / 2, 16, 1, 8, 1};
    if (tmap(&c.tmA[0], pl.edgeE, 5, dims, str, box)) return 1;
    if (tmap(&c.tmA[1], pl.edgeE, 5, dims, str, boxh, 64)) return 1;
    c.tmA[2] = c.tmA[0];
    uint64_t bd[2] = {(uint64_t)5 * nch9 * kbe, 192};
    uint64_t bs[1] = {(uint64_t)5 * nch9 * kbe};
    uint32_t bb[2] = {kbe, 48};
    if (tmap(&c.tmB, g->BE, 2, bd, bs, bb)) return 1;
    set_tiles(c.p, 1, (int)S, (int)N, 16, 1, 8, 4, 3);
    c.p.ntile_coord = 2;
    c.p.num_kb = 5 * nch9;
    for (int t = 0; t < 5; ++t)
      for (int ch = 0; ch < nch9; ++ch) {
        KBlock& k = c.p.kb[t * nch9 + ch];
        const bool hf = half9 && ch == nch9 - 1;
        k.src = hf ? 1 : 0; k.half = hf; k.o0 = (int16_t)(ch * kbe); k.o1 = (int16_t)(t + 2); k.o2 = 0; k.o3 = 0;
      }
    affine_epi(c.p.ep, g->zero48, g->one48, g->zero48, pl.deltaD, (long long)S * 192, 0, 192, 0, 0);
    c.p.ep.out_f32 = 1;
    c.bn = 48; c.epi = EPI_AFFINE; c.shallow = (shallow_mask >> 1) & 1; c.grid = grid_for(c.p, c.shallow ? 2 : 1);
  }
  // ---------------- L9: fused bilinear x2 + ConvT 5x5 on the flattened zero-padded concat image catp
  {
    ConvLaunch& c = pl.L9;
    std::memset(&c.p, 0, sizeof c.p);
    const uint64_t I = I9, PW = S2 + 4, flat = N * PW * PW;
    uint64_t dims[5] = {I, flat, 1, 1, 1};
    uint64_t str[4] = {CI, flat * CI, flat * CI, flat * CI};
    uint32_t box[5] = {kbe, 128, 1, 1, 1};
    uint32_t boxh[5] = {kbe / 2, 128, 1, 1, 1};
    if (tmap(&c.tmA[0], pl.catp, 5, dims, str, box)) return 1;
    if (tmap(&c.tmA[1], pl.catp, 5, dims, str, boxh, 64)) return 1;
    c.tmA[2] = c.tmA[0];
    uint64_t bd[2] = {(uint64_t)16 * nch9 * kbe, 64};
    uint64_t bs[1] = {(uint64_t)16 * nch9 * kbe};
    uint32_t bb[2] = {kbe, 64};
    if (tmap(&c.tmB, g->B9, 2, bd, bs, bb)) return 1;
    set_tiles(c.p, 1, (int)flat, 1, 128, 1, 1, 1, 3);
    c.p.N = (int)N;  // images, for the epilogue's row decode
    c.p.num_kb = 16 * nch9;
    if (c.p.num_kb > MAX_KB) return fail("upsample conv K-blocks exceed MAX_KB");
    for (int tap = 0; tap < 16; ++tap)
      for (int ch = 0; ch < nch9; ++ch) {
        KBlock& k = c.p.kb[tap * nch9 + ch];
        const bool hf = half9 && ch == nch9 - 1;
        k.src = hf ? 1 : 0; k.half = hf; k.o0 = (int16_t)(ch * kbe);
        k.o1 = (int16_t)((tap / 4) * PW + (tap % 4)); k.o2 = 0; k.o3 = 0;
      }
    EpiParams& e = c.p.ep;
    std::memset(&e, 0, sizeof e);
    const long long G9C = F / 8, G9X = S + 8, G9Y = S + 2;      // padded g9 image [G9Y][G9X][G9C], interior at (1, 4)
    e.bias = g->bias9; e.scale = g->sc9; e.shift = g->sh9; e.out = at(pl.g9, (G9X + 4) * G9C);
    e.out_sn = G9Y * G9X * G9C; e.out_sy = G9X * G9C;
    e.up_pw = (int)PW; e.up_ph = (int)PW; e.up_S = (int)S; e.up_delta = pl.deltaD;
    c.bn = 64; c.epi = EPI_UPCONV; c.grid = grid_for(c.p);
    // halo-reuse variant (halo_conv.cuh): needs 256 + 3*PW + 3 <= 416 rows of shared memory
    pl.use_halo = pl.use_halo && (3 * PW + 3 + H_TILES * TILE_M) <= (uint64_t)H_ROWS && !getenv("WDG_NO_HALO");
    if (pl.use_halo) {
      uint64_t hd[2] = {CI, flat};
      uint64_t hs[1] = {CI};
      uint32_t hb[2] = {kbe, H_BOX_ROWS};
      if (tmap(&pl.hA, pl.catp, 2, hd, hs, hb)) return 1;
      if (tmap(&pl.hB, g->B9h, 2, bd, bs, bb)) return 1;
      HaloParams& h = pl.hp;
      std::memset(&h, 0, sizeof h);
      h.num_passes = (int)((flat + H_TILES * TILE_M - 1) / (H_TILES * TILE_M));
      h.n_img = (int)N; h.pw = (int)PW; h.ph = (int)PW; h.S = (int)S; h.delta = pl.deltaD; h.box_rows = H_BOX_ROWS;
      for (int tap = 0; tap < 16; ++tap) { h.tap_shift[tap] = (tap / 4) * (int)PW + tap % 4; h.kmask[tap] = 0xF; }
      // 160 input channels in chunks of kbe: the last bf16 chunk (channels 128..191) is half padding -- zero weights and
      // zero activations -- so its upper two 16-channel K slices are never issued (1/6 of the layer's MMAs; exact)
      {
        const int used = (int)(I9 - (uint64_t)(nch9 - 1) * kbe), per_slice = (int)kbe / 4, slices = (used + per_slice - 1) / per_slice;
        h.kskip_tail = (unsigned char)(0xF & ~((1u << slices) - 1u));
      }
      h.bias = g->bias9; h.scale = g->sc9; h.shift = g->sh9; h.out = at(pl.g9, (G9X + 4) * G9C);
      h.up_sn = G9Y * G9X * G9C; h.up_sy = G9X * G9C;
      pl.hgrid = h.num_passes < sms ? h.num_passes : sms;
      uint32_t bbh[2] = {kbe, 32};
      if (tmap(&pl.hB_half, g->B9h, 2, bd, bs, bbh)) return 1;
      const int pairs = (h.num_passes + 1) / 2;
      pl.hgrid_pair = 2 * (pairs < sms / 2 ? pairs : sms / 2);
      static const int halo_pair_env = getenv("WDG_HALO_PAIR") ? atoi(getenv("WDG_HALO_PAIR")) : 1;
      pl.halo_pair = halo_pair_env != 0;
    }
    // ---------------- L11: final 3x3 conv on the tensor cores, super-pixel form (halo_conv.cuh, HEPI_FINAL; bf16 only)
    const long long SPW = G9X / 4;                              // super-pixels per padded row
    pl.use_halo11 = g->prec == PREC_BF16 && G9C == 16 && (2 * SPW + 2 + H_TILES * TILE_M) <= (long long)H_ROWS && !getenv("WDG_NO_HALO11");
    if (pl.use_halo11) {
      const uint64_t rows = N * G9Y * SPW;
      uint64_t ad[2] = {64, rows};
      uint64_t as_[1] = {64};
      // only 256 + 2 SPW + 2 rows are needed per pass: two boxes of half that (rounded up to 8 rows) instead of 2 x 208
      const uint32_t box11 = (uint32_t)(((H_TILES * TILE_M + 2 * SPW + 2 + 1) / 2 + 7) / 8 * 8);
      uint32_t ab[2] = {64, box11};
      if (tmap(&pl.h11A, pl.g9, 2, ad, as_, ab)) return 1;
      uint64_t b11d[2] = {9 * 64, 16};
      uint64_t b11s[1] = {9 * 64};
      uint32_t b11b[2] = {64, 16};
      if (tmap(&pl.h11B, g->B11, 2, b11d, b11s, b11b)) return 1;
      HaloParams& h = pl.h11p;
      std::memset(&h, 0, sizeof h);
      h.num_passes = (int)((rows + H_TILES * TILE_M - 1) / (H_TILES * TILE_M));
      h.n_img = (int)N; h.pw = (int)SPW; h.ph = (int)G9Y; h.S = (int)S; h.box_rows = (int)box11;
      // K slice k of a super-tap = input pixel k of the super-pixel: the left neighbour only contributes its last pixel,
      // the right neighbour its first one (all other weights are structural zeros: their MMAs are skipped)
      for (int tap = 0; tap < 9; ++tap) {
        h.tap_shift[tap] = (tap / 3) * (int)SPW + tap % 3;
        h.kmask[tap] = tap % 3 == 0 ? 0x8 : (tap % 3 == 1 ? 0xF : 0x1);
      }
      h.bias = g->b11;
      pl.h11grid = h.num_passes < sms ? h.num_passes : sms;
    }
  }
  pl.B = B; pl.T = T;
  pl.launches = 1 + 2 + (pl.lstm_persist ? 1 : T) + 2 + 2 + 1 + 1;   // kernels (the step-counter memset of the persistent ConvLSTM is not one)
  return 0;
}


extern "C" int wdg_generator_bind(wdg_generator* g, int B, int T, void* workspace_dev, size_t bytes, void* stream_) {
  if (!g || !workspace_dev || B <= 0 || T <= 0) return fail("bad argument");
  if (!g->finalized) return fail("wdg_generator_finalize must be called before bind");
  cudaStream_t stream = (cudaStream_t)stream_;
  size_t need = 0;
  if (wdg_generator_workspace_bytes(g, B, T, &need)) return 1;
  if (bytes < need) return fail("workspace too small");
  if ((uintptr_t)workspace_dev % 1024) return fail("workspace must be 1024-byte aligned");
  CK(cudaMemsetAsync(workspace_dev, 0, need, stream));
  uint8_t* ws = (uint8_t*)workspace_dev;
  for (auto& pl : g->plans) pl.B = 0;
  g->piece_plans.clear();
  const Schedules S = make_schedules(B);
  g->sched_host = S.host; g->sched_dev = S.dev; g->max_piece = 0;
  if (build_plan(g, g->plans[0], B, T, ws)) return 1;
  ws += ws_layout(g, B, T).total;
  for (int n : distinct_sizes(S)) {
    if (build_plan(g, g->piece_plans[n], n, T, ws)) return 1;
    ws += ws_layout(g, n, T).total;
    g->max_piece = n > g->max_piece ? n : g->max_piece;
  }
  return 0;
}

// ---------------------------------------------------------------- launch
static bool fused_noise_ok(const wdg_generator* g) { return g->cin == 3 && g->cnoise == 20 && g->CP == 24; }

// One launch path for the tcgen05 kernels of the forward: optional cluster of 2 (CTA-pair kernels), cooperative launch
// (persistent ConvLSTM) and programmatic dependent launch (the kernel's prologue overlaps its predecessor's tail; every
// kernel calls griddepcontrol.wait before it touches dependent data).  WDG_PDL=0 turns the last off.  If the driver rejects
// an attribute combination, PDL is dropped for the rest of the process and the launch is retried.
static int g_pdl = -1;
template <typename Kern, typename... Args>
static int launch_ex(Kern kern, int grid, int block, size_t smem, cudaStream_t stream, bool cluster2, bool coop, Args... args) {
  if (g_pdl < 0) g_pdl = getenv("WDG_PDL") ? atoi(getenv("WDG_PDL")) : 1;
  for (int attempt = 0; attempt < 2; ++attempt) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cudaLaunchAttribute attrs[3];
    int n = 0;
    if (cluster2) {
      attrs[n].id = cudaLaunchAttributeClusterDimension;
      attrs[n].val.clusterDim.x = 2; attrs[n].val.clusterDim.y = 1; attrs[n].val.clusterDim.z = 1;
      ++n;
    }
    if (coop) { attrs[n].id = cudaLaunchAttributeCooperative; attrs[n].val.cooperative = 1; ++n; }
    const bool pdl = g_pdl != 0;
    if (pdl) { attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; attrs[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cfg.attrs = attrs; cfg.numAttrs = n;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e == cudaSuccess) return 0;
    if (pdl && (e == cudaErrorInvalidValue || e == cudaErrorNotSupported)) {
      cudaGetLastError();
      g_pdl = 0;
      continue;
    }
    return fail(std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(e));
  }
  return fail("cudaLaunchKernelEx failed");
}

// The opt-in shared-memory size is a per-device function attribute: set it once per (kernel, device).  `done` must be
// a static of the CALLER's template instantiation (one per kernel), hence the macro.
#define ENSURE_SMEM(kern, device, smem)                                                              \
  do {                                                                                               \
    static bool done_[64] = {};                                                                      \
    const int d_ = (device);                                                                         \
    if (d_ < 0 || d_ >= 64 || !done_[d_]) {                                                          \
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));             \
      if (d_ >= 0 && d_ < 64) done_[d_] = true;                                                      \
    }                                                                                                \
  } while (0)

template <int BN, int EPI, int PREC, int NSTAGE = 0>
static int launch_conv_t(const ConvLaunch& c, int device, cudaStream_t stream) {
  auto kern = conv_umma_kernel<BN, EPI, PREC, NSTAGE>;
  using Cfg = ConvCfg<BN, NSTAGE>;
  ENSURE_SMEM(kern, device, Cfg::SMEM_BYTES);
  return launch_ex(kern, c.grid, 192, Cfg::SMEM_BYTES, stream, false, false, c.tmA[0], c.tmA[1], c.tmA[2], c.tmB, c.p);
}
// Cooperative launch: every CTA of the grid is resident at once (the persistent ConvLSTM's CTAs wait for each other).
template <int BN, int EPI, int PREC>
static int launch_conv_coop(const ConvLaunch& c, int device, cudaStream_t stream) {
  auto kern = conv_umma_kernel<BN, EPI, PREC, 0>;
  using Cfg = ConvCfg<BN, 0>;
  ENSURE_SMEM(kern, device, Cfg::SMEM_BYTES);
  static const int coop = getenv("WDG_LSTM_COOP") ? atoi(getenv("WDG_LSTM_COOP")) : 1;     // see lstm_pair_config
  return launch_ex(kern, c.grid, 192, Cfg::SMEM_BYTES, stream, false, coop != 0, c.tmA[0], c.tmA[1], c.tmA[2], c.tmB, c.p);
}
// CTA-pair convolution (conv_umma2.cuh), plain cluster launch
template <int BN, int EPI, int PREC>
static int launch_conv_pair(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmB_half, const ConvParams& p, int grid,
                            int device, cudaStream_t stream) {
  auto kern = conv_pair_kernel<BN, EPI, PREC>;
  ENSURE_SMEM(kern, device, PairCfg<BN>::SMEM_BYTES);
  return launch_ex(kern, grid, 192, PairCfg<BN>::SMEM_BYTES, stream, true, false, tmA0, tmA1, tmB_half, p);
}
template <int PREC>
static int launch_lstm_pair(const Plan& pl, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[2];
  if (lstm_pair_config<PREC>(&cfg, attrs, pl.pair_grid, stream)) return 1;     // sets the shared-memory attribute; numAttrs == 2: cooperative
  return launch_ex(conv_pair_kernel<256, EPI_LSTM, PREC>, pl.pair_grid, 192, PairCfg<256>::SMEM_BYTES, stream, true, cfg.numAttrs == 2,
                   pl.LSP.tmA[0], pl.LSP.tmA[1], pl.tmB_half, pl.pair_p);
}
template <int PREC>
static int launch_conv(const ConvLaunch& c, int device, cudaStream_t stream) {
  if (c.epi == EPI_LSTM) return launch_conv_t<256, EPI_LSTM, PREC>(c, device, stream);
  if (c.epi == EPI_UPCONV) return launch_conv_t<64, EPI_UPCONV, PREC>(c, device, stream);
  switch (c.bn) {
    case 128: return c.shallow ? launch_conv_t<128, EPI_AFFINE, PREC, 3>(c, device, stream) : launch_conv_t<128, EPI_AFFINE, PREC>(c, device, stream);
    case 64: return launch_conv_t<64, EPI_AFFINE, PREC>(c, device, stream);
    case 48: return c.shallow ? launch_conv_t<48, EPI_AFFINE, PREC, 3>(c, device, stream) : launch_conv_t<48, EPI_AFFINE, PREC>(c, device, stream);
  }
  return fail("no kernel instantiated for this BN");
}
template <int BN, int NCHUNK, int NTAP, int TPS, int EPI, int PREC>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloParams& hp, int grid, int device, cudaStream_t stream) {
  auto kern = halo_conv_kernel<BN, NCHUNK, NTAP, TPS, EPI, PREC>;
  constexpr int smem = HaloCfg<BN, NCHUNK, TPS>::SMEM;
  ENSURE_SMEM(kern, device, smem);
  return launch_ex(kern, grid, 224, smem, stream, false, false, tmA, tmB, hp);
}

// CTA-pair form: clusters of 2 (the two SMs of a TPC), one tcgen05.mma.cta_group::2 per pair of passes
template <int BN, int NCHUNK, int NTAP, int TPS, int EPI, int PREC>
static int launch_halo_pair(const CUtensorMap& tmA, const CUtensorMap& tmB_half, const HaloParams& hp, int grid, int device,
                            cudaStream_t stream) {
  auto kern = halo_conv_kernel<BN, NCHUNK, NTAP, TPS, EPI, PREC, true>;
  constexpr int smem = HaloCfg<BN / 2, NCHUNK, TPS>::SMEM;
  ENSURE_SMEM(kern, device, smem);
  return launch_ex(kern, grid, 224, smem, stream, true, false, tmA, tmB_half, hp);
}

// noise_dev == nullptr: the noise is drawn inside the packing kernel from `ns` (std, key, first counter block)
template <int PREC>
static int run_plan_t(wdg_generator* g, const Plan& pl, const float* image_dev, const float* noise_dev, const NoiseSpec& ns,
                      float* out_dev, cudaStream_t stream, bool profile) {
  using act_t = typename Prec<PREC>::act_t;
  constexpr int NC128 = 128 / Prec<PREC>::KB_ELEMS;     // K chunks per 128 channels: 2 (bf16) / 4 (tf32)
  constexpr int NC192 = 192 / Prec<PREC>::KB_ELEMS;     // chunks of the 8x8 conv's 192-element window
  constexpr int NC160 = (160 + Prec<PREC>::KB_ELEMS - 1) / Prec<PREC>::KB_ELEMS;   // concat(g7, res_2): 3 / 5
  const long long N = (long long)pl.B * pl.T, S = g->S;
  const long long npix = N * S * S;
  const int dev = g->device;
  int stage_i = 0;
  auto mark = [&]() { if (g->profiling && profile) cudaEventRecord(g->ev[stage_i++], stream); };
  // step counters of the persistent ConvLSTM: cleared here, not next to its launch, so that the kernel chain stays unbroken
  if (pl.lstm_persist) CK(cudaMemsetAsync(pl.lstm_flags, 0, (size_t)pl.T * pl.LSP.p.tiles_n * sizeof(unsigned long long), stream));
  mark();
  if (noise_dev) {
    pack_input_s2d_kernel<PREC><<<(unsigned)(N * S), 128, (size_t)S * (g->cin + g->cnoise) * sizeof(float), stream>>>(
        image_dev, noise_dev, (act_t*)pl.xpad, (int)S, g->cin, g->cnoise, g->CP);
  } else {
    if (!fused_noise_ok(g)) return fail("in-kernel noise is built for 3 + 20 input channels");
    pack_input_gen_noise_kernel<PREC, 3, 20, 24><<<(unsigned)(N * S), 96, (size_t)S * 3 * sizeof(float), stream>>>(
        image_dev, ns, (act_t*)pl.xpad, (int)S);
  }
  CK(cudaGetLastError());
  mark();
  if (pl.use_halo) {
    if (pl.halo_pair) {
      if (launch_halo_pair<128, NC192, 8, 2, HEPI_AFFINE, PREC>(pl.h0A, pl.h0B_half, pl.h0p, pl.h0grid_pair, dev, stream)) return 1;
    } else if (launch_halo<128, NC192, 8, 2, HEPI_AFFINE, PREC>(pl.h0A, pl.h0B, pl.h0p, pl.h0grid, dev, stream)) return 1;
  } else if (launch_conv<PREC>(pl.L0, dev, stream)) return 1;
  mark();
  if (pl.l2_pair) {
    if (launch_conv_pair<128, EPI_AFFINE, PREC>(pl.L2.tmA[0], pl.L2.tmA[1], pl.l2B_half, pl.L2.p, pl.l2_pair_grid, dev, stream)) return 1;
  } else if (launch_conv<PREC>(pl.L2, dev, stream)) return 1;
  mark();
  if (pl.lstm_persist) {
    if (pl.lstm_pair) {
      if (launch_lstm_pair<PREC>(pl, stream)) return 1;
    } else if (launch_conv_coop<256, EPI_LSTM, PREC>(pl.LSP, dev, stream)) return 1;
  } else {
    for (int t = 0; t < pl.T; ++t)
      if (launch_conv<PREC>(pl.LS[t], dev, stream)) return 1;
  }
  mark();
  if (pl.use_halo5) {
    if (pl.halo_pair) {
      if (launch_halo_pair<64, NC128, 9, 3, HEPI_AFFINE, PREC>(pl.h5A, pl.h5B_half, pl.h5p, pl.h5grid_pair, dev, stream)) return 1;
    } else if (launch_halo<64, NC128, 9, 3, HEPI_AFFINE, PREC>(pl.h5A, pl.h5B, pl.h5p, pl.h5grid, dev, stream)) return 1;
  } else if (launch_conv<PREC>(pl.L5, dev, stream)) return 1;
  mark();
  if (launch_conv<PREC>(pl.L7, dev, stream)) return 1;
  mark();
  {
    const int CI = g->F / 4 + 128;
    const long long total = N * 4 * (S + 8) * (CI / 8);
    edge_lines_kernel<PREC><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((const act_t*)pl.catp, (act_t*)pl.edgeE, total,
                                                                               (int)(S / 2), CI, g->catp_pitch());
    CK(cudaGetLastError());
    if (launch_conv<PREC>(pl.LE, dev, stream)) return 1;
  }
  mark();
  if (pl.use_halo) {
    if (pl.halo_pair) {
      if (launch_halo_pair<64, NC160, 16, 4, HEPI_UPCONV, PREC>(pl.hA, pl.hB_half, pl.hp, pl.hgrid_pair, dev, stream)) return 1;
    } else if (launch_halo<64, NC160, 16, 4, HEPI_UPCONV, PREC>(pl.hA, pl.hB, pl.hp, pl.hgrid, dev, stream)) return 1;
  } else if (launch_conv<PREC>(pl.L9, dev, stream)) return 1;
  mark();
  bool done11 = false;
  if constexpr (PREC == PREC_BF16) {
    if (pl.use_halo11) {
      HaloParams hp = pl.h11p;
      hp.outf = out_dev;
      if (launch_halo<16, 1, 9, 3, HEPI_FINAL, PREC_BF16>(pl.h11A, pl.h11B, hp, pl.h11grid, dev, stream)) return 1;
      done11 = true;
    }
  }
  if (!done11) {
    // tf32 path: the output layer (0.14 % of the MACs) runs in full fp32 on the CUDA cores, on the unrounded g9.
    // Measured, 512 fields: tiled + register-blocked 0.25 ms (shared-memory port and uniform-register weight loads bound),
    // per-pixel gather form 0.41 ms; the bf16 super-pixel tensor-core form takes 0.078 ms, which is why bf16 keeps it.
    const long long G9C = g->F / 8, G9X = S + 8, G9Y = S + 2;
    const int smem = (FT_ROWS + 2) * ((int)S + 2) * FT_PITCH * (int)sizeof(float);
    if (G9C == 16 && smem <= 200 * 1024) {
      auto kern = final_conv3x3_tiled_kernel<PREC>;
      ENSURE_SMEM(kern, dev, smem);
      const unsigned blocks = (unsigned)(N * ((S + FT_ROWS - 1) / FT_ROWS));
      // view starting at padded (row 0, col 3): the left zero ring pixel of the interior
      kern<<<blocks, 192, smem, stream>>>((const act_t*)pl.g9 + 3 * G9C, G9Y * G9X * G9C, G9X * G9C, g->w11h, out_dev, (int)S);
    } else {
      final_conv3x3_kernel<16, 2, PREC><<<(unsigned)((npix + 127) / 128), 128, 0, stream>>>(
          (const act_t*)pl.g9 + (G9X + 4) * G9C, G9Y * G9X * G9C, G9X * G9C, g->w11h, out_dev, npix, (int)S);
    }
    CK(cudaGetLastError());
  }
  mark();
  return 0;
}

static int run_plan(wdg_generator* g, const Plan& pl, const float* image_dev, const float* noise_dev, const NoiseSpec& ns,
                    float* out_dev, cudaStream_t stream, bool profile) {
  return g->prec == PREC_TF32 ? run_plan_t<PREC_TF32>(g, pl, image_dev, noise_dev, ns, out_dev, stream, profile)
                              : run_plan_t<PREC_BF16>(g, pl, image_dev, noise_dev, ns, out_dev, stream, profile);
}
static NoiseSpec noise_spec(float stddev, uint64_t seed, uint64_t offset) {
  return NoiseSpec{stddev, (uint32_t)seed, (uint32_t)(seed >> 32), (unsigned long long)offset};
}


extern "C" int wdg_generator_forward(wdg_generator* g, const float* image_dev, const float* noise_dev, float* out_dev,
                                     void* stream_) {
  if (!g || !image_dev || !noise_dev || !out_dev) return fail("null argument");
  if (g->plans[0].B == 0) return fail("wdg_generator_bind must be called before forward");
  return run_plan(g, g->plans[0], image_dev, noise_dev, NoiseSpec{}, out_dev, (cudaStream_t)stream_, true);
}

extern "C" int wdg_generator_forward_gen_noise(wdg_generator* g, const float* image_dev, float noise_std, uint64_t noise_seed,
                                               uint64_t noise_offset, float* out_dev, void* noise_scratch_dev, void* stream_) {
  if (!g || !image_dev || !out_dev) return fail("null argument");
  const Plan& pl = g->plans[0];
  if (pl.B == 0) return fail("wdg_generator_bind must be called before forward");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (fused_noise_ok(g))
    return run_plan(g, pl, image_dev, nullptr, noise_spec(noise_std, noise_seed, noise_offset), out_dev, stream, true);
  // other channel counts: draw the tensor first
  if (!noise_scratch_dev) return fail("this channel configuration needs a (B,T,S,S,Cnoise) fp32 noise scratch buffer");
  const long long n = (long long)pl.B * pl.T * g->S * g->S * g->cnoise;
  if (wdg_noise_normal((float*)noise_scratch_dev, n, noise_std, noise_seed, noise_offset, stream)) return 1;
  return run_plan(g, pl, image_dev, (const float*)noise_scratch_dev, NoiseSpec{}, out_dev, stream, true);
}

extern "C" int wdg_generator_profile(wdg_generator* g, int enable) {
  if (!g) return fail("null handle");
  if (enable)
    for (auto& e : g->ev)
      if (!e) CK(cudaEventCreate(&e));
  g->profiling = enable != 0;
  return 0;
}

extern "C" int wdg_generator_stage_ms(wdg_generator* g, float* ms, int n) {
  if (!g || !ms || n != WDG_NUM_STAGES) return fail("bad argument");
  if (!g->profiling) return fail("profiling is off");
  CK(cudaEventSynchronize(g->ev[WDG_NUM_STAGES]));
  for (int i = 0; i < WDG_NUM_STAGES; ++i) CK(cudaEventElapsedTime(&ms[i], g->ev[i], g->ev[i + 1]));
  return 0;
}

// Host-buffer entry point.  For B >= 2*CHUNK_B the batch is cut into pieces (make_schedules) and pipelined over three
// streams: H2D copy of piece i+1 | forward of piece i | D2H copy of piece i-1 (double-buffered staging), so the step
// costs ~max(PCIe, compute) instead of their sum.  Sequences are independent, so the cut does not change results.
static int predict_host_impl(wdg_generator* g, const float* image_host, const float* noise_host, float noise_std,
                             uint64_t noise_seed, uint64_t noise_offset, float* out_host, void* io_dev, void* stream_) {
  if (!g || !image_host || !out_host || !io_dev) return fail("null argument");
  const Plan& full = g->plans[0];
  if (full.B == 0) return fail("wdg_generator_bind must be called before predict_host");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t seq_px = (size_t)full.T * g->S * g->S;
  const size_t s_img = seq_px * g->cin * 4, s_noise = seq_px * g->cnoise * 4, s_out = seq_px * g->cout * 4;
  uint8_t* io = (uint8_t*)io_dev;
  const std::vector<int>& sched = noise_host ? g->sched_host : g->sched_dev;
  if (sched.empty()) {
    const size_t b_img = full.B * s_img, b_noise = full.B * s_noise, b_out = full.B * s_out;
    float* d_img = (float*)io;
    float* d_noise = (float*)(io + align_up(b_img, 256));
    float* d_out = (float*)(io + align_up(b_img, 256) + align_up(b_noise, 256));
    CK(cudaMemcpyAsync(d_img, image_host, b_img, cudaMemcpyHostToDevice, stream));
    const bool fused = !noise_host && fused_noise_ok(g);
    if (noise_host) CK(cudaMemcpyAsync(d_noise, noise_host, b_noise, cudaMemcpyHostToDevice, stream));
    else if (!fused && wdg_noise_normal(d_noise, (long long)(b_noise / 4), noise_std, noise_seed, noise_offset, stream)) return 1;
    if (run_plan(g, full, d_img, fused ? nullptr : d_noise, noise_spec(noise_std, noise_seed, noise_offset), d_out, stream, false)) return 1;
    CK(cudaMemcpyAsync(out_host, d_out, b_out, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }
  if (!g->copy_in) {
    CK(cudaStreamCreateWithFlags(&g->copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&g->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&g->ev_h2d[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g->ev_fwd[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&g->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  const int cb = g->max_piece;                   // staging slots are sized for the largest piece
  const size_t slot = align_up(cb * s_img, 256) + align_up(cb * s_noise, 256) + align_up(cb * s_out, 256);
  // order the side streams after whatever the caller queued on `stream`
  CK(cudaEventRecord(g->ev_fwd[0], stream));
  CK(cudaStreamWaitEvent(g->copy_in, g->ev_fwd[0], 0));
  int b0 = 0;
  for (int i = 0; i < (int)sched.size(); b0 += sched[i], ++i) {
    const int nb = sched[i];
    const Plan& pl = g->piece_plans.at(nb);
    const int buf = i & 1;
    float* d_img = (float*)(io + buf * slot);
    float* d_noise = (float*)(io + buf * slot + align_up(cb * s_img, 256));
    float* d_out = (float*)(io + buf * slot + align_up(cb * s_img, 256) + align_up(cb * s_noise, 256));
    if (i >= 2) CK(cudaStreamWaitEvent(g->copy_in, g->ev_fwd[buf], 0));       // inputs of chunk i-2 consumed
    CK(cudaMemcpyAsync(d_img, (const uint8_t*)image_host + b0 * s_img, nb * s_img, cudaMemcpyHostToDevice, g->copy_in));
    if (noise_host)
      CK(cudaMemcpyAsync(d_noise, (const uint8_t*)noise_host + b0 * s_noise, nb * s_noise, cudaMemcpyHostToDevice, g->copy_in));
    CK(cudaEventRecord(g->ev_h2d[buf], g->copy_in));
    CK(cudaStreamWaitEvent(stream, g->ev_h2d[buf], 0));
    if (i >= 2) CK(cudaStreamWaitEvent(stream, g->ev_d2h[buf], 0));          // output of chunk i-2 copied out
    // element index of this chunk in the whole noise tensor is a multiple of 4: counter blocks line up
    const uint64_t chunk_off = noise_offset + b0 * (s_noise / 16);
    const bool fused = !noise_host && fused_noise_ok(g);
    if (!noise_host && !fused && wdg_noise_normal(d_noise, (long long)(nb * s_noise / 4), noise_std, noise_seed, chunk_off, stream))
      return 1;
    if (run_plan(g, pl, d_img, fused ? nullptr : d_noise, noise_spec(noise_std, noise_seed, chunk_off), d_out, stream, false)) return 1;
    CK(cudaEventRecord(g->ev_fwd[buf], stream));
    CK(cudaStreamWaitEvent(g->copy_out, g->ev_fwd[buf], 0));
    CK(cudaMemcpyAsync((uint8_t*)out_host + b0 * s_out, d_out, nb * s_out, cudaMemcpyDeviceToHost, g->copy_out));
    CK(cudaEventRecord(g->ev_d2h[buf], g->copy_out));
  }
  CK(cudaStreamSynchronize(g->copy_out));
  CK(cudaStreamSynchronize(stream));
  return 0;
}

extern "C" int wdg_generator_predict_host(wdg_generator* g, const float* image_host, const float* noise_host,
                                          float* out_host, void* io_dev, void* stream) {
  if (!noise_host) return fail("null argument");
  return predict_host_impl(g, image_host, noise_host, 0.f, 0, 0, out_host, io_dev, stream);
}

extern "C" int wdg_generator_predict_host_gen_noise(wdg_generator* g, const float* image_host, float noise_std,
                                                    uint64_t noise_seed, uint64_t noise_offset, float* out_host,
                                                    void* io_dev, void* stream) {
  return predict_host_impl(g, image_host, nullptr, noise_std, noise_seed, noise_offset, out_host, io_dev, stream);
}

extern "C" int wdg_generator_launches_per_forward(const wdg_generator* g) { return g ? g->plans[0].launches : 0; }

// --------------------------------------------------------------- debug
template <int PREC>
__global__ void act_to_f32_strided(const typename Prec<PREC>::act_t* src, float* dst, long long n_img, int H, int W, int C,
                                   long long sn, long long sy, long long sx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = n_img * H * W * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pix = i / C;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
  dst[i] = Prec<PREC>::load(src + n * sn + y * sy + x * sx + c);
}

extern "C" int wdg_generator_debug_read(const wdg_generator* g, int which, float* host_out, int64_t count) {
  if (!g || g->plans[0].B == 0 || !host_out) return fail("bad argument");
  const Plan& pl = g->plans[0];
  const long long N = (long long)pl.B * pl.T, S = g->S, S2 = S / 2, S4 = S / 4, F = g->F, CI = g->catp_pitch();
  const uint8_t* base;
  long long off;   // element offset
  int H, W, C;
  long long sn, sy, sx;
  switch (which) {
    case 0: H = W = (int)S2; C = 128; sx = CI; sy = (S2 + 4) * sx; sn = (S2 + 4) * sy; base = pl.catp; off = 2 * sy + 2 * sx + F / 4; break;
    case 1: H = W = (int)S4; C = (int)F; sx = F; sy = S4 * F; sn = S4 * sy; base = pl.res4; off = 0; break;
    case 2: H = W = (int)S4; C = (int)F; sx = F; sy = (S4 + 2) * F; sn = (S4 + 2) * sy; base = pl.hseq; off = sy + sx; break;
    case 3: H = W = (int)S4; C = (int)(F / 2); sx = C; sy = S4 * sx; sn = S4 * sy; base = pl.g5; off = 0; break;
    case 4: H = W = (int)S2; C = (int)(F / 4); sx = CI; sy = (S2 + 4) * sx; sn = (S2 + 4) * sy; base = pl.catp; off = 2 * sy + 2 * sx; break;
    case 5: H = W = (int)S; C = (int)(F / 8); sx = C; sy = (S + 8) * sx; sn = (S + 2) * sy; base = pl.g9; off = sy + 4 * sx; break;
    default: return fail("unknown intermediate");
  }
  const long long total = N * H * W * C;
  if (count != total) return fail("count mismatch");
  float* d = nullptr;
  CK(cudaMalloc(&d, total * sizeof(float)));
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (g->prec == PREC_TF32)
    act_to_f32_strided<PREC_TF32><<<grid, 256>>>((const float*)base + off, d, N, H, W, C, sn, sy, sx);
  else
    act_to_f32_strided<PREC_BF16><<<grid, 256>>>((const __nv_bfloat16*)base + off, d, N, H, W, C, sn, sy, sx);
  cudaError_t e = cudaMemcpy(host_out, d, total * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(std::string("debug_read: ") + cudaGetErrorString(e));
  return 0;
}

// ====================================================================================================================
// Training-path entry point for the generator's largest layer: Concatenate([x, res_2]) -> UpSampling2D(2, bilinear) ->
// Conv2DTranspose(16, 5x5, same) + bias -> LeakyReLU (models.py:60-64), fp32 tensors in, fp32 out, tf32 operands.
// The training graph used to materialise the 96x96x160 upsampled tensor (1.1 GB at batch 8 x 24), run a 1x1 GEMM into
// 25 tap columns (2.8 GB written and read back) and overlap-add them: 3.1 ms per forward, five forwards per WGAN step.
// This runs the inference engine's kernels instead (fused bilinear + conv over 4x4 low-res windows with host... here
// DEVICE-composed phase weights, exact border correction): ~0.5 ms.  The weights change with every optimizer step, so
// the composition (generator finalize does it on the host in double) is a kernel here.
namespace {

__device__ __forceinline__ double up_w(int m, int d) {     // bilinear taps U[m][d] of wdg_generator_finalize
  const double U[6][4] = {{.75, .25, 0, 0}, {.25, .75, 0, 0}, {0, .75, .25, 0}, {0, .25, .75, 0}, {0, 0, .75, .25}, {0, 0, .25, .75}};
  return U[m][d];
}
// w: [5][5][16][160] (kh, kw, out, in).  B9h: [64][5 chunks * 16 taps][32] (K-block = chunk*16 + tap);
// BE: [192][5 taps * 5 chunks][32] (K-block = t*5 + chunk).  tf32-rounded.
__global__ void compose_upconv_kernel(const float* __restrict__ w, float* __restrict__ B9h, float* __restrict__ BE) {
  const int O = 16, I = 160;
  auto Wf = [&](int ty, int tx, int c, int o) { return (double)w[(((4 - ty) * 5 + (4 - tx)) * O + o) * I + c]; };
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 80 * 32) {
    const int n = i / (80 * 32), kb = (i / 32) % 80, j = i % 32;
    const int tap = kb % 16, chunk = kb / 16, c = chunk * 32 + j;
    const int py = n / (2 * O), px = (n / O) % 2, o = n % O, dy = tap / 4, dx = tap % 4;
    double acc = 0;
    for (int ty = 0; ty < 5; ++ty) {
      const double uy = up_w(py + ty, dy);
      if (uy == 0) continue;
      for (int tx = 0; tx < 5; ++tx) acc += uy * up_w(px + tx, dx) * Wf(ty, tx, c, o);
    }
    B9h[i] = __uint_as_float(to_tf32((float)acc));
  }
  if (i < 192 * 25 * 32) {
    const int n = i / (25 * 32), kb = (i / 32) % 25, j = i % 32;
    const int edge = n / 48, e = (n % 48) / 16, o = n % 16, t = kb / 5, chunk = kb % 5, c = chunk * 32 + j;
    double v;
    switch (edge) {
      case 0: v = Wf(2 - e, t, c, o) - (e <= 1 ? Wf(1 - e, t, c, o) : 0.0); break;
      case 1: v = Wf(2 + e, t, c, o) - (e <= 1 ? Wf(3 + e, t, c, o) : 0.0); break;
      case 2: v = Wf(t, 2 - e, c, o) - (e <= 1 ? Wf(t, 1 - e, c, o) : 0.0); break;
      default: v = Wf(t, 2 + e, c, o) - (e <= 1 ? Wf(t, 3 + e, c, o) : 0.0); break;
    }
    BE[i] = __uint_as_float(to_tf32((float)(0.25 * v)));
  }
}

// Concatenate([a (Ca channels), b (Cb channels)]) of dense [N][h][w][*] fp32 tensors into the zero-ring-2 padded image
// [N][h+4][w+4][Ca+Cb], values rounded to tf32 (GEMM operand).  One thread per (padded pixel, 4 channels).
__global__ void pad_concat_tf32_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b, int Cb,
                                       float* __restrict__ out, long long total, int h, int w) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int C = Ca + Cb, g4 = C / 4;
  const int c = (int)(i % g4) * 4;
  const long long pp = i / g4;
  const int px = (int)(pp % (w + 4)), py = (int)((pp / (w + 4)) % (h + 4));
  const long long n = pp / ((long long)(w + 4) * (h + 4));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int y = py - 2, x = px - 2;
  if (y >= 0 && y < h && x >= 0 && x < w) {
    const long long pix = (n * h + y) * w + x;
    v = c < Ca ? *reinterpret_cast<const float4*>(a + pix * Ca + c) : *reinterpret_cast<const float4*>(b + pix * Cb + (c - Ca));
    v.x = __uint_as_float(to_tf32(v.x)); v.y = __uint_as_float(to_tf32(v.y));
    v.z = __uint_as_float(to_tf32(v.z)); v.w = __uint_as_float(to_tf32(v.w));
  }
  *reinterpret_cast<float4*>(out + pp * C + c) = v;
}

__global__ void fill_vec_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct UpconvWs {
  size_t catp, edgeE, deltaD, B9h, BE, vecs, total;
};
UpconvWs upconv_ws(long long N, int h) {
  const size_t S = 2 * (size_t)h, PW = h + 4;
  UpconvWs L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 1024); return r; };
  L.catp = take((size_t)N * PW * PW * 160 * 4 + 8192);
  L.edgeE = take((size_t)N * 4 * (S + 8) * 160 * 4);
  L.deltaD = take((size_t)N * S * 192 * 4);
  L.B9h = take((size_t)64 * 80 * 32 * 4);
  L.BE = take((size_t)192 * 25 * 32 * 4);
  L.vecs = take(3 * 192 * 4);      // zeros[192] | ones[192] | (unused)
  L.total = o;
  return L;
}

}  // namespace

extern "C" int wdg_upconv5x5_workspace_bytes(long long N, int h, size_t* bytes) {
  if (!bytes || N <= 0 || h <= 0) return fail("bad argument");
  *bytes = upconv_ws(N, h).total;
  return 0;
}

extern "C" int wdg_upconv5x5_fwd(const float* a, const float* b, const float* w, const float* bias, float* out, long long N,
                                 int h, void* workspace, size_t ws_bytes, void* stream_) {
  if (!a || !b || !w || !bias || !out || !workspace || N <= 0 || h <= 0) return fail("bad argument");
  const int S = 2 * h, PW = h + 4, sms_cap = 148;
  if ((3 * PW + 3 + H_TILES * TILE_M) > H_ROWS) return fail("wdg_upconv5x5_fwd: low-res size too large for the halo kernel (<= 48)");
  if ((uintptr_t)workspace % 1024) return fail("workspace must be 1024-byte aligned");
  const UpconvWs L = upconv_ws(N, h);
  if (ws_bytes < L.total) return fail("workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* ws = (uint8_t*)workspace;
  float* catp = (float*)(ws + L.catp);
  float* edgeE = (float*)(ws + L.edgeE);
  float* deltaD = (float*)(ws + L.deltaD);
  float* B9h = (float*)(ws + L.B9h);
  float* BE = (float*)(ws + L.BE);
  float* zeros = (float*)(ws + L.vecs);
  float* ones = zeros + 192;
  int dev = 0, sms = sms_cap;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // 1. operands: composed weights, constant vectors, padded concat
  compose_upconv_kernel<<<(64 * 80 * 32 + 255) / 256, 256, 0, stream>>>(w, B9h, BE);
  CK(cudaGetLastError());
  fill_vec_kernel<<<1, 192, 0, stream>>>(zeros, 192, 0.f);
  fill_vec_kernel<<<1, 192, 0, stream>>>(ones, 192, 1.f);
  {
    const long long total = N * PW * PW * 40;
    pad_concat_tf32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, 32, b, 128, catp, total, h, h);
    CK(cudaGetLastError());
  }
  // 2. border corrections: upsampled edge lines -> 5-tap GEMM -> deltaD [N][S][192]
  {
    const long long total = N * 4 * (S + 8) * 20;
    edge_lines_kernel<PREC_TF32><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(catp, edgeE, total, h, 160, 160);
    CK(cudaGetLastError());
    ConvLaunch c;
    std::memset(&c.p, 0, sizeof c.p);
    const uint64_t I = 160, P = S + 8;
    uint64_t dims[5] = {I, P, 4, (uint64_t)N, 1};
    uint64_t str[4] = {I, P * I, 4 * P * I, (uint64_t)N * 4 * P * I};
    uint32_t box[5] = {32, 16, 1, 8, 1};
    if (wdg_make_tmap(&c.tmA[0], edgeE, 5, dims, str, box, 128, 4)) return 1;
    c.tmA[1] = c.tmA[0]; c.tmA[2] = c.tmA[0];
    uint64_t bd[2] = {25 * 32, 192}, bs[1] = {25 * 32};
    uint32_t bb[2] = {32, 48};
    if (wdg_make_tmap(&c.tmB, BE, 2, bd, bs, bb, 128, 4)) return 1;
    set_tiles(c.p, 1, S, (int)N, 16, 1, 8, 4, 3);
    c.p.ntile_coord = 2;
    c.p.num_kb = 25;
    for (int t = 0; t < 5; ++t)
      for (int ch = 0; ch < 5; ++ch) {
        KBlock& k = c.p.kb[t * 5 + ch];
        k.src = 0; k.half = 0; k.o0 = (int16_t)(ch * 32); k.o1 = (int16_t)(t + 2); k.o2 = 0; k.o3 = 0;
      }
    affine_epi(c.p.ep, zeros, ones, zeros, deltaD, (long long)S * 192, 0, 192, 0, 0);
    c.p.ep.out_f32 = 1;
    c.bn = 48; c.epi = EPI_AFFINE; c.shallow = 1;
    const int total_tiles = c.p.tiles_x * c.p.tiles_y * c.p.tiles_n * c.p.n_tiles_N;
    c.grid = total_tiles < 2 * sms ? total_tiles : 2 * sms;
    if (launch_conv<PREC_TF32>(c, dev, stream)) return 1;
  }
  // 3. fused bilinear x2 + 5x5 transposed conv + bias + LeakyReLU, dense fp32 output (BatchNorm follows in training mode)
  {
    const uint64_t flat = (uint64_t)N * PW * PW;
    CUtensorMap hA, hB;
    uint64_t hd[2] = {160, flat}, hs[1] = {160};
    uint32_t hb[2] = {32, H_BOX_ROWS};
    if (wdg_make_tmap(&hA, catp, 2, hd, hs, hb, 128, 4)) return 1;
    uint64_t bd[2] = {80 * 32, 64}, bs[1] = {80 * 32};
    uint32_t bb[2] = {32, 64};
    if (wdg_make_tmap(&hB, B9h, 2, bd, bs, bb, 128, 4)) return 1;
    HaloParams hp;
    std::memset(&hp, 0, sizeof hp);
    hp.num_passes = (int)((flat + H_TILES * TILE_M - 1) / (H_TILES * TILE_M));
    hp.n_img = (int)N; hp.pw = PW; hp.ph = PW; hp.S = S; hp.delta = deltaD; hp.box_rows = H_BOX_ROWS;
    for (int tap = 0; tap < 16; ++tap) { hp.tap_shift[tap] = (tap / 4) * PW + tap % 4; hp.kmask[tap] = 0xF; }
    hp.bias = bias; hp.scale = ones; hp.shift = zeros; hp.out = out;
    hp.up_sn = (long long)S * S * 16; hp.up_sy = (long long)S * 16;
    const int grid = hp.num_passes < sms ? hp.num_passes : sms;
    if (launch_halo<64, 5, 16, 4, HEPI_UPCONV, PREC_TF32>(hA, hB, hp, grid, dev, stream)) return 1;
  }
  return 0;
}
