// Fused recurrent steps of the critic's full-resolution 16-filter ConvLSTM2D (reference models.py:100-101) for the
// TRAINING path -- the single most expensive family of launches of a WGAN step (ganbase.py:21-94 runs the critic 12
// times forward and 10 times backward per step, 24 timesteps each).  Per timestep the unfused path ran a gather-fed
// GEMM (K = 144, N = 64: latency-bound, 49 us) that read-modify-wrote the 64-channel pre-activations, then a gate
// kernel that read them again (9 us); backward likewise (gate backward 10 us + a K = 576, N = 16 GEMM 38-69 us).
// Here each step is ONE launch of the inference engine's TMA-fed tcgen05 kernel (conv_umma.cuh) with the gate math in
// the TMEM epilogue:
//   forward  step t: acc = conv3x3(h_{t-1}, R) [9 taps x 16 channels: 64-byte SWIZZLE_64B rows straight from the fp32
//                    h tensor, zero-filled by TMA at the image border]; epilogue adds the input conv of step t, applies
//                    i/f/c~/o, writes the activated gates (kept for backward), c_t and h_t.
//   backward step s: acc = conv3x3_bwd_data(dz_{s+1}, R) [9 taps x 64 channels, N = 16] = recurrent part of dL/dh_s;
//                    epilogue applies the gate backward of step s in place (dz_s over the gates) and carries dL/dc.
// Operands are tf32 (kind::tf32): h_t and dz_s are rounded to nearest by the epilogue that produces them (they are only
// ever GEMM operands, so this is the same arithmetic as rounding in the consumer's loader), R is rounded when packed.
// The exact-fp32 training mode keeps the unfused CUDA-core path.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/wdg.h"
#include "conv_umma.cuh"

using namespace wdg;

extern int wdg_set_error(const std::string& m);
extern int wdg_make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                         const uint32_t* box, int inner_bytes, int esz);

#define CKL(call)                                                                                      \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return wdg_set_error(std::string(#call) + ": " + cudaGetErrorString(_e));   \
  } while (0)

namespace {

constexpr int F16 = 16;

// Forward B operand: [64 gate columns][9 taps x 32]: elements 0..15 of each K-block = R[tap][ci][col], 16..31 = 0
// (the A rows are 64 bytes wide, so only the first two K = 8 slices of every block are multiplied).
// Backward B operand: [16 input channels][9 taps x 2 x 32]: R[ky][kx][ci][chunk*32 + j] for the tap the A map shifts by
// (1 - ky, 1 - kx).
__global__ void pack_lstm16_kernel(const float* __restrict__ R, float* __restrict__ Bf, float* __restrict__ Bb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 64 * 9 * 32) {
    const int col = i / (9 * 32), kb = (i / 32) % 9, j = i % 32;
    Bf[i] = j < F16 ? __uint_as_float(to_tf32(R[(kb * F16 + j) * 64 + col])) : 0.f;
  }
  if (i < 16 * 18 * 32) {
    const int ci = i / (18 * 32), kb = (i / 32) % 18, j = i % 32;
    const int tap = kb / 2, chunk = kb % 2;
    Bb[i] = __uint_as_float(to_tf32(R[(tap * F16 + ci) * 64 + chunk * 32 + j]));
  }
}

// 128-filter cell, forward B operand: [512 rows][36 K-blocks x 32]: row = n_tile*256 + gate*64 + cc  <->  recurrent-kernel
// column gate*128 + n_tile*64 + cc; K-block = tap*4 + chunk, element j = input channel chunk*32 + j.
__global__ void pack_lstm128_kernel(const float* __restrict__ R, float* __restrict__ Bf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 512 * 36 * 32) return;
  const int row = i / (36 * 32), kb = (i / 32) % 36, j = i % 32;
  const int n_tile = row / 256, gate = (row % 256) / 64, cc = row % 64;
  const int tap = kb / 4, ci = (kb % 4) * 32 + j;
  Bf[i] = __uint_as_float(to_tf32(R[((long long)tap * 128 + ci) * 512 + gate * 128 + n_tile * 64 + cc]));
}

__global__ void round_tf32_kernel(float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = __uint_as_float(to_tf32(x[i]));
}

void set_tiles(ConvParams& p, int H, int W, int N) {
  p.H = H; p.W = W; p.N = N;
  p.tile_w = 16; p.tile_h = 8; p.tile_n = 1;
  p.tiles_x = (W + 15) / 16; p.tiles_y = (H + 7) / 8; p.tiles_n = N;
  p.n_tiles_N = 1;
  p.n_coord = 3;
  p.ntile_coord = -1;
}

int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int BN, int EPI, int NSTAGE = 3>
int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvParams& p, cudaStream_t stream) {
  auto kern = conv_umma_kernel<BN, EPI, PREC_TF32, NSTAGE>;
  using Cfg = ConvCfg<BN, NSTAGE>;
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !done[dev]) {
    CKL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  const int total = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_N;
  const int cap = sm_count() * Cfg::CTAS_PER_SM;
  kern<<<total < cap ? total : cap, 192, Cfg::SMEM_BYTES, stream>>>(a0, a1, a0, b, p);
  CKL(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int wdg_lstm16_pack(const float* R, float* packed, void* stream) {
  if (!R || !packed) return wdg_set_error("null argument");
  pack_lstm16_kernel<<<(64 * 9 * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(R, packed, packed + WDG_LSTM16_PACK_FWD_FLOATS);
  CKL(cudaGetLastError());
  return 0;
}

extern "C" int wdg_round_tf32(float* x, long long n, void* stream) {
  if (!x || n < 0) return wdg_set_error("bad argument");
  if (n) round_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n);
  CKL(cudaGetLastError());
  return 0;
}

extern "C" int wdg_lstm16_fwd_step(float* gates, const float* h_prev, const float* packed, const float* c_prev, float* c_out,
                                   float* h_out, int N, int H, int W, void* stream) {
  if (!gates || !h_prev || !packed || !c_prev || !c_out || !h_out || N <= 0 || H <= 0 || W <= 0) return wdg_set_error("bad argument");
  CUtensorMap tmA, tmB;
  const uint64_t dims[5] = {F16, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
  const uint64_t str[4] = {F16, (uint64_t)W * F16, (uint64_t)H * W * F16, (uint64_t)N * H * W * F16};
  const uint32_t box[5] = {F16, 16, 8, 1, 1};
  if (wdg_make_tmap(&tmA, h_prev, 5, dims, str, box, 64, 4)) return 1;
  const uint64_t bd[2] = {9 * 32, 64}, bs[1] = {9 * 32};
  const uint32_t bb[2] = {32, 64};
  if (wdg_make_tmap(&tmB, packed, 2, bd, bs, bb, 128, 4)) return 1;
  ConvParams p;
  std::memset(&p, 0, sizeof p);
  set_tiles(p, H, W, N);
  p.num_kb = 9;
  for (int tap = 0; tap < 9; ++tap) {
    KBlock& k = p.kb[tap];
    k.src = 1; k.half = 1; k.o0 = 0; k.o1 = (int16_t)(tap % 3 - 1); k.o2 = (int16_t)(tap / 3 - 1); k.o3 = 0;
  }
  p.ep.t_gates = gates; p.ep.t_c_prev = c_prev; p.ep.t_c = c_out; p.ep.t_h = h_out;
  return launch<64, EPI_LSTM16_FWD>(tmA, tmA, tmB, p, (cudaStream_t)stream);
}

extern "C" int wdg_lstm16_bwd_step(const float* dz_next, const float* packed, float* gates_s, const float* c_prev, const float* c_cur,
                                   const float* dh, float* dc, int N, int H, int W, void* stream) {
  if (!dz_next || !packed || !gates_s || !c_cur || !dh || !dc || N <= 0 || H <= 0 || W <= 0) return wdg_set_error("bad argument");
  CUtensorMap tmA, tmB;
  const uint64_t C = 4 * F16;
  const uint64_t dims[5] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
  const uint64_t str[4] = {C, (uint64_t)W * C, (uint64_t)H * W * C, (uint64_t)N * H * W * C};
  const uint32_t box[5] = {32, 16, 8, 1, 1};
  if (wdg_make_tmap(&tmA, dz_next, 5, dims, str, box, 128, 4)) return 1;
  const uint64_t bd[2] = {18 * 32, 16}, bs[1] = {18 * 32};
  const uint32_t bb[2] = {32, 16};
  if (wdg_make_tmap(&tmB, packed + WDG_LSTM16_PACK_FWD_FLOATS, 2, bd, bs, bb, 128, 4)) return 1;
  ConvParams p;
  std::memset(&p, 0, sizeof p);
  set_tiles(p, H, W, N);
  p.num_kb = 18;
  for (int kb = 0; kb < 18; ++kb) {
    KBlock& k = p.kb[kb];
    const int tap = kb / 2, ky = tap / 3, kx = tap % 3;
    k.src = 0; k.half = 0; k.o0 = (int16_t)((kb % 2) * 32); k.o1 = (int16_t)(1 - kx); k.o2 = (int16_t)(1 - ky); k.o3 = 0;
  }
  p.ep.t_gates = gates_s; p.ep.t_c_prev = c_prev; p.ep.t_c = const_cast<float*>(c_cur); p.ep.t_dh = dh; p.ep.t_dc = dc;
  return launch<16, EPI_LSTM16_BWD>(tmA, tmA, tmB, p, (cudaStream_t)stream);
}

// ---- the generator's 128-filter ConvLSTM2D (models.py:45) in the training path: one launch per timestep t >= 1 of the
// inference engine's TMA-fed tcgen05 kernel (K = 9 taps x 128 channels, N = 512 = two tiles of [i|f|c~|o] x 64) with the
// gate math in the TMEM epilogue.  The unfused step was a gather-fed GEMM at M = B*576 rows (94 us, latency-bound) that
// read-modify-wrote the 512-channel pre-activations plus a gate kernel that read them again.
extern "C" int wdg_lstm128_pack(const float* R, float* packed, void* stream) {
  if (!R || !packed) return wdg_set_error("null argument");
  pack_lstm128_kernel<<<(512 * 36 * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(R, packed);
  CKL(cudaGetLastError());
  return 0;
}

extern "C" int wdg_lstm128_fwd_step(float* gates, const float* h_prev, const float* packed, const float* c_prev, float* c_out,
                                    float* h_out, int N, int H, int W, void* stream) {
  if (!gates || !h_prev || !packed || !c_prev || !c_out || !h_out || N <= 0 || H <= 0 || W <= 0) return wdg_set_error("bad argument");
  constexpr uint64_t F = 128;
  CUtensorMap tmA, tmB;
  const uint64_t dims[5] = {F, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
  const uint64_t str[4] = {F, (uint64_t)W * F, (uint64_t)H * W * F, (uint64_t)N * H * W * F};
  const uint32_t box[5] = {32, 8, 8, 2, 1};
  if (wdg_make_tmap(&tmA, h_prev, 5, dims, str, box, 128, 4)) return 1;
  const uint64_t bd[2] = {36 * 32, 512}, bs[1] = {36 * 32};
  const uint32_t bb[2] = {32, 256};
  if (wdg_make_tmap(&tmB, packed, 2, bd, bs, bb, 128, 4)) return 1;
  ConvParams p;
  std::memset(&p, 0, sizeof p);
  p.H = H; p.W = W; p.N = N;
  p.tile_w = 8; p.tile_h = 8; p.tile_n = 2;
  p.tiles_x = (W + 7) / 8; p.tiles_y = (H + 7) / 8; p.tiles_n = (N + 1) / 2;
  p.n_tiles_N = 2;
  p.n_coord = 3;
  p.ntile_coord = -1;
  p.num_kb = 36;
  for (int kb = 0; kb < 36; ++kb) {
    KBlock& k = p.kb[kb];
    const int tap = kb / 4;
    k.src = 0; k.half = 0; k.o0 = (int16_t)((kb % 4) * 32); k.o1 = (int16_t)(tap % 3 - 1); k.o2 = (int16_t)(tap / 3 - 1); k.o3 = 0;
  }
  p.ep.t_F = (int)F;
  p.ep.t_gates = gates; p.ep.t_c_prev = c_prev; p.ep.t_c = c_out; p.ep.t_h = h_out;
  return launch<256, EPI_LSTM128_FWD, 0>(tmA, tmA, tmB, p, (cudaStream_t)stream);
}
