// Convolution geometry shared by the training kernels (train_ops.cu: CUDA-core fp32 GEMMs; train_gemm_tc.cu:
// tcgen05 GEMMs).  All tensors are fp32 channels-last; `cs`/`co` are the channel stride / offset of a tensor
// inside a wider (concatenated) buffer.
#pragma once
#include <cuda_runtime.h>

#include <string>

extern int wdg_set_error(const std::string& m);

#define CKT(call)                                                                                    \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) return wdg_set_error(std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct ConvGeo {
  int N, H, W, Ci, kh, kw, Co, stride, pad_t, pad_l, Ho, Wo;
  int x_cs, x_co;   // channel stride / offset of x (input side, Ci channels)
  int y_cs, y_co;   // channel stride / offset of y (output side, Co channels)
};

static inline ConvGeo make_geo(const int* g) {
  ConvGeo c;
  c.N = g[0]; c.H = g[1]; c.W = g[2]; c.Ci = g[3]; c.kh = g[4]; c.kw = g[5]; c.Co = g[6]; c.stride = g[7];
  c.pad_t = g[8]; c.pad_l = g[9]; c.Ho = g[10]; c.Wo = g[11]; c.x_cs = g[12]; c.x_co = g[13]; c.y_cs = g[14]; c.y_co = g[15];
  return c;
}

// Residue class (ry, rx) of a strided backward-data problem (see BwdDataClassProblem in train_ops.cu).
struct BwdClass {
  int ry, rx, fy, fx, Hc, Wc, Jy, Jx;   // class residues, first pixel of the class, class grid, taps per axis
};
static inline BwdClass make_bwd_class(const ConvGeo& g, int ry, int rx) {
  const int s = g.stride;
  BwdClass c;
  c.ry = ry; c.rx = rx;
  c.fy = ((ry - g.pad_t) % s + s) % s; c.fx = ((rx - g.pad_l) % s + s) % s;
  c.Hc = g.H > c.fy ? (g.H - c.fy + s - 1) / s : 0; c.Wc = g.W > c.fx ? (g.W - c.fx + s - 1) / s : 0;
  c.Jy = (g.kh - ry + s - 1) / s; c.Jx = (g.kw - rx + s - 1) / s;
  return c;
}

// split-K scratch of wdg_conv2d_bwd_weight for a given arithmetic mode (train_ops.cu)
int wdg_wgrad_scratch_for_mode(const int* geo, int mode, size_t* bytes, int* splits_out);

// tcgen05 paths (train_gemm_tc.cu).  op: 1 = tf32 operands, 2 = bf16 operands; fp32 accumulation in TMEM.
// alpha: slope of the LeakyReLU applied to the final value (1 = linear)
int wdg_tc_conv2d_fwd(const ConvGeo& g, const float* x, const float* w, const float* bias, float* y, int accumulate, float alpha,
                      int op, cudaStream_t stream);
int wdg_tc_conv2d_bwd_data(const ConvGeo& g, const float* dy, const float* w, float* dx, int accumulate, int op,
                           cudaStream_t stream);
void wdg_tc_wgrad_plan(const ConvGeo& g, int op, int* splits, long long* k_per_split);
int wdg_tc_conv2d_bwd_weight(const ConvGeo& g, const float* x, const float* dy, float* part, int splits, long long k_per_split,
                             int op, cudaStream_t stream);
