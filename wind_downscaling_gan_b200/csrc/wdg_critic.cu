// wdg_critic -- the WGAN critic of `make_discriminator` (reference gan/models.py:76-142, tf_utils.py:7-32) behind the C ABI:
// graph construction from the seven hyper-parameters, forward, backward-to-input (gradient penalty / generator update,
// ganbase.py:32-35, :60) and backward-to-weights (critic update, :46), spectral-norm power iteration (TFA 0.14).
//
// Host-side graph walker only: every tensor operation is one of the library's own training kernels (train_ops.cu,
// train_gemm_tc.cu, train_lstm16.cu) called through the entry points of include/wdg.h, in the arithmetic selected by
// wdg_train_set_precision.  All device memory is the caller's:
//   variables  one flat fp32 buffer (wdg_critic_num_floats); trainable variables first, the spectral-norm `sn_u`
//              vectors after them, each variable at wdg_critic_weight_info's offset (256-byte aligned, padding zero);
//   gradients  the first wdg_critic_num_trainable_floats of the same layout (one Adam launch / one all-reduce per model);
//   context    what one forward call keeps for its backward (activations, LayerNorm statistics, ConvLSTM gates and, in
//              training mode, a snapshot of the variables the call read: a later in-place spectral-norm update must not
//              leak into this call's backward), sized by wdg_critic_context_bytes;
//   scratch    split-K partials / reduction scratch of the kernels and the backward's temporaries.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/wdg.h"
#include "train_geo.cuh"

namespace {

const float kAlpha = 0.2f;      // LeakyReLU slope (models.py:95 ...)
const float kLnEps = 1e-3f;     // Keras LayerNormalization default epsilon

struct Var {
  std::string name;
  int64_t dims[4];
  int ndim;
  int64_t count, offset;   // floats
  bool trainable;
};

struct ConvSpec {
  int k, stride, pad, cin, cout, size_in, size_out;
  int w, b, u, gamma, beta;   // variable indices
};

struct LstmSpec {
  int cin, F;
  int K, R, b;
};

}  // namespace

struct wdg_critic {
  int size, lr_ch, hr_ch, T, F, ckpt_topology;
  std::vector<Var> vars;
  std::map<std::string, int> index;
  int64_t n_train = 0, n_total = 0;
  LstmSpec lstm_hr, lstm_mix;
  ConvSpec c_hr, c_mix;
  std::vector<ConvSpec> convs;   // pyramid + second loop (models.py:111-126)
  int sc_from = 0;               // the shortcut branches off the output of convs[sc_from - 1]
  bool has_sc = false;
  ConvSpec sc;
  std::vector<ConvSpec> tail;    // models.py:132-136
  int dense_w, dense_b, flat, last_size, last_c;
};

namespace {

int add_var(wdg_critic* c, const std::string& name, std::initializer_list<int64_t> dims, bool trainable) {
  Var v;
  v.name = name;
  v.ndim = (int)dims.size();
  v.count = 1;
  int i = 0;
  for (int64_t d : dims) { v.dims[i++] = d; v.count *= d; }
  for (; i < 4; ++i) v.dims[i] = 1;
  v.offset = -1;
  v.trainable = trainable;
  c->vars.push_back(v);
  c->index[name] = (int)c->vars.size() - 1;
  return (int)c->vars.size() - 1;
}

std::string lw(int i, const char* leaf) { return "layer_with_weights-" + std::to_string(i) + "/" + leaf; }

void add_lstm(wdg_critic* c, LstmSpec& l, int idx, int cin, int F) {
  l.cin = cin; l.F = F;
  l.K = add_var(c, lw(idx, "cell/kernel"), {3, 3, cin, 4 * F}, true);
  l.R = add_var(c, lw(idx, "cell/recurrent_kernel"), {3, 3, F, 4 * F}, true);
  l.b = add_var(c, lw(idx, "cell/bias"), {4 * F}, true);
}
void add_sn_conv(wdg_critic* c, ConvSpec& e, int idx) {
  e.w = add_var(c, lw(idx, "layer/w"), {e.k, e.k, e.cin, e.cout}, true);
  e.b = add_var(c, lw(idx, "layer/layer/bias"), {e.cout}, true);
  e.u = add_var(c, lw(idx, "layer/sn_u"), {1, e.cout}, false);
}
void add_ln(wdg_critic* c, ConvSpec& e, int idx) {
  e.gamma = add_var(c, lw(idx, "gamma"), {e.cout}, true);
  e.beta = add_var(c, lw(idx, "beta"), {e.cout}, true);
}

// Walks the graph-building loops of models.py:93-140; variable indices follow the creation order of weighted layers,
// which is what the checkpoint's `layer_with_weights-N` keys count.
int build_plan(wdg_critic* c) {
  const int S = c->size, F = c->F;
  add_lstm(c, c->lstm_hr, 0, c->hr_ch, c->hr_ch);                    // :93
  add_lstm(c, c->lstm_mix, 1, c->lr_ch + c->hr_ch, F);               // :100-101
  c->c_hr = ConvSpec{3, 1, 1, c->hr_ch, F, S, S};                    // :94-96
  add_sn_conv(c, c->c_hr, 2);
  c->c_mix = ConvSpec{3, 1, 1, F, F, S, S};                          // :102-104
  add_sn_conv(c, c->c_mix, 3);
  add_ln(c, c->c_hr, 4);                                             // :97
  add_ln(c, c->c_mix, 5);                                            // :105
  int idx = 6, s = S, ch = 2 * F;
  auto conv7 = [&](int s_in, int c_in) {
    ConvSpec e{7, 3, 1, c_in, 2 * c_in, s_in, s_in + 2 >= 7 ? (s_in + 2 - 7) / 3 + 1 : 0};
    return e;
  };
  while (s >= 16) {                                                  // :111-116
    c->convs.push_back(conv7(s, ch));
    s = c->convs.back().size_out; ch *= 2;
  }
  c->sc_from = (int)c->convs.size();
  const int hs = s, hc = ch;
  int i = 0;
  while (s >= 4) {                                                   // :120-126
    ConvSpec e = conv7(s, ch);
    if (e.size_out < 1) return wdg_set_error("wdg_critic_create: invalid image size (a 7x7 stride-3 convolution has no output)");
    c->convs.push_back(e);
    s = e.size_out; ch *= 2; ++i;
  }
  // models.py:127 asks for `i > 1`, which no valid size reaches; the checkpoint the reference ships was written by a
  // revision whose condition held at i == 1 (discriminator.index holds shortcut_conv w[6,6,128,256]).
  c->has_sc = (i > 1) || (c->ckpt_topology && i >= 1);
  if (c->has_sc) {                                                   // tf_utils.py:15-32
    ConvSpec e{};
    if (s == 1) { e.k = hs; e.stride = 1; e.pad = 0; }
    else {
      e.stride = (int)std::ceil((2.0 + hs) / (s - 1));
      e.pad = (int)(std::ceil((e.stride * (s - 1) - hs) / 2.0) + 1 + 2);
      e.k = e.stride * (1 - s) + hs + 2 * e.pad;
    }
    e.cin = hc; e.cout = ch; e.size_in = hs; e.size_out = s;
    if (e.k < 1 || (hs + 2 * e.pad - e.k) / e.stride + 1 != s) return wdg_set_error("wdg_critic_create: shortcut geometry does not close");
    c->sc = e;
  }
  for (size_t n = 0; n < c->convs.size(); ++n) {
    const bool last = n + 1 == c->convs.size();
    add_sn_conv(c, c->convs[n], idx++);
    if (last && c->has_sc) add_sn_conv(c, c->sc, idx++);
    add_ln(c, c->convs[n], idx++);
    if (last && c->has_sc) add_ln(c, c->sc, idx++);
  }
  while (s > 2) {                                                    // :132-136
    ConvSpec e{3, 2, 0, ch, 2 * ch, s, (s - 3) / 2 + 1};
    add_sn_conv(c, e, idx);
    add_ln(c, e, idx + 1);
    idx += 2;
    c->tail.push_back(e);
    s = e.size_out; ch *= 2;
  }
  c->last_size = s; c->last_c = ch; c->flat = s * s * ch;
  c->dense_w = add_var(c, lw(idx, "layer/kernel"), {c->flat, 1}, true);   // :137-139
  c->dense_b = add_var(c, lw(idx, "layer/bias"), {1}, true);
  int64_t off = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (Var& v : c->vars)
      if (v.trainable == (pass == 0)) { v.offset = off; off += (v.count + 63) / 64 * 64; }
    if (pass == 0) c->n_train = off;
  }
  c->n_total = off;
  return 0;
}

// ---------------------------------------------------------------------------------------------- memory layout
struct Arena {
  uintptr_t base;
  size_t off = 0;
  explicit Arena(void* p) : base((uintptr_t)p) {}
  float* take(long long n_floats) {
    const size_t b = ((size_t)n_floats * 4 + 255) & ~(size_t)255;
    float* p = (float*)(base + off);
    off += b;
    return p;
  }
};

struct LstmBufs { float *xt, *hs, *cs, *gates, *packed, *hseq; };
struct ConvBufs { float *act, *mean, *inv, *y; };
struct FwdBufs {
  float *snap, *mix_in, *x, *sum;
  LstmBufs lhr, lmix;
  ConvBufs chr, cmix, sc;
  std::vector<ConvBufs> convs, tail;
};

bool lstm16_fused(int F) {
  return F == 16 && wdg_train_get_precision() != 0 && !getenv("WDG_NO_LSTM16");
}

void lay_lstm(Arena& a, LstmBufs& b, const LstmSpec& l, long long npx) {
  b.xt = a.take(npx * l.cin);
  b.hs = a.take(npx * l.F);
  b.cs = a.take(npx * l.F);
  b.gates = a.take(npx * 4 * l.F);
  b.packed = lstm16_fused(l.F) ? a.take(WDG_LSTM16_PACK_FLOATS) : nullptr;
  b.hseq = a.take(npx * l.F);
}
void lay_conv(Arena& a, ConvBufs& b, const ConvSpec& e, long long N, bool own_y) {
  const long long rows = N * e.size_out * e.size_out;
  b.act = a.take(rows * e.cout);
  b.mean = a.take(rows);
  b.inv = a.take(rows);
  b.y = own_y ? a.take(rows * e.cout) : nullptr;
}
// The same sequence of takes lays out the context for forward (which fills it) and backward (which reads it).
void lay_forward(const wdg_critic* c, Arena& a, int B, int T, int training, FwdBufs& f) {
  const long long N = (long long)B * T, npx = N * c->size * c->size;
  f.snap = training ? a.take(c->n_train) : nullptr;
  f.mix_in = a.take(npx * (c->lr_ch + c->hr_ch));
  lay_lstm(a, f.lhr, c->lstm_hr, npx);
  lay_conv(a, f.chr, c->c_hr, N, false);
  lay_lstm(a, f.lmix, c->lstm_mix, npx);
  lay_conv(a, f.cmix, c->c_mix, N, false);
  f.x = a.take(npx * 2 * c->F);
  f.convs.resize(c->convs.size());
  for (size_t i = 0; i < c->convs.size(); ++i) lay_conv(a, f.convs[i], c->convs[i], N, true);
  f.sum = nullptr;
  if (c->has_sc) {
    lay_conv(a, f.sc, c->sc, N, true);
    f.sum = a.take(N * c->sc.size_out * c->sc.size_out * c->sc.cout);
  }
  f.tail.resize(c->tail.size());
  for (size_t i = 0; i < c->tail.size(); ++i) lay_conv(a, f.tail[i], c->tail[i], N, true);
}

struct Geo { int v[16]; };
Geo conv_geo(long long N, const ConvSpec& e, int x_cs, int x_co, int y_cs, int y_co) {
  Geo g{{(int)N, e.size_in, e.size_in, e.cin, e.k, e.k, e.cout, e.stride, e.pad, e.pad, e.size_out, e.size_out, x_cs, x_co, y_cs, y_co}};
  return g;
}
Geo lstm_geo(long long N, int S, int cin, int cout) {
  Geo g{{(int)N, S, S, cin, 3, 3, cout, 1, 1, 1, S, S, cin, 0, cout, 0}};
  return g;
}

// scratch any single kernel of forward / backward may ask for (bytes), in every arithmetic mode
size_t op_scratch_bytes(const wdg_critic* c, int B, int T) {
  const long long N = (long long)B * T;
  size_t m = 1024 * 8;
  auto up = [&](size_t b) { if (b > m) m = b; };
  auto colsum = [&](int C) { up((size_t)512 * (C > 32 ? C : 32) * 4); };
  auto ln_bwd = [&](long long rows, int C) { up(((size_t)rows * C + (size_t)512 * (C > 32 ? C : 32)) * 4); };
  auto sn = [&](const ConvSpec& e) { up(((size_t)e.k * e.k * e.cin + 64 * (size_t)e.cout + 4) * 4); };
  auto wgrad = [&](const Geo& g) {      // every arithmetic mode: a buffer sized now stays valid if the caller switches precision
    for (int mode = 0; mode < 3; ++mode) {
      size_t b = 0;
      wdg_wgrad_scratch_for_mode(g.v, mode, &b, nullptr);
      up(b);
    }
  };
  auto conv = [&](const ConvSpec& e) {
    sn(e); colsum(e.cout);
    ln_bwd(N * e.size_out * e.size_out, e.cout);
    wgrad(conv_geo(N, e, e.cin, 0, e.cout, 0));
  };
  conv(c->c_hr); conv(c->c_mix);
  for (const ConvSpec& e : c->convs) conv(e);
  if (c->has_sc) conv(c->sc);
  for (const ConvSpec& e : c->tail) conv(e);
  for (const LstmSpec* l : {&c->lstm_hr, &c->lstm_mix}) {
    colsum(4 * l->F);
    wgrad(lstm_geo(N, c->size, l->cin, 4 * l->F));
    wgrad(lstm_geo(N, c->size, l->F, 4 * l->F));
  }
  return (m + 255) & ~(size_t)255;
}

#define RUN(call)                 \
  do {                            \
    if (int _rc = (call)) return _rc; \
  } while (0)

int sn_all(wdg_critic* c, float* vars, void* scratch, void* st) {
  auto one = [&](const ConvSpec& e) {
    return wdg_sn_update(vars + c->vars[e.w].offset, vars + c->vars[e.u].offset, e.k * e.k * e.cin, e.cout, scratch, st);
  };
  RUN(one(c->c_hr));
  RUN(one(c->c_mix));
  for (const ConvSpec& e : c->convs) RUN(one(e));
  if (c->has_sc) RUN(one(c->sc));
  for (const ConvSpec& e : c->tail) RUN(one(e));
  return 0;
}

// ConvLSTM2D(F, 3x3, same, return_sequences=True): x [B*T,S,S,cin] batch-major -> b.hseq [B*T,S,S,F] batch-major.
// The input convolution (+ bias) of all timesteps is one GEMM over the T*B images; only the recurrent part is serial.
int lstm_forward(const wdg_critic* c, const LstmSpec& l, LstmBufs& b, const float* W, const float* x, int B, int T, void* st) {
  const int S = c->size, F = l.F;
  const long long img = (long long)S * S, step = B * img;
  const float *K = W + c->vars[l.K].offset, *R = W + c->vars[l.R].offset, *bias = W + c->vars[l.b].offset;
  RUN(wdg_transpose01(x, b.xt, B, T, img * l.cin, st));
  const Geo gx = lstm_geo((long long)T * B, S, l.cin, 4 * F);
  RUN(wdg_conv2d_fwd(b.xt, K, bias, b.gates, gx.v, 0, st));
  const bool small = F == 1 || F == 2 || F == 4, fused = b.packed != nullptr;
  if (fused) RUN(wdg_lstm16_pack(R, b.packed, st));
  const Geo gh = lstm_geo(B, S, F, 4 * F);
  for (int t = 0; t < T; ++t) {
    float *z = b.gates + t * step * 4 * F, *ct = b.cs + t * step * F, *ht = b.hs + t * step * F;
    const float *cp = t ? b.cs + (t - 1) * step * F : nullptr, *hp = t ? b.hs + (t - 1) * step * F : nullptr;
    if (small) {
      RUN(wdg_lstm_small_fwd(z, hp, R, cp, ct, ht, B, S, S, F, st));
    } else if (fused) {
      if (t == 0) {
        RUN(wdg_lstm_gates_fwd(z, nullptr, ct, ht, step, F, st));
        RUN(wdg_round_tf32(ht, step * F, st));
      } else {
        RUN(wdg_lstm16_fwd_step(z, hp, b.packed, cp, ct, ht, B, S, S, st));
      }
    } else {
      if (t) RUN(wdg_conv2d_fwd(hp, R, nullptr, z, gh.v, 1, st));
      RUN(wdg_lstm_gates_fwd(z, cp, ct, ht, step, F, st));
    }
  }
  RUN(wdg_transpose01(b.hs, b.hseq, T, B, img * F, st));
  return 0;
}

// Backward of the cell: dh_seq [B*T,S,S,F] batch-major.  Leaves dz_t in b.gates.  dK/dR/db (NULL: skipped) receive the
// weight gradients, *dx_out (if asked for) the batch-major input gradient [B*T,S,S,cin] taken from `tmp`.
int lstm_backward(const wdg_critic* c, const LstmSpec& l, LstmBufs& b, const float* W, const float* dh_seq, int B, int T,
                  float* dK, float* dR, float* db, float** dx_out, Arena& tmp, void* ops, void* st) {
  const int S = c->size, F = l.F;
  const long long img = (long long)S * S, step = B * img;
  const float *K = W + c->vars[l.K].offset, *R = W + c->vars[l.R].offset;
  float* dhs = tmp.take(T * step * F);
  float* dc = tmp.take(step * F);
  float* dh_rec = tmp.take(step * F);
  RUN(wdg_transpose01(dh_seq, dhs, B, T, img * F, st));
  if (cudaMemsetAsync(dc, 0, (size_t)step * F * 4, (cudaStream_t)st) != cudaSuccess) return wdg_set_error("wdg_critic: cudaMemsetAsync failed");
  const bool small = F == 1 || F == 2 || F == 4, fused = b.packed != nullptr;
  const Geo gh = lstm_geo(B, S, F, 4 * F);
  auto Z = [&](int t) { return b.gates + t * step * 4 * F; };
  auto Cs = [&](int t) { return t >= 0 ? b.cs + t * step * F : (float*)nullptr; };
  if (fused) {
    // step T-1 has no recurrent gradient; every earlier step is one launch: recurrent backward-data of dz_{t+1}
    // (tcgen05) with the gate backward of step t in its epilogue
    RUN(wdg_lstm_gates_bwd(Z(T - 1), Cs(T - 2), Cs(T - 1), dhs + (T - 1) * step * F, nullptr, dc, step, F, st));
    RUN(wdg_round_tf32(Z(T - 1), step * 4 * F, st));
    for (int t = T - 2; t >= 0; --t)
      RUN(wdg_lstm16_bwd_step(Z(t + 1), b.packed, Z(t), Cs(t - 1), Cs(t), dhs + t * step * F, dc, B, S, S, st));
  } else {
    for (int t = T - 1; t >= 0; --t) {
      RUN(wdg_lstm_gates_bwd(Z(t), Cs(t - 1), Cs(t), dhs + t * step * F, t < T - 1 ? dh_rec : nullptr, dc, step, F, st));
      if (t > 0) {
        if (small) RUN(wdg_lstm_small_bwd_data(Z(t), R, dh_rec, B, S, S, F, st));
        else RUN(wdg_conv2d_bwd_data(Z(t), R, dh_rec, gh.v, 0, st));
      }
    }
  }
  // everything that only needs all dz_t: one GEMM each over the T*B images
  const Geo gx = lstm_geo((long long)T * B, S, l.cin, 4 * F);
  if (dK) {
    RUN(wdg_conv2d_bwd_weight(b.xt, b.gates, dK, gx.v, ops, 0, st));
    if (T > 1) {
      const Geo gr = lstm_geo((long long)(T - 1) * B, S, F, 4 * F);
      RUN(wdg_conv2d_bwd_weight(b.hs, Z(1), dR, gr.v, ops, 0, st));
    } else if (cudaMemsetAsync(dR, 0, (size_t)c->vars[l.R].count * 4, (cudaStream_t)st) != cudaSuccess) {
      return wdg_set_error("wdg_critic: cudaMemsetAsync failed");
    }
    RUN(wdg_colsum(0, b.gates, 4 * F, 0, b.gates, 4 * F, 0, T * step, 4 * F, db, ops, 0, st));
  }
  if (dx_out) {
    float* dxt = tmp.take(T * step * l.cin);
    float* dx = tmp.take(T * step * l.cin);
    RUN(wdg_conv2d_bwd_data(b.gates, K, dxt, gx.v, 0, st));
    RUN(wdg_transpose01(dxt, dx, T, B, img * l.cin, st));
    *dx_out = dx;
  }
  return 0;
}

// TimeDistributed(SN(Conv2D(cout, k, strides, activation=LeakyReLU(0.2)))) + LayerNormalization; the LayerNorm writes into
// channels [y_co, y_co + cout) of a buffer with pixel pitch y_cs.
int block_forward(const wdg_critic* c, const ConvSpec& e, ConvBufs& b, const float* W, const float* x, int x_cs, long long N,
                  float* y, int y_cs, int y_co, void* st) {
  const Geo g = conv_geo(N, e, x_cs, 0, e.cout, 0);
  RUN(wdg_conv2d_fwd_act(x, W + c->vars[e.w].offset, W + c->vars[e.b].offset, b.act, g.v, 0, kAlpha, st));
  RUN(wdg_ln_fwd(b.act, y, y_cs, y_co, W + c->vars[e.gamma].offset, W + c->vars[e.beta].offset, b.mean, b.inv,
                 N * e.size_out * e.size_out, e.cout, kLnEps, st));
  return 0;
}

// Backward of that block.  dy: channels [dy_co, ..) of a buffer with pitch dy_cs.  G: gradient buffer (variables'
// layout) or NULL (only the input gradient is wanted); junk: >= 2 * cout floats.  dx: [N,size_in,size_in,cin] dense.
int block_backward(const wdg_critic* c, const ConvSpec& e, ConvBufs& b, const float* W, const float* dy, int dy_cs, int dy_co,
                   const float* x, int x_cs, long long N, float* G, float* junk, float* dx, int accumulate_dx, Arena& tmp, void* ops,
                   void* st) {
  const long long rows = N * e.size_out * e.size_out;
  float* da = tmp.take(rows * e.cout);
  float* dgamma = G ? G + c->vars[e.gamma].offset : junk;
  float* dbeta = G ? G + c->vars[e.beta].offset : junk + e.cout;
  // the LeakyReLU backward of the convolution's activation is folded into the LayerNorm backward (act_alpha)
  RUN(wdg_ln_bwd(dy, dy_cs, dy_co, b.act, W + c->vars[e.gamma].offset, b.mean, b.inv, da, dgamma, dbeta, rows, e.cout, kAlpha, ops, st));
  if (G) {
    RUN(wdg_colsum(0, da, e.cout, 0, da, e.cout, 0, rows, e.cout, G + c->vars[e.b].offset, ops, 0, st));
    const Geo g = conv_geo(N, e, x_cs, 0, e.cout, 0);
    RUN(wdg_conv2d_bwd_weight(x, da, G + c->vars[e.w].offset, g.v, ops, 0, st));
  }
  if (dx) {
    const Geo g = conv_geo(N, e, e.cin, 0, e.cout, 0);
    RUN(wdg_conv2d_bwd_data(da, W + c->vars[e.w].offset, dx, g.v, accumulate_dx, st));
  }
  return 0;
}

int critic_backward(wdg_critic* c, const float* vars, void* context, int B, int T, int training, const float* dscore, float* G,
                    float* d_high_res, Arena& tmp, void* st) {
  const long long N = (long long)B * T, npx = N * c->size * c->size;
  const int F = c->F, cl = c->lr_ch, chn = c->hr_ch;
  Arena ctx(context);
  FwdBufs f;
  lay_forward(c, ctx, B, T, training, f);
  const float* W = training ? f.snap : vars;
  void* ops = tmp.take((long long)(op_scratch_bytes(c, B, T) / 4));
  int maxc = 2 * F;
  for (const ConvSpec& e : c->convs) maxc = e.cout > maxc ? e.cout : maxc;
  for (const ConvSpec& e : c->tail) maxc = e.cout > maxc ? e.cout : maxc;
  float* junk = tmp.take(2 * maxc + c->flat + 64);
  // activation that feeds the Dense layer
  const float* flat_act = !c->tail.empty() ? f.tail.back().y : (c->has_sc ? f.sum : (!c->convs.empty() ? f.convs.back().y : f.x));
  float* d = tmp.take(N * c->flat);
  RUN(wdg_dense_mean_bwd(dscore, flat_act, W + c->vars[c->dense_w].offset, d, G ? G + c->vars[c->dense_w].offset : junk + 2 * maxc,
                         G ? G + c->vars[c->dense_b].offset : junk + 2 * maxc + c->flat, B, T, c->flat, st));
  const int nA = (int)c->convs.size();
  auto input_of = [&](const std::vector<ConvSpec>& L, std::vector<ConvBufs>& Bf, int i, const float* first) {
    return i ? (const float*)Bf[i - 1].y : first;
  };
  const float* stageA_out = c->has_sc ? f.sum : (nA ? f.convs[nA - 1].y : f.x);
  for (int i = (int)c->tail.size() - 1; i >= 0; --i) {
    const ConvSpec& e = c->tail[i];
    float* dx = tmp.take(N * e.size_in * e.size_in * e.cin);
    RUN(block_backward(c, e, f.tail[i], W, d, e.cout, 0, input_of(c->tail, f.tail, i, stageA_out), e.cin, N, G, junk, dx, 0, tmp, ops, st));
    d = dx;
  }
  const float* d_top = d;    // gradient of the (summed) output of the two 7x7 loops
  for (int i = nA - 1; i >= 0; --i) {
    const ConvSpec& e = c->convs[i];
    float* dx = tmp.take(N * e.size_in * e.size_in * e.cin);
    RUN(block_backward(c, e, f.convs[i], W, d, e.cout, 0, input_of(c->convs, f.convs, i, f.x), e.cin, N, G, junk, dx, 0, tmp, ops, st));
    d = dx;
    if (c->has_sc && i == c->sc_from)      // d is now the main path's gradient of the shortcut's source: add the branch
      RUN(block_backward(c, c->sc, f.sc, W, d_top, c->sc.cout, 0, input_of(c->convs, f.convs, c->sc_from, f.x), c->sc.cin, N, G, junk,
                         d, 1, tmp, ops, st));
  }
  // d: [N,S,S,2F] gradient of Concatenate([hr, mix]) (models.py:108)
  float *dx_hr = nullptr, *dx_mix = nullptr;
  {
    float* dh = tmp.take(npx * c->lstm_hr.F);
    RUN(block_backward(c, c->c_hr, f.chr, W, d, 2 * F, 0, f.lhr.hseq, c->lstm_hr.F, N, G, junk, dh, 0, tmp, ops, st));
    const LstmSpec& l = c->lstm_hr;
    RUN(lstm_backward(c, l, f.lhr, W, dh, B, T, G ? G + c->vars[l.K].offset : nullptr, G ? G + c->vars[l.R].offset : nullptr,
                      G ? G + c->vars[l.b].offset : nullptr, d_high_res ? &dx_hr : nullptr, tmp, ops, st));
  }
  {
    float* dh = tmp.take(npx * c->lstm_mix.F);
    RUN(block_backward(c, c->c_mix, f.cmix, W, d, 2 * F, F, f.lmix.hseq, c->lstm_mix.F, N, G, junk, dh, 0, tmp, ops, st));
    const LstmSpec& l = c->lstm_mix;
    RUN(lstm_backward(c, l, f.lmix, W, dh, B, T, G ? G + c->vars[l.K].offset : nullptr, G ? G + c->vars[l.R].offset : nullptr,
                      G ? G + c->vars[l.b].offset : nullptr, d_high_res ? &dx_mix : nullptr, tmp, ops, st));
  }
  if (d_high_res)   // high_res feeds the hr cell directly and channels [cl, cl+ch) of the mixed cell's input
    RUN(wdg_axpby(d_high_res, chn, 0, dx_hr, chn, 0, 1.f, dx_mix, cl + chn, cl, 1.f, npx, chn, 0, st));
  return 0;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int wdg_critic_create(wdg_critic** out, int low_res_size, int high_res_size, int low_res_channels, int high_res_channels,
                                 int n_timesteps, int feature_channels, int ckpt_topology) {
  if (!out) return wdg_set_error("wdg_critic_create: null out");
  if (low_res_size != high_res_size)   // models.py:89-91
    return wdg_set_error("The discriminator assumes that the low res and high res images have the same size.");
  if (high_res_size < 16 || low_res_channels < 1 || high_res_channels < 1 || feature_channels < 1 || n_timesteps < 1)
    return wdg_set_error("wdg_critic_create: bad hyper-parameter");
  wdg_critic* c = new wdg_critic();
  c->size = high_res_size; c->lr_ch = low_res_channels; c->hr_ch = high_res_channels; c->T = n_timesteps; c->F = feature_channels;
  c->ckpt_topology = ckpt_topology ? 1 : 0;
  if (int rc = build_plan(c)) { delete c; return rc; }
  *out = c;
  return 0;
}
extern "C" void wdg_critic_destroy(wdg_critic* c) { delete c; }

extern "C" int wdg_critic_num_weights(const wdg_critic* c) { return c ? (int)c->vars.size() : 0; }
extern "C" int wdg_critic_weight_info(const wdg_critic* c, int index, const char** name, int64_t* dims, int* ndim, int64_t* offset,
                                      int* trainable) {
  if (!c || index < 0 || index >= (int)c->vars.size()) return wdg_set_error("wdg_critic_weight_info: index out of range");
  const Var& v = c->vars[index];
  if (name) *name = v.name.c_str();
  if (dims) for (int i = 0; i < 4; ++i) dims[i] = v.dims[i];
  if (ndim) *ndim = v.ndim;
  if (offset) *offset = v.offset;
  if (trainable) *trainable = v.trainable ? 1 : 0;
  return 0;
}
extern "C" int64_t wdg_critic_num_floats(const wdg_critic* c) { return c ? c->n_total : 0; }
extern "C" int64_t wdg_critic_num_trainable_floats(const wdg_critic* c) { return c ? c->n_train : 0; }

extern "C" int wdg_critic_context_bytes(const wdg_critic* c, int B, int T, int training, size_t* bytes) {
  if (!c || !bytes || B < 1 || T < 1) return wdg_set_error("wdg_critic_context_bytes: bad argument");
  Arena a(nullptr);
  FwdBufs f;
  lay_forward(c, a, B, T, training, f);
  *bytes = a.off;
  return 0;
}

extern "C" int wdg_critic_scratch_bytes(const wdg_critic* c, int B, int T, size_t* bytes) {
  if (!c || !bytes || B < 1 || T < 1) return wdg_set_error("wdg_critic_scratch_bytes: bad argument");
  // the backward's temporaries, walked with the allocation sequence critic_backward itself uses
  const long long N = (long long)B * T, npx = N * c->size * c->size;
  Arena a(nullptr);
  a.take((long long)(op_scratch_bytes(c, B, T) / 4));
  int maxc = 2 * c->F;
  for (const ConvSpec& e : c->convs) maxc = e.cout > maxc ? e.cout : maxc;
  for (const ConvSpec& e : c->tail) maxc = e.cout > maxc ? e.cout : maxc;
  a.take(2 * maxc + c->flat + 64);
  a.take(N * c->flat);
  auto block = [&](const ConvSpec& e, bool own_dx) {
    if (own_dx) a.take(N * e.size_in * e.size_in * e.cin);
    a.take(N * e.size_out * e.size_out * e.cout);
  };
  for (const ConvSpec& e : c->tail) block(e, true);
  for (const ConvSpec& e : c->convs) block(e, true);
  if (c->has_sc) block(c->sc, false);
  for (const LstmSpec* l : {&c->lstm_hr, &c->lstm_mix}) {
    a.take(npx * l->F);                       // dh
    a.take(npx * c->F);                       // da of the 3x3 convolution
    a.take(npx * l->F);                       // dhs
    a.take(npx / T * l->F); a.take(npx / T * l->F);   // dc, dh_rec
    a.take(npx * l->cin); a.take(npx * l->cin);       // dxt, dx
  }
  *bytes = a.off + 4096;
  return 0;
}

extern "C" int wdg_critic_sn_update(wdg_critic* c, float* vars_dev, void* scratch_dev, size_t scratch_bytes, void* stream) {
  if (!c || !vars_dev || !scratch_dev) return wdg_set_error("wdg_critic_sn_update: null argument");
  if (scratch_bytes < op_scratch_bytes(c, 1, 1)) return wdg_set_error("wdg_critic_sn_update: scratch too small");
  return sn_all(c, vars_dev, scratch_dev, stream);
}

extern "C" int wdg_critic_forward(wdg_critic* c, float* vars_dev, const float* low_res_dev, const float* high_res_dev, float* score_dev,
                                  int B, int T, int training, void* context_dev, size_t context_bytes, void* scratch_dev,
                                  size_t scratch_bytes, void* stream) {
  if (!c || !vars_dev || !low_res_dev || !high_res_dev || !score_dev || !context_dev || !scratch_dev || B < 1 || T < 1)
    return wdg_set_error("wdg_critic_forward: bad argument");
  if (((uintptr_t)context_dev | (uintptr_t)scratch_dev | (uintptr_t)vars_dev) & 255)
    return wdg_set_error("wdg_critic_forward: vars / context / scratch must be 256-byte aligned");
  Arena ctx(context_dev);
  FwdBufs f;
  lay_forward(c, ctx, B, T, training, f);
  if (ctx.off > context_bytes) return wdg_set_error("wdg_critic_forward: context too small (wdg_critic_context_bytes)");
  if (scratch_bytes < op_scratch_bytes(c, B, T)) return wdg_set_error("wdg_critic_forward: scratch too small (wdg_critic_scratch_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const float* W = vars_dev;
  if (training) {   // TFA SpectralNormalization: one in-place power iteration per wrapped layer per training-mode call
    RUN(sn_all(c, vars_dev, scratch_dev, stream));
    CKT(cudaMemcpyAsync(f.snap, vars_dev, (size_t)c->n_train * 4, cudaMemcpyDeviceToDevice, st));
    W = f.snap;
  }
  const long long N = (long long)B * T, npx = N * c->size * c->size;
  const int F = c->F, cl = c->lr_ch, ch = c->hr_ch;
  RUN(wdg_axpby(f.mix_in, cl + ch, 0, low_res_dev, cl, 0, 1.f, nullptr, cl, 0, 0.f, npx, cl, 0, stream));        // models.py:100
  RUN(wdg_axpby(f.mix_in, cl + ch, cl, high_res_dev, ch, 0, 1.f, nullptr, ch, 0, 0.f, npx, ch, 0, stream));
  RUN(lstm_forward(c, c->lstm_hr, f.lhr, W, high_res_dev, B, T, stream));                                          // :93
  RUN(block_forward(c, c->c_hr, f.chr, W, f.lhr.hseq, c->lstm_hr.F, N, f.x, 2 * F, 0, stream));                    // :94-97
  RUN(lstm_forward(c, c->lstm_mix, f.lmix, W, f.mix_in, B, T, stream));                                            // :101
  RUN(block_forward(c, c->c_mix, f.cmix, W, f.lmix.hseq, c->lstm_mix.F, N, f.x, 2 * F, F, stream));                // :102-105, :108
  const float* cur = f.x;
  const float* sc_src = f.x;
  for (size_t i = 0; i < c->convs.size(); ++i) {                                                                   // :111-126
    const ConvSpec& e = c->convs[i];
    RUN(block_forward(c, e, f.convs[i], W, cur, e.cin, N, f.convs[i].y, e.cout, 0, stream));
    cur = f.convs[i].y;
    if ((int)i + 1 == c->sc_from) sc_src = cur;
  }
  if (c->has_sc) {                                                                                                 // :127-130
    const ConvSpec& e = c->sc;
    RUN(block_forward(c, e, f.sc, W, sc_src, e.cin, N, f.sc.y, e.cout, 0, stream));
    RUN(wdg_axpby(f.sum, e.cout, 0, cur, e.cout, 0, 1.f, f.sc.y, e.cout, 0, 1.f, N * e.size_out * e.size_out, e.cout, 0, stream));
    cur = f.sum;
  }
  for (size_t i = 0; i < c->tail.size(); ++i) {                                                                    // :132-136
    const ConvSpec& e = c->tail[i];
    RUN(block_forward(c, e, f.tail[i], W, cur, e.cin, N, f.tail[i].y, e.cout, 0, stream));
    cur = f.tail[i].y;
  }
  RUN(wdg_dense_mean_fwd(cur, W + c->vars[c->dense_w].offset, W + c->vars[c->dense_b].offset, score_dev, B, T, c->flat, stream));   // :137-140
  return 0;
}

extern "C" int wdg_critic_backward(wdg_critic* c, const float* vars_dev, void* context_dev, int B, int T, int training,
                                   const float* dscore_dev, float* grads_dev, float* d_high_res_dev, void* scratch_dev,
                                   size_t scratch_bytes, void* stream) {
  if (!c || !vars_dev || !context_dev || !dscore_dev || !scratch_dev || B < 1 || T < 1) return wdg_set_error("wdg_critic_backward: bad argument");
  if (!grads_dev && !d_high_res_dev) return wdg_set_error("wdg_critic_backward: nothing to compute (grads_dev and d_high_res_dev are both NULL)");
  size_t need = 0;
  RUN(wdg_critic_scratch_bytes(c, B, T, &need));
  if (scratch_bytes < need) return wdg_set_error("wdg_critic_backward: scratch too small (wdg_critic_scratch_bytes)");
  if (((uintptr_t)context_dev | (uintptr_t)scratch_dev) & 255) return wdg_set_error("wdg_critic_backward: context / scratch must be 256-byte aligned");
  Arena tmp(scratch_dev);
  RUN(critic_backward(c, vars_dev, context_dev, B, T, training, dscore_dev, grads_dev, d_high_res_dev, tmp, stream));
  if (tmp.off > scratch_bytes) return wdg_set_error("wdg_critic_backward: internal error: scratch accounting");
  return 0;
}

extern "C" int wdg_critic_backward_input(wdg_critic* c, const float* vars_dev, void* context_dev, int B, int T, int training,
                                         const float* dscore_dev, float* d_high_res_dev, void* scratch_dev, size_t scratch_bytes,
                                         void* stream) {
  return wdg_critic_backward(c, vars_dev, context_dev, B, T, training, dscore_dev, nullptr, d_high_res_dev, scratch_dev, scratch_bytes, stream);
}
extern "C" int wdg_critic_backward_weights(wdg_critic* c, const float* vars_dev, void* context_dev, int B, int T, int training,
                                           const float* dscore_dev, float* grads_dev, void* scratch_dev, size_t scratch_bytes,
                                           void* stream) {
  return wdg_critic_backward(c, vars_dev, context_dev, B, T, training, dscore_dev, grads_dev, nullptr, scratch_dev, scratch_bytes, stream);
}
