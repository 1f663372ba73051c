/*
 * wdg.h -- C ABI of libwdg.so, the B200 (sm_100a) implementation of the
 * wind-downscaling-gan generator forward hot path.
 *
 * The reference (OpheliaMiralles/wind-downscaling-gan) has no FFI: its boundary for
 * this path is the Keras object surface (`models.py:9-73`, `api.py:89-152`).  The
 * entry points below are what a ctypes binding added to the reference would call in
 * place of `make_generator(...)`, `generator.load_weights(...)` and
 * `gen.predict([tensor, noise])` -- see INTEGRATION.md for that binding.
 *
 * Conventions: plain C types only; every function returns 0 on success or a
 * non-zero code with a message available from wdg_last_error(); no exceptions
 * cross the ABI; tensors are channels-last float32 exactly as the reference
 * passes them ((B, T, S, S, C), `models.py:24-25`).  `*_dev` pointers are device
 * memory owned by the caller; `stream` is a cudaStream_t (NULL = default stream).
 * There is no CPU fallback: every call fails if no sm_100 device is present.
 */
#ifndef WDG_H
#define WDG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wdg_generator wdg_generator;

/* Last error message of the calling thread ("" if none). */
const char* wdg_last_error(void);

/* Library / device probe: returns 0 and fills sm_major/sm_minor/sm_count of device `device`. */
int wdg_device_info(int device, int* sm_major, int* sm_minor, int* sm_count);

/* Replaces make_generator(image_size, in_channels, noise_channels, out_channels, n_timesteps,
 * batch_size=None, feature_channels=128) -- models.py:9-17.  Asserts of models.py:19-20 are
 * returned as errors.  n_timesteps is the default sequence length; forward() takes T per call. */
int wdg_generator_create(wdg_generator** out, int image_size, int in_channels, int noise_channels,
                         int out_channels, int n_timesteps, int feature_channels);
void wdg_generator_destroy(wdg_generator* g);

/* Number of weight tensors, and name/shape of the i-th one.  Names are the reference
 * checkpoint's variable names (weights-55.ckpt/generator.index), e.g.
 * "layer_with_weights-0/layer/w"; layouts are the reference's (Conv2D HWIO, Conv2DTranspose
 * (kh, kw, out, in)).  dims must have room for 4 entries. */
int wdg_generator_num_weights(const wdg_generator* g);
int wdg_generator_weight_info(const wdg_generator* g, int index, const char** name, int64_t* dims, int* ndim);

/* Replaces generator.load_weights / set_weights (ganbase.py:136-140): copies one fp32 HOST
 * tensor, checked against the expected shape, into the handle. */
int wdg_generator_set_weight(wdg_generator* g, const char* name, const float* host_data, const int64_t* dims,
                             int ndim);
/* Reads one tensor back (fp32 HOST), e.g. for save_weights (ganbase.py:132-134). */
int wdg_generator_get_weight(const wdg_generator* g, const char* name, float* host_data, int64_t count);

/* Re-packs the fp32 weights into the kernels' device layouts (bf16, K-major, gate-interleaved,
 * BatchNorm folded to scale/shift).  Must be called after the last set_weight and before forward. */
int wdg_generator_finalize(wdg_generator* g);

/* Operand precision of the GEMM stages (default WDG_PREC_BF16).  The reference generator is fp32
 * (models.py:24-73); north_star's tolerances are rel-L2 <= 1e-3 for fp32/TF32 and <= 1e-2 for bf16:
 *   WDG_PREC_BF16: bf16 activations/weights in HBM, tcgen05.mma kind::f16, fp32 accumulation and epilogue math;
 *   WDG_PREC_TF32: fp32 activations/weights rounded to nearest tf32 by their producer, tcgen05.mma kind::tf32,
 *                  fp32 accumulation, fp32 cell state, output 3x3 convolution in full fp32.
 * Changing it invalidates finalize() and the bound plan (workspace_bytes depends on it). */
#define WDG_PREC_BF16 0
#define WDG_PREC_TF32 1
int wdg_generator_set_precision(wdg_generator* g, int mode);
int wdg_generator_get_precision(const wdg_generator* g);

/* Workspace (activation buffers) needed for a forward of B sequences x T timesteps. */
int wdg_generator_workspace_bytes(const wdg_generator* g, int B, int T, size_t* bytes);
/* Binds a caller-owned device workspace of at least that size and builds the TMA descriptors /
 * launch plan for (B, T).  The workspace is zero-filled on `stream`. */
int wdg_generator_bind(wdg_generator* g, int B, int T, void* workspace_dev, size_t bytes, void* stream);

/* Replaces gen.predict([image, noise]) / generator([image, noise], training=False) -- api.py:137.
 * image_dev (B,T,S,S,Cin) fp32, noise_dev (B,T,S,S,Cnoise) fp32 -> out_dev (B,T,S,S,Cout) fp32,
 * all device pointers; asynchronous on `stream`.  (B, T) must equal the bound plan. */
int wdg_generator_forward(wdg_generator* g, const float* image_dev, const float* noise_dev, float* out_dev,
                          void* stream);

/* forward() with the noise drawn INSIDE the input-packing kernel (what the reference does per group with the TF
 * generator, api.py:136 / data_generator.py:327-335): noise element i = stddev * N(0,1) from Philox counter block
 * noise_offset + i/4 under key noise_seed -- bit-identical to wdg_noise_normal() of the whole (B,T,S,S,Cnoise) tensor
 * followed by forward(), without that tensor's HBM round trip.  noise_scratch_dev: NULL for the reference's 3 + 20
 * channels; other channel counts draw the tensor first and need B*T*S*S*Cnoise floats of scratch. */
int wdg_generator_forward_gen_noise(wdg_generator* g, const float* image_dev, float noise_std, uint64_t noise_seed,
                                    uint64_t noise_offset, float* out_dev, void* noise_scratch_dev, void* stream);

/* Same call with HOST buffers: copies inputs host->device, runs forward, copies the result
 * back and synchronises.  io_dev is a caller-owned device staging buffer of at least
 * wdg_generator_io_bytes() bytes. */
int wdg_generator_io_bytes(const wdg_generator* g, int B, int T, size_t* bytes);
int wdg_generator_predict_host(wdg_generator* g, const float* image_host, const float* noise_host,
                               float* out_host, void* io_dev, void* stream);

/* As predict_host, but the noise (B,T,S,S,Cnoise) ~ N(0, noise_std^2) is generated on the device (wdg_noise_normal with
 * key noise_seed, starting at counter block noise_offset), as the reference's FlexibleNoiseGenerator does with the
 * TensorFlow generator (api.py:136): only the image crosses PCIe. */
int wdg_generator_predict_host_gen_noise(wdg_generator* g, const float* image_host, float noise_std, uint64_t noise_seed,
                                         uint64_t noise_offset, float* out_host, void* io_dev, void* stream);

/* The piece sizes (sequences) predict_host / predict_host_gen_noise cut a batch of B sequences into: H2D of piece i+1,
 * forward of piece i and D2H of piece i-1 overlap on three streams.  host_noise != 0: the H2D-bound schedule of
 * predict_host (uniform chunks, short last piece); 0: the compute-bound one of predict_host_gen_noise (small head and tail,
 * doubling pieces in between).  Returns the number of pieces (0: the batch is run in one piece) and writes up to
 * `capacity` of them; -1 on a bad argument.  Host arithmetic only. */
int wdg_generator_pipeline_schedule(int B, int host_noise, int* pieces, int capacity);

/* Number of kernels one forward() launches for the bound plan (bench.py's gpu_launches). */
int wdg_generator_launches_per_forward(const wdg_generator* g);

/* Per-stage device timing of forward() with CUDA events recorded on the caller's stream
 * (bench.py's roofline).  Stages, in launch order: 0 pack_input, 1 conv 8x8 s2, 2 conv 4x4 s2,
 * 3 ConvLSTM (T launches), 4 conv 3x3, 5 convT 2x2 s2, 6 border lines + border-correction GEMM, 7 fused bilinear x2 + convT 5x5, 8 conv 3x3 out. */
#define WDG_NUM_STAGES 9
int wdg_generator_profile(wdg_generator* g, int enable);
int wdg_generator_stage_ms(wdg_generator* g, float* ms, int n);

/* Debug / parity hooks: copies an intermediate activation of the last forward() to the host as
 * fp32.  which: 0 res_2 (N,S/2,S/2,128)  1 res_4 (N,S/4,S/4,128)  2 lstm h (N,S/4,S/4,128)
 * 3 g5 (N,S/4,S/4,64)  4 g7 (N,S/2,S/2,32)  5 g9 (N,S,S,16), N = B*T. */
int wdg_generator_debug_read(const wdg_generator* g, int which, float* host_out, int64_t count);

/* ---- Tiling driver of predict() (api.py:98-151), device side.  All pointers are device memory.
 * Patch order is the reference's: index = (ix * ny + iy) * ntimeseq + k (api.py:117-124); patch row p maps to
 * domain row sy+img-1-p, or img-p when sy == 0 (api.py:119). */

/* Scratch needed by wdg_gather_normalise (sequences of up to WDG_MAX_SEQ timesteps; api.py:22 uses 24). */
#define WDG_MAX_SEQ 32
int wdg_patch_scratch_bytes(int nx, int ny, int ntimeseq, int img, size_t* bytes);

/* Replaces api.py:117-129: slices u10/v10 (T_total,H,W) and elevation_km (H,W) into patches, computes
 * nanmean / nanstd over axes (0,1,2) of the (N,seq,img,img,3) stack -- i.e. per (patch column, channel), written to
 * mean_dev/std_dev as fp64 [img][3] -- and writes the normalised fp32 tensor out_dev (N,seq,img,img,3). */
int wdg_gather_normalise(const float* u10_dev, const float* v10_dev, const float* elev_km_dev, int T_total, int H,
                         int W, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny, int seq, int img,
                         double* mean_dev, double* std_dev, float* out_dev, void* scratch_dev, void* stream);

/* Same, with the nearest-neighbour regridding of api.py:31-43 folded into the gather: u10/v10 stay on their coarse
 * (uh, uw) grid and the DEM on its raster (eh, ew); *_row_map[H] / *_col_map[W] give, for every hi-res template row / col,
 * the source row / col (`.sel(method='nearest')`), and the DEM is divided by dem_divisor (1e3, api.py:96). */
int wdg_gather_normalise_regrid(const float* u10_coarse_dev, const float* v10_coarse_dev, int T_total, int uh, int uw,
                                const int* uv_row_map_dev, const int* uv_col_map_dev, const float* dem_dev, int eh, int ew,
                                const int* dem_row_map_dev, const int* dem_col_map_dev, float dem_divisor, int H, int W,
                                const int* starts_x_dev, int nx, const int* starts_y_dev, int ny, int seq, int img,
                                double* mean_dev, double* std_dev, float* out_dev, void* scratch_dev, void* stream);

/* Replaces api.py:140-151: crops `crop` pixels off every patch side and averages overlapping predictions.
 * pred_dev (N,seq,img,img,channels) fp32; rows_dev/cols_dev: sorted covered domain rows/cols;
 * out_dev (channels, ntimeseq*seq, nrows, ncols) fp32.  Contributions are summed in fp64 in patch order. */
int wdg_stitch(const float* pred_dev, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny, int ntimeseq,
               int seq, int img, int crop, int channels, const int* rows_dev, int nrows, const int* cols_dev, int ncols,
               float* out_dev, void* stream);

/* Same with the accumulation type of the mean stated.  The reference's mean is pandas' `groupby(level=...).mean()`
 * (api.py:150) = Cython group_mean: Kahan-compensated sum in order of appearance, divided by the count --
 * WDG_STITCH_F64: in float64, result cast to float32 (pandas==1.3.3, the reference's pin, upcasts float32 columns; this is
 * what wdg_stitch does); WDG_STITCH_F32: in float32 (pandas >= 1.5 keeps float32).  Both are bit-exact against real
 * pandas output (tests/golden/stitch_pandas.npz). */
#define WDG_STITCH_F64 0
#define WDG_STITCH_F32 1
int wdg_stitch_accum(const float* pred_dev, const int* starts_x_dev, int nx, const int* starts_y_dev, int ny, int ntimeseq,
                     int seq, int img, int crop, int channels, const int* rows_dev, int nrows, const int* cols_dev, int ncols,
                     float* out_dev, int accum, void* stream);

/* ---- On-device noise for FlexibleNoiseGenerator (data/data_generator.py:319-335): out[i] ~ N(0, stddev^2) from
 * Philox4x32-10 + Box-Muller; element 4j..4j+3 come from counter block `offset + j` under key `seed`, so a generator
 * advances `offset` by ceil(n/4) per call.  wdg_philox4x32_10 is the host-callable block function (known-answer tests). */
void wdg_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int wdg_noise_normal(float* out_dev, long long n, float stddev, uint64_t seed, uint64_t offset, void* stream);

/* Graph-capturable form: state_dev[0] = key, state_dev[1] = next counter block, both in device memory; the draw is
 * followed by a one-thread kernel that advances the counter by ceil(n/4), so a replayed CUDA graph draws fresh noise. */
int wdg_noise_normal_state(float* out_dev, long long n, float stddev, uint64_t* state_dev, void* stream);
int wdg_rng_advance(uint64_t* state_dev, uint64_t blocks, void* stream);
/* out[i] ~ U[0, 1) from the same device-resident stream (the interpolation weights of ganbase.py:30) */
int wdg_uniform_state(float* out_dev, long long n, uint64_t* state_dev, void* stream);

/* ---- building blocks of the WGAN training step (ganbase.py:21-94): critic forward/backward, training-mode
 * generator, optimiser.  Channels-last fp32 device tensors; `*_cs` / `*_co` = channel stride / offset of a tensor
 * inside a wider (concatenated) buffer.  geo[16] = {N, H, W, Ci, kh, kw, Co, stride, pad_top, pad_left, Ho, Wo,
 * x_cs, x_co, y_cs, y_co}; weights are HWIO (a Conv2DTranspose kernel (kh,kw,out,in) is the HWIO kernel of the
 * convolution it transposes, so its forward is wdg_conv2d_bwd_data). */
/* Arithmetic of the three convolution GEMMs below: 0 = fp32 on CUDA cores (default), 1 = tf32 operands, 2 = bf16
 * operands on the tcgen05 tensor cores (operands rounded to nearest in the loaders, fp32 accumulation in TMEM,
 * fp32 tensors in memory either way).  Process-wide setting. */
int wdg_train_set_precision(int mode);
int wdg_train_get_precision(void);
int wdg_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, const int* geo, int accumulate, void* stream);
/* same with the LeakyReLU(alpha) of the reference's `activation=` argument fused into the epilogue (alpha = 1: linear) */
int wdg_conv2d_fwd_act(const float* x, const float* w, const float* bias, float* y, const int* geo, int accumulate, float alpha,
                       void* stream);
int wdg_conv2d_bwd_data(const float* dy, const float* w, float* dx, const int* geo, int accumulate, void* stream);
int wdg_conv2d_bwd_weight_scratch(const int* geo, size_t* bytes, int* splits);
int wdg_conv2d_bwd_weight(const float* x, const float* dy, float* dw, const int* geo, void* scratch, int accumulate, void* stream);
/* out[C] (+)= column sums over R rows; mode 0: a, 1: a*b, 2: a*a; scratch >= 512*max(C,32) floats */
int wdg_colsum(int mode, const float* a, int a_cs, int a_co, const float* b, int b_cs, int b_co, long long R, int C, float* out,
               void* scratch, int accumulate, void* stream);
int wdg_leaky_relu_fwd(float* x, long long n, float alpha, void* stream);
int wdg_leaky_relu_bwd(float* dy, const float* y, long long n, float alpha, void* stream);
int wdg_axpby(float* out, int o_cs, int o_co, const float* x, int x_cs, int x_co, float a, const float* y, int y_cs, int y_co,
              float b, long long rows, int C, int accumulate, void* stream);
/* x = leaky(x + bias, alpha) in place (alpha = 1: linear); out[b][a][:] = in[a][b][:] */
int wdg_bias_act(float* x, int cs, int co, const float* bias, long long rows, int C, float alpha, void* stream);
int wdg_transpose01(const float* in, float* out, int A, int B, long long inner, void* stream);
/* Overlap-add for a stride-1 transposed convolution computed as a 1x1 GEMM into per-pixel tap columns
 * cols[N,H,W,(kh,kw,Co)]: out[n,iy,ix,o] = sum_{ky,kx} cols[n, iy+pad-ky, ix+pad-kx, (ky,kx,o)] */
int wdg_col2im(const float* cols, float* out, int N, int H, int W, int kh, int kw, int Co, int pad, int o_cs, int o_co, void* stream);
int wdg_lerp_batch(float* out, const float* real, const float* fake, const float* eps, long long per_sample, long long n, void* stream);
/* BatchNormalization (axis -1, eps, momentum): training mode uses batch statistics and updates the moving ones.
 * scratch: wdg_bn_train_fwd >= 512*max(C,32) + 2*C floats, wdg_bn_bwd_sums / wdg_bn_train_bwd >= 512*max(C,32) floats,
 * wdg_bn_infer >= C floats.  act_alpha of the backward entry points (BatchNorm and LayerNorm): when x is the output of a
 * LeakyReLU(act_alpha) -- the reference's conv -> LeakyReLU -> norm blocks -- that activation's backward is folded into
 * the dx they write (1 = x is not an activation output) */
int wdg_bn_train_fwd(const float* x, float* y, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                     float* save_mean, float* save_invstd, long long rows, int C, float eps, float momentum, void* scratch, void* stream);
/* Split forms for data-parallel (synchronised) BatchNorm: per-channel sums are all-reduced by the caller */
int wdg_bn_finalize_apply(const float* x, float* y, const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                          const float* s1, const float* s2, float* save_mean, float* save_invstd, long long rows_local,
                          long long rows_global, int C, float eps, float momentum, void* stream);
int wdg_bn_bwd_sums(const float* dy, const float* x, const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                    long long rows, int C, void* scratch, void* stream);
int wdg_bn_bwd_dx(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                  const float* dgamma, const float* dbeta, float* dx, long long rows_local, long long rows_global, int C,
                  float act_alpha, void* stream);
int wdg_bn_infer(const float* x, float* y, const float* gamma, const float* beta, const float* mean, const float* var,
                 long long rows, int C, float eps, void* scratch, void* stream);
int wdg_bn_train_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean, const float* save_invstd,
                     float* dx, float* dgamma, float* dbeta, long long rows, int C, float act_alpha, void* scratch, void* stream);
/* LayerNormalization over the channel axis (wdg_ln_bwd scratch >= rows*C + 512*max(C,32) floats) */
int wdg_ln_fwd(const float* x, float* y, int y_cs, int y_co, const float* gamma, const float* beta, float* save_mean,
               float* save_invstd, long long rows, int C, float eps, void* stream);
int wdg_ln_bwd(const float* dy, int dy_cs, int dy_co, const float* x, const float* gamma, const float* save_mean,
               const float* save_invstd, float* dx, float* dgamma, float* dbeta, long long rows, int C, float act_alpha,
               void* scratch, void* stream);
/* ConvLSTM2D gate math (gates i, f, c~, o; hard_sigmoid / tanh), one timestep */
int wdg_lstm_gates_fwd(float* z, const float* c_prev, float* c_out, float* h_out, long long rows, int F, void* stream);
/* dh_rec (may be NULL): recurrent part of dL/dh_t carried from step t+1, added to dh inside the kernel */
int wdg_lstm_gates_bwd(float* gates, const float* c_prev, const float* c_cur, const float* dh, const float* dh_rec, float* dc,
                       long long rows, int F, void* stream);
/* Cells with F in {1, 2, 4} filters (critic high-resolution branch): recurrent 3x3 convolution fused with the gate math
 * (z: x-conv + bias in, activated gates out; h_prev / c_prev NULL at t = 0; R = recurrent kernel [3][3][F][4F]) and
 * the recurrent backward-data stencil dh_rec = conv_bwd_data(dz, R). */
int wdg_lstm_small_fwd(float* z, const float* h_prev, const float* R, const float* c_prev, float* c_out, float* h_out,
                       int N, int H, int W, int F, void* stream);
int wdg_lstm_small_bwd_data(const float* dz, const float* R, float* dh_rec, int N, int H, int W, int F, void* stream);
/* Cells with F = 16 filters (critic mixed branch, models.py:100-101), tensor-core training modes: ONE launch per
 * timestep -- a TMA-fed tcgen05 (kind::tf32) recurrent convolution with the gate math in the TMEM epilogue.
 * wdg_lstm16_pack: R [3][3][16][64] -> `packed` (WDG_LSTM16_PACK_FLOATS floats: forward and backward-data operands,
 * rounded to tf32); call again whenever R changes.  wdg_lstm16_fwd_step (t >= 1): gates [N,H,W,64] holds the input
 * convolution + bias of step t and receives the activated gates; h_prev / c_prev of step t-1; writes c_out, h_out (h_out
 * rounded to tf32 -- it is only ever a GEMM operand; round h_0 of the unfused first step with wdg_round_tf32).
 * wdg_lstm16_bwd_step: dz_next = dz of step s+1 (as left in its gates buffer); gates_s = activated gates of step s, receives
 * dz_s; c_prev = c_{s-1} (NULL at s = 0), c_cur = c_s, dh = dL/dh_s from above, dc carried in place. */
#define WDG_LSTM16_PACK_FWD_FLOATS (64 * 9 * 32)
#define WDG_LSTM16_PACK_FLOATS (64 * 9 * 32 + 16 * 18 * 32)
int wdg_lstm16_pack(const float* R, float* packed, void* stream);
int wdg_round_tf32(float* x, long long n, void* stream);
int wdg_lstm16_fwd_step(float* gates, const float* h_prev, const float* packed, const float* c_prev, float* c_out, float* h_out,
                        int N, int H, int W, void* stream);
int wdg_lstm16_bwd_step(const float* dz_next, const float* packed, float* gates_s, const float* c_prev, const float* c_cur,
                        const float* dh, float* dc, int N, int H, int W, void* stream);
/* Cells with F = 128 filters (the generator's ConvLSTM2D, models.py:45), tensor-core training modes: the same fused forward
 * step (t >= 1) -- TMA-fed tcgen05 recurrent convolution (K = 9 x 128, N = 512) with the gate math in the TMEM epilogue.
 * wdg_lstm128_pack: R [3][3][128][512] -> packed (WDG_LSTM128_PACK_FLOATS floats, tf32).  gates [N,H,W,512]: input
 * convolution + bias of the step in, activated gates out (reference gate order i, f, c~, o); h_out rounded to tf32. */
#define WDG_LSTM128_PACK_FLOATS (512 * 36 * 32)
int wdg_lstm128_pack(const float* R, float* packed, void* stream);
int wdg_lstm128_fwd_step(float* gates, const float* h_prev, const float* packed, const float* c_prev, float* c_out, float* h_out,
                         int N, int H, int W, void* stream);
/* Generator block Concatenate([a, b]) -> UpSampling2D(2, bilinear) -> Conv2DTranspose(16, 5x5, same) + bias -> LeakyReLU(0.2)
 * (models.py:60-64) in ONE fused pass for the training forward (tensor-core modes): a [N,h,h,32], b [N,h,h,128] dense fp32
 * -> out [N,2h,2h,16] dense fp32 (pre-BatchNorm).  w is the layer's kernel [5][5][16][160]; the 4x4 phase weights and the
 * border-correction weights are composed from it on the device at every call (the optimizer changes it every step).
 * tf32 operands, fp32 accumulation.  workspace: wdg_upconv5x5_workspace_bytes(N, h), 1024-byte aligned. */
int wdg_upconv5x5_workspace_bytes(long long N, int h, size_t* bytes);
int wdg_upconv5x5_fwd(const float* a, const float* b, const float* w, const float* bias, float* out, long long N, int h,
                      void* workspace, size_t ws_bytes, void* stream);
/* UpSampling2D(2, bilinear) and its adjoint */
int wdg_upsample2x_fwd(const float* x, float* y, long long n_img, int h, int w, int C, void* stream);
int wdg_upsample2x_bwd(const float* dy, float* dx, long long n_img, int h, int w, int C, void* stream);
/* TimeDistributed(Dense(1)) + GlobalAveragePooling1D and their backward */
int wdg_dense_mean_fwd(const float* flat, const float* w, const float* bias, float* score, int B, int T, int D, void* stream);
int wdg_dense_mean_bwd(const float* dscore, const float* flat, const float* w, float* dflat, float* dw, float* dbias, int B, int T, int D, void* stream);
/* out[0] = scale * sum(a | a*b | a*a); scratch >= 1024 doubles */
int wdg_reduce(int mode, const float* a, const float* b, long long n, double scale, float* out, void* scratch, void* stream);
/* gradient-penalty norms (ganbase.py:36): out[b*C+c] = sqrt(sum over (T,H,W) of g^2) */
int wdg_gp_norm(const float* g, float* out, int B, long long per_sample_px, int C, void* stream);
/* Keras Adam step (epsilon outside the square root) and one TFA SpectralNormalization power iteration (in place;
 * scratch >= (R + 64*C + 4) floats) */
int wdg_adam(float* w, float* m, float* v, const float* g, long long n, float lr_t, float b1, float b2, float eps, void* stream);
/* Graph-capturable Adam: wdg_adam_lr increments the device step counter t and writes lr_t = lr sqrt(1-b2^t)/(1-b1^t);
 * wdg_adam_dev reads lr_t from device memory. */
int wdg_adam_lr(float* lr_t_dev, int* step_dev, float lr, float b1, float b2, void* stream);
int wdg_adam_dev(float* w, float* m, float* v, const float* g, long long n, const float* lr_t_dev, float b1, float b2, float eps,
                 void* stream);
int wdg_sn_update(float* w, float* u, int R, int C, void* scratch, void* stream);

/* ---- the WGAN critic as a handle (reference gan/models.py:76-142 `make_discriminator`, tf_utils.py:7-32; train step
 * ganbase.py:21-94).  What a reference-side binding calls in place of `discriminator([low_res, high_res], training=...)`,
 * `reg_tape.gradient(score, interpolated)` (ganbase.py:35: backward to the image only), `disc_tape.gradient(loss,
 * trainable_weights)` (:46: backward to the weights) and `gen_tape.gradient` through the critic (:60).
 *
 * wdg_critic_create takes make_discriminator's hyper-parameters (models.py:76-84; batch_size is not needed).  It only builds
 * the layer plan, so it works without a device; sizes differing -> the reference's NotImplementedError text (models.py:89-91).
 * ckpt_topology = 0: the graph the current reference code builds (the `i > 1` shortcut of models.py:127 never triggers);
 * ckpt_topology = 1: the graph the shipped weights-55.ckpt/discriminator.index was written from -- the same plus the shortcut
 * convolution + LayerNorm + add of tf_utils.py:15-32 around the last 7x7 stage (at 96 px: 6x6, stride 11, pad 4).
 *
 * Variables: ONE caller-owned flat fp32 device buffer of wdg_critic_num_floats() floats; variable i lives at
 * weight_info's `offset` (floats, 256-byte aligned, padding must stay zero).  Names/shapes/layouts are the checkpoint's
 * (`layer_with_weights-N/...`, Conv2D HWIO).  The trainable variables come first (wdg_critic_num_trainable_floats());
 * a gradient buffer has that size and the same offsets, so the optimiser step and the data-parallel all-reduce of a whole
 * model are one launch / one collective over a flat range.
 *
 * forward: low_res (B,T,S,S,Cl), high_res (B,T,S,S,Ch) fp32 device -> score (B,1).  training != 0 reproduces what Keras does
 * in a training-mode call: every SpectralNormalization wrapper runs one power iteration and rewrites its kernel and sn_u
 * IN PLACE in vars_dev (TFA 0.14), then the call reads a snapshot of the variables kept in its context.  context_dev
 * (wdg_critic_context_bytes) keeps what this call's backward needs; several contexts may be alive at once (ganbase.py:41-46
 * differentiates two calls together).  scratch_dev (wdg_critic_scratch_bytes) is transient and may be shared by all calls
 * on one stream.  Sizes depend on wdg_train_set_precision's mode at the time of the call; keep it unchanged between a
 * forward and its backward.  backward: dscore (B,1) = dLoss/dscore; grads_dev (trainable floats, may be NULL) receives
 * dLoss/dvariables, d_high_res_dev ((B,T,S,S,Ch), may be NULL) dLoss/dhigh_res.  A context can be differentiated once
 * (the backward overwrites the saved gate activations). */
typedef struct wdg_critic wdg_critic;
int wdg_critic_create(wdg_critic** out, int low_res_size, int high_res_size, int low_res_channels, int high_res_channels,
                      int n_timesteps, int feature_channels, int ckpt_topology);
void wdg_critic_destroy(wdg_critic* c);
int wdg_critic_num_weights(const wdg_critic* c);
int wdg_critic_weight_info(const wdg_critic* c, int index, const char** name, int64_t* dims, int* ndim, int64_t* offset,
                           int* trainable);
int64_t wdg_critic_num_floats(const wdg_critic* c);
int64_t wdg_critic_num_trainable_floats(const wdg_critic* c);
int wdg_critic_context_bytes(const wdg_critic* c, int B, int T, int training, size_t* bytes);
int wdg_critic_scratch_bytes(const wdg_critic* c, int B, int T, size_t* bytes);
/* One training-mode call's effect on the variables without the forward pass (all SpectralNormalization wrappers). */
int wdg_critic_sn_update(wdg_critic* c, float* vars_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
int wdg_critic_forward(wdg_critic* c, float* vars_dev, const float* low_res_dev, const float* high_res_dev, float* score_dev,
                       int B, int T, int training, void* context_dev, size_t context_bytes, void* scratch_dev,
                       size_t scratch_bytes, void* stream);
int wdg_critic_backward(wdg_critic* c, const float* vars_dev, void* context_dev, int B, int T, int training,
                        const float* dscore_dev, float* grads_dev, float* d_high_res_dev, void* scratch_dev,
                        size_t scratch_bytes, void* stream);
int wdg_critic_backward_input(wdg_critic* c, const float* vars_dev, void* context_dev, int B, int T, int training,
                              const float* dscore_dev, float* d_high_res_dev, void* scratch_dev, size_t scratch_bytes,
                              void* stream);
int wdg_critic_backward_weights(wdg_critic* c, const float* vars_dev, void* context_dev, int B, int T, int training,
                                const float* dscore_dev, float* grads_dev, void* scratch_dev, size_t scratch_bytes,
                                void* stream);

/* ---- host helper of the TensorFlow checkpoint-V2 reader / writer (tf_checkpoint.py; ganbase.py:132-140): CRC-32C
 * (Castagnoli) of `n` bytes, continuing from `crc` (0 to start).  TensorFlow stores mask(crc) = rotr(crc, 15) + 0xa282ead8
 * in SSTable block trailers and BundleEntryProto.crc32c. */
uint32_t wdg_crc32c(uint32_t crc, const void* data, size_t n);

/* ---- on-device evaluation metrics (reference gan/metrics.py).  real / fake: fp32 [B,T,H,W,C].
 * wdg_metrics_pointwise: out[5][B] = wind_speed_weighted_rmse (metrics.py:32-45), wind_speed_rmse (:81-91),
 * angular_cosine_distance (:97-105), opposite_cosine_similarity (:108-111), extreme_weighted_rmse (:66-73), all in
 * one pass; pixels_per_sample = T*H*W; scratch from wdg_metrics_pointwise_scratch.
 * wdg_metric_lsd: log_spectral_distance (:121-137) -- like tf.signal.rfft2d on the [B,T,H,W,C] tensor the transform
 * runs over the two innermost axes (W, C); out[B]; scratch >= B*T*H doubles.
 * wdg_metric_spatial_ks: spatially_convolved_ks_stat (:155-187), patch x patch windows, stride 1; out[(H-patch+1)*(W-patch+1)]. */
int wdg_metrics_pointwise_scratch(int B, size_t* bytes);
int wdg_metrics_pointwise(const float* real, const float* fake, int B, long long pixels_per_sample, int C, float* out,
                          void* scratch, void* stream);
int wdg_metric_lsd(const float* real, const float* fake, int B, int T, int H, int W, int C, float* out, void* scratch, void* stream);
int wdg_metric_spatial_ks(const float* real, const float* fake, int B, int T, int H, int W, int C, int patch, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WDG_H */
