// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[M = 128 output pixels, N = BN channels] = sum over K-blocks A[128, 64] * B[BN, 64]^T
//
// * A (activations, bf16 channels-last) is never im2col'ed in memory: every K-block is one
//   TMA box {64 channels, tile_w, tile_h, (1,) tile_n} of a 5-D tensor map, shifted by the
//   filter tap; out-of-bounds rows/cols are zero-filled by TMA (= SAME padding).  The box
//   lands in shared memory as 128 rows x 128 B with SWIZZLE_128B, which is exactly the
//   canonical K-major UMMA operand layout.  Strided convolutions use tensor maps with
//   overlapping strides over a zero-padded image (window axis, ox, oy, row parity, n).
// * B (weights, bf16) is a pre-packed [N_total][num_kb*64] K-major matrix, one 2-D TMA box
//   {64, BN} per K-block.
// * Accumulators live in TMEM (2 stages x BN fp32 columns) so the epilogue of tile i overlaps
//   the main loop of tile i+1.  Persistent CTAs, static round-robin tile schedule.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 = epilogue.
#pragma once
#include "ptx.cuh"

namespace wdg {

constexpr int MAX_KB = 80;
constexpr int TILE_M = 128;
constexpr int A_STAGE_BYTES = TILE_M * 128;

enum { EPI_AFFINE = 0, EPI_LSTM = 1, EPI_UPCONV = 2, EPI_LSTM16_FWD = 3, EPI_LSTM16_BWD = 4, EPI_LSTM128_FWD = 5 };

struct KBlock {        // one slice of the GEMM K axis
  int8_t src;          // which A tensor map (0..2)
  int8_t half;         // 1: 32-channel slice (64-byte rows, SWIZZLE_64B) -> 2 MMAs instead of 4
  int16_t o0, o1, o2, o3;  // offsets added to the tile's base TMA coordinate (dims 0..3)
};

struct EpiParams {
  // ---- EPI_AFFINE: v = (lrelu ? leaky(acc + bias) : acc + bias) * scale + shift  (per GEMM column)
  const float* bias;
  const float* scale;
  const float* shift;
  void* out;                         // act_t (bf16 / tf32-in-fp32), or plain fp32 when out_f32
  long long out_sn, out_sy, out_sx;  // destination element strides for (image, row, col)
  int out_c0;                        // first destination channel
  int out_mul;                       // 1; 2 = 2x2 stride-2 transposed conv (pixel shuffle by column group)
  int group_cols;                    // columns per (ky,kx) group when out_mul == 2
  int lrelu;
  int out_f32;
  void* out2;                        // optional second act_t destination (same values), own strides
  long long out2_sn, out2_sy, out2_sx;
  int out2_c0;
  // ---- EPI_UPCONV (BN = 64 = 2x2 output phases x 16 channels; GEMM rows are window anchors on the
  //      flattened zero-padded low-res image, see wdg_generator.cu)
  int up_pw, up_ph;                  // padded low-res width / height (anchor grid pitch)
  int up_S;                          // high-res size
  const float* up_delta;             // fp32 border corrections [n][S][4 edges * 48]
  // ---- EPI_LSTM (BN = 256 = 4 gates x 64 channels per N tile)
  float* c_state;                    // fp32 [n][H][W][F], updated in place
  void* h_out;                       // act_t, pixel (n,y,x) channel c at n*h_sn + h_off + (y*h_pitch+x)*F + c
  long long h_sn, h_off;
  int h_pitch;                       // pixels per row of the (zero-ring padded) h image
  int F;
  long long h_step;                  // elements between consecutive timesteps of the h image sequence
  // Persistent time loop (ConvParams::t_begin .. t_end in ONE launch).  The recurrent half of a tile at step t reads
  // h_{t-1} of its own images only (3x3 halo, all channels), so the dependency is tracked per (step, image group of
  // tile_n images): every (tile, epilogue warp) that has stored its part of h_t adds 1 to sync_flags[t * tiles_n + tn]; the
  // TMA producer waits for sync_total (= every tile of that image group) arrivals on sync_flags[(t-1) * tiles_n + tn]
  // before it fetches the tile's first h_{t-1} operand.  No grid-wide barrier: CTAs drift through the (t, tile) schedule and
  // only neighbours in it ever wait for each other.  NULL: one step per launch (the kernel boundary orders the steps).
  unsigned long long* sync_flags;
  unsigned int sync_total;
  // ---- EPI_LSTM16_FWD / EPI_LSTM16_BWD: one recurrent step of the critic's 16-filter ConvLSTM2D in the TRAINING
  //      path (fp32 tensors, pixel-major [n][H][W][C]; train_lstm16.cu).  FWD (BN = 64 = [i|f|c~|o] x 16): acc =
  //      recurrent conv of h_{t-1}; t_gates holds the input conv + bias of this step and receives the activated gates;
  //      writes c_t and h_t (h_t rounded to tf32: it is only ever a GEMM operand).  BWD (BN = 16): acc = recurrent
  //      part of dL/dh_s carried from step s + 1; applies the gate backward of step s in place over t_gates (dz,
  //      rounded to tf32) and updates the carried dL/dc.
  //      EPI_LSTM128_FWD (BN = 256 = [i|f|c~|o] x 64 channels of N tile n; train_lstm16.cu): the same forward step for the
  //      generator's 128-filter cell (models.py:45): t_gates is [pix][4 * t_F] in the reference's gate-major order
  //      (column gate * t_F + channel), t_c_prev / t_c / t_h are [pix][t_F].
  int t_F;
  float* t_gates;                    // [pix][64]
  const float* t_c_prev;             // [pix][16] c_{s-1} (NULL: zero)
  float* t_c;                        // FWD: c_t out; BWD: c_s in
  float* t_h;                        // FWD: h_t out
  const float* t_dh;                 // BWD: dL/dh_s from the layers above
  float* t_dc;                       // BWD: in dL/dc_s carried, out dL/dc_{s-1}
};

struct ConvParams {
  int H, W, N;                 // GEMM M axis = N images of H x W output pixels
  int tile_w, tile_h, tile_n;  // TMA box in pixels; tile_w * tile_h * tile_n == 128
  int tiles_x, tiles_y, tiles_n;
  int n_tiles_N;               // number of BN-wide column tiles
  int num_kb;
  int num_kb_first;            // EPI_LSTM: K-blocks of step t = 0 (h_0 = 0: input half only)
  int t_begin, t_end;          // EPI_LSTM: timesteps run by this launch; KBlock::o3 is relative to t (x_t: 0, h_{t-1}: -1)
  int n_coord;                 // TMA coordinate (3 or 4) that carries the image index
  int ntile_coord;             // TMA coordinate that additionally receives the N-tile index (-1: none)
  EpiParams ep;
  KBlock kb[MAX_KB];
};

// NSTAGE = 0: as many shared-memory stages as one CTA per SM can hold (long-K layers, throughput-bound).
// NSTAGE = 3: a shallow ring sized so that TWO CTAs are resident per SM (short-K layers -- the 2x2 transposed conv has 3
// K-blocks per tile, the border GEMM 15 -- whose tiles are a TMA -> MMA -> epilogue latency chain: the second CTA
// fills the bubbles).  TMEM: 2 x BN columns per CTA, so two CTAs fit for BN <= 128.
template <int BN, int NSTAGE = 0>
struct ConvCfg {
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = NSTAGE ? NSTAGE : ((BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8));
  static constexpr int CTAS_PER_SM = (NSTAGE && NSTAGE * STAGE_BYTES <= 100 * 1024 && BN <= 128) ? 2 : 1;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int VEC_COLS = 512;   // per-column epilogue vectors (bias | scale | shift) staged in shared memory
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 3 * VEC_COLS * 4;
};

__device__ __forceinline__ float leaky02(float v) { return v >= 0.f ? v : 0.2f * v; }
__device__ __forceinline__ float hard_sigmoid(float v) { return fminf(fmaxf(0.2f * v + 0.5f, 0.f), 1.f); }

template <int BN, int EPI, int PREC, int NSTAGE = 0>
__global__ void __launch_bounds__(192, ConvCfg<BN, NSTAGE>::CTAS_PER_SM)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, NSTAGE>;
  using P = Prec<PREC>;
  using act_t = typename P::act_t;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  // Per-column epilogue vectors live in shared memory: read per 16-column group as warp-uniform LDS.128 broadcasts
  // (from global memory each group cost a full L2 round trip that nothing overlapped -- ncu: long-scoreboard stalls on
  // the first FADD of every group; the 3-K-block transposed conv spent most of its time there).
  float* sm_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);
  float* sm_scale = sm_bias + Cfg::VEC_COLS;
  float* sm_shift = sm_scale + Cfg::VEC_COLS;
  {
    const int ncols = (EPI == EPI_UPCONV) ? 16 : ((EPI == EPI_LSTM16_FWD || EPI == EPI_LSTM16_BWD || EPI == EPI_LSTM128_FWD) ? 0 : p.n_tiles_N * BN);
    for (int i = threadIdx.x; i < ncols && i < Cfg::VEC_COLS; i += blockDim.x) {
      sm_bias[i] = p.ep.bias[i];
      if constexpr (EPI == EPI_AFFINE || EPI == EPI_UPCONV) { sm_scale[i] = p.ep.scale[i]; sm_shift[i] = p.ep.shift[i]; }
    }
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t_begin = (EPI == EPI_LSTM) ? p.t_begin : 0, t_end = (EPI == EPI_LSTM) ? p.t_end : 1;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles_N;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // everything above overlapped the previous kernel's tail; its data is needed from here on
  pdl_launch_dependents();

  // Producer and MMA loops are executed by the WHOLE warp (warp-uniform control flow, so addresses and
  // descriptors stay in uniform registers); only the TMA / tcgen05 instructions themselves are issued by
  // one elected lane.
  if (warp == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int nkb = (EPI == EPI_LSTM && t == 0) ? p.num_kb_first : p.num_kb;
      // the input half of step t (x_t) does not depend on step t-1: it is fetched, and multiplied, while the tiles of
      // step t-1 that this tile's recurrent half needs are still being finished elsewhere
      const bool h_sync = (EPI == EPI_LSTM) && p.ep.sync_flags != nullptr && t > t_begin;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        bool h_ready = !h_sync;
        const int n_tile = tile % p.n_tiles_N;
        const int m_tile = tile / p.n_tiles_N;
        const int tx = m_tile % p.tiles_x;
        const int ty = (m_tile / p.tiles_x) % p.tiles_y;
        const int tn = m_tile / (p.tiles_x * p.tiles_y);
        const int n0 = tn * p.tile_n;
        const int b0 = 0 + (p.ntile_coord == 0 ? n_tile : 0);
        const int b1 = tx * p.tile_w + (p.ntile_coord == 1 ? n_tile : 0);
        const int b2 = ty * p.tile_h + (p.ntile_coord == 2 ? n_tile : 0);
        const int b3 = (p.n_coord == 3 ? n0 : 0) + (p.ntile_coord == 3 ? n_tile : 0) + (EPI == EPI_LSTM ? t : 0);
        const int b4 = (p.n_coord == 4 ? n0 : 0);
        for (int kb = 0; kb < nkb; ++kb) {
          const KBlock k = p.kb[kb];
          if constexpr (EPI == EPI_LSTM) {
            if (!h_ready && k.src == 1) {
              flag_wait(p.ep.sync_flags + (long long)(t - 1) * p.tiles_n + tn, p.ep.sync_total);
              fence_proxy_async_global();      // h_{t-1} was written through the generic proxy, TMA reads it through the async one
              h_ready = true;
            }
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t a_bytes = k.half ? (A_STAGE_BYTES / 2) : A_STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], a_bytes + Cfg::B_STAGE_BYTES);
            const CUtensorMap* tm = (k.src == 0) ? &tmA0 : ((k.src == 1) ? &tmA1 : &tmA2);
            tma_load_5d(smA + stage * A_STAGE_BYTES, tm, &full_bar[stage], b0 + k.o0, b1 + k.o1, b2 + k.o2, b3 + k.o3, b4);
            tma_load_2d(smB + stage * Cfg::B_STAGE_BYTES, &tmB, &full_bar[stage], kb * P::KB_ELEMS, n_tile * BN);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = P::idesc(TILE_M, BN);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
    const int nkb = (EPI == EPI_LSTM && t == 0) ? p.num_kb_first : p.num_kb;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const int half = p.kb[kb].half;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_kmajor(smem_u32(smA + stage * A_STAGE_BYTES), half ? 64u : 128u);
          const uint64_t db = umma_desc_kmajor(smem_u32(smB + stage * Cfg::B_STAGE_BYTES), 128u);
          // advancing K by one MMA (16 bf16 / 8 tf32 elements) = +32 bytes = +2 in the descriptor's (address >> 4) field
          P::mma(d_tmem, da, db, idesc, kb ? 1u : 0u);
          P::mma(d_tmem, da + 2, db + 2, idesc, 1u);
          if (!half) {
            P::mma(d_tmem, da + 4, db + 4, idesc, 1u);
            P::mma(d_tmem, da + 6, db + 6, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);                        // frees the smem slot once these MMAs have read it
          if (kb == nkb - 1) umma_commit(&tfull_bar[as]);        // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;      // GEMM row == pixel within the tile
    const int lx = row % p.tile_w;
    const int ly = (row / p.tile_w) % p.tile_h;
    const int ln = row / (p.tile_w * p.tile_h);
    int as = 0;
    uint32_t aphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles_N;
      const int m_tile = tile / p.n_tiles_N;
      const int tx = m_tile % p.tiles_x;
      const int ty = (m_tile / p.tiles_x) % p.tiles_y;
      const int tn = m_tile / (p.tiles_x * p.tiles_y);
      const int x = tx * p.tile_w + lx;
      const int y = ty * p.tile_h + ly;
      const int n = tn * p.tile_n + ln;
      const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;

      if constexpr (EPI == EPI_AFFINE) {
        const EpiParams& e = p.ep;
        constexpr int GROUPS = (BN % 32 == 0) ? 2 : 1;      // 16-column groups per TMEM round trip
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16 * GROUPS) {
          uint32_t r[GROUPS][16];
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) tmem_ld16(taddr + c0 + 16 * g, r[g]);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
              const int col = n_tile * BN + c0 + 16 * g;
              int oy = y, ox = x, oc = col;
              if (e.out_mul == 2) {
                const int grp = col / e.group_cols;
                oc = col - grp * e.group_cols;
                oy = 2 * y + (grp >> 1);
                ox = 2 * x + (grp & 1);
              }
              const long long off = (long long)n * e.out_sn + (long long)oy * e.out_sy + (long long)ox * e.out_sx +
                                    e.out_c0 + oc;
              float v[16];
              affine16(r[g], sm_bias + col, sm_scale + col, sm_shift + col, e.lrelu != 0, v);
              if (e.out_f32) {
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              } else {
                P::store16(reinterpret_cast<act_t*>(e.out) + off, v);
                if (e.out2)
                  P::store16(reinterpret_cast<act_t*>(e.out2) + (long long)n * e.out2_sn + (long long)oy * e.out2_sy +
                                 (long long)ox * e.out2_sx + e.out2_c0 + oc, v);
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_UPCONV) {
        // Fused bilinear-x2 + 5x5 transposed conv: row = anchor (r, s) of a 4x4 low-res window, the
        // 64 columns are the 2x2 high-res pixels (2r+3+py, 2s+3+px) x 16 channels.
        static_assert(EPI != EPI_UPCONV || BN == 64, "UPCONV epilogue expects 4 phases x 16 channels");
        const EpiParams& e = p.ep;
        const long long f = (long long)m_tile * TILE_M + row;      // flattened padded position
        const int per_img = e.up_pw * e.up_ph;
        const int img = (int)(f / per_img);
        const int rem = (int)(f - (long long)img * per_img);
        const int pr = rem / e.up_pw, ps = rem - pr * e.up_pw;
        const int S = e.up_S;
        const bool arow = img < p.N;
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          uint32_t r[16];
          tmem_ld16(taddr + g * 16, r);
          tmem_ld_wait();
          const int Y = 2 * pr - 1 + (g >> 1);   // 2*(pr-2)+3+py
          const int X = 2 * ps - 1 + (g & 1);
          if (arow && Y >= 0 && Y < S && X >= 0 && X < S) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
            const float* D = e.up_delta + (long long)img * S * 192;
            if (Y < 3) {
              const float* d = D + X * 192 + Y * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += d[i];
            } else if (Y > S - 4) {
              const float* d = D + X * 192 + 48 + (S - 1 - Y) * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += d[i];
            }
            if (X < 3) {
              const float* d = D + Y * 192 + 96 + X * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += d[i];
            } else if (X > S - 4) {
              const float* d = D + Y * 192 + 144 + (S - 1 - X) * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += d[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = leaky02(v[i] + sm_bias[i]);
              v[i] = a * sm_scale[i] + sm_shift[i];
            }
            P::store16_exact(reinterpret_cast<act_t*>(e.out) + (long long)img * e.out_sn + (long long)Y * e.out_sy + X * 16, v);
          }
        }
      } else if constexpr (EPI == EPI_LSTM16_FWD) {
        static_assert(EPI != EPI_LSTM16_FWD || BN == 64, "LSTM16 forward expects 4 gates x 16 channels");
        const EpiParams& e = p.ep;
        uint32_t zi[16], zf[16], zc[16], zo[16];
        tmem_ld16(taddr + 0, zi);
        tmem_ld16(taddr + 16, zf);
        tmem_ld16(taddr + 32, zc);
        tmem_ld16(taddr + 48, zo);
        tmem_ld_wait();
        if (valid) {
          const long long pix = ((long long)n * p.H + y) * p.W + x;
          float4* g4 = reinterpret_cast<float4*>(e.t_gates + pix * 64);
          const float4* cp4 = reinterpret_cast<const float4*>(e.t_c_prev + pix * 16);
          float4* c4 = reinterpret_cast<float4*>(e.t_c + pix * 16);
          float4* h4 = reinterpret_cast<float4*>(e.t_h + pix * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 xi = g4[q], xf = g4[4 + q], xc = g4[8 + q], xo = g4[12 + q], cp = cp4[q];
            const float xi_[4] = {xi.x, xi.y, xi.z, xi.w}, xf_[4] = {xf.x, xf.y, xf.z, xf.w};
            const float xc_[4] = {xc.x, xc.y, xc.z, xc.w}, xo_[4] = {xo.x, xo.y, xo.z, xo.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w};
            float gi[4], gf[4], gc[4], go[4], cn[4], hn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              gi[k] = hard_sigmoid(__uint_as_float(zi[4 * q + k]) + xi_[k]);
              gf[k] = hard_sigmoid(__uint_as_float(zf[4 * q + k]) + xf_[k]);
              gc[k] = tanhf(__uint_as_float(zc[4 * q + k]) + xc_[k]);
              go[k] = hard_sigmoid(__uint_as_float(zo[4 * q + k]) + xo_[k]);
              cn[k] = gf[k] * cp_[k] + gi[k] * gc[k];
              hn[k] = __uint_as_float(to_tf32(go[k] * tanhf(cn[k])));
            }
            g4[q] = make_float4(gi[0], gi[1], gi[2], gi[3]);
            g4[4 + q] = make_float4(gf[0], gf[1], gf[2], gf[3]);
            g4[8 + q] = make_float4(gc[0], gc[1], gc[2], gc[3]);
            g4[12 + q] = make_float4(go[0], go[1], go[2], go[3]);
            c4[q] = make_float4(cn[0], cn[1], cn[2], cn[3]);
            h4[q] = make_float4(hn[0], hn[1], hn[2], hn[3]);
          }
        }
      } else if constexpr (EPI == EPI_LSTM128_FWD) {
        static_assert(EPI != EPI_LSTM128_FWD || BN == 256, "LSTM128 forward expects 4 gates x 64 channels per N tile");
        const EpiParams& e = p.ep;
        const long long pix = ((long long)n * p.H + y) * p.W + x;
        const int Ft = e.t_F;
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          uint32_t zi[16], zf[16], zc[16], zo[16];
          tmem_ld16(taddr + 0 * 64 + s * 16, zi);
          tmem_ld16(taddr + 1 * 64 + s * 16, zf);
          tmem_ld16(taddr + 2 * 64 + s * 16, zc);
          tmem_ld16(taddr + 3 * 64 + s * 16, zo);
          tmem_ld_wait();
          if (valid) {
            const int ch0 = n_tile * 64 + s * 16;
            float* gp = e.t_gates + pix * 4 * Ft + ch0;          // gate g of this channel block at gp + g * Ft
            const float4* cp4 = reinterpret_cast<const float4*>(e.t_c_prev + pix * Ft + ch0);
            float4* c4 = reinterpret_cast<float4*>(e.t_c + pix * Ft + ch0);
            float4* h4 = reinterpret_cast<float4*>(e.t_h + pix * Ft + ch0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4* gi4 = reinterpret_cast<float4*>(gp) + q;
              float4* gf4 = reinterpret_cast<float4*>(gp + Ft) + q;
              float4* gc4 = reinterpret_cast<float4*>(gp + 2 * Ft) + q;
              float4* go4 = reinterpret_cast<float4*>(gp + 3 * Ft) + q;
              const float4 xi = *gi4, xf = *gf4, xc = *gc4, xo = *go4, cp = cp4[q];
              const float xi_[4] = {xi.x, xi.y, xi.z, xi.w}, xf_[4] = {xf.x, xf.y, xf.z, xf.w};
              const float xc_[4] = {xc.x, xc.y, xc.z, xc.w}, xo_[4] = {xo.x, xo.y, xo.z, xo.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w};
              float gi[4], gf[4], gc[4], go[4], cn[4], hn[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                gi[k] = hard_sigmoid(__uint_as_float(zi[4 * q + k]) + xi_[k]);
                gf[k] = hard_sigmoid(__uint_as_float(zf[4 * q + k]) + xf_[k]);
                gc[k] = tanhf(__uint_as_float(zc[4 * q + k]) + xc_[k]);
                go[k] = hard_sigmoid(__uint_as_float(zo[4 * q + k]) + xo_[k]);
                cn[k] = gf[k] * cp_[k] + gi[k] * gc[k];
                hn[k] = __uint_as_float(to_tf32(go[k] * tanhf(cn[k])));
              }
              *gi4 = make_float4(gi[0], gi[1], gi[2], gi[3]);
              *gf4 = make_float4(gf[0], gf[1], gf[2], gf[3]);
              *gc4 = make_float4(gc[0], gc[1], gc[2], gc[3]);
              *go4 = make_float4(go[0], go[1], go[2], go[3]);
              c4[q] = make_float4(cn[0], cn[1], cn[2], cn[3]);
              h4[q] = make_float4(hn[0], hn[1], hn[2], hn[3]);
            }
          }
        }
      } else if constexpr (EPI == EPI_LSTM16_BWD) {
        static_assert(EPI != EPI_LSTM16_BWD || BN == 16, "LSTM16 backward expects 16 channels");
        const EpiParams& e = p.ep;
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        if (valid) {
          const long long pix = ((long long)n * p.H + y) * p.W + x;
          float4* g4 = reinterpret_cast<float4*>(e.t_gates + pix * 64);
          const float4* cc4 = reinterpret_cast<const float4*>(e.t_c + pix * 16);
          const float4* dh4 = reinterpret_cast<const float4*>(e.t_dh + pix * 16);
          float4* dc4 = reinterpret_cast<float4*>(e.t_dc + pix * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 vi = g4[q], vf = g4[4 + q], vc = g4[8 + q], vo = g4[12 + q], cc = cc4[q], dh = dh4[q], dcv = dc4[q];
            float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.t_c_prev) cp = reinterpret_cast<const float4*>(e.t_c_prev + pix * 16)[q];
            const float gi[4] = {vi.x, vi.y, vi.z, vi.w}, gf[4] = {vf.x, vf.y, vf.z, vf.w}, gc[4] = {vc.x, vc.y, vc.z, vc.w};
            const float go[4] = {vo.x, vo.y, vo.z, vo.w}, cc_[4] = {cc.x, cc.y, cc.z, cc.w}, dh_[4] = {dh.x, dh.y, dh.z, dh.w};
            const float dc_[4] = {dcv.x, dcv.y, dcv.z, dcv.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w};
            float di[4], df[4], dg[4], dov[4], dcn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float tc = tanhf(cc_[k]);
              const float dhv = dh_[k] + __uint_as_float(r[4 * q + k]);
              const float dct = dc_[k] + dhv * go[k] * (1.f - tc * tc);
              di[k] = __uint_as_float(to_tf32(dct * gc[k] * ((gi[k] > 0.f && gi[k] < 1.f) ? 0.2f : 0.f)));
              df[k] = __uint_as_float(to_tf32(dct * cp_[k] * ((gf[k] > 0.f && gf[k] < 1.f) ? 0.2f : 0.f)));
              dg[k] = __uint_as_float(to_tf32(dct * gi[k] * (1.f - gc[k] * gc[k])));
              dov[k] = __uint_as_float(to_tf32(dhv * tc * ((go[k] > 0.f && go[k] < 1.f) ? 0.2f : 0.f)));
              dcn[k] = dct * gf[k];
            }
            g4[q] = make_float4(di[0], di[1], di[2], di[3]);
            g4[4 + q] = make_float4(df[0], df[1], df[2], df[3]);
            g4[8 + q] = make_float4(dg[0], dg[1], dg[2], dg[3]);
            g4[12 + q] = make_float4(dov[0], dov[1], dov[2], dov[3]);
            dc4[q] = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
          }
        }
      } else {
        // ConvLSTM gate epilogue.  Column layout of this N tile: [i | f | c~ | o] x 64 channels.
        static_assert(EPI != EPI_LSTM || BN == 256, "LSTM epilogue expects 4 gates x 64 channels");
        const EpiParams& e = p.ep;
        const long long pix = ((long long)n * p.H + y) * p.W + x;
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
          uint32_t zi[16], zf[16], zc[16], zo[16];
          tmem_ld16(taddr + 0 * 64 + s * 16, zi);
          tmem_ld16(taddr + 1 * 64 + s * 16, zf);
          tmem_ld16(taddr + 2 * 64 + s * 16, zc);
          tmem_ld16(taddr + 3 * 64 + s * 16, zo);
          tmem_ld_wait();
          if (valid) {
            const int ch0 = n_tile * 64 + s * 16;
            const float* bias = sm_bias + n_tile * 256 + s * 16;
            float* cptr = e.c_state + pix * e.F + ch0;
            float cprev[16];
            if (t == 0) {                      // c_{-1} = 0: skip the state read
#pragma unroll
              for (int i = 0; i < 16; ++i) cprev[i] = 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 cv = reinterpret_cast<const float4*>(cptr)[i];
                cprev[4 * i] = cv.x; cprev[4 * i + 1] = cv.y; cprev[4 * i + 2] = cv.z; cprev[4 * i + 3] = cv.w;
              }
            }
            float cn[16], hn[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float gi = hard_sigmoid(__uint_as_float(zi[i]) + bias[i]);
              const float gf = hard_sigmoid(__uint_as_float(zf[i]) + bias[64 + i]);
              const float gc = tanhf(__uint_as_float(zc[i]) + bias[128 + i]);
              const float go = hard_sigmoid(__uint_as_float(zo[i]) + bias[192 + i]);
              cn[i] = gf * cprev[i] + gi * gc;
              hn[i] = go * tanhf(cn[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(cptr)[i] = make_float4(cn[4 * i], cn[4 * i + 1], cn[4 * i + 2], cn[4 * i + 3]);
            P::store16(reinterpret_cast<act_t*>(e.h_out) + (long long)n * e.h_sn + e.h_off + t * e.h_step +
                           ((long long)y * e.h_pitch + x) * e.F + ch0, hn);
          }
        }
        if (e.sync_flags != nullptr && t + 1 < t_end) {      // publish this warp's part of h_t to the other CTAs' TMA loads
          __threadfence();
          fence_proxy_async_global();
          __syncwarp();
          if (lane == 0) flag_arrive(e.sync_flags + (long long)t * p.tiles_n + tn);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace wdg
