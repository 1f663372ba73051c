// Implicit-GEMM convolution with shared-memory HALO REUSE on tcgen05 (sm_100a).
//
// The image is addressed as a flat list of rows (one row = one pixel, or one sliding window of pixels);
// a GEMM tile is 128 CONSECUTIVE flat positions and filter tap `t` reads the same rows shifted by
// tap_shift[t] (= dy * pitch + dx).  So the A operand of every tap is the same block of shared memory:
//
//   pass  = 256 consecutive flat positions (2 GEMM tiles)
//   A     = per channel chunk 416 rows x 128 B (SWIZZLE_128B), loaded ONCE per pass by two TMA boxes into a
//           2-deep ring (K order is chunk-major, so chunk c+1 streams in while chunk c is consumed); the operand
//           of (tile, tap) is the plain UMMA descriptor that starts `tile*128 + shift` rows into the chunk buffer
//           -- SWIZZLE_128B is applied on absolute smem address bits, so any 128-byte row of the 1024-byte
//           aligned buffer is a valid start (verified on B200)
//   B     = packed weights [BN][NCHUNK*NTAP*64] (K-block = chunk*NTAP + tap), streamed through a ring of
//           TPS-tap stages; each stage feeds both tiles (2*4*TPS MMAs per barrier round trip)
//   D     = 2 tiles x BN fp32 columns in TMEM, double buffered
//
// Positions whose window crosses an image row/edge compute garbage and are masked by the epilogue.
// Warp roles: 0 = A producer, 1 = MMA issuer (+TMEM alloc), 2-5 = epilogue, 6 = B producer.
//
// PAIR = true: two CTAs on the two SMs of a TPC (cluster of 2) work on two consecutive passes with ONE
// tcgen05.mma.cta_group::2 per (tile, tap, K slice) (M = 256: tile t of both passes).  CTA rank r loads the A chunk of
// pass 2*pp + r and rows [r*BN/2, (r+1)*BN/2) of every weight block; each SM's tensor core reads its own A and both B
// halves, each half served by one SM's shared memory for both.  N = 128 MMAs read 8 KB of operands per 64 cycles in
// the single-CTA form -- the whole 128 B/clk port, before any TMA write -- and 6 KB in pair form.  Barrier protocol as in
// conv_umma2.cuh (loads complete on the leader's barriers, commits are multicast, remote tempty arrives).
#pragma once
#include "conv_umma.cuh"
#include "conv_umma2.cuh"

namespace wdg {

enum { HEPI_UPCONV = 0, HEPI_AFFINE = 1, HEPI_FINAL = 2 };

struct HaloParams {
  int num_passes;        // ceil(total flat positions / 256)
  int n_img;             // images
  int pw, ph;            // flat positions per image = pw * ph (row pitch pw)
  int tap_shift[16];     // row shift of each tap
  int box_rows;          // rows per A TMA box (two boxes per chunk; <= H_BOX_ROWS): 2 * box_rows >= 256 + max tap shift
  unsigned char kmask[16];   // per tap: bit k set = the k-th 16-element K slice of the tap has non-zero weights (issue its MMA)
  unsigned char kskip_tail;  // K slices of the LAST channel chunk that are pure padding (zero activations and weights): not issued
  const float* bias;     // [BN] (HEPI_UPCONV: [16])
  const float* scale;
  const float* shift;
  // ---- HEPI_UPCONV: anchors of the fused bilinear x2 + 5x5 transposed conv
  int S;                 // high-res size
  const float* delta;    // fp32 border corrections [n][S][192]
  void* out;             // act_t; pixel (n, Y, X) channel c at out + n*up_sn + Y*up_sy + X*16 + c
  long long up_sn, up_sy;
  // ---- HEPI_FINAL: 3x3 conv 16 -> 2 over "super-pixels" (4 pixels x 16 channels = one 128-byte row); GEMM columns
  //      0..7 = (pixel in super-pixel, output channel), 8..15 = the same with the bf16 residual of the weights;
  //      fp32 out[n][S][S][2], bias[2]
  float* outf;
  // ---- HEPI_AFFINE: valid outputs are (y < vh, x < vw); v = leaky(acc + bias) * scale + shift -> bf16, two destinations
  int vw, vh;
  void* out1;            // act_t
  long long o1_sn, o1_sy, o1_sx;
  void* out2;
  long long o2_sn, o2_sy, o2_sx;
  int o2_c0;
};

constexpr int H_TILES = 2;
constexpr int H_ROWS = 416;                      // 256 + max tap shift (<= 160), two TMA boxes of 208 rows
constexpr int H_BOX_ROWS = 208;
constexpr int H_A_BYTES = H_ROWS * 128;          // one channel chunk
constexpr int H_ABUFS = 2;                       // A ring depth (chunks in flight)

template <int BN, int NCHUNK, int TPS>
struct HaloCfg {
  static constexpr int B_STAGE_BYTES = TPS * BN * 128;
  static constexpr int A_BYTES = H_ABUFS * H_A_BYTES;
  static constexpr int VEC_BYTES = 3 * 128 * 4;   // bias | scale | shift of up to 128 columns, staged in shared memory
  static constexpr int BSTAGES = (226 * 1024 - 1024 - 512 - VEC_BYTES - A_BYTES) / B_STAGE_BYTES > 8
                                     ? 8 : (226 * 1024 - 1024 - 512 - VEC_BYTES - A_BYTES) / B_STAGE_BYTES;
  static constexpr int SMEM = A_BYTES + BSTAGES * B_STAGE_BYTES + 1024 + 512 + VEC_BYTES;
  static constexpr int TMEM_COLS = (2 * H_TILES * BN <= 256) ? 256 : 512;
  static_assert(BSTAGES >= 2, "not enough shared memory for the B ring");
  static_assert(2 * H_TILES * BN <= 512, "accumulators exceed TMEM");
};

template <int BN, int NCHUNK, int NTAP, int TPS, int EPI, int PREC, bool PAIR = false>
__global__ void __launch_bounds__(224, 1)
halo_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ HaloParams p) {
  using Cfg = HaloCfg<PAIR ? BN / 2 : BN, NCHUNK, TPS>;     // PAIR: this CTA's half of every weight block
  constexpr int BROWS = PAIR ? BN / 2 : BN;                 // weight rows per tap in this CTA's shared memory
  static_assert(!PAIR || (BN % 32 == 0), "pair form: each half of B must be whole 16-row groups");
  using P = Prec<PREC>;
  using act_t = typename P::act_t;
  constexpr int BSTAGES = Cfg::BSTAGES;
  static_assert(NTAP % TPS == 0, "taps per stage must divide the tap count");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + BSTAGES * Cfg::B_STAGE_BYTES);
  uint64_t* a_full = bars;                      // [H_ABUFS]
  uint64_t* a_empty = a_full + H_ABUFS;         // [H_ABUFS]
  uint64_t* b_full = a_empty + H_ABUFS;         // [BSTAGES]
  uint64_t* b_empty = b_full + BSTAGES;         // [BSTAGES]
  uint64_t* tfull = b_empty + BSTAGES;          // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sm_bias = reinterpret_cast<float*>(smB + BSTAGES * Cfg::B_STAGE_BYTES + 512);   // see conv_umma.cuh
  float* sm_scale = sm_bias + 128;
  float* sm_shift = sm_scale + 128;
  if constexpr (EPI != HEPI_FINAL) {
    const int ncols = (EPI == HEPI_UPCONV) ? 16 : BN;
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) { sm_bias[i] = p.bias[i]; sm_scale[i] = p.scale[i]; sm_shift[i] = p.shift[i]; }
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TMEM_COLS = (2 * H_TILES * BN <= 256) ? 256 : 512;     // the pair's accumulators span the full BN columns
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // work items: passes (single CTA) or pairs of consecutive passes (PAIR); this CTA's pass of item i is first + i*step
  const int n_items = PAIR ? (p.num_passes + 1) / 2 : p.num_passes;
  const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto pass_of = [&](int item) { return PAIR ? 2 * item + (int)rank : item; };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < H_ABUFS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], PAIR ? 8 : 4); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc2<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  // Producer / MMA loops run warp-uniformly; one elected lane issues the TMA / tcgen05 instructions.
  if (warp == 0) {
    // ================================================= A producer: ring over (pass, chunk)
    int ab = 0;
    uint32_t phase = 0;
    for (int item = item0; item < n_items; item += item_step) {
      const int f0 = pass_of(item) * (H_TILES * TILE_M);       // beyond the last pass (odd count, PAIR): TMA zero-fills
      for (int c = 0; c < NCHUNK; ++c) {
        mbar_wait(&a_empty[ab], phase ^ 1);
        if (elect_one()) {
          if constexpr (PAIR) {
            const uint32_t bar0 = mapa_u32(smem_u32(&a_full[ab]), 0);
            if (leader) mbar_arrive_expect_tx(&a_full[ab], 2 * 2 * p.box_rows * 128);     // both CTAs' boxes
            tma2_load_2d(smA + ab * H_A_BYTES, &tmA, bar0, c * P::KB_ELEMS, f0);
            tma2_load_2d(smA + ab * H_A_BYTES + p.box_rows * 128, &tmA, bar0, c * P::KB_ELEMS, f0 + p.box_rows);
          } else {
            mbar_arrive_expect_tx(&a_full[ab], 2 * p.box_rows * 128);
            tma_load_2d(smA + ab * H_A_BYTES, &tmA, &a_full[ab], c * P::KB_ELEMS, f0);
            tma_load_2d(smA + ab * H_A_BYTES + p.box_rows * 128, &tmA, &a_full[ab], c * P::KB_ELEMS, f0 + p.box_rows);
          }
        }
        __syncwarp();
        if (++ab == H_ABUFS) { ab = 0; phase ^= 1; }
      }
    }
  } else if (warp == 6) {
    // ================================================= B producer: ring over (pass, stage of TPS K-blocks)
    int stage = 0;
    uint32_t phase = 0;
    for (int item = item0; item < n_items; item += item_step) {
      for (int g = 0; g < NCHUNK * NTAP / TPS; ++g) {
        mbar_wait(&b_empty[stage], phase ^ 1);
        if (elect_one()) {
          if constexpr (PAIR) {
            const uint32_t bar0 = mapa_u32(smem_u32(&b_full[stage]), 0);
            if (leader) mbar_arrive_expect_tx(&b_full[stage], 2 * Cfg::B_STAGE_BYTES);    // both halves
#pragma unroll
            for (int j = 0; j < TPS; ++j)
              tma2_load_2d(smB + stage * Cfg::B_STAGE_BYTES + j * BROWS * 128, &tmB, bar0, (g * TPS + j) * P::KB_ELEMS, (int)rank * BROWS);
          } else {
            mbar_arrive_expect_tx(&b_full[stage], Cfg::B_STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < TPS; ++j)
              tma_load_2d(smB + stage * Cfg::B_STAGE_BYTES + j * BN * 128, &tmB, &b_full[stage], (g * TPS + j) * P::KB_ELEMS, 0);
          }
        }
        __syncwarp();
        if (++stage == BSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer
    constexpr uint32_t idesc = P::idesc(PAIR ? 2 * TILE_M : TILE_M, BN);
    int stage = 0, ab = 0;
    uint32_t bphase = 0, aphase = 0, tphase = 0;
    int as = 0;
    for (int item = item0; item < n_items && leader; item += item_step) {      // PAIR: the leader issues for both CTAs
      mbar_wait(&tempty[as], tphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (H_TILES * BN);
      for (int c = 0; c < NCHUNK; ++c) {
        mbar_wait(&a_full[ab], aphase);
        tc_fence_after();
        const uint64_t a_desc0 = umma_desc_kmajor(smem_u32(smA + ab * H_A_BYTES), 128u);
        for (int g = 0; g < NTAP / TPS; ++g) {
          mbar_wait(&b_full[stage], bphase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t b_desc0 = umma_desc_kmajor(smem_u32(smB + stage * Cfg::B_STAGE_BYTES), 128u);
#pragma unroll
            for (int j = 0; j < TPS; ++j) {
              const int tap = g * TPS + j;
              const uint32_t shift_rows = (uint32_t)p.tap_shift[tap];
              const uint32_t kmask = p.kmask[tap] & ~((c == NCHUNK - 1) ? (uint32_t)p.kskip_tail : 0u);
              const int first_k = __ffs((int)kmask) - 1;     // the first MMA of a tile overwrites the accumulator
              const uint64_t db = b_desc0 + (uint64_t)(j * BROWS * 8);                       // rows * 128 B >> 4
#pragma unroll
              for (int t = 0; t < H_TILES; ++t) {
                const uint64_t da = a_desc0 + (uint64_t)((t * TILE_M + shift_rows) * 8);     // rows * 128 B >> 4
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (kmask & (1u << k)) {
                    const uint32_t acc = (c | tap) ? 1u : (k != first_k ? 1u : 0u);
                    if constexpr (PAIR) umma2<PREC>(d_tmem + t * BN, da + 2 * k, db + 2 * k, idesc, acc);
                    else P::mma(d_tmem + t * BN, da + 2 * k, db + 2 * k, idesc, acc);
                  }
              }
            }
            if constexpr (PAIR) {
              umma2_commit_both(&b_empty[stage]);
              if (g == NTAP / TPS - 1) {
                umma2_commit_both(&a_empty[ab]);
                if (c == NCHUNK - 1) umma2_commit_both(&tfull[as]);
              }
            } else {
              umma_commit(&b_empty[stage]);
              if (g == NTAP / TPS - 1) {
                umma_commit(&a_empty[ab]);
                if (c == NCHUNK - 1) umma_commit(&tfull[as]);
              }
            }
          }
          __syncwarp();
          if (++stage == BSTAGES) { stage = 0; bphase ^= 1; }
        }
        if (++ab == H_ABUFS) { ab = 0; aphase ^= 1; }
      }
      if (++as == 2) { as = 0; tphase ^= 1; }
    }
  } else {
    // ================================================= epilogue (warps 2..5)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int per_img = p.pw * p.ph;
    int as = 0;
    uint32_t tphase = 0;
    for (int item = item0; item < n_items; item += item_step) {
      const int pass = pass_of(item);
      mbar_wait(&tfull[as], tphase);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < H_TILES; ++t) {
        const long long f = (long long)pass * (H_TILES * TILE_M) + t * TILE_M + row;
        const int img = (int)(f / per_img);
        const int rem = (int)(f - (long long)img * per_img);
        const int pr = rem / p.pw, ps = rem - pr * p.pw;
        const bool arow = img < p.n_img;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (H_TILES * BN) + t * BN;
        if constexpr (EPI == HEPI_UPCONV) {
          static_assert(EPI != HEPI_UPCONV || BN == 64, "UPCONV expects 2x2 phases x 16 channels");
          const int S = p.S;
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            uint32_t r[16];
            tmem_ld16(taddr + g * 16, r);
            tmem_ld_wait();
            const int Y = 2 * pr - 1 + (g >> 1);
            const int X = 2 * ps - 1 + (g & 1);
            if (arow && Y >= 0 && Y < S && X >= 0 && X < S) {
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
              const float* D = p.delta + (long long)img * S * 192;
              if (Y < 3 || Y > S - 4) {
                const float4* d = reinterpret_cast<const float4*>(D + X * 192 + (Y < 3 ? Y * 16 : 48 + (S - 1 - Y) * 16));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 q = __ldg(d + i);
                  v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
                }
              }
              if (X < 3 || X > S - 4) {
                const float4* d = reinterpret_cast<const float4*>(D + Y * 192 + (X < 3 ? 96 + X * 16 : 144 + (S - 1 - X) * 16));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 q = __ldg(d + i);
                  v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float a = leaky02(v[i] + sm_bias[i]);
                v[i] = a * sm_scale[i] + sm_shift[i];
              }
              P::store16_exact(reinterpret_cast<act_t*>(p.out) + (long long)img * p.up_sn + (long long)Y * p.up_sy + X * 16, v);
            }
          }
        } else if constexpr (EPI == HEPI_FINAL) {
          static_assert(EPI != HEPI_FINAL || (BN == 16 && PREC == PREC_BF16), "FINAL expects 4 bf16 pixels x 2 channels (+ 8 padding columns)");
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
          if (arow && pr < p.S && ps < (p.S >> 2)) {
            const float b0 = __ldg(p.bias), b1 = __ldg(p.bias + 1);
            float4* o = reinterpret_cast<float4*>(p.outf + (((long long)img * p.S + pr) * p.S + 4 * ps) * 2);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(r[j + 8]);   // bf16 weights + their residual
            o[0] = make_float4(v[0] + b0, v[1] + b1, v[2] + b0, v[3] + b1);
            o[1] = make_float4(v[4] + b0, v[5] + b1, v[6] + b0, v[7] + b1);
          }
        } else {
          const bool valid = arow && pr < p.vh && ps < p.vw;
          act_t* d1 = reinterpret_cast<act_t*>(p.out1) + (long long)img * p.o1_sn + (long long)pr * p.o1_sy + (long long)ps * p.o1_sx;
          act_t* d2 = reinterpret_cast<act_t*>(p.out2) + (long long)img * p.o2_sn + (long long)pr * p.o2_sy + (long long)ps * p.o2_sx + p.o2_c0;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r0[16], r1[16];
            tmem_ld16(taddr + c0, r0);
            tmem_ld16(taddr + c0 + 16, r1);
            tmem_ld_wait();
            if (valid) {
              float v0[16], v1[16];
              affine16(r0, sm_bias + c0, sm_scale + c0, sm_shift + c0, true, v0);
              affine16(r1, sm_bias + c0 + 16, sm_scale + c0 + 16, sm_shift + c0 + 16, true, v1);
              P::store16(d1 + c0, v0);
              P::store16(d1 + c0 + 16, v1);
              if (p.out2) {
                P::store16(d2 + c0, v0);
                P::store16(d2 + c0 + 16, v1);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0));     // the leader's barrier
        else mbar_arrive(&tempty[as]);
      }
      if (++as == 2) { as = 0; tphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();      // neither CTA leaves (or frees TMEM) while the other may still address it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace wdg
