// Fused UpSampling2D(2, bilinear) + Conv2DTranspose(16, 5x5, same) (models.py:62-64) as an anchor GEMM with
// shared-memory HALO REUSE: the A operand of all 16 filter taps is the same block of the flattened,
// zero-padded low-res image, shifted by (dy*PW + dx) rows.
//
//   pass  = 256 consecutive flat positions (2 GEMM tiles of 128 anchors)
//   A     = 3 channel chunks x 416 rows x 128 B (SWIZZLE_128B), loaded ONCE per pass; tap (dy,dx) of tile t is
//           the plain UMMA descriptor starting at row t*128 + dy*PW + dx of the halo buffer
//   B     = composed weights [64 = 2x2 phases x 16 ch][48 K-blocks x 64], streamed through a ring, each K-block
//           feeding both tiles
//   D     = 2 tiles x 64 fp32 columns in TMEM, double buffered (256 columns)
//
// Compared with one TMA box per (tap, chunk) this cuts L2->SM traffic from ~1070 KB to ~280 KB per 128 anchors.
// Warp roles: 0 = A producer, 1 = MMA issuer (+TMEM alloc), 2-5 = epilogue, 6 = B producer.
#pragma once
#include "conv_umma.cuh"

namespace wdg {

struct UpHaloParams {
  int num_passes;   // ceil(total flat positions / 256)
  int n_img;        // images (fields)
  int pw, ph;       // padded low-res width / height
  int S;            // high-res size
  const float* delta;   // fp32 border corrections [n][S][192]
  const float* bias;    // [16]
  const float* scale;
  const float* shift;
  __nv_bfloat16* out;   // [n][S][S][16]
};

constexpr int UH_TILES = 2;
constexpr int UH_ROWS = 416;                       // 256 + max tap shift (3*52+3 = 159) rounded up to 2 x 208
constexpr int UH_BOX_ROWS = 208;
constexpr int UH_A_BYTES = UH_ROWS * 128;          // one channel chunk
constexpr int UH_B_BYTES = 64 * 128;
constexpr int UH_BSTAGES = 6;
constexpr int UH_SMEM = 3 * UH_A_BYTES + UH_BSTAGES * UH_B_BYTES + 1024 + 256;
constexpr int UH_TMEM_COLS = 256;

__global__ void __launch_bounds__(224, 1)
upconv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ UpHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + 3 * UH_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + UH_BSTAGES * UH_B_BYTES);
  uint64_t* a_full = bars;                 // [3]
  uint64_t* a_empty = bars + 3;            // [3]
  uint64_t* b_full = bars + 6;             // [UH_BSTAGES]
  uint64_t* b_empty = b_full + UH_BSTAGES; // [UH_BSTAGES]
  uint64_t* tfull = b_empty + UH_BSTAGES;  // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < 3; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < UH_BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<UH_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer / MMA loops run warp-uniformly; one elected lane issues the TMA / tcgen05 instructions.
  if (warp == 0) {
    // ================================================= A producer: 3 chunk buffers, one fill per pass
    uint32_t phase = 0;
    for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x) {
      const int f0 = pass * (UH_TILES * TILE_M);
      for (int c = 0; c < 3; ++c) {
        mbar_wait(&a_empty[c], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[c], UH_A_BYTES);
          tma_load_2d(smA + c * UH_A_BYTES, &tmA, &a_full[c], c * 64, f0);
          tma_load_2d(smA + c * UH_A_BYTES + UH_BOX_ROWS * 128, &tmA, &a_full[c], c * 64, f0 + UH_BOX_ROWS);
        }
        __syncwarp();
      }
      phase ^= 1;
    }
  } else if (warp == 6) {
    // ================================================= B producer: ring over (pass, K-block)
    int stage = 0;
    uint32_t phase = 0;
    for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x) {
      for (int kb = 0; kb < 48; ++kb) {
        mbar_wait(&b_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[stage], UH_B_BYTES);
          tma_load_2d(smB + stage * UH_B_BYTES, &tmB, &b_full[stage], kb * 64, 0);
        }
        __syncwarp();
        if (++stage == UH_BSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, 64);
    int stage = 0;
    uint32_t bphase = 0, aphase = 0, tphase = 0;
    int as = 0;
    for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x) {
      mbar_wait(&tempty[as], tphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (UH_TILES * 64);
      for (int c = 0; c < 3; ++c) {
        mbar_wait(&a_full[c], aphase);
        tc_fence_after();
        // SWIZZLE_128B is applied on absolute shared-memory address bits, so a tile that starts at any 128-byte
        // row of the 1024-byte aligned halo buffer is addressed by the plain descriptor (matrix base offset 0).
        const uint64_t a_desc0 = umma_desc_kmajor(smem_u32(smA + c * UH_A_BYTES), 128u);
        for (int tap = 0; tap < 16; ++tap) {
          mbar_wait(&b_full[stage], bphase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t db = umma_desc_kmajor(smem_u32(smB + stage * UH_B_BYTES), 128u);
            const uint32_t shift_rows = (uint32_t)((tap >> 2) * p.pw + (tap & 3));
#pragma unroll
            for (int t = 0; t < UH_TILES; ++t) {
              const uint64_t da = a_desc0 + (uint64_t)((t * TILE_M + shift_rows) * 8);   // rows * 128 B >> 4
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem + t * 64, da + 2 * k, db + 2 * k, idesc, (c | tap | k) ? 1u : 0u);
            }
            umma_commit(&b_empty[stage]);
            if (tap == 15) {
              umma_commit(&a_empty[c]);
              if (c == 2) umma_commit(&tfull[as]);
            }
          }
          __syncwarp();
          if (++stage == UH_BSTAGES) { stage = 0; bphase ^= 1; }
        }
      }
      aphase ^= 1;
      if (++as == 2) { as = 0; tphase ^= 1; }
    }
  } else {
    // ================================================= epilogue (warps 2..5)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int per_img = p.pw * p.ph;
    const int S = p.S;
    int as = 0;
    uint32_t tphase = 0;
    for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x) {
      mbar_wait(&tfull[as], tphase);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < UH_TILES; ++t) {
        const long long f = (long long)pass * (UH_TILES * TILE_M) + t * TILE_M + row;
        const int img = (int)(f / per_img);
        const int rem = (int)(f - (long long)img * per_img);
        const int pr = rem / p.pw, ps = rem - pr * p.pw;
        const bool arow = img < p.n_img;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (UH_TILES * 64) + t * 64;
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
              const float a = leaky02(v[i] + __ldg(p.bias + i));
              v[i] = a * __ldg(p.scale + i) + __ldg(p.shift + i);
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + (((long long)img * S + Y) * S + X) * 16);
            dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                pack_bf16x2(v[6], v[7]));
            dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                pack_bf16x2(v[14], v[15]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; tphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<UH_TMEM_COLS>(tmem_base);
  }
}

}  // namespace wdg
