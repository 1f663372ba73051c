// ConvLSTM steps on CTA PAIRS: tcgen05.mma.cta_group::2 (sm_100a), persistent over all timesteps.
//
// Same implicit GEMM as conv_umma_kernel<256, EPI_LSTM> (conv_umma.cuh): per step D[pixels, 4 gates x 64 ch] =
// [x_t taps | h_{t-1} taps] * packed weights, gate math in the TMEM epilogue -- but two CTAs on the two SMs of a TPC share
// one 256 x 256 accumulator tile:
//   * CTA rank r of the pair owns the 128 pixels of M tile 2*pair + r: it TMA-loads its own A rows and HALF of the
//     B tile (weight rows [r*128, r*128+128) of the 256-column N tile) into its own shared memory;
//   * the leader (rank 0) issues ONE tcgen05.mma.cta_group::2 per K step (M = 256): each SM's tensor core reads its own A
//     and both halves of B, each half fetched from ONE SM's shared memory for both -- per SM and K-block the shared-memory
//     port moves 64 KB (32 written by TMA, 32 read) instead of 96 KB (48 + 48) in the single-CTA kernel, whose tensor pipe
//     the port capped at ~70 % (96 KB per 512 MMA cycles = 187 B/clk against the 128 B/clk port);
//   * 32 KB stages instead of 48 KB: six stages deep instead of four;
//   * each CTA runs its own epilogue on its own 128 TMEM lanes.
// Barriers: every TMA load (both CTAs) completes on the LEADER's `full` barrier (cp.async.bulk.tensor ... cta_group::2 with the
// mbarrier address mapped to rank 0); tcgen05.commit.cta_group::2.multicast arrives on the `empty` (stage free) and `tfull`
// (accumulator ready) barriers of BOTH CTAs; the peer's epilogue warps release a TMEM stage with a remote arrive on the
// leader's `tempty` barrier.  Time loop and per-image-group step flags: as in conv_umma.cuh (sync_flags).
#pragma once
#include "conv_umma.cuh"

namespace wdg {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes complete on the mbarrier at
// `bar_cluster_addr` (the leader's barrier)
__device__ __forceinline__ void tma2_load_2d(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
template <int PREC>
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (PREC == PREC_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive (once all previously issued MMAs of the pair have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int BN_>
struct PairCfg {
  static constexpr int BN = BN_;                       // accumulator columns (EPI_LSTM: 256 = 4 gates x 64 channels)
  static constexpr int B_HALF_BYTES = (BN / 2) * 128;  // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;   // 32 KB (BN 256) / 24 KB (BN 128)
  static constexpr int STAGES = BN >= 256 ? 6 : 8;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int VEC_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 3 * VEC_COLS * 4;
  static_assert(BN == 128 || BN == 256, "pair kernel is instantiated for 128 / 256 accumulator columns");
};

// EPI_LSTM (BN 256): the persistent ConvLSTM.  EPI_AFFINE (BN 128): bias -> LeakyReLU -> folded BatchNorm convolution layers
// with the same K-block / tile description as conv_umma_kernel (the 4x4 stride-2 convolution: its N = 128 MMAs read 8 KB of
// operands per 64 cycles in the single-CTA kernel -- the whole shared-memory port -- and 6 KB here).
template <int BN, int EPI, int PREC>
__global__ void __launch_bounds__(192, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvParams p) {
  static_assert(EPI == EPI_LSTM || EPI == EPI_AFFINE, "pair kernel epilogues: ConvLSTM gates, affine");
  static_assert(EPI != EPI_LSTM || BN == 256, "LSTM epilogue expects 4 gates x 64 channels");
  using Cfg = PairCfg<BN>;
  using P = Prec<PREC>;
  using act_t = typename P::act_t;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs (the pair's MMAs and multicast commits address both by the leader's offsets)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // leader only: bytes of all four TMA loads of a stage
  uint64_t* empty_bar = bars + STAGES;           // both: stage consumed (multicast commit)
  uint64_t* tfull_bar = bars + 2 * STAGES;       // both: accumulator complete (multicast commit)
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // leader only: 8 epilogue warps of the pair drained the TMEM stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* sm_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);
  float* sm_scale = sm_bias + Cfg::VEC_COLS;
  float* sm_shift = sm_scale + Cfg::VEC_COLS;
  for (int i = threadIdx.x; i < p.n_tiles_N * BN && i < Cfg::VEC_COLS; i += blockDim.x) {
    sm_bias[i] = p.ep.bias[i];
    if constexpr (EPI == EPI_AFFINE) { sm_scale[i] = p.ep.scale[i]; sm_shift[i] = p.ep.shift[i]; }
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int total_pair_tiles = ((m_tiles + 1) >> 1) * p.n_tiles_N;
  const int t_begin = (EPI == EPI_LSTM) ? p.t_begin : 0, t_end = (EPI == EPI_LSTM) ? p.t_end : 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmH);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of BOTH CTAs initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int nkb = (EPI == EPI_LSTM && t == 0) ? p.num_kb_first : p.num_kb;
      const bool h_sync = (EPI == EPI_LSTM) && p.ep.sync_flags != nullptr && t > t_begin;
      for (int pt = pair; pt < total_pair_tiles; pt += n_pairs) {
        const int n_tile = pt % p.n_tiles_N;
        const int m_tile = 2 * (pt / p.n_tiles_N) + (int)rank;       // beyond m_tiles (odd count): TMA zero-fills, epilogue masks
        const int tx = m_tile % p.tiles_x;
        const int ty = (m_tile / p.tiles_x) % p.tiles_y;
        const int tn = m_tile / (p.tiles_x * p.tiles_y);
        const int n0 = tn * p.tile_n;
        const int b0 = (p.ntile_coord == 0 ? n_tile : 0);
        const int b1 = tx * p.tile_w + (p.ntile_coord == 1 ? n_tile : 0);
        const int b2 = ty * p.tile_h + (p.ntile_coord == 2 ? n_tile : 0);
        const int b3 = (p.n_coord == 3 ? n0 : 0) + (p.ntile_coord == 3 ? n_tile : 0) + (EPI == EPI_LSTM ? t : 0);
        const int b4 = (p.n_coord == 4 ? n0 : 0);
        bool h_ready = !h_sync || tn >= p.tiles_n;          // (a padding tile of an odd tile count reads zeros only)
        for (int kb = 0; kb < nkb; ++kb) {
          const KBlock k = p.kb[kb];
          if (!h_ready && k.src == 1) {
            flag_wait(p.ep.sync_flags + (long long)(t - 1) * p.tiles_n + tn, p.ep.sync_total);
            fence_proxy_async_global();
            h_ready = true;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t full0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);     // both CTAs' A tile + B half
            tma2_load_5d(smA + stage * A_STAGE_BYTES, k.src == 0 ? &tmX : &tmH, full0, b0 + k.o0, b1 + k.o1, b2 + k.o2, b3 + k.o3, b4);
            tma2_load_2d(smB + stage * Cfg::B_HALF_BYTES, &tmB, full0, kb * P::KB_ELEMS, n_tile * BN + (int)rank * (BN / 2));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader only): M = 256 over the pair
    if (leader) {
      constexpr uint32_t idesc = P::idesc(2 * TILE_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int nkb = (EPI == EPI_LSTM && t == 0) ? p.num_kb_first : p.num_kb;
        for (int pt = pair; pt < total_pair_tiles; pt += n_pairs) {
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = umma_desc_kmajor(smem_u32(smA + stage * A_STAGE_BYTES), 128u);
              const uint64_t db = umma_desc_kmajor(smem_u32(smB + stage * Cfg::B_HALF_BYTES), 128u);
              umma2<PREC>(d_tmem, da, db, idesc, kb ? 1u : 0u);
              umma2<PREC>(d_tmem, da + 2, db + 2, idesc, 1u);
              umma2<PREC>(d_tmem, da + 4, db + 4, idesc, 1u);
              umma2<PREC>(d_tmem, da + 6, db + 6, idesc, 1u);
              umma2_commit_both(&empty_bar[stage]);
              if (kb == nkb - 1) umma2_commit_both(&tfull_bar[as]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5 of both CTAs, own 128 TMEM lanes)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int lx = row % p.tile_w;
    const int ly = (row / p.tile_w) % p.tile_h;
    const int ln = row / (p.tile_w * p.tile_h);
    const EpiParams& e = p.ep;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      for (int pt = pair; pt < total_pair_tiles; pt += n_pairs) {
        const int n_tile = pt % p.n_tiles_N;
        const int m_tile = 2 * (pt / p.n_tiles_N) + (int)rank;
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
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[2][16];
            tmem_ld16(taddr + c0, r[0]);
            tmem_ld16(taddr + c0 + 16, r[1]);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const int col = n_tile * BN + c0 + 16 * g;
                const long long off = (long long)n * e.out_sn + (long long)y * e.out_sy + (long long)x * e.out_sx + e.out_c0 + col;
                float v[16];
                affine16(r[g], sm_bias + col, sm_scale + col, sm_shift + col, e.lrelu != 0, v);
                P::store16(reinterpret_cast<act_t*>(e.out) + off, v);
              }
            }
          }
        } else {
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
            if (t == 0) {
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
        if (e.sync_flags != nullptr && t + 1 < t_end && tn < p.tiles_n) {
          __threadfence();
          fence_proxy_async_global();
          __syncwarp();
          if (lane == 0) flag_arrive(e.sync_flags + (long long)t * p.tiles_n + tn);
        }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));     // the leader's barrier
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA leaves (or frees TMEM) while the other may still address it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace wdg
