// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences).  No CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace wdg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("wdg: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// ------------------------------------------------- grid-wide step flags (persistent kernels: global-memory counters)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void flag_arrive(unsigned long long* flag) {
  asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(flag) : "memory");
}
// Every lane of the calling warp polls (one broadcast transaction per round) until `flag` has reached `target`.
// Bounded: a protocol bug (or CTAs that are not co-resident) traps instead of hanging the GPU box.
__device__ __forceinline__ void flag_wait(const unsigned long long* flag, unsigned long long target) {
  uint32_t spins = 0;
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= target) break;
    __nanosleep(32);
    if (++spins > (1u << 24)) {
      printf("wdg: step flag wait timed out (block %d, have %llu of %llu)\n", (int)blockIdx.x, v, target);
      __trap();
    }
  }
}

// ------------------------------------------------- programmatic dependent launch (PDL)
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream
// is still running: its prologue (barrier init, TMEM allocation, tensor-map prefetch, per-column vectors -> smem) overlaps
// the predecessor's tail.  pdl_wait() blocks until the predecessor grid has COMPLETED and its memory is visible -- it must
// precede every access to data the predecessor produced or still reads; without the launch attribute it returns at once.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by one thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 x tf32 -> fp32 (K = 8 per instruction; the tensor core ignores the low 13
// mantissa bits, so producers round to nearest with cvt.rna.tf32.f32 before the operand reaches shared memory).
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand tile whose rows are `row_bytes`
// (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B) wide; 8-row core groups are contiguous
// (SBO = 8 * row_bytes).  Field layout: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | base_offset [49,52) | layout_type [61,64).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                         // LBO (unused for swizzled K-major) = 1
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;   // SBO
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

// Instruction descriptor for kind::f16, A/B = bf16 K-major, D = fp32, shape M x N (K = 16).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Instruction descriptor for kind::tf32, A/B = tf32 K-major, D = fp32, shape M x N (K = 8).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------- misc helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t to_tf32(float v) {   // round to nearest (ties away) onto the 10-bit tf32 mantissa
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
// 256-bit global store (sm_100+): one full 32-byte sector per lane and instruction.
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// ------------------------------------------------------------- operand precision of the inference kernels
// PREC_BF16: activations / weights are bf16 in HBM, kind::f16 MMAs (K = 16), 64 channels per 128-byte K-block row.
// PREC_TF32: activations / weights are fp32 containers holding tf32-rounded values, kind::tf32 MMAs (K = 8), 32
// channels per 128-byte K-block row.  Shared-memory stage bytes, swizzle, descriptors and the number of MMAs per
// K-block (4, each advancing 32 bytes) are identical, so the two precisions share every kernel.
enum { PREC_BF16 = 0, PREC_TF32 = 1 };
template <int PREC> struct Prec;
template <> struct Prec<PREC_BF16> {
  using act_t = __nv_bfloat16;
  static constexpr int KB_ELEMS = 64;
  static __host__ __device__ constexpr uint32_t idesc(int M, int N) { return umma_idesc_bf16(M, N); }
  static __device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) { umma_bf16(d, a, b, id, acc); }
  static __device__ __forceinline__ float load(const act_t* p) { return __bfloat162float(*p); }
  // 16 consecutive channels -> 32 bytes
  static __device__ __forceinline__ void store16(act_t* dst, const float (&v)[16]) {
    const uint32_t pk[8] = {pack_bf16x2(v[0], v[1]),   pack_bf16x2(v[2], v[3]),   pack_bf16x2(v[4], v[5]),
                            pack_bf16x2(v[6], v[7]),   pack_bf16x2(v[8], v[9]),   pack_bf16x2(v[10], v[11]),
                            pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15])};
    st_global_v8(dst, pk);
  }
  // same, for a tensor that no GEMM reads (input of the CUDA-core output convolution): no operand rounding needed
  static __device__ __forceinline__ void store16_exact(act_t* dst, const float (&v)[16]) { store16(dst, v); }
  // 8 consecutive channels
  static __device__ __forceinline__ void load8(const act_t* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(hv[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  }
  static __device__ __forceinline__ void store8(act_t* dst, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                pack_bf16x2(v[6], v[7]));
  }
};
template <> struct Prec<PREC_TF32> {
  using act_t = float;
  static constexpr int KB_ELEMS = 32;
  static __host__ __device__ constexpr uint32_t idesc(int M, int N) { return umma_idesc_tf32(M, N); }
  static __device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) { umma_tf32(d, a, b, id, acc); }
  static __device__ __forceinline__ float load(const act_t* p) { return *p; }
  static __device__ __forceinline__ void store16(act_t* dst, const float (&v)[16]) {
    const uint32_t a[8] = {to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]), to_tf32(v[4]), to_tf32(v[5]), to_tf32(v[6]), to_tf32(v[7])};
    const uint32_t b[8] = {to_tf32(v[8]), to_tf32(v[9]), to_tf32(v[10]), to_tf32(v[11]), to_tf32(v[12]), to_tf32(v[13]), to_tf32(v[14]), to_tf32(v[15])};
    st_global_v8(dst, a);
    st_global_v8(dst + 8, b);
  }
  static __device__ __forceinline__ void store16_exact(act_t* dst, const float (&v)[16]) {
    float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  static __device__ __forceinline__ void load8(const act_t* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store8(act_t* dst, const float (&v)[8]) {
    const uint32_t a[8] = {to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]), to_tf32(v[4]), to_tf32(v[5]), to_tf32(v[6]), to_tf32(v[7])};
    st_global_v8(dst, a);
  }
};

// bias -> LeakyReLU(0.2) -> folded BatchNorm for 16 consecutive GEMM columns (fp32 math)
__device__ __forceinline__ void affine16(const uint32_t (&acc)[16], const float* __restrict__ bias,
                                         const float* __restrict__ scale, const float* __restrict__ shift, bool lrelu,
                                         float (&out)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = reinterpret_cast<const float4*>(bias)[q];      // shared-memory (or global) vectors, 16-byte aligned
    const float4 s = reinterpret_cast<const float4*>(scale)[q];
    const float4 t = reinterpret_cast<const float4*>(shift)[q];
    float a0 = __uint_as_float(acc[4 * q]) + b.x, a1 = __uint_as_float(acc[4 * q + 1]) + b.y;
    float a2 = __uint_as_float(acc[4 * q + 2]) + b.z, a3 = __uint_as_float(acc[4 * q + 3]) + b.w;
    if (lrelu) {
      a0 = a0 >= 0.f ? a0 : 0.2f * a0; a1 = a1 >= 0.f ? a1 : 0.2f * a1;
      a2 = a2 >= 0.f ? a2 : 0.2f * a2; a3 = a3 >= 0.f ? a3 : 0.2f * a3;
    }
    out[4 * q] = a0 * s.x + t.x; out[4 * q + 1] = a1 * s.y + t.y;
    out[4 * q + 2] = a2 * s.z + t.z; out[4 * q + 3] = a3 * s.w + t.w;
  }
}

}  // namespace wdg
