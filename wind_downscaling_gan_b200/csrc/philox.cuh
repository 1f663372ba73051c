// Counter-based Gaussian noise shared by wdg_noise.cu (noise tensors) and the generator's input-packing kernel (noise
// generated in registers, never stored in fp32): Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.  Counter block j
// under key (k0, k1) yields elements 4j .. 4j+3 of the noise tensor, whichever kernel produces them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wdg {

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

__host__ __device__ inline U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += W0; k1 += W1;
  }
  return c;
}

__device__ inline float u32_to_unit(uint32_t x) {   // 23 mantissa bits -> [0, 1)
  return __uint_as_float((x & 0x7fffffu) | 0x3f800000u) - 1.0f;
}

// Four standard normals of counter block `ctr` (Box-Muller on the SFU: __logf / __sincosf with the angle kept in
// [-pi, pi), where their absolute error is 2^-21.4; the radicand is clamped at 0 because __logf can come out a hair
// positive next to u = 1).  The accurate libm forms made the input-packing kernel of the generator compute-bound.
__device__ inline void philox_normal4(unsigned long long ctr, uint32_t k0, uint32_t k1, float (&v)[4]) {
  const U4 r = philox4x32_10(U4{(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u}, k0, k1);
  const float rad = sqrtf(fmaxf(-2.0f * __logf(fmaxf(u32_to_unit(r.x), 1.0e-7f)), 0.f));
  float s, c;
  __sincosf(6.283185307179586f * u32_to_unit(r.y) - 3.141592653589793f, &s, &c);
  v[0] = s * rad; v[1] = c * rad;
  const float rad2 = sqrtf(fmaxf(-2.0f * __logf(fmaxf(u32_to_unit(r.z), 1.0e-7f)), 0.f));
  __sincosf(6.283185307179586f * u32_to_unit(r.w) - 3.141592653589793f, &s, &c);
  v[2] = s * rad2; v[3] = c * rad2;
}

}  // namespace wdg
