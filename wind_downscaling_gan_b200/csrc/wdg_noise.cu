// On-device Gaussian noise for FlexibleNoiseGenerator (reference data/data_generator.py:319-335, which draws from
// TensorFlow's Philox generator on the compute device): counter-based Philox4x32-10 + Box-Muller, so the
// (B,T,S,S,20) noise tensor -- 87 % of the generator's input bytes -- never crosses PCIe.
// The bit stream is NOT TensorFlow's (its counter/key bookkeeping cannot be verified here); parity runs pass
// explicit noise tensors instead.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/wdg.h"
#include "philox.cuh"

extern int wdg_set_error(const std::string& m);

namespace {
using namespace wdg;

// Each thread produces 4 consecutive normals from one Philox block (counter = offset + block index).
__global__ void noise_normal_kernel(float* __restrict__ out, long long n, float stddev, uint32_t k0, uint32_t k1,
                                    unsigned long long offset) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = blk * 4;
  if (i0 >= n) return;
  const unsigned long long ctr = offset + (unsigned long long)blk;
  float v[4];
  philox_normal4(ctr, k0, k1, v);
  if (i0 + 3 < n && (reinterpret_cast<uintptr_t>(out + i0) & 15) == 0) {
    *reinterpret_cast<float4*>(out + i0) = make_float4(v[0] * stddev, v[1] * stddev, v[2] * stddev, v[3] * stddev);
  } else {
    for (int j = 0; j < 4 && i0 + j < n; ++j) out[i0 + j] = v[j] * stddev;
  }
}

// Same stream, with seed and counter in DEVICE memory (state[0] = key, state[1] = next counter block) so that a
// captured CUDA graph draws fresh noise on every replay: the draw reads the counter, a one-thread kernel advances it.
__global__ void noise_normal_state_kernel(float* __restrict__ out, long long n, float stddev, const unsigned long long* __restrict__ state) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = blk * 4;
  if (i0 >= n) return;
  const unsigned long long seed = state[0], ctr = state[1] + (unsigned long long)blk;
  float v[4];
  philox_normal4(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), v);
  if (i0 + 3 < n && (reinterpret_cast<uintptr_t>(out + i0) & 15) == 0) {
    *reinterpret_cast<float4*>(out + i0) = make_float4(v[0] * stddev, v[1] * stddev, v[2] * stddev, v[3] * stddev);
  } else {
    for (int j = 0; j < 4 && i0 + j < n; ++j) out[i0 + j] = v[j] * stddev;
  }
}
__global__ void rng_advance_kernel(unsigned long long* state, unsigned long long blocks) { state[1] += blocks; }
__global__ void uniform_state_kernel(float* __restrict__ out, long long n, const unsigned long long* __restrict__ state) {
  const long long blk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = blk * 4;
  if (i0 >= n) return;
  const unsigned long long seed = state[0], ctr = state[1] + (unsigned long long)blk;
  const U4 r = philox4x32_10(U4{(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u}, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float v[4] = {u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w)};
  for (int j = 0; j < 4 && i0 + j < n; ++j) out[i0 + j] = v[j];
}

}  // namespace

extern "C" void wdg_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const U4 r = philox4x32_10(U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

extern "C" int wdg_noise_normal(float* out_dev, long long n, float stddev, uint64_t seed, uint64_t offset, void* stream) {
  if (!out_dev || n < 0) return wdg_set_error("bad argument");
  if (n == 0) return 0;
  const long long blocks4 = (n + 3) / 4;
  noise_normal_kernel<<<(unsigned)((blocks4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, n, stddev, (uint32_t)seed,
                                                                                         (uint32_t)(seed >> 32), offset);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return wdg_set_error(std::string("noise_normal_kernel: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" int wdg_rng_advance(uint64_t* state_dev, uint64_t blocks, void* stream) {
  if (!state_dev) return wdg_set_error("null argument");
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)state_dev, (unsigned long long)blocks);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return wdg_set_error(std::string("rng_advance_kernel: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" int wdg_noise_normal_state(float* out_dev, long long n, float stddev, uint64_t* state_dev, void* stream) {
  if (!out_dev || !state_dev || n < 0) return wdg_set_error("bad argument");
  if (n == 0) return 0;
  const long long blocks4 = (n + 3) / 4;
  noise_normal_state_kernel<<<(unsigned)((blocks4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, n, stddev,
                                                                                               (const unsigned long long*)state_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return wdg_set_error(std::string("noise_normal_state_kernel: ") + cudaGetErrorString(e));
  return wdg_rng_advance(state_dev, (uint64_t)blocks4, stream);
}

extern "C" int wdg_uniform_state(float* out_dev, long long n, uint64_t* state_dev, void* stream) {
  if (!out_dev || !state_dev || n < 0) return wdg_set_error("bad argument");
  if (n == 0) return 0;
  const long long blocks4 = (n + 3) / 4;
  uniform_state_kernel<<<(unsigned)((blocks4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, n, (const unsigned long long*)state_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return wdg_set_error(std::string("uniform_state_kernel: ") + cudaGetErrorString(e));
  return wdg_rng_advance(state_dev, (uint64_t)blocks4, stream);
}
