// Microbenchmark: tcgen05.mma (kind::f16, M=128, cta_group::1) issue rate with both operands resident in
// shared memory (no TMA traffic), for several N and A-start alignments.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I wind_downscaling_gan_b200/csrc tools/umma_rate.cu -o /tmp/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace wdg;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int a_shift_rows, int n_acc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smA = smem;                 // 512 rows x 128 B
  uint8_t* smB = smem + 512 * 128;     // 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (512 + 256) * 128 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da0 = umma_desc_kmajor(smem_u32(smA) + a_shift_rows * 128, 128);
    const uint64_t db0 = umma_desc_kmajor(smem_u32(smB), 128);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tm + (it % n_acc) * N;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, da0 + 2 * k, db0 + 2 * k, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) {
      t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N>
void run(int shift, int n_acc, long long* d_out) {
  const int iters = 4096;
  const int smem = (512 + 256) * 128 + 1024;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N><<<148, 128, smem>>>(iters, shift, n_acc, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d shift=%2d n_acc=%d: %s  %.1f cycles/MMA (ideal %d)\n", N, shift, n_acc, cudaGetErrorString(e),
         (double)h / (iters * 4), N / 2);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<256>(0, 1, d_out); run<256>(0, 2, d_out);
  run<128>(0, 1, d_out); run<128>(0, 2, d_out); run<128>(3, 2, d_out);
  run<64>(0, 1, d_out); run<64>(0, 2, d_out); run<64>(1, 2, d_out); run<64>(4, 2, d_out); run<64>(52, 4, d_out);
  run<32>(0, 2, d_out); run<16>(0, 2, d_out);
  return 0;
}
