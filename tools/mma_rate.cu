// Microbenchmark: issue rate of tcgen05.mma kind::f16, M=128, K=16, both operands in shared memory
// (SWIZZLE_128B K-major), for N = 64 / 128 / 256, 1-CTA group, one CTA per SM.  Answers: is the N=64
// tile of conv3x3_tc_kernel capped by operand fetch?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../moephoto_b200/csrc/ptx.cuh"
using namespace moe;

template <int N, int DX>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* cycles, int iters)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_sm = base;                 // 3 row slots of 17 KB
  const uint32_t b_sm = base + 3 * 17408;     // 9 taps x N rows x 128 B
  constexpr int TAPS = N == 256 ? 4 : 9;   // N=256: 9 tap tiles do not fit, reuse 4 (same traffic per MMA)
  const uint32_t bar = b_sm + TAPS * N * 128;
  const uint32_t tslot = bar + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  for (uint32_t i = threadIdx.x; i < (3 * 17408 + TAPS * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  ptx::fence_proxy_async_smem();
  if (threadIdx.x < 32) ptx::tmem_alloc(tslot, 512);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tslot_ptr;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::idesc_f16_f32(128, N);
    const uint64_t a0 = ptx::smem_desc_sw128(a_sm, 1024, 0), b0 = ptx::smem_desc_sw128(b_sm, 1024, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (it & 1) * N;
      if (ptx::elect_one()) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16_ss(d, a0 + dy * 1088 + (DX ? dx * 8 : 0) + k * 2, b0 + ((dy * 3 + dx) % TAPS) * (N * 8) + k * 2, idesc, (dy | dx | k) != 0);
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(bar);
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

template <int N, int DX> void run(int grid, const char* name)
{
  long long* d; cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 1024 + 3 * 17408 + (N == 256 ? 4 : 9) * N * 128 + 64;
  cudaFuncSetAttribute(rate_kernel<N, DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    rate_kernel<N, DX><<<grid, 64, smem>>>(d, iters);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[256]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < grid; ++i) cyc += h[i]; cyc /= grid;
    const double mmas = 36.0 * iters;
    const double flops = 2.0 * 128 * N * 16 * mmas * grid;
    printf("%-22s grid=%3d rep=%d  %s  %.1f cyc/MMA (ideal %d)  %.3f ms  %.0f TFLOP/s\n", name, grid, rep, cudaGetErrorString(e), cyc / mmas, N / 2, ms, flops / ms / 1e9);
  }
  cudaFree(d);
}

int main()
{
  run<64, 1>(148, "N=64  shifted views");
  run<64, 0>(148, "N=64  aligned views");
  run<128, 1>(148, "N=128 shifted views");
  run<256, 1>(148, "N=256 shifted views");
  run<64, 1>(1, "N=64  single CTA");
  run<256, 1>(1, "N=256 single CTA");
  return 0;
}
