// Microbenchmark for the next step of the 64 -> 64 trunk convolution: the A operand in TENSOR memory.
// The N = 64 tile is operand-fetch bound with both operands in shared memory (48.5 cycles per 128x64x16 MMA against 32
// of math, profiles/r01_mma_rate_microbench.log).  With A in TMEM only B (2 KB) is fetched per MMA; the price is
// copying every input row into TMEM three times (the dx = -1, 0, +1 views; TMEM lanes cannot be shifted) with
// tcgen05.cp.128x256b: 12 copies of 4 KB per row against 36 MMAs.
//   part 1 (correctness): shifted SWIZZLE_128B views of a 130-px row -> tcgen05.cp -> .ts MMA must equal the .ss MMA
//                         on the same views, and both the scalar reference;
//   part 2 (rate): cycles per MMA for .ss, .ts without copies, .ts with the 12 copies per row interleaved.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ts mma_ts.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../moephoto_b200/csrc/ptx.cuh"
using namespace moe;

__device__ __forceinline__ float a_val(int p, int c) { return static_cast<float>((p * 5 + c * 3) % 13 - 6) * 0.125f; }
__device__ __forceinline__ float b_val(int n, int c) { return static_cast<float>((n * 7 + c) % 11 - 5) * 0.25f; }

// [row][64 fp16] with the 128-byte swizzle of a 1024-aligned tile: 16-byte group g of row r lives at group g ^ (r & 7)
__device__ __forceinline__ void put_sw128(uint8_t* tile, int row, int c, float v) {
  *reinterpret_cast<__half*>(tile + row * 128 + ((((c >> 3) ^ (row & 7)) << 4) | ((c & 7) * 2))) = __float2half(v);
}

__global__ void __launch_bounds__(128, 1) check_kernel(int* bad)   // bad[0..2]: ss mismatches per dx, bad[3..5]: ts mismatches
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_sm = base, b_sm = base + 17408, bar = b_sm + 8192, tslot = bar + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  for (int i = threadIdx.x; i < 130 * 64; i += blockDim.x) put_sw128(smem, i >> 6, i & 63, a_val(i >> 6, i & 63));
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) put_sw128(smem + 17408, i >> 6, i & 63, b_val(i >> 6, i & 63));
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  ptx::fence_proxy_async_smem();
  if (threadIdx.x < 32) ptx::tmem_alloc(tslot, 512);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tslot_ptr;
  constexpr uint32_t kDss = 0, kDts = 192, kA = 384;           // TMEM columns
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::idesc_f16_f32(128, 64);
    const uint64_t a0 = ptx::smem_desc_sw128(a_sm, 1024, 0), b0 = ptx::smem_desc_sw128(b_sm, 1024, 0);
    if (ptx::elect_one()) {
      for (int dx = 0; dx < 3; ++dx) {
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem + kDss + dx * 64, a0 + dx * 8 + k * 2, b0 + k * 2, idesc, k != 0);
        for (int k = 0; k < 4; ++k) ptx::tmem_cp_128x256b(tmem + kA + dx * 32 + k * 8, a0 + dx * 8 + k * 2);
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ts(tmem + kDts + dx * 64, tmem + kA + dx * 32 + k * 8, b0 + k * 2, idesc, k != 0);
      }
      ptx::mma_commit(bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, p = threadIdx.x;          // TMEM lane = pixel
  for (int dx = 0; dx < 3; ++dx) {
    for (int h = 0; h < 2; ++h) {
      uint32_t vs[32], vt[32];
      ptx::tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + kDss + dx * 64 + h * 32, vs);
      ptx::tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + kDts + dx * 64 + h * 32, vt);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) {
        const int n = h * 32 + j;
        float want = 0.f;
        for (int c = 0; c < 64; ++c) want += a_val(p + dx, c) * b_val(n, c);
        if (__uint_as_float(vs[j]) != want) atomicAdd(bad + dx, 1);
        if (__uint_as_float(vt[j]) != want) atomicAdd(bad + 3 + dx, 1);
        if (p == 5 && n == 3) printf("dx=%d p=5 n=3: want %f ss %f ts %f\n", dx, want, __uint_as_float(vs[j]), __uint_as_float(vt[j]));
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

// MODE 0: .ss   1: .ts, A never refreshed   2: .ts + 12 tcgen05.cp per 36 MMAs
template <int MODE>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* cycles, int iters)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_sm = base, b_sm = base + 3 * 17408, bar = b_sm + 9 * 8192, tslot = bar + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  for (uint32_t i = threadIdx.x; i < (3 * 17408 + 9 * 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  ptx::fence_proxy_async_smem();
  if (threadIdx.x < 32) ptx::tmem_alloc(tslot, 512);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tslot_ptr;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::idesc_f16_f32(128, 64);
    const uint64_t a0 = ptx::smem_desc_sw128(a_sm, 1024, 0), b0 = ptx::smem_desc_sw128(b_sm, 1024, 0);
    if (MODE != 0 && ptx::elect_one()) {        // fill the four row buffers once (4 rows x 3 views x 32 columns = TMEM columns 0..383)
      for (int r = 0; r < 4; ++r)
        for (int dx = 0; dx < 3; ++dx)
          for (int k = 0; k < 4; ++k) ptx::tmem_cp_128x256b(tmem + r * 96 + dx * 32 + k * 8, a0 + (r % 3) * 1088 + dx * 8 + k * 2);
    }
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + 384 + (it & 1) * 64;
      if (ptx::elect_one()) {
        if (MODE == 2) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::tmem_cp_128x256b(tmem + (it & 3) * 96 + dx * 32 + k * 8, a0 + (it % 3) * 1088 + dx * 8 + k * 2);
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (MODE == 0) ptx::mma_f16_ss(d, a0 + dy * 1088 + dx * 8 + k * 2, b0 + (dy * 3 + dx) * 512 + k * 2, idesc, (dy | dx | k) != 0);
              else ptx::mma_f16_ts(d, tmem + ((it + 1 + dy) & 3) * 96 + dx * 32 + k * 8, b0 + (dy * 3 + dx) * 512 + k * 2, idesc, (dy | dx | k) != 0);
            }
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

// MODE 3 of the rate test: four loader warps copy every input row into TMEM themselves (ld.shared of the swizzled row,
// three dx views, tcgen05.st) one row ahead of the MMA warp; afull / aempty mbarriers hand the 4 TMEM row buffers over.
// Also verifies the result of the last row against the .ss MMA (bad[0] = mismatches).
template <int GROUPS, int WHAT = 3>   // GROUPS 1: four loader warps copy all three views; 3: twelve loader warps, one view each.  WHAT bit 0: ld.shared, bit 1: tcgen05.st
__global__ void __launch_bounds__(32 + 128 * GROUPS, 1) st_rate_kernel(long long* cycles, int iters, int* bad)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_sm = base, b_sm = base + 3 * 17408, bars = b_sm + 9 * 8192;
  const uint32_t afull = bars, aempty = bars + 32, done = bars + 64, tslot = bars + 72;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  for (int i = threadIdx.x; i < 3 * 130 * 64; i += blockDim.x) {
    const int r = i / (130 * 64), pc = i % (130 * 64);
    put_sw128(smem + r * 17408, pc >> 6, pc & 63, a_val((pc >> 6) + r, pc & 63));
  }
  for (int i = threadIdx.x; i < 9 * 64 * 64; i += blockDim.x) {
    const int t = i / 4096, nc = i % 4096;
    put_sw128(smem + 3 * 17408 + t * 8192, nc >> 6, nc & 63, b_val((nc >> 6) + t, nc & 63));
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(afull + 8 * i, 4 * GROUPS); ptx::mbar_init(aempty + 8 * i, 1); }
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  ptx::fence_proxy_async_smem();
  if (threadIdx.x < 32) ptx::tmem_alloc(tslot, 512);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tslot_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t idesc = ptx::idesc_f16_f32(128, 64);
  const uint64_t a0 = ptx::smem_desc_sw128(a_sm, 1024, 0), b0 = ptx::smem_desc_sw128(b_sm, 1024, 0);
  if (warp == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // rows it-2, it-1, it must be in TMEM (rows < 0 are never loaded: skip their taps)
      ptx::mbar_wait(afull + 8 * (it & 3), (it >> 2) & 1);
      ptx::tc_fence_after_sync();
      const uint32_t d = tmem + 384 + (it & 1) * 64;
      if (ptx::elect_one()) {
        bool first = true;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int row = it - 2 + dy;
          if (row < 0) continue;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ptx::mma_f16_ts(d, tmem + (row & 3) * 96 + dx * 32 + k * 8, b0 + (dy * 3 + dx) * 512 + k * 2, idesc, first ? 0u : 1u);
              first = false;
            }
        }
        // buffer of row it-2 is free once these MMAs completed
        if (it >= 2) ptx::mma_commit(aempty + 8 * ((it - 2) & 3));
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(done);
    __syncwarp();
    ptx::mbar_wait(done, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  } else {
    const int quad = warp & 3, L = quad * 32 + lane;          // TMEM lane = pixel
    const int group = (warp - 1) >> 2;
    for (int it = 0; it < iters; ++it) {
      const int b = it & 3;
      if (it >= 4) ptx::mbar_wait(aempty + 8 * b, ((it - 4) >> 2) & 1);    // row it-4 (same buffer) has been consumed
      ptx::tc_fence_after_sync();
      const uint8_t* slot = smem + (it % 3) * 17408;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        if (GROUPS == 3 && dx != group) continue;
        const int p = L + dx;
        uint32_t v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint4 q = make_uint4(it, lane, g, dx);
          if (WHAT & 1) q = *reinterpret_cast<const uint4*>(slot + p * 128 + ((g ^ (p & 7)) << 4));
          v[g * 4] = q.x; v[g * 4 + 1] = q.y; v[g * 4 + 2] = q.z; v[g * 4 + 3] = q.w;
        }
        if (WHAT & 2) ptx::tmem_st32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + b * 96 + dx * 32, v);
        else if (v[0] + v[7] + v[13] + v[31] == 0x12345u) cycles[1] = 0;      // keep the loads alive
      }
      if (WHAT & 2) ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(afull + 8 * b);
    }
  }
  __syncthreads();
  ptx::tc_fence_after_sync();
  // check the last accumulator against the scalar reference (rows it-2..it of the last iteration)
  if (warp >= 1 && warp <= 4 && bad && WHAT == 3) {
    const int quad = warp & 3, p = quad * 32 + lane, it = iters - 1;
    for (int h = 0; h < 2; ++h) {
      uint32_t vs[32];
      ptx::tmem_ld32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + 384 + (it & 1) * 64 + h * 32, vs);
      ptx::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) {
        const int n = h * 32 + j;
        float want = 0.f;
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx)
            for (int c = 0; c < 64; ++c) want += a_val(p + dx + (it - 2 + dy) % 3, c) * b_val(n + dy * 3 + dx, c);
        if (__uint_as_float(vs[j]) != want) atomicAdd(bad, 1);
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

template <int GROUPS, int WHAT = 3> void run_st(int grid)
{
  long long* d; cudaMalloc(&d, grid * sizeof(long long));
  int* bad; cudaMalloc(&bad, sizeof(int)); cudaMemset(bad, 0, sizeof(int));
  const int smem = 1024 + 3 * 17408 + 9 * 8192 + 128;
  cudaFuncSetAttribute(st_rate_kernel<GROUPS, WHAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    st_rate_kernel<GROUPS, WHAT><<<grid, 32 + 128 * GROUPS, smem>>>(d, iters, bad);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[256]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    int hb = 0; cudaMemcpy(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < grid; ++i) cyc += h[i]; cyc /= grid;
    const double mmas = 36.0 * iters;
    printf("%-44s grid=%3d rep=%d  %s  %.1f cyc/MMA (math 32)  %.3f ms  %.0f TFLOP/s  mismatches %d of %d\n", WHAT == 1 ? "N=64 .ts, 4 warps ld.shared only (no tcgen05.st)" : WHAT == 2 ? "N=64 .ts, 4 warps tcgen05.st only (no ld.shared)" : GROUPS == 1 ? "N=64 .ts, rows copied by 4 warps (tcgen05.st)" : "N=64 .ts, rows copied by 12 warps (tcgen05.st)", grid, rep,
           cudaGetErrorString(e), cyc / mmas, ms, 2.0 * 128 * 64 * 16 * mmas * grid / ms / 1e9, hb, 8192 * grid * (rep + 1));
    if (e != cudaSuccess) break;
  }
  cudaFree(d); cudaFree(bad);
}

template <int MODE> void run(int grid, const char* name)
{
  long long* d; cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 1024 + 3 * 17408 + 9 * 8192 + 64;
  cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    rate_kernel<MODE><<<grid, 64, smem>>>(d, iters);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[256]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < grid; ++i) cyc += h[i]; cyc /= grid;
    const double mmas = 36.0 * iters;
    printf("%-44s grid=%3d rep=%d  %s  %.1f cyc/MMA (math 32)  %.3f ms  %.0f TFLOP/s\n", name, grid, rep, cudaGetErrorString(e), cyc / mmas, ms,
           2.0 * 128 * 64 * 16 * mmas * grid / ms / 1e9);
    if (e != cudaSuccess) break;
  }
  cudaFree(d);
}

int main()
{
  int* bad; cudaMalloc(&bad, 6 * sizeof(int)); cudaMemset(bad, 0, 6 * sizeof(int));
  const int smem = 1024 + 17408 + 8192 + 64;
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  check_kernel<<<1, 128, smem>>>(bad);
  cudaError_t e = cudaDeviceSynchronize();
  int h[6]; cudaMemcpy(h, bad, sizeof h, cudaMemcpyDeviceToHost);
  printf("check: %s  mismatches of 8192 per view  .ss dx=0/1/2: %d %d %d   tcgen05.cp + .ts dx=0/1/2: %d %d %d\n", cudaGetErrorString(e), h[0], h[1], h[2], h[3],
         h[4], h[5]);
  if (e != cudaSuccess) return 1;
  run<0>(148, "N=64 .ss (both operands in smem)");
  run<1>(148, "N=64 .ts (A resident in TMEM)");
  run<2>(148, "N=64 .ts + 12 tcgen05.cp.128x256b per row");
  run<2>(1, "N=64 .ts + copies, single CTA");
  run_st<1>(148);
  run_st<3>(148);
  run_st<1, 1>(148);
  run_st<1, 2>(148);
  return 0;
}
