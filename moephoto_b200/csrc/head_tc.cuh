// The two Conv3x3(F,1) heads (models.py:130,140,151,162), their sum, the fp16 rounding, the seam blend
// and the canvas store (imageProcess.py:164-170) in one bandwidth-bound kernel.
//
// A 64->1 3x3 convolution is rewritten as  out(Y,X) = sum_{dy,dx} P_{dy,dx}(Y+dy-1, X+dx-1)  with
// P^b_t(y,x) = <w_b[t], B(y,x,:)>, b = u, r — nine 64-long dot products per INPUT pixel and per branch, kept apart
// because the reference's half model rounds each head to fp16 before adding them (models.py:38).  The dot products are a GEMM: [128 pixels x 64] x [64 x 16] (9 taps padded to N=16), so the
// tensor cores produce P from the raw NHWC rows with 8 tcgen05.mma per row (2 branches x 4 K-steps,
// no shifted views), each activation is read from HBM exactly once, and the 3x3 stencil collapses to
// 9 fp32 adds per pixel on the CUDA cores: the horizontal part through a small smem exchange, the vertical
// part in two running registers while the persistent CTA walks down its 126-column strip.
// Warp roles: 0 = TMA producer (U row + R row per step), 1 = MMA issuer, 2..5 = stencil/blend/store.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include "ptx.cuh"
#include "kernels_simt.cuh"

namespace moe {

struct HeadTcParams {
  HeadParams g;            // geometry, seam and canvas (same meaning as the SIMT kernel; u/r/wu/wr unused)
  const uint8_t* w_img;    // [2 branches][16 rows][128 B] swizzled fp16: row t < 9 = tap t, rows 9..15 zero
  int strips, nseg, seg_rows, items;
};

struct HeadMaps {
  CUtensorMap u, r;        // (64, W, H, N), box (64,128,1,1)
};

constexpr int kHeadStripOut = 126;                 // output columns per strip (128 loaded, 1-px halo each side)
constexpr int kHeadThreads = 192;
constexpr int kHeadSlots = 5;
constexpr uint32_t kHeadSlotBytes = 2 * kStageBytes;   // U row + R row
constexpr int kHeadAccStages = 4;
constexpr uint32_t kHeadExFloats = 2 * 9 * 130;       // one exchange buffer: [branch][tap][130]
constexpr uint32_t kHeadTmemCols = kHeadAccStages * 32;   // per stage: 16 columns (9 taps used) for each branch
constexpr uint32_t kHeadSmemBytes = 1024 + kHeadSlots * kHeadSlotBytes + 4096 + 2 * kHeadExFloats * 4 + 1024;

__device__ __forceinline__ void head_decode_item(const HeadTcParams& p, int item, int& n, int& xl, int& y0, int& y1) {
  const int seg = item % p.nseg;
  int rest = item / p.nseg;
  const int strip = rest % p.strips;
  n = rest / p.strips;
  xl = strip * kHeadStripOut - 1;                  // first loaded column
  y0 = seg * p.seg_rows;
  y1 = min(p.g.H, y0 + p.seg_rows);
}

__global__ void __launch_bounds__(kHeadThreads, 1)
head_tc_kernel(const __grid_constant__ HeadMaps maps, const HeadTcParams p)
{
  constexpr int S = kHeadSlots, AS = kHeadAccStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ring = base;
  const uint32_t wsm = ring + S * kHeadSlotBytes;          // 4 KB of weights
  const uint32_t exs = wsm + 4096;                         // exchange buffers [2][2 branches][9][130] float
  const uint32_t bars = exs + 2 * kHeadExFloats * 4;
  const uint32_t barsa = (bars + 7u) & ~7u;
  const uint32_t full = barsa, empty = full + 8 * S, tfull = empty + 8 * S, tempty = tfull + 8 * AS;
  const uint32_t wbar = tempty + 8 * AS, tslot = wbar + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  float* ex = reinterpret_cast<float*>(smem + (exs - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(full + 8 * i, 1); ptx::mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) { ptx::mbar_init(tfull + 8 * i, 1); ptx::mbar_init(tempty + 8 * i, 128); }
    ptx::mbar_init(wbar, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&maps.u);
    ptx::prefetch_tmap(&maps.r);
  }
  if (warp == 1) ptx::tmem_alloc(tslot, kHeadTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tslot_ptr;

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues (keeps TMA operands in uniform registers)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, 4096);
      ptx::bulk_load_1d(wsm, p.w_img, 4096, wbar);
    }
    __syncwarp();
    uint32_t ld = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, xl, y0, y1;
      head_decode_item(p, item, n, xl, y0, y1);
      for (int yy = y0 - 1; yy <= y1; ++yy, ++ld) {
        const uint32_t slot = ld % S;
        ptx::mbar_wait(empty + 8 * slot, ((ld / S) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(full + 8 * slot, kHeadSlotBytes);
          ptx::tma_load_4d(ring + slot * kHeadSlotBytes, &maps.u, full + 8 * slot, 0, xl, yy, n);
          ptx::tma_load_4d(ring + slot * kHeadSlotBytes + kStageBytes, &maps.r, full + 8 * slot, 0, xl, yy, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::idesc_f16_f32(128, 16);
    const uint64_t adesc0 = ptx::smem_desc_sw128(ring, 1024, 0);
    const uint64_t bdesc0 = ptx::smem_desc_sw128(wsm, 1024, 0);
    ptx::mbar_wait(wbar, 0);
    ptx::tc_fence_after_sync();
    uint32_t ld = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, xl, y0, y1;
      head_decode_item(p, item, n, xl, y0, y1);
      for (int yy = y0 - 1; yy <= y1; ++yy, ++ld) {
        const uint32_t slot = ld % S, stage = ld % AS;
        ptx::mbar_wait(full + 8 * slot, (ld / S) & 1);
        ptx::mbar_wait(tempty + 8 * stage, ((ld / AS) & 1) ^ 1);
        ptx::tc_fence_after_sync();
        const uint64_t arow = adesc0 + static_cast<uint64_t>(slot * (kHeadSlotBytes >> 4));
        if (ptx::elect_one()) {
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_f16_ss(tmem_base + stage * 32 + b * 16, arow + (b * (kStageBytes >> 4) + k * 2), bdesc0 + (b * 128 + k * 2), idesc, k != 0);
          ptx::mma_commit(tfull + 8 * stage);
          ptx::mma_commit(empty + 8 * slot);
        }
        __syncwarp();
      }
    }
    if (ptx::elect_one()) ptx::mma_commit(wbar);
    __syncwarp();
    ptx::mbar_wait_drain(wbar, 1);
  } else {
    const int lgrp = warp & 3;
    const int L = lgrp * 32 + lane;                 // loaded pixel index 0..127
    const HeadParams& g = p.g;
    uint32_t ld = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, xl, y0, y1;
      head_decode_item(p, item, n, xl, y0, y1);
      const int X = xl + L;                          // column within the computed rectangle
      const int cx = g.ox + X;
      const bool col_ok = L >= 1 && L <= kHeadStripOut && X < g.W && cx >= g.keep_x0 && cx < g.keep_x1;
      float acc_a[2] = {0.f, 0.f}, prev_h0[2] = {0.f, 0.f};
      for (int yy = y0 - 1; yy <= y1; ++yy, ++ld) {
        const uint32_t stage = ld % AS;
        ptx::mbar_wait(tfull + 8 * stage, (ld / AS) & 1);
        ptx::tc_fence_after_sync();
        uint32_t v[2][16];
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + stage * 32, v[0]);
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + stage * 32 + 16, v[1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(tempty + 8 * stage);
        float* eb = ex + (ld & 1) * kHeadExFloats;
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int t = 0; t < 9; ++t) eb[(b * 9 + t) * 130 + L + 1] = __uint_as_float(v[b][t]);
        ptx::named_bar_sync(1, 128);
        float outv[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const float* e2 = eb + b * 9 * 130;
          float h[3];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
            h[dy] = (e2[(dy * 3 + 0) * 130 + L] + e2[(dy * 3 + 1) * 130 + L + 1]) + e2[(dy * 3 + 2) * 130 + L + 2];
          outv[b] = acc_a[b] + h[2];                 // row yy-1 is complete: h0(yy-2) + h1(yy-1) + h2(yy)
          acc_a[b] = prev_h0[b] + h[1];
          prev_h0[b] = h[0];
        }
        const int Y = yy - 1;
        const int cy = g.oy + Y;
        if (col_ok && Y >= y0 && cy >= g.keep_y0 && cy < g.keep_y1) {
          // each head is an fp16 tensor in the reference's half model, their sum a third op (models.py:38)
          float val = h_round(h_round(outv[0]) + h_round(outv[1]));
          __half* dst = g.canvas + n * g.plane_stride + static_cast<int64_t>(cy) * g.row_stride + cx;
          if (cy < g.blend_y1 || cx < g.blend_x1) {
            const float old = __half2float(*dst);
            if (cy < g.blend_y1) val = h_round(old + h_round(g.ramp[cy - g.ramp_y0] * h_round(val - old)));
            if (cx < g.blend_x1) val = h_round(old + h_round(g.ramp[cx - g.ramp_x0] * h_round(val - old)));
          }
          *dst = __float2half_rn(val);
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, kHeadTmemCols);
  }
}

}  // namespace moe
