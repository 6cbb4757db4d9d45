// conv3x3, 64 -> 64*r*r channels, NHWC fp16, as an implicit GEMM on the 5th-gen tensor cores.
//
// One persistent CTA per SM walks DOWN a 128-pixel-wide column strip of one image plane:
//   * M tile   = one output row segment of 128 pixels (TMEM lane = pixel, column = output channel)
//   * K        = 9 taps x 64 input channels = 36 tcgen05.mma of 128 x (64*NCH) x 16
//   * A operand= the input rows y-1, y, y+1, each TMA-loaded ONCE (130 px x 128 B, 128-byte swizzle)
//                into a ring of row slots; the nine taps are nine *views* of those three slots: the
//                smem descriptor's start address is moved by dx*128 B (next pixel) and k*32 B (next
//                16 channels) — no im2col copy, no re-read of a row for its three vertical uses.
//                (Measured on B200, profiles/r01_diag_first_contact.log: the 128B swizzle is a function
//                of the absolute smem address, so a start address that is not 1024-aligned needs NO
//                base_offset in the descriptor; setting base_offset = (addr>>7)&7 gives garbage.)
//   * B operand= the layer's weights for NCH 64-channel output chunks, resident in smem for the whole
//                kernel (pre-swizzled on the host, fetched with cp.async.bulk).
//   * zero padding at the tile border is TMA out-of-bounds fill (x = -1, x = W, y = -1, y = H).
//   * epilogue = 4 warps: tcgen05.ld -> PReLU | x scale + skip | + bias, PReLU -> fp16 -> 128-byte pixel
//                stores; PixelShuffle is only an address computation (chunk q = (i,j) -> pixel
//                (y*r+i, x*r+j)), see weights.py for the matching output-channel permutation.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include "ptx.cuh"

namespace moe {

enum ConvEpilogue : int { EPI_PLAIN = 0, EPI_PRELU = 1, EPI_SCALE_SKIP = 2, EPI_BIAS_PRELU = 3 };

struct ConvParams {
  const uint8_t* w_img;   // [r*r chunks][9 taps][64 rows][128 B], 128B-swizzled fp16
  const float* bias;      // [r*r][64] or nullptr
  const __half* in;       // NHWC (N,H,W,64)   (tcgen05 path reads it through the tensor map)
  __half* out;            // NHWC (N,H*r,W*r,64)
  const __half* skip;     // NHWC (N,H,W,64) when epi == EPI_SCALE_SKIP
  int N, H, W;
  int r;                  // 1, 2 or 3
  int epi;
  float param;            // PReLU slope or residual scale
  int strips, nseg, seg_rows, items;
};

constexpr int kStripW = 128;
constexpr int kRowPx = kStripW + 2;
constexpr uint32_t kRowBytes = kRowPx * 128;          // 16640, the TMA box
constexpr uint32_t kSlotBytes = 17 * 1024;            // row slot, 1024-aligned for the swizzle pattern
constexpr uint32_t kChunkImgBytes = 9 * 64 * 128;     // 73728
constexpr int kConvThreads = 192;

template <int NCH> struct ConvCfg {
  static constexpr int kSlots = NCH == 1 ? 8 : 4;
  static constexpr int kAccStages = 4;
  static constexpr int kN = 64 * NCH;
  static constexpr uint32_t kTmemCols = kAccStages * kN;
  static constexpr uint32_t kWBytes = NCH * kChunkImgBytes;
  static constexpr uint32_t kBarBytes = 1024;
  static constexpr uint32_t kSmemBytes = 1024 + kSlots * kSlotBytes + kWBytes + kBarBytes;
};

__device__ __forceinline__ void conv_decode_item(const ConvParams& p, int item, int ncg, int& n, int& x0, int& y0, int& y1) {
  int rest = item / ncg;
  const int seg = rest % p.nseg;
  rest /= p.nseg;
  const int strip = rest % p.strips;
  n = rest / p.strips;
  x0 = strip * kStripW;
  y0 = seg * p.seg_rows;
  y1 = min(p.H, y0 + p.seg_rows);
}

__device__ __forceinline__ float epi_apply(float v, int epi, float param, float bias, float skip) {
  if (epi == EPI_PRELU) return v >= 0.f ? v : param * v;
  if (epi == EPI_SCALE_SKIP) return fmaf(v, param, skip);
  if (epi == EPI_BIAS_PRELU) { v += bias; return v >= 0.f ? v : param * v; }
  return v;
}

template <int NCH>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap in_map, const ConvParams p)
{
  using Cfg = ConvCfg<NCH>;
  constexpr int S = Cfg::kSlots, AS = Cfg::kAccStages, NN = Cfg::kN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ring = base;
  const uint32_t wsm = base + S * kSlotBytes;
  const uint32_t bars = wsm + Cfg::kWBytes;
  const uint32_t full = bars, empty = bars + 8 * S, tfull = bars + 16 * S, tempty = tfull + 8 * AS;
  const uint32_t wbar = tempty + 8 * AS;
  const uint32_t tslot = wbar + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = p.r * p.r;
  const int ncg = nchunks / NCH;                 // chunk groups; gridDim.x % ncg == 0 (host guarantees)
  const int cg = blockIdx.x % ncg;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(full + 8 * i, 1); ptx::mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) { ptx::mbar_init(tfull + 8 * i, 1); ptx::mbar_init(tempty + 8 * i, 128); }
    ptx::mbar_init(wbar, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&in_map);
  }
  if (warp == 1) ptx::tmem_alloc(tslot, Cfg::kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tslot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_expect_tx(wbar, Cfg::kWBytes);
      for (int c = 0; c < NCH; ++c) {
        const uint8_t* src = p.w_img + static_cast<size_t>(cg * NCH + c) * kChunkImgBytes;
        for (int tap = 0; tap < 9; ++tap)
          ptx::bulk_load_1d(wsm + tap * (NCH * 8192) + c * 8192, src + tap * 8192, 8192, wbar);
      }
      uint32_t ld = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int n, x0, y0, y1;
        conv_decode_item(p, item, ncg, n, x0, y0, y1);
        for (int yy = y0 - 1; yy <= y1; ++yy, ++ld) {
          const uint32_t slot = ld % S;
          ptx::mbar_wait(empty + 8 * slot, ((ld / S) & 1) ^ 1);
          ptx::mbar_expect_tx(full + 8 * slot, kRowBytes);
          ptx::tma_load_4d(ring + slot * kSlotBytes, &in_map, full + 8 * slot, 0, x0 - 1, yy, n);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_f16_f32(128, NN);
      ptx::mbar_wait(wbar, 0);
      ptx::tc_fence_after_sync();
      uint32_t cons = 0, acc = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int n, x0, y0, y1;
        conv_decode_item(p, item, ncg, n, x0, y0, y1);
        const int nrows = y1 - y0;
        ptx::mbar_wait(full + 8 * (cons % S), (cons / S) & 1);
        ptx::mbar_wait(full + 8 * ((cons + 1) % S), ((cons + 1) / S) & 1);
        for (int j = 0; j < nrows; ++j) {
          const uint32_t l2 = cons + j + 2;
          ptx::mbar_wait(full + 8 * (l2 % S), (l2 / S) & 1);
          const uint32_t stage = acc % AS;
          ptx::mbar_wait(tempty + 8 * stage, ((acc / AS) & 1) ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + stage * NN;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t abase = ring + ((cons + j + dy) % S) * kSlotBytes;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const uint32_t a0 = abase + dx * 128;
              const uint32_t b0 = wsm + (dy * 3 + dx) * (NCH * 8192);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                ptx::mma_f16_ss(d_tmem, ptx::smem_desc_sw128(a0 + k * 32, 1024, 0),
                                ptx::smem_desc_sw128(b0 + k * 32, 1024, 0), idesc, (dy | dx | k) != 0);
              }
            }
          }
          ptx::mma_commit(tfull + 8 * stage);               // accumulator ready for the epilogue
          ptx::mma_commit(empty + 8 * ((cons + j) % S));    // input row y-1 is no longer needed
          ++acc;
        }
        ptx::mma_commit(empty + 8 * ((cons + nrows) % S));
        ptx::mma_commit(empty + 8 * ((cons + nrows + 1) % S));
        cons += nrows + 2;
      }
      // drain: commits complete in order, so once this one lands no arrive is still in flight
      ptx::mma_commit(wbar);
      ptx::mbar_wait(wbar, 1);
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int lgrp = warp & 3;                       // TMEM lanes this warp may read: 32*lgrp ..
    const int L = lgrp * 32 + lane;
    const int Ho = p.H * p.r, Wo = p.W * p.r;
    uint32_t acc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, x0, y0, y1;
      conv_decode_item(p, item, ncg, n, x0, y0, y1);
      const int x = x0 + L;
      const bool valid = x < p.W;
      for (int y = y0; y < y1; ++y, ++acc) {
        const uint32_t stage = acc % AS;
        ptx::mbar_wait(tfull + 8 * stage, (acc / AS) & 1);
        ptx::tc_fence_after_sync();
        const size_t ipix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int chunk = cg * NCH + c;
          const int sy = chunk / p.r, sx = chunk - sy * p.r;
          const size_t opix = (static_cast<size_t>(n) * Ho + (y * p.r + sy)) * Wo + (x * p.r + sx);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + stage * NN + c * 64 + h * 32, v);
            ptx::tmem_ld_wait();
            if (c == NCH - 1 && h == 1) {            // last read of this stage: hand TMEM back to the MMA warp
              ptx::tc_fence_before_sync();
              ptx::mbar_arrive(tempty + 8 * stage);
            }
            if (valid) {
              uint4 sk[4];
              if (p.epi == EPI_SCALE_SKIP) {
                const uint4* sp = reinterpret_cast<const uint4*>(p.skip + ipix * 64 + h * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) sk[q] = sp[q];   // plain loads: `skip` may alias `out`
              }
              uint4 o[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = q * 8 + e * 2;
                  float b0 = 0.f, b1 = 0.f, s0 = 0.f, s1 = 0.f;
                  if (p.epi == EPI_BIAS_PRELU) {
                    b0 = __ldg(p.bias + chunk * 64 + h * 32 + j);
                    b1 = __ldg(p.bias + chunk * 64 + h * 32 + j + 1);
                  }
                  if (p.epi == EPI_SCALE_SKIP) {
                    const uint32_t sw = reinterpret_cast<const uint32_t*>(&sk[q])[e];
                    const __half2 hs = *reinterpret_cast<const __half2*>(&sw);
                    s0 = __low2float(hs);
                    s1 = __high2float(hs);
                  }
                  const float f0 = epi_apply(__uint_as_float(v[j]), p.epi, p.param, b0, s0);
                  const float f1 = epi_apply(__uint_as_float(v[j + 1]), p.epi, p.param, b1, s1);
                  const __half2 hv = __floats2half2_rn(f0, f1);
                  w[e] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                o[q] = make_uint4(w[0], w[1], w[2], w[3]);
              }
              uint4* dp = reinterpret_cast<uint4*>(p.out + opix * 64 + h * 32);
#pragma unroll
              for (int q = 0; q < 4; ++q) dp[q] = o[q];
            }
          }
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace moe
