// conv3x3, 64 -> 64*r*r channels, NHWC fp16, as an implicit GEMM on the 5th-gen tensor cores.
//
// One persistent CTA per SM walks DOWN a 128-pixel-wide column strip of one image plane:
//   * M tile   = one output row segment of 128 pixels (TMEM lane = pixel, column = output channel)
//   * K        = 9 taps x 64 input channels = 36 tcgen05.mma of 128 x 64 x 16
//   * A operand= the input rows y-1, y, y+1, each TMA-loaded ONCE (130 px x 128 B, 128-byte swizzle)
//                into a ring of row slots; the nine taps are nine *views* of those three slots: the
//                smem descriptor's start address is moved by dx*128 B (next pixel) and k*32 B (next
//                16 channels) — no im2col copy, no re-read of a row for its three vertical uses.
//                (Measured on B200, profiles/r01_diag_first_contact.log: the 128B swizzle is a function
//                of the absolute smem address, so a start address that is not 1024-aligned needs NO
//                base_offset in the descriptor; setting base_offset = (addr>>7)&7 gives garbage.)
//   * B operand= the layer's weights for one 64-channel output chunk, resident in smem for the whole
//                kernel (pre-swizzled on the host, fetched with cp.async.bulk).
//   * zero padding at the tile border is TMA out-of-bounds fill (x = -1, x = W, y = -1, y = H).
//   * epilogue = 8 warps (2 per TMEM lane quadrant, 32 channels each): tcgen05.ld -> PReLU | x scale +
//                skip | + bias, PReLU -> fp16 -> a swizzled 128 px x 128 B staging tile in smem -> ONE TMA
//                tensor store per row.  The residual operand arrives the same way (TMA load into the
//                staging tile, updated in place).  PixelShuffle is the store's tensor map: chunk q = (i,j)
//                owns the view out[:, i::r, j::r, :] (weights.py permutes the output channels to match).
//                Right-edge clipping is the TMA store's bounds check.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..9 = epilogue.
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
  const __half* in;       // NHWC (N,H,W,64)   (the tcgen05 path reads it through the tensor map)
  __half* out;            // NHWC (N,H*r,W*r,64)
  const __half* skip;     // NHWC (N,H,W,64) when epi == EPI_SCALE_SKIP (may alias out)
  int N, H, W;
  int r;                  // 1, 2 or 3
  int epi;
  float param;            // PReLU slope or residual scale
  int strips, nseg, seg_rows, items;
  int center_only;        // 1: a 1x1 convolution packed as a centre-tap 3x3 filter (MoeNet_lite2): issue only tap (1,1)
  int ksteps;             // K = 16 * ksteps input channels are real: 4, or 3 for the 48-filter models (NetDN, MoeNet_lite2), whose
                          // channels 48..63 are exactly zero in every activation — their MMAs are skipped, same bits
  int bias_fused;         // EPI_BIAS_PRELU: 0 = q(q(conv) + bias) (the reference on the GPU), 1 = q(conv + bias) (the half model on the CPU)
  // pair kernels only (conv_pair.cuh, "item scheduler"):
  int dynamic;            // 1: pairs draw their next item from the global counters, 0: round-robin by pair index
  int* sched;             // device int[kSchedInts]: per-chunk-group item counters + the count of finished pairs; all 0 between launches
  unsigned long long* dbg;// optional [pairs][4]: start ns, end ns, SM id, items processed (moe_engine_debug_buffer)
};

// The bias vector of an upsample convolution as a KERNEL PARAMETER (constant bank): the CTA-pair epilogues need all 64 values of
// their chunk per pixel; read from shared memory that was 16 LDS.128 per thread and row — as many L1 data-pipe wavefronts as the
// staging stores of the result (106 M of 225 M in conv3x3_pair_head_kernel), on the pipe the tensor core fetches its operands
// through.  An LDC does not touch that pipe.  (It halved the kernel's LSU share, 31 -> 16 % of cycles, without changing its
// duration: profiles/r02_arsb_experiments.txt.)
struct ConvBias {
  float v[9 * 64];        // [r*r chunks][64]
};

struct ConvMaps {
  CUtensorMap in;         // (64, W, H, N), box (64,130,1,1)
  CUtensorMap skip;       // (64, W, H, N), box (64,128,1,1)   (EPI_SCALE_SKIP only)
  CUtensorMap out[9];     // chunk q = (i,j): the (64, W, H, N) view out[:, i::r, j::r, :], box (64,128,1,1)
};

constexpr int kStripW = 128;
constexpr int kRowPx = kStripW + 2;
constexpr uint32_t kRowBytes = kRowPx * 128;          // 16640, the input TMA box
constexpr uint32_t kSlotBytes = 17 * 1024;            // row slot, 1024-aligned for the swizzle pattern
constexpr uint32_t kChunkImgBytes = 9 * 64 * 128;     // 73728
constexpr uint32_t kStageBytes = kStripW * 128;       // 16384: one output row segment
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kConvThreads = 64 + kEpiThreads;        // 320

struct ConvCfg {
  static constexpr int kSlots = 6;
  static constexpr int kAccStages = 4;
  static constexpr int kOutStages = 2;
  static constexpr uint32_t kTmemCols = kAccStages * 64;
  static constexpr uint32_t kBarBytes = 1024;
  static constexpr uint32_t kSmemBytes = 1024 + kSlots * kSlotBytes + kChunkImgBytes + kOutStages * kStageBytes + kBarBytes;
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

__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }

// Epilogue arithmetic with the rounding points of the reference's half model: every aten op of
// models.py rounds its result to fp16 — the convolution (fp32 accumulate, bias inside), PReLU, the ScaleLayer
// multiply (models.py:73) and the residual add (models.py:60) are separate ops.  Returns the value BEFORE the final
// rounding of the store (the caller packs with __floats2half2_rn).  Products / sums of two fp16 values are exact
// in fp32, so rounding them to fp16 afterwards is the single rounding the reference's fp32-opmath kernels do.
// A convolution WITH a bias is two ops on the GPU: aten runs cudnn_convolution without the bias, then output.add_(bias)
// — q(q(conv) + bias), measured on B200 (profiles/r02_cudnn_rounding_probe.log).  bias_fused = 1 selects q(conv + bias),
// what the CPU execution of the half model does (oneDNN adds the bias inside the convolution; the committed `.ref16`
// goldens were produced that way): moe_engine_set_conv_path bit 5.
template <int EPI>
__device__ __forceinline__ float epi_apply(float v, float param, float bias, float skip, int bias_fused) {
  if (EPI == EPI_PRELU) { v = h_round(v); return v >= 0.f ? v : param * v; }
  if (EPI == EPI_SCALE_SKIP) return h_round(h_round(v) * param) + skip;
  if (EPI == EPI_BIAS_PRELU) {
    const float t = bias_fused ? v : h_round(v);       // a select, not a branch: the epilogue is unrolled 32 values deep
    v = h_round(t + bias);
    return v >= 0.f ? v : param * v;
  }
  return v;
}
// run-time epilogue selector for the cross-check kernel (kernels_simt.cuh); the tensor-core kernels are compiled per epilogue:
// with `epi` a run-time value the compiler branched per accumulator value, 525 branches in the trunk kernel, and the
// epilogue warps fell behind the MMA stream (trunk convolutions 26 -> 35 ms per 4K frame, profiles/r02_bench_epilogue_branches.txt)
__device__ __forceinline__ float epi_apply_rt(float v, int epi, float param, float bias, float skip, int bias_fused) {
  switch (epi) {
    case EPI_PRELU: return epi_apply<EPI_PRELU>(v, param, bias, skip, bias_fused);
    case EPI_SCALE_SKIP: return epi_apply<EPI_SCALE_SKIP>(v, param, bias, skip, bias_fused);
    case EPI_BIAS_PRELU: return epi_apply<EPI_BIAS_PRELU>(v, param, bias, skip, bias_fused);
    default: return v;
  }
}

template <int EPI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p)
{
  using Cfg = ConvCfg;
  constexpr int S = Cfg::kSlots, AS = Cfg::kAccStages, OS = Cfg::kOutStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ring = base;
  const uint32_t wsm = ring + S * kSlotBytes;
  const uint32_t stg = wsm + kChunkImgBytes;                       // OS staging tiles, 1024-aligned
  const uint32_t bars = stg + OS * kStageBytes;
  const uint32_t full = bars, empty = full + 8 * S, tfull = empty + 8 * S, tempty = tfull + 8 * AS;
  const uint32_t skfull = tempty + 8 * AS, wbar = skfull + 8 * OS, tslot = wbar + 8;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  uint8_t* stg_ptr = smem + (stg - base);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncg = p.r * p.r;                     // one chunk per CTA; gridDim.x % ncg == 0 (host guarantees)
  const int chunk = blockIdx.x % ncg;
  const CUtensorMap* omap = &maps.out[chunk];

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(full + 8 * i, 1); ptx::mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) { ptx::mbar_init(tfull + 8 * i, 1); ptx::mbar_init(tempty + 8 * i, kEpiThreads); }
    for (int i = 0; i < OS; ++i) ptx::mbar_init(skfull + 8 * i, 1);
    ptx::mbar_init(wbar, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&maps.in);
    ptx::prefetch_tmap(omap);
  }
  if (warp == 1) ptx::tmem_alloc(tslot, Cfg::kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tslot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp converged, one lane issues)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, kChunkImgBytes);
      const uint8_t* src = p.w_img + static_cast<size_t>(chunk) * kChunkImgBytes;
      for (int tap = 0; tap < 9; ++tap) ptx::bulk_load_1d(wsm + tap * 8192, src + tap * 8192, 8192, wbar);
    }
    __syncwarp();
    uint32_t ld = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, x0, y0, y1;
      conv_decode_item(p, item, ncg, n, x0, y0, y1);
      for (int yy = y0 - 1; yy <= y1; ++yy, ++ld) {
        const uint32_t slot = ld % S;
        ptx::mbar_wait(empty + 8 * slot, ((ld / S) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(full + 8 * slot, kRowBytes);
          ptx::tma_load_4d(ring + slot * kSlotBytes, &maps.in, full + 8 * slot, 0, x0 - 1, yy, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp converged, one lane issues)
    constexpr uint32_t idesc = ptx::idesc_f16_f32(128, 64);
    const uint64_t bdesc0 = ptx::smem_desc_sw128(wsm, 1024, 0);
    const uint64_t adesc0 = ptx::smem_desc_sw128(ring, 1024, 0);
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
        const uint32_t d_tmem = tmem_base + stage * 64;
        // the descriptor's address field counts 16-byte units: slot = 1088, pixel = 8, 16 channels = 2, tap = 512
        const uint64_t arow0 = adesc0 + static_cast<uint64_t>(((cons + j) % S) * (kSlotBytes >> 4));
        const uint64_t arow1 = adesc0 + static_cast<uint64_t>(((cons + j + 1) % S) * (kSlotBytes >> 4));
        const uint64_t arow2 = adesc0 + static_cast<uint64_t>(((cons + j + 2) % S) * (kSlotBytes >> 4));
        if (ptx::elect_one()) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t arow = dy == 0 ? arow0 : (dy == 1 ? arow1 : arow2);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              if (p.center_only && (dy != 1 || dx != 1)) continue;
#pragma unroll
              for (int k = 0; k < 4; ++k) if (k < p.ksteps) {
                ptx::mma_f16_ss(d_tmem, arow + (dx * 8 + k * 2), bdesc0 + ((dy * 3 + dx) * 512 + k * 2), idesc, p.center_only ? (k != 0) : ((dy | dx | k) != 0));
              }
            }
          }
          ptx::mma_commit(tfull + 8 * stage);               // accumulator ready for the epilogue
          ptx::mma_commit(empty + 8 * ((cons + j) % S));    // input row y-1 is no longer needed
        }
        __syncwarp();
        ++acc;
      }
      if (ptx::elect_one()) {
        ptx::mma_commit(empty + 8 * ((cons + nrows) % S));
        ptx::mma_commit(empty + 8 * ((cons + nrows + 1) % S));
      }
      __syncwarp();
      cons += nrows + 2;
    }
    // drain: commits complete in order, so once this one lands no arrive is still in flight
    if (ptx::elect_one()) ptx::mma_commit(wbar);
    __syncwarp();
    ptx::mbar_wait_drain(wbar, 1);
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int lgrp = warp & 3;                       // TMEM lanes this warp may read: 32*lgrp .. +31
    const int half = (warp - 2) >> 2;                // which 32 of the 64 channels
    const int L = lgrp * 32 + lane;                  // pixel within the strip == staging row
    const bool lead_warp = (warp == 2);               // its elected lane issues the TMA stores / residual loads
    constexpr bool has_skip = EPI == EPI_SCALE_SKIP;
    uint8_t* my_row = stg_ptr + L * 128;
    const int sw = L & 7;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = EPI == EPI_BIAS_PRELU ? __ldg(p.bias + chunk * 64 + half * 32 + j) : 0.f;

    uint32_t acc = 0;                                // output-row counter over ALL items
    if (has_skip && lead_warp && blockIdx.x < p.items) { // residual tile of the very first row
      int n, x0, y0, y1;
      conv_decode_item(p, blockIdx.x, ncg, n, x0, y0, y1);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(skfull, kStageBytes);
        ptx::tma_load_4d(stg, &maps.skip, skfull, 0, x0, y0, n);
      }
      __syncwarp();
    }
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      int n, x0, y0, y1;
      conv_decode_item(p, item, ncg, n, x0, y0, y1);
      for (int y = y0; y < y1; ++y, ++acc) {
        const uint32_t stage = acc % AS;
        const uint32_t os = acc % OS;
        uint8_t* row = my_row + os * kStageBytes;
        // (1) staging tile `os`: free (its previous TMA store has read it) or holding the residual
        if (has_skip) {
          ptx::mbar_wait(skfull + 8 * os, (acc / OS) & 1);
        } else {
          if (lead_warp) {
            if (ptx::elect_one()) ptx::bulk_wait_read<OS - 1>();
            __syncwarp();
          }
          ptx::named_bar_sync(1, kEpiThreads);
        }
        // (2) accumulator -> registers, TMEM stage back to the MMA warp
        ptx::mbar_wait(tfull + 8 * stage, (acc / AS) & 1);
        ptx::tc_fence_after_sync();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + stage * 64 + half * 32, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(tempty + 8 * stage);
        // (3) epilogue math, fp16 pack, swizzled staging write (16-byte chunk c of row L sits at c ^ (L & 7))
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cidx = ((half * 4 + q) ^ sw) << 4;
          uint4 sk = make_uint4(0, 0, 0, 0);
          if (has_skip) sk = *reinterpret_cast<const uint4*>(row + cidx);
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q * 8 + e * 2;
            float s0 = 0.f, s1 = 0.f;
            if (has_skip) {
              const uint32_t sw32 = reinterpret_cast<const uint32_t*>(&sk)[e];
              const __half2 hs = *reinterpret_cast<const __half2*>(&sw32);
              s0 = __low2float(hs);
              s1 = __high2float(hs);
            }
            const float f0 = epi_apply<EPI>(__uint_as_float(v[j]), p.param, bias_r[j], s0, p.bias_fused);
            const float f1 = epi_apply<EPI>(__uint_as_float(v[j + 1]), p.param, bias_r[j + 1], s1, p.bias_fused);
            const __half2 hv = __floats2half2_rn(f0, f1);
            w[e] = *reinterpret_cast<const uint32_t*>(&hv);
          }
          *reinterpret_cast<uint4*>(row + cidx) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // (4) hand the tile to the async proxy and store it
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(2, kEpiThreads);
        if (lead_warp) {
          // residual tile of the NEXT row goes into the other staging tile once that tile's store has read it
          int nn = n, nx0 = x0, ny = y + 1;
          bool more = has_skip;
          if (has_skip && ny >= y1) {
            const int nitem = item + gridDim.x;
            more = nitem < p.items;
            if (more) { int t1; conv_decode_item(p, nitem, ncg, nn, nx0, ny, t1); }
          }
          if (ptx::elect_one()) {
            ptx::tma_store_4d(omap, stg + os * kStageBytes, 0, x0, y, n);
            ptx::bulk_commit();
            if (more) {
              const uint32_t nos = (acc + 1) % OS;
              ptx::bulk_wait_read<OS - 1>();            // every store but the one just issued has read its tile
              ptx::mbar_expect_tx(skfull + 8 * nos, kStageBytes);
              ptx::tma_load_4d(stg + nos * kStageBytes, &maps.skip, skfull + 8 * nos, 0, nx0, ny, nn);
            }
          }
          __syncwarp();
        }
      }
    }
    if (lead_warp) {
      if (ptx::elect_one()) ptx::bulk_wait<0>();          // all stores complete before the CTA exits
      __syncwarp();
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
