// Last upsample convolution of a branch FUSED with the dot-product half of the Conv3x3(64,1) head
// (models.py:29-33 + :130/:151).  conv3x3_pair_kernel writes PReLU(PixelShuffle(conv+b)) to HBM as 64 fp16
// channels per output pixel (128 B) only for head_tc_kernel to read it back and reduce it to nine numbers.
// Here the reduction happens while the pixel is still in shared memory:
//   * the epilogue writes its fp16 tile to the swizzled staging tile exactly as before — which is exactly the
//     layout of a K-major UMMA A operand — and, instead of a TMA store, a SECOND tcgen05.mma stream
//     (M = 256 over the pair, N = 16 = 9 taps padded, K = 64) multiplies it with the head filter:
//     P_t(Y,X) = <w_head[t], act(Y,X,:)>, fp32 in TMEM;
//   * four reader warps take P (9 floats per output pixel) out of TMEM and do the HORIZONTAL third of the head's 3x3
//     stencil on the spot: H_dy(Y,X) = (P[dy,0](Y,X-1) + P[dy,1](Y,X)) + P[dy,2](Y,X+1).  A lane owns the output pixels
//     2x and 2x+1 of one row, its neighbours' terms arrive by warp shuffle (across the four warps through a 96-byte
//     shared-memory exchange); only the terms that cross the 256-output-pixel strip of the CTA are unavailable — they
//     are exported to a small edge array (6 floats per strip and row) and added by the stencil kernel.  What leaves
//     the kernel is 3 floats per output pixel (12 B) instead of 9 (36 B, round 1) or the 128 B of the unfused path; the
//     2 x 12.8 GB intermediate tensors of an a4 tile are never written or read.  Each branch keeps its own array: the
//     reference's half model rounds each head to fp16 BEFORE adding them (models.py:38);
//   * head_stencil_kernel (below) finishes vertically: S = (H_0(Y-1) + H_1(Y)) + H_2(Y+1) per branch, rounds each, adds,
//     rounds, blends, stores: 2 x 12 + 2 B per output pixel.
//   * PixelShuffle(3) (a3): the three sub-pixels of an output row belong to different chunk groups (five groups of two chunks),
//     i.e. to different CTA pairs, so the readers export all nine dot products of every sub-pixel chunk unreduced
//     (pbuf[n][y][chunk * 9 + tap][W], 324 B per input pixel-plane instead of the 1 152 B of activations) and
//     head_stencil9_kernel does the whole 3 x 3 gather.
// Warp roles per CTA: 0 TMA producer, 1 MMA issuer (leader) / weight handshake (peer), 2..9 epilogue,
// 10 head-MMA issuer (leader), 11..14 P readers.  TMEM: 3 accumulator stages x 128 columns + 2 P stages x 32.
#pragma once
#include "conv_pair.cuh"
#include "kernels_simt.cuh"

namespace moe {

struct PairHeadParams {
  ConvParams c;            // the convolution (r = 2 or 3, EPI_BIAS_PRELU); c.out is unused
  const uint8_t* head_img; // [16 rows][128 B] swizzled fp16: rows 0..8 = the 9 taps of THIS branch's head filter
  float* hbuf;             // r = 2: [N][3][2H][2W] fp32, this branch's H_dy planes; r = 3: [N][H][9 chunks x 9 taps][W] fp32 dot products
  float* ebuf;             // [N][2H][2 * strips][6] fp32: per CTA strip and output row, {first pixel's P[dy,2], last pixel's P[dy,0]}
};

constexpr int kPairHeadThreads = 15 * 32;

struct PairHeadCfg {
  static constexpr int kSlots = 6;
  static constexpr int kAccStages = 3;
  static constexpr int kPStages = 2;
  static constexpr uint32_t kPCol0 = kAccStages * 128;     // 384
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kSmemBytes = 1024 + kSlots * kSlotBytes + kChunkImgBytes + 2 * kStageBytes + 1024 + 1024 + 256;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairHeadThreads, 1)
conv3x3_pair_head_kernel(const __grid_constant__ ConvMaps maps, const PairHeadParams hp, const __grid_constant__ ConvBias cb)
{
  using Cfg = PairHeadCfg;
  constexpr int S = Cfg::kSlots, AS = Cfg::kAccStages, PS = Cfg::kPStages;
  const ConvParams& p = hp.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ring = base;
  const uint32_t wsm = ring + S * kSlotBytes;
  const uint32_t stg = wsm + kChunkImgBytes;                       // two staging tiles = the two chunks of a row
  const uint32_t hsm = stg + 2 * kStageBytes;                       // this CTA's 8 rows of the head filter (1 KB)
  const uint32_t bars = hsm + 1024;
  const uint32_t full = bars, empty = full + 8 * S, tfull = empty + 8 * S, tempty = tfull + 8 * AS;
  const uint32_t staged = tempty + 8 * AS, pfull = staged + 8 * PS, pempty = pfull + 8 * PS;
  const uint32_t wbar = pempty + 8 * PS, wpeer = wbar + 8, dbar = wpeer + 8, tslot = dbar + 8;
  const uint32_t sq_items = bars + 256, sq_bars = bars + 320;       // item queue (kSchedQ ints + kSchedQ barriers), ends at +448
  float* xch_ptr = reinterpret_cast<float*>(smem + (bars + 1024 - base));   // P readers' exchange: [2 rows in flight][4 warps][6] floats
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  uint8_t* stg_ptr = smem + (stg - base);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int groups = pair_groups(p.r), nchunks = p.r * p.r;   // 2 chunk groups for PixelShuffle(2), 5 for PixelShuffle(3)
  const int g_fixed = pair % groups;              // npairs is a multiple of the group count (host): a pair keeps its chunk group
  const int my_chunk = min(g_fixed * 2 + static_cast<int>(rank), nchunks - 1);   // the idle half of PixelShuffle(3)'s last group recomputes chunk 8
  const PairSched sc = sched_make(p, sq_items, sq_bars, pair, npairs, groups);
  const uint64_t t_start = p.dbg ? ptx::globaltimer_ns() : 0;
  int n_items = 0;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(full + 8 * i, 1); ptx::mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) { ptx::mbar_init(tfull + 8 * i, 1); ptx::mbar_init(tempty + 8 * i, 2 * kEpiWarps); }
    for (int i = 0; i < PS; ++i) { ptx::mbar_init(staged + 8 * i, 2); ptx::mbar_init(pfull + 8 * i, 1); ptx::mbar_init(pempty + 8 * i, 2 * 4); }
    ptx::mbar_init(wbar, 1);
    ptx::mbar_init(wpeer, 1);
    ptx::mbar_init(dbar, 1);
    sched_init_bars(sq_bars);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&maps.in);
  }
  if (warp == 1) ptx::tmem_alloc_pair(tslot, Cfg::kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tslot_ptr;
  ptx::grid_dep_launch();
  if (warp == 0) {                                 // weights before the dependency wait (see conv3x3_pair_kernel)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, kChunkImgBytes + 1024);
      const uint8_t* src = p.w_img + static_cast<size_t>(my_chunk) * kChunkImgBytes;
      for (int tap = 0; tap < 9; ++tap) ptx::bulk_load_1d(wsm + tap * 8192, src + tap * 8192, 8192, wbar);
      ptx::bulk_load_1d(hsm, hp.head_img + rank * 1024, 1024, wbar);       // taps 8*rank .. 8*rank+7
    }
    __syncwarp();
  }
  ptx::grid_dep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    uint32_t ld = 0, ord = 0;
    int j = pair / sc.groups;
    int item = sched_produce(sc, leader_cta, lane, j, ord++);
    while (item >= 0) {
      const int next_item = sched_produce(sc, leader_cta, lane, j, ord++);     // drawn one item ahead of the loads
      int g, n, sp, y0, y1;
      pair_decode_item(p, item, g, n, sp, y0, y1);
      const int x0 = (sp * 2 + static_cast<int>(rank)) * kStripW;
      ++n_items;
      item = next_item;
      for (int yy = y0 - 1; yy <= y1; ++yy, ++ld) {
        const uint32_t slot = ld % S;
        ptx::mbar_wait(empty + 8 * slot, ((ld / S) & 1) ^ 1);
        if (ptx::elect_one()) {
          if (leader_cta) ptx::mbar_expect_tx(full + 8 * slot, 2 * kRowBytes);
          ptx::tma_load_4d_pair(ring + slot * kSlotBytes, &maps.in, ptx::mapa(full + 8 * slot, 0), 0, x0 - 1, yy, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (!leader_cta) {
      ptx::mbar_wait(wbar, 0);
      if (ptx::elect_one()) ptx::mbar_arrive_cluster(ptx::mapa(wpeer, 0));
      __syncwarp();
    } else {
      // ---------------------------------------------------------- leader: main MMA stream (as conv3x3_pair_kernel, 3 stages)
      constexpr uint32_t idesc = ptx::idesc_f16_f32(256, 128);
      const uint64_t bdesc0 = ptx::smem_desc_sw128(wsm, 1024, 0);
      const uint64_t adesc0 = ptx::smem_desc_sw128(ring, 1024, 0);
      ptx::mbar_wait(wbar, 0);
      ptx::mbar_wait(wpeer, 0);
      ptx::tc_fence_after_sync();
      uint32_t cons = 0, acc = 0;
      uint32_t ord = 0;
      for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
        int g, n, sp, y0, y1;
        pair_decode_item(p, item, g, n, sp, y0, y1);
        const int nrows = y1 - y0;
        ptx::mbar_wait(full + 8 * (cons % S), (cons / S) & 1);
        ptx::mbar_wait(full + 8 * ((cons + 1) % S), ((cons + 1) / S) & 1);
        for (int j = 0; j < nrows; ++j) {
          const uint32_t l2 = cons + j + 2;
          ptx::mbar_wait(full + 8 * (l2 % S), (l2 / S) & 1);
          const uint32_t stage = acc % AS;
          ptx::mbar_wait(tempty + 8 * stage, ((acc / AS) & 1) ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + stage * 128;
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
                  ptx::mma_f16_ss_pair(d_tmem, arow + (dx * 8 + k * 2), bdesc0 + ((dy * 3 + dx) * 512 + k * 2), idesc, p.center_only ? (k != 0) : ((dy | dx | k) != 0));
                }
              }
            }
            ptx::mma_commit_pair_mc(tfull + 8 * stage, 3);
            ptx::mma_commit_pair_mc(empty + 8 * ((cons + j) % S), 3);
          }
          __syncwarp();
          ++acc;
        }
        if (ptx::elect_one()) {
          ptx::mma_commit_pair_mc(empty + 8 * ((cons + nrows) % S), 3);
          ptx::mma_commit_pair_mc(empty + 8 * ((cons + nrows + 1) % S), 3);
        }
        __syncwarp();
        cons += nrows + 2;
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------ epilogue: accumulator -> fp16 staging tile (= head A operand)
    const int lgrp = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int L = lgrp * 32 + lane;
    const bool lead_warp = (warp == 2);
    uint8_t* my_row = stg_ptr + ch * kStageBytes + L * 128;
    const int sw = L & 7;
    const uint32_t tempty_leader = ptx::mapa(tempty, 0);
    const uint32_t staged_leader = ptx::mapa(staged, 0);
    const int my_bias = min(g_fixed * 2 + ch, nchunks - 1) * 64;        // index into cb.v (constant bank)
    uint32_t acc = 0;
    uint32_t ord = 0;
    for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
      int g, n, sp, y0, y1;
      pair_decode_item(p, item, g, n, sp, y0, y1);
      for (int y = y0; y < y1; ++y, ++acc) {
        const uint32_t stage = acc % AS;
        ptx::mbar_wait(tfull + 8 * stage, (acc / AS) & 1);
        ptx::tc_fence_after_sync();
        uint4 pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + stage * 128 + ch * 64 + h * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = q * 8 + e * 2;
              const float f0 = epi_apply<EPI_BIAS_PRELU>(__uint_as_float(v[j]), p.param, cb.v[my_bias + h * 32 + j], 0.f, p.bias_fused);
              const float f1 = epi_apply<EPI_BIAS_PRELU>(__uint_as_float(v[j + 1]), p.param, cb.v[my_bias + h * 32 + j + 1], 0.f, p.bias_fused);
              const __half2 hv = __floats2half2_rn(f0, f1);
              w[e] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            pk[h * 4 + q] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + 8 * stage);
        // the staging tiles are the A operand of the previous row's head MMAs: wait until those completed
        if (acc > 0) ptx::mbar_wait(pfull + 8 * ((acc - 1) % PS), ((acc - 1) / PS) & 1);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) = pk[c];
        ptx::fence_proxy_async_smem();                  // generic-proxy writes -> visible to the tensor core's async proxy
        ptx::named_bar_sync(2, kEpiThreads);
        if (lead_warp) {
          if (ptx::elect_one()) ptx::mbar_arrive_cluster_release(staged_leader + 8 * (acc % PS));
          __syncwarp();
        }
      }
    }
  } else if (warp == 10) {
    if (leader_cta) {
      // ---------------------------------------------------------- leader: head MMA stream, P = staging tile x head filter
      constexpr uint32_t idesc = ptx::idesc_f16_f32(256, 16);
      const uint64_t a0 = ptx::smem_desc_sw128(stg, 1024, 0);
      const uint64_t b0 = ptx::smem_desc_sw128(hsm, 1024, 0);
      ptx::mbar_wait(wbar, 0);
      ptx::mbar_wait(wpeer, 0);
      uint32_t acc = 0;
      uint32_t ord = 0;
      for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
        int g, n, sp, y0, y1;
        pair_decode_item(p, item, g, n, sp, y0, y1);
        for (int y = y0; y < y1; ++y, ++acc) {
          const uint32_t ps = acc % PS;
          ptx::mbar_wait_cluster(staged + 8 * ps, (acc / PS) & 1);      // both CTAs' staging tiles are written
          ptx::mbar_wait(pempty + 8 * ps, ((acc / PS) & 1) ^ 1);        // both CTAs' readers drained this P stage
          ptx::tc_fence_after_sync();
          if (ptx::elect_one()) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int k = 0; k < 4; ++k) if (k < p.ksteps)
                ptx::mma_f16_ss_pair(tmem_base + Cfg::kPCol0 + ps * 32 + c * 16, a0 + (c * (kStageBytes >> 4) + k * 2), b0 + k * 2, idesc, k != 0);
            ptx::mma_commit_pair_mc(pfull + 8 * ps, 3);
          }
          __syncwarp();
        }
      }
      if (ptx::elect_one()) ptx::mma_commit_pair(dbar);     // drain (both MMA streams complete in issue order per thread; see below)
      __syncwarp();
      ptx::mbar_wait_drain(dbar, 0);
    }
  } else {
    // ------------------------------------------------------------ P readers (warps 11..14): TMEM -> horizontal stencil third -> HBM
    const int lgrp = warp & 3;                       // TMEM lane quadrant of this warp = pixels 32*lgrp .. +31 of the strip
    const int L = lgrp * 32 + lane;
    const uint32_t pempty_leader = ptx::mapa(pempty, 0);
    const int Ho = p.H * 2, Wo = p.W * 2;
    const size_t plane = static_cast<size_t>(Ho) * Wo;
    const int estrips = 2 * p.strips;                // CTA strips per row (p.strips counts strip PAIRS)
    uint32_t acc = 0;
    uint32_t ord = 0;
    for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
      int g, n, sp, y0, y1;
      pair_decode_item(p, item, g, n, sp, y0, y1);
      const int strip = sp * 2 + static_cast<int>(rank);
      const int x = strip * kStripW + L;
      const bool valid = x < p.W;                     // pixels right of the tile are zero padding for the head convolution
      for (int y = y0; y < y1; ++y, ++acc) {
        const uint32_t ps = acc % PS;
        ptx::mbar_wait(pfull + 8 * ps, (acc / PS) & 1);
        ptx::tc_fence_after_sync();
        uint32_t v0[16], v1[16];
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + Cfg::kPCol0 + ps * 32, v0);
        ptx::tmem_ld16(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + Cfg::kPCol0 + ps * 32 + 16, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(pempty_leader + 8 * ps);
        if (p.r == 3) {
          // PixelShuffle(3): the three sub-pixels of an output row sit in different chunk groups (= different CTA pairs), so
          // no horizontal sum here.  The nine dot products of each sub-pixel chunk leave at INPUT resolution as
          // pbuf[n][y][chunk * 9 + tap][W]: full 128-byte lines per warp, and the 81 row segments of an input row adjacent in
          // memory (as 81 separate planes head_stencil9_kernel read 243 DRAM pages at once and ran at a quarter of the HBM rate).
          if (valid) {
            float* dst = hp.hbuf + ((static_cast<size_t>(n) * p.H + y) * 81 + g_fixed * 18) * p.W + x;
#pragma unroll
            for (int t = 0; t < 9; ++t) dst[t * p.W] = __uint_as_float(v0[t]);
            if (g_fixed * 2 + 1 < nchunks) {
#pragma unroll
              for (int t = 0; t < 9; ++t) dst[(9 + t) * p.W] = __uint_as_float(v1[t]);
            }
          }
          continue;
        }
        // chunk c of this pair's group g is sub-pixel (i, j) = (g, c): v0 = output pixel (2y + g, 2x), v1 = (2y + g, 2x + 1)
        float c0[9], c1[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) { c0[t] = valid ? __uint_as_float(v0[t]) : 0.f; c1[t] = valid ? __uint_as_float(v1[t]) : 0.f; }
        float left[3], right[3];                      // P[dy,0] of output pixel 2x - 1, P[dy,2] of output pixel 2x + 2
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          left[dy] = __shfl_up_sync(0xffffffffu, c1[dy * 3 + 0], 1);
          right[dy] = __shfl_down_sync(0xffffffffu, c0[dy * 3 + 2], 1);
        }
        float* xr = xch_ptr + (acc & 1) * 24;
        if (lane == 31) { xr[lgrp * 6 + 0] = c1[0]; xr[lgrp * 6 + 1] = c1[3]; xr[lgrp * 6 + 2] = c1[6]; }
        if (lane == 0) { xr[lgrp * 6 + 3] = c0[2]; xr[lgrp * 6 + 4] = c0[5]; xr[lgrp * 6 + 5] = c0[8]; }
        ptx::named_bar_sync(3, 128);                  // the four reader warps; two exchange buffers, so one barrier per row suffices
        float* erow = hp.ebuf + ((static_cast<size_t>(n) * Ho + (2 * y + g_fixed)) * estrips + strip) * 6;
        if (lane == 0) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) left[dy] = lgrp > 0 ? xr[(lgrp - 1) * 6 + dy] : 0.f;
          if (lgrp == 0) { erow[0] = c0[2]; erow[1] = c0[5]; erow[2] = c0[8]; }      // the strip to the left needs these
        }
        if (lane == 31) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) right[dy] = lgrp < 3 ? xr[(lgrp + 1) * 6 + 3 + dy] : 0.f;
          if (lgrp == 3) { erow[3] = c1[0]; erow[4] = c1[3]; erow[5] = c1[6]; }      // the strip to the right needs these
        }
        if (valid) {
          float* dst = hp.hbuf + static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(2 * y + g_fixed) * Wo + 2 * x;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
            *reinterpret_cast<float2*>(dst + dy * plane) =
                make_float2((left[dy] + c0[dy * 3 + 1]) + c1[dy * 3 + 2], (c0[dy * 3 + 0] + c1[dy * 3 + 1]) + right[dy]);
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();
  if (leader_cta && tid == 0) { sched_finish(p, npairs); pair_debug(p, pair, t_start, n_items); }
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// out(Y,X) = round16( round16(S_u) + round16(S_r) ), S_b = (H^b_0(Y-1,X) + H^b_1(Y,X)) + H^b_2(Y+1,X), zero outside the computed
// rectangle; at the two pixels of a row that touch the border of a 256-pixel CTA strip the horizontal term the convolution
// kernel could not reach comes from the edge array.  Then the seam blend and the canvas store of head_blend_kernel /
// head_tc_kernel.  A thread owns 4 consecutive pixels of a row: six aligned 16-byte loads (two 8-byte loads each when the
// row pitch is not a multiple of 4 floats), every H value is read exactly once.
struct HeadStencilParams {
  HeadParams g;            // geometry, seam and canvas (u/r/wu/wr unused)
  const float* hu;         // [N][3][H][W]: H_dy planes of branch `u`
  const float* hr;         // same for branch `convt_R1`
  const float* eu;         // [N][H][estrips][6] edge terms of branch `u`
  const float* er;
  int estrips;
};

constexpr int kStencilThreads = 128;
constexpr int kStencilPx = 4;
constexpr int kStencilStripOut = 2 * kStripW;     // output pixels per CTA strip of the convolution kernel

__global__ void __launch_bounds__(kStencilThreads) head_stencil_kernel(const HeadStencilParams p)
{
  const HeadParams& g = p.g;
  const int x = (blockIdx.x * kStencilThreads + threadIdx.x) * kStencilPx;
  const int n = blockIdx.z;
  if (x >= g.W) return;
  const size_t plane = static_cast<size_t>(g.H) * g.W;
  const float* hb[2] = {p.hu + static_cast<size_t>(n) * 3 * plane, p.hr + static_cast<size_t>(n) * 3 * plane};
  const float* eb[2] = {p.eu + static_cast<size_t>(n) * g.H * p.estrips * 6, p.er + static_cast<size_t>(n) * g.H * p.estrips * 6};
  const bool vec4 = (g.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(hb[0]) | reinterpret_cast<uintptr_t>(hb[1])) & 15) == 0 &&
                    (plane & 3) == 0;                            // rows of every plane 16-byte aligned
  const int nvalid = min(kStencilPx, g.W - x);
  const int strip = x / kStencilStripOut;
  const bool fix_left = x % kStencilStripOut == 0 && strip > 0;                              // pixel 0 of this thread
  const bool fix_right = x % kStencilStripOut == kStencilStripOut - kStencilPx && x + kStencilPx < g.W;   // pixel 3
  for (int y = blockIdx.y; y < g.H; y += gridDim.y) {            // gridDim.y is capped at 65535 rows
    const int cy = g.oy + y;
    if (cy < g.keep_y0 || cy >= g.keep_y1) continue;
    float head[2][kStencilPx];                                   // the two heads' stencil sums, each rounded to fp16
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      float h[3][kStencilPx];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
#pragma unroll
        for (int i = 0; i < kStencilPx; ++i) h[dy][i] = 0.f;
        if (yy < 0 || yy >= g.H) continue;
        const float* row = hb[b] + static_cast<size_t>(dy) * plane + static_cast<size_t>(yy) * g.W + x;
        if (nvalid == kStencilPx && vec4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(row));
          h[dy][0] = q.x; h[dy][1] = q.y; h[dy][2] = q.z; h[dy][3] = q.w;
        } else if (nvalid == kStencilPx && ((reinterpret_cast<uintptr_t>(row) & 7) == 0)) {
          const float2 q0 = __ldg(reinterpret_cast<const float2*>(row));
          const float2 q1 = __ldg(reinterpret_cast<const float2*>(row + 2));
          h[dy][0] = q0.x; h[dy][1] = q0.y; h[dy][2] = q1.x; h[dy][3] = q1.y;
        } else {
#pragma unroll
          for (int i = 0; i < kStencilPx; ++i) if (i < nvalid) h[dy][i] = __ldg(row + i);
        }
        const float* e = eb[b] + (static_cast<size_t>(yy) * p.estrips) * 6;
        if (fix_left) h[dy][0] += __ldg(e + (strip - 1) * 6 + 3 + dy);        // P[dy,0] of the last pixel of the strip to the left
        if (fix_right) h[dy][3] += __ldg(e + (strip + 1) * 6 + dy);           // P[dy,2] of the first pixel of the strip to the right
      }
#pragma unroll
      for (int i = 0; i < kStencilPx; ++i) head[b][i] = h_round((h[0][i] + h[1][i]) + h[2][i]);
    }
    __half* dst = g.canvas + n * g.plane_stride + static_cast<int64_t>(cy) * g.row_stride + (g.ox + x);
    float out[kStencilPx];
    bool keep[kStencilPx];
#pragma unroll
    for (int i = 0; i < kStencilPx; ++i) {
      const int cx = g.ox + x + i;
      keep[i] = i < nvalid && cx >= g.keep_x0 && cx < g.keep_x1;
      float v = h_round(head[0][i] + head[1][i]);               // u + convt_R1(t), models.py:38
      if (keep[i] && (cy < g.blend_y1 || cx < g.blend_x1)) {
        const float old = __half2float(dst[i]);
        if (cy < g.blend_y1) v = h_round(old + h_round(g.ramp[cy - g.ramp_y0] * h_round(v - old)));
        if (cx < g.blend_x1) v = h_round(old + h_round(g.ramp[cx - g.ramp_x0] * h_round(v - old)));
      }
      out[i] = v;
    }
    if (keep[0] && keep[1] && keep[2] && keep[3] && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
      const __half2 h0 = __floats2half2_rn(out[0], out[1]), h1 = __floats2half2_rn(out[2], out[3]);
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&h0);
      w.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(dst) = w;
    } else {
#pragma unroll
      for (int i = 0; i < kStencilPx; ++i) if (keep[i]) dst[i] = __float2half_rn(out[i]);
    }
  }
}

// PixelShuffle(3) (a3 = Net3x, models.py:135-143): the fused convolution leaves, per branch, the nine dot products
// P_t = <w_head[t], act> of every output pixel at input resolution, pbuf[n][y][chunk (i,j) * 9 + tap][w] with output pixel
// (3y + i, 3x + j).  out(Y,X) = round16(round16(S_u) + round16(S_r)), S = sum over taps (dy,dx) of P_(dy,dx)(Y+dy-1, X+dx-1) in the
// order of head_stencil_kernel: ((P_dy0 + P_dy1) + P_dy2) per dy, then (H_0 + H_1) + H_2; zero outside the computed rectangle.  A
// thread owns one input pixel = a 3 x 3 block of outputs: 81 loads per branch, every one a full line per warp; 72 B per output
// pixel instead of the 256 B the unfused head reads.  Then the seam blend and canvas store of the other head kernels.
struct HeadStencil9Params {
  HeadParams g;            // geometry at OUTPUT resolution (g.H = 3h, g.W = 3w), seam and canvas (u/r/wu/wr unused)
  const float* pu;         // [N][h][81][w]
  const float* pr;
  int h, w;
};

__global__ void __launch_bounds__(128) head_stencil9_kernel(const HeadStencil9Params p)
{
  const HeadParams& g = p.g;
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int n = blockIdx.z;
  if (x >= p.w) return;
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  const float* pb[2] = {p.pu + static_cast<size_t>(n) * 81 * plane, p.pr + static_cast<size_t>(n) * 81 * plane};
  for (int y = blockIdx.y; y < p.h; y += gridDim.y) {
    if (g.oy + 3 * y + 2 < g.keep_y0 || g.oy + 3 * y >= g.keep_y1) continue;
    float head[2][9];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          float hsum[3];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int a = i + dy - 1;                                  // output row 3y + a: input row y + oyy, sub-pixel row ii
            const int oyy = a < 0 ? -1 : (a > 2 ? 1 : 0), ii = (a + 3) % 3;
            const int yy = y + oyy;
            float term[3];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const int c = j + dx - 1;
              const int oxx = c < 0 ? -1 : (c > 2 ? 1 : 0), jj = (c + 3) % 3;
              const int xx = x + oxx;
              const bool in = yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
              term[dx] = in ? __ldg(pb[b] + (static_cast<size_t>(yy) * 81 + ((ii * 3 + jj) * 9 + dy * 3 + dx)) * p.w + xx) : 0.f;
            }
            hsum[dy] = (term[0] + term[1]) + term[2];
          }
          head[b][i * 3 + j] = h_round((hsum[0] + hsum[1]) + hsum[2]);
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int cy = g.oy + 3 * y + i;
      if (cy < g.keep_y0 || cy >= g.keep_y1) continue;
      __half* dst = g.canvas + n * g.plane_stride + static_cast<int64_t>(cy) * g.row_stride + (g.ox + 3 * x);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int cx = g.ox + 3 * x + j;
        if (cx < g.keep_x0 || cx >= g.keep_x1) continue;
        float v = h_round(head[0][i * 3 + j] + head[1][i * 3 + j]);   // u + convt_R1(t), models.py:38
        if (cy < g.blend_y1 || cx < g.blend_x1) {
          const float old = __half2float(dst[j]);
          if (cy < g.blend_y1) v = h_round(old + h_round(g.ramp[cy - g.ramp_y0] * h_round(v - old)));
          if (cx < g.blend_x1) v = h_round(old + h_round(g.ramp[cx - g.ramp_x0] * h_round(v - old)));
        }
        dst[j] = __float2half_rn(v);
      }
    }
  }
}

}  // namespace moe
