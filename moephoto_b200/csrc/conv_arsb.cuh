// One residual block of the trunk, t' = t + scale * conv_2(PReLU(conv_1(t)))  (ARSB, models.py:76-80), as ONE kernel.
//
// Run as two launches of conv3x3_pair_trunk_kernel the block moves 640 B per pixel-plane through HBM (read t, write mid,
// read mid, read t again as the residual, write t') for 147 456 FLOP: it sits on the HBM side of the ridge and ran at
// 0.66 of the tensor peak with DRAM 84 % busy (profiles/r01_final_kernels_ncu.txt).  Here `mid` never leaves the SM:
//   * a CTA pair (cta_group::2) walks down a pair of 126-pixel-wide column strips.  Per step ONE new row of t arrives by
//     TMA (130 px: the strip + 2 px of halo each side) into a 5-slot ring;
//   * conv_1: M = 256 (128 px of each CTA) x N = 64 x K = 9 x 64 on the three newest t rows -> TMEM; four epilogue
//     warps round, apply PReLU, round (the reference's two ops), zero what lies outside the tile (conv_2's zero
//     padding) and write the row as fp16 into a 3-slot ring of `mid` rows in exactly the layout of a K-major
//     SWIZZLE_128B A operand (the trick of conv_pair_head.cuh);
//   * conv_2: the same MMA shape on the three newest mid rows (shifted descriptor views again; mid position i is pixel
//     x0 - 1 + i, so output pixel j reads mid rows j, j+1, j+2 and the last two of the 128 outputs are scrap) -> TMEM;
//     four more epilogue warps apply q(q(q(acc) * scale) + t): the residual row is still in the t ring, read from
//     shared memory instead of HBM, and the 126-pixel result row leaves through a staging tile and one TMA store.
// HBM traffic per pixel-plane: 128 B read (x 130/126 and + 4 halo rows per row segment) + 128 B written.
// What bounds it is the SM's SHARED-MEMORY DATA PIPE (one 128-byte wavefront per cycle), not HBM, not MMA issue and not the B
// exchange between the two SMs (profiles/r02_arsb_experiments.txt): an M = 256 x N = 64 x K = 16 MMA fetches per SM 4 KB of A +
// 1 KB of its B half and serves 1 KB to its partner = 48 wavefronts for 32 cycles of math; every N = 64 kernel of the engine runs
// at 49-54 cycles per MMA.  Moving the output epilogue to global loads and stores, issuing the two convolutions from two warps
// (freely, or taking turns), or giving every SM all the weights (conv_arsb_solo.cuh) changed nothing or made it slower; taking
// conv_2's A operand out of shared memory (kMidTmem, the default) is worth 2-3 % inside the power-capped frame.
// MMA issue order (one thread of the leader CTA): conv_1 runs TWO mid rows ahead of conv_2 — C1(y+2), C2(y), C1(y+3),
// C2(y+1), ... — so the epilogue that turns an accumulator into a mid row overlaps the other convolution's MMAs.
// Shared memory (kMidTmem = false): 5 x 17 KB t ring + 3 x 16 KB mid ring + 2 x 36 KB weights (this CTA's 32 output channels of both
// convolutions) + 16 KB staging = 223 KB; TMEM: 4 + 4 accumulator stages of 64 columns.  The default form (kMidTmem = true, below)
// keeps the mid rows in tensor memory instead.
// Warp roles per CTA: 0 TMA producer, 1 MMA issuer (leader) / weight handshake (peer), 2..5 mid epilogue, 6..9 output epilogue.
#pragma once
#include "conv_pair.cuh"

namespace moe {

struct ArsbParams {
  ConvParams c;            // conv_1: w_img, in (= t), out (= t'), N/H/W, param = PReLU slope, scheduler fields; strips = strip PAIRS of 126 px
  const uint8_t* w2_img;   // conv_2 weights, same image format
  float scale;             // ScaleLayer (models.py:66-73)
};

struct ArsbMaps {
  CUtensorMap in;          // (64, W, H, N), box (64,130,1,1)
  CUtensorMap out;         // (64, W, H, N), box (64,126,1,1)
};

constexpr int kArsbStripW = 126;

// kMidTmem = false: `mid` rows live in a shared-memory ring and conv_2 is an .ss MMA like conv_1 (round 2, first form).
// kMidTmem = true : `mid` rows live in TENSOR memory and conv_2 is a .ts MMA — its A operand costs no shared-memory bandwidth.
//   TMEM lanes cannot be shifted, so the mid epilogue writes every row three times (the dx = 0, 1, 2 views: lane j holds mid
//   position j + dx; the neighbours' pixels arrive by warp shuffle, across warps through 2 KB of shared memory) with tcgen05.st:
//   4 rows x 3 views x 32 columns = 384 TMEM columns, which leaves one 64-column accumulator per convolution.  One stage each is
//   enough because the two convolutions alternate in the tensor pipe: an accumulator is drained while the OTHER convolution runs.
template <bool kMidTmem>
struct ArsbCfgT {
  static constexpr int kTSlots = kMidTmem ? 6 : 5;
  static constexpr int kMSlots = kMidTmem ? 4 : 3;
  static constexpr int kAcc = kMidTmem ? 1 : 4;                   // accumulator stages per convolution
  static constexpr uint32_t kMSlotBytes = 16384;                  // (smem form) 128 mid pixels; the scrap outputs read 2 rows past it
  static constexpr uint32_t kMidCols = 96;                        // (TMEM form) columns of one mid row: 3 views x 64 fp16
  static constexpr uint32_t kAcc1Col = kMidTmem ? 4 * kMidCols : 0;
  static constexpr uint32_t kAcc2Col = kMidTmem ? 4 * kMidCols + 64 : 4 * 64;
  static constexpr uint32_t kWBytes = 9 * 32 * 128;               // one convolution, this CTA's half of the output channels
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kMidBytes = kMidTmem ? 2048 : kMSlots * kMSlotBytes;    // TMEM form: the cross-warp exchange of the mid epilogue
  static constexpr uint32_t kSmemBytes = 1024 + kTSlots * kSlotBytes + kMidBytes + 2 * kWBytes + kStageBytes + 1024;
};
using ArsbCfg = ArsbCfgT<false>;

// rows of one item: outputs [y0, y1), mid rows [m_lo, m_hi] (the ones inside the tile), t rows [y0 - 2, y1 + 1]
struct ArsbItem { int n, sp, y0, y1, m_lo, m_hi; };
__device__ __forceinline__ ArsbItem arsb_decode(const ConvParams& p, int item) {
  ArsbItem it;
  pair_trunk_decode(p, item, it.n, it.sp, it.y0, it.y1);
  it.m_lo = max(it.y0 - 1, 0);
  it.m_hi = min(it.y1, p.H - 1);
  return it;
}

// KS: K steps of 16 input channels per tap, 4 or 3 (ConvParams::ksteps) — a template parameter because the one thread that issues
// every MMA of the CTA pair has no cycle to spare for a predicate (a run-time test made the kernel 1.6x slower)
template <bool kMidTmem, int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
arsb_pair_kernel(const __grid_constant__ ArsbMaps maps, const ArsbParams ap)
{
  using Cfg = ArsbCfgT<kMidTmem>;
  constexpr int TS = Cfg::kTSlots, MS = Cfg::kMSlots, AS = Cfg::kAcc;
  // the 48-filter models: N = 48 output channels per MMA as well as K = 48 (their channels 48..63 are zero in and must be zero out);
  // cta_group::2 takes half of B from each CTA: 24 of the tap's 64 weight rows, three whole swizzle groups
  constexpr int kN = KS == 3 ? 48 : 64, kHalf = kN / 2;
  const ConvParams& p = ap.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t tring = base;
  const uint32_t mring = tring + TS * kSlotBytes;
  const uint32_t w1sm = mring + Cfg::kMidBytes;
  const uint32_t w2sm = w1sm + Cfg::kWBytes;
  const uint32_t stg = w2sm + Cfg::kWBytes;
  const uint32_t bars = stg + kStageBytes;
  const uint32_t tfull_t = bars, tempty_t = tfull_t + 8 * TS;                       // t ring (full: leader; empty: each CTA)
  const uint32_t mfull = tempty_t + 8 * TS, mempty = mfull + 8 * MS;                // mid ring (full: leader; empty: each CTA)
  const uint32_t a1full = mempty + 8 * MS, a1empty = a1full + 8 * AS;               // conv_1 accumulators
  const uint32_t a2full = a1empty + 8 * AS, a2empty = a2full + 8 * AS;              // conv_2 accumulators
  const uint32_t wbar = a2empty + 8 * AS, wpeer = wbar + 8, dbar = wpeer + 8, tslot = dbar + 8;
  const uint32_t sq_items = bars + 512, sq_bars = bars + 576;
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  uint8_t* tring_ptr = smem + (tring - base);
  uint8_t* mring_ptr = smem + (mring - base);
  uint8_t* stg_ptr = smem + (stg - base);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const PairSched sc = sched_make(p, sq_items, sq_bars, pair, npairs, 1);
  const uint64_t t_start = p.dbg ? ptx::globaltimer_ns() : 0;
  int n_items = 0;

  if (tid == 0) {
    for (int i = 0; i < TS; ++i) { ptx::mbar_init(tfull_t + 8 * i, 1); ptx::mbar_init(tempty_t + 8 * i, 4); }
    for (int i = 0; i < MS; ++i) { ptx::mbar_init(mfull + 8 * i, kMidTmem ? 2 * 4 : 2); ptx::mbar_init(mempty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) {
      ptx::mbar_init(a1full + 8 * i, 1); ptx::mbar_init(a1empty + 8 * i, 2 * 4);
      ptx::mbar_init(a2full + 8 * i, 1); ptx::mbar_init(a2empty + 8 * i, 2 * 4);
    }
    ptx::mbar_init(wbar, 1);
    ptx::mbar_init(wpeer, 1);
    ptx::mbar_init(dbar, 1);
    sched_init_bars(sq_bars);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&maps.in);
    ptx::prefetch_tmap(&maps.out);
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
      // output channels kHalf*rank .. +kHalf-1 of every tap, both convolutions (rows of 128 B in 8-row swizzle groups of 1 KB)
      ptx::mbar_expect_tx(wbar, 2 * 9 * kHalf * 128);
      for (int tap = 0; tap < 9; ++tap) {
        ptx::bulk_load_1d(w1sm + tap * 4096, p.w_img + tap * 8192 + rank * (kHalf * 128), kHalf * 128, wbar);
        ptx::bulk_load_1d(w2sm + tap * 4096, ap.w2_img + tap * 8192 + rank * (kHalf * 128), kHalf * 128, wbar);
      }
    }
    __syncwarp();
  }
  ptx::grid_dep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the t rows of every item
    uint32_t ld = 0, ord = 0;
    int j = pair / sc.groups;
    int item = sched_produce(sc, leader_cta, lane, j, ord++);
    while (item >= 0) {
      const int next_item = sched_produce(sc, leader_cta, lane, j, ord++);     // drawn one item ahead of the loads
      const ArsbItem it = arsb_decode(p, item);
      const int x0 = (it.sp * 2 + static_cast<int>(rank)) * kArsbStripW;
      ++n_items;
      item = next_item;
      for (int yy = it.y0 - 2; yy <= it.y1 + 1; ++yy, ++ld) {
        const uint32_t slot = ld % TS;
        ptx::mbar_wait(tempty_t + 8 * slot, ((ld / TS) & 1) ^ 1);
        if (ptx::elect_one()) {
          if (leader_cta) ptx::mbar_expect_tx(tfull_t + 8 * slot, 2 * kRowBytes);
          ptx::tma_load_4d_pair(tring + slot * kSlotBytes, &maps.in, ptx::mapa(tfull_t + 8 * slot, 0), 0, x0 - 2, yy, it.n);
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
      // ---------------------------------------------------------- leader: both MMA streams, conv_1 two mid rows ahead of conv_2
      constexpr uint32_t idesc = ptx::idesc_f16_f32(256, kN);
      const uint64_t b1 = ptx::smem_desc_sw128(w1sm, 1024, 0);
      const uint64_t b2 = ptx::smem_desc_sw128(w2sm, 1024, 0);
      const uint64_t at0 = ptx::smem_desc_sw128(tring, 1024, 0);
      const uint64_t am0 = ptx::smem_desc_sw128(mring, 1024, 0);
      ptx::mbar_wait(wbar, 0);
      ptx::mbar_wait(wpeer, 0);
      ptx::tc_fence_after_sync();
      uint32_t tcnt = 0, mcnt = 0, n1 = 0, n2 = 0;       // t slots / mid slots consumed before this item; accumulators issued
      uint32_t twaited = 0;                               // t rows (global sequence) whose `full` barrier has been observed
      uint32_t ord = 0;
      for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
        const ArsbItem it = arsb_decode(p, item);
        int next_m = it.m_lo;
        // conv_1 for mid row m: t rows m-1, m, m+1 = ring entries tcnt + (m-1 - (y0-2)) ...
        auto conv1 = [&](int m) {
          const uint32_t e0 = tcnt + static_cast<uint32_t>(m - 1 - (it.y0 - 2));
          while (twaited < e0 + 3) { ptx::mbar_wait(tfull_t + 8 * (twaited % TS), (twaited / TS) & 1); ++twaited; }
          const uint32_t stage = n1 % AS;
          ptx::mbar_wait(a1empty + 8 * stage, ((n1 / AS) & 1) ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + Cfg::kAcc1Col + stage * 64;
          const uint64_t r0 = at0 + static_cast<uint64_t>((e0 % TS) * (kSlotBytes >> 4));
          const uint64_t r1 = at0 + static_cast<uint64_t>(((e0 + 1) % TS) * (kSlotBytes >> 4));
          const uint64_t r2 = at0 + static_cast<uint64_t>(((e0 + 2) % TS) * (kSlotBytes >> 4));
          if (ptx::elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint64_t arow = dy == 0 ? r0 : (dy == 1 ? r1 : r2);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < KS; ++k)
                  ptx::mma_f16_ss_pair(d_tmem, arow + (dx * 8 + k * 2), b1 + ((dy * 3 + dx) * 256 + k * 2), idesc, (dy | dx | k) != 0);
            }
            ptx::mma_commit_pair_mc(a1full + 8 * stage, 3);
          }
          __syncwarp();
          ++n1;
        };
        for (int y = it.y0; y < it.y1; ++y) {
          // conv_1 runs two mid rows ahead of conv_2 — except before the item's first output row: C1(y0+2) needs t row y0+3, the
          // SIXTH row of the item, and the 5-slot ring only turns over once the output epilogue of row y0 has released its slots
          const int want = min((y == it.y0 && TS < 6) ? y + 1 : y + 2, it.m_hi);
          while (next_m <= want) conv1(next_m++);
          if (y == it.y1 - 1) {
            // the output epilogue releases the item's last t slots after THIS conv_2: every load of the item must have landed
            // (rows below the image are loaded but read by no MMA)
            const uint32_t tend = tcnt + static_cast<uint32_t>(it.y1 - it.y0 + 4);
            while (twaited < tend) { ptx::mbar_wait(tfull_t + 8 * (twaited % TS), (twaited / TS) & 1); ++twaited; }
          }
          // conv_2 for output row y: mid rows y-1, y, y+1 that exist
          const int lo = max(y - 1, it.m_lo), hi = min(y + 1, it.m_hi);
          for (int m = lo; m <= hi; ++m) {
            const uint32_t e = mcnt + static_cast<uint32_t>(m - it.m_lo);
            if (kMidTmem) ptx::mbar_wait(mfull + 8 * (e % MS), (e / MS) & 1);  // TMEM rows: ordered by the tcgen05 fences
            else ptx::mbar_wait_cluster(mfull + 8 * (e % MS), (e / MS) & 1);   // smem rows written by both CTAs' mid epilogues
          }
          const uint32_t stage = n2 % AS;
          ptx::mbar_wait(a2empty + 8 * stage, ((n2 / AS) & 1) ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + Cfg::kAcc2Col + stage * 64;
          if (ptx::elect_one()) {
            bool first = true;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int m = y + dy - 1;
              if (m < it.m_lo || m > it.m_hi) continue;                       // zero padding above / below the tile
              const uint32_t e = mcnt + static_cast<uint32_t>(m - it.m_lo);
              const uint64_t arow = am0 + static_cast<uint64_t>((e % MS) * (Cfg::kMSlotBytes >> 4));
              const uint32_t trow = tmem_base + (e % MS) * Cfg::kMidCols;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                  if (kMidTmem) ptx::mma_f16_ts_pair(d_tmem, trow + dx * 32 + k * 8, b2 + ((dy * 3 + dx) * 256 + k * 2), idesc, first ? 0u : 1u);
                  else ptx::mma_f16_ss_pair(d_tmem, arow + (dx * 8 + k * 2), b2 + ((dy * 3 + dx) * 256 + k * 2), idesc, first ? 0u : 1u);
                  first = false;
                }
            }
            ptx::mma_commit_pair_mc(a2full + 8 * stage, 3);
            // mid row y-1 was read for the last time (the last output row also retires y and y+1)
            if (y - 1 >= it.m_lo) ptx::mma_commit_pair_mc(mempty + 8 * ((mcnt + static_cast<uint32_t>(y - 1 - it.m_lo)) % MS), 3);
            if (y == it.y1 - 1)
              for (int m = max(y, it.m_lo); m <= it.m_hi; ++m) ptx::mma_commit_pair_mc(mempty + 8 * ((mcnt + static_cast<uint32_t>(m - it.m_lo)) % MS), 3);
          }
          __syncwarp();
          ++n2;
        }
        tcnt += static_cast<uint32_t>(it.y1 - it.y0 + 4);
        mcnt += static_cast<uint32_t>(it.m_hi - it.m_lo + 1);
      }
      if (ptx::elect_one()) ptx::mma_commit_pair(dbar);
      __syncwarp();
      ptx::mbar_wait_drain(dbar, 0);
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ mid epilogue (warps 2..5): conv_1 accumulator -> fp16 mid row in smem
    const int lgrp = warp & 3;
    const int L = lgrp * 32 + lane;                  // mid position: pixel x0 - 1 + L
    const int sw = L & 7;
    const uint32_t a1empty_leader = ptx::mapa(a1empty, 0);
    const uint32_t mfull_leader = ptx::mapa(mfull, 0);
    uint32_t n1 = 0, mc = 0;
    uint32_t ord = 0;
    for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
      const ArsbItem it = arsb_decode(p, item);
      const int x = (it.sp * 2 + static_cast<int>(rank)) * kArsbStripW - 1 + L;
      const bool inside = x >= 0 && x < p.W;         // conv_2 sees zeros outside the tile
      for (int m = it.m_lo; m <= it.m_hi; ++m, ++n1, ++mc) {
        const uint32_t stage = n1 % AS;
        ptx::mbar_wait(a1full + 8 * stage, (n1 / AS) & 1);
        ptx::tc_fence_after_sync();
        uint4 pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + Cfg::kAcc1Col + stage * 64 + h * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = q * 8 + e * 2;
              const float f0 = epi_apply<EPI_PRELU>(__uint_as_float(v[jj]), p.param, 0.f, 0.f, 0);
              const float f1 = epi_apply<EPI_PRELU>(__uint_as_float(v[jj + 1]), p.param, 0.f, 0.f, 0);
              const __half2 hv = __floats2half2_rn(f0, f1);
              w[e] = inside && h * 32 + jj < kN ? *reinterpret_cast<const uint32_t*>(&hv) : 0u;   // accumulator columns >= kN are never written
            }
            pk[h * 4 + q] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(a1empty_leader + 8 * stage);   // accumulator stage back to the MMA warp
        const uint32_t ms = mc % MS;
        if (!kMidTmem) {
          ptx::mbar_wait(mempty + 8 * ms, ((mc / MS) & 1) ^ 1);                // conv_2 has finished with this slot's previous row
          uint8_t* row = mring_ptr + ms * Cfg::kMSlotBytes + L * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = pk[c];
          ptx::fence_proxy_async_smem();                                       // generic-proxy writes -> the tensor core's async proxy
          ptx::named_bar_sync(1, 128);
          if (warp == 2) {
            if (ptx::elect_one()) ptx::mbar_arrive_cluster_release(mfull_leader + 8 * ms);
            __syncwarp();
          }
        } else {
          // view dx of the row: THIS lane must hold mid position L + dx.  Positions L+1, L+2 sit in the next lanes; the last two
          // lanes of a warp take them from the first two lanes of the next warp through shared memory (positions >= 128 feed
          // only the two scrap outputs: zeros).  xbuf[row parity][warp][lane 0 / 1][32 words]
          uint32_t* xb = reinterpret_cast<uint32_t*>(mring_ptr) + (mc & 1) * 256;
          if (lane < 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(xb + (lgrp * 2 + lane) * 32 + c * 4) = pk[c];
          }
          ptx::named_bar_sync(1, 128);
          ptx::mbar_wait(mempty + 8 * ms, ((mc / MS) & 1) ^ 1);                // conv_2 has finished with this TMEM row's previous tenant
          ptx::tc_fence_after_sync();
          const uint32_t trow = tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + ms * Cfg::kMidCols;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t o[4] = {pk[c].x, pk[c].y, pk[c].z, pk[c].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                uint32_t w = dx == 0 ? o[e] : __shfl_down_sync(0xffffffffu, o[e], dx);
                if (dx > 0 && lane + dx >= 32) w = lgrp < 3 ? xb[((lgrp + 1) * 2 + (lane + dx - 32)) * 32 + c * 4 + e] : 0u;
                v[c * 4 + e] = w;
              }
            }
            ptx::tmem_st32(trow + dx * 32, v);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(mfull_leader + 8 * ms);      // count 2 x 4 warps
        }
      }
    }
  } else {
    // ------------------------------------------------------------ output epilogue (warps 6..9): q(q(q(acc) * scale) + t) -> staging -> TMA store
    const int lgrp = warp & 3;
    const int L = lgrp * 32 + lane;                  // output pixel x0 + L (L >= 126: scrap)
    const int sw = L & 7, swr = (L + 2) & 7;         // the residual sits 2 rows further down its slot (t row = pixel x0 - 2 + row)
    const bool lead_warp = warp == 6;
    const uint32_t a2empty_leader = ptx::mapa(a2empty, 0);
    uint8_t* my_row = stg_ptr + L * 128;
    uint32_t n2 = 0, tcnt = 0;
    uint32_t ord = 0;
    uint4 resid[8];                                  // this pixel's 64 residual channels of the row being finished
    auto load_resid = [&](uint32_t entry) {
      const uint8_t* res = tring_ptr + (entry % TS) * kSlotBytes + (L + 2) * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) resid[c] = *reinterpret_cast<const uint4*>(res + ((c ^ swr) << 4));
    };
    for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
      const ArsbItem it = arsb_decode(p, item);
      const int x0 = (it.sp * 2 + static_cast<int>(rank)) * kArsbStripW;
      const uint32_t nent = static_cast<uint32_t>(it.y1 - it.y0 + 4);          // t rows y0-2 .. y1+1 = ring entries tcnt .. tcnt+nent-1
      uint32_t rel = 0;                                                          // entries of this item already released
      for (int y = it.y0; y < it.y1; ++y, ++n2) {
        const uint32_t stage = n2 % AS;
        const uint32_t te = tcnt + static_cast<uint32_t>(y - (it.y0 - 2));     // ring entry of t row y
        // conv_2(y) complete => (in-order completion, issue order of the MMA warp) conv_1 is complete up to mid row y+2 (y+1 on
        // the item's first row): every t row <= y+3 (y+2) has landed and no MMA still reads t rows <= y+1 (y)
        ptx::mbar_wait(a2full + 8 * stage, (n2 / AS) & 1);
        ptx::tc_fence_after_sync();
        ptx::fence_proxy_async_smem();               // the t rows were written through the async proxy (TMA), read here through the generic one
        if (y == it.y0) load_resid(te);
        uint4 pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + Cfg::kAcc2Col + stage * 64 + h * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 sk = resid[h * 4 + q];
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = q * 8 + e * 2;
              const uint32_t sw32 = reinterpret_cast<const uint32_t*>(&sk)[e];
              const __half2 hs = *reinterpret_cast<const __half2*>(&sw32);
              const float f0 = epi_apply<EPI_SCALE_SKIP>(__uint_as_float(v[jj]), ap.scale, 0.f, __low2float(hs), 0);
              const float f1 = epi_apply<EPI_SCALE_SKIP>(__uint_as_float(v[jj + 1]), ap.scale, 0.f, __high2float(hs), 0);
              const __half2 hv = __floats2half2_rn(f0, f1);
              w[e] = h * 32 + jj < kN ? *reinterpret_cast<const uint32_t*>(&hv) : 0u;
            }
            pk[h * 4 + q] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        if (y + 1 < it.y1) load_resid(te + 1);       // next row's residual now, so that its slot can be released a row early
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(a2empty_leader + 8 * stage);
          // t ring slots are released HERE (count 4 = these warps), not by the MMA warp.  Through t row y on the item's first
          // row, through y+1 afterwards (its residual is in registers), everything on the last row.
          const uint32_t upto = y == it.y1 - 1 ? nent : static_cast<uint32_t>(y - it.y0) + (y == it.y0 ? 3u : 4u);
          for (; rel < upto; ++rel) ptx::mbar_arrive(tempty_t + 8 * ((tcnt + rel) % TS));
        }
        rel = y == it.y1 - 1 ? nent : static_cast<uint32_t>(y - it.y0) + (y == it.y0 ? 3u : 4u);   // all lanes track it
        // single staging tile: the previous row's store must have read it
        if (lead_warp) {
          if (ptx::elect_one()) ptx::bulk_wait_read<0>();
          __syncwarp();
        }
        ptx::named_bar_sync(2, 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) = pk[c];
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(3, 128);
        if (lead_warp) {
          if (ptx::elect_one()) {
            ptx::tma_store_4d(&maps.out, stg, 0, x0, y, it.n);                 // 126 pixels; clipped at the right edge of the tile
            ptx::bulk_commit();
          }
          __syncwarp();
        }
      }
      tcnt += nent;
    }
    if (lead_warp) {
      if (ptx::elect_one()) ptx::bulk_wait<0>();
      __syncwarp();
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

}  // namespace moe
