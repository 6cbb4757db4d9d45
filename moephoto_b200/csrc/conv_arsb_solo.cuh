// The fused residual block (conv_arsb.cuh) on ONE CTA per SM instead of a CTA pair — an A/B variant (engine switch bit 8), slower than
// the pair form, kept because it is what located the limit of the 64 -> 64 layers (profiles/r02_arsb_experiments.txt).
//
// Built to test the hypothesis that the N = 64 CTA-pair kernels (49-54 cycles per MMA for 32 of math) wait for the half of the B
// operand that cta_group::2 fetches from the partner SM.  Here every SM holds ALL 64 output channels of both convolutions (2 x 72 KB);
// conv_1 reads its A operand (the t rows, TMA-loaded) from shared memory, conv_2 reads its A operand (the mid rows, written by the
// mid epilogue with tcgen05.st as three shifted views) from tensor memory.  Measured in the kernel: the MMAs issue at exactly the
// single-SM microbenchmark rates (49 cycles .ss, 38 .ts) — and the kernel is still slower than the pair form, because
//   * tcgen05.mma blocks the issuing thread while the short queue is full and every wait / commit around the MMAs costs that thread
//     100-250 cycles, so the tensor pipe idles through the bookkeeping (hence two issuing warps that take turns, below), and
//   * with 144 KB of weights there is no room for a staging tile: the result row leaves as 16-byte global stores, one 128-byte line
//     per thread, and those 2 000 L1 wavefronts per row compete with the 2 300 wavefronts of the tensor core's operand fetch for
//     the one-wavefront-per-cycle shared-memory data pipe (without them: 3 760 instead of 4 700 cycles per row).
// The per-SM operand bytes per MMA are the same as in the pair form (4 KB A + 2 KB B against 4 KB A + 1 KB B + 1 KB served to the
// partner): the B exchange was never the limit, the A fetch is.
// 216 KB of shared memory (4-slot t ring released by the MMA warp's commits, residual from global memory = an L2 hit), all 512
// TMEM columns.  Rounding points and accumulation order are those of arsb_pair_kernel and of the two-launch form: bit-identical.
// Warp roles: 0 TMA producer + item scheduler, 1 conv_1 issuer, 2..5 mid epilogue, 6..9 output epilogue, 10 conv_2 issuer.
#pragma once
#include "conv_arsb.cuh"

namespace moe {

constexpr int kArsbSoloThreads = kConvThreads + 32;     // warp 10: the second MMA issuer

struct ArsbSoloCfg {
  static constexpr int kTSlots = 4;
  static constexpr int kMSlots = 4;                               // mid rows in TMEM
  static constexpr uint32_t kMidCols = 96;                        // 3 views x 64 fp16 per mid row
  static constexpr uint32_t kAcc1Col = 4 * kMidCols, kAcc2Col = 4 * kMidCols + 64;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kXchBytes = 2048;
  static constexpr uint32_t kSmemBytes = 1024 + kTSlots * kSlotBytes + kXchBytes + 2 * kChunkImgBytes + 1024;
};

// item -> (plane, strip, rows); p.strips = number of 126-px strips
__device__ __forceinline__ ArsbItem arsb_solo_decode(const ConvParams& p, int item) {
  ArsbItem it;
  const int seg = item % p.nseg;
  const int rest = item / p.nseg;
  it.sp = rest % p.strips;
  it.n = rest / p.strips;
  it.y0 = seg * p.seg_rows;
  it.y1 = min(p.H, it.y0 + p.seg_rows);
  it.m_lo = max(it.y0 - 1, 0);
  it.m_hi = min(it.y1, p.H - 1);
  return it;
}

template <int KS>
__global__ void __launch_bounds__(kArsbSoloThreads, 1)
arsb_solo_kernel(const __grid_constant__ ArsbMaps maps, const ArsbParams ap)
{
  using Cfg = ArsbSoloCfg;
  constexpr int TS = Cfg::kTSlots, MS = Cfg::kMSlots;
  const ConvParams& p = ap.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t tring = base;
  const uint32_t xch = tring + TS * kSlotBytes;
  const uint32_t w1sm = xch + Cfg::kXchBytes;
  const uint32_t w2sm = w1sm + kChunkImgBytes;
  const uint32_t bars = w2sm + kChunkImgBytes;
  const uint32_t tfull_t = bars, tempty_t = tfull_t + 8 * TS;
  const uint32_t mfull = tempty_t + 8 * TS, mempty = mfull + 8 * MS;
  const uint32_t a1full = mempty + 8 * MS, a1empty = a1full + 8, a2full = a1empty + 8, a2empty = a2full + 8;
  const uint32_t wbar = a2empty + 8, dbar = wbar + 8, tslot = dbar + 8, dbar2 = tslot + 8, tok1 = dbar2 + 8, tok2 = tok1 + 8;
  const uint32_t sq_items = bars + 512, sq_bars = bars + 576;     // item queue: kSchedQ ints + kSchedQ barriers
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  uint32_t* xch_ptr = reinterpret_cast<uint32_t*>(smem + (xch - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint64_t t_start = p.dbg ? ptx::globaltimer_ns() : 0;
  int n_items = 0;

  if (tid == 0) {
    for (int i = 0; i < TS; ++i) { ptx::mbar_init(tfull_t + 8 * i, 1); ptx::mbar_init(tempty_t + 8 * i, 1); }
    for (int i = 0; i < MS; ++i) { ptx::mbar_init(mfull + 8 * i, 4); ptx::mbar_init(mempty + 8 * i, 1); }
    ptx::mbar_init(a1full, 1); ptx::mbar_init(a1empty, 4);
    ptx::mbar_init(a2full, 1); ptx::mbar_init(a2empty, 4);
    ptx::mbar_init(wbar, 1);
    ptx::mbar_init(dbar, 1);
    ptx::mbar_init(dbar2, 1);
    ptx::mbar_init(tok1, 1); ptx::mbar_init(tok2, 1);
    for (int i = 0; i < kSchedQ; ++i) ptx::mbar_init(sq_bars + 8 * i, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&maps.in);
  }
  if (warp == 1) ptx::tmem_alloc(tslot, Cfg::kTmemCols);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tslot_ptr;
  ptx::grid_dep_launch();
  if (warp == 0) {                                 // weights are nobody's output: fetched before the dependency wait
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, 2 * kChunkImgBytes);
      for (int tap = 0; tap < 9; ++tap) {
        ptx::bulk_load_1d(w1sm + tap * 8192, p.w_img + tap * 8192, 8192, wbar);
        ptx::bulk_load_1d(w2sm + tap * 8192, ap.w2_img + tap * 8192, 8192, wbar);
      }
    }
    __syncwarp();
  }
  ptx::grid_dep_wait();

  // the item stream: warp 0 draws item numbers from the launch's counter (the first one is the CTA's index) and publishes them
  // through a queue of kSchedQ entries; -1 ends the stream.  Consumers never trail by kSchedQ items (an item is >= 16 rows).
  auto take = [&](uint32_t ord) -> int {
    const uint32_t q = ord % kSchedQ;
    if (!ptx::mbar_wait(sq_bars + 8 * q, (ord / kSchedQ) & 1)) return -1;
    return static_cast<int>(ptx::ld_shared_u32(sq_items + 4 * q));
  };

  if (warp == 0) {
    // ------------------------------------------------------------ scheduler + TMA producer
    const int ncta = static_cast<int>(gridDim.x);
    auto draw = [&](uint32_t ord) -> int {        // whole warp, converged
      int item = 0;
      if (lane == 0) {
        item = ord == 0 ? static_cast<int>(blockIdx.x) : (p.dynamic ? ncta + atomicAdd(p.sched, 1) : static_cast<int>(blockIdx.x) + static_cast<int>(ord) * ncta);
        if (item >= p.items) item = -1;
        ptx::st_shared_u32(sq_items + 4 * (ord % kSchedQ), static_cast<uint32_t>(item));
        ptx::mbar_arrive(sq_bars + 8 * (ord % kSchedQ));
      }
      return __shfl_sync(0xffffffffu, item, 0);
    };
    uint32_t ld = 0, ord = 0;
    int item = draw(ord++);
    while (item >= 0) {
      const int next_item = draw(ord++);           // drawn one item ahead of the loads
      const ArsbItem it = arsb_solo_decode(p, item);
      const int x0 = it.sp * kArsbStripW;
      ++n_items;
      item = next_item;
      for (int yy = it.y0 - 2; yy <= it.y1 + 1; ++yy, ++ld) {
        const uint32_t slot = ld % TS;
        ptx::mbar_wait(tempty_t + 8 * slot, ((ld / TS) & 1) ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(tfull_t + 8 * slot, kRowBytes);
          ptx::tma_load_4d(tring + slot * kSlotBytes, &maps.in, tfull_t + 8 * slot, 0, x0 - 2, yy, it.n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == 10) {
    // -------------------------------------------------------------- the two MMA issuers: warp 1 conv_1 (.ss: t rows and weights in
    // shared memory), warp 10 conv_2 (.ts: mid rows in tensor memory).
    // A thread's tcgen05.mma blocks it while the (short) queue is full, and every wait and commit around the MMAs costs it 100-250
    // cycles: ONE thread issuing both convolutions left the tensor pipe idle 45 % of the time (ncu, profiles/r02_arsb_experiments.txt).
    // Two threads issuing freely interleave their MMAs one by one, which the pipe executes no faster (65 cycles per MMA).  So the
    // two threads take turns: both walk the same sequence of steps (conv_1 two mid rows ahead of conv_2), a thread does its waits for
    // its next step while the other issues, then waits for the token, issues its 36 MMAs, passes the token and only then commits and
    // releases ring slots.  The token orders nothing but the issue; the data dependencies are the barriers'.
    const bool is_c1 = warp == 1;
    constexpr uint32_t idesc = ptx::idesc_f16_f32(128, 64);
    const uint64_t b1 = ptx::smem_desc_sw128(w1sm, 1024, 0);
    const uint64_t b2 = ptx::smem_desc_sw128(w2sm, 1024, 0);
    const uint64_t at0 = ptx::smem_desc_sw128(tring, 1024, 0);
    const uint32_t tok_mine = is_c1 ? tok1 : tok2, tok_other = is_c1 ? tok2 : tok1;
    ptx::mbar_wait(wbar, 0);
    ptx::tc_fence_after_sync();
    uint32_t tcnt = 0, mcnt = 0, n1 = 0, n2 = 0, twaited = 0, mwaited = 0, tokw = 0;
    bool prev_mine = is_c1;                          // the sequence starts with a conv_1
    uint32_t ord = 0;
    const bool prof = p.center_only && p.dbg && blockIdx.x < 8;
    long long w_data = 0, w_acc = 0, w_tok = 0, t_mma = 0, t_tail = 0, rows = 0;
    const long long t_begin = clock64();
    for (int item = take(ord++); item >= 0; item = take(ord++)) {
      const ArsbItem it = arsb_solo_decode(p, item);
      const uint32_t nent = static_cast<uint32_t>(it.y1 - it.y0 + 4);        // t rows y0-2 .. y1+1 = ring entries tcnt .. tcnt+nent-1
      int next_m = it.m_lo;
      uint32_t trel = 0;                                                     // entries of this item handed back to the producer
      for (int y = it.y0; y < it.y1; ++y) {
        const int want = min(y + 2, it.m_hi);
        for (; next_m <= want; ++next_m) {
          // ---- step: conv_1 of mid row m
          if (!is_c1) { prev_mine = false; continue; }
          const int m = next_m;
          const uint32_t e0 = tcnt + static_cast<uint32_t>(m - 1 - (it.y0 - 2));
          long long tic = prof ? clock64() : 0;
          while (twaited < e0 + 3) { ptx::mbar_wait(tfull_t + 8 * (twaited % TS), (twaited / TS) & 1); ++twaited; }
          if (prof) { const long long t = clock64(); w_data += t - tic; tic = t; }
          ptx::mbar_wait(a1empty, (n1 & 1) ^ 1);
          if (prof) { const long long t = clock64(); w_acc += t - tic; tic = t; }
          if (!prev_mine) { ptx::mbar_wait(tok_mine, tokw & 1); ++tokw; }
          ptx::tc_fence_after_sync();
          if (prof) { const long long t = clock64(); w_tok += t - tic; tic = t; }
          const uint32_t d_tmem = tmem_base + Cfg::kAcc1Col;
          const uint64_t r0 = at0 + static_cast<uint64_t>((e0 % TS) * (kSlotBytes >> 4));
          const uint64_t r1 = at0 + static_cast<uint64_t>(((e0 + 1) % TS) * (kSlotBytes >> 4));
          const uint64_t r2 = at0 + static_cast<uint64_t>(((e0 + 2) % TS) * (kSlotBytes >> 4));
          const bool pass = m + 1 > want;                                    // the next step is a conv_2
          if (ptx::elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint64_t arow = dy == 0 ? r0 : (dy == 1 ? r1 : r2);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                  ptx::mma_f16_ss(d_tmem, arow + (dx * 8 + k * 2), b1 + ((dy * 3 + dx) * 512 + k * 2), idesc, (dy | dx | k) != 0);
                }
            }
            if (pass) ptx::mbar_arrive(tok_other);
            ptx::mma_commit(a1full);
          }
          __syncwarp();
          if (prof) { const long long t = clock64(); t_mma += t - tic; tic = t; ++rows; }
          ++n1;
          prev_mine = true;
          // t row m-1 was read for the last time (after the item's last conv_1: every remaining row).  A row no MMA reads (outside
          // the image) must at least have LANDED before its slot is loaded again.
          const uint32_t upto = m == it.m_hi ? nent : e0 - tcnt;
          for (; trel <= upto && trel < nent; ++trel) {
            while (twaited <= tcnt + trel) { ptx::mbar_wait(tfull_t + 8 * (twaited % TS), (twaited / TS) & 1); ++twaited; }
            if (ptx::elect_one()) ptx::mma_commit(tempty_t + 8 * ((tcnt + trel) % TS));
            __syncwarp();
          }
          if (prof) t_tail += clock64() - tic;
        }
        // ---- step: conv_2 of output row y
        if (is_c1) { prev_mine = false; continue; }
        const int lo = max(y - 1, it.m_lo), hi = min(y + 1, it.m_hi);
        const uint32_t ehi = mcnt + static_cast<uint32_t>(hi - it.m_lo);
        long long tic = prof ? clock64() : 0;
        while (mwaited <= ehi) { ptx::mbar_wait(mfull + 8 * (mwaited % MS), (mwaited / MS) & 1); ++mwaited; }
        if (prof) { const long long t = clock64(); w_data += t - tic; tic = t; }
        ptx::mbar_wait(a2empty, (n2 & 1) ^ 1);
        if (prof) { const long long t = clock64(); w_acc += t - tic; tic = t; }
        if (!prev_mine) { ptx::mbar_wait(tok_mine, tokw & 1); ++tokw; }
        ptx::tc_fence_after_sync();
        if (prof) { const long long t = clock64(); w_tok += t - tic; tic = t; }
        const uint32_t d_tmem = tmem_base + Cfg::kAcc2Col;
        // the next step is a conv_1 unless no mid row is left to compute; the first step of the next item always is
        const bool pass = y + 1 == it.y1 || next_m <= min(y + 3, it.m_hi);
        if (ptx::elect_one()) {
          bool first = true;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int m = y + dy - 1;
            if (m < lo || m > hi) continue;                                   // zero padding above / below the tile
            const uint32_t e = mcnt + static_cast<uint32_t>(m - it.m_lo);
            const uint32_t trow = tmem_base + (e % MS) * Cfg::kMidCols;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                ptx::mma_f16_ts(d_tmem, trow + dx * 32 + k * 8, b2 + ((dy * 3 + dx) * 512 + k * 2), idesc, first ? 0u : 1u);
                first = false;
              }
          }
          if (pass) ptx::mbar_arrive(tok_other);
          ptx::mma_commit(a2full);
          if (y - 1 >= it.m_lo) ptx::mma_commit(mempty + 8 * ((mcnt + static_cast<uint32_t>(y - 1 - it.m_lo)) % MS));
          if (y == it.y1 - 1)
            for (int m = max(y, it.m_lo); m <= it.m_hi; ++m) ptx::mma_commit(mempty + 8 * ((mcnt + static_cast<uint32_t>(m - it.m_lo)) % MS));
        }
        __syncwarp();
        if (prof) { t_mma += clock64() - tic; ++rows; }
        ++n2;
        prev_mine = true;
      }
      tcnt += nent;
      mcnt += static_cast<uint32_t>(it.m_hi - it.m_lo + 1);
    }
    const uint32_t my_dbar = is_c1 ? dbar : dbar2;
    if (ptx::elect_one()) ptx::mma_commit(my_dbar);
    __syncwarp();
    ptx::mbar_wait_drain(my_dbar, 0);
    if (prof && lane == 0) {
      unsigned long long* d = p.dbg + 74 * 4 + blockIdx.x * 12 + (is_c1 ? 0 : 6);
      d[0] = rows; d[1] = clock64() - t_begin; d[2] = w_data + (static_cast<long long>(w_acc) << 32); d[3] = w_tok; d[4] = t_mma; d[5] = t_tail;
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ mid epilogue (warps 2..5): conv_1 accumulator -> three views in TMEM
    const int lgrp = warp & 3;
    const int L = lgrp * 32 + lane;                  // mid position: pixel x0 - 1 + L
    uint32_t n1 = 0;
    uint32_t ord = 0;
    for (int item = take(ord++); item >= 0; item = take(ord++)) {
      const ArsbItem it = arsb_solo_decode(p, item);
      const int x = it.sp * kArsbStripW - 1 + L;
      const bool inside = x >= 0 && x < p.W;         // conv_2 sees zeros outside the tile
      for (int m = it.m_lo; m <= it.m_hi; ++m, ++n1) {
        ptx::mbar_wait(a1full, n1 & 1);
        ptx::tc_fence_after_sync();
        uint4 pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + Cfg::kAcc1Col + h * 32, v);
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
              w[e] = inside ? *reinterpret_cast<const uint32_t*>(&hv) : 0u;
            }
            pk[h * 4 + q] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a1empty);                              // the accumulator back to the MMA warp
        // view dx: THIS lane must hold mid position L + dx (conv_arsb.cuh)
        uint32_t* xb = xch_ptr + (n1 & 1) * 256;
        if (lane < 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(xb + (lgrp * 2 + lane) * 32 + c * 4) = pk[c];
        }
        ptx::named_bar_sync(1, 128);
        const uint32_t ms = n1 % MS;
        ptx::mbar_wait(mempty + 8 * ms, ((n1 / MS) & 1) ^ 1);                  // conv_2 has finished with this TMEM row's previous tenant
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
        if (lane == 0) ptx::mbar_arrive(mfull + 8 * ms);
      }
    }
  } else {
    // ------------------------------------------------------------ output epilogue (warps 6..9): q(q(q(acc) * scale) + t) -> global
    const int lgrp = warp & 3;
    const int L = lgrp * 32 + lane;                  // output pixel x0 + L (L >= 126: scrap)
    uint32_t n2 = 0;
    uint32_t ord = 0;
    for (int item = take(ord++); item >= 0; item = take(ord++)) {
      const ArsbItem it = arsb_solo_decode(p, item);
      const int x = it.sp * kArsbStripW + L;
      const bool mine = L < kArsbStripW && x < p.W;
      const size_t pix0 = (static_cast<size_t>(it.n) * p.H + it.y0) * p.W + (mine ? x : 0);
      const uint4* res = reinterpret_cast<const uint4*>(p.in + pix0 * 64);     // this pixel's 128-byte line of t, row y0
      uint4* dst = reinterpret_cast<uint4*>(p.out + pix0 * 64);
      const size_t row_u4 = static_cast<size_t>(p.W) * 8;                        // uint4 per image row
      for (int y = it.y0; y < it.y1; ++y, ++n2, res += row_u4, dst += row_u4) {
        uint4 resid[8];                              // issued before the wait: an L2 hit hidden behind it
#pragma unroll
        for (int c = 0; c < 8; ++c) resid[c] = mine ? __ldg(res + c) : make_uint4(0, 0, 0, 0);
        ptx::mbar_wait(a2full, n2 & 1);
        ptx::tc_fence_after_sync();
        uint4 pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + Cfg::kAcc2Col + h * 32, v);
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
              w[e] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            pk[h * 4 + q] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a2empty);
        if (mine) {
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c] = pk[c];
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    if (p.dynamic) {                                // the last CTA to finish zeroes the launch's counter block
      __threadfence();
      if (atomicAdd(p.sched + kSchedGroups, 1) == static_cast<int>(gridDim.x) - 1) {
        for (int i = 0; i < kSchedInts; ++i) p.sched[i] = 0;
        __threadfence();
      }
    }
    if (p.dbg && blockIdx.x < 74) pair_debug(p, blockIdx.x, t_start, n_items);
  }
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace moe
