// The upsample convolutions (64 -> 256, PixelShuffle(2), +bias, PReLU: models.py:29-33) hold 68 % of a4's
// FLOPs, and a 128x64x16 tcgen05.mma cannot run at full rate: its 6 KB of shared-memory operands take 48.5
// cycles to fetch where the math takes 32 (profiles/r01_mma_rate_microbench.log).  This kernel gives those
// layers an N = 128 tile with the SAME shared-memory footprint per SM by pairing two CTAs (cta_group::2):
//   * the pair covers 256 pixels of an output row: CTA0 the 128-px strip 2s, CTA1 the strip 2s+1 (M = 256);
//   * each CTA keeps ONE 64-channel weight chunk (73.7 KB); the MMA reads B rows 0..63 from CTA0 and 64..127
//     from CTA1, so every input row is multiplied by two sub-pixel chunks per instruction: per MMA each SM
//     fetches 4 KB (A) + 2 KB (its half of B) for 64 cycles of math instead of 32;
//   * accumulators: 128 lanes x 128 columns per CTA and stage, 4 stages = all 512 TMEM columns;
//   * everything else is conv_tc.cuh: 130-px TMA row ring, nine shifted descriptor views, TMA-store
//     epilogue (two staging tiles = the two chunks of one row).
// Protocol across the pair: the leader (cluster rank 0) issues every MMA.  `full` and `tempty` barriers live
// in the leader (both CTAs' TMA loads complete_tx there; the 16 epilogue warps arrive there); `empty`
// and `tfull` exist in both CTAs and are signalled by multicast tcgen05.commit.
#pragma once
#include "conv_tc.cuh"

namespace moe {

struct PairCfg {
  static constexpr int kSlots = 6;
  static constexpr int kAccStages = 4;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kSmemBytes = 1024 + kSlots * kSlotBytes + kChunkImgBytes + 2 * kStageBytes + 1024;
};

// ---------------------------------------------------------------------------------------------------
// Item scheduler.  A launch is cut into items (a row segment of one strip pair of one plane for one chunk group);
// a pair keeps the weights of ONE chunk group resident, so each group has its own item sequence j = 0, 1, ...
// Dealing the items round-robin leaves SMs idle at the end of every launch — ncu showed sm__cycles_active between
// 6.5M and 8.1M cycles across the SMs of one fused-upsample launch (profiles/r01_final_kernels_ncu.txt) although every
// pair had the same number of rows: pairs do not run at the same speed.  So the pairs DRAW items instead: the leader
// CTA's producer warp takes the next j of its group from a global counter (one atomicAdd per item, one item ahead of
// its loads) and publishes the item to every warp role of both CTAs through a small shared-memory queue:
//   items[q] (int) + bars[q] (mbarrier, count 1), q = ordinal % kSchedQ, written locally and into the peer CTA
//   (st.shared::cluster + mbarrier.arrive.release.cluster), consumed with an acquire wait; item -1 ends the stream.
// The queue needs no "slot free" handshake: the publisher runs at most one item ahead of the TMA producer, and the
// slowest consumer (the P readers of the fused kernel) trails the producer by at most ring + accumulator + P stages
// = 12 rows = 4 items of the smallest possible size (1 row + 2 halo rows), so nobody is ever kSchedQ = 16 behind.
constexpr int kSchedQ = 16;
constexpr int kSchedGroups = 16;                  // counters [0, 16): next j of a chunk group; [16]: finished pairs
constexpr int kSchedInts = kSchedGroups + 1;

struct PairSched {
  uint32_t items, bars;       // shared::cta addresses of the queue in THIS CTA
  int group, groups;          // this pair's chunk group / number of groups
  int j_stride, j_count;      // pairs per group (= first j drawn dynamically, = static stride) / items per group
  int* ctr;
  int dynamic;
};

__device__ __forceinline__ PairSched sched_make(const ConvParams& p, uint32_t items, uint32_t bars, int pair, int npairs, int groups) {
  PairSched s;
  s.items = items; s.bars = bars;
  s.group = pair % groups; s.groups = groups;
  s.j_stride = npairs / groups; s.j_count = p.items / groups;
  s.ctr = p.sched; s.dynamic = p.dynamic;
  return s;
}
__device__ __forceinline__ void sched_init_bars(uint32_t bars) {      // one thread, before fence_mbar_init
  for (int i = 0; i < kSchedQ; ++i) ptx::mbar_init(bars + 8 * i, 1);
}
__device__ __forceinline__ int sched_item(const PairSched& s, int j) { return j < s.j_count ? s.group + s.groups * j : -1; }
// leader CTA's producer warp, converged: the j after `j` (the pair's first j is pair / groups)
__device__ __forceinline__ int sched_next_j(const PairSched& s, int j, int lane) {
  if (!s.dynamic) return j + s.j_stride;
  int nj = 0;
  if (lane == 0) nj = s.j_stride + atomicAdd(s.ctr + s.group, 1);
  return __shfl_sync(0xffffffffu, nj, 0);
}
__device__ __forceinline__ void sched_publish(const PairSched& s, uint32_t ord, int item) {   // ONE lane of the leader's producer warp
  const uint32_t q = ord % kSchedQ;
  ptx::st_shared_u32(s.items + 4 * q, static_cast<uint32_t>(item));
  ptx::st_shared_cluster_u32(ptx::mapa(s.items + 4 * q, 1), static_cast<uint32_t>(item));
  ptx::mbar_arrive(s.bars + 8 * q);                                   // release.cta: orders the local store
  ptx::mbar_arrive_cluster_release(ptx::mapa(s.bars + 8 * q, 1));     // release.cluster: orders the store into the peer
}
__device__ __forceinline__ int sched_take(const PairSched& s, uint32_t ord) {                 // any warp of either CTA, all lanes
  const uint32_t q = ord % kSchedQ;
  if (!ptx::mbar_wait_cluster(s.bars + 8 * q, (ord / kSchedQ) & 1)) return -1;     // aborted (ptx.cuh): end of the stream for this warp
  return static_cast<int>(ptx::ld_shared_u32(s.items + 4 * q));
}
// producer warps of both CTAs, converged: item number `ord` of this pair (ord 0 = the pair's first item, j = pair / groups)
__device__ __forceinline__ int sched_produce(const PairSched& s, bool leader, int lane, int& j, uint32_t ord) {
  if (!leader) return sched_take(s, ord);
  if (ord != 0) j = sched_next_j(s, j, lane);
  const int item = sched_item(s, j);
  if (lane == 0) sched_publish(s, ord, item);
  __syncwarp();
  return item;
}
// end of kernel, one thread of the leader CTA: the last pair to finish zeroes the counters for the next launch
__device__ __forceinline__ void sched_finish(const ConvParams& p, int npairs) {
  if (!p.dynamic) return;
  __threadfence();
  if (atomicAdd(p.sched + kSchedGroups, 1) == npairs - 1) {
    for (int i = 0; i < kSchedInts; ++i) p.sched[i] = 0;
    __threadfence();
  }
}
__device__ __forceinline__ void pair_debug(const ConvParams& p, int pair, uint64_t t0, int n_items) {
  if (!p.dbg) return;
  unsigned long long* d = p.dbg + static_cast<size_t>(pair) * 4;
  d[0] = t0; d[1] = ptx::globaltimer_ns(); d[2] = ptx::smid(); d[3] = static_cast<unsigned long long>(n_items);
}

// item -> (chunk group g, plane n, strip pair sp, rows)
// chunk groups of an upsample convolution: the r*r sub-pixel chunks two at a time (a pair computes N = 128 = two chunks per MMA);
// PixelShuffle(3) has nine, so its fifth group holds one real chunk and one idle half (10 % of the MMA work)
__device__ __host__ __forceinline__ int pair_groups(int r) { return (r * r + 1) / 2; }

__device__ __forceinline__ void pair_decode_item(const ConvParams& p, int item, int& g, int& n, int& sp, int& y0, int& y1) {
  const int groups = pair_groups(p.r);
  g = item % groups;
  int rest = item / groups;
  const int seg = rest % p.nseg;
  rest /= p.nseg;
  sp = rest % p.strips;          // p.strips holds the number of strip PAIRS here
  n = rest / p.strips;
  y0 = seg * p.seg_rows;
  y1 = min(p.H, y0 + p.seg_rows);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p, const __grid_constant__ ConvBias cb)
{
  using Cfg = PairCfg;
  constexpr int S = Cfg::kSlots, AS = Cfg::kAccStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ring = base;
  const uint32_t wsm = ring + S * kSlotBytes;
  const uint32_t stg = wsm + kChunkImgBytes;                       // two staging tiles (one per chunk), 1024-aligned
  const uint32_t bars = stg + 2 * kStageBytes;
  const uint32_t full = bars, empty = full + 8 * S, tfull = empty + 8 * S, tempty = tfull + 8 * AS;
  const uint32_t wbar = tempty + 8 * AS, wpeer = wbar + 8, dbar = wpeer + 8, tslot = dbar + 8;
  const uint32_t sq_items = bars + 768, sq_bars = bars + 832;       // item queue (kSchedQ ints + kSchedQ barriers)
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  uint8_t* stg_ptr = smem + (stg - base);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int groups = pair_groups(p.r), nchunks = p.r * p.r;
  const int g_fixed = pair % groups;              // npairs is a multiple of the group count (host), so a pair keeps its chunk group
  const int my_chunk = min(g_fixed * 2 + static_cast<int>(rank), nchunks - 1);   // the idle half of PixelShuffle(3)'s last group recomputes chunk 8
  const PairSched sc = sched_make(p, sq_items, sq_bars, pair, npairs, groups);
  const uint64_t t_start = p.dbg ? ptx::globaltimer_ns() : 0;
  int n_items = 0;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(full + 8 * i, 1); ptx::mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) { ptx::mbar_init(tfull + 8 * i, 1); ptx::mbar_init(tempty + 8 * i, 2 * kEpiWarps); }
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
  ptx::cluster_sync_all();                        // the peer's barriers are initialised before any remote arrive
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tslot_ptr;
  ptx::grid_dep_launch();
  if (warp == 0) {                                 // the weights are nobody's output: fetched before the dependency wait
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, kChunkImgBytes);
      const uint8_t* src = p.w_img + static_cast<size_t>(my_chunk) * kChunkImgBytes;
      for (int tap = 0; tap < 9; ++tap) ptx::bulk_load_1d(wsm + tap * 8192, src + tap * 8192, 8192, wbar);
    }
    __syncwarp();
  }
  ptx::grid_dep_wait();                            // the previous kernel on the stream has finished writing our input

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (each CTA loads its own strip)
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
          if (leader_cta) ptx::mbar_expect_tx(full + 8 * slot, 2 * kRowBytes);     // both CTAs' rows
          ptx::tma_load_4d_pair(ring + slot * kSlotBytes, &maps.in, ptx::mapa(full + 8 * slot, 0), 0, x0 - 1, yy, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (!leader_cta) {
      // ---------------------------------------------------------- peer: report "my weights are in smem" to the leader
      ptx::mbar_wait(wbar, 0);
      if (ptx::elect_one()) ptx::mbar_arrive_cluster(ptx::mapa(wpeer, 0));
      __syncwarp();
    } else {
      // ---------------------------------------------------------- leader: MMA issuer for the pair
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
            ptx::mma_commit_pair_mc(tfull + 8 * stage, 3);              // both CTAs' epilogues
            ptx::mma_commit_pair_mc(empty + 8 * ((cons + j) % S), 3);   // both CTAs' producers
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
      if (ptx::elect_one()) ptx::mma_commit_pair(dbar);     // drain
      __syncwarp();
      ptx::mbar_wait_drain(dbar, 0);
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9 of both CTAs)
    const int lgrp = warp & 3;                       // TMEM lanes 32*lgrp .. +31
    const int ch = (warp - 2) >> 2;                  // which of the pair's two chunks (64 accumulator columns)
    const int L = lgrp * 32 + lane;                  // pixel within this CTA's strip == staging row
    const bool lead_warp = (warp == 2);
    uint8_t* my_row = stg_ptr + ch * kStageBytes + L * 128;
    const int sw = L & 7;
    const uint32_t tempty_leader = ptx::mapa(tempty, 0);
    const int my_bias = min(g_fixed * 2 + ch, nchunks - 1) * 64;        // index into cb.v (constant bank)
    const CUtensorMap* omap0 = &maps.out[g_fixed * 2];
    const CUtensorMap* omap1 = &maps.out[min(g_fixed * 2 + 1, nchunks - 1)];
    const bool two_chunks = g_fixed * 2 + 1 < nchunks;
    uint32_t acc = 0;
    uint32_t ord = 0;
    for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
      int g, n, sp, y0, y1;
      pair_decode_item(p, item, g, n, sp, y0, y1);
      const int x0 = (sp * 2 + static_cast<int>(rank)) * kStripW;
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
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + 8 * stage);   // this warp is done with the TMEM stage
        // the staging tiles are single-buffered: the previous row's stores must have read them
        if (lead_warp) {
          if (ptx::elect_one()) ptx::bulk_wait_read<0>();
          __syncwarp();
        }
        ptx::named_bar_sync(1, kEpiThreads);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(my_row + ((c ^ sw) << 4)) = pk[c];
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(2, kEpiThreads);
        if (lead_warp) {
          if (ptx::elect_one()) {
            ptx::tma_store_4d(omap0, stg, 0, x0, y, n);
            if (two_chunks) ptx::tma_store_4d(omap1, stg + kStageBytes, 0, x0, y, n);
            ptx::bulk_commit();
          }
          __syncwarp();
        }
      }
    }
    if (lead_warp) {
      if (ptx::elect_one()) ptx::bulk_wait<0>();
      __syncwarp();
    }
  }
  __syncwarp();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync_all();                          // nobody frees TMEM / exits while the peer may still use it
  if (leader_cta && tid == 0) { sched_finish(p, npairs); pair_debug(p, pair, t_start, n_items); }
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------
// The same pairing for the 64 -> 64 trunk convolutions (conv_input2 and the twelve ARSB convs): M = 256,
// N = 64, each CTA holding HALF of the output channels' weights (32 rows of every tap).  Per MMA an SM now
// fetches 4 KB (A) + 1 KB (B half) instead of 4 + 2 KB: 40 instead of 48.5 cycles at the operand-fetch limit.
// Epilogue exactly as in conv_tc.cuh (plain / PReLU / x scale + residual via TMA / + bias, PReLU with the PixelShuffle
// store map: the nine sub-pixel chunks of Net3x's 64 -> 576 upsample conv run as nine groups of pairs).
struct PairTrunkCfg {
  static constexpr int kSlots = 6;
  static constexpr int kAccStages = 4;
  static constexpr int kOutStages = 4;                           // staging tiles; the residual is prefetched 2 rows ahead
  static constexpr int kSkipAhead = 2;
  static constexpr uint32_t kWBytes = 9 * 32 * 128;              // this CTA's half of the weights
  static constexpr uint32_t kTmemCols = kAccStages * 64;
  static constexpr uint32_t kSmemBytes = 1024 + kSlots * kSlotBytes + kWBytes + kOutStages * kStageBytes + 1024;
};

// item -> (plane n, strip pair sp, rows); p.strips = number of strip pairs.  With PixelShuffle (r > 1) the r*r
// 64-channel chunks are the fastest-varying part of the item index and a pair keeps ONE chunk for its lifetime
// (the host launches a multiple of r*r pairs), so `item / (r*r)` enumerates the (plane, strip pair, segment) triples.
__device__ __forceinline__ void pair_trunk_decode(const ConvParams& p, int item, int& n, int& sp, int& y0, int& y1) {
  item /= (p.r * p.r);
  const int seg = item % p.nseg;
  int rest = item / p.nseg;
  sp = rest % p.strips;
  n = rest / p.strips;
  y0 = seg * p.seg_rows;
  y1 = min(p.H, y0 + p.seg_rows);
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv3x3_pair_trunk_kernel(const __grid_constant__ ConvMaps maps, const ConvParams p)
{
  using Cfg = PairTrunkCfg;
  constexpr int S = Cfg::kSlots, AS = Cfg::kAccStages, OS = Cfg::kOutStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ring = base;
  const uint32_t wsm = ring + S * kSlotBytes;
  const uint32_t stg = wsm + Cfg::kWBytes;                          // 36 KB of weights keep this 1024-aligned
  const uint32_t bars = stg + OS * kStageBytes;
  const uint32_t full = bars, empty = full + 8 * S, tfull = empty + 8 * S, tempty = tfull + 8 * AS;
  const uint32_t skfull = tempty + 8 * AS, wbar = skfull + 8 * OS, wpeer = wbar + 8, dbar = wpeer + 8, tslot = dbar + 8;
  const uint32_t sq_items = bars + 512, sq_bars = bars + 576;       // item queue (kSchedQ ints + kSchedQ barriers)
  volatile uint32_t* tslot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tslot - base));
  uint8_t* stg_ptr = smem + (stg - base);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int chunk = pair % (p.r * p.r);            // npairs is a multiple of r*r (host)
  const CUtensorMap* omap = &maps.out[chunk];
  const PairSched sc = sched_make(p, sq_items, sq_bars, pair, npairs, p.r * p.r);
  const uint64_t t_start = p.dbg ? ptx::globaltimer_ns() : 0;
  int n_items = 0;

  if (tid == 0) {
    for (int i = 0; i < S; ++i) { ptx::mbar_init(full + 8 * i, 1); ptx::mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < AS; ++i) { ptx::mbar_init(tfull + 8 * i, 1); ptx::mbar_init(tempty + 8 * i, 2 * kEpiWarps); }
    for (int i = 0; i < OS; ++i) ptx::mbar_init(skfull + 8 * i, 1);
    ptx::mbar_init(wbar, 1);
    ptx::mbar_init(wpeer, 1);
    ptx::mbar_init(dbar, 1);
    sched_init_bars(sq_bars);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&maps.in);
    ptx::prefetch_tmap(omap);
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
      ptx::mbar_expect_tx(wbar, Cfg::kWBytes);
      for (int tap = 0; tap < 9; ++tap)      // output channels 32*rank .. +31 of every tap
        ptx::bulk_load_1d(wsm + tap * 4096, p.w_img + static_cast<size_t>(chunk) * kChunkImgBytes + tap * 8192 + rank * 4096, 4096, wbar);
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
      int n, sp, y0, y1;
      pair_trunk_decode(p, item, n, sp, y0, y1);
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
      constexpr uint32_t idesc = ptx::idesc_f16_f32(256, 64);
      const uint64_t bdesc0 = ptx::smem_desc_sw128(wsm, 1024, 0);
      const uint64_t adesc0 = ptx::smem_desc_sw128(ring, 1024, 0);
      ptx::mbar_wait(wbar, 0);
      ptx::mbar_wait(wpeer, 0);
      ptx::tc_fence_after_sync();
      uint32_t cons = 0, acc = 0;
      uint32_t ord = 0;
      for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
        int n, sp, y0, y1;
        pair_trunk_decode(p, item, n, sp, y0, y1);
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
                  ptx::mma_f16_ss_pair(d_tmem, arow + (dx * 8 + k * 2), bdesc0 + ((dy * 3 + dx) * 256 + k * 2), idesc, p.center_only ? (k != 0) : ((dy | dx | k) != 0));
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
      if (ptx::elect_one()) ptx::mma_commit_pair(dbar);
      __syncwarp();
      ptx::mbar_wait_drain(dbar, 0);
    }
  } else {
    // ------------------------------------------------------------ epilogue (as conv_tc.cuh; tempty lives in the leader)
    const int lgrp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int L = lgrp * 32 + lane;
    const bool lead_warp = (warp == 2);
    constexpr bool has_skip = EPI == EPI_SCALE_SKIP;
    uint8_t* my_row = stg_ptr + L * 128;
    const int sw = L & 7;
    const uint32_t tempty_leader = ptx::mapa(tempty, 0);
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = EPI == EPI_BIAS_PRELU ? __ldg(p.bias + chunk * 64 + half * 32 + j) : 0.f;
    uint32_t acc = 0;
    // residual prefetch cursor: walks the same row sequence kSkipAhead rows ahead of the epilogue
    int c_item = -1, c_n = 0, c_sp = 0, c_y = 0, c_y1 = 0;
    uint32_t c_idx = 0, c_ord = 0;
    bool c_ok = has_skip && lead_warp;
    if (c_ok) { c_item = sched_take(sc, c_ord++); c_ok = c_item >= 0; }
    if (c_ok) pair_trunk_decode(p, c_item, c_n, c_sp, c_y, c_y1);
    auto prefetch_skip = [&]() {            // lead warp, converged: issue the residual load of row c_idx and advance
      if (!c_ok) return;
      const uint32_t t = c_idx % OS;
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(skfull + 8 * t, kStageBytes);
        ptx::tma_load_4d(stg + t * kStageBytes, &maps.skip, skfull + 8 * t, 0, (c_sp * 2 + static_cast<int>(rank)) * kStripW, c_y, c_n);
      }
      __syncwarp();
      ++c_idx;
      if (++c_y >= c_y1) {                   // the item after this one was published before this one's loads began
        c_item = sched_take(sc, c_ord++);
        c_ok = c_item >= 0;
        if (c_ok) pair_trunk_decode(p, c_item, c_n, c_sp, c_y, c_y1);
      }
    };
    if (lead_warp) for (int i = 0; i < Cfg::kSkipAhead; ++i) prefetch_skip();
    uint32_t ord = 0;
    for (int item = sched_take(sc, ord++); item >= 0; item = sched_take(sc, ord++)) {
      int n, sp, y0, y1;
      pair_trunk_decode(p, item, n, sp, y0, y1);
      const int x0 = (sp * 2 + static_cast<int>(rank)) * kStripW;
      for (int y = y0; y < y1; ++y, ++acc) {
        const uint32_t stage = acc % AS;
        const uint32_t os = acc % OS;
        uint8_t* row = my_row + os * kStageBytes;
        if (has_skip) {
          ptx::mbar_wait(skfull + 8 * os, (acc / OS) & 1);
        } else {
          if (lead_warp) {
            if (ptx::elect_one()) ptx::bulk_wait_read<OS - 1>();
            __syncwarp();
          }
          ptx::named_bar_sync(1, kEpiThreads);
        }
        ptx::mbar_wait(tfull + 8 * stage, (acc / AS) & 1);
        ptx::tc_fence_after_sync();
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(lgrp * 32) << 16) + stage * 64 + half * 32, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tempty_leader + 8 * stage);
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
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(2, kEpiThreads);
        if (lead_warp) {
          if (ptx::elect_one()) {
            ptx::tma_store_4d(omap, stg + os * kStageBytes, 0, x0, y, n);
            ptx::bulk_commit();
            // the tile the next prefetch targets was stored kOutStages - kSkipAhead rows ago: everything older
            // than the (kOutStages - kSkipAhead) most recent stores must have read its tile
            if (has_skip) ptx::bulk_wait_read<OS - Cfg::kSkipAhead>();
          }
          __syncwarp();
          prefetch_skip();
        }
      }
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
