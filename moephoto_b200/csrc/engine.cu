// C ABI of the engine (include/moephoto_b200.h): model loading, the per-tile network schedule, the
// on-device doCrop, and the frame conversions.  Host-side C++ only orchestrates kernel launches.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/moephoto_b200.h"
#include "blob.h"
#include "conv_tc.cuh"
#include "conv_pair.cuh"
#include "conv_pair_head.cuh"
#include "conv_arsb.cuh"
#include "conv_arsb_solo.cuh"
#include "kernels_simt.cuh"
#include "head_tc.cuh"

using namespace moe;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define MOE_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) return fail(MOE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kHaloLR = 16;   // receptive-field radius of the whole net is 15.75 LR px
constexpr int kNumBufs = 5;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct MoeEngine {
  int device = 0;
  int sm_count = 0;
  EncodeTiledFn encode = nullptr;
  std::atomic<int64_t> launches{0};
  int simt = 0;
  bool smem_attr_set = false, head_attr_set = false, pair_attr_set = false;
  int no_pair = 0;         // 1 = keep every conv on the single-CTA kernel (A/B switch)
  int no_pair_trunk = 0;   // 1 = only the 64->64 convs stay on the single-CTA kernel
  int no_fuse = 0;         // 1 = last upsample conv and heads stay separate kernels (conv3x3_pair_kernel + head_tc_kernel)
  int arsb_smem_mid = 0;   // 1 = the fused residual block keeps its mid rows in shared memory (.ss conv_2) instead of TMEM (.ts)
  int full_k = 0;          // 1 = the 48-filter models issue all four K steps per tap like the 64-filter ones (test switch: same bits)
  int arsb_solo = 0;       // 1 = the fused residual block on single CTAs with full weights per SM (conv_arsb_solo.cuh) instead of CTA pairs
  int no_arsb = 0;         // 1 = every residual block as two launches of the trunk kernel instead of arsb_pair_kernel (A/B switch)
  bool arsb_attr_set = false;
  bool arsb_solo_attr_set = false;
  int bias_fused = 0;      // 1 = biased convolutions round once, q(conv + bias): the half model executed on the CPU (goldens); 0 = the GPU's two ops
  int static_sched = 0;    // 1 = pair kernels deal their items round-robin instead of drawing them (conv_pair.cuh, item scheduler)
  // the item scheduler's counters: a ring of kSchedRing blocks of kSchedInts ints, one block per pair-kernel launch (the
  // kernel's last pair zeroes its block again), so launches on DIFFERENT streams never draw each other's items
  int* d_sched = nullptr;
  std::atomic<uint32_t> sched_seq{0};
  std::mutex host_mutex;   // moe_enhance_host: the engine-owned staging buffers serve one call at a time
  size_t dbg_bytes = 0;
  unsigned long long* dbg = nullptr;   // moe_engine_debug_buffer: per-pair {start ns, end ns, SM id, items} of the LAST pair-kernel launch
  bool pair_head_attr_set = false;
  bool pair_trunk_attr_set = false;
  // optional per-launch CUDA-event timing (moe_engine_profile): class 0 conv_input, 1 conv3x3 r=1, 2 heads, 3 upsample conv3x3
  bool profiling = false;
  struct Span { int cls; cudaEvent_t a, b; double work; };
  std::vector<Span> spans;
  double prof_ms[MOE_PROFILE_CLASSES] = {0}, prof_work[MOE_PROFILE_CLASSES] = {0};
  int64_t prof_n[MOE_PROFILE_CLASSES] = {0};
  // grown-on-demand device buffers of moe_enhance_host: raw in, planar in, canvas, raw out, workspace
  cudaStream_t copy_stream = nullptr;   // moe_enhance_host: conversion + device->host copy of finished canvas columns
  cudaEvent_t copy_event = nullptr;
  void* buf[kNumBufs] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[kNumBufs] = {0, 0, 0, 0, 0};
};

struct MoeModel {
  MoeEngine* e = nullptr;
  int arch = 0, feat = 64, n_up = 0, r = 0, scale = 1;
  uint8_t* d_blob = nullptr;
  const float* first_w = nullptr;
  float scalars[32] = {0};
  const uint8_t* trunk_img[13] = {nullptr};
  const uint8_t* up_img[8] = {nullptr};     // [4*branch + stage]
  const float* up_bias[8] = {nullptr};
  ConvBias up_bias_h[8];                    // host mirror: the CTA-pair kernels take the bias as a kernel parameter (conv_tc.cuh)
  const float* frm[3] = {nullptr, nullptr, nullptr};   // MoeNet_lite2's FRM gates
  const float* head_w[2] = {nullptr, nullptr};
  uint8_t* d_head_img = nullptr;   // [2][16 rows][128 B] swizzled fp16 image of the two head filters (head_tc.cuh)
};

namespace {

struct Guard {   // make the engine's device current for the duration of a call
  int prev = -1;
  bool ok = true;
  explicit Guard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

// RAII bracket: records an event pair around one kernel launch when profiling is on
struct Timed {
  MoeEngine* e; cudaStream_t st; int idx = -1;
  Timed(MoeEngine* e_, cudaStream_t st_, int cls, double work) : e(e_), st(st_) {
    if (!e->profiling) return;
    MoeEngine::Span s{cls, nullptr, nullptr, work};
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    cudaEventRecord(s.a, st);
    e->spans.push_back(s);
    idx = static_cast<int>(e->spans.size()) - 1;
  }
  ~Timed() { if (idx >= 0) cudaEventRecord(e->spans[idx].b, st); }
};

constexpr uint32_t kSchedRing = 512;           // pair-kernel launches that may be in flight at once across all streams of an engine
constexpr uint32_t kSchedStride = 32;          // ints per block (kSchedInts = 17, padded to a 128-byte line)
int* next_sched_block(MoeEngine* e) { return e->d_sched + (e->sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedRing) * kSchedStride; }

int check_launch(MoeEngine* e, const char* what) {
  cudaError_t err = cudaPeekAtLastError();
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(MOE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(err));
  }
  e->launches.fetch_add(1, std::memory_order_relaxed);
  return MOE_OK;
}

// launch a CTA-pair kernel, optionally with programmatic stream serialization (MOE_B200_PDL=1): its prologue (barrier init, TMEM
// allocation, weight loads) may then overlap the tail of the previous kernel on the stream; the kernel orders itself with
// griddepcontrol.wait (ptx.cuh).  OFF by default: measured gain 0.0-0.5 ms per 4K frame on one GPU, and with two streams
// sharing the engine one run in four of test_two_streams_share_an_engine came back with a wrong row segment
// (profiles/r02_pdl_experiment.txt) — not understood, so not shipped.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args)
{
  static const bool no_pdl = [] { const char* v = getenv("MOE_B200_PDL"); return !(v && v[0] == '1'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

int grid_for(int64_t threads, int block, int sm_count) {
  int64_t blocks = (threads + block - 1) / block;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(blocks, static_cast<int64_t>(sm_count) * 16)));
}

// Row segmentation of a persistent kernel: `base_items` strips (x planes x chunk groups) are each cut into nseg
// row segments and dealt round-robin to `workers` CTAs (or CTA pairs).  The kernel's time is waves x seg_rows
// (+ ~2 rows of pipeline fill per item), so nseg is chosen to make the item count land just below a multiple of
// `workers`: e.g. the a4 trunk conv (12 strip pairs, 74 pairs, 2160 rows) ran 4.05 waves of 87 rows = 5 x 87 row times
// with the old "4 items per worker" rule and runs 6 x 59 with this one.
void choose_segments(int64_t base_items, int workers, int H, int min_rows, int* seg_rows_out, int* nseg_out)
{
  int64_t best_cost = -1;
  int best_rows = H, best_n = 1;
  const int max_nseg = std::max(1, H / std::max(1, min_rows));
  for (int nseg = 1; nseg <= max_nseg; ++nseg) {
    const int rows = (H + nseg - 1) / nseg;
    const int n = (H + rows - 1) / rows;
    const int64_t items = base_items * n;
    const int64_t waves = (items + workers - 1) / workers;
    const int64_t cost = waves * (rows + 3);             // +3: two halo rows of pipeline fill and scheduling slack per item
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; best_n = n; }
    if (items > 64ll * workers) break;
  }
  *seg_rows_out = best_rows;
  *nseg_out = best_n;
}

// ---- one 3x3 convolution 64 -> 64*r*r ----------------------------------------------------------
// bias: device vector (single-CTA and SIMT kernels); bias_h: the same values on the host (CTA-pair kernels, as a kernel parameter)
int launch_conv(MoeEngine* e, cudaStream_t st, const __half* in, __half* out, const __half* skip,
                const uint8_t* w_img, const float* bias, const ConvBias* bias_h, int feat, int N, int H, int W, int r, int epi, float param, int center_only = 0)
{
  ConvParams p{};
  p.w_img = w_img; p.bias = bias; p.in = in; p.out = out; p.skip = skip;
  p.N = N; p.H = H; p.W = W; p.r = r; p.epi = epi; p.param = param; p.center_only = center_only;
  p.dynamic = !e->static_sched; p.sched = next_sched_block(e); p.dbg = e->dbg; p.ksteps = (feat + 15) / 16; p.bias_fused = e->bias_fused;
  // algorithmic FLOPs: 2 * taps * Cin * Cout per input pixel (padded channels are not counted)
  Timed timed(e, st, r == 1 ? 1 : 3, 2.0 * (center_only ? 1 : 9) * feat * (static_cast<double>(feat) * r * r) * N * H * W);
  if (e->simt) {
    const int64_t threads = static_cast<int64_t>(N) * H * W * r * r * 8;
    conv3x3_simt_kernel<<<grid_for(threads, 256, e->sm_count), 256, 0, st>>>(p);
    return check_launch(e, "conv3x3_simt_kernel");
  }
  const int groups = pair_groups(r);                        // chunk groups of the N = 128 pair kernel (2 for PixelShuffle(2), 5 for (3))
  const bool pair_path = r >= 2 && epi == EPI_BIAS_PRELU && !e->no_pair && e->sm_count / 2 >= groups;
  const int ncg1 = r * r;
  const bool pair_trunk = r == 1 && epi != EPI_BIAS_PRELU && !e->no_pair && !e->no_pair_trunk && e->sm_count / 2 >= ncg1;
  if (pair_trunk) {
    // CTA pairs, 256 px x 64 channels per MMA (conv_pair.cuh): the 64 -> 64 convolutions, and Net3x's nine sub-pixel chunks
    const int npairs = e->sm_count / 2 / ncg1 * ncg1;       // a multiple of r*r: a pair keeps its chunk
    const int strips1 = (W + kStripW - 1) / kStripW;
    p.strips = (strips1 + 1) / 2;                            // strip PAIRS
    const int64_t base_items = static_cast<int64_t>(N) * p.strips * ncg1;
    choose_segments(base_items, npairs, H, 8, &p.seg_rows, &p.nseg);
    const int64_t items = base_items * p.nseg;
    if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "conv problem too large");
    p.items = static_cast<int>(items);
  }
  if (pair_path) {
    // CTA pairs (cta_group::2): 256 px x 128 channels per MMA, conv_pair.cuh
    const int npairs = e->sm_count / 2 / groups * groups;   // a multiple of the group count, so a pair keeps its chunk group (weights stay resident)
    const int strips1 = (W + kStripW - 1) / kStripW;
    p.strips = (strips1 + 1) / 2;                            // strip PAIRS
    const int64_t base_items = static_cast<int64_t>(groups) * N * p.strips;
    choose_segments(base_items, npairs, H, 8, &p.seg_rows, &p.nseg);
    const int64_t items = base_items * p.nseg;
    if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "conv problem too large");
    p.items = static_cast<int>(items);
  }
  const int ncg = r * r;                                    // one 64-channel chunk per CTA
  const int G = std::max(ncg, e->sm_count / ncg * ncg);    // CTAs; multiple of ncg so a CTA keeps its chunk
  int grid = 0;
  if (!pair_path && !pair_trunk) {
    p.strips = (W + kStripW - 1) / kStripW;
    const int64_t base_items = static_cast<int64_t>(N) * p.strips * ncg;
    choose_segments(base_items, G, H, 8, &p.seg_rows, &p.nseg);
    const int64_t items = base_items * p.nseg;
    if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "conv problem too large");
    p.items = static_cast<int>(items);
    grid = static_cast<int>(std::min<int64_t>(G, items));   // items is a multiple of ncg
  }

  // tensor maps: input rows (130-px box), residual rows and one output view per PixelShuffle sub-pixel
  ConvMaps maps;
  memset(&maps, 0, sizeof maps);
  auto encode = [&](CUtensorMap* tm, const void* ptr, cuuint64_t w, cuuint64_t h, cuuint64_t n, cuuint64_t pix_stride,
                    cuuint64_t row_stride, cuuint64_t plane_stride, cuuint32_t box_w) -> CUresult {
    const cuuint64_t dims[4] = {64, w, h, n};
    const cuuint64_t strides[3] = {pix_stride, row_stride, plane_stride};
    const cuuint32_t box[4] = {64, box_w, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return e->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  const cuuint64_t uW = W, uH = H, uN = N, Wo = static_cast<cuuint64_t>(W) * r, Ho = static_cast<cuuint64_t>(H) * r;
  CUresult cr = encode(&maps.in, in, uW, uH, uN, 128, uW * 128, uW * uH * 128, kRowPx);
  if (cr == CUDA_SUCCESS && epi == EPI_SCALE_SKIP) cr = encode(&maps.skip, skip, uW, uH, uN, 128, uW * 128, uW * uH * 128, kStripW);
  for (int q = 0; q < r * r && cr == CUDA_SUCCESS; ++q) {
    const int sy = q / r, sx = q % r;
    cr = encode(&maps.out[q], out + (static_cast<size_t>(sy) * Wo + sx) * 64, uW, uH, uN, static_cast<cuuint64_t>(r) * 128,
                static_cast<cuuint64_t>(r) * Wo * 128, Ho * Wo * 128, kStripW);
  }
  if (cr != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for N=%d H=%d W=%d r=%d", (int)cr, N, H, W, r);
  if (pair_trunk) {
    // one instantiation per epilogue: the epilogue arithmetic is straight-line code for the 32 accumulator values a thread owns
    typedef void (*TrunkFn)(const ConvMaps, const ConvParams);
    static const TrunkFn trunk_fn[4] = {conv3x3_pair_trunk_kernel<EPI_PLAIN>, conv3x3_pair_trunk_kernel<EPI_PRELU>,
                                        conv3x3_pair_trunk_kernel<EPI_SCALE_SKIP>, conv3x3_pair_trunk_kernel<EPI_BIAS_PRELU>};
    if (!e->pair_trunk_attr_set) {
      for (int i = 0; i < 4; ++i) MOE_CUDA(cudaFuncSetAttribute(trunk_fn[i], cudaFuncAttributeMaxDynamicSharedMemorySize, PairTrunkCfg::kSmemBytes));
      e->pair_trunk_attr_set = true;
    }
    const int npairs = static_cast<int>(std::min<int64_t>(e->sm_count / 2 / ncg1 * ncg1, p.items));   // p.items is a multiple of r*r
    MOE_CUDA(launch_pdl(trunk_fn[epi], 2 * npairs, kConvThreads, PairTrunkCfg::kSmemBytes, st, maps, p));
    return check_launch(e, "conv3x3_pair_trunk_kernel");
  }
  if (pair_path) {
    if (!e->pair_attr_set) {
      MOE_CUDA(cudaFuncSetAttribute(conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
      e->pair_attr_set = true;
    }
    const int npairs = static_cast<int>(std::min<int64_t>(e->sm_count / 2 / groups * groups, (p.items + groups - 1) / groups * groups));
    if (!bias_h) return fail(MOE_ERR_INVALID, "the CTA-pair convolution needs the host copy of its bias");
    MOE_CUDA(launch_pdl(conv3x3_pair_kernel, 2 * npairs, kConvThreads, PairCfg::kSmemBytes, st, maps, p, *bias_h));
    return check_launch(e, "conv3x3_pair_kernel");
  }
  typedef void (*ConvFn)(const ConvMaps, const ConvParams);
  static const ConvFn conv_fn[4] = {conv3x3_tc_kernel<EPI_PLAIN>, conv3x3_tc_kernel<EPI_PRELU>, conv3x3_tc_kernel<EPI_SCALE_SKIP>,
                                    conv3x3_tc_kernel<EPI_BIAS_PRELU>};
  if (!e->smem_attr_set) {
    for (int i = 0; i < 4; ++i) MOE_CUDA(cudaFuncSetAttribute(conv_fn[i], cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg::kSmemBytes));
    e->smem_attr_set = true;
  }
  conv_fn[epi]<<<grid, kConvThreads, ConvCfg::kSmemBytes, st>>>(maps, p);
  return check_launch(e, "conv3x3_tc_kernel");
}

// ---- one residual block t' = t + scale * conv_2(PReLU(conv_1(t))) as one kernel (conv_arsb.cuh) ----
int launch_arsb(MoeEngine* e, cudaStream_t st, const __half* in, __half* out, const uint8_t* w1_img, const uint8_t* w2_img,
                int feat, int N, int H, int W, float slope, float scale)
{
  ArsbParams ap{};
  ConvParams& p = ap.c;
  p.w_img = w1_img; p.in = in; p.out = out; p.N = N; p.H = H; p.W = W; p.r = 1; p.epi = EPI_PRELU; p.param = slope;
  p.dynamic = !e->static_sched; p.sched = next_sched_block(e); p.dbg = e->dbg; p.ksteps = (feat + 15) / 16;
  ap.w2_img = w2_img; ap.scale = scale;
  Timed timed(e, st, 4, 2 * 2.0 * 9 * feat * static_cast<double>(feat) * N * H * W);   // both convolutions
  if (e->arsb_solo) {
    // in-kernel cycle accounting of the MMA warps: 8 CTAs x 12 values behind the per-pair records of the debug buffer (tools/arsb_waits.py)
    if (const char* x = getenv("MOE_ARSB_EXP")) p.center_only = atoi(x) != 0 && e->dbg_bytes >= (74 * 4 + 96) * sizeof(unsigned long long);
    // one CTA per SM, full weights of both convolutions per SM (conv_arsb_solo.cuh)
    p.strips = (W + kArsbStripW - 1) / kArsbStripW;
    const int64_t base_items = static_cast<int64_t>(N) * p.strips;
    choose_segments(base_items, e->sm_count, H, 16, &p.seg_rows, &p.nseg);
    const int64_t items = base_items * p.nseg;
    if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "conv problem too large");
    p.items = static_cast<int>(items);
    ArsbMaps maps;
    memset(&maps, 0, sizeof maps);
    const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {128, dims[1] * 128, dims[1] * dims[2] * 128};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box_in[4] = {64, kRowPx, 1, 1};
    const CUresult cr = e->encode(&maps.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(in), dims, strides, box_in, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the residual block", (int)cr);
    if (!e->arsb_solo_attr_set) {
      MOE_CUDA(cudaFuncSetAttribute(arsb_solo_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ArsbSoloCfg::kSmemBytes));
      MOE_CUDA(cudaFuncSetAttribute(arsb_solo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ArsbSoloCfg::kSmemBytes));
      e->arsb_solo_attr_set = true;
    }
    const int grid = static_cast<int>(std::min<int64_t>(e->sm_count, p.items));
    MOE_CUDA(launch_pdl(p.ksteps == 3 ? arsb_solo_kernel<3> : arsb_solo_kernel<4>, grid, kArsbSoloThreads, ArsbSoloCfg::kSmemBytes, st, maps, ap));
    return check_launch(e, "arsb_solo_kernel");
  }
  const int npairs_max = e->sm_count / 2;
  const int strips1 = (W + kArsbStripW - 1) / kArsbStripW;
  p.strips = (strips1 + 1) / 2;                                // strip PAIRS of 2 x 126 px
  const int64_t base_items = static_cast<int64_t>(N) * p.strips;
  choose_segments(base_items, npairs_max, H, 16, &p.seg_rows, &p.nseg);
  const int64_t items = base_items * p.nseg;
  if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "conv problem too large");
  p.items = static_cast<int>(items);
  ArsbMaps maps;
  memset(&maps, 0, sizeof maps);
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {128, dims[1] * 128, dims[1] * dims[2] * 128};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t box_in[4] = {64, kRowPx, 1, 1}, box_out[4] = {64, kArsbStripW, 1, 1};
  CUresult cr = e->encode(&maps.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(in), dims, strides, box_in, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr == CUDA_SUCCESS)
    cr = e->encode(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, out, dims, strides, box_out, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the residual block", (int)cr);
  if (!e->arsb_attr_set) {
    MOE_CUDA(cudaFuncSetAttribute(arsb_pair_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ArsbCfgT<false>::kSmemBytes));
    MOE_CUDA(cudaFuncSetAttribute(arsb_pair_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ArsbCfgT<false>::kSmemBytes));
    MOE_CUDA(cudaFuncSetAttribute(arsb_pair_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ArsbCfgT<true>::kSmemBytes));
    MOE_CUDA(cudaFuncSetAttribute(arsb_pair_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ArsbCfgT<true>::kSmemBytes));
    e->arsb_attr_set = true;
  }
  const int npairs = static_cast<int>(std::min<int64_t>(npairs_max, p.items));
  if (e->arsb_smem_mid)
    MOE_CUDA(launch_pdl(p.ksteps == 3 ? arsb_pair_kernel<false, 3> : arsb_pair_kernel<false, 4>, 2 * npairs, kConvThreads, ArsbCfgT<false>::kSmemBytes, st, maps, ap));
  else
    MOE_CUDA(launch_pdl(p.ksteps == 3 ? arsb_pair_kernel<true, 3> : arsb_pair_kernel<true, 4>, 2 * npairs, kConvThreads, ArsbCfgT<true>::kSmemBytes, st, maps, ap));
  return check_launch(e, "arsb_pair_kernel");
}

// ---- last upsample conv of a branch fused with the head's dot products (conv_pair_head.cuh) ----
int launch_conv_head(MoeEngine* e, cudaStream_t st, const __half* in, const uint8_t* w_img, const float* bias, const ConvBias* bias_h, int feat, int N, int H, int W,
                     int r, float slope, const uint8_t* head_img, float* hbuf, float* ebuf, int center_only = 0)
{
  PairHeadParams hp{};
  ConvParams& p = hp.c;
  p.w_img = w_img; p.bias = bias; p.in = in; p.out = nullptr; p.skip = nullptr;
  p.N = N; p.H = H; p.W = W; p.r = r; p.epi = EPI_BIAS_PRELU; p.param = slope; p.center_only = center_only;
  p.dynamic = !e->static_sched; p.sched = next_sched_block(e); p.dbg = e->dbg; p.ksteps = (feat + 15) / 16; p.bias_fused = e->bias_fused;
  hp.head_img = head_img; hp.hbuf = hbuf; hp.ebuf = ebuf;
  Timed timed(e, st, 5, 2.0 * (center_only ? 1 : 9) * feat * (static_cast<double>(feat) * r * r) * N * H * W);
  const int groups = pair_groups(r);                          // 2 for PixelShuffle(2), 5 for PixelShuffle(3)
  const int npairs_max = e->sm_count / 2 / groups * groups;   // a multiple of the group count: a pair keeps its chunk group
  const int strips1 = (W + kStripW - 1) / kStripW;
  p.strips = (strips1 + 1) / 2;
  const int64_t base_items = static_cast<int64_t>(groups) * N * p.strips;
  choose_segments(base_items, npairs_max, H, 8, &p.seg_rows, &p.nseg);
  const int64_t items = base_items * p.nseg;
  if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "conv problem too large");
  p.items = static_cast<int>(items);
  ConvMaps maps;
  memset(&maps, 0, sizeof maps);
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {128, dims[1] * 128, dims[1] * dims[2] * 128};
  const cuuint32_t box[4] = {64, kRowPx, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = e->encode(&maps.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(in), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
  if (!e->pair_head_attr_set) {
    MOE_CUDA(cudaFuncSetAttribute(conv3x3_pair_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PairHeadCfg::kSmemBytes));
    e->pair_head_attr_set = true;
  }
  const int npairs = static_cast<int>(std::min<int64_t>(npairs_max, (p.items + groups - 1) / groups * groups));
  MOE_CUDA(launch_pdl(conv3x3_pair_head_kernel, 2 * npairs, kPairHeadThreads, PairHeadCfg::kSmemBytes, st, maps, hp, *bias_h));
  return check_launch(e, "conv3x3_pair_head_kernel");
}

// ---- the two heads + blend + store on the tensor cores (head_tc.cuh) ---------------------------
int launch_head_tc(MoeEngine* e, cudaStream_t st, const MoeModel* m, const HeadParams& hp)
{
  HeadTcParams p{};
  p.g = hp;
  p.w_img = m->d_head_img;
  p.strips = (hp.W + kHeadStripOut - 1) / kHeadStripOut;
  const int G = e->sm_count;
  const int64_t base_items = static_cast<int64_t>(hp.N) * p.strips;
  choose_segments(base_items, G, hp.H, 16, &p.seg_rows, &p.nseg);
  const int64_t items = base_items * p.nseg;
  if (items > 0x7fffffff) return fail(MOE_ERR_INVALID, "head problem too large");
  p.items = static_cast<int>(items);
  HeadMaps maps;
  memset(&maps, 0, sizeof maps);
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(hp.W), static_cast<cuuint64_t>(hp.H), static_cast<cuuint64_t>(hp.N)};
  const cuuint64_t strides[3] = {128, dims[1] * 128, dims[1] * dims[2] * 128};
  const cuuint32_t box[4] = {64, 128, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int b = 0; b < 2; ++b) {
    CUresult cr = e->encode(b ? &maps.r : &maps.u, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(b ? hp.r : hp.u), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(MOE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the head input", (int)cr);
  }
  if (!e->head_attr_set) {
    MOE_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmemBytes));
    e->head_attr_set = true;
  }
  {
    Timed timed(e, st, 2, static_cast<double>(hp.N) * hp.H * hp.W * (2 * 128 + 2));   // bytes: two 64-ch fp16 reads, one fp16 write
    head_tc_kernel<<<static_cast<int>(std::min<int64_t>(G, items)), kHeadThreads, kHeadSmemBytes, st>>>(maps, p);
  }
  return MOE_OK;
}

// ---- geometry of one tile restricted to a canvas row window ------------------------------------
struct TileGeom {
  bool active;
  int c0, c1;            // computed LR rows, tile-relative
  int H, W;              // computed LR rectangle
  int keep_y0, keep_y1, keep_x0, keep_x1, ramp_y0, ramp_x0, blend_y1, blend_x1;
};

TileGeom tile_geom(const MoePlan& pl, const MoeTile& t, int row_lo, int row_hi) {
  TileGeom g{};
  const int sc = pl.scale, psc = pl.pad_sc;
  const int th = t.bottom - t.top;
  const int oy = t.top * sc, ox = t.left * sc;
  const int lh = t.bsc - oy, lw = t.rsc - ox;          // tile extent on the canvas after unpad
  int lt_r = t.top_t < 0 ? lh + t.top_t : t.top_t;     // blend(): negative anchors count from the end
  int lt_c = t.left_t < 0 ? lw + t.left_t : t.left_t;
  const int start_r = lt_r < 1 ? 0 : lt_r - psc;
  const int start_c = lt_c < 1 ? 0 : lt_c - psc;
  if (lt_r < 1) lt_r = 0;
  if (lt_c < 1) lt_c = 0;
  g.ramp_y0 = oy + start_r; g.blend_y1 = oy + lt_r;
  g.ramp_x0 = ox + start_c; g.blend_x1 = ox + lt_c;
  if (lt_r == 0) g.blend_y1 = g.ramp_y0;
  if (lt_c == 0) g.blend_x1 = g.ramp_x0;
  g.keep_y0 = std::max(g.ramp_y0, row_lo); g.keep_y1 = std::min(t.bsc, row_hi);
  g.keep_x0 = g.ramp_x0; g.keep_x1 = t.rsc;
  g.active = g.keep_y0 < g.keep_y1 && g.keep_x0 < g.keep_x1;
  if (!g.active) return g;
  const int lo = (g.keep_y0 - oy) / sc;
  const int hi = (g.keep_y1 - oy + sc - 1) / sc;
  g.c0 = std::max(0, lo - kHaloLR);
  g.c1 = std::min(th, hi + kHaloLR);
  g.H = g.c1 - g.c0;
  g.W = t.right - t.left;
  return g;
}

// workspace of one tile in units of (planes x h x w x 128 B): A, T, M, C, then for PixelShuffle(2) nets the shared
// stage buffers S1 (4 units, if >= 2 stages) and S2 (16 units, if 3 stages) and the final region — two 4^n_up-unit
// tensors (unfused) or the P array (fused, smaller); PixelShuffle(3): two 9-unit tensors.
size_t stage_units(const MoeModel* m) { return (m->n_up >= 2 ? 4 : 0) + (m->n_up >= 3 ? 16 : 0); }
size_t tile_units(const MoeModel* m) {
  if (m->n_up == 0) return 4;
  if (m->r == 3) return 4 + 2 * 9;
  size_t fin = 2;
  for (int i = 0; i < m->n_up; ++i) fin *= 4;
  return 4 + stage_units(m) + fin;
}

// MoeNet_lite2's FRM reduction (partial sums per plane + gates) lives behind the tile's activation buffers in the CALLER's
// workspace, so two calls on two streams with two workspaces never share it
size_t frm_bytes(const MoeModel* m, int planes) {
  return m->arch == MOE_ARCH_LITE ? align_up((static_cast<size_t>(kFrmBlocks) + 1) * 64 * sizeof(float) * planes, 1024) : 0;
}

int validate_plan(const MoeModel* m, const MoePlan* pl, int planes, int row_lo, int row_hi) {
  if (!m || !pl || !pl->tiles || pl->n_tiles <= 0) return fail(MOE_ERR_INVALID, "null model or empty plan");
  if (pl->scale != m->scale) return fail(MOE_ERR_INVALID, "plan scale %d does not match model scale %d", pl->scale, m->scale);
  if (planes <= 0 || pl->in_h <= 0 || pl->in_w <= 0) return fail(MOE_ERR_INVALID, "bad image shape");
  if (pl->out_h != pl->in_h * pl->scale || pl->out_w != pl->in_w * pl->scale) return fail(MOE_ERR_INVALID, "canvas size mismatch");
  if (pl->pad_sc > 0 && !pl->ramp) return fail(MOE_ERR_INVALID, "plan has a seam but no ramp");
  if (pl->pad_sc < 0 || pl->pad_sc > kMaxSeam) return fail(MOE_ERR_INVALID, "seam of %d px is wider than %d", pl->pad_sc, kMaxSeam);
  if (row_lo < 0 || row_hi > pl->out_h || row_lo >= row_hi) return fail(MOE_ERR_INVALID, "bad row window [%d,%d)", row_lo, row_hi);
  for (int i = 0; i < pl->n_tiles; ++i) {
    const MoeTile& t = pl->tiles[i];
    if (t.top < 0 || t.left < 0 || t.bottom <= t.top || t.right <= t.left || t.bottom > pl->in_h + pl->pad_h ||
        t.right > pl->in_w + pl->pad_w || t.bsc > pl->out_h || t.rsc > pl->out_w || t.bsc <= t.top * pl->scale ||
        t.rsc <= t.left * pl->scale)
      return fail(MOE_ERR_INVALID, "tile %d is outside the padded image", i);
  }
  return MOE_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int moe_abi_version(void) { return MOE_ABI_VERSION; }
const char* moe_last_error(void) { return g_err.c_str(); }

int moe_engine_create(int device_id, MoeEngine** out)
{
  if (!out) return fail(MOE_ERR_INVALID, "out is null");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return fail(MOE_ERR_NO_DEVICE, "no CUDA device visible (this engine has no CPU fallback)");
  }
  if (device_id < 0 || device_id >= count) return fail(MOE_ERR_INVALID, "device %d out of range (%d devices)", device_id, count);
  cudaDeviceProp prop;
  MOE_CUDA(cudaGetDeviceProperties(&prop, device_id));
  if (prop.major != 10) return fail(MOE_ERR_NO_DEVICE, "device %d is sm_%d%d; this build is sm_100a only", device_id, prop.major, prop.minor);
  MoeEngine* e = new MoeEngine();
  e->device = device_id;
  e->sm_count = prop.multiProcessorCount;
  Guard g(device_id);
  if (!g.ok) { delete e; return fail(MOE_ERR_CUDA, "cudaSetDevice(%d) failed", device_id); }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    delete e;
    return fail(MOE_ERR_CUDA, "driver has no cuTensorMapEncodeTiled");
  }
  e->encode = reinterpret_cast<EncodeTiledFn>(fn);
  static_assert(kSchedInts <= kSchedStride, "scheduler block too small");
  if (cudaMalloc(&e->d_sched, kSchedRing * kSchedStride * sizeof(int)) != cudaSuccess ||
      cudaMemset(e->d_sched, 0, kSchedRing * kSchedStride * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    if (e->d_sched) cudaFree(e->d_sched);
    delete e;
    return fail(MOE_ERR_NOMEM, "scheduler counter allocation failed");
  }
  const char* env = getenv("MOE_B200_SIMT");
  e->simt = env && env[0] == '1';
  *out = e;
  return MOE_OK;
}

void moe_engine_destroy(MoeEngine* e)
{
  if (!e) return;
  Guard g(e->device);
  for (int i = 0; i < kNumBufs; ++i) if (e->buf[i]) cudaFree(e->buf[i]);
  if (e->d_sched) cudaFree(e->d_sched);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->copy_event) cudaEventDestroy(e->copy_event);
  delete e;
}

int64_t moe_engine_launch_count(const MoeEngine* e) { return e ? e->launches.load() : 0; }

int moe_engine_profile(MoeEngine* e, int enable)
{
  if (!e) return fail(MOE_ERR_INVALID, "engine is null");
  e->profiling = enable != 0;
  return MOE_OK;
}

int moe_engine_profile_read(MoeEngine* e, double ms[MOE_PROFILE_CLASSES], double work[MOE_PROFILE_CLASSES], int64_t launches[MOE_PROFILE_CLASSES])
{
  if (!e || !ms || !work || !launches) return fail(MOE_ERR_INVALID, "null argument");
  Guard g(e->device);
  for (auto& s : e->spans) {
    float t = 0.f;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
      e->prof_ms[s.cls] += t; e->prof_work[s.cls] += s.work; e->prof_n[s.cls] += 1;
    }
    cudaEventDestroy(s.a); cudaEventDestroy(s.b);
  }
  e->spans.clear();
  for (int i = 0; i < MOE_PROFILE_CLASSES; ++i) {
    ms[i] = e->prof_ms[i]; work[i] = e->prof_work[i]; launches[i] = e->prof_n[i];
    e->prof_ms[i] = e->prof_work[i] = 0; e->prof_n[i] = 0;
  }
  return MOE_OK;
}

int moe_engine_set_conv_path(MoeEngine* e, int simt)
{
  if (!e) return fail(MOE_ERR_INVALID, "engine is null");
  e->simt = simt & 1;
  e->no_pair = (simt >> 1) & 1;
  e->no_pair_trunk = (simt >> 2) & 1;
  e->no_fuse = (simt >> 3) & 1;
  e->static_sched = (simt >> 4) & 1;
  e->bias_fused = (simt >> 5) & 1;
  e->no_arsb = (simt >> 6) & 1;
  e->arsb_smem_mid = (simt >> 7) & 1;
  e->arsb_solo = (simt >> 8) & 1;
  e->full_k = (simt >> 9) & 1;
  return MOE_OK;
}

int moe_engine_check(MoeEngine* e, void* stream)
{
  if (!e) return fail(MOE_ERR_INVALID, "engine is null");
  Guard g(e->device);
  MOE_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  unsigned int flag = 0;
  MOE_CUDA(cudaMemcpyFromSymbol(&flag, ptx::g_mbar_abort, sizeof flag));
  if (!flag) return MOE_OK;
  flag = 0;
  MOE_CUDA(cudaMemcpyToSymbol(ptx::g_mbar_abort, &flag, sizeof flag));
  // the work-item counters of the aborted launches were left mid-count: start every block from zero again
  MOE_CUDA(cudaMemset(e->d_sched, 0, kSchedRing * kSchedStride * sizeof(int)));
  return fail(MOE_ERR_CUDA, "a kernel gave up waiting on an mbarrier (pipeline protocol error or a stalled GPU): the results since the last check are invalid; the engine remains usable");
}

int moe_engine_debug_timeout(MoeEngine* e, uint64_t ns)
{
  if (!e) return fail(MOE_ERR_INVALID, "engine is null");
  Guard g(e->device);
  unsigned long long v = ns ? ns : 4000000000ull;
  MOE_CUDA(cudaMemcpyToSymbol(ptx::g_mbar_timeout_ns, &v, sizeof v));
  return MOE_OK;
}

int moe_engine_debug_buffer(MoeEngine* e, void* dev, size_t nbytes)
{
  if (!e) return fail(MOE_ERR_INVALID, "engine is null");
  if (dev && nbytes < static_cast<size_t>(e->sm_count / 2) * 4 * sizeof(unsigned long long))
    return fail(MOE_ERR_INVALID, "debug buffer needs %zu bytes", static_cast<size_t>(e->sm_count / 2) * 4 * sizeof(unsigned long long));
  e->dbg = static_cast<unsigned long long*>(dev);
  e->dbg_bytes = dev ? nbytes : 0;
  return MOE_OK;
}

int moe_model_load(MoeEngine* e, int arch, const void* blob, size_t nbytes, MoeModel** out)
{
  if (!e || !blob || !out) return fail(MOE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (nbytes < sizeof(BlobHeader)) return fail(MOE_ERR_INVALID, "blob too small");
  BlobHeader h;
  memcpy(&h, blob, sizeof h);
  if (h.magic != kBlobMagic || h.version != kBlobVersion) return fail(MOE_ERR_INVALID, "not a moephoto_b200 weight blob");
  if (static_cast<int>(h.arch) != arch) return fail(MOE_ERR_INVALID, "blob is arch %u, asked for %d", h.arch, arch);
  if (h.n_up > 3 || (h.n_up && h.r != 2 && h.r != 3) || (h.n_up >= 2 && h.r != 2)) return fail(MOE_ERR_INVALID, "unsupported upsample layout");
  if (arch != MOE_ARCH_NET2X && arch != MOE_ARCH_NET3X && arch != MOE_ARCH_NET4X && arch != MOE_ARCH_NETDN && arch != MOE_ARCH_LITE)
    return fail(MOE_ERR_INVALID, "unknown architecture %d", arch);
  if (h.n_sections > 256 || nbytes < sizeof h + h.n_sections * sizeof(BlobEntry)) return fail(MOE_ERR_INVALID, "truncated directory");
  {
    // the header must describe the architecture it names (the table of moephoto_b200/weights.py::pack)
    static const struct { int arch; uint32_t feat, n_up, r; } kLayouts[] = {
      {MOE_ARCH_NETDN, 48, 0, 0}, {MOE_ARCH_NET2X, 64, 1, 2}, {MOE_ARCH_NET3X, 64, 1, 3}, {MOE_ARCH_NET4X, 64, 2, 2}};
    bool known = arch == MOE_ARCH_LITE && h.feat == 48 && h.r == 2 && h.n_up >= 1 && h.n_up <= 3;
    for (const auto& l : kLayouts) known = known || (l.arch == arch && l.feat == h.feat && l.n_up == h.n_up && l.r == h.r);
    if (!known) return fail(MOE_ERR_INVALID, "header (feat %u, n_up %u, r %u) does not describe architecture %d", h.feat, h.n_up, h.r, arch);
  }
  Guard g(e->device);
  if (!g.ok) return fail(MOE_ERR_CUDA, "cudaSetDevice failed");
  MoeModel* m = new MoeModel();
  m->e = e; m->arch = arch; m->feat = h.feat; m->n_up = h.n_up; m->r = h.r;
  m->scale = 1;
  for (uint32_t i = 0; i < h.n_up; ++i) m->scale *= static_cast<int>(h.r);
  // every failure exit below goes through moe_model_free: nothing allocated so far is leaked
  if (cudaMalloc(&m->d_blob, nbytes) != cudaSuccess) { cudaGetLastError(); moe_model_free(m); return fail(MOE_ERR_NOMEM, "cudaMalloc(%zu) for weights failed", nbytes); }
  if (cudaMemcpy(m->d_blob, blob, nbytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); moe_model_free(m); return fail(MOE_ERR_CUDA, "weight upload failed"); }
  const uint8_t* hb = static_cast<const uint8_t*>(blob);
  bool have_scalars = false;
  int n_trunk = 0, n_head = 0, n_up_img = 0, n_up_bias = 0, n_frm = 0;
  const size_t n_chunks = h.n_up ? static_cast<size_t>(h.r) * h.r : 0;
  for (uint32_t i = 0; i < h.n_sections; ++i) {
    BlobEntry en;
    memcpy(&en, hb + sizeof h + i * sizeof en, sizeof en);
    bool bad = en.offset % 256 || en.nbytes > nbytes || en.offset > nbytes - en.nbytes;       // no wrap-around for a crafted blob
    const uint8_t* dp = m->d_blob + en.offset;
    switch (en.kind) {
      case SEC_FIRST_W: bad |= en.nbytes != 9 * 64 * 4; m->first_w = reinterpret_cast<const float*>(dp); break;
      case SEC_SCALARS: bad |= en.nbytes != sizeof m->scalars; if (!bad) { memcpy(m->scalars, hb + en.offset, sizeof m->scalars); have_scalars = true; } break;
      case SEC_TRUNK_IMG: bad |= en.index >= 13 || en.nbytes != kChunkImgBytes; if (!bad) { m->trunk_img[en.index] = dp; ++n_trunk; } break;
      case SEC_UP_IMG: bad |= en.index >= 8 || en.nbytes != n_chunks * kChunkImgBytes; if (!bad) { m->up_img[en.index] = dp; ++n_up_img; } break;
      case SEC_UP_BIAS: bad |= en.index >= 8 || en.nbytes != n_chunks * 64 * 4; if (!bad) { m->up_bias[en.index] = reinterpret_cast<const float*>(dp); memcpy(m->up_bias_h[en.index].v, hb + en.offset, en.nbytes); ++n_up_bias; } break;
      case SEC_FRM: bad |= en.index >= 3 || en.nbytes != 516 * 4; if (!bad) { m->frm[en.index] = reinterpret_cast<const float*>(dp); ++n_frm; } break;
      case SEC_HEAD_W: bad |= en.index >= 2 || en.nbytes != 9 * 64 * 4; if (!bad) { m->head_w[en.index] = reinterpret_cast<const float*>(dp); ++n_head; } break;
      default: bad = true;
    }
    if (bad) { moe_model_free(m); return fail(MOE_ERR_INVALID, "bad blob section %u (kind %u index %u)", i, en.kind, en.index); }
  }
  const int want_up = 2 * static_cast<int>(h.n_up);
  const bool lite = arch == MOE_ARCH_LITE;
  bool complete = m->first_w && have_scalars && n_trunk == (lite ? 7 : 13) && n_head == 2 && n_up_img == want_up && n_up_bias == want_up &&
                  n_frm == (lite ? 3 : 0);
  for (int l = 0; l < (lite ? 7 : 13) && complete; ++l) complete = m->trunk_img[l] != nullptr;
  for (int b = 0; b < 2 && complete; ++b)
    for (uint32_t s = 0; s < h.n_up && complete; ++s) complete = m->up_img[4 * b + s] && m->up_bias[4 * b + s];
  if (!complete) { moe_model_free(m); return fail(MOE_ERR_INVALID, "blob is missing sections"); }
  {
    // head filters as a K-major SWIZZLE_128B B operand: row = tap (9 of 16 used), 64 input channels
    std::vector<__half> img(2 * 16 * 64, __float2half(0.f));
    for (int b = 0; b < 2; ++b) {
      const float* hw = reinterpret_cast<const float*>(hb + (reinterpret_cast<const uint8_t*>(m->head_w[b]) - m->d_blob));
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < 64; ++c)
          img[(b * 16 + t) * 64 + (((c >> 3) ^ (t & 7)) << 3) + (c & 7)] = __float2half(hw[t * 64 + c]);
    }
    if (cudaMalloc(&m->d_head_img, 4096) != cudaSuccess ||
        cudaMemcpy(m->d_head_img, img.data(), 4096, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError(); moe_model_free(m);
      return fail(MOE_ERR_NOMEM, "head weight upload failed");
    }
  }
  *out = m;
  return MOE_OK;
}

void moe_model_free(MoeModel* m)
{
  if (!m) return;
  Guard g(m->e->device);
  if (m->d_blob) cudaFree(m->d_blob);
  if (m->d_head_img) cudaFree(m->d_head_img);
  delete m;
}

int moe_model_scale(const MoeModel* m) { return m ? m->scale : 0; }

size_t moe_plan_workspace_bytes(const MoeModel* m, int planes, const MoePlan* plan, int row_lo, int row_hi)
{
  if (validate_plan(m, plan, planes, row_lo, row_hi) != MOE_OK) return 0;
  size_t worst = 0;
  for (int i = 0; i < plan->n_tiles; ++i) {
    TileGeom g = tile_geom(*plan, plan->tiles[i], row_lo, row_hi);
    if (!g.active) continue;
    if (m->arch == MOE_ARCH_LITE) g.H = plan->tiles[i].bottom - plan->tiles[i].top;
    const size_t unit = align_up(static_cast<size_t>(planes) * g.H * g.W * 128, 1024);
    worst = std::max(worst, unit * tile_units(m) + frm_bytes(m, planes));
  }
  return worst + 1024;
}

}  // extern "C"

namespace {
// after_tile(ti, ctx): called on the host right after tile ti's kernels were enqueued (moe_enhance_host overlaps the
// device->host copy of finished canvas columns with the next tile's compute)
typedef int (*AfterTileFn)(int ti, void* ctx);
int run_plan_core(MoeModel* m, const void* in, int64_t in_plane_stride, int64_t in_row_stride, int planes,
                  void* canvas, int64_t out_plane_stride, int64_t out_row_stride,
                  const MoePlan* plan, int row_lo, int row_hi, void* workspace, size_t workspace_bytes, void* stream,
                  AfterTileFn after_tile, void* after_ctx);
}  // namespace

extern "C" {

int moe_run_plan(MoeModel* m, const void* in, int64_t in_plane_stride, int64_t in_row_stride, int planes,
                 void* canvas, int64_t out_plane_stride, int64_t out_row_stride,
                 const MoePlan* plan, int row_lo, int row_hi, void* workspace, size_t workspace_bytes, void* stream)
{
  return run_plan_core(m, in, in_plane_stride, in_row_stride, planes, canvas, out_plane_stride, out_row_stride, plan, row_lo, row_hi,
                       workspace, workspace_bytes, stream, nullptr, nullptr);
}

}  // extern "C"

namespace {
int run_plan_core(MoeModel* m, const void* in, int64_t in_plane_stride, int64_t in_row_stride, int planes,
                  void* canvas, int64_t out_plane_stride, int64_t out_row_stride,
                  const MoePlan* plan, int row_lo, int row_hi, void* workspace, size_t workspace_bytes, void* stream,
                  AfterTileFn after_tile, void* after_ctx)
{
  int rc = validate_plan(m, plan, planes, row_lo, row_hi);
  if (rc != MOE_OK) return rc;
  if (!in || !canvas || !workspace) return fail(MOE_ERR_INVALID, "null buffer");
  if (workspace_bytes < moe_plan_workspace_bytes(m, planes, plan, row_lo, row_hi))
    return fail(MOE_ERR_NOMEM, "workspace of %zu bytes is too small (need %zu)", workspace_bytes,
                moe_plan_workspace_bytes(m, planes, plan, row_lo, row_hi));
  MoeEngine* e = m->e;
  Guard guard(e->device);
  if (!guard.ok) return fail(MOE_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  const int sc = plan->scale, N = planes;

  for (int ti = 0; ti < plan->n_tiles; ++ti) {
    const MoeTile& t = plan->tiles[ti];
    TileGeom g = tile_geom(*plan, t, row_lo, row_hi);
    if (!g.active) continue;
    if (m->arch == MOE_ARCH_LITE) { g.c0 = 0; g.c1 = t.bottom - t.top; g.H = g.c1; }   // FRM averages over the WHOLE tile: no row window
    const int H = g.H, W = g.W;
    const size_t unit = align_up(static_cast<size_t>(N) * H * W * 128, 1024);
    __half* bufA = reinterpret_cast<__half*>(ws);              // `out`  = PReLU(conv_input(x))
    __half* bufT = reinterpret_cast<__half*>(ws + unit);       // trunk  t
    __half* bufM = reinterpret_cast<__half*>(ws + 2 * unit);   // ARSB / LB mid
    __half* bufC = reinterpret_cast<__half*>(ws + 3 * unit);   // LB conv_2 output before the FRM gate (MoeNet_lite2)
    uint8_t* up0 = ws + 4 * unit;

    // conv_input + PReLU                                                       models.py:118
    FirstParams fp{};
    fp.img = static_cast<const __half*>(in); fp.plane_stride = in_plane_stride; fp.row_stride = in_row_stride;
    fp.in_h = plan->in_h; fp.in_w = plan->in_w; fp.pad_h = plan->pad_h; fp.pad_w = plan->pad_w;
    fp.top = t.top + g.c0; fp.left = t.left; fp.N = N; fp.H = H; fp.W = W;
    fp.w = m->first_w; fp.slope = m->scalars[0]; fp.out = bufA;
    if (N > 65535) return fail(MOE_ERR_INVALID, "too many planes for the conv_input grid");
    {
    Timed timed(e, st, 0, static_cast<double>(N) * H * W * (2 + 128));     // bytes: read 1 fp16, write 64 fp16
    int first_rows = 0, first_nseg = 0;
    // (the kernel is FP32-FMA-issue bound — 576 FMAs per pixel — not occupancy bound: 2 to 12 blocks per SM all give 1.39-1.46 ms per 4K frame)
    choose_segments(static_cast<int64_t>(N) * ((W + 127) / 128), 2 * e->sm_count, H, 8, &first_rows, &first_nseg);
    conv_first_kernel<<<dim3((W + 127) / 128, first_nseg, N), 256, 0, st>>>(fp, first_rows);
    }
    if ((rc = check_launch(e, "conv_first_kernel")) != MOE_OK) return rc;

    const int feat = e->full_k ? 64 : m->feat;                  // real input channels per tap (K steps of 16); full_k: test switch
    const int one_by_one = m->arch == MOE_ARCH_LITE;            // MoeNet_lite2: conv_input2 and the upsample convs are 1x1
    if ((rc = launch_conv(e, st, bufA, bufT, nullptr, m->trunk_img[0], nullptr, nullptr, feat, N, H, W, 1, EPI_PLAIN, 0.f, one_by_one)) != MOE_OK) return rc;   // conv_input2
    if (m->arch == MOE_ARCH_LITE) {
      // three LB blocks: t = FRM(conv_2(PReLU(conv_1(t)))) + t                  MoeNet_lite2.py:7-20, models.py:270-287
      float* partial = reinterpret_cast<float*>(ws + tile_units(m) * unit);
      float* gate = partial + static_cast<size_t>(kFrmBlocks) * 64 * N;
      const int64_t px = static_cast<int64_t>(H) * W;
      for (int b = 0; b < 3; ++b) {
        const int l1 = 1 + 2 * b, l2 = 2 + 2 * b;
        if ((rc = launch_conv(e, st, bufT, bufM, nullptr, m->trunk_img[l1], nullptr, nullptr, feat, N, H, W, 1, EPI_PRELU, m->scalars[1 + l1])) != MOE_OK) return rc;
        if ((rc = launch_conv(e, st, bufM, bufC, nullptr, m->trunk_img[l2], nullptr, nullptr, feat, N, H, W, 1, EPI_PLAIN, 0.f)) != MOE_OK) return rc;
        Timed timed(e, st, 6, static_cast<double>(N) * px * 128 * 4);               // bytes: v read twice, t read and written
        frm_partial_kernel<<<dim3(kFrmBlocks, N), 256, 0, st>>>(bufC, partial, px);
        if ((rc = check_launch(e, "frm_partial_kernel")) != MOE_OK) return rc;
        frm_gate_kernel<<<N, 64, 0, st>>>(partial, m->frm[b], gate, 1.0f / static_cast<float>(px), e->bias_fused);
        if ((rc = check_launch(e, "frm_gate_kernel")) != MOE_OK) return rc;
        frm_apply_kernel<<<grid_for(px * N * 8, 256, e->sm_count), 256, 0, st>>>(bufC, bufT, gate, px, N);
        if ((rc = check_launch(e, "frm_apply_kernel")) != MOE_OK) return rc;
      }
    } else {
      // six ARSBs: t += scale * conv_2(PReLU(conv_1(t)))                        models.py:76-80
      const bool fused_arsb = !e->simt && !e->no_pair && !e->no_pair_trunk && !e->no_arsb && e->sm_count >= 2;
      __half* cur = bufT;
      __half* other = bufM;                                     // the fused kernel cannot update t in place (strips read each other's halo columns)
      for (int b = 0; b < 6; ++b) {
        const int l1 = 1 + 2 * b, l2 = 2 + 2 * b;
        if (fused_arsb) {
          if ((rc = launch_arsb(e, st, cur, other, m->trunk_img[l1], m->trunk_img[l2], feat, N, H, W, m->scalars[1 + l1], m->scalars[1 + l2])) != MOE_OK) return rc;
          std::swap(cur, other);                                 // six blocks: the result ends up in bufT again
        } else {
          if ((rc = launch_conv(e, st, bufT, bufM, nullptr, m->trunk_img[l1], nullptr, nullptr, feat, N, H, W, 1, EPI_PRELU, m->scalars[1 + l1])) != MOE_OK) return rc;
          if ((rc = launch_conv(e, st, bufM, bufT, bufT, m->trunk_img[l2], nullptr, nullptr, feat, N, H, W, 1, EPI_SCALE_SKIP, m->scalars[1 + l2])) != MOE_OK) return rc;
        }
      }
    }

    // the two upsample stacks: branch 0 on `out`, branch 1 on the trunk           models.py:29-33,125-154; MoeNet_lite2.py:47-50
    const __half* head_in[2] = {bufA, bufT};
    const bool fuse = !e->simt && !e->no_pair && !e->no_fuse && m->n_up >= 1 && e->sm_count / 2 >= pair_groups(m->r);
    float* hbuf[2] = {nullptr, nullptr};
    float* ebuf[2] = {nullptr, nullptr};
    int estrips = 0;
    if (m->n_up >= 1 && m->r == 2) {
      __half* stage_buf[2] = {reinterpret_cast<__half*>(up0), reinterpret_cast<__half*>(up0 + 4 * unit)};   // S1 (4 units), S2 (16 units)
      uint8_t* fin = up0 + stage_units(m) * unit;
      size_t fin_units = 1;
      for (int i = 0; i < m->n_up; ++i) fin_units *= 4;
      // fused: per branch the three H_dy planes (12 of the 128 B per output pixel) and, behind them, the strip-edge terms
      estrips = 2 * ((((W << (m->n_up - 1)) + kStripW - 1) / kStripW + 1) / 2);
      for (int b = 0; b < 2; ++b) {
        hbuf[b] = reinterpret_cast<float*>(fin + b * fin_units * unit);
        ebuf[b] = hbuf[b] + static_cast<size_t>(N) * 3 * (static_cast<size_t>(H) * sc) * (static_cast<size_t>(W) * sc);
      }
      for (int b = 0; b < 2; ++b) {
        const __half* src = b ? bufT : bufA;
        for (int s2 = 0; s2 < m->n_up; ++s2) {
          const int hs = H << s2, wsz = W << s2;
          const bool last = s2 == m->n_up - 1;
          const uint8_t* wimg = m->up_img[4 * b + s2];
          const float* wb = m->up_bias[4 * b + s2];
          const float slope = m->scalars[14 + 4 * b + s2];
          if (last && fuse) {
            if ((rc = launch_conv_head(e, st, src, wimg, wb, &m->up_bias_h[4 * b + s2], feat, N, hs, wsz, 2, slope, m->d_head_img + b * 2048, hbuf[b], ebuf[b], one_by_one)) != MOE_OK) return rc;
          } else {
            __half* dst = last ? reinterpret_cast<__half*>(fin + b * fin_units * unit) : stage_buf[s2];
            if ((rc = launch_conv(e, st, src, dst, nullptr, wimg, wb, &m->up_bias_h[4 * b + s2], feat, N, hs, wsz, 2, EPI_BIAS_PRELU, slope, one_by_one)) != MOE_OK) return rc;
            src = dst;
            if (last) head_in[b] = dst;
          }
        }
      }
    } else if (m->n_up == 1) {                                     // PixelShuffle(3)
      const size_t usz = unit * 9;
      for (int b = 0; b < 2; ++b) {
        __half* dst = reinterpret_cast<__half*>(up0 + b * usz);
        if (fuse) {                                                // 81 floats per input pixel-plane (324 of the 1 152 B) per branch
          hbuf[b] = reinterpret_cast<float*>(dst);
          if ((rc = launch_conv_head(e, st, b ? bufT : bufA, m->up_img[4 * b], m->up_bias[4 * b], &m->up_bias_h[4 * b], feat, N, H, W, 3,
                                     m->scalars[14 + 4 * b], m->d_head_img + b * 2048, hbuf[b], nullptr)) != MOE_OK) return rc;
          continue;
        }
        if ((rc = launch_conv(e, st, b ? bufT : bufA, dst, nullptr, m->up_img[4 * b], m->up_bias[4 * b], &m->up_bias_h[4 * b], feat, N, H, W, 3,
                              EPI_BIAS_PRELU, m->scalars[14 + 4 * b])) != MOE_OK) return rc;
        head_in[b] = dst;
      }
    }

    // Conv3x3(F,1) heads, branch sum, seam blend, canvas store          models.py:38; imageProcess.py:164-170
    HeadParams hp{};
    hp.u = head_in[0]; hp.r = head_in[1]; hp.wu = m->head_w[0]; hp.wr = m->head_w[1];
    hp.N = N; hp.H = H * sc; hp.W = W * sc;
    hp.oy = (t.top + g.c0) * sc; hp.ox = t.left * sc;
    hp.keep_y0 = g.keep_y0; hp.keep_y1 = g.keep_y1; hp.keep_x0 = g.keep_x0; hp.keep_x1 = g.keep_x1;
    hp.ramp_y0 = g.ramp_y0; hp.ramp_x0 = g.ramp_x0; hp.blend_y1 = g.blend_y1; hp.blend_x1 = g.blend_x1;
    for (int i = 0; i < plan->pad_sc; ++i) hp.ramp[i] = plan->ramp[i];
    hp.canvas = static_cast<__half*>(canvas);
    hp.plane_stride = out_plane_stride; hp.row_stride = out_row_stride;
    if (fuse && m->r == 3) {
      HeadStencil9Params sp{};
      sp.g = hp; sp.pu = hbuf[0]; sp.pr = hbuf[1]; sp.h = H; sp.w = W;
      dim3 sgrid((W + 127) / 128, std::min(H, 65535), N);
      if (sgrid.z > 65535u) return fail(MOE_ERR_INVALID, "too many planes for the stencil kernel grid");
      Timed timed(e, st, 2, static_cast<double>(N) * hp.H * hp.W * (72 + 2));   // bytes: two 9-float reads, one fp16 write
      head_stencil9_kernel<<<sgrid, 128, 0, st>>>(sp);
    } else if (fuse) {
      HeadStencilParams sp{};
      sp.g = hp; sp.hu = hbuf[0]; sp.hr = hbuf[1]; sp.eu = ebuf[0]; sp.er = ebuf[1]; sp.estrips = estrips;
      dim3 sgrid((hp.W + kStencilThreads * kStencilPx - 1) / (kStencilThreads * kStencilPx), std::min(hp.H, 65535), N);
      if (sgrid.z > 65535u) return fail(MOE_ERR_INVALID, "too many planes for the stencil kernel grid");
      Timed timed(e, st, 2, static_cast<double>(N) * hp.H * hp.W * (24 + 2));   // bytes: two 3-float reads, one fp16 write
      head_stencil_kernel<<<sgrid, kStencilThreads, 0, st>>>(sp);
    } else if (e->simt) {
      dim3 hgrid((hp.W + 127) / 128, hp.H, N);
      if (hgrid.y > 65535u || hgrid.z > 65535u) return fail(MOE_ERR_INVALID, "tile too tall for the head kernel grid");
      Timed timed(e, st, 2, static_cast<double>(N) * hp.H * hp.W * (2 * 128 + 2));   // bytes: two 64-ch fp16 reads, one fp16 write
      head_blend_kernel<<<hgrid, 128, 0, st>>>(hp);
    } else {
      if ((rc = launch_head_tc(e, st, m, hp)) != MOE_OK) return rc;
    }
    if ((rc = check_launch(e, "head / stencil kernel")) != MOE_OK) return rc;
    if (after_tile && (rc = after_tile(ti, after_ctx)) != MOE_OK) return rc;
  }
  return MOE_OK;
}
}  // namespace

extern "C" {

int moe_conv3x3_c64(MoeEngine* e, const void* in, void* out, const void* skip, const void* w_img, const float* bias,
                    int n, int h, int w, int r, int epi, float param, void* stream)
{
  if (!e || !in || !out || !w_img || n <= 0 || h <= 0 || w <= 0 || r < 1 || r > 3 || epi < 0 || epi > 3)
    return fail(MOE_ERR_INVALID, "bad argument");
  if ((epi == EPI_SCALE_SKIP && !skip) || (epi == EPI_BIAS_PRELU && !bias)) return fail(MOE_ERR_INVALID, "epilogue operand missing");
  Guard g(e->device);
  ConvBias bias_h{};                           // diagnostic entry: the bias comes as a device vector, the pair kernels want it as a parameter
  if (epi == EPI_BIAS_PRELU) {
    MOE_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    MOE_CUDA(cudaMemcpy(bias_h.v, bias, static_cast<size_t>(r) * r * 64 * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return launch_conv(e, static_cast<cudaStream_t>(stream), static_cast<const __half*>(in), static_cast<__half*>(out),
                     static_cast<const __half*>(skip), static_cast<const uint8_t*>(w_img), bias, &bias_h, 64, n, h, w, r, epi, param);
}

int moe_axpby_f16(MoeEngine* e, void* y, const void* x, float s, size_t count, void* stream)
{
  if (!e || !y || !x) return fail(MOE_ERR_INVALID, "null argument");
  if (count == 0) return MOE_OK;
  Guard g(e->device);
  axpby_f16_kernel<<<grid_for(static_cast<int64_t>(count), 256, e->sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__half*>(y), static_cast<const __half*>(x), s, 1.f - s, count);
  return check_launch(e, "axpby_f16_kernel");
}

int moe_to_planar_f16(MoeEngine* e, const void* src, int bits, int h, int w, int c, int swap_rb, void* dst, void* stream)
{
  if (!e || !src || !dst || h <= 0 || w <= 0 || c <= 0 || bits < 1 || bits > 16) return fail(MOE_ERR_INVALID, "bad argument");
  Guard g(e->device);
  const int grid = grid_for(static_cast<int64_t>(h) * w, 256, e->sm_count);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bits <= 8)
    to_planar_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(src), h, w, c, swap_rb, 0.f, 1, static_cast<__half*>(dst));
  else
    to_planar_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(src), h, w, c, swap_rb, 1.f / static_cast<float>(1 << bits), 0,
                                                      static_cast<__half*>(dst));
  return check_launch(e, "to_planar_kernel");
}

int moe_to_output(MoeEngine* e, const void* src, int bits, int h, int w, int c, int swap_rb, void* dst, void* stream)
{
  if (!e || !src || !dst || h <= 0 || w <= 0 || c <= 0 || bits < 1 || bits > 16) return fail(MOE_ERR_INVALID, "bad argument");
  Guard g(e->device);
  const int grid = grid_for(static_cast<int64_t>(h) * w, 256, e->sm_count);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float quant = static_cast<float>(1 << bits);
  if (bits <= 8)
    to_output_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const __half*>(src), h, w, c, swap_rb, quant, static_cast<uint8_t*>(dst));
  else
    to_output_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const __half*>(src), h, w, c, swap_rb, quant, static_cast<uint16_t*>(dst));
  return check_launch(e, "to_output_kernel");
}

static int ensure_buf(MoeEngine* e, int i, size_t bytes)
{
  if (e->cap[i] >= bytes) return MOE_OK;
  if (e->buf[i]) { cudaFree(e->buf[i]); e->buf[i] = nullptr; e->cap[i] = 0; }
  if (cudaMalloc(&e->buf[i], bytes) != cudaSuccess) { cudaGetLastError(); return fail(MOE_ERR_NOMEM, "cudaMalloc(%zu) failed", bytes); }
  e->cap[i] = bytes;
  return MOE_OK;
}

}  // extern "C"

namespace {
// toFloat + toOutput on a column range [x0,x1) of the canvas (same arithmetic as to_output_kernel)
template <typename T>
__global__ void to_output_cols_kernel(const __half* src, size_t plane, int h, int w, int c, int x0, int x1, float quant, T* dst)
{
  const int cols = x1 - x0;
  const size_t total = static_cast<size_t>(h) * cols;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t y = i / cols, px = y * w + x0 + (i - y * cols);
    for (int ch = 0; ch < c; ++ch) {
      float v = __half2float(src[ch * plane + px]) * quant;
      v = fminf(fmaxf(v, 0.f), quant - 1.f);
      dst[px * c + ch] = static_cast<T>(v);
    }
  }
}

// `canvas` points at canvas row `row_lo` of plane 0 (planes `plane` halfs apart), d_out / host_out at row `row_lo` of the
// interleaved result: the overlap works on a row band of the frame (the whole frame at N = 1)
struct OverlapCtx {
  MoeEngine* e; const MoePlan* plan; cudaStream_t compute, copy; cudaEvent_t ev;
  const __half* canvas; size_t plane; uint8_t* d_out; uint8_t* host_out; int bits_out; int done_x; int row_lo, row_hi; int channels;
};

// convert + copy the canvas columns [c->done_x, x_end) once everything enqueued so far on the compute stream is done
int flush_columns(OverlapCtx* c, int x_end)
{
  if (x_end <= c->done_x) return MOE_OK;
  const MoePlan* pl = c->plan;
  const int bpp = c->bits_out <= 8 ? 1 : 2, rows = c->row_hi - c->row_lo;
  MOE_CUDA(cudaEventRecord(c->ev, c->compute));
  MOE_CUDA(cudaStreamWaitEvent(c->copy, c->ev, 0));
  const int64_t threads = static_cast<int64_t>(rows) * (x_end - c->done_x);
  const int grid = grid_for(threads, 256, c->e->sm_count);
  const float quant = static_cast<float>(1 << c->bits_out);
  if (bpp == 1)
    to_output_cols_kernel<uint8_t><<<grid, 256, 0, c->copy>>>(c->canvas, c->plane, rows, pl->out_w, c->channels, c->done_x, x_end, quant, c->d_out);
  else
    to_output_cols_kernel<uint16_t><<<grid, 256, 0, c->copy>>>(c->canvas, c->plane, rows, pl->out_w, c->channels, c->done_x, x_end, quant, reinterpret_cast<uint16_t*>(c->d_out));
  int rc = check_launch(c->e, "to_output_cols_kernel");
  if (rc != MOE_OK) return rc;
  const size_t pitch = static_cast<size_t>(pl->out_w) * c->channels * bpp, off = static_cast<size_t>(c->done_x) * c->channels * bpp;
  MOE_CUDA(cudaMemcpy2DAsync(c->host_out + off, pitch, c->d_out + off, pitch, static_cast<size_t>(x_end - c->done_x) * c->channels * bpp, rows,
                             cudaMemcpyDeviceToHost, c->copy));
  c->done_x = x_end;
  return MOE_OK;
}

int overlap_after_tile(int ti, void* ctx)
{
  OverlapCtx* c = static_cast<OverlapCtx*>(ctx);
  const MoePlan* pl = c->plan;
  // a later tile rewrites the canvas from its own kept-region start: everything left of that is final
  const int x_end = ti + 1 < pl->n_tiles ? tile_geom(*pl, pl->tiles[ti + 1], c->row_lo, c->row_hi).ramp_x0 : pl->out_w;
  return flush_columns(c, x_end);
}
}  // namespace

extern "C" {

int moe_enhance_host_c(MoeModel* m, const void* host_in, int bits_in, int channels, const MoePlan* plan, void* host_out, int bits_out, void* stream)
{
  if (channels != 1 && channels != 3 && channels != 4) return fail(MOE_ERR_INVALID, "frames have 1, 3 or 4 channels, got %d", channels);
  // the reference's SR path upscales an alpha plane like a colour plane (a 4th batch element); its RGBFilter (DN) strips the
  // alpha plane, filters the colour planes and re-attaches it untouched (imageProcess.py:370-377)
  const bool alpha_bypass = m && m->scale == 1 && channels == 4;
  const int net_planes = alpha_bypass ? 3 : channels;
  int rc = validate_plan(m, plan, net_planes, 0, plan ? plan->out_h : 0);
  if (rc != MOE_OK) return rc;
  if (!host_in || !host_out) return fail(MOE_ERR_INVALID, "null host buffer");
  if (bits_in < 1 || bits_in > 16 || bits_out < 1 || bits_out > 16) return fail(MOE_ERR_INVALID, "bad bit depth");
  MoeEngine* e = m->e;
  Guard guard(e->device);
  std::lock_guard<std::mutex> lock(e->host_mutex);      // the engine-owned staging buffers below serve one call at a time
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t ipx = static_cast<size_t>(plan->in_h) * plan->in_w, opx = static_cast<size_t>(plan->out_h) * plan->out_w;
  const size_t in_raw = ipx * channels * (bits_in <= 8 ? 1 : 2), out_raw = opx * channels * (bits_out <= 8 ? 1 : 2);
  const size_t wsb = moe_plan_workspace_bytes(m, net_planes, plan, 0, plan->out_h);
  if ((rc = ensure_buf(e, 0, in_raw)) || (rc = ensure_buf(e, 1, ipx * channels * 2)) || (rc = ensure_buf(e, 2, opx * channels * 2)) ||
      (rc = ensure_buf(e, 3, out_raw)) || (rc = ensure_buf(e, 4, wsb)))
    return rc;
  MOE_CUDA(cudaMemcpyAsync(e->buf[0], host_in, in_raw, cudaMemcpyHostToDevice, st));
  if ((rc = moe_to_planar_f16(e, e->buf[0], bits_in, plan->in_h, plan->in_w, channels, 0, e->buf[1], st)) != MOE_OK) return rc;
  if (alpha_bypass)
    MOE_CUDA(cudaMemcpyAsync(static_cast<__half*>(e->buf[2]) + 3 * opx, static_cast<const __half*>(e->buf[1]) + 3 * ipx, ipx * 2, cudaMemcpyDeviceToDevice, st));
  // One row of tiles (the reference's plan for big frames: column strips): a tile's columns are final as soon as the
  // NEXT tile's kept region starts, so their conversion and device->host copy run on a second stream under the next
  // tile's compute.  Several tile rows: convert and copy everything at the end.
  bool one_row = true;
  for (int i = 0; i < plan->n_tiles; ++i) one_row = one_row && plan->tiles[i].top == 0 && plan->tiles[i].bsc == plan->out_h;
  if (!e->copy_stream) {
    MOE_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    MOE_CUDA(cudaEventCreateWithFlags(&e->copy_event, cudaEventDisableTiming));
  }
  OverlapCtx ctx{e, plan, st, e->copy_stream, e->copy_event, static_cast<const __half*>(e->buf[2]), opx, static_cast<uint8_t*>(e->buf[3]),
                 static_cast<uint8_t*>(host_out), bits_out, 0, 0, plan->out_h, channels};
  if ((rc = run_plan_core(m, e->buf[1], static_cast<int64_t>(ipx), plan->in_w, net_planes, e->buf[2], static_cast<int64_t>(opx), plan->out_w,
                          plan, 0, plan->out_h, e->buf[4], e->cap[4], st, one_row ? overlap_after_tile : nullptr, &ctx)) != MOE_OK) return rc;
  if ((rc = flush_columns(&ctx, plan->out_w)) != MOE_OK) return rc;
  MOE_CUDA(cudaStreamSynchronize(e->copy_stream));
  return moe_engine_check(e, stream);
}

int moe_enhance_host(MoeModel* m, const void* host_in, int bits_in, const MoePlan* plan, void* host_out, int bits_out, void* stream)
{
  return moe_enhance_host_c(m, host_in, bits_in, 3, plan, host_out, bits_out, stream);
}

int moe_run_band_to_host(MoeModel* m, const void* in, int64_t in_plane_stride, int64_t in_row_stride, int planes, const MoePlan* plan,
                         int row_lo, int row_hi, void* host_out, int bits_out, void* stream)
{
  int rc = validate_plan(m, plan, planes, row_lo, row_hi);
  if (rc != MOE_OK) return rc;
  if (!in || !host_out) return fail(MOE_ERR_INVALID, "null buffer");
  if (bits_out < 1 || bits_out > 16) return fail(MOE_ERR_INVALID, "bad bit depth");
  MoeEngine* e = m->e;
  Guard guard(e->device);
  std::lock_guard<std::mutex> lock(e->host_mutex);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = row_hi - row_lo, bpp = bits_out <= 8 ? 1 : 2;
  const size_t band_px = static_cast<size_t>(rows) * plan->out_w;
  const size_t wsb = moe_plan_workspace_bytes(m, planes, plan, row_lo, row_hi);
  if ((rc = ensure_buf(e, 2, band_px * planes * 2)) || (rc = ensure_buf(e, 3, band_px * planes * bpp)) || (rc = ensure_buf(e, 4, wsb))) return rc;
  bool one_row = true;
  for (int i = 0; i < plan->n_tiles; ++i) one_row = one_row && plan->tiles[i].top == 0 && plan->tiles[i].bsc == plan->out_h;
  if (!e->copy_stream) {
    MOE_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    MOE_CUDA(cudaEventCreateWithFlags(&e->copy_event, cudaEventDisableTiming));
  }
  // the band canvas holds rows [row_lo, row_hi) only: frame row r lives at band row r - row_lo (same strides, shifted base)
  __half* band = static_cast<__half*>(e->buf[2]);
  __half* base = band - static_cast<int64_t>(row_lo) * plan->out_w;
  uint8_t* host_band = static_cast<uint8_t*>(host_out) + static_cast<size_t>(row_lo) * plan->out_w * planes * bpp;
  OverlapCtx ctx{e, plan, st, e->copy_stream, e->copy_event, band, band_px, static_cast<uint8_t*>(e->buf[3]), host_band, bits_out, 0, row_lo, row_hi, planes};
  if ((rc = run_plan_core(m, in, in_plane_stride, in_row_stride, planes, base, static_cast<int64_t>(band_px), plan->out_w, plan, row_lo, row_hi,
                          e->buf[4], e->cap[4], st, one_row ? overlap_after_tile : nullptr, &ctx)) != MOE_OK) return rc;
  if ((rc = flush_columns(&ctx, plan->out_w)) != MOE_OK) return rc;
  MOE_CUDA(cudaStreamSynchronize(e->copy_stream));
  return moe_engine_check(e, stream);
}

}  // extern "C"
