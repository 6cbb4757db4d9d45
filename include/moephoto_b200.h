/* moephoto_b200 — C ABI of the Blackwell (sm_100a) tiled SR / denoise engine.
 *
 * This is the drop-in boundary for MoePhoto's L3->L2 seam.  The reference has no
 * FFI on this path (it is PyTorch eager); the entry points below are what a binding for the path
 * replaces, one for one:
 *
 *   moe_model_load      <- imageProcess.initModel / getStateDict / castModel   imageProcess.py:304-334
 *                          (+ the nn.Module constructors Net2x/Net3x/Net4x/NetDN  models.py:125-164)
 *   moe_plan_workspace_bytes / moe_run_plan
 *                       <- imageProcess.doCrop  (per-tile net call + blend + canvas store)
 *                                                                          imageProcess.py:157-172
 *                          with the per-tile network  MyNet.forward        models.py:117-123
 *                          the seam blend             blend                 imageProcess.py:120-131
 *                          and padImage / unpad       getPad                imageProcess.py:48-56,100-108
 *   moe_axpby_f16       <- strengthOp                                       imageProcess.py:562
 *   moe_to_planar_f16   <- toTorch (uint8 /255, wider /2^bits)              imageProcess.py:259-263
 *   moe_to_output       <- toFloat + toOutput (x2^bits, clamp, truncate)    imageProcess.py:238-257
 *   moe_enhance_host    <- the whole file->SR->output step chain of procedure.genProcess
 *                          (procedure.py:156-201) for one image held in HOST memory.
 *
 * The tile list itself (prepare / getAnchors, imageProcess.py:19-35,73-118) is integer host logic and
 * stays on the host: the caller passes it in as a MoePlan.
 *
 * Conventions: plain pointers and sizes only; every function returns MOE_OK (0) or a negative
 * MoeStatus and never aborts the process (not even on a device-side pipeline time-out: moe_engine_check); moe_last_error() gives the message for the calling thread's
 * last failure.  All device pointers are on the engine's device.  `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream); work is enqueued on it and the call does not synchronise.
 * Concurrency: calls on one engine may come from several host threads and several streams.  Every convolution launch draws
 * its work items from its own counter block, MoeNet_lite2's reduction scratch lives in the workspace the CALLER passes, so two
 * moe_run_plan calls may overlap on two streams provided they are given two workspaces (and two canvases);
 * moe_enhance_host* uses engine-owned staging buffers and serialises itself on a mutex.  The A/B switches
 * (moe_engine_set_conv_path), the profiler and the debug buffer are engine-wide settings: change them only while no call is in
 * flight.
 * There is NO CPU fallback: without a usable sm_100 device every compute entry point fails with
 * MOE_ERR_NO_DEVICE.
 */
#ifndef MOEPHOTO_B200_H
#define MOEPHOTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOE_ABI_VERSION 1

typedef enum {
  MOE_OK = 0,
  MOE_ERR_INVALID = -1,      /* bad argument / malformed blob or plan            -> ValueError     */
  MOE_ERR_NO_DEVICE = -2,    /* no CUDA device, or not compute capability 10.x   -> RuntimeError   */
  MOE_ERR_NOMEM = -3,        /* workspace too small / allocation failed          -> MemoryError    */
  MOE_ERR_CUDA = -4          /* a CUDA call or kernel launch failed              -> RuntimeError   */
} MoeStatus;

/* network families on the hot path (runSR.py:11-16, runDN.py:16-18) */
typedef enum {
  MOE_ARCH_NET2X = 2,        /* models.Net2x : a2 / p2 */
  MOE_ARCH_NET3X = 3,        /* models.Net3x : a3 / p3 */
  MOE_ARCH_NET4X = 4,        /* models.Net4x : a4 / p4 */
  MOE_ARCH_NETDN = 1,        /* models.NetDN : dn_lite5/10/15 */
  MOE_ARCH_LITE = 5          /* MoeNet_lite2.Net : lite2 / lite4 / lite8 (runSR.py:21-23) */
} MoeArch;

typedef struct MoeEngine MoeEngine;
typedef struct MoeModel MoeModel;

/* One element of prepare()'s iterClip() (imageProcess.py:111-117), LR units unless noted. */
typedef struct {
  int32_t top, bottom, left, right;  /* tile rectangle in the (right/bottom padded) input image   */
  int32_t top_t, left_t;             /* blend anchors: 0 = first tile, >0 = seam ends here,
                                        <0 = seam ends |v| px before the tile's end (output px)     */
  int32_t bsc, rsc;                  /* bottom / right edge of the tile on the canvas (output px)   */
} MoeTile;

typedef struct {
  int32_t n_tiles;
  int32_t scale;                     /* 1 (DN), 2, 3, 4                                              */
  int32_t pad_sc;                    /* seam width in output px = padding * scale                    */
  int32_t in_h, in_w;                /* input image size (before padImage)                           */
  int32_t pad_h, pad_w;              /* rows / cols appended by padImage (reflect, then zeros)       */
  int32_t out_h, out_w;              /* canvas size = in_h*scale, in_w*scale                         */
  const MoeTile* tiles;              /* n_tiles entries, processed in this order                     */
  const float* ramp;                 /* pad_sc blend weights sigma(9(i/pad_sc - 1/2)) (:109), already
                                        rounded to the canvas dtype (fp16)                            */
} MoePlan;

/* ---- engine ------------------------------------------------------------------------------- */
int moe_abi_version(void);
const char* moe_last_error(void);
int moe_engine_create(int device_id, MoeEngine** out);
void moe_engine_destroy(MoeEngine* e);
/* number of this library's kernels launched since the engine was created (bench.py "gpu_launches") */
int64_t moe_engine_launch_count(const MoeEngine* e);
/* Per-launch device timing for roofline reports: while enabled, every kernel launch is bracketed by a CUDA event pair on its
 * stream (two extra driver calls per launch: keep it OFF for throughput measurements).  _read waits for them and returns per
 * kernel class the summed milliseconds, the summed ALGORITHMIC work (FLOPs for the convolution classes — padded channels,
 * halo columns and scrap outputs are not counted — bytes otherwise) and the launch count, and resets the counters.  Classes:
 *   0 conv_first_kernel (bytes)                      4 arsb_pair_kernel: one residual block, two convolutions (FLOPs)
 *   1 3x3 conv 64->64 as its own launch (FLOPs)      5 conv3x3_pair_head_kernel: last upsample conv + head dot products (FLOPs)
 *   2 head / stencil / blend / store (bytes)         6 MoeNet_lite2 FRM reduction + gate + apply (bytes)
 *   3 upsample conv as its own launch (FLOPs)        7 reserved */
#define MOE_PROFILE_CLASSES 8
int moe_engine_profile(MoeEngine* e, int enable);
int moe_engine_profile_read(MoeEngine* e, double ms[MOE_PROFILE_CLASSES], double work[MOE_PROFILE_CLASSES], int64_t launches[MOE_PROFILE_CLASSES]);
/* bit 0: 0 = tcgen05 tensor-core kernels (default), 1 = plain SIMT kernels (debug cross-check);
 * bit 1: 1 = keep every convolution on the single-CTA kernel instead of CTA pairs (A/B switch);
 * bit 2: 1 = only the 64->64 trunk convolutions stay on the single-CTA kernel;
 * bit 3: 1 = do not fuse the last upsample convolution with the heads' dot products;
 * bit 4: 1 = the CTA-pair kernels deal their work items round-robin instead of drawing them from a counter;
 * bit 5: numerics of a convolution WITH a bias (models.py:29-30): 0 (default) = q(q(conv) + bias), what the reference's GPU
 *        path computes (aten: cudnn_convolution, then add_ of the bias — two ops, two fp16 roundings), 1 = q(conv + bias), what
 *        the same half model computes on the CPU (oneDNN adds the bias inside the convolution; tests/golden `.ref16`);
 * bit 6: 1 = run every residual block (ARSB) as two convolution launches instead of the fused arsb_pair_kernel;
 * bit 7: 1 = the fused residual block keeps its intermediate rows in shared memory (.ss conv_2) instead of tensor memory (.ts);
 * bit 8: 1 = the fused residual block on single CTAs with all weights per SM (arsb_solo_kernel) instead of CTA pairs — slower;
 *        the experiment that located the 64->64 layers' limit (profiles/r02_arsb_experiments.txt);
 * bit 9: 1 = NetDN / MoeNet_lite2 (48 filters, zero-padded to 64) issue all four K steps of every tap instead of skipping the
 *        all-zero fourth one (test switch: the results are bit-identical) */
int moe_engine_set_conv_path(MoeEngine* e, int simt);
/* Kernels wait on mbarriers with a time-out (4 s of wall time).  A wait that gives up does NOT trap — round 1's __trap() destroyed
 * the CUDA context of the whole host process, i.e. MoePhoto's worker and every cached model, and a slow wait (a time-sliced or
 * profiled GPU) is not an error at all — it raises a device-side flag and the kernel runs on to its end; what it computes after that
 * is invalid.  moe_engine_check synchronises `stream` and, if the flag is up, clears it, resets the work-item counters and returns
 * MOE_ERR_CUDA.  moe_enhance_host* / moe_run_band_to_host call it themselves; after moe_run_plan (asynchronous) it is the caller's to
 * call where it synchronises anyway.  moe_engine_debug_timeout sets the time-out in ns (0 = the default). */
int moe_engine_check(MoeEngine* e, void* stream);
int moe_engine_debug_timeout(MoeEngine* e, uint64_t ns);
/* Diagnostics: `dev` = device buffer of >= 4 * 8 bytes per SM pair (or NULL to switch off).  Every CTA-pair convolution
 * launch then leaves {start ns, end ns, SM id, items processed} per pair in it (the last launch wins). */
int moe_engine_debug_buffer(MoeEngine* e, void* dev, size_t nbytes);

/* ---- model -------------------------------------------------------------------------------- */
/* `blob` is HOST memory in the packed format produced by moephoto_b200/weights.py (layout documented
 * in csrc/blob.h); it is copied to the device, the caller may free it afterwards. */
int moe_model_load(MoeEngine* e, int arch, const void* blob, size_t nbytes, MoeModel** out);
void moe_model_free(MoeModel* m);
int moe_model_scale(const MoeModel* m);

/* ---- doCrop ------------------------------------------------------------------------------- */
/* Only canvas rows [row_lo,row_hi) are produced (row_lo = 0, row_hi = out_h for the whole image):
 * a rank that owns a row band computes each tile restricted to the band plus a 16-LR-px recompute
 * halo, which is exact (receptive field radius 15.75 LR px). */
size_t moe_plan_workspace_bytes(const MoeModel* m, int planes, const MoePlan* plan, int row_lo, int row_hi);
/* in : planes x in_h x in_w fp16, element strides given (column stride 1)
 * out: planes x out_h x out_w fp16 canvas, element strides given (column stride 1)                */
int moe_run_plan(MoeModel* m,
                 const void* in, int64_t in_plane_stride, int64_t in_row_stride, int planes,
                 void* canvas, int64_t out_plane_stride, int64_t out_row_stride,
                 const MoePlan* plan, int row_lo, int row_hi,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- one layer (unit-test / profiling hook) -------------------------------------------------- */
/* 3x3 convolution, 64 -> 64*r*r channels, NHWC fp16 (the building block of ARSB, models.py:76-80, and of
 * genUpsampleBlock, models.py:29-33).  in: (n,h,w,64); out: (n,h*r,w*r,64); w_img: r*r swizzled 73 728-byte
 * images ON THE DEVICE (csrc/blob.h); bias: r*r*64 floats on the device or NULL.
 * epi: 0 plain, 1 PReLU(param), 2 out = skip + param*conv (skip may alias out), 3 PReLU(conv+bias).
 * With epi 3 this hook synchronises `stream` and copies the bias to the host (the CTA-pair kernels take it as a kernel
 * parameter; a loaded model keeps a host copy, so moe_run_plan never does this). */
int moe_conv3x3_c64(MoeEngine* e, const void* in, void* out, const void* skip, const void* w_img, const float* bias,
                    int n, int h, int w, int r, int epi, float param, void* stream);

/* ---- elementwise neighbours of the path --------------------------------------------------- */
/* y = s*y + (1-s)*x on fp16, each product and the sum rounded to fp16 (strengthOp) */
int moe_axpby_f16(MoeEngine* e, void* y, const void* x, float s, size_t count, void* stream);
/* src: h x w x c interleaved integers, bits<=8 -> uint8 (/255), else uint16 (/2^bits);
 * dst: c x h x w fp16.  `swap_rb` reads BGR as RGB (BGR2RGBTorch, procedure.py:129-135). */
int moe_to_planar_f16(MoeEngine* e, const void* src, int bits, int h, int w, int c, int swap_rb,
                      void* dst, void* stream);
/* src: c x h x w fp16 ; dst: h x w x c, bits<=8 -> uint8, <=16 -> uint16 (the reference's int16/int32
 * detour ends in the same bytes once toBuffer casts, imageProcess.py:231-236) */
int moe_to_output(MoeEngine* e, const void* src, int bits, int h, int w, int c, int swap_rb,
                  void* dst, void* stream);

/* ---- whole step chain on HOST buffers (file -> SR/DN -> output) --------------------------- */
/* host_in : in_h x in_w x 3 integers (bits as above); host_out: out_h x out_w x 3.
 * Copies in, converts, runs the plan, converts and copies out on `stream`, then synchronises it.
 * Uses the engine's internal device buffers (grown on demand). */
int moe_enhance_host(MoeModel* m, const void* host_in, int bits_in, const MoePlan* plan,
                     void* host_out, int bits_out, void* stream);
/* The same for frames of `channels` = 1 (grey), 3 or 4 (RGBA) interleaved samples.  As in the reference, an SR model upscales
 * the alpha plane like a colour plane (one more element of the plane batch, runSR.py:39-40), a DN model (scale 1) filters the
 * colour planes and passes alpha through untouched (_RGBFilter, imageProcess.py:370-377). */
int moe_enhance_host_c(MoeModel* m, const void* host_in, int bits_in, int channels, const MoePlan* plan,
                       void* host_out, int bits_out, void* stream);

/* One row band of the result, for a frame sharded over several GPUs (moephoto_b200/parallel.py): `in` is the planar fp16 frame in
 * DEVICE memory (it may be peer-mapped memory of another GPU: conv_first_kernel reads it in place), canvas rows [row_lo,row_hi)
 * are computed into an engine-owned band buffer, converted and copied into rows [row_lo,row_hi) of `host_out` (the whole
 * out_h x out_w x planes frame, e.g. a page-locked /dev/shm mapping every rank writes its own band of).  As in moe_enhance_host
 * the conversion and the device->host copy of a tile's finished columns run under the next tile's compute. */
int moe_run_band_to_host(MoeModel* m, const void* in, int64_t in_plane_stride, int64_t in_row_stride, int planes,
                         const MoePlan* plan, int row_lo, int row_hi, void* host_out, int bits_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOEPHOTO_B200_H */
