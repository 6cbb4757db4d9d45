// CUDA-core kernels of the path: the degenerate 1->F and F->1 convolutions (bandwidth-bound, no GEMM
// shape worth a tensor core tile yet), the fused seam blend + canvas store, the integer<->fp16 frame
// conversions, and a plain direct convolution used ONLY as an on-device cross-check of the tcgen05 path.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>
#include "conv_tc.cuh"

namespace moe {

// index in the source image of padded-image coordinate v (padImage: reflect first, then zeros;
// imageProcess.py:48-56).  -1 = zero.
__device__ __forceinline__ int pad_src(int v, int size, int padded) {
  if (v < 0 || v >= padded) return -1;
  if (v < size) return v;
  const int e = v - size;
  const int refl = min(size - 1, padded - size);
  return e < refl ? size - 2 - e : -1;
}

struct FirstParams {
  const __half* img;               // planes x in_h x in_w (strided)
  int64_t plane_stride, row_stride;
  int in_h, in_w, pad_h, pad_w;    // padded image is (in_h+pad_h) x (in_w+pad_w)
  int top, left;                   // origin of the computed rectangle in padded-image coordinates
  int N, H, W;                     // computed rectangle; zero outside it (tile-border zero padding)
  const float* w;                  // [9][64]
  float slope;
  __half* out;                     // NHWC (N,H,W,64)
};

// conv_input (1 -> 64, 3x3) + PReLU  (models.py:112,118).  A block owns a 128-pixel-wide column strip of one plane
// and walks down `seg_rows` rows with a rolling 4-row window in shared memory: one new input row (130 values,
// padImage / tile-border logic evaluated once per element) is fetched while the current row is computed, the 72
// weights a thread needs stay in registers for the whole walk.  Every thread computes 8 output channels of 4
// consecutive pixels (8 threads = one 4-pixel group, so each store instruction of a warp writes four full 128-byte
// pixels).  2 B read, 128 B written per pixel.
__device__ __forceinline__ float first_fetch(const FirstParams& p, int n, int yy, int xx) {
  if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) return 0.f;
  const int sy = pad_src(p.top + yy, p.in_h, p.in_h + p.pad_h);
  const int sx = pad_src(p.left + xx, p.in_w, p.in_w + p.pad_w);
  return (sy >= 0 && sx >= 0) ? __half2float(p.img[n * p.plane_stride + sy * p.row_stride + sx]) : 0.f;
}

__global__ void __launch_bounds__(256) conv_first_kernel(const FirstParams p, int seg_rows)
{
  __shared__ float win_s[4][132];
  const int x0 = blockIdx.x * 128;
  const int y0 = blockIdx.y * seg_rows;
  const int y1 = min(p.H, y0 + seg_rows);
  const int n = blockIdx.z;
  const int t = threadIdx.x;
  // rows y0-1, y0, y0+1 of the window
  for (int i = t; i < 3 * 130; i += 256) {
    const int r = i / 130, c = i - r * 130;
    win_s[(y0 - 1 + r) & 3][c] = first_fetch(p, n, y0 - 1 + r, x0 + c - 1);
  }
  const int g = t & 7;
  const int xl = (t >> 3) * 4;                     // first of this thread's 4 pixels, relative to x0
  const int x = x0 + xl;
  float wreg[9][8];                                // this thread's 8 output channels of every tap
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + k * 64 + g * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + k * 64 + g * 8 + 4));
    wreg[k][0] = w0.x; wreg[k][1] = w0.y; wreg[k][2] = w0.z; wreg[k][3] = w0.w;
    wreg[k][4] = w1.x; wreg[k][5] = w1.y; wreg[k][6] = w1.z; wreg[k][7] = w1.w;
  }
  __syncthreads();
  for (int y = y0; y < y1; ++y) {
    // fetch the row two below early (its latency hides behind this row's math); stored after the compute
    const float nxt = t < 130 ? first_fetch(p, n, y + 2, x0 + t - 1) : 0.f;
    if (x < p.W) {
      float win[3][6];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 6; ++dx) win[dy][dx] = win_s[(y - 1 + dy) & 3][xl + dx];
      const size_t pix0 = (static_cast<size_t>(n) * p.H + y) * p.W + x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (x + i >= p.W) break;
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float v = win[dy][i + dx];
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = fmaf(v, wreg[dy * 3 + dx][c], a[c]);
          }
        uint32_t w[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float f0 = h_round(a[2 * c]), f1 = h_round(a[2 * c + 1]);      // conv2d and prelu are two aten ops: two roundings
          f0 = f0 >= 0.f ? f0 : p.slope * f0;
          f1 = f1 >= 0.f ? f1 : p.slope * f1;
          const __half2 hv = __floats2half2_rn(f0, f1);
          w[c] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        reinterpret_cast<uint4*>(p.out + (pix0 + i) * 64)[g] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    // row y+2 replaces row y-2 in the ring: nobody reads slot (y+2)&3 during this iteration (rows y-1..y+1 are read)
    if (t < 130) win_s[(y + 2) & 3][t] = nxt;
    __syncthreads();
  }
}

constexpr int kMaxSeam = 64;

struct HeadParams {
  const __half* u;                 // NHWC (N,H,W,64): input of Conv3x3(F,1) in branch `u`
  const __half* r;                 // same for branch `convt_R1`
  const float* wu;                 // [9][64]
  const float* wr;                 // [9][64]
  int N, H, W;                     // computed rectangle at OUTPUT resolution (zero outside)
  int oy, ox;                      // canvas coordinates of its pixel (0,0)
  // kept region and seam blend of this tile, canvas coordinates (doCrop, imageProcess.py:164-170)
  int keep_y0, keep_y1, keep_x0, keep_x1;   // rows/cols [y0,y1) x [x0,x1) are written
  int ramp_y0, ramp_x0;            // first row / col of the seam (unclipped start of the kept region)
  int blend_y1, blend_x1;          // rows [ramp_y0, blend_y1) / cols [ramp_x0, blend_x1) are blended with the canvas
  float ramp[kMaxSeam];            // pad_sc weights, by value (the plan's ramp lives in host memory)
  __half* canvas;                  // planes x out_h x out_w (strided)
  int64_t plane_stride, row_stride;
};

// Conv3x3(F,1)(u) + Conv3x3(F,1)(r), rounded to fp16, seam-blended against what the canvas holds and
// stored.  One thread per output pixel.
__global__ void __launch_bounds__(128) head_blend_kernel(const HeadParams p)
{
  __shared__ float ws[2 * 9 * 64];
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) { ws[i] = p.wu[i]; ws[576 + i] = p.wr[i]; }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;   // within the computed rectangle
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (x >= p.W) return;
  const int cy = p.oy + y, cx = p.ox + x;
  if (cy < p.keep_y0 || cy >= p.keep_y1 || cx < p.keep_x0 || cx >= p.keep_x1) return;
  float head[2] = {0.f, 0.f};
#pragma unroll 1
  for (int b = 0; b < 2; ++b) {
    float acc = 0.f;
    const __half* src = b ? p.r : p.u;
    const float* wb = ws + b * 576;
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= p.H) continue;
#pragma unroll 1
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= p.W) continue;
        const uint4* sp = reinterpret_cast<const uint4*>(src + ((static_cast<size_t>(n) * p.H + yy) * p.W + xx) * 64);
        const float* wt = wb + (dy * 3 + dx) * 64;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 v = __ldg(sp + q);
          const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hv[e]);
            acc = fmaf(f.x, wt[q * 8 + e * 2], acc);
            acc = fmaf(f.y, wt[q * 8 + e * 2 + 1], acc);
          }
        }
      }
    }
    head[b] = h_round(acc);
  }
  float v = h_round(head[0] + head[1]);              // u + convt_R1(t): each head is an fp16 tensor, the add a third op
  __half* dst = p.canvas + n * p.plane_stride + cy * p.row_stride + cx;
  if (cy < p.blend_y1 || cx < p.blend_x1) {
    const float old = __half2float(*dst);
    // blend(): b = bx + w*(b - bx), each op rounded to the canvas dtype (imageProcess.py:129)
    if (cy < p.blend_y1) v = h_round(old + h_round(p.ramp[cy - p.ramp_y0] * h_round(v - old)));
    if (cx < p.blend_x1) v = h_round(old + h_round(p.ramp[cx - p.ramp_x0] * h_round(v - old)));
  }
  *dst = __float2half_rn(v);
}

// Direct 3x3 convolution with exactly the interface of conv3x3_tc_kernel (cross-check only).
// 8 threads per OUTPUT pixel (one per 8-channel group) reading the same swizzled weight images.
__global__ void __launch_bounds__(256) conv3x3_simt_kernel(const ConvParams p)
{
  const int nchunks = p.r * p.r;
  const int64_t total = static_cast<int64_t>(p.N) * p.H * p.W * nchunks * 8;
  const int Ho = p.H * p.r, Wo = p.W * p.r;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx & 7);
    int64_t rest = idx >> 3;
    const int chunk = static_cast<int>(rest % nchunks);
    rest /= nchunks;
    const int x = static_cast<int>(rest % p.W);
    const int y = static_cast<int>((rest / p.W) % p.H);
    const int n = static_cast<int>(rest / (static_cast<int64_t>(p.W) * p.H));
    float a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = 0.f;
    const uint8_t* img = p.w_img + static_cast<size_t>(chunk) * kChunkImgBytes;
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
      const __half* ip = p.in + ((static_cast<size_t>(n) * p.H + yy) * p.W + xx) * 64;
      for (int kc = 0; kc < 8; ++kc) {           // 8 input channels at a time
        const uint4 iv = __ldg(reinterpret_cast<const uint4*>(ip) + kc);
        const __half* ih = reinterpret_cast<const __half*>(&iv);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int row = g * 8 + c;                 // output channel within the chunk
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(img + tap * 8192 + row * 128 + ((kc ^ (row & 7)) << 4)));
          const __half* wh = reinterpret_cast<const __half*>(&wv);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[c] = fmaf(__half2float(ih[e]), __half2float(wh[e]), a[c]);
        }
      }
    }
    const int sy = chunk / p.r, sx = chunk - sy * p.r;
    const size_t ipix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
    const size_t opix = (static_cast<size_t>(n) * Ho + (y * p.r + sy)) * Wo + (x * p.r + sx);
    uint32_t w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float f[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ch = g * 8 + c * 2 + e;
        const float b = p.epi == EPI_BIAS_PRELU ? p.bias[chunk * 64 + ch] : 0.f;
        const float s = p.epi == EPI_SCALE_SKIP ? __half2float(p.skip[ipix * 64 + ch]) : 0.f;
        f[e] = epi_apply_rt(a[c * 2 + e], p.epi, p.param, b, s, p.bias_fused);
      }
      const __half2 hv = __floats2half2_rn(f[0], f[1]);
      w[c] = *reinterpret_cast<const uint32_t*>(&hv);
    }
    reinterpret_cast<uint4*>(p.out + opix * 64)[g] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- FRM, the feature recalibration module of MoeNet_lite2 (models.py:270-287; used by LB, MoeNet_lite2.py:15-19):
// t <- v * sigmoid(W1 relu(W0 mean_hw(v) + b0) + b1) + t, every op rounded to fp16 as in the reference's half model.
// The mean is over the whole reference tile, so the
// reduction is two deterministic passes (fixed chunking, fixed summation order: the result is reproducible).
constexpr int kFrmBlocks = 512;                    // partial sums per plane

// partial[n][blk][64] = sum of v[n, pixels of chunk blk, :]
__global__ void __launch_bounds__(256) frm_partial_kernel(const __half* v, float* partial, int64_t pixels_per_plane)
{
  __shared__ float red[32][65];
  const int n = blockIdx.y, blk = blockIdx.x;
  const int64_t chunk = (pixels_per_plane + kFrmBlocks - 1) / kFrmBlocks;
  const int64_t p0 = blk * chunk, p1 = min(pixels_per_plane, p0 + chunk);
  const int g = threadIdx.x & 7, lane_px = threadIdx.x >> 3;           // 8 channel groups x 32 pixel lanes
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4* base = reinterpret_cast<const uint4*>(v + static_cast<size_t>(n) * pixels_per_plane * 64);
  for (int64_t px = p0 + lane_px; px < p1; px += 32) {
    const uint4 q = __ldg(base + px * 8 + g);
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); a[2 * e] += f.x; a[2 * e + 1] += f.y; }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) red[lane_px][g * 8 + c] = a[c];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += red[l][threadIdx.x];
    partial[(static_cast<size_t>(n) * kFrmBlocks + blk) * 64 + threadIdx.x] = s;
  }
}

// gate[n][c] = sigmoid(b1[c] + sum_j w1[c][j] * relu(b0[j] + sum_k w0[j][k] * mean[n][k]));  frm = w0[3][64], b0[4], w1[64][4], b1[64]
__global__ void __launch_bounds__(64) frm_gate_kernel(const float* partial, const float* frm, float* gate, float inv_pixels, int bias_fused)
{
  __shared__ float mean[64];
  __shared__ float hid[4];
  const int n = blockIdx.x, c = threadIdx.x;
  float s = 0.f;
  for (int b = 0; b < kFrmBlocks; ++b) s += partial[(static_cast<size_t>(n) * kFrmBlocks + b) * 64 + c];
  mean[c] = h_round(s * inv_pixels);          // adaptive_avg_pool2d of a half tensor: fp32 accumulation, fp16 result
  __syncthreads();
  if (c < 3) {
    float h = 0.f;
    for (int k = 0; k < 64; ++k) h += frm[c * 64 + k] * mean[k];
    h = bias_fused ? h_round(h + frm[192 + c]) : h_round(h_round(h) + frm[192 + c]);     // conv_du.0, then (GPU) the bias add as its own op
    hid[c] = fmaxf(h, 0.f);                   // ReLU is exact
  }
  __syncthreads();
  float z = frm[196 + c * 4] * hid[0] + frm[196 + c * 4 + 1] * hid[1] + frm[196 + c * 4 + 2] * hid[2];                              // conv_du.2
  z = bias_fused ? h_round(z + frm[452 + c]) : h_round(h_round(z) + frm[452 + c]);
  gate[n * 64 + c] = h_round(1.f / (1.f + expf(-z)));                                                                               // Sigmoid
}

// t = round16(round16(v * gate) + t), NHWC, 8 channels per thread
__global__ void __launch_bounds__(256) frm_apply_kernel(const __half* v, __half* t, const float* gate, int64_t pixels_per_plane, int planes)
{
  const int64_t total = pixels_per_plane * planes * 8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i & 7);
    const int n = static_cast<int>((i >> 3) / pixels_per_plane);
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(v) + i);
    uint4 qt = reinterpret_cast<const uint4*>(t)[i];
    const __half2* hv = reinterpret_cast<const __half2*>(&qv);
    __half2* ht = reinterpret_cast<__half2*>(&qt);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + n * 64 + g * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + n * 64 + g * 8 + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fv = __half22float2(hv[e]), ft = __half22float2(ht[e]);
      // x * y (models.py:287) and + x (MoeNet_lite2.py:19) are two aten ops on half tensors: two roundings
      ht[e] = __floats2half2_rn(__fadd_rn(h_round(__fmul_rn(fv.x, gg[2 * e])), ft.x), __fadd_rn(h_round(__fmul_rn(fv.y, gg[2 * e + 1])), ft.y));
    }
    reinterpret_cast<uint4*>(t)[i] = qt;
  }
}

// y = s*y + (1-s)*x, every op rounded to fp16 (strengthOp, imageProcess.py:562)
__global__ void axpby_f16_kernel(__half* y, const __half* x, float s, float t, size_t count)
{
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float a = h_round(s * __half2float(y[i]));
    const float b = h_round(t * __half2float(x[i]));
    y[i] = __float2half_rn(a + b);
  }
}

// toTorch (imageProcess.py:259-263): interleaved integers -> planar fp16 in [0,1)
template <typename T>
__global__ void to_planar_kernel(const T* src, int h, int w, int c, int swap_rb, float inv, int exact255, __half* dst)
{
  const size_t total = static_cast<size_t>(h) * w;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    for (int ch = 0; ch < c; ++ch) {
      const int sc = (swap_rb && ch < 3) ? 2 - ch : ch;
      const float v = static_cast<float>(src[i * c + sc]);
      // to_tensor divides by 255 (a true division, not a multiply by the rounded reciprocal)
      dst[static_cast<size_t>(ch) * total + i] = __float2half_rn(exact255 ? __fdiv_rn(v, 255.f) : v * inv);
    }
  }
}

// toFloat + toOutput (imageProcess.py:238-257): planar fp16 -> x2^bits, clamp, truncate, interleave
template <typename T>
__global__ void to_output_kernel(const __half* src, int h, int w, int c, int swap_rb, float quant, T* dst)
{
  const size_t total = static_cast<size_t>(h) * w;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    for (int ch = 0; ch < c; ++ch) {
      const int sc = (swap_rb && ch < 3) ? 2 - ch : ch;
      float v = __half2float(src[static_cast<size_t>(sc) * total + i]) * quant;
      v = fminf(fmaxf(v, 0.f), quant - 1.f);
      dst[i * c + ch] = static_cast<T>(v);
    }
  }
}

}  // namespace moe
