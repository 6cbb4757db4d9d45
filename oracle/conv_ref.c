/* TEST INFRASTRUCTURE — CPU oracle, not product code.  Nothing under moephoto_b200/ may call this.
 *
 * Plain-C restatement of the three tensor primitives the reference's SR/DN path uses
 * (they live in PyTorch, an un-vendored dependency: requirements.txt:3 `torch>=1.10`):
 *   - conv2d 3x3, stride 1, zero padding 1, optional bias   (reference call sites models.py:35, :29-30)
 *   - PReLU with one scalar slope                           (models.py:30, :78, :114)
 *   - PixelShuffle(r)                                       (models.py:30)
 * Published definitions (torch.nn docs): out[n,co,y,x] = b[co] + sum_{ci,ky,kx} w[co,ci,ky,kx] *
 * in[n,ci,y+ky-1,x+kx-1] (zero outside); prelu(v)= v>=0 ? v : a*v;
 * shuffle: out[n,c,y*r+i,x*r+j] = in[n,c*r*r+i*r+j,y,x].
 * fp32 accumulation in (ci,ky,kx) order.  Layout NCHW, contiguous float32.
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC (oracle/build.py).
 */
#include <stddef.h>
#include <string.h>

void oracle_conv3x3(const float* in, const float* w, const float* bias, float* out,
                    int n, int cin, int cout, int h, int wd)
{
  const size_t plane = (size_t)h * wd;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < n; ++b) {
    for (int co = 0; co < cout; ++co) {
      float* o = out + ((size_t)b * cout + co) * plane;
      const float b0 = bias ? bias[co] : 0.0f;
      for (size_t i = 0; i < plane; ++i) o[i] = b0;
      for (int ci = 0; ci < cin; ++ci) {
        const float* src = in + ((size_t)b * cin + ci) * plane;
        const float* k = w + ((size_t)co * cin + ci) * 9;
        for (int ky = 0; ky < 3; ++ky) {
          for (int kx = 0; kx < 3; ++kx) {
            const float kv = k[ky * 3 + kx];
            const int dy = ky - 1, dx = kx - 1;
            const int y0 = dy < 0 ? 1 : 0, y1 = dy > 0 ? h - 1 : h;
            const int x0 = dx < 0 ? 1 : 0, x1 = dx > 0 ? wd - 1 : wd;
            for (int y = y0; y < y1; ++y) {
              float* orow = o + (size_t)y * wd;
              const float* irow = src + (size_t)(y + dy) * wd + dx;
              for (int x = x0; x < x1; ++x) orow[x] += kv * irow[x];
            }
          }
        }
      }
    }
  }
}

void oracle_prelu(float* x, size_t count, float slope)
{
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < count; ++i) x[i] = x[i] >= 0.0f ? x[i] : slope * x[i];
}

void oracle_pixel_shuffle(const float* in, float* out, int n, int c, int r, int h, int wd)
{
  const int oh = h * r, ow = wd * r;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < n; ++b)
    for (int ch = 0; ch < c; ++ch)
      for (int i = 0; i < r; ++i)
        for (int j = 0; j < r; ++j) {
          const float* src = in + (((size_t)b * c + ch) * r * r + (size_t)i * r + j) * h * wd;
          float* dst = out + ((size_t)b * c + ch) * oh * ow;
          for (int y = 0; y < h; ++y)
            for (int x = 0; x < wd; ++x)
              dst[(size_t)(y * r + i) * ow + (x * r + j)] = src[(size_t)y * wd + x];
        }
}
