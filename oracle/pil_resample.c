/* TEST INFRASTRUCTURE ONLY — plain-C restatement of Pillow's 8-bit bicubic resample (libImaging/Resample.c:
 * precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc) for the
 * image-preprocessing row (actionmesh/model/image_encoder.py:48-51 -> HF BitImageProcessor -> PIL.Image.resize).
 * Pillow is a third-party dependency of the reference (not under /root/reference); this file restates its published
 * algorithm and tests/test_preprocess_cpu.py pins it bit-exactly against the installed Pillow.  Built by
 * __graft_entry__.build() into oracle/_build/libpil_resample.so; only tests load it.
 *
 *   int amb_oracle_resize_u8(const uint8_t* src, int h, int w, int c, int out_h, int out_w, uint8_t* dst)
 *     src (h, w, c) interleaved, dst (out_h, out_w, c); horizontal pass first, uint8 between the passes; a pass whose size
 *     does not change is skipped.  Returns 0, or -1 on allocation failure.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

/* bounds[2*i] = first source index, bounds[2*i+1] = taps; kk[i*ksize + t] = fixed-point weight */
static int coeffs(int in_size, int out_size, int** bounds_out, int32_t** kk_out, int* ksize_out) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  int* bounds = (int*)malloc(sizeof(int) * 2 * out_size);
  int32_t* kk = (int32_t*)calloc((size_t)out_size * ksize, sizeof(int32_t));
  double* k = (double*)malloc(sizeof(double) * ksize);
  if (!bounds || !kk || !k) { free(bounds); free(kk); free(k); return -1; }
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      k[x] = bicubic((x + xmin - center + 0.5) * ss);
      ww += k[x];
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << PRECISION_BITS);
      kk[(size_t)xx * ksize + x] = (int32_t)(k[x] < 0 ? -0.5 + v : 0.5 + v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  free(k);
  *bounds_out = bounds; *kk_out = kk; *ksize_out = ksize;
  return 0;
}

static uint8_t clip8(int32_t acc) {
  const int32_t v = acc >> PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

int amb_oracle_resize_u8(const uint8_t* src, int h, int w, int c, int out_h, int out_w, uint8_t* dst) {
  const uint8_t* cur = src;
  uint8_t* mid = NULL;
  if (w != out_w) {  /* horizontal pass: (h, w, c) -> (h, out_w, c) */
    int *b; int32_t* kk; int ks;
    if (coeffs(w, out_w, &b, &kk, &ks)) return -1;
    mid = (uint8_t*)malloc((size_t)h * out_w * c);
    if (!mid) { free(b); free(kk); return -1; }
    for (int y = 0; y < h; ++y)
      for (int xx = 0; xx < out_w; ++xx)
        for (int ch = 0; ch < c; ++ch) {
          int32_t acc = 1 << (PRECISION_BITS - 1);
          for (int x = 0; x < b[2 * xx + 1]; ++x)
            acc += (int32_t)cur[((size_t)y * w + b[2 * xx] + x) * c + ch] * kk[(size_t)xx * ks + x];
          mid[((size_t)y * out_w + xx) * c + ch] = clip8(acc);
        }
    free(b); free(kk);
    cur = mid;
  }
  if (h != out_h) {  /* vertical pass: (h, out_w, c) -> (out_h, out_w, c) */
    int *b; int32_t* kk; int ks;
    if (coeffs(h, out_h, &b, &kk, &ks)) { free(mid); return -1; }
    for (int yy = 0; yy < out_h; ++yy)
      for (int i = 0; i < out_w * c; ++i) {
        int32_t acc = 1 << (PRECISION_BITS - 1);
        for (int y = 0; y < b[2 * yy + 1]; ++y) acc += (int32_t)cur[((size_t)(b[2 * yy] + y)) * out_w * c + i] * kk[(size_t)yy * ks + y];
        dst[(size_t)yy * out_w * c + i] = clip8(acc);
      }
    free(b); free(kk);
  } else {
    memcpy(dst, cur, (size_t)out_h * out_w * c);
  }
  free(mid);
  return 0;
}
