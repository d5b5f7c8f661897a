// Image preprocessing for the DinoV2 frame encoder on the GPU: byte/integer work, HBM-bound.
//
// Replaces the host-side HF BitImageProcessor call at actionmesh/model/image_encoder.py:48-51 (transformers < 5 as pinned by
// the reference's requirements.txt:10: PIL bicubic resize to shortest edge 256 -> centre crop 224 -> x 1/255 -> ImageNet
// mean/std -> channels first).  The resize is Pillow's two-pass separable convolution on uint8 (libImaging/Resample.c,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc): int32 coefficients with 22 fractional bits, accumulator seeded with
// 1 << 21, result (acc >> 22) clamped to [0, 255] and stored as uint8 BETWEEN the passes.  Both kernels do exactly that
// integer arithmetic, so the uint8 image is bit-identical to PIL's; the coefficient tables come from the host
// (actionmesh_b200/preprocess.py, float64 like Pillow's C doubles) already restricted to the cropped output window.
#include "common.cuh"
#include "../../include/actionmesh_b200.h"

namespace amb {

constexpr int kPrecisionBits = 32 - 8 - 2;

// horizontal pass: src (n, in_h, in_w, cin) u8 -> dst (n, n_rows, out_w, 3) u8 for the source rows [y0, y0 + n_rows)
__global__ void __launch_bounds__(256) resize_h_u8_kernel(const uint8_t* __restrict__ src, int in_h, int in_w, int cin, int y0,
                                                          int n_rows, const int32_t* __restrict__ bounds,
                                                          const int32_t* __restrict__ coeffs, int ksize, int out_w,
                                                          uint8_t* __restrict__ dst, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % out_w);
    const long long t = i / out_w;
    const int r = (int)(t % n_rows);
    const long long img = t / n_rows;
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int32_t* k = coeffs + (long long)xx * ksize;
    const uint8_t* p = src + ((img * in_h + (y0 + r)) * in_w + xmin) * cin;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < n; ++x) {
      const int w = k[x];
      a0 += (int)p[0] * w;
      a1 += (int)p[1] * w;
      a2 += (int)p[2] * w;
      p += cin;
    }
    uint8_t* d = dst + i * 3;
    d[0] = (uint8_t)min(max(a0 >> kPrecisionBits, 0), 255);
    d[1] = (uint8_t)min(max(a1 >> kPrecisionBits, 0), 255);
    d[2] = (uint8_t)min(max(a2 >> kPrecisionBits, 0), 255);
  }
}

// vertical pass + rescale + normalise + channels-first: src (n, n_rows, out_w, 3) u8 (row r == source row y0 + r)
// -> dst (n, 3, out_h, out_w) fp32 = (lut[u8] - mean[c]) / std[c]   (lut[v] = float(double(v) * rescale), host-built)
__global__ void __launch_bounds__(256) resize_v_normalize_kernel(const uint8_t* __restrict__ src, int n_rows, int y0, int out_w,
                                                                 const int32_t* __restrict__ bounds,
                                                                 const int32_t* __restrict__ coeffs, int ksize, int out_h,
                                                                 const float* __restrict__ lut, float m0, float m1, float m2,
                                                                 float s0, float s1, float s2, float* __restrict__ dst,
                                                                 uint8_t* __restrict__ dst_u8, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % out_w);
    const long long t = i / out_w;
    const int yy = (int)(t % out_h);
    const long long img = t / out_h;
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int32_t* k = coeffs + (long long)yy * ksize;
    const uint8_t* p = src + ((img * n_rows + (ymin - y0)) * out_w + xx) * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < n; ++y) {
      const int w = k[y];
      a0 += (int)p[0] * w;
      a1 += (int)p[1] * w;
      a2 += (int)p[2] * w;
      p += (long long)out_w * 3;
    }
    const int v0 = min(max(a0 >> kPrecisionBits, 0), 255), v1 = min(max(a1 >> kPrecisionBits, 0), 255),
              v2 = min(max(a2 >> kPrecisionBits, 0), 255);
    const long long plane = (long long)out_h * out_w;
    float* d = dst + img * 3 * plane + (long long)yy * out_w + xx;
    d[0] = __fdiv_rn(__fsub_rn(lut[v0], m0), s0);
    d[plane] = __fdiv_rn(__fsub_rn(lut[v1], m1), s1);
    d[2 * plane] = __fdiv_rn(__fsub_rn(lut[v2], m2), s2);
    if (dst_u8) {  // optional: the resized + cropped uint8 image (n, out_h, out_w, 3), for bit-exact parity checks
      uint8_t* u = dst_u8 + i * 3;
      u[0] = (uint8_t)v0; u[1] = (uint8_t)v1; u[2] = (uint8_t)v2;
    }
  }
}

static int grid_for_items(long long items, int block) {
  long long g = (items + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace amb

using namespace amb;

// ---- ImagePreprocessor (actionmesh/preprocessing/image_processor.py:26-146): RGBA frames -> composite on white -> crop to
// the foreground bounding box -> pad to a square with a margin.  Two kernels around a few host integers (the bounding boxes):
// alpha statistics, then composite + crop + pad straight to the uint8 image the reference hands on as PIL.

// per image: stats[0..3] = xmin, ymin, xmax, ymax of alpha > 0 (image_processor.py:57-64), stats[4] = count of alpha > 127
// (is_valid_alpha, :15-23).  One warp-aggregated atomic per row segment; stats are initialised by alpha_stats_init_kernel.
__global__ void alpha_stats_init_kernel(int32_t* stats, int n, int h, int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    stats[5 * i + 0] = w;
    stats[5 * i + 1] = h;
    stats[5 * i + 2] = -1;
    stats[5 * i + 3] = -1;
    stats[5 * i + 4] = 0;
  }
}
__global__ void __launch_bounds__(256) alpha_stats_kernel(const uint8_t* __restrict__ rgba, int h, int w, int32_t* stats) {
  const int img = blockIdx.z;
  const int y = blockIdx.y;
  const uint8_t* row = rgba + (((long long)img * h + y) * w) * 4;
  int xmin = w, xmax = -1, fg = 0;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
    const int a = row[4 * x + 3];
    if (a > 0) {
      xmin = min(xmin, x);
      xmax = max(xmax, x);
    }
    fg += a > 127;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    fg += __shfl_xor_sync(0xffffffffu, fg, o);
  }
  if ((threadIdx.x & 31) == 0) {
    int32_t* st = stats + 5 * img;
    if (xmax >= 0) {
      atomicMin(st + 0, xmin);
      atomicMax(st + 2, xmax);
      atomicMin(st + 1, y);
      atomicMax(st + 3, y);
    }
    if (fg) atomicAdd(st + 4, fg);
  }
}

// out (n, bh + 2 pad_y, bw + 2 pad_x, 3) u8: inside the box the composite rgb*a + 1*(1-a) in the reference's float32
// operation order (image_processor.py:44-52: (rgb*f32(1/255))*alpha + bg*(1-alpha), alpha = a*f32(1/255)), then *255 and
// truncation like `(img * 255).astype(np.uint8)` (:143-145); the padding is the background value 1.0 -> 255.
__global__ void __launch_bounds__(256) composite_crop_pad_kernel(const uint8_t* __restrict__ rgba, int h, int w, int bx, int by,
                                                                 int bw, int bh, int pad_x, int pad_y, uint8_t* __restrict__ out,
                                                                 long long total) {
  const int ow = bw + 2 * pad_x, oh = bh + 2 * pad_y;
  const float k = 1.0f / 255.0f;  // numpy multiplies the float32 array by float32(1.0 / 255.0)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    const long long t = i / ow;
    const int oy = (int)(t % oh);
    const long long img = t / oh;
    uint8_t r = 255, g = 255, b = 255;
    const int sx = ox - pad_x, sy = oy - pad_y;
    if (sx >= 0 && sx < bw && sy >= 0 && sy < bh) {
      const uint8_t* px = rgba + ((img * h + (by + sy)) * w + (bx + sx)) * 4;
      const float a = __fmul_rn((float)px[3], k);
      const float bgw = __fmul_rn(1.0f, __fsub_rn(1.0f, a));
      const float c0 = __fadd_rn(__fmul_rn(__fmul_rn((float)px[0], k), a), bgw);
      const float c1 = __fadd_rn(__fmul_rn(__fmul_rn((float)px[1], k), a), bgw);
      const float c2 = __fadd_rn(__fmul_rn(__fmul_rn((float)px[2], k), a), bgw);
      r = (uint8_t)(int)__fmul_rn(c0, 255.0f);
      g = (uint8_t)(int)__fmul_rn(c1, 255.0f);
      b = (uint8_t)(int)__fmul_rn(c2, 255.0f);
    }
    uint8_t* d = out + i * 3;
    d[0] = r;
    d[1] = g;
    d[2] = b;
  }
}

extern "C" {

int amb_alpha_stats(const uint8_t* rgba, int n_images, int height, int width, int32_t* stats, amb_stream_t stream) {
  AMB_CHECK_ARG(rgba && stats, "alpha_stats: null pointer");
  AMB_CHECK_ARG(height > 0 && width > 0 && height <= 65535, "alpha_stats: bad geometry %dx%d", height, width);
  if (n_images <= 0) return AMB_OK;
  AMB_CHECK_ARG(n_images <= 65535, "alpha_stats: at most 65535 images per call");
  cudaStream_t s = (cudaStream_t)stream;
  alpha_stats_init_kernel<<<(n_images + 127) / 128, 128, 0, s>>>(stats, n_images, height, width);
  AMB_CHECK_CUDA(cudaGetLastError());
  dim3 grid((width + 255) / 256 > 4 ? 4 : (width + 255) / 256, height, n_images);
  alpha_stats_kernel<<<grid, 256, 0, s>>>(rgba, height, width, stats);
  AMB_CHECK_CUDA(cudaGetLastError());
  return AMB_OK;
}

int amb_composite_crop_pad(const uint8_t* rgba, int n_images, int height, int width, int box_x, int box_y, int box_w, int box_h,
                           int pad_x, int pad_y, uint8_t* out, amb_stream_t stream) {
  AMB_CHECK_ARG(rgba && out, "composite_crop_pad: null pointer");
  AMB_CHECK_ARG(height > 0 && width > 0 && box_w > 0 && box_h > 0 && box_x >= 0 && box_y >= 0 && box_x + box_w <= width &&
                    box_y + box_h <= height && pad_x >= 0 && pad_y >= 0,
                "composite_crop_pad: box (%d,%d,%d,%d) / padding (%d,%d) do not fit the %dx%d image", box_x, box_y, box_w, box_h,
                pad_x, pad_y, height, width);
  if (n_images <= 0) return AMB_OK;
  const long long total = (long long)n_images * (box_h + 2 * pad_y) * (box_w + 2 * pad_x);
  composite_crop_pad_kernel<<<grid_for_items(total, 256), 256, 0, (cudaStream_t)stream>>>(rgba, height, width, box_x, box_y, box_w,
                                                                                         box_h, pad_x, pad_y, out, total);
  AMB_CHECK_CUDA(cudaGetLastError());
  return AMB_OK;
}

int amb_resize_h_u8(const uint8_t* src, int n_images, int in_h, int in_w, int channels_in, int y0, int n_rows,
                    const int32_t* bounds, const int32_t* coeffs, int ksize, int out_w, uint8_t* dst, amb_stream_t stream) {
  AMB_CHECK_ARG(src && bounds && coeffs && dst, "resize_h: null pointer");
  AMB_CHECK_ARG((channels_in == 3 || channels_in == 4) && in_h > 0 && in_w > 0 && out_w > 0 && ksize > 0,
                "resize_h: bad geometry in=%dx%dx%d out_w=%d ksize=%d", in_h, in_w, channels_in, out_w, ksize);
  AMB_CHECK_ARG(y0 >= 0 && n_rows > 0 && y0 + n_rows <= in_h, "resize_h: rows [%d, %d) outside the image height %d", y0,
                y0 + n_rows, in_h);
  if (n_images <= 0) return AMB_OK;
  const long long total = (long long)n_images * n_rows * out_w;
  resize_h_u8_kernel<<<grid_for_items(total, 256), 256, 0, (cudaStream_t)stream>>>(src, in_h, in_w, channels_in, y0, n_rows,
                                                                                   bounds, coeffs, ksize, out_w, dst, total);
  AMB_CHECK_CUDA(cudaGetLastError());
  return AMB_OK;
}

int amb_resize_v_normalize(const uint8_t* src, int n_images, int n_rows, int y0, int out_w, const int32_t* bounds,
                           const int32_t* coeffs, int ksize, int out_h, const float* lut256, const float* mean3_host,
                           const float* std3_host, float* dst, uint8_t* dst_u8, amb_stream_t stream) {
  AMB_CHECK_ARG(src && bounds && coeffs && lut256 && mean3_host && std3_host && dst, "resize_v_normalize: null pointer");
  AMB_CHECK_ARG(n_rows > 0 && out_w > 0 && out_h > 0 && ksize > 0 && y0 >= 0, "resize_v_normalize: bad geometry");
  if (n_images <= 0) return AMB_OK;
  const long long total = (long long)n_images * out_h * out_w;
  resize_v_normalize_kernel<<<grid_for_items(total, 256), 256, 0, (cudaStream_t)stream>>>(
      src, n_rows, y0, out_w, bounds, coeffs, ksize, out_h, lut256, mean3_host[0], mean3_host[1], mean3_host[2],
      std3_host[0], std3_host[1], std3_host[2], dst, dst_u8, total);
  AMB_CHECK_CUDA(cudaGetLastError());
  return AMB_OK;
}

}  // extern "C"
