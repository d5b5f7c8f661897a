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

extern "C" {

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
