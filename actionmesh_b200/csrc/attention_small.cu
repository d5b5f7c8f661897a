// fp32 attention for short sequences (DinoV2: 257 tokens, 16 heads of 64):  O = softmax(scale · Q Kᵀ) V, everything fp32.
// Replaces the F.scaled_dot_product_attention inside HF Dinov2SelfAttention (transformers/models/dinov2/modeling_dinov2.py),
// which the reference runs in fp32, outside autocast (actionmesh/pipeline.py:664-667, model/image_encoder.py:38-55).
//
// Why CUDA cores: the whole encoder is 2.5 TFLOP per clip, its attention 0.1 TFLOP — 0.02 % of a 30-step denoise — and
// the reference's precision here is fp32.  A tensor-core path would need three-way split operands for Q, K, P and V; the
// fp32 FMA pipe does the 0.1 TFLOP in a few milliseconds with nothing to split.
//
// One CTA per (head, frame): K (padded rows: conflict-free column reads) and V of the head live in shared memory.  A warp
// takes four query rows at a time: lane l scores keys l, l+32, ... (the q values are warp-wide broadcasts, reused by the
// four rows), the row maxima / sums are warp reductions, the un-normalised probabilities go through a per-warp smem buffer
// and lane l accumulates output columns 2l, 2l+1 (one 8-byte V read per key for all four rows).
#include "common.cuh"
#include "../../include/actionmesh_b200.h"

namespace amb {

constexpr int AS_D = 64;        // head_dim
constexpr int AS_WARPS = 8;
constexpr int AS_ROWS = 4;      // query rows a warp processes together
constexpr int AS_MAX_S = 320;   // keys per (head, frame): 10 per lane
constexpr int AS_KI = AS_MAX_S / 32;

__host__ __device__ inline int as_k_floats(int S) { return (S * (AS_D + 1) + 3) / 4 * 4; }  // V behind it stays 16-byte aligned
__host__ __device__ inline int as_smem_bytes(int S) {
  const int spad = (S + 31) / 32 * 32;
  return (as_k_floats(S) + S * AS_D + AS_WARPS * AS_ROWS * AS_D + AS_WARPS * AS_ROWS * spad) * 4;
}

__global__ void __launch_bounds__(AS_WARPS * 32)
attn_small_f32_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, long long ld,
                      int S, float scale, float* __restrict__ o, long long ldo) {
  extern __shared__ __align__(16) float as_smem[];
  const int spad = (S + 31) / 32 * 32;
  float* Ks = as_smem;                          // [S][65]
  float* Vs = Ks + as_k_floats(S);              // [S][64]
  float* Qs = Vs + S * AS_D;                    // [warps][rows][64]
  float* Ps = Qs + AS_WARPS * AS_ROWS * AS_D;   // [warps][rows][spad]
  const int head = blockIdx.x, frame = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)frame * S;
  const int col0 = head * AS_D;

  for (int i = threadIdx.x; i < S * AS_D; i += AS_WARPS * 32) {
    const int r = i >> 6, d = i & 63;
    Ks[r * (AS_D + 1) + d] = k[(row0 + r) * ld + col0 + d];
    Vs[i] = v[(row0 + r) * ld + col0 + d];
  }
  __syncthreads();

  float* qs = Qs + warp * AS_ROWS * AS_D;
  float* ps = Ps + warp * AS_ROWS * spad;
  const int nki = spad / 32;
  for (int r0 = warp * AS_ROWS; r0 < S; r0 += AS_WARPS * AS_ROWS) {
#pragma unroll
    for (int rr = 0; rr < AS_ROWS; ++rr) {
      const int r = r0 + rr;
      qs[rr * AS_D + lane] = r < S ? q[(row0 + r) * ld + col0 + lane] : 0.f;
      qs[rr * AS_D + lane + 32] = r < S ? q[(row0 + r) * ld + col0 + lane + 32] : 0.f;
    }
    __syncwarp();
    float acc[AS_ROWS][AS_KI];
#pragma unroll
    for (int rr = 0; rr < AS_ROWS; ++rr)
#pragma unroll
      for (int i = 0; i < AS_KI; ++i) acc[rr][i] = 0.f;
    int krow[AS_KI];
#pragma unroll
    for (int i = 0; i < AS_KI; ++i) {
      const int key = lane + 32 * i;
      krow[i] = (key < S ? key : S - 1) * (AS_D + 1);  // keys past the end read a valid row; their scores are masked below
    }
#pragma unroll 4
    for (int d = 0; d < AS_D; ++d) {
      float qv[AS_ROWS];
#pragma unroll
      for (int rr = 0; rr < AS_ROWS; ++rr) qv[rr] = qs[rr * AS_D + d];
#pragma unroll
      for (int i = 0; i < AS_KI; ++i) {
        if (i < nki) {
          const float kv = Ks[krow[i] + d];
#pragma unroll
          for (int rr = 0; rr < AS_ROWS; ++rr) acc[rr][i] = fmaf(qv[rr], kv, acc[rr][i]);
        }
      }
    }
    float inv[AS_ROWS];
#pragma unroll
    for (int rr = 0; rr < AS_ROWS; ++rr) {
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < AS_KI; ++i)
        if (i < nki && lane + 32 * i < S) m = fmaxf(m, acc[rr][i]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < AS_KI; ++i) {
        if (i < nki) {
          const float pv = lane + 32 * i < S ? expf((acc[rr][i] - m) * scale) : 0.f;
          ps[rr * spad + lane + 32 * i] = pv;
          sum += pv;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      inv[rr] = 1.0f / sum;
    }
    __syncwarp();
    float ox[AS_ROWS], oy[AS_ROWS];
#pragma unroll
    for (int rr = 0; rr < AS_ROWS; ++rr) ox[rr] = oy[rr] = 0.f;
    for (int key = 0; key < S; ++key) {
      const float2 vv = *reinterpret_cast<const float2*>(Vs + key * AS_D + 2 * lane);
#pragma unroll
      for (int rr = 0; rr < AS_ROWS; ++rr) {
        const float pv = ps[rr * spad + key];
        ox[rr] = fmaf(pv, vv.x, ox[rr]);
        oy[rr] = fmaf(pv, vv.y, oy[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < AS_ROWS; ++rr) {
      const int r = r0 + rr;
      if (r < S)
        *reinterpret_cast<float2*>(o + (row0 + r) * ldo + col0 + 2 * lane) = make_float2(ox[rr] * inv[rr], oy[rr] * inv[rr]);
    }
    __syncwarp();  // the next pass overwrites qs / ps
  }
}

}  // namespace amb

using namespace amb;

extern "C" int amb_attn_small_f32(const float* q, const float* k, const float* v, int64_t ld, int frames, int seq, int heads,
                                  float scale, float* out, int64_t ldo, amb_stream_t stream) {
  AMB_CHECK_ARG(q && k && v && out, "attn_small_f32: null pointer");
  AMB_CHECK_ARG(seq >= 1 && seq <= AS_MAX_S, "attn_small_f32: seq must be in [1, 320]");
  AMB_CHECK_ARG(heads >= 1 && frames >= 0, "attn_small_f32: bad heads / frames");
  AMB_CHECK_ARG(ld >= (int64_t)heads * AS_D && ldo >= (int64_t)heads * AS_D && ld % 2 == 0 && ldo % 2 == 0,
                "attn_small_f32: row strides must cover heads * 64 columns and be even");
  AMB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 7) == 0, "attn_small_f32: out must be 8-byte aligned");
  if (frames == 0) return AMB_OK;
  const int smem = as_smem_bytes(seq);
  {
    int r = ensure_smem_optin(attn_small_f32_kernel, as_smem_bytes(AS_MAX_S));  // opt in once, for the longest sequence
    if (r) return r;
  }
  attn_small_f32_kernel<<<dim3(heads, frames), AS_WARPS * 32, smem, (cudaStream_t)stream>>>(q, k, v, ld, seq, scale, out, ldo);
  AMB_CHECK_CUDA(cudaGetLastError());
  return AMB_OK;
}
