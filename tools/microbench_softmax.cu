// How fast can the exponential phase of the attention softmax run in isolation?  Each warp processes 64 register-resident
// scores per "tile step" exactly like flash_attn_fwd_v4 (FFMA2 scale, MUFU / polynomial exp2, FADD2 sum, bf16 pack), with
// W warps per SM sub-partition.  Prints cycles per tile step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I actionmesh_b200/csrc -o tools/microbench_softmax tools/microbench_softmax.cu
#include <cstdio>
#include "ptx.cuh"
using namespace amb;

template <int EMU>
__global__ void __launch_bounds__(1024) k(float* out, float seed, long long* cyc, int iters) {
  float sc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) sc[i] = -(seed * (i + 1) * 0.01f + threadIdx.x * 1e-5f);
  uint32_t acc = 0;
  float row_sum = 0.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const float mb = 0.5f + it * 1e-3f;
    const uint64_t scale2 = pk2(0.127f, 0.127f), nmb2 = pk2(-mb, -mb);
    uint64_t psum2 = pk2(0.f, 0.f), psum2b = pk2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int t = 0; t < 16; t += 2) {
        float x0, x1, e0, e1;
        upk2(fma2(pk2(sc[c * 16 + t], sc[c * 16 + t + 1]), scale2, nmb2), x0, x1);
        if (EMU > 0 && ((t >> 1) % EMU) == EMU - 1) exp2_poly2(x0, x1, e0, e1);
        else { e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
        if ((t >> 1) & 1) psum2b = add2(psum2b, pk2(e0, e1));
        else psum2 = add2(psum2, pk2(e0, e1));
        pk[t >> 1] = pack_bf16(e0, e1);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc ^= pk[u];
    }
    float s0, s1;
    upk2(add2(psum2, psum2b), s0, s1);
    row_sum += s0 + s1;
#pragma unroll
    for (int i = 0; i < 64; i += 16) sc[i] += 1e-6f * (float)(acc & 1);  // keep the loop from being hoisted
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = row_sum + __uint_as_float(acc & 0xff);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int EMU>
void run(int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 200;
  k<EMU><<<148, threads>>>(out, 1.0f, cyc, iters);
  k<EMU><<<148, threads>>>(out, 1.0f, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("EMU=%d warps/SMSP=%d  cycles per 64-score step = %.0f   (%s)\n", EMU, threads / 128, (double)h / iters,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512}) { run<0>(th); run<4>(th); run<2>(th); }
  return 0;
}
