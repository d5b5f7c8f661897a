// How fast can the exponential phase of the attention softmax run in isolation?  Each warp processes 64 register-resident
// scores per "tile step" like flash_attn_fwd_v4 (FFMA2 scale, MUFU / polynomial exp2, FADD2 sum, bf16 pack), W warps per
// SM sub-partition.  Variants isolate which part of the instruction stream sets the latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I actionmesh_b200/csrc -o tools/microbench_softmax tools/microbench_softmax.cu
#include <cstdio>
#include "ptx.cuh"
using namespace amb;

// VAR: 0 full packed, 1 no row-sum, 2 no bf16 pack, 3 MUFU only, 4 FFMA2 only, 5 scalar (unpacked) math,
//      6 phase-separated (all scales, all exps, all sums, all packs), 7 scalar phase-separated
template <int VAR, int EMU>
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
    const float sl = 0.127f;
    const uint64_t scale2 = pk2(sl, sl), nmb2 = pk2(-mb, -mb);
    uint64_t psum2 = pk2(0.f, 0.f), psum2b = pk2(0.f, 0.f);
    float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
    if (VAR == 6 || VAR == 7) {
      float x[64], e[64];
#pragma unroll
      for (int t = 0; t < 64; t += 2) {
        if (VAR == 6) upk2(fma2(pk2(sc[t], sc[t + 1]), scale2, nmb2), x[t], x[t + 1]);
        else { x[t] = fmaf(sc[t], sl, -mb); x[t + 1] = fmaf(sc[t + 1], sl, -mb); }
      }
#pragma unroll
      for (int t = 0; t < 64; ++t) e[t] = ex2_approx(x[t]);
#pragma unroll
      for (int t = 0; t < 64; t += 4) {
        if (VAR == 6) { psum2 = add2(psum2, pk2(e[t], e[t + 1])); psum2b = add2(psum2b, pk2(e[t + 2], e[t + 3])); }
        else { ps0 += e[t]; ps1 += e[t + 1]; ps2 += e[t + 2]; ps3 += e[t + 3]; }
      }
#pragma unroll
      for (int t = 0; t < 64; t += 2) acc ^= pack_bf16(e[t], e[t + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          float x0, x1, e0, e1;
          if (VAR == 3) { x0 = sc[c * 16 + t]; x1 = sc[c * 16 + t + 1]; }
          else if (VAR == 5) { x0 = fmaf(sc[c * 16 + t], sl, -mb); x1 = fmaf(sc[c * 16 + t + 1], sl, -mb); }
          else upk2(fma2(pk2(sc[c * 16 + t], sc[c * 16 + t + 1]), scale2, nmb2), x0, x1);
          if (VAR == 4) { e0 = x0; e1 = x1; }
          else if (EMU > 0 && ((t >> 1) % EMU) == EMU - 1) exp2_poly2(x0, x1, e0, e1);
          else { e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
          if (VAR == 5) { ps0 += e0; ps1 += e1; }
          else if (VAR != 1 && VAR != 3 && VAR != 4) {
            if ((t >> 1) & 1) psum2b = add2(psum2b, pk2(e0, e1));
            else psum2 = add2(psum2, pk2(e0, e1));
          }
          if (VAR == 2 || VAR == 3 || VAR == 4) pk[t >> 1] = __float_as_uint(e0) ^ __float_as_uint(e1);
          else pk[t >> 1] = pack_bf16(e0, e1);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= pk[u];
      }
    }
    float s0, s1;
    upk2(add2(psum2, psum2b), s0, s1);
    row_sum += s0 + s1 + ps0 + ps1 + ps2 + ps3;
#pragma unroll
    for (int i = 0; i < 64; i += 16) sc[i] += 1e-6f * (float)(acc & 1);  // keep the loop from being hoisted
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = row_sum + __uint_as_float(acc & 0xff);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int VAR, int EMU>
void run(const char* name, int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 200;
  k<VAR, EMU><<<148, threads>>>(out, 1.0f, cyc, iters);
  k<VAR, EMU><<<148, threads>>>(out, 1.0f, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s EMU=%d warps/SMSP=%d  cycles per 64-score step = %6.0f   (%s)\n", name, EMU, threads / 128, (double)h / iters,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0, 0>("full packed", th);
    run<0, 4>("full packed", th);
    run<1, 0>("no row sum", th);
    run<2, 0>("no bf16 pack", th);
    run<3, 0>("MUFU only", th);
    run<4, 0>("FFMA2 only", th);
    run<5, 0>("scalar math", th);
    run<6, 0>("phase-separated packed", th);
    run<7, 0>("phase-separated scalar", th);
  }
  return 0;
}
