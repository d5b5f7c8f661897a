// Throughput microbenchmarks that decide the softmax design of the attention kernel (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu && ./tools/microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float seed, long long* cyc) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (i + 1) * 1e-3f + threadIdx.x * 1e-6f; u[i] = __float_as_uint(a[i]); }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // MUFU.EX2 f32
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {  // MUFU.EX2 bf16x2
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      } else if (MODE == 2) {  // FFMA
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      } else if (MODE == 3) {  // FFMA2 (packed f32x2): 4 pairs
        if (i < 4) {
          unsigned long long p;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
          asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p));
        }
      } else if (MODE == 4) {  // cvt.rn.bf16x2.f32
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        a[i] += __uint_as_float(u[i] & 0x3f800000);
      } else if (MODE == 5) {  // FMNMX 3-input
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      } else if (MODE == 6) {  // emulated exp2: magic-round, 3 fma poly, exponent add (scalar)
        float x = a[i];
        float t = x + 12582912.0f;
        float n = t - 12582912.0f;
        float f = x - n;
        float p = fmaf(f, 0.0555041086f, 0.2402265069f);
        p = fmaf(p, f, 0.6931471805f);
        p = fmaf(p, f, 1.0f);
        a[i] = __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23)) * 1e-3f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<MODE><<<148, threads>>>(out, 1.0f, cyc);
  k<MODE><<<148, threads>>>(out, 1.0f, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  int warps_per_smsp = threads / 32 / 4;
  int n_ops = (MODE == 3) ? 4 : 8;
  double per = (double)h / (ITERS * n_ops * (warps_per_smsp > 0 ? warps_per_smsp : 1));
  printf("%-28s threads=%4d cycles=%9lld  cyc/warp-instr/SMSP=%.2f  %s\n", name, threads, h, per, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("ex2.f32", th);
    run<1>("ex2.bf16x2", th);
    run<2>("fma.f32", th);
    run<3>("fma.f32x2 (per instr)", th);
    run<4>("cvt.bf16x2+add", th);
    run<5>("max3.f32", th);
    run<6>("emulated exp2 (7 ops)", th);
  }
  return 0;
}
