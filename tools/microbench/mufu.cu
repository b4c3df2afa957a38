// MUFU.EX2 issue-rate microbenchmark: W warps per SM, each with 8 independent ex2 chains (+ optional FFMA
// per MUFU).  Prints cycles per warp-level MUFU per SMSP.  nvcc -arch=sm_100a -o mufu mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int FMA_PER>
__global__ void k(float* out, int iters, long long* cyc) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int f = 0; f < FMA_PER; ++f) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc) : "f"(1.0001f), "f"(x[(i + 4) & 7]));
    }
  }
  long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int fma = 0; fma <= 2; ++fma)
    for (int warps : {1, 2, 4, 8, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (fma == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        else if (fma == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        else k<3><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0;
      for (int i = 0; i < 148; ++i) c += h[i];
      c /= 148;
      const double mufu_per_smsp = double(iters) * 8 * warps / 4.0;  // warp-level MUFU instr per SMSP (warps>=4)
      printf("fma_per_mufu=%d warps/SM=%2d: %.0f cycles, %.2f cycles per warp-MUFU per warp, %.2f per SMSP-slot\n",
             fma == 2 ? 3 : fma, warps, c, c / (iters * 8.0), warps >= 4 ? c / mufu_per_smsp : c / (iters * 8.0));
    }
  return 0;
}
