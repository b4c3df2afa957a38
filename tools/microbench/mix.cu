// Which pipe does F2FP.BF16.PACK_AB use?  One warp per SMSP: 8 MUFU.EX2 [+ 4 F2FP] [+ 8 FFMA + 8 FADD] per iteration.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i), y[i] = 0.5f + i;
  unsigned acc = 0;
  float fs = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE != 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y[i]), "f"(y[i + 1]));
        acc ^= r;
      }
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(1.0001f), "f"(0.1f));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(fs) : "f"(y[i]));
      }
    }
  }
  long long t1 = clock64();
  float s = __uint_as_float(acc) + fs;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[] = {"8 MUFU", "4 F2FP", "8 MUFU + 4 F2FP", "8 MUFU + 4 F2FP + 8 FFMA + 8 FADD(chain)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0;
      for (int i = 0; i < 148; ++i) c += h[i];
      printf("%-45s warps/SM=%d: %.1f cycles per iteration\n", names[mode], warps, c / 148 / iters);
    }
  return 0;
}
