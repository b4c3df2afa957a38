// tcgen05.mma issue / execution rate at the FMHA shapes (M = 128, bf16, cta_group::1), one issuing warp:
// cycles per MMA for dependent chains (one accumulator) and for interleaved independent accumulators,
// SS (A from smem) and TS (A from TMEM) forms.  Operands are whatever the smem holds (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iln3diff_b200/csrc -Iinclude \
//        tools/microbench/umma_rate.cu -o tools/microbench/umma_rate
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace ln3;

struct Result { long long issue, total; };
__device__ Result g_res[32];

// pattern: 0 = N=64 chain (1 accumulator)   1 = N=64, 2 accumulators alternating   2 = N=96 chain
//          3 = N=128 chain                  4 = N=64 TS chain                       5 = N=64 TS, 2 accumulators
//          6 = FMHA visit: 6 x (N=64 -> O) interleaved with 4 x (N=96 -> S)         7 = same, not interleaved
//          8 = N=96, 3 accumulators round robin   9 = N=64 SS, 3 accumulators round robin
//          10 = N=256 chain  11 = N=192 chain
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;            // 128 x 64 bf16, 128B swizzle (16 KB)
  uint8_t* sB = smem + 16384;    // up to 256 x 64 (32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (warp == 0) {
    const uint64_t dA = make_smem_desc_sw128(smem_u32(sA), 0, 1024);
    const uint64_t dB = make_smem_desc_sw128(smem_u32(sB), 0, 1024);
    const uint64_t dBmn = make_smem_desc_sw128(smem_u32(sB), 1024, 1024);
    constexpr uint32_t i64 = make_idesc_bf16(128, 64, 0, 1), i96 = make_idesc_bf16(128, 96, 0, 0),
                       i128 = make_idesc_bf16(128, 128, 0, 0), i256 = make_idesc_bf16(128, 256, 0, 0),
                       i192 = make_idesc_bf16(128, 192, 0, 0);
    int phase = 0;
    for (int pat = 0; pat < 12; ++pat) {
      __syncwarp();
      long long t0 = clock64(), t1 = 0;
      int n = 0;
      if (elect_one_sync()) {
        for (int r = 0; r < reps; ++r) {
          switch (pat) {
            case 0: for (int k = 0; k < 8; ++k) umma_f16_ss(tm + 256, dA + (k & 3) * 2, dBmn + k * 128, i64, 1); n += 8; break;
            case 1: for (int k = 0; k < 8; ++k) umma_f16_ss(tm + 256 + (k & 1) * 64, dA + (k & 3) * 2, dBmn + k * 128, i64, 1); n += 8; break;
            case 2: for (int k = 0; k < 8; ++k) umma_f16_ss(tm, dA + (k & 3) * 2, dB + (k & 3) * 2, i96, 1); n += 8; break;
            case 3: for (int k = 0; k < 8; ++k) umma_f16_ss(tm, dA + (k & 3) * 2, dB + (k & 3) * 2, i128, 1); n += 8; break;
            case 4: for (int k = 0; k < 8; ++k) umma_f16_ts(tm + 256, tm + 384 + k * 8, dBmn + k * 128, i64, 1); n += 8; break;
            case 5: for (int k = 0; k < 8; ++k) umma_f16_ts(tm + 256 + (k & 1) * 64, tm + 384 + k * 8, dBmn + k * 128, i64, 1); n += 8; break;
            case 6:
              for (int k = 0; k < 6; ++k) {
                umma_f16_ss(tm + 288, dA + (k & 3) * 2, dBmn + k * 128, i64, 1);
                if (k < 4) umma_f16_ss(tm, dA + k * 2, dB + k * 2, i96, 1);
              }
              n += 10; break;
            case 7:
              for (int k = 0; k < 6; ++k) umma_f16_ss(tm + 288, dA + (k & 3) * 2, dBmn + k * 128, i64, 1);
              for (int k = 0; k < 4; ++k) umma_f16_ss(tm, dA + k * 2, dB + k * 2, i96, 1);
              n += 10; break;
            case 8: for (int k = 0; k < 12; ++k) umma_f16_ss(tm + (k % 3) * 96, dA + (k & 3) * 2, dB + (k & 3) * 2, i96, 1); n += 12; break;
            case 9: for (int k = 0; k < 12; ++k) umma_f16_ss(tm + 288 + (k % 3) * 64, dA + (k & 3) * 2, dBmn + (k & 7) * 128, i64, 1); n += 12; break;
            case 10: for (int k = 0; k < 8; ++k) umma_f16_ss(tm, dA + (k & 3) * 2, dB + (k & 3) * 2, i256, 1); n += 8; break;
            case 11: for (int k = 0; k < 8; ++k) umma_f16_ss(tm, dA + (k & 3) * 2, dB + (k & 3) * 2, i192, 1); n += 8; break;
          }
        }
        umma_commit(bar);
        t1 = clock64();
      }
      __syncwarp();
      mbar_wait(bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      n = __shfl_sync(0xffffffffu, n, 0);
      t1 = __reduce_max_sync(0xffffffffu, (unsigned)(t1 > 0 ? t1 - t0 : 0));
      if ((threadIdx.x & 31) == 0) { g_res[pat].issue = t1 * 1000 / (reps * (pat == 6 || pat == 7 ? 10 : (pat == 8 || pat == 9 ? 12 : 8))); g_res[pat].total = (t2 - t0) * 1000 / (reps * (pat == 6 || pat == 7 ? 10 : (pat == 8 || pat == 9 ? 12 : 8))); }
      tc_fence_after();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  const int smem = 1024 + 16384 + 32768 + 64;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[12] = {"N=64 SS chain", "N=64 SS 2 acc", "N=96 SS chain", "N=128 SS chain", "N=64 TS chain", "N=64 TS 2 acc",
                           "visit interleaved (6xN64 + 4xN96)", "visit sequential", "N=96 SS 3 acc", "N=64 SS 3 acc", "N=256 SS chain", "N=192 SS chain"};
  for (int reps : {1, 8, 64}) {
    umma_rate_kernel<<<1, 128, smem>>>(reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    Result r[32];
    cudaMemcpyFromSymbol(r, g_res, sizeof(r));
    printf("reps=%d (cycles per MMA: issue-side / until complete)\n", reps);
    for (int i = 0; i < 12; ++i) printf("  %-36s %7.1f %7.1f\n", names[i], r[i].issue / 1000.0, r[i].total / 1000.0);
  }
  return 0;
}
