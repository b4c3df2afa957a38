// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// Replaces the cuBLAS/cuBLASLt launches behind every nn.Linear of the reference DiT blocks
// (dit/dit_models_xformers.py:231-323 DiTBlock/TextCondDiTBlock, vit/vision_transformer.py:106-124
// qkv/proj, ldm/modules/attention.py:245-307 to_q/to_k/to_v/to_out, xformers FusedMLP) and fuses
// the elementwise tail that follows each of them in the reference (bias, GELU / SiLU, the
// adaLN-Zero `x + gate * f(.)` residual update) into the TMEM epilogue.
//
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor, 128B swizzle, kStages-deep ring)
//   warp 1        : TMEM allocator; lane 0 issues tcgen05.mma (UMMA 128 x BN x 16, fp32 in TMEM)
//   warps 2..5    : epilogue: tcgen05.ld -> registers -> bias/act/gate/residual -> global
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
#include "common.cuh"
#include "ln3_internal.h"

namespace ln3 {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 bytes = one 128B-swizzle row
static constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int act;        // LN3_ACT_*
  int out_kind;   // LN3_OUT_*
  const float* bias;       // [N] or null
  void* out;               // bf16 [M,ldo] or f32 [M,ldo] (for RESID: f32 residual, updated in place)
  long long ldo;           // leading dim of out, elements
  __nv_bfloat16* out2;     // optional bf16 copy of the updated residual (RESID only), ld = ldo2
  long long ldo2;
  const float* gate;       // RESID: gate[(m / gate_rows) * gate_ld + n]; null -> 1
  int gate_rows;
  long long gate_ld;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Tile order: consecutive CTAs walk M first inside a group of N panels so that the
  // concurrently resident tiles share W panels (L2 reuse); A panels stream.
  auto tile_coords = [&](int t, int& tm, int& tn) {
    tm = t % tiles_m;
    tn = t / tiles_m;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int tm, tn;
        tile_coords(t, tm, tn);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BK, tm * BM);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BK, tn * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // Epilogue warps 2..5 own TMEM lanes [32*(warp%4), +32).
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn;
      tile_coords(t, tm, tn);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = tm * BM + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      const float* gate_row = nullptr;
      if (p.out_kind == LN3_OUT_RESID_F32 && p.gate != nullptr && row_ok)
        gate_row = p.gate + static_cast<long long>(m / p.gate_rows) * p.gate_ld;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c, v);
        tmem_ld_wait();
        const int n0 = tn * BN + c;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
            f[i] += b.x;
            f[i + 1] += b.y;
            f[i + 2] += b.z;
            f[i + 3] += b.w;
          }
        }
        if (p.act == LN3_ACT_GELU_ERF) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
        } else if (p.act == LN3_ACT_GELU_TANH) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = gelu_tanh(f[i]);
        } else if (p.act == LN3_ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = silu(f[i]);
        }
        if (row_ok) {
          if (p.out_kind == LN3_OUT_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 q;
              q.x = pack_bf16x2(f[i], f[i + 1]);
              q.y = pack_bf16x2(f[i + 2], f[i + 3]);
              q.z = pack_bf16x2(f[i + 4], f[i + 5]);
              q.w = pack_bf16x2(f[i + 6], f[i + 7]);
              *reinterpret_cast<uint4*>(o + i) = q;
            }
          } else if (p.out_kind == LN3_OUT_F32) {
            float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          } else {  // LN3_OUT_RESID_F32: x[m,n] += gate * f
            float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + n0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 x = *reinterpret_cast<const float4*>(o + i);
              float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
              if (gate_row != nullptr) g = __ldg(reinterpret_cast<const float4*>(gate_row + n0 + i));
              x.x = fmaf(g.x, f[i], x.x);
              x.y = fmaf(g.y, f[i + 1], x.y);
              x.z = fmaf(g.z, f[i + 2], x.z);
              x.w = fmaf(g.w, f[i + 3], x.w);
              *reinterpret_cast<float4*>(o + i) = x;
              f[i] = x.x;
              f[i + 1] = x.y;
              f[i + 2] = x.z;
              f[i + 3] = x.w;
            }
            if (p.out2 != nullptr) {
              __nv_bfloat16* o2 = p.out2 + m * p.ldo2 + n0;
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 q;
                q.x = pack_bf16x2(f[i], f[i + 1]);
                q.y = pack_bf16x2(f[i + 2], f[i + 3]);
                q.z = pack_bf16x2(f[i + 4], f[i + 5]);
                q.w = pack_bf16x2(f[i + 6], f[i + 7]);
                *reinterpret_cast<uint4*>(o2 + i) = q;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------- host
template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error(LN3_ECUDA, "gemm: cudaFuncSetAttribute: %s",
                                           cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_bf16_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

int gemm_bf16(const ln3_gemm_args* a, cudaStream_t stream) {
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(LN3_EINVAL, "gemm: empty problem");
  if (a->K % BK != 0) return set_error(LN3_EINVAL, "gemm: K=%d must be a multiple of %d", a->K, BK);
  if (a->N % 128 != 0) return set_error(LN3_EINVAL, "gemm: N=%d must be a multiple of 128", a->N);
  if (a->lda % 8 != 0 || a->ldw % 8 != 0)
    return set_error(LN3_EINVAL, "gemm: lda/ldw must be multiples of 8 elements (16 bytes)");
  if ((reinterpret_cast<uintptr_t>(a->A) | reinterpret_cast<uintptr_t>(a->W) |
       reinterpret_cast<uintptr_t>(a->out)) & 15)
    return set_error(LN3_EINVAL, "gemm: pointers must be 16-byte aligned");
  if (a->out_kind == LN3_OUT_RESID_F32 && a->gate != nullptr && a->gate_rows <= 0)
    return set_error(LN3_EINVAL, "gemm: gate_rows must be > 0");
  const int bn = (a->N % 256 == 0) ? 256 : 128;

  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, a->A, a->M, a->K, a->lda, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, a->W, a->N, a->K, a->ldw, bn, BK);
  if (rc) return rc;

  GemmParams p;
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.act = a->act;
  p.out_kind = a->out_kind;
  p.bias = a->bias;
  p.out = a->out;
  p.ldo = a->ldo;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2);
  p.ldo2 = a->ldo2;
  p.gate = a->gate;
  p.gate_rows = a->gate_rows > 0 ? a->gate_rows : 1;
  p.gate_ld = a->gate_ld;
  const int sms = device_sm_count();
  if (bn == 256) return launch_gemm<256>(ta, tb, p, sms, stream);
  return launch_gemm<128>(ta, tb, p, sms, stream);
}

}  // namespace ln3
