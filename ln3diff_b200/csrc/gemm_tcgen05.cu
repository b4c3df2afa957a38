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
#include <cstdlib>

#include "common.cuh"
#include "ln3_internal.h"

namespace ln3 {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 bytes = one 128B-swizzle row
static constexpr int kGemmThreads = 192;
static constexpr int kStageLd = 36;  // floats per row of the epilogue transpose tile

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kEpiStageBytes = 4 * 32 * kStageLd * 4;  // 4 epilogue warps x 32x32 transpose tile
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiStageBytes;
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
  int epi_mode;            // bit0: stage bf16/f32 outputs through smem, bit1: stage residual updates
  int raster;              // pair kernel: 0 = row tile fastest (pairs of a round share a W tile), 1 = column tile fastest
                           // (the column tiles of one A row block run at the same time on neighbouring pairs)
  int wide_ok;             // out / out2 rows are 32-byte aligned -> 256-bit global accesses
  const float* hn_w;       // per-head RMSNorm weights [nsec][64] (HN kernels only)
  int hn_nsec, hn_sec_cols;
  float hn_eps;
  // stream-K tail of the CTA-pair kernel (sk_tiles == 0: plain data-parallel tiles)
  int sk_tiles;            // the last sk_tiles tiles are split along K over all pairs
  int* sk_flags;           // [pairs][2] counters, zero between launches (self-resetting)
  float* sk_partials;      // [pairs][2 ranks][32 column groups][128 rows][8] fp32 accumulator dumps
};

// One accumulator tile (this warp's 32 TMEM lanes x BN columns) -> global memory.
// Phase 1 (lane = row): tcgen05.ld, + bias, activation (or per-head RMSNorm), transpose through a
// private 32x32 smem tile.  Phase 2 (lane = column): every global access of the warp is one contiguous
// row segment (128 B fp32 / 64 B bf16) -> 1 L1 wavefront per instruction instead of 32.
template <int BN, bool HN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, float* stage, uint32_t t_row,
                                              int m_base, int tn, int lane) {
constexpr int CW = HN ? 64 : 32;  // columns per pass
#pragma unroll 1
for (int c = 0; c < BN; c += CW) {
  const int n0 = tn * BN + c;
  float f[CW];
  {
    uint32_t v[CW];
    tmem_ld_32x32(t_row + c, v);
    if constexpr (HN) tmem_ld_32x32(t_row + c + 32, v + 32);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < CW; ++i) f[i] = __uint_as_float(v[i]);
  }
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < CW; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
      f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
    }
  }
  if constexpr (HN) {
    const int sec = n0 / p.hn_sec_cols;
    if (sec < p.hn_nsec) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) ss = fmaf(f[i], f[i], ss);
      const float r = rsqrtf(ss * (1.0f / 64.0f) + p.hn_eps);
      const float* w = p.hn_w + sec * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w + i));
        f[i] *= r * ww.x; f[i + 1] *= r * ww.y; f[i + 2] *= r * ww.z; f[i + 3] *= r * ww.w;
      }
    }
  } else {
    if (p.act == LN3_ACT_GELU_ERF) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = gelu_erf_fast(f[i]);
    } else if (p.act == LN3_ACT_GELU_TANH) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = gelu_tanh(f[i]);
    } else if (p.act == LN3_ACT_SILU) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = silu(f[i]);
    } else if (p.act == LN3_ACT_QUICK_GELU) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = quick_gelu(f[i]);
    }
  }
  const bool staged = (p.out_kind == LN3_OUT_RESID_F32) ? (p.epi_mode & 2) != 0 : (p.epi_mode & 1) != 0;
  if (!staged) {
    // direct: thread = row, 16-byte vector accesses (each lane touches its own cache line)
    const int m = m_base + lane;
    if (m < p.M) {
#pragma unroll
      for (int h = 0; h < CW; h += 32) {
        const int nn = n0 + h;
        if (p.out_kind == LN3_OUT_BF16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + nn;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 q;
            q.x = pack_bf16x2(f[h + i], f[h + i + 1]);
            q.y = pack_bf16x2(f[h + i + 2], f[h + i + 3]);
            q.z = pack_bf16x2(f[h + i + 4], f[h + i + 5]);
            q.w = pack_bf16x2(f[h + i + 6], f[h + i + 7]);
            *reinterpret_cast<uint4*>(o + i) = q;
          }
        } else if (p.out_kind == LN3_OUT_F32) {
          float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + nn;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(f[h + i], f[h + i + 1], f[h + i + 2], f[h + i + 3]);
        } else {
          float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + nn;
          const float* gate_row = p.gate ? p.gate + static_cast<long long>(m / p.gate_rows) * p.gate_ld + nn : nullptr;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 x = *reinterpret_cast<const float4*>(o + i);
            float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
            if (gate_row != nullptr) g = __ldg(reinterpret_cast<const float4*>(gate_row + i));
            x.x = fmaf(g.x, f[h + i], x.x);
            x.y = fmaf(g.y, f[h + i + 1], x.y);
            x.z = fmaf(g.z, f[h + i + 2], x.z);
            x.w = fmaf(g.w, f[h + i + 3], x.w);
            *reinterpret_cast<float4*>(o + i) = x;
            f[h + i] = x.x; f[h + i + 1] = x.y; f[h + i + 2] = x.z; f[h + i + 3] = x.w;
          }
          if (p.out2 != nullptr) {
            __nv_bfloat16* o2 = p.out2 + m * p.ldo2 + nn;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 q;
              q.x = pack_bf16x2(f[h + i], f[h + i + 1]);
              q.y = pack_bf16x2(f[h + i + 2], f[h + i + 3]);
              q.z = pack_bf16x2(f[h + i + 4], f[h + i + 5]);
              q.w = pack_bf16x2(f[h + i + 6], f[h + i + 7]);
              *reinterpret_cast<uint4*>(o2 + i) = q;
            }
          }
        }
      }
    }
    continue;
  }
#pragma unroll
  for (int h = 0; h < CW; h += 32) {
    // transpose 32x32 through smem (row stride 36 floats: conflict-free STS.128 / LDS.32)
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      *reinterpret_cast<float4*>(stage + lane * kStageLd + i) =
          make_float4(f[h + i], f[h + i + 1], f[h + i + 2], f[h + i + 3]);
    __syncwarp();
    const int col = n0 + h + lane;
    if (p.out_kind == LN3_OUT_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + col;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        const int m = m_base + rr;
        if (m < p.M) o[m * p.ldo] = __float2bfloat16(stage[rr * kStageLd + lane]);
      }
    } else if (p.out_kind == LN3_OUT_F32) {
      float* o = reinterpret_cast<float*>(p.out) + col;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        const int m = m_base + rr;
        if (m < p.M) o[m * p.ldo] = stage[rr * kStageLd + lane];
      }
    } else {  // LN3_OUT_RESID_F32: x[m,n] += gate * val  (+ bf16 copy of the new x)
      float* o = reinterpret_cast<float*>(p.out) + col;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int m = m_base + rr;
        if (m < p.M) {
          const float g = p.gate ? __ldg(p.gate + static_cast<long long>(m / p.gate_rows) * p.gate_ld + col) : 1.f;
          const float xn = fmaf(g, stage[rr * kStageLd + lane], o[m * p.ldo]);
          o[m * p.ldo] = xn;
          if (p.out2 != nullptr) p.out2[m * p.ldo2 + col] = __float2bfloat16(xn);
        }
      }
    }
    __syncwarp();
  }
}
}

template <int BN, bool HN>
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
  float* stage_base = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256);

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
  pdl_launch_dependents();
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel
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
    {  // warp-uniform; elect_one_sync() only around the tcgen05 instructions (see the pair kernel)
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      const uint32_t tm_base = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem_a), 0, 1024);
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem_b), 0, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tm_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = a_desc0 + static_cast<uint32_t>(stage) * (Cfg::kABytes >> 4);
          const uint64_t db = b_desc0 + static_cast<uint32_t>(stage) * (Cfg::kBBytes >> 4);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
          }
          __syncwarp();
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one_sync()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // Epilogue warps 2..5 own TMEM lanes [32*(warp%4), +32).
    // Phase 1 (lane = row): tcgen05.ld 32 columns, + bias, activation (or per-head RMSNorm), then
    // transpose through a private 32x32 smem tile.  Phase 2 (lane = column): every global access of
    // the warp is one contiguous row segment (128 B fp32 / 64 B bf16) -> 1 L1 wavefront per
    // instruction instead of 32 (the per-thread-row layout was L1-wavefront bound on the residual
    // epilogue: 375 TF/s on the K=1024 projections).
    const int quarter = warp & 3;
    float* stage = stage_base + (warp - 2) * (32 * kStageLd);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int tm, tn;
      tile_coords(t, tm, tn);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m_base = tm * BM + quarter * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      epilogue_tile<BN, HN>(p, stage, t_row, m_base, tn, lane);
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
template <int BN, bool HN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static DeviceOnce once;
  if (int rc = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, HN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytes);
        return e == cudaSuccess ? LN3_OK : set_error(LN3_ECUDA, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      }))
    return rc;
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  cudaError_t e = launch_pdl(gemm_bf16_kernel<BN, HN>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, p);
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}



// Compile-time specialised, software-pipelined epilogue of the CTA-pair kernel: this warp's 32 TMEM
// lanes x columns [c_begin, c_end) of one accumulator tile.  The tcgen05.ld of chunk c+1 is in flight
// while chunk c is converted and stored; activation / output kind are template parameters (the
// runtime switch compiled to an indirect branch + constant loads that stalled the warp ~15 %).
// internal activation id (not in the ABI): erf-GELU by the packed polynomial of common.cuh
static constexpr int kActGeluErfPoly = 100;

template <int ACT, int OUT, bool HN>
__device__ __forceinline__ void epilogue_cols(const GemmParams& p, uint32_t t_row, int m, int n_tile0,
                                              int c_begin, int c_end, const float* bias_s,
                                              const float* part_row = nullptr, int nparts = 0, int c_lim = 1 << 30) {
  // c_lim (bf16 output only): tile columns >= c_lim are not stored -- the last 192-wide tile of a row ends at N,
  // possibly inside a 32-column chunk (multiples of 8)
  constexpr int CW = HN ? 64 : 32;
  // Plain epilogues software-pipeline the TMEM loads (chunk c+1 in flight while c is stored).  The
  // activation epilogues run with 16 warps and a 96-register budget instead: no prefetch registers,
  // the other three warps of the scheduler cover the tcgen05.ld latency.
  // (the head-norm epilogue holds a whole 64-column head per thread: with prefetch registers on top it spilled 1.5 KB)
  constexpr bool PREFETCH = ACT == LN3_ACT_NONE && !HN;
  const bool row_ok = m < p.M;
  uint32_t v[CW], vn[PREFETCH ? CW : 1];
  if constexpr (PREFETCH) {
    tmem_ld_32x32(t_row + c_begin, v);
    if constexpr (HN) tmem_ld_32x32(t_row + c_begin + 32, v + 32);
    tmem_ld_wait();
  }
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += CW) {
    const bool more = c + CW < c_end;
    if constexpr (PREFETCH) {
      if (more) {
        tmem_ld_32x32(t_row + c + CW, vn);
        if constexpr (HN) tmem_ld_32x32(t_row + c + CW + 32, vn + 32);
      }
    } else {
      tmem_ld_32x32(t_row + c, v);
      if constexpr (HN) tmem_ld_32x32(t_row + c + 32, v + 32);
      tmem_ld_wait();
    }
    const int n0 = n_tile0 + c;
    float f[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) f[i] = __uint_as_float(v[i]);
    // stream-K owner: add the K-range partial sums of the contributing pairs (this thread's row, fp32)
    // (slot layout [col / 8][row][8]: a warp's 32 rows of one 8-column group are 1 KB contiguous)
    for (int q = 0; q < nparts; ++q) {
      const float* pr = part_row + static_cast<long long>(q) * (2 * 128 * 256) + (c >> 3) * 1024;
      float t8[CW / 8][8];
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) ldg256_na(pr + i * 1024, t8[i]);
#pragma unroll
      for (int i = 0; i < CW / 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i * 8 + j] += t8[i][j];
    }
    if (p.bias != nullptr) {  // this tile's bias, staged in smem by the caller (a global load here stalls
#pragma unroll             // the chunk for an L2 round trip: 10 % of the kernel's stall samples)
      for (int i = 0; i < CW; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + c + i);
        f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
      }
    }
    if constexpr (HN) {
      const int sec = n0 / p.hn_sec_cols;
      if (sec < p.hn_nsec) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) ss = fmaf(f[i], f[i], ss);
        const float r = rsqrtf(ss * (1.0f / 64.0f) + p.hn_eps);
        const float* w = p.hn_w + sec * 64;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float4 ww = __ldg(reinterpret_cast<const float4*>(w + i));
          f[i] *= r * ww.x; f[i + 1] *= r * ww.y; f[i + 2] *= r * ww.z; f[i + 3] *= r * ww.w;
        }
      }
    }
    if constexpr (ACT == LN3_ACT_GELU_ERF) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = gelu_erf_fast(f[i]);
    } else if constexpr (ACT == kActGeluErfPoly) {
#pragma unroll
      for (int i = 0; i < CW; i += 2) gelu_erf_poly2(f[i], f[i + 1]);
    } else if constexpr (ACT == LN3_ACT_GELU_TANH) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = gelu_tanh(f[i]);
    } else if constexpr (ACT == LN3_ACT_SILU) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = silu(f[i]);
    } else if constexpr (ACT == LN3_ACT_QUICK_GELU) {
#pragma unroll
      for (int i = 0; i < CW; ++i) f[i] = quick_gelu(f[i]);
    }
    if (row_ok) {
      if constexpr (OUT == LN3_OUT_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n0;
        if (p.wide_ok && c + CW <= c_lim) {
#pragma unroll
          for (int i = 0; i < CW; i += 16) {
            uint32_t q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = pack_bf16x2(f[i + 2 * j], f[i + 2 * j + 1]);
            stg256_b32(o + i, q);
          }
        } else {
#pragma unroll
          for (int i = 0; i < CW; i += 8) {
            if (c + i + 8 > c_lim) break;
            uint4 q;
            q.x = pack_bf16x2(f[i], f[i + 1]);
            q.y = pack_bf16x2(f[i + 2], f[i + 3]);
            q.z = pack_bf16x2(f[i + 4], f[i + 5]);
            q.w = pack_bf16x2(f[i + 6], f[i + 7]);
            *reinterpret_cast<uint4*>(o + i) = q;
          }
        }
      } else if constexpr (OUT == LN3_OUT_F32) {
        float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + n0;
        if (p.wide_ok) {
#pragma unroll
          for (int i = 0; i < CW; i += 8) stg256_f32(o + i, f + i);
        } else {
#pragma unroll
          for (int i = 0; i < CW; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        }
      } else {  // LN3_OUT_RESID_F32: x += gate * val (256-bit accesses; wide_ok checked on the host)
        float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + n0;
        const float* gate_row = p.gate ? p.gate + static_cast<long long>(m / p.gate_rows) * p.gate_ld + n0 : nullptr;
        float x[CW];
#pragma unroll
        for (int i = 0; i < CW; i += 8) ldg256_na(o + i, x + i);
#pragma unroll
        for (int i = 0; i < CW; i += 4) {
          float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
          if (gate_row != nullptr) g = __ldg(reinterpret_cast<const float4*>(gate_row + i));
          x[i] = fmaf(g.x, f[i], x[i]);
          x[i + 1] = fmaf(g.y, f[i + 1], x[i + 1]);
          x[i + 2] = fmaf(g.z, f[i + 2], x[i + 2]);
          x[i + 3] = fmaf(g.w, f[i + 3], x[i + 3]);
        }
#pragma unroll
        for (int i = 0; i < CW; i += 8) stg256_f32(o + i, x + i);
        if (p.out2 != nullptr) {
          __nv_bfloat16* o2 = p.out2 + m * p.ldo2 + n0;
#pragma unroll
          for (int i = 0; i < CW; i += 16) {
            uint32_t q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = pack_bf16x2(x[i + 2 * j], x[i + 2 * j + 1]);
            stg256_b32(o2 + i, q);
          }
        }
      }
    }
    if constexpr (PREFETCH) {
      if (more) {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] = vn[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------- CTA-pair GEMM
// cta_group::2 variant: a cluster of two CTAs (one TPC) computes a 256 x 256 tile.  Each CTA loads its
// own 128 rows of A and HALF of the W tile (128 of the 256 output columns) per k-block -- 32 KB instead
// of 48 KB -- so the L2 -> SM operand traffic that bounds the 1-CTA kernel (148 x 96 B/clk > the ~6.3
// KB/clk L2 can deliver) drops by a third; the tensor cores of both SMs read both W halves.
//   both CTAs  : TMA producer (own A rows + own W half, bytes credited to the LEADER's full barrier),
//                TMEM allocation (cta_group::2), 4 epilogue warps on their own 128 accumulator rows
//   leader CTA : single-thread tcgen05.mma.cta_group::2 issuer (UMMA 256 x 256 x 16); its commits are
//                multicast to the empty / tmem_full barriers of both CTAs
//   tmem_empty : epilogue warps of both CTAs arrive on the leader's barrier (remote mbarrier arrive)
static constexpr int kStages2 = 6;
static constexpr int kStageBytes2 = (BM * BK + 128 * BK) * 2;  // A 128x64 + W half 128x64 = 32 KB
static constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256 + 2 * 256 * 4;  // + bias[2][256]
// Threads: TMA warp, MMA warp, EW epilogue warps (EW / 4 per TMEM lane quarter, each a column slice).
// EW = 8 for plain epilogues; EW = 16 for activation epilogues, whose per-element dependency chains
// (MUFU rcp/ex2 + Horner) leave a warp latency-bound: 4 warps per scheduler hide what 2 cannot.
template <int EW>
constexpr int gemm2_threads() { return 64 + 32 * EW; }
template <int ACT>
constexpr int gemm2_epi_warps() { return ACT == LN3_ACT_NONE ? 8 : 16; }

// BN = 256, or 192 / 176 (plain bf16 epilogue only): with T tiles on P pairs the cost is ceil(T / P) * BN.  For the
// N = 1024 GEMMs at M = 6144 (cond-only q / out of the DiT, proj / fc2 of the VAE decoder) 6 column tiles of 176
// (5 x 176 + 144) give 144 tiles = 1.95 -> 2 rounds of 0.69 instead of 96 tiles = 1.3 -> 2 rounds of 1.0: 20.3 -> 17.0 us.
// (At M = 12288 the model predicts 8 % -- 288 tiles = 3.89 rounds -- but the two extra passes over A cost more: fc2
// 85.5 -> 96.2 us, so the host only switches for a predicted gain >= 20 %.)
// The W box is BN / 2 rows per CTA; TMEM stages stay 256 columns apart; columns past N in the last tile of a row
// are zero-filled by TMA and never stored.
template <int ACT, int OUT, bool HN, int BN = 256>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm2_threads<gemm2_epi_warps<ACT>()>(), 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmParams p) {
  static_assert(BN == 256 || ((BN == 192 || BN == 176) && ACT == LN3_ACT_NONE && OUT == LN3_OUT_BF16 && !HN), "narrow tiles: plain bf16 epilogue");
  constexpr int EW = gemm2_epi_warps<ACT>();
  constexpr int kABytes = BM * BK * 2, kBBytes = 128 * BK * 2;   // smem strides (the W box uses BN / 2 of its 128 rows)
  constexpr int kStageTx = (BM * BK + (BN / 2) * BK) * 2;        // bytes one CTA's two TMA loads deliver per stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages2 * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2);
  uint64_t* full_bar = bars;                  // [kStages2]  (the leader's copy is the live one)
  uint64_t* empty_bar = bars + kStages2;      // [kStages2]  one per CTA, fed by multicast commits
  uint64_t* tmem_full = bars + 2 * kStages2;  // [2]         one per CTA, fed by multicast commits
  uint64_t* tmem_empty = tmem_full + 2;       // [2]         leader's copy: EW warps x 2 CTAs arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + kStages2 * kStageBytes2 + 256);  // [2 acc stages][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * EW);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int t, int& tm, int& tn) {
    if (p.raster) {
      tn = t % tiles_n;
      tm = t / tiles_n;
    } else {
      tm = t % tiles_m;
      tn = t / tiles_m;
    }
  };


  // Work list of this pair, identical in every warp role: data-parallel tiles pair, pair + P, ... below
  // dp_tiles, then (stream-K tail) this pair's contiguous share [u_lo, u_hi) of the sk_tiles * num_kb
  // k-block units of the last sk_tiles tiles.  A tile split over pairs a < b < ... is OWNED by a (it holds
  // the tile's first k-blocks and reaches them at the END of its share); b, ... hold later k-blocks at
  // the START of their shares, dump their raw accumulators early and never wait.
  const int dp_tiles = num_tiles - p.sk_tiles;
  const long long sk_total = static_cast<long long>(p.sk_tiles) * num_kb;
  const long long u_lo = sk_total * pair / num_pairs, u_hi = sk_total * (pair + 1) / num_pairs;
  // item iteration: returns false when exhausted.  mode 0 = whole tile, 1 = contributor, 2 = owner
  struct Item { int tile, kb0, kb1, mode, nparts; };
  auto next_item = [&](int& t_dp, long long& u, Item& it) -> bool {
    if (t_dp < dp_tiles) {
      it.tile = t_dp, it.kb0 = 0, it.kb1 = num_kb, it.mode = 0, it.nparts = 0;
      t_dp += num_pairs;
      return true;
    }
    if (u >= u_hi) return false;
    const int j = static_cast<int>(u / num_kb);
    it.tile = dp_tiles + j;
    it.kb0 = static_cast<int>(u - static_cast<long long>(j) * num_kb);
    const long long left = u_hi - u;
    it.kb1 = (num_kb - it.kb0 <= left) ? num_kb : it.kb0 + static_cast<int>(left);
    u += it.kb1 - it.kb0;
    it.nparts = 0;
    if (it.kb0 > 0) {
      it.mode = 1;
    } else if (it.kb1 == num_kb) {
      it.mode = 0;
    } else {
      it.mode = 2;
      const long long tile_end = static_cast<long long>(j + 1) * num_kb;
      for (int q = pair + 1; q < num_pairs && sk_total * q / num_pairs < tile_end; ++q) ++it.nparts;
    }
    return true;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int t_dp = pair;
      long long u = u_lo;
      Item it;
      while (next_item(t_dp, u, it)) {
        int tm, tn;
        tile_coords(it.tile, tm, tn);
        const int row_a = tm * 2 * BM + static_cast<int>(rank) * BM;
        const int row_b = tn * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t full_leader = mapa_u32(&full_bar[stage], 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageTx);  // bytes of both CTAs
          tma_load_2d_2sm(smem_a + stage * kABytes, &tmap_a, full_leader, kb * BK, row_a);
          tma_load_2d_2sm(smem_b + stage * kBBytes, &tmap_b, full_leader, kb * BK, row_b);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // Warp-uniform control flow + elect_one_sync() around the tcgen05 instructions only: with an
    // `if (lane == 0)` region ptxas wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall.
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);   // UMMA 256 x BN x 16 over the pair
      const uint32_t tm_base = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(smem_a), 0, 1024);
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem_b), 0, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int t_dp = pair;
      long long u = u_lo;
      Item it;
      while (next_item(t_dp, u, it)) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tm_base + acc * 256;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + static_cast<uint32_t>(stage) * (kABytes >> 4);
          const uint64_t b_desc = b_desc0 + static_cast<uint32_t>(stage) * (kBBytes >> 4);
          const uint32_t first = kb == it.kb0 ? 0u : 1u;
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0 ? 1u : first);
            umma_commit_2sm(&empty_bar[stage], 0b11);  // frees this stage in both CTAs
          }
          __syncwarp();
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one_sync()) umma_commit_2sm(&tmem_full[acc], 0b11);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int slice = (warp - 2) >> 2;  // which slice of the tile's columns
    constexpr int kSliceCols = BN / (EW / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    int t_dp = pair;
    long long u = u_lo;
    Item it;
    while (next_item(t_dp, u, it)) {
      int tm, tn;
      tile_coords(it.tile, tm, tn);
      // stage the tile's 256 bias values (one per epilogue thread) while the mainloop is still running;
      // two buffers + one barrier per tile: nobody can be two tiles ahead of the slowest warp
      if (p.bias != nullptr && threadIdx.x < 64 + BN) {
        const int bc = tn * BN + (threadIdx.x - 64);
        bias_s[acc * 256 + (threadIdx.x - 64)] = bc < p.N ? __ldg(p.bias + bc) : 0.f;
      }
      named_bar_sync(1, 32 * EW);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = tm * 2 * BM + static_cast<int>(rank) * BM + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256;
      // this thread's row inside a partial-sum slot: [pair][rank][256 / 8 column groups][128 rows][8]
      float* my_part = p.sk_partials + (static_cast<long long>(pair) * 2 + rank) * (128 * 256) + (quarter * 32 + lane) * 8;
      if (it.mode == 1) {
        // contributor: raw accumulator slice -> this pair's slot, then one flag tick per warp
#pragma unroll 1
        for (int c = slice * kSliceCols; c < (slice + 1) * kSliceCols; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 8) stg256_b32(my_part + ((c + i) >> 3) * 1024, v + i);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.sk_flags + pair * 2 + rank, 1);
      } else {
        if (it.mode == 2) {
          // owner: the contributors (pairs pair+1 .. pair+nparts) dumped their parts at the start of their
          // shares, long before this pair reaches the end of its own; wait for EW ticks per slot anyway
          if (lane == 0) {
            for (int q = 1; q <= it.nparts; ++q) {
              const int* f = p.sk_flags + (pair + q) * 2 + rank;
              long long spins = 0;
              while (ld_acquire_gpu(f) < EW) {
                if (++spins > (1ll << 28)) {
                  printf("ln3 gemm2: stream-K flag timeout (pair %d waits for %d)\n", pair, pair + q);
                  __trap();
                }
              }
            }
          }
          __syncwarp();
        }
        if constexpr (BN == 256) {
          epilogue_cols<ACT, OUT, HN>(p, t_row, m, tn * BN, slice * kSliceCols, (slice + 1) * kSliceCols, bias_s + acc * 256,
                                      my_part + 2 * 128 * 256, it.mode == 2 ? it.nparts : 0);
        } else {
          // BN columns as [0, 96) | [96, BN): both slices start on a 16-column boundary (aligned 256-bit stores); the
          // 176-wide tile ends inside its last 32-column chunk, the last tile of a row at N (multiples of 8 columns)
          const int c_lo = slice * 96, c_hi = slice == 0 ? 96 : BN;
          epilogue_cols<ACT, OUT, HN>(p, t_row, m, tn * BN, c_lo, c_hi, bias_s + acc * 256, nullptr, 0,
                                      min(c_hi, p.N - tn * BN));
        }
        if (it.mode == 2) {
          // second tick per warp; whoever brings a slot's counter to 2 * EW resets it for the next launch
          __syncwarp();
          if (lane == 0)
            for (int q = 1; q <= it.nparts; ++q) {
              int* f = p.sk_flags + (pair + q) * 2 + rank;
              if (atomicAdd(f, 1) == 2 * EW - 1) atomicExch(f, 0);
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(&tmem_empty[acc], 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer can still signal it
  tc_fence_after();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

size_t gemm_workspace_bytes() {
  return 1024 + static_cast<size_t>(device_sm_count() / 2) * 2 * 128 * 256 * sizeof(float);
}

template <int ACT, int OUT, bool HN, int BN = 256>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int num_sms,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  static DeviceOnce once;
  if (int rc = once.run([] {
        cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel<ACT, OUT, HN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmemBytes2);
        return e == cudaSuccess ? LN3_OK : set_error(LN3_ECUDA, "gemm2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      }))
    return rc;
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + BN - 1) / BN);
  const int num_kb = p.K / BK;
  const int all_pairs = num_sms / 2;
  int pairs = tiles < all_pairs ? tiles : all_pairs;
  // Stream-K tail: with T tiles on P pairs the last wave leaves P - T % P pairs idle for a whole tile
  // (192 tiles on 74 pairs: 3 rounds for 2.59 rounds of work).  When a workspace is given, the last
  // T % P tiles are split along K over all pairs instead.
  // Opt-in (LN3_GEMM_STREAMK=1): correct, but on the DiT shapes the fix-up currently costs more than the
  // idle tail it removes (qkv 56 -> 65 us, proj 25 -> 37 us, fc2 85 -> 84 us); kept for the next round.
  const char* sk_env = getenv("LN3_GEMM_STREAMK");
  const bool sk_off = !(sk_env && atoi(sk_env) != 0);
  p.sk_tiles = 0;
  p.sk_flags = nullptr;
  p.sk_partials = nullptr;
  if (BN == 256 && !sk_off && workspace != nullptr && workspace_bytes >= gemm_workspace_bytes() &&
      (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && num_kb >= 4) {
    int sk = 0;
    if (tiles >= all_pairs) {
      const int rem = tiles % all_pairs;
      if (rem != 0 && rem * 10 <= all_pairs * 9) sk = rem;
    } else if (tiles * 2 >= all_pairs && tiles * 10 <= all_pairs * 9) {
      sk = tiles;
    }
    if (sk > 0 && static_cast<long long>(sk) * num_kb >= 2LL * all_pairs) {
      p.sk_tiles = sk;
      p.sk_flags = reinterpret_cast<int*>(workspace);
      p.sk_partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 1024);
      pairs = all_pairs;
    }
  }
  cudaError_t e = launch_pdl(gemm2_bf16_kernel<ACT, OUT, HN, BN>, dim3(2 * pairs), dim3(gemm2_threads<gemm2_epi_warps<ACT>()>()),
                             kSmemBytes2, stream, ta, tb, p);
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "gemm2 launch: %s", cudaGetErrorString(e));
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
  static const bool force_1cta = getenv("LN3_GEMM_1CTA") != nullptr;
  bool use_pair = !force_1cta && bn == 256 && a->M >= 256;
  if (use_pair && a->out_kind != LN3_OUT_BF16 && a->act != LN3_ACT_NONE) use_pair = false;  // rare combos: generic kernel
  if (use_pair && a->out_kind == LN3_OUT_RESID_F32) {
    bool ok = (reinterpret_cast<uintptr_t>(a->out) % 32 == 0) && ((a->ldo * 4) % 32 == 0);
    if (a->out2 != nullptr) ok = ok && (reinterpret_cast<uintptr_t>(a->out2) % 32 == 0) && ((a->ldo2 * 2) % 32 == 0);
    if (!ok) use_pair = false;
  }

  // narrower column tiles (192 or 176) for the plain bf16 pair GEMM when they need fewer tile-rounds: cost = rounds x width
  int narrow = 0;
  if (use_pair && a->out_kind == LN3_OUT_BF16 && a->act == LN3_ACT_NONE && a->head_norm_w == nullptr) {
    static const int bn_env = getenv("LN3_GEMM_BN") ? atoi(getenv("LN3_GEMM_BN")) : 0;   // 256 / 192 / 176 force, 0 = cost model
    const int pairs = device_sm_count() / 2;
    const long long tm = (a->M + 2 * BM - 1) / (2 * BM);
    auto cost = [&](int w) { const long long t = tm * ((a->N + w - 1) / w); return ((t + pairs - 1) / pairs) * w; };
    if (bn_env == 192 || bn_env == 176) {
      narrow = bn_env;
    } else if (bn_env == 0 && tm * (a->N / 256) > 1) {
      const long long c256 = cost(256), c192 = cost(192), c176 = cost(176);
      const long long best = c192 <= c176 ? c192 : c176;
      // only when the model predicts >= 20 %: narrower tiles re-read the A operand once per column tile, and at
      // M = 12288, K = 4096 (fc2: 6 instead of 4 passes over 100 MB) the predicted 8 % became 12 % slower
      if (best * 100 < c256 * 80) narrow = c192 <= c176 ? 192 : 176;
    }
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_2d_bf16(&ta, a->A, a->M, a->K, a->lda, BM, BK);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tb, a->W, a->N, a->K, a->ldw, use_pair ? (narrow ? narrow / 2 : 128) : bn, BK);
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
  static const int epi_mode = getenv("LN3_GEMM_EPI") ? atoi(getenv("LN3_GEMM_EPI")) : 0;
  p.epi_mode = epi_mode;
  static const int raster_env = getenv("LN3_GEMM_RASTER") ? atoi(getenv("LN3_GEMM_RASTER")) : -1;
  // default: column tile fastest when a row block has few column tiles (N <= 2048): its A rows are then read by
  // neighbouring pairs at the same time and stream from HBM once (fc2 82.8 -> 75.5 us, step 11.52 -> 11.30 ms);
  // with many column tiles (qkv 12, fc1 16) row-tile-fastest keeps the W tile shared instead (qkv 58.2 vs 59.4 us)
  p.raster = raster_env >= 0 ? raster_env : (a->N <= 2048 ? 1 : 0);
  {
    const size_t esz = (a->out_kind == LN3_OUT_BF16) ? 2 : 4;
    bool ok = (reinterpret_cast<uintptr_t>(a->out) % 32 == 0) && ((a->ldo * esz) % 32 == 0);
    if (a->out2 != nullptr) ok = ok && (reinterpret_cast<uintptr_t>(a->out2) % 32 == 0) && ((a->ldo2 * 2) % 32 == 0);
    p.wide_ok = ok ? 1 : 0;
  }
  p.hn_w = a->head_norm_w;
  p.hn_nsec = a->head_norm_nsec;
  p.hn_sec_cols = a->head_norm_sec_cols;
  p.hn_eps = a->head_norm_eps;
  const int sms = device_sm_count();
  if (use_pair) {
    if (a->head_norm_w != nullptr) {
      if (a->out_kind != LN3_OUT_BF16 || a->act != LN3_ACT_NONE)
        return set_error(LN3_EINVAL, "gemm: head_norm needs LN3_OUT_BF16 and no activation");
      if (a->head_norm_nsec <= 0 || a->head_norm_sec_cols <= 0 || a->head_norm_sec_cols % 64 != 0)
        return set_error(LN3_EINVAL, "gemm: head_norm sections must be positive multiples of 64 columns");
      return launch_gemm2<LN3_ACT_NONE, LN3_OUT_BF16, true>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
    }
    if (a->out_kind == LN3_OUT_RESID_F32) {
      if (a->act != LN3_ACT_NONE) return set_error(LN3_EINVAL, "gemm: residual epilogue takes no activation");
      return launch_gemm2<LN3_ACT_NONE, LN3_OUT_RESID_F32, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
    }
    if (a->out_kind == LN3_OUT_F32) {
      if (a->act != LN3_ACT_NONE) return set_error(LN3_EUNSUPPORTED, "gemm: fp32 output with activation");
      return launch_gemm2<LN3_ACT_NONE, LN3_OUT_F32, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
    }
    if (narrow == 192) return launch_gemm2<LN3_ACT_NONE, LN3_OUT_BF16, false, 192>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
    if (narrow == 176) return launch_gemm2<LN3_ACT_NONE, LN3_OUT_BF16, false, 176>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
    switch (a->act) {
      case LN3_ACT_NONE: return launch_gemm2<LN3_ACT_NONE, LN3_OUT_BF16, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
      case LN3_ACT_GELU_ERF: {
        // LN3_GELU_EXACT=1: the A&S 7.1.26 form (|error| <= 1.5e-7, 2 MUFU per element) instead of the packed polynomial
        static const bool exact = getenv("LN3_GELU_EXACT") && atoi(getenv("LN3_GELU_EXACT")) != 0;
        if (exact) return launch_gemm2<LN3_ACT_GELU_ERF, LN3_OUT_BF16, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
        return launch_gemm2<kActGeluErfPoly, LN3_OUT_BF16, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
      }
      case LN3_ACT_GELU_TANH: return launch_gemm2<LN3_ACT_GELU_TANH, LN3_OUT_BF16, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
      case LN3_ACT_SILU: return launch_gemm2<LN3_ACT_SILU, LN3_OUT_BF16, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
      case LN3_ACT_QUICK_GELU: return launch_gemm2<LN3_ACT_QUICK_GELU, LN3_OUT_BF16, false>(ta, tb, p, sms, a->workspace, a->workspace_bytes, stream);
      default: return set_error(LN3_EINVAL, "gemm: unknown activation %d", a->act);
    }
  }
  if (a->head_norm_w != nullptr) {
    if (a->out_kind != LN3_OUT_BF16 || a->act != LN3_ACT_NONE)
      return set_error(LN3_EINVAL, "gemm: head_norm needs LN3_OUT_BF16 and no activation");
    if (a->head_norm_nsec <= 0 || a->head_norm_sec_cols <= 0 || a->head_norm_sec_cols % 64 != 0)
      return set_error(LN3_EINVAL, "gemm: head_norm sections must be positive multiples of 64 columns");
    if (bn == 256) return launch_gemm<256, true>(ta, tb, p, sms, stream);
    return launch_gemm<128, true>(ta, tb, p, sms, stream);
  }
  if (bn == 256) return launch_gemm<256, false>(ta, tb, p, sms, stream);
  return launch_gemm<128, false>(ta, tb, p, sms, stream);
}

}  // namespace ln3
