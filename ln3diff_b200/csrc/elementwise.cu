// HBM-bound glue kernels of the DiT / sampler path (fp32 SIMT, vectorised, coalesced):
//   norm_modulate      LayerNorm/RMSNorm + adaLN modulate -> bf16 GEMM operand
//   timestep_embedding sinusoidal features of the timestep
//   patch_embed        roll-out rearrange + 2x2 patch conv + pos_embed -> fp32 token stream
//   final_layer        LN + modulate + Linear(D -> 4*Cout) + unpatchify -> fp32 latent layout
//   sampler_update     x' = a x + w0 m0 + w1 m1 + s noise (all sampler engines, one launch/step)
// Reference call sites are cited per kernel in include/ln3b200.h.
#include <cstdlib>

#include "common.cuh"
#include "ln3_internal.h"

namespace ln3 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == LN3_ACT_SILU) return silu(v);
  if (act == LN3_ACT_GELU_ERF) return gelu_erf(v);
  if (act == LN3_ACT_GELU_TANH) return gelu_tanh(v);
  return v;
}

// ------------------------------------------------------------------ norm + modulate
// One warp per row; the row lives in registers (NV float4 per lane, D = 128 * NV).
template <int NV>
__global__ void __launch_bounds__(256)
norm_modulate_kernel(const ln3_norm_modulate_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // grid-stride over rows: the host sizes the grid to one resident wave, so there is no partial last wave
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < a.rows;
       row += gridDim.x * (blockDim.x >> 5)) {
  const float* x = a.x + static_cast<long long>(row) * a.ldx;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
  if (a.resid != nullptr) {
    // fused residual update: x += gate * resid (bf16), written back in place
    const __nv_bfloat16* rr = reinterpret_cast<const __nv_bfloat16*>(a.resid) + static_cast<long long>(row) * a.resid_ld;
    const __nv_bfloat16* r2 = nullptr;   // outside rows with resid_out_gate: own resid row first, then the bcast row
    const float* g2 = nullptr;
    if (a.resid_bcast != nullptr && (row < a.resid_row_begin || row >= a.resid_row_end)) {
      if (a.resid_out_gate != nullptr) {
        r2 = rr;
        g2 = a.resid_out_gate + static_cast<long long>(row / a.resid_out_gate_rows) * a.resid_out_gate_ld;
      }
      rr = reinterpret_cast<const __nv_bfloat16*>(a.resid_bcast) + static_cast<long long>(row / a.resid_bcast_rows) * a.resid_bcast_ld;
    }
    const float* gg = a.resid_gate ? a.resid_gate + static_cast<long long>(row / a.resid_gate_rows) * a.resid_gate_ld : nullptr;
    float* xw = const_cast<float*>(x);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (r2 != nullptr) {
        const uint2 qb = *reinterpret_cast<const uint2*>(r2 + c);
        const __nv_bfloat162 q01 = *reinterpret_cast<const __nv_bfloat162*>(&qb.x);
        const __nv_bfloat162 q23 = *reinterpret_cast<const __nv_bfloat162*>(&qb.y);
        const float4 h = __ldg(reinterpret_cast<const float4*>(g2 + c));
        v[i].x = fmaf(h.x, __low2float(q01), v[i].x);
        v[i].y = fmaf(h.y, __high2float(q01), v[i].y);
        v[i].z = fmaf(h.z, __low2float(q23), v[i].z);
        v[i].w = fmaf(h.w, __high2float(q23), v[i].w);
      }
      const uint2 rb = *reinterpret_cast<const uint2*>(rr + c);
      const __nv_bfloat162 r01 = *reinterpret_cast<const __nv_bfloat162*>(&rb.x);
      const __nv_bfloat162 r23 = *reinterpret_cast<const __nv_bfloat162*>(&rb.y);
      float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
      if (gg != nullptr && r2 == nullptr) g = __ldg(reinterpret_cast<const float4*>(gg + c));
      v[i].x = fmaf(g.x, __low2float(r01), v[i].x);
      v[i].y = fmaf(g.y, __high2float(r01), v[i].y);
      v[i].z = fmaf(g.z, __low2float(r23), v[i].z);
      v[i].w = fmaf(g.w, __high2float(r23), v[i].w);
      *reinterpret_cast<float4*>(xw + c) = v[i];
    }
    if (a.out == nullptr) continue;
  }

  float mean = 0.f, rstd = 1.f;
  if (a.norm == LN3_NORM_LAYER) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    mean = warp_sum(s) / static_cast<float>(a.D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(a.D) + a.eps);
  } else if (a.norm == LN3_NORM_RMS) {
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(a.D) + a.eps);
  }
  const float* sh = nullptr;
  const float* sc = nullptr;
  if (a.shift != nullptr) {
    const long long g = row / a.mod_rows;
    sh = a.shift + g * a.mod_ld;
    sc = a.scale + g * a.mod_ld;
  }
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<long long>(row) * a.ldo;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 y = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                           (v[i].w - mean) * rstd);
    if (a.weight != nullptr) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(a.weight + c));
      y.x *= w.x; y.y *= w.y; y.z *= w.z; y.w *= w.w;
    }
    if (sh != nullptr) {
      float4 s1 = __ldg(reinterpret_cast<const float4*>(sc + c));
      float4 s0 = __ldg(reinterpret_cast<const float4*>(sh + c));
      if (a.scale_tab != nullptr) {
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(a.scale_tab + c));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(a.shift_tab + c));
        s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
        s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
      }
      y.x = fmaf(y.x, 1.f + s1.x, s0.x);
      y.y = fmaf(y.y, 1.f + s1.y, s0.y);
      y.z = fmaf(y.z, 1.f + s1.z, s0.z);
      y.w = fmaf(y.w, 1.f + s1.w, s0.w);
    }
    if (a.act != LN3_ACT_NONE) {
      y.x = apply_act(y.x, a.act); y.y = apply_act(y.y, a.act);
      y.z = apply_act(y.z, a.act); y.w = apply_act(y.w, a.act);
    }
    uint2 pk;
    pk.x = pack_bf16x2(y.x, y.y);
    pk.y = pack_bf16x2(y.z, y.w);
    *reinterpret_cast<uint2*>(o + c) = pk;
  }
  }  // row loop
}

// ---- 256-bit kernels (D % 256 == 0, 32-byte aligned rows: every DiT / DiT2 call) -------------------------
// A lane owns NV8 chunks of 8 consecutive columns: the fp32 row moves as 256-bit accesses and the bf16 rows as
// 128-bit accesses.  nm_row_body is the arithmetic of one row, shared by the two kernels below; same operations in
// the same order per element as the float4 kernel above.
//   v      the row of x (registers)
//   rown   this row's own residual row resid[row] (bf16, one uint4 per chunk), meaningful iff nm_needs_own_row()
__device__ __forceinline__ bool nm_outside(const ln3_norm_modulate_args& a, int row) {
  return a.resid_bcast != nullptr && (row < a.resid_row_begin || row >= a.resid_row_end);
}
__device__ __forceinline__ bool nm_needs_own_row(const ln3_norm_modulate_args& a, int row) {
  return a.resid != nullptr && (!nm_outside(a, row) || a.resid_out_gate != nullptr);
}

// LN3_RESID_L2=1: residual-stream accesses carry the L2 evict_last priority (set once per process)
__constant__ int c_nm_l2_hint;

template <int NV8>
__device__ __forceinline__ void nm_row_body(const ln3_norm_modulate_args& a, int row, int lane, float (&v)[NV8][8],
                                            const uint4 (&rown)[NV8]) {
  float* x = const_cast<float*>(a.x) + static_cast<long long>(row) * a.ldx;
  auto ld8 = [&](const float* p8, float* d8) {
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(p8));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(p8 + 4));
    d8[0] = p0.x, d8[1] = p0.y, d8[2] = p0.z, d8[3] = p0.w, d8[4] = p1.x, d8[5] = p1.y, d8[6] = p1.z, d8[7] = p1.w;
  };
  auto axpy8 = [&](float* acc, const float* g8, const uint4& rb) {
    const uint32_t rw[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
      acc[2 * j] = fmaf(g8[2 * j], __low2float(r2), acc[2 * j]);
      acc[2 * j + 1] = fmaf(g8[2 * j + 1], __high2float(r2), acc[2 * j + 1]);
    }
  };
  if (a.resid != nullptr) {
    const bool outside = nm_outside(a, row);
    const bool own = !outside || a.resid_out_gate != nullptr;
    // gate of the own row: resid_gate inside, resid_out_gate outside
    const float* g_own = nullptr;
    if (!outside) {
      if (a.resid_gate) g_own = a.resid_gate + static_cast<long long>(row / a.resid_gate_rows) * a.resid_gate_ld;
    } else if (a.resid_out_gate) {
      g_own = a.resid_out_gate + static_cast<long long>(row / a.resid_out_gate_rows) * a.resid_out_gate_ld;
    }
    // outside rows: the per-group broadcast row, gated by resid_gate unless the own row took a gate of its own
    const __nv_bfloat16* rb = nullptr;
    const float* g_b = nullptr;
    if (outside) {
      rb = reinterpret_cast<const __nv_bfloat16*>(a.resid_bcast) + static_cast<long long>(row / a.resid_bcast_rows) * a.resid_bcast_ld;
      if (a.resid_out_gate == nullptr && a.resid_gate)
        g_b = a.resid_gate + static_cast<long long>(row / a.resid_gate_rows) * a.resid_gate_ld;
    }
#pragma unroll
    for (int i = 0; i < NV8; ++i) {
      const int c = (i * 32 + lane) * 8;
      float g[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (own) {
        if (g_own != nullptr) ld8(g_own + c, g);
        axpy8(v[i], g, rown[i]);
      }
      if (rb != nullptr) {
        float gb[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (g_b != nullptr) ld8(g_b + c, gb);
        axpy8(v[i], gb, *reinterpret_cast<const uint4*>(rb + c));
      }
      if (c_nm_l2_hint) stg256_f32_el(x + c, v[i]);
      else stg256_f32(x + c, v[i]);
    }
    if (a.out == nullptr) return;
  }
  float mean = 0.f, rstd = 1.f;
  if (a.norm == LN3_NORM_LAYER) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV8; ++i)  // same pairing as the float4 kernel: ((a+b)+(c+d)) per 4 columns
      s += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
    mean = warp_sum(s) / static_cast<float>(a.D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = v[i][j] - mean;
        q = fmaf(dlt, dlt, q);
      }
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(a.D) + a.eps);
  } else if (a.norm == LN3_NORM_RMS) {
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) q = fmaf(v[i][j], v[i][j], q);
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(a.D) + a.eps);
  }
  const float* sh = nullptr;
  const float* sc = nullptr;
  if (a.shift != nullptr) {
    const long long g = row / a.mod_rows;
    sh = a.shift + g * a.mod_ld;
    sc = a.scale + g * a.mod_ld;
  }
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + static_cast<long long>(row) * a.ldo;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const int c = (i * 32 + lane) * 8;
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = (v[i][j] - mean) * rstd;
    if (a.weight != nullptr) {
      float w[8];
      ld8(a.weight + c, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] *= w[j];
    }
    if (sh != nullptr) {
      float s1[8], s0[8];
      ld8(sc + c, s1);
      ld8(sh + c, s0);
      if (a.scale_tab != nullptr) {
        float t1[8], t0[8];
        ld8(a.scale_tab + c, t1);
        ld8(a.shift_tab + c, t0);
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] += t1[j], s0[j] += t0[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(y[j], 1.f + s1[j], s0[j]);
    }
    if (a.act != LN3_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = apply_act(y[j], a.act);
    }
    uint4 pk;
    pk.x = pack_bf16x2(y[0], y[1]);
    pk.y = pack_bf16x2(y[2], y[3]);
    pk.z = pack_bf16x2(y[4], y[5]);
    pk.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(o + c) = pk;
  }
}

// Warp per row straight from global memory.  (A software-pipelined variant -- next row's loads issued before the
// current row's arithmetic, 2 CTAs/SM at 128 registers -- measured slower: 36.9 vs 32.8 us; so did the shared-memory
// staged kernel below.  Occupancy at <= 80 registers beats explicit prefetch here.)
template <int NV8>
__device__ __forceinline__ void nm_load_row(const ln3_norm_modulate_args& a, int row, int lane, float (&v)[NV8][8],
                                            uint4 (&rown)[NV8]) {
  const float* x = a.x + static_cast<long long>(row) * a.ldx;
#pragma unroll
  if (c_nm_l2_hint) {
#pragma unroll
    for (int i = 0; i < NV8; ++i) ldg256_na_el(x + (i * 32 + lane) * 8, v[i]);
  } else {
#pragma unroll
    for (int i = 0; i < NV8; ++i) ldg256_na(x + (i * 32 + lane) * 8, v[i]);
  }
  if (nm_needs_own_row(a, row)) {
    const __nv_bfloat16* rr = reinterpret_cast<const __nv_bfloat16*>(a.resid) + static_cast<long long>(row) * a.resid_ld;
#pragma unroll
    for (int i = 0; i < NV8; ++i) rown[i] = *reinterpret_cast<const uint4*>(rr + (i * 32 + lane) * 8);
  } else {
#pragma unroll
    for (int i = 0; i < NV8; ++i) rown[i] = make_uint4(0, 0, 0, 0);
  }
}

template <int NV8>
__global__ void __launch_bounds__(256, 3)
norm_modulate_wide_kernel(const ln3_norm_modulate_args a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // A CTA owns a contiguous run of rows (not a grid-stride comb): consecutive token rows share their sample's
  // shift / scale / gate vectors, which then stay in L1 instead of every SM cycling through all samples' vectors
  // (LN + residual pass 36-38 -> 32.8 us at DiT-L/2 B'=16).
  const int per_cta = (a.rows + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int row_begin = static_cast<int>(blockIdx.x) * per_cta;
  const int row_end = min(a.rows, row_begin + per_cta);
  for (int row = row_begin + (threadIdx.x >> 5); row < row_end; row += (blockDim.x >> 5)) {
    float v[NV8][8];
    uint4 rown[NV8];
    nm_load_row<NV8>(a, row, lane, v, rown);
    nm_row_body<NV8>(a, row, lane, v, rown);
  }
}

// Staged variant for the big residual-stream passes (the three per DiT block: 150 MB each at DiT-L/2 B'=16).
// The warp-per-row kernel keeps at most one row per warp in flight and stops loading while it reduces and stores
// (ncu: 4.2-4.9 TB/s = 0.64-0.75 of the measured copy bandwidth, dram ~40 % busy).  Here a producer thread streams
// the fp32 rows (and their bf16 residual rows) into a 4-stage shared-memory ring with 1-D bulk copies
// (cp.async.bulk, byte-counted mbarriers), 8 rows per stage, so ~144 KB per SM is always in flight while 8 consumer
// warps (one row each) run the unchanged arithmetic out of shared memory and store straight to global.
static constexpr int kNmStages = 4, kNmRowsPerStage = 8;
template <int NV8>
constexpr int nm_stage_bytes() { return kNmRowsPerStage * NV8 * 256 * (4 + 2); }
template <int NV8>
constexpr int nm_staged_smem() { return kNmStages * nm_stage_bytes<NV8>() + 128; }

template <int NV8>
__global__ void __launch_bounds__((kNmRowsPerStage + 1) * 32, 1)
norm_modulate_staged_kernel(const ln3_norm_modulate_args a) {
  extern __shared__ __align__(128) uint8_t nm_smem[];
  constexpr int D = NV8 * 256;
  constexpr int kXBytes = D * 4, kRBytes = D * 2;
  uint64_t* full = reinterpret_cast<uint64_t*>(nm_smem + kNmStages * nm_stage_bytes<NV8>());
  uint64_t* empty = full + kNmStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kNmStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kNmRowsPerStage);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  const int n_groups = (a.rows + kNmRowsPerStage - 1) / kNmRowsPerStage;
  // contiguous run of row groups per CTA: the per-sample modulation vectors (shift / scale / gate rows, shared by
  // 768 consecutive token rows) stay hot in the 28 KB of L1 left beside the ring
  const int gpb = (n_groups + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int g_begin = static_cast<int>(blockIdx.x) * gpb;
  const int g_end = min(n_groups, g_begin + gpb);
  if (warp == kNmRowsPerStage) {
    // ---- producer: one lane issues the bulk copies of a whole stage
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int g = g_begin; g < g_end; ++g) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sx = nm_smem + stage * nm_stage_bytes<NV8>();
        uint8_t* sr = sx + kNmRowsPerStage * kXBytes;
        const int row0 = g * kNmRowsPerStage;
        const int nrows = min(kNmRowsPerStage, a.rows - row0);
        // contiguous rows (the usual case: x and resid are dense [rows, D] buffers) move as ONE bulk copy per
        // operand and stage -- a bulk-copy instruction costs hundreds of issue cycles on the producer thread, and
        // sixteen 2-4 KB copies per stage made the producer the bottleneck (2.4 TB/s)
        int n_own = 0;
        for (int r = 0; r < nrows; ++r) n_own += nm_needs_own_row(a, row0 + r) ? 1 : 0;
        mbar_arrive_expect_tx(&full[stage], nrows * kXBytes + n_own * kRBytes);
        if (a.ldx == D) {
          bulk_load_1d(sx, a.x + static_cast<long long>(row0) * a.ldx, nrows * kXBytes, &full[stage]);
        } else {
          for (int r = 0; r < nrows; ++r)
            bulk_load_1d(sx + r * kXBytes, a.x + static_cast<long long>(row0 + r) * a.ldx, kXBytes, &full[stage]);
        }
        if (n_own == nrows && a.resid_ld == D) {
          bulk_load_1d(sr, reinterpret_cast<const __nv_bfloat16*>(a.resid) + static_cast<long long>(row0) * a.resid_ld,
                       nrows * kRBytes, &full[stage]);
        } else if (n_own > 0) {
          for (int r = 0; r < nrows; ++r)
            if (nm_needs_own_row(a, row0 + r))
              bulk_load_1d(sr + r * kRBytes,
                           reinterpret_cast<const __nv_bfloat16*>(a.resid) + static_cast<long long>(row0 + r) * a.resid_ld,
                           kRBytes, &full[stage]);
        }
        if (++stage == kNmStages) stage = 0, phase ^= 1;
      }
    }
    return;
  }
  // ---- consumers: warp w owns row w of every stage
  int stage = 0;
  uint32_t phase = 0;
  for (int g = g_begin; g < g_end; ++g) {
    const int row = g * kNmRowsPerStage + warp;
    mbar_wait(&full[stage], phase);
    float v[NV8][8];
    uint4 rown[NV8];
    const bool live = row < a.rows;
    if (live) {
      const uint8_t* sx = nm_smem + stage * nm_stage_bytes<NV8>() + warp * kXBytes;
      const uint8_t* sr = nm_smem + stage * nm_stage_bytes<NV8>() + kNmRowsPerStage * kXBytes + warp * kRBytes;
      const bool own = nm_needs_own_row(a, row);
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        const float4 lo = *reinterpret_cast<const float4*>(sx + (i * 32 + lane) * 32);
        const float4 hi = *reinterpret_cast<const float4*>(sx + (i * 32 + lane) * 32 + 16);
        v[i][0] = lo.x, v[i][1] = lo.y, v[i][2] = lo.z, v[i][3] = lo.w;
        v[i][4] = hi.x, v[i][5] = hi.y, v[i][6] = hi.z, v[i][7] = hi.w;
        rown[i] = own ? *reinterpret_cast<const uint4*>(sr + (i * 32 + lane) * 16) : make_uint4(0, 0, 0, 0);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);   // the row is in registers: the producer may refill the slot
    if (live) nm_row_body<NV8>(a, row, lane, v, rown);
    if (++stage == kNmStages) stage = 0, phase ^= 1;
  }
}

int norm_modulate(const ln3_norm_modulate_args* a, cudaStream_t stream) {
  if (a->rows <= 0) return LN3_OK;
  if (a->D % 128 != 0 || a->D > 2048 || a->D <= 0)
    return set_error(LN3_EINVAL, "norm_modulate: D=%d must be a multiple of 128, <= 2048", a->D);
  if ((a->shift == nullptr) != (a->scale == nullptr))
    return set_error(LN3_EINVAL, "norm_modulate: shift and scale must be given together");
  if ((a->shift_tab == nullptr) != (a->scale_tab == nullptr))
    return set_error(LN3_EINVAL, "norm_modulate: shift_tab and scale_tab must be given together");
  if (a->shift != nullptr && a->mod_rows <= 0)
    return set_error(LN3_EINVAL, "norm_modulate: mod_rows must be > 0");
  if (a->ldx % 4 || a->ldo % 4 || (a->shift && a->mod_ld % 4))
    return set_error(LN3_EINVAL, "norm_modulate: leading dimensions must be multiples of 4");
  if (a->out == nullptr && a->resid == nullptr) return set_error(LN3_EINVAL, "norm_modulate: out is NULL");
  if (a->resid != nullptr) {
    if (a->resid_ld % 4) return set_error(LN3_EINVAL, "norm_modulate: resid_ld must be a multiple of 4");
    if (a->resid_gate != nullptr && (a->resid_gate_rows <= 0 || a->resid_gate_ld % 4))
      return set_error(LN3_EINVAL, "norm_modulate: bad resid_gate_rows / resid_gate_ld");
    if (a->resid_bcast != nullptr &&
        (a->resid_bcast_rows <= 0 || a->resid_bcast_ld % 8 || (reinterpret_cast<uintptr_t>(a->resid_bcast) & 15) ||
         a->resid_row_begin < 0 || a->resid_row_end < a->resid_row_begin || a->resid_row_end > a->rows))
      return set_error(LN3_EINVAL, "norm_modulate: bad resid_bcast arguments");
    if (a->resid_out_gate != nullptr &&
        (a->resid_bcast == nullptr || a->resid_out_gate_rows <= 0 || a->resid_out_gate_ld % 4))
      return set_error(LN3_EINVAL, "norm_modulate: resid_out_gate needs resid_bcast, rows > 0 and ld %% 4 == 0");
  } else if (a->resid_bcast != nullptr) {
    return set_error(LN3_EINVAL, "norm_modulate: resid_bcast needs resid");
  }
  const int warps = 8;
  const int blocks_needed = (a->rows + warps - 1) / warps;
  const int wave = device_sm_count() * 4;  // 4 x 256-thread blocks resident per SM (<= 64 regs/thread)
  dim3 grid(blocks_needed < wave ? blocks_needed : wave), block(warps * 32);
  static const bool wide_enabled = !(getenv("LN3_NORM_WIDE") && atoi(getenv("LN3_NORM_WIDE")) == 0);
  const bool wide = wide_enabled && a->D % 256 == 0 && a->D <= 1536 && a->ldx % 8 == 0 && a->ldo % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(a->x) & 31) == 0 && (a->out == nullptr || (reinterpret_cast<uintptr_t>(a->out) & 15) == 0) &&
                    (a->resid == nullptr || (a->resid_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0));
  // opt-in (LN3_NORM_STAGED=1): shared-memory staged kernel for the big residual-stream passes.  Measured slower
  // than the warp-per-row kernel (48 vs 36 us for LN + residual at DiT-L/2 B'=16): with the ring taking 196 KB the
  // CTA has 8 consumer warps and 28 KB of L1, and the per-row arithmetic (two reductions, the modulation-vector
  // loads) then bounds the pass, not the loads.
  {  // LN3_RESID_L2=1: evict_last hints on the residual stream (uploaded once per device)
    static DeviceOnce l2_once;
    if (int rc = l2_once.run([] {
          const int v = (getenv("LN3_RESID_L2") && atoi(getenv("LN3_RESID_L2")) != 0) ? 1 : 0;
          cudaError_t e = cudaMemcpyToSymbol(c_nm_l2_hint, &v, sizeof(v));
          return e == cudaSuccess ? LN3_OK : set_error(LN3_ECUDA, "norm_modulate: constant upload: %s", cudaGetErrorString(e));
        }))
      return rc;
  }
  static const bool staged_enabled = getenv("LN3_NORM_STAGED") && atoi(getenv("LN3_NORM_STAGED")) != 0;
  if (wide && staged_enabled && a->D <= 1024 && a->rows >= 2048 && (a->ldx * 4) % 16 == 0 &&
      (a->resid == nullptr || (a->resid_ld * 2) % 16 == 0)) {
    const int sms = device_sm_count();
    const int n_groups = (a->rows + kNmRowsPerStage - 1) / kNmRowsPerStage;
    const dim3 sgrid(n_groups < sms ? n_groups : sms), sblock((kNmRowsPerStage + 1) * 32);
    static DeviceOnce once;
    if (int rc = once.run([] {
          cudaError_t e = cudaFuncSetAttribute(norm_modulate_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, nm_staged_smem<1>());
          if (e == cudaSuccess) e = cudaFuncSetAttribute(norm_modulate_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, nm_staged_smem<2>());
          if (e == cudaSuccess) e = cudaFuncSetAttribute(norm_modulate_staged_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, nm_staged_smem<3>());
          if (e == cudaSuccess) e = cudaFuncSetAttribute(norm_modulate_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, nm_staged_smem<4>());
          return e == cudaSuccess ? LN3_OK : set_error(LN3_ECUDA, "norm_modulate: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }))
      return rc;
    cudaError_t le = cudaSuccess;
    switch (a->D / 256) {
      case 1: le = launch_pdl(norm_modulate_staged_kernel<1>, sgrid, sblock, nm_staged_smem<1>(), stream, *a); break;
      case 2: le = launch_pdl(norm_modulate_staged_kernel<2>, sgrid, sblock, nm_staged_smem<2>(), stream, *a); break;
      case 3: le = launch_pdl(norm_modulate_staged_kernel<3>, sgrid, sblock, nm_staged_smem<3>(), stream, *a); break;
      default: le = launch_pdl(norm_modulate_staged_kernel<4>, sgrid, sblock, nm_staged_smem<4>(), stream, *a); break;
    }
    cudaError_t e = le != cudaSuccess ? le : cudaGetLastError();
    if (e != cudaSuccess) return set_error(LN3_ECUDA, "norm_modulate launch: %s", cudaGetErrorString(e));
    count_launch();
    return LN3_OK;
  }
  if (wide) {
    cudaError_t le = cudaSuccess;
    const int wave3 = device_sm_count() * 3;  // 3 resident blocks per SM at <= 80 registers
    grid = dim3(blocks_needed < wave3 ? blocks_needed : wave3);
    switch (a->D / 256) {
      case 1: le = launch_pdl(norm_modulate_wide_kernel<1>, grid, block, 0, stream, *a); break;
      case 2: le = launch_pdl(norm_modulate_wide_kernel<2>, grid, block, 0, stream, *a); break;
      case 3: le = launch_pdl(norm_modulate_wide_kernel<3>, grid, block, 0, stream, *a); break;
      case 4: le = launch_pdl(norm_modulate_wide_kernel<4>, grid, block, 0, stream, *a); break;
      case 5: le = launch_pdl(norm_modulate_wide_kernel<5>, grid, block, 0, stream, *a); break;
      default: le = launch_pdl(norm_modulate_wide_kernel<6>, grid, block, 0, stream, *a); break;
    }
    cudaError_t e = le != cudaSuccess ? le : cudaGetLastError();
    if (e != cudaSuccess) return set_error(LN3_ECUDA, "norm_modulate launch: %s", cudaGetErrorString(e));
    count_launch();
    return LN3_OK;
  }
  switch (a->D / 128) {
#define LN3_NM_CASE(n) \
  case n: norm_modulate_kernel<n><<<grid, block, 0, stream>>>(*a); break;
    LN3_NM_CASE(1) LN3_NM_CASE(2) LN3_NM_CASE(3) LN3_NM_CASE(4) LN3_NM_CASE(5) LN3_NM_CASE(6)
    LN3_NM_CASE(7) LN3_NM_CASE(8) LN3_NM_CASE(9) LN3_NM_CASE(10) LN3_NM_CASE(11) LN3_NM_CASE(12)
    LN3_NM_CASE(13) LN3_NM_CASE(14) LN3_NM_CASE(15) LN3_NM_CASE(16)
#undef LN3_NM_CASE
    default: return set_error(LN3_EINVAL, "norm_modulate: unsupported D");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "norm_modulate launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

// ------------------------------------------------------------------ timestep embedding
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B,
                                          __nv_bfloat16* __restrict__ out) {
  const int b = blockIdx.x;
  const int i = threadIdx.x;  // 0..127
  if (b >= B) return;
  // freqs = exp(-ln(10000) * i / 128) in fp32, as torch.exp(fp32 tensor) computes it
  const float f = expf(-9.210340371976184f * static_cast<float>(i) / 128.0f);
  const float arg = t[b] * f;
  out[b * 256 + i] = __float2bfloat16(cosf(arg));
  out[b * 256 + 128 + i] = __float2bfloat16(sinf(arg));
}

int timestep_embedding(const float* t, int B, void* out_bf16, cudaStream_t stream) {
  if (B <= 0) return LN3_OK;
  timestep_embedding_kernel<<<B, 128, 0, stream>>>(t, B, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "timestep_embedding launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

// ------------------------------------------------------------------ patch embed
// Block = one token (b, n, l); threads stride over D.  K = Cin*4 (<= 64) inputs staged in smem.
__global__ void __launch_bounds__(256)
patch_embed_kernel(const ln3_patch_embed_args a) {
  __shared__ float xin[64];
  const int P = a.S / 2, L = P * P;
  const int tok = blockIdx.x;  // b * 3L + n * L + l
  const int b = tok / (3 * L);
  const int nl = tok - b * 3 * L;
  const int n = nl / L, l = nl - n * L;
  const int pi = l / P, pj = l - pi * P;
  const int K = a.Cin * 4;
  if (threadIdx.x < K) {
    const int c = threadIdx.x >> 2, p = (threadIdx.x >> 1) & 1, q = threadIdx.x & 1;
    const float s = a.in_scale ? a.in_scale[b] : 1.f;
    xin[threadIdx.x] =
        s * a.x[((static_cast<long long>(b) * (3 * a.Cin) + c * 3 + n) * a.S + 2 * pi + p) * a.S +
                2 * pj + q];
  }
  __syncthreads();
  float* o = a.tokens + static_cast<long long>(tok) * a.D;
  const float* pe = a.pos_embed ? a.pos_embed + static_cast<long long>(nl) * a.D : nullptr;
  for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
    float acc = a.bias ? a.bias[d] : 0.f;
    const float* w = a.weight + static_cast<long long>(d) * K;
    for (int k = 0; k < K; ++k) acc = fmaf(w[k], xin[k], acc);
    if (pe) acc += pe[d];
    o[d] = acc;
  }
}

// Cin == 4 fast path (the tri-latent: K = 16).  Block = 16 consecutive tokens; each thread keeps the
// 16 weights of its output channel in registers and walks the tokens, so the weight matrix is read once
// per block instead of once per token and every store is a coalesced 1 KB row segment.
constexpr int kPeTok = 16;
__global__ void __launch_bounds__(256)
patch_embed_k16_kernel(const ln3_patch_embed_args a) {
  __shared__ float xin[kPeTok][16];
  const int P = a.S / 2, L = P * P;
  const int ntok = a.B * 3 * L;
  const int tok0 = blockIdx.x * kPeTok;
  {
    const int t = threadIdx.x >> 4, k = threadIdx.x & 15;
    const int tok = tok0 + t;
    if (tok < ntok) {
      const int b = tok / (3 * L);
      const int nl = tok - b * 3 * L;
      const int n = nl / L, l = nl - n * L;
      const int pi = l / P, pj = l - pi * P;
      const int c = k >> 2, p = (k >> 1) & 1, q = k & 1;
      const float s = a.in_scale ? a.in_scale[b] : 1.f;
      xin[t][k] = s * a.x[((static_cast<long long>(b) * 12 + c * 3 + n) * a.S + 2 * pi + p) * a.S + 2 * pj + q];
    }
  }
  __syncthreads();
  // a thread owns four consecutive output channels: 64 weights in registers, 128-bit pos_embed loads and
  // token stores (the 4-byte version ran at 0.8 TB/s)
  for (int d = threadIdx.x * 4; d < a.D; d += 1024) {
    float w[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4* wp = reinterpret_cast<const float4*>(a.weight + static_cast<long long>(d + j) * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(wp + i);
        w[j][4 * i] = v.x, w[j][4 * i + 1] = v.y, w[j][4 * i + 2] = v.z, w[j][4 * i + 3] = v.w;
      }
    }
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bias = __ldg(reinterpret_cast<const float4*>(a.bias + d));
    const int nl0 = tok0 % (3 * L);
#pragma unroll 4
    for (int t = 0; t < kPeTok; ++t) {
      const int tok = tok0 + t;
      float acc[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float xv = xin[t][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(w[j][k], xv, acc[j]);
      }
      if (tok < ntok) {
        if (a.pos_embed) {
          const float4 pe = __ldg(reinterpret_cast<const float4*>(a.pos_embed + static_cast<long long>((nl0 + t) % (3 * L)) * a.D + d));
          acc[0] += pe.x, acc[1] += pe.y, acc[2] += pe.z, acc[3] += pe.w;
        }
        *reinterpret_cast<float4*>(a.tokens + static_cast<long long>(tok) * a.D + d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      }
    }
  }
}

int patch_embed(const ln3_patch_embed_args* a, cudaStream_t stream) {
  if (a->B <= 0) return LN3_OK;
  if (a->S % 2 || a->Cin <= 0 || a->Cin > 16 || a->D <= 0)
    return set_error(LN3_EINVAL, "patch_embed: need even S, 1 <= Cin <= 16");
  const int L = (a->S / 2) * (a->S / 2);
  if (a->Cin == 4 && a->D % 4 == 0)
    patch_embed_k16_kernel<<<(a->B * 3 * L + kPeTok - 1) / kPeTok, 256, 0, stream>>>(*a);
  else
    patch_embed_kernel<<<a->B * 3 * L, 256, 0, stream>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "patch_embed launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

// ------------------------------------------------------------------ final layer
// One warp per token: LN + modulate in registers, then 4*Cout dot products (warp reductions),
// scattered into the unpatchified '(b, c*3+n, 2i+p, 2j+q)' layout.
// Two tokens per warp: the 4*Cout weight rows are read once for both, and the 2 x 16 dot-product partials
// are reduced with one 31-shuffle reduce-scatter (lane tk*16 + o ends up with output o of token tk) instead
// of 32 five-step warp sums.  Requires 4*Cout == 16 (every release config: Cout = 4); other sizes take
// the one-token kernel below.
template <int NV>
__global__ void __launch_bounds__(128)
final_layer2_kernel(const ln3_final_layer_args a) {
  const int P = a.S / 2, L = P * P;
  const int ntok = a.B * 3 * L;
  const int tok0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;
  if (tok0 >= ntok) return;
  const int lane = threadIdx.x & 31;
  float4 v[2][NV];
#pragma unroll
  for (int tk = 0; tk < 2; ++tk) {
    const int tok = tok0 + tk < ntok ? tok0 + tk : ntok - 1;
    const int b = tok / (3 * L);
    const float* x = a.x + static_cast<long long>(tok) * a.D;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[tk][i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[tk][i].x + v[tk][i].y) + (v[tk][i].z + v[tk][i].w);
    const float mean = warp_sum(s) / static_cast<float>(a.D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float dx = v[tk][i].x - mean, dy = v[tk][i].y - mean, dz = v[tk][i].z - mean, dw = v[tk][i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(a.D) + 1e-6f);
    const float* sh = a.shift + static_cast<long long>(b) * a.mod_ld;
    const float* sc = a.scale + static_cast<long long>(b) * a.mod_ld;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 s1 = __ldg(reinterpret_cast<const float4*>(sc + c));
      float4 s0 = __ldg(reinterpret_cast<const float4*>(sh + c));
      if (a.scale_tab != nullptr) {
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(a.scale_tab + c));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(a.shift_tab + c));
        s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
        s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
      }
      v[tk][i].x = fmaf((v[tk][i].x - mean) * rstd, 1.f + s1.x, s0.x);
      v[tk][i].y = fmaf((v[tk][i].y - mean) * rstd, 1.f + s1.y, s0.y);
      v[tk][i].z = fmaf((v[tk][i].z - mean) * rstd, 1.f + s1.z, s0.z);
      v[tk][i].w = fmaf((v[tk][i].w - mean) * rstd, 1.f + s1.w, s0.w);
    }
  }
  float acc[32];  // [tk * 16 + o]
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    const float* w = a.weight + static_cast<long long>(o) * a.D;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (i * 32 + lane) * 4));
      a0 = fmaf(v[0][i].x, ww.x, a0); a0 = fmaf(v[0][i].y, ww.y, a0);
      a0 = fmaf(v[0][i].z, ww.z, a0); a0 = fmaf(v[0][i].w, ww.w, a0);
      a1 = fmaf(v[1][i].x, ww.x, a1); a1 = fmaf(v[1][i].y, ww.y, a1);
      a1 = fmaf(v[1][i].z, ww.z, a1); a1 = fmaf(v[1][i].w, ww.w, a1);
    }
    acc[o] = a0;
    acc[16 + o] = a1;
  }
  // reduce-scatter over the warp: after the stage with offset h, acc[0 .. h) holds partial sums of the
  // values whose index has bit h equal to the lane's bit h
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? acc[i] : acc[i + h];
      const float keep = up ? acc[i + h] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  const int tk = lane >> 4, oidx = lane & 15;
  const int tok = tok0 + tk;
  if (tok < ntok) {
    const int b = tok / (3 * L);
    const int nl = tok - b * 3 * L;
    const int n = nl / L, l = nl - n * L;
    const int pi = l / P, pj = l - pi * P;
    // unpatchify: feature index = (p * 2 + q) * Cout + c   ('nhwpqc->nchpwq')
    const int c = oidx % a.Cout, pq = oidx / a.Cout, pp = pq >> 1, qq = pq & 1;
    a.out[((static_cast<long long>(b) * (3 * a.Cout) + c * 3 + n) * a.S + 2 * pi + pp) * a.S + 2 * pj + qq] =
        acc[0] + (a.bias ? a.bias[oidx] : 0.f);
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
final_layer_kernel(const ln3_final_layer_args a) {
  const int P = a.S / 2, L = P * P;
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= a.B * 3 * L) return;
  const int lane = threadIdx.x & 31;
  const int b = tok / (3 * L);
  const int nl = tok - b * 3 * L;
  const int n = nl / L, l = nl - n * L;
  const int pi = l / P, pj = l - pi * P;
  const float* x = a.x + static_cast<long long>(tok) * a.D;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(a.D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(a.D) + 1e-6f);
  const float* sh = a.shift + static_cast<long long>(b) * a.mod_ld;
  const float* sc = a.scale + static_cast<long long>(b) * a.mod_ld;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 s1 = __ldg(reinterpret_cast<const float4*>(sc + c));
    float4 s0 = __ldg(reinterpret_cast<const float4*>(sh + c));
    if (a.scale_tab != nullptr) {
      const float4 t1 = __ldg(reinterpret_cast<const float4*>(a.scale_tab + c));
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(a.shift_tab + c));
      s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
      s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
    }
    v[i].x = fmaf((v[i].x - mean) * rstd, 1.f + s1.x, s0.x);
    v[i].y = fmaf((v[i].y - mean) * rstd, 1.f + s1.y, s0.y);
    v[i].z = fmaf((v[i].z - mean) * rstd, 1.f + s1.z, s0.z);
    v[i].w = fmaf((v[i].w - mean) * rstd, 1.f + s1.w, s0.w);
  }
  const int nout = 4 * a.Cout;
  for (int oidx = 0; oidx < nout; ++oidx) {
    const float* w = a.weight + static_cast<long long>(oidx) * a.D;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (i * 32 + lane) * 4));
      acc = fmaf(v[i].x, ww.x, acc);
      acc = fmaf(v[i].y, ww.y, acc);
      acc = fmaf(v[i].z, ww.z, acc);
      acc = fmaf(v[i].w, ww.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      // unpatchify: feature index = (p * 2 + q) * Cout + c   ('nhwpqc->nchpwq')
      const int c = oidx % a.Cout, pq = oidx / a.Cout, p = pq >> 1, qq = pq & 1;
      a.out[((static_cast<long long>(b) * (3 * a.Cout) + c * 3 + n) * a.S + 2 * pi + p) * a.S +
            2 * pj + qq] = acc + (a.bias ? a.bias[oidx] : 0.f);
    }
  }
}

int final_layer(const ln3_final_layer_args* a, cudaStream_t stream) {
  if (a->B <= 0) return LN3_OK;
  if (a->D % 128 != 0 || a->D > 2048) return set_error(LN3_EINVAL, "final_layer: bad D=%d", a->D);
  if (a->shift == nullptr || a->scale == nullptr)
    return set_error(LN3_EINVAL, "final_layer: shift/scale required");
  const int L = (a->S / 2) * (a->S / 2);
  const int toks = a->B * 3 * L;
  dim3 grid((toks + 3) / 4), block(128);
  if (4 * a->Cout == 16 && (a->D == 768 || a->D == 1024)) {
    dim3 grid2((toks + 7) / 8);
    if (a->D == 1024) final_layer2_kernel<8><<<grid2, block, 0, stream>>>(*a);
    else final_layer2_kernel<6><<<grid2, block, 0, stream>>>(*a);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return set_error(LN3_ECUDA, "final_layer launch: %s", cudaGetErrorString(e2));
    count_launch();
    return LN3_OK;
  }
  switch (a->D / 128) {
#define LN3_FL_CASE(n) \
  case n: final_layer_kernel<n><<<grid, block, 0, stream>>>(*a); break;
    LN3_FL_CASE(1) LN3_FL_CASE(2) LN3_FL_CASE(3) LN3_FL_CASE(4) LN3_FL_CASE(5) LN3_FL_CASE(6)
    LN3_FL_CASE(7) LN3_FL_CASE(8) LN3_FL_CASE(9) LN3_FL_CASE(10) LN3_FL_CASE(11) LN3_FL_CASE(12)
    LN3_FL_CASE(13) LN3_FL_CASE(14) LN3_FL_CASE(15) LN3_FL_CASE(16)
#undef LN3_FL_CASE
    default: return set_error(LN3_EINVAL, "final_layer: unsupported D");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "final_layer launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

// ------------------------------------------------------------------ sampler update
__global__ void __launch_bounds__(256)
sampler_update_kernel(const ln3_sampler_update_args a) {
  const int b = blockIdx.y;
  const float4 cf = *reinterpret_cast<const float4*>(a.coef + b * 4);
  const long long base = static_cast<long long>(b) * a.n_per_sample;
  const long long n4 = a.n_per_sample >> 2;
  for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long off = base + i * 4;
    const float4 x = *reinterpret_cast<const float4*>(a.x + off);
    const float4 m0 = *reinterpret_cast<const float4*>(a.m0 + off);
    float4 r = make_float4(cf.x * x.x, cf.x * x.y, cf.x * x.z, cf.x * x.w);
    r.x = fmaf(cf.y, m0.x, r.x); r.y = fmaf(cf.y, m0.y, r.y);
    r.z = fmaf(cf.y, m0.z, r.z); r.w = fmaf(cf.y, m0.w, r.w);
    if (a.m1 != nullptr) {
      const float4 m1 = *reinterpret_cast<const float4*>(a.m1 + off);
      r.x = fmaf(cf.z, m1.x, r.x); r.y = fmaf(cf.z, m1.y, r.y);
      r.z = fmaf(cf.z, m1.z, r.z); r.w = fmaf(cf.z, m1.w, r.w);
    }
    if (a.noise != nullptr) {
      const float4 nz = *reinterpret_cast<const float4*>(a.noise + off);
      r.x = fmaf(cf.w, nz.x, r.x); r.y = fmaf(cf.w, nz.y, r.y);
      r.z = fmaf(cf.w, nz.z, r.z); r.w = fmaf(cf.w, nz.w, r.w);
    }
    *reinterpret_cast<float4*>(a.x_out + off) = r;
  }
}

int sampler_affine_update(const ln3_sampler_update_args* a, cudaStream_t stream) {
  if (a->B <= 0 || a->n_per_sample <= 0) return LN3_OK;
  if (a->n_per_sample % 4) return set_error(LN3_EINVAL, "sampler_update: n_per_sample % 4 != 0");
  if (!a->x || !a->m0 || !a->coef || !a->x_out) return set_error(LN3_EINVAL, "sampler_update: null pointer");
  const long long n4 = a->n_per_sample / 4;
  int gx = static_cast<int>((n4 + 255) / 256);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, a->B);
  sampler_update_kernel<<<grid, 256, 0, stream>>>(*a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "sampler_update launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

}  // namespace ln3
