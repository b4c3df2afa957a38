// Fused multi-head attention forward for sm_100a (head_dim 64, bf16 in, fp32 softmax/accumulate).
//
// Replaces xformers.ops.memory_efficient_attention at its three call sites on the path:
//   vit/vision_transformer.py:114-118 (DiT self-attention, (B, N, 3, H, 64) packed qkv),
//   ldm/modules/attention.py:279-307  (cross-attention; the reference's three permute+contiguous
//                                       copies disappear: heads are addressed through the TMA map),
//   dit/dit_decoder.py                (in-plane / global attention of the DiT2 VAE decoder).
// Semantics: out = softmax(q k^T * scale) v, no mask (xformers FMHA with attn_bias=None).
//
// One CTA = 128 query rows of one (batch, head); 128 threads, thread r owns query row r (TMEM lane r),
// so the online softmax needs no cross-thread reduction.
//   S = Q K^T   : tcgen05.mma 128x128x16 x4, Q/K tiles K-major in 128B-swizzled smem (TMA)
//   P = exp2(..): TMEM -> registers -> bf16 -> smem (same swizzle, written by the owning thread)
//   O_j = P V   : tcgen05.mma 128x64x16 x8, V tile MN-major straight from the TMA layout
//   O   = O * alpha + O_j in registers.
// K/V tiles are double buffered; QK^T of block j+1 is issued right behind PV of block j so the
// tensor pipe works while the CUDA cores rescale O; two CTAs per SM interleave softmax and MMA.
#include "common.cuh"
#include "ln3_internal.h"

namespace ln3 {

static constexpr int kQT = 128;   // query rows per CTA
static constexpr int kKT = 128;   // kv rows per block
static constexpr int kHD = 64;    // head dim
static constexpr int kTileBytes = 128 * kHD * 2;  // 16 KB
static constexpr int kFmhaSmem = 1024 + kTileBytes * (1 + 2 + 2 + 2) + 128;
static constexpr int kFmhaTmemCols = 256;  // S: [0,128)  O_j: [128,192)

struct FmhaParams {
  int Lq, Lkv;
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  long long out_ld, out_bs;  // row / batch stride (elements); head h at column h*64
};

__global__ void __launch_bounds__(128, 2)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;      // [2]
  uint8_t* sV = sK + 2 * kTileBytes;  // [2]
  uint8_t* sP = sV + 2 * kTileBytes;  // two 64-column atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;  // [2]
  uint64_t* v_full = bars + 3;  // [2]
  uint64_t* s_done = bars + 5;
  uint64_t* o_done = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int q0 = blockIdx.x * kQT;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int nkv = (p.Lkv + kKT - 1) / kKT;

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kFmhaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;

  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major

  auto load_kv = [&](int j) {
    const int b = j & 1;
    mbar_arrive_expect_tx(&k_full[b], kTileBytes);
    tma_load_3d(sK + b * kTileBytes, &tmap_k, &k_full[b], head * kHD, j * kKT, batch);
    mbar_arrive_expect_tx(&v_full[b], kTileBytes);
    tma_load_3d(sV + b * kTileBytes, &tmap_v, &v_full[b], head * kHD, j * kKT, batch);
  };
  auto issue_qk = [&](int j) {
    const int b = j & 1;
    mbar_wait(&k_full[b], (j >> 1) & 1);
    tc_fence_after();
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + b * kTileBytes);
#pragma unroll
    for (int k = 0; k < kHD / 16; ++k)
      umma_f16_ss(tS, make_smem_desc_sw128(qa + k * 32, 0, 1024),
                  make_smem_desc_sw128(ka + k * 32, 0, 1024), idesc_s, k != 0);
    umma_commit(s_done);
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(q_full, kTileBytes);
    tma_load_3d(sQ, &tmap_q, q_full, head * kHD, q0, batch);
    load_kv(0);
    if (nkv > 1) load_kv(1);
    mbar_wait(q_full, 0);
    issue_qk(0);
  }

  float m_run = -INFINITY, l_run = 0.f;
  float o[kHD];
#pragma unroll
  for (int i = 0; i < kHD; ++i) o[i] = 0.f;

  const int row = tid;  // == TMEM lane
  const uint32_t p_row = smem_u32(sP) + row * 128;
  const int swz = row & 7;

  for (int j = 0; j < nkv; ++j) {
    const int kv_valid = min(kKT, p.Lkv - j * kKT);
    mbar_wait(s_done, j & 1);
    tc_fence_after();

    // pass 1: row max (log2 domain)
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < kKT; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tS + lane_off + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float s = (c + i < kv_valid) ? __uint_as_float(v[i]) : -INFINITY;
        mx = fmaxf(mx, s);
      }
    }
    const float m_new = fmaxf(m_run, mx * p.scale_log2);
    const float alpha = exp2f(m_run - m_new);  // 0 on the first block (m_run = -inf)
    // pass 2: p = exp2(s*scale - m_new) -> bf16 -> swizzled smem, row sum
    float rs = 0.f;
#pragma unroll 1
    for (int c = 0; c < kKT; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tS + lane_off + c, v);
      tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new));
        pv[i] = (c + i < kv_valid) ? e : 0.f;
      }
      const uint32_t atom = p_row + (c >> 6) * kTileBytes;
      const int chunk0 = (c & 63) >> 3;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]);
        const uint32_t y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
        const uint32_t z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]);
        const uint32_t w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
        // row sum of the values the tensor core will actually see (bf16-rounded)
        __nv_bfloat162 bx = *reinterpret_cast<const __nv_bfloat162*>(&x);
        __nv_bfloat162 by = *reinterpret_cast<const __nv_bfloat162*>(&y);
        __nv_bfloat162 bz = *reinterpret_cast<const __nv_bfloat162*>(&z);
        __nv_bfloat162 bw = *reinterpret_cast<const __nv_bfloat162*>(&w);
        rs += (__low2float(bx) + __high2float(bx)) + (__low2float(by) + __high2float(by)) +
              (__low2float(bz) + __high2float(bz)) + (__low2float(bw) + __high2float(bw));
        const uint32_t addr = atom + (((chunk0 + g) ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z),
                     "r"(w)
                     : "memory");
      }
    }
    l_run = l_run * alpha + rs;
    m_run = m_new;

    fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const int b = j & 1;
      mbar_wait(&v_full[b], (j >> 1) & 1);
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV + b * kTileBytes);
#pragma unroll
      for (int k = 0; k < kKT / 16; ++k)
        umma_f16_ss(tO, make_smem_desc_sw128(pa + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(va + k * 16 * 128, 1024, 1024), idesc_o, k != 0);
      umma_commit(o_done);
      if (j + 1 < nkv) issue_qk(j + 1);  // S is free: every thread finished pass 2 before the sync
    }
    mbar_wait(o_done, j & 1);
    tc_fence_after();
    if (tid == 0 && j + 2 < nkv) load_kv(j + 2);  // K/V buffer (j&1) is free once PV(j) retired
#pragma unroll
    for (int c = 0; c < kHD; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_off + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(v[i]));
    }
  }

  const float inv = 1.f / l_run;
  if (q0 + row < p.Lq) {
    __nv_bfloat16* dst = p.out + batch * p.out_bs + static_cast<long long>(q0 + row) * p.out_ld +
                         head * kHD;
#pragma unroll
    for (int i = 0; i < kHD; i += 8) {
      uint4 q;
      q.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
      q.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
      q.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
      q.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
      *reinterpret_cast<uint4*>(dst + i) = q;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, kFmhaTmemCols);
}

int fmha_fwd(const ln3_fmha_args* a, cudaStream_t stream) {
  if (a->head_dim != kHD) return set_error(LN3_EUNSUPPORTED, "fmha: head_dim must be 64");
  if (a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lkv <= 0)
    return set_error(LN3_EINVAL, "fmha: empty problem");
  if ((a->q_ld | a->k_ld | a->v_ld | a->o_ld | a->q_bs | a->k_bs | a->v_bs | a->o_bs) % 8)
    return set_error(LN3_EINVAL, "fmha: strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
       reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) & 15)
    return set_error(LN3_EINVAL, "fmha: pointers must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fmha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kFmhaSmem);
    if (e != cudaSuccess)
      return set_error(LN3_ECUDA, "fmha: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  int rc;
  const long long cols = static_cast<long long>(a->H) * kHD;
  if ((rc = make_tmap_3d_bf16(&tq, a->q, cols, a->Lq, a->B, a->q_ld, a->q_bs, kHD, kQT))) return rc;
  if ((rc = make_tmap_3d_bf16(&tk, a->k, cols, a->Lkv, a->B, a->k_ld, a->k_bs, kHD, kKT))) return rc;
  if ((rc = make_tmap_3d_bf16(&tv, a->v, cols, a->Lkv, a->B, a->v_ld, a->v_bs, kHD, kKT))) return rc;
  FmhaParams p;
  p.Lq = a->Lq;
  p.Lkv = a->Lkv;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.out_ld = a->o_ld;
  p.out_bs = a->o_bs;
  dim3 grid((a->Lq + kQT - 1) / kQT, a->H, a->B);
  fmha_fwd_kernel<<<grid, 128, kFmhaSmem, stream>>>(tq, tk, tv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "fmha launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

}  // namespace ln3
