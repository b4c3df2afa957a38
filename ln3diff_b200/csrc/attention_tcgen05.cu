// Fused multi-head attention forward for sm_100a (head_dim 64, bf16 in, fp32 softmax/accumulate).
//
// Replaces xformers.ops.memory_efficient_attention at its three call sites on the path:
//   vit/vision_transformer.py:114-118 (DiT self-attention, (B, N, 3, H, 64) packed qkv),
//   ldm/modules/attention.py:279-307  (cross-attention; the reference's three permute+contiguous
//                                       copies disappear: heads are addressed through the TMA map),
//   dit/dit_decoder.py                (in-plane / global attention of the DiT2 VAE decoder).
// Semantics: out = softmax(q k^T * scale) v, no mask (xformers FMHA with attn_bias=None).
//
// Warp-specialised, one CTA = 256 query rows (two 128-row tiles) of one (batch, head):
//   warps 0-3 / 4-7 : softmax warpgroup of tile 0 / 1; thread r owns query row r (= TMEM lane r), so
//                     the row max / sum need no cross-thread reduction.  S (128 fp32) is read once
//                     from TMEM into registers, P = exp2(S*scale - m) goes to 128B-swizzled smem as bf16.
//   warp 8          : TMA producer (Q once, K/V tiles double buffered)
//   warp 9          : tcgen05.mma issuer:  S_t = Q_t K^T (128x128x16 x4),  O_t += P_t V (128x64x16 x8,
//                     V MN-major straight from the TMA layout).  While warpgroup 0 runs its softmax the
//                     tensor core computes S_1 / P_1 V and vice versa.
// O accumulates in TMEM across KV blocks.  The running max is only refreshed when it grew by more
// than 2^8 (then the owning thread rescales its O row in TMEM); otherwise P is computed against the
// stale max, which is exact after the final 1/l normalisation.
#include "common.cuh"
#include "ln3_internal.h"

namespace ln3 {

static constexpr int kQT = 128;   // query rows per tile (2 tiles per CTA)
static constexpr int kKT = 128;   // kv rows per block
static constexpr int kHD = 64;    // head dim
static constexpr int kTileBytes = 128 * kHD * 2;  // 16 KB
// Q[2 buffers][2 tiles] | K[2] | V[2] | P0 (2 atoms) P1 (2 atoms)
static constexpr int kFmhaSmem = 1024 + kTileBytes * (4 + 2 + 2 + 4) + 256;
static constexpr int kFmhaTmemCols = 512;  // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)
static constexpr int kFmhaThreads = 320;
static constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct FmhaParams {
  int Lq, Lkv;
  int Lkv2;  // rows of the optional second K/V source (0 = none)
  int B, H, nq;  // work items = B * H * nq query-row pairs (256 rows each)
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  long long out_ld, out_bs;  // row / batch stride (elements); head h at column h*64
};

// Persistent: each CTA walks work items w = blockIdx.x, +gridDim.x, ... (item = one (batch, head,
// 256-query-row pair)); barriers, TMEM and the K/V ring carry over between items and Q is double
// buffered, so the TMA warp prefetches the next item's Q / K / V while the current item is still in its
// softmax -- the per-CTA prologue (TMEM allocation, barrier init, first-load latency) is paid once per
// SM instead of once per item (it was ~30 % of a self-attention item and most of a cross-attention one).
__global__ void __launch_bounds__(kFmhaThreads, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_k2,
                const __grid_constant__ CUtensorMap tmap_v2, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                  // [2 buffers][2 tiles]
  uint8_t* sK = sQ + 4 * kTileBytes;   // [2]
  uint8_t* sV = sK + 2 * kTileBytes;   // [2]
  uint8_t* sP = sV + 2 * kTileBytes;   // [2 tiles][2 atoms]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kTileBytes);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [2]
  uint64_t* kv_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;     // [2 tiles]
  uint64_t* p_full = bars + 10;    // [2 tiles], 128 arrivals
  uint64_t* o_full = bars + 12;    // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int nkv1 = (p.Lkv + kKT - 1) / kKT;  // with a second source Lkv is a multiple of 128
  const int nkv = nkv1 + (p.Lkv2 + kKT - 1) / kKT;
  const int nitems = p.B * p.H * p.nq;
  auto item_coords = [&](int w, int& q0, int& head, int& batch) {
    const int qp = w % p.nq;
    const int bh = w / p.nq;
    q0 = qp * 2 * kQT;
    head = bh % p.H;
    batch = bh / p.H;
  };

  if (tid == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    if (p.Lkv2 > 0) {
      tma_prefetch_desc(&tmap_k2);
      tma_prefetch_desc(&tmap_v2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, kFmhaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if ((tid & 31) == 0) {
      int it = 0, g = 0;
      for (int w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
        int q0, head, batch;
        item_coords(w, q0, head, batch);
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], 2 * kTileBytes);
        for (int t = 0; t < 2; ++t)
          tma_load_3d(sQ + (qb * 2 + t) * kTileBytes, &tmap_q, &q_full[qb], head * kHD, q0 + t * kQT, batch);
        for (int j = 0; j < nkv; ++j, ++g) {
          const int b = g & 1;
          mbar_wait(&kv_empty[b], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[b], 2 * kTileBytes);
          if (j < nkv1) {
            tma_load_3d(sK + b * kTileBytes, &tmap_k, &kv_full[b], head * kHD, j * kKT, batch);
            tma_load_3d(sV + b * kTileBytes, &tmap_v, &kv_full[b], head * kHD, j * kKT, batch);
          } else {
            tma_load_3d(sK + b * kTileBytes, &tmap_k2, &kv_full[b], head * kHD, (j - nkv1) * kKT, batch);
            tma_load_3d(sV + b * kTileBytes, &tmap_v2, &kv_full[b], head * kHD, (j - nkv1) * kKT, batch);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      auto issue_qk = [&](int t, int qb, int gblk) {
        const uint32_t qa = smem_u32(sQ + (qb * 2 + t) * kTileBytes), ka = smem_u32(sK + (gblk & 1) * kTileBytes);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_f16_ss(tmem_base + t * 128, make_smem_desc_sw128(qa + k * 32, 0, 1024),
                      make_smem_desc_sw128(ka + k * 32, 0, 1024), idesc_s, k != 0);
        umma_commit(&s_full[t]);
      };
      int it = 0, g = 0;
      for (int w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
        const int qb = it & 1;
        const bool has_next_item = w + static_cast<int>(gridDim.x) < nitems;
        if (it == 0) {  // very first block of this CTA: nothing has been issued ahead
          mbar_wait(&q_full[0], 0);
          mbar_wait(&kv_full[0], 0);
          tc_fence_after();
          issue_qk(0, 0, 0);
          issue_qk(1, 0, 0);
        }
        for (int j = 0; j < nkv; ++j, ++g) {
          const bool last = j + 1 == nkv;
          const bool next_exists = !last || has_next_item;
          const int next_qb = last ? ((it + 1) & 1) : qb;
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&p_full[t], g & 1);  // P_t in smem, S_t consumed, O_t rescaled if needed
            tc_fence_after();
            const uint32_t pa = smem_u32(sP + t * 2 * kTileBytes);
            const uint32_t va = smem_u32(sV + (g & 1) * kTileBytes);
#pragma unroll
            for (int k = 0; k < kKT / 16; ++k)
              umma_f16_ss(tmem_base + 256 + t * 64,
                          make_smem_desc_sw128(pa + (k >> 2) * kTileBytes + (k & 3) * 32, 0, 1024),
                          make_smem_desc_sw128(va + k * 16 * 128, 1024, 1024), idesc_o, (j | k) != 0);
            umma_commit(&o_full[t]);
            if (next_exists) {  // S of the next block (possibly the first block of the next item)
              if (t == 0) {
                if (last) mbar_wait(&q_full[next_qb], ((it + 1) >> 1) & 1);
                mbar_wait(&kv_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
                tc_fence_after();
              }
              issue_qk(t, next_qb, g + 1);
            }
          }
          umma_commit(&kv_empty[g & 1]);  // every MMA that read K / V of block g has been issued
          if (last) umma_commit(&q_empty[qb]);  // ... and every QK of this item
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ softmax warpgroups
    const int t = warp >> 2;
    const int row = tid & 127;  // TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 64 + lane_off;
    const uint32_t p_row = smem_u32(sP + t * 2 * kTileBytes) + row * 128;
    const int swz = row & 7;
    int g = 0;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
      int q0, head, batch;
      item_coords(w, q0, head, batch);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nkv; ++j, ++g) {
        const int kv_valid = (j < nkv1) ? p.Lkv - j * kKT : p.Lkv2 - (j - nkv1) * kKT;  // >= 1
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        uint32_t s[128];
        tmem_ld_32x32(tS + 0, s);
        tmem_ld_32x32(tS + 32, s + 32);
        tmem_ld_32x32(tS + 64, s + 64);
        tmem_ld_32x32(tS + 96, s + 96);
        tmem_ld_wait();
        if (kv_valid < kKT) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
        }
        float mx = fmax3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
#pragma unroll
        for (int i = 3; i < 127; i += 2) mx = fmax3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mx = fmaxf(mx, __uint_as_float(s[127]));
        const float m_cand = mx * p.scale_log2;
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_ref = m_cand;
        } else if (m_cand > m_ref + kRescaleThreshold) {
          need = true;
          alpha = fast_exp2(m_ref - m_cand);
          m_ref = m_cand;
          l_run *= alpha;
        }
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = fast_exp2(fmaf(__uint_as_float(s[c + i]), p.scale_log2, -m_ref));
          rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          const uint32_t addr = p_row + (c >> 6) * kTileBytes + ((((c & 63) >> 3) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(e[0], e[1])),
                       "r"(pack_bf16x2(e[2], e[3])), "r"(pack_bf16x2(e[4], e[5])),
                       "r"(pack_bf16x2(e[6], e[7]))
                       : "memory");
        }
        l_run += rs;
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // O_t of the previous block must be complete before it is rescaled in place
          mbar_wait(&o_full[t], (g - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kHD; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tO + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32(tO + c, v);
          }
          tmem_st_wait();
        }
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core
        tc_fence_before();
        mbar_arrive(&p_full[t]);
      }
      mbar_wait(&o_full[t], (g - 1) & 1);
      tc_fence_after();
      const float inv = 1.f / l_run;
      const int qrow = q0 + t * kQT + row;
      __nv_bfloat16* dst = p.out + batch * p.out_bs + static_cast<long long>(qrow) * p.out_ld + head * kHD;
#pragma unroll
      for (int c = 0; c < kHD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tO + c, v);
        tmem_ld_wait();
        if (qrow < p.Lq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 q;
            q.x = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
            q.y = pack_bf16x2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
            q.z = pack_bf16x2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
            q.w = pack_bf16x2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c + i) = q;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 9) tmem_dealloc(tmem_base, kFmhaTmemCols);
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
  if (a->k2 != nullptr || a->v2 != nullptr) {
    if (!a->k2 || !a->v2 || a->Lkv2 <= 0) return set_error(LN3_EINVAL, "fmha: k2/v2/Lkv2 must be given together");
    if (a->Lkv % kKT != 0) return set_error(LN3_EINVAL, "fmha: Lkv must be a multiple of 128 with a second K/V source");
    if ((a->k2_ld | a->v2_ld | a->k2_bs | a->v2_bs) % 8 ||
        ((reinterpret_cast<uintptr_t>(a->k2) | reinterpret_cast<uintptr_t>(a->v2)) & 15))
      return set_error(LN3_EINVAL, "fmha: k2/v2 alignment");
  }
  CUtensorMap tq, tk, tv, tk2, tv2;
  int rc;
  const long long cols = static_cast<long long>(a->H) * kHD;
  if ((rc = make_tmap_3d_bf16(&tq, a->q, cols, a->Lq, a->B, a->q_ld, a->q_bs, kHD, kQT))) return rc;
  if ((rc = make_tmap_3d_bf16(&tk, a->k, cols, a->Lkv, a->B, a->k_ld, a->k_bs, kHD, kKT))) return rc;
  if ((rc = make_tmap_3d_bf16(&tv, a->v, cols, a->Lkv, a->B, a->v_ld, a->v_bs, kHD, kKT))) return rc;
  const bool two = a->k2 != nullptr;
  if (two) {
    if ((rc = make_tmap_3d_bf16(&tk2, a->k2, cols, a->Lkv2, a->B, a->k2_ld, a->k2_bs, kHD, kKT))) return rc;
    if ((rc = make_tmap_3d_bf16(&tv2, a->v2, cols, a->Lkv2, a->B, a->v2_ld, a->v2_bs, kHD, kKT))) return rc;
  } else {
    tk2 = tk;
    tv2 = tv;
  }
  FmhaParams p;
  p.Lq = a->Lq;
  p.Lkv = a->Lkv;
  p.Lkv2 = two ? a->Lkv2 : 0;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.out_ld = a->o_ld;
  p.out_bs = a->o_bs;
  p.B = a->B;
  p.H = a->H;
  p.nq = (a->Lq + 2 * kQT - 1) / (2 * kQT);
  const long long nitems = static_cast<long long>(p.B) * p.H * p.nq;
  const int sms = device_sm_count();
  const int grid = static_cast<int>(nitems < sms ? nitems : sms);
  fmha_fwd_kernel<<<grid, kFmhaThreads, kFmhaSmem, stream>>>(tq, tk, tv, tk2, tv2, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "fmha launch: %s", cudaGetErrorString(e));
  count_launch();
  return LN3_OK;
}

}  // namespace ln3
