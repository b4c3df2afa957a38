// Frame sink: rendered fp32 views -> packed uint8 video frames on the device.
// Replaces the per-view host loop of TrainLoopDiffusionWithRec.render_video_given_triplane
// (nsr/train_util_diffusion.py:292-376): depth normalised by its per-view min/max, colour-mapped
// (matplotlib `plt.cm.viridis`, a 256-entry table), concatenated to the right of the RGB image,
// HWC, v*127.5+127.5 clipped to [0,255] and truncated to uint8.  HBM-bound: 12(+4) B read and
// 3(+3) B written per pixel; one float4 load per channel and three 32-bit stores per thread.
#include "ln3_internal.h"

namespace ln3 {
namespace {

__device__ __forceinline__ void atomic_min_f32(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

__global__ void minmax_init_kernel(float* ws, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    ws[2 * i] = __int_as_float(0x7f800000);      // +inf
    ws[2 * i + 1] = __int_as_float(0xff800000);  // -inf
  }
}

// grid (chunks, N): per-view min / max of depth [N, HW] (torch .min() / .max() over one view, :300-301)
__global__ void __launch_bounds__(256) depth_minmax_kernel(const float* __restrict__ depth, float* ws, int HW) {
  const int n = blockIdx.y;
  const float4* d4 = reinterpret_cast<const float4*>(depth + static_cast<size_t>(n) * HW);
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW / 4; i += gridDim.x * blockDim.x) {
    float4 v = __ldg(d4 + i);
    mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[8], smx[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; }
  __syncthreads();
  if (w == 0) {
    mn = l < 8 ? smn[l] : __int_as_float(0x7f800000);
    mx = l < 8 ? smx[l] : __int_as_float(0xff800000);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (l == 0) {
      atomic_min_f32(ws + 2 * n, mn);
      atomic_max_f32(ws + 2 * n + 1, mx);
    }
  }
}

// numpy: (v * 127.5 + 127.5).clip(0, 255).astype(uint8).  The video frame (:340-345,366-368) is the cat of the
// fp32 image with the float64 colour-mapped depth, which promotes it to float64 (F64 = true); the image-only
// path (:350-353, the per-view JPEG) stays a float32 array: two separately rounded fp32 operations, no FMA.
template <bool F64>
__device__ __forceinline__ unsigned to_u8(float v) {
  if (F64) {
    double d = static_cast<double>(v) * 127.5 + 127.5;
    d = d < 0.0 ? 0.0 : (d > 255.0 ? 255.0 : d);
    return static_cast<unsigned>(static_cast<int>(d));   // NaN -> 0
  }
  float f = __fadd_rn(__fmul_rn(v, 127.5f), 127.5f);
  f = f < 0.f ? 0.f : (f > 255.f ? 255.f : f);
  return static_cast<unsigned>(static_cast<int>(f));
}

// One thread = 4 horizontally adjacent output pixels = 12 bytes.
template <bool F64>
__global__ void __launch_bounds__(256)
pack_frames_kernel(const float* __restrict__ image, const float* __restrict__ depth,
                   const unsigned char* __restrict__ lut, const float* __restrict__ ws,
                   unsigned char* __restrict__ out, int N, int H, int W, int Wout) {
  __shared__ unsigned char s_lut[768];
  if (depth != nullptr)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const int q_per_row = Wout / 4;
  const long long total = static_cast<long long>(N) * H * q_per_row;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(t % q_per_row);
    const long long row = t / q_per_row;
    const int y = static_cast<int>(row % H);
    const int n = static_cast<int>(row / H);
    const int x = q * 4;
    unsigned b[12];
    if (x < W) {
      const size_t base = (static_cast<size_t>(n) * 3 * H + y) * W + x;
      const size_t cs = static_cast<size_t>(H) * W;
      const float4 r = __ldg(reinterpret_cast<const float4*>(image + base));
      const float4 g = __ldg(reinterpret_cast<const float4*>(image + base + cs));
      const float4 bl = __ldg(reinterpret_cast<const float4*>(image + base + 2 * cs));
      b[0] = to_u8<F64>(r.x); b[1] = to_u8<F64>(g.x); b[2] = to_u8<F64>(bl.x);
      b[3] = to_u8<F64>(r.y); b[4] = to_u8<F64>(g.y); b[5] = to_u8<F64>(bl.y);
      b[6] = to_u8<F64>(r.z); b[7] = to_u8<F64>(g.z); b[8] = to_u8<F64>(bl.z);
      b[9] = to_u8<F64>(r.w); b[10] = to_u8<F64>(g.w); b[11] = to_u8<F64>(bl.w);
    } else {
      const float mn = ws[2 * n], mx = ws[2 * n + 1];
      const float range = __fsub_rn(mx, mn);
      const float4 d = __ldg(reinterpret_cast<const float4*>(depth + (static_cast<size_t>(n) * H + y) * W + (x - W)));
      const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // (d - min) / (max - min) in fp32 (:300-301), then matplotlib Colormap.__call__ on a float array:
        // xa = x * 256 (fp32); xa == 256 -> 255; trunc; NaN (max == min) -> the "bad" colour (0,0,0,0)
        float xa = __fmul_rn(__fdiv_rn(__fsub_rn(dv[i], mn), range), 256.0f);
        if (xa != xa) { b[3 * i] = b[3 * i + 1] = b[3 * i + 2] = 0; continue; }
        int idx = xa >= 256.0f ? 255 : static_cast<int>(xa);
        idx = idx < 0 ? 0 : idx;
        b[3 * i] = s_lut[3 * idx]; b[3 * i + 1] = s_lut[3 * idx + 1]; b[3 * i + 2] = s_lut[3 * idx + 2];
      }
    }
    unsigned* o = reinterpret_cast<unsigned*>(out + ((static_cast<size_t>(n) * H + y) * Wout + x) * 3);
    o[0] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    o[1] = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    o[2] = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
  }
}

}  // namespace

int pack_frames(const ln3_pack_frames_args* a, cudaStream_t stream) {
  if (!a->image || !a->out) return set_error(LN3_EINVAL, "pack_frames: null pointer");
  if (a->N <= 0 || a->H <= 0 || a->W <= 0 || a->W % 4 != 0)
    return set_error(LN3_EINVAL, "pack_frames: need N,H > 0 and W %% 4 == 0 (got N=%d H=%d W=%d)", a->N, a->H, a->W);
  if ((reinterpret_cast<uintptr_t>(a->image) | reinterpret_cast<uintptr_t>(a->depth)) & 15)
    return set_error(LN3_EINVAL, "pack_frames: image / depth must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(a->out) & 3) return set_error(LN3_EINVAL, "pack_frames: out must be 4-byte aligned");
  if (a->depth && (!a->lut || !a->workspace))
    return set_error(LN3_EINVAL, "pack_frames: depth needs the colormap byte table and a 2*N float workspace");
  const int Wout = a->depth ? 2 * a->W : a->W;
  int launches = 1;
  if (a->depth) {
    const int HW = a->H * a->W;
    minmax_init_kernel<<<(a->N + 255) / 256, 256, 0, stream>>>(a->workspace, a->N);
    int chunks = (HW / 4 + 255) / 256;
    chunks = chunks > 16 ? 16 : (chunks < 1 ? 1 : chunks);
    depth_minmax_kernel<<<dim3(chunks, a->N), 256, 0, stream>>>(a->depth, a->workspace, HW);
    launches = 3;
  }
  const long long total = static_cast<long long>(a->N) * a->H * (Wout / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (a->depth)
    pack_frames_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a->image, a->depth, a->lut, a->workspace,
                                                                                a->out, a->N, a->H, a->W, Wout);
  else
    pack_frames_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a->image, nullptr, nullptr, nullptr, a->out,
                                                                                 a->N, a->H, a->W, Wout);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(LN3_ECUDA, "pack_frames launch: %s", cudaGetErrorString(e));
  count_launch(launches);
  return LN3_OK;
}

}  // namespace ln3
