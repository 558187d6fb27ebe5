// Shared device/host helpers for libshb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/shb200.h"

namespace shb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// SMs the persistent (one CTA per SM, all shared memory) kernels may occupy.  Data-parallel runs lower it by a few SMs so
// that NCCL's all-reduce CTAs can run BESIDE the backward kernels instead of waiting for one of them to retire
// (shb_set_persistent_sms; default: all 148).
int persistent_sms();

#define SHB_LAUNCH_CHECK()                           \
  do {                                               \
    cudaError_t e__ = cudaPeekAtLastError();         \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

// ---------------------------------------------------------------- storage-type traits (fp32 accumulate always)
template <typename T> struct Io;

template <> struct Io<float> {
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  // 8 contiguous elements, 16-byte aligned
  static __device__ __forceinline__ void ld8(const float* p, float* o) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void ld4(const float* p, float* o) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  }
  static __device__ __forceinline__ void st4(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <> struct Io<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* o) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void ld4(const __nv_bfloat16* p, float* o) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    o[0] = __uint_as_float(u.x << 16); o[1] = __uint_as_float(u.x & 0xffff0000u);
    o[2] = __uint_as_float(u.y << 16); o[3] = __uint_as_float(u.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void st4(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// ---------------------------------------------------------------- activations (models.py:19-32)
__device__ __forceinline__ float act_fwd(float v, int act) {
  switch (act) {
    case SHB_ACT_RELU: return v > 0.f ? v : 0.f;
    case SHB_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case SHB_ACT_LEAKY_RELU: return v > 0.f ? v : 0.02f * v;
    case SHB_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case SHB_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// Epilogue version: ELU / sigmoid through ex2.approx (absolute error ~2e-7 on outputs of O(1): far inside the 1e-4 budget of
// the fp32 mode).  expm1f / expf cost ~40 dependent instructions per value, which made the epilogue warps the bottleneck.
__device__ __forceinline__ float act_fwd_fast(float v, int act) {
  switch (act) {
    case SHB_ACT_RELU: return v > 0.f ? v : 0.f;
    case SHB_ACT_ELU: return v > 0.f ? v : __expf(v) - 1.f;
    case SHB_ACT_LEAKY_RELU: return v > 0.f ? v : 0.02f * v;
    case SHB_ACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-v));
    case SHB_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
// Eight values at once with the activation switch OUTSIDE the element loop: inside it the compiler if-converts the switch and
// evaluates every case (tanhf included) for every value -- measured 190 cycles per value in the conv epilogue.
template <int ACT> __device__ __forceinline__ void act_fwd8_as(float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = act_fwd_fast(v[i], ACT);
}
__device__ __forceinline__ void act_fwd8(float* v, int act) {
  switch (act) {
    case SHB_ACT_RELU: act_fwd8_as<SHB_ACT_RELU>(v); break;
    case SHB_ACT_ELU: act_fwd8_as<SHB_ACT_ELU>(v); break;
    case SHB_ACT_LEAKY_RELU: act_fwd8_as<SHB_ACT_LEAKY_RELU>(v); break;
    case SHB_ACT_SIGMOID: act_fwd8_as<SHB_ACT_SIGMOID>(v); break;
    case SHB_ACT_TANH: act_fwd8_as<SHB_ACT_TANH>(v); break;
    default: break;
  }
}
__device__ __forceinline__ float act_bwd_from_out(float y, int act);
template <int ACT> __device__ __forceinline__ void act_bwd8_as(float* v, const float* y) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= act_bwd_from_out(y[i], ACT);
}
// v[i] *= act'(y[i]) (derivative through the output)
__device__ __forceinline__ void act_bwd8(float* v, const float* y, int act) {
  switch (act) {
    case SHB_ACT_RELU: act_bwd8_as<SHB_ACT_RELU>(v, y); break;
    case SHB_ACT_ELU: act_bwd8_as<SHB_ACT_ELU>(v, y); break;
    case SHB_ACT_LEAKY_RELU: act_bwd8_as<SHB_ACT_LEAKY_RELU>(v, y); break;
    case SHB_ACT_SIGMOID: act_bwd8_as<SHB_ACT_SIGMOID>(v, y); break;
    case SHB_ACT_TANH: act_bwd8_as<SHB_ACT_TANH>(v, y); break;
    default: break;
  }
}
// derivative expressed through the OUTPUT y = act(v)
__device__ __forceinline__ float act_bwd_from_out(float y, int act) {
  switch (act) {
    case SHB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case SHB_ACT_ELU: return fminf(y, 0.f) + 1.f;   // == (y > 0 ? 1 : y + 1) in two instructions instead of three
    case SHB_ACT_LEAKY_RELU: return y > 0.f ? 1.f : 0.02f;
    case SHB_ACT_SIGMOID: return y * (1.f - y);
    case SHB_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in a fixed order (warp shuffles, then warp 0 over the per-warp partials). All threads get it.
template <int NT> __device__ __forceinline__ float block_sum(float v, float* smem /* >= NT/32 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float r = (lane < NT / 32) ? smem[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace shb
