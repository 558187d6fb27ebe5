// Adam (main.py:262: torch.optim.Adam(lr, weight_decay) -- L2 weight decay folded into the gradient, bias-corrected
// moments) as ONE multi-tensor kernel that also emits the bf16 shadow of the weights that the bf16 mode feeds to the FC
// GEMMs, so no stand-alone weight cast (and no cast of its gradient) runs in the step.
//
// HBM-bound by construction: per parameter 16 B read (p, g, m, v) + 12 B written (p, m, v) + 2 B of shadow.  The step
// counter lives on the device (bias corrections are computed from it in-kernel), so the launch is CUDA-graph replayable.
#include "shb_common.cuh"
#include "shb_internal.h"

namespace shb {

constexpr int AD_MAX_TENSORS = 32;
constexpr int AD_CHUNK = 8192;   // elements per CTA-iteration

struct AdamTable {
  float* p[AD_MAX_TENSORS];
  const void* g[AD_MAX_TENSORS];           // fp32, or bf16 where g_bf16[i] (a gradient bucket reduced in bf16)
  float* m[AD_MAX_TENSORS];
  float* v[AD_MAX_TENSORS];
  __nv_bfloat16* shadow[AD_MAX_TENSORS];   // or null
  long long n[AD_MAX_TENSORS];
  unsigned char g_bf16[AD_MAX_TENSORS];
  int chunk_start[AD_MAX_TENSORS + 1];     // prefix sum of ceil(n / AD_CHUNK)
  int count;
};

__global__ void adam_tick_kernel(float* step) { step[0] += 1.f; }

__global__ void __launch_bounds__(256) adam_kernel(const AdamTable t, const float* __restrict__ step, double lr_d, double b1_d,
                                                   double b2_d, float eps, float wd) {
  // torch.optim.Adam: step_size = lr / (1 - b1^t); denom = sqrt(v) / sqrt(1 - b2^t) + eps.  torch evaluates the bias
  // corrections in Python doubles; in fp32, 1 - 0.999f^t loses five digits for small t (0.999f != 0.999), so: doubles here too
  __shared__ float bc_s[2];
  if (threadIdx.x == 0) {
    const double tt = (double)__ldg(step);
    bc_s[0] = (float)(lr_d / (1.0 - pow(b1_d, tt)));
    bc_s[1] = (float)sqrt(1.0 - pow(b2_d, tt));
  }
  __syncthreads();
  const float step_size = bc_s[0], bc2_sqrt = bc_s[1];
  // 1 - beta in double, then rounded (as torch's Python scalars): 1.f - 0.999f is 1.3e-5 away from 0.001f
  const float b2 = (float)b2_d, omb1 = (float)(1.0 - b1_d), omb2 = (float)(1.0 - b2_d);
  const int total_chunks = t.chunk_start[t.count];
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    int ti = 0;
    while (ti + 1 < t.count && c >= t.chunk_start[ti + 1]) ++ti;
    const long long off = (long long)(c - t.chunk_start[ti]) * AD_CHUNK;
    const long long n = t.n[ti] - off < AD_CHUNK ? t.n[ti] - off : AD_CHUNK;
    float* p = t.p[ti] + off;
    const bool gb16 = t.g_bf16[ti] != 0;
    const float* g = reinterpret_cast<const float*>(t.g[ti]) + (gb16 ? 0 : off);
    const __nv_bfloat16* gh = reinterpret_cast<const __nv_bfloat16*>(t.g[ti]) + off;
    float* m = t.m[ti] + off;
    float* v = t.v[ti] + off;
    __nv_bfloat16* sh = t.shadow[ti] != nullptr ? t.shadow[ti] + off : nullptr;
    const long long n4 = n & ~3LL;
    for (long long i = (long long)threadIdx.x * 4; i < n4; i += 256 * 4) {
      float4 pp = *reinterpret_cast<const float4*>(p + i);
      float4 gg;
      if (gb16) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(gh + i));
        gg = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                         __uint_as_float(raw.y & 0xffff0000u));
      } else {
        gg = __ldg(reinterpret_cast<const float4*>(g + i));
      }
      float4 mm = *reinterpret_cast<const float4*>(m + i);
      float4 vv = *reinterpret_cast<const float4*>(v + i);
      float* pa = reinterpret_cast<float*>(&pp);
      const float* ga = reinterpret_cast<const float*>(&gg);
      float* ma = reinterpret_cast<float*>(&mm);
      float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = ga[k] + wd * pa[k];
        ma[k] = ma[k] + omb1 * (gk - ma[k]);          // lerp form, as torch: exp_avg.lerp_(grad, 1 - beta1)
        va[k] = b2 * va[k] + omb2 * gk * gk;
        const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
        pa[k] = pa[k] - step_size * (ma[k] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (sh != nullptr) Io<__nv_bfloat16>::st4(sh + i, pa);
    }
    for (long long i = n4 + threadIdx.x; i < n; i += 256) {
      const float gk = (gb16 ? __bfloat162float(gh[i]) : g[i]) + wd * p[i];
      const float mk = m[i] + omb1 * (gk - m[i]);
      const float vk = b2 * v[i] + omb2 * gk * gk;
      const float pk = p[i] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
      p[i] = pk; m[i] = mk; v[i] = vk;
      if (sh != nullptr) sh[i] = __float2bfloat16_rn(pk);
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * 256 * 4) {
    if (i + 4 <= n) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(src + i));
      const float a[4] = {x.x, x.y, x.z, x.w};
      Io<__nv_bfloat16>::st4(dst + i, a);
    } else {
      for (long long k = i; k < n; ++k) dst[k] = __float2bfloat16_rn(src[k]);
    }
  }
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb_adam_tick(float* step, void* stream) {
  if (!step) return SHB_E_ARG;
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_adam_step_mixed(int count, float* const* p, const void* const* g, const uint8_t* g_is_bf16, float* const* m,
                        float* const* v, void* const* shadow, const int64_t* numel, const float* step, double lr, double beta1,
                        double beta2, float eps, float weight_decay, void* stream) {
  if (count <= 0 || !p || !g || !m || !v || !numel || !step) return SHB_E_ARG;
  for (int base = 0; base < count; base += AD_MAX_TENSORS) {
    AdamTable t{};
    t.count = count - base < AD_MAX_TENSORS ? count - base : AD_MAX_TENSORS;
    t.chunk_start[0] = 0;
    for (int i = 0; i < t.count; ++i) {
      const int k = base + i;
      if (!p[k] || !g[k] || !m[k] || !v[k] || numel[k] <= 0) return SHB_E_ARG;
      if ((((uintptr_t)p[k] | (uintptr_t)g[k] | (uintptr_t)m[k] | (uintptr_t)v[k]) & 15) != 0) return SHB_E_ARG;
      t.p[i] = p[k]; t.g[i] = g[k]; t.m[i] = m[k]; t.v[i] = v[k];
      t.g_bf16[i] = (g_is_bf16 != nullptr && g_is_bf16[k]) ? 1 : 0;
      t.shadow[i] = shadow ? (__nv_bfloat16*)shadow[k] : nullptr;
      t.n[i] = numel[k];
      t.chunk_start[i + 1] = t.chunk_start[i] + (int)((numel[k] + AD_CHUNK - 1) / AD_CHUNK);
    }
    int grid = t.chunk_start[t.count];
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, step, (double)lr, (double)beta1, (double)beta2, eps, weight_decay);
    SHB_LAUNCH_CHECK();
  }
  return 0;
}

int shb_adam_step(int count, float* const* p, const float* const* g, float* const* m, float* const* v, void* const* shadow,
                  const int64_t* numel, const float* step, double lr, double beta1, double beta2, float eps, float weight_decay,
                  void* stream) {
  return shb_adam_step_mixed(count, p, reinterpret_cast<const void* const*>(g), nullptr, m, v, shadow, numel, step, lr, beta1,
                             beta2, eps, weight_decay, stream);
}

int shb_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst || n <= 0) return SHB_E_ARG;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  cast_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
