// Pool (CSR SpMM over batch-major features) and the loss reductions of libshb200.  All HBM-bound:
// coalesced 16-byte accesses along the channel axis, grids sized in multiples of the SM count,
// fixed-order reductions (no float atomics).
#include "shb_common.cuh"

namespace shb {

// ------------------------------------------------------------------------------------------------ Pool SpMM
// One thread per (b, r, V-wide channel vector): V = 16 bytes' worth when C allows (8 bf16 / 4 fp32), else 1.
// 32-bit index arithmetic (the host checks the item count fits).
template <typename T, int V>
__global__ void __launch_bounds__(256) pool_spmm_kernel(const T* __restrict__ x, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                        T* __restrict__ y, unsigned total, unsigned rows_in, unsigned rows_out,
                                                        unsigned CV) {
  // two independent items per thread per iteration: twice the loads in flight (the kernel is latency-, not issue-bound)
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
    unsigned it[2] = {i0, i0 + stride};
    unsigned br[2], cv[2];
    int e0[2], e1[2];
    const T* xb[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool on = it[u] < total;
      const unsigned i = on ? it[u] : 0;
      br[u] = i / CV;
      cv[u] = i - br[u] * CV;
      const unsigned b = br[u] / rows_out, r = br[u] - b * rows_out;
      e0[u] = on ? __ldg(rowptr + r) : 0;
      e1[u] = on ? __ldg(rowptr + r + 1) : 0;
      xb[u] = x + ((size_t)b * rows_in) * (CV * V) + cv[u] * V;
    }
    float acc[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int k = 0; k < V; ++k) acc[u][k] = 0.f;
    const int n0 = e1[0] - e0[0], n1 = e1[1] - e0[1], nmax = n0 > n1 ? n0 : n1;
    for (int t = 0; t < nmax; ++t) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = e0[u] + t;
        if (e < e1[u]) {
          const float a = __ldg(vals + e);
          const T* px = xb[u] + (size_t)__ldg(colidx + e) * (CV * V);
          float v[V];
          if (V == 8) Io<T>::ld8(px, v); else if (V == 4) Io<T>::ld4(px, v); else v[0] = Io<T>::ld(px);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[u][k] = fmaf(a, v[k], acc[u][k]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (it[u] < total) {
        T* py = y + (size_t)br[u] * (CV * V) + cv[u] * V;
        if (V == 8) { Io<T>::st4(py, acc[u]); Io<T>::st4(py + 4, acc[u] + 4); }
        else if (V == 4) Io<T>::st4(py, acc[u]);
        else Io<T>::st(py, acc[u][0]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ L1 loss
constexpr int L1_BLOCKS = 8 * kNumSMs;

template <typename T>
__global__ void __launch_bounds__(256) l1_partial_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n,
                                                         float* __restrict__ partials) {
  __shared__ float red[8];
  const long long stride = (long long)gridDim.x * blockDim.x;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    s += fabsf(Io<T>::ld(a + i) - Io<T>::ld(b + i));
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) l1_final_kernel(const float* __restrict__ partials, int np, float inv_n,
                                                       float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < np; i += 256) s += partials[i];
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) *out = s * inv_n;
}

template <typename T>
__global__ void __launch_bounds__(256) l1_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n,
                                                     const float* __restrict__ gscale, T* __restrict__ ga, T* __restrict__ gb) {
  const float g = __ldg(gscale) / (float)n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = Io<T>::ld(a + i) - Io<T>::ld(b + i);
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    if (ga) Io<T>::st(ga + i, g * sgn);
    if (gb) Io<T>::st(gb + i, -g * sgn);
  }
}

// ------------------------------------------------------------------------------------------------ part-norm loss
// One CTA: (b, i) terms strided over threads, fixed-order block reduction.  gz is zero-filled first.
__global__ void __launch_bounds__(256) partnorm_kernel(const float* __restrict__ z, const float* __restrict__ measure,
                                                       const int32_t* __restrict__ P, const int32_t* __restrict__ Q,
                                                       float* __restrict__ loss_out, float* __restrict__ gz, int B, int n_parts,
                                                       int L, int n_measure, int n_sel, int relative) {
  __shared__ float red[8];
  const long long nz = (long long)B * n_parts * L;
  for (long long i = threadIdx.x; i < nz; i += 256) gz[i] = 0.f;
  __syncthreads();
  const int terms = B * n_sel;
  const float inv = 1.f / (float)terms;
  float s = 0.f;
  for (int q = threadIdx.x; q < terms; q += 256) {
    const int b = q / n_sel, i = q - b * n_sel;
    const float* zp = z + ((long long)b * n_parts + P[i]) * L;
    float ss = 0.f;
    for (int l = 0; l < L; ++l) ss = fmaf(zp[l], zp[l], ss);
    const float m = sqrtf(ss);
    const float tgt = measure[(long long)b * n_measure + Q[i]];
    const float d = relative ? (m / tgt - 1.f) : (m - tgt);
    s += fabsf(d);
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    // d|d|/dz = sgn * (1/tgt or 1) * z/m ; sqrt'(0) -> inf in torch, guarded to 0 here
    const float coef = (m > 0.f) ? sgn * inv * (relative ? 1.f / tgt : 1.f) / m : 0.f;
    float* gp = gz + ((long long)b * n_parts + P[i]) * L;
    for (int l = 0; l < L; ++l) gp[l] = coef * zp[l];  // P has distinct entries: one writer per (b, part)
  }
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) *loss_out = s * inv;
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb_pool_spmm(const void* x, const int32_t* rowptr, const int32_t* colidx, const float* vals, void* y, int B,
                  int rows_in, int rows_out, int C, int dtype, void* stream) {
  if (!x || !rowptr || !colidx || !vals || !y) return SHB_E_ARG;
  if (B <= 0 || rows_in <= 0 || rows_out <= 0 || C <= 0) return SHB_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int V = 1;
  if (dtype == SHB_BF16 && C % 8 == 0) V = 8;
  else if (C % 4 == 0) V = 4;
  const unsigned CV = (unsigned)(C / V);
  const unsigned long long total64 = (unsigned long long)B * rows_out * CV;
  if (total64 >= (1ull << 32)) return SHB_E_SHAPE;
  const unsigned total = (unsigned)total64;
  unsigned blocks = (total / 2 + 255) / 256;  // two items per thread
  if (blocks == 0) blocks = 1;
  if (blocks > 32u * kNumSMs) blocks = 32u * kNumSMs;
#define SHB_POOL_LAUNCH(TT, VV)                                                                                        \
  pool_spmm_kernel<TT, VV><<<blocks, 256, 0, st>>>((const TT*)x, rowptr, colidx, vals, (TT*)y, total, (unsigned)rows_in, \
                                                   (unsigned)rows_out, CV)
  if (dtype == SHB_F32) {
    if (V == 4) SHB_POOL_LAUNCH(float, 4); else SHB_POOL_LAUNCH(float, 1);
  } else if (dtype == SHB_BF16) {
    if (V == 8) SHB_POOL_LAUNCH(__nv_bfloat16, 8); else if (V == 4) SHB_POOL_LAUNCH(__nv_bfloat16, 4); else SHB_POOL_LAUNCH(__nv_bfloat16, 1);
  } else {
    return SHB_E_DTYPE;
  }
#undef SHB_POOL_LAUNCH
  SHB_LAUNCH_CHECK();
  return 0;
}

size_t shb_l1_loss_workspace(int64_t n) { (void)n; return (size_t)L1_BLOCKS * sizeof(float); }

int shb_l1_loss_fwd(const void* a, const void* b, int64_t n, void* partials, size_t partials_bytes, float* loss_out,
                    int dtype, void* stream) {
  if (!a || !b || !partials || !loss_out || n <= 0) return SHB_E_ARG;
  if (partials_bytes < shb_l1_loss_workspace(n)) return SHB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SHB_F32) l1_partial_kernel<float><<<L1_BLOCKS, 256, 0, st>>>((const float*)a, (const float*)b, n, (float*)partials);
  else if (dtype == SHB_BF16) l1_partial_kernel<__nv_bfloat16><<<L1_BLOCKS, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n, (float*)partials);
  else return SHB_E_DTYPE;
  SHB_LAUNCH_CHECK();
  l1_final_kernel<<<1, 256, 0, st>>>((const float*)partials, L1_BLOCKS, 1.0f / (float)n, loss_out);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_l1_loss_bwd(const void* a, const void* b, int64_t n, const float* gscale, void* ga, void* gb, int dtype,
                    void* stream) {
  if (!a || !b || !gscale || n <= 0 || (!ga && !gb)) return SHB_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SHB_F32) l1_bwd_kernel<float><<<L1_BLOCKS, 256, 0, st>>>((const float*)a, (const float*)b, n, gscale, (float*)ga, (float*)gb);
  else if (dtype == SHB_BF16) l1_bwd_kernel<__nv_bfloat16><<<L1_BLOCKS, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n, gscale, (__nv_bfloat16*)ga, (__nv_bfloat16*)gb);
  else return SHB_E_DTYPE;
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_partnorm_loss_fwd_bwd(const float* z, const float* measure, const int32_t* P, const int32_t* Q, float* loss_out,
                              float* gz, int B, int n_parts, int L, int n_measure, int n_sel, int relative,
                              void* stream) {
  if (!z || !measure || !P || !Q || !loss_out || !gz) return SHB_E_ARG;
  if (B <= 0 || n_parts <= 0 || L <= 0 || n_measure <= 0 || n_sel <= 0) return SHB_E_ARG;
  partnorm_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(z, measure, P, Q, loss_out, gz, B, n_parts, L, n_measure, n_sel, relative);
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
