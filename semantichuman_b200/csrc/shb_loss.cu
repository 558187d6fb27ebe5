// Loss reductions of libshb200 (L1 reconstruction loss, part-measure latent loss).  HBM-bound: coalesced 16-byte accesses,
// grids sized in multiples of the SM count, fixed-order two-stage reductions (no float atomics).
#include "shb_common.cuh"

namespace shb {

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

// ------------------------------------------------------------------------------------------------ column sums
// out[c] = sum_b x[b][c] (the bias gradient of an nn.Linear: models.py:129,142), x row-major (B, N), fp32 result.
// A CTA owns VPC 8-column vectors: thread (v, rg) adds rows rg, rg + RG, ... of vector v (16-byte loads for bf16, coalesced
// across the vectors, eight in flight), the RG row groups are added in index order through shared memory.  Fixed order, no
// atomics.  (ATen's reduce kernel ran this with 2 CTAs for N = 256: 15 us of latency.)
constexpr int CS_ROWGROUPS = 8;
// VPC = 8-column vectors per CTA, RG = row groups: (16, 16) for long rows (N / 128 CTAs, all resident at once), (32, 32) for
// short ones (a single CTA for N = 256 -- its 1024 threads then need one round of eight loads each instead of four)
template <typename T, int VPC, int RG>
__global__ void __launch_bounds__(VPC * RG, 1024 / (VPC * RG)) colsum_kernel(const T* __restrict__ x, int B, int N, float* __restrict__ out) {
  __shared__ float part[RG][VPC][9];
  const int v = threadIdx.x % VPC, rg = threadIdx.x / VPC;
  const int c0 = (blockIdx.x * VPC + v) * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < N) {
    int b = rg;
    for (; b + 7 * RG < B; b += 8 * RG) {
      float r[8][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) Io<T>::ld8(x + (size_t)(b + k * RG) * N + c0, r[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += r[k][e];
    }
    for (; b < B; b += RG) {
      float r[8];
      Io<T>::ld8(x + (size_t)b * N + c0, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += r[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[rg][v][e] = acc[e];
  __syncthreads();
  const int col = threadIdx.x;
  if (col < VPC * 8) {
    const int c = blockIdx.x * VPC * 8 + col;
    if (c < N) {
      float s = 0.f;
#pragma unroll 8
      for (int k = 0; k < RG; ++k) s += part[k][col >> 3][col & 7];
      out[c] = s;
    }
  }
}

// any N: a CTA owns 32 columns (lane = column: coalesced), the same eight row groups
template <typename T>
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const T* __restrict__ x, int B, int N, float* __restrict__ out) {
  __shared__ float part[CS_ROWGROUPS][32];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < N)
    for (int b = rg; b < B; b += CS_ROWGROUPS) acc += Io<T>::ld(x + (size_t)b * N + c);
  part[rg][lane] = acc;
  __syncthreads();
  if (rg == 0 && c < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CS_ROWGROUPS; ++k) s += part[k][lane];
    out[c] = s;
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

int shb_colsum(const void* x, int dtype, int B, int N, float* out, void* stream) {
  if (!x || !out || B <= 0 || N <= 0) return SHB_E_ARG;
  if (dtype != SHB_F32 && dtype != SHB_BF16) return SHB_E_DTYPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (N % 8 != 0) {   // rows not 16-byte aligned: one column per thread
    const int grid = (N + 31) / 32;
    if (dtype == SHB_F32) colsum_scalar_kernel<float><<<grid, 256, 0, st>>>((const float*)x, B, N, out);
    else colsum_scalar_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, N, out);
    SHB_LAUNCH_CHECK();
    return 0;
  }
  if ((long long)N >= 128LL * kNumSMs) {
    const int grid = (N + 127) / 128;
    if (dtype == SHB_F32) colsum_kernel<float, 16, 16><<<grid, 256, 0, st>>>((const float*)x, B, N, out);
    else colsum_kernel<__nv_bfloat16, 16, 16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, N, out);
  } else {
    const int grid = (N + 255) / 256;
    if (dtype == SHB_F32) colsum_kernel<float, 32, 32><<<grid, 1024, 0, st>>>((const float*)x, B, N, out);
    else colsum_kernel<__nv_bfloat16, 32, 32><<<grid, 1024, 0, st>>>((const __nv_bfloat16*)x, B, N, out);
  }
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
