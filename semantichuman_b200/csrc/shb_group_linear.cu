// Bone-guided heads as grouped kernels (models.py:200-204, 233-236, 252-253, 269-273).
//
// The reference runs 17 + 17 + 17 tiny nn.Linear layers from Python list comprehensions, each on a fancy-indexed copy of
// the coarsest-level features, then scatters the decoder pieces back with a permutation and concatenates the dummy row.
// Here a "group" k owns the rows idx[gptr[k] .. gptr[k+1]) of a (B, rows, C) tensor; its weight and bias live at
// w + woff[k] and bias + boff[k] of packed fp32 buffers (nn.Linear layout, row-major):
//
//   gather ("encode") form :  z[b,k,o]         = bias_k[o] + sum_{p,c} W_k[o, p*C+c] * x[b, idx[gptr[k]+p], c]     W_k (L, n_k*C)
//   scatter ("decode") form:  y[b,idx[..+p],c] = bias_k[p*C+c] + sum_i W_k[p*C+c, i] * zz[b,k,i]                   W_k (n_k*C, Lin)
//
// One launch per direction instead of 17; the gathers, the permutation scatter and the bias are folded into the kernels.
// All reductions run in a fixed order (no float atomics).  fp32 only: the heads stay fp32 in both compute modes.
#include "shb_common.cuh"

namespace shb {

constexpr int GL_MAX = 32;      // widest latent side (L, Lin) the register accumulators hold
constexpr int GL_THREADS = 128;

// ---- gather form, forward (also: scatter form, gradient w.r.t. zz -- same contraction, no bias)
// grid (G, ceil(B / BT)).  out[b,k,o] = (bias ? bias_k[o] : 0) + sum_j Wt(o,j) * v[b,j]  where  v[b, p*C+c] = x[b, idx[g0+p], c]
// and Wt(o,j) = w_k[o*K + j]  (w_is_KxL == 0)   or   w_k[j*L + o]  (w_is_KxL != 0: the decode weight read transposed).
// A block handles BT samples at once (BT * LM = GL_MAX accumulators per thread) so that a weight element is loaded once per
// sample tile; L <= LM.
template <int BT, int LM>
__global__ void __launch_bounds__(GL_THREADS) gl_contract_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                                 const int32_t* __restrict__ gptr, const float* __restrict__ w,
                                                                 const int64_t* __restrict__ woff, const float* __restrict__ bias,
                                                                 const int64_t* __restrict__ boff, float* __restrict__ out, int B,
                                                                 int rows, int C, int G, int L, int w_is_KxL) {
  __shared__ float red[GL_THREADS / 32][BT][LM];
  const int k = blockIdx.x, b0 = blockIdx.y * BT, t = threadIdx.x;
  const int nb = min(BT, B - b0);
  const int g0 = gptr[k], n = gptr[k + 1] - g0, K = n * C;
  const float* wk = w + woff[k];
  const size_t xs = (size_t)rows * C;
  const float* xb = x + (size_t)b0 * xs;
  float acc[BT][LM];
#pragma unroll
  for (int u = 0; u < BT; ++u)
#pragma unroll
    for (int o = 0; o < LM; ++o) acc[u][o] = 0.f;
  for (int j = t; j < K; j += GL_THREADS) {
    const int p = j / C, c = j - p * C;
    const size_t col = (size_t)__ldg(idx + g0 + p) * C + c;
    float v[BT];
#pragma unroll
    for (int u = 0; u < BT; ++u) v[u] = u < nb ? xb[u * xs + col] : 0.f;
#pragma unroll
    for (int o = 0; o < LM; ++o) {
      if (o < L) {
        const float wv = __ldg(w_is_KxL ? wk + (size_t)j * L + o : wk + (size_t)o * K + j);
#pragma unroll
        for (int u = 0; u < BT; ++u) acc[u][o] = fmaf(wv, v[u], acc[u][o]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < BT; ++u)
#pragma unroll
    for (int o = 0; o < LM; ++o) {
      if (o < L) {  // L is uniform across the block
        const float s = warp_sum(acc[u][o]);
        if ((t & 31) == 0) red[t >> 5][u][o] = s;
      }
    }
  __syncthreads();
  if (t < nb * L) {
    const int u = t / L, o = t - u * L;
    float s = bias ? __ldg(bias + boff[k] + o) : 0.f;
#pragma unroll
    for (int q = 0; q < GL_THREADS / 32; ++q) s += red[q][u][o];
    out[((size_t)(b0 + u) * G + k) * L + o] = s;
  }
}

static void gl_contract_launch(const float* x, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                               const float* bias, const int64_t* boff, float* out, int B, int rows, int C, int G, int L,
                               int w_is_KxL, cudaStream_t st) {
  if (L <= 8)
    gl_contract_kernel<4, 8><<<dim3(G, (B + 3) / 4), GL_THREADS, 0, st>>>(x, idx, gptr, w, woff, bias, boff, out, B, rows, C, G, L, w_is_KxL);
  else if (L <= 16)
    gl_contract_kernel<2, 16><<<dim3(G, (B + 1) / 2), GL_THREADS, 0, st>>>(x, idx, gptr, w, woff, bias, boff, out, B, rows, C, G, L, w_is_KxL);
  else
    gl_contract_kernel<1, 32><<<dim3(G, B), GL_THREADS, 0, st>>>(x, idx, gptr, w, woff, bias, boff, out, B, rows, C, G, L, w_is_KxL);
}

// ---- scatter form, forward (also: gather form, gradient w.r.t. x -- same expansion, no bias)
// grid (G, ceil(B / GL_BT)).  y[b, idx[g0+p], c] = (bias ? bias_k[j] : 0) + sum_i Wt(j,i) * zz[b,k,i],  j = p*C+c,
// Wt(j,i) = w_k[j*L + i]  (w_is_KxL != 0)   or   w_k[i*K + j]  (the encode weight read transposed).
// A thread keeps its weight row in registers and walks a tile of GL_BT samples with it (the weights are read once per
// sample tile, not once per sample: they are 20x larger than the activations they produce).
constexpr int GL_BT = 16;
__global__ void __launch_bounds__(GL_THREADS) gl_expand_kernel(const float* __restrict__ zz, const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ gptr, const float* __restrict__ w,
                                                               const int64_t* __restrict__ woff, const float* __restrict__ bias,
                                                               const int64_t* __restrict__ boff, float* __restrict__ y, int B,
                                                               int rows, int C, int G, int L, int w_is_KxL) {
  __shared__ float zs[GL_BT][GL_MAX];
  const int k = blockIdx.x, b0 = blockIdx.y * GL_BT, t = threadIdx.x;
  const int nb = min(GL_BT, B - b0);
  const int g0 = gptr[k], n = gptr[k + 1] - g0, K = n * C;
  const float* wk = w + woff[k];
  for (int q = t; q < nb * L; q += GL_THREADS) {
    const int bb = q / L, i = q - bb * L;
    zs[bb][i] = zz[((size_t)(b0 + bb) * G + k) * L + i];
  }
  __syncthreads();
  for (int j = t; j < K; j += GL_THREADS) {
    const int p = j / C, c = j - p * C;
    const float bj = bias ? __ldg(bias + boff[k] + j) : 0.f;
    float wr[GL_MAX];
#pragma unroll
    for (int i = 0; i < GL_MAX; ++i)
      wr[i] = i < L ? __ldg(w_is_KxL ? wk + (size_t)j * L + i : wk + (size_t)i * K + j) : 0.f;
    float* yp = y + ((size_t)b0 * rows + __ldg(idx + g0 + p)) * C + c;
    for (int bb = 0; bb < nb; ++bb) {
      float a = bj;
#pragma unroll
      for (int i = 0; i < GL_MAX; ++i)
        if (i < L) a = fmaf(wr[i], zs[bb][i], a);
      yp[(size_t)bb * rows * C] = a;
    }
  }
}

// ---- weight (and scatter-form bias) gradients: one thread per K index j, sequential over the batch (fixed order)
// grid (G, ceil(Kmax / GL_THREADS)).  gw_k(o,j) = sum_b g[b,k,o] * v[b,j]  with v gathered from x as above;
// stored at gw_k[o*K + j] (w_is_KxL == 0) or gw_k[j*L + o].  gbias_j (scatter form only): gbk[j] = sum_b v[b,j].
__global__ void __launch_bounds__(GL_THREADS) gl_wgrad_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                              const int32_t* __restrict__ gptr, const float* __restrict__ g,
                                                              const int64_t* __restrict__ woff, const int64_t* __restrict__ boff,
                                                              float* __restrict__ gw, float* __restrict__ gbias_j, int B,
                                                              int rows, int C, int G, int L, int w_is_KxL) {
  __shared__ float gs[GL_MAX];
  const int k = blockIdx.x, t = threadIdx.x;
  const int g0 = gptr[k], n = gptr[k + 1] - g0, K = n * C;
  const int j = blockIdx.y * GL_THREADS + t;
  const bool on = j < K;  // block-uniform exit is not possible (the barrier below): inactive threads just idle
  const int p = on ? j / C : 0, c = on ? j - p * C : 0;
  const size_t col = on ? (size_t)__ldg(idx + g0 + p) * C + c : 0;
  if ((int)blockIdx.y * GL_THREADS >= K) return;  // whole block beyond this group's K
  float acc[GL_MAX];
#pragma unroll
  for (int o = 0; o < GL_MAX; ++o) acc[o] = 0.f;
  float vs = 0.f;
  for (int b = 0; b < B; ++b) {
    __syncthreads();
    if (t < L) gs[t] = g[((size_t)b * G + k) * L + t];
    __syncthreads();
    if (on) {
      const float v = x[(size_t)b * rows * C + col];
      vs += v;
#pragma unroll
      for (int o = 0; o < GL_MAX; ++o)
        if (o < L) acc[o] = fmaf(gs[o], v, acc[o]);
    }
  }
  if (!on) return;
  float* gwk = gw + woff[k];
#pragma unroll
  for (int o = 0; o < GL_MAX; ++o)
    if (o < L) gwk[w_is_KxL ? (size_t)j * L + o : (size_t)o * K + j] = acc[o];
  if (gbias_j) gbias_j[boff[k] + j] = vs;
}

// gather-form bias gradient: gb_k[o] = sum_b g[b,k,o]; one block per group
__global__ void __launch_bounds__(GL_THREADS) gl_bias_kernel(const float* __restrict__ g, const int64_t* __restrict__ boff,
                                                             float* __restrict__ gb, int B, int G, int L) {
  const int k = blockIdx.x, t = threadIdx.x;
  if (t >= L) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += g[((size_t)b * G + k) * L + t];
  gb[boff[k] + t] = s;
}

}  // namespace shb

using namespace shb;

extern "C" {

static int gl_args_ok(const void* a, const void* b, const void* c, const void* d, const void* e, int B, int rows, int C, int G,
                      int L) {
  if (!a || !b || !c || !d || !e) return SHB_E_ARG;
  if (B <= 0 || rows <= 0 || C <= 0 || G <= 0 || L <= 0) return SHB_E_ARG;
  if (L > GL_MAX || B > 65535) return SHB_E_SHAPE;
  return 0;
}

int shb_group_linear_gather_fwd(const float* x, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                const float* bias, const int64_t* boff, float* z, int B, int rows, int C, int G, int L,
                                void* stream) {
  int rc = gl_args_ok(x, idx, gptr, w, woff, B, rows, C, G, L);
  if (rc) return rc;
  if (!z || (bias && !boff)) return SHB_E_ARG;
  gl_contract_launch(x, idx, gptr, w, woff, bias, boff, z, B, rows, C, G, L, 0, (cudaStream_t)stream);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_group_linear_gather_bwd(const float* x, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                const int64_t* boff, const float* gz, float* gx, float* gw, float* gb, int B, int rows, int C,
                                int G, int L, int max_group_rows, void* stream) {
  int rc = gl_args_ok(x, idx, gptr, w, woff, B, rows, C, G, L);
  if (rc) return rc;
  if (!gz || max_group_rows <= 0) return SHB_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (gx) {  // rows outside every group (the dummy row) get zero
    cudaError_t e = cudaMemsetAsync(gx, 0, (size_t)B * rows * C * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    gl_expand_kernel<<<dim3(G, (B + GL_BT - 1) / GL_BT), GL_THREADS, 0, st>>>(gz, idx, gptr, w, woff, nullptr, nullptr, gx, B, rows, C, G, L, 0);
    SHB_LAUNCH_CHECK();
  }
  if (gw) {
    const int kb = (max_group_rows * C + GL_THREADS - 1) / GL_THREADS;
    gl_wgrad_kernel<<<dim3(G, kb), GL_THREADS, 0, st>>>(x, idx, gptr, gz, woff, nullptr, gw, nullptr, B, rows, C, G, L, 0);
    SHB_LAUNCH_CHECK();
  }
  if (gb) {
    if (!boff) return SHB_E_ARG;
    gl_bias_kernel<<<G, GL_THREADS, 0, st>>>(gz, boff, gb, B, G, L);
    SHB_LAUNCH_CHECK();
  }
  return 0;
}

int shb_group_linear_scatter_fwd(const float* zz, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                 const float* bias, const int64_t* boff, float* y, int B, int rows, int C, int G, int Lin,
                                 void* stream) {
  int rc = gl_args_ok(zz, idx, gptr, w, woff, B, rows, C, G, Lin);
  if (rc) return rc;
  if (!y || (bias && !boff)) return SHB_E_ARG;
  gl_expand_kernel<<<dim3(G, (B + GL_BT - 1) / GL_BT), GL_THREADS, 0, (cudaStream_t)stream>>>(zz, idx, gptr, w, woff, bias, boff, y, B, rows, C, G, Lin, 1);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_group_linear_scatter_bwd(const float* zz, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                                 const int64_t* boff, const float* gy, float* gzz, float* gw, float* gb, int B, int rows, int C,
                                 int G, int Lin, int max_group_rows, void* stream) {
  int rc = gl_args_ok(zz, idx, gptr, w, woff, B, rows, C, G, Lin);
  if (rc) return rc;
  if (!gy || max_group_rows <= 0) return SHB_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (gzz) {
    gl_contract_launch(gy, idx, gptr, w, woff, nullptr, nullptr, gzz, B, rows, C, G, Lin, 1, st);
    SHB_LAUNCH_CHECK();
  }
  if (gw || gb) {
    if (!gw || (gb && !boff)) return SHB_E_ARG;
    const int kb = (max_group_rows * C + GL_THREADS - 1) / GL_THREADS;
    gl_wgrad_kernel<<<dim3(G, kb), GL_THREADS, 0, st>>>(gy, idx, gptr, zz, woff, boff, gw, gb, B, rows, C, G, Lin, 1);
    SHB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
