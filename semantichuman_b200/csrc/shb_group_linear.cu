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
// grid (G, ceil(B / GC_BT)).  out[b,k,o] = (bias ? bias_k[o] : 0) + sum_j Wt(o,j) * v[b,j]  where  v[b, p*C+c] = x[b, idx[g0+p], c]
// and Wt(o,j) = w_k[o*K + j]  (w_is_KxL == 0)   or   w_k[j*L + o]  (w_is_KxL != 0: the decode weight read transposed).
// A shared-memory-tiled contraction: a block owns GC_BT samples of one group and walks K in chunks of 128 -- thread t fetches
// column j0 + t of all GC_BT samples (GC_BT independent, coalesced loads in flight per thread) and its share of the chunk's
// weights into shared memory, then thread (sample, output slice) accumulates its LT / 8 outputs over the chunk, j ascending
// (fixed order).  (First version: every block of BT = 2..4 samples streamed the group's whole weight matrix from L2 with one
// dependent load chain per j: 279 us per call for the 16-wide decode heads.)
constexpr int GC_BT = 16;
constexpr int GC_KC = GL_THREADS;   // K chunk: one column per thread
template <int LT>
__global__ void __launch_bounds__(GL_THREADS) gl_contract_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                                 const int32_t* __restrict__ gptr, const float* __restrict__ w,
                                                                 const int64_t* __restrict__ woff, const float* __restrict__ bias,
                                                                 const int64_t* __restrict__ boff, float* __restrict__ out, int B,
                                                                 int rows, int C, int G, int L, int w_is_KxL) {
  constexpr int NO = LT * GC_BT / GL_THREADS;         // outputs per thread
  static_assert(NO >= 1 && NO * GL_THREADS == LT * GC_BT, "tile shape");
  __shared__ float xt[GC_BT][GC_KC + 1];              // odd row stride: a warp's 16 samples x 2 hit distinct banks
  __shared__ __align__(16) float wt[GC_KC][LT];
  const int k = blockIdx.x, b0 = blockIdx.y * GC_BT, t = threadIdx.x;
  const int nb = min(GC_BT, B - b0);
  const int g0 = gptr[k], n = gptr[k + 1] - g0, K = n * C;
  const float* wk = w + woff[k];
  const size_t xs = (size_t)rows * C;
  const float* xb = x + (size_t)b0 * xs;
  const int bs = t & (GC_BT - 1), og = t / GC_BT;     // this thread's sample and output slice [og * NO, og * NO + NO)
  float acc[NO];
#pragma unroll
  for (int o = 0; o < NO; ++o) acc[o] = 0.f;
  // chunk loads into registers: column j0 + t of the GC_BT samples, and LT weights in the order they lie in memory
  float v[GC_BT], wv[LT];
  auto fetch = [&](int j0) {
    const int j = j0 + t;
    if (j < K) {
      const int p = j / C, c = j - p * C;
      const float* xc = xb + (size_t)__ldg(idx + g0 + p) * C + c;
#pragma unroll
      for (int u = 0; u < GC_BT; ++u) v[u] = u < nb ? xc[u * xs] : 0.f;
    } else {
#pragma unroll
      for (int u = 0; u < GC_BT; ++u) v[u] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < LT; ++i) {
      const int q = t + i * GL_THREADS;
      int jj, o;
      if (w_is_KxL) { jj = q / LT; o = q - jj * LT; } else { o = q / GC_KC; jj = q - o * GC_KC; }
      wv[i] = (j0 + jj < K && o < L) ? __ldg(w_is_KxL ? wk + (size_t)(j0 + jj) * L + o : wk + (size_t)o * K + j0 + jj) : 0.f;
    }
  };
  fetch(0);
  for (int j0 = 0; j0 < K; j0 += GC_KC) {
    __syncthreads();   // the previous chunk's readers are done
#pragma unroll
    for (int u = 0; u < GC_BT; ++u) xt[u][t] = v[u];
#pragma unroll
    for (int i = 0; i < LT; ++i) {
      const int q = t + i * GL_THREADS;
      int jj, o;
      if (w_is_KxL) { jj = q / LT; o = q - jj * LT; } else { o = q / GC_KC; jj = q - o * GC_KC; }
      wt[jj][o] = wv[i];
    }
    __syncthreads();
    if (j0 + GC_KC < K) fetch(j0 + GC_KC);   // the next chunk's loads are in flight while this one is contracted
#pragma unroll 8
    for (int jj = 0; jj < GC_KC; ++jj) {
      const float xv = xt[bs][jj];
#pragma unroll
      for (int o = 0; o < NO; ++o) acc[o] = fmaf(wt[jj][og * NO + o], xv, acc[o]);
    }
  }
  if (bs < nb) {
#pragma unroll
    for (int o = 0; o < NO; ++o) {
      const int oo = og * NO + o;
      if (oo < L) out[((size_t)(b0 + bs) * G + k) * L + oo] = acc[o] + (bias ? __ldg(bias + boff[k] + oo) : 0.f);
    }
  }
}

static void gl_contract_launch(const float* x, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                               const float* bias, const int64_t* boff, float* out, int B, int rows, int C, int G, int L,
                               int w_is_KxL, cudaStream_t st) {
  const dim3 grid(G, (B + GC_BT - 1) / GC_BT);
  if (L <= 8) gl_contract_kernel<8><<<grid, GL_THREADS, 0, st>>>(x, idx, gptr, w, woff, bias, boff, out, B, rows, C, G, L, w_is_KxL);
  else if (L <= 16) gl_contract_kernel<16><<<grid, GL_THREADS, 0, st>>>(x, idx, gptr, w, woff, bias, boff, out, B, rows, C, G, L, w_is_KxL);
  else gl_contract_kernel<32><<<grid, GL_THREADS, 0, st>>>(x, idx, gptr, w, woff, bias, boff, out, B, rows, C, G, L, w_is_KxL);
}

// ---- scatter form, forward (also: gather form, gradient w.r.t. x -- same expansion, no bias)
// grid (G, ceil(B / GL_BT), Z).  y[b, idx[g0+p], c] = (bias ? bias_k[j] : 0) + sum_i Wt(j,i) * zz[b,k,i],  j = p*C+c,
// Wt(j,i) = w_k[j*L + i]  (w_is_KxL != 0)   or   w_k[i*K + j]  (the encode weight read transposed).
// A thread keeps its weight row in registers and walks a tile of GL_BT samples with it (the weights are read once per
// sample tile, not once per sample: they are 20x larger than the activations they produce); the tile's latent vectors sit in
// shared memory and are read as broadcast float4s.  LT = L rounded up to 8 / 16 / 32: no predicated-off arithmetic.  The K
// outputs of a group are spread over the Z blocks of the grid's third dimension (first version: one block per (group, sample
// tile) -- 272 blocks of 128 threads for 56 MB of output, 250 us per call).
constexpr int GL_BT = 32;
template <int LT>
__global__ void __launch_bounds__(GL_THREADS) gl_expand_kernel(const float* __restrict__ zz, const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ gptr, const float* __restrict__ w,
                                                               const int64_t* __restrict__ woff, const float* __restrict__ bias,
                                                               const int64_t* __restrict__ boff, float* __restrict__ y, int B,
                                                               int rows, int C, int G, int L, int w_is_KxL) {
  __shared__ __align__(16) float zs[GL_BT][LT];
  const int k = blockIdx.x, b0 = blockIdx.y * GL_BT, t = threadIdx.x;
  const int nb = min(GL_BT, B - b0);
  const int g0 = gptr[k], n = gptr[k + 1] - g0, K = n * C;
  if ((int)blockIdx.z * GL_THREADS >= K) return;  // block-uniform
  const float* wk = w + woff[k];
  for (int q = t; q < GL_BT * LT; q += GL_THREADS) {
    const int bb = q / LT, i = q - bb * LT;
    zs[bb][i] = (bb < nb && i < L) ? zz[((size_t)(b0 + bb) * G + k) * L + i] : 0.f;
  }
  __syncthreads();
  const size_t ys = (size_t)rows * C;
  for (int j = blockIdx.z * GL_THREADS + t; j < K; j += gridDim.z * GL_THREADS) {
    const int p = j / C, c = j - p * C;
    const float bj = bias ? __ldg(bias + boff[k] + j) : 0.f;
    float wr[LT];
#pragma unroll
    for (int i = 0; i < LT; ++i) wr[i] = i < L ? __ldg(w_is_KxL ? wk + (size_t)j * L + i : wk + (size_t)i * K + j) : 0.f;
    float* yp = y + ((size_t)b0 * rows + __ldg(idx + g0 + p)) * C + c;
#pragma unroll 4
    for (int bb = 0; bb < nb; ++bb) {
      float a = bj;
#pragma unroll
      for (int i = 0; i < LT; i += 4) {
        const float4 z4 = *reinterpret_cast<const float4*>(&zs[bb][i]);
        a = fmaf(wr[i], z4.x, a);
        a = fmaf(wr[i + 1], z4.y, a);
        a = fmaf(wr[i + 2], z4.z, a);
        a = fmaf(wr[i + 3], z4.w, a);
      }
      yp[(size_t)bb * ys] = a;
    }
  }
}

static void gl_expand_launch(const float* zz, const int32_t* idx, const int32_t* gptr, const float* w, const int64_t* woff,
                             const float* bias, const int64_t* boff, float* y, int B, int rows, int C, int G, int L, int w_is_KxL,
                             cudaStream_t st) {
  long long kb = ((long long)rows * C + GL_THREADS - 1) / GL_THREADS;   // no group has more outputs than this
  const int Z = (int)(kb < 32 ? kb : 32);
  const dim3 grid(G, (B + GL_BT - 1) / GL_BT, Z);
  if (L <= 8) gl_expand_kernel<8><<<grid, GL_THREADS, 0, st>>>(zz, idx, gptr, w, woff, bias, boff, y, B, rows, C, G, L, w_is_KxL);
  else if (L <= 16) gl_expand_kernel<16><<<grid, GL_THREADS, 0, st>>>(zz, idx, gptr, w, woff, bias, boff, y, B, rows, C, G, L, w_is_KxL);
  else gl_expand_kernel<32><<<grid, GL_THREADS, 0, st>>>(zz, idx, gptr, w, woff, bias, boff, y, B, rows, C, G, L, w_is_KxL);
}

// ---- weight (and scatter-form bias) gradients: one thread per K index j, sequential over the batch (fixed order)
// grid (G, ceil(Kmax / GL_THREADS)).  gw_k(o,j) = sum_b g[b,k,o] * v[b,j]  with v gathered from x as above;
// stored at gw_k[o*K + j] (w_is_KxL == 0) or gw_k[j*L + o].  gbias_j (scatter form only): gbk[j] = sum_b v[b,j].
// The batch is walked in tiles of GL_WB samples: the tile's g vectors go to shared memory in one step, and the tile's loads of
// v are independent of each other (all GL_WB in flight), so a thread is not one global-load latency per sample (first version:
// two barriers and one dependent load per sample, 256 times).
constexpr int GL_WB = 32;
template <int LT>
__global__ void __launch_bounds__(GL_THREADS) gl_wgrad_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                                              const int32_t* __restrict__ gptr, const float* __restrict__ g,
                                                              const int64_t* __restrict__ woff, const int64_t* __restrict__ boff,
                                                              float* __restrict__ gw, float* __restrict__ gbias_j, int B,
                                                              int rows, int C, int G, int L, int w_is_KxL) {
  __shared__ __align__(16) float gs[GL_WB][LT];
  const int k = blockIdx.x, t = threadIdx.x;
  const int g0 = gptr[k], n = gptr[k + 1] - g0, K = n * C;
  if ((int)blockIdx.y * GL_THREADS >= K) return;  // whole block beyond this group's K (block-uniform)
  const int j = blockIdx.y * GL_THREADS + t;
  const bool on = j < K;  // inactive threads of the last block still take part in the barriers
  const int p = on ? j / C : 0, c = on ? j - p * C : 0;
  const size_t col = on ? (size_t)__ldg(idx + g0 + p) * C + c : 0;
  const size_t xs = (size_t)rows * C;
  float acc[LT];
#pragma unroll
  for (int o = 0; o < LT; ++o) acc[o] = 0.f;
  float vs = 0.f;
  for (int b0 = 0; b0 < B; b0 += GL_WB) {
    const int nb = min(GL_WB, B - b0);
    __syncthreads();
    for (int q = t; q < GL_WB * LT; q += GL_THREADS) {
      const int bb = q / LT, i = q - bb * LT;
      gs[bb][i] = (bb < nb && i < L) ? g[((size_t)(b0 + bb) * G + k) * L + i] : 0.f;
    }
    __syncthreads();
    if (on) {
      const float* xp = x + (size_t)b0 * xs + col;
      float v[GL_WB];   // the whole tile's loads in flight together (54 K threads in all: occupancy is low, so each thread
                        // has to keep many bytes in flight)
      if (nb == GL_WB) {   // block-uniform: the common case without per-load predicates
#pragma unroll
        for (int u = 0; u < GL_WB; ++u) v[u] = __ldg(xp + (size_t)u * xs);
      } else {
#pragma unroll
        for (int u = 0; u < GL_WB; ++u) v[u] = u < nb ? __ldg(xp + (size_t)u * xs) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < GL_WB; ++u) {
        vs += v[u];
#pragma unroll
        for (int o = 0; o < LT; o += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(&gs[u][o]);
          acc[o] = fmaf(g4.x, v[u], acc[o]);
          acc[o + 1] = fmaf(g4.y, v[u], acc[o + 1]);
          acc[o + 2] = fmaf(g4.z, v[u], acc[o + 2]);
          acc[o + 3] = fmaf(g4.w, v[u], acc[o + 3]);
        }
      }
    }
  }
  if (!on) return;
  float* gwk = gw + woff[k];
#pragma unroll
  for (int o = 0; o < LT; ++o)
    if (o < L) gwk[w_is_KxL ? (size_t)j * L + o : (size_t)o * K + j] = acc[o];
  if (gbias_j) gbias_j[boff[k] + j] = vs;
}

static void gl_wgrad_launch(const float* x, const int32_t* idx, const int32_t* gptr, const float* g, const int64_t* woff,
                            const int64_t* boff, float* gw, float* gbias_j, int B, int rows, int C, int G, int L, int w_is_KxL,
                            int max_group_rows, cudaStream_t st) {
  const int kb = (max_group_rows * C + GL_THREADS - 1) / GL_THREADS;
  const dim3 grid(G, kb);
  if (L <= 8) gl_wgrad_kernel<8><<<grid, GL_THREADS, 0, st>>>(x, idx, gptr, g, woff, boff, gw, gbias_j, B, rows, C, G, L, w_is_KxL);
  else if (L <= 16) gl_wgrad_kernel<16><<<grid, GL_THREADS, 0, st>>>(x, idx, gptr, g, woff, boff, gw, gbias_j, B, rows, C, G, L, w_is_KxL);
  else gl_wgrad_kernel<32><<<grid, GL_THREADS, 0, st>>>(x, idx, gptr, g, woff, boff, gw, gbias_j, B, rows, C, G, L, w_is_KxL);
}

// gather-form bias gradient: gb_k[o] = sum_b g[b,k,o]; one block per group.  Thread (o, s) adds the samples s, s + S, ... (S =
// threads / LT slices), thread (o, 0) then adds the S slice sums in slice order: fixed order, and 256 / S dependent loads
// instead of 256.
__global__ void __launch_bounds__(GL_THREADS) gl_bias_kernel(const float* __restrict__ g, const int64_t* __restrict__ boff,
                                                             float* __restrict__ gb, int B, int G, int L) {
  __shared__ float part[GL_THREADS];
  const int k = blockIdx.x, t = threadIdx.x;
  const int S = GL_THREADS / GL_MAX;            // 4 slices of up to 32 outputs
  const int o = t % GL_MAX, sl = t / GL_MAX;
  float s = 0.f;
  if (o < L)
    for (int b = sl; b < B; b += S) s += g[((size_t)b * G + k) * L + o];
  part[t] = s;
  __syncthreads();
  if (sl == 0 && o < L) {
    float r = 0.f;
    for (int q = 0; q < S; ++q) r += part[q * GL_MAX + o];
    gb[boff[k] + o] = r;
  }
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
    gl_expand_launch(gz, idx, gptr, w, woff, nullptr, nullptr, gx, B, rows, C, G, L, 0, st);
    SHB_LAUNCH_CHECK();
  }
  if (gw) {
    gl_wgrad_launch(x, idx, gptr, gz, woff, nullptr, gw, nullptr, B, rows, C, G, L, 0, max_group_rows, st);
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
  gl_expand_launch(zz, idx, gptr, w, woff, bias, boff, y, B, rows, C, G, Lin, 1, (cudaStream_t)stream);
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
    gl_wgrad_launch(gy, idx, gptr, zz, woff, boff, gw, gb, B, rows, C, G, Lin, 1, max_group_rows, st);
    SHB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
