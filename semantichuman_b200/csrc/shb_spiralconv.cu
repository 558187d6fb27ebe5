// SpiralConv forward / backward for libshb200 -- exact-fp32 (CUDA-core FFMA) path.
//
// One tiled gather-GEMM kernel serves both directions:
//   forward  (models.py:42-51):  y[(b,j), n]  = act( sum_{s,c} x[b, table[j,s], c] * W[n, s*Cin+c] + bias[n] )
//   dgrad    (autograd of :42,45): gx[(b,u), c] = sum_{s,n} ( sum_{j in inv(u,s)} gz[b,j,n] ) * W[n, s*Cin+c]
// i.e. C[M x N] = A[M x K] . Bop[K x N] where a K-slice of A for slot s is a gathered row (forward) or the
// fixed-order sum of the gathered rows of the (u,s) inverse list (dgrad).  A is never materialised in HBM.
// The weight gradient streams gz / gathered x once per slot s and reduces over rows in two fixed-order stages.
#include <stdlib.h>

#include "shb_common.cuh"
#include "shb_internal.h"

namespace shb {

constexpr int GG_BM = 128;  // rows of (b, vertex) per CTA
constexpr int GG_BK = 16;   // K elements per stage
constexpr int GG_NT = 256;
constexpr int GG_LDA = GG_BM + 4;

struct GGParams {
  const void* src;         // (B, rows_src, Cs)
  const int32_t* table;    // SINGLE: (rows_dst, S) rows of src.  MULTI: keyptr (rows_dst*S + 1)
  const int32_t* list;     // MULTI: concatenated src rows per key
  const void* w;           // nn.Linear weight (Cout, S*Cin)
  const void* bias;        // (Cd) or null
  void* dst;               // (B, rows_dst, Cd)
  long long M;             // B * rows_dst
  int rows_src, rows_dst, S, Cs, Cd, K;  // K = S*Cs
  int act, zero_last, skip_last;
};

// T: storage type. BN x (TM x TN): CTA / thread tile along N.  MULTI: gather-sum over inverse lists + dgrad weight view.
template <typename T, int BN, int TM, int TN, bool MULTI>
__global__ void __launch_bounds__(GG_NT) gather_gemm_kernel(const GGParams p) {
  constexpr int TX = BN / TN;
  static_assert((GG_BM / TM) * TX == GG_NT, "thread tiling must cover the CTA tile");
  constexpr int LDB = BN + 4;
  constexpr int B_PER_T = (GG_BK * BN + GG_NT - 1) / GG_NT;

  __shared__ __align__(16) float As[GG_BK][GG_LDA];
  __shared__ __align__(16) float Bs[GG_BK][LDB];

  const T* __restrict__ src = static_cast<const T*>(p.src);
  const T* __restrict__ w = static_cast<const T*>(p.w);
  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * GG_BM;
  const int n0 = blockIdx.y * BN;

  // ---- A loader role: row r = t/2, 8 consecutive k per thread
  const int lr = t >> 1, lh = (t & 1) * 8;
  const long long lm = m0 + lr;
  const bool lvalid = lm < p.M;
  int lj = 0;
  const T* srcb = src;
  if (lvalid) {
    const long long b = lm / p.rows_dst;
    lj = (int)(lm - b * p.rows_dst);
    srcb = src + b * (long long)p.rows_src * p.Cs;
  }
  const bool lskip = !lvalid || (MULTI && p.skip_last && lj == p.rows_dst - 1);
  const bool vec = (p.Cs % GG_BK) == 0;

  float areg[8];
  float breg[B_PER_T];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) areg[i] = 0.f;
    if (lskip) return;
    if (vec) {
      const int s = k0 / p.Cs;
      const int c0 = k0 - s * p.Cs + lh;
      if (!MULTI) {
        const int row = __ldg(p.table + (long long)lj * p.S + s);
        Io<T>::ld8(srcb + (long long)row * p.Cs + c0, areg);
      } else {
        const int e0 = __ldg(p.table + (long long)lj * p.S + s), e1 = __ldg(p.table + (long long)lj * p.S + s + 1);
        for (int e = e0; e < e1; ++e) {
          float v[8];
          Io<T>::ld8(srcb + (long long)__ldg(p.list + e) * p.Cs + c0, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) areg[i] += v[i];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + lh + i;
        if (k < p.K) {
          const int s = k / p.Cs, c = k - s * p.Cs;
          if (!MULTI) {
            const int row = __ldg(p.table + (long long)lj * p.S + s);
            areg[i] = Io<T>::ld(srcb + (long long)row * p.Cs + c);
          } else {
            const int e0 = __ldg(p.table + (long long)lj * p.S + s), e1 = __ldg(p.table + (long long)lj * p.S + s + 1);
            float a = 0.f;
            for (int e = e0; e < e1; ++e) a += Io<T>::ld(srcb + (long long)__ldg(p.list + e) * p.Cs + c);
            areg[i] = a;
          }
        }
      }
    }
  };

  // Bop[k][n]: forward W[n*K + k]; dgrad (k = s*Cs + c', Cs = Cout, n < Cd = Cin): W[c'*(S*Cd) + s*Cd + n]
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < B_PER_T; ++i) {
      const int e = t + i * GG_NT;
      float v = 0.f;
      if (e < GG_BK * BN) {
        int kk, n;
        if (!MULTI) { n = e / GG_BK; kk = e - n * GG_BK; } else { kk = e / BN; n = e - kk * BN; }
        const int k = k0 + kk, ng = n0 + n;
        if (k < p.K && ng < p.Cd) {
          if (!MULTI) {
            v = Io<T>::ld(w + (long long)ng * p.K + k);
          } else {
            const int s = k / p.Cs, c = k - s * p.Cs;
            v = Io<T>::ld(w + (long long)c * ((long long)p.S * p.Cd) + (long long)s * p.Cd + ng);
          }
        }
      }
      breg[i] = v;
    }
  };

  auto store_smem = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[lh + i][lr] = areg[i];
#pragma unroll
    for (int i = 0; i < B_PER_T; ++i) {
      const int e = t + i * GG_NT;
      if (e < GG_BK * BN) {
        int kk, n;
        if (!MULTI) { n = e / GG_BK; kk = e - n * GG_BK; } else { kk = e / BN; n = e - kk * BN; }
        Bs[kk][n] = breg[i];
      }
    }
  };

  const int tx = t % TX, ty = t / TX;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_a(0);
  load_b(0);
  store_smem();
  __syncthreads();
  for (int k0 = 0; k0 < p.K; k0 += GG_BK) {
    const bool more = k0 + GG_BK < p.K;
    if (more) { load_a(k0 + GG_BK); load_b(k0 + GG_BK); }
#pragma unroll
    for (int kk = 0; kk < GG_BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) { store_smem(); __syncthreads(); }
  }

  // ---- epilogue: bias + activation + dummy-row mask
  T* __restrict__ dst = static_cast<T*>(p.dst);
  const float* __restrict__ bias = static_cast<const float*>(p.bias);  // fp32 in every mode
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + tx * TN + j;
    bv[j] = (bias != nullptr && n < p.Cd) ? __ldg(bias + n) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    const int j_row = (int)(m % p.rows_dst);
    const bool zero = p.zero_last && (j_row == p.rows_dst - 1);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < p.Cd) {
        const float v = zero ? 0.f : act_fwd(acc[i][j] + bv[j], p.act);
        Io<T>::st(dst + m * p.Cd + n, v);
      }
    }
  }
}

template <typename T, bool MULTI> static int launch_gather_gemm(const GGParams& p, cudaStream_t st) {
  const unsigned gx = (unsigned)((p.M + GG_BM - 1) / GG_BM);
  auto grid = [&](int bn) { return dim3(gx, (unsigned)((p.Cd + bn - 1) / bn), 1); };
  if (p.Cd <= 4) gather_gemm_kernel<T, 4, 2, 1, MULTI><<<grid(4), GG_NT, 0, st>>>(p);
  else if (p.Cd <= 16) gather_gemm_kernel<T, 16, 4, 2, MULTI><<<grid(16), GG_NT, 0, st>>>(p);
  else if (p.Cd <= 32) gather_gemm_kernel<T, 32, 4, 4, MULTI><<<grid(32), GG_NT, 0, st>>>(p);
  else if (p.Cd <= 64) gather_gemm_kernel<T, 64, 8, 4, MULTI><<<grid(64), GG_NT, 0, st>>>(p);
  else gather_gemm_kernel<T, 128, 8, 8, MULTI><<<grid(128), GG_NT, 0, st>>>(p);
  SHB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ act backward
// gz[row, c] = c < C ? gy[row, c] * act'(y[row, c]) : 0  (gz has CP >= C channels; dummy rows zeroed)
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gz,
                                                      unsigned rows, unsigned rows_out, int C, int CP, int act, int zero_last) {
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned cp = (unsigned)CP;
  const unsigned long long total = (unsigned long long)rows * cp;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned row = (unsigned)(i / cp), c = (unsigned)(i - (unsigned long long)row * cp);
    float v = 0.f;
    if ((int)c < C && !(zero_last && (row % rows_out) == rows_out - 1)) {
      const size_t k = (size_t)row * C + c;
      v = Io<T>::ld(gy + k) * act_bwd_from_out(Io<T>::ld(y + k), act);
    }
    Io<T>::st(gz + i, v);
  }
}

// 16-byte vectorised variant for the un-padded case (C % VEC == 0): one thread per VEC-wide channel chunk, 32-bit math
// When `colpart` is given (chunks_per_row divides 256, so a thread always sees the same channel chunk) the kernel also
// accumulates the per-channel sums of gz -- the bias gradient -- and writes one fixed-order partial per block.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) act_bwd_vec_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gz,
                                                          unsigned chunks, unsigned chunks_per_row, unsigned rows_out, int act,
                                                          int zero_last, float* __restrict__ colpart) {
  __shared__ float red[256][VEC + 1];
  const unsigned stride = gridDim.x * blockDim.x;
  float csum[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) csum[k] = 0.f;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += stride) {
    const unsigned row = i / chunks_per_row;
    const bool zero = zero_last && (row % rows_out) == rows_out - 1;
    float g[VEC], o[VEC], r[VEC];
    if (VEC == 8) { Io<T>::ld8(gy + (size_t)i * VEC, g); Io<T>::ld8(y + (size_t)i * VEC, o); }
    else { Io<T>::ld4(gy + (size_t)i * VEC, g); Io<T>::ld4(y + (size_t)i * VEC, o); }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      r[k] = zero ? 0.f : g[k] * act_bwd_from_out(o[k], act);
      csum[k] += r[k];
    }
    Io<T>::st4(gz + (size_t)i * VEC, r);
    if (VEC == 8) Io<T>::st4(gz + (size_t)i * VEC + 4, r + 4);
  }
  if (colpart != nullptr) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) red[threadIdx.x][k] = csum[k];
    __syncthreads();
    const unsigned C = chunks_per_row * VEC;
    if (threadIdx.x < C) {  // column c: threads j*cpr + c/VEC (ascending j) hold its partial sums
      const unsigned cc = threadIdx.x / VEC, k = threadIdx.x % VEC;
      float s = 0.f;
      for (unsigned j = cc; j < 256; j += chunks_per_row) s += red[j][k];
      colpart[(size_t)blockIdx.x * C + threadIdx.x] = s;
    }
  }
}

// generic column sums of a (rows, C) tensor in fp32: per-block partials (fixed order), used when the fused path is off
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_generic(const T* __restrict__ g, unsigned long long rows, int C, int CP,
                                                              unsigned long long rows_per_block, float* __restrict__ part) {
  __shared__ float red[256];
  const int tx = threadIdx.x % CP, ty = threadIdx.x / CP, ny = 256 / CP;
  const unsigned long long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float a = 0.f;
  if (tx < C)
    for (unsigned long long r = r0 + ty; r < r1; r += ny) a += Io<T>::ld(g + r * C + tx);
  red[threadIdx.x] = a;
  __syncthreads();
  if (ty == 0 && tx < C) {
    float s = 0.f;
    for (int q = 0; q < ny; ++q) s += red[q * CP + tx];
    part[(size_t)blockIdx.x * C + tx] = s;
  }
}

// out[i] = sum_c part[c*n + i]: one CTA per output column, 256 part-lanes (each ascending), then the fixed-order
// block reduction -- deterministic, and ~10 dependent loads per thread instead of ~300 (parts is a few thousand)
__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ ws, int parts, int n, float* __restrict__ out) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  float a = 0.f;
  for (int c = threadIdx.x; c < parts; c += 256) a += ws[(size_t)c * n + i];
  a = block_sum<256>(a, red);
  if (threadIdx.x == 0) out[i] = a;
}

// dst[row, c] = c < C ? (Tdst)src[row, c] : 0   -- channel zero-padding (+ fp32 -> bf16 cast) so that 3-channel
// mesh coordinates can use the 16-byte-chunk tensor-core gather path
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) pad_channels_kernel(const TS* __restrict__ src, TD* __restrict__ dst,
                                                           unsigned long long rows, int C, int CP) {
  const unsigned long long total = rows * (unsigned)CP, stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned long long row = i / (unsigned)CP;
    const int c = (int)(i - row * (unsigned)CP);
    Io<TD>::st(dst + i, c < C ? Io<TS>::ld(src + row * C + c) : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int WG_BKM = 16;  // rows per stage
constexpr int WG_NT = 256;

struct WGParams {
  const void* x;         // (B, rows_in, Cin)
  const int32_t* table;  // (rows_out, S)
  const void* gz;        // (B, rows_out, Cout)
  float* ws;             // [splits][Cout*K + Cout] fp32 partials
  long long M;           // B*rows_out
  long long rows_per_split;
  int rows_in, rows_out, S, Cin, Cout, K;
  int tiles_b;           // number of tiles along Cin
};

// Output tile BA (Cout) x BB (Cin) for one slot s and one row split; thread tile TA x TB.
template <typename T, int BA, int BB, int TA, int TB>
__global__ void __launch_bounds__(WG_NT) wgrad_kernel(const WGParams p) {
  constexpr int TXB = BB / TB;
  static_assert((BA / TA) * TXB == WG_NT, "thread tiling must cover the tile");
  constexpr int G_PER_T = WG_BKM * BA / WG_NT, X_PER_T = WG_BKM * BB / WG_NT;
  static_assert(G_PER_T >= 1 && X_PER_T >= 1, "tile too small for the loader");
  __shared__ __align__(16) float Gs[WG_BKM][BA];
  __shared__ __align__(16) float Xs[WG_BKM][BB];

  const T* __restrict__ x = static_cast<const T*>(p.x);
  const T* __restrict__ gz = static_cast<const T*>(p.gz);
  const int t = threadIdx.x;
  const int split = blockIdx.x, s = blockIdx.y;
  const int tile_a = blockIdx.z / p.tiles_b, tile_b = blockIdx.z % p.tiles_b;
  const int a0 = tile_a * BA, c0 = tile_b * BB;
  const long long mbeg = (long long)split * p.rows_per_split;
  const long long mend = min(p.M, mbeg + p.rows_per_split);

  const int tx = t % TXB, ty = t / TXB;
  float acc[TA][TB];
  float bacc[TA];
#pragma unroll
  for (int i = 0; i < TA; ++i) {
    bacc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j] = 0.f;
  }
  const bool do_bias = (s == 0) && (tile_b == 0) && (tx == 0);

  float greg[G_PER_T], xreg[X_PER_T];
  auto load = [&](long long mm) {
#pragma unroll
    for (int i = 0; i < G_PER_T; ++i) {
      const int e = t + i * WG_NT;
      const int mk = e / BA, a = e - mk * BA;
      const long long m = mm + mk;
      greg[i] = (m < mend && a0 + a < p.Cout) ? Io<T>::ld(gz + m * p.Cout + a0 + a) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < X_PER_T; ++i) {
      const int e = t + i * WG_NT;
      const int mk = e / BB, c = e - mk * BB;
      const long long m = mm + mk;
      float v = 0.f;
      if (m < mend && c0 + c < p.Cin) {
        const long long b = m / p.rows_out;
        const int j = (int)(m - b * p.rows_out);
        const int row = __ldg(p.table + (long long)j * p.S + s);
        v = Io<T>::ld(x + (b * p.rows_in + row) * (long long)p.Cin + c0 + c);
      }
      xreg[i] = v;
    }
  };
  auto store = [&]() {
#pragma unroll
    for (int i = 0; i < G_PER_T; ++i) { const int e = t + i * WG_NT; Gs[e / BA][e % BA] = greg[i]; }
#pragma unroll
    for (int i = 0; i < X_PER_T; ++i) { const int e = t + i * WG_NT; Xs[e / BB][e % BB] = xreg[i]; }
  };

  if (mbeg < mend) {
    load(mbeg);
    store();
    __syncthreads();
    for (long long mm = mbeg; mm < mend; mm += WG_BKM) {
      const bool more = mm + WG_BKM < mend;
      if (more) load(mm + WG_BKM);
#pragma unroll
      for (int mk = 0; mk < WG_BKM; ++mk) {
        float a[TA], b[TB];
#pragma unroll
        for (int i = 0; i < TA; ++i) a[i] = Gs[mk][ty * TA + i];
#pragma unroll
        for (int j = 0; j < TB; ++j) b[j] = Xs[mk][tx * TB + j];
#pragma unroll
        for (int i = 0; i < TA; ++i) {
          if (do_bias) bacc[i] += a[i];
#pragma unroll
          for (int j = 0; j < TB; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
      }
      __syncthreads();
      if (more) { store(); __syncthreads(); }
    }
  }
  float* ws = p.ws + (long long)split * ((long long)p.Cout * p.K + p.Cout);
#pragma unroll
  for (int i = 0; i < TA; ++i) {
    const int n = a0 + ty * TA + i;
    if (n >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const int c = c0 + tx * TB + j;
      if (c < p.Cin) ws[(long long)n * p.K + (long long)s * p.Cin + c] = acc[i][j];
    }
    if (do_bias) ws[(long long)p.Cout * p.K + n] = bacc[i];
  }
}

// out[i] = sum over splits (ascending) of ws[split][i]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int splits, long long per_split,
                                                           long long n_w, float* __restrict__ gw, float* __restrict__ gb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_split) return;
  float a = 0.f;
  for (int sp = 0; sp < splits; ++sp) a += ws[(long long)sp * per_split + i];
  if (i < n_w) gw[i] = a;
  else if (gb != nullptr) gb[i - n_w] = a;
}

static int wgrad_splits(long long M, int S, int tiles) {
  // aim at ~4 CTAs per SM in flight, never fewer than 512 rows per split
  long long want = (4LL * kNumSMs + (long long)S * tiles - 1) / ((long long)S * tiles);
  long long cap = (M + 511) / 512;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

struct WGConfig { int ba, bb; };
static WGConfig wgrad_config(int Cin, int Cout) {
  // output tile BA (Cout) x BB (Cin), each dimension 16 / 32 / 64 by the layer's own size: a thread then owns a
  // (BA/16) x (BB/16) micro-tile instead of the 1 x 1 the old square 16 x 16 tile gave narrow layers
  auto pick = [](int c) { return c > 32 ? 64 : (c > 16 ? 32 : 16); };
  return {pick(Cout), pick(Cin)};
}

// ------------------------------------------------------------------------------------------------ dummy-row dgrad
// gx[b, rows_in-1, c] = sum_s ( sum_{j in inv(dummy, s)} gz[b,j,:] ) . W[:, s*Cin + c];  one CTA per sample.
template <typename T>
__global__ void __launch_bounds__(256) dgrad_dummy_kernel(const T* __restrict__ gz, const int32_t* __restrict__ keyptr,
                                                          const int32_t* __restrict__ rows, const T* __restrict__ w,
                                                          T* __restrict__ gx, int rows_in, int rows_out, int S, int Cin,
                                                          int Cout) {
  extern __shared__ float sm[];  // red[256] + G[Cout] + out[Cin]
  float* red = sm;
  float* G = sm + 256;
  float* out = G + Cout;
  const int b = blockIdx.x, t = threadIdx.x;
  const int u = rows_in - 1;
  const T* gzb = gz + (long long)b * rows_out * Cout;
  for (int c = t; c < Cin; c += 256) out[c] = 0.f;
  // lanes: groups of `Cout`-wide rows; thread t handles channel n = t % npad in entry-group t / npad
  int npad = 1;
  while (npad < Cout && npad < 256) npad <<= 1;
  const int groups = 256 / npad;
  const int n = t % npad, g = t / npad;
  const int K = S * Cin;
  for (int s = 0; s < S; ++s) {
    const int e0 = keyptr[(long long)u * S + s], e1 = keyptr[(long long)u * S + s + 1];
    for (int nb = 0; nb < Cout; nb += npad) {  // only loops when Cout > 256
      float a = 0.f;
      if (nb + n < Cout)
        for (int e = e0 + g; e < e1; e += groups) a += Io<T>::ld(gzb + (long long)rows[e] * Cout + nb + n);
      __syncthreads();
      red[t] = a;
      __syncthreads();
      if (g == 0 && nb + n < Cout) {
        float r = 0.f;
        for (int q = 0; q < groups; ++q) r += red[q * npad + n];
        G[nb + n] = r;
      }
    }
    __syncthreads();
    for (int c = t; c < Cin; c += 256) {
      float a = out[c];
      for (int nn = 0; nn < Cout; ++nn) a = fmaf(G[nn], Io<T>::ld(w + (long long)nn * K + (long long)s * Cin + c), a);
      out[c] = a;
    }
    __syncthreads();
  }
  for (int c = t; c < Cin; c += 256) Io<T>::st(gx + ((long long)b * rows_in + u) * Cin + c, out[c]);
}

// Vectorised version (Cout a multiple of V = 16 bytes' worth of T, S*Cout/V <= 512, Cin <= 512): all S slots at once.
// Phase 1: thread = (group g, slot s, V-wide channel chunk); entries e0+g, e0+g+NG, ... of key (dummy, s) in ascending
// order, four loads in flight; the NG groups are combined in ascending order.  Phase 2: out[c] = sum_s sum_n G[s][n] *
// W[n, s*Cin + c], slots strided over NQ = 512/Cin thread groups, combined in ascending order.  Deterministic.
template <typename T, int V>
__global__ void __launch_bounds__(512) dgrad_dummy_vec_kernel(const T* __restrict__ gz, const int32_t* __restrict__ keyptr,
                                                              const int32_t* __restrict__ rows, const T* __restrict__ w,
                                                              T* __restrict__ gx, int rows_in, int rows_out, int S, int Cin,
                                                              int Cout) {
  extern __shared__ float sm[];  // part[NG][S*Cout] | G[S*Cout] | outp[NQ][Cin]
  const int b = blockIdx.x, t = threadIdx.x;
  const int u = rows_in - 1;
  const int CPR = Cout / V, NI = S * CPR, NG = 512 / NI, SC = S * Cout;
  float* part = sm;
  float* G = part + (size_t)NG * SC;
  float* outp = G + SC;
  const T* gzb = gz + (long long)b * rows_out * Cout;
  if (t < NG * NI) {
    const int g = t / NI, item = t - g * NI, s = item / CPR, ch = item - s * CPR;
    const int e0 = keyptr[(long long)u * S + s], e1 = keyptr[(long long)u * S + s + 1];
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    for (int e = e0 + g; e < e1; e += 4 * NG) {
      float v[4][V];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int eq = e + q * NG;
        if (eq < e1) {
          const T* px = gzb + (long long)__ldg(rows + eq) * Cout + ch * V;
          if (V == 8) Io<T>::ld8(px, v[q]); else Io<T>::ld4(px, v[q]);
        } else {
#pragma unroll
          for (int k = 0; k < V; ++k) v[q][k] = 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += v[q][k];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) part[(size_t)g * SC + s * Cout + ch * V + k] = acc[k];
  }
  __syncthreads();
  for (int i = t; i < SC; i += 512) {
    float r = 0.f;
    for (int g = 0; g < NG; ++g) r += part[(size_t)g * SC + i];
    G[i] = r;
  }
  __syncthreads();
  const int NQ = 512 / Cin, K = S * Cin;
  if (t < NQ * Cin) {
    const int q = t / Cin, c = t - q * Cin;
    float a = 0.f;
    for (int s = q; s < S; s += NQ)
      for (int n = 0; n < Cout; ++n) a = fmaf(G[s * Cout + n], Io<T>::ld(w + (long long)n * K + (long long)s * Cin + c), a);
    outp[q * Cin + c] = a;
  }
  __syncthreads();
  if (t < Cin) {
    float a = 0.f;
    for (int q = 0; q < NQ; ++q) a += outp[q * Cin + t];
    Io<T>::st(gx + ((long long)b * rows_in + u) * Cin + t, a);
  }
}

template <typename T>
static int wgrad_launch(const WGParams& p0, int splits, void* gw, void* gb, cudaStream_t st) {
  WGParams p = p0;
  const WGConfig c = wgrad_config(p.Cin, p.Cout);
  const int tiles_a = ceil_div(p.Cout, c.ba);
  p.tiles_b = ceil_div(p.Cin, c.bb);
  dim3 grid((unsigned)splits, (unsigned)p.S, (unsigned)(tiles_a * p.tiles_b));
#define SHB_WG(A, B_) wgrad_kernel<T, A, B_, A / 16, B_ / 16><<<grid, WG_NT, 0, st>>>(p)
  switch (c.ba * 100 + c.bb) {
    case 6464: SHB_WG(64, 64); break;
    case 6432: SHB_WG(64, 32); break;
    case 6416: SHB_WG(64, 16); break;
    case 3264: SHB_WG(32, 64); break;
    case 3232: SHB_WG(32, 32); break;
    case 3216: SHB_WG(32, 16); break;
    case 1664: SHB_WG(16, 64); break;
    case 1632: SHB_WG(16, 32); break;
    default: SHB_WG(16, 16); break;
  }
#undef SHB_WG
  SHB_LAUNCH_CHECK();
  const long long n_w = (long long)p.Cout * p.K, per = n_w + p.Cout;
  wgrad_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(p.ws, splits, per, n_w, (float*)gw, (float*)gb);
  SHB_LAUNCH_CHECK();
  return 0;
}

bool umma_enabled() {
  static const bool on = [] {
    const char* e = getenv("SHB_DISABLE_UMMA");
    return !(e && e[0] == '1');
  }();
  return on;
}

}  // namespace shb

using namespace shb;

extern "C" {

// not part of the public ABI (not declared in include/shb200.h): perf-debug hook used by scripts/trace_umma.py
void shbdbg_set_trace(void* buf) { umma_set_trace((long long*)buf); }

int shb_spiralconv_fwd(const void* x, const int32_t* table, const void* w, const void* bias, void* y, int B,
                       int rows_in, int rows_out, int S, int Cin, int Cout, int act, int zero_last_row,
                       int src_dummy_zero, int dtype, void* stream) {
  if (!x || !table || !w || !y) return SHB_E_ARG;
  if (B <= 0 || rows_in <= 0 || rows_out <= 0 || S <= 0 || Cin <= 0 || Cout <= 0) return SHB_E_ARG;
  if (act < SHB_ACT_IDENTITY || act > SHB_ACT_TANH) return SHB_E_ARG;
  GGParams p{};
  p.src = x; p.table = table; p.list = nullptr; p.w = w; p.bias = bias; p.dst = y;
  p.M = (long long)B * rows_out;
  p.rows_src = rows_in; p.rows_dst = rows_out; p.S = S; p.Cs = Cin; p.Cd = Cout; p.K = S * Cin;
  p.act = act; p.zero_last = zero_last_row; p.skip_last = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SHB_F32) return launch_gather_gemm<float, false>(p, st);
  if (dtype == SHB_BF16) {
    // tensor-core path (tcgen05 + TMEM); its tile arithmetic is 32-bit, larger problems take the CUDA-core kernel
    if (umma_enabled() && p.M < (1LL << 31) && umma_gather_gemm_supported(Cin, Cout, S))
      return umma_gather_gemm(x, table, nullptr, nullptr, w, (const float*)bias, y, B, rows_in, rows_out, S, Cin, Cout, act,
                              zero_last_row, src_dummy_zero, false, st);
    return launch_gather_gemm<__nv_bfloat16, false>(p, st);
  }
  return SHB_E_DTYPE;
}

constexpr int ACT_MAX_BLOCKS = 16 * kNumSMs;

size_t shb_spiralconv_bwd_act_workspace(int gz_channels) {
  return ((size_t)ACT_MAX_BLOCKS + 1) * (size_t)(gz_channels > 0 ? gz_channels : 0) * sizeof(float);
}

int shb_spiralconv_bwd_act(const void* gy, const void* y, void* gz, int B, int rows_out, int Cout, int gz_channels,
                           int act, int zero_last_row, void* gb, void* workspace, size_t workspace_bytes, int dtype,
                           void* stream) {
  if (!gy || !y || !gz || B <= 0 || rows_out <= 0 || Cout <= 0 || gz_channels < Cout) return SHB_E_ARG;
  if (act < SHB_ACT_IDENTITY || act > SHB_ACT_TANH) return SHB_E_ARG;
  if (dtype != SHB_F32 && dtype != SHB_BF16) return SHB_E_DTYPE;
  if (gb != nullptr && (workspace == nullptr || workspace_bytes < shb_spiralconv_bwd_act_workspace(gz_channels)))
    return SHB_E_WORKSPACE;
  const unsigned long long rows = (unsigned long long)B * rows_out;
  if (rows >= (1ull << 32)) return SHB_E_SHAPE;
  const unsigned long long n = rows * gz_channels;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = dtype == SHB_BF16 ? 8 : 4;
  float* part = (float*)workspace;
  if (gz_channels == Cout && Cout % vec == 0 && n / vec < (1ull << 32)) {
    // 16-byte vectorised pass; the bias gradient rides along when a thread keeps one channel chunk
    const unsigned chunks = (unsigned)(n / vec), cpr = (unsigned)(Cout / vec);
    const unsigned want = (chunks + 255) / 256;
    const int vb = (int)(want < (unsigned)ACT_MAX_BLOCKS ? want : (unsigned)ACT_MAX_BLOCKS);
    const bool fuse = gb != nullptr && cpr <= 256 && (256 % cpr) == 0;
    if (dtype == SHB_F32)
      act_bwd_vec_kernel<float, 4><<<vb, 256, 0, st>>>((const float*)gy, (const float*)y, (float*)gz, chunks, cpr,
                                                       (unsigned)rows_out, act, zero_last_row, fuse ? part : nullptr);
    else
      act_bwd_vec_kernel<__nv_bfloat16, 8><<<vb, 256, 0, st>>>((const __nv_bfloat16*)gy, (const __nv_bfloat16*)y,
                                                               (__nv_bfloat16*)gz, chunks, cpr, (unsigned)rows_out, act,
                                                               zero_last_row, fuse ? part : nullptr);
    SHB_LAUNCH_CHECK();
    if (fuse) {
      partial_reduce_kernel<<<Cout, 256, 0, st>>>(part, vb, Cout, (float*)gb);
      SHB_LAUNCH_CHECK();
      return 0;
    }
  } else {
    const int blocks = (int)((n + 255) / 256 < (unsigned long long)ACT_MAX_BLOCKS ? (n + 255) / 256 : ACT_MAX_BLOCKS);
    if (dtype == SHB_F32)
      act_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)gy, (const float*)y, (float*)gz, (unsigned)rows,
                                                    (unsigned)rows_out, Cout, gz_channels, act, zero_last_row);
    else
      act_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)gy, (const __nv_bfloat16*)y,
                                                            (__nv_bfloat16*)gz, (unsigned)rows, (unsigned)rows_out, Cout,
                                                            gz_channels, act, zero_last_row);
    SHB_LAUNCH_CHECK();
  }
  if (gb == nullptr) return 0;
  // bias gradient from gz in a separate pass (odd channel counts / padded gz): two fixed-order stages
  int CP = 1;
  while (CP < gz_channels) CP <<= 1;
  if (CP > 256) return SHB_E_SHAPE;
  const int nb = 4 * kNumSMs;
  const unsigned long long rpb = (rows + nb - 1) / nb;
  float* tot = part + (size_t)nb * gz_channels;  // gz_channels floats after the partials
  if (dtype == SHB_F32)
    colsum_partial_generic<float><<<nb, 256, 0, st>>>((const float*)gz, rows, gz_channels, CP, rpb, part);
  else
    colsum_partial_generic<__nv_bfloat16><<<nb, 256, 0, st>>>((const __nv_bfloat16*)gz, rows, gz_channels, CP, rpb, part);
  SHB_LAUNCH_CHECK();
  partial_reduce_kernel<<<gz_channels, 256, 0, st>>>(part, nb, gz_channels, tot);
  SHB_LAUNCH_CHECK();
  const cudaError_t e = cudaMemcpyAsync(gb, tot, (size_t)Cout * sizeof(float), cudaMemcpyDeviceToDevice, st);
  return e == cudaSuccess ? 0 : (int)e;
}

int shb_pad_channels(const void* src, void* dst, int64_t rows, int C, int Cp, int dtype_src, int dtype_dst, void* stream) {
  if (!src || !dst || rows <= 0 || C <= 0 || Cp < C) return SHB_E_ARG;
  const unsigned long long n = (unsigned long long)rows * Cp;
  const int blocks = (int)((n + 255) / 256 < 16ULL * kNumSMs ? (n + 255) / 256 : 16ULL * kNumSMs);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_src == SHB_F32 && dtype_dst == SHB_F32)
    pad_channels_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, (float*)dst, rows, C, Cp);
  else if (dtype_src == SHB_F32 && dtype_dst == SHB_BF16)
    pad_channels_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, rows, C, Cp);
  else if (dtype_src == SHB_BF16 && dtype_dst == SHB_BF16)
    pad_channels_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)src,
                                                                               (__nv_bfloat16*)dst, rows, C, Cp);
  else if (dtype_src == SHB_BF16 && dtype_dst == SHB_F32)
    pad_channels_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, rows, C, Cp);
  else
    return SHB_E_DTYPE;
  SHB_LAUNCH_CHECK();
  return 0;
}

size_t shb_spiralconv_wgrad_workspace(int B, int rows_in, int rows_out, int S, int Cin, int Cout, int dtype) {
  (void)rows_in;
  if (dtype == SHB_BF16 && umma_enabled() && umma_wgrad_supported(Cin, Cout, S))
    return umma_wgrad_workspace(B, rows_out, S, Cin, Cout);
  const WGConfig c = wgrad_config(Cin, Cout);
  const int tiles = ceil_div(Cout, c.ba) * ceil_div(Cin, c.bb);
  const int splits = wgrad_splits((long long)B * rows_out, S, tiles);
  return (size_t)splits * ((size_t)Cout * S * Cin + Cout) * sizeof(float);
}

int shb_spiralconv_bwd_wgrad(const void* x, const int32_t* table, const void* gz, void* gw, void* gb,
                             void* workspace, size_t workspace_bytes, int B, int rows_in, int rows_out, int S,
                             int Cin, int Cout, int src_dummy_zero, int dtype, void* stream) {
  if (!x || !table || !gz || !gw || !workspace) return SHB_E_ARG;
  if (B <= 0 || rows_in <= 0 || rows_out <= 0 || S <= 0 || Cin <= 0 || Cout <= 0) return SHB_E_ARG;
  if (workspace_bytes < shb_spiralconv_wgrad_workspace(B, rows_in, rows_out, S, Cin, Cout, dtype)) return SHB_E_WORKSPACE;
  if (dtype == SHB_BF16 && umma_enabled() && umma_wgrad_supported(Cin, Cout, S))
    return umma_wgrad(x, table, gz, (float*)gw, (float*)gb, workspace, B, rows_in, rows_out, S, Cin, Cout,
                      src_dummy_zero, (cudaStream_t)stream);
  const WGConfig c = wgrad_config(Cin, Cout);
  const int tiles = ceil_div(Cout, c.ba) * ceil_div(Cin, c.bb);
  WGParams p{};
  p.x = x; p.table = table; p.gz = gz; p.ws = (float*)workspace;
  p.M = (long long)B * rows_out;
  const int splits = wgrad_splits(p.M, S, tiles);
  p.rows_per_split = ((p.M + splits - 1) / splits + WG_BKM - 1) / WG_BKM * WG_BKM;
  p.rows_in = rows_in; p.rows_out = rows_out; p.S = S; p.Cin = Cin; p.Cout = Cout; p.K = S * Cin;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SHB_F32) return wgrad_launch<float>(p, splits, gw, gb, st);
  if (dtype == SHB_BF16) return wgrad_launch<__nv_bfloat16>(p, splits, gw, gb, st);
  return SHB_E_DTYPE;
}

int shb_spiralconv_bwd_dgrad(const void* gz, const int32_t* keyptr, const int32_t* rows, const uint16_t* quads,
                             const void* w, void* gx,
                             int B, int rows_in, int rows_out, int S, int Cin, int Cout, int dummy_row_grad,
                             int dtype, void* stream) {
  if (!gz || !keyptr || !rows || !w || !gx) return SHB_E_ARG;
  if (B <= 0 || rows_in <= 0 || rows_out <= 0 || S <= 0 || Cin <= 0 || Cout <= 0) return SHB_E_ARG;
  GGParams p{};
  p.src = gz; p.table = keyptr; p.list = rows; p.w = w; p.bias = nullptr; p.dst = gx;
  p.M = (long long)B * rows_in;
  p.rows_src = rows_out; p.rows_dst = rows_in; p.S = S; p.Cs = Cout; p.Cd = Cin; p.K = S * Cout;
  p.act = SHB_ACT_IDENTITY; p.zero_last = 0; p.skip_last = 1;  // dummy row -> 0 here, filled below if wanted
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype == SHB_F32) rc = launch_gather_gemm<float, true>(p, st);
  else if (dtype == SHB_BF16) {
    if (umma_enabled() && quads != nullptr && p.M < (1LL << 31) && umma_gather_gemm_supported(Cout, Cin, S))
      rc = umma_gather_gemm(gz, (const int32_t*)quads, keyptr, rows, w, nullptr, gx, B, rows_out, rows_in, S, Cout, Cin, SHB_ACT_IDENTITY, 0, 1,
                            true, st);
    else
      rc = launch_gather_gemm<__nv_bfloat16, true>(p, st);
  } else return SHB_E_DTYPE;
  if (rc != 0 || !dummy_row_grad) return rc;
  const int V = dtype == SHB_F32 ? 4 : 8;
  if (Cout % V == 0 && S * (Cout / V) <= 512 && Cin <= 512) {
    const int NG = 512 / (S * (Cout / V)), NQ = 512 / Cin;
    const size_t smem = ((size_t)(NG + 1) * S * Cout + (size_t)NQ * Cin) * sizeof(float);  // <= 16 KB + 16 KB + 2 KB
    if (dtype == SHB_F32)
      dgrad_dummy_vec_kernel<float, 4><<<B, 512, smem, st>>>((const float*)gz, keyptr, rows, (const float*)w, (float*)gx,
                                                             rows_in, rows_out, S, Cin, Cout);
    else
      dgrad_dummy_vec_kernel<__nv_bfloat16, 8><<<B, 512, smem, st>>>((const __nv_bfloat16*)gz, keyptr, rows,
                                                                     (const __nv_bfloat16*)w, (__nv_bfloat16*)gx, rows_in,
                                                                     rows_out, S, Cin, Cout);
    SHB_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (256 + (size_t)Cout + Cin) * sizeof(float);
  if (dtype == SHB_F32)
    dgrad_dummy_kernel<float><<<B, 256, smem, st>>>((const float*)gz, keyptr, rows, (const float*)w, (float*)gx, rows_in,
                                                    rows_out, S, Cin, Cout);
  else
    dgrad_dummy_kernel<__nv_bfloat16><<<B, 256, smem, st>>>((const __nv_bfloat16*)gz, keyptr, rows,
                                                            (const __nv_bfloat16*)w, (__nv_bfloat16*)gx, rows_in,
                                                            rows_out, S, Cin, Cout);
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
