// Streaming kernels of the slab layout (shb_slab.cuh): Pool (CSR SpMM over whole slabs), and the two layout conversions at
// the ends of the trunks (caller's row-major (B, rows, C) <-> slab layout, with the vertex permutation, channel padding,
// dtype conversion and -- on the gradient path -- the activation derivative folded in).
//
// In the slab layout a Pool row is a weighted sum of <= a few whole slabs: every thread moves 16-byte vectors at the SAME
// offset of each source slab, so all traffic is perfectly coalesced, the CSR row is read once per (row, chunk) and there is
// no per-element index arithmetic (models.py:127,148: torch.matmul(D|U, x)).
#include "shb_common.cuh"
#include "shb_internal.h"
#include "shb_slab.cuh"

namespace shb {

using namespace slab;

// ---------------------------------------------------------------------------------------------- Pool
// dst[r] = act'(ymul[r]) * sum_k vals[k] * src[colidx[k]]   for k in rowptr[r] .. rowptr[r+1];   dummy row optionally zeroed
//
// A unit is one (row, chunk) slab, cut into nslice vector ranges when rows are few; ONE WARP owns a unit and streams its
// vectors, lane l taking vectors l, l + 32, ... of the range.  Two earlier versions (one thread per vector; a 256-thread
// block per unit) were bound by instruction issue and by the dependent load chain of a unit, not by bandwidth: ncu showed
// ~235 warp instructions per 16-byte output vector (unit decoding with three integer divisions, the row's indices and
// weights re-loaded and 64-bit addresses re-formed for every vector) at IPC 0.68, DRAM 30 %, L2 22 %.  Here everything that
// belongs to the ROW is done once per unit -- row range, indices, weights (one entry per lane, handed out by shuffles), the
// source pointers -- and is requested one unit ahead (row range: two units ahead), so the steady state of a warp is the
// vector loop alone: N + 1 independent 512-byte loads per vector (two vectors unrolled), 8 N fused multiply-adds, one store,
// pointers advanced by a constant.  The loop is specialised on the row's entry count N = 1..4 (rows of U have one or three
// entries; rows of a transpose have a handful); longer rows take four entries at a time.  CSR order is kept.
template <int P>
__device__ __forceinline__ void pool_load(const uint8_t* s, size_t plane_b, uint4* raw) {
  raw[0] = __ldg(reinterpret_cast<const uint4*>(s));
  if (P == 2) raw[P - 1] = __ldg(reinterpret_cast<const uint4*>(s + plane_b));
}
template <int P>
__device__ __forceinline__ void pool_value(const uint4* raw, float* x) {
  unpack8(raw[0], x);
  if (P == 2) {
    float l[8];
    unpack8(raw[P - 1], l);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += l[i];
  }
}
// acc (+)= sum_j w[j] * x_j for N entries whose data pointers are s[j] (+ off)
template <int P, int N, bool FIRST>
__device__ __forceinline__ void pool_accum(const uint8_t* const* s, const float* w, size_t off, size_t plane_b, float* acc) {
  uint4 raw[N][P];
#pragma unroll
  for (int j = 0; j < N; ++j) pool_load<P>(s[j] + off, plane_b, raw[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float x[8];
    pool_value<P>(raw[j], x);
    if (FIRST && j == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = w[0] * x[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w[j], x[i], acc[i]);
    }
  }
}
// the same for two consecutive vectors of a lane (off, off + 512): all 2 N loads first
template <int P, int N>
__device__ __forceinline__ void pool_accum2(const uint8_t* const* s, const float* w, size_t off, size_t plane_b, float (*acc)[8]) {
  uint4 raw[2][N][P];
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int j = 0; j < N; ++j) pool_load<P>(s[j] + off + v * 512, plane_b, raw[v][j]);
#pragma unroll
  for (int v = 0; v < 2; ++v)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      float x[8];
      pool_value<P>(raw[v][j], x);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] = fmaf(w[j], x[i], acc[v][i]);
    }
}
template <int P, int ACT>
__device__ __forceinline__ void pool_finish(float* acc, const uint4* yraw, uint8_t* d, size_t plane_b) {
  if (ACT != SHB_ACT_IDENTITY) {
    float yy[8];
    pool_value<P>(yraw, yy);
    act_bwd8_as<ACT>(acc, yy);
  }
  if (P == 1) {
    *reinterpret_cast<uint4*>(d) = pack8(acc);
  } else {
    uint4 hi, lo;
    split8(acc, hi, lo);
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + plane_b) = lo;
  }
}
// All `iters` vectors of a lane for a row of exactly N entries (N = 1..4): a branch-free body (the activation is a template
// parameter), unrolled twice so that the 2 (N + 1) loads of two vectors are in flight together.
template <int P, int ACT, int N>
__device__ __forceinline__ void pool_unit_fixed(const uint8_t* sv, size_t rstride, size_t plane_b, int col_l, float w_l,
                                                const uint8_t* yv, uint8_t* dv, int iters) {
  const uint8_t* s[N];
  float w[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    s[j] = sv + (size_t)__shfl_sync(0xffffffffu, col_l, j) * rstride;
    w[j] = __shfl_sync(0xffffffffu, w_l, j);
  }
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
    const size_t off = (size_t)i * 512;
    uint4 yraw[P];
    if (ACT != SHB_ACT_IDENTITY) pool_load<P>(yv + off, plane_b, yraw);
    float acc[8];
    pool_accum<P, N, true>(s, w, off, plane_b, acc);
    pool_finish<P, ACT>(acc, yraw, dv + off, plane_b);
  }
}
// Rows of 5..32 entries: four entries at a time, the indices / weights handed out by shuffles from the lanes that hold them;
// two vectors per pass so that eight loads are in flight (`iters` is even: host).
template <int P, int ACT>
__device__ __forceinline__ void pool_unit_long(const uint8_t* sv, size_t rstride, size_t plane_b, int col_l, float w_l, int n,
                                               const uint8_t* yv, uint8_t* dv, int iters) {
#pragma unroll 1
  for (int i = 0; i < iters; i += 2) {
    const size_t off = (size_t)i * 512;
    uint4 yraw[2][P];
    if (ACT != SHB_ACT_IDENTITY) {
      pool_load<P>(yv + off, plane_b, yraw[0]);
      pool_load<P>(yv + off + 512, plane_b, yraw[1]);
    }
    float acc[2][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < n; kb += 4) {
      const uint8_t* s[4];
      float w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] = sv + (size_t)__shfl_sync(0xffffffffu, col_l, (kb + j) & 31) * rstride;
        w[j] = __shfl_sync(0xffffffffu, w_l, (kb + j) & 31);
      }
      switch (n - kb) {
        case 1: pool_accum2<P, 1>(s, w, off, plane_b, acc); break;
        case 2: pool_accum2<P, 2>(s, w, off, plane_b, acc); break;
        case 3: pool_accum2<P, 3>(s, w, off, plane_b, acc); break;
        default: pool_accum2<P, 4>(s, w, off, plane_b, acc); break;
      }
    }
    pool_finish<P, ACT>(acc[0], yraw[0], dv + off, plane_b);
    pool_finish<P, ACT>(acc[1], yraw[1], dv + off + 512, plane_b);
  }
}
// Rows longer than a warp (no matrix of the models has one): entries fetched inside the vector loop, no overlap.
template <int P, int ACT>
__device__ __forceinline__ void pool_unit_huge(const uint8_t* sv, size_t rstride, size_t plane_b,
                                               const int32_t* __restrict__ colidx, const float* __restrict__ vals, int k0, int n,
                                               const uint8_t* yv, uint8_t* dv, int iters) {
  constexpr bool with_y = ACT != SHB_ACT_IDENTITY;
  for (int i = 0; i < iters; ++i) {
    const size_t off = (size_t)i * 512;
    uint4 yraw[P];
    if (with_y) pool_load<P>(yv + off, plane_b, yraw);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int k = 0; k < n; ++k) {
      uint4 raw[P];
      pool_load<P>(sv + (size_t)__ldg(colidx + k0 + k) * rstride + off, plane_b, raw);
      const float w = __ldg(vals + k0 + k);
      float x[8];
      pool_value<P>(raw, x);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, x[e], acc[e]);
    }
    pool_finish<P, ACT>(acc, yraw, dv + off, plane_b);
  }
}

constexpr int POOL_WARPS = 8;   // per block

template <int P, int ACT>
__global__ void __launch_bounds__(POOL_WARPS * 32, P == 1 ? 3 : 2) slab_pool_kernel(const uint8_t* __restrict__ src,
                                                                    const int32_t* __restrict__ rowptr,
                                                                    const int32_t* __restrict__ colidx,
                                                                    const float* __restrict__ vals, uint8_t* __restrict__ dst,
                                                                    const uint8_t* __restrict__ ymul, int NB, int rows_out, int C,
                                                                    int zero_last, int nslice) {
  const int lane = threadIdx.x & 31;
  const int nvec = C * 16;              // 16-byte vectors per plane of a slab
  const int vper = nvec / nslice;       // vectors per slice: a multiple of 32 (host)
  const int iters = vper >> 5;
  const size_t slab_b = slab_bytes(C, P);
  const size_t plane_b = (size_t)C * 256;
  const size_t rstride = (size_t)NB * slab_b;
  const int per_row = NB * nslice;
  const int units = rows_out * per_row;
  const int stride = gridDim.x * POOL_WARPS;
  int unit = blockIdx.x * POOL_WARPS + (threadIdx.x >> 5);
  if (unit >= units) return;
  // software pipeline over this warp's units: row range two units ahead, entries (one per lane) one unit ahead
  int r = unit / per_row;
  int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
  int r_n = 0, n0 = 0, n1 = 0;
  if (unit + stride < units) {
    r_n = (unit + stride) / per_row;
    n0 = __ldg(rowptr + r_n);
    n1 = __ldg(rowptr + r_n + 1);
  }
  int col_l = 0;
  float w_l = 0.f;
  if (lane < k1 - k0) {
    col_l = __ldg(colidx + k0 + lane);
    w_l = __ldg(vals + k0 + lane);
  }
  while (true) {
    const int unit_n = unit + stride, unit_nn = unit_n + stride;
    int r_nn = 0, m0 = 0, m1 = 0;
    if (unit_nn < units) {
      r_nn = unit_nn / per_row;
      m0 = __ldg(rowptr + r_nn);
      m1 = __ldg(rowptr + r_nn + 1);
    }
    int col_n = 0;
    float w_n = 0.f;
    if (unit_n < units && lane < n1 - n0) {
      col_n = __ldg(colidx + n0 + lane);
      w_n = __ldg(vals + n0 + lane);
    }
    const int rem = unit - r * per_row, q = rem / nslice, sl = rem - q * nslice;
    const bool zero = zero_last && r == rows_out - 1;
    const int n = zero ? 0 : k1 - k0;
    const size_t voff = ((size_t)sl * vper + lane) * 16;
    const size_t uoff = ((size_t)r * NB + q) * slab_b + voff;
    const uint8_t* sv = src + (size_t)q * slab_b + voff;
    const uint8_t* yv = ymul + uoff;   // read only when ACT != identity
    uint8_t* dv = dst + uoff;
    switch (n) {
      case 0:
        for (int i = 0; i < iters; ++i) {
          *reinterpret_cast<uint4*>(dv + (size_t)i * 512) = make_uint4(0, 0, 0, 0);
          if (P == 2) *reinterpret_cast<uint4*>(dv + (size_t)i * 512 + plane_b) = make_uint4(0, 0, 0, 0);
        }
        break;
      case 1: pool_unit_fixed<P, ACT, 1>(sv, rstride, plane_b, col_l, w_l, yv, dv, iters); break;
      case 2: pool_unit_fixed<P, ACT, 2>(sv, rstride, plane_b, col_l, w_l, yv, dv, iters); break;
      case 3: pool_unit_fixed<P, ACT, 3>(sv, rstride, plane_b, col_l, w_l, yv, dv, iters); break;
      case 4: pool_unit_fixed<P, ACT, 4>(sv, rstride, plane_b, col_l, w_l, yv, dv, iters); break;
      default:
        if (n <= 32) pool_unit_long<P, ACT>(sv, rstride, plane_b, col_l, w_l, n, yv, dv, iters);
        else pool_unit_huge<P, ACT>(sv, rstride, plane_b, colidx, vals, k0, n, yv, dv, iters);
        break;
    }
    if (unit_n >= units) break;
    unit = unit_n; r = r_n; k0 = n0; k1 = n1; col_l = col_n; w_l = w_n;
    r_n = r_nn; n0 = m0; n1 = m1;
  }
}

// ---------------------------------------------------------------------------------------------- rows -> slabs
// src: row-major (B, R, Cs), fp32 or bf16.  dst: slab tensor (R, B, Cp), Cp >= Cs padded with zeros; internal row i takes the
// caller's row perm[i] (perm == null: identity).  Optional: multiply by act'(ymul) (ymul: slab tensor shaped like dst) and zero
// the last row.  Samples beyond B (tail chunk) are written as zeros.  One thread per (row, chunk, 8-channel group, sample).
template <typename T, int P>
__global__ void __launch_bounds__(256) slab_from_rows_kernel(const T* __restrict__ src, const int32_t* __restrict__ perm,
                                                             uint8_t* __restrict__ dst, const uint8_t* __restrict__ ymul, int B,
                                                             int R, int Cs, int Cp, int act_mul, int zero_last) {
  const int NB = num_chunks(B), ncc = Cp / 8;
  // Work item = (row r, channel group cc), cc fastest: consecutive items are consecutive in the caller's row-major tensor.
  // A warp covers 8 items x 4 samples: row-major accesses are runs of 8 items (>= 96 B) per sample, slab accesses are runs
  // of 4 samples (64 B) per item -- whole sectors on both sides (one thread per sample made the row-major side one
  // 16-byte access per sector).
  const long long items = (long long)R * ncc, iblocks = (items + 7) / 8;
  const long long total = iblocks * 8 * CHUNK * NB;
  const size_t slab_b = slab_bytes(Cp, P);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // (single-channel-group tensors -- the 3-channel ends -- keep one warp = 32 samples of one item: their row-major side is
    // 12-byte pieces scattered by the vertex permutation either way, and the slab side then gets 512-byte runs)
    const int il = ncc >= 8 ? (int)(i & 7) : (int)((i >> 7) & 7), bl = ncc >= 8 ? (int)((i >> 3) & (CHUNK - 1)) : (int)(i & (CHUNK - 1));
    long long t = i >> 10;   // 8 items x 128 samples per block of work
    const long long ib = t % iblocks;
    const int q = (int)(t / iblocks);
    const long long item = ib * 8 + il;
    if (item >= items) continue;
    const int r = (int)(item / ncc), cc = (int)(item - (long long)r * ncc);
    const int b = q * CHUNK + bl;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    const bool live = b < B && !(zero_last && r == R - 1);
    const size_t doff = ((size_t)r * NB + q) * slab_b + (size_t)cc * PLANE_STRIDE + (size_t)bl * 16;
    if (live) {
      const int rs = perm != nullptr ? __ldg(perm + r) : r;
      const T* s = src + ((size_t)b * R + rs) * Cs + cc * 8;
      if (cc * 8 + 8 <= Cs && (Cs & 7) == 0) {
        Io<T>::ld8(s, v);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (cc * 8 + e < Cs) v[e] = Io<T>::ld(s + e);
      }
      if (ymul != nullptr) {
        float y[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(ymul + doff)), y);
        if (P == 2) {
          float l[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(ymul + doff + (size_t)Cp * 256)), l);
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] += l[e];
        }
        act_bwd8(v, y, act_mul);
      }
    }
    if (P == 1) {
      *reinterpret_cast<uint4*>(dst + doff) = pack8(v);
    } else {
      uint4 hi, lo;
      split8(v, hi, lo);
      *reinterpret_cast<uint4*>(dst + doff) = hi;
      *reinterpret_cast<uint4*>(dst + doff + (size_t)Cp * 256) = lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------- slabs -> rows
// src: slab tensor (R, B, Cp).  dst: row-major (B, R, Cd), Cd <= Cp, fp32 or bf16; the caller's row perm[i] receives internal
// row i.
template <typename T, int P>
__global__ void __launch_bounds__(256) slab_to_rows_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ perm,
                                                           T* __restrict__ dst, int B, int R, int Cp, int Cd) {
  const int NB = num_chunks(B), ncc = (Cd + 7) / 8;
  // same work decomposition as slab_from_rows_kernel: a warp = 8 (row, channel group) items x 4 samples
  const long long items = (long long)R * ncc, iblocks = (items + 7) / 8;
  const long long total = iblocks * 8 * CHUNK * NB;
  const size_t slab_b = slab_bytes(Cp, P);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int il = (int)(i & 7), bl = (int)((i >> 3) & (CHUNK - 1));
    long long t = i >> 10;
    const long long ib = t % iblocks;
    const int q = (int)(t / iblocks);
    const long long item = ib * 8 + il;
    if (item >= items) continue;
    const int r = (int)(item / ncc), cc = (int)(item - (long long)r * ncc);
    const int b = q * CHUNK + bl;
    if (b >= B) continue;
    const size_t soff = ((size_t)r * NB + q) * slab_b + (size_t)cc * PLANE_STRIDE + (size_t)bl * 16;
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + soff)), v);
    if (P == 2) {
      float l[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + soff + (size_t)Cp * 256)), l);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += l[e];
    }
    const int rd = perm != nullptr ? __ldg(perm + r) : r;
    T* d = dst + ((size_t)b * R + rd) * Cd + cc * 8;
    if (cc * 8 + 8 <= Cd && (Cd & 3) == 0) {
      Io<T>::st4(d, v);
      Io<T>::st4(d + 4, v + 4);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (cc * 8 + e < Cd) Io<T>::st(d + e, v[e]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- tiled conversions
// The two kernels above touch whole sectors on both sides but spend ~300 instructions of 64-bit index arithmetic per 16-byte
// item and move 64-byte runs; at the FC boundary (432 rows x 128 channels) and at the 3-channel ends they ran at 1.2-2.9 TB/s.
// The kernels below take a TILE through shared memory instead, so that both sides see long contiguous runs and the index
// arithmetic is per tile:
//   wide   (Cs == Cp in {16, 32, 64, 128, 256}): tile = one (row, chunk) x <= 128 channels, staged as fp32
//          [channel group][sample][8] (group stride padded by 16 B); row-major side: (<= 256 B) x 128 samples per tile,
//          slab side: 2 KB per channel group;
//   narrow (Cp == 8, a few channels, any vertex permutation): tile = 32 consecutive rows IN THE CALLER'S NUMBERING x one
//          chunk, staged as [sample][row x channel]; row-major side: 32 rows x Cs contiguous elements per sample (coalesced
//          4-byte accesses), slab side: 2 KB per row, scattered by the inverse permutation.
constexpr int CV_THREADS = 256;
constexpr int CV_GROUP_STRIDE = CHUNK * 32 + 16;   // bytes between channel groups of a wide tile (fp32 x 8 per sample, + pad)

template <int P> __device__ __forceinline__ void cv_load_slab8(const uint8_t* p, size_t plane_b, float* v) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), v);
  if (P == 2) {
    float l[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p + plane_b)), l);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += l[e];
  }
}
template <int P> __device__ __forceinline__ void cv_store_slab8(uint8_t* p, size_t plane_b, const float* v) {
  if (P == 1) {
    *reinterpret_cast<uint4*>(p) = pack8(v);
  } else {
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + plane_b) = lo;
  }
}

// LG = log2(channel groups per tile) (1..4); tiles = R * NB * nsl, nsl = Cp / (8 << LG)
template <typename T, int P, int LG>
__global__ void __launch_bounds__(CV_THREADS, 3) slab_from_rows_wide_kernel(const T* __restrict__ src, const int32_t* __restrict__ perm,
                                                                         uint8_t* __restrict__ dst,
                                                                         const uint8_t* __restrict__ ymul, int B, int R, int Cp,
                                                                         int act_mul, int zero_last, int nsl) {
  extern __shared__ __align__(16) uint8_t cv_smem[];
  constexpr int G = 1 << LG;                       // channel groups per tile
  constexpr int ITEMS = CHUNK * G / CV_THREADS;    // (group, sample) items per thread and phase
  const int NB = num_chunks(B), tid = threadIdx.x;
  const size_t slab_b = slab_bytes(Cp, P), plane_b = (size_t)Cp * 256;
  const int tiles = R * NB * nsl;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int sl = tile % nsl, rq = tile / nsl, q = rq % NB, r = rq / NB;
    const bool zero = zero_last && r == R - 1;
    const int rs = perm != nullptr ? __ldg(perm + r) : r;
    // phase 1: row-major side, channel group fastest (a sample's channels are contiguous)
    if constexpr (sizeof(T) == 2) {   // bf16 rows: the eight loads stay raw (4 registers each) until they are staged
      const int cc = tid & (G - 1);
      uint4 raw[ITEMS];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int bl = (tid >> LG) + k * (CV_THREADS >> LG), b = q * CHUNK + bl;
        raw[k] = (b < B && !zero) ? __ldg(reinterpret_cast<const uint4*>(src + ((size_t)b * R + rs) * Cp + (size_t)(sl * G + cc) * 8))
                                  : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int bl = (tid >> LG) + k * (CV_THREADS >> LG);
        float v[8];
        unpack8(raw[k], v);
        float4* d = reinterpret_cast<float4*>(cv_smem + (size_t)cc * CV_GROUP_STRIDE + (size_t)bl * 32);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      const int cc = tid & (G - 1);
      float v[ITEMS][8];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int bl = (tid >> LG) + k * (CV_THREADS >> LG), b = q * CHUNK + bl;
        if (b < B && !zero) {
          Io<T>::ld8(src + ((size_t)b * R + rs) * Cp + (size_t)(sl * G + cc) * 8, v[k]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[k][e] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int bl = (tid >> LG) + k * (CV_THREADS >> LG);
        float4* d = reinterpret_cast<float4*>(cv_smem + (size_t)cc * CV_GROUP_STRIDE + (size_t)bl * 32);
        d[0] = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
        d[1] = make_float4(v[k][4], v[k][5], v[k][6], v[k][7]);
      }
    }
    __syncthreads();
    // phase 2: slab side, sample fastest; the producer's outputs for act' are requested four items ahead (all eight would
    // cost a third resident CTA: measured 21 -> 30 us)
    const bool with_y = ymul != nullptr && !zero;
    constexpr int HB = ITEMS < 4 ? ITEMS : 4;
#pragma unroll
    for (int k0 = 0; k0 < ITEMS; k0 += HB) {
      uint4 yraw[HB][P];
      if (with_y) {
#pragma unroll
        for (int u = 0; u < HB; ++u) {
          const int it = tid + (k0 + u) * CV_THREADS, bl = it & (CHUNK - 1), cc = it >> 7;
          const uint8_t* a = ymul + ((size_t)r * NB + q) * slab_b + (size_t)(sl * G + cc) * PLANE_STRIDE + (size_t)bl * 16;
#pragma unroll
          for (int pl = 0; pl < P; ++pl) yraw[u][pl] = __ldg(reinterpret_cast<const uint4*>(a + pl * plane_b));
        }
      }
#pragma unroll
      for (int u = 0; u < HB; ++u) {
        const int it = tid + (k0 + u) * CV_THREADS, bl = it & (CHUNK - 1), cc = it >> 7;
        const float4* t4 = reinterpret_cast<const float4*>(cv_smem + (size_t)cc * CV_GROUP_STRIDE + (size_t)bl * 32);
        const float4 a = t4[0], c = t4[1];
        float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        const size_t doff = ((size_t)r * NB + q) * slab_b + (size_t)(sl * G + cc) * PLANE_STRIDE + (size_t)bl * 16;
        if (with_y && q * CHUNK + bl < B) {
          float y[8];
          unpack8(yraw[u][0], y);
          if (P == 2) {
            float l[8];
            unpack8(yraw[u][P - 1], l);
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] += l[e];
          }
          act_bwd8(v, y, act_mul);
        }
        cv_store_slab8<P>(dst + doff, plane_b, v);
      }
    }
    __syncthreads();
  }
}

template <typename T, int P, int LG>
__global__ void __launch_bounds__(CV_THREADS, 3) slab_to_rows_wide_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ perm,
                                                                       T* __restrict__ dst, int B, int R, int Cp, int nsl) {
  extern __shared__ __align__(16) uint8_t cv_smem[];
  constexpr int G = 1 << LG;
  constexpr int ITEMS = CHUNK * G / CV_THREADS;
  const int NB = num_chunks(B), tid = threadIdx.x;
  const size_t slab_b = slab_bytes(Cp, P), plane_b = (size_t)Cp * 256;
  const int tiles = R * NB * nsl;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int sl = tile % nsl, rq = tile / nsl, q = rq % NB, r = rq / NB;
    const int rd = perm != nullptr ? __ldg(perm + r) : r;
    {
      // all of a thread's slab vectors are requested before the first is unpacked (kept as raw 16-byte words: unpacked at
      // load they cost 8 registers each and the compiler then interleaves loads and shared-memory stores, two in flight)
      uint4 raw[ITEMS][P];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int it = tid + k * CV_THREADS, bl = it & (CHUNK - 1), cc = it >> 7;
        const uint8_t* a = src + ((size_t)r * NB + q) * slab_b + (size_t)(sl * G + cc) * PLANE_STRIDE + (size_t)bl * 16;
#pragma unroll
        for (int pl = 0; pl < P; ++pl) raw[k][pl] = __ldg(reinterpret_cast<const uint4*>(a + pl * plane_b));
      }
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int it = tid + k * CV_THREADS, bl = it & (CHUNK - 1), cc = it >> 7;
        float v[8];
        unpack8(raw[k][0], v);
        if (P == 2) {
          float l[8];
          unpack8(raw[k][P - 1], l);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += l[e];
        }
        float4* d = reinterpret_cast<float4*>(cv_smem + (size_t)cc * CV_GROUP_STRIDE + (size_t)bl * 32);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();
    const int cc = tid & (G - 1);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int bl = (tid >> LG) + k * (CV_THREADS >> LG), b = q * CHUNK + bl;
      if (b < B) {
        const float4* t4 = reinterpret_cast<const float4*>(cv_smem + (size_t)cc * CV_GROUP_STRIDE + (size_t)bl * 32);
        const float4 a = t4[0], c = t4[1];
        const float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        T* d = dst + ((size_t)b * R + rd) * Cp + (size_t)(sl * G + cc) * 8;
        Io<T>::st4(d, v);
        Io<T>::st4(d + 4, v + 4);
      }
    }
    __syncthreads();
  }
}

constexpr int CV_ROWS = 32;   // rows (caller's numbering) per narrow tile

// pos[c] = internal row of the caller's row c (the inverse of perm; null: identity).  Cs <= 8 channels, Cp == 8.
template <typename T, int P>
__global__ void __launch_bounds__(CV_THREADS, 3) slab_from_rows_narrow_kernel(const T* __restrict__ src, const int32_t* __restrict__ pos,
                                                                           uint8_t* __restrict__ dst,
                                                                           const uint8_t* __restrict__ ymul, int B, int R, int Cs,
                                                                           int act_mul, int zero_last) {
  extern __shared__ __align__(16) uint8_t cv_smem[];
  float* t = reinterpret_cast<float*>(cv_smem);
  const int NB = num_chunks(B), tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = (CV_ROWS * Cs) | 1;            // floats per sample in the tile: odd, so that a warp's samples hit 32 banks
  const size_t slab_b = slab_bytes(8, P), plane_b = 8 * 256;
  const int rblocks = (R + CV_ROWS - 1) / CV_ROWS, tiles = rblocks * NB;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int q = tile / rblocks, c0 = (tile - q * rblocks) * CV_ROWS;
    const int nrows = R - c0 < CV_ROWS ? R - c0 : CV_ROWS, len = nrows * Cs;
    // phase 1: a warp takes samples warp, warp + 8, ...; its lanes walk the len contiguous elements of a sample, the loads of
    // all 16 samples of the warp in flight together
    for (int e = lane; e < len; e += 32) {
      float tmp[CHUNK / (CV_THREADS / 32)];
#pragma unroll
      for (int k = 0; k < CHUNK / (CV_THREADS / 32); ++k) {
        const int b = q * CHUNK + warp + k * (CV_THREADS / 32);
        tmp[k] = b < B ? Io<T>::ld(src + ((size_t)b * R + c0) * Cs + e) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < CHUNK / (CV_THREADS / 32); ++k) t[(warp + k * (CV_THREADS / 32)) * stride + e] = tmp[k];
    }
    __syncthreads();
    // phase 2: one item = (row, sample), sample fastest: 2 KB contiguous per row on the slab side
    for (int it = tid; it < nrows * CHUNK; it += CV_THREADS) {
      const int j = it >> 7, bl = it & (CHUNK - 1);
      const int ri = pos != nullptr ? __ldg(pos + c0 + j) : c0 + j;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      const size_t doff = ((size_t)ri * NB + q) * slab_b + (size_t)bl * 16;
      if (q * CHUNK + bl < B && !(zero_last && ri == R - 1)) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e < Cs) v[e] = t[bl * stride + j * Cs + e];
        if (ymul != nullptr) {
          float y[8];
          cv_load_slab8<P>(ymul + doff, plane_b, y);
          act_bwd8(v, y, act_mul);
        }
      }
      cv_store_slab8<P>(dst + doff, plane_b, v);
    }
    __syncthreads();
  }
}

template <typename T, int P>
__global__ void __launch_bounds__(CV_THREADS, 3) slab_to_rows_narrow_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ pos,
                                                                         T* __restrict__ dst, int B, int R, int Cd) {
  extern __shared__ __align__(16) uint8_t cv_smem[];
  __shared__ int pos_s[CV_ROWS];
  float* t = reinterpret_cast<float*>(cv_smem);
  const int NB = num_chunks(B), tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = (CV_ROWS * Cd) | 1;
  const size_t slab_b = slab_bytes(8, P), plane_b = 8 * 256;
  const int rblocks = (R + CV_ROWS - 1) / CV_ROWS, tiles = rblocks * NB;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int q = tile / rblocks, c0 = (tile - q * rblocks) * CV_ROWS;
    const int nrows = R - c0 < CV_ROWS ? R - c0 : CV_ROWS, len = nrows * Cd;
    if (tid < nrows) pos_s[tid] = pos != nullptr ? __ldg(pos + c0 + tid) : c0 + tid;
    __syncthreads();
    // four items of a thread at a time, their slab vectors requested together (row positions from shared memory: fetched
    // per item from global memory they put a dependent load in front of every slab load)
    for (int it0 = tid; it0 < nrows * CHUNK; it0 += 4 * CV_THREADS) {
      uint4 raw[4][P];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * CV_THREADS;
        if (it < nrows * CHUNK) {
          const uint8_t* a = src + ((size_t)pos_s[it >> 7] * NB + q) * slab_b + (size_t)(it & (CHUNK - 1)) * 16;
#pragma unroll
          for (int pl = 0; pl < P; ++pl) raw[u][pl] = __ldg(reinterpret_cast<const uint4*>(a + pl * plane_b));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * CV_THREADS;
        if (it < nrows * CHUNK) {
          const int j = it >> 7, bl = it & (CHUNK - 1);
          float v[8];
          unpack8(raw[u][0], v);
          if (P == 2) {
            float l[8];
            unpack8(raw[u][P - 1], l);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += l[e];
          }
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < Cd) t[bl * stride + j * Cd + e] = v[e];
        }
      }
    }
    __syncthreads();
    for (int bl = warp; bl < CHUNK; bl += CV_THREADS / 32) {
      const int b = q * CHUNK + bl;
      if (b < B) {
        T* d = dst + ((size_t)b * R + c0) * Cd;
        for (int e = lane; e < len; e += 32) Io<T>::st(d + e, t[bl * stride + e]);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- L1 loss on a slab tensor
// loss = mean |rec - target| (train_funcs.py:501 on the output of models.py:159) with rec still in the slab layout (Cp == 8,
// Cs <= 8 real channels) and target row-major (B, R, Cs) in the caller's row order: the row-major reconstruction, its
// gradient and the two conversions between them never touch memory.  Tiles as in the narrow conversions (32 caller rows x
// one chunk, target staged in shared memory).  MODE 0: per-CTA partial sums of |d| (fixed order: a CTA's tiles in
// ascending order, block_sum; slab_l1_final_kernel adds the CTA partials in index order).  MODE 1: the gradient slab
// gscale / n * sign(d), times act'(rec) of the producer, dummy row zeroed when the producer masks it.
template <typename T, int P, int MODE>
__global__ void __launch_bounds__(CV_THREADS, 3) slab_l1_kernel(const uint8_t* __restrict__ rec, const T* __restrict__ target,
                                                             const int32_t* __restrict__ pos, float* __restrict__ partials,
                                                             const float* __restrict__ gscale, uint8_t* __restrict__ gdst, int B,
                                                             int R, int Cs, float n_elems, int act, int zero_last) {
  extern __shared__ __align__(16) uint8_t cv_smem[];
  __shared__ float red[CV_THREADS / 32];
  __shared__ int pos_s[CV_ROWS];   // internal row of each of the tile's rows: read once per tile, not once per item
  float* t = reinterpret_cast<float*>(cv_smem);
  const int NB = num_chunks(B), tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = (CV_ROWS * Cs) | 1;
  const size_t slab_b = slab_bytes(8, P), plane_b = 8 * 256;
  const int rblocks = (R + CV_ROWS - 1) / CV_ROWS, tiles = rblocks * NB;
  const float g = MODE == 1 ? __ldg(gscale) / n_elems : 0.f;   // as shb_l1_loss_bwd forms it: the two paths agree to the bit
  float acc = 0.f;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int q = tile / rblocks, c0 = (tile - q * rblocks) * CV_ROWS;
    const int nrows = R - c0 < CV_ROWS ? R - c0 : CV_ROWS, len = nrows * Cs;
    if (tid < nrows) pos_s[tid] = pos != nullptr ? __ldg(pos + c0 + tid) : c0 + tid;
    for (int e = lane; e < len; e += 32) {
      float tmp[CHUNK / (CV_THREADS / 32)];
#pragma unroll
      for (int k = 0; k < CHUNK / (CV_THREADS / 32); ++k) {
        const int b = q * CHUNK + warp + k * (CV_THREADS / 32);
        tmp[k] = b < B ? Io<T>::ld(target + ((size_t)b * R + c0) * Cs + e) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < CHUNK / (CV_THREADS / 32); ++k) t[(warp + k * (CV_THREADS / 32)) * stride + e] = tmp[k];
    }
    __syncthreads();
    // four items of a thread at a time: their slab vectors are requested together (unrolled by hand -- left to the
    // compiler the loads end up interleaved with the arithmetic that waits for them)
    constexpr int NBATCH = 4;
    for (int it0 = tid; it0 < nrows * CHUNK; it0 += NBATCH * CV_THREADS) {
      uint4 raw[NBATCH][P];
      size_t offs[NBATCH];
      int ris[NBATCH];
#pragma unroll
      for (int u = 0; u < NBATCH; ++u) {
        const int it = it0 + u * CV_THREADS;
        ris[u] = -1;
        offs[u] = 0;
        if (it < nrows * CHUNK) {
          const int j = it >> 7;
          ris[u] = pos_s[j];
          offs[u] = ((size_t)ris[u] * NB + q) * slab_b + (size_t)(it & (CHUNK - 1)) * 16;
#pragma unroll
          for (int pl = 0; pl < P; ++pl) raw[u][pl] = __ldg(reinterpret_cast<const uint4*>(rec + offs[u] + pl * plane_b));
        }
      }
#pragma unroll
      for (int u = 0; u < NBATCH; ++u) {
      if (ris[u] < 0) continue;
      const int it = it0 + u * CV_THREADS;
      const int j = it >> 7, bl = it & (CHUNK - 1);
      const int ri = ris[u];
      const size_t off = offs[u];
      const bool live = q * CHUNK + bl < B;
      float y[8];
      unpack8(raw[u][0], y);
      if (P == 2) {
        float l[8];
        unpack8(raw[u][P - 1], l);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] += l[e];
      }
      if (MODE == 0) {
        if (live) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < Cs) acc += fabsf(y[e] - t[bl * stride + j * Cs + e]);
        }
      } else {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (live && !(zero_last && ri == R - 1)) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < Cs) {
              const float d = y[e] - t[bl * stride + j * Cs + e];
              v[e] = d > 0.f ? g : (d < 0.f ? -g : 0.f);
            }
          if (act != SHB_ACT_IDENTITY) act_bwd8(v, y, act);
        }
        cv_store_slab8<P>(gdst + off, plane_b, v);
      }
      }
    }
    __syncthreads();
  }
  if (MODE == 0) {
    acc = block_sum<CV_THREADS>(acc, red);
    if (tid == 0) partials[blockIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(256) slab_l1_final_kernel(const float* __restrict__ partials, int np, float inv_n,
                                                            float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < np; i += 256) s += partials[i];
  s = block_sum<256>(s, red);
  if (threadIdx.x == 0) *out = s * inv_n;
}

constexpr int SL1_MAX_GRID = 8 * kNumSMs;

static inline int cv_log2_groups(int Cp) {   // wide tiles: 2, 4, 8 or 16 channel groups (Cp = 256: two tiles per slab)
  switch (Cp) {
    case 16: return 1;
    case 32: return 2;
    case 64: return 3;
    case 128: case 256: return 4;
    default: return 0;
  }
}

static inline int stream_grid(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb_slab_pool(const void* src, const int32_t* rowptr, const int32_t* colidx, const float* vals, void* dst, const void* ymul,
                  int B, int rows_out, int C, int act_mul, int zero_last, int planes, void* stream) {
  if (!src || !rowptr || !colidx || !vals || !dst || B <= 0 || rows_out <= 0 || C <= 0 || (C & 7)) return SHB_E_ARG;
  if (planes < 1 || planes > 2) return SHB_E_ARG;
  const int NB = slab::num_chunks(B), nvec = C * 16;
  // one warp per unit; a unit's vectors are cut into slices (>= 128 vectors: four per lane, an even count) until every
  // resident warp has a few units (few, long rows -- the per-slot sums of the dummy-row gradient -- are what needs the cut)
  int nslice = 1;
  const long long want = 8LL * kNumSMs * 24;
  while ((long long)rows_out * NB * nslice < want && nvec / (nslice * 2) >= 128 && (nvec / (nslice * 2)) % 64 == 0) nslice *= 2;
  if ((nvec / nslice) % 64) return SHB_E_ARG;   // cannot happen: nvec is a multiple of 128 and the loop keeps slices multiples of 64
  const long long units = (long long)rows_out * NB * nslice;
  if (units >= (1LL << 31)) return SHB_E_ARG;
  const long long blocks = (units + POOL_WARPS - 1) / POOL_WARPS;
  // as many blocks as are resident (launch bounds of the kernel): every warp walks its units with the requests of the next
  // one and two units in flight (measured against 8 blocks per SM: step 1.810 -> 1.799 ms)
  const long long resident = (long long)kNumSMs * (planes == 1 ? 3 : 2);
  const int grid = (int)(blocks < resident ? blocks : resident);
  const int threads = POOL_WARPS * 32;
  cudaStream_t st = (cudaStream_t)stream;
  const int act = ymul ? act_mul : SHB_ACT_IDENTITY;   // the kernel is compiled per activation: no switch inside the vector loop
  if (act < SHB_ACT_IDENTITY || act > SHB_ACT_TANH) return SHB_E_ARG;
#define SHB_POOL(PL, ACT)                                                                                                     \
  slab_pool_kernel<PL, ACT><<<grid, threads, 0, st>>>((const uint8_t*)src, rowptr, colidx, vals, (uint8_t*)dst,               \
                                                      (const uint8_t*)ymul, NB, rows_out, C, zero_last, nslice)
#define SHB_POOL_ACT(PL)                                           \
  switch (act) {                                                   \
    case SHB_ACT_RELU: SHB_POOL(PL, SHB_ACT_RELU); break;          \
    case SHB_ACT_ELU: SHB_POOL(PL, SHB_ACT_ELU); break;            \
    case SHB_ACT_LEAKY_RELU: SHB_POOL(PL, SHB_ACT_LEAKY_RELU); break; \
    case SHB_ACT_SIGMOID: SHB_POOL(PL, SHB_ACT_SIGMOID); break;    \
    case SHB_ACT_TANH: SHB_POOL(PL, SHB_ACT_TANH); break;          \
    default: SHB_POOL(PL, SHB_ACT_IDENTITY); break;                \
  }
  if (planes == 1) { SHB_POOL_ACT(1) } else { SHB_POOL_ACT(2) }
#undef SHB_POOL_ACT
#undef SHB_POOL
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_slab_from_rows(const void* src, int src_dtype, const int32_t* perm, const int32_t* perm_inv, void* dst, const void* ymul,
                       int B, int R, int Cs, int Cp, int act_mul, int zero_last, int planes, void* stream) {
  if (!src || !dst || B <= 0 || R <= 0 || Cs <= 0 || Cp < Cs || (Cp & 7) || planes < 1 || planes > 2) return SHB_E_ARG;
  if (src_dtype != SHB_F32 && src_dtype != SHB_BF16) return SHB_E_DTYPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int NB = slab::num_chunks(B);
  const int lg = Cs == Cp ? cv_log2_groups(Cp) : 0;
  if (Cp == 8 && (perm == nullptr || perm_inv != nullptr) && (long long)((R + CV_ROWS - 1) / CV_ROWS) * NB < (1LL << 31)) {
    // tiles in the caller's numbering
    const int tiles = (R + CV_ROWS - 1) / CV_ROWS * NB;
    const size_t smem = (size_t)CHUNK * ((CV_ROWS * Cs) | 1) * sizeof(float);
    const int per_sm = (int)(200 * 1024 / smem) < 8 ? (int)(200 * 1024 / smem) : 8;
    const int grid = tiles < kNumSMs * per_sm ? tiles : kNumSMs * per_sm;
#define SHB_FRN(T, PL)                                                                                                        \
  do {                                                                                                                        \
    if (smem > 48 * 1024) {                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(slab_from_rows_narrow_kernel<T, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           (int)smem);                                                                        \
      if (e != cudaSuccess) return (int)e;                                                                                    \
    }                                                                                                                         \
    slab_from_rows_narrow_kernel<T, PL><<<grid, CV_THREADS, smem, st>>>((const T*)src, perm_inv, (uint8_t*)dst,               \
                                                                        (const uint8_t*)ymul, B, R, Cs, act_mul, zero_last);  \
  } while (0)
    if (src_dtype == SHB_F32) { if (planes == 1) SHB_FRN(float, 1); else SHB_FRN(float, 2); }
    else { if (planes == 1) SHB_FRN(__nv_bfloat16, 1); else SHB_FRN(__nv_bfloat16, 2); }
#undef SHB_FRN
    SHB_LAUNCH_CHECK();
    return 0;
  }
  if (lg > 0 && (long long)R * NB * (Cp / (8 << lg)) < (1LL << 31)) {
    const int nsl = Cp / (8 << lg), tiles = R * NB * nsl;
    const size_t smem = (size_t)(1 << lg) * CV_GROUP_STRIDE;
    const int per_sm = (int)(200 * 1024 / smem) < 8 ? (int)(200 * 1024 / smem) : 8;
    const int grid = tiles < kNumSMs * per_sm ? tiles : kNumSMs * per_sm;
#define SHB_FRW(T, PL, LG)                                                                                                    \
  do {                                                                                                                        \
    if (smem > 48 * 1024) {                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(slab_from_rows_wide_kernel<T, PL, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize,\
                                           (int)smem);                                                                        \
      if (e != cudaSuccess) return (int)e;                                                                                    \
    }                                                                                                                         \
    slab_from_rows_wide_kernel<T, PL, LG><<<grid, CV_THREADS, smem, st>>>((const T*)src, perm, (uint8_t*)dst,                 \
                                                                          (const uint8_t*)ymul, B, R, Cp, act_mul, zero_last, \
                                                                          nsl);                                               \
  } while (0)
#define SHB_FRW_LG(T, PL)                   \
  switch (lg) {                             \
    case 1: SHB_FRW(T, PL, 1); break;       \
    case 2: SHB_FRW(T, PL, 2); break;       \
    case 3: SHB_FRW(T, PL, 3); break;       \
    default: SHB_FRW(T, PL, 4); break;      \
  }
    if (src_dtype == SHB_F32) { if (planes == 1) { SHB_FRW_LG(float, 1) } else { SHB_FRW_LG(float, 2) } }
    else { if (planes == 1) { SHB_FRW_LG(__nv_bfloat16, 1) } else { SHB_FRW_LG(__nv_bfloat16, 2) } }
#undef SHB_FRW_LG
#undef SHB_FRW
    SHB_LAUNCH_CHECK();
    return 0;
  }
  const long long total = (((long long)R * (Cp / 8) + 7) / 8) * 8 * CHUNK * NB;
  const int grid = stream_grid(total, 256);
#define SHB_FR(T, PL)                                                                                                          \
  slab_from_rows_kernel<T, PL><<<grid, 256, 0, st>>>((const T*)src, perm, (uint8_t*)dst, (const uint8_t*)ymul, B, R, Cs, Cp, \
                                                     act_mul, zero_last)
  if (src_dtype == SHB_F32) { if (planes == 1) SHB_FR(float, 1); else SHB_FR(float, 2); }
  else { if (planes == 1) SHB_FR(__nv_bfloat16, 1); else SHB_FR(__nv_bfloat16, 2); }
#undef SHB_FR
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_slab_to_rows(const void* src, const int32_t* perm, const int32_t* perm_inv, void* dst, int dst_dtype, int B, int R, int Cp,
                     int Cd, int planes, void* stream) {
  if (!src || !dst || B <= 0 || R <= 0 || Cd <= 0 || Cp < Cd || (Cp & 7) || planes < 1 || planes > 2) return SHB_E_ARG;
  if (dst_dtype != SHB_F32 && dst_dtype != SHB_BF16) return SHB_E_DTYPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int NB = slab::num_chunks(B);
  const int lg = Cd == Cp ? cv_log2_groups(Cp) : 0;
  if (Cp == 8 && (perm == nullptr || perm_inv != nullptr) && (long long)((R + CV_ROWS - 1) / CV_ROWS) * NB < (1LL << 31)) {
    const int tiles = (R + CV_ROWS - 1) / CV_ROWS * NB;
    const size_t smem = (size_t)CHUNK * ((CV_ROWS * Cd) | 1) * sizeof(float);
    const int per_sm = (int)(200 * 1024 / smem) < 8 ? (int)(200 * 1024 / smem) : 8;
    const int grid = tiles < kNumSMs * per_sm ? tiles : kNumSMs * per_sm;
#define SHB_TRN(T, PL)                                                                                                       \
  do {                                                                                                                       \
    if (smem > 48 * 1024) {                                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(slab_to_rows_narrow_kernel<T, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                           (int)smem);                                                                       \
      if (e != cudaSuccess) return (int)e;                                                                                   \
    }                                                                                                                        \
    slab_to_rows_narrow_kernel<T, PL><<<grid, CV_THREADS, smem, st>>>((const uint8_t*)src, perm_inv, (T*)dst, B, R, Cd);     \
  } while (0)
    if (dst_dtype == SHB_F32) { if (planes == 1) SHB_TRN(float, 1); else SHB_TRN(float, 2); }
    else { if (planes == 1) SHB_TRN(__nv_bfloat16, 1); else SHB_TRN(__nv_bfloat16, 2); }
#undef SHB_TRN
    SHB_LAUNCH_CHECK();
    return 0;
  }
  if (lg > 0 && (long long)R * NB * (Cp / (8 << lg)) < (1LL << 31)) {
    const int nsl = Cp / (8 << lg), tiles = R * NB * nsl;
    const size_t smem = (size_t)(1 << lg) * CV_GROUP_STRIDE;
    const int per_sm = (int)(200 * 1024 / smem) < 8 ? (int)(200 * 1024 / smem) : 8;
    const int grid = tiles < kNumSMs * per_sm ? tiles : kNumSMs * per_sm;
#define SHB_TRW(T, PL, LG)                                                                                                   \
  do {                                                                                                                       \
    if (smem > 48 * 1024) {                                                                                                  \
      cudaError_t e = cudaFuncSetAttribute(slab_to_rows_wide_kernel<T, PL, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem);                                                                       \
      if (e != cudaSuccess) return (int)e;                                                                                   \
    }                                                                                                                        \
    slab_to_rows_wide_kernel<T, PL, LG><<<grid, CV_THREADS, smem, st>>>((const uint8_t*)src, perm, (T*)dst, B, R, Cp, nsl);  \
  } while (0)
#define SHB_TRW_LG(T, PL)                   \
  switch (lg) {                             \
    case 1: SHB_TRW(T, PL, 1); break;       \
    case 2: SHB_TRW(T, PL, 2); break;       \
    case 3: SHB_TRW(T, PL, 3); break;       \
    default: SHB_TRW(T, PL, 4); break;      \
  }
    if (dst_dtype == SHB_F32) { if (planes == 1) { SHB_TRW_LG(float, 1) } else { SHB_TRW_LG(float, 2) } }
    else { if (planes == 1) { SHB_TRW_LG(__nv_bfloat16, 1) } else { SHB_TRW_LG(__nv_bfloat16, 2) } }
#undef SHB_TRW_LG
#undef SHB_TRW
    SHB_LAUNCH_CHECK();
    return 0;
  }
  const long long total = (((long long)R * ((Cd + 7) / 8) + 7) / 8) * 8 * CHUNK * NB;
  const int grid = stream_grid(total, 256);
#define SHB_TR(T, PL) slab_to_rows_kernel<T, PL><<<grid, 256, 0, st>>>((const uint8_t*)src, perm, (T*)dst, B, R, Cp, Cd)
  if (dst_dtype == SHB_F32) { if (planes == 1) SHB_TR(float, 1); else SHB_TR(float, 2); }
  else { if (planes == 1) SHB_TR(__nv_bfloat16, 1); else SHB_TR(__nv_bfloat16, 2); }
#undef SHB_TR
  SHB_LAUNCH_CHECK();
  return 0;
}

size_t shb_slab_l1_workspace(void) { return (size_t)SL1_MAX_GRID * sizeof(float); }

static int slab_l1_launch(int mode, const void* rec, const void* target, int target_dtype, const int32_t* perm_inv, float* partials,
                          const float* gscale, void* gdst, int B, int R, int Cs, int act, int zero_last, int planes, int* grid_out,
                          cudaStream_t st) {
  const int NB = slab::num_chunks(B);
  const long long tiles_ll = (long long)((R + CV_ROWS - 1) / CV_ROWS) * NB;
  if (tiles_ll >= (1LL << 31)) return SHB_E_SHAPE;
  const int tiles = (int)tiles_ll;
  const size_t smem = (size_t)CHUNK * ((CV_ROWS * Cs) | 1) * sizeof(float);
  const int per_sm = (int)(200 * 1024 / smem) < 8 ? (int)(200 * 1024 / smem) : 8;
  const int grid = tiles < kNumSMs * per_sm ? tiles : kNumSMs * per_sm;
  const float n_elems = (float)((long long)B * R * Cs);
#define SHB_SL1(T, PL, MODE)                                                                                                  \
  do {                                                                                                                        \
    if (smem > 48 * 1024) {                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(slab_l1_kernel<T, PL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return (int)e;                                                                                    \
    }                                                                                                                         \
    slab_l1_kernel<T, PL, MODE><<<grid, CV_THREADS, smem, st>>>((const uint8_t*)rec, (const T*)target, perm_inv, partials, gscale, \
                                                                (uint8_t*)gdst, B, R, Cs, n_elems, act, zero_last);           \
  } while (0)
#define SHB_SL1_M(T, PL) do { if (mode == 0) SHB_SL1(T, PL, 0); else SHB_SL1(T, PL, 1); } while (0)
  if (target_dtype == SHB_F32) { if (planes == 1) SHB_SL1_M(float, 1); else SHB_SL1_M(float, 2); }
  else { if (planes == 1) SHB_SL1_M(__nv_bfloat16, 1); else SHB_SL1_M(__nv_bfloat16, 2); }
#undef SHB_SL1_M
#undef SHB_SL1
  SHB_LAUNCH_CHECK();
  *grid_out = grid;
  return 0;
}

int shb_slab_l1_fwd(const void* rec, const void* target, int target_dtype, const int32_t* perm_inv, void* partials,
                    size_t partials_bytes, float* loss_out, int B, int R, int Cs, int planes, void* stream) {
  if (!rec || !target || !partials || !loss_out || B <= 0 || R <= 0 || Cs <= 0 || Cs > 8 || planes < 1 || planes > 2) return SHB_E_ARG;
  if (target_dtype != SHB_F32 && target_dtype != SHB_BF16) return SHB_E_DTYPE;
  if (partials_bytes < shb_slab_l1_workspace()) return SHB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = 0;
  const int rc = slab_l1_launch(0, rec, target, target_dtype, perm_inv, (float*)partials, nullptr, nullptr, B, R, Cs, 0, 0, planes,
                                &grid, st);
  if (rc != 0) return rc;
  slab_l1_final_kernel<<<1, 256, 0, st>>>((const float*)partials, grid, 1.0f / (float)((long long)B * R * Cs), loss_out);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_slab_l1_bwd(const void* rec, const void* target, int target_dtype, const int32_t* perm_inv, const float* gscale,
                    void* grad_slab, int B, int R, int Cs, int act_mul, int zero_last, int planes, void* stream) {
  if (!rec || !target || !gscale || !grad_slab || B <= 0 || R <= 0 || Cs <= 0 || Cs > 8 || planes < 1 || planes > 2)
    return SHB_E_ARG;
  if (target_dtype != SHB_F32 && target_dtype != SHB_BF16) return SHB_E_DTYPE;
  if (act_mul < SHB_ACT_IDENTITY || act_mul > SHB_ACT_TANH) return SHB_E_ARG;
  int grid = 0;
  return slab_l1_launch(1, rec, target, target_dtype, perm_inv, nullptr, gscale, grad_slab, B, R, Cs, act_mul, zero_last, planes,
                        &grid, (cudaStream_t)stream);
}

}  // extern "C"
