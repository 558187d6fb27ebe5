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
// A unit is one (row, chunk) slab (cut into nslice vector ranges when rows are few).  What bounds this kernel is the LENGTH OF
// THE DEPENDENT LOAD CHAIN per unit, not bandwidth (first version: one entry at a time, index load -> data load, per vector:
// 12 global-load latencies per unit, 30-40 % of HBM).  So: a unit's entries are fetched four at a time (indices and weights of
// all four in flight together), the data loads of all four entries are issued back to back, and the next unit's row range is
// requested before the current unit is processed: three latencies per unit (row range -> entries -> data).  One vector per
// thread (two doubled the registers, halved the resident warps and measured 10 % slower).  The kernel is also instruction-
// bound (IPC 0.5 at half occupancy, ~150 instructions per 16-byte vector with four predicated entries): the entry count of a
// row is uniform over the block, so the accumulation is specialised on it (no predicated-off work).  CSR order is kept.
// N entries (1..4) of one row into acc: the N index / weight loads, then the N data loads, all unconditional and back to back.
// FIRST: acc is written (first chunk of a row), not accumulated into.
template <int P, int N, bool FIRST>
__device__ __forceinline__ void pool_accum(const uint8_t* __restrict__ sv, size_t rstride, size_t plane_b,
                                           const int32_t* __restrict__ colidx, const float* __restrict__ vals, int kb, float* acc) {
  int col[N];
  float w[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    col[j] = __ldg(colidx + kb + j);
    w[j] = __ldg(vals + kb + j);
  }
  uint4 raw[N][P];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const uint8_t* s = sv + (size_t)col[j] * rstride;
    raw[j][0] = __ldg(reinterpret_cast<const uint4*>(s));
    if (P == 2) raw[j][P - 1] = __ldg(reinterpret_cast<const uint4*>(s + plane_b));
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float x[8];
    unpack8(raw[j][0], x);
    if (P == 2) {
      float l[8];
      unpack8(raw[j][P - 1], l);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] += l[i];
    }
    if (FIRST && j == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = w[0] * x[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w[j], x[i], acc[i]);
    }
  }
}

template <int P>
__global__ void __launch_bounds__(256) slab_pool_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                        uint8_t* __restrict__ dst, const uint8_t* __restrict__ ymul, int NB,
                                                        int rows_out, int C, int act_mul, int zero_last, int nslice) {
  const int nvec = C * 16;                         // 16-byte vectors per plane of a slab
  const int vper = (nvec + nslice - 1) / nslice;   // vectors per slice
  const size_t slab_b = slab_bytes(C, P);
  const size_t plane_b = (size_t)C * 256;
  const size_t rstride = (size_t)NB * slab_b;
  const int units = rows_out * NB * nslice;
  int unit = blockIdx.x;
  int k0 = 0, k1 = 0;
  if (unit < units) {
    const int r = unit / (nslice * NB);
    k0 = __ldg(rowptr + r);
    k1 = __ldg(rowptr + r + 1);
  }
  while (unit < units) {
    const int rq = unit / nslice, sl = unit - rq * nslice;
    const int r = rq / NB, q = rq - r * NB;
    const int unit_n = unit + gridDim.x;
    int n0 = 0, n1 = 0;
    if (unit_n < units) {  // next unit's row range: in flight while this unit is processed
      const int rn = unit_n / (nslice * NB);
      n0 = __ldg(rowptr + rn);
      n1 = __ldg(rowptr + rn + 1);
    }
    const bool zero = zero_last && r == rows_out - 1;
    const size_t uoff = ((size_t)r * NB + q) * slab_b;
    const int v1 = (sl + 1) * vper < nvec ? (sl + 1) * vper : nvec;
    const int n = zero ? 0 : k1 - k0;   // the whole block works on one row: the branches below are uniform
    for (int v = sl * vper + threadIdx.x; v < v1; v += blockDim.x) {
      float acc[8];
      uint4 yraw[P];
      if (ymul != nullptr && n > 0) {  // act' operand: requested first, consumed last
        yraw[0] = __ldg(reinterpret_cast<const uint4*>(ymul + uoff + (size_t)v * 16));
        if (P == 2) yraw[P - 1] = __ldg(reinterpret_cast<const uint4*>(ymul + uoff + plane_b + (size_t)v * 16));
      }
      const uint8_t* sv = src + (size_t)q * slab_b + (size_t)v * 16;
      // first (usually only) chunk of up to four entries, specialised on its length; longer rows continue four at a time
      switch (n < 4 ? n : 4) {
        case 0:
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
          break;
        case 1: pool_accum<P, 1, true>(sv, rstride, plane_b, colidx, vals, k0, acc); break;
        case 2: pool_accum<P, 2, true>(sv, rstride, plane_b, colidx, vals, k0, acc); break;
        case 3: pool_accum<P, 3, true>(sv, rstride, plane_b, colidx, vals, k0, acc); break;
        default: pool_accum<P, 4, true>(sv, rstride, plane_b, colidx, vals, k0, acc); break;
      }
      int kb = k0 + 4;
      for (; kb + 4 <= k1 && n > 4; kb += 4) pool_accum<P, 4, false>(sv, rstride, plane_b, colidx, vals, kb, acc);
      if (n > 4) {
        switch (k1 - kb) {
          case 1: pool_accum<P, 1, false>(sv, rstride, plane_b, colidx, vals, kb, acc); break;
          case 2: pool_accum<P, 2, false>(sv, rstride, plane_b, colidx, vals, kb, acc); break;
          case 3: pool_accum<P, 3, false>(sv, rstride, plane_b, colidx, vals, kb, acc); break;
          default: break;
        }
      }
      if (ymul != nullptr && n > 0) {
        float yy[8];
        unpack8(yraw[0], yy);
        if (P == 2) {
          float l[8];
          unpack8(yraw[P - 1], l);
#pragma unroll
          for (int i = 0; i < 8; ++i) yy[i] += l[i];
        }
        act_bwd8(acc, yy, act_mul);
      }
      uint8_t* d = dst + uoff + (size_t)v * 16;
      if (P == 1) {
        *reinterpret_cast<uint4*>(d) = pack8(acc);
      } else {
        uint4 hi, lo;
        split8(acc, hi, lo);
        *reinterpret_cast<uint4*>(d) = hi;
        *reinterpret_cast<uint4*>(d + plane_b) = lo;
      }
    }
    unit = unit_n; k0 = n0; k1 = n1;
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
  int nslice = 1;  // few, long rows (the per-slot sums of the dummy-row gradient): spread a unit's vectors over several CTAs
  while (rows_out * NB * nslice < 4 * kNumSMs && nvec / (nslice * 2) >= 32) nslice *= 2;
  const int units = rows_out * NB * nslice;
  const int per = nvec / nslice;
  const int threads = per >= 256 ? 256 : (per >= 128 ? 128 : (per >= 64 ? 64 : 32));
  int grid = units < kNumSMs * 8 ? units : kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (planes == 1)
    slab_pool_kernel<1><<<grid, threads, 0, st>>>((const uint8_t*)src, rowptr, colidx, vals, (uint8_t*)dst, (const uint8_t*)ymul,
                                                  NB, rows_out, C, act_mul, zero_last, nslice);
  else
    slab_pool_kernel<2><<<grid, threads, 0, st>>>((const uint8_t*)src, rowptr, colidx, vals, (uint8_t*)dst, (const uint8_t*)ymul,
                                                  NB, rows_out, C, act_mul, zero_last, nslice);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_slab_from_rows(const void* src, int src_dtype, const int32_t* perm, void* dst, const void* ymul, int B, int R, int Cs,
                       int Cp, int act_mul, int zero_last, int planes, void* stream) {
  if (!src || !dst || B <= 0 || R <= 0 || Cs <= 0 || Cp < Cs || (Cp & 7) || planes < 1 || planes > 2) return SHB_E_ARG;
  if (src_dtype != SHB_F32 && src_dtype != SHB_BF16) return SHB_E_DTYPE;
  const long long total = (((long long)R * (Cp / 8) + 7) / 8) * 8 * CHUNK * slab::num_chunks(B);
  const int grid = stream_grid(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
#define SHB_FR(T, PL)                                                                                                          \
  slab_from_rows_kernel<T, PL><<<grid, 256, 0, st>>>((const T*)src, perm, (uint8_t*)dst, (const uint8_t*)ymul, B, R, Cs, Cp, \
                                                     act_mul, zero_last)
  if (src_dtype == SHB_F32) { if (planes == 1) SHB_FR(float, 1); else SHB_FR(float, 2); }
  else { if (planes == 1) SHB_FR(__nv_bfloat16, 1); else SHB_FR(__nv_bfloat16, 2); }
#undef SHB_FR
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_slab_to_rows(const void* src, const int32_t* perm, void* dst, int dst_dtype, int B, int R, int Cp, int Cd, int planes,
                     void* stream) {
  if (!src || !dst || B <= 0 || R <= 0 || Cd <= 0 || Cp < Cd || (Cp & 7) || planes < 1 || planes > 2) return SHB_E_ARG;
  if (dst_dtype != SHB_F32 && dst_dtype != SHB_BF16) return SHB_E_DTYPE;
  const long long total = (((long long)R * ((Cd + 7) / 8) + 7) / 8) * 8 * CHUNK * slab::num_chunks(B);
  const int grid = stream_grid(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
#define SHB_TR(T, PL) slab_to_rows_kernel<T, PL><<<grid, 256, 0, st>>>((const uint8_t*)src, perm, (T*)dst, B, R, Cp, Cd)
  if (dst_dtype == SHB_F32) { if (planes == 1) SHB_TR(float, 1); else SHB_TR(float, 2); }
  else { if (planes == 1) SHB_TR(__nv_bfloat16, 1); else SHB_TR(__nv_bfloat16, 2); }
#undef SHB_TR
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
