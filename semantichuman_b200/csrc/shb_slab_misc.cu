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
template <int P>
__global__ void __launch_bounds__(256) slab_pool_kernel(const uint8_t* __restrict__ src, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                                        uint8_t* __restrict__ dst, const uint8_t* __restrict__ ymul, int NB,
                                                        int rows_out, int C, int act_mul, int zero_last, int nslice) {
  const int nvec = C * 16;                  // 16-byte vectors per plane of a slab
  const int vper = (nvec + nslice - 1) / nslice;   // a (row, chunk) unit is cut into nslice vector ranges when rows are few
  const size_t slab_b = slab_bytes(C, P);
  const int units = rows_out * NB * nslice;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int rq = unit / nslice, sl = unit - rq * nslice;
    const int r = rq / NB, q = rq - r * NB;
    const int k0 = __ldg(rowptr + r), k1 = __ldg(rowptr + r + 1);
    const bool zero = zero_last && r == rows_out - 1;
    uint8_t* d = dst + ((size_t)r * NB + q) * slab_b;
    const uint8_t* y = ymul != nullptr ? ymul + ((size_t)r * NB + q) * slab_b : nullptr;
    const int v1 = (sl + 1) * vper < nvec ? (sl + 1) * vper : nvec;
    for (int v = sl * vper + threadIdx.x; v < v1; v += blockDim.x) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      if (!zero) {
        const uint8_t* sv = src + (size_t)q * slab_b + (size_t)v * 16;
        const size_t rstride = (size_t)NB * slab_b;
        int k = k0;
        for (; k + 4 <= k1; k += 4) {  // four independent loads in flight per thread; accumulation stays in CSR order
          uint4 raw[4][P];
          float w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint8_t* s = sv + (size_t)__ldg(colidx + k + j) * rstride;
            w[j] = __ldg(vals + k + j);
            raw[j][0] = __ldg(reinterpret_cast<const uint4*>(s));
            if (P == 2) raw[j][P - 1] = __ldg(reinterpret_cast<const uint4*>(s + (size_t)C * 256));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x[8];
            unpack8(raw[j][0], x);
            if (P == 2) {
              float l[8];
              unpack8(raw[j][P - 1], l);
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] += l[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(w[j], x[i], acc[i]);
          }
        }
        for (; k < k1; ++k) {
          const float w = __ldg(vals + k);
          const uint8_t* s = sv + (size_t)__ldg(colidx + k) * rstride;
          float x[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(s)), x);
          if (P == 2) {
            float l[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(s + (size_t)C * 256)), l);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] += l[i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, x[i], acc[i]);
        }
        if (y != nullptr) {
          float yy[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(y + (size_t)v * 16)), yy);
          if (P == 2) {
            float l[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(y + (size_t)C * 256 + (size_t)v * 16)), l);
#pragma unroll
            for (int i = 0; i < 8; ++i) yy[i] += l[i];
          }
          act_bwd8(acc, yy, act_mul);
        }
      }
      if (P == 1) {
        *reinterpret_cast<uint4*>(d + (size_t)v * 16) = pack8(acc);
      } else {
        uint4 hi, lo;
        split8(acc, hi, lo);
        *reinterpret_cast<uint4*>(d + (size_t)v * 16) = hi;
        *reinterpret_cast<uint4*>(d + (size_t)C * 256 + (size_t)v * 16) = lo;
      }
    }
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
  const long long total = (long long)R * NB * ncc * CHUNK;
  const size_t slab_b = slab_bytes(Cp, P);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int bl = (int)(i % CHUNK);
    long long t = i / CHUNK;
    const int cc = (int)(t % ncc);
    t /= ncc;
    const int q = (int)(t % NB), r = (int)(t / NB);
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
  const long long total = (long long)R * NB * ncc * CHUNK;
  const size_t slab_b = slab_bytes(Cp, P);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int bl = (int)(i % CHUNK);
    long long t = i / CHUNK;
    const int cc = (int)(t % ncc);
    t /= ncc;
    const int q = (int)(t % NB), r = (int)(t / NB);
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
  const long long total = (long long)R * slab::num_chunks(B) * (Cp / 8) * CHUNK;
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
  const long long total = (long long)R * slab::num_chunks(B) * ((Cd + 7) / 8) * CHUNK;
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
