// Stand-alone probe (test infrastructure): pins down, on a real B200, the conventions the slab kernels
// (csrc/shb_slab_*.cu) rely on, and measures the TMA slab-streaming rate that bounds them.
//
// Part 1 -- descriptor semantics.  A tiny interpreter kernel executes host-built lists of TMA loads and tcgen05.mma
// instructions (one CTA), dumps the TMEM accumulator, and the host compares with a CPU product.  Cases:
//   K-major A operand written by a 3-D tiled TMA load with SWIZZLE_32B/64B/128B (rows of 32/64/128/256 bytes), and
//   16-byte rows without swizzle where one K=16 MMA spans two slabs through LBO;
//   MN-major A and B operands (the weight-gradient product, K = batch) for the same row widths, with both LBO/SBO
//   assignments tried so that the output says which one the hardware uses.
// Part 2 -- streaming rate.  148 persistent CTAs load 128-sample slabs of neighbouring rows of a (R, B, C) tensor
// through a ring of shared-memory stages (no MMA) and report the achieved L2->SM fill rate.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/slab_probe tests/cuda/slab_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Load { int map, c0, c1, c2; uint32_t off, bytes; };
struct Mma { uint64_t da, db; uint32_t idesc, col, acc, pad; };
constexpr int MAX_LOADS = 24, MAX_MMAS = 40;
struct Prog {
  Load loads[MAX_LOADS];
  Mma mmas[MAX_MMAS];
  int n_loads, n_mmas, ncols;
  uint32_t img_off, img_bytes;
};

__global__ void __launch_bounds__(128) interp_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                                                     const Prog* __restrict__ prog, const uint8_t* __restrict__ img,
                                                     float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = smem_u32(smem);
  for (uint32_t i = tid * 16; i < prog->img_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(smem + prog->img_off + i) = *reinterpret_cast<const uint4*>(img + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    uint32_t total = 0;
    for (int i = 0; i < prog->n_loads; ++i) total += prog->loads[i].bytes;
    mbar_expect_tx(smem_u32(&bars[0]), total);
    for (int i = 0; i < prog->n_loads; ++i) {
      const Load l = prog->loads[i];
      tma_load_3d(base + l.off, l.map == 0 ? &m0 : &m1, smem_u32(&bars[0]), l.c0, l.c1, l.c2);
    }
    mbar_wait(smem_u32(&bars[0]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t add = (uint64_t)((base & 0x3FFFFu) >> 4);
    for (int i = 0; i < prog->n_mmas; ++i) {
      const Mma m = prog->mmas[i];
      mma_bf16(tmem + m.col, m.da + add, m.db + add, m.idesc, m.acc);
    }
    mma_commit(smem_u32(&bars[1]));
    mbar_wait(smem_u32(&bars[1]), 0);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < prog->ncols; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int t = 0; t < 8; ++t) out[(size_t)(warp * 32 + lane) * prog->ncols + c + t] = __uint_as_float(v[t]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------ host helpers
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode;

static CUtensorMapSwizzle swz_for_bytes(int bytes) {
  return bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
         : bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
}
// (R, B, C) bf16 tensor; box = (cbox, 128, 1)
static CUtensorMap make_map(void* d, int R, int B, int C, int cbox) {
  CUtensorMap m;
  const cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)B, (cuuint64_t)R};
  const cuuint64_t gstr[2] = {(cuuint64_t)C * 2, (cuuint64_t)B * C * 2};
  const cuuint32_t box[3] = {(cuuint32_t)cbox, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult cr = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz_for_bytes(cbox * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed (%d) R=%d B=%d C=%d cbox=%d\n", (int)cr, R, B, C, cbox); exit(1); }
  return m;
}
static uint32_t layout_type_for_bytes(int bytes) { return bytes >= 128 ? 2u : bytes == 64 ? 4u : bytes == 32 ? 6u : 0u; }
static uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
static uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
static float bf(const __nv_bfloat16& x) { return __bfloat162float(x); }
static void fill(std::vector<__nv_bfloat16>& v, uint64_t seed) {
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
  for (auto& x : v) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = __float2bfloat16((float)((int)(s % 17) - 8) / 8.f); }
}

struct Ctx {
  Prog* d_prog; uint8_t* d_img; float* d_out;
};
static double run_and_compare(const Ctx& c, const CUtensorMap& m0, const CUtensorMap& m1, const Prog& p,
                              const std::vector<uint8_t>& img, const std::vector<float>& want, int rows, int ncols) {
  CK(cudaMemcpy(c.d_prog, &p, sizeof(Prog), cudaMemcpyHostToDevice));
  if (!img.empty()) CK(cudaMemcpy(c.d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(c.d_out, 0xFF, 128 * 512 * 4));
  interp_kernel<<<1, 128, 200 * 1024>>>(m0, m1, c.d_prog, c.d_img, c.d_out);
  CK(cudaDeviceSynchronize());
  std::vector<float> got((size_t)128 * ncols);
  CK(cudaMemcpy(got.data(), c.d_out, got.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int r = 0; r < rows; ++r)
    for (int n = 0; n < ncols; ++n) {
      double d = fabs((double)got[(size_t)r * ncols + n] - (double)want[(size_t)r * ncols + n]);
      if (!(d == d)) d = 1e30;
      if (d > worst) worst = d;
    }
  return worst;
}

// ---------------------------------------------------------------------------------------------- part 2: streaming
__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap map, int R, int nchunk, int S, int nboxes,
                                                    int cbox, int nst, uint32_t box_bytes, const int* __restrict__ offs,
                                                    long long* cycles, int mode, const uint8_t* __restrict__ base_ptr) {
  // mode bit 0: 1 = one contiguous bulk copy per slab (chunk-planar layout), 0 = tensor-map boxes
  // mode bit 1: 1 = tiles dealt round-robin to the CTAs (all CTAs work in the same region), 0 = contiguous ranges
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[32], empty[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < nst; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tiles = R * nchunk;
  const int per = (tiles + gridDim.x - 1) / gridDim.x;
  const bool rr = (mode & 2) != 0;
  const int t0 = rr ? blockIdx.x : blockIdx.x * per, t1 = rr ? tiles : min(tiles, blockIdx.x * per + per), tstep = rr ? gridDim.x : 1;
  const uint32_t stage_bytes = box_bytes * nboxes;
  const long long c0 = clock64();
  if (tid == 0) {
    uint32_t slot = 0, ph = 0;
    for (int t = t0; t < t1; t += tstep) {
      const int chunk = rr ? (t % nchunk) : t / R, j = rr ? t / nchunk : t - chunk * R;
      for (int s = 0; s < S; ++s) {
        int r = j + ((mode & 4) ? ((j * 7 + s * 13) % 161) - 80 : offs[(j * S + s) & 4095]);
        r = r < 0 ? 0 : (r >= R ? R - 1 : r);
        if (mode & 8) mbar_wait(smem_u32(&full[slot]), ph ^ 1);   // self-consume: the previous load into this slot has landed
        else mbar_wait(smem_u32(&empty[slot]), ph ^ 1);
        mbar_expect_tx(smem_u32(&full[slot]), stage_bytes);
        if (mode & 1)
          bulk_load(smem_u32(smem) + slot * stage_bytes, base_ptr + ((size_t)r * nchunk + chunk) * stage_bytes, stage_bytes, smem_u32(&full[slot]));
        else
          for (int b = 0; b < nboxes; ++b)
            tma_load_3d(smem_u32(smem) + slot * stage_bytes + b * box_bytes, &map, smem_u32(&full[slot]), b * cbox, chunk * 128, r);
        if (++slot == (uint32_t)nst) { slot = 0; ph ^= 1; }
      }
    }
  } else if (tid == 32 && !(mode & 8)) {
    uint32_t slot = 0, ph = 0;
    for (int t = t0; t < t1; t += tstep)
      for (int s = 0; s < S; ++s) {
        mbar_wait(smem_u32(&full[slot]), ph);
        mbar_arrive(smem_u32(&empty[slot]));
        if (++slot == (uint32_t)nst) { slot = 0; ph ^= 1; }
      }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - c0;
}


// ---------------------------------------------------------------------------------------------- part 3: issue / completion timeline
// One thread issues n bulk copies back to back (no waits); a second thread polls their barriers in order.
__global__ void __launch_bounds__(64) timeline_kernel(const uint8_t* __restrict__ base_ptr, uint32_t bytes, int n, size_t stride,
                                                      int nissuers, long long* stamps /* [grid][2][32] */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < n; ++i) mbar_init(smem_u32(&full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long* my = stamps + (size_t)blockIdx.x * 64;
  const long long c0 = clock64();
  if (tid < nissuers) {
    for (int i = tid; i < n; i += nissuers) {
      mbar_expect_tx(smem_u32(&full[i]), bytes);
      bulk_load(smem_u32(smem) + i * bytes, base_ptr + ((size_t)blockIdx.x * 97 + i * 5) * stride, bytes, smem_u32(&full[i]));
      my[i] = clock64() - c0;
    }
  } else if (tid == 32) {
    for (int i = 0; i < n; ++i) {
      mbar_wait(smem_u32(&full[i]), 0);
      my[32 + i] = clock64() - c0;
    }
  }
}


// ---------------------------------------------------------------------------------------------- part 4: steady-state timeline
// One thread issues n bulk copies through a ring of nst slots, waiting for the previous load of a slot before reusing it.
// Stamps: [i][0] before the wait, [i][1] after the wait, [i][2] after the issue.
__global__ void __launch_bounds__(64) ring_timeline_kernel(const uint8_t* __restrict__ base_ptr, uint32_t bytes, int n, int nst,
                                                           size_t stride, long long* stamps /* [grid][n][3] */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < nst; ++i) mbar_init(smem_u32(&full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long* my = stamps + (size_t)blockIdx.x * n * 3;
  if (tid == 0) {
    const long long c0 = clock64();
    uint32_t slot = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
      const long long a = clock64();
      mbar_wait(smem_u32(&full[slot]), ph ^ 1);
      const long long b = clock64();
      mbar_expect_tx(smem_u32(&full[slot]), bytes);
      bulk_load(smem_u32(smem) + slot * bytes, base_ptr + ((size_t)blockIdx.x * 97 + (size_t)i * 5) % 30000 * stride, bytes, smem_u32(&full[slot]));
      const long long c = clock64();
      my[i * 3 + 0] = a - c0; my[i * 3 + 1] = b - c0; my[i * 3 + 2] = c - c0;
      if (++slot == (uint32_t)nst) { slot = 0; ph ^= 1; }
    }
  }
}


// ---------------------------------------------------------------------------------------------- part 5: MMA issue/execute rate
// One thread issues `n` tcgen05.mma (M=128, K=16) with the given descriptors (operand contents irrelevant), alternating
// between `nacc` accumulators, then commits and waits.  Reports cycles per MMA.
__global__ void __launch_bounds__(128) mma_rate_kernel(uint64_t da, uint64_t db, uint32_t idesc, int n, int nacc, int ncols,
                                                       uint32_t a_step, int a_wrap, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint64_t add = (uint64_t)((smem_u32(smem) & 0x3FFFFu) >> 4);
    const long long c0 = clock64();
    for (int i = 0; i < n; ++i)
      mma_bf16(tmem + (uint32_t)((i % nacc) * ncols), da + add + (uint64_t)((i % a_wrap) * (a_step >> 4)), db + add, idesc, i >= nacc ? 1u : 0u);
    const long long c1 = clock64();
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long c2 = clock64();
    out[0] = c1 - c0; out[1] = c2 - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}


// part 5b: same measurement with a fully unrolled issue sequence (no per-MMA arithmetic at all)
__global__ void __launch_bounds__(128) mma_rate_unrolled_kernel(uint64_t da, uint64_t db, uint32_t idesc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint64_t add = (uint64_t)((smem_u32(smem) & 0x3FFFFu) >> 4);
    const uint64_t a = da + add, b = db + add;
    const long long c0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) mma_bf16(tmem, a, b, idesc, 1u);
    const long long c1 = clock64();
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long c2 = clock64();
    out[0] = c1 - c0; out[1] = c2 - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled unavailable\n"); return 1; }
  g_encode = (EncodeTiled)fn;
  CK(cudaFuncSetAttribute(interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  Ctx ctx;
  CK(cudaMalloc(&ctx.d_prog, sizeof(Prog)));
  CK(cudaMalloc(&ctx.d_img, 128 * 1024));
  CK(cudaMalloc(&ctx.d_out, 128 * 512 * 4));
  const int R = 6, B = 160;
  const uint32_t IMG = 128 * 1024;  // smem offset of the second operand
  int failures = 0;

  // ---------------------------------------------------------------- K-major A (forward / input-gradient product)
  for (int C : {8, 16, 32, 64, 128}) {
    for (int chunk = 0; chunk < 2; ++chunk) {  // chunk 1: samples 128..159 valid, the rest zero-filled by the TMA unit
      const int N = 32, cbox = C > 64 ? 64 : C, nbox = C / cbox, rowb = cbox * 2;
      std::vector<__nv_bfloat16> X((size_t)R * B * C);
      fill(X, 100 + C);
      __nv_bfloat16* dX;
      CK(cudaMalloc(&dX, X.size() * 2));
      CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
      const CUtensorMap map = make_map(dX, R, B, C, cbox);
      const int nslab = C == 8 ? 2 : 1;           // 16-byte rows: one K=16 MMA spans two slabs
      const int rows_used[2] = {4, 1};
      const int K = C * nslab;
      std::vector<__nv_bfloat16> W((size_t)N * K);
      fill(W, 200 + C);
      // un-swizzled K-major image of W: core matrix (n>>3, q) at ((n>>3)*Q + q)*128, row n&7 at +16*(n&7)
      const int Q = K / 8;
      std::vector<uint8_t> img((size_t)N * K * 2);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
          memcpy(&img[(((size_t)(n >> 3) * Q + (k >> 3)) * 128) + (n & 7) * 16 + (k & 7) * 2], &W[(size_t)n * K + k], 2);
      Prog p;
      memset(&p, 0, sizeof(p));
      p.img_off = IMG; p.img_bytes = (uint32_t)img.size(); p.ncols = N;
      const uint32_t slab_bytes = 128u * rowb;
      for (int sl = 0; sl < nslab; ++sl)
        for (int b = 0; b < nbox; ++b)
          p.loads[p.n_loads++] = Load{0, b * cbox, chunk * 128, rows_used[sl], (uint32_t)(sl * nbox + b) * slab_bytes, slab_bytes};
      const uint32_t idesc = make_idesc(128, N, 0, 0);
      for (int kk = 0; kk < K / 16; ++kk) {
        uint64_t da;
        if (C == 8) da = make_desc(0, slab_bytes, 128, 0);                       // LBO = next K core matrix = next slab
        else da = make_desc((uint32_t)(kk / (cbox / 16)) * slab_bytes + (kk % (cbox / 16)) * 32, 0, 8 * rowb, layout_type_for_bytes(rowb));
        const uint64_t db = make_desc(IMG + kk * 256, 128, Q * 128, 0);
        p.mmas[p.n_mmas++] = Mma{da, db, idesc, 0, kk > 0 ? 1u : 0u, 0};
      }
      std::vector<float> want((size_t)128 * N, 0.f);
      for (int b = 0; b < 128; ++b)
        for (int n = 0; n < N; ++n) {
          float acc = 0;
          const int bb = chunk * 128 + b;
          if (bb < B)
            for (int sl = 0; sl < nslab; ++sl)
              for (int c = 0; c < C; ++c) acc += bf(X[((size_t)rows_used[sl] * B + bb) * C + c]) * bf(W[(size_t)n * K + sl * C + c]);
          want[(size_t)b * N + n] = acc;
        }
      const double err = run_and_compare(ctx, map, map, p, img, want, 128, N);
      printf("K-major  A rows of %3d B (%s), chunk %d: max err %.3g  %s\n", C * 2, C == 8 ? "no swizzle, LBO across slabs" : "TMA swizzle",
             chunk, err, err < 1e-3 ? "OK" : "MISMATCH");
      failures += err >= 1e-3;
      CK(cudaFree(dX));
    }
  }

  // ---------------------------------------------------------------- MN-major A and B (weight-gradient product, K = batch)
  for (int C : {8, 16, 32, 64, 128}) {
    for (int Co : {8, 16, 32, 64, 128}) {
      if (!((C == Co) || (C == 32 && Co == 16) || (C == 128 && Co == 64) || (C == 64 && Co == 128) || (C == 16 && Co == 8) || (C == 8 && Co == 16)))
        continue;
      for (int variant = 0; variant < 2; ++variant) {
        const int cboxA = C > 64 ? 64 : C, rowA = cboxA * 2, nslabA = 128 / C > 0 ? 128 / C : 1, nboxA = C / cboxA;
        const int cboxB = Co > 64 ? 64 : Co, rowB = cboxB * 2, nboxB = Co / cboxB;
        std::vector<__nv_bfloat16> X((size_t)16 * B * C), G((size_t)R * B * Co);
        fill(X, 300 + C);
        fill(G, 400 + Co);
        __nv_bfloat16 *dX, *dG;
        CK(cudaMalloc(&dX, X.size() * 2)); CK(cudaMalloc(&dG, G.size() * 2));
        CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dG, G.data(), G.size() * 2, cudaMemcpyHostToDevice));
        const CUtensorMap mA = make_map(dX, 16, B, C, cboxA), mB = make_map(dG, R, B, Co, cboxB);
        Prog p;
        memset(&p, 0, sizeof(p));
        p.ncols = Co;
        const uint32_t slabA = 128u * rowA, slabB = 128u * rowB;
        int rowsA[16];
        for (int i = 0; i < nslabA; ++i) rowsA[i] = (i * 7 + 3) % 16;
        for (int i = 0; i < nslabA; ++i)
          for (int b = 0; b < nboxA; ++b)
            p.loads[p.n_loads++] = Load{0, b * cboxA, 0, rowsA[i], (uint32_t)(i * nboxA + b) * slabA, slabA};
        for (int b = 0; b < nboxB; ++b) p.loads[p.n_loads++] = Load{1, b * cboxB, 0, 2, IMG + b * slabB, slabB};
        const uint32_t idesc = make_idesc(128, Co < 16 ? 16 : Co, 1, 1);
        for (int kk = 0; kk < 8; ++kk) {  // 16 samples per MMA
          uint64_t da, db;
          if (C == 8) da = make_desc(kk * 16 * rowA, variant ? slabA : 128, variant ? 128 : slabA, 0);
          else da = make_desc(kk * 16 * rowA, variant ? 8 * rowA : slabA, variant ? slabA : 8 * rowA, layout_type_for_bytes(rowA));
          if (Co == 8) db = make_desc(IMG + kk * 16 * rowB, variant ? slabB : 128, variant ? 128 : slabB, 0);
          else db = make_desc(IMG + kk * 16 * rowB, variant ? 8 * rowB : slabB, variant ? slabB : 8 * rowB, layout_type_for_bytes(rowB));
          p.mmas[p.n_mmas++] = Mma{da, db, idesc, 0, kk > 0 ? 1u : 0u, 0};
        }
        std::vector<float> want((size_t)128 * Co, 0.f);
        for (int m = 0; m < 128; ++m) {
          const int slab = C >= 128 ? 0 : m / C, c = C >= 128 ? m : m % C;
          for (int n = 0; n < Co; ++n) {
            float acc = 0;
            for (int b = 0; b < 128; ++b) acc += bf(X[((size_t)rowsA[slab] * B + b) * C + c]) * bf(G[((size_t)2 * B + b) * Co + n]);
            want[(size_t)m * Co + n] = acc;
          }
        }
        const double err = run_and_compare(ctx, mA, mB, p, std::vector<uint8_t>(), want, 128, Co);
        printf("MN-major A %3d B rows x B %3d B rows, %s: max err %.3g  %s\n", C * 2, Co * 2,
               variant ? "LBO=8-sample group, SBO=MN group" : "LBO=MN group, SBO=8-sample group (CUTLASS comment)", err,
               err < 1e-3 ? "OK" : "mismatch");
        CK(cudaFree(dX)); CK(cudaFree(dG));
      }
    }
  }

  // ---------------------------------------------------------------- streaming rate
  {
    const int Rr = 6891, Bb = 256;
    std::vector<int> offs(4096);
    uint64_t s = 12345;
    for (auto& o : offs) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; o = (int)(s % 161) - 80; }
    int* d_offs;
    CK(cudaMalloc(&d_offs, offs.size() * 4));
    CK(cudaMemcpy(d_offs, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice));
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, 148 * 8));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    for (int C : {8, 16, 32, 64, 128}) {
      const int cbox = C > 64 ? 64 : C, nbox = C / cbox;
      __nv_bfloat16* dX;
      const size_t n = (size_t)Rr * Bb * C;
      CK(cudaMalloc(&dX, n * 2));
      CK(cudaMemset(dX, 0, n * 2));
      const CUtensorMap map = make_map(dX, Rr, Bb, C, cbox);
      const uint32_t box_bytes = 128u * cbox * 2, stage = box_bytes * nbox;
      for (int mode : {3, 7, 15})
      for (int nst_bytes : {32 * 1024, 128 * 1024, 200 * 1024}) {
        const int nst = nst_bytes / (int)stage > 32 ? 32 : nst_bytes / (int)stage;
        if (nst < 2) continue;
        if (C == 8 && !(mode & 1)) continue;
        const int S = 14;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        stream_kernel<<<148, 64, nst * stage>>>(map, Rr, 2, S, nbox, cbox, nst, box_bytes, d_offs, d_cyc, mode, (const uint8_t*)dX);
        CK(cudaEventRecord(e0));
        stream_kernel<<<148, 64, nst * stage>>>(map, Rr, 2, S, nbox, cbox, nst, box_bytes, d_offs, d_cyc, mode, (const uint8_t*)dX);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)Rr * 2 * S * stage;
        printf("stream[%s,%s]: rows of %3d B, %2d stages of %5u B in flight per SM: %.3f ms, %.0f GB/s L2->SM fill (%.1f B/clk/SM at 1.965 GHz), tensor %.0f MB\n",
               (mode & 1) ? "bulk" : "tmap", (mode & 8) ? "rr,arith,self-consume" : (mode & 4) ? "rr,arith" : "rr,table", C * 2, nst, stage, ms, bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.965, n * 2 / 1e6);
      }
      CK(cudaFree(dX));
    }
  }

  // ---------------------------------------------------------------- issue / completion timeline of bulk copies
  {
    uint8_t* d;
    const size_t total = (size_t)512 << 20;
    CK(cudaMalloc(&d, total));
    CK(cudaMemset(d, 0, total));
    long long* d_st;
    CK(cudaMalloc(&d_st, 148 * 64 * 8));
    CK(cudaFuncSetAttribute(timeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    for (int grid : {1, 148})
      for (uint32_t bytes : {2048u, 8192u})
        for (int nissuers : {1, 4}) {
          const int n = 24;
          for (int rep = 0; rep < 2; ++rep) {
            timeline_kernel<<<grid, 64, n * bytes>>>(d, bytes, n, 16384, nissuers, d_st);
            CK(cudaDeviceSynchronize());
          }
          std::vector<long long> st(64);
          CK(cudaMemcpy(st.data(), d_st, 64 * 8, cudaMemcpyDeviceToHost));
          printf("timeline grid=%d bytes=%u issuers=%d (warm L2): issue", grid, bytes, nissuers);
          for (int i = 0; i < n; i += 1) printf(" %lld", st[i]);
          printf(" | done");
          for (int i = 0; i < n; i += 1) printf(" %lld", st[32 + i]);
          printf("\n");
        }
  }

  // ---------------------------------------------------------------- steady-state ring timeline
  {
    uint8_t* d;
    const size_t total = (size_t)512 << 20;
    CK(cudaMalloc(&d, total));
    CK(cudaMemset(d, 0, total));
    const int n = 120;
    long long* d_st;
    CK(cudaMalloc(&d_st, 148 * n * 3 * 8));
    CK(cudaFuncSetAttribute(ring_timeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    for (int grid : {1, 148})
      for (uint32_t bytes : {2048u, 8192u})
        for (int nst : {4, 16}) {
          for (int rep = 0; rep < 2; ++rep) {
            ring_timeline_kernel<<<grid, 64, nst * bytes>>>(d, bytes, n, nst, 16384, d_st);
            CK(cudaDeviceSynchronize());
          }
          std::vector<long long> st(n * 3);
          CK(cudaMemcpy(st.data(), d_st, n * 3 * 8, cudaMemcpyDeviceToHost));
          printf("ring grid=%d bytes=%u nst=%d: (before-wait, after-wait, after-issue) ops 0..3:", grid, bytes, nst);
          for (int i = 0; i < 4; ++i) printf(" (%lld %lld %lld)", st[i * 3], st[i * 3 + 1], st[i * 3 + 2]);
          printf(" ... ops 100..107:");
          for (int i = 100; i < 108; ++i) printf(" (%lld %lld %lld)", st[i * 3], st[i * 3 + 1], st[i * 3 + 2]);
          printf("  => %.0f cycles/op over ops 40..119\n", (double)(st[119 * 3 + 2] - st[40 * 3 + 2]) / 79.0);
        }
  }

  // ---------------------------------------------------------------- MMA rate per operand layout
  {
    long long* d_out;
    CK(cudaMalloc(&d_out, 16));
    CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct L { const char* name; uint32_t layout, lbo, sbo, step; };
    // K-major A operand, 128 rows: un-swizzled slab planes (LBO 2048, SBO 128); SW32/64/128 row-major tiles
    const L la[] = {{"A none(LBO2048,SBO128)", 0, 2048, 128, 4096}, {"A SW32 ", 6, 0, 256, 4096}, {"A SW64 ", 4, 0, 512, 32},
                    {"A SW128", 2, 0, 1024, 32}};
    for (const L& a : la)
      for (int bsw = 0; bsw < 2; ++bsw)
        for (int N : {16, 32, 64, 128, 256})
          for (int nacc : {1, 2}) {
            if (nacc * N > 512) continue;
            const int n = 256;
            const uint64_t da = make_desc(0, a.lbo, a.sbo, a.layout);
            // B: N rows x K=16: un-swizzled image (LBO 128, SBO 256) or SW128 rows of 128 B
            const uint64_t db = bsw ? make_desc(IMG, 0, 1024, 2) : make_desc(IMG, 128, 256, 0);
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
              mma_rate_kernel<<<1, 128, 200 * 1024>>>(da, db, make_idesc(128, N, 0, 0), n, nacc, N, a.step, a.step == 32 ? 2 : 8, d_out);
              CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
            printf("mma rate: %s, B %s, N=%3d, %d accumulator(s): issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.1f)\n", a.name,
                   bsw ? "SW128" : "none ", N, nacc, (double)h[0] / n, (double)h[1] / n, 128.0 * N * 16 * 2 / 8192.0);
          }
  }

  {
    long long* d_out;
    CK(cudaMalloc(&d_out, 16));
    CK(cudaFuncSetAttribute(mma_rate_unrolled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int M : {128, 64})
      for (int sw = 0; sw < 2; ++sw)
        for (int N : {16, 32, 64, 128, 256}) {
          const uint64_t da = sw ? make_desc(0, 0, 1024, 2) : make_desc(0, 2048, 128, 0);
          const uint64_t db = sw ? make_desc(IMG, 0, 1024, 2) : make_desc(IMG, 128, 256, 0);
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) {
            mma_rate_unrolled_kernel<<<1, 128, 200 * 1024>>>(da, db, make_idesc(M, N, 0, 0), d_out);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
          printf("mma unrolled: M=%d %s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.1f)\n", M, sw ? "SW128" : "none ", N,
                 (double)h[0] / 64, (double)h[1] / 64, (double)M * N * 16 * 2 / 8192.0);
        }
  }
  printf("descriptor cases failed (K-major): %d\n", failures);
  return 0;
}
