// Stand-alone probe of tcgen05.mma operand-descriptor semantics on sm_100a (test infrastructure).
//
// The host builds the exact shared-memory byte image of A (128 x K) and B (N x K), plus the descriptor fields
// (leading/stride byte offsets, per-K-step start-address advance, major-ness), for several layout hypotheses.
// The kernel copies the images into smem, issues K/16 MMAs (bf16 x bf16 -> fp32 in TMEM), reads the accumulator
// back with tcgen05.ld and the host compares with a CPU GEMM.  Pass/fail per hypothesis is printed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/umma_probe tests/cuda/umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct ProbeParams {
  const uint8_t* a_img; const uint8_t* b_img; int a_bytes, b_bytes;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // bytes
  uint32_t kstep_a, kstep_b;             // bytes added to the start address per K=16 step
  uint32_t idesc; int ksteps; int N;
  float* d_out;                          // 128 x N
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;               // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128) probe_kernel(ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 32768;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t * 16; i < p.a_bytes; i += 128 * 16) *(uint4*)(sa + i) = *(const uint4*)(p.a_img + i);
  for (int i = t * 16; i < p.b_bytes; i += 128 * 16) *(uint4*)(sb + i) = *(const uint4*)(p.b_img + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (t == 0) {
    for (int k = 0; k < p.ksteps; ++k) {
      const uint64_t da = make_desc(smem_u32(sa) + k * p.kstep_a, p.lbo_a, p.sbo_a);
      const uint64_t db = make_desc(smem_u32(sb) + k * p.kstep_b, p.lbo_b, p.sbo_b);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(tmem_base), "l"(da), "l"(db), "r"(p.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32w..32w+31; thread lane i <-> row 32w+i; 16 columns per load
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) p.d_out[(size_t)t * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128));
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

// element (r, k) of an (R x K) operand -> byte offset in the image, for the two un-swizzled canonical layouts
//  K-major : core matrix = 8 rows x 16 B (8 consecutive k);   image [r/8][k/8][r%8][k%8]
//  MN-major: core matrix = 8 k   x 16 B (8 consecutive rows); image [k/8][r/8][k%8][r%8]
static size_t off_kmajor(int r, int k, int R, int K) { (void)R; return ((size_t)(r / 8) * (K / 8) + k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2; }
static size_t off_mnmajor(int r, int k, int R, int K) { (void)K; return ((size_t)(k / 8) * (R / 8) + r / 8) * 128 + (k % 8) * 16 + (r % 8) * 2; }

struct Hyp { const char* name; int a_mn, b_mn; int swap_a, swap_b; };

int main() {
  const int M = 128, K = 64;
  int fails = 0;
  const Hyp hyps[] = {
      {"K-major A,B: LBO=K-dir core stride, SBO=MN-dir core stride", 0, 0, 0, 0},
      {"K-major A,B: LBO/SBO swapped", 0, 0, 1, 1},
      {"MN-major A,B: LBO=K-dir, SBO=MN-dir", 1, 1, 0, 0},
      {"MN-major A,B: LBO/SBO swapped", 1, 1, 1, 1},
      {"MN-major A, K-major B", 1, 0, 0, 0},
      {"K-major A, MN-major B", 0, 1, 0, 0},
  };
  for (int N : {16, 32, 128}) {
    std::vector<float> A(M * K), B(N * K), D(M * N, 0.f);
    srand(7 + N);
    for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 500.0f));
    for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 500.0f));
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; D[m * N + n] = (float)s; }
    for (const Hyp& h : hyps) {
      std::vector<uint8_t> ai(M * K * 2, 0), bi(N * K * 2, 0);
      for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { uint16_t v = f2bf(A[m * K + k]); memcpy(&ai[h.a_mn ? off_mnmajor(m, k, M, K) : off_kmajor(m, k, M, K)], &v, 2); }
      for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { uint16_t v = f2bf(B[n * K + k]); memcpy(&bi[h.b_mn ? off_mnmajor(n, k, N, K) : off_kmajor(n, k, N, K)], &v, 2); }
      ProbeParams p{};
      // core-matrix strides of the images above
      uint32_t a_kdir = h.a_mn ? 128u * (M / 8) : 128u, a_mndir = h.a_mn ? 128u : 128u * (K / 8);
      uint32_t b_kdir = h.b_mn ? 128u * (N / 8) : 128u, b_mndir = h.b_mn ? 128u : 128u * (K / 8);
      p.lbo_a = h.swap_a ? a_mndir : a_kdir; p.sbo_a = h.swap_a ? a_kdir : a_mndir;
      p.lbo_b = h.swap_b ? b_mndir : b_kdir; p.sbo_b = h.swap_b ? b_kdir : b_mndir;
      p.kstep_a = 2 * a_kdir; p.kstep_b = 2 * b_kdir;  // K=16 per MMA = two 8-wide core matrices along K
      p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)h.a_mn << 15) | ((uint32_t)h.b_mn << 16) |
                ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      p.ksteps = K / 16; p.N = N; p.a_bytes = (int)ai.size(); p.b_bytes = (int)bi.size();
      uint8_t *da, *db; float* dd;
      CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dd, M * N * 4));
      CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
      CK(cudaMemset(dd, 0xFF, M * N * 4));
      p.a_img = da; p.b_img = db; p.d_out = dd;
      CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      probe_kernel<<<1, 128, 65536>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%3d  %-62s  LAUNCH ERROR %s\n", N, h.name, cudaGetErrorString(e)); return 2; }
      std::vector<float> out(M * N);
      CK(cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; int bad = 0;
      for (int i = 0; i < M * N; ++i) { double d = fabs((double)out[i] - D[i]); if (!(d <= 1e-2)) ++bad; if (d > maxerr || d != d) maxerr = d; }
      printf("N=%3d  %-62s  %s  maxerr %.3g  bad %d/%d\n", N, h.name, bad ? "FAIL" : "PASS", maxerr, bad, M * N);
      if (bad && !h.swap_a) ++fails;
      cudaFree(da); cudaFree(db); cudaFree(dd);
    }
  }
  printf("un-swapped hypotheses failing: %d\n", fails);
  return 0;
}
