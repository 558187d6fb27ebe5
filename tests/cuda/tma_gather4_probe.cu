// Stand-alone probe (test infrastructure, next-round groundwork): what does a TMA row gather cost on B200?
//
// The SpiralConv kernels gather ≤128-byte rows with cp.async (LDGSTS): measured 4-5 cycles per gathered 64-byte row per SM,
// bound by L1TEX wavefronts and instruction issue.  The alternative is
//     cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes
// one instruction = four rows of a 2-D tensor (row ids in registers) written by the async proxy, no LSU data-pipe traffic.
// Its tensor map is an ordinary tiled 2-D map with box = (columns, 1 row); a transaction moves 4 boxes.
//
// The probe fills a (rows x C) bf16 tensor, lets W warps per CTA (one CTA per SM) issue gather4 instructions with random
// row ids into a ring of shared-memory buffers, and reports cycles per instruction / per row / achieved GB/s for box widths
// of 32..256 bytes, plus a data check of the last buffer (no swizzle, so rows land verbatim).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/tma_gather4_probe tests/cuda/tma_gather4_probe.cu
//   ./tests/cuda/tma_gather4_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}

constexpr int RING = 4;  // buffers in flight per issuing warp

struct ProbeOut { long long cycles; unsigned mismatches; };

// Each issuing warp owns RING buffers of 32 lanes x 4 rows x box_bytes and its own barriers: per round, lane 0 posts the
// expected bytes, every lane issues one gather4, and the warp waits for the round RING-1 rounds back.
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ row_ids,
                                                    const __nv_bfloat16* __restrict__ src, int C, int box_cols, int rounds,
                                                    int n_ids, ProbeOut* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[4][RING];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const uint32_t box_bytes = (uint32_t)box_cols * 2, buf_bytes = 32 * 4 * box_bytes;
  uint8_t* my = smem + (size_t)warp * RING * buf_bytes;
  if (lane == 0)
    for (int i = 0; i < RING; ++i) mbar_init(smem_u32(&bars[warp][i]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int* ids = row_ids + (((size_t)blockIdx.x * nwarps + warp) * 32 + lane) * 4;
  const size_t id_stride = (size_t)gridDim.x * nwarps * 32 * 4;
  const long long t0 = clock64();
  int last_round_ids[4] = {0, 0, 0, 0};
  for (int r = 0; r < rounds; ++r) {
    const int slot = r % RING;
    const uint32_t bar = smem_u32(&bars[warp][slot]);
    if (r >= RING) mbar_wait(bar, ((r / RING) - 1) & 1);  // buffer free again: its previous transaction has landed
    __syncwarp();
    if (lane == 0) mbar_expect_tx(bar, buf_bytes);
    __syncwarp();
    const int* q = ids + (size_t)(r % (n_ids / (int)id_stride > 0 ? n_ids / (int)id_stride : 1)) * id_stride;
    const int a = q[0], b = q[1], c = q[2], d = q[3];
    tma_gather4(smem_u32(my + (size_t)slot * buf_bytes + (size_t)lane * 4 * box_bytes), &map, bar, 0, a, b, c, d);
    if (r == rounds - 1) { last_round_ids[0] = a; last_round_ids[1] = b; last_round_ids[2] = c; last_round_ids[3] = d; }
  }
  for (int r = rounds > RING ? rounds - RING : 0; r < rounds; ++r) mbar_wait(smem_u32(&bars[warp][r % RING]), (r / RING) & 1);
  const long long t1 = clock64();
  // data check of the last round (no swizzle: four rows of box_bytes land back to back)
  unsigned bad = 0;
  const uint8_t* got = my + (size_t)((rounds - 1) % RING) * buf_bytes + (size_t)lane * 4 * box_bytes;
  for (int k = 0; k < 4; ++k) {
    const uint8_t* want = reinterpret_cast<const uint8_t*>(src + (size_t)last_round_ids[k] * C);
    for (uint32_t i = 0; i < box_bytes; ++i) bad += got[k * box_bytes + i] != want[i];
  }
  for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
  if (lane == 0) atomicAdd(&out[blockIdx.x].mismatches, bad);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, %.0f MHz\n", prop.name, sms, prop.clockRate / 1e3);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled unavailable\n"); return 1; }
  EncodeTiled encode = (EncodeTiled)fn;

  const int rows = 256 * 6891, rounds = 2000;
  for (int C : {16, 32, 64, 128}) {  // row widths of the model's layers: 32..256 bytes
    std::vector<__nv_bfloat16> h((size_t)rows * C);
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)((i * 2654435761u) % 1021) / 64.f);
    __nv_bfloat16* d_src;
    CK(cudaMalloc(&d_src, h.size() * 2));
    CK(cudaMemcpy(d_src, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {(cuuint32_t)C, 1};  // gather4: one row per box, four boxes per instruction
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("C=%d: cuTensorMapEncodeTiled failed (%d)\n", C, (int)cr); continue; }
    for (int warps : {1, 2, 4}) {
      // locality like the model's: ids drawn from a 2000-row window of a random sample
      const size_t per_round = (size_t)sms * warps * 32 * 4;
      const int id_rounds = 64;
      std::vector<int> ids(per_round * id_rounds);
      uint64_t s = 88172645463325252ull;
      for (size_t i = 0; i < ids.size(); ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const int sample = (int)((i / (32 * 4)) % 256);
        ids[i] = sample * 6891 + (int)(s % 2000);
      }
      int* d_ids;
      CK(cudaMalloc(&d_ids, ids.size() * 4));
      CK(cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
      ProbeOut* d_out;
      CK(cudaMalloc(&d_out, sizeof(ProbeOut) * sms));
      CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * sms));
      const size_t smem = (size_t)warps * RING * 32 * 4 * C * 2;
      CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      probe_kernel<<<sms, warps * 32, smem>>>(map, d_ids, d_src, C, C, 50, (int)ids.size(), d_out);  // warm-up
      CK(cudaMemset(d_out, 0, sizeof(ProbeOut) * sms));
      CK(cudaEventRecord(e0));
      probe_kernel<<<sms, warps * 32, smem>>>(map, d_ids, d_src, C, C, rounds, (int)ids.size(), d_out);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<ProbeOut> o(sms);
      CK(cudaMemcpy(o.data(), d_out, sizeof(ProbeOut) * sms, cudaMemcpyDeviceToHost));
      long long cyc = 0; unsigned bad = 0;
      for (auto& x : o) { cyc = x.cycles > cyc ? x.cycles : cyc; bad += x.mismatches; }
      const double instr = (double)rounds * warps * 32, rows_sm = instr * 4;
      printf("row %3d B, %d warp(s)/SM: %7.1f cycles per gather4 per SM, %5.2f cycles per row, %7.1f GB/s chip, mismatching bytes %u\n",
             C * 2, warps, cyc / instr, cyc / rows_sm, rows_sm * sms * C * 2 / (ms * 1e6), bad);
      CK(cudaFree(d_ids)); CK(cudaFree(d_out));
    }
    CK(cudaFree(d_src));
  }
  printf("reference point: the cp.async gather of the forward kernel runs at ~4.6 cycles per 64-byte row per SM\n");
  return 0;
}
