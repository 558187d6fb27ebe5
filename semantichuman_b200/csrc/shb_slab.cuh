// Slab layout + the TMA bulk-copy pieces shared by the slab kernels (shb_slab_conv.cu, shb_slab_wgrad.cu,
// shb_slab_misc.cu).
//
// Inside the model trunks an activation tensor of `rows` vertices (dummy row included), batch B and C channels is stored
// BATCH-INNERMOST, in 128-sample chunks, channel-chunk planar ("slab layout"):
//
//     element (r, b, c), plane p  ->  bf16 index  ((((r*NB + b/128)*P + p)*(C/8) + c/8)*128 + b%128)*8 + c%8
//
//     NB = ceil(B/128) batch chunks (the tail chunk is zero-padded in memory), C a multiple of 8,
//     P = 1 (bf16 mode) or 2 (fp32 mode: plane 0 = bf16(x), plane 1 = bf16(x - plane 0); x ~ hi + lo to 2^-17).
//
// One (vertex, batch-chunk) pair -- a SLAB -- is P*C*256 contiguous bytes.  A slab lands in shared memory with ONE
// cp.async.bulk (TMA, no tensor map), and the landed bytes ARE the un-swizzled UMMA canonical layout, for both uses:
//   K-major  A operand (forward / input gradient: M = 128 samples, K = channels):  core matrix = 8 samples x 16 B,
//            SBO (next 8 samples) = 128 B, LBO (next 8 channels) = 2048 B;
//   MN-major A/B operand (weight gradient: K = 128 samples, M/N = channels):        LBO (next 8 samples) = 128 B,
//            SBO (next 8 channels) = 2048 B -- and consecutive slabs in shared memory continue the same 2048-byte stride, so
//            an M = 128 operand spans 128/C neighbour slabs with one descriptor.
// (Conventions verified on a B200 by tests/cuda/slab_probe.cu; log in profiles/r02_a_slab_probe.log.)
// What SpiralConv gathers per output vertex is therefore S contiguous slabs: no per-row gather, no LSU traffic, no index
// staging; the epilogue's stores are 512 contiguous bytes per warp.
#pragma once
#include <stdint.h>

#include "shb_common.cuh"
#include "shb_umma.cuh"

namespace shb {
namespace slab {

constexpr int CHUNK = 128;            // samples per slab == UMMA M (K-major use) == UMMA K x 8 (MN-major use)
constexpr int PLANE_STRIDE = 2048;    // bytes between consecutive 8-channel planes of a slab: 128 samples x 16 B

__host__ __device__ inline int num_chunks(int B) { return (B + CHUNK - 1) / CHUNK; }
__host__ __device__ inline size_t slab_bytes(int C, int planes) { return (size_t)planes * C * 256; }
__host__ __device__ inline size_t tensor_bytes(int rows, int B, int C, int planes) {
  return (size_t)rows * num_chunks(B) * slab_bytes(C, planes);
}

// ---------------------------------------------------------------------------------------------- TMA bulk copy, mbarrier tx
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared, completion (bytes) signalled on `bar`; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(ns);
  }
}
// Wait with a hardware suspend-time hint: the thread is parked (no polling instructions competing for the shared-memory
// pipe with the producer / MMA threads) until the phase completes or `hint_ns` elapse.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
  } while (ok == 0);
}
// One lane of a converged warp (elect.sync): unlike `lane == 0`, code under this predicate keeps warp-uniform values in the
// uniform datapath, which is what the single-thread tcgen05 / bulk-copy instructions take their operands from.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ void mma_commit_u32(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// fp32 -> (hi, lo) bf16 pair: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8(const uint4& u, float* o) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
// 8 fp32 values -> hi and lo packed bf16 vectors
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  float h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(v[i]);
    h[i] = __bfloat162float(hb);
    l[i] = v[i] - h[i];
  }
  hi = pack8(h);
  lo = pack8(l);
}

}  // namespace slab
}  // namespace shb
