// SpiralConv forward / input-gradient on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 operands, fp32 accumulate.
//
//   C[(b,j), n] = epilogue( sum_k A[(b,j), k] * Bop[n, k] )        M = B*rows_dst rows, K = S*CS, N = Cd (padded to 16)
//
//   forward : A[(b,j), s*CS+c] = x[b, table[j,s], c]                      Bop[n,k] = W[n, k]
//   dgrad   : A[(b,u), s*CS+c] = sum_{j in inv(u,s)} gz[b, j, c]          Bop[n, s*CS+c] = W[c, s*Cd + n]   (CS = Cout, Cd = Cin)
//
// Persistent warp-specialised CTAs (one or two per SM), 128-row tiles:
//   * the whole weight operand lives in shared memory for the lifetime of the CTA (un-swizzled K-major core-matrix layout);
//   * 8 (forward) or 16 (dgrad) producer warps gather 16-byte channel chunks of neighbour rows straight into the K-major
//     core-matrix layout of a multi-stage A ring -- cp.async (LDGSTS) for the forward gather, a register gather-sum
//     (packed bf16 adds in a fixed order, no atomics) for dgrad; the tile's index block is staged per warp in shared memory;
//   * one thread issues tcgen05.mma (M=128, N=Cd, K=16 per instruction) into a double-buffered TMEM accumulator and
//     releases A stages / publishes accumulators with tcgen05.commit -> mbarrier;
//   * 4 epilogue warps read the accumulator with tcgen05.ld, apply bias + activation + dummy-row mask, and store bf16.
// The gathered (M x K) matrix never exists in HBM: activations are read once from HBM (re-reads hit L2).
#include <stdlib.h>

#include "shb_common.cuh"
#include "shb_internal.h"
#include "shb_umma.cuh"

namespace shb {

using namespace umma;

constexpr int UG_BM = 128;             // rows per tile == UMMA M
constexpr int UG_KC = 8;               // 16-byte chunks (8 bf16) per row per stage  -> BK = 64
// K-direction core-matrix stride of the A ring: 128 B of payload + 16 B of padding.  LDGSTS writes shared memory one
// returning 32-byte SECTOR at a time; a sector holds two consecutive K chunks of a row, which an un-padded layout puts
// 128 B apart (same banks, 2-way conflict on every sector: ncu showed 35 smem wavefronts per LDGSTS).  The UMMA
// descriptor accepts any multiple of 16 as LBO/SBO.
constexpr int UG_LBO = 144;
constexpr int UG_SBO = UG_KC * UG_LBO;   // 8-row group stride
constexpr int UG_STAGE_BYTES = (UG_BM / 8) * UG_SBO;
constexpr int UG_EPI_WARPS = 4, UG_PROD_WARPS = 8;
constexpr int UG_PROD_THREADS = UG_PROD_WARPS * 32;
constexpr int UG_THREADS = (UG_EPI_WARPS + 1 + UG_PROD_WARPS) * 32;
// the gather-SUM (input-gradient) variant is bound by load latency, not by issue: it runs 16 producer warps (2 row-items
// per thread instead of 4), one CTA per SM
constexpr int ug_prod_warps(bool sum) { return sum ? 16 : 8; }
constexpr int ug_threads(bool sum) { return (UG_EPI_WARPS + 1 + ug_prod_warps(sum)) * 32; }
constexpr int UG_MAX_STAGES = 8;
constexpr int UG_MIN_STAGES = 3;       // shallowest A ring the kernels are launched with

struct UGParams {
  const __nv_bfloat16* src;  // (B, rows_src, CS)
  const int32_t* table;      // forward: (rows_dst, S) source rows.  dgrad: quads (rows_dst, S, 4 x uint16) viewed as 2 words per
                             //   key: first four source rows inline; 0xFFFF = none; [3] == 0xFFFE = 5+ entries -> CSR below
  const int32_t* keyptr;     // dgrad: (rows_dst*S + 1) CSR of the (u,s)-keyed inverse relation (overflow path only)
  const int32_t* list;       // dgrad: concatenated source rows per key
  const __nv_bfloat16* w;    // nn.Linear weight (Cout, S*Cin), bf16
  const float* bias;         // (Cd) fp32 or null
  __nv_bfloat16* dst;        // (B, rows_dst, Cd)
  long long M;
  int rows_src, rows_dst, S, Cd, NPAD;
  int Q;        // K / 8: 16-byte chunks per operand row
  int NS;       // stages per tile = ceil(Q / UG_KC)
  int nstage;   // depth of the A ring
  int num_tiles;
  int act, zero_last;
  int skip_last;  // gather-sum: destination dummy row not computed.  forward: source dummy row is known zero (zero-fill)
  long long* trace;  // debug timeline (CTA 0): [role 0..2][event 0..3][512] clock64 stamps, or null
  int dbg;      // SHB_UMMA_DEBUG bitmask (perf bisection only): 1 skip copies, 2 skip MMAs, 4 skip epilogue stores
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

#ifdef SHB_UMMA_TRACE
#define UG_TRACE(role, ev, idx)                                                              \
  do {                                                                                     \
    if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 512) p.trace[((role) * 4 + (ev)) * 512 + (idx)] = clock64(); \
  } while (0)
#else
#define UG_TRACE(role, ev, idx) do { } while (0)
#endif

// Every role's inner loop is latency-bound on a single warp's dependent instruction stream (measured with the
// clock64 timeline: ~6 cycles per instruction), so the loops below carry NO runtime divisions, keep ring slot/phase
// as incremented counters, and advance precomputed UMMA descriptors by adding to their low word.
template <int CS, bool SUM>
__global__ void __launch_bounds__(ug_threads(SUM), SUM ? 1 : 2) umma_gather_gemm_kernel(const UGParams p) {
  constexpr int NPW = ug_prod_warps(SUM);   // producer warps
  constexpr int NPT = NPW * 32;             // producer threads
  constexpr int NTH = ug_threads(SUM);      // CTA threads
  constexpr int NIT = UG_BM * UG_KC / NPT;  // (row, chunk) items per producer thread per stage: 4 or 2
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[UG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UG_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* w_img = dyn_smem;                                   // [NPAD/8][Q][8 rows][16 B]
  const uint32_t w_bytes = (uint32_t)p.NPAD * p.Q * 16;
  uint8_t* a_ring = dyn_smem + ((w_bytes + 127) / 128) * 128;  // nstage x [16 row groups][8 chunks][8 rows][16 B]
  const int K = p.Q * 8;
  const uint32_t nstage = (uint32_t)p.nstage;

  // ---------------------------------------------------------------- prologue: weights -> smem (core-matrix layout)
  {
    const int total_chunks = p.NPAD * p.Q;
    if (!SUM) {
      for (int i = tid; i < total_chunks; i += NTH) {
        const int n = i / p.Q, q = i - n * p.Q;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n < p.Cd) v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)n * K + (size_t)q * 8));
        *reinterpret_cast<uint4*>(w_img + ((size_t)(n >> 3) * p.Q + q) * 128 + (n & 7) * 16) = v;
      }
    } else if ((p.Cd & 7) == 0) {
      // Bop[n][s*CS + co] = W[co*(S*Cd) + s*Cd + n]: the operand is W transposed per slot.  A thread takes an 8 (co) x 8 (n)
      // block: eight 16-byte loads along n (consecutive lanes -> consecutive n blocks: coalesced), an in-register transpose,
      // eight 16-byte stores (one per n, contiguous 128 B).  The scalar version below issued 8 uncoalesced 2-byte loads per
      // chunk, ~100 us of prologue per CTA on the 128-channel layers.
      const int NB = p.NPAD >> 3, blocks = NB * p.Q;
      const size_t rs = (size_t)p.S * p.Cd;
      for (int i = tid; i < blocks; i += NTH) {
        const int q = i / NB, nb = i - q * NB, n0 = nb * 8;
        const int k0 = q * 8, sl = k0 / CS, co0 = k0 - sl * CS;
        uint4 r[8];
#pragma unroll
        for (int t = 0; t < 8; ++t)
          r[t] = n0 < p.Cd ? __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(co0 + t) * rs + (size_t)sl * p.Cd + n0))
                           : make_uint4(0, 0, 0, 0);
        uint8_t* dst = w_img + ((size_t)nb * p.Q + q) * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
          uint32_t wv[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&r[t]);
            wv[t] = rw[j >> 1];
          }
          *reinterpret_cast<uint4*>(dst + j * 16) =
              make_uint4(__byte_perm(wv[0], wv[1], sel), __byte_perm(wv[2], wv[3], sel), __byte_perm(wv[4], wv[5], sel),
                         __byte_perm(wv[6], wv[7], sel));
        }
      }
    } else {
      for (int i = tid; i < total_chunks; i += NTH) {
        const int q = i / p.NPAD, n = i - q * p.NPAD;  // n fastest: the 2-byte loads of a warp are contiguous
        uint4 v = make_uint4(0, 0, 0, 0);
        if (n < p.Cd) {
          const int k0 = q * 8, sl = k0 / CS, co0 = k0 - sl * CS;
          const unsigned short* wu = reinterpret_cast<const unsigned short*>(p.w);
          const size_t rs = (size_t)p.S * p.Cd;
          uint32_t e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) e[t] = __ldg(wu + (size_t)(co0 + t) * rs + (size_t)sl * p.Cd + n);
          v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        }
        *reinterpret_cast<uint4*>(w_img + ((size_t)(n >> 3) * p.Q + q) * 128 + (n & 7) * 16) = v;
      }
    }
    if (tid < 256) bias_s[tid] = (p.bias != nullptr && tid < p.Cd) ? __ldg(p.bias + tid) : 0.f;
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full_bar[i], SUM ? NPW : NPT);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], UG_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == UG_EPI_WARPS) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < UG_EPI_WARPS) {
    // ================================================================ epilogue warps: TMEM -> bias/act/mask -> bf16 -> HBM
    int tcount = 0;
    const unsigned rows_dst = (unsigned)p.rows_dst;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      mbar_wait_backoff(&tfull_bar[buf], (tcount >> 1) & 1, 100);
      tc_fence_after();
      const unsigned m = (unsigned)tile * UG_BM + warp * 32 + lane;
      const bool valid = (long long)m < p.M;
      const bool zero = p.zero_last && valid && (m % rows_dst == rows_dst - 1);
      __nv_bfloat16* out = p.dst + (size_t)(valid ? m : 0) * p.Cd;
      const uint32_t taddr = tmem_base + (uint32_t)(buf * p.NPAD) + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < p.NPAD; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (c0 + 16 >= p.NPAD) {  // accumulator fully read: hand the buffer back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        if (!valid || (p.dbg & 4)) continue;
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = zero ? 0.f : act_fwd(__uint_as_float(r[t]) + bias_s[c0 + t], p.act);
        if ((p.Cd & 7) == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (c0 + 8 * h < p.Cd) {
              const uint4 o = make_uint4(pack_bf16x2(v[8 * h], v[8 * h + 1]), pack_bf16x2(v[8 * h + 2], v[8 * h + 3]),
                                         pack_bf16x2(v[8 * h + 4], v[8 * h + 5]), pack_bf16x2(v[8 * h + 6], v[8 * h + 7]));
              *reinterpret_cast<uint4*>(out + c0 + 8 * h) = o;
            }
          }
        } else {
#pragma unroll
          for (int t = 0; t < 16; ++t)
            if (c0 + t < p.Cd) out[c0 + t] = __float2bfloat16_rn(v[t]);
        }
      }
    }
  } else if (warp == UG_EPI_WARPS) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(UG_BM, p.NPAD, 0, 0);
      // descriptor = hi (SBO, version) | lo (start address >> 4, LBO = 128 B); k-steps and stages only move the start address
      const uint64_t hi_a = ((uint64_t)(UG_SBO >> 4) << 32) | ((uint64_t)1 << 46);
      const uint64_t hi_b = ((uint64_t)(((uint32_t)p.Q * 128) >> 4) << 32) | ((uint64_t)1 << 46);
      const uint32_t lo_a0 = (smem_u32(a_ring) >> 4) | ((uint32_t)(UG_LBO >> 4) << 16);
      const uint32_t lo_b0 = (smem_u32(w_img) >> 4) | ((128u >> 4) << 16);
      uint32_t slot = 0, ph = 0, lo_a = lo_a0;
      int tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait_backoff(&tempty_bar[buf], ((tcount >> 1) & 1) ^ 1, 40);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.NPAD);
        uint32_t lo_b = lo_b0, acc = 0;
        int left = p.Q;  // chunks of this tile still to multiply
        for (int st = 0; st < p.NS; ++st) {
          UG_TRACE(1, 0, st);
          mbar_wait_backoff(&full_bar[slot], ph, 20);
          UG_TRACE(1, 1, st);
          // producers fill the stage through the generic proxy (cp.async / st.shared); tcgen05.mma reads it through the
          // async proxy: the cross-proxy fence sits here, after the acquire, on the one consuming thread
          fence_proxy_async_smem();
          tc_fence_after();
          if (!(p.dbg & 2)) {
            if (left >= UG_KC) {
#pragma unroll
              for (int kk = 0; kk < UG_KC / 2; ++kk) {
                mma_bf16(tmem_d, hi_a | (lo_a + kk * (2 * UG_LBO >> 4)), hi_b | (lo_b + kk * 16), idesc, acc);
                acc = 1;
              }
            } else {
              for (int kk = 0; kk < left / 2; ++kk) {
                mma_bf16(tmem_d, hi_a | (lo_a + kk * (2 * UG_LBO >> 4)), hi_b | (lo_b + kk * 16), idesc, acc);
                acc = 1;
              }
            }
          }
          mma_commit(&empty_bar[slot]);  // A stage reusable once these MMAs have read it
          UG_TRACE(1, 2, st);
          left -= UG_KC;
          lo_b += UG_KC * 8;             // 8 chunks x 128 B >> 4
          lo_a += UG_STAGE_BYTES >> 4;
          if (++slot == nstage) { slot = 0; ph ^= 1; lo_a = lo_a0; }
        }
        mma_commit(&tfull_bar[buf]);     // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ================================================================ producers: gather rows into the A ring
    // Lane mapping: 8 consecutive lanes fetch the 8 consecutive 16-byte chunks (128 B of K) of ONE row; a warp instruction
    // covers 4 rows; a thread owns chunk column kc of rows  i*32 + pw*4 + rr,  i = 0..3.  Measured with ncu: LDGSTS is
    // cheap (~6 shared-memory wavefronts per instruction instead of 35) only when BOTH hold:
    //   * consecutive lanes read consecutive global addresses (a returning 128-byte line is handled as one unit), and
    //   * the chunks of that line land in different banks -- hence the 144-byte K stride (UG_LBO) of the A ring.
    // The tile's block of the index table (128 rows x SP entries) is staged in shared memory, double-buffered, the NEXT
    // tile's block being prefetched into registers while the current tile is gathered.
    const int pt = tid - (UG_EPI_WARPS + 1) * 32;  // 0..NPT-1
    const int pw = pt >> 5, kc = pt & 7, rr = (pt >> 3) & 3;
    const uint32_t a_base = smem_u32(a_ring);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    const int SP = SUM ? 2 * p.S : p.S;            // int32 words per table row (dgrad: 4 x uint16 per key)
    int32_t* idx_s = reinterpret_cast<int32_t*>(a_ring + (size_t)p.nstage * UG_STAGE_BYTES);  // [2][128*SP]
    const int nidx = UG_BM * SP;
    const unsigned rows_dst = (unsigned)p.rows_dst;
    // slot/channel walk of this thread's chunk column (compile-time shape)
    constexpr int CPS = CS / 8;                          // chunks per spiral slot
    constexpr int SPS = CS <= 64 ? UG_KC / CPS : 1;      // slots per stage (CS = 128: one slot spans two stages)
    const int sl0 = CS <= 64 ? kc / CPS : 0;             // slot of this chunk column within a stage
    const int c0 = CS <= 64 ? (kc % CPS) * 8 : kc * 8;   // channel offset (+64 on odd stages when CS = 128)
    // The tile's block of the index table goes global -> shared with cp.async (no registers: a register-held prefetch
    // array spilled under the register cap and serialised its loads), double-buffered, one commit group per block.
    // Blocks are WARP-PRIVATE: a warp stages exactly the table rows its own lanes gather (NIT groups of 4 consecutive
    // tile rows), so a block needs only cp.async.wait_group + __syncwarp -- no CTA-wide barrier at tile boundaries
    // (ncu: 12 % of the gather-sum kernel's stall samples sat on that barrier).
    const uint32_t idx_base = smem_u32(idx_s);
    const int ush = (SP % 4 == 0) ? 2 : ((SP % 2 == 0) ? 1 : 0);  // log2(words per copy unit): widest that divides a row
    constexpr int RW = 4 * NIT;     // table rows per warp per tile
    constexpr int PARTS = 32 / RW;  // lanes sharing one table row
    const int cq = lane % RW, cpart = lane / RW;
    const unsigned crow = (unsigned)((cq >> 2) * (UG_BM / NIT) + pw * 4 + (cq & 3));  // tile row this lane stages
    auto prefetch_idx_block = [&](unsigned m0, unsigned j0, int buf) {
      const bool on = (long long)(m0 + crow) < p.M;
      unsigned j = j0 + crow;
      if (j >= rows_dst) j = rows_dst >= (unsigned)UG_BM ? j - rows_dst : j % rows_dst;
      const int32_t* src = p.table + (on ? (size_t)j * SP : 0);
      const uint32_t dst = idx_base + (uint32_t)(buf * nidx + (int)crow * SP) * 4;
      if (ush == 2) {
        for (int w = cpart * 4; w < SP; w += PARTS * 4) cp_async16(dst + w * 4, src + w, on ? 16u : 0u);
      } else if (ush == 1) {
        for (int w = cpart * 2; w < SP; w += PARTS * 2) cp_async8(dst + w * 4, src + w, on ? 8u : 0u);
      } else {
        for (int w = cpart; w < SP; w += PARTS) cp_async4(dst + w * 4, src + w, on ? 4u : 0u);
      }
      cp_async_commit();
    };
    // tile coordinates (sample b0, first row j0) advance by a constant per iteration: no division in the loop
    const unsigned tile_rows = gridDim.x * (unsigned)UG_BM;
    const unsigned step_b = tile_rows / rows_dst, step_j = tile_rows - step_b * rows_dst;
    unsigned m0 = blockIdx.x * (unsigned)UG_BM;
    unsigned b0 = m0 / rows_dst, j0 = m0 - b0 * rows_dst;
    prefetch_idx_block(m0, j0, 0);
    cp_async_wait_group<0>();
    __syncwarp();
    uint32_t slot = 0, ph = 0;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int32_t* idx_cur = idx_s + (tcount & 1) * nidx;
      unsigned bn = b0 + step_b, jn = j0 + step_j;
      if (jn >= rows_dst) { jn -= rows_dst; bn += 1; }
      prefetch_idx_block(m0 + tile_rows, jn, (tcount + 1) & 1);  // next tile's block: in flight during this tile's stages
      bool valid[NIT];
      const __nv_bfloat16* srcc[NIT];
      uint32_t row_off[NIT];
      int irow[NIT], urow[NIT];
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const unsigned r = i * (UG_BM / NIT) + pw * 4 + rr;
        unsigned j = j0 + r, b = b0;
        if (j >= rows_dst) {
          if (rows_dst >= (unsigned)UG_BM) { j -= rows_dst; b += 1; } else { b += j / rows_dst; j %= rows_dst; }
        }
        valid[i] = (long long)(m0 + r) < p.M;
        if (SUM && p.skip_last && j == rows_dst - 1) valid[i] = false;
        srcc[i] = p.src + (valid[i] ? (size_t)b * p.rows_src * CS : 0) + c0;
        irow[i] = r * SP;
        urow[i] = (int)j;
        row_off[i] = (uint32_t)(r >> 3) * UG_SBO + (uint32_t)(r & 7) * 16 + (uint32_t)kc * UG_LBO;
      }
      int s = sl0;  // spiral slot of this thread's chunk column in the current stage
      // gather-sum: raw 16-byte chunks of the (up to four) inline entries of every item, loaded one stage ahead of their
      // use.  Entries 0 and 1 are always requested (predicated per lane); entries 2 and 3 exist for 4 % / 0.7 % of the keys,
      // so their loads and adds sit behind warp votes.  `more` keeps the votes: bit (2i) = some lane has entry 2 of item i,
      // bit (2i+1) = entry 3.
      uint4 v0[NIT], v1[NIT], v2[NIT], v3[NIT];
      bool over[NIT];
      uint32_t more = 0;
      auto issue_sum_loads = [&](int sn, int coff) {
        more = 0;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          uint2 q = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
          if (valid[i] && sn < p.S) q = *reinterpret_cast<const uint2*>(idx_cur + irow[i] + 2 * sn);
          const uint32_t e0 = q.x & 0xFFFFu, e1 = q.x >> 16, e2 = q.y & 0xFFFFu, e3 = q.y >> 16;
          const __nv_bfloat16* base = srcc[i] + coff;
          const uint4 z = make_uint4(0, 0, 0, 0);
          v0[i] = e0 < 0xFFFEu ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)e0 * CS)) : z;
          v1[i] = e1 < 0xFFFEu ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)e1 * CS)) : z;
          const bool h2 = e2 < 0xFFFEu, h3 = e3 < 0xFFFEu;
          if (__any_sync(0xFFFFFFFFu, h2)) {
            v2[i] = h2 ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)e2 * CS)) : z;
            more |= 1u << (2 * i);
            if (__any_sync(0xFFFFFFFFu, h3)) {  // entries are packed: an entry 3 implies an entry 2
              v3[i] = h3 ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)e3 * CS)) : z;
              more |= 2u << (2 * i);
            }
          }
          over[i] = e3 == 0xFFFEu;
        }
      };
      if (SUM) issue_sum_loads(s, 0);
      for (int st = 0; st < p.NS; ++st) {
        const int coff = (CS == 128 && (st & 1)) ? 64 : 0;
        const int step = (CS == 128) ? (st & 1) : SPS;
        UG_TRACE(0, 0, st);
        mbar_wait(empty_u32 + slot * 8, ph ^ 1);
        UG_TRACE(0, 1, st);
        const uint32_t dst0 = a_base + slot * UG_STAGE_BYTES;
        if (!SUM) {
          if (s < p.S && !(p.dbg & 1)) {
            int row[NIT];
#pragma unroll
            for (int i = 0; i < NIT; ++i) row[i] = idx_cur[irow[i] + s];
            const int zrow = p.skip_last ? p.rows_src - 1 : -1;  // known-zero source row: no memory access, zero-fill
#pragma unroll
            for (int i = 0; i < NIT; ++i)
              cp_async16(dst0 + row_off[i], srcc[i] + (size_t)row[i] * CS + coff, (valid[i] && row[i] != zrow) ? 16u : 0u);
          }
        } else if (!(p.dbg & 1)) {
          // fixed summation order: inline entries 0..3 ascending, then (rare: 5+ entries) the CSR tail from entry 3.
          // Packed bf16 adds: each step rounds the exact sum once, i.e. what an fp32 add + cast gives for two entries.
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            uint4 acc = bf16x8_add(v0[i], v1[i]);
            if (more & (1u << (2 * i))) acc = bf16x8_add(acc, v2[i]);
            if (more & (2u << (2 * i))) acc = bf16x8_add(acc, v3[i]);
            if (over[i]) {
              const int k = urow[i] * p.S + s;
              const int e0 = __ldg(p.keyptr + k), e1 = __ldg(p.keyptr + k + 1);
              for (int e = e0 + 3; e < e1; ++e)
                acc = bf16x8_add(acc, __ldg(reinterpret_cast<const uint4*>(srcc[i] + (size_t)__ldg(p.list + e) * CS + coff)));
            }
            if (s < p.S)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + row_off[i]), "r"(acc.x), "r"(acc.y),
                           "r"(acc.z), "r"(acc.w)
                           : "memory");
          }
        }
        UG_TRACE(0, 2, st);
        s += step;
        if (!SUM) {
          cp_async_mbar_arrive_noinc(full_u32 + slot * 8);  // one (counted) arrival when this thread's copies have landed
        } else {
          // Next stage's loads go out before the hand-shake and STAY in flight across it: the generic->async proxy fence is
          // executed by the CONSUMER (the MMA thread, after it acquires the full barrier), not here -- a producer-side
          // fence.proxy.async is a MEMBAR that drains this thread's outstanding loads, i.e. one exposed L2 round trip per
          // stage.  Ordering: st.shared -> __syncwarp -> arrive(release) -> try_wait(acquire) -> fence.proxy.async -> mma.
          if (st + 1 < p.NS) issue_sum_loads(s, (CS == 128 && ((st + 1) & 1)) ? 64 : 0);
          __syncwarp();
          if (lane == 0) mbar_arrive(full_u32 + slot * 8);
        }
        if (++slot == nstage) { slot = 0; ph ^= 1; }
      }
      // the next tile's index block must have landed, for every lane of this warp (blocks are warp-private)
      cp_async_wait_group<0>();
      __syncwarp();
      m0 += tile_rows; b0 = bn; j0 = jn;
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == UG_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static long long* g_trace = nullptr;
void umma_set_trace(long long* buf) { g_trace = buf; }

static size_t ug_smem_bytes(int NPAD, int Q, int nstage, int S, bool sum = true) {
  return ((size_t)NPAD * Q * 16 + 127) / 128 * 128 + (size_t)nstage * UG_STAGE_BYTES + (size_t)2 * UG_BM * ((sum ? 2 : 1) * S) * 4;
}
constexpr size_t UG_SMEM_MAX = 227 * 1024 - 2048;  // leave room for the static smem (barriers, bias)

static int ug_npad(int Cd) { return ((Cd + 15) / 16) * 16; }

bool umma_gather_gemm_supported(int Cs, int Cd, int S) {
  if (!(Cs == 8 || Cs == 16 || Cs == 32 || Cs == 64 || Cs == 128)) return false;
  if ((S * Cs / 8) & 1) return false;  // an MMA consumes two 16-byte K chunks
  const int NPAD = ug_npad(Cd);
  if (NPAD > 256 || Cd > 256) return false;
  if (S > 16) return false;  // per-tile index block: 8 entries per producer thread
  return ug_smem_bytes(NPAD, S * Cs / 8, UG_MIN_STAGES, S) <= UG_SMEM_MAX;
}

template <int CS, bool SUM> static int ug_launch(const UGParams& p, size_t smem, int want_per_sm, cudaStream_t st) {
  static bool attr_set = false;  // per instantiation; the attribute is sticky
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_gather_gemm_kernel<CS, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)UG_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // CTAs per SM from registers and shared memory (the runtime's occupancy query is conservative about the carveout);
  // CTAs are independent, so an over-estimate only costs a second wave, never correctness
  static int regs = 0;
  if (regs == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, umma_gather_gemm_kernel<CS, SUM>);
    if (e != cudaSuccess) return (int)e;
    regs = fa.numRegs;
    cudaFuncSetAttribute(umma_gather_gemm_kernel<CS, SUM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  }
  const int by_regs = 65536 / (((regs + 7) / 8 * 8) * ug_threads(SUM));
  const int by_smem = (int)((228 * 1024) / (smem + 2048));
  int occ = by_regs < by_smem ? by_regs : by_smem;
  if (occ < 1) occ = 1;
  int grid = kNumSMs * (occ < want_per_sm ? occ : want_per_sm);
  if (grid > p.num_tiles) grid = p.num_tiles;
  umma_gather_gemm_kernel<CS, SUM><<<grid, ug_threads(SUM), smem, st>>>(p);
  SHB_LAUNCH_CHECK();
  return 0;
}

int umma_gather_gemm(const void* src, const int32_t* table, const int32_t* keyptr, const int32_t* list, const void* w,
                     const float* bias,
                     void* dst, int B, int rows_src, int rows_dst, int S, int Cs, int Cd, int act, int zero_last,
                     int skip_last, bool sum_mode, cudaStream_t st) {
  UGParams p{};
  p.src = (const __nv_bfloat16*)src; p.table = table; p.keyptr = keyptr; p.list = list; p.w = (const __nv_bfloat16*)w;
  p.bias = bias;
  p.dst = (__nv_bfloat16*)dst;
  p.M = (long long)B * rows_dst;
  p.rows_src = rows_src; p.rows_dst = rows_dst; p.S = S; p.Cd = Cd; p.NPAD = ug_npad(Cd);
  p.Q = S * Cs / 8;
  p.NS = (p.Q + UG_KC - 1) / UG_KC;
  p.num_tiles = (int)((p.M + UG_BM - 1) / UG_BM);
  p.act = act; p.zero_last = zero_last; p.skip_last = skip_last;
  {
    static const int dbg = [] { const char* e = getenv("SHB_UMMA_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    p.trace = g_trace;
  }
  int nstage = 6;
  while (nstage > UG_MIN_STAGES && ug_smem_bytes(p.NPAD, p.Q, nstage, S, sum_mode) > UG_SMEM_MAX) --nstage;
  // two CTAs per SM when both fit with a deep enough ring: more gathers in flight
  int ctas_per_sm = 1;
  if (!sum_mode && ug_smem_bytes(p.NPAD, p.Q, 4, S, false) * 2 + 4096 <= UG_SMEM_MAX) { ctas_per_sm = 2; nstage = 4; }
  else if (!sum_mode && ug_smem_bytes(p.NPAD, p.Q, 3, S, false) * 2 + 4096 <= UG_SMEM_MAX) { ctas_per_sm = 2; nstage = 3; }
  p.nstage = nstage;
  uint32_t cols = 32;
  while (cols < 2u * p.NPAD) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = ug_smem_bytes(p.NPAD, p.Q, nstage, S, sum_mode);
#define UG_DISPATCH(CSV)                                                                   \
  case CSV:                                                                                \
    return sum_mode ? ug_launch<CSV, true>(p, smem, ctas_per_sm, st) : ug_launch<CSV, false>(p, smem, ctas_per_sm, st);
  switch (Cs) {
    UG_DISPATCH(8)
    UG_DISPATCH(16)
    UG_DISPATCH(32)
    UG_DISPATCH(64)
    UG_DISPATCH(128)
    default: return SHB_E_SHAPE;
  }
#undef UG_DISPATCH
}

}  // namespace shb

// ================================================================================================ weight gradient
//   gw[n, s*Cin + c] = sum_{m=(b,j)} gz[m, n] * x[b, table[j,s], c]
// as  D[k, n] = sum_m At[k, m] * Bt[n, m]:  UMMA-M = 128-wide blocks of k = s*Cin+c, UMMA-N = Cout, UMMA-K = rows m.
// Both operands are MN-major (a gathered x row holds consecutive k, a gz row holds consecutive n), so every 16-byte
// cp.async chunk lands unchanged in the un-swizzled core-matrix layout.  The ENTIRE (K x Cout) fp32 accumulator stays
// in TMEM (ceil(K/128)*NPAD <= 512 columns) while the CTA streams its contiguous share of the rows exactly once; each CTA
// then writes one fp32 partial, and a fixed-order reduction over CTAs finishes gw (bit-reproducible, no atomics).
namespace shb {

constexpr int UW_MAX_STAGES = 6;

struct UWParams {
  const __nv_bfloat16* x;   // (B, rows_in, CIN)
  const int32_t* table;     // (rows_out, S)
  const __nv_bfloat16* gz;  // (B, rows_out, Cout)
  float* ws;                // [grid][Cout*K] fp32 partials
  long long M, rows_per_cta;
  int rows_in, rows_out, S, Cout, NPAD, K, Q, KT, NQ;  // Q = K/8 chunks per x row set, NQ = NPAD/8
  int src_dummy_zero;  // x[b, rows_in-1, :] is known zero: zero-fill instead of gathering it
  int nstage;
  uint32_t a_stage_bytes, b_stage_bytes, tmem_cols;
};

// NI: index registers per producer thread per stage (>= Q / QL); the NI = 8 variant fits two CTAs per SM (MINB = 2)
template <int CIN, int BKM, int NI, int MINB>
__global__ void __launch_bounds__(UG_THREADS, MINB) umma_wgrad_kernel(const UWParams p) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[UW_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UW_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  const long long mbeg = (long long)blockIdx.x * p.rows_per_cta;
  const long long mend = min(p.M, mbeg + p.rows_per_cta);
  const int nst = mbeg < mend ? (int)((mend - mbeg + BKM - 1) / BKM) : 0;

  if (tid == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full_bar[i], UG_PROD_THREADS);  // one cp.async-completion arrival per producer thread
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == UG_EPI_WARPS) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // MN-direction core-matrix stride (+16 B of padding: the two 16-byte halves of a returning 32-byte sector are consecutive
  // MN chunks and must not share banks); K-direction (row groups of 8) stride = 128
  constexpr uint32_t SBO = BKM * 16 + 16;

  if (warp < UG_EPI_WARPS) {
    // ================================================================ epilogue: TMEM -> fp32 partial in the workspace
    float* ws = p.ws + (size_t)blockIdx.x * ((size_t)p.Cout * p.K);
    if (nst > 0) {
      mbar_wait_backoff(&done_bar, 0, 1000);
      tc_fence_after();
    }
    for (int t = 0; t < p.KT; ++t) {
      const int k = t * 128 + warp * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t)(t * p.NPAD) + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < p.NPAD; c0 += 16) {
        uint32_t r[16];
        if (nst > 0) {
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0u;
        }
        if (k < p.K) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < p.Cout) ws[(size_t)(c0 + i) * p.K + k] = __uint_as_float(r[i]);
        }
      }
    }
  } else if (warp == UG_EPI_WARPS) {
    // ================================================================ MMA issuer
    if (lane == 0 && nst > 0) {
      const uint32_t idesc = idesc_bf16_f32(128, p.NPAD, 1, 1);
      const uint32_t base = smem_u32(dyn_smem);
      // descriptor = hi (SBO, version) | lo (start address >> 4, LBO = 128 B); only the start address moves
      const uint64_t hi = ((uint64_t)(SBO >> 4) << 32) | ((uint64_t)1 << 46);
      const uint32_t lo_flags = (128u >> 4) << 16;
      uint32_t slot = 0, ph = 0;
      for (int st = 0; st < nst; ++st) {
        mbar_wait_backoff(&full_bar[slot], ph, 20);
        fence_proxy_async_smem();  // the stage was written by cp.async (generic proxy); tcgen05.mma reads through the async proxy
        tc_fence_after();
        const uint32_t a_lo = ((base + slot * stage_bytes) >> 4) | lo_flags, b_lo = a_lo + (p.a_stage_bytes >> 4);
#pragma unroll
        for (int kk = 0; kk < BKM / 16; ++kk) {
          const uint64_t db = hi | (b_lo + kk * 16);
          uint32_t a_t = a_lo + kk * 16;
          for (int t = 0; t < p.KT; ++t) {
            mma_bf16(tmem_base + (uint32_t)(t * p.NPAD), hi | a_t, db, idesc, (st | kk) != 0);
            a_t += SBO;  // 16 MN-chunks of SBO bytes, >> 4
          }
        }
        mma_commit(&empty_bar[slot]);
        if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
      }
      mma_commit(&done_bar);
    }
    __syncwarp();
  } else {
    // ================================================================ producers
    // consecutive lanes fetch consecutive 16-byte chunks of one gathered row (one L1TEX wavefront per row per 128 B);
    // each warp owns BKM/8 rows of the stage, a thread owns chunk lanes ql, ql+QL, ...
    const int pt = tid - (UG_EPI_WARPS + 1) * 32;
    const int pw = pt >> 5, l = pt & 31;
    constexpr int RPW = BKM / 8;   // rows per warp
    constexpr int QL = 32 / RPW;   // chunk lanes per row
    const int mr = pw * RPW + l / QL, ql = l % QL;
    const uint32_t base = smem_u32(dyn_smem);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    int rowc[NI];
    bool validc = false;
    long long mc = 0;
    const __nv_bfloat16* xbc = p.x;
    // (sample, row) of this thread's stage row, advanced by BKM per stage: no division in the loop
    long long mnext = mbeg + mr;
    unsigned bnext = (unsigned)(mnext / p.rows_out), jnext = (unsigned)(mnext - (long long)bnext * p.rows_out);
    auto fetch_stage_idx = [&](int st, int* rowv, bool& valid, long long& m, const __nv_bfloat16*& xb) {
      m = mnext;
      valid = st < nst && m < mend;
      const unsigned b = valid ? bnext : 0u, j = valid ? jnext : 0u;
      xb = p.x + (size_t)b * p.rows_in * CIN;
      const int32_t* trow = p.table + (size_t)j * p.S;
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int q = ql + u * QL;
        rowv[u] = (valid && q < p.Q) ? __ldg(trow + (q * 8) / CIN) : 0;
      }
      mnext += BKM;
      jnext += BKM;
      while (jnext >= (unsigned)p.rows_out) { jnext -= (unsigned)p.rows_out; ++bnext; }
    };
    fetch_stage_idx(0, rowc, validc, mc, xbc);
    const int zrow = p.src_dummy_zero ? p.rows_in - 1 : -1;
    uint32_t slot = 0, ph = 0;
    for (int st = 0; st < nst; ++st) {
      int rown[NI];
      bool validn;
      long long mn;
      const __nv_bfloat16* xbn;
      fetch_stage_idx(st + 1, rown, validn, mn, xbn);  // next stage's indices: in flight while this stage is issued
      mbar_wait(empty_u32 + slot * 8, ph ^ 1);
      const uint32_t a_dst = base + slot * stage_bytes + (uint32_t)mr * 16;
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int q = ql + u * QL;
        if (q < p.Q)
          cp_async16(a_dst + (uint32_t)q * SBO, xbc + (size_t)rowc[u] * CIN + (q * 8) % CIN,
                     (validc && rowc[u] != zrow) ? 16u : 0u);
      }
      const uint32_t b_dst = a_dst + p.a_stage_bytes;
      for (int q = ql; q < p.NQ; q += QL) {
        const bool in = validc && (q * 8 < p.Cout);
        cp_async16(b_dst + (uint32_t)q * SBO, p.gz + (validc ? mc : 0) * p.Cout + q * 8, in ? 16u : 0u);
      }
      // one (counted) arrival on the stage's barrier when this thread's copies have landed: nothing waits here, the whole
      // ring depth stays available to the gathers (the generic->async proxy fence is on the consumer side)
      cp_async_mbar_arrive_noinc(full_u32 + slot * 8);
#pragma unroll
      for (int u = 0; u < NI; ++u) rowc[u] = rown[u];
      validc = validn;
      mc = mn;
      xbc = xbn;
      if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == UG_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// out[i] = sum over parts of ws[c*n + i], in a FIXED order: warp y of the block sums parts c = y, y+8, ... (coalesced
// 128-byte rows of 32 consecutive outputs), then warp 0 adds the eight partials in ascending y.  Deterministic, and
// parallel over the parts (the naive one-thread-per-output loop was latency-bound: 148 dependent-free but serial loads).
__global__ void __launch_bounds__(256) uw_reduce_kernel(const float* __restrict__ ws, int parts, long long n,
                                                        float* __restrict__ out) {
  __shared__ float red[8][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + x;
  float a = 0.f;
  if (i < n)
    for (int c = y; c < parts; c += 8) a += ws[(size_t)c * n + i];
  red[y][x] = a;
  __syncthreads();
  if (y == 0 && i < n) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][x];
    out[i] = s;
  }
}

// column sums of gz (bias gradient): per-block partials in fixed order, then a fixed-order final pass
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __nv_bfloat16* __restrict__ gz, long long M, int C,
                                                             int CP /*pow2 >= C, <= 256*/, long long rows_per_block,
                                                             float* __restrict__ part) {
  __shared__ float red[256];
  const int tx = threadIdx.x % CP, ty = threadIdx.x / CP, ny = 256 / CP;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a = 0.f;
  if (tx < C)
    for (long long r = r0 + ty; r < r1; r += ny) a += __bfloat162float(gz[r * C + tx]);
  red[threadIdx.x] = a;
  __syncthreads();
  if (ty == 0 && tx < C) {
    float s = 0.f;
    for (int q = 0; q < ny; ++q) s += red[q * CP + tx];
    part[(size_t)blockIdx.x * C + tx] = s;
  }
}

static size_t uw_stage_bytes(int KT, int NPAD, int bkm) { return ((size_t)KT * 16 + (size_t)(NPAD / 8)) * (bkm * 16 + 16); }

struct UWPlan { int bkm, nstage, KT, NPAD, grid; size_t smem; uint32_t cols; bool ok; bool two; };

static UWPlan uw_plan(int Cin, int Cout, int S) {
  UWPlan pl{};
  pl.ok = false;
  if (!(Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128)) return pl;
  if ((Cout & 7) != 0) return pl;  // gz rows must be whole 16-byte chunks for cp.async
  pl.NPAD = ug_npad(Cout);
  const int K = S * Cin;
  pl.KT = (K + 127) / 128;
  if (pl.KT * pl.NPAD > 512 || pl.NPAD > 256) return pl;
  if (K / 8 > 16 * 8) return pl;  // producer keeps Q / QL <= 16 index registers per stage
  pl.cols = 32;
  while ((int)pl.cols < pl.KT * pl.NPAD) pl.cols <<= 1;
  for (int bkm : {32, 16}) {
    const size_t sb = uw_stage_bytes(pl.KT, pl.NPAD, bkm);
    int ns = (int)(UG_SMEM_MAX / sb);
    if (ns > UW_MAX_STAGES) ns = UW_MAX_STAGES;
    if (ns >= 3) { pl.bkm = bkm; pl.nstage = ns; pl.smem = sb * ns; pl.ok = true; break; }
  }
  pl.grid = kNumSMs;
  // two CTAs per SM when TMEM (2 x cols <= 512), shared memory (2 x 3 stages) and the index registers (Q / QL <= 8) allow:
  // the kernel is bound by gather latency/L1TEX issue, and a second CTA doubles the gathers in flight
  pl.two = false;
  if (pl.ok && pl.bkm == 32 && K / 8 <= 64 && pl.cols * 2 <= 512) {
    const size_t sb = uw_stage_bytes(pl.KT, pl.NPAD, 32);
    int ns = (int)((UG_SMEM_MAX / 2 - 2048) / sb);
    if (ns > 4) ns = 4;
    if (ns >= 3) { pl.two = true; pl.nstage = ns; pl.smem = sb * ns; pl.grid = 2 * kNumSMs; }
  }
  return pl;
}

bool umma_wgrad_supported(int Cin, int Cout, int S) { return uw_plan(Cin, Cout, S).ok; }

size_t umma_wgrad_workspace(int B, int rows_out, int S, int Cin, int Cout) {
  (void)B; (void)rows_out;
  // per-CTA fp32 partials of gw, then per-block partials of the bias gradient
  return (size_t)2 * kNumSMs * Cout * S * Cin * sizeof(float) + (size_t)4 * kNumSMs * Cout * sizeof(float);
}

template <int CIN, int BKM, int NI, int MINB> static int uw_launch(const UWParams& p, int grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_wgrad_kernel<CIN, BKM, NI, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)UG_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    cudaFuncSetAttribute(umma_wgrad_kernel<CIN, BKM, NI, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    attr_set = true;
  }
  umma_wgrad_kernel<CIN, BKM, NI, MINB><<<grid, UG_THREADS, smem, st>>>(p);
  SHB_LAUNCH_CHECK();
  return 0;
}

int umma_wgrad(const void* x, const int32_t* table, const void* gz, float* gw, float* gb, void* workspace, int B,
               int rows_in, int rows_out, int S, int Cin, int Cout, int src_dummy_zero, cudaStream_t st) {
  const UWPlan pl = uw_plan(Cin, Cout, S);
  if (!pl.ok) return SHB_E_SHAPE;
  UWParams p{};
  p.x = (const __nv_bfloat16*)x; p.table = table; p.gz = (const __nv_bfloat16*)gz; p.ws = (float*)workspace;
  p.M = (long long)B * rows_out;
  int grid = pl.grid;
  const long long per = ((p.M + grid - 1) / grid + pl.bkm - 1) / pl.bkm * pl.bkm;
  grid = (int)((p.M + per - 1) / per);
  p.rows_per_cta = per;
  p.rows_in = rows_in; p.rows_out = rows_out; p.S = S; p.Cout = Cout; p.NPAD = pl.NPAD; p.K = S * Cin; p.Q = p.K / 8;
  p.KT = pl.KT; p.NQ = pl.NPAD / 8; p.nstage = pl.nstage;
  p.a_stage_bytes = (uint32_t)pl.KT * 16 * (pl.bkm * 16 + 16); p.b_stage_bytes = (uint32_t)(pl.NPAD / 8) * (pl.bkm * 16 + 16);
  p.tmem_cols = pl.cols;
  p.src_dummy_zero = src_dummy_zero;
  int rc;
#define UW_DISPATCH(C)                                                                                          \
  case C:                                                                                                       \
    rc = pl.two ? uw_launch<C, 32, 8, 2>(p, grid, pl.smem, st)                                                   \
                : (pl.bkm == 32 ? uw_launch<C, 32, 16, 1>(p, grid, pl.smem, st) : uw_launch<C, 16, 16, 1>(p, grid, pl.smem, st)); \
    break;
  switch (Cin) {
    UW_DISPATCH(8)
    UW_DISPATCH(16)
    UW_DISPATCH(32)
    UW_DISPATCH(64)
    UW_DISPATCH(128)
    default: return SHB_E_SHAPE;
  }
#undef UW_DISPATCH
  if (rc != 0) return rc;
  const long long n = (long long)Cout * p.K;
  uw_reduce_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(p.ws, grid, n, gw);
  SHB_LAUNCH_CHECK();
  if (gb != nullptr) {
    float* part = (float*)workspace + (size_t)2 * kNumSMs * Cout * p.K;
    int CP = 1;
    while (CP < Cout) CP <<= 1;
    if (CP > 256) return SHB_E_SHAPE;
    const int blocks = 4 * kNumSMs;
    const long long rpb = (p.M + blocks - 1) / blocks;
    colsum_partial_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)gz, p.M, Cout, CP, rpb, part);
    SHB_LAUNCH_CHECK();
    uw_reduce_kernel<<<(Cout + 31) / 32, 256, 0, st>>>(part, blocks, Cout, gb);
    SHB_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace shb
