// SpiralConv forward and input gradient in the slab layout (shb_slab.cuh): one persistent, warp-specialised tcgen05 kernel.
//
//   forward (models.py:34-53):   y[j]  = mask * act( sum_s  x[table[j,s]] . W_s^T + b )        entries of j: (table[j,s], s)
//   input gradient            :  gx[u] = act'(y_prev[u]) * sum_{(j,s): table[j,s]=u} gz[j] . W_s   entries of u: (j, s)
//
// Both are "for every destination row: a list of (source row, slot) entries; accumulate  slab(source) x Wop[slot]".
// A tile is one destination row x one 128-sample batch chunk: M = 128 samples, N = destination channels, and every entry
// contributes K = source channels.  The source slab arrives with ONE cp.async.bulk (TMA) already in UMMA K-major form, the
// per-slot weight operands sit in shared memory for the lifetime of the CTA, the sum over entries happens in the TMEM
// accumulator in list order -- fixed order, fp32, no atomics, no register gather-sum -- and the epilogue (bias, activation or
// activation derivative, dummy-row mask, bf16 / hi-lo split) writes 512 contiguous bytes per warp and channel chunk.
//
// Roles: warp 0 = TMA producer (entry lists -> bulk copies into a ring of stages), warp 1 = MMA issuer (one thread),
// warps 2-5 = epilogue (TMEM -> registers -> HBM), double-buffered accumulator.
#include <stdlib.h>

#include "shb_common.cuh"
#include "shb_internal.h"
#include "shb_slab.cuh"

namespace shb {

using namespace umma;
using namespace slab;

constexpr int SC_LEAD_WARPS = 3;                // two producer warps (even / odd ring stages) and the MMA warp
constexpr int SC_MAX_THREADS = SC_LEAD_WARPS * 32 + 4 * 128;  // + up to four groups of four epilogue warps
constexpr int SC_MAX_STAGES = 12;
constexpr int SC_MAX_SPS = 8;  // slabs per ring stage: a stage is up to 32 KB behind one barrier round trip

struct SlabConvParams {
  const uint8_t* src;      // slab tensor, CS channels, P planes
  const int32_t* ptr;      // (rows_dst + 1) entry ranges
  const int32_t* ent;      // (source row << 5) | slot
  const uint8_t* w_img;    // this pass's weight operand image(s): [P][NP/8][Q][8 n][8 k] bf16 (plane stride img_plane_stride)
  const float* bias;       // fp32, already offset to this pass's first channel; or null
  uint8_t* dst;            // slab tensor, Cd channels, P planes
  const uint8_t* ymul;     // slab tensor shaped like dst or null: dst = act'(ymul) * acc
  uint32_t img_plane_stride;
  int NB, rows_dst;
  int CS, Q, NP;           // source channels; S*CS/8; accumulator columns of this pass (multiple of 16)
  int n0, ncols, Cd;       // first destination channel of this pass, channels written (multiple of 8), destination channels
  int nbias;               // entries of `bias` that exist from n0 on
  int act, act_mul, zero_last;
  int P;
  int nstage, SPS;
  int cpw, EG;             // accumulator columns per epilogue warp (8/16/32), epilogue groups (threads = 96 + 128*EG)
  int num_tiles;
  uint32_t tmem_cols;
};

// Debug timeline (build with SHB_NVCC_FLAGS=-DSHB_SLAB_TRACE): per CTA, cycles each role spends waiting / working.
//   [0] producer wait-empty  [1] producer total   [2] MMA wait-full  [3] MMA wait-tmem  [4] MMA total
//   [5] epilogue wait-accumulator (warp 2)  [6] epilogue total (warp 2)  [7] tiles
#ifdef SHB_SLAB_TRACE
__device__ long long g_slab_trace[2 * kNumSMs * 16];
#define SC_T0(var) const long long var = clock64()
#define SC_ACC(idx, t0) trace_acc[idx] += clock64() - (t0)
#else
#define SC_T0(var) do { } while (0)
#define SC_ACC(idx, t0) do { } while (0)
#endif

// DUAL: two CTAs per SM (half the shared memory and at most two epilogue groups each).  Every role of a CTA is a dependent
// instruction chain that leaves its SM sub-partition idle most of the time (ncu: 0.3 issue slots / cycle, 17 % of the warp
// slots); a sibling CTA fills those gaps.  Used whenever the weight image leaves room for a ring in half an SM.
constexpr int SC_DUAL_THREADS = SC_LEAD_WARPS * 32 + 2 * 128;
template <int P, int NK, bool DUAL>
__global__ void __launch_bounds__(DUAL ? SC_DUAL_THREADS : SC_MAX_THREADS, DUAL ? 2 : 1) slab_conv_kernel(const SlabConvParams p) {
#ifdef SHB_SLAB_TRACE
  long long trace_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long trace_start = clock64();
#endif
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[SC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[SC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t img_bar;
  __shared__ __align__(16) unsigned long long src_s[SC_MAX_STAGES * SC_MAX_SPS];  // per slab: global source address
  __shared__ uint32_t hdr_s[SC_MAX_STAGES];                       // per stage: first/last-of-tile flags, slab count
  __shared__ __align__(16) uint32_t lb_s[SC_MAX_STAGES * SC_MAX_SPS];  // per slab: B-descriptor low word of its slot
  __shared__ uint32_t tile_empty_s[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t img_bytes1 = (uint32_t)p.NP * p.Q * 16;            // one plane
  // 8-channel sources read one core matrix past their slot's (see the MMA issuer): 128 B of zeros behind the image
  const uint32_t img_region = ((P * img_bytes1 + (p.CS == 8 ? 128u : 0u) + 1023) / 1024) * 1024;
  const uint32_t slab_b = (uint32_t)P * p.CS * 256;
  const uint32_t stage_b = slab_b * p.SPS;
  const uint32_t smem0 = smem_u32(dyn_smem);
  const uint32_t ring0 = smem0 + img_region;
  const uint32_t zero0 = ring0 + (uint32_t)p.nstage * stage_b;     // 2 KB of zeros (8-channel sources only)
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);

  // ---------------------------------------------------------------- prologue
  if (p.CS == 8) {
    for (int i = tid; i < 128 / 16; i += blockDim.x) *reinterpret_cast<uint4*>(dyn_smem + P * img_bytes1 + i * 16) = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < PLANE_STRIDE / 16; i += blockDim.x)
      *reinterpret_cast<uint4*>(dyn_smem + img_region + (size_t)p.nstage * stage_b + i * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < 256; i += blockDim.x) bias_s[i] = (p.bias != nullptr && i < p.nbias) ? __ldg(p.bias + i) : 0.f;
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * p.EG);
    }
    mbar_init(&img_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 2) {
    // ================================================================ producers: entry lists -> bulk copies
    // Two warps run the same walk over tiles and stages; warp w does the serial part (barrier round trip, bulk-copy issue) of
    // the stages whose running number is w mod 2 -- that part is a dependent instruction chain of ~1000 cycles per stage-triple.
    const uint32_t pw = (uint32_t)warp;
    uint32_t stage_no = 0;
    // One thread's mbarrier / bulk-copy instructions cost ~50 cycles each back to back (measured, tests/cuda/slab_probe.cu:
    // 200 cycles per wait + expect_tx + copy), so a ring stage carries up to SPS slabs behind ONE wait and ONE expect_tx.
    if (warp == 0 && lane == 0) {  // weight operand image(s): resident for the whole kernel
      mbar_expect_tx(smem_u32(&img_bar), P * img_bytes1);
      for (int pl = 0; pl < P; ++pl)
        for (uint32_t off = 0; off < img_bytes1; off += 32768) {
          const uint32_t n = img_bytes1 - off < 32768 ? img_bytes1 - off : 32768;
          bulk_load(smem0 + pl * img_bytes1 + off, p.w_img + (size_t)pl * p.img_plane_stride + off, n, smem_u32(&img_bar));
        }
    }
    uint32_t slot = 0, ph = 0;
    const size_t row_stride = (size_t)p.NB * slab_b;
    const uint32_t lo_b0 = (smem0 >> 4) | ((128u >> 4) << 16);          // B descriptor low word of slot 0 (LBO = 128)
    const uint32_t b_step = (uint32_t)p.CS;                             // (CS/8 chunks) * 128 B >> 4 per slot
    const int sps_shift = p.SPS == 8 ? 3 : (p.SPS == 4 ? 2 : (p.SPS == 2 ? 1 : 0));
    const int kpos = lane & (p.SPS - 1), sidx_of_lane = lane >> sps_shift;
    // tile -> (row u, chunk q) without divisions: a tile step adds gridDim.x = gdiv * NB + gmod
    const int gdiv = (int)gridDim.x / p.NB, gmod = (int)gridDim.x - gdiv * p.NB;
    int t = blockIdx.x;
    int u = t / p.NB, q = t - u * p.NB;
    int e0 = 0, e1 = 0, mine = 0;
    if (t < p.num_tiles) {
      e0 = __ldg(p.ptr + u);
      e1 = __ldg(p.ptr + u + 1);
      if (e0 + lane < e1) mine = __ldg(p.ent + e0 + lane);
    }
    // two tiles ahead: entry range; one tile ahead: range + first 32 entries (registers)
    int t1 = t + gridDim.x, u1 = u + gdiv, q1 = q + gmod, f0 = 0, f1 = 0;
    if (q1 >= p.NB) { q1 -= p.NB; ++u1; }
    if (t1 < p.num_tiles) { f0 = __ldg(p.ptr + u1); f1 = __ldg(p.ptr + u1 + 1); }
    while (t < p.num_tiles) {
      int mine_n = 0;
      if (f0 + lane < f1) mine_n = __ldg(p.ent + f0 + lane);
      const int t2 = t1 + gridDim.x;
      int u2 = u1 + gdiv, q2 = q1 + gmod;
      if (q2 >= p.NB) { q2 -= p.NB; ++u2; }
      int g0 = 0, g1 = 0;
      if (t2 < p.num_tiles) { g0 = __ldg(p.ptr + u2); g1 = __ldg(p.ptr + u2 + 1); }
      const uint8_t* srcq = p.src + (size_t)q * slab_b;
      if (e0 == e1) {  // no entries: the epilogue writes act(bias) / zeros; the MMA thread still hands the buffer over
        if (lane == 0 && (stage_no & 1u) == pw) {
          mbar_wait(empty0 + slot * 8, ph ^ 1);
          hdr_s[slot] = 3u;
          mbar_arrive(full0 + slot * 8);
        }
        ++stage_no;
        if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
      }
      // Lane L of the warp owns entry L of the current block of 32: it computes its own source address and B-descriptor word
      // (vector code, all lanes at once) and parks them in shared memory; ONE elected lane then does the barrier round trip of
      // the stage and issues the <= SPS bulk copies from warp-uniform operands (a per-lane issue costs ~100 cycles per copy).
      for (int base = e0; base < e1; base += 32) {
        const int cnt = e1 - base < 32 ? e1 - base : 32;
        if (base > e0) mine = lane < cnt ? __ldg(p.ent + base + lane) : 0;  // long lists: the rest is fetched in place
        const unsigned long long my_src = (unsigned long long)(srcq + (size_t)(mine >> 5) * row_stride);
        const uint32_t my_lb = lo_b0 + (uint32_t)(mine & 31) * b_step;
        const int nstg = (cnt + p.SPS - 1) >> sps_shift;
        for (int sidx = 0; sidx < nstg; ++sidx) {
          if (((stage_no++) & 1u) != pw) {  // the other producer warp's stage
            if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
            continue;
          }
          const int ns = cnt - (sidx << sps_shift) < p.SPS ? cnt - (sidx << sps_shift) : p.SPS;
          const bool active = sidx_of_lane == sidx && lane < cnt;
          const uint32_t hdr = (base + (sidx << sps_shift) == e0 ? 1u : 0u) | (base + (sidx << sps_shift) + ns == e1 ? 2u : 0u) |
                               ((uint32_t)ns << 2);
          if (elect_one()) {
            SC_T0(tw);
            mbar_wait(empty0 + slot * 8, ph ^ 1);
            SC_ACC(0, tw);
            hdr_s[slot] = hdr;
          }
          __syncwarp();
          if (active) {
            lb_s[slot * SC_MAX_SPS + kpos] = my_lb;
            src_s[slot * SC_MAX_SPS + kpos] = my_src;
          }
          __syncwarp();
          if (elect_one()) {
            SC_T0(tc);
            const uint32_t bar = full0 + slot * 8;
            mbar_expect_tx(bar, (uint32_t)ns * slab_b);
            const uint32_t dst = ring0 + slot * stage_b;
            const unsigned long long* sp = &src_s[slot * SC_MAX_SPS];
#pragma unroll
            for (int k = 0; k < SC_MAX_SPS; ++k)
              if (k < ns) bulk_load(dst + k * slab_b, reinterpret_cast<const void*>(sp[k]), slab_b, bar);
            SC_ACC(10, tc);
          }
          if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
        }
      }
      t = t1; u = u1; q = q1; e0 = f0; e1 = f1; mine = mine_n;
      t1 = t2; u1 = u2; q1 = q2; f0 = g0; f1 = g1;
    }
#ifdef SHB_SLAB_TRACE
    if (lane == 0 && warp == 0) {
      g_slab_trace[blockIdx.x * 16 + 0] = trace_acc[0];
      g_slab_trace[blockIdx.x * 16 + 1] = clock64() - trace_start;
      g_slab_trace[blockIdx.x * 16 + 8] = trace_acc[8];
      g_slab_trace[blockIdx.x * 16 + 9] = trace_acc[9];
      g_slab_trace[blockIdx.x * 16 + 10] = trace_acc[10];
    }
#endif
  } else if (warp == 2) {
    // ================================================================ MMA issuer (one thread)
    // tcgen05.commit only tracks the issuing thread's MMAs, so one thread issues them all; what it executes per entry is cut
    // to the bone (every instruction of this dependent chain costs ~5 cycles): the B-descriptor word comes ready-made from the
    // producer, the A-descriptor word advances by a constant.
    if (elect_one()) {
      const uint32_t idesc = idesc_bf16_f32(CHUNK, p.NP, 0, 0);
      const uint64_t hi_b = ((uint64_t)(((uint32_t)p.Q * 128) >> 4) << 32) | ((uint64_t)1 << 46);       // SBO = Q*128
      const uint64_t hi_a = ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);                         // SBO = 128
      const uint32_t slab16 = slab_b >> 4;
      // A: K-major, SBO = 128 (next 8 samples), LBO = 2048 (next 8 channels).  8-channel sources pair the slab with a block of
      // zeros through LBO = zero0 - slab: moving to the next slab adds slab16 to the address field and takes it off the LBO field
      const uint32_t a_step = p.CS >= 16 ? slab16 : slab16 - (slab16 << 16);
      const uint32_t plane_a = p.CS >= 16 ? ((uint32_t)p.CS * 256) >> 4 : (((uint32_t)p.CS * 256) >> 4) - ((((uint32_t)p.CS * 256) >> 4) << 16);
      const uint32_t plane_b = img_bytes1 >> 4;
      mbar_wait(smem_u32(&img_bar), 0);
      uint32_t slot = 0, ph = 0;
      int tcount = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        SC_T0(tw3);
        mbar_wait_parked(smem_u32(&tempty_bar[buf]), ((tcount >> 1) & 1) ^ 1, 1000);
        SC_ACC(3, tw3);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.NP);
        uint32_t acc = 0;
        for (;;) {
          SC_T0(tw2);
          mbar_wait_parked(full0 + slot * 8, ph, 1000);
          SC_ACC(2, tw2);
          const uint32_t hdr = hdr_s[slot];
          const uint4 w0 = *reinterpret_cast<const uint4*>(&lb_s[slot * SC_MAX_SPS]);
          const uint4 w1 = *reinterpret_cast<const uint4*>(&lb_s[slot * SC_MAX_SPS + 4]);
          const uint32_t lbs[SC_MAX_SPS] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          tc_fence_after();
          const uint32_t ns = (hdr >> 2) & 15u;
          SC_T0(ti);
          const uint32_t a_slab = ring0 + slot * stage_b;
          const uint32_t lbo_a = p.CS >= 16 ? (uint32_t)PLANE_STRIDE : zero0 - a_slab;
          uint32_t la0 = ((a_slab & 0x3FFFFu) >> 4) | ((lbo_a >> 4) << 16);
#pragma unroll
          for (uint32_t k = 0; k < (uint32_t)SC_MAX_SPS; ++k) {
            if (k < ns) {
              uint32_t la = la0, lb = lbs[k];
#pragma unroll
              for (uint32_t kk = 0; kk < (uint32_t)NK; ++kk) {
                mma_bf16(tmem_d, hi_a | la, hi_b | lb, idesc, acc);
                acc = 1;
                if (P == 2) {  // x ~ xh + xl, w ~ wh + wl: xh.wh + xl.wh + xh.wl; xl.wl (<= 2^-18 of the term) is dropped
                  mma_bf16(tmem_d, hi_a | (la + plane_a), hi_b | lb, idesc, 1);
                  mma_bf16(tmem_d, hi_a | la, hi_b | (lb + plane_b), idesc, 1);
                }
                la += 2 * PLANE_STRIDE >> 4;
                lb += 16;
              }
              la0 += a_step;
            }
          }
          SC_ACC(11, ti);
          SC_T0(tj);
          mma_commit_u32(empty0 + slot * 8);  // stage reusable once these MMAs have read it
          SC_ACC(12, tj);
          if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
          if (hdr & 2u) {
            tile_empty_s[buf] = ns == 0 ? 1u : 0u;
            break;
          }
        }
        __threadfence_block();
        mma_commit_u32(smem_u32(&tfull_bar[buf]));  // accumulator complete
      }
#ifdef SHB_SLAB_TRACE
      g_slab_trace[blockIdx.x * 16 + 2] = trace_acc[2];
      g_slab_trace[blockIdx.x * 16 + 3] = trace_acc[3];
      g_slab_trace[blockIdx.x * 16 + 4] = clock64() - trace_start;
      g_slab_trace[blockIdx.x * 16 + 11] = trace_acc[11];
      g_slab_trace[blockIdx.x * 16 + 12] = trace_acc[12];
#endif
    }
    __syncwarp();
  } else {
    // ================================================================ epilogue warps: TMEM -> bias / act / act' / mask -> HBM
    // EG groups of 4 warps; a group owns `cpw` consecutive accumulator columns (a warp can only read its own TMEM lane
    // quarter, warp_id % 4).  The instruction stream of a warp is one dependent chain (~5 cycles per instruction) and a global
    // load issued next to a saturated TMA ring takes ~2000 cycles, so: many narrow warps, and the act' operands are requested
    // before the wait on the accumulator.
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int grp = (warp - SC_LEAD_WARPS) >> 2;  // column group
    const int b = quarter * 32 + lane;       // sample within the chunk
    const int cpw = p.cpw, ngr = cpw >> 3;   // columns per warp (8, 16 or 32), 8-column vectors per warp
    const int col0 = grp * cpw;
    const size_t dslab = slab_bytes(p.Cd, P);
    const uint32_t plane_d = (uint32_t)(p.Cd / 8) * PLANE_STRIDE;
    int tcount = 0;
    const int gdiv = (int)gridDim.x / p.NB, gmod = (int)gridDim.x - gdiv * p.NB;
    int u = (int)blockIdx.x / p.NB, q = (int)blockIdx.x - u * p.NB;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      const bool zero = p.zero_last && u == p.rows_dst - 1;
      const size_t off = ((size_t)u * p.NB + q) * dslab + (size_t)((p.n0 + col0) / 8) * PLANE_STRIDE + (size_t)b * 16;
      uint8_t* drow = p.dst + off;
      uint4 yv[4][P];
      if (p.ymul != nullptr) {
        const uint8_t* yrow = p.ymul + off;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g < ngr) {
            yv[g][0] = __ldg(reinterpret_cast<const uint4*>(yrow + (size_t)g * PLANE_STRIDE));
            if (P == 2) yv[g][P - 1] = __ldg(reinterpret_cast<const uint4*>(yrow + (size_t)g * PLANE_STRIDE + plane_d));
          }
      }
      SC_T0(tw5);
      mbar_wait_parked(smem_u32(&tfull_bar[buf]), (tcount >> 1) & 1, 2000);
      SC_ACC(5, tw5);
      tc_fence_after();
      const bool empty_tile = tile_empty_s[buf] != 0;
      const uint32_t taddr = tmem_base + (uint32_t)(buf * p.NP + col0) + ((uint32_t)(quarter * 32) << 16);
      uint32_t r[4][8];
      SC_T0(te1);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g < ngr) tmem_ld8(taddr + g * 8, r[g]);
      tmem_ld_wait();
      SC_ACC(13, te1);
      SC_T0(te2);
      tc_fence_before();   // accumulator columns of this warp are in registers: hand the buffer back before the math
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      SC_ACC(14, te2);
      SC_T0(te3);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g >= ngr) break;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (empty_tile ? 0.f : __uint_as_float(r[g][i])) + bias_s[col0 + 8 * g + i];
        act_fwd8(v, p.act);
        if (p.ymul != nullptr) {
          float y[8];
          unpack8(yv[g][0], y);
          if (P == 2) {
            float yl[8];
            unpack8(yv[g][P - 1], yl);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] += yl[i];
          }
          act_bwd8(v, y, p.act_mul);
        }
        if (zero) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        if (P == 1) {
          *reinterpret_cast<uint4*>(drow + (size_t)g * PLANE_STRIDE) = pack8(v);
        } else {
          uint4 hi, lo;
          split8(v, hi, lo);
          *reinterpret_cast<uint4*>(drow + (size_t)g * PLANE_STRIDE) = hi;
          *reinterpret_cast<uint4*>(drow + (size_t)g * PLANE_STRIDE + plane_d) = lo;
        }
      }
      SC_ACC(15, te3);
      u += gdiv; q += gmod;
      if (q >= p.NB) { q -= p.NB; ++u; }
    }
  }

#ifdef SHB_SLAB_TRACE
  if (tid == SC_LEAD_WARPS * 32) {
    g_slab_trace[blockIdx.x * 16 + 5] = trace_acc[5];
    g_slab_trace[blockIdx.x * 16 + 6] = clock64() - trace_start;
    g_slab_trace[blockIdx.x * 16 + 7] = (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    g_slab_trace[blockIdx.x * 16 + 13] = trace_acc[13];
    g_slab_trace[blockIdx.x * 16 + 14] = trace_acc[14];
    g_slab_trace[blockIdx.x * 16 + 15] = trace_acc[15];
  }
#endif
  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ weight operand images
// nn.Linear weight (Cout, S*Cin) fp32  ->  bf16 operand images in the un-swizzled K-major core-matrix layout
//   [plane][N/8][Q][8 n][8 k],  Q = K/8:
//   forward : N = pad16(Cout_p), K = S*Cin_p,   img[n][s*Cin_p + c]  = W[n][s*Cin + c]
//   backward: N = pad16(Cin_p),  K = S*Cout_p,  img[n][s*Cout_p + o] = W[o][s*Cin + n]      (per-slot transpose)
// Padded channels are zero.  planes == 2 adds the lo image (w - bf16(w)) behind the hi image.
struct WeightImageJob {
  const float* w;
  uint8_t* img_f;
  uint8_t* img_b;
  int S, Cin, Cout, Cin_p, Cout_p;
  int first_item;   // prefix sum of work items (one item = 8 consecutive k of one image row)
};
constexpr int WI_MAX_JOBS = 24;
struct WeightImageTable {
  WeightImageJob job[WI_MAX_JOBS];
  int count, total_items;
};

// All conv layers of a model in ONE launch (they are tiny: nine launches cost more in launch latency than in work).
__global__ void slab_weight_image_kernel(const WeightImageTable t, int planes) {
  for (int gi = blockIdx.x * blockDim.x + threadIdx.x; gi < t.total_items; gi += gridDim.x * blockDim.x) {
    int ji = 0;
    while (ji + 1 < t.count && gi >= t.job[ji + 1].first_item) ++ji;
    const WeightImageJob& jb = t.job[ji];
    const float* __restrict__ w = jb.w;
    const int S = jb.S, Cin = jb.Cin, Cout = jb.Cout, Cin_p = jb.Cin_p, Cout_p = jb.Cout_p;
    const int i = gi - jb.first_item;
    const int Nf = (Cout_p + 15) / 16 * 16, Qf = S * Cin_p / 8, Nb = (Cin_p + 15) / 16 * 16, Qb = S * Cout_p / 8;
    const int nf = Nf * Qf;
    const bool fwd = i < nf;
    const int c = fwd ? i : i - nf;
    const int Q = fwd ? Qf : Qb, N = fwd ? Nf : Nb;
    const int n = c / Q, q = c - n * Q;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = q * 8 + e;
      float x = 0.f;
      if (fwd) {
        const int s = k / Cin_p, ci = k - s * Cin_p;
        if (n < Cout && ci < Cin) x = __ldg(w + (size_t)n * S * Cin + s * Cin + ci);
      } else {
        const int s = k / Cout_p, o = k - s * Cout_p;
        if (n < Cin && o < Cout) x = __ldg(w + (size_t)o * S * Cin + s * Cin + n);
      }
      v[e] = x;
    }
    uint8_t* img = fwd ? jb.img_f : jb.img_b;
    const size_t off = (((size_t)(n >> 3) * Q + q) * 8 + (n & 7)) * 16;
    if (planes == 1) {
      *reinterpret_cast<uint4*>(img + off) = pack8(v);
    } else {
      uint4 hi, lo;
      split8(v, hi, lo);
      *reinterpret_cast<uint4*>(img + off) = hi;
      *reinterpret_cast<uint4*>(img + (size_t)N * Q * 16 + off) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
constexpr size_t SC_SMEM_MAX = 227 * 1024 - 3072;  // dynamic budget: leave room for the static part (barriers, tables, bias)
constexpr size_t SC_SMEM_DUAL = 108 * 1024;         // per CTA when two share an SM (static part and the per-CTA reserve on top)

static inline int pad16(int c) { return (c + 15) / 16 * 16; }

struct SlabConvPlan { int NP, nstage, SPS; size_t smem; };

// Columns per pass (all of them unless the weight image would not leave room for a 3-stage ring), ring depth, smem bytes.
static bool slab_conv_plan(int S, int CS, int NPt, int P, SlabConvPlan* out, size_t budget = SC_SMEM_MAX) {
  const int Q = S * CS / 8;
  const size_t slab_b = (size_t)P * CS * 256;
  const size_t zero_b = CS == 8 ? PLANE_STRIDE : 0;
  constexpr size_t stage_cap = 64 * 1024;
  for (int NP = NPt > 256 ? 256 : NPt; NP >= 16; NP -= 16) {
    if (NPt % NP != 0 && NP != NPt) continue;  // equal passes
    const size_t img_region = (((size_t)P * NP * Q * 16 + (CS == 8 ? 128 : 0) + 1023) / 1024) * 1024;
    if (img_region + zero_b + 2 * slab_b > budget) continue;
    const size_t room = budget - img_region - zero_b;
    // slabs per stage: as many as keep a stage <= stage_cap (one barrier round trip per stage) with >= 3 stages in the ring
    int SPS = 1;
    while (SPS < SC_MAX_SPS && slab_b * SPS * 2 <= stage_cap && room / (slab_b * SPS * 2) >= 3) SPS *= 2;
    const size_t stage_b = slab_b * SPS;
    size_t n = room / stage_b;
    if (n < 3 && NP > 16) continue;  // prefer a narrower pass with a deeper ring
    if (n > SC_MAX_STAGES) n = SC_MAX_STAGES;
    out->NP = NP; out->nstage = (int)n; out->SPS = SPS; out->smem = img_region + zero_b + n * stage_b;
    return true;
  }
  return false;
}

template <int P, int NK, bool DUAL> static int slab_conv_go(const SlabConvParams& p, size_t smem, cudaStream_t st) {
  static bool attr_set = false;  // per instantiation; the attribute is sticky
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(slab_conv_kernel<P, NK, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(DUAL ? SC_SMEM_DUAL : SC_SMEM_MAX));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(slab_conv_kernel<P, NK, DUAL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int slots = persistent_sms() * (DUAL ? 2 : 1);
  const int grid = p.num_tiles < slots ? p.num_tiles : slots;
  const int threads = SC_LEAD_WARPS * 32 + 128 * p.EG;
  slab_conv_kernel<P, NK, DUAL><<<grid, threads, smem, st>>>(p);
  SHB_LAUNCH_CHECK();
  return 0;
}

static int slab_conv_launch(const SlabConvParams& p, size_t smem, bool dual, cudaStream_t st) {
  const int nk = p.CS >= 16 ? p.CS / 16 : 1;  // MMAs (K = 16) per slab and plane
#define SHB_SC_GO(PL, K) return dual ? slab_conv_go<PL, K, true>(p, smem, st) : slab_conv_go<PL, K, false>(p, smem, st)
  if (p.P == 1) {
    switch (nk) {
      case 1: SHB_SC_GO(1, 1);
      case 2: SHB_SC_GO(1, 2);
      case 4: SHB_SC_GO(1, 4);
      case 8: SHB_SC_GO(1, 8);
      case 16: SHB_SC_GO(1, 16);
      default: return SHB_E_UNSUPPORTED;
    }
  } else {
    switch (nk) {
      case 1: SHB_SC_GO(2, 1);
      case 2: SHB_SC_GO(2, 2);
      case 4: SHB_SC_GO(2, 4);
      case 8: SHB_SC_GO(2, 8);
      default: return SHB_E_UNSUPPORTED;
    }
  }
#undef SHB_SC_GO
}

}  // namespace shb

using namespace shb;

#ifdef SHB_SLAB_TRACE
extern "C" int shb_slab_trace_read(long long* host_out) {  // debug builds only: not part of the ABI
  return (int)cudaMemcpyFromSymbol(host_out, g_slab_trace, sizeof(long long) * kNumSMs * 16);
}
#endif

extern "C" {

size_t shb_slab_tensor_bytes(int rows, int B, int C, int planes) {
  if (rows <= 0 || B <= 0 || C <= 0 || (C & 7) || planes < 1 || planes > 2) return 0;
  return slab::tensor_bytes(rows, B, C, planes);
}

size_t shb_slab_weight_image_bytes(int S, int Ck, int Cn, int planes) {
  if (S <= 0 || Ck <= 0 || Cn <= 0 || (Ck & 7) || (Cn & 7) || planes < 1 || planes > 2) return 0;
  return (size_t)planes * pad16(Cn) * (S * Ck / 8) * 16;
}

int shb_slab_weight_images_batch(int count, const float* const* w, void* const* img_fwd, void* const* img_bwd, const int* S,
                                 const int* Cin, const int* Cout, const int* Cin_p, const int* Cout_p, int planes, void* stream) {
  if (count <= 0 || !w || !img_fwd || !img_bwd || !S || !Cin || !Cout || !Cin_p || !Cout_p || planes < 1 || planes > 2)
    return SHB_E_ARG;
  for (int base = 0; base < count; base += WI_MAX_JOBS) {
    WeightImageTable t{};
    t.count = count - base < WI_MAX_JOBS ? count - base : WI_MAX_JOBS;
    int items = 0;
    for (int i = 0; i < t.count; ++i) {
      const int k = base + i;
      if (!w[k] || !img_fwd[k] || !img_bwd[k] || S[k] <= 0 || S[k] > 32 || Cin[k] <= 0 || Cout[k] <= 0 || Cin_p[k] < Cin[k] ||
          Cout_p[k] < Cout[k] || (Cin_p[k] & 7) || (Cout_p[k] & 7))
        return SHB_E_ARG;
      t.job[i] = WeightImageJob{w[k], (uint8_t*)img_fwd[k], (uint8_t*)img_bwd[k], S[k], Cin[k], Cout[k], Cin_p[k], Cout_p[k], items};
      items += pad16(Cout_p[k]) * (S[k] * Cin_p[k] / 8) + pad16(Cin_p[k]) * (S[k] * Cout_p[k] / 8);
    }
    t.total_items = items;
    int grid = ceil_div(items, 256);
    if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
    slab_weight_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, planes);
    SHB_LAUNCH_CHECK();
  }
  return 0;
}

int shb_slab_weight_images(const float* w, void* img_fwd, void* img_bwd, int S, int Cin, int Cout, int Cin_p, int Cout_p, int planes,
                           void* stream) {
  return shb_slab_weight_images_batch(1, &w, &img_fwd, &img_bwd, &S, &Cin, &Cout, &Cin_p, &Cout_p, planes, stream);
}

int shb_slab_conv_supported(int S, int Cs, int Cd, int planes) {
  SlabConvPlan plan;
  if (S <= 0 || S > 32 || planes < 1 || planes > 2) return 0;
  if (!(Cs == 8 || Cs == 16 || Cs == 32 || Cs == 64 || Cs == 128 || (Cs == 256 && planes == 1)) || (Cd & 7) || Cd <= 0 ||
      Cd > 1024)
    return 0;
  return slab_conv_plan(S, Cs, pad16(Cd), planes, &plan) ? 1 : 0;
}

int shb_slab_conv(const void* src, const int32_t* ptr, const int32_t* entries, const void* w_img, const float* bias, void* dst,
                  const void* ymul, int B, int rows_dst, int S, int Cs, int Cd, int Cd_real, int act, int act_mul, int zero_last,
                  int planes, void* stream) {
  if (!src || !ptr || !entries || !w_img || !dst || B <= 0 || rows_dst <= 0) return SHB_E_ARG;
  if (!shb_slab_conv_supported(S, Cs, Cd, planes)) return SHB_E_UNSUPPORTED;
  SlabConvPlan plan;
  const int NPt = pad16(Cd);
  slab_conv_plan(S, Cs, NPt, planes, &plan);
  // two CTAs per SM when the whole weight image and a >= 3-stage ring fit in half an SM and two epilogue groups cover the
  // accumulator (<= 64 columns)
  SlabConvPlan dplan;
  const bool dual = Cd <= 64 && slab_conv_plan(S, Cs, NPt, planes, &dplan, SC_SMEM_DUAL) && dplan.NP == NPt &&
                    dplan.nstage >= 3;
  if (dual) plan = dplan;
  SlabConvParams p{};
  p.src = (const uint8_t*)src; p.ptr = ptr; p.ent = entries; p.dst = (uint8_t*)dst; p.ymul = (const uint8_t*)ymul;
  p.NB = slab::num_chunks(B); p.rows_dst = rows_dst;
  p.CS = Cs; p.Q = S * Cs / 8; p.Cd = Cd;
  p.act = act; p.act_mul = act_mul; p.zero_last = zero_last; p.P = planes;
  p.nstage = plan.nstage; p.SPS = plan.SPS;
  p.num_tiles = rows_dst * p.NB;
  p.img_plane_stride = (uint32_t)NPt * p.Q * 16;
  for (int n0 = 0; n0 < NPt; n0 += plan.NP) {
    if (n0 >= Cd) break;
    p.NP = plan.NP; p.n0 = n0;
    p.ncols = Cd - n0 < plan.NP ? Cd - n0 : plan.NP;
    p.w_img = (const uint8_t*)w_img + (size_t)(n0 / 8) * p.Q * 128;
    p.cpw = p.ncols >= 128 ? 32 : (p.ncols >= 64 ? 16 : 8);
    if (dual) p.cpw = p.ncols >= 64 ? 32 : (p.ncols >= 32 ? 16 : 8);
    p.EG = p.ncols / p.cpw;
    while (p.EG > 4 && p.cpw < 32) { p.cpw *= 2; p.EG /= 2; }   // at most four epilogue groups (measured best)
    p.bias = bias ? bias + n0 : nullptr;
    p.nbias = Cd_real - n0 < p.ncols ? (Cd_real - n0 > 0 ? Cd_real - n0 : 0) : p.ncols;  // bias holds Cd_real entries
    uint32_t cols = 32;
    while (cols < 2u * p.NP) cols <<= 1;
    p.tmem_cols = cols;
    int rc = slab_conv_launch(p, plan.smem, dual, (cudaStream_t)stream);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // extern "C"
