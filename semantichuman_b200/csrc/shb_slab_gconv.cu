// SpiralConv forward and input gradient in the slab layout with SHARED-SOURCE GROUPS: the second-generation conv kernel.
//
// shb_slab_conv.cu runs one destination row per tile and pulls its S neighbour slabs out of L2 -- every activation slab
// crosses the L2 -> SM fabric ~S times (11 live entries per row at level 0), which is where the level-0/1 layers sit
// (ncu: ~7400 B/cycle through the LTS at 42 % "LTS throughput", DRAM 16 %, tensor pipe 8 %).  Here a tile is a GROUP of up to R
// destination rows (clustered on the host so that their entry lists share sources, shb_build_conv_groups) x one 128-sample
// chunk: every distinct source slab of the group is loaded ONCE by TMA, and each (destination, slot) pair that uses it is one
// MMA group  slab x Wop[slot]  into that destination's own TMEM accumulator (R accumulators of NP columns, double-buffered:
// 2 R NP <= 512 columns).  Loads per destination row drop from ~11 to ~4.9 (R = 8) / ~3.7 (R = 16) at level 0; the sums keep a
// fixed order (ascending source row), fp32, no atomics.
//
//   forward (models.py:34-53):   y[j]  = mask * act( sum_s  x[table[j,s]] . W_s^T + b )
//   input gradient            :  gx[u] = act'(y_prev[u]) * sum_{(j,s): table[j,s]=u} gz[j] . W_s
//
// Roles: warps 0-1 = TMA producers (alternate records), warps 2-3 = MMA issuers (one thread each: the pairs of even / odd
// destinations -- independent accumulators, so no ordering between the two is needed, and the issue rate of the dependent
// per-pair instruction chain doubles), warps 4.. = epilogue
// (TMEM -> bias / act / act' / mask -> HBM; work items = (destination, 8-column vector), dealt round-robin to the warp groups,
// act' operands prefetched four items ahead).
#include "shb_common.cuh"
#include "shb_internal.h"
#include "shb_slab.cuh"

namespace shb {

using namespace umma;
using namespace slab;

constexpr int GC_LEAD_WARPS = 4;        // two producers, two MMA issuers
constexpr int GC_MAX_THREADS = GC_LEAD_WARPS * 32 + 4 * 128;
constexpr int GC_DUAL_THREADS = GC_LEAD_WARPS * 32 + 2 * 128;
constexpr int GC_MAX_STAGES = 12;
constexpr int GC_MAX_SPS = 8;     // sources per record
constexpr int GC_MAX_PAIRS = 32;  // (destination, slot) pairs per record
constexpr int GC_REC_WORDS = 48;
// act' operand prefetch depth of the epilogue (work items): requested BEFORE the wait on the accumulators, so that the
// ~2000-cycle latency of a global load issued next to a saturated TMA ring hides under the tile's MMA phase
template <int P> struct GcPf { static constexpr int value = P == 1 ? 8 : 4; };

struct SlabGConvParams {
  const uint8_t* src;      // slab tensor, CS channels, P planes
  const int32_t* gptr;     // (n_groups + 1) record ranges
  const int32_t* recs;     // GC_REC_WORDS words per record (include/shb200.h: shb_build_conv_groups)
  const int32_t* gdst;     // (n_groups * R) destination rows, -1 = absent
  const uint32_t* gmask;   // bit i: destination i of the group has no entries (its accumulator is never written)
  const uint8_t* w_img;    // weight operand image(s) of this pass
  const float* bias;
  uint8_t* dst;
  const uint8_t* ymul;
  uint32_t img_plane_stride;
  int NB, rows_dst, R;
  int CS, Q, NP;
  int n0, ncols, Cd, nbias;
  int act, act_mul, zero_last;
  int nstage, SPS, EG;
  int num_tiles;
  uint32_t tmem_cols;
};

// ------------------------------------------------------------------------------------------------ epilogue
// Work items of a tile: (destination i, 8-column vector v), n = i * nvec + v; warp group e takes items e, e + EG, ...
// (a warp can only read its own TMEM lane quarter, warp_id % 4, so a group is four warps = 128 samples).
// ACT / MUL: activation applied to the accumulator / activation whose derivative (through ymul) multiplies it; -1 = run-time
// value from the parameters, MUL == 0 = no ymul operand at all.
template <int P, int ACT, int MUL>
__device__ __forceinline__ void gc_epilogue(const SlabGConvParams& p, uint32_t tmem_base, const float* bias_s, int32_t* my_rows,
                                            uint64_t* tfull_bar, uint64_t* tempty_bar, int gdiv, int gmod, long long* trace_acc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const int grp = (warp - GC_LEAD_WARPS) >> 2;
  const int b = quarter * 32 + lane;  // sample within the chunk
  const int nvec = p.ncols >> 3;      // power of two
  const int vshift = 31 - __clz(nvec);
  const int nitems = p.R * nvec;
  const size_t dslab = slab_bytes(p.Cd, P);
  const uint32_t plane_d = (uint32_t)(p.Cd / 8) * PLANE_STRIDE;
  constexpr int GC_PF = GcPf<P>::value;
  int tcount = 0;
  int g = (int)blockIdx.x / p.NB, q = (int)blockIdx.x - g * p.NB;
  for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
    const int buf = tcount & 1;
    // The tile's destination rows go through shared memory, not through shuffles of the loaded register: a register that came
    // from a global load keeps its scoreboard, which the act' prefetches below re-arm -- every later shuffle would wait for
    // ALL outstanding prefetches (measured: ~2000 cycles per work item, the epilogue as the kernel's bottleneck).
    __syncwarp();
    if (lane <= 16) {
      int w = -1;
      if (lane < p.R) w = __ldg(p.gdst + (size_t)g * p.R + lane);
      if (lane == 16) w = (int)__ldg(p.gmask + g);
      my_rows[lane] = w;
    }
    __syncwarp();
    const uint32_t untouched = (uint32_t)my_rows[16];
    const size_t qoff = (size_t)q * dslab + (size_t)(p.n0 / 8) * PLANE_STRIDE + (size_t)b * 16;
    uint4 yv[GC_PF][P];
    // item n -> byte offset of its 16-byte vector in dst / ymul; row < 0 when the destination is absent
    auto item_off = [&](int n, int& row) -> size_t {
      const int i = n >> vshift, v = n & (nvec - 1);
      row = my_rows[i];
      return (size_t)row * p.NB * dslab + qoff + (size_t)v * PLANE_STRIDE;
    };
    if (MUL != 0) {
#pragma unroll
      for (int k = 0; k < GC_PF; ++k) {
        const int n = grp + k * p.EG;
        if (n < nitems) {
          int row;
          const size_t off = item_off(n, row);
          if (row >= 0) {
            yv[k][0] = __ldg(reinterpret_cast<const uint4*>(p.ymul + off));
            if (P == 2) yv[k][P - 1] = __ldg(reinterpret_cast<const uint4*>(p.ymul + off + plane_d));
          }
        }
      }
    }
#ifdef SHB_GCONV_TRACE
    const long long tw5 = clock64();
#endif
    mbar_wait_parked(smem_u32(&tfull_bar[buf]), (tcount >> 1) & 1, 2000);
#ifdef SHB_GCONV_TRACE
    trace_acc[5] += clock64() - tw5;
#endif
    tc_fence_after();
    const uint32_t tbase = tmem_base + (uint32_t)(buf * p.R * p.NP) + ((uint32_t)(quarter * 32) << 16);
    for (int n0 = grp; n0 < nitems; n0 += GC_PF * p.EG) {
#pragma unroll
      for (int k = 0; k < GC_PF; ++k) {
        const int n = n0 + k * p.EG;
        if (n < nitems) {
          int row;
          const size_t off = item_off(n, row);
          const int i = n >> vshift, v = n & (nvec - 1);
          if (row >= 0) {
            uint32_t rr[8];
            tmem_ld8(tbase + (uint32_t)(i * p.NP + v * 8), rr);
            tmem_ld_wait();
            const bool empty = (untouched >> i) & 1u;
            float val[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) val[e] = (empty ? 0.f : __uint_as_float(rr[e])) + bias_s[v * 8 + e];
            if (ACT >= 0) act_fwd8_as<ACT>(val); else act_fwd8(val, p.act);
            if (MUL != 0) {
              float y[8];
              unpack8(yv[k][0], y);
              if (P == 2) {
                float yl[8];
                unpack8(yv[k][P - 1], yl);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] += yl[e];
              }
              if (MUL > 0) act_bwd8_as<MUL>(val, y); else act_bwd8(val, y, p.act_mul);
            }
            if (p.zero_last && row == p.rows_dst - 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) val[e] = 0.f;
            }
            uint8_t* d = p.dst + off;
            if (P == 1) {
              *reinterpret_cast<uint4*>(d) = pack8(val);
            } else {
              uint4 hi, lo;
              split8(val, hi, lo);
              *reinterpret_cast<uint4*>(d) = hi;
              *reinterpret_cast<uint4*>(d + plane_d) = lo;
            }
          }
        }
        // refill this prefetch register with the item GC_PF rounds ahead
        if (MUL != 0) {
          const int nn = n + GC_PF * p.EG;
          if (nn < nitems) {
            int row2;
            const size_t off2 = item_off(nn, row2);
            if (row2 >= 0) {
              yv[k][0] = __ldg(reinterpret_cast<const uint4*>(p.ymul + off2));
              if (P == 2) yv[k][P - 1] = __ldg(reinterpret_cast<const uint4*>(p.ymul + off2 + plane_d));
            }
          }
        }
      }
    }
    tc_fence_before();  // every accumulator column this warp owns is in registers / stored: hand the buffer back
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    g += gdiv; q += gmod;
    if (q >= p.NB) { q -= p.NB; ++g; }
  }
}

// Debug timeline (build with SHB_NVCC_FLAGS=-DSHB_GCONV_TRACE): per CTA, cycles each role spends waiting / working.
//   [0] producer 0 wait-empty  [1] producer 0 total  [2] MMA wait-full  [3] MMA wait-accumulator-free  [4] MMA total
//   [5] epilogue wait-accumulator (first epilogue warp)  [6] epilogue total  [7] tiles  [8] MMA issue  [9] records
#ifdef SHB_GCONV_TRACE
__device__ long long g_gconv_trace[2 * kNumSMs * 16];
#define GC_T0(var) const long long var = clock64()
#define GC_ACC(idx, t0) trace_acc[idx] += clock64() - (t0)
#else
#define GC_T0(var) do { } while (0)
#define GC_ACC(idx, t0) do { } while (0)
#endif

template <int P, int NK, bool DUAL>
__global__ void __launch_bounds__(DUAL ? GC_DUAL_THREADS : GC_MAX_THREADS, DUAL ? 2 : 1) slab_gconv_kernel(const SlabGConvParams p) {
#ifdef SHB_GCONV_TRACE
  long long trace_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long trace_start = clock64();
#endif
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[GC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[GC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t img_bar;
  __shared__ __align__(16) unsigned long long src_s[GC_MAX_STAGES * GC_MAX_SPS];  // per slab: global source address
  __shared__ __align__(16) uint4 meta_s[GC_MAX_STAGES * GC_MAX_PAIRS];            // per pair: see the producer
  __shared__ uint32_t hdr_s[GC_MAX_STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[256];
  __shared__ int32_t rows_s[16][20];  // per epilogue warp: the tile's destination rows [0..15] and untouched mask [16]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t img_bytes1 = (uint32_t)p.NP * p.Q * 16;  // one plane
  // 8-channel sources read one core matrix past their slot's (see the MMA issuer): 128 B of zeros behind the image
  const uint32_t img_region = ((P * img_bytes1 + (p.CS == 8 ? 128u : 0u) + 1023) / 1024) * 1024;
  const uint32_t slab_b = (uint32_t)P * p.CS * 256;
  const uint32_t stage_b = slab_b * p.SPS;
  const uint32_t smem0 = smem_u32(dyn_smem);
  const uint32_t ring0 = smem0 + img_region;
  const uint32_t zero0 = ring0 + (uint32_t)p.nstage * stage_b;  // 2 KB of zeros (8-channel sources only)
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
      mbar_init(&empty_bar[i], 2);   // both MMA issuers commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 2);
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
  const int gdiv = (int)gridDim.x / p.NB, gmod = (int)gridDim.x - gdiv * p.NB;  // a tile step adds gridDim.x = gdiv*NB + gmod

  if (warp < 2) {
    // ================================================================ producers: records -> bulk copies + pair words
    // Both warps walk the same sequence of (tile, record); warp w serves the records whose running number is w mod 2.  The
    // record of the NEXT served stage is requested before the current one is processed (a record is 192 B in L2).
    const uint32_t pw = (uint32_t)warp;
    if (warp == 0 && lane == 0) {  // weight operand image(s): resident for the whole kernel
      mbar_expect_tx(smem_u32(&img_bar), P * img_bytes1);
      for (int pl = 0; pl < P; ++pl)
        for (uint32_t off = 0; off < img_bytes1; off += 32768) {
          const uint32_t n = img_bytes1 - off < 32768 ? img_bytes1 - off : 32768;
          bulk_load(smem0 + pl * img_bytes1 + off, p.w_img + (size_t)pl * p.img_plane_stride + off, n, smem_u32(&img_bar));
        }
    }
    const size_t row_stride = (size_t)p.NB * slab_b;
    const uint32_t b_step = (uint32_t)p.CS;  // (CS/8 chunks) * 128 B >> 4 per slot
    const uint32_t lo_b0 = ((smem0 & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);  // B descriptor low word of slot 0 (LBO = 128)
    int tc = 0;  // tiles entered so far: parity = accumulator buffer
    // walk state: tile t = (group g, chunk q), record r of [r, r1)
    int t = blockIdx.x, g = t / p.NB, q = t - g * p.NB;
    int r = 0, r1 = 0;
    if (t < p.num_tiles) { r = __ldg(p.gptr + g); r1 = __ldg(p.gptr + g + 1); }
    uint32_t stage_no = 0, slot = 0, ph = 0;
    // skip to this warp's first record
    auto advance = [&]() {  // one record forward (crossing into the next tile when the group is exhausted)
      ++r;
      ++stage_no;
      if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
      if (r >= r1) {
        ++tc;
        t += gridDim.x; g += gdiv; q += gmod;
        if (q >= p.NB) { q -= p.NB; ++g; }
        if (t < p.num_tiles) { r = __ldg(p.gptr + g); r1 = __ldg(p.gptr + g + 1); }
      }
    };
    if (pw == 1 && t < p.num_tiles) advance();
    int w0 = 0, w1 = 0;
    if (t < p.num_tiles) {
      const int32_t* rec = p.recs + (size_t)r * GC_REC_WORDS;
      w0 = __ldg(rec + lane);
      if (lane < 16) w1 = __ldg(rec + 32 + lane);
    }
    while (t < p.num_tiles) {
      // current record: words in (w0, w1), destination stage (slot, ph), chunk q
      const int c0 = w0, c1 = w1;
      const uint32_t cslot = slot, cph = ph;
      const int cbuf = tc & 1;
      const uint8_t* srcq = p.src + (size_t)q * slab_b;
      // move the walk two records on and request that record
      advance();
      if (t < p.num_tiles) advance();
      if (t < p.num_tiles) {
        const int32_t* rec = p.recs + (size_t)r * GC_REC_WORDS;
        w0 = __ldg(rec + lane);
        if (lane < 16) w1 = __ldg(rec + 32 + lane);
      }
      const uint32_t hdr = (uint32_t)__shfl_sync(0xFFFFFFFFu, c0, 0);
      const int ns = (int)(hdr & 15u), np = (int)((hdr >> 4) & 63u);
      if (elect_one()) {
        GC_T0(tw);
        mbar_wait(empty0 + cslot * 8, cph ^ 1);
        GC_ACC(0, tw);
      }
      __syncwarp();
      if (lane >= 1 && lane <= ns) src_s[cslot * GC_MAX_SPS + lane - 1] = (unsigned long long)(srcq + (size_t)c0 * row_stride);
      {
        // Per pair, ready-made for the MMA threads (whose every instruction is on the critical path, and whose operands must be
        // moved into uniform registers one by one): A-descriptor low word of the pair's slab, B-descriptor low word of its
        // slot, TMEM address of its destination's accumulator, accumulate flag.  Pairs of even destinations are listed from
        // the front of the stage's array (issuer 0), pairs of odd destinations from its back (issuer 1), each in record order.
        const int pidx = lane >= 16 ? lane - 16 : lane + 16;
        const int wv = lane >= 16 ? c0 : c1;
        const bool live = pidx < np;
        const uint32_t dl = ((uint32_t)wv >> 3) & 31u;
        const uint32_t odd_b = __ballot_sync(0xFFFFFFFFu, live && (dl & 1u));
        const uint32_t evn_b = __ballot_sync(0xFFFFFFFFu, live && !(dl & 1u));
        const uint32_t below = (1u << pidx) - 1u;   // ballots are by lane; rotating by 16 puts bit pidx in place
        const uint32_t odd_r = (odd_b >> 16) | (odd_b << 16), evn_r = (evn_b >> 16) | (evn_b << 16);
        if (live) {
          const uint32_t k = (uint32_t)wv & 7u, first = ((uint32_t)wv >> 8) & 1u, sl = ((uint32_t)wv >> 9) & 31u;
          const uint32_t a_k = ring0 + cslot * stage_b + k * slab_b;
          const uint32_t lbo_a = p.CS >= 16 ? (uint32_t)PLANE_STRIDE : zero0 - a_k;
          uint4 m;
          m.x = ((a_k & 0x3FFFFu) >> 4) | ((lbo_a >> 4) << 16);
          m.y = lo_b0 + sl * b_step;
          m.z = tmem_base + (uint32_t)(cbuf * p.R * p.NP) + dl * (uint32_t)p.NP;
          m.w = first ^ 1u;
          const uint32_t pos = (dl & 1u) ? (uint32_t)(GC_MAX_PAIRS - 1) - __popc(odd_r & below) : __popc(evn_r & below);
          meta_s[cslot * GC_MAX_PAIRS + pos] = m;
        }
        if (lane == 0) hdr_s[cslot] = (hdr & ~(63u << 4)) | ((uint32_t)__popc(evn_b) << 4) | ((uint32_t)__popc(odd_b) << 16);
      }
      __syncwarp();
      if (elect_one()) {
        const uint32_t bar = full0 + cslot * 8;
        if (ns > 0) {
          mbar_expect_tx(bar, (uint32_t)ns * slab_b);
          const uint32_t dst = ring0 + cslot * stage_b;
          const unsigned long long* sp = &src_s[cslot * GC_MAX_SPS];
#pragma unroll
          for (int k = 0; k < GC_MAX_SPS; ++k)
            if (k < ns) bulk_load(dst + k * slab_b, reinterpret_cast<const void*>(sp[k]), slab_b, bar);
        } else {
          mbar_arrive(bar);  // a group without entries: the record only carries the first/last flags
        }
      }
#ifdef SHB_GCONV_TRACE
      trace_acc[9] += 1;
#endif
    }
#ifdef SHB_GCONV_TRACE
    if (lane == 0 && warp == 0) {
      g_gconv_trace[blockIdx.x * 16 + 0] = trace_acc[0];
      g_gconv_trace[blockIdx.x * 16 + 1] = clock64() - trace_start;
      g_gconv_trace[blockIdx.x * 16 + 9] = trace_acc[9];
    }
#endif
  } else if (warp < GC_LEAD_WARPS) {
    // ================================================================ MMA issuers (one thread each)
    const uint32_t mw = (uint32_t)warp - 2u;   // 0: even destinations (list grows up from 0), 1: odd (down from the back)
    if (elect_one()) {
      const uint32_t idesc = idesc_bf16_f32(CHUNK, p.NP, 0, 0);
      const uint64_t hi_b = ((uint64_t)(((uint32_t)p.Q * 128) >> 4) << 32) | ((uint64_t)1 << 46);  // SBO = Q*128
      const uint64_t hi_a = ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);                    // SBO = 128
      // A: K-major, SBO = 128 (next 8 samples), LBO = 2048 (next 8 channels); 8-channel sources pair the slab with a block of
      // zeros through LBO = zero0 - slab (the producer builds the word).  The second plane of a slab sits CS*256 bytes on.
      const uint32_t plane_a = p.CS >= 16 ? ((uint32_t)p.CS * 256) >> 4 : (((uint32_t)p.CS * 256) >> 4) - ((((uint32_t)p.CS * 256) >> 4) << 16);
      const uint32_t plane_b = img_bytes1 >> 4;
      mbar_wait(smem_u32(&img_bar), 0);
      uint32_t slot = 0, ph = 0;
      int tcount = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        GC_T0(tw3);
        mbar_wait_parked(smem_u32(&tempty_bar[buf]), ((tcount >> 1) & 1) ^ 1, 1000);
        GC_ACC(3, tw3);
        tc_fence_after();
        for (;;) {
          GC_T0(tw2);
          mbar_wait_parked(full0 + slot * 8, ph, 1000);
          GC_ACC(2, tw2);
          GC_T0(ti);
          const uint32_t hdr = hdr_s[slot];
          tc_fence_after();
          const uint32_t np = mw == 0 ? (hdr >> 4) & 63u : (hdr >> 16) & 63u;
          const uint4* mp = &meta_s[slot * GC_MAX_PAIRS + (mw == 0 ? 0 : GC_MAX_PAIRS - 1)];
          const int mstep = mw == 0 ? 1 : -1;
          uint4 nxt = mp[0];
#pragma unroll 1
          for (uint32_t pi = 0; pi < np; ++pi) {
            const uint4 m = nxt;
            mp += mstep;
            if (pi + 1 < np) nxt = mp[0];  // the next pair's words are on their way while this pair's MMAs issue
            uint32_t la = m.x, lb = m.y;
            mma_bf16(m.z, hi_a | la, hi_b | lb, idesc, m.w);
            if (P == 2) {  // x ~ xh + xl, w ~ wh + wl: xh.wh + xl.wh + xh.wl; xl.wl (<= 2^-18 of the term) is dropped
              mma_bf16(m.z, hi_a | (la + plane_a), hi_b | lb, idesc, 1);
              mma_bf16(m.z, hi_a | la, hi_b | (lb + plane_b), idesc, 1);
            }
#pragma unroll
            for (uint32_t kk = 1; kk < (uint32_t)NK; ++kk) {
              la += 2 * PLANE_STRIDE >> 4;
              lb += 16;
              mma_bf16(m.z, hi_a | la, hi_b | lb, idesc, 1);
              if (P == 2) {
                mma_bf16(m.z, hi_a | (la + plane_a), hi_b | lb, idesc, 1);
                mma_bf16(m.z, hi_a | la, hi_b | (lb + plane_b), idesc, 1);
              }
            }
          }
          mma_commit_u32(empty0 + slot * 8);  // stage reusable once these MMAs have read it
          GC_ACC(8, ti);
          if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
          if (hdr & (1u << 11)) break;
        }
        mma_commit_u32(smem_u32(&tfull_bar[buf]));  // all accumulators of the group complete
      }
#ifdef SHB_GCONV_TRACE
      if (mw == 0) {
      g_gconv_trace[blockIdx.x * 16 + 2] = trace_acc[2];
      g_gconv_trace[blockIdx.x * 16 + 3] = trace_acc[3];
      g_gconv_trace[blockIdx.x * 16 + 4] = clock64() - trace_start;
      g_gconv_trace[blockIdx.x * 16 + 8] = trace_acc[8];
      }
#endif
    }
    __syncwarp();
  } else {
    // ================================================================ epilogue warps
    // The activation is a compile-time constant of the epilogue body for the combinations a training step uses (ELU forward,
    // identity forward, ELU derivative); anything else runs the generic body with run-time switches.  One body, executed 8
    // items deep, must stay small: with the switch inlined per item the loop was 4000 instructions and ran out of the
    // instruction cache (~1000 cycles per 8-column work item).
#ifdef SHB_GCONV_TRACE
#define GC_EPI(A, M) gc_epilogue<P, A, M>(p, tmem_base, bias_s, rows_s[warp - GC_LEAD_WARPS], tfull_bar, tempty_bar, gdiv, gmod, trace_acc)
#else
#define GC_EPI(A, M) gc_epilogue<P, A, M>(p, tmem_base, bias_s, rows_s[warp - GC_LEAD_WARPS], tfull_bar, tempty_bar, gdiv, gmod, nullptr)
#endif
    if (p.ymul == nullptr || p.act_mul == 0) {
      if (p.act == SHB_ACT_ELU) GC_EPI(SHB_ACT_ELU, 0);
      else if (p.act == SHB_ACT_IDENTITY) GC_EPI(SHB_ACT_IDENTITY, 0);
      else GC_EPI(-1, 0);
    } else {
      if (p.act == SHB_ACT_IDENTITY && p.act_mul == SHB_ACT_ELU) GC_EPI(SHB_ACT_IDENTITY, SHB_ACT_ELU);
      else GC_EPI(-1, -1);
    }
#undef GC_EPI
  }

#ifdef SHB_GCONV_TRACE
  if (tid == GC_LEAD_WARPS * 32) {
    g_gconv_trace[blockIdx.x * 16 + 5] = trace_acc[5];
    g_gconv_trace[blockIdx.x * 16 + 6] = clock64() - trace_start;
    g_gconv_trace[blockIdx.x * 16 + 7] = (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
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

// ------------------------------------------------------------------------------------------------ host side
constexpr size_t GC_SMEM_MAX = 227 * 1024 - 10240;   // dynamic budget: the static part (barriers, tables, bias) is ~4 KB
constexpr size_t GC_SMEM_DUAL = 102 * 1024;          // per CTA when two share an SM

static inline int gpad16(int c) { return (c + 15) / 16 * 16; }

struct SlabGConvPlan { int NP, nstage, SPS, R, dual; size_t smem; };

// Ring geometry for a budget: all destination columns in one pass (the grouped kernel is for the layers whose weight image
// leaves room; wider ones stay on shb_slab_conv), slabs per stage, ring depth.
static bool gconv_ring(int S, int CS, int NP, int P, size_t budget, SlabGConvPlan* o) {
  const int Q = S * CS / 8;
  const size_t slab_b = (size_t)P * CS * 256;
  const size_t zero_b = CS == 8 ? PLANE_STRIDE : 0;
  const size_t img_region = (((size_t)P * NP * Q * 16 + (CS == 8 ? 128 : 0) + 1023) / 1024) * 1024;
  if (img_region + zero_b + 3 * slab_b > budget) return false;
  const size_t room = budget - img_region - zero_b;
  int SPS = 1;
  while (SPS < GC_MAX_SPS && slab_b * SPS * 2 <= 32768 && room / (slab_b * SPS * 2) >= 4) SPS *= 2;
  size_t n = room / (slab_b * SPS);
  if (n < 3) return false;
  if (n > GC_MAX_STAGES) n = GC_MAX_STAGES;
  o->NP = NP; o->nstage = (int)n; o->SPS = SPS; o->smem = img_region + zero_b + n * slab_b * SPS;
  return true;
}

static bool slab_gconv_plan(int S, int CS, int Cd, int P, SlabGConvPlan* o) {
  if (S <= 0 || S > 32 || P < 1 || P > 2) return false;
  if (!(CS == 8 || CS == 16 || CS == 32 || CS == 64 || CS == 128) || (Cd & 7) || Cd <= 0 || Cd > 128) return false;
  const int NP = gpad16(Cd);
  SlabGConvPlan d{}, s{};
  const bool can_dual = Cd <= 64 && gconv_ring(S, CS, NP, P, GC_SMEM_DUAL, &d) && 256 / (2 * NP) >= 2;
  const bool can_single = gconv_ring(S, CS, NP, P, GC_SMEM_MAX, &s);
  if (!can_dual && !can_single) return false;
  const bool dual = can_dual;  // measured: two CTAs per SM beat one CTA with twice the group size on every layer that has both
  *o = dual ? d : s;
  o->dual = dual ? 1 : 0;
  int R = (dual ? 256 : 512) / (2 * NP);
  if (R > 16) R = 16;
  if (R < 1) return false;
  o->R = R;
  return true;
}

template <int P, int NK, bool DUAL> static int slab_gconv_go(const SlabGConvParams& p, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(slab_gconv_kernel<P, NK, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(DUAL ? GC_SMEM_DUAL : GC_SMEM_MAX));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(slab_gconv_kernel<P, NK, DUAL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int slots = persistent_sms() * (DUAL ? 2 : 1);
  const int grid = p.num_tiles < slots ? p.num_tiles : slots;
  const int threads = GC_LEAD_WARPS * 32 + 128 * p.EG;
  slab_gconv_kernel<P, NK, DUAL><<<grid, threads, smem, st>>>(p);
  SHB_LAUNCH_CHECK();
  return 0;
}

static int slab_gconv_launch(const SlabGConvParams& p, int planes, size_t smem, bool dual, cudaStream_t st) {
  const int nk = p.CS >= 16 ? p.CS / 16 : 1;
#define SHB_GC_GO(PL, K) return dual ? slab_gconv_go<PL, K, true>(p, smem, st) : slab_gconv_go<PL, K, false>(p, smem, st)
  if (planes == 1) {
    switch (nk) {
      case 1: SHB_GC_GO(1, 1);
      case 2: SHB_GC_GO(1, 2);
      case 4: SHB_GC_GO(1, 4);
      case 8: SHB_GC_GO(1, 8);
      default: return SHB_E_UNSUPPORTED;
    }
  } else {
    switch (nk) {
      case 1: SHB_GC_GO(2, 1);
      case 2: SHB_GC_GO(2, 2);
      case 4: SHB_GC_GO(2, 4);
      case 8: SHB_GC_GO(2, 8);
      default: return SHB_E_UNSUPPORTED;
    }
  }
#undef SHB_GC_GO
}

}  // namespace shb

using namespace shb;

#ifdef SHB_GCONV_TRACE
extern "C" int shb_gconv_trace_read(long long* host_out) {  // debug builds only: not part of the ABI
  return (int)cudaMemcpyFromSymbol(host_out, g_gconv_trace, sizeof(long long) * 2 * kNumSMs * 16);
}
#endif

extern "C" {

int shb_slab_gconv_plan(int S, int Cs, int Cd, int planes, int* R_max, int* SPS) {
  SlabGConvPlan plan;
  if (!R_max || !SPS) return SHB_E_ARG;
  if (!slab_gconv_plan(S, Cs, Cd, planes, &plan)) return SHB_E_UNSUPPORTED;
  *R_max = plan.R;
  *SPS = plan.SPS;
  return 0;
}

int shb_slab_gconv(const void* src, const int32_t* gptr, const int32_t* recs, const int32_t* gdst, const uint32_t* gmask,
                   int n_groups, int R, int SPS, const void* w_img, const float* bias, void* dst, const void* ymul, int B,
                   int rows_dst, int S, int Cs, int Cd, int Cd_real, int act, int act_mul, int zero_last, int planes,
                   void* stream) {
  if (!src || !gptr || !recs || !gdst || !gmask || !w_img || !dst || B <= 0 || rows_dst <= 0 || n_groups <= 0) return SHB_E_ARG;
  SlabGConvPlan plan;
  if (!slab_gconv_plan(S, Cs, Cd, planes, &plan)) return SHB_E_UNSUPPORTED;
  if (R < 1 || R > plan.R || SPS != plan.SPS) return SHB_E_SHAPE;  // the program was built for another ring geometry
  SlabGConvParams p{};
  p.src = (const uint8_t*)src; p.gptr = gptr; p.recs = recs; p.gdst = gdst; p.gmask = gmask;
  p.w_img = (const uint8_t*)w_img; p.dst = (uint8_t*)dst; p.ymul = (const uint8_t*)ymul;
  p.NB = slab::num_chunks(B); p.rows_dst = rows_dst; p.R = R;
  p.CS = Cs; p.Q = S * Cs / 8; p.Cd = Cd; p.NP = plan.NP;
  p.n0 = 0; p.ncols = Cd;
  p.act = act; p.act_mul = act_mul; p.zero_last = zero_last;
  p.nstage = plan.nstage; p.SPS = plan.SPS;
  p.num_tiles = n_groups * p.NB;
  p.img_plane_stride = (uint32_t)plan.NP * p.Q * 16;
  p.bias = bias;
  p.nbias = Cd_real < Cd ? Cd_real : Cd;
  // epilogue groups: as many as there are work items per tile to deal out, up to 4 (2 when two CTAs share an SM)
  const int items = R * (Cd / 8);
  int EG = plan.dual ? 2 : 4;
  while (EG > 1 && items < EG) EG /= 2;
  p.EG = EG;
  uint32_t cols = 32;
  while (cols < 2u * R * plan.NP) cols <<= 1;
  p.tmem_cols = cols;
  return slab_gconv_launch(p, planes, plan.smem, plan.dual != 0, (cudaStream_t)stream);
}

}  // extern "C"
