// SpiralConv weight and bias gradients in the slab layout (shb_slab.cuh), on tcgen05 with K = batch.
//
//   gW[o][s*Cin + c] = sum_{j, b}  x[table[j,s]][b][c] * gz[j][b][o]          gb[o] = sum_{j, b} gz[j][b][o]
//
// A tile is one output vertex j x one 128-sample chunk.  The S source slabs of j and the gz slab arrive by cp.async.bulk
// (TMA) and ARE the MN-major UMMA operands as they land (K = the 128 samples of the chunk, M = 128 consecutive source
// channels = 128/Cin neighbouring slabs, N = output channels).  The whole (S*Cin x Cout) fp32 accumulator lives in TMEM for
// the lifetime of the CTA (one 128-row group per slab group, <= 512 columns); every CTA streams its tiles once and writes one
// partial; a second kernel adds the partials in CTA order (fixed order: bit-reproducible) and un-pads into the nn.Linear
// layout.  The bias gradient is summed from the gz slab in shared memory by the four warps that otherwise only run the final
// TMEM read-out.
//
// Roles: warps 0-1 = TMA producers (alternate ring stages), warp 2 = MMA issuer (one elected thread), warps 3-6 = gz column
// sums + final read-out.
#include <stdlib.h>

#include "shb_common.cuh"
#include "shb_internal.h"
#include "shb_slab.cuh"

namespace shb {

using namespace umma;
using namespace slab;

constexpr int SW_THREADS = 224;  // two producer warps, the MMA warp, four column-sum / read-out warps
constexpr int SW_MAX_STAGES = 8;
constexpr uint32_t SW_GROUP_BYTES = 32768;  // one plane of one stage: 128 channels x 128 samples x 2 B

struct SlabWgradParams {
  const uint8_t* x;      // slab tensor, Cin channels, P planes
  const uint8_t* gz;     // slab tensor, Cout_p channels, P planes
  const int32_t* table;  // (rows_out, S) source rows
  float* partial;        // [grid][G*128][NPt]: this pass writes rows [g0*128, (g0+Gp)*128) x columns [n0, n0+N)
  float* bias_partial;   // [grid][NPt] or null
  int NB, S, Cin, Cout_p, NPt;
  int G;                 // groups in total
  int g0, Gp, n0, N;     // this pass
  int ncols_gz;          // gz channels that exist in [n0, n0+N): min(N, Cout_p - n0)
  int PC, PPS, PPG;      // channels per piece (min(Cin,128)), pieces per slab, pieces per 128-channel group
  int P, nstage, num_tiles;
  int zero_row;          // source row known to be all zero (the dummy vertex behind padded spiral slots), or -1
  uint32_t tmem_cols;
};

// Operand groups of this pass a tile needs: trailing groups whose slots all read the known-zero source row contribute nothing
// and are neither loaded nor multiplied (level 0-1 spirals: 22 % of the 4-slot groups).  `mine` = the tile's table row, one
// entry per lane.  A CTA's FIRST tile always runs every group: it is what zero-initialises the TMEM accumulators.
__device__ __forceinline__ int wgrad_groups(const SlabWgradParams& p, int mine, int lane, bool first) {
  if (first || p.zero_row < 0) return p.Gp;
  const uint32_t real = __ballot_sync(0xFFFFFFFFu, lane < p.S && mine != p.zero_row);
  const int slots = 32 - __clz(real);                         // 1 + last slot with a live source (0: none)
  int g = (slots * p.PPS + p.PPG - 1) / p.PPG - p.g0;         // groups of the whole operand, minus the passes before this one
  return g < 0 ? 0 : (g > p.Gp ? p.Gp : g);
}

// DUAL: two CTAs per SM (half the ring each) -- the sibling fills the issue slots the dependent role chains leave idle.
template <int P, bool DUAL>
__global__ void __launch_bounds__(SW_THREADS, DUAL ? 2 : 1) slab_wgrad_kernel(const SlabWgradParams p) {
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[SW_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[SW_MAX_STAGES];
  __shared__ __align__(8) uint64_t gfull_bar[2];
  __shared__ __align__(8) uint64_t gempty_bar[2];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_b = P * SW_GROUP_BYTES;
  const uint32_t gz_plane_b = (uint32_t)p.N * 256;                 // this pass's channel slice of one plane
  const uint32_t gz_load_b = (uint32_t)p.ncols_gz * 256;           // bytes that exist (the rest of the operand is never read back)
  const uint32_t gz_buf_b = ((P * gz_plane_b + 1023) / 1024) * 1024;
  const uint32_t smem0 = smem_u32(dyn_smem);
  const uint32_t gz0 = smem0 + (uint32_t)p.nstage * stage_b;
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t gfull0 = smem_u32(gfull_bar), gempty0 = smem_u32(gempty_bar);
  const bool want_bias = p.bias_partial != nullptr;

  if (tid == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&gfull_bar[i], 1);
      mbar_init(&gempty_bar[i], want_bias ? 5 : 1);
    }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int pieces_total = p.S * p.PPS;
  const uint32_t piece_b = (uint32_t)p.PC * 256;
  const size_t xslab = slab_bytes(p.Cin, P), gslab = slab_bytes(p.Cout_p, P);

  if (warp < 2) {
    // ================================================================ producers (warp w: stages with running number w mod 2)
    const uint32_t pw = (uint32_t)warp;
    uint32_t stage_no = 0;
    // Lane L owns piece L of a stage (<= 16 pieces of 128/PPG channels): it picks its source row out of the tile's table row
    // (held one entry per lane, fetched a tile ahead) and issues its own bulk copies; lane 0 does the barrier round trips.
    uint32_t slot = 0, ph = 0;
    int tcount = 0;
    const int pps_shift = p.PPS == 2 ? 1 : 0;
    int t = blockIdx.x;
    int mine = 0;
    if (t < p.num_tiles && lane < p.S) mine = __ldg(p.table + (size_t)(t / p.NB) * p.S + lane);
    for (; t < p.num_tiles; t += gridDim.x, ++tcount) {
      const int j = t / p.NB, q = t - j * p.NB;
      const int tn = t + gridDim.x;
      int mine_n = 0;
      if (tn < p.num_tiles && lane < p.S) mine_n = __ldg(p.table + (size_t)(tn / p.NB) * p.S + lane);
      const int buf = tcount & 1;
      if (lane == 0 && warp == 0) {
        mbar_wait(gempty0 + buf * 8, ((tcount >> 1) & 1) ^ 1);
        mbar_expect_tx(gfull0 + buf * 8, P * gz_load_b);
        const uint8_t* g = p.gz + ((size_t)j * p.NB + q) * gslab + (size_t)(p.n0 / 8) * PLANE_STRIDE;
        for (int pl = 0; pl < P; ++pl)
          bulk_load(gz0 + buf * gz_buf_b + pl * gz_plane_b, g + (size_t)pl * p.Cout_p * 256, gz_load_b, gfull0 + buf * 8);
      }
      const int ng = wgrad_groups(p, mine, lane, tcount == 0);
      for (int gi = 0; gi < ng; ++gi) {
        if (((stage_no++) & 1u) != pw) {  // the other producer warp's stage
          if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
          continue;
        }
        const int pi0 = (p.g0 + gi) * p.PPG;
        const int np = pieces_total - pi0 < p.PPG ? pieces_total - pi0 : p.PPG;
        const int pi = pi0 + lane, s = pi >> pps_shift, h = pi & (p.PPS - 1);
        const int row = __shfl_sync(0xFFFFFFFFu, mine, s & 31);
        if (lane == 0) {
          mbar_wait(empty0 + slot * 8, ph ^ 1);
          mbar_expect_tx(full0 + slot * 8, (uint32_t)np * P * piece_b);
        }
        __syncwarp();
        if (lane < np) {
          const uint8_t* src = p.x + ((size_t)row * p.NB + q) * xslab + (size_t)h * SW_GROUP_BYTES;
          for (int pl = 0; pl < P; ++pl)
            bulk_load(smem0 + slot * stage_b + pl * SW_GROUP_BYTES + lane * piece_b, src + (size_t)pl * p.Cin * 256, piece_b,
                      full0 + slot * 8);
        }
        if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
      }
      mine = mine_n;
    }
  } else if (warp == 2) {
    // ================================================================ MMA issuer (one elected thread: warp-uniform operands;
    // the whole warp reads the tile's table row, a tile ahead, to know how many operand groups the producers send)
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16_f32(CHUNK, p.N, 1, 1);
    // MN-major, un-swizzled: LBO = 128 B (next 8 samples), SBO = 2048 B (next 8 channels)
    const uint64_t hi = ((uint64_t)((uint32_t)PLANE_STRIDE >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t lbo = (128u >> 4) << 16;
    uint32_t slot = 0, ph = 0;
    int tcount = 0;
    int mine = 0;
    if ((int)blockIdx.x < p.num_tiles && lane < p.S && p.zero_row >= 0)
      mine = __ldg(p.table + (size_t)(blockIdx.x / p.NB) * p.S + lane);
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
      const int tn = t + gridDim.x;
      int mine_n = 0;
      if (tn < p.num_tiles && lane < p.S && p.zero_row >= 0) mine_n = __ldg(p.table + (size_t)(tn / p.NB) * p.S + lane);
      const int ng = wgrad_groups(p, mine, lane, tcount == 0);
      if (leader) {
        const int buf = tcount & 1;
        mbar_wait_parked(gfull0 + buf * 8, (tcount >> 1) & 1, 1000);
        const uint32_t lo_g = (((gz0 + buf * gz_buf_b) & 0x3FFFFu) >> 4) | lbo;
        for (int gi = 0; gi < ng; ++gi) {
          mbar_wait_parked(full0 + slot * 8, ph, 1000);
          tc_fence_after();
          const uint32_t lo_a = (((smem0 + slot * stage_b) & 0x3FFFFu) >> 4) | lbo;
          const uint32_t tmem_d = tmem_base + (uint32_t)(gi * p.N);
#pragma unroll
          for (uint32_t kk = 0; kk < 8; ++kk) {  // 16 samples per MMA: 2 x 128 B
            const uint32_t la = lo_a + kk * 16, lg = lo_g + kk * 16;
            mma_bf16(tmem_d, hi | la, hi | lg, idesc, (tcount > 0 || kk > 0) ? 1u : 0u);
            if (P == 2) {  // xh.gh + xl.gh + xh.gl; xl.gl (<= 2^-18 of the term) is dropped
              mma_bf16(tmem_d, hi | (la + (SW_GROUP_BYTES >> 4)), hi | lg, idesc, 1);
              mma_bf16(tmem_d, hi | la, hi | (lg + (gz_plane_b >> 4)), idesc, 1);
            }
          }
          mma_commit_u32(empty0 + slot * 8);
          if (++slot == (uint32_t)p.nstage) { slot = 0; ph ^= 1; }
        }
        mma_commit_u32(gempty0 + buf * 8);  // gz buffer reusable once the MMAs above have read it
      }
      mine = mine_n;
      __syncwarp();
    }
    if (leader) mma_commit_u32(smem_u32(&done_bar));
    __syncwarp();
  } else {
    // ================================================================ warps 3-6: bias column sums, then the read-out
    const int ew = warp - 3;          // 0..3: column-sum work split; TMEM quarter is warp & 3
    float bsum[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bsum[i] = 0.f;
    if (want_bias) {
      int tcount = 0;
      const int nchunk = p.ncols_gz / 8;  // 8-channel chunks of gz in this pass
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait_parked(gfull0 + buf * 8, (tcount >> 1) & 1, 2000);
        const uint8_t* g = dyn_smem + (size_t)p.nstage * stage_b + (size_t)buf * gz_buf_b;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int cc = ew + 4 * ci;
          if (cc < nchunk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float v[8];
              unpack8(*reinterpret_cast<const uint4*>(g + (size_t)cc * PLANE_STRIDE + (size_t)(lane + 32 * i) * 16), v);
              if (P == 2) {
                float l[8];
                unpack8(*reinterpret_cast<const uint4*>(g + gz_plane_b + (size_t)cc * PLANE_STRIDE + (size_t)(lane + 32 * i) * 16), l);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] += l[e];
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) bsum[ci * 8 + e] += v[e];
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gempty0 + buf * 8);
      }
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int cc = ew + 4 * ci;
        if (cc < nchunk) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float s = warp_sum(bsum[ci * 8 + e]);
            if (lane == 0) p.bias_partial[(size_t)blockIdx.x * p.NPt + p.n0 + cc * 8 + e] = s;
          }
        }
      }
    }
    // final read-out: TMEM (Gp groups x N columns, 128 rows) -> this CTA's partial
    mbar_wait_sleep(smem_u32(&done_bar), 0, 200);
    tc_fence_after();
    const int quarter = warp & 3, m = quarter * 32 + lane;
    float* out = p.partial + (size_t)blockIdx.x * p.G * 128 * p.NPt;
    for (int gi = 0; gi < p.Gp; ++gi) {
      float* orow = out + ((size_t)(p.g0 + gi) * 128 + m) * p.NPt + p.n0;
      for (int c0 = 0; c0 < p.N; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + (uint32_t)(gi * p.N + c0) + ((uint32_t)(quarter * 32) << 16), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(orow + c0 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                  __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// gw[o][s*Cin + c] = sum over CTAs of partial[cta][row(s,c)][o];  gb likewise.  Fixed order, hence bit-reproducible: a block
// owns 32 consecutive outputs (o fastest: the 32 loads of a warp are one or two contiguous runs); warp w of its NW warps adds the
// partials q = w, w + NW, ... in ascending order, eight loads in flight, and warp 0 adds the NW sums in warp order.
// (First version: one thread per output walking all 148-296 partials -- 28 blocks for the level-0 layer, 18 us per launch,
// 8 % of the step for 0.2 % of its bytes.)
// Layers with few outputs (level 0-1: 700-13 000) give few blocks, and each of their warps then walks 37 partials in five
// dependent rounds: those run with 32 warps per block (one or two rounds) -- NW is the number of warps sharing a block's partials.
template <int NW>
__global__ void __launch_bounds__(NW * 32) slab_wgrad_reduce_kernel(const float* __restrict__ partial,
                                                                    const float* __restrict__ bias_partial, int nparts, int G,
                                                                    int NPt, int S, int Cin, int Cin_p, int Cout, int PC, int PPS,
                                                                    int PPG, float* __restrict__ gw, float* __restrict__ gb) {
  __shared__ float part_s[NW][32];
  const int K = S * Cin;
  const int total = K * Cout;
  const size_t pstride = (size_t)G * 128 * NPt;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nblk = (total + Cout + 31) / 32;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int i = blk * 32 + lane;
    const float* src = nullptr;
    size_t stride = 0;
    float* out = nullptr;
    if (i < total) {
      if (gw != nullptr) {
        const int k = i / Cout, o = i - k * Cout;
        const int s = k / Cin, c = k - s * Cin;
        const int piece = s * PPS + c / PC;
        const int row = (piece / PPG) * 128 + (piece % PPG) * PC + c % PC;
        src = partial + (size_t)row * NPt + o;
        stride = pstride;
        out = gw + (size_t)o * K + k;
      }
    } else if (i < total + Cout && gb != nullptr && bias_partial != nullptr) {
      src = bias_partial + (i - total);
      stride = (size_t)NPt;
      out = gb + (i - total);
    }
    float acc = 0.f;
    if (src != nullptr) {
      int q = w;
      for (; q + 7 * NW < nparts; q += 8 * NW) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (size_t)(q + j * NW) * stride);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
      }
      if (q + 3 * NW < nparts) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(src + (size_t)(q + j * NW) * stride);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j];
        q += 4 * NW;
      }
      for (; q < nparts; q += NW) acc += __ldg(src + (size_t)q * stride);
    }
    __syncthreads();   // previous round's readers are done with part_s
    part_s[w][lane] = acc;
    __syncthreads();
    if (w == 0 && out != nullptr) {
      float r = 0.f;
#pragma unroll
      for (int j = 0; j < NW; ++j) r += part_s[j][lane];
      *out = r;
    }
  }
  (void)Cin_p;
}

// ------------------------------------------------------------------------------------------------ host side
constexpr size_t SW_SMEM_MAX = 227 * 1024 - 2048;

static inline int pad16w(int c) { return (c + 15) / 16 * 16; }

struct SlabWgradPlan { int PC, PPS, PPG, G, N, Gp, nstage; size_t smem; };

constexpr size_t SW_SMEM_DUAL = 108 * 1024;

static bool slab_wgrad_plan(int S, int Cin, int Cout_p, int P, SlabWgradPlan* o, size_t budget = SW_SMEM_MAX, int tmem_cols = 512) {
  if (!(Cin == 8 || Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256)) return false;
  o->PC = Cin < 128 ? Cin : 128;
  o->PPS = Cin / o->PC;
  o->PPG = 128 / o->PC;
  o->G = (S * o->PPS + o->PPG - 1) / o->PPG;
  const int NPt = pad16w(Cout_p);
  // columns per pass: all of them unless two gz buffers would not leave room for a 2-stage ring, or TMEM overflows
  for (int N = NPt > 128 ? 128 : NPt; N >= 16; N -= 16) {  // <= 128: the column-sum warps keep 4 chunks each
    if (NPt % N != 0) continue;
    const size_t gz_buf = (((size_t)P * N * 256 + 1023) / 1024) * 1024;
    const size_t stage = (size_t)P * SW_GROUP_BYTES;
    if (2 * gz_buf + 2 * stage > budget) continue;
    size_t n = (budget - 2 * gz_buf) / stage;
    if (n > SW_MAX_STAGES) n = SW_MAX_STAGES;
    int Gp = tmem_cols / N;
    if (Gp > o->G) Gp = o->G;
    if (Gp < 1) continue;
    o->N = N; o->Gp = Gp; o->nstage = (int)n; o->smem = 2 * gz_buf + n * stage;
    return true;
  }
  return false;
}

}  // namespace shb

using namespace shb;

extern "C" {

int shb_slab_wgrad_supported(int S, int Cin, int Cout_p, int planes) {
  SlabWgradPlan plan;
  if (S <= 0 || S > 32 || planes < 1 || planes > 2 || Cout_p <= 0 || (Cout_p & 7) || Cout_p > 1024) return 0;
  return slab_wgrad_plan(S, Cin, Cout_p, planes, &plan) ? 1 : 0;
}

size_t shb_slab_wgrad_workspace(int S, int Cin, int Cout_p, int planes) {
  SlabWgradPlan plan;
  if (!shb_slab_wgrad_supported(S, Cin, Cout_p, planes)) return 0;
  slab_wgrad_plan(S, Cin, Cout_p, planes, &plan);
  const size_t NPt = pad16w(Cout_p);
  return (size_t)2 * kNumSMs * ((size_t)plan.G * 128 * NPt + NPt) * sizeof(float);  // up to two CTAs per SM, one partial each
}

/* x: slab tensor (rows_in, B, Cin_p); gz: slab tensor (rows_out, B, Cout_p); table (rows_out, S).
 * gw (Cout, S*Cin) fp32, gb (Cout) fp32 or null.  skip_last: the last output row's gz is zero by construction (mask).
 * zero_src_row: a row of x known to be all zero (the dummy vertex behind padded spiral slots), or -1: trailing operand groups
 * that read nothing else are skipped. */
int shb_slab_wgrad(const void* x, const int32_t* table, const void* gz, float* gw, float* gb, void* workspace,
                   size_t workspace_bytes, int B, int rows_out, int S, int Cin, int Cin_p, int Cout, int Cout_p, int skip_last,
                   int zero_src_row, int planes, void* stream) {
  if (!x || !table || !gz || !workspace || B <= 0 || rows_out <= 0 || Cin <= 0 || Cin > Cin_p || Cout <= 0 || Cout > Cout_p)
    return SHB_E_ARG;
  if (!shb_slab_wgrad_supported(S, Cin_p, Cout_p, planes)) return SHB_E_UNSUPPORTED;
  if (workspace_bytes < shb_slab_wgrad_workspace(S, Cin_p, Cout_p, planes)) return SHB_E_WORKSPACE;
  SlabWgradPlan plan;
  slab_wgrad_plan(S, Cin_p, Cout_p, planes, &plan);
  cudaStream_t st = (cudaStream_t)stream;
  // two CTAs per SM when a >= 3-stage ring fits in half an SM and the whole accumulator in half of TMEM, in one pass
  SlabWgradPlan dplan;
  // (only where a CTA has many tiles: every CTA costs one partial in the fixed-order reduction)
  const long long tiles_all = (long long)(skip_last ? rows_out - 1 : rows_out) * slab::num_chunks(B);
  const bool dual = tiles_all >= 40LL * kNumSMs &&
                    slab_wgrad_plan(S, Cin_p, Cout_p, planes, &dplan, SW_SMEM_DUAL, 256) && dplan.nstage >= 3 &&
                    dplan.Gp == dplan.G && dplan.N == pad16w(Cout_p);
  if (dual) plan = dplan;
  static bool attr_set[3][2] = {{false, false}, {false, false}, {false, false}};
  if (!attr_set[planes][dual]) {
    cudaError_t e;
    if (planes == 1)
      e = dual ? cudaFuncSetAttribute(slab_wgrad_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_DUAL)
               : cudaFuncSetAttribute(slab_wgrad_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_MAX);
    else
      e = dual ? cudaFuncSetAttribute(slab_wgrad_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_DUAL)
               : cudaFuncSetAttribute(slab_wgrad_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    attr_set[planes][dual] = true;
  }
  const int NPt = pad16w(Cout_p);
  SlabWgradParams p{};
  p.x = (const uint8_t*)x; p.gz = (const uint8_t*)gz; p.table = table;
  p.NB = slab::num_chunks(B); p.S = S; p.Cin = Cin_p; p.Cout_p = Cout_p; p.NPt = NPt;
  p.G = plan.G; p.PC = plan.PC; p.PPS = plan.PPS; p.PPG = plan.PPG;
  p.P = planes; p.nstage = plan.nstage;
  p.zero_row = zero_src_row >= 0 ? zero_src_row : -1;
  const int rows_eff = skip_last ? rows_out - 1 : rows_out;
  p.num_tiles = rows_eff * p.NB;
  float* partial = (float*)workspace;
  float* bias_partial = partial + (size_t)2 * kNumSMs * plan.G * 128 * NPt;
  const int slots = persistent_sms() * (dual ? 2 : 1);
  const int grid = p.num_tiles < slots ? p.num_tiles : slots;
  if (grid > 0) {
    for (int n0 = 0; n0 < NPt; n0 += plan.N) {
      if (n0 >= Cout_p) break;
      for (int g0 = 0; g0 < plan.G; g0 += plan.Gp) {
        p.n0 = n0; p.N = plan.N; p.g0 = g0;
        p.ncols_gz = Cout_p - n0 < plan.N ? Cout_p - n0 : plan.N;
        p.Gp = plan.G - g0 < plan.Gp ? plan.G - g0 : plan.Gp;
        p.partial = partial;
        p.bias_partial = (gb != nullptr && g0 == 0) ? bias_partial : nullptr;
        uint32_t cols = 32;
        while (cols < (uint32_t)(p.Gp * p.N)) cols <<= 1;
        p.tmem_cols = cols;
        if (planes == 1) {
          if (dual) slab_wgrad_kernel<1, true><<<grid, SW_THREADS, plan.smem, st>>>(p);
          else slab_wgrad_kernel<1, false><<<grid, SW_THREADS, plan.smem, st>>>(p);
        } else {
          if (dual) slab_wgrad_kernel<2, true><<<grid, SW_THREADS, plan.smem, st>>>(p);
          else slab_wgrad_kernel<2, false><<<grid, SW_THREADS, plan.smem, st>>>(p);
        }
        SHB_LAUNCH_CHECK();
      }
    }
  }
  const int total = S * Cin * Cout + Cout;
  int rgrid = ceil_div(total, 32);
  if (rgrid > 8 * kNumSMs) rgrid = 8 * kNumSMs;
  if (rgrid < 2 * kNumSMs)
    slab_wgrad_reduce_kernel<32><<<rgrid, 32 * 32, 0, st>>>(partial, gb ? bias_partial : nullptr, grid, plan.G, NPt, S, Cin, Cin_p, Cout,
                                                            plan.PC, plan.PPS, plan.PPG, gw, gb);
  else
    slab_wgrad_reduce_kernel<8><<<rgrid, 8 * 32, 0, st>>>(partial, gb ? bias_partial : nullptr, grid, plan.G, NPt, S, Cin, Cin_p, Cout,
                                                          plan.PC, plan.PPS, plan.PPG, gw, gb);
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
