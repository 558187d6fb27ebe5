// SpiralConv forward / input-gradient on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 operands, fp32 accumulate.
//
//   C[(b,j), n] = epilogue( sum_k A[(b,j), k] * Bop[n, k] )        M = B*rows_dst rows, K = S*CS, N = Cd (padded to 16)
//
//   forward : A[(b,j), s*CS+c] = x[b, table[j,s], c]                      Bop[n,k] = W[n, k]
//   dgrad   : A[(b,u), s*CS+c] = sum_{j in inv(u,s)} gz[b, j, c]          Bop[n, s*CS+c] = W[c, s*Cd + n]   (CS = Cout, Cd = Cin)
//
// Persistent warp-specialised CTAs (one or two per SM), 128-row tiles:
//   * the whole weight operand lives in shared memory for the lifetime of the CTA (un-swizzled K-major core-matrix layout);
//   * 8 producer warps gather 16-byte channel chunks of neighbour rows straight into the K-major core-matrix layout of a
//     multi-stage A ring -- cp.async (LDGSTS) for the forward gather, register gather-sum (fixed order, no atomics) for dgrad;
//   * one thread issues tcgen05.mma (M=128, N=Cd, K=16 per instruction) into a double-buffered TMEM accumulator and
//     releases A stages / publishes accumulators with tcgen05.commit -> mbarrier;
//   * 4 epilogue warps read the accumulator with tcgen05.ld, apply bias + activation + dummy-row mask, and store bf16.
// The gathered (M x K) matrix never exists in HBM: activations are read once from HBM (re-reads hit L2).
#include "shb_common.cuh"
#include "shb_internal.h"
#include "shb_umma.cuh"

namespace shb {

using namespace umma;

constexpr int UG_BM = 128;             // rows per tile == UMMA M
constexpr int UG_KC = 8;               // 16-byte chunks (8 bf16) per row per stage  -> BK = 64
constexpr int UG_STAGE_BYTES = UG_BM * UG_KC * 16;
constexpr int UG_EPI_WARPS = 4, UG_PROD_WARPS = 8;
constexpr int UG_PROD_THREADS = UG_PROD_WARPS * 32;
constexpr int UG_THREADS = (UG_EPI_WARPS + 1 + UG_PROD_WARPS) * 32;
constexpr int UG_MAX_STAGES = 8;

struct UGParams {
  const __nv_bfloat16* src;  // (B, rows_src, CS)
  const int32_t* table;      // forward: (rows_dst, S) source rows.  dgrad: keyptr (rows_dst*S + 1)
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
  int act, zero_last, skip_last;
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int CS, bool SUM>
__global__ void __launch_bounds__(UG_THREADS, 1) umma_gather_gemm_kernel(const UGParams p) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[UG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UG_MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* w_img = dyn_smem;                                   // [NPAD/8][Q][8 rows][16 B]
  const uint32_t w_bytes = (uint32_t)p.NPAD * p.Q * 16;
  uint8_t* a_ring = dyn_smem + ((w_bytes + 127) / 128) * 128;  // nstage x [16 row groups][8 chunks][8 rows][16 B]
  const int K = p.Q * 8;

  // ---------------------------------------------------------------- prologue: weights -> smem (core-matrix layout)
  {
    const int total_chunks = p.NPAD * p.Q;
    for (int i = tid; i < total_chunks; i += UG_THREADS) {
      const int n = i / p.Q, q = i - n * p.Q;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (n < p.Cd) {
        if (!SUM) {
          v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)n * K + (size_t)q * 8));
        } else {
          // Bop[n][s*CS + co] = W[co*(S*Cd) + s*Cd + n]
          const int k0 = q * 8, s = k0 / CS, co0 = k0 - s * CS;
          const unsigned short* wu = reinterpret_cast<const unsigned short*>(p.w);
          const size_t rs = (size_t)p.S * p.Cd;
          uint32_t e[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) e[t] = __ldg(wu + (size_t)(co0 + t) * rs + (size_t)s * p.Cd + n);
          v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        }
      }
      *reinterpret_cast<uint4*>(w_img + ((size_t)(n >> 3) * p.Q + q) * 128 + (n & 7) * 16) = v;
    }
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full_bar[i], UG_PROD_THREADS);
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
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      mbar_wait(&tfull_bar[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const long long m = (long long)tile * UG_BM + warp * 32 + lane;
      const bool valid = m < p.M;
      const int j = valid ? (int)(m % p.rows_dst) : 0;
      const bool zero = p.zero_last && (j == p.rows_dst - 1);
      __nv_bfloat16* out = p.dst + (valid ? m : 0) * p.Cd;
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
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int n = c0 + t;
          const float b = (p.bias != nullptr && n < p.Cd) ? __ldg(p.bias + n) : 0.f;
          v[t] = zero ? 0.f : act_fwd(__uint_as_float(r[t]) + b, p.act);
        }
        if (!valid) continue;
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
      const uint32_t a_base = smem_u32(a_ring), w_base = smem_u32(w_img);
      const uint32_t sbo_b = (uint32_t)p.Q * 128;
      uint32_t it = 0;
      int tcount = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(&tempty_bar[buf], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.NPAD);
        for (int st = 0; st < p.NS; ++st, ++it) {
          const uint32_t slot = it % p.nstage, ph = (it / p.nstage) & 1;
          mbar_wait(&full_bar[slot], ph);
          tc_fence_after();
          const int chunks = min(UG_KC, p.Q - st * UG_KC);
          const uint32_t a_st = a_base + slot * UG_STAGE_BYTES;
          for (int kk = 0; kk < chunks / 2; ++kk) {
            const uint64_t da = smem_desc(a_st + kk * 256, 128, UG_KC * 128);
            const uint64_t db = smem_desc(w_base + (uint32_t)(st * UG_KC + 2 * kk) * 128, 128, sbo_b);
            mma_bf16(tmem_d, da, db, idesc, (st | kk) != 0);
          }
          mma_commit(&empty_bar[slot]);  // A stage reusable once these MMAs have read it
        }
        mma_commit(&tfull_bar[buf]);     // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ================================================================ producers: gather rows into the A ring
    const int pt = tid - (UG_EPI_WARPS + 1) * 32;  // 0..255
    const int r = pt & (UG_BM - 1), half = pt >> 7;
    const uint32_t a_base = smem_u32(a_ring);
    const uint32_t row_off = (uint32_t)(r >> 3) * (UG_KC * 128) + (uint32_t)(r & 7) * 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long m = (long long)tile * UG_BM + r;
      bool valid = m < p.M;
      int j = 0;
      const __nv_bfloat16* srcb = p.src;
      if (valid) {
        const long long b = m / p.rows_dst;
        j = (int)(m - b * p.rows_dst);
        srcb = p.src + b * (long long)p.rows_src * CS;
      }
      if (SUM && p.skip_last && j == p.rows_dst - 1) valid = false;
      const int32_t* trow = p.table + (long long)j * p.S;
      for (int st = 0; st < p.NS; ++st, ++it) {
        const uint32_t slot = it % p.nstage, ph = (it / p.nstage) & 1;
        mbar_wait(&empty_bar[slot], ph ^ 1);
        const uint32_t dst0 = a_base + slot * UG_STAGE_BYTES + row_off;
#pragma unroll
        for (int i = 0; i < UG_KC / 2; ++i) {
          const int kc = half * (UG_KC / 2) + i;
          const int q = st * UG_KC + kc;
          if (q < p.Q) {
            const int k = q * 8, s = k / CS, c = k - s * CS;
            if (!SUM) {
              const int row = valid ? __ldg(trow + s) : 0;
              cp_async16(dst0 + kc * 128, srcb + (size_t)row * CS + c, valid ? 16u : 0u);
            } else {
              float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              if (valid) {
                const int e0 = __ldg(trow + s), e1 = __ldg(trow + s + 1);
                for (int e = e0; e < e1; ++e) {
                  float v[8];
                  Io<__nv_bfloat16>::ld8(srcb + (size_t)__ldg(p.list + e) * CS + c, v);
#pragma unroll
                  for (int t = 0; t < 8; ++t) acc[t] += v[t];
                }
              }
              const uint4 o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                         pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + kc * 128), "r"(o.x), "r"(o.y),
                           "r"(o.z), "r"(o.w)
                           : "memory");
            }
          }
        }
        if (!SUM) {
          cp_async_mbar_arrive_noinc(&full_bar[slot]);
        } else {
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[slot]);
        }
      }
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
static size_t ug_smem_bytes(int NPAD, int Q, int nstage) {
  return ((size_t)NPAD * Q * 16 + 127) / 128 * 128 + (size_t)nstage * UG_STAGE_BYTES;
}
constexpr size_t UG_SMEM_MAX = 227 * 1024 - 1024;  // leave room for the static barriers

static int ug_npad(int Cd) { return ((Cd + 15) / 16) * 16; }

bool umma_gather_gemm_supported(int Cs, int Cd, int S) {
  if (!(Cs == 16 || Cs == 32 || Cs == 64 || Cs == 128)) return false;
  const int NPAD = ug_npad(Cd);
  if (NPAD > 256) return false;
  return ug_smem_bytes(NPAD, S * Cs / 8, 2) <= UG_SMEM_MAX;
}

template <int CS, bool SUM> static int ug_launch(const UGParams& p, size_t smem, int want_per_sm, cudaStream_t st) {
  static bool attr_set = false;  // per instantiation; the attribute is sticky
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_gather_gemm_kernel<CS, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)UG_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // persistent kernel with static tile striding: the grid must be fully co-resident
  int occ = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, umma_gather_gemm_kernel<CS, SUM>, UG_THREADS, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) return SHB_E_SHAPE;
  int grid = kNumSMs * (occ < want_per_sm ? occ : want_per_sm);
  if (grid > p.num_tiles) grid = p.num_tiles;
  umma_gather_gemm_kernel<CS, SUM><<<grid, UG_THREADS, smem, st>>>(p);
  SHB_LAUNCH_CHECK();
  return 0;
}

int umma_gather_gemm(const void* src, const int32_t* table, const int32_t* list, const void* w, const float* bias,
                     void* dst, int B, int rows_src, int rows_dst, int S, int Cs, int Cd, int act, int zero_last,
                     int skip_last, bool sum_mode, cudaStream_t st) {
  UGParams p{};
  p.src = (const __nv_bfloat16*)src; p.table = table; p.list = list; p.w = (const __nv_bfloat16*)w; p.bias = bias;
  p.dst = (__nv_bfloat16*)dst;
  p.M = (long long)B * rows_dst;
  p.rows_src = rows_src; p.rows_dst = rows_dst; p.S = S; p.Cd = Cd; p.NPAD = ug_npad(Cd);
  p.Q = S * Cs / 8;
  p.NS = (p.Q + UG_KC - 1) / UG_KC;
  p.num_tiles = (int)((p.M + UG_BM - 1) / UG_BM);
  p.act = act; p.zero_last = zero_last; p.skip_last = skip_last;
  int nstage = 6;
  while (nstage > 2 && ug_smem_bytes(p.NPAD, p.Q, nstage) > UG_SMEM_MAX) --nstage;
  // two CTAs per SM when both fit with a deep enough ring: more gathers in flight
  int ctas_per_sm = 1;
  if (ug_smem_bytes(p.NPAD, p.Q, 4) * 2 + 4096 <= UG_SMEM_MAX) { ctas_per_sm = 2; nstage = 4; }
  p.nstage = nstage;
  uint32_t cols = 32;
  while (cols < 2u * p.NPAD) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = ug_smem_bytes(p.NPAD, p.Q, nstage);
#define UG_DISPATCH(CSV)                                                                   \
  case CSV:                                                                                \
    return sum_mode ? ug_launch<CSV, true>(p, smem, ctas_per_sm, st) : ug_launch<CSV, false>(p, smem, ctas_per_sm, st);
  switch (Cs) {
    UG_DISPATCH(16)
    UG_DISPATCH(32)
    UG_DISPATCH(64)
    UG_DISPATCH(128)
    default: return SHB_E_SHAPE;
  }
#undef UG_DISPATCH
}

}  // namespace shb
