// Orientation-adaptive pairwise-distance loss, fused:
//   angle weights      utils_SH.py:442-478    (angle_skl: angle between v_i - v_j and the part's bone, |cos| clamped, NaN -> 1)
//   distance matrices  utils_distance.py:366-376 (calc_euclidean_dist_matrix) of ground truth and reconstruction
//   loss               train_funcs.py:243-284 (== :353-389): per part, over the entries with w*De != 0 and i != j,
//                      mean |w*De_r/De - w| (relat_flag) or mean |w*De_r - w*De|, summed with per-part weights.
// The reference materialises (B, n, n, 3) twice plus four (B, n, n) matrices per part and loops over 17 parts in Python; here
// nothing of size n x n ever exists: a thread owns one vertex i of one part of one sample and walks the part's vertices j
// through shared-memory tiles.  Distances are formed from coordinate differences (the reference's |x|^2 - 2x.y + |y|^2
// loses ~3 digits for neighbouring vertices in fp32; against the float64 evaluation of the reference formula this path is
// within 1e-4, the reference's own fp32 evaluation is at 1e-3).  All reductions run in a fixed order; no atomics.
#include "shb_common.cuh"

namespace shb {

constexpr int PL_THREADS = 128;

struct PLParams {
  const float* tx;      // (B, V, 3) ground truth
  const float* rec;     // (B, V, 3) reconstruction
  const float* kps;     // (B, NK, 3)
  const int32_t* idx;   // concatenated part vertex ids
  const int32_t* gptr;  // (G+1)
  const int32_t* bone;  // (G, 3): keypoint ids, third = -1 when the bone has two ends
  const int32_t* wmode; // (G): 0 all-one, 1 linear, 2 sin, 3 threshold
  const float* scale;   // (B, G) or null: per-sample scale of the ground-truth distances
  float w_threshold;
  int relative;
  int B, V, NK, G, nblk;
};

// One UNORDERED pair {i, j}: w, De and the mask are symmetric in (i, j), so are |e| and the magnitude of the gradient term; the
// matrix of the reference holds both (i, j) and (j, i).  Returns false when the entry is outside the mask ((w * De) == 0).
// Square roots and divisions are formed from two rsqrt (relative error ~1e-7, against the 1e-4 bar of this loss):
//   |d| = dd * rsqrt(dd),   cos = |d . k| * rsqrt(dd) / |k|,   Der / De = (rr * rsqrt(rr)) * rsqrt(dd) / sc.
// abs_e = |e_ij|; q = sign(e) (w [/ De]) / Der, the factor of (r_i - r_j) in d|e|/d r_i (0 when the reconstructed points
// coincide: the reference's sqrt'(0) gives NaN; 0 here).
// acos on [0, 1]: sqrt(1 - x) * P7(x) (Abramowitz & Stegun 4.4.46, |error| <= 2e-8 before rounding) -- the library acosf is
// ~25 instructions of the ~120 a pair costs.  acos(1) == 0 exactly (the mask (w * De) != 0 depends on it).
__device__ __forceinline__ float pl_acos01(float x) {
  float p = -0.0012624911f;
  p = fmaf(p, x, 0.0066700901f);
  p = fmaf(p, x, -0.0170881256f);
  p = fmaf(p, x, 0.0308918810f);
  p = fmaf(p, x, -0.0501743046f);
  p = fmaf(p, x, 0.0889789874f);
  p = fmaf(p, x, -0.2145988016f);
  p = fmaf(p, x, 1.5707963050f);
  const float om = 1.f - x;
  return om > 0.f ? om * rsqrtf(om) * p : 0.f;
}

__device__ __forceinline__ bool pl_pair(const float3 vi, const float3 vj, const float3 ri, const float3 rj, const float3 kd,
                                        float inv_km, int mode, float thr, float sc, float inv_sc, int relative, float& abs_e,
                                        float& q, float& rx, float& ry, float& rz) {
  const float dx = vi.x - vj.x, dy = vi.y - vj.y, dz = vi.z - vj.z;
  const float dd = dx * dx + dy * dy + dz * dz;
  if (dd == 0.f) return false;   // De == 0 (with any weight): w * De == 0
  const float inv_dm = rsqrtf(dd);
  const float De = dd * inv_dm * sc;
  float w;
  if (mode == 0) {
    w = 1.f;
  } else {
    float c = fabsf(dx * kd.x + dy * kd.y + dz * kd.z) * inv_dm * inv_km;
    if (c != c) c = 1.f;  // NaN (zero-length bone) -> 1, utils_SH.py:462
    c = fminf(fmaxf(c, 0.f), 1.f);
    if (mode == 2) {
      w = sqrtf(fmaxf(1.f - c * c, 0.f));   // sin(acos(c)): the reference's sin(angle / 180 * pi) of the angle in degrees
    } else {
      w = pl_acos01(c) * (2.f / 3.14159265358979323846f);   // angle / 90 with the angle in degrees
      if (mode == 3 && w < thr) w = 0.f;
    }
  }
  if ((w * De) == 0.f) return false;
  rx = ri.x - rj.x; ry = ri.y - rj.y; rz = ri.z - rj.z;
  const float rr = rx * rx + ry * ry + rz * rz;
  const float inv_Der = rr > 0.f ? rsqrtf(rr) : 0.f;
  const float Der = rr * inv_Der;
  const float inv_De = inv_dm * inv_sc;
  const float e = relative ? (w * Der * inv_De - w) : (w * Der - w * De);
  abs_e = fabsf(e);
  const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
  q = sg * (relative ? w * inv_De : w) * inv_Der;
  return true;
}

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }

// block-uniform per-(b, part) constants: bone direction and the reciprocal of its length
__device__ __forceinline__ void pl_bone(const PLParams& p, int b, int k, float3& kd, float& inv_km) {
  const float* kb = p.kps + (size_t)b * p.NK * 3;
  const int k0 = p.bone[k * 3], k1 = p.bone[k * 3 + 1], k2 = p.bone[k * 3 + 2];
  const float3 a = ld3(kb + k0 * 3), c1 = ld3(kb + k1 * 3);
  if (k2 < 0) {
    kd = make_float3(a.x - c1.x, a.y - c1.y, a.z - c1.z);
  } else {
    const float3 c2 = ld3(kb + k2 * 3);
    kd = make_float3(a.x - (c1.x + c2.x) / 2, a.y - (c1.y + c2.y) / 2, a.z - (c1.z + c2.z) / 2);
  }
  inv_km = 1.f / sqrtf(kd.x * kd.x + kd.y * kd.y + kd.z * kd.z);   // zero-length bone: inf -> cos is inf or NaN -> 1
}

// Gradient accumulator: per (sample, part) one slot of PL_THREADS x 3 floats for every tile pair (J, I), I <= J, at
// J (J + 1) / 2 + I.  Slot (T, T) holds what block T accumulated for its own vertices, slot (J, I < J) what block I accumulated
// for the vertices of tile J; the gradient of a vertex of tile T is the sum of slots (T, 0..T), added in that order.
__host__ __device__ inline int pl_tri(int nblk) { return nblk * (nblk + 1) / 2; }

// grid (nblk, G, B): block (I, k, b) owns tile I (128 vertices of part k of sample b) and walks the tiles J >= I, so every
// unordered pair is evaluated ONCE (the pair arithmetic -- rsqrt, acos -- is the whole cost of this loss): the diagonal tile in
// 64 rotations j = (t + s) mod 128, s = 1..64 (s = 64 by the lower half only), the others in 128 rotations.  A pair adds
// 2 |e| and 2 to the block's (sum of terms, number of masked entries) -> partials[((b*G + k)*nblk + I)*2 + {0,1}].
// With gacc != null the same pass leaves the UNSCALED gradient sum_j sign(e_ij) (w_ij [/ De_ij]) (r_i - r_j) / |r_i - r_j|:
// +q (r_i - r_j) to vertex i in registers, -q (r_i - r_j) to vertex j in a per-warp shared-memory array (the lanes of a warp
// hold 32 different j in every rotation: no conflicts, no atomics, fixed order), written out per tile pair.
__global__ void __launch_bounds__(PL_THREADS) pl_partial_kernel(const PLParams p, float* __restrict__ partials,
                                                                float* __restrict__ gacc) {
  __shared__ float3 sv[PL_THREADS], sr[PL_THREADS];
  __shared__ float sg[PL_THREADS / 32][3][PL_THREADS];
  __shared__ float red[PL_THREADS / 32];
  const int I = blockIdx.x, k = blockIdx.y, b = blockIdx.z, t = threadIdx.x, warp = t >> 5;
  const int g0 = p.gptr[k], n = p.gptr[k + 1] - g0;
  float* out = partials + (((size_t)b * p.G + k) * p.nblk + I) * 2;
  if (I * PL_THREADS >= n) {  // block-uniform
    if (t == 0) { out[0] = 0.f; out[1] = 0.f; }
    return;
  }
  float3 kd; float inv_km;
  pl_bone(p, b, k, kd, inv_km);
  const int mode = p.wmode[k];
  const float sc = p.scale ? p.scale[(size_t)b * p.G + k] : 1.f;
  const float inv_sc = 1.f / sc;
  const float* txb = p.tx + (size_t)b * p.V * 3;
  const float* rcb = p.rec + (size_t)b * p.V * 3;
  const int i = I * PL_THREADS + t;
  const bool on = i < n;
  const int vi_id = on ? p.idx[g0 + i] : 0;
  const float3 vi = ld3(txb + (size_t)vi_id * 3), ri = ld3(rcb + (size_t)vi_id * 3);
  const bool want_g = gacc != nullptr;
  float* gslots = want_g ? gacc + ((size_t)b * p.G + k) * pl_tri(p.nblk) * PL_THREADS * 3 : nullptr;
  float sum = 0.f, cnt = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
  for (int J = I; J * PL_THREADS < n; ++J) {
    __syncthreads();   // the previous tile's readers are done with sv / sr / sg
    const int j0 = J * PL_THREADS;
    if (j0 + t < n) {
      const int vj = p.idx[g0 + j0 + t];
      sv[t] = ld3(txb + (size_t)vj * 3);
      sr[t] = ld3(rcb + (size_t)vj * 3);
    }
#pragma unroll
    for (int w = 0; w < PL_THREADS / 32; ++w)
#pragma unroll
      for (int c = 0; c < 3; ++c) sg[w][c][t] = 0.f;
    __syncthreads();
    const int jn = min(PL_THREADS, n - j0);
    const bool diag = J == I;
    const int s0 = diag ? 1 : 0, s1 = diag ? PL_THREADS / 2 : PL_THREADS - 1;
    for (int s = s0; s <= s1; ++s) {
      __syncwarp();   // the lanes of a warp stay on the same rotation: their 32 targets in sg are distinct
      const int jj = (t + s) & (PL_THREADS - 1);
      if (!on || jj >= jn || (diag && s == PL_THREADS / 2 && t >= PL_THREADS / 2)) continue;
      float abs_e, q, rx, ry, rz;
      if (!pl_pair(vi, sv[jj], ri, sr[jj], kd, inv_km, mode, p.w_threshold, sc, inv_sc, p.relative, abs_e, q, rx, ry, rz)) continue;
      sum += 2.f * abs_e;
      cnt += 2.f;
      if (want_g) {
        gx = fmaf(q, rx, gx);
        gy = fmaf(q, ry, gy);
        gz = fmaf(q, rz, gz);
        sg[warp][0][jj] = fmaf(-q, rx, sg[warp][0][jj]);
        sg[warp][1][jj] = fmaf(-q, ry, sg[warp][1][jj]);
        sg[warp][2][jj] = fmaf(-q, rz, sg[warp][2][jj]);
      }
    }
    __syncthreads();
    if (want_g) {
      float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
      for (int w = 0; w < PL_THREADS / 32; ++w) { ax += sg[w][0][t]; ay += sg[w][1][t]; az += sg[w][2][t]; }
      if (diag) {   // this thread's own vertex as the j of its tile mates
        gx += ax; gy += ay; gz += az;
      } else if (t < jn) {
        float* gp = gslots + ((size_t)(J * (J + 1) / 2 + I) * PL_THREADS + t) * 3;
        gp[0] = ax; gp[1] = ay; gp[2] = az;
      }
    }
  }
  if (want_g && on) {
    float* gp = gslots + ((size_t)(I * (I + 1) / 2 + I) * PL_THREADS + t) * 3;
    gp[0] = gx; gp[1] = gy; gp[2] = gz;
  }
  __syncthreads();
  sum = block_sum<PL_THREADS>(sum, red);
  __syncthreads();
  cnt = block_sum<PL_THREADS>(cnt, red);
  if (t == 0) { out[0] = sum; out[1] = cnt; }
}

// one block: loss = sum_k pw[k] * S_k / N_k;  coef[k] = pw[k] / N_k (what the backward scales by)
__global__ void __launch_bounds__(256) pl_final_kernel(const float* __restrict__ partials, const float* __restrict__ pw,
                                                       int per_part /* entries per (b): nblk */, int B, int G,
                                                       float* __restrict__ coef, float* __restrict__ loss_out) {
  __shared__ float red[8];
  float loss = 0.f;
  for (int k = 0; k < G; ++k) {
    float s = 0.f, c = 0.f;
    for (int q = threadIdx.x; q < B * per_part; q += 256) {
      const int b = q / per_part, blk = q - b * per_part;
      const float* e = partials + (((size_t)b * G + k) * per_part + blk) * 2;
      s += e[0];
      c += e[1];
    }
    __syncthreads();
    s = block_sum<256>(s, red);
    __syncthreads();
    c = block_sum<256>(c, red);
    if (threadIdx.x == 0) {
      coef[k] = pw[k] / c;          // an empty mask gives inf/NaN, as F.l1_loss of an empty selection does
      loss += pw[k] * (s / c);
    }
  }
  if (threadIdx.x == 0) *loss_out = loss;
}

// grid (nblk, G, B): d loss / d rec = 2 coef[k] gscale (sum of the vertex's slots) for the vertices of every part (vertices
// outside every part: zeroed by the caller).  The matrix holds both (i,j) and (j,i), and w, De and the mask are symmetric:
// factor two.
__global__ void __launch_bounds__(PL_THREADS) pl_bwd_kernel(const PLParams p, const float* __restrict__ coef,
                                                            const float* __restrict__ gscale, const float* __restrict__ gacc,
                                                            float* __restrict__ grec) {
  const int T = blockIdx.x, k = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int g0 = p.gptr[k], n = p.gptr[k + 1] - g0;
  const int i = T * PL_THREADS + t;
  if (i >= n) return;
  const int vi_id = p.idx[g0 + i];
  const float c = 2.f * coef[k] * __ldg(gscale);
  const float* gs = gacc + (((size_t)b * p.G + k) * pl_tri(p.nblk) + (size_t)T * (T + 1) / 2) * PL_THREADS * 3 + (size_t)t * 3;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int I = 0; I <= T; ++I) {
    gx += gs[(size_t)I * PL_THREADS * 3];
    gy += gs[(size_t)I * PL_THREADS * 3 + 1];
    gz += gs[(size_t)I * PL_THREADS * 3 + 2];
  }
  const size_t o = ((size_t)b * p.V + vi_id) * 3;
  grec[o] = c * gx;
  grec[o + 1] = c * gy;
  grec[o + 2] = c * gz;
}

}  // namespace shb

using namespace shb;

extern "C" {

static size_t pl_partials_floats(int B, int G, int nblk) { return (size_t)B * G * nblk * 2; }

size_t shb_pair_loss_grad_acc_bytes(int B, int G, int max_part_rows) {
  if (B <= 0 || G <= 0 || max_part_rows <= 0) return 0;
  const int nblk = (max_part_rows + PL_THREADS - 1) / PL_THREADS;
  return (size_t)B * G * pl_tri(nblk) * PL_THREADS * 3 * sizeof(float);
}

size_t shb_pair_loss_workspace(int B, int G, int max_part_rows) {
  if (B <= 0 || G <= 0 || max_part_rows <= 0) return 0;
  const int nblk = (max_part_rows + PL_THREADS - 1) / PL_THREADS;
  return (pl_partials_floats(B, G, nblk) + (size_t)G) * sizeof(float);
}

static int pl_fill(PLParams& p, const float* tx, const float* rec, const float* kps, const int32_t* idx, const int32_t* gptr,
                   const int32_t* bone, const int32_t* wmode, const float* scale, float w_threshold, int relative, int B,
                   int V, int NK, int G, int max_part_rows) {
  if (!tx || !rec || !kps || !idx || !gptr || !bone || !wmode) return SHB_E_ARG;
  if (B <= 0 || V <= 0 || NK <= 0 || G <= 0 || max_part_rows <= 0) return SHB_E_ARG;
  if (B > 65535 || G > 65535) return SHB_E_SHAPE;
  p.tx = tx; p.rec = rec; p.kps = kps; p.idx = idx; p.gptr = gptr; p.bone = bone; p.wmode = wmode; p.scale = scale;
  p.w_threshold = w_threshold; p.relative = relative;
  p.B = B; p.V = V; p.NK = NK; p.G = G; p.nblk = (max_part_rows + PL_THREADS - 1) / PL_THREADS;
  return 0;
}

int shb_pair_loss_fwd(const float* tx, const float* rec, const float* kps, const int32_t* idx, const int32_t* gptr,
                      const int32_t* bone, const int32_t* wmode, const float* part_weight, const float* scale,
                      float w_threshold, int relative, float* loss_out, float* grad_acc, void* workspace, size_t workspace_bytes,
                      int B, int V, int NK, int G, int max_part_rows, void* stream) {
  PLParams p{};
  int rc = pl_fill(p, tx, rec, kps, idx, gptr, bone, wmode, scale, w_threshold, relative, B, V, NK, G, max_part_rows);
  if (rc) return rc;
  if (!part_weight || !loss_out || !workspace) return SHB_E_ARG;
  if (workspace_bytes < shb_pair_loss_workspace(B, G, max_part_rows)) return SHB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* partials = (float*)workspace;
  float* coef = partials + pl_partials_floats(B, G, p.nblk);
  pl_partial_kernel<<<dim3(p.nblk, G, B), PL_THREADS, 0, st>>>(p, partials, grad_acc);
  SHB_LAUNCH_CHECK();
  pl_final_kernel<<<1, 256, 0, st>>>(partials, part_weight, p.nblk, B, G, coef, loss_out);
  SHB_LAUNCH_CHECK();
  return 0;
}

int shb_pair_loss_bwd(const float* grad_acc, const int32_t* idx, const int32_t* gptr, const float* gscale, float* grec,
                      const void* workspace, size_t workspace_bytes, int B, int V, int G, int max_part_rows, void* stream) {
  if (!grad_acc || !idx || !gptr || !gscale || !grec || !workspace) return SHB_E_ARG;
  if (B <= 0 || V <= 0 || G <= 0 || max_part_rows <= 0) return SHB_E_ARG;
  if (B > 65535 || G > 65535) return SHB_E_SHAPE;
  PLParams p{};
  p.idx = idx; p.gptr = gptr; p.B = B; p.V = V; p.G = G; p.nblk = (max_part_rows + PL_THREADS - 1) / PL_THREADS;
  if (workspace_bytes < shb_pair_loss_workspace(B, G, max_part_rows)) return SHB_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const float* coef = (const float*)workspace + pl_partials_floats(B, G, p.nblk);
  cudaError_t e = cudaMemsetAsync(grec, 0, (size_t)B * V * 3 * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  pl_bwd_kernel<<<dim3(p.nblk, G, B), PL_THREADS, 0, st>>>(p, coef, gscale, grad_acc, grec);
  SHB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
