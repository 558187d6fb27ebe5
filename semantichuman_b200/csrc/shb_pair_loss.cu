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

// weight and masked term ingredients of one ordered pair; returns false when the entry is outside the mask
__device__ __forceinline__ bool pl_pair(const float3 vi, const float3 vj, const float3 kd, float km, int mode, float thr,
                                        float sc, float& w, float& De) {
  const float dx = vi.x - vj.x, dy = vi.y - vj.y, dz = vi.z - vj.z;
  const float dd = dx * dx + dy * dy + dz * dz;
  const float dm = sqrtf(dd);
  De = dm * sc;
  if (mode == 0) {
    w = 1.f;
  } else {
    float c = fabsf((dx * kd.x + dy * kd.y + dz * kd.z) / (dm * km));
    if (c != c) c = 1.f;  // NaN (coincident points) -> 1, utils_SH.py:462
    c = fminf(fmaxf(c, 0.f), 1.f);
    const float ang = acosf(c) * (180.f / 3.14159265358979323846f);
    if (mode == 2) {
      w = sinf(ang / 180.f * 3.14159265358979323846f);
    } else {
      w = ang / 90.f;
      if (mode == 3 && w < thr) w = 0.f;
    }
  }
  return (w * De) != 0.f;
}

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }

// block-uniform per-(b, part) constants: bone direction and its length
__device__ __forceinline__ void pl_bone(const PLParams& p, int b, int k, float3& kd, float& km) {
  const float* kb = p.kps + (size_t)b * p.NK * 3;
  const int k0 = p.bone[k * 3], k1 = p.bone[k * 3 + 1], k2 = p.bone[k * 3 + 2];
  const float3 a = ld3(kb + k0 * 3), c1 = ld3(kb + k1 * 3);
  if (k2 < 0) {
    kd = make_float3(a.x - c1.x, a.y - c1.y, a.z - c1.z);
  } else {
    const float3 c2 = ld3(kb + k2 * 3);
    kd = make_float3(a.x - (c1.x + c2.x) / 2, a.y - (c1.y + c2.y) / 2, a.z - (c1.z + c2.z) / 2);
  }
  km = sqrtf(kd.x * kd.x + kd.y * kd.y + kd.z * kd.z);
}

// grid (nblk, G, B): per-block (sum of terms, number of masked entries) -> partials[((b*G + k)*nblk + blk)*2 + {0,1}]
// With gacc != null the same pass also leaves the UNSCALED gradient sum_j sign(e_ij) (w_ij [/ De_ij]) (r_i - r_j) / |r_i - r_j|
// of every part vertex in gacc (B, V, 3): the backward is then an elementwise scaling by 2 coef[k] gscale instead of a second
// walk over all pairs (the pair arithmetic -- sqrt, division, acos -- is the whole cost of this loss).
__global__ void __launch_bounds__(PL_THREADS) pl_partial_kernel(const PLParams p, float* __restrict__ partials,
                                                                float* __restrict__ gacc) {
  __shared__ float3 sv[PL_THREADS], sr[PL_THREADS];
  __shared__ float red[PL_THREADS / 32];
  const int blk = blockIdx.x, k = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int g0 = p.gptr[k], n = p.gptr[k + 1] - g0;
  float* out = partials + (((size_t)b * p.G + k) * p.nblk + blk) * 2;
  if (blk * PL_THREADS >= n) {  // block-uniform
    if (t == 0) { out[0] = 0.f; out[1] = 0.f; }
    return;
  }
  float3 kd; float km;
  pl_bone(p, b, k, kd, km);
  const int mode = p.wmode[k];
  const float sc = p.scale ? p.scale[(size_t)b * p.G + k] : 1.f;
  const float* txb = p.tx + (size_t)b * p.V * 3;
  const float* rcb = p.rec + (size_t)b * p.V * 3;
  const int i = blk * PL_THREADS + t;
  const bool on = i < n;
  const int vi_id = on ? p.idx[g0 + i] : 0;
  const float3 vi = ld3(txb + (size_t)vi_id * 3), ri = ld3(rcb + (size_t)vi_id * 3);
  float sum = 0.f, cnt = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
  for (int j0 = 0; j0 < n; j0 += PL_THREADS) {
    __syncthreads();
    if (j0 + t < n) {
      const int vj = p.idx[g0 + j0 + t];
      sv[t] = ld3(txb + (size_t)vj * 3);
      sr[t] = ld3(rcb + (size_t)vj * 3);
    }
    __syncthreads();
    const int jn = min(PL_THREADS, n - j0);
    if (on) {
      for (int jj = 0; jj < jn; ++jj) {
        if (j0 + jj == i) continue;  // diagonal of w is zeroed, train_funcs.py:268-269
        float w, De;
        if (!pl_pair(vi, sv[jj], kd, km, mode, p.w_threshold, sc, w, De)) continue;
        const float rx = ri.x - sr[jj].x, ry = ri.y - sr[jj].y, rz = ri.z - sr[jj].z;
        const float Der = sqrtf(rx * rx + ry * ry + rz * rz);
        const float e = p.relative ? (w * Der / De - w) : (w * Der - w * De);
        sum += fabsf(e);
        cnt += 1.f;
        if (gacc != nullptr && Der != 0.f) {  // coincident reconstructed points: the reference's sqrt'(0) gives NaN; 0 here
          const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
          const float q = sg * (p.relative ? w / De : w) / Der;
          gx = fmaf(q, rx, gx);
          gy = fmaf(q, ry, gy);
          gz = fmaf(q, rz, gz);
        }
      }
    }
  }
  if (gacc != nullptr && on) {
    float* gp = gacc + ((size_t)b * p.V + vi_id) * 3;
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

// grid (nblk, G, B): d loss / d rec = 2 coef[k] gscale gacc for the vertices of every part (vertices outside every part:
// zeroed by the caller).  The matrix holds both (i,j) and (j,i), and w, De and the mask are symmetric: factor two.
__global__ void __launch_bounds__(PL_THREADS) pl_bwd_kernel(const PLParams p, const float* __restrict__ coef,
                                                            const float* __restrict__ gscale, const float* __restrict__ gacc,
                                                            float* __restrict__ grec) {
  const int blk = blockIdx.x, k = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int g0 = p.gptr[k], n = p.gptr[k + 1] - g0;
  const int i = blk * PL_THREADS + t;
  if (i >= n) return;
  const int vi_id = p.idx[g0 + i];
  const float c = 2.f * coef[k] * __ldg(gscale);
  const size_t o = ((size_t)b * p.V + vi_id) * 3;
  grec[o] = c * gacc[o];
  grec[o + 1] = c * gacc[o + 1];
  grec[o + 2] = c * gacc[o + 2];
}

}  // namespace shb

using namespace shb;

extern "C" {

static size_t pl_partials_floats(int B, int G, int nblk) { return (size_t)B * G * nblk * 2; }

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
