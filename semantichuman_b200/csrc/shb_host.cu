// Host-side index construction for libshb200: inverse-spiral tables and CSR forms of the D/U sampling matrices.
// Pure CPU code (no CUDA calls) so that it also works on a box without a GPU.
#include <algorithm>
#include <utility>
#include <vector>

#include "shb_common.cuh"

namespace shb {
static int g_persistent_sms = kNumSMs;
int persistent_sms() { return g_persistent_sms; }
}  // namespace shb

extern "C" {

int shb_set_persistent_sms(int n) {
  if (n < 1 || n > shb::kNumSMs) return SHB_E_ARG;
  shb::g_persistent_sms = n;
  return 0;
}

int shb_abi_version(void) { return SHB_ABI_VERSION; }

const char* shb_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SHB_E_ARG: return "shb200: invalid argument (null pointer or non-positive size)";
    case SHB_E_DTYPE: return "shb200: unsupported dtype enum";
    case SHB_E_SHAPE: return "shb200: unsupported shape for this kernel";
    case SHB_E_WORKSPACE: return "shb200: workspace too small";
    case SHB_E_UNSUPPORTED: return "shb200: operation not supported in this build";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "shb200: unknown error code";
}

// Stable counting sort of the flattened table by source row.
int shb_build_inverse_spiral_csr(const int32_t* table, int rows_out, int S, int rows_in, int32_t* rowptr,
                                 int32_t* slots) {
  if (!table || !rowptr || !slots || rows_out <= 0 || S <= 0 || rows_in <= 0) return SHB_E_ARG;
  const long long n = (long long)rows_out * S;
  for (int u = 0; u <= rows_in; ++u) rowptr[u] = 0;
  for (long long i = 0; i < n; ++i) {
    const int u = table[i];
    if (u < 0 || u >= rows_in) return SHB_E_SHAPE;
    rowptr[u + 1]++;
  }
  for (int u = 0; u < rows_in; ++u) rowptr[u + 1] += rowptr[u];
  std::vector<int32_t> cur(rowptr, rowptr + rows_in);
  for (long long i = 0; i < n; ++i) slots[cur[table[i]]++] = (int32_t)i;
  return 0;
}

// Key (u, s) -> ascending output rows j with table[j, s] == u.
int shb_dense_to_csr(const float* dense, int rows, int cols, int32_t* rowptr, int32_t* colidx, float* vals,
                     int64_t cap, int64_t* nnz_out) {
  if (!dense || rows <= 0 || cols <= 0 || !nnz_out) return SHB_E_ARG;
  int64_t nnz = 0;
  for (int r = 0; r < rows; ++r) {
    if (rowptr) rowptr[r] = (int32_t)nnz;
    const float* row = dense + (size_t)r * cols;
    for (int c = 0; c < cols; ++c) {
      if (row[c] != 0.0f) {
        if (colidx) {
          if (nnz >= cap) return SHB_E_WORKSPACE;
          colidx[nnz] = c;
          vals[nnz] = row[c];
        }
        ++nnz;
      }
    }
  }
  if (rowptr) rowptr[rows] = (int32_t)nnz;
  *nnz_out = nnz;
  return 0;
}

int shb_csr_transpose(const int32_t* rowptr, const int32_t* colidx, const float* vals, int rows, int cols,
                      int32_t* t_rowptr, int32_t* t_colidx, float* t_vals) {
  if (!rowptr || !colidx || !vals || !t_rowptr || !t_colidx || !t_vals || rows <= 0 || cols <= 0) return SHB_E_ARG;
  const int nnz = rowptr[rows];
  for (int c = 0; c <= cols; ++c) t_rowptr[c] = 0;
  for (int e = 0; e < nnz; ++e) {
    if (colidx[e] < 0 || colidx[e] >= cols) return SHB_E_SHAPE;
    t_rowptr[colidx[e] + 1]++;
  }
  for (int c = 0; c < cols; ++c) t_rowptr[c + 1] += t_rowptr[c];
  std::vector<int32_t> cur(t_rowptr, t_rowptr + cols);
  for (int r = 0; r < rows; ++r)
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const int d = cur[colidx[e]]++;
      t_colidx[d] = r;
      t_vals[d] = vals[e];
    }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ conv groups
// Destination rows whose entry lists share source rows are clustered into groups of <= R; a group's distinct sources are then
// loaded ONCE by the conv kernel and every (destination, slot) pair that uses a source is one MMA group on the slab in shared
// memory (shb_slab_gconv.cu).  Greedy clustering in destination order: a group grows by the unassigned destination that adds
// the fewest new sources (ties: most shared sources, then lowest index).  Deterministic; pure function of the lists.
namespace {

struct GroupBuild {
  std::vector<int32_t> gptr, recs, gdst;
  std::vector<uint32_t> gmask;
};

constexpr int GREC_WORDS = 48, GREC_MAX_PAIRS = 32, GREC_MAX_SRC = 8;

int build_groups(const int32_t* ptr, const int32_t* ent, int rows_dst, int rows_src, int R, int SPS, GroupBuild& out) {
  const int n_ent = ptr[rows_dst];
  // inverse lists: source -> destinations (with multiplicity removed per destination)
  std::vector<int32_t> inv_ptr(rows_src + 1, 0), inv;
  std::vector<int32_t> nsrc(rows_dst, 0);      // distinct sources per destination
  {
    std::vector<int32_t> last(rows_src, -1);
    for (int j = 0; j < rows_dst; ++j)
      for (int e = ptr[j]; e < ptr[j + 1]; ++e) {
        const int u = ent[e] >> 5;
        if (u < 0 || u >= rows_src) return SHB_E_SHAPE;
        if (last[u] != j) { last[u] = j; inv_ptr[u + 1]++; nsrc[j]++; }
      }
    for (int u = 0; u < rows_src; ++u) inv_ptr[u + 1] += inv_ptr[u];
    inv.resize(inv_ptr[rows_src]);
    std::vector<int32_t> cur(inv_ptr.begin(), inv_ptr.end() - 1);
    std::fill(last.begin(), last.end(), -1);
    for (int j = 0; j < rows_dst; ++j)
      for (int e = ptr[j]; e < ptr[j + 1]; ++e) {
        const int u = ent[e] >> 5;
        if (last[u] != j) { last[u] = j; inv[cur[u]++] = j; }
      }
  }
  (void)n_ent;
  std::vector<char> assigned(rows_dst, 0);
  std::vector<int32_t> stamp(rows_src, -1);     // source already in the current group's union
  std::vector<int32_t> cnt(rows_dst, 0);        // shared sources of an unassigned destination with the current union
  std::vector<int32_t> touched;
  std::vector<int32_t> group, uni;
  std::vector<std::pair<int32_t, int32_t>> pairs;  // (source, (dest_local << 5) | slot)
  out.gptr.assign(1, 0);
  int gid = 0;
  for (int seed = 0; seed < rows_dst; ++seed) {
    if (assigned[seed]) continue;
    group.clear(); uni.clear(); touched.clear();
    auto add = [&](int j) {
      assigned[j] = 1;
      group.push_back(j);
      for (int e = ptr[j]; e < ptr[j + 1]; ++e) {
        const int u = ent[e] >> 5;
        if (stamp[u] == gid) continue;
        stamp[u] = gid;
        uni.push_back(u);
        for (int k = inv_ptr[u]; k < inv_ptr[u + 1]; ++k) {
          const int d = inv[k];
          if (assigned[d]) continue;
          if (cnt[d]++ == 0) touched.push_back(d);
        }
      }
    };
    add(seed);
    while ((int)group.size() < R) {
      int best = -1, best_new = 0, best_cnt = 0;
      for (int d : touched) {
        if (assigned[d]) continue;
        const int nn = nsrc[d] - cnt[d];
        if (best < 0 || nn < best_new || (nn == best_new && (cnt[d] > best_cnt || (cnt[d] == best_cnt && d < best)))) {
          best = d; best_new = nn; best_cnt = cnt[d];
        }
      }
      if (best < 0) break;
      add(best);
    }
    for (int d : touched) cnt[d] = 0;
    // destinations, untouched mask
    uint32_t mask = 0;
    for (int i = 0; i < R; ++i) {
      const int j = i < (int)group.size() ? group[i] : -1;
      out.gdst.push_back(j);
      if (j >= 0 && ptr[j] == ptr[j + 1]) mask |= 1u << i;
    }
    out.gmask.push_back(mask);
    // pairs sorted by (source, destination order in the group, list order): the fixed accumulation order of every destination
    pairs.clear();
    for (int i = 0; i < (int)group.size(); ++i)
      for (int e = ptr[group[i]]; e < ptr[group[i] + 1]; ++e) pairs.push_back({ent[e] >> 5, (i << 5) | (ent[e] & 31)});
    std::stable_sort(pairs.begin(), pairs.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    // records: <= SPS sources and <= 32 pairs each; a source with more pairs than fit is continued in the next record
    std::vector<char> seen(group.size(), 0);
    const size_t rec0 = out.recs.size();
    int32_t* rec = nullptr;
    int ns = 0, np = 0;
    auto open = [&]() {
      out.recs.resize(out.recs.size() + GREC_WORDS, 0);
      rec = out.recs.data() + out.recs.size() - GREC_WORDS;
      ns = np = 0;
    };
    auto close = [&]() { rec[0] = ns | (np << 4); };
    open();
    size_t i = 0;
    while (i < pairs.size()) {
      size_t j = i;
      while (j < pairs.size() && pairs[j].first == pairs[i].first) ++j;
      size_t k = i;
      while (k < j) {
        const int want = (int)(j - k);
        if (ns == SPS || np == GREC_MAX_PAIRS || (np + want > GREC_MAX_PAIRS && want <= GREC_MAX_PAIRS && ns > 0)) {
          close();
          open();
        }
        const int take = want < GREC_MAX_PAIRS - np ? want : GREC_MAX_PAIRS - np;
        rec[1 + ns] = pairs[i].first;
        for (int t = 0; t < take; ++t) {
          const int dl = pairs[k + t].second >> 5, slot = pairs[k + t].second & 31;
          const int first = seen[dl] ? 0 : 1;
          seen[dl] = 1;
          rec[16 + np + t] = ns | (dl << 3) | (first << 8) | (slot << 9);
        }
        np += take;
        ++ns;
        k += take;
      }
      i = j;
    }
    close();
    const size_t nrec = (out.recs.size() - rec0) / GREC_WORDS;
    out.recs[rec0] |= 1 << 10;                                   // first record of the group
    out.recs[rec0 + (nrec - 1) * GREC_WORDS] |= 1 << 11;        // last
    out.gptr.push_back((int32_t)(out.recs.size() / GREC_WORDS));
    ++gid;
  }
  return 0;
}

}  // namespace

extern "C" {

int shb_build_conv_groups(const int32_t* ptr, const int32_t* ent, int rows_dst, int rows_src, int R, int SPS, int32_t* n_groups,
                          int32_t* n_records, int32_t* gptr, int32_t* recs, int32_t* gdst, uint32_t* gmask) {
  if (!ptr || !ent || !n_groups || !n_records || rows_dst <= 0 || rows_src <= 0 || R < 1 || R > 32 || SPS < 1 ||
      SPS > GREC_MAX_SRC)
    return SHB_E_ARG;
  GroupBuild b;
  const int rc = build_groups(ptr, ent, rows_dst, rows_src, R, SPS, b);
  if (rc != 0) return rc;
  const int ng = (int)b.gmask.size(), nr = (int)(b.recs.size() / GREC_WORDS);
  if (gptr || recs || gdst || gmask) {  // second call: the caller allocated from the counts of the first
    if (!gptr || !recs || !gdst || !gmask || *n_groups != ng || *n_records != nr) return SHB_E_WORKSPACE;
    std::copy(b.gptr.begin(), b.gptr.end(), gptr);
    std::copy(b.recs.begin(), b.recs.end(), recs);
    std::copy(b.gdst.begin(), b.gdst.end(), gdst);
    std::copy(b.gmask.begin(), b.gmask.end(), gmask);
  }
  *n_groups = ng;
  *n_records = nr;
  return 0;
}

}  // extern "C"
