// Host-side index construction for libshb200: inverse-spiral tables and CSR forms of the D/U sampling matrices.
// Pure CPU code (no CUDA calls) so that it also works on a box without a GPU.
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
