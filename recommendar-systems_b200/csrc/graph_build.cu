// Device-side adjacency construction (K11/K12) -- see include/mmrec_b200.h.
//
// One 64-bit radix sort of (row << 32 | col) keys yields the row-major, column-sorted order the
// reference gets from scipy's CSR->COO conversion; row pointers are lower bounds of (row << 32)
// in the sorted keys (no scan, no atomics), and values come from a degree-indexed look-up table
// of deg^-1/2 that the caller evaluates with the reference's own host pow(), so the float
// recipe is bit-exact by construction. CUB's radix sort (part of the CUDA toolkit) is the only
// library piece; it runs once per graph, not on the per-step path.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kThreads = 256;

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int bits_for(uint64_t n) {
  int b = 1;
  while (b < 32 && (1ull << b) < n) ++b;
  return b;
}

__global__ void ui_keys_kernel(const int64_t *__restrict__ users, const int64_t *__restrict__ items,
                               int64_t n_edges, uint32_t n_users, uint64_t *__restrict__ keys) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const uint64_t u = (uint64_t)users[e], i = (uint64_t)items[e] + n_users;
  keys[e] = (u << 32) | i;
  keys[n_edges + e] = (i << 32) | u;
}

__global__ void coo_keys_kernel(const int64_t *__restrict__ rows, const int64_t *__restrict__ cols,
                                int64_t nnz, int transpose, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ idx) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const uint64_t r = (uint64_t)(transpose ? cols[e] : rows[e]);
  const uint64_t c = (uint64_t)(transpose ? rows[e] : cols[e]);
  keys[e] = (r << 32) | c;
  idx[e] = (uint32_t)e;
}

// row_ptr[r] = first position whose key >= (r << 32), r = 0..n_rows
__global__ void row_ptr_kernel(const uint64_t *__restrict__ keys, int64_t nnz, int n_rows,
                               int32_t *__restrict__ row_ptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const uint64_t target = (uint64_t)r << 32;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  row_ptr[r] = (int32_t)lo;
}

template <bool F64>
__global__ void ui_fill_kernel(const uint64_t *__restrict__ keys, int64_t nnz,
                               const int32_t *__restrict__ row_ptr, const void *__restrict__ lut,
                               int lut_len, int32_t *__restrict__ col_idx, float *__restrict__ vals) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const uint64_t key = keys[k];
  const int r = (int)(key >> 32), c = (int)(key & 0xffffffffu);
  const int dr = min(row_ptr[r + 1] - row_ptr[r], lut_len - 1);
  const int dc = min(row_ptr[c + 1] - row_ptr[c], lut_len - 1);
  col_idx[k] = c;
  if constexpr (F64) {
    const double *t = static_cast<const double *>(lut);
    vals[k] = (float)__dmul_rn(t[dr], t[dc]);          // scipy D*A*D in float64, rounded once
  } else {
    const float *t = static_cast<const float *>(lut);
    vals[k] = __fmul_rn(__fmul_rn(t[dr], 1.0f), t[dc]);  // float32 products (mgcn.py recipe)
  }
}

__global__ void degree_kernel(const int32_t *__restrict__ row_ptr, int n_rows, int32_t *__restrict__ deg) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) deg[r] = row_ptr[r + 1] - row_ptr[r];
}

__global__ void coo_fill_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx,
                                const float *__restrict__ vals, int64_t nnz, int32_t *__restrict__ col_idx,
                                float *__restrict__ out_vals, int64_t *__restrict__ perm) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  col_idx[k] = (int32_t)(keys[k] & 0xffffffffu);
  const uint32_t src = idx[k];
  if (out_vals) out_vals[k] = vals[src];
  if (perm) perm[k] = (int64_t)src;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" size_t mmrec_ui_adj_workspace_bytes(int64_t n_edges, int32_t, int32_t) {
  const int64_t nnz = 2 * n_edges;
  size_t temp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, temp, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                 (int)nnz, 0, 64);
  return 2 * align_up(sizeof(uint64_t) * (size_t)nnz) + align_up(temp) + 256;
}

extern "C" int mmrec_ui_adj_build(const int64_t *users, const int64_t *items, int64_t n_edges,
                                  int32_t n_users, int32_t n_items, const void *lut, int32_t lut_len,
                                  int32_t lut_is_f64, int32_t *row_ptr, int32_t *col_idx, float *vals,
                                  int32_t *deg_out, void *workspace, size_t workspace_bytes,
                                  void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(users && items && lut && row_ptr && col_idx && vals && workspace, MMREC_E_BADARG,
                "ui_adj_build: null pointer");
  MMREC_REQUIRE(n_edges >= 0 && n_users > 0 && n_items > 0 && lut_len > 0, MMREC_E_BADARG,
                "ui_adj_build: bad sizes");
  const int64_t nnz = 2 * n_edges;
  MMREC_REQUIRE(nnz < (int64_t)INT32_MAX, MMREC_E_OVERFLOW,
                "ui_adj_build: 2E = %lld does not fit int32 row pointers", (long long)nnz);
  MMREC_REQUIRE(workspace_bytes >= mmrec_ui_adj_workspace_bytes(n_edges, n_users, n_items),
                MMREC_E_WORKSPACE, "ui_adj_build: workspace too small");
  const int n = n_users + n_items;
  char *ws = static_cast<char *>(workspace);
  ws = reinterpret_cast<char *>(align_up(reinterpret_cast<size_t>(ws)));
  uint64_t *keys_in = reinterpret_cast<uint64_t *>(ws);
  ws += align_up(sizeof(uint64_t) * (size_t)nnz);
  uint64_t *keys_out = reinterpret_cast<uint64_t *>(ws);
  ws += align_up(sizeof(uint64_t) * (size_t)nnz);
  size_t temp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, temp, keys_in, keys_out, (int)nnz, 0, 64);
  if (n_edges > 0) {
    ui_keys_kernel<<<grid_for(n_edges), kThreads, 0, stream>>>(users, items, n_edges, (uint32_t)n_users,
                                                              keys_in);
    MMREC_CHECK_LAUNCH("ui_keys_kernel");
    MMREC_CUDA(cub::DeviceRadixSort::SortKeys(ws, temp, keys_in, keys_out, (int)nnz, 0,
                                              32 + bits_for((uint64_t)n), stream));
    count_launch(4);
  }
  row_ptr_kernel<<<grid_for(n + 1), kThreads, 0, stream>>>(keys_out, nnz, n, row_ptr);
  MMREC_CHECK_LAUNCH("row_ptr_kernel");
  if (nnz > 0) {
    if (lut_is_f64)
      ui_fill_kernel<true><<<grid_for(nnz), kThreads, 0, stream>>>(keys_out, nnz, row_ptr, lut, lut_len,
                                                                  col_idx, vals);
    else
      ui_fill_kernel<false><<<grid_for(nnz), kThreads, 0, stream>>>(keys_out, nnz, row_ptr, lut, lut_len,
                                                                   col_idx, vals);
    MMREC_CHECK_LAUNCH("ui_fill_kernel");
  }
  if (deg_out) {
    degree_kernel<<<grid_for(n), kThreads, 0, stream>>>(row_ptr, n, deg_out);
    MMREC_CHECK_LAUNCH("degree_kernel");
  }
  return MMREC_OK;
}

extern "C" size_t mmrec_csr_from_coo_workspace_bytes(int64_t nnz) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)nnz, 0, 64);
  return 2 * align_up(sizeof(uint64_t) * (size_t)nnz) + 2 * align_up(sizeof(uint32_t) * (size_t)nnz) +
         align_up(temp) + 256;
}

extern "C" int mmrec_csr_from_coo(const int64_t *rows, const int64_t *cols, const float *vals, int64_t nnz,
                                  int32_t n_rows, int32_t n_cols, int32_t transpose, int32_t *row_ptr,
                                  int32_t *col_idx, float *out_vals, int64_t *perm_out, void *workspace,
                                  size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(rows && cols && row_ptr && col_idx && workspace, MMREC_E_BADARG, "csr_from_coo: null pointer");
  MMREC_REQUIRE(!out_vals || vals, MMREC_E_BADARG, "csr_from_coo: out_vals without vals");
  MMREC_REQUIRE(nnz >= 0 && n_rows > 0 && n_cols > 0, MMREC_E_BADARG, "csr_from_coo: bad sizes");
  MMREC_REQUIRE(nnz < (int64_t)INT32_MAX, MMREC_E_OVERFLOW, "csr_from_coo: nnz does not fit int32");
  MMREC_REQUIRE(workspace_bytes >= mmrec_csr_from_coo_workspace_bytes(nnz), MMREC_E_WORKSPACE,
                "csr_from_coo: workspace too small");
  const int out_rows = transpose ? n_cols : n_rows;
  const int out_cols = transpose ? n_rows : n_cols;
  char *ws = reinterpret_cast<char *>(align_up(reinterpret_cast<size_t>(workspace)));
  uint64_t *keys_in = reinterpret_cast<uint64_t *>(ws);
  ws += align_up(sizeof(uint64_t) * (size_t)nnz);
  uint64_t *keys_out = reinterpret_cast<uint64_t *>(ws);
  ws += align_up(sizeof(uint64_t) * (size_t)nnz);
  uint32_t *idx_in = reinterpret_cast<uint32_t *>(ws);
  ws += align_up(sizeof(uint32_t) * (size_t)nnz);
  uint32_t *idx_out = reinterpret_cast<uint32_t *>(ws);
  ws += align_up(sizeof(uint32_t) * (size_t)nnz);
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, keys_in, keys_out, idx_in, idx_out, (int)nnz, 0, 64);
  if (nnz > 0) {
    coo_keys_kernel<<<grid_for(nnz), kThreads, 0, stream>>>(rows, cols, nnz, transpose, keys_in, idx_in);
    MMREC_CHECK_LAUNCH("coo_keys_kernel");
    // stable LSD radix sort: equal (row, col) keys keep their COO order
    MMREC_CUDA(cub::DeviceRadixSort::SortPairs(ws, temp, keys_in, keys_out, idx_in, idx_out, (int)nnz, 0,
                                               32 + bits_for((uint64_t)out_rows), stream));
    count_launch(4);
    (void)out_cols;
  }
  row_ptr_kernel<<<grid_for(out_rows + 1), kThreads, 0, stream>>>(keys_out, nnz, out_rows, row_ptr);
  MMREC_CHECK_LAUNCH("row_ptr_kernel");
  if (nnz > 0) {
    coo_fill_kernel<<<grid_for(nnz), kThreads, 0, stream>>>(keys_out, idx_out, vals, nnz, col_idx, out_vals,
                                                           perm_out);
    MMREC_CHECK_LAUNCH("coo_fill_kernel");
  }
  return MMREC_OK;
}
