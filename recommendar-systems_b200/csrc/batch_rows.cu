// Batch rows of node tables: gather before / scatter-add after a row-local module.
//
// SMORE's modality-aware preference module (smore.py:321-341) and the `content + side` sum are
// ROW-LOCAL: row r of (all_embeds, side_embeds) is a function of row r of (fusion, image, text,
// content) and of the shared weights. A training step consumes only the rows of its batch
// (smore.py:395-407: ua[users], ia[pos], ia[neg], side[users], side[pos], content[users], content[pos]
// -- 3 B = 6 144 of the 26 495 rows at Baby size), and a row that nothing consumes receives a zero
// gradient and adds nothing to any weight gradient. The training forward therefore evaluates the
// module on the gathered rows only and the backward scatter-adds the four input gradients back into
// dense zero tables (rows repeat inside a batch: vector reductions, like the BPR scatter of loss.cu).
// The reference computes all rows; the loss and every gradient are the same numbers (the sums over
// rows run over fewer, non-zero terms).
//
//   row m of the compact tables = users[m]                 m <  B
//                                 n_users + pos[m - B]      B <= m < 2 B
//                                 n_users + neg[m - 2 B]    2 B <= m < 3 B
#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kMaxBatchTables = 4;
struct BatchTables {
  const float *src[kMaxBatchTables];
  float *dst[kMaxBatchTables];
  int n;
};

__device__ __forceinline__ void red_add4_rows(float *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// one thread per (compact row, 16-byte column chunk); the row id is written out once (chunk 0)
__global__ void __launch_bounds__(256)
gather_batch_rows_kernel(BatchTables T, const long long *__restrict__ users, const long long *__restrict__ pos,
                         const long long *__restrict__ neg, int B, int n_users, int d4,
                         long long *__restrict__ idx_out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= 3ll * B * d4) return;
  const int m = (int)(i / d4), c = (int)(i % d4);
  const long long row = m < B ? users[m] : m < 2 * B ? n_users + pos[m - B] : n_users + neg[m - 2 * B];
  if (c == 0) idx_out[m] = row;
#pragma unroll
  for (int t = 0; t < kMaxBatchTables; ++t)
    if (t < T.n)
      reinterpret_cast<float4 *>(T.dst[t])[(size_t)m * d4 + c] = ldg4(T.src[t] + ((size_t)row * d4 + c) * 4);
}

__global__ void __launch_bounds__(256)
scatter_batch_rows_add_kernel(BatchTables T, const long long *__restrict__ idx, int R, int d4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)R * d4) return;
  const int m = (int)(i / d4), c = (int)(i % d4);
  const long long row = idx[m];
#pragma unroll
  for (int t = 0; t < kMaxBatchTables; ++t)
    if (t < T.n && T.src[t] != nullptr)
      red_add4_rows(T.dst[t] + ((size_t)row * d4 + c) * 4, ldg4(T.src[t] + ((size_t)m * d4 + c) * 4));
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_gather_batch_rows_f32(const float *const *src_host, int32_t n_tables, const int64_t *users,
                                           const int64_t *pos, const int64_t *neg, int32_t batch, int32_t n_users,
                                           int32_t d, float *const *dst_host, int64_t *idx_out, void *stream) {
  MMREC_REQUIRE(src_host && dst_host && users && pos && neg && idx_out, MMREC_E_BADARG, "gather_batch_rows: null pointer");
  MMREC_REQUIRE(n_tables >= 1 && n_tables <= kMaxBatchTables && batch >= 0 && n_users >= 0 && d > 0 && d % 4 == 0,
                MMREC_E_BADARG, "gather_batch_rows: 1..%d tables, d a multiple of 4", kMaxBatchTables);
  BatchTables T{};
  T.n = n_tables;
  for (int t = 0; t < n_tables; ++t) {
    MMREC_REQUIRE(src_host[t] && dst_host[t] && aligned16(src_host[t]) && aligned16(dst_host[t]), MMREC_E_ALIGN,
                  "gather_batch_rows: table %d is null or not 16-byte aligned", t);
    T.src[t] = src_host[t];
    T.dst[t] = dst_host[t];
  }
  if (batch == 0) return MMREC_OK;
  const long long total = 3ll * batch * (d / 4);
  gather_batch_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      T, reinterpret_cast<const long long *>(users), reinterpret_cast<const long long *>(pos),
      reinterpret_cast<const long long *>(neg), batch, n_users, d / 4, reinterpret_cast<long long *>(idx_out));
  MMREC_CHECK_LAUNCH("gather_batch_rows_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_scatter_batch_rows_add_f32(const float *const *dsrc_host, int32_t n_tables, const int64_t *idx,
                                                int32_t n_rows, int32_t d, float *const *ddst_host, void *stream) {
  MMREC_REQUIRE(dsrc_host && ddst_host && idx, MMREC_E_BADARG, "scatter_batch_rows_add: null pointer");
  MMREC_REQUIRE(n_tables >= 1 && n_tables <= kMaxBatchTables && n_rows >= 0 && d > 0 && d % 4 == 0, MMREC_E_BADARG,
                "scatter_batch_rows_add: 1..%d tables, d a multiple of 4", kMaxBatchTables);
  BatchTables T{};
  T.n = n_tables;
  for (int t = 0; t < n_tables; ++t) {
    // a NULL source = no gradient for that table (its destination stays zero)
    MMREC_REQUIRE(ddst_host[t] && aligned16(dsrc_host[t]) && aligned16(ddst_host[t]), MMREC_E_ALIGN,
                  "scatter_batch_rows_add: table %d is null or not 16-byte aligned", t);
    T.src[t] = dsrc_host[t];
    T.dst[t] = ddst_host[t];
  }
  if (n_rows == 0) return MMREC_OK;
  const long long total = (long long)n_rows * (d / 4);
  scatter_batch_rows_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      T, reinterpret_cast<const long long *>(idx), n_rows, d / 4);
  MMREC_CHECK_LAUNCH("scatter_batch_rows_add_kernel");
  return MMREC_OK;
}
