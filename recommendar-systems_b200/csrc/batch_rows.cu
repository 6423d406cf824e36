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
#include <type_traits>

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

// ---- the same for the modality views cat([R x', x']) (smore.py:289-317, mgcn.py:170-184) ---------
// A view's user rows are R x' (one more SpMM over all 19 445 users at Baby size, and R^T in the
// backward); the batch consumes 2 048 of them. Compact row m < B is the CSR row of user users[m] of R
// times the item tables x'_v (all views in one pass over the row's non-zeros), rows B .. 3 B are item
// rows copied from x'_v; the content table is gathered alongside. The backward scatter-adds
// val * g into the item tables along the same non-zeros (vector reductions).
struct ViewTables {
  const float *x[3];       // item tables x'_v [I, d]  (gather) / compact gradients [3 B, d] (scatter)
  float *y[3];             // compact views [3 B, d]   (gather) / dense zeroed gradients [I, d] (scatter)
  const float *content;    // [U + I, d] or NULL       (gather) / compact gradient (scatter)
  float *content_out;      // [3 B, d]                 (gather) / dense zeroed gradient [U + I, d] (scatter)
  int n_views;
};

// One CTA owns 256 / LANES compact rows. A user row is cut into chunks of 64 non-zeros and the chunks of all
// rows of the CTA form one task list that the CTA's sub-warps walk together: the batch draws users in proportion
// to their interactions, so the heaviest user (872 interactions at Baby size) sits in almost every batch, and one
// sub-warp walking its row alone was the whole kernel time (265 us). Gather: chunk sums go to shared memory and the
// row's owner adds them in chunk order (fixed order: the forward is bit-reproducible); scatter: plain reductions.
template <int LANES, bool SCATTER>
__global__ void __launch_bounds__(256)
batch_views_kernel(ViewTables T, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                   const float *__restrict__ vals, int col_offset, const long long *__restrict__ users,
                   const long long *__restrict__ pos, const long long *__restrict__ neg, int B, int n_users,
                   long long *__restrict__ idx_out) {
  constexpr int D = LANES * 4, GROUPS = 256 / LANES, CAP = 768 / LANES;   // CAP chunk sums of 3 views: 36 KB
  __shared__ int s_k0[GROUPS], s_k1[GROUPS], s_base[GROUPS + 1];
  __shared__ float4 s_part[SCATTER ? 1 : CAP * 3 * LANES];
  const int grp = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const int m = blockIdx.x * GROUPS + grp;
  const size_t oc = (size_t)m * D + lane * 4;                       // this lane's chunk of compact row m
  int k0 = 0, k1 = 0;
  if (m < 3 * B) {
    const bool is_user = m < B;
    const long long id = is_user ? users[m] : m < 2 * B ? pos[m - B] : neg[m - 2 * B];
    const long long grow = is_user ? id : n_users + id;             // row of the stacked [U + I] tables
    if (!SCATTER && lane == 0) idx_out[m] = grow;
    if (T.content != nullptr) {
      const size_t og = (size_t)grow * D + lane * 4;
      if (SCATTER) red_add4_rows(T.content_out + og, ldg4(T.content + oc));
      else *reinterpret_cast<float4 *>(T.content_out + oc) = ldg4(T.content + og);
    }
    if (is_user) {
      k0 = row_ptr[id];
      k1 = row_ptr[id + 1];
    } else {
      const size_t oi = (size_t)id * D + lane * 4;
#pragma unroll
      for (int v = 0; v < 3; ++v)
        if (v < T.n_views) {
          if (SCATTER) { if (T.x[v] != nullptr) red_add4_rows(T.y[v] + oi, ldg4(T.x[v] + oc)); }
          else *reinterpret_cast<float4 *>(T.y[v] + oc) = ldg4(T.x[v] + oi);
        }
    }
  }
  if (lane == 0) { s_k0[grp] = k0; s_k1[grp] = k1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = 0;
    for (int r = 0; r < GROUPS; ++r) { s_base[r] = b; b += (s_k1[r] - s_k0[r] + 63) / 64; }
    s_base[GROUPS] = b;
  }
  __syncthreads();
  const int n_tasks = s_base[GROUPS];
  constexpr int NIDX = 64 / LANES, DEPTH = 4;
  const unsigned lane_in_warp = threadIdx.x & 31u;
  const unsigned mask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (LANES * (lane_in_warp / LANES)));
  float4 acc[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t0 = 0; t0 < n_tasks; t0 += CAP) {
    const int t1 = min(n_tasks, t0 + CAP);
    for (int t = t0 + grp; t < t1; t += GROUPS) {
      int r = 0;
      while (s_base[r + 1] <= t) ++r;                               // the row this chunk belongs to
      const int ka = s_k0[r] + (t - s_base[r]) * 64, kb = min(s_k1[r], ka + 64);
      float4 part[3], g[3];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        part[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        g[v] = SCATTER && v < T.n_views && T.x[v] != nullptr
                   ? ldg4(T.x[v] + (size_t)(blockIdx.x * GROUPS + r) * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // all 64 indices / values of the chunk in one round, rows gathered four at a time (like the SpMM kernel)
      int c[NIDX];
      float w[NIDX];
#pragma unroll
      for (int i = 0; i < NIDX; ++i) {
        const int kk = ka + i * LANES + lane;
        c[i] = kk < kb ? __ldg(col_idx + kk) - col_offset : 0;
        w[i] = kk < kb ? __ldg(vals + kk) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < NIDX; ++i) {
#pragma unroll
        for (int j = 0; j < LANES; j += DEPTH) {
          const int base = ka + i * LANES + j;
          if (base < kb) {
            int cc[DEPTH];
            float ww[DEPTH];
#pragma unroll
            for (int q = 0; q < DEPTH; ++q) {
              cc[q] = __shfl_sync(mask, c[i], j + q, LANES);
              ww[q] = __shfl_sync(mask, w[i], j + q, LANES);
            }
            if (SCATTER) {
#pragma unroll
              for (int q = 0; q < DEPTH; ++q)
                if (base + q < kb) {
#pragma unroll
                  for (int v = 0; v < 3; ++v)
                    if (v < T.n_views && T.x[v] != nullptr)
                      red_add4_rows(T.y[v] + (size_t)cc[q] * D + lane * 4,
                                    make_float4(ww[q] * g[v].x, ww[q] * g[v].y, ww[q] * g[v].z, ww[q] * g[v].w));
                }
            } else {
              float4 x[DEPTH][3];
#pragma unroll
              for (int q = 0; q < DEPTH; ++q)
#pragma unroll
                for (int v = 0; v < 3; ++v)
                  x[q][v] = v < T.n_views && base + q < kb ? ldg4(T.x[v] + (size_t)cc[q] * D + lane * 4)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int q = 0; q < DEPTH; ++q)        // CSR order inside the chunk
#pragma unroll
                for (int v = 0; v < 3; ++v) fma4(part[v], ww[q], x[q][v]);
            }
          }
        }
      }
      if (!SCATTER) {
#pragma unroll
        for (int v = 0; v < 3; ++v) s_part[((t - t0) * 3 + v) * LANES + lane] = part[v];
      }
    }
    if (!SCATTER) {
      __syncthreads();
      for (int t = max(s_base[grp], t0); t < min(s_base[grp + 1], t1); ++t) {     // chunk order
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const float4 p = s_part[((t - t0) * 3 + v) * LANES + lane];
          acc[v].x += p.x; acc[v].y += p.y; acc[v].z += p.z; acc[v].w += p.w;
        }
      }
      __syncthreads();
    }
  }
  if (!SCATTER && m < B) {
#pragma unroll
    for (int v = 0; v < 3; ++v)
      if (v < T.n_views) *reinterpret_cast<float4 *>(T.y[v] + oc) = acc[v];
  }
}

template <bool SCATTER>
int batch_views_launch(const ViewTables &T, const int32_t *row_ptr, const int32_t *col_idx, const float *vals,
                       int col_offset, const int64_t *users, const int64_t *pos, const int64_t *neg, int B, int n_users,
                       int d, int64_t *idx_out, cudaStream_t st) {
  const long long *u = reinterpret_cast<const long long *>(users), *p = reinterpret_cast<const long long *>(pos),
                  *q = reinterpret_cast<const long long *>(neg);
  long long *io = reinterpret_cast<long long *>(idx_out);
  auto go = [&](auto lanes) {
    constexpr int L = decltype(lanes)::value, GROUPS = 256 / L;
    batch_views_kernel<L, SCATTER><<<(3 * B + GROUPS - 1) / GROUPS, 256, 0, st>>>(T, row_ptr, col_idx, vals, col_offset, u, p,
                                                                                  q, B, n_users, io);
  };
  switch (d) {
    case 32: go(std::integral_constant<int, 8>{}); break;
    case 64: go(std::integral_constant<int, 16>{}); break;
    case 128: go(std::integral_constant<int, 32>{}); break;
    default: set_error("batch_views: unsupported embedding width d=%d (32, 64, 128)", d); return MMREC_E_BADARG;
  }
  MMREC_CHECK_LAUNCH(SCATTER ? "batch_views_kernel<scatter>" : "batch_views_kernel<gather>");
  return MMREC_OK;
}

int batch_views_args(ViewTables &T, const float *const *a_host, int32_t n_views, const float *content, float *const *b_host,
                     float *content_out, bool scatter) {
  MMREC_REQUIRE(a_host && b_host && n_views >= 1 && n_views <= 3, MMREC_E_BADARG, "batch_views: 1..3 views");
  T = ViewTables{};
  T.n_views = n_views;
  for (int v = 0; v < n_views; ++v) {
    MMREC_REQUIRE((a_host[v] || scatter) && b_host[v] && aligned16(a_host[v]) && aligned16(b_host[v]), MMREC_E_ALIGN,
                  "batch_views: table %d is null or not 16-byte aligned", v);
    T.x[v] = a_host[v];
    T.y[v] = b_host[v];
  }
  MMREC_REQUIRE((content == nullptr || content_out != nullptr) && aligned16(content) && aligned16(content_out), MMREC_E_ALIGN,
                "batch_views: content tables must come in pairs, 16-byte aligned");
  T.content = content;
  T.content_out = content_out;
  return MMREC_OK;
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_gather_batch_views_f32(const int32_t *row_ptr, const int32_t *col_idx, const float *vals,
                                            int32_t col_offset, const float *const *item_tables_host, int32_t n_views,
                                            const float *content, const int64_t *users, const int64_t *pos,
                                            const int64_t *neg, int32_t batch, int32_t n_users, int32_t d,
                                            float *const *views_out_host, float *content_out, int64_t *idx_out,
                                            void *stream) {
  MMREC_REQUIRE(row_ptr && col_idx && vals && users && pos && neg && idx_out && batch >= 0, MMREC_E_BADARG,
                "gather_batch_views: null pointer");
  ViewTables T;
  const int rc = batch_views_args(T, item_tables_host, n_views, content, views_out_host, content_out, false);
  if (rc != MMREC_OK) return rc;
  if (batch == 0) return MMREC_OK;
  return batch_views_launch<false>(T, row_ptr, col_idx, vals, col_offset, users, pos, neg, batch, n_users, d, idx_out,
                                   (cudaStream_t)stream);
}

extern "C" int mmrec_scatter_batch_views_add_f32(const int32_t *row_ptr, const int32_t *col_idx, const float *vals,
                                                 int32_t col_offset, const float *const *d_views_host, int32_t n_views,
                                                 const float *d_content, const int64_t *users, const int64_t *pos,
                                                 const int64_t *neg, int32_t batch, int32_t n_users, int32_t d,
                                                 float *const *d_item_tables_host, float *d_content_out, void *stream) {
  MMREC_REQUIRE(row_ptr && col_idx && vals && users && pos && neg && batch >= 0, MMREC_E_BADARG,
                "scatter_batch_views_add: null pointer");
  ViewTables T;
  const int rc = batch_views_args(T, d_views_host, n_views, d_content, d_item_tables_host, d_content_out, true);
  if (rc != MMREC_OK) return rc;
  if (batch == 0) return MMREC_OK;
  return batch_views_launch<true>(T, row_ptr, col_idx, vals, col_offset, users, pos, neg, batch, n_users, d, nullptr,
                                  (cudaStream_t)stream);
}

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
