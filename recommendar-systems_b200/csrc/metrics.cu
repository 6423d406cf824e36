// Top-K evaluation metrics on the device (K15): the hit matrix and Recall / Recall2 / Precision /
// NDCG / MAP @1..K of TopKEvaluator._calculate_metrics (utils/topk_evaluator.py:88-101,
// utils/metrics.py:12-109) without moving the [n_users, K] id matrix to the host. The reference
// builds the hit matrix with a Python membership loop (67 % of its CPU evaluation time) and
// reduces it with numpy in float64.
//
// One warp per evaluated user: lane r tests rank r and r + 32 against the user's ground-truth
// items (ascending int32 CSR, binary search), two ballots give the 64-bit hit mask, and every lane
// rebuilds the prefix quantities of its rank in float64 IN RANK ORDER (the order of numpy's
// cumsum), so per-user values are bit-identical to the reference's. `disc` (1/log2(r+2)) and
// `idcg_all` (its cumsum) come from the host, computed with numpy. Per-CTA partial sums over
// users are added in CTA order by a second kernel (deterministic; numpy's pairwise mean differs
// from it in the last bits only).
#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kThreadsM = 256;
constexpr int kWarpsM = kThreadsM / 32;
constexpr int kNumMetrics = 5;      // recall, recall2 numerator, precision, ndcg, map

__global__ void __launch_bounds__(kThreadsM)
topk_metrics_kernel(const int64_t *__restrict__ topk, int n_users, int k, const int32_t *__restrict__ gt_rowptr,
                    const int32_t *__restrict__ gt_items, const double *__restrict__ disc,
                    const double *__restrict__ idcg_all, uint8_t *__restrict__ hits_out,
                    double *__restrict__ partial) {
  __shared__ double acc[kWarpsM][kNumMetrics][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * kWarpsM + warp;
  double val[2][kNumMetrics];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int m = 0; m < kNumMetrics; ++m) val[h][m] = 0.0;
  if (u < n_users) {
    const int g0 = gt_rowptr[u], g1 = gt_rowptr[u + 1], pos_len = g1 - g0;
    unsigned bits[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lane + 32 * h;
      bool hit = false;
      if (r < k) {
        const int64_t id = topk[(size_t)u * k + r];
        int lo = g0, hi = g1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((int64_t)gt_items[mid] < id) lo = mid + 1; else hi = mid;
        }
        hit = lo < g1 && (int64_t)gt_items[lo] == id;
        if (hits_out != nullptr) hits_out[(size_t)u * k + r] = hit;
      }
      bits[h] = __ballot_sync(0xffffffffu, hit);
    }
    const unsigned long long mask = (unsigned long long)bits[0] | ((unsigned long long)bits[1] << 32);
    const int idcg_len = min(pos_len, k);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lane + 32 * h;
      if (r < k) {
        // prefix sums over ranks 0..r in rank order (numpy cumsum order)
        double dcg = 0.0, sum_pre = 0.0;
        int c = 0;
        for (int j = 0; j <= r; ++j) {
          if ((mask >> j) & 1ull) {
            ++c;
            dcg += disc[j];
            sum_pre += (double)c / (double)(j + 1);
          }
        }
        const double ch = (double)c;
        val[h][0] = ch / (double)pos_len;                                   // recall_
        val[h][1] = ch;                                                     // recall2_ numerator
        val[h][2] = ch / (double)(r + 1);                                   // precision_
        val[h][3] = dcg / idcg_all[min(r, idcg_len - 1)];                   // ndcg_
        val[h][4] = sum_pre / (double)min(r + 1, idcg_len);                 // map_
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int m = 0; m < kNumMetrics; ++m) acc[warp][m][lane + 32 * h] = val[h][m];
  __syncthreads();
  // warps of the CTA in order -> one partial per CTA
  for (int i = threadIdx.x; i < kNumMetrics * 64; i += kThreadsM) {
    const int m = i / 64, r = i % 64;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarpsM; ++w) s += acc[w][m][r];
    partial[((size_t)blockIdx.x * kNumMetrics + m) * 64 + r] = s;
  }
}

// Fixed-order sum of the per-block partials: one CTA per metric, 16 slices of the part list per
// rank (slice q adds parts q, q + 16, ... in order, four independent chains), then the slices in
// order. (One thread per rank walking all ~1100 parts was a 116 us chain of dependent loads.)
constexpr int kRedSlices = 16;
__global__ void __launch_bounds__(64 * kRedSlices)
topk_metrics_reduce_kernel(const double *__restrict__ partial, int n_parts, int k, double *__restrict__ out) {
  __shared__ double sh[kRedSlices][64];
  const int m = blockIdx.x, r = threadIdx.x & 63, q = threadIdx.x >> 6;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (r < k) {
    int p = q;
    for (; p + 3 * kRedSlices < n_parts; p += 4 * kRedSlices) {
      s0 += partial[((size_t)p * kNumMetrics + m) * 64 + r];
      s1 += partial[((size_t)(p + kRedSlices) * kNumMetrics + m) * 64 + r];
      s2 += partial[((size_t)(p + 2 * kRedSlices) * kNumMetrics + m) * 64 + r];
      s3 += partial[((size_t)(p + 3 * kRedSlices) * kNumMetrics + m) * 64 + r];
    }
    for (; p < n_parts; p += kRedSlices) s0 += partial[((size_t)p * kNumMetrics + m) * 64 + r];
  }
  sh[q][r] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (q == 0 && r < k) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kRedSlices; ++j) s += sh[j][r];
    out[m * k + r] = s;
  }
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" size_t mmrec_topk_metrics_workspace_bytes(int32_t n_users) {
  return sizeof(double) * (size_t)((n_users + kWarpsM - 1) / kWarpsM) * kNumMetrics * 64;
}

extern "C" int mmrec_topk_metrics_f64(const int64_t *topk, int32_t n_users, int32_t k, const int32_t *gt_rowptr,
                                      const int32_t *gt_items, const double *disc, const double *idcg_all,
                                      uint8_t *hits_out, double *sums_out, void *workspace, void *stream) {
  MMREC_REQUIRE(topk && gt_rowptr && gt_items && disc && idcg_all && sums_out && workspace, MMREC_E_BADARG,
                "topk_metrics: null pointer");
  MMREC_REQUIRE(n_users > 0 && k >= 1 && k <= 64, MMREC_E_BADARG, "topk_metrics: need n_users > 0 and 1 <= k <= 64");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (n_users + kWarpsM - 1) / kWarpsM;
  double *partial = reinterpret_cast<double *>(workspace);
  topk_metrics_kernel<<<blocks, kThreadsM, 0, st>>>(topk, n_users, k, gt_rowptr, gt_items, disc, idcg_all, hits_out,
                                                   partial);
  MMREC_CHECK_LAUNCH("topk_metrics_kernel");
  topk_metrics_reduce_kernel<<<kNumMetrics, 64 * kRedSlices, 0, st>>>(partial, blocks, k, sums_out);
  MMREC_CHECK_LAUNCH("topk_metrics_reduce_kernel");
  return MMREC_OK;
}
