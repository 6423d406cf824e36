// Item-item kNN modality graphs -- include/mmrec_b200.h.
//   build_sim (utils/utils.py:134-137)                     -> row_normalize + the library GEMM
//   torch.topk(sim, k) (utils.py:172, freedom.py:81)       -> row_topk_kernel
//   get_sparse_laplacian 'sym' (utils.py:139-152)          -> knn_weights_kernel (mode 0)
//   compute_normalized_laplacian (freedom.py:87-100)       -> knn_weights_kernel (mode 1)
// The reference walks the [I, k] neighbour lists element by element in a Python list comprehension
// (utils.py:175: minutes on Clothing); here the lists never leave the device.
#include <math_constants.h>

#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kThreads = 256;

// out[r] = x[r] / ||x[r]||_2 ; one warp per row, float4 loads (d % 4 == 0)
__global__ void __launch_bounds__(kThreads)
row_normalize_kernel(const float *__restrict__ x, int n_rows, int d, float *__restrict__ out) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float *src = x + (size_t)row * d;
  float s = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 v = ldg4(src + c);
    s += dot4(v, v);
  }
  s = warp_sum(s);
  const float nrm = sqrtf(s);
  float *dst = out + (size_t)row * d;
  for (int c = lane * 4; c < d; c += 128) {
    float4 v = ldg4(src + c);
    v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm;        // feat.div(norm): 0/0 = nan like torch
    *reinterpret_cast<float4 *>(dst + c) = v;
  }
}

// Strict total order used for the selection: a ranks before b iff higher value, ties -> lower column.
__device__ __forceinline__ bool before(float av, int ai, float bv, int bi) {
  return av > bv || (av == bv && ai < bi);
}

// Top-k of every row of a dense [n_rows, ld] matrix (first n_cols columns), descending, ties ->
// lower column. One CTA per row: the row is read once into shared memory (or re-read from L2 when
// it does not fit), then k rounds of "best element ranking after the previous pick": each thread
// scans its strided share, a shuffle + shared-memory tournament picks the block winner.
__global__ void __launch_bounds__(kThreads)
row_topk_kernel(const float *__restrict__ mat, int n_rows, int n_cols, int64_t ld, int k, int in_smem,
                float *__restrict__ out_val, int32_t *__restrict__ out_idx) {
  extern __shared__ float srow[];
  __shared__ float w_val[kThreads / 32];
  __shared__ int w_idx[kThreads / 32];
  __shared__ float pick_val;
  __shared__ int pick_idx;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *src = mat + (size_t)row * ld;
  if (in_smem) {
    for (int c = tid; c < n_cols; c += kThreads) srow[c] = src[c];
    __syncthreads();
  }
  const float *rowp = in_smem ? srow : src;
  float last_v = CUDART_INF_F;
  int last_i = -1;
  for (int r = 0; r < k; ++r) {
    float bv = -CUDART_INF_F;
    int bi = INT_MAX;
    for (int c = tid; c < n_cols; c += kThreads) {
      const float v = rowp[c];
      // eligible: ranks strictly after the previous pick (NaN never compares true: skipped)
      const bool elig = r == 0 ? (v == v) : before(last_v, last_i, v, c);
      if (elig && before(v, c, bv, bi)) { bv = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { w_val[warp] = bv; w_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kThreads / 32; ++w)
        if (before(w_val[w], w_idx[w], bv, bi)) { bv = w_val[w]; bi = w_idx[w]; }
      pick_val = bv;
      pick_idx = bi;
      out_val[(size_t)row * k + r] = bi == INT_MAX ? -CUDART_INF_F : bv;
      out_idx[(size_t)row * k + r] = bi == INT_MAX ? -1 : bi;
    }
    __syncthreads();
    last_v = pick_val;
    last_i = pick_idx;
    if (last_i == INT_MAX) {              // fewer than k comparable entries: the rest stays empty
      if (tid == 0)
        for (int q = r + 1; q < k; ++q) { out_val[(size_t)row * k + q] = -CUDART_INF_F; out_idx[(size_t)row * k + q] = -1; }
      break;
    }
  }
}

// deg[r] = sum_j w[r][j] (mode 0) -- the row sums get_sparse_laplacian scatter-adds (utils.py:143-146)
__global__ void knn_degree_kernel(const float *__restrict__ val, int n, int k, float *__restrict__ dis) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float s = 0.f;
  for (int j = 0; j < k; ++j) s += val[(size_t)r * k + j];
  float d = powf(s, -0.5f);                                  // torch.pow(deg, -0.5)
  if (isinf(d)) d = 0.f;                                     // masked_fill_(== inf, 0)
  dis[r] = d;
}
// mode 0: out = dis[r] * w * dis[c];  mode 1 (FREEDOM, binary edges): out = s * s, s = (k + 1e-7)^-1/2
__global__ void knn_weights_kernel(const int32_t *__restrict__ idx, const float *__restrict__ val,
                                   const float *__restrict__ dis, int n, int k, int mode, float *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * k) return;
  if (mode == 1) {
    const float s = powf(1e-7f + (float)k, -0.5f);           // freedom.py:92-94 in float32
    out[e] = s * s;
    return;
  }
  const int r = (int)(e / k), c = idx[e];
  out[e] = c >= 0 ? dis[r] * val[e] * dis[c] : 0.f;
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_row_normalize_f32(const float *x, int32_t n_rows, int32_t d, float *out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(x && out, MMREC_E_BADARG, "row_normalize: null pointer");
  MMREC_REQUIRE(n_rows > 0 && d > 0 && d % 4 == 0, MMREC_E_BADARG, "row_normalize: need n_rows > 0, d %% 4 == 0");
  MMREC_REQUIRE(aligned16(x) && aligned16(out), MMREC_E_ALIGN, "row_normalize: rows must be 16-byte aligned");
  row_normalize_kernel<<<(n_rows + 7) / 8, kThreads, 0, stream>>>(x, n_rows, d, out);
  MMREC_CHECK_LAUNCH("row_normalize_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_row_topk_f32(const float *mat, int32_t n_rows, int32_t n_cols, int64_t ld, int32_t k,
                                  float *out_val, int32_t *out_idx, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(mat && out_val && out_idx, MMREC_E_BADARG, "row_topk: null pointer");
  MMREC_REQUIRE(n_rows > 0 && n_cols > 0 && ld >= n_cols && k > 0 && k <= n_cols, MMREC_E_BADARG,
                "row_topk: need 0 < k <= n_cols <= ld");
  const size_t bytes = (size_t)n_cols * sizeof(float);
  const int in_smem = bytes <= 200 * 1024;
  if (in_smem) MMREC_CUDA(cudaFuncSetAttribute(row_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  row_topk_kernel<<<n_rows, kThreads, in_smem ? bytes : 0, stream>>>(mat, n_rows, n_cols, ld, k, in_smem, out_val,
                                                                     out_idx);
  MMREC_CHECK_LAUNCH("row_topk_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_knn_weights_f32(const int32_t *idx, const float *val, int32_t n, int32_t k, int32_t mode,
                                     float *dis_ws, float *out_vals, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(idx && out_vals && (mode == 1 || (val && dis_ws)), MMREC_E_BADARG, "knn_weights: null pointer");
  MMREC_REQUIRE(n > 0 && k > 0 && (mode == 0 || mode == 1), MMREC_E_BADARG, "knn_weights: bad arguments");
  if (mode == 0) {
    knn_degree_kernel<<<(n + 255) / 256, 256, 0, stream>>>(val, n, k, dis_ws);
    MMREC_CHECK_LAUNCH("knn_degree_kernel");
  }
  const int64_t total = (int64_t)n * k;
  knn_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(idx, val, dis_ws, n, k, mode, out_vals);
  MMREC_CHECK_LAUNCH("knn_weights_kernel");
  return MMREC_OK;
}

