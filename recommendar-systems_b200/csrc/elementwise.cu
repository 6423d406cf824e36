// Small single-launch replacements for chains of torch elementwise / reduction kernels on the
// training path -- include/mmrec_b200.h:
//   mmrec_colsum_f32        bias gradient dy.sum(0) of the table projections
//   mmrec_inject3_fwd/bwd   SMORE's residual modality injection item + scale * gate_m and its autograd
#include <algorithm>

#include "common.cuh"

using namespace mmrec;

// ---- column sums: the bias gradient db = dy.sum(0) of the table projections -------------------
// (smore.py:257-259 autograd). One CTA per 4 columns, rows strided over the threads, fixed-order
// tree in shared memory: one launch, no atomics, no counters (safe on concurrent streams). The
// input is the [I, d] gradient that was just written: it comes from L2.
namespace mmrec {
namespace {
__global__ void __launch_bounds__(1024)
colsum4_kernel(const float *__restrict__ x, int M, int N, float *__restrict__ out) {
  __shared__ float4 sh[32];
  const int c0 = blockIdx.x * 4;
  // 1024 threads: ~7 rows each at I = 7050, four independent partial sums -> two memory round trips
  float4 s[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = threadIdx.x;
  for (; r + 3 * 1024 < M; r += 4 * 1024) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 v = ldg4(x + (size_t)(r + u * 1024) * N + c0);
      s[u].x += v.x; s[u].y += v.y; s[u].z += v.z; s[u].w += v.w;
    }
  }
  {   // up to three more rows per thread: issue the loads together, statically indexed accumulators
    float4 v[3];
#pragma unroll
    for (int u = 0; u < 3; ++u)
      v[u] = r + u * 1024 < M ? ldg4(x + (size_t)(r + u * 1024) * N + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 3; ++u) { s[u].x += v[u].x; s[u].y += v[u].y; s[u].z += v[u].z; s[u].w += v[u].w; }
  }
  float4 t = make_float4((s[0].x + s[1].x) + (s[2].x + s[3].x), (s[0].y + s[1].y) + (s[2].y + s[3].y),
                         (s[0].z + s[1].z) + (s[2].z + s[3].z), (s[0].w + s[1].w) + (s[2].w + s[3].w));
  auto warp_sum4 = [](float4 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
      v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
    }
    return v;
  };
  t = warp_sum4(t);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    t = warp_sum4(sh[threadIdx.x]);
    if (threadIdx.x == 0) sh[0] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<float4 *>(out + c0) = sh[0];
}
}  // namespace
}  // namespace mmrec

// Tall matrices (tens of thousands of rows: the [N_nodes, d] gradients of the side-network layers):
// one CTA per 4 columns leaves 32 CTAs each reading 16 bytes of every 512-byte row. Stage 1 gives
// every CTA a band of whole rows (coalesced) and writes one partial row per CTA, stage 2 is the
// kernel above over the partials. Fixed order in both stages: deterministic.
namespace mmrec {
namespace {
constexpr int kBandThreads = 256;
__global__ void __launch_bounds__(kBandThreads)
colsum_band_kernel(const float *__restrict__ x, int M, int N, int rows_per_cta, float *__restrict__ partial) {
  extern __shared__ float4 sh4[];                       // [row groups][N / 4]
  const int cg = N / 4, rg = kBandThreads / cg;         // column groups, row groups (N <= 1024, N / 4 divides 256)
  const int c = threadIdx.x % cg, r0 = threadIdx.x / cg;
  const int begin = blockIdx.x * rows_per_cta, end = min(M, begin + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), t = s;
  int r = begin + r0;
  for (; r + rg < end; r += 2 * rg) {                   // two independent loads in flight
    const float4 a = ldg4(x + (size_t)r * N + c * 4), b = ldg4(x + (size_t)(r + rg) * N + c * 4);
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
  }
  if (r < end) {
    const float4 a = ldg4(x + (size_t)r * N + c * 4);
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
  }
  sh4[r0 * cg + c] = make_float4(s.x + t.x, s.y + t.y, s.z + t.z, s.w + t.w);
  __syncthreads();
  if (r0 == 0) {
    float4 a = sh4[c];
    for (int g = 1; g < rg; ++g) {
      const float4 b = sh4[g * cg + c];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4 *>(partial + (size_t)blockIdx.x * N + c * 4) = a;
  }
}
}  // namespace
}  // namespace mmrec

extern "C" size_t mmrec_colsum_workspace_bytes(int32_t M, int32_t N) {
  // tall matrices with a width whose float4 groups tile a 256-thread CTA take the two-stage path
  if (M < 16384 || N % 4 != 0 || N > 1024 || 256 % (N / 4) != 0) return 0;
  return (size_t)4 * kNumSMs * N * sizeof(float);
}

extern "C" int mmrec_colsum_f32(const float *x, int32_t M, int32_t N, float *out, void *stream_) {
  return mmrec_colsum_ws_f32(x, M, N, out, nullptr, stream_);
}

extern "C" int mmrec_colsum_ws_f32(const float *x, int32_t M, int32_t N, float *out, float *workspace, void *stream_) {
  MMREC_REQUIRE(x && out, MMREC_E_BADARG, "colsum: null pointer");
  MMREC_REQUIRE(M > 0 && N > 0 && N % 4 == 0, MMREC_E_BADARG, "colsum: need M > 0 and N %% 4 == 0");
  MMREC_REQUIRE(aligned16(x) && aligned16(out) && aligned16(workspace), MMREC_E_ALIGN,
                "colsum: operands must be 16-byte aligned");
  if (workspace != nullptr && mmrec_colsum_workspace_bytes(M, N) > 0) {
    const int ctas = 4 * kNumSMs, rows_per_cta = (M + ctas - 1) / ctas;
    const int used = (M + rows_per_cta - 1) / rows_per_cta;
    colsum_band_kernel<<<used, kBandThreads, (size_t)kBandThreads * sizeof(float4), (cudaStream_t)stream_>>>(
        x, M, N, rows_per_cta, workspace);
    MMREC_CHECK_LAUNCH("colsum_band_kernel");
    colsum4_kernel<<<N / 4, 1024, 0, (cudaStream_t)stream_>>>(workspace, used, N, out);
    MMREC_CHECK_LAUNCH("colsum4_kernel");
    return MMREC_OK;
  }
  colsum4_kernel<<<N / 4, 1024, 0, (cudaStream_t)stream_>>>(x, M, N, out);
  MMREC_CHECK_LAUNCH("colsum4_kernel");
  return MMREC_OK;
}

// ---- modality injection: item + scale * gate_m, m = image / text / fusion (smore.py:269-272) ----
// Three axpys forward and, in autograd, three scalings plus a three-way sum backward: 12 torch
// launches per pass; here one launch each way over the [I, d] tensors.
namespace mmrec {
namespace {
__global__ void __launch_bounds__(256)
inject3_fwd_kernel(const float4 *__restrict__ item, const float4 *__restrict__ g0, const float4 *__restrict__ g1,
                   const float4 *__restrict__ g2, float scale, int64_t n4, float4 *__restrict__ o0,
                   float4 *__restrict__ o1, float4 *__restrict__ o2) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 it = item[i], a = g0[i], b = g1[i], c = g2[i];
  // item + (scale * gate): the product is rounded first, like torch's mul followed by add
  o0[i] = make_float4(it.x + __fmul_rn(scale, a.x), it.y + __fmul_rn(scale, a.y), it.z + __fmul_rn(scale, a.z), it.w + __fmul_rn(scale, a.w));
  o1[i] = make_float4(it.x + __fmul_rn(scale, b.x), it.y + __fmul_rn(scale, b.y), it.z + __fmul_rn(scale, b.z), it.w + __fmul_rn(scale, b.w));
  o2[i] = make_float4(it.x + __fmul_rn(scale, c.x), it.y + __fmul_rn(scale, c.y), it.z + __fmul_rn(scale, c.z), it.w + __fmul_rn(scale, c.w));
}
__global__ void __launch_bounds__(256)
inject3_bwd_kernel(const float4 *__restrict__ d0, const float4 *__restrict__ d1, const float4 *__restrict__ d2,
                   float scale, int64_t n4, float4 *__restrict__ d_item, float4 *__restrict__ dg0,
                   float4 *__restrict__ dg1, float4 *__restrict__ dg2) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 a = d0[i], b = d1[i], c = d2[i];
  d_item[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
  dg0[i] = make_float4(scale * a.x, scale * a.y, scale * a.z, scale * a.w);
  dg1[i] = make_float4(scale * b.x, scale * b.y, scale * b.z, scale * b.w);
  dg2[i] = make_float4(scale * c.x, scale * c.y, scale * c.z, scale * c.w);
}
}  // namespace
}  // namespace mmrec

extern "C" int mmrec_inject3_fwd_f32(const float *item, const float *g0, const float *g1, const float *g2, float scale,
                                     int64_t numel, float *o0, float *o1, float *o2, void *stream_) {
  MMREC_REQUIRE(item && g0 && g1 && g2 && o0 && o1 && o2, MMREC_E_BADARG, "inject3_fwd: null pointer");
  MMREC_REQUIRE(numel > 0 && numel % 4 == 0, MMREC_E_BADARG, "inject3_fwd: numel must be a positive multiple of 4");
  MMREC_REQUIRE(aligned16(item) && aligned16(g0) && aligned16(g1) && aligned16(g2) && aligned16(o0) && aligned16(o1) &&
                    aligned16(o2), MMREC_E_ALIGN, "inject3_fwd: operands must be 16-byte aligned");
  const int64_t n4 = numel / 4;
  inject3_fwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      (const float4 *)item, (const float4 *)g0, (const float4 *)g1, (const float4 *)g2, scale, n4, (float4 *)o0,
      (float4 *)o1, (float4 *)o2);
  MMREC_CHECK_LAUNCH("inject3_fwd_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_inject3_bwd_f32(const float *d0, const float *d1, const float *d2, float scale, int64_t numel,
                                     float *d_item, float *dg0, float *dg1, float *dg2, void *stream_) {
  MMREC_REQUIRE(d0 && d1 && d2 && d_item && dg0 && dg1 && dg2, MMREC_E_BADARG, "inject3_bwd: null pointer");
  MMREC_REQUIRE(numel > 0 && numel % 4 == 0, MMREC_E_BADARG, "inject3_bwd: numel must be a positive multiple of 4");
  MMREC_REQUIRE(aligned16(d0) && aligned16(d1) && aligned16(d2) && aligned16(d_item) && aligned16(dg0) &&
                    aligned16(dg1) && aligned16(dg2), MMREC_E_ALIGN, "inject3_bwd: operands must be 16-byte aligned");
  const int64_t n4 = numel / 4;
  inject3_bwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      (const float4 *)d0, (const float4 *)d1, (const float4 *)d2, scale, n4, (float4 *)d_item, (float4 *)dg0,
      (float4 *)dg1, (float4 *)dg2);
  MMREC_CHECK_LAUNCH("inject3_bwd_kernel");
  return MMREC_OK;
}

// ---- loss head of MGCN / SMORE (mgcn.py:241-253, smore.py:396-411) ------------------------------
//   loss = bpr_sum / B + reg_weight * (reg_sum / train_batch_size) + cl_loss * (cl_items + cl_users)
// Six scalar torch kernels forward and as many backward, per pass; one single-thread launch each
// way here. Same float32 operations in the same order as the tensor expression (torch divides by a
// host scalar as a multiplication by its float32 reciprocal), so the value is bit-identical.
namespace mmrec {
namespace {
__global__ void loss_head_fwd_kernel(const float *__restrict__ o2, const float *__restrict__ cl2, float inv_b,
                                     float rw, float inv_bs, float clw, float *__restrict__ out) {
  const float bpr = __fmul_rn(o2[0], inv_b);
  const float reg = __fmul_rn(rw, __fmul_rn(o2[1], inv_bs));
  const float cl = __fmul_rn(clw, __fadd_rn(cl2[0], cl2[1]));
  out[0] = __fadd_rn(__fadd_rn(bpr, reg), cl);
}
__global__ void loss_head_bwd_kernel(const float *__restrict__ g, float inv_b, float rw, float inv_bs, float clw,
                                     float *__restrict__ d_o2, float *__restrict__ d_cl2) {
  const float gg = g[0];
  d_o2[0] = __fmul_rn(gg, inv_b);
  d_o2[1] = __fmul_rn(__fmul_rn(gg, rw), inv_bs);
  const float c = __fmul_rn(gg, clw);
  d_cl2[0] = c;
  d_cl2[1] = c;
}
}  // namespace
}  // namespace mmrec

extern "C" int mmrec_loss_head_fwd_f32(const float *o2, const float *cl2, float inv_batch, float reg_weight,
                                       float inv_train_batch_size, float cl_weight, float *out, void *stream_) {
  MMREC_REQUIRE(o2 && cl2 && out, MMREC_E_BADARG, "loss_head_fwd: null pointer");
  loss_head_fwd_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(o2, cl2, inv_batch, reg_weight, inv_train_batch_size,
                                                          cl_weight, out);
  MMREC_CHECK_LAUNCH("loss_head_fwd_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_loss_head_bwd_f32(const float *g, float inv_batch, float reg_weight, float inv_train_batch_size,
                                       float cl_weight, float *d_o2, float *d_cl2, void *stream_) {
  MMREC_REQUIRE(g && d_o2 && d_cl2, MMREC_E_BADARG, "loss_head_bwd: null pointer");
  loss_head_bwd_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(g, inv_batch, reg_weight, inv_train_batch_size, cl_weight,
                                                          d_o2, d_cl2);
  MMREC_CHECK_LAUNCH("loss_head_bwd_kernel");
  return MMREC_OK;
}

// ---- the fused kernels' activation forms, elementwise (accuracy tests) ---------------------------
namespace mmrec {
namespace {
__global__ void __launch_bounds__(256)
activation_kernel(const float *__restrict__ x, int64_t n, int act, float *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = act == 1 ? fast_tanh(v) : act == 2 ? fast_sigmoid(v) : fast_exp(v);
  }
}
}  // namespace
}  // namespace mmrec

extern "C" int mmrec_activation_f32(const float *x, int64_t numel, int32_t act, float *y, void *stream_) {
  MMREC_REQUIRE(x && y && numel >= 0 && act >= 1 && act <= 3, MMREC_E_BADARG, "activation: bad arguments");
  if (numel == 0) return MMREC_OK;
  const int blocks = (int)std::min<int64_t>((numel + 255) / 256, (int64_t)kNumSMs * 16);
  activation_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(x, numel, act, y);
  MMREC_CHECK_LAUNCH("activation_kernel");
  return MMREC_OK;
}
