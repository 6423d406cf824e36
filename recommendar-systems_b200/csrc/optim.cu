// Fused multi-tensor Adam (K13) -- one launch updates every parameter tensor of the model.
//
// torch.optim.Adam's default (foreach) path makes ~8 passes over param/grad/exp_avg/exp_avg_sq
// (trainer.py:126-143); for SMORE's 33.6 M parameters (28.9 M of them the trainable 7050 x 4096
// image table) that is the largest HBM consumer of a step. Here each element is read once
// (p, g, m, v) and written once (p, m, v): 28 bytes per parameter, the roofline of the update.
// The arithmetic follows torch's single-tensor formulas (lerp for exp_avg, mul+addcmul for
// exp_avg_sq, sqrt / bias_correction2_sqrt + eps, addcdiv with lr / bias_correction1).
#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kMaxTensors = 48;
constexpr int kChunk = 16384;      // elements per CTA-iteration
constexpr int kThreads = 256;

struct AdamPack {
  float *p[kMaxTensors];
  const float *g[kMaxTensors];
  float *m[kMaxTensors];
  float *v[kMaxTensors];
  const float *undo[kMaxTensors];     // optional: p += undo_coef[0] * undo[t] before the update
  int chunk_begin[kMaxTensors + 1];   // prefix sum of ceil(numel / kChunk)
  int64_t numel[kMaxTensors];
  int n_tensors;
};

struct AxpyPack {
  float *y[kMaxTensors];
  const float *x[kMaxTensors];
  int chunk_begin[kMaxTensors + 1];
  int64_t numel[kMaxTensors];
  int n_tensors;
};

// hyper[0] = learning rate, hyper[1] = number of updates including this one (device memory, so a
// captured CUDA graph replays with the current values).
__global__ void adam_tick_kernel(double *hyper) { hyper[1] += 1.0; }

__global__ void __launch_bounds__(kThreads)
adam_kernel(const __grid_constant__ AdamPack pack, const double *__restrict__ hyper, double beta1d,
            double beta2d, float eps, float weight_decay, float grad_scale, const float *__restrict__ undo_coef) {
  __shared__ float s_hyper[2];
  if (threadIdx.x == 0) {
    const double lr = hyper[0], step = hyper[1];
    s_hyper[0] = (float)(lr / (1.0 - pow(beta1d, step)));      // lr / bias_correction1
    s_hyper[1] = (float)sqrt(1.0 - pow(beta2d, step));         // sqrt(bias_correction2)
  }
  __syncthreads();
  const float step_size = s_hyper[0], bc2_sqrt = s_hyper[1];
  const float beta1 = (float)beta1d, beta2 = (float)beta2d;
  const int chunk = blockIdx.x;
  int t = 0;
  while (t + 1 < pack.n_tensors && pack.chunk_begin[t + 1] <= chunk) ++t;
  const int64_t begin = (int64_t)(chunk - pack.chunk_begin[t]) * kChunk;
  const int64_t n = pack.numel[t];
  const int64_t end = min(n, begin + kChunk);
  float *__restrict__ p = pack.p[t];
  const float *__restrict__ g = pack.g[t];
  float *__restrict__ m = pack.m[t];
  float *__restrict__ v = pack.v[t];
  const float w1 = (float)(1.0 - beta1d), w2 = (float)(1.0 - beta2d);
  (void)beta1;
  // mirror-gradient step (trainer.py:307-335): the parameters were moved to theta - c*g_old for the
  // second backward; they return to theta here, in the pass that applies the update, with the same
  // fmaf the separate axpy kernel would have used (bit-identical, one 12-byte-per-element pass less)
  const float *__restrict__ ux = undo_coef != nullptr ? pack.undo[t] : nullptr;
  const float uc = undo_coef != nullptr ? undo_coef[0] : 0.f;
  auto upd = [&](float &pp, float gg, float &mm, float &vv) {
    gg *= grad_scale;                                          // 1.0f is exact
    if (weight_decay != 0.f) gg = fmaf(weight_decay, pp, gg);
    mm = mm + w1 * (gg - mm);
    vv = vv * beta2 + w2 * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp = pp - step_size * (mm / denom);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  if (vec) {
    const int64_t end4 = begin + ((end - begin) & ~int64_t(3));
    for (int64_t i = begin + threadIdx.x * 4; i < end4; i += kThreads * 4) {
      float4 pp = *reinterpret_cast<float4 *>(p + i), mm = *reinterpret_cast<float4 *>(m + i),
             vv = *reinterpret_cast<float4 *>(v + i);
      const float4 gg = __ldcs(reinterpret_cast<const float4 *>(g + i));
      if (ux != nullptr) {
        const float4 xx = __ldcs(reinterpret_cast<const float4 *>(ux + i));
        pp.x = fmaf(uc, xx.x, pp.x); pp.y = fmaf(uc, xx.y, pp.y); pp.z = fmaf(uc, xx.z, pp.z); pp.w = fmaf(uc, xx.w, pp.w);
      }
      upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
      upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
      *reinterpret_cast<float4 *>(p + i) = pp;
      *reinterpret_cast<float4 *>(m + i) = mm;
      *reinterpret_cast<float4 *>(v + i) = vv;
    }
    for (int64_t i = end4 + threadIdx.x; i < end; i += kThreads) {
      if (ux != nullptr) p[i] = fmaf(uc, ux[i], p[i]);
      upd(p[i], g[i], m[i], v[i]);
    }
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += kThreads) {
      if (ux != nullptr) p[i] = fmaf(uc, ux[i], p[i]);
      upd(p[i], g[i], m[i], v[i]);
    }
  }
}

// ---- mirror-gradient step size (trainer.py:289-305) ---------------------------------------
// alpha_eff = clamp(target * rms(theta) / (lr * rms(g) + 1e-12), alpha, alpha * max_scale) needs
// sum g^2 and sum theta^2 over every parameter: one pass over both lists (per-CTA partials in
// double), then one CTA adds the partials in a fixed order and writes coef = alpha_eff * lr.
// Replaces two _foreach_norm passes and ~20 scalar torch kernels.
__global__ void __launch_bounds__(kThreads)
sumsq_pair_kernel(const __grid_constant__ AxpyPack pack, double *__restrict__ partial) {
  __shared__ float red[kThreads / 32];
  const int chunk = blockIdx.x;
  int t = 0;
  while (t + 1 < pack.n_tensors && pack.chunk_begin[t + 1] <= chunk) ++t;
  const int64_t begin = (int64_t)(chunk - pack.chunk_begin[t]) * kChunk;
  const int64_t end = min(pack.numel[t], begin + kChunk);
  const float *__restrict__ y = pack.y[t];
  const float *__restrict__ x = pack.x[t];
  float sy = 0.f, sx = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0;
  if (vec) {
    const int64_t end4 = begin + ((end - begin) & ~int64_t(3));
    for (int64_t i = begin + threadIdx.x * 4; i < end4; i += kThreads * 4) {
      const float4 yy = *reinterpret_cast<const float4 *>(y + i), xx = *reinterpret_cast<const float4 *>(x + i);
      sy += dot4(yy, yy);
      sx += dot4(xx, xx);
    }
    for (int64_t i = end4 + threadIdx.x; i < end; i += kThreads) { sy = fmaf(y[i], y[i], sy); sx = fmaf(x[i], x[i], sx); }
  } else {
    for (int64_t i = begin + threadIdx.x; i < end; i += kThreads) { sy = fmaf(y[i], y[i], sy); sx = fmaf(x[i], x[i], sx); }
  }
  const float ty = block_sum<kThreads>(sy, red);
  const float tx = block_sum<kThreads>(sx, red);
  if (threadIdx.x == 0) {
    partial[2 * (size_t)chunk] = (double)ty;
    partial[2 * (size_t)chunk + 1] = (double)tx;
  }
}

__global__ void __launch_bounds__(kThreads)
mirror_coef_kernel(const double *__restrict__ partial, int n_parts, const double *__restrict__ extra, int n_extra_p2,
                   int n_extra_g2, const double *__restrict__ hyper, double numel, float alpha_base,
                   float alpha_max_scale, float target_rel, float *__restrict__ out) {
  __shared__ double sh[2][kThreads];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_parts; i += kThreads) { a += partial[2 * (size_t)i]; b += partial[2 * (size_t)i + 1]; }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    double p2 = 0.0, g2 = 0.0;
    for (int i = 0; i < kThreads; ++i) { p2 += sh[0][i]; g2 += sh[1][i]; }
    // tensors whose gradient is a never-materialised low-rank product bring their two sums along
    for (int i = 0; i < n_extra_p2; ++i) p2 += extra[i];
    for (int i = 0; i < n_extra_g2; ++i) g2 += extra[n_extra_p2 + i];
    // float32 arithmetic in the order of the reference's tensor expression
    const float lr = (float)hyper[0], n = (float)numel;
    const float grad_rms = sqrtf((float)g2 / n);
    const float param_rms = sqrtf((float)p2 / n) + 1e-12f;
    float alpha = target_rel * param_rms / (lr * grad_rms + 1e-12f);
    alpha = fminf(fmaxf(alpha, alpha_base), alpha_base * alpha_max_scale);
    out[0] = alpha * lr;      // coef: theta' = theta - coef * g
    out[1] = alpha;           // alpha_eff (logging)
  }
}

// y_t += sign * coef[0] * x_t for every tensor t (coef is a device scalar)
__global__ void __launch_bounds__(kThreads)
axpy_multi_kernel(const __grid_constant__ AxpyPack pack, const float *__restrict__ coef, float sign) {
  const int chunk = blockIdx.x;
  int t = 0;
  while (t + 1 < pack.n_tensors && pack.chunk_begin[t + 1] <= chunk) ++t;
  const int64_t begin = (int64_t)(chunk - pack.chunk_begin[t]) * kChunk;
  const int64_t end = min(pack.numel[t], begin + kChunk);
  float *__restrict__ y = pack.y[t];
  const float *__restrict__ x = pack.x[t];
  const float a = sign * coef[0];
  const bool vec = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0;
  int64_t i = begin + (vec ? threadIdx.x * 4 : threadIdx.x);
  if (vec) {
    const int64_t end4 = begin + ((end - begin) & ~int64_t(3));
    for (; i < end4; i += kThreads * 4) {
      float4 yy = *reinterpret_cast<float4 *>(y + i);
      const float4 xx = *reinterpret_cast<const float4 *>(x + i);
      yy.x = fmaf(a, xx.x, yy.x); yy.y = fmaf(a, xx.y, yy.y); yy.z = fmaf(a, xx.z, yy.z); yy.w = fmaf(a, xx.w, yy.w);
      *reinterpret_cast<float4 *>(y + i) = yy;
    }
    for (i = end4 + threadIdx.x; i < end; i += kThreads) y[i] = fmaf(a, x[i], y[i]);
  } else {
    for (; i < end; i += kThreads) y[i] = fmaf(a, x[i], y[i]);
  }
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_adam_tick(double *hyper, void *stream) {
  MMREC_REQUIRE(hyper, MMREC_E_BADARG, "adam_tick: null pointer");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper);
  MMREC_CHECK_LAUNCH("adam_tick_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_adam_step_f32(float *const *params_host, const float *const *grads_host,
                                   float *const *exp_avg_host, float *const *exp_avg_sq_host,
                                   const int64_t *numel_host, int32_t n_tensors, double *hyper, double beta1,
                                   double beta2, double eps, double weight_decay, double grad_scale,
                                   const float *const *undo_host, const float *undo_coef, int32_t tick,
                                   void *stream) {
  MMREC_REQUIRE(params_host && grads_host && exp_avg_host && exp_avg_sq_host && numel_host && hyper,
                MMREC_E_BADARG, "adam: null pointer");
  MMREC_REQUIRE((undo_host == nullptr) == (undo_coef == nullptr), MMREC_E_BADARG,
                "adam: undo tensors and undo coefficient must be given together");
  MMREC_REQUIRE(n_tensors >= 0, MMREC_E_BADARG, "adam: bad sizes");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper);
    MMREC_CHECK_LAUNCH("adam_tick_kernel");
  }
  for (int base = 0; base < n_tensors; base += kMaxTensors) {
    AdamPack pack;
    pack.n_tensors = min(kMaxTensors, n_tensors - base);
    int chunks = 0;
    for (int i = 0; i < pack.n_tensors; ++i) {
      pack.p[i] = params_host[base + i];
      pack.g[i] = grads_host[base + i];
      pack.m[i] = exp_avg_host[base + i];
      pack.v[i] = exp_avg_sq_host[base + i];
      pack.undo[i] = undo_host != nullptr ? undo_host[base + i] : nullptr;
      MMREC_REQUIRE(undo_host == nullptr || pack.undo[i] != nullptr, MMREC_E_BADARG, "adam: bad undo tensor %d", base + i);
      pack.numel[i] = numel_host[base + i];
      MMREC_REQUIRE(pack.p[i] && pack.g[i] && pack.m[i] && pack.v[i] && pack.numel[i] >= 0, MMREC_E_BADARG,
                    "adam: bad tensor %d", base + i);
      pack.chunk_begin[i] = chunks;
      chunks += (int)((pack.numel[i] + kChunk - 1) / kChunk);
    }
    pack.chunk_begin[pack.n_tensors] = chunks;
    if (chunks == 0) continue;
    adam_kernel<<<chunks, kThreads, 0, (cudaStream_t)stream>>>(pack, hyper, beta1, beta2, (float)eps,
                                                              (float)weight_decay, (float)grad_scale,
                                                              undo_host != nullptr ? undo_coef : nullptr);
    MMREC_CHECK_LAUNCH("adam_kernel");
  }
  return MMREC_OK;
}

extern "C" int mmrec_axpy_multi_f32(float *const *y_host, const float *const *x_host, const int64_t *numel_host,
                                    int32_t n_tensors, const float *coef, float sign, void *stream) {
  MMREC_REQUIRE(y_host && x_host && numel_host && coef, MMREC_E_BADARG, "axpy_multi: null pointer");
  for (int base = 0; base < n_tensors; base += kMaxTensors) {
    AxpyPack pack;
    pack.n_tensors = min(kMaxTensors, n_tensors - base);
    int chunks = 0;
    for (int i = 0; i < pack.n_tensors; ++i) {
      pack.y[i] = y_host[base + i];
      pack.x[i] = x_host[base + i];
      pack.numel[i] = numel_host[base + i];
      MMREC_REQUIRE(pack.y[i] && pack.x[i] && pack.numel[i] >= 0, MMREC_E_BADARG, "axpy_multi: bad tensor %d",
                    base + i);
      pack.chunk_begin[i] = chunks;
      chunks += (int)((pack.numel[i] + kChunk - 1) / kChunk);
    }
    pack.chunk_begin[pack.n_tensors] = chunks;
    if (chunks == 0) continue;
    axpy_multi_kernel<<<chunks, kThreads, 0, (cudaStream_t)stream>>>(pack, coef, sign);
    MMREC_CHECK_LAUNCH("axpy_multi_kernel");
  }
  return MMREC_OK;
}

extern "C" size_t mmrec_mirror_coef_workspace_bytes(const int64_t *numel_host, int32_t n_tensors) {
  size_t chunks = 0;
  for (int i = 0; i < n_tensors; ++i) chunks += (size_t)((numel_host[i] + kChunk - 1) / kChunk);
  return 2 * sizeof(double) * (chunks + 1);
}

extern "C" int mmrec_mirror_coef_f32(const float *const *params_host, const float *const *grads_host,
                                     const int64_t *numel_host, int32_t n_tensors, const double *hyper,
                                     double numel_total, double alpha_base, double alpha_max_scale,
                                     double target_rel_step, const double *extra, int32_t n_extra_p2,
                                     int32_t n_extra_g2, void *workspace, float *coef_out, void *stream) {
  MMREC_REQUIRE(params_host && grads_host && numel_host && hyper && workspace && coef_out, MMREC_E_BADARG,
                "mirror_coef: null pointer");
  MMREC_REQUIRE(n_tensors > 0 && numel_total > 0, MMREC_E_BADARG, "mirror_coef: bad sizes");
  double *partial = static_cast<double *>(workspace);
  int done = 0;
  for (int base = 0; base < n_tensors; base += kMaxTensors) {
    AxpyPack pack;
    pack.n_tensors = min(kMaxTensors, n_tensors - base);
    int chunks = 0;
    for (int i = 0; i < pack.n_tensors; ++i) {
      pack.y[i] = const_cast<float *>(params_host[base + i]);
      pack.x[i] = grads_host[base + i];
      pack.numel[i] = numel_host[base + i];
      MMREC_REQUIRE(pack.y[i] && pack.x[i] && pack.numel[i] >= 0, MMREC_E_BADARG, "mirror_coef: bad tensor %d",
                    base + i);
      pack.chunk_begin[i] = chunks;
      chunks += (int)((pack.numel[i] + kChunk - 1) / kChunk);
    }
    pack.chunk_begin[pack.n_tensors] = chunks;
    if (chunks == 0) continue;
    sumsq_pair_kernel<<<chunks, kThreads, 0, (cudaStream_t)stream>>>(pack, partial + 2 * (size_t)done);
    MMREC_CHECK_LAUNCH("sumsq_pair_kernel");
    done += chunks;
  }
  MMREC_REQUIRE(n_extra_p2 >= 0 && n_extra_g2 >= 0 && (extra != nullptr || n_extra_p2 + n_extra_g2 == 0),
                MMREC_E_BADARG, "mirror_coef: bad extra sums");
  mirror_coef_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(partial, done, extra, n_extra_p2, n_extra_g2, hyper,
                                                             numel_total, (float)alpha_base,
                                                             (float)alpha_max_scale, (float)target_rel_step,
                                                             coef_out);
  MMREC_CHECK_LAUNCH("mirror_coef_kernel");
  return MMREC_OK;
}

// ---- feature tables whose gradient is the rank-d product dY W (SMORE / MGCN / FREEDOM image and
// text tables: d x 4096 / d x 384 projections, smore.py:257-259) -------------------------------
namespace mmrec {
int table_adam_ctas(int rows, int cols);
int table_sumsq_ctas(int rows, int cols);
int table_adam_dispatch(float *P, float *Mo, float *V, const float *dY, const float *W, int rows, int cols, int d,
                        const double *hyper, double beta1, double beta2, double eps, double weight_decay,
                        double grad_scale, double *sumsq_partial, cudaStream_t stream);
int table_sumsq_dispatch(const float *dY, const float *W, int rows, int cols, int d, double *sumsq_partial,
                         cudaStream_t stream);
namespace {
__global__ void __launch_bounds__(kThreads) sum_partials_kernel(const double *__restrict__ partial, int n,
                                                                double *__restrict__ out) {
  __shared__ double sh[kThreads];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += kThreads) a += partial[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kThreads; ++i) t += sh[i];
    out[0] = t;
  }
}
bool lowrank_shape_ok(int rows, int cols, int d) {
  return rows > 0 && cols > 0 && cols % 64 == 0 && (d == 32 || d == 64 || d == 128);
}
}  // namespace
}  // namespace mmrec

extern "C" int mmrec_table_lowrank_supported(int32_t rows, int32_t cols, int32_t d) {
  return lowrank_shape_ok(rows, cols, d);
}

extern "C" size_t mmrec_table_lowrank_workspace_bytes(int32_t rows, int32_t cols) {
  if (rows <= 0 || cols <= 0 || cols % 64) return 0;
  const int a = table_adam_ctas(rows, cols), b = table_sumsq_ctas(rows, cols);
  return sizeof(double) * (size_t)(a > b ? a : b);
}

extern "C" int mmrec_table_adam_lowrank_f32(float *table, float *exp_avg, float *exp_avg_sq, const float *dY,
                                            const float *W, int32_t rows, int32_t cols, int32_t d, double *hyper,
                                            double beta1, double beta2, double eps, double weight_decay,
                                            double grad_scale, int32_t tick, double *sumsq_out, void *workspace,
                                            void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(table && exp_avg && exp_avg_sq && dY && W && hyper, MMREC_E_BADARG, "table_adam: null pointer");
  MMREC_REQUIRE(lowrank_shape_ok(rows, cols, d), MMREC_E_BADARG,
                "table_adam: need cols %% 64 == 0 and d in {32, 64, 128} (got %d x %d, d = %d)", rows, cols, d);
  MMREC_REQUIRE(aligned16(table) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(dY) && aligned16(W),
                MMREC_E_ALIGN, "table_adam: operands must be 16-byte aligned");
  MMREC_REQUIRE(sumsq_out == nullptr || workspace != nullptr, MMREC_E_WORKSPACE,
                "table_adam: the parameter square sum needs the workspace");
  if (tick) {
    adam_tick_kernel<<<1, 1, 0, stream>>>(hyper);
    MMREC_CHECK_LAUNCH("adam_tick_kernel");
  }
  double *partial = sumsq_out ? static_cast<double *>(workspace) : nullptr;
  const int rc = table_adam_dispatch(table, exp_avg, exp_avg_sq, dY, W, rows, cols, d, hyper, beta1, beta2, eps,
                                     weight_decay, grad_scale, partial, stream);
  if (rc != MMREC_OK) return rc;
  if (sumsq_out) {
    sum_partials_kernel<<<1, kThreads, 0, stream>>>(partial, table_adam_ctas(rows, cols), sumsq_out);
    MMREC_CHECK_LAUNCH("sum_partials_kernel");
  }
  return MMREC_OK;
}

extern "C" int mmrec_table_lowrank_sumsq_f64(const float *dY, const float *W, int32_t rows, int32_t cols, int32_t d,
                                             void *workspace, double *out, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(dY && W && workspace && out, MMREC_E_BADARG, "table_sumsq: null pointer");
  MMREC_REQUIRE(lowrank_shape_ok(rows, cols, d), MMREC_E_BADARG,
                "table_sumsq: need cols %% 64 == 0 and d in {32, 64, 128} (got %d x %d, d = %d)", rows, cols, d);
  MMREC_REQUIRE(aligned16(dY) && aligned16(W), MMREC_E_ALIGN, "table_sumsq: operands must be 16-byte aligned");
  double *partial = static_cast<double *>(workspace);
  const int rc = table_sumsq_dispatch(dY, W, rows, cols, d, partial, stream);
  if (rc != MMREC_OK) return rc;
  sum_partials_kernel<<<1, kThreads, 0, stream>>>(partial, table_sumsq_ctas(rows, cols), out);
  MMREC_CHECK_LAUNCH("sum_partials_kernel");
  return MMREC_OK;
}
