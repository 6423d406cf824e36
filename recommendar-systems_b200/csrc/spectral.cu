// SMORE spectrum-based modality fusion (K5) as real circulant operators -- include/mmrec_b200.h.
//
// irfft(rfft(x) * w, norm='ortho') is a circular convolution x (*) h with h = irfft_backward(w)
// (the imaginary parts of the DC and Nyquist bins are ignored by irfft), and
// irfft(rfft(t) * rfft(v) * w_f) = ((t (*) v) (*) h_f) / sqrt(d). For d = 64 that is 4 x 64-tap
// circular convolutions per item row: one thread per output element, operands staged in shared
// memory (x broadcast, taps lane-contiguous -> conflict-free). No cuFFT plans, no complex
// intermediates in HBM, no host syncs: traffic is exactly 2 rows in, 3 rows out.
#include "common.cuh"

namespace mmrec {
namespace {

// taps[f][n] = irfft_backward(w_hat_f)[n], w_hat = w / (|w| + 1e-8) when weight_norm.
__global__ void spectral_taps_kernel(const float *__restrict__ w_img, const float *__restrict__ w_txt,
                                     const float *__restrict__ w_fus, int d, int weight_norm,
                                     float *__restrict__ taps) {
  const int f = blockIdx.x, n = threadIdx.x;
  if (n >= d) return;
  const float *w = f == 0 ? w_img : (f == 1 ? w_txt : w_fus);
  const int half = d / 2;
  float acc = 0.f;
  for (int k = 0; k <= half; ++k) {
    float a = w[2 * k], b = w[2 * k + 1];
    if (weight_norm) {
      const float s = 1.f / (sqrtf(a * a + b * b) + 1e-8f);
      a *= s;
      b *= s;
    }
    if (k == 0) {
      acc += a;
    } else if (k == half) {
      acc += (n & 1) ? -a : a;
    } else {
      float sn, cs;
      sincospif(2.f * (float)((k * n) % d) / (float)d, &sn, &cs);
      acc += 2.f * (a * cs - b * sn);
    }
  }
  taps[f * d + n] = acc / (float)d;
}

// y[n] = sum_m a[m] * b[(n - m) mod D]  (a, b in shared memory)
template <int D>
__device__ __forceinline__ float circ_conv(const float *a, const float *b, int n) {
  float s = 0.f;
#pragma unroll 8
  for (int m = 0; m < D; ++m) s = fmaf(a[m], b[(n - m) & (D - 1)], s);
  return s;
}
// z[m] = sum_n g[n] * b[(n - m) mod D]   (correlation: adjoint of circ_conv w.r.t. a)
template <int D>
__device__ __forceinline__ float circ_corr(const float *g, const float *b, int m) {
  float s = 0.f;
#pragma unroll 8
  for (int n = 0; n < D; ++n) s = fmaf(g[n], b[(n - m) & (D - 1)], s);
  return s;
}

template <int D>
__global__ void __launch_bounds__(256)
spectral_fwd_kernel(const float *__restrict__ img, const float *__restrict__ txt, int n_rows,
                    const float *__restrict__ taps, float *__restrict__ ic, float *__restrict__ tc,
                    float *__restrict__ fc) {
  constexpr int ROWS = 256 / D;
  __shared__ float sh[3][D];
  __shared__ float sx[ROWS][D], st[ROWS][D], sc[ROWS][D];
  for (int t = threadIdx.x; t < 3 * D; t += 256) sh[t / D][t % D] = taps[t];
  const int r = threadIdx.x / D, n = threadIdx.x % D;
  const int row = blockIdx.x * ROWS + r;
  const bool live = row < n_rows;
  sx[r][n] = live ? img[(size_t)row * D + n] : 0.f;
  st[r][n] = live ? txt[(size_t)row * D + n] : 0.f;
  __syncthreads();
  const float vi = circ_conv<D>(sx[r], sh[0], n);
  const float vt = circ_conv<D>(st[r], sh[1], n);
  sc[r][n] = circ_conv<D>(st[r], sx[r], n);
  __syncthreads();
  const float vf = circ_conv<D>(sc[r], sh[2], n) * rsqrtf((float)D);
  if (live) {
    ic[(size_t)row * D + n] = vi;
    tc[(size_t)row * D + n] = vt;
    fc[(size_t)row * D + n] = vf;
  }
}

template <int D>
__global__ void __launch_bounds__(256)
spectral_bwd_kernel(const float *__restrict__ img, const float *__restrict__ txt, int n_rows,
                    const float *__restrict__ taps, const float *__restrict__ g_ic,
                    const float *__restrict__ g_tc, const float *__restrict__ g_fc,
                    float *__restrict__ d_img, float *__restrict__ d_txt, float *__restrict__ dh) {
  constexpr int ROWS = 256 / D;
  __shared__ float sh[3][D];
  __shared__ float sx[ROWS][D], st[ROWS][D], sc[ROWS][D];
  __shared__ float gi[ROWS][D], gt[ROWS][D], gf[ROWS][D], gc[ROWS][D];
  __shared__ float sdh[3][ROWS][D];
  for (int t = threadIdx.x; t < 3 * D; t += 256) sh[t / D][t % D] = taps[t];
  const int r = threadIdx.x / D, n = threadIdx.x % D;
  const int row = blockIdx.x * ROWS + r;
  const bool live = row < n_rows;
  const size_t o = (size_t)row * D + n;
  sx[r][n] = live ? img[o] : 0.f;
  st[r][n] = live ? txt[o] : 0.f;
  gi[r][n] = live ? g_ic[o] : 0.f;
  gt[r][n] = live ? g_tc[o] : 0.f;
  gf[r][n] = live ? g_fc[o] * rsqrtf((float)D) : 0.f;     // fold the 1/sqrt(d) of the fusion path
  __syncthreads();
  sc[r][n] = circ_conv<D>(st[r], sx[r], n);               // c = t (*) x
  gc[r][n] = circ_corr<D>(gf[r], sh[2], n);               // dL/dc
  __syncthreads();
  // tap gradients: dh[j] = sum_n g[n] * a[(n - j)]  == circ_corr(g, a, j)
  sdh[0][r][n] = circ_corr<D>(gi[r], sx[r], n);
  sdh[1][r][n] = circ_corr<D>(gt[r], st[r], n);
  sdh[2][r][n] = circ_corr<D>(gf[r], sc[r], n);
  const float dx = circ_corr<D>(gi[r], sh[0], n) + circ_corr<D>(gc[r], st[r], n);
  const float dt = circ_corr<D>(gt[r], sh[1], n) + circ_corr<D>(gc[r], sx[r], n);
  if (live) {
    d_img[o] = dx;
    d_txt[o] = dt;
  }
  __syncthreads();
  if (threadIdx.x < D) {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < ROWS; ++q) s += sdh[f][q][threadIdx.x];
      atomicAdd(dh + f * D + threadIdx.x, s);
    }
  }
}

// dh (taps) -> raw weight gradients through irfft_backward and the unit-magnitude map.
__global__ void spectral_weight_bwd_kernel(const float *__restrict__ w_img, const float *__restrict__ w_txt,
                                           const float *__restrict__ w_fus, int d, int weight_norm,
                                           const float *__restrict__ dh, float *__restrict__ d_w_img,
                                           float *__restrict__ d_w_txt, float *__restrict__ d_w_fus) {
  const int f = blockIdx.x, k = threadIdx.x;
  const int half = d / 2;
  if (k > half) return;
  const float *w = f == 0 ? w_img : (f == 1 ? w_txt : w_fus);
  float *dw = f == 0 ? d_w_img : (f == 1 ? d_w_txt : d_w_fus);
  const float *g = dh + f * d;
  float da = 0.f, db = 0.f;   // gradient w.r.t. the (normalised) real / imaginary parts
  for (int n = 0; n < d; ++n) {
    if (k == 0) {
      da += g[n];
    } else if (k == half) {
      da += (n & 1) ? -g[n] : g[n];
    } else {
      float sn, cs;
      sincospif(2.f * (float)((k * n) % d) / (float)d, &sn, &cs);
      da += 2.f * g[n] * cs;
      db -= 2.f * g[n] * sn;
    }
  }
  da /= (float)d;
  db /= (float)d;
  const float p = w[2 * k], q = w[2 * k + 1];
  if (weight_norm) {
    const float r = sqrtf(p * p + q * q), s = r + 1e-8f;
    // a_hat = p/s, b_hat = q/s; d|w|/dp = p/r (0 at the origin, like torch.abs)
    const float rp = r > 0.f ? p / r : 0.f, rq = r > 0.f ? q / r : 0.f;
    const float t = (da * p + db * q) / (s * s);
    dw[2 * k] = da / s - t * rp;
    dw[2 * k + 1] = db / s - t * rq;
  } else {
    dw[2 * k] = da;
    dw[2 * k + 1] = db;
  }
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_spectral_fwd_f32(const float *img, const float *txt, int32_t n_rows, int32_t d,
                                      const float *w_img, const float *w_txt, const float *w_fus,
                                      int32_t weight_norm, float *taps_ws, float *img_conv, float *txt_conv,
                                      float *fus_conv, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(img && txt && w_img && w_txt && w_fus && taps_ws && img_conv && txt_conv && fus_conv,
                MMREC_E_BADARG, "spectral_fwd: null pointer");
  MMREC_REQUIRE(n_rows > 0, MMREC_E_BADARG, "spectral_fwd: empty input");
  MMREC_REQUIRE(d == 32 || d == 64 || d == 128, MMREC_E_BADARG, "spectral_fwd: unsupported d=%d (32, 64, 128)", d);
  spectral_taps_kernel<<<3, d, 0, stream>>>(w_img, w_txt, w_fus, d, weight_norm, taps_ws);
  MMREC_CHECK_LAUNCH("spectral_taps_kernel");
  const int rows_per_block = 256 / d;
  const int blocks = (n_rows + rows_per_block - 1) / rows_per_block;
  if (d == 32) spectral_fwd_kernel<32><<<blocks, 256, 0, stream>>>(img, txt, n_rows, taps_ws, img_conv, txt_conv, fus_conv);
  else if (d == 64) spectral_fwd_kernel<64><<<blocks, 256, 0, stream>>>(img, txt, n_rows, taps_ws, img_conv, txt_conv, fus_conv);
  else spectral_fwd_kernel<128><<<blocks, 256, 0, stream>>>(img, txt, n_rows, taps_ws, img_conv, txt_conv, fus_conv);
  MMREC_CHECK_LAUNCH("spectral_fwd_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_spectral_bwd_f32(const float *img, const float *txt, int32_t n_rows, int32_t d,
                                      const float *w_img, const float *w_txt, const float *w_fus,
                                      int32_t weight_norm, const float *taps_ws, const float *g_img_conv,
                                      const float *g_txt_conv, const float *g_fus_conv, float *d_img, float *d_txt,
                                      float *dh_ws, float *d_w_img, float *d_w_txt, float *d_w_fus,
                                      void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(img && txt && w_img && w_txt && w_fus && taps_ws && g_img_conv && g_txt_conv && g_fus_conv &&
                    d_img && d_txt && dh_ws && d_w_img && d_w_txt && d_w_fus,
                MMREC_E_BADARG, "spectral_bwd: null pointer");
  MMREC_REQUIRE(n_rows > 0, MMREC_E_BADARG, "spectral_bwd: empty input");
  MMREC_REQUIRE(d == 32 || d == 64 || d == 128, MMREC_E_BADARG, "spectral_bwd: unsupported d=%d (32, 64, 128)", d);
  const int rows_per_block = 256 / d;
  const int blocks = (n_rows + rows_per_block - 1) / rows_per_block;
  if (d == 32) spectral_bwd_kernel<32><<<blocks, 256, 0, stream>>>(img, txt, n_rows, taps_ws, g_img_conv, g_txt_conv, g_fus_conv, d_img, d_txt, dh_ws);
  else if (d == 64) spectral_bwd_kernel<64><<<blocks, 256, 0, stream>>>(img, txt, n_rows, taps_ws, g_img_conv, g_txt_conv, g_fus_conv, d_img, d_txt, dh_ws);
  else spectral_bwd_kernel<128><<<blocks, 256, 0, stream>>>(img, txt, n_rows, taps_ws, g_img_conv, g_txt_conv, g_fus_conv, d_img, d_txt, dh_ws);
  MMREC_CHECK_LAUNCH("spectral_bwd_kernel");
  spectral_weight_bwd_kernel<<<3, d / 2 + 1, 0, stream>>>(w_img, w_txt, w_fus, d, weight_norm, dh_ws, d_w_img, d_w_txt, d_w_fus);
  MMREC_CHECK_LAUNCH("spectral_weight_bwd_kernel");
  return MMREC_OK;
}
