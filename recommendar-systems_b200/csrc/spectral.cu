// SMORE spectrum-based modality fusion (K5) as real circulant operators -- include/mmrec_b200.h.
//
// irfft(rfft(x) * w, norm='ortho') is a circular convolution x (*) h with h = irfft_backward(w)
// (the imaginary parts of the DC and Nyquist bins are ignored by irfft), and
// irfft(rfft(t) * rfft(v) * w_f) = ((t (*) v) (*) h_f) / sqrt(d). For d = 64 that is 4 x 64-tap
// circular convolutions per item row (9 in the backward), register-tiled on the CUDA cores (see
// circ_tile). No cuFFT plans, no complex intermediates in HBM, no host syncs: traffic is exactly
// 2 rows in, 3 rows out.
#include "common.cuh"

namespace mmrec {
namespace {

// taps[f][n] = irfft_backward(w_hat_f)[n], w_hat = w / (|w| + 1e-8) when weight_norm.
// One CTA per filter; the d twiddles cos / sin(2 pi m / d) and the (normalised) bins are computed
// once into shared memory, so the d/2 + 1 terms of every tap are two table reads and two FMAs
// (evaluating sincospif + sqrt + divide inside the loop made this 3-CTA kernel cost 16 us).
__global__ void spectral_taps_kernel(const float *__restrict__ w_img, const float *__restrict__ w_txt,
                                     const float *__restrict__ w_fus, int d, int weight_norm,
                                     float *__restrict__ taps) {
  __shared__ float cs_t[128], sn_t[128], wa[65], wb[65];
  const int f = blockIdx.x, n = threadIdx.x;
  const float *w = f == 0 ? w_img : (f == 1 ? w_txt : w_fus);
  const int half = d / 2;
  if (n < d) sincospif(2.f * (float)n / (float)d, &sn_t[n], &cs_t[n]);
  if (n <= half) {
    float a = w[2 * n], b = w[2 * n + 1];
    if (weight_norm) {
      const float s = 1.f / (sqrtf(a * a + b * b) + 1e-8f);
      a *= s;
      b *= s;
    }
    wa[n] = a;
    wb[n] = b;
  }
  __syncthreads();
  if (n >= d) return;
  float acc = 0.f;
  for (int k = 0; k <= half; ++k) {
    const float a = wa[k], b = wb[k];
    if (k == 0) {
      acc += a;
    } else if (k == half) {
      acc += (n & 1) ? -a : a;
    } else {
      const int m = (k * n) & (d - 1);      // d is a power of two
      acc += 2.f * (a * cs_t[m] - b * sn_t[m]);
    }
  }
  taps[f * d + n] = acc / (float)d;
}

// ---- register-tiled circular convolution / correlation -----------------------------------
// A CTA owns 32 item rows (lane = row) and its 8 warps split the D outputs of every row into
// tiles of TN = D / 8. With one output per thread every FMA needed two shared-memory loads (the
// kernels ran at 16 % of the FMA pipe, LSU-bound); here a thread keeps TN accumulators and a
// sliding window of the second operand in registers: per 4 taps it issues two 128-bit loads
// (4 values of `a`, the 4 values of `b` that enter the window) for 4 * TN FMAs. Rows are staged
// with a pitch of D + 4 floats, so the 8 lanes of a 128-bit shared-memory phase hit distinct banks;
// taps are read at one address by the whole warp (broadcast).
//   conv: acc[j] += sum_s a[s] * b[(n0 + j - s) mod D]
//   corr: acc[j] += sum_s a[s] * b[(s - n0 - j) mod D]      (adjoint of conv w.r.t. its first operand)
// Summation runs over s ascending for every output, like the one-output-per-thread version.
constexpr int kRows = 32;        // rows per CTA
constexpr int kThreads = 256;

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
template <int E>
__device__ __forceinline__ float elem(const float4 &v) {
  if constexpr (E == 0) return v.x;
  if constexpr (E == 1) return v.y;
  if constexpr (E == 2) return v.z;
  return v.w;
}

template <int TN, bool CORR, int G, int CC, int I, int J>
__device__ __forceinline__ void fma_one(float (&acc)[TN], const float4 &av, const float4 (&w)[G]) {
  constexpr int q = CORR ? I - J + TN : J - I + 4;          // window element of this (tap, output) pair
  constexpr int grp = q / 4, e = q % 4;
  constexpr int slot = CORR ? (CC + grp) % G : ((grp - CC) % G + G) % G;
  acc[J] = fmaf(elem<I>(av), elem<e>(w[slot]), acc[J]);
}
template <int TN, bool CORR, int G, int CC, int I, int J>
struct FmaLoop {
  __device__ __forceinline__ static void run(float (&acc)[TN], const float4 &av, const float4 (&w)[G]) {
    fma_one<TN, CORR, G, CC, I, J>(acc, av, w);
    if constexpr (J + 1 < TN) FmaLoop<TN, CORR, G, CC, I, J + 1>::run(acc, av, w);
    else if constexpr (I + 1 < 4) FmaLoop<TN, CORR, G, CC, I + 1, 0>::run(acc, av, w);
  }
};
template <int D, int TN, bool CORR, int G, int CC>
struct ChunkLoop {
  __device__ __forceinline__ static void run(float (&acc)[TN], const float *a, const float *b, int b0, int co,
                                             float4 (&w)[G], float4 &av) {
    constexpr int LIVE = TN / 4 + 1;
    const int c = co + CC;
    // the group that enters the window at chunk c + 1 and the taps of chunk c + 1, both in flight
    // under the FMAs of chunk c (always loaded: the indices wrap harmlessly)
    constexpr int slot = CORR ? (CC + LIVE) % G : ((-(CC + 1)) % G + G) % G;
    const int idx = CORR ? (b0 + 4 * (c + LIVE)) & (D - 1) : (b0 - 4 * (c + 1)) & (D - 1);
    const float4 nw = lds4(b + idx);
    const float4 an = lds4(a + ((4 * (c + 1)) & (D - 1)));
    FmaLoop<TN, CORR, G, CC, 0, 0>::run(acc, av, w);
    w[slot] = nw;
    av = an;
    if constexpr (CC + 1 < G) ChunkLoop<D, TN, CORR, G, CC + 1>::run(acc, a, b, b0, co, w, av);
  }
};

template <int D, int TN, bool CORR>
__device__ __forceinline__ void circ_tile(float (&acc)[TN], const float *a, const float *b, int n0) {
  constexpr int LIVE = TN / 4 + 1;            // float4 groups alive in the window
  constexpr int G = LIVE < 4 ? 4 : 8;         // register slots: the window rotates through them with period G
  static_assert((D / 4) % G == 0 && G > LIVE, "window rotation does not close");
  float4 w[G];
  // index (mod D, multiple of 4) of window element 0 at chunk 0
  const int b0 = CORR ? (2 * D - n0 - TN) & (D - 1) : (n0 - 4 + D) & (D - 1);
#pragma unroll
  for (int g = 0; g < G; ++g) w[g] = lds4(b + ((b0 + 4 * g) & (D - 1)));      // slots LIVE.. are overwritten before use
  float4 av = lds4(a);
#pragma unroll 1
  for (int co = 0; co < D / 4; co += G) ChunkLoop<D, TN, CORR, G, 0>::run(acc, a, b, b0, co, w, av);
}

template <int TN>
__device__ __forceinline__ void zero(float (&acc)[TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) acc[j] = 0.f;
}

// rows [row0, row0 + 32) of a row-major [n_rows, D] array -> shared memory, pitch D + 4 (rows past the end: zeros)
template <int D>
__device__ __forceinline__ void stage_rows(float *dst, const float *__restrict__ src, int row0, int n_rows, float scale) {
  constexpr int V = D / 4;
  for (int t = threadIdx.x; t < kRows * V; t += kThreads) {
    const int r = t / V, c4 = t % V;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n_rows) v = ldg4(src + (size_t)(row0 + r) * D + c4 * 4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4 *>(dst + r * (D + 4) + c4 * 4) = v;
  }
}
template <int D, int TN>
__device__ __forceinline__ void store_tile(float *dst_row, int n0, const float (&acc)[TN], float scale = 1.f) {
#pragma unroll
  for (int j = 0; j < TN; j += 4)
    *reinterpret_cast<float4 *>(dst_row + n0 + j) =
        make_float4(acc[j] * scale, acc[j + 1] * scale, acc[j + 2] * scale, acc[j + 3] * scale);
}

template <int D>
__global__ void __launch_bounds__(kThreads)
spectral_fwd_kernel(const float *__restrict__ img, const float *__restrict__ txt, int n_rows,
                    const float *__restrict__ taps, float *__restrict__ ic, float *__restrict__ tc,
                    float *__restrict__ fc) {
  constexpr int TN = D / 8, P = D + 4;
  extern __shared__ __align__(16) float smem[];
  float *sh = smem;                           // [3][D] taps
  float *sx = sh + 3 * D, *st = sx + kRows * P, *sc = st + kRows * P;
  for (int t = threadIdx.x; t < 3 * D; t += kThreads) sh[t] = taps[t];
  const int row0 = blockIdx.x * kRows;
  stage_rows<D>(sx, img, row0, n_rows, 1.f);
  stage_rows<D>(st, txt, row0, n_rows, 1.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, n0 = (threadIdx.x >> 5) * TN, row = row0 + lane;
  const bool live = row < n_rows;
  const float *x = sx + lane * P, *t_ = st + lane * P;
  float acc[TN];
  zero(acc);
  circ_tile<D, TN, false>(acc, x, sh, n0);                  // image_conv = x (*) h_img
  if (live) store_tile<D, TN>(ic + (size_t)row * D, n0, acc);
  zero(acc);
  circ_tile<D, TN, false>(acc, t_, sh + D, n0);             // text_conv = t (*) h_txt
  if (live) store_tile<D, TN>(tc + (size_t)row * D, n0, acc);
  zero(acc);
  circ_tile<D, TN, false>(acc, t_, x, n0);                  // c = t (*) x
  store_tile<D, TN>(sc + lane * P, n0, acc);
  __syncthreads();
  zero(acc);
  circ_tile<D, TN, false>(acc, sc + lane * P, sh + 2 * D, n0);   // fusion_conv = (c (*) h_fus) / sqrt(d)
  if (live) store_tile<D, TN>(fc + (size_t)row * D, n0, acc, rsqrtf((float)D));
}

// sum of v over the 32 rows of the CTA (lanes), added to dst by lane 0
template <int TN>
__device__ __forceinline__ void reduce_rows_atomic(float *dst, const float (&acc)[TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(dst + j, v);
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads)
spectral_bwd_kernel(const float *__restrict__ img, const float *__restrict__ txt, int n_rows,
                    const float *__restrict__ taps, const float *__restrict__ g_ic,
                    const float *__restrict__ g_tc, const float *__restrict__ g_fc,
                    float *__restrict__ d_img, float *__restrict__ d_txt, float *__restrict__ dh) {
  constexpr int TN = D / 8, P = D + 4;
  extern __shared__ __align__(16) float smem[];
  float *sh = smem;
  float *sx = sh + 3 * D, *st = sx + kRows * P, *sc = st + kRows * P;
  float *gi = sc + kRows * P, *gt = gi + kRows * P, *gf = gt + kRows * P, *gc = gf + kRows * P;
  for (int t = threadIdx.x; t < 3 * D; t += kThreads) sh[t] = taps[t];
  const int row0 = blockIdx.x * kRows;
  stage_rows<D>(sx, img, row0, n_rows, 1.f);
  stage_rows<D>(st, txt, row0, n_rows, 1.f);
  stage_rows<D>(gi, g_ic, row0, n_rows, 1.f);
  stage_rows<D>(gt, g_tc, row0, n_rows, 1.f);
  stage_rows<D>(gf, g_fc, row0, n_rows, rsqrtf((float)D));     // fold the 1/sqrt(d) of the fusion path
  __syncthreads();
  const int lane = threadIdx.x & 31, n0 = (threadIdx.x >> 5) * TN, row = row0 + lane;
  const bool live = row < n_rows;
  const int o = lane * P;
  float acc[TN], acc2[TN];
  zero(acc);
  circ_tile<D, TN, false>(acc, st + o, sx + o, n0);         // c = t (*) x
  store_tile<D, TN>(sc + o, n0, acc);
  zero(acc);
  circ_tile<D, TN, true>(acc, gf + o, sh + 2 * D, n0);      // dL/dc
  store_tile<D, TN>(gc + o, n0, acc);
  __syncthreads();
  // tap gradients: dh_f[j] = sum over rows of corr(g, a)[j]
  zero(acc);
  circ_tile<D, TN, true>(acc, gi + o, sx + o, n0);
  reduce_rows_atomic<TN>(dh + n0, acc);
  zero(acc);
  circ_tile<D, TN, true>(acc, gt + o, st + o, n0);
  reduce_rows_atomic<TN>(dh + D + n0, acc);
  zero(acc);
  circ_tile<D, TN, true>(acc, gf + o, sc + o, n0);
  reduce_rows_atomic<TN>(dh + 2 * D + n0, acc);
  // d_img = corr(g_ic, h_img) + corr(dL/dc, t);  d_txt = corr(g_tc, h_txt) + corr(dL/dc, x)
  zero(acc);
  zero(acc2);
  circ_tile<D, TN, true>(acc, gi + o, sh, n0);
  circ_tile<D, TN, true>(acc2, gc + o, st + o, n0);
#pragma unroll
  for (int j = 0; j < TN; ++j) acc[j] += acc2[j];
  if (live) store_tile<D, TN>(d_img + (size_t)row * D, n0, acc);
  zero(acc);
  zero(acc2);
  circ_tile<D, TN, true>(acc, gt + o, sh + D, n0);
  circ_tile<D, TN, true>(acc2, gc + o, sx + o, n0);
#pragma unroll
  for (int j = 0; j < TN; ++j) acc[j] += acc2[j];
  if (live) store_tile<D, TN>(d_txt + (size_t)row * D, n0, acc);
}

template <int D>
constexpr size_t spectral_smem(int arrays) { return (3 * D + (size_t)arrays * kRows * (D + 4)) * sizeof(float); }

// dh (taps) -> raw weight gradients through irfft_backward and the unit-magnitude map.
__global__ void spectral_weight_bwd_kernel(const float *__restrict__ w_img, const float *__restrict__ w_txt,
                                           const float *__restrict__ w_fus, int d, int weight_norm,
                                           const float *__restrict__ dh, float *__restrict__ d_w_img,
                                           float *__restrict__ d_w_txt, float *__restrict__ d_w_fus) {
  __shared__ float cs_t[128], sn_t[128], sg[128];
  const int f = blockIdx.x, k = threadIdx.x;
  const int half = d / 2;
  const float *w = f == 0 ? w_img : (f == 1 ? w_txt : w_fus);
  float *dw = f == 0 ? d_w_img : (f == 1 ? d_w_txt : d_w_fus);
  for (int n = k; n < d; n += blockDim.x) {     // twiddle table + this filter's tap gradients
    sincospif(2.f * (float)n / (float)d, &sn_t[n], &cs_t[n]);
    sg[n] = dh[f * d + n];
  }
  __syncthreads();
  if (k > half) return;
  const float *g = sg;
  float da = 0.f, db = 0.f;   // gradient w.r.t. the (normalised) real / imaginary parts
  for (int n = 0; n < d; ++n) {
    if (k == 0) {
      da += g[n];
    } else if (k == half) {
      da += (n & 1) ? -g[n] : g[n];
    } else {
      const int m = (k * n) & (d - 1);      // d is a power of two
      da += 2.f * g[n] * cs_t[m];
      db -= 2.f * g[n] * sn_t[m];
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
  spectral_taps_kernel<<<3, d < 64 ? 64 : d, 0, stream>>>(w_img, w_txt, w_fus, d, weight_norm, taps_ws);
  MMREC_CHECK_LAUNCH("spectral_taps_kernel");
  const int blocks = (n_rows + kRows - 1) / kRows;
#define MMREC_SPEC_FWD(D_)                                                                                       \
  do {                                                                                                           \
    MMREC_CUDA(cudaFuncSetAttribute(spectral_fwd_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    (int)spectral_smem<D_>(3)));                                                 \
    spectral_fwd_kernel<D_><<<blocks, kThreads, spectral_smem<D_>(3), stream>>>(img, txt, n_rows, taps_ws,       \
                                                                                img_conv, txt_conv, fus_conv);   \
  } while (0)
  if (d == 32) MMREC_SPEC_FWD(32);
  else if (d == 64) MMREC_SPEC_FWD(64);
  else MMREC_SPEC_FWD(128);
#undef MMREC_SPEC_FWD
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
  const int blocks = (n_rows + kRows - 1) / kRows;
#define MMREC_SPEC_BWD(D_)                                                                                       \
  do {                                                                                                           \
    MMREC_CUDA(cudaFuncSetAttribute(spectral_bwd_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    (int)spectral_smem<D_>(7)));                                                 \
    spectral_bwd_kernel<D_><<<blocks, kThreads, spectral_smem<D_>(7), stream>>>(                                 \
        img, txt, n_rows, taps_ws, g_img_conv, g_txt_conv, g_fus_conv, d_img, d_txt, dh_ws);                     \
  } while (0)
  if (d == 32) MMREC_SPEC_BWD(32);
  else if (d == 64) MMREC_SPEC_BWD(64);
  else MMREC_SPEC_BWD(128);
#undef MMREC_SPEC_BWD
  MMREC_CHECK_LAUNCH("spectral_bwd_kernel");
  spectral_weight_bwd_kernel<<<3, d < 64 ? 64 : d, 0, stream>>>(w_img, w_txt, w_fus, d, weight_norm, dh_ws, d_w_img, d_w_txt, d_w_fus);
  MMREC_CHECK_LAUNCH("spectral_weight_bwd_kernel");
  return MMREC_OK;
}
