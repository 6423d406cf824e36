// MGCN's two-way attention fuser (mgcn.py:188-205) as one launch each way -- include/mmrec_b200.h:
//   mmrec_mgcn_fuse_fwd_f32 / mmrec_mgcn_fuse_bwd_f32
//
//   a_m   = h_m . w2                      h_m = tanh(Linear(d,d)(E_m)) from the d x d layer kernel,
//                                          w2 = query_common.2.weight [1, d] (the Linear(d, 1))
//   (w0, w1) = softmax(a_img, a_txt)
//   common = w0 E_img + w1 E_txt
//   side   = (P_img (E_img - common) + P_txt (E_txt - common) + common) / 3
//   all    = content + side
//
// The reference runs this as ~20 elementwise / reduction kernels plus a cuBLAS [N, d] x [d, 1]
// product each way. Here a sub-warp of d/4 lanes owns a row (one float4 per lane and operand), the
// two row dots are butterfly sums inside the sub-warp, every operand is read once and every result
// written once: 8 reads + 2 writes of [N, d] forward (HBM-bound: 4 d (8 + 2) bytes per row).
// Backward: one launch for the seven row gradients; the gradient of w2 (a column reduction over
// all rows) leaves as per-CTA partial sums [n_blocks, d] in fixed order and is finished by
// mmrec_colsum_f32 -- no float atomics, bit-reproducible.
#include "common.cuh"

using namespace mmrec;

namespace mmrec {
namespace {

constexpr int kFuseThreads = 256;

__device__ __forceinline__ float4 f4_sub(const float4 &a, const float4 &b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}

template <int LANES>
__global__ void __launch_bounds__(kFuseThreads)
mgcn_fuse_fwd_kernel(const float *__restrict__ Hi, const float *__restrict__ Ht, const float *__restrict__ w2,
                     const float *__restrict__ Ei, const float *__restrict__ Et, const float *__restrict__ Pi,
                     const float *__restrict__ Pt, const float *__restrict__ C, int n, float *__restrict__ att,
                     float *__restrict__ side, float *__restrict__ all) {
  constexpr int D = LANES * 4;
  constexpr int ROWS = kFuseThreads / LANES;
  const int sub = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const int r = blockIdx.x * ROWS + sub;
  if (r >= n) return;                                   // whole sub-warps leave together
  const size_t o = (size_t)r * D + lane * 4;
  const float4 w = ldg4(w2 + lane * 4);
  const float4 hi = ldg4(Hi + o), ht = ldg4(Ht + o);
  const float4 ei = ldg4(Ei + o), et = ldg4(Et + o), pi = ldg4(Pi + o), pt = ldg4(Pt + o), c = ldg4(C + o);
  const float ai = group_sum<LANES>(dot4(hi, w));
  const float at = group_sum<LANES>(dot4(ht, w));
  const float m = fmaxf(ai, at);
  const float xi = expf(ai - m), xt = expf(at - m);
  const float inv = 1.f / (xi + xt);
  const float w0 = xi * inv, w1 = xt * inv;
  if (lane == 0) *reinterpret_cast<float2 *>(att + 2 * (size_t)r) = make_float2(w0, w1);
  const float4 com = make_float4(w0 * ei.x + w1 * et.x, w0 * ei.y + w1 * et.y, w0 * ei.z + w1 * et.z,
                                 w0 * ei.w + w1 * et.w);
  const float4 di = f4_sub(ei, com), dt = f4_sub(et, com);
  constexpr float third = 1.f / 3.f;
  const float4 s = make_float4((pi.x * di.x + pt.x * dt.x + com.x) * third, (pi.y * di.y + pt.y * dt.y + com.y) * third,
                               (pi.z * di.z + pt.z * dt.z + com.z) * third, (pi.w * di.w + pt.w * dt.w + com.w) * third);
  *reinterpret_cast<float4 *>(side + o) = s;
  *reinterpret_cast<float4 *>(all + o) = make_float4(c.x + s.x, c.y + s.y, c.z + s.z, c.w + s.w);
}

template <int LANES>
__global__ void __launch_bounds__(kFuseThreads)
mgcn_fuse_bwd_kernel(const float *__restrict__ g_all, const float *__restrict__ g_side, const float *__restrict__ Hi,
                     const float *__restrict__ Ht, const float *__restrict__ w2, const float *__restrict__ Ei,
                     const float *__restrict__ Et, const float *__restrict__ Pi, const float *__restrict__ Pt,
                     const float *__restrict__ att, int n, float *__restrict__ dHi, float *__restrict__ dHt,
                     float *__restrict__ dEi, float *__restrict__ dEt, float *__restrict__ dPi,
                     float *__restrict__ dPt, float *__restrict__ dC, float *__restrict__ dw2_partial) {
  constexpr int D = LANES * 4;
  constexpr int ROWS = kFuseThreads / LANES;
  __shared__ float4 red[kFuseThreads];
  const int sub = threadIdx.x / LANES, lane = threadIdx.x % LANES;
  const int r = blockIdx.x * ROWS + sub;
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < n) {
    const size_t o = (size_t)r * D + lane * 4;
    const float4 ga = g_all ? ldg4(g_all + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 gsd = g_side ? ldg4(g_side + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr float third = 1.f / 3.f;
    const float4 gs = make_float4((ga.x + gsd.x) * third, (ga.y + gsd.y) * third, (ga.z + gsd.z) * third,
                                  (ga.w + gsd.w) * third);                  // d side / 3
    const float4 w = ldg4(w2 + lane * 4);
    const float4 hi = ldg4(Hi + o), ht = ldg4(Ht + o);
    const float4 ei = ldg4(Ei + o), et = ldg4(Et + o), pi = ldg4(Pi + o), pt = ldg4(Pt + o);
    const float2 a = __ldg(reinterpret_cast<const float2 *>(att + 2 * (size_t)r));
    const float w0 = a.x, w1 = a.y;
    const float4 com = make_float4(w0 * ei.x + w1 * et.x, w0 * ei.y + w1 * et.y, w0 * ei.z + w1 * et.z,
                                   w0 * ei.w + w1 * et.w);
    const float4 di = f4_sub(ei, com), dt = f4_sub(et, com);
    *reinterpret_cast<float4 *>(dC + o) = ga;
    *reinterpret_cast<float4 *>(dPi + o) = make_float4(gs.x * di.x, gs.y * di.y, gs.z * di.z, gs.w * di.w);
    *reinterpret_cast<float4 *>(dPt + o) = make_float4(gs.x * dt.x, gs.y * dt.y, gs.z * dt.z, gs.w * dt.w);
    // d common = gs (1 - P_img - P_txt)
    const float4 dc = make_float4(gs.x * (1.f - pi.x - pt.x), gs.y * (1.f - pi.y - pt.y), gs.z * (1.f - pi.z - pt.z),
                                  gs.w * (1.f - pi.w - pt.w));
    *reinterpret_cast<float4 *>(dEi + o) = make_float4(gs.x * pi.x + w0 * dc.x, gs.y * pi.y + w0 * dc.y,
                                                       gs.z * pi.z + w0 * dc.z, gs.w * pi.w + w0 * dc.w);
    *reinterpret_cast<float4 *>(dEt + o) = make_float4(gs.x * pt.x + w1 * dc.x, gs.y * pt.y + w1 * dc.y,
                                                       gs.z * pt.z + w1 * dc.z, gs.w * pt.w + w1 * dc.w);
    const float dw0 = group_sum<LANES>(dot4(dc, ei));
    const float dw1 = group_sum<LANES>(dot4(dc, et));
    const float mix = w0 * dw0 + w1 * dw1;                      // softmax backward
    const float dai = w0 * (dw0 - mix), dat = w1 * (dw1 - mix);
    *reinterpret_cast<float4 *>(dHi + o) = make_float4(dai * w.x, dai * w.y, dai * w.z, dai * w.w);
    *reinterpret_cast<float4 *>(dHt + o) = make_float4(dat * w.x, dat * w.y, dat * w.z, dat * w.w);
    dw = make_float4(dai * hi.x + dat * ht.x, dai * hi.y + dat * ht.y, dai * hi.z + dat * ht.z,
                     dai * hi.w + dat * ht.w);
  }
  // column partial of this CTA's rows for d w2, fixed order over the sub-warps
  red[threadIdx.x] = dw;
  __syncthreads();
  if (threadIdx.x < LANES) {
    float4 s = red[threadIdx.x];
#pragma unroll
    for (int k = 1; k < ROWS; ++k) {
      const float4 v = red[k * LANES + threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4 *>(dw2_partial + (size_t)blockIdx.x * D + threadIdx.x * 4) = s;
  }
}

inline int fuse_blocks(int n, int d) { return (n + (kFuseThreads / (d / 4)) - 1) / (kFuseThreads / (d / 4)); }

}  // namespace
}  // namespace mmrec

extern "C" int mmrec_mgcn_fuse_supported(int32_t d) { return d == 32 || d == 64 || d == 128; }

extern "C" int32_t mmrec_mgcn_fuse_bwd_blocks(int32_t n_rows, int32_t d) {
  return mmrec_mgcn_fuse_supported(d) && n_rows > 0 ? fuse_blocks(n_rows, d) : 0;
}

extern "C" int mmrec_mgcn_fuse_fwd_f32(const float *Hi, const float *Ht, const float *w2, const float *Ei,
                                       const float *Et, const float *Pi, const float *Pt, const float *content,
                                       int32_t n_rows, int32_t d, float *att, float *side, float *all,
                                       void *stream_) {
  MMREC_REQUIRE(Hi && Ht && w2 && Ei && Et && Pi && Pt && content && att && side && all, MMREC_E_BADARG,
                "mgcn_fuse_fwd: null pointer");
  MMREC_REQUIRE(n_rows > 0 && mmrec_mgcn_fuse_supported(d), MMREC_E_BADARG,
                "mgcn_fuse_fwd: need n_rows > 0 and d in {32, 64, 128} (got %d, %d)", n_rows, d);
  MMREC_REQUIRE(aligned16(Hi) && aligned16(Ht) && aligned16(w2) && aligned16(Ei) && aligned16(Et) && aligned16(Pi) &&
                    aligned16(Pt) && aligned16(content) && aligned16(side) && aligned16(all) &&
                    (reinterpret_cast<uintptr_t>(att) & 7u) == 0,
                MMREC_E_ALIGN, "mgcn_fuse_fwd: operands must be 16-byte aligned");
  auto s = (cudaStream_t)stream_;
  const int blocks = fuse_blocks(n_rows, d);
  if (d == 32) mgcn_fuse_fwd_kernel<8><<<blocks, kFuseThreads, 0, s>>>(Hi, Ht, w2, Ei, Et, Pi, Pt, content, n_rows, att, side, all);
  else if (d == 64) mgcn_fuse_fwd_kernel<16><<<blocks, kFuseThreads, 0, s>>>(Hi, Ht, w2, Ei, Et, Pi, Pt, content, n_rows, att, side, all);
  else mgcn_fuse_fwd_kernel<32><<<blocks, kFuseThreads, 0, s>>>(Hi, Ht, w2, Ei, Et, Pi, Pt, content, n_rows, att, side, all);
  MMREC_CHECK_LAUNCH("mgcn_fuse_fwd_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_mgcn_fuse_bwd_f32(const float *g_all, const float *g_side, const float *Hi, const float *Ht,
                                       const float *w2, const float *Ei, const float *Et, const float *Pi,
                                       const float *Pt, const float *att, int32_t n_rows, int32_t d, float *dHi,
                                       float *dHt, float *dEi, float *dEt, float *dPi, float *dPt, float *dC,
                                       float *dw2_partial, float *dw2, void *stream_) {
  MMREC_REQUIRE(g_all || g_side, MMREC_E_BADARG, "mgcn_fuse_bwd: both output gradients are null");
  MMREC_REQUIRE(Hi && Ht && w2 && Ei && Et && Pi && Pt && att && dHi && dHt && dEi && dEt && dPi && dPt && dC &&
                    dw2_partial && dw2, MMREC_E_BADARG, "mgcn_fuse_bwd: null pointer");
  MMREC_REQUIRE(n_rows > 0 && mmrec_mgcn_fuse_supported(d), MMREC_E_BADARG,
                "mgcn_fuse_bwd: need n_rows > 0 and d in {32, 64, 128} (got %d, %d)", n_rows, d);
  MMREC_REQUIRE((!g_all || aligned16(g_all)) && (!g_side || aligned16(g_side)) && aligned16(Hi) && aligned16(Ht) &&
                    aligned16(w2) && aligned16(Ei) && aligned16(Et) && aligned16(Pi) && aligned16(Pt) &&
                    aligned16(dHi) && aligned16(dHt) && aligned16(dEi) && aligned16(dEt) && aligned16(dPi) &&
                    aligned16(dPt) && aligned16(dC) && aligned16(dw2_partial) && aligned16(dw2),
                MMREC_E_ALIGN, "mgcn_fuse_bwd: operands must be 16-byte aligned");
  auto s = (cudaStream_t)stream_;
  const int blocks = fuse_blocks(n_rows, d);
#define MMREC_FUSE_BWD(L)                                                                                       \
  mgcn_fuse_bwd_kernel<L><<<blocks, kFuseThreads, 0, s>>>(g_all, g_side, Hi, Ht, w2, Ei, Et, Pi, Pt, att, n_rows, \
                                                          dHi, dHt, dEi, dEt, dPi, dPt, dC, dw2_partial)
  if (d == 32) MMREC_FUSE_BWD(8);
  else if (d == 64) MMREC_FUSE_BWD(16);
  else MMREC_FUSE_BWD(32);
#undef MMREC_FUSE_BWD
  MMREC_CHECK_LAUNCH("mgcn_fuse_bwd_kernel");
  return mmrec_colsum_f32(dw2_partial, blocks, d, dw2, stream_);
}
