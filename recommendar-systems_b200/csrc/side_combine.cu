// Row part of SMORE's modality-aware preference module (smore.py:321-341) for the widths whose
// seven d x d layers run as tcgen05 GEMM launches (d = 128: see mmrec_linear_act_tc_f32) instead of
// inside the fused mma.sync kernel of side_net.cu -- include/mmrec_b200.h:
//   mmrec_smore_combine_fwd_f32 / mmrec_smore_combine_bwd_f32
//
//   sv = softmax_d(zv), st = softmax_d(zt)          zv = query_v(fusion), zt = query_t(fusion)
//   side = (p_i * sv * V + p_t * st * T + p_f * F) / 3,   p_m = g_m * mask_m   (nn.Dropout multipliers)
//   all  = C + side
//
// torch runs this as ~20 elementwise / softmax kernels forward and ~35 backward over [N, d] tensors
// (1.6 ms of the 14 ms Clothing step). Here a sub-warp of d/4 lanes owns a row (one float4 per lane
// and operand): every operand is read once, the two softmax reductions are butterflies inside the
// sub-warp, the backward recomputes the softmax from zv / zt instead of saving it. HBM-bound:
// forward 9 (+3 masks) reads + 2 writes, backward 11 (+3) reads + 9 writes of [N, d].
#include <algorithm>

#include "common.cuh"

using namespace mmrec;

namespace mmrec {
namespace {

constexpr int kCombThreads = 256;

template <int W>
__device__ __forceinline__ float sub_max(float v) {
  unsigned mask = 0xffffffffu;
  if constexpr (W < 32) {
    const unsigned lane = threadIdx.x & 31u;
    mask = ((1u << W) - 1u) << (W * (lane / W));
  }
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, W));
  return v;
}

template <int LANES>
__device__ __forceinline__ float4 row_softmax(const float4 &z) {
  const float m = sub_max<LANES>(fmaxf(fmaxf(z.x, z.y), fmaxf(z.z, z.w)));
  float4 e = make_float4(fast_exp(z.x - m), fast_exp(z.y - m), fast_exp(z.z - m), fast_exp(z.w - m));
  const float inv = fast_rcp(group_sum<LANES>((e.x + e.y) + (e.z + e.w)));
  e.x *= inv; e.y *= inv; e.z *= inv; e.w *= inv;
  return e;
}

__device__ __forceinline__ float4 mul4(const float4 &a, const float4 &b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 scale4(const float4 &a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

struct CombIn {
  const float *zv, *zt, *V, *T, *F, *C, *gi, *gt, *gf;
  const float *mask;     // [3, n, d] dropout multipliers or NULL
  DropSpec drop;         // mask == NULL and drop.p > 0: multipliers generated here (common.cuh)
};

// dropout multipliers of (plane g, row r, columns 4*lane ..) -- the [3, n, d] index space of side_net.cu
template <int LANES>
__device__ __forceinline__ float4 comb_mask(const CombIn &in, uint64_t key, int g, int r, int lane, int n, size_t o) {
  if (in.mask != nullptr) return ldg4(in.mask + (size_t)g * n * (LANES * 4) + o);
  if (in.drop.p > 0.f) return drop_mask4(key, drop_row4(in.drop, g, r, n, LANES) + lane, in.drop.p);
  return make_float4(1.f, 1.f, 1.f, 1.f);
}

template <int LANES>
__global__ void __launch_bounds__(kCombThreads)
smore_combine_fwd_kernel(const CombIn in, int n, float *__restrict__ side, float *__restrict__ all) {
  constexpr int D = LANES * 4, ROWS = kCombThreads / LANES;
  const int r = blockIdx.x * ROWS + threadIdx.x / LANES, lane = threadIdx.x % LANES;
  if (r >= n) return;
  const size_t o = (size_t)r * D + lane * 4;
  const float4 sv = row_softmax<LANES>(ldg4(in.zv + o)), st = row_softmax<LANES>(ldg4(in.zt + o));
  float4 pi = ldg4(in.gi + o), pt = ldg4(in.gt + o), pf = ldg4(in.gf + o);
  if (in.mask != nullptr || in.drop.p > 0.f) {
    const uint64_t key = in.drop.p > 0.f ? drop_stream(in.drop) : 0ull;
    pi = mul4(pi, comb_mask<LANES>(in, key, 0, r, lane, n, o));
    pt = mul4(pt, comb_mask<LANES>(in, key, 1, r, lane, n, o));
    pf = mul4(pf, comb_mask<LANES>(in, key, 2, r, lane, n, o));
  }
  const float4 av = mul4(sv, ldg4(in.V + o)), at = mul4(st, ldg4(in.T + o)), f = ldg4(in.F + o), c = ldg4(in.C + o);
  constexpr float third = 1.f / 3.f;
  const float4 s = make_float4((pi.x * av.x + pt.x * at.x + pf.x * f.x) * third, (pi.y * av.y + pt.y * at.y + pf.y * f.y) * third,
                               (pi.z * av.z + pt.z * at.z + pf.z * f.z) * third, (pi.w * av.w + pt.w * at.w + pf.w * f.w) * third);
  *reinterpret_cast<float4 *>(side + o) = s;
  *reinterpret_cast<float4 *>(all + o) = make_float4(c.x + s.x, c.y + s.y, c.z + s.z, c.w + s.w);
}

struct CombGrad {
  float *dzv, *dzt, *dV, *dT, *dF, *dC, *dgi, *dgt, *dgf;
};

template <int LANES>
__global__ void __launch_bounds__(kCombThreads)
smore_combine_bwd_kernel(const CombIn in, const float *__restrict__ g_all, const float *__restrict__ g_side, int n,
                         const CombGrad out) {
  constexpr int D = LANES * 4, ROWS = kCombThreads / LANES;
  const int r = blockIdx.x * ROWS + threadIdx.x / LANES, lane = threadIdx.x % LANES;
  if (r >= n) return;
  const size_t o = (size_t)r * D + lane * 4;
  const float4 ga = g_all ? ldg4(g_all + o) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 gs = g_side ? ldg4(g_side + o) : make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr float third = 1.f / 3.f;
  const float4 g = make_float4((ga.x + gs.x) * third, (ga.y + gs.y) * third, (ga.z + gs.z) * third, (ga.w + gs.w) * third);
  *reinterpret_cast<float4 *>(out.dC + o) = ga;
  const float4 sv = row_softmax<LANES>(ldg4(in.zv + o)), st = row_softmax<LANES>(ldg4(in.zt + o));
  const float4 gi = ldg4(in.gi + o), gt = ldg4(in.gt + o), gf = ldg4(in.gf + o);
  float4 mi = make_float4(1.f, 1.f, 1.f, 1.f), mt = mi, mf = mi;
  if (in.mask != nullptr || in.drop.p > 0.f) {
    const uint64_t key = in.drop.p > 0.f ? drop_stream(in.drop) : 0ull;
    mi = comb_mask<LANES>(in, key, 0, r, lane, n, o);
    mt = comb_mask<LANES>(in, key, 1, r, lane, n, o);
    mf = comb_mask<LANES>(in, key, 2, r, lane, n, o);
  }
  const float4 V = ldg4(in.V + o), T = ldg4(in.T + o), f = ldg4(in.F + o);
  const float4 pi = mul4(gi, mi), pt = mul4(gt, mt), pf = mul4(gf, mf);
  // gates: d g_m = g * agg_m * mask_m
  *reinterpret_cast<float4 *>(out.dgi + o) = mul4(mul4(g, mul4(sv, V)), mi);
  *reinterpret_cast<float4 *>(out.dgt + o) = mul4(mul4(g, mul4(st, T)), mt);
  *reinterpret_cast<float4 *>(out.dgf + o) = mul4(mul4(g, f), mf);
  *reinterpret_cast<float4 *>(out.dF + o) = mul4(g, pf);
  // aggregated modalities: d agg = g * p;  dV = d agg * s;  ds = d agg * V;  dz = s * (ds - <ds, s>)
  const float4 dav = mul4(g, pi), dat = mul4(g, pt);
  *reinterpret_cast<float4 *>(out.dV + o) = mul4(dav, sv);
  *reinterpret_cast<float4 *>(out.dT + o) = mul4(dat, st);
  const float4 dsv = mul4(dav, V), dst = mul4(dat, T);
  const float dotv = group_sum<LANES>(dot4(dsv, sv)), dott = group_sum<LANES>(dot4(dst, st));
  *reinterpret_cast<float4 *>(out.dzv + o) =
      make_float4(sv.x * (dsv.x - dotv), sv.y * (dsv.y - dotv), sv.z * (dsv.z - dotv), sv.w * (dsv.w - dotv));
  *reinterpret_cast<float4 *>(out.dzt + o) =
      make_float4(st.x * (dst.x - dott), st.y * (dst.y - dott), st.z * (dst.z - dott), st.w * (dst.w - dott));
}

// the multipliers a *_drop_* call generates, written out: [planes, n, d]
__global__ void __launch_bounds__(256)
dropout_mask_kernel(float4 *__restrict__ out, size_t n4, const DropSpec drop) {
  const uint64_t key = drop_stream(drop);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    out[i] = drop.p > 0.f ? drop_mask4(key, i, drop.p) : make_float4(1.f, 1.f, 1.f, 1.f);
}

inline int comb_drop(const MmrecDropout *h, DropSpec &D) {
  D = DropSpec{nullptr, 0ull, 0.f, nullptr, 0};
  if (h == nullptr) return MMREC_OK;
  MMREC_REQUIRE(h->p >= 0.f && h->p < 1.f, MMREC_E_BADARG, "dropout: p must be in [0, 1) (got %g)", (double)h->p);
  MMREC_REQUIRE(h->row_ids == nullptr || h->n_total > 0, MMREC_E_BADARG, "dropout: row_ids needs n_total");
  D = DropSpec{h->counter, h->seed, h->p, reinterpret_cast<const long long *>(h->row_ids), h->n_total};
  return MMREC_OK;
}

inline bool comb_ok(int d) { return d == 32 || d == 64 || d == 128; }
inline int comb_blocks(int n, int d) { const int rows = kCombThreads / (d / 4); return (n + rows - 1) / rows; }

}  // namespace
}  // namespace mmrec

extern "C" int mmrec_smore_combine_supported(int32_t d) { return comb_ok(d); }

extern "C" int mmrec_dropout_mask_f32(float *out, int32_t planes, int32_t n, int32_t d, const MmrecDropout *drop,
                                      void *stream_) {
  MMREC_REQUIRE(out && drop && planes > 0 && n >= 0 && d > 0 && d % 4 == 0, MMREC_E_BADARG, "dropout_mask: bad arguments");
  MMREC_REQUIRE(aligned16(out), MMREC_E_ALIGN, "dropout_mask: out must be 16-byte aligned");
  DropSpec D;
  int rc = comb_drop(drop, D);
  if (rc != MMREC_OK) return rc;
  const size_t n4 = (size_t)planes * n * (d / 4);
  if (n4 == 0) return MMREC_OK;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, (size_t)kNumSMs * 8);
  dropout_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(reinterpret_cast<float4 *>(out), n4, D);
  MMREC_CHECK_LAUNCH("dropout_mask_kernel");
  return MMREC_OK;
}

static int combine_fwd_impl(const float *zv, const float *zt, const float *V, const float *T,
                                           const float *F, const float *C, const float *gi, const float *gt,
                                           const float *gf, const float *masks, const MmrecDropout *drop_host,
                                           int32_t n, int32_t d, float *side, float *all, void *stream_) {
  MMREC_REQUIRE(zv && zt && V && T && F && C && gi && gt && gf && side && all, MMREC_E_BADARG,
                "smore_combine_fwd: null pointer");
  MMREC_REQUIRE(n > 0 && comb_ok(d), MMREC_E_BADARG, "smore_combine_fwd: need n > 0 and d in {32, 64, 128}");
  MMREC_REQUIRE(aligned16(zv) && aligned16(zt) && aligned16(V) && aligned16(T) && aligned16(F) && aligned16(C) &&
                    aligned16(gi) && aligned16(gt) && aligned16(gf) && aligned16(masks) && aligned16(side) &&
                    aligned16(all), MMREC_E_ALIGN, "smore_combine_fwd: operands must be 16-byte aligned");
  DropSpec drop;
  int rc = comb_drop(drop_host, drop);
  if (rc != MMREC_OK) return rc;
  const CombIn in{zv, zt, V, T, F, C, gi, gt, gf, masks, drop};
  auto s = (cudaStream_t)stream_;
  const int blocks = comb_blocks(n, d);
  if (d == 32) smore_combine_fwd_kernel<8><<<blocks, kCombThreads, 0, s>>>(in, n, side, all);
  else if (d == 64) smore_combine_fwd_kernel<16><<<blocks, kCombThreads, 0, s>>>(in, n, side, all);
  else smore_combine_fwd_kernel<32><<<blocks, kCombThreads, 0, s>>>(in, n, side, all);
  MMREC_CHECK_LAUNCH("smore_combine_fwd_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_smore_combine_fwd_f32(const float *zv, const float *zt, const float *V, const float *T,
                                           const float *F, const float *C, const float *gi, const float *gt,
                                           const float *gf, const float *masks, int32_t n, int32_t d, float *side,
                                           float *all, void *stream_) {
  return combine_fwd_impl(zv, zt, V, T, F, C, gi, gt, gf, masks, nullptr, n, d, side, all, stream_);
}

extern "C" int mmrec_smore_combine_fwd_drop_f32(const float *zv, const float *zt, const float *V, const float *T,
                                                const float *F, const float *C, const float *gi, const float *gt,
                                                const float *gf, const MmrecDropout *drop, int32_t n, int32_t d,
                                                float *side, float *all, void *stream_) {
  return combine_fwd_impl(zv, zt, V, T, F, C, gi, gt, gf, nullptr, drop, n, d, side, all, stream_);
}

static int combine_bwd_impl(const float *g_all, const float *g_side, const float *zv, const float *zt,
                                           const float *V, const float *T, const float *F, const float *gi,
                                           const float *gt, const float *gf, const float *masks,
                                           const MmrecDropout *drop_host, int32_t n, int32_t d,
                                           float *dzv, float *dzt, float *dV, float *dT, float *dF, float *dC,
                                           float *dgi, float *dgt, float *dgf, void *stream_) {
  MMREC_REQUIRE(g_all || g_side, MMREC_E_BADARG, "smore_combine_bwd: both output gradients are null");
  MMREC_REQUIRE(zv && zt && V && T && F && gi && gt && gf && dzv && dzt && dV && dT && dF && dC && dgi && dgt && dgf,
                MMREC_E_BADARG, "smore_combine_bwd: null pointer");
  MMREC_REQUIRE(n > 0 && comb_ok(d), MMREC_E_BADARG, "smore_combine_bwd: need n > 0 and d in {32, 64, 128}");
  MMREC_REQUIRE(aligned16(g_all) && aligned16(g_side) && aligned16(zv) && aligned16(zt) && aligned16(V) && aligned16(T) &&
                    aligned16(F) && aligned16(gi) && aligned16(gt) && aligned16(gf) && aligned16(masks) &&
                    aligned16(dzv) && aligned16(dzt) && aligned16(dV) && aligned16(dT) && aligned16(dF) &&
                    aligned16(dC) && aligned16(dgi) && aligned16(dgt) && aligned16(dgf),
                MMREC_E_ALIGN, "smore_combine_bwd: operands must be 16-byte aligned");
  DropSpec drop;
  int rc = comb_drop(drop_host, drop);
  if (rc != MMREC_OK) return rc;
  const CombIn in{zv, zt, V, T, F, nullptr, gi, gt, gf, masks, drop};
  const CombGrad out{dzv, dzt, dV, dT, dF, dC, dgi, dgt, dgf};
  auto s = (cudaStream_t)stream_;
  const int blocks = comb_blocks(n, d);
  if (d == 32) smore_combine_bwd_kernel<8><<<blocks, kCombThreads, 0, s>>>(in, g_all, g_side, n, out);
  else if (d == 64) smore_combine_bwd_kernel<16><<<blocks, kCombThreads, 0, s>>>(in, g_all, g_side, n, out);
  else smore_combine_bwd_kernel<32><<<blocks, kCombThreads, 0, s>>>(in, g_all, g_side, n, out);
  MMREC_CHECK_LAUNCH("smore_combine_bwd_kernel");
  return MMREC_OK;
}

extern "C" int mmrec_smore_combine_bwd_f32(const float *g_all, const float *g_side, const float *zv, const float *zt,
                                           const float *V, const float *T, const float *F, const float *gi,
                                           const float *gt, const float *gf, const float *masks, int32_t n, int32_t d,
                                           float *dzv, float *dzt, float *dV, float *dT, float *dF, float *dC,
                                           float *dgi, float *dgt, float *dgf, void *stream_) {
  return combine_bwd_impl(g_all, g_side, zv, zt, V, T, F, gi, gt, gf, masks, nullptr, n, d, dzv, dzt, dV, dT, dF, dC,
                          dgi, dgt, dgf, stream_);
}

extern "C" int mmrec_smore_combine_bwd_drop_f32(const float *g_all, const float *g_side, const float *zv,
                                                const float *zt, const float *V, const float *T, const float *F,
                                                const float *gi, const float *gt, const float *gf,
                                                const MmrecDropout *drop, int32_t n, int32_t d, float *dzv, float *dzt,
                                                float *dV, float *dT, float *dF, float *dC, float *dgi, float *dgt,
                                                float *dgf, void *stream_) {
  return combine_bwd_impl(g_all, g_side, zv, zt, V, T, F, gi, gt, gf, nullptr, drop, n, d, dzv, dzt, dV, dT, dF, dC,
                          dgi, dgt, dgf, stream_);
}
