// SMORE's modality-aware preference module (smore.py:321-341) as ONE forward and ONE backward
// kernel (K14). Per node row (f = fusion_embeds, v = image_embeds,
// t = text_embeds, c = content_embeds, all [n, d]):
//   hv = tanh(Wq1v f + b);  sv = softmax_d(Wq2v hv);  av = sv * v          query_v, agg_image
//   ht = tanh(Wq1t f + b);  st = softmax_d(Wq2t ht);  at = st * t          query_t, agg_text
//   gi = sigmoid(Wgi c + b), gt = ..., gf = ...;  p* = mask* * g*           gate_*_prefer + nn.Dropout
//   side = (pi*av + pt*at + pf*f) / 3;   all = c + side                     stack/mean, smore.py:341
// The reference runs this as ~45 launches forward and ~90 backward (7 cuBLAS GEMMs, softmax, tanh,
// sigmoid, dropout, mul, stack, mean, add and their gradients), each a full pass over [n, d]
// tensors. Here a CTA owns a 64-row tile: the four inputs are read once, the seven 64x64 weight
// matrices stream through shared memory one after another, every intermediate lives in registers
// or shared memory, and only what the backward needs (hv, sv, ht, st, gi, gt, gf) is written.
// The backward re-reads those, chains all seven dX products and dW outer products on the same
// tiles and emits dF, dV, dT, dC plus per-CTA dW/db partials summed in a fixed order (no
// floating-point atomics). The 7 (fwd) + 14 (bwd) tile products of n*d*d MACs run on mma.sync
// tensor cores with the 3xTF32 split (fp32-class accuracy); the kernels are tensor/latency-bound
// (Baby: 0.76 / 1.5 GFMA x 3), not HBM-bound.
#include <stdlib.h>

#include "dense_tile.cuh"

namespace mmrec {
namespace {

using namespace dense;

constexpr int kNW = 7;        // q1v q2v q1t q2t gi gt gf
constexpr int kMaxSideParts = 1024;

struct SideWeights {
  const float *W[kNW];
  const float *b[kNW];        // NULL where the layer has no bias (q2v, q2t)
};
struct SideGrads {
  float *dW[kNW];
  float *db[kNW];
};

template <int D>
struct Cfg {
  static constexpr int TM = 4, CG = D / 4, RG = kT / CG, BM = RG * TM, P = D + 4;
  static constexpr int PW = D + 8;     // pitch of the staged weight matrix (conflict-free MMA B fragments)
  static constexpr int PW2 = D + 4;    // forward: pitch in float2 (hi, lo) pairs, PW2 % 16 == 4
  static constexpr int WPER = D * D / 4 / kT;
};

template <int W>
__device__ __forceinline__ float group_max(float v) {
  unsigned mask = 0xffffffffu;
  if constexpr (W < 32) {
    const unsigned lane = threadIdx.x & 31u;
    mask = ((1u << W) - 1u) << (W * (lane / W));
  }
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, W));
  return v;
}

// ---- fragments: thread (rg, cg) owns rows rg + RG*i (i < 4), columns 4*cg .. 4*cg+3 ----------
template <int D>
struct Frag {
  using C = Cfg<D>;
  float v[4][4];
  __device__ __forceinline__ void fill(float x) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = x;
  }
  __device__ __forceinline__ void load(const float *__restrict__ p, int m0, int n, int rg, int c0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + rg + C::RG * i;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < n) t = ldg4(p + (size_t)m * D + c0);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
  }
  __device__ __forceinline__ void store(float *__restrict__ p, int m0, int n, int rg, int c0) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + rg + C::RG * i;
      if (m < n) *reinterpret_cast<float4 *>(p + (size_t)m * D + c0) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
    }
  }
  // dropout multipliers of plane `g` ([3, n, D] layout) for this thread's rows / columns
  __device__ __forceinline__ void fill_dropout(uint64_t stream, const DropSpec &ds, int g, int m0, int n, int rg, int c0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + rg + C::RG * i;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < n) t = drop_mask4(stream, drop_row4(ds, g, m, n, D / 4) + c0 / 4, ds.p);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
  }
  __device__ __forceinline__ void load_smem(const float *s, int rg, int c0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = *reinterpret_cast<const float4 *>(s + (rg + C::RG * i) * C::P + c0);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    }
  }
  __device__ __forceinline__ void store_smem(float *s, int rg, int c0) const {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4 *>(s + (rg + C::RG * i) * C::P + c0) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  }
};

// A weight matrix on its way to shared memory: transposed (forward: Ws[k][n] = W[n][k]) or natural.
template <int D>
struct WStage {
  float4 w[Cfg<D>::WPER];
  __device__ __forceinline__ void load_t(const float *__restrict__ W) {
#pragma unroll
    for (int i = 0; i < Cfg<D>::WPER; ++i) {
      const int idx = threadIdx.x + i * kT, n = idx % D, k4 = idx / D;
      w[i] = ldg4(W + (size_t)n * D + k4 * 4);
    }
  }
  __device__ __forceinline__ void store_t(float *Ws) const {
#pragma unroll
    for (int i = 0; i < Cfg<D>::WPER; ++i) {
      const int idx = threadIdx.x + i * kT, n = idx % D, k4 = idx / D;
      Ws[(k4 * 4 + 0) * Cfg<D>::PW + n] = w[i].x;
      Ws[(k4 * 4 + 1) * Cfg<D>::PW + n] = w[i].y;
      Ws[(k4 * 4 + 2) * Cfg<D>::PW + n] = w[i].z;
      Ws[(k4 * 4 + 3) * Cfg<D>::PW + n] = w[i].w;
    }
  }
  // transposed, split and paired for tile_mma_tc_b2: element k of output column n goes to half
  // (k % 8) / 4 of the float2 at [(k / 8) * 4 + k % 4][n] of the hi plane (Ws2) and of the lo plane
  // (Ws2 + (D / 2) * PW2)
  __device__ __forceinline__ void store_t2(float2 *Ws2) const {
    float *hi_plane = reinterpret_cast<float *>(Ws2);
    float *lo_plane = reinterpret_cast<float *>(Ws2 + (D / 2) * Cfg<D>::PW2);
#pragma unroll
    for (int i = 0; i < Cfg<D>::WPER; ++i) {
      const int idx = threadIdx.x + i * kT, n = idx % D, k4 = idx / D;
      const float e[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t hi, lo;
        split_tf32_u(e[q], hi, lo);
        const int o = 2 * (((k4 >> 1) * 4 + q) * Cfg<D>::PW2 + n) + (k4 & 1);
        hi_plane[o] = __uint_as_float(hi);
        lo_plane[o] = __uint_as_float(lo);
      }
    }
  }
  __device__ __forceinline__ void load_n(const float *__restrict__ W) {
#pragma unroll
    for (int i = 0; i < Cfg<D>::WPER; ++i) w[i] = ldg4(W + (size_t)(threadIdx.x + i * kT) * 4);
  }
  __device__ __forceinline__ void store_n(float *Wn) const {
#pragma unroll
    for (int i = 0; i < Cfg<D>::WPER; ++i) {
      const int idx = threadIdx.x + i * kT, n = idx / (D / 4), k4 = idx % (D / 4);
      *reinterpret_cast<float4 *>(Wn + n * Cfg<D>::PW + k4 * 4) = w[i];
    }
  }
};

template <int D>
__device__ __forceinline__ void bias_add(float (&acc)[4][4], const float *__restrict__ b, int c0) {
  if (b == nullptr) return;
  const float4 bv = ldg4(b + c0);
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] += bv.x; acc[i][1] += bv.y; acc[i][2] += bv.z; acc[i][3] += bv.w; }
}

// acc = A tile (shared, pitch P) x staged weight (pitch PW) on the tensor cores; the product passes
// through the shared tile sO so that every thread gets the rows/columns its fragment owns.
// Leaves the CTA synchronised after the product is visible.
template <int D>
__device__ __forceinline__ void product(Frag<D> &acc, float *sO, const float *sA, const float2 *sW2, int rg, int c0) {
  using C = Cfg<D>;
  tile_mma_tc_b2<D, D, C::BM, C::P, C::PW2>(sO, C::P, sA, sW2, sW2 + (D / 2) * C::PW2);
  __syncthreads();
  acc.load_smem(sO, rg, c0);
}

// softmax over the d columns of every row; a row is spread over the CG threads of one sub-warp
template <int D>
__device__ __forceinline__ void softmax_rows(float (&a)[4][4]) {
  constexpr int CG = Cfg<D>::CG;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float mx = fmaxf(fmaxf(a[i][0], a[i][1]), fmaxf(a[i][2], a[i][3]));
    mx = group_max<CG>(mx);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[i][j] = fast_exp(a[i][j] - mx); s += a[i][j]; }
    s = fast_rcp(group_sum<CG>(s));
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = a[i][j] * s;
  }
}

// ================================================================================== forward
template <int D>
__global__ void __launch_bounds__(kT, D <= 64 ? 2 : 1)
side_fwd_kernel(const float *__restrict__ F, const float *__restrict__ V, const float *__restrict__ T,
                const float *__restrict__ C_, SideWeights P, const float *__restrict__ masks, const DropSpec drop,
                float *__restrict__ saved, float *__restrict__ side, float *__restrict__ all, int n, int n_tiles) {
  const uint64_t drop_key = drop.p > 0.f ? drop_stream(drop) : 0ull;
  using C = Cfg<D>;
  extern __shared__ float4 smem4[];
  float2 *Ws = reinterpret_cast<float2 *>(smem4); // transposed weight, hi plane [D/2][PW2] + lo plane, k-paired float2
  float *sF = reinterpret_cast<float *>(Ws + D * C::PW2);   // [BM][P]
  float *sC = sF + C::BM * C::P;
  float *sH = sC + C::BM * C::P;
  float *sO = sH + C::BM * C::P;                  // product staging
  const int cg = threadIdx.x % C::CG, rg = threadIdx.x / C::CG, c0 = cg * 4;
  const size_t nd = (size_t)n * D;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * C::BM;
    RowStage<D, C::BM> fs, cs;
    fs.load(F, m0, n);
    cs.load(C_, m0, n);
    WStage<D> w;
    w.load_t(P.W[0]);
    if (tile != (int)blockIdx.x) __syncthreads();       // previous tile done with shared memory
    fs.store(sF);
    cs.store(sC);
    Frag<D> acc, agg_v, agg_t, x;
    // ---- query_v / query_t: tanh(W1 f + b) -> W2 h -> softmax -> * v | t ----
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      w.store_t2(Ws);
      w.load_t(P.W[2 * br + 1]);
      __syncthreads();
      product<D>(acc, sO, sF, Ws, rg, c0);
      bias_add<D>(acc.v, P.b[2 * br], c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.v[i][j] = fast_tanh(acc.v[i][j]);
      acc.store_smem(sH, rg, c0);
      if (saved != nullptr) acc.store(saved + (2 * br) * nd, m0, n, rg, c0);
      __syncthreads();                                  // sH complete, Ws free
      w.store_t2(Ws);
      w.load_t(P.W[br == 0 ? 2 : 4]);
      __syncthreads();
      product<D>(acc, sO, sH, Ws, rg, c0);
      softmax_rows<D>(acc.v);
      if (saved != nullptr) acc.store(saved + (2 * br + 1) * nd, m0, n, rg, c0);
      x.load(br == 0 ? V : T, m0, n, rg, c0);
      Frag<D> &agg = br == 0 ? agg_v : agg_t;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) agg.v[i][j] = acc.v[i][j] * x.v[i][j];
      __syncthreads();                                  // Ws and sH free
    }
    // ---- preference gates on the content tile; side = mean of the three gated views ----
    Frag<D> sd;
    sd.fill(0.f);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      w.store_t2(Ws);
      if (g < 2) w.load_t(P.W[5 + g]);
      __syncthreads();
      product<D>(acc, sO, sC, Ws, rg, c0);
      bias_add<D>(acc.v, P.b[4 + g], c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.v[i][j] = fast_sigmoid(acc.v[i][j]);
      if (saved != nullptr) acc.store(saved + (4 + g) * nd, m0, n, rg, c0);
      if (masks != nullptr || drop.p > 0.f) {
        if (masks != nullptr) x.load(masks + g * nd, m0, n, rg, c0);
        else x.fill_dropout(drop_key, drop, g, m0, n, rg, c0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc.v[i][j] *= x.v[i][j];
      }
      if (g == 2) x.load_smem(sF, rg, c0);
      const Frag<D> &agg = g == 0 ? agg_v : g == 1 ? agg_t : x;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sd.v[i][j] = fmaf(acc.v[i][j], agg.v[i][j], sd.v[i][j]);
      if (g < 2) __syncthreads();                       // Ws free for the next gate
    }
    x.load_smem(sC, rg, c0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sd.v[i][j] *= (1.f / 3.f);
        x.v[i][j] += sd.v[i][j];
      }
    sd.store(side, m0, n, rg, c0);
    x.store(all, m0, n, rg, c0);
  }
}

// ================================================================================== backward
// partial layout: [kNW][D + 1 rows (dW rows, then db)][n_parts][D]; a CTA that owns several tiles accumulates into its slot.
template <int D>
__global__ void __launch_bounds__(kT, D <= 64 ? 2 : 1)
side_bwd_kernel(const float *__restrict__ d_all, const float *__restrict__ d_side, const float *__restrict__ F,
                const float *__restrict__ V, const float *__restrict__ T, const float *__restrict__ C_,
                SideWeights P, const float *__restrict__ masks, const DropSpec drop, const float *__restrict__ saved,
                float *__restrict__ dF, float *__restrict__ dV, float *__restrict__ dT, float *__restrict__ dC,
                float *__restrict__ partial, int n, int n_tiles) {
  const uint64_t drop_key = drop.p > 0.f ? drop_stream(drop) : 0ull;
  using C = Cfg<D>;
  extern __shared__ float4 smem4[];
  float *Wn = reinterpret_cast<float *>(smem4);   // [D][PW] natural layout (output-major)
  float *sZ = Wn + D * C::PW;                     // [BM][P] dz tile
  float *sC = sZ + C::BM * C::P;
  float *sF = sC + C::BM * C::P;
  float *sH = sF + C::BM * C::P;
  float *sO = sH + C::BM * C::P;                  // product staging
  const int cg = threadIdx.x % C::CG, rg = threadIdx.x / C::CG, c0 = cg * 4;
  const size_t nd = (size_t)n * D;
  const int n_parts = gridDim.x;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * C::BM;
    const bool first = tile == (int)blockIdx.x;
    RowStage<D, C::BM> fs, cs;
    fs.load(F, m0, n);
    cs.load(C_, m0, n);
    Frag<D> ds, dc, df, acc;
    ds.fill(0.f);
    if (d_all != nullptr) ds.load(d_all, m0, n, rg, c0);
    dc = ds;                                            // all = c + side
    if (d_side != nullptr) {
      acc.load(d_side, m0, n, rg, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ds.v[i][j] += acc.v[i][j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) ds.v[i][j] *= (1.f / 3.f);
    df.fill(0.f);
    if (!first) __syncthreads();
    fs.store(sF);
    cs.store(sC);
    WStage<D> w;
    // One dense layer of the backward chain: dz is in `acc`; X is the staged input tile of the
    // layer. out += dz W;  partial[w_idx] (+)= dz^T X, sum_rows dz. Both products on tensor cores.
    auto layer_bwd = [&](int w_idx, const float *sX, Frag<D> &out, bool with_bias) {
      w.load_n(P.W[w_idx]);
      __syncthreads();                                  // earlier readers of sZ / Wn / sO are done
      acc.store_smem(sZ, rg, c0);
      w.store_n(Wn);
      __syncthreads();
      // row r of this CTA's dW partial (r = D: the bias row) sits at [w][r][part][D]: the reduce
      // kernel then streams n_parts contiguous rows per output row instead of 256 bytes out of
      // every 116 KB slab
      float *slot = partial + (size_t)w_idx * (D + 1) * n_parts * D + (size_t)blockIdx.x * D;
      const int pitch = n_parts * D;
      tile_mma_tc<D, D, C::BM, C::P, C::PW, false>(sO, C::P, C::BM, sZ, Wn);          // dz W
      tile_mma_tc<C::BM, D, D, C::P, C::P, true>(slot, pitch, D, sZ, sX, !first);     // dz^T X
      if (with_bias && threadIdx.x < D) {
        float sdb = 0.f;
#pragma unroll 8
        for (int m = 0; m < C::BM; ++m) sdb += sZ[m * C::P + threadIdx.x];
        float *bslot = slot + (size_t)D * pitch + threadIdx.x;
        *bslot = first ? sdb : *bslot + sdb;
      }
      __syncthreads();
      Frag<D> tmp;
      tmp.load_smem(sO, rg, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out.v[i][j] += tmp.v[i][j];
    };
#pragma unroll 1
    for (int br = 0; br < 3; ++br) {
      // ---- preference gate br: p = mask * g, view a = (sv*v | st*t | f) ----
      Frag<D> g, s, x, da;
      g.load(saved + (4 + br) * nd, m0, n, rg, c0);
      if (br < 2) {
        s.load(saved + (2 * br + 1) * nd, m0, n, rg, c0);
        x.load(br == 0 ? V : T, m0, n, rg, c0);
      } else {
        x.load_smem(sF, rg, c0);
      }
      Frag<D> mk;
      mk.fill(1.f);
      if (masks != nullptr) mk.load(masks + br * nd, m0, n, rg, c0);
      else if (drop.p > 0.f) mk.fill_dropout(drop_key, drop, br, m0, n, rg, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gg = g.v[i][j], a = br < 2 ? s.v[i][j] * x.v[i][j] : x.v[i][j];
          da.v[i][j] = ds.v[i][j] * (mk.v[i][j] * gg);                        // d(view)
          acc.v[i][j] = ds.v[i][j] * a * mk.v[i][j] * ((1.f - gg) * gg);      // dz of the gate
        }
      layer_bwd(4 + br, sC, dc, true);
      if (br == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) df.v[i][j] += da.v[i][j];
        break;
      }
      // ---- query branch: a = s * x, s = softmax(W2 h), h = tanh(W1 f + b) ----
      RowStage<D, C::BM> hs;
      hs.load(saved + (2 * br) * nd, m0, n);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float rs = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d_s = da.v[i][j] * x.v[i][j];
          x.v[i][j] = da.v[i][j] * s.v[i][j];            // d(v | t)
          da.v[i][j] = d_s;
          rs = fmaf(d_s, s.v[i][j], rs);
        }
        rs = group_sum<C::CG>(rs);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.v[i][j] = s.v[i][j] * (da.v[i][j] - rs);   // dq
      }
      x.store(br == 0 ? dV : dT, m0, n, rg, c0);
      __syncthreads();                                  // sH free (previous branch's readers done)
      hs.store(sH);
      Frag<D> dh;
      dh.fill(0.f);
      layer_bwd(2 * br + 1, sH, dh, false);             // syncs inside make sH visible
      g.load_smem(sH, rg, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.v[i][j] = dh.v[i][j] * (1.f - g.v[i][j] * g.v[i][j]);
      layer_bwd(2 * br, sF, df, true);
    }
    df.store(dF, m0, n, rg, c0);
    dc.store(dC, m0, n, rg, c0);
  }
}

// out[w][row][col] = sum over parts of partial[w][row][part][col] (row D = the bias). A CTA owns 64
// consecutive outputs (one row at D = 64) and ALL their partials: thread (tx, ty) takes the float4
// column tx of parts ty, ty + TY, ... -- every load of a thread is in flight at once (the kernel used to
// walk its parts in dependent batches: four memory round trips per CTA times 1.5 waves = 26 us for a
// 48 MB stream), a warp reads 512 contiguous bytes per request, the sums run in a fixed order.
// TY = 64 (1024 threads, one CTA per SM) for the ~400 partial slabs of an all-rows backward; TY = 16 (256
// threads, eight CTAs per SM: the 455 CTAs are one wave instead of three) for the <= 128 slabs of a
// batch-row backward, where the 1024-thread CTAs spent 19 us on an 11 MB stream (ncu,
// profiles/r02_batch_rows_ncu_summary.txt).
template <int TY>
__global__ void __launch_bounds__(16 * TY, TY == 64 ? 1 : 4)
side_partial_reduce_kernel(const float *__restrict__ partial, int n_parts, int n_w, int n_b, SideGrads G) {
  static_assert(TY % 8 == 0 && TY >= 8, "two-stage tree over groups of eight");
  __shared__ float4 sm[TY][16];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j = (blockIdx.x * 16 + tx) * 4, wi = blockIdx.y;
  const int D = n_b, row = j / D, col = j % D;
  const bool live = j < n_w + n_b;
  const float *base = partial + ((size_t)wi * (D + 1) + row) * n_parts * D + col;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p0 = ty; p0 < n_parts; p0 += TY * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = live && p0 + TY * u < n_parts ? __ldcs(reinterpret_cast<const float4 *>(base + (size_t)(p0 + TY * u) * D))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty < 8) {                                   // parts ty, ty + 8, ... of the TY partial sums
    float4 a = sm[ty][tx];
#pragma unroll
    for (int g = 1; g < TY / 8; ++g) {
      const float4 b = sm[ty + 8 * g][tx];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    sm[ty][tx] = a;
  }
  __syncthreads();
  if (ty == 0 && live) {
    float4 a = sm[0][tx];
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      const float4 b = sm[g][tx];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (j < n_w) *reinterpret_cast<float4 *>(G.dW[wi] + j) = a;
    else if (G.db[wi] != nullptr) *reinterpret_cast<float4 *>(G.db[wi] + (j - n_w)) = a;
  }
}

template <int D>
constexpr size_t side_fwd_smem() { return sizeof(float) * (2 * D * Cfg<D>::PW2 + 4 * Cfg<D>::BM * Cfg<D>::P); }
template <int D>
constexpr size_t side_bwd_smem() { return sizeof(float) * (D * Cfg<D>::PW + 5 * Cfg<D>::BM * Cfg<D>::P); }

inline int side_parts(int n, int d) {
  const int bm = (kT / (d / 4)) * 4;
  // backward CTAs = dW / db partial slabs: two resident CTAs per SM loop over the tiles (MMREC_SIDE_PARTS overrides)
  static const int cap = getenv("MMREC_SIDE_PARTS") ? atoi(getenv("MMREC_SIDE_PARTS")) : kMaxSideParts;
  return max(1, min(min((n + bm - 1) / bm, kMaxSideParts), cap));
}

template <int D>
int side_fwd_launch(const float *F, const float *V, const float *T, const float *C_, const SideWeights &P,
                    const float *masks, const DropSpec &drop, float *saved, float *side, float *all, int n,
                    cudaStream_t st) {
  constexpr size_t smem = side_fwd_smem<D>();
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(side_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int n_tiles = (n + Cfg<D>::BM - 1) / Cfg<D>::BM;
  side_fwd_kernel<D><<<min(n_tiles, 8 * kNumSMs), kT, smem, st>>>(F, V, T, C_, P, masks, drop, saved, side, all, n, n_tiles);
  MMREC_CHECK_LAUNCH("side_fwd_kernel");
  return MMREC_OK;
}

template <int D>
int side_bwd_launch(const float *d_all, const float *d_side, const float *F, const float *V, const float *T,
                    const float *C_, const SideWeights &P, const float *masks, const DropSpec &drop, const float *saved,
                    float *dF, float *dV, float *dT, float *dC, const SideGrads &G, float *ws, int n, cudaStream_t st) {
  constexpr size_t smem = side_bwd_smem<D>();
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(side_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int n_tiles = (n + Cfg<D>::BM - 1) / Cfg<D>::BM;
  const int parts = side_parts(n, D);
  side_bwd_kernel<D><<<parts, kT, smem, st>>>(d_all, d_side, F, V, T, C_, P, masks, drop, saved, dF, dV, dT, dC, ws, n,
                                              n_tiles);
  MMREC_CHECK_LAUNCH("side_bwd_kernel");
  const dim3 rgrid((D * D + D + 63) / 64, kNW);
  static const bool wide_only = getenv("MMREC_SIDE_REDUCE_WIDE") && atoi(getenv("MMREC_SIDE_REDUCE_WIDE")) != 0;   // A/B
  if (parts <= 128 && !wide_only) side_partial_reduce_kernel<16><<<rgrid, 256, 0, st>>>(ws, parts, D * D, D, G);
  else side_partial_reduce_kernel<64><<<rgrid, 1024, 0, st>>>(ws, parts, D * D, D, G);
  MMREC_CHECK_LAUNCH("side_partial_reduce_kernel");
  return MMREC_OK;
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_smore_side_supported(int32_t d) { return d == 32 || d == 64 || d == 128; }

extern "C" size_t mmrec_smore_side_bwd_workspace_bytes(int32_t n, int32_t d) {
  if (!mmrec_smore_side_supported(d)) return 0;
  return sizeof(float) * (size_t)kNW * side_parts(n, d) * ((size_t)d * d + d);
}

static int unpack_weights(const float *const *W, const float *const *b, int d, SideWeights &P) {
  for (int i = 0; i < kNW; ++i) {
    MMREC_REQUIRE(W[i] != nullptr && aligned16(W[i]) && aligned16(b[i]), MMREC_E_BADARG,
                  "smore_side: weight %d is null or misaligned", i);
    P.W[i] = W[i];
    P.b[i] = b[i];
  }
  (void)d;
  return MMREC_OK;
}

static int unpack_drop(const MmrecDropout *drop, DropSpec &D) {
  D = DropSpec{nullptr, 0ull, 0.f, nullptr, 0};
  if (drop == nullptr) return MMREC_OK;
  MMREC_REQUIRE(drop->p >= 0.f && drop->p < 1.f, MMREC_E_BADARG, "dropout: p must be in [0, 1) (got %g)", (double)drop->p);
  MMREC_REQUIRE(drop->row_ids == nullptr || drop->n_total > 0, MMREC_E_BADARG, "dropout: row_ids needs n_total");
  D = DropSpec{drop->counter, drop->seed, drop->p, reinterpret_cast<const long long *>(drop->row_ids), drop->n_total};
  return MMREC_OK;
}

static int side_fwd_impl(const float *F, const float *V, const float *T, const float *C_, const float *const *W_host,
                         const float *const *b_host, const float *masks, const MmrecDropout *drop_host, float *saved,
                         float *side, float *all, int32_t n, int32_t d, void *stream) {
  MMREC_REQUIRE(F && V && T && C_ && W_host && b_host && side && all, MMREC_E_BADARG,
                "smore_side_fwd: null pointer");      // saved may be NULL: inference, nothing kept for a backward
  MMREC_REQUIRE(mmrec_smore_side_supported(d), MMREC_E_BADARG, "smore_side_fwd: d must be 32, 64 or 128 (got %d)", d);
  MMREC_REQUIRE(n >= 0, MMREC_E_BADARG, "smore_side_fwd: bad n");
  MMREC_REQUIRE(aligned16(F) && aligned16(V) && aligned16(T) && aligned16(C_) && aligned16(masks) &&
                    aligned16(saved) && aligned16(side) && aligned16(all), MMREC_E_ALIGN,
                "smore_side_fwd: operands must be 16-byte aligned");
  SideWeights P;
  int rc = unpack_weights(W_host, b_host, d, P);
  if (rc != MMREC_OK) return rc;
  DropSpec drop;
  rc = unpack_drop(drop_host, drop);
  if (rc != MMREC_OK) return rc;
  if (n == 0) return MMREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return d == 64    ? side_fwd_launch<64>(F, V, T, C_, P, masks, drop, saved, side, all, n, st)
         : d == 128 ? side_fwd_launch<128>(F, V, T, C_, P, masks, drop, saved, side, all, n, st)
                    : side_fwd_launch<32>(F, V, T, C_, P, masks, drop, saved, side, all, n, st);
}

static int side_bwd_impl(const float *d_all, const float *d_side, const float *F, const float *V, const float *T,
                         const float *C_, const float *const *W_host, const float *const *b_host, const float *masks,
                         const MmrecDropout *drop_host, const float *saved, float *dF, float *dV, float *dT, float *dC,
                         float *const *dW_host, float *const *db_host, float *ws, int32_t n, int32_t d, void *stream) {
  MMREC_REQUIRE(F && V && T && C_ && W_host && b_host && saved && dF && dV && dT && dC && dW_host && db_host && ws,
                MMREC_E_BADARG, "smore_side_bwd: null pointer");
  MMREC_REQUIRE(d_all || d_side, MMREC_E_BADARG, "smore_side_bwd: no incoming gradient");
  MMREC_REQUIRE(mmrec_smore_side_supported(d), MMREC_E_BADARG, "smore_side_bwd: d must be 32, 64 or 128 (got %d)", d);
  MMREC_REQUIRE(n > 0, MMREC_E_BADARG, "smore_side_bwd: bad n");
  MMREC_REQUIRE(aligned16(d_all) && aligned16(d_side) && aligned16(F) && aligned16(V) && aligned16(T) &&
                    aligned16(C_) && aligned16(masks) && aligned16(saved) && aligned16(dF) && aligned16(dV) &&
                    aligned16(dT) && aligned16(dC), MMREC_E_ALIGN, "smore_side_bwd: operands must be 16-byte aligned");
  SideWeights P;
  int rc = unpack_weights(W_host, b_host, d, P);
  if (rc != MMREC_OK) return rc;
  DropSpec drop;
  rc = unpack_drop(drop_host, drop);
  if (rc != MMREC_OK) return rc;
  SideGrads G;
  for (int i = 0; i < kNW; ++i) {
    MMREC_REQUIRE(dW_host[i] != nullptr, MMREC_E_BADARG, "smore_side_bwd: dW[%d] is null", i);
    G.dW[i] = dW_host[i];
    G.db[i] = db_host[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  return d == 64    ? side_bwd_launch<64>(d_all, d_side, F, V, T, C_, P, masks, drop, saved, dF, dV, dT, dC, G, ws, n, st)
         : d == 128 ? side_bwd_launch<128>(d_all, d_side, F, V, T, C_, P, masks, drop, saved, dF, dV, dT, dC, G, ws, n, st)
                    : side_bwd_launch<32>(d_all, d_side, F, V, T, C_, P, masks, drop, saved, dF, dV, dT, dC, G, ws, n, st);
}

extern "C" int mmrec_smore_side_fwd_f32(const float *F, const float *V, const float *T, const float *C_,
                                        const float *const *W_host, const float *const *b_host,
                                        const float *masks, float *saved, float *side, float *all, int32_t n,
                                        int32_t d, void *stream) {
  return side_fwd_impl(F, V, T, C_, W_host, b_host, masks, nullptr, saved, side, all, n, d, stream);
}

extern "C" int mmrec_smore_side_fwd_drop_f32(const float *F, const float *V, const float *T, const float *C_,
                                             const float *const *W_host, const float *const *b_host,
                                             const MmrecDropout *drop, float *saved, float *side, float *all,
                                             int32_t n, int32_t d, void *stream) {
  return side_fwd_impl(F, V, T, C_, W_host, b_host, nullptr, drop, saved, side, all, n, d, stream);
}

extern "C" int mmrec_smore_side_bwd_f32(const float *d_all, const float *d_side, const float *F, const float *V,
                                        const float *T, const float *C_, const float *const *W_host,
                                        const float *const *b_host, const float *masks, const float *saved,
                                        float *dF, float *dV, float *dT, float *dC, float *const *dW_host,
                                        float *const *db_host, float *ws, int32_t n, int32_t d, void *stream) {
  return side_bwd_impl(d_all, d_side, F, V, T, C_, W_host, b_host, masks, nullptr, saved, dF, dV, dT, dC, dW_host,
                       db_host, ws, n, d, stream);
}

extern "C" int mmrec_smore_side_bwd_drop_f32(const float *d_all, const float *d_side, const float *F, const float *V,
                                             const float *T, const float *C_, const float *const *W_host,
                                             const float *const *b_host, const MmrecDropout *drop, const float *saved,
                                             float *dF, float *dV, float *dT, float *dC, float *const *dW_host,
                                             float *const *db_host, float *ws, int32_t n, int32_t d, void *stream) {
  return side_bwd_impl(d_all, d_side, F, V, T, C_, W_host, b_host, nullptr, drop, saved, dF, dV, dT, dC, dW_host,
                       db_host, ws, n, d, stream);
}
