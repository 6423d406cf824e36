// Shared pieces of the fp32 FMA tile kernels for the d x d dense layers (dense_small.cu) and the
// fused SMORE side network (side_net.cu): activation helpers, the two-phase global -> shared row
// stager, the register-tile product and the dW outer-product accumulation.
#pragma once
#include "common.cuh"

namespace mmrec {
namespace dense {

constexpr int kT = 256;

enum Act { kNone = 0, kTanh = 1, kSigmoid = 2 };

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if constexpr (ACT == kTanh) return fast_tanh(z);
  if constexpr (ACT == kSigmoid) return fast_sigmoid(z);
  return z;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float dy, float y) {
  if constexpr (ACT == kTanh) return dy * (1.f - y * y);
  if constexpr (ACT == kSigmoid) return dy * ((1.f - y) * y);
  return dy;
}

template <int K, int N, int TM>
struct Tile {
  static constexpr int CG = N / 4;          // column groups (4 output columns each)
  static constexpr int RG = kT / CG;        // row groups
  static constexpr int BM = RG * TM;        // rows per tile
  static constexpr int XP = K + 4;          // pitch of a staged X row (floats)
};

// Rows [m0, m0+BM) of a row-major [M, C] array on their way to shared memory (pitch C+4; rows
// >= M are zero), in two phases: load() puts every global load of the thread in flight, store()
// parks the values. Whatever sits between the two (the W staging of the first tile) overlaps the
// DRAM round trip instead of adding one.
template <int C, int BM>
struct RowStage {
  static constexpr int V = C / 4, PER = BM * V / kT;
  static_assert(BM * V % kT == 0, "tile must be a multiple of the CTA");
  float4 v[PER];
  __device__ __forceinline__ void load(const float *__restrict__ src, int m0, int M) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = threadIdx.x + i * kT, r = idx / V, c4 = idx % V;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v[i] = ldg4(src + (size_t)(m0 + r) * C + c4 * 4);
    }
  }
  __device__ __forceinline__ void store(float *dst) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = threadIdx.x + i * kT, r = idx / V, c4 = idx % V;
      *reinterpret_cast<float4 *>(dst + r * (C + 4) + c4 * 4) = v[i];
    }
  }
};

// acc[i][0..3] += sum_k A[row_i][k] * B[k][c0..c0+3]; A staged with pitch KK+4, B k-major pitch NB
template <int KK, int TM, int RG, int NB>
__device__ __forceinline__ void tile_mma(float (&acc)[TM][4], const float *__restrict__ As, int rg,
                                         const float *__restrict__ Bs, int c0) {
#pragma unroll 4
  for (int k4 = 0; k4 < KK / 4; ++k4) {
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i)
      a[i] = *reinterpret_cast<const float4 *>(As + (rg + RG * i) * (KK + 4) + k4 * 4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 b = *reinterpret_cast<const float4 *>(Bs + (k4 * 4 + kk) * NB + c0);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, b.x, acc[i][0]);
        acc[i][1] = fmaf(av, b.y, acc[i][1]);
        acc[i][2] = fmaf(av, b.z, acc[i][2]);
        acc[i][3] = fmaf(av, b.w, acc[i][3]);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// Tensor-core versions of the two tile products (mma.sync m16n8k8 tf32, 3xTF32 split in
// registers: lo*hi + hi*lo + hi*hi, fp32 accumulate -> fp32-class accuracy for K <= 128). The d x d
// layers are 0.2-1.5 GFMA per launch and FMA-bound on the CUDA cores; the warp-level MMA does
// the same product with ~2.5x fewer issue slots. Results go back through shared memory (or
// straight to global for dW), so the surrounding fragment code is unchanged.
__device__ __forceinline__ void split_tf32_u(float x, uint32_t &hi, uint32_t &lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;        // round-to-nearest tf32 (no inf/nan guard)
  lo = __float_as_uint(x - __uint_as_float(hi));            // the tensor core ignores the low 13 bits
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Warp tiling of a [BM x NN] product over the 8 warps of the CTA: BM/16 row slabs, the warps of a
// slab split the columns.
template <int BM, int NN>
struct TcTile {
  static constexpr int SLABS = BM / 16, WPS = (kT / 32) / SLABS, WN = NN / WPS, NF = WN / 8;
  static_assert(BM % 16 == 0 && (kT / 32) % SLABS == 0 && NN % (WPS * 8) == 0, "tile does not fit 8 warps");
};

struct EpiIdentity {
  __device__ __forceinline__ float2 operator()(float2 v, int /*n*/) const { return v; }
};

// out[m][n] (+)= epi(sum_k A[m][k] * B[k][n]): A at As[m * PA + k] (TRANS_A: As[k * PA + m]), B at
// Bs[k * PB + n] (TRANS_B: Bs[n * PB + k]); out at out[m * PO + n] (shared or global; rows >= m_lim
// are skipped). `epi` maps the two adjacent columns (n, n + 1) of a result row, e.g. bias + activation.
template <int KK, int NN, int BM, int PA, int PB, bool TRANS_A, typename Epi = EpiIdentity, bool TRANS_B = false>
__device__ __forceinline__ void tile_mma_tc(float *__restrict__ out, int PO, int m_lim, const float *__restrict__ As,
                                            const float *__restrict__ Bs, bool accumulate = false, Epi epi = Epi()) {
  using T = TcTile<BM, NN>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rs = (warp % T::SLABS) * 16, cs = (warp / T::SLABS) * T::WN;
  float c[T::NF][4];
#pragma unroll
  for (int j = 0; j < T::NF; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll 2
  for (int k8 = 0; k8 < KK; k8 += 8) {
    uint32_t ah[4], al[4];
    auto A = [&](int m, int k) { return TRANS_A ? As[k * PA + m] : As[m * PA + k]; };
    // Both operands k-major in shared memory (the dW = dz^T X products): with the MMA's own k slots
    // (t, t + 4) a warp's 32 loads fall on banks 4 t + g of a pitch = 4 (mod 32) tile -- two-way
    // conflicts on every A and B fragment load (29 % of side_bwd's shared wavefronts, ncu). The sum
    // over k does not care which staged row feeds which slot as long as A and B agree: slot t reads row
    // 2 t, slot t + 4 row 2 t + 1 -> banks 8 t + g, conflict-free.
    constexpr bool KPERM = TRANS_A && !TRANS_B;
    const int ka = KPERM ? k8 + 2 * t : k8 + t, kb = KPERM ? k8 + 2 * t + 1 : k8 + t + 4;
    split_tf32_u(A(rs + g, ka), ah[0], al[0]);
    split_tf32_u(A(rs + g + 8, ka), ah[1], al[1]);
    split_tf32_u(A(rs + g, kb), ah[2], al[2]);
    split_tf32_u(A(rs + g + 8, kb), ah[3], al[3]);
#pragma unroll
    for (int j = 0; j < T::NF; ++j) {
      uint32_t bh[2], bl[2];
      const int n = cs + j * 8 + g;
      split_tf32_u(TRANS_B ? Bs[n * PB + ka] : Bs[ka * PB + n], bh[0], bl[0]);
      split_tf32_u(TRANS_B ? Bs[n * PB + kb] : Bs[kb * PB + n], bh[1], bl[1]);
      mma_tf32_16x8x8(c[j], al, bh);
      mma_tf32_16x8x8(c[j], ah, bl);
      mma_tf32_16x8x8(c[j], ah, bh);
    }
  }
#pragma unroll
  for (int j = 0; j < T::NF; ++j) {
    const int n = cs + j * 8 + 2 * t;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = rs + g + 8 * h;
      if (m < m_lim) {
        float2 *dst = reinterpret_cast<float2 *>(out + (size_t)m * PO + n);
        float2 v = epi(make_float2(c[j][2 * h], c[j][2 * h + 1]), n);
        if (accumulate) {
          const float2 o = *dst;
          v.x += o.x; v.y += o.y;
        }
        *dst = v;
      }
    }
  }
}

// The same product with a PRE-SPLIT B operand in MMA fragment order: for the k-step kb (8 rows of
// B) and t < 4, Bhi[(kb * 4 + t) * PB2 + n] = (hi of B[8 kb + t][n], hi of B[8 kb + t + 4][n]) as one
// float2 -- exactly the register pair {b0, b1} of mma.m16n8k8 -- and Blo the low parts (PB2 in float2
// units; PB2 % 16 == 4 keeps the 64-bit loads conflict-free). A weight matrix staged once per tile
// is read by every warp of the CTA: splitting it at staging time instead of at every fragment load
// takes the 24 split instructions of a k-step out of the inner loop, and the pair layout the
// register moves an interleaved (hi, lo) layout would need.
template <int KK, int NN, int BM, int PA, int PB2>
__device__ __forceinline__ void tile_mma_tc_b2(float *__restrict__ out, int PO, const float *__restrict__ As,
                                               const float2 *__restrict__ Bhi, const float2 *__restrict__ Blo) {
  using T = TcTile<BM, NN>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int rs = (warp % T::SLABS) * 16, cs = (warp / T::SLABS) * T::WN;
  float c[T::NF][4];
#pragma unroll
  for (int j = 0; j < T::NF; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
#pragma unroll 2
  for (int k8 = 0; k8 < KK; k8 += 8) {
    uint32_t ah[4], al[4];
    split_tf32_u(As[(rs + g) * PA + k8 + t], ah[0], al[0]);
    split_tf32_u(As[(rs + g + 8) * PA + k8 + t], ah[1], al[1]);
    split_tf32_u(As[(rs + g) * PA + k8 + t + 4], ah[2], al[2]);
    split_tf32_u(As[(rs + g + 8) * PA + k8 + t + 4], ah[3], al[3]);
#pragma unroll
    for (int j = 0; j < T::NF; ++j) {
      const int o = ((k8 >> 1) + t) * PB2 + cs + j * 8 + g;
      const float2 h = Bhi[o], l = Blo[o];
      const uint32_t bh[2] = {__float_as_uint(h.x), __float_as_uint(h.y)};
      const uint32_t bl[2] = {__float_as_uint(l.x), __float_as_uint(l.y)};
      mma_tf32_16x8x8(c[j], al, bh);
      mma_tf32_16x8x8(c[j], ah, bl);
      mma_tf32_16x8x8(c[j], ah, bh);
    }
  }
#pragma unroll
  for (int j = 0; j < T::NF; ++j) {
    const int n = cs + j * 8 + 2 * t;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      *reinterpret_cast<float2 *>(out + (size_t)(rs + g + 8 * h) * PO + n) = make_float2(c[j][2 * h], c[j][2 * h + 1]);
  }
}

// dw[a][b] += sum_m Z[m][n0 + a] * X[m][k0 + b] over the BM staged rows (pitches NN+4 / KK+4);
// db[a] += sum_m Z[m][n0 + a] in the threads with `with_db`. TN, TK in {2, 4, 8}.
template <int NN, int KK, int BM, int TN, int TK>
__device__ __forceinline__ void tile_outer(float (&dw)[TN][TK], float (&db)[TN], const float *__restrict__ Zs, int n0,
                                           const float *__restrict__ Xs, int k0, bool with_db) {
#pragma unroll 4
  for (int m = 0; m < BM; ++m) {
    float z[TN], x[TK];
#pragma unroll
    for (int a = 0; a < TN; a += (TN >= 4 ? 4 : TN)) {
      if constexpr (TN >= 4) {
        const float4 v = *reinterpret_cast<const float4 *>(Zs + m * (NN + 4) + n0 + a);
        z[a] = v.x; z[a + 1] = v.y; z[a + 2] = v.z; z[a + 3] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2 *>(Zs + m * (NN + 4) + n0 + a);
        z[a] = v.x; z[a + 1] = v.y;
      }
    }
#pragma unroll
    for (int b = 0; b < TK; b += (TK >= 4 ? 4 : TK)) {
      if constexpr (TK >= 4) {
        const float4 v = *reinterpret_cast<const float4 *>(Xs + m * (KK + 4) + k0 + b);
        x[b] = v.x; x[b + 1] = v.y; x[b + 2] = v.z; x[b + 3] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2 *>(Xs + m * (KK + 4) + k0 + b);
        x[b] = v.x; x[b + 1] = v.y;
      }
    }
#pragma unroll
    for (int a = 0; a < TN; ++a) {
      if (with_db) db[a] += z[a];
#pragma unroll
      for (int b = 0; b < TK; ++b) dw[a][b] = fmaf(z[a], x[b], dw[a][b]);
    }
  }
}

}  // namespace dense
}  // namespace mmrec
