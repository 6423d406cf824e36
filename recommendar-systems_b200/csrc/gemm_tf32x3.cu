// fp32-accurate tensor-core GEMM (3xTF32 error-compensated split) for the dense projections (K4):
//   C[M,N] = op(A) * op(B) (+ bias)      -- see include/mmrec_b200.h, mmrec_gemm_tf32x3_f32.
//
// The modality projections ([I x 4096] x [4096 x 64] and their two backward GEMMs) are 3.7 GFLOP
// each but only 117 MB of traffic: on fp32 CUDA cores they are compute-bound (~60 us), on tensor
// cores they become HBM-bound (~18 us). Plain TF32 would break the 1e-5 parity bound, so every
// operand is split in registers into hi = tf32(x) and lo = tf32(x - hi) and three MMAs
// (lo*hi + hi*lo + hi*hi, fp32 accumulate) reproduce fp32 products to ~2^-22.
//
// Three operand layouts cover forward, weight-gradient and input-gradient without ever
// transposing the 115 MB feature table:
//   A K-contiguous [M,K] or M-contiguous [K,M];  B K-contiguous [N,K] or N-contiguous [K,N].
// Tiles stream through a 3-stage cp.async ring (16-byte copies, zero-fill on the ragged edge);
// fragments are read with conflict-free scalar LDS thanks to the +4 / +8 word row padding.
// Split-K (grid.z) fills the 148 SMs when M*N is small (weight gradients); partial tiles go to a
// workspace and are summed in split order by a second kernel (deterministic, no atomics).
#include <stdlib.h>

#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kThreads = 256;
constexpr int BK = 32;
constexpr int kStages = 3;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int n = pred ? 16 : 0;   // src-size 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// hi = round-to-nearest tf32 (integer arithmetic on the bit pattern: cvt.rna.tf32.f32 without its
// inf/nan guards, 2 instructions instead of 4); lo = x - hi, left as fp32: the tensor core ignores
// the 13 low mantissa bits of a tf32 operand.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool KCONTIG, int ROWS>   // ROWS = BM or BN
struct TileLayout {
  // K-contiguous: [ROWS][BK + 4]; otherwise [BK][ROWS + 8]
  static constexpr int LD = KCONTIG ? BK + 4 : ROWS + 8;
  static constexpr int FLOATS = KCONTIG ? ROWS * LD : BK * LD;
  __device__ static __forceinline__ int at(int r, int k) { return KCONTIG ? r * LD + k : k * LD + r; }
};

// Copy one operand tile (rows r0.., k-range k0..) into shared memory.
template <bool KCONTIG, int ROWS>
__device__ __forceinline__ void load_tile(float *s, const float *__restrict__ g, int n_rows, int n_k, int r0,
                                          int k0, int k_end) {
  using L = TileLayout<KCONTIG, ROWS>;
  if constexpr (KCONTIG) {
    // global [n_rows][n_k]: 16-byte chunks along k
    constexpr int CH = BK / 4;
    for (int t = threadIdx.x; t < ROWS * CH; t += kThreads) {
      const int r = t / CH, c = (t % CH) * 4;
      const bool ok = (r0 + r < n_rows) && (k0 + c < k_end);
      const float *src = ok ? g + (size_t)(r0 + r) * n_k + k0 + c : g;
      cp_async16(s + r * L::LD + c, src, ok);
    }
  } else {
    // global [n_k][n_rows]: 16-byte chunks along the row index
    constexpr int CH = ROWS / 4;
    for (int t = threadIdx.x; t < BK * CH; t += kThreads) {
      const int k = t / CH, c = (t % CH) * 4;
      const bool ok = (k0 + k < k_end) && (r0 + c < n_rows);
      const float *src = ok ? g + (size_t)(k0 + k) * n_rows + r0 + c : g;
      cp_async16(s + k * L::LD + c, src, ok);
    }
  }
}

template <bool A_K, bool B_K, int BM, int BN>
__global__ void __launch_bounds__(kThreads)
gemm_tf32x3_kernel(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ bias,
                   float *__restrict__ C, float *__restrict__ ws, int M, int N, int K, int k_per_split) {
  using LA = TileLayout<A_K, BM>;
  using LB = TileLayout<B_K, BN>;
  constexpr int WARPS_M = BM / 32, WARPS_N = 8 / WARPS_M, WN = BN / WARPS_N, NF = WN / 8;
  extern __shared__ __align__(16) float smem[];
  float *sA = smem, *sB = smem + kStages * LA::FLOATS;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  const int n_kb = (k_end - k_begin + BK - 1) / BK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = (warp % WARPS_M) * 32, wn = (warp / WARPS_M) * WN;

  float acc[2][NF][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

  auto issue = [&](int kb) {
    const int stage = kb % kStages;
    load_tile<A_K, BM>(sA + stage * LA::FLOATS, A, M, K, m0, k_begin + kb * BK, k_end);
    load_tile<B_K, BN>(sB + stage * LB::FLOATS, B, N, K, n0, k_begin + kb * BK, k_end);
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < n_kb) issue(s);
    cp_async_commit();
  }
  for (int kb = 0; kb < n_kb; ++kb) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    if (kb + kStages - 1 < n_kb) issue(kb + kStages - 1);
    cp_async_commit();
    const float *a = sA + (kb % kStages) * LA::FLOATS;
    const float *b = sB + (kb % kStages) * LB::FLOATS;
    // per-k-block accumulator: the tensor core truncates when it accumulates, so long K chains
    // are summed with round-to-nearest FADDs every 32 columns instead
    float part[2][NF][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) part[i][j][q] = 0.f;
#pragma unroll
    for (int k8 = 0; k8 < BK; k8 += 8) {
      uint32_t ah[2][4], al[2][4], bh[NF][2], bl[NF][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = wm + i * 16 + g;
        split_tf32(a[LA::at(r, k8 + t)], ah[i][0], al[i][0]);
        split_tf32(a[LA::at(r + 8, k8 + t)], ah[i][1], al[i][1]);
        split_tf32(a[LA::at(r, k8 + t + 4)], ah[i][2], al[i][2]);
        split_tf32(a[LA::at(r + 8, k8 + t + 4)], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const int c = wn + j * 8 + g;
        split_tf32(b[LB::at(c, k8 + t)], bh[j][0], bl[j][0]);
        split_tf32(b[LB::at(c, k8 + t + 4)], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          mma_tf32(part[i][j], al[i], bh[j]);
          mma_tf32(part[i][j], ah[i], bl[j]);
          mma_tf32(part[i][j], ah[i], bh[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] += part[i][j][q];
  }
  cp_async_wait<0>();
  // epilogue: direct store (+bias) or partial tile to the split-K workspace
  float *out = gridDim.z == 1 ? C : ws + (size_t)blockIdx.z * M * N;
  const bool add_bias = gridDim.z == 1 && bias != nullptr;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const int c = n0 + wn + j * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = m0 + wm + i * 16 + g + 8 * h;
        if (r >= M) continue;
        float v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        if (c + 1 < N) {
          if (add_bias) { v0 += bias[c]; v1 += bias[c + 1]; }
          *reinterpret_cast<float2 *>(out + (size_t)r * N + c) = make_float2(v0, v1);
        } else if (c < N) {
          if (add_bias) v0 += bias[c];
          out[(size_t)r * N + c] = v0;
        }
      }
    }
}

__global__ void __launch_bounds__(kThreads)
splitk_reduce_kernel(const float *__restrict__ ws, const float *__restrict__ bias, float *__restrict__ C,
                     int64_t mn, int N, int splits) {
  const int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
  if (i >= mn) return;
  float4 s = *reinterpret_cast<const float4 *>(ws + i);
  for (int z = 1; z < splits; ++z) {
    const float4 v = *reinterpret_cast<const float4 *>(ws + (size_t)z * mn + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  if (bias) {
    const int c = (int)(i % N);
    s.x += bias[c]; s.y += bias[c + 1]; s.z += bias[c + 2]; s.w += bias[c + 3];
  }
  *reinterpret_cast<float4 *>(C + i) = s;
}

template <bool A_K, bool B_K, int BM, int BN>
int launch(const float *A, const float *B, const float *bias, float *C, float *ws, int M, int N, int K,
           int splits, cudaStream_t stream) {
  using LA = TileLayout<A_K, BM>;
  using LB = TileLayout<B_K, BN>;
  const size_t smem = (size_t)kStages * (LA::FLOATS + LB::FLOATS) * sizeof(float);
  auto kern = gemm_tf32x3_kernel<A_K, B_K, BM, BN>;
  MMREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int k_per_split = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (K + k_per_split - 1) / k_per_split;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  kern<<<grid, kThreads, smem, stream>>>(A, B, bias, C, ws, M, N, K, k_per_split);
  MMREC_CHECK_LAUNCH("gemm_tf32x3_kernel");
  if (splits > 1) {
    const int64_t mn = (int64_t)M * N;
    splitk_reduce_kernel<<<(unsigned)((mn / 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(ws, bias, C, mn, N,
                                                                                                 splits);
    MMREC_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return MMREC_OK;
}

}  // namespace

// gemm_tc05.cu: the tcgen05 path for the table-sized projections
int gemm_tc05_kind(int M, int N, int K, int a_kcontig, int b_kcontig);
int gemm_tc05_splits(int M, int N, int K, int kind);
int gemm_tc05_dispatch(const float *A, int a_kcontig, const float *B, int b_kcontig, const float *bias, float *C,
                       int M, int N, int K, int splits, float *ws, cudaStream_t stream, int act);
static bool tc05_enabled() {
  static const bool on = !(getenv("MMREC_GEMM_TC") && atoi(getenv("MMREC_GEMM_TC")) == 0);
  return on;
}
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_gemm_splits(int32_t M, int32_t N, int32_t K, int32_t a_kcontig, int32_t b_kcontig) {
  if (tc05_enabled()) {
    const int kind = gemm_tc05_kind(M, N, K, a_kcontig, b_kcontig);
    if (kind) return gemm_tc05_splits(M, N, K, kind);
  }
  // enough CTAs for ~2 waves of 148 SMs; never split a short K
  const int bm = (!a_kcontig || M <= 64) ? 64 : 128;
  const int bn = (!b_kcontig && N > 64) ? 128 : 64;
  const long tiles = (long)((M + bm - 1) / bm) * ((N + bn - 1) / bn);
  long s = (2 * kNumSMs + tiles - 1) / tiles;
  const long max_by_k = (K + 4 * BK - 1) / (4 * BK);
  if (s > max_by_k) s = max_by_k;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return (int)s;
}

extern "C" int mmrec_gemm_tf32x3_f32(const float *A, int32_t a_kcontig, const float *B, int32_t b_kcontig,
                                     const float *bias, float *C, int32_t M, int32_t N, int32_t K, int32_t splits,
                                     float *ws, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MMREC_REQUIRE(A && B && C, MMREC_E_BADARG, "gemm: null pointer");
  MMREC_REQUIRE(M > 0 && N > 0 && K > 0 && splits >= 1, MMREC_E_BADARG, "gemm: bad sizes");
  MMREC_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C) && aligned16(ws), MMREC_E_ALIGN,
                "gemm: operands must be 16-byte aligned");
  MMREC_REQUIRE(N % 4 == 0, MMREC_E_BADARG, "gemm: N=%d must be a multiple of 4", N);
  MMREC_REQUIRE((a_kcontig ? K : M) % 4 == 0 && (b_kcontig ? K : N) % 4 == 0, MMREC_E_BADARG,
                "gemm: contiguous dimensions must be multiples of 4 floats");
  MMREC_REQUIRE(splits == 1 || ws, MMREC_E_WORKSPACE, "gemm: split-K needs a workspace of splits*M*N floats");
  if (tc05_enabled()) {     // tcgen05 kernel for the table-sized projections; 1 = shape not covered
    const int rc = gemm_tc05_dispatch(A, a_kcontig, B, b_kcontig, bias, C, M, N, K, splits, ws, stream, 0);
    if (rc <= 0) {
      if (rc == MMREC_OK && splits > 1) {
        const int64_t mn = (int64_t)M * N;
        splitk_reduce_kernel<<<(unsigned)((mn / 4 + kThreads - 1) / kThreads), kThreads, 0, stream>>>(ws, bias, C, mn,
                                                                                                     N, splits);
        MMREC_CHECK_LAUNCH("splitk_reduce_kernel");
      }
      return rc;
    }
  }
  const bool big_m = a_kcontig && M > 64;
  const bool big_n = !b_kcontig && N > 64;
  if (a_kcontig && b_kcontig) {
    return big_m ? launch<true, true, 128, 64>(A, B, bias, C, ws, M, N, K, splits, stream)
                 : launch<true, true, 64, 64>(A, B, bias, C, ws, M, N, K, splits, stream);
  }
  if (a_kcontig && !b_kcontig) {
    if (big_m) return big_n ? launch<true, false, 128, 128>(A, B, bias, C, ws, M, N, K, splits, stream)
                            : launch<true, false, 128, 64>(A, B, bias, C, ws, M, N, K, splits, stream);
    return big_n ? launch<true, false, 64, 128>(A, B, bias, C, ws, M, N, K, splits, stream)
                 : launch<true, false, 64, 64>(A, B, bias, C, ws, M, N, K, splits, stream);
  }
  if (!a_kcontig && !b_kcontig) {
    return big_n ? launch<false, false, 64, 128>(A, B, bias, C, ws, M, N, K, splits, stream)
                 : launch<false, false, 64, 64>(A, B, bias, C, ws, M, N, K, splits, stream);
  }
  set_error("gemm: layout (A M-contiguous, B K-contiguous) is not instantiated");
  return MMREC_E_BADARG;
}

// ---- Linear + activation over many rows on the tcgen05 kernel (d x d layers at d = 128) --------
namespace mmrec {
namespace {
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ y, int64_t n4, int act, float4 *__restrict__ dz) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 g = dy[i], v = y[i];
  float4 o;
  if (act == 1) o = make_float4(g.x * (1.f - v.x * v.x), g.y * (1.f - v.y * v.y), g.z * (1.f - v.z * v.z), g.w * (1.f - v.w * v.w));
  else o = make_float4(g.x * ((1.f - v.x) * v.x), g.y * ((1.f - v.y) * v.y), g.z * ((1.f - v.z) * v.z), g.w * ((1.f - v.w) * v.w));
  dz[i] = o;
}
}  // namespace
}  // namespace mmrec

extern "C" int mmrec_linear_act_tc_supported(int32_t M, int32_t K, int32_t N) {
  return tc05_enabled() && gemm_tc05_kind(M, N, K, 1, 1) == 1 && gemm_tc05_splits(M, N, K, 1) == 1;
}

extern "C" int mmrec_linear_act_tc_f32(const float *x, const float *W, const float *b, float *y, int32_t M, int32_t K,
                                       int32_t N, int32_t act, void *stream_) {
  MMREC_REQUIRE(x && W && y, MMREC_E_BADARG, "linear_act_tc: null pointer");
  MMREC_REQUIRE(act >= 0 && act <= 2, MMREC_E_BADARG, "linear_act_tc: act must be 0 (none), 1 (tanh) or 2 (sigmoid)");
  MMREC_REQUIRE(aligned16(x) && aligned16(W) && aligned16(y) && aligned16(b), MMREC_E_ALIGN,
                "linear_act_tc: operands must be 16-byte aligned");
  MMREC_REQUIRE(mmrec_linear_act_tc_supported(M, K, N), MMREC_E_BADARG,
                "linear_act_tc: shape %d x %d -> %d is not covered by the tcgen05 kernel without split-K", M, K, N);
  const int rc = gemm_tc05_dispatch(x, 1, W, 1, b, y, M, N, K, 1, nullptr, (cudaStream_t)stream_, act);
  if (rc > 0) {
    set_error("linear_act_tc: dispatch refused the shape");
    return MMREC_E_BADARG;
  }
  return rc;
}

extern "C" int mmrec_act_bwd_f32(const float *dy, const float *y, int64_t numel, int32_t act, float *dz, void *stream_) {
  MMREC_REQUIRE(dy && y && dz, MMREC_E_BADARG, "act_bwd: null pointer");
  MMREC_REQUIRE(numel > 0 && numel % 4 == 0 && (act == 1 || act == 2), MMREC_E_BADARG, "act_bwd: bad arguments");
  MMREC_REQUIRE(aligned16(dy) && aligned16(y) && aligned16(dz), MMREC_E_ALIGN, "act_bwd: operands must be 16-byte aligned");
  const int64_t n4 = numel / 4;
  act_bwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      (const float4 *)dy, (const float4 *)y, n4, act, (float4 *)dz);
  MMREC_CHECK_LAUNCH("act_bwd_kernel");
  return MMREC_OK;
}
