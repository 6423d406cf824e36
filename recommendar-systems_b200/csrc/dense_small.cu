// d x d dense layers of the modality side networks, fused with bias + activation (K4b):
//   forward   Y = act(X W^T + b)                                  one launch
//   backward  dZ = dY * act'(Y);  dX = dZ W;  dW = dZ^T X;  db = sum_m dZ     one launch + reduce
// for X [M, K], W [N, K] with K, N in {32, 64, 128} (the embedding width): gate_v/t/f, query_v/t,
// gate_*_prefer of SMORE (smore.py:265-272, 321-330) and the MGCN gates (mgcn.py:153-154,
// 188-203). These GEMMs are tall and skinny (M = 7k..60k rows, K = N = 64): 0.2 GFLOP and 14 MB
// each; the tile products run on mma.sync tensor cores with the 3xTF32 split (dense_tile.cuh), and
// what matters most is one pass over the operands, no intermediate tensors and few launches.
//
// Tiling: 256 threads = (256 / (N/4)) row groups x (N/4) column groups; a thread owns TM rows x 4
// columns. X tiles are staged row-major in shared memory (pitch K+4, conflict-free float4 reads:
// a warp touches at most 4 rows, 4 banks apart), W is staged once per CTA k-major so that the 4
// output columns of a thread are one float4. The backward shares the dZ / X tiles between the
// dX product and the dW outer-product accumulation; per-CTA dW / db partials are summed by a
// second small kernel in a fixed order (no floating-point atomics: bit-reproducible). Grids are
// sized so that the whole layer is one wave of 3 CTAs per SM (64-row tiles): loads of one CTA
// overlap the FMA loop of its neighbours.
#include "dense_tile.cuh"

namespace mmrec {
namespace {

using namespace dense;

constexpr int kWP = 8;   // extra pitch of the staged weight matrix (conflict-free MMA B-fragment reads)

template <int ACT>
struct EpiBiasAct {
  const float *bias;
  __device__ __forceinline__ float2 operator()(float2 v, int n) const {
    if (bias != nullptr) { v.x += __ldg(bias + n); v.y += __ldg(bias + n + 1); }
    return make_float2(act_fwd<ACT>(v.x), act_fwd<ACT>(v.y));
  }
};

// Up to kMaxBatch independent layers of the same shape and activation go out as ONE launch
// (blockIdx.y = layer): SMORE's three modality gates gate_v / gate_t / gate_f (smore.py:269-272)
// are 7050-row problems of half a wave each -- launch-bound one by one.
constexpr int kMaxBatch = 4;
struct FwdBatch {
  const float *X[kMaxBatch], *W[kMaxBatch], *bias[kMaxBatch];
  float *Y[kMaxBatch];
};
struct BwdBatch {
  const float *dY[kMaxBatch], *Y[kMaxBatch], *X[kMaxBatch], *W[kMaxBatch];
  float *dX[kMaxBatch], *partial[kMaxBatch], *dW[kMaxBatch], *db[kMaxBatch];
};

template <int K, int N, int TM, int ACT>
__global__ void __launch_bounds__(kT, K <= 64 ? 3 : 2)
dense_fwd_kernel(const __grid_constant__ FwdBatch B, int M, int n_tiles) {
  using T = Tile<K, N, TM>;
  const float *__restrict__ X = B.X[blockIdx.y], *__restrict__ W = B.W[blockIdx.y],
                           *__restrict__ bias = B.bias[blockIdx.y];
  float *__restrict__ Y = B.Y[blockIdx.y];
  extern __shared__ float4 smem4[];
  float *Ws = reinterpret_cast<float *>(smem4);      // [K][N + 8]   Ws[k][n] = W[n][k]
  float *Xs = Ws + K * (N + kWP);                    // [BM][K + 4]
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * T::BM;
    RowStage<K, T::BM> xs;
    xs.load(X, m0, M);
    if (tile == (int)blockIdx.x) {
      // first tile: transpose W into shared memory while the X rows are in flight
      constexpr int WPER = N * (K / 4) / kT;
      float4 w[WPER];
#pragma unroll
      for (int i = 0; i < WPER; ++i) {
        const int idx = threadIdx.x + i * kT, n = idx % N, k4 = idx / N;
        w[i] = ldg4(W + (size_t)n * K + k4 * 4);
      }
#pragma unroll
      for (int i = 0; i < WPER; ++i) {
        const int idx = threadIdx.x + i * kT, n = idx % N, k4 = idx / N;
        Ws[(k4 * 4 + 0) * (N + kWP) + n] = w[i].x;
        Ws[(k4 * 4 + 1) * (N + kWP) + n] = w[i].y;
        Ws[(k4 * 4 + 2) * (N + kWP) + n] = w[i].z;
        Ws[(k4 * 4 + 3) * (N + kWP) + n] = w[i].w;
      }
    } else {
      __syncthreads();                               // previous tile fully consumed
    }
    xs.store(Xs);
    __syncthreads();
    // Y tile = act(X tile * W^T + b) on the tensor cores, written from the accumulator fragments
    tile_mma_tc<K, N, T::BM, K + 4, N + kWP, false>(Y + (size_t)m0 * N, N, M - m0, Xs, Ws, false,
                                                    EpiBiasAct<ACT>{bias});
  }
}

// Backward. partial layout per CTA: [N*K] dW then [N] db.
template <int K, int N, int TM, int ACT>
__global__ void __launch_bounds__(kT, K <= 64 ? 3 : 1)
dense_bwd_kernel(const __grid_constant__ BwdBatch B, int M, int n_tiles) {
  constexpr int CGK = K / 4, RGK = kT / CGK, BM = RGK * TM;
  const float *__restrict__ dY = B.dY[blockIdx.y], *__restrict__ Y = B.Y[blockIdx.y],
                           *__restrict__ X = B.X[blockIdx.y], *__restrict__ W = B.W[blockIdx.y];
  float *__restrict__ dX = B.dX[blockIdx.y], *__restrict__ partial = B.partial[blockIdx.y];
  extern __shared__ float4 smem4[];
  float *Wn = reinterpret_cast<float *>(smem4);     // [N][K + 8]   natural layout (n-major)
  float *Zs = Wn + N * (K + kWP);                   // [BM][N + 4]  dZ tile
  float *Xs = Zs + BM * (N + 4);                    // [BM][K + 4]  X tile
  float *p = partial + (size_t)blockIdx.x * (N * K + N);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int m0 = tile * BM;
    const bool first = tile == (int)blockIdx.x;
    // every global load of the tile (and, first time, of W) is in flight before the first store
    RowStage<N, BM> gs, ys;
    RowStage<K, BM> xs;
    static_assert(K == N, "square layers only");
    gs.load(dY, m0, M);
    xs.load(X, m0, M);
    if constexpr (ACT != kNone) ys.load(Y, m0, M);
    if (first) {
      constexpr int WPER = N * K / 4 / kT;
      float4 w[WPER];
#pragma unroll
      for (int i = 0; i < WPER; ++i) w[i] = ldg4(W + (threadIdx.x + i * kT) * 4);
#pragma unroll
      for (int i = 0; i < WPER; ++i) {
        const int idx = threadIdx.x + i * kT, n = idx / (K / 4), k4 = idx % (K / 4);
        *reinterpret_cast<float4 *>(Wn + n * (K + kWP) + k4 * 4) = w[i];
      }
    } else {
      __syncthreads();                               // previous tile fully consumed
    }
    if constexpr (ACT != kNone) {                    // dZ = dY * act'(Y)
#pragma unroll
      for (int i = 0; i < RowStage<N, BM>::PER; ++i) {
        gs.v[i].x = act_bwd<ACT>(gs.v[i].x, ys.v[i].x); gs.v[i].y = act_bwd<ACT>(gs.v[i].y, ys.v[i].y);
        gs.v[i].z = act_bwd<ACT>(gs.v[i].z, ys.v[i].z); gs.v[i].w = act_bwd<ACT>(gs.v[i].w, ys.v[i].w);
      }
    }
    gs.store(Zs);
    xs.store(Xs);
    __syncthreads();
    // dX tile = dZ W  (rows of the tile x K columns, reduction over N)
    if (dX != nullptr)
      tile_mma_tc<N, K, BM, N + 4, K + kWP, false>(dX + (size_t)m0 * K, K, M - m0, Zs, Wn);
    // dW (+)= dZ^T X  (N x K, reduction over the rows of the tile; rows beyond M are zero in both)
    tile_mma_tc<BM, K, N, N + 4, K + 4, true>(p, K, N, Zs, Xs, !first);
    if (threadIdx.x < N) {                           // db (+)= column sums of dZ
      float sdb = 0.f;
#pragma unroll 8
      for (int m = 0; m < BM; ++m) sdb += Zs[m * (N + 4) + threadIdx.x];
      p[N * K + threadIdx.x] = first ? sdb : p[N * K + threadIdx.x] + sdb;
    }
  }
}

// out[j] = sum over CTAs of partial[cta][j]; j < n_w -> dW, the rest -> db. 64 outputs per CTA,
// 16 thread rows each summing every 16th partial (independent coalesced loads), combined in a
// fixed order through shared memory: deterministic, and two memory round trips deep.
__global__ void __launch_bounds__(1024)
dense_partial_reduce_kernel(const __grid_constant__ BwdBatch B, int n_parts, int n_w, int n_b) {
  __shared__ float sm[16][64];
  const float *__restrict__ partial = B.partial[blockIdx.y];
  float *__restrict__ dW = B.dW[blockIdx.y], *__restrict__ db = B.db[blockIdx.y];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + tx;
  const size_t stride = (size_t)n_w + n_b;
  float s[8];                                  // eight independent loads in flight per thread
#pragma unroll
  for (int u = 0; u < 8; ++u) s[u] = 0.f;
  if (j < n_w + n_b) {
    int p = ty;
    for (; p + 16 * 7 < n_parts; p += 16 * 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += __ldcs(partial + (size_t)(p + 16 * u) * stride + j);
    }
    for (; p < n_parts; p += 16) s[0] += __ldcs(partial + (size_t)p * stride + j);
  }
  sm[ty][tx] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  __syncthreads();
  if (ty == 0 && j < n_w + n_b) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) s += sm[g][tx];
    if (j < n_w) dW[j] = s;
    else if (db != nullptr) db[j - n_w] = s;
  }
}

template <int D> struct RowsPerThread { static constexpr int value = D == 128 ? 8 : 4; };
constexpr int kMaxParts = 4 * kNumSMs;   // cap on CTAs (= dW partials) of the backward

template <int D, int TM>
constexpr size_t fwd_smem() { return sizeof(float) * (D * (D + kWP) + Tile<D, D, TM>::BM * (D + 4)); }
template <int D, int TM>
constexpr size_t bwd_smem() { return sizeof(float) * (D * (D + kWP) + 2 * Tile<D, D, TM>::BM * (D + 4)); }

template <int D, int ACT>
int launch_fwd(const FwdBatch &B, int n_batch, int M, cudaStream_t st) {
  constexpr int TM = RowsPerThread<D>::value;
  using T = Tile<D, D, TM>;
  constexpr size_t smem = fwd_smem<D, TM>();
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(dense_fwd_kernel<D, D, TM, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int n_tiles = (M + T::BM - 1) / T::BM;
  const int grid = min(n_tiles, 8 * kNumSMs);
  dense_fwd_kernel<D, D, TM, ACT><<<dim3(grid, n_batch), kT, smem, st>>>(B, M, n_tiles);
  MMREC_CHECK_LAUNCH("dense_fwd_kernel");
  return MMREC_OK;
}

template <int D, int ACT>
int launch_bwd(BwdBatch B, int n_batch, float *ws, int M, cudaStream_t st) {
  constexpr int TM = RowsPerThread<D>::value;
  using T = Tile<D, D, TM>;
  constexpr size_t smem = bwd_smem<D, TM>();
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(dense_bwd_kernel<D, D, TM, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int n_tiles = (M + T::BM - 1) / T::BM;
  const int grid = max(1, min(n_tiles, kMaxParts));
  for (int i = 0; i < n_batch; ++i) B.partial[i] = ws + (size_t)i * kMaxParts * (D * D + D);
  dense_bwd_kernel<D, D, TM, ACT><<<dim3(grid, n_batch), kT, smem, st>>>(B, M, n_tiles);
  MMREC_CHECK_LAUNCH("dense_bwd_kernel");
  const int n_out = D * D + D;
  dense_partial_reduce_kernel<<<dim3((n_out + 63) / 64, n_batch), 1024, 0, st>>>(B, grid, D * D, D);
  MMREC_CHECK_LAUNCH("dense_partial_reduce_kernel");
  return MMREC_OK;
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_dense_act_supported(int32_t K, int32_t N) {
  return K == N && (K == 32 || K == 64 || K == 128);
}

extern "C" size_t mmrec_dense_act_bwd_workspace_bytes(int32_t K, int32_t N) {
  return sizeof(float) * (size_t)kMaxParts * ((size_t)N * K + N);
}

#define MMREC_DENSE_DISPATCH(FN, ...)                          \
  switch (K * 4 + act) {                                       \
    case 32 * 4 + 0: return FN<32, kNone>(__VA_ARGS__);        \
    case 32 * 4 + 1: return FN<32, kTanh>(__VA_ARGS__);        \
    case 32 * 4 + 2: return FN<32, kSigmoid>(__VA_ARGS__);     \
    case 64 * 4 + 0: return FN<64, kNone>(__VA_ARGS__);        \
    case 64 * 4 + 1: return FN<64, kTanh>(__VA_ARGS__);        \
    case 64 * 4 + 2: return FN<64, kSigmoid>(__VA_ARGS__);     \
    case 128 * 4 + 0: return FN<128, kNone>(__VA_ARGS__);      \
    case 128 * 4 + 1: return FN<128, kTanh>(__VA_ARGS__);      \
    case 128 * 4 + 2: return FN<128, kSigmoid>(__VA_ARGS__);   \
  }

extern "C" int mmrec_dense_act_fwd_f32(const float *X, const float *W, const float *bias, float *Y, int32_t M,
                                       int32_t K, int32_t N, int32_t act, void *stream) {
  MMREC_REQUIRE(X && W && Y, MMREC_E_BADARG, "dense_act_fwd: null pointer");
  MMREC_REQUIRE(mmrec_dense_act_supported(K, N), MMREC_E_BADARG,
                "dense_act_fwd: K = N in {32, 64, 128} required (got K=%d N=%d)", K, N);
  MMREC_REQUIRE(act >= 0 && act <= 2 && M >= 0, MMREC_E_BADARG, "dense_act_fwd: bad act / M");
  MMREC_REQUIRE(aligned16(X) && aligned16(W) && aligned16(Y) && aligned16(bias), MMREC_E_ALIGN,
                "dense_act_fwd: operands must be 16-byte aligned");
  if (M == 0) return MMREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FwdBatch B{};
  B.X[0] = X; B.W[0] = W; B.bias[0] = bias; B.Y[0] = Y;
  MMREC_DENSE_DISPATCH(launch_fwd, B, 1, M, st);
  return MMREC_E_BADARG;
}

extern "C" int mmrec_dense_act_batch_fwd_f32(const float *const *X_host, const float *const *W_host,
                                             const float *const *bias_host, float *const *Y_host, int32_t n_batch,
                                             int32_t M, int32_t K, int32_t N, int32_t act, void *stream) {
  MMREC_REQUIRE(X_host && W_host && bias_host && Y_host, MMREC_E_BADARG, "dense_act_batch_fwd: null pointer");
  MMREC_REQUIRE(n_batch >= 1 && n_batch <= kMaxBatch, MMREC_E_BADARG, "dense_act_batch_fwd: 1 <= n_batch <= %d", kMaxBatch);
  MMREC_REQUIRE(mmrec_dense_act_supported(K, N), MMREC_E_BADARG,
                "dense_act_batch_fwd: K = N in {32, 64, 128} required (got K=%d N=%d)", K, N);
  MMREC_REQUIRE(act >= 0 && act <= 2 && M > 0, MMREC_E_BADARG, "dense_act_batch_fwd: bad act / M");
  FwdBatch B{};
  for (int i = 0; i < n_batch; ++i) {
    B.X[i] = X_host[i]; B.W[i] = W_host[i]; B.bias[i] = bias_host[i]; B.Y[i] = Y_host[i];
    MMREC_REQUIRE(B.X[i] && B.W[i] && B.Y[i], MMREC_E_BADARG, "dense_act_batch_fwd: null tensor %d", i);
    MMREC_REQUIRE(aligned16(B.X[i]) && aligned16(B.W[i]) && aligned16(B.Y[i]) && aligned16(B.bias[i]), MMREC_E_ALIGN,
                  "dense_act_batch_fwd: operands must be 16-byte aligned");
  }
  cudaStream_t st = (cudaStream_t)stream;
  MMREC_DENSE_DISPATCH(launch_fwd, B, n_batch, M, st);
  return MMREC_E_BADARG;
}

extern "C" int mmrec_dense_act_bwd_f32(const float *dY, const float *Y, const float *X, const float *W,
                                       float *dX, float *dW, float *db, float *ws, int32_t M, int32_t K,
                                       int32_t N, int32_t act, void *stream) {
  MMREC_REQUIRE(dY && X && W && dW && ws, MMREC_E_BADARG, "dense_act_bwd: null pointer");
  MMREC_REQUIRE(act == 0 || Y != nullptr, MMREC_E_BADARG, "dense_act_bwd: the activation needs Y");
  MMREC_REQUIRE(mmrec_dense_act_supported(K, N), MMREC_E_BADARG,
                "dense_act_bwd: K = N in {32, 64, 128} required (got K=%d N=%d)", K, N);
  MMREC_REQUIRE(act >= 0 && act <= 2 && M >= 0, MMREC_E_BADARG, "dense_act_bwd: bad act / M");
  MMREC_REQUIRE(aligned16(dY) && aligned16(Y) && aligned16(X) && aligned16(W) && aligned16(dX), MMREC_E_ALIGN,
                "dense_act_bwd: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) {
    MMREC_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * N * K, st));
    if (db) MMREC_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
    return MMREC_OK;
  }
  BwdBatch B{};
  B.dY[0] = dY; B.Y[0] = Y; B.X[0] = X; B.W[0] = W; B.dX[0] = dX; B.dW[0] = dW; B.db[0] = db;
  MMREC_DENSE_DISPATCH(launch_bwd, B, 1, ws, M, st);
  return MMREC_E_BADARG;
}

extern "C" int mmrec_dense_act_batch_bwd_f32(const float *const *dY_host, const float *const *Y_host,
                                             const float *const *X_host, const float *const *W_host,
                                             float *const *dX_host, float *const *dW_host, float *const *db_host,
                                             float *ws, int32_t n_batch, int32_t M, int32_t K, int32_t N, int32_t act,
                                             void *stream) {
  MMREC_REQUIRE(dY_host && Y_host && X_host && W_host && dX_host && dW_host && db_host && ws, MMREC_E_BADARG,
                "dense_act_batch_bwd: null pointer");
  MMREC_REQUIRE(n_batch >= 1 && n_batch <= kMaxBatch, MMREC_E_BADARG, "dense_act_batch_bwd: 1 <= n_batch <= %d", kMaxBatch);
  MMREC_REQUIRE(mmrec_dense_act_supported(K, N), MMREC_E_BADARG,
                "dense_act_batch_bwd: K = N in {32, 64, 128} required (got K=%d N=%d)", K, N);
  MMREC_REQUIRE(act >= 0 && act <= 2 && M > 0, MMREC_E_BADARG, "dense_act_batch_bwd: bad act / M");
  BwdBatch B{};
  for (int i = 0; i < n_batch; ++i) {
    B.dY[i] = dY_host[i]; B.Y[i] = Y_host[i]; B.X[i] = X_host[i]; B.W[i] = W_host[i];
    B.dX[i] = dX_host[i]; B.dW[i] = dW_host[i]; B.db[i] = db_host[i];
    MMREC_REQUIRE(B.dY[i] && B.X[i] && B.W[i] && B.dW[i] && (act == 0 || B.Y[i]), MMREC_E_BADARG,
                  "dense_act_batch_bwd: null tensor %d", i);
    MMREC_REQUIRE(aligned16(B.dY[i]) && aligned16(B.Y[i]) && aligned16(B.X[i]) && aligned16(B.W[i]) && aligned16(B.dX[i]),
                  MMREC_E_ALIGN, "dense_act_batch_bwd: operands must be 16-byte aligned");
  }
  cudaStream_t st = (cudaStream_t)stream;
  MMREC_DENSE_DISPATCH(launch_bwd, B, n_batch, ws, M, st);
  return MMREC_E_BADARG;
}
