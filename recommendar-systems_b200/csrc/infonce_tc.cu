// InfoNCE (K7) on the 5th-generation tensor cores for d = 64: forward row sums and both backward
// passes as ONE tcgen05 kernel template (mgcn.py:224-231, smore.py:380-387).
//
// A CTA owns 128 "own" rows (TMEM lane = row) and walks a range of 64-row "other" tiles:
//   GEMM 1   S = Own Other^T            tcgen05.mma kind::tf32, 3xTF32, A and B from shared memory
//   epilogue (warps 0-7: two warps per TMEM lane quadrant, each takes 32 of a tile's 64 columns
//             -- the epilogue is a latency chain per warp, so two warps per scheduler):
//             tcgen05.ld S ->
//              forward : rowsum += exp(s / t), the diagonal score
//              backward: g = coef/B/t * (exp(s / t) / ttl - delta), split into tf32 hi / lo and
//                        written back to TENSOR MEMORY with tcgen05.st as the A operand of
//   GEMM 2   dOwn += G Other            tcgen05.mma in its A-from-TMEM form, B = the same other
//                                       tile in MN-major layout; the accumulator stays in TMEM for
//                                       the whole CTA and is stored once at the end.
// The B x B matrices S and G never exist outside TMEM. Producers (two independent groups of two
// warps, one tile in flight each) write every other tile twice -- K-major SWIZZLE_128B for GEMM 1,
// MN-major SWIZZLE_128B_BASE32B for GEMM 2 -- both as hi / lo halves. The MMA warp is software
// pipelined: GEMM 1 of tile t + 1 is issued before GEMM 2 of tile t, so the epilogue of one tile
// overlaps the tensor-core work of its neighbours (S and G are double buffered in TMEM).
// The mma.sync version (loss.cu) needed 23 + 2 x 35 us per call at B = 2048; it stays the path for
// d = 32 / 128.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace mmrec {
namespace {

using namespace tc05;

constexpr int kTM = 128;         // own rows per CTA (UMMA M)
constexpr int kTN = 64;          // other rows per tile
constexpr int kThreadsI = 416;   // 8 epilogue + 4 producer + 1 MMA warp
constexpr int kProdWarp0 = 8, kMmaWarp = 12;
constexpr int kStages = 2;
constexpr int kGroupThreads = 64;

enum Mode { kFwd = 0, kBwdRow = 1, kBwdCol = 2 };

// MN-major SWIZZLE_128B_BASE32B tile (extent E floats along N, K rows): byte offset of the 16-byte
// chunk q (4 floats of N) of K row k. Atoms of 4 K rows x 128 B, [K group][N atom] order.
template <int E>
__device__ __forceinline__ uint32_t mn32_off(int k, int q) {
  return (uint32_t)((k >> 2) * (E / 32) + (q >> 3)) * 512u + (uint32_t)(k & 3) * 128u +
         (uint32_t)(((((q & 7) >> 1) ^ (k & 3)) << 5) | ((q & 1) << 4));
}

template <int D>
struct ICfg {
  static constexpr int KB = D / 32;
  static constexpr uint32_t A_HALF = KB * kTM * 128;       // hi (or lo) own tile, K-major
  static constexpr uint32_t B1_HALF = KB * kTN * 128;      // hi (or lo) other tile, K-major (GEMM 1)
  static constexpr uint32_t B2_HALF = kTN * D * 4;         // hi (or lo) other tile, MN-major (GEMM 2)
  static constexpr uint32_t OFF_B = 2 * A_HALF;
  static constexpr int S_COL = 0;                          // two S buffers of kTN columns
  static constexpr int O_COL = 2 * kTN;                    // dOwn accumulator, D columns
  static constexpr int P_COL = 2 * kTN + D;                // two G buffers: hi kTN + lo kTN columns each
  static_assert(P_COL + 4 * kTN <= 512, "TMEM budget");
};

// Up to two independent problems of the same batch size per launch (blockIdx.z): MGCN / SMORE call
// InfoNCE for the item rows and for the user rows of every batch (mgcn.py:250-251, smore.py:406-407).
constexpr int kMaxProb = 2;
struct TcArgs {
  const float *own[kMaxProb], *other[kMaxProb], *ttl[kMaxProb], *coef[kMaxProb];
  float *out_a[kMaxProb], *out_b[kMaxProb];
};

template <int D, int MODE>
__global__ void __launch_bounds__(kThreadsI, 1)
infonce_tc_kernel(const __grid_constant__ TcArgs A, int batch, float inv_temp, int tiles_per_split) {
  using C = ICfg<D>;
  const float *__restrict__ Own = A.own[blockIdx.z], *__restrict__ Other = A.other[blockIdx.z],
                           *__restrict__ ttl = A.ttl[blockIdx.z], *__restrict__ coef = A.coef[blockIdx.z];
  float *__restrict__ out_a = A.out_a[blockIdx.z], *__restrict__ out_b = A.out_b[blockIdx.z];
  constexpr bool BWD = MODE != kFwd;
  constexpr uint32_t STAGE = 2 * C::B1_HALF + (BWD ? 2 * C::B2_HALF : 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_B + kStages * STAGE);
  uint64_t *full = bars, *empty = full + kStages, *s_full = empty + kStages, *s_empty = s_full + 2,
           *p_full = s_empty + 2, *p_empty = p_full + 2, *o_full = p_empty + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_full + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int own0 = blockIdx.x * kTM;
  const int n_other = (batch + kTN - 1) / kTN;
  const int t_begin = blockIdx.y * tiles_per_split, t_end = min(n_other, t_begin + tiles_per_split);
  const int n_t = max(0, t_end - t_begin);
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, kGroupThreads); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full + b, 1); mbar_init(s_empty + b, 256);
      mbar_init(p_full + b, 256); mbar_init(p_empty + b, 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);

  // ---- own tile -> hi / lo A operand of GEMM 1 (thread = row) ------------------------------
  const int o = own0 + (tid & (kTM - 1));         // epilogue warps w and w + 4 share the rows of a quadrant
  const bool live = tid < 2 * kTM && o < batch;
  if (tid < kTM) {
    const float *src = live ? Own + (size_t)o * D : nullptr;
#pragma unroll
    for (int c4 = 0; c4 < D / 4; ++c4) {
      float4 v = live ? ldg4(src + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
      split_tf32x4(v, hi, lo);
      const uint32_t off = (c4 / 8) * (kTM * 128) + sw128_off(tid, c4 % 8);
      *reinterpret_cast<float4 *>(smem + off) = hi;
      *reinterpret_cast<float4 *>(smem + C::A_HALF + off) = lo;
    }
    fence_proxy_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp >= kProdWarp0 && warp < kMmaWarp) {
    // =============================== producers ===============================================
    // two independent groups, group g owns tiles g, g + 2, ...: one tile of loads in flight per
    // group (GROUPS == STAGES, so a parity wait on `empty` cannot alias)
    const int grp = (tid - 32 * kProdWarp0) / kGroupThreads, ptid = (tid - 32 * kProdWarp0) % kGroupThreads;
    constexpr int VEC = kTN * (D / 4) / kGroupThreads;
    for (int t = grp; t < n_t; t += 2) {
      const int x0 = (t_begin + t) * kTN;
      float4 v[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int idx = ptid + kGroupThreads * i, row = idx / (D / 4), c4 = idx % (D / 4);
        v[i] = (x0 + row < batch) ? ldg4(Other + (size_t)(x0 + row) * D + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int s = t % kStages;
      mbar_wait(empty + s, ((t / kStages) & 1) ^ 1);
      uint8_t *stage = smem + C::OFF_B + s * STAGE;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int idx = ptid + kGroupThreads * i, row = idx / (D / 4), c4 = idx % (D / 4);
        float4 hi, lo;
        split_tf32x4(v[i], hi, lo);
        const uint32_t off1 = (c4 / 8) * (kTN * 128) + sw128_off(row, c4 % 8);
        *reinterpret_cast<float4 *>(stage + off1) = hi;
        *reinterpret_cast<float4 *>(stage + C::B1_HALF + off1) = lo;
        if constexpr (BWD) {
          const uint32_t off2 = mn32_off<D>(row, c4);
          *reinterpret_cast<float4 *>(stage + 2 * C::B1_HALF + off2) = hi;
          *reinterpret_cast<float4 *>(stage + 2 * C::B1_HALF + C::B2_HALF + off2) = lo;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(full + s);
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ==============================================
    constexpr uint32_t idesc1 = idesc_tf32(kTM, kTN, false, false);
    constexpr uint32_t idesc2 = idesc_tf32(kTM, D, false, true);
    constexpr uint32_t b2_sbo = (D / 32) * 512, b2_step = 2 * b2_sbo;
    const uint64_t a_hi = smem_desc_sw128(smem_base, 16, 1024);
    const uint64_t a_lo = smem_desc_sw128(smem_base + C::A_HALF, 16, 1024);
    for (int t = 0; t <= n_t; ++t) {
      if (t < n_t) {
        // ---- GEMM 1 of tile t: S[t & 1] = Own Other_t^T
        const int s = t % kStages, b = t & 1;
        mbar_wait(full + s, (t / kStages) & 1);
        mbar_wait(s_empty + b, ((t >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + C::S_COL + b * kTN;
        const uint64_t b_hi = smem_desc_sw128(smem_base + C::OFF_B + s * STAGE, 16, 1024);
        const uint64_t b_lo = b_hi + (C::B1_HALF >> 4);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t a0 = pass == 0 ? a_lo : a_hi;
          const uint64_t b0 = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int kb = 0; kb < C::KB; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = a0 + ((kb * (kTM * 128) + ks * 32) >> 4);
              const uint64_t bd = b0 + ((kb * (kTN * 128) + ks * 32) >> 4);
              if (elect_one()) umma_tf32_ss(d_tmem, ad, bd, idesc1, (pass | kb | ks) != 0);
            }
        }
        if (elect_one()) {
          umma_commit(s_full + b);
          if constexpr (!BWD) umma_commit(empty + s);
        }
        __syncwarp();
      }
      if constexpr (BWD) {
        if (t >= 1) {
          // ---- GEMM 2 of tile u = t - 1: dOwn += G_u Other_u  (A = G from TMEM, hi / lo)
          const int u = t - 1, s = u % kStages, pb = u & 1;
          mbar_wait(p_full + pb, (u >> 1) & 1);
          fence_after_sync();
          const uint32_t d_tmem = tmem_base + C::O_COL;
          const uint32_t g_hi = tmem_base + C::P_COL + pb * (2 * kTN), g_lo = g_hi + kTN;
          const uint32_t st = smem_base + C::OFF_B + s * STAGE + 2 * C::B1_HALF;
          const uint64_t v_hi = smem_desc_mn32(st, 512, b2_sbo);
          const uint64_t v_lo = smem_desc_mn32(st + C::B2_HALF, 512, b2_sbo);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a0 = pass == 0 ? g_lo : g_hi;
            const uint64_t b0 = pass == 1 ? v_lo : v_hi;
#pragma unroll
            for (int ks = 0; ks < kTN / 8; ++ks) {
              const uint64_t bd = b0 + ((ks * b2_step) >> 4);
              const uint32_t acc = (u | pass | ks) != 0;
              if (elect_one()) umma_tf32_ts(d_tmem, a0 + ks * 8, bd, idesc2, acc);
            }
          }
          if (elect_one()) {
            umma_commit(empty + s);          // the stage is free once both GEMMs have read it
            umma_commit(p_empty + pb);
            if (u == n_t - 1) umma_commit(o_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== epilogue ================================================
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int half = warp >> 2;                 // which 32 of a tile's 64 columns this warp takes
    float rowsum = 0.f, diag = 0.f;
    bool has_diag = false;
    float my_inv = 1.f, scale = 0.f;
    if constexpr (BWD) {
      scale = coef[0] / (float)batch * inv_temp;
      if (MODE == kBwdRow && live) my_inv = 1.f / ttl[o];
    }
    // exp(s / t) = ex2(s * log2(e) / t): |s / t| <= 1 / t, the approximate unit is good to ~1e-7 there
    const float it2 = inv_temp * 1.4426950408889634f;
    for (int t = 0; t < n_t; ++t) {
      const int b = t & 1, x0 = (t_begin + t) * kTN + half * 32;
      mbar_wait(s_full + b, (t >> 1) & 1);
      fence_after_sync();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + lane_base + C::S_COL + b * kTN + half * 32, r);
      tmem_ld_wait();
      fence_before_sync();
      mbar_arrive(s_empty + b);                // this warp's share of S[b] is in registers
      if constexpr (!BWD) {
        float part[4] = {0.f, 0.f, 0.f, 0.f};  // four chains instead of one 32-long dependent sum
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int x = x0 + c;
          const float sc = __uint_as_float(r[c]);
          if (x < batch) part[c & 3] += exp2f(sc * it2);
          if (x == o) { diag = sc; has_diag = true; }
        }
        rowsum += (part[0] + part[1]) + (part[2] + part[3]);
      } else {
        uint32_t hi[32], lo[32];
        // column pass: 1 / ttl of this warp's 32 columns, one per lane, handed round with shuffles
        float col_inv = 1.f;
        if (MODE == kBwdCol) {
          const int xl = x0 + (tid & 31);
          col_inv = xl < batch ? __frcp_rn(__ldg(ttl + xl)) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int x = x0 + c;
          const float inv_c = MODE == kBwdCol ? __shfl_sync(0xffffffffu, col_inv, c) : my_inv;
          float g = 0.f;
          if (live && x < batch) {
            const float inv = inv_c;
            g = exp2f(__uint_as_float(r[c]) * it2) * inv;
            if (x == o) g -= 1.f;
            g *= scale;
          }
          float h, l;
          split_tf32(g, h, l);
          hi[c] = __float_as_uint(h);
          lo[c] = __float_as_uint(l);
        }
        mbar_wait(p_empty + b, ((t >> 1) & 1) ^ 1);            // GEMM 2 of tile t - 2 has read G[b]
        fence_after_sync();
        const uint32_t g_hi = tmem_base + lane_base + C::P_COL + b * (2 * kTN) + half * 32;
        tmem_st_32x32b_x32(g_hi, hi);
        tmem_st_32x32b_x32(g_hi + kTN, lo);
        tmem_st_wait();
        fence_before_sync();
        mbar_arrive(p_full + b);
      }
    }
    if constexpr (!BWD) {
      // out_a = ttl_part[2 * split + half][row], out_b = pos[row] (written where the diagonal falls)
      if (live) {
        out_a[(size_t)(2 * blockIdx.y + half) * batch + o] = rowsum;
        if (has_diag) out_b[o] = diag;
      }
    } else {
      static_assert(D == 64, "the two warps of a quadrant store 32 columns of dOwn each");
      float *dst = out_a + ((size_t)blockIdx.y * batch + o) * D + half * 32;     // this split's slab
      if (n_t > 0) {
        mbar_wait(o_full, 0);
        fence_after_sync();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + lane_base + C::O_COL + half * 32, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            *reinterpret_cast<float4 *>(dst + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                                               __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        }
      } else if (live) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4 *>(dst + q) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// ttl[r] = sum over splits, loss = mean_r( log ttl_r - s_rr / t ); one CTA, fixed order.
struct FinishArgs {
  const float *ttl_part[kMaxProb], *pos[kMaxProb];
  float *ttl[kMaxProb], *loss_out[kMaxProb];
};
__global__ void __launch_bounds__(1024)
infonce_finish_kernel(const __grid_constant__ FinishArgs F, int n_splits, int batch, float inv_temp) {
  __shared__ float red[32];
  const float *__restrict__ ttl_part = F.ttl_part[blockIdx.x], *__restrict__ pos = F.pos[blockIdx.x];
  float *__restrict__ ttl = F.ttl[blockIdx.x], *__restrict__ loss_out = F.loss_out[blockIdx.x];
  float l = 0.f;
  for (int r = threadIdx.x; r < batch; r += 1024) {
    float s = 0.f;
    for (int sp = 0; sp < n_splits; ++sp) s += ttl_part[(size_t)sp * batch + r];
    ttl[r] = s;
    l += logf(s) - pos[r] * inv_temp;
  }
  l = block_sum<1024>(l, red);
  if (threadIdx.x == 0) loss_out[0] = l / (float)batch;
}

template <int MODE>
int launch_mode(const TcArgs &A, int n_prob, int batch, float inv_temp, int splits, int tps, cudaStream_t stream) {
  constexpr int D = 64;
  using C = ICfg<D>;
  constexpr uint32_t STAGE = 2 * C::B1_HALF + (MODE != kFwd ? 2 * C::B2_HALF : 0);
  const size_t smem = 1024 + C::OFF_B + kStages * STAGE + 16 * 8 + 16;
  auto kern = infonce_tc_kernel<D, MODE>;
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((batch + kTM - 1) / kTM, splits, n_prob);
  kern<<<grid, kThreadsI, smem, stream>>>(A, batch, inv_temp, tps);
  MMREC_CHECK_LAUNCH("infonce_tc_kernel");
  return MMREC_OK;
}

}  // namespace

bool infonce_tc_enabled(int d) {
  static int on = getenv("MMREC_INFONCE_TC") ? atoi(getenv("MMREC_INFONCE_TC")) : 1;
  return on != 0 && d == 64;
}

// "Other"-dimension splits of the tcgen05 path: one wave of 128-row CTAs, never more than `cap`
// (the split count the caller sized its workspaces for).
int infonce_tc_splits(int batch, int cap, int *tiles_per_split, int n_prob = 1) {
  const int n_rt = (batch + kTM - 1) / kTM, n_ot = (batch + kTN - 1) / kTN;
  int s = max(1, min(min(n_ot, cap / 2), kNumSMs / (n_rt * n_prob)));   // one wave; the forward writes two slabs per split
  const int tps = (n_ot + s - 1) / s;
  *tiles_per_split = tps;
  return (n_ot + tps - 1) / tps;
}

// forward: partial_p = pos[batch] | ttl_part[2 * splits][batch] for each of the n_prob problems
int infonce_fwd_tc(int n_prob, const float *const *V1n, const float *const *V2n, int batch, float inv_temp,
                   int cap_splits, float *const *partial, float *const *ttl, float *const *loss_out,
                   cudaStream_t stream) {
  int tps;
  const int splits = infonce_tc_splits(batch, cap_splits, &tps, n_prob);
  TcArgs A{};
  FinishArgs F{};
  for (int p = 0; p < n_prob; ++p) {
    A.own[p] = V1n[p]; A.other[p] = V2n[p];
    A.out_a[p] = partial[p] + batch;      // ttl_part
    A.out_b[p] = partial[p];              // pos
    F.ttl_part[p] = partial[p] + batch; F.pos[p] = partial[p]; F.ttl[p] = ttl[p]; F.loss_out[p] = loss_out[p];
  }
  const int rc = launch_mode<kFwd>(A, n_prob, batch, inv_temp, splits, tps, stream);
  if (rc != MMREC_OK) return rc;
  infonce_finish_kernel<<<n_prob, 1024, 0, stream>>>(F, 2 * splits, batch, inv_temp);
  MMREC_CHECK_LAUNCH("infonce_finish_kernel");
  return MMREC_OK;
}

// backward: dV1_p / dV2_p hold `*splits_out` slabs of [batch, 64] partial sums for the scatter kernel
int infonce_bwd_tc(int n_prob, const float *const *V1n, const float *const *V2n, const float *const *ttl, int batch,
                   float inv_temp, int cap_splits, const float *const *coef, float *const *dV1, float *const *dV2,
                   int *splits_out, cudaStream_t stream) {
  int tps;
  const int splits = infonce_tc_splits(batch, cap_splits, &tps, n_prob);
  *splits_out = splits;
  TcArgs R{}, Cc{};
  for (int p = 0; p < n_prob; ++p) {
    R.own[p] = V1n[p]; R.other[p] = V2n[p]; R.ttl[p] = ttl[p]; R.coef[p] = coef[p]; R.out_a[p] = dV1[p];
    Cc.own[p] = V2n[p]; Cc.other[p] = V1n[p]; Cc.ttl[p] = ttl[p]; Cc.coef[p] = coef[p]; Cc.out_a[p] = dV2[p];
  }
  int rc = launch_mode<kBwdRow>(R, n_prob, batch, inv_temp, splits, tps, stream);
  if (rc != MMREC_OK) return rc;
  return launch_mode<kBwdCol>(Cc, n_prob, batch, inv_temp, splits, tps, stream);
}

}  // namespace mmrec
