// Full-rank scoring on the 5th-generation tensor cores, fused with train-item masking and a
// per-user top-K (K8/K9/K10): the tcgen05 path behind mmrec_score_mask_topk_f32 for d = 32 / 64.
//
// One CTA owns 128 users (the M = 128 rows of the MMA = the 128 TMEM lanes) and walks its item
// range in tiles of 64 items. Warp roles:
//   warps 4-7  producers : stream item rows from HBM/L2 (coalesced float4, prefetched one tile
//                          ahead in registers), split every value into tf32 hi/lo and store both
//                          halves into 128-byte-swizzled K-major shared-memory tiles;
//   warp  8    MMA       : one elected thread issues tcgen05.mma kind::tf32, 3 x (d/8) per tile
//                          (lo*hi + hi*lo + hi*hi -> fp32-accurate scores), accumulating in one of
//                          four 64-column TMEM buffers; tcgen05.commit frees the smem stage and
//                          publishes the accumulator;
//   warps 0-3  epilogue  : tcgen05.ld -- TMEM lane i is user i, so every thread reads the scores
//                          of its own user, 32 items at a time, filters them against the running
//                          K-th best, applies the train-item mask (ascending cursor) and pushes the
//                          survivors into a K-entry heap in shared memory ordered by (score, -id).
// Scores never reach HBM. Stages hand over through mbarriers (smem full/empty, TMEM full/empty).
// Output: per item-split partial lists, descending score, ties -> lower id, merged by
// topk_merge_kernel (score_topk.cu). The user tile is split once per CTA (hi/lo A operand tiles).
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace mmrec {
namespace {

using namespace tc05;

constexpr int kTileM = 128;     // users per CTA
constexpr int kTileN = 64;      // items per MMA tile
constexpr int kAcc = 4;         // TMEM accumulator buffers (64 columns each)
constexpr int kThreadsTC = 288; // 4 epilogue + 4 producer + 1 MMA warp
constexpr int kChunk = 32;      // scores handled per tcgen05.ld

template <int D>
struct Cfg {
  static constexpr int KB = D / 32;                         // 128-byte K atoms per row
  static constexpr int STAGES = D <= 32 ? 4 : 2;
  static constexpr uint32_t A_HALF = KB * kTileM * 128;     // bytes of the hi (or lo) user tile
  static constexpr uint32_t B_HALF = KB * kTileN * 128;     // bytes of the hi (or lo) item tile
  static constexpr uint32_t STAGE = 2 * B_HALF;
  static constexpr uint32_t OFF_B = 2 * A_HALF;
  static constexpr uint32_t OFF_HEAP = OFF_B + STAGES * STAGE;
  static constexpr int VEC = kTileN * (D / 4) / 128;        // float4 per producer thread per tile
};

struct Better {   // strict total order on (score, id): a ranks before b
  __device__ static bool worse(float av, int ai, float bv, int bi) { return av < bv || (av == bv && ai > bi); }
};

// The K best of a user live in an 8-ary heap in shared memory, column `tid` of [k][128] arrays
// (bank = lane: conflict-free), root = the worst kept entry. K <= 73 gives depth 2: replacing the
// root is two rounds of eight independent shared-memory loads instead of the six dependent
// rounds of a binary heap -- the epilogue is a chain of dependent smem latencies, not throughput.
constexpr int kAry = 8;

// Place (sc, gid) at `pos` of the heap prefix [0, lim) and sift it down.
__device__ __forceinline__ void heap_sift_down(float *hv, int32_t *hi_, int pos, int lim, float sc, int gid) {
  for (;;) {
    const int c0 = kAry * pos + 1;
    if (c0 >= lim) break;
    float v[kAry];
    int id[kAry];
#pragma unroll
    for (int j = 0; j < kAry; ++j) {
      const bool in = c0 + j < lim;
      v[j] = in ? hv[(c0 + j) * kTileM] : CUDART_INF_F;      // +inf never wins "worst"
      id[j] = in ? hi_[(c0 + j) * kTileM] : 0;
    }
    // worst child = lowest score, ties -> highest id. Shallow trees instead of a 7-step scan: the
    // epilogue is one warp per scheduler, so the DEPTH of the dependent chain is what costs.
    const float wv = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), fminf(fminf(v[4], v[5]), fminf(v[6], v[7])));
    int t[kAry];
#pragma unroll
    for (int j = 0; j < kAry; ++j) t[j] = v[j] == wv ? id[j] : -1;
    const int wi = max(max(max(t[0], t[1]), max(t[2], t[3])), max(max(t[4], t[5]), max(t[6], t[7])));
    int w = 0;
#pragma unroll
    for (int j = 1; j < kAry; ++j) w = t[j] == wi ? j : w;
    w = t[0] == wi ? 0 : w;
    if (!Better::worse(wv, wi, sc, gid)) break;
    hv[pos * kTileM] = wv;
    hi_[pos * kTileM] = wi;
    pos = c0 + w;
  }
  hv[pos * kTileM] = sc;
  hi_[pos * kTileM] = gid;
}

// Append (sc, gid) at `pos` (= current size) and sift it up.
__device__ __forceinline__ void heap_sift_up(float *hv, int32_t *hi_, int pos, float sc, int gid) {
  while (pos > 0) {
    const int par = (pos - 1) / kAry;
    const float pv = hv[par * kTileM];
    const int pi = hi_[par * kTileM];
    if (!Better::worse(sc, gid, pv, pi)) break;
    hv[pos * kTileM] = pv;
    hi_[pos * kTileM] = pi;
    pos = par;
  }
  hv[pos * kTileM] = sc;
  hi_[pos * kTileM] = gid;
}

template <int D>
__global__ void __launch_bounds__(kThreadsTC, 1)
score_topk_tc_kernel(const float *__restrict__ user_emb, const int64_t *__restrict__ users, int n_users,
                     const float *__restrict__ item_emb, int n_items, int item_offset,
                     const int32_t *__restrict__ mask_rowptr, const int32_t *__restrict__ mask_cols, int k,
                     int items_per_split, float *__restrict__ ws_val, int32_t *__restrict__ ws_idx, int dbg) {
  using C = Cfg<D>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (SW128 atoms) by an offset on the __shared__ array itself: pointers derived
  // through an integer cast would lose their address space and compile to generic LD/ST
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float *h_val = reinterpret_cast<float *>(smem + C::OFF_HEAP);              // [k][128]
  int32_t *h_idx = reinterpret_cast<int32_t *>(h_val + (size_t)k * kTileM);  // [k][128]
  float *c_val = reinterpret_cast<float *>(h_idx + (size_t)k * kTileM);      // [32][128]
  uint8_t *c_col = reinterpret_cast<uint8_t *>(c_val + kChunk * kTileM);     // [32][128]
  uint64_t *bars = reinterpret_cast<uint64_t *>(c_col + kChunk * kTileM);
  uint64_t *full = bars, *empty = bars + C::STAGES, *tfull = bars + 2 * C::STAGES, *tempty = tfull + kAcc;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + kAcc);

  // warp index through a broadcast: provably warp-uniform, so role branches and the MMA issue
  // loop (descriptor arithmetic included) compile to the uniform datapath
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int j_begin = blockIdx.y * items_per_split;
  const int j_end = min(n_items, j_begin + items_per_split);
  const int n_tiles = (j_end - j_begin + kTileN - 1) / kTileN;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full + s, 128); mbar_init(empty + s, 1); }
    for (int b = 0; b < kAcc; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 128); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, kAcc * kTileN);

  // ---- user tile -> hi/lo A operand (thread = user row = TMEM lane) ------------------------
  const int b_user = blockIdx.x * kTileM + tid;
  const bool live = tid < kTileM && b_user < n_users;
  if (tid < kTileM) {
    const float *src = live ? user_emb + (size_t)users[b_user] * D : nullptr;
#pragma unroll
    for (int c4 = 0; c4 < D / 4; ++c4) {
      float4 v = live ? ldg4(src + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f), hi, lo;
      split_tf32x4(v, hi, lo);
      const uint32_t off = (c4 / 8) * (kTileM * 128) + sw128_off(tid, c4 % 8);
      *reinterpret_cast<float4 *>(smem + off) = hi;
      *reinterpret_cast<float4 *>(smem + C::A_HALF + off) = lo;
    }
    fence_proxy_async_smem();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp >= 4 && warp < 8) {
    // =============================== producers ===============================================
    const int ptid = tid - 128;
    float4 cur[C::VEC], nxt[C::VEC];
    auto load_tile = [&](float4 (&dst)[C::VEC], int t) {
      const int j0 = j_begin + t * kTileN;
#pragma unroll
      for (int i = 0; i < C::VEC; ++i) {
        const int idx = ptid + 128 * i, row = idx / (D / 4), c4 = idx % (D / 4);
        dst[i] = (j0 + row < j_end) ? ldg4(item_emb + (size_t)(j0 + row) * D + c4 * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_tile(cur, 0);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 1 < n_tiles) load_tile(nxt, t + 1);
      const int s = t % C::STAGES;
      mbar_wait(empty + s, ((t / C::STAGES) & 1) ^ 1);
      uint8_t *stage = smem + C::OFF_B + s * C::STAGE;
#pragma unroll
      for (int i = 0; i < C::VEC; ++i) {
        const int idx = ptid + 128 * i, row = idx / (D / 4), c4 = idx % (D / 4);
        float4 hi, lo;
        split_tf32x4(cur[i], hi, lo);
        const uint32_t off = (c4 / 8) * (kTileN * 128) + sw128_off(row, c4 % 8);
        *reinterpret_cast<float4 *>(stage + off) = hi;
        *reinterpret_cast<float4 *>(stage + C::B_HALF + off) = lo;
      }
      fence_proxy_async_smem();
      mbar_arrive(full + s);
#pragma unroll
      for (int i = 0; i < C::VEC; ++i) cur[i] = nxt[i];
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ==============================================
    constexpr uint32_t idesc = idesc_tf32(kTileM, kTileN, false, false);
    const uint64_t a_hi = smem_desc_sw128(smem_base, 16, 1024);
    const uint64_t a_lo = smem_desc_sw128(smem_base + C::A_HALF, 16, 1024);
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % C::STAGES, b = t % kAcc;
      mbar_wait(full + s, (t / C::STAGES) & 1);
      mbar_wait(tempty + b, ((t / kAcc) & 1) ^ 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + b * kTileN;
      const uint64_t b_hi = smem_desc_sw128(smem_base + C::OFF_B + s * C::STAGE, 16, 1024);
      const uint64_t b_lo = b_hi + (C::B_HALF >> 4);
      // small cross terms first, the dominant hi*hi product last; a K step inside the swizzle atom
      // (and the next atom) is a plain offset in the descriptor's 16-byte-unit address field
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint64_t a0 = pass == 0 ? a_lo : a_hi;
        const uint64_t b0 = pass == 1 ? b_lo : b_hi;
#pragma unroll
        for (int kb = 0; kb < C::KB; ++kb)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = a0 + ((kb * (kTileM * 128) + ks * 32) >> 4);
            const uint64_t bd = b0 + ((kb * (kTileN * 128) + ks * 32) >> 4);
            if (elect_one()) umma_tf32_ss(d_tmem, ad, bd, idesc, (pass | kb | ks) != 0);
          }
      }
      if (elect_one()) {
        umma_commit(empty + s);     // smem stage reusable once these MMAs have read it
        umma_commit(tfull + b);     // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue: mask + top-K ==================================
    int mp = 0, mend = 0, next_masked = INT_MAX;
    if (live && mask_rowptr != nullptr) {
      int lo = mask_rowptr[b_user];
      mend = mask_rowptr[b_user + 1];
      int hi = mend;
      const int target = item_offset + j_begin;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (mask_cols[mid] < target) lo = mid + 1; else hi = mid;
      }
      mp = lo;
      if (mp < mend) next_masked = mask_cols[mp];
    }
    int cnt = 0;
    float thr = -CUDART_INF_F;
    float *hv = h_val + tid;
    int32_t *hi_ = h_idx + tid;
    float *cv = c_val + tid;
    uint8_t *cc = c_col + tid;

    for (int t = 0; t < n_tiles; ++t) {
      const int b = t % kAcc;
      mbar_wait(tfull + b, (t / kAcc) & 1);
      fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + b * kTileN;
#pragma unroll
      for (int half = 0; half < kTileN / kChunk; ++half) {
        uint32_t r[kChunk];
        tmem_ld_32x32b_x32(taddr + half * kChunk, r);
        tmem_ld_wait();
        if (half == kTileN / kChunk - 1) {
          fence_before_sync();
          mbar_arrive(tempty + b);       // all columns of this buffer are in registers
        }
        const int j0 = j_begin + t * kTileN + half * kChunk;
        const int valid = min(kChunk, j_end - j0);
        // cheap reject: nothing in this chunk beats the current K-th best
        float m = -CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < kChunk; ++c) m = fmaxf(m, c < valid ? __uint_as_float(r[c]) : -CUDART_INF_F);
        if (dbg == 1) m = -CUDART_INF_F;
        if (live && valid > 0 && (cnt < k || m > thr) && dbg != 2) {
          int n = 0;
#pragma unroll
          for (int c = 0; c < kChunk; ++c) {
            const float sc = __uint_as_float(r[c]);
            if (c < valid && (cnt < k || sc > thr)) {
              cv[n * kTileM] = sc;
              cc[n * kTileM] = (uint8_t)c;
              ++n;
            }
          }
          for (int i = 0; i < n; ++i) {
            float sc = cv[i * kTileM];
            const int gid = item_offset + j0 + cc[i * kTileM];
            while (next_masked < gid) {
              ++mp;
              next_masked = mp < mend ? mask_cols[mp] : INT_MAX;
            }
            if (gid == next_masked) sc = -1e10f;                  // trainer.py:524
            if (cnt < k) {
              heap_sift_up(hv, hi_, cnt++, sc, gid);
              if (cnt == k) thr = hv[0];
            } else if (sc > thr) {
              heap_sift_down(hv, hi_, 0, k, sc, gid);             // replace the worst kept entry
              thr = hv[0];
            }
          }
        }
        __syncwarp();       // tcgen05.ld / mbarrier waits below are warp-collective
      }
    }
    // heap -> descending list (in-place heap sort: the worst entry moves to the end each round)
    if (live) {
      for (int size = cnt; size > 1; --size) {
        const float lv = hv[(size - 1) * kTileM];
        const int li = hi_[(size - 1) * kTileM];
        hv[(size - 1) * kTileM] = hv[0];
        hi_[(size - 1) * kTileM] = hi_[0];
        heap_sift_down(hv, hi_, 0, size - 1, lv, li);
      }
      float *ov = ws_val + ((size_t)blockIdx.y * n_users + b_user) * k;
      int32_t *oi = ws_idx + ((size_t)blockIdx.y * n_users + b_user) * k;
      for (int t = 0; t < k; ++t) {
        ov[t] = t < cnt ? hv[t * kTileM] : -CUDART_INF_F;        // index 0 = best after the sort
        oi[t] = t < cnt ? hi_[t * kTileM] : INT_MAX;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    fence_after_sync();
    tmem_dealloc(tmem_base, kAcc * kTileN);
  }
}

inline int dbg_mode() {
  static int m = getenv("MMREC_TOPK_DEBUG") ? atoi(getenv("MMREC_TOPK_DEBUG")) : 0;
  return m;
}

template <int D>
int launch_tc(const float *user_emb, const int64_t *users, int n_users, const float *item_emb, int n_items,
              int item_offset, const int32_t *mask_rowptr, const int32_t *mask_cols, int k, int n_splits,
              float *ws_val, int32_t *ws_idx, cudaStream_t stream) {
  using C = Cfg<D>;
  const size_t smem = 1024 + C::OFF_HEAP + (size_t)k * kTileM * 8 + kChunk * kTileM * 5 +
                      (2 * C::STAGES + 2 * kAcc) * 8 + 16;
  if (smem > 227 * 1024) return 1;          // K too large for this tiling: SIMT path
  MMREC_CUDA(cudaFuncSetAttribute(score_topk_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items_per_split = ((n_items + n_splits - 1) / n_splits + kTileN - 1) / kTileN * kTileN;
  dim3 grid((n_users + kTileM - 1) / kTileM, n_splits);
  score_topk_tc_kernel<D><<<grid, kThreadsTC, smem, stream>>>(user_emb, users, n_users, item_emb, n_items,
                                                              item_offset, mask_rowptr, mask_cols, k,
                                                              items_per_split, ws_val, ws_idx, dbg_mode());
  MMREC_CHECK_LAUNCH("score_topk_tc_kernel");
  return MMREC_OK;
}

}  // namespace

// Called by mmrec_score_mask_topk_f32 (score_topk.cu). Returns 1 if d is not covered here.
int score_topk_tc_dispatch(const float *user_emb, const int64_t *users, int n_users, const float *item_emb,
                           int n_items, int item_offset, int d, const int32_t *mask_rowptr,
                           const int32_t *mask_cols, int k, int n_splits, float *ws_val, int32_t *ws_idx,
                           cudaStream_t stream) {
  switch (d) {
    case 32: return launch_tc<32>(user_emb, users, n_users, item_emb, n_items, item_offset, mask_rowptr, mask_cols,
                                  k, n_splits, ws_val, ws_idx, stream);
    case 64: return launch_tc<64>(user_emb, users, n_users, item_emb, n_items, item_offset, mask_rowptr, mask_cols,
                                  k, n_splits, ws_val, ws_idx, stream);
    default: return 1;
  }
}

}  // namespace mmrec
