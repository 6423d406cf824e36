// SMORE's modality-aware preference module (smore.py:321-341), FORWARD, on the 5th-generation tensor
// cores for d = 64 -- the tcgen05 counterpart of side_fwd_kernel (side_net.cu), same contract:
//   hv = tanh(Wq1v f + b);  sv = softmax_d(Wq2v hv);  ht, st likewise;  gi/gt/gf = sigmoid(W c + b)
//   side = (m_i gi sv v + m_t gt st t + m_f gf f) / 3;   all = c + side;   saved = hv sv ht st gi gt gf
//
// A persistent CTA owns 128-row tiles (TMEM lane = row). Every A operand lives in TENSOR MEMORY as
// tf32 hi / lo column blocks (tcgen05.mma in its A-from-TMEM form), every B operand is a weight image
// that a tiny prologue kernel pre-splits into hi / lo K-major SWIZZLE_128B tiles (224 KB for the seven
// matrices) and that the MMA warp streams from L2 with 64 KB bulk copies through a two-stage ring:
//   S1  [q1v | q1t] = F [Wq1v ; Wq1t]^T                 one N = 128 product, A = F
//   E1  hv, ht = tanh(. + b)  -> written back to TMEM as the A operands of
//   S2  zv = Hv Wq2v^T,  zt = Ht Wq2t^T                 two N = 64 products
//   E2  sv, st = softmax rows;  a_v = sv v, a_t = st t
//   S3  [gi | gt] = C [Wgi ; Wgt]^T,  gf = C Wgf^T      N = 128 and N = 64, A = C
//   E3  sigmoid, dropout multipliers, side, all
// TMEM columns: [0,128) F -> Hv, [128,256) C, [256,384) accumulator Q, [384,512) Ht -> accumulator G.
//
// The round-1 tcgen05 forward lost to the mma.sync kernel because its row-per-thread epilogue was
// ~8 k instructions per thread with one warp per scheduler. Here SIXTEEN epilogue warps share a
// tile: warp w owns the rows of TMEM lane quadrant w % 4 and the 16 columns 16 (w / 4) .. of every
// [128, 64] operand (tcgen05.ld/st 32x32b.x16), activations are the SFU forms of common.cuh, the row
// softmax is reduced across the four column warps through shared memory. Every global tensor
// crosses a padded shared-memory staging tile so that loads and stores are whole 256-byte rows
// (a row-per-thread access touches 32 lines per instruction).
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace mmrec {
namespace {

using namespace tc05;

constexpr int kD = 64;
constexpr int kRows = 128;                 // rows per tile = UMMA M
constexpr int kEpiThreads = 512;           // 16 epilogue warps
constexpr int kThreadsS = kEpiThreads + 32;   // + the MMA / weight-loader warp
constexpr int kMmaWarpS = 16;
constexpr int kPitch = kD + 4;             // staging row pitch (floats)
constexpr uint32_t kStageBytes = 65536;    // weight ring stage
constexpr uint32_t kBufBytes = kRows * kPitch * 4;

// weight images in the workspace (bytes): hi tile | lo tile, K-major SWIZZLE_128B, two 32-float K atoms
constexpr uint32_t kImgQ1 = 0;             // N = 128: rows 0-63 query_v.0, 64-127 query_t.0
constexpr uint32_t kImgQ2v = 65536;        // N = 64
constexpr uint32_t kImgQ2t = 98304;        // N = 64
constexpr uint32_t kImgG12 = 131072;       // N = 128: gate_image_prefer.0, gate_text_prefer.0
constexpr uint32_t kImgGf = 196608;        // N = 64
constexpr uint32_t kImgTotal = 229376;

// TMEM columns
constexpr uint32_t kColF = 0, kColC = 128, kColQ = 256, kColG = 384;

struct SideTcWeights {
  const float *W[7];
  const float *b[7];
};

// ---- prologue: the seven [64, 64] weights -> five pre-split UMMA B images -------------------------
__global__ void __launch_bounds__(256)
side_w_images_kernel(SideTcWeights P, uint8_t *__restrict__ ws) {
  // chunk = 4 consecutive K values of one output row; 16 chunks per row; rows: 128 + 64 + 64 + 128 + 64 = 448
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= 448 * 16) return;
  const int grow = idx >> 4, c4 = idx & 15;
  int img_n, row;
  uint32_t base;
  const float *src;
  if (grow < 128) { img_n = 128; row = grow; base = kImgQ1; src = P.W[row < 64 ? 0 : 2] + (size_t)(row & 63) * kD; }
  else if (grow < 192) { img_n = 64; row = grow - 128; base = kImgQ2v; src = P.W[1] + (size_t)row * kD; }
  else if (grow < 256) { img_n = 64; row = grow - 192; base = kImgQ2t; src = P.W[3] + (size_t)row * kD; }
  else if (grow < 384) { img_n = 128; row = grow - 256; base = kImgG12; src = P.W[row < 64 ? 4 : 5] + (size_t)(row & 63) * kD; }
  else { img_n = 64; row = grow - 384; base = kImgGf; src = P.W[6] + (size_t)row * kD; }
  float4 hi, lo;
  split_tf32x4(ldg4(src + c4 * 4), hi, lo);
  const uint32_t off = (uint32_t)(c4 >> 3) * (uint32_t)(img_n * 128) + sw128_off(row, c4 & 7);
  *reinterpret_cast<float4 *>(ws + base + off) = hi;
  *reinterpret_cast<float4 *>(ws + base + (uint32_t)img_n * 256u + off) = lo;
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// D[tmem d_col .. + N) (=) A[tmem a_hi / a_lo, 64 columns each] x B^T, B image at smem address b (hi | lo, N rows)
template <int N>
__device__ __forceinline__ void issue_product(uint32_t tmem_base, uint32_t d_col, uint32_t a_col, uint32_t b_addr) {
  constexpr uint32_t idesc = idesc_tf32(kRows, N, false, false);
  const uint64_t b_hi = smem_desc_sw128(b_addr, 16, 1024);
  const uint64_t b_lo = smem_desc_sw128(b_addr + N * 256, 16, 1024);
  const uint32_t a_hi = tmem_base + a_col, a_lo = a_hi + kD;
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a0 = pass == 0 ? a_lo : a_hi;
    const uint64_t b0 = pass == 1 ? b_lo : b_hi;
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bd = b0 + ((kb * (N * 128) + ks * 32) >> 4);
        if (elect_one()) umma_tf32_ts(tmem_base + d_col, a0 + kb * 32 + ks * 8, bd, idesc, (pass | kb | ks) != 0);
      }
  }
}

struct SideTcArgs {
  const float *F, *V, *T, *C;
  const float *b[7];
  DropSpec drop;
  float *saved, *side, *all;
  const uint8_t *ws;
  int n, n_tiles;
};

__global__ void __launch_bounds__(kThreadsS, 1)
side_fwd_tc_kernel(const __grid_constant__ SideTcArgs A) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *stage0 = smem, *stage1 = smem + kStageBytes;
  float *buf0 = reinterpret_cast<float *>(smem + 2 * kStageBytes);
  float *buf1 = buf0 + kRows * kPitch;
  float *red = buf1 + kRows * kPitch;                          // [4][4][128] softmax exchanges
  float *sbias = red + 4 * 4 * kRows;                          // [5][64]: b of q1v, q1t, gi, gt, gf (0 where absent)
  uint64_t *bars = reinterpret_cast<uint64_t *>(sbias + 5 * kD);
  uint64_t *wfull = bars, *wempty = bars + 2, *a_ready = bars + 4, *s1_done = bars + 5, *e1_done = bars + 6,
           *s2_done = bars + 7, *e2_done = bars + 8, *s3_done = bars + 9;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    mbar_init(wfull + 0, 1); mbar_init(wfull + 1, 1); mbar_init(wempty + 0, 1); mbar_init(wempty + 1, 1);
    mbar_init(a_ready, kEpiThreads); mbar_init(s1_done, 1); mbar_init(e1_done, kEpiThreads);
    mbar_init(s2_done, 1); mbar_init(e2_done, kEpiThreads); mbar_init(s3_done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarpS) tmem_alloc(tmem_slot, 512);
  if (tid < 5 * kD) {                                          // a global round trip per bias use otherwise
    const int k = tid / kD;
    const float *b = A.b[k == 0 ? 0 : k == 1 ? 2 : k + 2];
    sbias[tid] = b != nullptr ? b[tid % kD] : 0.f;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int n = A.n;

  if (warp == kMmaWarpS) {
    // =============================== MMA issuer + weight loader ==================================
    const uint32_t s0 = smem_u32(stage0), s1 = smem_u32(stage1);
    int it = 0;
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t p = it & 1;
      // use 2 it of both stages: [q1v | q1t] -> stage 0, q2v + q2t -> stage 1
      mbar_wait(wempty + 0, 1);                       // second use of the previous tile has been consumed
      mbar_wait(wempty + 1, 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(wfull + 0, 65536);
        bulk_load(stage0, A.ws + kImgQ1, 65536, wfull + 0);
        mbar_arrive_expect_tx(wfull + 1, 65536);
        bulk_load(stage1, A.ws + kImgQ2v, 65536, wfull + 1);
      }
      __syncwarp();
      mbar_wait(a_ready, p);
      mbar_wait(wfull + 0, 0);
      fence_after_sync();
      issue_product<128>(tmem_base, kColQ, kColF, s0);                       // S1
      if (elect_one()) { umma_commit(wempty + 0); umma_commit(s1_done); }
      __syncwarp();
      // use 2 it + 1 of stage 0: [gi | gt], once S1 has read it
      mbar_wait(wempty + 0, 0);
      if (lane == 0) {
        mbar_arrive_expect_tx(wfull + 0, 65536);
        bulk_load(stage0, A.ws + kImgG12, 65536, wfull + 0);
      }
      __syncwarp();
      mbar_wait(e1_done, p);
      mbar_wait(wfull + 1, 0);
      fence_after_sync();
      issue_product<64>(tmem_base, kColQ, kColF, s1);                        // S2: zv = Hv Wq2v^T
      issue_product<64>(tmem_base, kColQ + 64, kColG, s1 + 32768);           //     zt = Ht Wq2t^T
      if (elect_one()) { umma_commit(wempty + 1); umma_commit(s2_done); }
      __syncwarp();
      mbar_wait(wempty + 1, 0);
      if (lane == 0) {
        mbar_arrive_expect_tx(wfull + 1, 32768);
        bulk_load(stage1, A.ws + kImgGf, 32768, wfull + 1);
      }
      __syncwarp();
      mbar_wait(e2_done, p);
      mbar_wait(wfull + 0, 1);
      mbar_wait(wfull + 1, 1);
      fence_after_sync();
      issue_product<128>(tmem_base, kColQ, kColC, s0);                       // S3: [gi | gt]
      issue_product<64>(tmem_base, kColG, kColC, s1);                        //     gf
      if (elect_one()) { umma_commit(wempty + 0); umma_commit(wempty + 1); umma_commit(s3_done); }
      __syncwarp();
    }
  } else {
    // =============================== epilogue (16 warps) ==========================================
    const int q = warp & 3, cg = warp >> 2;              // TMEM lane quadrant, column group
    const int r = q * 32 + lane, c0 = cg * 16;           // own row of the tile, first own column
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint64_t drop_key = A.drop.p > 0.f ? drop_stream(A.drop) : 0ull;
    const size_t nd = (size_t)n * kD;

    // coalesced global -> staging of one [128, 64] tile (rows past n: zeros). (Hoisting these loads above the
    // previous phase's write-out -- two-phase staging with the values parked in registers -- was measured:
    // 45.7 vs 45.3 us, the 32 extra live registers spill at the 96-register cap of a 17-warp CTA.)
    auto stage_in = [&](float *buf, const float *__restrict__ src, int row0) {
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + kEpiThreads * i, rr = idx >> 4, c4 = idx & 15;
        v[i] = row0 + rr < n ? ldg4(src + (size_t)(row0 + rr) * kD + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + kEpiThreads * i, rr = idx >> 4, c4 = idx & 15;
        *reinterpret_cast<float4 *>(buf + rr * kPitch + c4 * 4) = v[i];
      }
    };
    auto own_load = [&](const float *buf, float (&x)[16]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = *reinterpret_cast<const float4 *>(buf + r * kPitch + c0 + 4 * j);
        x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
      }
    };
    auto own_store = [&](float *buf, const float (&x)[16]) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4 *>(buf + r * kPitch + c0 + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    };
    // staging -> global, whole rows
    auto stage_out = [&](const float *buf, float *__restrict__ dst, int row0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + kEpiThreads * i, rr = idx >> 4, c4 = idx & 15;
        if (row0 + rr < n)
          *reinterpret_cast<float4 *>(dst + (size_t)(row0 + rr) * kD + c4 * 4) =
              *reinterpret_cast<const float4 *>(buf + rr * kPitch + c4 * 4);
      }
    };
    // own 16 values -> tf32 hi / lo column blocks of an A operand in TMEM
    auto to_tmem = [&](uint32_t col, const float (&x)[16]) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float h, l;
        split_tf32(x[j], h, l);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(l);
      }
      tmem_st_32x32b_x16(lane_addr + col + c0, hi);
      tmem_st_32x32b_x16(lane_addr + col + kD + c0, lo);
    };
    auto from_tmem = [&](uint32_t col, float (&x)[16]) {
      uint32_t t[16];
      tmem_ld_32x32b_x16(lane_addr + col + c0, t);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(t[j]);
    };
    auto add_bias = [&](float (&x)[16], int k) {               // k: row of sbias
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = *reinterpret_cast<const float4 *>(sbias + k * kD + c0 + 4 * j);
        x[4 * j] += t.x; x[4 * j + 1] += t.y; x[4 * j + 2] += t.z; x[4 * j + 3] += t.w;
      }
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t p = it & 1;
      const int row0 = tile * kRows, grow = row0 + r;
      float x[16], y[16], f[16];
      // ---------------- E0: F, C -> A operands in TMEM
      stage_in(buf0, A.F, row0);
      stage_in(buf1, A.C, row0);
      epi_bar();
      own_load(buf0, f);                         // kept in registers for E3 (Hv takes its TMEM columns)
      own_load(buf1, y);
      to_tmem(kColF, f);
      to_tmem(kColC, y);
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(a_ready);
      // ---------------- E1: hv, ht = tanh(q1 + b) -> A operands of S2
      mbar_wait(s1_done, p);                     // (every thread has left buf0 / buf1: S1 needed all 512 arrivals)
      fence_after_sync();
      from_tmem(kColQ, x);
      from_tmem(kColQ + 64, y);
      add_bias(x, 0);
      add_bias(y, 1);
#pragma unroll
      for (int j = 0; j < 16; ++j) { x[j] = fast_tanh(x[j]); y[j] = fast_tanh(y[j]); }
      to_tmem(kColF, x);                         // Hv over F (S1 is complete)
      to_tmem(kColG, y);                         // Ht
      tmem_st_wait();
      fence_before_sync();
      mbar_arrive(e1_done);
      if (A.saved != nullptr) {
        own_store(buf0, x);
        own_store(buf1, y);
        epi_bar();
        stage_out(buf0, A.saved + 0 * nd, row0);
        stage_out(buf1, A.saved + 2 * nd, row0);
        epi_bar();                               // both buffers are written again below
      }
      // ---------------- E2: sv, st = softmax rows of zv, zt;  a_v = sv v, a_t = st t
      stage_in(buf0, A.V, row0);
      stage_in(buf1, A.T, row0);
      mbar_wait(s2_done, p);
      fence_after_sync();
      from_tmem(kColQ, x);
      from_tmem(kColQ + 64, y);
      fence_before_sync();
      mbar_arrive(e2_done);                      // Q and the Ht columns are free for S3
      {
        float mx = x[0], my = y[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) { mx = fmaxf(mx, x[j]); my = fmaxf(my, y[j]); }
        red[(0 * 4 + cg) * kRows + r] = mx;
        red[(1 * 4 + cg) * kRows + r] = my;
        epi_bar();                               // also: V, T tiles are staged
        mx = fmaxf(fmaxf(red[(0 * 4 + 0) * kRows + r], red[(0 * 4 + 1) * kRows + r]),
                   fmaxf(red[(0 * 4 + 2) * kRows + r], red[(0 * 4 + 3) * kRows + r]));
        my = fmaxf(fmaxf(red[(1 * 4 + 0) * kRows + r], red[(1 * 4 + 1) * kRows + r]),
                   fmaxf(red[(1 * 4 + 2) * kRows + r], red[(1 * 4 + 3) * kRows + r]));
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          x[j] = fast_exp(x[j] - mx); sx += x[j];
          y[j] = fast_exp(y[j] - my); sy += y[j];
        }
        red[(2 * 4 + cg) * kRows + r] = sx;
        red[(3 * 4 + cg) * kRows + r] = sy;
        epi_bar();
        sx = (red[(2 * 4 + 0) * kRows + r] + red[(2 * 4 + 1) * kRows + r]) +
             (red[(2 * 4 + 2) * kRows + r] + red[(2 * 4 + 3) * kRows + r]);
        sy = (red[(3 * 4 + 0) * kRows + r] + red[(3 * 4 + 1) * kRows + r]) +
             (red[(3 * 4 + 2) * kRows + r] + red[(3 * 4 + 3) * kRows + r]);
        const float ix = fast_rcp(sx), iy = fast_rcp(sy);
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[j] *= ix; y[j] *= iy; }
      }
      float av[16], at[16];
      own_load(buf0, av);                        // v
      own_load(buf1, at);                        // t
#pragma unroll
      for (int j = 0; j < 16; ++j) { av[j] *= x[j]; at[j] *= y[j]; }
      epi_bar();                                 // every thread has read V / T: the buffers can be reused
      if (A.saved != nullptr) {
        own_store(buf0, x);
        own_store(buf1, y);
        epi_bar();
        stage_out(buf0, A.saved + 1 * nd, row0);
        stage_out(buf1, A.saved + 3 * nd, row0);
        epi_bar();
      }
      // ---------------- E3: gates, dropout, side, all
      mbar_wait(s3_done, p);
      fence_after_sync();
      float sd[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) sd[j] = 0.f;
      float *const sv_out[3] = {A.saved != nullptr ? A.saved + 4 * nd : nullptr, A.saved != nullptr ? A.saved + 5 * nd : nullptr,
                                A.saved != nullptr ? A.saved + 6 * nd : nullptr};
      uint64_t drop_base[3] = {0, 0, 0};         // first float4 index of this row in each mask plane
      if (A.drop.p > 0.f && grow < n) {
#pragma unroll
        for (int g = 0; g < 3; ++g) drop_base[g] = drop_row4(A.drop, g, grow, n, kD / 4);
      }
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        from_tmem(g == 0 ? kColQ : g == 1 ? kColQ + 64 : kColG, x);
        add_bias(x, 2 + g);
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fast_sigmoid(x[j]);
        float *buf = (g & 1) ? buf1 : buf0;
        if (A.saved != nullptr) own_store(buf, x);
        if (A.drop.p > 0.f) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 m = drop_mask4(drop_key, drop_base[g] + (c0 >> 2) + j, A.drop.p);
            x[4 * j] *= m.x; x[4 * j + 1] *= m.y; x[4 * j + 2] *= m.z; x[4 * j + 3] *= m.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) sd[j] = fmaf(x[j], g == 0 ? av[j] : g == 1 ? at[j] : f[j], sd[j]);
        if (A.saved != nullptr) {
          epi_bar();
          stage_out(buf, sv_out[g], row0);       // (the next gate stages into the other buffer)
        }
      }
      // c comes back from its A-operand columns: hi + lo is the fp32 value exactly (lo = c - hi is exact and
      // the columns keep all 32 bits; the tensor core is what ignores the low ones)
      from_tmem(kColC, y);
      from_tmem(kColC + kD, x);
      epi_bar();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        sd[j] *= (1.f / 3.f);
        y[j] = (y[j] + x[j]) + sd[j];
      }
      own_store(buf0, sd);
      own_store(buf1, y);
      epi_bar();
      stage_out(buf0, A.side, row0);
      stage_out(buf1, A.all, row0);
      epi_bar();                                 // the next tile stages into both buffers
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarpS) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// Which forward runs (read per call: the tests compare both inside one process): the tcgen05 kernel for
// d = 64 unless MMREC_SIDE_TC=0. Measured inside the captured SMORE / Baby step (profiles/r02_side_tc.txt):
// 50 + 2 us (kernel + weight images) against 61 us for the mma.sync kernel, step 2.374 -> 2.325 ms; Sports
// (54 k rows): 89 against 125 us. (Timed as an eager op the order flips at Baby size -- 77 vs 69 us -- because
// the two launches and the workspace allocation cost ~10 us of host time that a graph replay does not pay.)
inline bool side_tc_pick(int n, int d) {
  (void)n;
  if (d != kD) return false;
  const char *e = getenv("MMREC_SIDE_TC");
  return !(e && atoi(e) == 0);
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" size_t mmrec_smore_side_fwd_tc_workspace_bytes(int32_t n, int32_t d) {
  return side_tc_pick(n, d) ? (size_t)kImgTotal : 0;
}

extern "C" int mmrec_smore_side_fwd_tc_f32(const float *F, const float *V, const float *T, const float *C_,
                                           const float *const *W_host, const float *const *b_host,
                                           const MmrecDropout *drop, float *saved, float *side, float *all,
                                           int32_t n, int32_t d, void *ws, void *stream) {
  MMREC_REQUIRE(F && V && T && C_ && W_host && b_host && side && all && ws, MMREC_E_BADARG, "smore_side_fwd_tc: null pointer");
  MMREC_REQUIRE(d == kD, MMREC_E_BADARG, "smore_side_fwd_tc: d must be 64 (got %d)", d);
  MMREC_REQUIRE(n >= 0, MMREC_E_BADARG, "smore_side_fwd_tc: bad n");
  MMREC_REQUIRE(aligned16(F) && aligned16(V) && aligned16(T) && aligned16(C_) && aligned16(saved) && aligned16(side) &&
                    aligned16(all) && (reinterpret_cast<uintptr_t>(ws) & 1023u) == 0, MMREC_E_ALIGN,
                "smore_side_fwd_tc: operands must be 16-byte aligned, the workspace 1024-byte aligned");
  SideTcWeights P;
  SideTcArgs A{};
  for (int i = 0; i < 7; ++i) {
    MMREC_REQUIRE(W_host[i] != nullptr && aligned16(W_host[i]) && aligned16(b_host[i]), MMREC_E_BADARG,
                  "smore_side_fwd_tc: weight %d is null or misaligned", i);
    P.W[i] = W_host[i];
    P.b[i] = b_host[i];
    A.b[i] = b_host[i];
  }
  A.drop = DropSpec{nullptr, 0ull, 0.f, nullptr, 0};
  if (drop != nullptr) {
    MMREC_REQUIRE(drop->p >= 0.f && drop->p < 1.f, MMREC_E_BADARG, "dropout: p must be in [0, 1) (got %g)", (double)drop->p);
    MMREC_REQUIRE(drop->row_ids == nullptr || drop->n_total > 0, MMREC_E_BADARG, "dropout: row_ids needs n_total");
    A.drop = DropSpec{drop->counter, drop->seed, drop->p, reinterpret_cast<const long long *>(drop->row_ids), drop->n_total};
  }
  if (n == 0) return MMREC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  side_w_images_kernel<<<(448 * 16 + 255) / 256, 256, 0, st>>>(P, static_cast<uint8_t *>(ws));
  MMREC_CHECK_LAUNCH("side_w_images_kernel");
  A.F = F; A.V = V; A.T = T; A.C = C_; A.saved = saved; A.side = side; A.all = all;
  A.ws = static_cast<const uint8_t *>(ws);
  A.n = n;
  A.n_tiles = (n + kRows - 1) / kRows;
  const size_t smem = 1024 + 2 * (size_t)kStageBytes + 2 * (size_t)kBufBytes + 4 * 4 * kRows * 4 + 5 * kD * 4 + 10 * 8 + 16;
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(side_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  side_fwd_tc_kernel<<<min(A.n_tiles, kNumSMs), kThreadsS, smem, st>>>(A);
  MMREC_CHECK_LAUNCH("side_fwd_tc_kernel");
  return MMREC_OK;
}
