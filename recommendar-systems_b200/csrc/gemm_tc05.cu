// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 split: the
// modality projections of SMORE / MGCN / FREEDOM and their two backward GEMMs (K4), i.e. the only
// dense contractions of a training step that touch the 115 MB feature tables:
//   forward  y  = x W^T      x [I, F] (F = 4096 / 384), W [d, F]      A K-major,  B K-major, split-K
//   dW       dW = dy^T x     reduction over the I items               A MN-major, B MN-major, split-K,
//                                                                     computed as dW^T and stored transposed
//   dx       dx = dy W       output [I, F]                            A K-major,  B MN-major, N ranges
// One CTA owns 128 rows of the UMMA M dimension. Warp roles:
//   warps 4-11 producers : four independent groups stream 32-wide K blocks of both operands from
//                          HBM/L2 (coalesced float4, one block in flight per group), split every value into tf32
//                          hi/lo and store both halves into swizzled canonical UMMA tiles (K-major
//                          SWIZZLE_128B or MN-major SWIZZLE_128B_BASE32B, so neither backward
//                          GEMM transposes the table);
//   warp  12   MMA       : tcgen05.mma kind::tf32, lo*hi + hi*lo + hi*hi per K step, fp32
//                          accumulation in one of two TMEM buffers; tcgen05.commit hands the smem
//                          stage back to the producers and the accumulator to the epilogue;
//   warps 0-3  epilogue  : tcgen05.ld (TMEM lane = output row), bias, plain or transposed store.
// The kernels are HBM-bound by construction: per 16 KB of table streamed an SM issues 12 MMAs
// (384 tensor cycles) and ~200 producer issue slots against ~700 cycles of its HBM share.
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "tma_host.h"

namespace mmrec {
namespace {

using namespace tc05;

constexpr int kBM = 128;            // UMMA M
constexpr int kKB = 32;             // floats of K per pipeline stage (one 128-byte swizzle atom)
constexpr int kChunkKB = 2;         // K blocks per tensor-core accumulation chain
// Producers work in kGroups independent groups of kGroupThreads; group g owns K blocks g, g +
// kGroups, ... and keeps exactly ONE block of loads in flight (load -> wait -> split -> store).
// Keeping several blocks in flight per THREAD (the first version: 4 per thread) does not work:
// the compiler maps the loads of different blocks onto the same few scoreboards, so waiting for
// the oldest block also waits for the one issued just before -- the kernel ran at one block per
// DRAM round trip (20 GB/s per SM, 28 % of HBM peak). Independent warps cannot alias.
// 256 producer threads: four groups of 64 (two of 128 for the 128-wide N tile, whose block would
// not fit the registers of 64 threads).
constexpr int kProdThreads = 256;
constexpr int kMmaWarp = 4 + kProdThreads / 32;
constexpr int kThreadsG = 128 + kProdThreads + 32;

__device__ __forceinline__ void group_sync(int id, int n_threads) {     // named barrier of one producer group
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

struct GemmArgs {
  const float *A, *B, *bias;
  float *C;
  int M, N, K;                      // UMMA-space problem: C[M, N] = A[M, K] B[N, K]^T
  int lda, ldb, ldc;
  int kb_per_split;                 // K blocks per blockIdx.y
  int nt_per_cta;                   // N tiles per blockIdx.z
  size_t slab;                      // floats between the output slabs of consecutive K splits
  // EPI_ADAM: the product tile is a gradient that is never stored -- it is applied to the parameter
  // tile it belongs to (C = parameters, in place) with the Adam moments beside it
  float *exp_avg, *exp_avg_sq;
  const double *hyper;              // device: [lr, update count]
  double beta1, beta2;
  float eps, weight_decay, grad_scale;
  double *sumsq_partial;            // optional: per-CTA sum of the updated parameters' squares
  int debug;                        // MMREC_TA_DEBUG (timing experiments only): 1 = no arithmetic, 2 = no copies
  int act;                          // EPI_STORE, no split-K: 0 = none, 1 = tanh, 2 = sigmoid after the bias
  int flat;                         // table epilogues (no split-K): the (row tile, N tile) space is cut into gridDim.x
                                    // equal contiguous shares instead of rectangles -- 56 x 128 tiles on 148 SMs are
                                    // 48.4 tiles per CTA in ONE wave, where the best rectangle (5 column ranges) was
                                    // 280 CTAs = two waves of 26 tiles, each paying its own pipeline fill
  int mt_per_cta;                   // consecutive 128-row tiles walked by one CTA (0 / 1: one). With more row tiles
                                    // than SMs a CTA keeps its pipeline full across tiles instead of paying the
                                    // fill (TMEM allocation, first DRAM round trips, drain) once per 64 KB of rows
};

// Byte offset of element chunk inside one operand tile (extent E along M/N, 32 along K).
// K-major: 16-byte chunk c (4 floats of K) of row r.  MN-major: chunk q (4 floats of M/N) of K row k.
template <int E>
__device__ __forceinline__ uint32_t off_kmajor(int r, int c) {
  return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}
template <int E>
__device__ __forceinline__ uint32_t off_mnmajor(int k, int q) {
  // SWIZZLE_128B_BASE32B: atoms of 4 K rows x 128 B; [K group][M/N atom] order; 32-byte chunks swizzled
  return (uint32_t)((k >> 2) * (E / 32) + (q >> 3)) * 512u + (uint32_t)(k & 3) * 128u +
         (uint32_t)(((((q & 7) >> 1) ^ (k & 3)) << 5) | ((q & 1) << 4));
}

// One operand block [E x 32] in flight: PER float4 per producer thread.
template <int E, bool MN, int kGroupThreads>
struct Block {
  static constexpr int PER = E * 8 / kGroupThreads;
  static_assert(E * 8 % kGroupThreads == 0, "operand block must be a multiple of the producer group");
  float4 v[PER];
  // src element (i, k): K-major src[i * ld + k]; MN-major src[k * ld + i]. i < lim_i, k < lim_k else 0.
  __device__ __forceinline__ void load(const float *__restrict__ src, int ld, int i0, int lim_i, int k0, int lim_k,
                                       int ptid) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int idx = ptid + j * kGroupThreads;
      int i, k;
      if constexpr (!MN) { i = i0 + idx / 8; k = k0 + (idx % 8) * 4; }
      else { k = k0 + idx / (E / 4); i = i0 + (idx % (E / 4)) * 4; }
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < lim_i && k < lim_k) v[j] = ldg4(MN ? src + (size_t)k * ld + i : src + (size_t)i * ld + k);
    }
  }
  __device__ __forceinline__ void store(uint8_t *hi_tile, uint8_t *lo_tile, int ptid) const {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int idx = ptid + j * kGroupThreads;
      const uint32_t off = MN ? off_mnmajor<E>(idx / (E / 4), idx % (E / 4)) : off_kmajor<E>(idx / 8, idx % 8);
      float4 hi, lo;
      split_tf32x4(v[j], hi, lo);
      *reinterpret_cast<float4 *>(hi_tile + off) = hi;
      *reinterpret_cast<float4 *>(lo_tile + off) = lo;
    }
  }
};

enum { EPI_STORE = 0, EPI_ADAM = 1, EPI_SUMSQ = 2 };

template <int NT, int EPI_MODE = EPI_STORE>
struct GCfg {
  static constexpr uint32_t A_HALF = kBM * 128, B_HALF = NT * 128;
  static constexpr uint32_t STAGE = 2 * (A_HALF + B_HALF);
  // EPI_ADAM keeps the parameter / moment tiles of every epilogue warp in shared memory (3 x the
  // staging of a plain store): two operand stages (one K = 64 tile) are all that is left, and all
  // that is needed -- the epilogue (HBM streaming) is several times longer than a tile's products
  static constexpr int STAGES = EPI_MODE == EPI_ADAM ? 2 : NT <= 32 ? 4 : NT <= 64 ? 3 : 2;
  // epilogue staging: 4 warps x [32][NT + 4] floats (x 3 arrays for EPI_ADAM)
  static constexpr uint32_t EPI = EPI_MODE == EPI_ADAM ? 1024 + 4 * 3 * 32 * 64 * 4 : 4 * 32 * (NT + 4) * 4;
  static constexpr int EXTRA_BARS = EPI_MODE == EPI_ADAM ? 8 : 0;     // tile-landed mbarriers: epilogue warp x buffer
  static constexpr int TMEM_COLS = 2 * NT < 32 ? 32 : 2 * NT;
  static constexpr int GROUPS = NT >= 128 ? 2 : 4;          // independent producer groups
  static constexpr int GROUP_THREADS = kProdThreads / GROUPS;
};

template <bool A_MN, bool B_MN, int NT, bool TRANS_OUT, int EPI = EPI_STORE>
__global__ void __launch_bounds__(kThreadsG, 1)
gemm_tc05_kernel(const GemmArgs g, const __grid_constant__ TmaMaps3 maps) {
  using C = GCfg<NT, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::STAGES * C::STAGE);
  uint64_t *full = bars, *empty = bars + C::STAGES, *tfull = empty + C::STAGES, *tempty = tfull + 2;
  uint64_t *pbar = tempty + 2;                               // EPI_ADAM: tile-landed barrier per epilogue warp
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pbar + C::EXTRA_BARS);
  volatile int *passed = reinterpret_cast<volatile int *>(tmem_slot + 1);   // producer wait chain (see below)
  float *stg = reinterpret_cast<float *>(tmem_slot + 4);     // 4 warps x [32][NT + 4] epilogue staging

  // warp index through a broadcast: provably warp-uniform, so role branches and the MMA issue
  // loop (descriptor arithmetic included) compile to the uniform datapath
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int mt_per = g.mt_per_cta > 1 ? g.mt_per_cta : 1;
  const int m_tiles_total = (g.M + kBM - 1) / kBM;
  const int n_mt = max(0, min(mt_per, m_tiles_total - (int)blockIdx.x * mt_per));
  const int m0 = blockIdx.x * mt_per * kBM;        // first row tile of this CTA; tile j starts at m0 + (j / n_nt) * kBM
  const int n_kb_total = (g.K + kKB - 1) / kKB;
  const int kb_begin = blockIdx.y * g.kb_per_split, kb_end = min(n_kb_total, kb_begin + g.kb_per_split);
  const int n_nt_total = (g.N + NT - 1) / NT;
  const int nt_begin = blockIdx.z * g.nt_per_cta, nt_end = min(n_nt_total, nt_begin + g.nt_per_cta);
  const int n_kb = max(0, kb_end - kb_begin), n_nt = max(0, nt_end - nt_begin);
  // tile j of this CTA -> (N tile, first row): rectangular (row tiles x N range) or a contiguous share of the
  // flattened (row tile, N tile) space
  const long flat_total = (long)m_tiles_total * n_nt_total;
  const int u0 = g.flat ? (int)(flat_total * blockIdx.x / gridDim.x) : 0;
  const int u1 = g.flat ? (int)(flat_total * (blockIdx.x + 1) / gridDim.x) : 0;
  const int n_tiles_cta = g.flat ? u1 - u0 : n_nt * n_mt;
  auto tile_nt = [&](int j) { return g.flat ? (u0 + j) % n_nt_total : nt_begin + j % n_nt; };
  auto tile_m0 = [&](int j) { return g.flat ? ((u0 + j) / n_nt_total) * kBM : m0 + (j / n_nt) * kBM; };
  const int n_iter = n_kb * n_tiles_cta;
  const int n_ck = (n_kb + kChunkKB - 1) / kChunkKB;   // accumulator chunks per N tile
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full + s, C::GROUP_THREADS); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 128); }
    for (int b = 0; b < C::EXTRA_BARS; ++b) mbar_init(pbar + b, 1);
    *passed = 0;
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, C::TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp >= 4 && warp < kMmaWarp) {
    // =============================== producers ===============================================
    const int grp = (tid - 128) / C::GROUP_THREADS, ptid = (tid - 128) % C::GROUP_THREADS;
    Block<kBM, A_MN, C::GROUP_THREADS> a;
    Block<NT, B_MN, C::GROUP_THREADS> b;
    for (int it = grp; it < n_iter; it += C::GROUPS) {
      const int tile = it / n_kb, kb = kb_begin + it % n_kb;
      const int nt = tile_nt(tile), m0t = tile_m0(tile);
      const int k_lim = min(g.K, kb_end * kKB);
      a.load(g.A, g.lda, m0t, g.M, kb * kKB, k_lim, ptid);
      b.load(g.B, g.ldb, nt * NT, g.N, kb * kKB, k_lim, ptid);
      const int s = it % C::STAGES;
      // There are more groups than stages, so a group could reach its wait on empty[s] while that
      // barrier is still TWO phases behind (the consumer of it - 2 * STAGES has not committed yet)
      // and a parity wait would then pass at once. The waits are therefore chained in iteration
      // order: block `it` waits only after the producer of block it - 1 is through its own wait.
      if (ptid == 0) {
        for (uint32_t spin = 0; *passed < it; ++spin)
          if (spin > (1u << 26)) __trap();
      }
      group_sync(1 + grp, C::GROUP_THREADS);
      mbar_wait(empty + s, ((it / C::STAGES) & 1) ^ 1);
      if (ptid == 0) *passed = it + 1;
      uint8_t *stage = smem + s * C::STAGE;
      a.store(stage, stage + C::A_HALF, ptid);
      b.store(stage + 2 * C::A_HALF, stage + 2 * C::A_HALF + C::B_HALF, ptid);
      fence_proxy_async_smem();
      mbar_arrive(full + s);
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ==============================================
    constexpr uint32_t idesc = idesc_tf32(kBM, NT, A_MN, B_MN);
    // K-major (SWIZZLE_128B): SBO = 1024 (next 8 rows); a K step of 8 floats is +32 bytes in the atom.
    // MN-major (SWIZZLE_128B_BASE32B): LBO = 512 (next 32 floats of M/N), SBO = next 4 K rows;
    // a K step of 8 is two such groups.
    constexpr uint32_t a_sbo = A_MN ? (kBM / 32) * 512 : 1024, b_sbo = B_MN ? (NT / 32) * 512 : 1024;
    constexpr uint32_t a_step = A_MN ? 2 * a_sbo : 32, b_step = B_MN ? 2 * b_sbo : 32;
    for (int it = 0; it < n_iter; ++it) {
      // accumulator chunk: kChunkKB K blocks of one N tile; chunks alternate between the two buffers
      const int s = it % C::STAGES, j = it / n_kb, kb = it % n_kb;
      const int ck = j * n_ck + kb / kChunkKB, kc = kb % kChunkKB, buf = ck & 1;
      mbar_wait(full + s, (it / C::STAGES) & 1);
      if (kc == 0) mbar_wait(tempty + buf, ((ck >> 1) & 1) ^ 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + buf * NT;
      const uint32_t st = smem_base + s * C::STAGE;
      const uint64_t a_hi = A_MN ? smem_desc_mn32(st, 512, a_sbo) : smem_desc_sw128(st, 16, 1024);
      const uint64_t a_lo = A_MN ? smem_desc_mn32(st + C::A_HALF, 512, a_sbo) : smem_desc_sw128(st + C::A_HALF, 16, 1024);
      const uint64_t b_hi = B_MN ? smem_desc_mn32(st + 2 * C::A_HALF, 512, b_sbo)
                                 : smem_desc_sw128(st + 2 * C::A_HALF, 16, 1024);
      const uint64_t b_lo = B_MN ? smem_desc_mn32(st + 2 * C::A_HALF + C::B_HALF, 512, b_sbo)
                                 : smem_desc_sw128(st + 2 * C::A_HALF + C::B_HALF, 16, 1024);
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint64_t a0 = pass == 0 ? a_lo : a_hi;
        const uint64_t b0 = pass == 1 ? b_lo : b_hi;
#pragma unroll
        for (int ks = 0; ks < kKB / 8; ++ks) {
          const uint64_t ad = a0 + ((ks * a_step) >> 4);
          const uint64_t bd = b0 + ((ks * b_step) >> 4);
          const uint32_t accumulate = (kc | pass | ks) != 0;
          if (elect_one()) umma_tf32_ss(d_tmem, ad, bd, idesc, accumulate);
        }
      }
      if (elect_one()) {
        umma_commit(empty + s);
        if (kc == kChunkKB - 1 || kb == n_kb - 1) umma_commit(tfull + buf);
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue ================================================
    // The tensor core truncates its fp32 accumulator after every instruction, a bias that grows with
    // the length of the chain; chains are therefore cut every kChunkKB K blocks (24 instructions) and
    // the chunks are added here in registers with round-to-nearest (fp32-class accuracy at K = 20k+).
    float *Cs = g.C + (size_t)blockIdx.y * g.slab;
    // EPI_ADAM: bias-corrected step size and sqrt(bias_correction2) exactly as adam_kernel (optim.cu)
    float step_size = 0.f, bc2_sqrt = 1.f, p2_acc = 0.f;
    if constexpr (EPI == EPI_ADAM) {
      const double lr = g.hyper[0], step = g.hyper[1];
      step_size = (float)(lr / (1.0 - pow(g.beta1, step)));
      bc2_sqrt = (float)sqrt(1.0 - pow(g.beta2, step));
    }
    // EPI_ADAM: this warp's 32 rows x NT columns of the table and of both moments travel HBM ->
    // shared memory -> HBM as TMA tile copies (boxes of 32 rows x 32 columns, 128-byte swizzle, three
    // arrays x NT / 32 boxes per tile): no registers and no LSU address work are spent on the 24
    // bytes per element that bound this kernel, the whole tile (3 x 32 x NT x 4 bytes per warp) is
    // in flight at once, and rows beyond the table are clipped by the copy engine.
    constexpr int NBOX = NT / 32;
    constexpr int NBUF = 64 / NT >= 2 ? 2 : 1;           // tile buffers per warp (same shared memory either way)
    constexpr uint32_t TILE_BYTES = 3u * NBOX * 4096u;
    uint8_t *tiles0 = nullptr;
    const int n_tiles = n_kb > 0 ? n_tiles_cta : 0;
    // tile jj of this warp -> its buffer jj % NBUF (issued by lane 0)
    auto issue_loads = [&](int jj) {
      uint8_t *t = tiles0 + (jj % NBUF) * TILE_BYTES;
      uint64_t *bar = pbar + warp * NBUF + jj % NBUF;
      const int nn = tile_nt(jj) * NT, mm = tile_m0(jj) + warp * 32;
      mbar_arrive_expect_tx(bar, TILE_BYTES);
#pragma unroll
      for (int b = 0; b < NBOX; ++b) {
        tma_load_2d(t + (0 * NBOX + b) * 4096, &maps.a, nn + 32 * b, mm, bar);
        tma_load_2d(t + (1 * NBOX + b) * 4096, &maps.b, nn + 32 * b, mm, bar);
        tma_load_2d(t + (2 * NBOX + b) * 4096, &maps.c, nn + 32 * b, mm, bar);
      }
    };
    if constexpr (EPI == EPI_ADAM) {
      uint8_t *base = reinterpret_cast<uint8_t *>(stg);
      base += (1024u - (smem_u32(base) & 1023u)) & 1023u;
      tiles0 = base + warp * (NBUF * TILE_BYTES);
      if (lane == 0) {
        tma_prefetch_desc(&maps.a); tma_prefetch_desc(&maps.b); tma_prefetch_desc(&maps.c);
        if (n_tiles > 0 && g.debug != 2) issue_loads(0);
      }
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int n0 = tile_nt(j) * NT, m0j = tile_m0(j);
      const int m = m0j + warp * 32 + lane;
      const int rows_valid = min(32, g.M - (m0j + warp * 32));
      if constexpr (EPI == EPI_ADAM) {
        // two buffers: the loads of tile j + 1 go out before tile j is touched (its buffer was last
        // read by the stores of tile j - 1); one buffer: tile j is fetched once tile j - 1 has left
        const int next = NBUF == 2 ? j + 1 : j;
        if (lane == 0 && g.debug != 2 && next > 0 && next < n_tiles) {
          bulk_wait_read();
          issue_loads(next);
        }
        __syncwarp();
      }
      float acc[NT];
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = 0.f;
      for (int c = 0; c < n_ck; ++c) {
        const int ck = j * n_ck + c, buf = ck & 1;
        mbar_wait(tfull + buf, (ck >> 1) & 1);
        fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * NT;
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) acc[c0 + q] += __uint_as_float(r[q]);
        }
        fence_before_sync();
        mbar_arrive(tempty + buf);
      }
      if constexpr (EPI == EPI_ADAM) {
        // lane = table row (the TMEM lane): acc[] is that row's gradient G = dY W. Same float
        // operations in the same order as adam_kernel (optim.cu), on the row held in shared memory
        // (16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4) inside its box: a warp's
        // access touches every bank exactly four times, the minimum for 512 bytes).
        uint8_t *tiles = tiles0 + (j % NBUF) * TILE_BYTES;
        if (g.debug != 2) mbar_wait(pbar + warp * NBUF + j % NBUF, (j / NBUF) & 1);
        if (lane < rows_valid && g.debug != 1) {
          const float beta2 = (float)g.beta2, w1 = (float)(1.0 - g.beta1), w2 = (float)(1.0 - g.beta2);
          const float inv_bc2 = 1.f / bc2_sqrt;
#pragma unroll
          for (int q = 0; q < NT; q += 4) {
            const uint32_t off = (uint32_t)(q / 32) * 4096u + (uint32_t)lane * 128u +
                                 (uint32_t)((((q % 32) / 4) ^ (lane & 7)) << 4);
            float4 *pq = reinterpret_cast<float4 *>(tiles + off);
            float4 *mq = reinterpret_cast<float4 *>(tiles + NBOX * 4096 + off);
            float4 *vq = reinterpret_cast<float4 *>(tiles + 2 * NBOX * 4096 + off);
            float4 p4 = *pq, m4 = *mq, v4 = *vq;
            // adam_kernel's update (optim.cu) with the square root and the two divisions on the
            // special-function unit (sqrt.approx / rcp.approx, <= 1 ulp each): the IEEE sequences
            // are ~10 dependent instructions with a slow-path branch each, and with one epilogue
            // warp per scheduler nothing hides them -- the arithmetic alone took 2.7x the HBM time
            // of the kernel. The update term m / denom is bounded by ~1, so the parameters move by
            // < 1e-9 relative against the exact formula (parity bar: 1e-5).
            auto upd = [&](float &p_, float g_, float &m_, float &v_) {
              g_ *= g.grad_scale;
              if (g.weight_decay != 0.f) g_ = fmaf(g.weight_decay, p_, g_);
              m_ = m_ + w1 * (g_ - m_);
              v_ = v_ * beta2 + w2 * g_ * g_;
              float sq, rc;
              asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v_));
              const float denom = fmaf(sq, inv_bc2, g.eps);
              asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(denom));
              p_ = p_ - step_size * (m_ * rc);
              p2_acc = fmaf(p_, p_, p2_acc);
            };
            upd(p4.x, acc[q], m4.x, v4.x); upd(p4.y, acc[q + 1], m4.y, v4.y);
            upd(p4.z, acc[q + 2], m4.z, v4.z); upd(p4.w, acc[q + 3], m4.w, v4.w);
            *pq = p4; *mq = m4; *vq = v4;
          }
        }
        fence_proxy_async_smem();                    // the rows written above -> visible to the copy engine
        __syncwarp();
        if (lane == 0 && g.debug != 2) {
#pragma unroll
          for (int b = 0; b < NBOX; ++b) {
            tma_store_2d(&maps.a, n0 + 32 * b, m0j + warp * 32, tiles + (0 * NBOX + b) * 4096);
            tma_store_2d(&maps.b, n0 + 32 * b, m0j + warp * 32, tiles + (1 * NBOX + b) * 4096);
            tma_store_2d(&maps.c, n0 + 32 * b, m0j + warp * 32, tiles + (2 * NBOX + b) * 4096);
          }
          bulk_commit();
        }
        continue;
      }
      if constexpr (EPI == EPI_SUMSQ) {
        // sum of squares of the product itself (the squared norm of a gradient that is never stored)
        if (m < g.M) {
#pragma unroll
          for (int q = 0; q < NT; ++q)
            if (n0 + q < g.N) p2_acc = fmaf(acc[q], acc[q], p2_acc);
        }
        continue;
      }
      if constexpr (!TRANS_OUT) {
        // TMEM hands every thread one ROW; storing it as-is touches 32 different 128-byte lines per
        // instruction (16 bytes each) and the LSU serialises them -- it was the bottleneck of the dx
        // GEMM. The warp's 32 x NT tile goes through shared memory instead and leaves as whole
        // rows: one STG.128 per NT/4 lanes, 512 contiguous bytes per instruction.
        constexpr int P = NT + 4, LPR = NT / 4, RPI = 32 / LPR;      // pitch, lanes per row, rows per instr
        float *tr = stg + warp * (32 * P);
        const int mw = m0j + warp * 32;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NT; q += 4)
          *reinterpret_cast<float4 *>(tr + lane * P + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
        __syncwarp();
        const int cq = (lane % LPR) * 4, n = n0 + cq;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias != nullptr && n < g.N) bv = ldg4(g.bias + n);
        if (n < g.N) {
#pragma unroll 8
          for (int r0 = 0; r0 < 32; r0 += RPI) {
            const int rr = r0 + lane / LPR;
            if (mw + rr < g.M) {
              float4 o = *reinterpret_cast<const float4 *>(tr + rr * P + cq);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              if (g.act == 1) { o.x = fast_tanh(o.x); o.y = fast_tanh(o.y); o.z = fast_tanh(o.z); o.w = fast_tanh(o.w); }
              else if (g.act == 2) {
                o.x = fast_sigmoid(o.x); o.y = fast_sigmoid(o.y); o.z = fast_sigmoid(o.z); o.w = fast_sigmoid(o.w);
              }
              *reinterpret_cast<float4 *>(Cs + (size_t)(mw + rr) * g.ldc + n) = o;
            }
          }
        }
      } else if (m < g.M) {
        // C^T: lane = column of the stored matrix, so every register is one coalesced row segment
#pragma unroll
        for (int q = 0; q < NT; ++q)
          if (n0 + q < g.N) Cs[(size_t)(n0 + q) * g.ldc + m] = acc[q];
      }
      __syncwarp();
    }
    if constexpr (EPI == EPI_ADAM) bulk_wait_all();          // this thread's last stores are performed
    if constexpr (EPI != EPI_STORE) {
      // EPI_ADAM: sum of the updated parameters' squares of this CTA -- the mirror-gradient step size
      // needs sum theta^2 right after this update (trainer.py:289-305), and this pass already holds
      // every theta in registers. EPI_SUMSQ: sum of the squared products. Fixed order: lanes by
      // butterfly, the four warps by index.
      if (g.sumsq_partial != nullptr) {
        const float w = warp_sum(p2_acc);
        asm volatile("bar.sync 9, 128;" ::: "memory");      // every epilogue warp is done with its staging tile
        if (lane == 0) stg[warp] = w;
        asm volatile("bar.sync 9, 128;" ::: "memory");
        if (tid == 0) {
          const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
          g.sumsq_partial[cta] = ((double)stg[0] + (double)stg[1]) + ((double)stg[2] + (double)stg[3]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    fence_after_sync();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN, int NT, bool TRANS_OUT>
int launch_tc05(const GemmArgs &g, int k_splits, int n_chunks, cudaStream_t stream) {
  using C = GCfg<NT>;
  const size_t smem = 1024 + (size_t)C::STAGES * C::STAGE + (2 * C::STAGES + 4) * 8 + 16 + C::EPI;
  auto kern = gemm_tc05_kernel<A_MN, B_MN, NT, TRANS_OUT>;
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int m_tiles = (g.M + kBM - 1) / kBM, mt_per = g.mt_per_cta > 1 ? g.mt_per_cta : 1;
  dim3 grid((m_tiles + mt_per - 1) / mt_per, k_splits, n_chunks);
  kern<<<grid, kThreadsG, smem, stream>>>(g, TmaMaps3{});
  MMREC_CHECK_LAUNCH("gemm_tc05_kernel");
  return MMREC_OK;
}

template <int NT, int EPI>
int launch_table_epi(const GemmArgs &g, const TmaMaps3 &maps, int n_chunks, cudaStream_t stream) {
  using C = GCfg<NT, EPI>;
  const size_t smem = 1024 + (size_t)C::STAGES * C::STAGE + (2 * C::STAGES + 4 + C::EXTRA_BARS) * 8 + 16 + C::EPI;
  auto kern = gemm_tc05_kernel<false, true, NT, false, EPI>;
  static bool attr = false;
  if (!attr) {
    MMREC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((g.M + kBM - 1) / kBM, 1, n_chunks);
  if (g.flat) grid = dim3(n_chunks, 1, 1);          // n_chunks = CTAs of the flat partition
  kern<<<grid, kThreadsG, smem, stream>>>(g, maps);
  MMREC_CHECK_LAUNCH(EPI == EPI_ADAM ? "gemm_tc05_kernel<adam>" : "gemm_tc05_kernel<sumsq>");
  return MMREC_OK;
}

template <bool A_MN, bool B_MN, bool TRANS_OUT>
int launch_tc05_nt(int nt, const GemmArgs &g, int k_splits, int n_chunks, cudaStream_t stream) {
  switch (nt) {
    case 32: return launch_tc05<A_MN, B_MN, 32, TRANS_OUT>(g, k_splits, n_chunks, stream);
    case 64: return launch_tc05<A_MN, B_MN, 64, TRANS_OUT>(g, k_splits, n_chunks, stream);
    case 128: return launch_tc05<A_MN, B_MN, 128, TRANS_OUT>(g, k_splits, n_chunks, stream);
    default: return 1;
  }
}

}  // namespace

// Which (layout, shape) combinations run on the tcgen05 kernel. The caller-facing problem is
// C[M,N] = op(A) op(B) with the a_kcontig / b_kcontig flags of mmrec_gemm_tf32x3_f32.
//   kind 1: A [M,K], B [N,K]      (forward)   -> split-K, N tile = N
//   kind 2: A [K,M], B [K,N]      (dW)        -> computed as C^T: UMMA M = N, UMMA N = M, split-K
//   kind 3: A [M,K], B [K,N]      (dx)        -> N ranges, no split-K
int gemm_tc05_kind(int M, int N, int K, int a_kcontig, int b_kcontig) {
  auto tile_ok = [](int n) { return n == 32 || n == 64 || n == 128; };
  // the table-sized projections, and the 128 x 128 dense layers over many rows (SMORE / Clothing,
  // d = 128: 62k node rows) where the mma.sync tile kernels run at a tenth of these
  const bool wide_dense = K == 128 || N == 128;
  if (a_kcontig && b_kcontig)
    return (((M >= 1024 && K >= 256) || (M >= 8192 && K == 128 && N == 128)) && tile_ok(N) && K % 4 == 0) ? 1 : 0;
  if (!a_kcontig && !b_kcontig)
    return (((N >= 1024 && K >= 256) || (K >= 8192 && M == 128 && N == 128)) && tile_ok(M) && N % 4 == 0) ? 2 : 0;
  if (a_kcontig && !b_kcontig)
    return (((M >= 1024 && N >= 1024) || (M >= 8192 && N == 128 && wide_dense)) && N % 128 == 0 && K % 4 == 0 && K <= 256) ? 3 : 0;
  return 0;
}

// Pick the number of equal parts (<= max_parts) of `units` work units per row tile that wastes the
// least of the last wave: one CTA per SM is resident (192 KB of shared memory), so 448 CTAs would
// run as 3 full waves + 4 stragglers, and every wave pays its own pipeline fill (first DRAM round
// trips, TMEM allocation). Cost = waves x (units per CTA + fill).
static int best_parts(int row_tiles, int units, int max_parts, int fill) {
  int best = 1;
  long best_cost = -1;
  for (int p = 1; p <= max_parts; ++p) {
    const int per = (units + p - 1) / p, parts = (units + per - 1) / per;
    if (parts != p) continue;
    const long ctas = (long)row_tiles * parts, waves = (ctas + kNumSMs - 1) / kNumSMs;
    const long cost = waves * (per + fill);            // fill: pipeline ramp of a CTA, in work units
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = p; }
  }
  return best;
}

// The split-K GEMMs (forward, dW) with several row tiles: a CTA may also walk `mt` consecutive row tiles of its K
// range with the pipeline kept full (mt_per_cta), which turns e.g. 56 row tiles x 5 splits = 280 CTAs = two waves of
// 13 units + two fills into 28 x 5 = 140 CTAs = one wave of 26 units + one fill. Same cost model, searched over
// (parts, mt); returns the parts, *mt_out the row tiles per CTA. MMREC_GEMM_MT=0 keeps mt = 1 (A/B switch).
static long parts_cost(int row_tiles, int units, int p, int mt, int fill) {
  const int per = (units + p - 1) / p, parts = (units + per - 1) / per;
  if (parts != p) return -1;
  const long ctas = (long)((row_tiles + mt - 1) / mt) * parts, waves = (ctas + kNumSMs - 1) / kNumSMs;
  return waves * ((long)per * mt + fill);
}
static int best_mt(int row_tiles, int units, int p, int fill) {
  static const bool on = !(getenv("MMREC_GEMM_MT") && atoi(getenv("MMREC_GEMM_MT")) == 0);
  int best = 1;
  long best_cost = parts_cost(row_tiles, units, p, 1, fill);
  if (!on || best_cost < 0) return 1;
  for (int mt = 2; mt <= 4 && mt <= row_tiles; ++mt) {
    const long c = parts_cost(row_tiles, units, p, mt, fill);
    if (c >= 0 && c < best_cost) { best_cost = c; best = mt; }
  }
  return best;
}
static int best_parts_mt(int row_tiles, int units, int max_parts, int fill) {
  int best = 1;
  long best_cost = -1;
  for (int p = 1; p <= max_parts; ++p) {
    const int mt = best_mt(row_tiles, units, p, fill);
    long cost = parts_cost(row_tiles, units, p, mt, fill);
    if (cost < 0) continue;
    cost = 2 * cost + p;          // every split writes a slab and the reduce reads it: ~half a work unit each
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = p; }
  }
  return best;
}

int gemm_tc05_splits(int M, int N, int K, int kind) {
  if (kind == 3) return 1;
  const int m_tiles = kind == 2 ? (N + kBM - 1) / kBM : (M + kBM - 1) / kBM;
  const int n_kb = (K + kKB - 1) / kKB;
  // a single row tile (128 x 128 weight gradients over tens of thousands of rows): the reduction
  // dimension is all the parallelism there is
  const int units = (n_kb + kChunkKB - 1) / kChunkKB;
  if (m_tiles == 1) return best_parts(m_tiles, units, 128, 3);
  return best_parts_mt(m_tiles, units, 32, 3);
}

// Returns MMREC_OK, a negative error, or 1 when the shape is not covered.
int gemm_tc05_dispatch(const float *A, int a_kcontig, const float *B, int b_kcontig, const float *bias, float *C,
                       int M, int N, int K, int splits, float *ws, cudaStream_t stream, int act = 0) {
  const int kind = gemm_tc05_kind(M, N, K, a_kcontig, b_kcontig);
  if (kind == 0) return 1;
  if (act != 0 && (kind != 1 || splits != 1)) return 1;    // the activation rides in the store epilogue only
  GemmArgs g{};
  g.act = act;
  const int n_kb = (K + kKB - 1) / kKB, n_units = (n_kb + kChunkKB - 1) / kChunkKB;
  g.kb_per_split = (n_units + splits - 1) / splits * kChunkKB;
  const int k_splits = (n_kb + g.kb_per_split - 1) / g.kb_per_split;
  if (kind != 3 && k_splits != splits) return 1;       // workspace laid out for another split count
  float *out = splits > 1 ? ws : C;
  g.bias = splits > 1 ? nullptr : bias;
  g.slab = (size_t)M * N;
  // more row tiles than SMs (and nothing else to parallelise over): consecutive tiles per CTA
  auto tiles_per_cta = [](int m_tiles, int other) {
    return other == 1 && m_tiles > kNumSMs ? (m_tiles + kNumSMs - 1) / kNumSMs : 1;
  };
  if (kind == 1) {
    g.A = A; g.lda = K; g.B = B; g.ldb = K; g.C = out; g.ldc = N; g.M = M; g.N = N; g.K = K; g.nt_per_cta = 1;
    g.mt_per_cta = k_splits > 1 ? best_mt((M + kBM - 1) / kBM, n_units, k_splits, 3) : tiles_per_cta((M + kBM - 1) / kBM, k_splits);
    return launch_tc05_nt<false, false, false>(N, g, k_splits, 1, stream);
  }
  if (kind == 2) {
    // C^T [N, M] = B^T [N, K] * A [K, M]: UMMA A = caller's B (MN-major), UMMA B = caller's A (MN-major)
    g.A = B; g.lda = N; g.B = A; g.ldb = M; g.C = out; g.ldc = N; g.M = N; g.N = M; g.K = K; g.nt_per_cta = 1;
    if (bias != nullptr && splits == 1) return 1;      // bias indexes the other axis here
    if (k_splits > 1) g.mt_per_cta = best_mt((N + kBM - 1) / kBM, n_units, k_splits, 3);
    return launch_tc05_nt<true, true, true>(M, g, k_splits, 1, stream);
  }
  // kind 3
  if (splits != 1) return 1;
  g.A = A; g.lda = K; g.B = B; g.ldb = N; g.C = C; g.ldc = N; g.M = M; g.N = N; g.K = K; g.bias = bias;
  g.kb_per_split = n_kb;
  // 64-wide N tiles: four producer groups and three stages instead of two and two, no register
  // spills in the epilogue -- 60 -> 50 us for the 7050 x 4096 image-table gradient (MMREC_DX_NT=128: old tiling)
  static const int dx_nt = getenv("MMREC_DX_NT") ? atoi(getenv("MMREC_DX_NT")) : 64;
  const int nt = dx_nt == 128 ? 128 : 64;
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = N / nt;
  int chunks = best_parts(m_tiles, n_tiles, n_tiles, 1);
  g.nt_per_cta = (n_tiles + chunks - 1) / chunks;
  chunks = (n_tiles + g.nt_per_cta - 1) / g.nt_per_cta;
  g.mt_per_cta = tiles_per_cta(m_tiles, chunks);
  if (nt == 64) return launch_tc05<false, true, 64, false>(g, 1, chunks, stream);
  return launch_tc05<false, true, 128, false>(g, 1, chunks, stream);
}


// ---- Adam on a feature table whose gradient is the rank-d product dY W (never materialised) ----
static int table_adam_nt() {
  static const int nt = getenv("MMREC_TA_NT") && atoi(getenv("MMREC_TA_NT")) == 64 ? 64 : 32;
  return nt;
}

static bool table_flat() {      // MMREC_TA_FLAT=0: the rectangular partition (A/B switch)
  static const bool on = !(getenv("MMREC_TA_FLAT") && atoi(getenv("MMREC_TA_FLAT")) == 0);
  return on;
}

static int table_grid(int rows, int cols, int *nt_per_cta, int nt) {
  const int m_tiles = (rows + kBM - 1) / kBM, n_tiles = cols / nt;
  if (table_flat()) {
    if (nt_per_cta) *nt_per_cta = n_tiles;
    return (int)min((long)kNumSMs, (long)m_tiles * n_tiles);
  }
  int chunks = best_parts(m_tiles, n_tiles, n_tiles, 1);
  const int per = (n_tiles + chunks - 1) / chunks;
  chunks = (n_tiles + per - 1) / per;
  if (nt_per_cta) *nt_per_cta = per;
  return m_tiles * chunks;
}

// number of CTAs (= per-CTA partial sums) of the two table kernels
int table_adam_ctas(int rows, int cols) { return table_grid(rows, cols, nullptr, table_adam_nt()); }
int table_sumsq_ctas(int rows, int cols) { return table_grid(rows, cols, nullptr, 64); }

int table_adam_dispatch(float *P, float *Mo, float *V, const float *dY, const float *W, int rows, int cols, int d,
                        const double *hyper, double beta1, double beta2, double eps, double weight_decay,
                        double grad_scale, double *sumsq_partial, cudaStream_t stream) {
  GemmArgs g{};
  g.A = dY; g.lda = d; g.B = W; g.ldb = cols; g.C = P; g.ldc = cols; g.M = rows; g.N = cols; g.K = d;
  g.kb_per_split = (d + kKB - 1) / kKB;
  g.exp_avg = Mo; g.exp_avg_sq = V; g.hyper = hyper; g.beta1 = beta1; g.beta2 = beta2; g.eps = (float)eps;
  g.weight_decay = (float)weight_decay; g.grad_scale = (float)grad_scale; g.sumsq_partial = sumsq_partial;
  const int nt = table_adam_nt();
  const int ctas = table_grid(rows, cols, &g.nt_per_cta, nt);
  const int m_tiles = (rows + kBM - 1) / kBM;
  static const int dbg = getenv("MMREC_TA_DEBUG") ? atoi(getenv("MMREC_TA_DEBUG")) : 0;
  g.debug = dbg;
  TmaMaps3 maps;
  if (!tma_encode_2d_f32(&maps.a, P, rows, cols, cols, 32, 32) || !tma_encode_2d_f32(&maps.b, Mo, rows, cols, cols, 32, 32) ||
      !tma_encode_2d_f32(&maps.c, V, rows, cols, cols, 32, 32)) {
    set_error("table_adam: cuTensorMapEncodeTiled failed");
    return MMREC_E_CUDA;
  }
  g.flat = table_flat();
  if (nt == 32) return launch_table_epi<32, EPI_ADAM>(g, maps, g.flat ? ctas : ctas / m_tiles, stream);
  return launch_table_epi<64, EPI_ADAM>(g, maps, g.flat ? ctas : ctas / m_tiles, stream);
}

// per-CTA partial sums of ||dY W||_F^2 (same grid as table_adam_dispatch)
int table_sumsq_dispatch(const float *dY, const float *W, int rows, int cols, int d, double *sumsq_partial,
                         cudaStream_t stream) {
  GemmArgs g{};
  g.A = dY; g.lda = d; g.B = W; g.ldb = cols; g.C = nullptr; g.ldc = cols; g.M = rows; g.N = cols; g.K = d;
  g.kb_per_split = (d + kKB - 1) / kKB;
  g.sumsq_partial = sumsq_partial;
  const int ctas = table_grid(rows, cols, &g.nt_per_cta, 64);
  const int m_tiles = (rows + kBM - 1) / kBM;
  g.flat = table_flat();
  return launch_table_epi<64, EPI_SUMSQ>(g, TmaMaps3{}, g.flat ? ctas : ctas / m_tiles, stream);
}

}  // namespace mmrec
