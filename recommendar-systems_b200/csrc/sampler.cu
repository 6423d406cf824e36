// Host-side negative sampler: a bit-exact replay of the reference's draw
//   iid = random.sample(all_items, 1)[0]; while iid in history[u]: redraw
// (utils/dataloader.py:267-275, 307-309) on CPython's global Mersenne Twister.
//
// CPython's random.sample(population, 1) reduces to population[_randbelow(n)], and
// _randbelow_with_getrandbits(n) is "k = n.bit_length(); r = getrandbits(k); while r >= n: redraw",
// with getrandbits(k <= 32) = genrand_uint32() >> (32 - k) (Modules/_randommodule.c). The caller
// hands over random.getstate()[1] (624 state words + position) and writes it back afterwards, so
// the Python stream continues exactly where the reference's would. No CUDA here: this is the
// producer side of the [3, B] batch, ~100x faster than the Python loop it replaces.
#include "common.cuh"

namespace mmrec {
namespace {

constexpr int kN = 624, kM = 397;

inline void mt_reload(uint32_t *mt) {
  constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  int kk;
  uint32_t y;
  for (kk = 0; kk < kN - kM; ++kk) {
    y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + kM] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; kk < kN - 1; ++kk) {
    y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + (kM - kN)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  y = (mt[kN - 1] & UPPER) | (mt[0] & LOWER);
  mt[kN - 1] = mt[kM - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

inline uint32_t mt_next(uint32_t *mt, uint32_t &pos) {
  if (pos >= (uint32_t)kN) {
    mt_reload(mt);
    pos = 0;
  }
  uint32_t y = mt[pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

}  // namespace
}  // namespace mmrec

using namespace mmrec;

extern "C" int mmrec_neg_sample_mt19937_host(uint32_t *mt_state_host, const int64_t *all_items_host,
                                             int64_t n_items, const int64_t *hist_rowptr_host,
                                             const int64_t *hist_cols_host, int64_t n_hist_users,
                                             const int64_t *users_host, int64_t n, int64_t *neg_out_host) {
  MMREC_REQUIRE(mt_state_host && all_items_host && hist_rowptr_host && hist_cols_host && users_host && neg_out_host,
                MMREC_E_BADARG, "neg_sample: null pointer");
  MMREC_REQUIRE(n_items > 0 && n_items < (1ll << 32) && n >= 0, MMREC_E_BADARG,
                "neg_sample: need 0 < n_items < 2^32 (getrandbits path for wider ranges is not replayed)");
  MMREC_REQUIRE(mt_state_host[kN] <= (uint32_t)kN, MMREC_E_BADARG, "neg_sample: bad Mersenne Twister position");
  uint32_t pos = mt_state_host[kN];
  int bits = 0;
  for (uint64_t v = (uint64_t)n_items; v; v >>= 1) ++bits;   // n.bit_length()
  const int shift = 32 - bits;
  for (int64_t b = 0; b < n; ++b) {
    const int64_t u = users_host[b];
    MMREC_REQUIRE(u >= 0 && u < n_hist_users, MMREC_E_BADARG, "neg_sample: user id %lld out of range", (long long)u);
    const int64_t lo = hist_rowptr_host[u], hi = hist_rowptr_host[u + 1];
    MMREC_REQUIRE(hi - lo < n_items, MMREC_E_BADARG,
                  "neg_sample: user %lld has interacted with every item (the reference would loop forever)",
                  (long long)u);
    for (;;) {
      uint32_t r;
      do {
        r = mt_next(mt_state_host, pos) >> shift;
      } while ((int64_t)r >= n_items);
      const int64_t item = all_items_host[r];
      // binary search in the user's ascending history
      int64_t a = lo, z = hi;
      while (a < z) {
        const int64_t mid = (a + z) >> 1;
        if (hist_cols_host[mid] < item) a = mid + 1; else z = mid;
      }
      if (a < hi && hist_cols_host[a] == item) continue;
      neg_out_host[b] = item;
      break;
    }
  }
  mt_state_host[kN] = pos;
  return MMREC_OK;
}

// ---- counter-based negative sampler on the device ---------------------
// The reference draws `iid = random.sample(all_items, 1)[0]` and re-draws while iid is in the
// user's training history (utils/dataloader.py:267-275, 307-309) -- one Python iteration per
// interaction, 500 M per epoch at config-5 scale. Here every (step, position) owns a counter-based
// stream: draw a = 0, 1, ... is item all_items[mix(seed, step, position, a) mod n_items], the first
// one outside the user's (ascending) history wins. Stateless, so any rank regenerates any batch
// without communication; same distribution as the reference (uniform over the items, rejection on
// the history), a different stream by construction (the bit-exact Mersenne-Twister replay above is
// what the parity configurations use). oracle/sampler.py restates the mixer in numpy: the ids are
// compared bit for bit.
namespace mmrec {
namespace {

__host__ __device__ inline uint64_t mix64(uint64_t z) {       // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
neg_sample_counter_kernel(const int64_t *__restrict__ users, int64_t n, const int64_t *__restrict__ all_items,
                          int64_t n_items, const int64_t *__restrict__ hist_rowptr,
                          const int32_t *__restrict__ hist_cols, uint64_t seed, uint64_t step, int32_t max_draws,
                          int64_t *__restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (b >= n) return;
  const int64_t u = users[b];
  const int64_t lo0 = hist_rowptr[u], hi0 = hist_rowptr[u + 1];
  const uint64_t base = mix64(mix64(seed ^ (step * 0xd1342543de82ef95ull)) + (uint64_t)b);
  int64_t item = -1;
  for (int a = 0; a < max_draws; ++a) {
    const uint64_t r = mix64(base + (uint64_t)a * 0x2545f4914f6cdd1dull);
    const int64_t cand = all_items ? all_items[(r >> 11) % (uint64_t)n_items] : (int64_t)((r >> 11) % (uint64_t)n_items);
    // binary search in the user's ascending history
    int64_t lo = lo0, hi = hi0;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)hist_cols[mid] < cand) lo = mid + 1; else hi = mid;
    }
    if (!(lo < hi0 && (int64_t)hist_cols[lo] == cand)) { item = cand; break; }
  }
  out[b] = item;            // -1: max_draws exhausted (a user who interacted with ~every item)
}

}  // namespace
}  // namespace mmrec

extern "C" int mmrec_neg_sample_counter(const int64_t *users, int64_t n, const int64_t *all_items, int64_t n_items,
                                        const int64_t *hist_rowptr, const int32_t *hist_cols, uint64_t seed,
                                        uint64_t step, int32_t max_draws, int64_t *neg_out, void *stream) {
  MMREC_REQUIRE(users && hist_rowptr && hist_cols && neg_out, MMREC_E_BADARG, "neg_sample_counter: null pointer");
  MMREC_REQUIRE(n >= 0 && n_items > 0 && max_draws > 0, MMREC_E_BADARG, "neg_sample_counter: bad sizes");
  if (n == 0) return MMREC_OK;
  mmrec::neg_sample_counter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      users, n, all_items, n_items, hist_rowptr, hist_cols, seed, step, max_draws, neg_out);
  MMREC_CHECK_LAUNCH("neg_sample_counter_kernel");
  return MMREC_OK;
}
