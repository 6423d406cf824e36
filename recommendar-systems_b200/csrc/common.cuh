// Shared helpers for the mmrec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmrec_b200.h"

namespace mmrec {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define MMREC_REQUIRE(cond, code, ...)        \
  do {                                        \
    if (!(cond)) {                            \
      ::mmrec::set_error(__VA_ARGS__);        \
      return (code);                          \
    }                                         \
  } while (0)

#define MMREC_CHECK_LAUNCH(name)                                                  \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::mmrec::set_error("%s: %s", name, cudaGetErrorString(e__));                \
      return MMREC_E_CUDA;                                                        \
    }                                                                             \
    ::mmrec::count_launch();                                                      \
  } while (0)

#define MMREC_CUDA(call)                                                          \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::mmrec::set_error("%s: %s", #call, cudaGetErrorString(e__));               \
      return MMREC_E_CUDA;                                                        \
    }                                                                             \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (griddepcontrol) ----------------------------------------------
// A kernel launched with launch_pdl(..., pdl = true) may start while the kernel before it on the
// stream is still running, once every CTA of that kernel has executed pdl_trigger() (or exited). What
// it does before pdl_wait() overlaps the predecessor's tail: only data no kernel of the step writes
// (CSR arrays, task lists) may be touched there. pdl_wait() returns when the predecessor grid has
// completed and its writes are visible. Without the launch attribute both instructions are no-ops.
// Works inside stream capture (the edge becomes a programmatic graph edge). The attribute is set only
// under MMREC_PDL=1 (see pdl_enabled() in runtime.cu for the measurement).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- in-kernel dropout (nn.Dropout of smore.py:331-333 without mask tensors) --------------------
// The multiplier of element (plane, row, 4-column group) is a pure function of a 64-bit stream key
// and the element index: forward and backward regenerate the same mask, nothing is written to HBM
// (a [3, n, d] mask costs one fill, one dropout kernel and two 4 n d reads per pass). The stream key
// mixes a host constant (model seed, call index: baked into a captured graph) with a DEVICE counter
// (the optimizer's update count), so replays of one captured step draw fresh masks.
struct DropSpec {
  const double *counter;  // device: one double holding an integer count (FusedAdam's update count), or NULL
  uint64_t seed;          // host constant of this call
  float p;                // drop probability, applied in steps of 2^-16; 0 = no dropout
  // Batch-row calls (the module evaluated on gathered rows only): row m of the call draws the multipliers
  // of row row_ids[m] of a dense [n_total, d] call -- the compact and the dense evaluation drop the same
  // elements. NULL = identity (row m of n).
  const long long *row_ids;
  int n_total;
};
// first float4 index of (plane g, row m) in the [planes, n, d] index space; m < n
__device__ __forceinline__ uint64_t drop_row4(const DropSpec &d, int g, int m, int n, int d4) {
  const uint64_t row = d.row_ids != nullptr ? (uint64_t)d.row_ids[m] : (uint64_t)m;
  const uint64_t nt = d.row_ids != nullptr ? (uint64_t)d.n_total : (uint64_t)n;
  return ((uint64_t)g * nt + row) * (uint64_t)d4;
}
__host__ __device__ inline uint64_t drop_mix64(uint64_t z) {       // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t drop_stream(const DropSpec &d) {
  const uint64_t c = d.counter != nullptr ? (uint64_t)d.counter[0] : 0ull;
  return drop_mix64(d.seed ^ (c * 0xd1342543de82ef95ull));
}
// multipliers of the 4 consecutive elements starting at flat index `idx4 * 4` (16 bits each)
__device__ __forceinline__ float4 drop_mask4(uint64_t stream, uint64_t idx4, float p) {
  const uint64_t r = drop_mix64(stream + idx4 * 0x2545f4914f6cdd1dull);
  const uint32_t thr = (uint32_t)(p * 65536.f);
  const float s = 1.f / (1.f - p);
  return make_float4(((r) & 0xffffu) >= thr ? s : 0.f, ((r >> 16) & 0xffffu) >= thr ? s : 0.f,
                     ((r >> 32) & 0xffffu) >= thr ? s : 0.f, ((r >> 48) & 0xffffu) >= thr ? s : 0.f);
}

// ---- activations on the special-function unit ----------------------------------------------------
// expf / tanhf / IEEE division are 12-25 instruction sequences with slow-path branches; the fused
// kernels evaluate 7 activations per element of an [n, d] tile (the SMORE preference module spends a
// third of its issue slots there). These forms are 4-14 instructions: ex2.approx / rcp.approx
// (<= 2 ulp each) plus, for tanh, a degree-4 polynomial in x^2 below |x| = 0.6 where the exponential
// form would cancel. Measured against float64 over the whole range: tanh 3e-7, sigmoid 2e-7, exp 4e-7
// relative (parity bar 1e-5; tests/test_gpu_round2b.py::test_fast_activations).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_exp(float x) { return fast_ex2(x * 1.4426950408889634f); }
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.f + fast_exp(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float u = x * x, a = fabsf(x);
  float q = fmaf(u, -0.006276389118283987f, 0.021116215735673904f);
  q = fmaf(q, u, -0.053875137120485306f);
  q = fmaf(q, u, 0.13332924246788025f);
  q = fmaf(q, u, -0.3333333134651184f);
  const float small = fmaf(x * u, q, x);                                   // x + x^3 Q(x^2), |x| < 0.6
  const float e = fast_ex2(a * 2.8853900817779268f);                       // e^{2|x|}; inf -> rcp = 0 -> 1
  const float big = copysignf(fmaf(-2.f, fast_rcp(1.f + e), 1.f), x);
  return a < 0.6f ? small : big;
}

__device__ __forceinline__ float4 ldg4(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void fma4(float4 &a, float s, const float4 &x) {
  a.x = fmaf(s, x.x, a.x);
  a.y = fmaf(s, x.y, a.y);
  a.z = fmaf(s, x.z, a.z);
  a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// Butterfly sum over an aligned sub-warp of WIDTH lanes. The shuffle mask names only that
// sub-warp, so sibling sub-warps of the same warp may be divergent or exited.
template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {
  unsigned mask = 0xffffffffu;
  if constexpr (WIDTH < 32) {
    const unsigned lane = threadIdx.x & 31u;
    mask = ((1u << WIDTH) - 1u) << (WIDTH * (lane / WIDTH));
  }
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, WIDTH);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v); }

// Deterministic block reduction of `v` (valid in every thread) -> result in thread 0.
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float *smem /* THREADS/32 floats */) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < THREADS / 32 ? smem[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

}  // namespace mmrec
