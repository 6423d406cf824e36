// Access-pattern micro-benchmark: a CTA streams a [128 rows x KC floats] panel of a row-major
// [R x 4096] float table, RUN bytes contiguous per row visit, DEPTH float4 loads in flight per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int RUN_F4, int DEPTH>   // RUN_F4: float4 per row visit (8 = 128 B, 32 = 512 B, 128 = 2 KB)
__global__ void __launch_bounds__(256) stream_kernel(const float4 *__restrict__ T, int R, int ld4, int kc4, float *out) {
  // CTA tile: 128 rows, k range [blockIdx.y*kc4, +kc4) in float4 units
  const int m0 = blockIdx.x * 128, k0 = blockIdx.y * kc4;
  float acc = 0.f;
  constexpr int ROWS_PER_STEP = 256 / RUN_F4 > 128 ? 128 : 256 / RUN_F4;       // rows covered by one CTA-wide load
  const int lr = threadIdx.x / RUN_F4, lc = threadIdx.x % RUN_F4;
  // iterate: for kb (RUN_F4 wide) over kc4, for row group over 128 rows
  const int n_kb = kc4 / RUN_F4, n_rg = 128 / ROWS_PER_STEP, n = n_kb * n_rg;
  float4 v[DEPTH];
  for (int i0 = 0; i0 < n; i0 += DEPTH) {
#pragma unroll
    for (int u = 0; u < DEPTH; ++u) {
      const int i = i0 + u, kb = i / n_rg, rg = i % n_rg;
      const int row = m0 + rg * ROWS_PER_STEP + lr;
      v[u] = (i < n && row < R && lr < ROWS_PER_STEP) ? __ldg(T + (size_t)row * ld4 + k0 + kb * RUN_F4 + lc) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < DEPTH; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}
template <int RUN_F4, int DEPTH>
void run(const float4 *T, int R, int splits, float *out, const char *name) {
  const int ld4 = 1024, kc4 = ld4 / splits;
  dim3 grid((R + 127) / 128, splits);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 2; ++w) stream_kernel<RUN_F4, DEPTH><<<grid, 256>>>(T, R, ld4, kc4, out);
  cudaEventRecord(a);
  for (int w = 0; w < 10; ++w) stream_kernel<RUN_F4, DEPTH><<<grid, 256>>>(T, R, ld4, kc4, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
  printf("%-28s R=%d splits=%d ctas=%d: %.1f us  %.0f GB/s\n", name, R, splits, grid.x * grid.y, ms * 1e3, (double)R * 16384 / ms / 1e6);
}
int main() {
  for (int R : {7050, 28200}) {
    float4 *T; float *out; cudaMalloc(&T, (size_t)R * 16384); cudaMalloc(&out, 4); cudaMemset(T, 0, (size_t)R * 16384);
    run<8, 8>(T, R, 4, out, "run128B depth8");
    run<8, 16>(T, R, 4, out, "run128B depth16");
    run<8, 16>(T, R, 8, out, "run128B depth16");
    run<32, 8>(T, R, 4, out, "run512B depth8");
    run<32, 16>(T, R, 4, out, "run512B depth16");
    run<32, 16>(T, R, 8, out, "run512B depth16");
    run<128, 16>(T, R, 4, out, "run2KB depth16");
    run<128, 16>(T, R, 8, out, "run2KB depth16");
    run<256, 16>(T, R, 4, out, "run4KB depth16");
    cudaFree(T); cudaFree(out);
  }
  return 0;
}
