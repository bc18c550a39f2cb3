// Microbenchmark behind DESIGN.md §4.1: how fast can the SMs pull random, fully coalesced 1-KB rows
// out of an L2-resident table?  (The forward kernels do exactly this, ~230 rows per anchor.)
// A warp reads whole rows (two 512-byte warp loads), U rows in flight per lane, pseudo-random row
// indices; the table (default 88 MB) fits the 126 MB L2 and is touched once before timing.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int U>
__global__ void __launch_bounds__(256) gather_rows(const float4 *__restrict__ table, uint32_t rows,
                                                    int iters, float *__restrict__ sink) {
  const int lane = threadIdx.x & 31;
  uint32_t state = (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float4 v[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      state = state * 1664525u + 1013904223u;
      const uint32_t r = __shfl_sync(0xffffffffu, state, 0) % rows;  // one row per warp
      const float4 *p = table + static_cast<size_t>(r) * 64 + lane;
      v[u][0] = __ldg(p);
      v[u][1] = __ldg(p + 32);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u][0].x + v[u][1].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int U>
static void run(const float4 *table, uint32_t rows, float *sink, int ctas_per_sm) {
  const int grid = 148 * ctas_per_sm, iters = 4096 / U;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  gather_rows<U><<<grid, 256>>>(table, rows, iters / 8, sink);  // warm-up: table into L2
  cudaEventRecord(e0);
  gather_rows<U><<<grid, 256>>>(table, rows, iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = static_cast<double>(grid) * 8 * iters * U * 1024.0;
  printf("rows in flight per lane %d, %d CTAs/SM: %.2f TB/s (%.1f GB in %.3f ms)\n", U, ctas_per_sm,
         bytes / ms / 1e9, bytes / 1e9, ms);
}

int main(int argc, char **argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 88;
  const uint32_t rows = static_cast<uint32_t>(mb * 1024);
  float4 *table;
  float *sink;
  cudaMalloc(&table, static_cast<size_t>(rows) * 1024);
  cudaMalloc(&sink, 4);
  cudaMemset(table, 0, static_cast<size_t>(rows) * 1024);
  printf("table %zu MB (%u rows of 1 KB)\n", mb, rows);
  run<2>(table, rows, sink, 8);
  run<4>(table, rows, sink, 8);
  run<8>(table, rows, sink, 4);
  run<4>(table, rows, sink, 6);
  run<8>(table, rows, sink, 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
