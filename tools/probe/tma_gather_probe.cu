// Microbenchmark behind DESIGN.md §4.1: random 1-KB rows pulled by TMA bulk copies (cp.async.bulk
// global -> shared, completion on mbarriers) instead of register loads.  The question it answers: can
// shared memory serve as the landing zone that lets an SM keep more bytes in flight than its register
// file allows (98 KB with the row-sliced kernel), and what does the TMA unit sustain with 1-KB copies?
// Every warp runs its own ring of R slots: lane 0 issues the copies, all lanes wait on the slot's
// mbarrier, read the row back (two LDS.128 per lane) and accumulate, then the slot is refilled.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int R, int NW>
__global__ void __launch_bounds__(NW * 32) tma_rows(const float4 *__restrict__ table, uint32_t rows, int iters,
                                                    float *__restrict__ sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *ring = reinterpret_cast<float4 *>(smem) + static_cast<size_t>(warp) * R * 64;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(NW) * R * 1024) + warp * R;
  uint32_t state = (blockIdx.x * NW + warp) * 2654435761u + 12345u;
  if (lane == 0) {
    for (int r = 0; r < R; ++r)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + r)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int slot) {
    state = state * 1664525u + 1013904223u;
    const float4 *src = table + static_cast<size_t>(state % rows) * 64;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(s32(bars + slot)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::
                     "r"(s32(ring + slot * 64)), "l"(src), "r"(s32(bars + slot)) : "memory");
  };
  if (lane == 0)
    for (int r = 0; r < R; ++r) issue(r);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int slot = it % R;
    const uint32_t parity = (it / R) & 1;
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\t"
        "bra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bars + slot)), "r"(parity) : "memory");
    const float4 a = ring[slot * 64 + lane], b = ring[slot * 64 + 32 + lane];
    acc = fmaf(a.x, 1.0001f, acc) + b.w + a.z * b.y;
    __syncwarp();
    if (lane == 0 && it + R < iters) issue(slot);
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int R, int NW>
static void run(const float4 *table, uint32_t rows, float *sink, int ctas_per_sm) {
  const int grid = 148 * ctas_per_sm, iters = 2048;
  const size_t smem = static_cast<size_t>(NW) * R * 1024 + NW * R * 8;
  cudaFuncSetAttribute(tma_rows<R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  tma_rows<R, NW><<<grid, NW * 32, smem>>>(table, rows, iters / 8, sink);
  cudaEventRecord(e0);
  tma_rows<R, NW><<<grid, NW * 32, smem>>>(table, rows, iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = static_cast<double>(grid) * NW * iters * 1024.0;
  printf("TMA rows: %d slots/warp x %d warps x %d CTAs/SM = %3d KB in flight per SM: %6.2f TB/s (%s)\n", R, NW,
         ctas_per_sm, R * NW * ctas_per_sm, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
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
  run<2, 8>(table, rows, sink, 6);
  run<4, 8>(table, rows, sink, 6);
  run<4, 8>(table, rows, sink, 3);
  run<8, 8>(table, rows, sink, 3);
  run<8, 4>(table, rows, sink, 6);
  run<4, 4>(table, rows, sink, 12);
  run<16, 4>(table, rows, sink, 3);
  run<24, 8>(table, rows, sink, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
