// Probe: in which order does cluster-launch-control hand out the pending clusters of a 1-D grid?
// Launches `n` clusters of 2 CTAs that each need a whole SM (200 KB smem); every running cluster records
// (globaltimer, claimed cluster id), spins ~2 us, then cancels another pending cluster.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/clc_order csrc/tests/clc_order.cu && build/clc_order 2048
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "../common/ptx.cuh"
using namespace tb;

struct Rec { unsigned long long t; int id; int smid; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(Rec* rec, int* count) {
  extern __shared__ uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 127u) & ~127u;
  const uint32_t bar = base, resp = base + 64;
  const bool leader = cluster_ctarank() == 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  cluster_sync();
  if (threadIdx.x == 0) {
    int id = blockIdx.x / 2;
    uint32_t smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    for (int it = 0;; ++it) {
      if (leader) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int slot = atomicAdd(count, 1);
        rec[slot] = Rec{t, id, (int)smid};
      }
      const long long t0 = clock64();
      while (clock64() - t0 < 3000) {}
      mbar_arrive_expect_tx(bar, 16);
      if (leader) clc_try_cancel<true>(resp, bar);
      mbar_wait(bar, it & 1);
      uint32_t x;
      if (!clc_query(resp, x)) break;
      id = (int)(x / 2);
    }
  }
  cluster_sync();
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1024;
  Rec* rec; int* count;
  cudaMalloc(&rec, sizeof(Rec) * n); cudaMalloc(&count, 4); cudaMemset(count, 0, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<n * 2, 128, 200 * 1024>>>(rec, count);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  int c; cudaMemcpy(&c, count, 4, cudaMemcpyDeviceToHost);
  std::vector<Rec> h(c);
  cudaMemcpy(h.data(), rec, sizeof(Rec) * c, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end(), [](const Rec& a, const Rec& b) { return a.t < b.t; });
  printf("%d records (expect %d)\n", c, n);
  for (int i = 0; i < c && i < 400; ++i) printf("%d%s", h[i].id, (i % 20 == 19) ? "\n" : " ");
  printf("\n");
  // monotonicity after the first wave
  int inv = 0; for (int i = 75; i + 1 < c; ++i) inv += h[i + 1].id < h[i].id;
  printf("inversions after first wave: %d of %d\n", inv, c - 76);
  return 0;
}
