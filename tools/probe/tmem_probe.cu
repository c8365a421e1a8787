// Tensor-memory read / write throughput probe (sm_100a):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
// One CTA; W warps (warp w reads the lanes of quarter w % 4) issue R loads of 16 / 32 / 64 columns with D of them in
// flight per tcgen05.wait::ld.  Prints SM clocks per load and bytes per clock for the CTA.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void st8(uint32_t a, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int DEPTH, bool STORE>
__global__ void probe(int reps, long long *out, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 448);
  uint32_t r[DEPTH][16];
  uint32_t acc = 0;
#pragma unroll
  for (int d = 0; d < DEPTH; ++d)
#pragma unroll
    for (int e = 0; e < 16; ++e) r[d][e] = threadIdx.x + e;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    if (STORE) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) pk[e] = r[d][e] + i;
        st8(base + d * 8, pk);
      }
      wait_st();
    } else {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) ld16(base + (d * 16) % 64, r[d]);
      wait_ld();
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc += r[d][0] ^ r[d][15];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int DEPTH, bool STORE>
void run(int warps, int reps) {
  long long *d;
  uint32_t *sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  probe<DEPTH, STORE><<<1, warps * 32>>>(reps, d, sink);
  probe<DEPTH, STORE><<<1, warps * 32>>>(reps, d, sink);
  long long clk = 0;
  cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double ops = (double)reps * DEPTH;
  const double bytes = ops * warps * 32 * (STORE ? 8 : 16) * 4;
  printf("%s warps=%2d depth=%d: %7.1f clk per %s per warp, %7.1f B/clk per SM  (%s)\n", STORE ? "st.x8 " : "ld.x16", warps, DEPTH,
         clk / ops, STORE ? "store" : "load", bytes / clk, cudaGetErrorString(e));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  const int reps = 2000;
  for (int w : {1, 4, 8, 12, 16}) {
    run<1, false>(w, reps);
    run<2, false>(w, reps);
    run<4, false>(w, reps);
    run<7, false>(w, reps);
  }
  for (int w : {1, 4, 8}) {
    run<1, true>(w, reps);
    run<4, true>(w, reps);
  }
  return 0;
}
