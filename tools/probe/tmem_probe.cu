// Tensor-memory read bandwidth probe (sm_100a): W warps of one CTA per SM stream tcgen05.ld over the CTA's 512 columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
// Prints bytes per SM clock for 4 / 8 / 16 warps and the x16 / x32 / x64 load shapes, with and without a wait per load.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(a) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int SHAPE, int WAIT_EVERY>
__global__ void __launch_bounds__(1024, 1) probe(int iters, long long *cycles, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * SHAPE + (warp >> 2) * 64) & 255);
    if (SHAPE == 16) {
      uint32_t r[16];
      ld16(base + col, r);
      if (WAIT_EVERY && (i % WAIT_EVERY) == WAIT_EVERY - 1) ld_wait();
      acc ^= r[0] ^ r[15];
    } else {
      uint32_t r[32];
      ld32(base + col, r);
      if (WAIT_EVERY && (i % WAIT_EVERY) == WAIT_EVERY - 1) ld_wait();
      acc ^= r[0] ^ r[31];
    }
  }
  ld_wait();
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int SHAPE, int WAIT_EVERY>
void run(int warps, int iters, long long *d_cycles, uint32_t *d_sink) {
  probe<SHAPE, WAIT_EVERY><<<148, warps * 32>>>(iters, d_cycles, d_sink);
  cudaDeviceSynchronize();
  probe<SHAPE, WAIT_EVERY><<<148, warps * 32>>>(iters, d_cycles, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[148];
  cudaMemcpy(c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
  const double bytes = (double)warps * iters * 32.0 * SHAPE * 4.0;
  printf("x%-2d warps=%2d wait_every=%d : %8lld cycles for %8.0f KB per SM = %6.1f B/clk/SM  (%s)\n", SHAPE, warps, WAIT_EVERY,
         mx, bytes / 1024.0, bytes / (double)mx, cudaGetErrorString(e));
}

int main() {
  long long *d_cycles;
  uint32_t *d_sink;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {
    run<16, 0>(warps, iters, d_cycles, d_sink);
    run<16, 4>(warps, iters, d_cycles, d_sink);
    run<16, 1>(warps, iters, d_cycles, d_sink);
    run<32, 0>(warps, iters, d_cycles, d_sink);
    run<32, 1>(warps, iters, d_cycles, d_sink);
  }
  return 0;
}
