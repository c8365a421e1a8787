// Micro-benchmark (tuning aid): how fast can one SM pull 128B-swizzled 2-D tiles through TMA, as a function of the box
// height (bytes per cp.async.bulk.tensor op) and of the number of issuing threads?  One CTA per SM; every issuing thread
// keeps DEPTH loads in flight on its own mbarriers.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_probe tools/probe/tma_probe.cu -lcuda && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}



// ISSUERS threads (lane 0 of warps 0..ISSUERS-1) each stream `iters` boxes of `rows` x 128 bytes
__global__ void probe(const __grid_constant__ CUtensorMap map, int rows, int iters, int total_rows, int kcols,
                      long long *out, int DEPTH) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int issuers = blockDim.x >> 5;
  uint64_t *bars = (uint64_t *)(smem + issuers * DEPTH * rows * 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < issuers * DEPTH; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) {
    const uint32_t base = smem_u32(smem) + warp * DEPTH * rows * 128;
    const uint32_t bar0 = smem_u32(bars + warp * DEPTH);
    const int bytes = rows * 128;
    // each (CTA, issuer) walks its own region of the matrix (cheap wrap-around arithmetic: no division in the loop)
    int r = (int)((((long long)blockIdx.x * issuers + warp) * 977 * rows) % (total_rows - rows));
    int kc = (blockIdx.x * 7 + warp) % kcols;
    long long t0 = clock64();
    for (int i = 0; i < iters + DEPTH; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) mbar_wait(bar0 + 8 * s, ((i / DEPTH) - 1) & 1);
      if (i < iters) {
        mbar_expect(bar0 + 8 * s, bytes);
        tma_load_2d(base + s * bytes, &map, bar0 + 8 * s, kc * 64, r);
        r += rows;
        if (r > total_rows - rows) {
          r -= total_rows - rows;
          kc = kc + 1 < kcols ? kc + 1 : 0;
        }
      }
    }
    long long t1 = clock64();
    out[blockIdx.x * 8 + warp] = t1 - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  int sms = 0;
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long *out;
  CHECK(cudaMalloc(&out, 148 * 8 * sizeof(long long)));
  // two footprints: 16384 x 4096 fp16 = 128 MB (HBM) and 2048 x 1024 fp16 = 4 MB (L2 resident)
  for (int big = 1; big >= 0; --big) {
    const int total_rows = big ? 16384 : 2048, cols = big ? 4096 : 1024;
    void *buf;
    CHECK(cudaMalloc(&buf, (size_t)total_rows * cols * 2));
    CHECK(cudaMemset(buf, 0, (size_t)total_rows * cols * 2));
    printf("== footprint %d MB (%s)\n", (int)((size_t)total_rows * cols * 2 >> 20), big ? "HBM" : "L2 resident");
    printf("box rows | issuers | depth | cycles per op (1 issuer) | implied latency | bytes/cycle/SM | chip TB/s @1.9GHz\n");
    for (int rows : {64, 128, 256}) {
      CUtensorMap map;
      cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)total_rows};
      cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      for (int issuers : {1, 2}) {
        for (int depth : {2, 4, 8, 16}) {
          const int smem = issuers * depth * rows * 128 + 512 + 1024;
          if (smem > 227 * 1024) continue;
          CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          const int iters = 2000;
          for (int rep = 0; rep < 2; ++rep)
            probe<<<sms, issuers * 32, smem>>>(map, rows, iters, total_rows, cols / 64, out, depth);
          CHECK(cudaDeviceSynchronize());
          long long h[148 * 8];
          CHECK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
          double sum = 0;
          for (int b = 0; b < sms; ++b) sum += (double)h[b * 8];
          const double cyc = sum / sms / iters;
          const double bpc = issuers * rows * 128.0 / cyc;
          printf("%8d | %7d | %5d | %24.1f | %15.0f | %14.1f | %6.2f\n", rows, issuers, depth, cyc, cyc * depth, bpc,
                 bpc * sms * 1.9e9 / 1e12);
        }
      }
    }
    CHECK(cudaFree(buf));
  }
  return 0;
}
