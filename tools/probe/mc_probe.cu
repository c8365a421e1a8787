// Micro-benchmark (tuning aid): does TMA multicast raise the operand bandwidth an SM can sustain when the whole chip
// streams GEMM operand tiles out of L2?  Clusters of 4 CTAs (one CTA per SM), every CTA receives 32 KB per step
// (a 128 x 64 "A" tile of its own + a 128 x 64 "B" tile), DEPTH steps in flight:
//   mode 0  unicast, every tile distinct            : 32 KB read from L2 per CTA and step
//   mode 1  unicast, CTA r and r^2 read the SAME B  : 24 KB unique per CTA (does L2 merge the duplicate requests?)
//   mode 2  multicast: CTA r loads half of the shared B tile (64 rows) and multicasts it to {r, r^2}
//   mode 3  multicast of B over all 4 CTAs (32 rows each): 20 KB unique per CTA
// Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mc_probe tools/probe/mc_probe.cu -lcuda && /tmp/mc_probe
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
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int TILE = 128 * 128;  // bytes of a 128-row x 64-column 16-bit tile

__global__ void __cluster_dims__(4, 1, 1) probe(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map64,
                                                const __grid_constant__ CUtensorMap map32, int mode, int iters, int total_rows,
                                                int kcols, long long *out, int DEPTH) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + DEPTH * 2 * TILE);
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int sharers = mode == 3 ? 4 : (mode == 2 ? 2 : 1);
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) {
      mbar_init(smem_u32(bars + i), 1);                   // full[i]
      mbar_init(smem_u32(bars + DEPTH + i), sharers);     // empty[i]: every CTA that writes into this slot's B half
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem), full0 = smem_u32(bars), empty0 = smem_u32(bars + DEPTH);
    const int cluster = blockIdx.x >> 2;
    // own A rows, and B rows shared at the granularity the mode asks for
    int ra = (int)(((long long)blockIdx.x * 977 * 128) % (total_rows - 128));
    const int owner = mode == 0 ? blockIdx.x : (mode == 3 ? cluster * 4 : cluster * 4 + (rank & 1));
    int rb = (int)(((long long)(owner * 2 + 1) * 1409 * 128) % (total_rows - 128));
    int kc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters + DEPTH; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) {
        mbar_wait(full0 + 8 * s, ((i / DEPTH) - 1) & 1);
        // slot consumed: tell every CTA that multicasts into it
        if (mode == 2) {
          arrive_remote(mapa(empty0 + 8 * s, rank));
          arrive_remote(mapa(empty0 + 8 * s, rank ^ 2));
        } else if (mode == 3) {
          for (uint32_t r = 0; r < 4; ++r) arrive_remote(mapa(empty0 + 8 * s, r));
        }
      }
      if (i < iters) {
        if (mode >= 2 && i >= DEPTH) mbar_wait(empty0 + 8 * s, ((i / DEPTH) - 1) & 1);
        mbar_expect(full0 + 8 * s, 2 * TILE);
        tma_load_2d(base + s * 2 * TILE, &map128, full0 + 8 * s, kc * 64, ra);
        const uint32_t bdst = base + s * 2 * TILE + TILE;
        if (mode <= 1) {
          tma_load_2d(bdst, &map128, full0 + 8 * s, kc * 64, rb);
        } else if (mode == 2) {
          const uint32_t part = rank >> 1;  // which half of the shared tile this CTA fetches
          tma_load_2d_mc(bdst + part * (TILE / 2), &map64, full0 + 8 * s, kc * 64, rb + part * 64,
                         (uint16_t)((1u << rank) | (1u << (rank ^ 2))));
        } else {
          tma_load_2d_mc(bdst + rank * (TILE / 4), &map32, full0 + 8 * s, kc * 64, rb + rank * 32, (uint16_t)0xF);
        }
        ra += 128;
        rb += 128;
        if (ra > total_rows - 128) ra -= total_rows - 128;
        if (rb > total_rows - 128) rb -= total_rows - 128;
        kc = kc + 1 < kcols ? kc + 1 : 0;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
  CHECK(cudaMalloc(&out, 148 * sizeof(long long)));
  const int total_rows = 8192, cols = 1024;  // 16 MB: L2 resident, like the operands of one GEMM
  void *buf;
  CHECK(cudaMalloc(&buf, (size_t)total_rows * cols * 2));
  CHECK(cudaMemset(buf, 0, (size_t)total_rows * cols * 2));
  CUtensorMap maps[3];
  int boxrows[3] = {128, 64, 32};
  for (int m = 0; m < 3; ++m) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)boxrows[m]};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  const int grid = (sms / 4) * 4;
  printf("grid %d CTAs (clusters of 4), 32 KB received per CTA and step\n", grid);
  printf("mode | depth | cycles per step | bytes/cycle/SM received | chip TB/s received @1.9GHz\n");
  for (int mode = 0; mode < 4; ++mode) {
    for (int depth : {3, 6}) {
      const int smem = depth * 2 * TILE + 256 + 1024;
      CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int iters = 3000;
      for (int rep = 0; rep < 2; ++rep)
        probe<<<grid, 32, smem>>>(maps[0], maps[1], maps[2], mode, iters, total_rows, cols / 64, out, depth);
      CHECK(cudaDeviceSynchronize());
      long long h[148];
      CHECK(cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
      double sum = 0;
      for (int b = 0; b < grid; ++b) sum += (double)h[b];
      const double cyc = sum / grid / iters;
      printf("%4d | %5d | %15.1f | %23.1f | %6.2f\n", mode, depth, cyc, 2.0 * TILE / cyc, 2.0 * TILE / cyc * grid * 1.9e9 / 1e12);
    }
  }
  return 0;
}
