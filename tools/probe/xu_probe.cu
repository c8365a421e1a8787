// Issue-rate probe for the softmax inner loop (sm_100a): clocks per warp instruction on one scheduler for
// ex2.approx, cvt.rn.f16x2.f32 (F2FP), their 2:1 mix, fma.rn.f32x2, ex2.approx.f16x2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_probe xu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void probe(int reps, long long *out, float *sink) {
  float x[16];
  uint32_t pk[8];
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = -0.001f * (threadIdx.x + e);
#pragma unroll
  for (int e = 0; e < 8; ++e) pk[e] = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    if (MODE == 0) {  // 16 ex2
#pragma unroll
      for (int e = 0; e < 16; ++e) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[e]));
    } else if (MODE == 1) {  // 16 F2FP
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[e]) : "f"(x[2 * e]), "f"(x[2 * e + 1]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[e]) : "f"(x[2 * e + 1]), "f"(x[2 * e]));
      }
    } else if (MODE == 2) {  // 16 ex2 + 8 F2FP
#pragma unroll
      for (int e = 0; e < 16; ++e) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[e]));
#pragma unroll
      for (int e = 0; e < 8; ++e) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[e]) : "f"(x[2 * e]), "f"(x[2 * e + 1]));
    } else if (MODE == 3) {  // 16 fma.f32x2
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; fma.rn.f32x2 a, a, b, b; mov.b64 {%0, %1}, a;}"
                     : "+f"(x[2 * e]), "+f"(x[2 * e + 1]) : "f"(0.999f));
        asm volatile("{.reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; fma.rn.f32x2 a, a, b, b; mov.b64 {%0, %1}, a;}"
                     : "+f"(x[2 * e]), "+f"(x[2 * e + 1]) : "f"(0.998f));
      }
    } else if (MODE == 4) {  // 16 ex2.f16x2
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(pk[e]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(pk[e]));
      }
    } else if (MODE == 5) {  // 16 ex2 + 16 ffma + 16 fadd + 8 F2FP (the loop body)
      float l = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float y = fmaf(x[e], 0.18f, -1.0f);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y));
        l += y;
        x[e] = y;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[e]) : "f"(x[2 * e]), "f"(x[2 * e + 1]));
      x[0] += l;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  float acc = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) acc += x[e];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc += __uint_as_float(pk[e]);
  if (acc == 1.2345f) sink[0] = acc;
}

template <int MODE>
void run(const char *what, int warps_per_sched) {
  long long *d;
  float *sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  const int reps = 4000;
  probe<MODE><<<1, 128 * warps_per_sched>>>(reps, d, sink);
  probe<MODE><<<1, 128 * warps_per_sched>>>(reps, d, sink);
  long long clk = 0;
  cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/scheduler=%d: %6.2f clk per 16-instruction group per scheduler\n", what, warps_per_sched,
         (double)clk / reps / 1.0);
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<0>("16 x ex2.approx.ftz.f32", w);
    run<1>("16 x cvt.rn.f16x2.f32", w);
    run<2>("16 x ex2 + 8 x cvt.f16x2", w);
    run<3>("16 x fma.rn.f32x2", w);
    run<4>("16 x ex2.approx.f16x2", w);
    run<5>("16 x (ffma, ex2, fadd) + 8 x cvt.f16x2", w);
  }
  return 0;
}
