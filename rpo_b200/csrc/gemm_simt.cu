// Generic-stride SIMT GEMM with exact f32 FMA accumulation:  C[m,n] = sum_k A(m,k) * B(n,k).
// Serves (a) the whole RPO_F32 precision mode, where tensor cores (tf32) would break the 1e-5
// parity bar, (b) the small odd-shaped contractions of the logit block (trainers/rpo.py:221-227,
// M = batch, strided per-pair slices) and (c) a cross-check for the tcgen05 kernel in the tests.
// The dense 16-bit GEMMs of the towers go through gemm_tc.cu.
#include "common.cuh"

namespace rpo {

static constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T *__restrict__ A, long long sam, long long sak,
                                                        const T *__restrict__ B, long long sbn, long long sbk,
                                                        T *__restrict__ C, long long ldc, long long M, int N, int Kd,
                                                        Epilogue<T> ep, long long bsa, long long bsb, long long bsc) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  A += (long long)blockIdx.z * bsa;
  B += (long long)blockIdx.z * bsb;
  C += (long long)blockIdx.z * bsc;
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * SBM;
  const int n0 = blockIdx.x * SBN;
  const int lr = tid >> 2;        // 0..63 : tile row loaded by this thread
  const int lk = (tid & 3) * 4;   // 0,4,8,12 : first k loaded by this thread
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // register double buffer: the global loads of slab k0 + SBK are in flight while slab k0 is multiplied
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int k = k0 + lk + e;
      long long m = m0 + lr;
      int n = n0 + lr;
      ra[e] = (m < M && k < Kd) ? tof<T>(A[m * sam + k * sak]) : 0.f;
      rb[e] = (n < N && k < Kd) ? tof<T>(B[n * sbn + k * sbk]) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < Kd; k0 += SBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[lk + e][lr] = ra[e];
      Bs[lk + e][lr] = rb[e];
    }
    __syncthreads();
    if (k0 + SBK < Kd) fetch(k0 + SBK);
    // blocked summation: the 16 products of a slab are summed on their own and then added to the running sum, so
    // the rounding error grows with K/16 + 16 terms instead of K (the fp32 parity bar is 1e-5 on gradients that
    // pass through a dozen K = 768 .. 3072 contractions)
    float part[4][4];
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = k == 0 ? av[i] * bv[j] : fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) C[m * ldc + n] = fromf<T>(ep.apply(acc[i][j], m, n, ldc));
    }
  }
}

template <typename T>
int gemm_simt(const T *A, long long sam, long long sak, const T *B, long long sbn, long long sbk, T *C, long long ldc,
              long long M, int N, int Kd, const Epilogue<T> &ep, int batch, long long bsa, long long bsb,
              long long bsc, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return RPO_OK;
  RPO_REQUIRE((M + SBM - 1) / SBM <= 65535 && batch <= 65535, "SIMT GEMM grid too large");
  dim3 grid((N + SBN - 1) / SBN, (unsigned)((M + SBM - 1) / SBM), batch);
  gemm_simt_kernel<T><<<grid, 256, 0, st>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, Kd, ep, bsa, bsb, bsc);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

template <typename T>
int gemm_dispatch(int backend, const T *A, long long lda, const T *B, long long ldb, T *C, long long ldc, long long M,
                  int N, int Kd, const Epilogue<T> &ep, cudaStream_t st) {
  if (M <= 0) return RPO_OK;
  prof_tag("gemm M=%lld N=%d K=%d%s%s%s%s", M, N, Kd, ep.bias ? " +bias" : "", ep.act ? " +gelu" : "",
           ep.residual ? " +res" : "", ep.gelu_grad_aux ? " *gelu'" : "");
  bool tc_ok = gemm_tcgen05_supported(Num<T>::dtype, lda, ldb, ldc, M, N, Kd, A, B, C);
  if (backend == RPO_GEMM_TCGEN05) {
    RPO_REQUIRE(tc_ok, "shape/dtype not supported by the tcgen05 GEMM");
    return gemm_tcgen05<T>(A, lda, B, ldb, C, ldc, M, N, Kd, ep, st);
  }
  if (backend == RPO_GEMM_AUTO && tc_ok) return gemm_tcgen05<T>(A, lda, B, ldb, C, ldc, M, N, Kd, ep, st);
  return gemm_simt<T>(A, lda, 1, B, ldb, 1, C, ldc, M, N, Kd, ep, 1, 0, 0, 0, st);
}

#define INSTANTIATE(T)                                                                                              \
  template int gemm_simt<T>(const T *, long long, long long, const T *, long long, long long, T *, long long,       \
                            long long, int, int, const Epilogue<T> &, int, long long, long long, long long,         \
                            cudaStream_t);                                                                          \
  template int gemm_dispatch<T>(int, const T *, long long, const T *, long long, T *, long long, long long, int,    \
                                int, const Epilogue<T> &, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(__half)
INSTANTIATE(__nv_bfloat16)

}  // namespace rpo
