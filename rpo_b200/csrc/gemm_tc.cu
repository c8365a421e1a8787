// tcgen05 GEMM for the dense 16-bit contractions of the towers (QKV / out-proj / MLP / patch-embed /
// projections and their input-gradient counterparts):
//     C[M,N] = epilogue( A[M,Kd] . B[N,Kd]^T )        A, B K-major (row-major, K contiguous)
//
// Blackwell-native structure (one 128 x BN output tile per CTA, 2 CTAs resident per SM so one
// CTA's epilogue overlaps the other's main loop):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor 2D tiles (128B swizzle) into a 3-4 stage
//                 shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1      : MMA issuer    -- one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                 (M=128, N=BN, K=16) with the f32 accumulator in TMEM; tcgen05.commit releases
//                 ring slots and finally signals the epilogue
//   warps 2..5  : epilogue      -- tcgen05.ld (32 lanes x 32 columns per warp), fused bias /
//                 QuickGELU / gelu-grad / residual, 16-byte stores
// M tails are handled by TMA out-of-bounds zero fill on load and row predicates on store.
#include <cuda.h>

#include "common.cuh"

namespace rpo {

namespace tc {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 x 2 B = 128 B = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int THREADS = 192;

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN >= 128) ? 3 : 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as
// a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format f32 (1<<4), a/b format (0 f16, 1 bf16) at
// bits 7 and 10, both operands K-major, N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int fmt16, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt16 << 7) | ((uint32_t)fmt16 << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <typename T, int BN>
__global__ void __launch_bounds__(THREADS, 2)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   T *__restrict__ C, long long ldc, long long M, int N, int Kd, Epilogue<T> ep) {
  using C_ = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C_::STAGES * C_::STAGE_BYTES);
  // bars[0..S) full, bars[S..2S) empty, bars[2S] accumulator-ready; then the TMEM base address slot
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * C_::STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int num_kb = Kd / BK;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C_::STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * C_::STAGES);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    for (int s = 0; s < C_::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), C_::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        int s = kb % C_::STAGES;
        uint32_t ph = (kb / C_::STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);  // fresh barrier: parity-1 wait passes immediately
        mbar_arrive_expect_tx(full_bar(s), C_::STAGE_BYTES);
        uint32_t a_dst = smem_base + s * C_::STAGE_BYTES;
        tma_load_2d(a_dst, &map_a, full_bar(s), kb * BK, (int)m0);
        tma_load_2d(a_dst + C_::A_BYTES, &map_b, full_bar(s), kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(Num<T>::dtype == RPO_BF16 ? 1 : 0, BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        int s = kb % C_::STAGES;
        uint32_t ph = (kb / C_::STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        uint32_t a_addr = smem_base + s * C_::STAGE_BYTES;
        uint64_t adesc = make_smem_desc(a_addr);
        uint64_t bdesc = make_smem_desc(a_addr + C_::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in 16-byte units
          umma_f16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
        }
        umma_commit(empty_bar(s));  // frees the ring slot once these MMAs have read it
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const long long m = m0 + q * 32 + lane;
    const bool row_ok = m < M;
    constexpr int VEC = 8;  // 16-bit elements per 16-byte vector
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 32 / VEC; ++g) {
          const int n = n0 + c0 + g * VEC;
          const long long off = m * ldc + n;
          Vec16<T> res, aux, out, pre, bias;
          if (ep.bias) bias = ld16(ep.bias + n);
          if (ep.residual) res = ld16(ep.residual + off);
          if (ep.gelu_grad_aux) aux = ld16(ep.gelu_grad_aux + off);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            float v = __uint_as_float(acc[g * VEC + e]);
            if (ep.bias) v += tof<T>(bias.v[e]);
            v = rnd<T>(v);
            pre.v[e] = fromf<T>(v);
            if (ep.act == RPO_ACT_QUICKGELU) v = rnd<T>(quickgelu_rounded<T>(v));
            if (ep.gelu_grad_aux) v = rnd<T>(v * quickgelu_grad(tof<T>(aux.v[e])));
            if (ep.residual) v += tof<T>(res.v[e]);
            out.v[e] = fromf<T>(v);
          }
          if (ep.aux_out && m >= ep.aux_row0) st16(ep.aux_out + (m - ep.aux_row0) * ldc + n, pre);
          st16(C + off, out);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C_::TMEM_COLS);
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D row-major [rows, cols] 16-bit tensor with row stride ld (elements); box = [box_rows, 64 cols]
static int make_map(CUtensorMap *map, int dtype, const void *ptr, long long rows, int cols, long long ld,
                    int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return RPO_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == RPO_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return RPO_ERR_CUDA;
  }
  return RPO_OK;
}

template <typename T, int BN>
static int launch(const T *A, long long lda, const T *B, long long ldb, T *C, long long ldc, long long M, int N,
                  int Kd, const Epilogue<T> &ep, cudaStream_t st) {
  using C_ = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C_::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap map_a, map_b;
  RPO_TRY(make_map(&map_a, Num<T>::dtype, A, M, Kd, lda, BM));
  RPO_TRY(make_map(&map_b, Num<T>::dtype, B, N, Kd, ldb, BN));
  dim3 grid(N / BN, (unsigned)((M + BM - 1) / BM));
  gemm_tc_kernel<T, BN><<<grid, THREADS, C_::SMEM_BYTES, st>>>(map_a, map_b, C, ldc, M, N, Kd, ep);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

}  // namespace tc

bool gemm_tcgen05_supported(int dtype, long long lda, long long ldb, long long ldc, long long M, int N, int Kd,
                            const void *A, const void *B, const void *C) {
  if (dtype != RPO_F16 && dtype != RPO_BF16) return false;
  if (M <= 0 || N <= 0 || Kd <= 0) return false;
  if (Kd % tc::BK != 0 || N % 32 != 0) return false;
  if (lda % 8 != 0 || ldb % 8 != 0 || ldc % 8 != 0) return false;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return false;
  if ((M + tc::BM - 1) / tc::BM > 65535) return false;
  return true;
}

template <typename T>
int gemm_tcgen05(const T *A, long long lda, const T *B, long long ldb, T *C, long long ldc, long long M, int N, int Kd,
                 const Epilogue<T> &ep, cudaStream_t st) {
  if constexpr (sizeof(T) != 2) {
    set_error("tcgen05 GEMM supports f16/bf16 only");
    return RPO_ERR_INVALID;
  } else {
    RPO_REQUIRE(gemm_tcgen05_supported(Num<T>::dtype, lda, ldb, ldc, M, N, Kd, A, B, C), "tcgen05 GEMM shape");
    if (ep.bias) RPO_REQUIRE(((uintptr_t)ep.bias & 15) == 0, "bias must be 16-byte aligned");
    if (ep.residual) RPO_REQUIRE(((uintptr_t)ep.residual & 15) == 0, "residual must be 16-byte aligned");
    if (ep.gelu_grad_aux) RPO_REQUIRE(((uintptr_t)ep.gelu_grad_aux & 15) == 0, "aux must be 16-byte aligned");
    if (ep.aux_out) RPO_REQUIRE(((uintptr_t)ep.aux_out & 15) == 0, "aux_out must be 16-byte aligned");
    // tile width: the widest BN that still yields at least ~one wave of CTAs on 148 SMs
    long long mt = (M + tc::BM - 1) / tc::BM;
    const int target = 148;
    if (N % 128 == 0 && mt * (N / 128) >= target) return tc::launch<T, 128>(A, lda, B, ldb, C, ldc, M, N, Kd, ep, st);
    if (N % 64 == 0 && mt * (N / 64) >= target) return tc::launch<T, 64>(A, lda, B, ldb, C, ldc, M, N, Kd, ep, st);
    return tc::launch<T, 32>(A, lda, B, ldb, C, ldc, M, N, Kd, ep, st);
  }
}

template int gemm_tcgen05<float>(const float *, long long, const float *, long long, float *, long long, long long,
                                 int, int, const Epilogue<float> &, cudaStream_t);
template int gemm_tcgen05<__half>(const __half *, long long, const __half *, long long, __half *, long long,
                                  long long, int, int, const Epilogue<__half> &, cudaStream_t);
template int gemm_tcgen05<__nv_bfloat16>(const __nv_bfloat16 *, long long, const __nv_bfloat16 *, long long,
                                         __nv_bfloat16 *, long long, long long, int, int,
                                         const Epilogue<__nv_bfloat16> &, cudaStream_t);

}  // namespace rpo
