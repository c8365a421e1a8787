// The three K-pair contractions of the logit block (trainers/rpo.py:221-227 and their backward) on tcgen05:
//   forward   pair[k][b][c]      = dtype( img_s[b,k,:] . text_n[c,k,:] )            M = C, N = B, Kd = E
//   backward  d_img_s[b,k,:]     = dtype( sum_c dl[b,c] text_n[c,k,:] )            M = B, N = E, Kd = C
//             d_text_n[c,k,:]    = dtype( sum_b dl[b,c] img_s[b,k,:] )             M = C, N = E, Kd = B
// One batched launch each (grid.y = pair k).  Forward: both operands are 16-byte aligned K-major feature matrices and
// both go through the TMA ring (plain SS MMA).  Backward: the shapes are small and odd (B = 16..64 rows, dl rows of
// 2 C = 200 bytes), so the operand that has them goes the way the attention kernel sends its probabilities: the 128
// threads that own the 128 rows (= TMEM lanes) of the tile read it from global memory with whatever strides it has,
// and write it into TENSOR MEMORY as the A operand (tcgen05.st, two 16-bit values per column; 128 contraction
// elements per segment, two segments in flight).  The other operand is always one of the two big, aligned feature
// matrices ([rows, K * E], 16-byte aligned rows): it is streamed by TMA into a 4-stage ring, K-major ([N rows, 64
// contraction columns], forward) or MN-major ([64 contraction rows, 64 N columns] boxes, consumed as loaded like V
// in the attention kernel, backward).  One thread issues tcgen05.mma (A in TMEM, K = 16 per instruction), f32
// accumulation in TMEM; the row threads read the accumulator back, round to the dtype and store -- along the row
// when the output is row-contiguous, across the lanes when it is the transposed forward output.
// Rows of the tile past M are never loaded and never stored (a row of garbage only ever meets its own output row);
// contraction elements past Kd are zero on both sides (explicitly in A, by TMA's out-of-bounds fill in B).
#include "common.cuh"
#include "tc_common.cuh"

namespace rpo {
namespace ltc {

using namespace tc;

static constexpr int THREADS = 192;      // warps 0..3: rows (A operand, epilogue); warp 4: TMA; warp 5: MMA issue
static constexpr int STAGES = 4;
static constexpr int B_STAGE_BYTES = 64 * 128;  // K-major: <= 64 rows x 128 B; MN-major: 64 contraction rows x 128 B
static constexpr int A_STAGE_BYTES = 128 * 128;  // forward only: the A tile through TMA as well (128 rows x 64 columns)
static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
static constexpr int SEG = 128;          // contraction elements of A per TMEM buffer (64 columns)
static constexpr int KB_PER_SEG = SEG / 64;
static constexpr int A_COLS = SEG / 2;
static constexpr int D_COL = 2 * A_COLS;
static constexpr int TMEM_COLS = 256;    // two A buffers + 64 accumulator columns; two CTAs share an SM
static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;

struct Args {
  const void *a;               // A(m, kd, batch) = a[m * a_m + kd * a_k + batch * a_b]
  long long a_m, a_k, a_b;
  int a_vec;                   // 1: rows are contiguous (a_k == 1) and every row start is 16-byte aligned
  void *out;                   // D(m, n, batch) -> out[m * o_m + n * o_n + batch * o_b]
  long long o_m, o_n, o_b;
  int M, N, Kd;                // problem size per batch; N <= 64 per tile
  int n_tiles;                 // tiles along N (64 wide) -- 1 for the forward
  int b_col_batch;             // TMA column coordinate of B: batch * b_col_batch + (MN-major: n0; K-major: kd0)
  int a_col_batch;             // A through TMA (forward): column coordinate batch * a_col_batch + kd0, row m0
  int nb;                      // UMMA N (multiple of 16, <= 64)
};

// A_TMA (forward): both operands are aligned K-major feature matrices, so A takes the TMA ring too (plain SS MMA) and
// the row threads only run the epilogue
template <typename T, bool B_MN, bool A_TMA>
__global__ void __launch_bounds__(THREADS, 2)
    pair_gemm_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a, Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t ring = smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_FULL = 0, B_EMPTY = STAGES, A_FULL = 2 * STAGES, A_FREE = 2 * STAGES + 2, D_FULL = 2 * STAGES + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + D_FULL + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int batch = blockIdx.y;
  const int mt = blockIdx.x / p.n_tiles, nt = blockIdx.x - mt * p.n_tiles;
  const int m0 = mt * 128, n0 = nt * 64;
  const int num_kb = (p.Kd + 63) >> 6;
  const int num_seg = (p.Kd + SEG - 1) / SEG;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(BAR(B_FULL + i), 1);
      mbar_init(BAR(B_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(A_FULL + i), 4);
      mbar_init(BAR(A_FREE + i), 1);
    }
    mbar_init(BAR(D_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA producer: the B tile of every 64-wide contraction block =====
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(BAR(B_EMPTY + s), (uint32_t)(((kb / STAGES) - 1) & 1));
        const uint32_t dst = ring + s * STAGE_BYTES + A_STAGE_BYTES;
        if (B_MN) {
          mbar_arrive_expect_tx(BAR(B_FULL + s), 64 * 128);
          tma_load_2d(dst, &map_b, BAR(B_FULL + s), batch * p.b_col_batch + n0, kb * 64);
        } else {
          mbar_arrive_expect_tx(BAR(B_FULL + s), (uint32_t)(p.nb * 128) + (A_TMA ? (uint32_t)A_STAGE_BYTES : 0u));
          tma_load_2d(dst, &map_b, BAR(B_FULL + s), batch * p.b_col_batch + kb * 64, 0);
          if (A_TMA) tma_load_2d(ring + s * STAGE_BYTES, &map_a, BAR(B_FULL + s), batch * p.a_col_batch + kb * 64, m0);
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      // ===== MMA issuer =====
      const uint32_t fmt = Num<T>::dtype == RPO_BF16 ? 1u : 0u;
      const uint32_t idesc = make_idesc((int)fmt, 128, p.nb) | (B_MN ? (1u << 16) : 0u);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int seg = kb / KB_PER_SEG, buf = seg & 1, s = kb % STAGES;
        if (!A_TMA && kb % KB_PER_SEG == 0) mbar_wait(BAR(A_FULL + buf), (uint32_t)((seg >> 1) & 1));
        mbar_wait(BAR(B_FULL + s), (uint32_t)((kb / STAGES) & 1));
        tc_fence_after();
        const uint64_t bdesc = make_smem_desc(ring + s * STAGE_BYTES + A_STAGE_BYTES);
        if (A_TMA) {
          const uint64_t adesc = make_smem_desc(ring + s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + D_COL, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
        } else {
          const uint32_t acol = tmem_base + (uint32_t)(buf * A_COLS + (kb % KB_PER_SEG) * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 contraction elements: 8 columns of A; K-major B: 32 bytes along the row, MN-major: 16 rows
            umma_f16_ts(tmem_base + D_COL, acol + (uint32_t)(k * 8), bdesc + (uint64_t)(B_MN ? k * 128 : k * 2), idesc,
                        (kb | k) != 0);
        }
        umma_commit(BAR(B_EMPTY + s));
        if (!A_TMA && (kb % KB_PER_SEG == KB_PER_SEG - 1 || kb == num_kb - 1)) umma_commit(BAR(A_FREE + buf));
      }
      umma_commit(BAR(D_FULL));
    }
  } else if (warp < 4) {
    // ===== row threads: A operand into tensor memory, then the epilogue =====
    const int row = warp * 32 + lane;
    const int m = m0 + row;
    const bool valid = m < p.M;
    const bool warp_valid = m0 + warp * 32 < p.M;
    const T *arow = reinterpret_cast<const T *>(p.a) + (long long)m * p.a_m + (long long)batch * p.a_b;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int seg = 0; seg < (A_TMA ? 0 : num_seg); ++seg) {
      const int buf = seg & 1;
      if (seg >= 2) mbar_wait(BAR(A_FREE + buf), (uint32_t)(((seg >> 1) - 1) & 1));
      if (warp_valid) {
        tc_fence_after();
        const int kd0 = seg * SEG;
        const int kd_end = min(p.Kd, kd0 + SEG);
        const int steps = (min(num_kb * 64, kd0 + SEG) - kd0) >> 4;  // 16 contraction elements per step, whole 64-blocks
        // all of the segment's loads first (they are independent: one memory latency per segment, not per step), then
        // the tensor-memory stores
        uint32_t pk[SEG / 16][8];
#pragma unroll
        for (int st = 0; st < SEG / 16; ++st) {
          const int kd = kd0 + st * 16;
          if (st < steps) {
            if (valid && p.a_vec && kd + 16 <= kd_end) {
              const uint4 v0 = *reinterpret_cast<const uint4 *>(arow + kd);
              const uint4 v1 = *reinterpret_cast<const uint4 *>(arow + kd + 8);
              pk[st][0] = v0.x, pk[st][1] = v0.y, pk[st][2] = v0.z, pk[st][3] = v0.w;
              pk[st][4] = v1.x, pk[st][5] = v1.y, pk[st][6] = v1.z, pk[st][7] = v1.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                unsigned short lo = 0, hi = 0;
                if (valid && kd + 2 * e < kd_end)
                  lo = *reinterpret_cast<const unsigned short *>(arow + (long long)(kd + 2 * e) * p.a_k);
                if (valid && kd + 2 * e + 1 < kd_end)
                  hi = *reinterpret_cast<const unsigned short *>(arow + (long long)(kd + 2 * e + 1) * p.a_k);
                pk[st][e] = (uint32_t)lo | ((uint32_t)hi << 16);
              }
            }
          }
        }
#pragma unroll
        for (int st = 0; st < SEG / 16; ++st)
          if (st < steps) tmem_st8(lane_base + (uint32_t)(buf * A_COLS + st * 8), pk[st]);
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(A_FULL + buf));
    }
    // ---- epilogue ----
    mbar_wait(BAR(D_FULL), 0);
    tc_fence_after();
    if (warp_valid) {
      T *obase = reinterpret_cast<T *>(p.out) + (long long)batch * p.o_b + (long long)m * p.o_m;
      const int ncols = min(64, p.N - n0);
#pragma unroll 1
      for (int c = 0; c < p.nb; c += 16) {
        uint32_t acc[16];
        tmem_ld16(lane_base + (uint32_t)(D_COL + c), acc);
        if (!valid) continue;
        if (p.o_n == 1) {
          // row-contiguous output: 16 values = 32 bytes per step
          T *dst = obase + n0 + c;
          if (c + 16 <= ncols && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const T lo = fromf<T>(__uint_as_float(acc[2 * e])), hi = fromf<T>(__uint_as_float(acc[2 * e + 1]));
              pk[e] = (uint32_t)(*reinterpret_cast<const unsigned short *>(&lo)) |
                      ((uint32_t)(*reinterpret_cast<const unsigned short *>(&hi)) << 16);
            }
            *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c + e < ncols) dst[e] = fromf<T>(__uint_as_float(acc[e]));
          }
        } else {
          // transposed output (consecutive lanes = consecutive addresses)
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c + e < ncols) obase[(long long)(n0 + c + e) * p.o_n] = fromf<T>(__uint_as_float(acc[e]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// 2D row-major [rows, cols] 16-bit tensor, box = [box_rows, 64 cols], 128B swizzle (tc::make_map with the row count free)
template <typename T, bool B_MN, bool A_TMA>
static int launch(const CUtensorMap &map_b, const CUtensorMap &map_a, const Args &p, int batches, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(pair_gemm_kernel<T, B_MN, A_TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)(((p.M + 127) / 128) * p.n_tiles), (unsigned)batches);
  pair_gemm_kernel<T, B_MN, A_TMA><<<grid, THREADS, SMEM_BYTES, st>>>(map_b, map_a, p);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

}  // namespace ltc

bool logits_tc_supported(int dtype, int B, int C, int K, int E) {
  if (dtype != RPO_F16 && dtype != RPO_BF16) return false;
  if (E % 64 != 0 || B < 1 || B > 64 || C < 1 || K < 1) return false;
  if ((long long)K * E >= (1LL << 31) || (long long)C * K >= (1LL << 31)) return false;
  return true;
}

// pair[k][b][c] = dtype( img_s[b,k,:] . text_n[c,k,:] )
template <typename T>
int logits_pair_fwd_tc(const T *img_s, const T *text_n, T *pair, int B, int C, int K, int E, cudaStream_t st) {
  using namespace ltc;
  RPO_REQUIRE(logits_tc_supported(Num<T>::dtype, B, C, K, E), "logit block shape for the tcgen05 path");
  RPO_REQUIRE((((uintptr_t)img_s | (uintptr_t)text_n) & 15) == 0, "feature matrices must be 16-byte aligned");
  const int nb = (B + 15) & ~15;
  CUtensorMap map_b, map_a;  // B operand = img_s [B rows, K*E], K-major boxes of nb rows x 64 columns; A = text_n, 128-row boxes
  RPO_TRY(tc::make_map(&map_b, Num<T>::dtype, img_s, B, K * E, (long long)K * E, nb));
  RPO_TRY(tc::make_map(&map_a, Num<T>::dtype, text_n, C, K * E, (long long)K * E, 128));
  Args p{};
  p.a = text_n, p.a_m = (long long)K * E, p.a_k = 1, p.a_b = E, p.a_vec = 1;
  p.out = pair, p.o_m = 1, p.o_n = C, p.o_b = (long long)B * C;
  p.M = C, p.N = B, p.Kd = E, p.n_tiles = 1, p.b_col_batch = E, p.a_col_batch = E, p.nb = nb;
  prof_tag("logits_pair_fwd_tc B=%d C=%d K=%d E=%d", B, C, K, E);
  return launch<T, false, true>(map_b, map_a, p, K, st);
}

// d_img_s[b,k,:] = dtype( sum_c dl[b,c] text_n[c,k,:] ),  d_text_n[c,k,:] = dtype( sum_b dl[b,c] img_s[b,k,:] )
template <typename T>
int logits_pair_bwd_tc(const T *dl, const T *img_s, const T *text_n, T *d_img_s, T *d_text_n, int B, int C, int K,
                       int E, cudaStream_t st) {
  using namespace ltc;
  RPO_REQUIRE(logits_tc_supported(Num<T>::dtype, B, C, K, E), "logit block shape for the tcgen05 path");
  RPO_REQUIRE((((uintptr_t)img_s | (uintptr_t)text_n | (uintptr_t)d_img_s | (uintptr_t)d_text_n) & 15) == 0,
              "feature matrices must be 16-byte aligned");
  {
    CUtensorMap map_b;  // B operand = text_n [C rows, K*E], MN-major boxes of 64 contraction rows x 64 columns
    RPO_TRY(tc::make_map(&map_b, Num<T>::dtype, text_n, C, K * E, (long long)K * E, 64));
    Args p{};
    p.a = dl, p.a_m = C, p.a_k = 1, p.a_b = 0, p.a_vec = (C % 8 == 0 && ((uintptr_t)dl & 15) == 0) ? 1 : 0;
    p.out = d_img_s, p.o_m = (long long)K * E, p.o_n = 1, p.o_b = E;
    p.M = B, p.N = E, p.Kd = C, p.n_tiles = E / 64, p.b_col_batch = E, p.nb = 64;
    prof_tag("logits_pair_dimg_tc B=%d C=%d K=%d E=%d", B, C, K, E);
    RPO_TRY((launch<T, true, false>(map_b, map_b, p, K, st)));
  }
  {
    CUtensorMap map_b;  // B operand = img_s [B rows, K*E], MN-major
    RPO_TRY(tc::make_map(&map_b, Num<T>::dtype, img_s, B, K * E, (long long)K * E, 64));
    Args p{};
    p.a = dl, p.a_m = 1, p.a_k = C, p.a_b = 0, p.a_vec = 0;
    p.out = d_text_n, p.o_m = (long long)K * E, p.o_n = 1, p.o_b = E;
    p.M = C, p.N = E, p.Kd = B, p.n_tiles = E / 64, p.b_col_batch = E, p.nb = 64;
    prof_tag("logits_pair_dtext_tc B=%d C=%d K=%d E=%d", B, C, K, E);
    RPO_TRY((launch<T, true, false>(map_b, map_b, p, K, st)));
  }
  return RPO_OK;
}

template int logits_pair_fwd_tc<__half>(const __half *, const __half *, __half *, int, int, int, int, cudaStream_t);
template int logits_pair_fwd_tc<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, __nv_bfloat16 *, int, int,
                                               int, int, cudaStream_t);
template int logits_pair_bwd_tc<__half>(const __half *, const __half *, const __half *, __half *, __half *, int, int, int,
                                        int, cudaStream_t);
template int logits_pair_bwd_tc<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, const __nv_bfloat16 *,
                                               __nv_bfloat16 *, __nv_bfloat16 *, int, int, int, int, cudaStream_t);

}  // namespace rpo
