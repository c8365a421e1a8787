// tcgen05 form of the read-only masked attention forward for the vision tower (every group has the same
// number n of context rows, no causal mask): clip/model.py:186 under visual_mask of trainers/rpo.py:153-159.
//
// One CTA = one (image, head, 128-query tile).  The query tile holds context rows and -- in the last
// tile -- the K prompt rows right behind them; keys / values are the n context rows only (the mask).
//   warp 0, one thread : TMA (cp.async.bulk.tensor, 128B swizzle) of the Q tile, K and V of the head
//                        straight out of the [rows, 3D] q|k|v matrix and the [G*K, D] prompt-q matrix;
//                        S = Q K^T   as 4 x tcgen05.mma (M=128, N=n16, K=16), f32 accumulator in TMEM;
//                        O = P V     as n16/16 x tcgen05.mma (M=128, N=64, K=16): A = P from shared
//                        memory (K-major, 32B swizzle, one 128x16 block per MMA), B = V as loaded
//                        (keys x head-dim rows = MN-major operand, no transpose pass), accumulator
//                        over the TMEM columns of the consumed S.
//   warps 1..8         : softmax, TWO threads per query row (TMEM lane; warps w and w+4 share a lane
//                        quarter and split the keys 7 : 6 in 16-key blocks): pass 1 reads S for the row
//                        maximum (halves exchanged through shared memory), pass 2 re-reads it,
//                        p = ex2((s - max) / 8 log2 e), accumulates the f32 row sum, rounds p to the
//                        dtype and stores it as the A operand of P V.  TMEM loads are software
//                        pipelined (block b+1 in flight while block b is processed).
//                        epilogue: each thread takes 32 of the row's 64 output columns from TMEM,
//                        times 1/sum, 64 contiguous bytes to global.
// No shuffles and no ldmatrix: the tensor core reads operands from shared memory itself, so the SM's
// issue slots carry only the softmax (the mma.sync kernel in attention_mma.cu is issue-bound: ~3400
// instructions per 16 query rows).  Shared memory: K 26 KB + V 26 KB + max(Q 16 KB, P 52 KB) = 104 KB and
// 256 TMEM columns, so two CTAs share an SM and hide each other's load / MMA latency.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace rpo {

namespace atc {

using namespace tc;

static constexpr int HD = 64;
static constexpr int ROW_BYTES = HD * 2;
static constexpr int QT = 128;  // query rows per CTA = UMMA M
static constexpr int SM_WARPS = 8;  // softmax warps
static constexpr int THREADS = 32 + SM_WARPS * 32;
static constexpr int TMEM_COLS = 256;
static constexpr int P_BLOCK_BYTES = QT * 32;  // one 128 x 16 block of P, 32-byte rows

// K-major operand with 32-byte rows (16 x 16-bit = one UMMA K step), 32B swizzle, 8-row groups 256 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// The wait names the destination registers of the load it completes as read-write operands, so the
// compiler cannot move a use of them above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// packed f32x2 arithmetic (sm_100: FFMA2 / FADD2, two lanes per issue slot): (a, b) = (a, b) * s + o ; (a, b) += (c, d)
__device__ __forceinline__ void ffma2(float &a, float &b, float s, float o) {
  asm("{ .reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; mov.b64 z, {%3, %3}; fma.rn.f32x2 x, x, y, z; "
      "mov.b64 {%0, %1}, x; }"
      : "+f"(a), "+f"(b)
      : "f"(s), "f"(o));
}
__device__ __forceinline__ void fadd2(float &a, float &b, float c, float d) {
  asm("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
      : "+f"(a), "+f"(b)
      : "f"(c), "f"(d));
}
__device__ __forceinline__ void pair_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

struct Geo {
  int n;            // context rows (keys) per group
  int n16;          // keys rounded up to the UMMA K step
  int K;            // prompt rows per group
  int prompt_tile;  // query tile that holds the prompt rows ...
  int prompt_row;   // ... starting at this row of the tile (== context rows in that tile)
  int H;
  int tiles;        // query tiles per (group, head)
  int phase_delay;  // SM clocks the second CTA of an SM holds back its first loads (0: none), see the kernel
  int flags;        // bit 0: packed f32x2 arithmetic in the probability pass
};

template <typename T>
__global__ void __launch_bounds__(THREADS, 2)
    ro_attn_fwd_tc(const __grid_constant__ CUtensorMap map_full,   // q|k|v matrix, box 64 x 128
                   const __grid_constant__ CUtensorMap map_kvt,    // q|k|v matrix, box 64 x (n16 % 128)
                   const __grid_constant__ CUtensorMap map_qt,     // q|k|v matrix, box 64 x (n % 128)
                   const __grid_constant__ CUtensorMap map_prompt, // prompt-q matrix, box 64 x K
                   T *__restrict__ out_ctx, T *__restrict__ out_prompt, Geo geo, int num_items, long long *trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int n = geo.n, n16 = geo.n16, K = geo.K, H = geo.H;
  const int D = H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_bytes = n16 * ROW_BYTES;
  const int p_bytes = (n16 >> 4) * P_BLOCK_BYTES;
  const int qp_bytes = p_bytes > QT * ROW_BYTES ? p_bytes : QT * ROW_BYTES;
  const uint32_t Ks = smem_u32(smem), Vs = Ks + kv_bytes, QPs = Vs + kv_bytes;
  uint8_t *QP_gen = smem + 2 * kv_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * kv_bytes + qp_bytes);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 7);
  float *red_max = reinterpret_cast<float *>(bars + 8);  // [2][QT]
  float *red_sum = red_max + 2 * QT;                      // [2][QT]
  // one phase of every barrier per work item: K landed, Q landed, V landed, S ready, P stored, O ready, O read
  const uint32_t bar_k = smem_u32(bars), bar_q = bar_k + 8, bar_v = bar_k + 16, bar_s = bar_k + 24, bar_p = bar_k + 32,
                 bar_o = bar_k + 40, bar_e = bar_k + 48;
  const int tiles = geo.tiles;
  // phase timestamps of the first item of each CTA (tuning aid, RPO_ATTN_TRACE): [cta][8] SM clock values
  long long *tr = trace ? trace + (size_t)blockIdx.x * 8 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_full)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_kvt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_qt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_prompt)) : "memory");
    mbar_init(bar_k, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, SM_WARPS);
    mbar_init(bar_o, 1);
    mbar_init(bar_e, SM_WARPS);
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
  pdl_trigger();
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  // work items: (query tile, head, image), tile fastest so that the two tiles of a head run side by side
  // on neighbouring CTAs and share K/V through L2
  struct Item {
    int t, h, g, c_rows, p_rows;
  };
  auto item_of = [&](int id) {
    Item it;
    it.t = id % tiles;
    it.h = (id / tiles) % H;
    it.g = id / (tiles * H);
    it.c_rows = min(max(n - it.t * QT, 0), QT);       // context query rows of this tile
    it.p_rows = (it.t == geo.prompt_tile) ? K : 0;     // prompt query rows, right behind them
    return it;
  };

  if (warp == 0) {
    if (lane == 0) {
      auto load_k = [&](const Item &it) {
        mbar_arrive_expect_tx(bar_k, (uint32_t)kv_bytes);
        for (int r = 0; r < n16; r += 128)
          tma_load_2d(Ks + r * ROW_BYTES, (n16 - r >= 128) ? &map_full : &map_kvt, bar_k, D + it.h * HD, it.g * n + r);
      };
      auto load_qv = [&](const Item &it) {
        mbar_arrive_expect_tx(bar_q, (uint32_t)((it.c_rows + it.p_rows) * ROW_BYTES));
        if (it.c_rows == QT)
          tma_load_2d(QPs, &map_full, bar_q, it.h * HD, it.g * n + it.t * QT);
        else if (it.c_rows > 0)
          tma_load_2d(QPs, &map_qt, bar_q, it.h * HD, it.g * n + it.t * QT);
        if (it.p_rows > 0) tma_load_2d(QPs + geo.prompt_row * ROW_BYTES, &map_prompt, bar_q, it.h * HD, it.g * K);
        mbar_arrive_expect_tx(bar_v, (uint32_t)kv_bytes);
        for (int r = 0; r < n16; r += 128)
          tma_load_2d(Vs + r * ROW_BYTES, (n16 - r >= 128) ? &map_full : &map_kvt, bar_v, 2 * D + it.h * HD, it.g * n + r);
      };
      const uint32_t fmt = Num<T>::dtype == RPO_BF16 ? 1u : 0u;
      const uint32_t idesc_s = make_idesc((int)fmt, QT, n16);
      const uint32_t idesc_o = make_idesc((int)fmt, QT, HD) | (1u << 16);  // B (= V) is MN-major
      const int nsteps = n16 >> 4;
      pdl_wait();
      // Two CTAs share an SM and would otherwise run in lockstep -- both in the MUFU-bound probability pass at the
      // same time, both waiting on the tensor core at the same time.  The CTA that got the upper half of the SM's
      // tensor memory starts late by a fraction of an item so that one CTA's exponentials run beside the other's
      // MMA / row-maximum / epilogue phases.
      if (geo.phase_delay > 0 && (tmem_base & 0xFFFFu) >= (uint32_t)TMEM_COLS) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)geo.phase_delay) __nanosleep(100);
      }
      int id = blockIdx.x;
      if (id < num_items) {
        const Item first = item_of(id);
        load_k(first);
        load_qv(first);
      }
      for (uint32_t i = 0; id < num_items; id += gridDim.x, ++i) {
        const uint32_t par = i & 1;
        const int next = id + gridDim.x;
        // ---- S = Q K^T (the previous item's O has been read out of TMEM) ----
        mbar_wait(bar_k, par);
        mbar_wait(bar_q, par);
        if (i > 0) mbar_wait(bar_e, par ^ 1);
        tc_fence_after();
        if (tr && i == 0) tr[2] = clock64();
        {
          const uint64_t adesc = make_smem_desc(QPs), bdesc = make_smem_desc(Ks);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_f16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc_s, k != 0);
          umma_commit(bar_s);
        }
        // ---- K is free once S is complete: prefetch the next item's K behind this item's softmax ----
        if (next < num_items) {
          mbar_wait(bar_s, par);
          load_k(item_of(next));
        }
        // ---- O = P V (P written by the softmax warps over the Q tile) ----
        mbar_wait(bar_p, par);
        if (tr && i == 0) tr[5] = clock64();
        mbar_wait(bar_v, par);
        tc_fence_after();
        for (int j = 0; j < nsteps; ++j)
          umma_f16(tmem_base, make_smem_desc_sw32(QPs + j * P_BLOCK_BYTES), make_smem_desc(Vs + j * 16 * ROW_BYTES),
                   idesc_o, j != 0);
        umma_commit(bar_o);
        // ---- V and the Q/P region are free once O is complete: next item's Q and V ----
        if (next < num_items) {
          mbar_wait(bar_o, par);
          load_qv(item_of(next));
        }
      }
    }
  } else {
    // ===== softmax + epilogue: warps w and w+4 own TMEM lanes (= query rows) 32*(w%4) .. +31 =====
    const int q = warp & 3;
    const int hf = (warp - 1) >> 2;  // which share of the key blocks / of the output columns
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const int nblk = n16 >> 4;
    const int b_mid = (nblk + 1) >> 1;
    const int b0 = hf ? b_mid : 0, b1 = hf ? nblk : b_mid;  // this thread's 16-key blocks
    const uint32_t sw = (uint32_t)((row >> 2) & 1);  // 32B swizzle: 16-byte chunk index ^= address bit 7
    uint8_t *prow = QP_gen + row * 32;
    uint32_t i = 0;
    for (int id = blockIdx.x; id < num_items; id += gridDim.x, ++i) {
      const uint32_t par = i & 1;
      const Item it = item_of(id);
      const int rows_here = it.c_rows + it.p_rows;
      const bool warp_valid = q * 32 < rows_here;  // warp-uniform, identical for the two partner warps
      // every warp waits for S, also those without valid rows: S(i) exists only after all warps arrived for item
      // i-1, so no warp can arrive twice in one barrier phase
      mbar_wait(bar_s, par);
      if (warp_valid) {
        tc_fence_after();
        if (tr && i == 0 && threadIdx.x == 32) tr[3] = clock64();
        // ---- pass 1: row maximum over this thread's keys ----
        float mx = -INFINITY;
        {
          uint32_t cur[16], nxt[16];
          tmem_ld16_nowait(taddr + (uint32_t)(b0 * 16), cur);
          tmem_ld_wait(cur);
          for (int b = b0; b < b1; ++b) {
            if (b + 1 < b1) tmem_ld16_nowait(taddr + (uint32_t)((b + 1) * 16), nxt);
            if (b * 16 + 16 <= n) {
#pragma unroll
              for (int e = 0; e < 16; e += 2)
                mx = fmaxf(mx, fmaxf(__uint_as_float(cur[e]), __uint_as_float(cur[e + 1])));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (b * 16 + e < n) mx = fmaxf(mx, __uint_as_float(cur[e]));
            }
            if (b + 1 < b1) {
              tmem_ld_wait(nxt);
#pragma unroll
              for (int e = 0; e < 16; ++e) cur[e] = nxt[e];
            }
          }
        }
        red_max[hf * QT + row] = mx;
        pair_bar_sync(1 + q);
        if (tr && i == 0 && threadIdx.x == 32) tr[4] = clock64();
        mx = fmaxf(red_max[row], red_max[QT + row]);  // every row sees key 0, so the maximum is finite
        const float off = mx * sl2;
        // ---- pass 2: probabilities -> P blocks, row sum ----
        float l = 0.f, l_odd = 0.f;
        {
          uint32_t cur[16], nxt[16];
          tmem_ld16_nowait(taddr + (uint32_t)(b0 * 16), cur);
          tmem_ld_wait(cur);
          for (int b = b0; b < b1; ++b) {
            if (b + 1 < b1) tmem_ld16_nowait(taddr + (uint32_t)((b + 1) * 16), nxt);
            uint32_t pk[8];
            if (b * 16 + 16 <= n && (geo.flags & 1)) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float x0 = __uint_as_float(cur[2 * e]), x1 = __uint_as_float(cur[2 * e + 1]);
                ffma2(x0, x1, sl2, -off);
                const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                fadd2(l, l_odd, p0, p1);
                pk[e] = pack2<T>(p0, p1);
              }
            } else if (b * 16 + 16 <= n) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(cur[2 * e]), sl2, -off));
                const float p1 = ex2_approx(fmaf(__uint_as_float(cur[2 * e + 1]), sl2, -off));
                l += p0 + p1;
                pk[e] = pack2<T>(p0, p1);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int col = b * 16 + 2 * e;
                const float p0 = col < n ? ex2_approx(fmaf(__uint_as_float(cur[2 * e]), sl2, -off)) : 0.f;
                const float p1 = col + 1 < n ? ex2_approx(fmaf(__uint_as_float(cur[2 * e + 1]), sl2, -off)) : 0.f;
                l += p0 + p1;
                pk[e] = pack2<T>(p0, p1);
              }
            }
            uint8_t *dst = prow + b * P_BLOCK_BYTES;
            *reinterpret_cast<uint4 *>(dst + ((0u ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(dst + ((1u ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (b + 1 < b1) {
              tmem_ld_wait(nxt);
#pragma unroll
              for (int e = 0; e < 16; ++e) cur[e] = nxt[e];
            }
          }
        }
        red_sum[hf * QT + row] = l + l_odd;
        // make the generic-proxy stores of P visible to the tensor core (async proxy), release S
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (warp_valid) {
        mbar_wait(bar_o, par);
        tc_fence_after();
        if (tr && i == 0 && threadIdx.x == 32) tr[6] = clock64();
        const float inv = 1.0f / (red_sum[row] + red_sum[QT + row]);
        T *dst = nullptr;
        if (row < it.c_rows)
          dst = out_ctx + ((long long)it.g * n + it.t * QT + row) * D + it.h * HD;
        else if (row < rows_here)
          dst = out_prompt + ((long long)it.g * K + (row - it.c_rows)) * D + it.h * HD;
        uint32_t acc[32];
        tmem_ld32(taddr + (uint32_t)(hf * 32), acc);  // this thread's 32 of the 64 output columns
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_e);  // O is out of TMEM: the next S may overwrite it
        if (dst) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 o;
            o.x = pack2<T>(__uint_as_float(acc[8 * v + 0]) * inv, __uint_as_float(acc[8 * v + 1]) * inv);
            o.y = pack2<T>(__uint_as_float(acc[8 * v + 2]) * inv, __uint_as_float(acc[8 * v + 3]) * inv);
            o.z = pack2<T>(__uint_as_float(acc[8 * v + 4]) * inv, __uint_as_float(acc[8 * v + 5]) * inv);
            o.w = pack2<T>(__uint_as_float(acc[8 * v + 6]) * inv, __uint_as_float(acc[8 * v + 7]) * inv);
            *reinterpret_cast<uint4 *>(dst + hf * 32 + v * 8) = o;
          }
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_e);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) tr[7] = clock64();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace atc

bool ro_attention_fwd_dense_supported(int dtype, int n, int K, int H) {
  if (dtype != RPO_F16 && dtype != RPO_BF16) return false;
  if (n < 1 || K < 0 || H < 1) return false;
  const int n16 = (n + 15) & ~15;
  if (n16 > 256) return false;                 // one UMMA N, 256 TMEM columns
  if (K > 0 && (n % 128) + K > 128) return false;  // all prompt rows in one query tile
  static const bool off = [] { const char *e = getenv("RPO_ATTN_NO_TC"); return e && e[0] == '1'; }();
  return !off;
}

template <typename T>
int ro_attention_fwd_dense(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, int G, int n, int K, int H,
                           cudaStream_t st) {
  if constexpr (sizeof(T) != 2) {
    set_error("tcgen05 attention supports f16/bf16 only");
    return RPO_ERR_INVALID;
  } else {
    using namespace atc;
    RPO_REQUIRE(ro_attention_fwd_dense_supported(Num<T>::dtype, n, K, H), "tcgen05 attention shape");
    RPO_REQUIRE(G >= 1, "at least one group");
    RPO_REQUIRE((((uintptr_t)qkv_ctx | (uintptr_t)q_prompt | (uintptr_t)out_ctx | (uintptr_t)out_prompt) & 15) == 0,
                "attention buffers must be 16-byte aligned");
    const int D = H * HD;
    const int n16 = (n + 15) & ~15;
    Geo geo;
    geo.n = n;
    geo.n16 = n16;
    geo.K = K;
    geo.prompt_tile = n / 128;
    geo.prompt_row = n % 128;
    geo.H = H;
    geo.phase_delay = 0;
    geo.flags = 0;
    if (const char *e = getenv("RPO_ATTN_PHASE_DELAY")) geo.phase_delay = atoi(e);
    if (const char *e = getenv("RPO_ATTN_F32X2")) geo.flags |= (e[0] == '1') ? 1 : 0;
    const int tiles = K > 0 ? geo.prompt_tile + 1 : (n + 127) / 128;
    CUtensorMap map_full, map_kvt, map_qt, map_prompt;
    const long long Mc = (long long)G * n;
    RPO_TRY(make_map(&map_full, Num<T>::dtype, qkv_ctx, Mc, 3 * D, 3LL * D, 128));
    RPO_TRY(make_map(&map_kvt, Num<T>::dtype, qkv_ctx, Mc, 3 * D, 3LL * D, n16 % 128 ? n16 % 128 : 128));
    RPO_TRY(make_map(&map_qt, Num<T>::dtype, qkv_ctx, Mc, 3 * D, 3LL * D, n % 128 ? n % 128 : 128));
    if (K > 0)
      RPO_TRY(make_map(&map_prompt, Num<T>::dtype, q_prompt, (long long)G * K, D, D, K));
    else
      map_prompt = map_full;
    const int kv_bytes = n16 * ROW_BYTES;
    const int p_bytes = (n16 >> 4) * P_BLOCK_BYTES;
    const int smem = 2 * kv_bytes + (p_bytes > QT * ROW_BYTES ? p_bytes : QT * ROW_BYTES) + 64 + 4 * QT * 4 + 1024;
    static int configured = 0;
    if (smem > configured) {
      RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_fwd_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured = smem;
    }
    geo.tiles = tiles;
    const long long items = (long long)tiles * H * G;
    RPO_REQUIRE(items <= 0x7fffffffLL, "grid limits");
    const int slots = 2 * sm_count();  // two CTAs per SM (104 KB of shared memory, 256 TMEM columns each)
    dim3 grid((unsigned)(items < slots ? items : slots));
    long long *trace = nullptr;
    if (const char *e = getenv("RPO_ATTN_TRACE")) trace = reinterpret_cast<long long *>(strtoull(e, nullptr, 0));
    prof_tag("attn_fwd_tc G=%d H=%d K=%d n=%d", G, H, K, n);
    RPO_CHECK_CUDA(launch_pdl(ro_attn_fwd_tc<T>, grid, dim3(THREADS), (size_t)smem, st, map_full, map_kvt, map_qt,
                              map_prompt, out_ctx, out_prompt, geo, (int)items, trace));
    RPO_LAUNCH_CHECK();
    return RPO_OK;
  }
}

template int ro_attention_fwd_dense<float>(const float *, const float *, float *, float *, int, int, int, int,
                                           cudaStream_t);
template int ro_attention_fwd_dense<__half>(const __half *, const __half *, __half *, __half *, int, int, int, int,
                                            cudaStream_t);
template int ro_attention_fwd_dense<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, __nv_bfloat16 *,
                                                   __nv_bfloat16 *, int, int, int, int, cudaStream_t);

}  // namespace rpo
