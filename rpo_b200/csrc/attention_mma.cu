// Tensor-core path of the read-only masked attention for the 16-bit dtypes (see attention.cu for the
// semantics and the exact-f32 SIMT path).
//
// One CTA = one (group, head, 64-query tile); 4 warps x 16 query rows.  K and V of the (group, head)
// (n <= 288 rows of 64) and the Q tile are staged in shared memory with cp.async (16-byte chunks,
// XOR-swizzled so ldmatrix is bank-conflict free).  Because a whole key row fits on chip, softmax is
// single pass: S = Q K^T lives in registers (mma.sync m16n8k16, f32 accumulate), masking is a column
// bound per row (the read-only mask is "keys j < n_vis"), P is rounded to the dtype exactly where the
// reference materialises the probability tensor, and O = P V reuses the accumulator registers as the
// A operand.  The kernel is HBM/L2-bound by design (AI ~ 104 FLOP/B): what matters is
// 128-byte coalesced row loads, one pass over Q/K/V/O, and enough CTAs (G*H*4) to fill 148 SMs.
//
// The backward kernel produces dQ for the prompt queries only (nothing else on the path needs a
// gradient): P is recomputed, delta = rowsum(dO * O) comes from the saved forward output, and
// dS = P * (dP - delta) / 8 is formed 16 keys at a time so that only S stays resident in registers.
#include <stdlib.h>

#include "common.cuh"

namespace rpo {

namespace amma {

static constexpr int HD = 64;
static constexpr int ROW_BYTES = HD * 2;  // 128 B per row of Q/K/V for one head
static constexpr int QT = 64;             // query rows per CTA
static constexpr int THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// byte offset of 16-byte chunk `c` (0..7) of row `r` in a swizzled [rows][128 B] tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * ROW_BYTES + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void st_zero16(uint32_t dst) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
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
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// stage `n` rows (zero-filled up to n16) of a [rows][64] head slice whose row r lives at base + r*ld
template <typename T>
__device__ __forceinline__ void stage_rows(uint32_t dst, const T *base, long long ld, int n, int n16) {
  for (int idx = threadIdx.x; idx < n16 * 8; idx += THREADS) {
    int r = idx >> 3, c = idx & 7;
    if (r < n)
      cp_async16(dst + swz(r, c), base + (long long)r * ld + c * 8);
    else
      st_zero16(dst + swz(r, c));
  }
}

// S = Q K^T for one warp's 16 query rows against all staged keys; acc[nt] is the m16n8 tile of keys
// nt*8 .. nt*8+7.  Q fragments are read from the warp's rows of the swizzled Q tile.
template <typename T, int NT>
__device__ __forceinline__ void qk_scores(uint32_t Qs, uint32_t Ks, int r_base, int n16, int lane,
                                          float (&acc)[NT][4]) {
  uint32_t qf[4][4];
  {
    const int m = lane >> 3;
    const int row = r_base + (m & 1) * 8 + (lane & 7);
#pragma unroll
    for (int kd = 0; kd < 4; ++kd) ldsm_x4(Qs + swz(row, kd * 2 + (m >> 1)), qf[kd][0], qf[kd][1], qf[kd][2], qf[kd][3]);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  const int m = lane >> 3;
#pragma unroll
  for (int np = 0; np < NT / 2; ++np) {
    if (np * 16 < n16) {
      const int key = (np * 2 + (m >> 1)) * 8 + (lane & 7);
#pragma unroll
      for (int kd = 0; kd < 4; ++kd) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(Ks + swz(key, kd * 2 + (m & 1)), b0, b1, b2, b3);
        mma16816<T>(acc[np * 2], qf[kd], b0, b1);
        mma16816<T>(acc[np * 2 + 1], qf[kd], b2, b3);
      }
    }
  }
}

// masked single-pass softmax over the register-resident scores of rows (lane/4) and (lane/4 + 8);
// leaves probabilities rounded through T in acc.
template <typename T, int NT>
__device__ __forceinline__ void softmax_rows(float (&acc)[NT][4], int n16, int nvis_a, int nvis_b, int lane) {
  const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int c0 = 2 * (lane & 3);
  float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt * 8 < n16) {
      const int c = nt * 8 + c0;
      if (c >= nvis_a) acc[nt][0] = -INFINITY;
      if (c + 1 >= nvis_a) acc[nt][1] = -INFINITY;
      if (c >= nvis_b) acc[nt][2] = -INFINITY;
      if (c + 1 >= nvis_b) acc[nt][3] = -INFINITY;
      mxa = fmaxf(mxa, fmaxf(acc[nt][0], acc[nt][1]));
      mxb = fmaxf(mxb, fmaxf(acc[nt][2], acc[nt][3]));
    }
  }
  mxa = quad_max(mxa);
  mxb = quad_max(mxb);
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt * 8 < n16) {
      acc[nt][0] = exp2f((acc[nt][0] - mxa) * sl2);
      acc[nt][1] = exp2f((acc[nt][1] - mxa) * sl2);
      acc[nt][2] = exp2f((acc[nt][2] - mxb) * sl2);
      acc[nt][3] = exp2f((acc[nt][3] - mxb) * sl2);
      sa += acc[nt][0] + acc[nt][1];
      sb += acc[nt][2] + acc[nt][3];
    }
  }
  const float ia = 1.0f / quad_sum(sa), ib = 1.0f / quad_sum(sb);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt * 8 < n16) {
      acc[nt][0] = rnd<T>(acc[nt][0] * ia);
      acc[nt][1] = rnd<T>(acc[nt][1] * ia);
      acc[nt][2] = rnd<T>(acc[nt][2] * ib);
      acc[nt][3] = rnd<T>(acc[nt][3] * ib);
    }
  }
}

// writes a warp's 16 x 64 f32 fragment tile through its own rows of a swizzled staging tile and then
// to global memory with 16-byte stores; row_ptr(r) gives the destination of local row r or nullptr.
template <typename T, typename RowPtr>
__device__ __forceinline__ void store_tile(uint32_t stage, uint8_t *stage_gen, int r_base, const float (&o)[8][4],
                                           int lane, RowPtr row_ptr) {
  const int ra = r_base + (lane >> 2), rb = ra + 8;
  const int cw = (lane & 3) * 4;  // byte offset inside the 16-byte chunk
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t va = pack2<T>(o[nt][0], o[nt][1]), vb = pack2<T>(o[nt][2], o[nt][3]);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage + swz(ra, nt) + cw), "r"(va) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage + swz(rb, nt) + cw), "r"(vb) : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i;
    const int r = r_base + (id >> 3), c = id & 7;
    T *dst = row_ptr(r);
    if (dst) *reinterpret_cast<uint4 *>(dst + c * 8) = *reinterpret_cast<const uint4 *>(stage_gen + swz(r, c));
  }
}

template <typename T, int NT>
__global__ void __launch_bounds__(THREADS)
    ro_attn_fwd_mma(const T *__restrict__ qkv_ctx, const T *__restrict__ q_prompt, T *__restrict__ out_ctx,
                    T *__restrict__ out_prompt, const int *__restrict__ ctx_off, int K, int H, int causal, int do_ctx) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int g = blockIdx.z, h = blockIdx.y;
  const int D = H * HD;
  pdl_wait();
  pdl_trigger();
  const int row0 = ctx_off[g];
  const int n = ctx_off[g + 1] - row0;
  const int n_ctx_q = do_ctx ? n : 0;
  const int n_q = n_ctx_q + K;
  const int q_begin = blockIdx.x * QT;
  if (q_begin >= n_q) return;
  const int n16 = (n + 15) & ~15;
  const uint32_t Ks = smem_u32(sm), Vs = Ks + n16 * ROW_BYTES, Qs = Vs + n16 * ROW_BYTES;
  uint8_t *Qs_gen = sm + 2 * n16 * ROW_BYTES;
  const T *kbase = qkv_ctx + (long long)row0 * 3 * D + D + h * HD;
  stage_rows<T>(Ks, kbase, 3LL * D, n, n16);
  stage_rows<T>(Vs, kbase + D, 3LL * D, n, n16);
  for (int idx = threadIdx.x; idx < QT * 8; idx += THREADS) {
    int r = idx >> 3, c = idx & 7;
    int qi = q_begin + r;
    if (qi < n_q) {
      const T *src = qi < n_ctx_q ? qkv_ctx + (long long)(row0 + qi) * 3 * D + h * HD
                                  : q_prompt + ((long long)g * K + (qi - n_ctx_q)) * D + h * HD;
      cp_async16(Qs + swz(r, c), src + c * 8);
    } else {
      st_zero16(Qs + swz(r, c));
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_base = warp * 16;
  if (q_begin + r_base >= n_q) return;  // no block-level sync below
  float acc[NT][4];
  qk_scores<T, NT>(Qs, Ks, r_base, n16, lane, acc);
  const int qa = q_begin + r_base + (lane >> 2), qb = qa + 8;
  const int nvis_a = (causal && qa < n_ctx_q) ? min(n, qa + 1) : n;
  const int nvis_b = (causal && qb < n_ctx_q) ? min(n, qb + 1) : n;
  softmax_rows<T, NT>(acc, n16, nvis_a, nvis_b, lane);
  // O = P V
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  const int m = lane >> 3;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    if (kk * 16 < n16) {
      uint32_t a[4];
      a[0] = pack2<T>(acc[2 * kk][0], acc[2 * kk][1]);
      a[1] = pack2<T>(acc[2 * kk][2], acc[2 * kk][3]);
      a[2] = pack2<T>(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
      a[3] = pack2<T>(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
      const int key = kk * 16 + (m & 1) * 8 + (lane & 7);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(Vs + swz(key, dp * 2 + (m >> 1)), b0, b1, b2, b3);
        mma16816<T>(o[dp * 2], a, b0, b1);
        mma16816<T>(o[dp * 2 + 1], a, b2, b3);
      }
    }
  }
  __syncwarp();  // all lanes have consumed the warp's Q rows: reuse them as the output staging tile
  store_tile<T>(Qs, Qs_gen, r_base, o, lane, [&](int r) -> T * {
    int qi = q_begin + r;
    if (qi >= n_q) return nullptr;
    return qi < n_ctx_q ? out_ctx + (long long)(row0 + qi) * D + h * HD
                        : out_prompt + ((long long)g * K + (qi - n_ctx_q)) * D + h * HD;
  });
}

template <typename T, int NT>
__global__ void __launch_bounds__(THREADS)
    ro_attn_bwd_mma(const T *__restrict__ qkv_ctx, const T *__restrict__ q_prompt, const T *__restrict__ o_prompt,
                    const T *__restrict__ d_out, T *__restrict__ dq, const int *__restrict__ ctx_off, int K, int H,
                    int settled) {
  // settled: see ro_attn_bwd_ks -- K, V, Q and the attention output rows are fetched ahead of the dependency wait
  extern __shared__ __align__(128) uint8_t sm[];
  const int g = blockIdx.y, h = blockIdx.x;
  const int D = H * HD;
  if (!settled) {
    pdl_wait();
    pdl_trigger();
  }
  const int row0 = ctx_off[g];
  const int n = ctx_off[g + 1] - row0;
  const int n16 = (n + 15) & ~15;
  const int q_begin = blockIdx.z * QT;  // K > 64 prompts: several tiles
  const uint32_t Ks = smem_u32(sm), Vs = Ks + n16 * ROW_BYTES, Qs = Vs + n16 * ROW_BYTES, dOs = Qs + QT * ROW_BYTES;
  uint8_t *Qs_gen = sm + 2 * n16 * ROW_BYTES;
  const T *kbase = qkv_ctx + (long long)row0 * 3 * D + D + h * HD;
  stage_rows<T>(Ks, kbase, 3LL * D, n, n16);
  stage_rows<T>(Vs, kbase + D, 3LL * D, n, n16);
  const long long pbase = ((long long)g * K + q_begin) * D + h * HD;
  const int rows_here = min(QT, K - q_begin);
  stage_rows<T>(Qs, q_prompt + pbase, D, rows_here, QT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_base = warp * 16;
  // attention output rows of delta_r = sum_d dO[r,d] * O[r,d] (two lanes per row, 32 columns each): loaded next to the
  // copies instead of in a round trip of their own after the barrier
  const int dr = r_base + (lane >> 1);
  uint4 ov[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) ov[v] = make_uint4(0u, 0u, 0u, 0u);
  const T *po = o_prompt + pbase + (long long)dr * D + (lane & 1) * 32;
  if (settled) {
    if (dr < rows_here) {
#pragma unroll
      for (int v = 0; v < 4; ++v) ov[v] = __ldg(reinterpret_cast<const uint4 *>(po + v * 8));
    }
    pdl_wait();
    pdl_trigger();
  } else if (dr < rows_here) {
#pragma unroll
    for (int v = 0; v < 4; ++v) ov[v] = *reinterpret_cast<const uint4 *>(po + v * 8);
  }
  stage_rows<T>(dOs, d_out + pbase, D, rows_here, QT);
  cp_async_wait_all();
  __syncthreads();
  if (r_base >= rows_here) return;
  float delta_a, delta_b;
  {
    float s = 0.f;
    if (dr < rows_here) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        Vec16<T> a, b;  // dO from the staged tile (the same values the MMA fragments read)
        *reinterpret_cast<uint4 *>(&a) = ov[v];
        uint4 bq;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(bq.x), "=r"(bq.y), "=r"(bq.z), "=r"(bq.w)
                     : "r"(dOs + swz(dr, (lane & 1) * 4 + v)));
        *reinterpret_cast<uint4 *>(&b) = bq;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += tof<T>(a.v[e]) * tof<T>(b.v[e]);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    delta_a = __shfl_sync(0xffffffffu, s, 2 * (lane >> 2));
    delta_b = __shfl_sync(0xffffffffu, s, 2 * ((lane >> 2) + 8));
  }
  float acc[NT][4];
  qk_scores<T, NT>(Qs, Ks, r_base, n16, lane, acc);
  softmax_rows<T, NT>(acc, n16, n, n, lane);
  // dO fragments (A operand of dP = dO V^T)
  uint32_t df[4][4];
  const int m = lane >> 3;
  {
    const int row = r_base + (m & 1) * 8 + (lane & 7);
#pragma unroll
    for (int kd = 0; kd < 4; ++kd) ldsm_x4(dOs + swz(row, kd * 2 + (m >> 1)), df[kd][0], df[kd][1], df[kd][2], df[kd][3]);
  }
  // dS is handed to the tensor core in the 16-bit dtype; for fp16 it is pre-scaled by 2^8 (exact) so
  // that products of small probabilities and small gradients stay out of the subnormal range
  constexpr float kDs = Num<T>::dtype == RPO_F16 ? 32.0f : 0.125f;
  constexpr float kUn = Num<T>::dtype == RPO_F16 ? 1.0f / 256.0f : 1.0f;
  float gq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) gq[nt][0] = gq[nt][1] = gq[nt][2] = gq[nt][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    if (kk * 16 < n16) {
      float dp0[4] = {0.f, 0.f, 0.f, 0.f}, dp1[4] = {0.f, 0.f, 0.f, 0.f};
      const int keyn = (kk * 2 + (m >> 1)) * 8 + (lane & 7);
#pragma unroll
      for (int kd = 0; kd < 4; ++kd) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(Vs + swz(keyn, kd * 2 + (m & 1)), b0, b1, b2, b3);
        mma16816<T>(dp0, df[kd], b0, b1);
        mma16816<T>(dp1, df[kd], b2, b3);
      }
      uint32_t a[4];
      a[0] = pack2<T>(acc[2 * kk][0] * (dp0[0] - delta_a) * kDs, acc[2 * kk][1] * (dp0[1] - delta_a) * kDs);
      a[1] = pack2<T>(acc[2 * kk][2] * (dp0[2] - delta_b) * kDs, acc[2 * kk][3] * (dp0[3] - delta_b) * kDs);
      a[2] = pack2<T>(acc[2 * kk + 1][0] * (dp1[0] - delta_a) * kDs,
                      acc[2 * kk + 1][1] * (dp1[1] - delta_a) * kDs);
      a[3] = pack2<T>(acc[2 * kk + 1][2] * (dp1[2] - delta_b) * kDs,
                      acc[2 * kk + 1][3] * (dp1[3] - delta_b) * kDs);
      const int keyk = kk * 16 + (m & 1) * 8 + (lane & 7);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(Ks + swz(keyk, dp * 2 + (m >> 1)), b0, b1, b2, b3);
        mma16816<T>(gq[dp * 2], a, b0, b1);
        mma16816<T>(gq[dp * 2 + 1], a, b2, b3);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    gq[nt][0] *= kUn;
    gq[nt][1] *= kUn;
    gq[nt][2] *= kUn;
    gq[nt][3] *= kUn;
  }
  __syncwarp();
  store_tile<T>(Qs, Qs_gen, r_base, gq, lane, [&](int r) -> T * {
    return r < rows_here ? dq + pbase + (long long)r * D : nullptr;
  });
}

// ---- forward, streaming ("flash") form -----------------------------------------------------------
// One CTA = one (group, head) and up to 128 query rows (blockDim.x/32 warps x 16 rows).  K and V of
// the (group, head) are staged ONCE per CTA in 64-key chunks, one cp.async commit group per chunk, so
// the tensor-core work on chunk c overlaps the loads of chunks c+1.. ; softmax is the online form
// (running row max m and row sum l, accumulator rescaled by exp2((m_old - m_new)/8 log2e) per chunk),
// which keeps the live score fragment at 64 keys = 32 registers and lets two CTAs (16 warps) share an
// SM.  Probabilities are rounded to the 16-bit dtype where they feed the P.V tensor-core product (the
// reference rounds the normalised probability tensor; here the division by l happens once at the
// end in f32 -- same 2^-11 relative rounding, within the 1e-3 parity bar).
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
  }
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
static constexpr int KC = 64;          // keys per chunk
static constexpr int MAX_CHUNKS = 5;   // 288 context rows
static constexpr int FL_MAX_WARPS = 8;

// rows [r_lo, r_hi) of a head slice into swizzled smem rows (zero-filled at and beyond n)
template <typename T>
__device__ __forceinline__ void stage_row_range(uint32_t dst, const T *base, long long ld, int r_lo, int r_hi, int n) {
  for (int idx = threadIdx.x + r_lo * 8; idx < r_hi * 8; idx += blockDim.x) {
    int r = idx >> 3, c = idx & 7;
    if (r < n)
      cp_async16(dst + swz(r, c), base + (long long)r * ld + c * 8);
    else
      st_zero16(dst + swz(r, c));
  }
}

template <typename T>
__global__ void __launch_bounds__(FL_MAX_WARPS * 32, 2)
    ro_attn_fwd_flash(const T *__restrict__ qkv_ctx, const T *__restrict__ q_prompt, T *__restrict__ out_ctx,
                      T *__restrict__ out_prompt, const int *__restrict__ ctx_off, int K, int H, int causal,
                      int do_ctx) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int g = blockIdx.z, h = blockIdx.y;
  const int D = H * HD;
  pdl_wait();
  pdl_trigger();
  const int row0 = ctx_off[g];
  const int n = ctx_off[g + 1] - row0;
  const int n_ctx_q = do_ctx ? n : 0;
  const int n_q = n_ctx_q + K;
  const int rows_cta = (blockDim.x >> 5) * 16;
  const int q_begin = blockIdx.x * rows_cta;
  if (q_begin >= n_q) return;
  const int n16 = (n + 15) & ~15;
  const int nchunks = (n16 + KC - 1) / KC;
  const uint32_t Ks = smem_u32(sm), Vs = Ks + n16 * ROW_BYTES, Qs = Vs + n16 * ROW_BYTES;
  uint8_t *Qs_gen = sm + 2 * n16 * ROW_BYTES;
  const T *kbase = qkv_ctx + (long long)row0 * 3 * D + D + h * HD;
  // group 0: Q tile + first K/V chunk; group c: K/V chunk c
  for (int idx = threadIdx.x; idx < rows_cta * 8; idx += blockDim.x) {
    int r = idx >> 3, c = idx & 7;
    int qi = q_begin + r;
    if (qi < n_q) {
      const T *src = qi < n_ctx_q ? qkv_ctx + (long long)(row0 + qi) * 3 * D + h * HD
                                  : q_prompt + ((long long)g * K + (qi - n_ctx_q)) * D + h * HD;
      cp_async16(Qs + swz(r, c), src + c * 8);
    } else {
      st_zero16(Qs + swz(r, c));
    }
  }
  for (int c = 0; c < nchunks; ++c) {
    const int lo = c * KC, hi = min(n16, lo + KC);
    stage_row_range<T>(Ks, kbase, 3LL * D, lo, hi, n);
    stage_row_range<T>(Vs, kbase + D, 3LL * D, lo, hi, n);
    cp_async_commit();
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_base = warp * 16;
  const bool active = q_begin + r_base < n_q;  // warp-uniform
  const int m = lane >> 3;
  const int qa = q_begin + r_base + (lane >> 2), qb = qa + 8;
  const int nvis_a = (causal && qa < n_ctx_q) ? min(n, qa + 1) : n;
  const int nvis_b = (causal && qb < n_ctx_q) ? min(n, qb + 1) : n;
  const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int c0 = 2 * (lane & 3);

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
  // per-lane ldmatrix offsets.  Key rows advance in multiples of 8, so (row & 7) == (lane & 7) and the
  // XOR swizzle term depends only on the lane and the (compile-time) 16-byte column index.
  uint32_t xk[4], xv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xk[i] = (uint32_t)(((i * 2 + (m & 1)) ^ (lane & 7)) << 4);
    xv[i] = (uint32_t)(((i * 2 + (m >> 1)) ^ (lane & 7)) << 4);
  }
  const uint32_t k_lane = Ks + (uint32_t)(((m >> 1) * 8 + (lane & 7)) * ROW_BYTES);
  const uint32_t v_lane = Vs + (uint32_t)(((m & 1) * 8 + (lane & 7)) * ROW_BYTES);

  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait_pending(nchunks - 1 - c);
    __syncthreads();
    if (!active) continue;
    if (c == 0) {
      const int row = r_base + (m & 1) * 8 + (lane & 7);
#pragma unroll
      for (int kd = 0; kd < 4; ++kd)
        ldsm_x4(Qs + swz(row, kd * 2 + (m >> 1)), qf[kd][0], qf[kd][1], qf[kd][2], qf[kd][3]);
    }
    const int k0 = c * KC;
    const int ng = (min(n16, k0 + KC) - k0) >> 4;  // 16-key groups in this chunk (1..4), CTA-uniform
    const uint32_t kc = k_lane + (uint32_t)(k0 * ROW_BYTES), vc = v_lane + (uint32_t)(k0 * ROW_BYTES);
    float s[8][4];
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      s[2 * np][0] = s[2 * np][1] = s[2 * np][2] = s[2 * np][3] = 0.f;
      s[2 * np + 1][0] = s[2 * np + 1][1] = s[2 * np + 1][2] = s[2 * np + 1][3] = 0.f;
      if (np < ng) {
#pragma unroll
        for (int kd = 0; kd < 4; ++kd) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(kc + (uint32_t)(np * 16 * ROW_BYTES) + xk[kd], b0, b1, b2, b3);
          mma16816<T>(s[2 * np], qf[kd], b0, b1);
          mma16816<T>(s[2 * np + 1], qf[kd], b2, b3);
        }
      }
    }
    // mask (only where something is masked: causal rows, the ragged tail of the last chunk) + chunk max
    float mxa = -INFINITY, mxb = -INFINITY;
    if (causal || k0 + KC > n) {  // CTA-uniform
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = k0 + nt * 8 + c0;
        if (col >= nvis_a) s[nt][0] = -INFINITY;
        if (col + 1 >= nvis_a) s[nt][1] = -INFINITY;
        if (col >= nvis_b) s[nt][2] = -INFINITY;
        if (col + 1 >= nvis_b) s[nt][3] = -INFINITY;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mxa = fmaxf(mxa, fmaxf(s[nt][0], s[nt][1]));
      mxb = fmaxf(mxb, fmaxf(s[nt][2], s[nt][3]));
    }
    // every row sees key 0, so after chunk 0 the running max is finite
    const float mna = fmaxf(m_a, quad_max(mxa)), mnb = fmaxf(m_b, quad_max(mxb));
    const float al_a = ex2_approx((m_a - mna) * sl2), al_b = ex2_approx((m_b - mnb) * sl2);
    m_a = mna;
    m_b = mnb;
    const float oa = mna * sl2, ob = mnb * sl2;
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = ex2_approx(fmaf(s[nt][0], sl2, -oa));
      s[nt][1] = ex2_approx(fmaf(s[nt][1], sl2, -oa));
      s[nt][2] = ex2_approx(fmaf(s[nt][2], sl2, -ob));
      s[nt][3] = ex2_approx(fmaf(s[nt][3], sl2, -ob));
      sa += s[nt][0] + s[nt][1];
      sb += s[nt][2] + s[nt][3];
    }
    l_a = l_a * al_a + sa;
    l_b = l_b * al_b + sb;
    if (__any_sync(0xffffffffu, al_a != 1.0f || al_b != 1.0f)) {  // running max moved for some row of the warp
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o[nt][0] *= al_a;
        o[nt][1] *= al_a;
        o[nt][2] *= al_b;
        o[nt][3] *= al_b;
      }
    }
    // O += P V_chunk
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk < ng) {
        uint32_t a[4];
        a[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(vc + (uint32_t)(kk * 16 * ROW_BYTES) + xv[dp], b0, b1, b2, b3);
          mma16816<T>(o[dp * 2], a, b0, b1);
          mma16816<T>(o[dp * 2 + 1], a, b2, b3);
        }
      }
    }
  }
  if (!active) return;
  const float ia = 1.0f / quad_sum(l_a), ib = 1.0f / quad_sum(l_b);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    o[nt][0] *= ia;
    o[nt][1] *= ia;
    o[nt][2] *= ib;
    o[nt][3] *= ib;
  }
  __syncwarp();  // the warp's Q rows were consumed into registers at chunk 0: reuse them as staging
  store_tile<T>(Qs, Qs_gen, r_base, o, lane, [&](int r) -> T * {
    int qi = q_begin + r;
    if (qi >= n_q) return nullptr;
    return qi < n_ctx_q ? out_ctx + (long long)(row0 + qi) * D + h * HD
                        : out_prompt + ((long long)g * K + (qi - n_ctx_q)) * D + h * HD;
  });
}

template <typename T>
static int launch_fwd_flash(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G,
                            int K, int H, int max_ctx, int causal, int do_ctx, cudaStream_t st) {
  const int n16 = (max_ctx + 15) & ~15;
  const int n_q = (do_ctx ? max_ctx : 0) + K;
  const int tiles = (n_q + 15) / 16;
  const int ncta = (tiles + FL_MAX_WARPS - 1) / FL_MAX_WARPS;
  const int warps = (tiles + ncta - 1) / ncta;  // L=221: 2 CTAs x 7 warps; L=281: 3 x 6
  const int smem = (2 * n16 + warps * 16) * ROW_BYTES;
  static int configured = 0;
  if (smem > configured) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_fwd_flash<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(ncta, H, G);
  prof_tag("attn_fwd G=%d H=%d K=%d max_ctx=%d do_ctx=%d", G, H, K, max_ctx, do_ctx);
  RPO_CHECK_CUDA(launch_pdl(ro_attn_fwd_flash<T>, grid, dim3(warps * 32), smem, st, qkv_ctx, q_prompt, out_ctx, out_prompt,
                            ctx_off, K, H, causal, do_ctx));
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// ---- backward, key-split form ---------------------------------------------------------------------------
// The gradient exists for the K prompt queries only (K = 24: two 16-row MMA tiles per (group, head)), so the
// single-pass kernel above keeps just two warps per CTA busy with a 26-tile score fragment each.  Here the four
// warps of a CTA split the KEYS as well: warp (mt, ks) owns query tile mt and every KS-th share of the 16-key
// groups; row maxima, row sums and the dQ partial tiles are exchanged through shared memory.  Half the score
// registers per warp (three CTAs per SM instead of two) and twice the active warps per CTA.
template <typename T, int NG>  // NG: 16-key groups per warp (compile-time bound of the fragment arrays)
__global__ void __launch_bounds__(THREADS)
    ro_attn_bwd_ks(const T *__restrict__ qkv_ctx, const T *__restrict__ q_prompt, const T *__restrict__ o_prompt,
                   const T *__restrict__ d_out, T *__restrict__ dq, const int *__restrict__ ctx_off, int K, int H,
                   int settled) {
  // settled (common.cuh, SettledOperands): keys, values and prompt queries date from the forward pass, so their copies
  // are issued before the dependency wait and land while the upstream GEMM (which produces d_out) is still running
  extern __shared__ __align__(128) uint8_t sm[];
  const int g = blockIdx.y, h = blockIdx.x;
  const int D = H * HD;
  if (!settled) {
    pdl_wait();
    pdl_trigger();
  }
  const int row0 = ctx_off[g];
  const int n = ctx_off[g + 1] - row0;
  const int n16 = (n + 15) & ~15;
  const int q_begin = blockIdx.z * QT;
  const int rows_here = min(QT, K - q_begin);
  const int mt_n = (rows_here + 15) >> 4;            // query tiles with rows (1..4)
  const int mt_slots = mt_n <= 1 ? 1 : (mt_n == 2 ? 2 : 4);
  const int KS = 4 / mt_slots;                        // key shares per query tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp / KS, ks = warp % KS;
  const int ng_total = n16 >> 4;
  const int ng_per = (ng_total + KS - 1) / KS;        // <= NG (checked by the launcher)
  const int g_lo = min(ks * ng_per, ng_total), g_hi = min(g_lo + ng_per, ng_total);
  // [K | V, later the f32 dQ partial tiles][Q][dO][row maxima and sums]
  const int kv_bytes = 2 * n16 * ROW_BYTES, part_bytes = 4 * 16 * HD * 4;
  const int q_off = kv_bytes > part_bytes ? kv_bytes : part_bytes;
  const uint32_t Ks = smem_u32(sm), Vs = Ks + n16 * ROW_BYTES, Qs = Ks + q_off, dOs = Qs + QT * ROW_BYTES;
  float *red = reinterpret_cast<float *>(sm + q_off + 2 * QT * ROW_BYTES);  // [2][4 shares][QT] max, sum
  const T *kbase = qkv_ctx + (long long)row0 * 3 * D + D + h * HD;
  // two copy groups: K and Q feed the score pass; V (half of the bytes) is first read by the dP pass two barriers
  // further down, so its copy stays in flight behind the score pass and the row-maximum / row-sum exchanges
  stage_rows<T>(Ks, kbase, 3LL * D, n, n16);
  const long long pbase = ((long long)g * K + q_begin) * D + h * HD;
  stage_rows<T>(Qs, q_prompt + pbase, D, rows_here, QT);
  cp_async_commit();
  stage_rows<T>(Vs, kbase + D, 3LL * D, n, n16);
  cp_async_commit();
  const int r_base = mt * 16;
  const bool active = mt < mt_n;  // warp-uniform; inactive warps only take part in the barriers
  // the attention output rows of delta_r = sum_d dO[r,d] * O[r,d] (two lanes per row, 32 columns each) also date from
  // the forward pass: fetched here, next to the copies, instead of in a round trip of their own after the barrier
  const int dr = r_base + (lane >> 1);
  uint4 ov[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) ov[v] = make_uint4(0u, 0u, 0u, 0u);
  if (settled && active && dr < rows_here) {
    const T *po = o_prompt + pbase + (long long)dr * D + (lane & 1) * 32;
#pragma unroll
    for (int v = 0; v < 4; ++v) ov[v] = __ldg(reinterpret_cast<const uint4 *>(po + v * 8));
  }
  if (settled) {
    pdl_wait();
    pdl_trigger();
  } else if (active && dr < rows_here) {
    const T *po = o_prompt + pbase + (long long)dr * D + (lane & 1) * 32;
#pragma unroll
    for (int v = 0; v < 4; ++v) ov[v] = *reinterpret_cast<const uint4 *>(po + v * 8);
  }
  // dO (a few KB, the one operand the upstream kernel wrote) through registers: it must not queue behind V's group
  {
    constexpr int PER = QT * 8 / THREADS;
    uint4 dv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {  // all loads first: one round trip, not PER
      const int idx = threadIdx.x + i * THREADS, r = idx >> 3, c = idx & 7;
      dv[i] = make_uint4(0u, 0u, 0u, 0u);
      if (r < rows_here) dv[i] = *reinterpret_cast<const uint4 *>(d_out + pbase + (long long)r * D + c * 8);
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int idx = threadIdx.x + i * THREADS, r = idx >> 3, c = idx & 7;
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dOs + swz(r, c)), "r"(dv[i].x), "r"(dv[i].y),
                   "r"(dv[i].z), "r"(dv[i].w)
                   : "memory");
    }
  }
  cp_async_wait_pending(1);
  __syncthreads();
  const int ra = r_base + (lane >> 2), rb = ra + 8;  // this lane's two rows inside the CTA's row block
  float delta_a = 0.f, delta_b = 0.f;
  if (active) {
    float sdl = 0.f;
    if (dr < rows_here) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        Vec16<T> a, b;  // dO from the staged tile (the same values the MMA fragments read)
        *reinterpret_cast<uint4 *>(&a) = ov[v];
        uint4 bq;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(bq.x), "=r"(bq.y), "=r"(bq.z), "=r"(bq.w)
                     : "r"(dOs + swz(dr, (lane & 1) * 4 + v)));
        *reinterpret_cast<uint4 *>(&b) = bq;
#pragma unroll
        for (int e = 0; e < 8; ++e) sdl += tof<T>(a.v[e]) * tof<T>(b.v[e]);
      }
    }
    sdl += __shfl_xor_sync(0xffffffffu, sdl, 1);
    delta_a = __shfl_sync(0xffffffffu, sdl, 2 * (lane >> 2));
    delta_b = __shfl_sync(0xffffffffu, sdl, 2 * ((lane >> 2) + 8));
  }
  const int m = lane >> 3;
  const float sl2 = 0.125f * 1.4426950408889634f;
  const int c0 = 2 * (lane & 3);
  float acc[2 * NG][4];
  uint32_t qf[4][4], df[4][4];
  if (active) {
    const int row = r_base + (m & 1) * 8 + (lane & 7);
#pragma unroll
    for (int kd = 0; kd < 4; ++kd) {
      ldsm_x4(Qs + swz(row, kd * 2 + (m >> 1)), qf[kd][0], qf[kd][1], qf[kd][2], qf[kd][3]);
      ldsm_x4(dOs + swz(row, kd * 2 + (m >> 1)), df[kd][0], df[kd][1], df[kd][2], df[kd][3]);
    }
    // S = Q K^T over this warp's key groups; partial row maximum
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      acc[2 * gi][0] = acc[2 * gi][1] = acc[2 * gi][2] = acc[2 * gi][3] = 0.f;
      acc[2 * gi + 1][0] = acc[2 * gi + 1][1] = acc[2 * gi + 1][2] = acc[2 * gi + 1][3] = 0.f;
      if (g_lo + gi < g_hi) {
        const int key = (g_lo + gi) * 16 + (m >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int kd = 0; kd < 4; ++kd) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(Ks + swz(key, kd * 2 + (m & 1)), b0, b1, b2, b3);
          mma16816<T>(acc[2 * gi], qf[kd], b0, b1);
          mma16816<T>(acc[2 * gi + 1], qf[kd], b2, b3);
        }
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const int col = (g_lo + gi) * 16 + t2 * 8 + c0;
          float(&a4)[4] = acc[2 * gi + t2];
          if (col >= n) a4[0] = a4[2] = -INFINITY;
          if (col + 1 >= n) a4[1] = a4[3] = -INFINITY;
          mxa = fmaxf(mxa, fmaxf(a4[0], a4[1]));
          mxb = fmaxf(mxb, fmaxf(a4[2], a4[3]));
        }
      }
    }
    mxa = quad_max(mxa);
    mxb = quad_max(mxb);
    if ((lane & 3) == 0) {
      red[ks * QT + ra] = mxa;
      red[ks * QT + rb] = mxb;
    }
  }
  __syncthreads();
  float sa = 0.f, sb = 0.f, oa = 0.f, ob = 0.f;
  if (active) {
    float mxa = -INFINITY, mxb = -INFINITY;
    for (int k2 = 0; k2 < KS; ++k2) {
      mxa = fmaxf(mxa, red[k2 * QT + ra]);
      mxb = fmaxf(mxb, red[k2 * QT + rb]);
    }
    oa = mxa * sl2;  // key 0 is visible to every row: the maximum is finite
    ob = mxb * sl2;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      if (g_lo + gi < g_hi) {
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          float(&a4)[4] = acc[2 * gi + t2];
          a4[0] = ex2_approx(fmaf(a4[0], sl2, -oa));
          a4[1] = ex2_approx(fmaf(a4[1], sl2, -oa));
          a4[2] = ex2_approx(fmaf(a4[2], sl2, -ob));
          a4[3] = ex2_approx(fmaf(a4[3], sl2, -ob));
          sa += a4[0] + a4[1];
          sb += a4[2] + a4[3];
        }
      }
    }
    sa = quad_sum(sa);
    sb = quad_sum(sb);
    if ((lane & 3) == 0) {
      red[(4 + ks) * QT + ra] = sa;
      red[(4 + ks) * QT + rb] = sb;
    }
  }
  cp_async_wait_pending(0);  // this thread's share of V has landed; the barrier publishes everybody's
  __syncthreads();
  // dS is handed to the tensor core in the 16-bit dtype; for fp16 it is pre-scaled by 2^8 (exact) so
  // that products of small probabilities and small gradients stay out of the subnormal range
  constexpr float kDs = Num<T>::dtype == RPO_F16 ? 32.0f : 0.125f;
  constexpr float kUn = Num<T>::dtype == RPO_F16 ? 1.0f / 256.0f : 1.0f;
  float gq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) gq[nt][0] = gq[nt][1] = gq[nt][2] = gq[nt][3] = 0.f;
  if (active) {
    float la = 0.f, lb = 0.f;
    for (int k2 = 0; k2 < KS; ++k2) {
      la += red[(4 + k2) * QT + ra];
      lb += red[(4 + k2) * QT + rb];
    }
    const float ia = 1.0f / la, ib = 1.0f / lb;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      if (g_lo + gi < g_hi) {
        float dp0[4] = {0.f, 0.f, 0.f, 0.f}, dp1[4] = {0.f, 0.f, 0.f, 0.f};
        const int keyn = (g_lo + gi) * 16 + (m >> 1) * 8 + (lane & 7);
#pragma unroll
        for (int kd = 0; kd < 4; ++kd) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4(Vs + swz(keyn, kd * 2 + (m & 1)), b0, b1, b2, b3);
          mma16816<T>(dp0, df[kd], b0, b1);
          mma16816<T>(dp1, df[kd], b2, b3);
        }
        // probabilities rounded through T, as the reference materialises them
        float(&p0)[4] = acc[2 * gi];
        float(&p1)[4] = acc[2 * gi + 1];
        uint32_t a[4];
        a[0] = pack2<T>(rnd<T>(p0[0] * ia) * (dp0[0] - delta_a) * kDs, rnd<T>(p0[1] * ia) * (dp0[1] - delta_a) * kDs);
        a[1] = pack2<T>(rnd<T>(p0[2] * ib) * (dp0[2] - delta_b) * kDs, rnd<T>(p0[3] * ib) * (dp0[3] - delta_b) * kDs);
        a[2] = pack2<T>(rnd<T>(p1[0] * ia) * (dp1[0] - delta_a) * kDs, rnd<T>(p1[1] * ia) * (dp1[1] - delta_a) * kDs);
        a[3] = pack2<T>(rnd<T>(p1[2] * ib) * (dp1[2] - delta_b) * kDs, rnd<T>(p1[3] * ib) * (dp1[3] - delta_b) * kDs);
        const int keyk = (g_lo + gi) * 16 + (m & 1) * 8 + (lane & 7);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_trans(Ks + swz(keyk, dp * 2 + (m >> 1)), b0, b1, b2, b3);
          mma16816<T>(gq[dp * 2], a, b0, b1);
          mma16816<T>(gq[dp * 2 + 1], a, b2, b3);
        }
      }
    }
  }
  __syncthreads();  // every warp is done with K and V: their space takes the f32 dQ partial tiles
  float *part = reinterpret_cast<float *>(sm);  // [4 warps][16 rows][64] f32
  if (active && ks > 0) {
    float *pw = part + warp * (16 * HD);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<float2 *>(pw + (lane >> 2) * HD + nt * 8 + c0) = make_float2(gq[nt][0], gq[nt][1]);
      *reinterpret_cast<float2 *>(pw + ((lane >> 2) + 8) * HD + nt * 8 + c0) = make_float2(gq[nt][2], gq[nt][3]);
    }
  }
  __syncthreads();
  if (active && ks == 0) {
    for (int k2 = 1; k2 < KS; ++k2) {
      const float *pw = part + (warp + k2) * (16 * HD);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 va = *reinterpret_cast<const float2 *>(pw + (lane >> 2) * HD + nt * 8 + c0);
        const float2 vb = *reinterpret_cast<const float2 *>(pw + ((lane >> 2) + 8) * HD + nt * 8 + c0);
        gq[nt][0] += va.x;
        gq[nt][1] += va.y;
        gq[nt][2] += vb.x;
        gq[nt][3] += vb.y;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      gq[nt][0] *= kUn;
      gq[nt][1] *= kUn;
      gq[nt][2] *= kUn;
      gq[nt][3] *= kUn;
    }
    __syncwarp();
    uint8_t *Qs_gen = sm + q_off;
    store_tile<T>(Qs, Qs_gen, r_base, gq, lane, [&](int r) -> T * {
      return r < rows_here ? dq + pbase + (long long)r * D : nullptr;
    });
  }
}

template <typename T, int NG>
static int launch_bwd_ks(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq,
                         const int *ctx_off, int G, int K, int H, int max_ctx, cudaStream_t st) {
  const int n16 = (max_ctx + 15) & ~15;
  const int kv = 2 * n16 * ROW_BYTES;
  const int part = 4 * 16 * HD * 4;  // dQ partial tiles reuse the K/V space
  const int smem = (kv > part ? kv : part) + 2 * QT * ROW_BYTES + 2 * 4 * QT * 4;
  static int configured = 0;
  if (smem > configured) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_bwd_ks<T, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(H, G, (K + QT - 1) / QT);
  prof_tag("attn_bwd G=%d H=%d K=%d max_ctx=%d", G, H, K, max_ctx);
  RPO_CHECK_CUDA(launch_pdl(ro_attn_bwd_ks<T, NG>, grid, dim3(THREADS), smem, st, qkv_ctx, q_prompt, o_prompt, d_out, dq,
                            ctx_off, K, H, g_operands_settled ? 1 : 0));
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

template <typename T, int NT>
static int launch_fwd(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G, int K,
                      int H, int max_ctx, int causal, int do_ctx, cudaStream_t st) {
  const int n16 = (max_ctx + 15) & ~15;
  const int smem = (2 * n16 + QT) * ROW_BYTES;
  static int configured = 0;
  if (smem > configured) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_fwd_mma<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int n_q = (do_ctx ? max_ctx : 0) + K;
  dim3 grid((n_q + QT - 1) / QT, H, G);
  prof_tag("attn_fwd G=%d H=%d K=%d max_ctx=%d do_ctx=%d", G, H, K, max_ctx, do_ctx);
  RPO_CHECK_CUDA(launch_pdl(ro_attn_fwd_mma<T, NT>, grid, dim3(THREADS), smem, st, qkv_ctx, q_prompt, out_ctx, out_prompt,
                            ctx_off, K, H, causal, do_ctx));
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

template <typename T, int NT>
static int launch_bwd(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq,
                      const int *ctx_off, int G, int K, int H, int max_ctx, cudaStream_t st) {
  const int n16 = (max_ctx + 15) & ~15;
  const int smem = (2 * n16 + 2 * QT) * ROW_BYTES;
  static int configured = 0;
  if (smem > configured) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_bwd_mma<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(H, G, (K + QT - 1) / QT);
  prof_tag("attn_bwd G=%d H=%d K=%d max_ctx=%d", G, H, K, max_ctx);
  RPO_CHECK_CUDA(launch_pdl(ro_attn_bwd_mma<T, NT>, grid, dim3(THREADS), smem, st, qkv_ctx, q_prompt, o_prompt, d_out, dq,
                            ctx_off, K, H, g_operands_settled ? 1 : 0));
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

}  // namespace amma

#define AMMA_PICK(FN, ...)                                                  \
  do {                                                                      \
    if (max_ctx <= 32) return amma::FN<T, 4>(__VA_ARGS__);                  \
    if (max_ctx <= 80) return amma::FN<T, 10>(__VA_ARGS__);                 \
    if (max_ctx <= 208) return amma::FN<T, 26>(__VA_ARGS__);                \
    return amma::FN<T, 36>(__VA_ARGS__);                                    \
  } while (0)

template <typename T>
int ro_attention_fwd_mma(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G,
                         int K, int H, int max_ctx, int causal, int do_ctx, cudaStream_t st) {
  RPO_REQUIRE(max_ctx >= 1 && max_ctx <= 288, "at most 288 context rows per group on the tensor-core path");
  RPO_REQUIRE(G <= 65535 && H <= 65535, "grid limits");
  if (G == 0 || (!do_ctx && K == 0)) return RPO_OK;
  static const bool single_pass = [] { const char *e = diag_env("RPO_ATTN_SINGLEPASS"); return e && e[0] == '1'; }();
  if (!single_pass)
    return amma::launch_fwd_flash<T>(qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, G, K, H, max_ctx, causal, do_ctx, st);
  AMMA_PICK(launch_fwd, qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, G, K, H, max_ctx, causal, do_ctx, st);
}

template <typename T>
int ro_attention_bwd_mma(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq,
                         const int *ctx_off, int G, int K, int H, int max_ctx, cudaStream_t st) {
  RPO_REQUIRE(max_ctx >= 1 && max_ctx <= 288, "at most 288 context rows per group on the tensor-core path");
  RPO_REQUIRE(G <= 65535, "grid limits");
  if (G == 0 || K == 0) return RPO_OK;
  static const bool single_pass = [] { const char *e = diag_env("RPO_ATTN_SINGLEPASS"); return e && e[0] == '1'; }();
  if (!single_pass && max_ctx > 64) {  // short contexts (text: ~10 keys) are faster in the single-pass kernel
    // 16-key groups per warp: the groups are shared by KS = 4 / (query tiles, rounded up to 1, 2 or 4) warps
    const int rows = K < amma::QT ? K : amma::QT;
    const int mt = (rows + 15) / 16, slots = mt <= 1 ? 1 : (mt == 2 ? 2 : 4), ks = 4 / slots;
    const int ng = (((max_ctx + 15) / 16) + ks - 1) / ks;
    if (ng <= 1) return amma::launch_bwd_ks<T, 1>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
    if (ng <= 2) return amma::launch_bwd_ks<T, 2>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
    if (ng <= 4) return amma::launch_bwd_ks<T, 4>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
    if (ng <= 7) return amma::launch_bwd_ks<T, 7>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
    if (ng <= 9) return amma::launch_bwd_ks<T, 9>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
  }
  AMMA_PICK(launch_bwd, qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
}

#define INSTANTIATE(T)                                                                                            \
  template int ro_attention_fwd_mma<T>(const T *, const T *, T *, T *, const int *, int, int, int, int, int, int, \
                                       cudaStream_t);                                                             \
  template int ro_attention_bwd_mma<T>(const T *, const T *, const T *, const T *, T *, const int *, int, int,    \
                                       int, int, cudaStream_t);
INSTANTIATE(__half)
INSTANTIATE(__nv_bfloat16)

}  // namespace rpo
