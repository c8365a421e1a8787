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
                    const T *__restrict__ d_out, T *__restrict__ dq, const int *__restrict__ ctx_off, int K, int H) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int g = blockIdx.y, h = blockIdx.x;
  const int D = H * HD;
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
  stage_rows<T>(dOs, d_out + pbase, D, rows_here, QT);
  cp_async_wait_all();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_base = warp * 16;
  if (r_base >= rows_here) return;
  // delta_r = sum_d dO[r,d] * O[r,d]; two lanes per row, 32 columns each
  float delta_a, delta_b;
  {
    const int r = r_base + (lane >> 1);
    float s = 0.f;
    if (r < rows_here) {
      const T *po = o_prompt + pbase + (long long)r * D + (lane & 1) * 32;
      const T *pd = d_out + pbase + (long long)r * D + (lane & 1) * 32;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        Vec16<T> a = ld16(po + v * 8), b = ld16(pd + v * 8);
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
  ro_attn_fwd_mma<T, NT><<<grid, THREADS, smem, st>>>(qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, K, H, causal,
                                                      do_ctx);
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
  ro_attn_bwd_mma<T, NT><<<grid, THREADS, smem, st>>>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, K, H);
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
  AMMA_PICK(launch_fwd, qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, G, K, H, max_ctx, causal, do_ctx, st);
}

template <typename T>
int ro_attention_bwd_mma(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq,
                         const int *ctx_off, int G, int K, int H, int max_ctx, cudaStream_t st) {
  RPO_REQUIRE(max_ctx >= 1 && max_ctx <= 288, "at most 288 context rows per group on the tensor-core path");
  RPO_REQUIRE(G <= 65535, "grid limits");
  if (G == 0 || K == 0) return RPO_OK;
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
