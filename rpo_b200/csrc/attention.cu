// Read-only masked multi-head attention (clip/model.py:186 under the masks of
// trainers/rpo.py:140-159), forward and the prompt-query gradient.
//
// The reference's additive masks put -inf on every prompt COLUMN, for every row, so a prompt is
// never a key or a value.  The mask is therefore not materialised here: a group
// (image / class) has n_g context rows that are keys, values and (optionally) queries, plus K
// prompt rows that are queries only.  Context queries read keys j < n_g (vision) or j <= r
// (causal, text); prompt queries read all n_g keys.
//
// This file holds the exact-f32 SIMT kernels (one warp per query row, keys/values of one
// (group, head) staged in shared memory).  They serve RPO_F32, where tensor cores (tf32) would
// break the 1e-5 parity bar, and are the cross-check for the tensor-core path in attention_mma.cu.
#include <stdlib.h>

#include "common.cuh"

namespace rpo {

static constexpr int HD = 64;           // head dim of every CLIP model
static constexpr int KV_LD = HD + 1;    // padded f32 row stride in shared memory (conflict-free)
static constexpr int ATT_WARPS = 8;
static constexpr int MAX_KT = 10;       // keys per lane: supports up to 320 context rows
static constexpr int Q_CHUNK = 32;      // query rows per block

template <typename T>
__device__ __forceinline__ void load_kv_smem(const T *__restrict__ qkv_ctx, int row0, int n, int D, int h,
                                             float *Ks, float *Vs) {
  // qkv_ctx row layout: [ q(D) | k(D) | v(D) ], head h at columns h*64..h*64+63 of each part
  for (int idx = threadIdx.x; idx < n * (HD / 2); idx += blockDim.x) {
    int j = idx / (HD / 2);
    int d = (idx % (HD / 2)) * 2;
    const T *kp = qkv_ctx + (size_t)(row0 + j) * 3 * D + D + h * HD + d;
    const T *vp = kp + D;
    Ks[j * KV_LD + d] = tof<T>(kp[0]);
    Ks[j * KV_LD + d + 1] = tof<T>(kp[1]);
    Vs[j * KV_LD + d] = tof<T>(vp[0]);
    Vs[j * KV_LD + d + 1] = tof<T>(vp[1]);
  }
}

template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
    ro_attn_fwd_simt(const T *__restrict__ qkv_ctx, const T *__restrict__ q_prompt, T *__restrict__ out_ctx,
                     T *__restrict__ out_prompt, const int *__restrict__ ctx_off, int K, int H, int causal,
                     int do_ctx) {
  extern __shared__ float sm[];
  const int g = blockIdx.z, h = blockIdx.y;
  const int D = H * HD;
  const int row0 = ctx_off[g];
  const int n = ctx_off[g + 1] - row0;
  const int n_q = (do_ctx ? n : 0) + K;  // query rows of this group: context rows first, then prompts
  const int q_begin = blockIdx.x * Q_CHUNK;
  if (q_begin >= n_q) return;
  const int q_end = min(n_q, q_begin + Q_CHUNK);
  float *Ks = sm;
  float *Vs = Ks + n * KV_LD;
  float *wbuf = Vs + n * KV_LD;  // per warp: q[64] + p[MAX_KT*32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *qs = wbuf + warp * (HD + MAX_KT * 32);
  float *ps = qs + HD;
  load_kv_smem(qkv_ctx, row0, n, D, h, Ks, Vs);
  __syncthreads();

  for (int qi = q_begin + warp; qi < q_end; qi += ATT_WARPS) {
    const bool is_ctx = do_ctx && qi < n;
    const int pi = qi - (do_ctx ? n : 0);
    const T *qp = is_ctx ? qkv_ctx + (size_t)(row0 + qi) * 3 * D + h * HD
                         : q_prompt + ((size_t)g * K + pi) * D + h * HD;
    qs[lane] = tof<T>(qp[lane]);
    qs[lane + 32] = tof<T>(qp[lane + 32]);
    __syncwarp();
    const int n_vis = (is_ctx && causal) ? min(n, qi + 1) : n;  // readable keys: j < n_vis
    float sc[MAX_KT];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      int j = lane + 32 * t;
      float s = -INFINITY;
      if (j < n_vis) {
        const float *kr = Ks + j * KV_LD;
        float a = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; ++d) a = fmaf(qs[d], kr[d], a);
        s = a * 0.125f;  // 1/sqrt(64)
      }
      sc[t] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      float e = (lane + 32 * t < n_vis) ? expf(sc[t] - mx) : 0.f;
      sc[t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      int j = lane + 32 * t;
      if (j < n_vis) ps[j] = rnd<T>(sc[t] * inv);  // probabilities are a dtype tensor in the reference
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < n_vis; ++j) {
      float p = ps[j];
      o0 = fmaf(p, Vs[j * KV_LD + lane], o0);
      o1 = fmaf(p, Vs[j * KV_LD + lane + 32], o1);
    }
    T *op = is_ctx ? out_ctx + (size_t)(row0 + qi) * D + h * HD : out_prompt + ((size_t)g * K + pi) * D + h * HD;
    op[lane] = fromf<T>(o0);
    op[lane + 32] = fromf<T>(o1);
    __syncwarp();
  }
}

// dq for the prompt queries:  p = softmax(q k^T / 8);  dp_j = dO . v_j;
// ds_j = p_j (dp_j - sum_i p_i dp_i);  dq = (1/8) sum_j ds_j k_j.
template <typename T>
__global__ void __launch_bounds__(ATT_WARPS * 32)
    ro_attn_bwd_simt(const T *__restrict__ qkv_ctx, const T *__restrict__ q_prompt, const T *__restrict__ d_out,
                     T *__restrict__ dq, const int *__restrict__ ctx_off, int K, int H) {
  extern __shared__ float sm[];
  const int g = blockIdx.y, h = blockIdx.x;
  const int D = H * HD;
  const int row0 = ctx_off[g];
  const int n = ctx_off[g + 1] - row0;
  float *Ks = sm;
  float *Vs = Ks + n * KV_LD;
  float *wbuf = Vs + n * KV_LD;  // per warp: q[64] + dO[64] + ds[MAX_KT*32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *qs = wbuf + warp * (2 * HD + MAX_KT * 32);
  float *dos = qs + HD;
  float *ds = dos + HD;
  load_kv_smem(qkv_ctx, row0, n, D, h, Ks, Vs);
  __syncthreads();
  for (int pi = warp; pi < K; pi += ATT_WARPS) {
    const size_t ro = ((size_t)g * K + pi) * D + h * HD;
    qs[lane] = tof<T>(q_prompt[ro + lane]);
    qs[lane + 32] = tof<T>(q_prompt[ro + lane + 32]);
    dos[lane] = tof<T>(d_out[ro + lane]);
    dos[lane + 32] = tof<T>(d_out[ro + lane + 32]);
    __syncwarp();
    float sc[MAX_KT], dp[MAX_KT];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      int j = lane + 32 * t;
      float s = -INFINITY, dpj = 0.f;
      if (j < n) {
        const float *kr = Ks + j * KV_LD;
        const float *vr = Vs + j * KV_LD;
        float a = 0.f;
#pragma unroll 16
        for (int d = 0; d < HD; ++d) {
          a = fmaf(qs[d], kr[d], a);
          dpj = fmaf(dos[d], vr[d], dpj);
        }
        s = a * 0.125f;
      }
      sc[t] = s;
      dp[t] = dpj;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      float e = (lane + 32 * t < n) ? expf(sc[t] - mx) : 0.f;
      sc[t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float delta = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      sc[t] *= inv;
      delta += sc[t] * dp[t];
    }
    delta = warp_sum(delta);
#pragma unroll
    for (int t = 0; t < MAX_KT; ++t) {
      int j = lane + 32 * t;
      if (j < n) ds[j] = sc[t] * (dp[t] - delta) * 0.125f;
    }
    __syncwarp();
    float g0 = 0.f, g1 = 0.f;
    for (int j = 0; j < n; ++j) {
      float w = ds[j];
      g0 = fmaf(w, Ks[j * KV_LD + lane], g0);
      g1 = fmaf(w, Ks[j * KV_LD + lane + 32], g1);
    }
    dq[ro + lane] = fromf<T>(g0);
    dq[ro + lane + 32] = fromf<T>(g1);
    __syncwarp();
  }
}

template <typename T>
int ro_attention_fwd_simt(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G,
                          int K, int H, int max_ctx, int causal, int do_ctx, cudaStream_t st) {
  RPO_REQUIRE(max_ctx >= 1 && max_ctx <= MAX_KT * 32, "at most 320 context rows per group");
  RPO_REQUIRE(G <= 65535 && H <= 65535, "grid limits");
  if (G == 0) return RPO_OK;
  size_t smem = sizeof(float) * ((size_t)2 * max_ctx * KV_LD + ATT_WARPS * (HD + MAX_KT * 32));
  static size_t configured = 0;
  if (smem > configured) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_fwd_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int n_q = (do_ctx ? max_ctx : 0) + K;
  dim3 grid((n_q + Q_CHUNK - 1) / Q_CHUNK, H, G);
  ro_attn_fwd_simt<T><<<grid, ATT_WARPS * 32, smem, st>>>(qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, K, H,
                                                          causal, do_ctx);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

template <typename T>
int ro_attention_bwd_simt(const T *qkv_ctx, const T *q_prompt, const T *d_out, T *dq, const int *ctx_off, int G, int K,
                          int H, int max_ctx, cudaStream_t st) {
  RPO_REQUIRE(max_ctx >= 1 && max_ctx <= MAX_KT * 32, "at most 320 context rows per group");
  RPO_REQUIRE(G <= 65535, "grid limits");
  if (G == 0) return RPO_OK;
  size_t smem = sizeof(float) * ((size_t)2 * max_ctx * KV_LD + ATT_WARPS * (2 * HD + MAX_KT * 32));
  static size_t configured = 0;
  if (smem > configured) {
    RPO_CHECK_CUDA(cudaFuncSetAttribute(ro_attn_bwd_simt<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(H, G);
  ro_attn_bwd_simt<T><<<grid, ATT_WARPS * 32, smem, st>>>(qkv_ctx, q_prompt, d_out, dq, ctx_off, K, H);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// tensor-core kernels (attention_mma.cu), 16-bit dtypes only
template <typename T>
int ro_attention_fwd_mma(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G,
                         int K, int H, int max_ctx, int causal, int do_ctx, cudaStream_t st);
template <typename T>
int ro_attention_bwd_mma(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq,
                         const int *ctx_off, int G, int K, int H, int max_ctx, cudaStream_t st);

// RPO_ATTN_SIMT=1 (diagnostics build) forces the exact-f32 SIMT kernels for every dtype
static bool force_simt() {
  const char *e = diag_env("RPO_ATTN_SIMT");
  return e && e[0] == '1';
}

template <typename T>
int ro_attention_fwd(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G, int K,
                     int H, int max_ctx, int causal, int do_ctx, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    if (max_ctx <= 288 && !force_simt())
      return ro_attention_fwd_mma<T>(qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, G, K, H, max_ctx, causal, do_ctx,
                                     st);
  }
  return ro_attention_fwd_simt<T>(qkv_ctx, q_prompt, out_ctx, out_prompt, ctx_off, G, K, H, max_ctx, causal, do_ctx,
                                  st);
}

template <typename T>
int ro_attention_bwd(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq, const int *ctx_off,
                     int G, int K, int H, int max_ctx, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    if (max_ctx <= 288 && !force_simt())
      return ro_attention_bwd_mma<T>(qkv_ctx, q_prompt, o_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
  }
  return ro_attention_bwd_simt<T>(qkv_ctx, q_prompt, d_out, dq, ctx_off, G, K, H, max_ctx, st);
}

#define INSTANTIATE(T)                                                                                           \
  template int ro_attention_fwd<T>(const T *, const T *, T *, T *, const int *, int, int, int, int, int, int,    \
                                   cudaStream_t);                                                                \
  template int ro_attention_bwd<T>(const T *, const T *, const T *, const T *, T *, const int *, int, int, int,  \
                                   int, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(__half)
INSTANTIATE(__nv_bfloat16)

}  // namespace rpo
