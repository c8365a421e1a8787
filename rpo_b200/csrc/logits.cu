// Logit block of CustomCLIP.forward (trainers/rpo.py:215-230) and its backward:
//   text_f, img_f <- row-wise L2 normalise                                   (:215, :218)
//   logits[b,c] = (1/K) sum_k  dtype( (dtype(exp(s)) * img_f[b,k,:]) . text_f[c,k,:] )   (:221-227)
//   loss = mean_b CE(logits[b,:], label_b)                                   (:230)
// The K per-pair products are K small GEMMs [B,E]x[E,C]: one batched tcgen05 launch (logits_tc.cu) for the 16-bit
// types, the strided exact-f32 SIMT GEMM for RPO_F32 and for shapes the tensor-core kernel does not take.  Each
// pair's product is rounded to the dtype, the pair sum is accumulated in f32 exactly like the reference's f32
// `logits` buffer.
#include "common.cuh"

namespace rpo {

// one warp per row: norm (rounded through T like the reference's dtype `norm` tensor), x/norm, and
// optionally the exp(logit_scale)-scaled copy used on the image side.
template <typename T>
__global__ void l2norm_fwd_kernel(const T *__restrict__ x, T *__restrict__ xn, T *__restrict__ xs,
                                  float *__restrict__ norm_out, const float *__restrict__ logit_scale, long long rows,
                                  int E) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T *xr = x + row * E;
  float ss = 0.f;
  for (int c = lane; c < E; c += 32) {
    float v = tof<T>(xr[c]);
    ss += v * v;
  }
  float norm = rnd<T>(sqrtf(warp_sum(ss)));
  if (lane == 0) norm_out[row] = norm;
  float s16 = xs ? rnd<T>(expf(*logit_scale)) : 0.f;
  for (int c = lane; c < E; c += 32) {
    float v = rnd<T>(tof<T>(xr[c]) / norm);
    xn[row * E + c] = fromf<T>(v);
    if (xs) xs[row * E + c] = fromf<T>(s16 * v);
  }
}

// dx = (g - xn * (xn . g)) / norm,   g = gscale * dxn    (gscale = dtype(exp(s)) on the image side)
template <typename T>
__global__ void l2norm_bwd_kernel(const T *__restrict__ dxn, const T *__restrict__ xn, const float *__restrict__ norm,
                                  const float *__restrict__ logit_scale, T *__restrict__ dx, long long rows, int E) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float gs = logit_scale ? rnd<T>(expf(*logit_scale)) : 1.0f;
  float dot = 0.f;
  for (int c = lane; c < E; c += 32) dot += rnd<T>(gs * tof<T>(dxn[row * E + c])) * tof<T>(xn[row * E + c]);
  dot = warp_sum(dot);
  float inv = 1.0f / norm[row];
  for (int c = lane; c < E; c += 32) {
    float g = rnd<T>(gs * tof<T>(dxn[row * E + c]));
    dx[row * E + c] = fromf<T>((g - tof<T>(xn[row * E + c]) * dot) * inv);
  }
}

// one block per image: f32 pair sum, /K, and (label != null) the CE row loss and d loss/d logits
template <typename T>
__global__ void __launch_bounds__(256) logits_reduce_ce_kernel(const T *__restrict__ pair, const int64_t *__restrict__ label,
                                                               float *__restrict__ logits_out, float *__restrict__ row_loss,
                                                               float *__restrict__ dlogits, float *__restrict__ lg_scratch,
                                                               int B, int C, int K) {
  __shared__ float red[32];
  __shared__ float bcast;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *lg = lg_scratch + (size_t)b * C;
  float mx = -INFINITY;
  for (int c = tid; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += tof<T>(pair[((size_t)k * B + b) * C + c]);
    acc /= (float)K;
    lg[c] = acc;
    if (logits_out) logits_out[(size_t)b * C + c] = acc;
    mx = fmaxf(mx, acc);
  }
  if (!label) return;
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : -INFINITY;
    v = warp_max(v);
    if (lane == 0) bcast = v;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int c = tid; c < C; c += blockDim.x) sum += expf(lg[c] - mx);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) bcast = v;
  }
  __syncthreads();
  sum = bcast;
  const int y = (int)label[b];
  const float lse = mx + logf(sum);
  const float ly = lg[y];
  __syncthreads();  // lg may alias dlogits: everyone has read lg[y] before it is overwritten below
  if (tid == 0 && row_loss) row_loss[b] = lse - ly;
  if (dlogits) {
    const float invB = 1.0f / (float)B;
    for (int c = tid; c < C; c += blockDim.x) {
      float p = expf(lg[c] - lse);
      dlogits[(size_t)b * C + c] = (p - (c == y ? 1.f : 0.f)) * invB;
    }
  }
}

__global__ void mean_rows_kernel(const float *__restrict__ row_loss, float *__restrict__ loss, int B) {
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += 32) s += row_loss[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *loss = s / (float)B;
}

template <typename T>
__global__ void scale_cast_kernel(const float *__restrict__ src, T *__restrict__ dst, long long n, float scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = fromf<T>(src[i] * scale);
}

template <typename T>
int logits_ce_fwd(const T *img_feat, const T *text_feat, const float *logit_scale, const int64_t *label, int B, int C,
                  int K, int E, T *img_n, T *img_s, T *text_n, float *img_norm, float *text_norm, T *pair_logits,
                  float *logits, float *loss, float *dlogits, cudaStream_t st) {
  RPO_REQUIRE(B > 0 && C > 0 && K > 0 && E > 0, "logit block shape");
  RPO_REQUIRE(!loss || label, "loss needs labels");
  // dlogits doubles as the f32 logit scratch row when the caller does not want logits
  long long rows_i = (long long)B * K, rows_t = (long long)C * K;
  l2norm_fwd_kernel<T><<<(unsigned)((rows_i + 7) / 8), 256, 0, st>>>(img_feat, img_n, img_s, img_norm, logit_scale,
                                                                      rows_i, E);
  RPO_LAUNCH_CHECK();
  l2norm_fwd_kernel<T><<<(unsigned)((rows_t + 7) / 8), 256, 0, st>>>(text_feat, text_n, nullptr, text_norm, nullptr,
                                                                      rows_t, E);
  RPO_LAUNCH_CHECK();
  if constexpr (sizeof(T) == 2) {
    if (logits_tc_supported(Num<T>::dtype, B, C, K, E)) {
      RPO_TRY(logits_pair_fwd_tc<T>(img_s, text_n, pair_logits, B, C, K, E, st));
    } else {
      Epilogue<T> ep{};
      RPO_TRY(gemm_simt<T>(img_s, (long long)K * E, 1, text_n, (long long)K * E, 1, pair_logits, C, B, C, E, ep, K, E,
                           E, (long long)B * C, st));
    }
  } else {
    Epilogue<T> ep{};
    RPO_TRY(gemm_simt<T>(img_s, (long long)K * E, 1, text_n, (long long)K * E, 1, pair_logits, C, B, C, E, ep, K, E, E,
                         (long long)B * C, st));
  }
  // scratch for the f32 logits row: reuse `logits` if given, else `dlogits`
  float *scratch = logits ? logits : dlogits;
  RPO_REQUIRE(scratch != nullptr, "need logits or dlogits as f32 scratch");
  // row losses are staged at the tail of img_norm? no: keep them in text_norm's tail-free dedicated
  // area -- the caller sizes img_norm as [B*K + B] so that row losses live after the norms.
  float *row_loss = img_norm + rows_i;
  logits_reduce_ce_kernel<T><<<B, 256, 0, st>>>(pair_logits, label, logits, label ? row_loss : nullptr,
                                                label ? dlogits : nullptr, scratch, B, C, K);
  RPO_LAUNCH_CHECK();
  if (loss) {
    mean_rows_kernel<<<1, 32, 0, st>>>(row_loss, loss, B);
    RPO_LAUNCH_CHECK();
  }
  return RPO_OK;
}

template <typename T>
int logits_ce_bwd(const float *dlogits, const T *img_feat, const T *text_feat, const T *img_n, const T *img_s,
                  const T *text_n, const float *img_norm, const float *text_norm, const float *logit_scale, int B,
                  int C, int K, int E, T *dl_t, T *d_img_s, T *d_text_n, T *d_img_feat, T *d_text_feat,
                  float grad_scale, cudaStream_t st) {
  (void)img_feat;
  (void)text_feat;
  long long n = (long long)B * C;
  // `logits /= K` then the dtype cast of the gradient flowing into each per-pair GEMM output
  scale_cast_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dlogits, dl_t, n, grad_scale / (float)K);
  RPO_LAUNCH_CHECK();
  bool on_tc = false;
  if constexpr (sizeof(T) == 2) {
    if (logits_tc_supported(Num<T>::dtype, B, C, K, E)) {
      RPO_TRY(logits_pair_bwd_tc<T>(dl_t, img_s, text_n, d_img_s, d_text_n, B, C, K, E, st));
      on_tc = true;
    }
  }
  if (!on_tc) {
    Epilogue<T> ep{};
    // d img_s[b,k,:] = sum_c dl[b,c] text_n[c,k,:]
    RPO_TRY(gemm_simt<T>(dl_t, C, 1, text_n, 1, (long long)K * E, d_img_s, (long long)K * E, B, E, C, ep, K, 0, E, E,
                         st));
    // d text_n[c,k,:] = sum_b dl[b,c] img_s[b,k,:]
    RPO_TRY(gemm_simt<T>(dl_t, 1, C, img_s, 1, (long long)K * E, d_text_n, (long long)K * E, C, E, B, ep, K, 0, E, E,
                         st));
  }
  long long rows_i = (long long)B * K, rows_t = (long long)C * K;
  l2norm_bwd_kernel<T><<<(unsigned)((rows_i + 7) / 8), 256, 0, st>>>(d_img_s, img_n, img_norm, logit_scale, d_img_feat,
                                                                      rows_i, E);
  RPO_LAUNCH_CHECK();
  l2norm_bwd_kernel<T><<<(unsigned)((rows_t + 7) / 8), 256, 0, st>>>(d_text_n, text_n, text_norm, nullptr, d_text_feat,
                                                                      rows_t, E);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

#define INSTANTIATE(T)                                                                                              \
  template int logits_ce_fwd<T>(const T *, const T *, const float *, const int64_t *, int, int, int, int, T *, T *, \
                                T *, float *, float *, T *, float *, float *, float *, cudaStream_t);               \
  template int logits_ce_bwd<T>(const float *, const T *, const T *, const T *, const T *, const T *,               \
                                const float *, const float *, const float *, int, int, int, int, T *, T *, T *,     \
                                T *, T *, float, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(__half)
INSTANTIATE(__nv_bfloat16)

}  // namespace rpo
