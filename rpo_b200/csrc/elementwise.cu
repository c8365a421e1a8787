// HBM-streaming kernels of the path: LayerNorm fwd/bwd (clip/model.py:153-159), patch extraction,
// embedding assembly (trainers/rpo.py:198-204), prompt broadcast / gradient reduction, SGD.
// All are bandwidth-bound: one warp per row, 16-byte vector loads, warp-shuffle reductions.
#include "common.cuh"

namespace rpo {

static constexpr int LN_WARPS = 8;
static constexpr float LN_EPS = 1e-5f;

// Loads one row (D elements of T) spread over a warp into registers as f32.
// Lane l owns vectors l, l+32, ... ; NV vectors of VEC elements each (D <= 32*NV*VEC).
template <typename T, int NV>
__device__ __forceinline__ void load_row(const T *row, int nv, int lane, float (&vals)[NV * Vec16<T>::N]) {
  constexpr int VEC = Vec16<T>::N;
  Vec16<T> v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)  // all loads of the row in flight before the first conversion
    if (lane + 32 * i < nv) v[i] = ld16(row + (size_t)(lane + 32 * i) * VEC);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    bool in = lane + 32 * i < nv;
#pragma unroll
    for (int e = 0; e < VEC; ++e) vals[i * VEC + e] = in ? tof<T>(v[i].v[e]) : 0.f;
  }
}

// N consecutive f32 parameters with 16-byte loads; p is 16-byte aligned
template <int N>
__device__ __forceinline__ void load_f32(const float *__restrict__ p, float (&out)[N]) {
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(p) + i);
    out[4 * i] = v.x;
    out[4 * i + 1] = v.y;
    out[4 * i + 2] = v.z;
    out[4 * i + 3] = v.w;
  }
}

// LayerNorm affine parameters staged in shared memory, 16-byte piece j of vector vi at [j * nv + vi] (consecutive
// lanes read consecutive float4: conflict-free).  They are frozen CLIP weights, written by no kernel of the step, so
// the block copies them BEFORE pdl_wait(): the copy overlaps the tail of the upstream kernel, and afterwards nobody
// holds them in registers across the row reductions (the f32 registers of w and b were what capped the old kernel
// at 16 warps per SM).
template <int VEC>
__device__ __forceinline__ void stage_param(const float *__restrict__ p, float4 *sp, int nv) {
  constexpr int Q = VEC / 4;
  for (int i = threadIdx.x; i < nv * Q; i += blockDim.x) {
    int vi = i / Q, j = i % Q;
    sp[j * nv + vi] = __ldg(reinterpret_cast<const float4 *>(p) + i);
  }
}
template <int VEC>
__device__ __forceinline__ void read_param(const float4 *sp, int nv, int vi, float (&out)[VEC]) {
#pragma unroll
  for (int j = 0; j < VEC / 4; ++j) {
    float4 v = sp[j * nv + vi];
    out[4 * j] = v.x;
    out[4 * j + 1] = v.y;
    out[4 * j + 2] = v.z;
    out[4 * j + 3] = v.w;
  }
}

template <typename T, int NV>
__device__ __forceinline__ void row_stats(const float (&vals)[NV * Vec16<T>::N], int nv, int lane, int D,
                                          float &mean, float &rstd) {
  constexpr int VEC = Vec16<T>::N;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV * VEC; ++i) s += vals[i];  // padding entries are zero
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + 32 * i < nv) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float d = vals[i * VEC + e] - mean;
        q += d * d;
      }
    }
  }
  float var = warp_sum(q) / (float)D;
  rstd = 1.0f / sqrtf(var + LN_EPS);
}

// resident blocks per SM the register budget is held to: 4 x 8 warps while a lane owns at most 4 vectors of the row
// (every width of the path: 512 / 768 / 1024 in 16-bit types), 2 for the wide f32 rows of the unit tests
constexpr int ln_min_blocks(int NV) { return NV <= 4 ? 4 : 2; }

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, ln_min_blocks(NV))
    ln_fwd_kernel(const T *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b, T *__restrict__ y,
                  long long rows, int D) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float4 sw[1024 / 4], sb[1024 / 4];
  long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  int nv = D / VEC;
  stage_param<VEC>(w, sw, nv);
  stage_param<VEC>(b, sb, nv);
  pdl_wait();
  pdl_trigger();
  float vals[NV * VEC];
  if (row < rows) load_row<T, NV>(x + row * D, nv, lane, vals);
  __syncthreads();  // sw / sb complete (after the row loads are issued)
  if (row >= rows) return;
  float mean, rstd;
  row_stats<T, NV>(vals, nv, lane, D, mean, rstd);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nv) {
      float wv[VEC], bv[VEC];
      read_param<VEC>(sw, nv, vi, wv);
      read_param<VEC>(sb, nv, vi, bv);
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < VEC; ++e) o.v[e] = fromf<T>((vals[i * VEC + e] - mean) * rstd * wv[e] + bv[e]);
      st16(y + row * D + (size_t)vi * VEC, o);
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy * w,  xhat = (x-mean)*rstd
// settled (common.cuh, SettledOperands): x is a row the forward pass left behind, so it is loaded and its statistics
// are reduced BEFORE the dependency wait, while the upstream GEMM still produces dy -- same arithmetic, two of the four
// reductions and one memory round trip off the chain.
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, NV <= 3 ? 3 : (NV <= 4 ? 2 : 1))
    ln_bwd_kernel(const T *__restrict__ dy, const T *__restrict__ x, const float *__restrict__ w,
                  const T *__restrict__ dres, T *__restrict__ dx, long long rows, int D, int settled) {
  constexpr int VEC = Vec16<T>::N;
  __shared__ float4 sw[1024 / 4];
  const long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nv = D / VEC;
  const bool live = row < rows;  // warp-uniform
  stage_param<VEC>(w, sw, nv);
  float xv[NV * VEC], gv[NV * VEC];
  float mean = 0.f, rstd = 0.f;
  if (settled) {
    if (live) {
      load_row<T, NV>(x + row * D, nv, lane, xv);
      row_stats<T, NV>(xv, nv, lane, D, mean, rstd);
    }
    pdl_wait();
    pdl_trigger();
  } else {
    pdl_wait();
    pdl_trigger();
    if (live) load_row<T, NV>(x + row * D, nv, lane, xv);
  }
  Vec16<T> r[NV];
  if (live) {
    load_row<T, NV>(dy + row * D, nv, lane, gv);
    // the residual gradient is needed last but depends on nothing: its load goes out with the others instead of
    // adding one more memory round trip after the reductions (the prompt-row launches are latency chains)
    if (dres) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + 32 * i < nv) r[i] = ld16(dres + row * D + (size_t)(lane + 32 * i) * VEC);
    }
  }
  __syncthreads();  // sw complete
  if (!live) return;
  // g = dy * w does not depend on the statistics
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nv) {
      float wv[VEC];
      read_param<VEC>(sw, nv, vi, wv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) gv[i * VEC + e] *= wv[e];
    }
  }
  if (!settled) row_stats<T, NV>(xv, nv, lane, D, mean, rstd);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nv) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float g = gv[i * VEC + e];
        float xh = (xv[i * VEC + e] - mean) * rstd;
        xv[i * VEC + e] = xh;
        sg += g;
        sgx += g * xh;
      }
    }
  }
  sg = warp_sum(sg) / (float)D;
  sgx = warp_sum(sgx) / (float)D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int vi = lane + 32 * i;
    if (vi < nv) {
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float v = rstd * (gv[i * VEC + e] - sg - xv[i * VEC + e] * sgx);
        if (dres) v += tof<T>(r[i].v[e]);
        o.v[e] = fromf<T>(v);
      }
      st16(dx + row * D + (size_t)vi * VEC, o);
    }
  }
}

static bool ln_shape_ok(int D, size_t esz) { return D > 0 && D <= 1024 && (D * esz) % 16 == 0; }

// vectors of the row per lane, rounded up to an instantiated count
template <typename T>
static int ln_vectors_per_lane(int D) {
  int nv = D / Vec16<T>::N, per = (nv + 31) / 32;
  return per <= 4 ? per : 8;
}

template <typename T>
int layernorm_fwd(const T *x, const float *w, const float *b, T *y, long long rows, int D, cudaStream_t st) {
  RPO_REQUIRE(ln_shape_ok(D, sizeof(T)), "LayerNorm width must be <= 1024 and a multiple of 16 bytes");
  RPO_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)w | (uintptr_t)b) & 15) == 0, "LayerNorm buffers must be 16-byte aligned");
  if (rows <= 0) return RPO_OK;
  dim3 grid((unsigned)((rows + LN_WARPS - 1) / LN_WARPS)), block(LN_WARPS * 32);
  prof_tag("ln_fwd rows=%lld D=%d", rows, D);
  cudaError_t e = cudaErrorInvalidValue;
  switch (ln_vectors_per_lane<T>(D)) {
    case 1: e = launch_pdl(ln_fwd_kernel<T, 1>, grid, block, 0, st, x, w, b, y, rows, D); break;
    case 2: e = launch_pdl(ln_fwd_kernel<T, 2>, grid, block, 0, st, x, w, b, y, rows, D); break;
    case 3: e = launch_pdl(ln_fwd_kernel<T, 3>, grid, block, 0, st, x, w, b, y, rows, D); break;
    case 4: e = launch_pdl(ln_fwd_kernel<T, 4>, grid, block, 0, st, x, w, b, y, rows, D); break;
    default:
      if constexpr (sizeof(T) == 4) e = launch_pdl(ln_fwd_kernel<T, 8>, grid, block, 0, st, x, w, b, y, rows, D);
  }
  RPO_CHECK_CUDA(e);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}
template <typename T>
int layernorm_bwd(const T *dy, const T *x, const float *w, const T *dres, T *dx, long long rows, int D,
                  cudaStream_t st) {
  RPO_REQUIRE(ln_shape_ok(D, sizeof(T)), "LayerNorm width must be <= 1024 and a multiple of 16 bytes");
  RPO_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)w) & 15) == 0, "LayerNorm buffers must be 16-byte aligned");
  if (rows <= 0) return RPO_OK;
  dim3 grid((unsigned)((rows + LN_WARPS - 1) / LN_WARPS)), block(LN_WARPS * 32);
  prof_tag("ln_bwd rows=%lld D=%d", rows, D);
  cudaError_t e = cudaErrorInvalidValue;
  const int settled = g_operands_settled ? 1 : 0;
  switch (ln_vectors_per_lane<T>(D)) {
    case 1: e = launch_pdl(ln_bwd_kernel<T, 1>, grid, block, 0, st, dy, x, w, dres, dx, rows, D, settled); break;
    case 2: e = launch_pdl(ln_bwd_kernel<T, 2>, grid, block, 0, st, dy, x, w, dres, dx, rows, D, settled); break;
    case 3: e = launch_pdl(ln_bwd_kernel<T, 3>, grid, block, 0, st, dy, x, w, dres, dx, rows, D, settled); break;
    case 4: e = launch_pdl(ln_bwd_kernel<T, 4>, grid, block, 0, st, dy, x, w, dres, dx, rows, D, settled); break;
    default:
      if constexpr (sizeof(T) == 4)
        e = launch_pdl(ln_bwd_kernel<T, 8>, grid, block, 0, st, dy, x, w, dres, dx, rows, D, settled);
  }
  RPO_CHECK_CUDA(e);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// ---- patch extraction: conv1 with stride == kernel (clip/model.py:215) is a GEMM over patches ----
// out[(b*NP + py*G + px), c*P*P + ky*P + kx] = T(image[b, c, py*P+ky, px*P+kx])
// uint8 pixels: ToTensor (x / 255) and Normalize ((x - mean[c]) / std[c]) of clip/clip.py:75-78 in f32, operation for
// operation (true divisions, no fused multiply-add), then the cast to the model dtype of trainers/rpo.py:198 -- the same
// value the reference's data pipeline + `image.type(self.dtype)` produce, from a quarter of the bytes over PCIe.
__device__ __forceinline__ float pixel_norm(uint8_t v, float mean, float std) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean), std);
}

template <typename T, typename TI>
__global__ void im2col_kernel(const TI *__restrict__ img, T *__restrict__ out, int B, int res, int P, int ld,
                              long long total, PixelNorm nm) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int G = res / P;
  int col = (int)(idx % ld);
  long long row = idx / ld;
  if (col >= 3 * P * P) {  // zero padding up to the GEMM K tile
    out[idx] = fromf<T>(0.f);
    return;
  }
  int kx = col % P, ky = (col / P) % P, c = col / (P * P);
  int p = (int)(row % (G * G));
  int b = (int)(row / (G * G));
  int py = p / G, px = p % G;
  const TI raw = img[(((size_t)b * 3 + c) * res + (py * P + ky)) * res + (px * P + kx)];
  float v;
  if constexpr (sizeof(TI) == 1)
    v = pixel_norm(raw, nm.mean[c], nm.std[c]);
  else
    v = Num<TI>::to_f(raw);
  out[idx] = fromf<T>(v);
}

// P % 8 == 0 (ViT-B/16, B/32) and no K padding: one thread moves one run of 8 pixels -- 32 (f32) or 16 (16-bit)
// contiguous bytes in, 16 contiguous bytes out; consecutive threads write consecutive 16-byte chunks of a
// patch row, so stores are fully coalesced and loads are whole sectors.
template <typename T, typename TI>
__global__ void im2col_vec8_kernel(const TI *__restrict__ img, T *__restrict__ out, int res, int P, long long total8,
                                   PixelNorm nm) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total8) return;
  const int G = res / P, runs = P / 8, cols8 = 3 * P * runs;  // 16-byte chunks per patch row
  const int j = (int)(idx % cols8);
  const long long row = idx / cols8;
  const int kx8 = j % runs, ky = (j / runs) % P, c = j / (runs * P);
  const int p = (int)(row % (G * G));
  const long long b = row / (G * G);
  const int py = p / G, px = p % G;
  const TI *src = img + ((b * 3 + c) * res + (py * P + ky)) * (long long)res + px * P + kx8 * 8;
  Vec16<T> o;
  if constexpr (sizeof(TI) == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(src)), bq = __ldg(reinterpret_cast<const float4 *>(src) + 1);
    o.v[0] = fromf<T>(a.x); o.v[1] = fromf<T>(a.y); o.v[2] = fromf<T>(a.z); o.v[3] = fromf<T>(a.w);
    o.v[4] = fromf<T>(bq.x); o.v[5] = fromf<T>(bq.y); o.v[6] = fromf<T>(bq.z); o.v[7] = fromf<T>(bq.w);
  } else if constexpr (sizeof(TI) == 1) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(src));  // 8 pixels
    const float mean = nm.mean[c], std = nm.std[c];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o.v[e] = fromf<T>(pixel_norm((uint8_t)(raw.x >> (8 * e)), mean, std));
      o.v[4 + e] = fromf<T>(pixel_norm((uint8_t)(raw.y >> (8 * e)), mean, std));
    }
  } else {
    o = ld16(reinterpret_cast<const T *>(src));
  }
  st16(out + idx * 8, o);
}

template <typename T>
int im2col_patches(const void *image, int image_dtype, T *out, int B, int res, int patch, int ld, const PixelNorm &nm,
                   cudaStream_t st) {
  int G = res / patch;
  long long total = (long long)B * G * G * ld;
  if (total == 0) return RPO_OK;
  RPO_REQUIRE(image_dtype == RPO_F32 || image_dtype == RPO_U8 || image_dtype == Num<T>::dtype,
              "image must be f32, uint8 or the model dtype");
  if (sizeof(T) == 2 && patch % 8 == 0 && res % 8 == 0 && ld == 3 * patch * patch && ((uintptr_t)image & 31) == 0 &&
      ((uintptr_t)out & 15) == 0) {
    const long long total8 = total / 8;
    const unsigned grid8 = (unsigned)((total8 + 255) / 256);
    if (image_dtype == RPO_F32)
      im2col_vec8_kernel<T, float><<<grid8, 256, 0, st>>>((const float *)image, out, res, patch, total8, nm);
    else if (image_dtype == RPO_U8)
      im2col_vec8_kernel<T, uint8_t><<<grid8, 256, 0, st>>>((const uint8_t *)image, out, res, patch, total8, nm);
    else
      im2col_vec8_kernel<T, T><<<grid8, 256, 0, st>>>((const T *)image, out, res, patch, total8, nm);
    RPO_LAUNCH_CHECK();
    return RPO_OK;
  }
  unsigned grid = (unsigned)((total + 255) / 256);
  if (image_dtype == RPO_F32)
    im2col_kernel<T, float><<<grid, 256, 0, st>>>((const float *)image, out, B, res, patch, ld, total, nm);
  else if (image_dtype == RPO_U8)
    im2col_kernel<T, uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)image, out, B, res, patch, ld, total, nm);
  else
    im2col_kernel<T, T><<<grid, 256, 0, st>>>((const T *)image, out, B, res, patch, ld, total, nm);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// trainers/rpo.py:199-204: [cls ; patches] + pos (dtype add), then the K image prompts appended
// (prompts get no positional embedding).  Context rows are image-major [B*S, D]; prompt rows
// [B*K, D].  ln_pre (:206) is applied afterwards by layernorm_fwd over both row sets.
template <typename T>
__global__ void vision_assemble_kernel(const T *__restrict__ patch_emb, const float *__restrict__ cls,
                                       const float *__restrict__ pos, const T *__restrict__ img_prompt,
                                       T *__restrict__ x_ctx, T *__restrict__ x_prompt, int B, int S, int K, int D) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n_ctx = (long long)B * S * D;
  long long total = n_ctx + (long long)B * K * D;
  if (idx >= total) return;
  if (idx < n_ctx) {
    int d = (int)(idx % D);
    long long r = idx / D;
    int s = (int)(r % S);
    int b = (int)(r / S);
    float e = (s == 0) ? rnd<T>(cls[d]) : tof<T>(patch_emb[((size_t)b * (S - 1) + (s - 1)) * D + d]);
    x_ctx[idx] = fromf<T>(e + rnd<T>(pos[(size_t)s * D + d]));
  } else {
    long long j = idx - n_ctx;
    int d = (int)(j % D);
    int i = (int)((j / D) % K);
    x_prompt[j] = img_prompt[(size_t)i * D + d];
  }
}

// 8 elements (16 bytes) per thread
template <typename T>
__global__ void vision_assemble_vec8_kernel(const T *__restrict__ patch_emb, const float *__restrict__ cls,
                                            const float *__restrict__ pos, const T *__restrict__ img_prompt,
                                            T *__restrict__ x_ctx, T *__restrict__ x_prompt, int B, int S, int K, int D) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int D8 = D / 8;
  const long long n_ctx = (long long)B * S * D8;
  const long long total = n_ctx + (long long)B * K * D8;
  if (idx >= total) return;
  if (idx < n_ctx) {
    const int d = (int)(idx % D8) * 8;
    const long long r = idx / D8;
    const int s = (int)(r % S);
    const long long b = r / S;
    float pv[8], ev[8];
    load_f32<8>(pos + (size_t)s * D + d, pv);
    if (s == 0) {
      load_f32<8>(cls + d, ev);
#pragma unroll
      for (int e = 0; e < 8; ++e) ev[e] = rnd<T>(ev[e]);
    } else {
      Vec16<T> pe = ld16(patch_emb + ((size_t)b * (S - 1) + (s - 1)) * D + d);
#pragma unroll
      for (int e = 0; e < 8; ++e) ev[e] = tof<T>(pe.v[e]);
    }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = fromf<T>(ev[e] + rnd<T>(pv[e]));
    st16(x_ctx + idx * 8, o);
  } else {
    const long long j = idx - n_ctx;
    const int d = (int)(j % D8) * 8;
    const int i = (int)((j / D8) % K);
    st16(x_prompt + j * 8, ld16(img_prompt + (size_t)i * D + d));
  }
}

template <typename T>
int vision_assemble_lnpre(const T *patch_emb, const float *cls, const float *pos, const float *, const float *,
                          const T *img_prompt, T *x_ctx, T *x_prompt, int B, int S, int K, int D, cudaStream_t st) {
  long long total = (long long)B * (S + K) * D;
  if (total == 0) return RPO_OK;
  if (sizeof(T) == 2 && D % 8 == 0 &&
      (((uintptr_t)patch_emb | (uintptr_t)cls | (uintptr_t)pos | (uintptr_t)img_prompt | (uintptr_t)x_ctx |
        (uintptr_t)x_prompt) & 15) == 0) {
    const long long total8 = total / 8;
    vision_assemble_vec8_kernel<T><<<(unsigned)((total8 + 255) / 256), 256, 0, st>>>(patch_emb, cls, pos, img_prompt,
                                                                                    x_ctx, x_prompt, B, S, K, D);
    RPO_LAUNCH_CHECK();
    return RPO_OK;
  }
  vision_assemble_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(patch_emb, cls, pos, img_prompt, x_ctx,
                                                                            x_prompt, B, S, K, D);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// context rows of the text tower: x_ctx[r, :] = text_x[row_cls[r], row_pos[r], :]
template <typename T>
__global__ void text_gather_kernel(const T *__restrict__ text_x, const int *__restrict__ row_cls,
                                   const int *__restrict__ row_pos, T *__restrict__ x_ctx, long long Mc, int T_len,
                                   int D) {
  constexpr int VEC = Vec16<T>::N;
  int nv = D / VEC;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Mc * nv) return;
  long long r = idx / nv;
  int v = (int)(idx % nv);
  const T *src = text_x + ((size_t)row_cls[r] * T_len + row_pos[r]) * D + (size_t)v * VEC;
  st16(x_ctx + r * D + (size_t)v * VEC, ld16(src));
}
template <typename T>
int text_gather_ctx(const T *text_x, const int *, const int *row_cls, const int *row_pos, T *x_ctx, long long Mc,
                    int T_len, int D, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  long long total = Mc * (D / VEC);
  if (total == 0) return RPO_OK;
  text_gather_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(text_x, row_cls, row_pos, x_ctx, Mc, T_len, D);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// dst[g*K + i, :] = src[i, :]   (the shared text prompt spliced into every class, trainers/rpo.py:176-177)
template <typename T>
__global__ void broadcast_rows_kernel(const T *__restrict__ src, T *__restrict__ dst, int G, int K, int D) {
  constexpr int VEC = Vec16<T>::N;
  int nv = D / VEC;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)G * K * nv;
  if (idx >= total) return;
  int v = (int)(idx % nv);
  int i = (int)((idx / nv) % K);
  st16(dst + idx * VEC, ld16(src + ((size_t)i * nv + v) * VEC));
}
template <typename T>
int broadcast_rows(const T *src, T *dst, int G, int K, int D, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::N;
  long long total = (long long)G * K * (D / VEC);
  if (total == 0) return RPO_OK;
  broadcast_rows_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dst, G, K, D);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// dst[i, d] = sum_g src[g*K + i, d]   (backward of the broadcast; f32 accumulation).
// grid.y splits the groups; partial sums are combined with atomics into a zeroed dst.
template <typename T>
__global__ void reduce_groups_kernel(const T *__restrict__ src, float *__restrict__ dst, int G, int K, int D,
                                     int g_per_block, float scale) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * D) return;
  int g0 = blockIdx.y * g_per_block;
  int g1 = min(G, g0 + g_per_block);
  float acc = 0.f;
  for (int g = g0; g < g1; ++g) acc += tof<T>(src[(size_t)g * K * D + idx]);
  atomicAdd(dst + idx, acc * scale);
}
template <typename T>
int reduce_groups_f32(const T *src, float *dst, int G, int K, int D, float scale, cudaStream_t st) {
  RPO_CHECK_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * K * D, st));
  if (G == 0) return RPO_OK;
  int g_per_block = 16;
  dim3 grid((K * D + 255) / 256, (G + g_per_block - 1) / g_per_block);
  reduce_groups_kernel<T><<<grid, 256, 0, st>>>(src, dst, G, K, D, g_per_block, scale);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// gradient of ln_pre w.r.t. the image prompts: every image sees the same prompt row, so
// d img_prompt[i] = dLN/dx(sum_b dy[b,i]; x = img_prompt[i]).  One warp per prompt row, D <= 1024.
template <typename T>
__global__ void lnpre_prompt_bwd_kernel(const float *__restrict__ dsum, const T *__restrict__ xp,
                                        const float *__restrict__ w, float *__restrict__ grad, int K, int D) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= K) return;
  const T *x = xp + (size_t)row * D;
  const float *dy = dsum + (size_t)row * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += tof<T>(x[c]);
  float mean = warp_sum(s) / D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) {
    float d = tof<T>(x[c]) - mean;
    q += d * d;
  }
  float rstd = 1.0f / sqrtf(warp_sum(q) / D + LN_EPS);
  float sg = 0.f, sgx = 0.f;
  for (int c = lane; c < D; c += 32) {
    float g = dy[c] * w[c];
    float xh = (tof<T>(x[c]) - mean) * rstd;
    sg += g;
    sgx += g * xh;
  }
  sg = warp_sum(sg) / D;
  sgx = warp_sum(sgx) / D;
  for (int c = lane; c < D; c += 32) {
    float g = dy[c] * w[c];
    float xh = (tof<T>(x[c]) - mean) * rstd;
    grad[(size_t)row * D + c] = rstd * (g - sg - xh * sgx);
  }
}
template <typename T>
int lnpre_prompt_bwd(const float *dsum, const T *img_prompt, const float *w, float *grad, int K, int D,
                     cudaStream_t st) {
  if (K == 0) return RPO_OK;
  lnpre_prompt_bwd_kernel<T><<<(K + 3) / 4, 128, 0, st>>>(dsum, img_prompt, w, grad, K, D);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// dst[c, r] = src[r, c]  (one-off at weight-bind time: K-major copies for the backward GEMMs)
template <typename T>
__global__ void transpose_kernel(const T *__restrict__ src, T *__restrict__ dst, int rows, int cols) {
  __shared__ T tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * cols + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) dst[(size_t)c2 * rows + r2] = tile[threadIdx.x][j];
  }
}
template <typename T>
int transpose_2d(const T *src, T *dst, int rows, int cols, cudaStream_t st) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<T><<<grid, block, 0, st>>>(src, dst, rows, cols);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

// torch.optim.SGD step (momentum, dampening 0, L2 weight decay), parameter stored in T
template <typename T>
__global__ void sgd_kernel(T *__restrict__ p, const float *__restrict__ g, float *__restrict__ buf, long long n,
                           const float *__restrict__ lr, float mom, float wd, float gscale,
                           const int *__restrict__ first) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pv = tof<T>(p[i]);
  float gv = gscale * g[i] + wd * pv;
  float b = gv;
  if (mom != 0.f) {
    b = (first && *first) ? gv : mom * buf[i] + gv;
    buf[i] = b;
  }
  p[i] = fromf<T>(pv - (*lr) * b);
}
template <typename T>
int sgd_step(T *p, const float *g, float *buf, long long n, const float *lr, float mom, float wd, float gscale,
             const int *first, cudaStream_t st) {
  if (n == 0) return RPO_OK;
  sgd_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, buf, n, lr, mom, wd, gscale, first);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

#define INSTANTIATE(T)                                                                                              \
  template int layernorm_fwd<T>(const T *, const float *, const float *, T *, long long, int, cudaStream_t);        \
  template int layernorm_bwd<T>(const T *, const T *, const float *, const T *, T *, long long, int, cudaStream_t); \
  template int im2col_patches<T>(const void *, int, T *, int, int, int, int, const PixelNorm &, cudaStream_t);                              \
  template int vision_assemble_lnpre<T>(const T *, const float *, const float *, const float *, const float *,      \
                                        const T *, T *, T *, int, int, int, int, cudaStream_t);                     \
  template int text_gather_ctx<T>(const T *, const int *, const int *, const int *, T *, long long, int, int,       \
                                  cudaStream_t);                                                                    \
  template int broadcast_rows<T>(const T *, T *, int, int, int, cudaStream_t);                                      \
  template int reduce_groups_f32<T>(const T *, float *, int, int, int, float, cudaStream_t);                            \
  template int lnpre_prompt_bwd<T>(const float *, const T *, const float *, float *, int, int, cudaStream_t);       \
  template int transpose_2d<T>(const T *, T *, int, int, cudaStream_t);                                             \
  template int sgd_step<T>(T *, const float *, float *, long long, const float *, float, float, float, const int *, \
                           cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(__half)
INSTANTIATE(__nv_bfloat16)

}  // namespace rpo
