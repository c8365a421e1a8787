// Shared device/host helpers for the rpo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "../../include/rpo_b200.h"

namespace rpo {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const std::string &msg);
extern thread_local int64_t g_launch_count;

#define RPO_CHECK_CUDA(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      rpo::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                     std::to_string(__LINE__));                                                  \
      return RPO_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)

#define RPO_REQUIRE(cond, msg)                                                                   \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      rpo::set_error(std::string("invalid argument: ") + msg + " (" #cond ") at " + __FILE__ + ":" + \
                     std::to_string(__LINE__));                                                  \
      return RPO_ERR_INVALID;                                                                    \
    }                                                                                            \
  } while (0)

#define RPO_TRY(expr)              \
  do {                             \
    int _s = (expr);               \
    if (_s != RPO_OK) return _s;   \
  } while (0)

// every kernel launch goes through this so that launches are counted and checked; `st` (the stream
// of the launch) must be in scope.  With the launch profiler armed (rpo_profile_begin) a CUDA event
// is recorded after the launch so that rpo_profile_end can report per-launch device times measured
// at the clocks of a real, back-to-back step (ncu's serialised launches run at idle clocks).
void prof_mark(const char *file, int line, cudaStream_t st);
void prof_tag(const char *fmt, ...);
extern thread_local bool g_prof_on;
#define RPO_LAUNCH_CHECK()                                      \
  do {                                                          \
    rpo::g_launch_count++;                                      \
    RPO_CHECK_CUDA(cudaPeekAtLastError());                      \
    if (rpo::g_prof_on) rpo::prof_mark(__FILE__, __LINE__, st); \
  } while (0)

// Tuning / fault-injection switches exist only in the diagnostics build (RPO_DIAG=1 python -m rpo_b200.build): the
// release library answers "unset" for them, and the strings are not even in it.
#ifdef RPO_DIAG
inline const char *diag_env(const char *name) { return getenv(name); }
#else
inline const char *diag_env(const char *) { return nullptr; }
#endif

inline size_t dtype_size(int dtype) { return dtype == RPO_F32 ? 4 : 2; }

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// The towers are chains of ~170 short dependent kernels per step.  Kernels launched through
// launch_pdl() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may be scheduled
// while the previous kernel on the stream is still draining, run their prologue (barrier init, TMEM
// allocation, tensor-map and weight-tile prefetch), and block in pdl_wait() -- griddepcontrol.wait
// returns only when the upstream grid has COMPLETED and its writes are visible -- before touching
// anything an upstream kernel produces or still reads.  Every kernel launched this way must call
// pdl_wait() in each thread that reads or writes activations.  RPO_NO_PDL=1 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

// "Settled operands": the prompt-row backward re-reads tensors the FORWARD pass left behind (context keys / values,
// prompt queries, the residual stream a LayerNorm normalised).  Inside a step those were written hundreds of launches
// earlier; every backward stage opens with a head GEMM and a LayerNorm backward that triggers its dependents only
// after its own wait (so nothing behind it is scheduled before the head GEMM -- and with it everything older on the
// stream -- has completed), and the logit stage in between consists of plain launches.  The tensors are therefore
// complete before any kernel of tower_backward can even be scheduled, and such a kernel may fetch them AHEAD of
// its pdl_wait(), while the upstream kernel is still running (what the GEMMs do with the frozen weights).  Only the
// engine's backward chain makes that promise (SettledOperands scope around its launches); the unit entry points of the
// C ABI cannot know who produced their arguments and keep the wait first.
extern thread_local bool g_operands_settled;
struct SettledOperands {
  bool prev;
  SettledOperands() : prev(g_operands_settled) { g_operands_settled = pdl_enabled(); }
  ~SettledOperands() { g_operands_settled = prev; }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- dtype traits -------------------------------------------------------------------------------
template <typename T>
struct Num;
template <>
struct Num<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static constexpr int dtype = RPO_F32;
};
template <>
struct Num<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static constexpr int dtype = RPO_F16;
};
template <>
struct Num<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static constexpr int dtype = RPO_BF16;
};

template <typename T>
__device__ __forceinline__ float tof(T v) {
  return Num<T>::to_f(v);
}
template <typename T>
__device__ __forceinline__ T fromf(float v) {
  return Num<T>::from_f(v);
}
// round an f32 value through T (what storing it in a T tensor would do)
template <typename T>
__device__ __forceinline__ float rnd(float v) {
  return tof<T>(fromf<T>(v));
}

// 16-byte vector of T
template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};
template <typename T>
__device__ __forceinline__ Vec16<T> ld16(const T *p) {
  Vec16<T> r;
  *reinterpret_cast<uint4 *>(&r) = *reinterpret_cast<const uint4 *>(p);
  return r;
}
template <typename T>
__device__ __forceinline__ void st16(T *p, const Vec16<T> &r) {
  *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(&r);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// QuickGELU (clip/model.py:162-164) evaluated the way the reference's dtype-typed tensor ops do:
// t = T(1.702*x); s = T(sigmoid(t)); y = T(x*s).  For T = float the roundings are identities.
template <typename T>
__device__ __forceinline__ float quickgelu_rounded(float x) {
  float t = rnd<T>(1.702f * x);
  float s = rnd<T>(1.0f / (1.0f + __expf(-t)));
  return x * s;
}
template <>
__device__ __forceinline__ float quickgelu_rounded<float>(float x) {
  return x * (1.0f / (1.0f + expf(-1.702f * x)));
}
// d/dz [ z * sigmoid(1.702 z) ]
__device__ __forceinline__ float quickgelu_grad(float z) {
  float s = 1.0f / (1.0f + expf(-1.702f * z));
  return s * (1.0f + 1.702f * z * (1.0f - s));
}

// ---- fused GEMM epilogue description (shared by the SIMT and tcgen05 kernels) --------------------
template <typename T>
struct Epilogue {
  const T *bias;           // [N] or null
  const T *residual;       // [M, ldc] or null
  const T *gelu_grad_aux;  // [M, ldc] or null: multiply by quickgelu'(aux)
  T *aux_out;              // [M - aux_row0, ldc] or null: pre-activation copy for rows >= aux_row0
  long long aux_row0;
  int act;
  // operand B is a frozen weight (never written on the step's streams): its first tiles may be
  // fetched before the upstream kernel has finished (programmatic dependent launch)
  int b_frozen;
  // Row split of the output (tcgen05 kernels only): rows >= split_row go to c2[(m - split_row) * ldc2 + n] for
  // columns n < ncols2 and are dropped for the other columns.  The vision tower's QKV projection uses it to run
  // over context AND prompt rows in one launch: context rows -> [Mc, 3D] q|k|v, prompt rows -> their q only
  // (prompts are never keys or values).  Not combinable with residual / gelu' rows / aux_out.  c2 == null: off.
  T *c2;
  long long ldc2, split_row;
  int ncols2;

  // v: f32 accumulator for element (m, n); returns the value to store (already rounded through T
  // at the points where the reference materialises a dtype tensor).
  __device__ __forceinline__ float apply(float v, long long m, int n, long long ldc) const {
    if (bias) v += tof<T>(bias[n]);
    v = rnd<T>(v);  // nn.Linear output tensor
    if (aux_out && m >= aux_row0) aux_out[(m - aux_row0) * ldc + n] = fromf<T>(v);
    if (act == RPO_ACT_QUICKGELU) v = rnd<T>(quickgelu_rounded<T>(v));
    if (gelu_grad_aux) v = rnd<T>(v * quickgelu_grad(tof<T>(gelu_grad_aux[m * ldc + n])));
    if (residual) v = v + tof<T>(residual[m * ldc + n]);
    return v;
  }
};

// ---- kernel launchers implemented across the .cu files -------------------------------------------
// tcgen05 form of the three K-pair contractions of the logit block (logits_tc.cu)
bool logits_tc_supported(int dtype, int B, int C, int K, int E);
template <typename T>
int logits_pair_fwd_tc(const T *img_s, const T *text_n, T *pair, int B, int C, int K, int E, cudaStream_t st);
template <typename T>
int logits_pair_bwd_tc(const T *dl, const T *img_s, const T *text_n, T *d_img_s, T *d_text_n, int B, int C, int K,
                       int E, cudaStream_t st);
template <typename T>
int gemm_simt(const T *A, long long sam, long long sak, const T *B, long long sbn, long long sbk, T *C, long long ldc,
              long long M, int N, int Kd, const Epilogue<T> &ep, int batch, long long bsa, long long bsb,
              long long bsc, cudaStream_t st);

// tcgen05 path; returns RPO_ERR_INVALID (without launching) if the shape is not supported
template <typename T>
int gemm_tcgen05(const T *A, long long lda, const T *B, long long ldb, T *C, long long ldc, long long M, int N, int Kd,
                 const Epilogue<T> &ep, cudaStream_t st);
bool gemm_tcgen05_supported(int dtype, long long lda, long long ldb, long long ldc, long long M, int N, int Kd,
                            const void *A, const void *B, const void *C);

template <typename T>
int gemm_dispatch(int backend, const T *A, long long lda, const T *B, long long ldb, T *C, long long ldc, long long M,
                  int N, int Kd, const Epilogue<T> &ep, cudaStream_t st);

template <typename T>
int layernorm_fwd(const T *x, const float *w, const float *b, T *y, long long rows, int D, cudaStream_t st);
template <typename T>
int layernorm_bwd(const T *dy, const T *x, const float *w, const T *dres, T *dx, long long rows, int D,
                  cudaStream_t st);

template <typename T>
int ro_attention_fwd(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, const int *ctx_off, int G, int K,
                     int H, int max_ctx, int causal, int do_ctx, cudaStream_t st);
// tcgen05 forward for uniform groups without causal mask (the vision tower): attention_tc.cu
bool ro_attention_fwd_dense_supported(int dtype, int n, int K, int H);
template <typename T>
int ro_attention_fwd_dense(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, int G, int n, int K, int H,
                           cudaStream_t st);
template <typename T>
int ro_attention_bwd(const T *qkv_ctx, const T *q_prompt, const T *o_prompt, const T *d_out, T *dq, const int *ctx_off,
                     int G, int K, int H, int max_ctx, cudaStream_t st);

template <typename T>
int logits_ce_fwd(const T *img_feat, const T *text_feat, const float *logit_scale, const int64_t *label, int B, int C,
                  int K, int E, T *img_n, T *img_s, T *text_n, float *img_rnorm, float *text_rnorm, T *pair_logits,
                  float *logits, float *loss, float *dlogits, cudaStream_t st);
template <typename T>
int logits_ce_bwd(const float *dlogits, const T *img_feat, const T *text_feat, const T *img_n, const T *img_s,
                  const T *text_n, const float *img_rnorm, const float *text_rnorm, const float *logit_scale, int B,
                  int C, int K, int E, T *dl_t, T *d_img_s, T *d_text_n, T *d_img_feat, T *d_text_feat,
                  float grad_scale, cudaStream_t st);

// embedding / glue kernels (elementwise.cu)
struct PixelNorm {  // per-channel Normalize of uint8 images (clip/clip.py:77)
  float mean[3], std[3];
};
template <typename T>
int im2col_patches(const void *image, int image_dtype, T *out, int B, int res, int patch, int ld, const PixelNorm &nm,
                   cudaStream_t st);
template <typename T>
int vision_assemble_lnpre(const T *patch_emb, const float *cls, const float *pos, const float *w, const float *b,
                          const T *img_prompt, T *x_ctx, T *x_prompt, int B, int S, int K, int D, cudaStream_t st);
template <typename T>
int text_gather_ctx(const T *text_x, const int *ctx_off, const int *row_cls, const int *row_pos, T *x_ctx,
                    long long Mc, int T_len, int D, cudaStream_t st);
template <typename T>
int broadcast_rows(const T *src, T *dst, int G, int K, int D, cudaStream_t st);
template <typename T>
int reduce_groups_f32(const T *src, float *dst, int G, int K, int D, float scale, cudaStream_t st);
template <typename T>
int lnpre_prompt_bwd(const float *dsum, const T *img_prompt, const float *w, float *grad, int K, int D,
                     cudaStream_t st);
template <typename T>
int transpose_2d(const T *src, T *dst, int rows, int cols, cudaStream_t st);
template <typename T>
int sgd_step(T *p, const float *g, float *buf, long long n, const float *lr, float mom, float wd, float gscale,
             const int *first, cudaStream_t st);

}  // namespace rpo
