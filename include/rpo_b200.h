/*
 * rpo_b200 -- C ABI of the B200-native RPO hot path (CLIP ViT towers with K read-only prompts).
 *
 * The reference (mlvlab/RPO) has no FFI: its boundary is the Python class surface of
 * trainers/rpo.py (CustomCLIP / PromptLearner, lines 41-232).  This header is what the host-side
 * mirror of that surface (rpo_b200/model.py) binds through ctypes; each entry point names the
 * reference code it replaces.  Plain C: pointers, sizes, an explicit cudaStream_t (passed as
 * void*), int status returns (0 = ok, negative = error; rpo_last_error() gives the message).
 * All device pointers are caller-owned unless stated; no entry point below allocates or
 * synchronises on the hot path (rpo_forward / rpo_backward / rpo_sgd_step / the unit kernels),
 * so the whole step can be captured into a CUDA graph by the caller.
 *
 * One handle per process/GPU, used from one host thread (same contract as the reference's single
 * Python training thread, trainers/rpo.py:290-316).
 */
#ifndef RPO_B200_H
#define RPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPO_OK 0
#define RPO_ERR_INVALID (-1)
#define RPO_ERR_CUDA (-2)
#define RPO_ERR_STATE (-3)

/* element type of activations and of the Linear/Conv/projection weights
 * (clip/model.py:379-400 convert_weights; LayerNorm params, embeddings, logit_scale are always f32) */
enum { RPO_F32 = 0, RPO_F16 = 1, RPO_BF16 = 2,
       RPO_U8 = 3 /* image_dtype of rpo_forward only: raw pixels, normalised on the GPU */ };

/* GEMM backends: tcgen05 (TMA + UMMA + TMEM, 16-bit types) or the generic SIMT kernel (any type,
 * exact fp32 FMA; the only backend for RPO_F32).  AUTO picks tcgen05 whenever the shape allows. */
enum { RPO_GEMM_AUTO = 0, RPO_GEMM_SIMT = 1, RPO_GEMM_TCGEN05 = 2 };

/* epilogue activation */
enum { RPO_ACT_NONE = 0, RPO_ACT_QUICKGELU = 1 };

typedef struct RpoHandle RpoHandle;

/* Model geometry.  Replaces the constants the reference hard-codes or reads off the CLIP module:
 * trainers/rpo.py:52 (d_v), :142 (attn_head), :154 (1+14*14), :185 (512), clip/model.py:403-440. */
typedef struct {
  int32_t dtype;        /* RPO_F32 / RPO_F16 / RPO_BF16 */
  int32_t K;            /* number of prompt pairs, cfg.TRAINER.RPO.K (trainers/rpo.py:49) */
  int32_t n_cls;        /* C, number of class prompts */
  int32_t ctx_len;      /* T = 77 */
  int32_t embed_dim;    /* E */
  int32_t v_width, v_layers, v_heads, v_patch, v_res;
  int32_t t_width, t_layers, t_heads;
  int32_t max_batch;    /* largest B rpo_forward will be called with (workspace is sized for it) */
  int32_t gemm_backend; /* RPO_GEMM_* */
  /* Class shard of the text tower (data-parallel ranks each run C/G of the class prompts that
   * trainers/rpo.py:180-192 runs in full on every GPU).  cls_local == 0: the handle owns all n_cls classes.
   * Otherwise this handle's text tower covers classes [cls_first, cls_first + cls_local) only: rpo_set_classes
   * takes the text_x / len_prompts of those classes, and the step is driven through the stage entry points below
   * with an all-gather of the text features and a reduce-scatter of their gradient in between.  Logits, loss and
   * the image side always span all n_cls classes. */
  int32_t cls_first, cls_local;
  /* 2: the handle keeps two sets of vision-tower activations so that the context rows of the NEXT batch
   * (rpo_forward_image_context, slot s) can be computed while the prompt rows and the backward of the current batch
   * (rpo_forward_image_prompts, other slot) are still in flight.  0 / 1: one set. */
  int32_t image_slots;
} RpoConfig;

/* One ResidualAttentionBlock (clip/model.py:167-191).  ln_* are f32 [D]; the rest are `dtype`:
 * in_w [3D,D], in_b [3D], out_w [D,D], out_b [D], fc_w [4D,D], fc_b [4D], proj_w [D,4D], proj_b [D]. */
typedef struct {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const void *in_w, *in_b, *out_w, *out_b, *fc_w, *fc_b, *proj_w, *proj_b;
} RpoBlockWeights;

/* Frozen CLIP weights on the path (trainers/rpo.py:104-120).  Device pointers, not owned; they
 * must outlive the handle.  `*_blocks` are HOST arrays of v_layers / t_layers entries. */
typedef struct {
  const RpoBlockWeights *v_blocks;
  const RpoBlockWeights *t_blocks;
  const void *conv_w;       /* visual.conv1.weight [Dv, 3*p*p] dtype (clip/model.py:215) */
  const float *cls_emb;     /* visual.class_embedding [Dv] f32 (:218) */
  const float *v_pos;       /* visual.positional_embedding [S, Dv] f32 (:219) */
  const float *ln_pre_w, *ln_pre_b, *ln_post_w, *ln_post_b; /* f32 [Dv] (:220,:224) */
  const void *v_proj;       /* visual.proj [Dv, E] dtype (:225) */
  const float *ln_final_w, *ln_final_b;                     /* f32 [Dt] */
  const void *t_proj;       /* text_projection [Dt, E] dtype */
  const float *logit_scale; /* f32 device scalar */
} RpoWeights;

const char *rpo_last_error(void);
int rpo_version(void);

/* lifetime ------------------------------------------------------------------------------------ */
int rpo_create(const RpoConfig *cfg, RpoHandle **out);
void rpo_destroy(RpoHandle *h);
/* bytes of device memory the handle owns (workspace + transposed weight copies + context cache) */
size_t rpo_device_bytes(const RpoHandle *h);

/* Replaces CustomCLIP.__init__'s aliasing of the frozen sub-modules (trainers/rpo.py:104-120).
 * Builds the K-major transposed copies the backward GEMMs read (weights are frozen, so this is a
 * one-off).  Synchronises `stream` before returning. */
int rpo_bind_weights(RpoHandle *h, const RpoWeights *w, void *stream);

/* Replaces make_prompts' text_x/len_prompts (trainers/rpo.py:135-137), define_mask (:140-159) and
 * the prompt-independent part of the text tower call (:180-181): runs the n_c = len_prompts[c]
 * readable context tokens of every class through the text transformer once and caches their
 * per-layer K/V.  text_x: device [C, T, Dt] dtype (token + positional embedding);
 * len_prompts: HOST int32 [C], each in [1, T-K].  Synchronises `stream`.
 * With a class shard (RpoConfig.cls_local > 0) both arrays hold the cls_local classes of the handle only. */
int rpo_set_classes(RpoHandle *h, const void *text_x, const int32_t *len_prompts, void *stream);

/* CustomCLIP.forward (trainers/rpo.py:161-232).  image: device [B,3,res,res], f32 (image_dtype =
 * RPO_F32, cast to dtype inside like `image.type(self.dtype)`, :198), already `dtype`, or uint8
 * (RPO_U8: raw pixels; ToTensor + Normalize of clip/clip.py:75-78 are applied in f32 inside the patch
 * extraction, constants from rpo_set_image_norm -- a quarter of the host-to-device bytes).
 * text_prompt [K,Dt], img_prompt [K,Dv]: device, dtype (PromptLearner.forward, :89-90).
 * label: device int64 [B] or NULL.  logits: device f32 [B,C] or NULL.  loss: device f32 scalar or
 * NULL (requires label).  Saves what rpo_backward needs inside the handle.
 * text_prompt == NULL (inference only: label must be NULL) reuses the text features of the last call
 * that was given a text prompt: at test time the reference runs the whole text tower again for every
 * batch (trainers/rpo.py:173-192 under TrainerX.test) although the prompts no longer change. */
int rpo_forward(RpoHandle *h, const void *image, int32_t image_dtype, int32_t B, const void *text_prompt,
                const void *img_prompt, const int64_t *label, float *logits, float *loss, void *stream);

/* Per-channel mean / std of the uint8 image path (defaults: the CLIP constants of clip/clip.py:77). */
int rpo_set_image_norm(RpoHandle *h, const float mean[3], const float std[3]);

/* loss.backward() of trainers/rpo.py:308 restricted to what has a gradient (:258-260): writes
 * d loss / d text_prompt into grad_flat[0 : K*Dt] and d loss / d img_prompt into
 * grad_flat[K*Dt : K*Dt + K*Dv] (f32, contiguous: all-reduce ready).  Must follow an rpo_forward
 * that was given a label. */
int rpo_backward(RpoHandle *h, float *grad_flat, void *stream);

/* Stage entry points ------------------------------------------------------------------------------
 * rpo_forward == rpo_forward_text (on a side stream) || rpo_forward_image, then rpo_forward_logits;
 * rpo_backward == rpo_backward_logits, then rpo_backward_text (side stream) || rpo_backward_image.
 * They exist so that a class-sharded text tower (RpoConfig.cls_local) can put its two collectives between the
 * stages: text features are all-gathered after rpo_forward_text, their gradient is reduce-scattered (sum) after
 * rpo_backward_logits.  Each call only enqueues kernels on `stream`; ordering between streams is the caller's
 * (events).  The exchange buffers are caller-owned so that the collective library can address them:
 * text_feat and d_text_feat, both [C_pad, K, E] `dtype` with C_pad >= n_cls (rows >= n_cls are never touched by
 * the library: keep them zero).  rpo_forward_text writes rows [cls_first*K, (cls_first+cls_local)*K) of
 * text_feat; rpo_forward_logits reads rows [0, n_cls*K); rpo_backward_logits writes rows [0, n_cls*K) of
 * d_text_feat (the contribution of THIS rank's images); rpo_backward_text reads rows
 * [cls_first*K, (cls_first+cls_local)*K) of d_text_feat and writes grad_flat[0 : K*Dt] (sum over the local
 * classes); rpo_backward_image writes grad_flat[K*Dt : K*Dt + K*Dv]. */
int rpo_bind_text_exchange(RpoHandle *h, void *text_feat, void *d_text_feat);
/* trainers/rpo.py:173-192 for the handle's classes: splice, text tower (prompt rows), ln_final, gather, projection */
int rpo_forward_text(RpoHandle *h, const void *text_prompt, void *stream);
/* trainers/rpo.py:198-211: patch embedding, prompt concat, ln_pre, vision tower, ln_post, projection */
int rpo_forward_image(RpoHandle *h, const void *image, int32_t image_dtype, int32_t B, const void *img_prompt,
                      void *stream);
/* rpo_forward_image in two passes.  The context rows (cls + patch tokens) of the vision tower depend on the image
 * and the frozen weights only -- visual_mask (trainers/rpo.py:155-156) hides the prompt columns from every row --
 * so they can be computed before this step's prompts exist, e.g. while the previous batch's prompt rows, backward
 * and SGD update are still running on another stream.  _context: patch embedding, ln_pre and all blocks over the
 * B*S context rows of `slot` (their per-layer q|k|v stay in the slot).  _prompts: the B*K prompt rows of the slot
 * (queries only), ln_post, projection; makes `slot` the one rpo_forward_logits / rpo_backward_image refer to.
 * slot < RpoConfig.image_slots.  Results equal rpo_forward_image's up to the rounding of the attention kernels
 * (prompt rows go through the mma.sync kernel here, the tcgen05 one there). */
int rpo_forward_image_context(RpoHandle *h, const void *image, int32_t image_dtype, int32_t B, int32_t slot,
                              void *stream);
int rpo_forward_image_prompts(RpoHandle *h, const void *img_prompt, int32_t slot, void *stream);
/* trainers/rpo.py:215-230: normalise, K-pair logits, cross-entropy (same argument rules as rpo_forward) */
int rpo_forward_logits(RpoHandle *h, const int64_t *label, float *logits, float *loss, void *stream);
/* d loss / d img_feat and d loss / d text_feat (all n_cls classes, this rank's images) */
int rpo_backward_logits(RpoHandle *h, void *stream);
int rpo_backward_text(RpoHandle *h, float *grad_flat, void *stream);
int rpo_backward_image(RpoHandle *h, float *grad_flat, void *stream);

/* optim.step() of trainers/rpo.py:309 for torch.optim.SGD semantics (momentum, dampening 0, L2
 * weight decay, no nesterov) applied to one parameter in `dtype` with an f32 gradient:
 *   g = scale*grad + wd*p ; buf = first ? g : mom*buf + g ; p -= lr*buf.
 * `first_step` is a device int32 flag (non-zero before any momentum exists) so the call is
 * graph-capturable; lr is a device f32 scalar for the same reason. */
int rpo_sgd_step(void *param, int32_t dtype, const float *grad, float *momentum_buf, int64_t n, const float *lr,
                 float momentum, float weight_decay, float grad_scale, const int32_t *first_step, void *stream);

/* ---- exchanges between the data-parallel ranks of one node through peer-mapped memory (NVLink / NVSwitch) ------
 * "one ncclAllReduce over the flat [K*Dt + K*Dv] buffer, then the identical SGD step on every
 * rank" (the reference's own multi-GPU is nn.DataParallel, trainers/rpo.py:282-285, which cannot train RPO).
 * Here the exchange is done by this library's own kernels over buffers every rank has mapped from every peer
 * (CUDA IPC / symmetric memory: the CALLER allocates and maps them, e.g. torch.distributed._symmetric_memory), so
 * the calls are plain kernel launches -- capturable into the step's CUDA graph, no host-issued collective between
 * graph segments.  Every rank must make the same calls in the same order.
 *
 * RpoPeerComm: signals[r] = rank r's signal words (rpo_peer_signal_bytes() bytes, zeroed once before the first
 * call) as mapped into THIS process; epoch = rpo_peer_epoch_bytes() bytes of local device memory, zeroed once. */
#define RPO_PEER_MAX_WORLD 8
#define RPO_PEER_MAX_BLOCKS 128
typedef struct {
  void *signals[RPO_PEER_MAX_WORLD];
  void *epoch;
  int32_t rank, world;
} RpoPeerComm;
size_t rpo_peer_signal_bytes(void);
size_t rpo_peer_epoch_bytes(void);
/* optim.step() (trainers/rpo.py:309) fused with the gradient all-reduce: grad_flat[r] = rank r's flat f32 gradient
 * [n_total] (text part first, n_text elements) as mapped here; the kernel sums them in rank order (bit-identical on
 * every replica), then updates both prompt tensors exactly as rpo_sgd_step does with grad = the sum. */
int rpo_peer_allreduce_sgd(const RpoPeerComm *comm, void *const *grad_flat, void *text_prompt, void *img_prompt,
                           int32_t dtype, int64_t n_text, int64_t n_total, float *momentum_buf, const float *lr,
                           float momentum, float weight_decay, float grad_scale, const int32_t *first_step,
                           void *stream);
/* class-sharded text tower: bufs[r] = rank r's exchange buffer (rpo_bind_text_exchange's text_feat /
 * d_text_feat).  all_gather pushes bytes [offset, offset + nbytes) of this rank's buffer to the same offset of every
 * peer's; reduce_scatter replaces elements [offset, offset + n) of this rank's buffer by the sum over all ranks'
 * buffers (f32 accumulation in rank order, one rounding).  *_max = the largest part of any rank (sizes the grid:
 * every rank must launch the same number of blocks). */
int rpo_peer_all_gather(const RpoPeerComm *comm, void *const *bufs, int64_t offset_bytes, int64_t nbytes,
                        int64_t nbytes_max, void *stream);
int rpo_peer_reduce_scatter(const RpoPeerComm *comm, void *const *bufs, int32_t dtype, int64_t offset_elems,
                            int64_t n_elems, int64_t n_elems_max, void *stream);

/* unit kernels (one per hot-path op; used by the parity tests) -------------------------------- */

/* clip/model.py:153-159 LayerNorm.forward: y = LN_f32(x; w, b, eps=1e-5) cast to dtype. x,y [rows,D].
 * w / b are parameters, not activations: the kernels copy them before they wait for the work enqueued ahead of them
 * on `stream`, so their contents must be final when the call is made (in the path they are frozen CLIP weights). */
int rpo_layernorm_fwd(const void *x, const float *w, const float *b, void *y, int64_t rows, int32_t D, int32_t dtype,
                      void *stream);
/* its input-gradient: dx = dres (nullable) + dLN/dx(dy; x, w).  All [rows,D] dtype. */
int rpo_layernorm_bwd(const void *dy, const void *x, const float *w, const void *dres, void *dx, int64_t rows,
                      int32_t D, int32_t dtype, void *stream);

/* nn.Linear / `@` (clip/model.py:174-176,186; trainers/rpo.py:191,210): C[M,N] = A[M,Kd] . B[N,Kd]^T
 * with fused epilogue: v = acc + bias[n]; aux_out (rows >= aux_row0) = v; v = act(v);
 * v *= quickgelu'(gelu_grad_aux[m,n]); C = v + residual[m,n].  Row strides lda/ldb/ldc in elements;
 * residual / aux use ldc.  Nullable: bias, residual, gelu_grad_aux, aux_out. */
int rpo_gemm_bias_act(const void *A, int64_t lda, const void *B, int64_t ldb, void *C, int64_t ldc, int64_t M,
                      int32_t N, int32_t Kd, const void *bias, int32_t act, const void *residual,
                      const void *gelu_grad_aux, void *aux_out, int64_t aux_row0, int32_t dtype, int32_t backend,
                      void *stream);

/* Read-only masked multi-head attention (clip/model.py:186 with the masks of trainers/rpo.py:140-159).
 * G groups (images / classes).  Group g has n_g = ctx_off[g+1]-ctx_off[g] context rows (its keys
 * and values, and -- if do_ctx -- also queries) stored at rows ctx_off[g].. of qkv_ctx [Mc, 3D]
 * (q | k | v, head-major inside each D), and K prompt rows (queries only) at rows g*K.. of
 * q_prompt [G*K, D].  Context row r reads keys j<n_g (vision) or j<=r (causal, text); prompt rows
 * read every key j<n_g.  Nothing reads a prompt row.  Output rows: out_ctx [Mc, D] (if do_ctx)
 * and out_prompt [G*K, D].  head_dim is 64.  ctx_off: device int32 [G+1]. */
int rpo_ro_attention_fwd(const void *qkv_ctx, const void *q_prompt, void *out_ctx, void *out_prompt,
                         const int32_t *ctx_off, int32_t G, int32_t K, int32_t H, int32_t max_ctx, int32_t causal,
                         int32_t do_ctx, int32_t dtype, void *stream);
/* Same contract for the vision tower's shape -- every group has exactly n_ctx context rows, no causal
 * mask (visual_mask, trainers/rpo.py:153-159), context and prompt rows all queries -- on the tcgen05
 * path: TMA-staged Q/K/V, S = QK^T and O = PV as tcgen05.mma with TMEM accumulators, one softmax
 * thread per query row.  16-bit dtypes; n_ctx <= 256 and (n_ctx % 128) + K <= 128 (returns
 * RPO_ERR_INVALID otherwise: use rpo_ro_attention_fwd).  qkv_ctx [G*n_ctx, 3D], q_prompt [G*K, D]. */
int rpo_ro_attention_fwd_dense(const void *qkv_ctx, const void *q_prompt, void *out_ctx, void *out_prompt, int32_t G,
                               int32_t n_ctx, int32_t K, int32_t H, int32_t dtype, void *stream);
/* 1 if rpo_ro_attention_fwd_dense takes this shape (dtype RPO_F16 / RPO_BF16, n_ctx context keys, K prompt queries,
 * H heads of 64), else 0 -- callers then use rpo_ro_attention_fwd */
int rpo_ro_attention_fwd_dense_supported(int32_t dtype, int32_t n_ctx, int32_t K, int32_t H);
/* gradient w.r.t. the prompt queries only (keys/values come from rows that carry no gradient):
 * dq_prompt [G*K, D] from d_out_prompt [G*K, D]; out_prompt is the forward output of the same rows
 * (used for the softmax-gradient row term sum_d dO*O). */
int rpo_ro_attention_bwd(const void *qkv_ctx, const void *q_prompt, const void *out_prompt, const void *d_out_prompt,
                         void *dq_prompt,
                         const int32_t *ctx_off, int32_t G, int32_t K, int32_t H, int32_t max_ctx, int32_t dtype,
                         void *stream);

/* trainers/rpo.py:215-230: L2-normalise img_feat [B,K,E] and text_feat [C,K,E] (dtype), K-pair
 * logits with exp(logit_scale) folded into the image side in `dtype`, f32 accumulation over the K
 * pairs, /K, and (label != NULL) mean cross-entropy.  Scratch is caller-provided:
 * img_n, img_s [B,K,E], text_n [C,K,E], pair_logits [K,B,C] (dtype), inv norms f32 [B*K], [C*K].
 * Outputs logits f32 [B,C], loss f32 scalar, dlogits f32 [B,C] (d loss / d logits; nullable). */
int rpo_logits_ce_fwd(const void *img_feat, const void *text_feat, const float *logit_scale, const int64_t *label,
                      int32_t B, int32_t C, int32_t K, int32_t E, void *img_n, void *img_s, void *text_n,
                      float *img_rnorm, float *text_rnorm, void *pair_logits, float *logits, float *loss,
                      float *dlogits, int32_t dtype, void *stream);
/* backward of the above to d img_feat [B,K,E] and d text_feat [C,K,E] (dtype).
 * Scratch: dl_t [B,C] dtype, d_img_s [B,K,E], d_text_n [C,K,E] dtype. */
int rpo_logits_ce_bwd(const float *dlogits, const void *img_feat, const void *text_feat, const void *img_n,
                      const void *img_s, const void *text_n, const float *img_rnorm, const float *text_rnorm,
                      const float *logit_scale, int32_t B, int32_t C, int32_t K, int32_t E, void *dl_t, void *d_img_s,
                      void *d_text_n, void *d_img_feat, void *d_text_feat, int32_t dtype, void *stream);

/* introspection for tests: copies of internal activations after rpo_forward.
 * which: 0 = vision residual stream after block `layer` ([B*(S+K), Dv]: context rows image-major,
 * then prompt rows image-major), 1 = text residual stream after block `layer` ([Mc + C*K, Dt]),
 * 2 = img_feat [B,K,E], 3 = text_feat [C,K,E].  Returns the number of elements, copies at most
 * `cap` elements (dtype) to `dst` (device).  layer = -1 means the tower input. */
int64_t rpo_debug_fetch(RpoHandle *h, int32_t which, int32_t layer, void *dst, int64_t cap, void *stream);

/* number of kernel launches issued by the last rpo_forward + rpo_backward pair (or the last call of each stage) */
int64_t rpo_launch_count(const RpoHandle *h);

/* Launch profiler (diagnostics; not for use inside CUDA-graph capture).  Between rpo_profile_begin and
 * rpo_profile_end every kernel this library launches from the calling thread is followed by a CUDA
 * event on its stream.  rpo_profile_end synchronises, writes one line per launch into buf --
 * "<source file:line>\t<microseconds since the previous launch finished>\t<tag: GEMM shape etc.>" --
 * and returns the number of launches (-1 on error). */
int rpo_profile_begin(void *stream);
int64_t rpo_profile_end(char *buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* RPO_B200_H */
