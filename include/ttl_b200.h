/* libttl_b200.so -- C ABI of the B200-native TTL per-sample test-time-adaptation path.
 *
 * The reference (Razaimam45/TTL-Test-Time-Low-Rank-Adaptation) has NO FFI for this path: it sits behind a Python
 * module API (clip/custom_clip.py + ttl.py).  This header is the boundary a reference maintainer would bind with
 * ctypes; each entry cites the reference interface it stands in for.  Plain pointers and
 * sizes only; no torch types.  Every function returns 0 on success or a negative TTL_E_* code and never throws;
 * the message is available from ttl_last_error().  No CPU fallback and no non-sm_100 fallback: ttl_create fails
 * with TTL_E_ARCH on any device whose compute capability is not 10.x.
 *
 * Threading: one context per (process, device); calls are not thread-safe.  All work is enqueued on the
 * caller's stream (a cudaStream_t passed as void*; NULL = legacy default stream) and nothing synchronises unless
 * stated ("host" variants and getters do).
 *
 * Ownership: the library owns weights, activations, workspaces, LoRA factors/gradients/AdamW moments and the
 * reset snapshot.  Callers own the image / logits buffers they pass and keep them alive until the stream has
 * consumed them.
 */
#ifndef TTL_B200_H_
#define TTL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTL_OK 0
#define TTL_E_INVALID (-1)
#define TTL_E_SHAPE (-2)
#define TTL_E_ARCH (-3)
#define TTL_E_CUDA (-4)
#define TTL_E_NOMEM (-5)
#define TTL_E_STATE (-6)

typedef struct ttl_ctx ttl_ctx;

/* Architecture + adapter geometry.  Stands in for the arguments of get_coop()/ClipTestTimeTuning.__init__
 * (clip/custom_clip.py:706-723, 571-624) and LoraConfig (clip/custom_clip.py:583-591). */
typedef struct ttl_config {
  int32_t image_size;    /* 224 */
  int32_t patch;         /* 16 (ViT-B/16), 14 (ViT-L/14) */
  int32_t width;         /* 768 / 1024; multiple of 128, head_dim must be 64 */
  int32_t layers;        /* 12 / 24 */
  int32_t heads;         /* 12 / 16 */
  int32_t mlp_dim;       /* 3072 / 4096 */
  int32_t proj_dim;      /* 512 / 768 */
  int32_t max_views;     /* views per forward call (--batch-size, ttl.py:389) */
  int32_t max_classes;   /* upper bound on C for buffer sizing */
  int32_t lora_rank;     /* 16 (or 32) */
  float lora_alpha;      /* 32 */
  int32_t lora_layer_lo; /* --layer_range, inclusive (ttl.py:402) */
  int32_t lora_layer_hi;
  float ln_eps;          /* 1e-5 */
  int32_t device;        /* CUDA ordinal (--gpu) */
  int32_t max_samples;   /* test samples adapted concurrently per call (BASELINE config 5); 0 or 1 = one at a time */
  int32_t precision;     /* enum ttl_precision: 0 = bf16 operands / fp32 accumulation on the tensor cores (the product);
                            1 = fp32 validation mode: every activation and contraction in fp32 on the CUDA cores, one sample
                            per call, no graphs -- held to the fp32 tolerance (1e-4) against the reference's fp32 CPU run */
  int32_t text_mode;     /* 1 = this context is the CLIP TEXT tower carrying the adapter (--lora_encoder text, ttl.py:145-149,
                            190-191; clip/custom_clip.py:602-606): width/layers/heads/mlp_dim/proj_dim describe the text tower,
                            max_views bounds the class prompts, max_classes the image views per test sample; image_size / patch
                            are ignored.  0 = image tower (the TTL configuration) */
  int32_t context;       /* text mode: tokens per prompt (77) */
  int32_t vocab;         /* text mode: 49408 */
} ttl_config;
enum ttl_precision { TTL_PRECISION_BF16 = 0, TTL_PRECISION_FP32 = 1 };

/* Frozen-weight slots; names follow the HF CLIP vision tower state_dict that
 * CLIPModel.from_pretrained (clip/custom_clip.py:581) yields. */
enum ttl_weight_kind {
  TTL_W_CLASS_EMB = 0, /* vision_model.embeddings.class_embedding           [d]           */
  TTL_W_PATCH_EMB = 1, /* vision_model.embeddings.patch_embedding.weight    [d,3,p,p]     */
  TTL_W_POS_EMB = 2,   /* vision_model.embeddings.position_embedding.weight [tokens,d]    */
  TTL_W_PRE_LN_G = 3, TTL_W_PRE_LN_B = 4,   /* vision_model.pre_layrnorm                  */
  TTL_W_POST_LN_G = 5, TTL_W_POST_LN_B = 6, /* vision_model.post_layernorm                */
  TTL_W_VIS_PROJ = 7,  /* visual_projection.weight [P,d]                                  */
  TTL_W_TOKEN_EMB = 8, /* text mode: text_model.embeddings.token_embedding.weight [vocab,d]; in text mode POS_EMB is
                          text_model.embeddings.position_embedding.weight [context,d], POST_LN_* the final_layer_norm and VIS_PROJ
                          text_projection.weight [P,d]; the per-layer slots below take text_model.encoder.layers.{i}.* */
  /* per encoder layer (layer index argument) */
  TTL_W_LN1_G = 16, TTL_W_LN1_B = 17,
  TTL_W_Q_W = 18, TTL_W_Q_B = 19, TTL_W_K_W = 20, TTL_W_K_B = 21, TTL_W_V_W = 22, TTL_W_V_B = 23,
  TTL_W_O_W = 24, TTL_W_O_B = 25,
  TTL_W_LN2_G = 26, TTL_W_LN2_B = 27,
  TTL_W_FC1_W = 28, TTL_W_FC1_B = 29, TTL_W_FC2_W = 30, TTL_W_FC2_B = 31
};

/* LoRA tensor slots of one layer, in the tuple order of LoRA_AB.init_weights (clip/custom_clip.py:193-200). */
enum ttl_lora_which { TTL_LORA_A_Q = 0, TTL_LORA_B_Q = 1, TTL_LORA_A_V = 2, TTL_LORA_B_V = 3 };
enum ttl_lora_what { TTL_LORA_PARAM = 0, TTL_LORA_GRAD = 1, TTL_LORA_INIT = 2 };

/* Loss heads of test_time_tuning (ttl.py:70-110). */
enum ttl_head {
  TTL_HEAD_TPT = 0,  /* top-p confidence selection + marginal entropy: ttl.py:50-61, 86-110 (deyo_selection falsy) */
  TTL_HEAD_DEYO = 1  /* weighted entropy over all views: deyo.py:93-196 with default flags (deyo_selection truthy) */
};

/* Optimiser / loop hyper-parameters (argparse defaults ttl.py:367-424; AdamW defaults ttl.py:218). */
typedef struct ttl_hparams {
  int32_t head;        /* enum ttl_head */
  int32_t tta_steps;   /* --tta_steps; the DeYO head performs tta_steps^2 optimiser steps like the reference */
  double selection_p;  /* --selection_p 0.1  -> K = (int)(V * p), evaluated in double like Python's int(V * top) at ttl.py:52 */
  float lr;            /* 5e-3 */
  float beta1, beta2;  /* 0.9, 0.999 */
  float eps;           /* 1e-8 */
  float weight_decay;  /* 1e-2 */
  float deyo_margin_e0;/* 0.4 */
} ttl_hparams;

/* Optional outputs of ttl_adapt_predict*: any pointer may be NULL.  The batch entry points return one block per
 * sample, sample-major (S = n_samples). */
typedef struct ttl_outputs {
  float* logits0;      /* [S,V,C] first-forward logits                                            */
  float* entropy;      /* [S,V]   per-view entropies                                              */
  int32_t* idx;        /* [S,K]   selected views (index within the sample) in argsort order (TPT) */
  float* loss;         /* [S]     loss of the last optimiser step                                 */
  float* pred_logits;  /* [S,C]   adapted prediction on view 0 (ttl.py:350-352)                   */
} ttl_outputs;

/* ---- lifetime ----------------------------------------------------------------------------------------- */
int ttl_create(ttl_ctx** out, const ttl_config* cfg);               /* get_coop(), clip/custom_clip.py:706 */
void ttl_destroy(ttl_ctx* ctx);
const char* ttl_last_error(const ttl_ctx* ctx);                     /* ctx may be NULL: last create error  */
int ttl_version(void);

/* ---- frozen state ------------------------------------------------------------------------------------- */
/* CLIPModel.from_pretrained(...) weights, fp32 host pointers, copied/converted once (clip/custom_clip.py:581). */
int ttl_set_weight(ttl_ctx* ctx, int32_t layer, int32_t kind, const float* host, int64_t numel);
/* Cached, L2-normalised text features [C,P] and logit_scale (log domain): get_text_features(),
 * clip/custom_clip.py:651-663 and :619 -- computed once per class-name set instead of twice per sample. */
int ttl_set_text_features(ttl_ctx* ctx, const float* host_text, int32_t n_classes, int32_t proj_dim, float logit_scale);

/* ---- adapter ------------------------------------------------------------------------------------------ */
/* LoRA_AB.initialize_layer_weights snapshot (clip/custom_clip.py:176-200): sets both the live factor and p0. */
int ttl_lora_set_init(ttl_ctx* ctx, int32_t layer, int32_t which, const float* host, int64_t numel);
/* ClipTestTimeTuning.LoRA_reset() + optimizer.load_state_dict(optim_state)  (ttl.py:338-344). */
int ttl_lora_reset(ttl_ctx* ctx, void* stream);
/* Synchronising getter: what = param | grad | init. */
int ttl_lora_get(ttl_ctx* ctx, int32_t layer, int32_t which, int32_t what, float* host_out, int64_t numel);
/* The same for sample `sample` of the last ttl_adapt_predict_batch* call (each concurrent sample owns its factors,
 * gradients and AdamW moments until the next call resets them; ttl.py:338-344 per sample). */
int ttl_lora_get_sample(ttl_ctx* ctx, int32_t sample, int32_t layer, int32_t which, int32_t what, float* host_out,
                        int64_t numel);
/* Device aliases of one tensor so a host framework can wrap them as parameters/gradients
 * (the names ttl.py:159-160,197-201 walk).  After writing through them call ttl_lora_touch(). */
int ttl_lora_device_ptr(ttl_ctx* ctx, int32_t layer, int32_t which, int32_t what, float** dev_ptr, int64_t* numel);
int ttl_lora_touch(ttl_ctx* ctx, void* stream);
/* torch.optim.AdamW.step() on the 4*n_layers LoRA tensors (ttl.py:105-108 via GradScaler, disabled in bf16). */
int ttl_adamw_step(ttl_ctx* ctx, const ttl_hparams* hp, void* stream);

/* ---- model calls -------------------------------------------------------------------------------------- */
/* ClipTestTimeTuning.forward/inference (clip/custom_clip.py:665-703): images fp32 [n_views,3,S,S] on device ->
 * logits fp32 [n_views,C] on device.  train != 0 keeps what ttl_backward needs (autograd-enabled forward). */
int ttl_forward(ttl_ctx* ctx, const float* images_dev, int32_t n_views, int32_t train, float* logits_dev, void* stream);
/* loss.backward() restricted to the LoRA factors (ttl.py:106): dlogits fp32 [n_views,C] of the last train forward. */
int ttl_backward(ttl_ctx* ctx, const float* dlogits_dev, void* stream);
/* The whole per-sample body of test_time_adapt_eval (ttl.py:338-352): reset -> test_time_tuning -> predict.
 * forced_idx_dev (nullable, TPT head) teacher-forces the selected views (selected_idx reuse, ttl.py:97-98). */
int ttl_adapt_predict(ttl_ctx* ctx, const float* images_dev, int32_t n_views, const ttl_hparams* hp,
                      const int32_t* forced_idx_dev, const ttl_outputs* out_dev, void* stream);
/* Same with HOST buffers (pinned recommended): H2D of the views, D2H of the requested outputs, stream sync. */
int ttl_adapt_predict_host(ttl_ctx* ctx, const float* images_host, int32_t n_views, const ttl_hparams* hp,
                           const int32_t* forced_idx_host, const ttl_outputs* out_host, void* stream);
/* n_samples independent test samples adapted concurrently (each: reset -> adapt -> predict with its own LoRA factors
 * and AdamW state, exactly as n_samples consecutive ttl_adapt_predict calls): images [n_samples, n_views, 3, S, S].
 * The frozen 64-view forward of all samples is one GEMM/attention pass; n_samples <= ttl_config.max_samples. */
int ttl_adapt_predict_batch(ttl_ctx* ctx, const float* images_dev, int32_t n_samples, int32_t n_views,
                            const ttl_hparams* hp, const int32_t* forced_idx_dev, const ttl_outputs* out_dev, void* stream);
int ttl_adapt_predict_batch_host(ttl_ctx* ctx, const float* images_host, int32_t n_samples, int32_t n_views,
                                 const ttl_hparams* hp, const int32_t* forced_idx_host, const ttl_outputs* out_host,
                                 void* stream);
/* As ttl_adapt_predict_batch_host but returns without synchronising: the views are copied on the library's copy stream
 * into one of two staging buffers, so the H2D of this call overlaps the kernels of the previous one (the role of the
 * reference's pin_memory + .cuda(non_blocking=True) loader, ttl.py:277,324-334).  images_host and out_host must stay
 * valid (pinned for real overlap) until `stream` has been synchronised. */
int ttl_adapt_predict_batch_host_async(ttl_ctx* ctx, const float* images_host, int32_t n_samples, int32_t n_views,
                                       const ttl_hparams* hp, const int32_t* forced_idx_host,
                                       const ttl_outputs* out_host, void* stream);
/* ---- adapter on the text tower: `--lora_encoder text` ---------------------------------------------
 * Reference: ttl.py:145-149 (requires-grad filter on text_encoder), :190-191 (optimizer groups over
 * text_encoder.text_model.encoder.layers), clip/custom_clip.py:602-606 (peft on the text tower), :672-678 (image features
 * under no_grad, class features with gradient).  Two contexts: an image-tower context (adapter unused) that yields the frozen
 * image features of the views, and a text-mode context (ttl_config.text_mode = 1) that owns the adapter. */
/* Image-tower context: images fp32 [n_views,3,S,S] on the device -> raw image features fp32 [n_views, P] on the device
 * (VisionEncoder.forward, clip/custom_clip.py:62-71; normalisation happens where they are used). */
int ttl_image_features(ttl_ctx* ctx, const float* images_dev, int32_t n_views, float* feats_dev, void* stream);
/* Text-mode context: the tokenised class prompts int32 [n_prompts, context] (clip.tokenize layout) and logit_scale (log
 * domain); runs the layers below the adapter once (reset_classnames, clip/custom_clip.py:343-372). */
int ttl_text_set_prompts(ttl_ctx* ctx, const int32_t* tokens_host, int32_t n_prompts, float logit_scale, void* stream);
/* L2-normalised class features fp32 [n_prompts, P] with the current factors (get_text_features, :651-663); synchronises. */
int ttl_text_features(ttl_ctx* ctx, float* feats_host, void* stream);
/* One test sample (ttl.py:338-352 with lora_encoder == 'text'): reset -> tta steps over the class features -> prediction for
 * view 0.  img_feats_dev: what ttl_image_features returned for the sample's views.  Outputs as ttl_adapt_predict with S = 1,
 * C = n_prompts.  The ttl_lora_* entry points address the text-tower factors of this context. */
int ttl_text_adapt_predict(ttl_ctx* ctx, const float* img_feats_dev, int32_t n_views, const ttl_hparams* hp,
                           const int32_t* forced_idx_dev, const ttl_outputs* out_dev, void* stream);

/* ---- optional branches of the weighted-entropy head ----------------------------------------
 * deyo.py:103-151 behind the flags of ttl.py:410-424, for n_samples concurrent test samples, fused like ttl_adapt_predict_batch:
 *   filter_ent    keep the int(V * selection_p) lowest-entropy views instead of every view with H <= ln 1000 (deyo.py:103-108)
 *   filter_plpd   second forward on x' = the kept views with their structure destroyed (--aug_type occ | patch | pixel,
 *                 deyo.py:115-136); keep the views whose top-class probability drops by more than plpd_threshold (:137-148)
 *   reweight_ent / reweight_plpd   deyo.py:159-179: coefficient reweight_ent * exp(-(H - margin)) when either is set, else 1
 * A sample whose views are all filtered out takes no optimiser step (deyo.py:184).  The head runs tta_steps^2 times like the
 * default one.  The random draws stay on the host (the torch seed decides, as in the reference): for aug_type patch,
 * perm_host = int32 [n_samples][tta_steps^2][n_kept][patch_len^2], row = torch.argsort(torch.rand(patch_len^2)) (deyo.py:127),
 * n_kept = filter_ent ? int(V * selection_p) : V; for pixel, int32 [n_samples][tta_steps^2][image_size^2] = torch.randperm;
 * NULL for occ.  Runs eagerly (no graph replay).  out_dev->idx receives the kept views of the first step when filter_ent. */
enum ttl_aug_type { TTL_AUG_OCC = 0, TTL_AUG_PATCH = 1, TTL_AUG_PIXEL = 2 };
typedef struct ttl_deyo_options {
  int32_t filter_ent, filter_plpd, reweight_ent, reweight_plpd;
  float plpd_threshold;     /* --plpd_threshold 0.2 */
  int32_t aug_type;         /* enum ttl_aug_type */
  int32_t occlusion_size, row_start, column_start, patch_len;   /* 112, 56, 56, 6 */
  const int32_t* perm_host;
  int64_t perm_numel;
  const int32_t* forced_keep_host;  /* nullable, for parity tests: [n_samples][n_kept] 0/1 flags that replace the outcome of the
                                       PLPD filter (PLPD is still computed and reported); the analogue of forced_idx for the
                                       selection, needed because a threshold inside the bf16 noise of the PLPD values flips views */
} ttl_deyo_options;
int ttl_adapt_predict_batch_deyo(ttl_ctx* ctx, const float* images_dev, int32_t n_samples, int32_t n_views,
                                 const ttl_hparams* hp, const ttl_deyo_options* opt, const ttl_outputs* out_dev, void* stream);
/* PLPD values [n_samples * n_kept] and kept-view counts [n_samples] of the last optimiser step of that call (synchronises). */
int ttl_deyo_last_plpd(ttl_ctx* ctx, float* plpd_host, int32_t* n_final_host, int32_t n_samples, int32_t n_kept);

/* ---- view generator on the device ------------------------------------------------------
 * Stands in for the host-side AugMixAugmenter the reference's DataLoader workers run per test image
 * (data/datautils.py:98-157 with the empty augmentation list; ttl.py:232-241): the caller ships the decoded uint8 RGB
 * image [H,W,3] plus one spec per view, and the library resamples on the GPU, bit-exactly as Pillow's 8-bit antialiased
 * resampler + torchvision's ToTensor/Normalize would on the host.
 *   TTL_VIEW_CLEAN: Resize(image_size, BICUBIC, antialias) + CenterCrop(image_size)   (ttl.py:232-234); box ignored
 *   TTL_VIEW_CROP : RandomResizedCrop box (top, left, height, width) as drawn by torchvision's
 *                   RandomResizedCrop.get_params, bilinear resize to image_size, optional horizontal flip
 *                   (data/datautils.py:98-101).  The RNG stays on the host so the torch seed (ttl.py:114) still decides. */
enum ttl_view_kind { TTL_VIEW_CLEAN = 0, TTL_VIEW_CROP = 1 };
typedef struct ttl_view_spec { int32_t kind, top, left, height, width, flip; } ttl_view_spec;
/* transforms.Normalize(mean, std) of ttl.py:226-227; the CLIP constants are the default. */
int ttl_set_pixel_norm(ttl_ctx* ctx, const float* mean3, const float* std3);
/* n_images host images (uint8 [H_i, W_i, 3]) x n_views specs each (specs_host [n_images, n_views]) ->
 * normalised fp32 views on the device [n_images * n_views, 3, S, S].  Enqueued on `stream`; the host buffers may be
 * reused once the call returns only if they are pageable (pinned buffers: after the stream has been synchronised). */
int ttl_make_views(ttl_ctx* ctx, const uint8_t* const* images_host, const int32_t* heights, const int32_t* widths,
                   int32_t n_images, const ttl_view_spec* specs_host, int32_t n_views, float* views_dev, void* stream);
/* ttl_adapt_predict_batch_host_async fed by uint8 images: H2D of the images (not of 64 fp32 views each), view
 * generation straight into the patch-embedding operand (bf16 patch matrix; the fp32 views never exist), then the
 * adapt/predict graph.  Same staging/overlap and lifetime rules as ttl_adapt_predict_batch_host_async. */
int ttl_adapt_predict_images_async(ttl_ctx* ctx, const uint8_t* const* images_host, const int32_t* heights,
                                   const int32_t* widths, int32_t n_samples, const ttl_view_spec* specs_host,
                                   int32_t n_views, const ttl_hparams* hp, const int32_t* forced_idx_host,
                                   const ttl_outputs* out_host, void* stream);

/* ---- class-feature builder: the text tower, once per class-name set --------------------------
 * Stands in for ClipTestTimeTuning.get_text_features (clip/custom_clip.py:651-663) -> PromptEncoder (:73-82) -> HF
 * CLIPModel.get_text_features, which the reference re-runs inside every forward although nothing in it is trainable on
 * the TTL path.  Own context (own geometry and weights); the result feeds ttl_set_text_features. */
typedef struct ttl_text_ctx ttl_text_ctx;
typedef struct ttl_text_config {
  int32_t vocab;        /* 49408 */
  int32_t context;      /* 77 (<= 128) */
  int32_t width;        /* 512 (B/16) / 768 (L/14); multiple of 128, head_dim 64 */
  int32_t layers;       /* 12 */
  int32_t heads;        /* 8 / 12 */
  int32_t mlp_dim;      /* 2048 / 3072 */
  int32_t proj_dim;     /* 512 / 768 */
  int32_t max_prompts;  /* prompts encoded per pass (buffer sizing; longer lists are chunked) */
  float ln_eps;         /* 1e-5 */
  int32_t device;
} ttl_text_config;
/* Non-layer slots (HF names: text_model.embeddings.token_embedding.weight [vocab,d], .position_embedding.weight [ctx,d],
 * text_model.final_layer_norm.{weight,bias}, text_projection.weight [P,d]); per-layer slots reuse TTL_W_LN1_G..TTL_W_FC2_B. */
enum ttl_text_weight_kind { TTL_TW_TOKEN_EMB = 0, TTL_TW_POS_EMB = 1, TTL_TW_FINAL_LN_G = 2, TTL_TW_FINAL_LN_B = 3, TTL_TW_TEXT_PROJ = 4 };
int ttl_text_create(ttl_text_ctx** out, const ttl_text_config* cfg);
void ttl_text_destroy(ttl_text_ctx* ctx);
const char* ttl_text_last_error(const ttl_text_ctx* ctx);
int ttl_text_set_weight(ttl_text_ctx* ctx, int32_t layer, int32_t kind, const float* host, int64_t numel);
/* tokens_host int32 [n_prompts, context] (clip.tokenize layout: SOT, ids, EOT = highest id, zero padding) ->
 * feats_host fp32 [n_prompts, proj_dim], L2-normalised.  Synchronises `stream`. */
int ttl_text_encode(ttl_text_ctx* ctx, const int32_t* tokens_host, int32_t n_prompts, float* feats_host, void* stream);

/* Toggle CUDA-graph replay of ttl_adapt_predict (default on). */
int ttl_set_graphs(ttl_ctx* ctx, int32_t enabled);
/* Kernel launches issued by the last ttl_adapt_predict* call (for bench.py's gpu_launches). */
int64_t ttl_last_launch_count(const ttl_ctx* ctx);

/* Per-launch timing of the dominant kernel (the tcgen05 GEMM) with CUDA events on the launching stream; while enabled
 * ttl_adapt_predict runs eagerly (no graph replay).  ttl_profile_read synchronises, returns and clears the records. */
typedef struct ttl_gemm_record { int32_t M, N, K, epi; float ms; } ttl_gemm_record;
int ttl_profile_gemm(ttl_ctx* ctx, int32_t enable);
int ttl_profile_read(ttl_ctx* ctx, ttl_gemm_record* out, int32_t max_records, int32_t* n_records);

/* ---- head pieces (select_confident_samples ttl.py:50-54, avg_entropy ttl.py:56-61, deyo.py:85-181) ---------- */
int ttl_op_logits_entropy(const float* feats_dev, const float* text_dev, float scale, float* logits_dev,
                          float* entropy_dev, int32_t V, int32_t C, int32_t P, void* stream);
/* batch_entropy of select_confident_samples (ttl.py:51) / softmax_entropy (deyo.py:85-90) */
int ttl_op_entropy(const float* logits_dev, float* entropy_dev, int32_t V, int32_t C, void* stream);
int ttl_op_select(const float* entropy_dev, int32_t V, int32_t K, int32_t* idx_dev, void* stream);
int ttl_op_tpt_loss(const float* logits_dev, const int32_t* idx_dev, int32_t K, int32_t C, float* loss_dev,
                    float* dlogits_dev, void* stream);
int ttl_op_deyo_loss(const float* logits_dev, int32_t V, int32_t C, float margin_e0, float* loss_dev,
                     float* dlogits_dev, void* stream);

/* ---- single kernels, exposed for parity tests --------------------------------------------------------- */
/* C[M,N] = epi(A[M,K] B[N,K]^T (+ A2[M,K2] B2[N,K2]^T)); bf16 operands (uint16 storage), see csrc/gemm.cuh. */
int ttl_op_gemm(const void* a, const void* b, const void* a2, const void* b2, int32_t M, int32_t N, int32_t K,
                int32_t K2, int32_t epi, const float* bias, void* out, void* out2, const float* resid,
                const void* aux, const float* pos, int32_t tokens_per_view, int32_t block_n, void* stream);
int ttl_op_layernorm(const float* x, void* y_bf16, const float* gamma, const float* beta, int32_t rows, int32_t d,
                     float eps, void* stream);
int ttl_op_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* dres, float* dx,
                         void* dx_bf16, int32_t rows, int32_t d, float eps, void* stream);
int ttl_op_attention_fwd(const void* qkv_bf16, void* out_bf16, float* lse, int32_t V, int32_t tokens, int32_t heads,
                         float scale, void* stream);
int ttl_op_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse,
                         void* dqkv_bf16, int32_t V, int32_t tokens, int32_t heads, float scale, void* stream);
int ttl_op_im2col(const float* images, void* patches_bf16, int32_t V, int32_t S, int32_t p, void* stream);
int ttl_op_adamw(float* p, const float* g, float* m, float* v, int32_t n, int32_t step, float lr, float b1, float b2,
                 float eps, float wd, void* stream);
int ttl_op_skinny_reduce(const void* wide_bf16, int32_t ldw, int32_t nw, const void* narrow_bf16, int32_t ldn,
                         int32_t nn, int32_t M, float scale, float* out, int32_t transpose_out, float* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTL_B200_H_ */
