// Launchers of the non-GEMM kernels (rowops.cu, attention.cu, head.cu, lora.cu).  All enqueue on the given
// stream and never synchronise.  Shapes use the reference's vocabulary: a "view" is one augmented crop of the
// test image (ttl.py:324-336), a view has `tokens` rows (CLS + patches) of width d.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ttl {

typedef __nv_bfloat16 bf16;

// ---- rowops.cu
// images fp32 [V,3,S,S] -> patches bf16 [V*T, 3*p*p] in conv-weight order (c, i, j)  (HF CLIPVisionEmbeddings conv k=p s=p)
void launch_im2col(const float* images, bf16* patches, int V, int S, int p, cudaStream_t st);
// x[V*tokens, d] holds patch rows = conv + pos (GEMM epilogue); writes CLS rows = cls + pos[0], then pre-LN in place.
void launch_embed_preln(float* x, const float* cls, const float* pos, const float* gamma, const float* beta, int V,
                        int tokens, int d, float eps, cudaStream_t st);
// `descending` (here, launch_attention_fwd, GemmArgs::descending): walk the rows / units from the last to the first, so a consumer
// starts on what its producer wrote last and still sits in L2 (engine.cu zigzag).  Same arithmetic per row, same result.
// y(bf16) = LN(x) * gamma + beta, one warp per row, fp32 statistics.
void launch_layernorm(const float* x, bf16* y, const float* gamma, const float* beta, int rows, int d, float eps,
                      cudaStream_t st, int descending = 0);
// dx = dres + LN'(x)^T (dy * gamma); statistics recomputed from x.  dres nullable.  dx_bf16 nullable.
void launch_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* dres, float* dx,
                          bf16* dx_bf16, int rows, int d, float eps, cudaStream_t st);
// dst[g*tokens + t, :] = src[v(g)*tokens + t, :],  v(g) = view_idx[g] (device array) or g * view_stride when view_idx == nullptr
// sample_views > 0 (concurrent samples): n_sel = S*K entries, entry g reads view (g / k_per_sample) * sample_views + view_idx[g]
void launch_gather_views(const float* src, float* dst, const int* view_idx, int n_sel, int view_stride, int tokens, int d,
                         cudaStream_t st, int sample_views = 0, int k_per_sample = 1);
// x bf16 [M, 64*S]: zero every 64-column block except the one of the row's own sample (row / rows_per_sample)
void launch_block_mask(bf16* x, int M, int ncols, int rows_per_sample, cudaStream_t st);

// ---- attention.cu   qkv bf16 [V*tokens, 3d] (q | k | v, heads concatenated, 64 per head)
// out bf16 [V*tokens, d]; lse (nullable) fp32 [V, heads, tokens] natural-log softmax normaliser of scale*q.k
// causal != 0: key j is visible to query i only for j <= i (text tower); served by the general mma.sync kernels
void launch_attention_fwd(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                          cudaStream_t st, int descending = 0, int causal = 0);
// dqkv bf16 [V*tokens, 3d]  from dout bf16 [V*tokens, d], qkv, out, lse
// delta_ws (nullable; fp32 [V, heads, tokens], tokens <= 128): Delta = rowsum(P o dP) is computed exactly into it first instead of
// being taken as rowsum(dO o O) from the bf16 O (text tower: strong cancellation in dP - Delta, see attention_delta_kernel)
void launch_attention_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int V,
                          int tokens, int heads, float scale, cudaStream_t st, int causal = 0, float* delta_ws = nullptr);
// CLS-query attention of the last layer in inference: q_cls bf16 [V, d] (one query row per view), K/V from qkv rows;
// out_cls bf16 [V, d]
void launch_attention_cls(const bf16* q_cls, const bf16* qkv, bf16* out_cls, int V, int tokens, int heads, float scale,
                          cudaStream_t st);
size_t attention_fwd_smem(int tokens);
size_t attention_bwd_smem(int tokens);

// ---- head.cu
// pooled = LN(x[v*tokens + 0, :]) ; feats[v,:] = Wp[P,d] @ pooled  (HF post_layernorm + visual_projection)
// pool_row (nullable): row of each sequence that is pooled (text tower: the EOT position) instead of row 0
void launch_pool_project(const float* x, const float* gamma, const float* beta, const float* Wp, float* pooled,
                         float* feats, int V, int tokens, int d, int P, float eps, cudaStream_t st, const int* pool_row = nullptr);
// y = x / |x| row-wise, [rows, P]
void launch_l2norm_rows(const float* x, float* y, int rows, int P, cudaStream_t st);
// `--lora_encoder text`: gradient of the loss w.r.t. the last text-tower hidden state from dlogits [K, C] (rows = the views
// idx[k], or k when idx == nullptr), the L2-normalised image features fhat [V, P] and the raw class features tfeats [C, P]:
// dx fp32 [C*tokens, d] (+ bf16 copy), non-zero on the EOT rows only
void launch_text_head_bwd(const float* dlogits, const int* idx, int K, const float* fhat, float scale, const float* tfeats,
                          const float* Wp, const float* x, const float* gamma, const int* eot, float* dfh, float* dpool, float* dx,
                          bf16* dx_bf16, int C, int P, int tokens, int d, float eps, cudaStream_t st);
// logits[v,c] = scale * <feats[v]/|feats[v]|, T[c]> ; entropy[v] = H(softmax(logits[v]))   (custom_clip.py:680-687, ttl.py:51)
void launch_logits_entropy(const float* feats, const float* text, float scale, float* logits, float* entropy, int V,
                           int C, int P, cudaStream_t st);
// entropy[v] = H(softmax(logits[v]))  (ttl.py:51 / deyo.py:85-90); logits are read-only in effect (scaled by 1)
void launch_entropy(float* logits, float* entropy, int V, int C, cudaStream_t st);
// idx[0..K) = argsort(entropy, stable)[:K]  (ttl.py:52; ties -> lowest index).  forced_idx (nullable) overrides.
// n_samples > 1: sample s uses entropy + s*V, idx + s*K (forced_idx likewise), one CTA each
void launch_select(const float* entropy, int V, int K, const int* forced_idx, int* idx, cudaStream_t st, int n_samples = 1);
// marginal-entropy loss of the K rows logits[idx[k]] (ttl.py:56-61) and its gradient, compact: dlogits[K,C]
// n_samples > 1: sample s uses logits + s*logits_sstride, idx + s*K, loss + s, dlogits + s*K*C, one CTA each
void launch_tpt_loss(const float* logits, const int* idx, int K, int C, float* loss, float* dlogits, cudaStream_t st,
                     int n_samples = 1, size_t logits_sstride = 0);
// weighted-entropy (DeYO, default flags) loss over all V rows and gradient dlogits[V,C]  (deyo.py:97-181)
void launch_deyo_loss(const float* logits, int V, int C, float margin_e0, float* loss, float* dlogits, cudaStream_t st,
                      int n_samples = 1);
// head backward for G compact views: dlogits[G,C] -> dx[G*tokens, d] (fp32, zero except CLS rows) + bf16 copy.
void launch_head_bwd(const float* dlogits, const float* text, float scale, const float* feats, const float* Wp,
                     const float* x, const float* gamma, float* dfh, float* dpool, float* dx, bf16* dx_bf16, int G, int C,
                     int P, int tokens, int d, float eps, cudaStream_t st);

// ---- lora.cu   per-layer fp32 master tensors in the reference's tuple order (A_q[r,d], B_q[d,r], A_v[r,d], B_v[d,r])
struct LoraPacked {       // bf16 operands consumed by the GEMM's second operand pair; kc = 64 * S (64 = padded 2r per sample)
  bf16* a_ext;            // [kc, d]   per sample block: rows 0..r-1 = A_q, r..2r-1 = A_v, rest 0   (T = h1 @ a_ext^T)
  bf16* a_ext_t;          // [d, kc]   transpose of a_ext                                  (dh1 += U @ a_ext_t^T)
  bf16* b_ext;            // [3d, kc]  per sample block: rows of q: s*B_q in cols 0..r-1; rows of v: s*B_v in cols r..2r-1
  bf16* b_ext_t;          // [kc, 3d]  transpose of b_ext                                  (U = dqkv @ b_ext_t^T)
};
// params: this layer's tensors of sample 0; sample i at params + i * sample_stride.  S samples are K-concatenated.
void launch_lora_pack(const float* params, int64_t sample_stride, LoraPacked pk, int d, int r, float s, int S, cudaStream_t st);
// all LoRA layers in one launch: layer i reads params + i*layer_stride and writes pks[i]  (n_layers <= 16 per launch)
void launch_lora_pack_layers(const float* params, int64_t layer_stride, int64_t sample_stride, const LoraPacked* pks,
                             int n_layers, int d, int r, float s, int S, cudaStream_t st);
// out[w, j] (or out[j, w] if transpose_out) = scale * sum_m Wd[m, w] * Nr[m, j];  w < nw (multiple of 64), j < 16 * nj16
// Deterministic two-pass reduction (partials in workspace `ws`, >= groups*ceil(M/128)*nw*nn floats).
// groups > 1: group g reduces rows [g*M, (g+1)*M) with the narrow operand shifted by g*narrow_gstride columns and writes
// out + g*out_gstride (the per-sample weight gradients of concurrently adapted samples).
void launch_skinny_reduce(const bf16* wide, int ldw, int nw, const bf16* narrow, int ldn, int nn, int M, float scale,
                          float* out, int transpose_out, float* ws, int groups, int narrow_gstride, int64_t out_gstride,
                          cudaStream_t st);
// Fused AdamW (torch.optim.AdamW rule, ttl.py:218 defaults) over n contiguous fp32 elements.
void launch_adamw(float* p, const float* g, float* m, float* v, int n, int step, float lr, float b1, float b2,
                  float eps, float wd, cudaStream_t st);
// p[i] <- p0[i % n0], m <- 0, v <- 0   (LoRA_AB.reset + optimizer.load_state_dict(optim_state), ttl.py:338-344; n = S * n0)
void launch_lora_reset(float* p, const float* p0, float* m, float* v, int n, int n0, cudaStream_t st);

// ---- fp32.cu   fp32 validation mode (ttl_config.precision = TTL_PRECISION_FP32): CUDA-core kernels, fp32 everywhere
enum SgemmEpi : int {
  SE_LINEAR = 0,    // out = alpha*acc (+ bias[n]) (+ resid[m,n])
  SE_GELU = 1,      // z = alpha*acc + bias; out2 (optional) = z; out = z*sigmoid(1.702 z)
  SE_GELU_BWD = 2,  // out = alpha*acc * dQuickGELU(aux[m,n])
  SE_PATCH = 3      // patch-embed scatter: row (view, patch) -> token row view*(T+1)+1+patch, + pos[1+patch, n]
};
struct SgemmArgs {
  const float* A = nullptr; int lda = 0;       // [M, K], K contiguous
  const float* B = nullptr; int ldb = 0;       // b_kn == 0: [N, K] (K contiguous);  b_kn == 1: [K, N] (N contiguous)
  int b_kn = 0;
  int M = 0, N = 0, K = 0;
  float alpha = 1.f;
  const float* bias = nullptr;
  const float* resid = nullptr; int ldr = 0;
  float* out = nullptr; int ldo = 0;
  float* out2 = nullptr;
  const float* aux = nullptr;
  const float* pos = nullptr; int tpv = 0;
  int epi = SE_LINEAR;
};
cudaError_t launch_sgemm(const SgemmArgs& a, cudaStream_t st);
// qkv fp32 [V*tokens, 3d] -> out fp32 [V*tokens, d], lse (nullable) [V, heads, tokens]
void launch_attention_f32_fwd(const float* qkv, float* out, float* lse, int V, int tokens, int heads, float scale, cudaStream_t st,
                              int causal = 0);
void launch_attention_f32_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int V,
                              int tokens, int heads, float scale, cudaStream_t st, int causal = 0);
size_t attention_f32_fwd_smem(int tokens);
size_t attention_f32_bwd_smem(int tokens);
void launch_layernorm_f32(const float* x, float* y, const float* gamma, const float* beta, int rows, int d, float eps,
                          cudaStream_t st);
void launch_im2col_f32(const float* images, float* patches, int V, int S, int p, cudaStream_t st);
// out[w, j] (or out[j, w] if transpose_out) = scale * sum_m wide[m, w] * narrow[m, j]
void launch_reduce_tn_f32(const float* wide, int ldw, int nw, const float* narrow, int ldn, int nn, int M, float scale,
                          float* out, int transpose_out, cudaStream_t st);

// ---- text.cu   x[m, :] = token_embedding[tokens[m], :] + position_embedding[m % ctx, :]
void launch_text_embed(const int* tokens, const float* tok_emb, const float* pos_emb, float* x, int rows, int ctx, int d, int vocab,
                       cudaStream_t st);

// ---- deyo.cu   optional branches of the weighted-entropy head (deyo.py:103-151).  Kept entry b of sample s is view
// s * V + idx[s * n1 + b] (idx == nullptr: b); x' fp32 [S * n1, 3, size, size]
void launch_destroy_occ(const float* images, const int* idx, float* xprime, int S, int V, int n1, int size, int occ, int r0, int c0,
                        cudaStream_t st);
void launch_destroy_pixel(const float* images, const int* idx, const int* perm, float* xprime, int S, int V, int n1, int size,
                          cudaStream_t st);
void launch_destroy_patch(const float* images, const int* idx, const int* perm, float* scratch, float* xprime, int S, int V, int n1,
                          int size, int patch_len, cudaStream_t st);
void launch_plpd(const float* logits, const float* logits_prime, int G, int C, float thr, int* keep, float* plpd_out, cudaStream_t st);
void launch_deyo_general_loss(const float* logits, const int* keep, int n1, int C, float e0, int filter_ent, int reweight,
                              float reweight_ent, float* loss, float* dlogits, int* active, int* steps, int* n_kept, int S,
                              cudaStream_t st);
void launch_adamw_masked(float* p, const float* g, float* m, float* v, int n_per_sample, int S, const int* active, const int* steps,
                         float lr, float b1, float b2, float eps, float wd, cudaStream_t st);

}  // namespace ttl
