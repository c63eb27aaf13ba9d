// Class-feature builder ("next" row N2): the CLIP text tower, run ONCE per class-name set on the device.
// The reference recomputes it inside every forward -- twice per test sample, 5.96 TFLOP per pass at 1000 classes
// (clip/custom_clip.py:651-663 get_text_features -> PromptEncoder :73-82 -> HF CLIPModel.get_text_features); nothing in
// it is trainable on the TTL path (--lora_encoder image), so its result is a constant of the dataset.
//   tokens [n, ctx] -> token_embedding + position_embedding -> L pre-LN layers (causal self-attention, QuickGELU MLP)
//   -> final_layer_norm on the EOT position (argmax of the token ids, HF CLIPTextTransformer) -> text_projection -> L2 norm.
// The dense contractions reuse the tcgen05 GEMM of the image tower (gemm.cu) and its LayerNorm; the 77-token causal
// attention is a small warp-per-query-row kernel (fp32 scores and softmax, K/V of one (prompt, head) staged in smem).
#include "../../include/ttl_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

using namespace ttl;

namespace {

thread_local std::string g_text_err;

constexpr int DH = 64;

// x[m, :] = token_embedding[tokens[m], :] + position_embedding[m % ctx, :]
__global__ void text_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ tok_emb,
                                  const float* __restrict__ pos_emb, float* __restrict__ x, int rows, int ctx, int d4, int vocab) {
  pdl_wait();
  pdl_trigger();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < static_cast<size_t>(rows) * d4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / d4), c = static_cast<int>(i % d4);
    int t = tokens[m];
    t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);
    const float4 a = reinterpret_cast<const float4*>(tok_emb)[static_cast<size_t>(t) * d4 + c];
    const float4 b = reinterpret_cast<const float4*>(pos_emb)[static_cast<size_t>(m % ctx) * d4 + c];
    reinterpret_cast<float4*>(x)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// Causal self-attention of one (prompt, head): qkv bf16 [rows, 3d] -> out bf16 [rows, d].  One warp per query row;
// lane j scores keys j, j+32, j+64 (<= row), fp32 softmax, then every lane accumulates two output dimensions.
__global__ void __launch_bounds__(256)
text_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int ctx, int heads, float scale) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm_text[];
  const int ldk = DH + 1;                       // +1: lanes read different key rows of the same column
  float* sK = sm_text;                          // [ctx][65]
  float* sV = sK + ctx * ldk;                   // [ctx][65]
  float* sQ = sV + ctx * ldk;                   // [warps][64]
  const int h = blockIdx.x, prompt = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bf16* base = qkv + static_cast<size_t>(prompt) * ctx * ld + h * DH;
  for (int i = threadIdx.x; i < ctx * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    sK[r * ldk + c] = __bfloat162float(base[static_cast<size_t>(r) * ld + d + c]);
    sV[r * ldk + c] = __bfloat162float(base[static_cast<size_t>(r) * ld + 2 * d + c]);
  }
  __syncthreads();
  float* q = sQ + warp * DH;
  const int nchunk = (ctx + 31) / 32;
  for (int r = warp; r < ctx; r += nw) {
    q[lane] = __bfloat162float(base[static_cast<size_t>(r) * ld + lane]);
    q[lane + 32] = __bfloat162float(base[static_cast<size_t>(r) * ld + lane + 32]);
    __syncwarp();
    float s[4];                                 // ctx <= 128
    float mx = -INFINITY;
    for (int cidx = 0; cidx < nchunk; ++cidx) {
      const int j = cidx * 32 + lane;
      float acc = -INFINITY;
      if (j <= r) {
        acc = 0.f;
        const float* kr = sK + j * ldk;
#pragma unroll 16
        for (int c = 0; c < DH; ++c) acc += q[c] * kr[c];
        acc *= scale;
      }
      s[cidx] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int cidx = 0; cidx < nchunk; ++cidx) {
      s[cidx] = s[cidx] == -INFINITY ? 0.f : __expf(s[cidx] - mx);
      sum += s[cidx];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float o0 = 0.f, o1 = 0.f;
    for (int cidx = 0; cidx < nchunk; ++cidx) {
      const int jmax = min(32, r + 1 - cidx * 32);
      for (int jj = 0; jj < jmax; ++jj) {
        const float p = __shfl_sync(0xffffffffu, s[cidx], jj);
        const float* vr = sV + (cidx * 32 + jj) * ldk;
        o0 += p * vr[lane];
        o1 += p * vr[lane + 32];
      }
    }
    bf16* orow = out + (static_cast<size_t>(prompt) * ctx + r) * d + h * DH;
    orow[lane] = __float2bfloat16(o0 * inv);
    orow[lane + 32] = __float2bfloat16(o1 * inv);
    __syncwarp();
  }
}

// xg[p, :] = x[p*ctx + eot[p], :]   with eot[p] = first argmax of tokens[p, :]   (HF CLIPTextTransformer pooling)
__global__ void text_gather_eot_kernel(const int* __restrict__ tokens, const float* __restrict__ x, float* __restrict__ xg,
                                       int ctx, int d) {
  pdl_wait();
  pdl_trigger();
  const int p = blockIdx.x;
  __shared__ int s_eot;
  if (threadIdx.x == 0) {
    int best = 0, bv = tokens[static_cast<size_t>(p) * ctx];
    for (int j = 1; j < ctx; ++j) {
      const int v = tokens[static_cast<size_t>(p) * ctx + j];
      if (v > bv) { bv = v; best = j; }
    }
    s_eot = best;
  }
  __syncthreads();
  const float* src = x + (static_cast<size_t>(p) * ctx + s_eot) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) xg[static_cast<size_t>(p) * d + i] = src[i];
}

struct TextLayer {
  bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr, *ln1g = nullptr, *ln1b = nullptr, *ln2g = nullptr, *ln2b = nullptr;
};

inline uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace

struct ttl_text_ctx {
  ttl_text_config cfg{};
  int d = 0, F = 0, P = 0, L = 0, H = 0, ctx = 0, vocab = 0, maxp = 0, num_sms = 148;
  std::string err;
  std::vector<void*> allocs;
  float *tok = nullptr, *pos = nullptr, *lnfg = nullptr, *lnfb = nullptr, *proj = nullptr;
  std::vector<TextLayer> lw;
  int* tokens = nullptr;
  float *XA = nullptr, *XB = nullptr, *XG = nullptr, *pooled = nullptr, *feats = nullptr;
  bf16 *Hb = nullptr, *QKV = nullptr, *AO = nullptr, *Gb = nullptr;
};

namespace {

#define TCK(expr)                                                            \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      c->err = std::string(#expr) + ": " + cudaGetErrorString(_e);           \
      return TTL_E_CUDA;                                                     \
    }                                                                        \
  } while (0)

template <typename Tp>
int talloc(ttl_text_ctx* c, Tp** p, size_t n) {
  void* q = nullptr;
  if (cudaMalloc(&q, n * sizeof(Tp) + 256) != cudaSuccess) {
    cudaGetLastError();
    c->err = "cudaMalloc failed";
    return TTL_E_NOMEM;
  }
  cudaMemset(q, 0, n * sizeof(Tp) + 256);
  c->allocs.push_back(q);
  *p = static_cast<Tp*>(q);
  return TTL_OK;
}

int up_f32(ttl_text_ctx* c, float* dst, const float* host, int64_t n) {
  TCK(cudaMemcpy(dst, host, sizeof(float) * n, cudaMemcpyHostToDevice));
  return TTL_OK;
}
int up_bf16(ttl_text_ctx* c, bf16* dst, const float* host, int64_t n) {
  std::vector<uint16_t> tmp(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) tmp[i] = f2bf(host[i]);
  TCK(cudaMemcpy(dst, tmp.data(), sizeof(uint16_t) * n, cudaMemcpyHostToDevice));
  return TTL_OK;
}

GemmOperand opnd(const bf16* p, int rows, int k, int ld) {
  GemmOperand o;
  o.ptr = p; o.rows = rows; o.k = k; o.ld = ld;
  return o;
}

int tgemm(ttl_text_ctx* c, GemmArgs& g, cudaStream_t st) {
  cudaError_t e = gemm_launch(g, st, c->num_sms);
  if (e != cudaSuccess) {
    c->err = std::string("gemm: ") + gemm_last_error() + " / " + cudaGetErrorString(e);
    return e == cudaErrorInvalidValue ? TTL_E_SHAPE : TTL_E_CUDA;
  }
  return TTL_OK;
}

// n prompts (<= maxp) whose tokens are in c->tokens -> c->feats [n, P] (not yet normalised)
int encode_chunk(ttl_text_ctx* c, int n, cudaStream_t st) {
  const int M = n * c->ctx, d = c->d, F = c->F;
  {
    const size_t total = static_cast<size_t>(M) * (d / 4);
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(text_embed_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const int*>(c->tokens),
               static_cast<const float*>(c->tok), static_cast<const float*>(c->pos), c->XA, M, c->ctx, d / 4, c->vocab);
  }
  const size_t att_smem = (2 * static_cast<size_t>(c->ctx) * (DH + 1) + 8 * DH) * sizeof(float);
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (att_smem > configured) {
    cudaFuncSetAttribute(text_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(att_smem));
    configured = att_smem;
  }
  for (int l = 0; l < c->L; ++l) {
    const TextLayer& w = c->lw[l];
    launch_layernorm(c->XA, c->Hb, w.ln1g, w.ln1b, M, d, c->cfg.ln_eps, st);
    {
      GemmArgs g;
      g.a1 = opnd(c->Hb, M, d, d); g.b1 = opnd(w.wqkv, 3 * d, d, d);
      g.M = M; g.N = 3 * d; g.epi = EPI_BF16; g.bias = w.bqkv; g.out = c->QKV; g.ldo = 3 * d;
      if (int r = tgemm(c, g, st)) return r;
    }
    launch_pdl(text_attention_kernel, dim3(c->H, n), dim3(256), att_smem, st, static_cast<const bf16*>(c->QKV), c->AO, c->ctx,
               c->H, 0.125f);
    {
      GemmArgs g;
      g.a1 = opnd(c->AO, M, d, d); g.b1 = opnd(w.wo, d, d, d);
      g.M = M; g.N = d; g.epi = EPI_RESID_F32; g.bias = w.bo; g.out = c->XB; g.ldo = d; g.resid = c->XA; g.ldr = d;
      if (int r = tgemm(c, g, st)) return r;
    }
    launch_layernorm(c->XB, c->Hb, w.ln2g, w.ln2b, M, d, c->cfg.ln_eps, st);
    {
      GemmArgs g;
      g.a1 = opnd(c->Hb, M, d, d); g.b1 = opnd(w.w1, F, d, d);
      g.M = M; g.N = F; g.epi = EPI_GELU; g.bias = w.b1; g.out = c->Gb; g.ldo = F;
      if (int r = tgemm(c, g, st)) return r;
    }
    {
      GemmArgs g;
      g.a1 = opnd(c->Gb, M, F, F); g.b1 = opnd(w.w2, d, F, F);
      g.M = M; g.N = d; g.epi = EPI_RESID_F32; g.bias = w.b2; g.out = c->XA; g.ldo = d; g.resid = c->XB; g.ldr = d;
      if (int r = tgemm(c, g, st)) return r;
    }
  }
  launch_pdl(text_gather_eot_kernel, dim3(n), dim3(128), 0, st, static_cast<const int*>(c->tokens),
             static_cast<const float*>(c->XA), c->XG, c->ctx, d);
  launch_pool_project(c->XG, c->lnfg, c->lnfb, c->proj, c->pooled, c->feats, n, 1, d, c->P, c->cfg.ln_eps, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { c->err = std::string("text tower: ") + cudaGetErrorString(e); return TTL_E_CUDA; }
  return TTL_OK;
}

}  // namespace

namespace ttl {
void launch_text_embed(const int* tokens, const float* tok_emb, const float* pos_emb, float* x, int rows, int ctx, int d, int vocab,
                       cudaStream_t st) {
  const size_t total = static_cast<size_t>(rows) * (d / 4);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(text_embed_kernel, dim3(blocks), dim3(256), 0, st, tokens, tok_emb, pos_emb, x, rows, ctx, d / 4, vocab);
}
}  // namespace ttl

extern "C" {

const char* ttl_text_last_error(const ttl_text_ctx* c) { return c ? c->err.c_str() : g_text_err.c_str(); }

int ttl_text_create(ttl_text_ctx** out, const ttl_text_config* cfg) {
  if (!out || !cfg) { g_text_err = "null argument"; return TTL_E_INVALID; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_text_err = "no CUDA device (no CPU fallback exists)"; return TTL_E_ARCH; }
  if (cfg->device < 0 || cfg->device >= ndev) { g_text_err = "bad device ordinal"; return TTL_E_INVALID; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10) { g_text_err = "device is not compute capability 10.x (sm_100a only; no fallback)"; return TTL_E_ARCH; }
  if (cfg->width % 128 != 0 || cfg->width > 1024 || cfg->width != cfg->heads * DH || cfg->mlp_dim % 64 != 0 || cfg->layers <= 0 ||
      cfg->context <= 0 || cfg->context > 128 || cfg->vocab <= 0 || cfg->proj_dim <= 0 || cfg->max_prompts <= 0) {
    g_text_err = "unsupported text geometry (width%128, head_dim 64, context <= 128)";
    return TTL_E_SHAPE;
  }
  cudaSetDevice(cfg->device);
  ttl_text_ctx* c = new ttl_text_ctx();
  c->cfg = *cfg;
  c->num_sms = prop.multiProcessorCount;
  c->d = cfg->width; c->F = cfg->mlp_dim; c->P = cfg->proj_dim; c->L = cfg->layers; c->H = cfg->heads;
  c->ctx = cfg->context; c->vocab = cfg->vocab; c->maxp = cfg->max_prompts;
  const int d = c->d, F = c->F;
  const size_t M = static_cast<size_t>(c->maxp) * c->ctx;
  int rc = TTL_OK;
#define A(p, n) if (rc == TTL_OK) rc = talloc(c, &(p), static_cast<size_t>(n))
  A(c->tok, static_cast<size_t>(c->vocab) * d); A(c->pos, static_cast<size_t>(c->ctx) * d); A(c->lnfg, d); A(c->lnfb, d);
  A(c->proj, static_cast<size_t>(c->P) * d);
  c->lw.resize(c->L);
  for (int l = 0; l < c->L && rc == TTL_OK; ++l) {
    TextLayer& w = c->lw[l];
    A(w.wqkv, 3 * d * d); A(w.wo, d * d); A(w.w1, F * d); A(w.w2, d * F);
    A(w.bqkv, 3 * d); A(w.bo, d); A(w.b1, F); A(w.b2, d); A(w.ln1g, d); A(w.ln1b, d); A(w.ln2g, d); A(w.ln2b, d);
  }
  A(c->tokens, M); A(c->XA, M * d); A(c->XB, M * d); A(c->Hb, M * d); A(c->QKV, M * 3 * d); A(c->AO, M * d); A(c->Gb, M * F);
  A(c->XG, static_cast<size_t>(c->maxp) * d); A(c->pooled, static_cast<size_t>(c->maxp) * d);
  A(c->feats, static_cast<size_t>(c->maxp) * c->P);
#undef A
  if (rc != TTL_OK) {
    g_text_err = c->err;
    ttl_text_destroy(c);
    return rc;
  }
  cudaDeviceSynchronize();
  *out = c;
  return TTL_OK;
}

void ttl_text_destroy(ttl_text_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : c->allocs) cudaFree(p);
  delete c;
}

int ttl_text_set_weight(ttl_text_ctx* c, int32_t layer, int32_t kind, const float* host, int64_t numel) {
  if (!c || !host) return TTL_E_INVALID;
  cudaSetDevice(c->cfg.device);
  const int64_t d = c->d, F = c->F;
  auto need = [&](int64_t n) { if (numel != n) { c->err = "ttl_text_set_weight: wrong numel"; return false; } return true; };
  switch (kind) {
    case TTL_TW_TOKEN_EMB: if (!need(c->vocab * d)) return TTL_E_SHAPE; return up_f32(c, c->tok, host, numel);
    case TTL_TW_POS_EMB: if (!need(c->ctx * d)) return TTL_E_SHAPE; return up_f32(c, c->pos, host, numel);
    case TTL_TW_FINAL_LN_G: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, c->lnfg, host, numel);
    case TTL_TW_FINAL_LN_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, c->lnfb, host, numel);
    case TTL_TW_TEXT_PROJ: if (!need(c->P * d)) return TTL_E_SHAPE; return up_f32(c, c->proj, host, numel);
    default: break;
  }
  if (layer < 0 || layer >= c->L) { c->err = "ttl_text_set_weight: bad layer / kind"; return TTL_E_INVALID; }
  TextLayer& w = c->lw[layer];
  switch (kind) {
    case TTL_W_LN1_G: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.ln1g, host, numel);
    case TTL_W_LN1_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.ln1b, host, numel);
    case TTL_W_LN2_G: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.ln2g, host, numel);
    case TTL_W_LN2_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.ln2b, host, numel);
    case TTL_W_Q_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.bqkv, host, numel);
    case TTL_W_K_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.bqkv + d, host, numel);
    case TTL_W_V_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.bqkv + 2 * d, host, numel);
    case TTL_W_O_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.bo, host, numel);
    case TTL_W_FC1_B: if (!need(F)) return TTL_E_SHAPE; return up_f32(c, w.b1, host, numel);
    case TTL_W_FC2_B: if (!need(d)) return TTL_E_SHAPE; return up_f32(c, w.b2, host, numel);
    case TTL_W_Q_W: if (!need(d * d)) return TTL_E_SHAPE; return up_bf16(c, w.wqkv, host, numel);
    case TTL_W_K_W: if (!need(d * d)) return TTL_E_SHAPE; return up_bf16(c, w.wqkv + d * d, host, numel);
    case TTL_W_V_W: if (!need(d * d)) return TTL_E_SHAPE; return up_bf16(c, w.wqkv + 2 * d * d, host, numel);
    case TTL_W_O_W: if (!need(d * d)) return TTL_E_SHAPE; return up_bf16(c, w.wo, host, numel);
    case TTL_W_FC1_W: if (!need(F * d)) return TTL_E_SHAPE; return up_bf16(c, w.w1, host, numel);
    case TTL_W_FC2_W: if (!need(d * F)) return TTL_E_SHAPE; return up_bf16(c, w.w2, host, numel);
    default: c->err = "ttl_text_set_weight: unknown kind"; return TTL_E_INVALID;
  }
}

int ttl_text_encode(ttl_text_ctx* c, const int32_t* tokens_host, int32_t n_prompts, float* feats_host, void* stream) {
  if (!c || !tokens_host || !feats_host) return TTL_E_INVALID;
  if (n_prompts <= 0) { c->err = "n_prompts must be positive"; return TTL_E_SHAPE; }
  cudaSetDevice(c->cfg.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int p0 = 0; p0 < n_prompts; p0 += c->maxp) {
    const int n = n_prompts - p0 < c->maxp ? n_prompts - p0 : c->maxp;
    TCK(cudaMemcpyAsync(c->tokens, tokens_host + static_cast<size_t>(p0) * c->ctx, sizeof(int) * n * c->ctx,
                        cudaMemcpyHostToDevice, st));
    if (int r = encode_chunk(c, n, st)) return r;
    TCK(cudaMemcpyAsync(feats_host + static_cast<size_t>(p0) * c->P, c->feats, sizeof(float) * n * c->P,
                        cudaMemcpyDeviceToHost, st));
    TCK(cudaStreamSynchronize(st));
  }
  for (int p = 0; p < n_prompts; ++p) {   // text_features / text_features.norm(dim=-1, keepdim=True)  (custom_clip.py:662)
    float* f = feats_host + static_cast<size_t>(p) * c->P;
    double s = 0.0;
    for (int i = 0; i < c->P; ++i) s += static_cast<double>(f[i]) * f[i];
    const float inv = static_cast<float>(1.0 / std::sqrt(s > 0.0 ? s : 1.0));
    for (int i = 0; i < c->P; ++i) f[i] *= inv;
  }
  return TTL_OK;
}

}  // extern "C"
