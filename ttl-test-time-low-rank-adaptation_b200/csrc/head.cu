// The small fp32 "head" of the TTL loop: CLS pooling + post-LN + visual projection, cosine logits against the
// cached text features, per-view entropy, top-K confidence selection, the marginal-entropy (ttl.py:56-61) and
// weighted-entropy (deyo.py:97-181) losses with their closed-form gradients, and the backward of the head down
// to the CLS rows of the last hidden state.  Everything here is latency-bound (<= 64 x 1000 floats).
#include "kernels.cuh"
#include "ptx.cuh"

#include <cfloat>

namespace ttl {

namespace {

constexpr int HT = 256;  // threads per CTA for head kernels

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

// pooled[v,:] = LN(x[v*tokens + 0, :])   (HF post_layernorm on the CLS token); one CTA per view
__global__ void __launch_bounds__(HT)
cls_ln_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              float* __restrict__ pooled, int tokens, int d, float eps, const int* __restrict__ pool_row) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int v = blockIdx.x;
  // pool_row (text tower): the pooled token of sequence v is row pool_row[v] (the EOT position) instead of row 0 (CLS)
  const float* xr = x + (static_cast<size_t>(v) * tokens + (pool_row != nullptr ? pool_row[v] : 0)) * d;
  float s = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) s += xr[i];
  const float mean = block_sum(s, red) / d;
  float q = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) { const float t = xr[i] - mean; q += t * t; }
  const float rstd = rsqrtf(block_sum(q, red) / d + eps);
  for (int i = threadIdx.x; i < d; i += blockDim.x)
    pooled[static_cast<size_t>(v) * d + i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
}

// out[m, n] = sum_k A[m, k] * B[n, k]      fp32.  A CTA stages 16 rows of A in smem; each of its 8 warps owns one row n of
// B at a time (contiguous, read with float4 across the lanes) and keeps 16 running dot products, reduced with shuffles.
// Grid (ceil(N/32), ceil(M/16)): >= 1000 CTAs for the [192 x 1000 x 512] logits GEMM, 32 for the 3-view prediction.
constexpr int SG_TM = 16, SG_NPC = 32;   // rows of A per CTA, columns (rows of B) per CTA
__global__ void __launch_bounds__(256)
small_gemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int M, int N,
                     int K) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sA[];            // [SG_TM][K]
  const int m0 = blockIdx.y * SG_TM, n0 = blockIdx.x * SG_NPC;
  const int K4 = K >> 2;
  for (int i = threadIdx.x; i < SG_TM * K4; i += 256) {
    const int r = i / K4, c = i - r * K4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < M) v = reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + r) * K)[c];
    reinterpret_cast<float4*>(sA)[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int nn = warp; nn < SG_NPC; nn += 8) {
    const int n = n0 + nn;
    if (n >= N) break;
    const float4* b4 = reinterpret_cast<const float4*>(B + static_cast<size_t>(n) * K);
    float acc[SG_TM];
#pragma unroll
    for (int i = 0; i < SG_TM; ++i) acc[i] = 0.f;
    for (int c = lane; c < K4; c += 32) {
      const float4 b = __ldg(b4 + c);
#pragma unroll
      for (int i = 0; i < SG_TM; ++i) {
        const float4 a = reinterpret_cast<const float4*>(sA)[i * K4 + c];
        acc[i] += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      }
    }
#pragma unroll
    for (int i = 0; i < SG_TM; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0 && m0 + i < M) out[static_cast<size_t>(m0 + i) * N + n] = v;
    }
  }
}

// out[m, n] = alpha * sum_k A[m, k] * B[k, n]     fp32; 8 rows of A per CTA, 64 columns, K split over 4 thread groups
__global__ void __launch_bounds__(256)
small_gemm_nn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int M, int N,
                     int K, float alpha) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh[];
  float* As = sh;                    // [8][K]
  float* part = sh + 8 * K;          // [4][8][64]
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < 8 * K; i += 256) {
    const int r = i / K, c = i - r * K;
    As[i] = (m0 + r < M) ? A[static_cast<size_t>(m0 + r) * K + c] : 0.f;
  }
  __syncthreads();
  const int tn = threadIdx.x & 63, ks = threadIdx.x >> 6;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (n0 + tn < N) {
#pragma unroll 4
    for (int k = ks; k < K; k += 4) {
      const float b = B[static_cast<size_t>(k) * N + n0 + tn];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += As[i * K + k] * b;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[(ks * 8 + i) * 64 + tn] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 64; i += 256) {
    const int r = i >> 6, c = i & 63;
    if (m0 + r < M && n0 + c < N)
      out[static_cast<size_t>(m0 + r) * N + n0 + c] =
          alpha * (part[(0 * 8 + r) * 64 + c] + part[(1 * 8 + r) * 64 + c] + part[(2 * 8 + r) * 64 + c] + part[(3 * 8 + r) * 64 + c]);
  }
}

// logits[v, :] = raw[v, :] * scale / |feats[v]| (in place) ; entropy[v] = H(softmax(logits[v]));  one CTA per view
__global__ void __launch_bounds__(HT)
scale_entropy_kernel(const float* __restrict__ feats, float scale, float* __restrict__ logits,
                     float* __restrict__ entropy, int C, int P) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int v = blockIdx.x;
  float mul = 1.f;   // feats == nullptr: logits are final, only the entropy is wanted
  if (feats != nullptr) {
    float s = 0.f;
    for (int i = threadIdx.x; i < P; i += blockDim.x) { const float t = feats[static_cast<size_t>(v) * P + i]; s += t * t; }
    mul = rsqrtf(block_sum(s, red)) * scale;
  }
  float* lg = logits + static_cast<size_t>(v) * C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { const float t = lg[c] * mul; lg[c] = t; mx = fmaxf(mx, t); }
  mx = block_max(mx, red);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(lg[c] - mx);
  const float lse = mx + logf(block_sum(se, red));
  float h = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { const float lp = lg[c] - lse; h -= expf(lp) * lp; }
  h = block_sum(h, red);
  if (threadIdx.x == 0) entropy[v] = h;
}

// rank-by-counting stable argsort prefix: idx[rank] = v for rank < K
__global__ void select_kernel(const float* __restrict__ entropy, int V, int K, const int* __restrict__ forced,
                              int* __restrict__ idx) {
  pdl_wait();
  pdl_trigger();
  entropy += static_cast<size_t>(blockIdx.x) * V;      // one CTA per test sample
  idx += static_cast<size_t>(blockIdx.x) * K;
  if (forced != nullptr) forced += static_cast<size_t>(blockIdx.x) * K;
  if (forced != nullptr) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) idx[k] = forced[k];
    return;
  }
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float hv = entropy[v];
    int rank = 0;
    for (int u = 0; u < V; ++u) {
      const float hu = entropy[u];
      rank += (hu < hv) || (hu == hv && u < v);
    }
    if (rank < K) idx[rank] = v;
  }
}

// single CTA.  a_c = logsumexp_k(lp[k,c]) - ln K ; L = -sum_c a_c e^{a_c} ; dL/dx[k,c] = -(1/K) p[k,c] (a_c - sum_j p[k,j] a_j)
__global__ void __launch_bounds__(1024)
tpt_loss_kernel(const float* __restrict__ logits, const int* __restrict__ idx, int K, int C, float* __restrict__ loss,
                float* __restrict__ dlogits, size_t logits_sstride) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh[];
  float* lse = sh;            // [K]
  float* sk = sh + K;         // [K]  sum_j p[k,j] a_j
  float* red = sh + 2 * K;    // [32]
  float* a = red + 32;        // [C]
  logits += static_cast<size_t>(blockIdx.x) * logits_sstride;   // one CTA per test sample
  if (idx != nullptr) idx += static_cast<size_t>(blockIdx.x) * K;
  loss += blockIdx.x;
  dlogits += static_cast<size_t>(blockIdx.x) * K * C;
  for (int k = 0; k < K; ++k) {
    const float* x = logits + static_cast<size_t>(idx ? idx[k] : k) * C;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, x[c]);
    mx = block_max(mx, red);
    float se = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(x[c] - mx);
    se = block_sum(se, red);
    if (threadIdx.x == 0) lse[k] = mx + logf(se);
  }
  __syncthreads();
  const float lnK = logf(static_cast<float>(K));
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, logits[static_cast<size_t>(idx ? idx[k] : k) * C + c] - lse[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(logits[static_cast<size_t>(idx ? idx[k] : k) * C + c] - lse[k] - mx);
    float ac = mx + logf(se) - lnK;
    ac = fmaxf(ac, -FLT_MAX);
    a[c] = ac;
    part -= ac * expf(ac);
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) *loss = part;
  for (int k = 0; k < K; ++k) {
    const float* x = logits + static_cast<size_t>(idx ? idx[k] : k) * C;
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(x[c] - lse[k]) * a[c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) sk[k] = s;
  }
  __syncthreads();
  const float invK = 1.f / K;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    const float p = expf(logits[static_cast<size_t>(idx ? idx[k] : k) * C + c] - lse[k]);
    dlogits[i] = -invK * p * (a[c] - sk[k]);
  }
}

// one CTA per test sample (V <= 1024 views).  H_v; keep H_v <= ln 1000; w_v = exp(-(H_v - e0)); L = mean_kept(w H);
// dL/dx[v,c] = -(w_v / n) p (log p + H_v)
__global__ void __launch_bounds__(1024)
deyo_loss_kernel(const float* __restrict__ logits, int V, int C, float e0, float* __restrict__ loss,
                 float* __restrict__ dlogits) {
  pdl_wait();
  pdl_trigger();
  logits += static_cast<size_t>(blockIdx.x) * V * C;     // samples are sample-major: [S][V][C] logits, [S] losses
  dlogits += static_cast<size_t>(blockIdx.x) * V * C;
  loss += blockIdx.x;
  extern __shared__ float sh[];
  float* lse = sh;          // [V]
  float* H = sh + V;        // [V]
  float* w = sh + 2 * V;    // [V]
  __shared__ float s_loss;
  __shared__ int s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int v = warp; v < V; v += nw) {
    const float* x = logits + static_cast<size_t>(v) * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
    const float l = mx + logf(warp_sum(se));
    float h = 0.f;
    for (int c = lane; c < C; c += 32) { const float lp = x[c] - l; h -= expf(lp) * lp; }
    h = warp_sum(h);
    if (lane == 0) { lse[v] = l; H[v] = h; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float thr = logf(1000.f);   // hard-coded in the reference (deyo.py:107)
    int n = 0;
    float L = 0.f;
    for (int v = 0; v < V; ++v) {
      if (H[v] <= thr) { w[v] = expf(-(H[v] - e0)); L += w[v] * H[v]; ++n; } else w[v] = 0.f;
    }
    s_n = n;
    s_loss = n > 0 ? L / n : 0.f;
    *loss = s_loss;
  }
  __syncthreads();
  const float invn = s_n > 0 ? 1.f / s_n : 0.f;
  for (int i = threadIdx.x; i < V * C; i += blockDim.x) {
    const int v = i / C;
    const float lp = logits[i] - lse[v];
    dlogits[i] = -(w[v] * invn) * expf(lp) * (lp + H[v]);
  }
}

// d/d f from d/d fhat:  df = (dfh - fhat <fhat, dfh>) / |f|     (in place on dfh); one CTA per compact view
__global__ void __launch_bounds__(HT)
l2norm_bwd_kernel(const float* __restrict__ feats, float* __restrict__ dfh, int P) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int g = blockIdx.x;
  const float* f = feats + static_cast<size_t>(g) * P;
  float* dd = dfh + static_cast<size_t>(g) * P;
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) s += f[p] * f[p];
  const float inv = rsqrtf(block_sum(s, red));
  float dt = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) dt += f[p] * inv * dd[p];
  dt = block_sum(dt, red);
  for (int p = threadIdx.x; p < P; p += blockDim.x) dd[p] = (dd[p] - f[p] * inv * dt) * inv;
}

// y[r, :] = x[r, :] / |x[r, :]|
__global__ void __launch_bounds__(HT)
l2norm_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int P) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const float* xr = x + static_cast<size_t>(blockIdx.x) * P;
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) s += xr[p] * xr[p];
  const float inv = rsqrtf(block_sum(s, red));
  for (int p = threadIdx.x; p < P; p += blockDim.x) y[static_cast<size_t>(blockIdx.x) * P + p] = xr[p] * inv;
}

// `--lora_encoder text`: the class features carry the gradient.  d that[c, :] = scale * sum_k dlogits[k, c] * fhat[view(k), :],
// view(k) = idx[k] (compact rows of the selected views) or k.  One CTA per class.
__global__ void __launch_bounds__(HT)
text_dfeat_kernel(const float* __restrict__ dlogits, const int* __restrict__ idx, const float* __restrict__ fhat, float scale,
                  float* __restrict__ dth, int K, int C, int P) {
  pdl_wait();
  pdl_trigger();
  const int c = blockIdx.x;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k)
      acc += dlogits[static_cast<size_t>(k) * C + c] * fhat[static_cast<size_t>(idx != nullptr ? idx[k] : k) * P + p];
    dth[static_cast<size_t>(c) * P + p] = acc * scale;
  }
}

// LayerNorm backward on the CLS row of compact view g: dpool[g,:] -> dx[g*tokens + 0, :] (+ bf16 copy)
__global__ void __launch_bounds__(HT)
cls_ln_bwd_kernel(const float* __restrict__ dpool, const float* __restrict__ x, const float* __restrict__ gamma,
                  float* __restrict__ dx, bf16* __restrict__ dxb, int tokens, int d, float eps, const int* __restrict__ pool_row) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int g = blockIdx.x;
  const size_t row = static_cast<size_t>(g) * tokens + (pool_row != nullptr ? pool_row[g] : 0);
  const float* xr = x + row * d;
  const float* dp = dpool + static_cast<size_t>(g) * d;
  float sm = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) sm += xr[i];
  const float mean = block_sum(sm, red) / d;
  float q = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) { const float t = xr[i] - mean; q += t * t; }
  const float rstd = rsqrtf(block_sum(q, red) / d + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float gy = gamma[i] * dp[i], xh = (xr[i] - mean) * rstd;
    s1 += gy; s2 += gy * xh;
  }
  s1 = block_sum(s1, red) / d;
  s2 = block_sum(s2, red) / d;
  float* dxr = dx + row * d;
  bf16* dbr = dxb + row * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float gy = gamma[i] * dp[i], xh = (xr[i] - mean) * rstd;
    const float o = rstd * (gy - s1 - xh * s2);
    dxr[i] = o;
    dbr[i] = __float2bfloat16(o);
  }
}

}  // namespace

static void small_gemm_nt_smem(size_t bytes) {
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (configured == 0) configured = 48 * 1024;
  if (bytes > configured) {
    cudaFuncSetAttribute(small_gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    configured = bytes;
  }
}

void launch_pool_project(const float* x, const float* gamma, const float* beta, const float* Wp, float* pooled,
                         float* feats, int V, int tokens, int d, int P, float eps, cudaStream_t st, const int* pool_row) {
  small_gemm_nt_smem(SG_TM * d * sizeof(float));
  launch_pdl(cls_ln_kernel, dim3(V), dim3(HT), 0, st, x, gamma, beta, pooled, tokens, d, eps, pool_row);
  launch_pdl(small_gemm_nt_kernel, dim3(dim3((P + SG_NPC - 1) / SG_NPC, (V + SG_TM - 1) / SG_TM)), dim3(256), SG_TM * d * sizeof(float), st, pooled, Wp, feats, V, P, d);
}
void launch_logits_entropy(const float* feats, const float* text, float scale, float* logits, float* entropy, int V,
                           int C, int P, cudaStream_t st) {
  small_gemm_nt_smem(SG_TM * P * sizeof(float));
  launch_pdl(small_gemm_nt_kernel, dim3(dim3((C + SG_NPC - 1) / SG_NPC, (V + SG_TM - 1) / SG_TM)), dim3(256), SG_TM * P * sizeof(float), st, feats, text, logits, V, C, P);
  launch_pdl(scale_entropy_kernel, dim3(V), dim3(HT), 0, st, feats, scale, logits, entropy, C, P);
}
void launch_entropy(float* logits, float* entropy, int V, int C, cudaStream_t st) {
  launch_pdl(scale_entropy_kernel, dim3(V), dim3(HT), 0, st, nullptr, 1.f, logits, entropy, C, 0);
}
void launch_select(const float* entropy, int V, int K, const int* forced_idx, int* idx, cudaStream_t st, int n_samples) {
  launch_pdl(select_kernel, dim3(n_samples), dim3(256), 0, st, entropy, V, K, forced_idx, idx);
}
void launch_tpt_loss(const float* logits, const int* idx, int K, int C, float* loss, float* dlogits, cudaStream_t st,
                     int n_samples, size_t logits_sstride) {
  launch_pdl(tpt_loss_kernel, dim3(n_samples), dim3(1024), (2 * K + 32 + C) * sizeof(float), st, logits, idx, K, C, loss, dlogits,
             logits_sstride);
}
void launch_deyo_loss(const float* logits, int V, int C, float margin_e0, float* loss, float* dlogits, cudaStream_t st,
                      int n_samples) {
  launch_pdl(deyo_loss_kernel, dim3(n_samples), dim3(1024), 3 * V * sizeof(float), st, logits, V, C, margin_e0, loss, dlogits);
}
void launch_head_bwd(const float* dlogits, const float* text, float scale, const float* feats, const float* Wp,
                     const float* x, const float* gamma, float* dfh, float* dpool, float* dx, bf16* dx_bf16, int G, int C,
                     int P, int tokens, int d, float eps, cudaStream_t st) {
  cudaMemsetAsync(dx, 0, static_cast<size_t>(G) * tokens * d * sizeof(float), st);
  cudaMemsetAsync(dx_bf16, 0, static_cast<size_t>(G) * tokens * d * sizeof(bf16), st);
  // d fhat = scale * dlogits @ T ; d f ; d pooled = d f @ Wp ; LN backward on the CLS rows
  launch_pdl(small_gemm_nn_kernel, dim3(dim3((P + 63) / 64, (G + 7) / 8)), dim3(256), (8 * C + 4 * 8 * 64) * sizeof(float), st, 
      dlogits, text, dfh, G, P, C, scale);
  launch_pdl(l2norm_bwd_kernel, dim3(G), dim3(HT), 0, st, feats, dfh, P);
  launch_pdl(small_gemm_nn_kernel, dim3(dim3((d + 63) / 64, (G + 7) / 8)), dim3(256), (8 * P + 4 * 8 * 64) * sizeof(float), st, 
      dfh, Wp, dpool, G, d, P, 1.0f);
  launch_pdl(cls_ln_bwd_kernel, dim3(G), dim3(HT), 0, st, dpool, x, gamma, dx, dx_bf16, tokens, d, eps, static_cast<const int*>(nullptr));
}
void launch_l2norm_rows(const float* x, float* y, int rows, int P, cudaStream_t st) {
  launch_pdl(l2norm_rows_kernel, dim3(rows), dim3(HT), 0, st, x, y, P);
}
void launch_text_head_bwd(const float* dlogits, const int* idx, int K, const float* fhat, float scale, const float* tfeats,
                          const float* Wp, const float* x, const float* gamma, const int* eot, float* dfh, float* dpool, float* dx,
                          bf16* dx_bf16, int C, int P, int tokens, int d, float eps, cudaStream_t st) {
  cudaMemsetAsync(dx, 0, static_cast<size_t>(C) * tokens * d * sizeof(float), st);
  cudaMemsetAsync(dx_bf16, 0, static_cast<size_t>(C) * tokens * d * sizeof(bf16), st);
  // d that = scale * dlogits^T fhat ; d t (L2-norm backward) ; d pooled = d t @ text_projection ; final-LN backward on the EOT rows
  launch_pdl(text_dfeat_kernel, dim3(C), dim3(HT), 0, st, dlogits, idx, fhat, scale, dfh, K, C, P);
  launch_pdl(l2norm_bwd_kernel, dim3(C), dim3(HT), 0, st, tfeats, dfh, P);
  launch_pdl(small_gemm_nn_kernel, dim3(dim3((d + 63) / 64, (C + 7) / 8)), dim3(256), (8 * P + 4 * 8 * 64) * sizeof(float), st,
      static_cast<const float*>(dfh), Wp, dpool, C, d, P, 1.0f);
  launch_pdl(cls_ln_bwd_kernel, dim3(C), dim3(HT), 0, st, static_cast<const float*>(dpool), x, gamma, dx, dx_bf16, tokens, d, eps, eot);
}

}  // namespace ttl
