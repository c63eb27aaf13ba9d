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

// one CTA per view: pooled = LN(x_cls); feats = Wp @ pooled  (one warp per output feature, coalesced over d)
__global__ void __launch_bounds__(HT)
pool_project_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ Wp, float* __restrict__ feats, int tokens, int d, int P, float eps) {
  extern __shared__ float sh[];
  float* pooled = sh;       // [d]
  float* red = sh + d;      // [32]
  const int v = blockIdx.x;
  const float* xr = x + static_cast<size_t>(v) * tokens * d;
  float s = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) s += xr[i];
  const float mean = block_sum(s, red) / d;
  float q = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) { const float t = xr[i] - mean; q += t * t; }
  const float rstd = rsqrtf(block_sum(q, red) / d + eps);
  for (int i = threadIdx.x; i < d; i += blockDim.x) pooled[i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = warp; p < P; p += nw) {
    const float* w = Wp + static_cast<size_t>(p) * d;
    float a = 0.f;
    for (int i = lane; i < d; i += 32) a += w[i] * pooled[i];
    a = warp_sum(a);
    if (lane == 0) feats[static_cast<size_t>(v) * P + p] = a;
  }
}

// one CTA per view: logits over all classes + entropy
__global__ void __launch_bounds__(HT)
logits_entropy_kernel(const float* __restrict__ feats, const float* __restrict__ text, float scale,
                      float* __restrict__ logits, float* __restrict__ entropy, int C, int P) {
  extern __shared__ float sh[];
  float* f = sh;        // [P] normalised feature * scale
  float* red = sh + P;  // [32]
  const int v = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) { const float t = feats[static_cast<size_t>(v) * P + i]; s += t * t; }
  const float inv = rsqrtf(block_sum(s, red));
  for (int i = threadIdx.x; i < P; i += blockDim.x) f[i] = feats[static_cast<size_t>(v) * P + i] * inv * scale;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* lg = logits + static_cast<size_t>(v) * C;
  for (int c = warp; c < C; c += nw) {
    const float* t = text + static_cast<size_t>(c) * P;
    float a = 0.f;
    for (int i = lane; i < P; i += 32) a += t[i] * f[i];
    a = warp_sum(a);
    if (lane == 0) lg[c] = a;
  }
  __syncthreads();
  // entropy = -sum p log p  with log p = x - lse
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lg[c]);
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
                float* __restrict__ dlogits) {
  extern __shared__ float sh[];
  float* lse = sh;            // [K]
  float* sk = sh + K;         // [K]  sum_j p[k,j] a_j
  float* red = sh + 2 * K;    // [32]
  float* a = red + 32;        // [C]
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

// single CTA (V <= 1024 views).  H_v; keep H_v <= ln 1000; w_v = exp(-(H_v - e0)); L = mean_kept(w H);
// dL/dx[v,c] = -(w_v / n) p (log p + H_v)
__global__ void __launch_bounds__(1024)
deyo_loss_kernel(const float* __restrict__ logits, int V, int C, float e0, float* __restrict__ loss,
                 float* __restrict__ dlogits) {
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

// one CTA per compact view g: dlogits[g,:] -> dx[g*tokens + 0, :]
__global__ void __launch_bounds__(HT)
head_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ text, float scale,
                const float* __restrict__ feats, const float* __restrict__ Wp, const float* __restrict__ x,
                const float* __restrict__ gamma, float* __restrict__ dx, bf16* __restrict__ dxb, int C, int P,
                int tokens, int d, float eps) {
  extern __shared__ float sh[];
  float* dfh = sh;           // [P] d/d fhat, then d/d f
  float* dpool = sh + P;     // [d]
  float* red = dpool + d;    // [32]
  const int g = blockIdx.x;
  const float* dl = dlogits + static_cast<size_t>(g) * C;
  const float* f = feats + static_cast<size_t>(g) * P;
  // dfhat[p] = scale * sum_c dl[c] T[c,p]
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a += dl[c] * text[static_cast<size_t>(c) * P + p];
    dfh[p] = a * scale;
  }
  float s = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) s += f[p] * f[p];
  const float nrm2 = block_sum(s, red);
  const float inv = rsqrtf(nrm2);
  float dt = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) dt += f[p] * inv * dfh[p];
  dt = block_sum(dt, red);
  for (int p = threadIdx.x; p < P; p += blockDim.x) dfh[p] = (dfh[p] - f[p] * inv * dt) * inv;   // d/d f
  __syncthreads();
  // dpooled[k] = sum_p df[p] Wp[p,k]
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    float a = 0.f;
    for (int p = 0; p < P; ++p) a += dfh[p] * Wp[static_cast<size_t>(p) * d + k];
    dpool[k] = a;
  }
  __syncthreads();
  // LayerNorm backward on the CLS row
  const float* xr = x + static_cast<size_t>(g) * tokens * d;
  float sm = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) sm += xr[i];
  const float mean = block_sum(sm, red) / d;
  float q = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) { const float t = xr[i] - mean; q += t * t; }
  const float rstd = rsqrtf(block_sum(q, red) / d + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float gy = gamma[i] * dpool[i], xh = (xr[i] - mean) * rstd;
    s1 += gy; s2 += gy * xh;
  }
  s1 = block_sum(s1, red) / d;
  s2 = block_sum(s2, red) / d;
  float* dxr = dx + static_cast<size_t>(g) * tokens * d;
  bf16* dbr = dxb + static_cast<size_t>(g) * tokens * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float gy = gamma[i] * dpool[i], xh = (xr[i] - mean) * rstd;
    const float o = rstd * (gy - s1 - xh * s2);
    dxr[i] = o;
    dbr[i] = __float2bfloat16(o);
  }
}

}  // namespace

void launch_pool_project(const float* x, const float* gamma, const float* beta, const float* Wp, float* feats, int V,
                         int tokens, int d, int P, float eps, cudaStream_t st) {
  pool_project_kernel<<<V, HT, (d + 32) * sizeof(float), st>>>(x, gamma, beta, Wp, feats, tokens, d, P, eps);
}
void launch_logits_entropy(const float* feats, const float* text, float scale, float* logits, float* entropy, int V,
                           int C, int P, cudaStream_t st) {
  logits_entropy_kernel<<<V, HT, (P + 32) * sizeof(float), st>>>(feats, text, scale, logits, entropy, C, P);
}
void launch_select(const float* entropy, int V, int K, const int* forced_idx, int* idx, cudaStream_t st) {
  select_kernel<<<1, 256, 0, st>>>(entropy, V, K, forced_idx, idx);
}
void launch_tpt_loss(const float* logits, const int* idx, int K, int C, float* loss, float* dlogits, cudaStream_t st) {
  tpt_loss_kernel<<<1, 1024, (2 * K + 32 + C) * sizeof(float), st>>>(logits, idx, K, C, loss, dlogits);
}
void launch_deyo_loss(const float* logits, int V, int C, float margin_e0, float* loss, float* dlogits, cudaStream_t st) {
  deyo_loss_kernel<<<1, 1024, 3 * V * sizeof(float), st>>>(logits, V, C, margin_e0, loss, dlogits);
}
void launch_head_bwd(const float* dlogits, const float* text, float scale, const float* feats, const float* Wp,
                     const float* x, const float* gamma, float* dx, bf16* dx_bf16, int G, int C, int P, int tokens,
                     int d, float eps, cudaStream_t st) {
  cudaMemsetAsync(dx, 0, static_cast<size_t>(G) * tokens * d * sizeof(float), st);
  cudaMemsetAsync(dx_bf16, 0, static_cast<size_t>(G) * tokens * d * sizeof(bf16), st);
  head_bwd_kernel<<<G, HT, (P + d + 32) * sizeof(float), st>>>(dlogits, text, scale, feats, Wp, x, gamma, dx, dx_bf16,
                                                               C, P, tokens, d, eps);
}

}  // namespace ttl
