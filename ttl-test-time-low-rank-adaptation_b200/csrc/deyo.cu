// Optional branches of the reference's weighted-entropy head (deyo.py:93-196; flags ttl.py:410-424), on the device:
//   filter_ent   keep the int(V * selection_p) lowest-entropy views (deyo.py:103-105; the selection kernel of head.cu)
//   filter_plpd  x' = the kept views with their object structure destroyed (deyo.py:115-136: occlusion window / shuffled
//                patch tiles between two antialiased resizes / one pixel permutation), a second forward on x', and
//                PLPD = p(x)[argmax p(x)] - p(x')[argmax p(x)] > plpd_threshold (deyo.py:137-148)
//   reweight_*   the detached coefficient exp(-(H - margin)) (deyo.py:159-179); off = plain mean entropy
// plus an AdamW step that a test sample skips when every one of its views was filtered out (deyo.py:184: the reference
// only steps when final_backward != 0) -- with several samples adapted concurrently that is a per-sample decision.
// The random draws (tile orders, pixel permutation) stay on the host so that the torch seed decides as in the reference.
#include "kernels.cuh"
#include "ptx.cuh"

#include <cmath>

namespace ttl {

namespace {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// view index (within the whole call) of kept entry b of sample s: idx == nullptr -> the b-th view of the sample
__device__ __forceinline__ int kept_view(const int* idx, int s, int b, int V, int n1) {
  return s * V + (idx != nullptr ? idx[s * n1 + b] : b);
}

// aug_type == 'occ' (deyo.py:117-121): the occlusion window of every channel is filled with that channel's mean over the view.
// One CTA per (kept view, channel) plane.
__global__ void __launch_bounds__(256)
destroy_occ_kernel(const float* __restrict__ images, const int* __restrict__ idx, float* __restrict__ xprime, int V, int n1,
                   int size, int occ, int r0, int c0) {
  __shared__ float red[8];
  const int plane = blockIdx.x, g = plane / 3, ch = plane - g * 3;
  const int s = g / n1, b = g - s * n1;
  const int hw = size * size;
  const float* src = images + (static_cast<size_t>(kept_view(idx, s, b, V, n1)) * 3 + ch) * hw;
  float* dst = xprime + (static_cast<size_t>(g) * 3 + ch) * hw;
  float acc = 0.f;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) acc += src[i];
  const float mean = block_sum_256(acc, red) / static_cast<float>(hw);
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const int y = i / size, x = i - y * size;
    const bool in = y >= r0 && y < r0 + occ && x >= c0 && x < c0 + occ;
    dst[i] = in ? mean : src[i];
  }
}

// aug_type == 'pixel' (deyo.py:131-134): x'[b, c, p] = x[b, c, perm[p]], one permutation per forward shared by views and channels
__global__ void destroy_pixel_kernel(const float* __restrict__ images, const int* __restrict__ idx, const int* __restrict__ perm,
                                     float* __restrict__ xprime, int V, int n1, int hw, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const size_t plane = i / hw;
  const int g = static_cast<int>(plane / 3), ch = static_cast<int>(plane - static_cast<size_t>(g) * 3);
  const int s = g / n1, b = g - s * n1;
  xprime[i] = images[(static_cast<size_t>(kept_view(idx, s, b, V, n1)) * 3 + ch) * hw + perm[static_cast<size_t>(s) * hw + p]];
}

// One axis of torch's antialiased bilinear resize (aten upsample_bilinear2d_aa, which torchvision's Resize calls for tensors;
// align_corners = False): output index i of `out_size` reads input taps [x0, x0 + n) with normalised triangle weights.
struct Taps { int x0, n; float w[4]; };
__device__ __forceinline__ Taps aa_taps(int i, int in_size, int out_size) {
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const float support = scale >= 1.f ? scale : 1.f, invscale = scale >= 1.f ? 1.f / scale : 1.f;
  const float center = scale * (static_cast<float>(i) + 0.5f);
  Taps t;
  t.x0 = max(static_cast<int>(center - support + 0.5f), 0);
  t.n = min(min(static_cast<int>(center + support + 0.5f), in_size) - t.x0, 4);
  float total = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float x = fabsf((static_cast<float>(j + t.x0) - center + 0.5f) * invscale);
    t.w[j] = j < t.n && x < 1.f ? 1.f - x : 0.f;
    total += t.w[j];
  }
  const float inv = total != 0.f ? 1.f / total : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) t.w[j] *= inv;
  return t;
}

// aug_type == 'patch', first half (deyo.py:123-125): t = Resize((side, side))(x[kept]), side = (size / patch_len) * patch_len
__global__ void patch_down_kernel(const float* __restrict__ images, const int* __restrict__ idx, float* __restrict__ t, int V, int n1,
                                  int size, int side, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % side), y = static_cast<int>((i / side) % side);
  const size_t plane = i / (static_cast<size_t>(side) * side);
  const int g = static_cast<int>(plane / 3), ch = static_cast<int>(plane - static_cast<size_t>(g) * 3);
  const int s = g / n1, b = g - s * n1;
  const float* src = images + (static_cast<size_t>(kept_view(idx, s, b, V, n1)) * 3 + ch) * size * size;
  const Taps ty = aa_taps(y, size, side), tx = aa_taps(x, size, side);
  float acc = 0.f;
  for (int jy = 0; jy < ty.n; ++jy) {
    float row = 0.f;
    for (int jx = 0; jx < tx.n; ++jx) row += tx.w[jx] * src[(ty.x0 + jy) * size + tx.x0 + jx];
    acc += ty.w[jy] * row;
  }
  t[i] = acc;
}

// second half (deyo.py:126-130): the patch_len^2 tiles of every view are reordered (new tile p = old tile perm[b][p]), then
// x' = Resize((size, size))(shuffled).  perm: [n_kept_total][patch_len^2]
__global__ void patch_up_kernel(const float* __restrict__ t, const int* __restrict__ perm, float* __restrict__ xprime, int size, int side,
                                int pl, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % size), y = static_cast<int>((i / size) % size);
  const size_t plane = i / (static_cast<size_t>(size) * size);
  const int g = static_cast<int>(plane / 3);
  const float* src = t + plane * side * side;
  const int* pm = perm + static_cast<size_t>(g) * pl * pl;
  const int ph = side / pl;
  const Taps ty = aa_taps(y, side, size), tx = aa_taps(x, side, size);
  float acc = 0.f;
  for (int jy = 0; jy < ty.n; ++jy) {
    const int ys = ty.x0 + jy, tyi = ys / ph, yin = ys - tyi * ph;
    float row = 0.f;
    for (int jx = 0; jx < tx.n; ++jx) {
      const int xs = tx.x0 + jx, txi = xs / ph, xin = xs - txi * ph;
      const int q = pm[tyi * pl + txi];                       // source tile of destination tile (tyi, txi)
      row += tx.w[jx] * src[((q / pl) * ph + yin) * side + (q % pl) * ph + xin];
    }
    acc += ty.w[jy] * row;
  }
  xprime[i] = acc;
}

// PLPD of deyo.py:139-146, one warp per kept view: top = argmax softmax(x) (lowest index on ties, as torch.argmax),
// keep = softmax(x)[top] - softmax(x')[top] > threshold
__global__ void plpd_kernel(const float* __restrict__ logits, const float* __restrict__ logits_prime, int G, int C, float thr,
                            int* __restrict__ keep, float* __restrict__ plpd_out) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= G) return;
  const float* x = logits + static_cast<size_t>(g) * C;
  const float* xp = logits_prime + static_cast<size_t>(g) * C;
  float mx = -INFINITY, mxp = -INFINITY;
  int am = 0;
  for (int c = lane; c < C; c += 32) {
    if (x[c] > mx) { mx = x[c]; am = c; }
    mxp = fmaxf(mxp, xp[c]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  mxp = warp_max(mxp);
  float se = 0.f, sep = 0.f;
  for (int c = lane; c < C; c += 32) { se += expf(x[c] - mx); sep += expf(xp[c] - mxp); }
  se = warp_sum(se);
  sep = warp_sum(sep);
  if (lane == 0) {
    const float p = 1.f / se, pp = expf(xp[am] - mxp) / sep;      // exp(x[top] - max) = 1
    keep[g] = (p - pp) > thr ? 1 : 0;
    if (plpd_out != nullptr) plpd_out[g] = p - pp;
  }
}

// deyo.py:102-181 over the n1 candidate views of one sample (one CTA per sample): H_v; kept = keep[v] (nullable) and, without
// filter_ent, H_v <= ln 1000 (deyo.py:107); w_v = reweight ? reweight_ent * exp(-(H_v - e0)) : 1; L = mean_kept(w H);
// dL/dx[v, c] = -(w_v / n) p (log p + H_v) on kept rows, 0 elsewhere.  active[s] = n > 0, steps[s] += active[s].
__global__ void __launch_bounds__(1024)
deyo_general_loss_kernel(const float* __restrict__ logits, const int* __restrict__ keep, int n1, int C, float e0, int filter_ent,
                         int reweight, float reweight_ent, float* __restrict__ loss, float* __restrict__ dlogits,
                         int* __restrict__ active, int* __restrict__ steps, int* __restrict__ n_kept) {
  const int s = blockIdx.x;
  logits += static_cast<size_t>(s) * n1 * C;
  dlogits += static_cast<size_t>(s) * n1 * C;
  if (keep != nullptr) keep += static_cast<size_t>(s) * n1;
  extern __shared__ float sh[];
  float* lse = sh;           // [n1]
  float* H = sh + n1;        // [n1]
  float* w = sh + 2 * n1;    // [n1]
  __shared__ int s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int v = warp; v < n1; v += nw) {
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
    const float thr = logf(1000.f);
    int n = 0;
    float L = 0.f;
    for (int v = 0; v < n1; ++v) {
      const bool k = (keep == nullptr || keep[v] != 0) && (filter_ent != 0 || H[v] <= thr);
      if (k) { w[v] = reweight ? reweight_ent * expf(-(H[v] - e0)) : 1.f; L += w[v] * H[v]; ++n; } else w[v] = 0.f;
    }
    s_n = n;
    loss[s] = n > 0 ? L / n : 0.f;
    active[s] = n > 0 ? 1 : 0;
    if (n > 0) steps[s] += 1;
    if (n_kept != nullptr) n_kept[s] = n;
  }
  __syncthreads();
  const float invn = s_n > 0 ? 1.f / s_n : 0.f;
  for (int i = threadIdx.x; i < n1 * C; i += blockDim.x) {
    const int v = i / C;
    const float lp = logits[i] - lse[v];
    dlogits[i] = -(w[v] * invn) * expf(lp) * (lp + H[v]);
  }
}

// torch.optim.AdamW on the factors of the samples that took a step (active[s] != 0), each with its own step count
__global__ void adamw_masked_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                    int n_per_sample, const int* __restrict__ active, const int* __restrict__ steps, float lr, float b1,
                                    float b2, float eps, float wd) {
  const int s = blockIdx.y;
  if (active[s] == 0) return;
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double bc1 = 1.0 - pow(static_cast<double>(b1), steps[s]), bc2 = 1.0 - pow(static_cast<double>(b2), steps[s]);
    s_step_size = static_cast<float>(lr / bc1);
    s_bc2_sqrt = static_cast<float>(sqrt(bc2));
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per_sample) return;
  const size_t k = static_cast<size_t>(s) * n_per_sample + i;
  const float gi = g[k];
  float pi = p[k] * (1.0f - lr * wd);
  const float mi = m[k] + (gi - m[k]) * (1.0f - b1);
  const float vi = v[k] * b2 + (1.0f - b2) * gi * gi;
  const float denom = sqrtf(vi) / s_bc2_sqrt + eps;
  pi -= s_step_size * (mi / denom);
  p[k] = pi; m[k] = mi; v[k] = vi;
}

}  // namespace

void launch_destroy_occ(const float* images, const int* idx, float* xprime, int S, int V, int n1, int size, int occ, int r0, int c0,
                        cudaStream_t st) {
  destroy_occ_kernel<<<S * n1 * 3, 256, 0, st>>>(images, idx, xprime, V, n1, size, occ, r0, c0);
}
void launch_destroy_pixel(const float* images, const int* idx, const int* perm, float* xprime, int S, int V, int n1, int size,
                          cudaStream_t st) {
  const size_t total = static_cast<size_t>(S) * n1 * 3 * size * size;
  destroy_pixel_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(images, idx, perm, xprime, V, n1, size * size, total);
}
void launch_destroy_patch(const float* images, const int* idx, const int* perm, float* scratch, float* xprime, int S, int V, int n1,
                          int size, int patch_len, cudaStream_t st) {
  const int side = (size / patch_len) * patch_len;
  const size_t t1 = static_cast<size_t>(S) * n1 * 3 * side * side, t2 = static_cast<size_t>(S) * n1 * 3 * size * size;
  patch_down_kernel<<<static_cast<unsigned>((t1 + 255) / 256), 256, 0, st>>>(images, idx, scratch, V, n1, size, side, t1);
  patch_up_kernel<<<static_cast<unsigned>((t2 + 255) / 256), 256, 0, st>>>(scratch, perm, xprime, size, side, patch_len, t2);
}
void launch_plpd(const float* logits, const float* logits_prime, int G, int C, float thr, int* keep, float* plpd_out, cudaStream_t st) {
  plpd_kernel<<<(G + 3) / 4, 128, 0, st>>>(logits, logits_prime, G, C, thr, keep, plpd_out);
}
void launch_deyo_general_loss(const float* logits, const int* keep, int n1, int C, float e0, int filter_ent, int reweight,
                              float reweight_ent, float* loss, float* dlogits, int* active, int* steps, int* n_kept, int S,
                              cudaStream_t st) {
  deyo_general_loss_kernel<<<S, 1024, 3 * n1 * sizeof(float), st>>>(logits, keep, n1, C, e0, filter_ent, reweight, reweight_ent, loss,
                                                                  dlogits, active, steps, n_kept);
}
void launch_adamw_masked(float* p, const float* g, float* m, float* v, int n_per_sample, int S, const int* active, const int* steps,
                         float lr, float b1, float b2, float eps, float wd, cudaStream_t st) {
  adamw_masked_kernel<<<dim3((n_per_sample + 255) / 256, S), 256, 0, st>>>(p, g, m, v, n_per_sample, active, steps, lr, b1, b2, eps, wd);
}

}  // namespace ttl
