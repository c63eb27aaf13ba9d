// Row-wise HBM-bound kernels: im2col for the patch embedding, CLS/pos + pre-LN, LayerNorm forward/backward,
// view gather.  One warp per row, 128-bit loads/stores, warp-shuffle reductions, fp32 statistics.
// Replaces ATen layer_norm / elementwise launches behind HF CLIPVisionEmbeddings, CLIPEncoderLayer
// (layer_norm1/2, pre_layrnorm) on the path clip/custom_clip.py:69-71 -> CLIPModel.get_image_features.
#include "kernels.cuh"
#include "ptx.cuh"

namespace ttl {

namespace {

constexpr int MAXV = 8;            // d <= 1024, d % 128 == 0
constexpr int ROWS_PER_BLOCK = 8;  // 8 warps per CTA

// One thread per 8 consecutive output columns = 8 consecutive pixels of one patch row (p % 8 == 0) or a generic
// 2-pixel path (p = 14).  Reads 32 B, writes 16 B; padding columns (K..Kp) are zero.
template <int VEC>
__global__ void im2col_kernel(const float* __restrict__ img, bf16* __restrict__ out, int V, int S, int p, int Kp) {
  pdl_wait();
  pdl_trigger();
  const int gp = S / p, T = gp * gp, K = 3 * p * p;
  const int cols = Kp / VEC;
  const size_t total = static_cast<size_t>(V) * T * cols;
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(e % cols) * VEC;
    const size_t row = e / cols;
    float v[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    if (col < K) {
      const int view = static_cast<int>(row / T), patch = static_cast<int>(row % T);
      const int py = patch / gp, px = patch % gp;
      const int c = col / (p * p), rem = col % (p * p), i = rem / p, j = rem % p;
      const float* src = img + ((static_cast<size_t>(view) * 3 + c) * S + (py * p + i)) * S + px * p + j;
      if (VEC == 8) {
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(src);
        v[0] = a.x; v[1] = a.y;
      }
    }
    if (VEC == 8) {
      uint4 o;
      o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(out + row * Kp + col) = o;
    } else {
      *reinterpret_cast<__nv_bfloat162*>(out + row * Kp + col) = __floats2bfloat162_rn(v[0], v[1]);
    }
  }
}

__device__ __forceinline__ void row_stats(const float4 (&v)[MAXV], int nv, int d, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) s += v[i].x + v[i].y + v[i].z + v[i].w;
  mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += a * a + b * b + c * c + e * e;
    }
  rstd = rsqrtf(warp_sum(q) / d + eps);
}

__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
embed_preln_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                   const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int tokens, int d,
                   float eps) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nv = d / 128;
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * d);
  const bool is_cls = (row % tokens) == 0;
  float4 v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int c4 = i * 32 + lane;
      if (is_cls) {
        float4 a = __ldg(reinterpret_cast<const float4*>(cls) + c4), b = __ldg(reinterpret_cast<const float4*>(pos) + c4);
        v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      } else {
        v[i] = xr[c4];
      }
    }
  float mean, rstd;
  row_stats(v, nv, d, eps, mean, rstd);
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int c4 = i * 32 + lane;
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      xr[c4] = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                           (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    }
}

__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
layernorm_kernel(const float* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int rows, int d, float eps, int rev) {
  pdl_wait();
  pdl_trigger();
  const int blk = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int row = blk * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nv = d / 128;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * d);
  float4 v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) v[i] = xr[i * 32 + lane];
  float mean, rstd;
  row_stats(v, nv, d, eps, mean, rstd);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * d);
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int c4 = i * 32 + lane;
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      uint2 o;
      o.x = pack_bf16((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      o.y = pack_bf16((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      yr[c4] = o;
    }
}

// dx = dres + rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat))
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ dres, float* __restrict__ dx, bf16* __restrict__ dxb, int rows, int d,
                     float eps) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nv = d / 128;
  const size_t base = static_cast<size_t>(row) * d;
  const float4* xr = reinterpret_cast<const float4*>(x + base);
  const float4* dyr = reinterpret_cast<const float4*>(dy + base);
  // every load of the row (x, dy, residual gradient) is issued before the first reduction: three dependent load phases cost
  // 461 us at 113 k rows (2.6 TB/s), half the HBM roofline
  float4 v[MAXV], gy[MAXV], rs[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int c4 = i * 32 + lane;
      v[i] = xr[c4];
      gy[i] = dyr[c4];
      rs[i] = dres != nullptr ? reinterpret_cast<const float4*>(dres + base)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  float mean, rstd;
  row_stats(v, nv, d, eps, mean, rstd);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int c4 = i * 32 + lane;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), t = gy[i];
      gy[i] = make_float4(g.x * t.x, g.y * t.y, g.z * t.z, g.w * t.w);
      v[i] = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
      s1 += gy[i].x + gy[i].y + gy[i].z + gy[i].w;
      s2 += gy[i].x * v[i].x + gy[i].y * v[i].y + gy[i].z * v[i].z + gy[i].w * v[i].w;
    }
  s1 = warp_sum(s1) / d;
  s2 = warp_sum(s2) / d;
  float4* dxr = reinterpret_cast<float4*>(dx + base);
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const int c4 = i * 32 + lane;
      float4 o = make_float4(rstd * (gy[i].x - s1 - v[i].x * s2), rstd * (gy[i].y - s1 - v[i].y * s2),
                             rstd * (gy[i].z - s1 - v[i].z * s2), rstd * (gy[i].w - s1 - v[i].w * s2));
      o.x += rs[i].x; o.y += rs[i].y; o.z += rs[i].z; o.w += rs[i].w;
      dxr[c4] = o;
      if (dxb != nullptr) {
        uint2 u;
        u.x = pack_bf16(o.x, o.y);
        u.y = pack_bf16(o.z, o.w);
        reinterpret_cast<uint2*>(dxb + base)[c4] = u;
      }
    }
}

__global__ void gather_views_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                    const int* __restrict__ idx, int n_sel, int view_stride, int per_view4, int sample_views,
                                    int k_per_sample) {
  pdl_wait();
  pdl_trigger();
  const int g = blockIdx.y;
  const int view = idx != nullptr ? (g / k_per_sample) * sample_views + idx[g] : g * view_stride;
  const float4* s = src + static_cast<size_t>(view) * per_view4;
  float4* o = dst + static_cast<size_t>(g) * per_view4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_view4; i += gridDim.x * blockDim.x) o[i] = s[i];
}

__global__ void block_mask_kernel(bf16* __restrict__ x, int M, int ncols, int rows_per_sample) {
  pdl_wait();
  pdl_trigger();
  const int chunks = ncols / 8;   // 16-byte chunks per row
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < static_cast<size_t>(M) * chunks;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / chunks), col = static_cast<int>(i % chunks) * 8;
    if (col / 64 != row / rows_per_sample) *reinterpret_cast<uint4*>(x + static_cast<size_t>(row) * ncols + col) = make_uint4(0, 0, 0, 0);
  }
}

}  // namespace

void launch_im2col(const float* images, bf16* patches, int V, int S, int p, cudaStream_t st) {
  const int K = 3 * p * p, Kp = (K + 63) / 64 * 64;
  const bool vec8 = (p % 8 == 0) && (S % 4 == 0);
  const size_t total = static_cast<size_t>(V) * (S / p) * (S / p) * (Kp / (vec8 ? 8 : 2));
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (vec8) im2col_kernel<8><<<blocks, 256, 0, st>>>(images, patches, V, S, p, Kp);
  else im2col_kernel<2><<<blocks, 256, 0, st>>>(images, patches, V, S, p, Kp);
}

void launch_embed_preln(float* x, const float* cls, const float* pos, const float* gamma, const float* beta, int V,
                        int tokens, int d, float eps, cudaStream_t st) {
  const int rows = V * tokens;
  launch_pdl(embed_preln_kernel, dim3((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), dim3(ROWS_PER_BLOCK * 32), 0, st, 
      x, cls, pos, gamma, beta, rows, tokens, d, eps);
}

void launch_layernorm(const float* x, bf16* y, const float* gamma, const float* beta, int rows, int d, float eps,
                      cudaStream_t st, int descending) {
  launch_pdl(layernorm_kernel, dim3((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), dim3(ROWS_PER_BLOCK * 32), 0, st, x, y, gamma, beta,
                                                                                                rows, d, eps, descending);
}

void launch_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* dres, float* dx,
                          bf16* dx_bf16, int rows, int d, float eps, cudaStream_t st) {
  launch_pdl(layernorm_bwd_kernel, dim3((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), dim3(ROWS_PER_BLOCK * 32), 0, st, 
      dy, x, gamma, dres, dx, dx_bf16, rows, d, eps);
}

void launch_gather_views(const float* src, float* dst, const int* view_idx, int n_sel, int view_stride, int tokens, int d,
                         cudaStream_t st, int sample_views, int k_per_sample) {
  const int per_view4 = tokens * d / 4;
  dim3 grid((per_view4 + 255) / 256 < 32 ? (per_view4 + 255) / 256 : 32, n_sel);
  launch_pdl(gather_views_kernel, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst),
                                            view_idx, n_sel, view_stride, per_view4, sample_views, k_per_sample);
}

void launch_block_mask(bf16* x, int M, int ncols, int rows_per_sample, cudaStream_t st) {
  const size_t total = static_cast<size_t>(M) * (ncols / 8);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(block_mask_kernel, dim3(blocks), dim3(256), 0, st, x, M, ncols, rows_per_sample);
}

}  // namespace ttl
