// View generator on the GPU ("next" row N1): the reference builds the 64 views of every test image on the
// host with PIL + torchvision and ships 38.5 MB of fp32 pixels per sample to the GPU
// (data/datautils.py:98-157 AugMixAugmenter with its empty augmentation list, ttl.py:232-241, 324-336):
//     view 0      = Normalize(ToTensor(CenterCrop(S)(Resize(S, BICUBIC, antialias)(img))))
//     views 1..   = Normalize(ToTensor(RandomHorizontalFlip()(RandomResizedCrop(S)(img))))
// Here the host ships the decoded uint8 image (H x W x 3) and one 24-byte spec per view (the crop box torchvision's RNG
// drew + the flip coin); the resampling runs on the device and is BIT-EXACT with Pillow's 8-bit antialiased resampler
// (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc):
//   1. views_coeff_kernel: per (view, axis, output index) the filter window and its 22-bit fixed-point coefficients, computed
//      in IEEE double with explicitly rounded operations (no FMA contraction) exactly as the C code does;
//   2. views_hpass_kernel: horizontal pass over the rows of the source window -> uint8 (Pillow rounds between the passes);
//   3. views_vpass_kernel: vertical pass -> uint8 -> (x / 255 - mean) / std in fp32 (torchvision's operation order) ->
//      either fp32 views [n, 3, S, S] or, fused with the patch-embedding im2col, the bf16 patch matrix [n*T, Kp] the
//      tcgen05 GEMM consumes -- the fp32 views never exist in that mode.
// All integer / byte work, HBM/L2-bound and tiny next to the encoder (~1 MB of source pixels per sample).
#include "views.cuh"
#include "ptx.cuh"

#include <cmath>

namespace ttl {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Pillow Resample.c

// Pillow's bilinear_filter / bicubic_filter (a = -0.5), every operation individually rounded.
__device__ __forceinline__ double resample_filter(int bicubic, double x) {
  if (x < 0.0) x = -x;
  if (!bicubic) return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// One thread per (view, axis, output index): entry = [first source index, tap count, taps...] (stride 2 + ksize ints).
__global__ void views_coeff_kernel(const ViewDesc* __restrict__ desc, int* __restrict__ coef, int size) {
  pdl_wait();
  pdl_trigger();
  const ViewDesc d = desc[blockIdx.x >> 1];
  const int axis = blockIdx.x & 1;   // 0 = horizontal, 1 = vertical
  const int in_size = axis ? d.h : d.w, out_size = axis ? d.oh : d.ow, first = axis ? d.oy : d.ox;
  const int ks = axis ? d.ksv : d.ksh;
  int* base = coef + (axis ? d.coef_v_off : d.coef_h_off);
  const double support0 = d.bicubic ? 2.0 : 1.0;
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(support0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int i = threadIdx.x; i < size; i += blockDim.x) {
    const double center = __dmul_rn(__dadd_rn(static_cast<double>(first + i), 0.5), scale);
    int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x)
      ww = __dadd_rn(ww, resample_filter(d.bicubic, __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss)));
    int* e = base + static_cast<size_t>(i) * (2 + ks);
    e[0] = xmin;
    e[1] = n;
    for (int x = 0; x < ks; ++x) {
      int k = 0;
      if (x < n) {
        double w = resample_filter(d.bicubic, __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        const double f = __dmul_rn(w, static_cast<double>(1 << PRECISION_BITS));
        k = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, f)) : static_cast<int>(__dadd_rn(0.5, f));
      }
      e[2 + x] = k;
    }
  }
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[view][y][x][c] = clip8(sum_k img[y0 + y][x0 + xmin + k][c] * k_x[k]) for the rows of the source window.
__global__ void __launch_bounds__(256)
views_hpass_kernel(const uint8_t* __restrict__ img, const ViewDesc* __restrict__ desc, const int* __restrict__ coef,
                   uint8_t* __restrict__ tmp, int size) {
  pdl_wait();
  pdl_trigger();
  const ViewDesc d = desc[blockIdx.y];
  const int total = d.h * size;
  const int* cbase = coef + d.coef_h_off;
  const int stride = 2 + d.ksh;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / size, x = idx - y * size;
    const int* e = cbase + static_cast<size_t>(x) * stride;
    const int xmin = e[0], n = e[1];
    const uint8_t* src = img + d.img_off + (static_cast<size_t>(d.y0 + y) * d.W + d.x0 + xmin) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int k = 0; k < n; ++k) {
      const int w = e[2 + k];
      s0 += src[3 * k] * w;
      s1 += src[3 * k + 1] * w;
      s2 += src[3 * k + 2] * w;
    }
    uint8_t* o = tmp + d.tmp_off + static_cast<size_t>(idx) * 3;
    o[0] = clip8(s0);
    o[1] = clip8(s1);
    o[2] = clip8(s2);
  }
}

// Vertical pass + ToTensor + Normalize (+ flip).  views != nullptr: fp32 [n,3,S,S]; patches != nullptr: bf16 [n*T, Kp] in
// the patch-embedding GEMM's operand layout (column = c*p*p + i*p + j, HF CLIPVisionEmbeddings conv k=p s=p).
__global__ void __launch_bounds__(256)
views_vpass_kernel(const uint8_t* __restrict__ tmp, const ViewDesc* __restrict__ desc, const int* __restrict__ coef,
                   float* __restrict__ views, bf16* __restrict__ patches, int size, int p, int Kp, float m0, float m1,
                   float m2, float sd0, float sd1, float sd2) {
  pdl_wait();
  pdl_trigger();
  const int v = blockIdx.y;
  const ViewDesc d = desc[v];
  const int total = size * size;
  const int* cbase = coef + d.coef_v_off;
  const int stride = 2 + d.ksv;
  const float mean[3] = {m0, m1, m2}, sd[3] = {sd0, sd1, sd2};
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / size, x = idx - y * size;
    const int* e = cbase + static_cast<size_t>(y) * stride;
    const int ymin = e[0], n = e[1];
    const uint8_t* src = tmp + d.tmp_off + (static_cast<size_t>(ymin) * size + x) * 3;
    int s[3] = {1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1)};
    for (int k = 0; k < n; ++k) {
      const int w = e[2 + k];
      const uint8_t* q = src + static_cast<size_t>(k) * size * 3;
      s[0] += q[0] * w;
      s[1] += q[1] * w;
      s[2] += q[2] * w;
    }
    const int xo = d.flip ? size - 1 - x : x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // ToTensor: uint8 -> float / 255 ; Normalize: (t - mean) / std, each step rounded to fp32 like torch
      const float val = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(clip8(s[c])), 255.0f), mean[c]), sd[c]);
      if (views != nullptr) views[((static_cast<size_t>(v) * 3 + c) * size + y) * size + xo] = val;
      if (patches != nullptr) {
        const int gp = size / p;
        const size_t row = static_cast<size_t>(v) * gp * gp + (y / p) * gp + xo / p;
        patches[row * Kp + c * p * p + (y % p) * p + xo % p] = __float2bfloat16(val);
      }
    }
  }
}

// zero the K..Kp padding columns of the patch matrix (p = 14: K = 588, Kp = 640); a no-op for p = 16 (K = Kp = 768)
__global__ void views_pad_kernel(bf16* __restrict__ patches, int rows, int K, int Kp) {
  pdl_wait();
  pdl_trigger();
  const int pad = Kp - K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * pad; i += gridDim.x * blockDim.x)
    patches[static_cast<size_t>(i / pad) * Kp + K + i % pad] = __float2bfloat16(0.f);
}

// torchvision.transforms.functional._compute_resized_output_size for an int size: smaller edge -> size
void resized_size(int h, int w, int size, int* nh, int* nw) {
  const int shrt = w <= h ? w : h, lng = w <= h ? h : w;
  const int new_long = static_cast<int>(static_cast<double>(static_cast<long long>(size) * lng) / shrt);
  if (w <= h) { *nw = size; *nh = new_long; } else { *nh = size; *nw = new_long; }
}

int ksize_of(int in_size, int out_size, bool bicubic) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double fs = scale < 1.0 ? 1.0 : scale;
  return static_cast<int>(std::ceil((bicubic ? 2.0 : 1.0) * fs)) * 2 + 1;
}

}  // namespace

const char* views_plan(const ttl_view_spec* specs, int n_views, int H, int W, int size, long long img_off, ViewDesc* out,
                       size_t* coef_ints, size_t* tmp_bytes) {
  if (H <= 0 || W <= 0) return "views: empty image";
  for (int i = 0; i < n_views; ++i) {
    const ttl_view_spec& s = specs[i];
    ViewDesc d{};
    d.img_off = img_off;
    d.W = W;
    d.flip = s.flip != 0;
    if (s.kind == TTL_VIEW_CLEAN) {            // Resize(size, BICUBIC) + CenterCrop(size): ttl.py:232-234
      int nh, nw;
      resized_size(H, W, size, &nh, &nw);
      d.x0 = 0; d.y0 = 0; d.w = W; d.h = H; d.ow = nw; d.oh = nh; d.bicubic = 1;
      d.oy = static_cast<int>(std::nearbyint((nh - size) / 2.0));   // Python round(): half to even
      d.ox = static_cast<int>(std::nearbyint((nw - size) / 2.0));
    } else if (s.kind == TTL_VIEW_CROP) {      // RandomResizedCrop(size) box + flip: data/datautils.py:98-101
      if (s.height <= 0 || s.width <= 0 || s.top < 0 || s.left < 0 || s.top + s.height > H || s.left + s.width > W)
        return "views: crop box outside the image";
      d.x0 = s.left; d.y0 = s.top; d.w = s.width; d.h = s.height; d.ow = size; d.oh = size; d.bicubic = 0;
      d.ox = 0; d.oy = 0;
    } else {
      return "views: unknown view kind";
    }
    d.ksh = ksize_of(d.w, d.ow, d.bicubic != 0);
    d.ksv = ksize_of(d.h, d.oh, d.bicubic != 0);
    d.coef_h_off = static_cast<long long>(*coef_ints);
    *coef_ints += static_cast<size_t>(size) * (2 + d.ksh);
    d.coef_v_off = static_cast<long long>(*coef_ints);
    *coef_ints += static_cast<size_t>(size) * (2 + d.ksv);
    d.tmp_off = static_cast<long long>(*tmp_bytes);
    *tmp_bytes += (static_cast<size_t>(d.h) * size * 3 + 15) / 16 * 16;
    out[i] = d;
  }
  return nullptr;
}

void launch_views(const uint8_t* img, const ViewDesc* desc, int n_views, int max_h, int* coef, uint8_t* tmp, float* views,
                  bf16* patches, int size, int p, const float* mean, const float* sd, cudaStream_t st) {
  launch_pdl(views_coeff_kernel, dim3(2 * n_views), dim3(256), 0, st, desc, coef, size);
  int bx = (max_h * size + 255) / 256;
  if (bx > 64) bx = 64;
  launch_pdl(views_hpass_kernel, dim3(bx, n_views), dim3(256), 0, st, img, desc, static_cast<const int*>(coef), tmp, size);
  int bv = (size * size + 255) / 256;
  if (bv > 64) bv = 64;
  const int K = 3 * p * p, Kp = (K + 63) / 64 * 64;
  launch_pdl(views_vpass_kernel, dim3(bv, n_views), dim3(256), 0, st, static_cast<const uint8_t*>(tmp), desc,
             static_cast<const int*>(coef), views, patches, size, p, Kp, mean[0], mean[1], mean[2], sd[0], sd[1], sd[2]);
  if (patches != nullptr && Kp > K) {
    const int rows = n_views * (size / p) * (size / p);
    launch_pdl(views_pad_kernel, dim3(148), dim3(256), 0, st, patches, rows, K, Kp);
  }
}

}  // namespace ttl
