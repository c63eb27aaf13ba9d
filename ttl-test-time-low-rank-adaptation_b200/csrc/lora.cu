// LoRA-side kernels: packing the fp32 master factors into the bf16 operands of the GEMM's second operand pair,
// the tall-skinny weight-gradient reductions (dB = s dY^T (X A^T), dA = s (dY B)^T X; peft LoRA Linear backward),
// and the fused AdamW step / adapter+optimiser reset (torch.optim.AdamW at ttl.py:218; LoRA_AB.reset at
// clip/custom_clip.py:202-215 + optimizer.load_state_dict at ttl.py:344).
#include "kernels.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdlib>

namespace ttl {

namespace {

// params layout per layer: A_q[r,d] | B_q[d,r] | A_v[r,d] | B_v[d,r]   (fp32, contiguous)
struct LoraPackedArr { LoraPacked p[16]; };

__global__ void lora_pack_kernel(const float* __restrict__ prm0, int64_t sample_stride, LoraPackedArr pks, int64_t layer_stride,
                                 int d, int r, float s, int S) {
  pdl_wait();
  pdl_trigger();
  const LoraPacked pk = pks.p[blockIdx.z];          // one grid plane per LoRA layer
  prm0 += blockIdx.z * layer_stride;
  const int n_a = 64 * d;          // a_ext / a_ext_t elements of one sample block
  const int n_b = 3 * d * 64;      // b_ext / b_ext_t elements of one sample block
  const int smp = blockIdx.y, kc = 64 * S, c0 = 64 * smp;   // this sample's 64-column block of the K-concatenation
  const float* prm = prm0 + smp * sample_stride;
  const float* A_q = prm;
  const float* B_q = prm + r * d;
  const float* A_v = prm + 2 * r * d;
  const float* B_v = prm + 3 * r * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_a + n_b; i += gridDim.x * blockDim.x) {
    if (i < n_a) {
      const int j = i / d, k = i - j * d;   // a_ext[j, k]
      float v = 0.f;
      if (j < r) v = A_q[j * d + k];
      else if (j < 2 * r) v = A_v[(j - r) * d + k];
      const bf16 b = __float2bfloat16(v);
      pk.a_ext[static_cast<size_t>(c0 + j) * d + k] = b;
      pk.a_ext_t[static_cast<size_t>(k) * kc + c0 + j] = b;
    } else {
      const int e = i - n_a;
      const int n = e / 64, j = e - n * 64;  // b_ext[n, j], n in [0, 3d)
      float v = 0.f;
      if (n < d) { if (j < r) v = s * B_q[n * r + j]; }
      else if (n >= 2 * d) { if (j >= r && j < 2 * r) v = s * B_v[(n - 2 * d) * r + (j - r)]; }
      const bf16 b = __float2bfloat16(v);
      pk.b_ext[static_cast<size_t>(n) * kc + c0 + j] = b;
      pk.b_ext_t[static_cast<size_t>(c0 + j) * 3 * d + n] = b;
    }
  }
}

// partial[mc, w, j] = sum_{m in chunk mc} wide[m, w] * narrow[m, j]     (w in a 64-wide block, j < 16*?)
constexpr int SR_MC = 128;  // rows per CTA
template <int NN>           // narrow width (16 or 32)
__global__ void __launch_bounds__(256)
skinny_partial_kernel(const bf16* __restrict__ wide, int ldw, const bf16* __restrict__ narrow, int ldn, int M, int nw,
                      float* __restrict__ ws, int narrow_gstride) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) bf16 sW[SR_MC][64 + 8];
  __shared__ __align__(16) bf16 sN[SR_MC][NN + 8];
  const int w0 = blockIdx.x * 64, m0 = blockIdx.y * SR_MC, grp = blockIdx.z;
  wide += static_cast<size_t>(grp) * M * ldw;                              // this group's rows
  narrow += static_cast<size_t>(grp) * M * ldn + grp * narrow_gstride;
  ws += static_cast<size_t>(grp) * gridDim.y * nw * NN;
  for (int i = threadIdx.x; i < SR_MC * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m0 + r < M) v = *reinterpret_cast<const uint4*>(wide + static_cast<size_t>(m0 + r) * ldw + w0 + c);
    *reinterpret_cast<uint4*>(&sW[r][c]) = v;
  }
  for (int i = threadIdx.x; i < SR_MC * (NN / 8); i += blockDim.x) {
    const int r = i / (NN / 8), c = (i % (NN / 8)) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m0 + r < M) v = *reinterpret_cast<const uint4*>(narrow + static_cast<size_t>(m0 + r) * ldn + c);
    *reinterpret_cast<uint4*>(&sN[r][c]) = v;
  }
  __syncthreads();
  constexpr int JPT = NN / 4;              // outputs per thread along j
  const int w = threadIdx.x & 63, jg = threadIdx.x >> 6;  // 4 groups of JPT
  float acc[JPT];
#pragma unroll
  for (int j = 0; j < JPT; ++j) acc[j] = 0.f;
  for (int m = 0; m < SR_MC; ++m) {
    const float a = __bfloat162float(sW[m][w]);
#pragma unroll
    for (int j = 0; j < JPT; ++j) acc[j] += a * __bfloat162float(sN[m][jg * JPT + j]);
  }
  float* o = ws + (static_cast<size_t>(blockIdx.y) * nw + w0 + w) * NN + jg * JPT;
#pragma unroll
  for (int j = 0; j < JPT; ++j) o[j] = acc[j];
}

// The same partial products on the tensor cores (narrow width 16): out[64 w x 16 j] = W^T [64 x 128 rows] . N [128 rows x 16] per CTA,
// four warps, one m16 tile of w each, mma.sync.m16n8k16 with both operands taken transposed out of the row-major staging
// (ldmatrix.trans).  The scalar kernel above needed 3.5 instructions per MAC: 198 us per launch at 64 views x 9 samples (DeYO head).
__global__ void __launch_bounds__(128)
skinny_partial_mma_kernel(const bf16* __restrict__ wide, int ldw, const bf16* __restrict__ narrow, int ldn, int M, int nw,
                          float* __restrict__ ws, int narrow_gstride) {
  pdl_wait();
  pdl_trigger();
  constexpr int NN = 16;
  __shared__ __align__(16) bf16 sW[SR_MC][64 + 8];
  __shared__ __align__(16) bf16 sN[SR_MC][NN + 8];
  const int w0 = blockIdx.x * 64, m0 = blockIdx.y * SR_MC, grp = blockIdx.z;
  wide += static_cast<size_t>(grp) * M * ldw;
  narrow += static_cast<size_t>(grp) * M * ldn + grp * narrow_gstride;
  ws += static_cast<size_t>(grp) * gridDim.y * nw * NN;
  for (int i = threadIdx.x; i < SR_MC * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m0 + r < M) v = *reinterpret_cast<const uint4*>(wide + static_cast<size_t>(m0 + r) * ldw + w0 + c);
    *reinterpret_cast<uint4*>(&sW[r][c]) = v;
  }
  for (int i = threadIdx.x; i < SR_MC * (NN / 8); i += blockDim.x) {
    const int r = i / (NN / 8), c = (i % (NN / 8)) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m0 + r < M) v = *reinterpret_cast<const uint4*>(narrow + static_cast<size_t>(m0 + r) * ldn + c);
    *reinterpret_cast<uint4*>(&sN[r][c]) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, mi = lane >> 3;
  const int wt = warp * 16;                      // this warp's 16 columns of `wide` = rows of the output tile
  float acc[2][4] = {};
#pragma unroll
  for (int k0 = 0; k0 < SR_MC; k0 += 16) {
    uint32_t a[4], b[4];
    // A = W^T (16 w x 16 rows): 8x8 blocks (w 0-7 | 8-15) x (rows 0-7 | 8-15), stored row-major by rows -> .trans
    ldsm_x4_t(smem_u32(&sW[k0 + (mi >> 1) * 8 + (lane & 7)][wt + (mi & 1) * 8]), a[0], a[1], a[2], a[3]);
    // B = N (16 rows x 16 j) stored [rows][j] -> .trans; (b0, b1): j 0-7, (b2, b3): j 8-15
    ldsm_x4_t(smem_u32(&sN[k0 + (mi & 1) * 8 + (lane & 7)][(mi >> 1) * 8]), b[0], b[1], b[2], b[3]);
    mma_bf16_16816(acc[0], a, b[0], b[1]);
    mma_bf16_16816(acc[1], a, b[2], b[3]);
  }
  const int r0 = wt + (lane >> 2), c0 = (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    float* o0 = ws + (static_cast<size_t>(blockIdx.y) * nw + w0 + r0) * NN + nt * 8 + c0;
    float* o1 = o0 + 8 * NN;
    *reinterpret_cast<float2*>(o0) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(o1) = make_float2(acc[nt][2], acc[nt][3]);
  }
}

// The weight-gradient reduction on the 5th-gen tensor cores: partial[seg, w, j] = sum_{m in segment} wide[m, w] * narrow[m, j] for a
// 128-column block of `wide` (dY or X) and the 16 columns of `narrow` (X A^T or dY B), i.e. D[128 x 16] += W^T[128 x rows] N[rows x 16]
// with the reduction dimension = rows.  Both operands are consumed as they sit in memory: row-major [rows][cols] is the MN-major
// form of the transposed operand (TMA tiles of 128 rows, 128B swizzle for the 2 x 64-column atoms of `wide`, 32B swizzle for the
// 16 columns of `narrow`); one tcgen05.mma 128 x 16 x 16 per 16 rows, fp32 accumulator in TMEM (32 columns).  A CTA streams one
// segment of rows through a 3-stage TMA ring (two CTAs per SM); the mma.sync version staged every tile through registers and ran at 39 % of the DRAM peak.
constexpr int ST_STAGES = 3, ST_ROWS = 128;      // 3 x 36 KB: two CTAs per SM
constexpr int ST_STAGE_BYTES = 2 * ST_ROWS * 128 + ST_ROWS * 32;      // wide: two [128][64] bf16 atoms; narrow: [128][16] bf16

__global__ void __launch_bounds__(128)
skinny_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN, int M, int nw, int seg_rows,
                 float* __restrict__ ws) {
  extern __shared__ uint8_t smem_st_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_st_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST_STAGES * ST_STAGE_BYTES);
  uint64_t* full = bars;                    // [ST_STAGES] tile landed
  uint64_t* empty = bars + ST_STAGES;       // [ST_STAGES] the MMAs that read the tile have completed
  uint64_t* done = bars + 2 * ST_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = blockIdx.x * 128, seg = blockIdx.y, grp = blockIdx.z;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmN);
    for (int i = 0; i < ST_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();
  pdl_trigger();
  const int row0 = seg * seg_rows;
  const int rows = M - row0 < seg_rows ? M - row0 : seg_rows;            // rows of this segment (> 0 by construction of the grid)
  const int nck = (rows + ST_ROWS - 1) / ST_ROWS;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, 16, 1) | (1u << 15);      // A and B MN-major
    auto load = [&](int i) {
      const int st = i % ST_STAGES;
      uint8_t* a = smem + st * ST_STAGE_BYTES;
      mbar_expect_tx(&full[st], ST_STAGE_BYTES);
      tma_load_3d(&tmW, &full[st], a, w0, row0 + i * ST_ROWS, grp);                        // rows beyond the group's M are zero-filled
      tma_load_3d(&tmW, &full[st], a + ST_ROWS * 128, w0 + 64, row0 + i * ST_ROWS, grp);
      tma_load_3d(&tmN, &full[st], a + 2 * ST_ROWS * 128, 0, row0 + i * ST_ROWS, grp);
    };
    for (int i = 0; i < nck && i < ST_STAGES; ++i) load(i);
    for (int i = 0; i < nck; ++i) {
      const int st = i % ST_STAGES;
      const uint32_t ph = (i / ST_STAGES) & 1;
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t a = smem_u32(smem + st * ST_STAGE_BYTES), b = a + 2 * ST_ROWS * 128;
      // seg_rows is a multiple of the 128-row tile, so a tile never straddles two segments; rows beyond the group's M are zero
#pragma unroll
      for (int ks = 0; ks < ST_ROWS / 16; ++ks)
        umma_bf16(tmem, umma_desc(a + ks * 2048, 1024, ST_ROWS * 128, 2), umma_desc(b + ks * 512, 256, 0, 6), idesc, (i | ks) != 0 ? 1u : 0u);
      umma_commit(&empty[st]);
      if (i + ST_STAGES < nck) {
        mbar_wait(&empty[st], ph);
        load(i + ST_STAGES);
      }
    }
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  {
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    tmem_ld_wait();
    float* o = ws + ((static_cast<size_t>(grp) * gridDim.y + seg) * nw + w0 + warp * 32 + lane) * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      reinterpret_cast<float4*>(o)[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                                    __uint_as_float(r[4 * q + 3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

__global__ void skinny_final_kernel(const float* __restrict__ ws, int chunks, int nw, int nn, float scale,
                                    float* __restrict__ out, int transpose_out, int64_t out_gstride) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw * nn) return;
  ws += static_cast<size_t>(blockIdx.y) * chunks * nw * nn;
  out += blockIdx.y * out_gstride;
  float a = 0.f;
  for (int c = 0; c < chunks; ++c) a += ws[static_cast<size_t>(c) * nw * nn + i];
  const int w = i / nn, j = i - w * nn;
  if (transpose_out) out[static_cast<size_t>(j) * nw + w] = a * scale;
  else out[i] = a * scale;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int n, float lr, float b1, float b2, float eps, float wd,
                             float step_size, float bc2_sqrt) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = m[i] + (gi - m[i]) * (1.0f - b1);            // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
  const float denom = sqrtf(vi) / bc2_sqrt + eps;           // (sqrt(v) / sqrt(bc2)).add_(eps)
  pi -= step_size * (mi / denom);                               // addcdiv_(m, denom, -lr / bc1)
  p[i] = pi; m[i] = mi; v[i] = vi;
}

__global__ void lora_reset_kernel(float* __restrict__ p, const float* __restrict__ p0, float* __restrict__ m,
                                  float* __restrict__ v, int n, int n0) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = p0[i % n0]; m[i] = 0.f; v[i] = 0.f;
}

}  // namespace

void launch_lora_pack(const float* params, int64_t sample_stride, LoraPacked pk, int d, int r, float s, int S, cudaStream_t st) {
  const int total = 64 * d + 3 * d * 64;
  LoraPackedArr a;
  a.p[0] = pk;
  launch_pdl(lora_pack_kernel, dim3((total + 255) / 256, S, 1), dim3(256), 0, st, params, sample_stride, a, static_cast<int64_t>(0), d, r,
             s, S);
}

void launch_lora_pack_layers(const float* params, int64_t layer_stride, int64_t sample_stride, const LoraPacked* pks,
                             int n_layers, int d, int r, float s, int S, cudaStream_t st) {
  const int total = 64 * d + 3 * d * 64;
  for (int l0 = 0; l0 < n_layers; l0 += 16) {
    const int n = n_layers - l0 < 16 ? n_layers - l0 : 16;
    LoraPackedArr a;
    for (int i = 0; i < n; ++i) a.p[i] = pks[l0 + i];
    launch_pdl(lora_pack_kernel, dim3((total + 255) / 256, S, n), dim3(256), 0, st, params + l0 * layer_stride, sample_stride, a,
               layer_stride, d, r, s, S);
  }
}

// tcgen05 / TMA route of launch_skinny_reduce (narrow width 16, nw a multiple of 128); false = not applicable, use the mma.sync kernels
static bool launch_skinny_tc(const bf16* wide, int ldw, int nw, const bf16* narrow, int ldn, int M, float scale, float* out,
                             int transpose_out, float* ws, int groups, int narrow_gstride, int64_t out_gstride, cudaStream_t st) {
  if (nw % 128 != 0 || M < 2 * ST_ROWS || (reinterpret_cast<uintptr_t>(wide) & 15) || (reinterpret_cast<uintptr_t>(narrow) & 15) ||
      (ldw % 8) || (ldn % 8) || (narrow_gstride % 8))
    return false;
  const int dv = current_device_slot();
  static int num_sms_dev[MAX_DEVICES] = {};
  static bool configured_dev[MAX_DEVICES] = {};
  if (num_sms_dev[dv] == 0) cudaDeviceGetAttribute(&num_sms_dev[dv], cudaDevAttrMultiProcessorCount, dv);
  const size_t smem = ST_STAGES * ST_STAGE_BYTES + 256 + 1024;
  if (!configured_dev[dv]) {
    if (cudaFuncSetAttribute(skinny_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured_dev[dv] = true;
  }
  // segments of rows per (column block, group): ONE wave of at most two CTAs per SM (a second, partial wave cost 25 % at
  // M = 113 472), at least two 128-row tiles per CTA; the workspace the callers provide holds (ceil(M / 128) + groups) x nw x 32
  // floats per group set, far more than nseg x nw x 16 per group
  const int col_blocks = nw / 128, nck_total = (M + ST_ROWS - 1) / ST_ROWS;
  int nseg = 2 * num_sms_dev[dv] / (col_blocks * groups);
  if (nseg > nck_total / 2) nseg = nck_total / 2;
  if (nseg < 1) nseg = 1;
  const int seg_rows = (nck_total + nseg - 1) / nseg * ST_ROWS;
  nseg = (M + seg_rows - 1) / seg_rows;
  CUtensorMap tw, tn;
  const uint64_t wd[3] = {static_cast<uint64_t>(nw), static_cast<uint64_t>(M), static_cast<uint64_t>(groups)};
  const uint64_t wstr[2] = {static_cast<uint64_t>(ldw) * 2, static_cast<uint64_t>(M) * ldw * 2};
  const uint32_t wbox[3] = {64, ST_ROWS, 1};
  if (!encode_tiled_map(&tw, 0, wide, 3, wd, wstr, wbox, 128)) return false;
  const uint64_t nd[3] = {16, static_cast<uint64_t>(M), static_cast<uint64_t>(groups)};
  const uint64_t nstr[2] = {static_cast<uint64_t>(ldn) * 2, (static_cast<uint64_t>(M) * ldn + narrow_gstride) * 2};
  const uint32_t nbox[3] = {16, ST_ROWS, 1};
  if (!encode_tiled_map(&tn, 0, narrow, 3, nd, nstr, nbox, 32)) return false;
  launch_pdl(skinny_tc_kernel, dim3(col_blocks, nseg, groups), dim3(128), smem, st, tw, tn, M, nw, seg_rows, ws);
  launch_pdl(skinny_final_kernel, dim3(dim3((nw * 16 + 255) / 256, groups)), dim3(256), 0, st, static_cast<const float*>(ws), nseg, nw, 16, scale,
             out, transpose_out, out_gstride);
  return true;
}

void launch_skinny_reduce(const bf16* wide, int ldw, int nw, const bf16* narrow, int ldn, int nn, int M, float scale,
                          float* out, int transpose_out, float* ws, int groups, int narrow_gstride, int64_t out_gstride,
                          cudaStream_t st) {
  // TTL_SKINNY: unset / "tc" = tcgen05 + TMA kernel where applicable, "mma" = mma.sync kernel, "scalar" = CUDA-core kernel
  static const char* sk_mode = std::getenv("TTL_SKINNY");
  if (nn == 16 && (sk_mode == nullptr || sk_mode[0] == 't') &&
      launch_skinny_tc(wide, ldw, nw, narrow, ldn, M, scale, out, transpose_out, ws, groups, narrow_gstride, out_gstride, st))
    return;
  const int chunks = (M + SR_MC - 1) / SR_MC;
  dim3 grid(nw / 64, chunks, groups);
  static const char* sk_env = std::getenv("TTL_SKINNY");      // "scalar": the CUDA-core kernel (A/B reference)
  static const bool scalar = sk_env != nullptr && sk_env[0] == 's';
  if (nn == 16 && !scalar) launch_pdl(skinny_partial_mma_kernel, dim3(grid), dim3(128), 0, st, wide, ldw, narrow, ldn, M, nw, ws, narrow_gstride);
  else if (nn == 16) launch_pdl(skinny_partial_kernel<16>, dim3(grid), dim3(256), 0, st, wide, ldw, narrow, ldn, M, nw, ws, narrow_gstride);
  else launch_pdl(skinny_partial_kernel<32>, dim3(grid), dim3(256), 0, st, wide, ldw, narrow, ldn, M, nw, ws, narrow_gstride);
  launch_pdl(skinny_final_kernel, dim3(dim3((nw * nn + 255) / 256, groups)), dim3(256), 0, st, ws, chunks, nw, nn, scale, out, transpose_out,
                                                                           out_gstride);
}

void launch_adamw(float* p, const float* g, float* m, float* v, int n, int step, float lr, float b1, float b2,
                  float eps, float wd, cudaStream_t st) {
  const double bc1 = 1.0 - pow(static_cast<double>(b1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(b2), step);
  launch_pdl(adamw_kernel, dim3((n + 255) / 256), dim3(256), 0, st, p, g, m, v, n, lr, b1, b2, eps, wd, static_cast<float>(lr / bc1),
                                                static_cast<float>(sqrt(bc2)));
}

void launch_lora_reset(float* p, const float* p0, float* m, float* v, int n, int n0, cudaStream_t st) {
  launch_pdl(lora_reset_kernel, dim3((n + 255) / 256), dim3(256), 0, st, p, p0, m, v, n, n0);
}

}  // namespace ttl
