// fp32 validation mode (ttl_config.precision = TTL_PRECISION_FP32): every activation of the path stays fp32 and every
// contraction accumulates fp32 products of fp32 operands, so the library can be held to the north-star's fp32 tolerance
// (logits / LoRA updates within 1e-4 of the reference's fp32 CPU run) -- the bf16 tensor-core path is held to 1e-2.
// It is a checker for the fast path, not the product: plain CUDA-core kernels, one sample at a time, no graphs.
//   sgemm_kernel        C = alpha * A[M,K] B^T (+ bias, + residual | QuickGELU | dQuickGELU | patch scatter), B as [N,K] or [K,N]
//   attention_f32_*     per (head, view) CTA, K/V (and Q/dO for the backward) staged in smem, one warp per row, exact softmax
//   layernorm_f32       LayerNorm with fp32 output; im2col_f32; reduce_tn_f32 (weight-gradient reductions dB, dA)
#include "kernels.cuh"
#include "ptx.cuh"

namespace ttl {

namespace {

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;
constexpr int DH = 64;

__device__ __forceinline__ float qgelu_sig(float z) { return 1.0f / (1.0f + expf(-1.702f * z)); }

template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(const SgemmArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sA[SG_BK][SG_BM + 4];
  __shared__ float sB[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;      // 16 x 16 threads, 8 rows x 4 columns each
  const int row0 = blockIdx.y * SG_BM, col0 = blockIdx.x * SG_BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < a.K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {       // A tile: 128 rows x 16 k, one float4 along k per (thread, i)
      const int r = (tid >> 2) + 64 * i, c4 = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < a.M && k0 + c4 < a.K) v = *reinterpret_cast<const float4*>(a.A + static_cast<size_t>(row0 + r) * a.lda + k0 + c4);
      sA[c4][r] = v.x; sA[c4 + 1][r] = v.y; sA[c4 + 2][r] = v.z; sA[c4 + 3][r] = v.w;
    }
    if (!a.b_kn) {                      // B[N][K]: 64 rows x 16 k
      const int n = tid >> 2, c4 = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + n < a.N && k0 + c4 < a.K) v = *reinterpret_cast<const float4*>(a.B + static_cast<size_t>(col0 + n) * a.ldb + k0 + c4);
      sB[c4][n] = v.x; sB[c4 + 1][n] = v.y; sB[c4 + 2][n] = v.z; sB[c4 + 3][n] = v.w;
    } else {                            // B[K][N]: 16 k x 64 columns
      const int kk = tid >> 4, n4 = (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < a.K && col0 + n4 < a.N) v = *reinterpret_cast<const float4*>(a.B + static_cast<size_t>(k0 + kk) * a.ldb + col0 + n4);
      *reinterpret_cast<float4*>(&sB[kk][n4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float av[8], bv[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = sA[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + ty * 8 + i;
    if (row >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= a.N) continue;
      float v = a.alpha * acc[i][j];
      const size_t o = static_cast<size_t>(row) * a.ldo + col;
      if (EPI == SE_LINEAR) {
        if (a.bias != nullptr) v += a.bias[col];
        if (a.resid != nullptr) v += a.resid[static_cast<size_t>(row) * a.ldr + col];
        a.out[o] = v;
      } else if (EPI == SE_GELU) {
        if (a.bias != nullptr) v += a.bias[col];
        if (a.out2 != nullptr) a.out2[o] = v;
        a.out[o] = v * qgelu_sig(v);
      } else if (EPI == SE_GELU_BWD) {
        const float z = a.aux[o], s = qgelu_sig(z);
        a.out[o] = v * s * (1.0f + 1.702f * z * (1.0f - s));
      } else {   // SE_PATCH: row = (view, patch) -> token row view*(T+1)+1+patch, + position embedding
        const int view = row / a.tpv, patch = row - view * a.tpv;
        a.out[(static_cast<size_t>(view) * (a.tpv + 1) + 1 + patch) * a.ldo + col] = v + a.pos[static_cast<size_t>(1 + patch) * a.N + col];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention, fp32
// smem rows padded to 65 floats: lanes read different rows of the same column.
constexpr int LDF = DH + 1;

__global__ void __launch_bounds__(256)
attention_f32_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ lse, int tokens, int heads,
                         float scale, int causal) {      // causal: query r sees keys 0..r (text tower)
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm_att[];
  float* sK = sm_att;
  float* sV = sK + tokens * LDF;
  float* sQ = sV + tokens * LDF;            // [warps][64]
  float* sP = sQ + 8 * DH;                  // [warps][tokens]
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* base = qkv + static_cast<size_t>(view) * tokens * ld + h * DH;
  for (int i = threadIdx.x; i < tokens * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    sK[r * LDF + c] = base[static_cast<size_t>(r) * ld + d + c];
    sV[r * LDF + c] = base[static_cast<size_t>(r) * ld + 2 * d + c];
  }
  __syncthreads();
  float* q = sQ + warp * DH;
  float* p = sP + warp * tokens;
  for (int r = warp; r < tokens; r += nw) {
    q[lane] = base[static_cast<size_t>(r) * ld + lane];
    q[lane + 32] = base[static_cast<size_t>(r) * ld + lane + 32];
    __syncwarp();
    const int nk = causal ? r + 1 : tokens;
    float mx = -INFINITY;
    for (int j = lane; j < nk; j += 32) {
      const float* kr = sK + j * LDF;
      float acc = 0.f;
#pragma unroll 16
      for (int c = 0; c < DH; ++c) acc = fmaf(q[c], kr[c], acc);
      acc *= scale;
      p[j] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < nk; j += 32) {
      const float e = expf(p[j] - mx);
      p[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < nk; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, sV[j * LDF + lane], o0);
      o1 = fmaf(pj, sV[j * LDF + lane + 32], o1);
    }
    const float inv = 1.f / sum;
    float* orow = out + (static_cast<size_t>(view) * tokens + r) * d + h * DH;
    orow[lane] = o0 * inv;
    orow[lane + 32] = o1 * inv;
    if (lse != nullptr && lane == 0) lse[(static_cast<size_t>(view) * heads + h) * tokens + r] = mx + logf(sum);
    __syncwarp();
  }
}

// Backward: phase A (one warp per query row i): D_i = dO_i . O_i, dS_ij = p_ij (dO_i . V_j - D_i), dQ_i = scale sum_j dS_ij K_j;
// phase B (one warp per key row j): dV_j = sum_i p_ij dO_i, dK_j = scale sum_i dS_ij Q_i.  p recomputed from lse (exact).
__global__ void __launch_bounds__(256)
attention_f32_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                         const float* __restrict__ lse, float* __restrict__ dqkv, int tokens, int heads, float scale, int causal) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm_att[];
  float* sQ = sm_att;
  float* sK = sQ + tokens * LDF;
  float* sV = sK + tokens * LDF;
  float* sDO = sV + tokens * LDF;
  float* sD = sDO + tokens * LDF;           // [tokens]
  float* sL = sD + tokens;                  // [tokens]
  float* sW = sL + tokens;                  // [warps][tokens] scratch
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t vrow = static_cast<size_t>(view) * tokens;
  const float* base = qkv + vrow * ld + h * DH;
  for (int i = threadIdx.x; i < tokens * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    sQ[r * LDF + c] = base[static_cast<size_t>(r) * ld + c];
    sK[r * LDF + c] = base[static_cast<size_t>(r) * ld + d + c];
    sV[r * LDF + c] = base[static_cast<size_t>(r) * ld + 2 * d + c];
    sDO[r * LDF + c] = dout[(vrow + r) * d + h * DH + c];
  }
  __syncthreads();
  for (int r = warp; r < tokens; r += nw) {      // D_i and lse_i
    const float* orow = out + (vrow + r) * d + h * DH;
    float t = sDO[r * LDF + lane] * orow[lane] + sDO[r * LDF + lane + 32] * orow[lane + 32];
    t = warp_sum(t);
    if (lane == 0) { sD[r] = t; sL[r] = lse[(static_cast<size_t>(view) * heads + h) * tokens + r]; }
  }
  __syncthreads();
  float* w = sW + warp * tokens;
  // ---- phase A: dQ
  for (int i = warp; i < tokens; i += nw) {
    const float* qi = sQ + i * LDF;
    const float* doi = sDO + i * LDF;
    const int nk = causal ? i + 1 : tokens;
    for (int j = lane; j < nk; j += 32) {
      const float* kj = sK + j * LDF;
      const float* vj = sV + j * LDF;
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int c = 0; c < DH; ++c) { s = fmaf(qi[c], kj[c], s); dp = fmaf(doi[c], vj[c], dp); }
      const float pij = expf(s * scale - sL[i]);
      w[j] = pij * (dp - sD[i]);
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < nk; ++j) {
      const float ds = w[j];
      a0 = fmaf(ds, sK[j * LDF + lane], a0);
      a1 = fmaf(ds, sK[j * LDF + lane + 32], a1);
    }
    float* o = dqkv + (vrow + i) * ld + h * DH;
    o[lane] = a0 * scale;
    o[lane + 32] = a1 * scale;
    __syncwarp();
  }
  // ---- phase B: dK, dV
  for (int j = warp; j < tokens; j += nw) {
    const float* kj = sK + j * LDF;
    const float* vj = sV + j * LDF;
    float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
    // lanes split the query rows to compute p_ij and dS_ij, then broadcast through smem
    float* pw = w;                      // p_ij for this j, i in [0, tokens)
    for (int i0 = 0; i0 < tokens; i0 += 32) {
      const int i = i0 + lane;
      float pij = 0.f, ds = 0.f;
      if (i < tokens && (!causal || i >= j)) {
        const float* qi = sQ + i * LDF;
        const float* doi = sDO + i * LDF;
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int c = 0; c < DH; ++c) { s = fmaf(qi[c], kj[c], s); dp = fmaf(doi[c], vj[c], dp); }
        pij = expf(s * scale - sL[i]);
        ds = pij * (dp - sD[i]);
      }
      const int n = min(32, tokens - i0);
      for (int ii = 0; ii < n; ++ii) {
        const float pb = __shfl_sync(0xffffffffu, pij, ii), db = __shfl_sync(0xffffffffu, ds, ii);
        const float* qi = sQ + (i0 + ii) * LDF;
        const float* doi = sDO + (i0 + ii) * LDF;
        dv0 = fmaf(pb, doi[lane], dv0);
        dv1 = fmaf(pb, doi[lane + 32], dv1);
        dk0 = fmaf(db, qi[lane], dk0);
        dk1 = fmaf(db, qi[lane + 32], dk1);
      }
    }
    (void)pw;
    float* o = dqkv + (vrow + j) * ld + h * DH;
    o[d + lane] = dk0 * scale;
    o[d + lane + 32] = dk1 * scale;
    o[2 * d + lane] = dv0;
    o[2 * d + lane + 32] = dv1;
  }
}

// The same backward as two launches for token counts whose four staged matrices exceed shared memory (257 tokens: 267 KB):
// PHASE_DQ keeps K and V resident (Q_i, dO_i per warp), PHASE_DKV keeps Q and dO resident (K_j, V_j per warp).  Same arithmetic
// and summation order as attention_f32_bwd_kernel.
template <int PHASE_DKV>
__global__ void __launch_bounds__(256)
attention_f32_bwd_split_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                               const float* __restrict__ lse, float* __restrict__ dqkv, int tokens, int heads, float scale,
                               int causal) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm_att[];
  float* sA = sm_att;                       // PHASE_DQ: K        PHASE_DKV: Q
  float* sB = sA + tokens * LDF;            // PHASE_DQ: V        PHASE_DKV: dO
  float* sD = sB + tokens * LDF;            // [tokens] D_i = dO_i . O_i
  float* sL = sD + tokens;                  // [tokens] lse_i
  float* sR = sL + tokens;                  // [warps][2][64] this warp's two rows (Q_i, dO_i  or  K_j, V_j)
  float* sW = sR + 8 * 2 * DH;              // [warps][tokens] scratch (PHASE_DQ)
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t vrow = static_cast<size_t>(view) * tokens;
  const float* base = qkv + vrow * ld + h * DH;
  for (int i = threadIdx.x; i < tokens * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    if (PHASE_DKV) {
      sA[r * LDF + c] = base[static_cast<size_t>(r) * ld + c];
      sB[r * LDF + c] = dout[(vrow + r) * d + h * DH + c];
    } else {
      sA[r * LDF + c] = base[static_cast<size_t>(r) * ld + d + c];
      sB[r * LDF + c] = base[static_cast<size_t>(r) * ld + 2 * d + c];
    }
  }
  for (int r = warp; r < tokens; r += nw) {      // D_i and lse_i
    const float* orow = out + (vrow + r) * d + h * DH;
    const float* dor = dout + (vrow + r) * d + h * DH;
    float t = dor[lane] * orow[lane] + dor[lane + 32] * orow[lane + 32];
    t = warp_sum(t);
    if (lane == 0) { sD[r] = t; sL[r] = lse[(static_cast<size_t>(view) * heads + h) * tokens + r]; }
  }
  __syncthreads();
  float* r0 = sR + warp * 2 * DH;
  float* r1 = r0 + DH;
  if (!PHASE_DKV) {
    float* w = sW + warp * tokens;
    for (int i = warp; i < tokens; i += nw) {
      r0[lane] = base[static_cast<size_t>(i) * ld + lane];
      r0[lane + 32] = base[static_cast<size_t>(i) * ld + lane + 32];
      r1[lane] = dout[(vrow + i) * d + h * DH + lane];
      r1[lane + 32] = dout[(vrow + i) * d + h * DH + lane + 32];
      __syncwarp();
      const int nk = causal ? i + 1 : tokens;
      for (int j = lane; j < nk; j += 32) {
        const float* kj = sA + j * LDF;
        const float* vj = sB + j * LDF;
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int c = 0; c < DH; ++c) { s = fmaf(r0[c], kj[c], s); dp = fmaf(r1[c], vj[c], dp); }
        const float pij = expf(s * scale - sL[i]);
        w[j] = pij * (dp - sD[i]);
      }
      __syncwarp();
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < nk; ++j) {
        const float ds = w[j];
        a0 = fmaf(ds, sA[j * LDF + lane], a0);
        a1 = fmaf(ds, sA[j * LDF + lane + 32], a1);
      }
      float* o = dqkv + (vrow + i) * ld + h * DH;
      o[lane] = a0 * scale;
      o[lane + 32] = a1 * scale;
      __syncwarp();
    }
  } else {
    for (int j = warp; j < tokens; j += nw) {
      r0[lane] = base[static_cast<size_t>(j) * ld + d + lane];
      r0[lane + 32] = base[static_cast<size_t>(j) * ld + d + lane + 32];
      r1[lane] = base[static_cast<size_t>(j) * ld + 2 * d + lane];
      r1[lane + 32] = base[static_cast<size_t>(j) * ld + 2 * d + lane + 32];
      __syncwarp();
      float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
      for (int i0 = 0; i0 < tokens; i0 += 32) {
        const int i = i0 + lane;
        float pij = 0.f, ds = 0.f;
        if (i < tokens && (!causal || i >= j)) {
          const float* qi = sA + i * LDF;
          const float* doi = sB + i * LDF;
          float s = 0.f, dp = 0.f;
#pragma unroll 16
          for (int c = 0; c < DH; ++c) { s = fmaf(qi[c], r0[c], s); dp = fmaf(doi[c], r1[c], dp); }
          pij = expf(s * scale - sL[i]);
          ds = pij * (dp - sD[i]);
        }
        const int n = min(32, tokens - i0);
        for (int ii = 0; ii < n; ++ii) {
          const float pb = __shfl_sync(0xffffffffu, pij, ii), db = __shfl_sync(0xffffffffu, ds, ii);
          const float* qi = sA + (i0 + ii) * LDF;
          const float* doi = sB + (i0 + ii) * LDF;
          dv0 = fmaf(pb, doi[lane], dv0);
          dv1 = fmaf(pb, doi[lane + 32], dv1);
          dk0 = fmaf(db, qi[lane], dk0);
          dk1 = fmaf(db, qi[lane + 32], dk1);
        }
      }
      float* o = dqkv + (vrow + j) * ld + h * DH;
      o[d + lane] = dk0 * scale;
      o[d + lane + 32] = dk1 * scale;
      o[2 * d + lane] = dv0;
      o[2 * d + lane + 32] = dv1;
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------ row kernels, fp32 out
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int rows, int d, float eps) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + static_cast<size_t>(row) * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += xr[i];
  const float mean = warp_sum(s) / d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) { const float t = xr[i] - mean; q += t * t; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / d + eps);
  float* yr = y + static_cast<size_t>(row) * d;
  for (int i = lane; i < d; i += 32) yr[i] = (xr[i] - mean) * rstd * gamma[i] + beta[i];
}

__global__ void im2col_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int V, int S, int p, int Kp) {
  pdl_wait();
  pdl_trigger();
  const int gp = S / p, T = gp * gp, K = 3 * p * p;
  const size_t total = static_cast<size_t>(V) * T * Kp;
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(e % Kp);
    const size_t row = e / Kp;
    float v = 0.f;
    if (col < K) {
      const int view = static_cast<int>(row / T), patch = static_cast<int>(row % T);
      const int py = patch / gp, px = patch % gp;
      const int c = col / (p * p), rem = col % (p * p), i = rem / p, j = rem % p;
      v = img[((static_cast<size_t>(view) * 3 + c) * S + (py * p + i)) * S + px * p + j];
    }
    out[e] = v;
  }
}

// out[w, j] (or out[j, w]) = scale * sum_m Wd[m, w] * Nr[m, j];  one thread per output, rows summed in order (deterministic)
__global__ void __launch_bounds__(256)
reduce_tn_f32_kernel(const float* __restrict__ wide, int ldw, int nw, const float* __restrict__ narrow, int ldn, int nn, int M,
                     float scale, float* __restrict__ out, int transpose_out) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nw * nn) return;
  const int w = idx % nw, j = idx / nw;        // consecutive threads -> consecutive w: coalesced reads of `wide`
  float acc = 0.f;
  for (int m = 0; m < M; ++m) acc = fmaf(wide[static_cast<size_t>(m) * ldw + w], narrow[static_cast<size_t>(m) * ldn + j], acc);
  if (transpose_out) out[static_cast<size_t>(j) * nw + w] = acc * scale;
  else out[static_cast<size_t>(w) * nn + j] = acc * scale;
}

}  // namespace

cudaError_t launch_sgemm(const SgemmArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0 || a.K % 4 != 0 || a.lda % 4 != 0 || a.ldb % 4 != 0 || (a.b_kn && a.N % 4 != 0) ||
      a.A == nullptr || a.B == nullptr || a.out == nullptr)
    return cudaErrorInvalidValue;
  const dim3 grid((a.N + SG_BN - 1) / SG_BN, (a.M + SG_BM - 1) / SG_BM), block(256);
  switch (a.epi) {
    case SE_LINEAR: return launch_pdl(sgemm_kernel<SE_LINEAR>, grid, block, 0, st, a);
    case SE_GELU: return launch_pdl(sgemm_kernel<SE_GELU>, grid, block, 0, st, a);
    case SE_GELU_BWD: return launch_pdl(sgemm_kernel<SE_GELU_BWD>, grid, block, 0, st, a);
    case SE_PATCH: return launch_pdl(sgemm_kernel<SE_PATCH>, grid, block, 0, st, a);
    default: return cudaErrorInvalidValue;
  }
}

size_t attention_f32_fwd_smem(int tokens) { return (2 * static_cast<size_t>(tokens) * LDF + 8 * DH + 8 * tokens) * sizeof(float); }
static size_t attention_f32_bwd_smem_whole(int tokens) { return (4 * static_cast<size_t>(tokens) * LDF + 2 * tokens + 8 * tokens) * sizeof(float); }
static size_t attention_f32_bwd_smem_split(int tokens) { return (2 * static_cast<size_t>(tokens) * LDF + 2 * tokens + 16 * DH + 8 * tokens) * sizeof(float); }
// shared memory the backward needs: the one-launch kernel where its four matrices fit, the two-launch form otherwise
size_t attention_f32_bwd_smem(int tokens) {
  const size_t whole = attention_f32_bwd_smem_whole(tokens);
  return whole <= 227 * 1024 ? whole : attention_f32_bwd_smem_split(tokens);
}

void launch_attention_f32_fwd(const float* qkv, float* out, float* lse, int V, int tokens, int heads, float scale, cudaStream_t st,
                              int causal) {
  const size_t smem = attention_f32_fwd_smem(tokens);
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    cudaFuncSetAttribute(attention_f32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  launch_pdl(attention_f32_fwd_kernel, dim3(heads, V), dim3(256), smem, st, qkv, out, lse, tokens, heads, scale, causal);
}

void launch_attention_f32_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, int V,
                              int tokens, int heads, float scale, cudaStream_t st, int causal) {
  if (attention_f32_bwd_smem_whole(tokens) > 227 * 1024) {      // e.g. 257 tokens (ViT-L/14): dQ and dK/dV as two launches
    const size_t smem2 = attention_f32_bwd_smem_split(tokens);
    static size_t configured2_dev[MAX_DEVICES] = {};
    size_t& configured2 = configured2_dev[current_device_slot()];
    if (smem2 > configured2) {
      cudaFuncSetAttribute(attention_f32_bwd_split_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
      cudaFuncSetAttribute(attention_f32_bwd_split_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
      configured2 = smem2;
    }
    launch_pdl(attention_f32_bwd_split_kernel<0>, dim3(heads, V), dim3(256), smem2, st, qkv, out, dout, lse, dqkv, tokens, heads, scale, causal);
    launch_pdl(attention_f32_bwd_split_kernel<1>, dim3(heads, V), dim3(256), smem2, st, qkv, out, dout, lse, dqkv, tokens, heads, scale, causal);
    return;
  }
  const size_t smem = attention_f32_bwd_smem_whole(tokens);
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    cudaFuncSetAttribute(attention_f32_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  launch_pdl(attention_f32_bwd_kernel, dim3(heads, V), dim3(256), smem, st, qkv, out, dout, lse, dqkv, tokens, heads, scale, causal);
}

void launch_layernorm_f32(const float* x, float* y, const float* gamma, const float* beta, int rows, int d, float eps,
                          cudaStream_t st) {
  launch_pdl(layernorm_f32_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, x, y, gamma, beta, rows, d, eps);
}

void launch_im2col_f32(const float* images, float* patches, int V, int S, int p, cudaStream_t st) {
  const int K = 3 * p * p, Kp = (K + 63) / 64 * 64;
  const size_t total = static_cast<size_t>(V) * (S / p) * (S / p) * Kp;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(im2col_f32_kernel, dim3(blocks), dim3(256), 0, st, images, patches, V, S, p, Kp);
}

void launch_reduce_tn_f32(const float* wide, int ldw, int nw, const float* narrow, int ldn, int nn, int M, float scale,
                          float* out, int transpose_out, cudaStream_t st) {
  launch_pdl(reduce_tn_f32_kernel, dim3((nw * nn + 255) / 256), dim3(256), 0, st, wide, ldw, nw, narrow, ldn, nn, M, scale, out,
             transpose_out);
}

}  // namespace ttl
