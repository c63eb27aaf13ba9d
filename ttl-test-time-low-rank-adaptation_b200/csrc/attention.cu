// Fused multi-head attention over one view's tokens (197 for ViT-B/16), head_dim 64, no mask.
// One CTA per (head, view); Q/K/V (and dO) of that head live in shared memory; each warp owns 16-row tiles and
// walks the other dimension in chunks of 64 with an online (flash-style) softmax, so nothing of size
// tokens x tokens ever touches HBM.  fp32 softmax statistics (as HF eager_attention_forward: softmax in fp32).
// Tensor path: warp-level mma.sync m16n8k16 bf16 + ldmatrix (a tcgen05 version is the planned upgrade; attention
// is ~4 % of the path's FLOPs).
//   forward : O = softmax(scale * Q K^T) V, optional LSE for the backward
//   backward: phase A per 16-query tile -> dQ ; phase B per 16-key tile (transposed problem) -> dK, dV.
//             No atomics, deterministic.
#include "kernels.cuh"
#include "ptx.cuh"

namespace ttl {

namespace {

constexpr int DH = 64;
constexpr int LDS = 72;  // smem row pitch (elements): 144 B keeps ldmatrix rows on distinct banks
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void load_a(const bf16* s, int row0, int col0, int lane, uint32_t (&a)[4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = col0 + (lane >> 4) * 8;
  ldsm_x4(smem_u32(s + r * LDS + c), a[0], a[1], a[2], a[3]);
}
// B operand stored [n][k]:  (b0,b1) -> n-tile n0..n0+7, (b2,b3) -> n-tile n0+8..n0+15, k-step k0..k0+15
__device__ __forceinline__ void load_b_nk(const bf16* s, int n0, int k0, int lane, uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  const int r = n0 + (mi >> 1) * 8 + (lane & 7);
  const int c = k0 + (mi & 1) * 8;
  ldsm_x4(smem_u32(s + r * LDS + c), b[0], b[1], b[2], b[3]);
}
// B operand stored [k][n] (transposed on load)
__device__ __forceinline__ void load_b_kn(const bf16* s, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  const int r = k0 + (mi & 1) * 8 + (lane & 7);
  const int c = n0 + (mi >> 1) * 8;
  ldsm_x4_t(smem_u32(s + r * LDS + c), b[0], b[1], b[2], b[3]);
}

// acc[16 x 64] += A[16 x 64(k = head dim)] * B[n0..n0+63][k]^T          (scores: Q K^T, dO V^T, K Q^T, V dO^T)
__device__ __forceinline__ void mma_rows_nk(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* sB, int n0,
                                            int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_nk(sB, n0 + np * 16, ks * 16, lane, b);
      mma_bf16_16816(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16_16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// acc[16 x 64(n = head dim)] += P[16 x 64(k)] * B[k0..k0+63][n]              (P V, dS K, P^T dO, dS^T Q)
__device__ __forceinline__ void mma_rows_kn(float (&acc)[8][4], const uint32_t (&p)[4][4], const bf16* sB, int k0,
                                            int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      load_b_kn(sB, k0 + ks * 16, np * 16, lane, b);
      mma_bf16_16816(acc[2 * np], p[ks], b[0], b[1]);
      mma_bf16_16816(acc[2 * np + 1], p[ks], b[2], b[3]);
    }
  }
}
// 16x64 fp32 accumulator tile -> bf16 A-operand fragments for the next MMA
__device__ __forceinline__ void acc_to_afrag(const float (&s)[8][4], uint32_t (&p)[4][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    p[ks][0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
    p[ks][1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
    p[ks][2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
    p[ks][3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
  }
}
__device__ __forceinline__ void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// Copy one head's [tokens x 64] slice (row pitch ld elements) into smem [rows_pad][LDS], zero-filling the padding rows.
__device__ __forceinline__ void load_head_tile(bf16* s, const bf16* g, int ld, int tokens, int rows_pad) {
  for (int i = threadIdx.x; i < rows_pad * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < tokens) v = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(r) * ld + c);
    *reinterpret_cast<uint4*>(s + r * LDS + c) = v;
  }
}
// Write a warp-owned 16x64 fp32 tile as bf16 to global rows row0.. (< tokens), staging through the warp's own smem rows.
__device__ __forceinline__ void store_tile(bf16* stage, const float (&acc)[8][4], bf16* g, int ld, int row0, int tokens,
                                           int lane) {
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + (lane & 3) * 2;
    *reinterpret_cast<uint32_t*>(stage + (lane >> 2) * LDS + c) = pack_bf16(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<uint32_t*>(stage + ((lane >> 2) + 8) * LDS + c) = pack_bf16(acc[nt][2], acc[nt][3]);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = it * 32 + lane, r = i >> 3, c = (i & 7) * 8;
    if (row0 + r < tokens)
      *reinterpret_cast<uint4*>(g + static_cast<size_t>(row0 + r) * ld + c) = *reinterpret_cast<const uint4*>(stage + r * LDS + c);
  }
  __syncwarp();
}

__global__ void attention_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                                     int tokens, int heads, float scale_log2, int q_tiles, int nkp) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_att);
  bf16* sK = sQ + q_tiles * 16 * LDS;
  bf16* sV = sK + nkp * LDS;
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const bf16* base = qkv + static_cast<size_t>(view) * tokens * ld + h * DH;
  load_head_tile(sQ, base, ld, tokens, q_tiles * 16);
  load_head_tile(sK, base + d, ld, tokens, nkp);
  load_head_tile(sV, base + 2 * d, ld, tokens, nkp);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int qt = warp; qt < q_tiles; qt += nwarps) {
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) load_a(sQ, qt * 16, ks * 16, lane, aq[ks]);
    float o[8][4];
    zero_acc(o);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kc = 0; kc < nkp; kc += 64) {
      float s[8][4];
      zero_acc(s);
      mma_rows_nk(s, aq, sK, kc, lane);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kc + nt * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = (key + (e & 1) < tokens) ? s[nt][e] * scale_log2 : -INFINITY;
          s[nt][e] = v;
          if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: key 0 is always valid in chunk 0
      const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1;
        s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
        s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      uint32_t p[4][4];
      acc_to_afrag(s, p);
      mma_rows_kn(o, p, sV, kc, lane);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= i0; o[nt][1] *= i0; o[nt][2] *= i1; o[nt][3] *= i1; }
    bf16* gout = out + static_cast<size_t>(view) * tokens * d + h * DH;
    store_tile(sQ + qt * 16 * LDS, o, gout, d, qt * 16, tokens, lane);   // this warp's own (already consumed) Q rows
    if (lse != nullptr && (lane & 3) == 0) {
      const int r0 = qt * 16 + (lane >> 2), r1 = r0 + 8;
      float* L = lse + (static_cast<size_t>(view) * heads + h) * tokens;
      if (r0 < tokens) L[r0] = (m0 + log2f(l0)) * LN2;
      if (r1 < tokens) L[r1] = (m1 + log2f(l1)) * LN2;
    }
  }
}

__global__ void attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                                     const bf16* __restrict__ dout, const float* __restrict__ lse,
                                     bf16* __restrict__ dqkv, int tokens, int heads, float scale, int tiles, int nkp) {
  extern __shared__ __align__(16) uint8_t smem_att[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_att);
  bf16* sK = sQ + nkp * LDS;
  bf16* sV = sK + nkp * LDS;
  bf16* sDO = sV + nkp * LDS;
  float* sD = reinterpret_cast<float*>(sDO + nkp * LDS);  // rowsum(dO * O)
  float* sL = sD + nkp;                                    // lse in log2 units
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const bf16* base = qkv + static_cast<size_t>(view) * tokens * ld + h * DH;
  const bf16* gO = out + static_cast<size_t>(view) * tokens * d + h * DH;
  const bf16* gDO = dout + static_cast<size_t>(view) * tokens * d + h * DH;
  load_head_tile(sQ, base, ld, tokens, nkp);
  load_head_tile(sK, base + d, ld, tokens, nkp);
  load_head_tile(sV, base + 2 * d, ld, tokens, nkp);
  load_head_tile(sDO, gDO, d, tokens, nkp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* L = lse + (static_cast<size_t>(view) * heads + h) * tokens;
  for (int r = warp; r < nkp; r += nwarps) {
    float acc = 0.f;
    if (r < tokens) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(gO + static_cast<size_t>(r) * d + lane * 2);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(gDO + static_cast<size_t>(r) * d + lane * 2);
      const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
      acc = fa.x * fb.x + fa.y * fb.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      sD[r] = acc;
      sL[r] = r < tokens ? L[r] * LOG2E : INFINITY;   // +inf -> exp2(x - inf) = 0 for padding rows
    }
  }
  __syncthreads();
  const float scale_log2 = scale * LOG2E;
  bf16* gdq = dqkv + static_cast<size_t>(view) * tokens * ld + h * DH;

  // ---------------- phase A: dQ for 16-query tiles
  for (int qt = warp; qt < tiles; qt += nwarps) {
    uint32_t aq[4][4], ado[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      load_a(sQ, qt * 16, ks * 16, lane, aq[ks]);
      load_a(sDO, qt * 16, ks * 16, lane, ado[ks]);
    }
    const int r0 = qt * 16 + (lane >> 2), r1 = r0 + 8;
    const float L0 = sL[r0], L1 = sL[r1], D0 = sD[r0], D1 = sD[r1];
    float dq[8][4];
    zero_acc(dq);
    for (int kc = 0; kc < nkp; kc += 64) {
      float s[8][4], dp[8][4];
      zero_acc(s);
      zero_acc(dp);
      mma_rows_nk(s, aq, sK, kc, lane);
      mma_rows_nk(dp, ado, sV, kc, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kc + nt * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = key + (e & 1) < tokens;
          const float p = ok ? exp2f(s[nt][e] * scale_log2 - (e < 2 ? L0 : L1)) : 0.f;
          s[nt][e] = p * (dp[nt][e] - (e < 2 ? D0 : D1)) * scale;   // dS
        }
      }
      uint32_t ds[4][4];
      acc_to_afrag(s, ds);
      mma_rows_kn(dq, ds, sK, kc, lane);
    }
    // stage through global directly (Q rows in smem are still needed by phase B of other warps)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + (lane & 3) * 2;
      if (r0 < tokens) *reinterpret_cast<uint32_t*>(gdq + static_cast<size_t>(r0) * ld + c) = pack_bf16(dq[nt][0], dq[nt][1]);
      if (r1 < tokens) *reinterpret_cast<uint32_t*>(gdq + static_cast<size_t>(r1) * ld + c) = pack_bf16(dq[nt][2], dq[nt][3]);
    }
  }

  // ---------------- phase B: dK, dV for 16-key tiles (transposed problem: rows = keys, columns = queries)
  for (int kt = warp; kt < tiles; kt += nwarps) {
    uint32_t ak[4][4], av[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      load_a(sK, kt * 16, ks * 16, lane, ak[ks]);
      load_a(sV, kt * 16, ks * 16, lane, av[ks]);
    }
    float dk[8][4], dv[8][4];
    zero_acc(dk);
    zero_acc(dv);
    for (int qc = 0; qc < nkp; qc += 64) {
      float st[8][4], dpt[8][4];
      zero_acc(st);
      zero_acc(dpt);
      mma_rows_nk(st, ak, sQ, qc, lane);     // S^T = K_t Q^T
      mma_rows_nk(dpt, av, sDO, qc, lane);   // dP^T = V_t dO^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int q = qc + nt * 8 + (lane & 3) * 2;
        const float La = sL[q], Lb = sL[q + 1], Da = sD[q], Db = sD[q + 1];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = exp2f(st[nt][e] * scale_log2 - ((e & 1) ? Lb : La));   // 0 for padded queries (L = +inf)
          dpt[nt][e] = p * (dpt[nt][e] - ((e & 1) ? Db : Da)) * scale;           // dS^T
          st[nt][e] = p;                                                         // P^T
        }
      }
      uint32_t pt[4][4], dst[4][4];
      acc_to_afrag(st, pt);
      acc_to_afrag(dpt, dst);
      mma_rows_kn(dv, pt, sDO, qc, lane);    // dV += P^T dO
      mma_rows_kn(dk, dst, sQ, qc, lane);    // dK += dS^T Q
    }
    const int r0 = kt * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + (lane & 3) * 2;
      if (r0 < tokens) {
        *reinterpret_cast<uint32_t*>(gdq + static_cast<size_t>(r0) * ld + d + c) = pack_bf16(dk[nt][0], dk[nt][1]);
        *reinterpret_cast<uint32_t*>(gdq + static_cast<size_t>(r0) * ld + 2 * d + c) = pack_bf16(dv[nt][0], dv[nt][1]);
      }
      if (r1 < tokens) {
        *reinterpret_cast<uint32_t*>(gdq + static_cast<size_t>(r1) * ld + d + c) = pack_bf16(dk[nt][2], dk[nt][3]);
        *reinterpret_cast<uint32_t*>(gdq + static_cast<size_t>(r1) * ld + 2 * d + c) = pack_bf16(dv[nt][2], dv[nt][3]);
      }
    }
  }
}

inline int pick_warps(int tiles) {
  const int rounds = (tiles + 7) / 8;
  return (tiles + rounds - 1) / rounds;
}

}  // namespace

size_t attention_fwd_smem(int tokens) {
  const int q_tiles = (tokens + 15) / 16, nkp = (tokens + 63) / 64 * 64;
  return static_cast<size_t>(q_tiles * 16 + 2 * nkp) * LDS * sizeof(bf16);
}
size_t attention_bwd_smem(int tokens) {
  const int nkp = (tokens + 63) / 64 * 64;
  return static_cast<size_t>(4 * nkp) * LDS * sizeof(bf16) + 2 * nkp * sizeof(float);
}

void launch_attention_fwd(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                          cudaStream_t st) {
  const int q_tiles = (tokens + 15) / 16, nkp = (tokens + 63) / 64 * 64;
  const size_t smem = attention_fwd_smem(tokens);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  attention_fwd_kernel<<<dim3(heads, V), pick_warps(q_tiles) * 32, smem, st>>>(qkv, out, lse, tokens, heads,
                                                                               scale * LOG2E, q_tiles, nkp);
}

void launch_attention_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int V,
                          int tokens, int heads, float scale, cudaStream_t st) {
  const int tiles = (tokens + 15) / 16, nkp = (tokens + 63) / 64 * 64;
  const size_t smem = attention_bwd_smem(tokens);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  attention_bwd_kernel<<<dim3(heads, V), pick_warps(tiles) * 32, smem, st>>>(qkv, out, dout, lse, dqkv, tokens, heads,
                                                                             scale, tiles, nkp);
}

}  // namespace ttl
