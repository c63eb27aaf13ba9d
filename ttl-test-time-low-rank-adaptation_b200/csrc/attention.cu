// Fused multi-head attention over one view's tokens (197 for ViT-B/16), head_dim 64, no mask.
// One CTA per (head, view); Q/K/V (and dO) of that head live in shared memory; each warp owns 16-row tiles and
// walks the other dimension in chunks of 64 with an online (flash-style) softmax, so nothing of size
// tokens x tokens ever touches HBM.  fp32 softmax statistics (as HF eager_attention_forward: softmax in fp32).
// Tensor path: warp-level mma.sync m16n8k16 bf16 + ldmatrix (a tcgen05 version is the planned upgrade; attention
// is ~4 % of the path's FLOPs).
//   forward : O = softmax(scale * Q K^T) V, optional LSE for the backward
//   backward: phase A per 16-query tile -> dQ ; phase B per 16-key tile (transposed problem) -> dK, dV.
//             No atomics, deterministic.
#include <type_traits>

#include "kernels.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
// NV (compile time): only the first NV 16-wide groups of the 64 columns hold real rows of B; the rest (padding) is skipped
template <int NV = 4>
__device__ __forceinline__ void mma_rows_nk(float (&acc)[8][4], const uint32_t (&a)[4][4], const bf16* sB, int n0,
                                            int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < NV; ++np) {
      uint32_t b[4];
      load_b_nk(sB, n0 + np * 16, ks * 16, lane, b);
      mma_bf16_16816(acc[2 * np], a[ks], b[0], b[1]);
      mma_bf16_16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// acc[16 x 64(n = head dim)] += P[16 x 64(k)] * B[k0..k0+63][n]              (P V, dS K, P^T dO, dS^T Q)
// NV (compile time): only the first NV 16-deep k-steps of P are non-zero
template <int NV = 4>
__device__ __forceinline__ void mma_rows_kn(float (&acc)[8][4], const uint32_t (&p)[4][4], const bf16* sB, int k0,
                                            int lane) {
#pragma unroll
  for (int ks = 0; ks < NV; ++ks) {
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
// Four independent 16-byte loads are in flight per thread before the first store (the plain loop serialised on the
// global-load latency: ~9 dependent round trips per operand).
__device__ __forceinline__ void load_head_tile(bf16* s, const bf16* g, int ld, int tokens, int rows_pad) {
  const int total = rows_pad * 8, step = blockDim.x;
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * step) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * step, r = i >> 3, c = (i & 7) * 8;
      v[u] = make_uint4(0, 0, 0, 0);
      if (i < total && r < tokens) v[u] = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(r) * ld + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * step, r = i >> 3, c = (i & 7) * 8;
      if (i < total) *reinterpret_cast<uint4*>(s + r * LDS + c) = v[u];
    }
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

// causal != 0: key j contributes to query i only when j <= i (the text tower of `--lora_encoder text`, HF CLIPTextTransformer's
// causal mask); here and in attention_bwd_kernel below.
__global__ void attention_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                                     int tokens, int heads, float scale_log2, int q_tiles, int nkp, int causal) {
  pdl_wait();
  pdl_trigger();
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
      const int qr0 = qt * 16 + (lane >> 2), qr1 = qr0 + 8;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kc + nt * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kk = key + (e & 1);
          const bool ok = kk < tokens && (!causal || kk <= (e < 2 ? qr0 : qr1));
          const float v = ok ? s[nt][e] * scale_log2 : -INFINITY;
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
                                     bf16* __restrict__ dqkv, int tokens, int heads, float scale, int tiles, int nkp, int causal,
                                     const float* __restrict__ delta) {
  pdl_wait();
  pdl_trigger();
  // Shared memory per CTA: the two matrices the phase uses as B operands in full (phase A: K, V; phase B: Q, dO), the row
  // statistics, and per warp a 2 x 16-row staging area for the A-operand rows of the tile it is working on (phase A: Q, dO rows;
  // phase B: K, V rows).  Half of what keeping all four matrices resident took: two CTAs per SM at 197 tokens.
  extern __shared__ __align__(16) uint8_t smem_att[];
  const bool phase_a = blockIdx.z == 0;
  bf16* sF0 = reinterpret_cast<bf16*>(smem_att);          // phase A: K      phase B: Q
  bf16* sF1 = sF0 + nkp * LDS;                            // phase A: V      phase B: dO
  float* sD = reinterpret_cast<float*>(sF1 + nkp * LDS);  // rowsum(dO * O)
  float* sL = sD + nkp;                                    // lse in log2 units
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  bf16* sW0 = reinterpret_cast<bf16*>(sL + nkp) + warp * (2 * 16 * LDS);
  bf16* sW1 = sW0 + 16 * LDS;
  const bf16* base = qkv + static_cast<size_t>(view) * tokens * ld + h * DH;
  const bf16* gO = out + static_cast<size_t>(view) * tokens * d + h * DH;
  const bf16* gDO = dout + static_cast<size_t>(view) * tokens * d + h * DH;
  bf16 *sQ, *sK, *sV, *sDO;
  if (phase_a) {
    sK = sF0; sV = sF1; sQ = nullptr; sDO = nullptr;
    load_head_tile(sK, base + d, ld, tokens, nkp);
    load_head_tile(sV, base + 2 * d, ld, tokens, nkp);
  } else {
    sQ = sF0; sDO = sF1; sK = nullptr; sV = nullptr;
    load_head_tile(sQ, base, ld, tokens, nkp);
    load_head_tile(sDO, gDO, d, tokens, nkp);
  }
  // 16 rows x 64 columns of a global matrix into the warp's staging area (rows >= tokens zero)
  auto stage16 = [&](bf16* dst, const bf16* g, int g_ld, int row0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = lane + u * 32, r = i >> 3, c = (i & 7) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row0 + r < tokens) v = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(row0 + r) * g_ld + c);
      *reinterpret_cast<uint4*>(dst + r * LDS + c) = v;
    }
  };
  const float* L = lse + (static_cast<size_t>(view) * heads + h) * tokens;
  // D[r] = rowsum(dO[r,:] * O[r,:]) and the log-sum-exp in log2 units: one thread per row, eight independent 16-byte loads
  // delta != nullptr: rowsum(P o dP) computed exactly by attention_delta_kernel (text tower, see there)
  const float* Dg = delta != nullptr ? delta + (static_cast<size_t>(view) * heads + h) * tokens : nullptr;
  for (int r = threadIdx.x; r < nkp; r += blockDim.x) {
    float acc = 0.f;
    if (r < tokens && Dg != nullptr) {
      acc = Dg[r];
    } else if (r < tokens) {
      uint4 a[8], b[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        a[c] = *reinterpret_cast<const uint4*>(gO + static_cast<size_t>(r) * d + c * 8);
        b[c] = *reinterpret_cast<const uint4*>(gDO + static_cast<size_t>(r) * d + c * 8);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a[c]);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b[c]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __bfloat1622float2(pa[e]), fb = __bfloat1622float2(pb[e]);
          acc += fa.x * fb.x + fa.y * fb.y;
        }
      }
    }
    sD[r] = acc;
    sL[r] = r < tokens ? L[r] * LOG2E : INFINITY;   // +inf -> exp2(x - inf) = 0 for padding rows
  }
  __syncthreads();
  const float scale_log2 = scale * LOG2E;
  const int tok16 = (tokens + 15) & ~15;
  bf16* gdq = dqkv + static_cast<size_t>(view) * tokens * ld + h * DH;

  // The two phases are independent: blockIdx.z picks one (gridDim.z == 2).
  const bool do_a = phase_a, do_b = !phase_a;
  // ---------------- phase A: dQ for 16-query tiles
  for (int qt = warp; do_a && qt < tiles; qt += nwarps) {
    uint32_t aq[4][4], ado[4][4];
    __syncwarp();
    stage16(sW0, base, ld, qt * 16);
    stage16(sW1, gDO, d, qt * 16);
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      load_a(sW0, 0, ks * 16, lane, aq[ks]);
      load_a(sW1, 0, ks * 16, lane, ado[ks]);
    }
    const int r0 = qt * 16 + (lane >> 2), r1 = r0 + 8;
    const float L0 = sL[r0], L1 = sL[r1], D0 = sD[r0], D1 = sD[r1];
    float dq[8][4];
    zero_acc(dq);
    // one 64-key chunk; NV = real 16-key groups in it (4 except for the last chunk: 197 tokens -> 1), a compile-time constant
    auto chunk_a = [&](const int kc, auto nvc) {
      constexpr int NV = decltype(nvc)::value;
      float s[8][4], dp[8][4];
      zero_acc(s);
      zero_acc(dp);
      mma_rows_nk<NV>(s, aq, sK, kc, lane);
      mma_rows_nk<NV>(dp, ado, sV, kc, lane);
#pragma unroll
      for (int nt = 0; nt < 2 * NV; ++nt) {
        const int key = kc + nt * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = key + (e & 1) < tokens && (!causal || key + (e & 1) <= (e < 2 ? r0 : r1));
          const float p = ok ? ex2_approx(s[nt][e] * scale_log2 - (e < 2 ? L0 : L1)) : 0.f;
          s[nt][e] = p * (dp[nt][e] - (e < 2 ? D0 : D1)) * scale;   // dS
        }
      }
      uint32_t ds[4][4];
      acc_to_afrag(s, ds);
      mma_rows_kn<NV>(dq, ds, sK, kc, lane);
    };
    int kc = 0;
    for (; kc + 64 <= tok16; kc += 64) chunk_a(kc, std::integral_constant<int, 4>{});
    switch ((tok16 - kc) >> 4) {
      case 1: chunk_a(kc, std::integral_constant<int, 1>{}); break;
      case 2: chunk_a(kc, std::integral_constant<int, 2>{}); break;
      case 3: chunk_a(kc, std::integral_constant<int, 3>{}); break;
      default: break;
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
  for (int kt = warp; do_b && kt < tiles; kt += nwarps) {
    uint32_t ak[4][4], av[4][4];
    __syncwarp();
    stage16(sW0, base + d, ld, kt * 16);
    stage16(sW1, base + 2 * d, ld, kt * 16);
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      load_a(sW0, 0, ks * 16, lane, ak[ks]);
      load_a(sW1, 0, ks * 16, lane, av[ks]);
    }
    float dk[8][4], dv[8][4];
    zero_acc(dk);
    zero_acc(dv);
    const int kr0 = kt * 16 + (lane >> 2), kr1 = kr0 + 8;     // the key rows of this thread's accumulator elements
    auto chunk_b = [&](const int qc, auto nvc) {
      constexpr int NV = decltype(nvc)::value;
      float st[8][4], dpt[8][4];
      zero_acc(st);
      zero_acc(dpt);
      mma_rows_nk<NV>(st, ak, sQ, qc, lane);     // S^T = K_t Q^T
      mma_rows_nk<NV>(dpt, av, sDO, qc, lane);   // dP^T = V_t dO^T
#pragma unroll
      for (int nt = 0; nt < 2 * NV; ++nt) {
        const int q = qc + nt * 8 + (lane & 3) * 2;
        const float La = sL[q], Lb = sL[q + 1], Da = sD[q], Db = sD[q + 1];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float p = ex2_approx(st[nt][e] * scale_log2 - ((e & 1) ? Lb : La));   // 0 for padded queries (L = +inf)
          if (causal && (e < 2 ? kr0 : kr1) > q + (e & 1)) p = 0.f;
          dpt[nt][e] = p * (dpt[nt][e] - ((e & 1) ? Db : Da)) * scale;           // dS^T
          st[nt][e] = p;                                                         // P^T
        }
      }
      uint32_t pt[4][4], dst[4][4];
      acc_to_afrag(st, pt);
      acc_to_afrag(dpt, dst);
      mma_rows_kn<NV>(dv, pt, sDO, qc, lane);    // dV += P^T dO
      mma_rows_kn<NV>(dk, dst, sQ, qc, lane);    // dK += dS^T Q
    };
    int qc = 0;
    for (; qc + 64 <= tok16; qc += 64) chunk_b(qc, std::integral_constant<int, 4>{});
    switch ((tok16 - qc) >> 4) {
      case 1: chunk_b(qc, std::integral_constant<int, 1>{}); break;
      case 2: chunk_b(qc, std::integral_constant<int, 2>{}); break;
      case 3: chunk_b(qc, std::integral_constant<int, 3>{}); break;
      default: break;
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


// Delta_i = sum_j P_ij dP_ij (= dO_i . O_i in exact arithmetic) for the backward, in fp32 from Q, K, V, dO and the forward's lse.
// The backward kernels normally take Delta as rowsum(dO o O) with the bf16 O of the tape: dS = P o (dP - Delta) then carries
// |Delta| * 2^-9 of rounding noise.  In the image tower that is harmless; in the text tower of `--lora_encoder text` the V rows
// of a prompt are nearly identical (shared prompt prefix), dP_ij barely depends on j, |Delta| is ~25 x |dP - Delta| and the
// noise reaches 10-17 % of dQ (measured against the reference).  One CTA per (head, sequence), one warp per query row; K / V
// staged in smem as fp32 (tokens <= 128).
__global__ void __launch_bounds__(256)
attention_delta_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ lse,
                       float* __restrict__ delta, int tokens, int heads, float scale, int causal) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm_delta[];
  const int ldk = DH + 1;
  float* sK = sm_delta;                         // [tokens][65]
  float* sV = sK + tokens * ldk;                // [tokens][65]
  float* sQ = sV + tokens * ldk;                // [warps][2][64]: q row, dO row
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const bf16* base = qkv + static_cast<size_t>(view) * tokens * ld + h * DH;
  const bf16* gdo = dout + static_cast<size_t>(view) * tokens * d + h * DH;
  for (int i = threadIdx.x; i < tokens * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    sK[r * ldk + c] = __bfloat162float(base[static_cast<size_t>(r) * ld + d + c]);
    sV[r * ldk + c] = __bfloat162float(base[static_cast<size_t>(r) * ld + 2 * d + c]);
  }
  __syncthreads();
  float* q = sQ + warp * 2 * DH;
  float* dorow = q + DH;
  const float* L = lse + (static_cast<size_t>(view) * heads + h) * tokens;
  for (int r = warp; r < tokens; r += nw) {
    q[lane] = __bfloat162float(base[static_cast<size_t>(r) * ld + lane]);
    q[lane + 32] = __bfloat162float(base[static_cast<size_t>(r) * ld + lane + 32]);
    dorow[lane] = __bfloat162float(gdo[static_cast<size_t>(r) * d + lane]);
    dorow[lane + 32] = __bfloat162float(gdo[static_cast<size_t>(r) * d + lane + 32]);
    __syncwarp();
    const float lr = L[r];
    float acc = 0.f;
    const int jend = causal ? r + 1 : tokens;
    for (int j = lane; j < jend; j += 32) {
      const float* kr = sK + j * ldk;
      const float* vr = sV + j * ldk;
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int c = 0; c < DH; ++c) { s += q[c] * kr[c]; dp += dorow[c] * vr[c]; }
      acc += __expf(s * scale - lr) * dp;
    }
    acc = warp_sum(acc);
    if (lane == 0) delta[(static_cast<size_t>(view) * heads + h) * tokens + r] = acc;
    __syncwarp();
  }
}

// =============================================================================================== TMA-fed forward
// Persistent forward kernel for the 64-view pass: one CTA per SM walks (view, head) units; Q/K/V of the NEXT unit are
// fetched by three 3-D TMA loads (box 64 x rows x 1 of the [views][tokens][3d] tensor, 128-byte swizzle, tokens beyond
// the view zero-filled by the TMA) into the other half of a double buffer while the current unit is computed.
// Each warp owns 32 query rows (two m16 tiles share every K/V fragment read from smem) and walks the keys in chunks of
// 64 + one exact 16-key tail (197 tokens -> 208 keys, not 256).  The normalised tile goes back through the warp's own
// Q rows (swizzled) and one TMA store per warp; rows >= tokens are clipped by the TMA.
constexpr int ATT_MAX_WARPS = 8;   // up to 256 tokens; longer sequences (ViT-L/14: 257) use the cp-style kernel above

__device__ __forceinline__ uint32_t sw128(uint32_t base, int r, int c) {   // c: element column, multiple of 8
  return base + r * 128 + ((((c >> 3) ^ r) & 7) << 4);
}
__device__ __forceinline__ void lda_sw(uint32_t base, int row0, int col0, int lane, uint32_t (&a)[4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = col0 + (lane >> 4) * 8;
  ldsm_x4(sw128(base, r, c), a[0], a[1], a[2], a[3]);
}
__device__ __forceinline__ void ldb_nk_sw(uint32_t base, int n0, int k0, int lane, uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  ldsm_x4(sw128(base, n0 + (mi >> 1) * 8 + (lane & 7), k0 + (mi & 1) * 8), b[0], b[1], b[2], b[3]);
}
__device__ __forceinline__ void ldb_kn_sw(uint32_t base, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  ldsm_x4_t(sw128(base, k0 + (mi & 1) * 8 + (lane & 7), n0 + (mi >> 1) * 8), b[0], b[1], b[2], b[3]);
}

// One chunk of NT*8 keys for NMT m16 query tiles of one warp (online softmax, exp2 domain).
template <int NMT, int NT>
__device__ __forceinline__ void att_chunk(float (&o)[2][8][4], float (&mrow)[2][2], float (&lrow)[2][2],
                                          const uint32_t (&aq)[2][4][4], uint32_t sK, uint32_t sV, int kc, int tokens,
                                          float scale_log2, int lane) {
  float s[NMT][NT][4];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldb_nk_sw(sK, kc + np * 16, ks * 16, lane, b);
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt) {
        mma_bf16_16816(s[mt][2 * np], aq[mt][ks], b[0], b[1]);
        mma_bf16_16816(s[mt][2 * np + 1], aq[mt][ks], b[2], b[3]);
      }
    }
  }
  const bool tail = kc + NT * 8 > tokens;   // warp-uniform: some keys of this chunk are padding
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt) {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int key = kc + nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[mt][nt][e] * scale_log2;
        if (tail && key + (e & 1) >= tokens) v = -INFINITY;
        s[mt][nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(mrow[mt][0], mx0), mn1 = fmaxf(mrow[mt][1], mx1);   // finite: every chunk holds a real key
    const float c0 = exp2f(mrow[mt][0] - mn0), c1 = exp2f(mrow[mt][1] - mn1);
    mrow[mt][0] = mn0; mrow[mt][1] = mn1;
    float l0 = lrow[mt][0] * c0, l1 = lrow[mt][1] * c1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[mt][nt][0] *= c0; o[mt][nt][1] *= c0; o[mt][nt][2] *= c1; o[mt][nt][3] *= c1; }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[mt][nt][0] = exp2f(s[mt][nt][0] - mn0); s[mt][nt][1] = exp2f(s[mt][nt][1] - mn0);
      s[mt][nt][2] = exp2f(s[mt][nt][2] - mn1); s[mt][nt][3] = exp2f(s[mt][nt][3] - mn1);
      l0 += s[mt][nt][0] + s[mt][nt][1];
      l1 += s[mt][nt][2] + s[mt][nt][3];
    }
    lrow[mt][0] = l0; lrow[mt][1] = l1;
  }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t pf[NMT][4];
#pragma unroll
    for (int mt = 0; mt < NMT; ++mt) {
      pf[mt][0] = pack_bf16(s[mt][2 * kk][0], s[mt][2 * kk][1]);
      pf[mt][1] = pack_bf16(s[mt][2 * kk][2], s[mt][2 * kk][3]);
      pf[mt][2] = pack_bf16(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
      pf[mt][3] = pack_bf16(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldb_kn_sw(sV, kc + kk * 16, np * 16, lane, b);
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt) {
        mma_bf16_16816(o[mt][2 * np], pf[mt], b[0], b[1]);
        mma_bf16_16816(o[mt][2 * np + 1], pf[mt], b[2], b[3]);
      }
    }
  }
}

template <int NMT>
__device__ __forceinline__ void att_warp_rows(uint32_t sQ, uint32_t sK, uint32_t sV, int m0, int tokens, int rows_pad,
                                              float scale_log2, int lane, float* lse_unit) {
  uint32_t aq[2][4][4];
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lda_sw(sQ, m0 + mt * 16, ks * 16, lane, aq[mt][ks]);
  float o[2][8][4];
  float mrow[2][2], lrow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    mrow[mt][0] = mrow[mt][1] = -INFINITY;
    lrow[mt][0] = lrow[mt][1] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
  }
  int kc = 0;
  for (; kc + 64 <= rows_pad; kc += 64) att_chunk<NMT, 8>(o, mrow, lrow, aq, sK, sV, kc, tokens, scale_log2, lane);
  const int rem = rows_pad - kc;   // 0, 16, 32 or 48
  if (rem == 16) att_chunk<NMT, 2>(o, mrow, lrow, aq, sK, sV, kc, tokens, scale_log2, lane);
  else if (rem == 32) att_chunk<NMT, 4>(o, mrow, lrow, aq, sK, sV, kc, tokens, scale_log2, lane);
  else if (rem == 48) att_chunk<NMT, 6>(o, mrow, lrow, aq, sK, sV, kc, tokens, scale_log2, lane);
  __syncwarp();   // all lanes are done with this warp's Q rows (aq loaded long ago) -> reuse them as the output stage
#pragma unroll
  for (int mt = 0; mt < NMT; ++mt) {
    float l0 = lrow[mt][0], l1 = lrow[mt][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = m0 + mt * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t lo = pack_bf16(o[mt][nt][0] * i0, o[mt][nt][1] * i0), hi = pack_bf16(o[mt][nt][2] * i1, o[mt][nt][3] * i1);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw128(sQ, r0, nt * 8) + (lane & 3) * 4), "r"(lo) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw128(sQ, r1, nt * 8) + (lane & 3) * 4), "r"(hi) : "memory");
    }
    if (lse_unit != nullptr && (lane & 3) == 0) {
      if (r0 < tokens) lse_unit[r0] = (mrow[mt][0] + log2f(l0)) * LN2;
      if (r1 < tokens) lse_unit[r1] = (mrow[mt][1] + log2f(l1)) * LN2;
    }
  }
}

__global__ void __maxnreg__(224)
attention_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut,
                         float* __restrict__ lse, int tokens, int heads, int units, float scale_log2, int rows_alloc,
                         int rows_pad, int nbox, int box_rows) {
  extern __shared__ uint8_t smem_att_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_att_raw) + 1023) & ~uintptr_t(1023));
  const int RB = rows_alloc * 128;                      // bytes of one operand (rows_alloc x 64 bf16)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 6 * RB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = heads * DH;
  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmOut);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  auto issue = [&](int unit, int buf) {   // thread 0 only
    const int view = unit / heads, h = unit - view * heads;
    uint8_t* dst = smem + buf * 3 * RB;
    mbar_expect_tx(&full[buf], 3 * RB);
    for (int j = 0; j < nbox; ++j)
#pragma unroll
      for (int mat = 0; mat < 3; ++mat)
        tma_load_3d(&tmQKV, &full[buf], dst + mat * RB + j * box_rows * 128, mat * d + h * DH, j * box_rows, view);
  };
  if (tid == 0 && static_cast<int>(blockIdx.x) < units) issue(blockIdx.x, 0);
  int it = 0;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tid == 0 && unit + static_cast<int>(gridDim.x) < units) issue(unit + gridDim.x, buf ^ 1);
    mbar_wait(&full[buf], (it >> 1) & 1);
    const int view = unit / heads, h = unit - view * heads;
    const uint32_t sQ = smem_u32(smem + buf * 3 * RB), sK = sQ + RB, sV = sK + RB;
    const int m0 = warp * 32;
    if (m0 < rows_pad) {
      float* lse_unit = lse != nullptr ? lse + (static_cast<size_t>(view) * heads + h) * tokens : nullptr;
      if (rows_pad - m0 >= 32) att_warp_rows<2>(sQ, sK, sV, m0, tokens, rows_pad, scale_log2, lane, lse_unit);
      else att_warp_rows<1>(sQ, sK, sV, m0, tokens, rows_pad, scale_log2, lane, lse_unit);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && m0 < tokens) {
        tma_store_3d(&tmOut, smem + buf * 3 * RB + m0 * 128, h * DH, m0, view);
        bulk_commit();
        bulk_wait_read<0>();   // the stage is refilled by the TMA load issued after the barrier below
      }
    }
    __syncthreads();
  }
  if (lane == 0) bulk_wait<0>();
}


// =============================================================================================== tcgen05 forward
// Forward attention on the 5th-gen tensor cores.  Work item = (view, head, 128-query tile); two CTAs per SM (84 KB smem,
// 256 TMEM columns each) so one CTA's softmax overlaps the other's loads and MMAs.
//   warp 0 (one thread): TMA loads of Q tile / K / V (3-D maps, 128B swizzle, rows >= tokens zero-filled), then
//                        S = Q K^T   : 4 x tcgen05.mma 128 x keys x 16 (both operands K-major)           -> TMEM cols [0, keys)
//                        O = P V     : keys/16 x tcgen05.mma 128 x 64 x 16, A = P (bf16, smem, K-major), B = V as it sits
//                                      in memory ([key][dh] = MN-major operand)                          -> TMEM cols [0, 64)
//   warps 1-4: one thread per query row: two passes over the row in TMEM (max; exp2 + sum), P written as bf16 into the
//              UMMA K-major swizzled layout (64-key blocks with 128B swizzle over the dead Q/K tiles + a 16-key tail block
//              with 32B swizzle), then O / rowsum -> swizzled smem -> one TMA store per tile (rows >= tokens clipped).
// fp32 softmax statistics as HF eager attention (modeling_clip.py:261-279); P rounded to bf16 like the mma.sync kernels.
constexpr int TC_THREADS = 160;

__global__ void __launch_bounds__(TC_THREADS, 2)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmOut, float* __restrict__ lse, int tokens, int heads,
                        int items, int q_tiles, int keys, float scale_log2, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_tc_raw[];
#define TC_STAMP(k) do { if (dbg != nullptr && blockIdx.x < 8 && local_it < 16) dbg[(blockIdx.x * 16 + local_it) * 8 + (k)] = clock64(); } while (0)
  int local_it = 0;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_tc_raw) + 1023) & ~uintptr_t(1023));
  const int KB = keys * 128;                                   // bytes of the K (or V) tile, multiple of 1024 (keys % 16 == 0, see launcher)
  uint8_t* sQ = smem;                                          // 16 KB  [128 q][64 dh]      -> later P block 1
  uint8_t* sK = smem + 16384;                                  // KB     [keys][64 dh]       -> later P block 0 (+ 16-key tail block)
  uint8_t* sV = sK + KB;                                       // KB     [keys][64 dh]
  uint8_t* sP2 = sV + KB;                                      // 16 KB  P block 2           -> later the output stage
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP2 + 16384);
  uint64_t* bar_qk = bars;          // Q tile + K landed
  uint64_t* bar_v = bars + 1;       // V landed
  uint64_t* bar_s = bars + 2;       // S complete in TMEM
  uint64_t* bar_o = bars + 3;       // O complete in TMEM (and every smem operand of the item is dead)
  uint64_t* bar_tfree = bars + 4;   // O copied to registers: TMEM reusable (4 softmax warps)
  uint64_t* bar_p = bars + 5;       // [4] P block b written (4 softmax warps each); index 3 = the 16-key tail block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * DH;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_tfree, 4);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_p[i], 4);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();      // the prologue above overlaps the tail of the previous kernel
  pdl_trigger();
  const int n_full = keys / 64, tail = keys - n_full * 64;     // 64-key P blocks + a 16-key tail block (tail in {0,16})
  // P block b lives at: 0 -> sK, 1 -> sQ, 2 -> sP2 ; tail block -> sK + 16384
  const uint32_t p_addr0 = smem_u32(sK), p_addr1 = smem_u32(sQ), p_addr2 = smem_u32(sP2);
  const uint32_t p_tail = smem_u32(sK) + 16384;

  uint32_t ph = 0;   // parity of the per-item barriers
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA + MMA issue (one thread)
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, static_cast<uint32_t>(keys));
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV);
      const uint32_t v_lbo = static_cast<uint32_t>(KB);
      for (int item = blockIdx.x; item < items; item += gridDim.x, ph ^= 1, ++local_it) {
        const int unit = item / q_tiles, qt = item - unit * q_tiles;
        const int view = unit / heads, h = unit - view * heads;
        // every smem operand of the previous item is dead (bar_o was waited at the end of the previous iteration)
        TC_STAMP(0);
        mbar_expect_tx(bar_qk, 16384 + KB);
        tma_load_3d(&tmQ, bar_qk, sQ, h * DH, qt * 128, view);
        tma_load_3d(&tmKV, bar_qk, sK, d + h * DH, 0, view);
        mbar_expect_tx(bar_v, KB);
        tma_load_3d(&tmKV, bar_v, sV, 2 * d + h * DH, 0, view);
        if (item + static_cast<int>(gridDim.x) < items) {   // pull the next item's operands into L2 behind this item's compute
          const int nu = (item + gridDim.x) / q_tiles, nq = (item + gridDim.x) - nu * q_tiles;
          const int nv = nu / heads, nh = nu - nv * heads;
          tma_prefetch_l2_3d(&tmQ, nh * DH, nq * 128, nv);
          tma_prefetch_l2_3d(&tmKV, d + nh * DH, 0, nv);
          tma_prefetch_l2_3d(&tmKV, 2 * d + nh * DH, 0, nv);
        }
        mbar_wait(bar_qk, ph);
        TC_STAMP(1);
        if (local_it > 0) mbar_wait(bar_tfree, ph ^ 1);     // previous O has left TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)     // S = Q K^T
          umma_bf16(tmem, umma_desc_k_sw128(qa + k * 32), umma_desc_k_sw128(ka + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar_s);
        int kk = 0;                     // O = P V, block by block as the softmax warps deliver P
        for (int b = 0; b < n_full; ++b) {
          mbar_wait(&bar_p[b], ph);
          if (b == 0) mbar_wait(bar_v, ph);
          tc_fence_after();
          const uint32_t pa = b == 0 ? p_addr0 : (b == 1 ? p_addr1 : p_addr2);
#pragma unroll
          for (int j = 0; j < 4; ++j, ++kk)
            umma_bf16(tmem, umma_desc_k_sw128(pa + j * 32), umma_desc(va + kk * 2048, 1024, v_lbo, 2), idesc_o, kk != 0 ? 1u : 0u);
        }
        if (tail) {
          mbar_wait(&bar_p[3], ph);
          if (n_full == 0) mbar_wait(bar_v, ph);
          tc_fence_after();
          umma_bf16(tmem, umma_desc(p_tail, 256, 0, 6), umma_desc(va + kk * 2048, 1024, v_lbo, 2), idesc_o, kk != 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
        mbar_wait(bar_o, ph);           // MMAs complete: Q/K/V/P regions may be overwritten by the next item's loads
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: one thread per query row
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const int n32 = keys / 32, rem16 = keys - n32 * 32;   // full 32-column chunks + optional 16-column chunk
    for (int item = blockIdx.x; item < items; item += gridDim.x, ph ^= 1, ++local_it) {
      const int unit = item / q_tiles, qt = item - unit * q_tiles;
      const int view = unit / heads, h = unit - view * heads;
      const int r0 = qt * 128;
      const bool active = r0 + quad * 32 < tokens;    // warp-uniform: at least one real query row
      mbar_wait(bar_s, ph);
      if (threadIdx.x == 32) TC_STAMP(2);
      tc_fence_after();
      float m = -INFINITY, l = 0.f;
      if (active) {
        // ---- pass 1: exact row maximum (ALU + TMEM only; it overlaps the other resident CTA's MUFU-bound pass 2).
        // Columns >= tokens hold exact zeros (K rows beyond the view are zero-filled by the TMA): including them can only
        // raise the subtracted constant to 0 when every real score is negative, which softmax is invariant to.
        for (int c = 0; c < n32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(trow + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) m = max3(m, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        }
        if (rem16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(trow + n32 * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) m = max3(m, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        }
      }
      if (threadIdx.x == 32) {
        TC_STAMP(3);
        bulk_wait_read<0>();          // the previous item's output store has finished reading sP2 (= P block 2)
      }
      named_bar_sync(2, 128);
      if (active) {
        // ---- pass 2: p = 2^(s*scale - m*scale), row sum, bf16 P into the UMMA K-major layout, block by block
        const float ms = m * scale_log2;
        float l0 = 0.f, l1 = 0.f;
        for (int c = 0; c < n32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(trow + c * 32, r);
          tmem_ld_wait();
          const int blk = c >> 1;
          const uint32_t base = (blk == 0 ? p_addr0 : (blk == 1 ? p_addr1 : p_addr2)) + row * 128;
          if (c * 32 + 32 <= tokens) {     // warp-uniform: chunk entirely inside the real keys
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float e[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(r[q4 * 8 + i]), scale_log2, -ms));
              l0 += (e[0] + e[1]) + (e[2] + e[3]);
              l1 += (e[4] + e[5]) + (e[6] + e[7]);
              const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((chunk ^ (row & 7)) << 4)),
                           "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                           "r"(pack_bf16(e[6], e[7]))
                           : "memory");
            }
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float e[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float v = ex2_approx(fmaf(__uint_as_float(r[q4 * 8 + i]), scale_log2, -ms));
                if (c * 32 + q4 * 8 + i >= tokens) v = 0.f;
                e[i] = v;
                l0 += v;
              }
              const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((chunk ^ (row & 7)) << 4)),
                           "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                           "r"(pack_bf16(e[6], e[7]))
                           : "memory");
            }
          }
          if (c & 1) {                  // a 64-key block is complete: hand it to the MMA thread
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
            tc_fence_before();          // this warp's tcgen05.ld of S columns [0, 64*(blk+1)) are complete
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[blk]);
          }
        }
        if (rem16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(trow + n32 * 32, r);
          tmem_ld_wait();
          const uint32_t base = p_tail + row * 32;
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v = ex2_approx(fmaf(__uint_as_float(r[q2 * 8 + i]), scale_log2, -ms));
              if (n32 * 32 + q2 * 8 + i >= tokens) v = 0.f;
              e[i] = v;
              l0 += v;
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((q2 ^ ((row >> 2) & 1)) << 4)),
                         "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                         "r"(pack_bf16(e[6], e[7]))
                         : "memory");
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[3]);
        }
        l = l0 + l1;
      } else {
        tc_fence_before();
        if (lane == 0) {
          for (int b = 0; b < n_full; ++b) mbar_arrive(&bar_p[b]);
          if (tail) mbar_arrive(&bar_p[3]);
        }
      }
      if (threadIdx.x == 32) TC_STAMP(4);
      mbar_wait(bar_o, ph);
      if (threadIdx.x == 32) TC_STAMP(5);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (active) {
        tmem_ld_32x32b_x32(trow, o0);
        tmem_ld_32x32b_x32(trow + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tfree);      // the next item's S may overwrite the TMEM columns
      if (active) {
        const float inv = 1.f / l;
        const uint32_t obase = smem_u32(sP2) + row * 128;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const uint32_t* r = q4 < 4 ? o0 + q4 * 8 : o1 + (q4 - 4) * 8;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obase + ((static_cast<uint32_t>(q4) ^ (row & 7)) << 4)),
                       "r"(pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv))
                       : "memory");
        }
        if (lse != nullptr && r0 + row < tokens)
          lse[(static_cast<size_t>(view) * heads + h) * tokens + r0 + row] = (m * scale_log2 + log2f(l)) * LN2;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);       // the four softmax warps: output stage complete
      if (threadIdx.x == 32) {
        tma_store_3d(&tmOut, sP2, h * DH, r0, view);
        bulk_commit();
        TC_STAMP(6);
      }
    }
    if (threadIdx.x == 32) bulk_wait<0>();
  }
#undef TC_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}


// =============================================================================================== tcgen05 forward, two tiles in flight
// Same arithmetic as attention_fwd_tc_kernel, re-plumbed so the MUFU-bound softmax never waits for loads or for the other
// tile's MMAs.  One CTA per SM (220 KB smem, all 512 TMEM columns) walks (view, head) units; the two 128-query tiles of a
// unit are processed CONCURRENTLY by two softmax warpgroups and share one copy of K and V:
//   warp 0 (one thread)  TMA producer: K + both Q tiles of unit u+1 are (re)loaded as soon as both S MMAs of unit u have
//                        retired (P has its own smem, so K/Q die early); V is double-buffered across units
//   warp 1 (one thread)  tcgen05.mma issue: S0 = Q0 K^T -> TMEM [0,208), S1 = Q1 K^T -> TMEM [256,464), then O_t = P_t V block
//                        by block, alternating tiles as the warpgroups deliver P
//   warps 2-5 / 6-9      softmax warpgroup of tile 0 / tile 1: one thread per query row, exact two-pass softmax from TMEM,
//                        bf16 P into the UMMA K-major swizzled layout, O / rowsum -> swizzled smem -> one TMA store per tile
// K/V are read from L2 once per unit instead of once per tile, and the load latency of unit u+1 hides behind unit u.
// XK = true is the 257-token form (ViT-L/14): 256 keys and queries go through the MMAs (S0 / S1 fill all 512 TMEM columns, four
// 64-key P blocks per tile, ONE V buffer, 226 KB smem); key 256 is folded in by the softmax threads (q_i . k_256 from global
// memory while the S MMAs run, p_256 v_256 added to O in the epilogue) and query 256 is computed by
//   warps 10-11 (XK)     on the CUDA cores from the K / V tiles in shared memory; the producer waits for their bar_xk / bar_xv
//                        arrivals before it overwrites K / V.
constexpr int PP_THREADS = 320;
// Pass 2 is MUFU-bound (16 ex2 per clock and SM): PP_POLY_MASK picks the elements of every group of 8 whose 2^x is evaluated on
// the FMA pipe instead (Cody-Waite split with the 1.5 * 2^23 trick + cubic on [-0.5, 0.5], max relative error 7.5e-5, far
// below the bf16 rounding of P).  x <= 0 here (the row maximum has been subtracted).
#ifndef PP_POLY_MASK
#define PP_POLY_MASK 0x80      // 1 of 8: 77.1-77.8 us per layer at S=3 against 78.0-79.1 with 0x88 and 80.9-81.2 with 0 (gpurun s108)
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float xf = x + 12582912.f;                 // integer part (round to nearest) lands in the low mantissa bits
  const float fr = x - (xf - 12582912.f);          // [-0.5, 0.5]
  float p = fmaf(fr, 0.0551716685f, 0.2426111251f);
  p = fmaf(p, fr, 0.6932609677f);
  p = fmaf(p, fr, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
constexpr int PP_PTILE = 3 * 16384 + 4096;     // P of one tile: three 64-key blocks (128 rows x 128 B) + the 16-key tail block

constexpr int PP_THREADS_XK = PP_THREADS + 64;    // + two warps for the extra query row
template <bool XK>     // XK: the 257-token form (one extra key and query, four 64-key P blocks, one V buffer); false: 129..208 tokens
__global__ void __launch_bounds__(XK ? PP_THREADS_XK : PP_THREADS, 1)
attention_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmOut, float* __restrict__ lse, int tokens, int heads,
                        int units, int keys, float scale_log2, long long* __restrict__ dbg, int rev, int ptile_rt, int vbufs_rt,
                        int xkey_rt, const bf16* __restrict__ qkv, bf16* __restrict__ out_x) {
  // compile-time constants in the classic form, so that its code is what it was before the 257-token form existed
  const int ptile = XK ? ptile_rt : PP_PTILE, vbufs = XK ? vbufs_rt : 2, xkey = XK ? xkey_rt : -1;
  // ptile: bytes of one tile's P (PP_PTILE, or four 64-key blocks); vbufs: V buffers (2, or 1 when shared memory is short);
  // xkey >= 0: ONE extra key / value row (index xkey = keys) that the MMAs do not cover -- 257 tokens = 256 keys in TMEM
  // (2 tiles x 256 columns = all 512) + key 256 folded in by the softmax threads from global memory (qkv).
  extern __shared__ uint8_t smem_pp_raw[];
  // development aid (TTL_ATTN_DBG): clock64 stamps of the first units of CTA 0, per warpgroup: [it][t][stage]
#define PP_STAMP(k) do { if (dbg != nullptr && blockIdx.x == 0 && it < 12 && wg_tid == 0) dbg[(it * 2 + t) * 8 + (k)] = clock64(); } while (0)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_pp_raw) + 1023) & ~uintptr_t(1023));
  const int KB = keys * 128;                       // bytes of K (or V): multiple of 1024 (keys % 16 == 0, keys >= 64)
  uint8_t* sQ = smem;                              // [2 tiles] 16 KB each
  uint8_t* sK = sQ + 2 * 16384;
  uint8_t* sV = sK + KB;                           // [2 buffers]
  uint8_t* sP = sV + vbufs * KB;                   // [2 tiles] ptile bytes each; block 2 doubles as the tile's output stage
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ptile);
  uint64_t* bar_kq = bars;            // K + Q0 + Q1 of the unit landed
  uint64_t* bar_v = bars + 1;         // [2] V buffer landed
  uint64_t* bar_vfree = bars + 3;     // [2] every MMA that reads the V buffer has retired
  uint64_t* bar_s = bars + 5;         // [2] S_t complete in TMEM
  uint64_t* bar_o = bars + 7;         // [2] O_t complete in TMEM (P_t dead)
  uint64_t* bar_tfree = bars + 9;     // [2] O_t copied to registers: TMEM region t reusable
  uint64_t* bar_p = bars + 11;        // [2][4] P block b of tile t written (index 3 = tail block)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  uint64_t* bar_xk = bars + 20;       // XK: the extra-query warps have read K of the unit (2 arrivals)
  uint64_t* bar_xv = bars + 21;       // XK: ... and V

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * DH;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
    mbar_init(bar_kq, 1);
    mbar_init(bar_xk, 2);
    mbar_init(bar_xv, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_vfree[i], 1);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_tfree[i], 4);
      for (int b = 0; b < 4; ++b) mbar_init(&bar_p[i * 4 + b], 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();
  pdl_trigger();
  const int n_full = keys / 64, tail = keys - n_full * 64;      // 64-key P blocks + an optional 16-key tail block

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const int u2 = rev ? units - 1 - unit : unit;            // descending walk (kernels.cuh)
        const int view = u2 / heads, h = u2 - view * heads;
        if (it > 0) mbar_wait(&bar_s[1], (it - 1) & 1);          // both S MMAs of the previous unit retired: K, Q0, Q1 are dead
        if (XK && it > 0) mbar_wait(bar_xk, (it - 1) & 1);       // ... and the extra-query warps are done with K
        mbar_expect_tx(bar_kq, KB + 2 * 16384);
        tma_load_3d(&tmKV, bar_kq, sK, d + h * DH, 0, view);
        tma_load_3d(&tmQ, bar_kq, sQ, h * DH, 0, view);
        tma_load_3d(&tmQ, bar_kq, sQ + 16384, h * DH, 128, view);
        const int b = vbufs == 2 ? (it & 1) : 0, use = vbufs == 2 ? (it >> 1) : it;     // buffer and how often it has been used
        if (use >= 1) mbar_wait(&bar_vfree[b], (use - 1) & 1);
        if (XK && it > 0) mbar_wait(bar_xv, (it - 1) & 1);       // one V buffer in this form: the extra-query warps are done with it
        mbar_expect_tx(&bar_v[b], KB);
        tma_load_3d(&tmKV, &bar_v[b], sV + b * KB, 2 * d + h * DH, 0, view);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue
    // elect.sync rather than `lane == 0`: ptxas then emits the tcgen05.mma of a block back to back instead of one
    // ELECT / BRA.U.ANY retry loop per instruction (see attention_bwd_tc_kernel)
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, static_cast<uint32_t>(keys));
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
      const uint32_t v_lbo = static_cast<uint32_t>(KB);
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int b = vbufs == 2 ? (it & 1) : 0, use = vbufs == 2 ? (it >> 1) : it;
        mbar_wait(bar_kq, ph);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (it > 0) mbar_wait(&bar_tfree[t], ph ^ 1);           // the previous unit's O_t has left TMEM
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + t * 256, umma_desc_k_sw128(qa + t * 16384 + k * 32), umma_desc_k_sw128(ka + k * 32), idesc_s,
                      k != 0 ? 1u : 0u);
          umma_commit(&bar_s[t]);
        }
        mbar_wait(&bar_v[b], use & 1);
        const uint32_t va = smem_u32(sV + b * KB);
        int kk = 0;
        for (int blk = 0; blk < n_full; ++blk) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&bar_p[t * 4 + blk], ph);
            tc_fence_after();
            const uint32_t pa = smem_u32(sP + t * ptile + blk * 16384);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_bf16(tmem + t * 256, umma_desc_k_sw128(pa + j * 32), umma_desc(va + (kk + j) * 2048, 1024, v_lbo, 2), idesc_o,
                        (kk + j) != 0 ? 1u : 0u);
            if (!tail && blk == n_full - 1) umma_commit(&bar_o[t]);
          }
          kk += 4;
        }
        if (tail) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&bar_p[t * 4 + 3], ph);
            tc_fence_after();
            umma_bf16(tmem + t * 256, umma_desc(smem_u32(sP + t * ptile + 3 * 16384), 256, 0, 6),
                      umma_desc(va + kk * 2048, 1024, v_lbo, 2), idesc_o, kk != 0 ? 1u : 0u);
            umma_commit(&bar_o[t]);
          }
        }
        umma_commit(&bar_vfree[b]);
      }
    }
    __syncwarp();
  } else if (XK && warp >= 10) {
    // ------------------------------------------------------------------ query row xkey (XK): two warps, CUDA cores, K / V from the
    // tiles the unit has in shared memory anyway (a separate kernel would read all K and V a second time).
    // Scores: thread x takes keys x, x+64, x+128, x+192 (+ key xkey, redundantly, from global memory); softmax over the 257
    // scores through shuffles + a 64-thread named barrier; P V: thread x = output dimension x.
    const int x = threadIdx.x - PP_THREADS;
    float* xP = reinterpret_cast<float*>(bars + 32);     // [256] probabilities of the keys in shared memory
    float* xR = xP + 256;                                // [2] partial maxima, [2] partial sums
    const size_t ld = static_cast<size_t>(3) * d;
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int u2 = rev ? units - 1 - unit : unit;
      const int view = u2 / heads, h = u2 - view * heads;
      const bf16* rowx = qkv + (static_cast<size_t>(view) * tokens + xkey) * ld + h * DH;     // q | k | v of token xkey, this head
      float q[64];
      float sxx = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(rowx) + c8), b = __ldg(reinterpret_cast<const uint4*>(rowx + d) + c8);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fa = __bfloat1622float2(a2[j]), fb = __bfloat1622float2(b2[j]);
          q[c8 * 8 + 2 * j] = fa.x;
          q[c8 * 8 + 2 * j + 1] = fa.y;
          sxx = fmaf(fa.x, fb.x, sxx);
          sxx = fmaf(fa.y, fb.y, sxx);
        }
      }
      mbar_wait(bar_kq, ph);                     // K of the unit has landed (TMA, 128-byte swizzle: chunk c of row j at (c ^ (j & 7)) * 16)
      float sc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = x + 64 * i;
        const uint32_t rowa = smem_u32(sK) + j * 128;
        float acc = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 kk;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(kk.x), "=r"(kk.y), "=r"(kk.z), "=r"(kk.w)
                       : "r"(rowa + ((static_cast<uint32_t>(c8) ^ (j & 7)) << 4)));
          const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fk = __bfloat1622float2(k2[e]);
            acc = fmaf(q[c8 * 8 + 2 * e], fk.x, acc);
            acc = fmaf(q[c8 * 8 + 2 * e + 1], fk.y, acc);
          }
        }
        sc[i] = acc;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_xk);        // K may be overwritten
      float m = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), sxx);
      m = warp_max(m);
      if (lane == 0) xR[warp - 10] = m;
      named_bar_sync(7, 64);
      m = fmaxf(xR[0], xR[1]);
      const float ms = m * scale_log2;
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pv = ex2_approx(fmaf(sc[i], scale_log2, -ms));
        xP[x + 64 * i] = pv;
        lsum += pv;
      }
      const float pxx = ex2_approx(fmaf(sxx, scale_log2, -ms));
      if (x == 0) lsum += pxx;
      lsum = warp_sum(lsum);
      if (lane == 0) xR[2 + warp - 10] = lsum;
      named_bar_sync(7, 64);                     // partial sums and all of xP visible
      const float l = xR[2] + xR[3];
      mbar_wait(&bar_v[0], it & 1);              // V of the unit (one buffer in this form)
      const uint32_t vcol = smem_u32(sV) + (x & 7) * 2;
      const uint32_t vch = static_cast<uint32_t>(x >> 3);
      float acc = 0.f;
#pragma unroll 8
      for (int j = 0; j < 256; ++j) {
        uint16_t raw;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(raw) : "r"(vcol + j * 128 + ((vch ^ (j & 7)) << 4)));
        acc = fmaf(xP[j], __uint_as_float(static_cast<uint32_t>(raw) << 16), acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_xv);        // V may be overwritten
      acc = fmaf(pxx, __bfloat162float(rowx[2 * d + x]), acc);
      out_x[(static_cast<size_t>(view) * tokens + xkey) * d + h * DH + x] = __float2bfloat16(acc / l);
      if (x == 0 && lse != nullptr) lse[(static_cast<size_t>(view) * heads + h) * tokens + xkey] = (ms + log2f(l)) * LN2;
      named_bar_sync(7, 64);                     // xP / xR are rewritten in the next unit
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ softmax + epilogue warpgroup of tile t
    const int t = (warp - 2) >> 2;
    const int quad = warp & 3, row = quad * 32 + lane;
    const int wg_tid = threadIdx.x - 64 - t * 128;           // 0..127 inside the warpgroup
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16) + t * 256;
    const int n32 = keys / 32, rem16 = keys - n32 * 32;
    uint8_t* myP = sP + t * ptile;
    const uint32_t p_blk0 = smem_u32(myP), p_tail = p_blk0 + 3 * 16384;
    uint8_t* ostage = myP + 2 * 16384;
    const int r0 = t * 128;
    const bool active = r0 + quad * 32 < tokens;            // warp-uniform: at least one real query row
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int u2 = rev ? units - 1 - unit : unit;
      const int view = u2 / heads, h = u2 - view * heads;
      PP_STAMP(0);
      // extra key (xkey >= 0): its score q_i . k_x from global memory while the S MMAs run; every query row r0 + row is real here
      float sx = 0.f, px = 0.f;
      if (xkey >= 0 && active) {
        const size_t ld = static_cast<size_t>(3) * d;
        const uint4* qrow = reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(view) * tokens + r0 + row) * ld + h * DH);
        const uint4* krow = reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(view) * tokens + xkey) * ld + d + h * DH);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 a = __ldg(qrow + c8), b = __ldg(krow + c8);
          const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(a2[j]), fb = __bfloat1622float2(b2[j]);
            sx = fmaf(fa.x, fb.x, sx);
            sx = fmaf(fa.y, fb.y, sx);
          }
        }
      }
      mbar_wait(&bar_s[t], ph);
      PP_STAMP(1);
      tc_fence_after();
      float m = -INFINITY, l = 0.f;
      if (active) {
        // pass 1: exact row maximum.  Columns >= tokens hold exact zeros (zero-filled K rows): harmless for softmax.
        // Two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is reduced.
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32(trow, ra);
        for (int c = 0; c < n32; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n32) tmem_ld_32x32b_x32(trow + (c + 1) * 32, rb);
#pragma unroll
          for (int i = 0; i < 32; i += 2) m = max3(m, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
          if (c + 1 < n32) {
            tmem_ld_wait();
            if (c + 2 < n32) tmem_ld_32x32b_x32(trow + (c + 2) * 32, ra);
#pragma unroll
            for (int i = 0; i < 32; i += 2) m = max3(m, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
          }
        }
        if (rem16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(trow + n32 * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 2) m = max3(m, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        }
      }
      if (xkey >= 0) m = fmaxf(m, sx);
      PP_STAMP(2);
      if (wg_tid == 0) bulk_wait_read<0>();      // the previous unit's output store has finished reading the stage (= P block 2)
      named_bar_sync(1 + t, 128);
      PP_STAMP(3);
      if (active) {
        // pass 2: p = 2^(s*scale - m*scale), row sum, bf16 P into the UMMA K-major layout, block by block
        const float ms = m * scale_log2;
        float l0 = 0.f, l1 = 0.f;
        for (int c = 0; c < n32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(trow + c * 32, r);
          tmem_ld_wait();
          const int blk = c >> 1;
          const uint32_t base = p_blk0 + blk * 16384 + row * 128;
          if (c * 32 + 32 <= tokens) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float e[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x = fmaf(__uint_as_float(r[q4 * 8 + i]), scale_log2, -ms);
                e[i] = ((PP_POLY_MASK >> i) & 1) ? ex2_poly(x) : ex2_approx(x);
              }
              l0 += (e[0] + e[1]) + (e[2] + e[3]);
              l1 += (e[4] + e[5]) + (e[6] + e[7]);
              const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((chunk ^ (row & 7)) << 4)),
                           "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                           "r"(pack_bf16(e[6], e[7]))
                           : "memory");
            }
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float e[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float v = ex2_approx(fmaf(__uint_as_float(r[q4 * 8 + i]), scale_log2, -ms));
                if (c * 32 + q4 * 8 + i >= tokens) v = 0.f;
                e[i] = v;
                l0 += v;
              }
              const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((chunk ^ (row & 7)) << 4)),
                           "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                           "r"(pack_bf16(e[6], e[7]))
                           : "memory");
            }
          }
          if (c & 1) {
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[t * 4 + blk]);
          }
        }
        if (rem16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(trow + n32 * 32, r);
          tmem_ld_wait();
          const uint32_t base = p_tail + row * 32;
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float v = ex2_approx(fmaf(__uint_as_float(r[q2 * 8 + i]), scale_log2, -ms));
              if (n32 * 32 + q2 * 8 + i >= tokens) v = 0.f;
              e[i] = v;
              l0 += v;
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((q2 ^ ((row >> 2) & 1)) << 4)),
                         "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                         "r"(pack_bf16(e[6], e[7]))
                         : "memory");
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[t * 4 + 3]);
        }
        if (xkey >= 0) px = ex2_approx(fmaf(sx, scale_log2, -ms));
        l = l0 + l1 + px;
      } else {
        tc_fence_before();
        if (lane == 0) {
          for (int b = 0; b < n_full; ++b) mbar_arrive(&bar_p[t * 4 + b]);
          if (tail) mbar_arrive(&bar_p[t * 4 + 3]);
        }
      }
      PP_STAMP(4);
      mbar_wait(&bar_o[t], ph);
      PP_STAMP(5);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (active) {
        tmem_ld_32x32b_x32(trow, o0);
        tmem_ld_32x32b_x32(trow + 32, o1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tfree[t]);
      if (xkey >= 0 && active) {            // O_i += p_x v_x
        const uint4* vrow = reinterpret_cast<const uint4*>(qkv + (static_cast<size_t>(view) * tokens + xkey) * (3 * static_cast<size_t>(d)) +
                                                           2 * d + h * DH);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 vv = __ldg(vrow + c8);
          const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vv);
          uint32_t* o = c8 < 4 ? o0 + c8 * 8 : o1 + (c8 - 4) * 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fv = __bfloat1622float2(v2[j]);
            o[2 * j] = __float_as_uint(fmaf(px, fv.x, __uint_as_float(o[2 * j])));
            o[2 * j + 1] = __float_as_uint(fmaf(px, fv.y, __uint_as_float(o[2 * j + 1])));
          }
        }
      }
      if (active) {
        const float inv = 1.f / l;
        const uint32_t obase = smem_u32(ostage) + row * 128;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const uint32_t* r = q4 < 4 ? o0 + q4 * 8 : o1 + (q4 - 4) * 8;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obase + ((static_cast<uint32_t>(q4) ^ (row & 7)) << 4)),
                       "r"(pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv))
                       : "memory");
        }
        if (lse != nullptr && r0 + row < tokens)
          lse[(static_cast<size_t>(view) * heads + h) * tokens + r0 + row] = (m * scale_log2 + log2f(l)) * LN2;
      }
      fence_proxy_async_smem();
      named_bar_sync(3 + t, 128);
      if (wg_tid == 0 && r0 < tokens) {
        tma_store_3d(&tmOut, ostage, h * DH, r0, view);
        bulk_commit();
      }
      PP_STAMP(6);
    }
    if (wg_tid == 0) bulk_wait<0>();
  }
#undef PP_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================== tcgen05 forward, P kept in TMEM
// attention_fwd_pp_kernel with (1) the probabilities handed to the tensor core through TMEM instead of shared memory (the TS form
// of tcgen05.mma: A operand = 128 lanes x 8 columns of packed bf16 pairs per 16-key step) and (2) the two query tiles of a unit
// DECOUPLED: each tile t is a stream of its own -- softmax warpgroup, MMA-issuing thread, Q double buffer, half of TMEM -- and the
// streams only share the (double-buffered) K and V of a unit; every softmax WARP stores its own 32 output rows (own staging
// slice, own TMA store), so no CTA- or warpgroup-wide barrier is left in the unit loop.  193..208 tokens (keys = 208).
// Per tile t, TMEM columns [256 t, 256 t + 256):
//   S_t  [0, 208)       fp32 scores, written by the S MMAs
//   P_t  [0, 104)       bf16 pairs, written IN PLACE by the softmax threads: pass 2 walks S in 16-key steps and stores the 8 columns
//                       of packed P of step h at [8 + 8 h, 16 + 8 h), which only covers S columns the thread has already loaded;
//                       the 16-key tail's P goes to [0, 8)
//   O_t  [192, 256)     fp32 accumulator of P V.  It overlaps the 16-column tail of S, so pass 2 reads that tail FIRST; the first
//                       P V MMA is issued after all four warps of the tile delivered block 0, i.e. after every lane has read its tail.
// No swizzled shared-memory P stores, no proxy fences; the shared memory that P occupied holds the second K and Q buffers
// (200 KB).  Exact two-pass softmax in fp32; 2 of 8 exponentials on the FMA pipe (163.3 us at 576 views against 166.4 with 1 of 8
// and 168.9 with none); scale and row sums as packed fp32 pairs.
//   warp 0 (one thread)   TMA producer: K + Q0 + Q1 of unit u into buffer u & 1 once both streams' S MMAs of unit u - 2 retired;
//                         V likewise behind both streams' last P V MMA
//   warp 1 / warp 10      MMA issue of stream 0 / 1 (one elected thread each; tcgen05.commit tracks the issuing thread's MMAs)
//   warps 2-5 / 6-9       softmax + epilogue warps of stream 0 / 1
// Measured on the way: a softmax warp that is alone on its scheduler runs pass 2 at about half the speed of two
// (ptxas schedules for latency hiding by other warps: the static stall counts of the loop alone add up to more than the MUFU time),
// so letting the streams take turns in pass 2 ("stagger") gains nothing; polling warps are not what slows it (suspend hint and
// nanosleep back-off change nothing); neither do the tcgen05.ld / .st of the pass nor the P V MMAs.
constexpr int PT_THREADS = 352;
constexpr int PT_KEYS = 208, PT_KB = PT_KEYS * 128;      // bytes of K (or V) of a unit
template <int MASK, bool DBG>     // MASK: which of every 8 exponentials run on the FMA pipe; DBG: clock64 stamps of CTA 0
__global__ void __launch_bounds__(PT_THREADS, 1)
attention_fwd_pt_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmOut, float* __restrict__ lse, int tokens, int heads,
                        int units, float scale_log2, long long* __restrict__ dbg, int rev) {
  extern __shared__ uint8_t smem_pt_raw[];
#define PT_STAMP(k) do { if (DBG && blockIdx.x == 0 && it < 12 && quad == 0 && lane == 0) dbg[(it * 2 + t) * 16 + (k)] = clock64(); } while (0)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_pt_raw) + 1023) & ~uintptr_t(1023));
  constexpr int KB = PT_KB;
  uint8_t* sQ = smem;                              // [2 buffers][2 tiles] 16 KB each
  uint8_t* sK = sQ + 4 * 16384;                    // [2 buffers]
  uint8_t* sV = sK + 2 * KB;                       // [2 buffers]
  uint8_t* sO = sV + 2 * KB;                       // [2 tiles][4 warps] output stage, 4 KB (32 rows) each
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * 16384);
  uint64_t* bar_kq = bars;            // [2] K + Q0 + Q1 of a unit landed in buffer b
  uint64_t* bar_kqfree = bars + 2;    // [2] both streams' S MMAs on buffer b retired (2 commits)
  uint64_t* bar_v = bars + 4;         // [2] V buffer landed
  uint64_t* bar_vfree = bars + 6;     // [2] both streams' P V MMAs on the V buffer retired (2 commits)
  uint64_t* bar_s = bars + 8;         // [2] S_t complete in TMEM
  uint64_t* bar_o = bars + 10;        // [2] O_t complete in TMEM (P_t dead)
  uint64_t* bar_tfree = bars + 12;    // [2] O_t copied to registers: TMEM region t reusable
  uint64_t* bar_p = bars + 14;        // [2][4] P block b (four 16-key steps; block 0 also the tail) of tile t stored in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * DH;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_kq[i], 1);
      mbar_init(&bar_kqfree[i], 2);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_vfree[i], 2);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_tfree[i], 4);
      for (int b = 0; b < 4; ++b) mbar_init(&bar_p[i * 4 + b], 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const int u2 = rev ? units - 1 - unit : unit;            // descending walk (kernels.cuh)
        const int view = u2 / heads, h = u2 - view * heads;
        const int b = it & 1, use = it >> 1;
        if (use >= 1) mbar_wait(&bar_kqfree[b], (use - 1) & 1);  // both S MMAs of unit it - 2 retired
        mbar_expect_tx(&bar_kq[b], KB + 2 * 16384);
        tma_load_3d(&tmKV, &bar_kq[b], sK + b * KB, d + h * DH, 0, view);
        tma_load_3d(&tmQ, &bar_kq[b], sQ + b * 32768, h * DH, 0, view);
        tma_load_3d(&tmQ, &bar_kq[b], sQ + b * 32768 + 16384, h * DH, 128, view);
        if (use >= 1) mbar_wait(&bar_vfree[b], (use - 1) & 1);
        mbar_expect_tx(&bar_v[b], KB);
        tma_load_3d(&tmKV, &bar_v[b], sV + b * KB, 2 * d + h * DH, 0, view);
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------------------------ MMA issue of stream t (one elected thread, counted loops:
    // see attention_bwd_tc_kernel for the two ptxas pitfalls this avoids)
    const int t = warp == 1 ? 0 : 1;
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, PT_KEYS);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      const uint32_t tt = tmem + t * 256;
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int b = it & 1, use = it >> 1;
        const uint32_t qa = smem_u32(sQ + b * 32768 + t * 16384), ka = smem_u32(sK + b * KB);
        mbar_wait(&bar_kq[b], use & 1);
        if (it > 0) mbar_wait(&bar_tfree[t], ph ^ 1);             // the previous unit's O_t has left TMEM
        tc_fence_after();
        // (S split into keys [0, 192) issued right behind the previous unit's last P V MMA and the 16-key tail behind bar_tfree:
        //  S ready 200 cycles earlier, but the 384-cycle head then sits in front of the OTHER stream's last P V block in the one
        //  tensor pipe: 171.0 against 166.7 us at 576 views)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tt, umma_desc_k_sw128(qa + k * 32), umma_desc_k_sw128(ka + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[t]);
        umma_commit(&bar_kqfree[b]);
        mbar_wait(&bar_v[b], use & 1);
        const uint64_t vdesc = umma_desc(smem_u32(sV + b * KB), 1024, KB, 2);     // + 128 per 16-key step (2048 B >> 4)
#pragma unroll 1
        for (int blk = 0; blk < 3; ++blk) {
          mbar_wait(&bar_p[t * 4 + blk], ph);
          tc_fence_after();
          if (blk == 0) umma_bf16_ts(tt + 192, tt, vdesc + 12ull * 128, idesc_o, 0u);      // the tail first: O starts from it
#pragma unroll 1
          for (int ks = 4 * blk; ks < 4 * blk + 4; ++ks)
            umma_bf16_ts(tt + 192, tt + 8 + 8 * ks, vdesc + static_cast<uint64_t>(ks) * 128, idesc_o, 1u);
        }
        umma_commit(&bar_o[t]);
        umma_commit(&bar_vfree[b]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps of stream t
    const int t = (warp - 2) >> 2;
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16) + t * 256;
    uint8_t* ostage = sO + t * 16384 + quad * 4096;         // this warp's 32 rows x 128 B (1024-aligned: swizzle atoms intact)
    const int r0 = t * 128 + quad * 32;                     // first query row of this warp
    const bool active = r0 < tokens;                        // warp-uniform: at least one real query row
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int u2 = rev ? units - 1 - unit : unit;
      const int view = u2 / heads, h = u2 - view * heads;
      PT_STAMP(0);
      mbar_wait(&bar_s[t], ph);
      PT_STAMP(1);
      tc_fence_after();
      float m = -INFINITY, l = 0.f;
      if (active) {
        // pass 1: exact row maximum.  Columns >= tokens hold exact zeros (zero-filled K rows): harmless for softmax.
        uint32_t ra[32], rb[32], rt[16];
        tmem_ld_32x32b_x16(trow + 192, rt);
        tmem_ld_32x32b_x32(trow, ra);
        tmem_ld_32x32b_x32(trow + 32, rb);
        tmem_ld_wait();
        float m1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m = max3(m, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
          m1 = max3(m1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
        }
        tmem_ld_32x32b_x32(trow + 64, ra);
        tmem_ld_32x32b_x32(trow + 96, rb);
#pragma unroll
        for (int i = 0; i < 16; i += 2) m = max3(m, __uint_as_float(rt[i]), __uint_as_float(rt[i + 1]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m = max3(m, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
          m1 = max3(m1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
        }
        tmem_ld_32x32b_x32(trow + 128, ra);
        tmem_ld_32x32b_x32(trow + 160, rb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m = max3(m, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
          m1 = max3(m1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
        }
        m = fmaxf(m, m1);
      }
      PT_STAMP(2);
      if (active) {
        // pass 2: p = 2^(s*scale - m*scale), row sum, packed bf16 P back into TMEM (layout above).  16-key steps through two
        // 16-register buffers: the load of step h + 1 flies while step h is evaluated.
        const float ms = m * scale_log2;
        const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2), nms2 = f32x2_pack(-ms, -ms);
        uint64_t l2 = 0;
        uint32_t ba[16], bb[16], pk[8];
        auto eval16 = [&](const uint32_t (&r)[16]) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint64_t e2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint64_t x2 = f32x2_fma(f32x2_pack_bits(r[8 * q + 2 * i], r[8 * q + 2 * i + 1]), sc2, nms2);
              const float e0 = ((MASK >> (2 * i)) & 1) ? ex2_poly(f32x2_lo(x2)) : ex2_approx(f32x2_lo(x2));
              const float e1 = ((MASK >> (2 * i + 1)) & 1) ? ex2_poly(f32x2_hi(x2)) : ex2_approx(f32x2_hi(x2));
              e2[i] = f32x2_pack(e0, e1);
              pk[4 * q + i] = pack_bf16(e0, e1);
            }
            l2 = f32x2_add(l2, f32x2_add(f32x2_add(e2[0], e2[1]), f32x2_add(e2[2], e2[3])));
          }
        };
        tmem_ld_32x32b_x16(trow + 192, ba);
        tmem_ld_32x32b_x16(trow, bb);
        tmem_ld_wait();
        float lt = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v0 = ex2_approx(fmaf(__uint_as_float(ba[2 * i]), scale_log2, -ms));
          float v1 = ex2_approx(fmaf(__uint_as_float(ba[2 * i + 1]), scale_log2, -ms));
          if (192 + 2 * i >= tokens) v0 = 0.f;
          if (192 + 2 * i + 1 >= tokens) v1 = 0.f;
          lt += v0 + v1;
          pk[i] = pack_bf16(v0, v1);
        }
        tmem_ld_32x32b_x16(trow + 16, ba);
        tmem_st_32x32b_x8(trow, pk);               // S columns [0, 8) are in bb already
        auto deliver = [&](int blk) {              // P block blk (four steps) is in TMEM: tell the MMA thread
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[t * 4 + blk]);
          PT_STAMP(8 + blk);
        };
#pragma unroll 1
        for (int i = 0; i < 6; ++i) {              // 16-key steps 2 i (in bb) and 2 i + 1 (in ba, in flight)
          eval16(bb);
          // a block's delivery rides one evaluation behind its last store: tcgen05.wait::st then finds the stores retired
          // (delivered right behind the store it cost ~150 cycles of the pass per block)
          if (i == 2 || i == 4) deliver((i >> 1) - 1);
          tmem_ld_wait();
          if (i < 5) tmem_ld_32x32b_x16(trow + 32 * i + 32, bb);
          tmem_st_32x32b_x8(trow + 8 + 16 * i, pk);
          eval16(ba);
          tmem_ld_wait();
          if (i < 5) tmem_ld_32x32b_x16(trow + 32 * i + 48, ba);
          tmem_st_32x32b_x8(trow + 16 + 16 * i, pk);
        }
        deliver(2);
        l = lt + f32x2_lo(l2) + f32x2_hi(l2);
      } else {
        tc_fence_before();
        if (lane == 0)
          for (int b = 0; b < 3; ++b) mbar_arrive(&bar_p[t * 4 + b]);
      }
      PT_STAMP(4);
      mbar_wait(&bar_o[t], ph);
      PT_STAMP(5);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      if (active) {
        tmem_ld_32x32b_x32(trow + 192, o0);
        tmem_ld_32x32b_x32(trow + 224, o1);
        if (lane == 0) bulk_wait_read<0>();      // this warp's previous output store has finished reading its stage
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tfree[t]);
      if (active) {
        const float inv = 1.f / l;
        const uint32_t obase = smem_u32(ostage) + lane * 128;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const uint32_t* r = q4 < 4 ? o0 + q4 * 8 : o1 + (q4 - 4) * 8;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obase + ((static_cast<uint32_t>(q4) ^ (lane & 7)) << 4)),
                       "r"(pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv)),
                       "r"(pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv))
                       : "memory");
        }
        if (lse != nullptr && r0 + lane < tokens)
          lse[(static_cast<size_t>(view) * heads + h) * tokens + r0 + lane] = (m * scale_log2 + log2f(l)) * LN2;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmOut, ostage, h * DH, r0, view);      // 32-row box; rows >= tokens are clipped
          bulk_commit();
        }
      }
      PT_STAMP(6);
    }
    if (lane == 0) bulk_wait<0>();
    (void)row;
  }
#undef PT_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================== tcgen05 forward, P kept in TMEM, 257 tokens
// attention_fwd_pt_kernel for ViT-L/14's 257 tokens: 256 keys and 256 queries go through the MMAs (two full 128-row query tiles,
// S_t = 128 x 256 fp32 fills the tile's 256 TMEM columns), key 256 is folded in by the softmax threads (score from global memory
// while the S MMAs run, p_256 v_256 added to O in the epilogue) and query row 256 is computed by one extra warp on the CUDA cores
// from the K / V tiles the unit has in shared memory anyway -- as in attention_fwd_pp_kernel<true>, which this kernel replaces.
// Per tile t, TMEM columns [256 t, 256 t + 256):
//   S_t   [0, 256)
//   O_t   [192, 256)    overlaps the scores of keys 192..255: pass 2 evaluates those four 16-key steps FIRST and keeps their packed
//                       probabilities in registers
//   P_t   keys 128..255 at [128, 192) (step ks = 8..15 at column 64 + 8 ks): written while / after the steps 8..11 are evaluated,
//                       i.e. only over S columns the thread has already loaded; block 0 (these eight steps) starts O
//         keys 0..127   at [0, 64) (step ks at column 8 ks), in place behind the loads as in the 208-key kernel: blocks 1 and 2
// The q / k / v rows of token 256 ride with the unit's K / V loads (three 128-byte TMA boxes per buffer, no swizzle), so no compute
// warp touches global memory on its critical path (first version: __ldg of those rows in the softmax warps and the extra warp --
// 609 us at 576 views x 16 heads, 264 us with both switched off, 573 us for attention_fwd_pp_kernel<true>).
// Shared memory: Q 2 x 2 x 16 KB, K and V double-buffered (4 x 32 KB), per-warp output stages 32 KB: 224 KB (+ 1 KB: barriers and
// the three rows).
constexpr int PX_THREADS = PT_THREADS + 32;
constexpr int PX_KB = 256 * 128;      // bytes of K (or V) of a unit
template <int MASK>
__global__ void __launch_bounds__(PX_THREADS, 1)
attention_fwd_px_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRow, float* __restrict__ lse, int heads, int units, float scale_log2,
                        int rev, const bf16* __restrict__ qkv, bf16* __restrict__ out, int dbg) {
  extern __shared__ uint8_t smem_px_raw[];
  constexpr int tokens = 257, XKEY = 256, KB = PX_KB;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_px_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2 buffers][2 tiles] 16 KB each
  uint8_t* sK = sQ + 4 * 16384;                    // [2 buffers]
  uint8_t* sV = sK + 2 * KB;                       // [2 buffers]
  uint8_t* sO = sV + 2 * KB;                       // [2 tiles][4 warps] output stage, 4 KB (32 rows) each
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * 16384);
  uint64_t* bar_kq = bars;            // [2] K + Q0 + Q1 of a unit landed in buffer b
  uint64_t* bar_kqfree = bars + 2;    // [2] both streams' S MMAs on buffer b retired (2 commits)
  uint64_t* bar_v = bars + 4;         // [2] V buffer landed
  uint64_t* bar_vfree = bars + 6;     // [2] both streams' P V MMAs on the V buffer retired (2 commits)
  uint64_t* bar_s = bars + 8;         // [2] S_t complete in TMEM
  uint64_t* bar_o = bars + 10;        // [2] O_t complete in TMEM (P_t dead)
  uint64_t* bar_tfree = bars + 12;    // [2] O_t copied to registers: TMEM region t reusable
  uint64_t* bar_p = bars + 14;        // [2][4] P block b of tile t stored in TMEM (three blocks used)
  uint64_t* bar_xk = bars + 22;       // [2] the extra-query warp has read K buffer b, every softmax warp its q row and k_256 (9 arrivals)
  uint64_t* bar_xv = bars + 24;       // [2] the extra-query warp has read V buffer b
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  uint8_t* xrow = reinterpret_cast<uint8_t*>(bars + 32);     // [2 buffers][q_256 | k_256 | v_256] 128 B each, not swizzled

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * DH;
  const size_t ld = static_cast<size_t>(3) * d;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRow);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_kq[i], 1);
      mbar_init(&bar_kqfree[i], 2);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_vfree[i], 2);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_o[i], 1);
      mbar_init(&bar_tfree[i], 4);
      mbar_init(&bar_xk[i], 9);
      mbar_init(&bar_xv[i], 1);
      for (int b = 0; b < 4; ++b) mbar_init(&bar_p[i * 4 + b], 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const int u2 = rev ? units - 1 - unit : unit;
        const int view = u2 / heads, h = u2 - view * heads;
        const int b = it & 1, use = it >> 1;
        if (use >= 1) {
          mbar_wait(&bar_kqfree[b], (use - 1) & 1);            // both S MMAs of unit it - 2 retired
          mbar_wait(&bar_xk[b], (use - 1) & 1);                // ... and the extra-query warp is done with that K
        }
        mbar_expect_tx(&bar_kq[b], KB + 2 * 16384 + 256);
        tma_load_3d(&tmRow, &bar_kq[b], xrow + b * 384, h * DH, XKEY, view);
        tma_load_3d(&tmRow, &bar_kq[b], xrow + b * 384 + 128, d + h * DH, XKEY, view);
        tma_load_3d(&tmKV, &bar_kq[b], sK + b * KB, d + h * DH, 0, view);
        tma_load_3d(&tmQ, &bar_kq[b], sQ + b * 32768, h * DH, 0, view);
        tma_load_3d(&tmQ, &bar_kq[b], sQ + b * 32768 + 16384, h * DH, 128, view);
        if (use >= 1) {
          mbar_wait(&bar_vfree[b], (use - 1) & 1);
          mbar_wait(&bar_xv[b], (use - 1) & 1);
        }
        mbar_expect_tx(&bar_v[b], KB + 128);
        tma_load_3d(&tmRow, &bar_v[b], xrow + b * 384 + 256, 2 * d + h * DH, XKEY, view);
        tma_load_3d(&tmKV, &bar_v[b], sV + b * KB, 2 * d + h * DH, 0, view);
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------------------------------ MMA issue of stream t (one elected thread, counted loops)
    const int t = warp == 1 ? 0 : 1;
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 256);
      const uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      const uint32_t tt = tmem + t * 256;
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int b = it & 1, use = it >> 1;
        const uint32_t qa = smem_u32(sQ + b * 32768 + t * 16384), ka = smem_u32(sK + b * KB);
        mbar_wait(&bar_kq[b], use & 1);
        if (it > 0) mbar_wait(&bar_tfree[t], ph ^ 1);             // the previous unit's O_t has left TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tt, umma_desc_k_sw128(qa + k * 32), umma_desc_k_sw128(ka + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[t]);
        umma_commit(&bar_kqfree[b]);
        mbar_wait(&bar_v[b], use & 1);
        const uint64_t vdesc = umma_desc(smem_u32(sV + b * KB), 1024, KB, 2);     // + 128 per 16-key step (2048 B >> 4)
        mbar_wait(&bar_p[t * 4 + 0], ph);                         // block 0: keys 128..255, starts O
        tc_fence_after();
#pragma unroll 1
        for (int ks = 8; ks < 16; ++ks)
          umma_bf16_ts(tt + 192, tt + 64 + 8 * ks, vdesc + static_cast<uint64_t>(ks) * 128, idesc_o, ks != 8 ? 1u : 0u);
#pragma unroll 1
        for (int blk = 1; blk < 3; ++blk) {
          mbar_wait(&bar_p[t * 4 + blk], ph);
          tc_fence_after();
#pragma unroll 1
          for (int ks = 4 * blk - 4; ks < 4 * blk; ++ks)
            umma_bf16_ts(tt + 192, tt + 8 * ks, vdesc + static_cast<uint64_t>(ks) * 128, idesc_o, 1u);
        }
        umma_commit(&bar_o[t]);
        umma_commit(&bar_vfree[b]);
      }
    }
    __syncwarp();
  } else if (warp == 11) {
    // ------------------------------------------------------------------ query row 256: one warp, mma.sync (m16n8k16) on the K / V
    // tiles in shared memory (first version: CUDA cores, lane = key for the scores and lane = two output dimensions for P V, 2600
    // instructions per unit on a scheduler it shares with two softmax warps -- 464 us at 576 views, 319 us with this warp idle)
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const int b = it & 1, use = it >> 1;
      const int u2 = rev ? units - 1 - unit : unit;
      const int view = u2 / heads, h = u2 - view * heads;
      const uint8_t* xr = xrow + b * 384;
      mbar_wait(&bar_kq[b], use & 1);            // K of the unit and the rows of token 256 have landed
      if (dbg & 1) {      // development: no work, only the hand-overs
        if (lane == 0) mbar_arrive(&bar_xk[b]);
        mbar_wait(&bar_v[b], use & 1);
        if (lane == 0) mbar_arrive(&bar_xv[b]);
        continue;
      }
      // q_256 as row 0 of an m16 A tile (rows 1..15 zero): lanes 0..3 hold its words, S = q K^T and P V run on mma.sync over four
      // 64-key chunks with an online softmax (exp2 domain); only the accumulators of row lane / 4 = 0 (c0, c1) mean anything
      const uint32_t* qw = reinterpret_cast<const uint32_t*>(xr);
      uint32_t aq[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        aq[ks][0] = lane < 4 ? qw[ks * 8 + lane] : 0u;
        aq[ks][2] = lane < 4 ? qw[ks * 8 + 4 + lane] : 0u;
        aq[ks][1] = 0u;
        aq[ks][3] = 0u;
      }
      float sxx;
      {
        const float2 fq = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(xr)[lane]);
        const float2 fk = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(xr + 128)[lane]);
        sxx = warp_sum(fmaf(fq.x, fk.x, fq.y * fk.y));
      }
      const uint32_t sKa = smem_u32(sK + b * KB), sVa = smem_u32(sV + b * KB);
      float o[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
      float mrow = -INFINITY, lrow = 0.f;
#pragma unroll 1
      for (int kc = 0; kc < 256; kc += 64) {
        float sa[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) sa[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t bf[4];
            ldb_nk_sw(sKa, kc + np * 16, ks * 16, lane, bf);
            mma_bf16_16816(sa[2 * np], aq[ks], bf[0], bf[1]);
            mma_bf16_16816(sa[2 * np + 1], aq[ks], bf[2], bf[3]);
          }
        }
        if (kc == 192) {                           // last read of K
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_xk[b]);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          sa[nt][0] *= scale_log2;
          sa[nt][1] *= scale_log2;
          mx = max3(mx, sa[nt][0], sa[nt][1]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mn = fmaxf(mrow, mx);
        const float c = ex2_approx(mrow - mn);     // first chunk: 2^-inf = 0
        mrow = mn;
        lrow *= c;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          o[nt][0] *= c;
          o[nt][1] *= c;
          sa[nt][0] = ex2_approx(sa[nt][0] - mn);
          sa[nt][1] = ex2_approx(sa[nt][1] - mn);
          lrow += sa[nt][0] + sa[nt][1];
        }
        if (kc == 0) mbar_wait(&bar_v[b], use & 1);      // V of the unit
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t pf[4];
          pf[0] = pack_bf16(sa[2 * kk][0], sa[2 * kk][1]);
          pf[2] = pack_bf16(sa[2 * kk + 1][0], sa[2 * kk + 1][1]);
          pf[1] = 0u;
          pf[3] = 0u;
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t bf[4];
            ldb_kn_sw(sVa, kc + kk * 16, np * 16, lane, bf);
            mma_bf16_16816(o[2 * np], pf, bf[0], bf[1]);
            mma_bf16_16816(o[2 * np + 1], pf, bf[2], bf[3]);
          }
        }
      }
      // key 256, then the row sum across the four lanes of row 0
      const float sx2 = sxx * scale_log2;
      const float mn = fmaxf(mrow, sx2);
      const float c = ex2_approx(mrow - mn), pxx = ex2_approx(sx2 - mn);
      float l = lrow * c;
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      l += pxx;
      const float inv = 1.f / l;
      if (lane < 4) {
        bf16* orow = out + (static_cast<size_t>(view) * tokens + XKEY) * d + h * DH;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float2 vx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + 256 + (nt * 8 + 2 * lane) * 2));
          *reinterpret_cast<__nv_bfloat162*>(orow + nt * 8 + 2 * lane) =
              __floats2bfloat162_rn(fmaf(pxx, vx.x, o[nt][0] * c) * inv, fmaf(pxx, vx.y, o[nt][1] * c) * inv);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_xv[b]);      // V (and v_256) may be overwritten
      if (lane == 0 && lse != nullptr) lse[(static_cast<size_t>(view) * heads + h) * tokens + XKEY] = (mn + log2f(l)) * LN2;
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps of stream t (every query row is real)
    const int t = (warp - 2) >> 2;
    const int quad = warp & 3;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16) + t * 256;
    uint8_t* ostage = sO + t * 16384 + quad * 4096;         // this warp's 32 rows x 128 B (1024-aligned: swizzle atoms intact)
    const int r0 = t * 128 + quad * 32;                     // first query row of this warp
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int u2 = rev ? units - 1 - unit : unit;
      const int view = u2 / heads, h = u2 - view * heads;
      // key 256: its score q_i . k_256 while the S MMAs run, q row and k_256 from shared memory (the Q tile is swizzled)
      const int b = it & 1, use = it >> 1;
      float sx = 0.f;
      mbar_wait(&bar_kq[b], use & 1);
      if (!(dbg & 2)) {
        const int row = quad * 32 + lane;
        const uint32_t qa = smem_u32(sQ + b * 32768 + t * 16384) + row * 128;
        const uint8_t* kr = xrow + b * 384 + 128;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint4 a;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                       : "r"(qa + ((static_cast<uint32_t>(c8) ^ (row & 7)) << 4)));
          const uint4 kx = *reinterpret_cast<const uint4*>(kr + c8 * 16);
          const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kx);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(a2[j]), fb = __bfloat1622float2(k2[j]);
            sx = fmaf(fa.x, fb.x, sx);
            sx = fmaf(fa.y, fb.y, sx);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_xk[b]);    // this warp is done with the Q tile and k_256 of buffer b
      mbar_wait(&bar_s[t], ph);
      tc_fence_after();
      float m = sx;
      {
        // pass 1: exact row maximum over the 256 scores in TMEM (two register buffers, loads one pair ahead)
        uint32_t ra[32], rb[32];
        float m1 = -INFINITY;
        tmem_ld_32x32b_x32(trow, ra);
        tmem_ld_32x32b_x32(trow + 32, rb);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) m = max3(m, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
          if (c < 3) tmem_ld_32x32b_x32(trow + 64 * c + 64, ra);
#pragma unroll
          for (int i = 0; i < 32; i += 2) m1 = max3(m1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
          if (c < 3) tmem_ld_32x32b_x32(trow + 64 * c + 96, rb);
        }
        m = fmaxf(m, m1);
      }
      float l;
      uint4 vx[8];
      const float ms = m * scale_log2;
      const float px = ex2_approx(fmaf(sx, scale_log2, -ms));
      {
        // pass 2: p = 2^(s*scale - m*scale), row sum, packed bf16 P back into TMEM (layout above).  16-key steps through two
        // 16-register buffers: the load of the next step flies while a step is evaluated.
        const uint64_t sc2 = f32x2_pack(scale_log2, scale_log2), nms2 = f32x2_pack(-ms, -ms);
        uint64_t l2 = 0;
        uint32_t ba[16], bb[16], pk[8], pa0[8], pa1[8], pa2[8], pa3[8];
        auto eval16 = [&](const uint32_t (&r)[16], uint32_t (&o)[8]) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint64_t e2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint64_t x2 = f32x2_fma(f32x2_pack_bits(r[8 * q + 2 * i], r[8 * q + 2 * i + 1]), sc2, nms2);
              const float e0 = ((MASK >> (2 * i)) & 1) ? ex2_poly(f32x2_lo(x2)) : ex2_approx(f32x2_lo(x2));
              const float e1 = ((MASK >> (2 * i + 1)) & 1) ? ex2_poly(f32x2_hi(x2)) : ex2_approx(f32x2_hi(x2));
              e2[i] = f32x2_pack(e0, e1);
              o[4 * q + i] = pack_bf16(e0, e1);
            }
            l2 = f32x2_add(l2, f32x2_add(f32x2_add(e2[0], e2[1]), f32x2_add(e2[2], e2[3])));
          }
        };
        auto deliver = [&](int blk) {              // P block blk is in TMEM: tell the MMA thread
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[t * 4 + blk]);
        };
        // keys 192..255 (their S columns are O's home): evaluated first, kept in registers
        tmem_ld_32x32b_x16(trow + 192, ba);
        tmem_ld_32x32b_x16(trow + 208, bb);
        tmem_ld_wait();
        eval16(ba, pa0);
        tmem_ld_32x32b_x16(trow + 224, ba);
        eval16(bb, pa1);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + 240, bb);
        eval16(ba, pa2);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + 128, ba);
        eval16(bb, pa3);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + 144, bb);
        // keys 128..191: step ks = 8..11 -> columns [128 + 8 (ks - 8), + 8), inside S columns this thread has loaded
        eval16(ba, pk);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + 160, ba);
        tmem_st_32x32b_x8(trow + 128, pk);
        eval16(bb, pk);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + 176, bb);
        tmem_st_32x32b_x8(trow + 136, pk);
        eval16(ba, pk);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow, ba);
        tmem_st_32x32b_x8(trow + 144, pk);
        eval16(bb, pk);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(trow + 16, bb);
        tmem_st_32x32b_x8(trow + 152, pk);
        tmem_st_32x32b_x8(trow + 160, pa0);        // keys 192..255 behind them: S columns 160..191 are in registers / done
        tmem_st_32x32b_x8(trow + 168, pa1);
        tmem_st_32x32b_x8(trow + 176, pa2);
        tmem_st_32x32b_x8(trow + 184, pa3);
        deliver(0);
        // keys 0..127: step ks -> columns [8 ks, 8 ks + 8)
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {              // steps 2 i (in ba) and 2 i + 1 (in bb, landed by the wait below)
          eval16(ba, pk);
          tmem_ld_wait();
          if (i < 3) tmem_ld_32x32b_x16(trow + 32 * i + 32, ba);
          tmem_st_32x32b_x8(trow + 16 * i, pk);
          eval16(bb, pk);
          tmem_ld_wait();
          if (i < 3) tmem_ld_32x32b_x16(trow + 32 * i + 48, bb);
          tmem_st_32x32b_x8(trow + 16 * i + 8, pk);
          if (i == 1) deliver(1);
        }
        // v_256 (lands with V) into registers BEFORE the last block is handed over: the buffer cannot be reloaded until then
        mbar_wait(&bar_v[b], use & 1);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) vx[c8] = *reinterpret_cast<const uint4*>(xrow + b * 384 + 256 + c8 * 16);
        deliver(2);
        l = f32x2_lo(l2) + f32x2_hi(l2) + px;
      }
      mbar_wait(&bar_o[t], ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(trow + 192, o0);
      tmem_ld_32x32b_x32(trow + 224, o1);
      if (lane == 0) bulk_wait_read<0>();      // this warp's previous output store has finished reading its stage
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tfree[t]);
      if (!(dbg & 2)) {                        // O_i += p_256 v_256
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vx[c8]);
          uint32_t* o = c8 < 4 ? o0 + c8 * 8 : o1 + (c8 - 4) * 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fv = __bfloat1622float2(v2[j]);
            o[2 * j] = __float_as_uint(fmaf(px, fv.x, __uint_as_float(o[2 * j])));
            o[2 * j + 1] = __float_as_uint(fmaf(px, fv.y, __uint_as_float(o[2 * j + 1])));
          }
        }
      }
      const float inv = 1.f / l;
      const uint32_t obase = smem_u32(ostage) + lane * 128;
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const uint32_t* r = q4 < 4 ? o0 + q4 * 8 : o1 + (q4 - 4) * 8;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obase + ((static_cast<uint32_t>(q4) ^ (lane & 7)) << 4)),
                     "r"(pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv)),
                     "r"(pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv)),
                     "r"(pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv)),
                     "r"(pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv))
                     : "memory");
      }
      if (lse != nullptr) lse[(static_cast<size_t>(view) * heads + h) * tokens + r0 + lane] = (ms + log2f(l)) * LN2;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmOut, ostage, h * DH, r0, view);      // 32-row box, rows r0 .. r0 + 31 < 256
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================== CLS-query attention
// Last encoder layer in inference: only the CLS token of each view feeds post_layernorm / visual_projection
// (HF CLIPVisionTransformer.forward: pooled_output = last_hidden_state[:, 0]), so only its query row is needed.
// One CTA per (head, view): phase 1 one thread per key (q . k_j, fp32), block softmax, phase 2 one thread per output
// dimension (sum_j p_j v_j[d], coalesced over d).  q comes from a compact [V, d] tensor, K/V from the usual qkv rows.
__global__ void __launch_bounds__(128)
attention_cls_kernel(const bf16* __restrict__ q_cls, size_t q_view_stride, const bf16* __restrict__ qkv, bf16* __restrict__ out_cls,
                     size_t out_view_stride, float* __restrict__ lse, int lse_row, int tokens, int heads, float scale_log2) {
  // General single-query-row form: q row of view v at q_cls + v * q_view_stride, output row at out_cls + v * out_view_stride
  // (compact [V, d] tensors for the CLS path; the last token row of qkv / out for the 257-token path, which also wants lse).
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh_cls[];
  float* sq = sh_cls;            // [64]
  float* sp = sh_cls + 64;       // [tokens]
  __shared__ float red[8];
  const int h = blockIdx.x, view = blockIdx.y, d = heads * DH, ld = 3 * d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < DH) sq[tid] = __bfloat162float(q_cls[static_cast<size_t>(view) * q_view_stride + h * DH + tid]);
  __syncthreads();
  const bf16* kbase = qkv + static_cast<size_t>(view) * tokens * ld + d + h * DH;
  float mx = -INFINITY;
  for (int j = tid; j < tokens; j += blockDim.x) {
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(j) * ld);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = kr[c];
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(p2[e]);
        acc += f.x * sq[c * 8 + 2 * e] + f.y * sq[c * 8 + 2 * e + 1];
      }
    }
    acc *= scale_log2;
    sp[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float sum = 0.f;
  for (int j = tid; j < tokens; j += blockDim.x) {
    const float e = exp2f(sp[j] - mx);
    sp[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[4 + warp] = sum;
  __syncthreads();
  const float tot = red[4] + red[5] + red[6] + red[7];
  const float inv = 1.f / tot;
  if (lse != nullptr && tid == 0) lse[(static_cast<size_t>(view) * heads + h) * tokens + lse_row] = (mx + log2f(tot)) * LN2;
  // phase 2: threads 0..63 take the even keys, 64..127 the odd keys of output dimension tid & 63
  const int dd = tid & 63, half = tid >> 6;
  const bf16* vbase = qkv + static_cast<size_t>(view) * tokens * ld + 2 * d + h * DH + dd;
  float o = 0.f;
  for (int j = half; j < tokens; j += 2) o += sp[j] * __bfloat162float(vbase[static_cast<size_t>(j) * ld]);
  __syncthreads();
  if (half == 1) sq[dd] = o;
  __syncthreads();
  if (half == 0) out_cls[static_cast<size_t>(view) * out_view_stride + h * DH + dd] = __float2bfloat16((o + sq[dd]) * inv);
}

// =============================================================================================== tcgen05 backward
// Attention backward on the 5th-gen tensor cores, one persistent CTA per SM walking (view, head) units of 129..208 tokens.
// All five contractions of a unit run as tcgen05.mma with fp32 accumulators in TMEM; nothing of size tokens x tokens leaves
// the SM.  With P = exp(scale S - lse), dP = dO V^T, Delta_q = sum_d dO_qd O_qd, dS = scale P o (dP - Delta):
//     dQ = dS K        dK = dS^T Q        dV = P^T dO
// The unit is walked as (key half kh) x (query tile t): keys [0,128) then [128, keys), queries [0,128) then [128,256)
// (rows >= tokens are zero-filled by the TMA).  TMEM (all 512 columns):
//     [  0,128) S = Q_t K_kh^T      [128,256) dP = dO_t V_kh^T        (overwritten every sub-step)
//     [256,320) dQ_0   [320,384) dQ_1   (accumulate over both key halves)
//     [384,448) dV_kh  [448,512) dK_kh  (accumulate over both query tiles, drained after t = 1)
// Shared memory: K, V (keys x 64), both Q tiles, both dO tiles (TMA, 128B swizzle), and one P and one dS tile
// [128 queries][128 keys] bf16 written by the softmax threads as two 64-key blocks in the UMMA K-major 128B-swizzle layout.
// That same tile is the A operand twice: K-major for dQ = dS K (M = queries, K = keys) and -- through an MN-major
// descriptor (instruction-descriptor bit 15) -- transposed for dV = P^T dO and dK = dS^T Q (M = keys, K = queries): the 8-row
// x 128-byte swizzle atom is the same physical layout under both readings, so no transposed copy is ever written.
// K / Q_t / dO_t are B operands as they sit in memory (MN-major, like V in the forward's P V).
//   warp 0 (one thread): TMA loads, all MMAs, commits.
//   warps 1-8: two threads per query row (each takes 64 of the 128 key columns of a sub-step): tcgen05.ld of S and dP,
//              P / dS in fp32, bf16 tiles to smem; then the drains: dV / dK rows (one key per TMEM lane) and dQ rows.
// Keys >= tokens need no mask: their K / V rows are zero, so S = dP = 0, P and dS stay finite, and every product they enter
// is with a zero K row (dQ) or lands in a dK / dV row that is not stored.
constexpr int BT_THREADS = 288;

// 8 scores + 8 dP values -> 4 packed bf16 pairs of P and of dS (registers; written to smem once the tiles are free)
__device__ __forceinline__ void bt_group8(const uint32_t* s8, const uint32_t* dp8, float scale_log2, float lse2, float delta,
                                          float scale, uint32_t* p4, uint32_t* ds4) {
  float p[8], ds[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i] = ex2_approx(fmaf(__uint_as_float(s8[i]), scale_log2, -lse2));
    ds[i] = p[i] * (__uint_as_float(dp8[i]) - delta) * scale;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p4[i] = pack_bf16(p[2 * i], p[2 * i + 1]);
    ds4[i] = pack_bf16(ds[2 * i], ds[2 * i + 1]);
  }
}
__device__ __forceinline__ void bt_st16(uint32_t addr, const uint32_t* v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// 64 fp32 accumulator columns of this thread's TMEM lane -> one bf16 row of the 128B-swizzled staging tile [128][64]
__device__ __forceinline__ void bt_stage_row(const uint32_t (&a)[32], const uint32_t (&b)[32], uint32_t row_addr, uint32_t sw) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t* r = q < 4 ? a + q * 8 : b + (q - 4) * 8;
    const uint32_t v[4] = {pack_bf16(__uint_as_float(r[0]), __uint_as_float(r[1])), pack_bf16(__uint_as_float(r[2]), __uint_as_float(r[3])),
                           pack_bf16(__uint_as_float(r[4]), __uint_as_float(r[5])), pack_bf16(__uint_as_float(r[6]), __uint_as_float(r[7]))};
    bt_st16(row_addr + ((static_cast<uint32_t>(q) ^ sw) << 4), v);
  }
}

__global__ void __launch_bounds__(BT_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmO,
                        const __grid_constant__ CUtensorMap tmDst, const float* __restrict__ lse, int tokens, int heads, int units,
                        int keys, float scale, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_bt_raw[];
  // development aid (TTL_ATTN_DBG): clock64 stamps of the first units of CTA 0; slots 0..31 = MMA thread, 32..63 = softmax warp 1
#define BT_STAMP(k) do { if (dbg != nullptr && blockIdx.x == 0 && local_u < 6) dbg[local_u * 64 + (k)] = clock64(); } while (0)
  int local_u = 0;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_bt_raw) + 1023) & ~uintptr_t(1023));
  const int KB = keys * 128;                     // bytes of K (or V): keys rows of 128 bytes, a multiple of 2048
  uint8_t* sK = smem;
  uint8_t* sV = sK + KB;
  uint8_t* sQ = sV + KB;                         // 2 x 16 KB  [128 q][64 dh]
  uint8_t* sdO = sQ + 32768;                     // 2 x 16 KB
  uint8_t* sP = sdO + 32768;                     // 2 x 16 KB  blocks of 64 keys: [128 q][64 keys]
  uint8_t* sdS = sP + 32768;                     // 2 x 16 KB
  uint8_t* sOut = sdS + 32768;                   // 2 x 16 KB  output staging, one tile per softmax warpgroup
  float* sDelta = reinterpret_cast<float*>(sOut + 32768);   // [256] Delta_q of the unit
  float* sLse2 = sDelta + 256;                              // [256] lse_q * log2(e), +inf for rows >= tokens
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLse2 + 256);
  uint64_t* bar_load0 = bars;       // K, V, Q0, dO0 of the unit landed
  uint64_t* bar_load1 = bars + 1;   // Q1, dO1 landed
  uint64_t* bar_sdp = bars + 2;     // S and dP of a sub-step complete in TMEM
  uint64_t* bar_pds = bars + 3;     // P and dS tiles of a sub-step written, its S / dP drained (8 softmax warps)
  uint64_t* bar_s2 = bars + 4;      // dQ / dV / dK MMAs of a sub-step complete: the P / dS tiles may be overwritten
  uint64_t* bar_kv = bars + 5;      // every MMA of the unit has completed (one phase per unit)
  uint64_t* bar_drn = bars + 6;     // S / dP of a sub-step have been read out of TMEM (8 softmax warps): the next S / dP may be issued
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = heads * DH;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmDst);
    mbar_init(bar_load0, 1);
    mbar_init(bar_load1, 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 8);
    mbar_init(bar_s2, 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_drn, 8);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();
  pdl_trigger();
  const int nh1 = keys - 128;                                   // keys of the second half (16..80)
  const int nq1 = (tokens - 128 + 15) / 16;                     // 16-row k-steps that hold real queries in tile 1
  const float scale_log2 = scale * LOG2E;

  if (warp == 0) {
    // One elected lane runs the whole loop (elect.sync, not `lane == 0`: ptxas then issues the tcgen05.mma of a case back to
    // back instead of wrapping each in an ELECT / BRA.U.ANY retry loop).  The other form -- the whole warp in the loop, each
    // instruction behind its own elect.sync -- was measured slower (644 vs 444 us at 576 views: ~130 issue cycles per MMA).
    if (elect_one()) {
      const uint32_t idesc_s0 = umma_idesc_bf16(128, 128), idesc_s1 = umma_idesc_bf16(128, static_cast<uint32_t>(nh1));
      const uint32_t idesc_dq = umma_idesc_bf16(128, 64, 1);                  // A K-major, B MN-major
      const uint32_t idesc_tr = umma_idesc_bf16(128, 64, 1) | (1u << 15);     // A MN-major (transposed read), B MN-major
      // Base descriptors once; per MMA only the 16-byte-granular start address moves (one 64-bit add): the issue loop of
      // the first version rebuilt both descriptors per MMA and needed ~115 cycles per tcgen05.mma, more than the MMA itself.
      const uint64_t kQ = umma_desc_k_sw128(smem_u32(sQ)), kK = umma_desc_k_sw128(smem_u32(sK)), kdO = umma_desc_k_sw128(smem_u32(sdO)),
                     kV = umma_desc_k_sw128(smem_u32(sV)), kdS = umma_desc_k_sw128(smem_u32(sdS));
      const uint64_t mK = umma_desc(smem_u32(sK), 1024, 16384, 2), mP = umma_desc(smem_u32(sP), 1024, 16384, 2),
                     mdS = umma_desc(smem_u32(sdS), 1024, 16384, 2), mdO = umma_desc(smem_u32(sdO), 1024, 16384, 2),
                     mQ = umma_desc(smem_u32(sQ), 1024, 16384, 2);
      // Consecutive tcgen05.mma into the SAME accumulator are a dependent chain (~115 cycles per link measured here, against a
      // 32 / 64 cycle issue floor at N = 64 / 128): the first version, chain after chain, spent 3.7 k cycles per sub-step in the
      // tensor pipe.  So the five chains that are ready together -- S and dP of the NEXT sub-step, dQ / dV / dK of this one --
      // are issued round-robin, one k-step of each per round.  Every round is a switch over the set of chains still running:
      // straight-line, unpredicated MMAs per case (an unrolled `if (j < n) mma` form was compiled to predicated UTCHMMA with
      // wrong uniform descriptor registers and faulted with out-of-range shared addresses).
      auto mma_S = [&](int sub, int k) {
        const int kh = sub >> 1, t = sub & 1;
        umma_bf16(tmem, kQ + static_cast<uint64_t>(t * 1024 + 2 * k), kK + static_cast<uint64_t>(kh * 1024 + 2 * k),
                  kh == 0 ? idesc_s0 : idesc_s1, k != 0 ? 1u : 0u);
      };
      auto mma_dP = [&](int sub, int k) {
        const int kh = sub >> 1, t = sub & 1;
        umma_bf16(tmem + 128, kdO + static_cast<uint64_t>(t * 1024 + 2 * k), kV + static_cast<uint64_t>(kh * 1024 + 2 * k),
                  kh == 0 ? idesc_s0 : idesc_s1, k != 0 ? 1u : 0u);
      };
      auto mma_dQ = [&](int sub, int j) {     // dQ_t (+)= dS K_kh : A = dS tile (K-major), B = K rows as they sit (MN-major)
        const int kh = sub >> 1, t = sub & 1;
        umma_bf16(tmem + 256 + 64 * t, kdS + static_cast<uint64_t>((j >> 2) * 1024 + (j & 3) * 2),
                  mK + static_cast<uint64_t>(kh * 1024 + j * 128), idesc_dq, (kh | j) != 0 ? 1u : 0u);
      };
      auto mma_dV = [&](int sub, int ks) {    // dV_kh (+)= P^T dO_t : A = P tile read MN-major (M = keys, K = queries)
        const int t = sub & 1;
        umma_bf16(tmem + 384, mP + static_cast<uint64_t>(ks * 128), mdO + static_cast<uint64_t>(t * 1024 + ks * 128), idesc_tr,
                  (t | ks) != 0 ? 1u : 0u);
      };
      auto mma_dK = [&](int sub, int ks) {    // dK_kh (+)= dS^T Q_t
        const int t = sub & 1;
        umma_bf16(tmem + 448, mdS + static_cast<uint64_t>(ks * 128), mQ + static_cast<uint64_t>(t * 1024 + ks * 128), idesc_tr,
                  (t | ks) != 0 ? 1u : 0u);
      };
      uint32_t n_sub = 0, n_load = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++local_u) {
        const int view = unit / heads, h = unit - view * heads;
        BT_STAMP(0);
        // O_0 / O_1 ride in the (dead) P tile: the softmax threads take Delta = rowsum(dO o O) from shared memory before the
        // first P is written there (global row loads for Delta cost ~10 k cycles per unit behind the TMA traffic)
        mbar_expect_tx(bar_load0, 2 * KB + 3 * 16384);
        tma_load_3d(&tmKV, bar_load0, sK, d + h * DH, 0, view);
        tma_load_3d(&tmQ, bar_load0, sQ, h * DH, 0, view);
        tma_load_3d(&tmKV, bar_load0, sV, 2 * d + h * DH, 0, view);
        tma_load_3d(&tmdO, bar_load0, sdO, h * DH, 0, view);
        tma_load_3d(&tmO, bar_load0, sP, h * DH, 0, view);
        mbar_expect_tx(bar_load1, 3 * 16384);
        tma_load_3d(&tmQ, bar_load1, sQ + 16384, h * DH, 128, view);
        tma_load_3d(&tmdO, bar_load1, sdO + 16384, h * DH, 128, view);
        tma_load_3d(&tmO, bar_load1, sP + 16384, h * DH, 128, view);
        mbar_wait(bar_load0, n_load & 1);
        BT_STAMP(1);
        tc_fence_after();
#pragma unroll 1
        for (int k = 0; k < 4; ++k) { mma_S(0, k); mma_dP(0, k); }
        umma_commit(bar_sdp);
        if (unit + static_cast<int>(gridDim.x) < units) {      // the next unit's operands: HBM -> L2 behind this unit's compute
          const int nu = unit + gridDim.x, nv = nu / heads, nhd = nu - nv * heads;
          tma_prefetch_l2_3d(&tmKV, d + nhd * DH, 0, nv);
          tma_prefetch_l2_3d(&tmKV, 2 * d + nhd * DH, 0, nv);
          tma_prefetch_l2_3d(&tmQ, nhd * DH, 0, nv);
          tma_prefetch_l2_3d(&tmQ, nhd * DH, 128, nv);
          tma_prefetch_l2_3d(&tmdO, nhd * DH, 0, nv);
          tma_prefetch_l2_3d(&tmdO, nhd * DH, 128, nv);
          tma_prefetch_l2_3d(&tmO, nhd * DH, 0, nv);
          tma_prefetch_l2_3d(&tmO, nhd * DH, 128, nv);
        }
        for (int sub = 0; sub < 4; ++sub) {
          const int kh = sub >> 1, t = sub & 1;
          const int nk = (kh == 0 ? 128 : nh1) >> 4;     // 16-key k-steps of dQ
          const int nq = t == 0 ? 8 : nq1;                // 16-query k-steps of dV / dK
          // S / dP of the NEXT sub-step as soon as this sub-step's S / dP have left TMEM (the softmax threads are still writing
          // the tiles then): they are complete by the time the tiles are handed over, so the threads start on them at once
          mbar_wait(bar_drn, n_sub & 1);
          tc_fence_after();
          if (sub == 0) { mbar_wait(bar_load1, n_load & 1); tc_fence_after(); }
          if (sub < 3) {
#pragma unroll 1
            for (int k = 0; k < 4; ++k) { mma_S(sub + 1, k); mma_dP(sub + 1, k); }
            umma_commit(bar_sdp);
          }
          mbar_wait(bar_pds, n_sub & 1);      // P / dS of this sub-step are in smem
          ++n_sub;
          BT_STAMP(2 + 2 * sub);
          tc_fence_after();
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {       // dQ / dV / dK of this sub-step: three chains, round-robin
            switch ((i < nk ? 1 : 0) | (i < nq ? 2 : 0)) {
              case 3: mma_dQ(sub, i); mma_dV(sub, i); mma_dK(sub, i); break;
              case 2: mma_dV(sub, i); mma_dK(sub, i); break;
              case 1: mma_dQ(sub, i); break;
              default: break;
            }
          }
          umma_commit(bar_s2);
          if (sub == 3) umma_commit(bar_kv);
          BT_STAMP(3 + 2 * sub);
        }
        mbar_wait(bar_kv, n_load & 1);      // every MMA of the unit has completed: the smem operands may be reloaded
        ++n_load;
        BT_STAMP(10);
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;                  // TMEM lane quadrant a warp may access = warp index % 4
    const int half = (warp - 1) >> 2;           // warps 1-4: key columns [0,64) of a sub-step; warps 5-8: [64,128)
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    // this thread's 64 key columns = one 64-key block of the tiles: block `half`, row `row`
    const uint32_t p_row = smem_u32(sP) + half * 16384 + row * 128, ds_row = smem_u32(sdS) + half * 16384 + row * 128;
    uint8_t* my_out = sOut + half * 16384;       // staging tile of this warpgroup (half 0: dV then dQ_0; half 1: dK then dQ_1)
    const uint32_t out_row = smem_u32(my_out) + row * 128;
    const bool issuer = quad == 1 && lane == 0;  // warps 1 and 5: the TMA-store thread of each warpgroup
    const bool st = threadIdx.x == 32;
    uint32_t n_sub = 0;             // sub-steps seen (parity of bar_sdp; bar_s2 of sub-step n - 1 has parity (n - 1) & 1)
    uint32_t n_unit = 0;            // units seen (parity of the load barriers)

    // one accumulator tile (64 TMEM columns at `col`, lane = row) -> bf16 staging tile -> one TMA store; rows >= tokens clipped
    auto drain_tile = [&](uint32_t col, int c0, int r0, int view) {
      uint32_t a[32], b[32];
      tmem_ld_32x32b_x32(trow + col, a);
      tmem_ld_32x32b_x32(trow + col + 32, b);
      tmem_ld_wait();
      if (issuer) bulk_wait_read<0>();         // the previous store of this warpgroup has finished reading the staging tile
      named_bar_sync(1 + half, 128);
      bt_stage_row(a, b, out_row, sw);
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (issuer) {
        tma_store_3d(&tmDst, my_out, c0, r0, view);
        bulk_commit();
      }
    };

    float lse_next = INFINITY;
    if (static_cast<int>(blockIdx.x) < units && half * 128 + row < tokens) {
      const int v0 = blockIdx.x / heads, h0 = blockIdx.x - v0 * heads;
      lse_next = __ldg(lse + (static_cast<size_t>(v0) * heads + h0) * tokens + half * 128 + row) * LOG2E;
    }
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++local_u) {
      const int view = unit / heads, h = unit - view * heads;
      if (st) BT_STAMP(32);
      // Delta_q = sum_d dO_qd O_qd of the unit's 256 query rows: warpgroup `half` takes tile `half`, one row per thread, O from
      // the P tile (TMA-loaded there while it is dead) and dO from its own tile, both 128B-swizzled (conflict-free 16-byte
      // reads); rows >= tokens are zero-filled.  lse_q comes from global memory; rows >= tokens get +inf: P = 2^(-inf) = 0 and
      // dS = 0 without a branch in the inner loop.  Both are exchanged through smem (every thread needs both tiles' values).
      {
        const int q = half * 128 + row;
        const float l = lse_next;      // loaded one unit ahead (a global load here waits ~2 k cycles behind the TMA traffic)
        {
          const int nu = unit + gridDim.x, nv = nu / heads, nhd = nu - nv * heads;
          lse_next = nu < units && q < tokens ? __ldg(lse + (static_cast<size_t>(nv) * heads + nhd) * tokens + q) * LOG2E : INFINITY;
        }
        mbar_wait(half == 0 ? bar_load0 : bar_load1, n_unit & 1);
        const uint32_t o_row = smem_u32(sP) + half * 16384 + row * 128, do_row = smem_u32(sdO) + half * 16384 + row * 128;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t uo[4], ud[4];
          const uint32_t o16 = (static_cast<uint32_t>(c) ^ sw) << 4;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(uo[0]), "=r"(uo[1]), "=r"(uo[2]), "=r"(uo[3]) : "r"(o_row + o16));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ud[0]), "=r"(ud[1]), "=r"(ud[2]), "=r"(ud[3]) : "r"(do_row + o16));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 fo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uo[e]));
            const float2 fd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ud[e]));
            acc = fmaf(fo.x, fd.x, acc);
            acc = fmaf(fo.y, fd.y, acc);
          }
        }
        sDelta[q] = acc;
        sLse2[q] = l;
      }
      ++n_unit;
      if (st) BT_STAMP(53);
      if (issuer) bulk_wait_read<0>();   // the previous unit's last stores have finished reading the dS tile (used as staging)
      named_bar_sync(3, 256);
      const float delta0 = sDelta[row], delta1 = sDelta[128 + row], lse20 = sLse2[row], lse21 = sLse2[128 + row];
      if (st) BT_STAMP(33);
      for (int sub = 0; sub < 4; ++sub) {
        const int kh = sub >> 1, t = sub & 1;
        const int nh = kh == 0 ? 128 : nh1;
        const int ncol = nh - half * 64 < 64 ? nh - half * 64 : 64;   // this thread's key columns in this sub-step: 64/48/32/16/<=0
        mbar_wait(bar_sdp, n_sub & 1);
        if (st) BT_STAMP(34 + 4 * sub);
        tc_fence_after();
        const bool warp_active = t * 128 + quad * 32 < tokens;        // warp-uniform: at least one real query row
        uint32_t pk[32], dk[32];                                      // packed bf16 pairs of P and dS: up to 64 columns
        if (warp_active) {
          const float l2 = t == 0 ? lse20 : lse21, dl = t == 0 ? delta0 : delta1;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = half * 64 + cc * 32;
            if (cc * 32 + 32 <= ncol) {
              uint32_t s[32], dp[32];
              tmem_ld_32x32b_x32(trow + c, s);
              tmem_ld_32x32b_x32(trow + 128 + c, dp);
              tmem_ld_wait();
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                bt_group8(s + q4 * 8, dp + q4 * 8, scale_log2, l2, dl, scale, pk + cc * 16 + q4 * 4, dk + cc * 16 + q4 * 4);
            } else if (cc * 32 < ncol) {
              uint32_t s[16], dp[16];
              tmem_ld_32x32b_x16(trow + c, s);
              tmem_ld_32x32b_x16(trow + 128 + c, dp);
              tmem_ld_wait();
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4)
                bt_group8(s + q4 * 8, dp + q4 * 8, scale_log2, l2, dl, scale, pk + cc * 16 + q4 * 4, dk + cc * 16 + q4 * 4);
            }
          }
        }
        tc_fence_before();            // this warp's tcgen05.ld of S / dP have completed
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_drn);
        if (st) BT_STAMP(35 + 4 * sub);
        // The tiles are still being read by the dQ / dV / dK MMAs of the previous sub-step (they were issued behind this
        // sub-step's S / dP): wait for them, and drain what they completed, before overwriting the tiles.
        if (n_sub > 0) mbar_wait(bar_s2, (n_sub - 1) & 1);
        if (sub == 2) {               // the first key half is complete: dV_0 (warpgroup 0) / dK_0 (warpgroup 1), TMEM lane = key
          tc_fence_after();
          drain_tile(384 + 64 * half, (2 - half) * d + h * DH, 0, view);
          tc_fence_before();
        }
        ++n_sub;
        if (st) BT_STAMP(36 + 4 * sub);
        if (warp_active) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (g * 8 < ncol) {
              const uint32_t o16 = (static_cast<uint32_t>(g) ^ sw) << 4;
              bt_st16(p_row + o16, pk + g * 4);
              bt_st16(ds_row + o16, dk + g * 4);
            }
          }
        }
        fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pds);
        if (st) BT_STAMP(37 + 4 * sub);
      }
      // end of the unit: dV_1 / dK_1 (TMEM lane = key 128 + row) and dQ_0 / dQ_1 (TMEM lane = query of tile `half`)
      mbar_wait(bar_kv, (n_unit - 1) & 1);
      if (st) BT_STAMP(50);
      tc_fence_after();
      {   // dV_1 / dK_1 -> this warpgroup's staging tile, dQ_half -> its (dead) 64-key block of the dS tile; one hand-over for both
        uint32_t a[32], b[32];
        tmem_ld_32x32b_x32(trow + 384 + 64 * half, a);
        tmem_ld_32x32b_x32(trow + 384 + 64 * half + 32, b);
        tmem_ld_wait();
        if (issuer) bulk_wait_read<0>();         // the mid-unit store has finished reading the staging tile
        named_bar_sync(1 + half, 128);
        bt_stage_row(a, b, out_row, sw);
        tmem_ld_32x32b_x32(trow + 256 + 64 * half, a);
        tmem_ld_32x32b_x32(trow + 256 + 64 * half + 32, b);
        tmem_ld_wait();
        bt_stage_row(a, b, ds_row, sw);
        fence_proxy_async_smem();
        named_bar_sync(1 + half, 128);
        if (issuer) {
          tma_store_3d(&tmDst, my_out, (2 - half) * d + h * DH, 128, view);
          tma_store_3d(&tmDst, sdS + half * 16384, h * DH, half * 128, view);
          bulk_commit();
        }
      }
      tc_fence_before();
      if (st) BT_STAMP(51);
    }
    if (issuer) bulk_wait<0>();
  }
#undef BT_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
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
size_t attention_bwd_smem(int tokens) {      // two full matrices + row statistics + per-warp staging (see the kernel)
  const int nkp = (tokens + 63) / 64 * 64, tiles = (tokens + 15) / 16;
  return static_cast<size_t>(2 * nkp) * LDS * sizeof(bf16) + 2 * nkp * sizeof(float) +
         static_cast<size_t>(pick_warps(tiles)) * 2 * 16 * LDS * sizeof(bf16);
}

static bool launch_attention_fwd_tc(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                                    cudaStream_t st) {
  const int keys = (tokens + 15) / 16 * 16;
  const int tail = keys % 64;
  if (keys < 64 || keys > 208 || (tail != 0 && tail != 16)) return false;   // P-block placement covers <= 3 blocks + 16-key tail
  const int q_tiles = (tokens + 127) / 128;
  const int d = heads * DH;
  const size_t smem = 16384 + 2 * static_cast<size_t>(keys) * 128 + 16384 + 64 + 1024;
  if (tail && 16384 + 4096 > keys * 128) return false;
  CUtensorMap tq, tkv, to;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(tokens) * 3 * d * 2};
  const uint32_t boxq[3] = {64, 128, 1}, boxkv[3] = {64, static_cast<uint32_t>(keys), 1};
  if (!encode_tiled_map(&tq, 0, qkv, 3, dims, strides, boxq, 128)) return false;
  if (!encode_tiled_map(&tkv, 0, qkv, 3, dims, strides, boxkv, 128)) return false;
  const uint64_t odims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t ostrides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(tokens) * d * 2};
  const uint32_t obox[3] = {64, 128, 1};
  if (!encode_tiled_map(&to, 0, out, 3, odims, ostrides, obox, 128)) return false;
  const int dv = current_device_slot();
  static size_t configured_dev[MAX_DEVICES] = {};
  static int num_sms_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[dv];
  int& num_sms = num_sms_dev[dv];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dv);
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = smem;
  }
  const int items = V * heads * q_tiles;
  const int grid = items < 2 * num_sms ? items : 2 * num_sms;
  static long long* dbg = nullptr;
  static const bool want_dbg = std::getenv("TTL_ATTN_DBG") != nullptr;
  if (want_dbg && dbg == nullptr) { cudaMallocManaged(&dbg, 8 * 16 * 8 * sizeof(long long)); }
  if (want_dbg) std::memset(dbg, 0, 8 * 16 * 8 * sizeof(long long));
  launch_pdl(attention_fwd_tc_kernel, dim3(grid), dim3(TC_THREADS), smem, st, tq, tkv, to, lse, tokens, heads, items, q_tiles, keys,
             scale * LOG2E, want_dbg ? dbg : nullptr);
  if (want_dbg) {   // development aid: per-stage clock64 deltas of the first items of CTAs 0..7
    cudaStreamSynchronize(st);
    static int printed = 0;
    if (items >= 2000 && printed++ == 3) {
      const char* nm[7] = {"issue", "landed", "S ready", "pass1", "pass2", "O ready", "retired"};
      for (int cta = 0; cta < 8; cta += 7)
        for (int it = 1; it < 6; ++it) {
          const long long* t = dbg + (cta * 16 + it) * 8;
          std::fprintf(stderr, "cta %d item %d:", cta, it);
          for (int k = 1; k < 7; ++k) std::fprintf(stderr, " %s +%lld", nm[k], t[k] - t[0]);
          std::fprintf(stderr, " | next issue +%lld\n", (dbg + (cta * 16 + it + 1) * 8)[0] - t[0]);
        }
    }
  }
  return true;
}

static bool launch_attention_fwd_pp(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                                    cudaStream_t st, int descending) {
  // 257 tokens (ViT-L/14): 256 keys / queries through the MMAs (2 tiles x 256 TMEM columns), key 256 folded in by the softmax
  // threads, query 256 by the single-row kernel below
  const int xkey = tokens == 257 ? 256 : -1;
  const int keys = xkey >= 0 ? 256 : (tokens + 15) / 16 * 16;
  const int tail = keys % 64;
  if (xkey < 0 && (keys < 128 || keys > 208 || (tail != 0 && tail != 16) || tokens <= 128 || tokens > 256)) return false;   // exactly two query tiles
  const int d = heads * DH;
  const int n_full = keys / 64;
  const int ptile = n_full * 16384 + (tail ? 4096 : 0) > PP_PTILE ? n_full * 16384 + (tail ? 4096 : 0) : PP_PTILE;
  int vbufs = 2;
  size_t smem = 2 * 16384 + static_cast<size_t>(1 + vbufs) * keys * 128 + 2 * static_cast<size_t>(ptile) + 256 + 1024;
  if (smem > 227 * 1024) {
    vbufs = 1;
    smem = 2 * 16384 + static_cast<size_t>(1 + vbufs) * keys * 128 + 2 * static_cast<size_t>(ptile) + 256 + 1024;
  }
  if (xkey >= 0) {
    if (vbufs != 1) return false;          // the extra-query warps' barriers assume the single V buffer of this form
    smem += 1280;                          // probabilities + partials of the extra query row
  }
  if (smem > 227 * 1024) return false;
  CUtensorMap tq, tkv, to;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(tokens) * 3 * d * 2};
  const uint32_t boxq[3] = {64, 128, 1}, boxkv[3] = {64, static_cast<uint32_t>(keys), 1};
  if (!encode_tiled_map(&tq, 0, qkv, 3, dims, strides, boxq, 128)) return false;
  if (!encode_tiled_map(&tkv, 0, qkv, 3, dims, strides, boxkv, 128)) return false;
  const uint64_t odims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t ostrides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(tokens) * d * 2};
  const uint32_t obox[3] = {64, 128, 1};
  if (!encode_tiled_map(&to, 0, out, 3, odims, ostrides, obox, 128)) return false;
  const int dv = current_device_slot();
  static size_t configured_dev[2][MAX_DEVICES] = {};      // per instantiation and device
  static int num_sms_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[xkey >= 0 ? 1 : 0][dv];
  int& num_sms = num_sms_dev[dv];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dv);
  if (smem > configured) {
    const cudaError_t ea = xkey >= 0 ? cudaFuncSetAttribute(attention_fwd_pp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                            static_cast<int>(smem))
                                     : cudaFuncSetAttribute(attention_fwd_pp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                            static_cast<int>(smem));
    if (ea != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = smem;
  }
  const int units = V * heads;
  const int grid = units < num_sms ? units : num_sms;
  static long long* dbg = nullptr;
  static const bool want_dbg = std::getenv("TTL_ATTN_DBG") != nullptr;
  if (want_dbg && dbg == nullptr) cudaMallocManaged(&dbg, 12 * 2 * 16 * sizeof(long long));
  if (want_dbg) std::memset(dbg, 0, 12 * 2 * 16 * sizeof(long long));
  const bool ok = (xkey >= 0 ? launch_pdl(attention_fwd_pp_kernel<true>, dim3(grid), dim3(PP_THREADS_XK), smem, st, tq, tkv, to, lse, tokens,
                                          heads, units, keys, scale * LOG2E, want_dbg ? dbg : nullptr, descending, ptile, vbufs, xkey, qkv, out)
                             : launch_pdl(attention_fwd_pp_kernel<false>, dim3(grid), dim3(PP_THREADS), smem, st, tq, tkv, to, lse, tokens,
                                          heads, units, keys, scale * LOG2E, want_dbg ? dbg : nullptr, descending, ptile, vbufs, xkey, qkv, out)) ==
                  cudaSuccess;
  if (want_dbg) {
    cudaStreamSynchronize(st);
    static int printed = 0;
    if (units >= 2000 && printed++ == 3) {
      const char* nm[7] = {"loop top", "S ready", "pass1", "stage free", "pass2", "O ready", "stored"};
      for (int it = 1; it < 8; ++it)
        for (int t = 0; t < 2; ++t) {
          const long long* q = dbg + (it * 2 + t) * 16;
          std::fprintf(stderr, "unit %d tile %d:", it, t);
          for (int k = 1; k < 7; ++k) std::fprintf(stderr, " %s +%lld", nm[k], q[k] - q[0]);
          std::fprintf(stderr, " | P blocks +%lld +%lld +%lld", q[8] - q[0], q[9] - q[0], q[10] - q[0]);
          std::fprintf(stderr, " | next top +%lld | since tile0 top %+lld\n", (dbg + ((it + 1) * 2 + t) * 16)[0] - q[0], q[0] - (dbg + it * 2 * 16)[0]);
        }
    }
  }
  return ok;
}

static bool launch_attention_fwd_pt(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                                    cudaStream_t st, int descending) {
  const int keys = (tokens + 15) / 16 * 16;
  if (keys != PT_KEYS) return false;      // 193..208 tokens: two query tiles, twelve 16-key steps + a 16-key tail
  const int d = heads * DH;
  const size_t smem = 4 * 16384 + static_cast<size_t>(4) * PT_KB + 2 * 16384 + 256 + 1024;
  CUtensorMap tq, tkv, to;
  static const int variant = std::getenv("TTL_PT_VARIANT") ? std::atoi(std::getenv("TTL_PT_VARIANT")) : 0;
  static const bool want_dbg = std::getenv("TTL_ATTN_DBG") != nullptr;
  auto kern = want_dbg ? attention_fwd_pt_kernel<0x88, true>
                       : variant == 1 ? attention_fwd_pt_kernel<0x80, false>
                                      : variant == 2 ? attention_fwd_pt_kernel<0x00, false> : attention_fwd_pt_kernel<0x88, false>;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(tokens) * 3 * d * 2};
  const uint32_t boxq[3] = {64, 128, 1}, boxkv[3] = {64, static_cast<uint32_t>(keys), 1};
  if (!encode_tiled_map(&tq, 0, qkv, 3, dims, strides, boxq, 128)) return false;
  if (!encode_tiled_map(&tkv, 0, qkv, 3, dims, strides, boxkv, 128)) return false;
  const uint64_t odims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t ostrides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(tokens) * d * 2};
  const uint32_t obox[3] = {64, 32, 1};      // one softmax warp's rows
  if (!encode_tiled_map(&to, 0, out, 3, odims, ostrides, obox, 128)) return false;
  const int dv = current_device_slot();
  static bool configured_dev[MAX_DEVICES] = {};
  static int num_sms_dev[MAX_DEVICES] = {};
  int& num_sms = num_sms_dev[dv];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dv);
  if (!configured_dev[dv]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured_dev[dv] = true;
  }
  const int units = V * heads;
  const int grid = units < num_sms ? units : num_sms;
  static long long* dbg = nullptr;
  if (want_dbg && dbg == nullptr) cudaMallocManaged(&dbg, 12 * 2 * 16 * sizeof(long long));
  if (want_dbg) std::memset(dbg, 0, 12 * 2 * 16 * sizeof(long long));
  const bool ok = launch_pdl(kern, dim3(grid), dim3(PT_THREADS), smem, st, tq, tkv, to, lse, tokens, heads, units,
                             scale * LOG2E, want_dbg ? dbg : nullptr, descending) == cudaSuccess;
  if (want_dbg) {
    cudaStreamSynchronize(st);
    static int printed = 0;
    if (units >= 2000 && printed++ == 3) {
      const char* nm[7] = {"loop top", "S ready", "pass1", "", "pass2", "O ready", "stored"};
      for (int it = 3; it < 6; ++it)
        for (int t = 0; t < 2; ++t) {
          const long long* q = dbg + (it * 2 + t) * 16;
          std::fprintf(stderr, "pt unit %d tile %d:", it, t);
          for (int k = 1; k < 7; ++k)
            if (k != 3) std::fprintf(stderr, " %s +%lld", nm[k], q[k] - q[0]);
          std::fprintf(stderr, " | P blocks +%lld +%lld +%lld", q[8] - q[0], q[9] - q[0], q[10] - q[0]);
          std::fprintf(stderr, " | next top +%lld | since tile0 top %+lld\n", (dbg + ((it + 1) * 2 + t) * 16)[0] - q[0], q[0] - (dbg + it * 2 * 16)[0]);
        }
    }
  }
  return ok;
}

// 257 tokens (ViT-L/14): the P-in-TMEM kernel with 256 keys / queries through the MMAs, key 256 folded in by the softmax threads and
// query row 256 on one extra warp (attention_fwd_px_kernel)
static bool launch_attention_fwd_px(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                                    cudaStream_t st, int descending) {
  if (tokens != 257) return false;
  const int d = heads * DH;
  const size_t smem = 4 * 16384 + static_cast<size_t>(4) * PX_KB + 2 * 16384 + 1024 + 1024;      // 226 KB
  CUtensorMap tq, tkv, to, trow;
  auto kern = attention_fwd_px_kernel<0x88>;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(tokens) * 3 * d * 2};
  const uint32_t boxq[3] = {64, 128, 1}, boxkv[3] = {64, 256, 1}, boxrow[3] = {64, 1, 1};
  if (!encode_tiled_map(&tq, 0, qkv, 3, dims, strides, boxq, 128)) return false;
  if (!encode_tiled_map(&tkv, 0, qkv, 3, dims, strides, boxkv, 128)) return false;
  if (!encode_tiled_map(&trow, 0, qkv, 3, dims, strides, boxrow, 0)) return false;      // one token's 64 values of a head, not swizzled
  const uint64_t odims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t ostrides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(tokens) * d * 2};
  const uint32_t obox[3] = {64, 32, 1};      // one softmax warp's rows
  if (!encode_tiled_map(&to, 0, out, 3, odims, ostrides, obox, 128)) return false;
  const int dv = current_device_slot();
  static bool configured_dev[MAX_DEVICES] = {};
  static int num_sms_dev[MAX_DEVICES] = {};
  int& num_sms = num_sms_dev[dv];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dv);
  if (!configured_dev[dv]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured_dev[dv] = true;
  }
  const int units = V * heads;
  const int grid = units < num_sms ? units : num_sms;
  static const int dbg = std::getenv("TTL_PX_DBG") ? std::atoi(std::getenv("TTL_PX_DBG")) : 0;      // development: 1 = idle extra warp, 2 = no key 256
  return launch_pdl(kern, dim3(grid), dim3(PX_THREADS), smem, st, tq, tkv, to, trow, lse, heads, units, scale * LOG2E, descending,
                    qkv, out, dbg) == cudaSuccess;
}

static bool launch_attention_fwd_tma(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                                     cudaStream_t st) {
  const int rows_pad = (tokens + 15) / 16 * 16, rows_alloc = (rows_pad + 31) / 32 * 32;
  if (rows_alloc > ATT_MAX_WARPS * 32) return false;
  const int nbox = (rows_alloc + 255) / 256;
  if (rows_alloc % nbox != 0) return false;
  const int box_rows = rows_alloc / nbox;
  const size_t smem = static_cast<size_t>(6) * rows_alloc * 128 + 64 + 1024;
  if (smem > 227 * 1024) return false;
  const int d = heads * DH;
  CUtensorMap tq, to;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
    const uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(tokens) * 3 * d * 2};
    const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
    if (!encode_tiled_map(&tq, 0, qkv, 3, dims, strides, box, 128)) return false;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
    const uint64_t strides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(tokens) * d * 2};
    const uint32_t box[3] = {64, 32, 1};
    if (!encode_tiled_map(&to, 0, out, 3, dims, strides, box, 128)) return false;
  }
  const int dv = current_device_slot();
  static size_t configured_dev[MAX_DEVICES] = {};
  static int num_sms_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[dv];
  int& num_sms = num_sms_dev[dv];
  if (num_sms == 0) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dv);
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = smem;
  }
  const int units = V * heads;
  const int grid = units < num_sms ? units : num_sms;
  launch_pdl(attention_fwd_tma_kernel, dim3(grid), dim3((rows_alloc / 32) * 32), smem, st, tq, to, lse, tokens, heads, units,
             scale * LOG2E, rows_alloc, rows_pad, nbox, box_rows);
  return true;
}

void launch_attention_fwd(const bf16* qkv, bf16* out, float* lse, int V, int tokens, int heads, float scale,
                          cudaStream_t st, int descending, int causal) {   // only the default kernel honours `descending`
  // TTL_ATTN: unset = the fastest kernel the geometry allows: "pt" (193..208 tokens) / "px" (257 tokens) = tcgen05 with P kept in
  // TMEM, "pp" = tcgen05 with both query tiles of a unit in flight and P through shared memory (129..208 and 257 tokens), "tc" =
  // tcgen05 with one tile per work item and two CTAs per SM, "mma" = TMA-fed mma.sync kernel, "legacy" = first kernel
  static const char* mode = std::getenv("TTL_ATTN");
  const bool want_pt = mode == nullptr || (mode[0] == 'p' && mode[1] == 't');      // default for 193..208 tokens: P kept in TMEM
  const bool want_px = mode == nullptr || (mode[0] == 'p' && mode[1] == 'x');      // default for 257 tokens: the same with an extra key / query
  const bool want_pp = mode == nullptr || mode[0] == 'p';
  const bool want_tc = mode == nullptr || mode[0] == 't' || mode[0] == 'p';
  const bool want_tma = mode == nullptr || mode[0] != 'l';
  // the causal form (77-token text tower) runs on the general kernel below
  if (!causal && want_pt && launch_attention_fwd_pt(qkv, out, lse, V, tokens, heads, scale, st, descending)) return;
  if (!causal && want_px && launch_attention_fwd_px(qkv, out, lse, V, tokens, heads, scale, st, descending)) return;
  if (!causal && want_pp && launch_attention_fwd_pp(qkv, out, lse, V, tokens, heads, scale, st, descending)) return;
  if (!causal && want_tc && launch_attention_fwd_tc(qkv, out, lse, V, tokens, heads, scale, st)) return;
  if (!causal && want_tma && launch_attention_fwd_tma(qkv, out, lse, V, tokens, heads, scale, st)) return;
  const int q_tiles = (tokens + 15) / 16, nkp = (tokens + 63) / 64 * 64;
  const size_t smem = attention_fwd_smem(tokens);
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  launch_pdl(attention_fwd_kernel, dim3(heads, V), dim3(pick_warps(q_tiles) * 32), smem, st, qkv, out, lse, tokens, heads,
             scale * LOG2E, q_tiles, nkp, causal);
}

void launch_attention_cls(const bf16* q_cls, const bf16* qkv, bf16* out_cls, int V, int tokens, int heads, float scale,
                          cudaStream_t st) {
  const size_t dd = static_cast<size_t>(heads) * DH;
  launch_pdl(attention_cls_kernel, dim3(heads, V), dim3(128), (64 + tokens) * sizeof(float), st, q_cls, dd, qkv, out_cls, dd,
             static_cast<float*>(nullptr), 0, tokens, heads, scale * LOG2E);
}

static bool launch_attention_bwd_tc(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int V,
                                    int tokens, int heads, float scale, cudaStream_t st) {
  if (tokens < 129 || tokens > 208) return false;       // two 128-query tiles, keys split 128 + (16..80)
  const int keys = (tokens + 15) / 16 * 16;
  const int d = heads * DH;
  const size_t smem = 2 * static_cast<size_t>(keys) * 128 + 5 * 32768 + 2048 + 256 + 1024;
  CUtensorMap tq, tkv, tdo;
  const uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(tokens) * 3 * d * 2};
  const uint32_t boxq[3] = {64, 128, 1}, boxkv[3] = {64, static_cast<uint32_t>(keys), 1};
  if (!encode_tiled_map(&tq, 0, qkv, 3, dims, strides, boxq, 128)) return false;
  if (!encode_tiled_map(&tkv, 0, qkv, 3, dims, strides, boxkv, 128)) return false;
  const uint64_t odims[3] = {static_cast<uint64_t>(d), static_cast<uint64_t>(tokens), static_cast<uint64_t>(V)};
  const uint64_t ostrides[2] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(tokens) * d * 2};
  if (!encode_tiled_map(&tdo, 0, dout, 3, odims, ostrides, boxq, 128)) return false;
  CUtensorMap tdst, to;
  if (!encode_tiled_map(&tdst, 0, dqkv, 3, dims, strides, boxq, 128)) return false;
  if (!encode_tiled_map(&to, 0, out, 3, odims, ostrides, boxq, 128)) return false;
  const int dv = current_device_slot();
  static size_t configured_dev[MAX_DEVICES] = {};
  static int num_sms_dev[MAX_DEVICES] = {};
  if (num_sms_dev[dv] == 0) cudaDeviceGetAttribute(&num_sms_dev[dv], cudaDevAttrMultiProcessorCount, dv);
  if (smem > configured_dev[dv]) {
    if (cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured_dev[dv] = smem;
  }
  const int units = V * heads;
  const int grid = units < num_sms_dev[dv] ? units : num_sms_dev[dv];
  static long long* dbg = nullptr;
  static const bool want_dbg = std::getenv("TTL_ATTN_DBG") != nullptr;
  if (want_dbg && dbg == nullptr) cudaMallocManaged(&dbg, 6 * 64 * sizeof(long long));
  if (want_dbg) { cudaStreamSynchronize(st); std::memset(dbg, 0, 6 * 64 * sizeof(long long)); }
  launch_pdl(attention_bwd_tc_kernel, dim3(grid), dim3(BT_THREADS), smem, st, tq, tkv, tdo, to, tdst, lse, tokens, heads, units, keys,
             scale, want_dbg ? dbg : nullptr);
  if (want_dbg) {   // per-stage clock64 deltas of units 1..3 of CTA 0
    cudaStreamSynchronize(st);
    static int printed = 0;
    if (units >= 2000 && printed++ == 2) {
      for (int u = 1; u < 4; ++u) {
        const long long* t = dbg + u * 64;
        std::fprintf(stderr, "bwd unit %d MMA thread: landed +%lld", u, t[1] - t[0]);
        for (int sb = 0; sb < 4; ++sb) std::fprintf(stderr, " | pds%d +%lld issued +%lld", sb, t[2 + 2 * sb] - t[0], t[3 + 2 * sb] - t[0]);
        std::fprintf(stderr, " | all done +%lld | next unit +%lld\n", t[10] - t[0], (dbg + (u + 1) * 64)[0] - t[0]);
        std::fprintf(stderr, "bwd unit %d softmax warp: start %+lld delta +%lld prologue +%lld", u, t[32] - t[0], t[53] - t[0], t[33] - t[0]);
        for (int sb = 0; sb < 4; ++sb)
          std::fprintf(stderr, " | sdp%d +%lld computed +%lld tiles free +%lld stored +%lld", sb, t[34 + 4 * sb] - t[0], t[35 + 4 * sb] - t[0],
                       t[36 + 4 * sb] - t[0], t[37 + 4 * sb] - t[0]);
        std::fprintf(stderr, " | kv +%lld drained +%lld\n", t[50] - t[0], t[51] - t[0]);
      }
    }
  }
  return true;
}

void launch_attention_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int V,
                          int tokens, int heads, float scale, cudaStream_t st, int causal, float* delta_ws) {
  // TTL_ATTN_BWD: unset / "tc" = tcgen05 kernel where the geometry allows (129..208 tokens), "mma" = the mma.sync kernel
  static const char* mode = std::getenv("TTL_ATTN_BWD");
  if (!causal && (mode == nullptr || mode[0] == 't') && launch_attention_bwd_tc(qkv, out, dout, lse, dqkv, V, tokens, heads, scale, st)) return;
  const int tiles = (tokens + 15) / 16, nkp = (tokens + 63) / 64 * 64;
  const size_t smem = attention_bwd_smem(tokens);
  static size_t configured_dev[MAX_DEVICES] = {};
  size_t& configured = configured_dev[current_device_slot()];
  if (smem > configured) {
    cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  const float* delta = nullptr;
  if (delta_ws != nullptr && tokens <= 128) {     // exact Delta (see attention_delta_kernel)
    const size_t dsm = (2 * static_cast<size_t>(tokens) * (DH + 1) + 8 * 2 * DH) * sizeof(float);
    static size_t dconf_dev[MAX_DEVICES] = {};
    size_t& dconf = dconf_dev[current_device_slot()];
    if (dsm > dconf) {
      cudaFuncSetAttribute(attention_delta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dsm));
      dconf = dsm;
    }
    launch_pdl(attention_delta_kernel, dim3(heads, V), dim3(256), dsm, st, qkv, dout, lse, delta_ws, tokens, heads, scale, causal);
    delta = delta_ws;
  }
  launch_pdl(attention_bwd_kernel, dim3(heads, V, 2), dim3(pick_warps(tiles) * 32), smem, st, qkv, out, dout, lse, dqkv, tokens,   // z: phase
             heads, scale, tiles, nkp, causal, delta);
}

}  // namespace ttl
