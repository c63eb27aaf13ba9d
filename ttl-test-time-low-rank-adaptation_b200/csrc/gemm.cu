// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epi( A[M,K] * B[N,K]^T ).
//   * operands: TMA (cp.async.bulk.tensor.2d, 128B swizzle) -> multi-stage smem ring
//   * math:     tcgen05.mma cta_group::1 kind::f16, M=128 x N=BLOCK_N x K=16 per instruction, issued by one
//               thread; fp32 accumulators live in TMEM, double-buffered (2 x BLOCK_N columns) so the
//               epilogue of tile i overlaps the main loop of tile i+1
//   * epilogue: 8 warps, tcgen05.ld 32x32b.x32 -> registers -> fused bias / QuickGELU / residual /
//               patch-embed scatter / dQuickGELU -> vectorised global stores
// Replaces the cuBLAS/cuDNN calls behind HF CLIPVisionEmbeddings/CLIPAttention/CLIPMLP and peft's LoRA
// Linear; an optional second operand pair adds the rank-r LoRA term into the same
// accumulator (y = W x + s B (A x), clip/custom_clip.py:583-591).
#include "gemm.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ttl {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + NUM_EPI_WARPS * 32;

template <int BLOCK_N>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 4096;   // one 32x32 fp32 block per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + STAGING_BYTES + 1024;  // +1024: alignment slack
};

struct EpiParams {
  int M, N;
  int kb1, kb2;
  const float* bias;
  void* out;
  int ldo;
  void* out2;
  const float* resid;
  int ldr;
  const __nv_bfloat16* aux;
  const float* pos;
  int tpv;
};

__device__ __forceinline__ float quick_gelu_sig(float z) { return __fdividef(1.0f, 1.0f + __expf(-1.702f * z)); }

__device__ __forceinline__ uint2 pack4(const float4& a) {
  uint2 u;
  u.x = pack_bf16(a.x, a.y);
  u.y = pack_bf16(a.z, a.w);
  return u;
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// One 32-row x 32-column block of the accumulator, owned by one warp.
//   phase 1: every lane holds ONE ROW (32 fp32 straight from tcgen05.ld) -> 16-byte chunks into the warp's smem block,
//            XOR-swizzled so both phases are bank-conflict free;
//   phase 2: lane l handles columns 4*(l&7)..+3 of rows (l>>3)+4j: global accesses are 128 B (fp32) / 64 B (bf16)
//            contiguous per row, 4 rows per instruction -> fully coalesced residual reads and output writes.
template <int EPI>
__device__ __forceinline__ void epilogue_block(const EpiParams& p, float* stag, int row0, int col0, int lane,
                                               const uint32_t (&r)[32]) {
  float4* s4 = reinterpret_cast<float4*>(stag);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s4[lane * 8 + (i ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                  __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
  __syncwarp();
  const int ci = lane & 7, col = col0 + ci * 4;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI != EPI_PATCH_F32 && EPI != EPI_GELU_BWD && p.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  // Issue every global read of this block before any use (8 independent 16-byte loads in flight per lane).
  float4 ext[8];
  uint2 zext[8];
  if (EPI == EPI_RESID_F32 || EPI == EPI_PATCH_F32 || EPI == EPI_GELU_BWD) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int row = row0 + (lane >> 3) + 4 * j;
      row = row < p.M ? row : p.M - 1;   // clamp instead of branching; out-of-range rows are not stored
      if (EPI == EPI_RESID_F32) {
        ext[j] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(row) * p.ldr + col);
      } else if (EPI == EPI_PATCH_F32) {
        const int patch = row % p.tpv;
        ext[j] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + patch) * p.N + col));
      } else {
        zext[j] = *reinterpret_cast<const uint2*>(p.aux + static_cast<size_t>(row) * p.ldo + col);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int rr = (lane >> 3) + 4 * j;
    const int row = row0 + rr;
    float4 a = add4(s4[rr * 8 + (ci ^ (rr & 7))], bias4);
    const size_t o = static_cast<size_t>(row) * p.ldo + col;
    if (EPI == EPI_BF16) {
      if (row < p.M) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack4(a);
    } else if (EPI == EPI_GELU) {
      const uint2 zp = pack4(a);
      a.x *= quick_gelu_sig(a.x); a.y *= quick_gelu_sig(a.y); a.z *= quick_gelu_sig(a.z); a.w *= quick_gelu_sig(a.w);
      if (row < p.M) {
        if (p.out2 != nullptr) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + o) = zp;
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack4(a);
      }
    } else if (EPI == EPI_RESID_F32) {
      a = add4(a, ext[j]);
      if (row < p.M) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) = a;
        if (p.out2 != nullptr) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + o) = pack4(a);
      }
    } else if (EPI == EPI_PATCH_F32) {
      a = add4(a, ext[j]);
      if (row < p.M) {
        const int view = row / p.tpv, patch = row - view * p.tpv;
        const size_t orow = static_cast<size_t>(view) * (p.tpv + 1) + 1 + patch;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col) = a;
      }
    } else if (EPI == EPI_F32) {
      if (row < p.M) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) = a;
        if (p.out2 != nullptr) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + o) = pack4(a);
      }
    } else if (EPI == EPI_GELU_BWD) {
      const float2 z01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zext[j].x));
      const float2 z23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zext[j].y));
      const float s0 = quick_gelu_sig(z01.x), s1 = quick_gelu_sig(z01.y), s2 = quick_gelu_sig(z23.x), s3 = quick_gelu_sig(z23.y);
      a.x *= s0 * (1.0f + 1.702f * z01.x * (1.0f - s0));
      a.y *= s1 * (1.0f + 1.702f * z01.y * (1.0f - s1));
      a.z *= s2 * (1.0f + 1.702f * z23.x * (1.0f - s2));
      a.w *= s3 * (1.0f + 1.702f * z23.y * (1.0f - s3));
      if (row < p.M) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pack4(a);
    }
  }
  __syncwarp();   // the block is rewritten by the next chunk
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const EpiParams p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* staging = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    if (p.kb2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], NUM_EPI_WARPS);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();      // prologue above overlaps the tail of the previous kernel; global memory is touched only below
  pdl_trigger();

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.kb1 + p.kb2;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES;
          if (kb < p.kb1) {
            tma_load_2d(&tmA1, &full[stage], a_dst, kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(&tmB1, &full[stage], b_dst, kb * BLOCK_K, n_blk * BLOCK_N);
          } else {
            tma_load_2d(&tmA2, &full[stage], a_dst, (kb - p.kb1) * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(&tmB2, &full[stage], b_dst, (kb - p.kb1) * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            umma_bf16(d_tmem, umma_desc_k_sw128(a_addr + k * UMMA_K * 2), umma_desc_k_sw128(b_addr + k * UMMA_K * 2),
                      idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // frees the smem slot when these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);       // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- epilogue warps
    const int ew = warp - 2;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = ew >> 2;           // which half of the BLOCK_N columns
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const int row0 = m_blk * BLOCK_M + quad * 32;
      float* stag = staging + ew * 1024;
#pragma unroll 1
      for (int c = half * (BLOCK_N / 2); c < (half + 1) * (BLOCK_N / 2); c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BLOCK_N + c, r);
        tmem_ld_wait();
        if (row0 < p.M) epilogue_block<EPI>(p, stag, row0, n_blk * BLOCK_N + c, lane, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}


// =============================================================================================== CTA-pair kernel
// Same contract as above for the big-M GEMMs of the 64-view forward, built around tcgen05 cta_group::2:
//   * a cluster of two CTAs (one SM each) owns a 256 x BLOCK_N output tile; each CTA stages its own 128 rows of A and
//     HALF of the B tile (BLOCK_N/2 rows) per k-block, so the pair reads (256 + BLOCK_N) x 64 bf16 from L2 per
//     256 x BLOCK_N x 64 MMA block instead of 2 x (128 + BLOCK_N) x 64 -- the 1-CTA kernel is L2->SM bandwidth bound;
//   * the leader CTA's single MMA thread issues 256 x BLOCK_N x 16 MMAs for both; completion is multicast to both
//     CTAs' mbarriers (tcgen05.commit ... multicast::cluster);
//   * epilogue: 8 warps per CTA, tcgen05.ld -> registers -> (bias / QuickGELU / + residual tile that TMA prefetched
//     into smem) -> 128B/64B-swizzled smem -> TMA store.  No per-thread global addressing, rows >= M are clipped by TMA.
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_THREADS = 64 + G2_EPI_WARPS * 32;
constexpr int G2_SMEM_MAX = 232448;   // 227 KB

template <int BLOCK_N, int EPI>
struct Cfg2 {
  static constexpr bool OUT_F32 = (EPI == EPI_RESID_F32 || EPI == EPI_F32);
  static constexpr int A_BYTES = 128 * BLOCK_K * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EBUF_BYTES = OUT_F32 ? 4096 : 2048;          // 32 rows x 32 columns
  static constexpr int EPI_BYTES = G2_EPI_WARPS * 2 * EBUF_BYTES;   // two buffers per epilogue warp
  static constexpr int BAR_BYTES = 512;   // (3 * STAGES + 4 + 2 * G2_EPI_WARPS) mbarriers + the TMEM slot: <= 44 * 8 + 4
  static constexpr int STAGES_RAW = (G2_SMEM_MAX - 1024 - BAR_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int CHUNKS = BLOCK_N / 64;                       // 32-column chunks per epilogue warp
  static_assert(BLOCK_N % 64 == 0 && BLOCK_N <= 256, "BLOCK_N");
  static_assert(B_BYTES % 1024 == 0, "B stage must keep 1024 B alignment");
};

struct Epi2Params {
  int M, N;
  int kb1, kb2;
  const float* bias;
  int dbg;   // TTL_GEMM_DBG bits (development only): 1 = no epilogue stores, 2 = no MMA, 4 = no TMA loads
  int rev;   // tiles walked from the last row block to the first
  const __nv_bfloat16* aux;   // EPI_GELU_BWD: z [M, N] bf16, row pitch ldaux
  int ldaux;
  __nv_bfloat16* out2;        // EPI_GELU: optional copy of the pre-activation z = acc + bias (bf16, row pitch ldaux), for the backward
  // EPI_RESID_F32 with a fused LayerNorm of the rows it writes (ln_out != nullptr).  The 2 N / BLOCK_N epilogue warps (of up to
  // N / BLOCK_N clusters) that own a piece of the same 32 rows each publish Welford statistics of their BLOCK_N / 2 columns and
  // count an arrival; one tile later every one of them re-reads ITS OWN 32 x BLOCK_N / 2 piece (fp32, L2-hot), by then with
  // the statistics of the complete rows at hand, and writes LayerNorm(row) as bf16 to ln_out: the same extra work for every warp
  // and tile.  (First form: the warp that arrived last swept all N columns of the 32 rows -- every tile then waited for its
  // slowest warp: 338 vs 396 samples/s.)  The tile walk stays round-robin, so the clusters working on one row block still
  // share its A tiles through L2 (the strip form, commit a995abe, walked n-inner and re-read A from HBM: 2.8 vs 1.1 GB per launch).
  const float* ln_gamma;
  const float* ln_beta;
  __nv_bfloat16* ln_out;
  int ld_ln;
  float ln_eps;
  float2* ln_stats;           // [rows][2 N / BLOCK_N]
  int* ln_cnt;                // [rows / 32], zeroed by the launcher
  const float* ln_src;        // == the output matrix (generic-proxy view of what tmOut stores), row pitch ld_src
  int ld_src;
};

// CL = 2: one CTA pair per cluster (above).  CL = 4: two pairs stacked along M share every B tile: each CTA loads a QUARTER of
// the B tile and TMA-multicasts it to the CTA of the same parity in the other pair, so a pair pulls (256 + BLOCK_N/2) x 64
// bf16 from L2 per 256 x BLOCK_N x 64 MMA block instead of (256 + BLOCK_N) x 64.  Parity-0 CTAs are the pair leaders: their
// B quarters signal the leaders' `full` barriers directly; parity-1 CTAs collect theirs on a local `fullB` barrier and the
// (otherwise idle) MMA warp of the non-leader forwards one remote arrival per stage to its leader.  A smem slot is
// released to all four producers only when BOTH pairs' MMAs have read it (`empty` counts two multicast commits).
template <int BLOCK_N, int EPI, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
             const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
             const Epi2Params p) {
  using C = Cfg2<BLOCK_N, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * C::A_BYTES;
  uint8_t* sE = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sE + C::EPI_BYTES);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* lbar = tempty + 2;                       // [G2_EPI_WARPS][2] residual-tile arrivals
  uint64_t* fullB = lbar + 2 * G2_EPI_WARPS;         // [STAGES] CL == 4, parity-1 CTAs: this CTA's B half has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fullB + C::STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0 .. CL-1
  const uint32_t rank = crank & 1;                   // parity inside the pair: 0 = leader
  const uint32_t pair = crank >> 1;                  // 0 (CL == 2) or 0/1 (CL == 4)
  const uint32_t lead_rank = crank & ~1u;            // cluster rank of this pair's leader
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmOut);
    if (EPI == EPI_RESID_F32) tma_prefetch_desc(&tmRes);
    if (p.kb2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], CL > 2 ? 3 : 2);    // one arrival per CTA's producer (+ the peer's B forwarder), leader's copy is used
      mbar_init(&empty[s], CL / 2);           // multicast tcgen05.commit of every pair in the cluster
      mbar_init(&fullB[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);                      // multicast tcgen05.commit
      mbar_init(&tempty[a], 2 * G2_EPI_WARPS);      // every epilogue warp of both CTAs (leader's copy is used)
    }
    for (int i = 0; i < 2 * G2_EPI_WARPS; ++i) mbar_init(&lbar[i], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();      // prologue above overlaps the tail of the previous kernel; global memory is touched only below
  pdl_trigger();

  constexpr int PAIRS = CL / 2;                       // pairs per cluster, stacked along M
  const int m_pairs = ((p.M + 255) / 256 + PAIRS - 1) / PAIRS;   // row blocks of 256 * PAIRS rows
  const int n_tiles = p.N / BLOCK_N;
  const int num_tiles = m_pairs * n_tiles;
  const int num_kb = p.kb1 + p.kb2;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * pair));
  const bool fuse_ln = EPI == EPI_RESID_F32 && CL == 2 && BLOCK_N == 256 && p.ln_out != nullptr;
  // the k-th tile of this cluster: round-robin over all tiles
  auto tile_at = [&](int k, int& m_pair, int& n_blk) -> bool {
    const int tile = cluster_id + k * num_clusters;
    if (tile >= num_tiles) return false;
    const int t2 = p.rev ? num_tiles - 1 - tile : tile;
    m_pair = t2 / n_tiles;
    n_blk = t2 - m_pair * n_tiles;
    return true;
  };

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (one thread per CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = mapa_u32(smem_u32(&full[0]), lead_rank);   // the pair LEADER's full barriers
      uint16_t bmask = 0;                                      // CL > 2: the CTAs of the same parity in every pair
      for (int q = 0; q < PAIRS; ++q) bmask |= static_cast<uint16_t>(1u << (rank + 2 * q));
      int m_pair = 0, n_blk = 0;
      for (int tk = 0; tile_at(tk, m_pair, n_blk); ++tk) {
        const int a_row = (m_pair * PAIRS + static_cast<int>(pair)) * 256 + static_cast<int>(rank) * 128;
        const int b_row = n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2) + static_cast<int>(pair) * (BLOCK_N / CL);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t fb = full0 + stage * 8;
          if (leader) mbar_expect_tx(&full[stage], (p.dbg & 4) ? 0 : 2 * C::A_BYTES + (CL > 2 ? 1 : 2) * C::B_BYTES);
          else mbar_arrive_cluster(fb);
          uint8_t* a_dst = sA + stage * C::A_BYTES;
          uint8_t* b_dst = sB + stage * C::B_BYTES + pair * (C::B_BYTES / PAIRS);
          const void* tA = kb < p.kb1 ? static_cast<const void*>(&tmA1) : static_cast<const void*>(&tmA2);
          const void* tB = kb < p.kb1 ? static_cast<const void*>(&tmB1) : static_cast<const void*>(&tmB2);
          const int kc = (kb < p.kb1 ? kb : kb - p.kb1) * BLOCK_K;
          if (p.dbg & 4) {
          } else {
            tma_load_2d_cg2(tA, fb, a_dst, kc, a_row);
            if (CL > 2) tma_load_2d_mc(tB, leader ? &full[stage] : &fullB[stage], b_dst, kc, b_row, bmask);
            else tma_load_2d_cg2(tB, fb, b_dst, kc, b_row);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int m_pair = 0, n_blk = 0;
      for (int tk = 0; tile_at(tk, m_pair, n_blk); ++tk) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (p.dbg & 2) break;
            umma_bf16_cg2(d_tmem, umma_desc_k_sw128(a_addr + k * UMMA_K * 2), umma_desc_k_sw128(b_addr + k * UMMA_K * 2),
                          idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_mc2(&empty[stage], static_cast<uint16_t>((1u << CL) - 1));   // frees the slot in every CTA that loads into it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_mc2(&tfull[as], pair_mask);   // accumulator complete -> both CTAs' epilogues
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      // drain: the peer's epilogue arrives remotely on OUR tempty barriers; do not exit under them
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&tempty[as], aphase ^ 1);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    } else if (CL > 2 && !leader && lane == 0 && !(p.dbg & 4)) {
      // B forwarder of the non-leader CTA: its half of the B tile arrives as two multicast quarters on the local fullB
      // barrier; one remote arrival per stage tells the leader's MMA thread that this CTA's operands are complete.
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = mapa_u32(smem_u32(&full[0]), lead_rank);
      int m_pair = 0, n_blk = 0;
      for (int tk = 0; tile_at(tk, m_pair, n_blk); ++tk) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_expect_tx(&fullB[stage], C::B_BYTES);
          mbar_wait(&fullB[stage], phase);
          mbar_arrive_cluster(full0 + stage * 8);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- epilogue warps
    const int ew = warp - 2;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = ew >> 2;           // which half of the BLOCK_N columns
    uint8_t* ebuf = sE + ew * 2 * C::EBUF_BYTES;
    uint64_t* lb = lbar + ew * 2;
    const uint32_t tempty0 = mapa_u32(smem_u32(&tempty[0]), lead_rank);
    int as = 0;
    uint32_t aphase = 0;
    uint32_t g = 0;                     // chunks this warp has pushed through its two smem buffers
    float ln_mean = 0.f, ln_m2 = 0.f, ln_n = 0.f;      // fuse_ln: running statistics of this lane's row over this warp's columns of the tile
    // fuse_ln: the statistics of a finished tile are published one tile later (after the next tile's residual prefetch has been
    // issued: the wait for this warp's output stores to COMPLETE is off the critical path), and the warp normalises its own
    // 32 x BLOCK_N / 2 piece one tile after that: the other warps that own columns of these rows have published theirs by then
    // (they run the same tile round), and the rows -- stored with an evict_last L2 policy -- have not left L2 yet (normalising
    // three tiles later re-read them from HBM: +372 MB per fc2 launch in ncu)
    bool ln_pend = false;
    int ln_prow0 = 0, ln_pcol = 0;
    float ln_pmean = 0.f, ln_pm2 = 0.f;
    int lq_row0 = -1, lq_col0 = 0;                    // published, not yet normalised
    const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
    const int ln_parts = 2 * n_tiles;
    auto ln_publish = [&]() {
      // this lane's row, this warp's column half of its tile
      p.ln_stats[static_cast<size_t>(ln_prow0 + lane) * ln_parts + ln_pcol / (BLOCK_N / 2)] = make_float2(ln_pmean, ln_pm2);
      __threadfence();
      if (lane == 0) bulk_wait<0>();                // the stores of this warp's rows have completed (not just been read from smem)
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(p.ln_cnt + (ln_prow0 >> 5), 1);
      }
    };
    auto ln_normalise = [&](int row0, int col0) {
      // every owner of a piece of these 32 rows has published (normally long ago: no spinning)
      if (lane == 0) {
        const volatile int* cnt = p.ln_cnt + (row0 >> 5);
        for (int spin = 0; *cnt < ln_parts && spin < (1 << 22); ++spin) __nanosleep(200);     // bounded: a miscount shows as wrong rows, not a hang
        __threadfence();
      }
      __syncwarp();
      float mean = 0.f, m2 = 0.f;
      {
        const float2* sp = p.ln_stats + static_cast<size_t>(row0 + lane) * ln_parts;
        float2 part[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) part[j] = j < ln_parts ? __ldcg(sp + j) : make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) mean += part[j].x;
        mean /= static_cast<float>(ln_parts);
        const float cnt = static_cast<float>(BLOCK_N / 2);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < ln_parts) { const float dl = part[j].x - mean; m2 += part[j].y + cnt * dl * dl; }
      }
      const float rstd = rsqrtf(m2 / static_cast<float>(p.N) + p.ln_eps);
      // lane <-> 4 columns of the piece (BLOCK_N / 2 == 128 columns), 16 rows in flight
      static_assert(BLOCK_N != 256 || BLOCK_N / 2 == 128, "one float4 per lane and row");
      const float4 gm = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0) + lane);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0) + lane);
      const float4* src = reinterpret_cast<const float4*>(p.ln_src + static_cast<size_t>(row0) * p.ld_src + col0) + lane;
      uint2* dst = reinterpret_cast<uint2*>(p.ln_out + static_cast<size_t>(row0) * p.ld_ln + col0) + lane;
      const int rows = p.M - row0 < 32 ? p.M - row0 : 32;
#pragma unroll
      for (int rb = 0; rb < 32; rb += 16) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (rb + j < rows)
            v[j] = ld_cg_hint_f4(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + static_cast<size_t>(rb + j) * p.ld_src), pol_drop);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float mr = __shfl_sync(0xffffffffu, mean, rb + j), rs = __shfl_sync(0xffffffffu, rstd, rb + j);
          if (rb + j < rows) {
            uint2 o;
            o.x = pack_bf16(fmaf((v[j].x - mr) * rs, gm.x, bt.x), fmaf((v[j].y - mr) * rs, gm.y, bt.y));
            o.y = pack_bf16(fmaf((v[j].z - mr) * rs, gm.z, bt.z), fmaf((v[j].w - mr) * rs, gm.w, bt.w));
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + static_cast<size_t>(rb + j) * p.ld_ln) = o;
          }
        }
      }
    };
    auto ln_step = [&]() {          // at the top of a tile: publish the tile just finished, normalise the one before it
      ln_publish();
      ln_pend = false;
      if (lq_row0 >= 0) ln_normalise(lq_row0, lq_col0);
      lq_row0 = ln_prow0; lq_col0 = ln_pcol;
    };
    int m_pair = 0, n_blk = 0;
    for (int tk = 0; tile_at(tk, m_pair, n_blk); ++tk) {
      const int row0 = (m_pair * PAIRS + static_cast<int>(pair)) * 256 + static_cast<int>(rank) * 128 + quad * 32;
      const int col_base = n_blk * BLOCK_N + half * (BLOCK_N / 2);
      const bool live = row0 < p.M;     // warp-uniform
      if (EPI == EPI_RESID_F32 && live && lane == 0 && !(p.dbg & 1)) {
        bulk_wait_read<0>();            // both buffers are free of pending store reads
#pragma unroll
        for (int c = 0; c < 2 && c < C::CHUNKS; ++c) {
          const uint32_t b = (g + c) & 1;
          mbar_expect_tx(&lb[b], C::EBUF_BYTES);
          tma_load_2d(&tmRes, &lb[b], ebuf + b * C::EBUF_BYTES, col_base + c * 32, row0);
        }
      }
      if (EPI == EPI_RESID_F32 && ln_pend) ln_step();
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < C::CHUNKS; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 256 + half * (BLOCK_N / 2) + c * 32, r);
        uint4 zq[4];      // EPI_GELU_BWD: this lane's 32 pre-activations (64 contiguous bytes), in flight with the TMEM load
        if (EPI == EPI_GELU_BWD) {
          const int zrow = row0 + lane < p.M ? row0 + lane : p.M - 1;
          const uint4* zp = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(zrow) * p.ldaux + col_base + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) zq[i] = __ldg(zp + i);
        }
        tmem_ld_wait();
        if (c == C::CHUNKS - 1) {       // accumulator fully read: hand the TMEM stage back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty0 + as * 8);
        }
        if (!live || (p.dbg & 1)) continue;
        const uint32_t b = g & 1;
        uint8_t* buf = ebuf + b * C::EBUF_BYTES;
        const int col = col_base + c * 32;
        if (EPI == EPI_RESID_F32) {
          mbar_wait(&lb[b], (g >> 1) & 1);
        } else {
          if (lane == 0) bulk_wait_read<1>();   // the store that last used this buffer has finished reading it
          __syncwarp();
        }
        if (C::OUT_F32) {
          float4* row = reinterpret_cast<float4*>(buf + lane * 128);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                   __uint_as_float(r[4 * i + 3]));
            if (p.bias != nullptr) v = add4(v, __ldg(reinterpret_cast<const float4*>(p.bias + col) + i));
            float4* slot = row + (i ^ (lane & 7));
            if (EPI == EPI_RESID_F32) v = add4(v, *slot);
            *slot = v;
            if (EPI == EPI_RESID_F32) {      // keep the final values for the row statistics below
              r[4 * i] = __float_as_uint(v.x); r[4 * i + 1] = __float_as_uint(v.y);
              r[4 * i + 2] = __float_as_uint(v.z); r[4 * i + 3] = __float_as_uint(v.w);
            }
          }
          if (EPI == EPI_RESID_F32 && fuse_ln) {
            // statistics of these 32 values (two-pass in registers), merged into the running ones (Chan et al.): no E[x^2] - mean^2
            if (c == 0) { ln_mean = 0.f; ln_m2 = 0.f; ln_n = 0.f; }
            float sm = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sm += __uint_as_float(r[i]);
            const float mc = sm * (1.f / 32.f);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float t = __uint_as_float(r[i]) - mc; q = fmaf(t, t, q); }
            const float tot = ln_n + 32.f, delta = mc - ln_mean;
            ln_mean += delta * (32.f / tot);
            ln_m2 += q + delta * delta * (ln_n * 32.f / tot);
            ln_n = tot;
          }
        } else {
          uint4* row = reinterpret_cast<uint4*>(buf + lane * 64);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
            if (p.bias != nullptr) {      // packed fp32 pairs: one issue slot per two bias adds
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + 2 * i);
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + 2 * i + 1);
              const uint64_t s0 = f32x2_add(f32x2_pack_bits(r[8 * i], r[8 * i + 1]), f32x2_pack(b0.x, b0.y));
              const uint64_t s1 = f32x2_add(f32x2_pack_bits(r[8 * i + 2], r[8 * i + 3]), f32x2_pack(b0.z, b0.w));
              const uint64_t s2 = f32x2_add(f32x2_pack_bits(r[8 * i + 4], r[8 * i + 5]), f32x2_pack(b1.x, b1.y));
              const uint64_t s3 = f32x2_add(f32x2_pack_bits(r[8 * i + 6], r[8 * i + 7]), f32x2_pack(b1.z, b1.w));
              v[0] = f32x2_lo(s0); v[1] = f32x2_hi(s0); v[2] = f32x2_lo(s1); v[3] = f32x2_hi(s1);
              v[4] = f32x2_lo(s2); v[5] = f32x2_hi(s2); v[6] = f32x2_lo(s3); v[7] = f32x2_hi(s3);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 * i + j]);
            }
            if (EPI == EPI_GELU) {
              if (p.out2 != nullptr && row0 + lane < p.M) {      // pre-activation copy: 16 of this lane's 64 contiguous bytes
                uint4 zz;
                zz.x = pack_bf16(v[0], v[1]); zz.y = pack_bf16(v[2], v[3]); zz.z = pack_bf16(v[4], v[5]); zz.w = pack_bf16(v[6], v[7]);
                reinterpret_cast<uint4*>(p.out2 + static_cast<size_t>(row0 + lane) * p.ldaux + col)[i] = zz;
              }
              {
                // z * sigmoid(1.702 z) = z (0.5 + 0.5 tanh(0.851 z)): three roundings per element (0.851 z; 0.5 t + 0.5; z s), issued as
                // packed fp32 pairs (bit-identical to the scalar form, half the issue slots; fc1 alone and the whole step are
                // unchanged by it, 1248-1264 vs 1233-1256 TFLOP/s, gpurun s59: the epilogue is not what bounds this GEMM)
                const uint64_t k2 = f32x2_pack(0.851f, 0.851f), h2 = f32x2_pack(0.5f, 0.5f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint64_t z2 = f32x2_pack(v[2 * j], v[2 * j + 1]);
                  const uint64_t a2 = f32x2_mul(z2, k2);
                  const uint64_t s2 = f32x2_fma(f32x2_pack(tanh_approx(f32x2_lo(a2)), tanh_approx(f32x2_hi(a2))), h2, h2);
                  const uint64_t g2 = f32x2_mul(z2, s2);
                  v[2 * j] = f32x2_lo(g2); v[2 * j + 1] = f32x2_hi(g2);
                }
              }
            }
            if (EPI == EPI_GELU_BWD) {     // dz = dg * sigma(1.702 z) (1 + 1.702 z (1 - sigma(1.702 z)))
              const __nv_bfloat162* z2 = reinterpret_cast<const __nv_bfloat162*>(&zq[i]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 z = __bfloat1622float2(z2[j]);
                const float s0 = 0.5f + 0.5f * tanh_approx(0.851f * z.x), s1 = 0.5f + 0.5f * tanh_approx(0.851f * z.y);   // sigma(1.702 z), one MUFU op
                v[2 * j] *= s0 * (1.0f + 1.702f * z.x * (1.0f - s0));
                v[2 * j + 1] *= s1 * (1.0f + 1.702f * z.y * (1.0f - s1));
              }
            }
            uint4 o;
            o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
            row[i ^ ((lane >> 1) & 3)] = o;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (EPI == EPI_RESID_F32 && fuse_ln) tma_store_2d_hint(&tmOut, buf, col, row0, pol_keep);   // re-read below: stay in L2
          else tma_store_2d(&tmOut, buf, col, row0);
          bulk_commit();
          if (EPI == EPI_RESID_F32 && c + 2 < C::CHUNKS) {
            bulk_wait_read<0>();        // the store just issued must be done with the buffer before it is refilled
            mbar_expect_tx(&lb[b], C::EBUF_BYTES);
            tma_load_2d(&tmRes, &lb[b], buf, col + 64, row0);
          }
        }
        ++g;
      }
      if (EPI == EPI_RESID_F32 && fuse_ln && live && !(p.dbg & 1)) {
        ln_pend = true;
        ln_prow0 = row0; ln_pcol = col_base;
        ln_pmean = ln_mean; ln_pm2 = ln_m2;
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (EPI == EPI_RESID_F32 && fuse_ln) {
      if (ln_pend) ln_step();
      if (lq_row0 >= 0) ln_normalise(lq_row0, lq_col0);
    }
    if (lane == 0) bulk_wait<0>();      // all output tiles written before the CTA retires
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

thread_local char g_err[256] = "";
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
int g_pairs_hint = 0;   // co-resident CTA pairs reported by the occupancy API (0 until the first pair launch)
int g_quads_hint = 0;   // co-resident 4-CTA clusters

void set_err(const char* msg) { std::snprintf(g_err, sizeof(g_err), "%s", msg); }

EncodeTiledFn get_encode() {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

bool encode_map(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* ptr, int inner, int rows, int ld,
                int box_inner, int box_rows, CUtensorMapSwizzle swz) {
  const uint64_t dims[2] = {static_cast<uint64_t>(inner), static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {static_cast<uint64_t>(ld) * elem_bytes};
  const uint32_t box[2] = {static_cast<uint32_t>(box_inner), static_cast<uint32_t>(box_rows)};
  return encode_tiled_map(m, dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 1 : 0, ptr, 2, dims, strides, box,
                          swz == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : (swz == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : 0));
}

bool make_map(CUtensorMap* m, const GemmOperand& op, int box_rows) {
  return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, op.ptr, op.k, op.rows, op.ld, BLOCK_K, box_rows,
                    CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BLOCK_N, int EPI, int CL>
cudaError_t launch2_t(const GemmArgs& g, cudaStream_t stream, int num_sms) {
  using C = Cfg2<BLOCK_N, EPI>;
  CUtensorMap tA1, tB1, tA2, tB2, tOut, tRes;
  if (!make_map(&tA1, g.a1, 128) || !make_map(&tB1, g.b1, BLOCK_N / CL)) return cudaErrorInvalidValue;
  const bool two = g.a2.ptr != nullptr && g.a2.k > 0;
  if (two) {
    if (!make_map(&tA2, g.a2, 128) || !make_map(&tB2, g.b2, BLOCK_N / CL)) return cudaErrorInvalidValue;
  } else {
    tA2 = tA1;
    tB2 = tB1;
  }
  if (C::OUT_F32) {
    if (!encode_map(&tOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.out, g.N, g.M, g.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
      return cudaErrorInvalidValue;
  } else {
    if (!encode_map(&tOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out, g.N, g.M, g.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))
      return cudaErrorInvalidValue;
  }
  if (EPI == EPI_RESID_F32) {
    if (!encode_map(&tRes, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.resid, g.N, g.M, g.ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
      return cudaErrorInvalidValue;
  } else {
    tRes = tOut;
  }
  Epi2Params p;
  p.M = g.M; p.N = g.N;
  p.kb1 = g.a1.k / BLOCK_K;
  p.kb2 = two ? g.a2.k / BLOCK_K : 0;
  p.bias = g.bias;
  static const char* dbg_env = std::getenv("TTL_GEMM_DBG");
  p.dbg = dbg_env ? std::atoi(dbg_env) : 0;
  p.rev = g.descending;
  p.aux = g.aux; p.ldaux = g.ldo;
  p.out2 = EPI == EPI_GELU ? static_cast<__nv_bfloat16*>(g.out2) : nullptr;
  p.ln_gamma = g.ln_gamma; p.ln_beta = g.ln_beta; p.ln_out = g.ln_out; p.ld_ln = g.ld_ln; p.ln_eps = g.ln_eps;
  p.ln_stats = reinterpret_cast<float2*>(g.ln_stats); p.ln_cnt = g.ln_cnt;
  p.ln_src = static_cast<const float*>(g.out); p.ld_src = g.ldo;
  if (g.ln_out != nullptr && (EPI != EPI_RESID_F32 || CL != 2 || BLOCK_N != 256 || g.ln_gamma == nullptr || g.ln_beta == nullptr ||
                              g.ln_stats == nullptr || g.ln_cnt == nullptr || g.ld_ln % 4 != 0 || g.ldo % 4 != 0 ||
                              g.N % 256 != 0 || g.N > 1024)) {
    set_err("gemm2: fused LayerNorm needs the residual epilogue on CTA pairs, BLOCK_N = 256, N <= 1024, gamma / beta / stats / counters");
    return cudaErrorInvalidValue;
  }
  if (g.out2 != nullptr && EPI != EPI_GELU) { set_err("gemm2: out2 only with the QuickGELU epilogue"); return cudaErrorInvalidValue; }
  if (EPI == EPI_GELU_BWD && (g.aux == nullptr || g.ldo % 8 != 0)) { set_err("gemm2: EPI_GELU_BWD needs aux (z), ld % 8 == 0"); return cudaErrorInvalidValue; }
  auto kern = gemm2_kernel<BLOCK_N, EPI, CL>;
  const int dv = current_device_slot();
  static bool attr_done[MAX_DEVICES] = {};  // per instantiation and device
  if (!attr_done[dv]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_err("cudaFuncSetAttribute(max dynamic smem) failed (gemm2)"); return e; }
    attr_done[dv] = true;
  }
  // A persistent grid must be co-resident: not every TPC of a B200 has both SMs enabled, so the number of CTA pairs
  // that fit at once can be below num_sms / 2 -- ask the occupancy API once per instantiation.
  static int max_pairs_dev[MAX_DEVICES] = {};
  int& max_pairs = max_pairs_dev[dv];
  if (max_pairs == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms / CL * CL);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / CL; }
    max_pairs = n < num_sms / CL ? n : num_sms / CL;
    if (CL == 2) g_pairs_hint = max_pairs;
    else g_quads_hint = max_pairs;
    if (std::getenv("TTL_DEBUG")) std::fprintf(stderr, "ttl: gemm2<%d,%d,%d> co-resident clusters: %d (SMs %d)\n", BLOCK_N, EPI, CL, n, num_sms);
  }
  const int tiles = (((g.M + 255) / 256 + CL / 2 - 1) / (CL / 2)) * (g.N / BLOCK_N);
  int clusters = tiles < max_pairs ? tiles : max_pairs;
  if (g.max_clusters > 0 && g.max_clusters < clusters) clusters = g.max_clusters;
  const int grid = CL * clusters;
  return launch_pdl(kern, dim3(grid), dim3(G2_THREADS), C::SMEM_BYTES, stream, tA1, tB1, tA2, tB2, tOut, tRes, p);
}

template <int BLOCK_N, int CL = 2>
cudaError_t launch2_n(const GemmArgs& g, cudaStream_t s, int sms) {
  switch (g.epi) {
    case EPI_BF16: return launch2_t<BLOCK_N, EPI_BF16, CL>(g, s, sms);
    case EPI_GELU: return launch2_t<BLOCK_N, EPI_GELU, CL>(g, s, sms);
    case EPI_RESID_F32: return launch2_t<BLOCK_N, EPI_RESID_F32, CL>(g, s, sms);
    case EPI_F32: return launch2_t<BLOCK_N, EPI_F32, CL>(g, s, sms);
    case EPI_GELU_BWD: return launch2_t<BLOCK_N, EPI_GELU_BWD, CL>(g, s, sms);
    default: set_err("gemm2: epilogue not supported by the CTA-pair kernel"); return cudaErrorInvalidValue;
  }
}

// The CTA-pair kernel covers the plain epilogues (no second output, no scatter); BLOCK_N by a per-k-block cost model:
// rounds of the persistent schedule x bytes a pair pulls from L2 per k-block (the kernel is L2->SM bound).
int pick_pair_block_n(const GemmArgs& g, int num_sms) {
  if ((g.out2 != nullptr && g.epi != EPI_GELU) || g.M < 1024) return 0;
  if (g.epi != EPI_BF16 && g.epi != EPI_GELU && g.epi != EPI_RESID_F32 && g.epi != EPI_F32 && g.epi != EPI_GELU_BWD) return 0;
  if (g.ldo % 4 != 0 || (g.epi == EPI_RESID_F32 && (g.resid == nullptr || g.ldr % 4 != 0))) return 0;
  const int pairs = g_pairs_hint > 0 ? g_pairs_hint : num_sms / 2, m_pairs = (g.M + 255) / 256;
  int best = 0;
  double best_cost = 0;
  const int cands[3] = {256, 192, 128};
  for (int bn : cands) {
    if (g.N % bn != 0) continue;
    const int tiles = m_pairs * (g.N / bn);
    const int rounds = (tiles + pairs - 1) / pairs;
    const double cost = rounds * (256.0 + bn);
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

template <int BLOCK_N, int EPI>
cudaError_t launch_t(const GemmArgs& g, cudaStream_t stream, int num_sms) {
  using C = Cfg<BLOCK_N>;
  CUtensorMap tA1, tB1, tA2, tB2;
  if (!make_map(&tA1, g.a1, BLOCK_M) || !make_map(&tB1, g.b1, BLOCK_N)) return cudaErrorInvalidValue;
  const bool two = g.a2.ptr != nullptr && g.a2.k > 0;
  if (two) {
    if (!make_map(&tA2, g.a2, BLOCK_M) || !make_map(&tB2, g.b2, BLOCK_N)) return cudaErrorInvalidValue;
  } else {
    tA2 = tA1;
    tB2 = tB1;
  }
  EpiParams p;
  p.M = g.M; p.N = g.N;
  p.kb1 = g.a1.k / BLOCK_K;
  p.kb2 = two ? g.a2.k / BLOCK_K : 0;
  p.bias = g.bias; p.out = g.out; p.ldo = g.ldo; p.out2 = g.out2;
  p.resid = g.resid; p.ldr = g.ldr; p.aux = g.aux; p.pos = g.pos; p.tpv = g.tokens_per_view;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, EPI>;
  static bool attr_done[MAX_DEVICES] = {};  // per instantiation and device
  const int dv = current_device_slot();
  if (!attr_done[dv]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_err("cudaFuncSetAttribute(max dynamic smem) failed"); return e; }
    attr_done[dv] = true;
  }
  const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles = m_tiles * (g.N / BLOCK_N);
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), C::SMEM_BYTES, stream, tA1, tB1, tA2, tB2, p);
}

template <int BLOCK_N>
cudaError_t launch_n(const GemmArgs& g, cudaStream_t s, int sms) {
  switch (g.epi) {
    case EPI_BF16: return launch_t<BLOCK_N, EPI_BF16>(g, s, sms);
    case EPI_GELU: return launch_t<BLOCK_N, EPI_GELU>(g, s, sms);
    case EPI_RESID_F32: return launch_t<BLOCK_N, EPI_RESID_F32>(g, s, sms);
    case EPI_PATCH_F32: return launch_t<BLOCK_N, EPI_PATCH_F32>(g, s, sms);
    case EPI_F32: return launch_t<BLOCK_N, EPI_F32>(g, s, sms);
    case EPI_GELU_BWD: return launch_t<BLOCK_N, EPI_GELU_BWD>(g, s, sms);
    default: set_err("unknown epilogue"); return cudaErrorInvalidValue;
  }
}

}  // namespace

bool encode_tiled_map(void* map_out, int dtype, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_err("cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                    : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  CUresult r = enc(static_cast<CUtensorMap*>(map_out), dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[240];
    std::snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d): ptr=%p rank=%d dims=%llu,%llu box=%u,%u swizzle=%d", int(r),
                  ptr, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], swizzle_bytes);
    set_err(msg);
    return false;
  }
  return true;
}

const char* gemm_last_error() { return g_err; }

cudaError_t gemm_launch(const GemmArgs& g, cudaStream_t stream, int num_sms) {
  g_err[0] = 0;
  if (g.M <= 0 || g.N <= 0 || g.N % 64 != 0 || g.a1.k <= 0 || g.a1.k % BLOCK_K != 0 || g.a1.k != g.b1.k ||
      (g.a2.ptr && (g.a2.k % BLOCK_K != 0 || g.a2.k != g.b2.k)) || g.ldo % 8 != 0 || g.a1.ld % 8 != 0 ||
      g.b1.ld % 8 != 0 || g.out == nullptr) {
    set_err("gemm_launch: invalid shape/alignment");
    return cudaErrorInvalidValue;
  }
  if (g.a1.rows < g.M || g.b1.rows < g.N) { set_err("gemm_launch: operand rows smaller than M/N"); return cudaErrorInvalidValue; }
  int bn = g.force_block_n;
  if (g.ln_out != nullptr) {      // fused LayerNorm: CTA pairs, BLOCK_N = 256 (checked again in launch2_t)
    if (g.epi != EPI_RESID_F32 || g.N % 256 != 0 || g.M < 1024 || g.resid == nullptr) {
      set_err("gemm_launch: fused LayerNorm needs EPI_RESID_F32, N % 256 == 0, M >= 1024");
      return cudaErrorInvalidValue;
    }
    if (g.ln_cnt == nullptr) { set_err("gemm_launch: fused LayerNorm needs the arrival counters"); return cudaErrorInvalidValue; }
    cudaError_t e = cudaMemsetAsync(g.ln_cnt, 0, static_cast<size_t>((g.M + 31) / 32) * sizeof(int), stream);
    if (e != cudaSuccess) { set_err("gemm_launch: clearing the fused-LayerNorm counters failed"); return e; }
    return launch2_t<256, EPI_RESID_F32, 2>(g, stream, num_sms);
  }
  {
    // force_block_n: 0 = heuristic; 64/128/256 = 1-CTA kernel; 1000 + {128,192,256} = CTA-pair kernel;
    // 2256 = 4-CTA cluster (two pairs, multicast B tiles), BLOCK_N = 256
    // 3192 = 6-CTA cluster (three pairs), BLOCK_N = 192
    if (bn == 2256) {
      if (g.N % 256 != 0 || g.out2 != nullptr) { set_err("gemm_launch: 4-CTA cluster kernel needs N % 256 == 0, no out2"); return cudaErrorInvalidValue; }
      return launch2_n<256, 4>(g, stream, num_sms);
    }
    if (bn == 3192) {
      if (g.N % 192 != 0 || g.out2 != nullptr) { set_err("gemm_launch: 6-CTA cluster kernel needs N % 192 == 0, no out2"); return cudaErrorInvalidValue; }
      return launch2_n<192, 6>(g, stream, num_sms);
    }
    int bn2 = bn >= 1000 ? bn - 1000 : (bn == 0 ? pick_pair_block_n(g, num_sms) : 0);
    static const char* env = std::getenv("TTL_GEMM_PAIR");
    if (bn == 0 && env != nullptr) {
      const int v = std::atoi(env);
      if (v == 0) bn2 = 0;
      else if (v > 1 && bn2 != 0 && g.N % v == 0) bn2 = v;
    }
    if (bn2 != 0) {
      if (g.N % bn2 != 0 || (g.out2 != nullptr && g.epi != EPI_GELU)) { set_err("gemm_launch: CTA-pair kernel: bad BLOCK_N / out2"); return cudaErrorInvalidValue; }
      // 4-CTA clusters with multicast B tiles: TTL_GEMM_CLUSTER=4 (BLOCK_N = 256, at least two row blocks of 256).  Opt-in:
      // measured on B200 (gpurun s59/s60) only 33 clusters of 4 are co-resident (132 of 148 SMs): +6-7 % per SM from the
      // smaller L2->SM operand traffic, but -4 % per kernel and -2 % per adapted sample with 16 SMs stranded.
      static const char* cl_env = std::getenv("TTL_GEMM_CLUSTER");
      static const int cl = cl_env ? std::atoi(cl_env) : 2;
      if (cl == 4 && bn2 == 256 && g.M > 2048) return launch2_n<256, 4>(g, stream, num_sms);
      if (cl == 6 && g.N % 192 == 0 && g.M > 2048) return launch2_n<192, 6>(g, stream, num_sms);
      switch (bn2) {
        case 256: return launch2_n<256>(g, stream, num_sms);
        case 192: return launch2_n<192>(g, stream, num_sms);
        case 128: return launch2_n<128>(g, stream, num_sms);
        default: set_err("gemm_launch: CTA-pair BLOCK_N must be 128/192/256"); return cudaErrorInvalidValue;
      }
    }
  }
  if (bn == 0) {
    const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
    if (g.N % 256 == 0 && m_tiles * (g.N / 256) >= num_sms) bn = 256;
    else if (g.N % 128 == 0 && m_tiles * (g.N / 128) >= num_sms) bn = 128;
    else bn = 64;
  }
  if (g.N % bn != 0) { set_err("gemm_launch: N not a multiple of BLOCK_N"); return cudaErrorInvalidValue; }
  switch (bn) {
    case 256: return launch_n<256>(g, stream, num_sms);
    case 128: return launch_n<128>(g, stream, num_sms);
    case 64: return launch_n<64>(g, stream, num_sms);
    default: set_err("gemm_launch: BLOCK_N must be 64/128/256"); return cudaErrorInvalidValue;
  }
}

}  // namespace ttl
