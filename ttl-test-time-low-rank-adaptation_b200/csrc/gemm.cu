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

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

thread_local char g_err[256] = "";
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;

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

bool make_map(CUtensorMap* m, const GemmOperand& op, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_err("cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(op.k), static_cast<cuuint64_t>(op.rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(op.ld) * 2};
  cuuint32_t box[2] = {BLOCK_K, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(op.ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[200];
    std::snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%d k=%d ld=%d box_rows=%d", int(r),
                  (const void*)op.ptr, op.rows, op.k, op.ld, box_rows);
    set_err(b);
    return false;
  }
  return true;
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
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_err("cudaFuncSetAttribute(max dynamic smem) failed"); return e; }
    attr_done = true;
  }
  const int m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles = m_tiles * (g.N / BLOCK_N);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(tA1, tB1, tA2, tB2, p);
  return cudaGetLastError();
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
