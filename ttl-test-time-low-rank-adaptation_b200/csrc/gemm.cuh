// Host-side interface of the tcgen05 GEMM (gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace ttl {

// Epilogues fused into the GEMM (applied to acc = sum_k A[m,k]*B[n,k], fp32, straight out of TMEM).
enum GemmEpi : int {
  EPI_BF16 = 0,      // out(bf16)[m,n] = acc + bias[n]
  EPI_GELU = 1,      // z = acc + bias; out2(bf16, optional) = z; out(bf16) = z*sigmoid(1.702 z)   (QuickGELU)
  EPI_RESID_F32 = 2, // out(f32)[m,n] = resid(f32)[m,n] + acc + bias[n]
  EPI_PATCH_F32 = 3, // patch-embed: row m = (view, patch) -> token row view*(T+1)+1+patch; out(f32) = acc + pos[1+patch, n]
  EPI_F32 = 4,       // out(f32)[m,n] = acc + bias[n]
  EPI_GELU_BWD = 5,  // out(bf16) = acc * dQuickGELU(aux(bf16)[m,n])
  EPI_COUNT = 6
};

struct GemmOperand {
  const __nv_bfloat16* ptr = nullptr;  // row-major [rows, k], k contiguous ("K-major")
  int rows = 0;
  int k = 0;      // multiple of 64
  int ld = 0;     // elements; multiple of 8
};

struct GemmArgs {
  // C[M,N] = A1[M,K1]*B1[N,K1]^T (+ A2[M,K2]*B2[N,K2]^T)  -- the optional second pair is the rank-r
  // LoRA extension riding in the same TMEM accumulator.
  GemmOperand a1, b1, a2, b2;
  int M = 0, N = 0;
  int epi = EPI_BF16;
  const float* bias = nullptr;
  void* out = nullptr;
  int ldo = 0;
  void* out2 = nullptr;   // EPI_GELU: pre-activation z (bf16, ld = ldo);  EPI_RESID_F32/EPI_F32: bf16 copy (ld = ldo)
  const float* resid = nullptr;
  int ldr = 0;
  const __nv_bfloat16* aux = nullptr;  // EPI_GELU_BWD: z, ld = ldo
  const float* pos = nullptr;          // EPI_PATCH_F32: [T+1, N]
  int tokens_per_view = 0;             // EPI_PATCH_F32: T (patches per view)
  int force_block_n = 0;               // 0 = heuristic
  int max_clusters = 0;                // CTA-pair / cluster kernels: cap on the persistent grid (0 = every co-resident cluster)
  int descending = 0;                  // CTA-pair kernel: walk the row blocks from the last to the first (engine.cu zigzag)
  // EPI_RESID_F32 on the CTA-pair kernel: also write LayerNorm(out row) * gamma + beta as bf16 [M, N] (row pitch ld_ln) -- the next
  // LayerNorm fused into this GEMM.  Every epilogue warp publishes the statistics of its 32 rows x BLOCK_N / 2 columns to
  // ln_stats ([M][2 N / BLOCK_N] float2) and counts an arrival in ln_cnt[row / 32] (cleared by the launcher); one tile later it
  // normalises its own piece from L2 with the statistics of the complete rows (gemm.cu).
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  __nv_bfloat16* ln_out = nullptr;
  int ld_ln = 0;
  float ln_eps = 1e-5f;
  float* ln_stats = nullptr;           // 2 * (M rounded up to 32) * (2 N / BLOCK_N) floats
  int* ln_cnt = nullptr;               // (M + 31) / 32 counters
};

// Generic tiled tensor-map encoder (cuTensorMapEncodeTiled through the runtime's driver entry point), shared with the
// TMA-fed attention kernel.  dtype: 0 = bf16, 1 = fp32.  swizzle_bytes: 0 / 64 / 128.  dims/box innermost first;
// strides_bytes has rank-1 entries (stride of dims 1..rank-1).  Returns false and sets gemm_last_error() on failure.
bool encode_tiled_map(void* map_out, int dtype, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// Returns cudaSuccess or an error; never synchronises.
cudaError_t gemm_launch(const GemmArgs& g, cudaStream_t stream, int num_sms);
const char* gemm_last_error();

}  // namespace ttl
