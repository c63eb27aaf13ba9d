// Micro-benchmark of the instruction mix of the attention kernel's pass 2 WITHOUT TMEM, mbarriers or MMAs:
// 8 warps per SM (two per scheduler), each thread turns 32 fp32 scores into bf16 probabilities + row sum and stores them into
// the swizzled P layout, chunk after chunk.  Prints cycles per 32-column chunk (both warps of a scheduler together), to be put
// against 4.4 k cycles / 6.5 chunks = 680 in the real kernel.   nvcc -arch=sm_100a -O3 -o pass2_mix pass2_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float xf = x + 12582912.f;
  const float fr = x - (xf - 12582912.f);
  float p = fmaf(fr, 0.0551716685f, 0.2426111251f);
  p = fmaf(p, fr, 0.6932609677f);
  p = fmaf(p, fr, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// MASK: which of every 8 elements use the polynomial; FLAGS bit 0: no stores, bit 1: no row sums, bit 2: no exponentials
template <int MASK, int FLAGS>
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, long long* cycles, float scale_log2) {
  extern __shared__ uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = warp >> 2, row = (warp & 3) * 32 + lane;
  const uint32_t p_blk0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + t * 53248;
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = -0.01f * ((threadIdx.x * 7 + i * 13) % 97);
  float l0 = 0.f, l1 = 0.f;
  const float ms = 0.25f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int c = it % 6;
    const int blk = c >> 1;
    const uint32_t base = p_blk0 + blk * 16384 + row * 128;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = fmaf(r[q4 * 8 + i], scale_log2, -ms);
        e[i] = (FLAGS & 4) ? x : (((MASK >> i) & 1) ? ex2_poly(x) : ex2_approx(x));
      }
      if (!(FLAGS & 2)) {
        l0 += (e[0] + e[1]) + (e[2] + e[3]);
        l1 += (e[4] + e[5]) + (e[6] + e[7]);
      }
      const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q4);
      if (!(FLAGS & 1))
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + ((chunk ^ (row & 7)) << 4)),
                     "r"(pack_bf16(e[0], e[1])), "r"(pack_bf16(e[2], e[3])), "r"(pack_bf16(e[4], e[5])),
                     "r"(pack_bf16(e[6], e[7]))
                     : "memory");
      else
        l1 += __uint_as_float(pack_bf16(e[0], e[1]) ^ pack_bf16(e[2], e[3]) ^ pack_bf16(e[4], e[5]) ^ pack_bf16(e[6], e[7])) * 1e-30f;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] -= 1e-4f;      // stands for the next chunk's scores (1 FADD per element extra)
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MASK, int FLAGS>
void run(const char* name) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 256);
  cudaMallocManaged(&cyc, sizeof(long long) * sms);
  const int iters = 6 * 400;
  const size_t smem = 2 * 53248 + 100 * 1024;     // one CTA per SM
  cudaFuncSetAttribute(k<MASK, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) k<MASK, FLAGS><<<sms, 256, smem>>>(out, iters, cyc, 0.18f);
  cudaError_t e = cudaDeviceSynchronize();
  double c = 0; for (int i = 0; i < sms; ++i) c += cyc[i]; c /= sms;
  printf("%-52s %7.1f cycles per 32-column chunk (2 warps per scheduler)  %s\n", name, c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0x88, 0>("as in the kernel (2 of 8 polynomial)");
  run<0x00, 0>("all MUFU");
  run<0x80, 0>("1 of 8 polynomial");
  run<0xAA, 0>("4 of 8 polynomial");
  run<0xFF, 0>("all polynomial");
  run<0x88, 1>("2 of 8 polynomial, no stores");
  run<0x88, 2>("2 of 8 polynomial, no row sums");
  run<0x88, 3>("2 of 8 polynomial, no stores, no row sums");
  run<0x88, 4>("no exponentials");
  run<0x00, 2>("all MUFU, no row sums");
  return 0;
}
