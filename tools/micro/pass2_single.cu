// Micro-benchmark: the attention forward's pass-2 instruction mix with ONE warp per scheduler (the decoupled-stream kernel,
// attention_fwd_pt_kernel) against two, in several formulations.  No TMEM / mbarriers / MMAs: registers only.
//   nvcc -arch=sm_100a -O3 -o pass2_single pass2_single.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx_v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
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
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return (static_cast<uint64_t>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ float lo32(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float hi32(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }

// MODE 0: as in the kernel (compiler-scheduled).  MODE 1: f32x2 scale FMA + f32x2 row sums.  MODE 2: volatile MUFU in program
// order, each followed by the pack/sum work of an EARLIER element (manual interleave).  MODE 3: no exponentials (mix floor).
template <int MASK, int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, int iters, long long* cycles, float scale_log2, uint32_t* sink) {
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = -0.01f * ((threadIdx.x * 7 + i * 13) % 97);
  float l0 = 0.f, l1 = 0.f;
  uint64_t l2 = 0;
  uint32_t acc = 0;
  const float ms = 0.25f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
    if (MODE == 0 || MODE == 3) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = fmaf(r[q4 * 8 + i], scale_log2, -ms);
          e[i] = MODE == 3 ? x : (((MASK >> i) & 1) ? ex2_poly(x) : ex2_approx(x));
        }
        l0 += (e[0] + e[1]) + (e[2] + e[3]);
        l1 += (e[4] + e[5]) + (e[6] + e[7]);
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[q4 * 4 + i] = pack_bf16(e[2 * i], e[2 * i + 1]);
      }
    } else if (MODE == 1) {
      const uint64_t sc2 = pack2(scale_log2, scale_log2), ms2 = pack2(-ms, -ms);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint64_t e2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint64_t x2 = ffma2(pack2(r[q4 * 8 + 2 * i], r[q4 * 8 + 2 * i + 1]), sc2, ms2);
          const float a = (((MASK >> (2 * i)) & 1) ? ex2_poly(lo32(x2)) : ex2_approx(lo32(x2)));
          const float b = (((MASK >> (2 * i + 1)) & 1) ? ex2_poly(hi32(x2)) : ex2_approx(hi32(x2)));
          e2[i] = pack2(a, b);
          pk[q4 * 4 + i] = pack_bf16(a, b);
        }
        l2 = fadd2(l2, fadd2(fadd2(e2[0], e2[1]), fadd2(e2[2], e2[3])));
      }
    } else if (MODE == 2) {
      float x[32], e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaf(r[i], scale_log2, -ms);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        e[i] = ex2_approx_v(x[i]);
        if (i >= 8 && (i & 1)) {          // work of elements i - 8, i - 7 rides behind this MUFU
          const int j = i - 9;
          pk[j >> 1] = pack_bf16(e[j], e[j + 1]);
          if (j & 2) l1 += e[j] + e[j + 1]; else l0 += e[j] + e[j + 1];
          asm volatile("" :: "r"(pk[j >> 1]), "f"(l0), "f"(l1));
        }
      }
#pragma unroll
      for (int j = 24; j < 32; j += 2) {
        pk[j >> 1] = pack_bf16(e[j], e[j + 1]);
        if (j & 2) l1 += e[j] + e[j + 1]; else l0 += e[j] + e[j + 1];
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];      // stands for the tcgen05.st (1 LOP3 per 2 elements)
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] -= 1e-4f;      // stands for the next chunk's scores (1 FADD per element extra)
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + lo32(l2) + hi32(l2);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MASK, int MODE>
void run(const char* name, int warps) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc; uint32_t* sink;
  cudaMalloc(&out, sizeof(float) * sms * 256);
  cudaMalloc(&sink, sizeof(uint32_t) * sms * 256);
  cudaMallocManaged(&cyc, sizeof(long long) * sms);
  const int iters = 6 * 400;
  for (int rep = 0; rep < 2; ++rep) k<MASK, MODE><<<sms, warps * 32>>>(out, iters, cyc, 0.18f, sink);
  cudaError_t e = cudaDeviceSynchronize();
  double c = 0; for (int i = 0; i < sms; ++i) c += cyc[i]; c /= sms;
  printf("%-44s %d warp(s) per scheduler: %7.1f cycles per 32-column chunk and warp set  %s\n", name, warps / 4, c / iters,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int w : {4, 8}) {
    run<0x00, 0>("all MUFU, compiler order", w);
    run<0x80, 0>("1 of 8 polynomial, compiler order", w);
    run<0x88, 0>("2 of 8 polynomial, compiler order", w);
    run<0xAA, 0>("4 of 8 polynomial, compiler order", w);
    run<0x00, 1>("all MUFU, f32x2 scale + row sums", w);
    run<0x88, 1>("2 of 8 polynomial, f32x2 scale + row sums", w);
    run<0x00, 2>("all MUFU, manual interleave", w);
    run<0x00, 3>("no exponentials", w);
  }
  return 0;
}
