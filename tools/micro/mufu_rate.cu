// Micro-benchmark: MUFU.EX2 (ex2.approx.ftz.f32) and FFMA issue rates per SM and clock on this GPU, to put the attention
// kernel's pass 2 against a measured ceiling.  nvcc -arch=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cycles) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i + 1);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * warps * 32);
  cudaMallocManaged(&cyc, sizeof(long long) * sms);
  const int iters = 4096;
  k<MODE><<<sms, warps * 32>>>(out, iters, cyc);
  k<MODE><<<sms, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  double c = 0; for (int i = 0; i < sms; ++i) c += cyc[i]; c /= sms;
  const double ops = double(iters) * 8 * warps * 32 * (MODE == 2 ? 1 : 1);
  printf("%-28s %2d warps/SM: %.2f %s per clock and SM\n", name, warps, ops / c, MODE == 1 ? "FFMA lanes" : "ex2 lanes");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>("MUFU.EX2", w);
  for (int w : {4, 8, 16, 32}) run<1>("FFMA", w);
  for (int w : {8, 16}) run<2>("MUFU.EX2 + FFMA interleaved", w);
  return 0;
}
