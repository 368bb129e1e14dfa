// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, one CTA per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a x2_rate.cu -o x2_rate && ./x2_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
template <int MODE>
__global__ void kern(float *out, int iters, long long *cyc) {
  float a[8], b = 1.0001f, c = 0.5f;
  unsigned long long A[8], B = pk(1.0001f, 1.0002f), C = pk(0.5f, 0.25f);
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; A[i] = pk(threadIdx.x + i, i); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], b, c);
      else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B), "l"(C));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) { s += a[i]; float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(A[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) kern<0><<<148, warps * 32>>>(out, iters, cyc); else kern<1><<<148, warps * 32>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double inst = (double)iters * 8 * warps;           // warp instructions per SM
      printf("%s warps/SM %2d: %lld cycles, %.3f warp-instr/clk/SM (%.3f per SMSP), %.1f FMA lanes/clk/SM\n",
             mode ? "FFMA2" : "FFMA ", warps, h, inst / h, inst / h / 4, inst / h * 32 * (mode ? 2 : 1));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
