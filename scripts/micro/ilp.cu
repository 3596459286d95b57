// Single-warp issue rate on sm_100a as a function of instruction-level parallelism: K independent DFMA chains in one warp,
// and the same with 1, 2, 4 warps per SM sub-partition. cycles per DFMA (warp-instruction) as seen by one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/micro/ilp scripts/micro/ilp.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int K, int OP> __global__ void k(double* out, long long* cyc, double b0) {
  double a[K];
#pragma unroll
  for (int i = 0; i < K; i++) a[i] = 1.0 + threadIdx.x + i;
  const double b = b0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int n = 0; n < N; n++) {
#pragma unroll
    for (int i = 0; i < K; i++) {
      if (OP == 0) a[i] = __fma_rn(a[i], b, b);
      if (OP == 1) a[i] = (a[i] > b) ? a[i] : b + (double)i;  // DSETP + 2 FSEL
      if (OP == 2) { int h = __double2hiint(a[i]); h = (h ^ 0x1234567) + i; a[i] = __hiloint2double(h, __double2loint(a[i])); }  // ALU
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; i++) s += a[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int K, int OP> void run(const char* nm, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 8);
  k<K, OP><<<1, threads>>>(out, cyc, 0.9999999); cudaDeviceSynchronize();
  k<K, OP><<<1, threads>>>(out, cyc, 0.9999999); cudaDeviceSynchronize();
  printf("%-10s K=%d warps/SM=%2d : %6.2f cycles per op per warp\n", nm, K, threads / 32, (double)*cyc / (N * K));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 0>("DFMA", 32); run<2, 0>("DFMA", 32); run<4, 0>("DFMA", 32); run<8, 0>("DFMA", 32);
  run<8, 0>("DFMA", 128); run<8, 0>("DFMA", 256); run<8, 0>("DFMA", 512);
  run<1, 1>("DSETP+SEL", 32); run<4, 1>("DSETP+SEL", 32); run<8, 1>("DSETP+SEL", 32);
  run<1, 2>("ALU x2", 32); run<8, 2>("ALU x2", 32);
  return 0;
}
