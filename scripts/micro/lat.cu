// Dependent-issue latencies on sm_100a of the instructions the Newton kernels' critical path is made of (one warp, clock64).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -o /tmp/lat scripts/micro/lat.cu && /tmp/lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void k(double* out, long long* cyc, double a0, double b0) {
  double a = a0 + threadIdx.x, b = b0;
  __shared__ double sm[1024];
  sm[threadIdx.x] = a0; sm[threadIdx.x + 32] = 0.0;
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) {
    if (OP == 0) a = __fma_rn(a, b, b);
    if (OP == 1) a = __dadd_rn(a, b);
    if (OP == 2) a = __dmul_rn(a, b);
    if (OP == 3) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 8) & 31);
    if (OP == 4) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r; }
    if (OP == 5) { idx = (int)sm[idx & 63] + (idx & 31); a += idx; }
    if (OP == 6) a = sm[(__double2loint(a) & 31)];
    if (OP == 7) a = (a > b) ? a : __dmul_rn(a, b);
    if (OP == 8) a = a / b;
    if (OP == 9) a = sqrt(a);
    if (OP == 10) a = exp(a * 1e-3);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a + idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 256 * 8); cudaMallocManaged(&cyc, 8);
  const char* nm[] = {"DFMA", "DADD", "DMUL", "SHFL f64 (2x SHFL.IDX)", "MUFU.RCP64H", "LDS->cvt->index", "LDS f64 dependent", "DSETP+select+DMUL", "a/b (compiler)", "sqrt", "exp"};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); printf("%-28s %7.1f cycles per dependent op\n", nm[OP], (double)*cyc / N);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
  return 0;
}
