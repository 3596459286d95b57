// FP64 throughput of one B200, measured: the denominator of the "fp64" roofline that bench.py reports for device
// evaluation (north_star: "FP64 peak for device evaluation"). All SMs, WARPS warps per SM sub-partition, K independent
// DFMA chains per thread; also DADD/DMUL (the product is built -fmad=false: the reference never contracts a*b+c, so its
// arithmetic issues as separate DMUL and DADD and the attainable flop rate of the product's text is the DMUL/DADD issue
// rate, half the DFMA flop peak), MUFU.RCP64H and the full-precision division as the library emits it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/fp64_peak scripts/micro/fp64_peak.cu
//   scripts/micro/fp64_peak > profiles/fp64_peak.json
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>
#define N 4096
enum { OP_DFMA, OP_DMUL_DADD, OP_DADD, OP_DIV };
template <int K, int OP> __global__ void __launch_bounds__(1024) k(double* out, double b0, double c0) {
  double a[K];
#pragma unroll
  for (int i = 0; i < K; i++) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
  const double b = b0, c = c0;
#pragma unroll 1
  for (int n = 0; n < N; n++) {
#pragma unroll
    for (int i = 0; i < K; i++) {
      if (OP == OP_DFMA) a[i] = __fma_rn(a[i], b, c);
      if (OP == OP_DMUL_DADD) a[i] = __dadd_rn(__dmul_rn(a[i], b), c);
      if (OP == OP_DADD) a[i] = __dadd_rn(a[i], c);
      if (OP == OP_DIV) a[i] = __ddiv_rn(c, a[i]) + b;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; i++) s += a[i];
  if (s == 12345.678) out[threadIdx.x] = s;
}
template <int K, int OP> double run(int blocks, int threads, double flops_per_op) {
  double* out;
  cudaMalloc(&out, 1024 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int r = 0; r < 7; r++) {
    cudaEventRecord(e0);
    k<K, OP><<<blocks, threads>>>(out, 0.9999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1);
    if (r >= 2) ms.push_back(t);
  }
  cudaFree(out);
  float best = *std::min_element(ms.begin(), ms.end());
  double ops = (double)blocks * threads * (double)N * K;
  return ops * flops_per_op / (best * 1e-3) / 1e12;  // T(fl)op/s
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // 8 warps per sub-partition (1024 threads per SM), 8 chains per thread; and the sparser shapes for the curve
  double dfma = run<8, OP_DFMA>(sms, 1024, 2.0);
  double dfma_2w = run<8, OP_DFMA>(sms, 256, 2.0);
  double dfma_1c = run<1, OP_DFMA>(sms, 1024, 2.0);
  double dmuladd = run<8, OP_DMUL_DADD>(sms, 1024, 2.0);
  double dadd = run<8, OP_DADD>(sms, 1024, 1.0);
  double ddiv = run<4, OP_DIV>(sms, 1024, 1.0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_attr\": %d,\n"
         " \"dfma_tflops\": %.3f, \"dfma_tflops_2warps_per_smsp\": %.3f, \"dfma_tflops_1chain\": %.3f,\n"
         " \"dmul_dadd_tflops\": %.3f, \"dadd_tops\": %.3f, \"ddiv_tops\": %.4f,\n"
         " \"dfma_per_clk_per_sm\": %.2f,\n"
         " \"how\": \"scripts/micro/fp64_peak.cu: %d SMs x 1024 threads x 8 independent chains x %d dependent ops, best of 5 after 2 warm-up launches, CUDA events; "
         "dfma = 2 flop per DFMA; dmul_dadd = the same a*b+c as separate DMUL and DADD (the product's -fmad=false text), 2 flop per pair; "
         "ddiv = full-precision divisions per second (4 chains)\"}\n",
         p.name, sms, clk, dfma, dfma_2w, dfma_1c, dmuladd, dadd, ddiv,
         dfma * 1e12 / 2.0 / sms / (clk * 1e3), sms, N);
  return 0;
}
