// FP64 latency / issue microbenchmark for sm_100a (B200): dependent DADD/DFMA chains (clock64 per op) and
// throughput versus warps per SM and independent chains per thread.  nvcc -arch=sm_100a -O3 -o fp64_mb fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int q = 0; q < ILP; ++q) x[q] = threadIdx.x + q;
  __syncthreads();
  const long long t0 = clock64();
  for (int n = 0; n < iters; ++n) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = __dadd_rn(x[q], a);
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = __fma_rn(x[q], b, a);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < ILP; ++q) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int sms) {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 2048);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const int threads = 32 * warps_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  chain<ILP><<<sms, threads>>>(out, cyc, 16, 1e-9, 1.0000001);
  cudaEventRecord(e0);
  chain<ILP><<<sms, threads>>>(out, cyc, iters, 1e-9, 1.0000001);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = 2.0 * ILP * iters;  // per thread
  printf("ILP=%d warps/SM=%2d: %6.1f cycles per dependent op (chain step), %7.1f Gop-lanes/s chip, %5.1f%% of 64 lanes/clk/SM @1.965GHz\n", ILP,
         warps_per_sm, double(c) / (2.0 * iters), ops * threads * sms / (ms * 1e-3) / 1e9,
         100.0 * ops * threads * sms / (ms * 1e-3) / (148.0 * 64 * 1.965e9) * (148.0 / sms));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16, 20, 32}) run<1>(w, 148);
  for (int w : {4, 8, 16, 20, 32}) run<2>(w, 148);
  for (int w : {4, 8, 16, 20, 32}) run<4>(w, 148);
  for (int w : {4, 8, 16, 20}) run<8>(w, 148);
  return 0;
}
