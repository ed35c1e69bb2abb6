// Dependent-chain latency and single-warp issue rate of DFMA / shuffle / sqrt on one SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void chain(double *out, long long *cyc, double a, double b, int n) {
  double x[CH];
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CH; ++c) s += x[c];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void chain_sqrt(double *out, long long *cyc, double a, int n) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = sqrt(x) + a;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_div(double *out, long long *cyc, double a, int n) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = a / x + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_rcp(double *out, long long *cyc, double a, int n) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = 1.0 / x + a;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_drcp(double *out, long long *cyc, double a, int n) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = __drcp_rn(x) + a;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_rsqrt(double *out, long long *cyc, double a, int n) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = rsqrt(x) + a;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_shfl(double *out, long long *cyc, int n) {
  double x = threadIdx.x + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc, h[8];
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  const int n = 4096;
#define RUN(CH, T) chain<CH><<<1, T>>>(out, cyc, 1.0000001, 1e-9, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DFMA chains=%d threads=%d: %.2f cycles per round, %.2f per DFMA\n", CH, T, (double)h[0] / n, (double)h[0] / n / CH);
  RUN(1, 32) RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(16, 32) RUN(1, 1) RUN(8, 128) RUN(16, 128) RUN(8, 256)
  chain_sqrt<<<1, 32>>>(out, cyc, 1.0, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("sqrt+add chain: %.1f cycles\n", (double)h[0] / n);
  chain_div<<<1, 32>>>(out, cyc, 3.0, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("div+add chain: %.1f cycles\n", (double)h[0] / n);
  chain_rcp<<<1, 32>>>(out, cyc, 1.5, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("1.0/x+add chain: %.1f cycles\n", (double)h[0] / n);
  chain_drcp<<<1, 32>>>(out, cyc, 1.5, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("__drcp_rn+add chain: %.1f cycles\n", (double)h[0] / n);
  chain_rsqrt<<<1, 32>>>(out, cyc, 1.5, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("rsqrt+add chain: %.1f cycles\n", (double)h[0] / n);
  chain_shfl<<<1, 32>>>(out, cyc, n); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("shfl64+add chain: %.1f cycles\n", (double)h[0] / n);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
