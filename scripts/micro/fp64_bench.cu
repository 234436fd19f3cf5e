// fp64 pipe throughput / latency probe (scratch)
#include <cstdio>
__global__ void k(long long* cyc, double* out, double x, int n) {
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    a0 = a0 * 1.0000001 + 0.5; a1 = a1 * 1.0000001 + 0.5; a2 = a2 * 1.0000001 + 0.5; a3 = a3 * 1.0000001 + 0.5;
    a4 = a4 * 1.0000001 + 0.5; a5 = a5 * 1.0000001 + 0.5; a6 = a6 * 1.0000001 + 0.5; a7 = a7 * 1.0000001 + 0.5;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void kdiv(long long* cyc, double* out, double x, int n) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = 1.0 / (1.0 + a);
  long long t1 = clock64();
  double b = x;
#pragma unroll 1
  for (int i = 0; i < n; ++i) b = sqrt(b + 1.0);
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  out[threadIdx.x] = a + b;
}
int main() {
  long long* dc; double* dout; cudaMalloc(&dc, 1024 * 8); cudaMalloc(&dout, 148 * 1024 * 8);
  int n = 2000;
  for (int threads : {32, 64, 128, 256, 512, 1024}) {
    k<<<1, threads>>>(dc, dout, 1.5, n); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("threads %4d: %.2f cycles per 8 independent DFMA per thread -> %.1f DFMA lanes/cycle/SM\n", threads, (double)c / n, 8.0 * threads * n / c);
  }
  kdiv<<<1, 32>>>(dc, dout, 1.5, n); cudaDeviceSynchronize();
  long long c[2]; cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
  printf("dependent fp64 division+add: %.0f cycles; sqrt+add: %.0f cycles\n", (double)c[0] / n, (double)c[1] / n);
  return 0;
}
