// Scratch: fp64 DFMA dependent latency / throughput on one SM as a function of warps and independent chains per thread
#include <cstdio>
template <int CH>
__global__ void k(long long* cyc, double* out, double x, int n) {
  double a[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) a[c] = x + c;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < CH; ++c) a[c] = a[c] * 1.0000001 + 0.5;
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0; for (int c = 0; c < CH; ++c) s += a[c];
  out[threadIdx.x] = s;
}
template <int CH> void run(long long* dc, double* dout) {
  const int n = 1000;
  for (int threads : {32, 128, 256, 512, 1024}) {
    k<CH><<<1, threads>>>(dc, dout, 1.5, n); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / (n * 4);      // cycles per "one DFMA on each of CH chains"
    printf("chains %d threads %4d: %.2f cyc per step (latency if 1 warp/SMSP) -> %.1f DFMA lanes/cycle/SM\n", CH, threads, per, CH * threads / per);
  }
}
int main() {
  long long* dc; double* dout; cudaMalloc(&dc, 64); cudaMalloc(&dout, 1024 * 8);
  run<1>(dc, dout); run<2>(dc, dout); run<4>(dc, dout); run<8>(dc, dout);
  return 0;
}
