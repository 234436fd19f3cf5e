// Micro-benchmark of the shared-memory FFT primitive (scratch; not part of the library).
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../pdhg-optimal-control_b200/csrc/pdhg_device.cuh"
using namespace pdhg;

__global__ void k_fft(FftPlan plan, const double2* tw_g, int rows, int ld, int reps, long long* cyc, double2* out) {
  extern __shared__ __align__(16) double2 sm[];
  double2* tw = sm;
  double2* b0 = tw + plan.n;
  double2* b1 = b0 + (size_t)rows * ld;
  for (int i = threadIdx.x; i < plan.n; i += blockDim.x) tw[i] = tw_g[i];
  for (int i = threadIdx.x; i < rows * ld; i += blockDim.x) b0[i] = make_double2(sin(0.1 * i), cos(0.3 * i));
  __syncthreads();
  long long t0 = clock64();
  double2* r = b0;
  for (int it = 0; it < reps; ++it) {
    r = fft_rows(b0, b1, plan, ld, tw, rows, 1.0);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r[threadIdx.x];
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 256, rows = argc > 2 ? atoi(argv[2]) : 8, threads = argc > 3 ? atoi(argv[3]) : 512;
  FftPlan plan; plan.n = n; plan.nstages = 0; int m = n;
  int maxr = argc > 4 ? atoi(argv[4]) : 16;
  while (maxr >= 16 && m % 16 == 0) { plan.radix[plan.nstages++] = 16; m /= 16; }
  while (maxr >= 8 && m % 8 == 0) { plan.radix[plan.nstages++] = 8; m /= 8; }
  while (m % 4 == 0) { plan.radix[plan.nstages++] = 4; m /= 4; }
  while (m % 2 == 0) { plan.radix[plan.nstages++] = 2; m /= 2; }
  while (m % 5 == 0) { plan.radix[plan.nstages++] = 5; m /= 5; }
  while (m % 3 == 0) { plan.radix[plan.nstages++] = 3; m /= 3; }
  std::vector<double2> tw(n);
  for (int i = 0; i < n; ++i) tw[i] = make_double2(cos(-2 * M_PI * i / n), sin(-2 * M_PI * i / n));
  double2 *dtw, *dout; long long* dc;
  cudaMalloc(&dtw, n * 16); cudaMemcpy(dtw, tw.data(), n * 16, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 148 * 1024 * 16); cudaMalloc(&dc, 148 * 8);
  int ld = fft_ld(n), reps = 200;
  size_t smem = (size_t)(n + 2 * rows * ld) * 16;
  cudaFuncSetAttribute(k_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int grid : {1, 148}) {
    k_fft<<<grid, threads, smem>>>(plan, dtw, rows, ld, reps, dc, dout);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[148]; cudaMemcpy(c, dc, grid * 8, cudaMemcpyDeviceToHost);
    printf("n=%d rows=%d threads=%d grid=%d: %s  cycles/transform-batch %.0f  per stage %.0f (stages %d)\n", n, rows, threads, grid,
           cudaGetErrorString(e), (double)c[0] / reps, (double)c[0] / reps / plan.nstages, plan.nstages);
  }
  return 0;
}
