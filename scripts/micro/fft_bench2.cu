// Floor test: fully compile-time radix-4 Stockham stages for n=256 (scratch).
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../pdhg-optimal-control_b200/csrc/pdhg_device.cuh"
using namespace pdhg;

template <int N, int NS, int ROWS, int LD>
__device__ __forceinline__ void st4(const double2* __restrict__ src, double2* __restrict__ dst, const double2* __restrict__ tw) {
  constexpr int nb = N / 4, twstep = nb / NS;
  for (int w = threadIdx.x; w < ROWS * nb; w += blockDim.x) {
    const int row = w / nb, j = w % nb, k = j % NS;
    const double2* s = src + row * LD + j;
    double2* d = dst + row * LD + (j - k) * 4 + k;
    double2 a0 = s[0], a1 = s[nb], a2 = s[2 * nb], a3 = s[3 * nb];
    if (NS > 1) {
      double2 w1 = tw[k * twstep], w2 = tw[2 * k * twstep], w3 = tw[3 * k * twstep];
      a1 = cmul(a1, w1); a2 = cmul(a2, w2); a3 = cmul(a3, w3);
    }
    double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot_mi(csub(a1, a3), 1.0);
    d[0] = cadd(t0, t2); d[NS] = cadd(t1, t3); d[2 * NS] = csub(t0, t2); d[3 * NS] = csub(t1, t3);
  }
}

template <int ROWS>
__global__ void k_fft(const double2* tw_g, int reps, long long* cyc, double2* out) {
  constexpr int N = 256, LD = 257;
  extern __shared__ __align__(16) double2 sm[];
  double2* tw = sm; double2* b0 = tw + N; double2* b1 = b0 + ROWS * LD;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = tw_g[i];
  for (int i = threadIdx.x; i < ROWS * LD; i += blockDim.x) b0[i] = make_double2(sin(0.1 * i), cos(0.3 * i));
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    st4<N, 1, ROWS, LD>(b0, b1, tw); __syncthreads();
    st4<N, 4, ROWS, LD>(b1, b0, tw); __syncthreads();
    st4<N, 16, ROWS, LD>(b0, b1, tw); __syncthreads();
    st4<N, 64, ROWS, LD>(b1, b0, tw); __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = b0[threadIdx.x];
}

// latency probes: dependent chains of DFMA / LDS
__global__ void k_lat(long long* cyc, double* out, double x) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = (double)((threadIdx.x + 1) & 63);
  __syncthreads();
  double a = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1000; ++i) a = a * 1.0000001 + 0.5;
  long long t1 = clock64();
  int idx = threadIdx.x & 63;
#pragma unroll 1
  for (int i = 0; i < 1000; ++i) idx = (int)sm[idx];
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  out[threadIdx.x] = a + idx;
}

int main() {
  const int n = 256, reps = 200;
  std::vector<double2> tw(n);
  for (int i = 0; i < n; ++i) tw[i] = make_double2(cos(-2 * M_PI * i / n), sin(-2 * M_PI * i / n));
  double2 *dtw, *dout; long long* dc;
  cudaMalloc(&dtw, n * 16); cudaMemcpy(dtw, tw.data(), n * 16, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 148 * 1024 * 16); cudaMalloc(&dc, 148 * 8);
  long long c[2];
  size_t smem = (size_t)(n + 2 * 8 * 257) * 16;
  cudaFuncSetAttribute(k_fft<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_fft<8><<<1, 512, smem>>>(dtw, reps, dc, dout); cudaDeviceSynchronize(); cudaMemcpy(c, dc, 8, cudaMemcpyDeviceToHost);
  printf("compile-time n=256 rows=8 threads=512: per stage %.0f cycles\n", (double)c[0] / reps / 4);
  k_fft<1><<<1, 64, smem>>>(dtw, reps, dc, dout); cudaDeviceSynchronize(); cudaMemcpy(c, dc, 8, cudaMemcpyDeviceToHost);
  printf("compile-time n=256 rows=1 threads=64: per stage %.0f cycles\n", (double)c[0] / reps / 4);
  k_fft<1><<<1, 256, smem>>>(dtw, reps, dc, dout); cudaDeviceSynchronize(); cudaMemcpy(c, dc, 8, cudaMemcpyDeviceToHost);
  printf("compile-time n=256 rows=1 threads=256: per stage %.0f cycles\n", (double)c[0] / reps / 4);
  k_lat<<<1, 32>>>(dc, (double*)dout, 1.5); cudaDeviceSynchronize(); cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
  printf("latency: DFMA dependent %.1f cycles, LDS.64+cvt dependent %.1f cycles\n", c[0] / 1000.0, c[1] / 1000.0);
  return 0;
}
