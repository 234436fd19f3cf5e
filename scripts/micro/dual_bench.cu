// Stand-alone dual-sweep kernel (2-D, egno 1) at different launch shapes: what can the memory system deliver for its
// 7 read + 5 write streams when the kernel has its own register budget / occupancy?  (scratch, round-2 planning)
#include <cstdio>
#include <vector>
#include "../../pdhg-optimal-control_b200/csrc/pdhg_device.cuh"
using namespace pdhg;

template <int VW> struct V { double e[VW]; };
template <int VW> __device__ __forceinline__ V<VW> ld(const double* p) { V<VW> r; if (VW == 2) { double2 t = *(const double2*)p; r.e[0] = t.x; r.e[VW-1] = t.y; } else r.e[0] = *p; return r; }
template <int VW> __device__ __forceinline__ void st(double* p, const V<VW>& v) { if (VW == 2) *(double2*)p = make_double2(v.e[0], v.e[VW-1]); else *p = v.e[0]; }

#ifndef RESTRICT
#define RESTRICT __restrict__
#endif
template <int VW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_dual(const double* RESTRICT phib, const double* RESTRICT rho_s, const double* RESTRICT alp_s,
                                                   double* RESTRICT rho_d, double* RESTRICT alp_d, const double* RESTRICT cxa,
                                                   const double* RESTRICT cya, int K, int nx, int ny, double sigma, double epsl, double dt,
                                                   double dx, double dy, double* sums) {
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const Recip rc(dt, dx, dy, sigma);
  const int ny2 = ny / VW;
  const size_t items = KN / VW;
  double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t it = (size_t)blockIdx.x * NT + threadIdx.x; it < items; it += (size_t)gridDim.x * NT) {
    const size_t r = it / ny2;
    const int jp = (int)(it - r * ny2), j = jp * VW;
    const int k = (int)(r / nx), i = (int)(r - (size_t)k * nx);
    const size_t row = (size_t)i * ny, g = (size_t)k * n + row + j;
    const double* pb1 = phib + (size_t)(k + 1) * n;
    const int im = (i == 0) ? nx - 1 : i - 1, ip = (i == nx - 1) ? 0 : i + 1;
    const int jm = (j == 0) ? ny - 1 : j - 1, jq = (j + VW == ny) ? 0 : j + VW;
    const V<VW> cc = ld<VW>(pb1 + row + j), pk = ld<VW>(phib + g), ro = ld<VW>(rho_s + g);
    V<VW> ao[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ao[q] = ld<VW>(alp_s + q * KN + g);
    const V<VW> cxm = ld<VW>(pb1 + (size_t)im * ny + j), cxp = ld<VW>(pb1 + (size_t)ip * ny + j);
    const double c_l = pb1[row + jm], c_r = pb1[row + jq];
    const double cx = cxa[i];
    V<VW> rn, an[4];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const double c0 = cc.e[e], cym = (e == 0) ? c_l : cc.e[0], cyp = (e == VW - 1) ? c_r : cc.e[VW - 1];
      const double cy = cya[j + e];
      const double pinv = (ro.e[e] + kRhoOffset) * rc.isig, rinv = prox_rinv(1, pinv);
      const double dyr = (cyp - c0) * rc.idy, dyl = (c0 - cym) * rc.idy, dxr = (cxp.e[e] - c0) * rc.idx, dxl = (c0 - cxm.e[e]) * rc.idx;
      const double a0 = prox_alp(1, ao[0].e[e], dxr, pinv, rinv, cx, true), a1 = prox_alp(1, ao[1].e[e], dxl, pinv, rinv, cx, false);
      const double a2 = prox_alp(1, ao[2].e[e], dyr, pinv, rinv, cy, true), a3 = prox_alp(1, ao[3].e[e], dyl, pinv, rinv, cy, false);
      double vec = (c0 - pk.e[e]) * rc.idt - epsl * ((cxp.e[e] + cxm.e[e] - 2 * c0) * rc.idx2) - epsl * ((cyp + cym - 2 * c0) * rc.idy2);
      vec -= dxr * f_plus(-(cx * a0)) + dxl * f_minus(-(cx * a1)) + dyr * f_plus(-(cy * a2)) + dyl * f_minus(-(cy * a3));
      vec -= lagr(1, a0) + lagr(1, a1) + lagr(1, a2) + lagr(1, a3);
      const double r1 = relu_nan(ro.e[e] + sigma * vec);
      rn.e[e] = r1; an[0].e[e] = a0; an[1].e[e] = a1; an[2].e[e] = a2; an[3].e[e] = a3;
      double d = r1 - ro.e[e]; s[0] += d * d; s[1] += r1 * r1;
      d = a0 - ao[0].e[e]; s[2] += d * d; s[3] += a0 * a0; d = a1 - ao[1].e[e]; s[4] += d * d; s[5] += a1 * a1;
      d = a2 - ao[2].e[e]; s[6] += d * d; s[7] += a2 * a2; d = a3 - ao[3].e[e]; s[8] += d * d; s[9] += a3 * a3;
    }
    st<VW>(rho_d + g, rn);
#pragma unroll
    for (int q = 0; q < 4; ++q) st<VW>(alp_d + q * KN + g, an[q]);
  }
  double t = 0;
#pragma unroll
  for (int q = 0; q < 10; ++q) t += s[q];
  if (t == 12345.678) sums[0] = t;   // keep the sums alive without a reduction
}

template <int VW, int NT, int MINB>
void run(const char* name, int ctas_per_sm, double* d[], int K, int nx, int ny, size_t dsm = 0) {
  cudaFuncSetAttribute(k_dual<VW, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  for (int w = 0; w < 2; ++w) k_dual<VW, NT, MINB><<<grid, NT, dsm>>>(d[0], d[1], d[2], d[3], d[4], d[5], d[6], K, nx, ny, 0.075, 0.0, 1.0 / 64, 2.0 / nx, 2.0 / ny, d[7]);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int w = 0; w < reps; ++w) k_dual<VW, NT, MINB><<<grid, NT, dsm>>>(d[0], d[1], d[2], d[3], d[4], d[5], d[6], K, nx, ny, 0.075, 0.0, 1.0 / 64, 2.0 / nx, 2.0 / ny, d[7]);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_dual<VW, NT, MINB>);
  const double bytes = 8.0 * K * nx * ny * 12;
  printf("%-28s regs %3d  %7.1f us/sweep  %.2f TB/s (12 words/pt)  err=%s\n", name, fa.numRegs, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int K = 64, nx = 256, ny = 256;
  const size_t n = (size_t)nx * ny, KN = K * n, NP = (K + 1) * n;
  std::vector<double> h(NP);
  for (size_t i = 0; i < NP; ++i) h[i] = sin(0.001 * i);
  double* d[8];
  cudaMalloc(&d[0], NP * 8); cudaMemcpy(d[0], h.data(), NP * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&d[1], KN * 8); cudaMalloc(&d[2], 4 * KN * 8); cudaMalloc(&d[3], KN * 8); cudaMalloc(&d[4], 4 * KN * 8);
  cudaMemset(d[2], 0, 4 * KN * 8);
  for (size_t i = 0; i < KN; ++i) h[i] = 70.0 + sin(0.01 * i);
  cudaMemcpy(d[1], h.data(), KN * 8, cudaMemcpyHostToDevice);
  std::vector<double> c(256); for (int i = 0; i < 256; ++i) c[i] = (i * 2.0 / 256 - 1) * (i * 2.0 / 256 - 1) + 0.1;
  cudaMalloc(&d[5], 2048); cudaMalloc(&d[6], 2048); cudaMalloc(&d[7], 64);
  cudaMemcpy(d[5], c.data(), 2048, cudaMemcpyHostToDevice); cudaMemcpy(d[6], c.data(), 2048, cudaMemcpyHostToDevice);
  run<2, 512, 1>("VW2 512thr x1 (as coop)", 1, d, K, nx, ny);
  run<2, 512, 1>("VW2 512thr x1 +80KB smem", 1, d, K, nx, ny, 80 * 1024);
  run<2, 512, 1>("VW2 512thr x1 +160KB smem", 1, d, K, nx, ny, 160 * 1024);
  run<2, 256, 2>("VW2 256thr x2", 2, d, K, nx, ny);
  run<2, 256, 3>("VW2 256thr x3", 3, d, K, nx, ny);
  run<2, 256, 4>("VW2 256thr x4", 4, d, K, nx, ny);
  run<1, 256, 4>("VW1 256thr x4", 4, d, K, nx, ny);
  run<1, 256, 6>("VW1 256thr x6", 6, d, K, nx, ny);
  run<1, 256, 8>("VW1 256thr x8", 8, d, K, nx, ny);
  run<2, 256, 2>("VW2 256thr x2 grid x8", 16, d, K, nx, ny);
  run<1, 256, 4>("VW1 256thr x4 grid x4", 16, d, K, nx, ny);
  return 0;
}
