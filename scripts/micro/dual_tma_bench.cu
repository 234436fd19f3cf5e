// Scratch micro-benchmark (round 2): the dual sweep (2-D, egno 1, 256 x 256 x K = 64) as stand-alone kernels —
//   direct global loads (the cooperative kernel's phase D shape)  vs.  TMA bulk-copy row pipelines in several shapes,
// plus "loads only" / "loads + stores" skeletons that separate the data movement from the arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dual_tma_bench dual_tma_bench.cu
#include <cstdio>
#include <vector>
#include "../../pdhg-optimal-control_b200/csrc/pdhg_device.cuh"
using namespace pdhg;

struct P { const double *phib, *rho_s, *alp_s; double *rho_d, *alp_d; const double *cxa, *cya; int K, nx, ny; double sigma, epsl, dt, dx, dy; double* sums; };

__device__ __forceinline__ void point(const Recip& rc, double sigma, double epsl, double c0, double cxm, double cxp, double cym, double cyp, double pk, double ro,
                                      const double (&ao)[4], double cx, double cy, double& rn, double (&an)[4], double (&s)[10]) {
  const double pinv = (ro + kRhoOffset) * rc.isig, rinv = prox_rinv(1, pinv);
  const double dyr = (cyp - c0) * rc.idy, dyl = (c0 - cym) * rc.idy, dxr = (cxp - c0) * rc.idx, dxl = (c0 - cxm) * rc.idx;
  an[0] = prox_alp(1, ao[0], dxr, pinv, rinv, cx, true); an[1] = prox_alp(1, ao[1], dxl, pinv, rinv, cx, false);
  an[2] = prox_alp(1, ao[2], dyr, pinv, rinv, cy, true); an[3] = prox_alp(1, ao[3], dyl, pinv, rinv, cy, false);
  double vec = (c0 - pk) * rc.idt - epsl * ((cxp + cxm - 2 * c0) * rc.idx2) - epsl * ((cyp + cym - 2 * c0) * rc.idy2);
  vec -= dxr * f_plus(-(cx * an[0])) + dxl * f_minus(-(cx * an[1])) + dyr * f_plus(-(cy * an[2])) + dyl * f_minus(-(cy * an[3]));
  vec -= lagr(1, an[0]) + lagr(1, an[1]) + lagr(1, an[2]) + lagr(1, an[3]);
  rn = relu_nan(ro + sigma * vec);
  double d = rn - ro; s[0] += d * d; s[1] += rn * rn;
#pragma unroll
  for (int q = 0; q < 4; ++q) { d = an[q] - ao[q]; s[2 + 2 * q] += d * d; s[3 + 2 * q] += an[q] * an[q]; }
}

// ---- direct loads, one double2 item per thread and iteration (= phase D of the cooperative kernel) ----
template <int NT, int MINB, bool CG>
__global__ void __launch_bounds__(NT, MINB) k_direct(P p) {
  const int K = p.K, nx = p.nx, ny = p.ny, ny2 = ny / 2;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const Recip rc(p.dt, p.dx, p.dy, p.sigma);
  double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int items = K * nx * ny2;
  auto ld2 = [](const double* q) { return CG ? __ldcg(reinterpret_cast<const double2*>(q)) : *reinterpret_cast<const double2*>(q); };
  auto ld1 = [](const double* q) { return CG ? __ldcg(q) : *q; };
  for (int it = blockIdx.x * NT + threadIdx.x; it < items; it += gridDim.x * NT) {
    const int r = it / ny2, jp = it - r * ny2, j = 2 * jp, k = r / nx, i = r - k * nx;
    const size_t row = (size_t)i * ny, g = (size_t)k * n + row + j;
    const double* pb1 = p.phib + (size_t)(k + 1) * n;
    const int im = (i == 0) ? nx - 1 : i - 1, ip = (i == nx - 1) ? 0 : i + 1, jm = (j == 0) ? ny - 1 : j - 1, jq = (j + 2 == ny) ? 0 : j + 2;
    const double2 cc = ld2(pb1 + row + j), pk = ld2(p.phib + g), ro = ld2(p.rho_s + g);
    double2 ao[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) ao[q] = ld2(p.alp_s + q * KN + g);
    const double2 cxm = ld2(pb1 + (size_t)im * ny + j), cxp = ld2(pb1 + (size_t)ip * ny + j);
    const double c_l = ld1(pb1 + row + jm), c_r = ld1(pb1 + row + jq);
    const double cx = p.cxa[i];
    double rn[2], an[2][4];
    { const double a0[4] = {ao[0].x, ao[1].x, ao[2].x, ao[3].x}; point(rc, p.sigma, p.epsl, cc.x, cxm.x, cxp.x, c_l, cc.y, pk.x, ro.x, a0, cx, p.cya[j], rn[0], an[0], s); }
    { const double a1[4] = {ao[0].y, ao[1].y, ao[2].y, ao[3].y}; point(rc, p.sigma, p.epsl, cc.y, cxm.y, cxp.y, cc.x, c_r, pk.y, ro.y, a1, cx, p.cya[j + 1], rn[1], an[1], s); }
    *reinterpret_cast<double2*>(p.rho_d + g) = make_double2(rn[0], rn[1]);
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<double2*>(p.alp_d + q * KN + g) = make_double2(an[0][q], an[1][q]);
  }
  double t = 0;
#pragma unroll
  for (int q = 0; q < 10; ++q) t += s[q];
  if (t == 12345.678) p.sums[0] = t;
}

// ---- TMA row pipeline.  Tile = R rows x ny of one time row.  MODE 0: full dual sweep; 1: loads only (consumers touch one word per
// row); 2: loads + stores (copy through), no arithmetic.  PROD = 1: a dedicated producer warp (NT = 32 * (consumer warps + 1)).
// SPLIT = 1: one bulk copy per row (2 KB) instead of one per array (R rows).
template <int NT, int R, int S, int MODE, int PROD, int SPLIT>
__global__ void __launch_bounds__(NT, 1) k_tma(P p) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 8;
  double* ring = reinterpret_cast<double*>(sm + 128);
  const int K = p.K, nx = p.nx, ny = p.ny, ny2 = ny / 2;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const Recip rc(p.dt, p.dx, p.dy, p.sigma);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NCW = NT / 32 - PROD;                 // consumer warps
  const int ctid = tid - 32 * PROD, NCT = 32 * NCW;   // consumer thread id
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
    mbar_fence_init();
  }
  __syncthreads();
  const int stage_d = (7 * R + 2) * ny;
  const int ntx = nx / R, T = K * ntx, G = gridDim.x;
  const uint32_t rowb = ny * 8u;
  double s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

  auto produce = [&](int tile, int st_i, uint32_t par) {
    const int k = tile / ntx, i0 = (tile - k * ntx) * R;
    mbar_wait(&empty[st_i], par ^ 1u);
    double* st = ring + (size_t)st_i * stage_d;
    if (lane == 0) mbar_arrive_expect_tx(&full[st_i], (uint32_t)(7 * R + 2) * rowb);
    __syncwarp();
    const double* pb1 = p.phib + (size_t)(k + 1) * n;
    const int im = (i0 == 0) ? nx - 1 : i0 - 1, ip = (i0 + R == nx) ? 0 : i0 + R;
    const size_t g0 = (size_t)k * n + (size_t)i0 * ny;
    if (!SPLIT) {
      if (lane == 0) bulk_g2s(st, pb1 + (size_t)im * ny, rowb, &full[st_i]);
      else if (lane == 1) bulk_g2s(st + ny, pb1 + (size_t)i0 * ny, rowb * R, &full[st_i]);
      else if (lane == 2) bulk_g2s(st + (size_t)(R + 1) * ny, pb1 + (size_t)ip * ny, rowb, &full[st_i]);
      else if (lane == 3) bulk_g2s(st + (size_t)(R + 2) * ny, p.phib + g0, rowb * R, &full[st_i]);
      else if (lane == 4) bulk_g2s(st + (size_t)(2 * R + 2) * ny, p.rho_s + g0, rowb * R, &full[st_i]);
      else if (lane < 9) bulk_g2s(st + (size_t)(3 * R + 2 + (lane - 5) * R) * ny, p.alp_s + (size_t)(lane - 5) * KN + g0, rowb * R, &full[st_i]);
    } else {
      for (int c = lane; c < 7 * R + 2; c += 32) {     // one copy per row
        const double* src;
        if (c == 0) src = pb1 + (size_t)im * ny;
        else if (c <= R) src = pb1 + (size_t)(i0 + c - 1) * ny;
        else if (c == R + 1) src = pb1 + (size_t)ip * ny;
        else { const int a = (c - (R + 2)) / R, r = (c - (R + 2)) - a * R; src = (a == 0 ? p.phib : (a == 1 ? p.rho_s : p.alp_s + (size_t)(a - 2) * KN)) + g0 + (size_t)r * ny; }
        bulk_g2s(st + (size_t)c * ny, src, rowb, &full[st_i]);
      }
    }
  };

  if (PROD) {
    if (warp == 0) {
      int ps = 0; uint32_t pp = 0;
      for (int tile = blockIdx.x; tile < T; tile += G) { produce(tile, ps, pp); if (++ps == S) { ps = 0; pp ^= 1u; } }
      return;
    }
  }
  int ps = 0; uint32_t pp = 0;
  if (!PROD && warp == 0) {
    for (int m = 0; m < S - 1; ++m) { const int tm = blockIdx.x + m * G; if (tm < T) produce(tm, ps, pp); if (++ps == S) { ps = 0; pp ^= 1u; } }
  }
  int cs = 0; uint32_t cp = 0;
  for (int tile = blockIdx.x; tile < T; tile += G) {
    if (!PROD && warp == 0) { const int tm = tile + (S - 1) * G; if (tm < T) produce(tm, ps, pp); if (++ps == S) { ps = 0; pp ^= 1u; } }
    const int k = tile / ntx, i0 = (tile - k * ntx) * R;
    mbar_wait(&full[cs], cp);
    const double* st = ring + (size_t)cs * stage_d;
    if (MODE == 1) {
      for (int c = ctid; c < 7 * R + 2; c += NCT) s[0] += st[(size_t)c * ny + 5];
    } else {
      for (int it = ctid; it < R * ny2; it += NCT) {
        const int r = it / ny2, jp = it - r * ny2, j = 2 * jp, i = i0 + r;
        const double* prow = st + (size_t)(r + 1) * ny;
        const int jm = (j == 0) ? ny - 1 : j - 1, jq = (j + 2 == ny) ? 0 : j + 2;
        const double2 cc = *reinterpret_cast<const double2*>(prow + j), cxm = *reinterpret_cast<const double2*>(prow - ny + j),
                      cxp = *reinterpret_cast<const double2*>(prow + ny + j);
        const double c_l = prow[jm], c_r = prow[jq];
        const double* own = st + (size_t)(R + 2 + r) * ny + j;
        const double2 pk = *reinterpret_cast<const double2*>(own), ro = *reinterpret_cast<const double2*>(own + (size_t)R * ny);
        double2 ao[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ao[q] = *reinterpret_cast<const double2*>(own + (size_t)(2 + q) * R * ny);
        const size_t g = (size_t)k * n + (size_t)i * ny + j;
        if (MODE == 2) {
          *reinterpret_cast<double2*>(p.rho_d + g) = make_double2(ro.x + cc.x + cxm.x + pk.x + c_l, ro.y + cxp.y + c_r);
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<double2*>(p.alp_d + q * KN + g) = ao[q];
        } else {
          const double cx = p.cxa[i];
          double rn[2], an[2][4];
          { const double a0[4] = {ao[0].x, ao[1].x, ao[2].x, ao[3].x}; point(rc, p.sigma, p.epsl, cc.x, cxm.x, cxp.x, c_l, cc.y, pk.x, ro.x, a0, cx, p.cya[j], rn[0], an[0], s); }
          { const double a1[4] = {ao[0].y, ao[1].y, ao[2].y, ao[3].y}; point(rc, p.sigma, p.epsl, cc.y, cxm.y, cxp.y, cc.x, c_r, pk.y, ro.y, a1, cx, p.cya[j + 1], rn[1], an[1], s); }
          *reinterpret_cast<double2*>(p.rho_d + g) = make_double2(rn[0], rn[1]);
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<double2*>(p.alp_d + q * KN + g) = make_double2(an[0][q], an[1][q]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[cs]);
    if (++cs == S) { cs = 0; cp ^= 1u; }
  }
  double t = 0;
#pragma unroll
  for (int q = 0; q < 10; ++q) t += s[q];
  if (t == 12345.678) p.sums[0] = t;
}

static double checksum(const double* d, size_t cnt) {
  std::vector<double> h(cnt);
  cudaMemcpy(h.data(), d, cnt * 8, cudaMemcpyDeviceToHost);
  double s = 0; for (size_t i = 0; i < cnt; i += 97) s += h[i] * (1 + (i % 13));
  return s;
}

template <typename F> static void run(const char* name, F launch, const P& p, double bytes, bool check) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t KN = (size_t)p.K * p.nx * p.ny;
  cudaMemset(p.rho_d, 0, KN * 8); cudaMemset(p.alp_d, 0, 4 * KN * 8);
  launch(); launch();
  cudaError_t err = cudaDeviceSynchronize();
  const int reps = 10;
  cudaEventRecord(e0);
  for (int w = 0; w < reps; ++w) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double c1 = 0, c2 = 0;
  if (check) { c1 = checksum(p.rho_d, KN); c2 = checksum(p.alp_d, 4 * KN); }
  printf("%-44s %7.1f us  %.2f TB/s  chk %.10e %.10e  %s\n", name, ms * 1e3 / reps, bytes / (ms * 1e-3 / reps) / 1e12, c1, c2,
         cudaGetErrorString(err != cudaSuccess ? err : cudaGetLastError()));
}

// all arrays carved from ONE allocation like the cooperative kernel's workspace: phi_bar [NP], rho_s [KN], alp_s [4 KN], rho_d [KN],
// alp_d [4 KN], every array start shifted by `skew` bytes more than the previous one (skew = 0: the dual arrays sit at exact
// multiples of 2^25 bytes from each other)
static void slab_test(int sms, size_t skew, const std::vector<double>& hphi, const std::vector<double>& hrho, const std::vector<double>& halp,
                      const double* cx, const double* cy, double* sums) {
  const int K = 64, nx = 256, ny = 256;
  const size_t n = (size_t)nx * ny, KN = K * n, NP = (K + 1) * n;
  char* base;
  const size_t total = (NP + 10 * KN) * 8 + 16 * skew + (1 << 20);
  cudaMalloc(&base, total);
  size_t off = 0;
  auto take = [&](size_t cnt, int idx) { double* q = (double*)(base + off + idx * skew); off += cnt * 8; return q; };
  double* phib = take(NP, 0);
  double* rho_s = take(KN, 1);
  double* alp_s = take(4 * KN, 2);
  double* rho_d = take(KN, 3);
  double* alp_d = take(4 * KN, 4);
  cudaMemcpy(phib, hphi.data(), NP * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(rho_s, hrho.data(), KN * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(alp_s, halp.data(), 4 * KN * 8, cudaMemcpyHostToDevice);
  P p{phib, rho_s, alp_s, rho_d, alp_d, cx, cy, K, nx, ny, 0.075, 0.0, 1.0 / 64, 2.0 / nx, 2.0 / ny, sums};
  char nm[96]; snprintf(nm, 96, "direct 512thr x1, one slab, skew %zu B", skew);
  run(nm, [&] { k_direct<512, 1, false><<<sms, 512>>>(p); }, p, 8.0 * KN * 12, true);
  cudaFree(base);
}

int main() {
  const int K = 64, nx = 256, ny = 256;
  const size_t n = (size_t)nx * ny, KN = K * n, NP = (K + 1) * n;
  std::vector<double> h(NP);
  for (size_t i = 0; i < NP; ++i) h[i] = sin(0.001 * i);
  double* d[8];
  cudaMalloc(&d[0], NP * 8); cudaMemcpy(d[0], h.data(), NP * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&d[1], KN * 8); cudaMalloc(&d[2], 4 * KN * 8); cudaMalloc(&d[3], KN * 8); cudaMalloc(&d[4], 4 * KN * 8);
  std::vector<double> ha(4 * KN);
  for (size_t i = 0; i < 4 * KN; ++i) ha[i] = 0.3 * sin(0.37 * i);
  cudaMemcpy(d[2], ha.data(), 4 * KN * 8, cudaMemcpyHostToDevice);
  for (size_t i = 0; i < KN; ++i) h[i] = 70.0 + sin(0.01 * i);
  cudaMemcpy(d[1], h.data(), KN * 8, cudaMemcpyHostToDevice);
  std::vector<double> c(256); for (int i = 0; i < 256; ++i) c[i] = (i * 2.0 / 256 - 1) * (i * 2.0 / 256 - 1) + 0.1;
  cudaMalloc(&d[5], 2048); cudaMalloc(&d[6], 2048); cudaMalloc(&d[7], 64);
  cudaMemcpy(d[5], c.data(), 2048, cudaMemcpyHostToDevice); cudaMemcpy(d[6], c.data(), 2048, cudaMemcpyHostToDevice);
  P p{d[0], d[1], d[2], d[3], d[4], d[5], d[6], K, nx, ny, 0.075, 0.0, 1.0 / 64, 2.0 / nx, 2.0 / ny, d[7]};
  const double b12 = 8.0 * KN * 12, b7 = 8.0 * KN * 7;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run("direct 512thr x1", [&] { k_direct<512, 1, false><<<sms, 512>>>(p); }, p, b12, true);
  {
    std::vector<double> hphi(NP), hrho(KN);
    for (size_t i = 0; i < NP; ++i) hphi[i] = sin(0.001 * i);
    for (size_t i = 0; i < KN; ++i) hrho[i] = 70.0 + sin(0.01 * i);
    for (size_t skew : {(size_t)0, (size_t)256, (size_t)4096 + 256, (size_t)65536 + 2304, (size_t)(1 << 20) + 8448, (size_t)(3 << 20) + 512})
      slab_test(sms, skew, hphi, hrho, ha, d[5], d[6], d[7]);
  }
  run("direct 512thr x1 ld.cg", [&] { k_direct<512, 1, true><<<sms, 512>>>(p); }, p, b12, true);
  run("direct 256thr x2", [&] { k_direct<256, 2, false><<<2 * sms, 256>>>(p); }, p, b12, true);
  run("direct 256thr x3 (80 regs)", [&] { k_direct<256, 3, false><<<3 * sms, 256>>>(p); }, p, b12, true);
#define TMA(NT, R, S, MODE, PROD, SPLIT, bytes, chk) { \
    const size_t smem = 128 + (size_t)S * (7 * R + 2) * ny * 8; \
    cudaFuncSetAttribute(k_tma<NT, R, S, MODE, PROD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    char nm[96]; snprintf(nm, 96, "tma NT%d R%d S%d mode%d prod%d split%d (%zu KB)", NT, R, S, MODE, PROD, SPLIT, smem / 1024); \
    run(nm, [&] { k_tma<NT, R, S, MODE, PROD, SPLIT><<<sms, NT, smem>>>(p); }, p, bytes, chk); }
  TMA(512, 4, 3, 0, 0, 0, b12, true)
  return 0;
}
