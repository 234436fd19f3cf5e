// Cooperative multi-CTA PDHG solver: 2-D grids of any size / any time_step_per_PDHG, and 1-D space-time blocks
// that do not fit one SM's shared memory.  ONE persistent cooperative launch runs the whole time-block march
// (PDHG_multi_step, jaxsrc/utils/utils_pdhg_solver.py:97-225) with every decision taken on the device.
//
// State lives in global memory (L2-resident for small grids, streamed from HBM for large ones), planar, t slowest,
// then x, then y (contiguous).  A 1-D problem is run as a 2-D one with nx = 1 (its x axis is the contiguous "y").
// One outer iteration = four grid-synchronised phases:
//   A  continuity residual (update_fns_in_pdhg.py:72-96) + real-to-complex FFT along y in smem (two rows per
//      complex transform) -> half spectrum stored transposed  zt[k][ky][i]
//   B  per ky: FFT along x in smem, per-mode t-solve (Thomas, utils_precond.py:10-35,164-169; tables of the
//      modified coefficients are precomputed once), inverse FFT along x, in place
//   C  inverse FFT along y, phi_next = phi_prev + tau*u, phi_bar = 2 phi_next - phi_prev, err1 partial sums
//   D  dual sweep(s) (update_fns_in_pdhg.py:150-180): alp prox, HJ residual, rho prox, inner/outer error sums
// followed by one grid reduction that drives the inner early exit, the convergence / NaN exits, the records and
// the step-size fallback exactly as utils_pdhg_solver.py:59-80,174-187 do on the host.
#include <cooperative_groups.h>
#include <string.h>

#include "pdhg_params.h"

namespace cg = cooperative_groups;

namespace pdhg {

constexpr int kNQ = 20;        // reduced quantities per epoch
constexpr int kThreads = 256;

struct CoopWs {
  double* phi[2];     // ping-pong phi [(K+1) n]
  double* phib;       // phi_bar [(K+1) n]
  double* rho[2];     // ping-pong rho [K n]
  double* alp[2];     // ping-pong alp [A][K n]
  double2* zt;        // [K][nyh][nx] half spectrum, transposed
  double* partials;   // [2][grid][kNQ]
  double* den;        // [K][nyh][nx]  Thomas pivots        (K > 1)
  double* tu;         // [K][nyh][nx]  modified super-diag  (K > 1)
};

enum : int { MODE_MARCH = 0, MODE_PRIMAL = 1, MODE_DUAL = 2, MODE_TABLES = 3 };

struct CoopArgs {
  MarchParams p;
  CoopWs w;
  int b;              // instance index
  int mode;
  int A;              // 2*ndim control arrays
  int nxe, nye;       // effective 2-D extents (1-D: nxe = 1, nye = nx)
  int nyh;            // nye/2 + 1
  int TR, TKY;        // rows per y-FFT tile (even), ky rows per x-FFT tile
  int has_x;          // 0 for a 1-D problem
  double dxe, dye;
  const double* coef_xe;
  const double* coef_ye;
  const double2* tw_xe;
  const double2* tw_ye;
  FftPlan plan_xe, plan_ye;
  // operator modes
  const double* op_phi_in;   // PRIMAL: phi_prev ; DUAL: phi_bar   [(K+1) n]
  double* op_phi_out;        // PRIMAL: phi_next
  double op_step;            // tau or sigma
  double op_eps;
  int* op_ninner;
  double* op_err;
};

struct Ctx {
  const CoopArgs& a;
  cg::grid_group grid;
  double2* sm;        // dynamic shared memory
  double* red;        // [kNQ*8] static scratch
  int epoch;
  __device__ Ctx(const CoopArgs& a_, double2* sm_, double* red_) : a(a_), grid(cg::this_grid()), sm(sm_), red(red_), epoch(0) {}
};

// ---- neighbour index helpers (bc 0 periodic; bc 1 Neumann, utils_diff_op.py:19-22,61-64,219-224) ----
struct Nbr { int m, p; double wm, wp; };   // wm/wp: 0 when the one-sided difference is forced to zero (Neumann edge)

__device__ __forceinline__ Nbr nbr(int i, int n, int bc) {
  Nbr r;
  if (bc == 0) { r.m = (i == 0) ? n - 1 : i - 1; r.p = (i == n - 1) ? 0 : i + 1; r.wm = 1.0; r.wp = 1.0; }
  else { r.m = (i == 0) ? 0 : i - 1; r.p = (i == n - 1) ? n - 1 : i + 1; r.wm = (i == 0) ? 0.0 : 1.0; r.wp = (i == n - 1) ? 0.0 : 1.0; }
  return r;
}

// grid-wide sum of kNQ per-thread values; result broadcast to every thread of every CTA (deterministic order).
__device__ void grid_sum(Ctx& c, double (&v)[kNQ]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int G = gridDim.x;
#pragma unroll
  for (int q = 0; q < kNQ; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < kNQ; ++q) c.red[q * 8 + warp] = v[q];
  }
  __syncthreads();
  double* part = c.a.w.partials + (size_t)(c.epoch & 1) * G * kNQ;
  if (tid < kNQ) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += c.red[tid * 8 + w];
    part[(size_t)blockIdx.x * kNQ + tid] = t;
  }
  c.grid.sync();
  // every CTA reduces all partials: warp w handles quantities w, w+nw, ...
  for (int q = warp; q < kNQ; q += nw) {
    double t = 0.0;
    for (int g = lane; g < G; g += 32) t += *((volatile double*)&part[(size_t)g * kNQ + q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) c.red[q * 8] = t;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kNQ; ++q) v[q] = c.red[q * 8];
  __syncthreads();
  c.epoch++;
}

// ---- phase A: residual rows -> y-FFT -> transposed half spectrum ----
__device__ void phase_A(Ctx& c, int cd, double epsl) {
  const CoopArgs& a = c.a;
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye, nyh = a.nyh, TR = a.TR;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const int rows = K * nx, ntiles = (rows + TR - 1) / TR;
  const int ld = ny + 1;
  double2* buf0 = c.sm;
  double2* buf1 = buf0 + (size_t)(TR / 2) * ld;
  const double* rho = a.w.rho[cd];
  const double* al = a.w.alp[cd];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int egno = p.egno;
  const double dt = p.dt, dx = a.dxe, dy = a.dye;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TR;
    const int nrows = min(TR, rows - r0), npairs = (nrows + 1) >> 1;
    for (int idx = tid; idx < npairs * 2 * ny; idx += nth) {
      const int lr = idx / ny, j = idx - lr * ny;
      double res = 0.0;
      if (lr < nrows) {
        const int r = r0 + lr, k = r / nx, i = r - k * nx;
        const Nbr by = nbr(j, ny, 0);
        const size_t o = (size_t)k * n + (size_t)i * ny;
        const double r00 = rho[o + j];
        const double rnext = (k + 1 < K) ? rho[o + n + j] : 0.0;
        const double rym = rho[o + by.m], ryp = rho[o + by.p];
        const double cy0 = a.coef_ye[j];
        // y direction (the only one in 1-D): control arrays A-2, A-1
        const double* a1y = al + (size_t)(a.A - 2) * KN;
        const double* a2y = al + (size_t)(a.A - 1) * KN;
        double f1y_0, f1y_m, f2y_0, f2y_p;
        if (egno == 3) {
          const double xi = a.coef_xe[i];
          f1y_0 = f_plus(xi); f1y_m = f1y_0; f2y_0 = f_minus(xi); f2y_p = f2y_0;
        } else {
          f1y_0 = f_plus(-(cy0 * a1y[o + j]));
          f1y_m = f_plus(-(a.coef_ye[by.m] * a1y[o + by.m]));
          f2y_0 = f_minus(-(cy0 * a2y[o + j]));
          f2y_p = f_minus(-(a.coef_ye[by.p] * a2y[o + by.p]));
        }
        const double m1y_0 = (r00 + kRhoOffset) * f1y_0, m1y_m = (rym + kRhoOffset) * f1y_m;
        const double m2y_0 = (r00 + kRhoOffset) * f2y_0, m2y_p = (ryp + kRhoOffset) * f2y_p;
        if (a.has_x) {
          const Nbr bx = nbr(i, nx, p.bc_x);
          const size_t om = (size_t)k * n + (size_t)bx.m * ny, op = (size_t)k * n + (size_t)bx.p * ny;
          const double rxm = rho[om + j], rxp = rho[op + j];
          const double* a1x = al;
          const double* a2x = al + KN;
          double f1x_0, f1x_m, f2x_0, f2x_p;
          if (egno == 3) {
            f1x_0 = f_plus(a1x[o + j]); f1x_m = f_plus(a1x[om + j]);
            f2x_0 = f_minus(a2x[o + j]); f2x_p = f_minus(a2x[op + j]);
          } else {
            f1x_0 = f_plus(-(a.coef_xe[i] * a1x[o + j]));
            f1x_m = f_plus(-(a.coef_xe[bx.m] * a1x[om + j]));
            f2x_0 = f_minus(-(a.coef_xe[i] * a2x[o + j]));
            f2x_p = f_minus(-(a.coef_xe[bx.p] * a2x[op + j]));
          }
          const double m1x_0 = (r00 + kRhoOffset) * f1x_0, m1x_m = (rxm + kRhoOffset) * f1x_m;
          const double m2x_0 = (r00 + kRhoOffset) * f2x_0, m2x_p = (rxp + kRhoOffset) * f2x_p;
          res = (rnext - r00) / dt + epsl * ((rxp + rxm - 2 * r00) / (dx * dx)) + epsl * ((ryp + rym - 2 * r00) / (dy * dy));
          res -= bx.wm * (m1x_0 - m1x_m) / dx + bx.wp * (m2x_p - m2x_0) / dx + (m1y_0 - m1y_m) / dy + (m2y_p - m2y_0) / dy;
        } else {
          res = (rnext - r00) / dt + epsl * ((ryp + rym - 2 * r00) / (dy * dy));
          res -= (m1y_0 - m1y_m) / dy + (m2y_p - m2y_0) / dy;
        }
        if (k == K - 1) res += p.c_on_rho / dt;
      }
      reinterpret_cast<double*>(&buf0[(size_t)(lr >> 1) * ld + j])[lr & 1] = res;
    }
    __syncthreads();
    double2* zf = fft_rows(buf0, buf1, a.plan_ye, ld, a.tw_ye, npairs, 1.0);
    for (int idx = tid; idx < npairs * nyh; idx += nth) {
      const int ky = idx / npairs, pr = idx - ky * npairs;
      const int kym = (ky == 0) ? 0 : ny - ky;
      const double2 z1 = zf[(size_t)pr * ld + ky], z2 = zf[(size_t)pr * ld + kym];
      const int ra = r0 + 2 * pr, ka = ra / nx, ia = ra - ka * nx;
      a.w.zt[((size_t)ka * nyh + ky) * nx + ia] = make_double2(0.5 * (z1.x + z2.x), 0.5 * (z1.y - z2.y));
      if (2 * pr + 1 < nrows) {
        const int rb = ra + 1, kb = rb / nx, ib = rb - kb * nx;
        a.w.zt[((size_t)kb * nyh + ky) * nx + ib] = make_double2(0.5 * (z1.y + z2.y), 0.5 * (z2.x - z1.x));
      }
    }
    __syncthreads();
  }
}

// ---- phase B: x-FFT, t-solve per mode, inverse x-FFT (in place on zt) ----
__device__ void phase_B(Ctx& c) {
  const CoopArgs& a = c.a;
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, nyh = a.nyh;
  const int tid = threadIdx.x, nth = blockDim.x;
  const double ct2 = p.Ct_over_dt2;
  double2* zt = a.w.zt;
  if (nx == 1) {
    // 1-D: no x transform; thread per mode walks the t direction
    for (int ky = blockIdx.x * nth + tid; ky < nyh; ky += gridDim.x * nth) {
      if (K == 1 || ct2 == 0.0) {
        for (int k = 0; k < K; ++k) {
          const double d = p.diag[ky] + ((K == 1) ? ct2 : 0.0);
          double2 v = zt[(size_t)k * nyh + ky];
          zt[(size_t)k * nyh + ky] = make_double2(v.x / d, v.y / d);
        }
      } else {
        double2 bp = make_double2(0.0, 0.0);
        for (int k = 0; k < K; ++k) {
          const double den = a.w.den[(size_t)k * nyh + ky];
          double2 v = zt[(size_t)k * nyh + ky];
          bp = make_double2((v.x + ct2 * bp.x) / den, (v.y + ct2 * bp.y) / den);
          zt[(size_t)k * nyh + ky] = bp;
        }
        double2 xs = bp;
        for (int k = K - 2; k >= 0; --k) {
          const double t = a.w.tu[(size_t)k * nyh + ky];
          double2 v = zt[(size_t)k * nyh + ky];
          xs = make_double2(v.x - t * xs.x, v.y - t * xs.y);
          zt[(size_t)k * nyh + ky] = xs;
        }
      }
    }
    return;
  }
  const int ld = nx + 1;
  const int TKY = a.TKY;
  double2* buf0 = c.sm;
  double2* buf1 = buf0 + (size_t)TKY * ld;
  double2* carry = buf1 + (size_t)TKY * ld;
  const int ntile = (nyh + TKY - 1) / TKY;
  if (K == 1 || ct2 == 0.0) {
    // independent (k, ky) rows: FFT -> divide -> IFFT
    for (int u = blockIdx.x; u < K * ntile; u += gridDim.x) {
      const int k = u / ntile, ky0 = (u - k * ntile) * TKY;
      const int nr = min(TKY, nyh - ky0);
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        buf0[(size_t)t * ld + kx] = zt[((size_t)k * nyh + ky0 + t) * nx + kx];
      }
      __syncthreads();
      double2* zf = fft_rows(buf0, buf1, a.plan_xe, ld, a.tw_xe, nr, 1.0);
      double2* zo = (zf == buf0) ? buf1 : buf0;
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        const double d = p.diag[(size_t)kx * nyh + ky0 + t] + ((K == 1) ? ct2 : 0.0);
        double2 v = zf[(size_t)t * ld + kx];
        zf[(size_t)t * ld + kx] = make_double2(v.x / d, v.y / d);
      }
      __syncthreads();
      double2* zu = fft_rows(zf, zo, a.plan_xe, ld, a.tw_xe, nr, -1.0);
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        zt[((size_t)k * nyh + ky0 + t) * nx + kx] = zu[(size_t)t * ld + kx];
      }
      __syncthreads();
    }
    return;
  }
  // K > 1: forward elimination over k (carry = b_{k-1}), then back-substitution + inverse FFT
  for (int u = blockIdx.x; u < ntile; u += gridDim.x) {
    const int ky0 = u * TKY;
    const int nr = min(TKY, nyh - ky0);
    for (int k = 0; k < K; ++k) {
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        buf0[(size_t)t * ld + kx] = zt[((size_t)k * nyh + ky0 + t) * nx + kx];
      }
      __syncthreads();
      double2* zf = fft_rows(buf0, buf1, a.plan_xe, ld, a.tw_xe, nr, 1.0);
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        const size_t g = ((size_t)k * nyh + ky0 + t) * nx + kx;
        const double den = a.w.den[g];
        const double2 v = zf[(size_t)t * ld + kx];
        const double2 bp = (k == 0) ? make_double2(0.0, 0.0) : carry[(size_t)t * nx + kx];
        const double2 bn = make_double2((v.x + ct2 * bp.x) / den, (v.y + ct2 * bp.y) / den);
        carry[(size_t)t * nx + kx] = bn;
        zt[g] = bn;
      }
      __syncthreads();
    }
    for (int k = K - 1; k >= 0; --k) {
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        const size_t g = ((size_t)k * nyh + ky0 + t) * nx + kx;
        double2 xs;
        if (k == K - 1) {
          xs = carry[(size_t)t * nx + kx];
        } else {
          const double tuv = a.w.tu[g];
          const double2 v = zt[g], xn = carry[(size_t)t * nx + kx];
          xs = make_double2(v.x - tuv * xn.x, v.y - tuv * xn.y);
          carry[(size_t)t * nx + kx] = xs;
        }
        buf0[(size_t)t * ld + kx] = xs;
      }
      __syncthreads();
      double2* zu = fft_rows(buf0, buf1, a.plan_xe, ld, a.tw_xe, nr, -1.0);
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = idx / nx, kx = idx - t * nx;
        zt[((size_t)k * nyh + ky0 + t) * nx + kx] = zu[(size_t)t * ld + kx];
      }
      __syncthreads();
    }
  }
}

// ---- phase C: inverse y-FFT, phi update.  Adds to v[15] (sum dphi^2), v[16] (sum phi_prev^2 rows>=1), v[17] (NaN count)
__device__ void phase_C(Ctx& c, const double* phi_prev, double* phi_next, double* phib, double tau, double (&v)[kNQ]) {
  const CoopArgs& a = c.a;
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye, nyh = a.nyh, TR = a.TR;
  const size_t n = (size_t)nx * ny;
  const int rows = K * nx, ntiles = (rows + TR - 1) / TR;
  const int ld = ny + 1;
  double2* buf0 = c.sm;
  double2* buf1 = buf0 + (size_t)(TR / 2) * ld;
  const int tid = threadIdx.x, nth = blockDim.x;
  const double inv_nn = 1.0 / ((double)nx * (double)ny);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TR;
    const int nrows = min(TR, rows - r0), npairs = (nrows + 1) >> 1;
    for (int idx = tid; idx < npairs * nyh; idx += nth) {
      const int ky = idx / npairs, pr = idx - ky * npairs;
      const int ra = r0 + 2 * pr, ka = ra / nx, ia = ra - ka * nx;
      const double2 ua = a.w.zt[((size_t)ka * nyh + ky) * nx + ia];
      double2 ub = make_double2(0.0, 0.0);
      if (2 * pr + 1 < nrows) {
        const int rb = ra + 1, kb = rb / nx, ib = rb - kb * nx;
        ub = a.w.zt[((size_t)kb * nyh + ky) * nx + ib];
      }
      buf0[(size_t)pr * ld + ky] = make_double2(ua.x - ub.y, ua.y + ub.x);
      const int kym = ny - ky;
      if (ky != 0 && kym != ky) buf0[(size_t)pr * ld + kym] = make_double2(ua.x + ub.y, ub.x - ua.y);
    }
    __syncthreads();
    double2* zu = fft_rows(buf0, buf1, a.plan_ye, ld, a.tw_ye, npairs, -1.0);
    for (int idx = tid; idx < nrows * ny; idx += nth) {
      const int lr = idx / ny, j = idx - lr * ny;
      const int r = r0 + lr, k = r / nx, i = r - k * nx;
      const double2 z = zu[(size_t)(lr >> 1) * ld + j];
      const double u = ((lr & 1) ? z.y : z.x) * inv_nn;
      const size_t g = (size_t)(k + 1) * n + (size_t)i * ny + j;
      const double pp = phi_prev[g];
      const double pn = pp + tau * u;
      const double df = pn - pp;
      v[15] += df * df;
      v[16] += pp * pp;
      v[17] += is_nan(pn) ? 1.0 : 0.0;
      phi_next[g] = pn;
      if (phib) phib[g] = 2 * pn - pp;
    }
    __syncthreads();
  }
}

// ---- phase D: one dual sweep.  src -> dst (may alias); outer differences against `ref` when ref != nullptr.
// v[0..1] rho (diff^2, next^2), v[2+2j..3+2j] alp j; v[10] outer rho diff^2, v[11+j] outer alp diff^2; v[18] NaN count of rho_next
__device__ void phase_D(Ctx& c, const double* phib, const double* rho_s, const double* alp_s, double* rho_d, double* alp_d,
                        const double* rho_ref, const double* alp_ref, double sigma, double epsl, double (&v)[kNQ]) {
  const CoopArgs& a = c.a;
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye, egno = p.egno, A = a.A;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const double dt = p.dt, dx = a.dxe, dy = a.dye;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < KN; g += stride) {
    const int k = (int)(g / n);
    const size_t rem = g - (size_t)k * n;
    const int i = (int)(rem / ny), j = (int)(rem - (size_t)i * ny);
    const double* pb1 = phib + (size_t)(k + 1) * n;
    const Nbr by = nbr(j, ny, 0);
    const size_t row = (size_t)i * ny;
    const double c0 = pb1[row + j], cym = pb1[row + by.m], cyp = pb1[row + by.p];
    const double dyr = (cyp - c0) / dy, dyl = (c0 - cym) / dy;
    const double ro = rho_s[g];
    const double pinv = (ro + kRhoOffset) / sigma;
    double vec = (c0 - phib[(size_t)k * n + row + j]) / dt;
    double adv = 0.0, L = 0.0;
    double an[4], ao[4];
    if (a.has_x) {
      const Nbr bx = nbr(i, nx, p.bc_x);
      const double cxm = pb1[(size_t)bx.m * ny + j], cxp = pb1[(size_t)bx.p * ny + j];
      const double dxr = bx.wp * (cxp - c0) / dx, dxl = bx.wm * (c0 - cxm) / dx;
      const double cx = a.coef_xe[i];
      ao[0] = alp_s[g]; ao[1] = alp_s[KN + g];
      an[0] = prox_alp(egno, ao[0], dxr, pinv, cx, true);
      an[1] = prox_alp(egno, ao[1], dxl, pinv, cx, false);
      const double f1 = (egno == 3) ? f_plus(an[0]) : f_plus(-(cx * an[0]));
      const double f2 = (egno == 3) ? f_minus(an[1]) : f_minus(-(cx * an[1]));
      vec -= epsl * ((cxp + cxm - 2 * c0) / (dx * dx));
      adv = dxr * f1 + dxl * f2;
      L = lagr(egno, an[0]) + lagr(egno, an[1]);
    }
    {
      const int jy = A - 2;
      const double cy = a.coef_ye[j];
      ao[jy] = alp_s[(size_t)jy * KN + g]; ao[jy + 1] = alp_s[(size_t)(jy + 1) * KN + g];
      double f1, f2;
      if (egno == 3) {
        an[jy] = ao[jy]; an[jy + 1] = ao[jy + 1];          // set_fns.py:110: the y pair is passed through
        const double xi = a.coef_xe[i];
        f1 = f_plus(xi); f2 = f_minus(xi);
      } else {
        an[jy] = prox_alp(egno, ao[jy], dyr, pinv, cy, true);
        an[jy + 1] = prox_alp(egno, ao[jy + 1], dyl, pinv, cy, false);
        f1 = f_plus(-(cy * an[jy])); f2 = f_minus(-(cy * an[jy + 1]));
        L += lagr(egno, an[jy]) + lagr(egno, an[jy + 1]);
      }
      vec -= epsl * ((cyp + cym - 2 * c0) / (dy * dy));
      adv += dyr * f1 + dyl * f2;
    }
    vec -= adv;
    vec -= L;
    const double rn = relu_nan(ro + sigma * vec);
    rho_d[g] = rn;
    double d = rn - ro;
    v[0] += d * d; v[1] += rn * rn;
    v[18] += is_nan(rn) ? 1.0 : 0.0;
    if (rho_ref) { d = rn - rho_ref[g]; v[10] += d * d; }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < A) {
        alp_d[(size_t)q * KN + g] = an[q];
        d = an[q] - ao[q];
        v[2 + 2 * q] += d * d; v[3 + 2 * q] += an[q] * an[q];
        if (alp_ref) { d = an[q] - alp_ref[(size_t)q * KN + g]; v[11 + q] += d * d; }
      }
    }
  }
}

__device__ void grid_copy(double* dst, const double* src, size_t count) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < count; g += stride) dst[g] = src[g];
}

// Thomas tables (utils_precond.py:13-27 recurrences for dl = du = -Ct/dt^2, diagonal diag + Ct*[2,..,2,1]/dt^2)
__device__ void build_tables(const CoopArgs& a) {
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, nyh = a.nyh;
  const double ct2 = p.Ct_over_dt2;
  const size_t modes = (size_t)nyh * nx, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < modes; m += stride) {
    const int ky = (int)(m / nx), kx = (int)(m - (size_t)ky * nx);
    const double dg = p.diag[(size_t)kx * nyh + ky];
    double tprev = 0.0;
    for (int k = 0; k < K; ++k) {
      const double dk = dg + ((k == K - 1) ? ct2 : 2.0 * ct2);
      const double den = (k == 0) ? dk : dk + ct2 * tprev;
      tprev = ((k == K - 1) ? 0.0 : -ct2) / den;
      a.w.den[(size_t)k * modes + m] = den;
      a.w.tu[(size_t)k * modes + m] = tprev;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) pdhg_coop_kernel(const CoopArgs a) {
  extern __shared__ __align__(16) double2 dynsm[];
  __shared__ double red[kNQ * 8];
  Ctx c(a, dynsm, red);
  const MarchParams& p = a.p;
  const int tid = threadIdx.x, b = a.b;
  const int K = p.K, A = a.A;
  const size_t n = (size_t)a.nxe * a.nye, KN = (size_t)K * n, NP = (size_t)(K + 1) * n;
  const CoopWs& w = a.w;
  const bool lead = (blockIdx.x == 0 && tid == 0);

  if (a.mode == MODE_TABLES) { build_tables(a); return; }

  double* gphi = p.st_phi + (size_t)b * NP;
  double* grho = p.st_rho + (size_t)b * KN;
  double* galp = p.st_alp + (size_t)b * A * KN;
  const double epsl = p.epsl[b];

  if (a.mode == MODE_PRIMAL) {
    // phi_next = phi_prev + tau * H1_precond(cont_residual(rho, alp))   (update_fns_in_pdhg.py:135-147)
    grid_copy(w.rho[0], grho, KN);
    grid_copy(w.alp[0], galp, (size_t)A * KN);
    c.grid.sync();
    phase_A(c, 0, epsl);
    c.grid.sync();
    phase_B(c);
    c.grid.sync();
    double v[kNQ];
#pragma unroll
    for (int q = 0; q < kNQ; ++q) v[q] = 0.0;
    phase_C(c, a.op_phi_in, a.op_phi_out, nullptr, a.op_step, v);
    grid_copy(a.op_phi_out, a.op_phi_in, n);   // row 0 is invariant (u[0] = 0, utils_precond.py:139)
    return;
  }

  if (a.mode == MODE_DUAL) {
    // update_dual_alternative (update_fns_in_pdhg.py:167-180) on the state buffers, in place
    grid_copy(w.rho[0], grho, KN);
    grid_copy(w.alp[0], galp, (size_t)A * KN);
    c.grid.sync();
    int j = 0;
    double err = 0.0;
    for (; j < p.rho_alp_iters; ++j) {
      double v[kNQ];
#pragma unroll
      for (int q = 0; q < kNQ; ++q) v[q] = 0.0;
      const int s = j & 1, d = s ^ 1;
      phase_D(c, a.op_phi_in, w.rho[s], w.alp[s], w.rho[d], w.alp[d], nullptr, nullptr, a.op_step, epsl, v);
      grid_sum(c, v);
      err = v[0] / v[1];
      for (int q = 0; q < A; ++q) err += v[2 + 2 * q] / v[3 + 2 * q];
      if (err < a.op_eps) { ++j; break; }
    }
    const int fin = j & 1;     // after j sweeps the latest iterate sits in buffer (j & 1)
    grid_copy(grho, w.rho[fin], KN);
    grid_copy(galp, w.alp[fin], (size_t)A * KN);
    if (lead) { a.op_ninner[b] = j; a.op_err[b] = err; }
    return;
  }

  // ------------------------------ MODE_MARCH ------------------------------
  double stepsz = p.stepsz[b];
  int status = ST_OK;
  int blocks_done = p.blk_begin;
  long long inner_total = 0;
  const int nt_all = p.nblocks * K + 1;

  for (int blk = p.blk_begin; blk < p.blk_end && status == ST_OK; ++blk) {
    const size_t lb = (size_t)b * p.nblocks + blk;
    double* elog = p.errlog + lb * p.max_rec * kLogCols;
    while (true) {
      int cp = 0, cd = 0;     // current phi / dual buffers
      grid_copy(w.phi[0], gphi, NP);
      grid_copy(w.phib, gphi, n);            // phi_bar row 0 = phi row 0
      grid_copy(w.phi[1], gphi, n);
      grid_copy(w.rho[0], grho, KN);
      grid_copy(w.alp[0], galp, (size_t)A * KN);
      const double tau = stepsz / 1.5, sigma = stepsz * 1.5;
      // initial norms
      double S_row0, S_rho, S_alp[4];
      {
        double v[kNQ];
#pragma unroll
        for (int q = 0; q < kNQ; ++q) v[q] = 0.0;
        const size_t stride = (size_t)gridDim.x * blockDim.x, g0 = (size_t)blockIdx.x * blockDim.x + tid;
        for (size_t g = g0; g < n; g += stride) { const double x = gphi[g]; v[0] += x * x; }
        for (size_t g = g0; g < KN; g += stride) {
          const double r = grho[g]; v[1] += r * r;
          for (int q = 0; q < A; ++q) { const double x = galp[(size_t)q * KN + g]; v[2 + q] += x * x; }
        }
        grid_sum(c, v);    // (also orders the copies above before phase A)
        S_row0 = v[0]; S_rho = v[1];
        for (int q = 0; q < 4; ++q) S_alp[q] = v[2 + q];
      }
      long long it = p.iter_begin;
      int reason = END_MAXITER, nrec = 0;
      bool lognan = false, logfull = false;
      double err1 = 0.0, err2 = 0.0, rmin = 0.0, rmax = 0.0;

      for (; it < p.n_maxiter; ++it) {
        if (it >= p.iter_pause) { reason = END_PAUSED; break; }
        phase_A(c, cd, epsl);
        c.grid.sync();
        phase_B(c);
        c.grid.sync();
        double v[kNQ];
#pragma unroll
        for (int q = 0; q < kNQ; ++q) v[q] = 0.0;
        phase_C(c, w.phi[cp], w.phi[cp ^ 1], w.phib, tau, v);
        c.grid.sync();
        // dual sweeps: the first goes cd -> cd^1, the rest in place on cd^1 with outer differences against cd
        const int nd = cd ^ 1;
        double e1s0 = 0.0, e1s1 = 0.0, e1nan = 0.0;
        int j = 0;
        for (; j < p.rho_alp_iters; ++j) {
          if (j > 0) {
#pragma unroll
            for (int q = 0; q < kNQ; ++q) v[q] = 0.0;
          }
          if (j == 0) phase_D(c, w.phib, w.rho[cd], w.alp[cd], w.rho[nd], w.alp[nd], nullptr, nullptr, sigma, epsl, v);
          else phase_D(c, w.phib, w.rho[nd], w.alp[nd], w.rho[nd], w.alp[nd], w.rho[cd], w.alp[cd], sigma, epsl, v);
          grid_sum(c, v);
          if (j == 0) { e1s0 = v[15]; e1s1 = v[16]; e1nan = v[17]; }
          double err = v[0] / v[1];
          for (int q = 0; q < A; ++q) err += v[2 + 2 * q] / v[3 + 2 * q];
          if (err < p.eps) { ++j; break; }
        }
        inner_total += j;
        const bool multi = (j > 1);
        err1 = sqrt(e1s0) / sqrt(S_row0 + e1s1);
        err2 = sqrt(multi ? v[10] : v[0]) / sqrt(S_rho);
        for (int q = 0; q < A; ++q) {
          const double na = sqrt(S_alp[q]), ne = sqrt(multi ? v[11 + q] : v[2 + 2 * q]);
          if (na < 1e-6 && ne > 1e-6) err2 += ne; else if (na >= 1e-6) err2 += ne / na;
        }
        S_rho = v[1];
        for (int q = 0; q < A; ++q) S_alp[q] = v[3 + 2 * q];
        const bool anynan = (e1nan > 0.0) || (v[18] > 0.0);
        cp ^= 1; cd = nd;       // accept phi_next, rho_next, alp_next
        if (err1 < p.eps && err2 < p.eps) { reason = END_CONVERGED; break; }
        if (anynan) { reason = END_NAN; break; }
        if (p.print_freq > 0 && it % p.print_freq == 0) {
          // min / max of rho_next for the log line (utils_pdhg_solver.py:84-85)
          double mn = 1e300, mx = -1e300;
          const size_t stride = (size_t)gridDim.x * blockDim.x;
          for (size_t g = (size_t)blockIdx.x * blockDim.x + tid; g < KN; g += stride) { const double r = w.rho[cd][g]; mn = fmin(mn, r); mx = fmax(mx, r); }
          for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
          __syncthreads();
          if ((tid & 31) == 0) { red[tid >> 5] = mn; red[8 + (tid >> 5)] = mx; }
          __syncthreads();
          double* part = w.partials + (size_t)(c.epoch & 1) * gridDim.x * kNQ;
          if (tid == 0) {
            for (int q = 1; q < (int)(blockDim.x >> 5); ++q) { mn = fmin(mn, red[q]); mx = fmax(mx, red[8 + q]); }
            part[(size_t)blockIdx.x * kNQ] = mn; part[(size_t)blockIdx.x * kNQ + 1] = mx;
          }
          c.grid.sync();
          mn = 1e300; mx = -1e300;
          for (int g = 0; g < (int)gridDim.x; ++g) {
            mn = fmin(mn, *((volatile double*)&part[(size_t)g * kNQ])); mx = fmax(mx, *((volatile double*)&part[(size_t)g * kNQ + 1]));
          }
          c.epoch++;
          rmin = mn; rmax = mx;
          if (nrec < p.max_rec - 1) {
            if (lead) { elog[nrec * kLogCols] = err1; elog[nrec * kLogCols + 1] = err2; elog[nrec * kLogCols + 2] = rmin; elog[nrec * kLogCols + 3] = rmax; }
            ++nrec;
          } else {
            logfull = true;
          }
          lognan = lognan || is_nan(err1) || is_nan(err2);
        }
      }
      const long long iters_done = (reason == END_MAXITER) ? p.n_maxiter : ((reason == END_PAUSED) ? it : it + 1);
      if (reason != END_PAUSED) {
        if (lead) { elog[nrec * kLogCols] = err1; elog[nrec * kLogCols + 1] = err2; elog[nrec * kLogCols + 2] = rmin; elog[nrec * kLogCols + 3] = rmax; }
        ++nrec;
        lognan = lognan || is_nan(err1) || is_nan(err2);
      }
      if (p.fallback && lognan) {
        if (stepsz > p.stepsz_floor[b]) { stepsz -= p.stepsz_delta[b]; c.grid.sync(); continue; }
        status = ST_SOL_NAN;
        break;
      }
      if (lead) { p.iters[lb] = iters_done; p.stepsz_used[lb] = stepsz; p.nrec[lb] = nrec; p.end_reason[lb] = reason; }
      if (logfull) status = ST_LOG_OVERFLOW;
      if (reason == END_PAUSED) status = ST_PAUSED;
      const double* phi = w.phi[cp];
      const double* rho = w.rho[cd];
      const double* alp = w.alp[cd];
      const size_t stride = (size_t)gridDim.x * blockDim.x, g0 = (size_t)blockIdx.x * blockDim.x + tid;
      if (p.handoff && reason != END_PAUSED) {
        double* pa = p.phi_all + ((size_t)b * nt_all + (size_t)blk * K) * n;
        const size_t nrows = (blk == p.nblocks - 1) ? K + 1 : K;
        grid_copy(pa, phi, nrows * n);
        grid_copy(p.rho_all + ((size_t)b * (nt_all - 1) + (size_t)blk * K) * n, rho, KN);
        for (int q = 0; q < A; ++q)
          grid_copy(p.alp_all + (((size_t)b * A + q) * (nt_all - 1) + (size_t)blk * K) * n, alp + (size_t)q * KN, KN);
        // phi0 <- phi0 + (phi_curr[-1] - phi0[0])   (utils_pdhg_solver.py:200-203); phi row 0 == phi0 row 0
        for (size_t g = g0; g < NP; g += stride) {
          const size_t x = g % n;
          gphi[g] = gphi[g] + (phi[(size_t)K * n + x] - phi[x]);
        }
      } else {
        grid_copy(gphi, phi, NP);
      }
      grid_copy(grho, rho, KN);
      grid_copy(galp, alp, (size_t)A * KN);
      c.grid.sync();
      if (reason != END_PAUSED) blocks_done = blk + 1;
      break;
    }
  }
  if (lead) {
    p.status[b] = status;
    p.blocks_done[b] = blocks_done;
    p.stepsz[b] = stepsz;
    p.inner_total[b] = inner_total;
  }
}

// ------------------------------------------- host side -------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct CoopGeom { int nxe, nye, nyh, TR, TKY, grid; size_t smem; };

static CoopGeom coop_geom(const MarchParams& p, int sm_count, size_t smem_cap) {
  CoopGeom g;
  g.nxe = (p.ndim == 1) ? 1 : p.nx;
  g.nye = (p.ndim == 1) ? p.nx : p.ny;
  g.nyh = g.nye / 2 + 1;
  const int rows = p.K * g.nxe;
  // y-FFT tile: TR rows (even) -> TR/2 complex transforms, two buffers of (TR/2)*(nye+1) complex
  int TR = 16;
  while (TR > 2 && ((rows + TR - 1) / TR < 2 * sm_count || (size_t)TR * (g.nye + 1) * 16 > smem_cap)) TR -= 2;
  g.TR = TR;
  int TKY = 4;
  while (TKY > 1 && ((g.nyh + TKY - 1) / TKY < 2 * sm_count || (size_t)3 * TKY * (g.nxe + 1) * 16 > smem_cap)) TKY -= 1;
  g.TKY = TKY;
  const size_t smA = (size_t)TR * (g.nye + 1) * 16;                 // 2 buffers * TR/2 rows
  const size_t smB = (g.nxe > 1) ? (size_t)3 * TKY * (g.nxe + 1) * 16 : 0;
  g.smem = smA > smB ? smA : smB;
  g.grid = sm_count;
  return g;
}

size_t pdhg_coop_workspace_bytes(const MarchParams& p, int B) {
  (void)B;
  const int nxe = (p.ndim == 1) ? 1 : p.nx, nye = (p.ndim == 1) ? p.nx : p.ny, nyh = nye / 2 + 1, A = 2 * p.ndim;
  const size_t n = (size_t)nxe * nye, KN = (size_t)p.K * n, NP = (size_t)(p.K + 1) * n, modes = (size_t)nyh * nxe;
  size_t bytes = 0;
  bytes += align_up(3 * NP * 8, 256);
  bytes += align_up(2 * KN * 8, 256);
  bytes += align_up(2 * A * KN * 8, 256);
  bytes += align_up((size_t)p.K * modes * 16, 256);
  bytes += align_up((size_t)2 * 1024 * kNQ * 8, 256);
  bytes += align_up(2 * (size_t)p.K * modes * 8, 256);
  return bytes + 1024;
}

static CoopWs carve(const MarchParams& p, void* ws) {
  const int nxe = (p.ndim == 1) ? 1 : p.nx, nye = (p.ndim == 1) ? p.nx : p.ny, nyh = nye / 2 + 1, A = 2 * p.ndim;
  const size_t n = (size_t)nxe * nye, KN = (size_t)p.K * n, NP = (size_t)(p.K + 1) * n, modes = (size_t)nyh * nxe;
  char* q = static_cast<char*>(ws);
  CoopWs w;
  w.phi[0] = (double*)q; w.phi[1] = w.phi[0] + NP; w.phib = w.phi[1] + NP; q += align_up(3 * NP * 8, 256);
  w.rho[0] = (double*)q; w.rho[1] = w.rho[0] + KN; q += align_up(2 * KN * 8, 256);
  w.alp[0] = (double*)q; w.alp[1] = w.alp[0] + (size_t)A * KN; q += align_up(2 * A * KN * 8, 256);
  w.zt = (double2*)q; q += align_up((size_t)p.K * modes * 16, 256);
  w.partials = (double*)q; q += align_up((size_t)2 * 1024 * kNQ * 8, 256);
  w.den = (double*)q; w.tu = w.den + (size_t)p.K * modes;
  return w;
}

static cudaError_t coop_launch(const MarchParams& p, void* ws, int b, int mode, const double* op_in, double* op_out,
                               double op_step, double op_eps, int* op_ninner, double* op_err, cudaStream_t stream) {
  int dev = 0, sms = 0, smem_cap = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const CoopGeom g = coop_geom(p, sms, (size_t)smem_cap - 4096);
  if (g.smem > (size_t)smem_cap - 2048) return cudaErrorInvalidConfiguration;
  CoopArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  a.w = carve(p, ws);
  a.b = b; a.mode = mode; a.A = 2 * p.ndim;
  a.nxe = g.nxe; a.nye = g.nye; a.nyh = g.nyh; a.TR = g.TR; a.TKY = g.TKY;
  a.has_x = (p.ndim == 2);
  if (p.ndim == 1) {
    a.dxe = 1.0; a.dye = p.dx; a.coef_xe = p.coef_x; a.coef_ye = p.coef_x; a.tw_xe = p.tw_x; a.tw_ye = p.tw_x;
    a.plan_ye = p.plan_x; a.plan_xe.n = 1; a.plan_xe.nstages = 0;
  } else {
    a.dxe = p.dx; a.dye = p.dy; a.coef_xe = p.coef_x; a.coef_ye = p.coef_y; a.tw_xe = p.tw_x; a.tw_ye = p.tw_y;
    a.plan_xe = p.plan_x; a.plan_ye = p.plan_y;
  }
  a.op_phi_in = op_in; a.op_phi_out = op_out; a.op_step = op_step; a.op_eps = op_eps; a.op_ninner = op_ninner; a.op_err = op_err;
  e = cudaFuncSetAttribute(pdhg_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  void* args[] = {&a};
  return cudaLaunchCooperativeKernel((void*)pdhg_coop_kernel, dim3(g.grid), dim3(kThreads), args, g.smem, stream);
}

static cudaError_t ensure_tables(const MarchParams& p, void* ws, cudaStream_t stream, long long* launches) {
  // tables depend only on the handle's constants; rebuilt per call (cheap) so the workspace needs no extra state
  if (p.K == 1 || p.Ct_over_dt2 == 0.0) return cudaSuccess;
  ++*launches;
  return coop_launch(p, ws, 0, MODE_TABLES, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, stream);
}

cudaError_t launch_pdhg_coop(const MarchParams& p, int B, void* ws, cudaStream_t stream, long long* launches) {
  cudaError_t e = ensure_tables(p, ws, stream, launches);
  for (int b = 0; b < B && e == cudaSuccess; ++b) {
    e = coop_launch(p, ws, b, MODE_MARCH, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, stream);
    ++*launches;
  }
  return e;
}

cudaError_t launch_update_primal(const MarchParams& p, int B, const double* phi_prev, double tau, double* phi_next, void* ws,
                                 cudaStream_t stream, long long* launches) {
  const size_t NP = (size_t)(p.K + 1) * p.nx * p.ny;
  cudaError_t e = ensure_tables(p, ws, stream, launches);
  for (int b = 0; b < B && e == cudaSuccess; ++b) {
    e = coop_launch(p, ws, b, MODE_PRIMAL, phi_prev + b * NP, phi_next + b * NP, tau, 0.0, nullptr, nullptr, stream);
    ++*launches;
  }
  return e;
}

cudaError_t launch_update_dual(const MarchParams& p, int B, const double* phi_bar, double sigma, double eps, int* n_inner_dev,
                               double* err_dev, void* ws, cudaStream_t stream, long long* launches) {
  const size_t NP = (size_t)(p.K + 1) * p.nx * p.ny;
  cudaError_t e = cudaSuccess;
  for (int b = 0; b < B && e == cudaSuccess; ++b) {
    e = coop_launch(p, ws, b, MODE_DUAL, phi_bar + b * NP, nullptr, sigma, eps, n_inner_dev, err_dev, stream);
    ++*launches;
  }
  return e;
}

}  // namespace pdhg
