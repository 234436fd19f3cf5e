// 1-D, time_step_per_PDHG = 2 (K = 1) specialisation of the single-CTA solver — the shape of every 1-D BASELINE config.
//
// Same contract as pdhg1d_cta_kernel (whole PDHG_multi_step on chip, one CTA per instance; reference lines:
// jaxsrc/utils/utils_pdhg_solver.py:9-225, jaxsrc/update_fns_in_pdhg.py:72-180), restructured for latency:
//   * every thread OWNS NPT grid points and keeps their phi, rho, alp1, alp2 in registers for the whole block solve;
//     shared memory only carries what neighbours need (rho, m1, m2, the source term, phi_bar);
//   * the H1 preconditioner for K = 1 is a per-mode scaling u^ = s^/d (utils_precond.py:122-136), i.e. the circular
//     convolution u = G (*) s with G = IFFT(1/d).  G decays like r^|j| (r ~ 0.61 on the BASELINE grids),
//     so it is applied directly with the taps |j| <= R beyond which |G| < 2^-58 G[0] (below fp64 rounding of the
//     result: exact to rounding).  G and R are computed on the host in long double.  This replaces two FFTs (6-8
//     barrier-separated stages) by one barrier-free pass of R+1 fused multiply-adds per point;
//   * one block reduction per outer iteration (9 sums: err1 pair, inner/outer dual sums) instead of two.
// The host falls back to the general kernel when R would exceed nx/2 - 1 or kMaxTaps (slowly decaying kernels).
#include "pdhg_params.h"

namespace pdhg {

constexpr int kMaxTaps = 160;

template <int NPT>
__global__ void __launch_bounds__(512, 1) pdhg1d_k1_kernel(const MarchParams p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int b = blockIdx.x;
  const int nx = p.nx, R = p.green_R;
  const int egno = p.egno;

  double* rho_s = sm;                    // [nx]   rho          (neighbour exchange)
  double* m1_s = rho_s + nx;             // [nx]   (rho+1e-4) f+  of alp1
  double* m2_s = m1_s + nx;              // [nx]   (rho+1e-4) f-  of alp2
  double* pb_s = m2_s + nx;              // [nx]   phi_bar row 1
  double* s_ext = pb_s + nx;             // [nx + 2R] source term with periodic halos
  double* g_s = s_ext + nx + 2 * R;      // [R + 1] Green's function taps
  double* red = g_s + R + 1;             // 9 * 32 reduction scratch

  double* gphi = p.st_phi + (size_t)b * 2 * nx;
  double* grho = p.st_rho + (size_t)b * nx;
  double* galp = p.st_alp + (size_t)b * 2 * nx;
  const double epsl = p.epsl[b];
  double stepsz = p.stepsz[b];
  const double dt = p.dt, dx = p.dx, eps = p.eps;
  const int nt_all = p.nblocks + 1;

  for (int j = tid; j <= R; j += nth) g_s[j] = p.green[j];

  int xs[NPT];
  bool own[NPT];
  double cf[NPT];
#pragma unroll
  for (int m = 0; m < NPT; ++m) {
    xs[m] = tid + m * nth;
    own[m] = xs[m] < nx;
    if (!own[m]) xs[m] = nx - 1;          // harmless duplicate; excluded from sums and stores
    cf[m] = p.coef_x[xs[m]];
  }

  int status = ST_OK;
  bool log_overflow = false;     // sticky: more periodic records than max_rec - 1 in some block (rows dropped, march goes on)
  int blocks_done = p.blk_begin;
  long long inner_total = 0;

  for (int blk = p.blk_begin; blk < p.blk_end && status == ST_OK; ++blk) {
    const size_t lb = (size_t)b * p.nblocks + blk;
    double* elog = p.errlog + lb * p.max_rec * kLogCols;
    while (true) {   // step-size fallback loop (utils_pdhg_solver.py:174-187)
      double ph0[NPT], ph1[NPT], rho[NPT], a1[NPT], a2[NPT];
#pragma unroll
      for (int m = 0; m < NPT; ++m) {
        ph0[m] = gphi[xs[m]]; ph1[m] = gphi[nx + xs[m]];
        rho[m] = grho[xs[m]]; a1[m] = galp[xs[m]]; a2[m] = galp[nx + xs[m]];
      }
      const double tau = stepsz / 1.5, sigma = stepsz * 1.5;   // utils_pdhg_solver.py:44-46
      const Recip rc(dt, dx, 1.0, sigma);
      const double c_dt = p.c_on_rho * rc.idt;

      double S_row0, S_rho, S_a1, S_a2;
      {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int m = 0; m < NPT; ++m) if (own[m]) { s[0] += ph0[m] * ph0[m]; s[1] += rho[m] * rho[m]; s[2] += a1[m] * a1[m]; s[3] += a2[m] * a2[m]; }
        block_sum<4>(s, red);
        S_row0 = s[0]; S_rho = s[1]; S_a1 = s[2]; S_a2 = s[3];
      }

      long long it = p.iter_begin;
      int reason = END_MAXITER, nrec = 0;
      bool lognan = false, logfull = false;
      double err1 = 0.0, err2 = 0.0, rmin = 0.0, rmax = 0.0;
      double v[9];

      for (; it < p.n_maxiter; ++it) {
        if (it >= p.iter_pause) { reason = END_PAUSED; break; }

        // ---- publish rho, m1, m2 (update_fns_in_pdhg.py:74-77) ----
#pragma unroll
        for (int m = 0; m < NPT; ++m) {
          if (own[m]) {
            rho_s[xs[m]] = rho[m];
            m1_s[xs[m]] = (rho[m] + kRhoOffset) * f_plus(-(cf[m] * a1[m]));
            m2_s[xs[m]] = (rho[m] + kRhoOffset) * f_minus(-(cf[m] * a2[m]));
          }
        }
        __syncthreads();
        // ---- continuity residual row 1 (update_fns_in_pdhg.py:78-81) -> s_ext (with periodic halos) ----
#pragma unroll
        for (int m = 0; m < NPT; ++m) {
          if (own[m]) {
            const int x = xs[m];
            const int xm = (x == 0) ? nx - 1 : x - 1, xp = (x == nx - 1) ? 0 : x + 1;
            const double r0 = rho[m], rm = rho_s[xm], rp = rho_s[xp];
            double res = (0.0 - r0) * rc.idt + epsl * ((rp + rm - 2 * r0) * rc.idx2);
            res -= (m1_s[x] - m1_s[xm]) * rc.idx + (m2_s[xp] - m2_s[x]) * rc.idx;
            res += c_dt;
            s_ext[R + x] = res;
            if (x < R) s_ext[R + nx + x] = res;
            if (x >= nx - R) s_ext[x - (nx - R)] = res;
          }
        }
        __syncthreads();
        // ---- u = G (*) s ; phi_next = phi_prev + tau u ; phi_bar = 2 phi_next - phi_prev (utils_pdhg_solver.py:53-55) ----
        int nanflag = 0;
        double e_d = 0.0, e_p = 0.0;
        double pbar[NPT];
        {
          double acc[NPT][2];
#pragma unroll
          for (int m = 0; m < NPT; ++m) { acc[m][0] = g_s[0] * s_ext[R + xs[m]]; acc[m][1] = 0.0; }
          for (int j = 1; j <= R; ++j) {
            const double g = g_s[j];
#pragma unroll
            for (int m = 0; m < NPT; ++m) {
              acc[m][0] = fma(g, s_ext[R + xs[m] + j], acc[m][0]);
              acc[m][1] = fma(g, s_ext[R + xs[m] - j], acc[m][1]);
            }
          }
#pragma unroll
          for (int m = 0; m < NPT; ++m) {
            const double u = acc[m][0] + acc[m][1];
            const double pp = ph1[m];
            const double pn = pp + tau * u;
            const double df = pn - pp;
            if (own[m]) { e_d += df * df; e_p += pp * pp; nanflag |= is_nan(pn) ? 1 : 0; }
            ph1[m] = pn;
            pbar[m] = 2 * pn - pp;
            if (own[m]) pb_s[xs[m]] = pbar[m];
          }
        }
        __syncthreads();

        // ---- dual sweeps (update_fns_in_pdhg.py:150-180); state in registers, outer-iteration copies kept for the norms ----
        double ro_o[NPT], a1_o[NPT], a2_o[NPT];
        double cl[NPT], cr[NPT];
#pragma unroll
        for (int m = 0; m < NPT; ++m) {
          ro_o[m] = rho[m]; a1_o[m] = a1[m]; a2_o[m] = a2[m];
          const int x = xs[m];
          cl[m] = pb_s[(x == 0) ? nx - 1 : x - 1];
          cr[m] = pb_s[(x == nx - 1) ? 0 : x + 1];
        }
        int j = 0;
        for (; j < p.rho_alp_iters; ++j) {
#pragma unroll
          for (int q = 0; q < 9; ++q) v[q] = 0.0;
          v[6] = e_d; v[7] = e_p;               // ride along with the first reduction (constant over the sweeps)
#pragma unroll
          for (int m = 0; m < NPT; ++m) {
            const double c0 = pbar[m];
            const double dxr = (cr[m] - c0) * rc.idx, dxl = (c0 - cl[m]) * rc.idx;
            const double ro = rho[m], a1p = a1[m], a2p = a2[m];
            const double pinv = (ro + kRhoOffset) * rc.isig;
            const double rinv = prox_rinv(egno, pinv);
            const double a1n = prox_alp(egno, a1p, dxr, pinv, rinv, cf[m], true);
            const double a2n = prox_alp(egno, a2p, dxl, pinv, rinv, cf[m], false);
            const double f1 = f_plus(-(cf[m] * a1n)), f2 = f_minus(-(cf[m] * a2n));
            double vec = (c0 - ph0[m]) * rc.idt - epsl * ((cr[m] + cl[m] - 2 * c0) * rc.idx2);
            vec -= dxr * f1 + dxl * f2;
            vec -= lagr(egno, a1n) + lagr(egno, a2n);
            const double rn = relu_nan(ro + sigma * vec);
            rho[m] = rn; a1[m] = a1n; a2[m] = a2n;
            if (own[m]) {
              double d;
              d = rn - ro;   v[0] += d * d; v[1] += rn * rn;
              d = a1n - a1p; v[2] += d * d; v[3] += a1n * a1n;
              d = a2n - a2p; v[4] += d * d; v[5] += a2n * a2n;
              nanflag |= is_nan(rn) ? 2 : 0;
            }
          }
          block_sum<9>(v, red);
          if (j == 0) { e_d = v[6]; e_p = v[7]; }
          const double err = v[0] / v[1] + v[2] / v[3] + v[4] / v[5];
          if (err < eps) { ++j; break; }
        }
        inner_total += j;
        double d_rho = v[0], d_a1 = v[2], d_a2 = v[4];
        if (j > 1) {
          // outer differences against the iterate the sweeps started from (rare: > 1 sweep)
          double w[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int m = 0; m < NPT; ++m) {
            if (own[m]) {
              double d;
              d = rho[m] - ro_o[m]; w[0] += d * d;
              d = a1[m] - a1_o[m];  w[1] += d * d;
              d = a2[m] - a2_o[m];  w[2] += d * d;
            }
          }
          block_sum<3>(w, red);
          d_rho = w[0]; d_a1 = w[1]; d_a2 = w[2];
        }
        // ---- outer errors and exits (utils_pdhg_solver.py:59-80) ----
        const double den1 = S_row0 + e_p;
        const bool rec = (p.print_freq > 0 && it % p.print_freq == 0);
        nanflag = __syncthreads_or(nanflag);
        // cheap screen: sqrt(e_d)/sqrt(den1) >= eps certainly (4-ulp margin) => not converged, exact values not needed
        const bool far = (e_d > eps * eps * den1 * (1.0 + 1e-12)) && !rec && !nanflag && (it + 1 < p.n_maxiter);
        if (!far) {
          err1 = sqrt(e_d) / sqrt(den1);
          err2 = sqrt(d_rho) / sqrt(S_rho);
          {
            const double na = sqrt(S_a1), ne = sqrt(d_a1);
            if (na < 1e-6 && ne > 1e-6) err2 += ne; else if (na >= 1e-6) err2 += ne / na;
          }
          {
            const double na = sqrt(S_a2), ne = sqrt(d_a2);
            if (na < 1e-6 && ne > 1e-6) err2 += ne; else if (na >= 1e-6) err2 += ne / na;
          }
        }
        S_rho = v[1]; S_a1 = v[3]; S_a2 = v[5];
        if (!far) {
          if (err1 < eps && err2 < eps) { reason = END_CONVERGED; break; }
          if (nanflag) { reason = END_NAN; break; }
          if (rec) {
            double mn = 1e300, mx = -1e300;
#pragma unroll
            for (int m = 0; m < NPT; ++m) if (own[m]) { mn = fmin(mn, rho[m]); mx = fmax(mx, rho[m]); }
            for (int o = 16; o > 0; o >>= 1) {
              mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
              mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            __syncthreads();
            if ((tid & 31) == 0) { red[tid >> 5] = mn; red[32 + (tid >> 5)] = mx; }
            __syncthreads();
            rmin = red[0]; rmax = red[32];
            for (int w = 1; w < ((nth + 31) >> 5); ++w) { rmin = fmin(rmin, red[w]); rmax = fmax(rmax, red[32 + w]); }
            if (nrec < p.max_rec - 1) {
              if (tid == 0) {
                elog[nrec * kLogCols + 0] = err1; elog[nrec * kLogCols + 1] = err2;
                elog[nrec * kLogCols + 2] = rmin; elog[nrec * kLogCols + 3] = rmax;
              }
              ++nrec;
            } else {
              logfull = true;
            }
            lognan = lognan || is_nan(err1) || is_nan(err2);
          }
        }
      }
      const long long iters_done = (reason == END_MAXITER) ? p.n_maxiter : ((reason == END_PAUSED) ? it : it + 1);
      if (reason != END_PAUSED) {
        if (tid == 0) {
          elog[nrec * kLogCols + 0] = err1; elog[nrec * kLogCols + 1] = err2;
          elog[nrec * kLogCols + 2] = rmin; elog[nrec * kLogCols + 3] = rmax;
        }
        ++nrec;
        lognan = lognan || is_nan(err1) || is_nan(err2);
      }
      if (p.fallback && lognan) {
        if (stepsz > p.stepsz_floor[b]) { stepsz -= p.stepsz_delta[b]; __syncthreads(); continue; }
        status = ST_SOL_NAN;
        break;
      }
      if (tid == 0) { p.iters[lb] = iters_done; p.stepsz_used[lb] = stepsz; p.nrec[lb] = nrec; p.end_reason[lb] = reason; }
      log_overflow = log_overflow || logfull;
      if (reason == END_PAUSED) status = ST_PAUSED;
      if (p.handoff && reason != END_PAUSED) {
        double* pa = p.phi_all + ((size_t)b * nt_all + (size_t)blk) * nx;
        double* ra = p.rho_all + ((size_t)b * (nt_all - 1) + (size_t)blk) * nx;
        double* aa1 = p.alp_all + (((size_t)b * 2 + 0) * (nt_all - 1) + (size_t)blk) * nx;
        double* aa2 = p.alp_all + (((size_t)b * 2 + 1) * (nt_all - 1) + (size_t)blk) * nx;
#pragma unroll
        for (int m = 0; m < NPT; ++m) {
          if (own[m]) {
            const int x = xs[m];
            pa[x] = ph0[m];
            if (blk == p.nblocks - 1) pa[nx + x] = ph1[m];
            ra[x] = rho[m]; aa1[x] = a1[m]; aa2[x] = a2[m];
            // warm start (utils_pdhg_solver.py:200-205): phi0 <- phi0 + (phi_curr[-1] - phi0[0]), row by row
            const double gd = ph1[m] - ph0[m];
            const double r1 = gphi[nx + x];
            gphi[x] = ph0[m] + gd;
            gphi[nx + x] = r1 + gd;
            grho[x] = rho[m]; galp[x] = a1[m]; galp[nx + x] = a2[m];
          }
        }
      } else {
#pragma unroll
        for (int m = 0; m < NPT; ++m) {
          if (own[m]) {
            const int x = xs[m];
            gphi[x] = ph0[m]; gphi[nx + x] = ph1[m];
            grho[x] = rho[m]; galp[x] = a1[m]; galp[nx + x] = a2[m];
          }
        }
      }
      __syncthreads();
      if (reason != END_PAUSED) blocks_done = blk + 1;
      break;
    }
  }
  if (tid == 0) {
    p.status[b] = (status == ST_OK && log_overflow) ? ST_LOG_OVERFLOW : status;   // overflow never stops the march (only NaN failure and pause do)
    p.blocks_done[b] = blocks_done;
    p.stepsz[b] = stepsz;
    p.inner_total[b] = inner_total;
  }
}

static int k1_npt(int nx) { return nx <= 512 ? 1 : (nx <= 1024 ? 2 : 4); }   // most threads that fit: the solve is latency-bound

bool pdhg1d_k1_supported(int nx, int K, int green_R) {
  return K == 1 && green_R >= 0 && green_R <= kMaxTaps && green_R < nx / 2 && nx <= 2048;
}

cudaError_t launch_pdhg1d_k1(const MarchParams& p, int B, cudaStream_t stream) {
  const int npt = k1_npt(p.nx);
  int nth = ((p.nx + npt - 1) / npt + 31) / 32 * 32;
  if (nth < 64) nth = 64;
  const size_t smem = ((size_t)5 * p.nx + 2 * p.green_R + p.green_R + 1 + 9 * 32) * sizeof(double);
  cudaError_t e;
  if (npt == 1) {
    e = cudaFuncSetAttribute(pdhg1d_k1_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pdhg1d_k1_kernel<1><<<B, nth, smem, stream>>>(p);
  } else if (npt == 2) {
    e = cudaFuncSetAttribute(pdhg1d_k1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pdhg1d_k1_kernel<2><<<B, nth, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(pdhg1d_k1_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pdhg1d_k1_kernel<4><<<B, nth, smem, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace pdhg
