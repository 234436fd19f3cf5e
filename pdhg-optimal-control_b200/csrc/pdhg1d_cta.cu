// 1-D time-block PDHG solver: ONE CTA PER PROBLEM INSTANCE, whole solve resident in shared memory.
//
// The CTA runs, without any host round trip, what the reference does with a Python loop and ~25 XLA
// dispatches + ~8 blocking scalar reads per iteration:
//   PDHG_multi_step      jaxsrc/utils/utils_pdhg_solver.py:97-225  (block marching, warm start, NaN fallback)
//   PDHG_solver_oneiter  jaxsrc/utils/utils_pdhg_solver.py:9-94    (outer loop, err1/err2, exits, records)
//   update_primal_1d     jaxsrc/update_fns_in_pdhg.py:135-140      (continuity residual -> H1 preconditioner -> phi)
//   update_dual_alternative / update_dual_oneiter   jaxsrc/update_fns_in_pdhg.py:150-180  (<=10 alp/rho sweeps)
// State per instance: phi[(K+1) nx], phi_bar, rho[K nx], alp1, alp2 + complex FFT workspace, all in smem
// ((9K+2) nx doubles).  The x-FFT is a mixed-radix Stockham transform in smem; the t-direction solve is a
// per-mode Thomas recurrence (one thread per Fourier mode).
#include "pdhg_params.h"

namespace pdhg {

__global__ void __launch_bounds__(256, 2) pdhg1d_cta_kernel(const MarchParams p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int b = blockIdx.x;
  const int nx = p.nx, K = p.K, N = K * nx, NP = (K + 1) * nx;
  const int egno = p.egno;

  double* phi = sm;                      // [K+1][nx]
  double* phib = phi + NP;               // [K+1][nx]   phi_bar = 2 phi_next - phi_prev
  double* rho = phib + NP;               // [K][nx]
  double* a1 = rho + N;
  double* a2 = a1 + N;
  const int ld = fft_ld(nx);                          // padded FFT row stride (fpad addressing)
  double2* z0 = reinterpret_cast<double2*>(a2 + N);   // [K][ld] complex
  double2* z1 = z0 + (size_t)K * ld;
  double* bak = reinterpret_cast<double*>(z0);         // outer-iteration copies of rho/alp (aliases the FFT
                                                       // workspace, which is idle during the dual sweeps)
  double* red = reinterpret_cast<double*>(z1 + (size_t)K * ld);     // 9*32 doubles of reduction scratch

  double* gphi = p.st_phi + (size_t)b * NP;
  double* grho = p.st_rho + (size_t)b * N;
  double* galp = p.st_alp + (size_t)b * 2 * N;
  const double epsl = p.epsl[b];
  double stepsz = p.stepsz[b];
  const double dt = p.dt, dx = p.dx, eps = p.eps;
  const double inv_n = 1.0 / (double)nx;
  const double ct2 = p.Ct_over_dt2;
  const int nt_all = p.nblocks * K + 1;

  int status = ST_OK;
  bool log_overflow = false;     // sticky: more periodic records than max_rec - 1 in some block (rows dropped, march goes on)
  int blocks_done = p.blk_begin;
  long long inner_total = 0;
  // diagnostic cycle counters (instance 0, thread 0): residual, FFT, solve, IFFT, phi update, dual sweeps, decisions
  long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tl = clock64();
  const bool prof = (p.dbg_ns != nullptr) && b == 0 && tid == 0;
#define TCK(slot) do { if (prof) { const long long t_ = clock64(); tc[slot] += t_ - tl; tl = t_; } } while (0)

  for (int blk = p.blk_begin; blk < p.blk_end && status == ST_OK; ++blk) {
    const size_t lb = (size_t)b * p.nblocks + blk;
    double* elog = p.errlog + lb * p.max_rec * kLogCols;
    while (true) {   // step-size fallback loop (utils_pdhg_solver.py:174-187)
      for (int i = tid; i < NP; i += nth) phi[i] = gphi[i];
      for (int i = tid; i < N; i += nth) { rho[i] = grho[i]; a1[i] = galp[i]; a2[i] = galp[N + i]; }
      __syncthreads();
      const double tau = stepsz / 1.5, sigma = stepsz * 1.5;   // utils_pdhg_solver.py:44-46
      const Recip rc(dt, dx, 1.0, sigma);
      const double c_dt = p.c_on_rho * rc.idt;

      double S_row0, S_rho, S_a1, S_a2;
      {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = tid; i < nx; i += nth) s[0] += phi[i] * phi[i];
        for (int i = tid; i < N; i += nth) { s[1] += rho[i] * rho[i]; s[2] += a1[i] * a1[i]; s[3] += a2[i] * a2[i]; }
        block_sum<4>(s, red);
        S_row0 = s[0]; S_rho = s[1]; S_a1 = s[2]; S_a2 = s[3];
      }

      long long it = p.iter_begin;
      int reason = END_MAXITER, nrec = 0;
      bool lognan = false, logfull = false;
      double err1 = 0.0, err2 = 0.0, rmin = 0.0, rmax = 0.0;

      for (; it < p.n_maxiter; ++it) {
        if (it >= p.iter_pause) { reason = END_PAUSED; break; }

        TCK(7);
        // ---- continuity residual rows 1..K -> z0 (update_fns_in_pdhg.py:72-81) ----
        for (int i = tid; i < N; i += nth) {
          const int k = i / nx, x = i - k * nx;
          const int xm = (x == 0) ? nx - 1 : x - 1, xp = (x == nx - 1) ? 0 : x + 1;
          const double* rk = rho + k * nx;
          const double r0 = rk[x], rm = rk[xm], rp = rk[xp];
          const double m1_0 = (r0 + kRhoOffset) * f_plus(-(p.coef_x[x] * a1[i]));
          const double m1_m = (rm + kRhoOffset) * f_plus(-(p.coef_x[xm] * a1[k * nx + xm]));
          const double m2_0 = (r0 + kRhoOffset) * f_minus(-(p.coef_x[x] * a2[i]));
          const double m2_p = (rp + kRhoOffset) * f_minus(-(p.coef_x[xp] * a2[k * nx + xp]));
          const double rnext = (k + 1 < K) ? rho[i + nx] : 0.0;
          double res = (rnext - r0) * rc.idt + epsl * ((rp + rm - 2 * r0) * rc.idx2);
          res -= (m1_0 - m1_m) * rc.idx + (m2_p - m2_0) * rc.idx;
          if (k == K - 1) res += c_dt;
          z0[k * ld + fpad(x)] = make_double2(res, 0.0);
        }
        __syncthreads();
        TCK(0);

        // ---- H1 preconditioner (utils_precond.py:105-140): FFT_x, per-mode t-solve, IFFT_x ----
        double2* zf = fft_rows<8>(z0, z1, p.plan_1d, ld, p.tw_x, K, 1.0);
        TCK(1);
        double2* zo = (zf == z0) ? z1 : z0;
        if (K == 1) {
          for (int m = tid; m < nx; m += nth) {
            const double d = p.diag[m] + ct2;
            const int e = fpad(m);
            zf[e] = make_double2(zf[e].x / d, zf[e].y / d);
          }
        } else if (ct2 == 0.0) {
          for (int i = tid; i < N; i += nth) {
            const int k = i / nx, m = i - k * nx;
            const double d = p.diag[m];
            const int e = k * ld + fpad(m);
            zf[e] = make_double2(zf[e].x / d, zf[e].y / d);
          }
        } else {
          double* tu = reinterpret_cast<double*>(zo);   // [K][nx] modified super-diagonal
          for (int m = tid; m < nx; m += nth) {
            const double dg = p.diag[m];
            double d0 = dg + ((K == 1) ? ct2 : 2.0 * ct2);
            double tprev = -ct2 / d0;
            tu[m] = tprev;
            const int em = fpad(m);
            double2 bp = make_double2(zf[em].x / d0, zf[em].y / d0);
            zf[em] = bp;
            for (int k = 1; k < K; ++k) {
              const double dk = dg + ((k == K - 1) ? ct2 : 2.0 * ct2);
              const double den = dk + ct2 * tprev;                      // d[k] - dl[k]*tu[k-1], dl = -ct2
              tprev = ((k == K - 1) ? 0.0 : -ct2) / den;
              tu[k * nx + m] = tprev;
              double2 v = zf[k * ld + em];
              bp = make_double2((v.x + ct2 * bp.x) / den, (v.y + ct2 * bp.y) / den);
              zf[k * ld + em] = bp;
            }
            double2 xs = bp;
            for (int k = K - 2; k >= 0; --k) {
              const double t = tu[k * nx + m];
              double2 v = zf[k * ld + em];
              xs = make_double2(v.x - t * xs.x, v.y - t * xs.y);
              zf[k * ld + em] = xs;
            }
          }
        }
        __syncthreads();
        TCK(2);
        double2* zu = fft_rows<8>(zf, zo, p.plan_1d, ld, p.tw_x, K, -1.0);
        TCK(3);

        // ---- phi_next = phi_prev + tau*u ; phi_bar = 2 phi_next - phi_prev (utils_pdhg_solver.py:53-55) ----
        int nanflag = 0;
        double sp[2] = {0.0, 0.0};
        for (int i = tid; i < N; i += nth) {
          const double pp = phi[nx + i];
          const int kk = i / nx;
          const double pn = pp + tau * (zu[kk * ld + fpad(i - kk * nx)].x * inv_n);
          const double df = pn - pp;
          sp[0] += df * df;
          sp[1] += pp * pp;
          nanflag |= is_nan(pn) ? 1 : 0;
          phi[nx + i] = pn;
          phib[nx + i] = 2 * pn - pp;
        }
        if (it == p.iter_begin) for (int i = tid; i < nx; i += nth) phib[i] = phi[i];
        block_sum<2>(sp, red);     // (its barriers also order the phib writes before the sweeps below)
        err1 = sqrt(sp[0]) / sqrt(S_row0 + sp[1]);
        TCK(4);

        // ---- dual sweeps (update_fns_in_pdhg.py:150-180) ----
        double v[9];
        int j = 0;
        for (; j < p.rho_alp_iters; ++j) {
#pragma unroll
          for (int q = 0; q < 9; ++q) v[q] = 0.0;
          for (int i = tid; i < N; i += nth) {
            const int k = i / nx, x = i - k * nx;
            const int xm = (x == 0) ? nx - 1 : x - 1, xp = (x == nx - 1) ? 0 : x + 1;
            const double* pb1 = phib + (k + 1) * nx;
            const double c0 = pb1[x], cm = pb1[xm], cp = pb1[xp];
            const double dxr = (cp - c0) * rc.idx, dxl = (c0 - cm) * rc.idx;
            const double ro = rho[i], a1o = a1[i], a2o = a2[i];
            const double cf = p.coef_x[x];
            const double pinv = (ro + kRhoOffset) * rc.isig;
            const double rinv = prox_rinv(egno, pinv);
            const double a1n = prox_alp(egno, a1o, dxr, pinv, rinv, cf, true);
            const double a2n = prox_alp(egno, a2o, dxl, pinv, rinv, cf, false);
            const double f1 = f_plus(-(cf * a1n)), f2 = f_minus(-(cf * a2n));
            double vec = (c0 - phib[k * nx + x]) * rc.idt - epsl * ((cp + cm - 2 * c0) * rc.idx2);
            vec -= dxr * f1 + dxl * f2;
            vec -= lagr(egno, a1n) + lagr(egno, a2n);
            const double rn = relu_nan(ro + sigma * vec);
            rho[i] = rn; a1[i] = a1n; a2[i] = a2n;
            double d;
            d = rn - ro;   v[0] += d * d; v[1] += rn * rn;
            d = a1n - a1o; v[2] += d * d; v[3] += a1n * a1n;
            d = a2n - a2o; v[4] += d * d; v[5] += a2n * a2n;
            if (j == 0) {
              bak[i] = ro; bak[N + i] = a1o; bak[2 * N + i] = a2o;
            } else {
              d = rn - bak[i];          v[6] += d * d;
              d = a1n - bak[N + i];     v[7] += d * d;
              d = a2n - bak[2 * N + i]; v[8] += d * d;
            }
            nanflag |= is_nan(rn) ? 2 : 0;
          }
          block_sum<9>(v, red);
          const double err = v[0] / v[1] + v[2] / v[3] + v[4] / v[5];
          if (err < eps) { ++j; break; }
        }
        inner_total += j;
        TCK(5);
        const bool multi = (j > 1);
        const double d_rho = multi ? v[6] : v[0], d_a1 = multi ? v[7] : v[2], d_a2 = multi ? v[8] : v[4];

        // ---- outer errors and exits (utils_pdhg_solver.py:59-80) ----
        err2 = sqrt(d_rho) / sqrt(S_rho);
        {
          const double na = sqrt(S_a1), ne = sqrt(d_a1);
          if (na < 1e-6 && ne > 1e-6) err2 += ne; else if (na >= 1e-6) err2 += ne / na;
        }
        {
          const double na = sqrt(S_a2), ne = sqrt(d_a2);
          if (na < 1e-6 && ne > 1e-6) err2 += ne; else if (na >= 1e-6) err2 += ne / na;
        }
        S_rho = v[1]; S_a1 = v[3]; S_a2 = v[5];
        nanflag = __syncthreads_or(nanflag);
        TCK(6);
        if (err1 < eps && err2 < eps) { reason = END_CONVERGED; break; }
        if (nanflag) { reason = END_NAN; break; }
        if (p.print_freq > 0 && it % p.print_freq == 0) {
          double mn = 1e300, mx = -1e300;
          for (int i = tid; i < N; i += nth) { mn = fmin(mn, rho[i]); mx = fmax(mx, rho[i]); }
          double mm[2] = {0.0, 0.0};
          // min/max through the sum reducer would be wrong; use shuffles + scratch directly
          for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          }
          __syncthreads();
          if ((tid & 31) == 0) { red[tid >> 5] = mn; red[32 + (tid >> 5)] = mx; }
          __syncthreads();
          mm[0] = red[0]; mm[1] = red[32];
          for (int w = 1; w < ((nth + 31) >> 5); ++w) { mm[0] = fmin(mm[0], red[w]); mm[1] = fmax(mm[1], red[32 + w]); }
          rmin = mm[0]; rmax = mm[1];
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
      const long long iters_done = (reason == END_MAXITER) ? p.n_maxiter : ((reason == END_PAUSED) ? it : it + 1);
      if (reason != END_PAUSED) {
        // final record (utils_pdhg_solver.py:91-93)
        if (tid == 0) {
          elog[nrec * kLogCols + 0] = err1; elog[nrec * kLogCols + 1] = err2;
          elog[nrec * kLogCols + 2] = rmin; elog[nrec * kLogCols + 3] = rmax;
        }
        ++nrec;
        lognan = lognan || is_nan(err1) || is_nan(err2);
      }
      if (p.fallback && lognan) {
        if (stepsz > p.stepsz_floor[b]) {      // utils_pdhg_solver.py:181-183
          stepsz -= p.stepsz_delta[b];
          __syncthreads();
          continue;
        }
        status = ST_SOL_NAN;                   // :184-187 "algorithm failed"
        break;
      }
      // ---- accept the block ----
      if (tid == 0) {
        p.iters[lb] = iters_done;
        p.stepsz_used[lb] = stepsz;
        p.nrec[lb] = nrec;
        p.end_reason[lb] = reason;
      }
      log_overflow = log_overflow || logfull;
      if (reason == END_PAUSED) status = ST_PAUSED;
      __syncthreads();
      if (p.handoff && reason != END_PAUSED) {
        // outputs (utils_pdhg_solver.py:193-199): all rows but the last, except for the final block
        double* pa = p.phi_all + ((size_t)b * nt_all + (size_t)blk * K) * nx;
        const int nrows = (blk == p.nblocks - 1) ? K + 1 : K;
        for (int i = tid; i < nrows * nx; i += nth) pa[i] = phi[i];
        double* ra = p.rho_all + ((size_t)b * (nt_all - 1) + (size_t)blk * K) * nx;
        double* aa1 = p.alp_all + (((size_t)b * 2 + 0) * (nt_all - 1) + (size_t)blk * K) * nx;
        double* aa2 = p.alp_all + (((size_t)b * 2 + 1) * (nt_all - 1) + (size_t)blk * K) * nx;
        for (int i = tid; i < N; i += nth) { ra[i] = rho[i]; aa1[i] = a1[i]; aa2[i] = a2[i]; }
        // warm start of the next block (:200-205): phi0 <- phi0 + (phi_curr[-1] - phi0[0])
        for (int i = tid; i < NP; i += nth) {
          const int x = i % nx;
          gphi[i] = gphi[i] + (phi[K * nx + x] - phi[x]);   // phi row 0 == phi0 row 0 (never updated)
        }
      } else {
        for (int i = tid; i < NP; i += nth) gphi[i] = phi[i];
      }
      for (int i = tid; i < N; i += nth) { grho[i] = rho[i]; galp[i] = a1[i]; galp[N + i] = a2[i]; }
      __syncthreads();
      if (reason != END_PAUSED) blocks_done = blk + 1;
      break;
    }
  }
  if (prof) for (int q = 0; q < 8; ++q) p.dbg_ns[q] = (double)tc[q];
#undef TCK
  if (tid == 0) {
    p.status[b] = (status == ST_OK && log_overflow) ? ST_LOG_OVERFLOW : status;   // overflow never stops the march (only NaN failure and pause do)
    p.blocks_done[b] = blocks_done;
    p.stepsz[b] = stepsz;
    p.inner_total[b] = inner_total;
  }
}

size_t pdhg1d_cta_smem_bytes(int nx, int K) {
  // phi, phi_bar [(K+1) nx] ; rho, alp1, alp2 [K nx] ; two complex FFT buffers [K][fft_ld(nx)] ; reduction scratch
  return ((size_t)(5 * K + 2) * nx + (size_t)4 * K * fft_ld(nx) + 9 * 32) * sizeof(double);
}

int pdhg1d_cta_threads(int nx, int K) {
  const int N = nx * K;
  const int rounds = (N + 255) / 256;
  int nt = ((N + rounds - 1) / rounds + 31) / 32 * 32;
  if (nt < 64) nt = 64;
  if (nt > 256) nt = 256;
  return nt;
}

cudaError_t launch_pdhg1d_cta(const MarchParams& p, int B, cudaStream_t stream) {
  const size_t smem = pdhg1d_cta_smem_bytes(p.nx, p.K);
  cudaError_t e = cudaFuncSetAttribute(pdhg1d_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  pdhg1d_cta_kernel<<<B, pdhg1d_cta_threads(p.nx, p.K), smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace pdhg
