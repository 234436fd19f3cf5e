/* pdhg_b200.h — C ABI of the B200-native PDHG hot path.
 *
 * Drop-in boundary for the PDHG iteration of TingweiMeng/PDHG-optimal-control (reference paths are
 * relative to /root/reference).  The reference has no FFI of its own (it is pure Python/JAX); the
 * injection point this library replaces is the pair of callables + the solver loop they are passed to:
 *
 *   fn_update_primal / fn_update_dual        jaxsrc/run_example.py:193-203
 *   PDHG_solver_oneiter(...)                 jaxsrc/utils/utils_pdhg_solver.py:9-94
 *   PDHG_multi_step(...)                     jaxsrc/utils/utils_pdhg_solver.py:97-225
 *
 * The Python host package `pdhg_b200` binds these entry points with ctypes and re-exports the reference's
 * names (`solve_HJ`, `PDHG_multi_step`, `PDHG_solver_oneiter`, `update_primal_1d/2d`,
 * `update_dual_alternative`), see INTEGRATION.md.
 *
 * Conventions
 *   - all arrays are C-contiguous fp64, t slowest, then x, then y, then control component — exactly the
 *     reference's layouts with a leading batch axis B of independent problem instances:
 *         phi [B][K+1][nx][ny]   rho [B][K][nx][ny]   alp [B][2*ndim][K][nx][ny][n_ctrl]
 *         phi_all [B][nt][nx][ny]  rho_all [B][nt-1][nx][ny]  alp_all [B][2*ndim][nt-1][nx][ny][n_ctrl]
 *     (ny = 1 in 1-D; K = time_step_per_PDHG - 1; nt = nblocks*K + 1);
 *   - "dev" pointers are device pointers on the handle's GPU, "host" pointers are ordinary host memory;
 *   - the caller owns every buffer; inputs are never modified; the handle owns its workspaces;
 *   - every entry returns 0 on success or a negative pdhg_status; pdhg_last_error() gives the text.
 *     No C++ exception crosses this boundary.  There is NO CPU fallback: without a CUDA device every
 *     compute entry fails with PDHG_ERR_CUDA.
 *   - a handle is bound to one device and is not thread-safe; `stream` is a cudaStream_t (NULL = default).
 */
#ifndef PDHG_B200_H_
#define PDHG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pdhg_handle pdhg_handle;

enum pdhg_status {
  PDHG_OK = 0,
  PDHG_ERR_ARG = -1,        /* invalid argument / unsupported configuration */
  PDHG_ERR_CUDA = -2,       /* CUDA runtime error (incl. no device) */
  PDHG_ERR_NOMEM = -3,
  PDHG_ERR_UNSUPPORTED = -4 /* valid reference configuration this build does not cover */
};

/* per-instance result of a solve (pdhg_logs.status) */
enum pdhg_inst_status {
  PDHG_INST_OK = 0,
  PDHG_INST_SOL_NAN = 1,      /* step-size fallback exhausted: "algorithm failed" (utils_pdhg_solver.py:184-187) */
  PDHG_INST_PAUSED = 3,       /* stopped at iter_pause; resume with iter_begin = iters */
  PDHG_INST_LOG_OVERFLOW = 4  /* more than max_rec-1 periodic records in some block: the extra rows were dropped, the march itself ran to its end */
};

/* how a time block ended (pdhg_logs.end_reason) */
enum pdhg_end_reason { PDHG_END_CONVERGED = 0, PDHG_END_NAN = 1, PDHG_END_MAXITER = 2, PDHG_END_PAUSED = 3 };

/* Static description of a problem family; replaces the closure state of run_example.py:157-203 and the
 * `fns_dict` namedtuple of set_fns.py:164-166 (the callables cannot cross a C ABI: `egno` selects the
 * compiled-in problem, `coef_x/coef_y` carry its coefficient tables). */
typedef struct pdhg_config {
  int32_t ndim;            /* 1 | 2 */
  int32_t egno;            /* 1 quadratic L, 2 indicator L, 3 Newton (2-D only)      set_fns.py:52-166 */
  int32_t nx, ny;          /* ny = 1 in 1-D */
  int32_t K;               /* time_step_per_PDHG - 1 */
  int32_t n_ctrl;          /* 1 (1-D, egno 3) | 2 (2-D egno 1,2)                     run_example.py:228-240 */
  int32_t bc_x, bc_y;      /* 0 periodic, 1 Neumann                                   utils_diff_op.py:5-7 */
  double dt, dx, dy;       /* run_example.py:160-162 */
  double c_on_rho;         /* FLAGS.c_on_rho */
  double C, pow, Ct;       /* preconditioner constants (pow, Ct ignored in 2-D)       utils_precond.py:105,142 */
  double eps;              /* stopping tolerance, also the inner tolerance            utils_pdhg_solver.py:52,57 */
  int32_t rho_alp_iters;   /* 10                                                      update_fns_in_pdhg.py:168 */
  int32_t batch;           /* B: independent instances solved per call */
  int32_t nblocks;         /* nt_PDHG = (nt-1)/K time blocks (1 for a single block solve) */
  int32_t max_rec;         /* rows reserved per block in the error log (>= 2) */
  int32_t device;          /* CUDA device ordinal */
  int32_t path;            /* 0 auto | 1 single-CTA shared-memory kernel | 2 cooperative multi-CTA kernel |
                            * 3 tables only (ndim 2, K 1, batch 1): per-mode table, twiddles and plans of the grid, NO state and NO
                            *   workspace - the handle only serves pdhg_ext_phase phase 1 on a caller-owned spectrum (the x-transform
                            *   handle of the slab decomposition: its memory must not grow with the full grid) */
} pdhg_config;

/* Host-side logs of a solve; any pointer may be NULL.  Shapes: [B][nblocks] unless noted. */
typedef struct pdhg_logs {
  int64_t* iters;          /* pdhg_iters of each accepted block (utils_pdhg_solver.py:91,189) */
  double* stepsz_used;     /* stepsz_param each block was accepted with */
  int32_t* nrec;           /* rows of errlog written for the block */
  double* errlog;          /* [B][nblocks][max_rec][4]: err1, err2, min rho, max rho (= error_all rows) */
  int32_t* end_reason;     /* pdhg_end_reason */
  int32_t* status;         /* [B] pdhg_inst_status */
  int32_t* blocks_done;    /* [B] number of accepted blocks */
  double* stepsz_final;    /* [B] step size after the last fallback */
  int64_t* inner_total;    /* [B] dual sweeps executed (statistics, not in the reference) */
} pdhg_logs;

/* coef_x[nx], coef_y[ny] (host): a(x)=(x-1)^2+0.1 per direction for egno 1,2 (set_fns.py:117-118,145);
 * for egno 3 coef_x = the x grid itself (f_y = x, set_fns.py:98).  coef_y may be NULL in 1-D. */
int pdhg_create(const pdhg_config* cfg, const double* coef_x, const double* coef_y, pdhg_handle** out);
void pdhg_destroy(pdhg_handle* h);
const char* pdhg_last_error(void);
/* 1 = single-CTA kernel, 2 = cooperative kernel (what `path = 0` resolved to) */
int pdhg_path(const pdhg_handle* h);
/* device time (ms, CUDA events on the launch stream) of the solver kernel(s) of the last pdhg_solve_block /
 * pdhg_multi_step call, excluding the staging copies around them; -1 if none.  Call after the stream is idle. */
double pdhg_last_kernel_ms(const pdhg_handle* h);
/* diagnostic: device nanoseconds the last cooperative-kernel march spent in each phase, out16[0..5] = {A residual+FFT_y,
 * B FFT_x+t-solve, C IFFT_y+phi update, D dual sweeps+reduction, -, setup/records/output}, out16[6..15] = sub-steps seen
 * by CTA 0 (A: compute, fft, store; B: pass 1, 2, 3; C: load, fft, update; spare); zeros on the single-CTA path */
int pdhg_phase_times(pdhg_handle* h, double* out16);
/* Slab mode — building block of the multi-GPU x-slab decomposition of ONE large 2-D grid (BASELINE configs[4]; the host
 * side with the NCCL halo exchange / all-to-all transposes / all-reduce is pdhg_b200/slab.py).  Runs ONE phase of the
 * cooperative kernel on caller-owned device buffers of a local, ghost-padded slab (null = the handle's workspace):
 *   phase 0 = A: (rho_in, alp_in) -> zt        continuity residual + y-FFT, half spectrum stored transposed [K][ny/2+1][nx]
 *   phase 1 = B: zt in place                   x-FFT, per-mode solve, inverse x-FFT; `nyh_override`/`ky_off`/`nyh_tab` select the
 *                                              ky-slab this rank owns after the transpose (table row length nyh_tab)
 *   phase 2 = C: (zt, phi_in) -> phi_out, phib inverse y-FFT + phi update (step = tau)
 *   phase 3 = D: (phib, rho_in, alp_in) -> rho_out, alp_out   one dual sweep (step = sigma), may run in place; `pass_mask` = 2 fuses two
 *                                              consecutive sweeps into the pass (not in place; the second sweep's sums go to slots 20..35)
 *   phase 4 = E: outer differences of (rho_out, alp_out) against (rho_in, alp_in)
 * Rows outside [sum_lo, sum_hi) (ghost rows) do not contribute to the error sums.  Phases 3 and 4 store the grid totals of the
 * reduced quantities (dual sums in 0..15, the preceding phase C's primal sums in 16..18, a fused second sweep in 20..35) to
 * bufs->sums (device, room for 20 + 16 (pass_mask - 1) doubles, 84 at most). */
typedef struct pdhg_ext_buffers {
  double *phi_in, *phi_out, *phib, *rho_in, *alp_in, *rho_out, *alp_out;
  void* zt;
  double* sums;
} pdhg_ext_buffers;
/* largest `pass_mask` (dual sweeps fused into one pass) phase 3 accepts for this handle; the sums of sweep s >= 2 of a pass land in
 * slots 20 + 16 (s - 2) .. of bufs->sums, which then needs room for 20 + 16 (pass_mask - 1) doubles */
int pdhg_max_fuse(pdhg_handle* h);
int pdhg_ext_phase(pdhg_handle* h, int phase, int pass_mask, double step, double epsl, const pdhg_ext_buffers* bufs,
                   int sum_lo, int sum_hi, int nyh_override, int ky_off, int nyh_tab, void* stream);
/* Slab mode, fused transposes (compute + exchange in ONE kernel over peer memory).  The half spectrum of the grid is a distributed
 * array [nyh][P nxl]: rank d holds columns [d nxl, (d + 1) nxl) of all rows as columns 1..nxl of its [nyh][nxl + 2] array `zt`.
 *  - phase 1 through a handle of the global x-extent, on this rank's ky-slab (nyh_override rows starting at ky_off), after
 *    pdhg_ext_set_exchange(h, P, rank, nxl, kyl, nyh, gather, ptrs) with ptrs[d] = rank d's `zt`: the result of (slab row t, column x) is
 *    stored straight into its owner's array, ptrs[x / nxl][ky_off + t][1 + x mod nxl], instead of bufs->zt; with gather != 0 the input
 *    rows are also loaded straight from there (bufs->zt is then not touched at all: all-gather -> transform / solve -> scatter in one
 *    kernel, the rank reads and writes only the rows of its own ky range);
 *  - phase 0 through the rank's local handle (nx = nxl + 2 rows incl. one ghost row per side) after pdhg_ext_set_exchange with
 *    ptrs[d] = rank d's ky-slab buffer ([kyl][P nxl] complex): the value of (ky, local row i in 1..nxl) is stored straight into
 *    ptrs[ky / kyl][ky mod kyl][rank nxl + i - 1] instead of bufs->zt (32-96-byte pieces: slower than a copy over NVLink; for
 *    emulation and measurement).
 * ptrs[d] are device pointers valid on THIS device: the rank's own buffer, peer-mapped buffers of the other ranks (CUDA VMM /
 * symmetric memory), or plain buffers when several ranks are emulated on one GPU.  The caller still separates the phases with a
 * barrier across the ranks.  P = 0 switches the exchange off again.  pdhg_ext_exchange_ok: 1 if the handle's transforms can do this
 * (generic variants, K = 1, periodic). */
int pdhg_ext_exchange_ok(pdhg_handle* h);
int pdhg_ext_set_exchange(pdhg_handle* h, int P, int rank, int nxl, int kyl, int nyh, int gather, void* const* peer_ptrs);

/* diagnostic (profiling aid): launches ONE phase of the cooperative kernel `reps` times on the workspace left by the last
 * march, so that ncu sees each phase as its own launch: phase 0 = A, 1 = B (pass_mask bit 0/1/2 = x-FFT / t-solve / inverse
 * x-FFT), 2 = C, 3 = D + reduction; `step` = tau or sigma.  Results are meaningless; caller buffers are not touched. */
int pdhg_debug_phase(pdhg_handle* h, int phase, int pass_mask, double step, int reps);
/* number of kernel launches issued through this handle so far */
int64_t pdhg_launch_count(const pdhg_handle* h);

/* fn_update_primal (run_example.py:193-195,199-201 -> update_fns_in_pdhg.py:135-147), batch = 1 per call slot b.
 * dev buffers: phi_prev/phi_next [B][K+1][n], rho_prev [B][K][n], alp_prev [B][2 ndim][K][n][n_ctrl]. */
int pdhg_update_primal(pdhg_handle* h, const double* phi_prev_dev, const double* rho_prev_dev,
                       const double* alp_prev_dev, const double* epsl_host, double tau,
                       double* phi_next_dev, void* stream);

/* fn_update_dual (run_example.py:196-197,202-203 -> update_fns_in_pdhg.py:167-180); n_inner_host[B] receives
 * the number of sweeps executed and err_host[B] the `err` of the last sweep (update_dual_oneiter's third
 * return value, update_fns_in_pdhg.py:162-165); either may be NULL.  Sweeps use cfg.rho_alp_iters. */
int pdhg_update_dual(pdhg_handle* h, const double* phi_bar_dev, const double* rho_prev_dev,
                     const double* alp_prev_dev, const double* epsl_host, double sigma, double eps,
                     double* rho_next_dev, double* alp_next_dev, int32_t* n_inner_host, double* err_host,
                     void* stream);

/* PDHG_solver_oneiter (utils_pdhg_solver.py:9-94) for B instances: iterations [iter_begin, ...) of ONE time
 * block from (phi0, rho0, alp0) with step size stepsz_host[b], until convergence / NaN / n_maxiter /
 * iter_pause.  Final iterate -> *_out (may alias nothing).  logs use nblocks = 1. Synchronises the stream. */
int pdhg_solve_block(pdhg_handle* h, const double* phi0_dev, const double* rho0_dev, const double* alp0_dev,
                     const double* epsl_host, const double* stepsz_host, int64_t n_maxiter, int64_t iter_begin,
                     int64_t iter_pause, int32_t print_freq, double* phi_out_dev, double* rho_out_dev,
                     double* alp_out_dev, pdhg_logs* logs, void* stream);

/* PDHG_multi_step (utils_pdhg_solver.py:97-225) for B instances: nblocks sequential time blocks from the
 * initial data g [B][n] (phi0 = tile(g), rho0 = c_on_rho, alp0 = 0), with warm starts and the NaN step-size
 * fallback.  stepsz_host[b] = initial stepsz_param.  Synchronises the stream. */
int pdhg_multi_step(pdhg_handle* h, const double* g_dev, const double* epsl_host, const double* stepsz_host,
                    int64_t n_maxiter, int32_t print_freq, double* phi_all_dev, double* rho_all_dev,
                    double* alp_all_dev, pdhg_logs* logs, void* stream);

/* pdhg_multi_step restricted to the time blocks [blk_begin, blk_end) of the march (utils_pdhg_solver.py:165-213, one pass of
 * the `for i in range(init_t_ind, nt_PDHG)` loop per block): what the host needs to write the reference's `save_middle` file
 * after EVERY block (:211-212) and to resume a march (`load_middle`, :139-154 — unwired and broken in the reference).
 * blk_begin = 0 starts from g like pdhg_multi_step; blk_begin > 0 continues from the handle's marching state — the next
 * block's warm-started phi0 and (rho0, alp0), left there by the previous range call or installed with pdhg_set_march_state —
 * with the CURRENT step size stepsz_cur_host[b] (NULL: the initial one).  The fallback decrement and floor always derive from
 * stepsz_host[b], the initial stepsz_param (:160-161).  Outputs land in rows [blk_begin K, blk_end K) of the *_all arrays; logs
 * are [B][nblocks] as in pdhg_multi_step (entries outside the range keep earlier values). */
int pdhg_multi_step_range(pdhg_handle* h, const double* g_dev, const double* epsl_host, const double* stepsz_host,
                          const double* stepsz_cur_host, int64_t n_maxiter, int32_t print_freq, int32_t blk_begin,
                          int32_t blk_end, double* phi_all_dev, double* rho_all_dev, double* alp_all_dev, pdhg_logs* logs,
                          void* stream);
/* marching state between two time blocks (device buffers, reference layouts): phi0 [B][K+1][n] of the NEXT block (already
 * shifted by phi_curr[-1] - phi0[0], utils_pdhg_solver.py:200-203), rho0 [B][K][n], alp0 [B][2 ndim][K][n][n_ctrl] */
int pdhg_get_march_state(pdhg_handle* h, double* phi0_dev, double* rho0_dev, double* alp0_dev, void* stream);
int pdhg_set_march_state(pdhg_handle* h, const double* phi0_dev, const double* rho0_dev, const double* alp0_dev, void* stream);

/* Same as pdhg_multi_step with HOST buffers for g and the three outputs (H2D / D2H copies inside). */
int pdhg_multi_step_host(pdhg_handle* h, const double* g_host, const double* epsl_host, const double* stepsz_host,
                         int64_t n_maxiter, int32_t print_freq, double* phi_all_host, double* rho_all_host,
                         double* alp_all_host, pdhg_logs* logs);

/* Closed-loop trajectories under the computed feedback control (run_example.py:18-155, compute_traj_1d / compute_traj_2d):
 * Euler-Maruyama  x_{k+1} = x_k + f(alp(x_k, t_k), x_k) dt_k + sqrt(2 epsl dt_k) z_k  for n_sample independent initial points, one
 * launch, one thread per sample.  All pointers are device pointers:
 *   alp      1-D: [2][nt-1][nx] (the reference passes alp[..., 0]);  2-D: [4][nt-1][nx][ny][n_ctrl] (the solver's returned layout,
 *            time already reversed by the caller as in run_example.py:356)
 *   x_nodes  [nx] ascending grid of the x axis (1-D: already reduced mod x_period and sorted, as numpy.interp does), y_nodes [ny]
 *   t_arr    [nt];  noise [nt-1][n_sample][ndim] standard normal draws (required iff epsl > 0; host-seeded, numpy.random order)
 *   x_init   [n_sample] (1-D) or [n_sample][2];  traj_x [nt][n_sample](x2),  traj_alp [nt-1][n_sample][n_ctrl]
 * nearest = 0: linear interpolation (numpy.interp with period / scipy interpn on the periodically extended or edge-clamped
 * grid), 1: nearest grid value (egno 2, run_example.py:349-350).  bc 0 periodic, 1 edge-clamped (Neumann, egno 3 x axis). */
int pdhg_compute_traj(int32_t ndim, int32_t egno, int32_t n_ctrl, int32_t nx, int32_t ny, int32_t nt, int32_t n_sample, int32_t bc_x,
                      int32_t bc_y, int32_t nearest, double x_period, double y_period, double epsl, const double* alp_dev,
                      const double* x_nodes_dev, const double* y_nodes_dev, const double* t_arr_dev, const double* noise_dev,
                      const double* x_init_dev, double* traj_x_dev, double* traj_alp_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDHG_B200_H_ */
