// Kernel parameter blocks shared between the C-ABI layer and the kernels.
#pragma once
#include <stdint.h>
#include "pdhg_device.cuh"

namespace pdhg {

// Status of one instance after a launch.
enum : int { ST_OK = 0, ST_SOL_NAN = 1, ST_PAUSED = 3, ST_LOG_OVERFLOW = 4 };

// How a block solve ended (per time block).
enum : int { END_CONVERGED = 0, END_NAN = 1, END_MAXITER = 2, END_PAUSED = 3 };

constexpr int kLogCols = 4;   // err1, err2, min rho_next, max rho_next

// Diagnostic knobs, read from the environment ONCE per handle (pdhg_create) — never on the launch path.  None is needed in
// normal use; the parity tests use them to force every kernel variant onto small grids.
struct Knobs {
  int no_w256, force_w256;   // PDHG_NO_W256 / PDHG_FORCE_W256: tiled vs warp-private 256-point transforms
  int dfuse;                 // PDHG_DFUSE=n: max. dual sweeps fused per pass (0 = automatic)
  int tma;                   // PDHG_TMA=0|1: TMA row pipelines off / forced wherever they are supported (-1 = automatic)
  int no_bslab;              // PDHG_NO_BSLAB: phase B in three grid-wide passes instead of one pass per ky-slab
  int no_k1;                 // PDHG_NO_K1: 1-D K = 1 problems on the generic single-CTA kernel
  int profile;               // PDHG_PROFILE: sub-step cycle counters of the single-CTA kernel
};

// Problem + marching description for the time-block solver kernels.  One "instance" = one independent
// HJ problem; instance b uses g/epsl/stepsz index b.  All pointers are device pointers.
struct MarchParams {
  // grid
  int ndim, egno, nx, ny, K;       // K = time_step_per_PDHG - 1
  int bc_x, bc_y;
  double dt, dx, dy, c_on_rho;
  double eps;                      // stopping tolerance, reused as the inner tolerance (utils_pdhg_solver.py:52,57)
  int rho_alp_iters;               // 10 (update_fns_in_pdhg.py:168)
  long long n_maxiter;             // absolute iteration bound of a block solve
  long long iter_begin;            // first iteration index of this launch (resume support; 0 normally)
  long long iter_pause;            // stop with END_PAUSED when i reaches this (>= n_maxiter: never)
  int print_freq;
  // marching
  int nblocks;                     // nt_PDHG
  int blk_begin, blk_end;          // time blocks processed by this launch
  int handoff;                     // 1: PDHG_multi_step semantics (outputs to *_all, warm start, fallback)
  int fallback;                    // 1: NaN -> stepsz -= delta and retry (utils_pdhg_solver.py:180-187)
  int max_rec;                     // rows of err log per block
  // per-instance inputs
  const double* epsl;              // [B]
  double* stepsz;                  // [B] in: step size to start with; out: step size after the launch
  const double* stepsz_delta;      // [B] initial stepsz/10
  const double* stepsz_floor;      // [B] stepsz_min + delta
  // tables (shared by all instances)
  const double* coef_x;            // [nx]  a(x) (egno 1,2) / x (egno 3)
  const double* coef_y;            // [ny]
  const double* diag;              // [nx*nyh] per-mode (lambda+C)^pow (1-D) or lambda+C (2-D)
  const double* green;             // [green_R + 1] taps of G = IFFT(1/d) for the K = 1 single-CTA kernel (1-D), else null
  int green_R;                     // -1: not available
  const double* dct_cos;           // [4 nx] cos(pi m/(2 nx)) for the Neumann-x DCT-II (bc_x = 1), else null
  const double2* tw_x;             // [nx] master twiddles exp(-2 pi i m/nx)
  const double2* tw_y;             // [ny]
  double Ct_over_dt2;              // Ct/dt^2 (0 => plain divide, utils_precond.py:134)
  FftPlan plan_x, plan_y;
  FftPlan plan_1d;                 // x plan restricted to radix <= 8 (single-CTA kernel: 128-register budget)
  // state in/out  (planar, active control components only): phi [B][K+1][n], rho [B][K][n], alp [B][A][K][n]
  double* st_phi;
  double* st_rho;
  double* st_alp;
  // marching outputs (handoff=1): phi_all [B][nt][n], rho_all [B][nt-1][n], alp_all [B][A][nt-1][n]
  double* phi_all;
  double* rho_all;
  double* alp_all;
  // logs
  long long* iters;                // [B][nblocks]
  double* stepsz_used;             // [B][nblocks]
  int* nrec;                       // [B][nblocks]
  double* errlog;                  // [B][nblocks][max_rec][kLogCols]
  int* end_reason;                 // [B][nblocks]
  int* status;                     // [B]
  int* blocks_done;                // [B]
  long long* inner_total;          // [B] total dual sweeps executed (statistics)
  Knobs knobs;
  double* dbg_ns;                  // [16] diagnostic: clock cycles per sub-step of instance 0 (single-CTA kernel); may be null
};

// Slab mode (multi-GPU x-slab decomposition): one phase of the cooperative kernel on CALLER-owned buffers of a local,
// ghost-padded slab (see pdhg_ext_phase in include/pdhg_b200.h).  Null pointers keep the handle's own workspace.
struct ExtPhaseDesc {
  double *phi_in, *phi_out, *phib, *rho_in, *alp_in, *rho_out, *alp_out;
  void* zt;
  double* sums;
  int sum_lo, sum_hi, nyh_override, ky_off, nyh_tab;
  // fused transpose (pdhg_ext_set_exchange): phase 0 scatters the half spectrum straight into the ky-slabs of the P ranks, phase 1 its
  // result straight back into their x-slabs (peer-mapped pointers), instead of writing `zt` for a separate exchange step
  int xch_P = 0, xch_rank = 0, xch_nxl = 0, xch_kyl = 0, xch_nyh = 0, xch_pull = 0;
  void* xch_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace pdhg
