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
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "pdhg_params.h"

namespace cg = cooperative_groups;

namespace pdhg {

constexpr int kFuseMax = 5;    // dual sweeps fused into one pass over memory (at most)
constexpr int kNQ = 20 + 16 * (kFuseMax - 1);   // reduced quantities per epoch: 0..15 dual sweep, 16..18 primal, 20 + 16 (s - 1) + (0..15): sweep s of a fused pass
constexpr int kNV = 20;        // the first kNV totals are handed to every thread in registers, the rest is read from shared memory on demand
#ifndef PDHG_COOP_THREADS
#define PDHG_COOP_THREADS 512
#endif
#ifndef PDHG_COOP_CTAS_PER_SM
#define PDHG_COOP_CTAS_PER_SM 1
#endif
constexpr int kThreads = PDHG_COOP_THREADS;
constexpr int kCtasPerSm = PDHG_COOP_CTAS_PER_SM;
constexpr int kWarps = kThreads / 32;

// diagnostic sub-phase timers share slots 0..2 between phase A and the single-pass phase B: -DPDHG_TICKS_B shows B's
#ifdef PDHG_TICKS_B
#define PDHG_TICK_A(s)
#define PDHG_TICK_B(s) c.tick(s)
#else
#define PDHG_TICK_A(s) c.tick(s)
#define PDHG_TICK_B(s)
#endif

struct CoopWs {
  double* phi[2];     // ping-pong phi [(K+1) n]
  double* phib;       // phi_bar [(K+1) n]
  double* rho[3];     // rho [K n]: current outer iterate + two work buffers of the inner dual loop
  double* alp[3];     // alp [A][K n], same roles
  double2* zt;        // [K][nyh][nx] half spectrum, transposed
  double* partials;   // [2][grid][kNQ]
  double* den;        // [K][nyh][nx]  reciprocal Thomas pivots (K > 1)
  double* tu;         // [K][nyh][nx]  modified super-diag  (K > 1)
  double* phase_ns;   // [8] accumulated device time per phase of the last march (A, B, C, D+reduce, records, setup/output)
};

enum : int { MODE_MARCH = 0, MODE_PRIMAL = 1, MODE_DUAL = 2, MODE_TABLES = 3, MODE_PHASE = 4 };

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
  int sum_lo, sum_hi;        // x-rows that contribute to the error sums (slab mode: ghost rows excluded); default [0, nxe)
  int ky_off, nyh_tab;       // phase B on an exchanged ky-slab: offset and row length of the per-mode table; default 0, nyh
  double* ext_sums;          // MODE_PHASE: device array [kNV = 20] receiving the grid totals of phases D / E (may be null)
  // MODE_PHASE, slab mode with fused transposes: xch_on = 1 (phase A): the spectrum value of (ky, local x-row i) goes to rank d = ky / kyl,
  // slab row ky - d kyl, column rank * nxl + i - 1 of its [kyl][nx_global] ky-slab; xch_on = 2 (phase B on a ky-slab): the value of
  // (slab row t, global column x) goes to rank d = x / nxl, row ky_off + t, column 1 + x - d nxl of its [nyh][nxl + 2] spectrum
  int xch_on, xch_rank, xch_nxl, xch_kyl, xch_nyh;
  int xch_ld;                // row length of the destination: P nxl (forward), nxl + 2 (backward)
  int xch_pull;              // xch_on = 2: phase B also GATHERS its input rows from the owners' x-slabs (same pointers)
  double2* xch_ptr[8];
  int dbg_phase, dbg_pass;   // MODE_PHASE (profiling): phase id 0..3 = A,B,C,D and pass mask of phase B (bit 0,1,2)
  int tma_d;          // 1: dual sweep through the TMA row pipeline (phase_D_tma); tma_R rows per tile, tma_S ring stages at most
  int tma_R, tma_S;
  int work_bytes;     // size of the work area (FFT buffers / accumulator slots / TMA ring)
  int fast_y, fast_x; // warp-private 256-point transforms along y (phases A, C) / along x (phase B)
  int b_slab;         // phase B in one pass per ky-slab (phase_B_slab)
  int d_fuse;         // max. inner dual sweeps fused per pass while the inner loop is long (1 = off; PDHG_DFUSE=n overrides)
  double dxe, dye;
  double r_idt, r_idx, r_idy, r_idx2, r_idy2;   // 1/dt, 1/dx, 1/dy, 1/dx^2, 1/dy^2 (IEEE divisions done once on the host)
  int o_twy, o_cx, o_cy, o_work;               // byte offsets of the shared-memory tables and of the work area
  float inv_nx;                                // 1/nxe for the division-free index arithmetic
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

// The argument block of the current launch lives in CONSTANT memory (copied there on the launch stream right before the launch):
// every phase function — they are not inlined — reads launch constants as constant-bank operands (c[3][..] in SASS, usable
// directly by DFMA / IMAD / address arithmetic), so grid constants, reciprocals, sizes and base pointers occupy no registers,
// no shared memory and no local memory in the streaming loops.  (Round 1 kept a copy in shared memory: every use was an LDS
// into a register, and under the 128-register cap ptxas split the dual sweep's load batch in two to make room — two exposed
// memory latencies per item instead of one, profiles/r02_phaseD_stalls.txt.)
__constant__ CoopArgs g_cargs;
// All shared memory is ONE dynamic region addressed from this symbol, so that every access compiles to LDS/STS (pointers
// stored in a struct would be generic and compile to LD/ST): [reduction scratch][barriers][twx][twy][cx][cy][work].
extern __shared__ __align__(16) unsigned char g_sm[];
constexpr int kArgsBytes = 0;
constexpr int kRedBytes = kNQ * kWarps * 8;
constexpr int kBarBytes = 128;    // mbarriers of the TMA row pipelines: full[8], empty[8]
__device__ __forceinline__ const CoopArgs& cargs() { return g_cargs; }

// TMA row pipelines (phase_D_tma ...): ring state kept by every thread; the mbarriers full[8], empty[8] live in shared memory
// behind the reduction scratch and are initialised once per launch (full: 1 arrival + transaction bytes, empty: one arrival
// per consumer warp).  Bit s of par_full / par_empty = parity of the next phase to wait for on full[s] / empty[s].
struct RowPipe { uint32_t par_full, par_empty; };
constexpr int kPipeMaxStages = 8;

struct Ctx {
  cg::grid_group grid;
  RowPipe pipe;
  int epoch;
  unsigned long long tsub[10], tl;   // diagnostic sub-phase timers (CTA 0, thread 0 only)
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(g_sm + kArgsBytes); }
  __device__ __forceinline__ uint64_t* bars() const { return reinterpret_cast<uint64_t*>(g_sm + kArgsBytes + kRedBytes); }
  __device__ __forceinline__ const double2* twx() const { return reinterpret_cast<const double2*>(g_sm + kArgsBytes + kRedBytes + kBarBytes); }
  __device__ __forceinline__ const double2* twy() const { return reinterpret_cast<const double2*>(g_sm + cargs().o_twy); }
  __device__ __forceinline__ const double* cx() const { return reinterpret_cast<const double*>(g_sm + cargs().o_cx); }
  __device__ __forceinline__ const double* cy() const { return reinterpret_cast<const double*>(g_sm + cargs().o_cy); }
  __device__ __forceinline__ double2* work() const { return reinterpret_cast<double2*>(g_sm + cargs().o_work); }
  __device__ __forceinline__ void tick(int slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      tsub[slot] += t - tl; tl = t;
    }
  }
  __device__ Ctx() : grid(cg::this_grid()), epoch(0) {
    const CoopArgs& a = cargs();
    for (int i = 0; i < 10; ++i) tsub[i] = 0;
    tl = 0;
    const int o_twx = kArgsBytes + kRedBytes + kBarBytes, o_twy = a.o_twy, o_cx = a.o_cx, o_cy = a.o_cy;
    double2* tx = reinterpret_cast<double2*>(g_sm + o_twx);
    double2* ty = reinterpret_cast<double2*>(g_sm + o_twy);
    double* px = reinterpret_cast<double*>(g_sm + o_cx);
    double* py = reinterpret_cast<double*>(g_sm + o_cy);
    for (int i = threadIdx.x; i < a.nxe; i += blockDim.x) { tx[i] = a.tw_xe[i]; px[i] = a.coef_xe[i]; }
    for (int i = threadIdx.x; i < a.nye; i += blockDim.x) { ty[i] = a.tw_ye[i]; py[i] = a.coef_ye[i]; }
    pipe.par_full = 0u; pipe.par_empty = 0u;
    if (threadIdx.x == 0) {
      for (int s = 0; s < kPipeMaxStages; ++s) { mbar_init(bars() + s, 1); mbar_init(bars() + kPipeMaxStages + s, kWarps); }
      mbar_fence_init();
    }
    __syncthreads();
  }
};

// ---- neighbour index helpers (bc 0 periodic; bc 1 Neumann, utils_diff_op.py:19-22,61-64,219-224) ----
struct Nbr { int m, p; double wm, wp; };   // wm/wp: 0 when the one-sided difference is forced to zero (Neumann edge)

__device__ __forceinline__ Nbr nbr(int i, int n, int bc) {
  Nbr r;
  if (bc == 0) { r.m = (i == 0) ? n - 1 : i - 1; r.p = (i == n - 1) ? 0 : i + 1; r.wm = 1.0; r.wp = 1.0; }
  else { r.m = (i == 0) ? 0 : i - 1; r.p = (i == n - 1) ? n - 1 : i + 1; r.wm = (i == 0) ? 0.0 : 1.0; r.wp = (i == n - 1) ? 0.0 : 1.0; }
  return r;
}

// Pointers reach the phase functions through structs, so the compiler would emit GENERIC loads/stores (LD/ST, slower
// address-space resolution, no read-only path).  Telling it that they are global turns them into LDG/STG.
template <typename T> __device__ __forceinline__ T* as_global(T* p) { __builtin_assume(__isGlobal(p)); return p; }

// VW-wide (1 or 2 doubles) global access; VW = 2 needs 16-byte alignment (even index on a 256-B aligned array)
template <int VW> struct Vec { double e[VW]; };
// Explicit global-space accesses (ld.global / st.global): the workspace pointers arrive through structs, so plain
// dereferences compile to GENERIC LD/ST (slower address-space resolution); these compile to LDG/STG.
#ifdef PDHG_ASM_LDST
// round-1 style: volatile inline-asm accesses (fixed program order, "memory" clobbers on the stores)
__device__ __forceinline__ double ldg1(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(__cvta_generic_to_global(p)));
  return v;
}
__device__ __forceinline__ double2 ldg2(const void* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(__cvta_generic_to_global(p)));
  return v;
}
__device__ __forceinline__ void stg1(double* p, double v) {
  asm volatile("st.global.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
__device__ __forceinline__ void stg2(void* p, double2 v) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(p)), "d"(v.x), "d"(v.y) : "memory");
}
#else
// Plain accesses on pointers the compiler knows to be global (as_global): LDG / STG as well, but without the scheduling
// barriers of volatile asm — no "memory" clobber that forces every shared- / local-memory resident value (the argument block,
// the context) to be re-read after each store, and loads may be hoisted and batched freely.
__device__ __forceinline__ double ldg1(const double* p) { return *as_global(p); }
__device__ __forceinline__ double2 ldg2(const void* p) { return *as_global(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void stg1(double* p, double v) { *as_global(p) = v; }
__device__ __forceinline__ void stg2(void* p, double2 v) { *as_global(reinterpret_cast<double2*>(p)) = v; }
#endif
template <int VW> __device__ __forceinline__ Vec<VW> ldv(const double* p) {
  Vec<VW> r;
  if (VW == 2) { const double2 t = ldg2(p); r.e[0] = t.x; r.e[VW - 1] = t.y; }
  else r.e[0] = ldg1(p);
  return r;
}
template <int VW> __device__ __forceinline__ void stv(double* p, const Vec<VW>& v) {
  if (VW == 2) stg2(p, make_double2(v.e[0], v.e[VW - 1]));
  else stg1(p, v.e[0]);
}

// Reductions.  Every phase block-reduces its per-thread sums and writes ONE row of CTA partials (slot0..slot0+N-1 of
// the current epoch's buffer); after the next grid sync `grid_gather` sums the rows of all CTAs in a fixed order, so
// every thread of every CTA holds bit-identical totals and takes identical decisions.
template <int N>
__device__ __forceinline__ void cta_partials(Ctx& c, const double (&vals)[N], int slot0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  double t[N];
#pragma unroll
  for (int q = 0; q < N; ++q) {
    t[q] = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < N; ++q) c.red()[q * kWarps + warp] = t[q];
  }
  __syncthreads();
  if (tid < N) {
    double acc = 0.0;
    for (int w = 0; w < nw; ++w) acc += c.red()[tid * kWarps + w];
    double* part = cargs().w.partials + (size_t)(c.epoch & 1) * gridDim.x * kNQ;
    part[(size_t)(slot0 + tid) * gridDim.x + blockIdx.x] = acc;      // [quantity][CTA]: the gather reads rows contiguously
  }
}

__device__ __forceinline__ void grid_gather(Ctx& c, double (&v)[kNV], int nq = 20) {      // nq: slots [0, nq) are in use this epoch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int G = gridDim.x;
  const double* part = cargs().w.partials + (size_t)(c.epoch & 1) * G * kNQ;
  c.grid.sync();
  for (int q = warp; q < nq; q += nw) {
    double t = 0.0;
    for (int g = lane; g < G; g += 32) t += *((const volatile double*)&part[(size_t)q * G + g]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) c.red()[q * kWarps] = t;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kNV; ++q) v[q] = c.red()[q * kWarps];
  __syncthreads();
  c.epoch++;
}
// total of slot q >= kNV of the last gather (valid until the next cta_partials call of this CTA)
__device__ __forceinline__ double gathered(const Ctx& c, int q) { return c.red()[q * kWarps]; }

// continuity residual at one point (update_fns_in_pdhg.py:72-96).  Index 0 = the point, m/p = its -/+ neighbour.
template <int ND>
__device__ __forceinline__ double cont_point(int egno, bool last_k, double r00, double rnext, double rym, double ryp, double rxm,
                                             double rxp, double a1y_0, double a1y_m, double a2y_0, double a2y_p, double a1x_0,
                                             double a1x_m, double a2x_0, double a2x_p, double cy0, double cym, double cyp, double cx0,
                                             double cxm, double cxp, double wxm, double wxp, double epsl, const Recip& rc,
                                             double c_dt) {
  double f1y_0, f1y_m, f2y_0, f2y_p;
  if (egno == 3) {
    f1y_0 = f_plus(cx0); f1y_m = f1y_0; f2y_0 = f_minus(cx0); f2y_p = f2y_0;     // f_y = x (velocity), set_fns.py:98
  } else {
    f1y_0 = f_plus(-(cy0 * a1y_0)); f1y_m = f_plus(-(cym * a1y_m));
    f2y_0 = f_minus(-(cy0 * a2y_0)); f2y_p = f_minus(-(cyp * a2y_p));
  }
  const double m1y_0 = (r00 + kRhoOffset) * f1y_0, m1y_m = (rym + kRhoOffset) * f1y_m;
  const double m2y_0 = (r00 + kRhoOffset) * f2y_0, m2y_p = (ryp + kRhoOffset) * f2y_p;
  double res;
  if (ND == 2) {
    double f1x_0, f1x_m, f2x_0, f2x_p;
    if (egno == 3) {
      f1x_0 = f_plus(a1x_0); f1x_m = f_plus(a1x_m); f2x_0 = f_minus(a2x_0); f2x_p = f_minus(a2x_p);
    } else {
      f1x_0 = f_plus(-(cx0 * a1x_0)); f1x_m = f_plus(-(cxm * a1x_m));
      f2x_0 = f_minus(-(cx0 * a2x_0)); f2x_p = f_minus(-(cxp * a2x_p));
    }
    const double m1x_0 = (r00 + kRhoOffset) * f1x_0, m1x_m = (rxm + kRhoOffset) * f1x_m;
    const double m2x_0 = (r00 + kRhoOffset) * f2x_0, m2x_p = (rxp + kRhoOffset) * f2x_p;
    res = (rnext - r00) * rc.idt + epsl * ((rxp + rxm - 2 * r00) * rc.idx2) + epsl * ((ryp + rym - 2 * r00) * rc.idy2);
    res -= wxm * (m1x_0 - m1x_m) * rc.idx + wxp * (m2x_p - m2x_0) * rc.idx + (m1y_0 - m1y_m) * rc.idy + (m2y_p - m2y_0) * rc.idy;
  } else {
    res = (rnext - r00) * rc.idt + epsl * ((ryp + rym - 2 * r00) * rc.idy2);
    res -= (m1y_0 - m1y_m) * rc.idy + (m2y_p - m2y_0) * rc.idy;
  }
  if (last_k) res += c_dt;
  return res;
}

// residual of the VW-wide item at (k, i, j .. j+VW-1): loads + cont_point (shared by the tiled and the warp-private phase A)
template <int ND, int VW, int EG>
__device__ __forceinline__ Vec<VW> cont_item(const Ctx& c, const double* rho, const double* a1x, const double* a2x, const double* a1y,
                                             const double* a2y, int k, int i, int j, double epsl, const Recip& rc, double c_dt) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye;
  const size_t n = (size_t)nx * ny;
  constexpr int egno = EG;
  Vec<VW> res;
  const size_t o = (size_t)k * n + (size_t)i * ny;
  const int jm = (j == 0) ? ny - 1 : j - 1, jq = (j + VW == ny) ? 0 : j + VW;
  const Vec<VW> r00 = ldv<VW>(rho + o + j);
  Vec<VW> rnext;
  if (k + 1 < K) rnext = ldv<VW>(rho + o + n + j);
  else {
#pragma unroll
    for (int e = 0; e < VW; ++e) rnext.e[e] = 0.0;
  }
  const double r_l = ldg1(rho + o + jm), r_r = ldg1(rho + o + jq);
  Vec<VW> v1y, v2y;
  double a1y_l = 0.0, a2y_r = 0.0;
  if (egno != 3) {
    v1y = ldv<VW>(a1y + o + j); v2y = ldv<VW>(a2y + o + j);
    a1y_l = ldg1(a1y + o + jm); a2y_r = ldg1(a2y + o + jq);
  } else {
#pragma unroll
    for (int e = 0; e < VW; ++e) { v1y.e[e] = 0.0; v2y.e[e] = 0.0; }
  }
  Vec<VW> rxm, rxp, v1x, v1xm, v2x, v2xp;
  Nbr bx = nbr(i, nx, p.bc_x);
  if (ND == 2) {
    const size_t om = (size_t)k * n + (size_t)bx.m * ny, op = (size_t)k * n + (size_t)bx.p * ny;
    rxm = ldv<VW>(rho + om + j); rxp = ldv<VW>(rho + op + j);
    v1x = ldv<VW>(a1x + o + j); v1xm = ldv<VW>(a1x + om + j);
    v2x = ldv<VW>(a2x + o + j); v2xp = ldv<VW>(a2x + op + j);
  }
  const double cx0 = (ND == 2 || egno == 3) ? c.cx()[i] : 0.0;
  const double cxm = (ND == 2) ? c.cx()[bx.m] : 0.0, cxp = (ND == 2) ? c.cx()[bx.p] : 0.0;
#pragma unroll
  for (int e = 0; e < VW; ++e) {
    const double rym = (e == 0) ? r_l : r00.e[0], ryp = (e == VW - 1) ? r_r : r00.e[VW - 1];
    const double a1m = (e == 0) ? a1y_l : v1y.e[0], a2p = (e == VW - 1) ? a2y_r : v2y.e[VW - 1];
    const int je = j + e;
    const double cym = c.cy()[(e == 0) ? jm : j], cyp = c.cy()[(e == VW - 1) ? jq : j + VW - 1];
    res.e[e] = cont_point<ND>(egno, k == K - 1, r00.e[e], rnext.e[e], rym, ryp, (ND == 2) ? rxm.e[e] : 0.0, (ND == 2) ? rxp.e[e] : 0.0,
                              v1y.e[e], a1m, v2y.e[e], a2p, (ND == 2) ? v1x.e[e] : 0.0, (ND == 2) ? v1xm.e[e] : 0.0,
                              (ND == 2) ? v2x.e[e] : 0.0, (ND == 2) ? v2xp.e[e] : 0.0, c.cy()[je], cym, cyp, cx0, cxm, cxp, bx.wm, bx.wp,
                              epsl, rc, c_dt);
  }
  return res;
}

// residuals of the two x-adjacent double2 items (k, i, j..j+1) and (k, i+1, j..j+1), i even, of a 2-D grid, for a warp whose
// lanes cover 64 consecutive columns (j = j0 + 2 lane): the x-neighbour rows are shared between the two items (16 vector loads
// instead of 20, all issued before the arithmetic) and the y-neighbours come from the adjacent lanes by shuffle (only the two
// edge lanes load a halo word).  Same operands as two cont_item calls => bitwise the same results.  All 32 lanes must call.
template <int EG>
__device__ __forceinline__ void cont_pair2d(const Ctx& c, const double* rho, const double* a1x, const double* a2x, const double* a1y,
                                            const double* a2y, int k, int i, int j, int lane, double epsl, const Recip& rc, double c_dt,
                                            Vec<2>& xa, Vec<2>& xb) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye;
  const size_t n = (size_t)nx * ny, base = (size_t)k * n;
  constexpr int egno = EG;
  const Nbr ba = nbr(i, nx, p.bc_x), bb = nbr(i + 1, nx, p.bc_x);
  const size_t o_m = base + (size_t)ba.m * ny + j, o_a = base + (size_t)i * ny + j, o_b = o_a + ny, o_p = base + (size_t)bb.p * ny + j;
  const Vec<2> R_m = ldv<2>(rho + o_m), R_a = ldv<2>(rho + o_a), R_b = ldv<2>(rho + o_b), R_p = ldv<2>(rho + o_p);
  Vec<2> Rn_a, Rn_b;
  if (k + 1 < K) { Rn_a = ldv<2>(rho + o_a + n); Rn_b = ldv<2>(rho + o_b + n); }
  else { Rn_a.e[0] = Rn_a.e[1] = Rn_b.e[0] = Rn_b.e[1] = 0.0; }
  const Vec<2> X1_m = ldv<2>(a1x + o_m), X1_a = ldv<2>(a1x + o_a), X1_b = ldv<2>(a1x + o_b);
  const Vec<2> X2_a = ldv<2>(a2x + o_a), X2_b = ldv<2>(a2x + o_b), X2_p = ldv<2>(a2x + o_p);
  Vec<2> Y1_a, Y1_b, Y2_a, Y2_b;
  if (egno != 3) { Y1_a = ldv<2>(a1y + o_a); Y1_b = ldv<2>(a1y + o_b); Y2_a = ldv<2>(a2y + o_a); Y2_b = ldv<2>(a2y + o_b); }
  else { Y1_a.e[0] = Y1_a.e[1] = Y1_b.e[0] = Y1_b.e[1] = Y2_a.e[0] = Y2_a.e[1] = Y2_b.e[0] = Y2_b.e[1] = 0.0; }
  const int jm = (j == 0) ? ny - 1 : j - 1, jq = (j + 2 == ny) ? 0 : j + 2;
  const size_t ra = base + (size_t)i * ny, rbo = ra + ny;
  double rl_a = 0.0, rl_b = 0.0, rr_a = 0.0, rr_b = 0.0, y1l_a = 0.0, y1l_b = 0.0, y2r_a = 0.0, y2r_b = 0.0;
  if (lane == 0) {
    rl_a = ldg1(rho + ra + jm); rl_b = ldg1(rho + rbo + jm);
    if (egno != 3) { y1l_a = ldg1(a1y + ra + jm); y1l_b = ldg1(a1y + rbo + jm); }
  }
  if (lane == 31) {
    rr_a = ldg1(rho + ra + jq); rr_b = ldg1(rho + rbo + jq);
    if (egno != 3) { y2r_a = ldg1(a2y + ra + jq); y2r_b = ldg1(a2y + rbo + jq); }
  }
  {
    double t;
    t = __shfl_up_sync(0xffffffffu, R_a.e[1], 1); if (lane != 0) rl_a = t;
    t = __shfl_up_sync(0xffffffffu, R_b.e[1], 1); if (lane != 0) rl_b = t;
    t = __shfl_down_sync(0xffffffffu, R_a.e[0], 1); if (lane != 31) rr_a = t;
    t = __shfl_down_sync(0xffffffffu, R_b.e[0], 1); if (lane != 31) rr_b = t;
    if (egno != 3) {
      t = __shfl_up_sync(0xffffffffu, Y1_a.e[1], 1); if (lane != 0) y1l_a = t;
      t = __shfl_up_sync(0xffffffffu, Y1_b.e[1], 1); if (lane != 0) y1l_b = t;
      t = __shfl_down_sync(0xffffffffu, Y2_a.e[0], 1); if (lane != 31) y2r_a = t;
      t = __shfl_down_sync(0xffffffffu, Y2_b.e[0], 1); if (lane != 31) y2r_b = t;
    }
  }
  const double cx_a = c.cx()[i], cx_b = c.cx()[i + 1], cx_am = c.cx()[ba.m], cx_bp = c.cx()[bb.p];
  const bool last = (k == K - 1);
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const double cy0 = c.cy()[j + e], cym = c.cy()[(e == 0) ? jm : j], cyp = c.cy()[(e == 1) ? jq : j + 1];
    xa.e[e] = cont_point<2>(egno, last, R_a.e[e], Rn_a.e[e], (e == 0) ? rl_a : R_a.e[0], (e == 1) ? rr_a : R_a.e[1], R_m.e[e], R_b.e[e],
                            Y1_a.e[e], (e == 0) ? y1l_a : Y1_a.e[0], Y2_a.e[e], (e == 1) ? y2r_a : Y2_a.e[1], X1_a.e[e], X1_m.e[e],
                            X2_a.e[e], X2_b.e[e], cy0, cym, cyp, cx_a, cx_am, cx_b, ba.wm, ba.wp, epsl, rc, c_dt);
    xb.e[e] = cont_point<2>(egno, last, R_b.e[e], Rn_b.e[e], (e == 0) ? rl_b : R_b.e[0], (e == 1) ? rr_b : R_b.e[1], R_a.e[e], R_p.e[e],
                            Y1_b.e[e], (e == 0) ? y1l_b : Y1_b.e[0], Y2_b.e[e], (e == 1) ? y2r_b : Y2_b.e[1], X1_b.e[e], X1_a.e[e],
                            X2_b.e[e], X2_p.e[e], cy0, cym, cyp, cx_b, cx_a, cx_bp, bb.wm, bb.wp, epsl, rc, c_dt);
  }
}

// ---- phase A: residual rows -> y-FFT -> transposed half spectrum ----
template <int ND, int VW, int EG>
__device__ __noinline__ void phase_A(Ctx& c, int cd, double epsl) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye, nyh = a.nyh, TR = a.TR;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const int rows = K * nx, ntiles = (rows + TR - 1) / TR;
  const int ld = fft_ld(ny);
  double2* buf0 = c.work();
  double2* buf1 = buf0 + (size_t)(TR / 2) * ld;
  const double* rho = as_global(a.w.rho[cd]);
  const double* al = as_global(a.w.alp[cd]);
  double2* ztg = as_global(a.w.zt);
  const double* a1x = al;
  const double* a2x = al + KN;
  const double* a1y = al + (size_t)(2 * ND - 2) * KN;
  const double* a2y = al + (size_t)(2 * ND - 1) * KN;
  const int tid = threadIdx.x, nth = blockDim.x;
  const Recip rc(a.r_idt, a.r_idx, a.r_idy, a.r_idx2, a.r_idy2, 1.0);
  const double c_dt = p.c_on_rho * rc.idt;
  const int ny2 = ny / VW;
  const int dlr = nth / ny2, djp = nth - dlr * ny2;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TR;
    const int nrows = min(TR, rows - r0), npairs = (nrows + 1) >> 1;
    int lr = tid / ny2, jp = tid - lr * ny2;
    while (lr < 2 * npairs) {
      const int j = jp * VW;
      Vec<VW> res;
#pragma unroll
      for (int e = 0; e < VW; ++e) res.e[e] = 0.0;
      if (lr < nrows) {
        const int r = r0 + lr, k = fast_div_exact(r, nx, cargs().inv_nx), i = r - k * nx;
        res = cont_item<ND, VW, EG>(c, rho, a1x, a2x, a1y, a2y, k, i, j, epsl, rc, c_dt);
      }
      double* dst = reinterpret_cast<double*>(&buf0[(size_t)(lr >> 1) * ld + fpad(j)]) + (lr & 1);
#pragma unroll
      for (int e = 0; e < VW; ++e) dst[2 * e] = res.e[e];
      lr += dlr; jp += djp;
      if (jp >= ny2) { jp -= ny2; ++lr; }
    }
    __syncthreads();
    c.tick(0);
    double2* zf = fft_rows(buf0, buf1, a.plan_ye, ld, c.twy(), npairs, 1.0);
    c.tick(1);
    const float inv_np = 1.0f / (float)npairs;
    for (int idx = tid; idx < npairs * nyh; idx += nth) {
      const int ky = fast_div_exact(idx, npairs, inv_np), pr = idx - ky * npairs;
      const int kym = (ky == 0) ? 0 : ny - ky;
      const double2 z1 = zf[(size_t)pr * ld + fpad(ky)], z2 = zf[(size_t)pr * ld + fpad(kym)];
      const int ra = r0 + 2 * pr, ka = fast_div_exact(ra, nx, cargs().inv_nx), ia = ra - ka * nx;
      const double2 va = make_double2(0.5 * (z1.x + z2.x), 0.5 * (z1.y - z2.y));
      if (a.xch_on == 1) {
        // fused forward transpose: interior x-rows go straight into the owner rank's ky-slab (peer memory); ghost rows are dropped
        const int d = ky / a.xch_kyl;
        double2* base = as_global(a.xch_ptr[d]) + (size_t)(ky - d * a.xch_kyl) * (size_t)a.xch_ld + (size_t)a.xch_rank * a.xch_nxl - 1;
        const size_t kstr = (size_t)a.xch_kyl * (size_t)a.xch_ld;
        if (ia >= 1 && ia <= a.xch_nxl) stg2(base + (size_t)ka * kstr + ia, va);
        if (2 * pr + 1 < nrows) {
          const int rb = ra + 1, kb = fast_div_exact(rb, nx, cargs().inv_nx), ib = rb - kb * nx;
          if (ib >= 1 && ib <= a.xch_nxl) stg2(base + (size_t)kb * kstr + ib, make_double2(0.5 * (z1.y + z2.y), 0.5 * (z2.x - z1.x)));
        }
        continue;
      }
      stg2(&ztg[((size_t)ka * nyh + ky) * nx + ia], va);
      if (2 * pr + 1 < nrows) {
        const int rb = ra + 1, kb = fast_div_exact(rb, nx, cargs().inv_nx), ib = rb - kb * nx;
        stg2(&ztg[((size_t)kb * nyh + ky) * nx + ib], make_double2(0.5 * (z1.y + z2.y), 0.5 * (z2.x - z1.x)));
      }
    }
    __syncthreads();
    c.tick(2);
  }
}

// ===================================================================================================================
// Warp-private fast path for 256-point transforms (the BASELINE 2-D grid): 256 = 16 x 16, every lane holds 16 complex
// elements of one row in registers, the only exchange is a 16 x 16 transpose through the warp's own two-row shared-memory
// buffer, and the only synchronisation is __syncwarp().  No block barrier anywhere in phases A / B / C, so the 16 warps of
// an SM run decoupled (loads of some overlap the butterflies of others) and global memory <-> registers moves are direct:
// 16 (B) or 32 (C) independent 16-byte loads per lane in flight.
//   lane = 16 p + jj: complex row p of the warp's pair, elements jj + 16 t.   in: v[t] = x[jj + 16 t]   out: v[q] = X[jj + 16 q]
// ===================================================================================================================
constexpr int kW256Ld = 273;      // fft_ld(256): padded row stride (complex elements)

template <bool INV>
__device__ __forceinline__ void wfft256_first(double2 (&v)[16], double2* rb, int jj) {
  Dft<16, INV>::run(v);                                       // Stockham stage 1 (Ns = 1): outputs q -> element 16 jj + q
  __syncwarp();                                               // earlier reads of the buffer by other lanes are done
#pragma unroll
  for (int q = 0; q < 16; ++q) rb[17 * jj + q] = v[q];        // fpad(16 jj + q)
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = rb[jj + 17 * t];        // fpad(jj + 16 t)
}
template <bool INV>
__device__ __forceinline__ void wfft256_second(double2 (&v)[16], int jj, const double2* tw) {
  // stage 2 (Ns = 16): twiddles W256^(t jj); table loads for t = 1..3 and 4, 8, 12, products for the rest
  double2 wb[4];
  wb[0] = make_double2(1.0, 0.0);
#pragma unroll
  for (int t = 1; t < 4; ++t) { wb[t] = tw[t * jj]; if (INV) wb[t].y = -wb[t].y; }
#pragma unroll
  for (int t = 1; t < 4; ++t) v[t] = cmul(v[t], wb[t]);
#pragma unroll
  for (int a4 = 1; a4 < 4; ++a4) {
    double2 wa = tw[4 * a4 * jj];
    if (INV) wa.y = -wa.y;
    v[4 * a4] = cmul(v[4 * a4], wa);
#pragma unroll
    for (int t = 1; t < 4; ++t) v[4 * a4 + t] = cmul(v[4 * a4 + t], cmul(wa, wb[t]));
  }
  Dft<16, INV>::run(v);                                       // outputs q -> element jj + 16 q
}
template <bool INV>
__device__ __forceinline__ void wfft256(double2 (&v)[16], double2* rb, int jj, const double2* tw) {
  wfft256_first<INV>(v, rb, jj);
  wfft256_second<INV>(v, jj, tw);
}

// phase A, ny == 256: a warp takes 4 consecutive residual rows (= 2 complex rows), computes them with the coalesced
// double2 pattern into its buffer, transforms, splits the two real spectra with lane shuffles (X[m] and X[256 - m] sit
// in lanes jj and 16 - jj) and stores the half spectrum transposed.
template <int ND, int EG>
__device__ __noinline__ void phase_A_w256(Ctx& c, int cd, double epsl) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, nyh = a.nyh;
  constexpr int ny = 256;
  const size_t KN = (size_t)K * nx * ny;
  const int rows = K * nx, nunits = (rows + 3) >> 2;
  const double* rho = as_global(a.w.rho[cd]);
  const double* al = as_global(a.w.alp[cd]);
  double2* ztg = as_global(a.w.zt);
  const double* a1x = al;
  const double* a2x = al + KN;
  const double* a1y = al + (size_t)(2 * ND - 2) * KN;
  const double* a2y = al + (size_t)(2 * ND - 1) * KN;
  const Recip rc(a.r_idt, a.r_idx, a.r_idy, a.r_idx2, a.r_idy2, 1.0);
  const double c_dt = p.c_on_rho * rc.idt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pr = lane >> 4, jj = lane & 15;
  double2* wbuf = c.work() + (size_t)warp * 2 * kW256Ld;
  double2* rb = wbuf + pr * kW256Ld;
  const int src = (lane & 16) | ((16 - jj) & 15);
  const int nwt = gridDim.x * kWarps;
  const bool paired = (ND == 2) && ((nx & 1) == 0);     // rows 2p, 2p+1 of a unit are x-neighbours in the same time row
  for (int u = warp * gridDim.x + blockIdx.x; u < nunits; u += nwt) {
    const int r0 = 4 * u;
    __syncwarp();
#pragma unroll 1
    for (int pp = 0; pp < 2; ++pp) {
      const int ra = r0 + 2 * pp, rbw = ra + 1;
      const bool va = ra < rows, vb = rbw < rows;
      const int ka = fast_div_exact(va ? ra : 0, nx, cargs().inv_nx), ia = (va ? ra : 0) - ka * nx;
      const int kb = fast_div_exact(vb ? rbw : 0, nx, cargs().inv_nx), ib = (vb ? rbw : 0) - kb * nx;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int j = 2 * lane + 64 * cc;
        Vec<2> xa, xb;
        xa.e[0] = xa.e[1] = xb.e[0] = xb.e[1] = 0.0;
        if (ND == 2 && paired) {
          if (va) cont_pair2d<EG>(c, rho, a1x, a2x, a1y, a2y, ka, ia, j, lane, epsl, rc, c_dt, xa, xb);     // (va == vb, warp-uniform)
        } else {
          if (va) xa = cont_item<ND, 2, EG>(c, rho, a1x, a2x, a1y, a2y, ka, ia, j, epsl, rc, c_dt);
          if (vb) xb = cont_item<ND, 2, EG>(c, rho, a1x, a2x, a1y, a2y, kb, ib, j, epsl, rc, c_dt);
        }
        double2* d = wbuf + pp * kW256Ld + fpad(j);
        d[0] = make_double2(xa.e[0], xb.e[0]);
        d[1] = make_double2(xa.e[1], xb.e[1]);
      }
    }
    __syncwarp();
    PDHG_TICK_A(0);
    double2 v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = rb[jj + 17 * t];
    wfft256<false>(v, rb, jj, c.twy());
    PDHG_TICK_A(1);
    // split: Z_a[m] = (X[m] + conj X[N-m]) / 2, Z_b[m] = (X[m] - conj X[N-m]) / (2i), m = jj + 16 q <= 128
    const int ra = r0 + 2 * pr, rbw = ra + 1;
    const bool va = ra < rows, vb = rbw < rows;
    const int ka = fast_div_exact(va ? ra : 0, nx, cargs().inv_nx), ia = (va ? ra : 0) - ka * nx;
    const int kb = fast_div_exact(vb ? rbw : 0, nx, cargs().inv_nx), ib = (vb ? rbw : 0) - kb * nx;
    double2* za = ztg + (size_t)ka * nyh * nx + ia;
    double2* zb = ztg + (size_t)kb * nyh * nx + ib;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double2 mine = (jj == 0) ? v[(16 - q) & 15] : v[15 - q];
      double2 z2;
      z2.x = __shfl_sync(0xffffffffu, mine.x, src);
      z2.y = __shfl_sync(0xffffffffu, mine.y, src);
      const double2 z1 = v[q];
      const size_t mo = (size_t)(jj + 16 * q) * nx;
      if (va) stg2(za + mo, make_double2(0.5 * (z1.x + z2.x), 0.5 * (z1.y - z2.y)));
      if (vb) stg2(zb + mo, make_double2(0.5 * (z1.y + z2.y), 0.5 * (z2.x - z1.x)));
    }
    if (jj == 0) {                       // m = 128 is its own mirror
      const double2 z1 = v[8];
      const size_t mo = (size_t)128 * nx;
      if (va) stg2(za + mo, make_double2(z1.x, 0.0));
      if (vb) stg2(zb + mo, make_double2(z1.y, 0.0));
    }
    PDHG_TICK_A(2);
  }
}

// phase B, nx == 256 (periodic): x-transforms of two (k, ky) rows per warp, global <-> registers.
//   DIR = +1 forward only (pass 1 of a coupled solve), -1 inverse only (pass 3), 0 forward + per-mode scaling + inverse
template <int DIR>
__device__ __noinline__ void phase_B_w256(Ctx& c, double2* zt, int nrows) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pr = lane >> 4, jj = lane & 15;
  double2* rb = c.work() + (size_t)(warp * 2 + pr) * kW256Ld;
  const int npairs = (nrows + 1) >> 1, nwt = gridDim.x * kWarps;
  const double ctk = (p.K == 1) ? p.Ct_over_dt2 : 0.0;
  for (int u = warp * gridDim.x + blockIdx.x; u < npairs; u += nwt) {
    const int row = 2 * u + pr;
    const bool valid = row < nrows;
    double2* g = zt + (size_t)(valid ? row : 0) * 256 + jj;
    double2 v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = valid ? ldg2(g + 16 * t) : make_double2(0.0, 0.0);
    if (DIR >= 0) wfft256<false>(v, rb, jj, c.twx());
    if (DIR == 0) {
      const int ky = row - (row / a.nyh) * a.nyh;
      const double* dg = p.diag + a.ky_off + (valid ? ky : 0);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double rd = 1.0 / (ldg1(dg + (size_t)(jj + 16 * q) * a.nyh_tab) + ctk);
        v[q].x *= rd; v[q].y *= rd;
      }
    }
    if (DIR <= 0) wfft256<true>(v, rb, jj, c.twx());
    if (valid) {
#pragma unroll
      for (int q = 0; q < 16; ++q) stg2(g + 16 * q, v[q]);
    }
  }
}

// phase C, ny == 256: rebuilds the full spectrum of the complex row (u_a + i u_b) from the two half spectra straight from
// global memory, inverse transform, phi update on the lane's 16 (strided) columns of both rows.
__device__ __noinline__ void phase_C_w256(Ctx& c, const double* phi_prev, double* phi_next, double* phib, double tau) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, nyh = a.nyh;
  constexpr int ny = 256;
  const size_t n = (size_t)nx * ny;
  const int rows = K * nx, nunits = (rows + 3) >> 2;
  const double inv_nn = 1.0 / ((double)nx * (double)ny);
  const double2* ztg = as_global(a.w.zt);
  phi_prev = as_global(phi_prev); phi_next = as_global(phi_next);
  if (phib) phib = as_global(phib);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pr = lane >> 4, jj = lane & 15;
  double2* rb = c.work() + (size_t)(warp * 2 + pr) * kW256Ld;
  const int nwt = gridDim.x * kWarps;
  double s_d = 0.0, s_p = 0.0, s_n = 0.0;
  for (int u = warp * gridDim.x + blockIdx.x; u < nunits; u += nwt) {
    const int ra = 4 * u + 2 * pr, rbw = ra + 1;
    const bool va = ra < rows, vb = rbw < rows;
    const int ka = fast_div_exact(va ? ra : 0, nx, cargs().inv_nx), ia = (va ? ra : 0) - ka * nx;
    const int kb = fast_div_exact(vb ? rbw : 0, nx, cargs().inv_nx), ib = (vb ? rbw : 0) - kb * nx;
    const double2* za = ztg + (size_t)ka * nyh * nx + ia;
    const double2* zb = ztg + (size_t)kb * nyh * nx + ib;
    double2 v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const bool lo = (t < 8) || (t == 8 && jj == 0);
      const int m = jj + 16 * t, mm = lo ? m : 256 - m;
      const double2 ua = va ? ldg2(za + (size_t)mm * nx) : make_double2(0.0, 0.0);
      const double2 ub = vb ? ldg2(zb + (size_t)mm * nx) : make_double2(0.0, 0.0);
      v[t] = lo ? make_double2(ua.x - ub.y, ua.y + ub.x) : make_double2(ua.x + ub.y, ub.x - ua.y);
    }
    c.tick(6);
    wfft256_first<true>(v, rb, jj);
    // phi_prev of row a is requested before the second butterfly pass, that of row b before row a's update: the latency of
    // both hides behind arithmetic (the loads are volatile asm, so they stay where they are written)
    const size_t ga = (size_t)(ka + 1) * n + (size_t)ia * ny + jj, gb = (size_t)(kb + 1) * n + (size_t)ib * ny + jj;
    double pa[16], pb[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) pa[q] = va ? ldg1(phi_prev + ga + 16 * q) : 0.0;
    wfft256_second<true>(v, jj, c.twy());
    c.tick(7);
    if (va) {
      const bool acc = (ia >= a.sum_lo && ia < a.sum_hi);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double pp = pa[q];
        const double pn = pp + tau * (v[q].x * inv_nn);
        const double df = pn - pp;
        if (acc) { s_d += df * df; s_p += pp * pp; s_n += is_nan(pn) ? 1.0 : 0.0; }
        stg1(phi_next + ga + 16 * q, pn);
        if (phib) stg1(phib + ga + 16 * q, 2 * pn - pp);
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) pb[q] = vb ? ldg1(phi_prev + gb + 16 * q) : 0.0;
    if (vb) {
      const bool acc = (ib >= a.sum_lo && ib < a.sum_hi);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double pp = pb[q];
        const double pn = pp + tau * (v[q].y * inv_nn);
        const double df = pn - pp;
        if (acc) { s_d += df * df; s_p += pp * pp; s_n += is_nan(pn) ? 1.0 : 0.0; }
        stg1(phi_next + gb + 16 * q, pn);
        if (phib) stg1(phib + gb + 16 * q, 2 * pn - pp);
      }
    }
    c.tick(8);
  }
  const double sums[3] = {s_d, s_p, s_n};
  cta_partials<3>(c, sums, 16);
}

// Thomas recurrences over k for one real component of one Fourier mode (item w of a [K][2*modes] real array)
__device__ __forceinline__ void thomas_component(double* ztd, const double* den, const double* tu, int K, size_t modes2, size_t w,
                                                 double ct2) {
  const size_t m = w >> 1, modes = modes2 >> 1;
  ztd = as_global(ztd); den = as_global(den); tu = as_global(tu);
  double bp = 0.0;
  int k = 0;
  for (; k + 16 <= K; k += 16) {    // 32 independent loads ahead of the dependent chain
    double v[16], dn[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) { v[q] = ldg1(ztd + (size_t)(k + q) * modes2 + w); dn[q] = ldg1(den + (size_t)(k + q) * modes + m); }
#pragma unroll
    for (int q = 0; q < 16; ++q) { bp = (v[q] + ct2 * bp) * dn[q]; stg1(ztd + (size_t)(k + q) * modes2 + w, bp); }
  }
  for (; k + 8 <= K; k += 8) {      // 16 independent loads ahead of the dependent chain
    double v[8], dn[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = ldg1(ztd + (size_t)(k + q) * modes2 + w); dn[q] = ldg1(den + (size_t)(k + q) * modes + m); }
#pragma unroll
    for (int q = 0; q < 8; ++q) { bp = (v[q] + ct2 * bp) * dn[q]; stg1(ztd + (size_t)(k + q) * modes2 + w, bp); }
  }
  for (; k + 4 <= K; k += 4) {
    double v[4], dn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = ldg1(ztd + (size_t)(k + q) * modes2 + w); dn[q] = ldg1(den + (size_t)(k + q) * modes + m); }
#pragma unroll
    for (int q = 0; q < 4; ++q) { bp = (v[q] + ct2 * bp) * dn[q]; stg1(ztd + (size_t)(k + q) * modes2 + w, bp); }
  }
  for (; k < K; ++k) {
    bp = (ldg1(ztd + (size_t)k * modes2 + w) + ct2 * bp) * ldg1(den + (size_t)k * modes + m);
    stg1(ztd + (size_t)k * modes2 + w, bp);
  }
  double xs = bp;
  k = K - 2;
  for (; k - 15 >= 0; k -= 16) {
    double v[16], tv[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) { v[q] = ldg1(ztd + (size_t)(k - q) * modes2 + w); tv[q] = ldg1(tu + (size_t)(k - q) * modes + m); }
#pragma unroll
    for (int q = 0; q < 16; ++q) { xs = v[q] - tv[q] * xs; stg1(ztd + (size_t)(k - q) * modes2 + w, xs); }
  }
  for (; k - 7 >= 0; k -= 8) {
    double v[8], tv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = ldg1(ztd + (size_t)(k - q) * modes2 + w); tv[q] = ldg1(tu + (size_t)(k - q) * modes + m); }
#pragma unroll
    for (int q = 0; q < 8; ++q) { xs = v[q] - tv[q] * xs; stg1(ztd + (size_t)(k - q) * modes2 + w, xs); }
  }
  for (; k - 3 >= 0; k -= 4) {
    double v[4], tv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { v[q] = ldg1(ztd + (size_t)(k - q) * modes2 + w); tv[q] = ldg1(tu + (size_t)(k - q) * modes + m); }
#pragma unroll
    for (int q = 0; q < 4; ++q) { xs = v[q] - tv[q] * xs; stg1(ztd + (size_t)(k - q) * modes2 + w, xs); }
  }
  for (; k >= 0; --k) {
    xs = ldg1(ztd + (size_t)k * modes2 + w) - ldg1(tu + (size_t)k * modes + m) * xs;
    stg1(ztd + (size_t)k * modes2 + w, xs);
  }
}

// ===================================================================================================================
// Phase B in ONE pass per ky-slab (nx == 256, coupled modes): a CTA owns the [K][256] complex slab of one ky.  The rows are
// transformed along x as they arrive (global -> registers -> warp-private 256-point FFT -> shared memory), the per-mode
// tridiagonal systems in t are solved IN SHARED MEMORY by 512 threads (one per real component of a mode, conflict-free), and
// the rows leave through the inverse transform straight to global memory: the spectrum is read once and written once instead
// of three times each (x-FFT pass, Thomas pass, inverse pass), and the two grid-wide barriers between the passes are gone.
// A slab larger than the work area (K = 64: 256 KB) is processed in nchunk balanced k-chunks: forward elimination chunk by
// chunk (the eliminated rows of all but the last chunk are parked in their own global rows), back substitution from the last
// chunk down (the parked rows are re-read by the Thomas threads themselves, fully coalesced).  Same butterflies, same
// recurrences (thomas_component) => bit-identical to the three-pass version.
// The transform's 16 x 16 exchange runs inside the row's own (unpadded) shared-memory slot with an XOR swizzle
// (element 16 J + Q at 16 J + (Q ^ J)): conflict-free for the 16-byte accesses of a quarter warp, no scratch rows.
// ===================================================================================================================
template <bool INV>
__device__ __forceinline__ void wfft256_first_sw(double2 (&v)[16], double2* row, int jj) {
  Dft<16, INV>::run(v);                                       // outputs q -> element 16 jj + q
  __syncwarp();                                               // every lane holds its inputs in registers: the slot is free
#pragma unroll
  for (int q = 0; q < 16; ++q) row[16 * jj + (q ^ jj)] = v[q];
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = row[16 * t + (jj ^ t)];  // element jj + 16 t
  __syncwarp();                                               // all exchange reads done before the slot is overwritten
}

__device__ __noinline__ void phase_B_slab(Ctx& c) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nyh = a.nyh;
  constexpr int nx = 256;
  const double ct2 = p.Ct_over_dt2;
  double2* zt = as_global(a.w.zt);
  const double* den = as_global(a.w.den);
  const size_t modes = (size_t)nyh * nx;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pr = lane >> 4, jj = lane & 15;
  // work area = Hmax row slots of the spectrum (4 KB each) + Hmax rows of reciprocal pivots (2 KB each, staged by TMA);
  // a chunk with an odd row count leaves one slot free (the idle half-warp's scratch slot)
  const int Hmax = a.work_bytes / (nx * 24);
  int nchunk = (K + Hmax - 1) / Hmax, H = (K + nchunk - 1) / nchunk;
  if (((H & 1) || ((K - (nchunk - 1) * H) & 1)) && H == Hmax) {   // an odd chunk needs the spare slot
    nchunk = (K + Hmax - 2) / (Hmax - 1); H = (K + nchunk - 1) / nchunk;
  }
  double2* S = c.work();
  double* Sd = reinterpret_cast<double*>(S);
  double* dsm = Sd + (size_t)Hmax * nx * 2;                    // [Hmax][256] reciprocal pivots of the current chunk
  uint64_t* bar = c.bars();                                    // full[0] of the pipeline barriers (1 arrival + transaction bytes)
  const int pmask = (a.mode == MODE_PHASE) ? a.dbg_pass : 7;
  if (!(pmask & 1)) return;
  fence_proxy_async_smem();                                    // generic-proxy accesses of the work area before the bulk copies
  __syncthreads();
  // the chunk's rows of the reciprocal-pivot table (2 KB each, contiguous over kx) -> shared memory, asynchronously
  auto stage_pivots = [&](int ky, int k0, int nr) {
    if (warp == 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)nr * nx * 8u);
      __syncwarp();
      for (int r = lane; r < nr; r += 32) bulk_g2s(dsm + (size_t)r * nx, den + (size_t)(k0 + r) * modes + (size_t)ky * nx, nx * 8u, bar);
    }
  };
  auto wait_pivots = [&]() { mbar_wait(bar, c.pipe.par_full & 1u); c.pipe.par_full ^= 1u; };
  for (int ky = blockIdx.x; ky < nyh; ky += gridDim.x) {
    double bp = 0.0;
    // ---- forward: x-FFT of the chunk's rows into shared memory, then forward elimination over its k ----
    for (int ch = 0; ch < nchunk; ++ch) {
      const int k0 = ch * H, nr = min(H, K - k0);
      stage_pivots(ky, k0, nr);                                // lands while the rows are loaded and transformed
      for (int r0 = 2 * warp; r0 < nr; r0 += 2 * kWarps) {      // a warp takes two rows; with an odd count its second half idles
        const int r = r0 + pr;                                 // (all 32 lanes stay in the loop: the exchange uses __syncwarp)
        const bool valid = r < nr;
        double2* g = zt + ((size_t)(k0 + (valid ? r : r0)) * nyh + ky) * nx + jj;
        double2 v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = valid ? ldg2(g + 16 * t) : make_double2(0.0, 0.0);
        double2* row = S + (size_t)(valid ? r : Hmax - 1) * nx;   // an idle half scribbles over the (unused) last slot
        wfft256_first_sw<false>(v, row, jj);
        wfft256_second<false>(v, jj, c.twx());
        if (valid) {
#pragma unroll
          for (int q = 0; q < 16; ++q) row[jj + 16 * q] = v[q];
        }
      }
      __syncthreads();
      PDHG_TICK_B(0);
      wait_pivots();
      // thread tid = real component tid & 1 of mode kx = tid >> 1; rows k0 .. k0 + nr - 1   (thomas_component's forward loop)
      const double* dn_t = dsm + (tid >> 1);
#pragma unroll 4
      for (int r = 0; r < nr; ++r) { bp = (Sd[(size_t)r * (2 * nx) + tid] + ct2 * bp) * dn_t[(size_t)r * nx]; Sd[(size_t)r * (2 * nx) + tid] = bp; }
      PDHG_TICK_B(1);
      if (ch < nchunk - 1) {
        // park the eliminated rows in their own global rows (coalesced: a row is 512 consecutive doubles)
        double* zd = reinterpret_cast<double*>(zt);
        for (int r = 0; r < nr; ++r) stg1(zd + (((size_t)(k0 + r) * nyh + ky) * nx) * 2 + tid, Sd[(size_t)r * (2 * nx) + tid]);
        fence_proxy_async_smem();
        __syncthreads();
        PDHG_TICK_B(2);
      }
    }
    // ---- backward: substitution from the last row down (super-diagonal of U: tu[k] = -ct2 * rpivot[k], as build_tables
    // defines it), inverse x-FFT chunk by chunk ----
    double xs = bp;                                            // x[K-1] = b'[K-1]
    for (int ch = nchunk - 1; ch >= 0; --ch) {
      const int k0 = ch * H, nr = min(H, K - k0);
      const bool last = (ch == nchunk - 1);
      const double* zd = reinterpret_cast<const double*>(zt);
      const double* dn_t = dsm + (tid >> 1);
      int r = last ? nr - 2 : nr - 1;                          // the very last row is already solved (and sits in S)
      if (last) {
#pragma unroll 4
        for (; r >= 0; --r) { xs = Sd[(size_t)r * (2 * nx) + tid] - (-ct2 * dn_t[(size_t)r * nx]) * xs; Sd[(size_t)r * (2 * nx) + tid] = xs; }
      } else {
        wait_pivots();                                         // staged while the previous chunk's rows were transformed back
        for (; r - 7 >= 0; r -= 8) {                           // the parked rows come back from global memory, 8 loads deep
          double bv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) bv[q] = ldg1(zd + (((size_t)(k0 + r - q) * nyh + ky) * nx) * 2 + tid);
#pragma unroll
          for (int q = 0; q < 8; ++q) { xs = bv[q] - (-ct2 * dn_t[(size_t)(r - q) * nx]) * xs; Sd[(size_t)(r - q) * (2 * nx) + tid] = xs; }
        }
        for (; r >= 0; --r) {
          xs = ldg1(zd + (((size_t)(k0 + r) * nyh + ky) * nx) * 2 + tid) - (-ct2 * dn_t[(size_t)r * nx]) * xs;
          Sd[(size_t)r * (2 * nx) + tid] = xs;
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      PDHG_TICK_B(3);
      if (ch > 0) stage_pivots(ky, (ch - 1) * H, min(H, K - (ch - 1) * H));   // the next (earlier) chunk's pivots, under the inverse transforms
      for (int r0 = 2 * warp; r0 < nr; r0 += 2 * kWarps) {
        const int rr = r0 + pr;
        const bool valid = rr < nr;
        double2* row = S + (size_t)(valid ? rr : Hmax - 1) * nx;
        double2 v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = valid ? row[jj + 16 * t] : make_double2(0.0, 0.0);
        wfft256_first_sw<true>(v, row, jj);
        wfft256_second<true>(v, jj, c.twx());
        if (valid) {
          double2* g = zt + ((size_t)(k0 + rr) * nyh + ky) * nx + jj;
#pragma unroll
          for (int q = 0; q < 16; ++q) stg2(g + 16 * q, v[q]);
        }
      }
      __syncthreads();
      PDHG_TICK_B(4);
    }
  }
}

// ---- phase B: x-FFT, t-solve per mode, inverse x-FFT (in place on zt) ----
__device__ __noinline__ void phase_B(Ctx& c) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, nyh = a.nyh;
  const int tid = threadIdx.x, nth = blockDim.x;
  const double ct2 = p.Ct_over_dt2;
  double2* zt = as_global(a.w.zt);
  const size_t modes = (size_t)nyh * nx;
  const bool coupled = (K > 1 && ct2 != 0.0);
  if (nx == 1) {
    // 1-D: no x transform
    if (coupled) {
      for (size_t w = (size_t)blockIdx.x * nth + tid; w < 2 * modes; w += (size_t)gridDim.x * nth)
        thomas_component(reinterpret_cast<double*>(zt), a.w.den, a.w.tu, K, 2 * modes, w, ct2);
    } else {
      const size_t total = (size_t)K * modes;
      for (size_t g = (size_t)blockIdx.x * nth + tid; g < total; g += (size_t)gridDim.x * nth) {
        const double d = p.diag[g % modes] + ((K == 1) ? ct2 : 0.0);
        const double2 v = zt[g];
        zt[g] = make_double2(v.x / d, v.y / d);
      }
    }
    return;
  }
  const int ld = fft_ld(nx);
  const int TKY = a.TKY;
  double2* buf0 = c.work();
  double2* buf1 = buf0 + (size_t)TKY * ld;
  const int ntile = (nyh + TKY - 1) / TKY;
  const int nunits = K * ntile;
  const int pmask = (a.mode == MODE_PHASE) ? a.dbg_pass : 7;
  if (a.fast_x) {
    if (!coupled) { if (pmask & 1) phase_B_w256<0>(c, zt, K * nyh); return; }
    if (a.b_slab) { phase_B_slab(c); return; }
    if (pmask & 1) phase_B_w256<1>(c, zt, K * nyh);
    c.grid.sync();
    c.tick(3);
    for (size_t w = (size_t)blockIdx.x * nth + tid; w < ((pmask & 2) ? 2 * modes : 0); w += (size_t)gridDim.x * nth)
      thomas_component(reinterpret_cast<double*>(zt), a.w.den, a.w.tu, K, 2 * modes, w, ct2);
    c.grid.sync();
    c.tick(4);
    if (pmask & 4) phase_B_w256<-1>(c, zt, K * nyh);
    return;
  }
  // pass 1: x-FFT of every (k, ky) row; uncoupled modes are solved and transformed back in the same pass
  for (int u = blockIdx.x; u < ((pmask & 1) ? nunits : 0); u += gridDim.x) {
    const int k = u / ntile, ky0 = (u - k * ntile) * TKY;
    const int nr = min(TKY, nyh - ky0);
    if (a.xch_on == 2 && a.xch_pull) {
      // fused forward transpose: the rows of this ky-slab are gathered straight from the x-slabs of their owners (peer memory)
      const float inv_nxl = 1.0f / (float)a.xch_nxl;
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
        const int d = fast_div_exact(kx, a.xch_nxl, inv_nxl);
        buf0[(size_t)t * ld + fpad(kx)] =
            ldg2(as_global(a.xch_ptr[d]) + ((size_t)k * a.xch_nyh + a.ky_off + ky0 + t) * (size_t)a.xch_ld + 1 + (kx - d * a.xch_nxl));
      }
    } else {
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
        buf0[(size_t)t * ld + fpad(kx)] = ldg2(&zt[((size_t)k * nyh + ky0 + t) * nx + kx]);
      }
    }
    __syncthreads();
    if (p.bc_x == 1) {
      // Neumann in x (egno 3): DCT-II along x, divide, inverse DCT (utils_precond.py:159-161,172-174), evaluated directly
      // (O(nx^2) per row; this path only serves the small Newton example and is uncoupled: K = 1)
      const double* ct = p.dct_cos;
      const int m4 = 4 * nx;
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
        double sx = 0.0, sy = 0.0;
        int m = kx % m4;                                  // index kx (2 i + 1) mod 4 nx, advanced by 2 kx per i
        for (int i = 0; i < nx; ++i) {
          const double2 v = buf0[(size_t)t * ld + fpad(i)];
          const double cv = ct[m];
          sx += v.x * cv; sy += v.y * cv;
          m += 2 * kx; if (m >= m4) m -= m4; if (m >= m4) m -= m4;
        }
        const double rd = 1.0 / (p.diag[(size_t)kx * a.nyh_tab + a.ky_off + ky0 + t] + ct2);
        buf1[(size_t)t * ld + fpad(kx)] = make_double2(2.0 * sx * rd, 2.0 * sy * rd);
      }
      __syncthreads();
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = fast_div_exact(idx, nx, cargs().inv_nx), i = idx - t * nx;
        const double2 y0 = buf1[(size_t)t * ld + fpad(0)];
        double sx = 0.5 * y0.x, sy = 0.5 * y0.y;
        const int step = (2 * i + 1) % m4;
        int m = step;
        for (int kx = 1; kx < nx; ++kx) {
          const double2 v = buf1[(size_t)t * ld + fpad(kx)];
          const double cv = ct[m];
          sx += v.x * cv; sy += v.y * cv;
          m += step; if (m >= m4) m -= m4;
        }
        zt[((size_t)k * nyh + ky0 + t) * nx + i] = make_double2(sx, sy);    // (the 1/nx of idct is applied in phase C)
      }
      __syncthreads();
      continue;
    }
    double2* zf = fft_rows(buf0, buf1, a.plan_xe, ld, c.twx(), nr, 1.0);
    if (!coupled) {
      double2* zo = (zf == buf0) ? buf1 : buf0;
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
        const double rd = 1.0 / (ldg1(p.diag + (size_t)kx * a.nyh_tab + a.ky_off + ky0 + t) + ((K == 1) ? ct2 : 0.0));
        const double2 v = zf[(size_t)t * ld + fpad(kx)];
        zf[(size_t)t * ld + fpad(kx)] = make_double2(v.x * rd, v.y * rd);
      }
      __syncthreads();
      zf = fft_rows(zf, zo, a.plan_xe, ld, c.twx(), nr, -1.0);
    }
    if (a.xch_on == 2) {
      // fused backward transpose (uncoupled modes only: K = 1): column x of this ky-slab goes straight into the x-slab of its owner
      const float inv_nxl = 1.0f / (float)a.xch_nxl;
      for (int idx = tid; idx < nr * nx; idx += nth) {
        const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
        const int d = fast_div_exact(kx, a.xch_nxl, inv_nxl);
        stg2(as_global(a.xch_ptr[d]) + ((size_t)k * a.xch_nyh + a.ky_off + ky0 + t) * (size_t)a.xch_ld + 1 + (kx - d * a.xch_nxl),
             zf[(size_t)t * ld + fpad(kx)]);
      }
      __syncthreads();
      continue;
    }
    for (int idx = tid; idx < nr * nx; idx += nth) {
      const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
      stg2(&zt[((size_t)k * nyh + ky0 + t) * nx + kx], zf[(size_t)t * ld + fpad(kx)]);
    }
    __syncthreads();
  }
  if (!coupled) return;
  c.grid.sync();
  c.tick(3);
  // pass 2: Thomas over k, one thread per real component of a mode (coalesced across modes)
  for (size_t w = (size_t)blockIdx.x * nth + tid; w < ((pmask & 2) ? 2 * modes : 0); w += (size_t)gridDim.x * nth)
    thomas_component(reinterpret_cast<double*>(zt), a.w.den, a.w.tu, K, 2 * modes, w, ct2);
  c.grid.sync();
  c.tick(4);
  // pass 3: inverse x-FFT of every row
  for (int u = blockIdx.x; u < ((pmask & 4) ? nunits : 0); u += gridDim.x) {
    const int k = u / ntile, ky0 = (u - k * ntile) * TKY;
    const int nr = min(TKY, nyh - ky0);
    for (int idx = tid; idx < nr * nx; idx += nth) {
      const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
      buf0[(size_t)t * ld + fpad(kx)] = ldg2(&zt[((size_t)k * nyh + ky0 + t) * nx + kx]);
    }
    __syncthreads();
    double2* zu = fft_rows(buf0, buf1, a.plan_xe, ld, c.twx(), nr, -1.0);
    for (int idx = tid; idx < nr * nx; idx += nth) {
      const int t = fast_div_exact(idx, nx, cargs().inv_nx), kx = idx - t * nx;
      stg2(&zt[((size_t)k * nyh + ky0 + t) * nx + kx], zu[(size_t)t * ld + fpad(kx)]);
    }
    __syncthreads();
  }
}

// ---- phase C: inverse y-FFT, phi update.  CTA partials: slot 16 (sum dphi^2), 17 (sum phi_prev^2 rows>=1), 18 (NaN count)
template <int VW>
__device__ __noinline__ void phase_C(Ctx& c, const double* phi_prev, double* phi_next, double* phib, double tau) {
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye, nyh = a.nyh, TR = a.TR;
  const size_t n = (size_t)nx * ny;
  const int rows = K * nx, ntiles = (rows + TR - 1) / TR;
  const int ld = fft_ld(ny);
  double2* buf0 = c.work();
  double2* buf1 = buf0 + (size_t)(TR / 2) * ld;
  const int tid = threadIdx.x, nth = blockDim.x;
  const double inv_nn = 1.0 / ((double)nx * (double)ny);
  const int ny2 = ny / VW;
  const int dlr = nth / ny2, djp = nth - dlr * ny2;
  const double2* ztg = as_global(a.w.zt);
  phi_prev = as_global(phi_prev); phi_next = as_global(phi_next);
  if (phib) phib = as_global(phib);
  double s_d = 0.0, s_p = 0.0, s_n = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TR;
    const int nrows = min(TR, rows - r0), npairs = (nrows + 1) >> 1;
    const float inv_np = 1.0f / (float)npairs;
    for (int idx = tid; idx < npairs * nyh; idx += nth) {
      const int ky = fast_div_exact(idx, npairs, inv_np), pr = idx - ky * npairs;
      const int ra = r0 + 2 * pr, ka = fast_div_exact(ra, nx, cargs().inv_nx), ia = ra - ka * nx;
      const double2 ua = ldg2(&ztg[((size_t)ka * nyh + ky) * nx + ia]);
      double2 ub = make_double2(0.0, 0.0);
      if (2 * pr + 1 < nrows) {
        const int rb = ra + 1, kb = fast_div_exact(rb, nx, cargs().inv_nx), ib = rb - kb * nx;
        ub = ldg2(&ztg[((size_t)kb * nyh + ky) * nx + ib]);
      }
      buf0[(size_t)pr * ld + fpad(ky)] = make_double2(ua.x - ub.y, ua.y + ub.x);
      const int kym = ny - ky;
      if (ky != 0 && kym != ky) buf0[(size_t)pr * ld + fpad(kym)] = make_double2(ua.x + ub.y, ub.x - ua.y);
    }
    __syncthreads();
    c.tick(6);
    double2* zu = fft_rows(buf0, buf1, a.plan_ye, ld, c.twy(), npairs, -1.0);
    c.tick(7);
    int lr = tid / ny2, jp = tid - lr * ny2;
    while (lr < nrows) {
      const int j = jp * VW;
      const int r = r0 + lr, k = fast_div_exact(r, nx, cargs().inv_nx), i = r - k * nx;
      const size_t g = (size_t)(k + 1) * n + (size_t)i * ny + j;
      const double* zsrc = reinterpret_cast<const double*>(&zu[(size_t)(lr >> 1) * ld + fpad(j)]) + (lr & 1);
      const Vec<VW> pp = ldv<VW>(phi_prev + g);
      Vec<VW> pn, pb;
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const double u = zsrc[2 * e] * inv_nn;
        pn.e[e] = pp.e[e] + tau * u;
        const double df = pn.e[e] - pp.e[e];
        if (i >= a.sum_lo && i < a.sum_hi) { s_d += df * df; s_p += pp.e[e] * pp.e[e]; s_n += is_nan(pn.e[e]) ? 1.0 : 0.0; }
        pb.e[e] = 2 * pn.e[e] - pp.e[e];
      }
      stv<VW>(phi_next + g, pn);
      if (phib) stv<VW>(phib + g, pb);
      lr += dlr; jp += djp;
      if (jp >= ny2) { jp -= ny2; ++lr; }
    }
    __syncthreads();
    c.tick(8);
  }
  const double sums[3] = {s_d, s_p, s_n};
  cta_partials<3>(c, sums, 16);
}

// dual update at one point (update_fns_in_pdhg.py:99-133,150-165; set_fns.py prox formulas)
template <int ND>
__device__ __forceinline__ void dual_point(int egno, double c0, double cxm, double cxp, double cym, double cyp, double pbk, double ro,
                                           const double (&ao)[2 * ND], double cx, double cy, double wxm, double wxp, double sigma,
                                           double epsl, const Recip& rc, double& rn, double (&an)[2 * ND]) {
  const double pinv = (ro + kRhoOffset) * rc.isig;
  const double rinv = prox_rinv(egno, pinv);
  const double dyr = (cyp - c0) * rc.idy, dyl = (c0 - cym) * rc.idy;
  double vec = (c0 - pbk) * rc.idt;
  double adv = 0.0, L = 0.0;
  if (ND == 2) {
    const double dxr = wxp * (cxp - c0) * rc.idx, dxl = wxm * (c0 - cxm) * rc.idx;
    an[0] = prox_alp(egno, ao[0], dxr, pinv, rinv, cx, true);
    an[1] = prox_alp(egno, ao[1], dxl, pinv, rinv, cx, false);
    const double f1 = (egno == 3) ? f_plus(an[0]) : f_plus(-(cx * an[0]));
    const double f2 = (egno == 3) ? f_minus(an[1]) : f_minus(-(cx * an[1]));
    vec -= epsl * ((cxp + cxm - 2 * c0) * rc.idx2);
    adv = dxr * f1 + dxl * f2;
    L = lagr(egno, an[0]) + lagr(egno, an[1]);
  }
  constexpr int jy = 2 * ND - 2;
  double f1, f2;
  if (egno == 3) {
    an[jy] = ao[jy]; an[jy + 1] = ao[jy + 1];          // set_fns.py:110: the y pair is passed through
    f1 = f_plus(cx); f2 = f_minus(cx);
  } else {
    an[jy] = prox_alp(egno, ao[jy], dyr, pinv, rinv, cy, true);
    an[jy + 1] = prox_alp(egno, ao[jy + 1], dyl, pinv, rinv, cy, false);
    f1 = f_plus(-(cy * an[jy])); f2 = f_minus(-(cy * an[jy + 1]));
    L += lagr(egno, an[jy]) + lagr(egno, an[jy + 1]);
  }
  vec -= epsl * ((cyp + cym - 2 * c0) * rc.idy2);
  adv += dyr * f1 + dyl * f2;
  vec -= adv;
  vec -= L;
  rn = relu_nan(ro + sigma * vec);
}

// ---- phase D: dual sweep(s).  src -> dst (may alias when NS == 1 and the direct-load variant runs).
// CTA partials: slots 0..1 rho (diff^2, next^2), 2+2q..3+2q alp q; 15 NaN count of rho_next  (10..14 belong to phase E)
// NS = 2 fuses two consecutive sweeps into one pass over memory: the dual update at a point needs phi_bar's stencil but only the
// point's own rho / alp, so the second sweep runs on the first one's results while they are still in registers (same arithmetic
// per point, same per-thread summation order => bit-identical iterates and error sums; its sums go to slots 20..35).
// With NS = 2, `nx_sw` (0..kFuseMax-2) further sweeps follow in the same pass; their sums are accumulated in per-thread
// shared-memory slots (same order of additions as registers would see) and go to slots 36.. .
template <int NA> struct DualSums {     // per-thread running sums of one pass: sweep 1 (s_*) and sweep 2 of a fused pair (t_*)
  double s_dr, s_rr, s_nan, s_da[NA], s_aa[NA];
  double t_dr, t_rr, t_nan, t_da[NA], t_aa[NA];
  __device__ __forceinline__ void clear() {
    s_dr = s_rr = s_nan = t_dr = t_rr = t_nan = 0.0;
#pragma unroll
    for (int q = 0; q < NA; ++q) { s_da[q] = s_aa[q] = t_da[q] = t_aa[q] = 0.0; }
  }
};
template <int NA> struct DualAccLayout { static constexpr int kAcc = 3 + 2 * NA; };   // shared-memory slots per extra sweep and thread

// all sweeps of one pass for the VW-wide item at column j, operands in registers: cc = phi_bar[k+1] at the item, pk = phi_bar[k],
// (cxm, cxp) = phi_bar[k+1] at the x-neighbours, (c_l, c_r) = its y-neighbours left of / right of the item, (ro, ao) = the prox centre
template <int ND, int VW, int EG, int NS, bool XTRA>
__device__ __forceinline__ void dual_item(const Ctx& c, const Vec<VW>& cc, const Vec<VW>& pk, const Vec<VW>& ro, const Vec<VW> (&ao)[2 * ND],
                                          const Vec<VW>& cxm, const Vec<VW>& cxp, double c_l, double c_r, double cx, int j, double wxm,
                                          double wxp, bool acc_on, double sigma, double epsl, const Recip& rc, int nx_sw, double* xacc,
                                          int astr, DualSums<2 * ND>& S, Vec<VW>& rn, Vec<VW> (&an)[2 * ND]) {
  constexpr int NA = 2 * ND, egno = EG;
  constexpr int kAcc = DualAccLayout<NA>::kAcc;
#pragma unroll
  for (int e = 0; e < VW; ++e) {
    const double cym = (e == 0) ? c_l : cc.e[0], cyp = (e == VW - 1) ? c_r : cc.e[VW - 1];
    double aoe[NA], ane[NA], rne;
#pragma unroll
    for (int q = 0; q < NA; ++q) aoe[q] = ao[q].e[e];
    dual_point<ND>(egno, cc.e[e], (ND == 2) ? cxm.e[e] : 0.0, (ND == 2) ? cxp.e[e] : 0.0, cym, cyp, pk.e[e], ro.e[e], aoe, cx,
                   c.cy()[j + e], wxm, wxp, sigma, epsl, rc, rne, ane);
    if (acc_on) {
      double d = rne - ro.e[e];
      S.s_dr += d * d; S.s_rr += rne * rne; S.s_nan += is_nan(rne) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < NA; ++q) { d = ane[q] - aoe[q]; S.s_da[q] += d * d; S.s_aa[q] += ane[q] * ane[q]; }
    }
    if (NS == 2) {
      double an2[NA], rn2;
      dual_point<ND>(egno, cc.e[e], (ND == 2) ? cxm.e[e] : 0.0, (ND == 2) ? cxp.e[e] : 0.0, cym, cyp, pk.e[e], rne, ane, cx,
                     c.cy()[j + e], wxm, wxp, sigma, epsl, rc, rn2, an2);
      if (acc_on) {
        double d = rn2 - rne;
        S.t_dr += d * d; S.t_rr += rn2 * rn2; S.t_nan += is_nan(rn2) ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < NA; ++q) { d = an2[q] - ane[q]; S.t_da[q] += d * d; S.t_aa[q] += an2[q] * an2[q]; }
      }
      rne = rn2;
#pragma unroll
      for (int q = 0; q < NA; ++q) ane[q] = an2[q];
    }
    rn.e[e] = rne;
#pragma unroll
    for (int q = 0; q < NA; ++q) an[q].e[e] = ane[q];
  }
  if (XTRA) {
    // further sweeps of the fused pass: the VW points of the item advance together (independent chains for the fp64
    // pipe), each sum is read from / written to its shared-memory slot once per item, additions in the order e = 0, 1
#pragma unroll 1
    for (int sw = 0; sw < nx_sw; ++sw) {
      double rn2[VW], an2[VW][NA];
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const double cym = (e == 0) ? c_l : cc.e[0], cyp = (e == VW - 1) ? c_r : cc.e[VW - 1];
        double aoe[NA];
#pragma unroll
        for (int q = 0; q < NA; ++q) aoe[q] = an[q].e[e];
        dual_point<ND>(egno, cc.e[e], (ND == 2) ? cxm.e[e] : 0.0, (ND == 2) ? cxp.e[e] : 0.0, cym, cyp, pk.e[e], rn.e[e], aoe, cx,
                       c.cy()[j + e], wxm, wxp, sigma, epsl, rc, rn2[e], an2[e]);
      }
      if (acc_on) {
        double* as = xacc + sw * kAcc * astr;
        double s0 = as[0], s1 = as[astr], s2 = as[2 * astr];
#pragma unroll
        for (int e = 0; e < VW; ++e) { const double d = rn2[e] - rn.e[e]; s0 += d * d; s1 += rn2[e] * rn2[e]; s2 += is_nan(rn2[e]) ? 1.0 : 0.0; }
        as[0] = s0; as[astr] = s1; as[2 * astr] = s2;
#pragma unroll
        for (int q = 0; q < NA; ++q) {
          double sd = as[(3 + 2 * q) * astr], sa = as[(4 + 2 * q) * astr];
#pragma unroll
          for (int e = 0; e < VW; ++e) { const double d = an2[e][q] - an[q].e[e]; sd += d * d; sa += an2[e][q] * an2[e][q]; }
          as[(3 + 2 * q) * astr] = sd; as[(4 + 2 * q) * astr] = sa;
        }
      }
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        rn.e[e] = rn2[e];
#pragma unroll
        for (int q = 0; q < NA; ++q) an[q].e[e] = an2[e][q];
      }
    }
  }
}

// block-reduce the per-thread sums of a pass into the CTA partial rows (slots 0..15, 20..35, 36 + 16 sw ..)
template <int ND, int NS, bool XTRA>
__device__ __forceinline__ void dual_publish(Ctx& c, const DualSums<2 * ND>& S, int nx_sw, const double* xacc, int astr) {
  constexpr int NA = 2 * ND, kAcc = DualAccLayout<NA>::kAcc;
  double sums[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) sums[q] = 0.0;
  sums[0] = S.s_dr; sums[1] = S.s_rr; sums[15] = S.s_nan;
#pragma unroll
  for (int q = 0; q < NA; ++q) { sums[2 + 2 * q] = S.s_da[q]; sums[3 + 2 * q] = S.s_aa[q]; }
  cta_partials<16>(c, sums, 0);
  if (NS == 2) {
#pragma unroll
    for (int q = 0; q < 16; ++q) sums[q] = 0.0;
    sums[0] = S.t_dr; sums[1] = S.t_rr; sums[15] = S.t_nan;
#pragma unroll
    for (int q = 0; q < NA; ++q) { sums[2 + 2 * q] = S.t_da[q]; sums[3 + 2 * q] = S.t_aa[q]; }
    cta_partials<16>(c, sums, 20);
    for (int sw = 0; sw < (XTRA ? nx_sw : 0); ++sw) {
      const double* as = xacc + sw * kAcc * astr;
#pragma unroll
      for (int q = 0; q < 16; ++q) sums[q] = 0.0;
      sums[0] = as[0]; sums[1] = as[astr]; sums[15] = as[2 * astr];
#pragma unroll
      for (int q = 0; q < NA; ++q) { sums[2 + 2 * q] = as[(3 + 2 * q) * astr]; sums[3 + 2 * q] = as[(4 + 2 * q) * astr]; }
      cta_partials<16>(c, sums, 36 + 16 * sw);
    }
  }
}

// direct-load variant: thread-linear grid-stride over VW-wide items, operands global -> registers (any grid, 1-D and 2-D)
template <int ND, int VW, int EG, int NS = 1, bool XTRA = false>
__device__ __noinline__ void phase_D(Ctx& c, const double* phib, const double* rho_s, const double* alp_s, double* rho_d, double* alp_d,
                                     double sigma, double epsl, int nx_sw = 0) {
  phib = as_global(phib); rho_s = as_global(rho_s); alp_s = as_global(alp_s); rho_d = as_global(rho_d); alp_d = as_global(alp_d);
  constexpr int NA = 2 * ND;
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const Recip rc(a.r_idt, a.r_idx, a.r_idy, a.r_idx2, a.r_idy2, 1.0 / sigma);
  const int ny2 = ny / VW;
  DualSums<NA> S;
  S.clear();
  constexpr int kAcc = DualAccLayout<NA>::kAcc;
  double* xacc = reinterpret_cast<double*>(c.work()) + threadIdx.x;
  const int astr = blockDim.x;
  if (XTRA) {
    for (int q = 0; q < nx_sw * kAcc; ++q) xacc[q * astr] = 0.0;
  }
  const float inv_ny2 = 1.0f / (float)ny2;
  const long long items = (long long)K * nx * ny2, istride = (long long)gridDim.x * blockDim.x;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < items; item += istride) {
    const int r = (items < (1LL << 24)) ? fast_div_exact((int)item, ny2, inv_ny2) : (int)(item / ny2);
    const int jp = (int)(item - (long long)r * ny2);
    const int k = fast_div_exact(r, nx, cargs().inv_nx), i = r - k * nx, j = jp * VW;
    const size_t row = (size_t)i * ny, g = (size_t)k * n + row + j;
    const double* pb1 = phib + (size_t)(k + 1) * n;
    const int jm = (j == 0) ? ny - 1 : j - 1, jq = (j + VW == ny) ? 0 : j + VW;
    const Vec<VW> cc = ldv<VW>(pb1 + row + j);
    const Vec<VW> pk = ldv<VW>(phib + g);
    const Vec<VW> ro = ldv<VW>(rho_s + g);
    Vec<VW> ao[NA];
#pragma unroll
    for (int q = 0; q < NA; ++q) ao[q] = ldv<VW>(alp_s + (size_t)q * KN + g);
    Vec<VW> cxm, cxp;
    const Nbr bx = nbr(i, nx, p.bc_x);
    if (ND == 2) { cxm = ldv<VW>(pb1 + (size_t)bx.m * ny + j); cxp = ldv<VW>(pb1 + (size_t)bx.p * ny + j); }
    const double c_l = ldg1(pb1 + row + jm), c_r = ldg1(pb1 + row + jq);
    const double cx = (ND == 2 || EG == 3) ? c.cx()[i] : 0.0;
    const bool acc_on = (i >= a.sum_lo && i < a.sum_hi);
    Vec<VW> rn, an[NA];
    dual_item<ND, VW, EG, NS, XTRA>(c, cc, pk, ro, ao, cxm, cxp, c_l, c_r, cx, j, bx.wm, bx.wp, acc_on, sigma, epsl, rc, nx_sw, xacc, astr, S,
                                    rn, an);
    stv<VW>(rho_d + g, rn);
#pragma unroll
    for (int q = 0; q < NA; ++q) stv<VW>(alp_d + (size_t)q * KN + g, an[q]);
  }
  c.tick(5);
  dual_publish<ND, NS, XTRA>(c, S, nx_sw, xacc, astr);
}

// ---- TMA row pipeline (2-D grids with an even number of columns) -----------------------------------------------------
// A tile = the R x-rows [i0, i0 + R) of one time row k.  ONE thread of the CTA streams the tile's operand rows — phi_bar[k+1]
// rows i0-1 .. i0+R (the x-halo comes with it), phi_bar[k], rho and the 4 control arrays: (7 R + 2) rows of ny doubles — into
// an S-stage shared-memory ring with bulk asynchronous copies (cp.async.bulk, one instruction per array), S - 1 tiles ahead
// of the consumers.  All 16 warps consume: wait on the stage's `full` mbarrier (transaction bytes), read their operands with
// LDS.128 at immediate offsets (no global address arithmetic, no load registers held across the wait, every phi_bar row is
// fetched from L2 once per tile instead of three times), compute, store the results straight from registers (coalesced
// 16-byte STG), and release the stage on its `empty` mbarrier.  Bytes in flight per SM are set by the ring (>= 64 KB), not
// by occupancy.  Per-point arithmetic is dual_item's, i.e. the same as the direct-load variant's, bit for bit.

template <int EG, int NS, bool XTRA>
__device__ __noinline__ void phase_D_tma(Ctx& c, const double* phib, const double* rho_s, const double* alp_s, double* rho_d,
                                         double* alp_d, double sigma, double epsl, int nx_sw = 0) {
  phib = as_global(phib); rho_s = as_global(rho_s); alp_s = as_global(alp_s); rho_d = as_global(rho_d); alp_d = as_global(alp_d);
  constexpr int ND = 2, NA = 4, VW = 2;
  const CoopArgs& a = cargs();
  const MarchParams& p = a.p;
  const int K = p.K, nx = a.nxe, ny = a.nye;
  const size_t n = (size_t)nx * ny, KN = (size_t)K * n;
  const Recip rc(a.r_idt, a.r_idx, a.r_idy, a.r_idx2, a.r_idy2, 1.0 / sigma);
  const int ny2 = ny >> 1;
  const int R = a.tma_R;
  constexpr int kAcc = DualAccLayout<NA>::kAcc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int astr = blockDim.x;
  double* xacc = reinterpret_cast<double*>(c.work()) + tid;
  if (XTRA) {
    for (int q = 0; q < nx_sw * kAcc; ++q) xacc[q * astr] = 0.0;
  }
  // ring geometry: stages start after the accumulator slots of the extra sweeps
  const int xbytes = XTRA ? ((nx_sw * kAcc * astr * 8 + 127) & ~127) : 0;
  const int stage_d = (7 * R + 2) * ny;                                   // doubles per stage
  int S = (a.work_bytes - xbytes) / (stage_d * 8);
  if (S > a.tma_S) S = a.tma_S;
  double* ring = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(c.work()) + xbytes);
  uint64_t* full = c.bars();
  uint64_t* empty = full + kPipeMaxStages;
  const int ntx = (nx + R - 1) / R;                                       // tiles per time row
  const int T = K * ntx, G = gridDim.x;
  const float inv_ntx = 1.0f / (float)ntx, inv_ny2 = 1.0f / (float)ny2;
  DualSums<NA> Sm;
  Sm.clear();
  // the ring overlays buffers the previous phase wrote through the generic proxy: order those accesses before the async-proxy
  // writes of the bulk copies
  fence_proxy_async_smem();
  __syncthreads();

  // producer (warp 0): lane q issues copy q of the tile; lane 0 arms the barrier with the tile's byte count
  auto produce = [&](int tile, int s) {
    const int k = fast_div_exact(tile, ntx, inv_ntx), i0 = (tile - k * ntx) * R, nr = min(R, nx - i0);
    mbar_wait(&empty[s], ((c.pipe.par_empty >> s) & 1u) ^ 1u);            // every consumer warp has released the stage
    c.pipe.par_empty ^= (1u << s);
    double* st = ring + (size_t)s * stage_d;
    const uint32_t rowb = (uint32_t)ny * 8u;
    if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)((nr + 2) + nr * (2 + NA)) * rowb);
    __syncwarp();
    const double* pb1 = phib + (size_t)(k + 1) * n;
    const Nbr lo = nbr(i0, nx, p.bc_x), hi = nbr(i0 + nr - 1, nx, p.bc_x);
    const size_t g0 = (size_t)k * n + (size_t)i0 * ny;
    if (lane == 0) bulk_g2s(st, pb1 + (size_t)lo.m * ny, rowb, &full[s]);
    else if (lane == 1) bulk_g2s(st + ny, pb1 + (size_t)i0 * ny, rowb * nr, &full[s]);
    else if (lane == 2) bulk_g2s(st + (size_t)(nr + 1) * ny, pb1 + (size_t)hi.p * ny, rowb, &full[s]);
    else if (lane == 3) bulk_g2s(st + (size_t)(R + 2) * ny, phib + g0, rowb * nr, &full[s]);
    else if (lane == 4) bulk_g2s(st + (size_t)(2 * R + 2) * ny, rho_s + g0, rowb * nr, &full[s]);
    else if (lane < 5 + NA) bulk_g2s(st + (size_t)(3 * R + 2 + (lane - 5) * R) * ny, alp_s + (size_t)(lane - 5) * KN + g0, rowb * nr, &full[s]);
  };

  int ps = 0;                                                             // producer stage cursor
  if (warp == 0) {
    for (int m = 0; m < S - 1; ++m) {
      const int tm = blockIdx.x + m * G;
      if (tm < T) produce(tm, ps);
      if (++ps == S) ps = 0;
    }
  }
  int cs = 0;                                                             // consumer stage cursor
  for (int tile = blockIdx.x; tile < T; tile += G) {
    if (warp == 0) {
      const int tm = tile + (S - 1) * G;
      if (tm < T) produce(tm, ps);
      if (++ps == S) ps = 0;
    }
    const int k = fast_div_exact(tile, ntx, inv_ntx), i0 = (tile - k * ntx) * R, nr = min(R, nx - i0);
    mbar_wait(&full[cs], (c.pipe.par_full >> cs) & 1u);
    c.pipe.par_full ^= (1u << cs);
    const double* st = ring + (size_t)cs * stage_d;
    const int nitems = nr * ny2;
    for (int it = tid; it < nitems; it += astr) {
      const int r = fast_div_exact(it, ny2, inv_ny2), jp = it - r * ny2, j = 2 * jp, i = i0 + r;
      const double* prow = st + (size_t)(r + 1) * ny;                     // phi_bar[k+1], row i
      const int jm = (j == 0) ? ny - 1 : j - 1, jq = (j + 2 == ny) ? 0 : j + 2;
      const double2 vcc = *reinterpret_cast<const double2*>(prow + j);
      const double2 vxm = *reinterpret_cast<const double2*>(prow - ny + j);
      const double2 vxp = *reinterpret_cast<const double2*>(prow + ny + j);
      const double c_l = prow[jm], c_r = prow[jq];
      const double* own = st + (size_t)(R + 2 + r) * ny + j;              // phi_bar[k], then rho, alp q at strides of R rows
      const double2 vpk = *reinterpret_cast<const double2*>(own);
      const double2 vro = *reinterpret_cast<const double2*>(own + (size_t)R * ny);
      Vec<2> cc, pk, ro, cxm, cxp, ao[NA];
      cc.e[0] = vcc.x; cc.e[1] = vcc.y; pk.e[0] = vpk.x; pk.e[1] = vpk.y; ro.e[0] = vro.x; ro.e[1] = vro.y;
      cxm.e[0] = vxm.x; cxm.e[1] = vxm.y; cxp.e[0] = vxp.x; cxp.e[1] = vxp.y;
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(own + (size_t)(2 + q) * R * ny);
        ao[q].e[0] = v.x; ao[q].e[1] = v.y;
      }
      const Nbr bx = nbr(i, nx, p.bc_x);
      const double cx = c.cx()[i];
      const bool acc_on = (i >= a.sum_lo && i < a.sum_hi);
      Vec<2> rn, an[NA];
      dual_item<ND, VW, EG, NS, XTRA>(c, cc, pk, ro, ao, cxm, cxp, c_l, c_r, cx, j, bx.wm, bx.wp, acc_on, sigma, epsl, rc, nx_sw, xacc, astr, Sm,
                                      rn, an);
      const size_t g = (size_t)k * n + (size_t)i * ny + j;
      stv<2>(rho_d + g, rn);
#pragma unroll
      for (int q = 0; q < NA; ++q) stv<2>(alp_d + (size_t)q * KN + g, an[q]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[cs]);                               // this warp is done reading the stage
    if (++cs == S) cs = 0;
  }
  c.tick(5);
  dual_publish<ND, NS, XTRA>(c, Sm, nx_sw, xacc, astr);
}

// ---- phase E (only after an inner loop of more than one sweep): outer-iteration differences of the dual variables,
// sum (x_next - x_prev)^2, for err2 (utils_pdhg_solver.py:61-68).  Reading the previous iterate once here instead of
// in every extra sweep keeps the extra sweeps at the traffic of the first one.  CTA partials: slot 10 rho, 11+q alp q.
template <int VW>
__device__ __noinline__ void phase_E(Ctx& c, const double* rho_n, const double* alp_n, const double* rho_o, const double* alp_o) {
  rho_n = as_global(rho_n); alp_n = as_global(alp_n); rho_o = as_global(rho_o); alp_o = as_global(alp_o);
  const CoopArgs& a = cargs();
  const size_t KN = (size_t)a.p.K * a.nxe * a.nye;
  const int A = a.A;
  const size_t stride = (size_t)gridDim.x * blockDim.x * VW;
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int ny_ = a.nye, nx_ = a.nxe;
  for (size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VW; g < KN; g += stride) {
    const int i_ = (int)((g / ny_) % nx_);
    if (i_ < a.sum_lo || i_ >= a.sum_hi) continue;
    {
      const Vec<VW> x = ldv<VW>(rho_n + g), y = ldv<VW>(rho_o + g);
#pragma unroll
      for (int e = 0; e < VW; ++e) { const double d = x.e[e] - y.e[e]; s[0] += d * d; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < A) {
        const Vec<VW> x = ldv<VW>(alp_n + (size_t)q * KN + g), y = ldv<VW>(alp_o + (size_t)q * KN + g);
#pragma unroll
        for (int e = 0; e < VW; ++e) { const double d = x.e[e] - y.e[e]; s[1 + q] += d * d; }
      }
    }
  }
  cta_partials<5>(c, s, 10);
}

// runtime -> compile-time dispatch of the templated phases
template <int EG>
__device__ __forceinline__ void run_A_eg(Ctx& c, int cd, double epsl) {
  const bool v2 = (cargs().nye & 1) == 0;
  if (cargs().fast_y) { if (cargs().has_x) phase_A_w256<2, EG>(c, cd, epsl); else phase_A_w256<1, EG>(c, cd, epsl); return; }
  if (cargs().has_x) { if (v2) phase_A<2, 2, EG>(c, cd, epsl); else phase_A<2, 1, EG>(c, cd, epsl); }
  else { if (v2) phase_A<1, 2, EG>(c, cd, epsl); else phase_A<1, 1, EG>(c, cd, epsl); }
}
__device__ __forceinline__ void run_A(Ctx& c, int cd, double epsl) {
  // the continuity residual only distinguishes the Newton example (egno 3: f = [alp, x]) from the a(x)-scaled ones
  if (cargs().p.egno == 3) run_A_eg<3>(c, cd, epsl); else run_A_eg<1>(c, cd, epsl);
}
__device__ __forceinline__ void run_C(Ctx& c, const double* pp, double* pn, double* pb, double tau) {
  if (cargs().fast_y) { phase_C_w256(c, pp, pn, pb, tau); return; }
  if ((cargs().nye & 1) == 0) phase_C<2>(c, pp, pn, pb, tau); else phase_C<1>(c, pp, pn, pb, tau);
}
template <int EG>
__device__ __forceinline__ void run_D_eg(Ctx& c, const double* phib, const double* rho_s, const double* alp_s, double* rho_d, double* alp_d,
                                         double sigma, double epsl, int ns) {
  const bool v2 = (cargs().nye & 1) == 0;
  if (cargs().tma_d) {      // 2-D, even ny (checked on the host): TMA row pipeline
    if (ns == 1) phase_D_tma<EG, 1, false>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl);
    else if (ns == 2) phase_D_tma<EG, 2, false>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl);
    else phase_D_tma<EG, 2, true>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns - 2);
    return;
  }
  if (ns == 2) {        // fused sweeps
    if (cargs().has_x) { if (v2) phase_D<2, 2, EG, 2>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl);
                         else phase_D<2, 1, EG, 2>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl); }
    else { if (v2) phase_D<1, 2, EG, 2>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl);
           else phase_D<1, 1, EG, 2>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl); }
    return;
  }
  if (ns > 2) {
    if (cargs().has_x) { if (v2) phase_D<2, 2, EG, 2, true>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns - 2);
                         else phase_D<2, 1, EG, 2, true>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns - 2); }
    else { if (v2) phase_D<1, 2, EG, 2, true>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns - 2);
           else phase_D<1, 1, EG, 2, true>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns - 2); }
    return;
  }
  if (cargs().has_x) { if (v2) phase_D<2, 2, EG>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl);
                       else phase_D<2, 1, EG>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl); }
  else { if (v2) phase_D<1, 2, EG>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl);
         else phase_D<1, 1, EG>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl); }
}
__device__ __forceinline__ void run_D(Ctx& c, const double* phib, const double* rho_s, const double* alp_s, double* rho_d, double* alp_d,
                                      double sigma, double epsl, int ns = 1) {
  const int eg = cargs().p.egno;
  if (eg == 1) run_D_eg<1>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns);
  else if (eg == 2) run_D_eg<2>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns);
  else run_D_eg<3>(c, phib, rho_s, alp_s, rho_d, alp_d, sigma, epsl, ns);
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ void grid_copy(double* dst, const double* src, size_t count) {
  dst = as_global(dst); src = as_global(src);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < count; g += stride) stg1(dst + g, ldg1(src + g));
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
      const double rden = 1.0 / den;                   // reciprocal pivot
      // modified super-diagonal du / pivot through the reciprocal (<= 1 ulp from the division of utils_precond.py:21): the
      // single-pass phase B then needs only the reciprocal-pivot table and forms this product on the fly
      tprev = (k == K - 1) ? 0.0 : -ct2 * rden;
      a.w.den[(size_t)k * modes + m] = rden;
      a.w.tu[(size_t)k * modes + m] = tprev;
    }
  }
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm) pdhg_coop_kernel() {
  const CoopArgs& a = cargs();
  Ctx c;
  double* red = c.red();
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

  if (a.mode == MODE_PHASE) {
    // profiling aid: ONE phase on whatever the workspace holds (left there by a previous march), so that ncu sees each
    // phase as its own launch.  Results are not meaningful; nothing outside the workspace is written.
    double v[kNV];
    switch (a.dbg_phase) {
      case 5: return;                          // diagnostic: the fixed cost of a launch (argument block, context set-up, exit)
      case 0: run_A(c, 0, epsl); break;
      case 1: phase_B(c); break;
      case 2: run_C(c, w.phi[0], w.phi[1], w.phib, a.op_step); break;
      case 3: {
        // dbg_pass = number of sweeps to fuse in this pass (slab mode; 2 needs no shared memory, more only up to d_fuse)
        const int ns = (a.dbg_pass == 2 || (a.dbg_pass >= 3 && a.dbg_pass <= a.d_fuse)) ? a.dbg_pass : 1;
        run_D(c, w.phib, w.rho[0], w.alp[0], w.rho[1], w.alp[1], a.op_step, epsl, ns);
        grid_gather(c, v, 20 + 16 * (ns - 1));
        if (a.ext_sums && lead) {
          for (int q = kNV; q < 20 + 16 * (ns - 1); ++q) a.ext_sums[q] = gathered(c, q);
        }
        break;
      }
      default:
        if ((a.nye & 1) == 0) phase_E<2>(c, w.rho[1], w.alp[1], w.rho[0], w.alp[0]); else phase_E<1>(c, w.rho[1], w.alp[1], w.rho[0], w.alp[0]);
        grid_gather(c, v);
        break;
    }
    // grid totals of this launch's reduction (incl. the partials a preceding phase-C launch left in slots 16..18) for the
    // host: slab mode all-reduces them across ranks and takes the decisions there
    if (a.ext_sums && a.dbg_phase >= 3 && lead) {
      for (int q = 0; q < kNV; ++q) a.ext_sums[q] = v[q];
    }
    return;
  }

  if (a.mode == MODE_PRIMAL) {
    // phi_next = phi_prev + tau * H1_precond(cont_residual(rho, alp))   (update_fns_in_pdhg.py:135-147)
    grid_copy(w.rho[0], grho, KN);
    grid_copy(w.alp[0], galp, (size_t)A * KN);
    c.grid.sync();
    run_A(c, 0, epsl);
    c.grid.sync();
    phase_B(c);
    c.grid.sync();
    run_C(c, a.op_phi_in, a.op_phi_out, nullptr, a.op_step);
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
      double v[kNV];
      const int s = j & 1, d = s ^ 1;
      run_D(c, a.op_phi_in, w.rho[s], w.alp[s], w.rho[d], w.alp[d], a.op_step, epsl);
      grid_gather(c, v);
      err = v[0] / v[1];
#pragma unroll
      for (int q = 0; q < 4; ++q) if (q < A) err += v[2 + 2 * q] / v[3 + 2 * q];
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
  bool log_overflow = false;     // sticky: more periodic records than max_rec - 1 in some block (rows dropped, march goes on)
  int blocks_done = p.blk_begin;
  long long inner_total = 0;
  const int nt_all = p.nblocks * K + 1;
  unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = gtimer();
#define TICK(slot) do { if (lead) { const unsigned long long t_ = gtimer(); tacc[slot] += t_ - tlast; tlast = t_; c.tl = t_; } } while (0)

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
        double s6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const size_t stride = (size_t)gridDim.x * blockDim.x, g0 = (size_t)blockIdx.x * blockDim.x + tid;
        for (size_t g = g0; g < n; g += stride) { const double x = gphi[g]; s6[0] += x * x; }
        for (size_t g = g0; g < KN; g += stride) {
          const double r = grho[g]; s6[1] += r * r;
#pragma unroll
          for (int q = 0; q < 4; ++q) if (q < A) { const double x = galp[(size_t)q * KN + g]; s6[2 + q] += x * x; }
        }
        cta_partials<6>(c, s6, 0);
        double v[kNV];
        grid_gather(c, v);    // (its grid sync also orders the copies above before phase A)
        S_row0 = v[0]; S_rho = v[1];
        for (int q = 0; q < 4; ++q) S_alp[q] = v[2 + q];
      }
      long long it = p.iter_begin;
      int reason = END_MAXITER, nrec = 0, prev_j = p.rho_alp_iters;
      bool lognan = false, logfull = false;
      double err1 = 0.0, err2 = 0.0, rmin = 0.0, rmax = 0.0;

      for (; it < p.n_maxiter; ++it) {
        if (it >= p.iter_pause) { reason = END_PAUSED; break; }
        TICK(5);
        run_A(c, cd, epsl);
        c.grid.sync();
        TICK(0);
        phase_B(c);
        c.grid.sync();
        TICK(1);
        double v[kNV];
        run_C(c, w.phi[cp], w.phi[cp ^ 1], w.phib, tau);
        c.grid.sync();
        TICK(2);
        // dual sweeps (update_dual_alternative, update_fns_in_pdhg.py:167-180): buffer cd holds the outer iterate and stays intact
        // (phase E needs it), the sweeps ping-pong between the two other buffers.  While the previous outer iteration needed
        // several sweeps, two sweeps are fused per pass (phase_D<NS = 2>); when the exit test already holds after the first
        // sweep of a pair, that single sweep is redone from the pair's (still intact) input, so iterates, sweep counts and
        // error sums are exactly those of the sweep-by-sweep loop.
        const int f1 = (cd + 1) % 3, f2 = (cd + 2) % 3;
        double e1s0 = 0.0, e1s1 = 0.0, e1nan = 0.0;
        int j = 0, last = cd;
        while (j < p.rho_alp_iters) {
          const int src = last, dst = (last == f1) ? f2 : f1;
          int ns = min(min(prev_j - j, p.rho_alp_iters - j), a.d_fuse);
          if (ns < 1) ns = 1;
          run_D(c, w.phib, w.rho[src], w.alp[src], w.rho[dst], w.alp[dst], sigma, epsl, ns);
          grid_gather(c, v, 20 + 16 * (ns - 1));
          c.tick(9);
          if (j == 0) { e1s0 = v[16]; e1s1 = v[17]; e1nan = v[18]; }
          // exit test after each of the ns sweeps of this pass, in order (sweep 0: v[0..15]; the others: shared memory)
          int hit = -1;
          {
            double err = v[0] / v[1];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < A) err += v[2 + 2 * q] / v[3 + 2 * q];
            if (err < p.eps) hit = 0;
          }
          for (int sw = 1; sw < ns && hit < 0; ++sw) {
            const int b0 = 20 + 16 * (sw - 1);
            double err = gathered(c, b0) / gathered(c, b0 + 1);
            for (int q = 0; q < 4; ++q) if (q < A) err += gathered(c, b0 + 2 + 2 * q) / gathered(c, b0 + 3 + 2 * q);
            if (err < p.eps) hit = sw;
          }
          int done_sw = ns;                     // sweeps of this pass that count
          if (hit >= 0 && hit < ns - 1) {
            // exit inside the fused pass: redo exactly hit + 1 sweeps from the same (intact) input
            done_sw = hit + 1;
            run_D(c, w.phib, w.rho[src], w.alp[src], w.rho[dst], w.alp[dst], sigma, epsl, done_sw);
            grid_gather(c, v, 20 + 16 * (done_sw - 1));
          }
          // v[0..15] <- sums of the last sweep done
          if (done_sw >= 2) {
            const int b0 = 20 + 16 * (done_sw - 2);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = gathered(c, b0 + q);
          }
          last = dst; j += done_sw;
          if (hit >= 0) break;
        }
        prev_j = j;
        const int nd = last;
        inner_total += j;
        const bool multi = (j > 1);
        double vo[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (multi) {
          if ((a.nye & 1) == 0) phase_E<2>(c, w.rho[nd], w.alp[nd], w.rho[cd], w.alp[cd]);
          else phase_E<1>(c, w.rho[nd], w.alp[nd], w.rho[cd], w.alp[cd]);
          double v2[kNV];
          grid_gather(c, v2);
#pragma unroll
          for (int q = 0; q < 5; ++q) vo[q] = v2[10 + q];
        }
        TICK(3);
        err1 = sqrt(e1s0) / sqrt(S_row0 + e1s1);
        err2 = sqrt(multi ? vo[0] : v[0]) / sqrt(S_rho);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < A) {
            const double na = sqrt(S_alp[q]), ne = sqrt(multi ? vo[1 + q] : v[2 + 2 * q]);
            if (na < 1e-6 && ne > 1e-6) err2 += ne; else if (na >= 1e-6) err2 += ne / na;
          }
        }
        S_rho = v[1];
#pragma unroll
        for (int q = 0; q < 4; ++q) S_alp[q] = v[3 + 2 * q];
        const bool anynan = (e1nan > 0.0) || (v[15] > 0.0);
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
          if ((tid & 31) == 0) { red[tid >> 5] = mn; red[kWarps + (tid >> 5)] = mx; }
          __syncthreads();
          double* part = w.partials + (size_t)(c.epoch & 1) * gridDim.x * kNQ;
          if (tid == 0) {
            for (int q = 1; q < (int)(blockDim.x >> 5); ++q) { mn = fmin(mn, red[q]); mx = fmax(mx, red[kWarps + q]); }
            part[blockIdx.x] = mn; part[(size_t)gridDim.x + blockIdx.x] = mx;
          }
          c.grid.sync();
          mn = 1e300; mx = -1e300;
          for (int g = 0; g < (int)gridDim.x; ++g) {
            mn = fmin(mn, *((volatile double*)&part[g])); mx = fmax(mx, *((volatile double*)&part[(size_t)gridDim.x + g]));
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
      log_overflow = log_overflow || logfull;
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
  TICK(5);
  if (lead) {
    p.status[b] = (status == ST_OK && log_overflow) ? ST_LOG_OVERFLOW : status;   // overflow never stops the march (only NaN failure and pause do)
    p.blocks_done[b] = blocks_done;
    p.stepsz[b] = stepsz;
    p.inner_total[b] = inner_total;
    for (int q = 0; q < 6; ++q) w.phase_ns[q] = (double)tacc[q];
    for (int q = 0; q < 10; ++q) w.phase_ns[6 + q] = (double)c.tsub[q];
  }
#undef TICK
}

// ------------------------------------------- host side -------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kMaxDevices = 16;
struct ArgsSlot { cudaEvent_t ev = nullptr; cudaStream_t stream = nullptr; bool used = false; };
static ArgsSlot g_args_slot[kMaxDevices];
static std::mutex g_args_mutex;

struct CoopGeom { int nxe, nye, nyh, TR, TKY, grid, d_fuse, d_fuse_ext, fast_y, fast_x, tma_d, tma_R, tma_S, work_bytes; size_t smem; };

static CoopGeom coop_geom(const MarchParams& p, int sm_count, size_t smem_cap) {
  CoopGeom g;
  const Knobs& kn = p.knobs;
  g.nxe = (p.ndim == 1) ? 1 : p.nx;
  g.nye = (p.ndim == 1) ? p.nx : p.ny;
  g.nyh = g.nye / 2 + 1;
  const int rows = p.K * g.nxe;
  // y-FFT tile: TR rows (even) -> TR/2 complex transforms, two buffers of (TR/2)*(nye+1) complex
  const size_t tab = (size_t)kArgsBytes + kRedBytes + kBarBytes + (size_t)24 * (g.nxe + g.nye) + 32;   // args, reduction scratch, barriers, twiddles + coefficient tables
  const size_t cap = smem_cap > tab ? smem_cap - tab : 0;
  int TR = 16;
  while (TR > 2 && ((rows + TR - 1) / TR < 2 * sm_count || (size_t)TR * fft_ld(g.nye) * 16 > cap)) TR -= 2;
  g.TR = TR;
  int TKY = 8;
  while (TKY > 1 && ((size_t)p.K * ((g.nyh + TKY - 1) / TKY) < (size_t)2 * sm_count || (size_t)2 * TKY * fft_ld(g.nxe) * 16 > cap)) TKY -= 1;
  g.TKY = TKY;
  const size_t smA = (size_t)TR * fft_ld(g.nye) * 16;                 // 2 buffers * TR/2 rows
  const size_t smB = (g.nxe > 1) ? (size_t)2 * TKY * fft_ld(g.nxe) * 16 : 0;
  size_t work = smA > smB ? smA : smB;
  // warp-private 256-point transforms: two padded rows per warp
  const size_t smW = (size_t)kWarps * 2 * kW256Ld * 16;
  const bool w256 = !kn.no_w256 && smW <= cap;
  // (a unit is one warp's work; below ~8 units per SM the tiled path, which spreads one tile over a whole CTA, is faster)
  const bool force = kn.force_w256 != 0;
  const bool many_y = force || (rows + 3) / 4 >= 8 * sm_count, many_x = force || (p.K * g.nyh + 1) / 2 >= 8 * sm_count;
  g.fast_y = (w256 && many_y && g.nye == 256 && (p.ndim == 1 ? p.bc_x == 0 : p.bc_y == 0)) ? 1 : 0;
  g.fast_x = (w256 && many_x && g.nxe == 256 && p.ndim == 2 && p.bc_x == 0) ? 1 : 0;
  if ((g.fast_y || g.fast_x) && smW > work) work = smW;
  // (L2-resident blocks are barrier / latency bound: fused passes and bulk-copy pipelines save them nothing)
  const bool hbm_bound = (size_t)p.K * g.nxe * g.nye * 8 * (4 + 4 * p.ndim) > ((size_t)64 << 20);
  // dual sweep through the TMA row pipeline: 2-D grids with even ny whose ring of >= 2 stages of (7 R + 2) rows fits
  g.tma_d = 0; g.tma_R = 0; g.tma_S = 0;
  size_t stage = 0;
  if (p.ndim == 2 && (g.nye & 1) == 0 && kn.tma != 0) {
    int R = (2 * kThreads + g.nye - 1) / g.nye;                       // rows per tile: every thread gets an item
    if (R > 8) R = 8;
    if (R > g.nxe) R = g.nxe;
    if (R < 1) R = 1;
    stage = (size_t)(7 * R + 2) * g.nye * 8;
    int S = (int)(cap / stage);
    if (S > 4) S = 4;
    if (S >= 2 && kn.tma == 1) { g.tma_d = 1; g.tma_R = R; g.tma_S = S; }   // opt-in: see DESIGN.md (direct loads already saturate HBM)
  }
  // fused dual sweeps: two per pass on the HBM-bound grids; more (up to kFuseMax) while the per-thread accumulator slots of
  // sweeps 3.. (3 + 4 ndim sums each) fit the work area (next to two ring stages when the TMA pipeline runs)
  g.d_fuse = !hbm_bound ? 1 : kFuseMax;
  if (kn.dfuse > 0) g.d_fuse = kn.dfuse;
  if (g.d_fuse > kFuseMax) g.d_fuse = kFuseMax;
  if (g.d_fuse < 1) g.d_fuse = 1;
  auto xacc_bytes = [&](int df) { return df > 2 ? (((size_t)(df - 2) * (3 + 4 * p.ndim) * kThreads * 8 + 127) & ~(size_t)127) : (size_t)0; };
  if (g.tma_d) {
    while (g.d_fuse > 2 && xacc_bytes(g.d_fuse) + 2 * stage > cap) --g.d_fuse;
    const size_t ring = xacc_bytes(g.d_fuse) + 2 * stage > (size_t)g.tma_S * stage ? xacc_bytes(g.d_fuse) + 2 * stage : (size_t)g.tma_S * stage;
    if (ring > work) work = ring;
  } else {
    while (g.d_fuse > 2 && xacc_bytes(g.d_fuse) > (work > cap ? cap : work)) --g.d_fuse;
  }
  // slab mode (MODE_PHASE, host-driven passes): fusing saves a launch and an all-reduce per sweep whatever the grid size, so the
  // cap only depends on what the work area holds
  g.d_fuse_ext = kFuseMax;
  if (kn.dfuse > 0 && kn.dfuse < g.d_fuse_ext) g.d_fuse_ext = kn.dfuse;
  while (g.d_fuse_ext > 2 && xacc_bytes(g.d_fuse_ext) + (g.tma_d ? 2 * stage : 0) > (work > cap ? cap : work)) --g.d_fuse_ext;
  g.work_bytes = (int)work;
  g.smem = tab + work;
  g.grid = sm_count * kCtasPerSm;
  return g;
}

size_t pdhg_coop_workspace_bytes(const MarchParams& p, int B) {
  (void)B;
  const int nxe = (p.ndim == 1) ? 1 : p.nx, nye = (p.ndim == 1) ? p.nx : p.ny, nyh = nye / 2 + 1, A = 2 * p.ndim;
  const size_t n = (size_t)nxe * nye, KN = (size_t)p.K * n, NP = (size_t)(p.K + 1) * n, modes = (size_t)nyh * nxe;
  size_t bytes = 0;
  bytes += align_up(3 * NP * 8, 256);
  bytes += align_up(3 * KN * 8, 256);
  bytes += align_up(3 * A * KN * 8, 256);
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
  w.rho[0] = (double*)q; w.rho[1] = w.rho[0] + KN; w.rho[2] = w.rho[1] + KN; q += align_up(3 * KN * 8, 256);
  w.alp[0] = (double*)q; w.alp[1] = w.alp[0] + (size_t)A * KN; w.alp[2] = w.alp[1] + (size_t)A * KN; q += align_up(3 * A * KN * 8, 256);
  w.zt = (double2*)q; q += align_up((size_t)p.K * modes * 16, 256);
  w.partials = (double*)q; q += align_up((size_t)2 * 1024 * kNQ * 8, 256);
  w.den = (double*)q; w.tu = w.den + (size_t)p.K * modes; q += align_up(2 * (size_t)p.K * modes * 8, 256);
  w.phase_ns = (double*)q;
  return w;
}

static void apply_ext(const ExtPhaseDesc& e, CoopArgs& a) {
  a.ext_sums = e.sums;
  if (e.phi_in) a.w.phi[0] = e.phi_in;
  if (e.phi_out) a.w.phi[1] = e.phi_out;
  if (e.phib) a.w.phib = e.phib;
  if (e.rho_in) a.w.rho[0] = e.rho_in;
  if (e.alp_in) a.w.alp[0] = e.alp_in;
  if (e.rho_out) a.w.rho[1] = e.rho_out;
  if (e.alp_out) a.w.alp[1] = e.alp_out;
  if (e.zt) a.w.zt = static_cast<double2*>(e.zt);
  a.sum_lo = e.sum_lo; a.sum_hi = e.sum_hi;
  if (e.nyh_override > 0) { a.nyh = e.nyh_override; a.ky_off = e.ky_off; a.nyh_tab = e.nyh_tab; a.b_slab = 0; }
  if (e.xch_P > 0 && (a.dbg_phase == 0 || a.dbg_phase == 1)) {
    a.xch_on = (a.dbg_phase == 0) ? 1 : 2;
    a.xch_rank = e.xch_rank; a.xch_nxl = e.xch_nxl; a.xch_kyl = e.xch_kyl; a.xch_nyh = e.xch_nyh;
    a.xch_ld = (a.dbg_phase == 0) ? e.xch_P * e.xch_nxl : e.xch_nxl + 2;
    a.xch_pull = (a.dbg_phase == 1) ? e.xch_pull : 0;
    for (int d = 0; d < 8; ++d) a.xch_ptr[d] = static_cast<double2*>(e.xch_ptr[d]);
  }
}

static cudaError_t coop_launch(const MarchParams& p, void* ws, int b, int mode, const double* op_in, double* op_out,
                               double op_step, double op_eps, int* op_ninner, double* op_err, cudaStream_t stream,
                               int dbg_phase = 0, int dbg_pass = 7, const ExtPhaseDesc* ext = nullptr) {
  int dev = 0, sms = 0, smem_cap = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const CoopGeom g = coop_geom(p, sms, ((size_t)smem_cap - 4096) / kCtasPerSm - (kCtasPerSm > 1 ? 2048 : 0));
  if (g.smem > (size_t)smem_cap - 2048) return cudaErrorInvalidConfiguration;
  CoopArgs a;
  memset(&a, 0, sizeof(a));
  a.p = p;
  a.w = carve(p, ws);
  a.b = b; a.mode = mode; a.A = 2 * p.ndim;
  a.nxe = g.nxe; a.nye = g.nye; a.nyh = g.nyh; a.TR = g.TR; a.TKY = g.TKY; a.fast_y = g.fast_y; a.fast_x = g.fast_x; a.d_fuse = g.d_fuse;
  a.b_slab = (g.fast_x && kThreads == 512 && p.K > 1 && !p.knobs.no_bslab) ? 1 : 0;
  a.tma_d = g.tma_d; a.tma_R = g.tma_R; a.tma_S = g.tma_S; a.work_bytes = g.work_bytes;
  a.has_x = (p.ndim == 2);
  if (p.ndim == 1) {
    a.dxe = 1.0; a.dye = p.dx; a.coef_xe = p.coef_x; a.coef_ye = p.coef_x; a.tw_xe = p.tw_x; a.tw_ye = p.tw_x;
    a.plan_ye = p.plan_x; a.plan_xe.n = 1; a.plan_xe.nstages = 0;
  } else {
    a.dxe = p.dx; a.dye = p.dy; a.coef_xe = p.coef_x; a.coef_ye = p.coef_y; a.tw_xe = p.tw_x; a.tw_ye = p.tw_y;
    a.plan_xe = p.plan_x; a.plan_ye = p.plan_y;
  }
  a.dbg_phase = dbg_phase; a.dbg_pass = dbg_pass;
  a.sum_lo = 0; a.sum_hi = g.nxe; a.ky_off = 0; a.nyh_tab = g.nyh;
  if (mode == MODE_PHASE) a.d_fuse = g.d_fuse_ext;
  if (ext) apply_ext(*ext, a);
  a.op_phi_in = op_in; a.op_phi_out = op_out; a.op_step = op_step; a.op_eps = op_eps; a.op_ninner = op_ninner; a.op_err = op_err;
  a.r_idt = 1.0 / p.dt; a.r_idx = 1.0 / a.dxe; a.r_idy = 1.0 / a.dye; a.r_idx2 = 1.0 / (a.dxe * a.dxe); a.r_idy2 = 1.0 / (a.dye * a.dye);
  a.inv_nx = 1.0f / (float)g.nxe;
  a.o_twy = kArgsBytes + kRedBytes + kBarBytes + 16 * g.nxe;
  a.o_cx = a.o_twy + 16 * g.nye;
  a.o_cy = a.o_cx + 8 * g.nxe;
  a.o_work = (a.o_cy + 8 * g.nye + 15) / 16 * 16;
  e = cudaFuncSetAttribute(pdhg_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  // The argument block goes to constant memory on the launch stream (ordered before the launch, after the previous kernel of
  // that stream).  The symbol is one per device: a launch from ANOTHER stream first waits (host side) for the last launch that
  // used it, so two handles driven from different streams serialise their launches instead of racing on it.
  {
    std::lock_guard<std::mutex> lock(g_args_mutex);
    ArgsSlot& slot = g_args_slot[dev & (kMaxDevices - 1)];
    if (slot.ev == nullptr) { e = cudaEventCreateWithFlags(&slot.ev, cudaEventDisableTiming); if (e != cudaSuccess) return e; slot.used = false; }
    if (slot.used && slot.stream != stream) { e = cudaEventSynchronize(slot.ev); if (e != cudaSuccess) return e; }
    e = cudaMemcpyToSymbolAsync(g_cargs, &a, sizeof(a), 0, cudaMemcpyHostToDevice, stream);   // (pageable source: staged before the call returns)
    if (e != cudaSuccess) return e;
    e = cudaLaunchCooperativeKernel((void*)pdhg_coop_kernel, dim3(g.grid), dim3(kThreads), nullptr, g.smem, stream);
    if (e != cudaSuccess) return e;
    slot.stream = stream; slot.used = true;
    return cudaEventRecord(slot.ev, stream);
  }
}

static cudaError_t ensure_tables(const MarchParams& p, void* ws, cudaStream_t stream, long long* launches) {
  // tables depend only on the handle's constants; rebuilt per call (cheap) so the workspace needs no extra state
  if (p.K == 1 || p.Ct_over_dt2 == 0.0) return cudaSuccess;
  ++*launches;
  return coop_launch(p, ws, 0, MODE_TABLES, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, stream);
}

cudaError_t launch_debug_phase(const MarchParams& p, void* ws, int phase, int pass_mask, double step, cudaStream_t stream) {
  return coop_launch(p, ws, 0, MODE_PHASE, nullptr, nullptr, step, 0.0, nullptr, nullptr, stream, phase, pass_mask);
}

cudaError_t launch_ext_phase(const MarchParams& p, void* ws, int phase, int pass_mask, double step, const ExtPhaseDesc& ext,
                             cudaStream_t stream) {
  return coop_launch(p, ws, 0, MODE_PHASE, nullptr, nullptr, step, 0.0, nullptr, nullptr, stream, phase, pass_mask, &ext);
}

// largest number of dual sweeps one pass may fuse for this problem (what MODE_PHASE / slab-mode callers may ask for)
int coop_max_fuse(const MarchParams& p) {
  int dev = 0, sms = 0, smem_cap = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return coop_geom(p, sms, ((size_t)smem_cap - 4096) / kCtasPerSm - (kCtasPerSm > 1 ? 2048 : 0)).d_fuse_ext;
}

// 1 if phases A and B of this problem run their generic (shared-memory tile) variants, the ones that can scatter their output
// into other ranks' buffers (fused transposes of the slab mode); the warp-private 256-point variants cannot
int coop_exchange_ok(const MarchParams& p) {
  int dev = 0, sms = 0, smem_cap = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const CoopGeom g = coop_geom(p, sms, ((size_t)smem_cap - 4096) / kCtasPerSm - (kCtasPerSm > 1 ? 2048 : 0));
  return (p.ndim == 2 && p.K == 1 && !g.fast_x && !g.fast_y && p.bc_x == 0) ? 1 : 0;
}

cudaError_t coop_phase_times(const MarchParams& p, void* ws, double* out6) {
  const CoopWs w = carve(p, ws);
  return cudaMemcpy(out6, w.phase_ns, 16 * sizeof(double), cudaMemcpyDeviceToHost);
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
