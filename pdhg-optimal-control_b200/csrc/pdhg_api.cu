// C-ABI layer of the B200-native PDHG hot path (see include/pdhg_b200.h for the contract).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/pdhg_b200.h"
#include "pdhg_params.h"

namespace pdhg {
// kernels / launchers implemented in the other translation units
cudaError_t launch_pdhg1d_cta(const MarchParams& p, int B, cudaStream_t stream);
cudaError_t launch_pdhg1d_k1(const MarchParams& p, int B, cudaStream_t stream);
bool pdhg1d_k1_supported(int nx, int K, int green_R);
size_t pdhg1d_cta_smem_bytes(int nx, int K);
cudaError_t launch_pdhg_coop(const MarchParams& p, int B, void* ws, cudaStream_t stream, long long* launches);
size_t pdhg_coop_workspace_bytes(const MarchParams& p, int B);
cudaError_t coop_phase_times(const MarchParams& p, void* ws, double* out6);
int coop_max_fuse(const MarchParams& p);
int coop_exchange_ok(const MarchParams& p);
cudaError_t launch_ext_phase(const MarchParams& p, void* ws, int phase, int pass_mask, double step, const ExtPhaseDesc& ext,
                             cudaStream_t stream);
cudaError_t launch_debug_phase(const MarchParams& p, void* ws, int phase, int pass_mask, double step, cudaStream_t stream);
cudaError_t launch_pack_alp(const double* ref_layout, double* planar, int B, int A, size_t kn, int n_ctrl, int ndim,
                            int egno, int to_planar, cudaStream_t stream);
cudaError_t launch_init_state(const MarchParams& p, const double* g, int B, cudaStream_t stream);
cudaError_t launch_update_primal(const MarchParams& p, int B, const double* phi_prev, double tau, double* phi_next,
                                 void* ws, cudaStream_t stream, long long* launches);
cudaError_t launch_update_dual(const MarchParams& p, int B, const double* phi_bar, double sigma, double eps,
                               int* n_inner_dev, double* err_dev, void* ws, cudaStream_t stream, long long* launches);
}  // namespace pdhg

using namespace pdhg;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// Every entry point runs on the handle's device and restores the caller's current device on every exit path.
struct DeviceGuard {
  int prev = -1;
  bool good = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) good = (cudaSetDevice(dev) == cudaSuccess); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  bool ok() const { return good; }
};

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? PDHG_ERR_NOMEM : PDHG_ERR_CUDA,                \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                             \
  } while (0)

struct pdhg_handle {
  pdhg_config cfg;
  int path = 0;
  int A = 0;            // active control arrays = 2*ndim
  size_t n = 0;         // nx*ny
  int nyh = 1;          // stored y-modes (ny/2+1 in 2-D)
  int B = 1;
  long long launches = 0;
  std::vector<void*> owned;
  double *coef_x = nullptr, *coef_y = nullptr, *diag = nullptr, *dct_cos = nullptr, *green = nullptr;
  int green_R = -1;
  double2 *tw_x = nullptr, *tw_y = nullptr;
  FftPlan plan_x{}, plan_y{}, plan_1d{};
  double *epsl = nullptr, *stepsz = nullptr, *delta = nullptr, *floor_ = nullptr;
  double *st_phi = nullptr, *st_rho = nullptr, *st_alp = nullptr;
  long long* iters = nullptr;
  double* stepsz_used = nullptr;
  int* nrec = nullptr;
  double* errlog = nullptr;
  int *end_reason = nullptr, *status = nullptr, *blocks_done = nullptr;
  long long* inner_total = nullptr;
  int* n_inner = nullptr;
  double* err_inner = nullptr;
  double* dbg_ns = nullptr;
  double ext_epsl = 0.0;
  bool ext_epsl_set = false;
  int xch_P = 0, xch_rank = 0, xch_nxl = 0, xch_kyl = 0, xch_nyh = 0, xch_pull = 0;     // fused transposes of the slab mode (pdhg_ext_set_exchange)
  void* xch_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  Knobs knobs{};              // diagnostic environment knobs, read once in pdhg_create
  int max_radix = 16;
  void* ws = nullptr;         // cooperative-kernel workspace
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // bracket the solver kernel(s) of the last march on its stream
  bool ev_valid = false;
  // lazily allocated device mirrors for the *_host entry points
  double *g_dev = nullptr, *phi_all = nullptr, *rho_all = nullptr, *alp_all_ref = nullptr, *alp_all_planar = nullptr;
  // scratch for reference-layout <-> planar conversion of single-block alp
  double* alp_tmp = nullptr;
};

template <typename T>
static cudaError_t dalloc(pdhg_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
  if (e != cudaSuccess) return e;
  h->owned.push_back(q);
  *p = static_cast<T*>(q);
  return cudaSuccess;
}

static bool make_plan(int n, FftPlan* plan, int max_radix = 16) {
  plan->n = n;
  plan->nstages = 0;
  int m = n;
  auto push = [&](int r) { if (plan->nstages < kMaxStages) plan->radix[plan->nstages++] = r; else m = -1; };
  while (max_radix >= 16 && m > 0 && m % 16 == 0) { push(16); m /= 16; }
  while (max_radix >= 8 && m > 0 && m % 8 == 0) { push(8); m /= 8; }
  while (m > 0 && m % 4 == 0) { push(4); m /= 4; }
  while (m > 0 && m % 2 == 0) { push(2); m /= 2; }
  while (m > 0 && m % 5 == 0) { push(5); m /= 5; }
  while (m > 0 && m % 3 == 0) { push(3); m /= 3; }
  for (int f = 7; m > 1 && f <= m; f += 2)
    while (m > 0 && m % f == 0) { push(f); m /= f; }
  return m == 1;
}

static std::vector<double2> make_twiddles(int n) {
  std::vector<double2> tw(n);
  for (int m = 0; m < n; ++m) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
    tw[m] = make_double2((double)cosl(a), (double)sinl(a));
  }
  return tw;
}

// lambda_k = (2 - 2 cos(2 pi k/n))/h^2 = 4 sin^2(pi k/n)/h^2 : minus the symbol of the periodic 3-point Laplacian
// (compute_Dxx_fft_fv, utils_precond.py:42-71, closed form per SURVEY.md A.6).
static double lap_symbol(int k, int n, double h) {
  const long double s = sinl(3.14159265358979323846264338327950288L * k / (long double)n);
  return (double)(4.0L * s * s / ((long double)h * (long double)h));
}

// -fv[kx][ky] for bc = (1, 0): the reference takes the DCT-II (x) and FFT (y) of the PERIODIC stencil
// (utils_precond.py:57-66), i.e. fv = 2 c_0 (-2/dx^2 - lambda_y) + 2 (c_1 + c_{nx-1})/dx^2, c_i = cos(pi kx (2i+1)/(2 nx)).
static double neumann_x_symbol(int kx, int nx, double dx, double lambda_y) {
  const long double PI = 3.14159265358979323846264338327950288L;
  auto c = [&](int i) { return cosl(PI * kx * (2.0L * i + 1.0L) / (2.0L * nx)); };
  const long double fv = 2.0L * c(0) * (-2.0L / ((long double)dx * dx) - (long double)lambda_y) + 2.0L * (c(1) + c(nx - 1)) / ((long double)dx * dx);
  return (double)(-fv);
}

extern "C" const char* pdhg_last_error(void) { return g_err.c_str(); }

extern "C" void pdhg_destroy(pdhg_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->cfg.device);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (void* p : h->owned) cudaFree(p);
  delete h;
}

extern "C" int pdhg_path(const pdhg_handle* h) { return h ? h->path : 0; }
extern "C" double pdhg_last_kernel_ms(const pdhg_handle* h) {
  if (!h || !h->ev_valid) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}
extern "C" int64_t pdhg_launch_count(const pdhg_handle* h) { return h ? h->launches : 0; }

static void fill_params(pdhg_handle* h, MarchParams* p);
static int upload_scalars(pdhg_handle* h, const double* epsl_host, const double* stepsz_host, cudaStream_t s, const double* stepsz_cur_host = nullptr);

extern "C" int pdhg_phase_times(pdhg_handle* h, double* out6 /* 16 doubles */) {
  if (!h || !out6) return fail(PDHG_ERR_ARG, "pdhg_phase_times: null argument");
  if (h->path != 2) {
    // single-CTA kernel: clock cycles of instance 0 per sub-step (only when PDHG_PROFILE is set)
    DeviceGuard guard(h->cfg.device);
    CU(cudaMemcpy(out6, h->dbg_ns, 16 * sizeof(double), cudaMemcpyDeviceToHost));
    return PDHG_OK;
  }
  MarchParams p;
  fill_params(h, &p);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  CU(coop_phase_times(p, h->ws, out6));
  return PDHG_OK;
}

extern "C" int pdhg_max_fuse(pdhg_handle* h) {
  if (!h || h->path != 2) return 1;
  DeviceGuard guard(h->cfg.device);
  MarchParams p;
  fill_params(h, &p);
  return coop_max_fuse(p);
}

extern "C" int pdhg_ext_phase(pdhg_handle* h, int phase, int pass_mask, double step, double epsl, const pdhg_ext_buffers* bufs,
                              int sum_lo, int sum_hi, int nyh_override, int ky_off, int nyh_tab, void* stream) {
  if (!h || !bufs || phase < 0 || phase > 5) return fail(PDHG_ERR_ARG, "pdhg_ext_phase: bad argument");   // 5: empty launch (diagnostic)
  if (h->B != 1) return fail(PDHG_ERR_ARG, "pdhg_ext_phase: slab mode needs a handle with batch = 1");
  if (h->path == 3 && !((phase == 1 && (bufs->zt || (h->xch_P > 0 && h->xch_pull))) || phase == 5))
    return fail(PDHG_ERR_ARG, "pdhg_ext_phase: a tables-only handle (path = 3) runs phase 1 on a caller-owned spectrum only");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  if (h->ext_epsl != epsl || !h->ext_epsl_set) {
    int rc = upload_scalars(h, &epsl, nullptr, s);
    if (rc) return rc;
    h->ext_epsl = epsl; h->ext_epsl_set = true;
  }
  MarchParams p;
  fill_params(h, &p);
  ExtPhaseDesc e;
  e.phi_in = bufs->phi_in; e.phi_out = bufs->phi_out; e.phib = bufs->phib; e.rho_in = bufs->rho_in; e.alp_in = bufs->alp_in;
  e.rho_out = bufs->rho_out; e.alp_out = bufs->alp_out; e.zt = bufs->zt; e.sums = bufs->sums;
  e.sum_lo = sum_lo; e.sum_hi = sum_hi; e.nyh_override = nyh_override; e.ky_off = ky_off; e.nyh_tab = nyh_tab;
  if (h->xch_P > 0 && phase <= 1) {
    e.xch_P = h->xch_P; e.xch_rank = h->xch_rank; e.xch_nxl = h->xch_nxl; e.xch_kyl = h->xch_kyl; e.xch_nyh = h->xch_nyh; e.xch_pull = h->xch_pull;
    for (int d = 0; d < 8; ++d) e.xch_ptr[d] = h->xch_ptr[d];
  }
  CU(launch_ext_phase(p, h->ws, phase, pass_mask, step, e, s));
  h->launches += 1;
  return PDHG_OK;
}

extern "C" int pdhg_ext_exchange_ok(pdhg_handle* h) {
  if (!h) return 0;
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return 0;
  MarchParams p;
  fill_params(h, &p);
  return (h->B == 1) ? coop_exchange_ok(p) : 0;
}

extern "C" int pdhg_ext_set_exchange(pdhg_handle* h, int P, int rank, int nxl, int kyl, int nyh, int gather, void* const* peer_ptrs) {
  if (!h) return fail(PDHG_ERR_ARG, "pdhg_ext_set_exchange: null handle");
  if (P == 0) { h->xch_P = 0; return PDHG_OK; }
  if (P < 1 || P > 8 || rank < 0 || rank >= P || nxl < 1 || kyl < 1 || nyh < 1 || !peer_ptrs)
    return fail(PDHG_ERR_ARG, "pdhg_ext_set_exchange: bad argument (1 <= P <= 8)");
  if (!pdhg_ext_exchange_ok(h))
    return fail(PDHG_ERR_UNSUPPORTED, "pdhg_ext_set_exchange: needs a 2-D periodic handle with K = 1 and batch = 1 whose transforms run the generic variants");
  for (int d = 0; d < P; ++d) if (!peer_ptrs[d]) return fail(PDHG_ERR_ARG, "pdhg_ext_set_exchange: null peer pointer");
  h->xch_P = P; h->xch_rank = rank; h->xch_nxl = nxl; h->xch_kyl = kyl; h->xch_nyh = nyh; h->xch_pull = gather ? 1 : 0;
  for (int d = 0; d < 8; ++d) h->xch_ptr[d] = (d < P) ? peer_ptrs[d] : nullptr;
  return PDHG_OK;
}

extern "C" int pdhg_debug_phase(pdhg_handle* h, int phase, int pass_mask, double step, int reps) {
  if (!h || phase < 0 || phase > 3 || reps < 1) return fail(PDHG_ERR_ARG, "pdhg_debug_phase: bad argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  MarchParams p;
  fill_params(h, &p);
  for (int r = 0; r < reps; ++r) { CU(launch_debug_phase(p, h->ws, phase, pass_mask, step, nullptr)); h->launches += 1; }
  CU(cudaDeviceSynchronize());
  return PDHG_OK;
}

extern "C" int pdhg_create(const pdhg_config* cfg, const double* coef_x, const double* coef_y, pdhg_handle** out) {
  if (!cfg || !out || !coef_x) return fail(PDHG_ERR_ARG, "pdhg_create: null argument");
  *out = nullptr;
  const pdhg_config& c = *cfg;
  if (c.ndim != 1 && c.ndim != 2) return fail(PDHG_ERR_ARG, "ndim must be 1 or 2");
  if (c.egno < 1 || c.egno > 3) return fail(PDHG_ERR_ARG, "egno must be 1, 2 or 3");
  if (c.egno == 3 && c.ndim != 2) return fail(PDHG_ERR_ARG, "egno 3 needs ndim 2");
  if (c.nx < 4 || (c.ndim == 2 && c.ny < 4)) return fail(PDHG_ERR_ARG, "nx, ny must be >= 4");
  if (c.ndim == 1 && c.ny != 1) return fail(PDHG_ERR_ARG, "ny must be 1 in 1-D");
  if (c.K < 1 || c.batch < 1 || c.nblocks < 1 || c.max_rec < 2) return fail(PDHG_ERR_ARG, "K, batch, nblocks >= 1, max_rec >= 2");
  if (c.ndim == 2 && !coef_y) return fail(PDHG_ERR_ARG, "coef_y required in 2-D");
  if (c.ndim == 1 && c.bc_x != 0) return fail(PDHG_ERR_UNSUPPORTED, "1-D supports periodic bc only (as the reference, utils_precond.py:121-124)");
  if (c.ndim == 2 && !((c.bc_x == 0 || c.bc_x == 1) && c.bc_y == 0))
    return fail(PDHG_ERR_UNSUPPORTED, "2-D supports bc (0,0) and (1,0) only (as the reference, utils_precond.py:157-163)");
  if (c.egno == 3 && c.K != 1) return fail(PDHG_ERR_UNSUPPORTED, "egno 3 needs time_step_per_PDHG = 2 (the reference's f_fn cannot broadcast otherwise, set_fns.py:98)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || c.device >= ndev)
    return fail(PDHG_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  DeviceGuard guard(c.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");

  pdhg_handle* h = new pdhg_handle();
  h->cfg = c;
  h->A = 2 * c.ndim;
  h->n = (size_t)c.nx * c.ny;
  h->B = c.batch;
  h->nyh = (c.ndim == 2) ? c.ny / 2 + 1 : 1;
  {
    auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    h->knobs.no_w256 = getenv("PDHG_NO_W256") != nullptr;
    h->knobs.force_w256 = getenv("PDHG_FORCE_W256") != nullptr;
    h->knobs.dfuse = env_int("PDHG_DFUSE", 0);
    h->knobs.tma = env_int("PDHG_TMA", -1);
    h->knobs.no_bslab = getenv("PDHG_NO_BSLAB") != nullptr;
    h->knobs.no_k1 = getenv("PDHG_NO_K1") != nullptr;
    h->knobs.profile = getenv("PDHG_PROFILE") != nullptr;
    h->max_radix = env_int("PDHG_MAX_RADIX", 16);
    if (h->max_radix > 16 || h->max_radix < 2) h->max_radix = 16;
  }
  if (!make_plan(c.nx, &h->plan_x, h->max_radix) || (c.ndim == 2 && !make_plan(c.ny, &h->plan_y, h->max_radix))) {
    delete h;
    return fail(PDHG_ERR_UNSUPPORTED, "grid size has too many prime factors for the FFT plan");
  }
  if (c.ndim == 1) { h->plan_y.n = 1; h->plan_y.nstages = 0; }
  make_plan(c.nx, &h->plan_1d, h->max_radix < 8 ? h->max_radix : 8);
  // path selection
  int dev_smem = 0;
  cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device);
  const bool fits1 = (c.ndim == 1) && pdhg1d_cta_smem_bytes(c.nx, c.K) <= (size_t)dev_smem;
  h->path = c.path;
  if (h->path == 3 && !(c.ndim == 2 && c.batch == 1 && c.K == 1)) { delete h; return fail(PDHG_ERR_ARG, "path 3 (tables only) needs ndim = 2, K = 1, batch = 1"); }
  if (h->path == 0) h->path = fits1 ? 1 : 2;
  if (h->path == 1 && !fits1) { delete h; return fail(PDHG_ERR_ARG, "path 1 (single-CTA) needs 1-D state that fits shared memory"); }

  auto bail = [&](cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    pdhg_destroy(h);
    return fail(e == cudaErrorMemoryAllocation ? PDHG_ERR_NOMEM : PDHG_ERR_CUDA, m);
  };
#define CB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(e_, #call); } while (0)

  // tables
  std::vector<double> diag((size_t)c.nx * h->nyh);
  for (int kx = 0; kx < c.nx; ++kx) {
    const double lx = lap_symbol(kx, c.nx, c.dx);
    if (c.ndim == 1) {
      diag[kx] = pow(lx + c.C, c.pow);                    // utils_precond.py:125-126
    } else {
      for (int ky = 0; ky < h->nyh; ++ky) {
        const double ly = lap_symbol(ky, c.ny, c.dy);
        diag[(size_t)kx * h->nyh + ky] = ((c.bc_x == 1) ? neumann_x_symbol(kx, c.nx, c.dx, ly) : lx + ly) + c.C;   // :168
      }
    }
  }
  CB(dalloc(h, &h->diag, diag.size()));
  CB(cudaMemcpy(h->diag, diag.data(), diag.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (c.ndim == 1 && c.K == 1 && c.nx <= 2048) {
    // G = IFFT(1/d), d = (lambda + C)^pow + Ct/dt^2: the K = 1 preconditioner as a circular convolution (pdhg1d_k1.cu).
    // Real and even; evaluated in long double; R = last tap with |G[j]| >= 2^-56 G[0].
    const int n = c.nx, half = n / 2;
    std::vector<long double> cs(n), inv(n);
    for (int m = 0; m < n; ++m) cs[m] = cosl(2.0L * 3.14159265358979323846264338327950288L * m / (long double)n);
    const long double ct2 = (long double)c.Ct / ((long double)c.dt * c.dt);
    for (int k = 0; k < n; ++k) inv[k] = 1.0L / ((long double)diag[k] + ct2);
    std::vector<double> G(half + 1);
    for (int j = 0; j <= half; ++j) {
      long double acc = 0.0L;
      for (int k = 0; k < n; ++k) acc += inv[k] * cs[(int)(((long long)k * j) % n)];
      G[j] = (double)(acc / n);
    }
    int R = 0;
    const double thr = ldexp(fabs(G[0]), -56);
    for (int j = 0; j <= half; ++j) if (fabs(G[j]) >= thr) R = j;
    if (pdhg1d_k1_supported(n, c.K, R)) {
      h->green_R = R;
      CB(dalloc(h, &h->green, (size_t)R + 1));
      CB(cudaMemcpy(h->green, G.data(), ((size_t)R + 1) * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  CB(dalloc(h, &h->coef_x, (size_t)c.nx));
  CB(cudaMemcpy(h->coef_x, coef_x, c.nx * sizeof(double), cudaMemcpyHostToDevice));
  CB(dalloc(h, &h->coef_y, (size_t)c.ny));
  if (coef_y) CB(cudaMemcpy(h->coef_y, coef_y, c.ny * sizeof(double), cudaMemcpyHostToDevice));
  if (c.bc_x == 1) {
    std::vector<double> ct((size_t)4 * c.nx);
    for (int m = 0; m < 4 * c.nx; ++m) ct[m] = (double)cosl(3.14159265358979323846264338327950288L * m / (2.0L * c.nx));
    CB(dalloc(h, &h->dct_cos, ct.size()));
    CB(cudaMemcpy(h->dct_cos, ct.data(), ct.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  {
    std::vector<double2> tx = make_twiddles(c.nx);
    CB(dalloc(h, &h->tw_x, tx.size()));
    CB(cudaMemcpy(h->tw_x, tx.data(), tx.size() * sizeof(double2), cudaMemcpyHostToDevice));
    std::vector<double2> ty = make_twiddles(c.ny);
    CB(dalloc(h, &h->tw_y, ty.size()));
    CB(cudaMemcpy(h->tw_y, ty.data(), ty.size() * sizeof(double2), cudaMemcpyHostToDevice));
  }
  const size_t B = h->B, NB = (size_t)c.nblocks;
  CB(dalloc(h, &h->epsl, B)); CB(dalloc(h, &h->stepsz, B)); CB(dalloc(h, &h->delta, B)); CB(dalloc(h, &h->floor_, B));
  const bool tables_only = (h->path == 3);     // slab mode's phase-B handle: per-mode table, twiddles, plans - no state, no workspace
  if (!tables_only) {
    CB(dalloc(h, &h->st_phi, B * (c.K + 1) * h->n));
    CB(dalloc(h, &h->st_rho, B * c.K * h->n));
    CB(dalloc(h, &h->st_alp, B * h->A * c.K * h->n));
  }
  CB(dalloc(h, &h->iters, B * NB)); CB(dalloc(h, &h->stepsz_used, B * NB)); CB(dalloc(h, &h->nrec, B * NB));
  CB(dalloc(h, &h->errlog, B * NB * c.max_rec * kLogCols)); CB(dalloc(h, &h->end_reason, B * NB));
  CB(dalloc(h, &h->status, B)); CB(dalloc(h, &h->blocks_done, B)); CB(dalloc(h, &h->inner_total, B));
  CB(dalloc(h, &h->n_inner, B));
  CB(dalloc(h, &h->err_inner, B));
  CB(dalloc(h, &h->dbg_ns, 16));
  CB(cudaMemset(h->dbg_ns, 0, 16 * sizeof(double)));
  if (c.ndim == 2 && c.n_ctrl > 1 && !tables_only) CB(dalloc(h, &h->alp_tmp, B * h->A * c.K * h->n * c.n_ctrl));
  if (tables_only) {
    char* w = nullptr;                         // (a token allocation: phase 1 on a caller-owned spectrum touches none of the carved arrays)
    CB(dalloc(h, &w, 4096));
    h->ws = w;
  } else {
    // the cooperative kernel also serves the operator-level entry points, so its workspace always exists
    MarchParams p{};
    p.ndim = c.ndim; p.nx = c.nx; p.ny = c.ny; p.K = c.K;
    const size_t wsb = pdhg_coop_workspace_bytes(p, h->B);
    char* w = nullptr;
    CB(dalloc(h, &w, wsb));
    CB(cudaMemset(w, 0, wsb));
    h->ws = w;
  }
#undef CB
  *out = h;
  return PDHG_OK;
}

static void fill_params(pdhg_handle* h, MarchParams* p) {
  const pdhg_config& c = h->cfg;
  memset(p, 0, sizeof(*p));
  p->ndim = c.ndim; p->egno = c.egno; p->nx = c.nx; p->ny = c.ny; p->K = c.K;
  p->bc_x = c.bc_x; p->bc_y = c.bc_y;
  p->dt = c.dt; p->dx = c.dx; p->dy = c.dy; p->c_on_rho = c.c_on_rho;
  p->eps = c.eps; p->rho_alp_iters = c.rho_alp_iters;
  p->max_rec = c.max_rec;
  p->epsl = h->epsl; p->stepsz = h->stepsz; p->stepsz_delta = h->delta; p->stepsz_floor = h->floor_;
  p->coef_x = h->coef_x; p->coef_y = h->coef_y; p->diag = h->diag; p->dct_cos = h->dct_cos;
  p->green = h->green; p->green_R = h->green_R; p->tw_x = h->tw_x; p->tw_y = h->tw_y;
  p->Ct_over_dt2 = (c.ndim == 1 ? c.Ct : 1.0) / (c.dt * c.dt);
  p->plan_x = h->plan_x; p->plan_y = h->plan_y; p->plan_1d = h->plan_1d;
  p->st_phi = h->st_phi; p->st_rho = h->st_rho; p->st_alp = h->st_alp;
  p->iters = h->iters; p->stepsz_used = h->stepsz_used; p->nrec = h->nrec; p->errlog = h->errlog;
  p->knobs = h->knobs;
  p->dbg_ns = h->knobs.profile ? h->dbg_ns : nullptr;
  p->end_reason = h->end_reason; p->status = h->status; p->blocks_done = h->blocks_done; p->inner_total = h->inner_total;
}

static int upload_scalars(pdhg_handle* h, const double* epsl_host, const double* stepsz_host, cudaStream_t s, const double* stepsz_cur_host) {
  const int B = h->B;
  std::vector<double> e(B), st(B), dl(B), fl(B);
  for (int b = 0; b < B; ++b) {
    e[b] = epsl_host ? epsl_host[b] : 0.0;
    const double st0 = stepsz_host ? stepsz_host[b] : 0.0;     // the initial stepsz_param: decrement and floor derive from it
    st[b] = stepsz_cur_host ? stepsz_cur_host[b] : st0;        // the step size this launch starts with (resume: after earlier fallbacks)
    const double mn = st0 / 10;            // stepsz_param_min  (utils_pdhg_solver.py:160)
    dl[b] = st0 / 10;                      // stepsz_param_delta (:161)
    fl[b] = mn + dl[b];                    // threshold of :181
  }
  CU(cudaMemcpyAsync(h->epsl, e.data(), B * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->stepsz, st.data(), B * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->delta, dl.data(), B * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(h->floor_, fl.data(), B * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));   // the staging vectors die at return
  return PDHG_OK;
}

static int download_logs(pdhg_handle* h, int nblocks, pdhg_logs* logs, cudaStream_t s) {
  CU(cudaStreamSynchronize(s));
  if (!logs) return PDHG_OK;
  const size_t B = h->B, NB = (size_t)nblocks;
  if (logs->iters) CU(cudaMemcpy(logs->iters, h->iters, B * NB * sizeof(long long), cudaMemcpyDeviceToHost));
  if (logs->stepsz_used) CU(cudaMemcpy(logs->stepsz_used, h->stepsz_used, B * NB * sizeof(double), cudaMemcpyDeviceToHost));
  if (logs->nrec) CU(cudaMemcpy(logs->nrec, h->nrec, B * NB * sizeof(int), cudaMemcpyDeviceToHost));
  if (logs->errlog) CU(cudaMemcpy(logs->errlog, h->errlog, B * NB * h->cfg.max_rec * kLogCols * sizeof(double), cudaMemcpyDeviceToHost));
  if (logs->end_reason) CU(cudaMemcpy(logs->end_reason, h->end_reason, B * NB * sizeof(int), cudaMemcpyDeviceToHost));
  if (logs->status) CU(cudaMemcpy(logs->status, h->status, B * sizeof(int), cudaMemcpyDeviceToHost));
  if (logs->blocks_done) CU(cudaMemcpy(logs->blocks_done, h->blocks_done, B * sizeof(int), cudaMemcpyDeviceToHost));
  if (logs->stepsz_final) CU(cudaMemcpy(logs->stepsz_final, h->stepsz, B * sizeof(double), cudaMemcpyDeviceToHost));
  if (logs->inner_total) CU(cudaMemcpy(logs->inner_total, h->inner_total, B * sizeof(long long), cudaMemcpyDeviceToHost));
  return PDHG_OK;
}

static int run_march(pdhg_handle* h, const MarchParams& p, cudaStream_t s) {
  if (!h->ev0) { CU(cudaEventCreate(&h->ev0)); CU(cudaEventCreate(&h->ev1)); }
  CU(cudaEventRecord(h->ev0, s));
  struct Rec { pdhg_handle* h; cudaStream_t s; ~Rec() { h->ev_valid = (cudaEventRecord(h->ev1, s) == cudaSuccess); } } rec{h, s};
  if (h->path == 1) {
    if (h->green_R >= 0 && !h->knobs.no_k1) CU(launch_pdhg1d_k1(p, h->B, s));   // K = 1 register-resident kernel
    else CU(launch_pdhg1d_cta(p, h->B, s));
    h->launches += 1;
  } else {
    CU(launch_pdhg_coop(p, h->B, h->ws, s, &h->launches));
  }
  return PDHG_OK;
}

// alp in reference layout [B][A][K][n][n_ctrl] <-> planar active-only [B][A][K][n]
static int alp_to_planar(pdhg_handle* h, const double* ref, double* planar, cudaStream_t s) {
  const pdhg_config& c = h->cfg;
  const size_t kn = (size_t)c.K * h->n;
  if (c.n_ctrl == 1) {
    CU(cudaMemcpyAsync(planar, ref, (size_t)h->B * h->A * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  } else {
    CU(launch_pack_alp(ref, planar, h->B, h->A, kn, c.n_ctrl, c.ndim, c.egno, 1, s));
    h->launches += 1;
  }
  return PDHG_OK;
}
static int alp_from_planar(pdhg_handle* h, const double* planar, double* ref, size_t kn, cudaStream_t s) {
  const pdhg_config& c = h->cfg;
  if (c.n_ctrl == 1) {
    CU(cudaMemcpyAsync(ref, planar, (size_t)h->B * h->A * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  } else {
    CU(launch_pack_alp(ref, const_cast<double*>(planar), h->B, h->A, kn, c.n_ctrl, c.ndim, c.egno, 0, s));
    h->launches += 1;
  }
  return PDHG_OK;
}

extern "C" int pdhg_solve_block(pdhg_handle* h, const double* phi0, const double* rho0, const double* alp0,
                                const double* epsl_host, const double* stepsz_host, int64_t n_maxiter,
                                int64_t iter_begin, int64_t iter_pause, int32_t print_freq, double* phi_out,
                                double* rho_out, double* alp_out, pdhg_logs* logs, void* stream) {
  if (!h || !phi0 || !rho0 || !alp0 || !stepsz_host || !phi_out || !rho_out || !alp_out)
    return fail(PDHG_ERR_ARG, "pdhg_solve_block: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  if (n_maxiter < 1 || iter_begin < 0 || iter_begin >= n_maxiter) return fail(PDHG_ERR_ARG, "pdhg_solve_block: bad iteration range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  const pdhg_config& c = h->cfg;
  const size_t B = h->B, kn = (size_t)c.K * h->n, np = (size_t)(c.K + 1) * h->n;
  int rc = upload_scalars(h, epsl_host, stepsz_host, s);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->st_phi, phi0, B * np * sizeof(double), cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(h->st_rho, rho0, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  rc = alp_to_planar(h, alp0, h->st_alp, s);
  if (rc) return rc;
  MarchParams p;
  fill_params(h, &p);
  p.n_maxiter = n_maxiter; p.iter_begin = iter_begin;
  p.iter_pause = (iter_pause > 0) ? iter_pause : n_maxiter;
  p.print_freq = print_freq;
  p.nblocks = 1; p.blk_begin = 0; p.blk_end = 1; p.handoff = 0; p.fallback = 0;
  rc = run_march(h, p, s);
  if (rc) return rc;
  CU(cudaMemcpyAsync(phi_out, h->st_phi, B * np * sizeof(double), cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(rho_out, h->st_rho, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  rc = alp_from_planar(h, h->st_alp, alp_out, kn, s);
  if (rc) return rc;
  return download_logs(h, 1, logs, s);
}

static int multi_step_impl(pdhg_handle* h, const double* g_dev, const double* epsl_host, const double* stepsz_host,
                           int64_t n_maxiter, int32_t print_freq, double* phi_all, double* rho_all, double* alp_all,
                           pdhg_logs* logs, cudaStream_t s, int blk_begin = 0, int blk_end = -1, const double* stepsz_cur_host = nullptr) {
  const pdhg_config& c = h->cfg;
  if (blk_end < 0) blk_end = c.nblocks;
  int rc = upload_scalars(h, epsl_host, stepsz_host, s, stepsz_cur_host);
  if (rc) return rc;
  MarchParams p;
  fill_params(h, &p);
  p.n_maxiter = n_maxiter; p.iter_begin = 0; p.iter_pause = n_maxiter; p.print_freq = print_freq;
  p.nblocks = c.nblocks; p.blk_begin = blk_begin; p.blk_end = blk_end; p.handoff = 1; p.fallback = 1;
  p.phi_all = phi_all; p.rho_all = rho_all;
  const size_t kn_all = (size_t)c.nblocks * c.K * h->n;
  if (c.n_ctrl == 1) {
    p.alp_all = alp_all;
  } else {
    if (!h->alp_all_planar) CU(dalloc(h, &h->alp_all_planar, (size_t)h->B * h->A * kn_all));
    p.alp_all = h->alp_all_planar;
  }
  if (blk_begin == 0) {      // phi0 = tile(g), rho0 = c_on_rho, alp0 = 0 (utils_pdhg_solver.py:123-137); later ranges continue from the handle's state
    CU(launch_init_state(p, g_dev, h->B, s));
    h->launches += 1;
  }
  rc = run_march(h, p, s);
  if (rc) return rc;
  if (c.n_ctrl != 1) {
    rc = alp_from_planar(h, h->alp_all_planar, alp_all, kn_all, s);
    if (rc) return rc;
  }
  return download_logs(h, c.nblocks, logs, s);
}

extern "C" int pdhg_multi_step_range(pdhg_handle* h, const double* g_dev, const double* epsl_host, const double* stepsz_host,
                                     const double* stepsz_cur_host, int64_t n_maxiter, int32_t print_freq, int32_t blk_begin,
                                     int32_t blk_end, double* phi_all, double* rho_all, double* alp_all, pdhg_logs* logs, void* stream) {
  if (!h || !stepsz_host || !phi_all || !rho_all || !alp_all) return fail(PDHG_ERR_ARG, "pdhg_multi_step_range: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  if (n_maxiter < 1) return fail(PDHG_ERR_ARG, "pdhg_multi_step_range: n_maxiter must be >= 1");
  if (blk_begin < 0 || blk_end > h->cfg.nblocks || blk_begin >= blk_end) return fail(PDHG_ERR_ARG, "pdhg_multi_step_range: bad block range");
  if (blk_begin == 0 && !g_dev) return fail(PDHG_ERR_ARG, "pdhg_multi_step_range: g is required when the range starts at block 0");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  return multi_step_impl(h, g_dev, epsl_host, stepsz_host, n_maxiter, print_freq, phi_all, rho_all, alp_all, logs,
                         static_cast<cudaStream_t>(stream), blk_begin, blk_end, stepsz_cur_host);
}

// marching state between time blocks: phi0 of the next block (already warm-started, utils_pdhg_solver.py:200-203), rho0, alp0
extern "C" int pdhg_get_march_state(pdhg_handle* h, double* phi0, double* rho0, double* alp0, void* stream) {
  if (!h || !phi0 || !rho0 || !alp0) return fail(PDHG_ERR_ARG, "pdhg_get_march_state: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const pdhg_config& c = h->cfg;
  const size_t B = h->B, kn = (size_t)c.K * h->n, np = (size_t)(c.K + 1) * h->n;
  CU(cudaMemcpyAsync(phi0, h->st_phi, B * np * sizeof(double), cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(rho0, h->st_rho, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return alp_from_planar(h, h->st_alp, alp0, kn, s);
}
extern "C" int pdhg_set_march_state(pdhg_handle* h, const double* phi0, const double* rho0, const double* alp0, void* stream) {
  if (!h || !phi0 || !rho0 || !alp0) return fail(PDHG_ERR_ARG, "pdhg_set_march_state: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const pdhg_config& c = h->cfg;
  const size_t B = h->B, kn = (size_t)c.K * h->n, np = (size_t)(c.K + 1) * h->n;
  CU(cudaMemcpyAsync(h->st_phi, phi0, B * np * sizeof(double), cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(h->st_rho, rho0, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return alp_to_planar(h, alp0, h->st_alp, s);
}

extern "C" int pdhg_multi_step(pdhg_handle* h, const double* g_dev, const double* epsl_host, const double* stepsz_host,
                               int64_t n_maxiter, int32_t print_freq, double* phi_all, double* rho_all, double* alp_all,
                               pdhg_logs* logs, void* stream) {
  if (!h || !g_dev || !stepsz_host || !phi_all || !rho_all || !alp_all) return fail(PDHG_ERR_ARG, "pdhg_multi_step: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  if (n_maxiter < 1) return fail(PDHG_ERR_ARG, "pdhg_multi_step: n_maxiter must be >= 1");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  return multi_step_impl(h, g_dev, epsl_host, stepsz_host, n_maxiter, print_freq, phi_all, rho_all, alp_all, logs,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int pdhg_multi_step_host(pdhg_handle* h, const double* g_host, const double* epsl_host,
                                    const double* stepsz_host, int64_t n_maxiter, int32_t print_freq,
                                    double* phi_all_host, double* rho_all_host, double* alp_all_host, pdhg_logs* logs) {
  if (!h || !g_host || !stepsz_host || !phi_all_host || !rho_all_host || !alp_all_host)
    return fail(PDHG_ERR_ARG, "pdhg_multi_step_host: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  if (n_maxiter < 1) return fail(PDHG_ERR_ARG, "pdhg_multi_step_host: n_maxiter must be >= 1");
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  const pdhg_config& c = h->cfg;
  const size_t B = h->B, n = h->n, nt = (size_t)c.nblocks * c.K + 1;
  const size_t n_phi = B * nt * n, n_rho = B * (nt - 1) * n, n_alp = B * h->A * (nt - 1) * n * c.n_ctrl;
  if (!h->g_dev) {
    CU(dalloc(h, &h->g_dev, B * n));
    CU(dalloc(h, &h->phi_all, n_phi));
    CU(dalloc(h, &h->rho_all, n_rho));
    CU(dalloc(h, &h->alp_all_ref, n_alp));
  }
  cudaStream_t s = nullptr;
  CU(cudaMemcpyAsync(h->g_dev, g_host, B * n * sizeof(double), cudaMemcpyHostToDevice, s));
  int rc = multi_step_impl(h, h->g_dev, epsl_host, stepsz_host, n_maxiter, print_freq, h->phi_all, h->rho_all,
                           h->alp_all_ref, logs, s);
  if (rc) return rc;
  CU(cudaMemcpyAsync(phi_all_host, h->phi_all, n_phi * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(rho_all_host, h->rho_all, n_rho * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(alp_all_host, h->alp_all_ref, n_alp * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return PDHG_OK;
}

extern "C" int pdhg_update_primal(pdhg_handle* h, const double* phi_prev, const double* rho_prev, const double* alp_prev,
                                  const double* epsl_host, double tau, double* phi_next, void* stream) {
  if (!h || !phi_prev || !rho_prev || !alp_prev || !phi_next) return fail(PDHG_ERR_ARG, "pdhg_update_primal: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  const pdhg_config& c = h->cfg;
  const size_t B = h->B, kn = (size_t)c.K * h->n;
  int rc = upload_scalars(h, epsl_host, nullptr, s);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->st_rho, rho_prev, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  rc = alp_to_planar(h, alp_prev, h->st_alp, s);
  if (rc) return rc;
  MarchParams p;
  fill_params(h, &p);
  CU(launch_update_primal(p, h->B, phi_prev, tau, phi_next, h->ws, s, &h->launches));
  return PDHG_OK;
}

extern "C" int pdhg_update_dual(pdhg_handle* h, const double* phi_bar, const double* rho_prev, const double* alp_prev,
                                const double* epsl_host, double sigma, double eps, double* rho_next, double* alp_next,
                                int32_t* n_inner_host, double* err_host, void* stream) {
  if (!h || !phi_bar || !rho_prev || !alp_prev || !rho_next || !alp_next) return fail(PDHG_ERR_ARG, "pdhg_update_dual: null argument");
  if (h->path == 3) return fail(PDHG_ERR_ARG, "handle created with path = 3 (tables only): it serves pdhg_ext_phase phase 1 on caller buffers, nothing else");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->cfg.device);
  if (!guard.ok()) return fail(PDHG_ERR_CUDA, "cudaSetDevice failed");
  const pdhg_config& c = h->cfg;
  const size_t B = h->B, kn = (size_t)c.K * h->n;
  int rc = upload_scalars(h, epsl_host, nullptr, s);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->st_rho, rho_prev, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  rc = alp_to_planar(h, alp_prev, h->st_alp, s);
  if (rc) return rc;
  MarchParams p;
  fill_params(h, &p);
  CU(launch_update_dual(p, h->B, phi_bar, sigma, eps, h->n_inner, h->err_inner, h->ws, s, &h->launches));
  CU(cudaMemcpyAsync(rho_next, h->st_rho, B * kn * sizeof(double), cudaMemcpyDeviceToDevice, s));
  rc = alp_from_planar(h, h->st_alp, alp_next, kn, s);
  if (rc) return rc;
  CU(cudaStreamSynchronize(s));
  if (n_inner_host) CU(cudaMemcpy(n_inner_host, h->n_inner, B * sizeof(int), cudaMemcpyDeviceToHost));
  if (err_host) CU(cudaMemcpy(err_host, h->err_inner, B * sizeof(double), cudaMemcpyDeviceToHost));
  return PDHG_OK;
}
