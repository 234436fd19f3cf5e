// Closed-loop trajectories under the computed feedback control (jaxsrc/run_example.py:18-155, compute_traj_1d / compute_traj_2d):
//   dx_t = f(alp(x_t, t), x_t) dt + sqrt(2 epsl) dW_t,   alp(x, t) interpolated from the solver's control arrays,
// Euler-Maruyama with the reference's own time grid.  One thread marches one sample through all nt - 1 steps (the steps of a
// sample are sequential, samples are independent), so a batch of initial points is one launch; the Gaussian increments are an
// INPUT (host-seeded, [nt-1][n_sample][ndim]) exactly as the reference draws them on the host with numpy.random.
//   interpolation: 1-D numpy.interp(x, x_arr, alp, period) (:35-36) or nearest grid point of x mod period (:37-44);
//                  2-D scipy interpn on the periodically extended / edge-clamped grid (extend_bdry_2d :53-110, :133-141),
//                  'linear' or 'nearest' (egno 2).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pdhg_b200.h"

namespace pdhg {

struct TrajParams {
  int ndim, egno, n_ctrl, nx, ny, nt, n_sample, bc_x, bc_y, nearest;
  double Px, Py, epsl;
  const double *alp, *xa, *ya, *t_arr, *noise, *x_init;
  double *traj_x, *traj_alp;
};

__device__ __forceinline__ double pymod(double x, double p) {     // Python / NumPy `%` for p > 0
  double r = fmod(x, p);
  if (r < 0.0) r += p;
  if (r >= p) r -= p;          // (-tiny) % p rounds to p in fmod + p
  return r;
}
// largest i in [0, n-1] with xa[i] <= x (xa ascending); -1 if x < xa[0]
__device__ __forceinline__ int lower_node(const double* xa, int n, double x) {
  int lo = -1, hi = n - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (xa[mid] <= x) lo = mid; else hi = mid - 1; }
  return lo;
}
struct Node { int i0, i1; double w; };     // value = (1 - w) v[i0] + w v[i1]

// 1-D numpy.interp with period: x and the (already sorted, in [0, period)) nodes are reduced mod period, the node set is
// closed periodically (numpy/lib/function_base.py interp: xp = [xp[-1] - period, xp, xp[0] + period])
__device__ __forceinline__ Node interp_periodic_np(const double* xa, int n, double P, double x) {
  const double xr = pymod(x, P);
  const int i = lower_node(xa, n, xr);
  Node r;
  if (i < 0) { r.i0 = n - 1; r.i1 = 0; const double x0 = xa[n - 1] - P; r.w = (xr - x0) / (xa[0] - x0); }
  else if (i == n - 1) { r.i0 = n - 1; r.i1 = 0; r.w = (xr - xa[n - 1]) / (xa[0] + P - xa[n - 1]); }
  else { r.i0 = i; r.i1 = i + 1; r.w = (xr - xa[i]) / (xa[i + 1] - xa[i]); }
  return r;
}
// nearest node of x mod period among the nodes themselves, no wrap (argmin |x_arr - x mod period|, first minimum, :38-39)
__device__ __forceinline__ int nearest_np(const double* xa, int n, double P, double x) {
  const double xr = pymod(x, P);
  const int i = lower_node(xa, n, xr);
  if (i < 0) return 0;
  if (i == n - 1) return n - 1;
  return (fabs(xa[i + 1] - xr) < fabs(xa[i] - xr)) ? i + 1 : i;
}
// one axis of scipy.interpolate.interpn on the grid of extend_bdry_2d: periodic copies of the nodes (bc 0) or the edge values
// repeated outside (bc 1); `nearest`: norm_dist <= 0.5 -> lower node (scipy _evaluate_nearest)
__device__ __forceinline__ Node interp_axis_ext(const double* xa, int n, double P, int bc, double x, int nearest) {
  Node r;
  if (bc == 0) {
    const double m = floor((x - xa[0]) / P);
    double xr = x - m * P;
    if (xr < xa[0]) xr = xa[0];
    int i = lower_node(xa, n, xr);
    if (i < 0) i = 0;
    const double right = (i == n - 1) ? xa[0] + P : xa[i + 1];
    r.i0 = i; r.i1 = (i == n - 1) ? 0 : i + 1;
    r.w = (xr - xa[i]) / (right - xa[i]);
  } else {
    if (x <= xa[0]) { r.i0 = r.i1 = 0; r.w = 0.0; }
    else if (x >= xa[n - 1]) { r.i0 = r.i1 = n - 1; r.w = 0.0; }
    else { const int i = lower_node(xa, n, x); r.i0 = i; r.i1 = i + 1; r.w = (x - xa[i]) / (xa[i + 1] - xa[i]); }
  }
  if (nearest) { if (r.w <= 0.5) r.i1 = r.i0; else r.i0 = r.i1; r.w = 0.0; }
  return r;
}
__device__ __forceinline__ double a_coef(double v) { return (v - 1.0) * (v - 1.0) + 0.1; }      // set_fns.py:117-118,145
__device__ __forceinline__ double keep_nonneg(double f) { return f * ((f >= 0.0) ? 1.0 : 0.0); }   // f * (f >= 0)
__device__ __forceinline__ double keep_neg(double f) { return f * ((f < 0.0) ? 1.0 : 0.0); }

__global__ void traj_kernel(const TrajParams p) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.n_sample) return;
  const int nt1 = p.nt - 1;
  if (p.ndim == 1) {
    double x = p.x_init[s];
    p.traj_x[s] = x;
    const size_t kn = (size_t)nt1 * p.nx;
    for (int k = 0; k < nt1; ++k) {
      const double dt = p.t_arr[k + 1] - p.t_arr[k];
      const double* a1 = p.alp + (size_t)k * p.nx;
      const double* a2 = a1 + kn;
      double v1, v2;
      if (p.nearest) { const int i = nearest_np(p.xa, p.nx, p.Px, x); v1 = a1[i]; v2 = a2[i]; }
      else {
        const Node nd = interp_periodic_np(p.xa, p.nx, p.Px, x);
        v1 = (a1[nd.i1] - a1[nd.i0]) * nd.w + a1[nd.i0];          // numpy: slope * (x - xp[j]) + fp[j]
        v2 = (a2[nd.i1] - a2[nd.i0]) * nd.w + a2[nd.i0];
      }
      p.traj_alp[(size_t)k * p.n_sample + s] = v1 + v2;
      const double ax = a_coef(pymod(x, p.Px));
      const double vel = keep_nonneg(-(v1 * ax)) + keep_neg(-(v2 * ax));       // get_f_vals_1d with f = -alp a(x)
      double nz = 0.0;
      if (p.noise) nz = sqrt(2.0 * p.epsl * dt) * p.noise[(size_t)k * p.n_sample + s];
      x = x + vel * dt + nz;
      p.traj_x[(size_t)(k + 1) * p.n_sample + s] = x;
    }
    return;
  }
  // ---- 2-D: alp [4][nt-1][nx][ny][n_ctrl] ----
  const int nc = p.n_ctrl;
  double x0 = p.x_init[2 * s], x1 = p.x_init[2 * s + 1];
  p.traj_x[2 * (size_t)s] = x0; p.traj_x[2 * (size_t)s + 1] = x1;
  const size_t plane = (size_t)p.nx * p.ny * nc, kn = (size_t)nt1 * plane;
  for (int k = 0; k < nt1; ++k) {
    const double dt = p.t_arr[k + 1] - p.t_arr[k];
    const Node nx_ = interp_axis_ext(p.xa, p.nx, p.Px, p.bc_x, x0, p.nearest);
    const Node ny_ = interp_axis_ext(p.ya, p.ny, p.Py, p.bc_y, x1, p.nearest);
    double v[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double* a = p.alp + (size_t)j * kn + (size_t)k * plane;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nc) {
          const double v00 = a[((size_t)nx_.i0 * p.ny + ny_.i0) * nc + c], v01 = a[((size_t)nx_.i0 * p.ny + ny_.i1) * nc + c];
          const double v10 = a[((size_t)nx_.i1 * p.ny + ny_.i0) * nc + c], v11 = a[((size_t)nx_.i1 * p.ny + ny_.i1) * nc + c];
          // scipy _evaluate_linear: sum over the corners of value * prod of (1 - w) or w, x-corner outer loop
          v[j][c] = v00 * ((1.0 - nx_.w) * (1.0 - ny_.w)) + v01 * ((1.0 - nx_.w) * ny_.w) + v10 * (nx_.w * (1.0 - ny_.w)) + v11 * (nx_.w * ny_.w);
        } else {
          v[j][c] = 0.0;
        }
      }
    }
    for (int c = 0; c < nc; ++c) p.traj_alp[((size_t)k * p.n_sample + s) * nc + c] = v[0][c] + v[1][c] + v[2][c] + v[3][c];
    const double xm0 = (p.bc_x == 0) ? pymod(x0, p.Px) : x0, xm1 = pymod(x1, p.Py);      // x_curr_in_period (:143-146)
    double vx, vy;
    if (p.egno == 3) {          // f = [alp, x']  (set_fns.py:98)
      vx = keep_nonneg(v[0][0]) + keep_neg(v[1][0]);
      vy = keep_nonneg(xm0) + keep_neg(xm0);
    } else {                    // f = -[a(x) alp_x, a(y) alp_y]
      const double ax = a_coef(xm0), ay = a_coef(xm1);
      vx = keep_nonneg(-(ax * v[0][0])) + keep_neg(-(ax * v[1][0]));
      vy = keep_nonneg(-(ay * v[2][1])) + keep_neg(-(ay * v[3][1]));
    }
    double n0 = 0.0, n1 = 0.0;
    if (p.noise) {
      const double sd = sqrt(2.0 * p.epsl * dt);
      n0 = sd * p.noise[((size_t)k * p.n_sample + s) * 2]; n1 = sd * p.noise[((size_t)k * p.n_sample + s) * 2 + 1];
    }
    x0 = x0 + vx * dt + n0; x1 = x1 + vy * dt + n1;
    p.traj_x[((size_t)(k + 1) * p.n_sample + s) * 2] = x0; p.traj_x[((size_t)(k + 1) * p.n_sample + s) * 2 + 1] = x1;
  }
}

}  // namespace pdhg

extern "C" int pdhg_compute_traj(int32_t ndim, int32_t egno, int32_t n_ctrl, int32_t nx, int32_t ny, int32_t nt, int32_t n_sample, int32_t bc_x,
                                 int32_t bc_y, int32_t nearest, double x_period, double y_period, double epsl, const double* alp_dev,
                                 const double* x_nodes_dev, const double* y_nodes_dev, const double* t_arr_dev, const double* noise_dev,
                                 const double* x_init_dev, double* traj_x_dev, double* traj_alp_dev, void* stream) {
  if ((ndim != 1 && ndim != 2) || nx < 2 || nt < 2 || n_sample < 1 || !alp_dev || !x_nodes_dev || !t_arr_dev || !x_init_dev || !traj_x_dev ||
      !traj_alp_dev || (ndim == 2 && (!y_nodes_dev || ny < 2)) || (egno == 3 && ndim != 2) || n_ctrl < 1 || n_ctrl > 2)
    return PDHG_ERR_ARG;
  pdhg::TrajParams p;
  p.ndim = ndim; p.egno = egno; p.n_ctrl = n_ctrl; p.nx = nx; p.ny = ny; p.nt = nt; p.n_sample = n_sample; p.bc_x = bc_x; p.bc_y = bc_y;
  p.nearest = nearest; p.Px = x_period; p.Py = y_period; p.epsl = epsl; p.alp = alp_dev; p.xa = x_nodes_dev; p.ya = y_nodes_dev;
  p.t_arr = t_arr_dev; p.noise = (epsl > 0.0) ? noise_dev : nullptr; p.x_init = x_init_dev; p.traj_x = traj_x_dev; p.traj_alp = traj_alp_dev;
  if (epsl > 0.0 && !noise_dev) return PDHG_ERR_ARG;
  pdhg::traj_kernel<<<(n_sample + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return cudaGetLastError() == cudaSuccess ? PDHG_OK : PDHG_ERR_CUDA;
}
