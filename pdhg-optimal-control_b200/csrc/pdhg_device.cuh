// Device-side building blocks shared by the PDHG kernels (sm_100a).
//
// Everything here is pointwise / intra-CTA: the Stockham shared-memory FFT, block reductions and the
// proximal-step arithmetic of the reference's update operators.  Reference citations are relative to
// /root/reference (see SURVEY.md Appendix A for the restated formulas).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdhg {

constexpr double kRhoOffset = 1e-4;   // update_fns_in_pdhg.py:74,86; set_fns.py:104,126,151
constexpr int kMaxStages = 16;

// ---------------------------------------------------------------------------------------------------
// Mixed-radix plan: n = prod radix[s].  Radices 2,3,4,5 have hand-written butterflies, any other prime
// factor goes through the generic O(r^2) butterfly.
// ---------------------------------------------------------------------------------------------------
struct FftPlan {
  int n;
  int nstages;
  int radix[kMaxStages];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i*sg  (sg=+1: forward rotation by -i; sg=-1: inverse rotation by +i)
__device__ __forceinline__ double2 rot_mi(double2 a, double sg) { return make_double2(sg * a.y, -sg * a.x); }

// One Stockham autosort stage over `rows` independent rows of length n (row stride `ld` complex elements).
//   src[j + t*n/R] * w^(t*k)  --radix-R DFT-->  dst[(j-k)*R + k + t*Ns],   k = j mod Ns,  w = exp(-+2*pi*i/(Ns*R))
// tw[m] = exp(-2*pi*i*m/n) is the master twiddle table (host-computed); sg = +1 forward, -1 inverse (unnormalised).
__device__ __forceinline__ void fft_stage(const double2* __restrict__ src, double2* __restrict__ dst, int n, int ld,
                                          int R, int Ns, const double2* __restrict__ tw, int rows, double sg,
                                          int tid, int nthreads) {
  const int nb = n / R;
  const int twstep = n / (Ns * R);
  const int total = rows * nb;
  for (int w = tid; w < total; w += nthreads) {
    const int row = w / nb;
    const int j = w - row * nb;
    const int k = j % Ns;
    const double2* s = src + (size_t)row * ld + j;
    double2* d = dst + (size_t)row * ld + (j - k) * R + k;
    if (R == 4) {
      double2 a0 = s[0], a1 = s[nb], a2 = s[2 * nb], a3 = s[3 * nb];
      if (k) {
        double2 w1 = tw[k * twstep], w2 = tw[2 * k * twstep], w3 = tw[3 * k * twstep];
        w1.y *= sg; w2.y *= sg; w3.y *= sg;
        a1 = cmul(a1, w1); a2 = cmul(a2, w2); a3 = cmul(a3, w3);
      }
      double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = rot_mi(csub(a1, a3), sg);
      d[0] = cadd(t0, t2); d[Ns] = cadd(t1, t3); d[2 * Ns] = csub(t0, t2); d[3 * Ns] = csub(t1, t3);
    } else if (R == 2) {
      double2 a0 = s[0], a1 = s[nb];
      if (k) { double2 w1 = tw[k * twstep]; w1.y *= sg; a1 = cmul(a1, w1); }
      d[0] = cadd(a0, a1); d[Ns] = csub(a0, a1);
    } else if (R == 5) {
      double2 a0 = s[0], a1 = s[nb], a2 = s[2 * nb], a3 = s[3 * nb], a4 = s[4 * nb];
      if (k) {
        double2 w1 = tw[k * twstep], w2 = tw[2 * k * twstep], w3 = tw[3 * k * twstep], w4 = tw[4 * k * twstep];
        w1.y *= sg; w2.y *= sg; w3.y *= sg; w4.y *= sg;
        a1 = cmul(a1, w1); a2 = cmul(a2, w2); a3 = cmul(a3, w3); a4 = cmul(a4, w4);
      }
      const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2pi/5), cos(4pi/5)
      const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin(2pi/5), sin(4pi/5)
      double2 p14 = cadd(a1, a4), m14 = csub(a1, a4), p23 = cadd(a2, a3), m23 = csub(a2, a3);
      double2 e1 = make_double2(a0.x + c1 * p14.x + c2 * p23.x, a0.y + c1 * p14.y + c2 * p23.y);
      double2 e2 = make_double2(a0.x + c2 * p14.x + c1 * p23.x, a0.y + c2 * p14.y + c1 * p23.y);
      double2 o1 = rot_mi(make_double2(s1 * m14.x + s2 * m23.x, s1 * m14.y + s2 * m23.y), sg);
      double2 o2 = rot_mi(make_double2(s2 * m14.x - s1 * m23.x, s2 * m14.y - s1 * m23.y), sg);
      d[0] = make_double2(a0.x + p14.x + p23.x, a0.y + p14.y + p23.y);
      d[Ns] = cadd(e1, o1); d[4 * Ns] = csub(e1, o1);
      d[2 * Ns] = cadd(e2, o2); d[3 * Ns] = csub(e2, o2);
    } else if (R == 3) {
      double2 a0 = s[0], a1 = s[nb], a2 = s[2 * nb];
      if (k) {
        double2 w1 = tw[k * twstep], w2 = tw[2 * k * twstep];
        w1.y *= sg; w2.y *= sg;
        a1 = cmul(a1, w1); a2 = cmul(a2, w2);
      }
      const double h = 0.86602540378443864676;   // sqrt(3)/2
      double2 p = cadd(a1, a2), m = csub(a1, a2);
      double2 e = make_double2(a0.x - 0.5 * p.x, a0.y - 0.5 * p.y);
      double2 o = rot_mi(make_double2(h * m.x, h * m.y), sg);
      d[0] = cadd(a0, p); d[Ns] = cadd(e, o); d[2 * Ns] = csub(e, o);
    } else {
      // generic radix (any other prime factor): out[q] = sum_t src[t] * w^(t*k) * W_R^(t*q)
      const int rstep = n / R;
      for (int q = 0; q < R; ++q) {
        double2 acc = make_double2(0.0, 0.0);
        for (int t = 0; t < R; ++t) {
          double2 v = s[t * nb];
          int m = (int)(((long long)t * k * twstep + (long long)((t * q) % R) * rstep) % n);
          double2 ww = tw[m]; ww.y *= sg;
          acc = cadd(acc, cmul(v, ww));
        }
        d[q * Ns] = acc;
      }
    }
  }
}

// Full transform of `rows` rows held in buf0; ping-pongs with buf1; returns the buffer that holds the result.
// All threads of the CTA must call it; it ends with a __syncthreads().
__device__ __forceinline__ double2* fft_rows(double2* buf0, double2* buf1, const FftPlan& plan, int ld,
                                             const double2* __restrict__ tw, int rows, double sg) {
  int Ns = 1;
  double2* a = buf0;
  double2* b = buf1;
  for (int s = 0; s < plan.nstages; ++s) {
    const int R = plan.radix[s];
    fft_stage(a, b, plan.n, ld, R, Ns, tw, rows, sg, threadIdx.x, blockDim.x);
    __syncthreads();
    Ns *= R;
    double2* t = a; a = b; b = t;
  }
  return a;
}

// ---------------------------------------------------------------------------------------------------
// Block reduction of N doubles; every thread gets all N totals.  Fixed order => bitwise reproducible.
// `scratch` needs N * 32 doubles of shared memory.
// ---------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  __syncthreads();   // protects scratch against the previous use
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += scratch[i * 32 + w];
    v[i] = t;
  }
}

// ---------------------------------------------------------------------------------------------------
// Pointwise arithmetic of the dual (proximal) updates.
// ---------------------------------------------------------------------------------------------------
// Reciprocals of the grid constants.  fp64 division costs ~40 DFMA-equivalents on the GPU and the dual sweep has a
// dozen of them per point; multiplying by a reciprocal differs from the reference's division by <= 1 ulp per
// operation (same order as FMA contraction) and keeps the kernels memory-bound instead of divide-bound.
struct Recip {
  double idt, idx, idy, idx2, idy2, isig;
  __device__ __forceinline__ Recip(double dt, double dx, double dy, double sigma)
      : idt(1.0 / dt), idx(1.0 / dx), idy(1.0 / dy), idx2(1.0 / (dx * dx)), idy2(1.0 / (dy * dy)), isig(1.0 / sigma) {}
};

// the one true division of the dual sweep: 1/(1/c_H + p) for egno 1,3 (c_H = 1), 1/p for egno 2
__device__ __forceinline__ double prox_rinv(int egno, double p) { return (egno == 2) ? 1.0 / p : 1.0 / (1.0 + p); }

// alp prox for one upwind copy (set_fns.py:63-77 egno 1, :79-95 egno 2, :100-108 egno 3) followed by the
// upwind mask (set_fns.py:128-138,153-159): keep where f(alp') >= 0 (want_nonneg) or < 0.
//   dphi: one-sided difference of phi_bar; coef: a(x) (egno 1,2); p = (rho+1e-4)/sigma; rinv = prox_rinv(egno, p)
__device__ __forceinline__ double prox_alp(int egno, double alp_prev, double dphi, double p, double rinv, double coef, bool want_nonneg) {
  double v;
  if (egno == 2) {
    v = dphi * coef * rinv + alp_prev;
    v = (v < -1.0) ? -1.0 : v;       // NaN-propagating clip to [-c_H, c_H], c_H = 1
    v = (v > 1.0) ? 1.0 : v;
  } else if (egno == 3) {
    v = (-dphi + p * alp_prev) * rinv;
  } else {
    v = (dphi * coef + p * alp_prev) * rinv;
  }
  const double f = (egno == 3) ? v : -(coef * v);
  const double keep = want_nonneg ? ((f >= 0.0) ? 1.0 : 0.0) : ((f < 0.0) ? 1.0 : 0.0);
  return v * keep;
}

// upwind split of the dynamics (update_fns_in_pdhg.py:22-27,38-47): f*[f>=0] or f*[f<0]
__device__ __forceinline__ double f_plus(double f) { return f * ((f >= 0.0) ? 1.0 : 0.0); }
__device__ __forceinline__ double f_minus(double f) { return f * ((f < 0.0) ? 1.0 : 0.0); }

// running Lagrangian term for one alp component (set_fns.py:32-36): alp^2/c_H/2, or 0*alp for egno 2
__device__ __forceinline__ double lagr(int egno, double a) { return (egno == 2) ? 0.0 * a : a * a * 0.5; }

// NaN-propagating max(v, 0) (jnp.maximum semantics, update_fns_in_pdhg.py:102,118)
__device__ __forceinline__ double relu_nan(double v) { return (v < 0.0) ? 0.0 : v; }

__device__ __forceinline__ bool is_nan(double v) { return v != v; }

}  // namespace pdhg
