// Small helper kernels at the API boundary: initial state of PDHG_multi_step and the conversion between the
// reference's alp layout ([.., n_ctrl] interleaved, with structurally-zero components) and the planar
// active-component layout the solver kernels use.
#include "pdhg_params.h"

namespace pdhg {

// phi0 = tile(g), rho0 = c_on_rho, alp0 = 0   (utils_pdhg_solver.py:123-137)
__global__ void init_state_kernel(const MarchParams p, const double* __restrict__ g, int B) {
  const size_t n = (size_t)p.nx * p.ny;
  const size_t np = (size_t)(p.K + 1) * n, kn = (size_t)p.K * n;
  const int A = 2 * p.ndim;
  const size_t total = (size_t)B * np;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t b = i / np, r = i - b * np;
    p.st_phi[i] = g[b * n + r % n];
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)B * kn; i += stride) p.st_rho[i] = 0.0 + p.c_on_rho;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)B * A * kn; i += stride) p.st_alp[i] = 0.0;
}

cudaError_t launch_init_state(const MarchParams& p, const double* g, int B, cudaStream_t stream) {
  const size_t total = (size_t)B * (p.K + 1) * p.nx * p.ny * 2;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  init_state_kernel<<<blocks, 256, 0, stream>>>(p, g, B);
  return cudaGetLastError();
}

// active component of control array j (0: alp1_x, 1: alp2_x, 2: alp1_y, 3: alp2_y):
// egno 1,2 in 2-D: x-pair uses component 0, y-pair component 1 (set_fns.py:117-118); n_ctrl == 1: component 0.
__device__ __forceinline__ int active_comp(int j, int n_ctrl) { return (n_ctrl == 1) ? 0 : ((j < 2) ? 0 : 1); }

// to_planar=1: planar[b][j][i] = ref[b][j][i][active]; to_planar=0: ref[b][j][i][c] = (c==active) ? planar : 0
__global__ void pack_alp_kernel(double* __restrict__ ref, double* __restrict__ planar, size_t total, int A, size_t kn,
                                int n_ctrl, int to_planar) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int j = (int)((i / kn) % A);
    const int ac = active_comp(j, n_ctrl);
    if (to_planar) {
      planar[i] = ref[i * n_ctrl + ac];
    } else {
      const double v = planar[i];
      for (int c = 0; c < n_ctrl; ++c) ref[i * n_ctrl + c] = (c == ac) ? v : 0.0;
    }
  }
}

cudaError_t launch_pack_alp(const double* ref_layout, double* planar, int B, int A, size_t kn, int n_ctrl, int ndim,
                            int egno, int to_planar, cudaStream_t stream) {
  (void)ndim; (void)egno;
  const size_t total = (size_t)B * A * kn;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  pack_alp_kernel<<<blocks, 256, 0, stream>>>(const_cast<double*>(ref_layout), planar, total, A, kn, n_ctrl, to_planar);
  return cudaGetLastError();
}

}  // namespace pdhg
