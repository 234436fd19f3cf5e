// Cooperative multi-CTA PDHG solver (2-D grids, large 1-D space-time blocks) — placeholder until implemented.
#include "pdhg_params.h"
namespace pdhg {
cudaError_t launch_pdhg_coop(const MarchParams&, int, void*, cudaStream_t, long long*) { return cudaErrorNotSupported; }
size_t pdhg_coop_workspace_bytes(const MarchParams&, int) { return 16; }
cudaError_t launch_update_primal(const MarchParams&, int, const double*, double, double*, void*, cudaStream_t, long long*) { return cudaErrorNotSupported; }
cudaError_t launch_update_dual(const MarchParams&, int, const double*, double, double, int*, void*, cudaStream_t, long long*) { return cudaErrorNotSupported; }
}
