// fused.cu — fused / temporally blocked kernels.  (Filled in after the unfused path is parity-green.)
#include "fused.cuh"

namespace mgb {

void fused_setup(Ctx&) {}
template <typename T> int fused_jacobi(Ctx&, Level&, int, T, T) { return 0; }
template int fused_jacobi<double>(Ctx&, Level&, int, double, double);
template int fused_jacobi<float>(Ctx&, Level&, int, float, float);
bool fused_cycle_level(Ctx&, int, int, int, int) { return false; }
bool fused_time_hook(Ctx&, int, bool) { return false; }

}  // namespace mgb
