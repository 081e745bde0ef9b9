// fused.cuh — fused / temporally blocked kernels (MG_FUSED) and the shared-memory
// coarse tail (MG_COARSE_TAIL).  Implemented in fused.cu.
#pragma once

#include "ctx.cuh"

namespace mgb {

void fused_setup(Ctx& ctx);
// run up to `remaining` temporally blocked Jacobi sweeps on level lv; returns the number done (0 = not applicable)
template <typename T> int fused_jacobi(Ctx& ctx, Level& lv, int remaining, T c0, T c1);
// `nu` red-black Gauss-Seidel sweeps through the streaming kernel (both colours per pass); false = not applicable
bool fused_rbgs(Ctx& ctx, Level& lv, int nu);
// run one whole cycle visit of `level` with fused kernels; false = caller runs the unfused sequence
bool fused_cycle_level(Ctx& ctx, int level, int nu1, int nu2, int gamma);
// `visits` (>= 2) consecutive visits of `level` with POST of visit v and PRE of visit v+1 fused into one POSTPRE launch
// (default; MGB200_CHAIN=0 turns it off); false = not applicable here, the caller runs the visits one by one
bool fused_cycle_chain(Ctx& ctx, int level, int nu1, int nu2, int gamma, int visits);
// choose chunk heights for the big levels before a cycle graph is captured
void fused_pretune(Ctx& ctx, int level, int nu1, int nu2, int gamma);
// one launch of the fused pre- (true) or post-smoothing (false) kernel for timing; false if unavailable
bool fused_time_hook(Ctx& ctx, int level, bool pre);
// one launch of the POST+PRE chain kernel (nu2 = nu1 = 2 Jacobi / 1 RB-GS) for timing; false if unavailable
bool fused_time_hook_postpre(Ctx& ctx, int level);
// micro-benchmark only: four weighted-Jacobi sweeps temporally blocked in one launch (not used by the cycles)
bool fused_time_sweeps4(Ctx& ctx, int level);

}  // namespace mgb
