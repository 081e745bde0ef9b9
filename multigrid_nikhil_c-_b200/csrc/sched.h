// sched.h — communication-avoiding V-cycle schedule for row slabs (pure host C++, no CUDA).
//
// The default multi-GPU path exchanges halos lazily (Ctx::ensure_halo): ~11 NCCL exchanges + 1 all-gather
// per V(2,2) with four distributed levels, each a latency-bound ~10-15 us step (profiles/r01_scaling_16385.md).
// Here the fused kernels compute a few rows BEYOND the rows a rank owns, redundantly and bit-identically
// to the neighbour, so that every input of the next kernel is already present:
//   * PRE(l) writes u' and the coarse right-hand side on own rows +- e_l; the zero coarse guess is local;
//   * POST(l) writes on own rows +- x_l, enough for POST(l+1)'s prolongation stencil;
//   * the only exchange left is ONE deep exchange of u (and of f when it is not static) on the top level
//     before the cycle, plus the all-gather at the agglomeration level.
// Extents (NS1 / NS2 = smoothing stages fused into PRE / POST; a stage reaches 1 row, residual +
// restriction 2 more):
//   x_top = 0;  x_{l-1} = ceil((x_l + NS2) / 2) + 1            coarse rows POST(l) interpolates from
//   e_low = x_low + NS2;  e_l = max(x_l + NS2, 2 (e_{l-1} + NS1 + 1))   f_{l-1} must cover PRE(l-1)'s reach
//   top-level input halo: u: e_top + NS1 + 2, f: e_top + NS1 + 1
// With NS1 = NS2 = 2 and levels 14..11 distributed: e = 90 / 42 / 18 / 6 rows: < 5 % extra rows on 2048.
//
// The schedule is DATA (a list of ops) produced here and executed by (a) fused.cu on the GPU and (b) the CPU
// emulator in tests/test_sched_comm_avoid.py, which runs the oracle on NaN-poisoned slabs: any row this plan
// fails to provide shows up as NaN in an owned row.
#pragma once

#include <algorithm>
#include <vector>

namespace mgb {

enum SchedKind {
    SCHED_EXCH = 0,       // a = which (0 u, 1 f), b = depth
    SCHED_PRE = 1,        // a = ya, b = yb  (output rows; coarse rows with centre in [ya, yb) are written too)
    SCHED_POST = 2,       // a = ya, b = yb
    SCHED_GATHER_F = 3,   // all-gather the right-hand side of `level` (first replicated level), zero its u
    SCHED_REPL_CYCLE = 4  // ordinary cycle on the replicated level `level`
};

struct SchedOp {
    int kind, level, a, b;
};

struct SchedPlan {
    std::vector<SchedOp> ops;
    int halo[32];    // stored halo rows needed per level (0 for replicated levels)
    int x[32], e[32];
    bool ok = false;
};

inline void sched_own_rows(int level, int rank, int world, int* lo, int* hi)
{
    const long long N = 1ll << level;
    long long a = (long long)rank * N / world, b = (long long)(rank + 1) * N / world;
    if (a < 1) a = 1;
    if (rank == world - 1) b = N;
    *lo = (int)a;
    *hi = (int)b;
}

// extents only (used at context creation to size the stored halos)
inline bool sched_extents(int top, int aggl, int ns1, int ns2, int* x, int* e, int* halo)
{
    for (int l = 0; l < 32; ++l) x[l] = e[l] = halo[l] = 0;
    if (top <= aggl) return false;
    x[top] = 0;
    for (int l = top; l > aggl + 1; --l) x[l - 1] = (x[l] + ns2 + 1) / 2 + 1;
    e[aggl + 1] = x[aggl + 1] + ns2;
    for (int l = aggl + 2; l <= top; ++l) e[l] = std::max(x[l] + ns2, 2 * (e[l - 1] + ns1 + 1));
    for (int l = aggl + 1; l <= top; ++l) halo[l] = e[l] + ns1 + 2;
    return true;
}

// hv_u_top / hv_f_top: valid halo rows of u / f on the top level when the cycle starts.
inline SchedPlan sched_plan_vcycle(int top, int aggl, int world, int rank, int ns1, int ns2, int hv_u_top, int hv_f_top)
{
    SchedPlan p;
    if (world < 2 || top <= aggl || ns1 < 1 || ns2 < 1) return p;
    if (!sched_extents(top, aggl, ns1, ns2, p.x, p.e, p.halo)) return p;
    // feasibility: a neighbour must own every row it is asked to send, on every distributed level
    for (int l = aggl + 1; l <= top; ++l) {
        const long long rows = (1ll << l) / world;
        if (p.halo[l] > rows - 1) return p;
    }
    const bool up = rank > 0, dn = rank < world - 1;
    auto range = [&](int l, int ext, int* ya, int* yb) {
        int lo, hi;
        sched_own_rows(l, rank, world, &lo, &hi);
        const int N = 1 << l;
        *ya = std::max(1, lo - (up ? ext : 0));
        *yb = std::min(N, hi + (dn ? ext : 0));
    };
    const int need_u = p.e[top] + ns1 + 2, need_f = p.e[top] + ns1 + 1;
    if (hv_u_top < need_u) p.ops.push_back({SCHED_EXCH, top, 0, need_u});
    if (hv_f_top < need_f) p.ops.push_back({SCHED_EXCH, top, 1, need_f});
    for (int l = top; l > aggl; --l) {
        int ya, yb;
        range(l, p.e[l], &ya, &yb);
        p.ops.push_back({SCHED_PRE, l, ya, yb});
        if (l - 1 == aggl) p.ops.push_back({SCHED_GATHER_F, aggl, 0, 0});
    }
    p.ops.push_back({SCHED_REPL_CYCLE, aggl, 0, 0});
    for (int l = aggl + 1; l <= top; ++l) {
        int ya, yb;
        range(l, p.x[l], &ya, &yb);
        p.ops.push_back({SCHED_POST, l, ya, yb});
    }
    p.ok = true;
    return p;
}

}  // namespace mgb
