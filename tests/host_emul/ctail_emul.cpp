// tests/host_emul/ctail_emul.cpp — TEST INFRASTRUCTURE.  Runs the op list of the cluster coarse tail
// (multigrid_nikhil_c-_b200/csrc/ctail_core.h, the exact code the CUDA kernel k_ctail executes) on the CPU:
// for every op, for every CTA of the cluster, for every thread.  A cluster barrier separates the ops on the
// GPU and an op only reads what earlier ops wrote (or, for red-black, points of the other colour), so running
// CTAs and threads one after the other is an exact emulation.  "Remote shared memory" is simply the other
// CTA's array.
#include <vector>

#include "ctail_core.h"

using namespace mgb;

template <typename T>
struct RemoteHost {
    T* const* bases;
    T* operator()(int rank) const { return bases[rank]; }
};

template <typename T>
static int run(int top, int coarsest, int nu1, int nu2, int gamma, int rbgs, int nctas, int nthr, long long pitch, T* u,
               const T* f, double c0, double c1, double w, int* nops_out)
{
    if (top < 1 || top > kCtailMaxLevel || nctas < 1 || nctas > kCtailMaxCtas || (nctas & (nctas - 1))) return 1;
    std::vector<CtailOp> ops = ctail_schedule(top, coarsest, nu1, nu2, gamma, rbgs != 0);
    const int total = ctail_off(top + 1, nctas);
    std::vector<std::vector<T>> smem(nctas, std::vector<T>(total, (T)0));
    std::vector<T*> bases(nctas);
    for (int c = 0; c < nctas; ++c) bases[c] = smem[c].data();
    CtailArgs<T> a;
    a.top = top; a.nctas = nctas; a.nops = (int)ops.size();
    a.c0 = (T)c0; a.c1 = (T)c1; a.w = (T)w;
    a.u = u; a.f = f; a.pitch = pitch; a.ops = ops.data();
    for (const CtailOp& op : ops)
        for (int c = 0; c < nctas; ++c) {
            CtailEnv<T, RemoteHost<T>> env{c, nctas, bases[c], RemoteHost<T>{bases.data()}};
            for (int tid = 0; tid < nthr; ++tid) ctail_op<T, RemoteHost<T>>(env, op, a, tid, nthr);
        }
    if (nops_out) *nops_out = (int)ops.size();
    return 0;
}

extern "C" {
int ctail_emul_f64(int top, int coarsest, int nu1, int nu2, int gamma, int rbgs, int nctas, int nthr, long long pitch,
                   double* u, const double* f, double c0, double c1, double w, int* nops_out)
{
    return run<double>(top, coarsest, nu1, nu2, gamma, rbgs, nctas, nthr, pitch, u, f, c0, c1, w, nops_out);
}
int ctail_emul_f32(int top, int coarsest, int nu1, int nu2, int gamma, int rbgs, int nctas, int nthr, long long pitch,
                   float* u, const float* f, double c0, double c1, double w, int* nops_out)
{
    return run<float>(top, coarsest, nu1, nu2, gamma, rbgs, nctas, nthr, pitch, u, f, c0, c1, w, nops_out);
}
long long ctail_smem_bytes_f64(int top, int nctas) { return (long long)ctail_smem_bytes<double>(top, nctas); }
long long ctail_smem_bytes_f32(int top, int nctas) { return (long long)ctail_smem_bytes<float>(top, nctas); }
}
