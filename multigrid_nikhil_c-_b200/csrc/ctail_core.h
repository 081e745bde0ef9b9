// ctail_core.h — CLUSTER coarse tail: every level <= 8 (257^2) of a cycle in ONE launch of ONE thread-block
// cluster, all arrays resident in the cluster's distributed shared memory (DSMEM).  Host-device core.
//
// The single-CTA tail (tail.cuh) stops at level 6 because level 7 does not fit 227 KB; levels 7 and 8 then cost
// two latency-bound launches each (and a W-cycle visits them 2^k times).  A cluster of C CTAs has C x 227 KB:
// C = 16 holds levels <= 8 in fp64 (2.1 MB), C = 4 holds levels <= 7.  Each CTA owns a block of rows of every
// level (the same proportional partition on every level, so coarse row I and fine row 2I mostly share an
// owner); a stencil that needs a row of another CTA reads it straight from that CTA's shared memory.
//
// Like the tile kernels this is written as PHASES separated by barriers (here: cluster barriers) in plain C++:
// the CUDA kernel (ctail.cuh) executes a host-generated op list, one op per phase,
//     for (i = 0; i < nops; ++i) { ctail_op(env, ops[i], ...); cluster.sync(); }
// and tests/host_emul/ctail_emul.cpp runs the same ops CTA by CTA, thread by thread, on the CPU, where the
// result is compared bit for bit with the oracle.  The cycle recursion (any nu1, nu2, gamma) is unrolled into
// the op list on the host, so control flow is trivially uniform across the cluster.
#pragma once

#include "formulas.h"

namespace mgb {

constexpr int kCtailMaxLevel = 8;
constexpr int kCtailMaxCtas = 16;

enum CtailKind {
    CT_LOAD = 0,      // global u, f of `level` -> buffers (a = u buffer id)
    CT_STORE = 1,     // buffer a of `level` -> global u
    CT_JACOBI = 2,    // a = source buffer, b = destination buffer
    CT_RBGS = 3,      // a = buffer (in place), b = colour
    CT_RESID = 4,     // a = u buffer, b = residual buffer
    CT_RESTRICT = 5,  // fine `level`: a = residual buffer; writes f of level-1 and zeroes u buffer 0 of level-1
    CT_PROLONG = 6    // fine `level`: a = fine u buffer (updated in place), b = coarse u buffer
};

struct CtailOp {
    int kind, level, a, b;
};

template <typename T>
struct CtailArgs {
    int top, nctas, nops;
    T c0, c1, w;
    T* u;            // level `top`, padded device layout
    const T* f;
    i64 pitch;
    const CtailOp* ops;
};

// ---- row partition: CTA c owns node rows [row_lo(c), row_hi(c)) of a level with N = 2^level ----
MG_HD int ctail_active(int level, int nctas) { const int N = 1 << level; return N < nctas ? N : nctas; }
MG_HD int ctail_row_lo(int level, int nctas, int c)
{
    const int N = 1 << level, ca = ctail_active(level, nctas);
    return c >= ca ? N + 1 : (int)((i64)c * N / ca);
}
MG_HD int ctail_row_hi(int level, int nctas, int c)
{
    const int N = 1 << level, ca = ctail_active(level, nctas);
    if (c >= ca) return N + 1;
    return c == ca - 1 ? N + 1 : (int)((i64)(c + 1) * N / ca);
}
MG_HD int ctail_owner(int level, int nctas, int y)
{
    const int N = 1 << level, ca = ctail_active(level, nctas);
    const int c = (int)((i64)y * ca / N);
    return c >= ca ? ca - 1 : c;
}
// rows reserved per CTA for a level (the last active CTA also holds the ring row N)
MG_HD int ctail_rows_cap(int level, int nctas)
{
    const int N = 1 << level, ca = ctail_active(level, nctas);
    return (N + ca - 1) / ca + 1;
}
// element offset of level `level`'s three buffers inside one CTA's shared memory (levels laid out from 1)
MG_HD int ctail_off(int level, int nctas)
{
    int off = 0;
    for (int l = 1; l < level; ++l) off += 3 * ctail_rows_cap(l, nctas) * ((1 << l) + 1);
    return off;
}
MG_HD int ctail_buf_elems(int level, int nctas) { return ctail_rows_cap(level, nctas) * ((1 << level) + 1); }

// Environment of one CTA: its rank, its shared memory and a way to reach another CTA's shared memory.
// Device: Remote = cooperative_groups cluster map_shared_rank; host emulation: a table of base pointers.
template <typename T, typename Remote>
struct CtailEnv {
    int me, nctas;
    T* smem;         // this CTA's shared memory
    Remote remote;   // remote(rank) -> base of rank's shared memory (same layout in every CTA)
};

// Geometry of one level inside a CTA's shared memory, computed once per op (not per point).
template <typename T, typename Remote>
struct CtailLevel {
    const CtailEnv<T, Remote>& env;
    int N, P, ca, rows_per, rp_shift, off, be;
    MG_HD CtailLevel(const CtailEnv<T, Remote>& e, int level) : env(e)
    {
        N = 1 << level;
        P = N + 1;
        ca = ctail_active(level, e.nctas);
        rows_per = N / ca;                       // power of two
        rp_shift = 0;
        while ((1 << rp_shift) < rows_per) ++rp_shift;
        off = ctail_off(level, e.nctas);
        be = ctail_buf_elems(level, e.nctas);
    }
    MG_HD int owner(int y) const { const int c = y >> rp_shift; return c >= ca ? ca - 1 : c; }
    // row y (node row 0..N) of buffer buf (0/1 = u ping / pong-or-residual, 2 = f), wherever it lives
    MG_HD const T* row(int buf, int y) const
    {
        const int c = owner(y);
        return (c == env.me ? env.smem : env.remote(c)) + off + buf * be + (y - (c << rp_shift)) * P;
    }
    MG_HD T* my_row(int buf, int y) const { return env.smem + off + buf * be + (y - (env.me << rp_shift)) * P; }
    MG_HD int lo() const { return env.me >= ca ? N + 1 : env.me * rows_per; }
    MG_HD int hi() const { return env.me >= ca ? N + 1 : (env.me == ca - 1 ? N + 1 : (env.me + 1) * rows_per); }
};

// One op (= one phase) executed by thread `tid` of `nthr` of CTA env.me.  Threads are arranged as rows of
// min(32, nthr) lanes: a "warp" takes a row, its lanes stride over the columns.
template <typename T, typename Remote>
MG_HD void ctail_op(const CtailEnv<T, Remote>& env, const CtailOp& op, const CtailArgs<T>& a, int tid, int nthr)
{
    const int lw = nthr < 32 ? nthr : 32, wid = tid / lw, lane = tid % lw, nw = nthr / lw;
    if (wid >= nw) return;
    const CtailLevel<T, Remote> L(env, op.level);
    const int N = L.N;
    const int ya = L.lo() < 1 ? 1 : L.lo(), yb = L.hi() > N ? N : L.hi();   // interior rows this CTA owns

    switch (op.kind) {
        case CT_LOAD: {
            for (int y = ya + wid; y < yb; y += nw) {
                T* u = L.my_row(op.a, y);
                T* f = L.my_row(2, y);
                for (int x = 1 + lane; x < N; x += lw) {
                    u[x] = a.u[(i64)y * a.pitch + x];
                    f[x] = a.f[(i64)y * a.pitch + x];
                }
            }
            break;
        }
        case CT_STORE: {
            for (int y = ya + wid; y < yb; y += nw) {
                const T* u = L.my_row(op.a, y);
                for (int x = 1 + lane; x < N; x += lw) a.u[(i64)y * a.pitch + x] = u[x];
            }
            break;
        }
        case CT_JACOBI: {
            for (int y = ya + wid; y < yb; y += nw) {
                const T* up = L.row(op.a, y - 1);
                const T* ce = L.my_row(op.a, y);
                const T* dn = L.row(op.a, y + 1);
                const T* f = L.my_row(2, y);
                T* o = L.my_row(op.b, y);
                for (int x = 1 + lane; x < N; x += lw)
                    o[x] = jacobi_pt<T>(a.c0, a.c1, ce[x], f[x], sigma4<T>(up[x], dn[x], ce[x - 1], ce[x + 1]));
            }
            break;
        }
        case CT_RBGS: {
            for (int y = ya + wid; y < yb; y += nw) {
                const T* up = L.row(op.a, y - 1);
                T* ce = L.my_row(op.a, y);
                const T* dn = L.row(op.a, y + 1);
                const T* f = L.my_row(2, y);
                for (int x = 1 + lane; x < N; x += lw)
                    if (((y + x) & 1) == op.b) ce[x] = gs_pt<T>(f[x], sigma4<T>(up[x], dn[x], ce[x - 1], ce[x + 1]));
            }
            break;
        }
        case CT_RESID: {
            for (int y = ya + wid; y < yb; y += nw) {
                const T* up = L.row(op.a, y - 1);
                const T* ce = L.my_row(op.a, y);
                const T* dn = L.row(op.a, y + 1);
                const T* f = L.my_row(2, y);
                T* o = L.my_row(op.b, y);
                for (int x = 1 + lane; x < N; x += lw)
                    o[x] = resid_pt<T>(ce[x], f[x], sigma4<T>(up[x], dn[x], ce[x - 1], ce[x + 1]));
            }
            break;
        }
        case CT_RESTRICT: {
            // coarse rows this CTA owns on level-1
            const CtailLevel<T, Remote> C(env, op.level - 1);
            const int Nc = C.N;
            const int Ia = C.lo() < 1 ? 1 : C.lo(), Ib = C.hi() > Nc ? Nc : C.hi();
            for (int I = Ia + wid; I < Ib; I += nw) {
                const T* rn = L.row(op.a, 2 * I - 1);
                const T* rc = L.row(op.a, 2 * I);
                const T* rs = L.row(op.a, 2 * I + 1);
                T* fc = C.my_row(2, I);
                T* uc = C.my_row(0, I);
                for (int J = 1 + lane; J < Nc; J += lw) {
                    const int x = 2 * J;
                    fc[J] = fw_pt<T>(a.w, rn[x - 1], rn[x + 1], rs[x - 1], rs[x + 1], rc[x - 1], rc[x + 1], rn[x], rs[x], rc[x]);
                    uc[J] = (T)0;                          // zero coarse guess (P:613) into buffer 0
                }
            }
            break;
        }
        case CT_PROLONG: {
            const CtailLevel<T, Remote> C(env, op.level - 1);
            for (int y = ya + wid; y < yb; y += nw) {
                const int I = y >> 1;
                const T* c0r = C.row(op.b, I);
                const T* c1r = (y & 1) ? C.row(op.b, I + 1) : c0r;
                T* u = L.my_row(op.a, y);
                for (int x = 1 + lane; x < N; x += lw) {
                    const int J = x >> 1;
                    const T c00 = c0r[J];
                    const T c01 = (x & 1) ? c0r[J + 1] : (T)0;
                    const T c10 = (y & 1) ? c1r[J] : (T)0;
                    const T c11 = ((y & 1) && (x & 1)) ? c1r[J + 1] : (T)0;
                    u[x] = u[x] + prolong_pt<T>(y, x, c00, c10, c01, c11);   // P:623
                }
            }
            break;
        }
    }
}

}  // namespace mgb

// ------------------------------------------------------------------------------------------------
// host side: the op list of one cycle visit of level `top` (vcyclemultigrid P:575-627 unrolled)
// ------------------------------------------------------------------------------------------------
#include <vector>
namespace mgb {

struct CtailSchedule {
    std::vector<CtailOp> ops;
    int cur[kCtailMaxLevel + 1];   // which buffer (0/1) holds u of each level while the list is being built
};

inline void ctail_emit_smooth(CtailSchedule& s, int l, int nu, bool rbgs)
{
    for (int k = 0; k < nu; ++k) {
        if (!rbgs) {
            s.ops.push_back({CT_JACOBI, l, s.cur[l], s.cur[l] ^ 1});
            s.cur[l] ^= 1;
        } else {
            s.ops.push_back({CT_RBGS, l, s.cur[l], 0});
            s.ops.push_back({CT_RBGS, l, s.cur[l], 1});
        }
    }
}

inline void ctail_emit_visit(CtailSchedule& s, int l, int coarsest, int nu1, int nu2, int gamma, bool rbgs)
{
    ctail_emit_smooth(s, l, nu1, rbgs);                                   // P:581
    if (l <= coarsest) {
        ctail_emit_smooth(s, l, nu2, rbgs);                               // P:585
        return;
    }
    s.ops.push_back({CT_RESID, l, s.cur[l], s.cur[l] ^ 1});               // P:604-608 (scratch = the other u buffer)
    s.ops.push_back({CT_RESTRICT, l, s.cur[l] ^ 1, 0});                   // P:611, P:613
    s.cur[l - 1] = 0;
    const int reps = (l - 1 <= coarsest) ? 1 : (gamma < 1 ? 1 : gamma);
    for (int g = 0; g < reps; ++g) ctail_emit_visit(s, l - 1, coarsest, nu1, nu2, gamma, rbgs);   // P:617
    s.ops.push_back({CT_PROLONG, l, s.cur[l], s.cur[l - 1]});             // P:620-624
    ctail_emit_smooth(s, l, nu2, rbgs);                                   // P:625
}

inline std::vector<CtailOp> ctail_schedule(int top, int coarsest, int nu1, int nu2, int gamma, bool rbgs)
{
    CtailSchedule s;
    for (int l = 0; l <= kCtailMaxLevel; ++l) s.cur[l] = 0;
    s.ops.push_back({CT_LOAD, top, 0, 0});
    ctail_emit_visit(s, top, coarsest, nu1, nu2, gamma, rbgs);
    s.ops.push_back({CT_STORE, top, s.cur[top], 0});
    return s.ops;
}

template <typename T>
inline size_t ctail_smem_bytes(int top, int nctas)
{
    return (size_t)ctail_off(top + 1, nctas) * sizeof(T) + 16;
}

}  // namespace mgb
