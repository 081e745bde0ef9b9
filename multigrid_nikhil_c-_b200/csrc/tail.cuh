// tail.cuh — the coarse tail (MG_COARSE_TAIL): every level <= kTailMaxLevel of a cycle
// runs in ONE launch of ONE CTA, with u, the ping-pong / residual scratch and f of all
// those levels resident in shared memory (fp64, levels 6..1: 137 KB).  Below ~65^2 a
// level is a few microseconds of launch latency and nothing else (profiles/r01_v1_*:
// ~20 launches of 3-20 us each); here a level visit costs a handful of __syncthreads.
// The cycle recursion of vcyclemultigrid (P:575-627), including gamma > 1, is unrolled
// into an explicit loop with per-level visit counters.  Point formulas and evaluation
// order are those of common.cuh, so results are bit-identical to the per-level kernels.
#pragma once

#include "common.cuh"

namespace mgb {

constexpr int kTailMaxLevel = 6;
constexpr int kTailThreads = 512;

template <typename T>
struct TailArgs {
    int top, coarsest, nu1, nu2, gamma;
    T c0, c1, w;
    T* u;        // level `top`, padded layout (in: current iterate, out: after the cycle)
    const T* f;
    i64 pitch;
};

template <typename T>
inline size_t tail_smem_bytes(int top, int coarsest)
{
    size_t elems = 0;
    for (int l = coarsest; l <= top; ++l) {
        const size_t n1 = ((size_t)1 << l) + 1;
        elems += 3 * n1 * n1;
    }
    return elems * sizeof(T) + 16;
}

template <typename T, bool RBGS>
__global__ void __launch_bounds__(kTailThreads)
k_tail(const TailArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char tail_smem[];
    T* base = reinterpret_cast<T*>(tail_smem);
    // per-level smem arrays (node grid incl. zero ring, pitch N+1): A = current u, B = scratch, F = rhs.
    // Kept as individually named registers through full unrolling over the (at most 6) levels.
    T* A[kTailMaxLevel + 1];
    T* B[kTailMaxLevel + 1];
    T* F[kTailMaxLevel + 1];
    int visits[kTailMaxLevel + 1];
    size_t total = 0;
#pragma unroll
    for (int l = 1; l <= kTailMaxLevel; ++l) {
        const size_t sz = (size_t)((1 << l) + 1) * ((1 << l) + 1);
        const bool on = (l >= a.coarsest) && (l <= a.top);
        A[l] = base + total;
        B[l] = base + total + sz;
        F[l] = base + total + 2 * sz;
        if (on) total += 3 * sz;
        visits[l] = 0;
    }
    A[0] = B[0] = F[0] = base;
    visits[0] = 0;
    for (size_t i = threadIdx.x; i < total; i += kTailThreads) base[i] = (T)0;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NWARP = kTailThreads / 32;

    // load level `top` (batched: all global loads of a thread are in flight before the first smem store)
    {
        const int N = 1 << a.top, P = N + 1;
        constexpr int MAXIT = ((1 << kTailMaxLevel) + NWARP - 1) / NWARP;
        T tu[MAXIT][2], tf[MAXIT][2];
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int y = 1 + warp + it * NWARP;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int x = 1 + lane + 32 * h;
                const bool ok = (y < N) && (x < N);
                tu[it][h] = ok ? a.u[(i64)y * a.pitch + x] : (T)0;
                tf[it][h] = ok ? a.f[(i64)y * a.pitch + x] : (T)0;
            }
        }
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int y = 1 + warp + it * NWARP;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int x = 1 + lane + 32 * h;
                if ((y < N) && (x < N)) {
                    A[a.top][y * P + x] = tu[it][h];
                    F[a.top][y * P + x] = tf[it][h];
                }
            }
        }
    }
    __syncthreads();

    auto smooth = [&](int l, int nu) {
        const int N = 1 << l, P = N + 1;
        for (int s = 0; s < nu; ++s) {
            if (!RBGS) {
                const T* src = A[l];
                T* dst = B[l];
                for (int y = 1 + warp; y < N; y += NWARP)
                    for (int x = 1 + lane; x < N; x += 32) {
                        const int i = y * P + x;
                        dst[i] = jacobi_pt<T>(a.c0, a.c1, src[i], F[l][i], sigma4<T>(src[i - P], src[i + P], src[i - 1], src[i + 1]));
                    }
                __syncthreads();
                T* t = A[l]; A[l] = B[l]; B[l] = t;
            } else {
                for (int colour = 0; colour < 2; ++colour) {
                    T* p = A[l];
                    for (int y = 1 + warp; y < N; y += NWARP)
                        for (int x = 1 + lane; x < N; x += 32) {
                            if (((y + x) & 1) != colour) continue;
                            const int i = y * P + x;
                            p[i] = gs_pt<T>(F[l][i], sigma4<T>(p[i - P], p[i + P], p[i - 1], p[i + 1]));
                        }
                    __syncthreads();
                }
            }
        }
    };

    auto residual_restrict = [&](int l) {
        const int N = 1 << l, P = N + 1;
        T* r = B[l];
        const T* u = A[l];
        for (int y = 1 + warp; y < N; y += NWARP)
            for (int x = 1 + lane; x < N; x += 32) {
                const int i = y * P + x;
                r[i] = resid_pt<T>(u[i], F[l][i], sigma4<T>(u[i - P], u[i + P], u[i - 1], u[i + 1]));
            }
        __syncthreads();
        const int Nc = N >> 1, Pc = Nc + 1;
        for (int I = 1 + warp; I < Nc; I += NWARP)
            for (int J = 1 + lane; J < Nc; J += 32) {
                const int i = (2 * I) * P + 2 * J;
                F[l - 1][I * Pc + J] = fw_pt<T>(a.w, r[i - P - 1], r[i - P + 1], r[i + P - 1], r[i + P + 1],
                                                r[i - 1], r[i + 1], r[i - P], r[i + P], r[i]);
                A[l - 1][I * Pc + J] = (T)0;   // zero coarse guess (P:613)
            }
        __syncthreads();
    };

    auto prolong_correct = [&](int l) {   // fine level l, coarse l-1
        const int N = 1 << l, P = N + 1, Pc = (N >> 1) + 1;
        const T* e = A[l - 1];
        T* u = A[l];
        for (int y = 1 + warp; y < N; y += NWARP)
            for (int x = 1 + lane; x < N; x += 32) {
                const int I = y >> 1, J = x >> 1;
                const T* c = e + I * Pc + J;
                T v;
                if (!(y & 1) && !(x & 1)) v = c[0];                                               // P:401
                else if ((y & 1) && !(x & 1)) v = (T)0.5 * (c[0] + c[Pc]);                        // P:407
                else if (!(y & 1)) v = (T)0.5 * (c[0] + c[1]);                                    // P:413
                else v = (T)0.25 * (((c[0] + c[Pc]) + c[1]) + c[Pc + 1]);                         // P:419
                u[y * P + x] = u[y * P + x] + v;                                                  // P:623
            }
        __syncthreads();
    };

    // vcyclemultigrid P:575-627 as a loop
    int l = a.top;
    bool descending = true;
    while (true) {
        if (descending) {
            smooth(l, a.nu1);                                   // P:581
            if (l <= a.coarsest) {
                smooth(l, a.nu2);                               // P:585
                descending = false;
            } else {
                residual_restrict(l);                           // P:604-613
                visits[l - 1] = (l - 1 <= a.coarsest) ? 1 : (a.gamma < 1 ? 1 : a.gamma);
                l = l - 1;
            }
        } else {
            // level l has just completed one cycle visit
            if (l == a.top) break;
            if (--visits[l] > 0) { descending = true; continue; }   // gamma > 1: cycle again on this level
            prolong_correct(l + 1);                             // P:620-624
            smooth(l + 1, a.nu2);                               // P:625
            l = l + 1;
        }
    }

    // store level `top`
    {
        const int N = 1 << a.top, P = N + 1;
        for (int y = 1 + warp; y < N; y += NWARP)
            for (int x = 1 + lane; x < N; x += 32) a.u[(i64)y * a.pitch + x] = A[a.top][y * P + x];
    }
}

}  // namespace mgb
