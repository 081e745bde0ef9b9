// tail.cuh — the coarse tail (MG_COARSE_TAIL): every level <= kTailMaxLevel of a cycle
// runs in ONE launch of ONE CTA, with u, the ping-pong / residual scratch and f of all
// those levels resident in shared memory (fp64, levels 6..1: 137 KB).  Below ~65^2 a
// level is a few microseconds of launch latency and nothing else (profiles/r01_v1_*:
// ~20 launches of 3-20 us each); here a level visit costs a handful of __syncthreads.
// The cycle recursion of vcyclemultigrid (P:575-627), including gamma > 1, is a template
// recursion over the level, so every loop bound, pitch and shared-memory offset is a
// compile-time constant.  Point formulas and evaluation order are those of common.cuh, so
// results are bit-identical to the per-level kernels.
#pragma once

#include "common.cuh"

namespace mgb {

constexpr int kTailMaxLevel = 6;
constexpr int kTailThreads = 1024;
constexpr int kTailWarps = kTailThreads / 32;

template <typename T>
struct TailArgs {
    int top, coarsest, nu1, nu2, gamma;
    T c0, c1, w;
    T* u;        // level `top`, padded layout (in: current iterate, out: after the cycle)
    const T* f;
    i64 pitch;
};

// shared-memory element offset of level L's three arrays (levels are laid out from level 1)
__host__ __device__ constexpr int tail_off(int L)
{
    int off = 0;
    for (int l = 1; l < L; ++l) off += 3 * ((1 << l) + 1) * ((1 << l) + 1);
    return off;
}

template <typename T>
inline size_t tail_smem_bytes(int top, int /*coarsest*/)
{
    return (size_t)tail_off(top + 1) * sizeof(T) + 16;
}

template <typename T, int L>
struct TailLv {
    static constexpr int N = 1 << L, P = N + 1, SZ = P * P;
    static constexpr int YIT = (N - 1 + kTailWarps - 1) / kTailWarps;   // row iterations per warp
    static constexpr int XIT = (N - 1 + 31) / 32;                       // column iterations per lane
    T* buf;      // [u ping][u pong / residual scratch][f]
    __device__ __forceinline__ TailLv(T* base) : buf(base + tail_off(L)) {}
    __device__ __forceinline__ T* A(unsigned cur) const { return buf + (((cur >> L) & 1u) ? SZ : 0); }
    __device__ __forceinline__ T* B(unsigned cur) const { return buf + (((cur >> L) & 1u) ? 0 : SZ); }
    __device__ __forceinline__ T* F() const { return buf + 2 * SZ; }
};

#define MG_TAIL_FOR_POINTS(LV)                                            \
    _Pragma("unroll") for (int it_ = 0; it_ < LV::YIT; ++it_)             \
    _Pragma("unroll") for (int h_ = 0; h_ < LV::XIT; ++h_)                \
        if (const int y = 1 + warp + it_ * kTailWarps; y < LV::N)         \
            if (const int x = 1 + lane + 32 * h_; x < LV::N)

template <typename T, bool RBGS, int L>
__device__ __forceinline__ void tail_smooth(T* base, const TailArgs<T>& a, unsigned& cur, int nu, int warp, int lane)
{
    typedef TailLv<T, L> LV;
    LV lv(base);
    constexpr int P = LV::P;
    const T* F = lv.F();
    for (int s = 0; s < nu; ++s) {
        if (!RBGS) {
            const T* src = lv.A(cur);
            T* dst = lv.B(cur);
            MG_TAIL_FOR_POINTS(LV)
            {
                const int i = y * P + x;
                dst[i] = jacobi_pt<T>(a.c0, a.c1, src[i], F[i], sigma4<T>(src[i - P], src[i + P], src[i - 1], src[i + 1]));
            }
            __syncthreads();
            cur ^= (1u << L);
        } else {
            T* p = lv.A(cur);
#pragma unroll
            for (int colour = 0; colour < 2; ++colour) {
                MG_TAIL_FOR_POINTS(LV)
                {
                    if (((y + x) & 1) == colour) {
                        const int i = y * P + x;
                        p[i] = gs_pt<T>(F[i], sigma4<T>(p[i - P], p[i + P], p[i - 1], p[i + 1]));
                    }
                }
                __syncthreads();
            }
        }
    }
}

// one cycle visit of level L (vcyclemultigrid P:575-627)
template <typename T, bool RBGS, int L>
__device__ __noinline__ void tail_visit(T* base, const TailArgs<T>& a, unsigned& cur, int warp, int lane)
{
    typedef TailLv<T, L> LV;
    tail_smooth<T, RBGS, L>(base, a, cur, a.nu1, warp, lane);                 // P:581
    if (L <= a.coarsest) {
        tail_smooth<T, RBGS, L>(base, a, cur, a.nu2, warp, lane);             // P:585
        return;
    }
    if constexpr (L > 1) {
        typedef TailLv<T, L - 1> LC;
        LV lv(base);
        LC lc(base);
        constexpr int P = LV::P, Pc = LC::P;
        {   // residual (P:604-607) into the scratch buffer, then full weighting (P:611) + zero guess (P:613)
            const T* u = lv.A(cur);
            const T* F = lv.F();
            T* r = lv.B(cur);
            MG_TAIL_FOR_POINTS(LV)
            {
                const int i = y * P + x;
                r[i] = resid_pt<T>(u[i], F[i], sigma4<T>(u[i - P], u[i + P], u[i - 1], u[i + 1]));
            }
            __syncthreads();
            cur &= ~(1u << (L - 1));
            T* fc = lc.F();
            T* uc = lc.A(cur);
            MG_TAIL_FOR_POINTS(LC)
            {
                const int i = (2 * y) * P + 2 * x;
                fc[y * Pc + x] = fw_pt<T>(a.w, r[i - P - 1], r[i - P + 1], r[i + P - 1], r[i + P + 1],
                                          r[i - 1], r[i + 1], r[i - P], r[i + P], r[i]);
                uc[y * Pc + x] = (T)0;
            }
            __syncthreads();
        }
        const int reps = (L - 1 <= a.coarsest) ? 1 : (a.gamma < 1 ? 1 : a.gamma);
        for (int g = 0; g < reps; ++g) tail_visit<T, RBGS, L - 1>(base, a, cur, warp, lane);   // P:617
        {   // prolongation + correction (P:620-624)
            const T* e = lc.A(cur);
            T* u = lv.A(cur);
            MG_TAIL_FOR_POINTS(LV)
            {
                const T* c = e + (y >> 1) * Pc + (x >> 1);
                T v;
                if (!(y & 1) && !(x & 1)) v = c[0];                                               // P:401
                else if ((y & 1) && !(x & 1)) v = (T)0.5 * (c[0] + c[Pc]);                        // P:407
                else if (!(y & 1)) v = (T)0.5 * (c[0] + c[1]);                                    // P:413
                else v = (T)0.25 * (((c[0] + c[Pc]) + c[1]) + c[Pc + 1]);                         // P:419
                u[y * P + x] = u[y * P + x] + v;                                                  // P:623
            }
            __syncthreads();
        }
        tail_smooth<T, RBGS, L>(base, a, cur, a.nu2, warp, lane);             // P:625
    }
}

template <typename T, bool RBGS, int TOP, bool ZG = false>
__device__ __forceinline__ void tail_run(T* base, const TailArgs<T>& a, int warp, int lane)
{
    typedef TailLv<T, TOP> LV;
    LV lv(base);
    constexpr int P = LV::P;
    unsigned cur = 0;
    {   // load level TOP: all global loads of a thread are in flight before the first smem store
        T tu[LV::YIT][LV::XIT], tf[LV::YIT][LV::XIT];
#pragma unroll
        for (int it = 0; it < LV::YIT; ++it)
#pragma unroll
            for (int h = 0; h < LV::XIT; ++h) {
                const int y = 1 + warp + it * kTailWarps, x = 1 + lane + 32 * h;
                const bool ok = (y < LV::N) && (x < LV::N);
                tu[it][h] = (ok && !ZG) ? a.u[(i64)y * a.pitch + x] : (T)0;
                tf[it][h] = ok ? a.f[(i64)y * a.pitch + x] : (T)0;
            }
        T* A = lv.A(cur);
        T* F = lv.F();
#pragma unroll
        for (int it = 0; it < LV::YIT; ++it)
#pragma unroll
            for (int h = 0; h < LV::XIT; ++h) {
                const int y = 1 + warp + it * kTailWarps, x = 1 + lane + 32 * h;
                if ((y < LV::N) && (x < LV::N)) {
                    A[y * P + x] = tu[it][h];
                    F[y * P + x] = tf[it][h];
                }
            }
    }
    __syncthreads();
    tail_visit<T, RBGS, TOP>(base, a, cur, warp, lane);
    {
        const T* A = lv.A(cur);
        MG_TAIL_FOR_POINTS(LV) { a.u[(i64)y * a.pitch + x] = A[y * P + x]; }
    }
}

template <typename T, bool RBGS>
__global__ void __launch_bounds__(kTailThreads)
k_tail(const TailArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char tail_smem[];
    T* base = reinterpret_cast<T*>(tail_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = tail_off(a.top + 1);
    for (int i = threadIdx.x; i < total; i += kTailThreads) base[i] = (T)0;   // zero rings (and everything else)
    __syncthreads();
    switch (a.top) {
        case 1: tail_run<T, RBGS, 1>(base, a, warp, lane); break;
        case 2: tail_run<T, RBGS, 2>(base, a, warp, lane); break;
        case 3: tail_run<T, RBGS, 3>(base, a, warp, lane); break;
        case 4: tail_run<T, RBGS, 4>(base, a, warp, lane); break;
        case 5: tail_run<T, RBGS, 5>(base, a, warp, lane); break;
        default: tail_run<T, RBGS, 6>(base, a, warp, lane); break;
    }
}

// zero-guess variant (opt-in, MGB200_ZERO_GUESS=1): u of the top level is known to be zero and is not read.
// A separate kernel so that k_tail stays byte-identical to the GPU-verified build.
template <typename T, bool RBGS>
__global__ void __launch_bounds__(kTailThreads)
k_tail_zg(const TailArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char tail_smem[];
    T* base = reinterpret_cast<T*>(tail_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = tail_off(a.top + 1);
    for (int i = threadIdx.x; i < total; i += kTailThreads) base[i] = (T)0;
    __syncthreads();
    switch (a.top) {
        case 1: tail_run<T, RBGS, 1, true>(base, a, warp, lane); break;
        case 2: tail_run<T, RBGS, 2, true>(base, a, warp, lane); break;
        case 3: tail_run<T, RBGS, 3, true>(base, a, warp, lane); break;
        case 4: tail_run<T, RBGS, 4, true>(base, a, warp, lane); break;
        case 5: tail_run<T, RBGS, 5, true>(base, a, warp, lane); break;
        default: tail_run<T, RBGS, 6, true>(base, a, warp, lane); break;
    }
}

}  // namespace mgb
