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

// arrays of levels 1..top, and one word after the arrays of ALL tail levels (warp 0 hands its buffer parities back there)
template <typename T>
inline size_t tail_smem_bytes(int /*top*/, int /*coarsest*/)
{
    return (size_t)tail_off(kTailMaxLevel + 1) * sizeof(T) + 16;
}

template <typename T, int L>
struct TailLv {
    static constexpr int N = 1 << L, P = N + 1, SZ = P * P;
    static constexpr int YIT = (N - 1 + kTailWarps - 1) / kTailWarps;   // row iterations per warp
    static constexpr int XIT = (N - 1 + 31) / 32;                       // column iterations per lane
    static constexpr int WIT = ((N - 1) * (N - 1) + 31) / 32;           // point iterations per lane in single-warp mode
    T* buf;      // [u ping][u pong / residual scratch][f]
    __device__ __forceinline__ TailLv(T* base) : buf(base + tail_off(L)) {}
    __device__ __forceinline__ T* A(unsigned cur) const { return buf + (((cur >> L) & 1u) ? SZ : 0); }
    __device__ __forceinline__ T* B(unsigned cur) const { return buf + (((cur >> L) & 1u) ? 0 : SZ); }
    __device__ __forceinline__ T* F() const { return buf + 2 * SZ; }
};

// Levels <= kTailWarpLevel are run by WARP 0 ALONE with __syncwarp() between the phases: a phase on a 7^2 (or smaller)
// grid is a few instructions per lane, and a 1024-thread block barrier costs more than the phase itself (ncu, round 2:
// barrier + instruction-fetch stalls dominate k_tail; a W-cycle visits these levels 40 times per tail launch).  The other
// 31 warps wait at one block barrier for the whole sub-cycle.  Same arrays, same formulas, same bits.
#ifndef MGB_TAIL_WARP_LEVEL
#define MGB_TAIL_WARP_LEVEL 3   // measured on the B200 (profiles/r02_tail_variants.txt): 3 beats 0 (off) and 4
#endif
constexpr int kTailWarpLevel = MGB_TAIL_WARP_LEVEL;

template <bool WARP> __device__ __forceinline__ void tail_sync()
{
    if constexpr (WARP) __syncwarp();
    else __syncthreads();
}

// the interior points of level LV, spread over the whole block (one row per warp) or over the lanes of warp 0
#define MG_TAIL_FOR_POINTS(LV, WARP)                                                                      \
    _Pragma("unroll") for (int it_ = 0; it_ < ((WARP) ? LV::WIT : LV::YIT * LV::XIT); ++it_)              \
        if (const int y = (WARP) ? 1 + (lane + 32 * it_) / (LV::N - 1) : 1 + warp + (it_ / LV::XIT) * kTailWarps; y < LV::N)   \
            if (const int x = (WARP) ? 1 + (lane + 32 * it_) % (LV::N - 1) : 1 + lane + 32 * (it_ % LV::XIT); x < LV::N)

template <typename T, bool RBGS, int L, bool WARP>
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
            MG_TAIL_FOR_POINTS(LV, WARP)
            {
                const int i = y * P + x;
                dst[i] = jacobi_pt<T>(a.c0, a.c1, src[i], F[i], sigma4<T>(src[i - P], src[i + P], src[i - 1], src[i + 1]));
            }
            tail_sync<WARP>();
            cur ^= (1u << L);
        } else {
            T* p = lv.A(cur);
#pragma unroll
            for (int colour = 0; colour < 2; ++colour) {
                MG_TAIL_FOR_POINTS(LV, WARP)
                {
                    if (((y + x) & 1) == colour) {
                        const int i = y * P + x;
                        p[i] = gs_pt<T>(F[i], sigma4<T>(p[i - P], p[i + P], p[i - 1], p[i + 1]));
                    }
                }
                tail_sync<WARP>();
            }
        }
    }
}

// one cycle visit of level L (vcyclemultigrid P:575-627)
// Inlined at its single call site per level (the recursion is a template recursion), so `cur` and the arguments live in
// registers: as a __noinline__ function taking `cur` by reference every use of it was a local-memory round trip.
template <typename T, bool RBGS, int L, bool WARP>
__device__ __forceinline__ void tail_visit(T* base, const TailArgs<T>& a, unsigned& cur, int warp, int lane)
{
    typedef TailLv<T, L> LV;
    tail_smooth<T, RBGS, L, WARP>(base, a, cur, a.nu1, warp, lane);           // P:581
    if (L <= a.coarsest) {
        tail_smooth<T, RBGS, L, WARP>(base, a, cur, a.nu2, warp, lane);       // P:585
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
            MG_TAIL_FOR_POINTS(LV, WARP)
            {
                const int i = y * P + x;
                r[i] = resid_pt<T>(u[i], F[i], sigma4<T>(u[i - P], u[i + P], u[i - 1], u[i + 1]));
            }
            tail_sync<WARP>();
            cur &= ~(1u << (L - 1));
            T* fc = lc.F();
            T* uc = lc.A(cur);
            MG_TAIL_FOR_POINTS(LC, WARP)
            {
                const int i = (2 * y) * P + 2 * x;
                fc[y * Pc + x] = fw_pt<T>(a.w, r[i - P - 1], r[i - P + 1], r[i + P - 1], r[i + P + 1],
                                          r[i - 1], r[i + 1], r[i - P], r[i + P], r[i]);
                uc[y * Pc + x] = (T)0;
            }
            tail_sync<WARP>();
        }
        const int reps = (L - 1 <= a.coarsest) ? 1 : (a.gamma < 1 ? 1 : a.gamma);
        if constexpr (!WARP && L - 1 <= kTailWarpLevel) {
            // hand the sub-cycle of the small levels to warp 0; its buffer parities come back through shared memory
            unsigned* cur_slot = reinterpret_cast<unsigned*>(base + tail_off(kTailMaxLevel + 1));
            if (warp == 0) {
                for (int g = 0; g < reps; ++g) tail_visit<T, RBGS, L - 1, true>(base, a, cur, warp, lane);   // P:617
                if (lane == 0) *cur_slot = cur;
            }
            __syncthreads();
            cur = *cur_slot;
        } else {
            for (int g = 0; g < reps; ++g) tail_visit<T, RBGS, L - 1, WARP>(base, a, cur, warp, lane);       // P:617
        }
        {   // prolongation + correction (P:620-624)
            const T* e = lc.A(cur);
            T* u = lv.A(cur);
            MG_TAIL_FOR_POINTS(LV, WARP)
            {
                const T* c = e + (y >> 1) * Pc + (x >> 1);
                T v;
                if (!(y & 1) && !(x & 1)) v = c[0];                                               // P:401
                else if ((y & 1) && !(x & 1)) v = (T)0.5 * (c[0] + c[Pc]);                        // P:407
                else if (!(y & 1)) v = (T)0.5 * (c[0] + c[1]);                                    // P:413
                else v = (T)0.25 * (((c[0] + c[Pc]) + c[1]) + c[Pc + 1]);                         // P:419
                u[y * P + x] = u[y * P + x] + v;                                                  // P:623
            }
            tail_sync<WARP>();
        }
        tail_smooth<T, RBGS, L, WARP>(base, a, cur, a.nu2, warp, lane);       // P:625
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
    if constexpr (TOP <= kTailWarpLevel) {      // the whole problem is a small level: warp 0 runs it alone
        unsigned* cur_slot = reinterpret_cast<unsigned*>(base + tail_off(kTailMaxLevel + 1));
        if (warp == 0) {
            tail_visit<T, RBGS, TOP, true>(base, a, cur, warp, lane);
            if (lane == 0) *cur_slot = cur;
        }
        __syncthreads();
        cur = *cur_slot;
    } else {
        tail_visit<T, RBGS, TOP, false>(base, a, cur, warp, lane);
    }
    {
        const T* A = lv.A(cur);
        MG_TAIL_FOR_POINTS(LV, false) { a.u[(i64)y * a.pitch + x] = A[y * P + x]; }
    }
}

template <typename T, bool RBGS>
__global__ void __launch_bounds__(kTailThreads, 1)   // one CTA per SM: 64 registers per thread (the default heuristic caps at 32 and spills)
k_tail(const TailArgs<T> a)
{
    pdl_wait();
    pdl_trigger();
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

// zero-guess variant (zero-guess chain): u of the top level is known to be zero and is not read.
// Its own entry point: the kernel neither loads nor waits for u of the top level.
template <typename T, bool RBGS>
__global__ void __launch_bounds__(kTailThreads, 1)   // one CTA per SM: 64 registers per thread (the default heuristic caps at 32 and spills)
k_tail_zg(const TailArgs<T> a)
{
    pdl_wait();
    pdl_trigger();
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
