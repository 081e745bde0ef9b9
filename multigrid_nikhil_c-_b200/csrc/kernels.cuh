// kernels.cuh — the unfused level operators (one launch == one reference library call
// chain).  Each thread owns one aligned 16-byte column vector (2 doubles / 4 floats) and
// marches down RY rows keeping the 3-row stencil window in registers, so every element
// of u is loaded from L2/HBM once per column strip (plus two scalar neighbours that hit
// L1).  Algorithmic traffic: 3 S bytes per point for smoother and residual.
#pragma once

#include <algorithm>

#include "common.cuh"

namespace mgb {

constexpr int kTX = 128;   // threads per block along x  (=> 256 fp64 / 512 fp32 columns per block)
constexpr int kRYMax = 32; // max rows marched per block (fewer on small levels, see pick_ry)

struct LaunchCounter { long long n = 0; };

// ---------------------------------------------------------------------------------
// Weighted Jacobi sweep, out of place (jacobirelaxation P:138-142).
// rows [ya, yb) must be interior rows (1 <= ya, yb <= N).
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kTX)
k_jacobi(const T* __restrict__ u, T* __restrict__ out, const T* __restrict__ f,
         i64 pitch, int N, int ya, int yb, int ry, T c0, T c1)
{
    constexpr int V = Vec<T>::N;
    const int c = V * (blockIdx.x * kTX + threadIdx.x);
    if (c >= N) return;
    const int y0 = ya + blockIdx.y * ry;
    const int y1 = min(y0 + ry, yb);
    if (y0 >= y1) return;

    const T* pu = u + (i64)(y0 - 1) * pitch + c;
    T up[V], ce[V], dn[V], ff[V], o[V];
    ldv<T>(pu, up);
    pu += pitch;
    ldv<T>(pu, ce);
    T lf = (c > 0) ? pu[-1] : (T)0;
    T rt = pu[V];
    const T* pf = f + (i64)y0 * pitch + c;
    T* po = out + (i64)y0 * pitch + c;

#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
        pu += pitch;
        ldv<T>(pu, dn);
        ldv<T>(pf, ff);
        const T nlf = (c > 0) ? pu[-1] : (T)0;
        const T nrt = pu[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const T l = (k == 0) ? lf : ce[k - 1];
            const T r = (k == V - 1) ? rt : ce[k + 1];
            const T val = jacobi_pt<T>(c0, c1, ce[k], ff[k], sigma4<T>(up[k], dn[k], l, r));
            const int x = c + k;
            o[k] = (x >= 1 && x < N) ? val : (T)0;
        }
        stv<T>(po, o);
#pragma unroll
        for (int k = 0; k < V; ++k) { up[k] = ce[k]; ce[k] = dn[k]; }
        lf = nlf; rt = nrt;
        pf += pitch; po += pitch;
    }
}

// ---------------------------------------------------------------------------------
// One colour of red-black Gauss-Seidel, in place.  colour 0 = red = (y + x) even
// (global node indices == 1-based interior indices), SURVEY App. B.2.
// The whole 16-byte vector is stored back (the other colour's values unchanged), so
// stores stay full-sector.
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kTX)
k_rbgs(T* __restrict__ u, const T* __restrict__ f, i64 pitch, int N, int ya, int yb, int ry, int colour)
{
    constexpr int V = Vec<T>::N;
    const int c = V * (blockIdx.x * kTX + threadIdx.x);
    if (c >= N) return;
    const int y0 = ya + blockIdx.y * ry;
    const int y1 = min(y0 + ry, yb);
    if (y0 >= y1) return;

    // NOTE: rows are NOT marched with a register window here: row y-1's points of the
    // *other* colour are final for this half sweep, but within one colour pass every
    // read is of the other colour, so plain loads are race free.
    for (int y = y0; y < y1; ++y) {
        T* pc = u + (i64)y * pitch + c;
        T up[V], ce[V], dn[V], ff[V], o[V];
        ldv<T>(pc - pitch, up);
        ldv<T>(pc, ce);
        ldv<T>(pc + pitch, dn);
        ldv<T>(f + (i64)y * pitch + c, ff);
        const T lf = (c > 0) ? pc[-1] : (T)0;
        const T rt = pc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int x = c + k;
            const T l = (k == 0) ? lf : ce[k - 1];
            const T r = (k == V - 1) ? rt : ce[k + 1];
            const T val = gs_pt<T>(ff[k], sigma4<T>(up[k], dn[k], l, r));
            const bool mine = (((y + x) & 1) == colour) && x >= 1 && x < N;
            o[k] = mine ? val : ce[k];
        }
        stv<T>(pc, o);
    }
}

// ---------------------------------------------------------------------------------
// Residual r = f - (4u - Sigma) (P:604-607), optional sum of squares.
// partials (may be null): one double per block, index blockIdx.y*gridDim.x+blockIdx.x.
// ---------------------------------------------------------------------------------
template <typename T, bool STORE>
__global__ void __launch_bounds__(kTX)
k_residual(const T* __restrict__ u, const T* __restrict__ f, T* __restrict__ r,
           i64 pitch, int N, int ya, int yb, int ry, double* __restrict__ partials)
{
    constexpr int V = Vec<T>::N;
    const int c = V * (blockIdx.x * kTX + threadIdx.x);
    const int y0 = ya + blockIdx.y * ry;
    const int y1 = min(y0 + ry, yb);
    double acc = 0.0;
    if (c < N && y0 < y1) {
        const T* pu = u + (i64)(y0 - 1) * pitch + c;
        T up[V], ce[V], dn[V], ff[V], o[V];
        ldv<T>(pu, up);
        pu += pitch;
        ldv<T>(pu, ce);
        T lf = (c > 0) ? pu[-1] : (T)0;
        T rt = pu[V];
        const T* pf = f + (i64)y0 * pitch + c;
        T* po = r + (i64)y0 * pitch + c;
#pragma unroll 2
        for (int y = y0; y < y1; ++y) {
            pu += pitch;
            ldv<T>(pu, dn);
            ldv<T>(pf, ff);
            const T nlf = (c > 0) ? pu[-1] : (T)0;
            const T nrt = pu[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T l = (k == 0) ? lf : ce[k - 1];
                const T rr = (k == V - 1) ? rt : ce[k + 1];
                const T val = resid_pt<T>(ce[k], ff[k], sigma4<T>(up[k], dn[k], l, rr));
                const int x = c + k;
                o[k] = (x >= 1 && x < N) ? val : (T)0;
                acc += (double)o[k] * (double)o[k];
            }
            if (STORE) stv<T>(po, o);
#pragma unroll
            for (int k = 0; k < V; ++k) { up[k] = ce[k]; ce[k] = dn[k]; }
            lf = nlf; rt = nrt;
            pf += pitch; po += pitch;
        }
    }
    if (partials) {
        // warp-shuffle tree, then one value per warp through shared memory
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
        __shared__ double wsum[kTX / 32];
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kTX / 32; ++w) t += wsum[w];
            partials[(i64)blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}

// fixed-order final reduction of the per-block partials (deterministic for a given grid)
static __global__ void __launch_bounds__(256)
k_sum_partials(const double* __restrict__ partials, int count, double* __restrict__ out)
{
    pdl_wait();
    pdl_trigger();
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += 256) acc += partials[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        *out = t;
    }
}

// ---------------------------------------------------------------------------------
// Full-weighting restriction (restriction2d P:531-546).  Coarse rows [Ia, Ib), all
// coarse columns.  Also writes the zero coarse initial guess (P:613) when uc != null.
// ---------------------------------------------------------------------------------
constexpr int kRTX = 64, kRTY = 4;
template <typename T>
__global__ void __launch_bounds__(kRTX* kRTY)
k_restrict(const T* __restrict__ rf, i64 pf, T* __restrict__ fc, T* __restrict__ uc, i64 pc,
           int Nc, int Ia, int Ib, T w)
{
    constexpr int V = Vec<T>::N;
    const int Jc = V * (blockIdx.x * kRTX + threadIdx.x);
    const int I = Ia + blockIdx.y * kRTY + threadIdx.y;
    if (Jc >= Nc || I >= Ib) return;
    const T* pN = rf + (i64)(2 * I - 1) * pf + 2 * Jc;
    const T* pC = pN + pf;
    const T* pS = pC + pf;
    // index k <-> fine column 2*Jc - 1 + k, k = 0..2V
    T n[2 * V + 1], cc[2 * V + 1], s[2 * V + 1];
    n[0] = (Jc > 0) ? pN[-1] : (T)0;
    cc[0] = (Jc > 0) ? pC[-1] : (T)0;
    s[0] = (Jc > 0) ? pS[-1] : (T)0;
    {
        T a[V], b[V];
        ldv<T>(pN, a); ldv<T>(pN + V, b);
#pragma unroll
        for (int k = 0; k < V; ++k) { n[1 + k] = a[k]; n[1 + V + k] = b[k]; }
        ldv<T>(pC, a); ldv<T>(pC + V, b);
#pragma unroll
        for (int k = 0; k < V; ++k) { cc[1 + k] = a[k]; cc[1 + V + k] = b[k]; }
        ldv<T>(pS, a); ldv<T>(pS + V, b);
#pragma unroll
        for (int k = 0; k < V; ++k) { s[1 + k] = a[k]; s[1 + V + k] = b[k]; }
    }
    T o[V], z[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const T val = fw_pt<T>(w, n[2 * j], n[2 * j + 2], s[2 * j], s[2 * j + 2],
                               cc[2 * j], cc[2 * j + 2], n[2 * j + 1], s[2 * j + 1], cc[2 * j + 1]);
        const int J = Jc + j;
        o[j] = (J >= 1 && J < Nc) ? val : (T)0;
        z[j] = (T)0;
    }
    stv<T>(fc + (i64)I * pc + Jc, o);
    if (uc) stv<T>(uc + (i64)I * pc + Jc, z);
}

// ---------------------------------------------------------------------------------
// Bilinear prolongation (interpolation2d P:337-425) fused with the correction
// (vm::add P:620-624): ADD ? u += P e : u = P e.   Fine rows [ya, yb).
// ---------------------------------------------------------------------------------
constexpr int kPTX = 64, kPTY = 4;
template <typename T, bool ADD>
__global__ void __launch_bounds__(kPTX* kPTY)
k_prolong(const T* __restrict__ ec, i64 pc, T* __restrict__ uf, i64 pf, int N, int ya, int yb)
{
    constexpr int V = Vec<T>::N;
    constexpr int H = V / 2;
    const int c = V * (blockIdx.x * kPTX + threadIdx.x);
    const int y = ya + blockIdx.y * kPTY + threadIdx.y;
    if (c >= N || y >= yb) return;
    const int J0 = c >> 1;
    T e[V];
    if ((y & 1) == 0) {
        const T* pa = ec + (i64)(y >> 1) * pc + J0;
        T a[H + 1];
#pragma unroll
        for (int k = 0; k <= H; ++k) a[k] = pa[k];
#pragma unroll
        for (int k = 0; k < H; ++k) {
            e[2 * k] = a[k];                                  // P:401
            e[2 * k + 1] = (T)0.5 * (a[k] + a[k + 1]);        // P:413
        }
    } else {
        const T* pa = ec + (i64)((y - 1) >> 1) * pc + J0;
        const T* pb = pa + pc;
        T a[H + 1], b[H + 1];
#pragma unroll
        for (int k = 0; k <= H; ++k) { a[k] = pa[k]; b[k] = pb[k]; }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            e[2 * k] = (T)0.5 * (a[k] + b[k]);                                    // P:407
            e[2 * k + 1] = (T)0.25 * (((a[k] + b[k]) + a[k + 1]) + b[k + 1]);     // P:419
        }
    }
    T* pu = uf + (i64)y * pf + c;
    T o[V];
    if (ADD) {
        T cur[V];
        ldv<T>(pu, cur);
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = cur[k] + e[k];                         // P:623
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = e[k];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) { const int x = c + k; if (!(x >= 1 && x < N)) o[k] = (T)0; }
    stv<T>(pu, o);
}

// ---------------------------------------------------------------------------------
// fill interior rows [ya, yb) with a constant (globalforcefunction P:283-335: f*h^2)
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_fill(T* __restrict__ p, i64 pitch, int N, int ya, int yb, T value)
{
    constexpr int V = Vec<T>::N;
    const int c = V * (blockIdx.x * 256 + threadIdx.x);
    const int y = ya + blockIdx.y;
    if (c >= N || y >= yb) return;
    T o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { const int x = c + k; o[k] = (x >= 1 && x < N) ? value : (T)0; }
    stv<T>(p + (i64)y * pitch + c, o);
}

// ---------------------------------------------------------------------------------
// Synthetic right-hand side for benchmarks (no reference counterpart: the reference only has the constant f = 4,
// P:123): b(y, x) = h^2 * (2 U - 1), U = top 53 bits of splitmix64(seed, idx) * 2^-53, idx = the reference's interior
// index (y-1)*n + (x-1) (P:227-228).  A pure function of the GLOBAL index, so every rank of a row-slab run -- and the
// single-GPU run of the same grid -- holds the same values without any host-side generation or transfer.
// ---------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long seed, unsigned long long idx)
{
    unsigned long long z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_fill_synthetic(T* __restrict__ p, i64 pitch, int N, int ya, int yb, double h2, unsigned long long seed)
{
    constexpr int V = Vec<T>::N;
    const int c = V * (blockIdx.x * 256 + threadIdx.x);
    const int y = ya + blockIdx.y;
    if (c >= N || y >= yb) return;
    const unsigned long long n = (unsigned long long)(N - 1);
    T o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int x = c + k;
        if (x >= 1 && x < N) {
            const unsigned long long z = splitmix64(seed, (unsigned long long)(y - 1) * n + (unsigned long long)(x - 1));
            const double u01 = (double)(z >> 11) * 0x1.0p-53;      // exact
            o[k] = (T)(h2 * (2.0 * u01 - 1.0));                    // exact in fp64 (h2 is a power of two); one rounding in fp32
        } else {
            o[k] = (T)0;
        }
    }
    stv<T>(p + (i64)y * pitch + c, o);
}

// ---------------------------------------------------------------------------------
// Order-independent 64-bit checksum of the interior values of rows [ya, yb): the sum (mod 2^64) over the points of
// splitmix64(value bits, global interior index).  Integer addition is associative, so the checksum of a grid is the sum
// of the checksums of its row slabs: N ranks and one rank agree exactly when every owned value agrees bit for bit.
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_checksum(const T* __restrict__ p, i64 pitch, int N, int ya, int yb, unsigned long long* __restrict__ out)
{
    const int x = blockIdx.x * 256 + threadIdx.x + 1;
    unsigned long long acc = 0;
    if (x < N) {
        const unsigned long long n = (unsigned long long)(N - 1);
        for (int y = ya + blockIdx.y; y < yb; y += gridDim.y) {
            const T v = p[(i64)y * pitch + x];
            unsigned long long bits;
            if constexpr (sizeof(T) == 8) bits = (unsigned long long)__double_as_longlong(v);
            else bits = (unsigned long long)__float_as_uint(v);
            acc += splitmix64(bits, (unsigned long long)(y - 1) * n + (unsigned long long)(x - 1));
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out, acc);
}

// pack / unpack between the reference's interior-only host layout (staged flat in device
// memory, n x n row-major) and the padded device layout.  Rows [ya, yb) (node rows).
template <typename T, bool TO_PADDED>
__global__ void __launch_bounds__(256)
k_repack(T* __restrict__ padded, i64 pitch, T* __restrict__ flat, int N, int ya, int yb)
{
    const int x = blockIdx.x * 256 + threadIdx.x + 1;   // node column 1..N-1
    const int y = ya + blockIdx.y;
    if (x >= N || y >= yb) return;
    const i64 n = N - 1;
    if (TO_PADDED) padded[(i64)y * pitch + x] = flat[(i64)(y - 1) * n + (x - 1)];
    else flat[(i64)(y - 1) * n + (x - 1)] = padded[(i64)y * pitch + x];
}

// ---------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------
inline unsigned cdiv(i64 a, i64 b) { return (unsigned)((a + b - 1) / b); }

// rows marched per block: 32 on big levels (halo rows re-read: 6%), fewer on small levels so
// that the grid still has >= ~4 blocks per SM (a 32-row march is a 32-deep latency chain)
inline int pick_ry(int rows, unsigned blocks_x)
{
    const i64 ry = (i64)rows * blocks_x / 592;
    return (int)(ry < 2 ? 2 : (ry > kRYMax ? kRYMax : ry));
}

template <typename T>
inline void launch_jacobi(cudaStream_t st, LaunchCounter& lc, const T* u, T* out, const T* f,
                          i64 pitch, int N, int ya, int yb, T c0, T c1)
{
    if (ya >= yb) return;
    const unsigned gx = cdiv(N, Vec<T>::N * kTX);
    const int ry = pick_ry(yb - ya, gx);
    dim3 grid(gx, cdiv(yb - ya, ry));
    k_jacobi<T><<<grid, kTX, 0, st>>>(u, out, f, pitch, N, ya, yb, ry, c0, c1);
    ++lc.n;
}

template <typename T>
inline void launch_rbgs(cudaStream_t st, LaunchCounter& lc, T* u, const T* f, i64 pitch, int N,
                        int ya, int yb, int colour)
{
    if (ya >= yb) return;
    const unsigned gx = cdiv(N, Vec<T>::N * kTX);
    const int ry = pick_ry(yb - ya, gx);
    dim3 grid(gx, cdiv(yb - ya, ry));
    k_rbgs<T><<<grid, kTX, 0, st>>>(u, f, pitch, N, ya, yb, ry, colour);
    ++lc.n;
}

// returns the number of partials written (0 when partials == null)
template <typename T>
inline int launch_residual(cudaStream_t st, LaunchCounter& lc, const T* u, const T* f, T* r,
                           i64 pitch, int N, int ya, int yb, double* partials, bool store)
{
    if (ya >= yb) return 0;
    const unsigned gx = cdiv(N, Vec<T>::N * kTX);
    const int ry = pick_ry(yb - ya, gx);
    dim3 grid(gx, cdiv(yb - ya, ry));
    if (store) k_residual<T, true><<<grid, kTX, 0, st>>>(u, f, r, pitch, N, ya, yb, ry, partials);
    else k_residual<T, false><<<grid, kTX, 0, st>>>(u, f, r, pitch, N, ya, yb, ry, partials);
    ++lc.n;
    return partials ? (int)(grid.x * grid.y) : 0;
}

inline void launch_sum_partials(cudaStream_t st, LaunchCounter& lc, const double* partials, int count, double* out)
{
    k_sum_partials<<<1, 256, 0, st>>>(partials, count, out);
    ++lc.n;
}

template <typename T>
inline void launch_restrict(cudaStream_t st, LaunchCounter& lc, const T* rf, i64 pf, T* fc, T* uc,
                            i64 pc, int Nc, int Ia, int Ib, T w)
{
    if (Ia >= Ib) return;
    dim3 grid(cdiv(Nc, Vec<T>::N * kRTX), cdiv(Ib - Ia, kRTY));
    k_restrict<T><<<grid, dim3(kRTX, kRTY), 0, st>>>(rf, pf, fc, uc, pc, Nc, Ia, Ib, w);
    ++lc.n;
}

template <typename T>
inline void launch_prolong(cudaStream_t st, LaunchCounter& lc, const T* ec, i64 pc, T* uf, i64 pf,
                           int N, int ya, int yb, bool add)
{
    if (ya >= yb) return;
    dim3 grid(cdiv(N, Vec<T>::N * kPTX), cdiv(yb - ya, kPTY));
    if (add) k_prolong<T, true><<<grid, dim3(kPTX, kPTY), 0, st>>>(ec, pc, uf, pf, N, ya, yb);
    else k_prolong<T, false><<<grid, dim3(kPTX, kPTY), 0, st>>>(ec, pc, uf, pf, N, ya, yb);
    ++lc.n;
}

template <typename T>
inline void launch_fill(cudaStream_t st, LaunchCounter& lc, T* p, i64 pitch, int N, int ya, int yb, T value)
{
    if (ya >= yb) return;
    dim3 grid(cdiv(N, Vec<T>::N * 256), (unsigned)(yb - ya));
    k_fill<T><<<grid, 256, 0, st>>>(p, pitch, N, ya, yb, value);
    ++lc.n;
}

template <typename T>
inline void launch_fill_synthetic(cudaStream_t st, LaunchCounter& lc, T* p, i64 pitch, int N, int ya, int yb, double h2,
                                  unsigned long long seed)
{
    if (ya >= yb) return;
    dim3 grid(cdiv(N, Vec<T>::N * 256), (unsigned)(yb - ya));
    k_fill_synthetic<T><<<grid, 256, 0, st>>>(p, pitch, N, ya, yb, h2, seed);
    ++lc.n;
}

template <typename T>
inline void launch_checksum(cudaStream_t st, LaunchCounter& lc, const T* p, i64 pitch, int N, int ya, int yb,
                            unsigned long long* out)
{
    if (ya >= yb || N < 2) return;
    dim3 grid(cdiv(N - 1, 256), (unsigned)std::min(yb - ya, 64));
    k_checksum<T><<<grid, 256, 0, st>>>(p, pitch, N, ya, yb, out);
    ++lc.n;
}

template <typename T>
inline void launch_repack(cudaStream_t st, LaunchCounter& lc, T* padded, i64 pitch, T* flat, int N,
                          int ya, int yb, bool to_padded)
{
    if (ya >= yb || N < 2) return;
    dim3 grid(cdiv(N - 1, 256), (unsigned)(yb - ya));
    if (to_padded) k_repack<T, true><<<grid, 256, 0, st>>>(padded, pitch, flat, N, ya, yb);
    else k_repack<T, false><<<grid, 256, 0, st>>>(padded, pitch, flat, N, ya, yb);
    ++lc.n;
}

}  // namespace mgb
