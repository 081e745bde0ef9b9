// coarse.cuh — exact solve on the coarsest level (MG_COARSE_EXACT).
//
// The reference's second version solves the coarsest level directly instead of smoothing it:
//     if (current_level == coarsest_level) { vec_h = direct_solver(obj.coarsest_level_matrix, f_h); return vec_h; }
// (Multigrid_functions.cpp M:136-139; direct_solver = Eigen SparseLU, M:63-72).  With the reference's own depth
// (coarsest = finest - 3, P:17-18) that is what makes the cycle converge at the textbook rate (0.22 per V(2,2) instead
// of 0.92-0.98 with nu1+nu2 sweeps on a 127^2 grid).
//
// Here the solve is a direct one too, but shaped for the GPU: A = [-1; -1 4 -1; -1] with the zero Dirichlet ring is
// diagonalised by the 2-D sine transform, so
//     u = c * S ((S F S) ./ Lambda) S,   S[j][k] = sin(pi (j+1)(k+1)/(n+1)),  Lambda[m][k] = d[m] + d[k],
//     d[j] = 2 - 2 cos(pi (j+1)/(n+1)),  c = (2/(n+1))^2,
// i.e. four small dense products (n <= 511), one thread per output element, the dot product in ascending index order
// with separate multiply and add -- exactly the order of the oracle's mgo_coarse_exact, so cycles stay bit-identical.
// The tables are formed on the host in double (same expressions as the oracle) and rounded to T once.
#pragma once

#include "common.cuh"

namespace mgb {

constexpr int kCoarseExactMaxLevel = 9;   // S is n x n: 2 MB in fp64 at 511^2

// C = A B (n x n, row-major with the given pitches).  EPI 0: store; 1: divide by d[row] + d[col]; 2: multiply by c.
template <typename T, int EPI>
__global__ void __launch_bounds__(256)
k_dense_product(const T* __restrict__ A, i64 pa, const T* __restrict__ B, i64 pb, T* __restrict__ C, i64 pc, int n,
                const T* __restrict__ d, T c)
{
    const int col = blockIdx.x * 64 + (threadIdx.x & 63);
    const int row = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (col >= n || row >= n) return;
    const T* a = A + (i64)row * pa;
    const T* b = B + col;
    T acc = (T)0;
    for (int t = 0; t < n; ++t) acc = acc + a[t] * b[(i64)t * pb];
    if (EPI == 1) acc = acc / (d[row] + d[col]);
    if (EPI == 2) acc = c * acc;
    C[(i64)row * pc + col] = acc;
}

template <typename T, int EPI>
inline void launch_dense_product(cudaStream_t st, LaunchCounter& lc, const T* A, i64 pa, const T* B, i64 pb, T* C, i64 pc,
                                 int n, const T* d, T c)
{
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((n + 3) / 4));
    k_dense_product<T, EPI><<<grid, 256, 0, st>>>(A, pa, B, pb, C, pc, n, d, c);
    ++lc.n;
}

}  // namespace mgb
