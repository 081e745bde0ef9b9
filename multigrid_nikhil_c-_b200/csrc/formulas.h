// formulas.h — the point formulas shared by every kernel, by the host-side emulation tests
// (tests/host_emul) and documented in DESIGN.md section 2.  Evaluation order is the parity
// contract with the oracle (oracle/mg_oracle_impl.inc); device code is
// compiled with --fmad=false and host code with -ffp-contract=off so nothing contracts into
// an FMA.
#pragma once

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif

namespace mgb {

typedef long long i64;

//   Sigma = (up + down) + (left + right)
template <typename T> MG_HD T sigma4(T up, T dn, T lf, T rt) { return (up + dn) + (lf + rt); }
// weighted Jacobi, P:138-142: ((1-w) v + (w/4) f) + (w/4) Sigma
template <typename T> MG_HD T jacobi_pt(T c0, T c1, T v, T f, T sig) { return (c0 * v + c1 * f) + c1 * sig; }
// Gauss-Seidel point update: 0.25 (f + Sigma)
template <typename T> MG_HD T gs_pt(T f, T sig) { return (T)0.25 * (f + sig); }
// residual, P:604-607: f - (4 v - Sigma)
template <typename T> MG_HD T resid_pt(T v, T f, T sig) { return f - ((T)4 * v - sig); }
// full weighting, P:539-542: w (((NW+NE+SW+SE) + 2 (W+E+N+S)) + 4 C), left to right
template <typename T>
MG_HD T fw_pt(T w, T nw, T ne, T sw, T se, T wv, T ev, T nv, T sv, T cv)
{
    T corners = ((nw + ne) + sw) + se;
    T edges = ((wv + ev) + nv) + sv;
    return w * ((corners + (T)2 * edges) + (T)4 * cv);
}
// bilinear prolongation of the coarse values around fine node (y, x), P:398-420:
//   c00 = c[y/2][x/2], c10 = next coarse row, c01 = next coarse column, c11 = both
template <typename T>
MG_HD T prolong_pt(int y, int x, T c00, T c10, T c01, T c11)
{
    if (!(y & 1) && !(x & 1)) return c00;                                   // P:401
    if ((y & 1) && !(x & 1)) return (T)0.5 * (c00 + c10);                   // P:407
    if (!(y & 1)) return (T)0.5 * (c00 + c01);                              // P:413
    return (T)0.25 * (((c00 + c10) + c01) + c11);                           // P:419
}

}  // namespace mgb
