// common.cuh — shared definitions for the sm_100a kernels and the context.
//
// Device layout of one level (DESIGN.md "Data layout in HBM"): the FULL node grid
// including the zero Dirichlet ring, rows 0..N (N = 2^level), `pitch` elements per row
// (a multiple of 32 elements, so every row starts on a 128-byte line and the 16-byte
// vector of columns [c, c+V) with c % V == 0 is aligned).  Column 0, column N, the
// padding columns and rows 0 / N hold zeros at all times; kernels only ever store zeros
// there.  Pointers handed to kernels are *virtual* row-0 bases: element (y, x) lives at
// base[y * pitch + x]; for a row slab only rows [stored_lo, stored_hi) are backed.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mgb {

typedef long long i64;

template <typename T> struct Vec;
template <> struct Vec<double> { static constexpr int N = 2; typedef double2 type; };
template <> struct Vec<float>  { static constexpr int N = 4; typedef float4 type; };

template <typename T> __device__ __forceinline__ void ldv(const T* p, T (&v)[Vec<T>::N]);
template <> __device__ __forceinline__ void ldv<double>(const double* p, double (&v)[2])
{
    double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ldv<float>(const float* p, float (&v)[4])
{
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <typename T> __device__ __forceinline__ void stv(T* p, const T (&v)[Vec<T>::N]);
template <> __device__ __forceinline__ void stv<double>(double* p, const double (&v)[2])
{
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}
template <> __device__ __forceinline__ void stv<float>(float* p, const float (&v)[4])
{
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// The three point formulas.  Evaluation order is the parity contract with the oracle
// (oracle/mg_oracle_impl.inc); the library is compiled with
// --fmad=false so none of these contracts into an FMA.
//   Sigma = (up + down) + (left + right)
template <typename T> __device__ __forceinline__ T sigma4(T up, T dn, T lf, T rt) { return (up + dn) + (lf + rt); }
// weighted Jacobi, P:138-142: ((1-w) v + (w/4) f) + (w/4) Sigma
template <typename T> __device__ __forceinline__ T jacobi_pt(T c0, T c1, T v, T f, T sig) { return (c0 * v + c1 * f) + c1 * sig; }
// Gauss-Seidel point update: 0.25 (f + Sigma)
template <typename T> __device__ __forceinline__ T gs_pt(T f, T sig) { return (T)0.25 * (f + sig); }
// residual, P:604-607: f - (4 v - Sigma)
template <typename T> __device__ __forceinline__ T resid_pt(T v, T f, T sig) { return f - ((T)4 * v - sig); }
// full weighting, P:539-542: w (((NW+NE+SW+SE) + 2 (W+E+N+S)) + 4 C), left to right
template <typename T>
__device__ __forceinline__ T fw_pt(T w, T nw, T ne, T sw, T se, T wv, T ev, T nv, T sv, T cv)
{
    T corners = ((nw + ne) + sw) + se;
    T edges = ((wv + ev) + nv) + sv;
    return w * ((corners + (T)2 * edges) + (T)4 * cv);
}

}  // namespace mgb
