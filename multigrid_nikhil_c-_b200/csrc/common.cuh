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

#include "formulas.h"

namespace mgb {


// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor grid has completed
// and its memory is visible, pdl_trigger() lets the NEXT kernel of the stream start its own launch early.  Every kernel
// of a cycle calls both first thing, so the ~2-3 us launch latency of each kernel hides behind the one before it while
// the data dependencies stay exactly those of ordinary stream order.  No-ops for kernels launched without the attribute.
#if defined(MGB_EMU) || !defined(__CUDA_ARCH__)
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

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

}  // namespace mgb
