// ctail.cuh — CUDA wrapper of the cluster coarse tail (ctail_core.h).  EXPERIMENTAL in round 1: compiled and
// checked on the CPU by exact emulation (tests/test_ctail_emulation.py), not yet run on a GPU; selected only
// with MGB200_CTAIL=1.
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"
#include "ctail_core.h"

namespace mgb {

constexpr int kCtailThreads = 512;

template <typename T>
struct CtailRemoteDev {
    T* base;   // this CTA's shared memory; every CTA of the cluster uses the same layout
    __device__ __forceinline__ T* operator()(int rank) const
    {
        return cooperative_groups::this_cluster().map_shared_rank(base, (unsigned)rank);
    }
};

template <typename T>
__global__ void __launch_bounds__(kCtailThreads)
k_ctail(const CtailArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char ctail_smem[];
    T* smem = reinterpret_cast<T*>(ctail_smem);
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const int me = (int)cluster.block_rank();
    const int total = ctail_off(a.top + 1, a.nctas);
    for (int i = threadIdx.x; i < total; i += kCtailThreads) smem[i] = (T)0;   // rings (and everything else) start at zero
    cluster.sync();
    CtailEnv<T, CtailRemoteDev<T>> env{me, a.nctas, smem, CtailRemoteDev<T>{smem}};
    for (int i = 0; i < a.nops; ++i) {
        const CtailOp op = a.ops[i];
        ctail_op<T, CtailRemoteDev<T>>(env, op, a, (int)threadIdx.x, kCtailThreads);
        cluster.sync();
    }
}

}  // namespace mgb
