// tile.cuh — CUDA wrapper of the shared-memory tile kernels (tile_core.h) for the mid levels.
// EXPERIMENTAL in round 1: compiled, checked on the CPU by the host emulation test
// (tests/test_tile_emulation.py) but not yet run on a GPU; selected only with MGB200_TILE=1.
#pragma once

#include "common.cuh"
#include "tile_core.h"

namespace mgb {

constexpr int kTileThreads = 256;

template <typename T, int NS, int MODE, bool RBGS, int TY, int TX>
__global__ void __launch_bounds__(kTileThreads)
k_tile(const TileArgs<T> a)
{
    typedef TileCfg<T, NS, MODE, TY, TX> C;
    extern __shared__ __align__(16) unsigned char tile_smem[];
    T* smem = reinterpret_cast<T*>(tile_smem);
#pragma unroll
    for (int ph = 0; ph < C::NPHASES; ++ph) {
        tile_phase<T, NS, MODE, RBGS, TY, TX>(a, smem, (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x, kTileThreads, ph);
        __syncthreads();
    }
}

}  // namespace mgb
