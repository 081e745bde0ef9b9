// tests/host_emul/cuda_emu/cooperative_groups.h — TEST INFRASTRUCTURE: the slice of cooperative_groups that
// csrc/ctail.cuh uses (thread-block clusters), on top of the fiber scheduler of emu_runtime.cpp.
#pragma once
#include "cuda_runtime.h"

namespace cooperative_groups {

struct cluster_group {
    unsigned block_rank() const { return emu::g_cur->cta_rank; }
    unsigned num_blocks() const { return emu::g_cur->cluster_size; }
    void sync() const { emu::cluster_barrier(); }
    // address of the same shared-memory object in the CTA with the given cluster rank (DSMEM)
    template <typename T>
    T* map_shared_rank(T* p, unsigned rank) const
    {
        const ptrdiff_t off = (unsigned char*)p - emu::g_cur->smem;
        return (T*)(emu::cluster_smem(rank) + off);
    }
};
inline cluster_group this_cluster() { return cluster_group(); }

}  // namespace cooperative_groups
