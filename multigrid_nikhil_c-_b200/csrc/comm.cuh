// comm.cuh — row-slab halo exchange between the GPUs of one box (no reference
// counterpart: the reference is single-device).  Implemented in comm.cu.
#pragma once

#include "ctx.cuh"

namespace mgb {

void slab_rows(int level, int rank, int world, int* lo, int* hi);

Comm* comm_create(Ctx& ctx);
void comm_destroy(Comm* c);
// exchange `depth` owned edge rows of the array at virtual base `base` with both neighbours
// (on `stream`, default: the context's stream)
void comm_halo_exchange(Ctx& ctx, Level& lv, char* base, int depth, cudaStream_t stream = nullptr);
// clear the halo rows of `base`
void comm_zero_halo(Ctx& ctx, Level& lv, char* base);
// replicated level: every rank contributed its slab of rows of `base`; gather all slabs everywhere
void comm_allgather_rows(Ctx& ctx, Level& lv, char* base);
// sum of one device double over ranks in fixed rank order; synchronises; same bits on every rank
double comm_sum(Ctx& ctx, const double* d_value);
int comm_unique_id(void* out128);

}  // namespace mgb
