// comm.cu — row-slab communication.  (world > 1 support lands after the single-GPU path.)
#include "comm.cuh"

#include <cstring>

namespace mgb {

struct Comm { int dummy; };

Comm* comm_create(Ctx&) { throw MgError(MG_ERR_COMM, "world > 1 is not available in this build"); }
void comm_destroy(Comm* c) { delete c; }
void comm_halo_exchange(Ctx&, Level&, char*, int) {}
void comm_zero_halo(Ctx&, Level&, char*) {}
void comm_allgather_rows(Ctx&, Level&, char*) {}
double comm_sum(Ctx&, const double*) { return 0.0; }
int comm_unique_id(void* out128) { std::memset(out128, 0, MG_COMM_ID_BYTES); return MG_ERR_COMM; }

}  // namespace mgb
