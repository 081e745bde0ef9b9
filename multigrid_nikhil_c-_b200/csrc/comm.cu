// comm.cu — row-slab communication between the GPUs of one box (one process per GPU).
//
// No reference counterpart (the reference is single-device).  NCCL is loaded
// with dlopen("libnccl.so.2") so that libmgb200.so has no link-time dependency on it and,
// inside a PyTorch process, shares the NCCL build torch already loaded.  The launcher
// (bench.py / tests, via torch.distributed; or any other out-of-band channel) broadcasts
// rank 0's mg_comm_id() bytes; mg_create(world > 1) then joins the communicator.
//
// Exchanges are stream-ordered on the context's stream:
//   halo exchange   grouped ncclSend/ncclRecv of `depth` contiguous rows with each neighbour
//                   (rows are full-pitch, so a block of rows is one contiguous message:
//                   131 KB per row at 16385^2 fp64) -- latency-, not bandwidth-bound
//   agglomeration   in-place ncclAllGather of the row slabs of the first replicated level
//   norm            ncclAllGather of one double per rank, summed on the host in rank order
//                   (bit-identical on every rank, independent of NCCL's reduction order)
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>

namespace mgb {

// minimal NCCL ABI (nccl.h is not required at build time)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclChar = 0, ncclFloat64 = 8 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl()
{
    static NcclApi api;
    if (api.handle) return api;
#ifdef MGB_EMU
    // CPU emulation build (tests/host_emul): file-based stand-in, one process per rank
    api.GetUniqueId = [](ncclUniqueId* id) -> ncclResult_t { return emu_nccl::GetUniqueId(id); };
    api.CommInitRank = [](ncclComm_t* c, int w, ncclUniqueId id, int r) -> ncclResult_t { return emu_nccl::CommInitRank((emu_nccl::Comm**)c, w, &id, r); };
    api.CommDestroy = [](ncclComm_t c) -> ncclResult_t { return emu_nccl::CommDestroy((emu_nccl::Comm*)c); };
    api.Send = [](const void* b, size_t n, int, int peer, ncclComm_t c, cudaStream_t s) -> ncclResult_t { return emu_nccl::Send(b, n, peer, (emu_nccl::Comm*)c, s); };
    api.Recv = [](void* b, size_t n, int, int peer, ncclComm_t c, cudaStream_t s) -> ncclResult_t { return emu_nccl::Recv(b, n, peer, (emu_nccl::Comm*)c, s); };
    api.AllGather = [](const void* sb, void* rb, size_t n, int dt, ncclComm_t c, cudaStream_t s) -> ncclResult_t {
        return emu_nccl::AllGather(sb, rb, n * (dt == ncclFloat64 ? 8 : 1), (emu_nccl::Comm*)c, s);
    };
    api.GroupStart = []() -> ncclResult_t { return 0; };
    api.GroupEnd = []() -> ncclResult_t { return 0; };
    api.GetErrorString = [](ncclResult_t r) -> const char* { return emu_nccl::GetErrorString(r); };
    api.handle = &api;
    return api;
#endif
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw MgError(MG_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) throw MgError(MG_ERR_COMM, std::string("libnccl lacks ") + name);
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.handle = h;
    return api;
}

#define NC(call)                                                                                    \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != 0)                                                                               \
            throw MgError(MG_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r__));     \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double* d_gather = nullptr;  // world doubles
    double* h_gather = nullptr;  // pinned
};

int comm_unique_id(void* out128)
{
    static_assert(sizeof(ncclUniqueId) == MG_COMM_ID_BYTES, "id size");
    ncclUniqueId id;
    NC(nccl().GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
    return MG_OK;
}

Comm* comm_create(Ctx& ctx)
{
    if (!ctx.cfg.comm_id) throw MgError(MG_ERR_ARG, "world > 1 needs cfg.comm_id (mg_comm_id() bytes of rank 0)");
    Comm* c = new Comm();
    c->rank = ctx.cfg.rank;
    c->world = ctx.cfg.world;
    ncclUniqueId id;
    std::memcpy(&id, ctx.cfg.comm_id, sizeof(id));
    try {
        NC(nccl().CommInitRank(&c->comm, c->world, id, c->rank));
        MG_CK(cudaMalloc(&c->d_gather, sizeof(double) * c->world));
        MG_CK(cudaMallocHost(&c->h_gather, sizeof(double) * c->world));
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

void comm_destroy(Comm* c)
{
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->d_gather) cudaFree(c->d_gather);
    if (c->h_gather) cudaFreeHost(c->h_gather);
    delete c;
}

void comm_halo_exchange(Ctx& ctx, Level& lv, char* base, int depth, cudaStream_t stream)
{
    Comm* c = ctx.comm;
    if (!c || !lv.distributed) return;
    if (!stream) stream = ctx.stream;
    const size_t row_bytes = (size_t)lv.pitch * ctx.esize;
    const size_t bytes = row_bytes * depth;
    auto row = [&](int y) { return base + (size_t)y * row_bytes; };
    NC(nccl().GroupStart());
    if (c->rank > 0) {  // neighbour above (smaller row indices)
        NC(nccl().Send(row(lv.own_lo), bytes, ncclChar, c->rank - 1, c->comm, stream));
        NC(nccl().Recv(row(lv.own_lo - depth), bytes, ncclChar, c->rank - 1, c->comm, stream));
    }
    if (c->rank < c->world - 1) {  // neighbour below
        NC(nccl().Send(row(lv.own_hi - depth), bytes, ncclChar, c->rank + 1, c->comm, stream));
        NC(nccl().Recv(row(lv.own_hi), bytes, ncclChar, c->rank + 1, c->comm, stream));
    }
    NC(nccl().GroupEnd());
}

void comm_zero_halo(Ctx& ctx, Level& lv, char* base)
{
    if (!lv.distributed) return;
    const size_t row_bytes = (size_t)lv.pitch * ctx.esize;
    if (lv.st_lo < lv.own_lo && ctx.cfg.rank > 0)
        MG_CK(cudaMemsetAsync(base + (size_t)lv.st_lo * row_bytes, 0, (size_t)(lv.own_lo - lv.st_lo) * row_bytes, ctx.stream));
    if (lv.st_hi > lv.own_hi && ctx.cfg.rank < ctx.cfg.world - 1)
        MG_CK(cudaMemsetAsync(base + (size_t)lv.own_hi * row_bytes, 0, (size_t)(lv.st_hi - lv.own_hi) * row_bytes, ctx.stream));
}

void comm_allgather_rows(Ctx& ctx, Level& lv, char* base)
{
    Comm* c = ctx.comm;
    if (!c) return;
    // rank r contributed node rows [r*N/R, (r+1)*N/R); equal counts, contiguous, in place
    const size_t row_bytes = (size_t)lv.pitch * ctx.esize;
    const size_t rows = (size_t)lv.N / c->world;
    const size_t bytes = rows * row_bytes;
    NC(nccl().AllGather(base + (size_t)c->rank * bytes, base, bytes, ncclChar, c->comm, ctx.stream));
}

double comm_sum(Ctx& ctx, const double* d_value)
{
    Comm* c = ctx.comm;
    if (!c) throw MgError(MG_ERR_STATE, "comm_sum without a communicator");
    NC(nccl().AllGather(d_value, c->d_gather, 1, ncclFloat64, c->comm, ctx.stream));
    MG_CK(cudaMemcpyAsync(c->h_gather, c->d_gather, sizeof(double) * c->world, cudaMemcpyDeviceToHost, ctx.stream));
    MG_CK(cudaStreamSynchronize(ctx.stream));
    double s = 0.0;
    for (int r = 0; r < c->world; ++r) s += c->h_gather[r];
    return s;
}

}  // namespace mgb
