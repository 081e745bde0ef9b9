// ctx.cuh — the multigrid context behind the C ABI (include/mgb200.h).
//
// It plays the role of the reference's `queue&` first argument plus its global level
// table `jacobi_matrices` (P:24-33): one Level per grid level, all device memory,
// the stream, and the CUDA-graph cache for whole cycles.  Matrix-free: there is no CSR
// operator; a level is just u (ping/pong), f and r in the padded layout of common.cuh.
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mgb200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace mgb {

struct MgError : std::runtime_error {
    int code;
    MgError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MG_CK(call)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            throw ::mgb::MgError(MG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + \
                                                  " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define MG_REQUIRE(cond, msg)                                          \
    do {                                                               \
        if (!(cond)) throw ::mgb::MgError(MG_ERR_ARG, std::string(msg)); \
    } while (0)

struct Comm;  // comm.cuh (row-slab halo exchange; null when world == 1)

constexpr int kHaloRows = 6;  // halo rows stored per side on distributed levels (deepest fused kernel: RB-GS PRE, NS=4)

struct Level {
    int level = 0;
    int N = 0;            // node index range 0..N, N = 2^level
    i64 pitch = 0;        // elements per stored row
    bool distributed = false;
    int own_lo = 0, own_hi = 0;  // owned interior node rows [own_lo, own_hi)
    int st_lo = 0, st_hi = 0;    // stored node rows [st_lo, st_hi) (owned + halo / ring)
    int halo = 0;                // halo rows stored per side on a distributed level (kHaloRows, or more for MGB200_COMM_AVOID)
    size_t bytes = 0;            // bytes of one array
    void* alloc[4] = {nullptr, nullptr, nullptr, nullptr};  // u0, u1, f, r (real allocations)
    char* u[2] = {nullptr, nullptr};                         // virtual row-0 bases
    char* f = nullptr;
    char* r = nullptr;
    int cur = 0;                 // which of u[] holds the current iterate
    // number of valid halo rows (distributed levels) of the current u, of f and of r;
    // an operator that rewrites only the owned rows sets it to 0, Ctx::ensure_halo exchanges lazily
    int hv_u = 0, hv_f = 0, hv_r = 0;
    // MGB200_ZERO_GUESS: the current u is LOGICALLY zero but its buffer was not written (the next kernel is a
    // zero-guess variant that does not read it); Ctx::materialize_u writes the zeros for every other reader
    bool u_zero = false;
    // MGB200_CHAIN, fullmultigrid: the current u is LOGICALLY the bare interpolation of the coarser level's iterate
    // (P:645) but was not computed: the first PRE of the level does it on the fly (k_stream_fmg_entry);
    // Ctx::materialize_u runs the real prolongation for every other reader
    bool u_interp = false;
};

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    long long kernels = 0;
    std::string state_after;   // Ctx::state_blob() once the captured cycle has run
};

struct Ctx {
    mg_config cfg;
    int device = 0;
    int esize = 8;
    cudaStream_t stream = nullptr;
    std::vector<Level> levels;  // indexed by level; valid for [coarsest, finest]
    std::string err;
    LaunchCounter lc;
    long long graph_launches = 0;
    size_t bytes_allocated = 0;
    double* d_partials = nullptr;
    int partials_cap = 0;
    double* d_norm = nullptr;
    double* h_norm = nullptr;  // pinned
    bool capturing = false;
    std::map<std::tuple<int, int, int, int, std::string>, GraphEntry> graphs;
    std::map<std::tuple<int, int, int, int>, int> stream_ry;  // tuned chunk height per (level, mode, NS, rbgs)
    std::map<std::tuple<int, int, int, int>, std::pair<void*, int>> ctail_ops;  // cluster-tail op lists on the device
    Comm* comm = nullptr;
    int aggl_level = 0;  // levels <= aggl_level are replicated on every rank
    bool graph_dist = false;  // capture NCCL exchanges into cycle graphs (MGB200_GRAPH_DIST=1)
    bool zero_guess = false;  // skip reading / writing the zero coarse guess (MGB200_ZERO_GUESS=1)
    bool comm_avoid = false;  // communication-avoiding slab schedule (MGB200_COMM_AVOID=1, csrc/sched.h)
    bool overlap = false;     // halo exchange on a second stream, overlapped with the interior rows (MGB200_OVERLAP=1)
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool chain = false;       // fuse POST of one visit of a level with PRE of the next (MGB200_CHAIN=1, stream.cuh MODE_POSTPRE)

    explicit Ctx(const mg_config& c);
    ~Ctx();

    Level& L(int level);
    const Level& L(int level) const;
    bool f64() const { return cfg.dtype == MG_F64; }

    // data movement
    enum Which { W_U = 0, W_F = 1, W_R = 2 };
    void set_host(int level, Which w, const void* host);
    void get_host(int level, Which w, void* host);
    void zero_u(int level);
    void force_constant(double f);

    // operators
    void smooth(int level, int nu);
    double residual(int level, bool want_norm, bool store);
    void restrict_to(int fine_level, bool from_rhs);
    void prolong(int fine_level, bool add);

    // cycles
    void cycle(int level, int nu1, int nu2, int gamma);
    void cycles(int level, int nu1, int nu2, int gamma, int count);   // `count` consecutive cycles (P:646-648 loop)
    void cycle_rec(int level, int nu1, int nu2, int gamma);
    void cycle_rec_visits(int level, int nu1, int nu2, int gamma, int visits);   // consecutive visits of one level
    void fmg(int cycles, int nu1, int nu2);
    int solve(double rtol, int max_cycles, int nu1, int nu2, int gamma, double* relres, double* history);
    float time_op(int op, int level, int reps);

    void sync();
    void ensure_halo(Level& lv, Which w, int depth);
    void materialize_u(Level& lv);
    void set_halo(Level& lv, Which w, int depth);
    // host-side state a captured cycle depends on and changes: which u buffer is current on each level and,
    // on distributed levels, how many halo rows of u / f / r are valid (decides which exchanges were captured)
    std::string state_blob() const;
    void set_state(const std::string& blob);

    template <typename T> void smooth_t(int level, int nu);
    template <typename T> double residual_t(int level, bool want_norm, bool store);
    template <typename T> void restrict_t(int fine_level, bool from_rhs);
    template <typename T> void prolong_t(int fine_level, bool add);
};

}  // namespace mgb
