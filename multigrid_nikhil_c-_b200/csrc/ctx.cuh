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

// a pair of CUDA events around a timed region on one stream; destroyed on every path, including exceptions
struct EventTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t stream;
    explicit EventTimer(cudaStream_t s) : stream(s)
    {
        MG_CK(cudaEventCreate(&e0));
        cudaError_t e = cudaEventCreate(&e1);
        if (e != cudaSuccess) { cudaEventDestroy(e0); e0 = nullptr; MG_CK(e); }
    }
    ~EventTimer()
    {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    EventTimer(const EventTimer&) = delete;
    EventTimer& operator=(const EventTimer&) = delete;
    void start() { MG_CK(cudaEventRecord(e0, stream)); }
    float stop()   // milliseconds since start(); waits for the region to finish
    {
        MG_CK(cudaEventRecord(e1, stream));
        MG_CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        MG_CK(cudaEventElapsedTime(&ms, e0, e1));
        return ms;
    }
};

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
    // zero-guess chain: the current u is LOGICALLY zero but its buffer was not written (the next kernel is a
    // zero-guess variant that does not read it); Ctx::materialize_u writes the zeros for every other reader
    bool u_zero = false;
    // visit chains, fullmultigrid: the current u is LOGICALLY the bare interpolation of the coarser level's iterate
    // (P:645) but was not computed: the first PRE of the level does it on the fly (k_stream_fmg_entry);
    // Ctx::materialize_u runs the real prolongation for every other reader
    bool u_interp = false;
};

// per-context tuning state of the streaming kernels (fused.cu).  Per context, not per process: two contexts on two
// devices of one process must not share the SM count or the knobs read at creation.
struct FusedKnobs {
    int num_sms = 148;
    int occ = 12;              // resident streaming warps per SM (MGB200_STREAM_OCC), enforced by padding dynamic smem
    bool autotune = true;      // MGB200_AUTOTUNE=0 disables the chunk-height tuner
    int force_ry = 0;          // MGB200_STREAM_RY / _MINN: force the chunk height on levels with N >= minN (tuning sweeps)
    int force_ry_minN = 4096;
    bool pdl = true;           // programmatic dependent launches between the kernels of a cycle (MGB200_PDL=0 turns them off)
};

// Staging of PAGEABLE host vectors (the std::vector call shape of include/mgb200_driver.hpp): a cudaMemcpy from pageable
// memory is copied through a driver bounce buffer by one thread; here a few host threads copy row chunks through the
// context's own pinned buffers on their own streams, so the copy engines stay busy.  Pinned callers skip all of this.
struct Stager {
    static constexpr int kThreads = 4, kBufs = 2;   // 4 threads: 15.6 ms for 402 MB at 4097^2; 8 threads measured no better
    static constexpr size_t kChunkBytes = 4u << 20;
    static constexpr size_t kMinBytes = 8u << 20;      // smaller transfers take the plain path
    char* pinned[kThreads][kBufs] = {};
    cudaStream_t st[kThreads] = {};
    cudaEvent_t ev[kThreads][kBufs] = {};
    cudaEvent_t ev_main = nullptr, ev_done[kThreads] = {};
    bool ready = false;
};

// one timed op of the communication-avoiding plan (mg_time_phases)
struct PhaseRec {
    int kind = 0, level = 0;           // SchedKind, level
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    long long kernels = 0;
    std::string state_after;   // Ctx::state_blob() once the captured cycle has run
    bool post_norm = false;    // the captured cycle's last kernel also produced the residual norm
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
    void* d_dst_S = nullptr;   // MG_COARSE_EXACT: sine-transform table of the coarsest level (n x n) ...
    void* d_dst_d = nullptr;   // ... and the 1-D eigenvalues d[j] (coarse.cuh)
    bool capturing = false;
    std::map<std::tuple<int, int, int, int, std::string>, GraphEntry> graphs;
    std::map<std::tuple<int, int, int, int>, int> stream_ry;  // tuned chunk height per (level, mode, NS, rbgs)
    FusedKnobs knobs;
    Comm* comm = nullptr;
    int aggl_level = 0;  // levels <= aggl_level are replicated on every rank
    bool graph_dist = true;   // capture NCCL exchanges into cycle graphs (MGB200_GRAPH_DIST=0 turns it off)
    bool zero_guess = true;   // skip reading / writing the zero coarse guess (P:613) where the next kernel does not need it (MGB200_ZERO_GUESS=0 turns it off)
    bool comm_avoid = true;   // communication-avoiding slab schedule (csrc/sched.h; MGB200_COMM_AVOID=0 turns it off)
    // tolerance loop: the cycle's last kernel on the finest level also leaves sum r^2 in d_norm (fused.cu: launch_post_norm)
    bool want_post_norm = false, post_norm_done = false;
    bool chain = true;        // fuse POST of one visit of a level with PRE of the next (stream.cuh MODE_POSTPRE; MGB200_CHAIN=0 turns it off)

    explicit Ctx(const mg_config& c);
    ~Ctx();
    void init();               // body of the constructor
    void release() noexcept;   // frees everything (destructor, and the constructor's error path)

    Level& L(int level);
    const Level& L(int level) const;
    bool f64() const { return cfg.dtype == MG_F64; }

    // data movement
    enum Which { W_U = 0, W_F = 1, W_R = 2 };
    // device-loop solve (ctx.cu: solve_device_loop): control block, body-capture stream, cached graphs
    void* d_solve_ctl = nullptr;
    int solve_ctl_cap = 0;
    cudaStream_t body_stream = nullptr;
    std::map<std::tuple<int, int, int, int, std::string>, std::pair<cudaGraphExec_t, std::string>> solve_graphs;
    double history_last = 0.0;
    int solve_loop_mode = -1;   // -1: read MGB200_SOLVE_GRAPH on first use; 0 host loop; 1 device loop
    Stager stager;
    void stager_init();
    // rows [ya, yb) of a padded device array <-> the same rows of an n x n host vector in ordinary (pageable) memory
    void copy_rows_staged(char* dev_row0, size_t dev_pitch_bytes, char* host_row0, size_t row_bytes, int rows, bool to_device);
    void copy_rows(char* dev_row0, size_t dev_pitch_bytes, char* host_row0, size_t row_bytes, int rows, bool to_device);
    void set_host(int level, Which w, const void* host);
    void get_host(int level, Which w, void* host);
    void zero_u(int level);
    void force_constant(double f);
    void force_synthetic(unsigned long long seed);          // b = h^2 (2U-1), U from splitmix64(seed, global index)
    unsigned long long checksum(int level, Which w);        // order-independent checksum of this rank's owned rows

    // operators
    void smooth(int level, int nu);
    double residual(int level, bool want_norm, bool store);
    void restrict_to(int fine_level, bool from_rhs);
    void prolong(int fine_level, bool add);
    void coarse_exact(int level);                            // direct solve on the coarsest level (M:63-72), coarse.cuh
    bool exact_coarse() const { return cfg.coarse_solver == MG_COARSE_EXACT; }

    // cycles
    void cycle(int level, int nu1, int nu2, int gamma);
    void cycles(int level, int nu1, int nu2, int gamma, int count);   // `count` consecutive cycles (P:646-648 loop)
    void cycle_rec(int level, int nu1, int nu2, int gamma);
    void cycle_rec_visits(int level, int nu1, int nu2, int gamma, int visits);   // consecutive visits of one level
    void fmg(int cycles, int nu1, int nu2);
    int solve(double rtol, int max_cycles, int nu1, int nu2, int gamma, double* relres, double* history);
    float time_op(int op, int level, int reps);
    // per-phase device time of the communication-avoiding cycle (world > 1): `reps` EAGER cycles with a pair of events around
    // every op of the plan; out[kind * 32 + level] = milliseconds per cycle (kinds: sched.h SchedKind).  Returns the number
    // of timed ops per cycle (0: this context does not run the plan).
    int time_phases(int level, int nu1, int nu2, int gamma, int reps, double* out160);
    bool phase_on = false, force_eager = false;
    std::vector<PhaseRec> phase_log;
    double read_norm(const Level& lv);
    // the whole tolerance loop as ONE graph launch: a conditional WHILE node whose body is the cycle (device-side decision)
    bool solve_device_loop(double rtol, int max_cycles, int nu1, int nu2, int gamma, double r0, int* k_out, double* history);

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
