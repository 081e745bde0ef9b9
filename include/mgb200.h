/* include/mgb200.h — C ABI of libmgb200.so, the B200-native (sm_100a) drop-in for the
 * geometric-multigrid Poisson path of nikhilTkur/Multigrid_Nikhil_C-.
 *
 * The reference has no FFI layer; its boundary is the C++ free-function surface of
 * Poissons_SYCL.cpp ("P:line") called by main() (P:658-731), and the context-object
 * sketch of Multigrid_functions.cpp ("M:line").  Each entry point below names the
 * reference interface it replaces.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain C: opaque context, pointers and sizes only; every call returns an int status
 *    (MG_OK == 0) and never throws; mg_last_error() gives the message of the last failure.
 *  - host vectors use the reference's layout: interior unknowns only, row-major n x n,
 *    n = 2^level - 1, index = (row-1)*n + (col-1), row <-> y (P:227-228, P:291, P:307).
 *  - the context owns all device memory, streams and CUDA graphs (the analogue of the
 *    reference's `queue&` first parameter plus its global level table P:24-33).
 *  - calls are stream-ordered and asynchronous; mg_sync() or any *_host copy / norm
 *    read-back synchronises (the reference waits after every call, P:143/P:608/P:624).
 *  - a context is not thread-safe; use one per host thread / per GPU.
 *  - there is no CPU fallback: without a CUDA device mg_create fails with MG_ERR_CUDA.
 */
#ifndef MGB200_H
#define MGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB200_VERSION 200

typedef struct mg_ctx mg_ctx;

enum { MG_OK = 0, MG_ERR_ARG = 1, MG_ERR_CUDA = 2, MG_ERR_STATE = 3, MG_ERR_COMM = 4, MG_ERR_ALLOC = 5 };
enum { MG_F64 = 0, MG_F32 = 1 };                 /* reference P is fp32 (E11), M is fp64 */
enum { MG_SMOOTH_JACOBI = 0, MG_SMOOTH_RBGS = 1 };
enum { MG_COARSE_SWEEPS = 0, /* coarsest level: nu1 + nu2 smoothing sweeps (P:583-587)                                  */
       MG_COARSE_EXACT = 1   /* coarsest level: exact solve, no smoothing (direct_solver M:63-72, called at M:136-139) */ };
enum { MG_GRAPH = 1,        /* replay whole cycles as CUDA graphs                         */
       MG_FUSED = 2,        /* use the fused / temporally blocked kernels where they apply */
       MG_COARSE_TAIL = 4   /* run the levels that fit one CTA's shared memory in one launch */ };

/* Replaces the compile-time globals P:17-22, P:127 and the level loop of main()
 * P:661-690 (one operator table entry per level; here matrix-free, so "setup" only
 * allocates u/f/r per level). */
typedef struct mg_config {
    int finest_level;          /* P:17  (N = 2^level + 1 nodes per side)                    */
    int coarsest_level;        /* P:18  (reference finest-3; default 1)                     */
    int dtype;                 /* MG_F64 | MG_F32                                           */
    int smoother;              /* MG_SMOOTH_JACOBI (P:125-147) | MG_SMOOTH_RBGS             */
    double omega;              /* P:127, 2/3                                                */
    double restrict_weight;    /* 0.25: P:539 with E2+E4 repaired; 0.0625 = literal FD weight */
    int device;                /* CUDA ordinal, -1 = current device                         */
    int flags;                 /* MG_GRAPH | MG_FUSED | MG_COARSE_TAIL                      */
    /* row-slab decomposition over the GPUs of one box (no reference counterpart) */
    int rank, world;           /* this context owns slab `rank` of `world`                  */
    int agglomerate_level;     /* levels <= this are solved redundantly on every rank; 0 = auto */
    const void* comm_id;       /* mg_comm_id() bytes from rank 0, required when world > 1   */
    int coarse_solver;         /* MG_COARSE_SWEEPS (P:583-587) | MG_COARSE_EXACT (M:63-72; coarsest_level <= 9) */
} mg_config;

void mg_config_default(mg_config* cfg);

int mg_create(mg_ctx** out, const mg_config* cfg);                 /* main() level loop P:661-690 */
int mg_destroy(mg_ctx* ctx);
const char* mg_last_error(const mg_ctx* ctx);                      /* ctx may be NULL: last create error */
int mg_sync(mg_ctx* ctx);                                          /* event.wait() P:143, P:608, P:624 */

/* 128-byte communicator id (rank 0 creates, the launcher broadcasts it to all ranks). */
#define MG_COMM_ID_BYTES 128
int mg_comm_id(void* out128);

/* ---- level queries (reference: int(log2(sqrt(size)+1)), P:583) ---- */
int mg_level_side(int level);                                      /* n = 2^level - 1              */
int mg_level_of_size(size_t vec_size);                             /* -1 if not (2^L-1)^2          */
/* owned interior rows [row_begin,row_end) (1-based node rows) of `rank` at `level` */
int mg_slab_rows(int level, int rank, int world, int* row_begin, int* row_end);
int mg_get_info(const mg_ctx* ctx, int what, int level, int64_t* out);
enum { MG_INFO_PITCH = 0, MG_INFO_ROWS_STORED = 1, MG_INFO_ROW_BEGIN = 2, MG_INFO_ROW_END = 3,
       MG_INFO_LAUNCHES = 4, MG_INFO_DISTRIBUTED = 5, MG_INFO_BYTES_ALLOCATED = 6,
       MG_INFO_GRAPH_LAUNCHES = 7, MG_INFO_AGGLOMERATE_LEVEL = 8,
       MG_INFO_STORED_ROW_BEGIN = 9, MG_INFO_STORED_ROW_END = 10 /* owned + halo / ring rows kept by this rank */ };

/* ---- data movement; host vectors are FULL-grid interior vectors (n*n), each rank
 *      takes / fills the rows of its slab (mg_get_* gathers nothing across ranks:
 *      rows outside the slab are left untouched) ---- */
int mg_force_constant(mg_ctx* ctx, double f);                      /* globalforcefunction P:283-335: b = f*h^2 on the finest level */
/* Synthetic right-hand side for benchmarks, generated on the device (no reference counterpart; the reference only has the
 * constant f = 4, P:123): b = h^2 (2U - 1), U = (splitmix64(seed + (idx+1)*0x9E3779B97F4A7C15) >> 11) * 2^-53, idx = the
 * reference's interior index (row-1)*n + (col-1) (P:227-228).  A function of the GLOBAL index only: every rank of a
 * row-slab run and a single-GPU run of the same grid hold identical values (tests restate it in numpy). */
int mg_force_synthetic(mg_ctx* ctx, uint64_t seed);
/* Order-independent 64-bit checksum of this rank's OWNED interior values of u (which = 0), f (1) or r (2) on `level`:
 * sum mod 2^64 of splitmix64(value bits + (idx+1)*0x9E37...) over the points.  The sum over the ranks of a row-slab run
 * equals the single-GPU checksum exactly when every value agrees bit for bit (bench.py's multi-GPU parity record). */
int mg_checksum(mg_ctx* ctx, int level, int which, uint64_t* out);
int mg_set_rhs_host(mg_ctx* ctx, int level, const void* f_host);   /* f_h argument of P:575 / P:629 */
int mg_set_u_host(mg_ctx* ctx, int level, const void* u_host);     /* vec_h argument of P:575       */
int mg_get_u_host(mg_ctx* ctx, int level, void* u_host);           /* returned vector P:626, P:649  */
int mg_get_rhs_host(mg_ctx* ctx, int level, void* f_host);
int mg_get_r_host(mg_ctx* ctx, int level, void* r_host);           /* `residual` P:594/P:607        */
int mg_zero_u(mg_ctx* ctx, int level);                             /* P:613, P:630                  */

/* ---- grid operators ---- */
int mg_smooth(mg_ctx* ctx, int level, int nu);                     /* jacobirelaxation P:125-147 (or RB-GS) */
int mg_residual(mg_ctx* ctx, int level, double* norm2_or_null);    /* P:589-608; optional ||r||_2 (whole grid) */
int mg_restrict(mg_ctx* ctx, int fine_level);                      /* restriction2d P:531-546 of r -> f[level-1]; zero coarse u (P:613) */
int mg_restrict_rhs(mg_ctx* ctx, int fine_level);                  /* restriction2d of f -> f[level-1] (FMG, P:641) */
int mg_prolong_correct(mg_ctx* ctx, int fine_level);               /* interpolation2d P:337-425 + vm::add P:620-624 */
int mg_prolong_set(mg_ctx* ctx, int fine_level);                   /* bare interpolation2d as FMG initial guess P:645 */

/* ---- cycles (rows a7-a10) ---- */
int mg_cycle(mg_ctx* ctx, int level, int nu1, int nu2, int gamma); /* vcyclemultigrid P:575-627; gamma=2 => W */
/* `count` consecutive cycles on the same right-hand side == the loop P:646-648 (`for i <= mu0: vec_h = vcyclemultigrid(...)`).
 * Same result as `count` mg_cycle calls, bit for bit; lets the library keep the iterate on chip between the post-smoothing
 * of one cycle and the pre-smoothing of the next (visit chains, the default) and replay the whole run as one CUDA graph. */
int mg_cycles(mg_ctx* ctx, int level, int nu1, int nu2, int gamma, int count);
int mg_fmg(mg_ctx* ctx, int cycles_per_level, int nu1, int nu2);   /* fullmultigrid P:629-650 (reference cycles = mu0+1) */
int mg_solve(mg_ctx* ctx, double rtol, int max_cycles, int nu1, int nu2, int gamma,
             int* cycles_out, double* relres_out, double* history_or_null /* max_cycles+1 */);

/* ---- host-vector entry points with the reference's call shape (copies inside):
 *      one call == one reference call on std::vector arguments ---- */
int mg_host_jacobirelaxation(mg_ctx* ctx, int level, void* v_inout, const void* fh, int mu);        /* P:125 */
int mg_host_restriction2d(mg_ctx* ctx, int fine_level, const void* vec_h, void* vec_2h);            /* P:531 */
int mg_host_interpolation2d(mg_ctx* ctx, int fine_level, const void* vec_2h, void* vec_h);          /* P:337 */
int mg_host_vcyclemultigrid(mg_ctx* ctx, int level, void* vec_h_inout, const void* f_h,
                            int nu1, int nu2, int gamma);                                           /* P:575 */
int mg_host_fullmultigrid(mg_ctx* ctx, const void* f_h, void* vec_h_out, int cycles_per_level,
                          int nu1, int nu2);                                                        /* P:629 */

/* ---- communication-avoiding row-slab schedule (csrc/sched.h; pure host logic, callable without a GPU).
 *      Fills `ops` with 4 ints per op {kind, level, a, b} (kinds: 0 EXCH a=which(0 u,1 f) b=depth; 1 PRE a=ya b=yb;
 *      2 POST a=ya b=yb; 3 GATHER_F; 4 REPL_CYCLE) and `halo32` with the halo rows each level must store.
 *      Returns the number of ops, or -1 when the schedule is not applicable (then the lazy-exchange path runs). ---- */
int mg_plan_vcycle(int top_level, int agglomerate_level, int world, int rank, int ns_pre, int ns_post,
                   int valid_halo_u_top, int valid_halo_f_top, int* ops, int max_ops, int* halo32);

/* ---- micro-benchmark hooks used by bench.py (device-timed, CUDA events on the
 *      context's stream; returns milliseconds for `reps` back-to-back launches) ---- */
int mg_time_op(mg_ctx* ctx, int op, int level, int reps, float* ms_out);
/* `reps` back-to-back mg_cycle calls bracketed by CUDA events on the context's stream */
int mg_time_cycle(mg_ctx* ctx, int level, int nu1, int nu2, int gamma, int reps, float* ms_out);
/* Row slabs (world > 1), communication-avoiding schedule: device milliseconds per cycle of every phase of the plan, from
 * `reps` eagerly launched cycles with a pair of CUDA events around each op (launch gaps of eager launches included).
 * out_ms_5x32[kind * 32 + level], kinds as in mg_plan_vcycle: 0 halo exchange, 1 PRE, 2 POST, 3 all-gather of the first
 * replicated level's right-hand side, 4 the replicated coarse cycle.  *ops_per_cycle = 0 when this context does not run
 * the plan (one GPU, lazy schedule). */
int mg_time_phases(mg_ctx* ctx, int level, int nu1, int nu2, int gamma, int reps, double* out_ms_5x32, int* ops_per_cycle);
enum { MG_OP_SMOOTH1 = 0, MG_OP_RESIDUAL = 1, MG_OP_RESTRICT = 2, MG_OP_PROLONG = 3,
       MG_OP_PRE_FUSED = 4, MG_OP_POST_FUSED = 5, MG_OP_RESIDUAL_NORM = 6, MG_OP_SMOOTH2 = 7,
       MG_OP_SMOOTH3 = 8, MG_OP_SMOOTH4 = 9, /* k sweeps temporally blocked in ONE launch (k = 4: Jacobi only) */
       MG_OP_POSTPRE_FUSED = 10 /* POST + PRE of consecutive visits in one launch (needs MGB200_CHAIN=1) */ };

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
