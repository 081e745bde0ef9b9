// capi.cu — the extern "C" boundary (include/mgb200.h).  Every entry point converts
// exceptions into a status code and records the message in the context.
#include <cmath>
#include <cstring>
#include <mutex>

#include "comm.cuh"
#include "ctx.cuh"
#include "sched.h"

using namespace mgb;

static std::string g_create_error;
static std::mutex g_mu;

template <typename F>
static int guarded(mg_ctx* c, F&& fn)
{
    Ctx* ctx = reinterpret_cast<Ctx*>(c);
    if (!ctx) return MG_ERR_ARG;
    try {
        cudaSetDevice(ctx->device);
        fn(*ctx);
        return MG_OK;
    } catch (const MgError& e) {
        ctx->err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        ctx->err = e.what();
        return MG_ERR_STATE;
    } catch (...) {
        ctx->err = "unknown error";
        return MG_ERR_STATE;
    }
}

extern "C" {

void mg_config_default(mg_config* cfg)
{
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->finest_level = 10;       // P:17
    cfg->coarsest_level = 1;      // reference P:18 uses finest-3; see SURVEY E5
    cfg->dtype = MG_F64;
    cfg->smoother = MG_SMOOTH_JACOBI;
    cfg->omega = 2.0 / 3.0;       // P:127
    cfg->restrict_weight = 0.25;  // P:539 with E2+E4 repaired
    cfg->device = -1;
    cfg->flags = MG_GRAPH | MG_FUSED | MG_COARSE_TAIL;
    cfg->rank = 0;
    cfg->world = 1;
    cfg->agglomerate_level = 0;
    cfg->comm_id = nullptr;
    cfg->coarse_solver = MG_COARSE_SWEEPS;   // P:583-587
}

int mg_create(mg_ctx** out, const mg_config* cfg)
{
    if (!out || !cfg) return MG_ERR_ARG;
    *out = nullptr;
    try {
        Ctx* ctx = new Ctx(*cfg);
        *out = reinterpret_cast<mg_ctx*>(ctx);
        return MG_OK;
    } catch (const MgError& e) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_create_error = e.what();
        return e.code;
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_create_error = e.what();
        return MG_ERR_STATE;
    } catch (...) {   // nothing may cross the C boundary
        std::lock_guard<std::mutex> lk(g_mu);
        g_create_error = "unknown error";
        return MG_ERR_STATE;
    }
}

int mg_destroy(mg_ctx* c)
{
    if (!c) return MG_ERR_ARG;
    delete reinterpret_cast<Ctx*>(c);
    return MG_OK;
}

const char* mg_last_error(const mg_ctx* c)
{
    if (!c) return g_create_error.c_str();
    return reinterpret_cast<const Ctx*>(c)->err.c_str();
}

int mg_sync(mg_ctx* c) { return guarded(c, [&](Ctx& x) { x.sync(); }); }

int mg_comm_id(void* out128)
{
    if (!out128) return MG_ERR_ARG;
    try {
        return comm_unique_id(out128);
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_create_error = e.what();
        return MG_ERR_COMM;
    } catch (...) {
        return MG_ERR_COMM;
    }
}

int mg_level_side(int level) { return (level >= 1 && level <= 30) ? (1 << level) - 1 : -1; }

int mg_level_of_size(size_t vec_size)
{
    for (int l = 1; l <= 30; ++l) {
        const size_t n = ((size_t)1 << l) - 1;
        if (n * n == vec_size) return l;
        if (n * n > vec_size) break;
    }
    return -1;
}

int mg_slab_rows(int level, int rank, int world, int* row_begin, int* row_end)
{
    if (level < 1 || level > 30 || world < 1 || rank < 0 || rank >= world || !row_begin || !row_end) return MG_ERR_ARG;
    if ((((long long)1 << level) % world) != 0) return MG_ERR_ARG;
    slab_rows(level, rank, world, row_begin, row_end);
    return MG_OK;
}

int mg_get_info(const mg_ctx* c, int what, int level, int64_t* out)
{
    if (!c || !out) return MG_ERR_ARG;
    return guarded(const_cast<mg_ctx*>(c), [&](Ctx& x) {
        switch (what) {
            case MG_INFO_PITCH: *out = x.L(level).pitch; break;
            case MG_INFO_ROWS_STORED: *out = x.L(level).st_hi - x.L(level).st_lo; break;
            case MG_INFO_ROW_BEGIN: *out = x.L(level).own_lo; break;
            case MG_INFO_ROW_END: *out = x.L(level).own_hi; break;
            case MG_INFO_LAUNCHES: *out = x.lc.n; break;
            case MG_INFO_DISTRIBUTED: *out = x.L(level).distributed ? 1 : 0; break;
            case MG_INFO_BYTES_ALLOCATED: *out = (int64_t)x.bytes_allocated; break;
            case MG_INFO_GRAPH_LAUNCHES: *out = x.graph_launches; break;
            case MG_INFO_AGGLOMERATE_LEVEL: *out = x.aggl_level; break;
            case MG_INFO_STORED_ROW_BEGIN: *out = x.L(level).st_lo; break;
            case MG_INFO_STORED_ROW_END: *out = x.L(level).st_hi; break;
            default: throw MgError(MG_ERR_ARG, "unknown info key");
        }
    });
}

int mg_force_constant(mg_ctx* c, double f) { return guarded(c, [&](Ctx& x) { x.force_constant(f); }); }
int mg_force_synthetic(mg_ctx* c, uint64_t seed) { return guarded(c, [&](Ctx& x) { x.force_synthetic(seed); }); }
int mg_checksum(mg_ctx* c, int level, int which, uint64_t* out)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(out != nullptr && which >= 0 && which <= 2, "null out or which not in {0 u, 1 f, 2 r}");
        *out = x.checksum(level, (Ctx::Which)which);
    });
}
int mg_set_rhs_host(mg_ctx* c, int level, const void* p) { return guarded(c, [&](Ctx& x) { x.set_host(level, Ctx::W_F, p); }); }
int mg_set_u_host(mg_ctx* c, int level, const void* p) { return guarded(c, [&](Ctx& x) { x.set_host(level, Ctx::W_U, p); }); }
int mg_get_u_host(mg_ctx* c, int level, void* p) { return guarded(c, [&](Ctx& x) { x.get_host(level, Ctx::W_U, p); }); }
int mg_get_rhs_host(mg_ctx* c, int level, void* p) { return guarded(c, [&](Ctx& x) { x.get_host(level, Ctx::W_F, p); }); }
int mg_get_r_host(mg_ctx* c, int level, void* p) { return guarded(c, [&](Ctx& x) { x.get_host(level, Ctx::W_R, p); }); }
int mg_zero_u(mg_ctx* c, int level) { return guarded(c, [&](Ctx& x) { x.zero_u(level); }); }

int mg_smooth(mg_ctx* c, int level, int nu)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(nu >= 0, "nu >= 0 required");
        x.L(level);
        x.smooth(level, nu);
    });
}

int mg_residual(mg_ctx* c, int level, double* norm2)
{
    return guarded(c, [&](Ctx& x) {
        x.L(level);
        const double v = x.residual(level, norm2 != nullptr, true);
        if (norm2) *norm2 = v;
    });
}

int mg_restrict(mg_ctx* c, int fine_level) { return guarded(c, [&](Ctx& x) { x.L(fine_level); x.restrict_to(fine_level, false); }); }
int mg_restrict_rhs(mg_ctx* c, int fine_level) { return guarded(c, [&](Ctx& x) { x.L(fine_level); x.restrict_to(fine_level, true); }); }
int mg_prolong_correct(mg_ctx* c, int fine_level) { return guarded(c, [&](Ctx& x) { x.L(fine_level); x.prolong(fine_level, true); }); }
int mg_prolong_set(mg_ctx* c, int fine_level) { return guarded(c, [&](Ctx& x) { x.L(fine_level); x.prolong(fine_level, false); }); }

int mg_cycle(mg_ctx* c, int level, int nu1, int nu2, int gamma)
{
    return guarded(c, [&](Ctx& x) { x.cycle(level, nu1, nu2, gamma); });
}

int mg_cycles(mg_ctx* c, int level, int nu1, int nu2, int gamma, int count)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(count >= 0, "count >= 0 required");
        x.cycles(level, nu1, nu2, gamma, count);
    });
}

int mg_fmg(mg_ctx* c, int cycles, int nu1, int nu2) { return guarded(c, [&](Ctx& x) { x.fmg(cycles, nu1, nu2); }); }

int mg_solve(mg_ctx* c, double rtol, int max_cycles, int nu1, int nu2, int gamma, int* cycles_out,
             double* relres_out, double* history)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(max_cycles >= 0 && nu1 >= 0 && nu2 >= 0 && gamma >= 1, "bad solve parameters");
        const int k = x.solve(rtol, max_cycles, nu1, nu2, gamma, relres_out, history);
        if (cycles_out) *cycles_out = k;
    });
}

// ---- host-vector entry points, one call == one reference call ----
int mg_host_jacobirelaxation(mg_ctx* c, int level, void* v, const void* fh, int mu)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(mu >= 0, "mu >= 0 required");
        x.set_host(level, Ctx::W_U, v);
        x.set_host(level, Ctx::W_F, fh);
        x.smooth(level, mu);
        x.get_host(level, Ctx::W_U, v);
    });
}

int mg_host_restriction2d(mg_ctx* c, int fine_level, const void* vec_h, void* vec_2h)
{
    return guarded(c, [&](Ctx& x) {
        x.L(fine_level);
        x.set_host(fine_level, Ctx::W_R, vec_h);
        x.restrict_to(fine_level, false);
        x.get_host(fine_level - 1, Ctx::W_F, vec_2h);
    });
}

int mg_host_interpolation2d(mg_ctx* c, int fine_level, const void* vec_2h, void* vec_h)
{
    return guarded(c, [&](Ctx& x) {
        x.L(fine_level);
        MG_REQUIRE(fine_level > x.cfg.coarsest_level, "no coarser level below coarsest_level");
        x.set_host(fine_level - 1, Ctx::W_U, vec_2h);
        x.prolong(fine_level, false);
        x.get_host(fine_level, Ctx::W_U, vec_h);
    });
}

int mg_host_vcyclemultigrid(mg_ctx* c, int level, void* vec_h, const void* f_h, int nu1, int nu2, int gamma)
{
    return guarded(c, [&](Ctx& x) {
        x.set_host(level, Ctx::W_U, vec_h);
        x.set_host(level, Ctx::W_F, f_h);
        x.cycle(level, nu1, nu2, gamma);
        x.get_host(level, Ctx::W_U, vec_h);
    });
}

int mg_host_fullmultigrid(mg_ctx* c, const void* f_h, void* vec_h_out, int cycles, int nu1, int nu2)
{
    return guarded(c, [&](Ctx& x) {
        x.set_host(x.cfg.finest_level, Ctx::W_F, f_h);
        x.fmg(cycles, nu1, nu2);
        x.get_host(x.cfg.finest_level, Ctx::W_U, vec_h_out);
    });
}

int mg_plan_vcycle(int top_level, int agglomerate_level, int world, int rank, int ns_pre, int ns_post,
                   int valid_halo_u_top, int valid_halo_f_top, int* ops, int max_ops, int* halo32)
{
    if (top_level < 1 || top_level > 30 || world < 1 || rank < 0 || rank >= world || !ops || !halo32) return -1;
    const SchedPlan p = sched_plan_vcycle(top_level, agglomerate_level, world, rank, ns_pre, ns_post, valid_halo_u_top,
                                          valid_halo_f_top);
    if (!p.ok || (int)p.ops.size() > max_ops) return -1;
    for (size_t i = 0; i < p.ops.size(); ++i) {
        ops[4 * i + 0] = p.ops[i].kind;
        ops[4 * i + 1] = p.ops[i].level;
        ops[4 * i + 2] = p.ops[i].a;
        ops[4 * i + 3] = p.ops[i].b;
    }
    for (int l = 0; l < 32; ++l) halo32[l] = p.halo[l];
    return (int)p.ops.size();
}

int mg_time_op(mg_ctx* c, int op, int level, int reps, float* ms_out)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(ms_out != nullptr, "null ms_out");
        *ms_out = x.time_op(op, level, reps);
    });
}

int mg_time_phases(mg_ctx* c, int level, int nu1, int nu2, int gamma, int reps, double* out_ms_5x32, int* ops_per_cycle)
{
    return guarded(c, [&](Ctx& x) {
        const int n = x.time_phases(level, nu1, nu2, gamma, reps, out_ms_5x32);
        if (ops_per_cycle) *ops_per_cycle = n;
    });
}

int mg_time_cycle(mg_ctx* c, int level, int nu1, int nu2, int gamma, int reps, float* ms_out)
{
    return guarded(c, [&](Ctx& x) {
        MG_REQUIRE(ms_out != nullptr && reps >= 1, "null ms_out or reps < 1");
        EventTimer t(x.stream);
        t.start();
        x.cycles(level, nu1, nu2, gamma, reps);
        *ms_out = t.stop();
    });
}

}  // extern "C"
