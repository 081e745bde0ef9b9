// fused.cu — host side of the fused / temporally blocked kernels (stream.cuh) and of the
// shared-memory coarse tail (tail.cuh).  Everything here is selected by cfg.flags and
// produces results bit-identical to the unfused path in ctx.cu.
#include "fused.cuh"

#include <algorithm>

#include "comm.cuh"
#include "stream.cuh"
#include "tail.cuh"

namespace mgb {

static int g_num_sms = 148;

template <typename T, int NS, int MODE, bool RBGS>
static void set_attr()
{
    typedef StreamCfg<T, NS, MODE> C;
    MG_CK(cudaFuncSetAttribute(k_stream<T, NS, MODE, RBGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)C::SMEM_BYTES));
}

template <typename T>
static void set_attrs_t()
{
    set_attr<T, 1, MODE_SWEEPS, false>();
    set_attr<T, 2, MODE_SWEEPS, false>();
    set_attr<T, 3, MODE_SWEEPS, false>();
    set_attr<T, 1, MODE_PRE, false>();
    set_attr<T, 2, MODE_PRE, false>();
    set_attr<T, 1, MODE_POST, false>();
    set_attr<T, 2, MODE_POST, false>();
    set_attr<T, 2, MODE_SWEEPS, true>();
    set_attr<T, 4, MODE_SWEEPS, true>();
    set_attr<T, 2, MODE_PRE, true>();
    set_attr<T, 4, MODE_PRE, true>();
    set_attr<T, 2, MODE_POST, true>();
    set_attr<T, 4, MODE_POST, true>();
    MG_CK(cudaFuncSetAttribute(k_tail<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem_bytes<T>(kTailMaxLevel, 1)));
    MG_CK(cudaFuncSetAttribute(k_tail<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem_bytes<T>(kTailMaxLevel, 1)));
}

void fused_setup(Ctx& ctx)
{
    cudaDeviceProp prop;
    MG_CK(cudaGetDeviceProperties(&prop, ctx.device));
    g_num_sms = prop.multiProcessorCount;
    if (ctx.f64()) set_attrs_t<double>();
    else set_attrs_t<float>();
}

// ---------------------------------------------------------------------------------
// one launch of a streaming kernel on level lv (lcv = next coarser level for PRE/POST)
// ---------------------------------------------------------------------------------
template <typename T, int NS, int MODE, bool RBGS>
static void launch_stream(Ctx& ctx, Level& lv, Level* lcv)
{
    typedef StreamCfg<T, NS, MODE> C;
    StreamArgs<T> a;
    a.u_in = (const T*)lv.u[lv.cur];
    a.u_out = (T*)lv.u[lv.cur ^ 1];
    a.f = (const T*)lv.f;
    a.pitch = lv.pitch;
    a.N = lv.N;
    a.ya = lv.own_lo;
    a.yb = lv.own_hi;
    a.row_lo = lv.st_lo;
    a.row_hi = lv.st_hi;
    const int rows = a.yb - a.ya;
    a.strips = std::max(1, (int)cdiv(std::max(1, lv.N - C::V * C::HLANES), C::OUTW));
    // enough independent warps to fill the chip once; chunks of 16..128 rows
    const int target = g_num_sms * 16;
    int chunks = std::max(1, target / a.strips);
    int ry = (rows + chunks - 1) / chunks;
    ry = std::min(128, std::max(16, ry));
    ry = (ry + 1) & ~1;
    chunks = (rows + ry - 1) / ry;
    a.ry = ry;
    a.nitems = a.strips * chunks;
    const T om = (T)ctx.cfg.omega;
    a.c0 = (T)(1.0 - (double)om);
    a.c1 = (T)((double)om / 4.0);
    a.w = (T)ctx.cfg.restrict_weight;
    a.fc = nullptr;
    a.uc = nullptr;
    a.ec = nullptr;
    a.pitch_c = 0;
    a.Nc = 0;
    a.crow_lo = a.crow_hi = 0;
    // lazy halo exchanges (no-ops on replicated levels): rows the stencil pipeline reaches into
    ctx.ensure_halo(lv, Ctx::W_U, C::HT - (MODE == MODE_POST ? 1 : 0));
    ctx.ensure_halo(lv, Ctx::W_F, NS + (MODE == MODE_PRE ? 1 : 0) - (MODE == MODE_SWEEPS ? 1 : 0));
    if (MODE == MODE_PRE) {
        lcv->cur = 0;
        a.fc = (T*)lcv->f;
        a.uc = (lv.distributed && !lcv->distributed) ? nullptr : (T*)lcv->u[0];
        a.pitch_c = lcv->pitch;
        a.Nc = lcv->N;
    } else if (MODE == MODE_POST) {
        ctx.ensure_halo(*lcv, Ctx::W_U, NS / 2 + 1);
        a.ec = (const T*)lcv->u[lcv->cur];
        a.pitch_c = lcv->pitch;
        a.Nc = lcv->N;
        a.crow_lo = lcv->st_lo;
        a.crow_hi = lcv->st_hi;
    }
    if (rows <= 0) return;
    const unsigned grid = cdiv(a.nitems, kStreamWarps);
    k_stream<T, NS, MODE, RBGS><<<grid, kStreamWarps * 32, C::SMEM_BYTES, ctx.stream>>>(a);
    ++ctx.lc.n;
    MG_CK(cudaGetLastError());
    lv.cur ^= 1;
    lv.hv_u = 0;
    if (MODE == MODE_PRE) {
        if (lv.distributed && !lcv->distributed) {
            // agglomeration: gather the coarse right-hand side everywhere, zero guess on the full grid
            comm_allgather_rows(ctx, *lcv, lcv->f);
            MG_CK(cudaMemsetAsync(lcv->alloc[0], 0, lcv->bytes, ctx.stream));
        } else if (lcv->distributed) {
            lcv->hv_f = 0;
            comm_zero_halo(ctx, *lcv, lcv->u[0]);
            lcv->hv_u = kHaloRows;
        }
    }
}

static bool stream_ok(const Ctx& ctx, const Level&)
{
    return (ctx.cfg.flags & MG_FUSED) != 0;
}

// temporally blocked sweeps: returns the number of sweeps performed (0 = not applicable)
template <typename T>
int fused_jacobi(Ctx& ctx, Level& lv, int remaining, T, T)
{
    if (!stream_ok(ctx, lv) || remaining < 2) return 0;
    if (ctx.cfg.smoother != MG_SMOOTH_JACOBI) return 0;
    if (remaining >= 3) { launch_stream<T, 3, MODE_SWEEPS, false>(ctx, lv, nullptr); return 3; }
    launch_stream<T, 2, MODE_SWEEPS, false>(ctx, lv, nullptr);
    return 2;
}
template int fused_jacobi<double>(Ctx&, Level&, int, double, double);
template int fused_jacobi<float>(Ctx&, Level&, int, float, float);

template <typename T>
static void stream_sweeps(Ctx& ctx, Level& lv, int nu)
{
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) {
        while (nu >= 3) { launch_stream<T, 3, MODE_SWEEPS, false>(ctx, lv, nullptr); nu -= 3; }
        if (nu == 2) launch_stream<T, 2, MODE_SWEEPS, false>(ctx, lv, nullptr);
        if (nu == 1) launch_stream<T, 1, MODE_SWEEPS, false>(ctx, lv, nullptr);
    } else {
        while (nu >= 2) { launch_stream<T, 4, MODE_SWEEPS, true>(ctx, lv, nullptr); nu -= 2; }
        if (nu == 1) launch_stream<T, 2, MODE_SWEEPS, true>(ctx, lv, nullptr);
    }
}

template <typename T>
static void pre_fused(Ctx& ctx, Level& lv, Level& lcv, int nu1)
{
    const int k = std::min(nu1, 2);
    stream_sweeps<T>(ctx, lv, nu1 - k);
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) {
        if (k == 2) launch_stream<T, 2, MODE_PRE, false>(ctx, lv, &lcv);
        else launch_stream<T, 1, MODE_PRE, false>(ctx, lv, &lcv);
    } else {
        if (k == 2) launch_stream<T, 4, MODE_PRE, true>(ctx, lv, &lcv);
        else launch_stream<T, 2, MODE_PRE, true>(ctx, lv, &lcv);
    }
}

template <typename T>
static void post_fused(Ctx& ctx, Level& lv, Level& lcv, int nu2)
{
    const int k = std::min(nu2, 2);
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) {
        if (k == 2) launch_stream<T, 2, MODE_POST, false>(ctx, lv, &lcv);
        else launch_stream<T, 1, MODE_POST, false>(ctx, lv, &lcv);
    } else {
        if (k == 2) launch_stream<T, 4, MODE_POST, true>(ctx, lv, &lcv);
        else launch_stream<T, 2, MODE_POST, true>(ctx, lv, &lcv);
    }
    stream_sweeps<T>(ctx, lv, nu2 - k);
}

// ---------------------------------------------------------------------------------
// coarse tail: levels <= tail_top run in ONE launch, entirely in one CTA's shared memory
// ---------------------------------------------------------------------------------
static int tail_top(const Ctx& ctx)
{
    if (!(ctx.cfg.flags & MG_COARSE_TAIL)) return -1;
    int top = std::min(kTailMaxLevel, ctx.cfg.finest_level);
    if (top < ctx.cfg.coarsest_level) return -1;
    if (ctx.L(top).distributed) return -1;
    return top;
}

template <typename T>
static void run_tail(Ctx& ctx, int level, int nu1, int nu2, int gamma)
{
    Level& lv = ctx.L(level);
    TailArgs<T> a;
    a.top = level;
    a.coarsest = ctx.cfg.coarsest_level;
    a.nu1 = nu1;
    a.nu2 = nu2;
    a.gamma = gamma;
    const T om = (T)ctx.cfg.omega;
    a.c0 = (T)(1.0 - (double)om);
    a.c1 = (T)((double)om / 4.0);
    a.w = (T)ctx.cfg.restrict_weight;
    a.u = (T*)lv.u[lv.cur];
    a.f = (const T*)lv.f;
    a.pitch = lv.pitch;
    const size_t smem = tail_smem_bytes<T>(level, a.coarsest);
    if (ctx.cfg.smoother == MG_SMOOTH_RBGS) k_tail<T, true><<<1, kTailThreads, smem, ctx.stream>>>(a);
    else k_tail<T, false><<<1, kTailThreads, smem, ctx.stream>>>(a);
    ++ctx.lc.n;
    MG_CK(cudaGetLastError());
}

bool fused_cycle_level(Ctx& ctx, int level, int nu1, int nu2, int gamma)
{
    if (level == tail_top(ctx)) {
        if (ctx.f64()) run_tail<double>(ctx, level, nu1, nu2, gamma);
        else run_tail<float>(ctx, level, nu1, nu2, gamma);
        return true;
    }
    if (level <= ctx.cfg.coarsest_level) return false;
    Level& lv = ctx.L(level);
    Level& lcv = ctx.L(level - 1);
    if (!stream_ok(ctx, lv)) return false;
    if (nu1 < 1 || nu2 < 1) return false;
    if (ctx.f64()) pre_fused<double>(ctx, lv, lcv, nu1);
    else pre_fused<float>(ctx, lv, lcv, nu1);
    const int reps = (level - 1 <= ctx.cfg.coarsest_level) ? 1 : std::max(1, gamma);
    for (int g = 0; g < reps; ++g) ctx.cycle_rec(level - 1, nu1, nu2, gamma);
    if (ctx.f64()) post_fused<double>(ctx, lv, lcv, nu2);
    else post_fused<float>(ctx, lv, lcv, nu2);
    return true;
}

bool fused_time_hook(Ctx& ctx, int level, bool pre)
{
    if (level <= ctx.cfg.coarsest_level) return false;
    Level& lv = ctx.L(level);
    Level& lcv = ctx.L(level - 1);
    if (!stream_ok(ctx, lv)) return false;
    if (pre) {
        if (ctx.f64()) pre_fused<double>(ctx, lv, lcv, 2);
        else pre_fused<float>(ctx, lv, lcv, 2);
    } else {
        if (ctx.f64()) post_fused<double>(ctx, lv, lcv, 2);
        else post_fused<float>(ctx, lv, lcv, 2);
    }
    return true;
}

}  // namespace mgb
