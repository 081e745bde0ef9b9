// fused.cu — host side of the fused / temporally blocked kernels (stream.cuh) and of the
// shared-memory coarse tail (tail.cuh).  Everything here is selected by cfg.flags and
// produces results bit-identical to the unfused path in ctx.cu.
#include "fused.cuh"

#include <algorithm>
#include <type_traits>
#include <vector>
#include <cstdlib>

#include "comm.cuh"
#include "sched.h"
#include "stream.cuh"
#include "tail.cuh"

namespace mgb {

template <typename T, int NS, int MODE, bool RBGS>
static void set_attr()
{
    typedef StreamCfg<T, NS, MODE> C;
    MG_CK(cudaFuncSetAttribute(k_stream<T, NS, MODE, RBGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)std::max<size_t>(C::SMEM_BYTES, 100 * 1024)));
}

template <typename T>
static void set_attrs_t()
{
    set_attr<T, 1, MODE_SWEEPS, false>();
    set_attr<T, 2, MODE_SWEEPS, false>();
    set_attr<T, 3, MODE_SWEEPS, false>();
    set_attr<T, 4, MODE_SWEEPS, false>();
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

template <typename T>
static void set_attrs_chain()
{
    const int big = 100 * 1024;
    MG_CK(cudaFuncSetAttribute(k_stream_norm<T, 1, MODE_POST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_norm<T, 2, MODE_POST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_norm<T, 2, MODE_POST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_norm<T, 4, MODE_POST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_chain<T, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_chain<T, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_chain<T, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_chain<T, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_fmg_entry<T, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_fmg_entry<T, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_fmg_entry<T, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_fmg_entry<T, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_pre_zg<T, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_pre_zg<T, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_pre_zg<T, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    MG_CK(cudaFuncSetAttribute(k_stream_pre_zg<T, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    const int ts = (int)tail_smem_bytes<T>(kTailMaxLevel, 1);
    MG_CK(cudaFuncSetAttribute(k_tail_zg<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts));
    MG_CK(cudaFuncSetAttribute(k_tail_zg<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ts));
}

// Per CONTEXT, on the context's device (function attributes are per device; a second context on another device of the
// same process gets its own), outside any stream capture.  Run-time knobs are read afresh for every context.
void fused_setup(Ctx& ctx)
{
    cudaDeviceProp prop;
    MG_CK(cudaGetDeviceProperties(&prop, ctx.device));
    FusedKnobs& k = ctx.knobs;
    k.num_sms = prop.multiProcessorCount;
    auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    k.force_ry = env_int("MGB200_STREAM_RY", 0);
    k.force_ry_minN = env_int("MGB200_STREAM_RY_MINN", 4096);
    k.occ = std::max(1, env_int("MGB200_STREAM_OCC", 12));
    k.autotune = env_int("MGB200_AUTOTUNE", 1) != 0;
    k.pdl = env_int("MGB200_PDL", 1) != 0;   // measured: -2.6 % on the 4097^2 cycle, -1.5 % on the 8193^2 W-cycle (profiles/r02_kernel_experiments.md)
    if (ctx.f64()) { set_attrs_t<double>(); set_attrs_chain<double>(); }
    else { set_attrs_t<float>(); set_attrs_chain<float>(); }
}

// ---------------------------------------------------------------------------------
// one launch of a streaming kernel on level lv (lcv = next coarser level for PRE/POST)
// ---------------------------------------------------------------------------------
template <typename T, int NS, int MODE>
static int default_ry(const Ctx& ctx, const Level& lv, int strips)
{
    // One wave of knobs.occ warps per SM: a partial last wave costs a whole chunk time, so the chunk height is
    // chosen to make strips x chunks land on one wave whenever that keeps chunks <= 256 rows.
    const FusedKnobs& kn = ctx.knobs;
    const int rows = lv.own_hi - lv.own_lo;
    const int wave = kn.num_sms * kn.occ;
    const int chunks = std::max(1, wave / strips);
    int ry = (rows + chunks - 1) / chunks;
    ry = std::max(8, std::min(256, ry));
    if (kn.force_ry > 0 && lv.N >= kn.force_ry_minN) ry = kn.force_ry;   // MGB200_STREAM_RY[_MINN]: tuning knobs
    return (ry + 1) & ~1;
}

template <typename T, int NS, int MODE, bool NORM = false>
static StreamArgs<T> make_args(Ctx& ctx, Level& lv, Level* lcv, int ry, int ya = -1, int yb = -1, bool write_zero_guess = true)
{
    typedef StreamCfg<T, NS, MODE, NORM> C;
    StreamArgs<T> a;
    a.u_in = (const T*)lv.u[lv.cur];
    a.u_out = (T*)lv.u[lv.cur ^ 1];
    a.f = (const T*)lv.f;
    a.pitch = lv.pitch;
    a.N = lv.N;
    a.ya = (ya >= 0) ? ya : lv.own_lo;
    a.yb = (yb >= 0) ? yb : lv.own_hi;
    a.row_lo = lv.st_lo;
    a.row_hi = lv.st_hi;
    const int rows = a.yb - a.ya;
    a.strips = std::max(1, (int)cdiv(std::max(1, lv.N - C::V * C::HLANES), C::OUTW));
    a.strips_pad = (a.strips + kStreamWarps - 1) / kStreamWarps * kStreamWarps;
    if (ry <= 0) ry = default_ry<T, NS, MODE>(ctx, lv, a.strips);
    a.ry = ry;
    a.nitems = a.strips_pad * ((rows + ry - 1) / ry);
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
    if (MODE == MODE_PRE) {
        a.fc = (T*)lcv->f;
        a.uc = ((lv.distributed && !lcv->distributed) || !write_zero_guess) ? nullptr : (T*)lcv->u[0];
        a.pitch_c = lcv->pitch;
        a.Nc = lcv->N;
    } else if (MODE == MODE_POST) {
        a.ec = (const T*)lcv->u[lcv->cur];
        a.pitch_c = lcv->pitch;
        a.Nc = lcv->N;
        a.crow_lo = lcv->st_lo;
        a.crow_hi = lcv->st_hi;
    } else if (MODE == MODE_POSTPRE) {
        // reads the coarse correction from the current coarse buffer and writes the next visit's zero guess into
        // the OTHER one (chunks run concurrently: the buffer being read must not be cleared under them)
        a.ec = (const T*)lcv->u[lcv->cur];
        a.fc = (T*)lcv->f;
        a.uc = write_zero_guess ? (T*)lcv->u[lcv->cur ^ 1] : nullptr;
        a.pitch_c = lcv->pitch;
        a.Nc = lcv->N;
        a.crow_lo = lcv->st_lo;
        a.crow_hi = lcv->st_hi;
    }
    return a;
}

// dynamic shared memory of a streaming launch: the ring, padded so that at most knobs.occ warps are resident per SM
template <typename C>
static size_t stream_smem(const Ctx& ctx)
{
    return std::min<size_t>(100 * 1024, std::max<size_t>(C::SMEM_BYTES, (size_t)(227 * 1024) / std::max(1, ctx.knobs.occ / kStreamWarps) - 1024));
}

// One kernel launch on the context's stream; with knobs.pdl as a programmatic dependent launch (common.cuh: pdl_wait),
// which also holds inside stream capture (the graph gets a programmatic edge).
template <typename... KP, typename... A>
static void launch_k(Ctx& ctx, void (*kernel)(KP...), unsigned grid, unsigned block, size_t smem, A&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx.stream;
    cudaLaunchAttribute at[1];
    if (ctx.knobs.pdl) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    MG_CK(cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...));
    ++ctx.lc.n;
}

template <typename T, int NS, int MODE, bool RBGS>
static void raw_launch(Ctx& ctx, const StreamArgs<T>& a)
{
    typedef StreamCfg<T, NS, MODE> C;
    if (a.yb <= a.ya) return;
    const unsigned grid = cdiv(a.nitems, kStreamWarps);
    const size_t smem = stream_smem<C>(ctx);
    if constexpr (MODE == MODE_POSTPRE) launch_k(ctx, k_stream_chain<T, NS, RBGS>, grid, kStreamWarps * 32, smem, a);
    else launch_k(ctx, k_stream<T, NS, MODE, RBGS>, grid, kStreamWarps * 32, smem, a);
}

// Chunk height per (level, kernel): on the big levels the best height depends on how strips x chunks maps
// onto waves and DRAM (measured spread at 4097^2: 79..117 us for the same kernel, profiles/r01_tune_stream.txt),
// so it is picked once by timing a few candidates.  Tuning launches write only scratch state (the non-current
// u buffer and, for PRE, the coarse f / zero guess that the real launch rewrites), results never depend on it.
constexpr int kTuneMinN = 2048;
constexpr int kTuneReps = 4;   // timed launches per candidate (after one warm-up)

template <typename T, int NS, int MODE, bool RBGS>
static int tuned_ry(Ctx& ctx, Level& lv, Level* lcv)
{
    const auto key = std::make_tuple(lv.level, MODE, NS, (int)RBGS);
    auto it = ctx.stream_ry.find(key);
    if (it != ctx.stream_ry.end()) return it->second;
    StreamArgs<T> a0 = make_args<T, NS, MODE>(ctx, lv, lcv, 0);
    if (lv.N < kTuneMinN || ctx.capturing || !ctx.knobs.autotune || ctx.knobs.force_ry > 0) return a0.ry;
    const int rows = lv.own_hi - lv.own_lo;
    std::vector<int> cand;
    auto add = [&](int ry) {
        ry = (std::max(8, std::min(512, ry)) + 1) & ~1;
        if (std::find(cand.begin(), cand.end(), ry) == cand.end()) cand.push_back(ry);
    };
    add(a0.ry);
    for (int w : {8, 10, 12, 16}) {
        const int chunks = std::max(1, ctx.knobs.num_sms * w / a0.strips);
        const int ry = (rows + chunks - 1) / chunks;
        if (ry <= 512) add(ry);
    }
    add(128); add(192); add(256);
    EventTimer timer(ctx.stream);
    float best = 1e30f;
    int best_ry = a0.ry;
    for (int ry : cand) {
        // no zero-guess store: lcv->u[0] may hold a coarse solution that fullmultigrid is about to interpolate (P:645)
        StreamArgs<T> a = make_args<T, NS, MODE>(ctx, lv, lcv, ry, -1, -1, false);
        raw_launch<T, NS, MODE, RBGS>(ctx, a);   // warm-up
        timer.start();
        for (int rep = 0; rep < kTuneReps; ++rep) raw_launch<T, NS, MODE, RBGS>(ctx, a);
        const float ms = timer.stop();
        if (ms < best) { best = ms; best_ry = ry; }
    }
    ctx.lc.n -= (1 + kTuneReps) * (long long)cand.size();   // tuning launches are not part of the work
    ctx.stream_ry[key] = best_ry;
    return best_ry;
}

template <typename T, int NS, int MODE, bool RBGS>
static void launch_stream(Ctx& ctx, Level& lv, Level* lcv)
{
    typedef StreamCfg<T, NS, MODE> C;
    ctx.materialize_u(lv);
    if (MODE == MODE_POST) ctx.materialize_u(*lcv);
    // lazy halo exchanges (no-ops on replicated levels): rows the stencil pipeline reaches into
    constexpr int du = C::HT - (MODE == MODE_POST ? 1 : 0);
    constexpr int df = NS + (MODE == MODE_PRE ? 1 : 0) - (MODE == MODE_SWEEPS ? 1 : 0);
    constexpr int dc = (MODE == MODE_POST) ? NS / 2 + 1 : 0;
    ctx.ensure_halo(lv, Ctx::W_U, du);
    ctx.ensure_halo(lv, Ctx::W_F, df);
    if (MODE == MODE_PRE) lcv->cur = 0;
    if (MODE == MODE_POST) ctx.ensure_halo(*lcv, Ctx::W_U, dc);
    const int ry = tuned_ry<T, NS, MODE, RBGS>(ctx, lv, lcv);
    const StreamArgs<T> a = make_args<T, NS, MODE>(ctx, lv, lcv, ry);
    raw_launch<T, NS, MODE, RBGS>(ctx, a);
    lv.cur ^= 1;
    lv.hv_u = 0;
    if (MODE == MODE_PRE) {
        lcv->u_zero = false;   // the zero guess is written (kernel stores or the memset below)
        if (lv.distributed && !lcv->distributed) {
            // agglomeration: gather the coarse right-hand side everywhere, zero guess on the full grid
            comm_allgather_rows(ctx, *lcv, lcv->f);
            MG_CK(cudaMemsetAsync(lcv->alloc[0], 0, lcv->bytes, ctx.stream));
        } else if (lcv->distributed) {
            lcv->hv_f = 0;
            comm_zero_halo(ctx, *lcv, lcv->u[0]);
            lcv->hv_u = lcv->halo;
        }
    }
}

static bool stream_ok(const Ctx& ctx, const Level&)
{
    return (ctx.cfg.flags & MG_FUSED) != 0;
}

template <typename T> static void stream_sweeps(Ctx& ctx, Level& lv, int nu);

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
// red-black Gauss-Seidel sweeps through the streaming kernel: both colours of a sweep (and two sweeps when nu >= 2) in
// one pass over the level, 3S bytes per point per launch instead of 6S per sweep with one k_rbgs launch per colour
bool fused_rbgs(Ctx& ctx, Level& lv, int nu)
{
    if (!stream_ok(ctx, lv) || ctx.cfg.smoother != MG_SMOOTH_RBGS || nu < 1) return false;
    if (ctx.f64()) stream_sweeps<double>(ctx, lv, nu);
    else stream_sweeps<float>(ctx, lv, nu);
    return true;
}

template int fused_jacobi<double>(Ctx&, Level&, int, double, double);
template int fused_jacobi<float>(Ctx&, Level&, int, float, float);

// ---------------------------------------------------------------------------------
// zero-guess chain (default; MGB200_ZERO_GUESS=0 turns it off): PRE(l) does not write the zero coarse guess when the first
// kernel of level l-1 is a zero-guess variant that does not read u either (the stream PRE or the tail).
// Saves S/4 bytes per fine point in PRE(l) and S per point in PRE(l-1).  Any other reader of a logically-zero
// iterate goes through Ctx::materialize_u, so a wrong prediction costs time, never correctness.
// ---------------------------------------------------------------------------------
static int tail_top(const Ctx& ctx);

static bool child_takes_zero_guess(Ctx& ctx, Level& lv, int nu1)
{
    if (!ctx.zero_guess || !(ctx.cfg.flags & MG_FUSED) || lv.distributed) return false;
    const int c = lv.level - 1;
    if (c < ctx.cfg.coarsest_level || ctx.L(c).distributed) return false;
    if (c == tail_top(ctx)) return true;
    if (c <= ctx.cfg.coarsest_level) return false;
    return nu1 >= 1 && nu1 <= 2;
}

template <typename T, int NS, bool RBGS>
static void launch_pre_zero_guess(Ctx& ctx, Level& lv, Level* lcv, bool write_zero_guess)
{
    // u == 0 needs no halo; f as for PRE
    ctx.ensure_halo(lv, Ctx::W_F, NS + 1);
    lcv->cur = 0;
    const int ry = tuned_ry<T, NS, MODE_PRE, RBGS>(ctx, lv, lcv);
    const StreamArgs<T> a = make_args<T, NS, MODE_PRE>(ctx, lv, lcv, ry, -1, -1, write_zero_guess);
    if (a.yb > a.ya) {
        typedef StreamCfg<T, NS, MODE_PRE> C;
        const size_t smem = stream_smem<C>(ctx);
        launch_k(ctx, k_stream_pre_zg<T, NS, RBGS>, cdiv(a.nitems, kStreamWarps), kStreamWarps * 32, smem, a);
    }
    lv.cur ^= 1;
    lv.hv_u = 0;
    lv.u_zero = false;
}

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

template <typename T> static bool fmg_entry_fused(Ctx& ctx, Level& lv, Level& lcv, int nu1);

template <typename T>
static void pre_fused(Ctx& ctx, Level& lv, Level& lcv, int nu1)
{
    const int k = std::min(nu1, 2);
    const bool rb = ctx.cfg.smoother == MG_SMOOTH_RBGS;
    if (lv.u_interp) {   // fullmultigrid entry (MGB200_CHAIN): interpolate the coarse solution inside this PRE
        if (fmg_entry_fused<T>(ctx, lv, lcv, nu1)) return;
        ctx.materialize_u(lv);
    }
    if (ctx.zero_guess && !lv.distributed) {
        // zero-guess chain
        const bool wz = !child_takes_zero_guess(ctx, lv, nu1);
        if (lv.u_zero && nu1 == k) {
            if (!rb) {
                if (k == 2) launch_pre_zero_guess<T, 2, false>(ctx, lv, &lcv, wz);
                else launch_pre_zero_guess<T, 1, false>(ctx, lv, &lcv, wz);
            } else {
                if (k == 2) launch_pre_zero_guess<T, 4, true>(ctx, lv, &lcv, wz);
                else launch_pre_zero_guess<T, 2, true>(ctx, lv, &lcv, wz);
            }
            lcv.u_zero = !wz;   // zeros were written unless the child skips reading them
            lcv.cur = 0;
            return;
        }
        if (!wz) {
            // regular PRE that skips the zero-guess store
            ctx.materialize_u(lv);
            stream_sweeps<T>(ctx, lv, nu1 - k);
            ctx.ensure_halo(lv, Ctx::W_U, (rb ? 2 * k : k) + 2);
            ctx.ensure_halo(lv, Ctx::W_F, (rb ? 2 * k : k) + 1);
            lcv.cur = 0;
            auto go = [&](auto ns_tag, auto rb_tag) {
                constexpr int NS = decltype(ns_tag)::value;
                constexpr bool RB = decltype(rb_tag)::value;
                const int ry = tuned_ry<T, NS, MODE_PRE, RB>(ctx, lv, &lcv);
                const StreamArgs<T> a = make_args<T, NS, MODE_PRE>(ctx, lv, &lcv, ry, -1, -1, false);
                raw_launch<T, NS, MODE_PRE, RB>(ctx, a);
            };
            if (!rb) {
                if (k == 2) go(std::integral_constant<int, 2>(), std::false_type());
                else go(std::integral_constant<int, 1>(), std::false_type());
            } else {
                if (k == 2) go(std::integral_constant<int, 4>(), std::true_type());
                else go(std::integral_constant<int, 2>(), std::true_type());
            }
            lv.cur ^= 1;
            lv.hv_u = 0;
            lcv.u_zero = true;
            return;
        }
    }
    stream_sweeps<T>(ctx, lv, nu1 - k);
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) {
        if (k == 2) launch_stream<T, 2, MODE_PRE, false>(ctx, lv, &lcv);
        else launch_stream<T, 1, MODE_PRE, false>(ctx, lv, &lcv);
    } else {
        if (k == 2) launch_stream<T, 4, MODE_PRE, true>(ctx, lv, &lcv);
        else launch_stream<T, 2, MODE_PRE, true>(ctx, lv, &lcv);
    }
}

// POST with the residual norm of its output folded in (k_stream_norm): the last kernel of a cycle on the finest level
// when the tolerance loop asked for the norm (Ctx::solve).  Leaves sum r^2 of this rank's rows in ctx.d_norm.
template <typename T, int NS, bool RBGS>
static bool launch_post_norm(Ctx& ctx, Level& lv, Level* lcv)
{
    typedef StreamCfg<T, NS, MODE_POST, true> C;
    const int ry = tuned_ry<T, NS, MODE_POST, RBGS>(ctx, lv, lcv);
    ctx.materialize_u(lv);
    ctx.materialize_u(*lcv);
    StreamArgs<T> a = make_args<T, NS, MODE_POST, true>(ctx, lv, lcv, ry);
    MG_REQUIRE(a.nitems <= ctx.partials_cap, "partials buffer too small for the norm kernel");   // (never rank-dependent)
    if (a.yb <= a.ya) return false;
    ctx.ensure_halo(lv, Ctx::W_U, C::HT - 1);
    ctx.ensure_halo(lv, Ctx::W_F, NS);
    ctx.ensure_halo(*lcv, Ctx::W_U, (NS + 2) / 2 + 1);
    launch_k(ctx, k_stream_norm<T, NS, MODE_POST, RBGS>, cdiv(a.nitems, kStreamWarps), kStreamWarps * 32, stream_smem<C>(ctx), a, ctx.d_partials);
    launch_k(ctx, k_sum_partials, 1u, 256u, (size_t)0, (const double*)ctx.d_partials, a.nitems, ctx.d_norm);
    lv.cur ^= 1;
    lv.hv_u = 0;
    return true;
}

template <typename T>
static void post_fused(Ctx& ctx, Level& lv, Level& lcv, int nu2)
{
    const int k = std::min(nu2, 2);
    if (ctx.want_post_norm && lv.level == ctx.cfg.finest_level && nu2 == k) {
        bool done;
        if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) done = (k == 2) ? launch_post_norm<T, 2, false>(ctx, lv, &lcv) : launch_post_norm<T, 1, false>(ctx, lv, &lcv);
        else done = (k == 2) ? launch_post_norm<T, 4, true>(ctx, lv, &lcv) : launch_post_norm<T, 2, true>(ctx, lv, &lcv);
        if (done) { ctx.post_norm_done = true; return; }
    }
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
// visit chains (default; MGB200_CHAIN=0 turns them off): when a level is visited several times in a row on the same right-hand
// side -- consecutive cycles on the top level (fullmultigrid runs mu0+1 per level, P:646-648; mg_cycles), the gamma
// visits of a W-cycle on every level below -- POST of visit v and PRE of visit v+1 are ONE streaming launch
// (stream.cuh MODE_POSTPRE): 3.5 S bytes per point instead of 6.5 S, the iterate between them never goes to memory.
// ---------------------------------------------------------------------------------
static int postpre_ns(const Ctx& ctx, const Level& lv, int nu1, int nu2)
{
    if (!ctx.chain || !(ctx.cfg.flags & MG_FUSED) || (lv.distributed && ctx.comm_avoid)) return 0;
    if (nu1 < 1 || nu2 < 1 || nu1 > 2 || nu2 > 2) return 0;
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) return nu1 + nu2;          // 2, 3 or 4 pipeline stages
    return (nu1 == 1 && nu2 == 1) ? 4 : 0;                               // RB-GS: one stage per colour
}

template <typename T, int NS, bool RBGS>
static void launch_postpre(Ctx& ctx, Level& lv, Level& lcv, bool write_zero_guess)
{
    ctx.materialize_u(lv);
    ctx.materialize_u(lcv);
    // row slabs: halo rows the pipeline reaches into (PRE's needs on u and f, POST's on the coarse correction)
    ctx.ensure_halo(lv, Ctx::W_U, NS + 2);
    ctx.ensure_halo(lv, Ctx::W_F, NS + 1);
    ctx.ensure_halo(lcv, Ctx::W_U, (NS + 2) / 2 + 1);
    const bool gather = lv.distributed && !lcv.distributed;   // agglomeration boundary: zero guess by memset after the all-gather
    const int ry = tuned_ry<T, NS, MODE_POSTPRE, RBGS>(ctx, lv, &lcv);
    const StreamArgs<T> a = make_args<T, NS, MODE_POSTPRE>(ctx, lv, &lcv, ry, -1, -1, write_zero_guess && !gather);
    raw_launch<T, NS, MODE_POSTPRE, RBGS>(ctx, a);
    lv.cur ^= 1;
    lv.hv_u = 0;
    if (write_zero_guess) {
        lcv.cur ^= 1;            // the zero guess went into the other coarse buffer
        lcv.u_zero = false;
    } else {
        lcv.u_zero = true;       // zero-guess chain: the child's first kernel does not read its iterate
    }
    if (gather) {
        comm_allgather_rows(ctx, lcv, lcv.f);
        if (write_zero_guess) MG_CK(cudaMemsetAsync(lcv.alloc[lcv.cur], 0, lcv.bytes, ctx.stream));
    } else if (lcv.distributed) {
        lcv.hv_f = 0;
        if (write_zero_guess) {
            comm_zero_halo(ctx, lcv, lcv.u[lcv.cur]);
            lcv.hv_u = lcv.halo;
        }
    }
}

template <typename T>
static void postpre_fused(Ctx& ctx, Level& lv, Level& lcv, int ns, int nu1)
{
    const bool wz = !child_takes_zero_guess(ctx, lv, nu1);
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) {
        if (ns == 2) launch_postpre<T, 2, false>(ctx, lv, lcv, wz);
        else if (ns == 3) launch_postpre<T, 3, false>(ctx, lv, lcv, wz);
        else launch_postpre<T, 4, false>(ctx, lv, lcv, wz);
    } else {
        launch_postpre<T, 4, true>(ctx, lv, lcv, wz);
    }
}

// fullmultigrid's entry into a level: interpolation of the coarse solution (P:645) + PRE of the first cycle in one launch
template <typename T, int NS, bool RBGS>
static void launch_fmg_entry(Ctx& ctx, Level& lv, Level& lcv, bool write_zero_guess)
{
    typedef StreamCfg<T, NS, MODE_POSTPRE> C;
    ctx.materialize_u(lcv);
    lv.u_interp = false;                       // consumed here: stage 0 of the kernel is P u_c
    StreamArgs<T> a = make_args<T, NS, MODE_POSTPRE>(ctx, lv, &lcv, 0, -1, -1, write_zero_guess);
    if (a.yb > a.ya) {
        const size_t smem = stream_smem<C>(ctx);
        launch_k(ctx, k_stream_fmg_entry<T, NS, RBGS>, cdiv(a.nitems, kStreamWarps), kStreamWarps * 32, smem, a);
    }
    lv.cur ^= 1;
    lv.hv_u = 0;
    if (write_zero_guess) {
        lcv.cur ^= 1;
        lcv.u_zero = false;
    } else {
        lcv.u_zero = true;
    }
}

template <typename T>
static bool fmg_entry_fused(Ctx& ctx, Level& lv, Level& lcv, int nu1)
{
    if (!ctx.chain || lv.distributed || nu1 < 1 || nu1 > 2) return false;
    const bool wz = !child_takes_zero_guess(ctx, lv, nu1);
    if (ctx.cfg.smoother == MG_SMOOTH_JACOBI) {
        if (nu1 == 2) launch_fmg_entry<T, 2, false>(ctx, lv, lcv, wz);
        else launch_fmg_entry<T, 1, false>(ctx, lv, lcv, wz);
    } else {
        if (nu1 == 2) launch_fmg_entry<T, 4, true>(ctx, lv, lcv, wz);
        else launch_fmg_entry<T, 2, true>(ctx, lv, lcv, wz);
    }
    return true;
}

// ---------------------------------------------------------------------------------
// coarse tail: levels <= tail_top run in ONE launch, entirely in one CTA's shared memory
// ---------------------------------------------------------------------------------
static int tail_top(const Ctx& ctx)
{
    if (!(ctx.cfg.flags & MG_COARSE_TAIL)) return -1;
    if (ctx.exact_coarse()) return -1;   // the coarsest level is a direct solve (coarse.cuh): every level runs its own kernels
    int top = std::min(kTailMaxLevel, ctx.cfg.finest_level);
    if (top < ctx.cfg.coarsest_level) return -1;
    if (ctx.L(top).distributed) return -1;
    return top;
}

template <typename T>
static void run_tail(Ctx& ctx, int level, int nu1, int nu2, int gamma)
{
    Level& lv = ctx.L(level);
    if (lv.u_interp) ctx.materialize_u(lv);
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
    if (lv.u_zero) {   // zero-guess chain: u is logically zero and is not read
        if (ctx.cfg.smoother == MG_SMOOTH_RBGS) launch_k(ctx, k_tail_zg<T, true>, 1u, (unsigned)kTailThreads, smem, a);
        else launch_k(ctx, k_tail_zg<T, false>, 1u, (unsigned)kTailThreads, smem, a);
        lv.u_zero = false;
        lv.hv_u = lv.halo;
    } else if (ctx.cfg.smoother == MG_SMOOTH_RBGS) launch_k(ctx, k_tail<T, true>, 1u, (unsigned)kTailThreads, smem, a);
    else launch_k(ctx, k_tail<T, false>, 1u, (unsigned)kTailThreads, smem, a);
}

// ---------------------------------------------------------------------------------
// communication-avoiding V-cycle over the distributed levels (sched.h), the default at world > 1 (MGB200_COMM_AVOID=0: lazy exchanges).
// The op list comes from the same planner the CPU emulation test executes with the oracle.
// ---------------------------------------------------------------------------------
template <typename T, int NS, int MODE, bool RBGS>
static void launch_stream_range(Ctx& ctx, Level& lv, Level* lcv, int ya, int yb)
{
    if (MODE == MODE_PRE) lcv->cur = 0;
    const int ry = tuned_ry<T, NS, MODE, RBGS>(ctx, lv, lcv);
    const StreamArgs<T> a = make_args<T, NS, MODE>(ctx, lv, lcv, ry, ya, yb);
    raw_launch<T, NS, MODE, RBGS>(ctx, a);
    lv.cur ^= 1;
}

template <typename T>
static void comm_avoid_op(Ctx& ctx, const SchedPlan& plan, const SchedOp& op, int nu1, int nu2)
{
    const bool rb = ctx.cfg.smoother == MG_SMOOTH_RBGS;
    Level& lv = ctx.L(op.level);
    switch (op.kind) {
        case SCHED_EXCH: {
            comm_halo_exchange(ctx, lv, op.a == 0 ? lv.u[lv.cur] : lv.f, op.b);
            if (op.a == 0) lv.hv_u = op.b; else lv.hv_f = op.b;
            break;
        }
        case SCHED_PRE: {
            Level& lcv = ctx.L(op.level - 1);
            if (lcv.distributed) comm_zero_halo(ctx, lcv, lcv.u[0]);   // rows of the zero guess the kernel does not reach
            if (!rb) {
                if (nu1 == 2) launch_stream_range<T, 2, MODE_PRE, false>(ctx, lv, &lcv, op.a, op.b);
                else launch_stream_range<T, 1, MODE_PRE, false>(ctx, lv, &lcv, op.a, op.b);
            } else {
                if (nu1 == 2) launch_stream_range<T, 4, MODE_PRE, true>(ctx, lv, &lcv, op.a, op.b);
                else launch_stream_range<T, 2, MODE_PRE, true>(ctx, lv, &lcv, op.a, op.b);
            }
            lv.hv_u = plan.e[op.level];
            lcv.u_zero = false;
            if (lcv.distributed) {
                lcv.hv_f = plan.e[op.level] / 2;
                lcv.hv_u = lcv.halo;
            }
            break;
        }
        case SCHED_GATHER_F: {
            comm_allgather_rows(ctx, lv, lv.f);
            lv.cur = 0;
            MG_CK(cudaMemsetAsync(lv.alloc[0], 0, lv.bytes, ctx.stream));
            break;
        }
        case SCHED_REPL_CYCLE: ctx.cycle_rec(op.level, nu1, nu2, 1); break;
        case SCHED_POST: {
            Level& lcv = ctx.L(op.level - 1);
            if (!rb) {
                if (nu2 == 2) launch_stream_range<T, 2, MODE_POST, false>(ctx, lv, &lcv, op.a, op.b);
                else launch_stream_range<T, 1, MODE_POST, false>(ctx, lv, &lcv, op.a, op.b);
            } else {
                if (nu2 == 2) launch_stream_range<T, 4, MODE_POST, true>(ctx, lv, &lcv, op.a, op.b);
                else launch_stream_range<T, 2, MODE_POST, true>(ctx, lv, &lcv, op.a, op.b);
            }
            lv.hv_u = plan.x[op.level];
            break;
        }
    }
}

template <typename T>
static void comm_avoid_run(Ctx& ctx, const SchedPlan& plan, int nu1, int nu2)
{
    ctx.materialize_u(ctx.L(plan.ops.front().level));
    for (const SchedOp& op : plan.ops) {
        // mg_time_phases: a pair of events around every op of the plan (eager launches only; see Ctx::time_phases)
        const bool timed = ctx.phase_on && !ctx.capturing;
        size_t slot = 0;
        if (timed) {
            // the record is in the log before its events exist: whatever throws below, Ctx::time_phases destroys them
            slot = ctx.phase_log.size();
            ctx.phase_log.emplace_back();
            PhaseRec& rec = ctx.phase_log.back();
            rec.kind = op.kind;
            rec.level = op.level;
            MG_CK(cudaEventCreate(&rec.e0));
            MG_CK(cudaEventCreate(&rec.e1));
            MG_CK(cudaEventRecord(rec.e0, ctx.stream));
        }
        comm_avoid_op<T>(ctx, plan, op, nu1, nu2);
        if (timed) MG_CK(cudaEventRecord(ctx.phase_log[slot].e1, ctx.stream));
    }
}

static bool comm_avoid_cycle(Ctx& ctx, int level, int nu1, int nu2, int gamma)
{
    if (!ctx.comm_avoid || ctx.cfg.world < 2 || gamma != 1 || !(ctx.cfg.flags & MG_FUSED)) return false;
    if (nu1 < 1 || nu1 > 2 || nu2 < 1 || nu2 > 2) return false;
    Level& lv = ctx.L(level);
    if (!lv.distributed || ctx.aggl_level < ctx.cfg.coarsest_level) return false;
    const bool rb = ctx.cfg.smoother == MG_SMOOTH_RBGS;
    const int ns1 = rb ? 2 * nu1 : nu1, ns2 = rb ? 2 * nu2 : nu2;
    const SchedPlan plan = sched_plan_vcycle(level, ctx.aggl_level, ctx.cfg.world, ctx.cfg.rank, ns1, ns2, lv.hv_u, lv.hv_f);
    if (!plan.ok) return false;
    for (int l = ctx.aggl_level + 1; l <= level; ++l)
        if (plan.halo[l] > ctx.L(l).halo) return false;
    if (ctx.f64()) comm_avoid_run<double>(ctx, plan, nu1, nu2);
    else comm_avoid_run<float>(ctx, plan, nu1, nu2);
    return true;
}

bool fused_cycle_level(Ctx& ctx, int level, int nu1, int nu2, int gamma)
{
    if (comm_avoid_cycle(ctx, level, nu1, nu2, gamma)) return true;
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
    ctx.cycle_rec_visits(level - 1, nu1, nu2, gamma, reps);
    if (ctx.f64()) post_fused<double>(ctx, lv, lcv, nu2);
    else post_fused<float>(ctx, lv, lcv, nu2);
    return true;
}

bool fused_cycle_chain(Ctx& ctx, int level, int nu1, int nu2, int gamma, int visits)
{
    if (visits < 2 || level <= ctx.cfg.coarsest_level) return false;
    if (level == tail_top(ctx)) return false;   // one launch per visit already
    Level& lv = ctx.L(level);
    Level& lcv = ctx.L(level - 1);
    const int ns = postpre_ns(ctx, lv, nu1, nu2);
    if (ns == 0) return false;
    if (ctx.f64()) pre_fused<double>(ctx, lv, lcv, nu1);
    else pre_fused<float>(ctx, lv, lcv, nu1);
    const int reps = (level - 1 <= ctx.cfg.coarsest_level) ? 1 : std::max(1, gamma);
    for (int v = 0; v < visits; ++v) {
        ctx.cycle_rec_visits(level - 1, nu1, nu2, gamma, reps);
        if (v + 1 < visits) {
            if (ctx.f64()) postpre_fused<double>(ctx, lv, lcv, ns, nu1);
            else postpre_fused<float>(ctx, lv, lcv, ns, nu1);
        } else {
            if (ctx.f64()) post_fused<double>(ctx, lv, lcv, nu2);
            else post_fused<float>(ctx, lv, lcv, nu2);
        }
    }
    return true;
}

// Pick the chunk heights of every big level a cycle from `level` will visit, eagerly (cannot be done while
// a CUDA graph is being captured).  Side-effect free for the solver state.
template <typename T>
static void pretune_t(Ctx& ctx, int level, int nu1, int nu2)
{
    for (int l = level; l > ctx.cfg.coarsest_level; --l) {
        Level& lv = ctx.L(l);
        Level& lcv = ctx.L(l - 1);
        if (lv.N < kTuneMinN) break;
        const bool rb = ctx.cfg.smoother == MG_SMOOTH_RBGS;
        const int k1 = std::min(nu1, 2), k2 = std::min(nu2, 2);
        if (!rb) {
            if (k1 == 2) tuned_ry<T, 2, MODE_PRE, false>(ctx, lv, &lcv); else if (k1 == 1) tuned_ry<T, 1, MODE_PRE, false>(ctx, lv, &lcv);
            if (k2 == 2) tuned_ry<T, 2, MODE_POST, false>(ctx, lv, &lcv); else if (k2 == 1) tuned_ry<T, 1, MODE_POST, false>(ctx, lv, &lcv);
            if (nu1 - k1 >= 3 || nu2 - k2 >= 3) tuned_ry<T, 3, MODE_SWEEPS, false>(ctx, lv, nullptr);
            switch (postpre_ns(ctx, lv, nu1, nu2)) {
                case 2: tuned_ry<T, 2, MODE_POSTPRE, false>(ctx, lv, &lcv); break;
                case 3: tuned_ry<T, 3, MODE_POSTPRE, false>(ctx, lv, &lcv); break;
                case 4: tuned_ry<T, 4, MODE_POSTPRE, false>(ctx, lv, &lcv); break;
                default: break;
            }
        } else {
            if (postpre_ns(ctx, lv, nu1, nu2) == 4) tuned_ry<T, 4, MODE_POSTPRE, true>(ctx, lv, &lcv);
            if (k1 == 2) tuned_ry<T, 4, MODE_PRE, true>(ctx, lv, &lcv); else if (k1 == 1) tuned_ry<T, 2, MODE_PRE, true>(ctx, lv, &lcv);
            if (k2 == 2) tuned_ry<T, 4, MODE_POST, true>(ctx, lv, &lcv); else if (k2 == 1) tuned_ry<T, 2, MODE_POST, true>(ctx, lv, &lcv);
            if (nu1 - k1 >= 2 || nu2 - k2 >= 2) tuned_ry<T, 4, MODE_SWEEPS, true>(ctx, lv, nullptr);
        }
    }
}

void fused_pretune(Ctx& ctx, int level, int nu1, int nu2, int gamma)
{
    if (!(ctx.cfg.flags & MG_FUSED) || nu1 < 1 || nu2 < 1) return;
    if (ctx.f64()) pretune_t<double>(ctx, level, nu1, nu2);
    else pretune_t<float>(ctx, level, nu1, nu2);
}

bool fused_time_sweeps4(Ctx& ctx, int level)
{
    Level& lv = ctx.L(level);
    if (!stream_ok(ctx, lv) || ctx.cfg.smoother != MG_SMOOTH_JACOBI) return false;
    if (ctx.f64()) launch_stream<double, 4, MODE_SWEEPS, false>(ctx, lv, nullptr);
    else launch_stream<float, 4, MODE_SWEEPS, false>(ctx, lv, nullptr);
    return true;
}

bool fused_time_hook_postpre(Ctx& ctx, int level)
{
    if (level <= ctx.cfg.coarsest_level) return false;
    Level& lv = ctx.L(level);
    Level& lcv = ctx.L(level - 1);
    const int nu = ctx.cfg.smoother == MG_SMOOTH_JACOBI ? 2 : 1;
    const int ns = postpre_ns(ctx, lv, nu, nu);
    if (ns == 0) return false;
    if (ctx.f64()) postpre_fused<double>(ctx, lv, lcv, ns, nu);
    else postpre_fused<float>(ctx, lv, lcv, ns, nu);
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
