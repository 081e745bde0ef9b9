// ctx.cu — context, level table, host<->device movement, cycles.  See ctx.cuh.
#include "ctx.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <thread>

#include "coarse.cuh"
#include "comm.cuh"
#include "fused.cuh"
#include "sched.h"

namespace mgb {

// ---------------------------------------------------------------------------------
// slab partition (shared with mg_slab_rows): rank r of R owns node rows
// [r*N/R, (r+1)*N/R); interior rows are 1..N-1.  Slab boundaries are multiples of
// N/R on every distributed level, so coarse row I (fine row 2I) and fine row 2I have the
// same owner and a coarse row never straddles ranks.
// ---------------------------------------------------------------------------------
void slab_rows(int level, int rank, int world, int* lo, int* hi)
{
    const i64 N = (i64)1 << level;
    i64 a = (i64)rank * N / world;
    i64 b = (i64)(rank + 1) * N / world;
    if (a < 1) a = 1;
    if (b > N) b = N;
    if (rank == world - 1) b = N;
    *lo = (int)a;
    *hi = (int)b;
}

static i64 round_up(i64 a, i64 m) { return (a + m - 1) / m * m; }

// A constructor that throws does not run the destructor: everything acquired so far (stream, device arrays of the
// levels already allocated, pinned buffers, communicator) is released here before the error leaves mg_create, so a
// caller that retries with a smaller grid starts from a clean device.
Ctx::Ctx(const mg_config& c) : cfg(c)
{
    try {
        init();
    } catch (...) {
        release();
        throw;
    }
}

void Ctx::init()
{
    MG_REQUIRE(cfg.finest_level >= 1 && cfg.finest_level <= 15, "finest_level must be in 1..15");
    MG_REQUIRE(cfg.coarsest_level >= 1 && cfg.coarsest_level <= cfg.finest_level,
               "coarsest_level must be in 1..finest_level");
    MG_REQUIRE(cfg.dtype == MG_F64 || cfg.dtype == MG_F32, "dtype must be MG_F64 or MG_F32");
    MG_REQUIRE(cfg.smoother == MG_SMOOTH_JACOBI || cfg.smoother == MG_SMOOTH_RBGS, "unknown smoother");
    MG_REQUIRE(cfg.coarse_solver == MG_COARSE_SWEEPS || cfg.coarse_solver == MG_COARSE_EXACT, "unknown coarse_solver");
    MG_REQUIRE(!exact_coarse() || cfg.coarsest_level <= kCoarseExactMaxLevel,
               "MG_COARSE_EXACT needs coarsest_level <= 9 (the sine-transform table is n x n)");
    MG_REQUIRE(cfg.world >= 1 && cfg.rank >= 0 && cfg.rank < cfg.world, "bad rank/world");
    MG_REQUIRE((cfg.world & (cfg.world - 1)) == 0, "world must be a power of two");
    esize = f64() ? 8 : 4;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw MgError(MG_ERR_CUDA, std::string("no CUDA device: libmgb200 has no CPU fallback (") +
                                       cudaGetErrorString(e) + ")");
    if (cfg.device >= 0) {
        MG_REQUIRE(cfg.device < ndev, "device ordinal out of range");
        MG_CK(cudaSetDevice(cfg.device));
    }
    MG_CK(cudaGetDevice(&device));
    MG_CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

    // agglomeration threshold: distributed levels need >= 8 rows per rank
    if (cfg.world > 1) {
        int min_dist = 3;
        while (((i64)1 << min_dist) / cfg.world < 8) ++min_dist;
        // default 10: measured best at 8 GPUs on 16385^2 (profiles/r01_scaling_16385.md: 0.764 ms vs 0.779 at 11, 0.998 at 12)
        aggl_level = cfg.agglomerate_level > 0 ? cfg.agglomerate_level : std::min(cfg.finest_level - 1, 10);
        aggl_level = std::max(aggl_level, min_dist - 1);
        aggl_level = std::max(aggl_level, cfg.coarsest_level - 1);
        MG_REQUIRE(aggl_level < cfg.finest_level || cfg.finest_level < min_dist,
                   "agglomerate_level must be below finest_level");
        if (aggl_level >= cfg.finest_level) aggl_level = cfg.finest_level;  // tiny problems: fully replicated
    } else {
        aggl_level = cfg.finest_level;
    }

    levels.resize(cfg.finest_level + 1);
    // stored halo rows per level: kHaloRows covers every fused kernel; the communication-avoiding schedule
    // (the default) computes rows beyond the slab and needs deeper halos on the fine levels (sched.h)
    int halo_rows[32];
    for (int l = 0; l < 32; ++l) halo_rows[l] = kHaloRows;
    if (cfg.world > 1) {
        // default (measured at 2 and 8 GPUs, profiles/r02_scaling.md); MGB200_COMM_AVOID=0 selects the lazy exchanges
        const char* e = getenv("MGB200_COMM_AVOID");
        comm_avoid = !(e && e[0] == '0');
        if (comm_avoid) {
            const int ns = (cfg.smoother == MG_SMOOTH_RBGS) ? 4 : 2;
            int x[32], ex[32], need[32];
            bool ok = sched_extents(cfg.finest_level, aggl_level, ns, ns, x, ex, need);
            for (int l = aggl_level + 1; ok && l <= cfg.finest_level; ++l)
                if (need[l] > ((i64)1 << l) / cfg.world - 1) ok = false;
            if (ok)
                for (int l = aggl_level + 1; l <= cfg.finest_level; ++l) halo_rows[l] = std::max(kHaloRows, need[l]);
            else
                comm_avoid = false;
        }
    }
    for (int l = cfg.coarsest_level; l <= cfg.finest_level; ++l) {
        Level& lv = levels[l];
        const int halo = halo_rows[l];
        lv.level = l;
        lv.N = 1 << l;
        lv.pitch = round_up((i64)lv.N + 1, 32);
        lv.distributed = (cfg.world > 1 && l > aggl_level);
        lv.halo = lv.distributed ? halo : 0;
        if (lv.distributed) {
            slab_rows(l, cfg.rank, cfg.world, &lv.own_lo, &lv.own_hi);
            lv.st_lo = std::max(0, lv.own_lo - halo);
            lv.st_hi = std::min(lv.N + 1, lv.own_hi + halo);
            if (cfg.rank == 0) lv.st_lo = 0;
        } else {
            lv.own_lo = 1;
            lv.own_hi = lv.N;
            lv.st_lo = 0;
            lv.st_hi = lv.N + 1;
        }
        lv.bytes = (size_t)(lv.st_hi - lv.st_lo) * (size_t)lv.pitch * (size_t)esize;
        for (int k = 0; k < 4; ++k) {
            cudaError_t ae = cudaMalloc(&lv.alloc[k], lv.bytes);
            if (ae != cudaSuccess)
                throw MgError(MG_ERR_ALLOC, std::string("cudaMalloc of level ") + std::to_string(l) + " failed: " +
                                                cudaGetErrorString(ae));
            MG_CK(cudaMemsetAsync(lv.alloc[k], 0, lv.bytes, stream));
            bytes_allocated += lv.bytes;
        }
        const i64 off = (i64)lv.st_lo * lv.pitch * esize;
        lv.u[0] = (char*)lv.alloc[0] - off;
        lv.u[1] = (char*)lv.alloc[1] - off;
        lv.f = (char*)lv.alloc[2] - off;
        lv.r = (char*)lv.alloc[3] - off;
    }
    {
        // One partial per residual block (k_residual) or per streaming work item (k_stream_norm: at most one item per 8 rows
        // and per strip of >= 48 columns).  The capacity is the maximum over all levels and must NOT depend on the rank: a
        // rank-dependent capacity once made edge ranks fall back to another kernel -- with other halo depths -- than their
        // neighbours.  Distributed levels therefore count N / world + 2 * halo rows whatever this rank stores.
        const int V = f64() ? 2 : 4;
        partials_cap = 0;
        for (int l = cfg.coarsest_level; l <= cfg.finest_level; ++l) {
            const Level& lv = levels[l];
            const i64 rows = lv.distributed ? (i64)lv.N / cfg.world + 2 * lv.halo + 2 : (i64)lv.N + 1;
            const i64 resid = (i64)cdiv(lv.N, V * kTX) * cdiv(rows, 2);
            const i64 items = (i64)cdiv(lv.N, 48) * (cdiv(rows, 8) + 1);
            partials_cap = std::max(partials_cap, (int)std::max(resid, items) + 8);
        }
        MG_CK(cudaMalloc(&d_partials, sizeof(double) * (size_t)partials_cap));
        MG_CK(cudaMalloc(&d_norm, sizeof(double) * 8));
        MG_CK(cudaMallocHost(&h_norm, sizeof(double) * 8));
    }
    if (exact_coarse()) {
        // sine-transform tables of the coarsest level, formed in double with the oracle's expressions, rounded to T once
        MG_REQUIRE(!levels[cfg.coarsest_level].distributed, "MG_COARSE_EXACT needs a replicated coarsest level");
        const int n = (1 << cfg.coarsest_level) - 1;
        const double pi = 3.14159265358979323846;
        std::vector<double> S((size_t)n * n), d(n);
        for (int j = 0; j < n; ++j) {
            d[j] = 2.0 - 2.0 * std::cos(pi * (double)(j + 1) / (double)(n + 1));
            for (int k = 0; k < n; ++k) {
                const long p = ((long)(j + 1) * (long)(k + 1)) % (2L * (n + 1));   // exact period of the argument
                S[(size_t)j * n + k] = std::sin(pi * (double)p / (double)(n + 1));
            }
        }
        MG_CK(cudaMalloc(&d_dst_S, S.size() * esize));
        MG_CK(cudaMalloc(&d_dst_d, d.size() * esize));
        if (f64()) {
            MG_CK(cudaMemcpy(d_dst_S, S.data(), S.size() * 8, cudaMemcpyHostToDevice));
            MG_CK(cudaMemcpy(d_dst_d, d.data(), d.size() * 8, cudaMemcpyHostToDevice));
        } else {
            std::vector<float> Sf(S.begin(), S.end()), df(d.begin(), d.end());
            MG_CK(cudaMemcpy(d_dst_S, Sf.data(), Sf.size() * 4, cudaMemcpyHostToDevice));
            MG_CK(cudaMemcpy(d_dst_d, df.data(), df.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    if (cfg.world > 1) {
        comm = comm_create(*this);
        // distributed cycles are captured (NCCL calls included) and replayed as CUDA graphs unless MGB200_GRAPH_DIST=0
        const char* e = getenv("MGB200_GRAPH_DIST");
        graph_dist = !(e && e[0] == '0');
    }
    {
        // both on by default (measured on the B200: profiles/r02a_*); "=0" switches one off for A/B timing
        const char* zg = getenv("MGB200_ZERO_GUESS");
        zero_guess = !(zg && zg[0] == '0');
        const char* ch = getenv("MGB200_CHAIN");
        chain = !(ch && ch[0] == '0');
    }
    fused_setup(*this);
    MG_CK(cudaStreamSynchronize(stream));
}

Ctx::~Ctx() { release(); }

void Ctx::release() noexcept
{
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& kv : graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (comm) comm_destroy(comm);
    for (auto& lv : levels)
        for (int k = 0; k < 4; ++k)
            if (lv.alloc[k]) cudaFree(lv.alloc[k]);
    {   // (every handle is checked: a failed stager_init leaves some of them null)
        for (int t = 0; t < Stager::kThreads; ++t) {
            if (stager.st[t]) { cudaStreamSynchronize(stager.st[t]); cudaStreamDestroy(stager.st[t]); }
            if (stager.ev_done[t]) cudaEventDestroy(stager.ev_done[t]);
            for (int b = 0; b < Stager::kBufs; ++b) {
                if (stager.pinned[t][b]) cudaFreeHost(stager.pinned[t][b]);
                if (stager.ev[t][b]) cudaEventDestroy(stager.ev[t][b]);
            }
        }
        if (stager.ev_main) cudaEventDestroy(stager.ev_main);
        stager = Stager();
    }
    for (auto& kv : solve_graphs)
        if (kv.second.first) cudaGraphExecDestroy(kv.second.first);
    solve_graphs.clear();
    if (d_solve_ctl) cudaFree(d_solve_ctl);
    d_solve_ctl = nullptr;
    solve_ctl_cap = 0;
    if (body_stream) cudaStreamDestroy(body_stream);
    body_stream = nullptr;
    if (d_partials) cudaFree(d_partials);
    if (d_dst_S) cudaFree(d_dst_S);
    if (d_dst_d) cudaFree(d_dst_d);
    d_dst_S = d_dst_d = nullptr;
    if (d_norm) cudaFree(d_norm);
    if (h_norm) cudaFreeHost(h_norm);
    if (stream) cudaStreamDestroy(stream);
    graphs.clear();
    levels.clear();
    comm = nullptr;
    d_partials = d_norm = h_norm = nullptr;
    stream = nullptr;
}

Level& Ctx::L(int level)
{
    if (level < cfg.coarsest_level || level > cfg.finest_level)
        throw MgError(MG_ERR_ARG, "level " + std::to_string(level) + " outside [coarsest_level, finest_level]");
    return levels[level];
}
const Level& Ctx::L(int level) const { return const_cast<Ctx*>(this)->L(level); }

void Ctx::sync() { MG_CK(cudaStreamSynchronize(stream)); }

// write the zeros of a logically-zero iterate (no-op unless MGB200_ZERO_GUESS left one pending)
void Ctx::materialize_u(Level& lv)
{
    if (lv.u_interp) {           // pending bare interpolation (fullmultigrid entry, MGB200_CHAIN)
        lv.u_interp = false;
        prolong(lv.level, false);
        return;
    }
    if (!lv.u_zero) return;
    MG_CK(cudaMemsetAsync(lv.alloc[lv.cur], 0, lv.bytes, stream));
    lv.u_zero = false;
    lv.hv_u = lv.halo;
}

static int& halo_ref(Level& lv, Ctx::Which w) { return w == Ctx::W_U ? lv.hv_u : (w == Ctx::W_F ? lv.hv_f : lv.hv_r); }

void Ctx::set_halo(Level& lv, Which w, int depth) { halo_ref(lv, w) = depth; }

// lazy halo exchange: make `depth` halo rows of the array valid (no-op on replicated levels)
void Ctx::ensure_halo(Level& lv, Which w, int depth)
{
    if (!lv.distributed || depth <= 0) return;
    MG_REQUIRE(depth <= lv.halo, "halo depth exceeds the stored halo");
    int& hv = halo_ref(lv, w);
    if (hv >= depth) return;
    char* base = (w == W_U) ? lv.u[lv.cur] : (w == W_F ? lv.f : lv.r);
    comm_halo_exchange(*this, lv, base, depth);
    hv = depth;
}

std::string Ctx::state_blob() const
{
    std::string b;
    for (int l = cfg.coarsest_level; l <= cfg.finest_level; ++l) {
        const Level& lv = levels[l];
        b.push_back((char)(lv.cur | (lv.u_zero ? 2 : 0) | (lv.u_interp ? 4 : 0)));
        if (lv.distributed)   // halo depths exceed 127 under the communication-avoiding schedule: 16 bits each
            for (int hv : {lv.hv_u, lv.hv_f, lv.hv_r}) {
                b.push_back((char)(hv & 0xff));
                b.push_back((char)((hv >> 8) & 0xff));
            }
    }
    return b;
}

void Ctx::set_state(const std::string& blob)
{
    size_t k = 0;
    for (int l = cfg.coarsest_level; l <= cfg.finest_level; ++l) {
        Level& lv = levels[l];
        lv.cur = blob[k] & 1;
        lv.u_zero = (blob[k] & 2) != 0;
        lv.u_interp = (blob[k] & 4) != 0;
        ++k;
        if (lv.distributed) {
            auto rd = [&]() { const int v = (unsigned char)blob[k] | ((unsigned char)blob[k + 1] << 8); k += 2; return v; };
            lv.hv_u = rd();
            lv.hv_f = rd();
            lv.hv_r = rd();
        }
    }
}

// ---------------------------------------------------------------------------------
// host <-> device.  Host vectors are full-grid interior vectors in the reference layout
// (n x n row-major, P:227-228); a rank moves every interior row it stores (owned + halo),
// so halos are valid after a set without an exchange.
// ---------------------------------------------------------------------------------
static char* which_ptr(Level& lv, Ctx::Which w)
{
    switch (w) {
        case Ctx::W_U: return lv.u[lv.cur];
        case Ctx::W_F: return lv.f;
        default: return lv.r;
    }
}

void Ctx::stager_init()
{
    if (stager.ready) return;
    for (int t = 0; t < Stager::kThreads; ++t) {
        MG_CK(cudaStreamCreateWithFlags(&stager.st[t], cudaStreamNonBlocking));
        MG_CK(cudaEventCreateWithFlags(&stager.ev_done[t], cudaEventDisableTiming));
        for (int b = 0; b < Stager::kBufs; ++b) {
            MG_CK(cudaMallocHost((void**)&stager.pinned[t][b], Stager::kChunkBytes));
            MG_CK(cudaEventCreateWithFlags(&stager.ev[t][b], cudaEventDisableTiming));
        }
    }
    MG_CK(cudaEventCreateWithFlags(&stager.ev_main, cudaEventDisableTiming));
    stager.ready = true;
}

void Ctx::copy_rows_staged(char* dev, size_t dpitch, char* host, size_t row_bytes, int rows, bool to_device)
{
    stager_init();
    const int rows_per_chunk = (int)std::max<size_t>(1, Stager::kChunkBytes / row_bytes);
    const int nchunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
    // the worker streams start after everything already queued on the context's stream (which may still use the array)
    MG_CK(cudaEventRecord(stager.ev_main, stream));
    cudaError_t errs[Stager::kThreads];
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        cudaStream_t st = stager.st[t];
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, stager.ev_main, 0);
        int k = 0;
        int pend_c = -1, pend_b = -1;   // device -> host: the chunk whose copy into a pinned buffer is in flight
        auto drain = [&]() {
            if (pend_c < 0 || e != cudaSuccess) return;
            e = cudaEventSynchronize(stager.ev[t][pend_b]);
            const int r0 = pend_c * rows_per_chunk, nr = std::min(rows_per_chunk, rows - r0);
            if (e == cudaSuccess) std::memcpy(host + (size_t)r0 * row_bytes, stager.pinned[t][pend_b], (size_t)nr * row_bytes);
            pend_c = -1;
        };
        for (int c = t; c < nchunks && e == cudaSuccess; c += Stager::kThreads, ++k) {
            const int b = k % Stager::kBufs;
            const int r0 = c * rows_per_chunk, nr = std::min(rows_per_chunk, rows - r0);
            char* pin = stager.pinned[t][b];
            if (to_device) {
                e = cudaEventSynchronize(stager.ev[t][b]);      // the previous copy out of this buffer has finished
                if (e != cudaSuccess) break;
                std::memcpy(pin, host + (size_t)r0 * row_bytes, (size_t)nr * row_bytes);
                e = cudaMemcpy2DAsync(dev + (size_t)r0 * dpitch, dpitch, pin, row_bytes, row_bytes, (size_t)nr, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(stager.ev[t][b], st);
            } else {
                e = cudaMemcpy2DAsync(pin, row_bytes, dev + (size_t)r0 * dpitch, dpitch, row_bytes, (size_t)nr, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaEventRecord(stager.ev[t][b], st);
                drain();                                         // the previous chunk, while this one is in flight
                pend_c = c;
                pend_b = b;
            }
        }
        drain();
        if (e == cudaSuccess) e = cudaEventRecord(stager.ev_done[t], st);
        errs[t] = e;
    };
    std::thread th[Stager::kThreads];
    for (int t = 1; t < Stager::kThreads; ++t) th[t] = std::thread(work, t);
    work(0);
    for (int t = 1; t < Stager::kThreads; ++t) th[t].join();
    for (int t = 0; t < Stager::kThreads; ++t) MG_CK(errs[t]);
    for (int t = 0; t < Stager::kThreads; ++t) MG_CK(cudaStreamWaitEvent(stream, stager.ev_done[t], 0));
}

// pinned (or registered / managed) host memory goes straight through the copy engine; ordinary pageable memory -- what
// a std::vector or a numpy array is -- through the staged multi-thread path when the transfer is large enough to matter
void Ctx::copy_rows(char* dev, size_t dpitch, char* host, size_t row_bytes, int rows, bool to_device)
{
    if (rows <= 0) return;
    bool pageable = false;
#ifndef MGB_EMU
    if ((size_t)rows * row_bytes >= Stager::kMinBytes) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, host) == cudaSuccess) pageable = (at.type == cudaMemoryTypeUnregistered);
        else cudaGetLastError();
    }
#endif
    if (pageable) {
        copy_rows_staged(dev, dpitch, host, row_bytes, rows, to_device);
        return;
    }
    if (to_device) MG_CK(cudaMemcpy2DAsync(dev, dpitch, host, row_bytes, row_bytes, (size_t)rows, cudaMemcpyHostToDevice, stream));
    else MG_CK(cudaMemcpy2DAsync(host, row_bytes, dev, dpitch, row_bytes, (size_t)rows, cudaMemcpyDeviceToHost, stream));
}

void Ctx::set_host(int level, Which w, const void* host)
{
    MG_REQUIRE(host != nullptr, "null host pointer");
    Level& lv = L(level);
    if (w == W_U) lv.u_zero = lv.u_interp = false;   // overwritten below (ring rows / columns of every buffer are zero already)
    const i64 n = lv.N - 1;
    const int ya = std::max(lv.st_lo, 1), yb = std::min(lv.st_hi, lv.N);
    char* dst = which_ptr(lv, w) + ((i64)ya * lv.pitch + 1) * esize;
    const char* src = (const char*)host + (i64)(ya - 1) * n * esize;
    copy_rows(dst, (size_t)lv.pitch * esize, const_cast<char*>(src), (size_t)n * esize, yb - ya, true);
    MG_CK(cudaStreamSynchronize(stream));
    set_halo(lv, w, lv.halo);
}

void Ctx::get_host(int level, Which w, void* host)
{
    MG_REQUIRE(host != nullptr, "null host pointer");
    Level& lv = L(level);
    if (w == W_U) materialize_u(lv);
    const i64 n = lv.N - 1;
    const int ya = lv.own_lo, yb = lv.own_hi;
    const char* src = which_ptr(lv, w) + ((i64)ya * lv.pitch + 1) * esize;
    char* dst = (char*)host + (i64)(ya - 1) * n * esize;
    copy_rows(const_cast<char*>(src), (size_t)lv.pitch * esize, dst, (size_t)n * esize, yb - ya, false);
    MG_CK(cudaStreamSynchronize(stream));
}

void Ctx::zero_u(int level)
{
    Level& lv = L(level);
    MG_CK(cudaMemsetAsync(lv.alloc[lv.cur], 0, lv.bytes, stream));
    lv.hv_u = lv.halo;
    lv.u_zero = lv.u_interp = false;
}

void Ctx::force_constant(double fval)
{
    Level& lv = L(cfg.finest_level);
    const double h = 1.0 / (double)lv.N;
    const double b = fval * h * h;  // P:182-184 summed over the six triangles of a node
    const int ya = std::max(lv.st_lo, 1), yb = std::min(lv.st_hi, lv.N);
    if (f64()) launch_fill<double>(stream, lc, (double*)lv.f, lv.pitch, lv.N, ya, yb, b);
    else launch_fill<float>(stream, lc, (float*)lv.f, lv.pitch, lv.N, ya, yb, (float)b);
    MG_CK(cudaGetLastError());
    lv.hv_f = lv.halo;
}

void Ctx::force_synthetic(unsigned long long seed)
{
    Level& lv = L(cfg.finest_level);
    const double h = 1.0 / (double)lv.N;
    const int ya = std::max(lv.st_lo, 1), yb = std::min(lv.st_hi, lv.N);   // owned rows and halo rows alike
    if (f64()) launch_fill_synthetic<double>(stream, lc, (double*)lv.f, lv.pitch, lv.N, ya, yb, h * h, seed);
    else launch_fill_synthetic<float>(stream, lc, (float*)lv.f, lv.pitch, lv.N, ya, yb, h * h, seed);
    MG_CK(cudaGetLastError());
    lv.hv_f = lv.halo;
}

unsigned long long Ctx::checksum(int level, Which w)
{
    Level& lv = L(level);
    if (w == W_U) materialize_u(lv);
    MG_REQUIRE(!capturing, "checksum inside a captured cycle");
    unsigned long long* d = reinterpret_cast<unsigned long long*>(d_norm);
    unsigned long long* h = reinterpret_cast<unsigned long long*>(h_norm);
    MG_CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), stream));
    const char* base = which_ptr(lv, w);
    if (f64()) launch_checksum<double>(stream, lc, (const double*)base, lv.pitch, lv.N, lv.own_lo, lv.own_hi, d);
    else launch_checksum<float>(stream, lc, (const float*)base, lv.pitch, lv.N, lv.own_lo, lv.own_hi, d);
    MG_CK(cudaGetLastError());
    MG_CK(cudaMemcpyAsync(h, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    MG_CK(cudaStreamSynchronize(stream));
    return *h;
}

// ---------------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------------
template <typename T>
void Ctx::smooth_t(int level, int nu)
{
    Level& lv = L(level);
    if (nu <= 0) return;
    materialize_u(lv);
    if (cfg.smoother == MG_SMOOTH_JACOBI) {
        // constants exactly as the oracle forms them (P:127, P:138-140)
        const T om = (T)cfg.omega;
        const T c0 = (T)(1.0 - (double)om);
        const T c1 = (T)((double)om / 4.0);
        int s = 0;
        while (s < nu) {
            const int did = fused_jacobi<T>(*this, lv, nu - s, c0, c1);  // temporally blocked sweeps (0 if n/a)
            if (did > 0) { s += did; continue; }
            ensure_halo(lv, W_U, 1);
            launch_jacobi<T>(stream, lc, (const T*)lv.u[lv.cur], (T*)lv.u[lv.cur ^ 1], (const T*)lv.f,
                             lv.pitch, lv.N, lv.own_lo, lv.own_hi, c0, c1);
            lv.cur ^= 1;
            lv.hv_u = 0;
            ++s;
        }
    } else if (!fused_rbgs(*this, lv, nu)) {
        for (int s = 0; s < nu; ++s)
            for (int colour = 0; colour < 2; ++colour) {
                ensure_halo(lv, W_U, 1);
                launch_rbgs<T>(stream, lc, (T*)lv.u[lv.cur], (const T*)lv.f, lv.pitch, lv.N, lv.own_lo,
                               lv.own_hi, colour);
                lv.hv_u = 0;
            }
    }
    MG_CK(cudaGetLastError());
}

template <typename T>
double Ctx::residual_t(int level, bool want_norm, bool store)
{
    Level& lv = L(level);
    materialize_u(lv);
    ensure_halo(lv, W_U, 1);
    if (store) lv.hv_r = 0;
    const int np = launch_residual<T>(stream, lc, (const T*)lv.u[lv.cur], (const T*)lv.f, (T*)lv.r, lv.pitch,
                                      lv.N, lv.own_lo, lv.own_hi, want_norm ? d_partials : nullptr, store);
    MG_CK(cudaGetLastError());
    if (!want_norm) return 0.0;
    MG_REQUIRE(!capturing, "norm read-back inside a captured cycle");
    launch_sum_partials(stream, lc, d_partials, np, d_norm);
    MG_CK(cudaGetLastError());
    return read_norm(lv);
}

template <typename T>
void Ctx::restrict_t(int fine_level, bool from_rhs)
{
    Level& lf = L(fine_level);
    Level& lcv = L(fine_level - 1);
    const T w = (T)cfg.restrict_weight;
    const T* src = (const T*)(from_rhs ? lf.f : lf.r);
    ensure_halo(lf, from_rhs ? W_F : W_R, 1);   // edge row of the neighbour's r (or f)
    if (!from_rhs) {
        lcv.cur = 0;                             // fixed buffer for the zero guess keeps graph replays valid
        lcv.u_zero = false;                      // real zeros are written below
    }
    if (lf.distributed && !lcv.distributed) {
        // agglomeration: every rank restricts its slab of coarse rows, then all-gathers
        int lo, hi;
        slab_rows(lcv.level, cfg.rank, cfg.world, &lo, &hi);
        launch_restrict<T>(stream, lc, src, lf.pitch, (T*)lcv.f, (T*)nullptr, lcv.pitch, lcv.N, lo, hi, w);
        MG_CK(cudaGetLastError());
        comm_allgather_rows(*this, lcv, lcv.f);
        if (!from_rhs) MG_CK(cudaMemsetAsync(lcv.alloc[lcv.cur], 0, lcv.bytes, stream));
        return;
    }
    launch_restrict<T>(stream, lc, src, lf.pitch, (T*)lcv.f, from_rhs ? nullptr : (T*)lcv.u[lcv.cur], lcv.pitch,
                       lcv.N, lcv.own_lo, lcv.own_hi, w);
    MG_CK(cudaGetLastError());
    lcv.hv_f = 0;
    if (lcv.distributed && !from_rhs) {
        comm_zero_halo(*this, lcv, lcv.u[lcv.cur]);   // halo rows of the zero guess
        lcv.hv_u = lcv.halo;
    }
}

template <typename T>
void Ctx::prolong_t(int fine_level, bool add)
{
    Level& lf = L(fine_level);
    Level& lcv = L(fine_level - 1);
    materialize_u(lcv);
    if (add) materialize_u(lf);
    else lf.u_zero = lf.u_interp = false;   // overwritten: whatever was pending is moot
    // the last owned fine row (odd) interpolates from the first coarse halo row
    ensure_halo(lcv, W_U, 1);
    launch_prolong<T>(stream, lc, (const T*)lcv.u[lcv.cur], lcv.pitch, (T*)lf.u[lf.cur], lf.pitch, lf.N, lf.own_lo,
                      lf.own_hi, add);
    MG_CK(cudaGetLastError());
    lf.hv_u = 0;
}

// u = A^-1 f on the coarsest level (see coarse.cuh).  Scratch: r and the non-current u buffer of that level (interior
// only: their zero rings stay intact, and every later writer of those arrays overwrites the whole interior).
template <typename T>
static void coarse_exact_t(Ctx& ctx, Level& lv)
{
    const int n = lv.N - 1;
    const i64 P = lv.pitch;
    const T* S = (const T*)ctx.d_dst_S;
    const T* d = (const T*)ctx.d_dst_d;
    auto interior = [&](char* base) { return (T*)base + P + 1; };
    const T* F = interior(lv.f);
    T* t1 = interior(lv.r);
    T* t2 = interior(lv.u[lv.cur ^ 1]);
    T* U = interior(lv.u[lv.cur]);
    const T c = (T)(4.0 / ((double)(n + 1) * (double)(n + 1)));
    launch_dense_product<T, 0>(ctx.stream, ctx.lc, F, P, S, n, t1, P, n, d, c);     // t1 = F S
    launch_dense_product<T, 1>(ctx.stream, ctx.lc, S, n, t1, P, t2, P, n, d, c);    // t2 = (S t1) ./ Lambda
    launch_dense_product<T, 0>(ctx.stream, ctx.lc, t2, P, S, n, t1, P, n, d, c);    // t1 = t2 S
    launch_dense_product<T, 2>(ctx.stream, ctx.lc, S, n, t1, P, U, P, n, d, c);     // u = c (S t1)
    MG_CK(cudaGetLastError());
}

void Ctx::coarse_exact(int level)
{
    Level& lv = L(level);
    MG_REQUIRE(exact_coarse() && level == cfg.coarsest_level && !lv.distributed, "coarse_exact: not the (replicated) coarsest level");
    lv.u_zero = lv.u_interp = false;     // the incoming iterate is ignored (M:137), the result overwrites it
    if (f64()) coarse_exact_t<double>(*this, lv);
    else coarse_exact_t<float>(*this, lv);
    lv.hv_u = lv.halo;
    lv.hv_r = 0;
}

void Ctx::smooth(int level, int nu) { f64() ? smooth_t<double>(level, nu) : smooth_t<float>(level, nu); }
double Ctx::residual(int level, bool want_norm, bool store)
{
    return f64() ? residual_t<double>(level, want_norm, store) : residual_t<float>(level, want_norm, store);
}
void Ctx::restrict_to(int fine_level, bool from_rhs)
{
    MG_REQUIRE(fine_level > cfg.coarsest_level, "no coarser level below coarsest_level");
    f64() ? restrict_t<double>(fine_level, from_rhs) : restrict_t<float>(fine_level, from_rhs);
}
void Ctx::prolong(int fine_level, bool add)
{
    MG_REQUIRE(fine_level > cfg.coarsest_level, "no coarser level below coarsest_level");
    f64() ? prolong_t<double>(fine_level, add) : prolong_t<float>(fine_level, add);
}

// ---------------------------------------------------------------------------------
// cycles.  cycle_rec mirrors vcyclemultigrid P:575-627 line by line.
// ---------------------------------------------------------------------------------
void Ctx::cycle_rec(int level, int nu1, int nu2, int gamma)
{
    if (level <= cfg.coarsest_level && exact_coarse()) {    // M:136-139: direct solve, no smoothing on this level
        coarse_exact(level);
        return;
    }
    if (fused_cycle_level(*this, level, nu1, nu2, gamma)) return;  // fused pre/post kernels or coarse tail
    smooth(level, nu1);                                     // P:581
    if (level <= cfg.coarsest_level) {                      // P:583
        smooth(level, nu2);                                 // P:585
        return;
    }
    residual(level, false, true);                           // P:604-608
    restrict_to(level, false);                              // P:611, P:613
    const int reps = (level - 1 <= cfg.coarsest_level) ? 1 : std::max(1, gamma);
    cycle_rec_visits(level - 1, nu1, nu2, gamma, reps);     // P:617
    prolong(level, true);                                   // P:620-624
    smooth(level, nu2);                                     // P:625
}

// `visits` consecutive visits of one level (same right-hand side, the iterate carries over): the gamma recursive
// calls of P:617 for a W-cycle, or consecutive cycles on the top level (P:646-648)
void Ctx::cycle_rec_visits(int level, int nu1, int nu2, int gamma, int visits)
{
    if (visits > 1 && chain && fused_cycle_chain(*this, level, nu1, nu2, gamma, visits)) return;
    for (int v = 0; v < visits; ++v) cycle_rec(level, nu1, nu2, gamma);
}

// `count` consecutive cycles from `level`, replayed as ONE CUDA graph when MG_GRAPH is set (cache key: level, sweep
// counts, cycle index, count and the host state the capture depends on)
void Ctx::cycles(int level, int nu1, int nu2, int gamma, int count)
{
    L(level);
    MG_REQUIRE(nu1 >= 0 && nu2 >= 0 && gamma >= 1 && gamma < 1000, "nu1, nu2 >= 0 and 1 <= gamma < 1000 required");
    if (count <= 0) return;
    if (count > 1 && !chain) {   // nothing to fuse across cycles: one (cached) graph per cycle
        for (int i = 0; i < count; ++i) cycles(level, nu1, nu2, gamma, 1);
        return;
    }
    if (!(cfg.flags & MG_GRAPH) || capturing || force_eager || (cfg.world > 1 && !graph_dist)) {
        cycle_rec_visits(level, nu1, nu2, gamma, count);
        return;
    }
    auto key = std::make_tuple(level, nu1, nu2, gamma + 1000 * count + (want_post_norm ? 500 : 0), state_blob());
    auto it = graphs.find(key);
    if (it == graphs.end()) {
        fused_pretune(*this, level, nu1, nu2, gamma);
        GraphEntry ge;
        const long long before = lc.n;
        cudaGraph_t g = nullptr;
        MG_CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        capturing = true;
        try {
            cycle_rec_visits(level, nu1, nu2, gamma, count);
        } catch (...) {
            capturing = false;
            cudaStreamEndCapture(stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        capturing = false;
        MG_CK(cudaStreamEndCapture(stream, &g));
        ge.kernels = lc.n - before;
        ge.state_after = state_blob();
        ge.post_norm = post_norm_done;
        cudaError_t ie = cudaGraphInstantiate(&ge.exec, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) throw MgError(MG_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
        it = graphs.emplace(key, ge).first;
        lc.n = before;  // counted again at launch below
    }
    MG_CK(cudaGraphLaunch(it->second.exec, stream));
    lc.n += it->second.kernels;
    ++graph_launches;
    set_state(it->second.state_after);
    if (want_post_norm) post_norm_done = it->second.post_norm;
}

void Ctx::cycle(int level, int nu1, int nu2, int gamma) { cycles(level, nu1, nu2, gamma, 1); }

void Ctx::fmg(int cycles, int nu1, int nu2)
{
    MG_REQUIRE(cycles >= 1, "cycles_per_level >= 1 required");
    for (int l = cfg.finest_level; l > cfg.coarsest_level; --l) restrict_to(l, true);  // P:641
    zero_u(cfg.coarsest_level);                                                          // P:630
    this->cycles(cfg.coarsest_level, nu1, nu2, 1, cycles);                               // P:635-637
    for (int l = cfg.coarsest_level + 1; l <= cfg.finest_level; ++l) {
        // P:645.  With visit chains on a single-GPU level the interpolation is left pending: the first PRE of the
        // level interpolates on the fly (fused.cu: k_stream_fmg_entry); any other reader materialises it
        if (chain && !L(l).distributed && !capturing) { materialize_u(L(l - 1)); L(l).u_zero = false; L(l).u_interp = true; }
        else prolong(l, false);
        this->cycles(l, nu1, nu2, 1, cycles);                                            // P:646-648
    }
}

// sqrt of the sum over the ranks of the sum r^2 a kernel left in d_norm
double Ctx::read_norm(const Level& lv)
{
    double sumsq = 0.0;
    if (lv.distributed) {
        sumsq = comm_sum(*this, d_norm);  // fixed rank order => same value on every rank
    } else {
        MG_CK(cudaMemcpyAsync(h_norm, d_norm, sizeof(double), cudaMemcpyDeviceToHost, stream));
        MG_CK(cudaStreamSynchronize(stream));
        sumsq = h_norm[0];
    }
    return std::sqrt(sumsq);
}

// ---------------------------------------------------------------------------------
// The tolerance loop as one graph launch (single GPU).  The host loop below pays a read-back, a host decision and a
// graph launch between two cycles (~40 us of idle GPU at 4097^2).  Here the decision is taken on the device: a CUDA
// graph with a conditional WHILE node whose body is one cycle (its last kernel leaves sum r^2 in d_norm) followed by
// k_solve_check, which records ||r_k||, decides `continue` and sets the node's condition.  The first cycle is peeled in
// front of the node, because the host-side buffer parities after the first cycle may differ from those it started from;
// from the second cycle on they must be a fixed point (checked at capture time, else the host loop runs).
// ---------------------------------------------------------------------------------
struct SolveCtl {
    double r0, rtol;
    int k, max_cycles;
    double hist[1];   // hist[0 .. max_cycles]
};

static __global__ void k_solve_check(cudaGraphConditionalHandle h, const double* __restrict__ sumsq, SolveCtl* __restrict__ c)
{
    if (threadIdx.x == 0) {
        const double rk = sqrt(*sumsq);
        const int k = ++c->k;
        c->hist[k] = rk;
        const bool done = (c->r0 == 0.0) || (rk <= c->rtol * c->r0) || (k >= c->max_cycles);
        cudaGraphSetConditional(h, done ? 0u : 1u);
    }
}

bool Ctx::solve_device_loop(double rtol, int max_cycles, int nu1, int nu2, int gamma, double r0, int* k_out, double* history)
{
    if (solve_loop_mode < 0) {
        const char* e = getenv("MGB200_SOLVE_GRAPH");     // on by default (verified on the B200); "0" = host loop
        solve_loop_mode = (e && e[0] == '0') ? 0 : 1;
    }
    if (solve_loop_mode != 1 || cfg.world != 1 || !(cfg.flags & MG_GRAPH) || !(cfg.flags & MG_FUSED) || capturing ||
        max_cycles < 1 || max_cycles > 100000)
        return false;
    const int top = cfg.finest_level;
    if (solve_ctl_cap < max_cycles) {
        if (d_solve_ctl) MG_CK(cudaFree(d_solve_ctl));
        d_solve_ctl = nullptr;
        MG_CK(cudaMalloc(&d_solve_ctl, sizeof(SolveCtl) + sizeof(double) * (size_t)(max_cycles + 1)));
        solve_ctl_cap = max_cycles;
    }
    SolveCtl* ctl = (SolveCtl*)d_solve_ctl;
    if (!body_stream) MG_CK(cudaStreamCreateWithFlags(&body_stream, cudaStreamNonBlocking));

    const std::string s0 = state_blob();
    auto key = std::make_tuple(top, nu1, nu2, gamma, s0);
    auto it = solve_graphs.find(key);
    if (it == solve_graphs.end()) {
        fused_pretune(*this, top, nu1, nu2, gamma);
        cudaGraph_t g = nullptr;
        cudaGraphExec_t exec = nullptr;
        std::string s1, s2;
        bool ok = true;
        cudaStream_t main_stream = stream;
        auto one_cycle = [&](cudaGraphConditionalHandle h) {
            want_post_norm = true;
            post_norm_done = false;
            cycle_rec_visits(top, nu1, nu2, gamma, 1);
            want_post_norm = false;
            if (!post_norm_done) ok = false;
            k_solve_check<<<1, 32, 0, stream>>>(h, d_norm, ctl);
            ++lc.n;
        };
        const long long launches_before = lc.n;
        try {
            MG_CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            capturing = true;
            cudaStreamCaptureStatus st;
            const cudaGraphNode_t* deps = nullptr;
            size_t ndeps = 0;
            MG_CK(cudaStreamGetCaptureInfo(stream, &st, nullptr, &g, &deps, &ndeps));
            cudaGraphConditionalHandle h;
            MG_CK(cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault));
            one_cycle(h);                                      // peeled first cycle: s0 -> s1
            s1 = state_blob();
            MG_CK(cudaStreamGetCaptureInfo(stream, &st, nullptr, &g, &deps, &ndeps));
            cudaGraphNodeParams np = {};
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = h;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            MG_CK(cudaGraphAddNode(&node, g, deps, ndeps, &np));
            cudaGraph_t body = np.conditional.phGraph_out[0];
            MG_CK(cudaStreamUpdateCaptureDependencies(stream, &node, 1, cudaStreamSetCaptureDependencies));
            MG_CK(cudaStreamEndCapture(stream, &g));
            // the body: one cycle from s1, which must lead back to s1
            MG_CK(cudaStreamBeginCaptureToGraph(body_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            stream = body_stream;
            one_cycle(h);
            s2 = state_blob();
            stream = main_stream;
            cudaGraph_t same = nullptr;
            MG_CK(cudaStreamEndCapture(body_stream, &same));
            capturing = false;
            if (s2 != s1) ok = false;
            if (ok) MG_CK(cudaGraphInstantiate(&exec, g, 0));
        } catch (...) {
            // leave no capture open, then fall back to the host loop for the life of this context
            stream = main_stream;
            capturing = false;
            want_post_norm = false;
            cudaGraph_t junk = nullptr;
            cudaStreamCaptureStatus st;
            if (cudaStreamIsCapturing(body_stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) cudaStreamEndCapture(body_stream, &junk);
            if (cudaStreamIsCapturing(main_stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) cudaStreamEndCapture(main_stream, &junk);
            cudaGetLastError();
            ok = false;
        }
        if (g) cudaGraphDestroy(g);
        lc.n = launches_before;
        set_state(s0);
        if (!ok) {
            if (exec) cudaGraphExecDestroy(exec);
            solve_loop_mode = 0;
            return false;
        }
        it = solve_graphs.emplace(key, std::make_pair(exec, s1)).first;
    }
    SolveCtl head;
    head.r0 = r0;
    head.rtol = rtol;
    head.k = 0;
    head.max_cycles = max_cycles;
    head.hist[0] = r0;
    MG_CK(cudaMemcpyAsync(ctl, &head, sizeof(SolveCtl), cudaMemcpyHostToDevice, stream));
    MG_CK(cudaGraphLaunch(it->second.first, stream));
    ++graph_launches;
    std::vector<double> hist((size_t)max_cycles + 1);
    SolveCtl tail;
    MG_CK(cudaMemcpyAsync(&tail, ctl, sizeof(SolveCtl), cudaMemcpyDeviceToHost, stream));
    MG_CK(cudaMemcpyAsync(hist.data(), (char*)ctl + offsetof(SolveCtl, hist), sizeof(double) * hist.size(), cudaMemcpyDeviceToHost, stream));
    MG_CK(cudaStreamSynchronize(stream));
    set_state(it->second.second);
    *k_out = tail.k;
    if (history)
        for (int i = 0; i <= tail.k; ++i) history[i] = hist[i];
    history_last = hist[tail.k];
    return true;
}

// Tolerance-controlled loop (the reference runs a fixed count, P:635).  The norm after each cycle comes out
// of the cycle's own last kernel on the finest level when the fused POST applies (k_stream_norm: no extra pass over the
// grid); otherwise from a residual pass without the store.
int Ctx::solve(double rtol, int max_cycles, int nu1, int nu2, int gamma, double* relres, double* history)
{
    const int top = cfg.finest_level;
    const double r0 = residual(top, true, false);
    if (history) history[0] = r0;
    int k = 0;
    double rk = r0;
    if (r0 > 0.0 && solve_device_loop(rtol, max_cycles, nu1, nu2, gamma, r0, &k, history)) {
        rk = history_last;
        if (relres) *relres = rk / r0;
        return k;
    }
    while (k < max_cycles) {
        want_post_norm = true;
        post_norm_done = false;
        try {
            cycle(top, nu1, nu2, gamma);
        } catch (...) {
            want_post_norm = false;
            throw;
        }
        want_post_norm = false;
        ++k;
        rk = post_norm_done ? read_norm(L(top)) : residual(top, true, false);
        if (history) history[k] = rk;
        if (r0 == 0.0 || rk <= rtol * r0) break;
    }
    if (relres) *relres = (r0 > 0.0) ? rk / r0 : 0.0;
    return k;
}

int Ctx::time_phases(int level, int nu1, int nu2, int gamma, int reps, double* out)
{
    MG_REQUIRE(reps >= 1 && out != nullptr, "reps >= 1 and an output array of 5 x 32 doubles required");
    L(level);
    for (int i = 0; i < 5 * 32; ++i) out[i] = 0.0;
    struct Restore {
        Ctx& c;
        ~Restore()
        {
            c.force_eager = c.phase_on = false;
            for (PhaseRec& r : c.phase_log) {
                if (r.e0) cudaEventDestroy(r.e0);
                if (r.e1) cudaEventDestroy(r.e1);
            }
            c.phase_log.clear();
        }
    } restore{*this};
    force_eager = true;
    cycles(level, nu1, nu2, gamma, 1);      // warm (tuner, first exchange of the static right-hand side), not logged
    phase_on = true;
    for (int i = 0; i < reps; ++i) cycles(level, nu1, nu2, gamma, 1);
    MG_CK(cudaStreamSynchronize(stream));
    for (const PhaseRec& r : phase_log) {
        float ms = 0.f;
        MG_CK(cudaEventElapsedTime(&ms, r.e0, r.e1));
        if (r.kind >= 0 && r.kind < 5 && r.level >= 0 && r.level < 32) out[r.kind * 32 + r.level] += (double)ms / reps;
    }
    return (int)(phase_log.size() / (size_t)reps);
}

float Ctx::time_op(int op, int level, int reps)
{
    MG_REQUIRE(reps >= 1, "reps >= 1 required");
    L(level);
    EventTimer timer(stream);
    auto run = [&]() {
        switch (op) {
            case MG_OP_SMOOTH1: smooth(level, 1); break;
            case MG_OP_SMOOTH2: smooth(level, 2); break;
            case MG_OP_SMOOTH3: smooth(level, 3); break;
            case MG_OP_SMOOTH4:
                if (!fused_time_sweeps4(*this, level)) throw MgError(MG_ERR_STATE, "4-sweep temporally blocked kernel not available for this config");
                break;
            case MG_OP_RESIDUAL: residual(level, false, true); break;
            case MG_OP_RESIDUAL_NORM: {
                Level& lv = L(level);
                if (f64()) launch_residual<double>(stream, lc, (const double*)lv.u[lv.cur], (const double*)lv.f, (double*)lv.r, lv.pitch, lv.N, lv.own_lo, lv.own_hi, d_partials, false);
                else launch_residual<float>(stream, lc, (const float*)lv.u[lv.cur], (const float*)lv.f, (float*)lv.r, lv.pitch, lv.N, lv.own_lo, lv.own_hi, d_partials, false);
                break;
            }
            case MG_OP_RESTRICT: restrict_to(level, false); break;
            case MG_OP_PROLONG: prolong(level, true); break;
            case MG_OP_PRE_FUSED:
                if (!fused_time_hook(*this, level, true)) throw MgError(MG_ERR_STATE, "fused pre-smoothing kernel not available for this level/config");
                break;
            case MG_OP_POST_FUSED:
                if (!fused_time_hook(*this, level, false)) throw MgError(MG_ERR_STATE, "fused post-smoothing kernel not available for this level/config");
                break;
            case MG_OP_POSTPRE_FUSED:
                if (!fused_time_hook_postpre(*this, level)) throw MgError(MG_ERR_STATE, "POST+PRE chain kernel not available (MGB200_CHAIN=1, fused, this level/config)");
                break;
            default: throw MgError(MG_ERR_ARG, "unknown op");
        }
    };
    run();  // warm
    MG_CK(cudaStreamSynchronize(stream));
    timer.start();
    for (int i = 0; i < reps; ++i) run();
    return timer.stop();
}

}  // namespace mgb
