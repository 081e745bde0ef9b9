// tests/host_emul/tile_emul.cpp — TEST INFRASTRUCTURE.  Runs the phases of the tile kernels
// (multigrid_nikhil_c-_b200/csrc/tile_core.h, the exact code the CUDA kernel k_tile executes)
// on the CPU: for every CTA, for every phase, for every thread.  A phase only reads what earlier
// phases wrote (that is what the block barrier between phases guarantees on the GPU), so running
// the threads of a phase one after the other is an exact emulation.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -I<csrc> tile_emul.cpp -o libtile_emul.so
#include <vector>

#include "tile_core.h"

using namespace mgb;

template <typename T, int NS, int MODE, bool RBGS, int TY, int TX>
static void run_tiles(const TileArgs<T>& a, int nthr)
{
    typedef TileCfg<T, NS, MODE, TY, TX> C;
    const int gx = (a.N + TX - 1) / TX, gy = (a.yb - a.ya + TY - 1) / TY;
    std::vector<T> smem(C::SMEM_ELEMS);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            for (auto& v : smem) v = (T)12345.678;   // poison: nothing may depend on stale shared memory
            for (int ph = 0; ph < C::NPHASES; ++ph)
                for (int tid = 0; tid < nthr; ++tid)
                    tile_phase<T, NS, MODE, RBGS, TY, TX>(a, smem.data(), bx, by, tid, nthr, ph);
        }
}

template <typename T, int MODE, bool RBGS, int TY, int TX>
static int dispatch_ns(const TileArgs<T>& a, int ns, int nthr)
{
    switch (ns) {
        case 1: run_tiles<T, 1, MODE, RBGS, TY, TX>(a, nthr); return 0;
        case 2: run_tiles<T, 2, MODE, RBGS, TY, TX>(a, nthr); return 0;
        case 3: run_tiles<T, 3, MODE, RBGS, TY, TX>(a, nthr); return 0;
        case 4: run_tiles<T, 4, MODE, RBGS, TY, TX>(a, nthr); return 0;
    }
    return 1;
}

template <typename T>
static int run(int mode, int ns, int rbgs, int tile, const TileArgs<T>& a, int nthr)
{
#define MG_CASE(M, R)                                                                          \
    if (mode == M && rbgs == R) {                                                              \
        if (tile == 0) return dispatch_ns<T, M, (R != 0), 16, 32>(a, ns, nthr);                \
        return dispatch_ns<T, M, (R != 0), 32, 64>(a, ns, nthr);                               \
    }
    MG_CASE(TILE_SWEEPS, 0) MG_CASE(TILE_SWEEPS, 1) MG_CASE(TILE_PRE, 0) MG_CASE(TILE_PRE, 1)
    MG_CASE(TILE_POST, 0) MG_CASE(TILE_POST, 1)
#undef MG_CASE
    return 2;
}

extern "C" {

// Arrays are in the padded device layout (rows 0..N, `pitch` elements per row, zero ring), given as
// pointers to the element (row_lo, 0) of the backed rows [row_lo, row_hi).
#define MG_EMUL(NAME, T)                                                                                        \
    int NAME(int mode, int ns, int rbgs, int tile, int nthr, int N, long long pitch, int ya, int yb, int row_lo, \
             int row_hi, const T* u_in, T* u_out, const T* f, double c0, double c1, double w, T* fc, T* uc,       \
             const T* ec, long long pitch_c, int crow_lo, int crow_hi)                                            \
    {                                                                                                             \
        TileArgs<T> a;                                                                                            \
        a.u_in = u_in - (long long)row_lo * pitch;                                                                \
        a.u_out = u_out - (long long)row_lo * pitch;                                                              \
        a.f = f - (long long)row_lo * pitch;                                                                      \
        a.pitch = pitch; a.N = N; a.ya = ya; a.yb = yb; a.row_lo = row_lo; a.row_hi = row_hi;                     \
        a.c0 = (T)c0; a.c1 = (T)c1; a.w = (T)w;                                                                   \
        a.fc = fc ? fc - (long long)crow_lo * pitch_c : nullptr;                                                  \
        a.uc = uc ? uc - (long long)crow_lo * pitch_c : nullptr;                                                  \
        a.ec = ec ? ec - (long long)crow_lo * pitch_c : nullptr;                                                  \
        a.pitch_c = pitch_c; a.Nc = N / 2; a.crow_lo = crow_lo; a.crow_hi = crow_hi;                              \
        return run<T>(mode, ns, rbgs, tile, a, nthr);                                                             \
    }
MG_EMUL(tile_emul_f64, double)
MG_EMUL(tile_emul_f32, float)

}  // extern "C"
