// tile_core.h — shared-memory TILE kernels for the mid levels (129^2 .. 1025^2), host-device core.
//
// On these levels a launch of the warp-streaming kernel (stream.cuh) is a 16..24-step dependent
// chain per warp and costs 6-11 us regardless of the grid size (profiles/r01_bench_fused_v6:
// levels 7-10 are 20 % of a 4097^2 V-cycle).  Here one CTA owns a TY x TX tile: it stages u and f
// with the halo the fused operator needs, runs every stage over the whole tile with all threads
// (a block barrier between stages), and writes the tile -- the latency of a launch becomes a
// handful of barriers instead of a row-by-row march.
//
//   SWEEPS: u_out = S^NS u        PRE: u_out = S^NS u, f_c = R(f - A u_out), u_c = 0
//   POST  : u_out = S^NS (u + P e_c)
//
// Every value is produced by the formulas of formulas.h in the same order as everywhere else, so
// results are bit-identical to the other paths.  The body is written as PHASES separated by block
// barriers and uses nothing but plain C++: the CUDA kernel (tile.cuh) runs
//     for (ph = 0; ph < NPHASES; ++ph) { tile_phase(..., ph); __syncthreads(); }
// and the host emulation (tests/host_emul/tile_emul.cpp) runs the same phases thread by thread,
// which lets the CPU test-suite check this code against the oracle without a GPU.
#pragma once

#include "formulas.h"

namespace mgb {

enum { TILE_SWEEPS = 0, TILE_PRE = 1, TILE_POST = 2 };

template <typename T>
struct TileArgs {
    const T* u_in;
    T* u_out;
    const T* f;
    i64 pitch;
    int N;
    int ya, yb;            // output rows [ya, yb)
    int row_lo, row_hi;    // rows backed by storage; anything else reads as zero
    T c0, c1, w;
    T* fc;                 // PRE: coarse right-hand side
    T* uc;                 //      zero coarse guess (may be null)
    const T* ec;           // POST: coarse correction
    i64 pitch_c;
    int Nc;
    int crow_lo, crow_hi;  // coarse rows backed by storage
};

template <typename T, int NS, int MODE, int TY, int TX>
struct TileCfg {
    static constexpr int HL = NS + (MODE == TILE_PRE ? 2 : 0);
    static constexpr int HR = NS + (MODE == TILE_PRE ? 1 : 0);
    static constexpr int HT = NS + (MODE == TILE_PRE ? 2 : 0);
    static constexpr int HB = NS + (MODE == TILE_PRE ? 2 : 0);
    static constexpr int W = TX + HL + HR;      // staged columns
    static constexpr int H = TY + HT + HB;      // staged rows
    static constexpr int P = W + 1;             // smem pitch (odd: rows fall on different banks)
    static constexpr int BUF = H * P;           // elements per buffer
    static constexpr int SMEM_ELEMS = 3 * BUF;  // u ping, u pong / residual, f
    // phases: load | NS smoothing stages | (PRE: residual) | store (+ PRE: restriction)
    static constexpr int NPHASES = 1 + NS + (MODE == TILE_PRE ? 1 : 0) + 1;
};

// value of the coarse grid at (I, J) with the zero ring / unbacked rows reading as zero
template <typename T>
MG_HD T tile_coarse_at(const TileArgs<T>& a, int I, int J)
{
    if (I < a.crow_lo || I >= a.crow_hi || J < 0 || J > a.Nc) return (T)0;
    return a.ec[(i64)I * a.pitch_c + J];
}

// One phase of one thread.  smem holds 3 buffers of TileCfg::BUF elements.
template <typename T, int NS, int MODE, bool RBGS, int TY, int TX>
MG_HD void tile_phase(const TileArgs<T>& a, T* smem, int bx, int by, int tid, int nthr, int phase)
{
    typedef TileCfg<T, NS, MODE, TY, TX> C;
    constexpr int W = C::W, H = C::H, P = C::P;
    T* const bufA = smem;
    T* const bufB = smem + C::BUF;
    T* const bufF = smem + 2 * C::BUF;
    const int x0 = bx * TX, y0 = a.ya + by * TY;
    const int x1 = x0 + TX, y1 = (y0 + TY < a.yb) ? y0 + TY : a.yb;
    const int gx0 = x0 - C::HL, gy0 = y0 - C::HT;   // global coordinates of staged element (0, 0)

    if (phase == 0) {
        // ---- load u (POST: u + P e) and f, zero outside the grid / the backed rows ----
        for (int idx = tid; idx < H * W; idx += nthr) {
            const int ry = idx / W, rx = idx - ry * W;
            const int gy = gy0 + ry, gx = gx0 + rx;
            const bool backed = (gy >= a.row_lo) && (gy < a.row_hi) && (gx >= 0) && (gx <= a.N);
            T u = backed ? a.u_in[(i64)gy * a.pitch + gx] : (T)0;
            const T fv = backed ? a.f[(i64)gy * a.pitch + gx] : (T)0;
            if (MODE == TILE_POST) {
                const bool interior = (gy >= 1) && (gy < a.N) && (gx >= 1) && (gx < a.N);
                if (interior) {
                    const int I = gy >> 1, J = gx >> 1;
                    const T c00 = tile_coarse_at<T>(a, I, J);
                    const T c10 = (gy & 1) ? tile_coarse_at<T>(a, I + 1, J) : (T)0;
                    const T c01 = (gx & 1) ? tile_coarse_at<T>(a, I, J + 1) : (T)0;
                    const T c11 = ((gy & 1) && (gx & 1)) ? tile_coarse_at<T>(a, I + 1, J + 1) : (T)0;
                    u = u + prolong_pt<T>(gy, gx, c00, c10, c01, c11);     // P:623
                } else {
                    u = (T)0;
                }
            }
            bufA[ry * P + rx] = u;
            bufF[ry * P + rx] = fv;
        }
        return;
    }

    if (phase <= NS) {
        // ---- smoothing stage s: region shrinks by one per side per stage ----
        const int s = phase;
        const T* src = (s & 1) ? bufA : bufB;
        T* dst = (s & 1) ? bufB : bufA;
        const int rw = W - 2 * s, rh = H - 2 * s;
        for (int idx = tid; idx < rh * rw; idx += nthr) {
            const int ry = s + idx / rw, rx = s + idx % rw;
            const int gy = gy0 + ry, gx = gx0 + rx;
            const int i = ry * P + rx;
            T val = (T)0;
            if ((gy >= 1) && (gy < a.N) && (gx >= 1) && (gx < a.N)) {
                const T sig = sigma4<T>(src[i - P], src[i + P], src[i - 1], src[i + 1]);
                if (RBGS) {
                    const int colour = (s - 1) & 1;
                    val = (((gy + gx) & 1) == colour) ? gs_pt<T>(bufF[i], sig) : src[i];
                } else {
                    val = jacobi_pt<T>(a.c0, a.c1, src[i], bufF[i], sig);
                }
            }
            dst[i] = val;
        }
        return;
    }

    // after NS stages u_NS lives in: NS odd -> bufB, NS even -> bufA
    const T* cur = (NS & 1) ? bufB : bufA;
    T* other = (NS & 1) ? bufA : bufB;

    if (MODE == TILE_PRE && phase == NS + 1) {
        // ---- residual of u_NS on the region one wider than the coarse stencils need ----
        const int s = NS + 1;
        const int rw = W - 2 * s, rh = H - 2 * s;
        for (int idx = tid; idx < rh * rw; idx += nthr) {
            const int ry = s + idx / rw, rx = s + idx % rw;
            const int gy = gy0 + ry, gx = gx0 + rx;
            const int i = ry * P + rx;
            T val = (T)0;
            if ((gy >= 1) && (gy < a.N) && (gx >= 1) && (gx < a.N))
                val = resid_pt<T>(cur[i], bufF[i], sigma4<T>(cur[i - P], cur[i + P], cur[i - 1], cur[i + 1]));
            other[i] = val;
        }
        return;
    }

    // ---- last phase: store the tile (PRE: and the coarse points whose centre lies in it) ----
    {
        const int th = y1 - y0;
        for (int idx = tid; idx < th * TX; idx += nthr) {
            const int ty = idx / TX, tx = idx - ty * TX;
            const int gy = y0 + ty, gx = x0 + tx;
            if (gx < a.N) a.u_out[(i64)gy * a.pitch + gx] = cur[(C::HT + ty) * P + (C::HL + tx)];
        }
        if (MODE == TILE_PRE) {
            const T* r = other;
            constexpr int CW = TX / 2 + 1, CH = TY / 2 + 1;
            const int I0 = (y0 + 1) >> 1, J0 = (x0 + 1) >> 1;   // first coarse row / column with centre >= y0 / x0
            for (int idx = tid; idx < CH * CW; idx += nthr) {
                const int I = I0 + idx / CW, J = J0 + idx % CW;
                const int gy = 2 * I, gx = 2 * J;
                if (gy >= y1 || gx >= x1 || J < 1 || J >= a.Nc || I < 1 || I >= a.Nc) continue;
                const int i = (gy - gy0) * P + (gx - gx0);
                const T val = fw_pt<T>(a.w, r[i - P - 1], r[i - P + 1], r[i + P - 1], r[i + P + 1],
                                       r[i - 1], r[i + 1], r[i - P], r[i + P], r[i]);
                a.fc[(i64)I * a.pitch_c + J] = val;
                if (a.uc) a.uc[(i64)I * a.pitch_c + J] = (T)0;
            }
        }
    }
}

}  // namespace mgb
