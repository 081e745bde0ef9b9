// stream.cuh — warp-streaming, temporally blocked level kernels (MG_FUSED).
//
// One WARP owns a column strip of 32 lanes x V columns (64 fp64 / 128 fp32 columns) and a
// chunk of rows, and streams down the rows.  Rows of u and f (and of the coarse
// correction for the post kernel) are prefetched D rows ahead with cp.async into a
// warp-private shared-memory ring (each lane only ever reads back the 16 bytes it copied
// itself, so no barrier of any kind is needed).  All NS smoothing stages -- and, in the
// PRE kernel, the residual and the full-weighting restriction; in the POST kernel, the
// bilinear prolongation + correction -- run as a software pipeline over that stream:
// stage s works on row y-s while row y arrives, every stage keeps its 3-row stencil
// window in registers and gets its left/right neighbours with warp shuffles.  The
// outermost HLANES lanes on each side only feed the stencils (their results are never
// stored): strips overlap by 2*V*HLANES columns.
//
//   SWEEPS: u_out = S^NS u                               3 S B/pt per launch (not per sweep)
//   PRE   : u_out = S^NS u;  f_c = R (f - A u_out); u_c = 0     (3 + 1/4 [+1/4]) S B/pt
//   POST  : u_out = S^NS (u + P e_c)                            (3 + 1/4) S B/pt
//
// S is one weighted-Jacobi sweep, or one *half* sweep (one colour) of red-black
// Gauss-Seidel (then NS = 2 x sweeps).  Every point value is produced by the same
// expression, in the same order, as in the unfused kernels (common.cuh), so results are
// bit-identical to them and to the oracle.
#pragma once

#include "common.cuh"

namespace mgb {

enum { MODE_SWEEPS = 0, MODE_PRE = 1, MODE_POST = 2 };

constexpr int kStreamWarps = 4;  // warps (independent work items) per CTA

template <typename T, int NS, int MODE>
struct StreamCfg {
    static constexpr int V = Vec<T>::N;
    static constexpr int HL = NS + (MODE == MODE_PRE ? 2 : 0);                          // columns needed to the left
    static constexpr int HR = NS + (MODE == MODE_PRE ? 1 : (MODE == MODE_POST ? 1 : 0));  // ... to the right
    static constexpr int HMAX = HL > HR ? HL : HR;
    static constexpr int HLANES = (HMAX + V - 1) / V;                                   // halo lanes per side
    static constexpr int OUTW = 32 * V - 2 * V * HLANES;                                // output columns per strip
    static constexpr int HT = NS + (MODE == MODE_PRE ? 2 : (MODE == MODE_POST ? 1 : 0));  // rows needed above
    static constexpr int HB = NS + (MODE == MODE_PRE ? 2 : 0);                          // rows needed below
    static constexpr int D = (NS >= 4) ? 5 : 6;                                         // prefetch distance (rows)
    static constexpr int DEPTH = D + NS + 3;                                            // ring slots
    static constexpr int NW = NS + (MODE == MODE_PRE ? 1 : 0);                          // register windows of u_s
    static constexpr int SLOT_ELEMS = 32 * V * 2 + (MODE == MODE_POST ? 32 * (V / 2) : 0);
    static constexpr size_t SMEM_BYTES = (size_t)kStreamWarps * DEPTH * SLOT_ELEMS * sizeof(T);
};

template <typename T>
struct StreamArgs {
    const T* u_in;
    T* u_out;
    const T* f;
    i64 pitch;
    int N;
    int ya, yb;          // output rows [ya, yb)
    int row_lo, row_hi;  // rows backed by storage [row_lo, row_hi); anything else reads as zero
    int ry;              // output rows per chunk
    int strips;
    int nitems;          // strips * chunks
    T c0, c1, w;
    T* fc;               // PRE: coarse right-hand side / zero guess
    T* uc;
    const T* ec;         // POST: coarse correction
    i64 pitch_c;
    int Nc;
    int crow_lo, crow_hi;  // coarse rows backed by storage
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <typename T, int NS, int MODE, bool RBGS>
struct Streamer {
    typedef StreamCfg<T, NS, MODE> C;
    static constexpr int V = C::V;
    static constexpr int H = V / 2;
    static constexpr unsigned FULL = 0xffffffffu;

    const StreamArgs<T>& a;
    T* ring;      // this lane's 16 bytes of slot 0 (u part)
    int lane, c;  // lane id, first column
    int out_lo, out_hi, y0, y1;
    int slot;     // ring slot of the row arriving in the current step

    // register windows: W[s][phase slot][k], neighbours of each row
    T W[C::NW > 0 ? C::NW : 1][3][V];
    T WL[C::NW > 0 ? C::NW : 1][3], WR[C::NW > 0 ? C::NW : 1][3];
    T R[3][V], RL[3];  // PRE: residual window (left neighbours only)

    __device__ __forceinline__ Streamer(const StreamArgs<T>& a_) : a(a_) {}

    __device__ __forceinline__ T* slot_u(int s) const { return ring + (size_t)s * C::SLOT_ELEMS; }
    __device__ __forceinline__ T* slot_f(int s) const { return ring + (size_t)s * C::SLOT_ELEMS + 32 * V; }
    __device__ __forceinline__ T* slot_c(int s) const { return ring + (size_t)s * C::SLOT_ELEMS + 64 * V - lane * V + lane * H; }
    __device__ __forceinline__ int slot_back(int k) const { int s = slot - k; return s < 0 ? s + C::DEPTH : s; }

    // prefetch row y into ring slot s (always commits exactly one group)
    __device__ __forceinline__ void issue(int y, int s)
    {
        const bool vrow = (y >= a.row_lo) && (y < a.row_hi);
        const bool v = vrow && (c < a.pitch);
        const i64 off = v ? ((i64)y * a.pitch + c) : ((i64)a.row_lo * a.pitch);
        cp_async16(slot_u(s), a.u_in + off, v);
        cp_async16(slot_f(s), a.f + off, v);
        if (MODE == MODE_POST) {
            const int ic = (y + 1) >> 1;
            const int jc = c >> 1;
            const bool vc = (ic >= a.crow_lo) && (ic < a.crow_hi) && (y >= 0) && (jc < a.pitch_c);
            const i64 offc = vc ? ((i64)ic * a.pitch_c + jc) : ((i64)a.crow_lo * a.pitch_c);
            cp_async8(slot_c(s), a.ec + offc, vc);
        }
        cp_async_commit();
    }

    __device__ __forceinline__ bool interior(int y, int x) const { return y >= 1 && y < a.N && x >= 1 && x < a.N; }

    template <int NEW>
    __device__ __forceinline__ void put_row(int s, const T (&v)[V])
    {
#pragma unroll
        for (int k = 0; k < V; ++k) W[s][NEW][k] = v[k];
        WL[s][NEW] = __shfl_up_sync(FULL, v[V - 1], 1);
        WR[s][NEW] = __shfl_down_sync(FULL, v[0], 1);
    }

    // one pipeline step: row y of the input arrives
    template <int PH>
    __device__ __forceinline__ void step(int y)
    {
        constexpr int NEW = PH, MID = (PH + 2) % 3, OLD = (PH + 1) % 3;
        T cur[V];
        // ---- stage 0: the incoming row (POST: plus the interpolated coarse correction) ----
        ldv<T>(slot_u(slot), cur);
        if (MODE == MODE_POST) {
            T ca[H + 1], cb[H + 1], e[V];
            const T* pa = slot_c(slot);
#pragma unroll
            for (int k = 0; k < H; ++k) cb[k] = pa[k];
            cb[H] = __shfl_down_sync(FULL, cb[0], 1);
            if (y & 1) {
                const T* pp = slot_c(slot_back(1));
#pragma unroll
                for (int k = 0; k < H; ++k) ca[k] = pp[k];
                ca[H] = __shfl_down_sync(FULL, ca[0], 1);
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    e[2 * k] = (T)0.5 * (ca[k] + cb[k]);                                       // P:407
                    e[2 * k + 1] = (T)0.25 * (((ca[k] + cb[k]) + ca[k + 1]) + cb[k + 1]);      // P:419
                }
            } else {  // (y parity is warp-uniform, so both branches are convergent)
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    e[2 * k] = cb[k];                                                          // P:401
                    e[2 * k + 1] = (T)0.5 * (cb[k] + cb[k + 1]);                               // P:413
                }
            }
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T val = cur[k] + e[k];                                                   // P:623
                cur[k] = interior(y, c + k) ? val : (T)0;
            }
        }
        put_row<NEW>(0, cur);

        // ---- smoothing stages: stage s produces row y-s of u_s from the window of u_{s-1} ----
#pragma unroll
        for (int s = 1; s <= NS; ++s) {
            const int rs = y - s;
            T ff[V], o[V];
            ldv<T>(slot_f(slot_back(s)), ff);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T l = (k == 0) ? WL[s - 1][MID] : W[s - 1][MID][k - 1];
                const T r = (k == V - 1) ? WR[s - 1][MID] : W[s - 1][MID][k + 1];
                const T sig = sigma4<T>(W[s - 1][OLD][k], W[s - 1][NEW][k], l, r);
                T val;
                if (RBGS) {
                    const int colour = (s - 1) & 1;
                    val = (((rs + c + k) & 1) == colour) ? gs_pt<T>(ff[k], sig) : W[s - 1][MID][k];
                } else {
                    val = jacobi_pt<T>(a.c0, a.c1, W[s - 1][MID][k], ff[k], sig);
                }
                o[k] = interior(rs, c + k) ? val : (T)0;
            }
            if (s < C::NW) put_row<NEW>(s, o);
            if (s == NS) {
                if (rs >= y0 && rs < y1 && c >= out_lo && c < out_hi && c < a.N)
                    stv<T>(a.u_out + (i64)rs * a.pitch + c, o);
            }
        }

        // ---- PRE: residual of u_NS (row y-NS-1) and full weighting (coarse row when that row is odd) ----
        if (MODE == MODE_PRE) {
            const int rr = y - NS - 1;
            T ff[V];
            ldv<T>(slot_f(slot_back(NS + 1)), ff);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const T l = (k == 0) ? WL[NS][MID] : W[NS][MID][k - 1];
                const T r = (k == V - 1) ? WR[NS][MID] : W[NS][MID][k + 1];
                const T val = resid_pt<T>(W[NS][MID][k], ff[k], sigma4<T>(W[NS][OLD][k], W[NS][NEW][k], l, r));
                R[NEW][k] = interior(rr, c + k) ? val : (T)0;
            }
            RL[NEW] = __shfl_up_sync(FULL, R[NEW][V - 1], 1);
            if (rr & 1) {
                const int yc = rr - 1;  // fine centre row 2I (MID of the residual window)
                T o[H];
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const int k = 2 * j;
                    const T nw = (k == 0) ? RL[OLD] : R[OLD][k - 1];
                    const T wv = (k == 0) ? RL[MID] : R[MID][k - 1];
                    const T sw = (k == 0) ? RL[NEW] : R[NEW][k - 1];
                    const T val = fw_pt<T>(a.w, nw, R[OLD][k + 1], sw, R[NEW][k + 1], wv, R[MID][k + 1],
                                           R[OLD][k], R[NEW][k], R[MID][k]);
                    const int J = (c + k) >> 1;
                    o[j] = (J >= 1 && J < a.Nc) ? val : (T)0;
                }
                if (yc >= y0 && yc < y1 && c >= out_lo && c < out_hi && c < a.N) {
                    const i64 offc = (i64)(yc >> 1) * a.pitch_c + (c >> 1);
                    if constexpr (H == 1) {
                        a.fc[offc] = o[0];
                        if (a.uc) a.uc[offc] = (T)0;
                    } else {
                        *reinterpret_cast<float2*>(a.fc + offc) = make_float2((float)o[0], (float)o[H - 1]);
                        if (a.uc) *reinterpret_cast<float2*>(a.uc + offc) = make_float2(0.f, 0.f);
                    }
                }
            }
        }
        slot = (slot + 1 == C::DEPTH) ? 0 : slot + 1;
    }

    __device__ __forceinline__ void run(T* ring_base, int warp, int lane_, int item)
    {
        lane = lane_;
        const int chunk = item / a.strips;
        const int strip = item - chunk * a.strips;
        const int X0 = strip * C::OUTW;
        c = X0 + V * lane;
        out_lo = (strip == 0) ? 0 : X0 + V * C::HLANES;
        out_hi = X0 + V * C::HLANES + C::OUTW;
        y0 = a.ya + chunk * a.ry;
        y1 = min(y0 + a.ry, a.yb);
        ring = ring_base + (size_t)warp * C::DEPTH * C::SLOT_ELEMS + lane * V;
        const int ylo = y0 - C::HT, yhi = y1 - 1 + C::HB;

#pragma unroll
        for (int s = 0; s < (C::NW > 0 ? C::NW : 1); ++s)
#pragma unroll
            for (int p = 0; p < 3; ++p) {
#pragma unroll
                for (int k = 0; k < V; ++k) W[s][p][k] = (T)0;
                WL[s][p] = (T)0;
                WR[s][p] = (T)0;
            }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int k = 0; k < V; ++k) R[p][k] = (T)0;
            RL[p] = (T)0;
        }

        // prologue: D rows in flight
#pragma unroll
        for (int i = 0; i < C::D; ++i) issue(ylo + i, i);
        slot = 0;
        int pf = C::D;  // ring slot the next prefetch goes to
        for (int y = ylo; y <= yhi; y += 3) {
            issue(y + C::D, pf);
            pf = (pf + 1 == C::DEPTH) ? 0 : pf + 1;
            cp_async_wait<C::D>();
            step<0>(y);
            issue(y + 1 + C::D, pf);
            pf = (pf + 1 == C::DEPTH) ? 0 : pf + 1;
            cp_async_wait<C::D>();
            step<1>(y + 1);
            issue(y + 2 + C::D, pf);
            pf = (pf + 1 == C::DEPTH) ? 0 : pf + 1;
            cp_async_wait<C::D>();
            step<2>(y + 2);
        }
        cp_async_wait<0>();
    }
};

template <typename T, int NS, int MODE, bool RBGS>
__global__ void __launch_bounds__(kStreamWarps * 32)
k_stream(const StreamArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kStreamWarps + warp;
    if (item >= a.nitems) return;  // warp-uniform
    Streamer<T, NS, MODE, RBGS> st(a);
    st.run(reinterpret_cast<T*>(stream_smem), warp, threadIdx.x & 31, item);
}

}  // namespace mgb
