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
//   POSTPRE: POST of one visit of a level fused with PRE of the NEXT visit of the same level (consecutive cycles on
//            the finest level -- fullmultigrid runs mu0+1 of them per level, P:646-648 -- and the gamma visits of a
//            W-cycle):  u_out = S^(nu2+nu1) (u + P e_c);  f_c = R (f - A u_out);  u_c = 0      (3 + 1/2) S B/pt
//            instead of 6.5 S for POST + PRE: the iterate between the two visits never goes to memory.
//
// S is one weighted-Jacobi sweep, or one *half* sweep (one colour) of red-black
// Gauss-Seidel (then NS = 2 x sweeps).  Every point value is produced by the same
// expression, in the same order, as in the unfused kernels (common.cuh), so results are
// bit-identical to them and to the oracle.
//
// The kernel is instruction-issue sensitive (profiles/r01_v2_*: 69 % issue utilisation,
// half of it integer work), hence:
//  - the ring has 12 slots = 4 blocks of 3; the row loop is unrolled by 3 so that window
//    rotation is pure register renaming and every ring address is (block pointer +
//    compile-time constant);
//  - Dirichlet masking is a warp-uniform branch taken only by strips / rows that touch the
//    boundary; interior warps execute no mask instructions at all;
//  - global pointers for prefetch and store advance by one pitch per step.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace mgb {

enum { MODE_SWEEPS = 0, MODE_PRE = 1, MODE_POST = 2,
       MODE_POSTPRE = 3 };   // POST of one level visit fused with PRE of the next visit of the same level (see below)
constexpr bool mode_has_post(int m) { return m == MODE_POST || m == MODE_POSTPRE; }   // stage 0 adds the coarse correction
constexpr bool mode_has_pre(int m) { return m == MODE_PRE || m == MODE_POSTPRE; }     // residual + full weighting after stage NS

constexpr int kStreamWarps = 1;  // warps (independent work items) per CTA.  4 lock-stepped warps per CTA (bar.sync every 3 rows)
                                 // were measured slower and more erratic (profiles/r01_tune_stream.txt)
// ring slots per warp: NB blocks of 3 slots, NB a power of two.  12 slots prefetch D = 9 - NS rows ahead.  A 24-slot ring
// (D = 21 - NS, 8 resident warps per SM instead of 12) was measured on the B200 and is slower for every kernel
// (profiles/r02_kernel_experiments.md); the build-time switch stays for re-measuring on other parts.
#ifndef MGB_DEEP_RING_MIN_NS
#define MGB_DEEP_RING_MIN_NS 99   // kernels with NS >= this would use the 24-slot ring
#endif
constexpr int ring_slots(int ns) { return ns >= MGB_DEEP_RING_MIN_NS ? 24 : 12; }

// NORM (POST only): one more pipeline stage computes the residual of the OUTPUT iterate and accumulates its sum of
// squares (the convergence check of the tolerance loop, P:604-608 shape) -- no separate residual pass over the grid.
template <typename T, int NS, int MODE, bool NORM = false>
struct StreamCfg {
    static constexpr int V = Vec<T>::N;
    static constexpr bool HAS_POST = mode_has_post(MODE), HAS_PRE = mode_has_pre(MODE);
    static constexpr bool HAS_RES = HAS_PRE || NORM;                                    // residual stage after stage NS
    static_assert(!(NORM && HAS_PRE), "NORM is a POST / SWEEPS option");
    static constexpr int HL = NS + (HAS_PRE ? 2 : 0) + (NORM ? 1 : 0);                  // columns needed to the left
    static constexpr int HR = NS + (HAS_RES ? 1 : 0) + (HAS_POST ? 1 : 0);              // ... to the right
    static constexpr int HMAX = HL > HR ? HL : HR;
    static constexpr int HLANES = (HMAX + V - 1) / V;                                   // halo lanes per side
    static constexpr int OUTW = 32 * V - 2 * V * HLANES;                                // output columns per strip
    static constexpr int HT = NS + (HAS_PRE ? 2 : 0) + (NORM ? 1 : 0) + (HAS_POST ? 1 : 0);   // rows needed above
    static constexpr int HB = NS + (HAS_PRE ? 2 : 0) + (NORM ? 1 : 0);                  // rows needed below
    static constexpr int DEPTH = ring_slots(NS);
    static constexpr int NB = DEPTH / 3;                                                // ring blocks (power of two)
    static constexpr int D = DEPTH - NS - 3;                                            // prefetch distance (rows)
    static constexpr int NW = NS + (HAS_RES ? 1 : 0);                                   // register windows of u_s
    // slot: [u: 32 V][f: 32 V]; POST keeps the coarse rows in a second, half-rate ring
    static constexpr int SLOT_ELEMS = 32 * V * 2;
    static constexpr int CSLOT_ELEMS = 32 * (V / 2);
    static constexpr int WARP_ELEMS = DEPTH * SLOT_ELEMS + (HAS_POST ? DEPTH * CSLOT_ELEMS : 0);
    static constexpr size_t SMEM_BYTES = (size_t)kStreamWarps * WARP_ELEMS * sizeof(T);
    static_assert(D >= 3, "ring too shallow");
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
    int strips_pad;      // strips rounded up to a multiple of kStreamWarps
    int nitems;          // strips * chunks
    T c0, c1, w;
    T* fc;               // PRE: coarse right-hand side / zero guess
    T* uc;
    const T* ec;         // POST: coarse correction
    i64 pitch_c;
    int Nc;
    int crow_lo, crow_hi;  // coarse rows backed by storage
};

#ifdef MGB_EMU
// CPU emulation build (tests/host_emul): the copy happens at issue time, which is one of the orders the hardware allows
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) { if (valid) memcpy(smem, gmem, 16); else memset(smem, 0, 16); }
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) { if (valid) memcpy(smem, gmem, 8); else memset(smem, 0, 8); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
#else
// cp.async with the `ignore-src` predicate: when !valid nothing is read and zeros are written
// (maps 1:1 onto LDGSTS.ZFILL with a predicate; `gmem` must still be a mapped address).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, %2, 0;\n\tcp.async.cg.shared.global [%0], [%1], 16, p;\n\t}\n"
                 ::"r"(s), "l"(gmem), "r"((int)valid) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, %2, 0;\n\tcp.async.ca.shared.global [%0], [%1], 8, p;\n\t}\n"
                 ::"r"(s), "l"(gmem), "r"((int)valid) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

// ZG ("zero guess"): the input iterate is known to be identically zero (first visit of a coarse level, P:613):
// u is neither prefetched nor read, stage 0 is the constant 0.  Same arithmetic on the same values => same bits.
template <typename T, int NS, int MODE, bool RBGS, bool ZG = false, bool NORM = false>
struct Streamer {
    typedef StreamCfg<T, NS, MODE, NORM> C;
    static constexpr int V = C::V;
    static constexpr int H = V / 2;
    static constexpr unsigned FULL = 0xffffffffu;
    static constexpr int BLK = 3 * C::SLOT_ELEMS;    // elements per ring block (3 slots)
    static constexpr int CBLK = 3 * C::CSLOT_ELEMS;

    const StreamArgs<T>& a;
    T* ring;       // this lane's 16 bytes of slot 0 (u part)
    T* cring;      // POST: this lane's 8 bytes of coarse slot 0
    static constexpr int NB = C::NB;
    T* blk[NB];    // blk[j] = ring block (q + j) & (NB - 1) for the current outer iteration q
    T* cblk[NB];
    int c;         // first column of this lane
    int y0, y1;
    bool lane_ld, lane_ldc, lane_st;  // per-lane load / store flags
    typedef typename std::conditional<sizeof(T) == 8, unsigned long long, unsigned>::type MaskT;
    MaskT cm[V];   // all-ones for interior columns, 0 where the column must hold zero (Dirichlet ring / beyond the grid)
    const T* g_u;  // global prefetch pointers (row of the next prefetch, this lane's column)
    const T* g_f;
    T* g_o;        // output pointer of the row the last smoothing stage produces in this step
    const T* g_c;  // POST: coarse row of the next prefetch
    const T* safe_u;  // mapped addresses handed to ignored copies
    const T* safe_f;
    const T* safe_c;
    T W[C::NW > 0 ? C::NW : 1][3][V];
    T WL[C::NW > 0 ? C::NW : 1][3], WR[C::NW > 0 ? C::NW : 1][3];
    T R[3][V], RL[3];  // PRE: residual window (left neighbours only)
    double nacc;       // NORM: this lane's sum of squared residuals

    __device__ __forceinline__ Streamer(const StreamArgs<T>& a_) : a(a_) {}

    // ring slot of row (current step row - back), for phase PH: block pointer + constant
    template <int PH, int BACK>
    __device__ __forceinline__ T* rslot() const
    {
        constexpr int rel = PH - BACK;                       // slot index relative to this iteration's block
        constexpr int b = (rel >= 0) ? rel / 3 : -((-rel + 2) / 3);
        constexpr int s = rel - 3 * b;
        return blk[(b + NB) & (NB - 1)] + s * C::SLOT_ELEMS;
    }
    template <int PH, int BACK>
    __device__ __forceinline__ T* cslot() const
    {
        constexpr int rel = PH - BACK;
        constexpr int b = (rel >= 0) ? rel / 3 : -((-rel + 2) / 3);
        constexpr int s = rel - 3 * b;
        return cblk[(b + NB) & (NB - 1)] + s * C::CSLOT_ELEMS;
    }

    // prefetch row y into the slot that is REL slots ahead of this iteration's block start
    template <int REL>
    __device__ __forceinline__ void issue(int y)
    {
        constexpr int b = REL / 3, s = REL % 3;
        T* dst = blk[b & (NB - 1)] + s * C::SLOT_ELEMS;
        const bool v = lane_ld && (y >= a.row_lo) && (y < a.row_hi);
        if (!ZG) cp_async16(dst, v ? g_u : safe_u, v);
        cp_async16(dst + 32 * V, v ? g_f : safe_f, v);
        g_u += a.pitch;
        g_f += a.pitch;
        if (C::HAS_POST) {
            // coarse row ceil(y/2): a new one starts at every odd y; even rows re-read the previous one (L1 hit)
            const int ic = (y + 1) >> 1;
            const bool vc = lane_ldc && (ic >= a.crow_lo) && (ic < a.crow_hi) && (y >= 0);
            cp_async8(cblk[b & (NB - 1)] + s * C::CSLOT_ELEMS, vc ? g_c : safe_c, vc);
            if (!(y & 1)) g_c += a.pitch_c;   // ceil((y+1)/2) > ceil(y/2) exactly when y is even
        }
        cp_async_commit();
    }

    // `need` (warp-uniform): bit 0 = the consumer of this row reads the left neighbour lane's last value, bit 1 = the
    // right neighbour lane's first value.  Jacobi and the residual need both; a red-black stage needs exactly one (below).
    template <int NEW>
    __device__ __forceinline__ void put_row(int s, const T (&v)[V], int need = 3)
    {
#pragma unroll
        for (int k = 0; k < V; ++k) W[s][NEW][k] = v[k];
        if (need & 1) WL[s][NEW] = __shfl_up_sync(FULL, v[V - 1], 1);
        if (need & 2) WR[s][NEW] = __shfl_down_sync(FULL, v[0], 1);
    }

    // Red-black Gauss-Seidel: stage t (1-based) updates colour (t-1)&1 on row `row`.  Every lane's first column c is even
    // (strip origins and lane offsets are multiples of V), so the columns of that colour are the k with k % 2 == kk,
    // kk = (row + colour) & 1 -- the same for the whole warp.  A stage therefore computes V/2 points per lane, not V,
    // and needs only ONE neighbour-lane value: the left one when kk == 0 (k = 0 is updated), the right one when kk == 1.
    static __device__ __forceinline__ int rb_kk(int row, int t) { return (row + ((t - 1) & 1)) & 1; }
    static __device__ __forceinline__ int rb_need(int row, int t) { return rb_kk(row, t) ? 2 : 1; }

    // Dirichlet ring: only rows 0 / N (rows beyond them are zero by construction) and, on
    // strips touching the boundary, the columns flagged in cz[] ever need forcing.
    // Dirichlet columns: AND with a per-lane bit mask (all-ones in the interior; exact, NaN-safe, gives +0).
    // Dirichlet rows 0 / N are handled by a warp-uniform branch around each stage (rows beyond them are
    // zero by construction), so interior rows pay two logic ops per value and nothing else.
    __device__ __forceinline__ void mask_cols(T (&o)[V]) const
    {
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if constexpr (sizeof(T) == 8) o[k] = __longlong_as_double((long long)((unsigned long long)__double_as_longlong(o[k]) & cm[k]));
            else o[k] = __uint_as_float(__float_as_uint(o[k]) & cm[k]);
        }
    }
    __device__ __forceinline__ void mask_col(T& o, int k) const
    {
        if constexpr (sizeof(T) == 8) o = __longlong_as_double((long long)((unsigned long long)__double_as_longlong(o) & cm[k]));
        else o = __uint_as_float(__float_as_uint(o) & cm[k]);
    }
    __device__ __forceinline__ bool ring_row(int row) const { return (row <= 0) || (row >= a.N); }

    // one pipeline step: row y of the input arrives
    template <int PH>
    __device__ __forceinline__ void step(int y)
    {
        constexpr int NEW = PH, MID = (PH + 2) % 3, OLD = (PH + 1) % 3;
        T cur[V];
        // ---- stage 0: the incoming row (POST: plus the interpolated coarse correction) ----
        if (ZG) {
#pragma unroll
            for (int k = 0; k < V; ++k) cur[k] = (T)0;
        } else {
            ldv<T>(rslot<PH, 0>(), cur);
        }
        if (C::HAS_POST) {
            T ca[H + 1], cb[H + 1], e[V];
            const T* pa = cslot<PH, 0>();
#pragma unroll
            for (int k = 0; k < H; ++k) cb[k] = pa[k];
            cb[H] = __shfl_down_sync(FULL, cb[0], 1);
            if (y & 1) {  // (y parity is warp-uniform, so both branches are convergent)
                const T* pp = cslot<PH, 1>();
#pragma unroll
                for (int k = 0; k < H; ++k) ca[k] = pp[k];
                ca[H] = __shfl_down_sync(FULL, ca[0], 1);
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    e[2 * k] = (T)0.5 * (ca[k] + cb[k]);                                       // P:407
                    e[2 * k + 1] = (T)0.25 * (((ca[k] + cb[k]) + ca[k + 1]) + cb[k + 1]);      // P:419
                }
            } else {
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    e[2 * k] = cb[k];                                                          // P:401
                    e[2 * k + 1] = (T)0.5 * (cb[k] + cb[k + 1]);                               // P:413
                }
            }
#pragma unroll
            for (int k = 0; k < V; ++k) cur[k] = ring_row(y) ? (T)0 : (ZG ? e[k] : cur[k] + e[k]);   // P:623 (ZG: bare interpolation, P:645)
            mask_cols(cur);
        }
        put_row<NEW>(0, cur, (RBGS && NS >= 1) ? rb_need(y, 1) : 3);

        // ---- smoothing stages: stage s produces row y-s of u_s from the window of u_{s-1} ----
#pragma unroll
        for (int s = 1; s <= NS; ++s) {
            const int rs = y - s;
            T ff[V], o[V];
            if (s == 1) ldv<T>(rslot<PH, 1>() + 32 * V, ff);
            else if (s == 2) ldv<T>(rslot<PH, 2>() + 32 * V, ff);
            else if (s == 3) ldv<T>(rslot<PH, 3>() + 32 * V, ff);
            else ldv<T>(rslot<PH, 4>() + 32 * V, ff);
            if (ring_row(rs)) {   // warp-uniform, rare
#pragma unroll
                for (int k = 0; k < V; ++k) o[k] = (T)0;
            } else if constexpr (RBGS) {
#pragma unroll
                for (int k = 0; k < V; ++k) o[k] = W[s - 1][MID][k];     // the other colour is carried over
                if (rb_kk(rs, s) == 0) {                                 // warp-uniform (see rb_kk)
#pragma unroll
                    for (int k = 0; k < V; k += 2) {
                        const T l = (k == 0) ? WL[s - 1][MID] : W[s - 1][MID][k - 1];
                        o[k] = gs_pt<T>(ff[k], sigma4<T>(W[s - 1][OLD][k], W[s - 1][NEW][k], l, W[s - 1][MID][k + 1]));
                        mask_col(o[k], k);
                    }
                } else {
#pragma unroll
                    for (int k = 1; k < V; k += 2) {
                        const T r = (k == V - 1) ? WR[s - 1][MID] : W[s - 1][MID][k + 1];
                        o[k] = gs_pt<T>(ff[k], sigma4<T>(W[s - 1][OLD][k], W[s - 1][NEW][k], W[s - 1][MID][k - 1], r));
                        mask_col(o[k], k);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const T l = (k == 0) ? WL[s - 1][MID] : W[s - 1][MID][k - 1];
                    const T r = (k == V - 1) ? WR[s - 1][MID] : W[s - 1][MID][k + 1];
                    o[k] = jacobi_pt<T>(a.c0, a.c1, W[s - 1][MID][k], ff[k], sigma4<T>(W[s - 1][OLD][k], W[s - 1][NEW][k], l, r));
                }
                mask_cols(o);
            }
            if (s < C::NW) put_row<NEW>(s, o, (RBGS && s < NS) ? rb_need(rs, s + 1) : 3);
            if (s == NS) {
                if (lane_st && rs >= y0 && rs < y1) stv<T>(g_o, o);
                g_o += a.pitch;
            }
        }

        // ---- NORM: residual of the output iterate u_NS (row y-NS-1), squared and accumulated on the stored points ----
        if constexpr (NORM) {
            const int rr = y - NS - 1;
            if (!ring_row(rr) && lane_st && rr >= y0 && rr < y1) {
                T ff[V];
                ldv<T>(rslot<PH, NS + 1>() + 32 * V, ff);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const T l = (k == 0) ? WL[NS][MID] : W[NS][MID][k - 1];
                    const T r = (k == V - 1) ? WR[NS][MID] : W[NS][MID][k + 1];
                    T o = resid_pt<T>(W[NS][MID][k], ff[k], sigma4<T>(W[NS][OLD][k], W[NS][NEW][k], l, r));
                    mask_col(o, k);
                    nacc += (double)o * (double)o;
                }
            }
        }

        // ---- PRE: residual of u_NS (row y-NS-1) and full weighting (coarse row when that row is odd) ----
        if (C::HAS_PRE) {
            const int rr = y - NS - 1;
            T ff[V], o[V];
            ldv<T>(rslot<PH, NS + 1>() + 32 * V, ff);
            if (ring_row(rr)) {
#pragma unroll
                for (int k = 0; k < V; ++k) o[k] = (T)0;
            } else {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const T l = (k == 0) ? WL[NS][MID] : W[NS][MID][k - 1];
                    const T r = (k == V - 1) ? WR[NS][MID] : W[NS][MID][k + 1];
                    o[k] = resid_pt<T>(W[NS][MID][k], ff[k], sigma4<T>(W[NS][OLD][k], W[NS][NEW][k], l, r));
                }
                mask_cols(o);
            }
#pragma unroll
            for (int k = 0; k < V; ++k) R[NEW][k] = o[k];
            RL[NEW] = __shfl_up_sync(FULL, o[V - 1], 1);
            if (rr & 1) {
                const int yc = rr - 1;  // fine centre row 2I (MID of the residual window)
                T oc[H];
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const int k = 2 * j;
                    const T nw = (k == 0) ? RL[OLD] : R[OLD][k - 1];
                    const T wv = (k == 0) ? RL[MID] : R[MID][k - 1];
                    const T sw = (k == 0) ? RL[NEW] : R[NEW][k - 1];
                    oc[j] = fw_pt<T>(a.w, nw, R[OLD][k + 1], sw, R[NEW][k + 1], wv, R[MID][k + 1],
                                     R[OLD][k], R[NEW][k], R[MID][k]);
                }
                {   // coarse ring columns: J = 0 <=> fine column 0, J >= Nc <=> fine column >= N (same mask)
                    T tmp[V];
#pragma unroll
                    for (int k = 0; k < V; ++k) tmp[k] = (k & 1) ? (T)0 : oc[k >> 1];
                    mask_cols(tmp);
#pragma unroll
                    for (int j = 0; j < H; ++j) oc[j] = tmp[2 * j];
                }
                if (lane_st && yc >= y0 && yc < y1) {
                    const i64 offc = (i64)(yc >> 1) * a.pitch_c + (c >> 1);
                    if constexpr (H == 1) {
                        a.fc[offc] = oc[0];
                        if (a.uc) a.uc[offc] = (T)0;
                    } else {
                        *reinterpret_cast<float2*>(a.fc + offc) = make_float2((float)oc[0], (float)oc[H - 1]);
                        if (a.uc) *reinterpret_cast<float2*>(a.uc + offc) = make_float2(0.f, 0.f);
                    }
                }
            }
        }
    }

    __device__ __forceinline__ void set_blocks(int q)
    {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            blk[j] = ring + ((q + j) & (NB - 1)) * BLK;
            if (C::HAS_POST) cblk[j] = cring + ((q + j) & (NB - 1)) * CBLK;
        }
    }

    __device__ __forceinline__ void run(T* ring_base, int warp, int lane, int item)
    {
        const int chunk = item / a.strips_pad;
        const int strip = item - chunk * a.strips_pad;
        const bool dead = strip >= a.strips;   // padding warp (only when kStreamWarps > 1): neither loads nor stores
        const int X0 = strip * C::OUTW;
        c = X0 + V * lane;
        const int out_lo = (strip == 0) ? 0 : X0 + V * C::HLANES;
        const int out_hi = X0 + V * C::HLANES + C::OUTW;
        y0 = a.ya + chunk * a.ry;
        y1 = min(y0 + a.ry, a.yb);
        ring = ring_base + (size_t)warp * C::WARP_ELEMS + lane * V;
        cring = ring_base + (size_t)warp * C::WARP_ELEMS + C::DEPTH * C::SLOT_ELEMS + lane * H;
        const int ylo = y0 - C::HT, yhi = y1 - 1 + C::HB;
        lane_ld = (c < a.pitch) && !dead;
        lane_ldc = ((c >> 1) < a.pitch_c) && !dead;
        lane_st = (c >= out_lo) && (c < out_hi) && (c < a.N) && !dead;
#pragma unroll
        for (int k = 0; k < V; ++k) cm[k] = ((c + k < 1) || (c + k >= a.N)) ? (MaskT)0 : ~(MaskT)0;

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
        nacc = 0.0;

        g_u = a.u_in + (i64)ylo * a.pitch + c;
        g_f = a.f + (i64)ylo * a.pitch + c;
        g_o = a.u_out + (i64)(ylo - NS) * a.pitch + c;   // row produced by stage NS in the first step
        safe_u = a.u_in + (i64)a.row_lo * a.pitch;
        safe_f = a.f + (i64)a.row_lo * a.pitch;
        safe_c = a.ec;
        g_c = a.ec;
        if (C::HAS_POST) {
            safe_c = a.ec + (i64)a.crow_lo * a.pitch_c;
            g_c = a.ec + (i64)((ylo + 1) >> 1) * a.pitch_c + (c >> 1);
        }

        // prologue: rows ylo .. ylo+D-1 into slots 0 .. D-1
        set_blocks(0);
        issue_prologue<0>(ylo);
        int q = 0;
        for (int y = ylo; y <= yhi; y += 3) {
            set_blocks(q);
            issue<C::D>(y + C::D);
            wait_row<0>();
            step<0>(y);
            issue<C::D + 1>(y + 1 + C::D);
            wait_row<1>();
            step<1>(y + 1);
            issue<C::D + 2>(y + 2 + C::D);
            wait_row<2>();
            step<2>(y + 2);
            q = (q + 1) & (NB - 1);
        }
        cp_async_wait<0>();
    }

    // the row that step<PH> is about to consume has landed
    template <int PH>
    __device__ __forceinline__ void wait_row()
    {
        cp_async_wait<C::D>();
    }

    template <int I>
    __device__ __forceinline__ void issue_prologue(int ylo)
    {
        if constexpr (I < C::D) {
            issue<I>(ylo + I);
            issue_prologue<I + 1>(ylo);
        }
    }
};

// launch bound = the 12 resident warps per SM the host caps these kernels at (FusedKnobs::occ): 170 registers per
// thread instead of the 128 a bound of 16 allows, which the red-black PRE kernel (NS = 4) needs to stay spill-free
constexpr int kStreamMinCtas = 12;
template <typename T, int NS, int MODE, bool RBGS>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamMinCtas / kStreamWarps)
k_stream(const StreamArgs<T> a)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kStreamWarps + warp;
    Streamer<T, NS, MODE, RBGS> st(a);
    st.run(reinterpret_cast<T*>(stream_smem), warp, threadIdx.x & 31, item);
}

// POST (or SWEEPS) with the residual norm of the output folded in: one partial sum of squares per work item (warp),
// reduced by warp shuffles in a fixed order; k_sum_partials adds the partials in index order
template <typename T, int NS, int MODE, bool RBGS>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamMinCtas / kStreamWarps)
k_stream_norm(const StreamArgs<T> a, double* __restrict__ partials)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kStreamWarps + warp;
    Streamer<T, NS, MODE, RBGS, false, true> st(a);
    st.run(reinterpret_cast<T*>(stream_smem), warp, threadIdx.x & 31, item);
    double acc = st.nacc;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) partials[item] = acc;
}

// POSTPRE (visit chains) needs ~142 registers at NS = 4: its own entry point with a launch bound
// of 12 resident CTAs per SM (the occupancy the host caps the streaming kernels at anyway) instead of 16 => no spills
constexpr int kStreamChainMinCtas = 12;
template <typename T, int NS, bool RBGS>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamChainMinCtas / kStreamWarps)
k_stream_chain(const StreamArgs<T> a)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kStreamWarps + warp;
    Streamer<T, NS, MODE_POSTPRE, RBGS> st(a);
    st.run(reinterpret_cast<T*>(stream_smem), warp, threadIdx.x & 31, item);
}

// fullmultigrid's entry into a level (P:645-646): the bare interpolation of the coarse solution as initial guess fused with
// the first cycle's PRE -- POSTPRE with "the incoming iterate is zero and is not read", stage 0 = P u_c
template <typename T, int NS, bool RBGS>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamChainMinCtas / kStreamWarps)
k_stream_fmg_entry(const StreamArgs<T> a)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kStreamWarps + warp;
    Streamer<T, NS, MODE_POSTPRE, RBGS, true> st(a);
    st.run(reinterpret_cast<T*>(stream_smem), warp, threadIdx.x & 31, item);
}

// zero-guess variant of the PRE kernel (zero-guess chain): its own entry point, so that k_stream keeps its code
template <typename T, int NS, bool RBGS>
__global__ void __launch_bounds__(kStreamWarps * 32, kStreamMinCtas / kStreamWarps)
k_stream_pre_zg(const StreamArgs<T> a)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char stream_smem[];
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * kStreamWarps + warp;
    Streamer<T, NS, MODE_PRE, RBGS, true> st(a);
    st.run(reinterpret_cast<T*>(stream_smem), warp, threadIdx.x & 31, item);
}

}  // namespace mgb
