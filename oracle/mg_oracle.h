/* oracle/mg_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the geometric-multigrid Poisson hot path: a plain-C restatement
 * of the reference's algorithm (/root/reference/Poissons_SYCL.cpp, "P:line"),
 * intended semantics.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (libmgb200.so) never links, loads or calls it.
 *
 * PARITY PINS.  The reference ships no tests, golden vectors or fixtures
 *  and its arithmetic lives in Intel oneMKL 2021.1.1
 * (sparse::gemv, blas::scal, vm::add/sub; not vendored, V:63), so nothing in the
 * reference pins results at the oneMKL boundary.  What IS pinned:
 *   - interpolation2d, globalforcefunction, the cycle call structure and the
 *     as-written end result are checked against the reference's OWN source
 *     compiled here against a stub oneMKL/SYCL header (oracle/_ref, built by
 *     oracle/Makefile; fixtures in tests/golden/ref_*.json);
 *   - restriction2d is pinned through the adjoint identity P = 4 R^T against the
 *     pinned interpolation (the reference's own weight `(1/16)` evaluates to 0, E2);
 *   - smoother / residual / whole cycles are pinned by the known answers of
 *     SURVEY.md Appendix C (exact rationals C.4, convergence histories C.3).
 * Smoother + residual arithmetic therefore remains "parity unpinned by the
 * reference's own tests"; DESIGN.md says so too.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int coarsest_level;     /* P:18 (reference 7 = finest-3; benchmark default 1)          */
    int nu1, nu2;           /* P:21-22 mu1, mu2 (reference 10; benchmark 2)                 */
    int gamma;              /* 1 = V (reference), 2 = W                                    */
    int smoother;           /* 0 = weighted Jacobi (P:125-147), 1 = red-black Gauss-Seidel */
    double omega;           /* P:127, 2/3                                                  */
    double restrict_weight; /* 0.25 intended (E2+E4 repaired); 1/16 literal FD weight      */
    int nthreads;           /* OpenMP threads used by every loop                           */
    int coarse_exact;       /* 0: coarsest level = nu1+nu2 sweeps (P:583-587); 1: exact solve, no
                               smoothing there (direct_solver, Multigrid_functions.cpp M:63-72 / M:136-139) */
} mgo_params;

void mgo_params_default(mgo_params* p);
int mgo_max_threads(void);

#define MGO_DECL(T, S)                                                                               \
    void mgo_jacobi_constants##S(double omega, T* c0, T* c1);                                        \
    void mgo_jacobirelaxation##S(T* v, const T* fh, int n, int mu, double omega, int nthreads);      \
    void mgo_rbgs##S(T* v, const T* fh, int n, int mu, int nthreads);                                \
    void mgo_rbgs_half##S(T* v, const T* fh, int n, int colour, int nthreads);                       \
    void mgo_residual##S(const T* v, const T* fh, T* r, int n, int nthreads);                        \
    double mgo_sumsq##S(const T* r, int n);                                                          \
    void mgo_restriction2d##S(const T* vec_h, int nh, T* vec_2h, double w, int nthreads);            \
    void mgo_interpolation2d##S(const T* vec_2h, int m, T* vec_h, int nthreads);                     \
    void mgo_prolong_correct##S(const T* vec_2h, int m, T* vec_h, int nthreads);                     \
    void mgo_globalforcefunction##S(T* out, int level, double f);                                    \
    void mgo_coarse_exact##S(T* u, const T* f, int n);                                               \
    void mgo_vcyclemultigrid##S(T* vec_h, const T* f_h, int level, const mgo_params* p);             \
    void mgo_fullmultigrid##S(T* vec_h, const T* f_h, int level, int cycles, const mgo_params* p);   \
    int mgo_solve##S(T* vec_h, const T* f_h, int level, double rtol, int max_cycles,                 \
                     double* history, const mgo_params* p);                                          \
    void* mgo_csr_build##S(int level);                                                               \
    void mgo_csr_free##S(void* h);                                                                   \
    void mgo_csr_jacobirelaxation##S(void* h, T* v, const T* fh, int mu, double omega, int nthreads);\
    void mgo_csr_residual##S(void* h, const T* v, const T* fh, T* r, int nthreads);                  \
    void mgo_csr_vcyclemultigrid##S(void** csr, T* vec_h, const T* f_h, int level, const mgo_params* p);

MGO_DECL(double, _f64)
MGO_DECL(float, _f32)

#ifdef __cplusplus
}
#endif
#endif
