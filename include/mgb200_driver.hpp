// include/mgb200_driver.hpp — the reference's C++ function surface over the C ABI.
//
// Header-only.  Same names, argument meaning and return-by-value behaviour as the free
// functions of /root/reference/Poissons_SYCL.cpp ("P:line"), with `mgb200::queue&` standing
// where `cl::sycl::queue&` stood and `mgb200::level` standing where
// `matrix_elements_for_jacobi&` / `matrix_handle_t` stood (the reference's per-level operator
// handle, P:24-30: there is no assembled matrix here, a level is just its index).
//
//   reference (P)                                           here
//   jacobirelaxation(q, a_lu, a_size, v, fh, mu)    P:125   jacobirelaxation(q, a_h, v, fh, mu)
//   restriction2d(vec_h)                            P:531   restriction2d(q, vec_h)
//   interpolation2d(vec_2h)                         P:337   interpolation2d(q, vec_2h)
//   vcyclemultigrid(q, a_h, vec_h, f_h)             P:575   vcyclemultigrid(q, a_h, vec_h, f_h)
//   fullmultigrid(q, a_h, f_h)                      P:629   fullmultigrid(q, a_h, f_h)
//   globalforcefunction()                           P:283   globalforcefunction(q)
//   main() level loop                               P:661   queue(finest_level, coarsest_level)
//   v2 sketch (Multigrid_functions.cpp, M):                 ProblemVar<T>, fullmultigrid(q, obj, f_h, level),
//     ProblemVar M:16, fullmultigrid M:175, multigrid_solver M:193   multigrid_solver(obj); sampled f / Dirichlet g:
//                                                           globalforcefunction(q, f, g)
//
// Errors: the reference has none; here every failing call throws
// std::runtime_error carrying mg_last_error().  T is float (as P) or double (as M).
#pragma once

#include <cmath>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "mgb200.h"

namespace mgb200 {

// the reference's compile-time globals (P:17-22), kept as defaults of the context
struct parameters {
    int finest_level = 10;   // P:17
    int coarsest_level = 7;  // P:18
    int mu0 = 30;            // P:20  (fullmultigrid runs mu0+1 cycles per level, P:635/P:646)
    int mu1 = 10;            // P:21
    int mu2 = 10;            // P:22
    double omega = 2.0 / 3.0;  // P:127
    double f = 4.0;            // P:123
    int gamma = 1;             // 1 = V-cycle (reference); 2 = W-cycle
    int smoother = MG_SMOOTH_JACOBI;
    int coarse_solver = MG_COARSE_SWEEPS;   // P:583-587; MG_COARSE_EXACT = the second version's direct_solver (M:63-72, M:136-139)
};

struct level { int index; };  // stands in for matrix_elements_for_jacobi (P:24-30)

template <typename T>
class queue {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "float or double");

public:
    explicit queue(const parameters& p = parameters(), int device = -1) : par(p)
    {
        mg_config cfg;
        mg_config_default(&cfg);
        cfg.finest_level = p.finest_level;
        cfg.coarsest_level = p.coarsest_level;
        cfg.dtype = std::is_same<T, double>::value ? MG_F64 : MG_F32;
        cfg.smoother = p.smoother;
        cfg.omega = p.omega;
        cfg.coarse_solver = p.coarse_solver;
        cfg.device = device;
        if (mg_create(&ctx, &cfg) != MG_OK) throw std::runtime_error(std::string("mg_create: ") + mg_last_error(nullptr));
    }
    ~queue() { if (ctx) mg_destroy(ctx); }
    queue(const queue&) = delete;
    queue& operator=(const queue&) = delete;

    mg_ctx* handle() const { return ctx; }
    void check(int rc, const char* what) const
    {
        if (rc != MG_OK) throw std::runtime_error(std::string(what) + ": " + mg_last_error(ctx));
    }
    // jacobi_matrices[level - coarsest_level] (P:33): the operator "handle" of a level
    level operator[](int lvl) const { return level{lvl}; }
    level finest() const { return level{par.finest_level}; }
    void wait() const { check(mg_sync(ctx), "mg_sync"); }  // event.wait()

    parameters par;

private:
    mg_ctx* ctx = nullptr;
};

// the reference infers the level from the vector length: int(log2(sqrt(size)+1)) (P:583)
inline int level_of_size(std::size_t size)
{
    const int l = mg_level_of_size(size);
    if (l < 0) throw std::runtime_error("vector length is not (2^L-1)^2");
    return l;
}

// the C ABI takes bare pointers: the length of every vector handed over is checked here against its level
inline void require_level_size(std::size_t size, int lvl, const char* what)
{
    const long long n = mg_level_side(lvl);
    if (n < 0 || size != (std::size_t)n * (std::size_t)n)
        throw std::runtime_error(std::string(what) + ": vector length " + std::to_string(size) + " is not (2^" +
                                 std::to_string(lvl) + "-1)^2");
}

// P:125-147.  Like the reference, v is smoothed in place AND returned by value (P:146).
template <typename T>
std::vector<T> jacobirelaxation(queue<T>& q, level a_h, std::vector<T>& v, std::vector<T>& fh, const int& mu)
{
    require_level_size(v.size(), a_h.index, "jacobirelaxation: v");
    require_level_size(fh.size(), a_h.index, "jacobirelaxation: fh");
    q.check(mg_host_jacobirelaxation(q.handle(), a_h.index, v.data(), fh.data(), mu), "jacobirelaxation");
    return v;
}

// P:531-546
template <typename T>
std::vector<T> restriction2d(queue<T>& q, std::vector<T>& vec_h)
{
    const int fine = level_of_size(vec_h.size());
    const std::size_t m = (std::size_t)mg_level_side(fine - 1);
    std::vector<T> vec_2h(m * m, 0);
    q.check(mg_host_restriction2d(q.handle(), fine, vec_h.data(), vec_2h.data()), "restriction2d");
    return vec_2h;
}

// P:337-425
template <typename T>
std::vector<T> interpolation2d(queue<T>& q, std::vector<T>& vec_2h)
{
    const int coarse = level_of_size(vec_2h.size());
    const std::size_t n = (std::size_t)mg_level_side(coarse + 1);
    std::vector<T> vec_h(n * n, 0);
    q.check(mg_host_interpolation2d(q.handle(), coarse + 1, vec_2h.data(), vec_h.data()), "interpolation2d");
    return vec_h;
}

// P:575-627: pre-smooth mu1, recurse, correct, post-smooth mu2.  Returns the new iterate.
template <typename T>
std::vector<T> vcyclemultigrid(queue<T>& q, level a_h, std::vector<T>& vec_h, std::vector<T>& f_h)
{
    require_level_size(vec_h.size(), a_h.index, "vcyclemultigrid: vec_h");
    require_level_size(f_h.size(), a_h.index, "vcyclemultigrid: f_h");
    std::vector<T> out(vec_h);
    q.check(mg_host_vcyclemultigrid(q.handle(), a_h.index, out.data(), f_h.data(), q.par.mu1, q.par.mu2, q.par.gamma),
            "vcyclemultigrid");
    return out;
}

// P:629-650: mu0+1 cycles per level on the way up.
template <typename T>
std::vector<T> fullmultigrid(queue<T>& q, level a_h, std::vector<T>& f_h)
{
    if (a_h.index != q.par.finest_level) throw std::runtime_error("fullmultigrid starts on the finest level");
    require_level_size(f_h.size(), a_h.index, "fullmultigrid: f_h");
    std::vector<T> out(f_h.size(), 0);
    q.check(mg_host_fullmultigrid(q.handle(), f_h.data(), out.data(), q.par.mu0 + 1, q.par.mu1, q.par.mu2), "fullmultigrid");
    return out;
}

// P:283-335: lumped P1 load vector of the finest level, b = f h^2.
template <typename T>
std::vector<T> globalforcefunction(queue<T>& q)
{
    const std::size_t n = (std::size_t)mg_level_side(q.par.finest_level);
    std::vector<T> b(n * n, 0);
    q.check(mg_force_constant(q.handle(), q.par.f), "globalforcefunction");
    q.check(mg_get_rhs_host(q.handle(), q.par.finest_level, b.data()), "globalforcefunction");
    return b;
}

// ---------------------------------------------------------------------------------------------
// globalforcefunction generalised from the constant f = 4 with a zero ring (P:123,
// P:283-335) to a sampled f(x, y) and Dirichlet data g(x, y).  b_i = f(x_i, y_i) h^2 (the lumped P1
// load of P:175-186); a boundary neighbour contributes +g to its interior node's row, so the device
// keeps the zero ring and the kernels do not change.  Host-side set-up, like the reference's.
// ---------------------------------------------------------------------------------------------
template <typename T, typename F, typename G>
std::vector<T> globalforcefunction(queue<T>& q, F f, G g)
{
    const int N = 1 << q.par.finest_level, n = N - 1;
    const double h = 1.0 / N;
    std::vector<T> b((std::size_t)n * n);
    for (int row = 1; row <= n; ++row)          // row <-> y, P:227-228
        for (int col = 1; col <= n; ++col) {
            const double x = col * h, y = row * h;
            double v = (double)f(x, y) * h * h;
            if (row == 1) v += (double)g(x, 0.0);
            if (row == n) v += (double)g(x, 1.0);
            if (col == 1) v += (double)g(0.0, y);
            if (col == n) v += (double)g(1.0, y);
            b[(std::size_t)(row - 1) * n + (col - 1)] = (T)v;
        }
    return b;
}

// ---------------------------------------------------------------------------------------------
// the call shape of the reference's second sketch (Multigrid_functions.cpp, "M:"):
// a problem object with a per-level load-vector dictionary (M:16-26), explicit level arguments, and
// multigrid_solver(obj) (M:193-197).  Operators are the structured-grid ones of libmgb200; like the second
// version, the coarsest level is solved directly (M:63-72, M:136-139: MG_COARSE_EXACT) -- set
// par.coarse_solver = MG_COARSE_SWEEPS for the first version's nu1+nu2 sweeps there (P:583-587).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct ProblemVar {
    parameters par = v2_parameters();
    std::unordered_map<int, std::vector<T>> b_dict;   // M:22: load vector per level; missing levels are restricted (P:641)

    static parameters v2_parameters()
    {
        parameters p;
        p.finest_level = 5;    // M:45
        p.coarsest_level = 1;  // M:44 has 0: no unknowns on a structured grid
        p.mu0 = 2;             // M:46
        p.mu1 = 1;             // M:47
        p.mu2 = 1;             // M:48
        p.coarse_solver = MG_COARSE_EXACT;   // M:136-139
        return p;              // omega: M:49 `4 / 5` is integer 0 (erratum); P:127's 2/3 is kept
    }
};

// M:75-96: in place, void, explicit level.  (M:87 hard-codes 10 sweeps; here obj.par.mu1.)
template <typename T>
void jacobirelaxation(queue<T>& q, std::vector<T>& vec, std::vector<T>& b, ProblemVar<T>& obj, int current_level)
{
    require_level_size(vec.size(), current_level, "jacobirelaxation: vec");
    require_level_size(b.size(), current_level, "jacobirelaxation: b");
    q.check(mg_host_jacobirelaxation(q.handle(), current_level, vec.data(), b.data(), obj.par.mu1), "jacobirelaxation");
}

// M:132-173
template <typename T>
std::vector<T> vcyclemultigrid(queue<T>& q, ProblemVar<T>& obj, std::vector<T>& vec_h, std::vector<T>& f_h, int current_level)
{
    require_level_size(vec_h.size(), current_level, "vcyclemultigrid: vec_h");
    require_level_size(f_h.size(), current_level, "vcyclemultigrid: f_h");
    std::vector<T> out(vec_h);
    q.check(mg_host_vcyclemultigrid(q.handle(), current_level, out.data(), f_h.data(), obj.par.mu1, obj.par.mu2, obj.par.gamma),
            "vcyclemultigrid");
    return out;
}

// M:175-191: the coarser problem first (its load vector from b_dict, M:183), interpolate (M:185), mu0+1 cycles (M:186-188)
template <typename T>
std::vector<T> fullmultigrid(queue<T>& q, ProblemVar<T>& obj, std::vector<T>& f_h, int current_level)
{
    mg_ctx* c = q.handle();
    const int lo = obj.par.coarsest_level;
    require_level_size(f_h.size(), current_level, "fullmultigrid: f_h");
    q.check(mg_set_rhs_host(c, current_level, f_h.data()), "fullmultigrid: f_h");
    for (int l = current_level - 1; l >= lo; --l) {
        auto it = obj.b_dict.find(l);
        if (it != obj.b_dict.end()) {
            require_level_size(it->second.size(), l, "fullmultigrid: b_dict");
            q.check(mg_set_rhs_host(c, l, it->second.data()), "fullmultigrid: b_dict");
        } else {
            q.check(mg_restrict_rhs(c, l + 1), "fullmultigrid: restrict");
        }
    }
    q.check(mg_zero_u(c, lo), "fullmultigrid");                                              // M:176
    // the loop `for i <= mu0: vec_h = vcyclemultigrid(...)` (M:186-188; P:646-648) as ONE call per level: same bits as mu0+1
    // mg_cycle calls, and the library may keep the iterate on chip between consecutive cycles (visit chains)
    q.check(mg_cycles(c, lo, obj.par.mu1, obj.par.mu2, 1, obj.par.mu0 + 1), "fullmultigrid");
    for (int l = lo + 1; l <= current_level; ++l) {
        q.check(mg_prolong_set(c, l), "fullmultigrid: interpolation");                       // M:185
        q.check(mg_cycles(c, l, obj.par.mu1, obj.par.mu2, 1, obj.par.mu0 + 1), "fullmultigrid");
    }
    const std::size_t n = (std::size_t)mg_level_side(current_level);
    std::vector<T> vec_h(n * n, 0);
    q.check(mg_get_u_host(c, current_level, vec_h.data()), "fullmultigrid: result");
    return vec_h;
}

// M:193-197
template <typename T>
std::vector<T> multigrid_solver(ProblemVar<T>& obj)
{
    queue<T> q(obj.par);
    auto it = obj.b_dict.find(obj.par.finest_level);
    if (it == obj.b_dict.end()) throw std::runtime_error("multigrid_solver: b_dict has no finest-level load vector");
    return fullmultigrid(q, obj, it->second, obj.par.finest_level);
}

}  // namespace mgb200
