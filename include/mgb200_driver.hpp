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
//
// Errors: the reference has none; here every failing call throws
// std::runtime_error carrying mg_last_error().  T is float (as P) or double (as M).
#pragma once

#include <cmath>
#include <stdexcept>
#include <string>
#include <type_traits>
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

// P:125-147.  Like the reference, v is smoothed in place AND returned by value (P:146).
template <typename T>
std::vector<T> jacobirelaxation(queue<T>& q, level a_h, std::vector<T>& v, std::vector<T>& fh, const int& mu)
{
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

}  // namespace mgb200
