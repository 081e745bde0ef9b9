// oracle/stub/CL/sycl.hpp — TEST INFRASTRUCTURE.  A host-only stand-in for the ~10 SYCL
// symbols /root/reference/Poissons_SYCL.cpp uses, so that the reference translation unit
// compiles with plain g++ (oracle/Makefile target `ref`).  Semantics: everything runs
// synchronously on the calling thread; parallel_for is a serial loop.  Not a SYCL
// implementation and not part of the product.
#pragma once
#include <cstddef>
#include <functional>
#include <vector>

namespace sycl {

struct event {
    void wait() {}
};

template <int D> struct range {
    std::size_t v[D];
    range(std::size_t a) : v{a} { static_assert(D == 1, ""); }
    range(std::size_t a, std::size_t b) : v{a, b} { static_assert(D == 2, ""); }
    std::size_t operator[](int i) const { return v[i]; }
};
template <int D> struct id {
    std::size_t v[D];
    std::size_t operator[](int i) const { return v[i]; }
};

template <typename T, int D> struct buffer {
    T* ptr;
    range<D> r;
    buffer(T* p, range<D> rr) : ptr(p), r(rr) {}
};

struct handler {
    template <typename F> void parallel_for(range<1> r, F f)
    {
        for (std::size_t i = 0; i < r[0]; ++i) f(id<1>{{i}});
    }
    template <typename F> void parallel_for(range<2> r, F f)
    {
        for (std::size_t i = 0; i < r[0]; ++i)
            for (std::size_t j = 0; j < r[1]; ++j) f(id<2>{{i, j}});
    }
};

template <typename T, int D> struct accessor;
template <typename T> struct accessor<T, 2> {
    T* ptr;
    std::size_t cols;
    accessor(buffer<T, 2>& b, handler&) : ptr(b.ptr), cols(b.r[1]) {}
    T* operator[](std::size_t i) const { return ptr + i * cols; }
};
template <typename T, int D> accessor(buffer<T, D>&, handler&) -> accessor<T, D>;

struct queue {
    template <typename F> event submit(F f)
    {
        handler h;
        f(h);
        return event();
    }
    void wait() {}
};

}  // namespace sycl

namespace cl { namespace sycl = ::sycl; }
