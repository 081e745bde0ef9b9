// oracle/stub/oneapi/mkl.hpp — TEST INFRASTRUCTURE.  Host stand-in for the oneMKL calls of
// /root/reference/Poissons_SYCL.cpp (sparse::gemv / init_matrix_handle / set_csr_data,
// blas::column_major::scal, vm::add, vm::sub) with their published semantics:
// y = alpha*A*x + beta*y (CSR, row order), x *= alpha, y = a + b, y = a - b, all in float.
// Counts calls so that tests can pin the reference's cycle structure.
#pragma once
#include <cstdint>
#include <vector>

#include <CL/sycl.hpp>

namespace refstub {
struct counters { long gemv = 0, scal = 0, add = 0, sub = 0; long long gemv_rows = 0; };
inline counters& C() { static counters c; return c; }
}  // namespace refstub

namespace oneapi { namespace mkl {

enum class transpose { nontrans, trans };
enum class index_base { zero, one };

namespace sparse {
struct matrix_handle {
    std::int32_t nrows = 0, ncols = 0;
    const std::int32_t* row_ptr = nullptr;
    const std::int32_t* col_ind = nullptr;
    const float* val = nullptr;
};
typedef matrix_handle* matrix_handle_t;
inline void init_matrix_handle(matrix_handle_t* h) { *h = new matrix_handle(); }
inline void set_csr_data(matrix_handle_t h, std::int32_t nrows, std::int32_t ncols, index_base, std::int32_t* row_ptr,
                         std::int32_t* col_ind, float* val)
{
    h->nrows = nrows; h->ncols = ncols; h->row_ptr = row_ptr; h->col_ind = col_ind; h->val = val;
}
inline sycl::event gemv(sycl::queue&, transpose, float alpha, matrix_handle_t A, const float* x, float beta, float* y,
                        const std::vector<sycl::event>& = {})
{
    refstub::C().gemv++;
    refstub::C().gemv_rows += A->nrows;
    for (std::int32_t r = 0; r < A->nrows; ++r) {
        float s = 0.0f;
        for (std::int32_t k = A->row_ptr[r]; k < A->row_ptr[r + 1]; ++k) s += A->val[k] * x[A->col_ind[k]];
        y[r] = (beta == 0.0f) ? alpha * s : alpha * s + beta * y[r];
    }
    return sycl::event();
}
}  // namespace sparse

namespace blas { namespace column_major {
inline sycl::event scal(sycl::queue&, std::int64_t n, float alpha, float* x, std::int64_t incx,
                        const std::vector<sycl::event>& = {})
{
    refstub::C().scal++;
    for (std::int64_t i = 0; i < n; ++i) x[i * incx] *= alpha;
    return sycl::event();
}
}}  // namespace blas::column_major

namespace vm {
inline sycl::event add(sycl::queue&, std::int64_t n, const float* a, const float* b, float* y,
                       const std::vector<sycl::event>& = {})
{
    refstub::C().add++;
    for (std::int64_t i = 0; i < n; ++i) y[i] = a[i] + b[i];
    return sycl::event();
}
inline sycl::event sub(sycl::queue&, std::int64_t n, const float* a, const float* b, float* y,
                       const std::vector<sycl::event>& = {})
{
    refstub::C().sub++;
    for (std::int64_t i = 0; i < n; ++i) y[i] = a[i] - b[i];
    return sycl::event();
}
}  // namespace vm

}}  // namespace oneapi::mkl
