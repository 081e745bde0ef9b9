// examples/problemvar_main.cpp — the reference's second call shape (Multigrid_functions.cpp M:193-197:
// multigrid_solver(ProblemVar&)) plus a Dirichlet problem with a sampled right-hand side.
//   -Lap u = -4 on the unit square, u = x^2 + y^2 on the boundary  =>  u = x^2 + y^2 (exact for the 5-point stencil)
//
//   g++ -O2 -std=c++17 -Iinclude examples/problemvar_main.cpp -o problemvar_main \
//       -L multigrid_nikhil_c-_b200/lib -lmgb200 -Wl,-rpath,$PWD/multigrid_nikhil_c-_b200/lib
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mgb200_driver.hpp"

int main(int argc, char** argv)
{
    try {
        mgb200::ProblemVar<double> obj;
        obj.par.finest_level = argc > 1 ? std::atoi(argv[1]) : 7;
        obj.par.mu0 = argc > 2 ? std::atoi(argv[2]) : 24;  // mu0+1 V(2,2) cycles per level
        obj.par.mu1 = obj.par.mu2 = 2;
        auto f = [](double, double) { return -4.0; };
        auto g = [](double x, double y) { return x * x + y * y; };
        {
            mgb200::queue<double> q(obj.par);
            obj.b_dict[obj.par.finest_level] = mgb200::globalforcefunction(q, f, g);
        }
        std::vector<double> u = mgb200::multigrid_solver(obj);
        const int N = 1 << obj.par.finest_level, n = N - 1;
        double err = 0;
        for (int row = 1; row <= n; ++row)
            for (int col = 1; col <= n; ++col)
                err = std::fmax(err, std::fabs(u[(std::size_t)(row - 1) * n + (col - 1)] - g(col / (double)N, row / (double)N)));
        std::printf("multigrid_solver: %zu unknowns, max |u - (x^2+y^2)| = %.3e\n", u.size(), err);
        return err < 1e-9 ? 0 : 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
