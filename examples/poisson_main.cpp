// examples/poisson_main.cpp — the reference's main() (P:658-731) on top of libmgb200:
// build the level table, form the load vector, run full multigrid, report.
// Unlike the reference (which prints only the vector length, P:728), this also prints the
// residual history of a tolerance-controlled solve.
//
//   g++ -O2 -std=c++17 -Iinclude examples/poisson_main.cpp -o poisson_main \
//       -L multigrid_nikhil_c-_b200/lib -lmgb200 -Wl,-rpath,$PWD/multigrid_nikhil_c-_b200/lib
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "mgb200_driver.hpp"

int main(int argc, char** argv)
{
    mgb200::parameters par;             // P:17-22 defaults: levels 10..7, mu0=30, mu1=mu2=10
    if (argc > 1) par.finest_level = std::atoi(argv[1]);
    if (argc > 2) par.coarsest_level = std::atoi(argv[2]);
    if (argc > 3) { par.mu0 = 0; par.mu1 = par.mu2 = std::atoi(argv[3]); }
    try {
        mgb200::queue<double> q(par);                                   // P:659 + level loop P:661-690
        std::vector<double> f_global = mgb200::globalforcefunction(q);  // P:725
        std::vector<double> solution_finest = mgb200::fullmultigrid(q, q.finest(), f_global);  // P:727
        std::cout << "Size of finest level solution is " << solution_finest.size() << "\n";    // P:728
        double umax = 0;
        for (double v : solution_finest) umax = v > umax ? v : umax;
        std::printf("max u = %.12f (analytic u(1/2,1/2) = 0.29468541 for -Lap u = 4 on the unit square)\n", umax);

        // tolerance-controlled V(2,2) solve from zero on the same right-hand side
        int cycles = 0;
        double relres = 0;
        std::vector<double> hist(61);
        q.check(mg_zero_u(q.handle(), par.finest_level), "mg_zero_u");
        q.check(mg_solve(q.handle(), 1e-8, 60, 2, 2, 1, &cycles, &relres, hist.data()), "mg_solve");
        std::printf("V(2,2) solve: %d cycles, ||r||/||r0|| = %.3e\n", cycles, relres);
        for (int k = 1; k <= cycles; ++k) std::printf("  cycle %2d  ||r|| = %.6e  factor %.4f\n", k, hist[k], hist[k] / hist[k - 1]);
        std::cout << "Program Running Correctly ";                      // P:729
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
