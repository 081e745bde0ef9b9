/* oracle/mg_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mg_oracle.h).
 * Build: gcc -O3 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile). */
#define _GNU_SOURCE   /* M_PI */
#include "mg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void mgo_params_default(mgo_params* p)
{
    p->coarsest_level = 1;
    p->nu1 = 2;
    p->nu2 = 2;
    p->gamma = 1;
    p->smoother = 0;
    p->omega = 2.0 / 3.0; /* P:127 */
    p->restrict_weight = 0.25;
    p->nthreads = 1;
    p->coarse_exact = 0;
}

int mgo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define T double
#define SUF _f64
#include "mg_oracle_impl.inc"
#undef T
#undef SUF

#define T float
#define SUF _f32
#include "mg_oracle_impl.inc"
#undef T
#undef SUF
