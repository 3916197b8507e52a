/*
 * ORACLE (test infrastructure, NOT product code) -- see nsf_oracle_impl.h.
 * Builds the float ("_f32") and double ("_f64") restatements of the NF-iSAM
 * autoregressive spline flow into liboracle_nsf.so (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL float
#define SUFFIX _f32
#include "nsf_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX _f64
#include "nsf_oracle_impl.h"
#undef REAL
#undef SUFFIX

/* Thread count of the OpenMP loops (bench.py sets it explicitly: torchrun exports OMP_NUM_THREADS=1).  Returns the
 * count in effect. */
#ifdef _OPENMP
int nsf_set_num_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
int nsf_set_num_threads(int n) { (void)n; return 1; }
#endif
