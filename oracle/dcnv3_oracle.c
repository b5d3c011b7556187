/*
 * TEST INFRASTRUCTURE ONLY (see dcnv3_oracle_impl.h): builds the f32 and f64 instances of the CPU
 * restatement of the reference DCNv3 kernels into oracle/_build/libgp_oracle.so.
 * Build: make -C oracle        (gcc -O2 -ffp-contract=off -fopenmp)
 */
#include <math.h>

#define REAL float
#define SUFFIX f32
#define FMA(a, b, c) fmaf((a), (b), (c))
#define FLOOR(x) floorf(x)
#include "dcnv3_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FMA
#undef FLOOR

#define REAL double
#define SUFFIX f64
#define FMA(a, b, c) fma((a), (b), (c))
#define FLOOR(x) floor(x)
#include "dcnv3_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FMA
#undef FLOOR

int gpo_abi_version(void) { return 1; }
