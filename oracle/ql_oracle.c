/*
 * ql_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU oracle for the QL hot path of
 * MatrixFactorizations.jl (see ql_oracle_impl.h for the per-function reference citations).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libqlb200.so) never links or calls it.
 *
 * Build:  make -C oracle      ->  oracle/liboracle.so
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stddef.h>

/* double instance: orc_d* */
#define T double
#define FN(name) orc_d##name
#define FABS fabs
#define SQRT sqrt
#define COPYSIGN copysign
#define HUGEV DBL_MAX
#define SAFMIN DBL_MIN            /* dlamch('S') = 2^-1022 */
#define EPSV (DBL_EPSILON * 0.5)  /* dlamch('E') = 2^-53   */
#include "ql_oracle_impl.h"
#undef T
#undef FN
#undef FABS
#undef SQRT
#undef COPYSIGN
#undef HUGEV
#undef SAFMIN
#undef EPSV

/* float instance: orc_s* */
#define T float
#define FN(name) orc_s##name
#define FABS fabsf
#define SQRT sqrtf
#define COPYSIGN copysignf
#define HUGEV FLT_MAX
#define SAFMIN FLT_MIN
#define EPSV (FLT_EPSILON * 0.5f)
#include "ql_oracle_impl.h"
