/*
 * krylov_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the
 * Krylov.jl hot path (see krylov_oracle_impl.h for the per-function
 * reference citations and the parity-pinning status).  Built by
 * oracle/Makefile into oracle/libkrylov_oracle.so; loaded by oracle/oracle.py.
 * The product library (krylov.jl_b200/) never links, loads or calls this.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

/* ---- Float64 instantiation ---- */
#define REAL double
#define SUF(name) name##_f64
#define SQRT sqrt
#define FABS fabs
#define COPYSIGN copysign
#define POW pow
#define EPS DBL_EPSILON
#include "krylov_oracle_impl.h"
#undef REAL
#undef SUF
#undef SQRT
#undef FABS
#undef COPYSIGN
#undef POW
#undef EPS

/* ---- Float32 instantiation ---- */
#define REAL float
#define SUF(name) name##_f32
#define SQRT sqrtf
#define FABS fabsf
#define COPYSIGN copysignf
#define POW powf
#define EPS FLT_EPSILON
#include "krylov_oracle_impl.h"

int oracle_abi_version(void) { return 1; }
