/*
 * cattl3_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the C-ATTL3 hot path: a plain-C restatement of the reference algorithms
 * (see cattl3_oracle_impl.h for the per-function reference citations and the parity status).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file; the product path (c-attl3_b200/) never does.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { /* == cattl3_conv_geom, include/cattl3_b200.h */
	int n, h, w, c, f, rh, rw, ph, pw, sh, sw, dh, dw;
} orc_geom;

#define S float
#define ORC_MAX FLT_MAX
#define FN(name) name##_f32
#include "cattl3_oracle_impl.h"
#undef S
#undef ORC_MAX
#undef FN

#define S double
#define ORC_MAX DBL_MAX
#define FN(name) name##_f64
#include "cattl3_oracle_impl.h"
#undef S
#undef ORC_MAX
#undef FN
