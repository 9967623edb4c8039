// TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> so that the DEVICE constitutive code
// (nairn_mpm_fea_b200/csrc/materials.cuh) can be compiled by g++ and exercised on a machine without a GPU
// (tests/test_device_laws_cpu.py).  Nothing here is used by the product build.
#pragma once
#include <math.h>
#include <stdint.h>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
static inline double atomicAdd(double *a, double v) { double o = *a; *a += v; return o; }
// round-to-nearest intrinsics = plain IEEE operations (the host build uses -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
