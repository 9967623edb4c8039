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
