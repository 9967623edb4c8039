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

// ---- kernels on the host (tests/devlaws/host_step.cpp): one CUDA thread after the other ------------------------------
// Only kernels whose threads do not cooperate (no shared memory, no warp intrinsics, no __syncthreads) can run this way:
// the per-task kernels of kernels_task.cuh.  Atomics are plain read-modify-writes.
#define __launch_bounds__(...)
struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 blockIdx, threadIdx, blockDim, gridDim;
static inline int atomicAdd(int *a, int v) { int o = *a; *a += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *a, unsigned long long v) { unsigned long long o = *a; *a += v; return o; }
static inline int atomicCAS(int *a, int cmp, int v) { int o = *a; if (o == cmp) *a = v; return o; }
static inline int atomicMin(int *a, int v) { int o = *a; if (v < o) *a = v; return o; }
#define EMU_LAUNCH(KERNEL, GRID, BLOCK, ...) do { \
    gridDim.x = (unsigned)(GRID); gridDim.y = gridDim.z = 1; blockDim.x = (unsigned)(BLOCK); blockDim.y = blockDim.z = 1; \
    for (unsigned b_ = 0; b_ < (unsigned)(GRID); b_++) for (unsigned t_ = 0; t_ < (unsigned)(BLOCK); t_++) { \
        blockIdx.x = b_; blockIdx.y = blockIdx.z = 0; threadIdx.x = t_; threadIdx.y = threadIdx.z = 0; KERNEL(__VA_ARGS__); } } while (0)
