// Micro-benchmark: cost of FP64 RED (atomicAdd without return) per warp instruction for the address patterns the
// P2G scatter can produce.  Answers: is the L1TEX cost per lane, per 32-byte sector or per 128-byte line?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_patterns red_patterns.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>

#define YP 117
#define ZP (117 * 117)
#define NN (117 * 117 * 117)

template <int PAT>
__device__ __forceinline__ bool lane_target(int lane, int base, int &node)
{
    if (PAT == 0) { if (lane >= 27) return false; node = base + (lane % 3 - 1) + ((lane / 3) % 3 - 1) * YP + (lane / 9 - 1) * ZP; return true; }
    if (PAT == 1) { if (lane >= 9) return false; node = base + (lane % 3 - 1) * YP + (lane / 3 - 1) * ZP; return true; }
    if (PAT == 2) { node = base + lane; return true; }
    if (PAT == 3) { node = base + (lane & 7) + (lane >> 3) * YP; return true; }
    if (PAT == 4) { node = base + lane * 16; return true; }
    if (PAT == 5) { if (lane >= 18) return false; node = base + (lane % 2) + ((lane / 2) % 3 - 1) * YP + (lane / 6 - 1) * ZP; return true; }
    if (PAT == 6) { node = base + (lane & 3) + ((lane >> 2) % 3 - 1) * YP + ((lane >> 2) / 3 - 1) * ZP; return (lane >> 2) < 8; }   // 8 rows of 4
    if (PAT == 7) { node = base + (lane & 15) + (lane >> 4) * YP; return true; }                                                   // 2 rows of 16
    if (PAT == 8) { if (lane >= 27) return false; node = base + (lane % 9) + (lane / 9 - 1) * YP; return true; }                   // 3 rows of 9
    if (PAT == 9) { node = base; return true; }                                                                                    // single address
    return false;
}

template <int PAT>
__global__ void k_red(double *a, int iters, int stepPerIter)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int base = ZP + YP + 1 + (int)(((long long)warp * iters * stepPerIter) % (NN - 4 * ZP - 600));
    for (int it = 0; it < iters; it++) {
        int node;
        if (lane_target<PAT>(lane, base, node)) atomicAdd(&a[node], 1.0);
        base += stepPerIter;
        if (base > NN - 3 * ZP - 600) base = ZP + YP + 1;
    }
}

template <int PAT>
void run(const char *name, double *a, int lanes, int stepPerIter)
{
    const int blocks = 148 * 8, threads = 128, iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_red<PAT><<<blocks, threads>>>(a, 200, stepPerIter);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_red<PAT><<<blocks, threads>>>(a, iters, stepPerIter);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double reds = (double)blocks * (threads / 32) * iters;
    const double nsPerRedPerSM = ms * 1e6 / (reds / 148.);
    printf("%-34s lanes %2d  %.3f ms  %.2f ns/warp-RED/SM  = %.1f cyc @1.9GHz  (%.2f cyc/lane)  %.1f G lane-atomics/s\n", name, lanes, ms, nsPerRedPerSM,
           nsPerRedPerSM * 1.9, nsPerRedPerSM * 1.9 / lanes, reds * lanes / ms / 1e6);
}

int main()
{
    double *a;
    cudaMalloc(&a, sizeof(double) * NN);
    cudaMemset(a, 0, sizeof(double) * NN);
    run<0>("27 lanes: 9 rows x 3 (dual cell)", a, 27, 1);
    run<1>("9 lanes: one column, 9 rows", a, 9, 1);
    run<5>("18 lanes: 9 rows x 2", a, 18, 2);
    run<6>("32 lanes: 8 rows x 4", a, 32, 4);
    run<3>("32 lanes: 4 rows x 8", a, 32, 8);
    run<8>("27 lanes: 3 rows x 9", a, 27, 9);
    run<7>("32 lanes: 2 rows x 16", a, 32, 16);
    run<2>("32 lanes: 1 row x 32 (coalesced)", a, 32, 32);
    run<4>("32 lanes: 32 distinct lines", a, 32, 1);
    run<9>("32 lanes: one address", a, 32, 1);
    run<0>("27 lanes dual cell, stride 7 cells", a, 27, 7);
    cudaFree(a);
    return 0;
}
