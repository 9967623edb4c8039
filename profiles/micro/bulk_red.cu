// Micro-benchmark: FP64 node accumulation of one 32-particle chunk of the warp-cooperative scatter done two ways.
//  A  as shipped: per dual-cell group NV RED instructions from 27 lanes (9 rows x 3 nodes), ~4.5 groups per chunk
//  B  per-warp shared-memory tile [NV][9 rows][W columns] handed to the grid by the TMA unit: one
//     cp.reduce.async.bulk.global.shared::cta.add.f64 per row (UBLKRED, issued once per warp through the uniform datapath,
//     does not occupy the LSU pipe), rows 16-byte aligned (even node index, even length)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_red bulk_red.cu ; run on one B200.
#include <cstdio>
#include <cuda_runtime.h>

#define YP 118
#define ZP (118 * 118)
#define NN (118 * 118 * 118)

template <int NV, int GROUPS>
__global__ void k_redg(double *a, int iters)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int base = ZP + YP + 2 + (int)(((long long)warp * iters * GROUPS) % (NN - 4 * ZP - 600));
    const int off = (lane % 3 - 1) + ((lane / 3) % 3 - 1) * YP + (lane / 9 - 1) * ZP;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int gI = 0; gI < GROUPS; gI++) {
            if (lane < 27) {
#pragma unroll
                for (int v = 0; v < NV; v++) atomicAdd(&a[(size_t)v * NN + base + off], 1.0);
            }
            base += 1;
        }
        if (base > NN - 3 * ZP - 600) base = ZP + YP + 2;
    }
}

// W = columns of the tile (even), WAIT_EVERY = iterations between waits for the bulk group (the tile must not be rewritten
// before the TMA unit has read it: the real kernel waits once per chunk)
template <int NV, int W, int WARPS>
__global__ void k_bulk(double *a, int iters, int advance)
{
    __shared__ alignas(16) double tile[WARPS][NV][9][W];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int base = ZP + YP + 2 + (int)(((long long)warp * iters * advance) % (NN - 4 * ZP - 600));
    for (int it = 0; it < iters; it++) {
        // the member loops' results land in the tile (here: a fill, 9*W*NV/32 stores per lane)
        for (int e = lane; e < NV * 9 * W; e += 32) (&tile[w][0][0][0])[e] = 1.0;
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) {
            const int b0 = base & ~1;          // even node index: 16-byte aligned row start
#pragma unroll
            for (int v = 0; v < NV; v++)
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    double *dst = a + (size_t)v * NN + b0 - 2 + (r % 3 - 1) * YP + (r / 3 - 1) * ZP;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(&tile[w][v][r][0]);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(W * 8) : "memory");
                }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        base += advance;
        if (base > NN - 3 * ZP - 600) base = ZP + YP + 2;
    }
}

template <class F>
static float timeit(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(2000);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return ms;
}

int main()
{
    double *a;
    cudaMalloc(&a, sizeof(double) * (size_t)NN * 4);
    cudaMemset(a, 0, sizeof(double) * (size_t)NN * 4);
    const int blocks = 148 * 8, threads = 128, iters = 2000;
    const double chunks = (double)blocks * (threads / 32) * iters;
    auto report = [&](const char *name, float ms) {
        const double ns = ms * 1e6 / (chunks / 148.);
        printf("%-64s %.3f ms  %.1f ns = %.0f cycles per 32-particle chunk per SM\n", name, ms, ns, ns * 1.965);
    };
    report("A  NV=4: 5 groups x 4 REDG (27 lanes)                [F1]", timeit([&](int it) { k_redg<4, 5><<<blocks, threads>>>(a, it); }));
    report("A  NV=3: 5 groups x 3 REDG (27 lanes)             [F2, F3]", timeit([&](int it) { k_redg<3, 5><<<blocks, threads>>>(a, it); }));
    report("B  NV=4: 36 UBLKRED of 64 B (W=8) + tile fill + wait", timeit([&](int it) { k_bulk<4, 8, 4><<<blocks, threads>>>(a, it, 5); }));
    report("B  NV=3: 27 UBLKRED of 64 B (W=8) + tile fill + wait", timeit([&](int it) { k_bulk<3, 8, 4><<<blocks, threads>>>(a, it, 5); }));
    report("B  NV=4: 36 UBLKRED of 80 B (W=10) + tile fill + wait", timeit([&](int it) { k_bulk<4, 10, 4><<<blocks, threads>>>(a, it, 5); }));
    report("B  NV=4: 36 UBLKRED of 128 B (W=16) + tile fill + wait, advance 12", timeit([&](int it) { k_bulk<4, 16, 4><<<blocks, threads>>>(a, it, 12); }));
    // correctness of the bulk reduction: total added must equal the number of tile entries handed over
    cudaMemset(a, 0, sizeof(double) * (size_t)NN * 4);
    k_bulk<4, 8, 4><<<blocks, threads>>>(a, 10, 5);
    cudaDeviceSynchronize();
    double *h = (double *)malloc(sizeof(double) * (size_t)NN * 4);
    cudaMemcpy(h, a, sizeof(double) * (size_t)NN * 4, cudaMemcpyDeviceToHost);
    double sum = 0.;
    for (size_t i = 0; i < (size_t)NN * 4; i++) sum += h[i];
    printf("bulk reduction check: sum %.1f, expected %.1f\n", sum, (double)blocks * 4 * 10 * 4 * 9 * 8);
    free(h);
    cudaFree(a);
    return 0;
}
