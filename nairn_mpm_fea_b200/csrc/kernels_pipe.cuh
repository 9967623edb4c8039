// TMA-pipelined persistent variants of the fused particle kernels.
//
// The per-particle code of F2/F4 (weights -> 27-node gather -> constitutive law) is latency-bound
// when every warp first waits for its own DRAM loads (profiles/r1b_*).  Here each warp owns a ring of
// shared-memory stages: one elected lane per field issues `cp.async.bulk` (TMA bulk copy, SASS
// UBLKCP) of the 256 contiguous bytes that hold the field for the warp's next 32 particles and the
// data's arrival is tracked by an mbarrier, so the DRAM fetch of chunk i+1 overlaps the arithmetic of
// chunk i without holding registers.  Blocks are persistent (grid = SMs x resident blocks) and warps
// stride over 32-particle chunks.
#pragma once
#include "kernels_fused.cuh"

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}

#define PIPE_STAGES 2
#define PIPE_MAXF 40

// which arrays a kernel stages: up to PIPE_MAXF double fields + 2 int fields
struct PipeFields {
    const double *d[PIPE_MAXF];
    const int *i[2];
    int nd, ni;
};

// bytes of one stage for one warp
__host__ __device__ inline size_t pipe_stage_bytes(int nd, int ni) { return (size_t)nd * 256 + (size_t)ni * 128; }

// issue the loads of `chunk` (32 particles starting at chunk*32) into stage memory `buf`
__device__ __forceinline__ void pipe_issue(const PipeFields &pf, int chunk, unsigned char *buf, unsigned long long *bar, int lane)
{
    const size_t p0 = (size_t)chunk * 32;
    if (lane == 0) mbar_expect_tx(bar, (unsigned)pipe_stage_bytes(pf.nd, pf.ni));
    __syncwarp();
    for (int f = lane; f < pf.nd; f += 32) bulk_g2s(buf + (size_t)f * 256, pf.d[f] + p0, 256, bar);
    if (lane < pf.ni) bulk_g2s(buf + (size_t)pf.nd * 256 + (size_t)lane * 128, pf.i[lane] + p0, 128, bar);
}

// field order staged for F4 / F2 (elastic):  0-2 ncpos, 3-11 F, 12-17 sp, 18 work, 19 heat, 20 entropy, 21 prevT,
// 22-24 pos (F4) | 22 mp (F2);  full: + 6 eplast, pressure, plast, res, 4 hist after those
enum { PF_NCPOS = 0, PF_F = 3, PF_SP = 12, PF_WORK = 18, PF_HEAT = 19, PF_ENTROPY = 20, PF_PREVT = 21, PF_X0 = 22 };

template <int SK>
__device__ __forceinline__ void stage_to_state(const double *st, int lane, int base, PState &s)
{
#pragma unroll
    for (int i = 0; i < 9; i++) s.F[i] = st[(PF_F + i) * 32 + lane];
#pragma unroll
    for (int i = 0; i < 6; i++) s.sp[i] = st[(PF_SP + i) * 32 + lane];
    s.work = st[PF_WORK * 32 + lane]; s.heat = st[PF_HEAT * 32 + lane]; s.entropy = st[PF_ENTROPY * 32 + lane];
    s.prevT = st[PF_PREVT * 32 + lane];
    if (SK == SK_FULL) {
#pragma unroll
        for (int i = 0; i < 6; i++) s.eplast[i] = st[(base + i) * 32 + lane];
        s.pressure = st[(base + 6) * 32 + lane]; s.plast = st[(base + 7) * 32 + lane]; s.res = st[(base + 8) * 32 + lane];
#pragma unroll
        for (int i = 0; i < MPM_MAX_HISTORY; i++) s.hist[i] = st[(base + 9 + i) * 32 + lane];
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) s.eplast[i] = 0.;
        s.pressure = 0.; s.plast = 0.; s.res = 0.;
#pragma unroll
        for (int i = 0; i < MPM_MAX_HISTORY; i++) s.hist[i] = 0.;
    }
}

// ---- F4 pipelined: second strain update + element reset --------------------------------------------------
template <int SK>
__global__ void __launch_bounds__(FUSED_THREADS) k_f4_pipe(Grid g, Particles P, FusedNodes FN, const Material *mats, PipeFields pf,
                                                           double strainTime, int doStrain, StatusFlags *flags, double dt, SlabInfo slab)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t stageBytes = pipe_stage_bytes(pf.nd, pf.ni);
    unsigned char *wbuf = smem + (size_t)warp * PIPE_STAGES * stageBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)FUSED_WARPS * PIPE_STAGES * stageBytes) + warp * PIPE_STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < PIPE_STAGES; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int nchunks = (P.n + 31) >> 5;
    const int stride = gridDim.x * FUSED_WARPS;
    int chunk = blockIdx.x * FUSED_WARPS + warp;
    // prologue: fill the ring
#pragma unroll
    for (int s = 0; s < PIPE_STAGES; s++) {
        const int c = chunk + s * stride;
        if (c < nchunks) pipe_issue(pf, c, wbuf + (size_t)s * stageBytes, &bars[s], lane);
    }
    const int fullBase = PF_X0 + 3;
    for (int it = 0; chunk < nchunks; chunk += stride, it++) {
        const int s = it % PIPE_STAGES;
        const unsigned phase = (it / PIPE_STAGES) & 1;
        mbar_wait(&bars[s], phase);
        const double *st = reinterpret_cast<const double *>(wbuf + (size_t)s * stageBytes);
        const int *sti = reinterpret_cast<const int *>(wbuf + (size_t)s * stageBytes + (size_t)pf.nd * 256);
        const int p = chunk * 32 + lane;
        const bool active = p < P.n;
        int elem = 1;
        double pos[3] = {0., 0., 0.};
        if (active) {
            elem = sti[lane];
            pos[0] = st[(PF_X0 + 0) * 32 + lane]; pos[1] = st[(PF_X0 + 1) * 32 + lane]; pos[2] = st[(PF_X0 + 2) * 32 + lane];
            if (doStrain && p < P.nNR) {
                double xi[3], lp[3];
                xi[0] = st[0 * 32 + lane]; xi[1] = st[1 * 32 + lane]; xi[2] = st[2 * 32 + lane];
                load_lp(g, P, p, lp);
                prefetch_stencil(g, dual_cell_center(g, elem, xi), FN.V);
                double dv[9];
                {
                    Weights3 w;
                    particle_weights<true>(g, dual_cell_center(g, elem, xi), xi, lp, w);
                    gather_gradv_impl<false>(g, w, FN.V, (const WarpTile *)0, 0, dv);
                }
#pragma unroll
                for (int i = 0; i < 9; i++) dv[i] *= strainTime;
                PState ps;
                stage_to_state<SK>(st, lane, fullBase, ps);
                constitutive_law<3, SK == SK_ELASTIC>(ps, dv, strainTime, g.np, mats[sti[32 + lane]]);
                store_state<SK>(P, p, ps);
            }
        }
        __syncwarp();                       // every lane has read its stage: refill it
        const int cnext = chunk + PIPE_STAGES * stride;
        if (cnext < nchunks) pipe_issue(pf, cnext, wbuf + (size_t)s * stageBytes, &bars[s], lane);
        if (active) {
            // element reset: the common case (still inside its element) is decided from the staged position
            if (!(pos[0] == pos[0] && pos[1] == pos[1] && pos[2] == pos[2]) || !pt_in_element<3>(g, elem, pos))
                reset_element_one<3>(g, P, p, flags, dt);
            if (slab.on) {
                const int k = (P.elem[p] - 1) / (g.horiz * g.vert);
                const int side = k < slab.cellLo ? 0 : (k >= slab.cellHi ? 1 : -1);
                if (side >= 0) {
                    const int slot = atomicAdd(&slab.leaveCount[side], 1);
                    if (slot < slab.leaveCap) slab.leaveIdx[side * slab.leaveCap + slot] = p;
                }
            }
        }
    }
}
