// Fused fast path for the 3D uGIMP step (USAVG/USF/USL, FLIP/PIC, any material in materials.cuh).
//
// Four particle kernels + three node sweeps per step replace the ten per-task kernels:
//   F1  ncpos + P2G mass/momentum                       (tasks 1,2)
//   N1  save momenta, velocity BCs, v = p/m              (tasks 3, 4a)
//   F2  G2P grad v + constitutive law + P2G forces       (tasks 4b, 5)
//   N2  restore p, gravity, BCs, p += f dt, BCs, v, a    (tasks 6, 7, 8a) + zero p for the re-extrapolation
//   F3  G2P particle update + P2G momentum (new v)       (tasks 8b, 9a)
//   N3  velocity BCs, v = p/m                            (task 9b)
//   F4  G2P grad v + constitutive law + element reset    (tasks 9c, 11)
//
// Particle -> grid transfers use the DUAL CELL of a particle: the node nearest to it.  With a
// particle no larger than a cell (lp <= 1) the uGIMP stencil is exactly the 3x3x3 nodes around that
// node (shape.cuh explains why this is the reference's node set).  Particles are kept physically
// sorted by dual cell (integer radix key = centre node index, re-sorted every few steps), so the
// ~8 particles of a dual cell sit in neighbouring lanes.  A warp takes 32 consecutive particles,
// each lane computes its particle's 1-D weights and payload into shared memory, then the warp
// walks the groups of equal key (__match_any_sync): lanes 0..26 BECOME the 27 nodes of that dual
// cell and accumulate the group's particles in registers; one FP64 RED per node and component
// leaves the warp.  That cuts global atomics ~8x relative to one atomic per particle-node pair and
// never depends on the sort being exact (an out-of-place particle just forms its own group).
//
// Grid -> particle transfers are one thread per particle reading 32-byte node records
// {vx,vy,vz,pad} / {ax,ay,az,pad} that the node sweeps write; the records of a whole 8M-particle
// grid fit in the 126 MB L2.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include "mpm_types.cuh"
#include "shape.cuh"
#include "materials.cuh"
#include "kernels_task.cuh"

#define FUSED_THREADS 128
#define FUSED_WARPS (FUSED_THREADS / 32)
// minimum resident blocks per SM asked of the register allocator (tuned on B200, profiles/tune_bounds_r1.txt)
#ifndef F1_MINB
#define F1_MINB 8
#endif
#ifndef F2_MINB
#define F2_MINB 5
#endif
#ifndef F3_MINB
#define F3_MINB 8
#endif
#ifndef F4_MINB
#define F4_MINB 6
#endif

struct FusedNodes {
    double4 *V;          // [nnodes] {vx,vy,vz,0}: vk[0]
    double4 *A;          // [nnodes] {ax,ay,az,0}: ftot/mass
    double4 *VS;         // [nnodes] {v*prev}: the vector the XPIC/FMPM iterations gather (order > 1)
    const int *bcOfNode; // [nnodes] index into VelBCs unique list or -1
    RigidBCs R;          // BCs projected from rigid particles (k_project_rigid_bcs)
};

// slab decomposition along z (one process per GPU): this rank owns cell planes [cellLo, cellHi)
struct SlabInfo {
    int on;
    int cellLo, cellHi;
    int *leaveCount;         // [2]: particles whose new element lies below / above the slab
    int *leaveIdx;           // [2][leaveCap] their indices
    int leaveCap;
};

struct TiledState {
    int enabled;         // fast path usable for this context
    int stateKind;       // SK_ELASTIC / SK_FULL: which particle fields the materials in use touch
    int usePipe, numSMs; // TMA-pipelined persistent kernels (kernels_pipe.cuh)
    int sortInterval;
    long long stepsSinceSort;
    FusedNodes FN;
    // sorting workspace
    int *keysIn, *keysOut, *idxIn, *idxOut;
    void *cubTemp; size_t cubTempBytes;
    double *altPool; int *altIntPool;       // second particle pool for the physical reorder
    size_t cap;
    // slab mode
    SlabInfo slab;
    int hasLower, hasUpper;
    int nodeLo, nodeCount;                  // node range the sweeps and the zeroing cover
    double *haloSend[2], *haloRecv[2];      // [lower, upper]; 3 planes x up to 5 values
    size_t planeNodes;
    double *migSend[2], *migRecv[2];        // rows of MIG_ROW doubles
    int *migSorted, *migKeys, *migFillers, *migFlags, *migPairs;     // device-side bookkeeping (k_mig_plan)
    void *migCubTemp; size_t migCubBytes;
    int migCap;
    int hLeave[2];                          // host copy of leaveCount after the step
};

static inline void tiled_state_init(TiledState &t) { memset(&t, 0, sizeof t); }
static inline void tiled_state_free(TiledState &t) {}
static inline void tiled_on_upload(TiledState &t) { t.stepsSinceSort = 1 << 30; }

// 32-byte node record through the read-only path
__device__ __forceinline__ double4 ldg4(const double4 *p)
{
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// Ask for the 9 rows (3 consecutive 32-byte records each) of a particle's 27-node stencil to be brought
// into L1 now; issued as soon as the centre node is known, long before the gather loop needs them, and
// costs no registers (the gather's own loads otherwise serialise into several L2 round trips).
__device__ __forceinline__ void prefetch_stencil(const Grid &g, int center, const double4 *R)
{
#pragma unroll
    for (int k = -1; k <= 1; k++)
#pragma unroll
        for (int j = -1; j <= 1; j++) {
            const double4 *row = R + (center + j * g.yplane + k * g.zplane - 1);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(row));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 2));
        }
}

// ---- per-warp node tile in shared memory -----------------------------------------------------------
// The particles of a warp are (nearly) sorted by dual cell, so their 27-node stencils overlap in a window
// of TILE_W consecutive node columns: nine linear runs of TILE_W 32-byte records (the node index is linear,
// so a run that crosses the end of a grid row simply continues into the next row).  The warp copies the
// window once with cp.async (one L2 round trip for the whole warp) and every lane gathers its 27 records
// from shared memory; a lane whose stencil is not inside the window (a particle that drifted since the
// last sort, or a warp spanning more than TILE_W-2 dual cells) reads global memory as before.
#ifndef TILE_W
#define TILE_W 8
#endif
struct WarpTile { double4 r[9][TILE_W]; };

// window start (centre-row node index of column 0 is anchor-1): covers centres anchor .. anchor+TILE_W-3
__device__ __forceinline__ int tile_anchor(int key, bool active)
{
    const unsigned full = 0xffffffffu;
    const unsigned act = __ballot_sync(full, active);
    const int kmin = __reduce_min_sync(full, active ? key : 0x7fffffff);
    const int ref = __shfl_sync(full, key, __popc(act) >> 1);       // active lanes are the low ones
    if (!act) return 1;                                             // a warp past the last particle: any valid window
    return max(kmin, ref - (TILE_W - 3));
}

__device__ __forceinline__ void tile_load_async(const Grid &g, WarpTile &t, int anchor, const double4 *__restrict__ R)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c0 = 0; c0 < 9 * TILE_W * 2; c0 += 32) {
        const int c = c0 + lane;
        if (c < 9 * TILE_W * 2) {
            const int rec = c >> 1, half = c & 1;
            const int row = rec / TILE_W, xo = rec - row * TILE_W;
            int node = anchor - 1 + xo + ((row % 3) - 1) * g.yplane + ((row / 3) - 1) * g.zplane;
            node = min(max(node, 0), g.nnodes - 1);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<char *>(&t.r[row][xo]) + 16 * half);
            const char *src = reinterpret_cast<const char *>(R + node) + 16 * half;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
    }
}

__device__ __forceinline__ void tile_wait()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
}

// centre node of the dual cell from the element and the sign of the natural coordinates
__device__ __forceinline__ int dual_cell_center(const Grid &g, int inElem, const double xi[3])
{
    const ElemIJK c = elem_ijk(g, inElem);
    return elem_node0(g, c) + (xi[0] < 0. ? 0 : 1) + (xi[1] < 0. ? 0 : 1) * g.yplane + (xi[2] < 0. ? 0 : 1) * g.zplane;
}

// ---- 3-node uGIMP weights around the dual-cell centre -------------------------------------------
// base = -1 (xi<0: nodes at -3,-1,1) or 0 (xi>=0: nodes at -1,1,3); returns S[3], dS[3] (signed, not
// yet scaled by 2/dx) and a 3-bit validity mask (xp < 2+lp).
// Written with selects instead of the reference's branch ladder (EightNodeIsoparamBrick.cpp:316-387): the lanes of a warp
// sit in different pieces of the piecewise weight, so branches would run every piece one after the other anyway.
// Each piece is the reference's expression, so the selected value is bit-identical to the ladder's result.  With the
// particle inside its element (|xi| <= 1) and lp <= 1 a side node is at least one cell away (xp >= 1 >= lp), so its
// first piece is never taken; the centre node is closer than 2+lp, so it is never outside the support.
template <bool GRAD, bool CENTRE>
__device__ __forceinline__ void gimp_piece(double xi, double xn, double lp, double q1, double q2, double inv_size, double inv2lp, double &S, double &dS, bool &ok)
{
    const double xp = fabs(xi - xn);
    const double arg = (q2 - xp) * inv_size;
    const double s3 = 2. * lp * arg * arg;
    const double s2 = 0.5 * (2. - xp);
    double s, d;
    if (CENTRE) {
        const double s1 = ((4. - lp) * lp - xp * xp) * inv_size;
        const bool in1 = xp < lp, in2 = xp <= q1;
        s = in1 ? s1 : (in2 ? s2 : s3);
        d = in1 ? -xp * inv2lp : (in2 ? -0.5 : -arg);
        ok = true;
    } else {
        const bool in2 = xp <= q1;
        ok = xp < q2;
        s = in2 ? s2 : (ok ? s3 : 0.);
        d = in2 ? -0.5 : (ok ? -arg : 0.);
    }
    S = s;
    if (GRAD) dS = (xi > xn) ? d : -d;
}

template <bool GRAD>
__device__ __forceinline__ int gimp3(double xi, double lp, double inv_size, double inv2lp, double S[3], double dS[3], unsigned &ok)
{
    const int base = xi < 0. ? -1 : 0;
    const double q1 = 2. - lp, q2 = 2. + lp;
    const double xc = xi < 0. ? -1. : 1.;          // centre node of the dual cell
    bool ok0, ok1, ok2;
    double d0 = 0., d1 = 0., d2 = 0.;
    gimp_piece<GRAD, false>(xi, xc - 2., lp, q1, q2, inv_size, inv2lp, S[0], d0, ok0);
    gimp_piece<GRAD, true>(xi, xc, lp, q1, q2, inv_size, inv2lp, S[1], d1, ok1);
    gimp_piece<GRAD, false>(xi, xc + 2., lp, q1, q2, inv_size, inv2lp, S[2], d2, ok2);
    if (GRAD) { dS[0] = d0; dS[1] = d1; dS[2] = d2; }
    ok = (ok0 ? 1u : 0u) | 2u | (ok2 ? 4u : 0u);
    return base;
}

struct Weights3 {
    double S[3][3];      // [axis][t]
    double dS[3][3];     // gradient factors, already multiplied by 2/delta of the element
    unsigned ok;         // 9 bits: axis*3+t
    int center;          // 0-based node index of the dual-cell centre
};

// center = dual_cell_center(element, xi): F1 computes it once per step and keeps it in P.key for the later kernels
template <bool GRAD>
__device__ __forceinline__ void particle_weights(const Grid &g, int center, const double xi[3], const double lp[3], Weights3 &w)
{
    unsigned okx, oky, okz;
    // inv_size and the branch-1 derivative divisor follow EightNodeIsoparamBrick.cpp:296-303,:365-387
    // (z uses 1/(4 lp.y)); with a uniform particle size the reciprocals are per-run constants
    double isx, isy, i2x, i2y, i2z;
    if (g.lpUniform) { isx = g.lpInvSize[0]; isy = g.lpInvSize[1]; i2x = g.lpInv2[0]; i2y = g.lpInv2[1]; i2z = g.lpInv2[2]; }
    else { isx = 1. / (4. * lp[0]); isy = 1. / (4. * lp[1]); i2x = 1. / (2. * lp[0]); i2y = 1. / (2. * lp[1]); i2z = 1. / (2. * lp[2]); }
    gimp3<GRAD>(xi[0], lp[0], isx, i2x, w.S[0], w.dS[0], okx);
    gimp3<GRAD>(xi[1], lp[1], isy, i2y, w.S[1], w.dS[1], oky);
    gimp3<GRAD>(xi[2], lp[2], isy, i2z, w.S[2], w.dS[2], okz);
    w.ok = okx | (oky << 3) | (okz << 6);
    w.center = center;
    if (GRAD) {
        // 2/delta of the element: equal element sizes on this path (the reference divides by the element's own
        // extent, which differs from the grid constant by at most an ulp)
        const double ix = g.inv2d[0], iy = g.inv2d[1], iz = g.inv2d[2];
#pragma unroll
        for (int t = 0; t < 3; t++) { w.dS[0][t] *= ix; w.dS[1][t] *= iy; w.dS[2][t] *= iz; }
    }
}

// ---- per-warp shared staging ----------------------------------------------------------------------
// Each particle lane stores FACTORISED weights of its 27-node stencil: the three x factors and the
// nine (y,z) products, so that a lane acting as node (i,j,k) forms S = X[i]*YZ[j+3k] with one
// multiply.  Rows are 33 doubles long: particle lanes write consecutive words, node lanes read with
// an odd stride -- both conflict-free.
#define WS_STRIDE 33
template <bool GRAD, int NQ>
struct WarpStage {
    double X[3][WS_STRIDE];                     // Sx_i
    double YZ[9][WS_STRIDE];                    // Sy_j * Sz_k
    alignas(16) double Q[32][NQ];               // payload per particle lane (NQ even: read as double2 broadcasts)
};
// With gradients the stage keeps the six per-axis factor triplets (18 rows instead of 33: shared memory is what
// limits the resident warps of the force kernel); the node lane forms the (y,z) products itself, in the same
// order as the particle lane would, so the sums are unchanged.
template <int NQ>
struct WarpStage<true, NQ> {
    double X[3][WS_STRIDE], Y[3][WS_STRIDE], Z[3][WS_STRIDE];       // S
    double DX[3][WS_STRIDE], DY[3][WS_STRIDE], DZ[3][WS_STRIDE];    // dS (scaled by 2/delta)
    alignas(16) double Q[32][NQ];
};

template <int NQ>
__device__ __forceinline__ void load_payload(const double (*Q)[NQ], int src, double q[NQ])
{
    const double2 *r = reinterpret_cast<const double2 *>(Q[src]);
#pragma unroll
    for (int i = 0; i < NQ / 2; i++) { const double2 t = r[i]; q[2 * i] = t.x; q[2 * i + 1] = t.y; }
}

// `scale` (the particle mass in the two momentum scatters) rides on the x factors, so the node lanes need one
// payload double less per member: S*mp = (Sx*mp)*(Sy*Sz)
template <int NQ>
__device__ __forceinline__ void stage_weights(WarpStage<false, NQ> &st, int lane, const Weights3 &w, double scale)
{
#pragma unroll
    for (int t = 0; t < 3; t++) st.X[t][lane] = w.S[0][t] * scale;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) st.YZ[j + 3 * k][lane] = w.S[1][j] * w.S[2][k];
}

template <int NQ>
__device__ __forceinline__ void stage_weights(WarpStage<true, NQ> &st, int lane, const Weights3 &w)
{
#pragma unroll
    for (int t = 0; t < 3; t++) {
        st.X[t][lane] = w.S[0][t]; st.Y[t][lane] = w.S[1][t]; st.Z[t][lane] = w.S[2][t];
        st.DX[t][lane] = w.dS[0][t]; st.DY[t][lane] = w.dS[1][t]; st.DZ[t][lane] = w.dS[2][t];
    }
}

// ---- the warp-cooperative scatter ----------------------------------------------------------------
// Groups = lanes whose particles share a dual cell (equal key).  For each group, lanes 0..26 act as
// the 27 nodes around the group's centre node: contrib(src, i, jk, acc[]) adds particle `src`'s
// contribution for node (i, jk=j+3k) into registers; one RED per node and component leaves the warp.
template <int NV, bool COUNT, class Contrib>
__device__ __forceinline__ void warp_scatter(const Grid &g, int key, bool active, double *const *dst, int *cnt, Contrib contrib)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned mymask = __match_any_sync(full, active ? key : (-1 - lane));
    unsigned remaining = __ballot_sync(full, active);
    const int li = lane % 3, ljk = lane / 3;                 // node of this lane (lane < 27): i, j+3k
    const int nodeOff = (li - 1) + ((ljk % 3) - 1) * g.yplane + ((ljk / 3) - 1) * g.zplane;
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const unsigned grp = __shfl_sync(full, mymask, leader);
        const int center = __shfl_sync(full, key, leader);
        remaining &= ~grp;
        if (lane < 27) {
            double acc[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) acc[v] = 0.;
            int n = 0;
            // sorted particles: the members of a group are consecutive lanes starting at the leader, so the
            // loop is a plain counted one (constant address strides); any other mask takes the bit scan
            const int members = __popc(grp);
            if ((grp >> leader) == (0xffffffffu >> (32 - members))) {
#pragma unroll 2
                for (int s = leader; s < leader + members; s++) n += contrib(s, li, ljk, acc);
            } else {
                unsigned mm = grp;
                while (mm) {
                    const int s = __ffs(mm) - 1;
                    mm &= mm - 1;
                    n += contrib(s, li, ljk, acc);
                }
            }
            if (n) {
                const int nd = center + nodeOff;
#pragma unroll
                for (int v = 0; v < NV; v++) atomAdd(&dst[v][nd], acc[v]);
                if (COUNT) atomicAdd(&cnt[nd], n);
            }
        }
    }
}

__device__ __forceinline__ void load_lp(const Grid &g, const Particles &P, int p, double lp[3])
{
    if (g.lpUniform) { lp[0] = g.lpU[0]; lp[1] = g.lpU[1]; lp[2] = g.lpU[2]; }
    else { lp[0] = P.lp[0][p]; lp[1] = P.lp[1][p]; lp[2] = P.lp[2][p]; }
}

// ---- particle state <-> registers, specialised by what the materials in use touch ------------------
// SK_ELASTIC: IsotropicMat only: F, sp, work, heat, entropy (+prevT read).  SK_FULL: everything.
enum { SK_ELASTIC = 0, SK_FULL = 1 };

template <int SK>
__device__ __forceinline__ void load_state(const Particles &P, int p, PState &s)
{
    if constexpr (SK == SK_FULL) load_pstate(P, p, s);
    else {
#pragma unroll
        for (int i = 0; i < 9; i++) s.F[i] = P.F[i][p];
#pragma unroll
        for (int i = 0; i < 6; i++) { s.sp[i] = P.sp[i][p]; s.eplast[i] = 0.; }
        s.pressure = 0.; s.res = 0.; s.plast = 0.;
        s.work = P.work[p]; s.heat = P.heat[p]; s.entropy = P.entropy[p];
        s.prevT = P.prevT[p];
        s.dT = 0.; s.dTad = 0.; s.adiabatic = 0;
#pragma unroll
        for (int i = 0; i < MPM_MAX_HISTORY; i++) s.hist[i] = 0.;
    }
}

template <int SK>
__device__ __forceinline__ void store_state(const Particles &P, int p, const PState &s)
{
    if constexpr (SK == SK_FULL) store_pstate(P, p, s);
    else {
#pragma unroll
        for (int i = 0; i < 9; i++) P.F[i][p] = s.F[i];
#pragma unroll
        for (int i = 0; i < 6; i++) P.sp[i][p] = s.sp[i];
        P.work[p] = s.work; P.heat[p] = s.heat; P.entropy[p] = s.entropy;
    }
}

// Pull the particle state this thread will need after its gather into L2 now (one lane per 128-byte
// line), so the loads issued after the gather loop hit L2 instead of exposing a second DRAM latency.
template <int SK>
__device__ __forceinline__ void prefetch_state(const Particles &P, int p)
{
    if ((threadIdx.x & 15) != 0) return;
#pragma unroll
    for (int i = 0; i < 9; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.F[i] + p));
#pragma unroll
    for (int i = 0; i < 6; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.sp[i] + p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.work + p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.heat + p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.entropy + p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.prevT + p));
    if (SK == SK_FULL) {
#pragma unroll
        for (int i = 0; i < 6; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.eplast[i] + p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.pressure + p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.plast + p));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.res + p));
#pragma unroll
        for (int i = 0; i < MPM_MAX_HISTORY; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.hist[i] + p));
    }
}

// ---- F1: ncpos + P2G mass and momentum ------------------------------------------------------------
__global__ void __launch_bounds__(FUSED_THREADS, F1_MINB) k_f1_mass_momentum(Grid g, Particles P, Nodes N)
{
    __shared__ WarpStage<false, 4> stage[FUSED_WARPS];
    WarpStage<false, 4> &st = stage[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = p < P.nNR;
    int key = 0;
    if (active) {
        const int e = P.elem[p];
        double pos[3] = {P.pos[0][p], P.pos[1][p], P.pos[2][p]};
        double xi[3], lp[3];
        load_lp(g, P, p, lp);
        get_xipos<3>(g, e, pos, xi);
        P.ncpos[0][p] = xi[0]; P.ncpos[1][p] = xi[1]; P.ncpos[2][p] = xi[2];
        Weights3 w;
        key = dual_cell_center(g, e, xi);
        P.key[p] = key;
        particle_weights<false>(g, key, xi, lp, w);
        stage_weights(st, lane, w, P.mp[p]);
        st.Q[lane][0] = P.vel[0][p]; st.Q[lane][1] = P.vel[1][p]; st.Q[lane][2] = P.vel[2][p];
    }
    __syncwarp();
    double *dst[4] = {N.mass, N.pk[0], N.pk[1], N.pk[2]};
    // (the node's particle count is only ever used as an activity flag: N1 derives it from mass != 0)
    warp_scatter<4, false>(g, key, active, dst, (int *)0, [&](int src, int i, int jk, double *acc) {
        const double f = st.X[i][src] * st.YZ[jk][src];          // fn * mp
        const double2 vxy = *reinterpret_cast<const double2 *>(st.Q[src]);
        const double vz = st.Q[src][2];
        acc[0] += f;
        acc[1] += vxy.x * f; acc[2] += vxy.y * f; acc[3] += vz * f;
        return f != 0. ? 1 : 0;
    });
}

// ---- gather of grad v from the V records -----------------------------------------------------------
template <bool TILED>
__device__ __forceinline__ void gather_gradv_impl(const Grid &g, const Weights3 &w, const double4 *__restrict__ V, const WarpTile *t, int d, double dv[9])
{
#pragma unroll
    for (int i = 0; i < 9; i++) dv[i] = 0.;
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double syz = w.S[1][j] * w.S[2][k];
            const double dyz = w.dS[1][j] * w.S[2][k];
            const double ydz = w.S[1][j] * w.dS[2][k];
            const int row = w.center + (j - 1) * g.yplane + (k - 1) * g.zplane - 1;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double4 v = TILED ? t->r[j + 3 * k][d + i] : ldg4(&V[row + i]);
                const double gx = w.dS[0][i] * syz;
                const double gy = w.S[0][i] * dyz;
                const double gz = w.S[0][i] * ydz;
                dv[0] += v.x * gx; dv[1] += v.x * gy; dv[2] += v.x * gz;
                dv[3] += v.y * gx; dv[4] += v.y * gy; dv[5] += v.y * gz;
                dv[6] += v.z * gx; dv[7] += v.z * gy; dv[8] += v.z * gz;
            }
        }
    }
}

// d = centre - anchor; the stencil is inside the warp's tile iff 0 <= d <= TILE_W-3
__device__ __forceinline__ void gather_gradv(const Grid &g, const Weights3 &w, const double4 *__restrict__ V, const WarpTile &t, int anchor, double dv[9])
{
    const int d = w.center - anchor;
    if ((unsigned)d <= (unsigned)(TILE_W - 3)) gather_gradv_impl<true>(g, w, V, &t, d, dv);
    else gather_gradv_impl<false>(g, w, V, &t, 0, dv);
}

// ---- F2: grad v + constitutive law + P2G forces ------------------------------------------------------
template <bool FEXT>
constexpr size_t f2_smem_bytes() { return FUSED_WARPS * (sizeof(WarpStage<true, FEXT ? 10 : 6>) + sizeof(WarpTile)); }

template <int SK, bool FEXT>
__global__ void __launch_bounds__(FUSED_THREADS, F2_MINB) k_f2_strain_forces(Grid g, Particles P, Nodes N, FusedNodes FN, const Material *mats,
                                                                    double strainTime, int doStrain)
{
    // dynamic shared memory (above the 48 KB static limit): per warp one weight stage + one node tile
    extern __shared__ __align__(16) unsigned char f2_smem[];
    typedef WarpStage<true, FEXT ? 10 : 6> Stage;
    Stage &st = reinterpret_cast<Stage *>(f2_smem)[threadIdx.x >> 5];
    WarpTile &tile = reinterpret_cast<WarpTile *>(f2_smem + FUSED_WARPS * sizeof(Stage))[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = p < P.nNR;
    int key = 0, anchor = 0;
    double xi[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
    if (active) {
        if (doStrain) prefetch_state<SK>(P, p);
        key = P.key[p];
        xi[0] = P.ncpos[0][p]; xi[1] = P.ncpos[1][p]; xi[2] = P.ncpos[2][p];
        load_lp(g, P, p, lp);
    }
    if (doStrain) {         // start the warp's node window on its way while the weights are computed
        anchor = tile_anchor(key, active);
        tile_load_async(g, tile, anchor, FN.V);
    }
    Weights3 w;
    if (active) {
        particle_weights<true>(g, key, xi, lp, w);
        stage_weights(st, lane, w);
    }
    if (doStrain) tile_wait();
    if (active) {
        double sp[6], pr = 0.;
        {
            if (doStrain) {
                double dv[9];
                gather_gradv(g, w, FN.V, tile, anchor, dv);
#pragma unroll
                for (int i = 0; i < 9; i++) dv[i] *= strainTime;
                PState s;
                load_state<SK>(P, p, s);
                constitutive_law<3, SK == SK_ELASTIC>(s, dv, strainTime, g.np, mats[P.mat[p]]);
                store_state<SK>(P, p, s);
#pragma unroll
                for (int i = 0; i < 6; i++) sp[i] = s.sp[i];
                pr = s.pressure;
            } else {
#pragma unroll
                for (int i = 0; i < 6; i++) sp[i] = P.sp[i][p];
                if (SK == SK_FULL) pr = P.pressure[p];
            }
        }
        const double nmp = -P.mp[p];            // f = -mp (sigma - p I) . grad S  (MatPoint3D.cpp:248-252)
        st.Q[lane][0] = nmp * (sp[XX] - pr); st.Q[lane][1] = nmp * (sp[YY] - pr); st.Q[lane][2] = nmp * (sp[ZZ] - pr);
        st.Q[lane][3] = nmp * sp[YZ]; st.Q[lane][4] = nmp * sp[XZ]; st.Q[lane][5] = nmp * sp[XY];
        if (FEXT) { st.Q[lane][6] = P.pfext[0][p]; st.Q[lane][7] = P.pfext[1][p]; st.Q[lane][8] = P.pfext[2][p]; st.Q[lane][9] = 0.; }
    }
    __syncwarp();
    double *dst[3] = {N.ftot[0], N.ftot[1], N.ftot[2]};
    constexpr int NQ = FEXT ? 10 : 6;
    warp_scatter<3, false>(g, key, active, dst, (int *)0, [&](int src, int i, int jk, double *acc) {
        const int j = jk % 3, k = jk / 3;
        const double Sx = st.X[i][src], Sy = st.Y[j][src], Sz = st.Z[k][src];
        const double gx = st.DX[i][src] * (Sy * Sz);
        const double gy = Sx * (st.DY[j][src] * Sz);
        const double gz = Sx * (Sy * st.DZ[k][src]);
        double q[NQ];
        load_payload<NQ>(st.Q, src, q);
        acc[0] += q[0] * gx + q[5] * gy + q[4] * gz;
        acc[1] += q[5] * gx + q[1] * gy + q[3] * gz;
        acc[2] += q[4] * gx + q[3] * gy + q[2] * gz;
        if (FEXT) {
            const double S = Sx * (Sy * Sz);
            acc[0] += S * q[6]; acc[1] += S * q[7]; acc[2] += S * q[8];
        }
        return (gx != 0. || gy != 0. || gz != 0.) ? 1 : 0;
    });
}

// ---- F3: particle update + P2G momentum with the new velocity ------------------------------------------
__global__ void __launch_bounds__(FUSED_THREADS, F3_MINB) k_f3_update_momentum(Grid g, Particles P, Nodes N, FusedNodes FN, const Material *mats,
                                                                      StepParams sp, int m, int doScatter)
{
    // the node tiles (gather) and the weight stage (scatter) are live one after the other: same shared memory
    union F3Shared { WarpTile t[2]; WarpStage<false, 4> st; };
    __shared__ F3Shared shared[FUSED_WARPS];
    WarpStage<false, 4> &st = shared[threadIdx.x >> 5].st;
    WarpTile &tv = shared[threadIdx.x >> 5].t[0], &ta = shared[threadIdx.x >> 5].t[1];
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = p < P.nNR;
    int key = 0;
    double xi[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
    if (active) {
        key = P.key[p];
        xi[0] = P.ncpos[0][p]; xi[1] = P.ncpos[1][p]; xi[2] = P.ncpos[2][p];
        load_lp(g, P, p, lp);
    }
    const int anchor = tile_anchor(key, active);
    tile_load_async(g, tv, anchor, FN.V);
    if (m <= 0) tile_load_async(g, ta, anchor, FN.A);
    Weights3 w;
    if (active) particle_weights<false>(g, key, xi, lp, w);
    tile_wait();
    double Svk[3] = {0., 0., 0.}, Sacc[3] = {0., 0., 0.};
    if (active) {
        const int d = key - anchor;
        const bool tiled = (unsigned)d <= (unsigned)(TILE_W - 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double syz = w.S[1][j] * w.S[2][k];
                const int row = w.center + (j - 1) * g.yplane + (k - 1) * g.zplane - 1;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double S = w.S[0][i] * syz;
                    const double4 v = tiled ? tv.r[j + 3 * k][d + i] : ldg4(&FN.V[row + i]);
                    Svk[0] += v.x * S; Svk[1] += v.y * S; Svk[2] += v.z * S;
                    if (m <= 0) {
                        const double4 a = tiled ? ta.r[j + 3 * k][d + i] : ldg4(&FN.A[row + i]);
                        Sacc[0] += a.x * S; Sacc[1] += a.y * S; Sacc[2] += a.z * S;
                    }
                }
            }
        }
    }
    __syncwarp();           // every lane is done with the tiles: reuse the memory for the scatter stage
    if (active) {
        if (doScatter) stage_weights(st, lane, w, P.mp[p]);
        const double dt = sp.dt;
        const double matDamp = mats[P.mat[p]].p[2];
        const double pAlpha = matDamp >= 0. ? matDamp : sp.particleAlpha;
        double vel[3] = {P.vel[0][p], P.vel[1][p], P.vel[2][p]};
        double pos[3] = {P.pos[0][p], P.pos[1][p], P.pos[2][p]};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double vm = Svk[c];
            if (m == 0) vm += Sacc[c] * (-dt);
            double Adamp0 = vel[c] * pAlpha;
            Adamp0 += vm * sp.gridAlpha;
            double delV;
            if (m > 0) {
                double delXRate = vel[c];
                vel[c] = vm;
                vel[c] += Adamp0 * (-dt);
                delV = vel[c] - delXRate;
                delXRate += vel[c];
                pos[c] += delXRate * (0.5 * dt);
            } else if (m == 0) {
                delV = (Sacc[c] - Adamp0) * dt;
                vel[c] += delV;
                double delXRate = vm + 0.5 * delV;
                pos[c] += delXRate * dt;
            } else {
                double delXRate = vel[c];
                vel[c] = Svk[c] - Adamp0 * dt;
                delV = vel[c] - delXRate;
                delXRate = vm + 0.5 * delV;
                pos[c] += delXRate * dt;
            }
            P.vel[c][p] = vel[c];
            P.pos[c][p] = pos[c];
            P.acc[c][p] = delV / dt;
        }
        if (doScatter) { st.Q[lane][0] = vel[0]; st.Q[lane][1] = vel[1]; st.Q[lane][2] = vel[2]; }
    }
    if (!doScatter) return;
    __syncwarp();
    double *dst[3] = {N.pk[0], N.pk[1], N.pk[2]};
    warp_scatter<3, false>(g, key, active, dst, (int *)0, [&](int src, int i, int jk, double *acc) {
        const double f = st.X[i][src] * st.YZ[jk][src];          // fn * mp
        const double2 vxy = *reinterpret_cast<const double2 *>(st.Q[src]);
        const double vz = st.Q[src][2];
        acc[0] += vxy.x * f; acc[1] += vxy.y * f; acc[2] += vz * f;
        return f != 0. ? 1 : 0;
    });
}

// ---- element reset as its own pass (slab mode) --------------------------------------------------------
// The reset needs only the updated positions, and the second strain update needs only the dual-cell key and
// ncpos of the start of the step, so in slab mode the reset runs BEFORE the strain kernel: the list of particles
// that left the slab is known one kernel earlier and the host does the migration handshake with its neighbours
// while the GPU is busy with the strain update.
__device__ __forceinline__ void reset_and_list(const Grid &g, const Particles &P, int p, StatusFlags *flags, double dt, const SlabInfo &slab)
{
    reset_element_one<3>(g, P, p, flags, dt);
    if (slab.on) {          // particle migration between slabs (replaces GridPatch::AddMovingParticle, GridPatch.cpp:214)
        const int k = (P.elem[p] - 1) / (g.horiz * g.vert);
        const int side = k < slab.cellLo ? 0 : (k >= slab.cellHi ? 1 : -1);
        if (side >= 0) {
            const int slot = atomicAdd(&slab.leaveCount[side], 1);
            if (slot < slab.leaveCap) slab.leaveIdx[side * slab.leaveCap + slot] = p;
        }
    }
}

__global__ void __launch_bounds__(256) k_reset_slab(Grid g, Particles P, StatusFlags *flags, double dt, SlabInfo slab)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P.n) reset_and_list(g, P, p, flags, dt, slab);
}

// ---- F4: second strain update + element reset --------------------------------------------------------
template <int SK>
__global__ void __launch_bounds__(FUSED_THREADS, F4_MINB) k_f4_strain_reset(Grid g, Particles P, FusedNodes FN, const Material *mats,
                                                                   double strainTime, int doStrain, int doReset, StatusFlags *flags, double dt, SlabInfo slab)
{
    __shared__ WarpTile tiles[FUSED_WARPS];
    WarpTile &tile = tiles[threadIdx.x >> 5];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (doStrain) {
        const bool active = p < P.nNR;
        int key = 0;
        double xi[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
        if (active) {
            prefetch_state<SK>(P, p);
            key = P.key[p];
            xi[0] = P.ncpos[0][p]; xi[1] = P.ncpos[1][p]; xi[2] = P.ncpos[2][p];
            load_lp(g, P, p, lp);
        }
        const int anchor = tile_anchor(key, active);
        tile_load_async(g, tile, anchor, FN.V);
        Weights3 w;
        if (active) particle_weights<true>(g, key, xi, lp, w);
        tile_wait();
        if (active) {
            double dv[9];
            gather_gradv(g, w, FN.V, tile, anchor, dv);
#pragma unroll
            for (int i = 0; i < 9; i++) dv[i] *= strainTime;
            PState s;               // (loading the state before the gather was measured slower: registers, not latency, limit this kernel)
            load_state<SK>(P, p, s);
            constitutive_law<3, SK == SK_ELASTIC>(s, dv, strainTime, g.np, mats[P.mat[p]]);
            store_state<SK>(P, p, s);
        }
    }
    if (doReset && p < P.n) reset_and_list(g, P, p, flags, dt, slab);
}

// ---- slab halo: pack partial sums of the three node planes shared with a neighbour / add the neighbour's ----
// which: 0 mass,pk (4 values)  1 ftot (3)  2 pk (3)  3 XPIC v*next sums (3).  Buffer layout [value][3 planes * planeNodes].
__global__ void k_halo_pack(int which, int node0, int count, Nodes N, double *buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int nd = node0 + i;
    if (which == 0) {
        buf[i] = N.mass[nd];
        buf[count + i] = N.pk[0][nd]; buf[2 * count + i] = N.pk[1][nd]; buf[3 * count + i] = N.pk[2][nd];
    } else if (which == 1) {
        buf[i] = N.ftot[0][nd]; buf[count + i] = N.ftot[1][nd]; buf[2 * count + i] = N.ftot[2][nd];
    } else if (which == 2) {
        buf[i] = N.pk[0][nd]; buf[count + i] = N.pk[1][nd]; buf[2 * count + i] = N.pk[2][nd];
    } else {
        buf[i] = N.vsn[0][nd]; buf[count + i] = N.vsn[1][nd]; buf[2 * count + i] = N.vsn[2][nd];
    }
}

__global__ void k_halo_add(int which, int node0, int count, Nodes N, const double *buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int nd = node0 + i;
    if (which == 0) {
        N.mass[nd] += buf[i];
        N.pk[0][nd] += buf[count + i]; N.pk[1][nd] += buf[2 * count + i]; N.pk[2][nd] += buf[3 * count + i];
    } else if (which == 1) {
        N.ftot[0][nd] += buf[i]; N.ftot[1][nd] += buf[count + i]; N.ftot[2][nd] += buf[2 * count + i];
    } else if (which == 2) {
        N.pk[0][nd] += buf[i]; N.pk[1][nd] += buf[count + i]; N.pk[2][nd] += buf[2 * count + i];
    } else {
        N.vsn[0][nd] += buf[i]; N.vsn[1][nd] += buf[count + i]; N.vsn[2][nd] += buf[2 * count + i];
    }
}

// ---- particle migration: rows of (nd doubles + ni ints bit-cast into ceil(ni/2) doubles) ----------------
__global__ void k_mig_pack(int nrows, const int *idx, size_t stride, int nd, const double *pool, int ni, const int *ipool, int rowLen, double *rows)
{
    const int r = blockIdx.x;
    if (r >= nrows) return;
    const int p = idx[r];
    for (int f = threadIdx.x; f < nd; f += blockDim.x) rows[(size_t)r * rowLen + f] = pool[(size_t)f * stride + p];
    int *irow = reinterpret_cast<int *>(rows + (size_t)r * rowLen + nd);
    for (int f = threadIdx.x; f < ni; f += blockDim.x) irow[f] = ipool[(size_t)f * stride + p];
}

__global__ void k_mig_unpack(int nrows, int first, size_t stride, int nd, double *pool, int ni, int *ipool, int rowLen, const double *rows)
{
    const int r = blockIdx.x;
    if (r >= nrows) return;
    const int p = first + r;
    for (int f = threadIdx.x; f < nd; f += blockDim.x) pool[(size_t)f * stride + p] = rows[(size_t)r * rowLen + f];
    const int *irow = reinterpret_cast<const int *>(rows + (size_t)r * rowLen + nd);
    for (int f = threadIdx.x; f < ni; f += blockDim.x) ipool[(size_t)f * stride + p] = irow[f];
}

// Bookkeeping of one migration on the device (one block): the n - L particles that stay must end up in slots
// [0, n-L).  sorted = all L leavers in ascending slot order.  Leavers below n-L are holes (the first *npairs entries of
// `sorted`); the slots [n-L, n) that are not leaving are the fillers, listed in ascending order.
#define MIG_PLAN_THREADS 1024
__global__ void __launch_bounds__(MIG_PLAN_THREADS) k_mig_plan(int n, int L, const int *sorted, int *fillers, int *npairs, int *flags)
{
    typedef cub::BlockScan<int, MIG_PLAN_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int base;
    const int nNew = n - L, tid = threadIdx.x;
    for (int s = tid; s < L; s += MIG_PLAN_THREADS) flags[s] = 0;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int t = tid; t < L; t += MIG_PLAN_THREADS) { const int q = sorted[t]; if (q >= nNew) flags[q - nNew] = 1; }
    __syncthreads();
    for (int c0 = 0; c0 < L; c0 += MIG_PLAN_THREADS) {
        const int s = c0 + tid;
        const int keep = (s < L && !flags[s]) ? 1 : 0;
        int pos, total;
        Scan(tmp).ExclusiveSum(keep, pos, total);
        if (keep) fillers[base + pos] = nNew + s;
        __syncthreads();
        if (tid == 0) base += total;
        __syncthreads();
    }
    if (tid == 0) *npairs = base;
}

// fill the holes left by departed particles with particles taken from the end of the arrays
__global__ void k_mig_fill(const int *npairs, const int *hole, const int *filler, size_t stride, int nd, double *pool, int ni, int *ipool)
{
    const int r = blockIdx.x;
    if (r >= *npairs) return;
    const int h = hole[r], q = filler[r];
    for (int f = threadIdx.x; f < nd; f += blockDim.x) pool[(size_t)f * stride + h] = pool[(size_t)f * stride + q];
    for (int f = threadIdx.x; f < ni; f += blockDim.x) ipool[(size_t)f * stride + h] = ipool[(size_t)f * stride + q];
}

// ---- node sweeps -------------------------------------------------------------------------------------
// (node_bcs / node_rigid_bcs: kernels_task.cuh)

// N1: tasks 3 + 4a.  pkc = pk; symmetry adjust; BCs(MASS_MOMENTUM) if a USF task exists; V = pk/mass
__global__ void k_n1_post_extrapolation(int n0, int nnodes, Nodes N, FusedNodes FN, VelBCs B, StepParams sp, int hasUSF)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + nnodes) return;
    double4 v = make_double4(0., 0., 0., 0.);
    // active node = some particle has a nonzero weight on it (numberPoints > 0 in the reference) = mass != 0, since
    // the weights and particle masses are positive; the flag is kept in N.cnt for the later sweeps
    const double mass = N.mass[i];
    N.cnt[i] = mass != 0. ? 1 : 0;
    if (mass != 0.) {
        double pk[3] = {N.pk[0][i], N.pk[1][i], N.pk[2][i]};
        double pkc[3] = {pk[0], pk[1], pk[2]};
        const int u = FN.bcOfNode ? FN.bcOfNode[i] : -1;
        if (u >= 0) {
            const int sd = B.symdir[u];
            if (sd & 32) pkc[0] = 0.;
            if (sd & 64) pkc[1] = 0.;
            if (sd & 128) pkc[2] = 0.;
        }
        if (hasUSF) {
            double ft[3] = {0., 0., 0.};
            bool changed = u >= 0;
            if (u >= 0) node_bcs(B, u, PASS_MASS_MOMENTUM, sp.dt, mass, pk, ft);
            if (FN.R.on) changed |= node_rigid_bcs(FN.R, i, PASS_MASS_MOMENTUM, sp.dt, mass, pk, ft);
            if (changed) { N.pk[0][i] = pk[0]; N.pk[1][i] = pk[1]; N.pk[2][i] = pk[2]; }
        }
        N.pkc[0][i] = pkc[0]; N.pkc[1][i] = pkc[1]; N.pkc[2][i] = pkc[2];
        if (mass != 0.) {
            const double rm = 1. / mass;
            v = make_double4(pk[0] * rm, pk[1] * rm, pk[2] * rm, 0.);
            N.vk[0][i] = v.x; N.vk[1][i] = v.y; N.vk[2][i] = v.z;
        }
    }
    FN.V[i] = v;
}

// N2: tasks 6 + 7 + 8a.  pk = pkc; ftot += m g; BCs(GRID_FORCES); pk += ftot dt; BCs(UPDATE_MOMENTUM);
// V = pk/mass; A = ftot/mass; then (when a re-extrapolation follows) pk = 0 for task 9a.
__global__ void k_n2_forces_momenta(int n0, int nnodes, Nodes N, FusedNodes FN, VelBCs B, StepParams sp, int rezero)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + nnodes) return;
    double4 v = make_double4(0., 0., 0., 0.), a = make_double4(0., 0., 0., 0.);
    if (N.cnt[i] > 0) {
        double pk[3] = {N.pkc[0][i], N.pkc[1][i], N.pkc[2][i]};
        double ft[3] = {N.ftot[0][i], N.ftot[1][i], N.ftot[2][i]};
        const double mass = N.mass[i];
        if (sp.hasGravity) { ft[0] += mass * sp.grav[0]; ft[1] += mass * sp.grav[1]; ft[2] += mass * sp.grav[2]; }
        const int u = FN.bcOfNode ? FN.bcOfNode[i] : -1;
        if (u >= 0) node_bcs(B, u, PASS_GRID_FORCES, sp.dt, mass, pk, ft);
        if (FN.R.on) node_rigid_bcs(FN.R, i, PASS_GRID_FORCES, sp.dt, mass, pk, ft);
        pk[0] += ft[0] * sp.dt; pk[1] += ft[1] * sp.dt; pk[2] += ft[2] * sp.dt;
        if (sp.xpicOrder <= 1) {
            if (u >= 0) node_bcs(B, u, PASS_UPDATE_MOMENTUM, sp.dt, mass, pk, ft);
            if (FN.R.on) node_rigid_bcs(FN.R, i, PASS_UPDATE_MOMENTUM, sp.dt, mass, pk, ft);
        }
        N.ftot[0][i] = ft[0]; N.ftot[1][i] = ft[1]; N.ftot[2][i] = ft[2];
        if (mass != 0.) {
            const double rm = 1. / mass;
            v = make_double4(pk[0] * rm, pk[1] * rm, pk[2] * rm, 0.);
            a = make_double4(ft[0] / mass, ft[1] / mass, ft[2] / mass, 0.);
            N.vk[0][i] = v.x; N.vk[1][i] = v.y; N.vk[2][i] = v.z;
        }
        if (rezero) { pk[0] = 0.; pk[1] = 0.; pk[2] = 0.; }
        N.pk[0][i] = pk[0]; N.pk[1][i] = pk[1]; N.pk[2][i] = pk[2];
    }
    FN.V[i] = v;
    FN.A[i] = a;
}

// N3: task 9b.  BCs(UPDATE_STRAINS_LAST); V = pk/mass
__global__ void k_n3_strains_last(int n0, int nnodes, Nodes N, FusedNodes FN, VelBCs B, StepParams sp)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + nnodes) return;
    double4 v = make_double4(0., 0., 0., 0.);
    if (N.cnt[i] > 0) {
        double pk[3] = {N.pk[0][i], N.pk[1][i], N.pk[2][i]};
        const double mass = N.mass[i];
        const int u = FN.bcOfNode ? FN.bcOfNode[i] : -1;
        {
            double ft[3] = {0., 0., 0.};
            bool changed = u >= 0;
            if (u >= 0) node_bcs(B, u, PASS_UPDATE_STRAINS_LAST, sp.dt, mass, pk, ft);
            if (FN.R.on) changed |= node_rigid_bcs(FN.R, i, PASS_UPDATE_STRAINS_LAST, sp.dt, mass, pk, ft);
            if (changed) { N.pk[0][i] = pk[0]; N.pk[1][i] = pk[1]; N.pk[2][i] = pk[2]; }
        }
        if (mass != 0.) {
            const double rm = 1. / mass;
            v = make_double4(pk[0] * rm, pk[1] * rm, pk[2] * rm, 0.);
            N.vk[0][i] = v.x; N.vk[1][i] = v.y; N.vk[2][i] = v.z;
        }
    }
    FN.V[i] = v;
}

// ---- XPIC(k) / FMPM(k), k > 1: XPICExtrapolationTask.cpp:49-161 in fused form ------------------------------------
// v*(1) = lumped velocity; for k = 2..m:  v*next_i = sum_p (mp S_ip / m_i) sum_j S_jp v*prev_j,  dv = v*prev - v*next,
// BCs zero the increment, v* += dv, v*prev = dv.  One node sweep to start, then per iteration one particle kernel
// (gather u_p from the VS records through the warp tile, warp-cooperative scatter of mp S_ip u_p) and one node sweep.
__global__ void k_nx_init(int n0, int nnodes, Nodes N, FusedNodes FN, double dt, int usingFMPM)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + nnodes) return;
    N.vsn[0][i] = 0.; N.vsn[1][i] = 0.; N.vsn[2][i] = 0.;
    double4 vs = make_double4(0., 0., 0., 0.);
    if (N.cnt[i] > 0) {
        const double mass = N.mass[i], rm = 1. / mass;
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (usingFMPM) {
                v[c] = N.pk[c][i] * rm;
                N.vsp[c][i] = v[c]; N.vk[c][i] = v[c];
            } else {            // XPIC: v*(1) is the velocity before the force increment (MatVelocityField.cpp:318-345)
                double t = N.pk[c][i];
                t += N.ftot[c][i] * (-dt);
                t *= rm;
                v[c] = t;
                N.vsp[c][i] = t;
                N.vk[c][i] = t + N.ftot[c][i] * (dt / mass);
            }
        }
        vs = make_double4(v[0], v[1], v[2], 0.);
    }
    FN.VS[i] = vs;
}

__global__ void __launch_bounds__(FUSED_THREADS, F3_MINB) k_fx_iterate(Grid g, Particles P, Nodes N, FusedNodes FN)
{
    union FxShared { WarpTile t; WarpStage<false, 4> st; };
    __shared__ FxShared shared[FUSED_WARPS];
    WarpStage<false, 4> &st = shared[threadIdx.x >> 5].st;
    WarpTile &tile = shared[threadIdx.x >> 5].t;
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = p < P.nNR;
    int key = 0;
    double xi[3] = {0., 0., 0.}, lp[3] = {0., 0., 0.};
    if (active) {
        key = P.key[p];
        xi[0] = P.ncpos[0][p]; xi[1] = P.ncpos[1][p]; xi[2] = P.ncpos[2][p];
        load_lp(g, P, p, lp);
    }
    const int anchor = tile_anchor(key, active);
    tile_load_async(g, tile, anchor, FN.VS);
    Weights3 w;
    if (active) particle_weights<false>(g, key, xi, lp, w);
    tile_wait();
    double u[3] = {0., 0., 0.};
    if (active) {
        const int d = key - anchor;
        const bool tiled = (unsigned)d <= (unsigned)(TILE_W - 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double syz = w.S[1][j] * w.S[2][k];
                const int row = w.center + (j - 1) * g.yplane + (k - 1) * g.zplane - 1;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double S = w.S[0][i] * syz;
                    const double4 v = tiled ? tile.r[j + 3 * k][d + i] : ldg4(&FN.VS[row + i]);
                    u[0] += S * v.x; u[1] += S * v.y; u[2] += S * v.z;
                }
            }
        }
    }
    __syncwarp();
    if (active) {
        stage_weights(st, lane, w, P.mp[p]);
        st.Q[lane][0] = u[0]; st.Q[lane][1] = u[1]; st.Q[lane][2] = u[2];
    }
    __syncwarp();
    double *dst[3] = {N.vsn[0], N.vsn[1], N.vsn[2]};
    warp_scatter<3, false>(g, key, active, dst, (int *)0, [&](int src, int i, int jk, double *acc) {
        const double f = st.X[i][src] * st.YZ[jk][src];          // fn * mp
        const double2 uxy = *reinterpret_cast<const double2 *>(st.Q[src]);
        const double uz = st.Q[src][2];
        acc[0] += uxy.x * f; acc[1] += uxy.y * f; acc[2] += uz * f;
        return f != 0. ? 1 : 0;
    });
}

// GET_DELTAV, BCs on the increment (XPIC_* pass of ZeroVelocityBC, MatVelocityField.cpp:529-541), UPDATE_VSTAR; after the
// last iteration the gather record V is v*(m)
__global__ void k_nx_finish(int n0, int nnodes, Nodes N, FusedNodes FN, VelBCs B, double dt, int particleUpdate, int usingFMPM, int last)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + nnodes) return;
    if (N.cnt[i] == 0) {
        FN.VS[i] = make_double4(0., 0., 0., 0.);
        if (last) FN.V[i] = make_double4(0., 0., 0., 0.);
        return;
    }
    const double mass = N.mass[i];
    double d[3] = {N.vsp[0][i] - N.vsn[0][i] / mass, N.vsp[1][i] - N.vsn[1][i] / mass, N.vsp[2][i] - N.vsn[2][i] / mass};
    const bool addReaction = particleUpdate && !usingFMPM;
    double ft[3] = {N.ftot[0][i], N.ftot[1][i], N.ftot[2][i]};
    const int u = FN.bcOfNode ? FN.bcOfNode[i] : -1;
    if (u >= 0) {
        for (int e = B.start[u]; e < B.start[u + 1]; e++) {
            if (!B.active[e]) continue;
            const double nx = B.norm[3 * e], ny = B.norm[3 * e + 1], nz = B.norm[3 * e + 2];
            const double dotn = d[0] * nx + d[1] * ny + d[2] * nz;
            d[0] += nx * (-dotn); d[1] += ny * (-dotn); d[2] += nz * (-dotn);
            if (particleUpdate && B.reaction) {      // lumped-mass addition to the BC's freaction (MatVelocityField.cpp:532-536)
                const double s = -mass * dotn / dt;
                const double r[3] = {nx * s, ny * s, nz * s};
                add_reaction(B.reaction + 3 * e, r);
            }
            if (addReaction) {
                const double s = -mass * dotn / dt;
                ft[0] += nx * s; ft[1] += ny * s; ft[2] += nz * s;
            }
        }
    }
    if (FN.R.on) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (FN.R.owner[c][i] == RIGID_NONE) continue;
            const double dotn = d[c];
            d[c] += -dotn;
            if (particleUpdate && FN.R.reaction && dotn != 0.) atomicAdd(FN.R.reaction + 3 * FN.R.mat[FN.R.owner[c][i]] + c, -mass * dotn / dt);
            if (addReaction) ft[c] += -mass * dotn / dt;
        }
    }
    if (addReaction) { N.ftot[0][i] = ft[0]; N.ftot[1][i] = ft[1]; N.ftot[2][i] = ft[2]; }
    double vk[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        N.vsp[c][i] = d[c];
        vk[c] = N.vk[c][i] + d[c];
        N.vk[c][i] = vk[c];
        N.vsn[c][i] = 0.;
    }
    FN.VS[i] = make_double4(d[0], d[1], d[2], 0.);
    if (last) FN.V[i] = make_double4(vk[0], vk[1], vk[2], 0.);
}

// lumped grid velocity V = pk/mass (strain update after the particle update when no re-extrapolation follows and the
// V records still hold v*(m) of an XPIC particle update: MatVelocityField::GridValueCalculation :239-251)
__global__ void k_n_grid_velocity(int n0, int nnodes, Nodes N, FusedNodes FN)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + nnodes) return;
    double4 v = make_double4(0., 0., 0., 0.);
    const double mass = N.mass[i];
    if (N.cnt[i] > 0 && mass != 0.) {
        const double rm = 1. / mass;
        v = make_double4(N.pk[0][i] * rm, N.pk[1][i] * rm, N.pk[2][i] * rm, 0.);
        N.vk[0][i] = v.x; N.vk[1][i] = v.y; N.vk[2][i] = v.z;
    }
    FN.V[i] = v;
}

// ---- physical sort by dual cell -----------------------------------------------------------------------
// lead = half a sort interval: the particles are ordered by the dual cell they will be in halfway to the next sort, so
// the order is most accurate in the middle of the interval instead of decaying from the sort onwards (the kernels
// group by the ACTUAL dual cell, so the order only affects speed)
__global__ void k_sort_keys(Grid g, Particles P, int *keys, int *idx, double lead)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    int key;
    if (p < P.nNR) {
        int e = P.elem[p];
        double pos[3] = {P.pos[0][p], P.pos[1][p], P.pos[2][p]};
        if (lead > 0.) {
            double ahead[3] = {pos[0] + lead * P.vel[0][p], pos[1] + lead * P.vel[1][p], pos[2] + lead * P.vel[2][p]};
            const int e2 = find_element_from_point<3>(g, ahead);
            if (e2 > 0 && !edge_element<3>(g, e2)) { e = e2; pos[0] = ahead[0]; pos[1] = ahead[1]; pos[2] = ahead[2]; }
        }
        double xi[3];
        get_xipos<3>(g, e, pos, xi);
        const ElemIJK c = elem_ijk(g, e);
        key = elem_node0(g, c) + (xi[0] < 0. ? 0 : 1) + (xi[1] < 0. ? 0 : 1) * g.yplane + (xi[2] < 0. ? 0 : 1) * g.zplane;
    } else {
        key = g.nnodes + (p - P.nNR);        // rigid particles keep their place after the nonrigid block
    }
    keys[p] = key;
    idx[p] = p;
}

// live: bit f set = double field f carries information at sort time (fields that are recomputed before their next
// read, or that the materials in use never touch, stay behind: both pools hold the same zeros there)
__global__ void k_permute_pool(int n, size_t stride, int nd, unsigned long long live, const double *__restrict__ src, double *__restrict__ dst,
                               int ni, unsigned ilive, const int *__restrict__ isrc, int *__restrict__ idst, const int *__restrict__ perm)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int q = perm[p];
    for (int f = 0; f < nd; f++) if (live >> f & 1ull) dst[(size_t)f * stride + p] = src[(size_t)f * stride + q];
    for (int f = 0; f < ni; f++) if (ilive >> f & 1u) idst[(size_t)f * stride + p] = isrc[(size_t)f * stride + q];
}
