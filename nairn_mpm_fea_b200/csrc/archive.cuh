// Output side of the step, on the device: the reference's binary particle-archive record
// (ArchiveData::ArchiveResults, NairnMPM/src/System/ArchiveData.cpp:806-1100; record size CalcArchiveSize :328-396) packed
// straight from the SoA state, and the raw sums behind the common GlobalQuantity rows
// (Global_Quantities/GlobalQuantity.cpp:394-1075), so that an archive step moves one record block or a few hundred bytes
// instead of the whole particle state.
//
// Depends on mpm_types.cuh only: tests/devlaws compiles this header for the host and checks the records byte for byte against
// nairn_mpm_fea_b200/archive.py (which is byte-identical to the reference CLI's files) without a GPU.
#pragma once
#include "mpm_types.cuh"

// <MPMArchiveOrder> byte positions (ArchiveData.hpp:22-32, enum starting at ARCH_Velocity = 2)
enum { ARCH_Velocity = 2, ARCH_Stress, ARCH_Strain, ARCH_PlasticStrain, ARCH_OldOrigPosition, ARCH_WorkEnergy, ARCH_DeltaTemp,
       ARCH_PlasticEnergy, ARCH_ver2Empty, ARCH_ShearComponents, ARCH_StrainEnergy, ARCH_History, ARCH_Concentration,
       ARCH_HeatEnergy, ARCH_ElementCrossings, ARCH_RotStrain, ARCH_DamageNormal, ARCH_SpinMomentum, ARCH_SpinVelocity,
       ARCH_History59, ARCH_History1014, ARCH_History1519, ARCH_Size };
#define ARCH_MAXMPMITEMS 24

struct ArchiveLayout {
    int dim;                 // 2 or 3
    unsigned items;          // bit b set: order[b] == 'Y' (history: any of its four)
    unsigned histMask;       // bit k: history variable k+1 is archived
    int recWords;            // record size in 32-bit words
    double thickness;        // 2D: the particles' thickness (uniform)
    const double *origpos;   // [3][nTotal] in the caller's order (never permuted: indexed through Particles::orig), or NULL
    const double *angles0;   // [3][nTotal] initial material angles z, y, x in radians, or NULL for 0
    size_t stride;           // nTotal
};

// Record size for an order string, following CalcArchiveSize; -1 when the order asks for an item this path does not produce
__host__ __device__ inline int archive_layout_from_order(const char *order, int dim, ArchiveLayout &L)
{
    char o[ARCH_MAXMPMITEMS];
    int len = 0;
    while (len < ARCH_MAXMPMITEMS && order[len] != 0) { o[len] = order[len]; len++; }
    for (int i = len; i < ARCH_MAXMPMITEMS; i++) o[i] = 'N';
    const int unsupported[8] = {ARCH_ShearComponents, ARCH_DamageNormal, ARCH_SpinMomentum, ARCH_SpinVelocity, ARCH_History59,
                                ARCH_History1014, ARCH_History1519, ARCH_Size};
    for (int i = 0; i < 8; i++) if (unsupported[i] < ARCH_MAXMPMITEMS && o[unsupported[i]] != 'N') return -1;
    L.dim = dim; L.items = 0; L.histMask = 0;
    for (int b = ARCH_Velocity; b < ARCH_MAXMPMITEMS; b++) if (o[b] == 'Y') L.items |= 1u << b;
    const char h = o[ARCH_History];
    if (h == 'Y') L.histMask = 1;
    else if (h != 'N') L.histMask = (unsigned)h & 15u;
    if (L.histMask) L.items |= 1u << ARCH_History; else L.items &= ~(1u << ARCH_History);
    const int nt = dim == 3 ? 6 : 4;
    int w = 1 + 2 + 1;                                   // element, mass, material + padding
    w += 2 * (dim == 3 ? 3 : 2);                         // angles (3D) or angle + thickness (2D)
    w += 2 * dim * 2;                                    // position, original position
    if (L.items & (1u << ARCH_Velocity)) w += 2 * dim;
    if (L.items & (1u << ARCH_Stress)) w += 2 * nt;
    if (L.items & (1u << ARCH_Strain)) w += 2 * nt;
    if (L.items & (1u << ARCH_PlasticStrain)) w += 2 * nt;
    if (L.items & (1u << ARCH_WorkEnergy)) w += 2;
    if (L.items & (1u << ARCH_DeltaTemp)) w += 2;
    if (L.items & (1u << ARCH_PlasticEnergy)) w += 2;
    if (L.items & (1u << ARCH_StrainEnergy)) w += 2;
    for (int k = 0; k < 4; k++) if (L.histMask & (1u << k)) w += 2;
    if (L.items & (1u << ARCH_Concentration)) w += 2 * (dim + 1);
    if (L.items & (1u << ARCH_HeatEnergy)) w += 2;
    if (L.items & (1u << ARCH_ElementCrossings)) w += 1;
    if (L.items & (1u << ARCH_RotStrain)) w += 2 * (dim == 3 ? 3 : 1);
    L.recWords = w;
    return w * 4;
}

struct RecordWriter {
    uint32_t *w;
    __host__ __device__ inline void i32(int v) { *w++ = (uint32_t)v; }
    __host__ __device__ inline void f64(double v)
    {   // doubles sit at 4-byte offsets inside the record (the element number comes first): two little-endian words
        union { double d; uint32_t u[2]; } c;
        c.d = v;
        *w++ = c.u[0]; *w++ = c.u[1];
    }
};

// One particle's record at `out` (recWords words).  slot = the particle's index in the caller's order.
__host__ __device__ inline void archive_record(const Particles &P, int p, int slot, const Material *mats, const ArchiveLayout &L, uint32_t *out)
{
    RecordWriter r;
    r.w = out;
    const int dim = L.dim;
    const Material &m = mats[P.mat[p]];
    const double mp = P.mp[p];
    const double pi = 3.141592653589793;         // PI_CONSTANT; operation order of MPMBase::GetRotation* (MPMBase.cpp:601-619)
    r.i32(P.elem[p]);
    r.f64(mp);
    r.i32((P.mat[p] + 1) & 0xffff);              // short material number + two zero bytes
    // deformation gradient -> strain and rotation strain (MatPoint3D::SetDeformationGradientMatrix, MatPoint3D.cpp:320-336)
    const double F0 = P.F[0][p], F1 = P.F[1][p], F2 = P.F[2][p], F3 = P.F[3][p], F4 = P.F[4][p], F5 = P.F[5][p], F6 = P.F[6][p],
                 F7 = P.F[7][p], F8 = P.F[8][p];
    const double wxy = F3 - F1, wxz = dim == 3 ? F6 - F2 : 0., wyz = dim == 3 ? F7 - F5 : 0.;
    double a0[3] = {0., 0., 0.};
    if (L.angles0) { a0[0] = L.angles0[slot]; a0[1] = L.angles0[L.stride + slot]; a0[2] = L.angles0[2 * L.stride + slot]; }
    if (dim == 3) {
        r.f64(180.0 * (a0[0] - 0.5 * wxy) / pi); r.f64(180.0 * (a0[1] + 0.5 * wxz) / pi); r.f64(180.0 * (a0[2] - 0.5 * wyz) / pi);
    } else {
        r.f64(180.0 * (a0[0] - 0.5 * wxy) / pi);
        r.f64(L.thickness);
    }
    for (int c = 0; c < dim; c++) r.f64(P.pos[c][p]);
    for (int c = 0; c < dim; c++) r.f64(L.origpos ? L.origpos[(size_t)c * L.stride + slot] : P.pos[c][p]);
    if (L.items & (1u << ARCH_Velocity)) for (int c = 0; c < dim; c++) r.f64(P.vel[c][p]);
    // record tensor order xx yy zz xy [xz yz]; state order xx yy zz yz xz xy
    const int tens[6] = {0, 1, 2, 5, 4, 3};
    const int nt = dim == 3 ? 6 : 4;
    if (L.items & (1u << ARCH_Stress)) {
        // Cauchy stress = rho * specific stress, rho = rho0 / relative volume (1 unless the material tracks J: Neohookean.cpp:374-376);
        // materials that keep the pressure apart add it back (MaterialBase::GetStressPandDev, MaterialBaseMPM.cpp:1635-1641)
        const double relvol = (m.kind == MAT_NEOHOOKEAN || m.kind == MAT_MOONEY) ? P.hist[0][p] : 1.0;
        const double rho = m.p[0] / relvol;
        const bool pand = m.kind == MAT_NEOHOOKEAN || m.kind == MAT_ISOPLASTICITY || m.kind == MAT_MOONEY;
        const double pr = P.pressure[p];
        for (int i = 0; i < nt; i++) {
            double s = P.sp[tens[i]][p];
            if (pand && tens[i] < 3) s = s - pr;
            r.f64(rho * s);
        }
    }
    if (L.items & (1u << ARCH_Strain)) {
        const double e[6] = {F0 - 1., F4 - 1., F8 - 1., dim == 3 ? F7 + F5 : 0., dim == 3 ? F6 + F2 : 0., F3 + F1};
        for (int i = 0; i < nt; i++) r.f64(e[tens[i]]);
    }
    if (L.items & (1u << ARCH_PlasticStrain)) for (int i = 0; i < nt; i++) r.f64(P.eplast[tens[i]][p]);
    const double sc = 1.0e-9 * mp;
    if (L.items & (1u << ARCH_WorkEnergy)) r.f64(sc * P.work[p]);
    if (L.items & (1u << ARCH_DeltaTemp)) r.f64(P.temp ? P.temp[p] : P.prevT[p]);        // pTemperature (ArchiveData.cpp:976); without conduction it equals prevT
    if (L.items & (1u << ARCH_PlasticEnergy)) r.f64(sc * P.plast[p]);
    if (L.items & (1u << ARCH_StrainEnergy)) { const double se = P.work[p] - P.res[p]; r.f64(sc * se); }
    for (int k = 0; k < 4; k++) if (L.histMask & (1u << k)) r.f64(P.hist[k][p]);
    if (L.items & (1u << ARCH_Concentration)) for (int c = 0; c < dim + 1; c++) r.f64(0.);      // no transport on this path
    if (L.items & (1u << ARCH_HeatEnergy)) r.f64(sc * P.heat[p]);
    if (L.items & (1u << ARCH_ElementCrossings)) { const int x = P.cross[p]; r.i32(x < 0 ? -x : x); }
    if (L.items & (1u << ARCH_RotStrain)) for (int c = 0; c < (dim == 3 ? 3 : 1); c++) r.f64(180.0 * a0[c] / pi);
}

// ---- raw sums behind the GlobalQuantity rows (GlobalQuantity.cpp:394-1075), per material ----------------------------------
// The host divides (volume-weighted averages: sum / sum of Vp) and applies the unit scalings.
enum { GS_MASS = 0,          // sum mp
       GS_VOLUME = 1,        // sum Vp,  Vp = J mp / rho0 (J = GetCurrentRelativeVolume)
       GS_LINMOM = 2,        // 2..4   sum mp v           (LINMOMX/Y/Z :1021-1050)
       GS_KINETIC = 5,       // sum mp |v|^2 / 2          (KINE_ENERGY :619-625)
       GS_WORK = 6,          // sum mp workEnergy         (WORK_ENERGY)
       GS_STRAIN_ENERGY = 7, // sum mp (work - residual)  (STRAIN_ENERGY)
       GS_HEAT = 8,          // sum mp heatEnergy         (HEAT_ENERGY)
       GS_ENTROPY = 9,       // sum mp entropy            (ENTROPY_ENERGY)
       GS_PLASTIC = 10,      // sum mp plastEnergy        (PLAS_ENERGY :665-675)
       GS_STRESS = 11,       // 11..16 sum mp (total specific stress) xx yy zz yz xz xy   (AVG_Sij :408-437: divide by volume)
       GS_VOL_VEL = 17,      // 17..19 sum Vp v           (AVG_VELX/Y/Z :677-723)
       GS_VOL_F = 20,        // 20..28 sum Vp F, row-major (AVG_Fij :568-600)
       GS_NSUMS = 29 };

__host__ __device__ inline void global_summands(const Particles &P, int p, const Material &m, int dim, double q[GS_NSUMS])
{
    const double mp = P.mp[p];
    const double relvol = (m.kind == MAT_NEOHOOKEAN || m.kind == MAT_MOONEY) ? P.hist[0][p] : 1.0;
    const double Vp = relvol * mp / m.p[0];
    const double vx = P.vel[0][p], vy = P.vel[1][p], vz = dim == 3 ? P.vel[2][p] : 0.;
    q[GS_MASS] = mp;
    q[GS_VOLUME] = Vp;
    q[GS_LINMOM] = mp * vx; q[GS_LINMOM + 1] = mp * vy; q[GS_LINMOM + 2] = mp * vz;
    q[GS_KINETIC] = 0.5 * mp * (vx * vx + vy * vy) + (dim == 3 ? 0.5 * mp * (vz * vz) : 0.);
    q[GS_WORK] = mp * P.work[p];
    q[GS_STRAIN_ENERGY] = mp * (P.work[p] - P.res[p]);
    q[GS_HEAT] = mp * P.heat[p];
    q[GS_ENTROPY] = mp * P.entropy[p];
    q[GS_PLASTIC] = mp * P.plast[p];
    const bool pand = m.kind == MAT_NEOHOOKEAN || m.kind == MAT_ISOPLASTICITY || m.kind == MAT_MOONEY;
    const double pr = pand ? P.pressure[p] : 0.;
    for (int c = 0; c < 6; c++) q[GS_STRESS + c] = mp * (c < 3 ? P.sp[c][p] - pr : P.sp[c][p]);
    q[GS_VOL_VEL] = Vp * vx; q[GS_VOL_VEL + 1] = Vp * vy; q[GS_VOL_VEL + 2] = Vp * vz;
    for (int i = 0; i < 9; i++) q[GS_VOL_F + i] = Vp * P.F[i][p];
}

#ifdef __CUDACC__
#define ARCHIVE_THREADS 256

// records of one particle set (non-rigid or rigid-BC) into the caller-ordered record block
// (slot == NULL: device order, record base + p)
__global__ void __launch_bounds__(ARCHIVE_THREADS) k_pack_archive(int cnt, Particles P, const int *slot, int base, const Material *mats, ArchiveLayout L,
                                                                   uint32_t *records)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cnt) return;
    const int o = slot ? slot[p] : base + p;
    archive_record(P, p, o, mats, L, records + (size_t)o * L.recWords);
}

// partial[(m * gridDim.x + block) * GS_NSUMS + k]: block-wise sums for material m = blockIdx.y, fixed order (no atomics), so the
// totals are the same from run to run
__global__ void __launch_bounds__(ARCHIVE_THREADS) k_global_partial(Particles P, const Material *mats, int dim, double *partial)
{
    const int m = blockIdx.y;
    double acc[GS_NSUMS];
#pragma unroll
    for (int k = 0; k < GS_NSUMS; k++) acc[k] = 0.;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P.nNR; p += gridDim.x * blockDim.x) {
        if (P.mat[p] != m) continue;
        double q[GS_NSUMS];
        global_summands(P, p, mats[m], dim, q);
#pragma unroll
        for (int k = 0; k < GS_NSUMS; k++) acc[k] += q[k];
    }
    __shared__ double sh[ARCHIVE_THREADS / 32][GS_NSUMS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < GS_NSUMS; k++) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < GS_NSUMS) {
        double v = 0.;
        for (int w = 0; w < ARCHIVE_THREADS / 32; w++) v += sh[w][threadIdx.x];
        partial[((size_t)m * gridDim.x + blockIdx.x) * GS_NSUMS + threadIdx.x] = v;
    }
}

__global__ void k_global_final(int nmat, int nblk, const double *partial, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmat * GS_NSUMS) return;
    const int m = i / GS_NSUMS, k = i % GS_NSUMS;
    double v = 0.;
    for (int b = 0; b < nblk; b++) v += partial[((size_t)m * nblk + b) * GS_NSUMS + k];
    out[i] = v;
}
#endif
