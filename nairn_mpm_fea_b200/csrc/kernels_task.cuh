// Per-task kernels: one kernel (or a node sweep) per reference MPMTask, particles in any order,
// node accumulation by global FP64 atomics.  This is the general path (every shape function,
// 2D and 3D, every material); the cell-sorted fused kernels in kernels_fused.cuh replace the
// particle<->grid transfers of the 3D uGIMP step when the configuration is eligible.
#pragma once
#include "mpm_types.cuh"
#include "shape.cuh"
#include "materials.cuh"

#define TASK_THREADS 128

__device__ __forceinline__ void load_xi_lp(const Particles &P, int p, double xi[3], double lp[3])
{
    xi[0] = P.ncpos[0][p]; xi[1] = P.ncpos[1][p]; xi[2] = P.ncpos[2][p];
    lp[0] = P.lp[0][p]; lp[1] = P.lp[1][p]; lp[2] = P.lp[2][p];
}

// nodes of particle p for the configured shape function; in multimaterial mode the node index handed to f is the
// particle's own material velocity field of that node (field-major node arrays, mpm_types.cuh)
template <int DIM, int SHAPE, bool GRAD, class F>
__device__ __forceinline__ void particle_nodes_one_field(const Grid &g, const Particles &P, int p, F &&f);

template <int DIM, int SHAPE, bool GRAD, class F>
__device__ __forceinline__ void particle_nodes(const Grid &g, const Particles &P, int p, F &&f)
{
    const int off = P.foff ? P.foff[p] : 0;
    particle_nodes_one_field<DIM, SHAPE, GRAD>(g, P, p, [&](int nd, double S, double gx, double gy, double gz) { f(nd + off, S, gx, gy, gz); });
}

template <int DIM, int SHAPE, bool GRAD, class F>
__device__ __forceinline__ void particle_nodes_one_field(const Grid &g, const Particles &P, int p, F &&f)
{
    if (SHAPE_IS_MERGED(SHAPE)) {
        for_each_node_cpdi_merged<DIM, SHAPE, GRAD>(g, P, p, f);
    } else if (SHAPE_IS_CPDI(SHAPE)) {
        for_each_node_cpdi<DIM, SHAPE, GRAD>(g, P, p, f);
    } else {
        double xi[3], lp[3];
        load_xi_lp(P, p, xi, lp);
        for_each_node<DIM, SHAPE, GRAD>(g, P.elem[p], xi, lp, f);
    }
}

// ---- task 1: InitializationTask (InitializationTask.cpp:49-85) ------------------------------
// node zeroing is a memset; this is the particle half: ncpos = GetXiPos(pos) in the current element
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_init_particles(Grid g, Particles P, StatusFlags *flags)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    int e = P.elem[p];
    double pos[3] = {P.pos[0][p], P.pos[1][p], DIM == 3 ? P.pos[2][p] : 0.};
    double xi[3];
    get_xipos<DIM>(g, e, pos, xi);
    P.ncpos[0][p] = xi[0]; P.ncpos[1][p] = xi[1]; P.ncpos[2][p] = xi[2];
    if (SHAPE_IS_CPDI(SHAPE) && p < P.nNR) {     // ElementBase::GetShapeFunctionData (MoreMPMElementBase.cpp:50-58)
        // (merged instantiation = every kernel of the step walks the merged window: the lean set-up will do)
        if (!cpdi_setup<DIM, SHAPE, DIM == 3 && SHAPE == SHAPE_LCPDI_MERGED>(g, P, p)) atomicCAS(&flags->cpdiLeft, 0, P.orig[p] + 1);
    }
}

// ---- task 2: MassAndMomentumTask (MassAndMomentumTask.cpp:62-98, NodalPointMPM.cpp:419-453) ---
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_mass_momentum(Grid g, Particles P, Nodes N)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const double mp = P.mp[p];
    const double vx = P.vel[0][p], vy = P.vel[1][p], vz = DIM == 3 ? P.vel[2][p] : 0.;
    particle_nodes<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        const double fnmp = S * mp;
        atomAdd(&N.pk[0][nd], vx * fnmp);
        atomAdd(&N.pk[1][nd], vy * fnmp);
        if (DIM == 3) atomAdd(&N.pk[2][nd], vz * fnmp);
        atomAdd(&N.mass[nd], fnmp);
        atomicAdd(&N.cnt[nd], 1);
    });
}

// ---- task 3: PostExtrapolationTask node pass (MatVelocityField.cpp:158-164) ------------------
// (fields without nonrigid points -- empty ones, and a rigid contact material's field in multimaterial mode -- keep the zeros of
// MatVelocityField::Zero: CrackVelocityFieldMulti::GetTotalMassAndCount copies nonrigid fields only)
__global__ void k_copy_momenta(int nnodes, Nodes N)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes || N.cnt[i] <= 0) return;
    N.pkc[0][i] = N.pk[0][i]; N.pkc[1][i] = N.pk[1][i]; N.pkc[2][i] = N.pk[2][i];
}

// ---- grid velocity BCs (NodalVelBC.cpp:321-400, MatVelocityField.cpp:490-575) ----------------
// zero pass of one BC with unit direction n (ZeroVelocityBC / SetFtotDirection / ZeroMomentumBC)
// react (grid-forces pass only): where the force the BC adds is summed too (NodalVelBC::freaction, MatVelocityField.cpp:519-524,:563-569)
__device__ __forceinline__ void bc_zero(int pass, double dt, double nx, double ny, double nz, double pk[3], double ft[3], double *react = nullptr)
{
    if (pass == PASS_GRID_FORCES) {
        double dotf = ft[0] * nx + ft[1] * ny + ft[2] * nz;
        double dotp = pk[0] * nx + pk[1] * ny + pk[2] * nz;
        double s = -dotf - dotp / dt;
        ft[0] += nx * s; ft[1] += ny * s; ft[2] += nz * s;
        if (react) { react[0] += nx * s; react[1] += ny * s; react[2] += nz * s; }
    } else {
        double dotn = pk[0] * nx + pk[1] * ny + pk[2] * nz;
        pk[0] += nx * (-dotn); pk[1] += ny * (-dotn); pk[2] += nz * (-dotn);
        if (pass == PASS_UPDATE_MOMENTUM) {
            double s = -dotn / dt;
            ft[0] += nx * s; ft[1] += ny * s; ft[2] += nz * s;
        }
    }
}

// add pass (AddVelocityBC / AddFtotDirection / AddMomentumBC)
__device__ __forceinline__ void bc_add(int pass, double dt, double mass, double vel, double nx, double ny, double nz, double pk[3], double ft[3], double *react = nullptr)
{
    if (pass == PASS_GRID_FORCES) {
        double s = mass * vel / dt;
        ft[0] += nx * s; ft[1] += ny * s; ft[2] += nz * s;
        if (react) { react[0] += nx * s; react[1] += ny * s; react[2] += nz * s; }
    } else {
        double pvel = mass * vel;
        pk[0] += nx * pvel; pk[1] += ny * pvel; pk[2] += nz * pvel;
        if (pass == PASS_UPDATE_MOMENTUM) {
            double s = pvel / dt;
            ft[0] += nx * s; ft[1] += ny * s; ft[2] += nz * s;
        }
    }
}

// one BC's share of the reaction force; the material fields of a node (multimaterial mode) add into the same entry
__device__ __forceinline__ void add_reaction(double *dst, const double r[3])
{
#pragma unroll
    for (int d = 0; d < 3; d++)
        if (r[d] != 0.) atomicAdd(dst + d, r[d]);
}

// BC application for one node: its entries are walked in list order, first the zero pass over all of
// them, then the add pass (VelocityBCLoop).  A BC next to a symmetry plane REFLECTS the velocity of the node on the other
// side of the plane (NodalVelBC::AddVelocityBC -> NodalPoint::ReflectVelocityBC, NodalVelBC.cpp:196-210,
// CrackVelocityFieldSingle.cpp:133-144): v = v0 + ratio (v0 - n.pk_r / m_r) when that node has particles, else v0.
// Only the per-task kernels pass N (the reflected node's momentum has to be complete: its own BCs act on other
// components); contexts with reflected BCs do not use the fused node sweeps.
__device__ __forceinline__ void node_bcs(const VelBCs &B, int u, int pass, double dt, double mass, double pk[3], double ft[3], const Nodes *N = nullptr, int off = 0)
{
    const int e0 = B.start[u], e1 = B.start[u + 1];
    const bool track = B.reaction != nullptr && pass == PASS_GRID_FORCES;
    for (int e = e0; e < e1; e++) {
        if (!B.active[e]) continue;
        double r[3] = {0., 0., 0.};
        bc_zero(pass, dt, B.norm[3 * e], B.norm[3 * e + 1], B.norm[3 * e + 2], pk, ft, track ? r : nullptr);
        if (track) add_reaction(B.reaction + 3 * e, r);
    }
    for (int e = e0; e < e1; e++) {
        if (!B.active[e]) continue;
        const double nx = B.norm[3 * e], ny = B.norm[3 * e + 1], nz = B.norm[3 * e + 2];
        double v = B.value[e];
        if (N && B.refl && B.refl[e] >= 0) {
            const int r = B.refl[e] + off;      // the same material field of the node across the plane (CrackVelocityFieldMulti::ReflectVelocityBC)
            if (N->cnt[r] > 0) {
                const double dotn = nx * N->pk[0][r] + ny * N->pk[1][r] + nz * N->pk[2][r];
                v = v + B.reflRatio[e] * (v - dotn / N->mass[r]);
            }
        }
        double r[3] = {0., 0., 0.};
        bc_add(pass, dt, mass, v, nx, ny, nz, pk, ft, track ? r : nullptr);
        if (track) add_reaction(B.reaction + 3 * e, r);
    }
}

// The BCs rigid particles made on this node.  They sit after the grid BCs in the reference's list and
// only on dofs no grid BC fixes (ProjectRigidBCsTask.cpp:202), each along one axis, so applying them
// after the node's grid BCs gives the same sums as the reference's zero-all-then-add-all walk.
__device__ __forceinline__ double rigid_bc_velocity(const RigidBCs &R, const Nodes &N, int nd, int d, int owner, bool &skip);

template <bool MIRROR = false>
__device__ __forceinline__ bool node_rigid_bcs(const RigidBCs &R, int nd, int pass, double dt, double mass, double pk[3], double ft[3], const Nodes *N = nullptr)
{
    int o[3];
    bool any = false;
#pragma unroll
    for (int d = 0; d < 3; d++) { o[d] = R.owner[d][nd]; any |= o[d] != RIGID_NONE; }
    if (!any) return false;
    const bool track = R.reaction != nullptr && pass == PASS_GRID_FORCES;
    double r[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}};
#pragma unroll
    for (int d = 0; d < 3; d++)
        if (o[d] != RIGID_NONE) bc_zero(pass, dt, d == 0 ? 1. : 0., d == 1 ? 1. : 0., d == 2 ? 1. : 0., pk, ft, track ? r[d] : nullptr);
#pragma unroll
    for (int d = 0; d < 3; d++)
        if (o[d] != RIGID_NONE) {
            double v = R.vel[d][o[d]];
            if (MIRROR) { bool skip; v = rigid_bc_velocity(R, *N, nd, d, o[d], skip); if (skip) continue; }
            bc_add(pass, dt, mass, v, d == 0 ? 1. : 0., d == 1 ? 1. : 0., d == 2 ? 1. : 0., pk, ft, track ? r[d] : nullptr);
        }
    if (track) {        // the BC's ID is its rigid particle's material (ProjectRigidBCsTask.cpp:241)
#pragma unroll
        for (int d = 0; d < 3; d++)
            if (o[d] != RIGID_NONE) add_reaction(R.reaction + 3 * R.mat[o[d]], r[d]);
    }
    return true;
}

// dof d of node nd is fixed by a grid BC or by a rigid particle this step (NodalPoint::fixedDirection)
__device__ __forceinline__ bool rigid_dof_fixed(const RigidBCs &R, int nd, int d)
{
    return (R.fixedBits && (R.fixedBits[nd] >> d & 1)) || R.owner[d][nd] != RIGID_NONE;
}

// Mirrored rigid BCs (RigidMaterial::mirrored, NodalVelBC::SetMirroredVelBC :214-244, ReflectVelocityBC): when the next node
// towards the body is fixed too and the one after it is a free node of the body, the BC velocity becomes
// v0 + (v0 - v_mirror).  Returns the velocity to impose, or sets skip when the reference adds nothing.
__device__ __forceinline__ double rigid_bc_velocity(const RigidBCs &R, const Nodes &N, int nd, int d, int owner, bool &skip)
{
    skip = false;
    const double v0 = R.vel[d][owner];
    if (!R.mirrored) return v0;
    const int mirrored = (int)R.mats[R.mat[owner]].p[9];
    if (mirrored == 0) return v0;
    const int s = mirrored < 0 ? R.stride[d] : -R.stride[d];
    const int neighbor = nd + s, mirror = nd + 2 * s;
    if (neighbor < 0 || neighbor >= R.nnodes || !rigid_dof_fixed(R, neighbor, d)) return v0;
    if (mirror < 0 || mirror >= R.nnodes || rigid_dof_fixed(R, mirror, d) || N.cnt[mirror] <= 0) return v0;
    return v0 + (v0 - N.pk[d][mirror] / N.mass[mirror]);
}

// One thread per node that has BCs and material velocity field.  BCs act only on active fields (numberPoints>0,
// NodalPointMPM.cpp:1849-1862; every active nonrigid field in multimaterial mode, CrackVelocityFieldMulti.cpp:1243-1260).
__global__ void k_velocity_bcs(VelBCs B, Nodes N, int pass, double dt, int adjustSym, int nf = 1, int nnodes = 0)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B.nUnique * nf) return;
    const int u = t % B.nUnique, off = (t / B.nUnique) * nnodes;
    const int nd = B.node[u] + off;
    if (N.cnt[nd] <= 0) return;
    if (adjustSym) {        // ADJUST_COPIED_PK==1 (NodalVelBC.cpp:339-353, MatVelocityField.cpp:579-586)
        int sd = B.symdir[u];
        if (sd & 32) N.pkc[0][nd] = 0.;
        if (sd & 64) N.pkc[1][nd] = 0.;
        if (sd & 128) N.pkc[2][nd] = 0.;
        if (adjustSym == 2) return;      // symmetry adjust only (no USF task: NodalVelBC.cpp:356)
    }
    double pk[3] = {N.pk[0][nd], N.pk[1][nd], N.pk[2][nd]};
    double ft[3] = {N.ftot[0][nd], N.ftot[1][nd], N.ftot[2][nd]};
    node_bcs(B, u, pass, dt, N.mass[nd], pk, ft, &N, off);
    N.pk[0][nd] = pk[0]; N.pk[1][nd] = pk[1]; N.pk[2][nd] = pk[2];
    N.ftot[0][nd] = ft[0]; N.ftot[1][nd] = ft[1]; N.ftot[2][nd] = ft[2];
}

// rigid-particle BCs on every node, launched after k_velocity_bcs
// (nnodes = fields x real nodes in multimaterial mode; the claims are per real node)
__global__ void k_rigid_velocity_bcs(int nnodes, RigidBCs R, Nodes N, int pass, double dt)
{
    int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= nnodes) return;
    if (N.cnt[nd] <= 0) return;
    double pk[3] = {N.pk[0][nd], N.pk[1][nd], N.pk[2][nd]};
    double ft[3] = {N.ftot[0][nd], N.ftot[1][nd], N.ftot[2][nd]};
    const int real = nd % R.nnodes;
    if (!(R.mirrored ? node_rigid_bcs<true>(R, nd, pass, dt, N.mass[nd], pk, ft, &N) : node_rigid_bcs<false>(R, real, pass, dt, N.mass[nd], pk, ft))) return;
    N.pk[0][nd] = pk[0]; N.pk[1][nd] = pk[1]; N.pk[2][nd] = pk[2];
    N.ftot[0][nd] = ft[0]; N.ftot[1][nd] = ft[1]; N.ftot[2][nd] = ft[2];
}

// ---- ProjectRigidBCsTask (ProjectRigidBCsTask.cpp:39-158) ---------------------------------------
// PR holds the rigid-BC particles in host order; the reference walks them serially and the first one
// to reach a free dof of a node keeps it, which is the minimum index here.
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_project_rigid_bcs(Grid g, Particles PR, const Material *mats, RigidBCs R, StatusFlags *flags)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= PR.n) return;
    const int dirs = (int)mats[PR.mat[p]].p[8];      // RigidMaterial::setDirection: x=1, y=2, z=4
    const bool setsT = R.ownerT != nullptr && mats[PR.mat[p]].p[10] != 0.;      // RigidMaterial::RigidTemperature
    auto claim = [&](int nd, double, double, double, double) {
        const int fixed = R.fixedBits ? R.fixedBits[nd] : 0;
#pragma unroll
        for (int d = 0; d < DIM; d++)
            if ((dirs >> d & 1) && !(fixed >> d & 1)) atomicMin(&R.owner[d][nd], p);
        if (setsT && !(R.fixedT && R.fixedT[nd])) atomicMin(&R.ownerT[nd], p);
    };
    if (SHAPE_IS_CPDI(SHAPE)) {
        // the nodes of the particle domain's corners (the reference's InitializationTask finds them for rigid particles too)
        if (!cpdi_setup<DIM, SHAPE>(g, PR, p)) { atomicCAS(&flags->cpdiLeft, 0, PR.orig[p] + 1); return; }
        for_each_node_cpdi<DIM, SHAPE, false>(g, PR, p, claim);
    } else {
        const int e = PR.elem[p];
        double pos[3] = {PR.pos[0][p], PR.pos[1][p], DIM == 3 ? PR.pos[2][p] : 0.};
        double xi[3], lp[3] = {PR.lp[0][p], PR.lp[1][p], PR.lp[2][p]};
        get_xipos<DIM>(g, e, pos, xi);
        for_each_node<DIM, SHAPE_IS_CPDI(SHAPE) ? SHAPE_LINEAR : SHAPE, false>(g, e, xi, lp, claim);
    }
}

// rigid particles move at their own velocity (UpdateParticlesTask.cpp:292-295, MatPoint3D.cpp:197-201)
template <int DIM>
__global__ void k_move_rigid(Particles PR, double dt)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= PR.n) return;
    PR.pos[0][p] += PR.vel[0][p] * dt;
    PR.pos[1][p] += PR.vel[1][p] * dt;
    if (DIM == 3) PR.pos[2][p] += PR.vel[2][p] * dt;
}

// ---- grid velocity for strain / particle update (MatVelocityField.cpp:239-251) ---------------
__global__ void k_grid_velocity(int nnodes, Nodes N)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    const double m = N.mass[i];
    if (N.cnt[i] == 0 || m == 0.) return;
    const double rm = 1. / m;
    N.vk[0][i] = N.pk[0][i] * rm; N.vk[1][i] = N.pk[1][i] * rm; N.vk[2][i] = N.pk[2][i] * rm;
}

// ---- particle state <-> registers -------------------------------------------------------------
__device__ __forceinline__ void load_pstate(const Particles &P, int p, PState &s)
{
#pragma unroll
    for (int i = 0; i < 9; i++) s.F[i] = P.F[i][p];
#pragma unroll
    for (int i = 0; i < 6; i++) { s.sp[i] = P.sp[i][p]; s.eplast[i] = P.eplast[i][p]; }
    s.pressure = P.pressure[p];
    s.work = P.work[p]; s.res = P.res[p]; s.heat = P.heat[p]; s.entropy = P.entropy[p]; s.plast = P.plast[p];
    s.prevT = P.prevT[p];
    s.dT = 0.; s.dTad = 0.; s.adiabatic = 0;
#pragma unroll
    for (int i = 0; i < MPM_MAX_HISTORY; i++) s.hist[i] = P.hist[i][p];
}

__device__ __forceinline__ void store_pstate(const Particles &P, int p, const PState &s)
{
#pragma unroll
    for (int i = 0; i < 9; i++) P.F[i][p] = s.F[i];
#pragma unroll
    for (int i = 0; i < 6; i++) { P.sp[i][p] = s.sp[i]; P.eplast[i][p] = s.eplast[i]; }
    P.pressure[p] = s.pressure;
    P.work[p] = s.work; P.res[p] = s.res; P.heat[p] = s.heat; P.entropy[p] = s.entropy; P.plast[p] = s.plast;
#pragma unroll
    for (int i = 0; i < MPM_MAX_HISTORY; i++) P.hist[i][p] = s.hist[i];
}

// ---- tasks 4 and 9b: FullStrainUpdate (UpdateStrainsFirstTask.cpp:101-168, MatPoint3D.cpp:45-93) ----
// dTscale: share of the step's temperature change this pass answers to (MPMBase::ScaledResidualStrains, MPMBase.cpp:346-360:
// fractionUSF / 1 - fractionUSF for the two passes of USAVG, 1 otherwise)
template <int DIM, int SHAPE, bool LRLAW>
__device__ __forceinline__ void update_strains_body(const Grid &g, const Particles &P, const Nodes &N, const Material *mats,
                                                    double strainTime, double dTscale)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    double dv[9] = {0., 0., 0., 0., 0., 0., 0., 0., 0.};
    particle_nodes<DIM, SHAPE, true>(g, P, p, [&](int nd, double S, double gx, double gy, double gz) {
        const double vx = N.vk[0][nd], vy = N.vk[1][nd];
        dv[0] += vx * gx; dv[1] += vx * gy;
        dv[3] += vy * gx; dv[4] += vy * gy;
        if (DIM == 3) {
            const double vz = N.vk[2][nd];
            dv[2] += vx * gz; dv[5] += vy * gz;
            dv[6] += vz * gx; dv[7] += vz * gy; dv[8] += vz * gz;
        }
    });
#pragma unroll
    for (int i = 0; i < 9; i++) dv[i] *= strainTime;
    PState s;
    load_pstate(P, p, s);
    if (P.dTr) s.dT = P.dTr[p] * dTscale;
    if (P.dTad) s.adiabatic = 1;
    if (LRLAW) constitutive_law_lr<DIM>(s, dv, strainTime, g.np, mats[P.mat[p]]);
    else constitutive_law<DIM>(s, dv, strainTime, g.np, mats[P.mat[p]]);
    store_pstate(P, p, s);
    if (P.dTad) P.dTad[p] += s.dTad;          // MPMBase::Add_dTad
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_update_strains(Grid g, Particles P, Nodes N, const Material *mats,
                                                                 double strainTime, double dTscale)
{
    update_strains_body<DIM, SHAPE, false>(g, P, N, mats, strainTime, dTscale);
}

// the same task when some material asks for the large-rotation hypoelastic update (Elastic::useLargeRotation)
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_update_strains_lr(Grid g, Particles P, Nodes N, const Material *mats,
                                                                    double strainTime, double dTscale)
{
    update_strains_body<DIM, SHAPE, true>(g, P, N, mats, strainTime, dTscale);
}

// ---- task 5: GridForcesTask (GridForcesTask.cpp:55-117, MatPoint3D.cpp:248-252) ---------------
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_forces(Grid g, Particles P, Nodes N, int hasFext)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const double mp = P.mp[p], pr = P.pressure[p];
    const double sxx = P.sp[XX][p] - pr, syy = P.sp[YY][p] - pr, sxy = P.sp[XY][p];
    double szz = 0., syz = 0., sxz = 0.;
    if (DIM == 3) { szz = P.sp[ZZ][p] - pr; syz = P.sp[YZ][p]; sxz = P.sp[XZ][p]; }
    double fx = 0., fy = 0., fz = 0.;
    if (hasFext) { fx = P.pfext[0][p]; fy = P.pfext[1][p]; fz = P.pfext[2][p]; }
    particle_nodes<DIM, SHAPE, true>(g, P, p, [&](int nd, double S, double gx, double gy, double gz) {
        if (DIM == 3) {
            atomAdd(&N.ftot[0][nd], -mp * (sxx * gx + sxy * gy + sxz * gz) + S * fx);
            atomAdd(&N.ftot[1][nd], -mp * (sxy * gx + syy * gy + syz * gz) + S * fy);
            atomAdd(&N.ftot[2][nd], -mp * (sxz * gx + syz * gy + szz * gz) + S * fz);
        } else {
            atomAdd(&N.ftot[0][nd], -mp * (sxx * gx + sxy * gy) + S * fx);
            atomAdd(&N.ftot[1][nd], -mp * (sxy * gx + syy * gy) + S * fy);
        }
    });
}

// Contact force quantities (GlobalQuantity.cpp:905-968 -> NodalPoint::AddGetContactForce -> CrackVelocityFieldMulti::
// SumAndClearRigidContactForces :1712-1728): the summed force row of every rigid material field over the nodes where that field is
// active now, cleared after reading
__global__ void k_contact_force_sum(int nnodes, int nf, unsigned rigidMask, ContactNodes C, int clear, double *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnodes * nf) return;
    const int f = t / nnodes;
    if (!(rigidMask >> f & 1) || C.rcnt[t] <= 0) return;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double v = C.rforce[c][t];
        if (v != 0.) atomicAdd(out + 3 * f + c, v);
        if (clear) C.rforce[c][t] = 0.;
    }
}

// ---- task 6a: particle traction BCs (PostForcesTask.cpp:51 -> MatPtTractionBC::AddMPFluxBC, MatPtTractionBC.cpp:64-226) ----
// One thread per particle; its traction entries load one face of its domain each.  The face's corners (2 in 2D, 4 in 3D) come from
// MatPoint2D/3D::GetSurfaceInfo (MatPoint2D.cpp, MatPoint3D.cpp): semi-side vectors F.lp for the CPDI shapes, the undeformed ones for
// the others with the weight taken from the deformed face all the same; every corner hands wtNorm * value * N_i(corner) to the
// nodes of its element (plain element shape functions, ElementBase::GetShapeFunctionsForTractions) that carry non-rigid particles
// and the particle's material field.  wtNorm = direction x (face area / corners).
// fluxQ != NULL: the same walk for particle heat-flux BCs (MatPtHeatFluxBC::AddMPFluxBC, MatPtHeatFluxBC.cpp:64-160, external flux): the
// value is a scalar flux, the direction argument of GetSurfaceInfo is x only to carry the face weight, and value x weight x N_i goes
// into the transport rate gQ of every node with non-rigid particles (TransportTask::AddFluxCondition, TransportTask.cpp:302-308).
// spline: the quadratic B-spline shape functions of the corner's element stand in for the plain element ones (B2SPLINE, B2GIMP, B2CPDI:
// ElementBase::GetShapeFunctionsForTractions, MoreMPMElementBase.cpp:91-107).
template <int DIM>
__global__ void __launch_bounds__(TASK_THREADS) k_particle_tractions(Grid g, Particles P, Nodes N, TractionBCs TB, int cpdi, double thickness, int nf, StatusFlags *flags,
                                                                     double *fluxQ = nullptr, int spline = 0)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const int o = P.orig[p];
    const int e0 = TB.start[o], e1 = TB.start[o + 1];
    if (e0 == e1) return;
    const int el = P.elem[p];
    const ElemIJK c0 = elem_ijk(g, el);
    const double psx = (g.xpts[c0.i + 1] - g.xpts[c0.i]) * (0.5 * P.lp[0][p]), psy = (g.ypts[c0.j + 1] - g.ypts[c0.j]) * (0.5 * P.lp[1][p]);
    const double psz = DIM == 3 ? (g.zpts[c0.k + 1] - g.zpts[c0.k]) * (0.5 * P.lp[2][p]) : 0.;
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = P.F[i][p];
    // deformed semi-side vectors (GetSemiSideVectors)
    const double d1[3] = {F[0] * psx, F[3] * psx, DIM == 3 ? F[6] * psx : 0.};
    const double d2[3] = {F[1] * psy, F[4] * psy, DIM == 3 ? F[7] * psy : 0.};
    const double d3[3] = {DIM == 3 ? F[2] * psz : 0., DIM == 3 ? F[5] * psz : 0., DIM == 3 ? F[8] * psz : 0.};
    const double pos[3] = {P.pos[0][p], P.pos[1][p], DIM == 3 ? P.pos[2][p] : 0.};
    const int off = nf > 1 ? P.foff[p] : 0;
    for (int e = e0; e < e1; e++) {
        const int face = TB.face[e], dof = fluxQ ? 1 : TB.dir[e];
        const double tmag = TB.value[e];
        double r1[3] = {d1[0], d1[1], d1[2]}, r2[3] = {d2[0], d2[1], d2[2]}, r3[3] = {d3[0], d3[1], d3[2]};
        double uSize = -1.;
        if (!cpdi) {        // undeformed edge for the extrapolation, deformed size for the weight
            if (DIM == 3) {
                const double *a = (face == 2 || face == 4) ? r2 : r1, *b = (face == 1 || face == 3 || face == 2 || face == 4) ? r3 : r2;
                const double Ap[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                uSize = sqrt(Ap[0] * Ap[0] + Ap[1] * Ap[1] + Ap[2] * Ap[2]);
            } else {
                const double *a = (face == 1 || face == 3) ? r1 : r2;
                uSize = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            }
            r1[0] = psx; r1[1] = 0.; r1[2] = 0.; r2[0] = 0.; r2[1] = psy; r2[2] = 0.; r3[0] = 0.; r3[1] = 0.; r3[2] = psz;
        }
        // corners c[0..NCF-1] of the face and one more point off the face (signs of r1, r2, r3)
        const int NCF = DIM == 3 ? 4 : 2;
        double sgn[5][3];
        if (DIM == 3) {
            const double S3[6][5][3] = {
                {{-1, -1, -1}, {1, -1, -1}, {1, -1, 1}, {-1, -1, 1}, {-1, 1, -1}},
                {{1, -1, -1}, {1, 1, -1}, {1, 1, 1}, {1, -1, 1}, {-1, -1, -1}},
                {{1, 1, -1}, {-1, 1, -1}, {-1, 1, 1}, {1, 1, 1}, {1, -1, -1}},
                {{-1, 1, -1}, {-1, -1, -1}, {-1, -1, 1}, {-1, 1, 1}, {1, 1, -1}},
                {{1, -1, -1}, {-1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {1, -1, 1}},
                {{-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}, {-1, -1, -1}}};
            const int f = face >= 1 && face <= 5 ? face - 1 : 5;
            for (int i = 0; i < 5; i++) for (int d = 0; d < 3; d++) sgn[i][d] = S3[f][i][d];
        } else {
            const double S2[4][3][2] = {{{-1, -1}, {1, -1}, {-1, 1}}, {{1, -1}, {1, 1}, {-1, -1}}, {{1, 1}, {-1, 1}, {1, -1}}, {{-1, 1}, {-1, -1}, {1, 1}}};
            const int f = face >= 1 && face <= 3 ? face - 1 : 3;
            for (int i = 0; i < 3; i++) { sgn[i][0] = S2[f][i][0]; sgn[i][1] = S2[f][i][1]; sgn[i][2] = 0.; }
            for (int d = 0; d < 3; d++) { sgn[3][d] = 0.; sgn[4][d] = 0.; }
        }
        double c[4][3];
        for (int i = 0; i < NCF; i++)
            for (int d = 0; d < 3; d++) c[i][d] = pos[d] + sgn[i][0] * r1[d] + sgn[i][1] * r2[d] + sgn[i][2] * r3[d];
        // edge radii and the weighted direction
        double ra[3], rb[3];
        for (int d = 0; d < 3; d++) { ra[d] = 0.5 * (c[1][d] - c[0][d]); rb[d] = DIM == 3 ? 0.5 * (c[3][d] - c[0][d]) : 0.; }
        double wt[3] = {0., 0., 0.};
        if (DIM == 3) {
            double Ap[3] = {ra[1] * rb[2] - ra[2] * rb[1], ra[2] * rb[0] - ra[0] * rb[2], ra[0] * rb[1] - ra[1] * rb[0]};
            const double faceWt = uSize < 0. ? sqrt(Ap[0] * Ap[0] + Ap[1] * Ap[1] + Ap[2] * Ap[2]) : uSize;
            if (dof == 1) wt[0] = faceWt;
            else if (dof == 2) wt[1] = faceWt;
            else if (dof == 11) {
                if (uSize > 0.) { const double sc = uSize / sqrt(Ap[0] * Ap[0] + Ap[1] * Ap[1] + Ap[2] * Ap[2]); Ap[0] *= sc; Ap[1] *= sc; Ap[2] *= sc; }
                wt[0] = Ap[0]; wt[1] = Ap[1]; wt[2] = Ap[2];
            } else wt[2] = faceWt;
        } else {
            double faceWt = uSize < 0. ? sqrt(ra[0] * ra[0] + ra[1] * ra[1]) : uSize;
            faceWt *= thickness;
            if (dof == 1) wt[0] = faceWt;
            else if (dof == 2) wt[1] = faceWt;
            else if (dof == 11) { const double ex = ra[1], ey = -ra[0], en = sqrt(ex * ex + ey * ey); wt[0] = ex * faceWt / en; wt[1] = ey * faceWt / en; }
            else if (dof == 12) { const double ex = ra[0], ey = ra[1], en = sqrt(ex * ex + ey * ey); wt[0] = ex * faceWt / en; wt[1] = ey * faceWt / en; }
            else wt[2] = faceWt;
        }
        const double lpz[3] = {0., 0., 0.};
        for (int i = 0; i < NCF; i++) {
            const int ce = find_element_from_point<DIM>(g, c[i]);
            if (ce <= 0) { atomicCAS(&flags->cpdiLeft, 0, o + 1); continue; }     // "A Traction edge node has left the grid"
            double xi[3];
            get_xipos<DIM>(g, ce, c[i], xi);
            auto hand = [&](int nd, double S, double, double, double) {
                bool any = false;       // NodalPoint::NodeHasNonrigidParticles, then the particle's own field (AddTractionTask3)
                for (int f = 0; f < nf; f++) any |= N.cnt[nd + f * g.nnodes] > 0;
                if (fluxQ) { if (any) atomAdd(&fluxQ[nd], (tmag * wt[0]) * S); return; }
                if (!any || N.cnt[nd + off] <= 0) return;
                const double s = tmag * S;
                if (wt[0] != 0.) atomAdd(&N.ftot[0][nd + off], wt[0] * s);
                if (wt[1] != 0.) atomAdd(&N.ftot[1][nd + off], wt[1] * s);
                if (DIM == 3 && wt[2] != 0.) atomAdd(&N.ftot[2][nd + off], wt[2] * s);
            };
            if (spline) for_each_node<DIM, SHAPE_B2SPLINE, false>(g, ce, xi, lpz, hand);
            else for_each_node<DIM, SHAPE_LINEAR, false>(g, ce, xi, lpz, hand);
        }
    }
}

// ---- task 6: PostForcesTask node pass (PostForcesTask.cpp:46-94) -------------------------------
__global__ void k_post_forces(int nnodes, Nodes N, StepParams sp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    if (N.cnt[i] == 0) return;
    N.pk[0][i] = N.pkc[0][i]; N.pk[1][i] = N.pkc[1][i]; N.pk[2][i] = N.pkc[2][i];     // RestoreMomenta
    if (sp.hasGravity) {                                                                  // AddGravityAndBodyForceTask3
        const double m = N.mass[i];
        N.ftot[0][i] += m * sp.grav[0]; N.ftot[1][i] += m * sp.grav[1]; N.ftot[2][i] += m * sp.grav[2];
    }
}

// ---- task 7: UpdateMomentaTask node pass (MatVelocityField.cpp:301-304) ------------------------
__global__ void k_update_momenta(int nnodes, Nodes N, double dt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    if (N.cnt[i] == 0) return;
    N.pk[0][i] += N.ftot[0][i] * dt; N.pk[1][i] += N.ftot[1][i] * dt; N.pk[2][i] += N.ftot[2][i] * dt;
}

// ---- task 8: UpdateParticlesTask (UpdateParticlesTask.cpp:99-286, MatPoint3D.cpp:104-194) ------
// m: 0 FLIP, >0 FMPM(m), <0 XPIC(-m)
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_update_particles(Grid g, Particles P, Nodes N, const Material *mats,
                                                                   StepParams sp, int m)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    double Svk[3] = {0., 0., 0.}, Sacc[3] = {0., 0., 0.};
    particle_nodes<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        Svk[0] += N.vk[0][nd] * S; Svk[1] += N.vk[1][nd] * S; Svk[2] += N.vk[2][nd] * S;
        if (m <= 0) {
            const double mnode = S / N.mass[nd];
            Sacc[0] += N.ftot[0][nd] * mnode; Sacc[1] += N.ftot[1][nd] * mnode; Sacc[2] += N.ftot[2][nd] * mnode;
        }
    });
    const double dt = sp.dt;
    const double matDamp = mats[P.mat[p]].p[2];
    const double pAlpha = matDamp >= 0. ? matDamp : sp.particleAlpha;      // MaterialBase::GetMaterialDamping
    double vel[3] = {P.vel[0][p], P.vel[1][p], P.vel[2][p]};
    double pos[3] = {P.pos[0][p], P.pos[1][p], P.pos[2][p]};
    double vm[3], delV[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        vm[c] = Svk[c];
        if (m == 0) vm[c] += Sacc[c] * (-dt);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (DIM == 2 && c == 2) { delV[c] = 0.; continue; }
        double Adamp0 = vel[c] * pAlpha;
        Adamp0 += vm[c] * sp.gridAlpha;
        if (m > 0) {
            double delXRate = vel[c];
            vel[c] = vm[c];
            vel[c] += Adamp0 * (-dt);
            delV[c] = vel[c] - delXRate;
            delXRate += vel[c];
            pos[c] += delXRate * (0.5 * dt);
        } else if (m == 0) {
            delV[c] = (Sacc[c] - Adamp0) * dt;
            vel[c] += delV[c];
            double delXRate = vm[c] + 0.5 * delV[c];
            pos[c] += delXRate * dt;
        } else {
            double delXRate = vel[c];
            vel[c] = Svk[c] - Adamp0 * dt;
            delV[c] = vel[c] - delXRate;
            delXRate = vm[c] + 0.5 * delV[c];
            pos[c] += delXRate * dt;
        }
    }
#pragma unroll
    for (int c = 0; c < DIM; c++) {
        P.vel[c][p] = vel[c];
        P.pos[c][p] = pos[c];
        P.acc[c][p] = delV[c] / dt;
    }
}

// ---- XPIC(k) / FMPM(k), k > 1: XPICExtrapolationTask.cpp:49-161, MatVelocityField::XPICSupport :318-397 ----
// vk = v* (VSTAR_VEC), vsp = v*prev, vsn = v*next.  The reference's per-particle double loop
//   v*next_i += (mp S_ip S_jp / m_i) v*prev_j          (XPICDoubleLoop :183-214, S^2 pair operations)
// is evaluated as a gather followed by a scatter, u_p = sum_j S_jp v*prev_j ; v*next_i += (mp S_ip / m_i) u_p.
__global__ void k_xpic_init(int nnodes, Nodes N, double dt, int usingFMPM)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    N.vsn[0][i] = 0.; N.vsn[1][i] = 0.; N.vsn[2][i] = 0.;
    if (N.cnt[i] == 0) return;
    const double mass = N.mass[i], rm = 1. / mass;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (usingFMPM) {
            const double v = N.pk[c][i] * rm;
            N.vsp[c][i] = v; N.vk[c][i] = v;
        } else {
            double v = N.pk[c][i];
            v += N.ftot[c][i] * (-dt);
            v *= rm;
            N.vsp[c][i] = v;
            N.vk[c][i] = v + N.ftot[c][i] * (dt / mass);
        }
    }
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_xpic_iterate(Grid g, Particles P, Nodes N)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    double u[3] = {0., 0., 0.};
    particle_nodes<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        u[0] += S * N.vsp[0][nd]; u[1] += S * N.vsp[1][nd]; u[2] += S * N.vsp[2][nd];
    });
    const double mp = P.mp[p];
    particle_nodes<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        const double w = mp * S / N.mass[nd];
        atomAdd(&N.vsn[0][nd], w * u[0]);
        atomAdd(&N.vsn[1][nd], w * u[1]);
        if (DIM == 3) atomAdd(&N.vsn[2][nd], w * u[2]);
    });
}

// GET_DELTAV, velocity BCs on the increment (XPIC_* pass of ZeroVelocityBC, MatVelocityField.cpp:529-541), UPDATE_VSTAR
__global__ void k_xpic_finish(int nnodes, Nodes N, VelBCs B, const int *bcOfNode, RigidBCs R, double dt, int particleUpdate, int usingFMPM)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    if (N.cnt[i] == 0) return;
    double d[3] = {N.vsp[0][i] - N.vsn[0][i], N.vsp[1][i] - N.vsn[1][i], N.vsp[2][i] - N.vsn[2][i]};
    const int u = bcOfNode ? bcOfNode[i] : -1;
    if (u >= 0) {
        const double mass = N.mass[i];
        double ft[3] = {N.ftot[0][i], N.ftot[1][i], N.ftot[2][i]};
        for (int e = B.start[u]; e < B.start[u + 1]; e++) {
            if (!B.active[e]) continue;
            const double nx = B.norm[3 * e], ny = B.norm[3 * e + 1], nz = B.norm[3 * e + 2];
            const double dotn = d[0] * nx + d[1] * ny + d[2] * nz;
            d[0] += nx * (-dotn); d[1] += ny * (-dotn); d[2] += nz * (-dotn);
            if (particleUpdate && B.reaction) {     // lumped-mass addition to freaction (MatVelocityField.cpp:532-536)
                const double s = -mass * dotn / dt;
                const double r[3] = {nx * s, ny * s, nz * s};
                add_reaction(B.reaction + 3 * e, r);
            }
            if (particleUpdate && !usingFMPM) {
                const double s = -mass * dotn / dt;
                ft[0] += nx * s; ft[1] += ny * s; ft[2] += nz * s;
            }
        }
        if (particleUpdate && !usingFMPM) { N.ftot[0][i] = ft[0]; N.ftot[1][i] = ft[1]; N.ftot[2][i] = ft[2]; }
    }
    if (R.on) {         // the rigid-particle BCs follow in the list, one axis each
        const double mass = N.mass[i];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (R.owner[c][i] == RIGID_NONE) continue;
            const double dotn = d[c];
            d[c] += -dotn;
            if (particleUpdate && R.reaction && dotn != 0.) atomicAdd(R.reaction + 3 * R.mat[R.owner[c][i]] + c, -mass * dotn / dt);
            if (particleUpdate && !usingFMPM) N.ftot[c][i] += -mass * dotn / dt;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        N.vsp[c][i] = d[c];
        N.vk[c][i] += d[c];
        N.vsn[c][i] = 0.;
    }
}

// ---- task 9a: UpdateStrainsLastContactTask re-extrapolation (UpdateStrainsLastContactTask.cpp:71-149) ----
__global__ void k_rezero_momenta(int nnodes, Nodes N)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    N.pk[0][i] = 0.; N.pk[1][i] = 0.; N.pk[2][i] = 0.;
}

template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_momentum_last(Grid g, Particles P, Nodes N)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const double mp = P.mp[p];
    const double vx = P.vel[0][p], vy = P.vel[1][p], vz = DIM == 3 ? P.vel[2][p] : 0.;
    particle_nodes<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        const double fnmp = S * mp;
        atomAdd(&N.pk[0][nd], vx * fnmp);
        atomAdd(&N.pk[1][nd], vy * fnmp);
        if (DIM == 3) atomAdd(&N.pk[2][nd], vz * fnmp);
    });
}

// ---- task 11: ResetElementsTask (ResetElementsTask.cpp:196-265) ---------------------------------
template <int DIM>
__device__ __forceinline__ void reset_element_one(const Grid &g, const Particles &P, int p, StatusFlags *flags, double dt)
{
    double pos[3] = {P.pos[0][p], P.pos[1][p], DIM == 3 ? P.pos[2][p] : 0.};
    if (pos[0] != pos[0] || pos[1] != pos[1] || pos[2] != pos[2]) {
        atomicCAS(&flags->nanParticle, 0, P.orig[p] + 1);
        return;
    }
    const int e = P.elem[p];
    if (pt_in_element<DIM>(g, e, pos)) return;                 // SAME_ELEMENT
    int ne = find_element_from_point<DIM>(g, pos);
    if (ne > 0 && !edge_element<DIM>(g, ne)) {                 // NEW_ELEMENT (MPMBase::ChangeElemID)
        P.elem[p] = ne;
        int c = P.cross[p];
        P.cross[p] = c >= 0 ? c + 1 : c - 1;
        atomicAdd(&flags->crossings, 1ull);
        return;
    }
    // LEFT_GRID: count it (IncrementElementCrossings flips the sign on the first exit) and push the
    // particle back into its element by bisection (ReturnToElement, ResetElementsTask.cpp:232-265)
    {
        int c = P.cross[p];
        P.cross[p] = c > 0 ? -(c + 1) : c - 1;
        // low word: exits; high word: particles leaving for the first time (MPMBase::HasLeftTheGridBefore = negative crossings) --
        // the ones the reference issues its "left the grid" warning for (ResetElementsTask.cpp:71-95)
        atomicAdd(&flags->leftGrid, c >= 0 ? ((1ull << 32) | 1ull) : 1ull);
    }
    double outside[3] = {pos[0], pos[1], pos[2]};
    double inside[3];
    inside[0] = outside[0] - dt * P.vel[0][p];
    inside[1] = outside[1] - dt * P.vel[1][p];
    inside[2] = DIM == 3 ? outside[2] - dt * P.vel[2][p] : 0.;
    if (!pt_in_element<DIM>(g, e, inside)) {
        ElemIJK c = elem_ijk(g, e);
        inside[0] = (g.xpts[c.i] + g.xpts[c.i + 1]) / 2.;
        inside[1] = (g.ypts[c.j] + g.ypts[c.j + 1]) / 2.;
        inside[2] = DIM == 3 ? (g.zpts[c.k] + g.zpts[c.k + 1]) / 2. : 0.;
    }
    for (int pass = 1; pass <= 10; pass++) {
        double middle[3] = {(outside[0] + inside[0]) / 2., (outside[1] + inside[1]) / 2., (outside[2] + inside[2]) / 2.};
        if (pt_in_element<DIM>(g, e, middle)) { inside[0] = middle[0]; inside[1] = middle[1]; inside[2] = middle[2]; }
        else { outside[0] = middle[0]; outside[1] = middle[1]; outside[2] = middle[2]; }
    }
    P.pos[0][p] = inside[0]; P.pos[1][p] = inside[1];
    if (DIM == 3) P.pos[2][p] = inside[2];
}

template <int DIM>
__global__ void __launch_bounds__(TASK_THREADS) k_reset_elements(Grid g, Particles P, StatusFlags *flags, double dt)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    reset_element_one<DIM>(g, P, p, flags, dt);
}

// =================================================================================================================
// Multimaterial mode: one velocity field per material field on every node + material contact
// (CrackVelocityFieldMulti with a single crack field)
// =================================================================================================================

// node offset of every particle's material velocity field (MaterialBase::GetField), fixed for the run
__global__ void k_set_field_offsets(int n, const int *mat, const int *fieldOfMat, int nnodes, int *foff)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) foff[p] = fieldOfMat[mat[p]] * nnodes;
}

// The extrapolations contact needs besides mass and momentum (NodalPoint::AddMassMomentum / AddMassMomentumLast,
// NodalPointMPM.cpp:419-453, :479-496): contact volume, mass-weighted displacement or position
// (CrackVelocityField::AddVolumeDisplacement, CrackVelocityField.cpp:395-401) and the volume gradient
// (CrackVelocityFieldMulti::AddVolumeGradient, CrackVelocityFieldMulti.cpp:108-112).
// origpos: [3][norig] in the caller's particle order (MPMBase::origpos).
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_contact_terms(Grid g, Particles P, const Material *mats, ContactNodes C,
                                                                    const double *origpos, size_t norig, int byDisplacements, int needGradient)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const double mp = P.mp[p], rho = mats[P.mat[p]].p[0];
    // MPMBase::GetVolume(DEFORMED_AREA): 3D det(F) mp / rho (MatPoint3D.cpp:395-408); 2D the in-plane part of F (MatPoint2D.cpp:407-417)
    double vol;
    if (DIM == 3) {
        const double F00 = P.F[0][p], F01 = P.F[1][p], F02 = P.F[2][p], F10 = P.F[3][p], F11 = P.F[4][p], F12 = P.F[5][p],
                     F20 = P.F[6][p], F21 = P.F[7][p], F22 = P.F[8][p];
        const double J = F00 * (F11 * F22 - F21 * F12) - F01 * (F10 * F22 - F20 * F12) + F02 * (F10 * F21 - F20 * F11);
        vol = J * mp / rho;
    } else {
        vol = (P.F[0][p] * P.F[4][p] - P.F[3][p] * P.F[1][p]) * mp / rho;
    }
    double d[3] = {P.pos[0][p], P.pos[1][p], DIM == 3 ? P.pos[2][p] : 0.};
    if (byDisplacements) {
        const size_t q = (size_t)P.orig[p];
        d[0] -= origpos[q]; d[1] -= origpos[norig + q];
        if (DIM == 3) d[2] -= origpos[2 * norig + q];
    }
    particle_nodes<DIM, SHAPE, true>(g, P, p, [&](int nd, double S, double gx, double gy, double gz) {
        const double fnmp = S * mp;
        atomAdd(&C.cvol[nd], S * vol);
        atomAdd(&C.cdisp[0][nd], d[0] * fnmp);
        atomAdd(&C.cdisp[1][nd], d[1] * fnmp);
        if (DIM == 3) atomAdd(&C.cdisp[2][nd], d[2] * fnmp);
        if (needGradient) {
            atomAdd(&C.cgrad[0][nd], gx * vol);
            atomAdd(&C.cgrad[1][nd], gy * vol);
            if (DIM == 3) atomAdd(&C.cgrad[2][nd], gz * vol);
        }
    });
}

// CrackVelocityFieldMulti::RezeroNodeTask6 (:157-176): nonrigid fields lose momentum and contact terms; a rigid field keeps its
// momentum and moves its displacement / position extrapolation on by pk dt (its particle mass is its volume)
__global__ void k_rezero_fields_task6(int nnodes, int nf, int rigidMask, Nodes N, ContactNodes C, double dt)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nnodes * nf) return;
    if (rigidMask >> (v / nnodes) & 1) {
        if (C.rcnt[v] > 0) { C.cdisp[0][v] += N.pk[0][v] * dt; C.cdisp[1][v] += N.pk[1][v] * dt; C.cdisp[2][v] += N.pk[2][v] * dt; }
        return;
    }
    N.pk[0][v] = 0.; N.pk[1][v] = 0.; N.pk[2][v] = 0.;
    C.cvol[v] = 0.;
#pragma unroll
    for (int c = 0; c < 3; c++) { C.cgrad[c][v] = 0.; C.cdisp[c][v] = 0.; }
}

// Rigid contact particles (RigidMaterial with SetDirection 8; the host's block FIRST_RIGID_CONTACT) extrapolate to the velocity
// field of their material: momentum, "mass" (= their volume: rho is 1), point count, contact volume from the unscaled volume,
// displacement / position and volume gradient (NodalPoint::AddMassMomentum with nonRigid == false, NodalPointMPM.cpp:419-453).
// They live in the rigid particle set PR beside the rigid-BC particles (which claim no field).
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_rigid_contact(Grid g, Particles PR, const Material *mats, Nodes N, ContactNodes C,
                                                                    const double *origpos, size_t norig, int byDisplacements, int needGradient, StatusFlags *flags)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= PR.n) return;
    const Material &m = mats[PR.mat[p]];
    if (m.kind != MAT_RIGIDCONTACT) return;
    double pos[3] = {PR.pos[0][p], PR.pos[1][p], DIM == 3 ? PR.pos[2][p] : 0.};
    double xi[3];
    get_xipos<DIM>(g, PR.elem[p], pos, xi);
    PR.ncpos[0][p] = xi[0]; PR.ncpos[1][p] = xi[1]; PR.ncpos[2][p] = xi[2];
    if (SHAPE_IS_CPDI(SHAPE) && !cpdi_setup<DIM, SHAPE>(g, PR, p)) { atomicCAS(&flags->cpdiLeft, 0, PR.orig[p] + 1); return; }
    const double mp = PR.mp[p];
    const double vol = mp / m.p[0];             // GetUnscaledVolume; F = I, so this is the volume for the gradient too
    const double vx = PR.vel[0][p], vy = PR.vel[1][p], vz = DIM == 3 ? PR.vel[2][p] : 0.;
    double d[3] = {pos[0], pos[1], pos[2]};
    if (byDisplacements) {
        const size_t q = (size_t)PR.orig[p];
        d[0] -= origpos[q]; d[1] -= origpos[norig + q];
        if (DIM == 3) d[2] -= origpos[2 * norig + q];
    }
    particle_nodes<DIM, SHAPE, true>(g, PR, p, [&](int nd, double S, double gx, double gy, double gz) {
        const double fnmp = S * mp;
        atomAdd(&N.pk[0][nd], vx * fnmp);
        atomAdd(&N.pk[1][nd], vy * fnmp);
        if (DIM == 3) atomAdd(&N.pk[2][nd], vz * fnmp);
        atomAdd(&N.mass[nd], fnmp);
        atomicAdd(&C.rcnt[nd], 1);
        atomAdd(&C.cvol[nd], S * vol);
        atomAdd(&C.cdisp[0][nd], d[0] * fnmp);
        atomAdd(&C.cdisp[1][nd], d[1] * fnmp);
        if (DIM == 3) atomAdd(&C.cdisp[2][nd], d[2] * fnmp);
        if (needGradient) {
            atomAdd(&C.cgrad[0][nd], gx * vol);
            atomAdd(&C.cgrad[1][nd], gy * vol);
            if (DIM == 3) atomAdd(&C.cgrad[2][nd], gz * vol);
        }
    });
}

// symmetry-plane bits of NodalPoint::fixedDirection (32 x, 64 y, 128 z) zero a vector's components (CrackVelocityField::AdjustForSymmetry)
__device__ __forceinline__ bool adjust_for_symmetry(int sd, double v[3])
{
    bool any = false;
    if (sd & 32) { v[0] = 0.; any = true; }
    if (sd & 64) { v[1] = 0.; any = true; }
    if (sd & 128) { v[2] = 0.; any = true; }
    return any;
}

// MeshInfo::GetPerpendicularDistance (MeshInfo.cpp:1741-1869), structured grid of equal elements: hperp along norm
__device__ __forceinline__ double perpendicular_distance(const Grid &g, int cubic, const double n[3])
{
    if (cubic) return g.gx;
    if (g.dim == 3) {
        double t1[3];
        if (n[2] > n[0] && n[2] > n[1]) { t1[0] = 0.; t1[1] = -n[2]; t1[2] = n[1]; }
        else { t1[0] = -n[1]; t1[1] = n[0]; t1[2] = 0.; }
        const double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
        const double a1 = t1[0] / g.gx, b1 = t1[1] / g.gy, c1 = t1[2] / g.gz;
        const double a2 = t2[0] / g.gx, b2 = t2[1] / g.gy, c2 = t2[2] / g.gz;
        return g.gx * g.gy * g.gz * sqrt((a1 * a1 + b1 * b1 + c1 * c1) * (a2 * a2 + b2 * b2 + c2 * c2));
    }
    const double a = g.gx * n[0], b = g.gy * n[1];
    return sqrt(a * a + b * b);
}

// CrackSurfaceContact::MaterialSeparation (CrackSurfaceContact.cpp:278-303)
__device__ __forceinline__ double material_separation(const Grid &g, const ContactParams &cp, double dbdotn, double dadotn, const double n[3],
                                                      const double xn[3])
{
    if (cp.byDisplacements) return dbdotn - dadotn;
    double r = cp.positionCutoff;
    const double hperp = perpendicular_distance(g, cp.cubic, n);
    if (r > 0.) return dbdotn - dadotn - r * hperp;
    r = -r;
    const double xdotn = xn[0] * n[0] + xn[1] * n[1] + xn[2] * n[2];
    const double pa = dadotn - xn[0] * n[0] - xn[1] * n[1] - xn[2] * n[2];
    const double da = pa > 0. ? 2. * pow(pa / (1.25 * hperp), r) - 1. : 1 - 2. * pow(-pa / (1.25 * hperp), r);
    const double pb = dbdotn - xn[0] * n[0] - xn[1] * n[1] - xn[2] * n[2];
    const double db = pb > 0. ? 2. * pow(pb / (1.25 * hperp), r) - 1. : 1 - 2. * pow(-pb / (1.25 * hperp), r);
    (void)xdotn;
    return (db - da) * hperp;
}

// CoulombFriction::GetFrictionalDeltaMomentum (CoulombFriction.cpp:150-272) without frictional heating (no conduction on this path).
// delFi != NULL in the momentum update.  Returns false when the law decides the materials are not in contact.
__device__ __forceinline__ bool frictional_delta_momentum(int kind, double mu, double muStatic, double delPi[3], const double n[3], double dotn,
                                                          double deltaDotn, double mred, double dt, const double *delFi)
{
    if (kind == LAW_STICK) return true;
    if (delFi) {
        const double fn = delFi[0] * n[0] + delFi[1] * n[1] + delFi[2] * n[2];
        deltaDotn += dt * (dotn - 0.5 * fn * dt) / mred;
    }
    const bool inContact = deltaDotn < 0. && dotn < 0.;
    if (kind == LAW_FRICTIONLESS) {
        if (!inContact) return false;
        delPi[0] = n[0] * dotn; delPi[1] = n[1] * dotn; delPi[2] = n[2] * dotn;
        return true;
    }
    if (!inContact) return false;                 // HasFreeSeparation
    double dott = 0.;
    double tang[3] = {delPi[0], delPi[1], delPi[2]};
    tang[0] += n[0] * (-dotn); tang[1] += n[1] * (-dotn); tang[2] += n[2] * (-dotn);
    const double tangMag = sqrt(tang[0] * tang[0] + tang[1] * tang[1] + tang[2] * tang[2]);
    if (tangMag > 0.) {
        const double s = 1. / tangMag;
        tang[0] *= s; tang[1] *= s; tang[2] *= s;
        dott = delPi[0] * tang[0] + delPi[1] * tang[1] + delPi[2] * tang[2];
        if (dott < 0.) { tang[0] *= -1.; tang[1] *= -1.; tang[2] *= -1.; dott = -dott; }
    }
    // GetSslideAcDt(-dotn, dott, ...)
    const double NAcDt = -dotn;
    double SslideAcDt;
    if (muStatic > 0. && dott <= muStatic * NAcDt) SslideAcDt = muStatic * NAcDt;
    else SslideAcDt = mu * NAcDt;
    if (SslideAcDt <= 0.) {
        delPi[0] = n[0] * dotn; delPi[1] = n[1] * dotn; delPi[2] = n[2] * dotn;
    } else if (dott > SslideAcDt) {
        delPi[0] = n[0] * dotn; delPi[1] = n[1] * dotn; delPi[2] = n[2] * dotn;
        delPi[0] += tang[0] * SslideAcDt; delPi[1] += tang[1] * SslideAcDt; delPi[2] += tang[2] * SslideAcDt;
    }
    return true;
}

// MaterialContactNode::ContactOnKnownNodes -> CrackVelocityFieldMulti::MaterialContactOnCVF / MaterialContactOnCVFLumped
// (CrackVelocityFieldMulti.cpp:302-674) for nonrigid materials: one thread per node walks the node's active material fields in
// field order, as the reference does (with three or more materials the lumped loop reads the forces the earlier
// iterations changed).  bcOfNode/B give the node's symmetry-plane bits.
__global__ void k_material_contact(Grid g, Nodes N, ContactNodes C, ContactParams cp, VelBCs B, const int *bcOfNode, int callType, double dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nn = g.nnodes;
    if (i >= nn) return;
    int act[MPM_MAX_FIELDS];
    int numMats = 0, rigidFld = -1, numberMaterials = 0;
    bool multiRigid = false;
    double Pc[3] = {0., 0., 0.}, Mc = 0.;
    for (int f = 0; f < cp.nf; f++) {
        const int v = f * nn + i;
        if (cp.rigidMask >> f & 1) {
            if (C.rcnt[v] > 0) { numberMaterials++; if (rigidFld >= 0) multiRigid = true; else rigidFld = f; }
        } else if (N.cnt[v] > 0) {
            Pc[0] += N.pk[0][v]; Pc[1] += N.pk[1][v]; Pc[2] += N.pk[2][v];
            Mc += N.mass[v];
            act[numMats++] = f;
            numberMaterials++;
        }
    }
    if (numberMaterials <= 1) return;
    int sd = 0;
    if (bcOfNode) { const int u = bcOfNode[i]; if (u >= 0) sd = B.symdir[u]; }
    const bool postUpdate = callType == CALL_UPDATE_MOMENTUM;
    const bool useGrad = cp.normalMethod != NORMALS_SPECIFIED;
    double xn[3];
    {   // node coordinates (power-law position cutoff only)
        const int iz = g.dim == 3 ? i / g.zplane : 0, r = i - iz * g.zplane, iy = r / g.yplane, ix = r - iy * g.yplane;
        xn[0] = g.xpts[ix]; xn[1] = g.ypts[iy]; xn[2] = g.dim == 3 ? g.zpts[iz] : 0.;
    }
    if (rigidFld >= 0) {
        // CrackVelocityFieldMulti::RigidMaterialContactOnCVF (:676-955): every nonrigid material of the node against the rigid
        // material (the one of largest contact volume when there are several); only the nonrigid momenta change
        double rigidVolume = C.cvol[rigidFld * nn + i];
        if (multiRigid)
            for (int f = rigidFld + 1; f < cp.nf; f++)
                if ((cp.rigidMask >> f & 1) && C.rcnt[f * nn + i] > 0) {
                    const double tv = C.cvol[f * nn + i];
                    if (tv > rigidVolume) { rigidVolume = tv; rigidFld = f; }
                }
        const int vr = rigidFld * nn + i;
        const double actualRigidVolume = N.mass[vr];
        double rvel[3];
        { const double s = 1. / actualRigidVolume; rvel[0] = N.pk[0][vr] * s; rvel[1] = N.pk[1][vr] * s; rvel[2] = N.pk[2][vr] * s; }
        for (int mi = 0; mi < numMats; mi++) {
            const int fi = act[mi], vi = fi * nn + i;
            const int law = cp.lawKind[fi * cp.nf + rigidFld];
            const double massi = N.mass[vi], mred = massi, voli = C.cvol[vi];
            double delPi[3] = {N.pk[0][vi] * -1., N.pk[1][vi] * -1., N.pk[2][vi] * -1.};
            delPi[0] += rvel[0] * massi; delPi[1] += rvel[1] * massi; delPi[2] += rvel[2] * massi;
            adjust_for_symmetry(sd, delPi);
            if (law != LAW_IGNORE) {
                double gradj[3] = {0., 0., 0.}, normi[3] = {0., 0., 0.}, norm[3] = {0., 0., 0.};
                if (useGrad) {
                    gradj[0] = C.cgrad[0][vr] * -1.; gradj[1] = C.cgrad[1][vr] * -1.; gradj[2] = C.cgrad[2][vr] * -1.;
                    adjust_for_symmetry(sd, gradj);
                    normi[0] = C.cgrad[0][vi]; normi[1] = C.cgrad[1][vi]; normi[2] = C.cgrad[2][vi];
                    adjust_for_symmetry(sd, normi);
                }
                const double jBias = cp.rigidBias;
                switch (cp.normalMethod) {      // GetNormalVector with j = the rigid field (:960-1071)
                case NORMALS_MAXG: {
                    const double magi = sqrt(normi[0] * normi[0] + normi[1] * normi[1] + normi[2] * normi[2]);
                    const double magj = sqrt(gradj[0] * gradj[0] + gradj[1] * gradj[1] + gradj[2] * gradj[2]);
                    if (magi >= jBias * magj) { const double s = 1. / magi; norm[0] = normi[0] * s; norm[1] = normi[1] * s; norm[2] = normi[2] * s; }
                    else { const double s = 1. / magj; norm[0] = gradj[0] * s; norm[1] = gradj[1] * s; norm[2] = gradj[2] * s; }
                    break;
                }
                case NORMALS_MAXV: {
                    if (voli >= rigidVolume) { norm[0] = normi[0]; norm[1] = normi[1]; norm[2] = normi[2]; }
                    else { norm[0] = gradj[0]; norm[1] = gradj[1]; norm[2] = gradj[2]; }
                    const double s = 1. / sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
                    norm[0] *= s; norm[1] *= s; norm[2] *= s;
                    break;
                }
                case NORMALS_AVGG: {
                    norm[0] = normi[0] + gradj[0]; norm[1] = normi[1] + gradj[1]; norm[2] = normi[2] + gradj[2];
                    const double magi = norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2];
                    const double vRatio = (voli + rigidVolume) / rigidVolume;
                    const double magj = gradj[0] * gradj[0] + gradj[1] * gradj[1] + gradj[2] * gradj[2];
                    if (magi >= jBias * magj * vRatio * vRatio) { const double s = 1. / sqrt(magi); norm[0] *= s; norm[1] *= s; norm[2] *= s; }
                    else { const double s = 1. / sqrt(magj); norm[0] = gradj[0] * s; norm[1] = gradj[1] * s; norm[2] = gradj[2] * s; }
                    break;
                }
                case NORMALS_OWNG: {
                    const double s = 1. / sqrt(normi[0] * normi[0] + normi[1] * normi[1] + normi[2] * normi[2]);
                    norm[0] = normi[0] * s; norm[1] = normi[1] * s; norm[2] = normi[2] * s;
                    break;
                }
                default: {
                    norm[0] = cp.normal[0]; norm[1] = cp.normal[1]; norm[2] = cp.normal[2];
                    if (adjust_for_symmetry(sd, norm)) {
                        const double s = 1. / sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
                        norm[0] *= s; norm[1] *= s; norm[2] *= s;
                    }
                    break;
                }
                }
                if (norm[0] != norm[0] || norm[1] != norm[1] || norm[2] != norm[2]) continue;
                const double dotn = delPi[0] * norm[0] + delPi[1] * norm[1] + delPi[2] * norm[2];
                double dispRigid[3], dispi[3];
                { const double s = 1. / actualRigidVolume; dispRigid[0] = C.cdisp[0][vr] * s; dispRigid[1] = C.cdisp[1][vr] * s; dispRigid[2] = C.cdisp[2][vr] * s; }
                adjust_for_symmetry(sd, dispRigid);
                { const double s = 1. / massi; dispi[0] = C.cdisp[0][vi] * s; dispi[1] = C.cdisp[1][vi] * s; dispi[2] = C.cdisp[2][vi] * s; }
                adjust_for_symmetry(sd, dispi);
                const double deln = material_separation(g, cp, dispRigid[0] * norm[0] + dispRigid[1] * norm[1] + dispRigid[2] * norm[2],
                                                        dispi[0] * norm[0] + dispi[1] * norm[1] + dispi[2] * norm[2], norm, xn);
                double delFi[3];
                if (postUpdate) { delFi[0] = N.ftot[0][vi] * -1.; delFi[1] = N.ftot[1][vi] * -1.; delFi[2] = N.ftot[2][vi] * -1.; }
                if (!frictional_delta_momentum(law, cp.lawFriction[fi * cp.nf + rigidFld], cp.lawStatic[fi * cp.nf + rigidFld], delPi, norm, dotn, deln, mred, dt,
                                               postUpdate ? delFi : (const double *)0)) continue;
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double pk = N.pk[c][vi] + delPi[c];
                N.pk[c][vi] = pk;
                if (callType == CALL_UPDATE_MOMENTUM) { N.ftot[c][vi] += delPi[c] * (1. / dt); C.rforce[c][vr] += delPi[c]; }    // + AddContactForce on the rigid field
                else if (callType == CALL_MASS_MOMENTUM) N.pkc[c][vi] = pk;
            }
        }
        return;
    }
    if (numMats <= 1) return;
    const bool doingPairs = numMats == 2 && cp.normalMethod != NORMALS_OWNG;
    const int miMax = doingPairs ? numMats - 1 : numMats;
    // centre-of-mass displacement (or position) and total contact volume of the nonrigid materials
    double dispc[3] = {0., 0., 0.}, volAll = 0.;
    for (int k = 0; k < numMats; k++) {
        const int v = act[k] * nn + i;
        dispc[0] += C.cdisp[0][v]; dispc[1] += C.cdisp[1][v]; dispc[2] += C.cdisp[2][v];
        volAll += C.cvol[v];
    }
    adjust_for_symmetry(sd, dispc);
    { const double s = 1. / Mc; dispc[0] *= s; dispc[1] *= s; dispc[2] *= s; }
    for (int mi = 0; mi < miMax; mi++) {
        const int fi = act[mi], vi = fi * nn + i;
        const double massi = N.mass[vi];
        const double voli = C.cvol[vi];
        const double volj = volAll - voli;
        int fj = -1;
        double gradj[3] = {0., 0., 0.};
        double maxOther = 0.;
        for (int kj = 0; kj < numMats; kj++) {
            if (kj == mi) continue;
            const int jj = act[kj], vj = jj * nn + i;
            const double matVolume = C.cvol[vj];
            if (matVolume > maxOther) { maxOther = matVolume; fj = jj; }
            if (useGrad) {      // GetVolumeGradient(jj, ndptr, &normj, -1.)
                double normj[3] = {C.cgrad[0][vj] * -1., C.cgrad[1][vj] * -1., C.cgrad[2][vj] * -1.};
                adjust_for_symmetry(sd, normj);
                gradj[0] += normj[0]; gradj[1] += normj[1]; gradj[2] += normj[2];
            }
        }
        if (fj < 0) continue;
        const int vj = fj * nn + i;
        const int law = cp.lawKind[fi * cp.nf + fj];
        const double massRatio = massi / Mc;
        double mred = 1. - massRatio;
        double norm[3] = {0., 0., 0.};
        double dotn = 0., deln = 0.;
        double delPi[3] = {N.pk[0][vi] * -1., N.pk[1][vi] * -1., N.pk[2][vi] * -1.};
        delPi[0] += Pc[0] * massRatio; delPi[1] += Pc[1] * massRatio; delPi[2] += Pc[2] * massRatio;
        adjust_for_symmetry(sd, delPi);
        if (law != LAW_IGNORE) {
            // CrackVelocityFieldMulti::GetNormalVector (:960-1071), volume-gradient methods and the specified normal
            double normi[3] = {0., 0., 0.};
            if (useGrad) { normi[0] = C.cgrad[0][vi]; normi[1] = C.cgrad[1][vi]; normi[2] = C.cgrad[2][vi]; adjust_for_symmetry(sd, normi); }
            switch (cp.normalMethod) {
            case NORMALS_MAXG: {
                const double magi = sqrt(normi[0] * normi[0] + normi[1] * normi[1] + normi[2] * normi[2]);
                const double magj = sqrt(gradj[0] * gradj[0] + gradj[1] * gradj[1] + gradj[2] * gradj[2]);
                if (magi >= magj) { const double s = 1. / magi; norm[0] = normi[0] * s; norm[1] = normi[1] * s; norm[2] = normi[2] * s; }
                else { const double s = 1. / magj; norm[0] = gradj[0] * s; norm[1] = gradj[1] * s; norm[2] = gradj[2] * s; }
                break;
            }
            case NORMALS_MAXV: {
                if (voli >= volj) { norm[0] = normi[0]; norm[1] = normi[1]; norm[2] = normi[2]; }
                else { norm[0] = gradj[0]; norm[1] = gradj[1]; norm[2] = gradj[2]; }
                const double s = 1. / sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
                norm[0] *= s; norm[1] *= s; norm[2] *= s;
                break;
            }
            case NORMALS_AVGG: {
                norm[0] = normi[0] + gradj[0]; norm[1] = normi[1] + gradj[1]; norm[2] = normi[2] + gradj[2];
                const double magi = norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2];
                const double s = 1. / sqrt(magi);
                norm[0] *= s; norm[1] *= s; norm[2] *= s;
                break;
            }
            case NORMALS_OWNG: {
                const double s = 1. / sqrt(normi[0] * normi[0] + normi[1] * normi[1] + normi[2] * normi[2]);
                norm[0] = normi[0] * s; norm[1] = normi[1] * s; norm[2] = normi[2] * s;
                break;
            }
            default: {          // NORMALS_SPECIFIED
                norm[0] = cp.normal[0]; norm[1] = cp.normal[1]; norm[2] = cp.normal[2];
                if (adjust_for_symmetry(sd, norm)) {
                    const double s = 1. / sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
                    norm[0] *= s; norm[1] *= s; norm[2] *= s;
                }
                break;
            }
            }
            // a volume gradient normal to a symmetry plane leaves no direction
            if (norm[0] != norm[0] || norm[1] != norm[1] || norm[2] != norm[2]) continue;
            dotn = delPi[0] * norm[0] + delPi[1] * norm[1] + delPi[2] * norm[2];
            double dispa[3];
            { const double s = 1. / massi; dispa[0] = C.cdisp[0][vi] * s; dispa[1] = C.cdisp[1][vi] * s; dispa[2] = C.cdisp[2][vi] * s; }
            adjust_for_symmetry(sd, dispa);
            double delta[3] = {dispc[0] - dispa[0], dispc[1] - dispa[1], dispc[2] - dispa[2]};
            { const double s = 1. / mred; delta[0] *= s; delta[1] *= s; delta[2] *= s; }
            const double dispb[3] = {delta[0] + dispa[0], delta[1] + dispa[1], delta[2] + dispa[2]};
            deln = material_separation(g, cp, dispb[0] * norm[0] + dispb[1] * norm[1] + dispb[2] * norm[2],
                                       dispa[0] * norm[0] + dispa[1] * norm[1] + dispa[2] * norm[2], norm, xn);
            mred *= massi;
            double delFi[3];
            if (postUpdate) {   // (ma Fc / Mc) - Fa, with the forces as they are now (GetCMatFtot)
                double fk[3] = {0., 0., 0.};
                for (int k = 0; k < numMats; k++) {
                    const int v = act[k] * nn + i;
                    fk[0] += N.ftot[0][v]; fk[1] += N.ftot[1][v]; fk[2] += N.ftot[2][v];
                }
                const double s = massi / Mc;
                delFi[0] = fk[0] * s; delFi[1] = fk[1] * s; delFi[2] = fk[2] * s;
                delFi[0] += N.ftot[0][vi] * -1.; delFi[1] += N.ftot[1][vi] * -1.; delFi[2] += N.ftot[2][vi] * -1.;
            }
#ifdef EMU_DEBUG_NODE
            if (i == EMU_DEBUG_NODE) printf("node %d call %d fi %d fj %d law %d norm %.17g %.17g %.17g dotn %.17g deln %.17g mred %.17g delPi %.17g %.17g %.17g\n", i, callType, fi, fj, law, norm[0], norm[1], norm[2], dotn, deln, mred, delPi[0], delPi[1], delPi[2]);
#endif
            if (!frictional_delta_momentum(law, cp.lawFriction[fi * cp.nf + fj], cp.lawStatic[fi * cp.nf + fj], delPi, norm, dotn, deln, mred, dt,
                                           postUpdate ? delFi : (const double *)0)) continue;
        }
        // MatVelocityField::ChangeMatMomentum (MatVelocityField.cpp:197-224)
        for (int side = 0; side < (doingPairs ? 2 : 1); side++) {
            const int v = side == 0 ? vi : vj;
            const double sg = side == 0 ? 1. : -1.;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double dp = side == 0 ? delPi[c] : delPi[c] * sg;
                const double pk = N.pk[c][v] + dp;
                N.pk[c][v] = pk;
                if (callType == CALL_UPDATE_MOMENTUM) N.ftot[c][v] += dp * (1. / dt);
                else if (callType == CALL_MASS_MOMENTUM) N.pkc[c][v] = pk;
            }
        }
    }
}

// Without a transport task the particle update still hands the laws a temperature change: the difference between the particle's
// temperature and the one its last strain update saw (UpdateParticlesTask.cpp:246-251) -- nonzero once, after a start off the
// stress-free temperature.
// With <EnergyCoupling/> the buffered adiabatic rise moves both temperatures first (:229-235); without conduction the difference
// taken afterwards does not see it.
__global__ void k_update_temperature_offsets(int n, Particles P)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (P.dTad) { const double dTad = P.dTad[p]; P.dTad[p] = 0.; P.temp[p] += dTad; P.prevT[p] += dTad; }
    P.dTr[p] = P.temp[p] - P.prevT[p];
    P.prevT[p] = P.temp[p];
}

// =================================================================================================================
// Conduction: the first transport task on the same scatter/gather skeleton (Custom_Tasks/ConductionTask.cpp,
// TransportTask.cpp).  One scalar per node and particle; isothermal energy mode;
// FLIP transport update; no temperature or heat-flux BCs (insulated boundaries); materials without thermal expansion.
// Transport values live on the NODE (NodalPoint::gCond), not on a material velocity field, so these kernels use the
// particle's real node numbers also in multimaterial mode.
// =================================================================================================================

// task 2: ConductionTask::Task1Extrapolation (ConductionTask.cpp:101-112), called from NodalPoint::AddMassMomentum
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_temperature(Grid g, Particles P, const Material *mats, TransportNodes T)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const double mpCv = P.mp[p] * mats[P.mat[p]].p[1];
    const double Tp = P.temp[p];
    particle_nodes_one_field<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        const double CTShape = mpCv * S;
        atomAdd(&T.gT[nd], Tp * CTShape);
        atomAdd(&T.gVCT[nd], CTShape);
    });
}

// a node "has nonrigid particles" when any of its material velocity fields saw one (NodalPoint::NodeHasNonrigidParticles)
__device__ __forceinline__ bool node_has_particles(const Nodes &N, int i, int nnodes, int nf)
{
    for (int f = 0; f < nf; f++) if (N.cnt[f * nnodes + i] > 0) return true;
    return false;
}

// task 1: the reference zeroes the transport field of the nodes on its ACTIVE list only -- the nodes that had particles in the
// previous step (InitializationTask.cpp:52-56 -> NodalPoint::InitializeForTimeStep).  A temperature-BC node without particles
// therefore keeps the value and rate its BC wrote (ImposeValueGridBCs runs on every BC node), and when particles reach it later
// the extrapolation starts from those stale numbers.  Kept: it is what the reference computes.  N.cnt still holds the previous
// step's point counts when this runs.
__global__ void k_transport_zero_active(int nnodes, int nf, Nodes N, TransportNodes T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes || !node_has_particles(N, i, nnodes, nf)) return;
    T.gT[i] = 0.; T.gVCT[i] = 0.; T.gQ[i] = 0.;
}

// task 3: TransportTask::GetTransportNodalValue (TransportTask.cpp:152-161)
__global__ void k_transport_nodal_value(int nnodes, int nf, Nodes N, TransportNodes T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes || !node_has_particles(N, i, nnodes, nf)) return;
    T.gT[i] /= T.gVCT[i];
}

// task 3: TransportTask::GetGradients (TransportTask.cpp:224-272): grad T on the particle from the nodal temperatures
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_transport_gradients(Grid g, Particles P, TransportNodes T)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    double tg[3] = {0., 0., 0.};
    particle_nodes_one_field<DIM, SHAPE, true>(g, P, p, [&](int nd, double, double gx, double gy, double gz) {
        const double Ti = T.gT[nd];
        tg[0] += gx * Ti; tg[1] += gy * Ti;
        if (DIM == 3) tg[2] += gz * Ti;
    });
    P.tgrad[0][p] = tg[0]; P.tgrad[1][p] = tg[1]; P.tgrad[2][p] = tg[2];
}

// task 3: TransportTask::ImposeValueBCs(copyFirst) before the gradients and RestoreValueBCs after them (TransportTask.cpp:167-222,
// called from TransportBCsAndGradients :550-565): on a node with active temperature BCs the gradients see the sum of the BC values
__global__ void k_temp_bcs_impose(TempBCs Q, TransportNodes T, int restore)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= Q.nUnique) return;
    const int nd = Q.node[u];
    bool any = false;
    double sum = 0.;
    for (int e = Q.start[u]; e < Q.start[u + 1]; e++) if (Q.active[e]) { any = true; sum += Q.value[e]; }
    if (!any) return;
    if (restore) T.gT[nd] = Q.saved[u];
    else { Q.saved[u] = T.gT[nd]; T.gT[nd] = sum; }
}

// task 7: TransportTask::ImposeValueGridBCs in the momentum update (TransportTask.cpp:316-404): the nodal value becomes the sum of
// the BC values and the rate what takes it there, so that the particles' FLIP update sees a consistent pair
__global__ void k_temp_bcs_grid(TempBCs Q, TransportNodes T, double dt)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= Q.nUnique) return;
    const int nd = Q.node[u];
    bool first = true;
    double gT = T.gT[nd], gQ = T.gQ[nd];
    for (int e = Q.start[u]; e < Q.start[u + 1]; e++) {
        if (!Q.active[e]) continue;
        if (first) { gQ += -gT / dt; gT = 0.; first = false; }
        const double bcT = Q.value[e];
        gT += bcT;
        gQ += bcT / dt;
    }
    if (!first) { T.gT[nd] = gT; T.gQ[nd] = gQ; }
}

// the same two steps for the temperature BCs rigid particles made this step (they follow the grid BCs in the reference's list and sit
// on nodes without one: ProjectRigidBCsTask.cpp:202; one BC per node)
__global__ void k_rigid_temp_bcs_impose(int nnodes, RigidBCs R, TransportNodes T, int restore)
{
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= nnodes) return;
    const int o = R.ownerT[nd];
    if (o == RIGID_NONE) return;
    if (restore) T.gT[nd] = R.savedT[nd];
    else { R.savedT[nd] = T.gT[nd]; T.gT[nd] = R.ptemp[o]; }
}

__global__ void k_rigid_temp_bcs_grid(int nnodes, RigidBCs R, TransportNodes T, double dt)
{
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= nnodes) return;
    const int o = R.ownerT[nd];
    if (o == RIGID_NONE) return;
    const double bcT = R.ptemp[o];
    double gQ = T.gQ[nd] + -T.gT[nd] / dt;
    gQ += bcT / dt;
    T.gT[nd] = bcT; T.gQ[nd] = gQ;
}

// task 5: ConductionTask::AddForces -> MatPoint3D::FCond / MatPoint2D::FCond (MatPoint3D.cpp:280-287, MatPoint2D.cpp:272-278)
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_p2g_conduction(Grid g, Particles P, TransportNodes T)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    const double k = T.kcond[P.mat[p]];
    double relVol;          // MatPoint3D/2D::GetRelativeVolume: det F
    if (DIM == 3) {
        const double F00 = P.F[0][p], F01 = P.F[1][p], F02 = P.F[2][p], F10 = P.F[3][p], F11 = P.F[4][p], F12 = P.F[5][p],
                     F20 = P.F[6][p], F21 = P.F[7][p], F22 = P.F[8][p];
        relVol = F00 * (F11 * F22 - F21 * F12) - F01 * (F10 * F22 - F20 * F12) + F02 * (F10 * F21 - F20 * F11);
    } else {
        relVol = P.F[8][p] * (P.F[0][p] * P.F[4][p] - P.F[3][p] * P.F[1][p]);
    }
    const double c = -P.mp[p] * relVol;
    const double qx = k * P.tgrad[0][p], qy = k * P.tgrad[1][p], qz = DIM == 3 ? k * P.tgrad[2][p] : 0.;
    particle_nodes_one_field<DIM, SHAPE, true>(g, P, p, [&](int nd, double, double gx, double gy, double gz) {
        atomAdd(&T.gQ[nd], DIM == 3 ? c * (qx * gx + qy * gy + qz * gz) : c * (qx * gx + qy * gy));
    });
}

// task 7: TransportTask::UpdateTransport (TransportTask.cpp:419-428)
__global__ void k_transport_update(int nnodes, int nf, Nodes N, TransportNodes T, double dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes || !node_has_particles(N, i, nnodes, nf)) return;
    const double rate = T.gQ[i] / T.gVCT[i];
    T.gQ[i] = rate;
    T.gT[i] += rate * dt;
}

// task 8: the transport part of UpdateParticlesTask (UpdateParticlesTask.cpp:134-245): value and rate from the grid, change of
// the grid-extrapolated temperature since the last update (TransportTask::GetDeltaValue), FLIP update of the particle
// temperature (MoveTransportValue), heat energy and entropy of the conducted heat.
template <int DIM, int SHAPE>
__global__ void __launch_bounds__(TASK_THREADS) k_update_temperature(Grid g, Particles P, const Material *mats, TransportNodes T, double dt)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nNR) return;
    double value = 0., rate = 0.;
    particle_nodes_one_field<DIM, SHAPE, false>(g, P, p, [&](int nd, double S, double, double, double) {
        value += T.gT[nd] * S;
        rate += T.gQ[nd] * S;
    });
    const double prev = P.prevT[p];
    const double dTcond = value - prev;
    double prevNew = value, dTres = dTcond, Tp = P.temp[p] + dt * rate;
    if (P.dTad) {           // <EnergyCoupling/>: the adiabatic rise the laws buffered moves both temperatures and res.dT, not dTcond (:229-235)
        const double dTad = P.dTad[p];
        P.dTad[p] = 0.;
        Tp += dTad; prevNew += dTad; dTres += dTad;
    }
    P.prevT[p] = prevNew;
    P.dTr[p] = dTres;                   // res.dT of the next strain updates (mpmptr->dTrans = res, UpdateParticlesTask.cpp:261)
    P.temp[p] = Tp;
    const double cv = mats[P.mat[p]].p[1];
    P.heat[p] += cv * dTcond;
    P.entropy[p] += cv * log(prevNew / (prevNew - dTcond));
}
