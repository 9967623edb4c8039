// TEST INFRASTRUCTURE: the per-task kernels of libmpmgpu (csrc/kernels_task.cuh, with shape.cuh and materials.cuh) compiled
// for the host through tests/devlaws/stub and run one CUDA thread after the other, in the task order of capi.cu::step_by_tasks
// (= the reference's MPMTask list), with grid velocity BCs (constant values) and rigid-BC particles (constant velocities).  tests/test_device_step_cpu.py
// compares the result with the golden dumps of the unmodified reference: the CUDA SOURCE of the general path is checked on
// every CPU run; the compiled kernels and the host orchestration of capi.cu are checked by the GPU parity tests.
// This is not a product path: nothing outside tests/ builds or loads it.
#include "kernels_task.cuh"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

struct EmuSim {
    Grid g;
    Particles P;
    Nodes N;
    StepParams sp;
    StatusFlags flags;
    std::vector<Material> mats;
    std::vector<double> xpts, ypts, zpts, ppool, npool, cpXi, cpWg;
    std::vector<int> ipool, ncnt, cpElem;
    // rigid-BC particles (host order, after the non-rigid ones) and the BCs they make (capi.cu: ctx->PR, ctx->R)
    Particles PR;
    RigidBCs R;
    std::vector<double> rpool, rcpXi, rcpWg;
    std::vector<int> ripool, rcpElem, owner;
    std::vector<unsigned char> fixedBits;
    // grid velocity BCs grouped by node, as capi.cu::mpmgpu_set_velocity_bcs builds them
    VelBCs B;
    bool hasBCs;
    std::vector<int> bcNode, bcStart, bcSym, bcActive, bcOfNode;
    std::vector<double> bcNorm, bcValue, bcRatio;
    std::vector<int> bcRefl, bcOrder;
    std::vector<double> bcReact, rigidReact;      // capi.cu: dReaction, dRigidReaction (always kept here)
    int dim, shape, n;
    bool largeRotation;
    long long mstep;
    // multimaterial mode (capi.cu: ctx->multimaterial, nf, nvn, C, cp)
    bool multimaterial = false;
    int nf = 1, nvn = 0;
    ContactNodes C;
    ContactParams cp;
    std::vector<double> cpool, origpos;
    std::vector<int> foff, fieldOfMat, foffR, rcnt;
    std::vector<double> rforce;
    // conduction (capi.cu: ctx->conduction, T)
    bool conduction = false, thermal = false;
    TransportNodes T;
    std::vector<double> tpool, kcond, temps;
    TempBCs Q;
    std::vector<int> tbcNode, tbcStart, tbcActive; std::vector<double> tbcValue, tbcSaved;
    // particle traction BCs (capi.cu: TB)
    TractionBCs TB; std::vector<int> trStart, trFace, trDir; std::vector<double> trValue; double thickness = 1.;
    TractionBCs HF; std::vector<int> hfStart, hfFace, hfDir; std::vector<double> hfValue;      // heat fluxes (capi.cu: flux)
    // temperature BCs of rigid particles (capi.cu: rigidTemp, R.ownerT / ptemp / fixedT / savedT)
    std::vector<int> ownerT; std::vector<double> rigidTemps, savedT; std::vector<unsigned char> fixedTemp;
};

struct HostArrays {
    const double *pos, *vel, *mp, *lp; const int *inElem, *matnum;
    const double *sp, *pressure, *ep, *wrot, *eplast, *energies, *history; const int *crossings;
};

// one particle set over its own pools: the fields of capi.cu::bind_particles, filled from host arrays [component][N] at [off, off+cnt)
void fill_set(Particles &P, std::vector<double> &pool, std::vector<int> &ipool, std::vector<int> &cpElem, std::vector<double> &cpXi,
              std::vector<double> &cpWg, int dim, int shape, size_t N, size_t off, size_t cnt, const HostArrays &H)
{
    const int ND = 3 + 3 + 1 + 3 + 3 + 9 + 6 + 1 + 6 + 6 + MPM_MAX_HISTORY + 3 + 3;
    const size_t C = cnt ? cnt : 1;
    pool.assign(C * ND, 0.);
    ipool.assign(C * 5, 0);
    memset(&P, 0, sizeof P);
    double *d = pool.data();
    auto take = [&]() { double *r = d; d += C; return r; };
    for (int c = 0; c < 3; c++) P.pos[c] = take();
    for (int c = 0; c < 3; c++) P.vel[c] = take();
    P.mp = take();
    for (int c = 0; c < 3; c++) P.lp[c] = take();
    for (int c = 0; c < 3; c++) P.ncpos[c] = take();
    for (int c = 0; c < 9; c++) P.F[c] = take();
    for (int c = 0; c < 6; c++) P.sp[c] = take();
    P.pressure = take();
    for (int c = 0; c < 6; c++) P.eplast[c] = take();
    P.work = take(); P.res = take(); P.heat = take(); P.entropy = take(); P.plast = take(); P.prevT = take();
    for (int c = 0; c < MPM_MAX_HISTORY; c++) P.hist[c] = take();
    for (int c = 0; c < 3; c++) P.pfext[c] = take();
    for (int c = 0; c < 3; c++) P.acc[c] = take();
    int *ii = ipool.data();
    P.elem = ii; P.mat = ii + C; P.cross = ii + 2 * C; P.orig = ii + 3 * C; P.key = ii + 4 * C;
    P.n = (int)cnt; P.nNR = (int)cnt;
    for (size_t p = 0; p < cnt; p++) {
        const size_t q = off + p;
        for (int c = 0; c < 3; c++) { P.pos[c][p] = H.pos[c * N + q]; P.vel[c][p] = H.vel[c * N + q]; P.lp[c][p] = H.lp[c * N + q]; }
        P.mp[p] = H.mp[q];
        // MatPoint3D::GetDeformationGradient from ep + wrot (MatPoint3D.cpp:363-376), as capi.cu::k_epwrot_to_F
        P.F[0][p] = 1. + H.ep[q]; P.F[4][p] = 1. + H.ep[N + q]; P.F[8][p] = 1. + H.ep[2 * N + q];
        P.F[1][p] = 0.5 * (H.ep[5 * N + q] - H.wrot[q]); P.F[3][p] = 0.5 * (H.ep[5 * N + q] + H.wrot[q]);
        if (dim == 3) {
            P.F[2][p] = 0.5 * (H.ep[4 * N + q] - H.wrot[N + q]); P.F[6][p] = 0.5 * (H.ep[4 * N + q] + H.wrot[N + q]);
            P.F[5][p] = 0.5 * (H.ep[3 * N + q] - H.wrot[2 * N + q]); P.F[7][p] = 0.5 * (H.ep[3 * N + q] + H.wrot[2 * N + q]);
        }
        for (int c = 0; c < 6; c++) { P.sp[c][p] = H.sp[c * N + q]; P.eplast[c][p] = H.eplast[c * N + q]; }
        P.pressure[p] = H.pressure[q];
        P.work[p] = H.energies[q]; P.res[p] = H.energies[N + q]; P.heat[p] = H.energies[2 * N + q]; P.entropy[p] = H.energies[3 * N + q];
        P.plast[p] = H.energies[4 * N + q]; P.prevT[p] = H.energies[5 * N + q];
        for (int c = 0; c < MPM_MAX_HISTORY; c++) P.hist[c][p] = H.history[c * N + q];
        P.elem[p] = H.inElem[q]; P.mat[p] = H.matnum[q] - 1; P.cross[p] = H.crossings[q]; P.orig[p] = (int)q;
    }
    if (SHAPE_IS_CPDI(shape)) {
        const int nc = dim == 3 ? 8 : (SHAPE_IS_QCPDI(shape) ? 9 : 4);
        cpElem.assign(C * nc, 0); cpXi.assign(C * (nc * 3 + 12), 0.); cpWg.assign(C * nc * 3, 0.);
        P.cpElem = cpElem.data(); P.cpXi = cpXi.data(); P.cpWg = cpWg.data(); P.cpStride = C;
        P.cpDom = cpXi.data() + C * nc * 3;
    }
}

inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

// node pool: mass, pk, ftot, vk, pkCopy, v*prev, v*next (capi.cu::alloc_node_arrays); nn = fields x nodes
void bind_nodes(EmuSim *S, size_t nn)
{
    S->npool.assign(nn * 22, 0.);
    S->ncnt.assign(nn, 0);
    S->nvn = (int)nn;
    Nodes &Nd = S->N;
    memset(&Nd, 0, sizeof Nd);
    double *qn = S->npool.data();
    Nd.mass = qn; qn += nn;
    for (int c = 0; c < 3; c++) { Nd.pk[c] = qn; qn += nn; }
    for (int c = 0; c < 3; c++) { Nd.ftot[c] = qn; qn += nn; }
    for (int c = 0; c < 3; c++) { Nd.vk[c] = qn; qn += nn; }
    for (int c = 0; c < 3; c++) { Nd.pkc[c] = qn; qn += nn; }
    for (int c = 0; c < 3; c++) { Nd.vsp[c] = qn; qn += nn; }
    for (int c = 0; c < 3; c++) { Nd.vsn[c] = qn; qn += nn; }
    Nd.cnt = S->ncnt.data();
}

#define DISPATCH(KERNEL, n, ...) do { \
    const int grid_ = nblk((n), TASK_THREADS); \
    if (S->dim == 3) { \
        if (S->shape == SHAPE_UGIMP) EMU_LAUNCH((KERNEL<3, SHAPE_UGIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_B2SPLINE) EMU_LAUNCH((KERNEL<3, SHAPE_B2SPLINE>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_B2GIMP) EMU_LAUNCH((KERNEL<3, SHAPE_B2GIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_B2CPDI) EMU_LAUNCH((KERNEL<3, SHAPE_B2CPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_LCPDI) EMU_LAUNCH((KERNEL<3, SHAPE_LCPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_LCPDI_MERGED) EMU_LAUNCH((KERNEL<3, SHAPE_LCPDI_MERGED>), grid_, TASK_THREADS, __VA_ARGS__); \
        else EMU_LAUNCH((KERNEL<3, SHAPE_LINEAR>), grid_, TASK_THREADS, __VA_ARGS__); \
    } else { \
        if (S->shape == SHAPE_UGIMP) EMU_LAUNCH((KERNEL<2, SHAPE_UGIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_B2SPLINE) EMU_LAUNCH((KERNEL<2, SHAPE_B2SPLINE>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_B2GIMP) EMU_LAUNCH((KERNEL<2, SHAPE_B2GIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_B2CPDI) EMU_LAUNCH((KERNEL<2, SHAPE_B2CPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_LCPDI) EMU_LAUNCH((KERNEL<2, SHAPE_LCPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_LCPDI_MERGED) EMU_LAUNCH((KERNEL<2, SHAPE_LCPDI_MERGED>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_QCPDI) EMU_LAUNCH((KERNEL<2, SHAPE_QCPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
        else if (S->shape == SHAPE_QCPDI_MERGED) EMU_LAUNCH((KERNEL<2, SHAPE_QCPDI_MERGED>), grid_, TASK_THREADS, __VA_ARGS__); \
        else EMU_LAUNCH((KERNEL<2, SHAPE_LINEAR>), grid_, TASK_THREADS, __VA_ARGS__); \
    } } while (0)

void apply_bcs(EmuSim *S, int pass, int adjustSym)
{
    if (S->hasBCs && S->B.nUnique > 0) EMU_LAUNCH(k_velocity_bcs, nblk((long long)S->B.nUnique * S->nf, 128), 128, S->B, S->N, pass, S->sp.dt, adjustSym, S->nf, S->g.nnodes);
    if (S->R.on && adjustSym != 2) EMU_LAUNCH(k_rigid_velocity_bcs, nblk(S->nvn, 256), 256, S->nvn, S->R, S->N, pass, S->sp.dt);
}

// capi.cu::contact_extrapolation / material_contact
void contact_extrapolation(EmuSim *S, bool rigidToo = false)
{
    if (!S->multimaterial) return;
    if (rigidToo && S->cp.rigidMask && S->PR.n > 0)
        DISPATCH(k_p2g_rigid_contact, S->PR.n, S->g, S->PR, S->mats.data(), S->N, S->C, S->origpos.data(), (size_t)S->n, S->cp.byDisplacements,
                 S->cp.normalMethod != NORMALS_SPECIFIED ? 1 : 0, &S->flags);
    DISPATCH(k_p2g_contact_terms, S->P.nNR, S->g, S->P, S->mats.data(), S->C, S->origpos.data(), (size_t)S->n, S->cp.byDisplacements,
             S->cp.normalMethod != NORMALS_SPECIFIED ? 1 : 0);
}

void material_contact(EmuSim *S, int callType)
{
    if (!S->multimaterial) return;
    if (getenv("EMU_SKIP_CONTACT") && atoi(getenv("EMU_SKIP_CONTACT")) == callType + 1 && S->mstep + 1 == atoi(getenv("EMU_SKIP_STEP"))) return;      // debugging aid
    EMU_LAUNCH(k_material_contact, nblk(S->g.nnodes, 128), 128, S->g, S->N, S->C, S->cp, S->B, S->hasBCs ? S->bcOfNode.data() : (const int *)NULL, callType, S->sp.dt);
}


void xpic_extrapolation(EmuSim *S, int particleUpdate)
{
    const int nn = S->nvn, fmpm = S->sp.usingFMPM ? 1 : 0;
    EMU_LAUNCH(k_xpic_init, nblk(nn, 256), 256, nn, S->N, S->sp.dt, fmpm);
    for (int k = 2; k <= S->sp.xpicOrder; k++) {
        DISPATCH(k_xpic_iterate, S->P.nNR, S->g, S->P, S->N);
        EMU_LAUNCH(k_xpic_finish, nblk(nn, 256), 256, nn, S->N, S->B, S->hasBCs ? S->bcOfNode.data() : (const int *)NULL, S->R, S->sp.dt, particleUpdate, fmpm);
    }
}

void strain_update(EmuSim *S, double strainTime, bool postUpdate)
{
    if (S->sp.usingFMPM && S->sp.xpicOrder > 1) {
        if (!postUpdate || !S->sp.skipPost) xpic_extrapolation(S, 0);
    } else
        EMU_LAUNCH(k_grid_velocity, nblk(S->nvn, 256), 256, S->nvn, S->N);
    const double dTscale = S->sp.method == METHOD_USAVG ? (postUpdate ? 1.0 - S->sp.fractionUSF : S->sp.fractionUSF) : 1.0;
    if (S->largeRotation) DISPATCH(k_update_strains_lr, S->P.nNR, S->g, S->P, S->N, S->mats.data(), strainTime, dTscale);
    else DISPATCH(k_update_strains, S->P.nNR, S->g, S->P, S->N, S->mats.data(), strainTime, dTscale);
}

// task numbers as in capi.cu: 0 initialization ... 9 reset elements
void run_task(EmuSim *S, int t)
{
    const int nn = S->nvn;
    switch (t) {
    case 0:
        if (S->conduction) EMU_LAUNCH(k_transport_zero_active, nblk(S->g.nnodes, 256), 256, S->g.nnodes, S->nf, S->N, S->T);
        std::fill(S->npool.begin(), S->npool.end(), 0.);
        std::fill(S->ncnt.begin(), S->ncnt.end(), 0);
        std::fill(S->cpool.begin(), S->cpool.end(), 0.);
        std::fill(S->rcnt.begin(), S->rcnt.end(), 0);
        DISPATCH(k_init_particles, S->P.n, S->g, S->P, &S->flags);
        break;
    case 1:
        DISPATCH(k_p2g_mass_momentum, S->P.nNR, S->g, S->P, S->N);
        if (S->conduction) DISPATCH(k_p2g_temperature, S->P.nNR, S->g, S->P, S->mats.data(), S->T);
        contact_extrapolation(S, true);
        break;
    case 2: {
        EMU_LAUNCH(k_copy_momenta, nblk(nn, 256), 256, nn, S->N);
        material_contact(S, CALL_MASS_MOMENTUM);
        const bool hasUSF = S->sp.method == METHOD_USF || S->sp.method == METHOD_USAVG;
        apply_bcs(S, PASS_MASS_MOMENTUM, hasUSF ? 1 : 2);
        if (S->conduction) {
            EMU_LAUNCH(k_transport_nodal_value, nblk(S->g.nnodes, 256), 256, S->g.nnodes, S->nf, S->N, S->T);
            if (S->Q.nUnique > 0) EMU_LAUNCH(k_temp_bcs_impose, nblk(S->Q.nUnique, 128), 128, S->Q, S->T, 0);
            if (S->R.ownerT) EMU_LAUNCH(k_rigid_temp_bcs_impose, nblk(S->g.nnodes, 256), 256, S->g.nnodes, S->R, S->T, 0);
            DISPATCH(k_transport_gradients, S->P.nNR, S->g, S->P, S->T);
            if (S->Q.nUnique > 0) EMU_LAUNCH(k_temp_bcs_impose, nblk(S->Q.nUnique, 128), 128, S->Q, S->T, 1);
            if (S->R.ownerT) EMU_LAUNCH(k_rigid_temp_bcs_impose, nblk(S->g.nnodes, 256), 256, S->g.nnodes, S->R, S->T, 1);
        }
        break;
    }
    case 3:
        if (S->sp.method != METHOD_USL) strain_update(S, S->sp.method == METHOD_USAVG ? S->sp.dtStrainFirst : S->sp.dt, false);
        break;
    case 4:
        DISPATCH(k_p2g_forces, S->P.nNR, S->g, S->P, S->N, 0);
        if (S->conduction) DISPATCH(k_p2g_conduction, S->P.nNR, S->g, S->P, S->T);
        break;
    case 5:
        if (S->TB.n > 0) {      // capi.cu: particle_tractions
            const int cpdi = SHAPE_IS_CPDI(S->shape) ? 1 : 0;
            const int spline = S->shape == SHAPE_B2SPLINE || S->shape == SHAPE_B2GIMP || S->shape == SHAPE_B2CPDI ? 1 : 0;
            if (S->dim == 3) EMU_LAUNCH((k_particle_tractions<3>), nblk(S->P.nNR, TASK_THREADS), TASK_THREADS, S->g, S->P, S->N, S->TB, cpdi, S->thickness, S->nf, &S->flags, (double *)NULL, spline);
            else EMU_LAUNCH((k_particle_tractions<2>), nblk(S->P.nNR, TASK_THREADS), TASK_THREADS, S->g, S->P, S->N, S->TB, cpdi, S->thickness, S->nf, &S->flags, (double *)NULL, spline);
        }
        EMU_LAUNCH(k_post_forces, nblk(nn, 256), 256, nn, S->N, S->sp);
        std::fill(S->bcReact.begin(), S->bcReact.end(), 0.); std::fill(S->rigidReact.begin(), S->rigidReact.end(), 0.);     // capi.cu: reactions_zero
        apply_bcs(S, PASS_GRID_FORCES, 0);
        if (S->conduction && S->HF.n > 0) {      // capi.cu: particle_heat_fluxes
            const int cpdi = SHAPE_IS_CPDI(S->shape) ? 1 : 0;
            const int spline = S->shape == SHAPE_B2SPLINE || S->shape == SHAPE_B2GIMP || S->shape == SHAPE_B2CPDI ? 1 : 0;
            if (S->dim == 3) EMU_LAUNCH((k_particle_tractions<3>), nblk(S->P.nNR, TASK_THREADS), TASK_THREADS, S->g, S->P, S->N, S->HF, cpdi, S->thickness, S->nf, &S->flags, S->T.gQ, spline);
            else EMU_LAUNCH((k_particle_tractions<2>), nblk(S->P.nNR, TASK_THREADS), TASK_THREADS, S->g, S->P, S->N, S->HF, cpdi, S->thickness, S->nf, &S->flags, S->T.gQ, spline);
        }
        break;
    case 6:
        EMU_LAUNCH(k_update_momenta, nblk(nn, 256), 256, nn, S->N, S->sp.dt);
        if (S->conduction) EMU_LAUNCH(k_transport_update, nblk(S->g.nnodes, 256), 256, S->g.nnodes, S->nf, S->N, S->T, S->sp.dt);
        material_contact(S, CALL_UPDATE_MOMENTUM);
        if (S->sp.xpicOrder <= 1) apply_bcs(S, PASS_UPDATE_MOMENTUM, 0);
        if (S->conduction && S->Q.nUnique > 0) EMU_LAUNCH(k_temp_bcs_grid, nblk(S->Q.nUnique, 128), 128, S->Q, S->T, S->sp.dt);
        if (S->conduction && S->R.ownerT) EMU_LAUNCH(k_rigid_temp_bcs_grid, nblk(S->g.nnodes, 256), 256, S->g.nnodes, S->R, S->T, S->sp.dt);
        break;
    case 7: {
        if (S->sp.xpicOrder > 1) xpic_extrapolation(S, 1);
        else EMU_LAUNCH(k_grid_velocity, nblk(nn, 256), 256, nn, S->N);
        int m = S->sp.xpicOrder;
        if (!S->sp.usingFMPM) m = -m;
        DISPATCH(k_update_particles, S->P.nNR, S->g, S->P, S->N, S->mats.data(), S->sp, m);
        if (S->conduction) DISPATCH(k_update_temperature, S->P.nNR, S->g, S->P, S->mats.data(), S->T, S->sp.dt);
        else if (S->thermal) EMU_LAUNCH(k_update_temperature_offsets, nblk(S->P.nNR, 256), 256, S->P.nNR, S->P);
        if (S->PR.n > 0) {
            if (S->dim == 3) EMU_LAUNCH(k_move_rigid<3>, nblk(S->PR.n, 128), 128, S->PR, S->sp.dt);
            else EMU_LAUNCH(k_move_rigid<2>, nblk(S->PR.n, 128), 128, S->PR, S->sp.dt);
        }
        break;
    }
    case 8:
        if (S->sp.method == METHOD_USF) break;
        if (!S->sp.skipPost) {
            if (S->multimaterial) EMU_LAUNCH(k_rezero_fields_task6, nblk(nn, 256), 256, S->g.nnodes, S->nf, S->cp.rigidMask, S->N, S->C, S->sp.dt);
            else EMU_LAUNCH(k_rezero_momenta, nblk(nn, 256), 256, nn, S->N);
            DISPATCH(k_p2g_momentum_last, S->P.nNR, S->g, S->P, S->N);
            contact_extrapolation(S);
            material_contact(S, CALL_UPDATE_STRAINS_LAST);
            apply_bcs(S, PASS_UPDATE_STRAINS_LAST, 0);
        }
        strain_update(S, S->sp.method == METHOD_USAVG ? S->sp.dtStrainLast : S->sp.dt, true);
        break;
    case 9:
        if (S->dim == 3) EMU_LAUNCH(k_reset_elements<3>, nblk(S->P.n, TASK_THREADS), TASK_THREADS, S->g, S->P, &S->flags, S->sp.dt);
        else EMU_LAUNCH(k_reset_elements<2>, nblk(S->P.n, TASK_THREADS), TASK_THREADS, S->g, S->P, &S->flags, S->sp.dt);
        if (S->PR.n > 0) {
            if (S->dim == 3) EMU_LAUNCH(k_reset_elements<3>, nblk(S->PR.n, TASK_THREADS), TASK_THREADS, S->g, S->PR, &S->flags, S->sp.dt);
            else EMU_LAUNCH(k_reset_elements<2>, nblk(S->PR.n, TASK_THREADS), TASK_THREADS, S->g, S->PR, &S->flags, S->sp.dt);
        }
        S->mstep++;
        break;
    case 10:            // ProjectRigidBCsTask: between mass/momentum and post-extrapolation (capi.cu::t_project_rigid_bcs)
        if (!S->R.on) break;
        std::fill(S->owner.begin(), S->owner.end(), RIGID_NONE);
        std::fill(S->ownerT.begin(), S->ownerT.end(), RIGID_NONE);
        DISPATCH(k_project_rigid_bcs, S->PR.n, S->g, S->PR, S->mats.data(), S->R, &S->flags);
        break;
    }
}

} // namespace

// Inputs follow include/mpmgpu.h: cfg fields by value, particle arrays [component][n].  Returns an opaque handle.
extern "C" void *emu_create(int np, int horiz, int vert, int depth, const double *xpts, const double *ypts, const double *zpts,
                            double gridx, double gridy, double gridz, int shape, double rcrit, int method, int skipPost, double fractionUSF,
                            int xpicOrder, int usingFMPM, double gridAlpha, double particleAlpha, const double *gravity,
                            double dt, double dtFirst, double dtLast,
                            int nmat, const int *kinds, const int *nhist, const double *params,
                            int n, int nNR, const double *pos, const double *vel, const double *mp, const double *lp, const int *inElem, const int *matnum,
                            const double *sp, const double *pressure, const double *ep, const double *wrot, const double *eplast,
                            const double *energies, const double *history, const int *crossings)
{
    EmuSim *S = new EmuSim;
    S->dim = np == NP_THREED ? 3 : 2;
    S->shape = shape; S->n = n; S->mstep = 0;
    S->hasBCs = false;
    memset(&S->B, 0, sizeof S->B);
    Grid &g = S->g;
    memset(&g, 0, sizeof g);
    g.dim = S->dim; g.np = np; g.horiz = horiz; g.vert = vert; g.depth = S->dim == 3 ? depth : 1;
    g.yplane = horiz + 1; g.zplane = (horiz + 1) * (vert + 1);
    g.nnodes = g.zplane * (S->dim == 3 ? depth + 1 : 1);
    g.nelems = g.horiz * g.vert * g.depth;
    S->xpts.assign(xpts, xpts + horiz + 1); S->ypts.assign(ypts, ypts + vert + 1);
    if (S->dim == 3) S->zpts.assign(zpts, zpts + depth + 1); else S->zpts.assign(2, 0.);
    g.xpts = S->xpts.data(); g.ypts = S->ypts.data(); g.zpts = S->zpts.data();
    g.gx = gridx; g.gy = gridy; g.gz = gridz;
    g.xmin = xpts[0]; g.ymin = ypts[0]; g.zmin = S->dim == 3 ? zpts[0] : 0.;
    g.rcrit = rcrit;
    g.lpUniform = 0;
    S->mats.resize(nmat);
    S->largeRotation = false;
    for (int i = 0; i < nmat; i++) {
        S->mats[i].kind = kinds[i]; S->mats[i].nhist = nhist[i];
        memcpy(S->mats[i].p, params + (size_t)i * MPM_MAT_NPARAMS, sizeof(double) * MPM_MAT_NPARAMS);
        S->mats[i].p[6] = S->dim == 3 ? (gridx + gridy + gridz) / 3. : (gridx + gridy) / 2.;       // as mpmgpu_set_materials
        if (S->mats[i].p[7] != 0. || kinds[i] == MAT_MOONEY || (kinds[i] == MAT_ISOPLASTICITY && S->mats[i].p[16] > 1.)) S->largeRotation = true;       // extended law dispatch, as mpmgpu_set_materials
    }
    StepParams &q = S->sp;
    memset(&q, 0, sizeof q);
    q.dt = dt; q.dtStrainFirst = dtFirst; q.dtStrainLast = dtLast; q.fractionUSF = fractionUSF;
    q.gridAlpha = gridAlpha; q.particleAlpha = particleAlpha;
    q.grav[0] = gravity[0]; q.grav[1] = gravity[1]; q.grav[2] = gravity[2];
    q.hasGravity = gravity[0] != 0. || gravity[1] != 0. || gravity[2] != 0.;
    q.method = method; q.skipPost = skipPost; q.xpicOrder = xpicOrder; q.usingFMPM = usingFMPM;
    memset(&S->flags, 0, sizeof S->flags);
    // particle pools: the non-rigid particles [0, nNR) and the rigid-BC particles [nNR, n), as capi.cu::mpmgpu_upload_particles
    const HostArrays H = {pos, vel, mp, lp, inElem, matnum, sp, pressure, ep, wrot, eplast, energies, history, crossings};
    fill_set(S->P, S->ppool, S->ipool, S->cpElem, S->cpXi, S->cpWg, S->dim, shape, (size_t)n, 0, (size_t)nNR, H);
    S->P.nNR = nNR;
    fill_set(S->PR, S->rpool, S->ripool, S->rcpElem, S->rcpXi, S->rcpWg, S->dim, shape, (size_t)n, (size_t)nNR, (size_t)(n - nNR), H);
    S->PR.nNR = 0;
    // BCs made by the rigid particles (capi.cu: ctx->R)
    RigidBCs &R = S->R;
    memset(&R, 0, sizeof R);
    R.on = n - nNR > 0 ? 1 : 0;
    S->owner.assign(3 * (size_t)g.nnodes, RIGID_NONE);
    S->fixedBits.assign((size_t)g.nnodes, 0);
    for (int dd = 0; dd < 3; dd++) { R.owner[dd] = S->owner.data() + (size_t)dd * g.nnodes; R.vel[dd] = S->PR.vel[dd]; }
    R.fixedBits = S->fixedBits.data();
    for (int p = nNR; p < n; p++) if (S->mats[matnum[p] - 1].p[9] != 0.) R.mirrored = 1;
    R.mat = S->PR.mat; R.mats = S->mats.data();
    R.stride[0] = 1; R.stride[1] = g.yplane; R.stride[2] = g.zplane; R.nnodes = g.nnodes;
    S->rigidReact.assign(3 * S->mats.size(), 0.);
    R.reaction = S->rigidReact.data();
    S->B.reaction = NULL;
    bind_nodes(S, (size_t)g.nnodes);
    memset(&S->C, 0, sizeof S->C); memset(&S->cp, 0, sizeof S->cp); memset(&S->Q, 0, sizeof S->Q); memset(&S->TB, 0, sizeof S->TB); memset(&S->HF, 0, sizeof S->HF);
    return S;
}

// capi.cu::mpmgpu_set_multimaterial: nf fields per node (field-major node arrays), contact extrapolations, particle field offsets
extern "C" void emu_set_multimaterial(void *h, int nf, const int *fieldOfMat, int normalMethod, int byDisplacements, double positionCutoff,
                                      const double *normal, const int *lawKind, const double *lawFriction, const double *lawStatic, const double *origpos,
                                      double rigidBias)
{
    EmuSim *S = (EmuSim *)h;
    S->multimaterial = true; S->nf = nf;
    const size_t nv = (size_t)nf * S->g.nnodes;
    bind_nodes(S, nv);
    S->cpool.assign(nv * 7, 0.);
    double *q = S->cpool.data();
    S->C.cvol = q; q += nv;
    for (int c = 0; c < 3; c++) { S->C.cgrad[c] = q; q += nv; }
    for (int c = 0; c < 3; c++) { S->C.cdisp[c] = q; q += nv; }
    ContactParams &cp = S->cp;
    memset(&cp, 0, sizeof cp);
    cp.nf = nf; cp.normalMethod = normalMethod; cp.byDisplacements = byDisplacements; cp.positionCutoff = positionCutoff;
    for (int c = 0; c < 3; c++) cp.normal[c] = normal[c];
    auto dbleEqual = [](double a, double b) { const double d = fabs(a - b); if (d <= 1.0e-16) return true; a = fabs(a); b = fabs(b); return d <= (b > a ? b : a) * 1.0e-7; };
    cp.cubic = dbleEqual(S->g.gx, S->g.gy) && (S->dim == 2 || dbleEqual(S->g.gx, S->g.gz)) ? 1 : 0;
    for (int i = 0; i < nf * nf; i++) { cp.lawKind[i] = (i / nf == i % nf) ? LAW_IGNORE : lawKind[i]; cp.lawFriction[i] = lawFriction[i]; cp.lawStatic[i] = lawStatic[i]; }
    S->fieldOfMat.assign(fieldOfMat, fieldOfMat + S->mats.size());
    cp.rigidBias = rigidBias > 0. ? rigidBias : 1.;
    for (size_t m = 0; m < S->mats.size(); m++) if (S->mats[m].kind == MAT_RIGIDCONTACT) { cp.rigidMask |= 1 << S->fieldOfMat[m]; S->mats[m].p[8] = 0.; S->mats[m].p[9] = 0.; }
    S->rcnt.assign(nv, 0);
    S->C.rcnt = S->rcnt.data();
    S->rforce.assign(nv * 3, 0.);
    for (int c = 0; c < 3; c++) S->C.rforce[c] = S->rforce.data() + (size_t)c * nv;
    S->foffR.assign(S->PR.n ? S->PR.n : 1, 0);
    for (int p = 0; p < S->PR.n; p++) S->foffR[p] = S->fieldOfMat[S->PR.mat[p]] * S->g.nnodes;
    S->PR.foff = S->foffR.data();
    S->foff.assign(S->P.n ? S->P.n : 1, 0);
    for (int p = 0; p < S->P.n; p++) S->foff[p] = S->fieldOfMat[S->P.mat[p]] * S->g.nnodes;
    S->P.foff = S->foff.data();
    S->origpos.assign(origpos, origpos + 3 * (size_t)S->n);
}

// capi.cu::mpmgpu_set_conduction + the temperature part of mpmgpu_upload_particles
// kcond NULL: particle temperatures without conduction (a start off the stress-free temperature)
extern "C" void emu_set_conduction(void *h, const double *kcond, const double *temperature)
{
    EmuSim *S = (EmuSim *)h;
    S->conduction = kcond != NULL;
    S->thermal = true;
    const size_t nn = (size_t)S->g.nnodes;
    if (kcond) {
        S->tpool.assign(nn * 3, 0.);
        S->kcond.assign(kcond, kcond + S->mats.size());
        S->T.gT = S->tpool.data(); S->T.gVCT = S->tpool.data() + nn; S->T.gQ = S->tpool.data() + 2 * nn; S->T.kcond = S->kcond.data();
    }
    const size_t C = S->P.n ? (size_t)S->P.n : 1;
    S->temps.assign(C * 6, 0.);
    for (int p = 0; p < S->P.n; p++) S->temps[p] = temperature[p];
    S->P.temp = S->temps.data();
    for (int c = 0; c < 3; c++) S->P.tgrad[c] = S->temps.data() + (size_t)(c + 1) * C;
    S->P.dTr = S->temps.data() + 4 * C;
    S->P.dTad = S->sp.adiabatic ? S->temps.data() + 5 * C : NULL;
    // rigid particles of a material that sets the temperature (capi.cu::mpmgpu_upload_particles: rigidTemp)
    bool rigidTemp = false;
    for (size_t i = 0; i < S->mats.size() && kcond && S->PR.n > 0; i++) if (S->mats[i].kind == MAT_RIGIDBC && S->mats[i].p[10] != 0.) rigidTemp = true;
    if (rigidTemp) {
        S->rigidTemps.assign(temperature + S->P.n, temperature + S->n);
        S->ownerT.assign(nn, RIGID_NONE); S->savedT.assign(nn, 0.);
        if (S->fixedTemp.size() != nn) S->fixedTemp.assign(nn, 0);
        S->PR.temp = S->rigidTemps.data();
        S->R.ownerT = S->ownerT.data(); S->R.ptemp = S->rigidTemps.data(); S->R.savedT = S->savedT.data(); S->R.fixedT = S->fixedTemp.data();
    }
}

extern "C" void emu_set_energy_coupling(void *h, int adiabatic) { ((EmuSim *)h)->sp.adiabatic = adiabatic; }

// capi.cu::mpmgpu_set_temperature_bcs: grouped by node, list order kept inside a node
extern "C" void emu_set_temperature_bcs(void *h, int n, const int *node, const double *value)
{
    EmuSim *S = (EmuSim *)h;
    S->fixedTemp.assign((size_t)S->g.nnodes, 0);
    for (int i = 0; i < n; i++) S->fixedTemp[node[i] - 1] = 1;
    S->R.fixedT = S->fixedTemp.data();
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return node[a] < node[b]; });
    S->tbcNode.clear(); S->tbcStart.clear(); S->tbcActive.assign(n, 1); S->tbcValue.assign(n, 0.); S->tbcSaved.assign(n ? n : 1, 0.);
    for (int e = 0; e < n; e++) {
        const int i = order[e];
        if (e == 0 || node[i] != node[order[e - 1]]) { S->tbcNode.push_back(node[i] - 1); S->tbcStart.push_back(e); }
        S->tbcValue[e] = value[i];
    }
    S->tbcStart.push_back(n);
    S->Q.nUnique = (int)S->tbcNode.size(); S->Q.node = S->tbcNode.data(); S->Q.start = S->tbcStart.data(); S->Q.value = S->tbcValue.data();
    S->Q.active = S->tbcActive.data(); S->Q.saved = S->tbcSaved.data();
}

extern "C" void emu_get_transport(void *h, double *gT, double *gVCT, double *gQ, double *temperature)
{
    EmuSim *S = (EmuSim *)h;
    for (int i = 0; i < S->g.nnodes && S->conduction; i++) { gT[i] = S->T.gT[i]; gVCT[i] = S->T.gVCT[i]; gQ[i] = S->T.gQ[i]; }
    for (int p = 0; p < S->n; p++) temperature[p] = p < S->P.n ? S->P.temp[p] : (S->R.ownerT ? S->PR.temp[p - S->P.n] : S->PR.prevT[p - S->P.n]);
}

extern "C" void emu_get_contact(void *h, double *cvol, double *cgrad, double *cdisp)
{
    EmuSim *S = (EmuSim *)h;
    const size_t nn = (size_t)S->nvn;
    for (size_t i = 0; i < nn; i++) {
        cvol[i] = S->C.cvol[i];
        for (int c = 0; c < 3; c++) { cgrad[c * nn + i] = S->C.cgrad[c][i]; cdisp[c * nn + i] = S->C.cdisp[c][i]; }
    }
}

// the grouping of capi.cu::mpmgpu_set_velocity_bcs: by node, list order kept inside a node
extern "C" void emu_set_bcs(void *h, int n, const int *node, const double *norm, const double *value, const int *active, const int *symdir)
{
    EmuSim *S = (EmuSim *)h;
    S->hasBCs = n > 0;
    // dofs fixed by grid BCs, for the rigid-particle projection (capi.cu: hFixedBits)
    std::fill(S->fixedBits.begin(), S->fixedBits.end(), (unsigned char)0);
    for (int i = 0; i < n; i++) {
        unsigned char b = symdir ? (unsigned char)(symdir[i] & 7) : 0;
        for (int dd = 0; dd < 3; dd++) if (norm[3 * i + dd] != 0.) b |= (unsigned char)(1 << dd);
        S->fixedBits[node[i] - 1] |= b;
    }
    if (n == 0) return;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return node[a] < node[b]; });
    S->bcNode.clear(); S->bcStart.clear(); S->bcSym.clear();
    S->bcActive.assign(n, 1); S->bcNorm.assign(3 * (size_t)n, 0.); S->bcValue.assign(n, 0.);
    for (int e = 0; e < n; e++) {
        const int i = order[e];
        if (e == 0 || node[i] != node[order[e - 1]]) { S->bcNode.push_back(node[i] - 1); S->bcStart.push_back(e); S->bcSym.push_back(0); }
        if (symdir) S->bcSym.back() |= symdir[i];
        for (int c = 0; c < 3; c++) S->bcNorm[3 * e + c] = norm[3 * i + c];
        S->bcValue[e] = value[i]; S->bcActive[e] = active ? active[i] : 1;
    }
    S->bcStart.push_back(n);
    S->B.nUnique = (int)S->bcNode.size(); S->B.node = S->bcNode.data(); S->B.start = S->bcStart.data(); S->B.symdir = S->bcSym.data();
    S->B.active = S->bcActive.data(); S->B.norm = S->bcNorm.data(); S->B.value = S->bcValue.data();
    S->B.refl = NULL; S->B.reflRatio = NULL;
    S->bcReact.assign(3 * (size_t)n, 0.);
    S->B.reaction = S->bcReact.data();
    S->bcOrder = order;
    S->bcOfNode.assign((size_t)S->g.nnodes, -1);
    for (int u = 0; u < S->B.nUnique; u++) S->bcOfNode[S->bcNode[u]] = u;
}

// capi.cu::mpmgpu_set_velocity_bc_reflections
extern "C" void emu_set_bc_reflections(void *h, int n, const int *reflected, const double *ratio)
{
    EmuSim *S = (EmuSim *)h;
    S->bcRefl.assign(n, -1); S->bcRatio.assign(n, 1.);
    for (int e = 0; e < n; e++) { const int i = S->bcOrder[e]; S->bcRefl[e] = reflected[i] > 0 ? reflected[i] - 1 : -1; S->bcRatio[e] = ratio[i]; }
    S->B.refl = S->bcRefl.data(); S->B.reflRatio = S->bcRatio.data();
}

// capi.cu::mpmgpu_set_particle_tractions (0-based particles)
extern "C" void emu_set_tractions(void *h, int n, const int *particle, const int *face, const int *direction, const double *value, double thickness)
{
    EmuSim *S = (EmuSim *)h;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return particle[a] < particle[b]; });
    S->trStart.assign((size_t)S->P.n + 1, 0); S->trFace.resize(n); S->trDir.resize(n); S->trValue.resize(n);
    for (int e = 0; e < n; e++) { const int i = order[e]; S->trStart[particle[i] + 1]++; S->trFace[e] = face[i]; S->trDir[e] = direction[i]; S->trValue[e] = value[i]; }
    for (int i = 0; i < S->P.n; i++) S->trStart[i + 1] += S->trStart[i];
    S->TB.n = n; S->TB.start = S->trStart.data(); S->TB.face = S->trFace.data(); S->TB.dir = S->trDir.data(); S->TB.value = S->trValue.data();
    S->thickness = thickness;
}

// capi.cu::mpmgpu_set_particle_heat_fluxes
extern "C" void emu_set_heat_fluxes(void *h, int n, const int *particle, const int *face, const double *value, double thickness)
{
    EmuSim *S = (EmuSim *)h;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return particle[a] < particle[b]; });
    S->hfStart.assign((size_t)S->P.n + 1, 0); S->hfFace.resize(n); S->hfDir.assign(n, 1); S->hfValue.resize(n);
    for (int e = 0; e < n; e++) { const int i = order[e]; S->hfStart[particle[i] + 1]++; S->hfFace[e] = face[i]; S->hfValue[e] = value[i]; }
    for (int i = 0; i < S->P.n; i++) S->hfStart[i + 1] += S->hfStart[i];
    S->HF.n = n; S->HF.start = S->hfStart.data(); S->HF.face = S->hfFace.data(); S->HF.dir = S->hfDir.data(); S->HF.value = S->hfValue.data();
    S->thickness = thickness;
}

// capi.cu::mpmgpu_download_reactions
extern "C" void emu_get_reactions(void *h, double *bc, double *rigid)
{
    EmuSim *S = (EmuSim *)h;
    for (size_t e = 0; e < S->bcOrder.size() && bc; e++)
        for (int c = 0; c < 3; c++) bc[3 * (size_t)S->bcOrder[e] + c] = S->bcReact[3 * e + c];
    for (size_t i = 0; i < S->rigidReact.size() && rigid; i++) rigid[i] = S->rigidReact[i];
}

extern "C" void emu_set_xpic(void *h, int order, int usingFMPM) { EmuSim *S = (EmuSim *)h; S->sp.xpicOrder = order; S->sp.usingFMPM = usingFMPM; }
extern "C" void emu_task(void *h, int t) { run_task((EmuSim *)h, t); }
extern "C" void emu_step(void *h, int nsteps)
{
    static const int seq[11] = {0, 1, 10, 2, 3, 4, 5, 6, 7, 8, 9};
    for (int s = 0; s < nsteps; s++) for (int t = 0; t < 11; t++) run_task((EmuSim *)h, seq[t]);
}
extern "C" int emu_flags(void *h, long long *crossings, long long *leftGrid, int *nanParticle, int *cpdiLeft)
{
    EmuSim *S = (EmuSim *)h;
    *crossings = (long long)S->flags.crossings; *leftGrid = (long long)S->flags.leftGrid; *nanParticle = S->flags.nanParticle; *cpdiLeft = S->flags.cpdiLeft;
    return 0;
}

static void get_set(const Particles &P, int dim, size_t N, size_t off, double *pos, double *vel, double *sp, double *pressure, double *ep, double *wrot,
                    double *eplast, double *energies, double *history, double *acc, int *inElem, int *crossings)
{
    for (size_t p = 0; p < (size_t)P.n; p++) {
        const size_t q = off + p;
        for (int c = 0; c < 3; c++) { pos[c * N + q] = P.pos[c][p]; vel[c * N + q] = P.vel[c][p]; acc[c * N + q] = P.acc[c][p]; }
        for (int c = 0; c < 6; c++) { sp[c * N + q] = P.sp[c][p]; eplast[c * N + q] = P.eplast[c][p]; }
        pressure[q] = P.pressure[p];
        // capi.cu::k_F_to_epwrot
        const double F0 = P.F[0][p], F1 = P.F[1][p], F2 = P.F[2][p], F3 = P.F[3][p], F4 = P.F[4][p], F5 = P.F[5][p], F6 = P.F[6][p], F7 = P.F[7][p], F8 = P.F[8][p];
        ep[q] = F0 - 1.; ep[N + q] = F4 - 1.; ep[2 * N + q] = F8 - 1.; ep[5 * N + q] = F3 + F1; wrot[q] = F3 - F1;
        if (dim == 3) { ep[4 * N + q] = F6 + F2; ep[3 * N + q] = F7 + F5; wrot[N + q] = F6 - F2; wrot[2 * N + q] = F7 - F5; }
        else { ep[4 * N + q] = 0.; ep[3 * N + q] = 0.; wrot[N + q] = 0.; wrot[2 * N + q] = 0.; }
        energies[q] = P.work[p]; energies[N + q] = P.res[p]; energies[2 * N + q] = P.heat[p]; energies[3 * N + q] = P.entropy[p];
        energies[4 * N + q] = P.plast[p]; energies[5 * N + q] = P.prevT[p];
        for (int c = 0; c < MPM_MAX_HISTORY; c++) history[c * N + q] = P.hist[c][p];
        inElem[q] = P.elem[p]; crossings[q] = P.cross[p];
    }
}

extern "C" void emu_get_particles(void *h, double *pos, double *vel, double *sp, double *pressure, double *ep, double *wrot, double *eplast,
                                  double *energies, double *history, double *acc, int *inElem, int *crossings)
{
    EmuSim *S = (EmuSim *)h;
    const size_t N = (size_t)S->n;
    get_set(S->P, S->dim, N, 0, pos, vel, sp, pressure, ep, wrot, eplast, energies, history, acc, inElem, crossings);
    get_set(S->PR, S->dim, N, (size_t)S->P.n, pos, vel, sp, pressure, ep, wrot, eplast, energies, history, acc, inElem, crossings);
}

extern "C" void emu_get_nodes(void *h, int *cnt, double *mass, double *pk, double *ftot, double *vk, double *pkc)
{
    EmuSim *S = (EmuSim *)h;
    const size_t nn = (size_t)S->nvn;
    for (size_t i = 0; i < nn; i++) {
        cnt[i] = S->N.cnt[i] + (S->multimaterial ? S->C.rcnt[i] : 0); mass[i] = S->N.mass[i];
        const bool rigidField = S->multimaterial && (S->cp.rigidMask >> (i / (size_t)S->g.nnodes) & 1);
        for (int c = 0; c < 3; c++) { pk[c * nn + i] = S->N.pk[c][i]; ftot[c * nn + i] = rigidField ? S->C.rforce[c][i] : S->N.ftot[c][i]; vk[c * nn + i] = S->N.vk[c][i]; pkc[c * nn + i] = S->N.pkc[c][i]; }
    }
}

// device-layout view of the current state for the archive / global-sum functions of csrc/archive.cuh (tests only): arrays [component][n],
// non-rigid particles first, in the caller's order (the emulation never permutes)
extern "C" void emu_get_device_state(void *h, double *pos, double *vel, double *mp, double *F, double *sp, double *pressure, double *eplast,
                                     double *energies, double *hist, int *elem, int *mat0, int *cross)
{
    EmuSim *S = (EmuSim *)h;
    const size_t N = (size_t)S->n;
    const Particles *sets[2] = {&S->P, &S->PR};
    size_t off = 0;
    for (int k = 0; k < 2; k++) {
        const Particles &P = *sets[k];
        for (size_t p = 0; p < (size_t)P.n; p++) {
            const size_t q = off + p;
            for (int c = 0; c < 3; c++) { pos[c * N + q] = P.pos[c][p]; vel[c * N + q] = P.vel[c][p]; }
            mp[q] = P.mp[p];
            for (int c = 0; c < 9; c++) F[c * N + q] = P.F[c][p];
            for (int c = 0; c < 6; c++) { sp[c * N + q] = P.sp[c][p]; eplast[c * N + q] = P.eplast[c][p]; }
            pressure[q] = P.pressure[p];
            energies[q] = P.work[p]; energies[N + q] = P.res[p]; energies[2 * N + q] = P.heat[p]; energies[3 * N + q] = P.entropy[p];
            energies[4 * N + q] = P.plast[p]; energies[5 * N + q] = P.prevT[p];
            for (int c = 0; c < MPM_MAX_HISTORY; c++) hist[c * N + q] = P.hist[c][p];
            elem[q] = P.elem[p]; mat0[q] = P.mat[p]; cross[q] = P.cross[p];
        }
        off += (size_t)P.n;
    }
}

extern "C" void emu_destroy(void *h) { delete (EmuSim *)h; }
