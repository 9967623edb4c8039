// GpuTasks.cpp -- the reference-side binding of libmpmgpu: MPMTask subclasses that replace the
// reference's CPU tasks 1-9 and 11 inside its own driver (NairnMPM::CreateTasks builds the list,
// NairnMPM_Class/NairnMPM.cpp:870-1110; NairnMPM::MPMStep walks it, :284-335).
//
// This file is compiled AGAINST the reference's headers and linked WITH the reference's objects
// (nairn_mpm_fea_b200/host/build_host.sh); it contains no reference code.  Flow:
//   main()  (same steps as Common/System/main.cpp:25-140)
//     ReadFile -> StartResultsOutput -> CMPreparations            reference code, unchanged
//     GpuTasks::Install()     harvest grid, particles, materials, BC list -> mpmgpu_create/upload,
//                             then swap each eligible CPU task object for a GpuTask with the same name
//     CMAnalysis()                                                 reference code, unchanged
// Host objects (mpm[]) are refreshed from the device before every archive the reference is about to
// write (ArchiveData::ArchiveResults, System/ArchiveData.cpp:722-760) and at the end of the run.
// Errors from the C ABI are re-thrown as the reference's CommonException.
#include <omp.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "NairnMPM_Class/NairnMPM.hpp"
#include "NairnMPM_Class/MPMTask.hpp"
#include "NairnMPM_Class/MeshInfo.hpp"
#include "MPM_Classes/MPMBase.hpp"
#include "Nodes/NodalPoint.hpp"
#include "Elements/ElementBase.hpp"
#include "Materials/MaterialBase.hpp"
#include "Materials/IsotropicMat.hpp"
#include "Materials/Neohookean.hpp"
#include "Materials/Mooney.hpp"
#include "Materials/IsoPlasticity.hpp"
#include "Materials/HardeningLawBase.hpp"
#include "Materials/LinearHardening.hpp"
#include "Materials/NonlinearHardening.hpp"
#include "Materials/Nonlinear2Hardening.hpp"
#include "Materials/JohnsonCook.hpp"
#include "Materials/SCGLHardening.hpp"
#include "Global_Quantities/ThermalRamp.hpp"
#include "Materials/RigidMaterial.hpp"
#include "Boundary_Conditions/NodalVelBC.hpp"
#include "Boundary_Conditions/MatPtLoadBC.hpp"
#include "Read_XML/Expression.hpp"
#include "Boundary_Conditions/MatPtTractionBC.hpp"
#include "Boundary_Conditions/MatPtHeatFluxBC.hpp"
#include "Boundary_Conditions/NodalTempBC.hpp"
#include "Global_Quantities/BodyForce.hpp"
#include "Custom_Tasks/CustomTask.hpp"
#include "Custom_Tasks/TransportTask.hpp"
#include "Custom_Tasks/ConductionTask.hpp"
#include "NairnMPM_Class/Reservoir.hpp"
#include "System/UnitsController.hpp"
#include "Cracks/CrackHeader.hpp"
#include "System/ArchiveData.hpp"
#include "Global_Quantities/GlobalQuantity.hpp"
#include "Exceptions/CommonException.hpp"
#include "Exceptions/MPMWarnings.hpp"
#include "Materials/Elastic.hpp"
#include "Materials/ContactLaw.hpp"
#include "Materials/CoulombFriction.hpp"
#undef private
#undef protected

#include "../../include/mpmgpu.h"

extern MPMTask *firstMPMTask;
extern double timestep, strainTimestepFirst, strainTimestepLast, fractionUSF, mtime;

namespace {

mpmgpu_ctx *gCtx = NULL;
// -gpus N: one context per GPU, each holding the particles of its z-slab of cell planes (mpmgpu_slab_*; the library does the halo
// and migrant exchanges itself over NCCL).  gCtx is slab 0 then; one OpenMP thread per GPU drives mpmgpu_slab_step.
std::vector<mpmgpu_ctx *> gSlabs;
int gNumGpus = 1;
bool gHostStale = false;            // device is ahead of mpm[]
std::vector<NodalVelBC *> gBCs;     // host BC list in list order
// "reactionx/y/z" global quantities (GlobalQuantity.cpp:971-986) read NodalVelBC::freaction of the host's BC objects: when one is
// asked for, the device keeps every BC's reaction force (mpmgpu_track_reactions) and SyncReactionsToHost writes them into the
// objects before the reference's own code sums them by ID.  The BCs rigid particles would make (ProjectRigidBCsTask, not run on
// the host any more) are stood in for by one carrier object per rigid-BC material at the end of the list, holding the material's
// summed reaction under its ID (= material number, ProjectRigidBCsTask.cpp:241).
bool gTrackReactions = false;
std::vector<NodalVelBC *> gRigidCarriers;       // by material index, NULL for the others
bool gBCsVary = false;
bool gRigidFunctions = false;       // some rigid-BC material sets its velocity by functions of time and position
bool gRigidValueFunctions = false;  // ... or its temperature by a value function (RigidMaterial::GetValueSetting)
long long gLeftGridWarned = 0;     // first-time grid leavers already handed to the reference's MPMWarnings
bool gCustomTasksReadParticles = false;     // (the only custom task the adapter admits, PeriodicXPIC, touches bodyFrc alone)
std::vector<MatPtTractionBC *> gTractions;      // particle traction BCs in list order
bool gTractionsVary = false;
std::vector<MatPtHeatFluxBC *> gHeatFluxes;     // particle heat-flux BCs in list order (conduction)
bool gHeatFluxesVary = false;
std::vector<int> gLoadPts;          // particles with load BCs (MatPtLoadBC), 0-based, each once
bool gLoadsSent = false;
std::vector<NodalTempBC *> gTempBCs;    // nodal temperature BCs in list order (conduction)
bool gTempBCsVary = false, gTempBCsSent = false;
bool gThermal = false;              // particle temperatures live on the device (conduction, or a start off the stress-free temperature)
bool gFusedStep = false;            // -fused: the whole step runs in the first task (mpmgpu_step, fused kernels); the other tasks are empty

void check(int rc, const char *where, mpmgpu_ctx *ctx = NULL)
{
    if (rc == MPMGPU_OK) return;
    throw CommonException(mpmgpu_last_error(ctx ? ctx : gCtx), where);
}

// the same call on every slab context (one context without -gpus)
template <class F>
void each_ctx(F f, const char *where)
{
    if (gSlabs.empty()) { check(f(gCtx), where); return; }
    for (mpmgpu_ctx *c : gSlabs) check(f(c), where, c);
}

void AssignParticle(MPMBase *m, int n, int q, const double *pos, const double *vel, const double *acc, const double *sp, const double *pr, const double *ep,
                    const double *wrot, const double *epl, const double *en, const double *hist, const int *elem, const int *cross);

// -gpus N: every slab hands back ITS particles in device order with their ids (the host's particle numbers); rigid-BC particles
// live on every slab and move identically, slab 0's copy is taken
void DownloadSlabsToHost(void)
{
    for (size_t r = 0; r < gSlabs.size(); r++) {
        mpmgpu_ctx *c = gSlabs[r];
        const int n = mpmgpu_num_particles(c);
        if (n == 0) continue;
        std::vector<double> pos(3 * (size_t)n), vel(3 * (size_t)n), sp(6 * (size_t)n), pr(n), ep(6 * (size_t)n), wrot(3 * (size_t)n), epl(6 * (size_t)n), en(6 * (size_t)n),
            hist((size_t)MPMGPU_MAX_HISTORY * n), acc(3 * (size_t)n);
        std::vector<int> elem(n), cross(n), ids(n);
        mpmgpu_particles h;
        memset(&h, 0, sizeof h);
        h.pos = pos.data(); h.vel = vel.data(); h.sp = sp.data(); h.pressure = pr.data(); h.ep = ep.data(); h.wrot = wrot.data();
        h.eplast = epl.data(); h.energies = en.data(); h.history = hist.data(); h.acc = acc.data(); h.in_elem = elem.data(); h.crossings = cross.data();
        h.ids = ids.data();
        check(mpmgpu_download_particles(c, &h, MPMGPU_F_ALL), "GpuTasks::DownloadSlabsToHost", c);
        for (int q = 0; q < n; q++) {
            const int p = ids[q];
            if (p < 0 || p >= nmpms) throw CommonException("a slab returned a particle id outside the host's list", "GpuTasks::DownloadSlabsToHost");
            if (p >= nmpmsNR && r != 0) continue;
            AssignParticle(mpm[p], n, q, pos.data(), vel.data(), acc.data(), sp.data(), pr.data(), ep.data(), wrot.data(), epl.data(), en.data(), hist.data(),
                           elem.data(), cross.data());
        }
    }
}

// device -> mpm[] (MPMBase fields; SetDeformationGradient is implicit: ep + wrot are downloaded)
// "Grid Kinetic Energy" (GlobalQuantity.cpp:1145-1151) sums 0.5 |pk|^2 / mass over the host's nodes: in single-material mode the
// mass, momentum and point count of every node are copied into the host's MatVelocityField objects before a global archive
bool gGridKineticEnergy = false;
void SyncNodesToHost(void)
{
    if (!gGridKineticEnergy) return;
    const size_t nn = (size_t)nnodes;
    std::vector<int> cnt(nn); std::vector<double> mass(nn), pk(3 * nn);
    mpmgpu_nodes h;
    memset(&h, 0, sizeof h);
    h.nnodes = nnodes; h.number_points = cnt.data(); h.mass = mass.data(); h.pk = pk.data();
    check(mpmgpu_download_nodes(gCtx, &h), "GpuTasks::SyncNodesToHost");
    for (size_t i = 0; i < nn; i++) {
        MatVelocityField *m = nd[i + 1]->cvf[0]->mvf[0];
        nd[i + 1]->cvf[0]->numberPoints = cnt[i];         // (CrackVelocityField::ActiveField reads the crack field's own count)
        m->numberPoints = cnt[i]; m->mass = mass[i]; m->pk = MakeVector(pk[i], pk[nn + i], pk[2 * nn + i]);
    }
}

void SyncReactionsToHost(void)
{
    SyncNodesToHost();
    if (!gTrackReactions) return;
    std::vector<double> bc(3 * gBCs.size() + 3), rigid(3 * (size_t)nmat);
    check(mpmgpu_download_reactions(gCtx, (int)gBCs.size(), bc.data(), rigid.data()), "GpuTasks::SyncReactionsToHost");
    for (size_t i = 0; i < gBCs.size(); i++) gBCs[i]->freaction = MakeVector(bc[3 * i], bc[3 * i + 1], bc[3 * i + 2]);
    for (int m = 0; m < nmat; m++)
        if (gRigidCarriers[m] != NULL) gRigidCarriers[m]->freaction = MakeVector(rigid[3 * m], rigid[3 * m + 1], rigid[3 * m + 2]);
}

void DownloadToHost(void)
{
    SyncReactionsToHost();
    if (!gHostStale) return;
    if (!gSlabs.empty()) { DownloadSlabsToHost(); gHostStale = false; return; }
    const int n = nmpms;
    std::vector<double> pos(3 * n), vel(3 * n), sp(6 * n), pr(n), ep(6 * n), wrot(3 * n), epl(6 * n), en(6 * n), hist(MPMGPU_MAX_HISTORY * n), acc(3 * n);
    std::vector<int> elem(n), cross(n);
    mpmgpu_particles h;
    memset(&h, 0, sizeof h);
    h.pos = pos.data(); h.vel = vel.data(); h.sp = sp.data(); h.pressure = pr.data(); h.ep = ep.data(); h.wrot = wrot.data();
    h.eplast = epl.data(); h.energies = en.data(); h.history = hist.data(); h.acc = acc.data(); h.in_elem = elem.data(); h.crossings = cross.data();
    std::vector<double> temp;
    if (gThermal) { temp.resize(n); h.temperature = temp.data(); }
    check(mpmgpu_download_particles(gCtx, &h, MPMGPU_F_ALL | (gThermal ? MPMGPU_F_TEMPERATURE : 0)), "GpuTasks::DownloadToHost");
    for (int p = 0; p < n; p++) {
        MPMBase *m = mpm[p];
        AssignParticle(m, n, p, pos.data(), vel.data(), acc.data(), sp.data(), pr.data(), ep.data(), wrot.data(), epl.data(), en.data(), hist.data(), elem.data(), cross.data());
        if (gThermal) { m->pTemperature = temp[p]; m->pPreviousTemperature = en[5 * n + p]; }
    }
    gHostStale = false;
}

// column q of component-major arrays of length n -> one host particle
void AssignParticle(MPMBase *m, int n, int q, const double *pos, const double *vel, const double *acc, const double *sp, const double *pr, const double *ep,
                    const double *wrot, const double *epl, const double *en, const double *hist, const int *elem, const int *cross)
{
    const size_t N = (size_t)n, p = (size_t)q;
    m->pos = MakeVector(pos[p], pos[N + p], pos[2 * N + p]);
    m->vel = MakeVector(vel[p], vel[N + p], vel[2 * N + p]);
    m->acc = MakeVector(acc[p], acc[N + p], acc[2 * N + p]);
    m->sp.xx = sp[p]; m->sp.yy = sp[N + p]; m->sp.zz = sp[2 * N + p]; m->sp.yz = sp[3 * N + p]; m->sp.xz = sp[4 * N + p]; m->sp.xy = sp[5 * N + p];
    m->pressure = pr[p];
    m->ep.xx = ep[p]; m->ep.yy = ep[N + p]; m->ep.zz = ep[2 * N + p]; m->ep.yz = ep[3 * N + p]; m->ep.xz = ep[4 * N + p]; m->ep.xy = ep[5 * N + p];
    m->wrot.xy = wrot[p]; m->wrot.xz = wrot[N + p]; m->wrot.yz = wrot[2 * N + p];
    m->eplast.xx = epl[p]; m->eplast.yy = epl[N + p]; m->eplast.zz = epl[2 * N + p]; m->eplast.yz = epl[3 * N + p]; m->eplast.xz = epl[4 * N + p]; m->eplast.xy = epl[5 * N + p];
    m->workEnergy = en[p]; m->resEnergy = en[N + p]; m->heatEnergy = en[2 * N + p]; m->entropy = en[3 * N + p]; m->plastEnergy = en[4 * N + p];
    if (elem[p] != m->inElem) { m->prevInElem = m->inElem; m->inElem = elem[p]; }
    m->elementCrossings = cross[p];
    const int nh = theMaterials[m->MatID()]->NumberOfHistoryDoubles();
    for (int k = 0; k < nh && k < MPMGPU_MAX_HISTORY && m->matData != NULL; k++) ((double *)m->matData)[k] = hist[(size_t)k * N + p];
}

// ---- output from the device --------------------------------------------------------------
// The reference writes a particle archive from mpm[] (ArchiveData::ArchiveResults, System/ArchiveData.cpp:720-1290) and the
// global quantities by looping over mpm[] once per quantity (GlobalQuantity::AppendQuantity, GlobalQuantity.cpp:394-1075).
// Neither is virtual, so this adapter runs them FIRST -- at the end of the step's last task, before NairnMPM::CMAnalysis calls
// ArchiveResults(mtime+timestep) -- from device data: the record block comes packed from mpmgpu_pack_archive, the sums from
// mpmgpu_global_sums; file names, header, the line printed to the output file and the advance of nextArchTime /
// nextGlobalTime repeat the reference's statements, so its own call then finds nothing due.  The last step of a run (which the
// reference archives unconditionally) and anything the device does not produce (an archive item or a global quantity outside
// the lists in include/mpmgpu.h, a tracer particle) take the old route: full download, reference code.
bool gDeviceOutput = true;          // -hostoutput switches it off
std::vector<char> gArchiveBlock;
long long gArchiveBytes = 0, gArchiveCount = 0;
double gArchiveSeconds = 0.;

// value of a global quantity from the per-material device sums; false = not covered
bool QuantityFromSums(GlobalQuantity *g, const std::vector<double> &sums, double &value)
{
    const int NS = MPMGPU_GS_NSUMS;
    if (g->ptNum >= 0 || g->ptPos != NULL) return false;
    double t[MPMGPU_GS_NSUMS];
    for (int k = 0; k < NS; k++) t[k] = 0.;
    for (int m = 0; m < nmat; m++) {
        if (!g->IncludeThisMaterial(m, false)) continue;
        if (theMaterials[m]->IsRigid()) return false;           // the sums run over the non-rigid particles
        for (int k = 0; k < NS; k++) { const double v = sums[(size_t)m * NS + k]; if (v == v) t[k] += v; }
    }
    const double vol = t[MPMGPU_GS_VOLUME];
    int c = -1;
    switch (g->quantity) {
    case AVG_SXX: c = 0; break; case AVG_SYY: c = 1; break; case AVG_SZZ: c = 2; break;
    case AVG_SYZ: c = 3; break; case AVG_SXZ: c = 4; break; case AVG_SXY: c = 5; break;
    default: break;
    }
    if (c >= 0) { value = t[MPMGPU_GS_STRESS + c]; if (vol > 0.) value /= vol; value *= UnitsController::Scaling(1.e-6); return true; }
    switch (g->quantity) {
    case KINE_ENERGY: value = t[MPMGPU_GS_KINETIC] * UnitsController::Scaling(1.e-9); return true;
    case WORK_ENERGY: value = t[MPMGPU_GS_WORK] * UnitsController::Scaling(1.e-9); return true;
    case STRAIN_ENERGY: value = t[MPMGPU_GS_STRAIN_ENERGY] * UnitsController::Scaling(1.e-9); return true;
    case HEAT_ENERGY: value = t[MPMGPU_GS_HEAT] * UnitsController::Scaling(1.e-9); return true;
    case ENTROPY_ENERGY: value = t[MPMGPU_GS_ENTROPY] * UnitsController::Scaling(1.e-9); return true;
    case INTERNAL_ENERGY: value = (t[MPMGPU_GS_WORK] + t[MPMGPU_GS_HEAT]) * UnitsController::Scaling(1.e-9); return true;
    case PLAS_ENERGY: value = t[MPMGPU_GS_PLASTIC] * UnitsController::Scaling(1.e-9); return true;
    case AVG_VELX: value = t[MPMGPU_GS_VOL_VEL]; if (vol > 0.) value /= vol; return true;
    case AVG_VELY: value = t[MPMGPU_GS_VOL_VEL + 1]; if (vol > 0.) value /= vol; return true;
    case AVG_VELZ: value = t[MPMGPU_GS_VOL_VEL + 2]; if (vol > 0.) value /= vol; return true;
    case LINMOMX: value = t[MPMGPU_GS_LINMOM] * UnitsController::Scaling(1.e-6); return true;
    case LINMOMY: value = t[MPMGPU_GS_LINMOM + 1] * UnitsController::Scaling(1.e-6); return true;
    case LINMOMZ: value = t[MPMGPU_GS_LINMOM + 2] * UnitsController::Scaling(1.e-6); return true;
    default: break;
    }
    static const int fq[9] = {AVG_FXX, AVG_FXY, AVG_FXZ, AVG_FYX, AVG_FYY, AVG_FYZ, AVG_FZX, AVG_FZY, AVG_FZZ};
    for (int i = 0; i < 9; i++)
        if (g->quantity == fq[i]) { value = t[MPMGPU_GS_VOL_F + i]; if (vol > 0.) value /= vol; value *= UnitsController::Scaling(100.); return true; }
    return false;
}

// "contactx/y/z": GlobalQuantity.cpp:905-968 with the node loop (NodalPoint::AddGetContactForce) replaced by the device's sum over
// the rigid material fields; the bookkeeping of the steps since the forces were last cleared is the archiver's own
bool gContactQuantities = false;
bool QuantityFromContact(GlobalQuantity *g, double &value)
{
    const int q = g->quantity;
    if (q != TOT_FCONX && q != TOT_FCONY && q != TOT_FCONZ) return false;
    if (!gContactQuantities) return false;
    Vector ftotal = MakeVector(0., 0., 0.);
    if (fmobj->mstep != 0) {
        const int totalSteps = archiver->GetArchiveContactStepInterval();
        Vector *forces = archiver->GetLastContactForcePtr();
        if (totalSteps > 0) {
            std::vector<double> f(3 * (size_t)maxMaterialFields, 0.);
            check(mpmgpu_contact_forces(gCtx, 1, f.data()), "GpuTasks::contact forces");
            // NodalPoint::AddGetContactForce: scale = -Scaling(1.e-6) * stepScale / timestep with stepScale = -1/totalSteps
            const double scale = UnitsController::Scaling(1.e-6) / ((double)totalSteps * timestep);
            for (int im = 0; im < maxMaterialFields; im++) forces[im] = MakeVector(f[3 * im] * scale, f[3 * im + 1] * scale, f[3 * im + 2] * scale);
        }
        for (int im = 0; im < maxMaterialFields; im++) {
            if (g->whichMat == 0) AddVector(&ftotal, &forces[im]);
            else if (g->whichMat == MaterialBase::GetFieldMatID(im) + 1) { AddVector(&ftotal, &forces[im]); break; }
        }
    }
    value = q == TOT_FCONX ? ftotal.x : (q == TOT_FCONY ? ftotal.y : ftotal.z);
    return true;
}

// quantities that do not read the particles (step number, times, grid damping values ...): the reference's own code serves
bool QuantityIsParticleFree(int q)
{
    // (reaction forces: the reference's code reads the BC objects SyncReactionsToHost has just filled)
    if (gTrackReactions && (q == TOT_REACTX || q == TOT_REACTY || q == TOT_REACTZ)) return true;
    if (gGridKineticEnergy && q == GRID_KINE_ENERGY) return true;       // (the host's nodes hold the device's values: SyncNodesToHost)
    return q == STEP_NUMBER || q == CPU_TIME || q == ELAPSED_TIME || q == GRID_ALPHA || q == PARTICLE_ALPHA;
}

bool GlobalsCoveredByDevice(void)
{
    std::vector<double> zero((size_t)nmat * MPMGPU_GS_NSUMS, 0.);
    for (GlobalQuantity *g = firstGlobal; g != NULL; g = g->GetNextGlobal()) {
        double v;
        if (gContactQuantities && (g->quantity == TOT_FCONX || g->quantity == TOT_FCONY || g->quantity == TOT_FCONZ)) continue;
        if (!QuantityIsParticleFree(g->quantity) && !QuantityFromSums(g, zero, v)) return false;
    }
    return true;
}

// ArchiveData::GlobalArchive (ArchiveData.cpp:1508-1553) with the particle loops replaced by the device sums
void GlobalArchiveFromDevice(double atime)
{
    if (archiver->globalFile == NULL) return;
    std::vector<double> sums((size_t)nmat * MPMGPU_GS_NSUMS, 0.);
    check(mpmgpu_global_sums(gCtx, sums.data()), "GpuTasks::GlobalArchive");
    SyncReactionsToHost();
    archiver->lastArchived.clear();
    archiver->lastArchivedStep = fmobj->mstep;
    for (GlobalQuantity *g = firstGlobal; g != NULL;) {
        double v;
        if (QuantityFromContact(g, v)) { archiver->lastArchived.push_back(v); g = g->GetNextGlobal(); }
        else if (QuantityIsParticleFree(g->quantity)) g = g->AppendQuantity(archiver->lastArchived);
        else { QuantityFromSums(g, sums, v); archiver->lastArchived.push_back(v); g = g->GetNextGlobal(); }
    }
    char fline[1000], numStr[100];
    snprintf(fline, sizeof fline, "%g", UnitsController::Scaling(1000.) * atime);
    for (size_t i = 0; i < archiver->lastArchived.size(); i++) { snprintf(numStr, sizeof numStr, "\t%e", archiver->lastArchived[i]); strcat(fline, numStr); }
    std::ofstream global;
    global.open(archiver->globalFile, std::ios::out | std::ios::app);
    if (!global.is_open()) throw CommonException("File error opening global results", "GpuTasks::GlobalArchive");
    global << fline << std::endl;
    global.close();
}

// the part of ArchiveData::ArchiveResults (ArchiveData.cpp:720-1290) that is due after this step, from device data.
// Returns false when the reference has to do it itself from a downloaded mpm[].
bool OutputFromDevice(double atime)
{
    if (gDeviceOutput && gContactQuantities && atime > fmobj->maxtime) {
        // last step: the reference writes the global row itself (ArchiveData.cpp:733), and its contact-force quantities would sum host
        // nodes nobody filled.  Read the forces from the device now: that leaves them in the archiver's own array and moves its
        // step mark, so the reference's code finds zero steps since the last reading and uses what is stored (GlobalQuantity.cpp:931-958)
        for (GlobalQuantity *g = firstGlobal; g != NULL; g = g->GetNextGlobal()) {
            double v;
            if (QuantityFromContact(g, v)) break;
        }
        return false;
    }
    if (!gDeviceOutput || atime > fmobj->maxtime) return false;         // last step: archived unconditionally by the reference
    const bool globalByTime = firstGlobal != NULL && archiver->globalTime >= 0.;
    const bool globalDue = globalByTime && atime > archiver->nextGlobalTime;
    const bool archiveDue = atime >= archiver->nextArchTime;
    if (!globalDue && !archiveDue) return true;
    if (firstGlobal != NULL && !GlobalsCoveredByDevice()) return false;
    char order[60];
    strncpy(order, archiver->mpmOrder, sizeof order - 1); order[sizeof order - 1] = 0;
    if (archiveDue) {
        const int rec = mpmgpu_archive_record_size(gCtx, order);
        if (rec < 0 || rec != archiver->mpmRecSize || archiver->recSize < rec || fmobj->GetReverseBytes()) return false;
    }
    if (globalDue) { GlobalArchiveFromDevice(atime); archiver->nextGlobalTime += archiver->globalTime; }
    if (!archiveDue) return true;
    // next archive time (ArchiveData.cpp:748-756)
    archiver->nextArchTime += archiver->archTimes[archiver->archBlock];
    if (archiver->archBlock + 1 < (int)archiver->firstArchTimes.size()) {
        if (archiver->nextArchTime > archiver->firstArchTimes[archiver->archBlock + 1]) {
            archiver->archBlock++;
            archiver->nextArchTime = atime + archiver->archTimes[archiver->archBlock];
        }
    }
    if (firstGlobal != NULL && archiver->globalTime < 0.) GlobalArchiveFromDevice(atime);
    if (mpmReservoir != NULL) archiver->ArchiveResizings(atime * UnitsController::Scaling(1.e3), fmobj->mstep);     // (reads no particles)
    const double t0 = fmobj->ElapsedTime();
    char fname[500], fline[600];
    archiver->GetFilePathNum(fname, sizeof fname, "%s%s.%d", fmobj->mstep);
    int i;
    for (i = (int)strlen(fname); i >= 0; i--) if (fname[i] == '/' || fname[i] == '\\') break;
    snprintf(fline, sizeof fline, "%7d %15.7e  %s", fmobj->mstep, atime * UnitsController::Scaling(1.e3), &fname[i + 1]);
    std::cout << fline << std::endl;
    const size_t rec = (size_t)archiver->mpmRecSize, bytes = rec * (size_t)nmpms;
    if (gArchiveBlock.size() < bytes) gArchiveBlock.resize(bytes);
    check(mpmgpu_pack_archive(gCtx, order, gArchiveBlock.data(), gArchiveBlock.size()), "GpuTasks::ArchiveResults");
    std::ofstream afile;
    afile.open(fname, std::ios::out | std::ios::binary);
    if (!afile.is_open()) throw CommonException("Cannot open an archive file", "GpuTasks::ArchiveResults");
    *archiver->timeStamp = (float)(atime * UnitsController::Scaling(1.e3));
    afile.write(archiver->archHeader, HEADER_LENGTH);
    if ((size_t)archiver->recSize == rec) afile.write(gArchiveBlock.data(), (std::streamsize)bytes);
    else {      // records padded to the size of the longer crack-segment record (ArchiveData.cpp:1274-1276)
        std::vector<char> padded((size_t)archiver->recSize * (size_t)nmpms, 0);
        for (int p = 0; p < nmpms; p++) memcpy(&padded[(size_t)p * archiver->recSize], &gArchiveBlock[(size_t)p * rec], rec);
        afile.write(padded.data(), (std::streamsize)padded.size());
    }
    if (afile.bad()) throw CommonException("File error writing archive file", "GpuTasks::ArchiveResults");
    afile.close();
    gArchiveBytes += (long long)bytes; gArchiveCount++; gArchiveSeconds += fmobj->ElapsedTime() - t0;
    return true;
}

enum { G_INIT, G_MASSMOM, G_POSTEXTRAP, G_USF, G_FORCES, G_POSTFORCES, G_MOMENTA, G_PARTICLES, G_USL, G_RESET, G_RIGIDBC };

// One class for all ten tasks: Execute() forwards to the matching C entry point.
class GpuTask : public MPMTask
{
  public:
    int which;
    GpuTask(const char *name, int w) : MPMTask(name), which(w) {}
    void UpdateBCValues(void)
    {   // NodalVelBC::GridVelocityBCValues (NodalVelBC.cpp:293-302): values at this step's mtime
        if (!gBCsVary) return;
        std::vector<double> v(gBCs.size()); std::vector<int> a(gBCs.size());
        for (size_t i = 0; i < gBCs.size(); i++) { a[i] = gBCs[i]->GetNodeNum(mtime) > 0; v[i] = a[i] ? gBCs[i]->BCValue(mtime) : 0.; }
        each_ctx([&](mpmgpu_ctx *c) { return mpmgpu_update_velocity_bc_values(c, (int)v.size(), v.data(), a.data()); }, "GpuTask(BC values)");
    }
    // nodal temperature BCs at this step's time (TransportTask::ImposeValueBCs / ImposeValueGridBCs read BCValue(mtime))
    void UpdateTemperatureBCs(void)
    {
        if (gTempBCs.empty() || (gTempBCsSent && !gTempBCsVary)) return;
        std::vector<int> node(gTempBCs.size()), act(gTempBCs.size()); std::vector<double> val(gTempBCs.size());
        for (size_t i = 0; i < gTempBCs.size(); i++) {
            node[i] = gTempBCs[i]->GetNodeNum();
            act[i] = gTempBCs[i]->GetNodeNum(mtime) != 0 ? 1 : 0;
            val[i] = act[i] ? gTempBCs[i]->BCValue(mtime) : 0.;
        }
        check(mpmgpu_set_temperature_bcs(gCtx, (int)node.size(), node.data(), val.data(), act.data()), "GpuTask(temperature BCs)");
        gTempBCsSent = true;
    }
    // MatPtLoadBC::SetParticleFext (InitializationTask.cpp:91) by the reference's own BC objects on mpm[]->pFext; the forces of the
    // loaded particles go to the device
    void UpdateParticleLoads(void)
    {
        if (gLoadPts.empty()) return;
        MatPtLoadBC::SetParticleFext(mtime);
        const size_t nl = gLoadPts.size();
        std::vector<double> f(3 * nl);
        for (size_t k = 0; k < nl; k++) { const Vector *pf = mpm[gLoadPts[k]]->GetPFext(); f[k] = pf->x; f[nl + k] = pf->y; f[2 * nl + k] = pf->z; }
        check(mpmgpu_update_particle_loads(gCtx, (int)nl, gLoadsSent ? NULL : gLoadPts.data(), f.data()), "GpuTask(particle loads)");
        gLoadsSent = true;
    }
    // MatPtTractionBC values at this step's time (MatPtTractionBC::AddMPFluxBC reads BCValue(bctime), MatPtTractionBC.cpp:68)
    void UpdateTractionValues(void)
    {
        if (gTractions.empty() || !gTractionsVary) return;
        std::vector<double> v(gTractions.size());
        for (size_t i = 0; i < gTractions.size(); i++) v[i] = gTractions[i]->BCValue(mtime);
        check(mpmgpu_update_particle_traction_values(gCtx, (int)v.size(), v.data()), "GpuTask(traction values)");
    }
    void UpdateHeatFluxValues(void)
    {
        if (gHeatFluxes.empty() || !gHeatFluxesVary) return;
        std::vector<double> v(gHeatFluxes.size());
        for (size_t i = 0; i < gHeatFluxes.size(); i++) v[i] = gHeatFluxes[i]->BCValue(mtime);
        check(mpmgpu_update_particle_heat_flux_values(gCtx, (int)v.size(), v.data()), "GpuTask(heat flux values)");
    }
    // RigidMaterial::GetVectorSetting evaluated by the reference's own Expression objects (ProjectRigidBCsTask.cpp:75-93)
    void UpdateRigidVelocities(void)
    {
        if (!gRigidFunctions) return;
        // (rigid contact particles [nmpmsRB, nmpmsRC): SetRigidContactVelTask.cpp:38-46; rigid-BC particles after them)
        const int nr = nmpms - nmpmsNR;
        std::vector<double> v(3 * (size_t)nr);
        for (int p = nmpmsNR; p < nmpms; p++) {
            MPMBase *m = mpm[p];
            bool hasDir[3];
            ((RigidMaterial *)theMaterials[m->MatID()])->GetVectorSetting(&m->vel, hasDir, mtime, &m->pos);
            const int j = p - nmpmsNR;
            v[j] = m->vel.x; v[nr + j] = m->vel.y; v[2 * (size_t)nr + j] = m->vel.z;
        }
        each_ctx([&](mpmgpu_ctx *c) { return mpmgpu_update_rigid_velocities(c, nr, v.data()); }, "GpuTask(rigid velocities)");
        if (gRigidValueFunctions) {     // ProjectRigidBCsTask.cpp:118-125: the value function sets the particle's temperature, which its BCs impose
            std::vector<double> t((size_t)nr);
            for (int p = nmpmsNR; p < nmpms; p++) {
                MPMBase *m = mpm[p];
                RigidMaterial *rm = (RigidMaterial *)theMaterials[m->MatID()];
                double rvalue;
                if (rm->RigidTemperature() && rm->GetValueSetting(&rvalue, mtime, &m->pos)) m->pTemperature = rvalue;
                t[p - nmpmsNR] = m->pTemperature;
            }
            check(mpmgpu_update_rigid_temperatures(gCtx, nr, t.data()), "GpuTask(rigid temperatures)");
        }
    }
    void AfterStep(void)
    {
        if (gRigidFunctions)        // keep the host copy of the rigid positions current for the next evaluation
            for (int p = nmpmsNR; p < nmpms; p++) mpm[p]->MovePosition(timestep);
        gHostStale = true;
        {   // particles pushed back into the grid: the reference warns once per particle and aborts at the <LeaveLimit>
            // threshold (ResetElementsTask.cpp:71-95); same warning object, same exception
            long long exits = 0, first = 0;
            each_ctx([&](mpmgpu_ctx *c) { long long e = 0, f = 0; const int rc = mpmgpu_left_grid_counts(c, &e, &f); exits += e; first += f; return rc; }, "GpuTask(ResetElements)");
            for (; gLeftGridWarned < first; gLeftGridWarned++)
                if (warnings.Issue(fmobj->warnParticleLeftGrid, -1) == REACHED_MAX_WARNINGS) {
                    DownloadToHost();
                    throw CommonException("Too many particles have left the grid\n  (plot x displacement to see last one).", "ResetElementsTask::Execute");
                }
        }
        // will the reference archive after this step?  (ArchiveResults(mtime+timestep,...), ArchiveData.cpp:731-746)
        const double atime = mtime + timestep;
        bool due = atime >= archiver->nextArchTime || atime + timestep > fmobj->maxtime;
        if (firstGlobal != NULL && archiver->globalTime >= 0. && atime > archiver->nextGlobalTime) due = true;
        if (due && OutputFromDevice(atime)) due = atime + timestep > fmobj->maxtime;       // written from device data: nothing left for the host
        if (due || (theTasks != NULL && gCustomTasksReadParticles)) DownloadToHost();
    }
    virtual bool Execute(int)
    {
        if (gFusedStep) {       // one call per step; the per-task rows of the timing report then show the whole step under "Initialize"
            if (which == G_INIT) {
                each_ctx([&](mpmgpu_ctx *c) { return mpmgpu_set_xpic(c, bodyFrc.GetXPICOrder(), bodyFrc.UsingFMPM() ? 1 : 0); }, "GpuTask(step)");
                UpdateBCValues();
                UpdateRigidVelocities();
                UpdateParticleLoads();
                UpdateTractionValues();
                UpdateHeatFluxValues();
                UpdateTemperatureBCs();
                if (!gSlabs.empty()) {
                    // every slab steps at the same time: the halo and migrant exchanges inside mpmgpu_slab_step are NCCL calls that
                    // wait for the neighbours
                    std::vector<int> rcs(gSlabs.size(), MPMGPU_OK);
#pragma omp parallel num_threads((int)gSlabs.size())
                    {
                        const int r = omp_get_thread_num();
                        if (r < (int)gSlabs.size()) rcs[r] = mpmgpu_slab_step(gSlabs[r], 1);
                    }
                    for (size_t r = 0; r < gSlabs.size(); r++) check(rcs[r], "GpuTask(slab step)", gSlabs[r]);
                } else
                check(mpmgpu_step(gCtx, 1), "GpuTask(step)");
            } else if (which == G_RESET) AfterStep();
            return true;
        }
        switch (which) {
        case G_INIT:
            // the PeriodicXPIC custom task changes the order between steps (Custom_Tasks/PeriodicXPIC.cpp:161-240)
            check(mpmgpu_set_xpic(gCtx, bodyFrc.GetXPICOrder(), bodyFrc.UsingFMPM() ? 1 : 0), "GpuTask(Initialize)");
            UpdateParticleLoads();
            UpdateTractionValues();
            UpdateHeatFluxValues();
            UpdateTemperatureBCs();
            if (nmpmsRC != nmpmsRB) UpdateRigidVelocities();        // rigid contact particles: SetRigidContactVelTask runs before the extrapolation
            check(mpmgpu_task_initialization(gCtx), "GpuTask(Initialize)");
            break;
        case G_RIGIDBC:
            UpdateRigidVelocities();
            check(mpmgpu_task_project_rigid_bcs(gCtx), "GpuTask(ProjectRigidBCs)");
            break;
        case G_MASSMOM: check(mpmgpu_task_mass_and_momentum(gCtx), "GpuTask(MassAndMomentum)"); break;
        case G_POSTEXTRAP:
            UpdateBCValues();
            check(mpmgpu_task_post_extrapolation(gCtx), "GpuTask(PostExtrapolation)");
            break;
        case G_USF: check(mpmgpu_task_update_strains_first(gCtx), "GpuTask(UpdateStrainsFirst)"); break;
        case G_FORCES: check(mpmgpu_task_grid_forces(gCtx), "GpuTask(GridForces)"); break;
        case G_POSTFORCES: check(mpmgpu_task_post_forces(gCtx), "GpuTask(PostForces)"); break;
        case G_MOMENTA: check(mpmgpu_task_update_momenta(gCtx), "GpuTask(UpdateMomenta)"); break;
        case G_PARTICLES: check(mpmgpu_task_update_particles(gCtx), "GpuTask(UpdateParticles)"); break;
        case G_USL: check(mpmgpu_task_update_strains_last(gCtx), "GpuTask(UpdateStrainsLast)"); break;
        case G_RESET:
            check(mpmgpu_task_reset_elements(gCtx), "GpuTask(ResetElements)");
            AfterStep();
            break;
        }
        return true;
    }
};

int TaskCode(const char *name)
{
    static const struct { const char *nm; int code; } map[] = {
        {"Initialize", G_INIT}, {"Extrapolate Mass and Momentum", G_MASSMOM}, {"Post Extrapolation Tasks", G_POSTEXTRAP},
        {"Rigid BCs by Projection", G_RIGIDBC},
        {"Update Strains First", G_USF}, {"Extrapolate Grid Forces", G_FORCES}, {"Post Force Extrapolation Tasks", G_POSTFORCES},
        {"Update Momenta", G_MOMENTA}, {"Update Particles", G_PARTICLES}, {"Update Strains Last with Extrapolation", G_USL},
        {"Update Strains Last", G_USL}, {"Reset Elements", G_RESET}};
    for (size_t i = 0; i < sizeof map / sizeof map[0]; i++) if (strcmp(name, map[i].nm) == 0) return map[i].code;
    return -1;
}

} // namespace

// Returns NULL when installed, else the reason the run stays on the CPU tasks.
// a particle BC's function may read the particle's position and rotation (MatPtLoadBC::GetPositionVars), which live on the device
// between archives: the host can evaluate it only when it reads the time alone (Expression::IsPositionIndependent)
static bool TimeOnlyFunction(MatPtLoadBC *lb) { return lb->function != NULL && lb->function->IsPositionIndependent(); }

const char *GpuTasks_Install(int device, bool fusedStep, int ngpus)
{
    gNumGpus = ngpus < 1 ? 1 : ngpus;
    if (gNumGpus > 1) {
        // -gpus N: z-slabs of cell planes, one per GPU, on the fused kernels; the whole step runs in mpmgpu_slab_step and the
        // output takes the host route (every slab downloads its particles, the reference's writers run on mpm[])
        fusedStep = true;
        gDeviceOutput = false;
        omp_set_dynamic(0);             // the slabs step in lock-step: exactly one thread per GPU
        if (!fmobj->IsThreeD()) return "-gpus N needs a 3D problem (slabs of cell planes along z)";
        if (firstLoadedPt != NULL) return "particle load BCs with -gpus N (particles change slabs)";
    }
    gFusedStep = fusedStep;
    if (firstCrack != NULL) return "cracks present";
    if (fmobj->multiMaterialMode) {
        // material velocity fields + contact (mpmgpu_set_multimaterial): what the device path covers, the rest stays refused
        if (mpmgrid.materialNormalMethod > SPECIFIED_NORMAL) return "multimaterial contact normals by linear or logistic regression (<MultiMaterialMode Normals=\"5|6\">)";
        if (mpmgrid.materialNormalMethod == EACH_MATERIALS_MASS_GRADIENT) return "multimaterial contact with each material's own normal (the reference itself fails there without FMPM order > 1)";
        if (mpmgrid.hasImperfectInterface) return "imperfect interfaces between materials";
        if (ConductionTask::matContactHeating) return "frictional heating in material contact";
        if (bodyFrc.GetXPICOrder() > 1) return "XPIC/FMPM of order > 1 in multimaterial mode";
        if (maxMaterialFields > 8) return "more than 8 material velocity fields";
    }
    if (transportTasks != NULL) {
        // heat conduction runs on the device (mpmgpu_set_conduction); every other transport task and every option of the
        // conduction task the device does not have stays refused
        if (transportTasks != conduction || conduction->GetNextTransportTask() != NULL) return "transport tasks other than conduction (diffusion, poroelasticity, ...)";
        if (firstRigidTempBC != NULL) return "temperature BCs set by rigid particles";
        // particle heat-flux BCs run on the device when they are external fluxes the host can evaluate (mpmgpu_set_particle_heat_fluxes)
        for (MatPtLoadBC *lb = firstHeatFluxPt; lb != NULL; lb = (MatPtLoadBC *)lb->GetNextObject()) {
            if (lb->style == SILENT || lb->direction != EXTERNAL_FLUX || (lb->style == FUNCTION_VALUE && !TimeOnlyFunction(lb)))
                return "particle heat-flux BCs that are silent, coupled or set by a function of position";
            if (lb->ptNum - 1 >= nmpmsNR) return "heat-flux BCs on rigid particles";
        }
        if (ConductionTask::crackTipHeating || ConductionTask::crackContactHeating || ConductionTask::matContactHeating) return "crack-tip or contact heating";
        if (TransportTask::hasXPICOption) return "XPIC/FMPM options for transport tasks";
        // (a mechanical XPIC/FMPM order > 1 does not touch the transport update: UpdateParticlesTask.cpp:85-97)
    }
    // everything the replaced CPU tasks would do on the side must be absent, or the run would silently differ:
    // particle loads / tractions are re-evaluated every step by InitializationTask and GridForcesTask
    // particle loads are re-evaluated every step (MatPtLoadBC::SetParticleFext): values that are functions of time only are
    // evaluated by the host and sent down; silent BCs (need the particle velocity) and function styles (may use the particle
    // position and rotation) would need the current particle state on the host
    for (MatPtLoadBC *lb = firstLoadedPt; lb != NULL; lb = (MatPtLoadBC *)lb->GetNextObject()) {
        if (lb->style == SILENT) return "silent particle load BCs";
        if (lb->style == FUNCTION_VALUE && !TimeOnlyFunction(lb)) return "particle load BCs set by a function of position";
        if (lb->ptNum - 1 >= nmpmsNR) return "load BCs on rigid particles";
    }
    // particle traction BCs run on the device (mpmgpu_set_particle_tractions); the host re-evaluates values that depend on time only
    for (MatPtLoadBC *lb = firstTractionPt; lb != NULL; lb = (MatPtLoadBC *)lb->GetNextObject()) {
        if (lb->style == FUNCTION_VALUE && !TimeOnlyFunction(lb)) return "particle traction BCs set by a function of position";
        if (lb->ptNum - 1 >= nmpmsNR) return "traction BCs on rigid particles";
        if (fmobj->exactTractions) return "<ExactTractions>";
        if (ngpus > 1) return "particle traction BCs with -gpus N";
    }
    // global quantities the reference reads from its nodes or BC objects, which the replaced tasks no longer fill
    for (GlobalQuantity *gq = firstGlobal; gq != NULL; gq = gq->GetNextGlobal()) {
        const int q = gq->quantity;
        if (q == INTERFACE_ENERGY || q == FRICTION_WORK)
            return "global quantities read from the grid (interface energy, friction work)";
        if (q == GRID_KINE_ENERGY) {
            if (fmobj->multiMaterialMode || ngpus > 1) return "grid kinetic energy in multimaterial mode or with -gpus N";
            gGridKineticEnergy = true;
        }
        if (q == TOT_FCONX || q == TOT_FCONY || q == TOT_FCONZ) {
            if (!fmobj->multiMaterialMode) return "contact-force global quantities outside multimaterial mode";
            if (ngpus > 1 || !gDeviceOutput) return "contact-force global quantities with -gpus N or -hostoutput (they are summed on the device)";
            if (archiver->globalTime < 0.) return "contact-force global quantities archived with the particle archives";
            gContactQuantities = true;
        }
        if (q == TOT_REACTX || q == TOT_REACTY || q == TOT_REACTZ) {
            if (ngpus > 1) return "reaction-force global quantities with -gpus N (the slabs do not keep them)";
            gTrackReactions = true;
        }
    }
    if (gContactQuantities && !GlobalsCoveredByDevice())
        return "contact-force global quantities next to a quantity the device does not sum (the host's nodes carry no contact forces)";
    // damping that changes during the run (functions of time, feedback on the kinetic energy: BodyForce.cpp:167-230)
    if (bodyFrc.useFeedback || bodyFrc.usePFeedback || bodyFrc.gridfunction != NULL || bodyFrc.pgridfunction != NULL)
        return "time-dependent or feedback damping";
    // grid body forces that are functions of position and time are added to the grid force by PostForcesTask
    // (BodyForce::GetGridBodyForce, BodyForce.cpp:97-117); the device applies constant gravity only
    if (bodyFrc.hasGridBodyForce) return "grid body force functions (<BodyXForce> ...)";
    // (<EnergyCoupling/>, ConductionTask::adiabatic, runs on the device: mpmgpu_set_energy_coupling)
    // (particle temperatures travel to the device -- mpmgpu_particles.temperature -- whenever they matter: with conduction they are
    // state of the transport task; without it a start off the stress-free temperature gives the first particle update a thermal
    // strain increment res.dT = pTemperature - pPreviousTemperature, UpdateParticlesTask.cpp:246-251, which the device laws carry)
    bool temperatureOffsets = false;
    for (int p = 0; p < nmpmsNR && !temperatureOffsets; p++) temperatureOffsets = mpm[p]->pTemperature != mpm[p]->pPreviousTemperature;
    // custom tasks run on the host particles between the step tasks; only the one that just switches the XPIC/FMPM order is safe
    for (CustomTask *ct = theTasks; ct != NULL; ct = ct->nextTask)
        if (strcmp(ct->TaskName(), "Periodic XPIC Implementation") != 0) return "custom tasks other than PeriodicXPIC";
    // failure handling the replaced ResetElements / PostForces tasks would do on the host objects
    if (fabs(fmobj->restartScaling) > 1.e-6) return "time-step restarts (<RestartScaling>)";
    if (fmobj->deleteLeavingParticles) return "deleting particles that leave the grid (<LeaveLimit> < 0)";
    if (warnings.GetMaxIssues(fmobj->warnParticleDeleted) >= 2) return "deleting nan particles (<DeleteLimit> > 1)";
    if (nmpmsRB != nmpmsNR) return "rigid block particles present";
    if (nmpmsRC != nmpmsRB && !fmobj->multiMaterialMode) return "rigid contact particles outside multimaterial mode";
    if (nmpms != nmpmsNR && MaterialBase::extrapolateRigidBCs) return "rigid BCs by extrapolation";
    if (fmobj->np != PLANE_STRAIN_MPM && fmobj->np != PLANE_STRESS_MPM && fmobj->np != THREED_MPM) return "analysis type";
    if (ElementBase::useGimp != POINT_GIMP && ElementBase::useGimp != UNIFORM_GIMP && ElementBase::useGimp != LINEAR_CPDI &&
        ElementBase::useGimp != QUADRATIC_CPDI && ElementBase::useGimp != BSPLINE_GIMP && ElementBase::useGimp != BSPLINE && ElementBase::useGimp != BSPLINE_CPDI) return "shape functions";
    if (!mpmgrid.IsStructuredEqualElementsGrid()) return "grid is not structured with equal elements";
    for (int i = 0; i < nmat; i++) {
        MaterialBase *mb = theMaterials[i];
        if (mb->artificialViscosity && mb->MaterialID() != 28 && mb->MaterialID() != 9 && mb->MaterialID() != 8) return "artificial viscosity on this material";
        // dF = exp(du) to <DefGradTerms> terms (Neohookean, large-rotation laws): the device uses the defaults, 1 in 3D and 2 in 2D
        if ((mb->MaterialID() == 28 || mb->MaterialID() == 8 || ((mb->MaterialID() == 1 || mb->MaterialID() == 9) && ((Elastic *)mb)->useLargeRotation)) &&
            MaterialBase::incrementalDefGradTerms != (fmobj->IsThreeD() ? 1 : 2)) return "<DefGradTerms> other than the default";
        switch (mb->MaterialID()) {
        case 1: break;          // small- and large-rotation hypoelasticity (Elastic::useLargeRotation -> material slot 7)
        case 28: break;
        case 8: if (((Mooney *)mb)->rubber) return "Mooney with the IdealRubber option"; break;
        case 9:
            {   HardeningLawBase *hl = ((IsoPlasticity *)mb)->plasticLaw;
                if (dynamic_cast<LinearHardening *>(hl) == NULL && dynamic_cast<NonlinearHardening *>(hl) == NULL && dynamic_cast<JohnsonCook *>(hl) == NULL &&
                    hl->lawID != SCGLHARDENING_ID)
                    return "IsoPlasticity hardening law other than Linear, Nonlinear, Nonlinear2, JohnsonCook and SCGL";
            }
            break;
        case 11: {
            RigidMaterial *rm = (RigidMaterial *)mb;
            if (rm->IsRigidBlock()) return "rigid block material";
            if (rm->useControlVelocity) return "rigid material with control velocity";
            if (rm->Vfunction != NULL) {        // the value function of a temperature-setting material is evaluated by the host every step
                if (!rm->setTemperature || rm->setConcentration || !ConductionTask::active || ngpus > 1) return "rigid material with a value function (other than a temperature with conduction)";
                gRigidFunctions = true; gRigidValueFunctions = true;
            }
            if (rm->setConcentration) return "rigid material that sets concentration";
            // (setTemperature: with conduction the device projects these particles' temperatures onto the nodes they touch; the value
            // function that would change the particle temperature in time was refused above)
            if (rm->function != NULL) gRigidFunctions = true;
            break;
        }
        case CONTACTLAW: case COULOMBFRICTIONLAW:        // a contact law's place in theMaterials[] (multimaterial mode): checked pair by pair below
            if (!fmobj->multiMaterialMode) return "material type";
            break;
        default: return "material type";
        }
    }
    // multimaterial mode: the contact law of every pair of material velocity fields must be one the device applies
    std::vector<int> mmField, mmKind; std::vector<double> mmFriction, mmStatic;
    if (fmobj->multiMaterialMode) {
        const int nf = maxMaterialFields;
        mmField.assign(nmat, 0); mmKind.assign((size_t)nf * nf, 0); mmFriction.assign((size_t)nf * nf, 0.); mmStatic.assign((size_t)nf * nf, -1.);
        for (int i = 0; i < nmat; i++) mmField[i] = theMaterials[i]->GetField() >= 0 ? theMaterials[i]->GetField() : 0;
        for (int i = 0; i < nf; i++)
            for (int j = 0; j < nf; j++) {
                if (i == j) continue;
                ContactLaw *cl = mpmgrid.GetMaterialContactLaw(i, j);
                if (cl == NULL) return "multimaterial mode without a contact law for a pair of materials";
                if (cl->IgnoreContact()) continue;                                           // kind 0
                CoulombFriction *cf = dynamic_cast<CoulombFriction *>(cl);
                if (cf == NULL || strcmp(cl->MaterialType(), "Coulomb Friction") != 0) return "contact law other than ignore / stick / frictionless / Coulomb friction";
                mmKind[(size_t)i * nf + j] = cf->IsStick() ? 1 : (cf->IsFrictionless() ? 2 : 3);
                mmFriction[(size_t)i * nf + j] = cf->frictionCoeff; mmStatic[(size_t)i * nf + j] = cf->frictionCoeffStatic;
            }
    }
    const bool is3D = fmobj->IsThreeD();

    // grid: node coordinates per axis exactly as generated (Read_MPM/Generators.cpp:1823-1833)
    const int nx = mpmgrid.horiz + 1, ny = mpmgrid.vert + 1, nz = is3D ? mpmgrid.depth + 1 : 1;
    std::vector<double> xp(nx), yp(ny), zp(nz);
    for (int i = 0; i < nx; i++) xp[i] = nd[1 + i]->x;
    for (int j = 0; j < ny; j++) yp[j] = nd[1 + j * nx]->y;
    for (int k = 0; k < nz && is3D; k++) zp[k] = nd[1 + k * nx * ny]->z;
    mpmgpu_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = MPMGPU_ABI_VERSION; cfg.device = device; cfg.np = fmobj->np;
    cfg.horiz = mpmgrid.horiz; cfg.vert = mpmgrid.vert; cfg.depth = is3D ? mpmgrid.depth : 0;
    cfg.xpts = xp.data(); cfg.ypts = yp.data(); cfg.zpts = is3D ? zp.data() : NULL;
    cfg.gridx = mpmgrid.grid.x; cfg.gridy = mpmgrid.grid.y; cfg.gridz = mpmgrid.grid.z;
    cfg.shape = ElementBase::useGimp; cfg.cpdi_rcrit = ElementBase::rcrit;
    cfg.method = fmobj->mpmApproach; cfg.skip_post_extrapolation = fmobj->skipPostExtrapolation ? 1 : 0;
    cfg.fraction_usf = fractionUSF;
    cfg.xpic_order = bodyFrc.GetXPICOrder(); cfg.using_fmpm = bodyFrc.UsingFMPM() ? 1 : 0;
    cfg.grid_damping = bodyFrc.GetGridDamping(mtime); cfg.particle_damping = bodyFrc.GetParticleDamping(mtime);
    if (bodyFrc.gravity) { cfg.gravity[0] = bodyFrc.gforce.x; cfg.gravity[1] = bodyFrc.gforce.y; cfg.gravity[2] = bodyFrc.gforce.z; }
    cfg.kernel_path = fusedStep ? 0 : 1;    // per-task entry points keep the reference's task timing report meaningful; -fused runs
                                            // mpmgpu_step (fused kernels when the problem is eligible, per-task kernels otherwise)
    // -gpus N: split the occupied cell planes evenly (the first and last slab reach the grid edges), as slab.py::slab_bounds
    std::vector<std::vector<int> > slabSel(gNumGpus);
    std::vector<int> slabLo(gNumGpus, 0), slabHi(gNumGpus, 0);
    if (gNumGpus > 1) {
        int ndev = 0;
        const int perPlane = mpmgrid.horiz * mpmgrid.vert;
        int kFirst = mpmgrid.depth, kLast = 0;
        for (int p = 0; p < nmpmsNR; p++) { const int k = (mpm[p]->inElem - 1) / perPlane; if (k < kFirst) kFirst = k; if (k + 1 > kLast) kLast = k + 1; }
        const int planes = kLast - kFirst;
        if (planes < gNumGpus) return "-gpus N: fewer occupied cell planes than GPUs";
        for (int r = 0; r < gNumGpus; r++) {
            slabLo[r] = r == 0 ? 0 : kFirst + (int)(((long long)planes * r) / gNumGpus);
            slabHi[r] = r == gNumGpus - 1 ? mpmgrid.depth : kFirst + (int)(((long long)planes * (r + 1)) / gNumGpus);
        }
        for (int p = 0; p < nmpms; p++) {
            if (p >= nmpmsNR) { for (int r = 0; r < gNumGpus; r++) slabSel[r].push_back(p); continue; }     // rigid-BC particles: on every slab
            const int k = (mpm[p]->inElem - 1) / perPlane;
            for (int r = 0; r < gNumGpus; r++) if (k >= slabLo[r] && k < slabHi[r]) { slabSel[r].push_back(p); break; }
        }
        (void)ndev;
        gSlabs.assign(gNumGpus, (mpmgpu_ctx *)NULL);
        for (int r = 0; r < gNumGpus; r++) {
            mpmgpu_config c = cfg;
            c.device = device + r;
            c.kernel_path = 2;
            const long long want = (long long)(1.3 * (double)slabSel[r].size()) + 1024;
            c.max_particles = (int)(want < 65536 ? 65536 : want);
            if (mpmgpu_create(&c, &gSlabs[r]) != MPMGPU_OK) return mpmgpu_last_error(NULL);
            if (mpmgpu_slab_configure(gSlabs[r], slabLo[r], slabHi[r], r > 0 ? 1 : 0, r < gNumGpus - 1 ? 1 : 0, 65536) != MPMGPU_OK) return mpmgpu_last_error(gSlabs[r]);
        }
        gCtx = gSlabs[0];
    }
    else if (mpmgpu_create(&cfg, &gCtx) != MPMGPU_OK) return mpmgpu_last_error(NULL);
    // one call on every context, first failure reported
#define ALL_CTX(CALL) do { if (gSlabs.empty()) { mpmgpu_ctx *ctx_ = gCtx; if ((CALL) != MPMGPU_OK) return mpmgpu_last_error(ctx_); } \
                           else for (mpmgpu_ctx *ctx_ : gSlabs) if ((CALL) != MPMGPU_OK) return mpmgpu_last_error(ctx_); } while (0)

    // materials: the block GetCopyOfMechanicalProps would hand out (Elastic::FillUnrotatedElasticProperties)
    std::vector<mpmgpu_material> mats(nmat);
    for (int i = 0; i < nmat; i++) {
        MaterialBase *mb = theMaterials[i];
        mpmgpu_material &m = mats[i];
        memset(&m, 0, sizeof m);
        m.p[0] = mb->rho; m.p[1] = mb->heatCapacity; m.p[2] = mb->matUsePDamping ? mb->matPdamping : -1.;
        m.n_history = mb->NumberOfHistoryDoubles();
        if (mb->artificialViscosity) { m.p[3] = 1.; m.p[4] = mb->avA1; m.p[5] = mb->avA2; }
        if (mb->MaterialID() == 1) {
            IsotropicMat *im = (IsotropicMat *)mb;
            m.kind = MPMGPU_MAT_ISOTROPIC;
            const ElasticProperties &e = im->pr;
            if (is3D) {
                m.p[8] = e.C[0][0]; m.p[9] = e.C[0][1]; m.p[10] = e.C[0][2]; m.p[11] = e.C[1][1]; m.p[12] = e.C[1][2]; m.p[13] = e.C[2][2];
                m.p[14] = e.C[3][3]; m.p[15] = e.C[4][4]; m.p[16] = e.C[5][5];
                m.p[17] = e.alpha[0]; m.p[18] = e.alpha[1]; m.p[19] = e.alpha[2];
            } else {
                m.p[8] = e.C[1][1]; m.p[9] = e.C[1][2]; m.p[11] = e.C[2][2]; m.p[16] = e.C[3][3];
                m.p[21] = e.C[4][1]; m.p[22] = e.C[4][2]; m.p[23] = e.C[4][4]; m.p[24] = e.C[5][1];
                m.p[17] = e.alpha[1]; m.p[18] = e.alpha[2]; m.p[19] = e.alpha[4];
            }
            m.p[20] = im->gamma0;
            m.p[7] = im->useLargeRotation ? 1. : 0.;
        } else if (mb->MaterialID() == 28) {        // Neohookean::GetCopyOfMechanicalProps hands out pr (Neohookean.cpp:143-150)
            Neohookean *nm = (Neohookean *)mb;
            m.kind = MPMGPU_MAT_NEOHOOKEAN;
            m.p[8] = nm->pr.Gsp; m.p[9] = nm->pr.Ksp; m.p[10] = nm->pr.Lamesp; m.p[11] = nm->UofJOption; m.p[12] = nm->CTE1; m.p[13] = nm->gamma0;
        } else if (mb->MaterialID() == 8) {         // Mooney: specific moduli (Mooney.cpp:139-147, HyperElastic.cpp:63-81)
            Mooney *mm = (Mooney *)mb;
            m.kind = MPMGPU_MAT_MOONEY;
            m.p[8] = mm->G1sp; m.p[9] = mm->G2sp; m.p[10] = mm->Ksp; m.p[11] = mm->UofJOption; m.p[12] = mm->CTE1; m.p[13] = mm->gamma0;
        } else if (mb->MaterialID() == 9) {         // IsoPlasticity::pr + LinearHardening reduced properties
            IsoPlasticity *pm = (IsoPlasticity *)mb;
            HardeningLawBase *hl = pm->plasticLaw;
            LinearHardening *lh = dynamic_cast<LinearHardening *>(hl);
            NonlinearHardening *nh = dynamic_cast<NonlinearHardening *>(hl);
            JohnsonCook *jc = dynamic_cast<JohnsonCook *>(hl);
            m.kind = MPMGPU_MAT_ISOPLASTICITY;
            m.p[8] = pm->pr.Gred; m.p[9] = pm->pr.Kred; m.p[10] = hl->yldred; m.p[12] = pm->CTE3; m.p[13] = pm->gamma0;
            m.p[14] = 1.e50; m.p[15] = hl->yldredMin;
            if (lh != NULL) { m.p[11] = lh->Epred; m.p[14] = lh->alphaMax; }
            else if (nh != NULL) {      // Nonlinear (2) and its subclass Nonlinear2 (6): NonlinearHardening.cpp:49-60, Nonlinear2Hardening.cpp:28-39
                m.p[16] = dynamic_cast<Nonlinear2Hardening *>(hl) != NULL ? 6. : 2.;
                m.p[17] = nh->beta; m.p[18] = nh->npow; m.p[14] = nh->alphaMax;
            } else if (hl->lawID == SCGLHARDENING_ID) {        // SCGL (4): SCGLHardening.cpp:72-91, :138-198
                SCGLHardening *sc = (SCGLHardening *)hl;
                m.p[16] = 4.;
                m.p[17] = sc->beta; m.p[18] = sc->nhard; m.p[19] = sc->yldMaxred; m.p[20] = sc->GPpred; m.p[21] = sc->GTp; m.p[25] = thermal.reference;
            } else {                    // Johnson-Cook (3): JohnsonCook.cpp:106-127
                m.p[16] = 3.;
                m.p[17] = jc->Bred; m.p[18] = jc->njc; m.p[19] = jc->Cjc; m.p[20] = jc->ep0jc; m.p[21] = jc->Djc; m.p[22] = jc->n2jc;
                m.p[23] = jc->Tmjc; m.p[24] = jc->mjc; m.p[25] = thermal.reference; m.p[26] = jc->edotMin; m.p[27] = jc->eminTerm;
            }
            m.p[7] = pm->useLargeRotation ? 1. : 0.;
        } else if (mb->MaterialID() == CONTACTLAW || mb->MaterialID() == COULOMBFRICTIONLAW) {
            m.kind = MPMGPU_MAT_NONE; m.n_history = 0;          // keeps the particles' material numbers in place
            memset(m.p, 0, sizeof m.p);
        } else if (((RigidMaterial *)mb)->IsRigidContact()) {       // rigid contact particles: their own velocity field (multimaterial mode)
            m.kind = MPMGPU_MAT_RIGIDCONTACT; m.n_history = 0;
            m.p[0] = mb->rho;
        } else {                                     // rigid BC particles: directions they control
            m.kind = MPMGPU_MAT_RIGIDBC; m.n_history = 0;
            m.p[8] = ((RigidMaterial *)mb)->setDirection;
            m.p[9] = ((RigidMaterial *)mb)->mirrored;
            m.p[10] = ((RigidMaterial *)mb)->setTemperature ? 1. : 0.;
        }
    }
    ALL_CTX(mpmgpu_set_materials(ctx_, nmat, mats.data()));
    if (ConductionTask::adiabatic && mpmgpu_set_energy_coupling(gCtx, 1) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    if (ConductionTask::active) {
        std::vector<double> kc(nmat, 0.);
        for (int i = 0; i < nmat; i++) kc[i] = theMaterials[i]->kCond;          // conductivity / rho (MaterialBaseMPM.cpp:233)
        if (mpmgpu_set_conduction(gCtx, nmat, kc.data()) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    }
    if (fmobj->multiMaterialMode) {
        mpmgpu_multimaterial mm;
        memset(&mm, 0, sizeof mm);
        mm.n_fields = maxMaterialFields; mm.field_of_material = mmField.data();
        mm.normal_method = mpmgrid.materialNormalMethod; mm.contact_by_displacements = mpmgrid.contactByDisplacements ? 1 : 0;
        mm.position_cutoff = mpmgrid.positionCutoff;
        mm.contact_normal[0] = mpmgrid.contactNormal.x; mm.contact_normal[1] = mpmgrid.contactNormal.y; mm.contact_normal[2] = mpmgrid.contactNormal.z;
        mm.law_kind = mmKind.data(); mm.law_friction = mmFriction.data(); mm.law_static = mmStatic.data();
        mm.rigid_gradient_bias = mpmgrid.rigidGradientBias;         // (squared by MeshInfo::MaterialOutput already)
        if (mpmgpu_set_multimaterial(gCtx, &mm) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    }

    // particles: AoS heap objects -> SoA
    const int n = nmpms;
    std::vector<double> pos(3 * n), vel(3 * n), mp(n), lp(3 * n), sp(6 * n), pr(n), ep(6 * n), wrot(3 * n), epl(6 * n), en(6 * n), pf(3 * n),
        hist((size_t)MPMGPU_MAX_HISTORY * n, 0.);
    std::vector<int> elem(n), matn(n), cross(n);
    bool anyFext = firstLoadedPt != NULL;
    {
        std::vector<char> seen((size_t)n, 0);
        for (MatPtLoadBC *lb = firstLoadedPt; lb != NULL; lb = (MatPtLoadBC *)lb->GetNextObject())
            if (!seen[lb->ptNum - 1]) { seen[lb->ptNum - 1] = 1; gLoadPts.push_back(lb->ptNum - 1); }
    }
    for (int p = 0; p < n; p++) {
        MPMBase *m = mpm[p];
        pos[p] = m->pos.x; pos[n + p] = m->pos.y; pos[2 * n + p] = m->pos.z;
        vel[p] = m->vel.x; vel[n + p] = m->vel.y; vel[2 * n + p] = m->vel.z;
        mp[p] = m->mp; lp[p] = m->mpm_lp.x; lp[n + p] = m->mpm_lp.y; lp[2 * n + p] = m->mpm_lp.z;
        sp[p] = m->sp.xx; sp[n + p] = m->sp.yy; sp[2 * n + p] = m->sp.zz; sp[3 * n + p] = m->sp.yz; sp[4 * n + p] = m->sp.xz; sp[5 * n + p] = m->sp.xy;
        pr[p] = m->pressure;
        ep[p] = m->ep.xx; ep[n + p] = m->ep.yy; ep[2 * n + p] = m->ep.zz; ep[3 * n + p] = m->ep.yz; ep[4 * n + p] = m->ep.xz; ep[5 * n + p] = m->ep.xy;
        wrot[p] = m->wrot.xy; wrot[n + p] = m->wrot.xz; wrot[2 * n + p] = m->wrot.yz;
        epl[p] = m->eplast.xx; epl[n + p] = m->eplast.yy; epl[2 * n + p] = m->eplast.zz; epl[3 * n + p] = m->eplast.yz; epl[4 * n + p] = m->eplast.xz; epl[5 * n + p] = m->eplast.xy;
        en[p] = m->workEnergy; en[n + p] = m->resEnergy; en[2 * n + p] = m->heatEnergy; en[3 * n + p] = m->entropy; en[4 * n + p] = m->plastEnergy;
        en[5 * n + p] = m->pPreviousTemperature;
        pf[p] = m->pFext.x; pf[n + p] = m->pFext.y; pf[2 * n + p] = m->pFext.z;
        if (m->pFext.x != 0. || m->pFext.y != 0. || m->pFext.z != 0.) anyFext = true;
        elem[p] = m->inElem; matn[p] = m->matnum; cross[p] = m->elementCrossings;
        const int nh = theMaterials[m->MatID()]->NumberOfHistoryDoubles();
        for (int k = 0; k < nh && k < MPMGPU_MAX_HISTORY && m->matData != NULL; k++) hist[(size_t)k * n + p] = ((double *)m->matData)[k];
    }
    mpmgpu_particles h;
    memset(&h, 0, sizeof h);
    h.n = n; h.n_nonrigid = nmpmsNR;
    h.pos = pos.data(); h.vel = vel.data(); h.mp = mp.data(); h.lp = lp.data(); h.in_elem = elem.data(); h.matnum = matn.data();
    h.sp = sp.data(); h.pressure = pr.data(); h.ep = ep.data(); h.wrot = wrot.data(); h.eplast = epl.data(); h.energies = en.data();
    h.pfext = anyFext ? pf.data() : NULL; h.crossings = cross.data(); h.history = hist.data();
    std::vector<double> temp0;
    gThermal = ConductionTask::active || ConductionTask::adiabatic || temperatureOffsets;
    if (gThermal) {
        temp0.resize(n);
        for (int p = 0; p < n; p++) temp0[p] = mpm[p]->pTemperature;
        h.temperature = temp0.data();
    }
    if (!gSlabs.empty()) {
        // every slab gets the columns of its particles, with the host's particle numbers as ids
        for (int r = 0; r < gNumGpus; r++) {
            const std::vector<int> &sel = slabSel[r];
            const size_t m = sel.size();
            auto cols = [&](const std::vector<double> &a, int ncomp) {
                std::vector<double> o((size_t)ncomp * m);
                for (int c = 0; c < ncomp; c++) for (size_t q = 0; q < m; q++) o[(size_t)c * m + q] = a[(size_t)c * n + sel[q]];
                return o;
            };
            auto icols = [&](const std::vector<int> &a) { std::vector<int> o(m); for (size_t q = 0; q < m; q++) o[q] = a[sel[q]]; return o; };
            std::vector<double> spos = cols(pos, 3), svel = cols(vel, 3), smp = cols(mp, 1), slp = cols(lp, 3), ssp = cols(sp, 6), spr = cols(pr, 1), sep = cols(ep, 6),
                swrot = cols(wrot, 3), sepl = cols(epl, 6), sen = cols(en, 6), shist = cols(hist, MPMGPU_MAX_HISTORY);
            std::vector<int> selem = icols(elem), smat = icols(matn), scross = icols(cross), ids(sel.begin(), sel.end());
            mpmgpu_particles hs;
            memset(&hs, 0, sizeof hs);
            hs.n = (int)m;
            hs.n_nonrigid = 0;
            for (size_t q = 0; q < m; q++) if (sel[q] < nmpmsNR) hs.n_nonrigid++;
            hs.pos = spos.data(); hs.vel = svel.data(); hs.mp = smp.data(); hs.lp = slp.data(); hs.in_elem = selem.data(); hs.matnum = smat.data();
            hs.sp = ssp.data(); hs.pressure = spr.data(); hs.ep = sep.data(); hs.wrot = swrot.data(); hs.eplast = sepl.data(); hs.energies = sen.data();
            hs.crossings = scross.data(); hs.history = shist.data(); hs.ids = ids.data();
            if (mpmgpu_upload_particles(gSlabs[r], &hs) != MPMGPU_OK) return mpmgpu_last_error(gSlabs[r]);
        }
    }
    else if (mpmgpu_upload_particles(gCtx, &h) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    ALL_CTX(mpmgpu_set_time_step(ctx_, timestep, strainTimestepFirst, strainTimestepLast));
    if (gSlabs.empty())
    {   // constants of the archive records (ArchiveData.cpp:820-875): original position, initial material angles, 2D thickness
        std::vector<double> op(3 * (size_t)n), ang(3 * (size_t)n);
        double thick = is3D ? 1. : mpm[0]->thickness();
        for (int p = 0; p < n; p++) {
            MPMBase *m = mpm[p];
            op[p] = m->origpos.x; op[n + p] = m->origpos.y; op[2 * (size_t)n + p] = m->origpos.z;
            ang[p] = m->GetAnglez0InRadians(); ang[n + p] = m->GetAngley0InRadians(); ang[2 * (size_t)n + p] = m->GetAnglex0InRadians();
            if (!is3D && m->thickness() != thick) gDeviceOutput = false;    // one thickness per run on the device: the host writes otherwise
        }
        if (mpmgpu_set_archive_origin(gCtx, op.data(), ang.data(), thick) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    }

    // grid velocity BCs in list order
    std::vector<int> bnode, bact, bsym, brefl; std::vector<double> bnorm, bval, bratio;
    bool anyReflected = false;
    for (NodalVelBC *bc = firstVelocityBC; bc != NULL; bc = (NodalVelBC *)bc->GetNextObject()) {
        gBCs.push_back(bc);
        bnode.push_back(bc->nodeNum);
        bnorm.push_back(bc->norm.x); bnorm.push_back(bc->norm.y); bnorm.push_back(bc->norm.z);
        bact.push_back(bc->GetNodeNum(mtime) > 0 ? 1 : 0);
        bval.push_back(bact.back() ? bc->BCValue(mtime) : 0.);
        bsym.push_back(nd[bc->nodeNum]->fixedDirection & ANYSYMMETRYPLANE_DIRECTION);
        if (bc->style != CONSTANT_VALUE || bc->GetBCFirstTime() > 0.) gBCsVary = true;
        // symmetry-plane neighbours reflect the node across the plane (Generators.cpp:2178-2190); BCs that rigid particles
        // would mirror at run time (SetMirroredVelBC) are handled by the device's rigid-BC projection instead
        brefl.push_back(bc->reflectedNode); bratio.push_back(bc->reflectRatio);
        if (bc->reflectedNode >= 0) anyReflected = true;
    }
    if (gTrackReactions) {
        if (mpmgpu_track_reactions(gCtx, 1) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
        // carriers for the reactions of the rigid-particle BCs, linked after the grid BCs (where ProjectRigidBCsTask would put its own)
        gRigidCarriers.assign((size_t)nmat, (NodalVelBC *)NULL);
        BoundaryCondition *last = gBCs.empty() ? NULL : gBCs.back();
        for (int m = 0; m < nmat; m++) {
            if (!theMaterials[m]->IsRigidBC()) continue;
            const int keep = nd[1]->fixedDirection;         // (the constructor marks the node's dof as fixed: undone, this BC fixes nothing)
            NodalVelBC *c = new NodalVelBC(1, X_DIRECTION, CONSTANT_VALUE, 0., 0., 0., 0.);
            nd[1]->fixedDirection = keep;
            c->SetID(m + 1);
            if (last != NULL) last->SetNextObject(c); else firstVelocityBC = c;
            if (firstRigidVelocityBC == NULL) firstRigidVelocityBC = c;
            last = c;
            gRigidCarriers[m] = c;
        }
    }
    ALL_CTX(mpmgpu_set_velocity_bcs(ctx_, (int)bnode.size(), bnode.data(), bnorm.data(), bval.data(), bact.data(), bsym.data()));
    if (anyReflected) ALL_CTX(mpmgpu_set_velocity_bc_reflections(ctx_, (int)bnode.size(), brefl.data(), bratio.data()));
    if (ConductionTask::active && firstHeatFluxPt != NULL) {      // MatPtHeatFluxBC list in list order
        std::vector<int> tp, tf; std::vector<double> tv;
        for (MatPtLoadBC *lb = firstHeatFluxPt; lb != NULL; lb = (MatPtLoadBC *)lb->GetNextObject()) {
            MatPtHeatFluxBC *hb = (MatPtHeatFluxBC *)lb;
            gHeatFluxes.push_back(hb);
            tp.push_back(hb->ptNum - 1); tf.push_back(hb->face); tv.push_back(hb->BCValue(mtime));
            if (hb->style != CONSTANT_VALUE || hb->GetBCFirstTime() > 0.) gHeatFluxesVary = true;
        }
        if (mpmgpu_set_particle_heat_fluxes(gCtx, (int)tp.size(), tp.data(), tf.data(), tv.data()) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    }
    if (firstTractionPt != NULL) {      // MatPtTractionBC list in list order
        std::vector<int> tp, tf, td; std::vector<double> tv;
        for (MatPtLoadBC *lb = firstTractionPt; lb != NULL; lb = (MatPtLoadBC *)lb->GetNextObject()) {
            MatPtTractionBC *tb = (MatPtTractionBC *)lb;
            gTractions.push_back(tb);
            tp.push_back(tb->ptNum - 1); tf.push_back(tb->face); td.push_back(tb->direction); tv.push_back(tb->BCValue(mtime));
            if (tb->style != CONSTANT_VALUE || tb->GetBCFirstTime() > 0.) gTractionsVary = true;
        }
        if (mpmgpu_set_particle_tractions(gCtx, (int)tp.size(), tp.data(), tf.data(), td.data(), tv.data()) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    }
    if (!gSlabs.empty()) {
        // the slabs join the two NCCL communicators of the run (collective: one thread per GPU)
        char ids[256];
        if (mpmgpu_nccl_unique_ids(ids) != MPMGPU_OK) return "mpmgpu_nccl_unique_ids failed (libnccl.so.2 not found?)";
        std::vector<int> rcs(gSlabs.size(), MPMGPU_OK);
#pragma omp parallel num_threads((int)gSlabs.size())
        {
            const int r = omp_get_thread_num();
            if (r < (int)gSlabs.size()) rcs[r] = mpmgpu_slab_connect(gSlabs[r], r, (int)gSlabs.size(), ids);
        }
        for (size_t r = 0; r < gSlabs.size(); r++) if (rcs[r] != MPMGPU_OK) return mpmgpu_last_error(gSlabs[r]);
    }

    if (ConductionTask::active)
        for (NodalTempBC *bc = firstTempBC; bc != NULL; bc = (NodalTempBC *)bc->GetNextObject()) {
            gTempBCs.push_back(bc);
            if (bc->style != CONSTANT_VALUE || bc->GetBCFirstTime() > 0.) gTempBCsVary = true;
        }
    // swap the CPU task objects for GPU ones, keeping order and names (custom-task runner stays)
    MPMTask *prev = NULL;
    for (MPMTask *t = firstMPMTask; t != NULL;) {
        MPMTask *next = (MPMTask *)t->GetNextTask();
        const int code = TaskCode(t->GetTaskName());
        if (strcmp(t->GetTaskName(), "Decipher Crack and Material Fields") == 0) {
            // InitVelocityFieldsTask (multimaterial mode without cracks): it assigns every particle's velocity field on the host
            // objects; the device knows the field from the particle's material.  Dropped from the list.
            if (prev) prev->SetNextTask(next); else firstMPMTask = next;
            t = next;
            continue;
        }
        if (code >= 0) {
            GpuTask *g = new GpuTask(t->GetTaskName(), code);
            g->SetNextTask(next);
            if (prev) prev->SetNextTask(g); else firstMPMTask = g;
            prev = g;
        } else prev = t;
        t = next;
    }
    extern int gPollInterval;
    if (fusedStep && gPollInterval > 1 && gSlabs.empty() && mpmgpu_set_poll_interval(gCtx, gPollInterval) != MPMGPU_OK) return mpmgpu_last_error(gCtx);
    std::cout << "GPU TASKS: tasks 1-9,11 run on libmpmgpu (device " << device << ", " << n << " particles"
              << (fusedStep ? ", whole-step entry point" : ", per-task entry points") << ")" << std::endl;
    if (!gSlabs.empty()) {
        std::cout << "GPU SLABS: " << gSlabs.size() << " GPUs, cell planes";
        for (int r = 0; r < gNumGpus; r++) std::cout << " [" << slabLo[r] << "," << slabHi[r] << "):" << slabSel[r].size();
        std::cout << " particles (rigid-BC particles on every slab)" << std::endl;
    }
    if (gContactQuantities && !gDeviceOutput) return "contact-force global quantities when the archives have to be written by the host";
    return NULL;
}

void GpuTasks_SetDeviceOutput(bool on) { gDeviceOutput = on; }

void GpuTasks_Finish(void)
{
    if (!gCtx) return;
    if (gArchiveCount > 0)
        std::cout << "GPU ARCHIVES: " << gArchiveCount << " particle archives packed on the device, " << gArchiveBytes / gArchiveCount
                  << " bytes each (D2H), " << 1.e3 * gArchiveSeconds / (double)gArchiveCount << " ms each including the file write" << std::endl;
    gHostStale = true;
    try { DownloadToHost(); } catch (...) {}
    if (!gSlabs.empty()) {
        long long out = 0, in = 0;
        for (mpmgpu_ctx *c : gSlabs) { long long o = 0, i = 0; mpmgpu_slab_migrated(c, &o, &i); out += o; in += i; }
        std::cout << "GPU SLABS: " << out << " particle rows changed slabs during the run" << std::endl;
        // (communicators are torn down collectively)
#pragma omp parallel num_threads((int)gSlabs.size())
        {
            const int r = omp_get_thread_num();
            if (r < (int)gSlabs.size()) mpmgpu_destroy(gSlabs[r]);
        }
        gSlabs.clear();
    } else
    mpmgpu_destroy(gCtx);
    gCtx = NULL;
}

// ---- the driver: Common/System/main.cpp steps with the install hook between preparations and analysis ----
static bool gDeviceOutputSwitch = true;
int gPollInterval = 1;              // -poll K (with -fused): look at the device's status word every K-th step (mpmgpu_set_poll_interval)
void GpuTasks_SetDeviceOutput(bool on);

int main(int argc, const char *argv[])
{
    int numProcs = 1, device = 0, arg = 1, ngpus = 1;
    bool useGpu = true, fused = false;
    for (; arg < argc && argv[arg][0] == '-'; arg++) {
        if (strcmp(argv[arg], "-np") == 0 && arg + 1 < argc) sscanf(argv[++arg], "%d", &numProcs);
        else if (strcmp(argv[arg], "-gpu") == 0 && arg + 1 < argc) sscanf(argv[++arg], "%d", &device);
        else if (strcmp(argv[arg], "-gpus") == 0 && arg + 1 < argc) sscanf(argv[++arg], "%d", &ngpus);
        else if (strcmp(argv[arg], "-cpu") == 0) useGpu = false;
        else if (strcmp(argv[arg], "-fused") == 0) fused = true;
        else if (strcmp(argv[arg], "-hostoutput") == 0) gDeviceOutputSwitch = false;
        else if (strcmp(argv[arg], "-poll") == 0 && arg + 1 < argc) sscanf(argv[++arg], "%d", &gPollInterval);
        else { std::cerr << "usage: NairnMPM_gpu [-np N] [-gpu DEVICE] [-gpus N] [-fused] [-poll K] [-hostoutput] [-cpu] input.fmcmd" << std::endl; return 1; }
    }
    if (arg + 1 != argc) { std::cerr << "usage: NairnMPM_gpu [-np N] [-gpu DEVICE] [-gpus N] [-fused] [-poll K] [-hostoutput] [-cpu] input.fmcmd" << std::endl; return 1; }
    fmobj = new NairnMPM();
    omp_set_num_threads(numProcs);
    fmobj->SetNumberOfProcessors(numProcs);
    int rv = fmobj->ReadFile(argv[arg], false);
    if (rv != 0) return rv;
    InitRandom(fmobj->randseed > 0 ? (unsigned int)fmobj->randseed : 0);
    try {
        fmobj->StartResultsOutput();
        fmobj->CMStartResultsOutput();
        fmobj->CMPreparations();
        if (useGpu) {
            GpuTasks_SetDeviceOutput(gDeviceOutputSwitch);
            const char *why = GpuTasks_Install(device, fused, ngpus);
            if (why != NULL) { std::cerr << "NairnMPM_gpu: cannot run this input on libmpmgpu: " << why << std::endl; return 2; }
        }
        fmobj->CMAnalysis(false);
        GpuTasks_Finish();
    }
    catch (CommonException &e) { e.Display(); return 3; }
    catch (const char *msg) { std::cerr << msg << std::endl; return 3; }
    return 0;
}
