// TEST INFRASTRUCTURE (oracle) -- not product code, never linked into libmpmgpu.
//
// C accessors around the UNMODIFIED reference NairnMPM objects, so that parity tests and
// bench.py's reference arm can (1) parse an XML input with the reference's own reader,
// (2) advance it one MPMStep -- or one MPMTask -- at a time, and (3) read back particle and
// node state as flat arrays.  Built by oracle/build_ref.sh into oracle/_ref/libnairnmpm_ref.so.
//
// What it drives (reference file:line):
//   main()                       Common/System/main.cpp:25-140      (steps 1-3 reproduced in ref_open)
//   CommonAnalysis::StartAnalysis Common/System/CommonAnalysis.cpp:49-62 (all but CMAnalysis)
//   NairnMPM::CMAnalysis         NairnMPM_Class/NairnMPM.cpp:171-200 (Step0 + ValidateOptions + loop body)
//   NairnMPM::MPMStep            NairnMPM_Class/NairnMPM.cpp:284-335 (ref_step / ref_run_task)
#include <omp.h>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <iostream>
#include <fstream>

// state lives in protected/private members; this TU only reads it
#define private public
#define protected public
#include "NairnMPM_Class/NairnMPM.hpp"
#include "NairnMPM_Class/MPMTask.hpp"
#include "NairnMPM_Class/MeshInfo.hpp"
#include "NairnMPM_Class/ResetElementsTask.hpp"
#include "MPM_Classes/MPMBase.hpp"
#include "Nodes/NodalPoint.hpp"
#include "Nodes/CrackVelocityField.hpp"
#include "Nodes/MatVelocityField.hpp"
#include "Elements/ElementBase.hpp"
#include "Materials/MaterialBase.hpp"
#include "Materials/IsotropicMat.hpp"
#include "Materials/Neohookean.hpp"
#include "Materials/NonlinearHardening.hpp"
#include "Materials/Nonlinear2Hardening.hpp"
#include "Materials/JohnsonCook.hpp"
#include "Boundary_Conditions/MatPtTractionBC.hpp"
#include "Materials/SCGLHardening.hpp"
#include "Global_Quantities/ThermalRamp.hpp"
#include "Materials/Mooney.hpp"
#include "Materials/IsoPlasticity.hpp"
#include "Materials/HardeningLawBase.hpp"
#include "Materials/LinearHardening.hpp"
#include "Materials/RigidMaterial.hpp"
#include "Boundary_Conditions/NodalVelBC.hpp"
#include "Materials/ContactLaw.hpp"
#include "Materials/CoulombFriction.hpp"
#include "Boundary_Conditions/MatPtLoadBC.hpp"
#include "Boundary_Conditions/NodalTempBC.hpp"
#include "Boundary_Conditions/MatPtHeatFluxBC.hpp"
#include "Custom_Tasks/ConductionTask.hpp"
#include "Custom_Tasks/TransportTask.hpp"
#include "Global_Quantities/BodyForce.hpp"
#include "Custom_Tasks/CustomTask.hpp"
#include "Custom_Tasks/ConductionTask.hpp"
#include "Exceptions/CommonException.hpp"
#include "System/UnitsController.hpp"
#undef private
#undef protected

extern MPMTask *firstMPMTask;
extern double timestep, strainTimestepFirst, strainTimestepLast, fractionUSF, mtime;
extern int maxShapeNodes;

static std::ofstream g_log;
static std::streambuf *g_coutbuf = NULL;
static char g_err[1024] = "";

static void set_err(const char *what) { snprintf(g_err, sizeof g_err, "%s", what); }

extern "C" {

const char *ref_last_error(void) { return g_err; }

// Parse the XML with the reference reader and run all set-up short of the time loop.
// The reference's text report (cout) goes to logpath (or /dev/null).
int ref_open(const char *xml, int nprocs, const char *logpath)
{
    try {
        g_log.open(logpath && logpath[0] ? logpath : "/dev/null");
        g_coutbuf = std::cout.rdbuf(g_log.rdbuf());
        fmobj = new NairnMPM();
        int maxThreads = omp_get_max_threads();
        int avail = omp_get_num_procs();
        int maxProcs = maxThreads > avail ? maxThreads : avail;
        if (nprocs <= 0 || nprocs > maxProcs) nprocs = maxProcs;
        omp_set_num_threads(nprocs);
        fmobj->SetNumberOfProcessors(nprocs);
        int rv = fmobj->ReadFile(xml, false);
        if (rv != 0) { set_err("ReadFile failed (see log)"); return rv; }
        InitRandom(fmobj->randseed > 0 ? (unsigned int)fmobj->randseed : 0);
        fmobj->StartResultsOutput();
        fmobj->CMStartResultsOutput();
        fmobj->CMPreparations();
        CustomTask *nextTask = theTasks;
        while (nextTask != NULL) nextTask = nextTask->Step0Calculation();
        fmobj->ValidateOptions();
        std::cout.flush();
        return 0;
    }
    catch (CommonException &e) { set_err(e.Message()); }
    catch (CommonException *e) { set_err(e->Message()); }
    catch (const char *m) { set_err(m); }
    catch (std::exception &e) { set_err(e.what()); }
    catch (...) { set_err("unknown exception in ref_open"); }
    return -1;
}

// n full steps: the body of the loop in NairnMPM::CMAnalysis minus archiving
int ref_step(int n)
{
    try {
        for (int i = 0; i < n; i++) {
            fmobj->mstep++;
            fmobj->MPMStep();
            mtime += timestep;
        }
        return 0;
    }
    catch (CommonException &e) { set_err(e.Message()); }
    catch (CommonException *e) { set_err(e->Message()); }
    catch (const char *m) { set_err(m); }
    catch (...) { set_err("unknown exception in ref_step"); }
    return -1;
}

int ref_num_tasks(void)
{
    int n = 0;
    for (MPMTask *t = firstMPMTask; t != NULL; t = (MPMTask *)t->GetNextTask()) n++;
    return n;
}

const char *ref_task_name(int i)
{
    MPMTask *t = firstMPMTask;
    while (i-- > 0 && t != NULL) t = (MPMTask *)t->GetNextTask();
    return t ? t->GetTaskName() : "";
}

// Execute task i of the current step.  Caller runs i=0..ntasks-1 in order; task 0 starts the
// step (mstep++) and the last task ends it (mtime += timestep).
int ref_run_task(int i)
{
    try {
        int n = ref_num_tasks();
        if (i == 0) fmobj->mstep++;
        MPMTask *t = firstMPMTask;
        for (int k = 0; k < i && t != NULL; k++) t = (MPMTask *)t->GetNextTask();
        if (t == NULL) { set_err("no such task"); return -1; }
        t->Execute(0);
        if (i == n - 1) mtime += timestep;
        return 0;
    }
    catch (CommonException &e) { set_err(e.Message()); }
    catch (CommonException *e) { set_err(e->Message()); }
    catch (const char *m) { set_err(m); }
    catch (...) { set_err("unknown exception in ref_run_task"); }
    return -1;
}

// scalar facts about the run; keys are positional, see tests/refharness.py
// ints:  0 np(analysis) 1 is3D 2 nmpms 3 nmpmsNR 4 nmpmsRB 5 nmpmsRC 6 nnodes 7 nelems 8 horiz 9 vert
//        10 depth 11 mpmApproach 12 useGimp 13 skipPostExtrapolation 14 XPICOrder 15 usingFMPM
//        16 mstep 17 nmat 18 maxShapeNodes 19 incrementalDefGradTerms 20 adiabatic 21 conduction active
//        22 numPatches 23 hasGravity 24 useDamping 25 usePDamping
// dbls:  0 xmin 1 ymin 2 zmin 3 gridx 4 gridy 5 gridz 6 timestep 7 strainTimestepFirst
//        8 strainTimestepLast 9 fractionUSF 10 mtime 11 damping(t) 12 pdamping(t) 13 gx 14 gy 15 gz
//        16 rcrit 17 thickness(2D) 18 maxtime
void ref_info(int *iv, double *dv)
{
    iv[0] = fmobj->np; iv[1] = fmobj->IsThreeD() ? 1 : 0;
    iv[2] = nmpms; iv[3] = nmpmsNR; iv[4] = nmpmsRB; iv[5] = nmpmsRC;
    iv[6] = nnodes; iv[7] = nelems;
    iv[8] = mpmgrid.horiz; iv[9] = mpmgrid.vert; iv[10] = mpmgrid.depth;
    iv[11] = fmobj->mpmApproach; iv[12] = ElementBase::useGimp;
    iv[13] = fmobj->skipPostExtrapolation ? 1 : 0;
    iv[14] = bodyFrc.GetXPICOrder(); iv[15] = bodyFrc.UsingFMPM() ? 1 : 0;
    iv[16] = fmobj->mstep; iv[17] = nmat; iv[18] = maxShapeNodes;
    iv[19] = MaterialBase::incrementalDefGradTerms;
    iv[20] = ConductionTask::adiabatic ? 1 : 0; iv[21] = ConductionTask::active ? 1 : 0;
    iv[22] = fmobj->GetTotalNumberOfPatches();
    iv[23] = bodyFrc.gravity ? 1 : 0; iv[24] = bodyFrc.useDamping ? 1 : 0; iv[25] = bodyFrc.usePDamping ? 1 : 0;
    dv[0] = mpmgrid.xmin; dv[1] = mpmgrid.ymin; dv[2] = mpmgrid.zmin;
    dv[3] = mpmgrid.grid.x; dv[4] = mpmgrid.grid.y; dv[5] = mpmgrid.grid.z;
    dv[6] = timestep; dv[7] = strainTimestepFirst; dv[8] = strainTimestepLast; dv[9] = fractionUSF;
    dv[10] = mtime;
    dv[11] = bodyFrc.GetGridDamping(mtime); dv[12] = bodyFrc.GetParticleDamping(mtime);
    dv[13] = bodyFrc.gforce.x; dv[14] = bodyFrc.gforce.y; dv[15] = bodyFrc.gforce.z;
    dv[16] = ElementBase::rcrit;
    dv[17] = fmobj->IsThreeD() ? 0. : mpmgrid.GetThickness();
    dv[18] = fmobj->maxtime;
}

// node coordinates (nnodes x 3), node order = reference numbering 1..nnodes
void ref_get_node_coords(double *xyz)
{
    for (int i = 1; i <= nnodes; i++) {
        xyz[3 * (i - 1)] = nd[i]->x; xyz[3 * (i - 1) + 1] = nd[i]->y; xyz[3 * (i - 1) + 2] = nd[i]->z;
    }
}

// element extents as the reference stores them: xmin,xmax,ymin,ymax,zmin,zmax (nelems x 6)
void ref_get_element_extents(double *ext)
{
    for (int e = 0; e < nelems; e++) {
        const ElementBase *el = theElements[e];
        ext[6 * e] = el->xmin; ext[6 * e + 1] = el->xmax; ext[6 * e + 2] = el->ymin; ext[6 * e + 3] = el->ymax;
        double zmn = 0., zmx = 0.;
        if (fmobj->IsThreeD()) { zmn = nd[el->nodes[0]]->z; zmx = nd[el->nodes[4]]->z; }
        ext[6 * e + 4] = zmn; ext[6 * e + 5] = zmx;
    }
}

// particle state, SoA, each array nmpms long (pass NULL to skip one)
// vec3 arrays are [3][nmpms]; tensors [6][nmpms] in order xx,yy,zz,yz,xz,xy; wrot [3][nmpms] xy,xz,yz
void ref_get_particles(double *pos, double *vel, double *mp, double *lp, int *inElem, int *matnum,
                       double *sp, double *pressure, double *ep, double *wrot, double *eplast,
                       double *energies /* [6][n]: work,res,heat,entropy,plast,prevTemp */,
                       double *hist /* [nh][n] */, int nh, double *pFext, double *origpos,
                       int *crossings, double *ncpos, double *acc)
{
    int n = nmpms;
    for (int p = 0; p < n; p++) {
        MPMBase *m = mpm[p];
        if (pos) { pos[p] = m->pos.x; pos[n + p] = m->pos.y; pos[2 * n + p] = m->pos.z; }
        if (vel) { vel[p] = m->vel.x; vel[n + p] = m->vel.y; vel[2 * n + p] = m->vel.z; }
        if (mp) mp[p] = m->mp;
        if (lp) { lp[p] = m->mpm_lp.x; lp[n + p] = m->mpm_lp.y; lp[2 * n + p] = m->mpm_lp.z; }
        if (inElem) inElem[p] = m->inElem;
        if (matnum) matnum[p] = m->matnum;
        if (sp) { sp[p] = m->sp.xx; sp[n + p] = m->sp.yy; sp[2 * n + p] = m->sp.zz;
                  sp[3 * n + p] = m->sp.yz; sp[4 * n + p] = m->sp.xz; sp[5 * n + p] = m->sp.xy; }
        if (pressure) pressure[p] = m->pressure;
        if (ep) { ep[p] = m->ep.xx; ep[n + p] = m->ep.yy; ep[2 * n + p] = m->ep.zz;
                  ep[3 * n + p] = m->ep.yz; ep[4 * n + p] = m->ep.xz; ep[5 * n + p] = m->ep.xy; }
        if (wrot) { wrot[p] = m->wrot.xy; wrot[n + p] = m->wrot.xz; wrot[2 * n + p] = m->wrot.yz; }
        if (eplast) { eplast[p] = m->eplast.xx; eplast[n + p] = m->eplast.yy; eplast[2 * n + p] = m->eplast.zz;
                      eplast[3 * n + p] = m->eplast.yz; eplast[4 * n + p] = m->eplast.xz; eplast[5 * n + p] = m->eplast.xy; }
        if (energies) { energies[p] = m->workEnergy; energies[n + p] = m->resEnergy; energies[2 * n + p] = m->heatEnergy;
                        energies[3 * n + p] = m->entropy; energies[4 * n + p] = m->plastEnergy;
                        energies[5 * n + p] = m->pPreviousTemperature; }
        if (hist) {
            const MaterialBase *mat = theMaterials[m->MatID()];
            int have = mat->NumberOfHistoryDoubles();
            for (int h = 0; h < nh; h++)
                hist[h * n + p] = (h < have && m->matData != NULL) ? ((double *)m->matData)[h] : 0.;
        }
        if (pFext) { pFext[p] = m->pFext.x; pFext[n + p] = m->pFext.y; pFext[2 * n + p] = m->pFext.z; }
        if (origpos) { origpos[p] = m->origpos.x; origpos[n + p] = m->origpos.y; origpos[2 * n + p] = m->origpos.z; }
        if (crossings) crossings[p] = m->elementCrossings;
        if (ncpos) { ncpos[p] = m->ncpos.x; ncpos[n + p] = m->ncpos.y; ncpos[2 * n + p] = m->ncpos.z; }
        if (acc) { acc[p] = m->acc.x; acc[n + p] = m->acc.y; acc[2 * n + p] = m->acc.z; }
    }
}

// node accumulators of the single velocity field cvf[0]->mvf[0]; arrays nnodes long, vec3 = [3][nnodes]
void ref_get_nodes(int *numberPoints, double *mass, double *pk, double *ftot, double *vk0, double *pkcopy,
                   int *fixedDirection)
{
    int n = nnodes;
    for (int i = 1; i <= n; i++) {
        int k = i - 1;
        MatVelocityField *mvf = NULL;
        if (nd[i]->cvf != NULL && nd[i]->cvf[0] != NULL && nd[i]->cvf[0]->mvf != NULL) mvf = nd[i]->cvf[0]->mvf[0];
        if (fixedDirection) fixedDirection[k] = nd[i]->fixedDirection;
        if (mvf == NULL) {
            if (numberPoints) numberPoints[k] = 0;
            if (mass) mass[k] = 0.;
            for (int c = 0; c < 3; c++) {
                if (pk) pk[c * n + k] = 0.;
                if (ftot) ftot[c * n + k] = 0.;
                if (vk0) vk0[c * n + k] = 0.;
                if (pkcopy) pkcopy[c * n + k] = 0.;
            }
            continue;
        }
        if (numberPoints) numberPoints[k] = mvf->numberPoints;
        if (mass) mass[k] = mvf->mass;
        if (pk) { pk[k] = mvf->pk.x; pk[n + k] = mvf->pk.y; pk[2 * n + k] = mvf->pk.z; }
        Vector f = mvf->GetFtot();
        if (ftot) { ftot[k] = f.x; ftot[n + k] = f.y; ftot[2 * n + k] = f.z; }
        if (vk0) { vk0[k] = mvf->vk[0].x; vk0[n + k] = mvf->vk[0].y; vk0[2 * n + k] = mvf->vk[0].z; }
        if (pkcopy) { Vector &c = mvf->vk[MatVelocityField::pkCopy];
                      pkcopy[k] = c.x; pkcopy[n + k] = c.y; pkcopy[2 * n + k] = c.z; }
    }
}

// ---- multimaterial mode: one set of node accumulators per material velocity field cvf[0]->mvf[f] -----------------------------
int ref_num_fields(void) { return fmobj->multiMaterialMode ? maxMaterialFields : 1; }

// field f of every node: the accumulators of ref_get_nodes plus the contact extrapolations (volume, volume gradient, displacement)
void ref_get_nodes_field(int f, int *numberPoints, double *mass, double *pk, double *ftot, double *vk0, double *pkcopy,
                         double *cvolume, double *cgrad, double *cdisp)
{
    int n = nnodes;
    for (int i = 1; i <= n; i++) {
        int k = i - 1;
        MatVelocityField *mvf = NULL;
        if (nd[i]->cvf != NULL && nd[i]->cvf[0] != NULL && nd[i]->cvf[0]->mvf != NULL && f < maxMaterialFields) mvf = nd[i]->cvf[0]->mvf[f];
        numberPoints[k] = 0; mass[k] = 0.; cvolume[k] = 0.;
        for (int c = 0; c < 3; c++) { pk[c * n + k] = 0.; ftot[c * n + k] = 0.; vk0[c * n + k] = 0.; pkcopy[c * n + k] = 0.; cgrad[c * n + k] = 0.; cdisp[c * n + k] = 0.; }
        if (mvf == NULL) continue;
        numberPoints[k] = mvf->numberPoints;
        mass[k] = mvf->mass;
        pk[k] = mvf->pk.x; pk[n + k] = mvf->pk.y; pk[2 * n + k] = mvf->pk.z;
        Vector ft = mvf->GetFtot();
        ftot[k] = ft.x; ftot[n + k] = ft.y; ftot[2 * n + k] = ft.z;
        vk0[k] = mvf->vk[0].x; vk0[n + k] = mvf->vk[0].y; vk0[2 * n + k] = mvf->vk[0].z;
        Vector &c = mvf->vk[MatVelocityField::pkCopy];
        pkcopy[k] = c.x; pkcopy[n + k] = c.y; pkcopy[2 * n + k] = c.z;
        if (mvf->contactInfo != NULL) {
            cvolume[k] = mvf->contactInfo->cvolume;
            if (mpmgrid.volumeGradientIndex >= 0) { Vector &v = mvf->contactInfo->terms[mpmgrid.volumeGradientIndex]; cgrad[k] = v.x; cgrad[n + k] = v.y; cgrad[2 * n + k] = v.z; }
            // the extrapolation contact detection uses: displacements (contactByDisplacements) or positions
            int di = mpmgrid.contactByDisplacements ? mpmgrid.displacementIndex : mpmgrid.positionIndex;
            if (di >= 0) { Vector &v = mvf->contactInfo->terms[di]; cdisp[k] = v.x; cdisp[n + k] = v.y; cdisp[2 * n + k] = v.z; }
        }
    }
}

// multimaterial settings: out[0] normal method, out[1] contactByDisplacements, out[2] number of materials; field[m] = velocity field of
// material m (-1 unused); law[(i*nmat+j)*4 ..]: contact law of the pair (kind: 0 ignore, 1 stick, 2 frictionless, 3 frictional, -1 other;
// friction coefficient; static coefficient; 0)
int ref_multimaterial_on(void) { return fmobj->multiMaterialMode ? 1 : 0; }

void ref_get_multimaterial(int *out, int *field, double *law, double *normal)
{
    out[0] = mpmgrid.materialNormalMethod; out[1] = mpmgrid.contactByDisplacements ? 1 : 0; out[2] = nmat;
    out[3] = fmobj->multiMaterialMode ? 1 : 0;
    normal[0] = mpmgrid.contactNormal.x; normal[1] = mpmgrid.contactNormal.y; normal[2] = mpmgrid.contactNormal.z;
    normal[3] = mpmgrid.positionCutoff;
    normal[4] = mpmgrid.rigidGradientBias;          // already squared by MeshInfo::MaterialOutput
    for (int i = 0; i < nmat; i++) field[i] = theMaterials[i]->GetField();
    for (int i = 0; i < nmat; i++)
        for (int j = 0; j < nmat; j++) {
            double *q = law + ((size_t)i * nmat + j) * 4;
            q[0] = -1.; q[1] = 0.; q[2] = 0.; q[3] = 0.;
            if (!fmobj->multiMaterialMode || field[i] < 0 || field[j] < 0 || i == j) continue;
            ContactLaw *cl = mpmgrid.GetMaterialContactLaw(field[i], field[j]);
            if (cl == NULL) continue;
            if (cl->IgnoreContact()) q[0] = 0.;
            else if (cl->IsImperfectInterface()) q[0] = -1.;
            else {
                CoulombFriction *cf = dynamic_cast<CoulombFriction *>(cl);
                if (cf == NULL || strcmp(cl->MaterialType(), "Coulomb Friction") != 0) continue;
                q[0] = cf->IsStick() ? 1. : (cf->IsFrictionless() ? 2. : 3.);
                q[1] = cf->frictionCoeff; q[2] = cf->frictionCoeffStatic;
            }
        }
}

// ---- conduction (the first transport task): settings, particle temperatures, nodal transport field ------------------------
// out[0] ConductionTask::active, [1] adiabatic, [2] number of temperature BCs, [3] number of heat-flux BCs, [4] material contact heating;
// kcond[m] = conductivity / rho (MaterialBase::kCond after VerifyAndLoadProperties)
void ref_get_conduction(int *out, double *kcond)
{
    out[0] = ConductionTask::active ? 1 : 0; out[1] = ConductionTask::adiabatic ? 1 : 0;
    int nb = 0; for (BoundaryCondition *b = (BoundaryCondition *)firstTempBC; b != NULL; b = (BoundaryCondition *)b->GetNextObject()) nb++;
    out[2] = nb;
    nb = 0; for (BoundaryCondition *b = (BoundaryCondition *)firstHeatFluxPt; b != NULL; b = (BoundaryCondition *)b->GetNextObject()) nb++;
    out[3] = nb;
    out[4] = ConductionTask::matContactHeating ? 1 : 0;
    for (int i = 0; i < nmat; i++) kcond[i] = theMaterials[i]->kCond;
}

int ref_conduction_on(void) { return ConductionTask::active ? 1 : 0; }

// nodal temperature BCs in list order: node (1-based, 0 when not active at mtime) and value at mtime
int ref_get_temp_bcs(int *node, double *value)
{
    int n = 0;
    for (NodalTempBC *b = firstTempBC; b != NULL; b = (NodalTempBC *)b->GetNextObject()) {
        if (node) { node[n] = b->GetNodeNum(); value[n] = b->GetNodeNum(mtime) != 0 ? b->BCValue(mtime) : 0.; }
        n++;
    }
    return n;
}

// pTemperature [n] and the particle's temperature gradient of the step [3][n]
void ref_get_temperatures(double *T, double *grad)
{
    const int n = nmpms;
    for (int p = 0; p < n; p++) {
        T[p] = mpm[p]->pTemperature;
        for (int c = 0; c < 3; c++) grad[c * n + p] = 0.;
        if (mpm[p]->pTemp != NULL && p < nmpmsNR) { grad[p] = mpm[p]->pTemp[0]; grad[n + p] = mpm[p]->pTemp[1]; if (fmobj->IsThreeD()) grad[2 * n + p] = mpm[p]->pTemp[2]; }
    }
}

// NodalPoint::gCond of every node
void ref_get_node_transport(double *gT, double *gVCT, double *gQ)
{
    for (int i = 1; i <= nnodes; i++) { gT[i - 1] = nd[i]->gCond.gTValue; gVCT[i - 1] = nd[i]->gCond.gVCT; gQ[i - 1] = nd[i]->gCond.gQ; }
}

// grid velocity BC list, in the reference's list order
int ref_num_velbcs(void)
{
    int n = 0;
    for (NodalVelBC *bc = firstVelocityBC; bc != NULL; bc = (NodalVelBC *)bc->GetNextObject()) n++;
    return n;
}

// per BC: node (1-based), dir bits, style, norm[3], value, ftime, offset, currentValue
void ref_get_velbcs(int *node, int *dir, int *style, double *norm, double *value, double *ftime,
                    double *offset, double *currentValue)
{
    int k = 0;
    for (NodalVelBC *bc = firstVelocityBC; bc != NULL; bc = (NodalVelBC *)bc->GetNextObject(), k++) {
        node[k] = bc->nodeNum; dir[k] = bc->dir; style[k] = bc->style;
        norm[3 * k] = bc->norm.x; norm[3 * k + 1] = bc->norm.y; norm[3 * k + 2] = bc->norm.z;
        value[k] = bc->value; ftime[k] = bc->ftime; offset[k] = bc->offset;
        currentValue[k] = bc->currentValue;
    }
}

// particle traction BCs (MatPtTractionBC list): 1-based particle, face, direction, style, BCValue at the current time
int ref_num_tractions(void)
{
    int k = 0;
    for (MatPtLoadBC *bc = firstTractionPt; bc != NULL; bc = (MatPtLoadBC *)bc->GetNextObject()) k++;
    return k;
}
void ref_get_tractions(int *particle, int *face, int *direction, int *style, double *value)
{
    int k = 0;
    for (MatPtLoadBC *b = firstTractionPt; b != NULL; b = (MatPtLoadBC *)b->GetNextObject(), k++) {
        MatPtTractionBC *bc = (MatPtTractionBC *)b;
        particle[k] = bc->ptNum; face[k] = bc->face; direction[k] = bc->direction; style[k] = bc->style; value[k] = bc->BCValue(mtime);
    }
}

// particle heat-flux BCs (MatPtHeatFluxBC list): 1-based particle, face, direction (1 external, 2 coupled), style, BCValue now
int ref_num_heat_fluxes(void)
{
    int k = 0;
    for (MatPtLoadBC *bc = firstHeatFluxPt; bc != NULL; bc = (MatPtLoadBC *)bc->GetNextObject()) k++;
    return k;
}
void ref_get_heat_fluxes(int *particle, int *face, int *direction, int *style, double *value)
{
    int k = 0;
    for (MatPtLoadBC *b = firstHeatFluxPt; b != NULL; b = (MatPtLoadBC *)b->GetNextObject(), k++) {
        MatPtHeatFluxBC *bc = (MatPtHeatFluxBC *)b;
        particle[k] = bc->ptNum; face[k] = bc->face; direction[k] = bc->direction; style[k] = bc->style; value[k] = bc->BCValue(mtime);
    }
}

// per BC: bcID (BoundaryCondition::GetID; the "id" attribute of the BC's XML block, the material number for rigid-particle BCs)
void ref_get_velbc_ids(int *ids)
{
    int k = 0;
    for (NodalVelBC *bc = firstVelocityBC; bc != NULL; bc = (NodalVelBC *)bc->GetNextObject(), k++) ids[k] = bc->GetID();
}

// what the "reactionx/y/z" global quantities read (GlobalQuantity.cpp:971-986): the summed freaction of the BCs with ID ids[k], all when 0
void ref_reaction_forces(int nids, const int *ids, double *out)
{
    for (int k = 0; k < nids; k++) {
        Vector f = NodalVelBC::TotalReactionForce(ids[k]);
        out[3 * k] = f.x; out[3 * k + 1] = f.y; out[3 * k + 2] = f.z;
    }
}

// per BC: NodalVelBC::reflectedNode (1-based node across a symmetry plane, -1 none) and reflectRatio
void ref_get_velbc_reflections(int *reflected, double *ratio)
{
    int k = 0;
    for (NodalVelBC *bc = firstVelocityBC; bc != NULL; bc = (NodalVelBC *)bc->GetNextObject(), k++) {
        reflected[k] = bc->reflectedNode; ratio[k] = bc->reflectRatio;
    }
}

// material facts. ids[i] = MaterialID(); params: 32 doubles per material, meaning by type
//  all:   0 rho  1 heatCapacity(Cv)  2 field  3 damping-or-(-1)  4 rigid flag
//  iso(1):      8 E 9 nu 10 G 11 CTE3 12 gamma0 13 useLargeRotation  14.. C11 C12 C44 (specific, /rho) from pr
//  neo(28):     8 G 9 K 10 Lame 11 Gsp 12 Ksp 13 Lamesp 14 UofJOption 15 CTE1 16 gamma0(as used)
//  mooney(8):   8 G1 9 G2 10 K 11 G1sp 12 G2sp 13 Ksp 14 UofJOption 15 CTE1 16 gamma0 17 IdealRubber
//  isoplas(9):  8 E 9 nu 10 G 11 CTE3 12 gamma0 13 Gred 14 Kred 15 yield 16 Ep 17 yldred 18 Epred 19 alphaMax 20 yldredMin 21 beta
//               22 useLargeRotation 23 hardening law id (1 Linear, 2 Nonlinear, 6 Nonlinear2, 3 JohnsonCook, 4 SCGL)
//               SCGL: 24 beta 25 nhard 26 yldMaxred 27 GPpred 28 GTp (16 thermal.reference)
//               Nonlinear/Nonlinear2: 24 beta 25 npow (19 alphaMax);  JohnsonCook: 24 Bred 25 njc 26 Cjc 27 ep0jc 28 Djc 29 n2jc 30 Tmjc 31 mjc
//               16 thermal.reference
int ref_get_materials(int *ids, double *params)
{
    for (int i = 0; i < nmat; i++) {
        MaterialBase *m = theMaterials[i];
        double *q = params + 32 * i;
        for (int k = 0; k < 32; k++) q[k] = 0.;
        ids[i] = m->MaterialID();
        q[0] = m->rho; q[1] = m->heatCapacity; q[2] = m->GetField();
        q[3] = m->matUsePDamping ? m->matPdamping : -1.;
        q[4] = m->IsRigid() ? 1. : 0.;
        q[5] = m->artificialViscosity ? 1. : 0.; q[6] = m->avA1; q[7] = m->avA2;
        if (ids[i] == 1) {
            IsotropicMat *im = (IsotropicMat *)m;
            q[8] = im->E; q[9] = im->nu; q[10] = im->G; q[11] = im->CTE3; q[12] = im->gamma0;
            q[13] = im->useLargeRotation;
            q[14] = im->pr.C[0][0]; q[15] = im->pr.C[0][1]; q[16] = im->pr.C[3][3];
        }
        else if (ids[i] == 28) {
            Neohookean *nm = (Neohookean *)m;
            q[8] = nm->G; q[9] = nm->Kbulk; q[10] = nm->Lame; q[11] = nm->pr.Gsp; q[12] = nm->pr.Ksp;
            q[13] = nm->pr.Lamesp; q[14] = nm->UofJOption; q[15] = nm->CTE1; q[16] = nm->gamma0;
        }
        else if (ids[i] == 8) {
            Mooney *mm = (Mooney *)m;
            q[8] = mm->G1; q[9] = mm->G2; q[10] = mm->Kbulk; q[11] = mm->G1sp; q[12] = mm->G2sp; q[13] = mm->Ksp;
            q[14] = mm->UofJOption; q[15] = mm->CTE1; q[16] = mm->gamma0; q[17] = mm->rubber ? 1. : 0.;
        }
        else if (ids[i] == 11) {
            RigidMaterial *rm = (RigidMaterial *)m;
            q[8] = rm->setDirection; q[9] = rm->mirrored;
            q[10] = (rm->function != NULL || rm->function2 != NULL || rm->function3 != NULL) ? 1. : 0.;
            q[11] = (rm->setTemperature ? 1. : 0.) + (rm->setConcentration ? 2. : 0.);
            q[12] = rm->Vfunction != NULL ? 1. : 0.;
        }
        else if (ids[i] == 9) {
            IsoPlasticity *pm = (IsoPlasticity *)m;
            q[8] = pm->E; q[9] = pm->nu; q[10] = pm->G; q[11] = pm->CTE3; q[12] = pm->gamma0;
            q[13] = pm->pr.Gred; q[14] = pm->pr.Kred; q[22] = pm->useLargeRotation;
            HardeningLawBase *h = pm->plasticLaw;
            if (h != NULL) { q[15] = h->yield; q[17] = h->yldred;
                             LinearHardening *lh = dynamic_cast<LinearHardening *>(h);
                             q[20] = h->yldredMin;
                             if (lh != NULL) { q[16] = lh->Ep; q[18] = lh->Epred; q[19] = lh->alphaMax; q[21] = lh->beta; q[23] = 1.; }
                             Nonlinear2Hardening *n2h = dynamic_cast<Nonlinear2Hardening *>(h);
                             NonlinearHardening *nh = dynamic_cast<NonlinearHardening *>(h);
                             JohnsonCook *jc = dynamic_cast<JohnsonCook *>(h);
                             if (nh != NULL) { q[23] = n2h != NULL ? 6. : 2.; q[24] = nh->beta; q[25] = nh->npow; q[19] = nh->alphaMax; }
                             SCGLHardening *sc = h->lawID == SCGLHARDENING_ID ? (SCGLHardening *)h : NULL;     // (SLMaterial, id 5, derives from it: not exported)
                             if (sc != NULL) { q[23] = 4.; q[24] = sc->beta; q[25] = sc->nhard; q[26] = sc->yldMaxred; q[27] = sc->GPpred; q[28] = sc->GTp;
                                               q[16] = thermal.reference; }
                             if (jc != NULL) { q[23] = 3.; q[24] = jc->Bred; q[25] = jc->njc; q[26] = jc->Cjc; q[27] = jc->ep0jc; q[28] = jc->Djc;
                                               q[29] = jc->n2jc; q[30] = jc->Tmjc; q[31] = jc->mjc; q[16] = thermal.reference; } }
        }
    }
    return nmat;
}

// Overwrite particle positions / velocities (arrays [3][nmpms]) before stepping, so that tests and the
// bench can give the reference exactly the (jittered, non-lattice) state they give the GPU path.
// Each particle's element is re-found with the reference's own ResetElementsTask::ResetElement.
int ref_set_particles(const double *pos, const double *vel)
{
    int n = nmpms, bad = 0;
    for (int p = 0; p < n; p++) {
        MPMBase *m = mpm[p];
        if (pos) {
            Vector x = MakeVector(pos[p], pos[n + p], pos[2 * n + p]);
            m->SetPosition(&x);
            m->SetOrigin(&x);
            int status = ResetElementsTask::ResetElement(m);
            if (status != SAME_ELEMENT && status != NEW_ELEMENT) bad++;
            m->SetElementCrossings(0);
        }
        if (vel) {
            Vector v = MakeVector(vel[p], vel[n + p], vel[2 * n + p]);
            m->SetVelocity(&v);
        }
    }
    return bad;
}

// The hardening law of IsoPlasticity material `mat` (0-based) evaluated alone (for tests of other implementations of the same law):
// out = {GetYield, GetKPrime, GetK2Prime(fnp1), GetYieldIncrement}.  particle 0 supplies the temperature for Johnson-Cook.
int ref_hardening_terms(int mat, double alpint, double dalpha, double delTime, double fnp1, double *out)
{
    if (mat < 0 || mat >= nmat || theMaterials[mat]->MaterialID() != 9) return -1;
    IsoPlasticity *pm = (IsoPlasticity *)theMaterials[mat];
    HardeningLawBase *h = pm->plasticLaw;
    HardeningAlpha a;
    a.alpint = alpint; a.dalpha = dalpha;
    char buffer[256];
    void *props = h->GetCopyOfHardeningProps(mpm[0], fmobj->np, (void *)buffer, 0);
    h->GetShearRatio(mpm[0], mpm[0]->GetPressure(), 1., props, 0);          // (fills SCGLProperties::Gratio, as IsoPlasticity::GetCopyOfMechanicalProps does)
    out[0] = h->GetYield(mpm[0], fmobj->np, delTime, &a, props);
    out[1] = h->GetKPrime(mpm[0], fmobj->np, delTime, &a, props);
    out[2] = h->GetK2Prime(mpm[0], fnp1, delTime, &a, props);
    out[3] = h->GetYieldIncrement(mpm[0], fmobj->np, delTime, &a, props);
    return 0;
}

// The reference's own MPMConstitutiveLaw applied to every non-rigid particle with a caller-given du[p][9] (row-major): the law alone,
// on the state the particles are in.  Lets tests compare another implementation of a law with the reference at the law level.
static double g_law_dT = 0.;        // ResidualStrains::dT handed to the laws by ref_constitutive_law_all
void ref_set_law_dT(double dT) { g_law_dT = dT; }

int ref_constitutive_law_all(const double *du, double delTime)
{
    static char *matBuf = NULL, *altBuf = NULL;
    if (matBuf == NULL) { matBuf = new char[MaterialBase::maxPropertyBufferSize + 64]; altBuf = new char[MaterialBase::maxAltBufferSize + 64]; }
    try {
        for (int p = 0; p < nmpmsNR; p++) {
            MPMBase *mptr = mpm[p];
            const MaterialBase *matRef = theMaterials[mptr->MatID()];
            void *props = matRef->GetCopyOfMechanicalProps(mptr, fmobj->np, (void *)matBuf, (void *)altBuf, 0);
            const double *d = du + (size_t)9 * p;
            ResidualStrains res;
            res.dT = g_law_dT; res.dC = 0.; res.doopse = 0.;
            if (fmobj->IsThreeD()) {
                Matrix3 dm(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8]);
                matRef->MPMConstitutiveLaw(mptr, dm, delTime, fmobj->np, props, &res, 0);
            } else {
                Matrix3 dm(d[0], d[1], d[3], d[4], d[8]);
                matRef->MPMConstitutiveLaw(mptr, dm, delTime, fmobj->np, props, &res, 0);
            }
        }
    } catch (CommonException &err) { set_err(err.Message()); return -1; }
    return 0;
}

void ref_close(void)
{
    if (g_coutbuf) { std::cout.flush(); std::cout.rdbuf(g_coutbuf); g_coutbuf = NULL; }
    if (g_log.is_open()) g_log.close();
}

} // extern "C"
