/*
 * mpm_oracle.c -- TEST INFRASTRUCTURE.  Plain-C, single-thread restatement of the reference NairnMPM
 * time step for the hot path (tasks 1-9, 11 + IsotropicMat), written to follow the REFERENCE's own
 * loop structure (per-particle candidate tables, serial node accumulation, serial BC list), not the
 * CUDA design.  It is the CPU checker ("port") for tests/ and the fallback cpu_baseline of bench.py.
 * It is never linked into libmpmgpu and never called by the product path.
 *
 * Parity of THIS file is pinned by tests/test_oracle_cpu.py against the golden dumps of the unmodified
 * reference (tests/golden/*.npz, produced by oracle/_ref = the reference's own sources compiled here).
 *
 * Data layout reuses the plain structs of include/mpmgpu.h (SoA host arrays); everything else is
 * independent of the product.
 *
 * Reference lines each function restates are cited at the function.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/mpmgpu.h"

typedef struct { double x, y, z; } Vec;

typedef struct {        /* MatVelocityField (Nodes/MatVelocityField.hpp:44-48) */
    int numberPoints;
    double mass;
    Vec pk, ftot, vk, pkCopy;
    Vec vprev, vnext;       /* vk[VSTARPREV_VEC], vk[VSTARNEXT_VEC] (vk = vk[VSTAR_VEC]) */
} NodeField;

typedef struct {
    mpmgpu_config cfg;
    int dim, horiz, vert, depth, nnodes, nelems, xplane, yplane, zplane;
    double *xpts, *ypts, *zpts;
    int nmat;
    mpmgpu_material *mats;
    int n, nNR;
    /* particle state (AoS-of-arrays, caller order) */
    double *pos, *vel, *mp, *lp, *ncpos, *sp, *pressure, *ep, *wrot, *eplast, *energies, *hist, *pfext, *acc;
    int *inElem, *matnum, *cross;
    NodeField *nd;      /* 0-based node index = reference node number - 1 */
    int nbc, nbcGrid, bcCap;        /* BCs [nbcGrid, nbc) are the ones rigid particles made this step */
    int *fixedDirection;            /* NodalPoint::fixedDirection bits x=1, y=2, z=4 */
    int *bcDir;
    int *bcMirror, *bcReflect;      /* rigid BCs: node spacing towards the mirrored side (0 = none), reflected node this step or -1 */
    double *gridRatio;              /* grid BCs next to a symmetry plane: bcReflect = the node across the plane (set once), reflectRatio */
    int *bcNode; double *bcNorm, *bcValue; int *bcActive, *bcSym;
    double dt, dtFirst, dtLast;
    long long mstep;
    /* CPDI domains (CPDIDomain, Common/System/DataTypes.hpp:110-115): up to 9 per particle */
    int ncorner;
    int *cpElem; double *cpXi, *cpWg, *cpWs;
    int cpdiLeftGrid;
} Oracle;

static Oracle *O = NULL;

#define P3(a, c, p) (O->a[(size_t)(c) * O->n + (p)])

/* ---- candidate tables: Elements/EightNodeIsoparamBrick.cpp:24-54, Common/Elements/FourNodeIsoparam.cpp:27-31 ---- */
static const double g3xii[64] = {-1, 1, 1, -1, -1, 1, 1, -1, -3, -1, 1, 3, 3, 3, 3, 1, -1, -3, -3, -3, -3, -1, 1, 3, 3, 3, 3, 1, -1, -3, -3, -3,
                                 -1, 1, 1, -1, -3, -1, 1, 3, 3, 3, 3, 1, -1, -3, -3, -3, -1, 1, 1, -1, -3, -1, 1, 3, 3, 3, 3, 1, -1, -3, -3, -3};
static const double g3eti[64] = {-1, -1, 1, 1, -1, -1, 1, 1, -3, -3, -3, -3, -1, 1, 3, 3, 3, 3, 1, -1, -3, -3, -3, -3, -1, 1, 3, 3, 3, 3, 1, -1,
                                 -1, -1, 1, 1, -3, -3, -3, -3, -1, 1, 3, 3, 3, 3, 1, -1, -1, -1, 1, 1, -3, -3, -3, -3, -1, 1, 3, 3, 3, 3, 1, -1};
static const double g3zti[64] = {-1, -1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                 -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, -3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3};
static const double gxii[16] = {-1, 1, 1, -1, -3, -1, 1, 3, 3, 3, 3, 1, -1, -3, -3, -3};
static const double geti[16] = {-1, -1, 1, 1, -3, -3, -3, -3, -1, 1, 3, 3, 3, 3, 1, -1};

static int off_of(double xi) { return (int)(0.5 * (xi + 1.)); }   /* -3->-1, -1->0, 1->1, 3->2 (tables :40-54) */

/* element (1-based) -> col,row,rank and first node (0-based): Read_MPM/Generators.cpp:1917-1929 */
static void elem_ijk(int inElem, int *i, int *j, int *k)
{
    int e0 = inElem - 1;
    *i = e0 % O->horiz; e0 /= O->horiz;
    *j = e0 % O->vert;
    *k = e0 / O->vert;
}

/* ---- shape functions: returns count, fills nds (0-based), fn, derivatives ------------------------------
 * uGIMP 3D: EightNodeIsoparamBrick.cpp:289-397;  2D: FourNodeIsoparam.cpp:431-519
 * Linear 3D: EightNodeIsoparamBrick.cpp:87-105;  2D: FourNodeIsoparam.cpp:189-210 */
static int shape(int p, int getDeriv, int *nds, double *fn, double *xd, double *yd, double *zd)
{
    int ei, ej, ek;
    elem_ijk(O->inElem[p], &ei, &ej, &ek);
    const int node0 = ek * O->zplane + ej * O->yplane + ei;
    const double xi = P3(ncpos, 0, p), eta = P3(ncpos, 1, p), zeta = P3(ncpos, 2, p);
    const double dx = O->xpts[ei + 1] - O->xpts[ei], dy = O->ypts[ej + 1] - O->ypts[ej];
    const double dz = O->dim == 3 ? O->zpts[ek + 1] - O->zpts[ek] : 1.;
    int i = 0;
    if (O->cfg.shape == MPMGPU_BSPLINE_CPDI) {
        /* B2CPDI: GetCPDIFunctions with the corners evaluated by SplineShapeFunction (MoreMPMElementBase.cpp:90-105,581-657) */
        int ind = 0;
        for (int c = 0; c < O->ncorner; c++) {
            const size_t q = (size_t)p * O->ncorner + c;
            int ci, cj, ck;
            elem_ijk(O->cpElem[q], &ci, &cj, &ck);
            const int cnode0 = ck * O->zplane + cj * O->yplane + ci;
            const double *cx = &O->cpXi[3 * q], *wg = &O->cpWg[3 * q];
            const double ws = O->cpWs[c];
            const int ncand = O->dim == 3 ? 64 : 16;
            for (int id = 0; id < ncand; id++) {
                const double xn = O->dim == 3 ? g3xii[id] : gxii[id], yn = O->dim == 3 ? g3eti[id] : geti[id], zn = O->dim == 3 ? g3zti[id] : 0.;
                double e1 = cx[0] - xn, t1 = fabs(e1); if (t1 >= 3.) continue;
                double e2 = cx[1] - yn, t2 = fabs(e2); if (t2 >= 3.) continue;
                double e3 = O->dim == 3 ? cx[2] - zn : 0., t3 = fabs(e3); if (O->dim == 3 && t3 >= 3.) continue;
                double sx = t1 <= 1.0 ? 0.25 * (3. - e1 * e1) : 0.125 * (3. - t1) * (3. - t1);
                double sy = t2 <= 1.0 ? 0.25 * (3. - e2 * e2) : 0.125 * (3. - t2) * (3. - t2);
                double sz = O->dim == 3 ? (t3 <= 1.0 ? 0.25 * (3. - e3 * e3) : 0.125 * (3. - t3) * (3. - t3)) : 1.;
                double cfn = O->dim == 3 ? sx * sy * sz : sx * sy;
                if (cfn < 1e-15) continue;
                int node = cnode0 + off_of(xn) + off_of(yn) * O->yplane + (O->dim == 3 ? off_of(zn) * O->zplane : 0);
                int look;
                for (look = ind - 1; look >= 0; look--) if (nds[look] == node) break;
                if (look >= 0) {
                    fn[look] += ws * cfn;
                    if (getDeriv) { xd[look] += wg[0] * cfn; yd[look] += wg[1] * cfn; zd[look] += wg[2] * cfn; }
                } else {
                    nds[ind] = node; fn[ind] = ws * cfn;
                    if (getDeriv) { xd[ind] = wg[0] * cfn; yd[ind] = wg[1] * cfn; zd[ind] = wg[2] * cfn; }
                    ind++;
                }
            }
        }
        return ind;
    }
    if (O->cfg.shape == MPMGPU_BSPLINE || O->cfg.shape == MPMGPU_BSPLINE_GIMP) {
        /* B2SPLINE: EightNodeIsoparamBrick::SplineShapeFunction :110-204, FourNodeIsoparam.cpp:217-294
         * B2GIMP:   EightNodeIsoparamBrick::BGimpShapeFunction :477-634 (z-gradient sign from xi.y, :586), FourNodeIsoparam.cpp:770-884 */
        const int spline = O->cfg.shape == MPMGPU_BSPLINE, ncand = O->dim == 3 ? 64 : 16;
        const double lpv[3] = {P3(lp, 0, p), P3(lp, 1, p), P3(lp, 2, p)};
        const double xv[3] = {xi, eta, zeta};
        const double inv_d[3] = {(spline ? 1. : 2.0) / dx, (spline ? 1. : 2.0) / dy, (spline ? 1. : 2.0) / dz};
        for (int id = 0; id < ncand; id++) {
            const double nn3[3] = {O->dim == 3 ? g3xii[id] : gxii[id], O->dim == 3 ? g3eti[id] : geti[id], O->dim == 3 ? g3zti[id] : 0.};
            double S[3] = {1., 1., 1.}, dS[3] = {0., 0., 0.};
            int skip = 0;
            for (int a = 0; a < O->dim && !skip; a++) {
                if (spline) {
                    double e = xv[a] - nn3[a], t = fabs(e);
                    if (t >= 3.) { skip = 1; break; }
                    if (t <= 1.0) { S[a] = 0.25 * (3. - e * e); dS[a] = -e; }
                    else { double arg = 3. - t; S[a] = 0.125 * arg * arg; dS[a] = e >= 0. ? 0.5 * (e - 3) : 0.5 * (e + 3); }
                } else {
                    const double l = lpv[a], b1 = 1. - l, b2 = 1. + l, b3 = 3. - l, b4 = 3. + l, inv_size = 1. / (48. * l), oneTwelth = 1. / 12.;
                    double xp = fabs(xv[a] - nn3[a]);
                    if (xp >= b4) { skip = 1; break; }
                    if (xp < b1) { S[a] = O->dim == 3 ? (9. - l * l - 3 * xp * xp) * oneTwelth : (9. - l * l - 3. * xp * xp) * oneTwelth; dS[a] = -0.5 * xp; }
                    else if (xp < b2) {
                        double arg = xp - 1., lp2 = l * l;
                        if (O->dim == 3) { double lp3 = lp2 * l; S[a] = (9. * lp2 * arg + 3. * arg * arg * arg + 3. * l * (15. - xp * (6. + xp)) - lp3) * inv_size; }
                        else S[a] = (lp2 * (9. * arg - l) + 3. * arg * arg * arg + 3. * l * (15. - xp * (6. + xp))) * inv_size;
                        dS[a] = (3 * lp2 + 3. * arg * arg - 2. * l * (3. + xp)) * 3. * inv_size;
                    }
                    else if (xp <= b3) { double arg = xp - 3.; S[a] = (l * l + 3. * arg * arg) * 0.5 * oneTwelth; dS[a] = 0.25 * (xp - 3.); }
                    else { double arg = 3. + l - xp; S[a] = arg * arg * arg * inv_size; dS[a] = -arg * arg * 3. * inv_size; }
                    /* xsign, ysign and (3D) zsign = xi.y > node z */
                    const double ref = (a == 2) ? xv[1] : xv[a];
                    if (!(ref > nn3[a])) dS[a] = -dS[a];
                }
            }
            if (skip) continue;
            if (O->dim == 3) {
                fn[i] = S[0] * S[1] * S[2];
                if (getDeriv) { xd[i] = dS[0] * S[1] * S[2] * inv_d[0]; yd[i] = S[0] * dS[1] * S[2] * inv_d[1]; zd[i] = S[0] * S[1] * dS[2] * inv_d[2]; }
                nds[i] = node0 + off_of(nn3[0]) + off_of(nn3[1]) * O->yplane + off_of(nn3[2]) * O->zplane;
            } else {
                fn[i] = S[0] * S[1];
                if (getDeriv) { xd[i] = dS[0] * S[1] * inv_d[0]; yd[i] = S[0] * dS[1] * inv_d[1]; zd[i] = 0.; }
                nds[i] = node0 + off_of(nn3[0]) + off_of(nn3[1]) * O->yplane;
            }
            i++;
        }
        return i;
    }
    if (O->cfg.shape == MPMGPU_LINEAR_CPDI || O->cfg.shape == MPMGPU_QUADRATIC_CPDI) {
        /* ElementBase::GetCPDIFunctions, Elements/MoreMPMElementBase.cpp:581-657: merge by node in first-seen order */
        static const double xii[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, eti[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, zti[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        const int nn = O->dim == 3 ? 8 : 4;
        int ind = 0;
        for (int c = 0; c < O->ncorner; c++) {
            const size_t q = (size_t)p * O->ncorner + c;
            int ci, cj, ck;
            elem_ijk(O->cpElem[q], &ci, &cj, &ck);
            const int cnode0 = ck * O->zplane + cj * O->yplane + ci;
            const double *cx = &O->cpXi[3 * q], *wg = &O->cpWg[3 * q];
            const double ws = O->cpWs[c];
            for (int j = 0; j < nn; j++) {
                double t1 = 1. + xii[j] * cx[0], t2 = 1. + eti[j] * cx[1], t3 = 1. + zti[j] * cx[2];
                double cfn = O->dim == 3 ? 0.125 * t1 * t2 * t3 : 0.25 * t1 * t2;
                if (cfn < 1e-15) continue;
                int node = cnode0 + (xii[j] > 0) + (eti[j] > 0) * O->yplane + (O->dim == 3 ? (zti[j] > 0) * O->zplane : 0);
                int look;
                for (look = ind - 1; look >= 0; look--) if (nds[look] == node) break;
                if (look >= 0) {
                    fn[look] += ws * cfn;
                    if (getDeriv) { xd[look] += wg[0] * cfn; yd[look] += wg[1] * cfn; zd[look] += wg[2] * cfn; }
                } else {
                    nds[ind] = node; fn[ind] = ws * cfn;
                    if (getDeriv) { xd[ind] = wg[0] * cfn; yd[ind] = wg[1] * cfn; zd[ind] = wg[2] * cfn; }
                    ind++;
                }
            }
        }
        return ind;
    }
    if (O->cfg.shape == MPMGPU_POINT_GIMP) {
        static const double xii[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, eti[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, zti[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        const int nn = O->dim == 3 ? 8 : 4;
        for (i = 0; i < nn; i++) {
            double t1 = 1. + xii[i] * xi, t2 = 1. + eti[i] * eta, t3 = 1. + zti[i] * zeta;
            int xo = xii[i] > 0, yo = eti[i] > 0, zo = zti[i] > 0;
            if (O->dim == 3) {
                fn[i] = 0.125 * t1 * t2 * t3;
                if (getDeriv) { xd[i] = 0.25 * xii[i] * t2 * t3 / dx; yd[i] = 0.25 * eti[i] * t1 * t3 / dy; zd[i] = 0.25 * zti[i] * t1 * t2 / dz; }
                nds[i] = node0 + xo + yo * O->yplane + zo * O->zplane;
            } else {
                fn[i] = 0.25 * t1 * t2;
                if (getDeriv) { xd[i] = 0.5 * xii[i] * t2 / dx; yd[i] = 0.5 * eti[i] * t1 / dy; zd[i] = 0.; }
                nds[i] = node0 + xo + yo * O->yplane;
            }
        }
        return nn;
    }
    const double lpx = P3(lp, 0, p), lpy = P3(lp, 1, p), lpz = P3(lp, 2, p);
    const double q1x = 2. - lpx, q2x = 2. + lpx, q1y = 2. - lpy, q2y = 2. + lpy, q1z = 2. - lpz, q2z = 2. + lpz;
    const double isx = 1. / (4. * lpx), isy = 1. / (4. * lpy), isz = 1. / (4. * lpy);     /* :303 uses lp.y for z */
    const double inv_dx = 2.0 / dx, inv_dy = 2.0 / dy, inv_dz = 2.0 / dz;
    if (O->dim == 3) {
        for (int id = 0; id < 64; id++) {
            double xp = fabs(xi - g3xii[id]); if (xp >= q2x) continue;
            double yp = fabs(eta - g3eti[id]); if (yp >= q2y) continue;
            double zp = fabs(zeta - g3zti[id]); if (zp >= q2z) continue;
            double Sx, Sy, Sz, argx = 0., argy = 0., argz = 0.;
            if (xp < lpx) Sx = ((4. - lpx) * lpx - xp * xp) * isx; else if (xp <= q1x) Sx = 0.5 * (2. - xp); else { argx = (q2x - xp) * isx; Sx = 2. * lpx * argx * argx; }
            if (yp < lpy) Sy = ((4. - lpy) * lpy - yp * yp) * isy; else if (yp <= q1y) Sy = 0.5 * (2. - yp); else { argy = (q2y - yp) * isy; Sy = 2. * lpy * argy * argy; }
            if (zp < lpz) Sz = ((4. - lpz) * lpz - zp * zp) * isz; else if (zp <= q1z) Sz = 0.5 * (2. - zp); else { argz = (q2z - zp) * isz; Sz = 2. * lpz * argz * argz; }
            fn[i] = Sx * Sy * Sz;
            if (getDeriv) {
                double xs = xi > g3xii[id] ? 1. : -1., ys = eta > g3eti[id] ? 1. : -1., zs = zeta > g3zti[id] ? 1. : -1.;
                double dSx = xp < lpx ? -xp / (2. * lpx) : (xp <= q1x ? -0.5 : -argx);
                double dSy = yp < lpy ? -yp / (2. * lpy) : (yp <= q1y ? -0.5 : -argy);
                double dSz = zp < lpz ? -zp / (2. * lpz) : (zp <= q1z ? -0.5 : -argz);
                xd[i] = xs * dSx * Sy * Sz * inv_dx; yd[i] = ys * Sx * dSy * Sz * inv_dy; zd[i] = zs * Sx * Sy * dSz * inv_dz;
            }
            nds[i] = node0 + off_of(g3xii[id]) + off_of(g3eti[id]) * O->yplane + off_of(g3zti[id]) * O->zplane;
            i++;
        }
    } else {
        for (int id = 0; id < 16; id++) {
            double xp = fabs(xi - gxii[id]); if (xp >= q2x) continue;
            double yp = fabs(eta - geti[id]); if (yp >= q2y) continue;
            double Sx, Sy, argx = 0., argy = 0.;
            if (xp < lpx) Sx = ((4. - lpx) * lpx - xp * xp) * isx; else if (xp <= q1x) Sx = (2. - xp) / 2.; else { argx = (q2x - xp) * isx; Sx = 2. * lpx * argx * argx; }
            if (yp < lpy) Sy = ((4. - lpy) * lpy - yp * yp) * isy; else if (yp <= q1y) Sy = (2. - yp) / 2.; else { argy = (q2y - yp) * isy; Sy = 2. * lpy * argy * argy; }
            fn[i] = Sx * Sy;
            if (getDeriv) {
                double xs = xi > gxii[id] ? 1. : -1., ys = eta > geti[id] ? 1. : -1.;
                double dSx = xp < lpx ? -xp * isx * 2.0 : (xp <= q1x ? -0.5 : -argx);
                double dSy = yp < lpy ? -yp * isy * 2.0 : (yp <= q1y ? -0.5 : -argy);
                xd[i] = xs * dSx * Sy * inv_dx; yd[i] = ys * Sx * dSy * inv_dy; zd[i] = 0.;
            }
            nds[i] = node0 + off_of(gxii[id]) + off_of(geti[id]) * O->yplane;
            i++;
        }
    }
    return i;
}

static void get_F(int p, double F[3][3]);

/* ---- CPDI domains: MatPoint3D.cpp:413-492,538-621; MatPoint2D.cpp:423-477,519-640 --------------------------------- */
static int find_element(const double *x)        /* MeshInfo::FindElementFromPoint, MeshInfo.cpp:593-633; 0 = off grid */
{
    int col = (int)((x[0] - O->xpts[0]) / O->cfg.gridx), row = (int)((x[1] - O->ypts[0]) / O->cfg.gridy), zrow = 0;
    if (col < 0 || col >= O->horiz) { if (x[0] == O->xpts[0] + O->horiz * O->cfg.gridx) col = O->horiz - 1; else return 0; }
    if (row < 0 || row >= O->vert) { if (x[1] == O->ypts[0] + O->vert * O->cfg.gridy) row = O->vert - 1; else return 0; }
    if (O->dim == 3) {
        zrow = (int)((x[2] - O->zpts[0]) / O->cfg.gridz);
        if (zrow < 0 || zrow >= O->depth) { if (x[2] == O->zpts[0] + O->depth * O->cfg.gridz) zrow = O->depth - 1; else return 0; }
        return O->horiz * (zrow * O->vert + row) + col + 1;
    }
    return row * O->horiz + col + 1;
}

static void corner_into(int p, int c, const double *x, int forceElem)
{
    const size_t q = (size_t)p * O->ncorner + c;
    int e = forceElem > 0 ? forceElem : find_element(x);
    if (e <= 0) { O->cpdiLeftGrid = p + 1; e = O->inElem[p]; }
    int i, j, k;
    elem_ijk(e, &i, &j, &k);
    O->cpElem[q] = e;
    O->cpXi[3 * q] = (2. * x[0] - O->xpts[i] - O->xpts[i + 1]) / (O->xpts[i + 1] - O->xpts[i]);
    O->cpXi[3 * q + 1] = (2. * x[1] - O->ypts[j] - O->ypts[j + 1]) / (O->ypts[j + 1] - O->ypts[j]);
    O->cpXi[3 * q + 2] = O->dim == 3 ? (2. * x[2] - O->zpts[k] - O->zpts[k + 1]) / (O->zpts[k + 1] - O->zpts[k]) : 0.;
}

static void cpdi_nodes_and_weights(int p)
{
    int ei, ej, ek;
    elem_ijk(O->inElem[p], &ei, &ej, &ek);
    const double cx = O->xpts[ei + 1] - O->xpts[ei], cy = O->ypts[ej + 1] - O->ypts[ej], cz = O->dim == 3 ? O->zpts[ek + 1] - O->zpts[ek] : 1.;
    const double psx = cx * (0.5 * P3(lp, 0, p)), psy = cy * (0.5 * P3(lp, 1, p)), psz = cz * (0.5 * P3(lp, 2, p));
    double F[3][3];
    get_F(p, F);
    const double pos[3] = {P3(pos, 0, p), P3(pos, 1, p), O->dim == 3 ? P3(pos, 2, p) : 0.};
    const size_t q0 = (size_t)p * O->ncorner;
    if (O->dim == 3) {
        double r1[3] = {F[0][0] * psx, F[1][0] * psx, F[2][0] * psx}, r2[3] = {F[0][1] * psy, F[1][1] * psy, F[2][1] * psy};
        double r3[3] = {F[0][2] * psz, F[1][2] * psz, F[2][2] * psz};
        if (O->cfg.cpdi_rcrit >= 0.) {
            double rc = O->cfg.cpdi_rcrit * fmin(cx, fmin(cy, cz)), l[4][3];
            static const double sg[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
            int rescale = 0;
            for (int a = 0; a < 4; a++) {
                for (int d = 0; d < 3; d++) l[a][d] = sg[a][0] * r1[d] + sg[a][1] * r2[d] + r3[d];
                double mag = sqrt(l[a][0] * l[a][0] + l[a][1] * l[a][1] + l[a][2] * l[a][2]);
                if (mag > rc) { for (int d = 0; d < 3; d++) l[a][d] *= rc / mag; rescale = 1; }
            }
            if (rescale) for (int d = 0; d < 3; d++) {
                r1[d] = 0.25 * (l[0][d] + l[1][d] - l[2][d] - l[3][d]);
                r2[d] = 0.25 * (l[0][d] - l[1][d] + l[2][d] - l[3][d]);
                r3[d] = 0.25 * (l[0][d] + l[1][d] + l[2][d] + l[3][d]);
            }
        }
        static const double r1s[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, r2s[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, r3s[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        for (int i = 0; i < 8; i++) {
            double c[3];
            for (int d = 0; d < 3; d++) c[d] = pos[d] + r1s[i] * r1[d] + r2s[i] * r2[d] + r3s[i] * r3[d];
            corner_into(p, i, c, 0);
        }
        double Vp = 8. * (r1[0] * (r2[1] * r3[2] - r2[2] * r3[1]) + r1[1] * (r2[2] * r3[0] - r2[0] * r3[2]) + r1[2] * (r2[0] * r3[1] - r2[1] * r3[0]));
        Vp = 1. / Vp;
        /* wg_c = (1/Vp) * [ s1 (r2 x r3) + s2 (r3 x r1) + s3 (r1 x r2) ] written out in the reference (:551-594);
           evaluated here from the same closed forms */
        for (int i = 0; i < 8; i++) {
            double s1 = r1s[i], s2 = r2s[i], s3 = r3s[i];
            double *w = &O->cpWg[3 * (q0 + i)];
            /* expand exactly the reference polynomials: terms are +-(ra_b * rc_d) */
            w[0] = (-s3 * r1[2] * r2[1] + s3 * r1[1] * r2[2] + s2 * r1[2] * r3[1] - s1 * r2[2] * r3[1] - s2 * r1[1] * r3[2] + s1 * r2[1] * r3[2]) * Vp;
            w[1] = (s3 * r1[2] * r2[0] - s3 * r1[0] * r2[2] - s2 * r1[2] * r3[0] + s1 * r2[2] * r3[0] + s2 * r1[0] * r3[2] - s1 * r2[0] * r3[2]) * Vp;
            w[2] = (-s3 * r1[1] * r2[0] + s3 * r1[0] * r2[1] + s2 * r1[1] * r3[0] - s1 * r2[1] * r3[0] - s2 * r1[0] * r3[1] + s1 * r2[0] * r3[1]) * Vp;
        }
    } else {
        double r1[2] = {F[0][0] * psx, F[1][0] * psx}, r2[2] = {F[0][1] * psy, F[1][1] * psy};
        if (O->cfg.cpdi_rcrit >= 0.) {
            double rc = O->cfg.cpdi_rcrit * fmin(cx, cy);
            double la[2] = {r1[0] + r2[0], r1[1] + r2[1]}, lb[2] = {r1[0] - r2[0], r1[1] - r2[1]};
            double lam = sqrt(la[0] * la[0] + la[1] * la[1]), lbm = sqrt(lb[0] * lb[0] + lb[1] * lb[1]);
            int rescale = 0;
            if (lam > rc) { la[0] *= rc / lam; la[1] *= rc / lam; rescale = 1; }
            if (lbm > rc) { lb[0] *= rc / lbm; lb[1] *= rc / lbm; rescale = 1; }
            if (rescale) { r1[0] = 0.5 * (la[0] + lb[0]); r1[1] = 0.5 * (la[1] + lb[1]); r2[0] = 0.5 * (la[0] - lb[0]); r2[1] = 0.5 * (la[1] - lb[1]); }
        }
        static const double s1[9] = {-1, 1, 1, -1, 0, 1, 0, -1, 0}, s2[9] = {-1, -1, 1, 1, -1, 0, 1, 0, 0};
        for (int i = 0; i < O->ncorner; i++) {
            double c[3] = {pos[0], pos[1], 0.};
            if (s1[i] != 0.) { c[0] += s1[i] * r1[0]; c[1] += s1[i] * r1[1]; }
            if (s2[i] != 0.) { c[0] += s2[i] * r2[0]; c[1] += s2[i] * r2[1]; }
            corner_into(p, i, c, i == 8 ? O->inElem[p] : 0);
        }
        double Ap = 4. * (r1[0] * r2[1] - r1[1] * r2[0]);
        Ap = O->ncorner == 9 ? 1. / (3. * Ap) : 1. / Ap;
        double (*w)[3] = (double (*)[3]) & O->cpWg[3 * q0];
        w[0][0] = (r1[1] - r2[1]) * Ap; w[0][1] = (-r1[0] + r2[0]) * Ap;
        w[1][0] = (r1[1] + r2[1]) * Ap; w[1][1] = (-r1[0] - r2[0]) * Ap;
        w[2][0] = (-r1[1] + r2[1]) * Ap; w[2][1] = (r1[0] - r2[0]) * Ap;
        w[3][0] = (-r1[1] - r2[1]) * Ap; w[3][1] = (r1[0] + r2[0]) * Ap;
        if (O->ncorner == 9) {
            w[4][0] = 4. * r1[1] * Ap; w[4][1] = -4. * r1[0] * Ap; w[5][0] = 4. * r2[1] * Ap; w[5][1] = -4. * r2[0] * Ap;
            w[6][0] = -4. * r1[1] * Ap; w[6][1] = 4. * r1[0] * Ap; w[7][0] = -4. * r2[1] * Ap; w[7][1] = 4. * r2[0] * Ap;
            w[8][0] = 0.; w[8][1] = 0.;
        }
        for (int i = 0; i < O->ncorner; i++) w[i][2] = 0.;
    }
}

/* ---- task 1: InitializationTask.cpp:49-85; GetXiPos EightNodeIsoparamBrick.cpp:276-281 ------------------- */
static void task_initialization(void)
{
    for (int i = 0; i < O->nnodes; i++) memset(&O->nd[i], 0, sizeof(NodeField));       /* MatVelocityField::Zero :92-104 */
    for (int p = 0; p < O->n; p++) {
        int ei, ej, ek;
        elem_ijk(O->inElem[p], &ei, &ej, &ek);
        P3(ncpos, 0, p) = (2. * P3(pos, 0, p) - O->xpts[ei] - O->xpts[ei + 1]) / (O->xpts[ei + 1] - O->xpts[ei]);
        P3(ncpos, 1, p) = (2. * P3(pos, 1, p) - O->ypts[ej] - O->ypts[ej + 1]) / (O->ypts[ej + 1] - O->ypts[ej]);
        P3(ncpos, 2, p) = O->dim == 3 ? (2. * P3(pos, 2, p) - O->zpts[ek] - O->zpts[ek + 1]) / (O->zpts[ek + 1] - O->zpts[ek]) : 0.;
        if (O->ncorner > 0) cpdi_nodes_and_weights(p);      /* all particles, rigid ones too (InitializationTask.cpp:62-67) */
    }
}

/* ---- task 2: MassAndMomentumTask.cpp:62-98, NodalPointMPM.cpp:419-453 ------------------------------------ */
static void task_mass_and_momentum(void)
{
    int nds[64]; double fn[64];
    for (int p = 0; p < O->nNR; p++) {
        int nn = shape(p, 0, nds, fn, NULL, NULL, NULL);
        for (int i = 0; i < nn; i++) {
            NodeField *f = &O->nd[nds[i]];
            double fnmp = fn[i] * O->mp[p];
            f->pk.x += P3(vel, 0, p) * fnmp; f->pk.y += P3(vel, 1, p) * fnmp; f->pk.z += P3(vel, 2, p) * fnmp;
            f->numberPoints += 1;
            f->mass += fnmp;
        }
    }
}

/* ---- velocity BCs: NodalVelBC.cpp:321-400, MatVelocityField.cpp:490-575 ----------------------------------- */
enum { MASS_MOMENTUM_CALL, GRID_FORCES_CALL, UPDATE_MOMENTUM_CALL, UPDATE_STRAINS_LAST_CALL };

static void velocity_bc_loop(int pass)
{
    const double dt = O->dt;
    for (int b = 0; b < O->nbc; b++) {         /* zero pass over the whole list */
        if (!O->bcActive[b]) continue;
        NodeField *f = &O->nd[O->bcNode[b] - 1];
        if (f->numberPoints <= 0) continue;
        const double *n = &O->bcNorm[3 * b];
        if (pass == GRID_FORCES_CALL) {
            double dotf = f->ftot.x * n[0] + f->ftot.y * n[1] + f->ftot.z * n[2];
            double dotp = f->pk.x * n[0] + f->pk.y * n[1] + f->pk.z * n[2];
            double s = -dotf - dotp / dt;
            f->ftot.x += n[0] * s; f->ftot.y += n[1] * s; f->ftot.z += n[2] * s;
        } else {
            double dotn = f->pk.x * n[0] + f->pk.y * n[1] + f->pk.z * n[2];
            f->pk.x += n[0] * (-dotn); f->pk.y += n[1] * (-dotn); f->pk.z += n[2] * (-dotn);
            if (pass == UPDATE_MOMENTUM_CALL) { double s = -dotn / dt; f->ftot.x += n[0] * s; f->ftot.y += n[1] * s; f->ftot.z += n[2] * s; }
        }
    }
    for (int b = 0; b < O->nbc; b++) {         /* then the add pass */
        if (!O->bcActive[b]) continue;
        NodeField *f = &O->nd[O->bcNode[b] - 1];
        if (f->numberPoints <= 0) continue;
        const double *n = &O->bcNorm[3 * b];
        double vel = O->bcValue[b];
        if (b >= O->nbcGrid && O->bcReflect[b] >= 0) {      /* CrackVelocityFieldSingle::ReflectVelocityBC :133-143 */
            const NodeField *r = &O->nd[O->bcReflect[b] - 1];
            if (r->numberPoints <= 0) continue;
            vel = vel + 1. * (vel - (n[0] * r->pk.x + n[1] * r->pk.y + n[2] * r->pk.z) / r->mass);
        } else if (b < O->nbcGrid && O->bcReflect[b] >= 0) {
            /* symmetry-plane neighbour (Generators.cpp:2178-2190): NodalPoint::ReflectVelocityBC, NodalPointMPM.cpp:1865-1881 --
               a reflected node without particles gives the plain BC */
            const NodeField *r = &O->nd[O->bcReflect[b] - 1];
            if (r->numberPoints > 0)
                vel = vel + O->gridRatio[b] * (vel - (n[0] * r->pk.x + n[1] * r->pk.y + n[2] * r->pk.z) / r->mass);
        }
        if (pass == GRID_FORCES_CALL) {
            double s = f->mass * vel / dt;
            f->ftot.x += n[0] * s; f->ftot.y += n[1] * s; f->ftot.z += n[2] * s;
        } else {
            double pvel = f->mass * vel;
            f->pk.x += n[0] * pvel; f->pk.y += n[1] * pvel; f->pk.z += n[2] * pvel;
            if (pass == UPDATE_MOMENTUM_CALL) { double s = pvel / dt; f->ftot.x += n[0] * s; f->ftot.y += n[1] * s; f->ftot.z += n[2] * s; }
        }
    }
}

static void grid_velocity_conditions(int pass)
{
    if (O->nbc == 0) return;
    if (pass == MASS_MOMENTUM_CALL) {
        for (int b = 0; b < O->nbc; b++) {     /* ADJUST_COPIED_PK == 1: NodalVelBC.cpp:339-353 */
            NodeField *f = &O->nd[O->bcNode[b] - 1];
            if (f->numberPoints <= 0) continue;
            int sd = O->bcSym ? O->bcSym[b] : 0;
            if (sd & 32) f->pkCopy.x = 0.;
            if (sd & 64) f->pkCopy.y = 0.;
            if (sd & 128) f->pkCopy.z = 0.;
        }
        if (O->cfg.method == MPMGPU_USL) return;       /* no USF task: NodalVelBC.cpp:356 */
    }
    if (pass == UPDATE_MOMENTUM_CALL && O->cfg.xpic_order > 1) return;
    velocity_bc_loop(pass);
}

/* ---- ProjectRigidBCsTask.cpp:39-158 (SetRigidBCs :194-266, UnsetRigidBCs :166-191) ---------------------------
 * Rigid-BC particles (mpm[nmpmsRC..nmpms)) append a constant velocity BC to the list on every node of their
 * shape-function stencil, in each direction their material sets, unless that dof is already fixed. */
static void task_project_rigid_bcs(void)
{
    if (O->n == O->nNR) return;
    for (int b = O->nbcGrid; b < O->nbc; b++) {             /* BoundaryCondition::UnsetDirection */
        int *fd = &O->fixedDirection[O->bcNode[b] - 1];
        if (*fd & O->bcDir[b]) *fd ^= O->bcDir[b];
    }
    O->nbc = O->nbcGrid;
    int nds[64]; double fn[64];
    for (int p = O->nNR; p < O->n; p++) {
        const mpmgpu_material *mat = &O->mats[O->matnum[p] - 1];
        const int setDirection = (int)mat->p[8];
        int nn = shape(p, 0, nds, fn, NULL, NULL, NULL);
        for (int i = 0; i < nn; i++) {
            for (int d = 0; d < 3; d++) {
                const int type = 1 << d;
                if ((setDirection & type) != type) continue;            /* RigidMaterial::RigidDirection */
                if (O->fixedDirection[nds[i]] & type) continue;         /* :202 */
                if (O->nbc == O->bcCap) {
                    O->bcCap = O->bcCap ? 2 * O->bcCap : 1024;
                    O->bcNode = (int *)realloc(O->bcNode, O->bcCap * sizeof(int)); O->bcDir = (int *)realloc(O->bcDir, O->bcCap * sizeof(int));
                    O->bcActive = (int *)realloc(O->bcActive, O->bcCap * sizeof(int)); O->bcSym = (int *)realloc(O->bcSym, O->bcCap * sizeof(int));
                    O->bcMirror = (int *)realloc(O->bcMirror, O->bcCap * sizeof(int)); O->bcReflect = (int *)realloc(O->bcReflect, O->bcCap * sizeof(int));
                    O->bcNorm = (double *)realloc(O->bcNorm, 3 * (size_t)O->bcCap * sizeof(double));
                    O->bcValue = (double *)realloc(O->bcValue, O->bcCap * sizeof(double));
                }
                const int b = O->nbc++;
                O->bcNode[b] = nds[i] + 1; O->bcDir[b] = type; O->bcActive[b] = 1; O->bcSym[b] = 0;
                O->bcNorm[3 * b] = d == 0; O->bcNorm[3 * b + 1] = d == 1; O->bcNorm[3 * b + 2] = d == 2;
                O->bcValue[b] = P3(vel, d, p);                           /* CONSTANT_VALUE, ftime 0 */
                {   /* NodalVelBC::SetMirrorSpacing :264-283 */
                    const int mirrored = (int)mat->p[9], plane = d == 0 ? O->xplane : (d == 1 ? O->yplane : O->zplane);
                    O->bcMirror[b] = mirrored == 0 ? 0 : (mirrored < 0 ? plane : -plane);
                    O->bcReflect[b] = -1;
                }
                O->fixedDirection[nds[i]] |= type;
            }
        }
    }
}

/* ---- task 3: PostExtrapolationTask.cpp:62-165 -------------------------------------------------------------- */
static void task_post_extrapolation(void)
{
    for (int i = 0; i < O->nnodes; i++) O->nd[i].pkCopy = O->nd[i].pk;         /* MatVelocityField.cpp:158-164 */
    for (int b = O->nbcGrid; b < O->nbc; b++) {        /* NodalVelBC::SetMirroredVelBC :214-244 (PostExtrapolationTask.cpp:146-151) */
        const int s = O->bcMirror[b], i = O->bcNode[b], dir = O->bcDir[b];
        O->bcReflect[b] = -1;
        if (s == 0) continue;
        const int neighbor = i + s;
        if (O->nd[i - 1].numberPoints > 0 && neighbor > 0 && neighbor <= O->nnodes && (O->fixedDirection[neighbor - 1] & dir)) {
            const int mirror = neighbor + s;
            if (mirror > 0 && mirror <= O->nnodes && (O->fixedDirection[mirror - 1] & dir) == 0 && O->nd[mirror - 1].numberPoints > 0)
                O->bcReflect[b] = mirror;
        }
    }
    grid_velocity_conditions(MASS_MOMENTUM_CALL);
}

/* ---- IsotropicMat small-rotation law: MoreIsotropicMat.cpp:33-45,185-348; ElasticMPM.cpp:392-403;
 *      Hypo3D/2D MaterialBaseMPM.cpp:1012-1047; IncrementHeatEnergy :982-1006 --------------------------------- */
enum { XX, YY, ZZ, YZ, XZ, XY };

static void isotropic_law(int p, const double du[3][3], const mpmgpu_material *m)
{
    const double *q = m->p;
    double F[3][3], Fn[3][3], dF[3][3];
    /* deformation gradient from ep + wrot (MatPoint3D.cpp:363-376) */
    F[0][0] = 1. + P3(ep, XX, p); F[1][1] = 1. + P3(ep, YY, p); F[2][2] = 1. + P3(ep, ZZ, p);
    F[0][1] = 0.5 * (P3(ep, XY, p) - P3(wrot, 0, p)); F[1][0] = 0.5 * (P3(ep, XY, p) + P3(wrot, 0, p));
    F[0][2] = 0.5 * (P3(ep, XZ, p) - P3(wrot, 1, p)); F[2][0] = 0.5 * (P3(ep, XZ, p) + P3(wrot, 1, p));
    F[1][2] = 0.5 * (P3(ep, YZ, p) - P3(wrot, 2, p)); F[2][1] = 0.5 * (P3(ep, YZ, p) + P3(wrot, 2, p));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dF[i][j] = du[i][j] + (i == j ? 1. : 0.);
    if (O->dim == 3) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Fn[i][j] = dF[i][0] * F[0][j] + dF[i][1] * F[1][j] + dF[i][2] * F[2][j];
    } else {
        memset(Fn, 0, sizeof Fn);
        Fn[0][0] = dF[0][0] * F[0][0] + dF[0][1] * F[1][0]; Fn[0][1] = dF[0][0] * F[0][1] + dF[0][1] * F[1][1];
        Fn[1][0] = dF[1][0] * F[0][0] + dF[1][1] * F[1][0]; Fn[1][1] = dF[1][0] * F[0][1] + dF[1][1] * F[1][1];
        Fn[2][2] = dF[2][2] * F[2][2];
    }
    /* SetDeformationGradientMatrix (MatPoint3D.cpp:320-336) */
    P3(ep, XX, p) = Fn[0][0] - 1.; P3(ep, YY, p) = Fn[1][1] - 1.; P3(ep, ZZ, p) = Fn[2][2] - 1.;
    P3(ep, XY, p) = Fn[1][0] + Fn[0][1]; P3(wrot, 0, p) = Fn[1][0] - Fn[0][1];
    if (O->dim == 3) {
        P3(ep, XZ, p) = Fn[2][0] + Fn[0][2]; P3(ep, YZ, p) = Fn[2][1] + Fn[1][2];
        P3(wrot, 1, p) = Fn[2][0] - Fn[0][2]; P3(wrot, 2, p) = Fn[2][1] - Fn[1][2];
    }
    const double gamma0 = q[20], Cv = q[1], prevT = P3(energies, 5, p);
    double st0[6];
    for (int c = 0; c < 6; c++) st0[c] = P3(sp, c, p);
    double dVoverV, work;
    if (O->dim == 3) {
        double dvxx = du[0][0], dvyy = du[1][1], dvzz = du[2][2];
        double dgamxy = du[0][1] + du[1][0], dgamxz = du[0][2] + du[2][0], dgamyz = du[1][2] + du[2][1];
        double dwxy = du[1][0] - du[0][1], dwxz = du[2][0] - du[0][2], dwyz = du[2][1] - du[1][2];
        dVoverV = dvxx + dvyy + dvzz;
        double ds[6];
        ds[XX] = q[8] * dvxx + q[9] * dvyy + q[10] * dvzz;
        ds[YY] = q[9] * dvxx + q[11] * dvyy + q[12] * dvzz;
        ds[ZZ] = q[10] * dvxx + q[12] * dvyy + q[13] * dvzz;
        ds[YZ] = q[14] * dgamyz; ds[XZ] = q[15] * dgamxz; ds[XY] = q[16] * dgamxy;
        double st[6];
        st[XX] = -dwxy * st0[XY] - dwxz * st0[XZ];
        st[YY] = dwxy * st0[XY] - dwyz * st0[YZ];
        st[ZZ] = dwxz * st0[XZ] + dwyz * st0[YZ];
        st[YZ] = 0.5 * (dwxy * st0[XZ] + dwxz * st0[XY] + dwyz * (st0[YY] - st0[ZZ]));
        st[XZ] = 0.5 * (-dwxy * st0[YZ] + dwxz * (st0[XX] - st0[ZZ]) + dwyz * st0[XY]);
        st[XY] = 0.5 * (dwxy * (st0[XX] - st0[YY]) - dwxz * st0[YZ] - dwyz * st0[XZ]);
        for (int c = 0; c < 6; c++) P3(sp, c, p) += ds[c] + st[c];
        work = 0.5 * ((st0[XX] + P3(sp, XX, p)) * dvxx + (st0[YY] + P3(sp, YY, p)) * dvyy + (st0[ZZ] + P3(sp, ZZ, p)) * dvzz +
                      (st0[YZ] + P3(sp, YZ, p)) * dgamyz + (st0[XZ] + P3(sp, XZ, p)) * dgamxz + (st0[XY] + P3(sp, XY, p)) * dgamxy);
    } else {
        double dvxx = du[0][0], dvyy = du[1][1], dgam = du[0][1] + du[1][0], dwxy = du[1][0] - du[0][1];
        dVoverV = dvxx + dvyy;
        double c1 = q[8] * dvxx + q[9] * dvyy, c2 = q[9] * dvxx + q[11] * dvyy, c3 = q[16] * dgam;
        double dnorm = dwxy * st0[XY], dshear = 0.5 * dwxy * (st0[XX] - st0[YY]);
        P3(sp, XX, p) += c1 - dnorm; P3(sp, YY, p) += c2 + dnorm; P3(sp, XY, p) += c3 + dshear;
        work = 0.5 * ((st0[XX] + P3(sp, XX, p)) * dvxx + (st0[YY] + P3(sp, YY, p)) * dvyy + (st0[XY] + P3(sp, XY, p)) * dgam);
        if (O->cfg.np == MPMGPU_PLANE_STRAIN_MPM) {
            P3(sp, ZZ, p) += q[21] * dvxx + q[22] * dvyy;
        } else {
            double dezz = q[21] * dvxx + q[22] * dvyy;
            P3(ep, ZZ, p) += dezz * (1. + P3(ep, ZZ, p));      /* MPMBase::IncrementDeformationGradientZZ, MPMBase.cpp:637-639 */
            work += 0.5 * (st0[ZZ] + P3(sp, ZZ, p)) * dezz;
            dVoverV += dezz;
        }
    }
    P3(energies, 0, p) += work;
    double dTq0 = -gamma0 * prevT * dVoverV;
    double baseHeat = -Cv * dTq0;
    P3(energies, 2, p) += baseHeat;
    P3(energies, 3, p) += baseHeat / prevT;
}

/* ---- Neohookean: Materials/Neohookean.cpp:177-331, HyperElastic.cpp:104-139,171-204 ------------------------------ */
static void get_F(int p, double F[3][3])
{
    memset(F, 0, 9 * sizeof(double));
    F[0][0] = 1. + P3(ep, XX, p); F[1][1] = 1. + P3(ep, YY, p); F[2][2] = 1. + P3(ep, ZZ, p);
    F[0][1] = 0.5 * (P3(ep, XY, p) - P3(wrot, 0, p)); F[1][0] = 0.5 * (P3(ep, XY, p) + P3(wrot, 0, p));
    if (O->dim == 3) {
        F[0][2] = 0.5 * (P3(ep, XZ, p) - P3(wrot, 1, p)); F[2][0] = 0.5 * (P3(ep, XZ, p) + P3(wrot, 1, p));
        F[1][2] = 0.5 * (P3(ep, YZ, p) - P3(wrot, 2, p)); F[2][1] = 0.5 * (P3(ep, YZ, p) + P3(wrot, 2, p));
    }
}

static void set_F(int p, double F[3][3])
{
    P3(ep, XX, p) = F[0][0] - 1.; P3(ep, YY, p) = F[1][1] - 1.; P3(ep, ZZ, p) = F[2][2] - 1.;
    P3(ep, XY, p) = F[1][0] + F[0][1]; P3(wrot, 0, p) = F[1][0] - F[0][1];
    if (O->dim == 3) {
        P3(ep, XZ, p) = F[2][0] + F[0][2]; P3(ep, YZ, p) = F[2][1] + F[1][2];
        P3(wrot, 1, p) = F[2][0] - F[0][2]; P3(wrot, 2, p) = F[2][1] - F[1][2];
    }
}

static void mat_mul(double a[3][3], double b[3][3], double c[3][3])
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}

/* ---- small-strain / large-rotation helpers (Elastic::useLargeRotation): MaterialBase::LRGetStrainIncrement
 *      (Materials/MaterialBaseMPM.cpp:882-927), Matrix3::Exponential (Common/System/Matrix3.cpp:312-384),
 *      Eigenvalues (:464-520), RightDecompose / LeftDecompose (:564-740), RVoightRT (:190-232) -------------------------- */
/* dF = exp(du) to incrementalDefGradTerms terms: 1 in 3D, 2 in 2D (StartOutput.cpp:116-120) */
static void exp_du(const double du[3][3], double dF[3][3])
{
    memset(dF, 0, 9 * sizeof(double));
    if (O->dim == 3) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dF[i][j] = du[i][j] + (i == j ? 1. : 0.);
        return;
    }
    double c0 = du[0][1] * du[1][0] - du[0][0] * du[1][1], c1 = du[0][0] + du[1][1];
    double beta0 = 0., beta1 = 1., alpha0 = 1., alpha1 = 1., betaz = du[2][2], ezz = 1. + betaz;
    for (int k = 2; k <= 2; k++) {
        double factor = 1 / (double)k, temp = beta1;
        beta1 = factor * (c1 * temp + beta0);
        beta0 = factor * c0 * temp;
        betaz *= factor * du[2][2];
        alpha0 += beta0; alpha1 += beta1; ezz += betaz;
    }
    dF[0][0] = alpha0 + alpha1 * du[0][0]; dF[0][1] = alpha1 * du[0][1];
    dF[1][0] = alpha1 * du[1][0]; dF[1][1] = alpha0 + alpha1 * du[1][1]; dF[2][2] = ezz;
}

/* eigenvalues of a symmetric positive definite 3x3 matrix, trigonometric solution of the characteristic cubic */
static void sym_eigenvalues3(double m[3][3], double lam[3])
{
    double de = m[0][1] * m[1][2], dd = m[0][1] * m[0][1], ee = m[1][2] * m[1][2], ff = m[0][2] * m[0][2];
    double mm = m[0][0] + m[1][1] + m[2][2];
    double c1 = (m[0][0] * m[1][1] + m[0][0] * m[2][2] + m[1][1] * m[2][2]) - (dd + ee + ff);
    double c0 = m[2][2] * dd + m[0][0] * ee + m[1][1] * ff - m[0][0] * m[1][1] * m[2][2] - 2.0 * m[0][2] * de;
    double pp = mm * mm - 3.0 * c1;
    double q = mm * (pp - (3.0 / 2.0) * c1) - (27.0 / 2.0) * c0;
    double sqrt_p = sqrt(fabs(pp));
    double phi = 27.0 * (0.25 * c1 * c1 * (pp - c1) + c0 * (q + 27.0 / 4.0 * c0));
    phi = (1.0 / 3.0) * atan2(sqrt(fabs(phi)), q);
    double c = sqrt_p * cos(phi), s = (1.0 / 1.73205080756887729352744634151) * sqrt_p * sin(phi);
    lam[1] = (1.0 / 3.0) * (mm - c);
    lam[2] = lam[1] + s;
    lam[0] = lam[1] + c;
    lam[1] -= s;
}

/* rotation R of the polar decomposition F = RU (left = 0, through C = F^T F) or F = VR (left = 1, through B = F F^T) */
static void polar_rotation(double F[3][3], int left, double R[3][3])
{
    if (O->dim == 2) {
        double Fsum = F[0][0] + F[1][1], Fdif = F[0][1] - F[1][0], denom = sqrt(Fsum * Fsum + Fdif * Fdif);
        Fsum /= denom; Fdif /= denom;
        memset(R, 0, 9 * sizeof(double));
        R[0][0] = Fsum; R[0][1] = Fdif; R[1][0] = -Fdif; R[1][1] = Fsum; R[2][2] = 1.;
        return;
    }
    double Ft[3][3], C[3][3], C2[3][3], lam[3], U[3][3], Ui[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ft[i][j] = F[j][i];
    if (left) mat_mul(F, Ft, C); else mat_mul(Ft, F, C);
    mat_mul(C, C, C2);
    sym_eigenvalues3(C, lam);
    double l1 = sqrt(lam[0]), l2 = sqrt(lam[1]), l3 = sqrt(lam[2]);
    double i1 = l1 + l2 + l3, i2 = l1 * l2 + l1 * l3 + l2 * l3, i3 = l1 * l2 * l3;
    double d1 = 1. / (i1 * i2 - i3), c2 = -d1, c1 = (i1 * i1 - i2) * d1, cI = i1 * i3 * d1;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        int a = i, b = j;
        if (!left && i > j) { a = j; b = i; }          /* RightDecompose builds U from the upper triangle */
        U[i][j] = c2 * C2[a][b] + c1 * C[a][b] + (i == j ? cI : 0.);
    }
    c1 = 1 / i3;
    double cU = -i1 * c1;
    cI = i2 * c1;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        int a = i, b = j;
        if (!left && i > j) { a = j; b = i; }
        Ui[i][j] = c1 * C[a][b] + cU * U[a][b] + (i == j ? cI : 0.);
    }
    if (left) mat_mul(Ui, F, R); else mat_mul(F, Ui, R);
}

/* Matrix3::RVoightRT: R t R^T of a Voight tensor, stress (engineering shear not doubled) or strain */
static void rotate_voight(double m[3][3], const double t[6], int stress, double o[6])
{
    const double ur = stress ? 2. : 1., ll = stress ? 1. : 2.;
    if (O->dim == 2) {
        o[XX] = m[0][0] * m[0][0] * t[XX] + m[0][1] * m[0][1] * t[YY] + ur * m[0][0] * m[0][1] * t[XY];
        o[YY] = m[1][0] * m[1][0] * t[XX] + m[1][1] * m[1][1] * t[YY] + ur * m[1][1] * m[1][0] * t[XY];
        o[XY] = ll * (m[0][0] * m[1][0] * t[XX] + m[1][1] * m[0][1] * t[YY]) + (m[0][1] * m[1][0] + m[0][0] * m[1][1]) * t[XY];
        o[ZZ] = t[ZZ]; o[YZ] = 0.; o[XZ] = 0.;
        return;
    }
    o[XX] = m[0][0] * m[0][0] * t[XX] + m[0][1] * m[0][1] * t[YY] + m[0][2] * m[0][2] * t[ZZ]
            + ur * (m[0][1] * m[0][2] * t[YZ] + m[0][0] * m[0][2] * t[XZ] + m[0][0] * m[0][1] * t[XY]);
    o[YY] = m[1][0] * m[1][0] * t[XX] + m[1][1] * m[1][1] * t[YY] + m[1][2] * m[1][2] * t[ZZ]
            + ur * (m[1][1] * m[1][2] * t[YZ] + m[1][0] * m[1][2] * t[XZ] + m[1][1] * m[1][0] * t[XY]);
    o[ZZ] = m[2][0] * m[2][0] * t[XX] + m[2][1] * m[2][1] * t[YY] + m[2][2] * m[2][2] * t[ZZ]
            + ur * (m[2][2] * m[2][1] * t[YZ] + m[2][2] * m[2][0] * t[XZ] + m[2][0] * m[2][1] * t[XY]);
    o[YZ] = ll * (m[1][0] * m[2][0] * t[XX] + m[1][1] * m[2][1] * t[YY] + m[2][2] * m[1][2] * t[ZZ])
            + (m[1][2] * m[2][1] + m[1][1] * m[2][2]) * t[YZ] + (m[1][2] * m[2][0] + m[1][0] * m[2][2]) * t[XZ]
            + (m[1][1] * m[2][0] + m[1][0] * m[2][1]) * t[XY];
    o[XZ] = ll * (m[0][0] * m[2][0] * t[XX] + m[0][1] * m[2][1] * t[YY] + m[2][2] * m[0][2] * t[ZZ])
            + (m[0][2] * m[2][1] + m[2][2] * m[0][1]) * t[YZ] + (m[0][2] * m[2][0] + m[0][0] * m[2][2]) * t[XZ]
            + (m[0][1] * m[2][0] + m[0][0] * m[2][1]) * t[XY];
    o[XY] = ll * (m[0][0] * m[1][0] * t[XX] + m[0][1] * m[1][1] * t[YY] + m[0][2] * m[1][2] * t[ZZ])
            + (m[1][1] * m[0][2] + m[1][2] * m[0][1]) * t[YZ] + (m[0][2] * m[1][0] + m[0][0] * m[1][2]) * t[XZ]
            + (m[0][1] * m[1][0] + m[0][0] * m[1][1]) * t[XY];
}

/* LRGetStrainIncrement(CURRENT_CONFIGURATION): F <- exp(du) F on the particle; returns the strain increment in the current
 * configuration de = (dF - dR) F(n-1) Rn^T and the incremental rotation dR = Rn Rn-1^T */
static void lr_strain_increment(int p, const double du[3][3], double de[3][3], double dR[3][3])
{
    double F0[3][3], dF[3][3], F1[3][3], Rnm1[3][3], Rn[3][3], RnT[3][3], Rnm1T[3][3], D[3][3], FR[3][3];
    get_F(p, F0);
    exp_du(du, dF);
    mat_mul(dF, F0, F1);
    set_F(p, F1);
    polar_rotation(F0, 0, Rnm1);
    polar_rotation(F1, 1, Rn);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { RnT[i][j] = Rn[j][i]; Rnm1T[i][j] = Rnm1[j][i]; }
    mat_mul(Rn, Rnm1T, dR);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D[i][j] = dF[i][j] - dR[i][j];
    mat_mul(F0, RnT, FR);
    mat_mul(D, FR, de);
}

/* IsotropicMat::LRConstitutiveLaw (Materials/MoreIsotropicMat.cpp:54-174); residual strains are zero on this path */
static void isotropic_lr_law(int p, const double du[3][3], const mpmgpu_material *m)
{
    const double *q = m->p;
    double de[3][3], dR[3][3], s0[6], sr[6];
    lr_strain_increment(p, du, de, dR);
    for (int c = 0; c < 6; c++) s0[c] = P3(sp, c, p);
    rotate_voight(dR, s0, 1, sr);
    const double gamma0 = q[20], Cv = q[1], prevT = P3(energies, 5, p);
    double dgamxy = de[0][1] + de[1][0], dVoverV = de[0][0] + de[1][1], work;
    if (O->dim == 3) {
        double dgamyz = de[1][2] + de[2][1], dgamxz = de[0][2] + de[2][0];
        dVoverV += de[2][2];
        sr[XX] += q[8] * de[0][0] + q[9] * de[1][1] + q[10] * de[2][2];
        sr[YY] += q[9] * de[0][0] + q[11] * de[1][1] + q[12] * de[2][2];
        sr[ZZ] += q[10] * de[0][0] + q[12] * de[1][1] + q[13] * de[2][2];
        sr[YZ] += q[14] * dgamyz; sr[XZ] += q[15] * dgamxz; sr[XY] += q[16] * dgamxy;
        work = sr[XX] * de[0][0] + sr[YY] * de[1][1] + sr[ZZ] * de[2][2] + sr[YZ] * dgamyz + sr[XZ] * dgamxz + sr[XY] * dgamxy;
    } else {
        sr[XX] += q[8] * de[0][0] + q[9] * de[1][1];
        sr[YY] += q[9] * de[0][0] + q[11] * de[1][1];
        sr[XY] += q[16] * dgamxy;
        work = sr[XX] * de[0][0] + sr[YY] * de[1][1] + sr[XY] * dgamxy;
        if (O->cfg.np == MPMGPU_PLANE_STRAIN_MPM) {
            sr[ZZ] += q[21] * de[0][0] + q[22] * de[1][1];
        } else {
            double dezz = q[21] * de[0][0] + q[22] * de[1][1];
            P3(ep, ZZ, p) += dezz * (1. + P3(ep, ZZ, p));
            work += sr[ZZ] * dezz;
            dVoverV += dezz;
        }
    }
    for (int c = 0; c < 6; c++) P3(sp, c, p) = sr[c];
    P3(energies, 0, p) += work;
    double dTq0 = -gamma0 * prevT * dVoverV, baseHeat = -Cv * dTq0;
    P3(energies, 2, p) += baseHeat;
    P3(energies, 3, p) += baseHeat / prevT;
}

/* MaterialBase::GetArtificialViscosity, MaterialBaseMPM.cpp:1824-1829; dcell = MeshInfo::GetAverageCellSize (equal elements) */
static double artificial_viscosity(double Dkk, double c, const mpmgpu_material *m)
{
    double divuij = fabs(Dkk);
    double dcell = O->dim == 3 ? (O->cfg.gridx + O->cfg.gridy + O->cfg.gridz) / 3. : (O->cfg.gridx + O->cfg.gridy) / 2.;
    return dcell * divuij * (m->p[4] * c + m->p[5] * dcell * divuij);
}

static void neohookean_law(int p, const double du[3][3], double delTime, const mpmgpu_material *m)
{
    const double Gsp = m->p[8], Ksp = m->p[9], Lamesp = m->p[10], gamma0 = m->p[13], Cv = m->p[1];
    const int UofJ = (int)m->p[11];
    double dF[3][3], F[3][3], Fn[3][3], detDf;
    memset(dF, 0, sizeof dF);
    if (O->dim == 3) {          /* Exponential(1) */
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dF[i][j] = du[i][j] + (i == j ? 1. : 0.);
    } else {                    /* 2D Exponential(2), Matrix3.cpp:312-345 */
        double c0 = du[0][1] * du[1][0] - du[0][0] * du[1][1], c1 = du[0][0] + du[1][1];
        double beta0 = 0., beta1 = 1., alpha0 = 1., alpha1 = 1., betaz = du[2][2], ezz = 1. + betaz;
        for (int k = 2; k <= 2; k++) {
            double factor = 1 / (double)k, temp = beta1;
            beta1 = factor * (c1 * temp + beta0);
            beta0 = factor * c0 * temp;
            betaz *= factor * du[2][2];
            alpha0 += beta0; alpha1 += beta1; ezz += betaz;
        }
        dF[0][0] = alpha0 + alpha1 * du[0][0]; dF[0][1] = alpha1 * du[0][1];
        dF[1][0] = alpha1 * du[1][0]; dF[1][1] = alpha0 + alpha1 * du[1][1]; dF[2][2] = ezz;
    }
    get_F(p, F);
    mat_mul(dF, F, Fn);
    set_F(p, Fn);
    double Bo[3][3] = {{P3(eplast, XX, p), P3(eplast, XY, p), P3(eplast, XZ, p)}, {P3(eplast, XY, p), P3(eplast, YY, p), P3(eplast, YZ, p)},
                       {P3(eplast, XZ, p), P3(eplast, YZ, p), P3(eplast, ZZ, p)}};
    if (O->dim == 2) { Bo[0][2] = Bo[2][0] = Bo[1][2] = Bo[2][1] = 0.; }
    double dB[3][3];
    mat_mul(dF, Bo, dB);
    double bxx = dB[0][0] * dF[0][0] + dB[0][1] * dF[0][1] + dB[0][2] * dF[0][2];
    double bxy = dB[0][0] * dF[1][0] + dB[0][1] * dF[1][1] + dB[0][2] * dF[1][2];
    double byy = dB[1][0] * dF[1][0] + dB[1][1] * dF[1][1] + dB[1][2] * dF[1][2];
    double bzz = dB[2][0] * dF[2][0] + dB[2][1] * dF[2][1] + dB[2][2] * dF[2][2];
    P3(eplast, XX, p) = bxx; P3(eplast, XY, p) = bxy; P3(eplast, YY, p) = byy; P3(eplast, ZZ, p) = bzz;
    if (O->dim == 3) {
        P3(eplast, XZ, p) = dB[0][0] * dF[2][0] + dB[0][1] * dF[2][1] + dB[0][2] * dF[2][2];
        P3(eplast, YZ, p) = dB[1][0] * dF[2][0] + dB[1][1] * dF[2][1] + dB[1][2] * dF[2][2];
        detDf = dF[0][0] * (dF[1][1] * dF[2][2] - dF[2][1] * dF[1][2]) - dF[1][0] * (dF[0][1] * dF[2][2] - dF[2][1] * dF[0][2]) +
                dF[2][0] * (dF[0][1] * dF[1][2] - dF[1][1] * dF[0][2]);
    } else detDf = dF[2][2] * (dF[0][0] * dF[1][1] - dF[1][0] * dF[0][1]);
    double Jres = P3(hist, 1, p), dJres = 1.;
    Jres *= dJres;
    P3(hist, 1, p) = Jres;
    double resStretch = pow(Jres, 1. / 3.), Jres23 = resStretch * resStretch;
    if (O->cfg.np == MPMGPU_PLANE_STRESS_MPM) {
        double arg = P3(eplast, XX, p) * P3(eplast, YY, p) - P3(eplast, XY, p) * P3(eplast, XY, p), xn;
        if (UofJ == 1) {
            double a = Lamesp * arg + Gsp * pow(Jres, 4. / 3.), b = Lamesp * sqrt(arg);
            xn = Jres * (b + sqrt(b * b + 4. * Gsp * a)) / (2. * a);
            xn *= xn;
        } else if (UofJ == 2) {
            xn = P3(eplast, ZZ, p);
            double J23 = pow(Jres, 2. / 3.);
            for (int iter = 1; iter < 20; iter++) {
                double fx = Gsp * (xn - J23) + 0.5 * Lamesp * J23 * log(xn * arg / (Jres * Jres));
                double fxp = Gsp + Lamesp * J23 / (2 * xn);
                double xnp1 = xn - fx / fxp;
                if (fabs(xn - xnp1) < 1e-10) break;
                xn = xnp1;
            }
        } else xn = Jres * Jres * (Lamesp + 2. * Gsp) / (Lamesp * arg + 2. * Gsp * pow(Jres, 4. / 3.));
        double dFzz = sqrt(xn / P3(eplast, ZZ, p));
        P3(eplast, ZZ, p) = xn;
        P3(ep, ZZ, p) = dFzz * (1. + P3(ep, ZZ, p)) - 1.;
        detDf *= dFzz;
    }
    double J = detDf * P3(hist, 0, p);
    P3(hist, 0, p) = J;
    double st0[6];
    for (int c = 0; c < 6; c++) st0[c] = P3(sp, c, p);
    double Jeff = J / Jres, p0 = O->pressure[p], Kterm;
    if (UofJ == 1) Kterm = Lamesp * (Jeff - 1.);
    else if (UofJ == 2) Kterm = Lamesp * log(Jeff) / Jeff;
    else Kterm = 0.5 * Lamesp * (Jeff - 1. / Jeff);
    double Bxx = P3(eplast, XX, p), Byy = P3(eplast, YY, p), Bzz = P3(eplast, ZZ, p);
    double Pterm = J * Kterm + Jres * Gsp * ((Bxx + Byy + Bzz) / (3. * Jres23) - 1.);
    double delV = 1. - 1. / detDf, QAVred = 0., AVEnergy = 0.;
    if (delV < 0. && m->p[3] != 0.) {           /* Neohookean.cpp:261-266 */
        QAVred = artificial_viscosity(delV / delTime, sqrt(Ksp * J), m);
        AVEnergy = fabs(QAVred * delV);
    }
    double Pfinal = -Pterm + QAVred;
    O->pressure[p] = Pfinal;
    double avgP = 0.5 * (p0 + Pfinal), dilEnergy = -avgP * delV, resEnergy = -avgP * (1. - 1. / dJres);
    double GJeff = resStretch * Gsp, I1third = (Bxx + Byy + Bzz) / 3.;
    P3(sp, XX, p) = GJeff * (Bxx - I1third); P3(sp, YY, p) = GJeff * (Byy - I1third); P3(sp, ZZ, p) = GJeff * (Bzz - I1third);
    P3(sp, XY, p) = GJeff * P3(eplast, XY, p);
    if (O->dim == 3) { P3(sp, XZ, p) = GJeff * P3(eplast, XZ, p); P3(sp, YZ, p) = GJeff * P3(eplast, YZ, p); }
    double shear = 0.5 * ((P3(sp, XX, p) + st0[XX]) * du[0][0] + (P3(sp, YY, p) + st0[YY]) * du[1][1] + (P3(sp, ZZ, p) + st0[ZZ]) * du[2][2] +
                          (P3(sp, XY, p) + st0[XY]) * (du[0][1] + du[1][0]));
    if (O->dim == 3) shear += 0.5 * ((P3(sp, XZ, p) + st0[XZ]) * (du[0][2] + du[2][0]) + (P3(sp, YZ, p) + st0[YZ]) * (du[1][2] + du[2][1]));
    P3(energies, 0, p) += dilEnergy + shear;
    P3(energies, 1, p) += resEnergy;
    double Gterm = Gsp * (3. - I1third / pow(Jres, 2. / 3.)) / (3. * Jeff), Kratio;
    if (UofJ == 1) Kratio = Lamesp * Jeff + Gterm;
    else if (UofJ == 2) Kratio = Lamesp * (1 - log(Jeff)) / Jeff + Gterm;
    else Kratio = 0.5 * Lamesp * (Jeff + 1. / Jeff) + Gterm;
    Kratio /= Ksp;
    double prevT = P3(energies, 5, p);
    double dTq0 = -J * Kratio * gamma0 * prevT * delV, baseHeat = -Cv * dTq0;
    P3(energies, 2, p) += baseHeat - AVEnergy;
    P3(energies, 3, p) += baseHeat / prevT;
}

/* ---- Mooney: Materials/Mooney.cpp:184-365, HyperElastic.cpp:104-139,171-238 ------------------------------------------ */
static void mooney_law(int p, const double du[3][3], double delTime, const mpmgpu_material *m)
{
    const double G1sp = m->p[8], G2sp = m->p[9], Ksp = m->p[10], gamma0 = m->p[13], Cv = m->p[1];
    const int UofJ = (int)m->p[11];
    double dF[3][3], F[3][3], Fn[3][3], detDf;
    exp_du(du, dF);
    get_F(p, F);
    mat_mul(dF, F, Fn);
    set_F(p, Fn);
    double Bo[3][3] = {{P3(eplast, XX, p), P3(eplast, XY, p), P3(eplast, XZ, p)}, {P3(eplast, XY, p), P3(eplast, YY, p), P3(eplast, YZ, p)},
                       {P3(eplast, XZ, p), P3(eplast, YZ, p), P3(eplast, ZZ, p)}};
    if (O->dim == 2) { Bo[0][2] = Bo[2][0] = Bo[1][2] = Bo[2][1] = 0.; }
    double dB[3][3], Bn[3][3], dFt[3][3];
    mat_mul(dF, Bo, dB);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dFt[i][j] = dF[j][i];
    mat_mul(dB, dFt, Bn);
    P3(eplast, XX, p) = Bn[0][0]; P3(eplast, YY, p) = Bn[1][1]; P3(eplast, ZZ, p) = Bn[2][2]; P3(eplast, XY, p) = Bn[0][1];
    if (O->dim == 3) {
        P3(eplast, XZ, p) = Bn[0][2]; P3(eplast, YZ, p) = Bn[1][2];
        detDf = dF[0][0] * (dF[1][1] * dF[2][2] - dF[2][1] * dF[1][2]) - dF[1][0] * (dF[0][1] * dF[2][2] - dF[2][1] * dF[0][2]) +
                dF[2][0] * (dF[0][1] * dF[1][2] - dF[1][1] * dF[0][2]);
    } else detDf = dF[2][2] * (dF[0][0] * dF[1][1] - dF[1][0] * dF[0][1]);
    double Jres = P3(hist, 1, p), dJres = 1.;
    Jres *= dJres;
    P3(hist, 1, p) = Jres;
    double Bxx = P3(eplast, XX, p), Byy = P3(eplast, YY, p), Bzz = P3(eplast, ZZ, p), Bxy = P3(eplast, XY, p);
    if (O->cfg.np == MPMGPU_PLANE_STRESS_MPM) {         /* :204-258 */
        double arg = Bxx * Byy - Bxy * Bxy, arg12 = sqrt(arg), arg16 = pow(arg, 1. / 6.), arg2 = Bxx + Byy;
        double xn = (1. + P3(ep, ZZ, p)) * (1. + P3(ep, ZZ, p));
        for (int iter = 1; iter < 20; iter++) {
            double xn16 = pow(xn, 1. / 6.), xn12 = sqrt(xn), J13 = xn16 * arg16, J0 = xn12 * arg12, Je = J0 / Jres, mJ2P, mdJ2PdJ;
            if (UofJ == 1) { mJ2P = Ksp * Je * Je * (Je - 1.); mdJ2PdJ = Ksp * Je * (3. * Je - 2.); }
            else if (UofJ == 2) { mJ2P = Ksp * Je * log(Je); mdJ2PdJ = Ksp * (log(Je) + 1.); }
            else { mJ2P = 0.5 * Ksp * Je * (Je * Je - 1.); mdJ2PdJ = 0.5 * Ksp * (3. * Je * Je - 1.); }
            double fx = 3. * Jres * mJ2P + G1sp * (2. * xn - arg2) * J13 + G2sp * (xn * arg2 - 2. * arg) / J13;
            double fxp = (1.5 * J0 / xn) * mdJ2PdJ + G1sp * J13 * (14. * xn - arg2) / (6. * xn) + G2sp * (2. * arg + 5. * xn * arg2) / (6. * J13 * xn);
            double xnp1 = xn - fx / fxp;
            if (fabs(xn - xnp1) < 1e-10) break;
            xn = xnp1;
        }
        double dFzz = sqrt(xn / Bzz);
        Bzz = xn;
        P3(eplast, ZZ, p) = xn;
        P3(ep, ZZ, p) = dFzz * (1. + P3(ep, ZZ, p)) - 1.;
        detDf *= dFzz;
    }
    double J = detDf * P3(hist, 0, p);
    P3(hist, 0, p) = J;
    double st0[6];
    for (int c = 0; c < 6; c++) st0[c] = P3(sp, c, p);
    double Jeff = J / Jres, p0 = O->pressure[p], Kvol;
    if (UofJ == 1) Kvol = Ksp * (Jeff - 1.);
    else if (UofJ == 2) Kvol = Ksp * log(Jeff) / Jeff;
    else Kvol = 0.5 * Ksp * (Jeff - 1. / Jeff);
    double Kterm = J * Kvol;
    double delV = 1. - 1. / detDf, QAVred = 0., AVEnergy = 0.;
    if (delV < 0. && m->p[3] != 0.) {
        QAVred = artificial_viscosity(delV / delTime, sqrt(Ksp * J), m);
        AVEnergy = fabs(QAVred * delV);
    }
    double Pfinal = -Kterm + QAVred;
    O->pressure[p] = Pfinal;
    double avgP = 0.5 * (p0 + Pfinal), dilEnergy = -avgP * delV, resEnergy = -avgP * (1. - 1. / dJres);
    double J23 = pow(J, 2. / 3.), J43 = J23 * J23, JforG1 = J23 / Jres, JforG2 = J43 / Jres;
    double Bxz = O->dim == 3 ? P3(eplast, XZ, p) : 0., Byz = O->dim == 3 ? P3(eplast, YZ, p) : 0.;
    double sn[6];
    sn[XX] = (2 * Bxx - Byy - Bzz) * G1sp / (3. * JforG1) + (Bxx * (Byy + Bzz) - 2 * Byy * Bzz - Bxy * Bxy) * G2sp / (3. * JforG2);
    sn[YY] = (2 * Byy - Bxx - Bzz) * G1sp / (3. * JforG1) + (Byy * (Bxx + Bzz) - 2 * Bxx * Bzz - Bxy * Bxy) * G2sp / (3. * JforG2);
    sn[ZZ] = (2 * Bzz - Bxx - Byy) * G1sp / (3. * JforG1) + (Bzz * (Bxx + Byy) - 2 * Bxx * Byy + 2. * Bxy * Bxy) * G2sp / (3. * JforG2);
    sn[XY] = Bxy * G1sp / JforG1 + (Bzz * Bxy) * G2sp / JforG2;
    sn[XZ] = st0[XZ]; sn[YZ] = st0[YZ];
    if (O->dim == 3) {
        sn[XX] += (2. * Byz * Byz - Bxz * Bxz) * G2sp / (3. * JforG2);
        sn[YY] += (2. * Bxz * Bxz - Byz * Byz) * G2sp / (3. * JforG2);
        sn[ZZ] -= (Bxz * Bxz + Byz * Byz) * G2sp / (3. * JforG2);
        sn[XY] -= Bxz * Byz * G2sp / JforG2;
        sn[XZ] = Bxz * G1sp / JforG1 + (Byy * Bxz - Bxy * Byz) * G2sp / JforG2;
        sn[YZ] = Byz * G1sp / JforG1 + (Bxx * Byz - Bxy * Bxz) * G2sp / JforG2;
    }
    for (int c = 0; c < 6; c++) P3(sp, c, p) = sn[c];
    double shear = 0.5 * ((sn[XX] + st0[XX]) * du[0][0] + (sn[YY] + st0[YY]) * du[1][1] + (sn[ZZ] + st0[ZZ]) * du[2][2] +
                          (sn[XY] + st0[XY]) * (du[0][1] + du[1][0]));
    if (O->dim == 3) shear += 0.5 * ((sn[XZ] + st0[XZ]) * (du[0][2] + du[2][0]) + (sn[YZ] + st0[YZ]) * (du[1][2] + du[2][1]));
    P3(energies, 0, p) += dilEnergy + shear;
    P3(energies, 1, p) += resEnergy;
    double Kratio;
    if (UofJ == 1) Kratio = Jeff;
    else if (UofJ == 2) Kratio = (1 - log(Jeff)) / (Jeff * Jeff);
    else Kratio = 0.5 * (Jeff + 1. / Jeff);
    double prevT = P3(energies, 5, p);
    double dTq0 = -J * Kratio * gamma0 * prevT * delV, baseHeat = -Cv * dTq0;
    P3(energies, 2, p) += baseHeat - AVEnergy;
    P3(energies, 3, p) += baseHeat / prevT;
}

/* ---- IsoPlasticity + LinearHardening: Materials/IsoPlasticity.cpp:128-517, LinearHardening.cpp:93-145 -------------- */
#define SQRT_TWOTHIRDS 0.8164965809277260
/* ---- hardening laws returned numerically: HardeningLawBase::SolveForLambdaBracketed + BracketSolution (HardeningLawBase.cpp:211-381),
 *      NonlinearHardening.cpp:49-74, Nonlinear2Hardening.cpp:28-60, JohnsonCook.cpp:130-249.  Material slots: include/mpmgpu.h ------- */
enum { HARD_LINEAR = 1, HARD_NONLINEAR = 2, HARD_JOHNSONCOOK = 3, HARD_NONLINEAR2 = 6 };
#define TWOTHIRDS 0.6666666666666667
#define SQRT_EIGHT27THS 0.5443310539518174
typedef struct { double alpint, dalpha; } HardAlpha;
typedef struct { int law; double TjcTerm, hmlgTemp; } HardProps;

static HardProps hard_props(const mpmgpu_material *m, double prevT)
{
    HardProps h;
    h.law = (int)m->p[16]; h.TjcTerm = 1.; h.hmlgTemp = 0.;
    if (h.law == HARD_JOHNSONCOOK) {
        h.hmlgTemp = (prevT - m->p[25]) / (m->p[23] - m->p[25]);
        if (h.hmlgTemp > 1.) h.TjcTerm = 0.;
        else if (h.hmlgTemp > 0.) h.TjcTerm = 1. - pow(h.hmlgTemp, m->p[24]);
        else h.TjcTerm = 1.;
    }
    return h;
}

static int dble_equal(double A, double B)       /* Common/System/CommonUtilities.cpp:46-62 */
{
    double diff = fabs(A - B);
    if (diff <= 1.0e-16) return 1;
    A = fabs(A); B = fabs(B);
    return diff <= (B > A ? B : A) * 1.0e-7;
}

/* Johnson-Cook rate term, its derivative factor; returns 1 above the minimum rate */
static int jc_rate_terms(const mpmgpu_material *m, double delTime, const HardAlpha *a, double *term2, double *dterm2)
{
    const double Cjc = m->p[19], ep0 = m->p[20], Djc = m->p[21], n2 = m->p[22];
    const double ep = a->dalpha / (delTime * ep0);
    *dterm2 = 0.;
    if (ep > m->p[26]) {
        *term2 = 1. + Cjc * log(ep);
        *dterm2 = Cjc * ep0 / a->dalpha;
        if (Djc != 0. && ep > 1.) { *term2 += Djc * pow(log(ep), n2); *dterm2 += Djc * ep0 * n2 * pow(log(ep), n2 - 1.) / a->dalpha; }
        return 1;
    }
    *term2 = m->p[27];
    return 0;
}

static double hard_yield(const mpmgpu_material *m, const HardProps *h, double delTime, const HardAlpha *a)
{
    const double yldred = m->p[10];
    if (h->law == HARD_NONLINEAR) return a->alpint < m->p[14] ? yldred * pow(1. + m->p[17] * a->alpint, m->p[18]) : m->p[15];
    if (h->law == HARD_NONLINEAR2) return a->alpint < m->p[14] ? yldred * (1. + m->p[17] * pow(a->alpint, m->p[18])) : m->p[15];
    if (h->hmlgTemp >= 1.) return 0.;
    double term1 = yldred + m->p[17] * pow(a->alpint, m->p[18]);
    double ep = a->dalpha / (delTime * m->p[20]);
    double term2 = ep > m->p[26] ? 1. + m->p[19] * log(ep) : m->p[27];
    if (m->p[21] != 0. && ep > 1.) term2 += m->p[21] * pow(log(ep), m->p[22]);
    return term1 * term2 * h->TjcTerm;
}

static double hard_kprime(const mpmgpu_material *m, const HardProps *h, double delTime, const HardAlpha *a)
{
    const double yldred = m->p[10];
    if (h->law == HARD_NONLINEAR) return a->alpint < m->p[14] ? TWOTHIRDS * yldred * m->p[17] * m->p[18] * pow(1. + m->p[17] * a->alpint, m->p[18] - 1) : 0.;
    if (h->law == HARD_NONLINEAR2) return a->alpint < m->p[14] ? TWOTHIRDS * yldred * m->p[17] * m->p[18] * pow(a->alpint, m->p[18] - 1.) : 0.;
    if (h->hmlgTemp >= 1.) return 0.;
    double dterm1 = m->p[17] * m->p[18] * pow(a->alpint, m->p[18] - 1.), term2, dterm2;
    if (jc_rate_terms(m, delTime, a, &term2, &dterm2)) {
        double term1 = yldred + m->p[17] * pow(a->alpint, m->p[18]);
        return TWOTHIRDS * h->TjcTerm * (dterm1 * term2 + term1 * dterm2);
    }
    return TWOTHIRDS * h->TjcTerm * dterm1 * term2;
}

static double hard_k2prime(const mpmgpu_material *m, const HardProps *h, double fnp1, double delTime, const HardAlpha *a)
{
    const double yldred = m->p[10];
    if (h->law == HARD_NONLINEAR)
        return a->alpint < m->p[14] ? SQRT_EIGHT27THS * yldred * yldred * m->p[17] * m->p[18] * pow(1. + m->p[17] * a->alpint, 2. * m->p[18] - 1) * fnp1 : 0.;
    if (h->law == HARD_NONLINEAR2) {
        if (dble_equal(a->alpint, 0.)) return 0.;
        if (a->alpint < m->p[14]) {
            double alphan = pow(a->alpint, m->p[18]);
            return SQRT_EIGHT27THS * yldred * yldred * m->p[17] * m->p[18] * (1. + m->p[17] * alphan) * alphan * fnp1 / a->alpint;
        }
        return 0.;
    }
    if (dble_equal(a->alpint, 0.)) return 0.;
    if (h->hmlgTemp >= 1.) return 0.;
    double term1 = yldred + m->p[17] * pow(a->alpint, m->p[18]);
    double dterm1 = m->p[17] * m->p[18] * pow(a->alpint, m->p[18] - 1.), term2, dterm2;
    if (jc_rate_terms(m, delTime, a, &term2, &dterm2))
        return SQRT_EIGHT27THS * term1 * term2 * fnp1 * h->TjcTerm * h->TjcTerm * (dterm1 * term2 + dterm2 * term1);
    return SQRT_EIGHT27THS * term1 * term2 * fnp1 * h->TjcTerm * h->TjcTerm * dterm1 * term2;
}

static double hard_yield_increment(const mpmgpu_material *m, const HardProps *h, double delTime, const HardAlpha *a)
{
    if (h->law != HARD_JOHNSONCOOK) return hard_yield(m, h, delTime, a) - m->p[10];
    if (h->hmlgTemp >= 1.) return 0.;
    double ep = a->dalpha / (delTime * m->p[20]);
    double term2 = ep > m->p[26] ? 1. + m->p[19] * log(ep) : m->p[27];
    if (m->p[21] != 0. && ep > 1.) term2 += m->p[21] * pow(log(ep), m->p[22]);
    return m->p[17] * pow(a->alpint, m->p[18]) * term2 * h->TjcTerm;
}

/* stk: trial deviatoric stress xx,yy,zz,yz,xz,xy.  On return a holds the solution's alpha.  NaN where the reference throws. */
static double solve_lambda_bracketed(int planeStress, const mpmgpu_material *m, const HardProps *h, double alpha0, double strial, const double *stk,
                                     double Gred, double psKred, double Pfinal, double delTime, HardAlpha *a)
{
    if (h->law == HARD_JOHNSONCOOK && h->hmlgTemp >= 1.) return strial / (2. * Gred);
    double xl = 0., xh = 0., n1trial = 0., n2trial = 0.;
    if (planeStress) {
        n2trial = -stk[XX] + stk[YY];
        n2trial *= 0.5 * n2trial;
        n2trial += 2. * stk[XY] * stk[XY];
        n1trial = stk[XX] + stk[YY] - 2. * Pfinal;
        n1trial *= n1trial / 6.;
        double epdot = 1.;
        int found = 0;
        for (int step = 0; step < 20; step++) {
            a->dalpha = epdot * delTime;
            a->alpint = alpha0 + a->dalpha;
            double lambdak = a->dalpha / SQRT_TWOTHIRDS;
            double d1 = (1 + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
            double fnp12 = n1trial / (d1 * d1) + n2trial / (d2 * d2);
            double kyld = hard_yield(m, h, delTime, a);
            double gmax = 0.5 * fnp12 - kyld * kyld / 3.;
            if (gmax < 0.) { xl = a->dalpha / SQRT_TWOTHIRDS; found = 1; break; }
            xh = lambdak;
            epdot *= 10.;
        }
        if (!found) return NAN;
    } else {
        double dalpha = strial / (2. * Gred);
        a->alpint = alpha0 + dalpha;
        if (hard_yield(m, h, delTime, a) <= 0.) xh = dalpha / SQRT_TWOTHIRDS;
        xl = dalpha / SQRT_TWOTHIRDS;
    }
    if (xh > xl) return xh;
    double lambdak = 0.5 * (xl + xh);
    a->dalpha = planeStress ? 0. : SQRT_TWOTHIRDS * lambdak;
    a->alpint = alpha0 + a->dalpha;
    double dxold = fabs(xh - xl), dx = dxold;
    for (int step = 1;;) {
        double glam, slope, fnp1 = 0.;
        if (planeStress) {
            double d1 = (1 + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
            double fnp12 = n1trial / (d1 * d1) + n2trial / (d2 * d2);
            double kyld = hard_yield(m, h, delTime, a);
            glam = 0.5 * fnp12 - kyld * kyld / 3.;
            fnp1 = sqrt(fnp12);
            slope = -(psKred * n1trial / (d1 * d1 * d1) + 2 * Gred * n2trial / (d2 * d2 * d2)) - hard_k2prime(m, h, fnp1, delTime, a);
        } else {
            glam = strial - 2 * Gred * lambdak - SQRT_TWOTHIRDS * hard_yield(m, h, delTime, a);
            slope = -2. * Gred - hard_kprime(m, h, delTime, a);
        }
        if (((lambdak - xh) * slope - glam) * ((lambdak - xl) * slope - glam) >= 0. || fabs(2. * glam) > fabs(dxold * slope)) {
            dxold = dx;
            dx = 0.5 * (xh - xl);
            lambdak = xl + dx;
            if (xl == lambdak) break;
        } else {
            dxold = dx;
            dx = glam / slope;
            double temp = lambdak;
            lambdak -= dx;
            if (temp == lambdak) break;
        }
        a->dalpha = planeStress ? SQRT_TWOTHIRDS * lambdak * fnp1 : SQRT_TWOTHIRDS * lambdak;
        a->alpint = alpha0 + a->dalpha;
        if (step++ > 20 || fabs(dx / lambdak) < 0.0001) break;
        if (glam < 0.) xl = lambdak; else xh = lambdak;
    }
    return lambdak;
}

static void isoplasticity_law(int p, const double (*de)[3], double delTime, const mpmgpu_material *m)
{
    const double Gred = m->p[8], Kred = m->p[9], yldred = m->p[10], Epred = m->p[11], gamma0 = m->p[13], Cv = m->p[1];
    const double alphaMax = m->p[14], yldredMin = m->p[15];
    const int is2D = O->dim == 2;
    const int largeRotation = m->p[7] != 0.;
    double dF[3][3], F[3][3], Fn[3][3], deLR[3][3], dR[3][3];
    if (largeRotation) {        /* :140-158: strain increment in the current configuration replaces du */
        lr_strain_increment(p, de, deLR, dR);
        de = deLR;
    } else {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dF[i][j] = de[i][j] + (i == j ? 1. : 0.);
        get_F(p, F);
        if (is2D) { dF[0][2] = dF[2][0] = dF[1][2] = dF[2][1] = 0.; }
        mat_mul(dF, F, Fn);
        set_F(p, Fn);
    }
    /* plane stress terms (IsoPlasticity::GetCopyOfMechanicalProps :551-556) */
    const int planeStress = O->cfg.np == MPMGPU_PLANE_STRESS_MPM;
    const double psRed = 1. / (Kred / (2. * Gred) + 2. / 3.), psLr2G = (Kred / (2. * Gred) - 1. / 3.) * psRed, psKred = Kred * psRed;
    double delV = planeStress ? psRed * (de[0][0] + de[1][1]) : de[0][0] + de[1][1] + de[2][2];       /* :171-174, eres = 0 */
    double dgxy = de[0][1] + de[1][0], dgxz = 0., dgyz = 0.;
    if (!is2D) { dgxz = de[0][2] + de[2][0]; dgyz = de[1][2] + de[2][1]; }
    const double P0 = O->pressure[p];
    /* UpdatePressure */
    double dP = -Kred * delV, dispEnergy = 0.;
    if (delV < 0. && m->p[3] != 0.) {           /* IsoPlasticity::UpdatePressure :474-479 */
        double QAVred = artificial_viscosity(delV / delTime, sqrt(Kred), m);
        dispEnergy += fabs(QAVred * delV);
        dP += QAVred;
    }
    O->pressure[p] += dP;
    double Pfinal = O->pressure[p], prevT = P3(energies, 5, p);
    P3(energies, 0, p) += -Pfinal * delV;
    double dTq0 = -gamma0 * prevT * delV;
    double e0[6], s0[6], st0[6];
    for (int c = 0; c < 6; c++) { e0[c] = P3(eplast, c, p); s0[c] = P3(sp, c, p); st0[c] = s0[c]; }
    double dwxy = de[1][0] - de[0][1];
    if (largeRotation) {        /* :214-222: rotate the plastic strain (stored) and the prior stress by dR */
        double er[6];
        rotate_voight(dR, e0, 0, er);
        rotate_voight(dR, s0, 1, st0);
        for (int c = 0; c < 6; c++) P3(eplast, c, p) = er[c];
    } else if (is2D) {
        double dnorm = 0.5 * dwxy * e0[XY];
        P3(eplast, XX, p) -= dnorm; P3(eplast, YY, p) += dnorm; P3(eplast, XY, p) += dwxy * (e0[XX] - e0[YY]);
        double dn = dwxy * s0[XY];
        st0[XX] -= dn; st0[YY] += dn; st0[XY] += 0.5 * dwxy * (e0[XX] - e0[YY]);        /* :252 uses plastic strain */
    } else {
        double dwxz = de[2][0] - de[0][2], dwyz = de[2][1] - de[1][2];
        double dxy = 0.5 * dwxy * e0[XY], dxz = 0.5 * dwxz * e0[XZ], dyz = 0.5 * dwyz * e0[YZ];
        P3(eplast, XX, p) += -dxy - dxz; P3(eplast, YY, p) += dxy - dyz; P3(eplast, ZZ, p) += dxz + dyz;
        P3(eplast, YZ, p) += dwyz * (e0[YY] - e0[ZZ]) + 0.5 * (dwxz * e0[XY] + dwxy * e0[XZ]);
        P3(eplast, XZ, p) += dwxz * (e0[XX] - e0[ZZ]) + 0.5 * (dwyz * e0[XY] - dwxy * e0[YZ]);
        P3(eplast, XY, p) += dwxy * (e0[XX] - e0[YY]) - 0.5 * (dwyz * e0[XZ] + dwxz * e0[YZ]);
        double sxy = dwxy * s0[XY], sxz = dwxz * s0[XZ], syz = dwyz * s0[YZ];
        st0[XX] += -sxy - sxz; st0[YY] += sxy - syz; st0[ZZ] += sxz + syz;
        st0[YZ] += 0.5 * (dwyz * (s0[YY] - s0[ZZ]) + dwxz * s0[XY] + dwxy * s0[XZ]);
        st0[XZ] += 0.5 * (dwxz * (s0[XX] - s0[ZZ]) + dwyz * s0[XY] - dwxy * s0[YZ]);
        st0[XY] += 0.5 * (dwxy * (s0[XX] - s0[YY]) - dwyz * s0[XZ] - dwxz * s0[YZ]);
    }
    double third = delV / 3., strial[6];
    strial[XX] = st0[XX] + 2. * Gred * (de[0][0] - third);
    strial[YY] = st0[YY] + 2. * Gred * (de[1][1] - third);
    strial[ZZ] = st0[ZZ] + (planeStress ? Pfinal - P0 : 2. * Gred * (de[2][2] - third));        /* :275-278 */
    strial[XY] = st0[XY] + Gred * dgxy;
    strial[YZ] = is2D ? st0[YZ] : st0[YZ] + Gred * dgyz;
    strial[XZ] = is2D ? st0[XZ] : st0[XZ] + Gred * dgxz;
    double alpha0 = P3(hist, 0, p);
    double ss = strial[XX] * strial[XX] + strial[YY] * strial[YY] + strial[ZZ] * strial[ZZ], tt = strial[XY] * strial[XY];
    if (!is2D) tt += strial[XZ] * strial[XZ] + strial[YZ] * strial[YZ];
    double smag = sqrt(ss + tt + tt);
    const int general = m->p[16] > 1.;
    HardProps hp = hard_props(m, prevT);
    HardAlpha ha;
    ha.alpint = alpha0; ha.dalpha = 0.;
    double yield0 = general ? hard_yield(m, &hp, delTime, &ha) : (alpha0 < alphaMax ? yldred + Epred * alpha0 : yldredMin);
    if (smag - SQRT_TWOTHIRDS * yield0 < 0.) {
        for (int c = 0; c < 6; c++) P3(sp, c, p) = strial[c];
        if (!is2D) P3(energies, 0, p) += strial[XX] * de[0][0] + strial[YY] * de[1][1] + strial[ZZ] * de[2][2] + strial[YZ] * dgyz + strial[XZ] * dgxz + strial[XY] * dgxy;
        else {
            if (planeStress) {          /* zz deformation :315-320 */
                double dezz = -psLr2G * (de[0][0] + de[1][1]);
                P3(ep, ZZ, p) += dezz * (1. + P3(ep, ZZ, p));
            }
            P3(energies, 0, p) += strial[XX] * de[0][0] + strial[YY] * de[1][1] + strial[XY] * dgxy;
        }
        double baseHeat = -Cv * dTq0;
        P3(energies, 2, p) += baseHeat - dispEnergy;
        P3(energies, 3, p) += baseHeat / prevT;
        return;
    }
    double lambdak, alpint, dfds[6], dezzTotal = de[2][2], spPS[6] = {0, 0, 0, 0, 0, 0};
    if (planeStress) {
        /* HardeningLawBase::SolveForLambda (unbracketed Newton, HardeningLawBase.cpp:157-202; LinearHardening.cpp:128-131) */
        lambdak = 0.;
        alpint = alpha0;
        if (general) {
            lambdak = solve_lambda_bracketed(1, m, &hp, alpha0, smag, strial, Gred, psKred, Pfinal, delTime, &ha);
            alpint = ha.alpint;
        } else {
        double n2trial = -strial[XX] + strial[YY];
        n2trial *= n2trial / 2;
        n2trial += 2. * strial[XY] * strial[XY];
        double n1trial = strial[XX] + strial[YY] - 2. * Pfinal;
        n1trial *= n1trial / 6.;
        for (int step = 1;; ) {
            double d1 = (1 + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
            double fnp12 = n1trial / (d1 * d1) + n2trial / (d2 * d2);
            double kyld = alpint < alphaMax ? yldred + Epred * alpint : yldredMin;
            double glam = 0.5 * fnp12 - kyld * kyld / 3.;
            double fnp1 = sqrt(fnp12);
            double k2prime = alpint < alphaMax ? 0.5443310539518174 * (yldred + Epred * alpint) * Epred * fnp1 : 0.;
            double slope = -(psKred * n1trial / (d1 * d1 * d1) + 2 * Gred * n2trial / (d2 * d2 * d2)) - k2prime;
            double delLam = -glam / slope;
            lambdak += delLam;
            alpint = alpha0 + SQRT_TWOTHIRDS * lambdak * fnp1;          /* UpdateTrialAlpha, plane stress */
            if (step++ > 20 || fabs(delLam / lambdak) < 0.0001) break;      /* LambdaConverged */
        }
        }
        /* :345-389 */
        double d1 = (1. + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
        double n1 = (strial[XX] + strial[YY] - 2. * Pfinal) / d1, n2 = (-strial[XX] + strial[YY]) / d2;
        double sxx = (n1 - n2) / 2., syy = (n1 + n2) / 2., txy = strial[XY] / d2;
        dfds[XX] = (2. * sxx - syy) / 3.; dfds[YY] = (2. * syy - sxx) / 3.; dfds[ZZ] = -(dfds[XX] + dfds[YY]); dfds[XY] = txy;
        dfds[XZ] = dfds[YZ] = 0.;
        double dPps = -n1 / 3. - Pfinal;
        O->pressure[p] += dPps;
        double dezzp = lambdak * dfds[ZZ];
        double dVoverV = delV + psRed * dezzp;
        P3(energies, 0, p) += -Pfinal * psRed * dezzp - dPps * dVoverV;
        Pfinal = O->pressure[p];
        dezzTotal = -psLr2G * (de[0][0] + de[1][1] - lambdak * (dfds[XX] + dfds[YY])) + dezzp;
        P3(ep, ZZ, p) += dezzTotal * (1. + P3(ep, ZZ, p));
        dTq0 -= gamma0 * prevT * dezzp;
        spPS[XX] = sxx + Pfinal; spPS[YY] = syy + Pfinal; spPS[XY] = txy; spPS[ZZ] = Pfinal;
    } else {
        if (general) {
            lambdak = solve_lambda_bracketed(0, m, &hp, alpha0, smag, strial, Gred, 0., Pfinal, delTime, &ha);
            alpint = ha.alpint;
        } else {
            lambdak = (smag - SQRT_TWOTHIRDS * (yldred + Epred * alpha0)) / (2. * (Gred + Epred / 3.));
            if (alpha0 + SQRT_TWOTHIRDS * lambdak > alphaMax) lambdak = (smag - SQRT_TWOTHIRDS * yldredMin) / (2. * Gred);
            alpint = alpha0 + SQRT_TWOTHIRDS * lambdak;
        }
        for (int c = 0; c < 6; c++) dfds[c] = strial[c] / smag;
    }
    double dep[6];
    dep[XX] = lambdak * dfds[XX]; dep[YY] = lambdak * dfds[YY]; dep[ZZ] = lambdak * dfds[ZZ];
    dep[XY] = 2. * lambdak * dfds[XY];
    dep[XZ] = is2D ? 0. : 2. * lambdak * dfds[XZ];
    dep[YZ] = is2D ? 0. : 2. * lambdak * dfds[YZ];
    P3(eplast, XX, p) += dep[XX]; P3(eplast, YY, p) += dep[YY]; P3(eplast, ZZ, p) += dep[ZZ]; P3(eplast, XY, p) += dep[XY];
    if (!is2D) { P3(eplast, XZ, p) += dep[XZ]; P3(eplast, YZ, p) += dep[YZ]; }
    double sn[6];
    sn[XX] = strial[XX] - 2. * Gred * dep[XX]; sn[YY] = strial[YY] - 2. * Gred * dep[YY]; sn[ZZ] = strial[ZZ] - 2. * Gred * dep[ZZ];
    sn[XY] = strial[XY] - Gred * dep[XY];
    sn[YZ] = is2D ? s0[YZ] : strial[YZ] - Gred * dep[YZ];
    sn[XZ] = is2D ? s0[XZ] : strial[XZ] - Gred * dep[XZ];
    if (planeStress) { sn[XX] = spPS[XX]; sn[YY] = spPS[YY]; sn[ZZ] = spPS[ZZ]; sn[XY] = spPS[XY]; }      /* set above (:386-389) */
    for (int c = 0; c < 6; c++) P3(sp, c, p) = sn[c];
    double work = sn[XX] * de[0][0] + sn[YY] * de[1][1] + sn[XY] * dgxy;
    if (!is2D) work += sn[ZZ] * de[2][2] + sn[YZ] * dgyz + sn[XZ] * dgxz;
    if (O->cfg.np != MPMGPU_PLANE_STRAIN_MPM) work += sn[ZZ] * dezzTotal;          /* :428-431: zz term twice in 3D */
    P3(energies, 0, p) += work;
    double plast = sn[XX] * dep[XX] + sn[YY] * dep[YY] + sn[ZZ] * dep[ZZ] + sn[XY] * dep[XY];
    if (!is2D) plast += sn[XZ] * dep[XZ] + sn[YZ] * dep[YZ];
    dispEnergy += plast - lambdak * SQRT_TWOTHIRDS * (general ? hard_yield_increment(m, &hp, delTime, &ha) : fmax(Epred * alpint, yldredMin - yldred));
    P3(energies, 4, p) += dispEnergy;
    double baseHeat = -Cv * dTq0;
    P3(energies, 2, p) += baseHeat - dispEnergy;
    P3(energies, 3, p) += baseHeat / prevT;
    P3(hist, 0, p) = alpint;
}

/* ---- XPIC(k)/FMPM(k): XPICExtrapolationTask.cpp:49-214, MatVelocityField::XPICSupport :318-397 ------------------- */
static void xpic_extrapolation(int particleUpdate)
{
    const int m = O->cfg.xpic_order, fmpm = O->cfg.using_fmpm;
    const double dt = O->dt;
    if (m <= 1) return;
    for (int i = 0; i < O->nnodes; i++) {          /* INITIALIZE_XPIC */
        NodeField *f = &O->nd[i];
        f->vnext.x = f->vnext.y = f->vnext.z = 0.;
        if (f->numberPoints == 0) continue;
        if (fmpm) {
            double rm = 1. / f->mass;
            f->vprev.x = f->pk.x * rm; f->vprev.y = f->pk.y * rm; f->vprev.z = f->pk.z * rm;
            f->vk = f->vprev;
        } else {
            f->vprev = f->pk;
            f->vprev.x += f->ftot.x * (-dt); f->vprev.y += f->ftot.y * (-dt); f->vprev.z += f->ftot.z * (-dt);
            double rm = 1. / f->mass;
            f->vprev.x *= rm; f->vprev.y *= rm; f->vprev.z *= rm;
            f->vk = f->vprev;
            double s = dt / f->mass;
            f->vk.x += f->ftot.x * s; f->vk.y += f->ftot.y * s; f->vk.z += f->ftot.z * s;
        }
    }
    int nds[64]; double fn[64];
    for (int k = 2; k <= m; k++) {
        for (int p = 0; p < O->nNR; p++) {         /* XPICDoubleLoop: all node pairs of the particle */
            int nn = shape(p, 0, nds, fn, NULL, NULL, NULL);
            for (int i = 0; i < nn; i++) {
                NodeField *fi = &O->nd[nds[i]];
                for (int j = 0; j < nn; j++) {
                    const NodeField *fj = &O->nd[nds[j]];
                    double w = O->mp[p] * (fn[i] * fn[j]) / fi->mass;
                    fi->vnext.x += w * fj->vprev.x; fi->vnext.y += w * fj->vprev.y; fi->vnext.z += w * fj->vprev.z;
                }
            }
        }
        for (int i = 0; i < O->nnodes; i++) {      /* GET_DELTAV */
            NodeField *f = &O->nd[i];
            if (f->numberPoints == 0) continue;
            f->vprev.x -= f->vnext.x; f->vprev.y -= f->vnext.y; f->vprev.z -= f->vnext.z;
        }
        for (int b = 0; b < O->nbc; b++) {         /* GridVelocityConditions(XPIC_*): zero pass only */
            if (!O->bcActive[b]) continue;
            NodeField *f = &O->nd[O->bcNode[b] - 1];
            if (f->numberPoints <= 0) continue;
            const double *n = &O->bcNorm[3 * b];
            double dotn = f->vprev.x * n[0] + f->vprev.y * n[1] + f->vprev.z * n[2];
            f->vprev.x += n[0] * (-dotn); f->vprev.y += n[1] * (-dotn); f->vprev.z += n[2] * (-dotn);
            if (particleUpdate && !fmpm) {
                double s = -f->mass * dotn / dt;
                f->ftot.x += n[0] * s; f->ftot.y += n[1] * s; f->ftot.z += n[2] * s;
            }
        }
        for (int i = 0; i < O->nnodes; i++) {      /* UPDATE_VSTAR */
            NodeField *f = &O->nd[i];
            if (f->numberPoints == 0) continue;
            f->vk.x += f->vprev.x; f->vk.y += f->vprev.y; f->vk.z += f->vprev.z;
            f->vnext.x = f->vnext.y = f->vnext.z = 0.;
        }
    }
}

/* ---- tasks 4 / 9: FullStrainUpdate UpdateStrainsFirstTask.cpp:101-168, MatPoint3D.cpp:45-93 ------------------ */
static void full_strain_update(double strainTime, int postUpdate)
{
    if (O->cfg.using_fmpm && O->cfg.xpic_order > 1) {          /* UpdateStrainsFirstTask.cpp:105-116 */
        if (!postUpdate || !O->cfg.skip_post_extrapolation) xpic_extrapolation(0);
    } else
    for (int i = 0; i < O->nnodes; i++) {      /* GridValueCalculation MatVelocityField.cpp:239-251 */
        NodeField *f = &O->nd[i];
        if (f->numberPoints == 0 || f->mass == 0.) continue;
        double rm = 1. / f->mass;
        f->vk.x = f->pk.x * rm; f->vk.y = f->pk.y * rm; f->vk.z = f->pk.z * rm;
    }
    int nds[64]; double fn[64], xd[64], yd[64], zd[64];
    for (int p = 0; p < O->nNR; p++) {
        int nn = shape(p, 1, nds, fn, xd, yd, zd);
        double dv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < nn; i++) {
            Vec v = O->nd[nds[i]].vk;
            dv[0][0] += v.x * xd[i]; dv[0][1] += v.x * yd[i]; dv[1][0] += v.y * xd[i]; dv[1][1] += v.y * yd[i];
            if (O->dim == 3) {
                dv[0][2] += v.x * zd[i]; dv[1][2] += v.y * zd[i];
                dv[2][0] += v.z * xd[i]; dv[2][1] += v.z * yd[i]; dv[2][2] += v.z * zd[i];
            }
        }
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) dv[a][b] *= strainTime;
        const mpmgpu_material *m = &O->mats[O->matnum[p] - 1];
        if (m->kind == MPMGPU_MAT_ISOTROPIC) { if (m->p[7] != 0.) isotropic_lr_law(p, dv, m); else isotropic_law(p, dv, m); }
        else if (m->kind == MPMGPU_MAT_NEOHOOKEAN) neohookean_law(p, dv, strainTime, m);
        else if (m->kind == MPMGPU_MAT_MOONEY) mooney_law(p, dv, strainTime, m);
        else if (m->kind == MPMGPU_MAT_ISOPLASTICITY) isoplasticity_law(p, dv, strainTime, m);
    }
}

static void task_update_strains_first(void)
{
    if (O->cfg.method == MPMGPU_USL) return;
    full_strain_update(O->cfg.method == MPMGPU_USAVG ? O->dtFirst : O->dt, 0);
}

/* ---- task 5: GridForcesTask.cpp:55-117, MatPoint3D.cpp:248-252 ------------------------------------------------ */
static void task_grid_forces(void)
{
    int nds[64]; double fn[64], xd[64], yd[64], zd[64];
    for (int p = 0; p < O->nNR; p++) {
        int nn = shape(p, 1, nds, fn, xd, yd, zd);
        double mp = O->mp[p], pr = O->pressure[p];
        double sxx = P3(sp, XX, p), syy = P3(sp, YY, p), szz = P3(sp, ZZ, p), syz = P3(sp, YZ, p), sxz = P3(sp, XZ, p), sxy = P3(sp, XY, p);
        double fx = O->pfext ? P3(pfext, 0, p) : 0., fy = O->pfext ? P3(pfext, 1, p) : 0., fz = O->pfext ? P3(pfext, 2, p) : 0.;
        for (int i = 0; i < nn; i++) {
            NodeField *f = &O->nd[nds[i]];
            if (O->dim == 3) {
                f->ftot.x += -mp * ((sxx - pr) * xd[i] + sxy * yd[i] + sxz * zd[i]) + fn[i] * fx;
                f->ftot.y += -mp * (sxy * xd[i] + (syy - pr) * yd[i] + syz * zd[i]) + fn[i] * fy;
                f->ftot.z += -mp * (sxz * xd[i] + syz * yd[i] + (szz - pr) * zd[i]) + fn[i] * fz;
            } else {
                f->ftot.x += -mp * ((sxx - pr) * xd[i] + sxy * yd[i]) + fn[i] * fx;
                f->ftot.y += -mp * (sxy * xd[i] + (syy - pr) * yd[i]) + fn[i] * fy;
            }
        }
    }
}

/* ---- task 6: PostForcesTask.cpp:46-94 ---------------------------------------------------------------------- */
static void task_post_forces(void)
{
    const double *g = O->cfg.gravity;
    int hasG = g[0] != 0. || g[1] != 0. || g[2] != 0.;
    for (int i = 0; i < O->nnodes; i++) {
        NodeField *f = &O->nd[i];
        if (f->numberPoints == 0) continue;
        f->pk = f->pkCopy;                                                                  /* RestoreMomenta */
        if (hasG) { f->ftot.x += f->mass * g[0]; f->ftot.y += f->mass * g[1]; f->ftot.z += f->mass * g[2]; }
    }
    grid_velocity_conditions(GRID_FORCES_CALL);
}

/* ---- task 7: UpdateMomentaTask.cpp:44-64 -------------------------------------------------------------------- */
static void task_update_momenta(void)
{
    for (int i = 0; i < O->nnodes; i++) {
        NodeField *f = &O->nd[i];
        if (f->numberPoints == 0) continue;
        f->pk.x += f->ftot.x * O->dt; f->pk.y += f->ftot.y * O->dt; f->pk.z += f->ftot.z * O->dt;
    }
    grid_velocity_conditions(UPDATE_MOMENTUM_CALL);
}

/* ---- task 8: UpdateParticlesTask.cpp:75-286, MatPoint3D::MoveParticle MatPoint3D.cpp:104-194 ------------------ */
static void task_update_particles(void)
{
    if (O->cfg.xpic_order > 1) xpic_extrapolation(1);          /* UpdateParticlesTask.cpp:66-71 */
    else
    for (int i = 0; i < O->nnodes; i++) {
        NodeField *f = &O->nd[i];
        if (f->numberPoints == 0 || f->mass == 0.) continue;
        double rm = 1. / f->mass;
        f->vk.x = f->pk.x * rm; f->vk.y = f->pk.y * rm; f->vk.z = f->pk.z * rm;
    }
    int m = O->cfg.xpic_order;
    if (!O->cfg.using_fmpm) m = -m;
    const double dt = O->dt;
    int nds[64]; double fn[64];
    for (int p = 0; p < O->nNR; p++) {
        int nn = shape(p, 0, nds, fn, NULL, NULL, NULL);
        double Svk[3] = {0, 0, 0}, Sacc[3] = {0, 0, 0};
        for (int i = 0; i < nn; i++) {
            NodeField *f = &O->nd[nds[i]];
            Svk[0] += f->vk.x * fn[i]; Svk[1] += f->vk.y * fn[i]; Svk[2] += f->vk.z * fn[i];
            if (m <= 0) { double mn = fn[i] / f->mass; Sacc[0] += f->ftot.x * mn; Sacc[1] += f->ftot.y * mn; Sacc[2] += f->ftot.z * mn; }
        }
        const mpmgpu_material *mat = &O->mats[O->matnum[p] - 1];
        double pAlpha = mat->p[2] >= 0. ? mat->p[2] : O->cfg.particle_damping, gAlpha = O->cfg.grid_damping;
        for (int c = 0; c < O->dim; c++) {
            double v = P3(vel, c, p), x = P3(pos, c, p), vm = Svk[c], delV;
            if (m == 0) vm += Sacc[c] * (-dt);
            double Adamp0 = v * pAlpha;
            Adamp0 += vm * gAlpha;
            if (m > 0) { double r = v; v = vm; v += Adamp0 * (-dt); delV = v - r; r += v; x += r * (0.5 * dt); }
            else if (m == 0) { delV = (Sacc[c] - Adamp0) * dt; v += delV; x += (vm + 0.5 * delV) * dt; }
            else { double r = v; v = Svk[c] - Adamp0 * dt; delV = v - r; x += (vm + 0.5 * delV) * dt; }
            P3(vel, c, p) = v; P3(pos, c, p) = x; P3(acc, c, p) = delV / dt;
        }
    }
    for (int p = O->nNR; p < O->n; p++)            /* rigid particles: MovePosition, UpdateParticlesTask.cpp:292-295 */
        for (int c = 0; c < O->dim; c++) P3(pos, c, p) += dt * P3(vel, c, p);
}

/* ---- task 9: UpdateStrainsLastContactTask.cpp:60-152 / UpdateStrainsLastTask.cpp:43-48 ------------------------- */
static void task_update_strains_last(void)
{
    if (O->cfg.method == MPMGPU_USF) return;
    if (!O->cfg.skip_post_extrapolation) {
        for (int i = 0; i < O->nnodes; i++) { O->nd[i].pk.x = O->nd[i].pk.y = O->nd[i].pk.z = 0.; }     /* RezeroNodeTask6 */
        int nds[64]; double fn[64];
        for (int p = 0; p < O->nNR; p++) {
            int nn = shape(p, 0, nds, fn, NULL, NULL, NULL);
            for (int i = 0; i < nn; i++) {
                NodeField *f = &O->nd[nds[i]];
                double fnmp = fn[i] * O->mp[p];
                f->pk.x += P3(vel, 0, p) * fnmp; f->pk.y += P3(vel, 1, p) * fnmp; f->pk.z += P3(vel, 2, p) * fnmp;
            }
        }
        grid_velocity_conditions(UPDATE_STRAINS_LAST_CALL);
    }
    full_strain_update(O->cfg.method == MPMGPU_USAVG ? O->dtLast : O->dt, 1);
}

/* ---- task 11: ResetElementsTask.cpp:196-265, MeshInfo.cpp:171-199,593-633 ---------------------------------------- */
static int pt_in_element(int inElem, const double *x)
{
    int i, j, k;
    elem_ijk(inElem, &i, &j, &k);
    if (x[0] < O->xpts[i] || x[0] >= O->xpts[i + 1]) return 0;
    if (x[1] < O->ypts[j] || x[1] >= O->ypts[j + 1]) return 0;
    if (O->dim == 3 && (x[2] < O->zpts[k] || x[2] >= O->zpts[k + 1])) return 0;
    return 1;
}

static int edge_element(int num)
{
    if (O->dim == 3) {
        int hv = O->horiz * O->vert;
        if (num <= hv || num > O->nelems - hv) return 1;
        if (num % O->horiz <= 1) return 1;
        int xz = num % hv;
        return xz <= O->horiz || xz > O->horiz * (O->vert - 1);
    }
    if (num <= O->horiz || num > O->nelems - O->horiz) return 1;
    return num % O->horiz <= 1;
}

static void task_reset_elements(void)
{
    const double gx = O->cfg.gridx, gy = O->cfg.gridy, gz = O->cfg.gridz;
    for (int p = 0; p < O->n; p++) {
        double x[3] = {P3(pos, 0, p), P3(pos, 1, p), O->dim == 3 ? P3(pos, 2, p) : 0.};
        if (pt_in_element(O->inElem[p], x)) continue;
        int col = (int)((x[0] - O->xpts[0]) / gx), row = (int)((x[1] - O->ypts[0]) / gy), zrow = 0, ok = 1;
        if (col < 0 || col >= O->horiz) { if (x[0] == O->xpts[0] + O->horiz * gx) col = O->horiz - 1; else ok = 0; }
        if (row < 0 || row >= O->vert) { if (x[1] == O->ypts[0] + O->vert * gy) row = O->vert - 1; else ok = 0; }
        if (O->dim == 3) {
            zrow = (int)((x[2] - O->zpts[0]) / gz);
            if (zrow < 0 || zrow >= O->depth) { if (x[2] == O->zpts[0] + O->depth * gz) zrow = O->depth - 1; else ok = 0; }
        }
        int ne = ok ? (O->dim == 3 ? O->horiz * (zrow * O->vert + row) + col + 1 : row * O->horiz + col + 1) : 0;
        if (ne > 0 && !edge_element(ne)) {
            O->inElem[p] = ne;
            O->cross[p] = O->cross[p] >= 0 ? O->cross[p] + 1 : O->cross[p] - 1;
            continue;
        }
        /* LEFT_GRID: count, mark, ReturnToElement by bisection (:232-265) */
        { int c = O->cross[p]; c = c >= 0 ? c + 1 : c - 1; O->cross[p] = c > 0 ? -c : c; }
        double outside[3] = {x[0], x[1], x[2]}, inside[3];
        for (int c = 0; c < 3; c++) inside[c] = c < O->dim ? outside[c] - O->dt * P3(vel, c, p) : 0.;
        if (!pt_in_element(O->inElem[p], inside)) {
            int i, j, k;
            elem_ijk(O->inElem[p], &i, &j, &k);
            inside[0] = (O->xpts[i] + O->xpts[i + 1]) / 2.; inside[1] = (O->ypts[j] + O->ypts[j + 1]) / 2.;
            inside[2] = O->dim == 3 ? (O->zpts[k] + O->zpts[k + 1]) / 2. : 0.;
        }
        for (int pass = 1; pass <= 10; pass++) {
            double mid[3] = {(outside[0] + inside[0]) / 2., (outside[1] + inside[1]) / 2., (outside[2] + inside[2]) / 2.};
            if (pt_in_element(O->inElem[p], mid)) memcpy(inside, mid, sizeof mid); else memcpy(outside, mid, sizeof mid);
        }
        for (int c = 0; c < O->dim; c++) P3(pos, c, p) = inside[c];
    }
}

/* ---- driver ------------------------------------------------------------------------------------------------------- */
typedef void (*taskfn)(void);
static const taskfn TASKS[11] = {task_initialization, task_mass_and_momentum, task_post_extrapolation, task_update_strains_first,
                                 task_grid_forces, task_post_forces, task_update_momenta, task_update_particles,
                                 task_update_strains_last, task_reset_elements,
                                 task_project_rigid_bcs};      /* 10: runs between tasks 1 and 2 */

static double *dupd(const double *src, size_t n) { double *d = (double *)calloc(n ? n : 1, sizeof(double)); if (src) memcpy(d, src, n * sizeof(double)); return d; }
static int *dupi(const int *src, size_t n, int fill) { int *d = (int *)malloc((n ? n : 1) * sizeof(int)); for (size_t i = 0; i < n; i++) d[i] = src ? src[i] : fill; return d; }

/* NodalVelBC::reflectedNode (1-based, <= 0 none) and reflectRatio of the grid BCs, in list order */
int oracle_set_bc_reflections(int n, const int *reflected, const double *ratio);

int oracle_create(const mpmgpu_config *cfg, int nmat, const mpmgpu_material *mats, const mpmgpu_particles *h,
                  int nbc, const int *bcNode, const double *bcNorm, const double *bcValue, const int *bcActive, const int *bcSym,
                  double dt, double dtFirst, double dtLast)
{
    O = (Oracle *)calloc(1, sizeof(Oracle));
    O->cfg = *cfg;
    O->dim = cfg->np == MPMGPU_THREED_MPM ? 3 : 2;
    O->horiz = cfg->horiz; O->vert = cfg->vert; O->depth = O->dim == 3 ? cfg->depth : 1;
    O->xplane = 1; O->yplane = O->horiz + 1; O->zplane = (O->horiz + 1) * (O->vert + 1);
    O->nnodes = O->zplane * (O->dim == 3 ? O->depth + 1 : 1);
    O->nelems = O->horiz * O->vert * O->depth;
    O->xpts = dupd(cfg->xpts, O->horiz + 1); O->ypts = dupd(cfg->ypts, O->vert + 1);
    O->zpts = O->dim == 3 ? dupd(cfg->zpts, O->depth + 1) : NULL;
    O->nmat = nmat;
    O->mats = (mpmgpu_material *)malloc(nmat * sizeof(mpmgpu_material));
    memcpy(O->mats, mats, nmat * sizeof(mpmgpu_material));
    size_t n = h->n;
    O->n = h->n; O->nNR = h->n_nonrigid;
    O->pos = dupd(h->pos, 3 * n); O->vel = dupd(h->vel, 3 * n); O->mp = dupd(h->mp, n); O->lp = dupd(h->lp, 3 * n);
    O->ncpos = dupd(NULL, 3 * n); O->sp = dupd(h->sp, 6 * n); O->pressure = dupd(h->pressure, n);
    O->ep = dupd(h->ep, 6 * n); O->wrot = dupd(h->wrot, 3 * n); O->eplast = dupd(h->eplast, 6 * n);
    O->energies = dupd(h->energies, 6 * n); O->hist = dupd(h->history, MPMGPU_MAX_HISTORY * n);
    O->pfext = h->pfext ? dupd(h->pfext, 3 * n) : NULL; O->acc = dupd(NULL, 3 * n);
    O->inElem = dupi(h->in_elem, n, 1); O->matnum = dupi(h->matnum, n, 1); O->cross = dupi(h->crossings, n, 0);
    O->nd = (NodeField *)calloc(O->nnodes, sizeof(NodeField));
    O->nbc = nbc;
    O->bcNode = dupi(bcNode, nbc, 0); O->bcNorm = dupd(bcNorm, 3 * (size_t)nbc); O->bcValue = dupd(bcValue, nbc);
    O->bcActive = dupi(bcActive, nbc, 1); O->bcSym = dupi(bcSym, nbc, 0);
    O->nbcGrid = nbc; O->bcCap = nbc;
    O->bcDir = dupi(NULL, nbc, 0);
    O->bcMirror = dupi(NULL, nbc, 0); O->bcReflect = dupi(NULL, nbc, -1);
    O->gridRatio = dupd(NULL, nbc);
    O->fixedDirection = (int *)calloc(O->nnodes, sizeof(int));
    for (int b = 0; b < nbc; b++) {            /* NodalVelBC.cpp:40-45: direction bits of the grid BCs */
        int bits = bcSym ? bcSym[b] & 7 : 0;
        for (int d = 0; d < 3; d++) if (bcNorm[3 * b + d] != 0.) bits |= 1 << d;
        O->bcDir[b] = bits;
        O->fixedDirection[bcNode[b] - 1] |= bits;
    }
    O->dt = dt; O->dtFirst = dtFirst; O->dtLast = dtLast;
    O->ncorner = 0;
    if (cfg->shape == MPMGPU_LINEAR_CPDI || cfg->shape == MPMGPU_BSPLINE_CPDI) O->ncorner = O->dim == 3 ? 8 : 4;
    if (cfg->shape == MPMGPU_QUADRATIC_CPDI) O->ncorner = 9;
    if (O->ncorner) {
        O->cpElem = dupi(NULL, n * O->ncorner, 1); O->cpXi = dupd(NULL, 3 * n * O->ncorner); O->cpWg = dupd(NULL, 3 * n * O->ncorner);
        O->cpWs = dupd(NULL, 9);
        for (int c = 0; c < O->ncorner; c++)        /* MPMBase::AllocateCPDIorGIMPStructures, MPMBase.cpp:160-175 */
            O->cpWs[c] = O->ncorner == 9 ? (c < 4 ? 1. / 36. : (c < 8 ? 1. / 9. : 4. / 9.)) : (O->dim == 3 ? 0.125 : 0.25);
    }
    return 0;
}

int oracle_cpdi_left_grid(void) { return O ? O->cpdiLeftGrid : 0; }

int oracle_set_xpic(int order, int using_fmpm) { if (!O) return -1; O->cfg.xpic_order = order; O->cfg.using_fmpm = using_fmpm; return 0; }

int oracle_task(int t) { if (!O || t < 0 || t > 10) return -1; TASKS[t](); if (t == 9) O->mstep++; return 0; }

int oracle_step(int nsteps)
{
    if (!O) return -1;
    for (int s = 0; s < nsteps; s++) for (int t = 0; t < 10; t++) { oracle_task(t); if (t == 1) oracle_task(10); }
    return 0;
}

int oracle_get_particles(mpmgpu_particles *h)
{
    size_t n = O->n;
    if (h->pos) memcpy(h->pos, O->pos, 3 * n * sizeof(double));
    if (h->vel) memcpy(h->vel, O->vel, 3 * n * sizeof(double));
    if (h->sp) memcpy(h->sp, O->sp, 6 * n * sizeof(double));
    if (h->pressure) memcpy(h->pressure, O->pressure, n * sizeof(double));
    if (h->ep) memcpy(h->ep, O->ep, 6 * n * sizeof(double));
    if (h->wrot) memcpy(h->wrot, O->wrot, 3 * n * sizeof(double));
    if (h->eplast) memcpy(h->eplast, O->eplast, 6 * n * sizeof(double));
    if (h->energies) memcpy(h->energies, O->energies, 6 * n * sizeof(double));
    if (h->history) memcpy(h->history, O->hist, MPMGPU_MAX_HISTORY * n * sizeof(double));
    if (h->acc) memcpy(h->acc, O->acc, 3 * n * sizeof(double));
    if (h->in_elem) memcpy(h->in_elem, O->inElem, n * sizeof(int));
    if (h->crossings) memcpy(h->crossings, O->cross, n * sizeof(int));
    return 0;
}

int oracle_get_nodes(mpmgpu_nodes *h)
{
    size_t nn = O->nnodes;
    for (size_t i = 0; i < nn; i++) {
        const NodeField *f = &O->nd[i];
        if (h->number_points) h->number_points[i] = f->numberPoints;
        if (h->mass) h->mass[i] = f->mass;
        if (h->pk) { h->pk[i] = f->pk.x; h->pk[nn + i] = f->pk.y; h->pk[2 * nn + i] = f->pk.z; }
        if (h->ftot) { h->ftot[i] = f->ftot.x; h->ftot[nn + i] = f->ftot.y; h->ftot[2 * nn + i] = f->ftot.z; }
        if (h->vk) { h->vk[i] = f->vk.x; h->vk[nn + i] = f->vk.y; h->vk[2 * nn + i] = f->vk.z; }
        if (h->pk_copy) { h->pk_copy[i] = f->pkCopy.x; h->pk_copy[nn + i] = f->pkCopy.y; h->pk_copy[2 * nn + i] = f->pkCopy.z; }
    }
    return 0;
}

/* The constitutive laws alone, on caller-owned state arrays ([component][n], as in mpmgpu_particles): lets tests check a
 * law against closed forms or against another implementation without building a grid.  du is [n][9], row-major. */
int oracle_law_batch(int np, double gridx, double gridy, double gridz, const mpmgpu_material *m, int n,
                     double *sp, double *pressure, double *ep, double *wrot, double *eplast, double *energies, double *hist,
                     const double *du, double delTime)
{
    Oracle *saved = O, tmp;
    memset(&tmp, 0, sizeof tmp);
    tmp.cfg.np = np; tmp.cfg.gridx = gridx; tmp.cfg.gridy = gridy; tmp.cfg.gridz = gridz;
    tmp.dim = np == MPMGPU_THREED_MPM ? 3 : 2;
    tmp.n = n; tmp.nNR = n;
    tmp.sp = sp; tmp.pressure = pressure; tmp.ep = ep; tmp.wrot = wrot; tmp.eplast = eplast; tmp.energies = energies; tmp.hist = hist;
    O = &tmp;
    for (int p = 0; p < n; p++) {
        double d[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[i][j] = du[(size_t)9 * p + 3 * i + j];
        if (m->kind == MPMGPU_MAT_ISOTROPIC) { if (m->p[7] != 0.) isotropic_lr_law(p, d, m); else isotropic_law(p, d, m); }
        else if (m->kind == MPMGPU_MAT_NEOHOOKEAN) neohookean_law(p, d, delTime, m);
        else if (m->kind == MPMGPU_MAT_MOONEY) mooney_law(p, d, delTime, m);
        else if (m->kind == MPMGPU_MAT_ISOPLASTICITY) isoplasticity_law(p, d, delTime, m);
        else { O = saved; return -1; }
    }
    O = saved;
    return 0;
}

/* Shape functions and element search alone, on caller-owned arrays: for tests that compare another implementation of the same
 * formulas (the device source compiled for the host, tests/test_device_shape_cpu.py) without building a run.
 * Linear and uGIMP only (cfg->shape); nds/fn/xd/yd/zd are [n][64], count[n]. */
static void batch_grid(Oracle *t, const mpmgpu_config *cfg)
{
    memset(t, 0, sizeof *t);
    t->cfg = *cfg;
    t->dim = cfg->np == MPMGPU_THREED_MPM ? 3 : 2;
    t->horiz = cfg->horiz; t->vert = cfg->vert; t->depth = t->dim == 3 ? cfg->depth : 1;
    t->xplane = 1; t->yplane = t->horiz + 1; t->zplane = (t->horiz + 1) * (t->vert + 1);
    t->nnodes = t->zplane * (t->dim == 3 ? t->depth + 1 : 1);
    t->nelems = t->horiz * t->vert * t->depth;
    t->xpts = (double *)cfg->xpts; t->ypts = (double *)cfg->ypts; t->zpts = (double *)cfg->zpts;
}

int oracle_shape_batch(const mpmgpu_config *cfg, int n, int *inElem, double *ncpos, double *lp, int getDeriv,
                       int *count, int *nds, double *fn, double *xd, double *yd, double *zd)
{
    if (cfg->shape != MPMGPU_POINT_GIMP && cfg->shape != MPMGPU_UNIFORM_GIMP) return -1;
    Oracle *saved = O, tmp;
    batch_grid(&tmp, cfg);
    tmp.n = n; tmp.nNR = n; tmp.inElem = inElem; tmp.ncpos = ncpos; tmp.lp = lp;
    O = &tmp;
    for (int p = 0; p < n; p++)
        count[p] = shape(p, getDeriv, nds + (size_t)64 * p, fn + (size_t)64 * p, xd + (size_t)64 * p, yd + (size_t)64 * p, zd + (size_t)64 * p);
    O = saved;
    return 0;
}

/* x is [n][3]; elem[n] = MeshInfo::FindElementFromPoint (0 = off the grid) */
int oracle_find_element_batch(const mpmgpu_config *cfg, int n, const double *x, int *elem)
{
    Oracle *saved = O, tmp;
    batch_grid(&tmp, cfg);
    O = &tmp;
    for (int p = 0; p < n; p++) elem[p] = find_element(x + (size_t)3 * p);
    O = saved;
    return 0;
}

void oracle_destroy(void)
{
    if (!O) return;
    free(O->xpts); free(O->ypts); free(O->zpts); free(O->mats);
    free(O->pos); free(O->vel); free(O->mp); free(O->lp); free(O->ncpos); free(O->sp); free(O->pressure); free(O->ep);
    free(O->wrot); free(O->eplast); free(O->energies); free(O->hist); free(O->pfext); free(O->acc);
    free(O->inElem); free(O->matnum); free(O->cross); free(O->nd);
    free(O->bcNode); free(O->bcNorm); free(O->bcValue); free(O->bcActive); free(O->bcSym); free(O->bcDir); free(O->fixedDirection); free(O->bcMirror); free(O->bcReflect); free(O->gridRatio);
    free(O);
    O = NULL;
}

int oracle_set_bc_reflections(int n, const int *reflected, const double *ratio)
{
    if (!O || n != O->nbcGrid) return 1;
    for (int b = 0; b < n; b++) { O->bcReflect[b] = reflected[b] > 0 ? reflected[b] : -1; O->gridRatio[b] = ratio[b]; }
    return 0;
}
