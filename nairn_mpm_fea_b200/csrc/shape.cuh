// Shape functions and element indexing on the structured grid.
//
// Linear ("Classic"/POINT_GIMP): reference Elements/EightNodeIsoparamBrick.cpp:87-105 (3D),
//                                Common/Elements/FourNodeIsoparam.cpp:189-210 (2D)
// uGIMP: reference Elements/EightNodeIsoparamBrick.cpp:289-397 (3D, incl. the inv_size_z = 1/(4 lp.y)
//        quirk at :303), Common/Elements/FourNodeIsoparam.cpp:431-519 (2D, dSvp written as
//        -xp*inv_size*2 at :494)
// Natural coordinates: EightNodeIsoparamBrick.cpp:276-281, FourNodeIsoparam.cpp:351-356
// Element search: NairnMPM_Class/MeshInfo.cpp:593-633, edge ring :171-199
//
// Rather than the reference's 64 (16 in 2D) candidate table, the 1-D GIMP weights of the four
// node columns at natural coordinate -3,-1,1,3 are evaluated once per axis and tensor-multiplied.
// The set of nodes with xp < 2+lp on every axis is the same set the reference keeps; the
// per-node products are formed in the reference's operation order.
#pragma once
#include "mpm_types.cuh"

// ---- element <-> (col,row,rank) and node indexing ------------------------------------------
struct ElemIJK { int i, j, k; };

__device__ __forceinline__ ElemIJK elem_ijk(const Grid &g, int inElem)
{
    int e0 = inElem - 1;
    ElemIJK c;
    c.i = e0 % g.horiz;
    int r = e0 / g.horiz;
    c.j = r % g.vert;
    c.k = r / g.vert;
    return c;
}

__device__ __forceinline__ int elem_node0(const Grid &g, ElemIJK c)
{
    return c.k * g.zplane + c.j * g.yplane + c.i;    // 0-based index of the element's first node
}

// (2x - xmin - xmax)/(xmax - xmin) with the element's own extents, no contraction
__device__ __forceinline__ double natural_coord(double x, double lo, double hi)
{
    double num = __dsub_rn(__dsub_rn(__dmul_rn(2., x), lo), hi);
    return __ddiv_rn(num, __dsub_rn(hi, lo));
}

template <int DIM>
__device__ __forceinline__ void get_xipos(const Grid &g, int inElem, const double pos[3], double xi[3])
{
    ElemIJK c = elem_ijk(g, inElem);
    xi[0] = natural_coord(pos[0], g.xpts[c.i], g.xpts[c.i + 1]);
    xi[1] = natural_coord(pos[1], g.ypts[c.j], g.ypts[c.j + 1]);
    xi[2] = (DIM == 3) ? natural_coord(pos[2], g.zpts[c.k], g.zpts[c.k + 1]) : 0.;
}

// PtInElement: xmin <= x < xmax on node coordinates (EightNodeIsoparamBrick.cpp:207-212)
template <int DIM>
__device__ __forceinline__ bool pt_in_element(const Grid &g, int inElem, const double pos[3])
{
    ElemIJK c = elem_ijk(g, inElem);
    if (pos[0] < g.xpts[c.i] || pos[0] >= g.xpts[c.i + 1]) return false;
    if (pos[1] < g.ypts[c.j] || pos[1] >= g.ypts[c.j + 1]) return false;
    if (DIM == 3) {
        if (pos[2] < g.zpts[c.k] || pos[2] >= g.zpts[c.k + 1]) return false;
    }
    return true;
}

// MeshInfo::FindElementFromPoint for equal element sizes; returns 1-based element or 0 if off grid
template <int DIM>
__device__ __forceinline__ int find_element_from_point(const Grid &g, const double pos[3])
{
    int col = (int)__ddiv_rn(__dsub_rn(pos[0], g.xmin), g.gx);
    if (col < 0 || col >= g.horiz) {
        if (pos[0] == __dadd_rn(g.xmin, __dmul_rn((double)g.horiz, g.gx))) col = g.horiz - 1;
        else return 0;
    }
    int row = (int)__ddiv_rn(__dsub_rn(pos[1], g.ymin), g.gy);
    if (row < 0 || row >= g.vert) {
        if (pos[1] == __dadd_rn(g.ymin, __dmul_rn((double)g.vert, g.gy))) row = g.vert - 1;
        else return 0;
    }
    if (DIM == 3) {
        int zrow = (int)__ddiv_rn(__dsub_rn(pos[2], g.zmin), g.gz);
        if (zrow < 0 || zrow >= g.depth) {
            if (pos[2] == __dadd_rn(g.zmin, __dmul_rn((double)g.depth, g.gz))) zrow = g.depth - 1;
            else return 0;
        }
        return g.horiz * (zrow * g.vert + row) + col + 1;
    }
    return row * g.horiz + col + 1;
}

// MeshInfo::EdgeElement2D/3D on the 1-based element number
template <int DIM>
__device__ __forceinline__ bool edge_element(const Grid &g, int num)
{
    if (DIM == 3) {
        int hv = g.horiz * g.vert;
        if (num <= hv || num > g.nelems - hv) return true;
        int yz = num % g.horiz;
        if (yz <= 1) return true;
        int xz = num % hv;
        if (xz <= g.horiz || xz > g.horiz * (g.vert - 1)) return true;
        return false;
    }
    if (num <= g.horiz || num > g.nelems - g.horiz) return true;
    int col = num % g.horiz;
    return col <= 1;
}

// ---- 1-D uGIMP weights ---------------------------------------------------------------------
// For node columns at natural coordinate xi_n = -3,-1,1,3 (grid offsets -1..2 from the element's
// first node).  S[o] = 0 and ok bit clear when |xi - xi_n| >= 2+lp (the reference skips those).
// dS[o] already carries the reference's xsign factor.
struct Gimp1D {
    double S[4], dS[4];
    unsigned ok;
};

template <bool GRAD, bool TWO_D_FORM>
__device__ __forceinline__ void gimp_1d(double xi, double lp, double inv_size, Gimp1D &w)
{
    const double q1 = 2. - lp, q2 = 2. + lp;
    w.ok = 0;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const double xn = (double)(2 * o - 3);
        double xp = fabs(xi - xn);
        double S = 0., dS = 0.;
        if (xp < q2) {
            w.ok |= 1u << o;
            if (xp < lp) {
                S = ((4. - lp) * lp - xp * xp) * inv_size;
                if (GRAD) dS = TWO_D_FORM ? -xp * inv_size * 2.0 : -xp / (2. * lp);
            } else if (xp <= q1) {
                S = 0.5 * (2. - xp);
                if (GRAD) dS = -0.5;
            } else {
                double arg = (q2 - xp) * inv_size;
                S = 2. * lp * arg * arg;
                if (GRAD) dS = -arg;
            }
            if (GRAD && !(xi > xn)) dS = -dS;      // xsign = xi>xn ? 1 : -1
        }
        w.S[o] = S;
        w.dS[o] = dS;
    }
}

// ---- 1-D quadratic B-spline weights (B2SPLINE) and their GIMP form (B2GIMP) ------------------------------------------------
// Same four node columns as uGIMP.  B2SPLINE: EightNodeIsoparamBrick::SplineShapeFunction (Elements/EightNodeIsoparamBrick.cpp:110-204),
// FourNodeIsoparam::SplineShapeFunction (Common/Elements/FourNodeIsoparam.cpp:217-294): support |eta| < 3, dS carries the sign, the caller
// multiplies by 1/dx.  B2GIMP: BGimpShapeFunction (EightNodeIsoparamBrick.cpp:477-634, FourNodeIsoparam.cpp:770-884): support
// |eta| < 3 + lp, the caller multiplies by 2/dx.
template <bool GRAD>
__device__ __forceinline__ void bspline_1d(double xi, Gimp1D &w)
{
    w.ok = 0;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const double xn = (double)(2 * o - 3);
        const double etai = xi - xn;
        const double t = fabs(etai);
        double S = 0., dS = 0.;
        if (!(t >= 3.)) {
            w.ok |= 1u << o;
            if (t <= 1.0) {
                S = 0.25 * (3. - etai * etai);
                if (GRAD) dS = -etai;
            } else {
                const double arg = 3. - t;
                S = 0.125 * arg * arg;
                if (GRAD) dS = etai >= 0. ? 0.5 * (etai - 3) : 0.5 * (etai + 3);
            }
        }
        w.S[o] = S;
        w.dS[o] = dS;
    }
}

// xiSign: the coordinate the reference compares with the node to pick the sign of the gradient (its own axis, except the 3D
// z gradient, which uses the y coordinate: EightNodeIsoparamBrick.cpp:586)
template <bool GRAD, bool TWO_D_FORM>
__device__ __forceinline__ void bgimp_1d(double xi, double xiSign, double lp, Gimp1D &w)
{
    const double b1 = 1. - lp, b2 = 1. + lp, b3 = 3. - lp, b4 = 3. + lp;
    const double inv_size = 1. / (48. * lp), oneTwelth = 1. / 12.;
    w.ok = 0;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const double xn = (double)(2 * o - 3);
        const double xp = fabs(xi - xn);
        double S = 0., dS = 0.;
        if (!(xp >= b4)) {
            w.ok |= 1u << o;
            if (xp < b1) {
                S = TWO_D_FORM ? (9. - lp * lp - 3. * xp * xp) * oneTwelth : (9. - lp * lp - 3 * xp * xp) * oneTwelth;
                if (GRAD) dS = -0.5 * xp;
            } else if (xp < b2) {
                const double arg = xp - 1.;
                const double lp2 = lp * lp;
                if (TWO_D_FORM) S = (lp2 * (9. * arg - lp) + 3. * arg * arg * arg + 3. * lp * (15. - xp * (6. + xp))) * inv_size;
                else { const double lp3 = lp2 * lp; S = (9. * lp2 * arg + 3. * arg * arg * arg + 3. * lp * (15. - xp * (6. + xp)) - lp3) * inv_size; }
                if (GRAD) dS = (3 * lp2 + 3. * arg * arg - 2. * lp * (3. + xp)) * 3. * inv_size;
            } else if (xp <= b3) {
                const double arg = xp - 3.;
                S = (lp * lp + 3. * arg * arg) * 0.5 * oneTwelth;
                if (GRAD) dS = 0.25 * (xp - 3.);
            } else {
                const double arg = 3. + lp - xp;
                S = arg * arg * arg * inv_size;
                if (GRAD) dS = -arg * arg * 3. * inv_size;
            }
            if (GRAD && !(xiSign > xn)) dS = -dS;
        }
        w.S[o] = S;
        w.dS[o] = dS;
    }
}

// ---- per-particle node loop ----------------------------------------------------------------
// f(node0based, S, dSdx, dSdy, dSdz) is called for every node of the particle's stencil.
template <int DIM, int SHAPE, bool GRAD, class F>
__device__ __forceinline__ void for_each_node(const Grid &g, int inElem, const double xi[3], const double lp[3], F &&f)
{
    const ElemIJK c = elem_ijk(g, inElem);
    const int n0 = elem_node0(g, c);
    if (SHAPE == SHAPE_LINEAR) {
        const double dx = g.xpts[c.i + 1] - g.xpts[c.i];
        const double dy = g.ypts[c.j + 1] - g.ypts[c.j];
        if (DIM == 3) {
            const double dz = g.zpts[c.k + 1] - g.zpts[c.k];
            const int xo[8] = {0, 1, 1, 0, 0, 1, 1, 0}, yo[8] = {0, 0, 1, 1, 0, 0, 1, 1}, zo[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const double sx = xo[a] ? 1. : -1., sy = yo[a] ? 1. : -1., sz = zo[a] ? 1. : -1.;
                const double t1 = 1. + sx * xi[0], t2 = 1. + sy * xi[1], t3 = 1. + sz * xi[2];
                const double S = 0.125 * t1 * t2 * t3;
                double gxv = 0., gyv = 0., gzv = 0.;
                if (GRAD) {
                    gxv = 0.25 * sx * t2 * t3 / dx;
                    gyv = 0.25 * sy * t1 * t3 / dy;
                    gzv = 0.25 * sz * t1 * t2 / dz;
                }
                f(n0 + xo[a] + yo[a] * g.yplane + zo[a] * g.zplane, S, gxv, gyv, gzv);
            }
        } else {
            const int xo[4] = {0, 1, 1, 0}, yo[4] = {0, 0, 1, 1};
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double sx = xo[a] ? 1. : -1., sy = yo[a] ? 1. : -1.;
                const double t1 = 1. + sx * xi[0], t2 = 1. + sy * xi[1];
                const double S = 0.25 * t1 * t2;
                double gxv = 0., gyv = 0.;
                if (GRAD) {
                    gxv = 0.5 * sx * t2 / dx;
                    gyv = 0.5 * sy * t1 / dy;
                }
                f(n0 + xo[a] + yo[a] * g.yplane, S, gxv, gyv, 0.);
            }
        }
    } else if (SHAPE == SHAPE_UGIMP || SHAPE == SHAPE_B2SPLINE || SHAPE == SHAPE_B2GIMP) {
        Gimp1D wx, wy, wz;
        const double gnum = SHAPE == SHAPE_B2SPLINE ? 1.0 : 2.0;       // the spline gradients are per half cell
        if (SHAPE == SHAPE_UGIMP) {
            gimp_1d<GRAD, DIM == 2>(xi[0], lp[0], 1. / (4. * lp[0]), wx);
            gimp_1d<GRAD, DIM == 2>(xi[1], lp[1], 1. / (4. * lp[1]), wy);
        } else if (SHAPE == SHAPE_B2SPLINE) {
            bspline_1d<GRAD>(xi[0], wx);
            bspline_1d<GRAD>(xi[1], wy);
        } else {
            bgimp_1d<GRAD, DIM == 2>(xi[0], xi[0], lp[0], wx);
            bgimp_1d<GRAD, DIM == 2>(xi[1], xi[1], lp[1], wy);
        }
        double inv_dx = 0., inv_dy = 0., inv_dz = 0.;
        if (GRAD) {
            inv_dx = gnum / (g.xpts[c.i + 1] - g.xpts[c.i]);
            inv_dy = gnum / (g.ypts[c.j + 1] - g.ypts[c.j]);
        }
        if (DIM == 3) {
            if (SHAPE == SHAPE_UGIMP) gimp_1d<GRAD, false>(xi[2], lp[2], 1. / (4. * lp[1]), wz);      // reference quirk: lp.y
            else if (SHAPE == SHAPE_B2SPLINE) bspline_1d<GRAD>(xi[2], wz);
            else bgimp_1d<GRAD, false>(xi[2], xi[1], lp[2], wz);              // reference quirk: the sign of the z gradient looks at xi.y
            if (GRAD) inv_dz = gnum / (g.zpts[c.k + 1] - g.zpts[c.k]);
#pragma unroll
            for (int kz = 0; kz < 4; kz++) {
                if (!(wz.ok >> kz & 1u)) continue;
#pragma unroll
                for (int jy = 0; jy < 4; jy++) {
                    if (!(wy.ok >> jy & 1u)) continue;
#pragma unroll
                    for (int ix = 0; ix < 4; ix++) {
                        if (!(wx.ok >> ix & 1u)) continue;
                        const double S = wx.S[ix] * wy.S[jy] * wz.S[kz];
                        double gxv = 0., gyv = 0., gzv = 0.;
                        if (GRAD) {
                            gxv = wx.dS[ix] * wy.S[jy] * wz.S[kz] * inv_dx;
                            gyv = wx.S[ix] * wy.dS[jy] * wz.S[kz] * inv_dy;
                            gzv = wx.S[ix] * wy.S[jy] * wz.dS[kz] * inv_dz;
                        }
                        f(n0 + (ix - 1) + (jy - 1) * g.yplane + (kz - 1) * g.zplane, S, gxv, gyv, gzv);
                    }
                }
            }
        } else {
#pragma unroll
            for (int jy = 0; jy < 4; jy++) {
                if (!(wy.ok >> jy & 1u)) continue;
#pragma unroll
                for (int ix = 0; ix < 4; ix++) {
                    if (!(wx.ok >> ix & 1u)) continue;
                    const double S = wx.S[ix] * wy.S[jy];
                    double gxv = 0., gyv = 0.;
                    if (GRAD) {
                        gxv = wx.dS[ix] * wy.S[jy] * inv_dx;
                        gyv = wx.S[ix] * wy.dS[jy] * inv_dy;
                    }
                    f(n0 + (ix - 1) + (jy - 1) * g.yplane, S, gxv, gyv, 0.);
                }
            }
        }
    }
}


// ---- CPDI (lCPDI 2D/3D, qCPDI 2D) ------------------------------------------------------------------------
// Domain set-up once per step: MatPoint3D::GetCPDINodesAndWeights (MPM_Classes/MatPoint3D.cpp:538-621) with
// GetSemiSideVectors (:413-433) and ScaleSemiSideVectorsForCPDI (:436-492); 2D: MatPoint2D.cpp:423-477,519-640.
// Returns false when a corner has left the grid (the reference throws, MatPoint3D.cpp:596-601).
template <int DIM, int SHAPE>
struct CpdiTraits { static const int NC = (DIM == 3 ? 8 : (SHAPE_IS_QCPDI(SHAPE) ? 9 : 4)); };

// LEAN (3D lCPDI with the merged walk in every kernel): a domain that fits the 3x3x3 window stores its 12 domain numbers only --
// nothing reads the 52 bytes per corner -- and cannot have a corner off the grid when the window is inside it.
template <int DIM, int SHAPE, bool LEAN = false>
__device__ __forceinline__ bool cpdi_setup(const Grid &g, const Particles &P, int p)
{
    const int NC = CpdiTraits<DIM, SHAPE>::NC;
    const int e = P.elem[p];
    const ElemIJK c0 = elem_ijk(g, e);
    const double cx = g.xpts[c0.i + 1] - g.xpts[c0.i], cy = g.ypts[c0.j + 1] - g.ypts[c0.j];
    const double cz = DIM == 3 ? g.zpts[c0.k + 1] - g.zpts[c0.k] : 1.;
    const double psx = cx * (0.5 * P.lp[0][p]), psy = cy * (0.5 * P.lp[1][p]), psz = DIM == 3 ? cz * (0.5 * P.lp[2][p]) : 0.;
    const double pos[3] = {P.pos[0][p], P.pos[1][p], DIM == 3 ? P.pos[2][p] : 0.};
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = P.F[i][p];
    double r1[3] = {F[0] * psx, F[3] * psx, DIM == 3 ? F[6] * psx : 0.};
    double r2[3] = {F[1] * psy, F[4] * psy, DIM == 3 ? F[7] * psy : 0.};
    double r3[3] = {F[2] * psz, F[5] * psz, F[8] * psz};
    if (g.rcrit >= 0.) {
        if (DIM == 3) {
            const double rc = g.rcrit * fmin(cx, fmin(cy, cz));
            double l[4][3];
            const double sg[4][2] = {{1., 1.}, {1., -1.}, {-1., 1.}, {-1., -1.}};     // la, lb, lc, ld: signs of r1, r2
            bool rescale = false;
#pragma unroll
            for (int a = 0; a < 4; a++) {
#pragma unroll
                for (int d = 0; d < 3; d++) l[a][d] = sg[a][0] * r1[d] + sg[a][1] * r2[d] + r3[d];
                const double mag = sqrt(l[a][0] * l[a][0] + l[a][1] * l[a][1] + l[a][2] * l[a][2]);
                if (mag > rc) {
                    const double sc = rc / mag;
                    l[a][0] *= sc; l[a][1] *= sc; l[a][2] *= sc;
                    rescale = true;
                }
            }
            if (rescale) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    r1[d] = 0.25 * (l[0][d] + l[1][d] - l[2][d] - l[3][d]);
                    r2[d] = 0.25 * (l[0][d] - l[1][d] + l[2][d] - l[3][d]);
                    r3[d] = 0.25 * (l[0][d] + l[1][d] + l[2][d] + l[3][d]);
                }
            }
        } else {
            const double rc = g.rcrit * fmin(cx, cy);
            double la[2] = {r1[0] + r2[0], r1[1] + r2[1]}, lb[2] = {r1[0] - r2[0], r1[1] - r2[1]};
            bool rescale = false;
            const double lam = sqrt(la[0] * la[0] + la[1] * la[1]), lbm = sqrt(lb[0] * lb[0] + lb[1] * lb[1]);
            if (lam > rc) { la[0] *= rc / lam; la[1] *= rc / lam; rescale = true; }
            if (lbm > rc) { lb[0] *= rc / lbm; lb[1] *= rc / lbm; rescale = true; }
            if (rescale) {
                r1[0] = 0.5 * (la[0] + lb[0]); r1[1] = 0.5 * (la[1] + lb[1]);
                r2[0] = 0.5 * (la[0] - lb[0]); r2[1] = 0.5 * (la[1] - lb[1]);
            }
        }
    }
    if (DIM == 3 && P.cpDom) {          // the domain in grid units for the merged walk (for_each_node_lcpdi3_hat)
        const double ic[3] = {1. / g.gx, 1. / g.gy, 1. / g.gz}, mn[3] = {g.xmin, g.ymin, g.zmin};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            P.cpDom[(size_t)d * P.cpStride + p] = (pos[d] - mn[d]) * ic[d];
            P.cpDom[(size_t)(3 + d) * P.cpStride + p] = r1[d] * ic[d];
            P.cpDom[(size_t)(6 + d) * P.cpStride + p] = r2[d] * ic[d];
            P.cpDom[(size_t)(9 + d) * P.cpStride + p] = r3[d] * ic[d];
        }
        if (LEAN) {         // the test of for_each_node_lcpdi3_hat
            bool fits = true;
            const int nmax[3] = {g.horiz, g.vert, g.depth};
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double ud = (pos[d] - mn[d]) * ic[d], ext = fabs(r1[d] * ic[d]) + fabs(r2[d] * ic[d]) + fabs(r3[d] * ic[d]);
                const int lo = (int)floor(ud - ext);
                if ((int)floor(ud + ext) - lo > 1 || lo < 0 || lo + 2 > nmax[d]) fits = false;
            }
            if (fits) return true;
        }
    }
    // corner positions
    double cs[NC][3];
    if (DIM == 3) {
        const double s1[8] = {-1., 1., 1., -1., -1., 1., 1., -1.}, s2[8] = {-1., -1., 1., 1., -1., -1., 1., 1.}, s3[8] = {-1., -1., -1., -1., 1., 1., 1., 1.};
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int d = 0; d < 3; d++) cs[i][d] = pos[d] + s1[i] * r1[d] + s2[i] * r2[d] + s3[i] * r3[d];
    } else {
        const double s1[9] = {-1., 1., 1., -1., 0., 1., 0., -1., 0.}, s2[9] = {-1., -1., 1., 1., -1., 0., 1., 0., 0.};
#pragma unroll
        for (int i = 0; i < NC; i++) {
            // the reference writes pos -r1 -r2 etc. as chained subtractions/additions of r1 then r2
            double x = pos[0], y = pos[1];
            if (s1[i] != 0.) { x += s1[i] * r1[0]; y += s1[i] * r1[1]; }
            if (s2[i] != 0.) { x += s2[i] * r2[0]; y += s2[i] * r2[1]; }
            cs[i][0] = x; cs[i][1] = y; cs[i][2] = 0.;
        }
    }
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NC; i++) {
        int ce;
        if (DIM == 2 && SHAPE_IS_QCPDI(SHAPE) && i == 8) ce = e;
        else ce = find_element_from_point<DIM>(g, cs[i]);
        if (ce <= 0) { ok = false; ce = e; }
        double xi[3];
        get_xipos<DIM>(g, ce, cs[i], xi);
        P.cpElem[(size_t)i * P.cpStride + p] = ce;
        P.cpXi[(size_t)(3 * i) * P.cpStride + p] = xi[0];
        P.cpXi[(size_t)(3 * i + 1) * P.cpStride + p] = xi[1];
        P.cpXi[(size_t)(3 * i + 2) * P.cpStride + p] = xi[2];
    }
    // gradient weights
    double wg[NC][3];
    if (DIM == 3) {
        double Vp = 8. * (r1[0] * (r2[1] * r3[2] - r2[2] * r3[1]) + r1[1] * (r2[2] * r3[0] - r2[0] * r3[2]) + r1[2] * (r2[0] * r3[1] - r2[1] * r3[0]));
        Vp = 1. / Vp;
        const double r1x = r1[0], r1y = r1[1], r1z = r1[2], r2x = r2[0], r2y = r2[1], r2z = r2[2], r3x = r3[0], r3y = r3[1], r3z = r3[2];
        wg[0][0] = (r1z * r2y - r1y * r2z - r1z * r3y + r2z * r3y + r1y * r3z - r2y * r3z) * Vp;
        wg[0][1] = (-(r1z * r2x) + r1x * r2z + r1z * r3x - r2z * r3x - r1x * r3z + r2x * r3z) * Vp;
        wg[0][2] = (r1y * r2x - r1x * r2y - r1y * r3x + r2y * r3x + r1x * r3y - r2x * r3y) * Vp;
        wg[1][0] = (r1z * r2y - r1y * r2z - r1z * r3y - r2z * r3y + r1y * r3z + r2y * r3z) * Vp;
        wg[1][1] = (-(r1z * r2x) + r1x * r2z + r1z * r3x + r2z * r3x - r1x * r3z - r2x * r3z) * Vp;
        wg[1][2] = (r1y * r2x - r1x * r2y - r1y * r3x - r2y * r3x + r1x * r3y + r2x * r3y) * Vp;
        wg[2][0] = (r1z * r2y - r1y * r2z + r1z * r3y - r2z * r3y - r1y * r3z + r2y * r3z) * Vp;
        wg[2][1] = (-(r1z * r2x) + r1x * r2z - r1z * r3x + r2z * r3x + r1x * r3z - r2x * r3z) * Vp;
        wg[2][2] = (r1y * r2x - r1x * r2y + r1y * r3x - r2y * r3x - r1x * r3y + r2x * r3y) * Vp;
        wg[3][0] = (r1z * r2y - r1y * r2z + r1z * r3y + r2z * r3y - r1y * r3z - r2y * r3z) * Vp;
        wg[3][1] = (-(r1z * r2x) + r1x * r2z - r1z * r3x - r2z * r3x + r1x * r3z + r2x * r3z) * Vp;
        wg[3][2] = (r1y * r2x - r1x * r2y + r1y * r3x + r2y * r3x - r1x * r3y - r2x * r3y) * Vp;
        wg[4][0] = (-(r1z * r2y) + r1y * r2z - r1z * r3y + r2z * r3y + r1y * r3z - r2y * r3z) * Vp;
        wg[4][1] = (r1z * r2x - r1x * r2z + r1z * r3x - r2z * r3x - r1x * r3z + r2x * r3z) * Vp;
        wg[4][2] = (-(r1y * r2x) + r1x * r2y - r1y * r3x + r2y * r3x + r1x * r3y - r2x * r3y) * Vp;
        wg[5][0] = (-(r1z * r2y) + r1y * r2z - r1z * r3y - r2z * r3y + r1y * r3z + r2y * r3z) * Vp;
        wg[5][1] = (r1z * r2x - r1x * r2z + r1z * r3x + r2z * r3x - r1x * r3z - r2x * r3z) * Vp;
        wg[5][2] = (-(r1y * r2x) + r1x * r2y - r1y * r3x - r2y * r3x + r1x * r3y + r2x * r3y) * Vp;
        wg[6][0] = (-(r1z * r2y) + r1y * r2z + r1z * r3y - r2z * r3y - r1y * r3z + r2y * r3z) * Vp;
        wg[6][1] = (r1z * r2x - r1x * r2z - r1z * r3x + r2z * r3x + r1x * r3z - r2x * r3z) * Vp;
        wg[6][2] = (-(r1y * r2x) + r1x * r2y + r1y * r3x - r2y * r3x - r1x * r3y + r2x * r3y) * Vp;
        wg[7][0] = (-(r1z * r2y) + r1y * r2z + r1z * r3y + r2z * r3y - r1y * r3z - r2y * r3z) * Vp;
        wg[7][1] = (r1z * r2x - r1x * r2z - r1z * r3x - r2z * r3x + r1x * r3z + r2x * r3z) * Vp;
        wg[7][2] = (-(r1y * r2x) + r1x * r2y + r1y * r3x + r2y * r3x - r1x * r3y - r2x * r3y) * Vp;
    } else {
        double Ap = 4. * (r1[0] * r2[1] - r1[1] * r2[0]);
        Ap = SHAPE_IS_QCPDI(SHAPE) ? 1. / (3. * Ap) : 1. / Ap;
        wg[0][0] = (r1[1] - r2[1]) * Ap; wg[0][1] = (-r1[0] + r2[0]) * Ap;
        wg[1][0] = (r1[1] + r2[1]) * Ap; wg[1][1] = (-r1[0] - r2[0]) * Ap;
        wg[2][0] = (-r1[1] + r2[1]) * Ap; wg[2][1] = (r1[0] - r2[0]) * Ap;
        wg[3][0] = (-r1[1] - r2[1]) * Ap; wg[3][1] = (r1[0] + r2[0]) * Ap;
        if (SHAPE_IS_QCPDI(SHAPE)) {
            wg[4 % NC][0] = 4. * r1[1] * Ap; wg[4 % NC][1] = -4. * r1[0] * Ap;
            wg[5 % NC][0] = 4. * r2[1] * Ap; wg[5 % NC][1] = -4. * r2[0] * Ap;
            wg[6 % NC][0] = -4. * r1[1] * Ap; wg[6 % NC][1] = 4. * r1[0] * Ap;
            wg[7 % NC][0] = -4. * r2[1] * Ap; wg[7 % NC][1] = 4. * r2[0] * Ap;
            wg[8 % NC][0] = 0.; wg[8 % NC][1] = 0.;
        }
#pragma unroll
        for (int i = 0; i < NC; i++) wg[i][2] = 0.;
    }
#pragma unroll
    for (int i = 0; i < NC; i++)
#pragma unroll
        for (int d = 0; d < 3; d++) P.cpWg[(size_t)(3 * i + d) * P.cpStride + p] = wg[i][d];
    return ok;
}

// ElementBase::GetCPDIFunctions (Elements/MoreMPMElementBase.cpp:581-657): node functions are the corner-weighted
// linear element functions, S_i = sum_c ws_c N_i(x_c), grad S_i = sum_c wg_c N_i(x_c); N < 1e-15 dropped (:612).
// The reference merges equal nodes before use; every consumer here is linear in (S, grad S), so the
// (corner, node) pairs are handed out unmerged.
template <int DIM, int SHAPE, bool GRAD, class F>
__device__ __forceinline__ void for_each_node_cpdi(const Grid &g, const Particles &P, int p, F &&f)
{
    const int NC = CpdiTraits<DIM, SHAPE>::NC;
#pragma unroll 1
    for (int c = 0; c < NC; c++) {
        const int ce = P.cpElem[(size_t)c * P.cpStride + p];
        const double xi = P.cpXi[(size_t)(3 * c) * P.cpStride + p], eta = P.cpXi[(size_t)(3 * c + 1) * P.cpStride + p];
        const double zeta = DIM == 3 ? P.cpXi[(size_t)(3 * c + 2) * P.cpStride + p] : 0.;
        double ws;
        if (DIM == 3) ws = 0.125;
        else if (SHAPE_IS_QCPDI(SHAPE)) ws = c < 4 ? 1. / 36. : (c < 8 ? 1. / 9. : 4. / 9.);
        else ws = 0.25;
        double wx = 0., wy = 0., wz = 0.;
        if (GRAD) {
            wx = P.cpWg[(size_t)(3 * c) * P.cpStride + p]; wy = P.cpWg[(size_t)(3 * c + 1) * P.cpStride + p];
            wz = DIM == 3 ? P.cpWg[(size_t)(3 * c + 2) * P.cpStride + p] : 0.;
        }
        const ElemIJK ec = elem_ijk(g, ce);
        const int n0 = elem_node0(g, ec);
        if (SHAPE == SHAPE_B2CPDI) {
            // B2CPDI: the corners are evaluated with the quadratic B-splines of the grid (ElementBase::GetShapeFunctionsForTractions
            // -> SplineShapeFunction, MoreMPMElementBase.cpp:90-105), 27 (9) nodes per corner
            Gimp1D sx, sy, sz;
            bspline_1d<false>(xi, sx);
            bspline_1d<false>(eta, sy);
            if (DIM == 3) bspline_1d<false>(zeta, sz); else { sz.ok = 2u; sz.S[1] = 1.; }
#pragma unroll 1
            for (int kz = 0; kz < 4; kz++) {
                if (!(sz.ok >> kz & 1u)) continue;
#pragma unroll 1
                for (int jy = 0; jy < 4; jy++) {
                    if (!(sy.ok >> jy & 1u)) continue;
#pragma unroll
                    for (int ix = 0; ix < 4; ix++) {
                        if (!(sx.ok >> ix & 1u)) continue;
                        const double N = DIM == 3 ? sx.S[ix] * sy.S[jy] * sz.S[kz] : sx.S[ix] * sy.S[jy];
                        if (N < 1e-15) continue;
                        f(n0 + (ix - 1) + (jy - 1) * g.yplane + (DIM == 3 ? (kz - 1) * g.zplane : 0), ws * N, wx * N, wy * N, DIM == 3 ? wz * N : 0.);
                    }
                }
            }
        } else if (DIM == 3) {
            const int xo[8] = {0, 1, 1, 0, 0, 1, 1, 0}, yo[8] = {0, 0, 1, 1, 0, 0, 1, 1}, zo[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const double t1 = 1. + (xo[a] ? 1. : -1.) * xi, t2 = 1. + (yo[a] ? 1. : -1.) * eta, t3 = 1. + (zo[a] ? 1. : -1.) * zeta;
                const double N = 0.125 * t1 * t2 * t3;
                if (N < 1e-15) continue;
                f(n0 + xo[a] + yo[a] * g.yplane + zo[a] * g.zplane, ws * N, wx * N, wy * N, wz * N);
            }
        } else {
            const int xo[4] = {0, 1, 1, 0}, yo[4] = {0, 0, 1, 1};
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double t1 = 1. + (xo[a] ? 1. : -1.) * xi, t2 = 1. + (yo[a] ? 1. : -1.) * eta;
                const double N = 0.25 * t1 * t2;
                if (N < 1e-15) continue;
                f(n0 + xo[a] + yo[a] * g.yplane, ws * N, wx * N, wy * N, 0.);
            }
        }
    }
}

// ---- 3D lCPDI with the corners merged per node, in registers ---------------------------------------------------------
// The shape function of corner c at node n of its element is the trilinear hat  N_n(x_c) = prod_d max(0, 1 - |v_cd - m_nd|)
// in grid units (v = (x - min)/cell, m = integer node coordinates: the same polynomial as (1 +- xi)(1 +- eta)(1 +- zeta)/8 of
// EightNodeIsoparamBrick::ShapeFunction :87-105), and the corner gradient weights of MatPoint3D::GetCPDINodesAndWeights
// (:551-594) are  wg_c = (s1 r2xr3 + s2 r3xr1 + s3 r1xr2)/Vp  with the corner's signs (s1, s2, s3).  So
//     S_n = 1/8 sum_c N_n(x_c),     grad S_n = [ (r2xr3) sum_c s1 N_n + (r3xr1) sum_c s2 N_n + (r1xr2) sum_c s3 N_n ] / Vp:
// four scalar sums per node.  A domain no longer than a cell per axis touches at most 3 nodes per axis; the 3x3x3 window is
// walked row by row (three nodes at a time: 12 running sums, no thread-local array), each node is handed to f ONCE --
// 8 to 27 calls instead of 64, which is what the P2G kernels pay for in atomics and the G2P kernels in node reads -- and
// nothing of the per-corner data cpdi_setup stores (52 bytes per corner) is read.  Returns false for a domain that spans
// more than two cells along some axis (stretched): the caller then walks the stored corners.
template <bool GRAD, class F>
__device__ __forceinline__ bool for_each_node_lcpdi3_hat(const Grid &g, const Particles &P, int p, F &&f)
{
    // the domain of this step (position and semi-side vectors in grid units), stored by cpdi_setup: like the reference's corner data
    // it is made once per step from the state at the start of the step and does not follow the strain and position updates
    double u[3], a1[3], a2[3], a3[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        u[d] = P.cpDom[(size_t)d * P.cpStride + p];
        a1[d] = P.cpDom[(size_t)(3 + d) * P.cpStride + p]; a2[d] = P.cpDom[(size_t)(6 + d) * P.cpStride + p]; a3[d] = P.cpDom[(size_t)(9 + d) * P.cpStride + p];
    }
    const double r1[3] = {a1[0] * g.gx, a1[1] * g.gy, a1[2] * g.gz}, r2[3] = {a2[0] * g.gx, a2[1] * g.gy, a2[2] * g.gz},
                 r3[3] = {a3[0] * g.gx, a3[1] * g.gy, a3[2] * g.gz};
    int lo[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double ext = fabs(a1[d]) + fabs(a2[d]) + fabs(a3[d]);
        lo[d] = (int)floor(u[d] - ext);
        if ((int)floor(u[d] + ext) - lo[d] > 1) return false;
    }
    if (lo[0] < 0 || lo[1] < 0 || lo[2] < 0 || lo[0] + 2 > g.horiz || lo[1] + 2 > g.vert || lo[2] + 2 > g.depth) return false;
    // corner coordinates relative to the window's first node, corner order and signs of MatPoint3D.cpp:23-25
    double cx[8], cy[8], cz[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const double s1 = ((c + 1) & 2) ? 1. : -1., s2 = (c & 2) ? 1. : -1., s3 = (c & 4) ? 1. : -1.;       // s1: - + + - - + + -
        cx[c] = (u[0] - lo[0]) + s1 * a1[0] + s2 * a2[0] + s3 * a3[0];
        cy[c] = (u[1] - lo[1]) + s1 * a1[1] + s2 * a2[1] + s3 * a3[1];
        cz[c] = (u[2] - lo[2]) + s1 * a1[2] + s2 * a2[2] + s3 * a3[2];
    }
    double A[3] = {0., 0., 0.}, B[3] = {0., 0., 0.}, C[3] = {0., 0., 0.};
    if (GRAD) {
        const double c23[3] = {r2[1] * r3[2] - r2[2] * r3[1], r2[2] * r3[0] - r2[0] * r3[2], r2[0] * r3[1] - r2[1] * r3[0]};
        const double c31[3] = {r3[1] * r1[2] - r3[2] * r1[1], r3[2] * r1[0] - r3[0] * r1[2], r3[0] * r1[1] - r3[1] * r1[0]};
        const double c12[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
        const double iVp = 1. / (8. * (r1[0] * c23[0] + r1[1] * c23[1] + r1[2] * c23[2]));
#pragma unroll
        for (int d = 0; d < 3; d++) { A[d] = c23[d] * iVp; B[d] = c31[d] * iVp; C[d] = c12[d] * iVp; }
    }
    const int n0 = lo[2] * g.zplane + lo[1] * g.yplane + lo[0];
#pragma unroll 1
    for (int kz = 0; kz < 3; kz++) {
#pragma unroll 1
        for (int jy = 0; jy < 3; jy++) {
            double s0[3] = {0., 0., 0.}, t1[3] = {0., 0., 0.}, t2[3] = {0., 0., 0.}, t3[3] = {0., 0., 0.};
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double hy = fmax(0., 1. - fabs(cy[c] - (double)jy)), hz = fmax(0., 1. - fabs(cz[c] - (double)kz));
                const double yz = hy * hz;
#pragma unroll
                for (int ix = 0; ix < 3; ix++) {
                    const double w = fmax(0., 1. - fabs(cx[c] - (double)ix)) * yz;
                    s0[ix] += w;
                    if (GRAD) {
                        if ((c + 1) & 2) t1[ix] += w; else t1[ix] -= w;
                        if (c & 2) t2[ix] += w; else t2[ix] -= w;
                        if (c & 4) t3[ix] += w; else t3[ix] -= w;
                    }
                }
            }
#pragma unroll
            for (int ix = 0; ix < 3; ix++) {
                if (s0[ix] < 1e-15) continue;           // (the reference drops corner weights below 1e-15, MoreMPMElementBase.cpp:618)
                const int nd = n0 + ix + jy * g.yplane + kz * g.zplane;
                if (GRAD) f(nd, 0.125 * s0[ix], A[0] * t1[ix] + B[0] * t2[ix] + C[0] * t3[ix], A[1] * t1[ix] + B[1] * t2[ix] + C[1] * t3[ix],
                            A[2] * t1[ix] + B[2] * t2[ix] + C[2] * t3[ix]);
                else f(nd, 0.125 * s0[ix], 0., 0., 0.);
            }
        }
    }
    return true;
}

// The same node set with the corners' contributions to one node merged before f sees them.  The corners of a domain lie in
// at most two elements per axis unless the domain is stretched beyond a cell, i.e. on a window of three nodes per axis anchored
// at the lowest corner element.  The weights are accumulated per window node in thread-local memory and f is called once per
// touched node -- 8 to 27 calls instead of 64 in 3D (8 when the domain sits inside one element, as on the undeformed lattice).
// f is linear in (S, gx, gy, gz) at every call site (P2G adds, G2P sums), so only the summation order changes -- towards
// the reference's, which compacts duplicate nodes in ElementBase::GetCPDIFunctions (MoreMPMElementBase.cpp:581-657).
// A corner outside the window (stretched domain) goes to f directly, as in for_each_node_cpdi.
template <int DIM, int SHAPE, bool GRAD, class F>
__device__ __forceinline__ void for_each_node_cpdi_merged(const Grid &g, const Particles &P, int p, F &&f)
{
    if constexpr (DIM == 3 && SHAPE == SHAPE_LCPDI_MERGED) {
        // 3D: the register-only walk of the window; a stretched domain (more than two cells along an axis) walks its corners
        if (!for_each_node_lcpdi3_hat<GRAD>(g, P, p, f)) for_each_node_cpdi<DIM, SHAPE, GRAD>(g, P, p, f);
        return;
    }
    const int NC = CpdiTraits<DIM, SHAPE>::NC;
    const int WN = DIM == 3 ? 27 : 9;
    int i0 = 0x7fffffff, j0 = 0x7fffffff, k0 = DIM == 3 ? 0x7fffffff : 0;
#pragma unroll 1
    for (int c = 0; c < NC; c++) {
        const ElemIJK ec = elem_ijk(g, P.cpElem[(size_t)c * P.cpStride + p]);
        i0 = ec.i < i0 ? ec.i : i0;
        j0 = ec.j < j0 ? ec.j : j0;
        if (DIM == 3) k0 = ec.k < k0 ? ec.k : k0;
    }
    double wS[WN], wX[GRAD ? WN : 1], wY[GRAD ? WN : 1], wZ[(GRAD && DIM == 3) ? WN : 1];
#pragma unroll
    for (int i = 0; i < WN; i++) {
        wS[i] = 0.;
        if (GRAD) { wX[i] = 0.; wY[i] = 0.; if (DIM == 3) wZ[i] = 0.; }
    }
#pragma unroll 1
    for (int c = 0; c < NC; c++) {
        const int ce = P.cpElem[(size_t)c * P.cpStride + p];
        const double xi = P.cpXi[(size_t)(3 * c) * P.cpStride + p], eta = P.cpXi[(size_t)(3 * c + 1) * P.cpStride + p];
        const double zeta = DIM == 3 ? P.cpXi[(size_t)(3 * c + 2) * P.cpStride + p] : 0.;
        double ws;
        if (DIM == 3) ws = 0.125;
        else if (SHAPE_IS_QCPDI(SHAPE)) ws = c < 4 ? 1. / 36. : (c < 8 ? 1. / 9. : 4. / 9.);
        else ws = 0.25;
        double wx = 0., wy = 0., wz = 0.;
        if (GRAD) {
            wx = P.cpWg[(size_t)(3 * c) * P.cpStride + p]; wy = P.cpWg[(size_t)(3 * c + 1) * P.cpStride + p];
            wz = DIM == 3 ? P.cpWg[(size_t)(3 * c + 2) * P.cpStride + p] : 0.;
        }
        const ElemIJK ec = elem_ijk(g, ce);
        const int di = ec.i - i0, dj = ec.j - j0, dk = DIM == 3 ? ec.k - k0 : 0;
        const bool inside = di <= 1 && dj <= 1 && dk <= 1;
        const int n0 = elem_node0(g, ec);
        const int NA = DIM == 3 ? 8 : 4;
        const int xo[8] = {0, 1, 1, 0, 0, 1, 1, 0}, yo[8] = {0, 0, 1, 1, 0, 0, 1, 1}, zo[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
        for (int a = 0; a < NA; a++) {
            const double t1 = 1. + (xo[a] ? 1. : -1.) * xi, t2 = 1. + (yo[a] ? 1. : -1.) * eta;
            double N;
            if (DIM == 3) { const double t3 = 1. + (zo[a] ? 1. : -1.) * zeta; N = 0.125 * t1 * t2 * t3; }
            else N = 0.25 * t1 * t2;
            if (N < 1e-15) continue;
            if (inside) {
                const int idx = (di + xo[a]) + 3 * (dj + yo[a]) + (DIM == 3 ? 9 * (dk + zo[a]) : 0);
                wS[idx] += ws * N;
                if (GRAD) { wX[idx] += wx * N; wY[idx] += wy * N; if (DIM == 3) wZ[idx] += wz * N; }
            } else {
                f(n0 + xo[a] + yo[a] * g.yplane + (DIM == 3 ? zo[a] * g.zplane : 0), ws * N, wx * N, wy * N, DIM == 3 ? wz * N : 0.);
            }
        }
    }
    const int nbase = k0 * g.zplane + j0 * g.yplane + i0;
#pragma unroll 1
    for (int idx = 0; idx < WN; idx++) {
        if (wS[idx] == 0.) continue;
        const int ix = idx % 3, iy = (idx / 3) % 3, iz = idx / 9;
        f(nbase + ix + iy * g.yplane + iz * g.zplane, wS[idx], GRAD ? wX[idx] : 0., GRAD ? wY[idx] : 0., (GRAD && DIM == 3) ? wZ[idx] : 0.);
    }
}
