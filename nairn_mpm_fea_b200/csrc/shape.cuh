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
    } else if (SHAPE == SHAPE_UGIMP) {
        Gimp1D wx, wy, wz;
        gimp_1d<GRAD, DIM == 2>(xi[0], lp[0], 1. / (4. * lp[0]), wx);
        gimp_1d<GRAD, DIM == 2>(xi[1], lp[1], 1. / (4. * lp[1]), wy);
        double inv_dx = 0., inv_dy = 0., inv_dz = 0.;
        if (GRAD) {
            inv_dx = 2.0 / (g.xpts[c.i + 1] - g.xpts[c.i]);
            inv_dy = 2.0 / (g.ypts[c.j + 1] - g.ypts[c.j]);
        }
        if (DIM == 3) {
            gimp_1d<GRAD, false>(xi[2], lp[2], 1. / (4. * lp[1]), wz);      // reference quirk: lp.y
            if (GRAD) inv_dz = 2.0 / (g.zpts[c.k + 1] - g.zpts[c.k]);
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
