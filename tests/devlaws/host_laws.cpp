// TEST INFRASTRUCTURE: the device constitutive laws of libmpmgpu (csrc/materials.cuh), compiled for the host through
// the stub in tests/devlaws/stub, behind one C entry point.  The test compares them with oracle/mpm_oracle.c on random
// states -- it checks the CUDA SOURCE without a GPU; the GPU parity tests check the compiled kernels.
#include "materials.cuh"
#include <string.h>

// temperature change handed to the laws (ResidualStrains::dT of the strain update)
static double g_dT = 0.;
extern "C" void devlaws_set_dT(double dT) { g_dT = dT; }

// state arrays are [component][n] as in mpmgpu_particles; F is [9][n] row-major components; du is [n][9]
extern "C" int devlaws_batch(int dim, int np, int kind, int nhist, const double *params, int n,
                             double *F, double *sp, double *pressure, double *eplast, double *energies /* work res heat entropy plast prevT */,
                             double *hist, const double *du, double delTime)
{
    Material m;
    m.kind = kind; m.nhist = nhist;
    memcpy(m.p, params, sizeof(double) * MPM_MAT_NPARAMS);
    for (int p = 0; p < n; p++) {
        PState s;
        s.dT = g_dT; s.dTad = 0.; s.adiabatic = 0;
        for (int i = 0; i < 9; i++) s.F[i] = F[(size_t)i * n + p];
        for (int i = 0; i < 6; i++) { s.sp[i] = sp[(size_t)i * n + p]; s.eplast[i] = eplast[(size_t)i * n + p]; }
        s.pressure = pressure[p];
        s.work = energies[p]; s.res = energies[(size_t)n + p]; s.heat = energies[(size_t)2 * n + p];
        s.entropy = energies[(size_t)3 * n + p]; s.plast = energies[(size_t)4 * n + p]; s.prevT = energies[(size_t)5 * n + p];
        for (int i = 0; i < MPM_MAX_HISTORY; i++) s.hist[i] = hist[(size_t)i * n + p];
        const double *d = du + (size_t)9 * p;
        if (dim == 3) constitutive_law_lr<3>(s, d, delTime, np, m); else constitutive_law_lr<2>(s, d, delTime, np, m);
        for (int i = 0; i < 9; i++) F[(size_t)i * n + p] = s.F[i];
        for (int i = 0; i < 6; i++) { sp[(size_t)i * n + p] = s.sp[i]; eplast[(size_t)i * n + p] = s.eplast[i]; }
        pressure[p] = s.pressure;
        energies[p] = s.work; energies[(size_t)n + p] = s.res; energies[(size_t)2 * n + p] = s.heat;
        energies[(size_t)3 * n + p] = s.entropy; energies[(size_t)4 * n + p] = s.plast;
        for (int i = 0; i < MPM_MAX_HISTORY; i++) hist[(size_t)i * n + p] = s.hist[i];
    }
    return 0;
}

// the plain dispatch (what every kernel except k_update_strains_lr calls): must agree with the _lr dispatch when p[7] = 0
extern "C" int devlaws_plain_one(int dim, int np, int kind, const double *params, double *F, double *sp, double *pressure, double *eplast,
                                 double *energies, double *hist, const double *du, double delTime)
{
    Material m;
    m.kind = kind; m.nhist = 0;
    memcpy(m.p, params, sizeof(double) * MPM_MAT_NPARAMS);
    PState s;
    s.dT = g_dT; s.dTad = 0.; s.adiabatic = 0;
    for (int i = 0; i < 9; i++) s.F[i] = F[i];
    for (int i = 0; i < 6; i++) { s.sp[i] = sp[i]; s.eplast[i] = eplast[i]; }
    s.pressure = *pressure;
    s.work = energies[0]; s.res = energies[1]; s.heat = energies[2]; s.entropy = energies[3]; s.plast = energies[4]; s.prevT = energies[5];
    for (int i = 0; i < MPM_MAX_HISTORY; i++) s.hist[i] = hist[i];
    if (dim == 3) constitutive_law<3>(s, du, delTime, np, m); else constitutive_law<2>(s, du, delTime, np, m);
    for (int i = 0; i < 9; i++) F[i] = s.F[i];
    for (int i = 0; i < 6; i++) { sp[i] = s.sp[i]; eplast[i] = s.eplast[i]; }
    *pressure = s.pressure;
    energies[0] = s.work; energies[1] = s.res; energies[2] = s.heat; energies[3] = s.entropy; energies[4] = s.plast;
    for (int i = 0; i < MPM_MAX_HISTORY; i++) hist[i] = s.hist[i];
    return 0;
}

// ---- archive records and global summands (csrc/archive.cuh) on the host -----------------------------------------------
#include "archive.cuh"

static void bind_host_particles(Particles &P, int n, double *pos, double *vel, double *mp, double *F, double *sp, double *pressure,
                                double *eplast, double *energies, double *hist, int *elem, int *mat0, int *cross)
{
    memset(&P, 0, sizeof P);
    P.n = n; P.nNR = n;
    for (int c = 0; c < 3; c++) { P.pos[c] = pos + (size_t)c * n; P.vel[c] = vel + (size_t)c * n; }
    P.mp = mp;
    for (int i = 0; i < 9; i++) P.F[i] = F + (size_t)i * n;
    for (int i = 0; i < 6; i++) { P.sp[i] = sp + (size_t)i * n; P.eplast[i] = eplast + (size_t)i * n; }
    P.pressure = pressure;
    P.work = energies; P.res = energies + n; P.heat = energies + (size_t)2 * n; P.entropy = energies + (size_t)3 * n;
    P.plast = energies + (size_t)4 * n; P.prevT = energies + (size_t)5 * n;
    for (int i = 0; i < MPM_MAX_HISTORY; i++) P.hist[i] = hist + (size_t)i * n;
    P.elem = elem; P.mat = mat0; P.cross = cross;
}

static void make_materials(Material *m, int nmat, const int *kinds, const double *params)
{
    for (int i = 0; i < nmat; i++) { m[i].kind = kinds[i]; m[i].nhist = 0; memcpy(m[i].p, params + (size_t)i * MPM_MAT_NPARAMS, sizeof(double) * MPM_MAT_NPARAMS); }
}

// returns the record size in bytes (or -1); out must hold n records
extern "C" int devarch_records(int dim, int n, const char *order, double *pos, double *vel, double *mp, double *F, double *sp, double *pressure,
                               double *eplast, double *energies, double *hist, int *elem, int *mat0, int *cross, int nmat, const int *kinds,
                               const double *params, const double *origpos, const double *angles0, double thickness, unsigned char *out)
{
    Particles P;
    bind_host_particles(P, n, pos, vel, mp, F, sp, pressure, eplast, energies, hist, elem, mat0, cross);
    Material mats[MPM_MAX_MATERIALS];
    make_materials(mats, nmat, kinds, params);
    ArchiveLayout L;
    const int rec = archive_layout_from_order(order, dim, L);
    if (rec < 0 || out == NULL) return rec;
    L.thickness = thickness; L.origpos = origpos; L.angles0 = angles0; L.stride = n;
    for (int p = 0; p < n; p++) archive_record(P, p, p, mats, L, (uint32_t *)out + (size_t)p * L.recWords);
    return rec;
}

// sums[m][GS_NSUMS], particle order
extern "C" int devarch_global_sums(int dim, int n, double *pos, double *vel, double *mp, double *F, double *sp, double *pressure, double *eplast,
                                   double *energies, double *hist, int *elem, int *mat0, int *cross, int nmat, const int *kinds,
                                   const double *params, double *sums)
{
    Particles P;
    bind_host_particles(P, n, pos, vel, mp, F, sp, pressure, eplast, energies, hist, elem, mat0, cross);
    Material mats[MPM_MAX_MATERIALS];
    make_materials(mats, nmat, kinds, params);
    for (int i = 0; i < nmat * GS_NSUMS; i++) sums[i] = 0.;
    for (int p = 0; p < n; p++) {
        double q[GS_NSUMS];
        global_summands(P, p, mats[mat0[p]], dim, q);
        for (int k = 0; k < GS_NSUMS; k++) sums[mat0[p] * GS_NSUMS + k] += q[k];
    }
    return GS_NSUMS;
}

// ---- CPDI node iteration (csrc/shape.cuh): plain (one call per corner node) against merged (one call per node) ----------
#include <vector>
#include "shape.cuh"

// Calls the device iteration for particle p and accumulates what f receives per node into out[4][nnodes] (S, gx, gy, gz) and the
// number of calls into calls[nnodes].  shape: SHAPE_LCPDI / SHAPE_QCPDI; merged selects the _MERGED variant.
template <int DIM, int SHAPE>
static void run_cpdi(const Grid &g, const Particles &P, int n, double *out, int *calls)
{
    const size_t nn = (size_t)g.nnodes;
    for (int p = 0; p < n; p++) {
        auto f = [&](int nd, double S, double gx, double gy, double gz) {
            out[nd] += S; out[nn + nd] += gx; out[2 * nn + nd] += gy; out[3 * nn + nd] += gz; calls[nd]++;
        };
        if (SHAPE_IS_MERGED(SHAPE)) for_each_node_cpdi_merged<DIM, SHAPE, true>(g, P, p, f);
        else for_each_node_cpdi<DIM, SHAPE, true>(g, P, p, f);
    }
}

extern "C" int devshape_cpdi_nodes(int dim, int shape, int merged, int horiz, int vert, int depth, int n, int *cpElem, double *cpXi, double *cpWg,
                                   double *out, int *calls)
{
    Grid g;
    memset(&g, 0, sizeof g);
    g.dim = dim; g.horiz = horiz; g.vert = vert; g.depth = dim == 3 ? depth : 1;
    g.yplane = horiz + 1; g.zplane = (horiz + 1) * (vert + 1);
    g.nnodes = g.zplane * (dim == 3 ? depth + 1 : 1);
    g.nelems = g.horiz * g.vert * g.depth;
    Particles P;
    memset(&P, 0, sizeof P);
    P.n = n; P.nNR = n; P.cpElem = cpElem; P.cpXi = cpXi; P.cpWg = cpWg; P.cpStride = (size_t)n;
    if (dim == 3 && shape == SHAPE_LCPDI) { if (merged) run_cpdi<3, SHAPE_LCPDI_MERGED>(g, P, n, out, calls); else run_cpdi<3, SHAPE_LCPDI>(g, P, n, out, calls); }
    else if (dim == 2 && shape == SHAPE_LCPDI) { if (merged) run_cpdi<2, SHAPE_LCPDI_MERGED>(g, P, n, out, calls); else run_cpdi<2, SHAPE_LCPDI>(g, P, n, out, calls); }
    else if (dim == 2 && shape == SHAPE_QCPDI) { if (merged) run_cpdi<2, SHAPE_QCPDI_MERGED>(g, P, n, out, calls); else run_cpdi<2, SHAPE_QCPDI>(g, P, n, out, calls); }
    else return -1;
    return 0;
}

// 3D lCPDI on REAL domains: the merged iteration derives the corners from pos, F and lp itself (for_each_node_lcpdi3_hat), so the two
// iterations are compared on particles whose corner data cpdi_setup has just made.  pos[3][n], F[9][n], lp[3][n]; out*[4][nnodes].
extern "C" int devshape_cpdi3_particles(int horiz, int vert, int depth, const double *xpts, const double *ypts, const double *zpts, double rcrit,
                                        int n, double *pos, double *F, double *lp, int *elem, double *outPlain, int *callsPlain,
                                        double *outMerged, int *callsMerged, int *fallbacks)
{
    Grid g;
    memset(&g, 0, sizeof g);
    g.dim = 3; g.np = 12; g.horiz = horiz; g.vert = vert; g.depth = depth;
    g.yplane = horiz + 1; g.zplane = (horiz + 1) * (vert + 1);
    g.nnodes = g.zplane * (depth + 1);
    g.nelems = horiz * vert * depth;
    g.xpts = xpts; g.ypts = ypts; g.zpts = zpts;
    g.xmin = xpts[0]; g.ymin = ypts[0]; g.zmin = zpts[0];
    g.gx = xpts[1] - xpts[0]; g.gy = ypts[1] - ypts[0]; g.gz = zpts[1] - zpts[0];
    g.rcrit = rcrit;
    Particles P;
    memset(&P, 0, sizeof P);
    P.n = n; P.nNR = n; P.elem = elem;
    for (int c = 0; c < 3; c++) { P.pos[c] = pos + (size_t)c * n; P.lp[c] = lp + (size_t)c * n; }
    for (int c = 0; c < 9; c++) P.F[c] = F + (size_t)c * n;
    std::vector<int> ce((size_t)8 * n);
    std::vector<double> cxi((size_t)36 * n), cwg((size_t)24 * n);
    P.cpElem = ce.data(); P.cpXi = cxi.data(); P.cpWg = cwg.data(); P.cpStride = (size_t)n;
    P.cpDom = cxi.data() + (size_t)24 * n;
    const size_t nn = (size_t)g.nnodes;
    *fallbacks = 0;
    for (int p = 0; p < n; p++) {
        if (!cpdi_setup<3, SHAPE_LCPDI>(g, P, p)) return -2;
        auto fp = [&](int nd, double S, double gx, double gy, double gz) {
            outPlain[nd] += S; outPlain[nn + nd] += gx; outPlain[2 * nn + nd] += gy; outPlain[3 * nn + nd] += gz; callsPlain[nd]++;
        };
        for_each_node_cpdi<3, SHAPE_LCPDI, true>(g, P, p, fp);
        auto fm = [&](int nd, double S, double gx, double gy, double gz) {
            outMerged[nd] += S; outMerged[nn + nd] += gx; outMerged[2 * nn + nd] += gy; outMerged[3 * nn + nd] += gz; callsMerged[nd]++;
        };
        auto none = [](int, double, double, double, double) {};
        if (!for_each_node_lcpdi3_hat<true>(g, P, p, none)) (*fallbacks)++;
        for_each_node_cpdi_merged<3, SHAPE_LCPDI_MERGED, true>(g, P, p, fm);
    }
    return 0;
}

// ---- Linear / uGIMP shape functions and the element search of the device source ---------------------------------------
static void host_grid(Grid &g, int dim, int np, int horiz, int vert, int depth, const double *xpts, const double *ypts, const double *zpts,
                      double gx, double gy, double gz)
{
    memset(&g, 0, sizeof g);
    g.dim = dim; g.np = np; g.horiz = horiz; g.vert = vert; g.depth = dim == 3 ? depth : 1;
    g.yplane = horiz + 1; g.zplane = (horiz + 1) * (vert + 1);
    g.nnodes = g.zplane * (dim == 3 ? depth + 1 : 1);
    g.nelems = g.horiz * g.vert * g.depth;
    g.xpts = xpts; g.ypts = ypts; g.zpts = zpts;
    g.gx = gx; g.gy = gy; g.gz = gz;
    g.xmin = xpts[0]; g.ymin = ypts[0]; g.zmin = dim == 3 ? zpts[0] : 0.;
    g.rcrit = -1.;
}

template <int DIM, int SHAPE>
static void run_shape(const Grid &g, int n, const int *inElem, const double *ncpos, const double *lp, int *count, int *nds, double *fn,
                      double *xd, double *yd, double *zd)
{
    for (int p = 0; p < n; p++) {
        const double xi[3] = {ncpos[p], ncpos[(size_t)n + p], ncpos[(size_t)2 * n + p]};
        const double l[3] = {lp[p], lp[(size_t)n + p], lp[(size_t)2 * n + p]};
        int k = 0;
        const size_t o = (size_t)64 * p;
        for_each_node<DIM, SHAPE, true>(g, inElem[p], xi, l, [&](int nd, double S, double gxv, double gyv, double gzv) {
            nds[o + k] = nd; fn[o + k] = S; xd[o + k] = gxv; yd[o + k] = gyv; zd[o + k] = gzv; k++;
        });
        count[p] = k;
    }
}

extern "C" int devshape_nodes(int dim, int np, int shape, int horiz, int vert, int depth, const double *xpts, const double *ypts, const double *zpts,
                              double gx, double gy, double gz, int n, const int *inElem, const double *ncpos, const double *lp,
                              int *count, int *nds, double *fn, double *xd, double *yd, double *zd)
{
    Grid g;
    host_grid(g, dim, np, horiz, vert, depth, xpts, ypts, zpts, gx, gy, gz);
    if (dim == 3 && shape == SHAPE_LINEAR) run_shape<3, SHAPE_LINEAR>(g, n, inElem, ncpos, lp, count, nds, fn, xd, yd, zd);
    else if (dim == 3 && shape == SHAPE_UGIMP) run_shape<3, SHAPE_UGIMP>(g, n, inElem, ncpos, lp, count, nds, fn, xd, yd, zd);
    else if (dim == 2 && shape == SHAPE_LINEAR) run_shape<2, SHAPE_LINEAR>(g, n, inElem, ncpos, lp, count, nds, fn, xd, yd, zd);
    else if (dim == 2 && shape == SHAPE_UGIMP) run_shape<2, SHAPE_UGIMP>(g, n, inElem, ncpos, lp, count, nds, fn, xd, yd, zd);
    else return -1;
    return 0;
}

// x [n][3] -> elem[n] (0 = off grid), xi[n][3] natural coordinates in that element, inside[n] = PtInElement(elem, x)
extern "C" int devshape_find_element(int dim, int np, int horiz, int vert, int depth, const double *xpts, const double *ypts, const double *zpts,
                                     double gx, double gy, double gz, int n, const double *x, int *elem, double *xi, int *inside)
{
    Grid g;
    host_grid(g, dim, np, horiz, vert, depth, xpts, ypts, zpts, gx, gy, gz);
    for (int p = 0; p < n; p++) {
        const double *pos = x + (size_t)3 * p;
        const int e = dim == 3 ? find_element_from_point<3>(g, pos) : find_element_from_point<2>(g, pos);
        elem[p] = e;
        inside[p] = 0;
        if (e > 0) {
            if (dim == 3) { get_xipos<3>(g, e, pos, xi + (size_t)3 * p); inside[p] = pt_in_element<3>(g, e, pos) ? 1 : 0; }
            else { get_xipos<2>(g, e, pos, xi + (size_t)3 * p); inside[p] = pt_in_element<2>(g, e, pos) ? 1 : 0; }
        }
    }
    return 0;
}

// the hardening-law terms of the device source alone: out = {yield, K', K2'(fnp1), yield increment}
extern "C" int devlaws_hardening_terms(const double *params, double prevT, double alpint, double dalpha, double delTime, double fnp1, double *out, double pressure)
{
    Material m;
    m.kind = MAT_ISOPLASTICITY; m.nhist = 1;
    memcpy(m.p, params, sizeof(double) * MPM_MAT_NPARAMS);
    const HardProps h = hard_props(m, prevT, pressure);
    HardAlpha a;
    a.alpint = alpint; a.dalpha = dalpha;
    out[0] = hard_yield(m, h, delTime, a);
    out[1] = hard_kprime(m, h, delTime, a);
    out[2] = hard_k2prime(m, h, fnp1, delTime, a);
    out[3] = hard_yield_increment(m, h, delTime, a);
    return 0;
}
