// Constitutive updates behind the reference's MaterialBase::MPMConstitutiveLaw interface
// (Common/Materials/MaterialBase.hpp:164-165): given du = grad(v)*dt for one particle, advance its
// deformation gradient, stress, energies and history.
//
// The per-particle state is handed over in registers (PState); the caller (a G2P kernel) loads it
// from and stores it to the SoA arrays.
#pragma once
#include "mpm_types.cuh"

struct PState {
    double F[9];        // row-major deformation gradient
    double sp[6];       // xx,yy,zz,yz,xz,xy (specific stress)
    double pressure;
    double eplast[6];   // plastic strain / elastic B
    double work, res, heat, entropy, plast;
    double prevT;       // pPreviousTemperature
    double hist[MPM_MAX_HISTORY];
    double dTad;        // adiabatic mode: temperature rise this law call adds to the particle's buffer (MPMBase::buffer_dTad)
    int adiabatic;      // ConductionTask::adiabatic (<EnergyCoupling/>): heat stays on the particle as a temperature rise
    double dT;          // temperature change this strain update answers to (ResidualStrains::dT, already scaled for the pass:
                        // MPMBase::ScaledResidualStrains); 0 unless the particle temperatures change (conduction, a start off the
                        // stress-free temperature)
};

enum { XX = 0, YY = 1, ZZ = 2, YZ = 3, XZ = 4, XY = 5 };

// 3x3 helpers (row-major)
__device__ __forceinline__ void mat3_mul(const double a[9], const double b[9], double c[9])
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

__device__ __forceinline__ double mat3_det(const double m[9])
{
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[6] * m[5]) + m[2] * (m[3] * m[7] - m[6] * m[4]);
}

// MaterialBase::IncrementHeatEnergy (Materials/MaterialBaseMPM.cpp:982-1006): isothermal -- the heat leaves, the temperature
// stays -- or, with <EnergyCoupling/> (ConductionTask::adiabatic), adiabatic -- dTq0 + dPhi/Cv is buffered as a temperature rise that
// the next particle update applies, and only the dissipated part makes entropy
__device__ __forceinline__ void increment_heat_energy(PState &s, double Cv, double dTq0, double dPhi)
{
    if (s.adiabatic) {
        const double dTphi = dPhi / Cv;
        s.dTad += dTq0 + dTphi;
        s.entropy += Cv * log((s.prevT + dTphi) / s.prevT);      // MPMBase::AddEntropy(Cv, T1, T2)
        return;
    }
    double baseHeat = -Cv * dTq0;
    s.heat += baseHeat - dPhi;
    s.entropy += baseHeat / s.prevT;
}

// Elastic::HypoIncrementDeformation (Materials/ElasticMPM.cpp:392-403): F <- (I + du) F
// MaterialBase::GetArtificialViscosity (MaterialBaseMPM.cpp:1824-1829): Dkk = relative volume change rate (< 0), c = current
// wave speed; material slots p[3] on/off, p[4] avA1, p[5] avA2, p[6] average cell size (MeshInfo::GetAverageCellSize)
__device__ __forceinline__ double artificial_viscosity(double Dkk, double c, const Material &m)
{
    const double divuij = fabs(Dkk);
    const double dcell = m.p[6];
    return dcell * divuij * (m.p[4] * c + m.p[5] * dcell * divuij);
}

template <int DIM>
__device__ __forceinline__ void hypo_increment_deformation(PState &s, const double du[9])
{
    double dF[9], Fn[9];
#pragma unroll
    for (int i = 0; i < 9; i++) dF[i] = du[i];
    dF[0] += 1.; dF[4] += 1.; dF[8] += 1.;
    if (DIM == 3) {
        mat3_mul(dF, s.F, Fn);
#pragma unroll
        for (int i = 0; i < 9; i++) s.F[i] = Fn[i];
    } else {
        // 2D Matrix3 product: in-plane block and zz only (Common/System/Matrix3.cpp operator*=)
        double f00 = dF[0] * s.F[0] + dF[1] * s.F[3];
        double f01 = dF[0] * s.F[1] + dF[1] * s.F[4];
        double f10 = dF[3] * s.F[0] + dF[4] * s.F[3];
        double f11 = dF[3] * s.F[1] + dF[4] * s.F[4];
        s.F[0] = f00; s.F[1] = f01; s.F[3] = f10; s.F[4] = f11;
        s.F[8] = dF[8] * s.F[8];
    }
}

// MaterialBase::Hypo3DCalculations (Materials/MaterialBaseMPM.cpp:1026-1047)
__device__ __forceinline__ void hypo3d(double sp[6], double dwxy, double dwxz, double dwyz, const double dsig[6])
{
    double st[6];
    st[XX] = -dwxy * sp[XY] - dwxz * sp[XZ];
    st[YY] = dwxy * sp[XY] - dwyz * sp[YZ];
    st[ZZ] = dwxz * sp[XZ] + dwyz * sp[YZ];
    st[YZ] = 0.5 * (dwxy * sp[XZ] + dwxz * sp[XY] + dwyz * (sp[YY] - sp[ZZ]));
    st[XZ] = 0.5 * (-dwxy * sp[YZ] + dwxz * (sp[XX] - sp[ZZ]) + dwyz * sp[XY]);
    st[XY] = 0.5 * (dwxy * (sp[XX] - sp[YY]) - dwxz * sp[YZ] - dwyz * sp[XZ]);
#pragma unroll
    for (int i = 0; i < 6; i++) sp[i] += dsig[i] + st[i];
}

// MaterialBase::Hypo2DCalculations (Materials/MaterialBaseMPM.cpp:1012-1020)
__device__ __forceinline__ void hypo2d(double sp[6], double dwrotxy, double dsxx, double dsyy, double dtxy)
{
    double dnorm = dwrotxy * sp[XY];
    double dshear = 0.5 * dwrotxy * (sp[XX] - sp[YY]);
    sp[XX] += dsxx - dnorm;
    sp[YY] += dsyy + dnorm;
    sp[XY] += dtxy + dshear;
}

// ---- IsotropicMat, small rotation ----------------------------------------------------------
// 3D: IsotropicMat::SRConstitutiveLaw3D (Materials/MoreIsotropicMat.cpp:286-348)
// 2D: IsotropicMat::SRConstitutiveLaw2D (Materials/MoreIsotropicMat.cpp:185-279), plane strain / plane stress
// Residual (thermal/moisture) strain increments are zero on this path (no transport tasks, no
// thermal ramp: UpdateParticlesTask.cpp:248-251 gives res.dT = 0); generalized-plane doopse = 0.
template <int DIM>
__device__ __forceinline__ void isotropic_law(PState &s, const double du[9], int np, const Material &m)
{
    hypo_increment_deformation<DIM>(s, du);
    const double *q = m.p;
    const double gamma0 = q[20], Cv = q[1];
    if (DIM == 3) {
        const double dvxx = du[0], dvyy = du[4], dvzz = du[8];
        const double dgamxy = du[1] + du[3], dgamxz = du[2] + du[6], dgamyz = du[5] + du[7];
        const double dwrotxy = du[3] - du[1], dwrotxz = du[6] - du[2], dwrotyz = du[7] - du[5];
        const double dVoverV = dvxx + dvyy + dvzz;
        const double eres = q[19] * s.dT;                         // CTE3 dT (:307)
        const double dvxxeff = dvxx - eres, dvyyeff = dvyy - eres, dvzzeff = dvzz - eres;
        double st0[6];
#pragma unroll
        for (int i = 0; i < 6; i++) st0[i] = s.sp[i];
        double delsp[6];
        delsp[XX] = q[8] * dvxxeff + q[9] * dvyyeff + q[10] * dvzzeff;     // C11 C12 C13
        delsp[YY] = q[9] * dvxxeff + q[11] * dvyyeff + q[12] * dvzzeff;    // C12 C22 C23
        delsp[ZZ] = q[10] * dvxxeff + q[12] * dvyyeff + q[13] * dvzzeff;   // C13 C23 C33
        delsp[YZ] = q[14] * dgamyz;                               // C44
        delsp[XZ] = q[15] * dgamxz;                               // C55
        delsp[XY] = q[16] * dgamxy;                               // C66
        hypo3d(s.sp, dwrotxy, dwrotxz, dwrotyz, delsp);
        s.work += 0.5 * ((st0[XX] + s.sp[XX]) * dvxx + (st0[YY] + s.sp[YY]) * dvyy + (st0[ZZ] + s.sp[ZZ]) * dvzz +
                         (st0[YZ] + s.sp[YZ]) * dgamyz + (st0[XZ] + s.sp[XZ]) * dgamxz + (st0[XY] + s.sp[XY]) * dgamxy);
        s.res += 0.5 * (st0[XX] + s.sp[XX] + st0[YY] + s.sp[YY] + st0[ZZ] + s.sp[ZZ]) * eres;
        double dTq0 = -gamma0 * s.prevT * dVoverV;
        increment_heat_energy(s, Cv, dTq0, 0.);
    } else {
        const double dvxx = du[0], dvyy = du[4];
        const double dgam = du[1] + du[3];
        const double dwrotxy = du[3] - du[1];
        double dVoverV = dvxx + dvyy;
        const double eres = q[17] * s.dT, ezzres = q[19] * s.dT;      // CTE1 (reduced in plane strain) and CTE3 (:198-199); doopse = 0
        const double dvxxeff = dvxx - eres, dvyyeff = dvyy - eres;
        double st0[6];
#pragma unroll
        for (int i = 0; i < 6; i++) st0[i] = s.sp[i];
        const double c1 = q[8] * dvxxeff + q[9] * dvyyeff;      // C[1][1] C[1][2]
        const double c2 = q[9] * dvxxeff + q[11] * dvyyeff;     // C[1][2] C[2][2]
        const double c3 = q[16] * dgam;                   // C[3][3]
        hypo2d(s.sp, dwrotxy, c1, c2, c3);
        double workEnergy = 0.5 * ((st0[XX] + s.sp[XX]) * dvxx + (st0[YY] + s.sp[YY]) * dvyy + (st0[XY] + s.sp[XY]) * dgam);
        double resEnergy = 0.5 * (st0[XX] + s.sp[XX] + st0[YY] + s.sp[YY]) * ezzres;
        if (np == NP_PLANE_STRAIN) {
            s.sp[ZZ] += q[21] * (dvxx - ezzres) + q[22] * (dvyy - ezzres) + q[23] * (0. - ezzres);      // C[4][1] C[4][2] C[4][4] (:241)
            resEnergy += 0.5 * (st0[ZZ] + s.sp[ZZ]) * ezzres;
        } else {
            // plane stress: out-of-plane strain increment (MoreIsotropicMat.cpp:249-258)
            double dezz = q[21] * (dvxx - ezzres) + q[22] * (dvyy - ezzres) + ezzres;
            s.F[8] += dezz * s.F[8];                      // MPMBase::IncrementDeformationGradientZZ: ep.zz += dezz*(1+ep.zz) (MPMBase.cpp:637-639)
            workEnergy += 0.5 * (st0[ZZ] + s.sp[ZZ]) * dezz;
            resEnergy += 0.5 * (st0[ZZ] + s.sp[ZZ]) * ezzres;
            dVoverV += dezz;
        }
        s.work += workEnergy;
        s.res += resEnergy;
        double dTq0 = -gamma0 * s.prevT * dVoverV;
        increment_heat_energy(s, Cv, dTq0, 0.);
    }
}

// ---- small-strain / large-rotation option (Elastic::useLargeRotation, <largeRotation>1</largeRotation>) ---------------
// MaterialBase::LRGetStrainIncrement (Materials/MaterialBaseMPM.cpp:882-927) over Matrix3::Exponential, Eigenvalues,
// RightDecompose / LeftDecompose and RVoightRT (Common/System/Matrix3.cpp:312-384, 464-520, 564-740, 190-232).
// The 3D decomposition goes through the trigonometric eigenvalues of C = F^T F; for the strain increments of an explicit
// step (|du| << 1) those are dominated by cancellation, which makes the reference's own result reproducible to ~1e-6 of the
// increment only (tests/parity.py TOL_LR3D); the same formulas are used here so the behaviour is the reference's.
template <int DIM>
__device__ __forceinline__ void exp_du(const double du[9], double dF[9])
{
    if (DIM == 3) {
#pragma unroll
        for (int i = 0; i < 9; i++) dF[i] = du[i];
        dF[0] += 1.; dF[4] += 1.; dF[8] += 1.;
    } else {
        // two terms in 2D (System/StartOutput.cpp:116-120): alpha0 I + alpha1 du, zz separately
        const double c0 = du[1] * du[3] - du[0] * du[4], c1 = du[0] + du[4];
        const double beta1 = 0.5 * (c1 * 1. + 0.), beta0 = 0.5 * c0 * 1.;
        const double betaz = du[8] * (0.5 * du[8]);
        const double alpha0 = 1. + beta0, alpha1 = 1. + beta1;
#pragma unroll
        for (int i = 0; i < 9; i++) dF[i] = 0.;
        dF[0] = alpha0 + alpha1 * du[0]; dF[1] = alpha1 * du[1]; dF[3] = alpha1 * du[3]; dF[4] = alpha0 + alpha1 * du[4];
        dF[8] = 1. + du[8] + betaz;
    }
}

__device__ __forceinline__ void sym_eigenvalues3(const double m[9], double lam[3])
{
    const double de = m[1] * m[5], dd = m[1] * m[1], ee = m[5] * m[5], ff = m[2] * m[2];
    const double mm = m[0] + m[4] + m[8];
    const double c1 = (m[0] * m[4] + m[0] * m[8] + m[4] * m[8]) - (dd + ee + ff);
    const double c0 = m[8] * dd + m[0] * ee + m[4] * ff - m[0] * m[4] * m[8] - 2.0 * m[2] * de;
    const double pp = mm * mm - 3.0 * c1;
    const double q = mm * (pp - 1.5 * c1) - 13.5 * c0;
    const double sqrt_p = sqrt(fabs(pp));
    double phi = 27.0 * (0.25 * c1 * c1 * (pp - c1) + c0 * (q + 6.75 * c0));
    phi = (1.0 / 3.0) * atan2(sqrt(fabs(phi)), q);
    double sn, cs;
    sincos(phi, &sn, &cs);
    const double c = sqrt_p * cs, s = (1.0 / 1.73205080756887729352744634151) * sqrt_p * sn;
    lam[1] = (1.0 / 3.0) * (mm - c);
    lam[2] = lam[1] + s;
    lam[0] = lam[1] + c;
    lam[1] -= s;
}

// rotation of the polar decomposition: F = R U (LEFT false, through C = F^T F) or F = V R (LEFT true, through B = F F^T)
template <int DIM, bool LEFT>
__device__ __forceinline__ void polar_rotation(const double F[9], double R[9])
{
    if (DIM == 2) {
        double Fsum = F[0] + F[4], Fdif = F[1] - F[3];
        const double denom = sqrt(Fsum * Fsum + Fdif * Fdif);
        Fsum /= denom; Fdif /= denom;
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = 0.;
        R[0] = Fsum; R[1] = Fdif; R[3] = -Fdif; R[4] = Fsum; R[8] = 1.;
        return;
    }
    double Ft[9], C[9], C2[9], lam[3], U[9], Ui[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Ft[3 * i + j] = F[3 * j + i];
    if (LEFT) mat3_mul(F, Ft, C); else mat3_mul(Ft, F, C);
    mat3_mul(C, C, C2);
    sym_eigenvalues3(C, lam);
    const double l1 = sqrt(lam[0]), l2 = sqrt(lam[1]), l3 = sqrt(lam[2]);
    const double i1 = l1 + l2 + l3, i2 = l1 * l2 + l1 * l3 + l2 * l3, i3 = l1 * l2 * l3;
    const double d1 = 1. / (i1 * i2 - i3), c2 = -d1, c1 = (i1 * i1 - i2) * d1, cI = i1 * i3 * d1;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int a = (!LEFT && i > j) ? 3 * j + i : 3 * i + j;         // RightDecompose fills U from the upper triangle
            U[3 * i + j] = c2 * C2[a] + c1 * C[a] + (i == j ? cI : 0.);
        }
    const double b1 = 1. / i3, bU = -i1 * b1, bI = i2 * b1;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int a = (!LEFT && i > j) ? 3 * j + i : 3 * i + j;
            Ui[3 * i + j] = b1 * C[a] + bU * U[a] + (i == j ? bI : 0.);
        }
    if (LEFT) mat3_mul(Ui, F, R); else mat3_mul(F, Ui, R);
}

// Matrix3::RVoightRT: o = R t R^T for a Voight tensor (stress: shear not doubled; strain: engineering shear)
template <int DIM>
__device__ __forceinline__ void rotate_voight(const double m[9], const double t[6], bool stress, double o[6])
{
    const double ur = stress ? 2. : 1., ll = stress ? 1. : 2.;
    const double txx = t[XX], tyy = t[YY], tzz = t[ZZ], tyz = t[YZ], txz = t[XZ], txy = t[XY];     // o may alias t
    if (DIM == 2) {
        o[XX] = m[0] * m[0] * txx + m[1] * m[1] * tyy + ur * m[0] * m[1] * txy;
        o[YY] = m[3] * m[3] * txx + m[4] * m[4] * tyy + ur * m[4] * m[3] * txy;
        o[XY] = ll * (m[0] * m[3] * txx + m[4] * m[1] * tyy) + (m[1] * m[3] + m[0] * m[4]) * txy;
        o[ZZ] = tzz; o[YZ] = 0.; o[XZ] = 0.;
        return;
    }
    o[XX] = m[0] * m[0] * txx + m[1] * m[1] * tyy + m[2] * m[2] * tzz + ur * (m[1] * m[2] * tyz + m[0] * m[2] * txz + m[0] * m[1] * txy);
    o[YY] = m[3] * m[3] * txx + m[4] * m[4] * tyy + m[5] * m[5] * tzz + ur * (m[4] * m[5] * tyz + m[3] * m[5] * txz + m[4] * m[3] * txy);
    o[ZZ] = m[6] * m[6] * txx + m[7] * m[7] * tyy + m[8] * m[8] * tzz + ur * (m[8] * m[7] * tyz + m[8] * m[6] * txz + m[6] * m[7] * txy);
    o[YZ] = ll * (m[3] * m[6] * txx + m[4] * m[7] * tyy + m[8] * m[5] * tzz)
            + (m[5] * m[7] + m[4] * m[8]) * tyz + (m[5] * m[6] + m[3] * m[8]) * txz + (m[4] * m[6] + m[3] * m[7]) * txy;
    o[XZ] = ll * (m[0] * m[6] * txx + m[1] * m[7] * tyy + m[8] * m[2] * tzz)
            + (m[2] * m[7] + m[8] * m[1]) * tyz + (m[2] * m[6] + m[0] * m[8]) * txz + (m[1] * m[6] + m[0] * m[7]) * txy;
    o[XY] = ll * (m[0] * m[3] * txx + m[1] * m[4] * tyy + m[2] * m[5] * tzz)
            + (m[4] * m[2] + m[5] * m[1]) * tyz + (m[2] * m[3] + m[0] * m[5]) * txz + (m[1] * m[3] + m[0] * m[4]) * txy;
}

// LRGetStrainIncrement(CURRENT_CONFIGURATION): F <- exp(du) F; de = (dF - dR) F(n-1) Rn^T, dR = Rn Rn-1^T
template <int DIM>
__device__ __forceinline__ void lr_strain_increment(PState &s, const double du[9], double de[9], double dR[9])
{
    double F0[9], dF[9], F1[9], Rm[9], Rn[9], T[9], FR[9];
#pragma unroll
    for (int i = 0; i < 9; i++) F0[i] = s.F[i];
    if (DIM == 2) { F0[2] = 0.; F0[5] = 0.; F0[6] = 0.; F0[7] = 0.; }
    exp_du<DIM>(du, dF);
    mat3_mul(dF, F0, F1);
#pragma unroll
    for (int i = 0; i < 9; i++) s.F[i] = F1[i];
    polar_rotation<DIM, false>(F0, Rm);
    polar_rotation<DIM, true>(F1, Rn);
    // dR = Rn Rm^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) dR[3 * i + j] = Rn[3 * i] * Rm[3 * j] + Rn[3 * i + 1] * Rm[3 * j + 1] + Rn[3 * i + 2] * Rm[3 * j + 2];
#pragma unroll
    for (int i = 0; i < 9; i++) T[i] = dF[i] - dR[i];
    // FR = F0 Rn^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) FR[3 * i + j] = F0[3 * i] * Rn[3 * j] + F0[3 * i + 1] * Rn[3 * j + 1] + F0[3 * i + 2] * Rn[3 * j + 2];
    mat3_mul(T, FR, de);
}

// IsotropicMat::LRConstitutiveLaw (Materials/MoreIsotropicMat.cpp:54-174); residual strains are zero on this path
template <int DIM>
__device__ __forceinline__ void isotropic_lr_law(PState &s, const double du[9], int np, const Material &m)
{
    const double *q = m.p;
    const double gamma0 = q[20], Cv = q[1];
    double de[9], dR[9];
    lr_strain_increment<DIM>(s, du, de, dR);
    rotate_voight<DIM>(dR, s.sp, true, s.sp);
    const double dgamxy = de[1] + de[3];
    double dVoverV = de[0] + de[4], work;
    if (DIM == 3) {
        const double dgamyz = de[5] + de[7], dgamxz = de[2] + de[6];
        dVoverV += de[8];
        s.sp[XX] += q[8] * de[0] + q[9] * de[4] + q[10] * de[8];
        s.sp[YY] += q[9] * de[0] + q[11] * de[4] + q[12] * de[8];
        s.sp[ZZ] += q[10] * de[0] + q[12] * de[4] + q[13] * de[8];
        s.sp[YZ] += q[14] * dgamyz; s.sp[XZ] += q[15] * dgamxz; s.sp[XY] += q[16] * dgamxy;
        work = s.sp[XX] * de[0] + s.sp[YY] * de[4] + s.sp[ZZ] * de[8] + s.sp[YZ] * dgamyz + s.sp[XZ] * dgamxz + s.sp[XY] * dgamxy;
    } else {
        s.sp[XX] += q[8] * de[0] + q[9] * de[4];
        s.sp[YY] += q[9] * de[0] + q[11] * de[4];
        s.sp[XY] += q[16] * dgamxy;
        work = s.sp[XX] * de[0] + s.sp[YY] * de[4] + s.sp[XY] * dgamxy;
        if (np == NP_PLANE_STRAIN) {
            s.sp[ZZ] += q[21] * de[0] + q[22] * de[4];
        } else {
            const double dezz = q[21] * de[0] + q[22] * de[4];
            s.F[8] += dezz * s.F[8];
            work += s.sp[ZZ] * dezz;
            dVoverV += dezz;
        }
    }
    s.work += work;
    increment_heat_energy(s, Cv, -gamma0 * s.prevT * dVoverV, 0.);
}

// ---- Neohookean (MaterialID 28) ---------------------------------------------------------------
// Neohookean::MPMConstitutiveLaw (Materials/Neohookean.cpp:177-331) over HyperElastic::IncrementDeformation
// (Materials/HyperElastic.cpp:104-139) and GetVolumetricTerms (:171-204).  Elastic left Cauchy-Green tensor B
// lives in eplast, pressure separately, deviatoric Kirchhoff stress/rho0 in sp, history = {J, Jres}.
// dF = exp(du) to incrementalDefGradTerms terms: 1 in 3D, 2 in 2D (System/StartOutput.cpp:116-120;
// Matrix3::Exponential, Common/System/Matrix3.cpp:312-353).  Residual stretch increment is 1 (no thermal load);
// artificial viscosity is off (rejected at set-up when on).
template <int DIM>
__device__ __forceinline__ void neohookean_law(PState &s, const double du[9], double delTime, int np, const Material &m)
{
    const double Gsp = m.p[8], Ksp = m.p[9], Lamesp = m.p[10];
    const int UofJ = (int)m.p[11];
    const double gamma0 = m.p[13], Cv = m.p[1];
    double dF[9], detDf;
    double *B = s.eplast;       // xx,yy,zz,yz,xz,xy
    if (DIM == 3) {
#pragma unroll
        for (int i = 0; i < 9; i++) dF[i] = du[i];
        dF[0] += 1.; dF[4] += 1.; dF[8] += 1.;
        double Fn[9];
        mat3_mul(dF, s.F, Fn);
#pragma unroll
        for (int i = 0; i < 9; i++) s.F[i] = Fn[i];
        const double Bm[9] = {B[XX], B[XY], B[XZ], B[XY], B[YY], B[YZ], B[XZ], B[YZ], B[ZZ]};
        double dB[9];
        mat3_mul(dF, Bm, dB);
        const double bxx = dB[0] * dF[0] + dB[1] * dF[1] + dB[2] * dF[2];
        const double bxy = dB[0] * dF[3] + dB[1] * dF[4] + dB[2] * dF[5];
        const double byy = dB[3] * dF[3] + dB[4] * dF[4] + dB[5] * dF[5];
        const double bzz = dB[6] * dF[6] + dB[7] * dF[7] + dB[8] * dF[8];
        const double bxz = dB[0] * dF[6] + dB[1] * dF[7] + dB[2] * dF[8];
        const double byz = dB[3] * dF[6] + dB[4] * dF[7] + dB[5] * dF[8];
        B[XX] = bxx; B[XY] = bxy; B[YY] = byy; B[ZZ] = bzz; B[XZ] = bxz; B[YZ] = byz;
        detDf = dF[0] * (dF[4] * dF[8] - dF[7] * dF[5]) - dF[3] * (dF[1] * dF[8] - dF[7] * dF[2]) + dF[6] * (dF[1] * dF[5] - dF[4] * dF[2]);
    } else {
        // two-term 2D exponential: alpha0 I + alpha1 m, zz separately
        const double c0 = du[1] * du[3] - du[0] * du[4], c1 = du[0] + du[4];
        const double beta1 = 0.5 * (c1 * 1. + 0.), beta0 = 0.5 * c0 * 1.;
        const double betaz = du[8] * (0.5 * du[8]);
        const double alpha0 = 1. + beta0, alpha1 = 1. + beta1, ezz = 1. + du[8] + betaz;
        const double d00 = alpha0 + alpha1 * du[0], d01 = alpha1 * du[1], d10 = alpha1 * du[3], d11 = alpha0 + alpha1 * du[4];
        const double f00 = d00 * s.F[0] + d01 * s.F[3], f01 = d00 * s.F[1] + d01 * s.F[4];
        const double f10 = d10 * s.F[0] + d11 * s.F[3], f11 = d10 * s.F[1] + d11 * s.F[4];
        s.F[0] = f00; s.F[1] = f01; s.F[3] = f10; s.F[4] = f11; s.F[8] = ezz * s.F[8];
        const double e00 = d00 * B[XX] + d01 * B[XY], e01 = d00 * B[XY] + d01 * B[YY];
        const double e10 = d10 * B[XX] + d11 * B[XY], e11 = d10 * B[XY] + d11 * B[YY];
        const double e22 = ezz * B[ZZ];
        B[XX] = e00 * d00 + e01 * d01;
        B[XY] = e00 * d10 + e01 * d11;
        B[YY] = e10 * d10 + e11 * d11;
        B[ZZ] = e22 * ezz;
        detDf = ezz * (d00 * d11 - d10 * d01);
    }
    double Jres = s.hist[1];
    const double dJres = s.dT != 0. ? exp(3. * m.p[12] * s.dT) : 1.;        // HyperElastic::GetIncrementalResJ (HyperElastic.cpp:89-95): exp(3 CTE1 dT)
    Jres *= dJres;
    s.hist[1] = Jres;
    const double resStretch = pow(Jres, 1. / 3.);
    const double Jres23 = resStretch * resStretch;
    if (DIM == 2 && np == NP_PLANE_STRESS) {
        const double arg = B[XX] * B[YY] - B[XY] * B[XY];
        double xn;
        if (UofJ == 1) {
            const double a = Lamesp * arg + Gsp * pow(Jres, 4. / 3.);
            const double b = Lamesp * sqrt(arg);
            xn = Jres * (b + sqrt(b * b + 4. * Gsp * a)) / (2. * a);
            xn *= xn;
        } else if (UofJ == 2) {
            xn = B[ZZ];
            const double J23 = pow(Jres, 2. / 3.);
            for (int iter = 1; iter < 20; iter++) {
                const double fx = Gsp * (xn - J23) + 0.5 * Lamesp * J23 * log(xn * arg / (Jres * Jres));
                const double fxp = Gsp + Lamesp * J23 / (2 * xn);
                const double xnp1 = xn - fx / fxp;
                if (fabs(xn - xnp1) < 1e-10) break;
                xn = xnp1;
            }
        } else {
            xn = Jres * Jres * (Lamesp + 2. * Gsp) / (Lamesp * arg + 2. * Gsp * pow(Jres, 4. / 3.));
        }
        const double dFzz = sqrt(xn / B[ZZ]);
        B[ZZ] = xn;
        s.F[8] = dFzz * s.F[8];          // ep.zz = dFzz*(1+ep.zz) - 1
        detDf *= dFzz;
    }
    const double J = detDf * s.hist[0];
    s.hist[0] = J;
    double st0[6];
#pragma unroll
    for (int i = 0; i < 6; i++) st0[i] = s.sp[i];
    const double Jeff = J / Jres;
    const double p0 = s.pressure;
    double Kterm;
    if (UofJ == 1) Kterm = Lamesp * (Jeff - 1.);
    else if (UofJ == 2) Kterm = Lamesp * log(Jeff) / Jeff;
    else Kterm = 0.5 * Lamesp * (Jeff - 1. / Jeff);
    const double Pterm = J * Kterm + Jres * Gsp * ((B[XX] + B[YY] + B[ZZ]) / (3. * Jres23) - 1.);
    const double delV = 1. - 1. / detDf;
    double QAVred = 0., AVEnergy = 0.;          // Neohookean.cpp:261-266
    if (delV < 0. && m.p[3] != 0.) {
        QAVred = artificial_viscosity(delV / delTime, sqrt(Ksp * J), m);
        AVEnergy = fabs(QAVred * delV);
    }
    const double Pfinal = -Pterm + QAVred;
    s.pressure = Pfinal;
    const double avgP = 0.5 * (p0 + Pfinal);
    const double dilEnergy = -avgP * delV;
    const double delVres = 1. - 1. / dJres;
    const double resEnergy = -avgP * delVres;
    const double GJeff = resStretch * Gsp;
    const double I1third = (B[XX] + B[YY] + B[ZZ]) / 3.;
    s.sp[XX] = GJeff * (B[XX] - I1third);
    s.sp[YY] = GJeff * (B[YY] - I1third);
    s.sp[ZZ] = GJeff * (B[ZZ] - I1third);
    s.sp[XY] = GJeff * B[XY];
    if (DIM == 3) { s.sp[XZ] = GJeff * B[XZ]; s.sp[YZ] = GJeff * B[YZ]; }
    double shearEnergy = 0.5 * ((s.sp[XX] + st0[XX]) * du[0] + (s.sp[YY] + st0[YY]) * du[4] + (s.sp[ZZ] + st0[ZZ]) * du[8] +
                                (s.sp[XY] + st0[XY]) * (du[1] + du[3]));
    if (DIM == 3) shearEnergy += 0.5 * ((s.sp[XZ] + st0[XZ]) * (du[2] + du[6]) + (s.sp[YZ] + st0[YZ]) * (du[5] + du[7]));
    s.work += dilEnergy + shearEnergy;
    s.res += resEnergy;
    const double Jres2third = pow(Jres, 2. / 3.);
    const double Gterm = Gsp * (3. - I1third / Jres2third) / (3. * Jeff);
    double Kratio;
    if (UofJ == 1) Kratio = Lamesp * Jeff + Gterm;
    else if (UofJ == 2) Kratio = Lamesp * (1 - log(Jeff)) / Jeff + Gterm;
    else Kratio = 0.5 * Lamesp * (Jeff + 1. / Jeff) + Gterm;
    Kratio /= Ksp;
    const double dTq0 = -J * Kratio * gamma0 * s.prevT * delV;
    increment_heat_energy(s, Cv, dTq0, AVEnergy);
}

// ---- Mooney (MaterialID 8) ----------------------------------------------------------------------------------------
// Mooney::MPMConstitutiveLaw (Materials/Mooney.cpp:184-365) over HyperElastic::IncrementDeformation (HyperElastic.cpp:104-139),
// GetVolumetricTerms (:171-204) and GetNewtonPressureTerms (:208-238).  Same state as Neohookean: elastic B in eplast,
// pressure apart, deviatoric Kirchhoff stress / rho0 in sp, history = {J, Jres}.  The IdealRubber option is refused at set-up.
template <int DIM>
__device__ __forceinline__ void mooney_law(PState &s, const double du[9], double delTime, int np, const Material &m)
{
    const double G1sp = m.p[8], G2sp = m.p[9], Ksp = m.p[10];
    const int UofJ = (int)m.p[11];
    const double gamma0 = m.p[13], Cv = m.p[1];
    double dF[9], detDf;
    double *B = s.eplast;       // xx,yy,zz,yz,xz,xy
    exp_du<DIM>(du, dF);
    if (DIM == 3) {
        double Fn[9];
        mat3_mul(dF, s.F, Fn);
#pragma unroll
        for (int i = 0; i < 9; i++) s.F[i] = Fn[i];
        const double Bm[9] = {B[XX], B[XY], B[XZ], B[XY], B[YY], B[YZ], B[XZ], B[YZ], B[ZZ]};
        double dB[9];
        mat3_mul(dF, Bm, dB);
        const double bxx = dB[0] * dF[0] + dB[1] * dF[1] + dB[2] * dF[2];
        const double bxy = dB[0] * dF[3] + dB[1] * dF[4] + dB[2] * dF[5];
        const double byy = dB[3] * dF[3] + dB[4] * dF[4] + dB[5] * dF[5];
        const double bzz = dB[6] * dF[6] + dB[7] * dF[7] + dB[8] * dF[8];
        const double bxz = dB[0] * dF[6] + dB[1] * dF[7] + dB[2] * dF[8];
        const double byz = dB[3] * dF[6] + dB[4] * dF[7] + dB[5] * dF[8];
        B[XX] = bxx; B[XY] = bxy; B[YY] = byy; B[ZZ] = bzz; B[XZ] = bxz; B[YZ] = byz;
        detDf = dF[0] * (dF[4] * dF[8] - dF[7] * dF[5]) - dF[3] * (dF[1] * dF[8] - dF[7] * dF[2]) + dF[6] * (dF[1] * dF[5] - dF[4] * dF[2]);
    } else {
        const double d00 = dF[0], d01 = dF[1], d10 = dF[3], d11 = dF[4], ezz = dF[8];
        const double f00 = d00 * s.F[0] + d01 * s.F[3], f01 = d00 * s.F[1] + d01 * s.F[4];
        const double f10 = d10 * s.F[0] + d11 * s.F[3], f11 = d10 * s.F[1] + d11 * s.F[4];
        s.F[0] = f00; s.F[1] = f01; s.F[3] = f10; s.F[4] = f11; s.F[8] = ezz * s.F[8];
        const double e00 = d00 * B[XX] + d01 * B[XY], e01 = d00 * B[XY] + d01 * B[YY];
        const double e10 = d10 * B[XX] + d11 * B[XY], e11 = d10 * B[XY] + d11 * B[YY];
        const double e22 = ezz * B[ZZ];
        B[XX] = e00 * d00 + e01 * d01;
        B[XY] = e00 * d10 + e01 * d11;
        B[YY] = e10 * d10 + e11 * d11;
        B[ZZ] = e22 * ezz;
        detDf = ezz * (d00 * d11 - d10 * d01);
    }
    double Jres = s.hist[1];
    const double dJres = s.dT != 0. ? exp(3. * m.p[12] * s.dT) : 1.;        // HyperElastic::GetIncrementalResJ (HyperElastic.cpp:89-95): exp(3 CTE1 dT)
    Jres *= dJres;
    s.hist[1] = Jres;
    if (DIM == 2 && np == NP_PLANE_STRESS) {
        // B.zz that zeroes the zz stress: Newton from the current (1 + ezz)^2 (:204-258)
        const double arg = B[XX] * B[YY] - B[XY] * B[XY];
        const double arg12 = sqrt(arg), arg16 = pow(arg, 1. / 6.), arg2 = B[XX] + B[YY];
        double xn = s.F[8] * s.F[8];
        for (int iter = 1; iter < 20; iter++) {
            const double xn16 = pow(xn, 1. / 6.), xn12 = sqrt(xn);
            const double J13 = xn16 * arg16, J0 = xn12 * arg12, Je = J0 / Jres;
            double mJ2P, mdJ2PdJ;
            if (UofJ == 1) { mJ2P = Ksp * Je * Je * (Je - 1.); mdJ2PdJ = Ksp * Je * (3. * Je - 2.); }
            else if (UofJ == 2) { mJ2P = Ksp * Je * log(Je); mdJ2PdJ = Ksp * (log(Je) + 1.); }
            else { mJ2P = 0.5 * Ksp * Je * (Je * Je - 1.); mdJ2PdJ = 0.5 * Ksp * (3. * Je * Je - 1.); }
            const double fx = 3. * Jres * mJ2P + G1sp * (2. * xn - arg2) * J13 + G2sp * (xn * arg2 - 2. * arg) / J13;
            const double fxp = (1.5 * J0 / xn) * mdJ2PdJ + G1sp * J13 * (14. * xn - arg2) / (6. * xn) + G2sp * (2. * arg + 5. * xn * arg2) / (6. * J13 * xn);
            const double xnp1 = xn - fx / fxp;
            if (fabs(xn - xnp1) < 1e-10) break;
            xn = xnp1;
        }
        const double dFzz = sqrt(xn / B[ZZ]);
        B[ZZ] = xn;
        s.F[8] = dFzz * s.F[8];
        detDf *= dFzz;
    }
    const double J = detDf * s.hist[0];
    s.hist[0] = J;
    double st0[6];
#pragma unroll
    for (int i = 0; i < 6; i++) st0[i] = s.sp[i];
    const double Jeff = J / Jres;
    const double p0 = s.pressure;
    double Kvol;
    if (UofJ == 1) Kvol = Ksp * (Jeff - 1.);
    else if (UofJ == 2) Kvol = Ksp * log(Jeff) / Jeff;
    else Kvol = 0.5 * Ksp * (Jeff - 1. / Jeff);
    const double Kterm = J * Kvol;
    const double delV = 1. - 1. / detDf;
    double QAVred = 0., AVEnergy = 0.;
    if (delV < 0. && m.p[3] != 0.) {
        QAVred = artificial_viscosity(delV / delTime, sqrt(Ksp * J), m);
        AVEnergy = fabs(QAVred * delV);
    }
    const double Pfinal = -Kterm + QAVred;
    s.pressure = Pfinal;
    const double avgP = 0.5 * (p0 + Pfinal);
    const double dilEnergy = -avgP * delV;
    const double delVres = 1. - 1. / dJres;
    const double resEnergy = -avgP * delVres;
    const double J23 = pow(J, 2. / 3.);
    const double J43 = J23 * J23;
    const double JforG1 = J23 / Jres, JforG2 = J43 / Jres;
    double *sp = s.sp;
    sp[XX] = (2 * B[XX] - B[YY] - B[ZZ]) * G1sp / (3. * JforG1) + (B[XX] * (B[YY] + B[ZZ]) - 2 * B[YY] * B[ZZ] - B[XY] * B[XY]) * G2sp / (3. * JforG2);
    sp[YY] = (2 * B[YY] - B[XX] - B[ZZ]) * G1sp / (3. * JforG1) + (B[YY] * (B[XX] + B[ZZ]) - 2 * B[XX] * B[ZZ] - B[XY] * B[XY]) * G2sp / (3. * JforG2);
    sp[ZZ] = (2 * B[ZZ] - B[XX] - B[YY]) * G1sp / (3. * JforG1) + (B[ZZ] * (B[XX] + B[YY]) - 2 * B[XX] * B[YY] + 2. * B[XY] * B[XY]) * G2sp / (3. * JforG2);
    sp[XY] = B[XY] * G1sp / JforG1 + (B[ZZ] * B[XY]) * G2sp / JforG2;
    if (DIM == 3) {
        sp[XX] += (2. * B[YZ] * B[YZ] - B[XZ] * B[XZ]) * G2sp / (3. * JforG2);
        sp[YY] += (2. * B[XZ] * B[XZ] - B[YZ] * B[YZ]) * G2sp / (3. * JforG2);
        sp[ZZ] -= (B[XZ] * B[XZ] + B[YZ] * B[YZ]) * G2sp / (3. * JforG2);
        sp[XY] -= B[XZ] * B[YZ] * G2sp / JforG2;
        sp[XZ] = B[XZ] * G1sp / JforG1 + (B[YY] * B[XZ] - B[XY] * B[YZ]) * G2sp / JforG2;
        sp[YZ] = B[YZ] * G1sp / JforG1 + (B[XX] * B[YZ] - B[XY] * B[XZ]) * G2sp / JforG2;
    }
    double shearEnergy = 0.5 * ((sp[XX] + st0[XX]) * du[0] + (sp[YY] + st0[YY]) * du[4] + (sp[ZZ] + st0[ZZ]) * du[8] + (sp[XY] + st0[XY]) * (du[1] + du[3]));
    if (DIM == 3) shearEnergy += 0.5 * ((sp[XZ] + st0[XZ]) * (du[2] + du[6]) + (sp[YZ] + st0[YZ]) * (du[5] + du[7]));
    s.work += dilEnergy + shearEnergy;
    s.res += resEnergy;
    double Kratio;
    if (UofJ == 1) Kratio = Jeff;
    else if (UofJ == 2) Kratio = (1 - log(Jeff)) / (Jeff * Jeff);
    else Kratio = 0.5 * (Jeff + 1. / Jeff);
    const double dTq0 = -J * Kratio * gamma0 * s.prevT * delV;
    increment_heat_energy(s, Cv, dTq0, AVEnergy);
}

// ---- IsoPlasticity + LinearHardening (MaterialID 9) ------------------------------------------------
// IsoPlasticity::MPMConstitutiveLaw / PlasticityConstLaw / UpdatePressure (Materials/IsoPlasticity.cpp:128-492),
// small-rotation branch, J2 potential (:513-517), closed-form radial return of LinearHardening
// (Materials/LinearHardening.cpp:93-145), history = cumulative plastic strain alpha.  3D, plane strain and plane stress
// (the latter with the numerical solve for lambda of HardeningLawBase::SolveForLambda).
// Reference quirks kept: 3D plastic-step work energy adds sp.zz*de.zz twice (:428-431); the 2D rotation of the
// prior shear stress uses the plastic strain (:252).
#define MPM_SQRT_TWOTHIRDS 0.8164965809277260

// ---- hardening laws other than Linear (HardeningLawBase and subclasses) ---------------------------------------------------
// Law ids as in MaterialBase::SetHardeningLaw (MaterialBaseMPM.cpp:548-595).  Parameter slots of an ISOPLASTICITY material:
// [16] law id (0 or 1 = Linear, closed-form return), then
//   Nonlinear  (2, Materials/NonlinearHardening.cpp):  yield = yldred (1 + beta alpha)^n        [17] beta [18] n
//   Nonlinear2 (6, Materials/Nonlinear2Hardening.cpp): yield = yldred (1 + beta alpha^n)        [17] beta [18] n
//   JohnsonCook (3, Materials/JohnsonCook.cpp): (yldred + Bred alpha^n)(1 + C ln(epdot/ep0) [+ D ln^n2])(1 - T*^m)
//        [17] Bred [18] n [19] C [20] ep0 [21] D [22] n2 [23] Tm [24] m [25] reference temperature [26] edotMin [27] eminTerm
//   SCGL (4, Materials/SCGLHardening.cpp): yield = min(yldred (1 + beta alpha)^n, yldMaxred) Gratio with the shear modulus ratio
//        Gratio = 1 + GPpred P + GTp (T - Tref) >= 0 of the particle's pressure and temperature, which also scales Gred
//        [17] beta [18] n [19] yldMaxred [20] GPpred [21] GTp [25] reference temperature
// alphaMax [14] and yldredMin [15] keep their meaning for the two power laws.
enum { HARD_LINEAR = 1, HARD_NONLINEAR = 2, HARD_JOHNSONCOOK = 3, HARD_SCGL = 4, HARD_NONLINEAR2 = 6 };
#define MPM_TWOTHIRDS 0.6666666666666667
#define MPM_SQRT_EIGHT27THS 0.5443310539518174

struct HardAlpha { double alpint, dalpha; };      // HardeningAlpha (Materials/HardeningLawBase.hpp)
struct HardProps { int law; double TjcTerm, hmlgTemp, Gratio; };       // JCProperties (JohnsonCook::GetCopyOfHardeningProps :130-152), SCGLProperties

__device__ __forceinline__ HardProps hard_props(const Material &m, double prevT, double pressure)
{
    HardProps h;
    h.law = (int)m.p[16]; h.TjcTerm = 1.; h.hmlgTemp = 0.; h.Gratio = 1.;
    if (h.law == HARD_SCGL) {        // SCGLHardening::GetShearRatio with J = 1 (SCGLHardening.cpp:141-152, IsoPlasticity.cpp:548)
        h.Gratio = 1. + m.p[20] * pressure + m.p[21] * (prevT - m.p[25]);
        if (h.Gratio < 0.) h.Gratio = 0.;
    }
    if (h.law == HARD_JOHNSONCOOK) {
        h.hmlgTemp = (prevT - m.p[25]) / (m.p[23] - m.p[25]);
        if (h.hmlgTemp > 1.) h.TjcTerm = 0.;
        else if (h.hmlgTemp > 0.) h.TjcTerm = 1. - pow(h.hmlgTemp, m.p[24]);
        else h.TjcTerm = 1.;
    }
    return h;
}

__device__ __forceinline__ bool dble_equal(double A, double B)      // Common/System/CommonUtilities.cpp:46-62
{
    const double diff = fabs(A - B);
    if (diff <= 1.0e-16) return true;
    A = fabs(A); B = fabs(B);
    const double largest = (B > A) ? B : A;
    return diff <= largest * 1.0e-7;
}

// the rate term of Johnson-Cook and its derivative factor (JohnsonCook.cpp:158-228)
__device__ __forceinline__ void jc_rate_terms(const Material &m, double delTime, const HardAlpha &a, double &term2, double &dterm2, bool &above)
{
    const double Cjc = m.p[19], ep0 = m.p[20], Djc = m.p[21], n2 = m.p[22], edotMin = m.p[26], eminTerm = m.p[27];
    const double ep = a.dalpha / (delTime * ep0);
    above = ep > edotMin;
    dterm2 = 0.;
    if (above) {
        term2 = 1. + Cjc * log(ep);
        dterm2 = Cjc * ep0 / a.dalpha;
        if (Djc != 0. && ep > 1.) {
            term2 += Djc * pow(log(ep), n2);
            dterm2 += Djc * ep0 * n2 * pow(log(ep), n2 - 1.) / a.dalpha;
        }
    } else term2 = eminTerm;
}

__device__ __forceinline__ double hard_yield(const Material &m, const HardProps &h, double delTime, const HardAlpha &a)
{
    const double yldred = m.p[10];
    if (h.law == HARD_SCGL) return fmin(yldred * pow(1. + m.p[17] * a.alpint, m.p[18]), m.p[19]) * h.Gratio;        // SCGLHardening.cpp:160-164
    if (h.law == HARD_NONLINEAR) return a.alpint < m.p[14] ? yldred * pow(1. + m.p[17] * a.alpint, m.p[18]) : m.p[15];
    if (h.law == HARD_NONLINEAR2) return a.alpint < m.p[14] ? yldred * (1. + m.p[17] * pow(a.alpint, m.p[18])) : m.p[15];
    if (h.hmlgTemp >= 1.) return 0.;
    const double term1 = yldred + m.p[17] * pow(a.alpint, m.p[18]);
    const double ep = a.dalpha / (delTime * m.p[20]);
    double term2 = ep > m.p[26] ? 1. + m.p[19] * log(ep) : m.p[27];
    if (m.p[21] != 0. && ep > 1.) term2 += m.p[21] * pow(log(ep), m.p[22]);
    return term1 * term2 * h.TjcTerm;
}

// d(sqrt(2/3) yield)/d lambda, 3D and plane strain
__device__ __forceinline__ double hard_kprime(const Material &m, const HardProps &h, double delTime, const HardAlpha &a)
{
    const double yldred = m.p[10];
    if (h.law == HARD_SCGL) {        // SCGLHardening.cpp:170-181
        if (yldred * pow(1. + m.p[17] * a.alpint, m.p[18]) >= m.p[19]) return 0.;
        const double bfactor = dble_equal(m.p[18], 1.) ? m.p[17] : m.p[17] * m.p[18] * pow(1. + m.p[17] * a.alpint, m.p[18] - 1.);
        return MPM_TWOTHIRDS * (yldred * h.Gratio) * bfactor;
    }
    if (h.law == HARD_NONLINEAR) return a.alpint < m.p[14] ? MPM_TWOTHIRDS * yldred * m.p[17] * m.p[18] * pow(1. + m.p[17] * a.alpint, m.p[18] - 1) : 0.;
    if (h.law == HARD_NONLINEAR2) return a.alpint < m.p[14] ? MPM_TWOTHIRDS * yldred * m.p[17] * m.p[18] * pow(a.alpint, m.p[18] - 1.) : 0.;
    if (h.hmlgTemp >= 1.) return 0.;
    const double dterm1 = m.p[17] * m.p[18] * pow(a.alpint, m.p[18] - 1.);
    double term2, dterm2; bool above;
    jc_rate_terms(m, delTime, a, term2, dterm2, above);
    if (above) {
        const double term1 = yldred + m.p[17] * pow(a.alpint, m.p[18]);
        return MPM_TWOTHIRDS * h.TjcTerm * (dterm1 * term2 + term1 * dterm2);
    }
    return MPM_TWOTHIRDS * h.TjcTerm * dterm1 * term2;
}

// d(yield^2 / 3)/d lambda, plane stress
__device__ __forceinline__ double hard_k2prime(const Material &m, const HardProps &h, double fnp1, double delTime, const HardAlpha &a)
{
    const double yldred = m.p[10];
    if (h.law == HARD_SCGL) {        // SCGLHardening.cpp:184-192
        if (yldred * pow(1. + m.p[17] * a.alpint, m.p[18]) >= m.p[19]) return 0.;
        const double factor = yldred * h.Gratio;
        return MPM_SQRT_EIGHT27THS * factor * factor * m.p[17] * m.p[18] * pow(1. + m.p[17] * a.alpint, 2. * m.p[18] - 1) * fnp1;
    }
    if (h.law == HARD_NONLINEAR)
        return a.alpint < m.p[14] ? MPM_SQRT_EIGHT27THS * yldred * yldred * m.p[17] * m.p[18] * pow(1. + m.p[17] * a.alpint, 2. * m.p[18] - 1) * fnp1 : 0.;
    if (h.law == HARD_NONLINEAR2) {
        if (dble_equal(a.alpint, 0.)) return 0.;
        if (a.alpint < m.p[14]) {
            const double alphan = pow(a.alpint, m.p[18]);
            return MPM_SQRT_EIGHT27THS * yldred * yldred * m.p[17] * m.p[18] * (1. + m.p[17] * alphan) * alphan * fnp1 / a.alpint;
        }
        return 0.;
    }
    if (dble_equal(a.alpint, 0.)) return 0.;
    if (h.hmlgTemp >= 1.) return 0.;
    const double term1 = yldred + m.p[17] * pow(a.alpint, m.p[18]);
    const double dterm1 = m.p[17] * m.p[18] * pow(a.alpint, m.p[18] - 1.);
    double term2, dterm2; bool above;
    jc_rate_terms(m, delTime, a, term2, dterm2, above);
    if (above) return MPM_SQRT_EIGHT27THS * term1 * term2 * fnp1 * h.TjcTerm * h.TjcTerm * (dterm1 * term2 + dterm2 * term1);
    return MPM_SQRT_EIGHT27THS * term1 * term2 * fnp1 * h.TjcTerm * h.TjcTerm * dterm1 * term2;
}

// K(alpha) - K(0) for the dissipated energy (HardeningLawBase.cpp:110-113; JohnsonCook.cpp:230-238)
__device__ __forceinline__ double hard_yield_increment(const Material &m, const HardProps &h, double delTime, const HardAlpha &a)
{
    if (h.law == HARD_SCGL) return (fmin(m.p[10] * pow(1. + m.p[17] * a.alpint, m.p[18]), m.p[19]) - m.p[10]) * h.Gratio;      // SCGLHardening.cpp:195-198
    if (h.law != HARD_JOHNSONCOOK) return hard_yield(m, h, delTime, a) - m.p[10];
    if (h.hmlgTemp >= 1.) return 0.;
    const double ep = a.dalpha / (delTime * m.p[20]);
    double term2 = ep > m.p[26] ? 1. + m.p[19] * log(ep) : m.p[27];
    if (m.p[21] != 0. && ep > 1.) term2 += m.p[21] * pow(log(ep), m.p[22]);
    return m.p[17] * pow(a.alpint, m.p[18]) * term2 * h.TjcTerm;
}

// HardeningLawBase::SolveForLambdaBracketed with BracketSolution (HardeningLawBase.cpp:211-381): Newton's method kept inside a
// bracket, bisecting when a step would leave it.  a holds alpha at the start of the step (dalpha = 0) and the solution at the end.
// ok = false: the plane-stress bracket was not found in 20 decades of strain rate (the reference throws); lambda = NaN then
// poisons the particle so that the run stops loudly (MPMGPU_ENAN) instead of continuing on a wrong state.
template <bool PLANE_STRESS>
__device__ __forceinline__ double solve_lambda_bracketed(const Material &m, const HardProps &h, double alpha0, double strial, const double stk[6],
                                                         double Gred, double psKred, double Pfinal, double delTime, HardAlpha &a, bool &ok)
{
    ok = true;
    if (h.law == HARD_JOHNSONCOOK && h.hmlgTemp >= 1.) return strial / (2. * Gred);          // melted (JohnsonCook.cpp:244-249)
    double xl = 0., xh = 0.;        // xl: g < 0 (higher lambda), xh: g > 0 (lower lambda)
    double n1trial = 0., n2trial = 0.;
    if (PLANE_STRESS) {
        n2trial = -stk[XX] + stk[YY];
        n2trial *= 0.5 * n2trial;
        n2trial += 2. * stk[XY] * stk[XY];
        n1trial = stk[XX] + stk[YY] - 2. * Pfinal;
        n1trial *= n1trial / 6.;
        double epdot = 1.;
        bool found = false;
        for (int step = 0; step < 20; step++) {
            a.dalpha = epdot * delTime;
            a.alpint = alpha0 + a.dalpha;
            const double lambdak = a.dalpha / MPM_SQRT_TWOTHIRDS;
            const double d1 = (1 + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
            const double fnp12 = n1trial / (d1 * d1) + n2trial / (d2 * d2);
            const double kyld = hard_yield(m, h, delTime, a);
            const double gmax = 0.5 * fnp12 - kyld * kyld / 3.;
            if (gmax < 0.) { xl = a.dalpha / MPM_SQRT_TWOTHIRDS; found = true; break; }
            xh = lambdak;
            epdot *= 10.;
        }
        if (!found) { ok = false; return (double)NAN; }        // the reference throws here; NaN stops the run with MPMGPU_ENAN at the next element reset
    } else {
        const double dalpha = strial / (2. * Gred);
        a.alpint = alpha0 + dalpha;
        if (hard_yield(m, h, delTime, a) <= 0.) xh = dalpha / MPM_SQRT_TWOTHIRDS;
        xl = dalpha / MPM_SQRT_TWOTHIRDS;
    }
    if (xh > xl) return xh;
    double lambdak = 0.5 * (xl + xh);
    a.dalpha = PLANE_STRESS ? 0. : MPM_SQRT_TWOTHIRDS * lambdak;        // UpdateTrialAlpha(lambdak, fnp1 = 0): zero in plane stress
    a.alpint = alpha0 + a.dalpha;
    double dxold = fabs(xh - xl), dx = dxold;
    for (int step = 1;;) {
        double glam, slope, fnp1 = 0.;
        if (PLANE_STRESS) {
            const double d1 = (1 + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
            const double fnp12 = n1trial / (d1 * d1) + n2trial / (d2 * d2);
            const double kyld = hard_yield(m, h, delTime, a);
            glam = 0.5 * fnp12 - kyld * kyld / 3.;
            fnp1 = sqrt(fnp12);
            slope = -(psKred * n1trial / (d1 * d1 * d1) + 2 * Gred * n2trial / (d2 * d2 * d2)) - hard_k2prime(m, h, fnp1, delTime, a);
        } else {
            glam = strial - 2 * Gred * lambdak - MPM_SQRT_TWOTHIRDS * hard_yield(m, h, delTime, a);
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
            const double temp = lambdak;
            lambdak -= dx;
            if (temp == lambdak) break;
        }
        a.dalpha = PLANE_STRESS ? MPM_SQRT_TWOTHIRDS * lambdak * fnp1 : MPM_SQRT_TWOTHIRDS * lambdak;
        a.alpint = alpha0 + a.dalpha;
        if (step++ > 20 || fabs(dx / lambdak) < 0.0001) break;
        if (glam < 0.) xl = lambdak; else xh = lambdak;
    }
    return lambdak;
}
template <int DIM, bool LR = false, bool GENERAL = false>
__device__ __forceinline__ void isoplasticity_law(PState &s, const double du[9], double delTime, int np, const Material &m)
{
    // GENERAL: a hardening law other than Linear (slot 16): numerical return by the bracketed Newton's method of HardeningLawBase
    HardProps hp;
    if (GENERAL) hp = hard_props(m, s.prevT, s.pressure);
    // (the shear modulus follows the hardening law's ratio: IsoPlasticity::GetCopyOfMechanicalProps :548-549)
    const double Gred = GENERAL ? m.p[8] * hp.Gratio : m.p[8], Kred = m.p[9], yldred = m.p[10], Epred = m.p[11];
    const double gamma0 = m.p[13], Cv = m.p[1], alphaMax = m.p[14], yldredMin = m.p[15];
    // large rotation (:140-158): the strain increment in the current configuration replaces du, state n-1 is rotated by dR
    double deLR[9], dR[9];
    if (LR) lr_strain_increment<DIM>(s, du, deLR, dR);
    else hypo_increment_deformation<DIM>(s, du);
    const double *de = LR ? deLR : du;
    // plane stress terms (IsoPlasticity::GetCopyOfMechanicalProps :551-556)
    const bool planeStress = DIM == 2 && np == NP_PLANE_STRESS;
    const double psRed = 1. / (Kred / (2. * Gred) + 2. / 3.), psLr2G = (Kred / (2. * Gred) - 1. / 3.) * psRed, psKred = Kred * psRed;
    const double eres = m.p[12] * s.dT;         // CTE3 dT (:133)
    const double delV = planeStress ? psRed * (de[0] + de[4] - 2. * eres) : de[0] + de[4] + de[8] - 3. * eres;          // :147-156, :165-174
    const double P0 = s.pressure;
    const double dgxy = de[1] + de[3];
    double dgxz = 0., dgyz = 0.;
    if (DIM == 3) { dgxz = de[2] + de[6]; dgyz = de[5] + de[7]; }
    const double dexxr = de[0] - eres, deyyr = de[4] - eres, dezzr = de[8] - eres;
    // UpdatePressure (:462-492)
    double dP = -Kred * delV;
    const double dVoverV = delV + 3. * eres;
    double dispEnergy = 0.;
    if (dVoverV < 0. && m.p[3] != 0.) {         // artificial viscosity (:474-479)
        const double QAVred = artificial_viscosity(dVoverV / delTime, sqrt(Kred), m);
        dispEnergy += fabs(QAVred * dVoverV);
        dP += QAVred;
    }
    s.pressure += dP;
    double Pfinal = s.pressure;
    s.work += -Pfinal * dVoverV;
    s.res += -3. * Pfinal * eres;
    double dTq0 = -gamma0 * s.prevT * dVoverV;
    // rotate plastic strain and prior stress (:218-268)
    double *ep = s.eplast, *sp = s.sp;
    double e0[6], st0[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { e0[i] = ep[i]; st0[i] = sp[i]; }
    const double dwrotxy = de[3] - de[1];
    if (LR) {                   // :214-222
        rotate_voight<DIM>(dR, e0, false, ep);
        rotate_voight<DIM>(dR, sp, true, st0);
    } else if (DIM == 2) {
        const double dnorm = 0.5 * dwrotxy * e0[XY];
        ep[XX] -= dnorm; ep[YY] += dnorm; ep[XY] += dwrotxy * (e0[XX] - e0[YY]);
        const double dn = dwrotxy * sp[XY];
        st0[XX] -= dn; st0[YY] += dn; st0[XY] += 0.5 * dwrotxy * (e0[XX] - e0[YY]);
    } else {
        const double dwrotxz = de[6] - de[2], dwrotyz = de[7] - de[5];
        const double dxy = 0.5 * dwrotxy * e0[XY], dxz = 0.5 * dwrotxz * e0[XZ], dyz = 0.5 * dwrotyz * e0[YZ];
        ep[XX] += -dxy - dxz; ep[YY] += dxy - dyz; ep[ZZ] += dxz + dyz;
        ep[YZ] += dwrotyz * (e0[YY] - e0[ZZ]) + 0.5 * (dwrotxz * e0[XY] + dwrotxy * e0[XZ]);
        ep[XZ] += dwrotxz * (e0[XX] - e0[ZZ]) + 0.5 * (dwrotyz * e0[XY] - dwrotxy * e0[YZ]);
        ep[XY] += dwrotxy * (e0[XX] - e0[YY]) - 0.5 * (dwrotyz * e0[XZ] + dwrotxz * e0[YZ]);
        const double sxy = dwrotxy * sp[XY], sxz = dwrotxz * sp[XZ], syz = dwrotyz * sp[YZ];
        st0[XX] += -sxy - sxz; st0[YY] += sxy - syz; st0[ZZ] += sxz + syz;
        st0[YZ] += 0.5 * (dwrotyz * (sp[YY] - sp[ZZ]) + dwrotxz * sp[XY] + dwrotxy * sp[XZ]);
        st0[XZ] += 0.5 * (dwrotxz * (sp[XX] - sp[ZZ]) + dwrotyz * sp[XY] - dwrotxy * sp[YZ]);
        st0[XY] += 0.5 * (dwrotxy * (sp[XX] - sp[YY]) - dwrotyz * sp[XZ] - dwrotxz * sp[YZ]);
    }
    // trial deviatoric stress (:271-295)
    const double thirdDelV = delV / 3.;
    double strial[6];
    strial[XX] = st0[XX] + 2. * Gred * (dexxr - thirdDelV);
    strial[YY] = st0[YY] + 2. * Gred * (deyyr - thirdDelV);
    strial[ZZ] = st0[ZZ] + (planeStress ? Pfinal - P0 : 2. * Gred * (dezzr - thirdDelV));       // :275-278
    strial[XY] = st0[XY] + Gred * dgxy;
    strial[YZ] = DIM == 3 ? st0[YZ] + Gred * dgyz : st0[YZ];
    strial[XZ] = DIM == 3 ? st0[XZ] + Gred * dgxz : st0[XZ];
    // plastic potential (:513-517, MoreIsotropicMat.cpp:354-372)
    const double alpha0 = s.hist[0];
    double ss = strial[XX] * strial[XX] + strial[YY] * strial[YY] + strial[ZZ] * strial[ZZ];
    double tt = strial[XY] * strial[XY];
    if (DIM == 3) tt += strial[XZ] * strial[XZ] + strial[YZ] * strial[YZ];
    const double smag = sqrt(ss + tt + tt);
    HardAlpha ha;
    ha.alpint = alpha0; ha.dalpha = 0.;                 // UpdateTrialAlpha(mptr, np, &alpha, 0)
    const double yield0 = GENERAL ? hard_yield(m, hp, delTime, ha) : (alpha0 < alphaMax ? yldred + Epred * alpha0 : yldredMin);
    const double ftrial = smag - MPM_SQRT_TWOTHIRDS * yield0;
    if (ftrial < 0.) {
#pragma unroll
        for (int i = 0; i < 6; i++) sp[i] = strial[i];
        if (DIM == 3) s.work += sp[XX] * de[0] + sp[YY] * de[4] + sp[ZZ] * de[8] + sp[YZ] * dgyz + sp[XZ] * dgxz + sp[XY] * dgxy;
        else {
            if (planeStress) {          // zz deformation (:315-320), MPMBase::IncrementDeformationGradientZZ
                const double dezz = -psLr2G * (dexxr + deyyr) + eres;
                s.F[8] += dezz * s.F[8];
            }
            s.work += sp[XX] * de[0] + sp[YY] * de[4] + sp[XY] * dgxy;
        }
        increment_heat_energy(s, Cv, dTq0, dispEnergy);
        return;
    }
    double lambdak, alpint, dfds[6], dezzTotal = de[8];
    double spPS[4] = {0., 0., 0., 0.};
    if (planeStress) {
        // plane stress: the return direction changes with lambda; unbracketed Newton of HardeningLawBase::SolveForLambda
        // (HardeningLawBase.cpp:157-202, chosen by LinearHardening.cpp:128-131)
        lambdak = 0.;
        alpint = alpha0;
        if (GENERAL) {
            bool ok;
            lambdak = solve_lambda_bracketed<true>(m, hp, alpha0, smag, strial, Gred, psKred, Pfinal, delTime, ha, ok);
            alpint = ha.alpint;
        } else {
        double n2trial = -strial[XX] + strial[YY];
        n2trial *= n2trial / 2;
        n2trial += 2. * strial[XY] * strial[XY];
        double n1trial = strial[XX] + strial[YY] - 2. * Pfinal;
        n1trial *= n1trial / 6.;
        for (int step = 1;;) {
            const double d1 = (1 + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
            const double fnp12 = n1trial / (d1 * d1) + n2trial / (d2 * d2);
            const double kyld = alpint < alphaMax ? yldred + Epred * alpint : yldredMin;
            const double glam = 0.5 * fnp12 - kyld * kyld / 3.;
            const double fnp1 = sqrt(fnp12);
            const double k2prime = alpint < alphaMax ? 0.5443310539518174 * (yldred + Epred * alpint) * Epred * fnp1 : 0.;    // GetK2Prime
            const double slope = -(psKred * n1trial / (d1 * d1 * d1) + 2 * Gred * n2trial / (d2 * d2 * d2)) - k2prime;
            const double delLam = -glam / slope;
            lambdak += delLam;
            alpint = alpha0 + MPM_SQRT_TWOTHIRDS * lambdak * fnp1;          // UpdateTrialAlpha, plane stress
            if (step++ > 20 || fabs(delLam / lambdak) < 0.0001) break;      // LambdaConverged
        }
        }
        // final stress and direction (:345-389)
        const double d1 = (1. + psKred * lambdak), d2 = (1. + 2. * Gred * lambdak);
        const double n1 = (strial[XX] + strial[YY] - 2. * Pfinal) / d1, n2 = (-strial[XX] + strial[YY]) / d2;
        const double sxx = (n1 - n2) / 2., syy = (n1 + n2) / 2., txy = strial[XY] / d2;
        dfds[XX] = (2. * sxx - syy) / 3.; dfds[YY] = (2. * syy - sxx) / 3.; dfds[ZZ] = -(dfds[XX] + dfds[YY]); dfds[XY] = txy;
        dfds[XZ] = 0.; dfds[YZ] = 0.;
        const double dPps = -n1 / 3. - Pfinal;
        s.pressure += dPps;
        const double dezzp = lambdak * dfds[ZZ];
        const double dVtot = delV + 3. * eres + psRed * dezzp;
        s.work += -Pfinal * psRed * dezzp - dPps * dVtot;
        s.res += -3. * dPps * eres;
        Pfinal = s.pressure;
        dezzTotal = -psLr2G * (dexxr + deyyr - lambdak * (dfds[XX] + dfds[YY])) + dezzp + eres;
        s.F[8] += dezzTotal * s.F[8];
        dTq0 -= gamma0 * s.prevT * dezzp;
        spPS[0] = sxx + Pfinal; spPS[1] = syy + Pfinal; spPS[2] = Pfinal; spPS[3] = txy;
    } else {
        if (GENERAL) {
            bool ok;
            lambdak = solve_lambda_bracketed<false>(m, hp, alpha0, smag, strial, Gred, 0., Pfinal, delTime, ha, ok);
            alpint = ha.alpint;
        } else {
            // radial return, closed form (LinearHardening.cpp:124-145)
            lambdak = (smag - MPM_SQRT_TWOTHIRDS * (yldred + Epred * alpha0)) / (2. * (Gred + Epred / 3.));
            if (alpha0 + MPM_SQRT_TWOTHIRDS * lambdak > alphaMax) lambdak = (smag - MPM_SQRT_TWOTHIRDS * yldredMin) / (2. * Gred);
            alpint = alpha0 + MPM_SQRT_TWOTHIRDS * lambdak;
        }
#pragma unroll
        for (int i = 0; i < 6; i++) dfds[i] = strial[i] / smag;           // df/dsigma = s/|s| (:496-509)
    }
    // plastic strain increment (:395-405)
    double dep[6];
    dep[XX] = lambdak * dfds[XX]; dep[YY] = lambdak * dfds[YY]; dep[ZZ] = lambdak * dfds[ZZ];
    dep[XY] = 2. * lambdak * dfds[XY];
    dep[XZ] = DIM == 3 ? 2. * lambdak * dfds[XZ] : 0.;
    dep[YZ] = DIM == 3 ? 2. * lambdak * dfds[YZ] : 0.;
    ep[XX] += dep[XX]; ep[YY] += dep[YY]; ep[ZZ] += dep[ZZ]; ep[XY] += dep[XY];
    if (DIM == 3) { ep[XZ] += dep[XZ]; ep[YZ] += dep[YZ]; }
    sp[XX] = strial[XX] - 2. * Gred * dep[XX];
    sp[YY] = strial[YY] - 2. * Gred * dep[YY];
    sp[ZZ] = strial[ZZ] - 2. * Gred * dep[ZZ];
    sp[XY] = strial[XY] - Gred * dep[XY];
    if (DIM == 3) { sp[YZ] = strial[YZ] - Gred * dep[YZ]; sp[XZ] = strial[XZ] - Gred * dep[XZ]; }
    if (planeStress) { sp[XX] = spPS[0]; sp[YY] = spPS[1]; sp[ZZ] = spPS[2]; sp[XY] = spPS[3]; }      // set by the plane-stress return (:386-389)
    double workEnergy = sp[XX] * de[0] + sp[YY] * de[4] + sp[XY] * dgxy;
    if (DIM == 3) workEnergy += sp[ZZ] * de[8] + sp[YZ] * dgyz + sp[XZ] * dgxz;
    if (np != NP_PLANE_STRAIN) workEnergy += sp[ZZ] * dezzTotal;
    s.work += workEnergy;
    double plastEnergy = sp[XX] * dep[XX] + sp[YY] * dep[YY] + sp[ZZ] * dep[ZZ] + sp[XY] * dep[XY];
    if (DIM == 3) plastEnergy += sp[XZ] * dep[XZ] + sp[YZ] * dep[YZ];
    const double yieldInc = GENERAL ? hard_yield_increment(m, hp, delTime, ha)
                                    : fmax(Epred * alpint, yldredMin - yldred);      // LinearHardening::GetYieldIncrement
    dispEnergy += plastEnergy - lambdak * MPM_SQRT_TWOTHIRDS * yieldInc;
    s.plast += dispEnergy;
    increment_heat_energy(s, Cv, dTq0, dispEnergy);
    s.hist[0] = alpint;
}

// Dispatch on the particle's material kind (MaterialBase::MPMConstitutiveLaw virtual call)
// ELASTIC_ONLY: the caller knows every material in use is IsotropicMat (keeps the other laws out of the kernel)
template <int DIM, bool ELASTIC_ONLY = false>
__device__ __forceinline__ void constitutive_law(PState &s, const double du[9], double delTime, int np, const Material &m)
{
    if (ELASTIC_ONLY) { isotropic_law<DIM>(s, du, np, m); return; }
    switch (m.kind) {
    case MAT_ISOTROPIC:
        isotropic_law<DIM>(s, du, np, m);
        break;
    case MAT_NEOHOOKEAN:
        neohookean_law<DIM>(s, du, delTime, np, m);
        break;
    case MAT_ISOPLASTICITY:
        isoplasticity_law<DIM>(s, du, delTime, np, m);
        break;
    default:
        break;
    }
}

// The same dispatch extended by the large-rotation variants of IsotropicMat and IsoPlasticity (material slot p[7],
// Elastic::useLargeRotation) and by Mooney; kept apart so that the kernels of the other configurations do not carry them
template <int DIM>
__device__ __forceinline__ void constitutive_law_lr(PState &s, const double du[9], double delTime, int np, const Material &m)
{
    const bool lr = m.p[7] != 0.;
    switch (m.kind) {
    case MAT_ISOTROPIC:
        if (lr) isotropic_lr_law<DIM>(s, du, np, m); else isotropic_law<DIM>(s, du, np, m);
        break;
    case MAT_NEOHOOKEAN:
        neohookean_law<DIM>(s, du, delTime, np, m);
        break;
    case MAT_ISOPLASTICITY:
        if (m.p[16] > 1.) { if (lr) isoplasticity_law<DIM, true, true>(s, du, delTime, np, m); else isoplasticity_law<DIM, false, true>(s, du, delTime, np, m); }
        else if (lr) isoplasticity_law<DIM, true>(s, du, delTime, np, m); else isoplasticity_law<DIM, false>(s, du, delTime, np, m);
        break;
    case MAT_MOONEY:
        mooney_law<DIM>(s, du, delTime, np, m);
        break;
    default:
        break;
    }
}
