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

// MaterialBase::IncrementHeatEnergy, isothermal branch (Materials/MaterialBaseMPM.cpp:982-1006;
// ConductionTask::adiabatic is false unless <EnergyCoupling/> -- transport is out of scope)
__device__ __forceinline__ void increment_heat_energy(PState &s, double Cv, double dTq0, double dPhi)
{
    double baseHeat = -Cv * dTq0;
    s.heat += baseHeat - dPhi;
    s.entropy += baseHeat / s.prevT;
}

// Elastic::HypoIncrementDeformation (Materials/ElasticMPM.cpp:392-403): F <- (I + du) F
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
        double st0[6];
#pragma unroll
        for (int i = 0; i < 6; i++) st0[i] = s.sp[i];
        double delsp[6];
        delsp[XX] = q[8] * dvxx + q[9] * dvyy + q[10] * dvzz;     // C11 C12 C13
        delsp[YY] = q[9] * dvxx + q[11] * dvyy + q[12] * dvzz;    // C12 C22 C23
        delsp[ZZ] = q[10] * dvxx + q[12] * dvyy + q[13] * dvzz;   // C13 C23 C33
        delsp[YZ] = q[14] * dgamyz;                               // C44
        delsp[XZ] = q[15] * dgamxz;                               // C55
        delsp[XY] = q[16] * dgamxy;                               // C66
        hypo3d(s.sp, dwrotxy, dwrotxz, dwrotyz, delsp);
        s.work += 0.5 * ((st0[XX] + s.sp[XX]) * dvxx + (st0[YY] + s.sp[YY]) * dvyy + (st0[ZZ] + s.sp[ZZ]) * dvzz +
                         (st0[YZ] + s.sp[YZ]) * dgamyz + (st0[XZ] + s.sp[XZ]) * dgamxz + (st0[XY] + s.sp[XY]) * dgamxy);
        // residual energy increment is 0.5*(trace sum)*eres with eres = 0
        double dTq0 = -gamma0 * s.prevT * dVoverV;
        increment_heat_energy(s, Cv, dTq0, 0.);
    } else {
        const double dvxx = du[0], dvyy = du[4];
        const double dgam = du[1] + du[3];
        const double dwrotxy = du[3] - du[1];
        double dVoverV = dvxx + dvyy;
        double st0[6];
#pragma unroll
        for (int i = 0; i < 6; i++) st0[i] = s.sp[i];
        const double c1 = q[8] * dvxx + q[9] * dvyy;      // C[1][1] C[1][2]
        const double c2 = q[9] * dvxx + q[11] * dvyy;     // C[1][2] C[2][2]
        const double c3 = q[16] * dgam;                   // C[3][3]
        hypo2d(s.sp, dwrotxy, c1, c2, c3);
        double workEnergy = 0.5 * ((st0[XX] + s.sp[XX]) * dvxx + (st0[YY] + s.sp[YY]) * dvyy + (st0[XY] + s.sp[XY]) * dgam);
        if (np == NP_PLANE_STRAIN) {
            s.sp[ZZ] += q[21] * dvxx + q[22] * dvyy;      // C[4][1] C[4][2]; (doopse - ezzres) = 0
        } else {
            // plane stress: out-of-plane strain increment (MoreIsotropicMat.cpp:249-258)
            double dezz = q[21] * dvxx + q[22] * dvyy;
            s.F[8] += dezz * s.F[8];                      // MPMBase::IncrementDeformationGradientZZ: ep.zz += dezz*(1+ep.zz) (MPMBase.cpp:637-639)
            workEnergy += 0.5 * (st0[ZZ] + s.sp[ZZ]) * dezz;
            dVoverV += dezz;
        }
        s.work += workEnergy;
        double dTq0 = -gamma0 * s.prevT * dVoverV;
        increment_heat_energy(s, Cv, dTq0, 0.);
    }
}

// Dispatch on the particle's material kind (MaterialBase::MPMConstitutiveLaw virtual call)
template <int DIM>
__device__ __forceinline__ void constitutive_law(PState &s, const double du[9], double delTime, int np, const Material &m)
{
    switch (m.kind) {
    case MAT_ISOTROPIC:
        isotropic_law<DIM>(s, du, np, m);
        break;
    default:
        break;
    }
}
