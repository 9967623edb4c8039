/*
 * mpmgpu.h -- C ABI of libmpmgpu: NairnMPM's explicit MPM time step on one B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's MPMTask pipeline
 * (NairnMPM/src/NairnMPM_Class/NairnMPM.cpp:870-1110 CreateTasks, :284-335 MPMStep).  The host
 * driver (the reference's own C++ driver through nairn_mpm_fea_b200/host/GpuTasks.cpp, or the Python
 * host in nairn_mpm_fea_b200/) owns input parsing, materials objects, BC lists and archiving; this
 * library owns device-resident structure-of-arrays copies of the particle state (MPMBase,
 * MPM_Classes/MPMBase.hpp:34-264) and node state (MatVelocityField, Nodes/MatVelocityField.hpp:44-48)
 * and runs tasks 1-9 and 11 of the step on them.
 *
 * Conventions
 *  - every entry point returns 0 on success, a negative MPMGPU_E* code on failure; the text is
 *    available from mpmgpu_last_error().  No C++ exception crosses this boundary: the C++ task
 *    wrapper turns a failure into the reference's CommonException (Common/Exceptions/CommonException.hpp).
 *  - plain pointers and sizes only; all floating point is IEEE double; all host arrays are
 *    structure-of-arrays, component-major ( v[c*n + p] ), in the caller's particle order
 *    (the reference's mpm[] order).  The device may keep particles in cell-sorted order internally;
 *    uploads/downloads always use the caller's order.
 *  - units are whatever the host uses consistently (the reference's internal mm-g-s units).
 *  - node numbers are the reference's 1-based numbers (Read_MPM/Generators.cpp:1910); element numbers
 *    are the reference's 1-based MPMBase::inElem (MPM_Classes/MPMBase.cpp:456-472).
 *  - one context is bound to one CUDA device and is used from one host thread at a time.
 *  - there is NO CPU fallback: without a CUDA device mpmgpu_create fails with MPMGPU_ENODEVICE.
 */
#ifndef MPMGPU_H
#define MPMGPU_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MPMGPU_ABI_VERSION 1

/* error codes */
#define MPMGPU_OK            0
#define MPMGPU_EINVAL       -1   /* bad argument / unsupported option (message says which) */
#define MPMGPU_ENODEVICE    -2   /* no usable CUDA device */
#define MPMGPU_ECUDA        -3   /* CUDA runtime error */
#define MPMGPU_ELEFTGRID    -4   /* a particle left the grid (ResetElementsTask.cpp:71-95) and could not be returned */
#define MPMGPU_ENAN         -5   /* particle position became NaN (ResetElementsTask.cpp:200-203) */
#define MPMGPU_ESTATE       -6   /* call out of order (e.g. step before upload) */
#define MPMGPU_ECPDI        -7   /* CPDI corner left the grid (MPM_Classes/MatPoint3D.cpp:596-601) */

/* analysis type: reference codes, System/MPMPrefix.hpp:114-115 */
#define MPMGPU_PLANE_STRAIN_MPM 10
#define MPMGPU_PLANE_STRESS_MPM 11
#define MPMGPU_THREED_MPM       12

/* update method: reference codes, System/MPMPrefix.hpp:110 */
#define MPMGPU_USF   0
#define MPMGPU_USAVG 2
#define MPMGPU_USL   3

/* shape functions: reference ElementBase::useGimp codes, System/MPMPrefix.hpp:127-140 */
#define MPMGPU_POINT_GIMP     0   /* "Classic": linear element shape functions */
#define MPMGPU_UNIFORM_GIMP   1   /* uGIMP */
#define MPMGPU_BSPLINE_GIMP    5   /* B2GIMP: quadratic B-spline GIMP (per-task kernels) */
#define MPMGPU_BSPLINE         6   /* B2SPLINE: quadratic B-splines (per-task kernels) */
#define MPMGPU_LINEAR_CPDI   10   /* lCPDI (2D and 3D) */
#define MPMGPU_QUADRATIC_CPDI 11  /* qCPDI (2D only, as in the reference) */
#define MPMGPU_BSPLINE_CPDI   13  /* B2CPDI: linear CPDI domains over quadratic B-splines (per-task kernels) */

/* material kinds: reference MaterialID() values, Common/Read_XML/MaterialController.cpp:105-232 */
#define MPMGPU_MAT_NONE           0   /* place holder for a host materials-list entry that is not a particle material (a contact law) */
#define MPMGPU_MAT_ISOTROPIC      1   /* IsotropicMat, small- or large-rotation hypoelastic */
#define MPMGPU_MAT_MOONEY         8   /* Mooney (Mooney-Rivlin hyperelastic; per-task kernels) */
#define MPMGPU_MAT_ISOPLASTICITY  9   /* IsoPlasticity + LinearHardening */
#define MPMGPU_MAT_RIGIDBC       11   /* RigidMaterial used as moving velocity BC */
#define MPMGPU_MAT_NEOHOOKEAN    28   /* Neohookean */
#define MPMGPU_MAT_RIGIDCONTACT  35   /* RigidMaterial in contact mode (SetDirection 8): multimaterial mode only.  Its particles follow the
                                         nonrigid ones and precede the rigid-BC particles in the host's order (NairnMPM.cpp:1121-1190), move at
                                         their own velocity (mpmgpu_update_rigid_velocities covers them) and extrapolate to their material's
                                         velocity field, against which every nonrigid material of a node makes contact
                                         (CrackVelocityFieldMulti::RigidMaterialContactOnCVF :676-955).  Slot [0] rho (1). */

#define MPMGPU_MAT_NPARAMS 32
#define MPMGPU_MAX_HISTORY 4

typedef struct mpmgpu_ctx mpmgpu_ctx;

/* Grid, method and global constants: everything CreateTasks/MPMStep read from fmobj, mpmgrid,
 * bodyFrc and ElementBase statics. */
typedef struct mpmgpu_config {
    int abi_version;        /* MPMGPU_ABI_VERSION */
    int device;             /* CUDA device ordinal */
    int np;                 /* analysis type (MPMGPU_*_MPM) */
    int horiz, vert, depth; /* cells per axis INCLUDING the automatic border cell on each side
                               (Read_MPM/Generators.cpp:1769-1785); depth = 0 in 2D */
    const double *xpts;     /* node coordinates per axis exactly as the host generated them, */
    const double *ypts;     /*   horiz+1 / vert+1 / depth+1 long (Generators.cpp:1823-1833);  */
    const double *zpts;     /*   zpts NULL in 2D.  Element extents are taken from these.      */
    double gridx, gridy, gridz; /* mpmgrid.grid (MeshInfo::SetCartesian); gridz = 0 in 2D */
    double thickness;       /* 2D grid thickness (unused by the step, kept for archives) */
    int shape;              /* MPMGPU_POINT_GIMP | _UNIFORM_GIMP | _BSPLINE_GIMP | _BSPLINE | _LINEAR_CPDI | _QUADRATIC_CPDI */
    double cpdi_rcrit;      /* ElementBase::rcrit, <0 for none (MatPoint3D.cpp:436-442) */
    int method;             /* MPMGPU_USF | _USAVG | _USL */
    int skip_post_extrapolation; /* <SkipPostExtrapolation/>: USL-/USAVG- (NairnMPM.cpp:1076-1087) */
    double fraction_usf;    /* fractionUSF (NairnMPM.cpp:78), 0.5 default */
    int xpic_order;         /* bodyFrc.GetXPICOrder(): 0 FLIP, 1 PIC, k>1 XPIC(k)/FMPM(k) */
    int using_fmpm;         /* bodyFrc.UsingFMPM() */
    double grid_damping;    /* bodyFrc.GetGridDamping(mtime) (constant part) */
    double particle_damping;/* bodyFrc.GetParticleDamping(mtime) */
    double gravity[3];      /* bodyFrc.gforce (zero vector if no gravity) */
    int max_particles;      /* capacity; 0 = size to the first upload */
    int sort_interval;      /* re-sort particles by cell every this many steps (0 = library default) */
    int kernel_path;        /* 0 = auto (tiled fast path when eligible), 1 = force reference-order
                               per-task kernels, 2 = force tiled path (error if not eligible) */
} mpmgpu_config;

/* One material: kind + POD parameter block harvested from the host's MaterialBase object
 * (what GetCopyOfMechanicalProps would hand to MPMConstitutiveLaw, MaterialBase.hpp:143-165).
 * Parameter slots by kind (all stiffness-like values are SPECIFIC, i.e. divided by rho):
 *  all kinds:  [0] rho   [1] heat capacity Cv   [2] particle damping override or -1
 *              [3] artificial viscosity on (1) / off (0)  [4] avA1  [5] avA2   (MaterialBaseMPM.cpp:202-216,1824-1829;
 *              Neohookean and IsoPlasticity only)   [6] reserved: the library stores the average cell size here
 *              [7] Elastic::useLargeRotation (<largeRotation>1</largeRotation>, Common/Materials/Elastic.cpp:27-33): ISOTROPIC
 *              and ISOPLASTICITY take the small-strain / large-rotation update (IsotropicMat::LRConstitutiveLaw,
 *              Materials/MoreIsotropicMat.cpp:54-174; IsoPlasticity.cpp:140-158,214-222) on the per-task kernels
 *  ISOTROPIC (3D):  [8] C11 [9] C12 [10] C13 [11] C22 [12] C23 [13] C33 [14] C44 [15] C55 [16] C66
 *                   [17] CTE1 [18] CTE2 [19] CTE3 [20] gamma0
 *  ISOTROPIC (2D, ElasticProperties 2D slots, Common/Materials/Elastic.cpp:90-140):
 *                   [8] C[1][1] [9] C[1][2] [11] C[2][2] [16] C[3][3] [21] C[4][1] [22] C[4][2]
 *                   [23] C[4][4] [24] C[5][1]  (+CTE/gamma0 as above)
 *  NEOHOOKEAN:      [8] Gsp [9] Ksp [10] Lamesp [11] UofJOption [12] CTE1 [13] gamma0
 *  MOONEY:          [8] G1sp [9] G2sp [10] Ksp [11] UofJOption [12] CTE1 [13] gamma0   (Materials/Mooney.cpp:104-147; history J, Jres)
 *  ISOPLASTICITY:   [8] Gred [9] Kred [10] yldred [11] Epred [12] CTE3 [13] gamma0
 *                   [14] alphaMax [15] yldredMin  (LinearHardening.cpp:55-80)
 *                   [16] hardening law, ids of MaterialBase::SetHardeningLaw (MaterialBaseMPM.cpp:548-595): 0 or 1 Linear (closed-form
 *                   return; fused path), else returned by the bracketed Newton's method of HardeningLawBase (per-task kernels):
 *                   2 Nonlinear  yldred (1 + beta alpha)^n   [17] beta [18] n
 *                   6 Nonlinear2 yldred (1 + beta alpha^n)   [17] beta [18] n
 *                   3 JohnsonCook  [17] Bred [18] n [19] C [20] ep0 [21] D [22] n2 [23] Tm [24] m [25] reference temperature
 *                                  [26] edotMin [27] eminTerm  (JohnsonCook.cpp:106-127)
 *                   4 SCGL  min(yldred (1 + beta alpha)^n, yldMaxred) Gratio, Gratio = max(0, 1 + GPpred P + GTp (T - Tref)) also
 *                           scales Gred  [17] beta [18] n [19] yldMaxred [20] GPpred [21] GTp [25] reference temperature
 *                           (SCGLHardening.cpp:72-91, :138-198; IsoPlasticity.cpp:548-549)
 *  RIGIDBC:         [8] direction bits (1 x, 2 y, 4 z: RigidMaterial setDirection)  [9] mirrored (-1, 0, +1)
 *                   [10] != 0: sets the temperature (RigidMaterial::setTemperature): with conduction the first such particle to reach a
 *                   node without a grid temperature BC holds it at its own temperature (mpmgpu_particles.temperature of the rigid
 *                   particle; ProjectRigidBCsTask.cpp:118-125)
 */
typedef struct mpmgpu_material {
    int kind;
    int n_history;          /* doubles of history per particle (MaterialBase::NumberOfHistoryDoubles) */
    double p[MPMGPU_MAT_NPARAMS];
} mpmgpu_material;

/* Host structure-of-arrays view of the particles.  Any pointer may be NULL: on upload the field is
 * then zero (or 1 for temperature-like defaults noted below); on download it is skipped.
 * n particles; nonrigid particles first, rigid-BC particles last (NairnMPM::ReorderParticles,
 * NairnMPM.cpp:1117-1194); n_nonrigid of them are nonrigid. */
typedef struct mpmgpu_particles {
    int n;
    int n_nonrigid;
    double *pos;        /* [3][n]  MPMBase::pos */
    double *vel;        /* [3][n]  MPMBase::vel */
    double *mp;         /* [n]     MPMBase::mp */
    double *lp;         /* [3][n]  MPMBase::mpm_lp (dimensionless semi-size) */
    int    *in_elem;    /* [n]     MPMBase::inElem (1-based) */
    int    *matnum;     /* [n]     MPMBase::matnum (1-based index into the materials array) */
    double *sp;         /* [6][n]  xx,yy,zz,yz,xz,xy  MPMBase::sp (specific stress) */
    double *pressure;   /* [n]     MPMBase::pressure */
    double *ep;         /* [6][n]  MPMBase::ep  } together the deformation gradient,           */
    double *wrot;       /* [3][n]  xy,xz,yz     } MatPoint3D.cpp:320-336,363-376                */
    double *eplast;     /* [6][n]  MPMBase::eplast (plastic strain, or elastic B for hyperelastic) */
    double *energies;   /* [6][n]  workEnergy,resEnergy,heatEnergy,entropy,plastEnergy,pPreviousTemperature */
    double *history;    /* [MPMGPU_MAX_HISTORY][n] material history doubles */
    double *pfext;      /* [3][n]  MPMBase::pFext (external particle force, MatPtLoadBC) */
    int    *crossings;  /* [n]     MPMBase::elementCrossings */
    double *acc;        /* [3][n]  MPMBase::acc (download only) */
    int    *ids;        /* [n]     optional caller-global particle ids.  Upload: when given (always in slab
                                   mode, where particles migrate between processes) the ids travel with the
                                   particles and downloads come back in DEVICE order with ids filled in;
                                   when NULL particles are identified by their upload index and downloads come
                                   back in upload order. */
    double *temperature;/* [n]     MPMBase::pTemperature.  Upload: give it with conduction (mpmgpu_set_conduction) and whenever particles start
                                   off the temperature of their last strain update (energies[5]): the first particle update then hands the
                                   laws dT = temperature - energies[5] (thermal strains, UpdateParticlesTask.cpp:246-251).  NULL = energies[5].
                                   Giving it switches the run to the per-task kernels. */
} mpmgpu_particles;

/* download masks */
#define MPMGPU_F_POS      0x001
#define MPMGPU_F_VEL      0x002
#define MPMGPU_F_STRESS   0x004   /* sp + pressure */
#define MPMGPU_F_STRAIN   0x008   /* ep + wrot */
#define MPMGPU_F_EPLAST   0x010
#define MPMGPU_F_ENERGY   0x020
#define MPMGPU_F_HISTORY  0x040
#define MPMGPU_F_ELEM     0x080   /* in_elem + crossings */
#define MPMGPU_F_ACC      0x100
#define MPMGPU_F_ALL      0x1ff
#define MPMGPU_F_TEMPERATURE 0x200 /* pTemperature (runs that uploaded temperatures or use conduction; not part of MPMGPU_F_ALL) */

/* Host view of the node accumulators (debug / global quantities / parity tests).
 * Arrays are nnodes long (vectors [3][nnodes]), reference node order; NULL = skip. */
typedef struct mpmgpu_nodes {
    int nnodes;
    int    *number_points;  /* MatVelocityField::numberPoints (per-task path); the fused path keeps only the 0/1 activity flag numberPoints>0 */
    double *mass;           /* MatVelocityField::mass */
    double *pk;             /* [3][nnodes] momentum */
    double *ftot;           /* [3][nnodes] force */
    double *vk;             /* [3][nnodes] vk[0] */
    double *pk_copy;        /* [3][nnodes] vk[pkCopy] */
    /* multimaterial mode only (mpmgpu_set_multimaterial): every array above and below is n_fields * (grid nodes) long,
     * field-major -- material velocity field f of node i at [f * nodes + i] -- and nnodes returns that length */
    double *contact_volume;   /* MatVelocityField::contactInfo->cvolume */
    double *contact_gradient; /* [3][nnodes] volume gradient (terms[volumeGradientIndex]) */
    double *contact_disp;     /* [3][nnodes] mass-weighted displacement or position (contactByDisplacements) */
    /* conduction only (mpmgpu_set_conduction): NodalPoint::gCond, one value per GRID node (also in multimaterial mode) */
    double *transport_value;    /* gTValue: nodal temperature (sum mp Cv T S before the division of task 3) */
    double *transport_capacity; /* gVCT: sum mp Cv S */
    double *transport_rate;     /* gQ: heat flow into the node, the temperature rate after the momentum update */
} mpmgpu_nodes;

/* ---- life cycle -------------------------------------------------------------------------- */
int mpmgpu_abi_version(void);
int mpmgpu_create(const mpmgpu_config *cfg, mpmgpu_ctx **out);
int mpmgpu_destroy(mpmgpu_ctx *ctx);
const char *mpmgpu_last_error(const mpmgpu_ctx *ctx);   /* ctx may be NULL for create() failures */

/* ---- set-up (host -> device) ------------------------------------------------------------- */
int mpmgpu_set_materials(mpmgpu_ctx *ctx, int nmat, const mpmgpu_material *mats);
/* Multimaterial mode (<MultiMaterialMode>, fmobj->multiMaterialMode): every node carries one velocity field per material
 * field (CrackVelocityFieldMulti::mvf[], Nodes/CrackVelocityFieldMulti.cpp:51-191 with a single crack field) and nodes seen by
 * two or more materials get material contact after the mass/momentum extrapolation, the momentum update and the
 * re-extrapolation of USAVG+/USL+ (MaterialContactNode::ContactOnKnownNodes -> MaterialContactOnCVFLumped,
 * CrackVelocityFieldMulti.cpp:302-674; UpdateMomentaTask::ContactAndMomentaBCs).  Built: nonrigid materials, normals from the
 * volume gradients (methods 0-3, GetNormalVector :960-1071) or specified (4), contact detected by displacements or positions
 * (CrackSurfaceContact::MaterialSeparation), contact laws ignore / stick / frictionless / Coulomb friction with optional static
 * coefficient (Materials/CoulombFriction.cpp:150-272), three or more materials lumped as the reference does, rigid contact
 * materials (MPMGPU_MAT_RIGIDCONTACT).  Refused: the regression normals (5, 6), imperfect interfaces, adhesion, XPIC/FMPM
 * order > 1 (FMPM contact increments), slab mode.  Runs on the per-task kernels.  Call after mpmgpu_set_materials and before mpmgpu_upload_particles;
 * displacements are taken against the original positions of mpmgpu_set_archive_origin (default: the positions at upload). */
typedef struct mpmgpu_multimaterial {
    int n_fields;                   /* maxMaterialFields: material velocity fields per node (<= 8) */
    const int *field_of_material;   /* [nmat] MaterialBase::GetField() of every material (ignored for rigid-BC materials) */
    int normal_method;              /* mpmgrid.materialNormalMethod: 0 MAXG, 1 MAXV, 2 AVGG, 3 OWNG, 4 SN (MeshInfo.hpp:31) */
    int contact_by_displacements;   /* mpmgrid.contactByDisplacements */
    double position_cutoff;         /* mpmgrid.positionCutoff (<ContactPosition>; < 0: the power-law form) */
    double contact_normal[3];       /* mpmgrid.contactNormal (method 4) */
    const int *law_kind;            /* [n_fields][n_fields] mpmgrid.GetMaterialContactLaw(i, j): 0 ignore, 1 stick, 2 frictionless, 3 Coulomb friction */
    const double *law_friction;     /* [n_fields][n_fields] CoulombFriction::frictionCoeff (NULL: 0) */
    const double *law_static;       /* [n_fields][n_fields] frictionCoeffStatic, <= 0 for none (NULL: none) */
    double rigid_gradient_bias;     /* mpmgrid.rigidGradientBias as the reference holds it after set-up (RigidBias squared,
                                       MeshInfo.cpp:1196): preference for the rigid material's normal; 0 = 1 */
} mpmgpu_multimaterial;
int mpmgpu_set_multimaterial(mpmgpu_ctx *ctx, const mpmgpu_multimaterial *mm);
/* Heat conduction, the first transport task (<Thermal><Conduction/></Thermal>; Custom_Tasks/ConductionTask.cpp + TransportTask.cpp,
 * hooked into the step at NodalPointMPM.cpp:444-447, PostExtrapolationTask.cpp:88,160, GridForcesTask.cpp:108-112,
 * UpdateMomentaTask.cpp:55, UpdateParticlesTask.cpp:134-245): nodal temperature sum(mp Cv T S)/sum(mp Cv S), temperature gradient
 * on the particles, conduction flow -mp (Vp/V0) (k/rho) grad T . grad S, grid rate and value update, FLIP update of the particle
 * temperature, heat energy and entropy of the conducted heat; the laws see the grid-extrapolated temperature as their
 * previous temperature.  kcond[m] = conductivity / rho of material m in the host's units (TransportProperties::kCondTensor,
 * isotropic: MaterialBaseMPM.cpp:320-326; ignored for rigid-BC materials).  Built: isothermal energy mode, insulated boundaries,
 * any number of materials (also in multimaterial mode: transport values live on the node, not on a velocity field), nodal
 * temperature BCs (mpmgpu_set_temperature_bcs), thermal expansion (the temperature change of a step reaches the laws as ResidualStrains::dT).  Refused: thermal expansion on the
 * large-rotation IsotropicMat, slab mode; silent / coupled heat-flux BCs, the transport task's own XPIC option and contact heating are the adapter's to refuse (a mechanical XPIC/FMPM order > 1 leaves the transport update FLIP, as in the reference).  Per-task kernels.  Call after
 * mpmgpu_set_materials and before mpmgpu_upload_particles. */
int mpmgpu_set_conduction(mpmgpu_ctx *ctx, int nmat, const double *kcond);
/* <EnergyCoupling/> (ConductionTask::adiabatic): MaterialBase::IncrementHeatEnergy buffers dTq0 + dPhi/Cv as a temperature rise on
 * the particle instead of releasing it as heat (MaterialBaseMPM.cpp:982-1006); the next particle update adds the buffer to the
 * particle's temperatures and, with conduction, to the dT the laws see (UpdateParticlesTask.cpp:229-235) -- plastic work heats the
 * material (thermal softening in Johnson-Cook hardening).  With or without conduction; per-task kernels.  Before
 * mpmgpu_upload_particles. */
int mpmgpu_set_energy_coupling(mpmgpu_ctx *ctx, int adiabatic);
/* Nodal temperature BCs (NodalTempBC list, firstTempBC ...; <TempBC> in <GridBCs>) in the host's list order: node[i] 1-based,
 * value[i] = BCValue at this step's time, active[i] = GetNodeNum(time) != 0 (NULL: all active).  Before the temperature gradients
 * are taken the BC nodes hold the sum of their BC values and get their own value back afterwards (TransportTask::ImposeValueBCs /
 * RestoreValueBCs, TransportTask.cpp:167-222); after the grid update the nodal value becomes that sum and the rate what takes it
 * there (ImposeValueGridBCs :316-404).  The heat the BCs feed in (NodalValueBC::qreaction, a global-quantity input) is not
 * tracked.  Call after mpmgpu_set_conduction; call again whenever values change. */
int mpmgpu_set_temperature_bcs(mpmgpu_ctx *ctx, int n, const int *node, const double *value, const int *active);
/* Rigid particles of a material that sets the temperature (material slot 10) through a value function of time and position
 * (RigidMaterial::GetValueSetting, which ProjectRigidBCsTask.cpp:118-121 evaluates for every particle every step): the host does that
 * and hands over pTemperature of all rigid particles, host order, before the step. */
int mpmgpu_update_rigid_temperatures(mpmgpu_ctx *ctx, int n_rigid, const double *temperature);
int mpmgpu_upload_particles(mpmgpu_ctx *ctx, const mpmgpu_particles *host);
/* Particle traction BCs (MatPtTractionBC list, firstTractionPt ...; <TractionBC> in <ParticleBCs>): entry i loads face[i] of the
 * domain of particle particle[i] (0-based host index, non-rigid) in direction[i] -- 1 x, 2 y, 3 z (3D), 11 normal to the deformed
 * face, 12 along it (2D) -- with the stress value[i] = BCValue at this step's time.  Faces: 2D 1 bottom, 2 right, 3 top, 4 left;
 * 3D 1 -y, 2 +x, 3 +y, 4 -x, 5 -z, 6 +z.  In the post-forces task (PostForcesTask.cpp:51) each of the face's corners (2 or 4,
 * MatPoint2D/3D::GetSurfaceInfo: the deformed domain for the CPDI shapes, the undeformed one weighted with the deformed face size
 * for the others) hands direction x area/corners x value x N_i to the nodes of its element that carry the particle's material
 * (MatPtTractionBC::AddMPFluxBC, MatPtTractionBC.cpp:64-226).  With the B-spline shapes the corner's weights are the quadratic splines of its element (ElementBase::GetShapeFunctionsForTractions).  Not built: axisymmetric, <ExactTractions>, slab mode.
 * Call after mpmgpu_upload_particles; mpmgpu_update_particle_traction_values hands over new values of the same list. */
int mpmgpu_set_particle_tractions(mpmgpu_ctx *ctx, int n, const int *particle, const int *face, const int *direction, const double *value);
int mpmgpu_update_particle_traction_values(mpmgpu_ctx *ctx, int n, const double *value);
/* Particle heat-flux BCs (MatPtHeatFluxBC list, firstHeatFluxPt ...; <HeatFluxBC dir="1"> in <ParticleBCs>: external flux): the same
 * walk over a face's corners at the end of the post-forces task (TransportTask::TransportForceBCs, PostForcesTask.cpp:97); value[i]
 * (energy per time and area at this step's time) x face area / corners x N_i goes into the conduction rate of every node around the
 * corner that carries non-rigid particles (MatPtHeatFluxBC::AddMPFluxBC, MatPtHeatFluxBC.cpp:64-160; TransportTask::AddFluxCondition).
 * Silent and coupled (dir="2", a function of the particle temperature) fluxes are the adapter's to refuse.  After
 * mpmgpu_set_conduction and mpmgpu_upload_particles. */
int mpmgpu_set_particle_heat_fluxes(mpmgpu_ctx *ctx, int n, const int *particle, const int *face, const double *value);
int mpmgpu_update_particle_heat_flux_values(mpmgpu_ctx *ctx, int n, const double *value);
/* Reaction forces of the velocity BCs (NodalVelBC::freaction, the input of the "reactionx/y/z" global quantities:
 * GlobalQuantity.cpp:971-986 -> NodalVelBC::TotalReactionForce, NodalVelBC.cpp:246-252).  Each BC's freaction starts from zero in
 * the grid-forces pass and collects the force that pass adds to the node's material fields, -(ftot.n + pk.n/dt) n for the zeroing
 * and m v/dt n for the imposed value (MatVelocityField.cpp:504-524, :563-569), plus with XPIC/FMPM order > 1 the lumped
 * -m dv*.n/dt n of the particle-update pass (:529-536).  mpmgpu_track_reactions(ctx, 1) turns the bookkeeping on (after
 * mpmgpu_set_materials; not in slab mode).  mpmgpu_download_reactions: bc_reaction[3*i..] = freaction of entry i of the list given
 * to mpmgpu_set_velocity_bcs (n = its length; the host sums by bcID), rigid_reaction[3*m..] = the summed freaction of the BCs made
 * by the rigid-BC particles of material m (0-based; their bcID is the material number, ProjectRigidBCsTask.cpp:241).  Either may
 * be NULL.  The values are those of the last completed step. */
int mpmgpu_track_reactions(mpmgpu_ctx *ctx, int on);
/* Contact forces on rigid contact materials (the "contactx/y/z" global quantities, GlobalQuantity.cpp:905-968 ->
 * NodalPoint::AddGetContactForce, NodalPointMPM.cpp:1458-1472): force[3*f..] = the summed force row of rigid material field f over
 * the nodes where it is active now -- the momentum its contacts gave the other materials since the last clearing
 * (MatVelocityField::AddContactForce); 0 for non-rigid fields; n_fields x 3 doubles.  clear != 0 zeroes the rows that were read, as
 * the reference does outside VTK archiving.  The caller scales by 1/(steps since the last clearing x timestep).  Multimaterial mode. */
int mpmgpu_contact_forces(mpmgpu_ctx *ctx, int clear, double *force);
int mpmgpu_download_reactions(mpmgpu_ctx *ctx, int n, double *bc_reaction, double *rigid_reaction);
/* timestep, strainTimestepFirst, strainTimestepLast (NairnMPM.cpp:1207-1240) */
int mpmgpu_set_time_step(mpmgpu_ctx *ctx, double dt, double dt_strain_first, double dt_strain_last);
/* XPIC/FMPM order can change per step (Custom_Tasks/PeriodicXPIC.cpp:161-240) */
int mpmgpu_set_xpic(mpmgpu_ctx *ctx, int order, int using_fmpm);
/* Grid velocity BCs in the host's list order (firstVelocityBC..., Boundary_Conditions/NodalVelBC.cpp:321-400):
 * node[i] 1-based, norm[3*i..] unit direction, value[i] = currentValue at this step's mtime
 * (BoundaryCondition::BCValue), active[i] = GetNodeNum(mtime)>0, symdir[i] = the node's
 * fixedDirection symmetry-plane bits (32|64|128, ADJUST_COPIED_PK) or 0.  Call again whenever values change. */
int mpmgpu_set_velocity_bcs(mpmgpu_ctx *ctx, int n, const int *node, const double *norm,
                            const double *value, const int *active, const int *symdir);
/* BCs next to a symmetry plane (<Horiz symmin=...>: Read_MPM/Generators.cpp:2178-2190) reflect the velocity of the node
 * across the plane instead of imposing value[i]: v = value + ratio (value - n.pk_r/m_r) when node r has particles
 * (NodalVelBC::AddVelocityBC -> NodalPoint::ReflectVelocityBC, NodalVelBC.cpp:196-210, CrackVelocityFieldSingle.cpp:133-144).
 * reflected_node[i] = NodalVelBC::reflectedNode (1-based, <= 0 for a plain BC), ratio[i] = reflectRatio; same n and order
 * as the list above; call after mpmgpu_set_velocity_bcs.  Runs on the per-task kernels (kernel_path 2 is refused). */
int mpmgpu_set_velocity_bc_reflections(mpmgpu_ctx *ctx, int n, const int *reflected_node, const double *ratio);
/* update only the values/active flags of the BC list set above (same n, same order) */
int mpmgpu_update_velocity_bc_values(mpmgpu_ctx *ctx, int n, const double *value, const int *active);
/* Particle load BCs (MatPtLoadBC): the reference re-evaluates them at the start of every step (MatPtLoadBC::SetParticleFext,
 * InitializationTask.cpp:91, MatPtLoadBC.cpp:210-222).  The host evaluates the BCs at this step's time and hands over the
 * external force fext[3][n_loaded] of the loaded particles only; particle[k] = 0-based index (upload order) of loaded particle k
 * on the first call, NULL afterwards (same particles).  The particles must have been uploaded with a pfext array. */
int mpmgpu_update_particle_loads(mpmgpu_ctx *ctx, int n_loaded, const int *particle, const double *fext);
/* velocities [3][n_rigid] of the rigid-BC particles (host order) for this step: the host evaluates the material's
 * setting functions (RigidMaterial::GetVectorSetting, Materials/RigidMaterial.cpp:376-531) before the projection task */
int mpmgpu_update_rigid_velocities(mpmgpu_ctx *ctx, int n_rigid, const double *vel);

/* ---- the step ---------------------------------------------------------------------------- */
/* nsteps full MPMSteps (tasks 1-9, 11) with the configured method; mtime advances by dt each step */
int mpmgpu_step(mpmgpu_ctx *ctx, int nsteps);
/* mpmgpu_step reads the device's status word (position NaN, CPDI corner off the grid: the reference's exceptions of
 * ResetElementsTask.cpp:200-203 and MatPoint3D.cpp:596-601) at the end of every call, which costs one stream synchronisation.
 * With k > 1 it does so every k-th call only: the host keeps enqueueing steps, an error is reported up to k-1 steps late and
 * mpmgpu_left_grid_counts returns the counts of the last poll in between.  Default 1 (the reference's behaviour). */
int mpmgpu_set_poll_interval(mpmgpu_ctx *ctx, int k);

/* The same step as separate entry points named after the reference's tasks, so that the
 * reference's per-task timing report (MPMTask.cpp:106-121) stays meaningful and each task can be
 * checked in isolation.  Call in pipeline order. */
int mpmgpu_task_initialization(mpmgpu_ctx *ctx);        /* InitializationTask.cpp:43-103 */
int mpmgpu_task_mass_and_momentum(mpmgpu_ctx *ctx);     /* MassAndMomentumTask.cpp:48-127 */
int mpmgpu_task_project_rigid_bcs(mpmgpu_ctx *ctx);     /* ProjectRigidBCsTask.cpp:39-158 (no-op without rigid-BC particles) */
int mpmgpu_task_post_extrapolation(mpmgpu_ctx *ctx);    /* PostExtrapolationTask.cpp:46-165 */
int mpmgpu_task_update_strains_first(mpmgpu_ctx *ctx);  /* UpdateStrainsFirstTask.cpp:52-168 */
int mpmgpu_task_grid_forces(mpmgpu_ctx *ctx);           /* GridForcesTask.cpp:40-153 */
int mpmgpu_task_post_forces(mpmgpu_ctx *ctx);           /* PostForcesTask.cpp:43-111 */
int mpmgpu_task_update_momenta(mpmgpu_ctx *ctx);        /* UpdateMomentaTask.cpp:44-64 */
int mpmgpu_task_update_particles(mpmgpu_ctx *ctx);      /* UpdateParticlesTask.cpp:44-298 */
int mpmgpu_task_update_strains_last(mpmgpu_ctx *ctx);   /* UpdateStrainsLast(Contact)Task.cpp */
int mpmgpu_task_reset_elements(mpmgpu_ctx *ctx);        /* ResetElementsTask.cpp:44-265 */

/* ---- device -> host ---------------------------------------------------------------------- */
int mpmgpu_download_particles(mpmgpu_ctx *ctx, mpmgpu_particles *host, unsigned mask);
int mpmgpu_download_nodes(mpmgpu_ctx *ctx, mpmgpu_nodes *host);
int mpmgpu_synchronize(mpmgpu_ctx *ctx);
/* step counter, simulated time, particles that crossed an element boundary / left the grid so far */
int mpmgpu_get_status(mpmgpu_ctx *ctx, long long *mstep, double *mtime, long long *crossings, long long *left_grid);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long mpmgpu_launch_count(const mpmgpu_ctx *ctx);
/* CUDA stream the context launches on (as a cudaStream_t cast to void*), for event timing */
void *mpmgpu_stream(mpmgpu_ctx *ctx);
/* per-task device time (the reference's task timing report, MPMTask.cpp:100-121): when profiling is
 * on every task is bracketed by CUDA events (this serialises the stream; leave it off when timing
 * whole steps).  mpmgpu_task_times fills ms[10] (accumulated) and calls[10] in task order above. */
int mpmgpu_set_profiling(mpmgpu_ctx *ctx, int on);
int mpmgpu_task_times(mpmgpu_ctx *ctx, double *ms, long long *calls);

/* ---- slab decomposition across the GPUs of one box (one process per GPU) ------------------------
 * Replaces the reference's GridPatch/GhostNode OpenMP decomposition (Patches/GridPatch.cpp:32-138,
 * Patches/GhostNode.cpp:58-185).  Every process creates a context over the WHOLE grid (node and element
 * numbers stay global, so indexing is bit-identical to the single-GPU run) but holds only the particles
 * whose element lies in its cell planes [cell_lo, cell_hi) along z, and touches only the node planes
 * [cell_lo-1, cell_hi+2).  After each of the three particle->grid passes the partial node sums of the
 * three node planes around an interior slab face are swapped with the neighbour and added (both sides
 * then hold identical complete sums and run the node sweeps redundantly on those planes -- no
 * broadcast step).  The library packs/adds; the HOST moves the buffers (NCCL send/recv via
 * torch.distributed, see nairn_mpm_fea_b200/slab.py).  Particles whose new element leaves the slab are
 * listed by the last phase and moved as rows. */
int mpmgpu_slab_configure(mpmgpu_ctx *ctx, int cell_lo, int cell_hi, int has_lower, int has_upper, int migration_capacity);
/* device pointers of the halo exchange buffers ([lower, upper] neighbour); a pass moves
 * nvalues*3*plane_nodes doubles per neighbour, nvalues = 4, 3, 3 for the exchanges after phases 0, 1, 2
 * (and 3 for the exchange inside an XPIC/FMPM iteration, halo kind 3) */
int mpmgpu_slab_halo_buffers(mpmgpu_ctx *ctx, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, long long *plane_nodes);
/* phase 0: zero + P2G mass/momentum | exchange | 1: node sweep + G2P strain + P2G forces | exchange |
 * 2: node sweep + G2P update + P2G momentum | exchange | 3: node sweep + G2P strain + element reset.
 * Phases 0-2 return with the send buffers packed and the stream idle; the next phase adds the recv buffers. */
int mpmgpu_slab_step_phase(mpmgpu_ctx *ctx, int phase);
/* XPIC(k)/FMPM(k), k > 1 (XPICExtrapolationTask.cpp:49-161) needs one more exchange per iteration, in the middle of a
 * phase: the library packs the send buffers, calls fn(user, 3) -- the host enqueues the swap of halo kind 3 on the
 * context's stream, exactly as it does between phases -- and then adds the recv buffers.  fn must not throw. */
typedef void (*mpmgpu_halo_fn)(void *user, int which);
int mpmgpu_slab_set_halo_callback(mpmgpu_ctx *ctx, mpmgpu_halo_fn fn, void *user);
/* after phase 3: how many particles must move to the lower / upper neighbour */
int mpmgpu_slab_migration_counts(mpmgpu_ctx *ctx, int *n_lo, int *n_hi);
int mpmgpu_slab_migration_buffers(mpmgpu_ctx *ctx, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, int *row_doubles, int *capacity_rows);
int mpmgpu_slab_pack_migrants(mpmgpu_ctx *ctx);            /* rows -> send buffers */
int mpmgpu_slab_finish_migration(mpmgpu_ctx *ctx, int n_from_lo, int n_from_hi);  /* drop leavers, append arrivals */

/* The same exchanges done by the library itself over NCCL (ncclSend/ncclRecv with the lower and upper z-neighbour on the
 * context's stream; libnccl.so.2 is resolved at run time): the multi-GPU form of the reference's shared-memory patch
 * reductions and particle moves (Patches/GhostNode.cpp:127-185, Patches/GridPatch.cpp:214-251).
 *   mpmgpu_nccl_unique_ids   rank 0: 2 x 128 bytes (ncclUniqueId of the data and of the count communicator); the caller hands
 *                            them to every rank (MPI, torch.distributed, a file ...)
 *   mpmgpu_slab_connect      every rank, after mpmgpu_slab_configure: joins the two communicators (collective)
 *   mpmgpu_slab_step         nsteps full steps: phases 0-3 with the three halo exchanges (and the one inside every XPIC/FMPM
 *                            iteration) in stream order, then the migration -- the two-integer count handshake runs on the
 *                            second communicator and stream while the last strain kernel is busy.  Every rank calls it with
 *                            the same nsteps.
 *   mpmgpu_slab_migrated     particle rows sent / received so far */
int mpmgpu_nccl_unique_ids(void *ids256);
int mpmgpu_slab_connect(mpmgpu_ctx *ctx, int rank, int world, const void *ids256);
int mpmgpu_slab_step(mpmgpu_ctx *ctx, int nsteps);
int mpmgpu_slab_migrated(const mpmgpu_ctx *ctx, long long *rows_out, long long *rows_in);
/* particles that left the grid and were pushed back (ResetElementsTask.cpp:71-151,232-265) since the upload: `exits` counts
 * every push-back, `particles` the particles leaving for the first time -- the events the reference issues its
 * "Particle has left the grid" warning for (abort threshold <LeaveLimit>, NairnMPM.cpp:814-832); the host feeds its own
 * MPMWarnings with the increase of `particles`.  Synchronises the stream. */
int mpmgpu_left_grid_counts(mpmgpu_ctx *ctx, long long *exits, long long *particles);
int mpmgpu_num_particles(const mpmgpu_ctx *ctx);
/* launch on the caller's CUDA stream (cudaStream_t as void*; NULL = back to the context's own), so the
 * host's NCCL calls and the kernels are ordered on one stream without host synchronisation */
int mpmgpu_set_stream(mpmgpu_ctx *ctx, void *cuda_stream);

/* ---- output side on the device --------------------------------------------------------
 * Particle-archive records in the reference's binary format (ArchiveData::ArchiveResults, System/ArchiveData.cpp:806-1100;
 * record size CalcArchiveSize :328-396), packed on the device in the caller's particle order: an archive step moves one
 * record block instead of the whole state.  `order` is the <MPMArchiveOrder> string (ArchiveData.hpp:22-32).  Items this
 * path does not produce (shear components, damage normal, spin, history 5-19, particle size) give MPMGPU_EINVAL.
 * The 64-byte file header (ArchiveData.cpp:464-489) is the host's to write (nairn_mpm_fea_b200/archive.py::header). */
int mpmgpu_archive_record_size(const mpmgpu_ctx *ctx, const char *order);      /* bytes per particle, or -1 */
/* constants the records carry and the step does not: original positions [3][n] and initial material angles [3][n]
 * (z, y, x; radians) in the caller's order; thickness: 2D particle thickness.  Defaults without this call: the positions
 * at mpmgpu_upload_particles, zero angles, mpmgpu_config.thickness.  Either array may be NULL (keeps the default).
 * Call after mpmgpu_upload_particles. */
int mpmgpu_set_archive_origin(mpmgpu_ctx *ctx, const double *origpos, const double *angles0, double thickness);
/* The "temperature" item writes the temperature of the particle's last strain update; this path is isothermal (particle
 * temperatures other than the stress-free one are refused by the adapter), so it equals ArchiveData.cpp:976's pTemperature.
 * One slab of a multi-GPU run packs ITS particles in device order (mpmgpu_download_ids gives the ids in the same order);
 * the original-position column then repeats the current position (the caller's constants are indexed by its own order). */
int mpmgpu_pack_archive(mpmgpu_ctx *ctx, const char *order, void *records, size_t capacity_bytes);
/* particle ids in device order: the ids handed over at upload (slab mode) or the caller's particle index */
int mpmgpu_download_ids(mpmgpu_ctx *ctx, int *ids, int capacity);

/* Raw sums behind the reference's GlobalQuantity rows (Global_Quantities/GlobalQuantity.cpp:394-1075) over the non-rigid
 * particles, per material: sums[m * MPMGPU_GS_NSUMS + k], internal units; the host divides the volume-weighted ones by
 * MPMGPU_GS_VOLUME and applies the reference's unit scalings.  Fixed summation order (no atomics): repeatable. */
#define MPMGPU_GS_MASS           0   /* sum mp */
#define MPMGPU_GS_VOLUME         1   /* sum Vp, Vp = J mp / rho0 */
#define MPMGPU_GS_LINMOM         2   /* 2..4   sum mp v                 (LINMOMX/Y/Z) */
#define MPMGPU_GS_KINETIC        5   /* sum mp |v|^2 / 2                (KINE_ENERGY) */
#define MPMGPU_GS_WORK           6   /* sum mp workEnergy               (WORK_ENERGY) */
#define MPMGPU_GS_STRAIN_ENERGY  7   /* sum mp (work - residual energy) (STRAIN_ENERGY) */
#define MPMGPU_GS_HEAT           8   /* sum mp heatEnergy               (HEAT_ENERGY) */
#define MPMGPU_GS_ENTROPY        9   /* sum mp entropy                  (ENTROPY_ENERGY) */
#define MPMGPU_GS_PLASTIC       10   /* sum mp plastEnergy              (PLAS_ENERGY) */
#define MPMGPU_GS_STRESS        11   /* 11..16 sum mp (total specific stress) xx yy zz yz xz xy   (AVG_Sij = this / volume) */
#define MPMGPU_GS_VOL_VEL       17   /* 17..19 sum Vp v                 (AVG_VELX/Y/Z = this / volume) */
#define MPMGPU_GS_VOL_F         20   /* 20..28 sum Vp F (row-major)     (AVG_Fij = this / volume) */
#define MPMGPU_GS_NSUMS         29
int mpmgpu_global_sums(mpmgpu_ctx *ctx, double *sums /* [nmat][MPMGPU_GS_NSUMS] */);

#ifdef __cplusplus
}
#endif
#endif /* MPMGPU_H */
