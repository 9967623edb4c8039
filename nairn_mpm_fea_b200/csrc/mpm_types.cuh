// Device-side data layout of libmpmgpu (structure-of-arrays, FP64) and small helpers.
//
// Particle fields mirror MPMBase (reference MPM_Classes/MPMBase.hpp:34-264); node fields mirror the
// single material velocity field cvf[0]->mvf[0] (reference Nodes/MatVelocityField.hpp:44-48).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MPM_MAX_HISTORY 4
#define MPM_MAT_NPARAMS 32
#define MPM_MAX_MATERIALS 16

// reference codes (System/MPMPrefix.hpp:110-140)
enum { NP_PLANE_STRAIN = 10, NP_PLANE_STRESS = 11, NP_THREED = 12 };
enum { METHOD_USF = 0, METHOD_USAVG = 2, METHOD_USL = 3 };
enum { SHAPE_LINEAR = 0, SHAPE_UGIMP = 1, SHAPE_B2GIMP = 5, SHAPE_B2SPLINE = 6, SHAPE_LCPDI = 10, SHAPE_QCPDI = 11, SHAPE_B2CPDI = 13,
       // internal variants of the two CPDI codes: same functions, corner contributions to one node merged before the node is
       // touched (shape.cuh::for_each_node_cpdi_merged); chosen at launch, never seen through the ABI
       SHAPE_LCPDI_MERGED = 20, SHAPE_QCPDI_MERGED = 21 };
#define SHAPE_IS_QCPDI(S) ((S) == SHAPE_QCPDI || (S) == SHAPE_QCPDI_MERGED)
#define SHAPE_IS_CPDI(S) ((S) == SHAPE_LCPDI || (S) == SHAPE_LCPDI_MERGED || (S) == SHAPE_B2CPDI || SHAPE_IS_QCPDI(S))
#define SHAPE_IS_MERGED(S) ((S) == SHAPE_LCPDI_MERGED || (S) == SHAPE_QCPDI_MERGED)
enum { MAT_NONE = 0, MAT_ISOTROPIC = 1, MAT_MOONEY = 8, MAT_ISOPLASTICITY = 9, MAT_RIGIDBC = 11, MAT_NEOHOOKEAN = 28,
       MAT_RIGIDCONTACT = 35 };     // RigidMaterial in contact mode (SetDirection 8, RIGID_MULTIMATERIAL_MODE): multimaterial mode only

// BC pass types (reference NodalVelBC.cpp:321-380)
enum { PASS_MASS_MOMENTUM = 0, PASS_GRID_FORCES = 1, PASS_UPDATE_MOMENTUM = 2, PASS_UPDATE_STRAINS_LAST = 3,
       PASS_XPIC = 4 };

struct Grid {
    int dim;                 // 2 or 3
    int np;                  // analysis type
    int horiz, vert, depth;  // cells per axis incl. border (depth>=1; 1 in 2D)
    int yplane, zplane;      // node strides (xplane = 1): MeshInfo.cpp:1510-1512
    int nnodes, nelems;
    const double *xpts, *ypts, *zpts;   // node coordinates per axis (device)
    double gx, gy, gz;       // mpmgrid.grid
    double xmin, ymin, zmin; // mpmgrid.xmin... (= xpts[0]...)
    double rcrit;            // CPDI critical radius or <0
    int lpUniform;           // every particle has the same dimensionless size lpU (skips the lp loads)
    double lpU[3];
    double lpInvSize[3];     // 1/(4 lpU)
    double lpInv2[3];        // 1/(2 lpU)
    double inv2d[3];         // 2/grid cell size per axis
};

// one velocity field per node, component-major
struct Nodes {
    double *mass;            // [nnodes]
    double *pk[3];           // momentum
    double *ftot[3];         // force
    double *vk[3];           // vk[0] (velocity for strain update / particle update)
    double *pkc[3];          // vk[pkCopy]: momentum saved after the first extrapolation
    double *vsp[3];          // XPIC/FMPM: vk[VSTARPREV]
    double *vsn[3];          // XPIC/FMPM: vk[VSTARNEXT]
    int *cnt;                // numberPoints
};

struct Particles {
    int n;                   // particles resident
    int nNR;                 // nonrigid particles come first
    double *pos[3], *vel[3], *mp, *lp[3];
    double *ncpos[3];        // natural coordinates fixed for the step (MPMBase::ncpos)
    double *F[9];            // deformation gradient, row-major (reference stores ep + wrot)
    double *sp[6];           // xx,yy,zz,yz,xz,xy
    double *pressure;
    double *eplast[6];       // xx,yy,zz,yz,xz,xy
    double *work, *res, *heat, *entropy, *plast, *prevT;
    double *hist[MPM_MAX_HISTORY];
    double *pfext[3];
    double *acc[3];
    int *elem;               // 1-based inElem
    int *key;                // fused path: centre node of the particle's dual cell at the start of the step (set by F1)
    int *mat;                // 0-based material index
    int *cross;              // elementCrossings
    int *orig;               // caller's index of this particle
    // CPDI domain data fixed for the step (CPDIDomain, Common/System/DataTypes.hpp:110-115), only allocated for
    // CPDI shape functions: corner c of particle p at [c][p]
    int *cpElem;             // [ncorner][cap]  1-based element holding the corner
    double *cpXi;            // [ncorner*3][cap] natural coordinates of the corner in that element
    double *cpWg;            // [ncorner*3][cap] gradient weights
    double *cpDom;           // [12][cap] the domain itself in grid units, frozen for the step like the corner data: centre (3), semi-side
                             // vectors (3 x 3); rows after the ncorner*3 rows of cpXi.  3D lCPDI: for_each_node_lcpdi3_hat (shape.cuh)
    size_t cpStride;         // cap
    // conduction (mpmgpu_set_conduction): pTemperature and the temperature gradient of the step (MPMBase::pTemp), NULL = off
    double *temp;
    double *tgrad[3];
    // ResidualStrains::dT of the step (MPMBase::dTrans.dT, set by the particle update and used by the next strain updates), NULL =
    // the particle temperatures never change
    double *dTr;
    double *dTad;            // adiabatic mode (<EnergyCoupling/>): MPMBase::buffer_dTad, the temperature rise the laws have buffered this step
    // multimaterial mode: node-index offset of the particle's material velocity field (field * nnodes), NULL = one field
    const int *foff;
};

// ---- multimaterial mode (<MultiMaterialMode>): CrackVelocityFieldMulti with one crack field -----------------------------
// Every node array holds nf fields, field-major: field f of node i at [f * nnodes + i] (a "virtual node"), so the
// particle<->grid kernels only add the particle's field offset to the node index.  Contact needs three more
// extrapolations per field (NodalPoint::AddMassMomentum, NodalPointMPM.cpp:419-453; ContactTerms of MatVelocityField).
#define MPM_MAX_FIELDS 8
enum { NORMALS_MAXG = 0, NORMALS_MAXV = 1, NORMALS_AVGG = 2, NORMALS_OWNG = 3, NORMALS_SPECIFIED = 4 };   // MeshInfo.hpp:31
enum { LAW_IGNORE = 0, LAW_STICK = 1, LAW_FRICTIONLESS = 2, LAW_FRICTIONAL = 3 };
enum { CALL_MASS_MOMENTUM = 0, CALL_UPDATE_MOMENTUM = 1, CALL_UPDATE_STRAINS_LAST = 2 };
struct ContactNodes {
    double *cvol;            // [nf*nnodes] contactInfo->cvolume
    double *cgrad[3];        // volume gradient (terms[volumeGradientIndex])
    double *cdisp[3];        // mass-weighted displacement (contactByDisplacements) or position
    double *rforce[3];       // [nf*nnodes] a rigid field's ftot: the momentum changes its contact gave the other materials, CUMULATIVE over
                             // the steps (MatVelocityField::Zero leaves it alone; the reference clears it when a contact-force quantity is read)
    int *rcnt;               // [nf*nnodes] rigid contact particles seen by a RIGID material's field (its numberPoints; Nodes::cnt stays
                             // 0 there, so every node kernel of the nonrigid fields passes a rigid field by)
};
// ---- conduction (ConductionTask, the first transport task) ------------------------------------
// NodalPoint::gCond (TransportField): one value per NODE (not per material velocity field)
struct TransportNodes {
    double *gT;              // gTValue: sum mp Cv T S, then the nodal temperature
    double *gVCT;            // gVCT: sum mp Cv S
    double *gQ;              // gQ: heat flow, then the temperature rate
    const double *kcond;     // [nmat] conductivity / rho of every material (TransportProperties::kCondTensor, isotropic)
};

// nodal temperature BCs (NodalTempBC list), grouped by node; entries keep the host's list order inside a node
struct TempBCs {
    int nUnique;
    const int *node;         // [nUnique] 0-based node
    const int *start;        // [nUnique+1] range into value/active
    const double *value;     // [nEntries] BCValue at this step's time
    const int *active;       // [nEntries] GetNodeNum(time) != 0
    double *saved;           // [nUnique] the no-BC nodal value between ImposeValueBCs and RestoreValueBCs
};

struct ContactParams {
    int nf;                  // material velocity fields per node (maxMaterialFields)
    int normalMethod;        // mpmgrid.materialNormalMethod (0..4)
    int byDisplacements;     // mpmgrid.contactByDisplacements
    int cubic;               // 3D cubic / 2D square cells (MeshInfo::GetPerpendicularDistance short cut)
    int rigidMask;           // bit f: field f belongs to a rigid contact material
    double rigidBias;        // mpmgrid.rigidGradientBias, squared as MeshInfo::MaterialOutput leaves it
    double positionCutoff;   // mpmgrid.positionCutoff
    double normal[3];        // SPECIFIED_NORMAL
    int lawKind[MPM_MAX_FIELDS * MPM_MAX_FIELDS];
    double lawFriction[MPM_MAX_FIELDS * MPM_MAX_FIELDS];
    double lawStatic[MPM_MAX_FIELDS * MPM_MAX_FIELDS];
};

struct Material {
    int kind;
    int nhist;
    double p[MPM_MAT_NPARAMS];
};

struct StepParams {
    double dt, dtStrainFirst, dtStrainLast;
    double fractionUSF;
    double gridAlpha, particleAlpha;
    double grav[3];
    int method, skipPost;
    int xpicOrder, usingFMPM;
    int hasGravity;
    int adiabatic;           // ConductionTask::adiabatic
};

// velocity BCs grouped by node (entries keep the host's list order within a node)
struct VelBCs {
    int nUnique;             // nodes with at least one BC
    const int *node;         // [nUnique] 0-based node index
    const int *start;        // [nUnique+1] range into the entry arrays
    const int *symdir;       // [nUnique] symmetry-plane bits of the node
    const double *norm;      // [3*nEntries]
    const double *value;     // [nEntries]
    const int *active;       // [nEntries]
    const int *refl;         // [nEntries] 0-based node whose velocity a symmetry-plane BC reflects, -1 = plain BC; NULL = none at all
    const double *reflRatio; // [nEntries] cell-size ratio across the plane (NodalVelBC::reflectRatio)
    double *reaction;        // [3*nEntries] NodalVelBC::freaction of each entry, summed over the material fields; NULL = not tracked
};

// Velocity BCs made by rigid-BC particles (ProjectRigidBCsTask.cpp:39-158): per node and direction the
// first rigid particle (lowest index) that claims a dof not fixed by a grid BC sets it to its velocity.
#define RIGID_NONE 0x7f7f7f7f
struct RigidBCs {
    int on;
    int *owner[3];           // [nnodes] claiming rigid particle per direction, RIGID_NONE when free
    const double *vel[3];    // rigid particle velocities
    const unsigned char *fixedBits;   // [nnodes] x=1,y=2,z=4 dofs fixed by grid BCs (NodalPoint::fixedDirection), or NULL
    int mirrored;            // some rigid material reflects (RigidMaterial::mirrored): per-task path only
    const int *mat;          // rigid particles' material index
    const Material *mats;
    int stride[3];           // node spacing along x, y, z
    int nnodes;
    double *reaction;        // [3*nmat] freaction of the rigid-particle BCs summed per rigid material (their bcID); NULL = not tracked
    // temperature BCs made by rigid particles whose material sets the temperature (RigidMaterial::setTemperature,
    // ProjectRigidBCsTask.cpp:118-125): the first such particle to reach a node without a grid temperature BC holds it at its own
    int *ownerT;             // [nnodes] claiming rigid particle, RIGID_NONE when free; NULL = no rigid temperature BCs
    const double *ptemp;     // rigid particles' pTemperature
    const unsigned char *fixedT;      // [nnodes] 1 = the node has a grid temperature BC (NodalPoint::fixedDirection & TEMP_DIRECTION), or NULL
    double *savedT;          // [nnodes] the node's own value while the BC value stands in for the gradients
};

// Particle traction BCs (MatPtTractionBC): a stress on one face of a particle's domain, handed to the nodes around the face's corners
struct TractionBCs {
    int n;                   // entries
    const int *start;        // [nParticles+1] entries of host particle i (CSR; the reference walks its list, sums commute)
    const int *face;         // [n] 1..4 (2D: bottom, right, top, left), 1..6 (3D: -y, +x, +y, -x, -z, +z)
    const int *dir;          // [n] 1 x, 2 y, 3 z, 11 normal, 12 tangent (2D)
    const double *value;     // [n] BCValue at this step's time (stress)
};

struct StatusFlags {         // device -> host error reporting (ResetElementsTask.cpp:71-151)
    unsigned long long crossings;
    unsigned long long leftGrid;
    int nanParticle;         // 1 + index of a particle with NaN position
    int cpdiLeft;            // 1 + index of a particle whose CPDI corner left the grid
};

__device__ __forceinline__ void atomAdd(double *a, double v) { atomicAdd(a, v); }
