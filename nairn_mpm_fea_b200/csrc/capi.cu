// libmpmgpu: C ABI implementation (include/mpmgpu.h).  Owns the device-resident SoA particle and
// node state and sequences the kernels of one MPM step in the reference's task order
// (NairnMPM_Class/NairnMPM.cpp:870-1110, :284-335).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpmgpu.h"
#include "mpm_types.cuh"
#include "shape.cuh"
#include "materials.cuh"
#include "kernels_task.cuh"
#include "kernels_fused.cuh"
#include "kernels_pipe.cuh"
#include "archive.cuh"
#include <dlfcn.h>
#include <nccl.h>        // types only: the library is resolved at run time (slab_nccl below)

static std::string g_create_error;

// result of the device-side check of an upload (k_validate_upload): lowest offending particle per rule, or 0x7fffffff
struct UploadCheck { int badMat, badElem, notRigid, rigidEarly, lpLarge, lpVaries; };

enum { T_INIT = 0, T_MASSMOM, T_POSTEXTRAP, T_USF, T_FORCES, T_POSTFORCES, T_MOMENTA, T_PARTICLES, T_USL, T_RESET, T_NTASKS };

struct mpmgpu_ctx {
    mpmgpu_config cfg;
    int dim;
    Grid g;
    Particles P;                        // nonrigid particles (P.n == P.nNR)
    Particles PR;                       // rigid-BC particles, host order, in their own small pool
    double *rigidPool; int *rigidIntPool; size_t rigidCap;
    RigidBCs R;                         // node dofs claimed by rigid particles
    std::vector<unsigned char> hFixedBits;
    Nodes N;
    StepParams sp;
    VelBCs B;
    int nBCEntries;
    int *dBcNode = NULL, *dBcStart = NULL, *dBcSym = NULL, *dBcActive = NULL, *dBcOfNode = NULL; double *dBcNorm = NULL, *dBcValue = NULL;
    int bcCapUnique = 0, bcCapEntries = 0, bcCapRefl = 0;
    int *dBcRefl = NULL; double *dBcReflRatio = NULL;
    bool trackReactions = false;        // mpmgpu_track_reactions: B.reaction / R.reaction are kept
    double *dReaction = NULL;           // [3*reactionCap] grid BC entries (device order)
    size_t reactionCap = 0;
    double *dRigidReaction = NULL;      // [3*nmat]
    std::vector<int> bcOrder;           // entry e on device = host list index bcOrder[e]
    Material *dMats;
    int nmat;
    std::vector<Material> hMats;
    StatusFlags *dFlags;
    StatusFlags hFlags;
    int pollInterval = 1, callsSincePoll = 0;       // mpmgpu_set_poll_interval: status word read (one stream sync) every k-th mpmgpu_step call
    cudaStream_t stream;
    std::vector<void *> allocs;         // everything cudaMalloc'd, for destroy
    double *particlePool;               // one slab for all particle doubles
    int *particleIntPool;
    double *nodePool;
    int *cpElemPool; double *cpXiPool, *cpWgPool;   // CPDI domain data (only for CPDI shape functions)
    size_t cpDomOffset = 0;             // rows of cpXiPool before the 12 rows of Particles::cpDom
    size_t cap;                         // particle capacity
    long long mstep;
    double mtime;
    long long launches;
    bool uploaded, hasFext, hasBCs, hasReflectedBCs;
    cudaStream_t ownStream; bool ownStreamSaved;
    const int *dlSlot, *dlSlotR;        // download slot maps: P.orig / PR.orig, or identity when ids are global
    bool globalIds;                     // particle ids are caller-global (slab mode): downloads come in device order + ids
    bool cpdiMerge = false;             // every CPDI kernel merges the corners' contributions per node (MPMGPU_CPDI_MERGE=1; shape.cuh)
    bool cpdiMergeValues = true;        // the value-only CPDI kernels do (MPMGPU_CPDI_MERGE_VALUES=0 switches it off)
    bool largeRotation = false;         // some material needs the extended law dispatch (Elastic::useLargeRotation, Mooney): per-task kernels, k_update_strains_lr
    double *archOrigin = NULL, *archAngles = NULL;   // [3][n] caller order, for the archive records (mpmgpu_set_archive_origin)
    double archThickness = 1.;
    bool archOriginFromCaller = false;
    size_t archOriginLen = 0;
    UploadCheck *dUploadCheck = NULL;
    int *dLoadOf = NULL; double *dLoadFext = NULL; int loadCap = 0, nLoaded = 0;     // particle loads (mpmgpu_update_particle_loads)
    double *stage = NULL; size_t stageLen = 0;       // staging buffer of uploads/downloads (grown on demand, freed with the context)
    uint32_t *archBuf = NULL; size_t archBufWords = 0;
    double *gsumBuf = NULL; size_t gsumBufLen = 0;
    std::string err;
    // profiling
    bool profiling;
    cudaEvent_t ev0, ev1;
    double taskMs[T_NTASKS];
    long long taskCalls[T_NTASKS];
    // multimaterial mode (mpmgpu_set_multimaterial): nf material velocity fields per node, node arrays field-major
    bool multimaterial = false;
    int nf = 1;                         // fields per node
    int nvn = 0;                        // "virtual" nodes = nf * g.nnodes: the length the per-task node kernels run over
    size_t nodePad = 0;                 // padded length of one node array
    ContactNodes C;                     // contact extrapolations (volume, volume gradient, displacement/position)
    double *contactPool = NULL;
    ContactParams cp;
    int *dFieldOfMat = NULL, *foffPool = NULL, *foffRigidPool = NULL;
    std::vector<int> hFieldOfMat;
    // conduction (mpmgpu_set_conduction): nodal transport field, particle temperature + gradient
    bool conduction = false;
    bool adiabatic = false;             // <EnergyCoupling/> (mpmgpu_set_energy_coupling): the laws buffer a temperature rise instead of releasing heat
    bool thermal = false;               // particle temperatures can change (conduction, or a start off the stress-free temperature): the laws get dT
    TransportNodes T;
    double *transportPool = NULL, *dKcond = NULL, *tempPool = NULL;
    struct FaceBCs {                    // particle BCs on a face of the particle domain: tractions, heat fluxes
        TractionBCs TB;
        int *dStart = NULL, *dFace = NULL, *dDir = NULL; double *dValue = NULL;
        int cap = 0, startLen = 0;
        std::vector<int> order;         // entry e on device = list index order[e]
    } trac, flux;
    bool rigidTemp = false;             // some rigid-BC material sets the temperature: R.ownerT / R.ptemp / R.savedT are allocated
    double *rigidTempPool = NULL;       // pTemperature of the rigid particles
    unsigned char *dFixedTemp = NULL;   // nodes with a grid temperature BC
    std::vector<unsigned char> hFixedTemp;
    TempBCs Q;                          // nodal temperature BCs (mpmgpu_set_temperature_bcs)
    int tbcEntries = 0, tbcCap = 0;
    std::vector<int> tbcOrder;          // entry e on device = host list index tbcOrder[e]
    int *dTbcNode = NULL, *dTbcStart = NULL, *dTbcActive = NULL; double *dTbcValue = NULL, *dTbcSaved = NULL;
    // CUDA graphs of one whole step (mpmgpu_step): keyed by a byte signature of everything the step's launches capture by value
    // (particle / node / BC structs with their device pointers, step parameters, mode flags); a setter that changes any of it,
    // or a physical sort that swaps the particle pools, simply selects or creates another graph.  MPMGPU_GRAPHS=0 switches it off.
    struct StepGraph { std::string sig; cudaGraphExec_t exec; long long launches; };
    std::vector<StepGraph> stepGraphs;
    bool useGraphs = true, inGraphStep = false;
    TiledState tiled;
    bool f2Attr[2][2];
    // slab mode: leave counts + status flags land here (pinned) right after the early element reset; the host
    // waits on slabEvent while the strain kernel of the same step is still running
    mpmgpu_halo_fn haloFn; void *haloUser;     // host hook that exchanges halo `which` with the neighbours (XPIC iterations)
    struct SlabHost { int leave[2]; StatusFlags flags; } *slabHost;
    // NCCL inside the library (mpmgpu_slab_connect / mpmgpu_slab_step): halo and migrant exchange on `comm` in stream order, the
    // two-integer count handshake on `commSide` + sideStream while the last kernel of the step is still running
    ncclComm_t comm = NULL, commSide = NULL; int rank = 0, world = 1;
    cudaStream_t sideStream = NULL; cudaEvent_t sideEvent = NULL;
    int *dCounts = NULL; int *hCounts = NULL;       // [4] device / pinned host: to lower, to upper, from lower, from upper
    long long migratedOut = 0, migratedIn = 0;
    cudaEvent_t slabEvent; bool slabPending;                  // dynamic shared memory opt-in done for k_f2_strain_forces<SK, FEXT> on this device
};

static int fail(mpmgpu_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(ctx, MPMGPU_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

template <class T>
static cudaError_t dalloc(mpmgpu_ctx *ctx, T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) ctx->allocs.push_back((void *)*p);
    return e;
}

static inline int nblocks(long long n, int t) { return (int)((n + t - 1) / t); }

// (a grid computed from a particle count can be empty: a slab that holds no particles at the moment)
#define LAUNCH(kernel, grid, block, ...) do { if ((grid) > 0) { kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); ctx->launches++; } } while (0)

// node arrays: mass + 7 vectors in one pool, `count` entries each (count = fields x nodes in multimaterial mode)
static int alloc_node_arrays(mpmgpu_ctx *ctx, size_t count)
{
    const size_t pad = (count + 31) & ~(size_t)31;
    if (dalloc(ctx, &ctx->nodePool, pad * 22) != cudaSuccess) return MPMGPU_ECUDA;
    cudaMemset(ctx->nodePool, 0, pad * 22 * sizeof(double));
    double *q = ctx->nodePool;
    ctx->N.mass = q; q += pad;
    for (int c = 0; c < 3; c++) { ctx->N.pk[c] = q; q += pad; }
    for (int c = 0; c < 3; c++) { ctx->N.ftot[c] = q; q += pad; }
    for (int c = 0; c < 3; c++) { ctx->N.vk[c] = q; q += pad; }
    for (int c = 0; c < 3; c++) { ctx->N.pkc[c] = q; q += pad; }
    for (int c = 0; c < 3; c++) { ctx->N.vsp[c] = q; q += pad; }
    for (int c = 0; c < 3; c++) { ctx->N.vsn[c] = q; q += pad; }
    if (dalloc(ctx, &ctx->N.cnt, pad) != cudaSuccess) return MPMGPU_ECUDA;
    cudaMemset(ctx->N.cnt, 0, pad * sizeof(int));
    ctx->nodePad = pad;
    ctx->nvn = (int)count;
    return MPMGPU_OK;
}

// ------------------------------------------------------------------------------------------------
static void slab_disconnect(mpmgpu_ctx *ctx);
static void drop_step_graphs(mpmgpu_ctx *ctx);
extern "C" int mpmgpu_abi_version(void) { return MPMGPU_ABI_VERSION; }

extern "C" const char *mpmgpu_last_error(const mpmgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int mpmgpu_create(const mpmgpu_config *cfg, mpmgpu_ctx **out)
{
    mpmgpu_ctx *ctx = NULL;
    if (!cfg || !out) return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: null argument");
    *out = NULL;
    if (cfg->abi_version != MPMGPU_ABI_VERSION) return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: ABI version %d, library is %d", cfg->abi_version, MPMGPU_ABI_VERSION);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(NULL, MPMGPU_ENODEVICE, "mpmgpu_create: no CUDA device (libmpmgpu has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: device %d of %d", cfg->device, ndev);
    if (cfg->np != MPMGPU_THREED_MPM && cfg->np != MPMGPU_PLANE_STRAIN_MPM && cfg->np != MPMGPU_PLANE_STRESS_MPM)
        return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: analysis type %d not supported (10, 11, 12)", cfg->np);
    const bool is3D = cfg->np == MPMGPU_THREED_MPM;
    if (cfg->horiz < 3 || cfg->vert < 3 || (is3D && cfg->depth < 3)) return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: grid needs >=3 cells per axis incl. border");
    if (!cfg->xpts || !cfg->ypts || (is3D && !cfg->zpts)) return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: node coordinate arrays missing");
    if (cfg->shape != MPMGPU_POINT_GIMP && cfg->shape != MPMGPU_UNIFORM_GIMP && cfg->shape != MPMGPU_LINEAR_CPDI && cfg->shape != MPMGPU_QUADRATIC_CPDI &&
        cfg->shape != MPMGPU_BSPLINE_GIMP && cfg->shape != MPMGPU_BSPLINE && cfg->shape != MPMGPU_BSPLINE_CPDI)
        return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: shape function code %d not supported", cfg->shape);
    if (cfg->shape == MPMGPU_QUADRATIC_CPDI && is3D) return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: qCPDI is 2D only (as in the reference)");
    if (cfg->method != MPMGPU_USF && cfg->method != MPMGPU_USAVG && cfg->method != MPMGPU_USL)
        return fail(NULL, MPMGPU_EINVAL, "mpmgpu_create: MPM method %d not supported", cfg->method);

    ctx = new mpmgpu_ctx();
    ctx->cfg = *cfg;
    {   // CPDI: one call per touched node instead of one per (corner, node) pair.  3D lCPDI walks the window in registers
        // (shape.cuh::for_each_node_lcpdi3_hat) in every kernel; 2D keeps thread-local sums, which pay off in the value-only
        // kernels and spill in the gradient kernels (measured on B200: profiles/r2_experiments/README.md).
        // MPMGPU_CPDI_MERGE=0/1 forces all kernels, MPMGPU_CPDI_MERGE_VALUES=0 switches the value-only kernels back.
        const char *ge = getenv("MPMGPU_GRAPHS");
        ctx->useGraphs = ge ? atoi(ge) != 0 : true;
        const char *e = getenv("MPMGPU_CPDI_MERGE");
        ctx->cpdiMerge = e ? atoi(e) != 0 : is3D;
        const char *v = getenv("MPMGPU_CPDI_MERGE_VALUES");
        ctx->cpdiMergeValues = ctx->cpdiMerge || !(v && atoi(v) == 0);
    }
    ctx->dim = is3D ? 3 : 2;
    ctx->dMats = NULL; ctx->nmat = 0; ctx->dFlags = NULL;
    ctx->cap = 0; ctx->mstep = 0; ctx->mtime = 0.; ctx->launches = 0;
    ctx->archThickness = cfg->thickness > 0. ? cfg->thickness : 1.;      // the archive records of a 2D run carry the particle thickness
    ctx->uploaded = false; ctx->hasFext = false; ctx->hasBCs = false; ctx->hasReflectedBCs = false; ctx->profiling = false; ctx->globalIds = false; ctx->ownStreamSaved = false;
    ctx->particlePool = NULL; ctx->particleIntPool = NULL; ctx->nodePool = NULL;
    ctx->cpElemPool = NULL; ctx->cpXiPool = NULL; ctx->cpWgPool = NULL;
    ctx->nBCEntries = 0;
    memset(&ctx->P, 0, sizeof ctx->P); memset(&ctx->PR, 0, sizeof ctx->PR); memset(&ctx->R, 0, sizeof ctx->R); memset(&ctx->trac.TB, 0, sizeof ctx->trac.TB); memset(&ctx->flux.TB, 0, sizeof ctx->flux.TB);
    ctx->rigidPool = NULL; ctx->rigidIntPool = NULL; ctx->rigidCap = 0;
    memset(&ctx->N, 0, sizeof ctx->N); memset(&ctx->B, 0, sizeof ctx->B); memset(&ctx->C, 0, sizeof ctx->C); memset(&ctx->cp, 0, sizeof ctx->cp); memset(&ctx->T, 0, sizeof ctx->T); memset(&ctx->Q, 0, sizeof ctx->Q);
    memset(ctx->taskMs, 0, sizeof ctx->taskMs); memset(ctx->taskCalls, 0, sizeof ctx->taskCalls);
    memset(&ctx->hFlags, 0, sizeof ctx->hFlags);
    tiled_state_init(ctx->tiled);
    memset(ctx->f2Attr, 0, sizeof ctx->f2Attr);
    ctx->slabHost = NULL; ctx->slabPending = false; ctx->haloFn = NULL; ctx->haloUser = NULL;

    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { int rc = fail(NULL, MPMGPU_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e)); delete ctx; return rc; }
    cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);

    Grid &g = ctx->g;
    g.dim = ctx->dim; g.np = cfg->np;
    g.horiz = cfg->horiz; g.vert = cfg->vert; g.depth = is3D ? cfg->depth : 1;
    g.yplane = g.horiz + 1; g.zplane = (g.horiz + 1) * (g.vert + 1);
    g.nnodes = is3D ? (g.horiz + 1) * (g.vert + 1) * (g.depth + 1) : (g.horiz + 1) * (g.vert + 1);
    g.nelems = g.horiz * g.vert * g.depth;
    g.gx = cfg->gridx; g.gy = cfg->gridy; g.gz = is3D ? cfg->gridz : 0.;
    g.xmin = cfg->xpts[0]; g.ymin = cfg->ypts[0]; g.zmin = is3D ? cfg->zpts[0] : 0.;
    g.rcrit = cfg->cpdi_rcrit;
    g.lpUniform = 0; g.lpU[0] = g.lpU[1] = g.lpU[2] = 0.;
    g.inv2d[0] = 2.0 / g.gx; g.inv2d[1] = 2.0 / g.gy; g.inv2d[2] = is3D ? 2.0 / g.gz : 0.;
    double *dx = NULL, *dy = NULL, *dz = NULL;
    int rc = MPMGPU_OK;
    do {
        if (dalloc(ctx, &dx, g.horiz + 1) != cudaSuccess || dalloc(ctx, &dy, g.vert + 1) != cudaSuccess) { rc = MPMGPU_ECUDA; break; }
        cudaMemcpy(dx, cfg->xpts, (g.horiz + 1) * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(dy, cfg->ypts, (g.vert + 1) * sizeof(double), cudaMemcpyHostToDevice);
        if (is3D) {
            if (dalloc(ctx, &dz, g.depth + 1) != cudaSuccess) { rc = MPMGPU_ECUDA; break; }
            cudaMemcpy(dz, cfg->zpts, (g.depth + 1) * sizeof(double), cudaMemcpyHostToDevice);
        }
        g.xpts = dx; g.ypts = dy; g.zpts = dz;
        size_t nn = (size_t)g.nnodes;
        size_t nnPad = (nn + 31) & ~(size_t)31;
        if (alloc_node_arrays(ctx, nn) != MPMGPU_OK) { rc = MPMGPU_ECUDA; break; }
        if (dalloc(ctx, &ctx->dFlags, 1) != cudaSuccess) { rc = MPMGPU_ECUDA; break; }
        cudaMemset(ctx->dFlags, 0, sizeof(StatusFlags));
        if (dalloc(ctx, &ctx->dMats, MPM_MAX_MATERIALS) != cudaSuccess) { rc = MPMGPU_ECUDA; break; }
        if (is3D) {
            if (dalloc(ctx, &ctx->tiled.FN.V, nnPad) != cudaSuccess || dalloc(ctx, &ctx->tiled.FN.A, nnPad) != cudaSuccess ||
                dalloc(ctx, &ctx->tiled.FN.VS, nnPad) != cudaSuccess) { rc = MPMGPU_ECUDA; break; }
            cudaMemset(ctx->tiled.FN.V, 0, nnPad * sizeof(double4)); cudaMemset(ctx->tiled.FN.A, 0, nnPad * sizeof(double4));
            cudaMemset(ctx->tiled.FN.VS, 0, nnPad * sizeof(double4));
        }
    } while (0);
    if (rc != MPMGPU_OK) { fail(NULL, rc, "mpmgpu_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError())); mpmgpu_destroy(ctx); return rc; }

    StepParams &sp = ctx->sp;
    memset(&sp, 0, sizeof sp);
    sp.method = cfg->method; sp.skipPost = cfg->skip_post_extrapolation;
    sp.fractionUSF = cfg->fraction_usf > 0. ? cfg->fraction_usf : 0.5;
    sp.xpicOrder = cfg->xpic_order; sp.usingFMPM = cfg->using_fmpm;
    sp.gridAlpha = cfg->grid_damping; sp.particleAlpha = cfg->particle_damping;
    for (int c = 0; c < 3; c++) sp.grav[c] = cfg->gravity[c];
    sp.hasGravity = (sp.grav[0] != 0. || sp.grav[1] != 0. || sp.grav[2] != 0.);
    *out = ctx;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_destroy(mpmgpu_ctx *ctx)
{
    if (!ctx) return MPMGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    tiled_state_free(ctx->tiled);
    drop_step_graphs(ctx);
    for (void *p : ctx->allocs) cudaFree(p);
    if (ctx->slabHost) { cudaFreeHost(ctx->slabHost); cudaEventDestroy(ctx->slabEvent); }
    slab_disconnect(ctx);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->ownStreamSaved ? ctx->ownStream : ctx->stream);
    delete ctx;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_set_materials(mpmgpu_ctx *ctx, int nmat, const mpmgpu_material *mats)
{
    if (!ctx || !mats || nmat < 1 || nmat > MPM_MAX_MATERIALS) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_materials: need 1..%d materials", MPM_MAX_MATERIALS);
    ctx->hMats.resize(nmat);
    ctx->largeRotation = false;
    for (int i = 0; i < nmat; i++) {
        int k = mats[i].kind;
        if (k == MAT_NONE) {        // a contact law's place in the host's materials list: no particle may use it
            ctx->hMats[i].kind = k; ctx->hMats[i].nhist = 0;
            memset(ctx->hMats[i].p, 0, sizeof ctx->hMats[i].p);
            continue;
        }
        if (k == MAT_RIGIDCONTACT) {    // RigidMaterial in contact mode: its particles ride in the rigid set and claim no BC (direction bits 0)
            ctx->hMats[i].kind = k; ctx->hMats[i].nhist = 0;
            memcpy(ctx->hMats[i].p, mats[i].p, sizeof(double) * MPM_MAT_NPARAMS);
            ctx->hMats[i].p[8] = 0.; ctx->hMats[i].p[9] = 0.;
            if (!(ctx->hMats[i].p[0] > 0.)) ctx->hMats[i].p[0] = 1.;
            continue;
        }
        if (k != MAT_ISOTROPIC && k != MAT_RIGIDBC && k != MAT_NEOHOOKEAN && k != MAT_ISOPLASTICITY && k != MAT_MOONEY)
            return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_materials: material kind %d is not supported (IsotropicMat 1, Mooney 8, IsoPlasticity 9, rigid BC 11, Neohookean 28)", k);
        if (mats[i].n_history < 0 || mats[i].n_history > MPM_MAX_HISTORY) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_materials: %d history doubles (max %d)", mats[i].n_history, MPM_MAX_HISTORY);
        ctx->hMats[i].kind = k; ctx->hMats[i].nhist = mats[i].n_history;
        memcpy(ctx->hMats[i].p, mats[i].p, sizeof(double) * MPM_MAT_NPARAMS);
        if (k == MAT_ISOPLASTICITY) {
            const int law = (int)mats[i].p[16];
            if (law != 0 && law != HARD_LINEAR && law != HARD_NONLINEAR && law != HARD_JOHNSONCOOK && law != HARD_SCGL && law != HARD_NONLINEAR2)
                return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_materials: hardening law %d is not supported (1 Linear, 2 Nonlinear, 3 JohnsonCook, 4 SCGL, 6 Nonlinear2)", law);
        }
        if (mats[i].p[3] != 0. && k != MAT_NEOHOOKEAN && k != MAT_ISOPLASTICITY && k != MAT_MOONEY)
            return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_materials: material kind %d does not support artificial viscosity (MaterialBase::SupportsArtificialViscosity)", k);
        // MeshInfo::GetAverageCellSize for equal elements (MeshInfo.cpp:1517-1523): a grid constant the law needs
        if (k == MAT_MOONEY || (k == MAT_ISOPLASTICITY && mats[i].p[16] > 1.)) ctx->largeRotation = true;     // laws of the extended dispatch (Mooney, hardening laws other than Linear): per-task kernels, k_update_strains_lr
        if (mats[i].p[7] != 0.) {
            if (k != MAT_ISOTROPIC && k != MAT_ISOPLASTICITY)
                return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_materials: material kind %d has no large-rotation mode (Elastic::useLargeRotation: IsotropicMat, IsoPlasticity)", k);
            ctx->largeRotation = true;
        }
        ctx->hMats[i].p[6] = ctx->dim == 3 ? (ctx->cfg.gridx + ctx->cfg.gridy + ctx->cfg.gridz) / 3. : (ctx->cfg.gridx + ctx->cfg.gridy) / 2.;
    }
    ctx->nmat = nmat;
    drop_step_graphs(ctx);
    CK(cudaMemcpyAsync(ctx->dMats, ctx->hMats.data(), nmat * sizeof(Material), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

// <MultiMaterialMode>: mvf[field] per node + material contact.  After mpmgpu_set_materials, before mpmgpu_upload_particles.
extern "C" int mpmgpu_set_multimaterial(mpmgpu_ctx *ctx, const mpmgpu_multimaterial *mm)
{
    if (!ctx || !mm) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: null argument");
    if (ctx->nmat == 0) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_multimaterial: call mpmgpu_set_materials first");
    if (ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_multimaterial: call before mpmgpu_upload_particles");
    if (ctx->tiled.slab.on) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: not available in slab mode");
    if (ctx->cfg.kernel_path == 2) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: multimaterial mode runs on the per-task kernels (kernel_path 2 asked for the fused path)");
    if (mm->n_fields < 1 || mm->n_fields > MPM_MAX_FIELDS) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: %d material velocity fields (1..%d)", mm->n_fields, MPM_MAX_FIELDS);
    if (mm->normal_method < NORMALS_MAXG || mm->normal_method > NORMALS_SPECIFIED)
        return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: normal method %d is not supported (0 MAXG, 1 MAXV, 2 AVGG, 3 OWNG, 4 SN; the regression methods 5 and 6 are not built)", mm->normal_method);
    if (!mm->field_of_material || !mm->law_kind) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: field_of_material and law_kind are required");
    if (ctx->sp.xpicOrder > 1) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: XPIC/FMPM of order > 1 with material contact (MaterialXPICIncrementOnCVF) is not built");
    cudaSetDevice(ctx->cfg.device);
    const int nf = mm->n_fields;
    ctx->hFieldOfMat.assign(ctx->nmat, 0);
    for (int m = 0; m < ctx->nmat; m++) {
        const int f = mm->field_of_material[m];
        if (ctx->hMats[m].kind == MAT_RIGIDBC || ctx->hMats[m].kind == MAT_NONE) { ctx->hFieldOfMat[m] = 0; continue; }       // rigid-BC particles have no field
        if (f < 0 || f >= nf) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: material %d has velocity field %d of %d", m + 1, f, nf);
        ctx->hFieldOfMat[m] = f;
    }
    ContactParams &cp = ctx->cp;
    memset(&cp, 0, sizeof cp);
    cp.nf = nf; cp.normalMethod = mm->normal_method; cp.byDisplacements = mm->contact_by_displacements ? 1 : 0;
    cp.positionCutoff = mm->position_cutoff;
    for (int c = 0; c < 3; c++) cp.normal[c] = mm->contact_normal[c];
    cp.rigidBias = mm->rigid_gradient_bias > 0. ? mm->rigid_gradient_bias : 1.;
    for (int m = 0; m < ctx->nmat; m++) if (ctx->hMats[m].kind == MAT_RIGIDCONTACT) cp.rigidMask |= 1 << ctx->hFieldOfMat[m];
    for (int m = 0; m < ctx->nmat; m++)
        if (ctx->hMats[m].kind != MAT_RIGIDCONTACT && ctx->hMats[m].kind != MAT_RIGIDBC && ctx->hMats[m].kind != MAT_NONE && (cp.rigidMask >> ctx->hFieldOfMat[m] & 1))
            return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: material %d shares velocity field %d with a rigid contact material", m + 1, ctx->hFieldOfMat[m]);
    {   // MeshInfo::SetCartesian (MeshInfo.cpp:1466-1480): square / cubic cells by DbleEqual
        auto dbleEqual = [](double a, double b) { const double d = fabs(a - b); if (d <= 1.0e-16) return true; a = fabs(a); b = fabs(b); return d <= (b > a ? b : a) * 1.0e-7; };
        cp.cubic = dbleEqual(ctx->g.gx, ctx->g.gy) && (ctx->dim == 2 || dbleEqual(ctx->g.gx, ctx->g.gz)) ? 1 : 0;
    }
    for (int i = 0; i < nf; i++)
        for (int j = 0; j < nf; j++) {
            const int k = mm->law_kind[i * nf + j];
            if (i != j && (k < LAW_IGNORE || k > LAW_FRICTIONAL))
                return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_multimaterial: contact law %d between fields %d and %d is not supported (0 ignore, 1 stick, 2 frictionless, 3 Coulomb friction)", k, i, j);
            cp.lawKind[i * nf + j] = i == j ? LAW_IGNORE : k;
            cp.lawFriction[i * nf + j] = mm->law_friction ? mm->law_friction[i * nf + j] : 0.;
            cp.lawStatic[i * nf + j] = mm->law_static ? mm->law_static[i * nf + j] : -1.;
        }
    // node arrays for nf fields per node, field-major, plus the contact extrapolations
    if (alloc_node_arrays(ctx, (size_t)nf * ctx->g.nnodes) != MPMGPU_OK) return fail(ctx, MPMGPU_ECUDA, "mpmgpu_set_multimaterial: node arrays: %s", cudaGetErrorString(cudaGetLastError()));
    CK(dalloc(ctx, &ctx->contactPool, ctx->nodePad * 7));
    CK(cudaMemset(ctx->contactPool, 0, ctx->nodePad * 7 * sizeof(double)));
    CK(dalloc(ctx, &ctx->C.rcnt, ctx->nodePad));
    CK(cudaMemset(ctx->C.rcnt, 0, ctx->nodePad * sizeof(int)));
    {
        double *rf = NULL;
        CK(dalloc(ctx, &rf, ctx->nodePad * 3));
        CK(cudaMemset(rf, 0, ctx->nodePad * 3 * sizeof(double)));
        for (int c = 0; c < 3; c++) ctx->C.rforce[c] = rf + (size_t)c * ctx->nodePad;
    }
    {
        double *q = ctx->contactPool;
        ctx->C.cvol = q; q += ctx->nodePad;
        for (int c = 0; c < 3; c++) { ctx->C.cgrad[c] = q; q += ctx->nodePad; }
        for (int c = 0; c < 3; c++) { ctx->C.cdisp[c] = q; q += ctx->nodePad; }
    }
    CK(dalloc(ctx, &ctx->dFieldOfMat, (size_t)MPM_MAX_MATERIALS));
    CK(cudaMemcpy(ctx->dFieldOfMat, ctx->hFieldOfMat.data(), ctx->nmat * sizeof(int), cudaMemcpyHostToDevice));
    ctx->nf = nf;
    ctx->multimaterial = true;
    return MPMGPU_OK;
}

// <Thermal><Conduction/></Thermal>: heat conduction on the grid.  After mpmgpu_set_materials, before mpmgpu_upload_particles.
extern "C" int mpmgpu_set_conduction(mpmgpu_ctx *ctx, int nmat, const double *kcond)
{
    if (!ctx || !kcond) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_conduction: null argument");
    if (ctx->nmat == 0) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_conduction: call mpmgpu_set_materials first");
    if (nmat != ctx->nmat) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_conduction: %d conductivities for %d materials", nmat, ctx->nmat);
    if (ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_conduction: call before mpmgpu_upload_particles");
    if (ctx->tiled.slab.on) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_conduction: not available in slab mode");
    if (ctx->cfg.kernel_path == 2) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_conduction: transport tasks run on the per-task kernels (kernel_path 2 asked for the fused path)");
    // (a mechanical XPIC/FMPM order > 1 leaves the transport update as it is -- FLIP -- unless the transport task has its own XPIC
    // option, XPICExtrapolationTaskTO, which is the adapter's to refuse: UpdateParticlesTask.cpp:85-97)
    for (int i = 0; i < nmat; i++) {
        const Material &m = ctx->hMats[i];
        // thermal strains are in the device laws (materials.cuh) except in IsotropicMat's large-rotation form
        const bool cte = m.kind == MAT_ISOTROPIC && m.p[7] != 0. && (m.p[17] != 0. || m.p[18] != 0. || m.p[19] != 0.);
        if (cte) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_conduction: material %d: thermal expansion with <largeRotation> on IsotropicMat is not built", i + 1);
        if (m.kind != MAT_NONE && m.kind != MAT_RIGIDBC && !(m.p[1] > 0.)) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_conduction: material %d has no heat capacity", i + 1);
    }
    cudaSetDevice(ctx->cfg.device);
    const size_t nnPad = ((size_t)ctx->g.nnodes + 31) & ~(size_t)31;
    if (!ctx->transportPool) CK(dalloc(ctx, &ctx->transportPool, nnPad * 3));
    CK(cudaMemset(ctx->transportPool, 0, nnPad * 3 * sizeof(double)));
    ctx->T.gT = ctx->transportPool; ctx->T.gVCT = ctx->transportPool + nnPad; ctx->T.gQ = ctx->transportPool + 2 * nnPad;
    if (!ctx->dKcond) CK(dalloc(ctx, &ctx->dKcond, (size_t)MPM_MAX_MATERIALS));
    CK(cudaMemcpy(ctx->dKcond, kcond, nmat * sizeof(double), cudaMemcpyHostToDevice));
    ctx->T.kcond = ctx->dKcond;
    ctx->conduction = true;
    return MPMGPU_OK;
}

// Reaction forces of the velocity BCs (NodalVelBC::freaction): buffers for the BC list set now and for the rigid materials.
static int reaction_buffers(mpmgpu_ctx *ctx)
{
    if (!ctx->trackReactions) { ctx->B.reaction = NULL; ctx->R.reaction = NULL; ctx->tiled.FN.R.reaction = NULL; return MPMGPU_OK; }
    if ((size_t)ctx->nBCEntries > ctx->reactionCap) {
        CK(dalloc(ctx, &ctx->dReaction, 3 * (size_t)ctx->nBCEntries));
        ctx->reactionCap = (size_t)ctx->nBCEntries;
    }
    if (ctx->nBCEntries > 0) CK(cudaMemsetAsync(ctx->dReaction, 0, 3 * (size_t)ctx->nBCEntries * sizeof(double), ctx->stream));
    if (!ctx->dRigidReaction && ctx->nmat > 0) {
        CK(dalloc(ctx, &ctx->dRigidReaction, 3 * (size_t)ctx->nmat));
        CK(cudaMemsetAsync(ctx->dRigidReaction, 0, 3 * (size_t)ctx->nmat * sizeof(double), ctx->stream));
    }
    ctx->B.reaction = ctx->nBCEntries > 0 ? ctx->dReaction : NULL;
    ctx->R.reaction = ctx->dRigidReaction;
    ctx->tiled.FN.R.reaction = ctx->dRigidReaction;
    return MPMGPU_OK;
}

// every BC's freaction starts from zero in the grid-forces pass (NodalVelBC::ZeroVelocityBC, NodalVelBC.cpp:190); the rigid-particle
// BCs are made anew each step
static int reactions_zero(mpmgpu_ctx *ctx)
{
    if (!ctx->trackReactions) return MPMGPU_OK;
    if (ctx->B.reaction) CK(cudaMemsetAsync(ctx->B.reaction, 0, 3 * (size_t)ctx->nBCEntries * sizeof(double), ctx->stream));
    if (ctx->R.reaction) CK(cudaMemsetAsync(ctx->R.reaction, 0, 3 * (size_t)ctx->nmat * sizeof(double), ctx->stream));
    return MPMGPU_OK;
}

static int stage_buffer(mpmgpu_ctx *ctx, size_t ndoubles, double **out);

// Contact forces on the rigid contact materials: force[3*f..] = sum over the nodes where rigid material field f is active of the
// momentum its contacts gave the other materials since the last clearing (a rigid field's ftot, MatVelocityField::AddContactForce);
// 0 for the other fields.  The caller divides by (steps since the last clearing) x timestep.
extern "C" int mpmgpu_contact_forces(mpmgpu_ctx *ctx, int clear, double *force)
{
    if (!ctx || !force) return MPMGPU_EINVAL;
    if (!ctx->multimaterial) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_contact_forces: not in multimaterial mode");
    cudaSetDevice(ctx->cfg.device);
    double *tmp = NULL;
    int rc = stage_buffer(ctx, (size_t)3 * MPM_MAX_FIELDS, &tmp);
    if (rc) return rc;
    CK(cudaMemsetAsync(tmp, 0, 3 * MPM_MAX_FIELDS * sizeof(double), ctx->stream));
    if (ctx->cp.rigidMask)
        LAUNCH(k_contact_force_sum, nblocks((long long)ctx->g.nnodes * ctx->nf, 256), 256, ctx->g.nnodes, ctx->nf, (unsigned)ctx->cp.rigidMask, ctx->C, clear ? 1 : 0, tmp);
    CK(cudaMemcpyAsync(force, tmp, (size_t)3 * ctx->nf * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_track_reactions(mpmgpu_ctx *ctx, int on)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (on && ctx->tiled.slab.on) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_track_reactions: not available in slab mode (halo nodes would count twice)");
    if (on && ctx->nmat <= 0) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_track_reactions: call after mpmgpu_set_materials");
    cudaSetDevice(ctx->cfg.device);
    ctx->trackReactions = on != 0;
    return reaction_buffers(ctx);
}

extern "C" int mpmgpu_download_reactions(mpmgpu_ctx *ctx, int n, double *bc_reaction, double *rigid_reaction)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (!ctx->trackReactions) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_download_reactions: mpmgpu_track_reactions was not called");
    if (bc_reaction && n != ctx->nBCEntries) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_download_reactions: n=%d but %d BCs are set", n, ctx->nBCEntries);
    cudaSetDevice(ctx->cfg.device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (bc_reaction && n > 0) {
        std::vector<double> dev(3 * (size_t)n);
        CK(cudaMemcpy(dev.data(), ctx->dReaction, dev.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int e = 0; e < n; e++) {
            const int i = ctx->bcOrder[e];
            for (int d = 0; d < 3; d++) bc_reaction[3 * (size_t)i + d] = dev[3 * (size_t)e + d];
        }
    }
    if (rigid_reaction) {
        if (ctx->dRigidReaction) CK(cudaMemcpy(rigid_reaction, ctx->dRigidReaction, 3 * (size_t)ctx->nmat * sizeof(double), cudaMemcpyDeviceToHost));
        else for (int i = 0; i < 3 * ctx->nmat; i++) rigid_reaction[i] = 0.;
    }
    return MPMGPU_OK;
}

// <EnergyCoupling/>: ConductionTask::adiabatic.  Before mpmgpu_upload_particles.
extern "C" int mpmgpu_set_energy_coupling(mpmgpu_ctx *ctx, int adiabatic)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_energy_coupling: call before mpmgpu_upload_particles");
    if (adiabatic && ctx->tiled.slab.on) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_energy_coupling: not available in slab mode");
    if (adiabatic && ctx->cfg.kernel_path == 2) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_energy_coupling: adiabatic coupling runs on the per-task kernels (kernel_path 2 asked for the fused path)");
    ctx->adiabatic = adiabatic != 0;
    ctx->sp.adiabatic = ctx->adiabatic ? 1 : 0;
    return MPMGPU_OK;
}

// Nodal temperature BCs in the host's list order (firstTempBC ...): node[i] 1-based, value[i] = BCValue at this step's time,
// active[i] = GetNodeNum(time) != 0.  Call again (same or another list) whenever values change.
extern "C" int mpmgpu_set_temperature_bcs(mpmgpu_ctx *ctx, int n, const int *node, const double *value, const int *active)
{
    if (!ctx || n < 0 || (n > 0 && (!node || !value))) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_temperature_bcs: bad arguments");
    if (!ctx->conduction) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_temperature_bcs: call mpmgpu_set_conduction first");
    cudaSetDevice(ctx->cfg.device);
    for (int i = 0; i < n; i++)
        if (node[i] < 1 || node[i] > ctx->g.nnodes) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_temperature_bcs: BC %d on node %d of %d", i, node[i], ctx->g.nnodes);
    {   // nodes whose temperature a grid BC holds (active or not: NodalTempBC's constructor sets the node's TEMP_DIRECTION bit): rigid
        // particles make no BC there (ProjectRigidBCsTask.cpp:202)
        std::vector<unsigned char> fixed((size_t)ctx->g.nnodes, 0);
        for (int i = 0; i < n; i++) fixed[node[i] - 1] = 1;
        if (fixed != ctx->hFixedTemp) {
            if (!ctx->dFixedTemp) CK(dalloc(ctx, &ctx->dFixedTemp, (size_t)ctx->g.nnodes));
            CK(cudaMemcpy(ctx->dFixedTemp, fixed.data(), fixed.size(), cudaMemcpyHostToDevice));
            ctx->hFixedTemp.swap(fixed);
        }
        ctx->R.fixedT = ctx->dFixedTemp;
    }
    if (n == 0) { ctx->Q.nUnique = 0; ctx->tbcEntries = 0; return MPMGPU_OK; }
    // group by node, list order kept inside a node (the reference walks its list: zero all, then add all)
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return node[a] < node[b]; });
    std::vector<int> un, st, ac(n); std::vector<double> va(n);
    for (int e = 0; e < n; e++) {
        const int i = order[e];
        if (e == 0 || node[i] != node[order[e - 1]]) { un.push_back(node[i] - 1); st.push_back(e); }
        va[e] = value[i]; ac[e] = active ? active[i] : 1;
    }
    st.push_back(n);
    if (n > ctx->tbcCap) {
        CK(dalloc(ctx, &ctx->dTbcNode, (size_t)n)); CK(dalloc(ctx, &ctx->dTbcStart, (size_t)n + 1)); CK(dalloc(ctx, &ctx->dTbcActive, (size_t)n));
        CK(dalloc(ctx, &ctx->dTbcValue, (size_t)n)); CK(dalloc(ctx, &ctx->dTbcSaved, (size_t)n));
        ctx->tbcCap = n;
    }
    CK(cudaMemcpyAsync(ctx->dTbcNode, un.data(), un.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dTbcStart, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dTbcActive, ac.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->dTbcValue, va.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->Q.nUnique = (int)un.size(); ctx->Q.node = ctx->dTbcNode; ctx->Q.start = ctx->dTbcStart; ctx->Q.value = ctx->dTbcValue;
    ctx->Q.active = ctx->dTbcActive; ctx->Q.saved = ctx->dTbcSaved;
    ctx->tbcEntries = n;
    return MPMGPU_OK;
}

// number of double arrays per particle in the pool
#define NPD (3 + 3 + 1 + 3 + 3 + 9 + 6 + 1 + 6 + 6 + MPM_MAX_HISTORY + 3 + 3)
#define NPI 5

static void bind_particles(Particles &P, double *pool, int *ipool, size_t capPad);

static int alloc_particles(mpmgpu_ctx *ctx, size_t cap)
{
    if (ctx->cap >= cap) return MPMGPU_OK;
    if (ctx->cap != 0) return fail(ctx, MPMGPU_EINVAL, "particle capacity %zu exceeded (%zu); set max_particles", ctx->cap, cap);
    size_t capPad = (cap + 31) & ~(size_t)31;
    CK(dalloc(ctx, &ctx->particlePool, capPad * NPD));
    CK(dalloc(ctx, &ctx->particleIntPool, capPad * NPI));
    CK(cudaMemsetAsync(ctx->particlePool, 0, capPad * NPD * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->particleIntPool, 0, capPad * NPI * sizeof(int), ctx->stream));
    bind_particles(ctx->P, ctx->particlePool, ctx->particleIntPool, capPad);
    ctx->cap = capPad;
    if (ctx->cfg.shape == MPMGPU_LINEAR_CPDI || ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI || ctx->cfg.shape == MPMGPU_BSPLINE_CPDI) {
        const int nc = ctx->dim == 3 ? 8 : (ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI ? 9 : 4);
        CK(dalloc(ctx, &ctx->cpElemPool, capPad * nc));
        CK(dalloc(ctx, &ctx->cpXiPool, capPad * (nc * 3 + 12)));         // + the 12 rows of Particles::cpDom
        CK(dalloc(ctx, &ctx->cpWgPool, capPad * nc * 3));
        CK(cudaMemsetAsync(ctx->cpElemPool, 0, capPad * nc * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->cpXiPool, 0, capPad * (nc * 3 + 12) * sizeof(double), ctx->stream));
        ctx->cpDomOffset = (size_t)nc * 3;
        CK(cudaMemsetAsync(ctx->cpWgPool, 0, capPad * nc * 3 * sizeof(double), ctx->stream));
    }
    ctx->P.cpElem = ctx->cpElemPool; ctx->P.cpXi = ctx->cpXiPool; ctx->P.cpWg = ctx->cpWgPool; ctx->P.cpStride = capPad;
    ctx->P.cpDom = ctx->cpXiPool ? ctx->cpXiPool + ctx->cpDomOffset * capPad : NULL;
    return MPMGPU_OK;
}

static void bind_particles(Particles &P, double *pool, int *ipool, size_t capPad)
{
    double *q = pool;
    auto take = [&]() { double *r = q; q += capPad; return r; };
    for (int c = 0; c < 3; c++) P.pos[c] = take();
    for (int c = 0; c < 3; c++) P.vel[c] = take();
    P.mp = take();
    for (int c = 0; c < 3; c++) P.lp[c] = take();
    for (int c = 0; c < 3; c++) P.ncpos[c] = take();
    for (int c = 0; c < 9; c++) P.F[c] = take();
    for (int c = 0; c < 6; c++) P.sp[c] = take();
    P.pressure = take();
    for (int c = 0; c < 6; c++) P.eplast[c] = take();
    P.work = take(); P.res = take(); P.heat = take(); P.entropy = take(); P.plast = take(); P.prevT = take();
    for (int c = 0; c < MPM_MAX_HISTORY; c++) P.hist[c] = take();
    for (int c = 0; c < 3; c++) P.pfext[c] = take();
    for (int c = 0; c < 3; c++) P.acc[c] = take();
    int *qi = ipool;
    P.elem = qi; qi += capPad; P.mat = qi; qi += capPad; P.cross = qi; qi += capPad; P.orig = qi; qi += capPad; P.key = qi;
}

// copy rows [off, off+cnt) of a host [ncomp][n] array to device component arrays (zero-fill when host is NULL)
static int up_field(mpmgpu_ctx *ctx, double *const *dev, const double *host, int ncomp, int n, int off, int cnt)
{
    for (int c = 0; c < ncomp; c++) {
        if (host) CK(cudaMemcpyAsync(dev[c], host + (size_t)c * n + off, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        else CK(cudaMemsetAsync(dev[c], 0, (size_t)cnt * sizeof(double), ctx->stream));
    }
    return MPMGPU_OK;
}

__global__ void k_epwrot_to_F(int cnt, int n, int dim, const double *ep, const double *wrot, Particles P)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cnt) return;
    // MatPoint3D::GetDeformationGradient (MatPoint3D.cpp:363-376), 2D: MatPoint2D.cpp:369-377
    double exx = ep ? ep[p] : 0., eyy = ep ? ep[n + p] : 0., ezz = ep ? ep[2 * n + p] : 0.;
    double eyz = ep ? ep[3 * n + p] : 0., exz = ep ? ep[4 * n + p] : 0., exy = ep ? ep[5 * n + p] : 0.;
    double wxy = wrot ? wrot[p] : 0., wxz = wrot ? wrot[n + p] : 0., wyz = wrot ? wrot[2 * n + p] : 0.;
    P.F[0][p] = 1. + exx; P.F[4][p] = 1. + eyy; P.F[8][p] = 1. + ezz;
    P.F[1][p] = 0.5 * (exy - wxy); P.F[3][p] = 0.5 * (exy + wxy);
    if (dim == 3) {
        P.F[2][p] = 0.5 * (exz - wxz); P.F[6][p] = 0.5 * (exz + wxz);
        P.F[5][p] = 0.5 * (eyz - wyz); P.F[7][p] = 0.5 * (eyz + wyz);
    } else {
        P.F[2][p] = 0.; P.F[6][p] = 0.; P.F[5][p] = 0.; P.F[7][p] = 0.;
    }
}

__global__ void k_F_to_epwrot(int cnt, int n, int dim, Particles P, const int *slot, double *ep, double *wrot)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cnt) return;
    int o = slot[p];        // caller's index
    // MatPoint3D::SetDeformationGradientMatrix (MatPoint3D.cpp:320-336)
    double F0 = P.F[0][p], F1 = P.F[1][p], F2 = P.F[2][p], F3 = P.F[3][p], F4 = P.F[4][p], F5 = P.F[5][p], F6 = P.F[6][p], F7 = P.F[7][p], F8 = P.F[8][p];
    ep[o] = F0 - 1.; ep[n + o] = F4 - 1.; ep[2 * n + o] = F8 - 1.;
    ep[5 * n + o] = F3 + F1; wrot[o] = F3 - F1;
    if (dim == 3) {
        ep[4 * n + o] = F6 + F2; ep[3 * n + o] = F7 + F5;
        wrot[n + o] = F6 - F2; wrot[2 * n + o] = F7 - F5;
    } else {
        ep[4 * n + o] = 0.; ep[3 * n + o] = 0.; wrot[n + o] = 0.; wrot[2 * n + o] = 0.;
    }
}

__global__ void k_iota(int n, int *a, int first = 0) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = first + i; }
__global__ void k_fill(int n, double *a, double v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = v; }
__global__ void k_dec(int n, int *a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] -= 1; }

// scatter device (internal order) -> staging buffer in the caller's order
__global__ void k_unpermute(int n, const double *src, const int *slot, double *dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[slot[p]] = src[p];
}
__global__ void k_unpermute_int(int n, const int *src, const int *slot, int *dst, int add)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[slot[p]] = src[p] + add;
}

// rows [off, off+cnt) of the caller's arrays -> one particle set
// one staging buffer per context, grown on demand: no cudaMalloc/cudaFree (and their device-wide syncs) per transfer
static int stage_buffer(mpmgpu_ctx *ctx, size_t ndoubles, double **out)
{
    if (ctx->stageLen < ndoubles) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->stage) { cudaFree(ctx->stage); for (auto &a : ctx->allocs) if (a == (void *)ctx->stage) a = NULL; ctx->stage = NULL; ctx->stageLen = 0; }
        CK(dalloc(ctx, &ctx->stage, ndoubles));
        ctx->stageLen = ndoubles;
    }
    *out = ctx->stage;
    return MPMGPU_OK;
}

static int upload_range(mpmgpu_ctx *ctx, Particles &P, const mpmgpu_particles *h, int off, int cnt)
{
    const int n = h->n;
    int rc;
    const int T = 256;
    if ((rc = up_field(ctx, P.pos, h->pos, 3, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.vel, h->vel, 3, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, &P.mp, h->mp, 1, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.lp, h->lp, 3, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.sp, h->sp, 6, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, &P.pressure, h->pressure, 1, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.eplast, h->eplast, 6, n, off, cnt))) return rc;
    if (h->energies) {
        double *const e5[6] = {P.work, P.res, P.heat, P.entropy, P.plast, P.prevT};
        if ((rc = up_field(ctx, e5, h->energies, 6, n, off, cnt))) return rc;
    } else {
        double *const e5[5] = {P.work, P.res, P.heat, P.entropy, P.plast};
        if ((rc = up_field(ctx, e5, NULL, 5, n, off, cnt))) return rc;
        LAUNCH(k_fill, nblocks(cnt, T), T, cnt, P.prevT, 1.);
    }
    if ((rc = up_field(ctx, P.hist, h->history, MPM_MAX_HISTORY, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.pfext, h->pfext, 3, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.acc, NULL, 3, n, off, cnt))) return rc;
    if ((rc = up_field(ctx, P.ncpos, NULL, 3, n, off, cnt))) return rc;
    CK(cudaMemcpyAsync(P.elem, h->in_elem + off, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (h->matnum) {
        CK(cudaMemcpyAsync(P.mat, h->matnum + off, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_dec, nblocks(cnt, T), T, cnt, P.mat);
    } else CK(cudaMemsetAsync(P.mat, 0, (size_t)cnt * sizeof(int), ctx->stream));
    if (h->crossings) CK(cudaMemcpyAsync(P.cross, h->crossings + off, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    else CK(cudaMemsetAsync(P.cross, 0, (size_t)cnt * sizeof(int), ctx->stream));
    if (h->ids) CK(cudaMemcpyAsync(P.orig, h->ids + off, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    else LAUNCH(k_iota, nblocks(cnt, T), T, cnt, P.orig, off);
    // strain + rotation -> deformation gradient (staged through a temporary device buffer)
    {
        double *tmp = NULL;
        int rcs = stage_buffer(ctx, (size_t)cnt * 9, &tmp);
        if (rcs) return rcs;
        const double *dep = NULL, *dw = NULL;
        if (h->ep) {
            for (int c = 0; c < 6; c++) CK(cudaMemcpyAsync(tmp + (size_t)c * cnt, h->ep + (size_t)c * n + off, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            dep = tmp;
        }
        if (h->wrot) {
            for (int c = 0; c < 3; c++) CK(cudaMemcpyAsync(tmp + (size_t)(6 + c) * cnt, h->wrot + (size_t)c * n + off, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            dw = tmp + (size_t)6 * cnt;
        }
        LAUNCH(k_epwrot_to_F, nblocks(cnt, T), T, cnt, cnt, ctx->dim, dep, dw, P);
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MPMGPU_OK;
}

static int alloc_rigid(mpmgpu_ctx *ctx, size_t cap)
{
    if (ctx->rigidCap < cap) {
        if (ctx->rigidCap != 0) return fail(ctx, MPMGPU_EINVAL, "rigid particle count grew from %zu to %zu", ctx->rigidCap, cap);
        size_t capPad = (cap + 31) & ~(size_t)31;
        CK(dalloc(ctx, &ctx->rigidPool, capPad * NPD));
        CK(dalloc(ctx, &ctx->rigidIntPool, capPad * NPI));
        CK(cudaMemsetAsync(ctx->rigidPool, 0, capPad * NPD * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(ctx->rigidIntPool, 0, capPad * NPI * sizeof(int), ctx->stream));
        bind_particles(ctx->PR, ctx->rigidPool, ctx->rigidIntPool, capPad);
        ctx->rigidCap = capPad;
        const size_t nn = (size_t)ctx->g.nnodes;
        for (int d = 0; d < 3; d++) {
            CK(dalloc(ctx, &ctx->R.owner[d], nn));
            CK(cudaMemsetAsync(ctx->R.owner[d], 0x7f, nn * sizeof(int), ctx->stream));        // RIGID_NONE until the projection task claims dofs
            ctx->R.vel[d] = ctx->PR.vel[d];
        }
        unsigned char *fb = NULL;
        CK(dalloc(ctx, &fb, nn));
        if (ctx->hFixedBits.size() == nn) CK(cudaMemcpyAsync(fb, ctx->hFixedBits.data(), nn, cudaMemcpyHostToDevice, ctx->stream));
        else CK(cudaMemsetAsync(fb, 0, nn, ctx->stream));
        ctx->R.fixedBits = fb;
        if (ctx->cfg.shape == MPMGPU_LINEAR_CPDI || ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI || ctx->cfg.shape == MPMGPU_BSPLINE_CPDI) {     // CPDI domains of the rigid particles
            const int nc = ctx->dim == 3 ? 8 : (ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI ? 9 : 4);
            int *ce; double *cx, *cw;
            CK(dalloc(ctx, &ce, capPad * nc)); CK(dalloc(ctx, &cx, capPad * (nc * 3 + 12))); CK(dalloc(ctx, &cw, capPad * nc * 3));
            CK(cudaMemsetAsync(ce, 0, capPad * nc * sizeof(int), ctx->stream));
            CK(cudaMemsetAsync(cx, 0, capPad * (nc * 3 + 12) * sizeof(double), ctx->stream));
            CK(cudaMemsetAsync(cw, 0, capPad * nc * 3 * sizeof(double), ctx->stream));
            ctx->PR.cpElem = ce; ctx->PR.cpXi = cx; ctx->PR.cpWg = cw; ctx->PR.cpStride = capPad;
            ctx->PR.cpDom = cx + (size_t)nc * 3 * capPad;
        }
    }
    return MPMGPU_OK;
}

__global__ void k_validate_upload(int cnt, int off, int rigidPart, Particles P, const Material *mats, int nmat, int nelems, UploadCheck *out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cnt) return;
    const int m = P.mat[p], e = P.elem[p];          // P.mat is 0-based on the device
    if (m < 0 || m >= nmat || mats[m].kind == MAT_NONE) { atomicMin(&out->badMat, off + p); return; }
    if (e < 1 || e > nelems) atomicMin(&out->badElem, off + p);
    const bool rigid = mats[m].kind == MAT_RIGIDBC || mats[m].kind == MAT_RIGIDCONTACT;
    if (rigidPart && !rigid) atomicMin(&out->notRigid, off + p);
    if (!rigidPart && rigid) atomicMin(&out->rigidEarly, off + p);
    if (!rigidPart) {
        bool large = false, varies = false;
#pragma unroll
        for (int c = 0; c < 3; c++) { const double l = P.lp[c][p]; large |= !(l <= 1.0); varies |= l != P.lp[c][0]; }
        if (large) out->lpLarge = 1;
        if (varies) out->lpVaries = 1;
    }
}

extern "C" int mpmgpu_upload_particles(mpmgpu_ctx *ctx, const mpmgpu_particles *h)
{
    if (!ctx || !h) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: null argument");
    if (ctx->nmat == 0) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_upload_particles: call mpmgpu_set_materials first");
    const int n = h->n;
    // (a slab of a multi-GPU run may start without particles: a body can enter it later)
    const bool emptySlab = n == 0 && ctx->tiled.slab.on;
    if ((n < 1 && !emptySlab) || h->n_nonrigid < 0 || h->n_nonrigid > n) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: bad counts n=%d nonrigid=%d", n, h->n_nonrigid);
    if (!emptySlab && (!h->pos || !h->mp || !h->in_elem || !h->lp)) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: pos, mp, lp and in_elem are required");
    cudaSetDevice(ctx->cfg.device);
    size_t want = ctx->cfg.max_particles > n ? (size_t)ctx->cfg.max_particles : (size_t)n;
    int rc = alloc_particles(ctx, want);
    if (rc) return rc;
    const int nNR = h->n_nonrigid, nR = n - nNR;
    if (nR > 0 && (rc = alloc_rigid(ctx, (size_t)nR))) return rc;
    ctx->P.n = nNR; ctx->P.nNR = nNR;
    ctx->PR.n = nR; ctx->PR.nNR = 0;
    ctx->hasFext = h->pfext != NULL;
    if (h->ids || emptySlab) ctx->globalIds = true;
    if (nNR > 0 && (rc = upload_range(ctx, ctx->P, h, 0, nNR))) return rc;
    if (nR > 0 && (rc = upload_range(ctx, ctx->PR, h, nNR, nR))) return rc;
    // validate what arrived, on the device (a host pass over 8M particles costs more than the copy): materials and elements in
    // range, rigid-BC particles after the non-rigid ones (the reference reorders them to the end, NairnMPM.cpp:1121-1150), and
    // the two facts the fused path depends on (particles no larger than a cell, one particle size)
    UploadCheck chk;
    {
        UploadCheck init = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff, 0, 0};
        if (!ctx->dUploadCheck) CK(dalloc(ctx, &ctx->dUploadCheck, 1));
        CK(cudaMemcpyAsync(ctx->dUploadCheck, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
        if (nNR) LAUNCH(k_validate_upload, nblocks(nNR, 256), 256, nNR, 0, 0, ctx->P, ctx->dMats, ctx->nmat, ctx->g.nelems, ctx->dUploadCheck);
        if (nR) LAUNCH(k_validate_upload, nblocks(nR, 256), 256, nR, nNR, 1, ctx->PR, ctx->dMats, ctx->nmat, ctx->g.nelems, ctx->dUploadCheck);
        CK(cudaMemcpyAsync(&chk, ctx->dUploadCheck, sizeof chk, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (chk.badMat != 0x7fffffff) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: particle %d has material %d of %d", chk.badMat, h->matnum ? h->matnum[chk.badMat] : 1, ctx->nmat);
    if (chk.badElem != 0x7fffffff) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: particle %d in element %d of %d", chk.badElem, h->in_elem[chk.badElem], ctx->g.nelems);
    if (chk.notRigid != 0x7fffffff) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: particle %d (after n_nonrigid) is not a rigid-BC material", chk.notRigid);
    if (chk.rigidEarly != 0x7fffffff)
        return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: rigid particle %d before n_nonrigid=%d (the reference reorders them to the end, NairnMPM.cpp:1121-1150)", chk.rigidEarly, nNR);
    // (rigid CONTACT particles ride in the rigid set too but make no velocity BCs)
    int nRigidBC = 0;
    for (int p = nNR; p < n; p++) if (ctx->hMats[(h->matnum ? h->matnum[p] : 1) - 1].kind == MAT_RIGIDBC) nRigidBC++;
    ctx->R.on = nRigidBC > 0 ? 1 : 0;
    ctx->R.mirrored = 0;
    for (int p = nNR; p < n; p++) if (ctx->hMats[(h->matnum ? h->matnum[p] : 1) - 1].p[9] != 0.) ctx->R.mirrored = 1;
    ctx->R.mat = ctx->PR.mat; ctx->R.mats = ctx->dMats;
    ctx->R.stride[0] = 1; ctx->R.stride[1] = ctx->g.yplane; ctx->R.stride[2] = ctx->g.zplane; ctx->R.nnodes = ctx->g.nnodes;
    ctx->R.reaction = ctx->trackReactions ? ctx->dRigidReaction : NULL;
    ctx->tiled.FN.R = ctx->R;
    if (!ctx->globalIds && !ctx->archOriginFromCaller) {
        // "original position" column of the archive records (ArchiveData.cpp:868-872): the positions at upload unless the caller
        // hands over others with mpmgpu_set_archive_origin
        if (ctx->archOrigin && ctx->archOriginLen < (size_t)3 * n) ctx->archOrigin = NULL;
        if (!ctx->archOrigin) { CK(dalloc(ctx, &ctx->archOrigin, (size_t)3 * n)); ctx->archOriginLen = (size_t)3 * n; }
        for (int c = 0; c < 3; c++) {
            if (nNR) CK(cudaMemcpyAsync(ctx->archOrigin + (size_t)c * n, ctx->P.pos[c], (size_t)nNR * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            if (nR) CK(cudaMemcpyAsync(ctx->archOrigin + (size_t)c * n + nNR, ctx->PR.pos[c], (size_t)nR * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    ctx->thermal = ctx->conduction || ctx->adiabatic || h->temperature != NULL;
    if (ctx->thermal) {
        if (ctx->globalIds) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: particle temperatures / conduction with caller-global particle ids (slab mode) are not built");
        if (!ctx->conduction)
            for (int i = 0; i < ctx->nmat; i++) {
                const Material &m = ctx->hMats[i];
                if (m.kind == MAT_ISOTROPIC && m.p[7] != 0. && (m.p[17] != 0. || m.p[18] != 0. || m.p[19] != 0.))
                    return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: material %d: thermal expansion with <largeRotation> on IsotropicMat is not built", i + 1);
            }
        if (!ctx->tempPool) { CK(dalloc(ctx, &ctx->tempPool, ctx->cap * 6)); CK(cudaMemsetAsync(ctx->tempPool, 0, ctx->cap * 6 * sizeof(double), ctx->stream)); }
        ctx->P.temp = ctx->tempPool;
        for (int c = 0; c < 3; c++) ctx->P.tgrad[c] = ctx->tempPool + (size_t)(c + 1) * ctx->cap;
        ctx->P.dTr = ctx->tempPool + (size_t)4 * ctx->cap;
        ctx->P.dTad = ctx->adiabatic ? ctx->tempPool + (size_t)5 * ctx->cap : NULL;
        if (ctx->adiabatic && nNR) CK(cudaMemsetAsync(ctx->P.dTad, 0, (size_t)nNR * sizeof(double), ctx->stream));
        ctx->rigidTemp = false;
        if (ctx->conduction && nR)
            for (int i = 0; i < ctx->nmat; i++) if (ctx->hMats[i].kind == MAT_RIGIDBC && ctx->hMats[i].p[10] != 0.) ctx->rigidTemp = true;
        if (ctx->rigidTemp) {           // rigid particles that hold the nodes they touch at their own temperature
            if (!h->temperature) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: a rigid material sets the temperature: mpmgpu_particles.temperature is needed");
            const size_t nn = (size_t)ctx->g.nnodes;
            if (!ctx->rigidTempPool) {
                CK(dalloc(ctx, &ctx->rigidTempPool, ctx->rigidCap));
                CK(dalloc(ctx, &ctx->R.ownerT, nn)); CK(dalloc(ctx, &ctx->R.savedT, nn));
                CK(cudaMemsetAsync(ctx->R.ownerT, 0x7f, nn * sizeof(int), ctx->stream));
            }
            ctx->PR.temp = ctx->rigidTempPool;
            if ((rc = up_field(ctx, &ctx->PR.temp, h->temperature, 1, n, nNR, nR))) return rc;
            ctx->R.ptemp = ctx->PR.temp;
            ctx->R.fixedT = ctx->dFixedTemp;
        } else { ctx->R.ownerT = NULL; ctx->R.ptemp = NULL; ctx->R.savedT = NULL; }
        if (nNR) {
            // pTemperature; without the array every particle starts at the temperature of its last strain update (energies[5])
            if (h->temperature) { if ((rc = up_field(ctx, &ctx->P.temp, h->temperature, 1, n, 0, nNR))) return rc; }
            else CK(cudaMemcpyAsync(ctx->P.temp, ctx->P.prevT, (size_t)nNR * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    if (!ctx->multimaterial)
        for (int i = 0; i < ctx->nmat; i++)
            if (ctx->hMats[i].kind == MAT_RIGIDCONTACT) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: material %d is a rigid contact material: call mpmgpu_set_multimaterial first", i + 1);
    if (ctx->multimaterial) {
        if (ctx->globalIds) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: multimaterial mode with caller-global particle ids (slab mode) is not built");
        if (ctx->R.mirrored) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_upload_particles: mirrored rigid BCs in multimaterial mode are not built");
        if (!ctx->foffPool) CK(dalloc(ctx, &ctx->foffPool, ctx->cap));
        if (nNR) LAUNCH(k_set_field_offsets, nblocks(nNR, 256), 256, nNR, ctx->P.mat, ctx->dFieldOfMat, ctx->g.nnodes, ctx->foffPool);
        ctx->P.foff = ctx->foffPool;
        if (nR) {       // rigid contact particles extrapolate to their material's field too
            if (!ctx->foffRigidPool) CK(dalloc(ctx, &ctx->foffRigidPool, ctx->rigidCap));
            LAUNCH(k_set_field_offsets, nblocks(nR, 256), 256, nR, ctx->PR.mat, ctx->dFieldOfMat, ctx->g.nnodes, ctx->foffRigidPool);
            ctx->PR.foff = ctx->foffRigidPool;
        }
    }
    CK(cudaMemsetAsync(ctx->dFlags, 0, sizeof(StatusFlags), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->uploaded = true;
    tiled_on_upload(ctx->tiled);
    {   // the fused path needs 3D uGIMP, particles no larger than a cell, FLIP/PIC, no rigid particles
        bool ok = ctx->dim == 3 && ctx->cfg.shape == MPMGPU_UNIFORM_GIMP && (nNR > 0 || emptySlab);   // (XPIC order is checked per step)
        if (chk.lpLarge) ok = false;
        ctx->g.lpUniform = (chk.lpVaries || emptySlab) ? 0 : 1;        // an empty slab learns the sizes from the particles that migrate in
        for (int c = 0; c < 3; c++) {
            ctx->g.lpU[c] = emptySlab ? 0.5 : h->lp[(size_t)c * n];
            ctx->g.lpInvSize[c] = 1. / (4. * ctx->g.lpU[c]);
            ctx->g.lpInv2[c] = 1. / (2. * ctx->g.lpU[c]);
        }
        ctx->tiled.stateKind = SK_ELASTIC;
        for (int i = 0; i < ctx->nmat; i++) if (ctx->hMats[i].kind != MAT_ISOTROPIC && ctx->hMats[i].kind != MAT_RIGIDBC) ctx->tiled.stateKind = SK_FULL;
        if (ctx->largeRotation) ok = false;     // large-rotation hypoelastic laws and Mooney live in the per-task strain kernel
        if (ctx->hasReflectedBCs) ok = false;   // symmetry-plane BCs read the momentum of the node across the plane: per-task kernels
        if (ctx->R.mirrored) ok = false;        // a mirrored rigid BC reads a neighbour node's momentum between the node updates: per-task kernels
        if (ctx->multimaterial) ok = false;     // material velocity fields + contact: per-task kernels
        if (ctx->conduction || ctx->thermal) ok = false;        // transport tasks, temperature changes handed to the laws: per-task kernels
        if (ctx->cfg.kernel_path == 1) ok = false;
        if (ctx->cfg.kernel_path == 2 && !ok)
            return fail(ctx, MPMGPU_EINVAL, "kernel_path=2 (fused) needs 3D uGIMP, lp<=1, no mirrored rigid BCs and no large-rotation or Mooney materials");
        ctx->tiled.enabled = ok ? 1 : 0;
        ctx->tiled.sortInterval = ctx->cfg.sort_interval > 0 ? ctx->cfg.sort_interval : 12;
        {
            // TMA-pipelined F4 (kernels_pipe.cuh): measured on B200 within 2 % of the plain kernel
            // (profiles/tune_bounds_r1.txt), so it is opt-in: MPMGPU_PIPE=1
            const char *e = getenv("MPMGPU_PIPE");
            ctx->tiled.usePipe = e ? atoi(e) : 0;
            cudaDeviceProp prop;
            cudaGetDeviceProperties(&prop, ctx->cfg.device);
            ctx->tiled.numSMs = prop.multiProcessorCount;
        }
    }
    return MPMGPU_OK;
}

extern "C" int mpmgpu_set_time_step(mpmgpu_ctx *ctx, double dt, double dtFirst, double dtLast)
{
    if (!ctx || !(dt > 0.)) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_time_step: dt must be > 0");
    ctx->sp.dt = dt; ctx->sp.dtStrainFirst = dtFirst; ctx->sp.dtStrainLast = dtLast;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_set_xpic(mpmgpu_ctx *ctx, int order, int usingFMPM)
{
    if (ctx && ctx->multimaterial && order > 1) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_xpic: XPIC/FMPM of order > 1 with material contact is not built");
    if (!ctx) return MPMGPU_EINVAL;
    if (order < 0) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_xpic: order %d", order);
    ctx->sp.xpicOrder = order; ctx->sp.usingFMPM = usingFMPM;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_set_velocity_bcs(mpmgpu_ctx *ctx, int n, const int *node, const double *norm,
                                       const double *value, const int *active, const int *symdir)
{
    if (!ctx || n < 0) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_velocity_bcs: bad argument");
    cudaSetDevice(ctx->cfg.device);
    ctx->hasBCs = n > 0;
    ctx->nBCEntries = n;
    {   // dofs fixed by grid BCs, for rigid-particle projection (nd[]->fixedDirection bits 1,2,4: NodalVelBC.cpp:40-45)
        ctx->hFixedBits.assign((size_t)ctx->g.nnodes, 0);
        for (int i = 0; i < n && node && norm; i++) {
            if (node[i] < 1 || node[i] > ctx->g.nnodes) continue;
            unsigned char b = symdir ? (unsigned char)(symdir[i] & 7) : 0;
            for (int d = 0; d < 3; d++) if (norm[3 * i + d] != 0.) b |= (unsigned char)(1 << d);
            ctx->hFixedBits[node[i] - 1] |= b;
        }
        if (ctx->R.fixedBits) cudaMemcpy((void *)ctx->R.fixedBits, ctx->hFixedBits.data(), ctx->hFixedBits.size(), cudaMemcpyHostToDevice);
    }
    if (n == 0) { ctx->B.nUnique = 0; ctx->tiled.FN.bcOfNode = NULL; return reaction_buffers(ctx); }
    if (!node || !norm || !value) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_velocity_bcs: null arrays");
    for (int i = 0; i < n; i++)
        if (node[i] < 1 || node[i] > ctx->g.nnodes) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_velocity_bcs: BC %d on node %d of %d", i, node[i], ctx->g.nnodes);
    // group by node, keeping list order inside a node
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return node[a] < node[b]; });
    std::vector<int> un, st, sd, ac(n);
    std::vector<double> nm(3 * (size_t)n), va(n);
    for (int e = 0; e < n; e++) {
        int i = order[e];
        if (e == 0 || node[i] != node[order[e - 1]]) { un.push_back(node[i] - 1); st.push_back(e); sd.push_back(0); }
        if (symdir) sd.back() |= symdir[i];
        nm[3 * e] = norm[3 * i]; nm[3 * e + 1] = norm[3 * i + 1]; nm[3 * e + 2] = norm[3 * i + 2];
        va[e] = value[i]; ac[e] = active ? active[i] : 1;
    }
    st.push_back(n);
    ctx->bcOrder = order;
    int nu = (int)un.size();
    // (buffers are kept between calls and grow on demand: a host may hand the list over every step)
    if (nu > ctx->bcCapUnique) {
        CK(dalloc(ctx, &ctx->dBcNode, nu)); CK(dalloc(ctx, &ctx->dBcStart, nu + 1)); CK(dalloc(ctx, &ctx->dBcSym, nu));
        ctx->bcCapUnique = nu;
    }
    if (n > ctx->bcCapEntries) {
        CK(dalloc(ctx, &ctx->dBcActive, n)); CK(dalloc(ctx, &ctx->dBcNorm, 3 * (size_t)n)); CK(dalloc(ctx, &ctx->dBcValue, n));
        ctx->bcCapEntries = n;
    }
    int *dn = ctx->dBcNode, *ds = ctx->dBcStart, *dsd = ctx->dBcSym, *da = ctx->dBcActive; double *dnm = ctx->dBcNorm, *dva = ctx->dBcValue;
    CK(cudaStreamSynchronize(ctx->stream));        // a step in flight may still read the old contents
    CK(cudaMemcpy(dn, un.data(), nu * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ds, st.data(), (nu + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dsd, sd.data(), nu * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, ac.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dnm, nm.data(), 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dva, va.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    ctx->B.refl = NULL; ctx->B.reflRatio = NULL;
    ctx->B.nUnique = nu; ctx->B.node = dn; ctx->B.start = ds; ctx->B.symdir = dsd; ctx->B.active = da; ctx->B.norm = dnm; ctx->B.value = dva;
    {   // node -> BC group lookup for the fused node sweeps
        std::vector<int> of(ctx->g.nnodes, -1);
        for (int u = 0; u < nu; u++) of[un[u]] = u;
        if (!ctx->dBcOfNode) CK(dalloc(ctx, &ctx->dBcOfNode, (size_t)ctx->g.nnodes));
        CK(cudaMemcpy(ctx->dBcOfNode, of.data(), (size_t)ctx->g.nnodes * sizeof(int), cudaMemcpyHostToDevice));
        ctx->tiled.FN.bcOfNode = ctx->dBcOfNode;
    }
    return reaction_buffers(ctx);
}

// Symmetry-plane BCs (<Horiz symmin=...>, Generators.cpp:2178-2190): entry i of the list set by mpmgpu_set_velocity_bcs
// reflects node reflected_node[i] (1-based; <= 0: a plain BC) with the cell-size ratio ratio[i].
extern "C" int mpmgpu_set_velocity_bc_reflections(mpmgpu_ctx *ctx, int n, const int *reflected_node, const double *ratio)
{
    if (!ctx || n != ctx->nBCEntries) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_velocity_bc_reflections: n=%d but %d BCs are set", n, ctx ? ctx->nBCEntries : 0);
    if (n == 0) return MPMGPU_OK;
    if (!reflected_node || !ratio) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_velocity_bc_reflections: null arrays");
    cudaSetDevice(ctx->cfg.device);
    std::vector<int> re(n); std::vector<double> ra(n);
    bool any = false;
    for (int e = 0; e < n; e++) {
        const int i = ctx->bcOrder[e];
        if (reflected_node[i] > ctx->g.nnodes) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_velocity_bc_reflections: BC %d reflects node %d of %d", i, reflected_node[i], ctx->g.nnodes);
        re[e] = reflected_node[i] > 0 ? reflected_node[i] - 1 : -1; ra[e] = ratio[i];
        any |= re[e] >= 0;
    }
    if (!any) { ctx->B.refl = NULL; ctx->B.reflRatio = NULL; return MPMGPU_OK; }
    if (ctx->cfg.kernel_path == 2) return fail(ctx, MPMGPU_EINVAL, "kernel_path=2 (fused) cannot apply reflected (symmetry-plane) velocity BCs");
    if (n > ctx->bcCapRefl) { CK(dalloc(ctx, &ctx->dBcRefl, n)); CK(dalloc(ctx, &ctx->dBcReflRatio, n)); ctx->bcCapRefl = n; }
    int *dre = ctx->dBcRefl; double *dra = ctx->dBcReflRatio;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(dre, re.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dra, ra.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    ctx->B.refl = dre; ctx->B.reflRatio = dra;
    ctx->hasReflectedBCs = true;
    ctx->tiled.enabled = 0;         // the reflected node's momentum must be complete before the BC reads it: per-task kernels
    return MPMGPU_OK;
}

// Particle loads (MatPtLoadBC::SetParticleFext at the start of every step, InitializationTask.cpp:91, MatPtLoadBC.cpp:210-222): the
// host evaluates the load BCs at this step's time and hands over the external force of the loaded particles only.
// index_of_particle[i] stays on the device after the first call (NULL afterwards = same particles as before).
__global__ void k_scatter_particle_loads(int n, Particles P, const int *loadOf, int nload, const double *fext)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int k = loadOf[P.orig[p]];
    if (k < 0) return;
    P.pfext[0][p] = fext[k]; P.pfext[1][p] = fext[nload + k]; P.pfext[2][p] = fext[2 * nload + k];
}

extern "C" int mpmgpu_update_particle_loads(mpmgpu_ctx *ctx, int n_loaded, const int *particle, const double *fext)
{
    if (!ctx || !fext || n_loaded < 0) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_particle_loads: bad argument");
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_update_particle_loads: upload the particles first");
    if (!ctx->hasFext) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_update_particle_loads: the particles were uploaded without an external-force array (pfext)");
    if (ctx->globalIds) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_update_particle_loads: not available in slab mode");
    if (n_loaded == 0) return MPMGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    const int nNR = ctx->P.n, nAll = nNR + ctx->PR.n;
    if (particle) {
        std::vector<int> of((size_t)nAll, -1);
        for (int k = 0; k < n_loaded; k++) {
            if (particle[k] < 0 || particle[k] >= nNR) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_particle_loads: entry %d is particle %d of %d non-rigid particles", k, particle[k], nNR);
            if (of[particle[k]] >= 0) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_particle_loads: particle %d listed twice", particle[k]);
            of[particle[k]] = k;
        }
        if (!ctx->dLoadOf) CK(dalloc(ctx, &ctx->dLoadOf, (size_t)nAll));
        if (ctx->loadCap < n_loaded) { CK(dalloc(ctx, &ctx->dLoadFext, (size_t)3 * n_loaded)); ctx->loadCap = n_loaded; }
        CK(cudaMemcpyAsync(ctx->dLoadOf, of.data(), (size_t)nAll * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));        // `of` goes out of scope
        ctx->nLoaded = n_loaded;
    } else if (n_loaded != ctx->nLoaded) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_particle_loads: %d forces for %d loaded particles", n_loaded, ctx->nLoaded);
    CK(cudaMemcpyAsync(ctx->dLoadFext, fext, (size_t)3 * n_loaded * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(k_scatter_particle_loads, nblocks(nNR, 256), 256, nNR, ctx->P, ctx->dLoadOf, n_loaded, ctx->dLoadFext);
    CK(cudaStreamSynchronize(ctx->stream));            // the caller's array may change after the return
    return MPMGPU_OK;
}

// Particle BCs on a face of the particle domain, in the host's list order: tractions (MatPtTractionBC list, firstTractionPt ...:
// stress value[i] in direction[i]) and heat fluxes (MatPtHeatFluxBC list, firstHeatFluxPt ...: external flux value[i]).  particle[i] is
// the 0-based host index of a non-rigid particle, value[i] = BCValue at this step's time.
static int set_face_bcs(mpmgpu_ctx *ctx, mpmgpu_ctx::FaceBCs &S, const char *who, int n, const int *particle, const int *face, const int *direction, const double *value)
{
    if (!ctx || n < 0 || (n > 0 && (!particle || !face || !value))) return fail(ctx, MPMGPU_EINVAL, "%s: bad argument", who);
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "%s: upload the particles first", who);
    if (ctx->globalIds || ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "%s: not available in slab mode", who);
    cudaSetDevice(ctx->cfg.device);
    const int nNR = ctx->P.n;
    if (n == 0) { S.TB.n = 0; S.order.clear(); return MPMGPU_OK; }
    const int nfaces = ctx->dim == 3 ? 6 : 4;
    for (int i = 0; i < n; i++) {
        if (particle[i] < 0 || particle[i] >= nNR) return fail(ctx, MPMGPU_EINVAL, "%s: entry %d is particle %d of %d non-rigid particles", who, i, particle[i], nNR);
        if (face[i] < 1 || face[i] > nfaces) return fail(ctx, MPMGPU_EINVAL, "%s: entry %d has face %d (1..%d)", who, i, face[i], nfaces);
        const int d = direction ? direction[i] : 1;
        if (!(d == 1 || d == 2 || (d == 3 && ctx->dim == 3) || d == 11 || (d == 12 && ctx->dim == 2)))
            return fail(ctx, MPMGPU_EINVAL, "%s: entry %d has direction %d (1 x, 2 y, 3 z in 3D, 11 normal, 12 tangent in 2D)", who, i, d);
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return particle[a] < particle[b]; });
    std::vector<int> st((size_t)nNR + 1, 0), fa(n), di(n);
    std::vector<double> va(n);
    for (int e = 0; e < n; e++) { const int i = order[e]; st[particle[i] + 1]++; fa[e] = face[i]; di[e] = direction ? direction[i] : 1; va[e] = value[i]; }
    for (int i = 0; i < nNR; i++) st[i + 1] += st[i];
    if (S.startLen < nNR + 1) { CK(dalloc(ctx, &S.dStart, (size_t)nNR + 1)); S.startLen = nNR + 1; }
    if (S.cap < n) {
        CK(dalloc(ctx, &S.dFace, (size_t)n)); CK(dalloc(ctx, &S.dDir, (size_t)n)); CK(dalloc(ctx, &S.dValue, (size_t)n));
        S.cap = n;
    }
    CK(cudaMemcpyAsync(S.dStart, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S.dFace, fa.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S.dDir, di.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(S.dValue, va.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    S.TB.n = n; S.TB.start = S.dStart; S.TB.face = S.dFace; S.TB.dir = S.dDir; S.TB.value = S.dValue;
    S.order.swap(order);
    return MPMGPU_OK;
}

// the values of the same list at a new time (BCs that vary)
static int update_face_bc_values(mpmgpu_ctx *ctx, mpmgpu_ctx::FaceBCs &S, const char *who, int n, const double *value)
{
    if (!ctx || !value || n != S.TB.n) return fail(ctx, MPMGPU_EINVAL, "%s: n=%d but %d BCs are set", who, n, ctx ? S.TB.n : 0);
    if (n == 0) return MPMGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    std::vector<double> va(n);
    for (int e = 0; e < n; e++) va[e] = value[S.order[e]];
    CK(cudaMemcpyAsync(S.dValue, va.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_set_particle_tractions(mpmgpu_ctx *ctx, int n, const int *particle, const int *face, const int *direction, const double *value)
{
    if (ctx && n > 0 && !direction) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_particle_tractions: bad argument");
    return set_face_bcs(ctx, ctx->trac, "mpmgpu_set_particle_tractions", n, particle, face, direction, value);
}
extern "C" int mpmgpu_update_particle_traction_values(mpmgpu_ctx *ctx, int n, const double *value)
{
    return update_face_bc_values(ctx, ctx->trac, "mpmgpu_update_particle_traction_values", n, value);
}
extern "C" int mpmgpu_set_particle_heat_fluxes(mpmgpu_ctx *ctx, int n, const int *particle, const int *face, const double *value)
{
    if (ctx && !ctx->conduction) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_particle_heat_fluxes: call mpmgpu_set_conduction first");
    return set_face_bcs(ctx, ctx->flux, "mpmgpu_set_particle_heat_fluxes", n, particle, face, NULL, value);
}
extern "C" int mpmgpu_update_particle_heat_flux_values(mpmgpu_ctx *ctx, int n, const double *value)
{
    return update_face_bc_values(ctx, ctx->flux, "mpmgpu_update_particle_heat_flux_values", n, value);
}

static int face_bc_launch(mpmgpu_ctx *ctx, const TractionBCs &TB, double *fluxQ)
{
    if (TB.n <= 0 || ctx->P.nNR <= 0) return MPMGPU_OK;
    const int cpdi = ctx->cfg.shape == MPMGPU_LINEAR_CPDI || ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI || ctx->cfg.shape == MPMGPU_BSPLINE_CPDI ? 1 : 0;
    const int spline = ctx->cfg.shape == MPMGPU_BSPLINE || ctx->cfg.shape == MPMGPU_BSPLINE_GIMP || ctx->cfg.shape == MPMGPU_BSPLINE_CPDI ? 1 : 0;
    const double thick = ctx->cfg.thickness > 0. ? ctx->cfg.thickness : 1.;
    if (ctx->dim == 3) LAUNCH((k_particle_tractions<3>), nblocks(ctx->P.nNR, TASK_THREADS), TASK_THREADS, ctx->g, ctx->P, ctx->N, TB, cpdi, thick, ctx->nf, ctx->dFlags, fluxQ, spline);
    else LAUNCH((k_particle_tractions<2>), nblocks(ctx->P.nNR, TASK_THREADS), TASK_THREADS, ctx->g, ctx->P, ctx->N, TB, cpdi, thick, ctx->nf, ctx->dFlags, fluxQ, spline);
    return MPMGPU_OK;
}
static int particle_tractions(mpmgpu_ctx *ctx) { return face_bc_launch(ctx, ctx->trac.TB, NULL); }
// TransportTask::TransportForceBCs at the end of the post-forces task (PostForcesTask.cpp:97)
static int particle_heat_fluxes(mpmgpu_ctx *ctx) { return ctx->conduction ? face_bc_launch(ctx, ctx->flux.TB, ctx->T.gQ) : MPMGPU_OK; }

// Rigid-BC particles whose material has setting functions: the host evaluates them each step
// (RigidMaterial::GetVectorSetting, Materials/RigidMaterial.cpp:376-531) and hands over the velocities
extern "C" int mpmgpu_update_rigid_velocities(mpmgpu_ctx *ctx, int n_rigid, const double *vel)
{
    if (!ctx || !vel) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_rigid_velocities: null argument");
    if (!ctx->uploaded || n_rigid != ctx->PR.n) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_rigid_velocities: %d velocities for %d rigid particles", n_rigid, ctx ? ctx->PR.n : 0);
    if (n_rigid == 0) return MPMGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    for (int c = 0; c < 3; c++)
        CK(cudaMemcpyAsync(ctx->PR.vel[c], vel + (size_t)c * n_rigid, (size_t)n_rigid * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

// Rigid particles whose material sets the temperature by a value function (RigidMaterial::GetValueSetting, evaluated by the host each
// step as ProjectRigidBCsTask.cpp:120 does): pTemperature of every rigid particle in host order
extern "C" int mpmgpu_update_rigid_temperatures(mpmgpu_ctx *ctx, int n_rigid, const double *temperature)
{
    if (!ctx || !temperature) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_rigid_temperatures: null argument");
    if (!ctx->uploaded || n_rigid != ctx->PR.n) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_rigid_temperatures: %d temperatures for %d rigid particles", n_rigid, ctx ? ctx->PR.n : 0);
    if (!ctx->rigidTemp) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_update_rigid_temperatures: no rigid material sets the temperature (material slot 10) or conduction is off");
    if (n_rigid == 0) return MPMGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    CK(cudaMemcpyAsync(ctx->PR.temp, temperature, (size_t)n_rigid * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_update_velocity_bc_values(mpmgpu_ctx *ctx, int n, const double *value, const int *active)
{
    if (!ctx || n != ctx->nBCEntries) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_update_velocity_bc_values: n=%d but %d BCs are set", n, ctx ? ctx->nBCEntries : 0);
    if (n == 0) return MPMGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    std::vector<double> va(n); std::vector<int> ac(n);
    for (int e = 0; e < n; e++) { int i = ctx->bcOrder[e]; va[e] = value[i]; ac[e] = active ? active[i] : 1; }
    CK(cudaMemcpyAsync((void *)ctx->B.value, va.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((void *)ctx->B.active, ac.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// task bodies
#define DISPATCH_DIM_SHAPE_M(MERGE, KERNEL, n, ...) do { \
    const int grid_ = nblocks((n), TASK_THREADS); const bool merge_ = (MERGE); \
    if (grid_ > 0) { \
        if (ctx->dim == 3) { \
            if (ctx->cfg.shape == MPMGPU_UNIFORM_GIMP) LAUNCH((KERNEL<3, SHAPE_UGIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_BSPLINE) LAUNCH((KERNEL<3, SHAPE_B2SPLINE>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_BSPLINE_GIMP) LAUNCH((KERNEL<3, SHAPE_B2GIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_BSPLINE_CPDI) LAUNCH((KERNEL<3, SHAPE_B2CPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_LINEAR_CPDI && merge_) LAUNCH((KERNEL<3, SHAPE_LCPDI_MERGED>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_LINEAR_CPDI) LAUNCH((KERNEL<3, SHAPE_LCPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
            else LAUNCH((KERNEL<3, SHAPE_LINEAR>), grid_, TASK_THREADS, __VA_ARGS__); \
        } else { \
            if (ctx->cfg.shape == MPMGPU_UNIFORM_GIMP) LAUNCH((KERNEL<2, SHAPE_UGIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_BSPLINE) LAUNCH((KERNEL<2, SHAPE_B2SPLINE>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_BSPLINE_GIMP) LAUNCH((KERNEL<2, SHAPE_B2GIMP>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_BSPLINE_CPDI) LAUNCH((KERNEL<2, SHAPE_B2CPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_LINEAR_CPDI && merge_) LAUNCH((KERNEL<2, SHAPE_LCPDI_MERGED>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_LINEAR_CPDI) LAUNCH((KERNEL<2, SHAPE_LCPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI && merge_) LAUNCH((KERNEL<2, SHAPE_QCPDI_MERGED>), grid_, TASK_THREADS, __VA_ARGS__); \
            else if (ctx->cfg.shape == MPMGPU_QUADRATIC_CPDI) LAUNCH((KERNEL<2, SHAPE_QCPDI>), grid_, TASK_THREADS, __VA_ARGS__); \
            else LAUNCH((KERNEL<2, SHAPE_LINEAR>), grid_, TASK_THREADS, __VA_ARGS__); \
        } } } while (0)
// kernels that use shape-function VALUES only (mass/momentum scatters, the XPIC iteration) merge the CPDI corners' contributions
// per node before touching the grid: 2.4x faster on B200 (neo8m mass+momentum 8.2 -> 3.4 ms; profiles/r2_experiments); the
// kernels that need gradients keep the plain corner loop (merged, the 27x4 thread-local sums spill: 2.4 -> 12.6 ms)
#define DISPATCH_DIM_SHAPE(KERNEL, n, ...) DISPATCH_DIM_SHAPE_M(ctx->cpdiMerge, KERNEL, n, __VA_ARGS__)
#define DISPATCH_DIM_SHAPE_VALUES(KERNEL, n, ...) DISPATCH_DIM_SHAPE_M(ctx->cpdiMergeValues, KERNEL, n, __VA_ARGS__)

static int check_ready(mpmgpu_ctx *ctx, const char *who)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "%s: no particles uploaded", who);
    if (!(ctx->sp.dt > 0.)) return fail(ctx, MPMGPU_ESTATE, "%s: time step not set", who);
    cudaSetDevice(ctx->cfg.device);
    return MPMGPU_OK;
}

static void prof_begin(mpmgpu_ctx *ctx) { if (ctx->profiling) cudaEventRecord(ctx->ev0, ctx->stream); }
static void prof_end(mpmgpu_ctx *ctx, int task)
{
    if (!ctx->profiling) return;
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaEventSynchronize(ctx->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->taskMs[task] += ms; ctx->taskCalls[task]++;
}

static int apply_bcs(mpmgpu_ctx *ctx, int pass, int adjustSym)
{
    if (ctx->hasBCs && ctx->B.nUnique > 0)
        LAUNCH(k_velocity_bcs, nblocks((long long)ctx->B.nUnique * ctx->nf, 128), 128, ctx->B, ctx->N, pass, ctx->sp.dt, adjustSym, ctx->nf, ctx->g.nnodes);
    if (ctx->R.on && adjustSym != 2)
        LAUNCH(k_rigid_velocity_bcs, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->R, ctx->N, pass, ctx->sp.dt);
    return MPMGPU_OK;
}

// multimaterial mode: the contact extrapolations of the particles (after a mass/momentum extrapolation) and the contact
// pass over the nodes (UpdateMomentaTask::ContactAndMomentaBCs, UpdateMomentaTask.cpp:76-99: contact first, then the BCs)
static int contact_extrapolation(mpmgpu_ctx *ctx, bool rigidToo = false)
{
    if (!ctx->multimaterial) return MPMGPU_OK;
    const size_t norig = (size_t)ctx->P.n + (size_t)ctx->PR.n;
    if (rigidToo && ctx->cp.rigidMask && ctx->PR.n > 0)
        DISPATCH_DIM_SHAPE(k_p2g_rigid_contact, ctx->PR.n, ctx->g, ctx->PR, ctx->dMats, ctx->N, ctx->C, ctx->archOrigin, norig, ctx->cp.byDisplacements,
                           ctx->cp.normalMethod != NORMALS_SPECIFIED ? 1 : 0, ctx->dFlags);
    DISPATCH_DIM_SHAPE(k_p2g_contact_terms, ctx->P.nNR, ctx->g, ctx->P, ctx->dMats, ctx->C, ctx->archOrigin, norig, ctx->cp.byDisplacements,
                       ctx->cp.normalMethod != NORMALS_SPECIFIED ? 1 : 0);
    return MPMGPU_OK;
}

static int material_contact(mpmgpu_ctx *ctx, int callType)
{
    if (!ctx->multimaterial) return MPMGPU_OK;
    LAUNCH(k_material_contact, nblocks(ctx->g.nnodes, 128), 128, ctx->g, ctx->N, ctx->C, ctx->cp, ctx->B,
           ctx->hasBCs ? ctx->tiled.FN.bcOfNode : (const int *)NULL, callType, ctx->sp.dt);
    return MPMGPU_OK;
}

// ProjectRigidBCsTask: runs between mass/momentum and post-extrapolation (NairnMPM.cpp:982-987)
static int t_project_rigid_bcs(mpmgpu_ctx *ctx)
{
    if (!ctx->R.on) return MPMGPU_OK;
    for (int d = 0; d < 3; d++) CK(cudaMemsetAsync(ctx->R.owner[d], 0x7f, (size_t)ctx->g.nnodes * sizeof(int), ctx->stream));
    if (ctx->R.ownerT) CK(cudaMemsetAsync(ctx->R.ownerT, 0x7f, (size_t)ctx->g.nnodes * sizeof(int), ctx->stream));
    ctx->launches += 3;
    DISPATCH_DIM_SHAPE(k_project_rigid_bcs, ctx->PR.n, ctx->g, ctx->PR, ctx->dMats, ctx->R, ctx->dFlags);
    return MPMGPU_OK;
}

static int move_rigid(mpmgpu_ctx *ctx)
{
    if (ctx->PR.n <= 0) return MPMGPU_OK;
    if (ctx->dim == 3) LAUNCH(k_move_rigid<3>, nblocks(ctx->PR.n, 128), 128, ctx->PR, ctx->sp.dt);
    else LAUNCH(k_move_rigid<2>, nblocks(ctx->PR.n, 128), 128, ctx->PR, ctx->sp.dt);
    return MPMGPU_OK;
}

static int reset_rigid(mpmgpu_ctx *ctx)
{
    if (ctx->PR.n <= 0) return MPMGPU_OK;
    if (ctx->dim == 3) LAUNCH(k_reset_elements<3>, nblocks(ctx->PR.n, TASK_THREADS), TASK_THREADS, ctx->g, ctx->PR, ctx->dFlags, ctx->sp.dt);
    else LAUNCH(k_reset_elements<2>, nblocks(ctx->PR.n, TASK_THREADS), TASK_THREADS, ctx->g, ctx->PR, ctx->dFlags, ctx->sp.dt);
    return MPMGPU_OK;
}

static int t_initialization(mpmgpu_ctx *ctx)
{
    const Grid &g = ctx->g;
    (void)g;
    // (before the point counts of the previous step are cleared: the transport field is zeroed on the nodes that were active)
    if (ctx->conduction) LAUNCH(k_transport_zero_active, nblocks(ctx->g.nnodes, 256), 256, ctx->g.nnodes, ctx->nf, ctx->N, ctx->T);
    size_t nnPad = ctx->nodePad;
    // MatVelocityField::Zero: mass, pk, ftot, vk[0], vk[pkCopy] (+XPIC vectors), numberPoints
    CK(cudaMemsetAsync(ctx->nodePool, 0, nnPad * 22 * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->N.cnt, 0, nnPad * sizeof(int), ctx->stream));
    ctx->launches += 2;
    if (ctx->multimaterial) {
        CK(cudaMemsetAsync(ctx->contactPool, 0, nnPad * 7 * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(ctx->C.rcnt, 0, nnPad * sizeof(int), ctx->stream));
        ctx->launches += 2;
    }

    DISPATCH_DIM_SHAPE(k_init_particles, ctx->P.n, ctx->g, ctx->P, ctx->dFlags);
    return MPMGPU_OK;
}

static int t_mass_and_momentum(mpmgpu_ctx *ctx)
{
    DISPATCH_DIM_SHAPE_VALUES(k_p2g_mass_momentum, ctx->P.nNR, ctx->g, ctx->P, ctx->N);
    if (ctx->conduction) DISPATCH_DIM_SHAPE_VALUES(k_p2g_temperature, ctx->P.nNR, ctx->g, ctx->P, ctx->dMats, ctx->T);
    return contact_extrapolation(ctx, true);
}

static int t_post_extrapolation(mpmgpu_ctx *ctx)
{
    LAUNCH(k_copy_momenta, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->N);
    { int rc = material_contact(ctx, CALL_MASS_MOMENTUM); if (rc) return rc; }
    // MASS_MOMENTUM_CALL: symmetry adjust always, BC loop only when a USF task exists (NodalVelBC.cpp:339-361)
    const bool hasUSF = ctx->sp.method == METHOD_USF || ctx->sp.method == METHOD_USAVG;
    { int rc = apply_bcs(ctx, PASS_MASS_MOMENTUM, hasUSF ? 1 : 2); if (rc) return rc; }
    if (ctx->conduction) {      // TransportTask::GetTransportValues + TransportBCsAndGradients (PostExtrapolationTask.cpp:88,160)
        LAUNCH(k_transport_nodal_value, nblocks(ctx->g.nnodes, 256), 256, ctx->g.nnodes, ctx->nf, ctx->N, ctx->T);
        if (ctx->Q.nUnique > 0) LAUNCH(k_temp_bcs_impose, nblocks(ctx->Q.nUnique, 128), 128, ctx->Q, ctx->T, 0);
        if (ctx->R.ownerT) LAUNCH(k_rigid_temp_bcs_impose, nblocks(ctx->g.nnodes, 256), 256, ctx->g.nnodes, ctx->R, ctx->T, 0);
        DISPATCH_DIM_SHAPE(k_transport_gradients, ctx->P.nNR, ctx->g, ctx->P, ctx->T);
        if (ctx->Q.nUnique > 0) LAUNCH(k_temp_bcs_impose, nblocks(ctx->Q.nUnique, 128), 128, ctx->Q, ctx->T, 1);
        if (ctx->R.ownerT) LAUNCH(k_rigid_temp_bcs_impose, nblocks(ctx->g.nnodes, 256), 256, ctx->g.nnodes, ctx->R, ctx->T, 1);
    }
    return MPMGPU_OK;
}

// XPICExtrapolationTask::Execute (XPICExtrapolationTask.cpp:49-161): grid velocity v(k) for order k > 1
static int xpic_extrapolation(mpmgpu_ctx *ctx, int particleUpdate)
{
    const int nn = ctx->nvn, fmpm = ctx->sp.usingFMPM ? 1 : 0;
    LAUNCH(k_xpic_init, nblocks(nn, 256), 256, nn, ctx->N, ctx->sp.dt, fmpm);
    for (int k = 2; k <= ctx->sp.xpicOrder; k++) {
        DISPATCH_DIM_SHAPE_VALUES(k_xpic_iterate, ctx->P.nNR, ctx->g, ctx->P, ctx->N);
        LAUNCH(k_xpic_finish, nblocks(nn, 256), 256, nn, ctx->N, ctx->B, ctx->hasBCs ? ctx->tiled.FN.bcOfNode : (const int *)NULL, ctx->R,
               ctx->sp.dt, particleUpdate, fmpm);
    }
    return MPMGPU_OK;
}

static int strain_update(mpmgpu_ctx *ctx, double strainTime, bool postUpdate = false)
{
    // FMPM(k>1) strain updates use v(k) (UpdateStrainsFirstTask.cpp:105-116); XPIC(k) and FLIP use the lumped velocity
    if (ctx->sp.usingFMPM && ctx->sp.xpicOrder > 1) {
        if (!postUpdate || !ctx->sp.skipPost) { int rc = xpic_extrapolation(ctx, 0); if (rc) return rc; }
    } else
    LAUNCH(k_grid_velocity, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->N);
    // MPMBase::ScaledResidualStrains: the two passes of USAVG share the step's temperature change
    const double dTscale = ctx->sp.method == METHOD_USAVG ? (postUpdate ? 1.0 - ctx->sp.fractionUSF : ctx->sp.fractionUSF) : 1.0;
    if (ctx->largeRotation) DISPATCH_DIM_SHAPE(k_update_strains_lr, ctx->P.nNR, ctx->g, ctx->P, ctx->N, ctx->dMats, strainTime, dTscale);
    else DISPATCH_DIM_SHAPE(k_update_strains, ctx->P.nNR, ctx->g, ctx->P, ctx->N, ctx->dMats, strainTime, dTscale);
    return MPMGPU_OK;
}

static int t_update_strains_first(mpmgpu_ctx *ctx)
{
    if (ctx->sp.method == METHOD_USL) return MPMGPU_OK;       // no USF task in the list (NairnMPM.cpp:1004-1010)
    double st = ctx->sp.method == METHOD_USAVG ? ctx->sp.dtStrainFirst : ctx->sp.dt;
    return strain_update(ctx, st);
}

static int t_grid_forces(mpmgpu_ctx *ctx)
{
    DISPATCH_DIM_SHAPE(k_p2g_forces, ctx->P.nNR, ctx->g, ctx->P, ctx->N, ctx->hasFext ? 1 : 0);
    if (ctx->conduction) DISPATCH_DIM_SHAPE(k_p2g_conduction, ctx->P.nNR, ctx->g, ctx->P, ctx->T);        // GridForcesTask.cpp:108-112
    return MPMGPU_OK;
}

static int t_post_forces(mpmgpu_ctx *ctx)
{
    int rc = particle_tractions(ctx);          // PostForcesTask.cpp:51, before the body forces
    if (rc) return rc;
    LAUNCH(k_post_forces, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->N, ctx->sp);
    rc = reactions_zero(ctx);
    if (rc) return rc;
    if ((rc = apply_bcs(ctx, PASS_GRID_FORCES, 0))) return rc;
    return particle_heat_fluxes(ctx);
}

static int t_update_momenta(mpmgpu_ctx *ctx)
{
    LAUNCH(k_update_momenta, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->N, ctx->sp.dt);
    if (ctx->conduction) LAUNCH(k_transport_update, nblocks(ctx->g.nnodes, 256), 256, ctx->g.nnodes, ctx->nf, ctx->N, ctx->T, ctx->sp.dt);    // UpdateMomentaTask.cpp:55
    { int rc = material_contact(ctx, CALL_UPDATE_MOMENTUM); if (rc) return rc; }
    if (ctx->sp.xpicOrder <= 1) { int rc = apply_bcs(ctx, PASS_UPDATE_MOMENTUM, 0); if (rc) return rc; }      // NodalVelBC.cpp:367-375
    // TransportTask::TransportGridBCs (UpdateMomentaTask.cpp:61)
    if (ctx->conduction && ctx->Q.nUnique > 0) LAUNCH(k_temp_bcs_grid, nblocks(ctx->Q.nUnique, 128), 128, ctx->Q, ctx->T, ctx->sp.dt);
    if (ctx->conduction && ctx->R.ownerT) LAUNCH(k_rigid_temp_bcs_grid, nblocks(ctx->g.nnodes, 256), 256, ctx->g.nnodes, ctx->R, ctx->T, ctx->sp.dt);
    return MPMGPU_OK;
}

static int t_update_particles(mpmgpu_ctx *ctx)
{
    if (ctx->sp.xpicOrder > 1) { int rc = xpic_extrapolation(ctx, 1); if (rc) return rc; }      // UpdateParticlesTask.cpp:66-71
    else LAUNCH(k_grid_velocity, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->N);
    int m = ctx->sp.xpicOrder;
    if (!ctx->sp.usingFMPM) m = -m;
    DISPATCH_DIM_SHAPE(k_update_particles, ctx->P.nNR, ctx->g, ctx->P, ctx->N, ctx->dMats, ctx->sp, m);
    if (ctx->conduction) DISPATCH_DIM_SHAPE_VALUES(k_update_temperature, ctx->P.nNR, ctx->g, ctx->P, ctx->dMats, ctx->T, ctx->sp.dt);
    else if (ctx->thermal) LAUNCH(k_update_temperature_offsets, nblocks(ctx->P.nNR, 256), 256, ctx->P.nNR, ctx->P);      // UpdateParticlesTask.cpp:246-251
    return move_rigid(ctx);
}

static int t_update_strains_last(mpmgpu_ctx *ctx)
{
    if (ctx->sp.method == METHOD_USF) return MPMGPU_OK;
    if (!ctx->sp.skipPost) {
        if (ctx->multimaterial) LAUNCH(k_rezero_fields_task6, nblocks(ctx->nvn, 256), 256, ctx->g.nnodes, ctx->nf, ctx->cp.rigidMask, ctx->N, ctx->C, ctx->sp.dt);
        else LAUNCH(k_rezero_momenta, nblocks(ctx->nvn, 256), 256, ctx->nvn, ctx->N);
        DISPATCH_DIM_SHAPE_VALUES(k_p2g_momentum_last, ctx->P.nNR, ctx->g, ctx->P, ctx->N);
        int rc = contact_extrapolation(ctx);
        if (!rc) rc = material_contact(ctx, CALL_UPDATE_STRAINS_LAST);
        if (!rc) rc = apply_bcs(ctx, PASS_UPDATE_STRAINS_LAST, 0);
        if (rc) return rc;
    }
    double st = ctx->sp.method == METHOD_USAVG ? ctx->sp.dtStrainLast : ctx->sp.dt;
    return strain_update(ctx, st, true);
}

static int t_reset_elements(mpmgpu_ctx *ctx)
{
    const int n = ctx->P.n;
    if (n == 0) return reset_rigid(ctx);
    if (ctx->dim == 3) LAUNCH(k_reset_elements<3>, nblocks(n, TASK_THREADS), TASK_THREADS, ctx->g, ctx->P, ctx->dFlags, ctx->sp.dt);
    else LAUNCH(k_reset_elements<2>, nblocks(n, TASK_THREADS), TASK_THREADS, ctx->g, ctx->P, ctx->dFlags, ctx->sp.dt);
    return reset_rigid(ctx);
}

static int poll_flags(mpmgpu_ctx *ctx, bool force = true)
{
    // the status word costs a stream synchronisation: with a poll interval k > 1 the host keeps enqueueing steps and looks at it
    // every k-th call (an error -- NaN position, CPDI corner off the grid -- is then reported up to k-1 steps late)
    if (!force && ++ctx->callsSincePoll < ctx->pollInterval) return MPMGPU_OK;
    ctx->callsSincePoll = 0;
    CK(cudaMemcpyAsync(&ctx->hFlags, ctx->dFlags, sizeof(StatusFlags), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    if (ctx->hFlags.nanParticle) return fail(ctx, MPMGPU_ENAN, "particle %d left grid with position nan (ResetElementsTask)", ctx->hFlags.nanParticle - 1);
    if (ctx->hFlags.cpdiLeft) return fail(ctx, MPMGPU_ECPDI, "a CPDI corner of particle %d has left the grid", ctx->hFlags.cpdiLeft - 1);
    return MPMGPU_OK;
}

#define TASK_ENTRY(NAME, BODY, TASKID) \
extern "C" int NAME(mpmgpu_ctx *ctx) \
{ \
    int rc = check_ready(ctx, #NAME); if (rc) return rc; \
    prof_begin(ctx); rc = BODY(ctx); prof_end(ctx, TASKID); \
    if (rc) return rc; \
    CK(cudaGetLastError()); \
    return MPMGPU_OK; \
}

TASK_ENTRY(mpmgpu_task_initialization, t_initialization, T_INIT)
TASK_ENTRY(mpmgpu_task_mass_and_momentum, t_mass_and_momentum, T_MASSMOM)
TASK_ENTRY(mpmgpu_task_project_rigid_bcs, t_project_rigid_bcs, T_POSTEXTRAP)
TASK_ENTRY(mpmgpu_task_post_extrapolation, t_post_extrapolation, T_POSTEXTRAP)
TASK_ENTRY(mpmgpu_task_update_strains_first, t_update_strains_first, T_USF)
TASK_ENTRY(mpmgpu_task_grid_forces, t_grid_forces, T_FORCES)
TASK_ENTRY(mpmgpu_task_post_forces, t_post_forces, T_POSTFORCES)
TASK_ENTRY(mpmgpu_task_update_momenta, t_update_momenta, T_MOMENTA)
TASK_ENTRY(mpmgpu_task_update_particles, t_update_particles, T_PARTICLES)
TASK_ENTRY(mpmgpu_task_update_strains_last, t_update_strains_last, T_USL)

extern "C" int mpmgpu_task_reset_elements(mpmgpu_ctx *ctx)
{
    int rc = check_ready(ctx, "mpmgpu_task_reset_elements"); if (rc) return rc;
    prof_begin(ctx); rc = t_reset_elements(ctx); prof_end(ctx, T_RESET);
    if (rc) return rc;
    ctx->mstep++; ctx->mtime += ctx->sp.dt;
    return poll_flags(ctx);
}

static int step_by_tasks(mpmgpu_ctx *ctx)
{
    int rc;
    typedef int (*taskfn)(mpmgpu_ctx *);
    static const taskfn seq[T_NTASKS] = {t_initialization, t_mass_and_momentum, t_post_extrapolation, t_update_strains_first,
                                         t_grid_forces, t_post_forces, t_update_momenta, t_update_particles,
                                         t_update_strains_last, t_reset_elements};
    for (int t = 0; t < T_NTASKS; t++) {
        prof_begin(ctx);
        rc = seq[t](ctx);
        if (!rc && t == T_MASSMOM) rc = t_project_rigid_bcs(ctx);
        prof_end(ctx, t);
        if (rc) return rc;
    }
    return MPMGPU_OK;
}


// ------------------------------------------------------------------------------------------------
// fused fast path (kernels_fused.cuh)
static int sort_particles(mpmgpu_ctx *ctx)
{
    TiledState &t = ctx->tiled;
    const int n = ctx->P.n;
    if (n == 0) { t.stepsSinceSort = 0; return MPMGPU_OK; }        // an empty slab
    if (t.cap < ctx->cap) {
        if (t.cap != 0) return fail(ctx, MPMGPU_EINVAL, "sort workspace capacity changed");
        CK(dalloc(ctx, &t.keysIn, ctx->cap)); CK(dalloc(ctx, &t.keysOut, ctx->cap));
        CK(dalloc(ctx, &t.idxIn, ctx->cap)); CK(dalloc(ctx, &t.idxOut, ctx->cap));
        CK(dalloc(ctx, &t.altPool, ctx->cap * NPD)); CK(dalloc(ctx, &t.altIntPool, ctx->cap * NPI));
        CK(cudaMemsetAsync(t.altPool, 0, ctx->cap * NPD * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(t.altIntPool, 0, ctx->cap * NPI * sizeof(int), ctx->stream));
        t.cubTempBytes = 0;
        cub::DeviceRadixSort::SortPairs(NULL, t.cubTempBytes, t.keysIn, t.keysOut, t.idxIn, t.idxOut, (int)ctx->cap, 0, 32, ctx->stream);
        CK(dalloc(ctx, (char **)&t.cubTemp, t.cubTempBytes));
        t.cap = ctx->cap;
    }
    const char *leadEnv = getenv("MPMGPU_SORT_LEAD");        // fraction of the sort interval to look ahead (default 0.5)
    const double lead = (leadEnv ? atof(leadEnv) : 0.5) * t.sortInterval * ctx->sp.dt;
    LAUNCH(k_sort_keys, nblocks(n, 256), 256, ctx->g, ctx->P, t.keysIn, t.idxIn, lead);
    int bits = 1;
    while ((1ll << bits) < (long long)ctx->g.nnodes + (n - ctx->P.nNR) + 1) bits++;
    CK(cub::DeviceRadixSort::SortPairs(t.cubTemp, t.cubTempBytes, t.keysIn, t.keysOut, t.idxIn, t.idxOut, n, 0, bits, ctx->stream));
    ctx->launches += 4;
    // field order = bind_particles: pos 0-2 vel 3-5 mp 6 lp 7-9 ncpos 10-12 F 13-21 sp 22-27 pressure 28 eplast 29-34 work 35 res 36
    // heat 37 entropy 38 plast 39 prevT 40 hist 41-44 pfext 45-47 acc 48-50 | ints: elem mat cross orig key.
    // ncpos, key (F1 rewrites them right after the sort) and acc (F3 writes it before anyone reads it) never move; with
    // IsotropicMat only, eplast, the plastic energy and the history are zero in both pools; likewise pfext without particle loads.
    static_assert(MPM_MAX_HISTORY == 4 && NPD == 51, "the field positions below follow bind_particles with 4 history variables");
    unsigned long long live = (1ull << NPD) - 1ull;
    live &= ~(7ull << 10); live &= ~(7ull << 48);
    if (t.stateKind == SK_ELASTIC) { live &= ~(63ull << 29); live &= ~(1ull << 39); live &= ~(15ull << 41); }
    if (!ctx->hasFext) live &= ~(7ull << 45);
    const unsigned ilive = 0xfu;
    LAUNCH(k_permute_pool, nblocks(n, 256), 256, n, ctx->cap, NPD, live, ctx->particlePool, t.altPool, NPI, ilive, ctx->particleIntPool, t.altIntPool, t.idxOut);
    std::swap(ctx->particlePool, t.altPool);
    std::swap(ctx->particleIntPool, t.altIntPool);
    int nn = ctx->P.n, nnr = ctx->P.nNR;
    bind_particles(ctx->P, ctx->particlePool, ctx->particleIntPool, ctx->cap);
    ctx->P.n = nn; ctx->P.nNR = nnr;
    ctx->P.cpElem = ctx->cpElemPool; ctx->P.cpXi = ctx->cpXiPool; ctx->P.cpWg = ctx->cpWgPool; ctx->P.cpStride = ctx->cap;
    ctx->P.cpDom = ctx->cpXiPool ? ctx->cpXiPool + ctx->cpDomOffset * ctx->cap : NULL;
    t.stepsSinceSort = 0;
    return MPMGPU_OK;
}

__global__ void k_zero_node_range(int n0, int count, Nodes N)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + count) return;
    N.mass[i] = 0.; N.cnt[i] = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) { N.pk[c][i] = 0.; N.ftot[c][i] = 0.; N.vk[c][i] = 0.; N.pkc[c][i] = 0.; }
}

__global__ void k_rezero_range(int n0, int count, Nodes N)      // NodalPoint::RezeroNodeTask6 (NodalPointMPM.cpp:725-730)
{
    const int i = n0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0 + count) return;
    N.pk[0][i] = 0.; N.pk[1][i] = 0.; N.pk[2][i] = 0.;
}

#define MIG_ROW (NPD + (NPI + 1) / 2)

// halo planes shared with the lower (side 0) / upper (side 1) neighbour: three node planes around the slab face
static inline void halo_range(const mpmgpu_ctx *ctx, int side, int &node0, int &count)
{
    const TiledState &t = ctx->tiled;
    const int face = side == 0 ? t.slab.cellLo : t.slab.cellHi;
    node0 = (face - 1) * ctx->g.zplane;
    count = 3 * ctx->g.zplane;
}

static int halo_pack(mpmgpu_ctx *ctx, int which, bool real)
{
    TiledState &t = ctx->tiled;
    const int nv = which == 0 ? 4 : 3;
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? t.hasLower : t.hasUpper)) continue;
        int node0, count;
        halo_range(ctx, side, node0, count);
        if (real) LAUNCH(k_halo_pack, nblocks(count, 256), 256, which, node0, count, ctx->N, t.haloSend[side]);
        else { CK(cudaMemsetAsync(t.haloSend[side], 0, (size_t)nv * count * sizeof(double), ctx->stream)); ctx->launches++; }
    }
    return MPMGPU_OK;
}

static int halo_add(mpmgpu_ctx *ctx, int which)
{
    TiledState &t = ctx->tiled;
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? t.hasLower : t.hasUpper)) continue;
        int node0, count;
        halo_range(ctx, side, node0, count);
        LAUNCH(k_halo_add, nblocks(count, 256), 256, which, node0, count, ctx->N, t.haloRecv[side]);
    }
    return MPMGPU_OK;
}

// XPICExtrapolationTask in fused form (kernels_fused.cuh): grid velocity v*(m) for order m > 1 into the V records
static int xpic_fused(mpmgpu_ctx *ctx, int particleUpdate)
{
    TiledState &t = ctx->tiled;
    const int n0 = t.slab.on ? t.nodeLo : 0, ncount = t.slab.on ? t.nodeCount : ctx->g.nnodes, ngrid = nblocks(ncount, 256);
    const int pgrid = nblocks(ctx->P.nNR, FUSED_THREADS);
    const int fmpm = ctx->sp.usingFMPM ? 1 : 0;
    const bool exchange = t.slab.on && (t.hasLower || t.hasUpper);
    int rc;
    LAUNCH(k_nx_init, ngrid, 256, n0, ncount, ctx->N, t.FN, ctx->sp.dt, fmpm);
    for (int k = 2; k <= ctx->sp.xpicOrder; k++) {
        if (pgrid) LAUNCH(k_fx_iterate, pgrid, FUSED_THREADS, ctx->g, ctx->P, ctx->N, t.FN);
        if (exchange) {         // partial v*next sums on the planes around the slab faces: swap with the neighbours and add
            if ((rc = halo_pack(ctx, 3, true))) return rc;
            ctx->haloFn(ctx->haloUser, 3);
            if ((rc = halo_add(ctx, 3))) return rc;
        }
        LAUNCH(k_nx_finish, ngrid, 256, n0, ncount, ctx->N, t.FN, ctx->B, ctx->sp.dt, particleUpdate, fmpm, k == ctx->sp.xpicOrder ? 1 : 0);
    }
    return MPMGPU_OK;
}

// One step = four phases; in slab mode the host exchanges halo buffers between them.
//  0: (sort) zero nodes, F1                      -> partial mass/momentum sums
//  1: N1, F2                                     -> partial forces
//  2: N2, F3                                     -> partial re-extrapolated momenta
//  3: N3, F4                                     -> particles updated, leavers listed
static int fused_phase(mpmgpu_ctx *ctx, int phase)
{
    TiledState &t = ctx->tiled;
    const Grid &g = ctx->g;
    const StepParams &sp = ctx->sp;
    int rc;
    const int n = ctx->P.n, nNR = ctx->P.nNR;
    const int pgrid = nblocks(nNR, FUSED_THREADS);
    const int n0 = t.slab.on ? t.nodeLo : 0, ncount = t.slab.on ? t.nodeCount : g.nnodes;
    const int ngrid = nblocks(ncount, 256);
    const bool hasUSF = sp.method == METHOD_USF || sp.method == METHOD_USAVG;
    const bool hasUSL = sp.method == METHOD_USL || sp.method == METHOD_USAVG;
    const bool reextrap = hasUSL && !sp.skipPost;
    const double stFirst = sp.method == METHOD_USAVG ? sp.dtStrainFirst : sp.dt;
    const double stLast = sp.method == METHOD_USAVG ? sp.dtStrainLast : sp.dt;
    int m = sp.xpicOrder;
    if (!sp.usingFMPM) m = -m;
    const bool highOrder = sp.xpicOrder > 1;

    if (phase == 0) {
        if (!ctx->inGraphStep) {        // (a step that is captured or replayed as a graph sorts and counts in mpmgpu_step)
            if (t.stepsSinceSort >= t.sortInterval) { if ((rc = sort_particles(ctx))) return rc; }
            t.stepsSinceSort++;
        }
        prof_begin(ctx);
        if (t.slab.on) LAUNCH(k_zero_node_range, ngrid, 256, n0, ncount, ctx->N);
        else {
            const size_t nnPad = ((size_t)g.nnodes + 31) & ~(size_t)31;
            CK(cudaMemsetAsync(ctx->nodePool, 0, nnPad * 13 * sizeof(double), ctx->stream));      // mass, pk, ftot, vk, pkc
            ctx->launches += 1;
        }
        prof_end(ctx, T_INIT);
        prof_begin(ctx);
        if (pgrid) LAUNCH(k_f1_mass_momentum, pgrid, FUSED_THREADS, g, ctx->P, ctx->N);
        if ((rc = t_project_rigid_bcs(ctx))) return rc;
        if (t.slab.on && (rc = halo_pack(ctx, 0, true))) return rc;
        prof_end(ctx, T_MASSMOM);
    } else if (phase == 1) {
        prof_begin(ctx);
        if (t.slab.on && (rc = halo_add(ctx, 0))) return rc;
        LAUNCH(k_n1_post_extrapolation, ngrid, 256, n0, ncount, ctx->N, t.FN, ctx->B, sp, hasUSF ? 1 : 0);
        // FMPM(k>1) strain updates use v*(k) (UpdateStrainsFirstTask.cpp:105-116)
        if (highOrder && sp.usingFMPM && hasUSF && (rc = xpic_fused(ctx, 0))) return rc;
        prof_end(ctx, T_POSTEXTRAP);
        prof_begin(ctx);
        if (pgrid) {
#define LAUNCH_F2(SKV, FX) do { \
                if (!ctx->f2Attr[SKV][FX ? 1 : 0]) { CK(cudaFuncSetAttribute(k_f2_strain_forces<SKV, FX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2_smem_bytes<FX>())); ctx->f2Attr[SKV][FX ? 1 : 0] = true; } \
                k_f2_strain_forces<SKV, FX><<<pgrid, FUSED_THREADS, f2_smem_bytes<FX>(), ctx->stream>>>(g, ctx->P, ctx->N, t.FN, ctx->dMats, stFirst, hasUSF ? 1 : 0); \
                ctx->launches++; } while (0)
            if (t.stateKind == SK_ELASTIC) {
                if (ctx->hasFext) LAUNCH_F2(SK_ELASTIC, true); else LAUNCH_F2(SK_ELASTIC, false);
            } else {
                if (ctx->hasFext) LAUNCH_F2(SK_FULL, true); else LAUNCH_F2(SK_FULL, false);
            }
        }
        if (t.slab.on && (rc = halo_pack(ctx, 1, true))) return rc;
        prof_end(ctx, T_USF);
    } else if (phase == 2) {
        prof_begin(ctx);
        if (t.slab.on && (rc = halo_add(ctx, 1))) return rc;
        // (with order > 1 the re-zeroing of pk for task 9a waits until v* has been formed from it)
        if ((rc = particle_tractions(ctx))) return rc;
        if ((rc = reactions_zero(ctx))) return rc;
        LAUNCH(k_n2_forces_momenta, ngrid, 256, n0, ncount, ctx->N, t.FN, ctx->B, sp, (reextrap && !highOrder) ? 1 : 0);
        if (highOrder) {                                       // UpdateParticlesTask.cpp:66-71
            if ((rc = xpic_fused(ctx, 1))) return rc;
            if (reextrap) LAUNCH(k_rezero_range, ngrid, 256, n0, ncount, ctx->N);
        }
        prof_end(ctx, T_POSTFORCES);
        prof_begin(ctx);
        if (pgrid) LAUNCH(k_f3_update_momentum, pgrid, FUSED_THREADS, g, ctx->P, ctx->N, t.FN, ctx->dMats, sp, m, reextrap ? 1 : 0);
        if ((rc = move_rigid(ctx))) return rc;
        if (t.slab.on && (rc = halo_pack(ctx, 2, reextrap))) return rc;
        prof_end(ctx, T_PARTICLES);
    } else {
        prof_begin(ctx);
        if (t.slab.on && (rc = halo_add(ctx, 2))) return rc;
        if (reextrap) LAUNCH(k_n3_strains_last, ngrid, 256, n0, ncount, ctx->N, t.FN, ctx->B, sp);
        if (highOrder && hasUSL) {
            if (sp.usingFMPM) { if (reextrap && (rc = xpic_fused(ctx, 0))) return rc; }        // second strain update on v*(k) of the new momenta
            else if (!reextrap) LAUNCH(k_n_grid_velocity, ngrid, 256, n0, ncount, ctx->N, t.FN);   // XPIC: lumped velocity again
        }
        const bool earlyReset = t.slab.on && !t.usePipe && n > 0;
        if (earlyReset) {
            LAUNCH(k_reset_slab, nblocks(n, 256), 256, g, ctx->P, ctx->dFlags, sp.dt, t.slab);
            CK(cudaMemcpyAsync(ctx->slabHost->leave, t.slab.leaveCount, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(&ctx->slabHost->flags, ctx->dFlags, sizeof(StatusFlags), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaEventRecord(ctx->slabEvent, ctx->stream));
            ctx->slabPending = true;
        }
        if (n && t.usePipe) {
            PipeFields pf;
            memset(&pf, 0, sizeof pf);
            int k = 0;
            for (int c = 0; c < 3; c++) pf.d[k++] = ctx->P.ncpos[c];
            for (int c = 0; c < 9; c++) pf.d[k++] = ctx->P.F[c];
            for (int c = 0; c < 6; c++) pf.d[k++] = ctx->P.sp[c];
            pf.d[k++] = ctx->P.work; pf.d[k++] = ctx->P.heat; pf.d[k++] = ctx->P.entropy; pf.d[k++] = ctx->P.prevT;
            for (int c = 0; c < 3; c++) pf.d[k++] = ctx->P.pos[c];
            if (t.stateKind == SK_FULL) {
                for (int c = 0; c < 6; c++) pf.d[k++] = ctx->P.eplast[c];
                pf.d[k++] = ctx->P.pressure; pf.d[k++] = ctx->P.plast; pf.d[k++] = ctx->P.res;
                for (int c = 0; c < MPM_MAX_HISTORY; c++) pf.d[k++] = ctx->P.hist[c];
            }
            pf.nd = k; pf.i[0] = ctx->P.elem; pf.i[1] = ctx->P.mat; pf.ni = 2;
            const size_t smemBytes = (size_t)FUSED_WARPS * PIPE_STAGES * pipe_stage_bytes(pf.nd, pf.ni) + FUSED_WARPS * PIPE_STAGES * sizeof(unsigned long long);
            const int nchunks = (n + 31) / 32;
            int blocksPerSM = (int)(220 * 1024 / smemBytes); if (blocksPerSM > 4) blocksPerSM = 4; if (blocksPerSM < 1) blocksPerSM = 1;
            int grid = std::min((nchunks + FUSED_WARPS - 1) / FUSED_WARPS, t.numSMs * blocksPerSM);
            if (t.stateKind == SK_ELASTIC) {
                CK(cudaFuncSetAttribute(k_f4_pipe<SK_ELASTIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
                k_f4_pipe<SK_ELASTIC><<<grid, FUSED_THREADS, smemBytes, ctx->stream>>>(g, ctx->P, t.FN, ctx->dMats, pf, stLast, hasUSL ? 1 : 0, ctx->dFlags, sp.dt, t.slab);
            } else {
                CK(cudaFuncSetAttribute(k_f4_pipe<SK_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
                k_f4_pipe<SK_FULL><<<grid, FUSED_THREADS, smemBytes, ctx->stream>>>(g, ctx->P, t.FN, ctx->dMats, pf, stLast, hasUSL ? 1 : 0, ctx->dFlags, sp.dt, t.slab);
            }
            ctx->launches++;
        } else if (n) {
            // slab mode: the reset ran as its own kernel before this one (see k_reset_slab) unless there is no strain update to hide behind
            const int doReset = earlyReset ? 0 : 1;
            if (!doReset && !hasUSL) { /* nothing left to do */ }
            else if (t.stateKind == SK_ELASTIC) LAUNCH(k_f4_strain_reset<SK_ELASTIC>, nblocks(n, FUSED_THREADS), FUSED_THREADS, g, ctx->P, t.FN, ctx->dMats, stLast, hasUSL ? 1 : 0, doReset, ctx->dFlags, sp.dt, t.slab);
            else LAUNCH(k_f4_strain_reset<SK_FULL>, nblocks(n, FUSED_THREADS), FUSED_THREADS, g, ctx->P, t.FN, ctx->dMats, stLast, hasUSL ? 1 : 0, doReset, ctx->dFlags, sp.dt, t.slab);
        }
        if ((rc = reset_rigid(ctx))) return rc;
        prof_end(ctx, T_USL);
    }
    return MPMGPU_OK;
}

static int fused_step(mpmgpu_ctx *ctx)
{
    int rc;
    for (int ph = 0; ph < 4; ph++) if ((rc = fused_phase(ctx, ph))) return rc;
    return MPMGPU_OK;
}

// everything one step's launches take by value: when these bytes are the same, a captured step can be replayed
static void step_signature(const mpmgpu_ctx *ctx, std::string &sig)
{
    sig.clear();
    auto add = [&](const void *p, size_t n) { sig.append((const char *)p, n); };
    add(&ctx->g, sizeof ctx->g); add(&ctx->P, sizeof ctx->P); add(&ctx->PR, sizeof ctx->PR); add(&ctx->N, sizeof ctx->N);
    add(&ctx->B, sizeof ctx->B); add(&ctx->R, sizeof ctx->R); add(&ctx->sp, sizeof ctx->sp); add(&ctx->tiled.FN, sizeof ctx->tiled.FN);
    add(&ctx->C, sizeof ctx->C); add(&ctx->cp, sizeof ctx->cp); add(&ctx->T, sizeof ctx->T); add(&ctx->Q, sizeof ctx->Q);
    add(&ctx->trac.TB, sizeof ctx->trac.TB); add(&ctx->flux.TB, sizeof ctx->flux.TB);
    const long long more[3] = {ctx->trackReactions, ctx->nBCEntries, ctx->nmat};        // (sizes of the memset nodes of reactions_zero)
    add(more, sizeof more);
    const long long misc[17] = {ctx->thermal, ctx->tiled.enabled, ctx->tiled.stateKind, ctx->tiled.usePipe, ctx->hasFext, ctx->hasBCs, ctx->largeRotation, ctx->multimaterial,
                                ctx->conduction, ctx->nf, ctx->nvn, ctx->cpdiMerge, ctx->cpdiMergeValues, (long long)(size_t)ctx->dMats, (long long)(size_t)ctx->archOrigin,
                                (long long)(size_t)ctx->nodePool, (long long)(size_t)ctx->stream};
    add(misc, sizeof misc);
}

static void drop_step_graphs(mpmgpu_ctx *ctx)
{
    for (auto &sg : ctx->stepGraphs) cudaGraphExecDestroy(sg.exec);
    ctx->stepGraphs.clear();
}

// the sort and the step count of this step are already done: run the phases as a graph step would, without a graph
static int graph_step_plain(mpmgpu_ctx *ctx, bool &done)
{
    ctx->inGraphStep = true;
    const int rc = ctx->tiled.enabled ? fused_step(ctx) : step_by_tasks(ctx);
    ctx->inGraphStep = false;
    done = true;
    return rc;
}

// one step as a CUDA graph: captured the first time a signature is seen (the capture only records, so the graph is launched
// for that step too), replayed afterwards -- one launch instead of 10-30, which is what a small problem's step costs
static int graph_step(mpmgpu_ctx *ctx, bool &done)
{
    done = false;
    int rc;
    TiledState &t = ctx->tiled;
    if (t.enabled) {            // the host-side part of phase 0
        if (t.stepsSinceSort >= t.sortInterval) { if ((rc = sort_particles(ctx))) return rc; }
        t.stepsSinceSort++;
        if (t.usePipe) return graph_step_plain(ctx, done);          // (sets a function attribute per launch: not captured)
        // the dynamic shared-memory opt-in of F2 is not a stream operation: do it before capturing
        const int fx = ctx->hasFext ? 1 : 0;
        if (!ctx->f2Attr[t.stateKind][fx]) {
            cudaError_t e;
            if (t.stateKind == SK_ELASTIC) e = fx ? cudaFuncSetAttribute(k_f2_strain_forces<SK_ELASTIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2_smem_bytes<true>())
                                                  : cudaFuncSetAttribute(k_f2_strain_forces<SK_ELASTIC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2_smem_bytes<false>());
            else e = fx ? cudaFuncSetAttribute(k_f2_strain_forces<SK_FULL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2_smem_bytes<true>())
                        : cudaFuncSetAttribute(k_f2_strain_forces<SK_FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f2_smem_bytes<false>());
            CK(e);
            ctx->f2Attr[t.stateKind][fx] = true;
        }
    }
    std::string sig;
    step_signature(ctx, sig);
    mpmgpu_ctx::StepGraph *hit = NULL;
    for (auto &sg : ctx->stepGraphs) if (sg.sig == sig) { hit = &sg; break; }
    if (!hit) {
        if (ctx->stepGraphs.size() >= 4) { cudaGraphExecDestroy(ctx->stepGraphs.front().exec); ctx->stepGraphs.erase(ctx->stepGraphs.begin()); }
        const long long before = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); ctx->useGraphs = false; return graph_step_plain(ctx, done); }
        ctx->inGraphStep = true;
        rc = t.enabled ? fused_step(ctx) : step_by_tasks(ctx);
        ctx->inGraphStep = false;
        cudaGraph_t graph = NULL;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        const long long captured = ctx->launches - before;
        ctx->launches = before;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        cudaGraphExec_t exec = NULL;
        if (e != cudaSuccess || graph == NULL || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            // something in this configuration cannot be captured: run it the plain way from now on
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            ctx->useGraphs = false;
            return graph_step_plain(ctx, done);
        }
        cudaGraphDestroy(graph);
        ctx->stepGraphs.push_back({sig, exec, captured});
        hit = &ctx->stepGraphs.back();
    }
    CK(cudaGraphLaunch(hit->exec, ctx->stream));
    ctx->launches += hit->launches;
    done = true;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_step(mpmgpu_ctx *ctx, int nsteps)
{
    int rc = check_ready(ctx, "mpmgpu_step"); if (rc) return rc;
    if (ctx->tiled.slab.on && (ctx->tiled.hasLower || ctx->tiled.hasUpper))
        return fail(ctx, MPMGPU_ESTATE, "mpmgpu_step: this context is one slab of a multi-GPU run; drive it with mpmgpu_slab_step_phase and the halo exchanges");
    for (int s = 0; s < nsteps; s++) {
        bool done = false;
        if (ctx->useGraphs && !ctx->profiling && !ctx->tiled.slab.on) { if ((rc = graph_step(ctx, done))) return rc; }
        if (!done) {
            rc = ctx->tiled.enabled ? fused_step(ctx) : step_by_tasks(ctx);
            if (rc) return rc;
        }
        ctx->mstep++; ctx->mtime += ctx->sp.dt;
    }
    return poll_flags(ctx, false);
}

extern "C" int mpmgpu_set_poll_interval(mpmgpu_ctx *ctx, int k)
{
    if (!ctx || k < 1) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_set_poll_interval: interval must be >= 1");
    ctx->pollInterval = k; ctx->callsSincePoll = 0;
    return MPMGPU_OK;
}

// ------------------------------------------------------------------------------------------------
static int down_field(mpmgpu_ctx *ctx, double *const *dev, double *const *devR, double *host, int ncomp, int n, double *dtmp)
{
    if (!host) return MPMGPU_OK;
    const int T = 256;
    const int nNR = ctx->P.n, nR = ctx->PR.n;
    for (int c = 0; c < ncomp; c++) {
        if (nNR) LAUNCH(k_unpermute, nblocks(nNR, T), T, nNR, dev[c], ctx->dlSlot, dtmp);
        if (nR) LAUNCH(k_unpermute, nblocks(nR, T), T, nR, devR[c], ctx->dlSlotR, dtmp);
        CK(cudaMemcpyAsync(host + (size_t)c * n, dtmp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return MPMGPU_OK;
}

extern "C" int mpmgpu_download_particles(mpmgpu_ctx *ctx, mpmgpu_particles *h, unsigned mask)
{
    if (!ctx || !h) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_download_particles: null argument");
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_download_particles: nothing uploaded");
    cudaSetDevice(ctx->cfg.device);
    Particles &P = ctx->P, &PR = ctx->PR;
    const int nNR = P.n, nR = PR.n, n = nNR + nR;
    h->n = n; h->n_nonrigid = nNR;
    double *dtmp = NULL;
    int rc = stage_buffer(ctx, (size_t)n * 10, &dtmp);        // 9 doubles per particle of field staging + the identity map of slab mode
    if (rc) return rc;
    const int T = 256;
    int *ident = NULL;
    ctx->dlSlot = P.orig; ctx->dlSlotR = PR.orig;
    if (ctx->globalIds) {       // device order (rigid particles last); the caller re-assembles by id
        ident = reinterpret_cast<int *>(dtmp + (size_t)n * 9);
        LAUNCH(k_iota, nblocks(n, T), T, n, ident, 0);
        ctx->dlSlot = ident; ctx->dlSlotR = ident + nNR;
        if (h->ids) {
            if (nNR) CK(cudaMemcpyAsync(h->ids, P.orig, (size_t)nNR * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            if (nR) CK(cudaMemcpyAsync(h->ids + nNR, PR.orig, (size_t)nR * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    do {
        if ((mask & MPMGPU_F_POS) && (rc = down_field(ctx, P.pos, PR.pos, h->pos, 3, n, dtmp))) break;
        if ((mask & MPMGPU_F_VEL) && (rc = down_field(ctx, P.vel, PR.vel, h->vel, 3, n, dtmp))) break;
        if (mask & MPMGPU_F_STRESS) {
            if ((rc = down_field(ctx, P.sp, PR.sp, h->sp, 6, n, dtmp))) break;
            if ((rc = down_field(ctx, &P.pressure, &PR.pressure, h->pressure, 1, n, dtmp))) break;
        }
        if ((mask & MPMGPU_F_STRAIN) && h->ep && h->wrot) {
            if (nNR) LAUNCH(k_F_to_epwrot, nblocks(nNR, T), T, nNR, n, ctx->dim, P, ctx->dlSlot, dtmp, dtmp + (size_t)6 * n);
            if (nR) LAUNCH(k_F_to_epwrot, nblocks(nR, T), T, nR, n, ctx->dim, PR, ctx->dlSlotR, dtmp, dtmp + (size_t)6 * n);
            CK(cudaMemcpyAsync(h->ep, dtmp, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(h->wrot, dtmp + (size_t)6 * n, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if ((mask & MPMGPU_F_EPLAST) && (rc = down_field(ctx, P.eplast, PR.eplast, h->eplast, 6, n, dtmp))) break;
        if ((mask & MPMGPU_F_ENERGY) && h->energies) {
            double *const e6[6] = {P.work, P.res, P.heat, P.entropy, P.plast, P.prevT};
            double *const e6R[6] = {PR.work, PR.res, PR.heat, PR.entropy, PR.plast, PR.prevT};
            if ((rc = down_field(ctx, e6, e6R, h->energies, 6, n, dtmp))) break;
        }
        if ((mask & MPMGPU_F_HISTORY) && (rc = down_field(ctx, P.hist, PR.hist, h->history, MPM_MAX_HISTORY, n, dtmp))) break;
        if ((mask & MPMGPU_F_ACC) && (rc = down_field(ctx, P.acc, PR.acc, h->acc, 3, n, dtmp))) break;
        if ((mask & MPMGPU_F_TEMPERATURE) && h->temperature && ctx->thermal) {
            // pTemperature of the nonrigid particles (rigid-BC particles: the temperature of energies[5])
            double *const t1[1] = {P.temp}, *const t1R[1] = {ctx->rigidTemp ? PR.temp : PR.prevT};
            if ((rc = down_field(ctx, t1, t1R, h->temperature, 1, n, dtmp))) break;
        }
        if (mask & MPMGPU_F_ELEM) {
            int *itmp = (int *)dtmp;
            if (h->in_elem) {
                if (nNR) LAUNCH(k_unpermute_int, nblocks(nNR, T), T, nNR, P.elem, ctx->dlSlot, itmp, 0);
                if (nR) LAUNCH(k_unpermute_int, nblocks(nR, T), T, nR, PR.elem, ctx->dlSlotR, itmp, 0);
                CK(cudaMemcpyAsync(h->in_elem, itmp, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            }
            if (h->crossings) {
                if (nNR) LAUNCH(k_unpermute_int, nblocks(nNR, T), T, nNR, P.cross, ctx->dlSlot, itmp + n, 0);
                if (nR) LAUNCH(k_unpermute_int, nblocks(nR, T), T, nR, PR.cross, ctx->dlSlotR, itmp + n, 0);
                CK(cudaMemcpyAsync(h->crossings, itmp + n, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            }
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(ctx, MPMGPU_ECUDA, "mpmgpu_download_particles: %s", cudaGetErrorString(e));
    return MPMGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// output side: archive records and global sums on the device (archive.cuh)
extern "C" int mpmgpu_archive_record_size(const mpmgpu_ctx *ctx, const char *order)
{
    if (!ctx || !order) return -1;
    ArchiveLayout L;
    return archive_layout_from_order(order, ctx->dim, L);
}

extern "C" int mpmgpu_set_archive_origin(mpmgpu_ctx *ctx, const double *origpos, const double *angles0, double thickness)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_archive_origin: upload the particles first");
    if (ctx->globalIds) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_set_archive_origin: not available in slab mode (each rank downloads its particles)");
    cudaSetDevice(ctx->cfg.device);
    const size_t n = (size_t)ctx->P.n + (size_t)ctx->PR.n;
    if (origpos) {
        if (ctx->archOrigin && ctx->archOriginLen < 3 * n) ctx->archOrigin = NULL;
        if (!ctx->archOrigin) { CK(dalloc(ctx, &ctx->archOrigin, 3 * n)); ctx->archOriginLen = 3 * n; }
        CK(cudaMemcpyAsync(ctx->archOrigin, origpos, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->archOriginFromCaller = true;
    }
    if (angles0) {
        if (!ctx->archAngles) CK(dalloc(ctx, &ctx->archAngles, 3 * n));
        CK(cudaMemcpyAsync(ctx->archAngles, angles0, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->archThickness = thickness;
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_pack_archive(mpmgpu_ctx *ctx, const char *order, void *records, size_t capacity_bytes)
{
    if (!ctx || !order || !records) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_pack_archive: null argument");
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_pack_archive: nothing uploaded");
    cudaSetDevice(ctx->cfg.device);
    ArchiveLayout L;
    const int recBytes = archive_layout_from_order(order, ctx->dim, L);
    if (recBytes < 0) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_pack_archive: <MPMArchiveOrder> \"%s\" asks for an item this path does not produce (shear components, damage normal, spin, history 5-19, size)", order);
    const int nNR = ctx->P.n, nR = ctx->PR.n;
    const size_t n = (size_t)nNR + (size_t)nR, words = n * (size_t)L.recWords;
    if (capacity_bytes < words * 4) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_pack_archive: %zu bytes needed (%zu particles x %d), %zu given", words * 4, n, recBytes, capacity_bytes);
    if (ctx->archBufWords < words) {
        CK(dalloc(ctx, &ctx->archBuf, words));       // earlier (smaller) buffers stay on the context's free list until destroy
        ctx->archBufWords = words;
    }
    L.thickness = ctx->archThickness; L.origpos = ctx->archOrigin; L.angles0 = ctx->archAngles; L.stride = n;
    if (ctx->globalIds) {
        // one slab of a multi-GPU run: records in this rank's device order (mpmgpu_download_ids gives the particle ids in the same
        // order); the constants indexed by the caller's particle order are not on this rank: "original position" = the current one
        L.origpos = NULL; L.angles0 = NULL;
        if (nNR) LAUNCH(k_pack_archive, nblocks(nNR, ARCHIVE_THREADS), ARCHIVE_THREADS, nNR, ctx->P, (const int *)NULL, 0, ctx->dMats, L, ctx->archBuf);
        if (nR) LAUNCH(k_pack_archive, nblocks(nR, ARCHIVE_THREADS), ARCHIVE_THREADS, nR, ctx->PR, (const int *)NULL, nNR, ctx->dMats, L, ctx->archBuf);
    } else {
        if (nNR) LAUNCH(k_pack_archive, nblocks(nNR, ARCHIVE_THREADS), ARCHIVE_THREADS, nNR, ctx->P, ctx->P.orig, 0, ctx->dMats, L, ctx->archBuf);
        if (nR) LAUNCH(k_pack_archive, nblocks(nR, ARCHIVE_THREADS), ARCHIVE_THREADS, nR, ctx->PR, ctx->PR.orig, 0, ctx->dMats, L, ctx->archBuf);
    }
    CK(cudaMemcpyAsync(records, ctx->archBuf, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

// particle ids in device order (slab mode: the global ids handed over at upload; otherwise the caller's particle index)
extern "C" int mpmgpu_download_ids(mpmgpu_ctx *ctx, int *ids, int capacity)
{
    if (!ctx || !ids) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_download_ids: null argument");
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_download_ids: nothing uploaded");
    const int nNR = ctx->P.n, nR = ctx->PR.n;
    if (capacity < nNR + nR) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_download_ids: %d ids, room for %d", nNR + nR, capacity);
    cudaSetDevice(ctx->cfg.device);
    if (nNR) CK(cudaMemcpyAsync(ids, ctx->P.orig, (size_t)nNR * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (nR) CK(cudaMemcpyAsync(ids + nNR, ctx->PR.orig, (size_t)nR * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_global_sums(mpmgpu_ctx *ctx, double *sums)
{
    if (!ctx || !sums) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_global_sums: null argument");
    if (!ctx->uploaded) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_global_sums: nothing uploaded");
    cudaSetDevice(ctx->cfg.device);
    const int nNR = ctx->P.n, nmat = ctx->nmat;
    int nblk = nblocks(nNR, ARCHIVE_THREADS);
    if (nblk > 592) nblk = 592;                       // 4 blocks per SM of a B200; grid-stride beyond that
    if (nblk < 1) nblk = 1;
    const size_t need = ((size_t)nblk + 1) * nmat * GS_NSUMS;
    if (ctx->gsumBufLen < need) { CK(dalloc(ctx, &ctx->gsumBuf, need)); ctx->gsumBufLen = need; }
    double *partial = ctx->gsumBuf, *out = ctx->gsumBuf + (size_t)nblk * nmat * GS_NSUMS;
    Particles P = ctx->P;
    P.nNR = nNR;
    {
        dim3 grid(nblk, nmat);
        k_global_partial<<<grid, ARCHIVE_THREADS, 0, ctx->stream>>>(P, ctx->dMats, ctx->dim, partial);
        ctx->launches++;
    }
    LAUNCH(k_global_final, nblocks(nmat * GS_NSUMS, 128), 128, nmat, nblk, partial, out);
    CK(cudaMemcpyAsync(sums, out, (size_t)nmat * GS_NSUMS * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_download_nodes(mpmgpu_ctx *ctx, mpmgpu_nodes *h)
{
    if (!ctx || !h) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_download_nodes: null argument");
    cudaSetDevice(ctx->cfg.device);
    const size_t nn = ctx->nvn;          // fields x nodes in multimaterial mode (field-major)
    h->nnodes = (int)nn;
    if (ctx->conduction) {
        const size_t nr = ctx->g.nnodes;
        if (h->transport_value) CK(cudaMemcpyAsync(h->transport_value, ctx->T.gT, nr * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h->transport_capacity) CK(cudaMemcpyAsync(h->transport_capacity, ctx->T.gVCT, nr * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h->transport_rate) CK(cudaMemcpyAsync(h->transport_rate, ctx->T.gQ, nr * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ctx->multimaterial) {
        if (h->contact_volume) CK(cudaMemcpyAsync(h->contact_volume, ctx->C.cvol, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        for (int c = 0; c < 3; c++) {
            if (h->contact_gradient) CK(cudaMemcpyAsync(h->contact_gradient + c * nn, ctx->C.cgrad[c], nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (h->contact_disp) CK(cudaMemcpyAsync(h->contact_disp + c * nn, ctx->C.cdisp[c], nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    if (h->number_points) CK(cudaMemcpyAsync(h->number_points, ctx->N.cnt, nn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> rcnt;
    if (h->number_points && ctx->multimaterial && ctx->cp.rigidMask) {
        rcnt.resize(nn);
        CK(cudaMemcpyAsync(rcnt.data(), ctx->C.rcnt, nn * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (h->mass) CK(cudaMemcpyAsync(h->mass, ctx->N.mass, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    for (int c = 0; c < 3; c++) {
        if (h->pk) CK(cudaMemcpyAsync(h->pk + c * nn, ctx->N.pk[c], nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h->ftot) CK(cudaMemcpyAsync(h->ftot + c * nn, ctx->N.ftot[c], nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h->vk) CK(cudaMemcpyAsync(h->vk + c * nn, ctx->N.vk[c], nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h->pk_copy) CK(cudaMemcpyAsync(h->pk_copy + c * nn, ctx->N.pkc[c], nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < rcnt.size(); i++) h->number_points[i] += rcnt[i];       // a rigid material's field counts its points apart
    if (h->ftot && ctx->multimaterial && ctx->cp.rigidMask) {
        // ... and its force row is the cumulative contact force (ContactNodes::rforce)
        const size_t nr = ctx->g.nnodes;
        for (int f = 0; f < ctx->nf; f++)
            if (ctx->cp.rigidMask >> f & 1)
                for (int c = 0; c < 3; c++) CK(cudaMemcpy(h->ftot + c * nn + f * nr, ctx->C.rforce[c] + f * nr, nr * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return MPMGPU_OK;
}

extern "C" int mpmgpu_synchronize(mpmgpu_ctx *ctx)
{
    if (!ctx) return MPMGPU_EINVAL;
    cudaSetDevice(ctx->cfg.device);
    CK(cudaStreamSynchronize(ctx->stream));
    return MPMGPU_OK;
}

extern "C" int mpmgpu_get_status(mpmgpu_ctx *ctx, long long *mstep, double *mtime, long long *crossings, long long *leftGrid)
{
    if (!ctx) return MPMGPU_EINVAL;
    cudaSetDevice(ctx->cfg.device);
    CK(cudaMemcpyAsync(&ctx->hFlags, ctx->dFlags, sizeof(StatusFlags), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (mstep) *mstep = ctx->mstep;
    if (mtime) *mtime = ctx->mtime;
    if (crossings) *crossings = (long long)ctx->hFlags.crossings;
    if (leftGrid) *leftGrid = (long long)(ctx->hFlags.leftGrid & 0xffffffffull);       // exits (the high word counts first-time leavers)
    return MPMGPU_OK;
}

extern "C" int mpmgpu_left_grid_counts(mpmgpu_ctx *ctx, long long *exits, long long *particles)
{
    if (!ctx) return MPMGPU_EINVAL;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->pollInterval <= 1 || ctx->callsSincePoll == 0) {      // (between polls: the counts of the last poll, no synchronisation)
        CK(cudaMemcpyAsync(&ctx->hFlags, ctx->dFlags, sizeof(StatusFlags), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (exits) *exits = (long long)(ctx->hFlags.leftGrid & 0xffffffffull);
    if (particles) *particles = (long long)(ctx->hFlags.leftGrid >> 32);
    return MPMGPU_OK;
}

extern "C" long long mpmgpu_launch_count(const mpmgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *mpmgpu_stream(mpmgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : NULL; }

extern "C" int mpmgpu_set_profiling(mpmgpu_ctx *ctx, int on)
{
    if (!ctx) return MPMGPU_EINVAL;
    ctx->profiling = on != 0;
    memset(ctx->taskMs, 0, sizeof ctx->taskMs); memset(ctx->taskCalls, 0, sizeof ctx->taskCalls);
    return MPMGPU_OK;
}

extern "C" int mpmgpu_task_times(mpmgpu_ctx *ctx, double *ms, long long *calls)
{
    if (!ctx || !ms) return MPMGPU_EINVAL;
    for (int t = 0; t < T_NTASKS; t++) { ms[t] = ctx->taskMs[t]; if (calls) calls[t] = ctx->taskCalls[t]; }
    return MPMGPU_OK;
}

// ------------------------------------------------------------------------------------------------
// slab decomposition across GPUs (one process per GPU; the host moves the buffers with NCCL)
extern "C" int mpmgpu_slab_configure(mpmgpu_ctx *ctx, int cell_lo, int cell_hi, int has_lower, int has_upper, int migration_capacity)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (ctx->dim != 3) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_configure: slabs are along z of a 3D grid");
    if (ctx->cfg.shape != MPMGPU_UNIFORM_GIMP || ctx->cfg.kernel_path == 1) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_configure: slab mode runs on the fused 3D uGIMP path");
    if (cell_lo < 0 || cell_hi > ctx->g.depth || cell_hi - cell_lo < 3) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_configure: slab [%d,%d) of %d cell planes (need >= 3)", cell_lo, cell_hi, ctx->g.depth);
    if ((has_lower && cell_lo < 2) || (has_upper && cell_hi > ctx->g.depth - 2)) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_configure: interior faces must be >= 2 planes from the grid edge");
    if (ctx->trackReactions) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_configure: reaction forces (mpmgpu_track_reactions) are not kept in slab mode");
    cudaSetDevice(ctx->cfg.device);
    TiledState &t = ctx->tiled;
    t.slab.on = 1; t.slab.cellLo = cell_lo; t.slab.cellHi = cell_hi;
    t.hasLower = has_lower ? 1 : 0; t.hasUpper = has_upper ? 1 : 0;
    const int planeLo = std::max(0, cell_lo - 1), planeHi = std::min(ctx->g.depth + 1, cell_hi + 2);
    t.nodeLo = planeLo * ctx->g.zplane; t.nodeCount = (planeHi - planeLo) * ctx->g.zplane;
    t.planeNodes = ctx->g.zplane;
    const size_t haloDoubles = (size_t)5 * 3 * ctx->g.zplane;
    t.migCap = migration_capacity > 0 ? migration_capacity : 65536;
    t.slab.leaveCap = t.migCap;
    for (int side = 0; side < 2; side++) {
        CK(dalloc(ctx, &t.haloSend[side], haloDoubles)); CK(dalloc(ctx, &t.haloRecv[side], haloDoubles));
        CK(dalloc(ctx, &t.migSend[side], (size_t)t.migCap * MIG_ROW)); CK(dalloc(ctx, &t.migRecv[side], (size_t)t.migCap * MIG_ROW));
        CK(cudaMemset(t.haloSend[side], 0, haloDoubles * sizeof(double))); CK(cudaMemset(t.haloRecv[side], 0, haloDoubles * sizeof(double)));
    }
    CK(dalloc(ctx, &t.slab.leaveCount, 2)); CK(dalloc(ctx, &t.slab.leaveIdx, (size_t)2 * t.migCap));
    CK(dalloc(ctx, &t.migSorted, (size_t)2 * t.migCap)); CK(dalloc(ctx, &t.migKeys, (size_t)2 * t.migCap));
    CK(dalloc(ctx, &t.migFillers, (size_t)2 * t.migCap)); CK(dalloc(ctx, &t.migFlags, (size_t)2 * t.migCap));
    CK(dalloc(ctx, &t.migPairs, 1));
    if (!ctx->slabHost) {
        CK(cudaMallocHost((void **)&ctx->slabHost, sizeof(*ctx->slabHost)));
        memset(ctx->slabHost, 0, sizeof(*ctx->slabHost));
        CK(cudaEventCreateWithFlags(&ctx->slabEvent, cudaEventDisableTiming));
    }
    t.migCubBytes = 0;
    cub::DeviceRadixSort::SortKeys(NULL, t.migCubBytes, t.migKeys, t.migSorted, 2 * t.migCap, 0, 32, ctx->stream);
    CK(dalloc(ctx, (char **)&t.migCubTemp, t.migCubBytes));
    CK(cudaMemset(t.slab.leaveCount, 0, 2 * sizeof(int)));
    t.hLeave[0] = t.hLeave[1] = 0;
    ctx->globalIds = true;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_slab_halo_buffers(mpmgpu_ctx *ctx, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, long long *plane_nodes)
{
    if (!ctx || !ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_halo_buffers: slab mode not configured");
    TiledState &t = ctx->tiled;
    if (send_lo) *send_lo = t.haloSend[0];
    if (send_hi) *send_hi = t.haloSend[1];
    if (recv_lo) *recv_lo = t.haloRecv[0];
    if (recv_hi) *recv_hi = t.haloRecv[1];
    if (plane_nodes) *plane_nodes = (long long)t.planeNodes;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_slab_set_halo_callback(mpmgpu_ctx *ctx, mpmgpu_halo_fn fn, void *user)
{
    if (!ctx) return MPMGPU_EINVAL;
    ctx->haloFn = fn; ctx->haloUser = user;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_slab_step_phase(mpmgpu_ctx *ctx, int phase)
{
    int rc = check_ready(ctx, "mpmgpu_slab_step_phase"); if (rc) return rc;
    if (!ctx->tiled.enabled) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_step_phase: fused path not enabled for this problem");
    if (phase < 0 || phase > 3) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_step_phase: phase %d", phase);
    if (ctx->sp.xpicOrder > 1 && ctx->tiled.slab.on && (ctx->tiled.hasLower || ctx->tiled.hasUpper) && !ctx->haloFn)
        return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_step_phase: XPIC/FMPM order %d > 1 needs one halo exchange per iteration: register mpmgpu_slab_set_halo_callback", ctx->sp.xpicOrder);
    rc = fused_phase(ctx, phase);
    if (rc) return rc;
    if (phase == 3) {
        ctx->mstep++; ctx->mtime += ctx->sp.dt;
        if (ctx->slabPending) return MPMGPU_OK;     // counts + flags are on their way: mpmgpu_slab_migration_counts waits for them
        if (ctx->tiled.slab.on) {
            CK(cudaMemcpyAsync(ctx->tiled.hLeave, ctx->tiled.slab.leaveCount, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        }
        rc = poll_flags(ctx);
        if (rc) return rc;
        if (ctx->tiled.slab.on && (ctx->tiled.hLeave[0] > ctx->tiled.migCap || ctx->tiled.hLeave[1] > ctx->tiled.migCap))
            return fail(ctx, MPMGPU_EINVAL, "slab migration capacity %d exceeded (%d, %d leavers)", ctx->tiled.migCap, ctx->tiled.hLeave[0], ctx->tiled.hLeave[1]);
    }
    return MPMGPU_OK;       // phases 0-2 are asynchronous: the host enqueues the halo exchange on the same stream
}

extern "C" int mpmgpu_slab_migration_counts(mpmgpu_ctx *ctx, int *n_lo, int *n_hi)
{
    if (!ctx || !ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_migration_counts: slab mode not configured");
    if (ctx->slabPending) {         // waits for the element reset only; the strain kernel behind it keeps the GPU busy
        cudaSetDevice(ctx->cfg.device);
        CK(cudaEventSynchronize(ctx->slabEvent));
        ctx->slabPending = false;
        TiledState &t = ctx->tiled;
        t.hLeave[0] = ctx->slabHost->leave[0]; t.hLeave[1] = ctx->slabHost->leave[1];
        ctx->hFlags = ctx->slabHost->flags;
        if (ctx->hFlags.nanParticle) return fail(ctx, MPMGPU_ENAN, "particle %d left grid with position nan (ResetElementsTask)", ctx->hFlags.nanParticle - 1);
        if (t.hLeave[0] > t.migCap || t.hLeave[1] > t.migCap)
            return fail(ctx, MPMGPU_EINVAL, "slab migration capacity %d exceeded (%d, %d leavers)", t.migCap, t.hLeave[0], t.hLeave[1]);
    }
    if (n_lo) *n_lo = ctx->tiled.hLeave[0];
    if (n_hi) *n_hi = ctx->tiled.hLeave[1];
    return MPMGPU_OK;
}

extern "C" int mpmgpu_slab_migration_buffers(mpmgpu_ctx *ctx, void **send_lo, void **send_hi, void **recv_lo, void **recv_hi, int *row_doubles, int *capacity_rows)
{
    if (!ctx || !ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_migration_buffers: slab mode not configured");
    TiledState &t = ctx->tiled;
    if (send_lo) *send_lo = t.migSend[0];
    if (send_hi) *send_hi = t.migSend[1];
    if (recv_lo) *recv_lo = t.migRecv[0];
    if (recv_hi) *recv_hi = t.migRecv[1];
    if (row_doubles) *row_doubles = MIG_ROW;
    if (capacity_rows) *capacity_rows = t.migCap;
    return MPMGPU_OK;
}

// Leavers are packed in ascending slot order (the order F4 appended them in is not reproducible), so the
// particle order on the receiving rank -- and with it every later sum -- is the same from run to run.
extern "C" int mpmgpu_slab_pack_migrants(mpmgpu_ctx *ctx)
{
    if (!ctx || !ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_pack_migrants: slab mode not configured");
    cudaSetDevice(ctx->cfg.device);
    TiledState &t = ctx->tiled;
    int bits = 1;
    while ((1ll << bits) < (long long)ctx->cap + 1) bits++;
    for (int side = 0; side < 2; side++) {
        const int nl = t.hLeave[side];
        if (nl <= 0) continue;
        int *list = t.slab.leaveIdx + (size_t)side * t.migCap;
        size_t tb = t.migCubBytes;
        CK(cub::DeviceRadixSort::SortKeys(t.migCubTemp, tb, list, t.migKeys, nl, 0, bits, ctx->stream));
        CK(cudaMemcpyAsync(list, t.migKeys, (size_t)nl * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->launches += 3;
        LAUNCH(k_mig_pack, nl, 64, nl, list, ctx->cap, NPD, ctx->particlePool, NPI, ctx->particleIntPool, MIG_ROW, t.migSend[side]);
    }
    return MPMGPU_OK;       // stream-ordered: the exchange is enqueued on the same stream
}

// remove the particles that left (fill their slots from the end), then append the arrivals; all on the device, no host sync
extern "C" int mpmgpu_slab_finish_migration(mpmgpu_ctx *ctx, int n_from_lo, int n_from_hi)
{
    if (!ctx || !ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_finish_migration: slab mode not configured");
    cudaSetDevice(ctx->cfg.device);
    TiledState &t = ctx->tiled;
    const int L = t.hLeave[0] + t.hLeave[1];
    int n = ctx->P.n;
    if (L > 0) {
        if (t.hLeave[0]) CK(cudaMemcpyAsync(t.migKeys, t.slab.leaveIdx, (size_t)t.hLeave[0] * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        if (t.hLeave[1]) CK(cudaMemcpyAsync(t.migKeys + t.hLeave[0], t.slab.leaveIdx + t.migCap, (size_t)t.hLeave[1] * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        int bits = 1;
        while ((1ll << bits) < (long long)ctx->cap + 1) bits++;
        size_t tb = t.migCubBytes;
        CK(cub::DeviceRadixSort::SortKeys(t.migCubTemp, tb, t.migKeys, t.migSorted, L, 0, bits, ctx->stream));
        ctx->launches += 4;
        LAUNCH(k_mig_plan, 1, MIG_PLAN_THREADS, n, L, t.migSorted, t.migFillers, t.migPairs, t.migFlags);
        LAUNCH(k_mig_fill, L, 64, t.migPairs, t.migSorted, t.migFillers, ctx->cap, NPD, ctx->particlePool, NPI, ctx->particleIntPool);
        n -= L;
    }
    const int R = n_from_lo + n_from_hi;
    if ((size_t)(n + R) > ctx->cap) return fail(ctx, MPMGPU_EINVAL, "particle capacity %zu exceeded by migration (%d + %d); raise max_particles", ctx->cap, n, R);
    if (n_from_lo > 0) LAUNCH(k_mig_unpack, n_from_lo, 64, n_from_lo, n, ctx->cap, NPD, ctx->particlePool, NPI, ctx->particleIntPool, MIG_ROW, t.migRecv[0]);
    if (n_from_hi > 0) LAUNCH(k_mig_unpack, n_from_hi, 64, n_from_hi, n + n_from_lo, ctx->cap, NPD, ctx->particlePool, NPI, ctx->particleIntPool, MIG_ROW, t.migRecv[1]);
    n += R;
    ctx->P.n = n; ctx->P.nNR = n;
    CK(cudaMemsetAsync(t.slab.leaveCount, 0, 2 * sizeof(int), ctx->stream));
    t.hLeave[0] = t.hLeave[1] = 0;
    return MPMGPU_OK;
}


// ---- NCCL inside the library ---------------------------------------------------------------------------------------------
// The slab exchange of the reference's GridPatch/GhostNode decomposition (Patches/GridPatch.cpp:214-251, GhostNode.cpp:127-185) as
// ncclSend/ncclRecv pairs with the lower and upper z-neighbour, issued by the library itself on the context's stream.
// libnccl is resolved at run time (the copy the process already holds -- torch's -- or the system's), so libmpmgpu has no link-time
// dependency on it and a single-GPU user never loads it.
struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi *slab_nccl(void)
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : NULL;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return NULL;
    api.GetUniqueId = (ncclResult_t (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (ncclResult_t (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (ncclResult_t (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    api.Send = (ncclResult_t (*)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
    api.Recv = (ncclResult_t (*)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
    api.GroupStart = (ncclResult_t (*)(void))dlsym(h, "ncclGroupStart");
    api.GroupEnd = (ncclResult_t (*)(void))dlsym(h, "ncclGroupEnd");
    api.GetErrorString = (const char *(*)(ncclResult_t))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Send || !api.Recv || !api.GroupStart || !api.GroupEnd || !api.GetErrorString) return NULL;
    api.lib = h;
    return &api;
}

#define NCCLCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(ctx, MPMGPU_ECUDA, "NCCL: %s (%s)", nccl->GetErrorString(r_), #call); } while (0)

static void slab_disconnect(mpmgpu_ctx *ctx)
{
    NcclApi *nccl = slab_nccl();
    if (nccl && ctx->comm) { nccl->CommDestroy(ctx->comm); ctx->comm = NULL; }
    if (nccl && ctx->commSide) { nccl->CommDestroy(ctx->commSide); ctx->commSide = NULL; }
    if (ctx->sideStream) { cudaStreamDestroy(ctx->sideStream); ctx->sideStream = NULL; }
    if (ctx->sideEvent) { cudaEventDestroy(ctx->sideEvent); ctx->sideEvent = NULL; }
    if (ctx->hCounts) { cudaFreeHost(ctx->hCounts); ctx->hCounts = NULL; }
}

// 2 x 128 bytes: the ids of the two communicators of a run (rank 0 makes them, the caller hands them to every rank)
extern "C" int mpmgpu_nccl_unique_ids(void *ids256)
{
    NcclApi *nccl = slab_nccl();
    if (!nccl || !ids256) { g_create_error = "mpmgpu_nccl_unique_ids: libnccl.so.2 not found"; return MPMGPU_EINVAL; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    for (int k = 0; k < 2; k++) {
        ncclUniqueId id;
        if (nccl->GetUniqueId(&id) != ncclSuccess) return MPMGPU_ECUDA;
        memcpy((char *)ids256 + 128 * k, &id, 128);
    }
    return MPMGPU_OK;
}

extern "C" int mpmgpu_slab_connect(mpmgpu_ctx *ctx, int rank, int world, const void *ids256)
{
    if (!ctx || !ids256 || world < 1 || rank < 0 || rank >= world) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_connect: bad argument");
    if (!ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_connect: call mpmgpu_slab_configure first");
    NcclApi *nccl = slab_nccl();
    if (!nccl) return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_connect: libnccl.so.2 not found");
    if ((ctx->tiled.hasLower != 0) != (rank > 0) || (ctx->tiled.hasUpper != 0) != (rank < world - 1))
        return fail(ctx, MPMGPU_EINVAL, "mpmgpu_slab_connect: rank %d of %d does not match the neighbours given to mpmgpu_slab_configure", rank, world);
    cudaSetDevice(ctx->cfg.device);
    ncclUniqueId id[2];
    memcpy(id, ids256, 256);
    NCCLCK(nccl->CommInitRank(&ctx->comm, world, id[0], rank));
    NCCLCK(nccl->CommInitRank(&ctx->commSide, world, id[1], rank));
    ctx->rank = rank; ctx->world = world;
    CK(cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->sideEvent, cudaEventDisableTiming));
    CK(dalloc(ctx, &ctx->dCounts, 4));
    CK(cudaMallocHost((void **)&ctx->hCounts, 4 * sizeof(int)));
    return MPMGPU_OK;
}

// swap halo kind `which` with the neighbours: both directions of both faces in one NCCL group, in stream order
static int slab_exchange_halo(mpmgpu_ctx *ctx, int which)
{
    NcclApi *nccl = slab_nccl();
    TiledState &t = ctx->tiled;
    static const int values[4] = {4, 3, 3, 3};
    const size_t nd = (size_t)values[which] * 3 * t.planeNodes;
    if (!t.hasLower && !t.hasUpper) return MPMGPU_OK;
    NCCLCK(nccl->GroupStart());
    if (t.hasLower) { NCCLCK(nccl->Send(t.haloSend[0], nd, ncclDouble, ctx->rank - 1, ctx->comm, ctx->stream)); NCCLCK(nccl->Recv(t.haloRecv[0], nd, ncclDouble, ctx->rank - 1, ctx->comm, ctx->stream)); }
    if (t.hasUpper) { NCCLCK(nccl->Send(t.haloSend[1], nd, ncclDouble, ctx->rank + 1, ctx->comm, ctx->stream)); NCCLCK(nccl->Recv(t.haloRecv[1], nd, ncclDouble, ctx->rank + 1, ctx->comm, ctx->stream)); }
    NCCLCK(nccl->GroupEnd());
    ctx->launches++;
    return MPMGPU_OK;
}

static void slab_halo_hook(void *user, int which) { slab_exchange_halo((mpmgpu_ctx *)user, which); }

// particle migration after the element reset: counts to the neighbours on the side communicator (while the second strain
// update is still running on the main stream), rows on the main one
static int slab_migrate(mpmgpu_ctx *ctx)
{
    NcclApi *nccl = slab_nccl();
    TiledState &t = ctx->tiled;
    int nLo = 0, nHi = 0, rc;
    if ((rc = mpmgpu_slab_migration_counts(ctx, &nLo, &nHi))) return rc;
    if (!t.hasLower && !t.hasUpper) return MPMGPU_OK;
    int *h = ctx->hCounts;
    h[0] = nLo; h[1] = nHi; h[2] = 0; h[3] = 0;
    CK(cudaMemcpyAsync(ctx->dCounts, h, 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->sideStream));
    NCCLCK(nccl->GroupStart());
    if (t.hasLower) { NCCLCK(nccl->Send(ctx->dCounts + 0, 1, ncclInt32, ctx->rank - 1, ctx->commSide, ctx->sideStream)); NCCLCK(nccl->Recv(ctx->dCounts + 2, 1, ncclInt32, ctx->rank - 1, ctx->commSide, ctx->sideStream)); }
    if (t.hasUpper) { NCCLCK(nccl->Send(ctx->dCounts + 1, 1, ncclInt32, ctx->rank + 1, ctx->commSide, ctx->sideStream)); NCCLCK(nccl->Recv(ctx->dCounts + 3, 1, ncclInt32, ctx->rank + 1, ctx->commSide, ctx->sideStream)); }
    NCCLCK(nccl->GroupEnd());
    CK(cudaMemcpyAsync(h + 2, ctx->dCounts + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->sideStream));
    CK(cudaStreamSynchronize(ctx->sideStream));
    const int fLo = t.hasLower ? h[2] : 0, fHi = t.hasUpper ? h[3] : 0;
    if (fLo > t.migCap || fHi > t.migCap) return fail(ctx, MPMGPU_EINVAL, "slab migration capacity %d exceeded by arrivals (%d, %d)", t.migCap, fLo, fHi);
    if (nLo || nHi) { if ((rc = mpmgpu_slab_pack_migrants(ctx))) return rc; }
    if (nLo || nHi || fLo || fHi) {
        NCCLCK(nccl->GroupStart());
        if (t.hasLower && nLo) NCCLCK(nccl->Send(t.migSend[0], (size_t)nLo * MIG_ROW, ncclDouble, ctx->rank - 1, ctx->comm, ctx->stream));
        if (t.hasLower && fLo) NCCLCK(nccl->Recv(t.migRecv[0], (size_t)fLo * MIG_ROW, ncclDouble, ctx->rank - 1, ctx->comm, ctx->stream));
        if (t.hasUpper && nHi) NCCLCK(nccl->Send(t.migSend[1], (size_t)nHi * MIG_ROW, ncclDouble, ctx->rank + 1, ctx->comm, ctx->stream));
        if (t.hasUpper && fHi) NCCLCK(nccl->Recv(t.migRecv[1], (size_t)fHi * MIG_ROW, ncclDouble, ctx->rank + 1, ctx->comm, ctx->stream));
        NCCLCK(nccl->GroupEnd());
        ctx->launches++;
        if ((rc = mpmgpu_slab_finish_migration(ctx, fLo, fHi))) return rc;
        ctx->migratedOut += nLo + nHi; ctx->migratedIn += fLo + fHi;
    }
    return MPMGPU_OK;
}

// nsteps full steps of one slab of a multi-GPU run, exchanges included (every rank calls it with the same nsteps)
extern "C" int mpmgpu_slab_step(mpmgpu_ctx *ctx, int nsteps)
{
    int rc = check_ready(ctx, "mpmgpu_slab_step"); if (rc) return rc;
    if (!ctx->tiled.slab.on) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_step: slab mode not configured");
    const bool alone = !ctx->tiled.hasLower && !ctx->tiled.hasUpper;
    if (!alone && !ctx->comm) return fail(ctx, MPMGPU_ESTATE, "mpmgpu_slab_step: call mpmgpu_slab_connect first");
    if (!alone && !ctx->haloFn) { ctx->haloFn = slab_halo_hook; ctx->haloUser = ctx; }     // the exchange inside each XPIC/FMPM iteration
    for (int s = 0; s < nsteps; s++) {
        for (int phase = 0; phase < 3; phase++) {
            if ((rc = mpmgpu_slab_step_phase(ctx, phase))) return rc;
            if (!alone && (rc = slab_exchange_halo(ctx, phase))) return rc;
        }
        if ((rc = mpmgpu_slab_step_phase(ctx, 3))) return rc;
        if ((rc = slab_migrate(ctx))) return rc;
    }
    return MPMGPU_OK;
}

extern "C" int mpmgpu_slab_migrated(const mpmgpu_ctx *ctx, long long *out, long long *in)
{
    if (!ctx) return MPMGPU_EINVAL;
    if (out) *out = ctx->migratedOut;
    if (in) *in = ctx->migratedIn;
    return MPMGPU_OK;
}

// Run on the caller's stream (e.g. torch's current stream, so NCCL calls and kernels are ordered
// without host synchronisation).  The context's own stream is kept for destroy.
extern "C" int mpmgpu_set_stream(mpmgpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return MPMGPU_EINVAL;
    cudaSetDevice(ctx->cfg.device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->ownStreamSaved) { ctx->ownStream = ctx->stream; ctx->ownStreamSaved = true; }
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->ownStream;
    return MPMGPU_OK;
}

extern "C" int mpmgpu_num_particles(const mpmgpu_ctx *ctx) { return ctx ? ctx->P.n + ctx->PR.n : 0; }
