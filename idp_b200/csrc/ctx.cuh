// Context, device buffers and launch helpers shared by the translation units of libidp_contact.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/idp_contact.h"

#define IDP_EVENT_POOL 512
// the sharded path keeps per-rank counts, thresholds and exchange matrices in fixed-size host / device arrays
#define IDP_MAX_RANKS 8

namespace idp {

// one cache-line-pair record per boundary primitive (point / edge / triangle): exact AABB + vertex ids.
// For the static phase the AABB is that of the current positions, for CCD it is the full-step swept AABB
// (min/max over x and x + dir, Math/Distance/CCD.h:187-235).
struct __align__(64) PrimRec {
    double lo[3];
    double hi[3];
    int v[3];   // vertex ids (-1 = unused)
    int flags;  // bit0: all vertices are Dirichlet
};
// integer lattice box of a primitive (cell coordinates of the broad-phase lattice)
struct __align__(32) IBox {
    int lo[3];
    int hi[3];
    int pad[2];
};
struct __align__(16) Row4 {
    int a, b, c, d;
    __host__ __device__ bool operator==(const Row4& o) const { return a == o.a && b == o.b && c == o.c && d == o.d; }
    __host__ __device__ bool operator!=(const Row4& o) const { return !(*this == o); }
};
struct RowLess {
    __host__ __device__ bool operator()(const Row4& x, const Row4& y) const
    {
        if (x.a != y.a) return x.a < y.a;
        if (x.b != y.b) return x.b < y.b;
        if (x.c != y.c) return x.c < y.c;
        return x.d < y.d;
    }
};

// lattice -> cell grid used by the broad phase
struct GridDesc {
    double lo[3];   // lattice origin
    double inv;     // 1 / lattice spacing
    int k;          // lattice voxels per grid cell (per axis)
    int n[3];       // grid cells per axis
    int clampLat;   // 1: lattice indices are clamped to the grid (static phase); 0: reference semantics (CCD)
};

inline long& device_alloc_counter() { static long n = 0; return n; } // cudaMalloc calls made by DBuf (steady state: 0 per step)

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0; // elements
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // grow (contents are NOT preserved unless keep=true)
    cudaError_t reserve(size_t n, bool keep = false, cudaStream_t s = 0)
    {
        if (n <= cap) return cudaSuccess;
        size_t ncap = std::max(n, cap + cap / 2);
        T* np = nullptr;
        cudaError_t e = cudaMalloc((void**)&np, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        ++device_alloc_counter();
        if (keep && p && cap) {
            e = cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return e;
            cudaStreamSynchronize(s);
        }
        if (p) cudaFree(p);
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
};

struct StageTimes {
    // milliseconds of the last call of each stage (CUDA events on ctx->stream), names follow the reference's
    // TIMER_FLAG scopes
    float v[IDP_STAGE_COUNT];
};

} // namespace idp

struct idp_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    bool own_stream = false;
    // ---- asynchronous result transfers (idp_get_*_begin / idp_transfers_end): a copy stream beside the compute stream ----
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evCopyFork = nullptr, evRows = nullptr, evBlk = nullptr, evVal = nullptr;
    bool pendRows = false, pendCsr = false;
    int* pendRowsOut = nullptr; double* pendInfoOut = nullptr; long pendRowsN = 0;
    int* pendPtr = nullptr; int* pendCol = nullptr; double* pendVal = nullptr;
    int* h_blk = nullptr; size_t h_blk_cap = 0;     // pinned: rowStart (nV + 1) + block column vertices
    void* rowsWorker = nullptr; void* csrWorker = nullptr; // std::thread*: host-side fill / expansion running beside the copies
    int workerStatus = 0;
    idp::DBuf<int> blkCol;
    cudaStream_t commStream = nullptr;              // NCCL collectives that overlap local work (idp_barrier_all)
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    std::string err;
    long launches = 0;      // kernels of this library launched since the last idp_reset_counters
    long lib_launches = 0;  // CUB device-wide primitives invoked (each expands to a few library kernels)
    int sm_count = 148;

    // sharding (multi-GPU): this context evaluates query primitives / rows of shard `rank` of `nranks`
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    void* local_group = nullptr; // in-process group (idp_comm_init_local): same sharded code path, peer-memory transport

    // ---- mesh (set once per time step) ----
    int nV = 0, nBN = 0, nBE = 0, nBT = 0;
    idp::DBuf<int> bnode;
    idp::DBuf<int2> bedge;
    idp::DBuf<int4> btri;
    idp::DBuf<unsigned char> dbc, projMask;
    // ---- per-iterate state ----
    idp::DBuf<double> stage;            // upload staging (AoS)
    idp::DBuf<double> xs, ys, zs;       // SoA positions (streaming kernels)
    idp::DBuf<double4> xp, x0p, dp, xnp; // 32-byte packed positions / rest positions / search direction / step-start positions (gathers)
    bool have_x = false, have_x0 = false, have_dir = false;
    long xVersion = 0, meanEdgeVersion = -1; // positions counter; mean boundary-edge length cached per positions (hash builds)
    double meanEdgeCached = 0;
    // ---- broad-phase scratch ----
    idp::DBuf<idp::PrimRec> recN, recE, recT;
    idp::DBuf<idp::IBox> boxNq, boxEq, boxEb, boxTb; // query boxes (inflated) and insert boxes
    idp::DBuf<idp::IBox> vbox;                        // per-vertex lattice box (CCD)
    idp::DBuf<int> cellStart, cellCursor, largeList, histScratch, prepRegion;
    idp::DBuf<int4> crec0, crec1;       // cell-sorted filter records, SoA 2 x 16 B (see k_cells)
    idp::DBuf<int2> candPT, candEE;
    long nCandPT = 0, nCandEE = 0;
    idp::DBuf<double> red;              // reduction scratch
    idp::DBuf<unsigned char> cubTemp;
    idp::DBuf<long long> commCounts;    // P x P exchange counts
    idp::DBuf<double> commTmp;          // in-process group: reduction result before it replaces the input
    idp::DBuf<long long> counters;      // device counters / flags (see enum in kernels)
    // ---- constraint set ----
    idp::DBuf<idp::Row4> rowsA, rowsB, rowsD, rowsD2, rows, rowsG;
    idp::DBuf<int> runCounts;
    idp::DBuf<unsigned long long> keyA, keyB, keyD, keyTmp; // sort keys of the row groups
    idp::DBuf<double> weights;
    long nRows = 0;
    // sharded LOCAL-ROWS mode (set by build_constraint_set): rows holds only this rank's [direct PT][direct EE][merged] rows
    bool rowsLocal = false;
    long nRowsGlobal = 0;
    long shardCnt[IDP_MAX_RANKS][3] = {};
    idp::DBuf<idp::Row4> rowsGlobal;     // gathered list (idp_get_constraints)
    idp::DBuf<double> dist2Global;
    double cs_dhat2 = 0; // dHat2 stored in stencilInfo (after the thickness offset, IPC.h:53-54)
    // ---- barrier outputs ----
    idp::DBuf<double> gbuf;             // 3*nV gradient (xyz interleaved)
    idp::DBuf<double> rowDist2;
    bool dist2Valid = false;            // rowDist2 holds the values of the current rows (set by idp_min_dist2)
    // evaluation order of this rank's rows: stable sort by row kind (uniform warps); rebuilt when the rows change
    idp::DBuf<unsigned char> rowKind, rowKindSorted;
    idp::DBuf<int> rowIota, rowPerm;
    idp::DBuf<idp::Row4> rowsK;         // this rank's rows in evaluation (kind-sorted) order: k_barrier streams them
    idp::DBuf<double> weightsK;
    long kindCount[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool permValid = false;
    // Hessian blocks bucketed by lower vertex (BucketEmit) and their per-vertex reduction (assemble_csr)
    idp::DBuf<int> vtxCnt, vtxOff, vtxCursor;
    idp::DBuf<unsigned long long> bktKey, bigKey;
    idp::DBuf<double> bktVal8, bktVal1, uVal;
    idp::DBuf<int> uCount, uCol, usrc, bigList;
    idp::DBuf<unsigned> bigPay, bigStart;
    long nBlocksEmitted = 0;
    idp::DBuf<unsigned long long> blkKey, blkKeySorted; // scratch of idp_get_candidates
    idp::DBuf<int> segId;
    idp::DBuf<int> vtxBlkStart, rowStart, lowerCount, lstart, urow, ucol, lkey, lkeySorted, lseg, lsegSorted;
    idp::DBuf<int> csrPtr, csrCol;
    idp::DBuf<double> csrVal;
    long nnz = 0, nBlocksUnique = 0;
    // ---- the rest of the Newton system assembled into the same CSR (flow term, lumped mass) ----
    idp::DBuf<int> flowElem;            // 3 vertex ids per triangle element
    idp::DBuf<double> flowVol, massDiag;
    int nFlowElem = 0;
    double flowH = 0;
    bool haveMass = false;
    bool csrProjected = false;          // idp_project_dbc has been applied to the current CSR
    // ---- elastic terms of the shell system (membrane triangles, bending hinges; elastic_kernels.cu) ----
    idp::DBuf<int> memElem, hingeV;      // 3 / 4 vertex ids per element
    idp::DBuf<double> memIB, memCoef, memLambda, memMu, hingeInfo; // IB (3 per triangle), h^2 vol, Lame parameters; (thetabar, ebar, hbar)
    idp::DBuf<double> elasticG;         // gradient of the elastic terms (3 nV)
    int nMem = 0, nHinge = 0;
    double hingeKh2 = 0;                // h^2 k
    // ---- lagged friction (friction_kernels.cu): rows frozen by idp_friction_update ----
    idp::DBuf<unsigned char> fricRows;  // FricRow records (friction.cuh)
    idp::DBuf<double> fricG;            // friction gradient (3 nV)
    long nFric = 0, nFricActive = 0;
    idp::DBuf<int> fricCompRange;       // Compute_Friction_Coef: component c holds the vertices below fricCompRange[c] (and above the previous bound)
    idp::DBuf<double> fricMuComp;       // nComp x nComp coefficients
    int nFricComp = 0;
    double fricMu = 0, fricEpsvh = 0;
    bool have_xn = false;
    // ---- device-side surface extraction (idp_set_mesh_from_triangles) ----
    idp::DBuf<int> surfTri;
    idp::DBuf<double> surfTriArea, surfTriAreaH, surfNodeArea, surfNodeAreaC, surfEdgeArea, surfEdgeArea2;
    bool surfValid = false;
    // ---- PCG (idp_solve_pcg) ----
    idp::DBuf<double> pcgX, pcgR, pcgZ, pcgP, pcgP2, pcgAp, pcgInvDiag, pcgScal;
    // ---- CCD ----
    long ccd_iters = 0;
    long nCcdPT = 0, nCcdEE = 0;
    // host pinned scratch for small readbacks
    long long* h_counters = nullptr;
    double* h_red = nullptr;
    void* h_stage = nullptr;            // pinned staging for gradient / weight read-backs
    size_t h_stage_bytes = 0;
    bool weights_all_one = false;

    idp::StageTimes times;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr; // microbenchmark timing
    cudaEvent_t evPool[2 * IDP_EVENT_POOL] = {};
    int pendStage[IDP_EVENT_POOL] = {};
    bool pendAccum[IDP_EVENT_POOL] = {};
    int nPending = 0, timerDepth = 0;
};

namespace idp {

static inline int fail(idp_ctx* c, int code, const char* fmt, const char* what, const char* file, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), fmt, what, file, line);
    c->err = buf;
    return code;
}
#define IDP_CK(ctx, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) return idp::fail(ctx, IDP_ERR_CUDA, "CUDA error: %s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define IDP_TRY(expr)                 \
    do {                              \
        int st__ = (expr);            \
        if (st__ != IDP_OK) return st__; \
    } while (0)
#define IDP_LAUNCH(ctx, kernel, grid, block, smem, ...)                    \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);  \
        ++(ctx)->launches;                                                 \
    } while (0)

static inline unsigned blocks_for(long n, int block) { return (unsigned)std::max(1L, (n + block - 1) / block); }

// device counter slots
enum Counter {
    CNT_CAND = 0,     // candidate pairs written by the current query kernel
    CNT_ROWS_A = 1,   // direct rows (PT / EE kinds)
    CNT_ROWS_D = 2,   // plain PP / PE rows awaiting duplicate merge
    CNT_ERR_DIST = 3, // rows with non-positive distance
    CNT_ERR_CCD = 4,  // ACCD iteration cap reached or zero step
    CNT_CCD_ITERS = 5,
    CNT_RUNS = 6,
    CNT_ALPHA_BITS = 7, // current CCD step as double bits (atomicMin on positive doubles)
    CNT_KINDS = 16,     // 8 slots: rows of this rank per kind (k_row_kinds)
    CNT_SHARD = 24,     // 3 x 8 slots: (direct PT, direct EE, merged) row counts of every rank
    CNT_ERR_EIG = 48,   // rows whose PSD projection hit the QL iteration cap
    CNT_STATUS = 49,    // status agreement cell (comm_agree_status)
    CNT_COUNT = 52
};

// Device-side timers that never block the host: a scope records an event pair from a small pool and the elapsed times
// are resolved lazily (idp_stage_ms, or when the pool wraps). Synchronising inside the destructors, as a first version
// did, cost one host round trip per stage, kernel and collective -- visible at 8 GPUs where a step is ~20 ms.
inline void timers_resolve(idp_ctx* c)
{
    if (c->nPending == 0) return;
    for (int i = 0; i < c->nPending; ++i) {
        float ms = 0;
        if (cudaEventSynchronize(c->evPool[2 * i + 1]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&ms, c->evPool[2 * i], c->evPool[2 * i + 1]) != cudaSuccess) continue;
        const int st = c->pendStage[i];
        if (c->pendAccum[i]) c->times.v[st] += ms;
        else c->times.v[st] = ms;
    }
    c->nPending = 0;
}
struct ScopeTimer {
    idp_ctx* c;
    int slot;
    ScopeTimer(idp_ctx* ctx, int stage, bool accumulate) : c(ctx), slot(-1)
    {
        if (c->nPending == IDP_EVENT_POOL && c->timerDepth == 0) timers_resolve(c); // only between top-level scopes
        ++c->timerDepth;
        if (c->nPending == IDP_EVENT_POOL) return; // pool exhausted inside a nested scope: this one goes untimed
        slot = c->nPending++;
        c->pendStage[slot] = stage;
        c->pendAccum[slot] = accumulate;
        cudaEventRecord(c->evPool[2 * slot], c->stream);
    }
    ~ScopeTimer()
    {
        --c->timerDepth;
        if (slot >= 0) cudaEventRecord(c->evPool[2 * slot + 1], c->stream);
    }
};
struct StageTimer : ScopeTimer {
    StageTimer(idp_ctx* ctx, int st) : ScopeTimer(ctx, st, false) {}
};
// times a single launch (nested inside a StageTimer scope)
struct KernelTimer : ScopeTimer {
    KernelTimer(idp_ctx* ctx, int st) : ScopeTimer(ctx, st, false) {}
};

// ---- host-side launchers implemented in the .cu files ----
int upload_positions(idp_ctx* c, const double* host_xyz, int stride, int which); // which: 0 X, 1 X0, 2 dir, 3 Xn (friction)
int build_constraint_set(idp_ctx* c, double dhat2, double thickness);
int sorted_candidates(idp_ctx* c, int which, int2* host_out);
int barrier_eval(idp_ctx* c, double dhat2, double kappa, double thickness, int want_e, int want_g, int want_h,
    int project_spd, double* E_out);
int assemble_csr(idp_ctx* c);
int project_dbc(idp_ctx* c, const unsigned char* host_mask = nullptr);
// elastic terms: bucket counts / Hessian blocks for the assembly in barrier_eval, energy + gradient on request
int elastic_block_counts(idp_ctx* c, int* vtxCnt, long* nElements);
int elastic_emit_blocks(idp_ctx* c, int project_spd, unsigned tagBase, int* vtxCursor, unsigned long long* bktKey, double* bktVal8, double* bktVal1);
int elastic_energy_gradient(idp_ctx* c, int want_e, int want_g, double* E_out);
// lagged friction: basis from the current rows / positions; bucket counts / Hessian blocks for the assembly; E and g
int friction_update(idp_ctx* c, double dhat2, double kappa, double thickness);
int friction_block_counts(idp_ctx* c, int* vtxCnt, long* nRows);
int friction_emit_blocks(idp_ctx* c, unsigned tagBase, int* vtxCursor, unsigned long long* bktKey, double* bktVal8, double* bktVal1);
int friction_energy_gradient(idp_ctx* c, int want_e, int want_g, double* E_out);
int friction_copy_rows(idp_ctx* c, int* rows4, double* closest2, double* basis6, double* normalForce);
int solve_pcg(idp_ctx* c, const double* rhs, double* sol, double rel_tol, int max_iter, int* iters, double* rel_res);
int extract_surface(idp_ctx* c, int nV, int nF, const int* tri, int stride, const double* x, int xstride, const unsigned char* dbc);
int min_dist2(idp_ctx* c, double thickness, double* host_dist2, double* min_out);
int ccd_step(idp_ctx* c, double thickness, double* alpha_inout, int keep_candidates);
int cub_scan_exclusive(idp_ctx* c, const int* in, int* out, long n);
// variable-size all-gather of nLocal elements of elemSize bytes into *outPtr (capacity *outCap elements, grown and
// preserved when too small) starting at element outOffset; *nTotal = sum over ranks
inline bool comm_on(const idp_ctx* c) { return c->nranks > 1 && (c->nccl_comm || c->local_group); }
int comm_allreduce_min(idp_ctx* c, double* dev, long n);
int comm_allreduce_sum(idp_ctx* c, double* dev, long n);
int comm_allreduce_max(idp_ctx* c, double* dev, long n);
int comm_agree_status(idp_ctx* c, int status); // max over ranks of a status code (collective; identity when unsharded)
int comm_allgather_i64(idp_ctx* c, long long* dev, long perRank); // in place: rank r's perRank values at dev + r*perRank
int comm_gather_groups(idp_ctx* c, const void* local, size_t elemSize, void* globalOut); // local [A|B|U] -> global [A..|B..|U..]
int comm_allreduce_min_u64(idp_ctx* c, unsigned long long* dev, long n);
int comm_exchange_keys(idp_ctx* c, const unsigned long long* keys, const long sendBegin[IDP_MAX_RANKS], const long sendCount[IDP_MAX_RANKS], DBuf<unsigned long long>& recv, long* nRecv);
int comm_allgatherv(idp_ctx* c, const void* local, long nLocal, size_t elemSize, void** outPtr, size_t* outCap, long outOffset, long* nTotal);

} // namespace idp
