// Barrier energy / gradient / Hessian per constraint row, per-row PSD projection, and the sort-reduce assembly of the
// 3x3 blocks into a device-resident scalar CSR.
//
// Reference operators replaced (relative to /root/reference/Library):
//   FEM/IPC.h:742-941 (Compute_Barrier), 943-1256 (Compute_Barrier_Gradient), 1258-1731 (Compute_Barrier_Hessian),
//   Math/UTILS.h:9-27 (makePD), Math/CSR_MATRIX.h:49-56 (Construct_From_Triplet = Eigen setFromTriplets).
// Tolerance of these outputs is 1e-10 relative (BASELINE.json), so FMA contraction is allowed in this unit.
#include "ctx.cuh"
#include "pair_deriv.cuh"
#include "psd_lowrank.cuh"
#include "bucket_emit.cuh"
#include <cub/cub.cuh>

#ifndef IDP_BARRIER_T0
#define IDP_BARRIER_T0 128
#define IDP_BARRIER_B0 2
#endif

namespace idp {

__device__ __forceinline__ V3 ldv4(const double4* __restrict__ p, int v)
{
    const double2* q = reinterpret_cast<const double2*>(p + v);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return mk3(a.x, a.y, b.x);
}

// Row -> rank assignment of the sharded path: the rank that owns the vertex chunk of the row's smallest vertex, chunks of
// 2^shift consecutive vertices dealt round-robin. Rows that touch the same vertices land on the same rank, so the
// per-rank partial CSRs are nearly disjoint (only chunk borders overlap; their sum is the global Hessian either way),
// while the round-robin over ~16 chunks per rank keeps the mix of cheap and expensive row kinds balanced.
__device__ __forceinline__ int row_owner(const Row4& r, int shift, int nranks)
{
    const RowDec d = decode_row(r.a, r.b, r.c, r.d);
    int mv = d.v[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) mv = (k < d.nv && d.v[k] < mv) ? d.v[k] : mv;
    return (mv >> shift) % nranks;
}
// Hessian blocks are bucketed by their LOWER vertex: the local Hessian is symmetric, so one block per unordered vertex pair
// (vlo <= vhi) is emitted, nv(nv+1)/2 per row (IPC.h:1372-1387 emits all nv^2 x 9 scalars), and vertex v collects the blocks
// whose lower vertex it is. This pass counts them: vertex v[k] of a row receives 1 + #{m : v[m] > v[k]} blocks.
__global__ void k_vertex_block_counts(const Row4* __restrict__ rows, long n, int rank, int nranks, int shift, int* __restrict__ cnt)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const Row4 r = rows[i];
        if (nranks != 1 && row_owner(r, shift, nranks) != rank) continue;
        const RowDec d = decode_row(r.a, r.b, r.c, r.d);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= d.nv) break;
            int m = 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) m += (j < d.nv && d.v[j] > d.v[k]) ? 1 : 0;
            atomicAdd(&cnt[d.v[k]], m);
        }
    }
}
// kind of every row of this rank (7 = row of another rank) + identity permutation; kinds are counted per block and added
// once (counts are order independent, so the result is deterministic)
__global__ void __launch_bounds__(256) k_row_kinds(const Row4* __restrict__ rows, long n, int rank, int nranks, int shift,
    unsigned char* __restrict__ kind, int* __restrict__ iota, unsigned long long* __restrict__ kindCount)
{
    __shared__ int sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const Row4 r = rows[i];
        int k = 7;
        if (nranks == 1 || row_owner(r, shift, nranks) == rank) k = decode_row(r.a, r.b, r.c, r.d).kind;
        kind[i] = (unsigned char)k;
        iota[i] = (int)i;
        atomicAdd(&sh[k], 1);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(&kindCount[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// copies of the rows / weights in evaluation order (cached with the permutation): the barrier kernel reads them coalesced and
// can fetch the next row while it works on the current one instead of chasing perm -> row -> positions
__global__ void __launch_bounds__(256) k_gather_rows(const Row4* __restrict__ rows, const double* __restrict__ weights, const int* __restrict__ perm, long n,
    Row4* __restrict__ rowsK, double* __restrict__ weightsK)
{
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long)gridDim.x * blockDim.x) {
        const int i = perm[j];
        rowsK[j] = rows[i];
        weightsK[j] = weights[i];
    }
}

struct BarrierArgs {
    const Row4* rows; const double* weights; long jBegin, jEnd; // rows [jBegin, jEnd) of the kind-sorted copies: one path
    const double4* xp; const double4* x0p;
    double dHat2, kappa, xi2;
    int projectSPD;
    double* partialE;                 // one slot per block
    double* g;                        // 3*nV, xyz interleaved (atomics)
    const int* vtxOff; int* vtxCursor; unsigned long long* bktKey; double* bktVal8; double* bktVal1; // Hessian block buckets
    unsigned long long* errDist;
    unsigned long long* errEig;
};

// ---- the other Hessian terms of the flow Newton system, emitted into the same buckets --------------
// Laplacian flow term (Shell/INC_POTENTIAL.h:323-339): per triangle element and axis d the reference appends the triplets
// (v_i d, v_i d, 2 h vol / 6) and (v_i d, v_j d, -h vol / 6) for the two other vertices j -- i.e. the 3x3 blocks
// (v_i, v_i) = (2 h vol / 6) I and (v_i, v_j) = (-h vol / 6) I; the symmetric assembly takes the upper blocks.
// Lumped mass (`sysMtr += M`, INC_POTENTIAL.h:383-386; M diagonal, Shell/DISCRETE_SHELL.h:279-318): block (v, v) = m_v I.
struct ExtraArgs {
    const int* elem; const double* vol; int eBegin, eEnd; double h;
    const double* mass; int vBegin, vEnd;
    int* vtxCnt; int* vtxCursor; unsigned long long* bktKey; double* bktVal8; double* bktVal1;
    unsigned tagBase;
};
__global__ void __launch_bounds__(256) k_extra_counts(ExtraArgs a)
{
    const int nE = a.eEnd - a.eBegin, nM = a.mass ? a.vEnd - a.vBegin : 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nE + nM; t += gridDim.x * blockDim.x) {
        if (t < nE) {
            const int e = a.eBegin + t;
            const int v[3] = {a.elem[3 * e], a.elem[3 * e + 1], a.elem[3 * e + 2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int m = 1;
#pragma unroll
                for (int j = 0; j < 3; ++j) m += v[j] > v[k] ? 1 : 0;
                atomicAdd(&a.vtxCnt[v[k]], m);
            }
        }
        else {
            const int v = a.vBegin + (t - nE);
            if (a.mass[v] != 0.0) atomicAdd(&a.vtxCnt[v], 1);
        }
    }
}
__global__ void __launch_bounds__(256) k_extra_emit(ExtraArgs a)
{
    const int nE = a.eEnd - a.eBegin, nM = a.mass ? a.vEnd - a.vBegin : 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nE + nM; t += gridDim.x * blockDim.x) {
        BucketEmit em;
        em.key = a.bktKey; em.val8 = a.bktVal8; em.val1 = a.bktVal1;
        em.rowTag = (a.tagBase + (unsigned)t) << 4;
        double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (t < nE) {
            const int e = a.eBegin + t;
            em.nv = 3; em.v[0] = a.elem[3 * e]; em.v[1] = a.elem[3 * e + 1]; em.v[2] = a.elem[3 * e + 2]; em.v[3] = -1;
            em.reserve(a.vtxCursor);
            const double w = a.h * a.vol[e] / 6.0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = i; j < 3; ++j) {
                    blk[0] = blk[4] = blk[8] = (i == j) ? 2.0 * w : -w;
                    em(i, j, blk);
                }
        }
        else {
            const int v = a.vBegin + (t - nE);
            const double m = a.mass[v];
            if (m == 0.0) continue;
            em.nv = 1; em.v[0] = v; em.v[1] = em.v[2] = em.v[3] = -1;
            em.reserve(a.vtxCursor);
            blk[0] = blk[4] = blk[8] = m;
            em(0, 0, blk);
        }
    }
}

// One thread per constraint row. PATH 0: four-vertex kinds (PT, EE and the three mollified kinds; 9x9 QL with the work
// store in shared memory), PATH 1: point-edge (6x6 QL), PATH 2: point-point (closed form).
// Rows are visited through a permutation that groups them by kind (stable, so memory locality of the constraint order is
// kept inside a kind): the lanes of a warp run the same distance formulas and the three launches see only their own
// rows. Outputs keep their positions (block offsets by row index), so the results do not depend on the visiting order.
// Threads per block / resident blocks per SM by path. PATH 0 is limited by the 99-word shared work store per row
// (792 B): 2 blocks x 128 rows = 8 warps per SM at 255 registers (9 warps would need 96-thread blocks, but registers are
// allocated per 4 warps, which caps them at 168 and spills 1.4 KB per row).
template <int PATH> struct BarrierCfg { static constexpr int T = PATH == 0 ? IDP_BARRIER_T0 : 128; static constexpr int MINB = PATH == 0 ? IDP_BARRIER_B0 : (PATH == 1 ? 3 : 4); };
template <int PATH, bool WANT_E, bool WANT_G, bool WANT_H, bool PROJECT>
__global__ void __launch_bounds__(BarrierCfg<PATH>::T, BarrierCfg<PATH>::MINB) k_barrier(BarrierArgs a)
{
    extern __shared__ double sV[];
    double Eacc = 0;
    const long stride = (long)gridDim.x * blockDim.x;
    long j = a.jBegin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    Row4 rNext = j < a.jEnd ? a.rows[j] : Row4{0, 0, 0, 0};
    for (; j < a.jEnd; j += stride) {
        const Row4 r = rNext;
        if (j + stride < a.jEnd) rNext = a.rows[j + stride]; // in flight during this row's work
        const RowDec d = decode_row(r.a, r.b, r.c, r.d);
        V3 x[4], xr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = ldv4(a.xp, d.v[k]);
        if (PATH == 0 && (d.kind == K_EE_M || d.kind == K_PE_M || d.kind == K_PP_M)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) xr[k] = ldv4(a.x0p, d.v[k]);
        }
        RowOut out;
        QlStore<9, BarrierCfg<PATH>::T> V9{sV + threadIdx.x};
        QlStore<6, BarrierCfg<PATH>::T> V6{sV + threadIdx.x};
        BucketEmit em;
        em.key = a.bktKey; em.val8 = a.bktVal8; em.val1 = a.bktVal1; em.nv = d.nv; em.v[0] = d.v[0]; em.v[1] = d.v[1]; em.v[2] = d.v[2]; em.v[3] = d.v[3]; em.rowTag = (unsigned)j << 4; // origin tag: any unique, reproducible id
        if (WANT_H) em.reserve(a.vtxCursor);
        const bool ok = row_eval<PATH>(d, x, xr, a.weights[j], a.dHat2, a.kappa, a.xi2, PROJECT, WANT_H, V9, V6, out, em);
        if (!ok) { atomicAdd(a.errDist, 1ull); continue; }
        if (WANT_H && out.eigFail) atomicAdd(a.errEig, 1ull);
        if (WANT_E) Eacc += out.E;
        if (WANT_G) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < d.nv) {
                    double* gp = a.g + 3 * (long)d.v[k];
                    atomicAdd(gp, out.g[3 * k]); atomicAdd(gp + 1, out.g[3 * k + 1]); atomicAdd(gp + 2, out.g[3 * k + 2]);
                }
            }
        }
    }
    if (WANT_E) {
        typedef cub::BlockReduce<double, BarrierCfg<PATH>::T> BR;
        __shared__ typename BR::TempStorage tmp;
        const double s = BR(tmp).Sum(Eacc);
        if (threadIdx.x == 0) a.partialE[blockIdx.x] = s;
    }
}

template <int PATH>
static int launch_barrier_path(idp_ctx* c, BarrierArgs a, unsigned grid, int sel)
{
    constexpr int T = BarrierCfg<PATH>::T;
    const size_t smem = PATH == 0 ? QlStore<9, T>::WORDS * sizeof(double) * T : (PATH == 1 ? QlStore<6, T>::WORDS * sizeof(double) * T : 0);
#define IDP_BARRIER_CASE1(E, G, H, P)                                                                                     \
    do {                                                                                                                  \
        if (smem) IDP_CK(c, cudaFuncSetAttribute(k_barrier<PATH, E, G, H, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        IDP_LAUNCH(c, (k_barrier<PATH, E, G, H, P>), grid, T, smem, a);                                               \
    } while (0)
#define IDP_BARRIER_CASE(E, G, H)                                                                                         \
    do {                                                                                                                  \
        if (H && a.projectSPD) IDP_BARRIER_CASE1(E, G, H, true);                                                          \
        else IDP_BARRIER_CASE1(E, G, H, false);                                                                           \
    } while (0)
    switch (sel) {
    case 1: IDP_BARRIER_CASE(true, false, false); break;
    case 2: IDP_BARRIER_CASE(false, true, false); break;
    case 3: IDP_BARRIER_CASE(true, true, false); break;
    case 4: IDP_BARRIER_CASE(false, false, true); break;
    case 5: IDP_BARRIER_CASE(true, false, true); break;
    case 6: IDP_BARRIER_CASE(false, true, true); break;
    case 7: IDP_BARRIER_CASE(true, true, true); break;
    default: break;
    }
#undef IDP_BARRIER_CASE
#undef IDP_BARRIER_CASE1
    IDP_CK(c, cudaGetLastError());
    return IDP_OK;
}

// deterministic final sum of the per-block partials (single block)
__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ p, int n, double* __restrict__ out)
{
    double s = 0;
    for (int i = threadIdx.x; i < n; i += 256) s += p[i];
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double t = BR(tmp).Sum(s);
    if (threadIdx.x == 0) *out = t;
}

int barrier_eval(idp_ctx* c, double dhat2, double kappa, double thickness, int want_e, int want_g, int want_h,
    int project_spd, double* E_out)
{
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    if (E_out) *E_out = 0;
    if (want_g) {
        IDP_CK(c, c->gbuf.reserve(3 * (size_t)c->nV));
        IDP_CK(c, cudaMemsetAsync(c->gbuf.p, 0, 3 * (size_t)c->nV * sizeof(double), c->stream));
    }
    if (want_h) { c->nnz = 0; c->nBlocksUnique = 0; c->nBlocksEmitted = 0; }
    const bool extraTerms = want_h && (c->nFlowElem > 0 || c->haveMass || c->nMem > 0 || c->nHinge > 0 || (c->nFricActive > 0 && c->fricMu > 0 && c->have_xn)); // the other terms are assembled even without contact rows
    if (c->nRows == 0 && !extraTerms) return IDP_OK;
    StageTimer tm(c, IDP_STAGE_BARRIER);
    IDP_CK(c, cudaMemsetAsync(c->counters.p + CNT_ERR_DIST, 0, sizeof(long long), c->stream));
    IDP_CK(c, cudaMemsetAsync(c->counters.p + CNT_ERR_EIG, 0, sizeof(long long), c->stream));
    // this rank's rows (row_owner) in an order grouped by kind
    const int ownRanks = c->rowsLocal ? 1 : c->nranks; // LOCAL-ROWS mode: every row in c->rows is this rank's
    int ownerShift = 8;
    while (ownerShift < 14 && (c->nV >> ownerShift) > 16 * c->nranks) ++ownerShift;
    if (c->nRows == 0) { for (int k = 0; k < 8; ++k) c->kindCount[k] = 0; }
    else if (!c->permValid) {
        IDP_CK(c, c->rowKind.reserve(c->nRows)); IDP_CK(c, c->rowKindSorted.reserve(c->nRows));
        IDP_CK(c, c->rowIota.reserve(c->nRows)); IDP_CK(c, c->rowPerm.reserve(c->nRows));
        unsigned long long* dKind = (unsigned long long*)(c->counters.p + CNT_KINDS);
        IDP_CK(c, cudaMemsetAsync(dKind, 0, 8 * sizeof(long long), c->stream));
        IDP_LAUNCH(c, k_row_kinds, std::min(blocks_for(c->nRows, 256), (unsigned)c->sm_count * 16), 256, 0, c->rows.p, c->nRows, c->rank, ownRanks, ownerShift,
            c->rowKind.p, c->rowIota.p, dKind);
        size_t bytes = 0;
        IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->rowKind.p, c->rowKindSorted.p, c->rowIota.p, c->rowPerm.p, (int)c->nRows, 0, 3, c->stream));
        IDP_CK(c, c->cubTemp.reserve(bytes));
        IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, c->rowKind.p, c->rowKindSorted.p, c->rowIota.p, c->rowPerm.p, (int)c->nRows, 0, 3, c->stream));
        ++c->lib_launches;
        long long hk[8];
        IDP_CK(c, cudaMemcpyAsync(hk, dKind, sizeof(hk), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 8; ++k) c->kindCount[k] = hk[k];
        const long nOwn = c->nRows - c->kindCount[7];
        IDP_CK(c, c->rowsK.reserve(std::max<long>(nOwn, 1))); IDP_CK(c, c->weightsK.reserve(std::max<long>(nOwn, 1)));
        if (nOwn > 0)
            IDP_LAUNCH(c, k_gather_rows, std::min(blocks_for(nOwn, 256), (unsigned)c->sm_count * 16), 256, 0, c->rows.p, c->weights.p, c->rowPerm.p, nOwn, c->rowsK.p, c->weightsK.p);
        c->permValid = true;
    }
    const long nPath[3] = {c->kindCount[K_EE] + c->kindCount[K_EE_M] + c->kindCount[K_PE_M] + c->kindCount[K_PP_M] + c->kindCount[K_PT],
        c->kindCount[K_PE], c->kindCount[K_PP]};
    const long nMine = nPath[0] + nPath[1] + nPath[2];
    if (!c->have_x0 && c->kindCount[K_EE_M] + c->kindCount[K_PE_M] + c->kindCount[K_PP_M] > 0)
        return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "mollified rows need the rest positions (idp_set_rest_positions)", __FILE__, __LINE__);
    const unsigned grid = std::max(1u, std::min(blocks_for(std::max(nPath[0], std::max(nPath[1], nPath[2])), 128), (unsigned)c->sm_count * 16));
    BarrierArgs a;
    a.rows = c->rowsK.p; a.weights = c->weightsK.p; a.jBegin = 0; a.jEnd = 0;
    a.xp = c->xp.p; a.x0p = c->x0p.p;
    a.dHat2 = dhat2 + 2 * std::sqrt(dhat2) * thickness; // IPC.h:757
    a.kappa = kappa; a.xi2 = thickness * thickness; a.projectSPD = project_spd;
    IDP_CK(c, c->red.reserve(3 * (size_t)grid + 8));
    a.partialE = c->red.p;
    a.g = c->gbuf.p;
    a.errDist = (unsigned long long*)(c->counters.p + CNT_ERR_DIST);
    a.errEig = (unsigned long long*)(c->counters.p + CNT_ERR_EIG);
    a.vtxOff = nullptr; a.vtxCursor = nullptr; a.bktKey = nullptr; a.bktVal8 = nullptr; a.bktVal1 = nullptr;
    long nBlocks = 0, nExtra = 0, nElastic = 0, nFricRows = 0;
    ExtraArgs xaKeep = {};
    if (want_h) {
        c->csrProjected = false;
        if (c->nRows + (long)c->nFlowElem + c->nV + c->nMem + c->nHinge + c->nFric >= (1L << 28)) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "more than 2^28 constraint rows (+ elements + vertices) on one rank", __FILE__, __LINE__);
        const size_t nV1 = (size_t)c->nV + 1;
        IDP_CK(c, c->vtxCnt.reserve(nV1)); IDP_CK(c, c->vtxOff.reserve(nV1)); IDP_CK(c, c->vtxCursor.reserve(nV1));
        IDP_CK(c, cudaMemsetAsync(c->vtxCnt.p, 0, nV1 * sizeof(int), c->stream));
        if (c->nRows > 0)
            IDP_LAUNCH(c, k_vertex_block_counts, std::min(blocks_for(c->nRows, 256), (unsigned)c->sm_count * 16), 256, 0, c->rows.p, c->nRows, c->rank, ownRanks, ownerShift, c->vtxCnt.p);
        // the other terms of the system matrix (flow Laplacian, lumped mass) share the buckets; sharded by element / vertex range
        ExtraArgs xa;
        xa.elem = c->flowElem.p; xa.vol = c->flowVol.p; xa.h = c->flowH; xa.mass = c->haveMass ? c->massDiag.p : nullptr;
        xa.eBegin = (int)((long)c->nFlowElem * c->rank / c->nranks); xa.eEnd = (int)((long)c->nFlowElem * (c->rank + 1) / c->nranks);
        xa.vBegin = (int)((long)c->nV * c->rank / c->nranks); xa.vEnd = (int)((long)c->nV * (c->rank + 1) / c->nranks);
        xa.vtxCnt = c->vtxCnt.p;
        nExtra = (xa.eEnd - xa.eBegin) + (xa.mass ? xa.vEnd - xa.vBegin : 0);
        if (nExtra > 0) IDP_LAUNCH(c, k_extra_counts, std::min(blocks_for(nExtra, 256), (unsigned)c->sm_count * 16), 256, 0, xa);
        xaKeep = xa;
        // membrane triangles and bending hinges (elastic_kernels.cu)
        if (c->nMem > 0 || c->nHinge > 0) IDP_TRY(elastic_block_counts(c, c->vtxCnt.p, &nElastic));
        IDP_TRY(friction_block_counts(c, c->vtxCnt.p, &nFricRows)); // lagged friction rows (friction_kernels.cu)
        IDP_TRY(cub_scan_exclusive(c, c->vtxCnt.p, c->vtxOff.p, (long)nV1));
        int total = 0;
        IDP_CK(c, cudaMemcpyAsync(&total, c->vtxOff.p + c->nV, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        IDP_CK(c, cudaMemcpyAsync(c->vtxCursor.p, c->vtxOff.p, nV1 * sizeof(int), cudaMemcpyDeviceToDevice, c->stream)); // cursors start at the bucket offsets
        nBlocks = total; // only this rank's rows are counted
        IDP_CK(c, c->bktKey.reserve(std::max<long>(nBlocks, 1)));
        IDP_CK(c, c->bktVal8.reserve(8 * (size_t)std::max<long>(nBlocks, 1)));
        IDP_CK(c, c->bktVal1.reserve(std::max<long>(nBlocks, 1) + 2)); // + 2: the bulk copies read whole 16-byte pairs
        a.vtxOff = c->vtxOff.p; a.vtxCursor = c->vtxCursor.p; a.bktKey = c->bktKey.p; a.bktVal8 = c->bktVal8.p; a.bktVal1 = c->bktVal1.p;
        c->nBlocksUnique = 0;
        c->nnz = 0;
        c->nBlocksEmitted = nBlocks;
    }
    if (nMine > 0) {
        KernelTimer kt(c, IDP_STAGE_K_BARRIER);
        const int sel = (want_e ? 1 : 0) | (want_g ? 2 : 0) | (want_h ? 4 : 0);
        BarrierArgs a0 = a, a1 = a, a2 = a;
        a0.jBegin = 0; a0.jEnd = nPath[0];
        a1.jBegin = nPath[0]; a1.jEnd = nPath[0] + nPath[1];
        a2.jBegin = nPath[0] + nPath[1]; a2.jEnd = nPath[0] + nPath[1] + nPath[2];
        a1.partialE = a.partialE + grid;
        a2.partialE = a.partialE + 2 * (size_t)grid;
        if (want_e) IDP_CK(c, cudaMemsetAsync(c->red.p, 0, 3 * (size_t)grid * sizeof(double), c->stream));
        const unsigned g0 = std::max(1u, std::min(blocks_for(nPath[0], BarrierCfg<0>::T), grid)), g1 = std::max(1u, std::min(blocks_for(nPath[1], 128), grid)),
                       g2 = std::max(1u, std::min(blocks_for(nPath[2], 128), grid));
        if (nPath[0] > 0) IDP_TRY(launch_barrier_path<0>(c, a0, g0, sel));
        if (nPath[1] > 0) IDP_TRY(launch_barrier_path<1>(c, a1, g1, sel));
        if (nPath[2] > 0) IDP_TRY(launch_barrier_path<2>(c, a2, g2, sel));
    }
    if (want_h && nExtra > 0) {
        xaKeep.vtxCursor = c->vtxCursor.p; xaKeep.bktKey = c->bktKey.p; xaKeep.bktVal8 = c->bktVal8.p; xaKeep.bktVal1 = c->bktVal1.p;
        xaKeep.tagBase = (unsigned)nMine;
        IDP_LAUNCH(c, k_extra_emit, std::min(blocks_for(nExtra, 256), (unsigned)c->sm_count * 16), 256, 0, xaKeep);
    }
    if (want_h && nElastic > 0) IDP_TRY(elastic_emit_blocks(c, project_spd, (unsigned)(nMine + nExtra), c->vtxCursor.p, c->bktKey.p, c->bktVal8.p, c->bktVal1.p));
    if (want_h && nFricRows > 0) IDP_TRY(friction_emit_blocks(c, (unsigned)(nMine + nExtra + nElastic), c->vtxCursor.p, c->bktKey.p, c->bktVal8.p, c->bktVal1.p));
    if (nMine > 0 && want_e) IDP_LAUNCH(c, k_sum_partials, 1, 256, 0, c->red.p, 3 * (int)grid, c->red.p + 3 * (size_t)grid);
    long long nerr = 0, neig = 0;
    IDP_CK(c, cudaMemcpyAsync(&nerr, c->counters.p + CNT_ERR_DIST, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaMemcpyAsync(&neig, c->counters.p + CNT_ERR_EIG, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    double E = 0;
    if (want_e && nMine > 0) IDP_CK(c, cudaMemcpyAsync(&E, c->red.p + 3 * (size_t)grid, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    if (nerr) return fail(c, IDP_ERR_NONPOSITIVE_DISTANCE, "%s (%s:%d)", "non-positive distance detected during barrier evaluation", __FILE__, __LINE__);
    if (neig) return fail(c, IDP_ERR_EIGEN_NO_CONVERGENCE, "%s (%s:%d)", "PSD projection: QL iteration cap reached", __FILE__, __LINE__);
    if (E_out) *E_out = E;
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// assembly (replaces Eigen's setFromTriplets, Math/CSR_MATRIX.h:49-56): the blocks arrive bucketed by lower vertex
// (k_barrier / BucketEmit). One warp per vertex sorts its bucket's keys (vhi, origin) in shared memory, finds the unique
// upper neighbours and sums the duplicates in origin order -- deterministic although the slots inside a bucket were handed
// out by atomics -- reading the 3x3 values from the bucket's own contiguous region. No global sort of the ~10 blocks per
// row is needed (the first version radix-sorted 200 M 64-bit keys and then gathered 72-byte values from all over HBM).
// The unique upper blocks are mirrored (a small stable sort by column vertex) and every block row is written as
// [mirrors ascending][diagonal + upper ascending] -> exactly Eigen's pattern: columns ascending, duplicates summed.
// ------------------------------------------------------------------------------------------------------------
// ascending bitonic network without direction flags ("flip" merge): elements past n behave as +inf and are never touched
template <class SyncF>
__device__ __forceinline__ void bitonic_sort_kp(unsigned long long* key, unsigned* pay, int n, int tid, int nthreads, SyncF sync)
{
    int P = 1;
    while (P < n) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        const int half = k >> 1;
        for (int t = tid; t < (P >> 1); t += nthreads) { // flip stage: i <-> mirror inside its block of k
            const int grp = t / half, idx = t - grp * half;
            const int i = grp * k + idx, l = grp * k + (k - 1 - idx);
            if (l < n) {
                const unsigned long long a = key[i], b = key[l];
                if (a > b) { key[i] = b; key[l] = a; const unsigned pa = pay[i]; pay[i] = pay[l]; pay[l] = pa; }
            }
        }
        sync();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nthreads) {
                const int i = 2 * j * (t / j) + (t % j), l = i + j;
                if (l < n) {
                    const unsigned long long a = key[i], b = key[l];
                    if (a > b) { key[i] = b; key[l] = a; const unsigned pa = pay[i]; pay[i] = pay[l]; pay[l] = pa; }
                }
            }
            sync();
        }
    }
}
struct ReduceArgs {
    const int* vtxOff; const unsigned long long* bktKey; const double* bktVal8; const double* bktVal1;
    int nV;
    int* uCount; int* uCol; double* uVal; int* lowerCount; // unique blocks of vertex v at sparse slots vtxOff[v] + u
    int* bigList; int* bigCount;                           // vertices whose bucket exceeds the shared-memory capacity
    unsigned long long* bigKey; unsigned* bigPay; unsigned* bigStart; // global scratch for those (aligned with the bucket slots)
};
// after the sort: unique upper neighbours and their sums. key/pay/ustart may live in shared or global memory.
template <class SyncF>
__device__ __forceinline__ void reduce_sorted_bucket(const ReduceArgs& a, int v, int off, int n, const unsigned long long* key, const unsigned* pay,
    unsigned* ustart, int tid, int nthreads, int* sCounter, SyncF sync)
{
    // heads of the runs of equal vhi -> ustart[u] (order preserved: a block-wide pass in chunks with a running base)
    if (tid == 0) *sCounter = 0;
    sync();
    for (int p0 = 0; p0 < n; p0 += nthreads) {
        const int p = p0 + tid;
        const bool head = p < n && (p == 0 || (unsigned)(key[p] >> 32) != (unsigned)(key[p - 1] >> 32));
        // order-preserving compaction inside the chunk: warp ballots + per-warp bases taken in warp order
        const unsigned m = __ballot_sync(0xffffffffu, head);
        const int lane = tid & 31, wid = tid >> 5, nw = (nthreads + 31) >> 5;
        for (int w = 0; w < nw; ++w) { // warps take their turn (nw = 1 for the warp-per-vertex path)
            if (w == wid) {
                int base = 0;
                if (lane == 0) { base = *sCounter; *sCounter = base + __popc(m); }
                base = __shfl_sync(0xffffffffu, base, 0);
                if (head) ustart[base + __popc(m & ((1u << lane) - 1u))] = (unsigned)p;
            }
            sync();
        }
    }
    const int U = *sCounter;
    sync();
    for (int t = tid; t < 9 * U; t += nthreads) {
        const int u = t / 9, comp = t - 9 * u;
        const int s0 = (int)ustart[u], s1 = (u + 1 < U) ? (int)ustart[u + 1] : n;
        double sum = 0;
        for (int p = s0; p < s1; ++p) {
            const long slot = (long)off + pay[p];
            sum += comp < 8 ? a.bktVal8[8 * slot + comp] : a.bktVal1[slot];
        }
        a.uVal[9 * ((long)off + u) + comp] = sum;
        if (comp == 0) {
            const int vj = (int)(unsigned)(key[s0] >> 32);
            a.uCol[off + u] = vj;
            if (vj != v) atomicAdd(&a.lowerCount[vj], 1);
        }
    }
    if (tid == 0) a.uCount[v] = U;
    sync();
}
#ifndef IDP_REDUCE_WARPS
#define IDP_REDUCE_WARPS 4
#endif
#define IDP_REDUCE_CAP 512 // bucket entries a warp sorts in registers (16 per lane)
#ifndef IDP_REDUCE_STAGE
#define IDP_REDUCE_STAGE 160 // bucket entries whose 3x3 values a warp stages in shared memory (72 B each)
#endif
// Warp-level bitonic sort of 32 E keys held E per lane (element index = lane * E + r). Stages with partner distance < E
// are compare-exchanges between registers of one lane (static indices), the others exchange with lane ^ (distance / E)
// through shuffles. Only the in-lane networks are unrolled; the merge levels and their cross-lane stages are rolled loops
// with a run-time lane mask -- the fully unrolled network (three sizes of it) did not fit the instruction cache and the
// kernel stalled on instruction fetch (ncu: no_instruction 8.5 of 14 cycles per issue).
// PAY = false: the keys are unique and carry their own payload (packed slot), compare-exchange is a plain min / max.
__device__ __forceinline__ unsigned long long selp64(bool p, unsigned long long a, unsigned long long b)
{
    unsigned long long r; // branch-free (the compiler turned 64-bit ternaries into divergent branches)
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.b64 %0, %1, %2, q; }" : "=l"(r) : "l"(a), "l"(b), "r"((unsigned)p));
    return r;
}
__device__ __forceinline__ unsigned selp32(bool p, unsigned a, unsigned b)
{
    unsigned r;
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.b32 %0, %1, %2, q; }" : "=r"(r) : "r"(a), "r"(b), "r"((unsigned)p));
    return r;
}
template <int E, bool PAY>
__device__ __forceinline__ void lane_stages(unsigned long long (&key)[E], unsigned (&pay)[E], int kFrom, bool ascLane)
{
    // the in-lane stages j = kFrom/2 .. 1 of one merge level; direction: static (r & k) inside the first levels (k < E),
    // otherwise the lane's direction
#pragma unroll
    for (int j = E >> 1; j > 0; j >>= 1) {
        if (j >= kFrom) continue;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const int p = r ^ j;
            if (p > r) {
                const bool asc = kFrom < E ? ((r & kFrom) == 0) : ascLane;
                const bool sw = (key[r] > key[p]) == asc;
                const unsigned long long ka = key[r], kb = key[p];
                key[r] = selp64(sw, kb, ka); key[p] = selp64(sw, ka, kb);
                if (PAY) { const unsigned pa = pay[r], pb = pay[p]; pay[r] = selp32(sw, pb, pa); pay[p] = selp32(sw, pa, pb); }
            }
        }
    }
}
template <int E, bool PAY>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&key)[E], unsigned (&pay)[E], int lane)
{
#pragma unroll
    for (int k = 2; k <= E; k <<= 1) lane_stages<E, PAY>(key, pay, k, (lane & 1) == 0); // levels inside a lane (k == E: by lane parity)
#pragma unroll 1
    for (int kk = 2; kk <= 32; kk <<= 1) { // levels of kk lanes = kk E elements
        const bool asc = (lane & kk) == 0;
#pragma unroll 1
        for (int lj = kk >> 1; lj > 0; lj >>= 1) {
            const bool keepMin = ((lane & lj) == 0) == asc;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key[r], lj);
                const bool take = PAY ? ((ok < key[r]) == keepMin && ok != key[r]) : ((ok < key[r]) == keepMin);
                if (PAY) { const unsigned op = __shfl_xor_sync(0xffffffffu, pay[r], lj); pay[r] = selp32(take, op, pay[r]); }
                key[r] = selp64(take, ok, key[r]);
            }
        }
        lane_stages<E, PAY>(key, pay, E, asc);
    }
}
// ---- bulk-copy (TMA, 1-D) staging of a bucket's values: one elected lane arms the warp's mbarrier with the byte count and
// issues cp.async.bulk for the 64-byte and the 8-byte parts; the copy runs while the warp sorts the keys in registers.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
struct alignas(128) ReduceSmem { // one per warp
    double val8[8 * IDP_REDUCE_STAGE];       // 64-byte parts, entry e at [8 e, 8 e + 8)
    double val1[IDP_REDUCE_STAGE + 2];       // ninth components (the source is only 8-byte aligned: copied from the even slot below)
    unsigned short pay[IDP_REDUCE_CAP];      // sorted position -> bucket slot
    unsigned short start[IDP_REDUCE_CAP];    // unique block -> first sorted position
    unsigned long long bar;
};
struct SharedF64 { // explicit shared-window loads (a generic pointer would compile to LD instead of LDS)
    unsigned base;
    __device__ __forceinline__ double2 ld2(int i) const { double2 r; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(base + 8u * (unsigned)i)); return r; }
    __device__ __forceinline__ double ld1(int i) const { double r; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(base + 8u * (unsigned)i)); return r; }
};
struct GlobalF64 {
    const double* base;
    __device__ __forceinline__ double2 ld2(int i) const { return *reinterpret_cast<const double2*>(base + i); }
    __device__ __forceinline__ double ld1(int i) const { return base[i]; }
};
// Sums of the duplicates in a fixed association of the sorted (= origin) order. Eight groups of four lanes; a lane adds
// two components of the 64-byte part (16-byte loads) of every entry its group visits and the ninth component of every
// fourth one. The first run (the diagonal block: every row at this vertex contributes) is strided over all eight groups
// and combined by a butterfly; the other runs go to one group each. v8 / v1: the bucket's values (shared or global).
template <class P8, class P1>
__device__ __forceinline__ void bucket_sums(const ReduceSmem& sm, P8 v8, P1 v1, int U, int n, int lane, double* outp)
{
    const int grp = lane >> 2, sub = lane & 3, c2 = 2 * sub;
    {
        const int s1 = U > 1 ? (int)sm.start[1] : n; // run 0 = [0, s1)
        double ax = 0, ay = 0, az = 0;
        for (int p = grp; p < s1; p += 8) {
            const int sl = sm.pay[p];
            const double2 x = v8.ld2(8 * sl + c2);
            ax += x.x; ay += x.y;
            if (sub == 0) az += v1.ld1(sl);
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); az += __shfl_xor_sync(0xffffffffu, az, o);
        }
        if (grp == 0) { outp[c2] = ax; outp[c2 + 1] = ay; if (sub == 0) outp[8] = az; }
    }
    for (int u0 = 1; u0 < U; u0 += 8) { // warp-uniform trip count: the shuffles below need all lanes
        const int u = u0 + grp;
        int s0 = 0, s1 = 0;
        if (u < U) { s0 = sm.start[u]; s1 = (u + 1 < U) ? (int)sm.start[u + 1] : n; }
        double ax = 0, ay = 0, az = 0;
        int p = s0;
        for (; p + 4 <= s1; p += 4) {
            const int l0 = sm.pay[p], l1 = sm.pay[p + 1], l2 = sm.pay[p + 2], l3 = sm.pay[p + 3];
            const double2 x0 = v8.ld2(8 * l0 + c2), x1 = v8.ld2(8 * l1 + c2);
            const double2 x2 = v8.ld2(8 * l2 + c2), x3 = v8.ld2(8 * l3 + c2);
            const double z = v1.ld1(sub == 0 ? l0 : (sub == 1 ? l1 : (sub == 2 ? l2 : l3)));
            ax = (((ax + x0.x) + x1.x) + x2.x) + x3.x; ay = (((ay + x0.y) + x1.y) + x2.y) + x3.y;
            az += z;
        }
        for (int q = 0; p < s1; ++p, ++q) {
            const int sl = sm.pay[p];
            const double2 x = v8.ld2(8 * sl + c2);
            ax += x.x; ay += x.y;
            if (sub == q) az += v1.ld1(sl);
        }
        az += __shfl_xor_sync(0xffffffffu, az, 1);
        az += __shfl_xor_sync(0xffffffffu, az, 2);
        if (u < U) { outp[9 * u + c2] = ax; outp[9 * u + c2 + 1] = ay; if (sub == 0) outp[9 * u + 8] = az; }
    }
}
// one vertex: sort its bucket by (vhi, origin), find the unique upper neighbours, sum the duplicates in origin order
// PACKED (nV <= 2^23): the sort key is (vhi : 23 | origin : 32 | slot : 9), unique, no payload registers.
template <int E, bool PACKED>
__device__ __forceinline__ void vertex_reduce_warp(const ReduceArgs& a, int v, int off, int n, int lane, ReduceSmem& sm, unsigned& phase)
{
    const bool staged = n <= IDP_REDUCE_STAGE;
    const int off1 = off & ~1; // bulk copies need 16-byte aligned sources
    if (staged && lane == 0) {
        const unsigned b8 = 64u * (unsigned)n, b1 = 8u * (unsigned)(((off + n + 1) & ~1) - off1);
        mbar_expect_tx(&sm.bar, b8 + b1);
        bulk_g2s(sm.val8, a.bktVal8 + 8 * (long)off, b8, &sm.bar);
        bulk_g2s(sm.val1, a.bktVal1 + off1, b1, &sm.bar);
    }
    unsigned long long key[E];
    unsigned pay[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int e = lane * E + r;
        unsigned long long k = ~0ull; // padding sorts to the end
        if (e < n) {
            k = a.bktKey[off + e];
            if (PACKED) k = ((k >> 32) << 41) | ((k & 0xffffffffull) << 9) | (unsigned long long)e;
        }
        key[r] = k;
        pay[r] = (unsigned)e;
    }
    warp_bitonic_sort<E, !PACKED>(key, pay, lane);
    // heads of the runs of equal vhi
    constexpr int VSH = PACKED ? 41 : 32;
    const unsigned prevLast = __shfl_up_sync(0xffffffffu, (unsigned)(key[E - 1] >> VSH), 1);
    unsigned heads = 0;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int e = lane * E + r;
        const unsigned vj = (unsigned)(key[r] >> VSH), pv = r ? (unsigned)(key[r - 1] >> VSH) : prevLast;
        if (e < n && (e == 0 || vj != pv)) heads |= 1u << r;
        if (e < n) sm.pay[e] = PACKED ? (unsigned short)((unsigned)key[r] & 511u) : (unsigned short)pay[r];
    }
    int base = __popc(heads); // exclusive prefix of the head counts over the lanes
    const int mine = base;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, base, o);
        if (lane >= o) base += t;
    }
    const int U = __shfl_sync(0xffffffffu, base, 31);
    base -= mine;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        if (heads & (1u << r)) {
            const int u = base + __popc(heads & ((1u << r) - 1u));
            const int vj = (int)(unsigned)(key[r] >> VSH);
            sm.start[u] = (unsigned short)(lane * E + r);
            a.uCol[off + u] = vj;
            if (vj != v) atomicAdd(&a.lowerCount[vj], 1);
        }
    }
    if (lane == 0) a.uCount[v] = U;
    __syncwarp();
    double* outp = a.uVal + 9 * (long)off;
    if (staged) {
        mbar_wait(&sm.bar, phase);
        phase ^= 1u;
        bucket_sums(sm, SharedF64{smem_u32(sm.val8)}, SharedF64{smem_u32(sm.val1 + (off - off1))}, U, n, lane, outp);
    }
    else bucket_sums(sm, GlobalF64{a.bktVal8 + 8 * (long)off}, GlobalF64{a.bktVal1 + off}, U, n, lane, outp);
    __syncwarp(); // every lane is done with the staged values before the next bulk copy is issued
}
template <bool PACKED>
__global__ void __launch_bounds__(32 * IDP_REDUCE_WARPS) k_vertex_reduce(ReduceArgs a)
{
    extern __shared__ __align__(128) unsigned char sRaw[];
    ReduceSmem* smAll = reinterpret_cast<ReduceSmem*>(sRaw);
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ReduceSmem& sm = smAll[wib];
    if (lane == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;
    for (int v = blockIdx.x * IDP_REDUCE_WARPS + wib; v < a.nV; v += gridDim.x * IDP_REDUCE_WARPS) {
        const int off = a.vtxOff[v], n = a.vtxOff[v + 1] - off;
        if (n == 0) { if (lane == 0) a.uCount[v] = 0; continue; }
        if (n > IDP_REDUCE_CAP) { if (lane == 0) { a.bigList[atomicAdd(a.bigCount, 1)] = v; a.uCount[v] = 0; } }
        else if (PACKED) {
            if (n <= 128) vertex_reduce_warp<4, true>(a, v, off, n, lane, sm, phase);
            else if (n <= 256) vertex_reduce_warp<8, true>(a, v, off, n, lane, sm, phase);
            else vertex_reduce_warp<16, true>(a, v, off, n, lane, sm, phase);
        }
        else vertex_reduce_warp<16, false>(a, v, off, n, lane, sm, phase); // > 2^23 vertices: (key, payload) pairs, one size
    }
}
// oversized buckets: one CTA per vertex, the same network on global scratch (correct for any size; only pathological
// inputs -- one vertex within dHat of thousands of primitives -- get here)
__global__ void __launch_bounds__(256) k_vertex_reduce_big(ReduceArgs a)
{
    __shared__ int sCnt;
    auto sync = [] { __threadfence_block(); __syncthreads(); };
    const int nBig = *a.bigCount;
    for (int b = blockIdx.x; b < nBig; b += gridDim.x) {
        const int v = a.bigList[b];
        const int off = a.vtxOff[v], n = a.vtxOff[v + 1] - off;
        unsigned long long* key = a.bigKey + off;
        unsigned* pay = a.bigPay + off;
        for (int p = threadIdx.x; p < n; p += blockDim.x) { key[p] = a.bktKey[off + p]; pay[p] = (unsigned)p; }
        sync();
        bitonic_sort_kp(key, pay, n, (int)threadIdx.x, (int)blockDim.x, sync);
        reduce_sorted_bucket(a, v, off, n, key, pay, a.bigStart + off, (int)threadIdx.x, (int)blockDim.x, &sCnt, sync);
    }
}
// compact list of the unique upper blocks in (row vertex, column vertex) order
__global__ void __launch_bounds__(256) k_compact_unique(const int* __restrict__ vtxOff, const int* __restrict__ uCount, const int* __restrict__ vtxBlkStart,
    const int* __restrict__ uCol, int nV, int* __restrict__ urow, int* __restrict__ ucol, int* __restrict__ usrc)
{
    const int lane = threadIdx.x & 31;
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < nV; v += (gridDim.x * blockDim.x) >> 5) {
        const int U = uCount[v], off = vtxOff[v], seg0 = vtxBlkStart[v];
        for (int u = lane; u < U; u += 32) { urow[seg0 + u] = v; ucol[seg0 + u] = uCol[off + u]; usrc[seg0 + u] = off + u; }
    }
}
// block-row layout: [lower blocks (mirrors, ascending column)] [upper blocks incl. diagonal (ascending column)]
// cnt[v] = lowerCount[v] + upperCount[v]
__global__ void k_row_totals(const int* __restrict__ uCount, const int* __restrict__ lowerCount, int nV, int* __restrict__ cnt)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= nV; v += gridDim.x * blockDim.x)
        cnt[v] = (v < nV) ? lowerCount[v] + uCount[v] : 0;
}
__global__ void k_csr_ptr(const int* __restrict__ rowStart, int nV, int* __restrict__ ptr)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= nV; v += gridDim.x * blockDim.x) {
        if (v == nV) { ptr[3 * (long)nV] = 9 * rowStart[nV]; continue; }
        const int s = rowStart[v], nb = rowStart[v + 1] - s;
        ptr[3 * (long)v] = 9 * s;
        ptr[3 * (long)v + 1] = 9 * s + 3 * nb;
        ptr[3 * (long)v + 2] = 9 * s + 6 * nb;
    }
}
// One warp per block row writes its three scalar rows: entry k of a row is component b = k % 3 of block slot k / 3, the first
// lowerCount[v] slots being the mirrored blocks (`order` lists the strictly-upper unique blocks stably sorted by their column
// vertex, so the members of one block row appear with ascending row vertex = ascending column in the mirror; lstart[v] = first
// entry of row v; stored transposed), the rest the diagonal + upper blocks in bucket order. Consecutive lanes write consecutive
// positions of all three rows (the first version ran one thread per (block, component): 24-byte pieces scattered over the rows,
// seven index loads per scalar).
__global__ void __launch_bounds__(256) k_write_rows(const double* __restrict__ uVal, const int* __restrict__ usrc, const int* __restrict__ urow,
    const int* __restrict__ ucol, const int* __restrict__ vtxBlkStart, const int* __restrict__ rowStart, const int* __restrict__ lowerCount,
    const int* __restrict__ order, const int* __restrict__ lstart, int nV, int* __restrict__ col, double* __restrict__ val)
{
    const int lane = threadIdx.x & 31;
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < nV; v += (gridDim.x * blockDim.x) >> 5) {
        const int bs = rowStart[v], nb = rowStart[v + 1] - bs, nl = lowerCount[v], ub = vtxBlkStart[v], le = lstart[v];
        const long base = 9L * bs;
        for (int k = lane; k < 3 * nb; k += 32) {
            const int slot = k / 3, b = k - 3 * slot;
            int cv;
            const double* blk;
            int s0, s1, s2; // offsets of component (a, b) for a = 0, 1, 2 inside the 9-value block
            if (slot < nl) {
                const int seg = order[le + slot];
                cv = urow[seg];
                blk = uVal + 9L * usrc[seg];
                s0 = 3 * b; s1 = 3 * b + 1; s2 = 3 * b + 2; // transposed
            }
            else {
                const int seg = ub + (slot - nl);
                cv = ucol[seg];
                blk = uVal + 9L * usrc[seg];
                s0 = b; s1 = 3 + b; s2 = 6 + b;
            }
            const double x0 = blk[s0], x1 = blk[s1], x2 = blk[s2];
            const int c3 = 3 * cv + b;
            const long p0 = base + k, p1 = p0 + 3L * nb, p2 = p1 + 3L * nb;
            col[p0] = c3; col[p1] = c3; col[p2] = c3;
            val[p0] = x0; val[p1] = x1; val[p2] = x2;
        }
    }
}
__global__ void k_lower_keys(const int* __restrict__ urow, const int* __restrict__ ucol, int nSeg, int* __restrict__ keyOut, int* __restrict__ segOut)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nSeg; s += gridDim.x * blockDim.x) {
        keyOut[s] = (urow[s] != ucol[s]) ? ucol[s] : 0x7fffffff; // diagonal blocks sort to the end and are ignored
        segOut[s] = s;
    }
}
__global__ void k_lower_starts(const int* __restrict__ sortedCol, long nLower, int nV, int* __restrict__ lstart)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= nV; v += gridDim.x * blockDim.x) {
        long lo = 0, hi = nLower;
        while (lo < hi) {
            const long mid = (lo + hi) >> 1;
            if (sortedCol[mid] < v) lo = mid + 1;
            else hi = mid;
        }
        lstart[v] = (int)lo;
    }
}

int assemble_csr(idp_ctx* c)
{
    if (c->pendCsr) IDP_CK(c, cudaStreamWaitEvent(c->stream, c->evVal, 0)); // a pending idp_get_hessian_csr_begin still reads the old CSR
    StageTimer tm(c, IDP_STAGE_CSR);
    IDP_CK(c, c->csrPtr.reserve(3 * (size_t)c->nV + 1));
    const long n = c->nBlocksEmitted;
    if (n <= 0) {
        IDP_CK(c, cudaMemsetAsync(c->csrPtr.p, 0, (3 * (size_t)c->nV + 1) * sizeof(int), c->stream));
        c->nnz = 0; c->nBlocksUnique = 0;
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        return IDP_OK;
    }
    const int nV = c->nV;
    const size_t nV2 = (size_t)nV + 2;
    // per-vertex sort / unique / sum
    IDP_CK(c, c->uCount.reserve(nV2)); IDP_CK(c, c->uCol.reserve(n)); IDP_CK(c, c->uVal.reserve(9 * (size_t)n));
    IDP_CK(c, c->lowerCount.reserve(nV2)); IDP_CK(c, c->lstart.reserve(nV2));
    IDP_CK(c, c->vtxBlkStart.reserve(nV2)); IDP_CK(c, c->rowStart.reserve(nV2)); IDP_CK(c, c->segId.reserve(nV2));
    IDP_CK(c, c->bigList.reserve(nV2)); IDP_CK(c, c->bigKey.reserve(n)); IDP_CK(c, c->bigPay.reserve(n)); IDP_CK(c, c->bigStart.reserve(n));
    IDP_CK(c, cudaMemsetAsync(c->lowerCount.p, 0, nV2 * sizeof(int), c->stream));
    IDP_CK(c, cudaMemsetAsync(c->uCount.p + nV, 0, 2 * sizeof(int), c->stream));
    int* dBigCount = c->bigList.p + nV + 1;
    IDP_CK(c, cudaMemsetAsync(dBigCount, 0, sizeof(int), c->stream));
    ReduceArgs ra;
    ra.vtxOff = c->vtxOff.p; ra.bktKey = c->bktKey.p; ra.bktVal8 = c->bktVal8.p; ra.bktVal1 = c->bktVal1.p; ra.nV = nV;
    ra.uCount = c->uCount.p; ra.uCol = c->uCol.p; ra.uVal = c->uVal.p; ra.lowerCount = c->lowerCount.p;
    ra.bigList = c->bigList.p; ra.bigCount = dBigCount; ra.bigKey = c->bigKey.p; ra.bigPay = c->bigPay.p; ra.bigStart = c->bigStart.p;
    const size_t reduceSmem = sizeof(ReduceSmem) * IDP_REDUCE_WARPS;
    const unsigned reduceGrid = std::min(blocks_for(nV, IDP_REDUCE_WARPS), (unsigned)c->sm_count * 64);
    if (nV <= (1 << 23) && !getenv("IDP_REDUCE_UNPACKED")) {
        IDP_CK(c, cudaFuncSetAttribute(k_vertex_reduce<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reduceSmem));
        IDP_LAUNCH(c, k_vertex_reduce<true>, reduceGrid, 32 * IDP_REDUCE_WARPS, reduceSmem, ra);
    }
    else {
        IDP_CK(c, cudaFuncSetAttribute(k_vertex_reduce<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reduceSmem));
        IDP_LAUNCH(c, k_vertex_reduce<false>, reduceGrid, 32 * IDP_REDUCE_WARPS, reduceSmem, ra);
    }
    IDP_LAUNCH(c, k_vertex_reduce_big, (unsigned)c->sm_count * 2, 256, 0, ra);
    IDP_TRY(cub_scan_exclusive(c, c->uCount.p, c->vtxBlkStart.p, (long)nV + 1));
    int nSeg = 0;
    IDP_CK(c, cudaMemcpyAsync(&nSeg, c->vtxBlkStart.p + nV, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    // block-row sizes and starts
    IDP_LAUNCH(c, k_row_totals, blocks_for(nV + 1, 256), 256, 0, c->uCount.p, c->lowerCount.p, nV, c->segId.p);
    IDP_TRY(cub_scan_exclusive(c, c->segId.p, c->rowStart.p, (long)nV + 1));
    int nBlkTotal = 0;
    IDP_CK(c, cudaMemcpyAsync(&nBlkTotal, c->rowStart.p + nV, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    IDP_CK(c, c->urow.reserve(std::max(nSeg, 1))); IDP_CK(c, c->ucol.reserve(std::max(nSeg, 1))); IDP_CK(c, c->usrc.reserve(std::max(nSeg, 1)));
    IDP_LAUNCH(c, k_compact_unique, std::min(blocks_for(32L * nV, 256), (unsigned)c->sm_count * 32), 256, 0, c->vtxOff.p, c->uCount.p, c->vtxBlkStart.p, c->uCol.p, nV,
        c->urow.p, c->ucol.p, c->usrc.p);
    // mirrored (lower) blocks: stable sort of the unique blocks by column vertex
    IDP_CK(c, c->lkey.reserve(std::max(nSeg, 1))); IDP_CK(c, c->lkeySorted.reserve(std::max(nSeg, 1)));
    IDP_CK(c, c->lseg.reserve(std::max(nSeg, 1))); IDP_CK(c, c->lsegSorted.reserve(std::max(nSeg, 1)));
    IDP_LAUNCH(c, k_lower_keys, blocks_for(nSeg, 256), 256, 0, c->urow.p, c->ucol.p, nSeg, c->lkey.p, c->lseg.p);
    size_t bytes = 0;
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->lkey.p, c->lkeySorted.p, c->lseg.p, c->lsegSorted.p, nSeg, 0, 31, c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, c->lkey.p, c->lkeySorted.p, c->lseg.p, c->lsegSorted.p, nSeg, 0, 31, c->stream));
    ++c->lib_launches;
    const long nLower = (long)nBlkTotal - nSeg; // every strictly-upper block has one mirror
    c->nBlocksUnique = nBlkTotal;
    c->nnz = 9L * nBlkTotal;
    IDP_CK(c, c->csrCol.reserve(std::max<long>(c->nnz, 1)));
    IDP_CK(c, c->csrVal.reserve(std::max<long>(c->nnz, 1)));
    IDP_LAUNCH(c, k_lower_starts, blocks_for(nV + 1, 256), 256, 0, c->lkeySorted.p, nLower, nV, c->lstart.p);
    IDP_LAUNCH(c, k_csr_ptr, blocks_for(nV + 1, 256), 256, 0, c->rowStart.p, nV, c->csrPtr.p);
    if (nSeg > 0)
        IDP_LAUNCH(c, k_write_rows, std::min(blocks_for(32L * nV, 256), (unsigned)c->sm_count * 32), 256, 0, c->uVal.p, c->usrc.p, c->urow.p, c->ucol.p, c->vtxBlkStart.p,
            c->rowStart.p, c->lowerCount.p, c->lsegSorted.p, c->lstart.p, nV, c->csrCol.p, c->csrVal.p);
    IDP_CK(c, cudaGetLastError());
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// FP64 pipe microbenchmark (roofline denominator): 8 independent DFMA chains per thread
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 1.2345) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace idp

extern "C" int idp_measure_fp64_tflops(idp_ctx* c, double* tflops)
{
    using namespace idp;
    const int iters = 1 << 15, blocks = c->sm_count * 8, threads = 256;
    IDP_CK(c, c->red.reserve((size_t)blocks * threads));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        IDP_CK(c, cudaEventRecord(c->ev0, c->stream));
        IDP_LAUNCH(c, k_dfma, blocks, threads, 0, c->red.p, iters, 0.999999, 1e-6);
        IDP_CK(c, cudaEventRecord(c->ev1, c->stream));
        IDP_CK(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        IDP_CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return IDP_OK;
}
