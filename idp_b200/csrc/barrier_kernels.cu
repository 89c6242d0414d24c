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
// number of 3x3 blocks a row contributes: nv(nv+1)/2 -- only blocks with vi <= vj are emitted, the assembly mirrors them
__global__ void k_row_block_counts(const Row4* __restrict__ rows, long n, int rank, int nranks, int shift, int* __restrict__ cnt)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (long)gridDim.x * blockDim.x) {
        int k = 0;
        if (i < n) {
            const Row4 r = rows[i];
            if (nranks == 1 || row_owner(r, shift, nranks) == rank)
                k = (r.a >= 0 || r.d >= 0) ? 10 : (r.c >= 0 ? 6 : 3); // upper triangle of the nv x nv blocks (IPC.h:1372-1387 counts all nv^2)
        }
        cnt[i] = k;
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

struct BarrierArgs {
    const Row4* rows; const double* weights; const int* perm; long jBegin, jEnd; // rows perm[jBegin..jEnd): one path, sorted by kind
    const double4* xp; const double4* x0p;
    double dHat2, kappa, xi2;
    int projectSPD;
    double* partialE;                 // one slot per block
    double* g;                        // 3*nV, xyz interleaved (atomics)
    const int* blkOff; unsigned long long* blkKey; int* blkIdx; double* blkVal; long long nVll;
    unsigned long long* errDist;
    unsigned long long* errEig;
};

// Hessian sink: the local Hessian is symmetric, so only one block per unordered vertex pair is stored, keyed (vlo, vhi)
// with vlo <= vhi; block (i,j) with v[i] > v[j] is stored transposed. Slot = upper-triangle index of (min(i,j), max(i,j)).
struct BlockEmit {
    unsigned long long* key; int* idx; double* val;
    long o; int nv; const int* v; long long nV;
    __device__ __forceinline__ bool wants(int i, int j) const { return i <= j; }
    __device__ __forceinline__ void operator()(int i, int j, const double* blk) const
    {
        const long s = o + (i * nv - (i * (i - 1)) / 2 + (j - i));
        const bool tr = v[i] > v[j];
        const long long vlo = tr ? v[j] : v[i], vhi = tr ? v[i] : v[j];
        key[s] = (unsigned long long)(vlo * nV + vhi);
        idx[s] = (int)s;
        double* dst = val + 9 * s;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) dst[3 * p + q] = tr ? blk[3 * q + p] : blk[3 * p + q];
    }
};

// One thread per constraint row. PATH 0: four-vertex kinds (PT, EE and the three mollified kinds; 9x9 QL with the work
// store in shared memory), PATH 1: point-edge (6x6 QL), PATH 2: point-point (closed form).
// Rows are visited through a permutation that groups them by kind (stable, so memory locality of the constraint order is
// kept inside a kind): the lanes of a warp run the same distance formulas and the three launches see only their own
// rows. Outputs keep their positions (block offsets by row index), so the results do not depend on the visiting order.
// Threads per block / resident blocks per SM by path. PATH 0 is limited by the 99-word shared work store per row
// (792 B): 2 blocks x 128 rows = 8 warps per SM at 255 registers (9 warps would need 96-thread blocks, but registers are
// allocated per 4 warps, which caps them at 168 and spills 1.4 KB per row).
template <int PATH> struct BarrierCfg { static constexpr int T = PATH == 0 ? IDP_BARRIER_T0 : 128; static constexpr int MINB = PATH == 0 ? IDP_BARRIER_B0 : (PATH == 1 ? 3 : 4); };
template <int PATH, bool WANT_E, bool WANT_G, bool WANT_H>
__global__ void __launch_bounds__(BarrierCfg<PATH>::T, BarrierCfg<PATH>::MINB) k_barrier(BarrierArgs a)
{
    extern __shared__ double sV[];
    double Eacc = 0;
    for (long j = a.jBegin + (long)blockIdx.x * blockDim.x + threadIdx.x; j < a.jEnd; j += (long)gridDim.x * blockDim.x) {
        const long i = a.perm[j];
        const Row4 r = a.rows[i];
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
        BlockEmit em{a.blkKey, a.blkIdx, a.blkVal, WANT_H ? (long)a.blkOff[i] : 0L, d.nv, d.v, a.nVll};
        const bool ok = row_eval<PATH>(d, x, xr, a.weights[i], a.dHat2, a.kappa, a.xi2, a.projectSPD != 0, WANT_H, V9, V6, out, em);
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
#define IDP_BARRIER_CASE(E, G, H)                                                                                         \
    do {                                                                                                                  \
        if (smem) IDP_CK(c, cudaFuncSetAttribute(k_barrier<PATH, E, G, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        IDP_LAUNCH(c, (k_barrier<PATH, E, G, H>), grid, T, smem, a);                                                  \
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
    if (want_h) { c->nnz = 0; c->nBlocksUnique = 0; }
    if (c->nRows == 0) return IDP_OK;
    StageTimer tm(c, IDP_STAGE_BARRIER);
    IDP_CK(c, cudaMemsetAsync(c->counters.p + CNT_ERR_DIST, 0, sizeof(long long), c->stream));
    IDP_CK(c, cudaMemsetAsync(c->counters.p + CNT_ERR_EIG, 0, sizeof(long long), c->stream));
    // this rank's rows (row_owner) in an order grouped by kind
    const int ownRanks = c->rowsLocal ? 1 : c->nranks; // LOCAL-ROWS mode: every row in c->rows is this rank's
    int ownerShift = 8;
    while (ownerShift < 14 && (c->nV >> ownerShift) > 16 * c->nranks) ++ownerShift;
    if (!c->permValid) {
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
        c->permValid = true;
    }
    const long nPath[3] = {c->kindCount[K_EE] + c->kindCount[K_EE_M] + c->kindCount[K_PE_M] + c->kindCount[K_PP_M] + c->kindCount[K_PT],
        c->kindCount[K_PE], c->kindCount[K_PP]};
    const long nMine = nPath[0] + nPath[1] + nPath[2];
    if (!c->have_x0 && c->kindCount[K_EE_M] + c->kindCount[K_PE_M] + c->kindCount[K_PP_M] > 0)
        return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "mollified rows need the rest positions (idp_set_rest_positions)", __FILE__, __LINE__);
    const unsigned grid = std::max(1u, std::min(blocks_for(std::max(nPath[0], std::max(nPath[1], nPath[2])), 128), (unsigned)c->sm_count * 16));
    BarrierArgs a;
    a.rows = c->rows.p; a.weights = c->weights.p; a.perm = c->rowPerm.p; a.jBegin = 0; a.jEnd = 0;
    a.xp = c->xp.p; a.x0p = c->x0p.p;
    a.dHat2 = dhat2 + 2 * std::sqrt(dhat2) * thickness; // IPC.h:757
    a.kappa = kappa; a.xi2 = thickness * thickness; a.projectSPD = project_spd;
    IDP_CK(c, c->red.reserve(3 * (size_t)grid + 8));
    a.partialE = c->red.p;
    a.g = c->gbuf.p;
    a.errDist = (unsigned long long*)(c->counters.p + CNT_ERR_DIST);
    a.errEig = (unsigned long long*)(c->counters.p + CNT_ERR_EIG);
    a.nVll = c->nV;
    a.blkOff = nullptr; a.blkKey = nullptr; a.blkIdx = nullptr; a.blkVal = nullptr;
    long nBlocks = 0;
    if (want_h) {
        IDP_CK(c, c->rowBlkOff.reserve(c->nRows + 1));
        IDP_CK(c, c->segId.reserve(c->nRows + 1));
        IDP_LAUNCH(c, k_row_block_counts, blocks_for(c->nRows + 1, 256), 256, 0, c->rows.p, c->nRows, c->rank, ownRanks, ownerShift, c->segId.p);
        IDP_TRY(cub_scan_exclusive(c, c->segId.p, c->rowBlkOff.p, c->nRows + 1));
        int ends[2] = {0, 0};
        IDP_CK(c, cudaMemcpyAsync(&ends[1], c->rowBlkOff.p + c->nRows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        nBlocks = ends[1]; // offsets count only this rank's rows, so its blocks are contiguous in [0, nBlocks)
        IDP_CK(c, c->blkKey.reserve(std::max<long>(nBlocks, 1)));
        IDP_CK(c, c->blkIdx.reserve(std::max<long>(nBlocks, 1)));
        IDP_CK(c, c->blkVal.reserve(9 * (size_t)std::max<long>(nBlocks, 1)));
        a.blkOff = c->rowBlkOff.p; a.blkKey = c->blkKey.p; a.blkIdx = c->blkIdx.p; a.blkVal = c->blkVal.p;
        c->nBlocksUnique = 0;
        c->nnz = 0;
        // remember the slice for the assembly
        c->h_counters[CNT_COUNT - 2] = ends[0];
        c->h_counters[CNT_COUNT - 1] = ends[1];
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
// assembly: sort (vi*nV + vj) keys, segmented sum of the 3x3 blocks, scalar CSR with ascending columns
// ------------------------------------------------------------------------------------------------------------
__global__ void k_head_flags(const unsigned long long* __restrict__ keys, long n, int* __restrict__ flag)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// segId = inclusive scan of flags - 1; scatter segment starts
__global__ void k_seg_starts(const int* __restrict__ flagScan, const int* __restrict__ flag, long n, int* __restrict__ segStart)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        if (flag[i]) segStart[flagScan[i]] = (int)i; // flagScan = exclusive scan -> segment index
}
// first unique (upper) block of every block row: vtxStart[v] = lower_bound(uniqueKey, v * nV)
__global__ void k_vertex_block_starts(const unsigned long long* __restrict__ keys, const int* __restrict__ segStart, int nSeg,
    int nV, int* __restrict__ vtxStart)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= nV; v += gridDim.x * blockDim.x) {
        const unsigned long long target = (unsigned long long)v * (unsigned long long)nV;
        int lo = 0, hi = nSeg;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[segStart[mid]] < target) lo = mid + 1;
            else hi = mid;
        }
        vtxStart[v] = lo;
    }
}
// one thread per (unique upper block, component): sum the members in sorted (= emission) order into ublk[seg][9];
// component 0 also records the block's column vertex and counts the strictly-upper blocks per column (= lower blocks per row)
__global__ void __launch_bounds__(288) k_reduce_blocks(const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
    const int* __restrict__ segStart, int nSeg, long nTot, const double* __restrict__ blkVal, long long nV,
    double* __restrict__ ublk, int* __restrict__ ucol, int* __restrict__ urow, int* __restrict__ lowerCount)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long seg = t / 9;
    const int comp = (int)(t - seg * 9);
    if (seg >= nSeg) return;
    const int s0 = segStart[seg];
    const int s1 = (seg + 1 < nSeg) ? segStart[seg + 1] : (int)nTot;
    double s = 0;
    for (int m = s0; m < s1; ++m) s += blkVal[9 * (long)idx[m] + comp];
    ublk[9 * seg + comp] = s;
    if (comp == 0) {
        const unsigned long long key = keys[s0];
        const int vi = (int)(key / (unsigned long long)nV), vj = (int)(key - (unsigned long long)vi * (unsigned long long)nV);
        urow[seg] = vi;
        ucol[seg] = vj;
        if (vi != vj) atomicAdd(&lowerCount[vj], 1);
    }
}
// block-row layout: [lower blocks (mirrors, ascending column)] [upper blocks incl. diagonal (ascending column)]
// cnt[v] = lowerCount[v] + upperCount[v]
__global__ void k_row_totals(const int* __restrict__ vtxStart, const int* __restrict__ lowerCount, int nV, int* __restrict__ cnt)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= nV; v += gridDim.x * blockDim.x)
        cnt[v] = (v < nV) ? lowerCount[v] + (vtxStart[v + 1] - vtxStart[v]) : 0;
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
// upper (and diagonal) blocks: one thread per (unique block, component)
__global__ void __launch_bounds__(288) k_write_upper(const double* __restrict__ ublk, const int* __restrict__ urow, const int* __restrict__ ucol,
    int nSeg, const int* __restrict__ vtxStart, const int* __restrict__ rowStart, const int* __restrict__ lowerCount,
    int* __restrict__ col, double* __restrict__ val)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long seg = t / 9;
    const int comp = (int)(t - seg * 9);
    if (seg >= nSeg) return;
    const int vi = urow[seg], vj = ucol[seg];
    const int bs = rowStart[vi], nb = rowStart[vi + 1] - bs;
    const int slot = lowerCount[vi] + (int)(seg - vtxStart[vi]);
    const int a = comp / 3, b = comp - 3 * a;
    const long pos = 9L * bs + (long)a * 3 * nb + 3L * slot + b;
    col[pos] = 3 * vj + b;
    val[pos] = ublk[9 * seg + comp];
}
// mirrored blocks: `order` lists the strictly-upper unique blocks stably sorted by their column vertex, so the members
// of one block row appear with ascending row vertex = ascending column in the mirror. lstart[v] = first entry of row v.
__global__ void __launch_bounds__(288) k_write_lower(const double* __restrict__ ublk, const int* __restrict__ urow, const int* __restrict__ order,
    const int* __restrict__ sortedCol, long nLower, const int* __restrict__ lstart, const int* __restrict__ rowStart,
    int* __restrict__ col, double* __restrict__ val)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long e = t / 9;
    const int comp = (int)(t - e * 9);
    if (e >= nLower) return;
    const int seg = order[e];
    const int row = sortedCol[e], cv = urow[seg]; // mirror: row vertex = original column, column vertex = original row
    const int bs = rowStart[row], nb = rowStart[row + 1] - bs;
    const int slot = (int)(e - lstart[row]);
    const int a = comp / 3, b = comp - 3 * a;
    const long pos = 9L * bs + (long)a * 3 * nb + 3L * slot + b;
    col[pos] = 3 * cv + b;
    val[pos] = ublk[9 * seg + 3 * b + a]; // transposed block
}
__global__ void k_lower_keys(const int* __restrict__ urow, const int* __restrict__ ucol, int nSeg, int* __restrict__ keyOut, int* __restrict__ segOut,
    unsigned long long* __restrict__ counter)
{
    // compacts the strictly-upper blocks in segment order (deterministic: slot = exclusive count of earlier ones is not
    // needed because the following sort is stable on (column) and the input order here is by segment via the scan below)
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
    StageTimer tm(c, IDP_STAGE_CSR);
    IDP_CK(c, c->csrPtr.reserve(3 * (size_t)c->nV + 1));
    const long b0 = c->nRows ? (long)c->h_counters[CNT_COUNT - 2] : 0, b1 = c->nRows ? (long)c->h_counters[CNT_COUNT - 1] : 0;
    const long n = b1 - b0;
    if (n <= 0) {
        IDP_CK(c, cudaMemsetAsync(c->csrPtr.p, 0, (3 * (size_t)c->nV + 1) * sizeof(int), c->stream));
        c->nnz = 0; c->nBlocksUnique = 0;
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        return IDP_OK;
    }
    const int nV = c->nV;
    IDP_CK(c, c->blkKeySorted.reserve(n));
    IDP_CK(c, c->blkIdxSorted.reserve(n));
    int bits = 1;
    while (bits < 64 && ((unsigned long long)nV * (unsigned long long)nV) >> bits) ++bits;
    size_t bytes = 0;
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->blkKey.p + b0, c->blkKeySorted.p, c->blkIdx.p + b0, c->blkIdxSorted.p, (int)n, 0, bits, c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, c->blkKey.p + b0, c->blkKeySorted.p, c->blkIdx.p + b0, c->blkIdxSorted.p, (int)n, 0, bits, c->stream));
    ++c->lib_launches;
    // unique upper blocks
    IDP_CK(c, c->segId.reserve(std::max<long>(n + 1, nV + 2)));
    IDP_CK(c, c->segStart.reserve(n + 1));
    IDP_CK(c, c->rowBlkOff.reserve(std::max<long>(n + 1, nV + 2))); // reused as scan output (row offsets are no longer needed)
    IDP_LAUNCH(c, k_head_flags, blocks_for(n, 256), 256, 0, c->blkKeySorted.p, n, c->segId.p);
    IDP_CK(c, cudaMemsetAsync(c->segId.p + n, 0, sizeof(int), c->stream));
    IDP_TRY(cub_scan_exclusive(c, c->segId.p, c->rowBlkOff.p, n + 1));
    int nSeg = 0;
    IDP_CK(c, cudaMemcpyAsync(&nSeg, c->rowBlkOff.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IDP_LAUNCH(c, k_seg_starts, blocks_for(n, 256), 256, 0, c->rowBlkOff.p, c->segId.p, n, c->segStart.p);
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    // reduce duplicates -> dense unique upper blocks
    IDP_CK(c, c->ublk.reserve(9 * (size_t)nSeg));
    IDP_CK(c, c->urow.reserve(nSeg)); IDP_CK(c, c->ucol.reserve(nSeg));
    IDP_CK(c, c->lowerCount.reserve((size_t)nV + 2)); IDP_CK(c, c->lstart.reserve((size_t)nV + 2));
    IDP_CK(c, c->vtxBlkStart.reserve((size_t)nV + 2)); IDP_CK(c, c->rowStart.reserve((size_t)nV + 2));
    IDP_CK(c, cudaMemsetAsync(c->lowerCount.p, 0, ((size_t)nV + 2) * sizeof(int), c->stream));
    IDP_LAUNCH(c, k_reduce_blocks, blocks_for(9L * nSeg, 288), 288, 0, c->blkKeySorted.p, c->blkIdxSorted.p, c->segStart.p, nSeg, n,
        c->blkVal.p, (long long)nV, c->ublk.p, c->ucol.p, c->urow.p, c->lowerCount.p);
    IDP_LAUNCH(c, k_vertex_block_starts, blocks_for(nV + 1, 256), 256, 0, c->blkKeySorted.p, c->segStart.p, nSeg, nV, c->vtxBlkStart.p);
    // block-row sizes and starts
    IDP_LAUNCH(c, k_row_totals, blocks_for(nV + 1, 256), 256, 0, c->vtxBlkStart.p, c->lowerCount.p, nV, c->segId.p);
    IDP_TRY(cub_scan_exclusive(c, c->segId.p, c->rowStart.p, (long)nV + 1));
    int nBlkTotal = 0;
    IDP_CK(c, cudaMemcpyAsync(&nBlkTotal, c->rowStart.p + nV, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    // mirrored (lower) blocks: stable sort of the unique blocks by column vertex
    IDP_CK(c, c->lkey.reserve(nSeg)); IDP_CK(c, c->lkeySorted.reserve(nSeg)); IDP_CK(c, c->lseg.reserve(nSeg)); IDP_CK(c, c->lsegSorted.reserve(nSeg));
    IDP_LAUNCH(c, k_lower_keys, blocks_for(nSeg, 256), 256, 0, c->urow.p, c->ucol.p, nSeg, c->lkey.p, c->lseg.p, (unsigned long long*)nullptr);
    int vbits = 1;
    while (vbits < 31 && (nV >> vbits)) ++vbits;
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->lkey.p, c->lkeySorted.p, c->lseg.p, c->lsegSorted.p, nSeg, 0, 31, c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, c->lkey.p, c->lkeySorted.p, c->lseg.p, c->lsegSorted.p, nSeg, 0, 31, c->stream));
    ++c->lib_launches;
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    const long nLower = (long)nBlkTotal - nSeg; // every strictly-upper block has one mirror
    c->nBlocksUnique = nBlkTotal;
    c->nnz = 9L * nBlkTotal;
    IDP_CK(c, c->csrCol.reserve(std::max<long>(c->nnz, 1)));
    IDP_CK(c, c->csrVal.reserve(std::max<long>(c->nnz, 1)));
    IDP_LAUNCH(c, k_lower_starts, blocks_for(nV + 1, 256), 256, 0, c->lkeySorted.p, nLower, nV, c->lstart.p);
    IDP_LAUNCH(c, k_csr_ptr, blocks_for(nV + 1, 256), 256, 0, c->rowStart.p, nV, c->csrPtr.p);
    IDP_LAUNCH(c, k_write_upper, blocks_for(9L * nSeg, 288), 288, 0, c->ublk.p, c->urow.p, c->ucol.p, nSeg, c->vtxBlkStart.p, c->rowStart.p,
        c->lowerCount.p, c->csrCol.p, c->csrVal.p);
    if (nLower > 0)
        IDP_LAUNCH(c, k_write_lower, blocks_for(9L * nLower, 288), 288, 0, c->ublk.p, c->urow.p, c->lsegSorted.p, c->lkeySorted.p, nLower, c->lstart.p,
            c->rowStart.p, c->csrCol.p, c->csrVal.p);
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
