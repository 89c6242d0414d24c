// Broad phase (static + swept), narrow-phase classification, duplicate merge, min-distance and additive CCD.
// COMPILED WITH --fmad=false: every floating-point result here feeds a comparison whose outcome must match the
// CPU reference bit for bit (voxel indices, AABB tests, distance types, d < dHat^2, ACCD step).
//
// Reference operators replaced (relative to /root/reference/Library):
//   Grid/SPATIAL_HASH.h:28-291 (static build + queries), 432-662 (CCD build + queries),
//   FEM/IPC.h:145-661 (PT / EE loops + merge), 1957-2243 (CCD loops), 2246-2388 (min distance),
//   Math/Distance/CCD.h:149-235 (AABB predicates), 279-395 (ACCD).
//
// Data layout in HBM: positions are kept twice — SoA x[],y[],z[] for streaming reductions and
// 32-byte packed double4 records for the per-pair gathers (one DRAM sector per vertex); every boundary primitive
// gets a 64-byte PrimRec (exact AABB + vertex ids) and a 32-byte integer lattice box.
#include "ctx.cuh"
#include "pair_exact.cuh"
#include <cub/cub.cuh>

namespace idp {

// ------------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 ldv(const double4* __restrict__ p, int v)
{
    const double2* q = reinterpret_cast<const double2*>(p + v);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return mk3(a.x, a.y, b.x);
}
// order-preserving map double -> uint64 (for atomicMin / atomicMax on doubles of either sign)
__device__ __forceinline__ unsigned long long enc_ord(double d)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_ord(unsigned long long u)
{
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// warp-aggregated append: returns the slot of this lane's item (valid only where pred), -1 if over capacity
__device__ __forceinline__ long warp_append(bool pred, unsigned long long* counter, long cap)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return -1;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!pred) return -1;
    const long slot = (long)base + __popc(m & ((1u << lane_id()) - 1u));
    return slot < cap ? slot : -1;
}

__device__ __forceinline__ int lat_index(double c, double lo, double inv) { return (int)floor((c - lo) * inv); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------------------------------------------------
// upload marshalling: host AoS (stride doubles per vertex) -> SoA + packed
// ------------------------------------------------------------------------------------------------------------
__global__ void k_scatter_positions(const double* __restrict__ aos, int stride, int nV, double* __restrict__ xs,
    double* __restrict__ ys, double* __restrict__ zs, double4* __restrict__ packed)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += gridDim.x * blockDim.x) {
        const double x = aos[(long)v * stride], y = aos[(long)v * stride + 1], z = aos[(long)v * stride + 2];
        if (xs) { xs[v] = x; ys[v] = y; zs[v] = z; }
        packed[v] = make_double4(x, y, z, 0.0);
    }
}

int upload_positions(idp_ctx* c, const double* host, int stride, int which)
{
    if (c->nV <= 0) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_set_mesh must be called first", __FILE__, __LINE__);
    if (stride < 3) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "stride must be >= 3", __FILE__, __LINE__);
    StageTimer tm(c, IDP_STAGE_UPLOAD);
    const size_t n = (size_t)c->nV * stride;
    IDP_CK(c, c->stage.reserve(n));
    IDP_CK(c, cudaMemcpyAsync(c->stage.p, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DBuf<double4>& dst = which == 0 ? c->xp : (which == 1 ? c->x0p : (which == 2 ? c->dp : c->xnp));
    IDP_CK(c, dst.reserve(c->nV));
    if (which == 0) {
        IDP_CK(c, c->xs.reserve(c->nV));
        IDP_CK(c, c->ys.reserve(c->nV));
        IDP_CK(c, c->zs.reserve(c->nV));
    }
    IDP_LAUNCH(c, k_scatter_positions, blocks_for(c->nV, 256), 256, 0, c->stage.p, stride, c->nV,
        which == 0 ? c->xs.p : nullptr, c->ys.p, c->zs.p, dst.p);
    IDP_CK(c, cudaGetLastError());
    if (which == 0) { c->have_x = true; ++c->xVersion; }
    if (which == 1) c->have_x0 = true;
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// reductions: bounding box and the balanced adjacent-pair tree sum (mean edge length / mean |dir|)
// ------------------------------------------------------------------------------------------------------------
// out[0..2] = min (encoded), out[3..5] = max (encoded). mode 0: positions; mode 1: swept min/max(x, x + alpha d)
__global__ void k_node_bbox(const int* __restrict__ bnode, int nBN, const double4* __restrict__ xp,
    const double4* __restrict__ dp, double alpha, int mode, unsigned long long* __restrict__ out)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nBN; s += gridDim.x * blockDim.x) {
        const int v = bnode[s];
        const V3 x = ldv(xp, v);
        V3 a = x, b = x;
        if (mode == 1) {
            const V3 d = ldv(dp, v);
            const V3 xt = mk3(x.x + alpha * d.x, x.y + alpha * d.y, x.z + alpha * d.z);
            a = min3(x, xt);
            b = max3(x, xt);
        }
        lo[0] = fmin(lo[0], a.x); lo[1] = fmin(lo[1], a.y); lo[2] = fmin(lo[2], a.z);
        hi[0] = fmax(hi[0], b.x); hi[1] = fmax(hi[1], b.y); hi[2] = fmax(hi[2], b.z);
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    for (int k = 0; k < 3; ++k) {
        const double mn = BR(tmp).Reduce(lo[k], cub::Min());
        __syncthreads();
        const double mx = BR(tmp).Reduce(hi[k], cub::Max());
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicMin(out + k, enc_ord(mn));
            atomicMax(out + 3 + k, enc_ord(mx));
        }
    }
}

__global__ void k_edge_lengths(const int2* __restrict__ bedge, int nBE, const double4* __restrict__ xp, double* __restrict__ len)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nBE; e += gridDim.x * blockDim.x) {
        const int2 ab = bedge[e];
        len[e] = sqrt(sqn3(ldv(xp, ab.x) - ldv(xp, ab.y)));
    }
}
__global__ void k_abs_dir(const int* __restrict__ bnode, int nBN, const double4* __restrict__ dp, double* __restrict__ out)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nBN; s += gridDim.x * blockDim.x) {
        const V3 d = ldv(dp, bnode[s]);
        out[3 * (long)s] = fabs(d.x); out[3 * (long)s + 1] = fabs(d.y); out[3 * (long)s + 2] = fabs(d.z);
    }
}
// One block reduces an aligned chunk of 2048 inputs with the adjacent-pair binary tree (zero padded), which makes the
// multi-pass result identical to the tree over the whole array padded to a power of two (oracle: orc::tree_sum).
__global__ void __launch_bounds__(1024) k_tree_reduce(const double* __restrict__ in, long n, double* __restrict__ out)
{
    __shared__ double s[1024];
    const long base = (long)blockIdx.x * 2048;
    const int t = threadIdx.x;
    const long i0 = base + 2 * t, i1 = i0 + 1;
    const double a = i0 < n ? in[i0] : 0.0, b = i1 < n ? in[i1] : 0.0;
    s[t] = a + b;
    __syncthreads();
    for (int len = 1024; len > 1; len >>= 1) {
        double v = 0;
        const bool act = t < (len >> 1);
        if (act) v = s[2 * t] + s[2 * t + 1];
        __syncthreads();
        if (act) s[t] = v;
        __syncthreads();
    }
    if (t == 0) out[blockIdx.x] = s[0];
}
// tree sum of n doubles in buf (clobbers scratch); result to host
static int tree_sum_device(idp_ctx* c, double* buf, long n, double* scratchA, double* scratchB, double* host_out)
{
    if (n <= 0) { *host_out = 0.0; return IDP_OK; }
    const double* in = buf;
    double* outs[2] = {scratchA, scratchB};
    int flip = 0;
    long m = n;
    while (true) {
        const long nb = (m + 2047) / 2048;
        IDP_LAUNCH(c, k_tree_reduce, (unsigned)nb, 1024, 0, in, m, outs[flip]);
        in = outs[flip];
        flip ^= 1;
        m = nb;
        if (nb == 1) break;
    }
    IDP_CK(c, cudaMemcpyAsync(host_out, in, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// primitive records and lattice boxes
// ------------------------------------------------------------------------------------------------------------
struct PrepArgs {
    const int* bnode; const int2* bedge; const int4* btri; const unsigned char* dbc;
    const double4* xp; const double4* dp; // dp only for CCD
    GridDesc g;
    double radius;      // static: inflated query radius r'; CCD: unused
    const IBox* vbox;   // CCD: per-vertex lattice boxes
    // Sharded contexts prepare only what their queries can meet: primitives [ownB, ownE) are this rank's queries (always
    // prepared, phase 0); the others (phase 1) are potential partners and are skipped -- insert box marked with the
    // sentinel pad[0] = -1, no 64-byte record written -- unless their lattice box overlaps `region`, the lattice bounding
    // box of this rank's query boxes (a candidate's insert box always overlaps the inflated query box).
    int ownB = 0, ownE = 0, phase = 0;
    const int* region = nullptr; // {lo[3], hi[3]} on the device; nullptr: no filter (single GPU)
};
__device__ __forceinline__ bool outside_region(const int* __restrict__ rg, const IBox& b)
{
    return b.hi[0] < rg[0] || b.hi[1] < rg[1] || b.hi[2] < rg[2] || b.lo[0] > rg[3] || b.lo[1] > rg[4] || b.lo[2] > rg[5];
}
// lattice bounding box of the query boxes [qb, qe)
__global__ void k_lat_region(const IBox* __restrict__ qbox, int qb, int qe, int* __restrict__ region)
{
    int lo[3] = {2147483647, 2147483647, 2147483647}, hi[3] = {-2147483647, -2147483647, -2147483647};
    for (int q = qb + blockIdx.x * blockDim.x + threadIdx.x; q < qe; q += gridDim.x * blockDim.x) {
        const IBox b = qbox[q];
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], b.lo[k]); hi[k] = max(hi[k], b.hi[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&region[k], lo[k]); atomicMax(&region[3 + k], hi[k]); }
    }
}
__device__ __forceinline__ void box_from_aabb(const GridDesc& g, const V3& lo, const V3& hi, double r, IBox& b)
{
    b.lo[0] = lat_index(lo.x - r, g.lo[0], g.inv); b.lo[1] = lat_index(lo.y - r, g.lo[1], g.inv); b.lo[2] = lat_index(lo.z - r, g.lo[2], g.inv);
    b.hi[0] = lat_index(hi.x + r, g.lo[0], g.inv); b.hi[1] = lat_index(hi.y + r, g.lo[1], g.inv); b.hi[2] = lat_index(hi.z + r, g.lo[2], g.inv);
    if (g.clampLat) {
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = clampi(b.lo[k], 0, g.n[k] * g.k - 1);
            b.hi[k] = clampi(b.hi[k], 0, g.n[k] * g.k - 1);
        }
    }
    b.pad[0] = b.pad[1] = 0;
}
__device__ __forceinline__ void rec_set(PrimRec& r, const V3& lo, const V3& hi, int v0, int v1, int v2, int flags)
{
    r.lo[0] = lo.x; r.lo[1] = lo.y; r.lo[2] = lo.z; r.hi[0] = hi.x; r.hi[1] = hi.y; r.hi[2] = hi.z;
    r.v[0] = v0; r.v[1] = v1; r.v[2] = v2; r.flags = flags;
}
__device__ __forceinline__ void swept(const PrepArgs& a, int v, V3& lo, V3& hi)
{
    const V3 x = ldv(a.xp, v);
    if (a.dp) {
        const V3 xe = x + ldv(a.dp, v);
        lo = min3(x, xe); hi = max3(x, xe);
    }
    else { lo = x; hi = x; }
}
__device__ __forceinline__ void box_union(IBox& o, const IBox& a, const IBox& b)
{
    for (int k = 0; k < 3; ++k) { o.lo[k] = min(a.lo[k], b.lo[k]); o.hi[k] = max(a.hi[k], b.hi[k]); }
    o.pad[0] = o.pad[1] = 0;
}

// nodes: query records. static: box = idx(p -+ r'); CCD: box = vbox[v]
__global__ void k_prep_nodes(PrepArgs a, int nBN, PrimRec* __restrict__ rec, IBox* __restrict__ qbox)
{
    (void)nBN; // nodes are only ever queries: this rank's range is all that is needed
    for (int s = a.ownB + blockIdx.x * blockDim.x + threadIdx.x; s < a.ownE; s += gridDim.x * blockDim.x) {
        const int v = a.bnode[s];
        V3 lo, hi;
        swept(a, v, lo, hi);
        PrimRec r;
        rec_set(r, lo, hi, v, -1, -1, a.dbc[v] ? 1 : 0);
        rec[s] = r;
        IBox b;
        if (a.vbox) b = a.vbox[v];
        else box_from_aabb(a.g, lo, hi, a.radius, b);
        qbox[s] = b;
    }
}
__global__ void k_prep_edges(PrepArgs a, int nBE, PrimRec* __restrict__ rec, IBox* __restrict__ qbox, IBox* __restrict__ bbox)
{
    // phase 0: [ownB, ownE); phase 1: the rest, filtered by the region
    const int n = a.phase == 0 ? a.ownE - a.ownB : nBE - (a.ownE - a.ownB);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int e = a.phase == 0 ? a.ownB + i : (i < a.ownB ? i : i + (a.ownE - a.ownB));
        const int2 ab = a.bedge[e];
        V3 l0, h0, l1, h1;
        swept(a, ab.x, l0, h0);
        swept(a, ab.y, l1, h1);
        const V3 lo = min3(l0, l1), hi = max3(h0, h1);
        IBox b;
        if (a.vbox) box_union(b, a.vbox[ab.x], a.vbox[ab.y]);
        else box_from_aabb(a.g, lo, hi, 0.0, b);
        if (a.phase == 1 && a.region && outside_region(a.region, b)) {
            b.pad[0] = -1;
            bbox[e] = b;
            continue;
        }
        bbox[e] = b; // query and insert boxes coincide for CCD
        PrimRec r;
        rec_set(r, lo, hi, ab.x, ab.y, -1, (a.dbc[ab.x] && a.dbc[ab.y]) ? 1 : 0);
        rec[e] = r;
        if (!a.vbox && a.phase == 0) {
            box_from_aabb(a.g, lo, hi, a.radius, b);
            qbox[e] = b;
        }
    }
}
__global__ void k_prep_tris(PrepArgs a, int nBT, PrimRec* __restrict__ rec, IBox* __restrict__ bbox)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nBT; t += gridDim.x * blockDim.x) {
        const int4 f = a.btri[t];
        V3 l0, h0, l1, h1, l2, h2;
        swept(a, f.x, l0, h0);
        swept(a, f.y, l1, h1);
        swept(a, f.z, l2, h2);
        const V3 lo = min3(min3(l0, l1), l2), hi = max3(max3(h0, h1), h2);
        IBox b;
        if (a.vbox) {
            IBox t01;
            box_union(t01, a.vbox[f.x], a.vbox[f.y]);
            box_union(b, t01, a.vbox[f.z]);
        }
        else box_from_aabb(a.g, lo, hi, 0.0, b);
        if (a.region && outside_region(a.region, b)) { // triangles are only ever partners
            b.pad[0] = -1;
            bbox[t] = b;
            continue;
        }
        bbox[t] = b;
        PrimRec r;
        rec_set(r, lo, hi, f.x, f.y, f.z, (a.dbc[f.x] && a.dbc[f.y] && a.dbc[f.z]) ? 1 : 0);
        rec[t] = r;
    }
}
// CCD: per-vertex lattice box of the alpha-scaled, xi/2-inflated swept segment (SPATIAL_HASH.h:525-533)
__global__ void k_ccd_vertex_boxes(const int* __restrict__ bnode, int nBN, const double4* __restrict__ xp,
    const double4* __restrict__ dp, double alpha, double half, GridDesc g, IBox* __restrict__ vbox)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nBN; s += gridDim.x * blockDim.x) {
        const int v = bnode[s];
        const V3 x = ldv(xp, v), d = ldv(dp, v);
        const V3 xt = mk3(x.x + alpha * d.x, x.y + alpha * d.y, x.z + alpha * d.z);
        const V3 mn = min3(x, xt), mx = max3(x, xt);
        IBox b;
        b.lo[0] = lat_index(mn.x - half, g.lo[0], g.inv); b.lo[1] = lat_index(mn.y - half, g.lo[1], g.inv); b.lo[2] = lat_index(mn.z - half, g.lo[2], g.inv);
        b.hi[0] = lat_index(mx.x + half, g.lo[0], g.inv); b.hi[1] = lat_index(mx.y + half, g.lo[1], g.inv); b.hi[2] = lat_index(mx.z + half, g.lo[2], g.inv);
        b.pad[0] = b.pad[1] = 0;
        vbox[v] = b;
    }
}

// ------------------------------------------------------------------------------------------------------------
// cell lists: histogram -> exclusive scan -> fill
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cell_range(const GridDesc& g, const IBox& b, int lo[3], int hi[3])
{
    if (g.k == 1) { // static phase and most CCD grids: no division
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = clampi(b.lo[k], 0, g.n[k] - 1);
            hi[k] = clampi(b.hi[k], 0, g.n[k] - 1);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = clampi(b.lo[k] / g.k, 0, g.n[k] - 1);
        hi[k] = clampi(b.hi[k] / g.k, 0, g.n[k] - 1);
    }
}
// note: lattice indices are >= 0 on the paths that use k > 1 (CCD lattice origin is the global minimum), so integer
// division is a monotone floor.
//
// Cell lists with ANCHOR insertion in LEVELS: a primitive is stored once, in the cell of the low corner of its box, in the
// first level that covers its extent. Bins 0 and 1 use the base cells and hold the primitives that span at most 1 / exactly
// 2 cells beyond their corner; bin j >= 2 has cells of 2^j base cells and holds what spans at most two of those per axis.
// A query visits, per non-empty level, the cells [qlo - ext, qhi] (ext = the largest span of the level per axis): every
// pair is met exactly once -- no de-duplication, no per-cell copies -- and the neighbourhood of a query does not grow with
// the largest primitive in the scene (the swept CCD boxes have a long-tailed extent distribution: a single level either
// visits 6^3 cells for everybody or parks the tail on a list that every query scans). The tail levels are sparse and have
// few, large cells, so they cost a query a handful of rows. A thinly populated bin 1 is folded into bin 2 (useBin1 = 0).
#define IDP_MAX_LEVELS 12   // active (non-empty) levels
#define IDP_LEVEL_BINS 28
struct Levels {
    int n;                          // active levels
    int useBin1;
    int shift[IDP_MAX_LEVELS];      // cell = base cell >> shift
    int ext[IDP_MAX_LEVELS][3];     // largest span (in level cells) of the level's primitives per axis
    int dim[IDP_MAX_LEVELS][3];     // cells per axis
    long off[IDP_MAX_LEVELS];       // first cell of the level in the concatenated cellStart array
    signed char slot[IDP_LEVEL_BINS]; // level bin -> active index (-1: empty)
};
__host__ __device__ __forceinline__ int level_shift(int bin) { return bin < 2 ? 0 : bin; }
__device__ __forceinline__ int level_of(const int lo[3], const int hi[3], int useBin1)
{
    const int m = max(hi[0] - lo[0], max(hi[1] - lo[1], hi[2] - lo[2]));
    if (m <= 1) return 0;
    if (m == 2 && useBin1) return 1;
    int s = 2;
    while (s < IDP_LEVEL_BINS - 1 && (((hi[0] >> s) - (lo[0] >> s)) > 1 || ((hi[1] >> s) - (lo[1] >> s)) > 1 || ((hi[2] >> s) - (lo[2] >> s)) > 1)) ++s;
    return s;
}
// hist[bin] = primitives of the level; hist[IDP_LEVEL_BINS + 3 bin + k] = their largest span on axis k
__global__ void k_level_hist(const IBox* __restrict__ bbox, int nB, GridDesc g, int* __restrict__ hist)
{
    __shared__ int sh[4 * IDP_LEVEL_BINS];
    for (int i = threadIdx.x; i < 4 * IDP_LEVEL_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nB; b += gridDim.x * blockDim.x) {
        int lo[3], hi[3];
        const IBox bx = bbox[b];
        if (bx.pad[0] < 0) continue; // not prepared on this shard (outside the query region)
        cell_range(g, bx, lo, hi);
        const int l = level_of(lo, hi, 1), s = level_shift(l);
        atomicAdd(&sh[l], 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = (hi[k] >> s) - (lo[k] >> s);
            if (e > sh[IDP_LEVEL_BINS + 3 * l + k]) atomicMax(&sh[IDP_LEVEL_BINS + 3 * l + k], e);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * IDP_LEVEL_BINS; i += blockDim.x)
        if (sh[i]) { if (i < IDP_LEVEL_BINS) atomicAdd(&hist[i], sh[i]); else atomicMax(&hist[i], sh[i]); }
}
// Shards only look at the cells their own queries visit: the bounding box (in cell coordinates) of the query ranges of
// this rank is reduced on the device and the fill kernels skip everything outside it, so the cost of building the cell
// lists shrinks with the shard instead of being replicated on every GPU. region = {lo[3], hi[3]}; one rank: whole grid.
__global__ void k_query_region(const IBox* __restrict__ qbox, int qBegin, int qEnd, GridDesc g, int* __restrict__ region)
{
    int lo[3] = {2147483647, 2147483647, 2147483647}, hi[3] = {-1, -1, -1};
    for (int q = qBegin + blockIdx.x * blockDim.x + threadIdx.x; q < qEnd; q += gridDim.x * blockDim.x) {
        int l[3], h[3];
        cell_range(g, qbox[q], l, h);
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], l[k]); hi[k] = max(hi[k], h[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
        hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&region[k], lo[k]); atomicMax(&region[3 + k], hi[k]); }
    }
}
// Cell-sorted filter records, SoA: rec0[i] = (lo.x, lo.y, lo.z, hi.x), rec1[i] = (hi.y, hi.z, primitive, 0): the primitive's
// AABB (static phase) or full-step swept AABB (CCD) rounded OUTWARD to float -- a conservative pre-filter; the exact
// predicate (FP64 AABB gap; for CCD also the integer "shares a voxel" test of the reference's hash) is applied to the
// survivors. Consecutive lanes of the query kernel read consecutive records, so both 16-byte streams are fully coalesced.
template <bool FILL>
__global__ void k_cells(const IBox* __restrict__ bbox, const PrimRec* __restrict__ rec, int nB, GridDesc g, Levels lv,
    const int* __restrict__ region, int* __restrict__ cellCountOrCursor, int4* __restrict__ rec0, int4* __restrict__ rec1)
{
    int rg[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) rg[k] = region[k];
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nB; b += gridDim.x * blockDim.x) {
        int lo[3], hi[3];
        const IBox bx = bbox[b];
        if (bx.pad[0] < 0) continue; // not prepared on this shard (outside the query region)
        cell_range(g, bx, lo, hi);
        const int li = lv.slot[level_of(lo, hi, lv.useBin1)], l = lv.shift[li];
        // queries visit the anchors in [qlo - ext, qhi] (level cells): anchors outside the shard's region are never met
        bool skip = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) skip = skip || (lo[k] >> l) < (rg[k] >> l) - lv.ext[li][k] || (lo[k] >> l) > (rg[3 + k] >> l);
        if (skip) continue;
        const long cell = lv.off[li] + ((long)(lo[2] >> l) * lv.dim[li][1] + (lo[1] >> l)) * lv.dim[li][0] + (lo[0] >> l);
        if (FILL) {
            const PrimRec r = rec[b];
            const int4 r0 = make_int4(__float_as_int(__double2float_rd(r.lo[0])), __float_as_int(__double2float_rd(r.lo[1])),
                __float_as_int(__double2float_rd(r.lo[2])), __float_as_int(__double2float_ru(r.hi[0])));
            const int4 r1 = make_int4(__float_as_int(__double2float_ru(r.hi[1])), __float_as_int(__double2float_ru(r.hi[2])), b, 0);
            const int slot = atomicAdd(&cellCountOrCursor[cell], 1);
            rec0[slot] = r0;
            rec1[slot] = r1;
        }
        else atomicAdd(&cellCountOrCursor[cell], 1);
    }
}

int cub_scan_exclusive(idp_ctx* c, const int* in, int* out, long n)
{
    size_t bytes = 0;
    IDP_CK(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceScan::ExclusiveSum(c->cubTemp.p, bytes, in, out, (int)n, c->stream));
    ++c->lib_launches;
    return IDP_OK;
}

static long grid_cells(const GridDesc& g) { return (long)g.n[0] * g.n[1] * g.n[2]; }

// device region of the cells visited by the queries [qb, qe) of this shard (6 ints behind the histogram scratch)
static int set_query_region(idp_ctx* c, const IBox* qbox, int qb, int qe, const GridDesc& g, int** regionOut)
{
    int* dRegion = (int*)c->histScratch.p + 4 * IDP_LEVEL_BINS + 2;
    int init[6] = {0, 0, 0, g.n[0] - 1, g.n[1] - 1, g.n[2] - 1};
    const bool shard = c->nranks > 1;
    if (shard) { init[0] = init[1] = init[2] = 2147483647; init[3] = init[4] = init[5] = -1; }
    IDP_CK(c, cudaMemcpyAsync(dRegion, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    if (shard && qe > qb) IDP_LAUNCH(c, k_query_region, std::min(blocks_for(qe - qb, 256), (unsigned)c->sm_count * 4), 256, 0, qbox, qb, qe, g, dRegion);
    *regionOut = dRegion;
    return IDP_OK;
}
// Build the level table, cellStart (all levels concatenated, + 1) and the cell-sorted records for the insert boxes
// bbox[0..nB). If the primitives spread over more than IDP_MAX_LEVELS levels the base cells are coarsened (g.k doubled: a
// cell is k^3 lattice voxels) and the choice is repeated.
static int build_cells(idp_ctx* c, const IBox* bbox, const PrimRec* rec, int nB, GridDesc& g, const int latN[3], Levels* levelsOut,
    const IBox* qbox, int qb, int qe, bool ccd)
{
    int* dHist = (int*)c->histScratch.p; // 4 * IDP_LEVEL_BINS ints
    int hist[4 * IDP_LEVEL_BINS];
    Levels lv;
    bool ok = false;
    for (int attempt = 0; attempt < 16 && !ok; ++attempt) {
        IDP_CK(c, cudaMemsetAsync(dHist, 0, sizeof(hist), c->stream));
        IDP_LAUNCH(c, k_level_hist, std::min(blocks_for(nB, 256), (unsigned)c->sm_count * 8), 256, 0, bbox, nB, g, dHist);
        IDP_CK(c, cudaMemcpyAsync(hist, dHist, sizeof(hist), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        lv.n = 0;
        lv.useBin1 = 1;
        if (hist[1] > 0 && (long)hist[1] * 32 < nB && !getenv("IDP_KEEP_BIN1")) { // thin bin 1: its rows cost every query more than its records
            lv.useBin1 = 0;
            hist[2] += hist[1];
            hist[1] = 0;
            for (int k = 0; k < 3; ++k) hist[IDP_LEVEL_BINS + 6 + k] = std::max(hist[IDP_LEVEL_BINS + 6 + k], hist[IDP_LEVEL_BINS + 3 + k] ? 1 : 0);
        }
        long off = 0;
        ok = true;
        for (int l = 0; l < IDP_LEVEL_BINS; ++l) {
            lv.slot[l] = -1;
            if (hist[l] == 0 && !(l == 0 && nB == 0)) continue;
            if (lv.n == IDP_MAX_LEVELS) { ok = false; break; }
            const int i = lv.n++;
            lv.slot[l] = (signed char)i;
            lv.shift[i] = level_shift(l);
            for (int k = 0; k < 3; ++k) {
                lv.ext[i][k] = hist[IDP_LEVEL_BINS + 3 * l + k];
                lv.dim[i][k] = ((g.n[k] - 1) >> lv.shift[i]) + 1;
            }
            lv.off[i] = off;
            off += (long)lv.dim[i][0] * lv.dim[i][1] * lv.dim[i][2];
        }
        if (lv.n == 0) { lv.n = 1; lv.slot[0] = 0; lv.shift[0] = 0; lv.off[0] = 0; for (int k = 0; k < 3; ++k) { lv.ext[0][k] = 0; lv.dim[0][k] = g.n[k]; } }
        if (ok) break;
        g.k *= 2;
        for (int d = 0; d < 3; ++d) g.n[d] = std::max(1, (latN[d] + g.k) / g.k);
    }
    if (!ok) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "broad-phase grid could not be sized", __FILE__, __LINE__);
    for (int i = lv.n; i < IDP_MAX_LEVELS; ++i) { lv.shift[i] = 0; lv.off[i] = 0; for (int k = 0; k < 3; ++k) { lv.ext[i][k] = 0; lv.dim[i][k] = 1; } }
    const int last = lv.n - 1;
    const long ncAll = lv.off[last] + (long)lv.dim[last][0] * lv.dim[last][1] * lv.dim[last][2];
    if (getenv("IDP_DEBUG")) {
        fprintf(stderr, "[idp] build_cells(%s) nB=%d base grid %dx%dx%d k=%d, %d level(s):", ccd ? "ccd" : "static", nB, g.n[0], g.n[1], g.n[2], g.k, lv.n);
        for (int l = 0; l < IDP_LEVEL_BINS; ++l)
            if (lv.slot[l] >= 0) fprintf(stderr, " [shift=%d n=%d ext=%d%d%d]", level_shift(l), hist[l], lv.ext[lv.slot[l]][0], lv.ext[lv.slot[l]][1], lv.ext[lv.slot[l]][2]);
        fprintf(stderr, "\n");
    }
    int* dRegion = nullptr;
    IDP_TRY(set_query_region(c, qbox, qb, qe, g, &dRegion)); // after the grid is final (it may have been coarsened above)
    IDP_CK(c, c->cellStart.reserve(ncAll + 1));
    IDP_CK(c, c->cellCursor.reserve(ncAll + 1));
    IDP_CK(c, cudaMemsetAsync(c->cellCursor.p, 0, (ncAll + 1) * sizeof(int), c->stream));
    IDP_CK(c, c->crec0.reserve((size_t)std::max(nB, 1)));
    IDP_CK(c, c->crec1.reserve((size_t)std::max(nB, 1)));
    const unsigned grid = blocks_for(nB, 256);
    IDP_LAUNCH(c, (k_cells<false>), grid, 256, 0, bbox, rec, nB, g, lv, dRegion, c->cellCursor.p, (int4*)nullptr, (int4*)nullptr);
    IDP_TRY(cub_scan_exclusive(c, c->cellCursor.p, c->cellStart.p, ncAll + 1));
    IDP_CK(c, cudaMemcpyAsync(c->cellCursor.p, c->cellStart.p, (ncAll + 1) * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    IDP_LAUNCH(c, (k_cells<true>), grid, 256, 0, bbox, rec, nB, g, lv, dRegion, c->cellCursor.p, c->crec0.p, c->crec1.p);
    IDP_CK(c, cudaGetLastError());
    *levelsOut = lv;
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// broad-phase query: one warp per query primitive, records streamed in FLAT order.
// With anchor insertion the records a query must look at are the (y,z) rows of the cell block [qlo - emax, qhi]; each row
// is one contiguous run of the cell-sorted record arrays. The warp lists the non-empty runs (one lane per row: two
// cellStart loads), prefix-sums their lengths and then walks the concatenation of the runs 32 records at a time: lane j
// takes flat index i + j, finds its run from a 32-bit mask of the run starts inside the chunk (one REDUX.OR + popc) and
// reads record start[run] + (flat - offset[run]). Consecutive lanes read consecutive records, so the two 16-byte record
// streams are coalesced and every lane is busy except in the last chunk (the first version gave every lane its own run:
// 32 different cache lines per warp load and lanes idling on short runs -- 85x the algorithmic bytes through L1).
// Coarse filter per record: index order, then outward-rounded float AABB (static) or lattice-box overlap (CCD, exact).
// Survivors are pushed to a per-warp queue in shared memory with ballot/popc; whenever 32 are waiting the whole warp runs
// the exact predicate on them (one 64-byte PrimRec each: shared vertices, Dirichlet flags, FP64 AABB gap) and appends the
// hits with one atomic. Every pair is met exactly once, so there is no de-duplication.
// MODE bit0: 0 = point queries vs triangles, 1 = edge queries vs edges (partner index > query index);
// MODE bit1: CCD (lattice-box overlap = "shares a voxel" of the reference's hash; swept AABBs)
// ------------------------------------------------------------------------------------------------------------
struct QueryArgs {
    const PrimRec* qrec; const IBox* qbox; int qBegin, qEnd;
    const PrimRec* brec; const IBox* bbox;
    const int* cellStart; const int4* rec0; const int4* rec1;
    GridDesc g;
    double dist; // dHat (static) or thickness (CCD)
    Levels lv;
    int2* out; long cap; unsigned long long* counter;
};
// exact predicate on a (query, partner) pair that passed the coarse filter (IPC.h:171-172, 384-385; CCD.h:149-235)
template <int MODE>
__device__ __forceinline__ bool pair_exact(const QueryArgs& a, int b, const IBox& qb, const PrimRec& qr, const V3& qL, const V3& qH)
{
    const PrimRec br = a.brec[b];
    bool ok;
    if (MODE & 2) { // the reference's CCD hash only pairs primitives that share a voxel: lattice boxes of the alpha-scaled sweeps overlap
        const IBox bb = a.bbox[b];
        if (!(qb.lo[0] <= bb.hi[0] && bb.lo[0] <= qb.hi[0] && qb.lo[1] <= bb.hi[1] && bb.lo[1] <= qb.hi[1] && qb.lo[2] <= bb.hi[2] && bb.lo[2] <= qb.hi[2]))
            return false;
    }
    if (MODE & 1) ok = !(qr.v[0] == br.v[0] || qr.v[0] == br.v[1] || qr.v[1] == br.v[0] || qr.v[1] == br.v[1]); // shared vertex (IPC.h:384)
    else ok = !(qr.v[0] == br.v[0] || qr.v[0] == br.v[1] || qr.v[0] == br.v[2]);                                   // incident triangle (IPC.h:171)
    ok = ok && !((qr.flags & 1) && (br.flags & 1));                                                                  // all Dirichlet (:172, :385)
    return ok && aabb_gap_ok(qL, qH, mk3(br.lo[0], br.lo[1], br.lo[2]), mk3(br.hi[0], br.hi[1], br.hi[2]), a.dist);
}
#define IDP_QUERY_WARPS 8
// per-warp table of the rows a query visits, one entry per active level
struct LevelRows { int rowEnd; int x0, x1, y0, ny, z0; int nx, nyDim; long off; };
template <int MODE>
__global__ void __launch_bounds__(32 * IDP_QUERY_WARPS) k_query(QueryArgs a)
{
    __shared__ int squeue[IDP_QUERY_WARPS][96];
    __shared__ int2 sruns[IDP_QUERY_WARPS][32];
    __shared__ LevelRows slev[IDP_QUERY_WARPS][IDP_MAX_LEVELS];
    const int wib = threadIdx.x >> 5;
    const long warp0 = (long)blockIdx.x * IDP_QUERY_WARPS + wib;
    const long nWarps = (long)gridDim.x * IDP_QUERY_WARPS;
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u, le = lt | (1u << lane);
    int* sq = squeue[wib];
    int2* sr = sruns[wib];
    LevelRows* sl = slev[wib];
    const int nLev = a.lv.n;
    for (long q = a.qBegin + warp0; q < a.qEnd; q += nWarps) {
        const PrimRec qr = a.qrec[q];
        const IBox qb = a.qbox[q];
        int qlo0[3], qhi0[3];
        cell_range(a.g, qb, qlo0, qhi0);
        const V3 qL = mk3(qr.lo[0], qr.lo[1], qr.lo[2]), qH = mk3(qr.hi[0], qr.hi[1], qr.hi[2]);
        // coarse float thresholds: a record is dropped only if its outward-rounded box lies beyond qL - dist - eps (or
        // qH + dist + eps) on some axis, which implies the exact FP64 gap test (aabb_gap_ok) fails too
        float loT[3], hiT[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double eps = 1e-12 * (fabs(qr.lo[k]) + fabs(qr.hi[k]) + a.dist) + 1e-300;
            loT[k] = __double2float_rd(qr.lo[k] - a.dist - eps);
            hiT[k] = __double2float_ru(qr.hi[k] + a.dist + eps);
        }
        // lane l < nLev describes level l: the block of cells [qlo - ext, qhi] in level cells, rows = (y, z) pairs
        int myRows = 0;
        LevelRows lr;
        if (lane < nLev) {
            const int sh = a.lv.shift[lane];
            lr.x0 = max((qlo0[0] >> sh) - a.lv.ext[lane][0], 0); lr.x1 = qhi0[0] >> sh;
            lr.y0 = max((qlo0[1] >> sh) - a.lv.ext[lane][1], 0); lr.ny = (qhi0[1] >> sh) - lr.y0 + 1;
            lr.z0 = max((qlo0[2] >> sh) - a.lv.ext[lane][2], 0);
            lr.nx = a.lv.dim[lane][0]; lr.nyDim = a.lv.dim[lane][1]; lr.off = a.lv.off[lane];
            myRows = lr.ny * ((qhi0[2] >> sh) - lr.z0 + 1);
        }
        int rowEnd = myRows; // inclusive prefix over the levels
#pragma unroll
        for (int o = 1; o < IDP_MAX_LEVELS; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, rowEnd, o);
            if (lane >= o) rowEnd += t;
        }
        const int nRows = __shfl_sync(0xffffffffu, rowEnd, nLev - 1);
        __syncwarp();
        if (lane < nLev) { lr.rowEnd = rowEnd; sl[lane] = lr; }
        __syncwarp();
        int nq = 0; // survivors waiting in the queue (warp-uniform)
        for (int rbase = 0; rbase < nRows; rbase += 32) {
            // one lane per row: its run of records (rows of all levels in one list)
            const int r = rbase + lane;
            int start = 0, len = 0;
            if (r < nRows) {
                int L = 0;
                while (r >= sl[L].rowEnd) ++L;
                const LevelRows v = sl[L];
                const int rr = r - (L ? sl[L - 1].rowEnd : 0);
                const int rz = rr / v.ny;
                const int* cs = a.cellStart + v.off + ((long)(v.z0 + rz) * v.nyDim + (v.y0 + (rr - rz * v.ny))) * v.nx;
                start = __ldg(cs + v.x0);
                len = __ldg(cs + v.x1 + 1) - start;
            }
            int off = len; // inclusive prefix sum of the run lengths
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, off, o);
                if (lane >= o) off += t;
            }
            const int total = __shfl_sync(0xffffffffu, off, 31);
            off -= len; // exclusive
            const unsigned ne = __ballot_sync(0xffffffffu, len > 0);
            __syncwarp();
            if (len > 0) sr[__popc(ne & lt)] = make_int2(start, off); // non-empty runs, offsets strictly increasing
            __syncwarp();
            int runsBefore = 0;
            for (int i = 0; i < total; i += 64) { // 64 records per trip: two per lane, four 16-byte loads in flight
                const unsigned bit = len > 0 ? (unsigned)(off - i) : 64u;
                const unsigned m0 = __reduce_or_sync(0xffffffffu, bit < 32u ? (1u << bit) : 0u);
                const unsigned m1 = __reduce_or_sync(0xffffffffu, (bit >= 32u && bit < 64u) ? (1u << (bit - 32u)) : 0u);
                const int f0 = i + lane, f1 = f0 + 32;
                int4 w0 = make_int4(0, 0, 0, 0), w1 = w0, w2 = w0, w3 = w0;
                if (f0 < total) {
                    const int2 run = sr[runsBefore + __popc(m0 & le) - 1];
                    const int rec = run.x + (f0 - run.y);
                    w0 = __ldg(a.rec0 + rec); w1 = __ldg(a.rec1 + rec); // lo.xyz hi.x | hi.y hi.z b 0
                }
                if (f1 < total) {
                    const int2 run = sr[runsBefore + __popc(m0) + __popc(m1 & le) - 1];
                    const int rec = run.x + (f1 - run.y);
                    w2 = __ldg(a.rec0 + rec); w3 = __ldg(a.rec1 + rec);
                }
                runsBefore += __popc(m0) + __popc(m1);
                bool c0 = f0 < total && (!(MODE & 1) || w1.z > (int)q); // eJ > eI (IPC.h:384, SPATIAL_HASH.h:265)
                bool c1 = f1 < total && (!(MODE & 1) || w3.z > (int)q);
                c0 = c0 && !(__int_as_float(w0.w) < loT[0] || __int_as_float(w1.x) < loT[1] || __int_as_float(w1.y) < loT[2] ||
                             __int_as_float(w0.x) > hiT[0] || __int_as_float(w0.y) > hiT[1] || __int_as_float(w0.z) > hiT[2]);
                c1 = c1 && !(__int_as_float(w2.w) < loT[0] || __int_as_float(w3.x) < loT[1] || __int_as_float(w3.y) < loT[2] ||
                             __int_as_float(w2.x) > hiT[0] || __int_as_float(w2.y) > hiT[1] || __int_as_float(w2.z) > hiT[2]);
                const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
                if (b0 | b1) {
                    if (c0) sq[nq + __popc(b0 & lt)] = w1.z;
                    if (c1) sq[nq + __popc(b0) + __popc(b1 & lt)] = w3.z;
                    nq += __popc(b0) + __popc(b1);
                    while (nq >= 32) { // a full warp of survivors: exact predicate, append
                        __syncwarp();
                        const int bb = sq[nq - 32 + lane];
                        const bool hit = pair_exact<MODE>(a, bb, qb, qr, qL, qH);
                        const long slot = warp_append(hit, a.counter, a.cap);
                        if (hit && slot >= 0) a.out[slot] = make_int2((int)q, bb);
                        nq -= 32;
                    }
                    __syncwarp();
                }
            }
        }
        if (nq > 0) { // remaining survivors
            __syncwarp();
            const int bb = lane < nq ? sq[lane] : -1;
            const bool hit = lane < nq && pair_exact<MODE>(a, bb, qb, qr, qL, qH);
            const long slot = warp_append(hit, a.counter, a.cap);
            if (hit && slot >= 0) a.out[slot] = make_int2((int)q, bb);
            __syncwarp();
        }
    }
}

// runs the query kernel, growing the candidate buffer until everything fits
template <int MODE>
static int run_query(idp_ctx* c, QueryArgs a, DBuf<int2>& out, long* nOut)
{
    const long nQ = a.qEnd - a.qBegin;
    if (out.cap < 1024) IDP_CK(c, out.reserve(std::max<long>(1024, 4 * nQ)));
    unsigned long long* counter = (unsigned long long*)(c->counters.p + CNT_CAND);
    for (int attempt = 0; attempt < 8; ++attempt) {
        IDP_CK(c, cudaMemsetAsync(counter, 0, sizeof(long long), c->stream));
        a.out = out.p;
        a.cap = (long)out.cap;
        a.counter = counter;
        const unsigned grid = (unsigned)std::min<long>((nQ + 7) / 8, (long)c->sm_count * 64);
        if (nQ > 0) {
            KernelTimer kt(c, IDP_STAGE_K_QUERY);
            IDP_LAUNCH(c, k_query<MODE>, std::max(grid, 1u), 32 * IDP_QUERY_WARPS, 0, a);
        }
        IDP_CK(c, cudaGetLastError());
        long long n = 0;
        IDP_CK(c, cudaMemcpyAsync(&n, counter, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        if (n <= (long long)out.cap) { *nOut = (long)n; return IDP_OK; }
        IDP_CK(c, out.reserve((size_t)(n + n / 8)));
    }
    return fail(c, IDP_ERR_CUDA, "%s (%s:%d)", "candidate buffer did not converge", __FILE__, __LINE__);
}

// ------------------------------------------------------------------------------------------------------------
// narrow phase: classification + constraint-row encoding (IPC.h:198-267, 421-564)
// ------------------------------------------------------------------------------------------------------------
struct ClassifyArgs {
    const int2* cand; long nCand;
    const int* bnode; const int2* bedge; const int4* btri;
    const double4* xp; const double4* x0p;
    double dHat2;
    Row4* rowsDirect; long capDirect; unsigned long long* cntDirect;
    Row4* rowsDup; long capDup; unsigned long long* cntDup;
    // sort keys: direct rows are ordered by their candidate pair (query * nPartner + partner) = the reference's loop
    // order with ascending partners; duplicate rows by the std::map key order of (k0,k1,k2), packed when 3*dupBits <= 64
    unsigned long long* keysDirect; unsigned long long* keysDup; long long nPartner; int dupBits; long long nV;
};
__device__ __forceinline__ unsigned long long dup_key(const Row4& r, int bits, long long nV)
{
    const unsigned long long p = (unsigned long long)(nV - 1 - (long long)(-r.a - 1)); // k0 = -p-1 ascending <=> p descending
    return (p << (2 * bits)) | ((unsigned long long)r.b << bits) | (unsigned long long)(r.c + 1);
}
__global__ void __launch_bounds__(256, 3) k_classify_pt(ClassifyArgs a)
{
    const long stride = (long)gridDim.x * blockDim.x;
    const long nRound = (a.nCand + 31) / 32 * 32;
    const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int2 zero2 = make_int2(0, 0);
    int2 c1 = i0 < a.nCand ? a.cand[i0] : zero2;
    int2 c2 = i0 + stride < a.nCand ? a.cand[i0 + stride] : zero2;
    int v1 = a.nCand ? a.bnode[c1.x] : 0;
    int4 t1n = a.nCand ? a.btri[c1.y] : make_int4(0, 0, 0, 0);
    for (long i = i0; i < nRound; i += stride) {
        bool direct = false, dup = false;
        Row4 r = {0, 0, 0, 0};
        long long ckey = 0;
        const int2 c = c1;
        const int vI = v1;
        const int4 t = t1n;
        c1 = c2;
        v1 = a.bnode[c1.x]; t1n = a.btri[c1.y];
        c2 = i + 2 * stride < a.nCand ? a.cand[i + 2 * stride] : zero2;
        if (i < a.nCand) {
            ckey = (long long)c.x * a.nPartner + c.y;
            const V3 p = ldv(a.xp, vI), t0 = ldv(a.xp, t.x), t1 = ldv(a.xp, t.y), t2 = ldv(a.xp, t.z);
            const int ty = pt_type(p, t0, t1, t2);
            const double d = dist2_pt_by_type(ty, p, t0, t1, t2);
            if (d < a.dHat2) {
                r.a = -vI - 1;
                switch (ty) {
                case 0: r.b = t.x; r.c = -1; r.d = -1; break;
                case 1: r.b = t.y; r.c = -1; r.d = -1; break;
                case 2: r.b = t.z; r.c = -1; r.d = -1; break;
                case 3: r.b = t.x; r.c = t.y; r.d = -1; break;
                case 4: r.b = t.y; r.c = t.z; r.d = -1; break;
                case 5: r.b = t.z; r.c = t.x; r.d = -1; break;
                default: r.b = t.x; r.c = t.y; r.d = t.z; break;
                }
                direct = (ty == 6);
                dup = !direct;
            }
        }
        long s = warp_append(direct, a.cntDirect, a.capDirect);
        if (direct && s >= 0) { a.rowsDirect[s] = r; a.keysDirect[s] = (unsigned long long)ckey; }
        s = warp_append(dup, a.cntDup, a.capDup);
        if (dup && s >= 0) { if (a.dupBits) a.keysDup[s] = dup_key(r, a.dupBits, a.nV); else a.rowsDup[s] = r; }
    }
}
// The gathers of a candidate form a dependent chain candidate -> edge vertices -> positions; the first two links are
// fetched one and two iterations ahead (the kernel was long-scoreboard bound at 16 warps/SM).
__global__ void __launch_bounds__(256, 3) k_classify_ee(ClassifyArgs a)
{
    const long stride = (long)gridDim.x * blockDim.x;
    const long nRound = (a.nCand + 31) / 32 * 32;
    const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int2 zero2 = make_int2(0, 0);
    int2 c1 = i0 < a.nCand ? a.cand[i0] : zero2;                     // candidate of this iteration
    int2 c2 = i0 + stride < a.nCand ? a.cand[i0 + stride] : zero2;   // ... of the next one
    int2 ea1 = a.bedge[c1.x], eb1 = a.bedge[c1.y];
    for (long i = i0; i < nRound; i += stride) {
        bool direct = false, dup = false;
        Row4 r = {0, 0, 0, 0};
        long long ckey = 0;
        const int2 c = c1, ea = ea1, eb = eb1;
        c1 = c2;
        ea1 = a.bedge[c1.x]; eb1 = a.bedge[c1.y];
        c2 = i + 2 * stride < a.nCand ? a.cand[i + 2 * stride] : zero2;
        if (i < a.nCand) {
            ckey = (long long)c.x * a.nPartner + c.y;
            const V3 a0 = ldv(a.xp, ea.x), a1 = ldv(a.xp, ea.y), b0 = ldv(a.xp, eb.x), b1 = ldv(a.xp, eb.y);
            const int ty = ee_type(a0, a1, b0, b1);
            const double d = dist2_ee_by_type(ty, a0, a1, b0, b1);
            if (d < a.dHat2) {
                const double cn = ee_cross_norm2(a0, a1, b0, b1);
                const double eps_x = ee_mollifier_threshold(ldv(a.x0p, ea.x), ldv(a.x0p, ea.y), ldv(a.x0p, eb.x), ldv(a.x0p, eb.y));
                const bool moll = cn < eps_x;
                const int A0 = ea.x, A1 = ea.y, B0 = eb.x, B1 = eb.y;
                if (moll) {
                    switch (ty) {
                    case 0: r = {A0, B0, -A1 - 1, -B1 - 1}; break;
                    case 1: r = {A0, B1, -A1 - 1, -B0 - 1}; break;
                    case 2: r = {A0, B0, B1, -A1 - 1}; break;
                    case 3: r = {A1, B0, -A0 - 1, -B1 - 1}; break;
                    case 4: r = {A1, B1, -A0 - 1, -B0 - 1}; break;
                    case 5: r = {A1, B0, B1, -A0 - 1}; break;
                    case 6: r = {B0, A0, A1, -B1 - 1}; break;
                    case 7: r = {B1, A0, A1, -B0 - 1}; break;
                    default: r = {A0, A1, -B0 - 1, B1}; break;
                    }
                    direct = true;
                }
                else {
                    switch (ty) {
                    case 0: r = {-A0 - 1, B0, -1, -1}; break;
                    case 1: r = {-A0 - 1, B1, -1, -1}; break;
                    case 2: r = {-A0 - 1, B0, B1, -1}; break;
                    case 3: r = {-A1 - 1, B0, -1, -1}; break;
                    case 4: r = {-A1 - 1, B1, -1, -1}; break;
                    case 5: r = {-A1 - 1, B0, B1, -1}; break;
                    case 6: r = {-B0 - 1, A0, A1, -1}; break;
                    case 7: r = {-B1 - 1, A0, A1, -1}; break;
                    default: r = {A0, A1, B0, B1}; break;
                    }
                    direct = (ty == 8);
                    dup = !direct;
                }
            }
        }
        long s = warp_append(direct, a.cntDirect, a.capDirect);
        if (direct && s >= 0) { a.rowsDirect[s] = r; a.keysDirect[s] = (unsigned long long)ckey; }
        s = warp_append(dup, a.cntDup, a.capDup);
        if (dup && s >= 0) { if (a.dupBits) a.keysDup[s] = dup_key(r, a.dupBits, a.nV); else a.rowsDup[s] = r; }
    }
}

__global__ void k_emit_merged_keys(const unsigned long long* __restrict__ uniq, const int* __restrict__ counts, long n, int bits,
    long long nV, Row4* __restrict__ out)
{
    const unsigned long long mask = (1ull << bits) - 1ull;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const unsigned long long k = uniq[i];
        const long long p = nV - 1 - (long long)(k >> (2 * bits));
        Row4 r;
        r.a = (int)(-p - 1);
        r.b = (int)((k >> bits) & mask);
        r.c = (int)(k & mask) - 1;
        r.d = -counts[i];
        out[i] = r;
    }
}

// lower_bound of P+1 thresholds in a sorted key array (one thread each)
__global__ void k_key_splits(const unsigned long long* __restrict__ sorted, long n, const unsigned long long* __restrict__ thr, int nThr, long long* __restrict__ out)
{
    const int t = threadIdx.x;
    if (t >= nThr) return;
    long lo = 0, hi = n;
    const unsigned long long target = thr[t];
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (sorted[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    out[t] = lo;
}

__global__ void k_emit_merged(const Row4* __restrict__ uniq, const int* __restrict__ counts, long n, Row4* __restrict__ out)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        Row4 r = uniq[i];
        r.d = -counts[i];
        out[i] = r;
    }
}
__global__ void k_fill_double(double* __restrict__ p, long n, double v)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

static int sort_rows(idp_ctx* c, Row4* rows, long n)
{
    if (n <= 1) return IDP_OK;
    size_t bytes = 0;
    IDP_CK(c, cub::DeviceMergeSort::SortKeys(nullptr, bytes, rows, (int)n, RowLess(), c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceMergeSort::SortKeys(c->cubTemp.p, bytes, rows, (int)n, RowLess(), c->stream));
    ++c->lib_launches;
    return IDP_OK;
}

static int bits_for(unsigned long long maxval)
{
    int b = 1;
    while (b < 64 && (maxval >> b)) ++b;
    return b;
}
// stable radix sort of rows by 64-bit keys (in: rows/keys, out: rowsOut); bits = significant key bits
static int sort_rows_by_key(idp_ctx* c, unsigned long long* keys, Row4* rows, Row4* rowsOut, long n, int bits)
{
    if (n == 0) return IDP_OK;
    IDP_CK(c, c->keyTmp.reserve(n));
    size_t bytes = 0;
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, c->keyTmp.p, rows, rowsOut, (int)n, 0, bits, c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, keys, c->keyTmp.p, rows, rowsOut, (int)n, 0, bits, c->stream));
    ++c->lib_launches;
    return IDP_OK;
}

static int read_counters(idp_ctx* c)
{
    IDP_CK(c, cudaMemcpyAsync(c->h_counters, c->counters.p, CNT_COUNT * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    return IDP_OK;
}

// choose the static broad-phase lattice: spacing h ~ max(mean edge length, query radius), at most 2^26 cells
static void choose_static_grid(const double lo[3], const double hi[3], double meanEdge, double radius, GridDesc& g)
{
    double h = std::max(meanEdge, radius);
    if (!(h > 0)) h = 1.0;
    if (const char* e = getenv("IDP_CELL_SCALE")) h *= atof(e); // tuning knob (results do not depend on the grid)
    const double ext[3] = {hi[0] - lo[0] + 2 * radius, hi[1] - lo[1] + 2 * radius, hi[2] - lo[2] + 2 * radius};
    for (int iter = 0; iter < 64; ++iter) {
        double cells = 1;
        for (int k = 0; k < 3; ++k) cells *= std::floor(ext[k] / h) + 2;
        if (cells <= 67108864.0) break;
        h *= 1.26;
    }
    g.inv = 1.0 / h;
    g.k = 1;
    g.clampLat = 1;
    for (int k = 0; k < 3; ++k) {
        g.lo[k] = lo[k] - radius;
        g.n[k] = (int)std::floor(ext[k] / h) + 2;
    }
}

static int mean_edge_length(idp_ctx* c, double* out)
{
    if (c->nBE == 0) { *out = 0; return IDP_OK; }
    if (c->meanEdgeVersion == c->xVersion) { *out = c->meanEdgeCached; return IDP_OK; } // same positions as the last call (constraint set -> CCD)
    const long need = (long)c->nBE + 2 * ((c->nBE + 2047) / 2048 + 1);
    IDP_CK(c, c->red.reserve(std::max<long>(need, 3L * c->nBN + 2 * ((3L * c->nBN + 2047) / 2048 + 1))));
    IDP_LAUNCH(c, k_edge_lengths, blocks_for(c->nBE, 256), 256, 0, c->bedge.p, c->nBE, c->xp.p, c->red.p);
    double s = 0;
    const long nb = (c->nBE + 2047) / 2048 + 1;
    IDP_TRY(tree_sum_device(c, c->red.p, c->nBE, c->red.p + c->nBE, c->red.p + c->nBE + nb, &s));
    *out = s / c->nBE;
    c->meanEdgeCached = *out;
    c->meanEdgeVersion = c->xVersion;
    return IDP_OK;
}

static int node_bbox(idp_ctx* c, int mode, double alpha, double lo[3], double hi[3])
{
    unsigned long long init[6];
    for (int k = 0; k < 3; ++k) { init[k] = ~0ull; init[3 + k] = 0ull; }
    unsigned long long* d = (unsigned long long*)(c->counters.p + 8);
    IDP_CK(c, cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    IDP_LAUNCH(c, k_node_bbox, std::min(blocks_for(c->nBN, 256), (unsigned)c->sm_count * 8), 256, 0, c->bnode.p, c->nBN,
        c->xp.p, c->dp.p, alpha, mode, d);
    IDP_CK(c, cudaMemcpyAsync(init, d, sizeof(init), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; ++k) { lo[k] = dec_ord(init[k]); hi[k] = dec_ord(init[3 + k]); }
    return IDP_OK;
}

// shard [0, n) by contiguous primitive ranges
static void shard_range(const idp_ctx* c, long n, int* b, int* e)
{
    *b = (int)(n * c->rank / c->nranks);
    *e = (int)(n * (c->rank + 1) / c->nranks);
}

// Primitive records and lattice boxes for one hash build. Single GPU: everything. Sharded: this rank's query ranges, then
// only the partners whose insert box overlaps the lattice bounding box of those queries (PrepArgs); the rest is marked
// with the sentinel box and costs one 32-byte store instead of the 64 + 32 (+ 32) bytes of records and boxes.
static int prep_primitives(idp_ctx* c, PrepArgs pa, bool ccd)
{
    int nb, ne, eb, ee;
    shard_range(c, c->nBN, &nb, &ne);
    shard_range(c, c->nBE, &eb, &ee);
    const bool shard = c->nranks > 1;
    int* reg = nullptr;
    if (shard) {
        IDP_CK(c, c->prepRegion.reserve(12));
        static const int init[12] = {2147483647, 2147483647, 2147483647, -2147483647, -2147483647, -2147483647,
            2147483647, 2147483647, 2147483647, -2147483647, -2147483647, -2147483647};
        IDP_CK(c, cudaMemcpyAsync(c->prepRegion.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
        reg = c->prepRegion.p;
    }
    pa.phase = 0; pa.region = nullptr;
    pa.ownB = nb; pa.ownE = ne;
    if (ne > nb) IDP_LAUNCH(c, k_prep_nodes, blocks_for(ne - nb, 256), 256, 0, pa, c->nBN, c->recN.p, c->boxNq.p);
    pa.ownB = eb; pa.ownE = ee;
    if (ee > eb) IDP_LAUNCH(c, k_prep_edges, blocks_for(ee - eb, 256), 256, 0, pa, c->nBE, c->recE.p, c->boxEq.p, c->boxEb.p);
    if (shard) {
        if (ne > nb) IDP_LAUNCH(c, k_lat_region, std::min(blocks_for(ne - nb, 256), (unsigned)c->sm_count * 4), 256, 0, c->boxNq.p, nb, ne, reg);
        if (ee > eb) IDP_LAUNCH(c, k_lat_region, std::min(blocks_for(ee - eb, 256), (unsigned)c->sm_count * 4), 256, 0, ccd ? c->boxEb.p : c->boxEq.p, eb, ee, reg + 6);
        pa.phase = 1; pa.region = reg + 6;
        if (c->nBE - (ee - eb) > 0) IDP_LAUNCH(c, k_prep_edges, blocks_for(c->nBE - (ee - eb), 256), 256, 0, pa, c->nBE, c->recE.p, c->boxEq.p, c->boxEb.p);
        pa.region = reg;
    }
    if (c->nBT > 0) IDP_LAUNCH(c, k_prep_tris, blocks_for(c->nBT, 256), 256, 0, pa, c->nBT, c->recT.p, c->boxTb.p);
    IDP_CK(c, cudaGetLastError());
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Compute_Constraint_Set
// ------------------------------------------------------------------------------------------------------------
int build_constraint_set(idp_ctx* c, double dhat2_in, double thickness)
{
    if (c->pendRows) IDP_CK(c, cudaStreamWaitEvent(c->stream, c->evRows, 0)); // a pending idp_get_constraints_begin still reads the old rows
    c->permValid = false;
    c->dist2Valid = false;
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    if (!c->have_x0) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "rest positions not set", __FILE__, __LINE__);
    const double dHat = std::sqrt(dhat2_in) + thickness; // IPC.h:53-54
    const double dHat2 = dHat * dHat;
    c->cs_dhat2 = dHat2;
    timers_resolve(c);
    c->times.v[IDP_STAGE_COMM] = 0;
    c->times.v[IDP_STAGE_CCS_MERGE] = 0; // accumulated over the scopes of the merge stage
    IDP_CK(c, cudaMemsetAsync(c->counters.p, 0, CNT_COUNT * sizeof(long long), c->stream));

    GridDesc g;
    PrepArgs pa;
    int latN[3] = {1, 1, 1};
    {
        StageTimer tm(c, IDP_STAGE_CCS_BUILD_HASH);
        double meanEdge = 0, lo[3], hi[3];
        IDP_TRY(mean_edge_length(c, &meanEdge));
        IDP_TRY(node_bbox(c, 0, 0.0, lo, hi));
        double amax = 0;
        for (int k = 0; k < 3; ++k) amax = std::max(amax, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
        // conservative query radius: the exact predicate (AABB gap <= dHat) is applied afterwards
        const double radius = dHat * (1.0 + 1e-9) + amax * 1e-13;
        choose_static_grid(lo, hi, meanEdge, radius, g);
        for (int k = 0; k < 3; ++k) latN[k] = g.n[k];
        pa.bnode = c->bnode.p; pa.bedge = c->bedge.p; pa.btri = c->btri.p; pa.dbc = c->dbc.p;
        pa.xp = c->xp.p; pa.dp = nullptr; pa.g = g; pa.radius = radius; pa.vbox = nullptr;
        IDP_CK(c, c->recN.reserve(c->nBN)); IDP_CK(c, c->boxNq.reserve(c->nBN));
        IDP_CK(c, c->recE.reserve(c->nBE)); IDP_CK(c, c->boxEq.reserve(c->nBE)); IDP_CK(c, c->boxEb.reserve(c->nBE));
        IDP_CK(c, c->recT.reserve(c->nBT)); IDP_CK(c, c->boxTb.reserve(c->nBT));
        IDP_TRY(prep_primitives(c, pa, false));
    }
    unsigned long long* cnt = (unsigned long long*)c->counters.p;
    int qb, qe;
    // duplicate (PP / PE) rows are merged by sorting a packed (nV-1-p, k1, k2+1) key when it fits 64 bits
    const int vbits = bits_for((unsigned long long)c->nV);
    // IDP_FORCE_ROW_MERGE=1 selects the 16-byte row merge sort that meshes with more than 2^21 vertices need (tests)
    const int dupBits = (3 * vbits <= 64 && !getenv("IDP_FORCE_ROW_MERGE")) ? vbits : 0;
    {
        StageTimer tm(c, IDP_STAGE_CCS_PT);
        Levels lv;
        shard_range(c, c->nBN, &qb, &qe);
        IDP_TRY(build_cells(c, c->boxTb.p, c->recT.p, c->nBT, g, latN, &lv, c->boxNq.p, qb, qe, false));
        QueryArgs qa;
        qa.qrec = c->recN.p; qa.qbox = c->boxNq.p; qa.qBegin = qb; qa.qEnd = qe;
        qa.brec = c->recT.p; qa.bbox = c->boxTb.p; qa.cellStart = c->cellStart.p; qa.rec0 = c->crec0.p; qa.rec1 = c->crec1.p;
        qa.g = g; qa.dist = dHat; qa.lv = lv;
        IDP_TRY(run_query<0>(c, qa, c->candPT, &c->nCandPT));
        // classification
        IDP_CK(c, c->rowsA.reserve(std::max<long>(c->nCandPT, 1)));
        if (!dupBits) IDP_CK(c, c->rowsD.reserve(std::max<long>(c->nCandPT, 1)));
        ClassifyArgs ca;
        ca.cand = c->candPT.p; ca.nCand = c->nCandPT; ca.bnode = c->bnode.p; ca.bedge = c->bedge.p; ca.btri = c->btri.p;
        ca.xp = c->xp.p; ca.x0p = c->x0p.p; ca.dHat2 = dHat2;
        ca.rowsDirect = c->rowsA.p; ca.capDirect = (long)c->rowsA.cap; ca.cntDirect = cnt + CNT_ROWS_A;
        ca.rowsDup = c->rowsD.p; ca.capDup = (long)c->rowsD.cap; ca.cntDup = cnt + CNT_ROWS_D;
        IDP_CK(c, c->keyA.reserve(std::max<long>(c->nCandPT, 1)));
        IDP_CK(c, c->keyD.reserve(std::max<long>(c->nCandPT, 1)));
        ca.keysDirect = c->keyA.p; ca.keysDup = c->keyD.p; ca.nPartner = c->nBT; ca.dupBits = dupBits; ca.nV = c->nV;
        if (dupBits) ca.capDup = (long)c->keyD.cap;
        if (c->nCandPT > 0) {
            KernelTimer kt(c, IDP_STAGE_K_CLASSIFY);
            IDP_LAUNCH(c, k_classify_pt, std::min(blocks_for(c->nCandPT, 256), (unsigned)c->sm_count * 16), 256, 0, ca);
        }
        IDP_CK(c, cudaGetLastError());
        IDP_TRY(read_counters(c));
    }
    const long nA = (long)c->h_counters[CNT_ROWS_A];
    const long nD_pt = (long)c->h_counters[CNT_ROWS_D];
    long nB = 0, nD = 0;
    {
        StageTimer tm(c, IDP_STAGE_CCS_EE);
        Levels lv;
        shard_range(c, c->nBE, &qb, &qe);
        IDP_TRY(build_cells(c, c->boxEb.p, c->recE.p, c->nBE, g, latN, &lv, c->boxEq.p, qb, qe, false));
        QueryArgs qa;
        qa.qrec = c->recE.p; qa.qbox = c->boxEq.p; qa.qBegin = qb; qa.qEnd = qe;
        qa.brec = c->recE.p; qa.bbox = c->boxEb.p; qa.cellStart = c->cellStart.p; qa.rec0 = c->crec0.p; qa.rec1 = c->crec1.p;
        qa.g = g; qa.dist = dHat; qa.lv = lv;
        IDP_TRY(run_query<1>(c, qa, c->candEE, &c->nCandEE));
        IDP_CK(c, c->rowsB.reserve(std::max<long>(c->nCandEE, 1)));
        if (!dupBits) IDP_CK(c, c->rowsD.reserve(std::max<long>(nD_pt + c->nCandEE, 1), true, c->stream));
        IDP_CK(c, cudaMemsetAsync(c->counters.p + CNT_ROWS_A, 0, sizeof(long long), c->stream));
        ClassifyArgs ca;
        ca.cand = c->candEE.p; ca.nCand = c->nCandEE; ca.bnode = c->bnode.p; ca.bedge = c->bedge.p; ca.btri = c->btri.p;
        ca.xp = c->xp.p; ca.x0p = c->x0p.p; ca.dHat2 = dHat2;
        ca.rowsDirect = c->rowsB.p; ca.capDirect = (long)c->rowsB.cap; ca.cntDirect = cnt + CNT_ROWS_A;
        ca.rowsDup = c->rowsD.p; ca.capDup = (long)c->rowsD.cap; ca.cntDup = cnt + CNT_ROWS_D;
        IDP_CK(c, c->keyB.reserve(std::max<long>(c->nCandEE, 1)));
        IDP_CK(c, c->keyD.reserve(std::max<long>(nD_pt + c->nCandEE, 1), true, c->stream));
        ca.keysDirect = c->keyB.p; ca.keysDup = c->keyD.p; ca.nPartner = c->nBE; ca.dupBits = dupBits; ca.nV = c->nV;
        if (dupBits) ca.capDup = (long)c->keyD.cap;
        if (c->nCandEE > 0) {
            KernelTimer kt(c, IDP_STAGE_K_CLASSIFY);
            IDP_LAUNCH(c, k_classify_ee, std::min(blocks_for(c->nCandEE, 256), (unsigned)c->sm_count * 16), 256, 0, ca);
        }
        IDP_CK(c, cudaGetLastError());
        IDP_TRY(read_counters(c));
        nB = (long)c->h_counters[CNT_ROWS_A];
        nD = (long)c->h_counters[CNT_ROWS_D];
    }
    // ---- merge (IPC.h:571-661). No buffer is swapped or re-sized here in the steady state: repeated calls reuse the same
    // allocations (cudaMalloc / cudaFree of multi-GB buffers would dominate the step).
    const bool sharded = comm_on(c);
    long nAg = nA, nBg = nB, nDg = nD;
    IDP_CK(c, c->rows.reserve(std::max<long>(nA + nB + nD, 1)));
    {
        ScopeTimer tm(c, IDP_STAGE_CCS_MERGE, true);
        // direct groups: order by candidate pair (query-major, partner ascending = the reference's loop order) with one
        // radix sort each, written straight into their final place when there is a single rank. A shard's rows come from
        // its own contiguous query range, so concatenating the sorted shards in rank order is globally sorted.
        Row4* dstA = c->rows.p;
        Row4* dstB = c->rows.p + nA;
        if (sharded && !dupBits) { // fallback (keys wider than 64 bits): the row groups are all-gathered below
            IDP_CK(c, c->rowsG.reserve(std::max<long>(nA + nB, 1)));
            dstA = c->rowsG.p;
            dstB = c->rowsG.p + nA;
        }
        IDP_TRY(sort_rows_by_key(c, c->keyA.p, c->rowsA.p, dstA, nA, bits_for((unsigned long long)c->nBN * (unsigned long long)c->nBT)));
        IDP_TRY(sort_rows_by_key(c, c->keyB.p, c->rowsB.p, dstB, nB, bits_for((unsigned long long)c->nBE * (unsigned long long)c->nBE)));
    }
    unsigned long long* dupKeys = c->keyD.p;
    bool distributedDup = false;
    c->rowsLocal = false;
    if (sharded) {
        if (!dupBits) {
            IDP_TRY(comm_allgatherv(c, c->rowsG.p, nA, sizeof(Row4), (void**)&c->rows.p, &c->rows.cap, 0, &nAg));
            IDP_TRY(comm_allgatherv(c, c->rowsG.p + nA, nB, sizeof(Row4), (void**)&c->rows.p, &c->rows.cap, nAg, &nBg));
            // the unmerged PP / PE rows of every rank are gathered too (rowsG is free again) and merged on every rank
            IDP_TRY(comm_allgatherv(c, c->rowsD.p, nD, sizeof(Row4), (void**)&c->rowsG.p, &c->rowsG.cap, 0, &nDg));
        }
        if (dupBits) {
            // LOCAL-ROWS mode: every rank keeps (and later evaluates) the rows it produced -- its own direct rows and the
            // merged rows of its key range -- so nothing is replicated; the global list is only materialised on request
            // (idp_get_constraints). Rows born from a contiguous primitive range touch a contiguous vertex range, which
            // keeps the per-rank partial CSRs nearly disjoint.
            // distributed duplicate merge: sort the local keys, route every key to the rank owning its range of the leading
            // field (nV-1-p), merge there; the merged rows are gathered below in rank order = key order.
            const int P = c->nranks;
            long sendOff[IDP_MAX_RANKS + 1];
            {
            ScopeTimer tm2(c, IDP_STAGE_CCS_MERGE, true);
            IDP_CK(c, c->keyTmp.reserve(std::max<long>(nD, 1)));
            size_t bytes = 0;
            if (nD > 0) {
                // routing only needs the keys grouped by destination: order them by the leading field alone (3 radix passes
                // instead of 8); the owner sorts what it receives completely
                IDP_CK(c, cub::DeviceRadixSort::SortKeys(nullptr, bytes, c->keyD.p, c->keyTmp.p, (int)nD, 2 * dupBits, 3 * dupBits, c->stream));
                IDP_CK(c, c->cubTemp.reserve(bytes));
                IDP_CK(c, cub::DeviceRadixSort::SortKeys(c->cubTemp.p, bytes, c->keyD.p, c->keyTmp.p, (int)nD, 2 * dupBits, 3 * dupBits, c->stream));
                ++c->lib_launches;
            }
            unsigned long long thr[IDP_MAX_RANKS + 1];
            for (int r = 0; r <= P; ++r) thr[r] = (r == P) ? ~0ull : (((unsigned long long)((long long)c->nV * r / P)) << (2 * dupBits));
            thr[0] = 0;
            unsigned long long* dThr = (unsigned long long*)c->histScratch.p; // 64 ints = 32 ull of scratch
            long long* dSplit = (long long*)(dThr + 12);
            IDP_CK(c, cudaMemcpyAsync(dThr, thr, (P + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
            IDP_LAUNCH(c, k_key_splits, 1, 32, 0, c->keyTmp.p, nD, dThr, P + 1, dSplit);
            long long split[IDP_MAX_RANKS + 1];
            IDP_CK(c, cudaMemcpyAsync(split, dSplit, (P + 1) * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            IDP_CK(c, cudaStreamSynchronize(c->stream));
            for (int r = 0; r <= P; ++r) sendOff[r] = (long)split[r];
            sendOff[P] = nD;
            }
            // segment s of the leading field (nV-1-p ascending) holds the vertices of slab P-1-s: send it to that rank, so the
            // merged rows a rank evaluates touch the same vertex slab as its direct rows (near-disjoint partial CSRs)
            long sendBegin[IDP_MAX_RANKS], sendCount[IDP_MAX_RANKS];
            for (int r = 0; r < P; ++r) { sendBegin[r] = sendOff[P - 1 - r]; sendCount[r] = sendOff[P - r] - sendOff[P - 1 - r]; }
            IDP_TRY(comm_exchange_keys(c, c->keyTmp.p, sendBegin, sendCount, c->keyB, &nDg));
            dupKeys = c->keyB.p;
            distributedDup = true;
        }
        if (!distributedDup) IDP_CK(c, c->rows.reserve(std::max<long>(nAg + nBg + nDg, 1), true, c->stream));
    }
    {
        const long nA = nAg, nB = nBg, nD = nDg;
        ScopeTimer tm(c, IDP_STAGE_CCS_MERGE, true);
        long nU = 0;
        int* dRuns = (int*)(c->counters.p + CNT_RUNS);
        if (nD > 0 && dupBits) {
            // merged group: sort the packed keys (key order == std::map<VECTOR<int,4>> order, IPC.h:599-654), run lengths = multiplicities
            IDP_CK(c, c->keyTmp.reserve(nD));
            IDP_CK(c, c->keyA.reserve(nD));
            IDP_CK(c, c->runCounts.reserve(nD));
            size_t bytes = 0;
            IDP_CK(c, cub::DeviceRadixSort::SortKeys(nullptr, bytes, dupKeys, c->keyTmp.p, (int)nD, 0, 3 * dupBits, c->stream));
            IDP_CK(c, c->cubTemp.reserve(bytes));
            IDP_CK(c, cub::DeviceRadixSort::SortKeys(c->cubTemp.p, bytes, dupKeys, c->keyTmp.p, (int)nD, 0, 3 * dupBits, c->stream));
            IDP_CK(c, cub::DeviceRunLengthEncode::Encode(nullptr, bytes, c->keyTmp.p, c->keyA.p, c->runCounts.p, dRuns, (int)nD, c->stream));
            IDP_CK(c, c->cubTemp.reserve(bytes));
            IDP_CK(c, cub::DeviceRunLengthEncode::Encode(c->cubTemp.p, bytes, c->keyTmp.p, c->keyA.p, c->runCounts.p, dRuns, (int)nD, c->stream));
            c->lib_launches += 2;
            int runs = 0;
            IDP_CK(c, cudaMemcpyAsync(&runs, dRuns, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            IDP_CK(c, cudaStreamSynchronize(c->stream));
            nU = runs;
        }
        else if (nD > 0) {
            Row4* dupRows = sharded ? c->rowsG.p : c->rowsD.p;
            IDP_TRY(sort_rows(c, dupRows, nD));
            IDP_CK(c, c->rowsD2.reserve(nD));
            IDP_CK(c, c->runCounts.reserve(nD));
            size_t bytes = 0;
            IDP_CK(c, cub::DeviceRunLengthEncode::Encode(nullptr, bytes, dupRows, c->rowsD2.p, c->runCounts.p, dRuns, (int)nD, c->stream));
            IDP_CK(c, c->cubTemp.reserve(bytes));
            IDP_CK(c, cub::DeviceRunLengthEncode::Encode(c->cubTemp.p, bytes, dupRows, c->rowsD2.p, c->runCounts.p, dRuns, (int)nD, c->stream));
            ++c->lib_launches;
            int runs = 0;
            IDP_CK(c, cudaMemcpyAsync(&runs, dRuns, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            IDP_CK(c, cudaStreamSynchronize(c->stream));
            nU = runs;
        }
        if (distributedDup) {
            // this rank merged its own key range: its rows go behind its direct groups; only the counts are exchanged
            IDP_CK(c, c->rows.reserve(std::max<long>(nA + nB + nU, 1), true, c->stream));
            long long mine[3] = {nA, nB, nU};
            long long* dcnt = c->counters.p + CNT_SHARD; // 3 x P slots
            IDP_CK(c, cudaMemcpyAsync(dcnt + 3 * c->rank, mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
            IDP_TRY(comm_allgather_i64(c, dcnt, 3));
            long long all[3 * IDP_MAX_RANKS];
            IDP_CK(c, cudaMemcpyAsync(all, dcnt, 3 * c->nranks * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            IDP_CK(c, cudaStreamSynchronize(c->stream));
            c->nRowsGlobal = 0;
            for (int r = 0; r < c->nranks; ++r)
                for (int k = 0; k < 3; ++k) { c->shardCnt[r][k] = (long)all[3 * r + k]; c->nRowsGlobal += (long)all[3 * r + k]; }
            c->rowsLocal = true;
        }
        c->nRows = nA + nB + nU;
        if (c->nranks > 1 && !sharded) { // idp_set_shard without a communicator: the rows of this shard only, all evaluated here
            c->rowsLocal = true;
            c->nRowsGlobal = c->nRows;
        }
        IDP_CK(c, c->weights.reserve(std::max<long>(c->nRows, 1)));
        if (nU && dupBits) IDP_LAUNCH(c, k_emit_merged_keys, blocks_for(nU, 256), 256, 0, c->keyA.p, c->runCounts.p, nU, dupBits, (long long)c->nV, c->rows.p + nA + nB);
        else if (nU) IDP_LAUNCH(c, k_emit_merged, blocks_for(nU, 256), 256, 0, c->rowsD2.p, c->runCounts.p, nU, c->rows.p + nA + nB);
        if (c->nRows) IDP_LAUNCH(c, k_fill_double, blocks_for(c->nRows, 256), 256, 0, c->weights.p, c->nRows, 1.0); // OIPC: weight 1 (IPC.h:656-660)
        c->weights_all_one = true;
        IDP_CK(c, cudaGetLastError());
    }
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// sorted candidate sets for the parity check
// ------------------------------------------------------------------------------------------------------------
__global__ void k_pack_pairs(const int2* __restrict__ in, long n, unsigned long long* __restrict__ out)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = ((unsigned long long)(unsigned)in[i].x << 32) | (unsigned)in[i].y;
}
__global__ void k_unpack_pairs(const unsigned long long* __restrict__ in, long n, int2* __restrict__ out)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = make_int2((int)(in[i] >> 32), (int)(in[i] & 0xffffffffu));
}
int sorted_candidates(idp_ctx* c, int which, int2* host_out)
{
    DBuf<int2>& src = (which == 0 || which == 2) ? c->candPT : c->candEE;
    const long n = which == 0 ? c->nCandPT : (which == 1 ? c->nCandEE : (which == 2 ? c->nCcdPT : c->nCcdEE));
    if (n == 0) return IDP_OK;
    IDP_CK(c, c->blkKey.reserve(n));
    IDP_CK(c, c->blkKeySorted.reserve(n));
    IDP_LAUNCH(c, k_pack_pairs, blocks_for(n, 256), 256, 0, src.p, n, c->blkKey.p);
    size_t bytes = 0;
    // 64-bit item count: the CCD-only stress config produces up to 7 G swept-AABB candidates (> 2^31)
    IDP_CK(c, cub::DeviceRadixSort::SortKeys(nullptr, bytes, c->blkKey.p, c->blkKeySorted.p, (long long)n, 0, 64, c->stream));
    IDP_CK(c, c->cubTemp.reserve(bytes));
    IDP_CK(c, cub::DeviceRadixSort::SortKeys(c->cubTemp.p, bytes, c->blkKey.p, c->blkKeySorted.p, (long long)n, 0, 64, c->stream));
    ++c->lib_launches;
    IDP_LAUNCH(c, k_unpack_pairs, blocks_for(n, 256), 256, 0, c->blkKeySorted.p, n, (int2*)c->blkKey.p);
    IDP_CK(c, cudaMemcpyAsync(host_out, c->blkKey.p, n * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Compute_Min_Dist2
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_min_dist(const Row4* __restrict__ rows, long begin, long n, const double4* __restrict__ xp,
    double* __restrict__ dist2, unsigned long long* __restrict__ minOut)
{
    double m = INFINITY;
    for (long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const Row4 r = rows[i];
        const RowDec d = decode_row(r.a, r.b, r.c, r.d);
        const double d2 = row_dist2(d.kind, ldv(xp, d.v[0]), ldv(xp, d.v[1]), ldv(xp, d.v[2]), ldv(xp, d.v[3]));
        dist2[i] = d2;
        m = fmin(m, d2);
    }
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double bm = BR(tmp).Reduce(m, cub::Min());
    if (threadIdx.x == 0) atomicMin(minOut, enc_ord(bm));
}
int min_dist2(idp_ctx* c, double thickness, double* host_dist2, double* min_out)
{
    if ((c->rowsLocal ? c->nRowsGlobal : c->nRows) == 0) return IDP_OK; // IPC.h:2253-2255 (a shard without rows still joins the collectives)
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    StageTimer tm(c, IDP_STAGE_MIN_DIST);
    IDP_CK(c, c->rowDist2.reserve(std::max<long>(c->nRows, 1)));
    unsigned long long init = ~0ull;
    unsigned long long* d = (unsigned long long*)(c->counters.p + 8);
    IDP_CK(c, cudaMemcpyAsync(d, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    // sharded: in LOCAL-ROWS mode every rank scans its own rows; with replicated rows (idp_set_constraints) and only the
    // minimum wanted, a contiguous slice each. The order-encoded minima are combined with one all-reduce. The per-row
    // vector is the one of the rows this rank holds (idp_get_constraints order; idp_gather_constraints for the global one).
    const bool sharded = comm_on(c);
    const bool slice = sharded && !c->rowsLocal && !host_dist2;
    const long rb = slice ? c->nRows * c->rank / c->nranks : 0, re = slice ? c->nRows * (c->rank + 1) / c->nranks : c->nRows;
    if (re > rb)
        IDP_LAUNCH(c, k_min_dist, std::min(blocks_for(re - rb, 256), (unsigned)c->sm_count * 16), 256, 0, c->rows.p, rb, re, c->xp.p, c->rowDist2.p, d);
    IDP_CK(c, cudaGetLastError());
    c->dist2Valid = !slice;
    if (slice || (sharded && c->rowsLocal)) IDP_TRY(comm_allreduce_min_u64(c, d, 1));
    IDP_CK(c, cudaMemcpyAsync(c->h_red, d, sizeof(init), cudaMemcpyDeviceToHost, c->stream));
    if (host_dist2 && c->nRows) IDP_CK(c, cudaMemcpyAsync(host_dist2, c->rowDist2.p, c->nRows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    memcpy(&init, c->h_red, sizeof(init));
    *min_out = dec_ord(init) - thickness * thickness;
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Compute_Intersection_Free_StepSize
// ------------------------------------------------------------------------------------------------------------
struct AccdArgs {
    const int2* cand; long nCand;
    const int* bnode; const int2* bedge; const int4* btri;
    const double4* xp; const double4* dp;
    double eta, xi;
    unsigned long long* alphaBits; // positive double as ordered bits (plain bit pattern: positive doubles order as integers)
    unsigned long long* iters; unsigned long long* err;
};
struct BoundLoad {
    const unsigned long long* p;
    __device__ __forceinline__ double operator()() const { return __longlong_as_double((long long)*(volatile const unsigned long long*)p); }
};
template <int EE>
__global__ void __launch_bounds__(128) k_accd(AccdArgs a)
{
    const long stride = (long)gridDim.x * blockDim.x;
    unsigned long long myIters = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nCand; i += stride) {
        const int2 c = a.cand[i];
        int v0, v1, v2, v3;
        if (EE) { const int2 ea = a.bedge[c.x], eb = a.bedge[c.y]; v0 = ea.x; v1 = ea.y; v2 = eb.x; v3 = eb.y; }
        else { const int4 t = a.btri[c.y]; v0 = a.bnode[c.x]; v1 = t.x; v2 = t.y; v3 = t.z; }
        BoundLoad bound{a.alphaBits};
        double toc = 0;
        int it = 0, res;
        if (EE) res = accd_ee(ldv(a.xp, v0), ldv(a.xp, v1), ldv(a.xp, v2), ldv(a.xp, v3), ldv(a.dp, v0), ldv(a.dp, v1), ldv(a.dp, v2), ldv(a.dp, v3), a.eta, a.xi, bound, toc, it);
        else res = accd_pt(ldv(a.xp, v0), ldv(a.xp, v1), ldv(a.xp, v2), ldv(a.xp, v3), ldv(a.dp, v0), ldv(a.dp, v1), ldv(a.dp, v2), ldv(a.dp, v3), a.eta, a.xi, bound, toc, it);
        myIters += it;
        if (res == 1) atomicMin(a.alphaBits, (unsigned long long)__double_as_longlong(toc));
        else if (res < 0) atomicAdd(a.err, 1ull);
    }
    // warp-reduce the trip counter
    for (int o = 16; o > 0; o >>= 1) myIters += __shfl_down_sync(0xffffffffu, myIters, o);
    if (lane_id() == 0 && myIters) atomicAdd(a.iters, myIters);
}

// restatement of SPATIAL_HASH::Build (CCD), grid sizing part (:504-519)
static void ccd_set_grid(const double lo[3], const double hi[3], double voxelSize, GridDesc& g, long latN[3])
{
    const double range[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    double inv = 1.0 / voxelSize;
    long amt = 1;
    for (int d = 0; d < 3; ++d) amt *= std::max(1L, (long)std::ceil(range[d] * inv));
    if (amt > 1e9) {
        voxelSize *= std::pow(amt / 1.0e9, 1.0 / 3);
        inv = 1.0 / voxelSize;
    }
    int mn = 2147483647;
    int n[3];
    for (int d = 0; d < 3; ++d) {
        n[d] = std::max(1, (int)std::ceil(range[d] * inv));
        mn = std::min(mn, n[d]);
    }
    if (mn <= 0) {
        inv = 1.0 / (std::max(std::max(range[0], range[1]), range[2]) * 1.01);
        n[0] = n[1] = n[2] = 1;
    }
    g.inv = inv;
    g.clampLat = 0;
    for (int d = 0; d < 3; ++d) { g.lo[d] = lo[d]; latN[d] = n[d]; }
    // cells: k^3 lattice voxels each, at most 2^26 cells. Lattice index n[d] (max corner) is clamped into the last cell.
    int k = 1;
    while (true) {
        double cells = 1;
        for (int d = 0; d < 3; ++d) cells *= (double)((n[d] + k) / k);
        if (cells <= 67108864.0) break;
        ++k;
    }
    g.k = k;
    for (int d = 0; d < 3; ++d) g.n[d] = (n[d] + k) / k;
}

__global__ void k_poison_step_on_error(const unsigned long long* __restrict__ err, unsigned long long* __restrict__ alphaBits)
{
    if (*err) *alphaBits = 0ull;
}
int ccd_step(idp_ctx* c, double thickness, double* alpha_inout, int keep_candidates)
{
    (void)keep_candidates;
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    double alpha = *alpha_inout;
    c->nCcdPT = c->nCcdEE = 0;
    c->ccd_iters = 0;
    if (c->nBN == 0 || (c->nBT == 0 && c->nBE == 0)) return IDP_OK; // nothing can collide: the step is returned unchanged (IPC.h:1957-2243 loops are empty)
    IDP_CK(c, cudaMemsetAsync(c->counters.p, 0, CNT_COUNT * sizeof(long long), c->stream));
    GridDesc g;
    PrepArgs pa;
    int latN[3] = {1, 1, 1};
    {
        StageTimer tm(c, IDP_STAGE_CCD_BUILD_HASH);
        double voxelSize = 1.0;
        if (c->nBE) {
            double meanEdge = 0;
            IDP_TRY(mean_edge_length(c, &meanEdge));
            voxelSize *= meanEdge; // SPATIAL_HASH.h:455-464
        }
        // pSize = mean |dir component| over boundary nodes (:466-475); summed with the balanced tree, see DESIGN.md
        const long n3 = 3L * c->nBN;
        const long nb = (n3 + 2047) / 2048 + 1;
        IDP_CK(c, c->red.reserve(n3 + 2 * nb));
        IDP_LAUNCH(c, k_abs_dir, blocks_for(c->nBN, 256), 256, 0, c->bnode.p, c->nBN, c->dp.p, c->red.p);
        double pSum = 0;
        IDP_TRY(tree_sum_device(c, c->red.p, n3, c->red.p + n3, c->red.p + n3 + nb, &pSum));
        const double pSize = pSum / (double)n3;
        const double spanSize = alpha * pSize / voxelSize;
        if (spanSize > 1) {
            alpha /= spanSize; // :477-482
            // the reference sums |dir| serially; a differently ordered sum can differ in the last bits, so the clamped
            // step is shaved by 2^-30 (<< the 1e-6 budget) to guarantee it never exceeds the reference's.
            alpha *= (1.0 - 9.313225746154785e-10);
        }
        double lo[3], hi[3];
        IDP_TRY(node_bbox(c, 1, alpha, lo, hi));
        const double half = thickness / 2;
        for (int k = 0; k < 3; ++k) { lo[k] -= half; hi[k] += half; } // :502-503
        long latNl[3];
        ccd_set_grid(lo, hi, voxelSize, g, latNl);
        for (int k = 0; k < 3; ++k) latN[k] = (int)latNl[k];
        IDP_CK(c, c->vbox.reserve(c->nV));
        IDP_LAUNCH(c, k_ccd_vertex_boxes, blocks_for(c->nBN, 256), 256, 0, c->bnode.p, c->nBN, c->xp.p, c->dp.p, alpha, half, g, c->vbox.p);
        pa.bnode = c->bnode.p; pa.bedge = c->bedge.p; pa.btri = c->btri.p; pa.dbc = c->dbc.p;
        pa.xp = c->xp.p; pa.dp = c->dp.p; pa.g = g; pa.radius = 0; pa.vbox = c->vbox.p;
        IDP_CK(c, c->recN.reserve(c->nBN)); IDP_CK(c, c->boxNq.reserve(c->nBN));
        IDP_CK(c, c->recE.reserve(c->nBE)); IDP_CK(c, c->boxEq.reserve(c->nBE)); IDP_CK(c, c->boxEb.reserve(c->nBE));
        IDP_CK(c, c->recT.reserve(c->nBT)); IDP_CK(c, c->boxTb.reserve(c->nBT));
        IDP_TRY(prep_primitives(c, pa, true));
    }
    unsigned long long* cnt = (unsigned long long*)c->counters.p;
    {
        unsigned long long bits;
        memcpy(&bits, &alpha, 8);
        IDP_CK(c, cudaMemcpyAsync(cnt + CNT_ALPHA_BITS, &bits, 8, cudaMemcpyHostToDevice, c->stream));
    }
    AccdArgs aa;
    aa.bnode = c->bnode.p; aa.bedge = c->bedge.p; aa.btri = c->btri.p; aa.xp = c->xp.p; aa.dp = c->dp.p;
    aa.eta = 0.1; aa.xi = thickness; // IPC.h:2009, 2233
    aa.alphaBits = cnt + CNT_ALPHA_BITS; aa.iters = cnt + CNT_CCD_ITERS; aa.err = cnt + CNT_ERR_CCD;
    int qb, qe;
    {
        StageTimer tm(c, IDP_STAGE_CCD_PT);
        Levels lv;
        shard_range(c, c->nBN, &qb, &qe);
        IDP_TRY(build_cells(c, c->boxTb.p, c->recT.p, c->nBT, g, latN, &lv, c->boxNq.p, qb, qe, true));
        QueryArgs qa;
        qa.qrec = c->recN.p; qa.qbox = c->boxNq.p; qa.qBegin = qb; qa.qEnd = qe;
        qa.brec = c->recT.p; qa.bbox = c->boxTb.p; qa.cellStart = c->cellStart.p; qa.rec0 = c->crec0.p; qa.rec1 = c->crec1.p;
        qa.g = g; qa.dist = thickness; qa.lv = lv;
        IDP_TRY(run_query<2>(c, qa, c->candPT, &c->nCcdPT));
        aa.cand = c->candPT.p; aa.nCand = c->nCcdPT;
        if (c->nCcdPT > 0) IDP_LAUNCH(c, k_accd<0>, std::min(blocks_for(c->nCcdPT, 128), (unsigned)c->sm_count * 32), 128, 0, aa);
        IDP_CK(c, cudaGetLastError());
    }
    {
        StageTimer tm(c, IDP_STAGE_CCD_EE);
        Levels lv;
        shard_range(c, c->nBE, &qb, &qe);
        IDP_TRY(build_cells(c, c->boxEb.p, c->recE.p, c->nBE, g, latN, &lv, c->boxEb.p, qb, qe, true));
        QueryArgs qa;
        qa.qrec = c->recE.p; qa.qbox = c->boxEb.p; qa.qBegin = qb; qa.qEnd = qe;
        qa.brec = c->recE.p; qa.bbox = c->boxEb.p; qa.cellStart = c->cellStart.p; qa.rec0 = c->crec0.p; qa.rec1 = c->crec1.p;
        qa.g = g; qa.dist = thickness; qa.lv = lv;
        IDP_TRY(run_query<3>(c, qa, c->candEE, &c->nCcdEE));
        aa.cand = c->candEE.p; aa.nCand = c->nCcdEE;
        if (c->nCcdEE > 0) {
            KernelTimer kt(c, IDP_STAGE_K_ACCD);
            IDP_LAUNCH(c, k_accd<1>, std::min(blocks_for(c->nCcdEE, 128), (unsigned)c->sm_count * 32), 128, 0, aa);
        }
        IDP_CK(c, cudaGetLastError());
    }
    // positive doubles order like their bit patterns, so the device cell is a valid double for the min all-reduce; a rank
    // that hit the ACCD iteration cap zeroes its cell first, so the same reduction tells every rank (no second collective)
    if (comm_on(c)) {
        IDP_LAUNCH(c, k_poison_step_on_error, 1, 1, 0, cnt + CNT_ERR_CCD, cnt + CNT_ALPHA_BITS);
        IDP_TRY(comm_allreduce_min(c, (double*)(cnt + CNT_ALPHA_BITS), 1));
    }
    IDP_TRY(read_counters(c));
    c->ccd_iters = (long)c->h_counters[CNT_CCD_ITERS];
    double out;
    memcpy(&out, &c->h_counters[CNT_ALPHA_BITS], 8);
    // the PT and EE candidate buffers are invalidated for the static phase
    c->nCandPT = 0; c->nCandEE = 0;
    // every rank leaves with the same verdict (the step itself is already the global minimum)
    const int st = (c->h_counters[CNT_ERR_CCD] || (comm_on(c) && c->h_counters[CNT_ALPHA_BITS] == 0)) ? IDP_ERR_CCD_ITERATION_CAP : IDP_OK;
    if (st != IDP_OK) return fail(c, st, "%s (%s:%d)", "additive CCD iteration cap reached (on this or another rank)", __FILE__, __LINE__);
    *alpha_inout = out;
    if (out == 0 && alpha > 0) return fail(c, IDP_ERR_CCD_ZERO_STEP, "%s (%s:%d)", "CCD returned a zero step", __FILE__, __LINE__);
    return IDP_OK;
}

} // namespace idp
