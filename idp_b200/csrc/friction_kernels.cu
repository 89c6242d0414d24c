// Lagged friction on the device (Library/FEM/FRICTION.h:17-662): Compute_Friction_Basis freezes the
// non-mollified contact rows with their closest-point weights, tangent bases and normal forces (k_friction_basis, one thread
// per contact row); Compute_Friction_Potential / _Gradient / _Hessian are one thread per frozen row (k_friction), the
// Hessian blocks w_i w_j (T M2 T^T) going into the same per-vertex buckets as the contact rows, so the assembled CSR holds
// the friction term too (INC_POTENTIAL.h:375-377). Sharded by contiguous row ranges of the frozen list (replicated rows).
#include "ctx.cuh"
#include "friction.cuh"
#include "bucket_emit.cuh"
#include <cub/cub.cuh>

namespace idp {

__device__ __forceinline__ V3 ldf(const double4* __restrict__ p, int v)
{
    const double2* q = reinterpret_cast<const double2*>(p + v);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return mk3(a.x, a.y, b.x);
}

__global__ void __launch_bounds__(256) k_friction_basis(const Row4* __restrict__ rows, const double* __restrict__ weights, long n, const double4* __restrict__ xp,
    double dHat2, double kappa, double xi2, FricRow* __restrict__ out, unsigned long long* __restrict__ nActive)
{
    int mine = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const Row4 r = rows[i];
        const RowDec d = decode_row(r.a, r.b, r.c, r.d);
        const V3 x[4] = {ldf(xp, d.v[0]), ldf(xp, d.v[1]), ldf(xp, d.v[2]), ldf(xp, d.v[3])};
        FricRow f;
        friction_basis(d, x, weights[i], dHat2, kappa, xi2, f);
        out[i] = f;
        mine += f.nv > 0 ? 1 : 0;
    }
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int s = BR(tmp).Sum(mine);
    if (threadIdx.x == 0 && s) atomicAdd(nActive, (unsigned long long)s);
}

// Compute_Friction_Coef (FRICTION.h:126-170): the normal force of a row is scaled by the coefficient of the two components its
// first and its opposite primitive belong to (vertex v is in the first component whose upper bound exceeds it); mu becomes 1
__global__ void __launch_bounds__(256) k_friction_coef(FricRow* __restrict__ rows, long n, const int* __restrict__ compRange, const double* __restrict__ muComp,
    int nComp, unsigned long long* __restrict__ bad)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        FricRow& f = rows[i];
        if (f.nv == 0) continue;
        const int va = f.v[0], vb = f.kind == K_EE ? f.v[2] : f.v[1];
        int ca = -1, cb = -1;
        for (int c = nComp - 1; c >= 0; --c) { if (va < compRange[c]) ca = c; if (vb < compRange[c]) cb = c; }
        if (ca < 0 || cb < 0) { atomicAdd(bad, 1ull); continue; } // "can't find node compI"
        f.lam *= muComp[ca + cb * nComp];
    }
}

struct FrictionArgs {
    const FricRow* rows; long rBegin, rEnd;
    const double4* xp; const double4* xnp;
    double epsvh, mu;
    double* partialE; double* g;
    int* vtxCnt; int* vtxCursor; unsigned long long* bktKey; double* bktVal8; double* bktVal1;
    unsigned tagBase;
};

__global__ void __launch_bounds__(256) k_friction_counts(FrictionArgs a)
{
    for (long i = a.rBegin + (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.rEnd; i += (long)gridDim.x * blockDim.x) {
        const FricRow& f = a.rows[i];
        const int nv = f.nv;
        for (int k = 0; k < nv; ++k) {
            int m = 1;
            for (int j = 0; j < nv; ++j) m += f.v[j] > f.v[k] ? 1 : 0;
            atomicAdd(&a.vtxCnt[f.v[k]], m);
        }
    }
}

template <bool WANT_E, bool WANT_G, bool WANT_H>
__global__ void __launch_bounds__(128) k_friction(FrictionArgs a)
{
    double Eacc = 0;
    for (long i = a.rBegin + (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.rEnd; i += (long)gridDim.x * blockDim.x) {
        const FricRow f = a.rows[i];
        if (f.nv == 0) continue;
        V3 dx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dx[k] = mk3(0, 0, 0);
            if (k < f.nv) {
                const V3 x = ldf(a.xp, f.v[k]);
                dx[k] = (f.pp_abs && k == 1) ? x : x - ldf(a.xnp, f.v[k]); // the reference's PP rows read X for the second point
            }
        }
        if (WANT_E) Eacc += friction_energy(f, dx, a.epsvh, a.mu);
        if (WANT_G) {
            double g3[3];
            friction_gradient(f, dx, a.epsvh, a.mu, g3);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < f.nv) {
                    double* gp = a.g + 3 * (long)f.v[k];
                    atomicAdd(gp, f.w[k] * g3[0]); atomicAdd(gp + 1, f.w[k] * g3[1]); atomicAdd(gp + 2, f.w[k] * g3[2]);
                }
        }
        if (WANT_H) {
            double B[9];
            friction_hessian_core(f, dx, a.epsvh, a.mu, B);
            BucketEmit em;
            em.key = a.bktKey; em.val8 = a.bktVal8; em.val1 = a.bktVal1; em.nv = f.nv;
            em.v[0] = f.v[0]; em.v[1] = f.v[1]; em.v[2] = f.nv > 2 ? f.v[2] : -1; em.v[3] = f.nv > 3 ? f.v[3] : -1;
            em.rowTag = (a.tagBase + (unsigned)(i - a.rBegin)) << 4;
            em.reserve(a.vtxCursor);
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = p; q < 4; ++q) {
                    if (q >= f.nv) continue;
                    double blk[9];
                    const double s = f.w[p] * f.w[q];
#pragma unroll
                    for (int t = 0; t < 9; ++t) blk[t] = s * B[t];
                    em(p, q, blk);
                }
        }
    }
    if (WANT_E) {
        typedef cub::BlockReduce<double, 128> BR;
        __shared__ typename BR::TempStorage tmp;
        const double s = BR(tmp).Sum(Eacc);
        if (threadIdx.x == 0) a.partialE[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) k_friction_sum(const double* __restrict__ p, int n, double* __restrict__ out)
{
    double s = 0;
    for (int i = threadIdx.x; i < n; i += 256) s += p[i];
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double t = BR(tmp).Sum(s);
    if (threadIdx.x == 0) *out = t;
}

int friction_update(idp_ctx* c, double dhat2, double kappa, double thickness)
{
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    if (c->rowsLocal && comm_on(c)) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_friction_update: sharded contexts need replicated rows (idp_gather_constraints + idp_set_constraints)", __FILE__, __LINE__);
    c->nFric = 0; c->nFricActive = 0;
    if (c->nRows == 0) return IDP_OK;
    IDP_CK(c, c->fricRows.reserve((size_t)c->nRows * sizeof(FricRow)));
    unsigned long long* cnt = (unsigned long long*)(c->counters.p + 8);
    IDP_CK(c, cudaMemsetAsync(cnt, 0, sizeof(long long), c->stream));
    IDP_LAUNCH(c, k_friction_basis, std::min(blocks_for(c->nRows, 256), (unsigned)c->sm_count * 16), 256, 0, c->rows.p, c->weights.p, c->nRows, c->xp.p,
        dhat2 + 2 * std::sqrt(dhat2) * thickness, kappa, thickness * thickness, (FricRow*)c->fricRows.p, cnt);
    IDP_CK(c, cudaGetLastError());
    long long n = 0;
    IDP_CK(c, cudaMemcpyAsync(&n, cnt, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->nFric = c->nRows;
    c->nFricActive = (long)n;
    if (c->nFricComp > 0 && n > 0) {
        IDP_CK(c, cudaMemsetAsync(cnt, 0, sizeof(long long), c->stream));
        IDP_LAUNCH(c, k_friction_coef, std::min(blocks_for(c->nRows, 256), (unsigned)c->sm_count * 16), 256, 0, (FricRow*)c->fricRows.p, c->nRows, c->fricCompRange.p,
            c->fricMuComp.p, c->nFricComp, cnt);
        long long bad = 0;
        IDP_CK(c, cudaMemcpyAsync(&bad, cnt, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        if (bad) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_friction_update: a vertex lies beyond the last component bound (can't find node compI)", __FILE__, __LINE__);
    }
    return IDP_OK;
}

static FrictionArgs fric_args(idp_ctx* c)
{
    FrictionArgs a = {};
    a.rows = (const FricRow*)c->fricRows.p;
    a.rBegin = c->nFric * c->rank / c->nranks; a.rEnd = c->nFric * (c->rank + 1) / c->nranks;
    a.xp = c->xp.p; a.xnp = c->xnp.p; a.epsvh = c->fricEpsvh; a.mu = c->fricMu;
    return a;
}
static bool friction_on(const idp_ctx* c) { return c->nFricActive > 0 && c->fricMu > 0 && c->have_xn; }

int friction_block_counts(idp_ctx* c, int* vtxCnt, long* nRows)
{
    *nRows = 0;
    if (!friction_on(c)) return IDP_OK;
    FrictionArgs a = fric_args(c);
    *nRows = a.rEnd - a.rBegin;
    if (*nRows == 0) return IDP_OK;
    a.vtxCnt = vtxCnt;
    IDP_LAUNCH(c, k_friction_counts, std::min(blocks_for(*nRows, 256), (unsigned)c->sm_count * 16), 256, 0, a);
    return IDP_OK;
}
int friction_emit_blocks(idp_ctx* c, unsigned tagBase, int* vtxCursor, unsigned long long* bktKey, double* bktVal8, double* bktVal1)
{
    FrictionArgs a = fric_args(c);
    a.vtxCursor = vtxCursor; a.bktKey = bktKey; a.bktVal8 = bktVal8; a.bktVal1 = bktVal1; a.tagBase = tagBase;
    if (a.rEnd > a.rBegin) IDP_LAUNCH(c, (k_friction<false, false, true>), std::min(blocks_for(a.rEnd - a.rBegin, 128), (unsigned)c->sm_count * 16), 128, 0, a);
    IDP_CK(c, cudaGetLastError());
    return IDP_OK;
}
int friction_energy_gradient(idp_ctx* c, int want_e, int want_g, double* E_out)
{
    if (E_out) *E_out = 0;
    if (want_g) {
        IDP_CK(c, c->fricG.reserve(3 * (size_t)c->nV));
        IDP_CK(c, cudaMemsetAsync(c->fricG.p, 0, 3 * (size_t)c->nV * sizeof(double), c->stream));
    }
    if (!friction_on(c)) return IDP_OK;
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    FrictionArgs a = fric_args(c);
    const long n = a.rEnd - a.rBegin;
    const unsigned grid = std::max(1u, std::min(blocks_for(n, 128), (unsigned)c->sm_count * 16));
    IDP_CK(c, c->red.reserve((size_t)grid + 8));
    IDP_CK(c, cudaMemsetAsync(c->red.p, 0, ((size_t)grid + 1) * sizeof(double), c->stream));
    a.partialE = c->red.p; a.g = c->fricG.p;
    if (n > 0) {
        if (want_e && want_g) IDP_LAUNCH(c, (k_friction<true, true, false>), grid, 128, 0, a);
        else if (want_e) IDP_LAUNCH(c, (k_friction<true, false, false>), grid, 128, 0, a);
        else if (want_g) IDP_LAUNCH(c, (k_friction<false, true, false>), grid, 128, 0, a);
    }
    IDP_CK(c, cudaGetLastError());
    if (want_e) {
        IDP_LAUNCH(c, k_friction_sum, 1, 256, 0, c->red.p, (int)grid, c->red.p + grid);
        if (comm_on(c)) IDP_TRY(comm_allreduce_sum(c, c->red.p + grid, 1));
        double E = 0;
        IDP_CK(c, cudaMemcpyAsync(&E, c->red.p + grid, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        if (E_out) *E_out = E;
    }
    if (want_g && comm_on(c)) IDP_TRY(comm_allreduce_sum(c, c->fricG.p, 3L * c->nV));
    return IDP_OK;
}

// the frozen rows in the reference's layout (the non-mollified contact rows, in order)
int friction_copy_rows(idp_ctx* c, int* rows4, double* closest2, double* basis6, double* normalForce)
{
    if (c->nFric == 0) return IDP_OK;
    std::vector<FricRow> h((size_t)c->nFric);
    IDP_CK(c, cudaMemcpyAsync(h.data(), c->fricRows.p, h.size() * sizeof(FricRow), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    long k = 0;
    for (long i = 0; i < c->nFric; ++i) {
        const FricRow& f = h[i];
        if (f.nv == 0) continue;
        if (rows4) { // the contact row as it was (SURVEY.md A.1 encoding)
            int* r = rows4 + 4 * k;
            if (f.kind == K_EE) { r[0] = f.v[0]; r[1] = f.v[1]; r[2] = f.v[2]; r[3] = f.v[3]; }
            else if (f.kind == K_PT) { r[0] = -f.v[0] - 1; r[1] = f.v[1]; r[2] = f.v[2]; r[3] = f.v[3]; }
            else if (f.kind == K_PE) { r[0] = -f.v[0] - 1; r[1] = f.v[1]; r[2] = f.v[2]; r[3] = -f.mult; }
            else { r[0] = -f.v[0] - 1; r[1] = f.v[1]; r[2] = -1; r[3] = -f.mult; }
        }
        if (closest2) { closest2[2 * k] = f.cp[0]; closest2[2 * k + 1] = f.cp[1]; }
        if (basis6) for (int a = 0; a < 3; ++a) { basis6[6 * k + a] = f.t0[a]; basis6[6 * k + 3 + a] = f.t1[a]; }
        if (normalForce) normalForce[k] = f.lam / (double)f.mult; // the reference applies the multiplicity when it evaluates the force
        ++k;
    }
    return IDP_OK;
}

} // namespace idp
