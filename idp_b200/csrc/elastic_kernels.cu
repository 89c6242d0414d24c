// Elastic terms of the discrete-shell Newton system on the device: membrane triangles and bending
// hinges, one thread per element, energy / gradient on request and PSD-projected Hessian blocks emitted into the SAME
// per-vertex buckets as the contact rows, so the assembled CSR is the whole system matrix of
// Compute_IncPotential_Hessian (Library/FEM/Shell/INC_POTENTIAL.h:340-394, non-flow branch):
//   Compute_Membrane_Energy / _Gradient / _Hessian   FEM/Shell/MEMBRANE.h:8-315
//   Compute_Bending_Energy / _Gradient / _Hessian    FEM/Shell/BENDING.h:52-80, 176-213, 438-497 (KL = false)
// Elements whose vertices are all Dirichlet nodes are skipped like the reference skips them (MEMBRANE.h:25, BENDING.h:56-60).
// Compiled with --fmad=false: the dihedral angle is an acos of a clamped cosine and only reproduces the reference's value
// where the same roundings are made (shell_elastic.cuh).
#include "ctx.cuh"
#include "shell_elastic.cuh"
#include "bucket_emit.cuh"
#include <cub/cub.cuh>

namespace idp {

__device__ __forceinline__ V3 ldx(const double4* __restrict__ p, int v)
{
    const double2* q = reinterpret_cast<const double2*>(p + v);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return mk3(a.x, a.y, b.x);
}

struct ElasticArgs {
    const double4* xp; const unsigned char* dbc;
    const int* memElem; const double* memIB; const double* memCoef; const double* memLambda; const double* memMu; int mBegin, mEnd;
    const int* hingeV; const double* hingeInfo; double hingeKh2; int hBegin, hEnd;
    double* partialE; double* g;
    int* vtxCnt; int* vtxCursor; unsigned long long* bktKey; double* bktVal8; double* bktVal1;
    unsigned tagBase;
    unsigned long long* errEig;
    int projectSPD;
};

__device__ __forceinline__ bool membrane_active(const ElasticArgs& a, int e, int (&v)[3])
{
    v[0] = a.memElem[3 * e]; v[1] = a.memElem[3 * e + 1]; v[2] = a.memElem[3 * e + 2];
    if (a.dbc[v[0]] && a.dbc[v[1]] && a.dbc[v[2]]) return false;
    const double b0 = a.memIB[3 * e], b1 = a.memIB[3 * e + 1], b2 = a.memIB[3 * e + 2];
    return b0 * b2 - b1 * b1 != 0.0;
}
__device__ __forceinline__ bool hinge_active(const ElasticArgs& a, int e, int (&v)[4])
{
    v[0] = a.hingeV[4 * e]; v[1] = a.hingeV[4 * e + 1]; v[2] = a.hingeV[4 * e + 2]; v[3] = a.hingeV[4 * e + 3];
    return !(a.dbc[v[0]] && a.dbc[v[1]] && a.dbc[v[2]] && a.dbc[v[3]]);
}

// bucket sizes: an active element contributes, to the bucket of each of its vertices, the diagonal block plus one block per
// stencil vertex with a larger id (the upper triangle of its local Hessian)
__global__ void __launch_bounds__(256) k_elastic_counts(ElasticArgs a)
{
    const int nM = a.mEnd - a.mBegin, nH = a.hEnd - a.hBegin;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nM + nH; t += gridDim.x * blockDim.x) {
        if (t < nM) {
            int v[3];
            if (!membrane_active(a, a.mBegin + t, v)) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                int m = 1;
#pragma unroll
                for (int j = 0; j < 3; ++j) m += v[j] > v[k] ? 1 : 0;
                atomicAdd(&a.vtxCnt[v[k]], m);
            }
        }
        else {
            int v[4];
            if (!hinge_active(a, a.hBegin + (t - nM), v)) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int m = 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) m += v[j] > v[k] ? 1 : 0;
                atomicAdd(&a.vtxCnt[v[k]], m);
            }
        }
    }
}

struct NoEmit {
    __device__ __forceinline__ bool wants(int, int) const { return false; }
    __device__ __forceinline__ void operator()(int, int, const double*) const {}
};

constexpr int ELASTIC_T = 128;

// membrane triangles: 6x6 projection store in shared memory (QlStore<6>)
template <bool WANT_E, bool WANT_G, bool WANT_H>
__global__ void __launch_bounds__(ELASTIC_T) k_membrane(ElasticArgs a)
{
    extern __shared__ double sV[];
    double Eacc = 0;
    for (int e = a.mBegin + blockIdx.x * blockDim.x + threadIdx.x; e < a.mEnd; e += gridDim.x * blockDim.x) {
        int v[3];
        if (!membrane_active(a, e, v)) continue;
        const V3 x[3] = {ldx(a.xp, v[0]), ldx(a.xp, v[1]), ldx(a.xp, v[2])};
        const double ib[3] = {a.memIB[3 * e], a.memIB[3 * e + 1], a.memIB[3 * e + 2]};
        QlStore<6, ELASTIC_T> V6{sV + threadIdx.x};
        ElasticOut out;
        if (WANT_H) {
            BucketEmit em;
            em.key = a.bktKey; em.val8 = a.bktVal8; em.val1 = a.bktVal1; em.nv = 3; em.v[0] = v[0]; em.v[1] = v[1]; em.v[2] = v[2]; em.v[3] = -1;
            em.rowTag = (a.tagBase + (unsigned)(e - a.mBegin)) << 4;
            em.reserve(a.vtxCursor);
            membrane_eval(x, ib, a.memCoef[e], a.memLambda[e], a.memMu[e], a.projectSPD != 0, WANT_G, true, V6, out, em);
            if (out.eigFail) atomicAdd(a.errEig, 1ull);
        }
        else {
            NoEmit em;
            membrane_eval(x, ib, a.memCoef[e], a.memLambda[e], a.memMu[e], false, WANT_G, false, V6, out, em);
        }
        if (WANT_E) Eacc += out.E;
        if (WANT_G) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double* gp = a.g + 3 * (long)v[k];
                atomicAdd(gp, out.g[3 * k]); atomicAdd(gp + 1, out.g[3 * k + 1]); atomicAdd(gp + 2, out.g[3 * k + 2]);
            }
        }
    }
    if (WANT_E) {
        typedef cub::BlockReduce<double, ELASTIC_T> BR;
        __shared__ typename BR::TempStorage tmp;
        const double s = BR(tmp).Sum(Eacc);
        if (threadIdx.x == 0) a.partialE[blockIdx.x] = s;
    }
}

// bending hinges: 9x9 projection store in shared memory (QlStore<9>)
template <bool WANT_E, bool WANT_G, bool WANT_H>
__global__ void __launch_bounds__(ELASTIC_T) k_hinge(ElasticArgs a)
{
    extern __shared__ double sV[];
    double Eacc = 0;
    const int nM = a.mEnd - a.mBegin;
    for (int e = a.hBegin + blockIdx.x * blockDim.x + threadIdx.x; e < a.hEnd; e += gridDim.x * blockDim.x) {
        int v[4];
        if (!hinge_active(a, e, v)) continue;
        const V3 x[4] = {ldx(a.xp, v[0]), ldx(a.xp, v[1]), ldx(a.xp, v[2]), ldx(a.xp, v[3])};
        const double thetabar = a.hingeInfo[3 * e], coef = a.hingeKh2 * a.hingeInfo[3 * e + 1] / a.hingeInfo[3 * e + 2];
        QlStore<9, ELASTIC_T> V9{sV + threadIdx.x};
        ElasticOut out;
        if (WANT_H) {
            BucketEmit em;
            em.key = a.bktKey; em.val8 = a.bktVal8; em.val1 = a.bktVal1; em.nv = 4; em.v[0] = v[0]; em.v[1] = v[1]; em.v[2] = v[2]; em.v[3] = v[3];
            em.rowTag = (a.tagBase + (unsigned)nM + (unsigned)(e - a.hBegin)) << 4;
            em.reserve(a.vtxCursor);
            hinge_eval(x, thetabar, coef, a.projectSPD != 0, WANT_G, true, V9, out, em);
            if (out.eigFail) atomicAdd(a.errEig, 1ull);
        }
        else {
            NoEmit em;
            hinge_eval(x, thetabar, coef, false, WANT_G, false, V9, out, em);
        }
        if (WANT_E) Eacc += out.E;
        if (WANT_G) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double* gp = a.g + 3 * (long)v[k];
                atomicAdd(gp, out.g[3 * k]); atomicAdd(gp + 1, out.g[3 * k + 1]); atomicAdd(gp + 2, out.g[3 * k + 2]);
            }
        }
    }
    if (WANT_E) {
        typedef cub::BlockReduce<double, ELASTIC_T> BR;
        __shared__ typename BR::TempStorage tmp;
        const double s = BR(tmp).Sum(Eacc);
        if (threadIdx.x == 0) a.partialE[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) k_elastic_sum(const double* __restrict__ p, int n, double* __restrict__ out)
{
    double s = 0;
    for (int i = threadIdx.x; i < n; i += 256) s += p[i];
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double t = BR(tmp).Sum(s);
    if (threadIdx.x == 0) *out = t;
}

static ElasticArgs base_args(idp_ctx* c)
{
    ElasticArgs a = {};
    a.xp = c->xp.p; a.dbc = c->dbc.p;
    a.memElem = c->memElem.p; a.memIB = c->memIB.p; a.memCoef = c->memCoef.p; a.memLambda = c->memLambda.p; a.memMu = c->memMu.p;
    a.hingeV = c->hingeV.p; a.hingeInfo = c->hingeInfo.p; a.hingeKh2 = c->hingeKh2;
    // sharded by contiguous element ranges, like the flow term and the query primitives
    a.mBegin = (int)((long)c->nMem * c->rank / c->nranks); a.mEnd = (int)((long)c->nMem * (c->rank + 1) / c->nranks);
    a.hBegin = (int)((long)c->nHinge * c->rank / c->nranks); a.hEnd = (int)((long)c->nHinge * (c->rank + 1) / c->nranks);
    return a;
}

int elastic_block_counts(idp_ctx* c, int* vtxCnt, long* nElements)
{
    ElasticArgs a = base_args(c);
    const long n = (long)(a.mEnd - a.mBegin) + (a.hEnd - a.hBegin);
    *nElements = n;
    if (n == 0) return IDP_OK;
    a.vtxCnt = vtxCnt;
    IDP_LAUNCH(c, k_elastic_counts, std::min(blocks_for(n, 256), (unsigned)c->sm_count * 16), 256, 0, a);
    return IDP_OK;
}

int elastic_emit_blocks(idp_ctx* c, int project_spd, unsigned tagBase, int* vtxCursor, unsigned long long* bktKey, double* bktVal8, double* bktVal1)
{
    ElasticArgs a = base_args(c);
    a.vtxCursor = vtxCursor; a.bktKey = bktKey; a.bktVal8 = bktVal8; a.bktVal1 = bktVal1; a.tagBase = tagBase; a.projectSPD = project_spd;
    a.errEig = (unsigned long long*)(c->counters.p + CNT_ERR_EIG);
    const int nM = a.mEnd - a.mBegin, nH = a.hEnd - a.hBegin;
    if (nM > 0) {
        const size_t smem = QlStore<6, ELASTIC_T>::WORDS * sizeof(double) * ELASTIC_T;
        IDP_CK(c, cudaFuncSetAttribute(k_membrane<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IDP_LAUNCH(c, (k_membrane<false, false, true>), std::min(blocks_for(nM, ELASTIC_T), (unsigned)c->sm_count * 8), ELASTIC_T, smem, a);
    }
    if (nH > 0) {
        const size_t smem = QlStore<9, ELASTIC_T>::WORDS * sizeof(double) * ELASTIC_T;
        IDP_CK(c, cudaFuncSetAttribute(k_hinge<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IDP_LAUNCH(c, (k_hinge<false, false, true>), std::min(blocks_for(nH, ELASTIC_T), (unsigned)c->sm_count * 8), ELASTIC_T, smem, a);
    }
    IDP_CK(c, cudaGetLastError());
    return IDP_OK;
}

// E (sum over this rank's elements; sharded: all-reduced) and / or the gradient into c->elasticG (3 nV, all-reduced)
int elastic_energy_gradient(idp_ctx* c, int want_e, int want_g, double* E_out)
{
    if (!c->have_x) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "positions not set", __FILE__, __LINE__);
    if (E_out) *E_out = 0;
    if (want_g) {
        IDP_CK(c, c->elasticG.reserve(3 * (size_t)c->nV));
        IDP_CK(c, cudaMemsetAsync(c->elasticG.p, 0, 3 * (size_t)c->nV * sizeof(double), c->stream));
    }
    ElasticArgs a = base_args(c);
    const int nM = a.mEnd - a.mBegin, nH = a.hEnd - a.hBegin;
    const unsigned gM = std::max(1u, std::min(blocks_for(nM, ELASTIC_T), (unsigned)c->sm_count * 8)), gH = std::max(1u, std::min(blocks_for(nH, ELASTIC_T), (unsigned)c->sm_count * 8));
    IDP_CK(c, c->red.reserve((size_t)gM + gH + 8));
    IDP_CK(c, cudaMemsetAsync(c->red.p, 0, ((size_t)gM + gH + 1) * sizeof(double), c->stream));
    a.g = c->elasticG.p;
    if (nM > 0) {
        a.partialE = c->red.p;
        if (want_e && want_g) IDP_LAUNCH(c, (k_membrane<true, true, false>), gM, ELASTIC_T, 0, a);
        else if (want_e) IDP_LAUNCH(c, (k_membrane<true, false, false>), gM, ELASTIC_T, 0, a);
        else if (want_g) IDP_LAUNCH(c, (k_membrane<false, true, false>), gM, ELASTIC_T, 0, a);
    }
    if (nH > 0) {
        a.partialE = c->red.p + gM;
        if (want_e && want_g) IDP_LAUNCH(c, (k_hinge<true, true, false>), gH, ELASTIC_T, 0, a);
        else if (want_e) IDP_LAUNCH(c, (k_hinge<true, false, false>), gH, ELASTIC_T, 0, a);
        else if (want_g) IDP_LAUNCH(c, (k_hinge<false, true, false>), gH, ELASTIC_T, 0, a);
    }
    IDP_CK(c, cudaGetLastError());
    if (want_e) {
        IDP_LAUNCH(c, k_elastic_sum, 1, 256, 0, c->red.p, (int)(gM + gH), c->red.p + gM + gH);
        if (comm_on(c)) IDP_TRY(comm_allreduce_sum(c, c->red.p + gM + gH, 1));
        double E = 0;
        IDP_CK(c, cudaMemcpyAsync(&E, c->red.p + gM + gH, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        if (E_out) *E_out = E;
    }
    if (want_g && comm_on(c)) IDP_TRY(comm_allreduce_sum(c, c->elasticG.p, 3L * c->nV));
    return IDP_OK;
}

} // namespace idp
