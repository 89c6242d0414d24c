// The caller-side steps around the barrier Hessian that keep the Newton system on the device:
//   * CSR_MATRIX::Project_DBC            /root/reference/Library/Math/CSR_MATRIX.h:130-141
//   * the role of Solve_Direct            Math/DIRECT_SOLVER.h:14-88 (CHOLMOD / SimplicialLDLT) -- here a block-Jacobi
//                                         preconditioned conjugate gradient on the device CSR (iterative: tolerance in the call)
//   * Find_Surface_Primitives_And_Compute_Area  Utils/MESHIO.h:768-834 (std::map ordering contract, areas)
// FP64 + int32; everything here is bound by HBM bandwidth (SpMV: 12 bytes per stored scalar).
#include "ctx.cuh"
#include <cooperative_groups.h>
#include <cub/cub.cuh>

namespace idp {

// ------------------------------------------------------------------------------------------------------------
// Project_DBC: entries whose row or column vertex is a Dirichlet node become (row == col)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_project_dbc(const int* __restrict__ ptr, const int* __restrict__ col, double* __restrict__ val,
    const unsigned char* __restrict__ dbc, int nRowsScalar)
{
    const int lane = threadIdx.x & 31;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nRowsScalar; r += (gridDim.x * blockDim.x) >> 5) {
        const int p0 = ptr[r], p1 = ptr[r + 1];
        const bool rowFixed = dbc[r / 3] != 0;
        for (int p = p0 + lane; p < p1; p += 32) {
            const int cI = col[p];
            if (rowFixed || dbc[cI / 3]) val[p] = (cI == r) ? 1.0 : 0.0;
        }
    }
}
int project_dbc(idp_ctx* c, const unsigned char* host_mask)
{
    if (c->nnz <= 0) return IDP_OK;
    if (c->nranks > 1) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_project_dbc: the per-rank CSRs are partial sums; project after summing them (single-GPU contexts only)", __FILE__, __LINE__);
    const unsigned char* mask = c->dbc.p;
    if (host_mask) { // a caller-supplied mask (the nodes that stay fixed while the augmented-Lagrangian Dirichlet penalty is active)
        IDP_CK(c, c->projMask.reserve(c->nV));
        IDP_CK(c, cudaMemcpyAsync(c->projMask.p, host_mask, c->nV, cudaMemcpyHostToDevice, c->stream));
        mask = c->projMask.p;
    }
    IDP_LAUNCH(c, k_project_dbc, std::min(blocks_for(32L * 3 * c->nV, 256), (unsigned)c->sm_count * 32), 256, 0, c->csrPtr.p, c->csrCol.p, c->csrVal.p, mask, 3 * c->nV);
    if (host_mask) IDP_CK(c, cudaStreamSynchronize(c->stream)); // the host buffer may go away
    IDP_CK(c, cudaGetLastError());
    c->csrProjected = true;
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Block-Jacobi PCG on the scalar CSR. The matrix comes from assemble_csr, so the three scalar rows of a vertex have the
// same block columns (3 nb entries each, stored back to back): one warp takes a block row, streams the three value rows
// and the column indices coalesced and gathers x.
// Scalars (rho = r.z, p.Ap, r.r) never visit the host inside the loop: every phase leaves per-block partial sums in a
// fixed slot array and the consumers add them up in slot order themselves (deterministic, no atomics); the host reads the
// iteration count and the residual once, after the solve.
// ------------------------------------------------------------------------------------------------------------
#define IDP_PCG_BLOCKS 1184 // 148 SMs x 8
#define IDP_PCG_CHECK 10
struct PcgScal { double part[3][IDP_PCG_BLOCKS]; }; // 0: p.Ap, 1: r.z (new), 2: r.r
// every thread of every block adds the same values in the same order (strided per-thread sums, xor-butterfly inside the warp,
// the eight warp sums in warp order): identical result in all threads and blocks, no atomics
__device__ __forceinline__ void sum_partials2(const double* __restrict__ p0, const double* __restrict__ p1, int n, double& s0, double& s1)
{
    __shared__ double sred[2][8];
    double a = 0, b = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { a += p0[i]; b += p1[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = a; sred[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    a = 0; b = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sred[0][w]; b += sred[1][w]; }
    __syncthreads();
    s0 = a; s1 = b;
}
__device__ __forceinline__ double sum_partials(const double* __restrict__ p, int n = IDP_PCG_BLOCKS)
{
    double a, b;
    sum_partials2(p, p, n, a, b);
    return a;
}
__device__ __forceinline__ void block_partial(double v, double* __restrict__ slot)
{
    typedef cub::BlockReduce<double, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const double s = BR(tmp).Sum(v);
    if (threadIdx.x == 0) *slot = s;
    __syncthreads();
}
// inverse of the 3x3 diagonal blocks (block-Jacobi); a singular block falls back to the inverse of its diagonal
__global__ void __launch_bounds__(256) k_pcg_inv_diag(const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val, int nV,
    double* __restrict__ inv)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += gridDim.x * blockDim.x) {
        double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; ++i) {
            const int r = 3 * v + i;
            int lo = ptr[r], hi = ptr[r + 1];
            while (lo < hi) { // first entry with col >= 3 v (columns ascend)
                const int mid = (lo + hi) >> 1;
                if (col[mid] < 3 * v) lo = mid + 1;
                else hi = mid;
            }
            for (int p = lo; p < ptr[r + 1] && col[p] < 3 * v + 3; ++p) a[3 * i + (col[p] - 3 * v)] = val[p];
        }
        const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
        const double det = a[0] * c0 + a[1] * c1 + a[2] * c2;
        const double scale = fabs(a[0] * a[4] * a[8]);
        double* o = inv + 9 * (long)v;
        if (fabs(det) > 1e-14 * scale && scale > 0) {
            const double id = 1.0 / det;
            o[0] = c0 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
            o[3] = c1 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
            o[6] = c2 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
        }
        else {
            for (int k = 0; k < 9; ++k) o[k] = 0;
            for (int k = 0; k < 3; ++k) o[4 * k] = a[4 * k] != 0 ? 1.0 / a[4 * k] : 1.0;
        }
    }
}
// r = b (x0 = 0), z = M^-1 r, p = 0; partial sums of r.z and r.r
__global__ void __launch_bounds__(256) k_pcg_init(const double* __restrict__ b, const double* __restrict__ inv, int nV, double* __restrict__ x,
    double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, PcgScal* __restrict__ sc)
{
    double rz = 0, rr = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += gridDim.x * blockDim.x) {
        const double r0 = b[3 * v], r1 = b[3 * v + 1], r2 = b[3 * v + 2];
        const double* m = inv + 9 * (long)v;
        const double z0 = m[0] * r0 + m[1] * r1 + m[2] * r2, z1 = m[3] * r0 + m[4] * r1 + m[5] * r2, z2 = m[6] * r0 + m[7] * r1 + m[8] * r2;
        x[3 * v] = 0; x[3 * v + 1] = 0; x[3 * v + 2] = 0;
        r[3 * v] = r0; r[3 * v + 1] = r1; r[3 * v + 2] = r2;
        z[3 * v] = z0; z[3 * v + 1] = z1; z[3 * v + 2] = z2;
        p[3 * v] = 0; p[3 * v + 1] = 0; p[3 * v + 2] = 0; // the first direction is formed as z + 0 * p by the iteration kernel
        rz += r0 * z0 + r1 * z1 + r2 * z2;
        rr += r0 * r0 + r1 * r1 + r2 * r2;
    }
    block_partial(rz, &sc->part[1][blockIdx.x]);
    block_partial(rr, &sc->part[2][blockIdx.x]);
}
__global__ void __launch_bounds__(256) k_pcg_read_rr(const PcgScal* __restrict__ sc, double* __restrict__ out)
{
    const double rr = sum_partials(sc->part[2]);
    if (threadIdx.x == 0) *out = rr;
}

// The whole iteration in ONE cooperative launch (one resident grid, three grid-wide barriers per iteration, convergence tested on
// the device after every iteration). The first version launched three kernels per iteration and read the residual every ten
// iterations: for the small systems of the paper examples (10^4 unknowns, ~250 iterations per Newton step) that was pure launch
// latency -- 24 us per iteration, 75 % of the whole normal-flow run. Phases: (1) Ap = A p, one warp per block row streaming the
// three value rows and the column indices coalesced, partial p.Ap; (2) alpha, x += alpha p, r -= alpha Ap, z = M^-1 r, partials of
// r.z and r.r; the next direction p = z + beta p is formed inside phase (1) of the following iteration (two barriers per iteration, not three).
// Partials per block, summed in slot order by every block (deterministic).
struct PcgPersistentArgs {
    const int* ptr; const int* col; const double* val; const double* inv; int nV;
    double* x; double* r; double* z; double* p0; double* p1; double* Ap; PcgScal* sc;
    double target; int maxIter; int* itersOut; double* rrOut;
};
__global__ void __launch_bounds__(256) k_pcg_persistent(PcgPersistentArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int nb = gridDim.x, lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    int cur = 0;
    double rho, rr;
    sum_partials2(a.sc[0].part[1], a.sc[0].part[2], nb, rho, rr);
    double beta = 0.0;           // p = z + beta p_old; the first direction is z itself (p_old = 0 from k_pcg_init)
    double* po = a.p0; double* pn = a.p1;
    int it = 0;
    while (it < a.maxIter && rr > a.target) {
        { // new direction and A p in one pass: the gathers form p[col] = z[col] + beta p_old[col] themselves (same fma as the owner's
          // write), so no grid-wide barrier is needed between the direction update and the product
            double acc = 0;
            for (int v = gtid >> 5; v < a.nV; v += gsize >> 5) {
                // the three scalar rows of a block row have the same columns (assemble_csr stores them back to back, 3 nb entries
                // each): one column load and one gathered direction value serve all three rows
                double s[3] = {0, 0, 0};
                const int q0 = a.ptr[3 * v], len = a.ptr[3 * v + 1] - q0;
                const double* v0 = a.val + q0; const double* v1 = v0 + len; const double* v2 = v1 + len;
                int k = lane;
                for (; k + 32 < len; k += 64) { // two entries per lane and trip: eight independent loads in flight before the first use
                    const int cA = a.col[q0 + k], cB = a.col[q0 + k + 32];
                    const double a0 = v0[k], a1 = v1[k], a2 = v2[k], b0 = v0[k + 32], b1 = v1[k + 32], b2 = v2[k + 32];
                    const double pA = fma(beta, po[cA], a.z[cA]), pB = fma(beta, po[cB], a.z[cB]);
                    s[0] += a0 * pA; s[1] += a1 * pA; s[2] += a2 * pA;
                    s[0] += b0 * pB; s[1] += b1 * pB; s[2] += b2 * pB;
                }
                if (k < len) {
                    const int cI = a.col[q0 + k];
                    const double pv = fma(beta, po[cI], a.z[cI]);
                    s[0] += v0[k] * pv; s[1] += v1[k] * pv; s[2] += v2[k] * pv;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o); s[1] += __shfl_xor_sync(0xffffffffu, s[1], o); s[2] += __shfl_xor_sync(0xffffffffu, s[2], o);
                }
                if (lane == 0) {
                    double pv[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { pv[k] = fma(beta, po[3 * v + k], a.z[3 * v + k]); pn[3 * v + k] = pv[k]; a.Ap[3 * v + k] = s[k]; }
                    acc += pv[0] * s[0] + pv[1] * s[1] + pv[2] * s[2];
                }
            }
            block_partial(acc, &a.sc[cur].part[0][blockIdx.x]);
        }
        grid.sync();
        { // x += alpha p; r -= alpha Ap; z = M^-1 r; partials of the new r.z and r.r
            const double pAp = sum_partials(a.sc[cur].part[0], nb);
            const double alpha = pAp != 0 ? rho / pAp : 0.0;
            double rz = 0, r2 = 0;
            for (int v = gtid; v < a.nV; v += gsize) {
                double rv[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    a.x[3 * v + k] += alpha * pn[3 * v + k];
                    rv[k] = a.r[3 * v + k] - alpha * a.Ap[3 * v + k];
                    a.r[3 * v + k] = rv[k];
                }
                const double* m = a.inv + 9 * (long)v;
                const double z0 = m[0] * rv[0] + m[1] * rv[1] + m[2] * rv[2], z1 = m[3] * rv[0] + m[4] * rv[1] + m[5] * rv[2], z2 = m[6] * rv[0] + m[7] * rv[1] + m[8] * rv[2];
                a.z[3 * v] = z0; a.z[3 * v + 1] = z1; a.z[3 * v + 2] = z2;
                rz += rv[0] * z0 + rv[1] * z1 + rv[2] * z2;
                r2 += rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
            }
            block_partial(rz, &a.sc[cur ^ 1].part[1][blockIdx.x]);
            block_partial(r2, &a.sc[cur ^ 1].part[2][blockIdx.x]);
        }
        grid.sync();
        double rhoNew;
        sum_partials2(a.sc[cur ^ 1].part[1], a.sc[cur ^ 1].part[2], nb, rhoNew, rr);
        beta = rho != 0 ? rhoNew / rho : 0.0;
        rho = rhoNew;
        double* t = po; po = pn; pn = t;
        cur ^= 1;
        ++it;
        if (!(rr == rr)) break; // breakdown: reported by the host
    }
    if (gtid == 0) { *a.itersOut = it; *a.rrOut = rr; }
}

int solve_pcg(idp_ctx* c, const double* rhs, double* sol, double rel_tol, int max_iter, int* iters, double* rel_res)
{
    if (c->nranks > 1) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_solve_pcg: single-GPU contexts only (a rank's CSR is a partial sum)", __FILE__, __LINE__);
    if (c->nnz <= 0) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_solve_pcg: no system matrix (call idp_barrier_hessian / idp_barrier_all first)", __FILE__, __LINE__);
    const int nV = c->nV;
    const size_t n = 3 * (size_t)nV;
    IDP_CK(c, c->pcgX.reserve(n)); IDP_CK(c, c->pcgR.reserve(n)); IDP_CK(c, c->pcgZ.reserve(n)); IDP_CK(c, c->pcgP.reserve(n)); IDP_CK(c, c->pcgP2.reserve(n)); IDP_CK(c, c->pcgAp.reserve(n));
    IDP_CK(c, c->pcgInvDiag.reserve(9 * (size_t)nV));
    IDP_CK(c, c->pcgScal.reserve(2 * sizeof(PcgScal) / sizeof(double) + 8)); // + the (rr, iterations) read-back slots
    PcgScal* sc[2] = {(PcgScal*)c->pcgScal.p, (PcgScal*)c->pcgScal.p + 1};
    double* dRR = c->pcgScal.p + 2 * sizeof(PcgScal) / sizeof(double);
    // rhs: host -> pcgAp (scratch) -> r
    IDP_CK(c, cudaMemcpyAsync(c->pcgAp.p, rhs, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemsetAsync(c->pcgScal.p, 0, 2 * sizeof(PcgScal), c->stream));
    IDP_LAUNCH(c, k_pcg_inv_diag, blocks_for(nV, 256), 256, 0, c->csrPtr.p, c->csrCol.p, c->csrVal.p, nV, c->pcgInvDiag.p);
    // one resident grid for the whole solve: as many blocks as the problem can use, at most what fits on the device at once
    int perSm = 0;
    IDP_CK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg_persistent, 256, 0));
    // at least ~4 block rows per warp: a grid-wide barrier costs more the more blocks take part, and small systems are all latency
    const int grid = std::max(1, std::min(std::min(perSm * c->sm_count, IDP_PCG_BLOCKS), (int)blocks_for(8L * nV, 256)));
    IDP_LAUNCH(c, k_pcg_init, grid, 256, 0, c->pcgAp.p, c->pcgInvDiag.p, nV, c->pcgX.p, c->pcgR.p, c->pcgZ.p, c->pcgP.p, sc[0]);
    IDP_LAUNCH(c, k_pcg_read_rr, 1, 256, 0, sc[0], dRR);
    double rr0 = 0;
    IDP_CK(c, cudaMemcpyAsync(&rr0, dRR, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    int it = 0;
    double rr = rr0;
    if (rr0 > 0 && max_iter > 0) {
        PcgPersistentArgs pa;
        pa.ptr = c->csrPtr.p; pa.col = c->csrCol.p; pa.val = c->csrVal.p; pa.inv = c->pcgInvDiag.p; pa.nV = nV;
        pa.x = c->pcgX.p; pa.r = c->pcgR.p; pa.z = c->pcgZ.p; pa.p0 = c->pcgP.p; pa.p1 = c->pcgP2.p; pa.Ap = c->pcgAp.p; pa.sc = sc[0];
        pa.target = rel_tol * rel_tol * rr0; pa.maxIter = max_iter;
        pa.itersOut = (int*)(dRR + 1); pa.rrOut = dRR;
        void* kargs[] = {&pa};
        IDP_CK(c, cudaLaunchCooperativeKernel((const void*)k_pcg_persistent, dim3(grid), dim3(256), kargs, 0, c->stream));
        ++c->launches;
        double out[2] = {0, 0};
        IDP_CK(c, cudaMemcpyAsync(out, dRR, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        rr = out[0];
        std::memcpy(&it, &out[1], sizeof(int));
        if (!(rr == rr)) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_solve_pcg: the iteration broke down (matrix not positive definite?)", __FILE__, __LINE__);
    }
    IDP_CK(c, cudaGetLastError());
    if (sol) {
        IDP_CK(c, cudaMemcpyAsync(sol, c->pcgX.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    if (iters) *iters = it;
    if (rel_res) *rel_res = rr0 > 0 ? std::sqrt(rr / rr0) : 0.0;
    return IDP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Surface primitives of a triangle mesh in the reference's ordering contract (Utils/MESHIO.h:768-834):
//   boundaryTri  = the elements in element order;
//   boundaryEdge = every undirected edge once, with the orientation of the first directed edge (element order, then
//                  (v0,v1), (v1,v2), (v2,v0)) that introduced it -- the reference looks for the reversed pair in a
//                  std::map and otherwise inserts / overwrites the pair itself, which leaves exactly that orientation --
//                  listed in the lexicographic (v0, v1) order of the map;
//   boundaryNode = the vertices whose summed incident triangle area is non-zero, ascending.
// Areas: BTArea = area / 2, BEArea = (sum of area / 3 over incident triangles) / 2, BNArea = sum of area / 3.
// Two radix sorts of 64-bit keys (3 nF directed edges by (min, max, sequence), then the unique ones by (v0, v1)).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_surface_edges(const int* __restrict__ tri, int stride, int nF, const double* __restrict__ x, int xstride, int nV,
    unsigned long long* __restrict__ key, int* __restrict__ seq, double* __restrict__ triArea, double* __restrict__ nodeArea)
{
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
        const int v[3] = {tri[(long)stride * f], tri[(long)stride * f + 1], tri[(long)stride * f + 2]};
        double area = 0;
        if (x) {
            const double* a = x + (long)xstride * v[0]; const double* b = x + (long)xstride * v[1]; const double* cc = x + (long)xstride * v[2];
            const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {cc[0] - a[0], cc[1] - a[1], cc[2] - a[2]};
            // no FMA contraction: a degenerate triangle must give exactly zero (the boundary-node rule tests the summed area != 0),
            // and the sum follows VECTOR::length (Math/VECTOR.h:136-151): (p0 + p1) + p2
            const double n0 = __dsub_rn(__dmul_rn(e1[1], e2[2]), __dmul_rn(e1[2], e2[1])), n1 = __dsub_rn(__dmul_rn(e1[2], e2[0]), __dmul_rn(e1[0], e2[2])),
                         n2 = __dsub_rn(__dmul_rn(e1[0], e2[1]), __dmul_rn(e1[1], e2[0]));
            area = 0.5 * sqrt(__dadd_rn(__dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1)), __dmul_rn(n2, n2)));
        }
        else area = 1.0; // topology only: every referenced vertex is a boundary node
        triArea[f] = area;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int a = v[k], b = v[(k + 1) % 3];
            const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
            key[3L * f + k] = ((unsigned long long)lo << 32) | hi;
            seq[3L * f + k] = 3 * f + k;
            atomicAdd(&nodeArea[v[k]], area / 3);
        }
    }
}
// after the stable sort by (min, max): heads of the runs are the first occurrences; emit the stored orientation as key (v0, v1)
__global__ void __launch_bounds__(256) k_surface_heads(const unsigned long long* __restrict__ key, const int* __restrict__ seq, long n, const int* __restrict__ tri, int stride,
    unsigned char* __restrict__ head)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
    (void)seq; (void)tri; (void)stride;
}
__global__ void __launch_bounds__(256) k_surface_unique(const unsigned long long* __restrict__ key, const int* __restrict__ seq, const int* __restrict__ headPos, long n,
    const int* __restrict__ tri, int stride, const double* __restrict__ triArea, unsigned long long* __restrict__ okey, double* __restrict__ oarea)
{
    // one thread per run head: orientation of the first occurrence, area summed over the run in sequence order
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        if (!(i == 0 || key[i] != key[i - 1])) continue;
        const int u = headPos[i];
        const int s = seq[i], f = s / 3, k = s - 3 * f;
        const int a = tri[(long)stride * f + k], b = tri[(long)stride * f + (k + 1) % 3];
        okey[u] = ((unsigned long long)(unsigned)a << 32) | (unsigned)b;
        // the reference looks for the REVERSED pair: found -> its area grows; not found -> map[(a, b)] = area / 3, which
        // OVERWRITES when a later triangle repeats the stored orientation (inconsistently oriented / non-manifold input)
        double sum = 0;
        for (long j = i; j < n && key[j] == key[i]; ++j) {
            const int sj = seq[j], fj = sj / 3, kj = sj - 3 * fj;
            const double w = triArea[fj] / 3;
            if (tri[(long)stride * fj + kj] == a) sum = w;
            else sum += w;
        }
        oarea[u] = sum / 2;
    }
}
__global__ void __launch_bounds__(256) k_surface_unpack(const unsigned long long* __restrict__ key, int n, int2* __restrict__ edge)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) edge[i] = make_int2((int)(key[i] >> 32), (int)(key[i] & 0xffffffffu));
}
__global__ void __launch_bounds__(256) k_surface_node_flags(const double* __restrict__ nodeArea, int nV, int* __restrict__ flag)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += gridDim.x * blockDim.x) flag[v] = nodeArea[v] != 0.0 ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_surface_nodes(const int* __restrict__ flag, const int* __restrict__ pos, const double* __restrict__ nodeArea, int nV,
    int* __restrict__ bnode, double* __restrict__ bnArea)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += gridDim.x * blockDim.x)
        if (flag[v]) { bnode[pos[v]] = v; bnArea[pos[v]] = nodeArea[v]; }
}
__global__ void __launch_bounds__(256) k_surface_tris(const int* __restrict__ tri, int stride, int nF, const double* __restrict__ triArea, int4* __restrict__ btri, double* __restrict__ btArea)
{
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
        btri[f] = make_int4(tri[(long)stride * f], tri[(long)stride * f + 1], tri[(long)stride * f + 2], 0);
        btArea[f] = triArea[f] / 2;
    }
}

// sets the context's mesh (like idp_set_mesh) from the element list; the areas stay on the device for idp_get_surface_primitives
int extract_surface(idp_ctx* c, int nV, int nF, const int* tri, int stride, const double* x, int xstride, const unsigned char* dbc)
{
    const long n3 = 3L * nF;
    IDP_CK(c, c->surfTri.reserve((size_t)std::max(1L, (long)stride * nF)));
    IDP_CK(c, cudaMemcpyAsync(c->surfTri.p, tri, (size_t)stride * nF * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const double* dx = nullptr;
    if (x) {
        IDP_CK(c, c->stage.reserve((size_t)xstride * nV));
        IDP_CK(c, cudaMemcpyAsync(c->stage.p, x, (size_t)xstride * nV * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        dx = c->stage.p;
    }
    IDP_CK(c, c->keyA.reserve(std::max(n3, 1L))); IDP_CK(c, c->keyTmp.reserve(std::max(n3, 1L)));
    IDP_CK(c, c->rowIota.reserve(std::max(n3, 1L))); IDP_CK(c, c->rowPerm.reserve(std::max(n3, 1L)));
    IDP_CK(c, c->surfTriArea.reserve(std::max(nF, 1))); IDP_CK(c, c->surfNodeArea.reserve(nV));
    IDP_CK(c, cudaMemsetAsync(c->surfNodeArea.p, 0, nV * sizeof(double), c->stream));
    if (nF > 0) IDP_LAUNCH(c, k_surface_edges, std::min(blocks_for(nF, 256), (unsigned)c->sm_count * 16), 256, 0, c->surfTri.p, stride, nF, dx, xstride, nV, c->keyA.p, c->rowIota.p, c->surfTriArea.p, c->surfNodeArea.p);
    // stable sort by (min, max): equal keys keep the sequence (= visiting) order
    size_t bytes = 0;
    int nBE = 0;
    if (n3 > 0) {
        IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->keyA.p, c->keyTmp.p, c->rowIota.p, c->rowPerm.p, (int)n3, 0, 64, c->stream));
        IDP_CK(c, c->cubTemp.reserve(bytes));
        IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, c->keyA.p, c->keyTmp.p, c->rowIota.p, c->rowPerm.p, (int)n3, 0, 64, c->stream));
        ++c->lib_launches;
        // positions of the run heads
        IDP_CK(c, c->rowKind.reserve(n3)); IDP_CK(c, c->vtxCnt.reserve(std::max<long>(n3 + 1, nV + 1))); IDP_CK(c, c->vtxOff.reserve(std::max<long>(n3 + 1, nV + 1)));
        IDP_LAUNCH(c, k_surface_heads, std::min(blocks_for(n3, 256), (unsigned)c->sm_count * 16), 256, 0, c->keyTmp.p, c->rowPerm.p, n3, c->surfTri.p, stride, c->rowKind.p);
        IDP_CK(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, c->rowKind.p, c->vtxOff.p, (int)n3, c->stream));
        IDP_CK(c, c->cubTemp.reserve(bytes));
        IDP_CK(c, cub::DeviceScan::ExclusiveSum(c->cubTemp.p, bytes, c->rowKind.p, c->vtxOff.p, (int)n3, c->stream));
        ++c->lib_launches;
        int lastPos = 0; unsigned char lastHead = 0;
        IDP_CK(c, cudaMemcpyAsync(&lastPos, c->vtxOff.p + (n3 - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaMemcpyAsync(&lastHead, c->rowKind.p + (n3 - 1), 1, cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        nBE = lastPos + lastHead;
        IDP_CK(c, c->keyB.reserve(nBE)); IDP_CK(c, c->keyD.reserve(nBE));
        IDP_CK(c, c->surfEdgeArea.reserve(nBE)); IDP_CK(c, c->surfEdgeArea2.reserve(nBE));
        IDP_LAUNCH(c, k_surface_unique, std::min(blocks_for(n3, 256), (unsigned)c->sm_count * 16), 256, 0, c->keyTmp.p, c->rowPerm.p, c->vtxOff.p, n3, c->surfTri.p, stride,
            c->surfTriArea.p, c->keyB.p, c->surfEdgeArea2.p);
        // the map's order: lexicographic (v0, v1) of the stored pairs (keys are unique: a plain sort)
        IDP_CK(c, cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->keyB.p, c->keyD.p, c->surfEdgeArea2.p, c->surfEdgeArea.p, nBE, 0, 64, c->stream));
        IDP_CK(c, c->cubTemp.reserve(bytes));
        IDP_CK(c, cub::DeviceRadixSort::SortPairs(c->cubTemp.p, bytes, c->keyB.p, c->keyD.p, c->surfEdgeArea2.p, c->surfEdgeArea.p, nBE, 0, 64, c->stream));
        ++c->lib_launches;
    }
    // nodes
    IDP_CK(c, c->vtxCnt.reserve(nV + 1)); IDP_CK(c, c->vtxOff.reserve(nV + 1));
    IDP_LAUNCH(c, k_surface_node_flags, blocks_for(nV, 256), 256, 0, c->surfNodeArea.p, nV, c->vtxCnt.p);
    IDP_CK(c, cudaMemsetAsync(c->vtxCnt.p + nV, 0, sizeof(int), c->stream));
    IDP_TRY(cub_scan_exclusive(c, c->vtxCnt.p, c->vtxOff.p, (long)nV + 1));
    int nBN = 0;
    IDP_CK(c, cudaMemcpyAsync(&nBN, c->vtxOff.p + nV, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    // commit as the context's mesh
    c->nV = nV; c->nBN = nBN; c->nBE = nBE; c->nBT = nF;
    c->have_x = c->have_x0 = c->have_dir = false;
    c->meanEdgeVersion = -1;
    c->nRows = 0; c->nCandPT = c->nCandEE = c->nCcdPT = c->nCcdEE = 0;
    c->permValid = false;
    IDP_CK(c, c->bnode.reserve(std::max(nBN, 1))); IDP_CK(c, c->bedge.reserve(std::max(nBE, 1))); IDP_CK(c, c->btri.reserve(std::max(nF, 1)));
    IDP_CK(c, c->surfNodeAreaC.reserve(std::max(nBN, 1))); IDP_CK(c, c->surfTriAreaH.reserve(std::max(nF, 1)));
    IDP_CK(c, c->dbc.reserve(nV));
    IDP_LAUNCH(c, k_surface_nodes, blocks_for(nV, 256), 256, 0, c->vtxCnt.p, c->vtxOff.p, c->surfNodeArea.p, nV, c->bnode.p, c->surfNodeAreaC.p);
    if (nBE > 0) IDP_LAUNCH(c, k_surface_unpack, blocks_for(nBE, 256), 256, 0, c->keyD.p, nBE, c->bedge.p);
    if (nF > 0) IDP_LAUNCH(c, k_surface_tris, blocks_for(nF, 256), 256, 0, c->surfTri.p, stride, nF, c->surfTriArea.p, c->btri.p, c->surfTriAreaH.p);
    if (dbc) IDP_CK(c, cudaMemcpyAsync(c->dbc.p, dbc, nV, cudaMemcpyHostToDevice, c->stream));
    else IDP_CK(c, cudaMemsetAsync(c->dbc.p, 0, nV, c->stream));
    IDP_CK(c, cudaGetLastError());
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->surfValid = true;
    return IDP_OK;
}

} // namespace idp
