// extern "C" launcher layer of libidp_contact.so (include/idp_contact.h). Thin: argument checks, marshalling of
// host buffers, and calls into the launchers of exact_kernels.cu / barrier_kernels.cu / comm.cu.
#include "ctx.cuh"
#include <chrono>
#include <string.h>
#include <thread>

using namespace idp;

// run f(begin, end) over [0, n) on a few host threads (marshalling loops over tens of millions of rows)
template <class F>
static void host_parallel(long n, F f)
{
    const int T = (int)std::max(1L, std::min<long>(8, n / 65536));
    if (T == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([=]() { f(n * t / T, n * (t + 1) / T); });
    for (auto& x : th) x.join();
}
// pinned host staging owned by the context (grown on demand)
static int host_stage(idp_ctx* c, size_t bytes, void** out)
{
    if (bytes > c->h_stage_bytes) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr;
        c->h_stage_bytes = 0;
        IDP_CK(c, cudaMallocHost(&c->h_stage, bytes + bytes / 8));
        c->h_stage_bytes = bytes + bytes / 8;
    }
    *out = c->h_stage;
    return IDP_OK;
}

namespace idp {
void comm_destroy(idp_ctx* c);
} // namespace idp

extern "C" {

int idp_create(int device, idp_ctx** out)
{
    if (!out) return IDP_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return IDP_ERR_CUDA; // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return IDP_ERR_CUDA;
    idp_ctx* c = new idp_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return IDP_ERR_CUDA; }
    c->own_stream = true;
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    for (int i = 0; i < 2 * IDP_EVENT_POOL; ++i) cudaEventCreate(&c->evPool[i]);
    if (c->counters.reserve(CNT_COUNT) != cudaSuccess || c->histScratch.reserve(256) != cudaSuccess) { delete c; return IDP_ERR_CUDA; }
    cudaMemset(c->counters.p, 0, CNT_COUNT * sizeof(long long));
    cudaMallocHost((void**)&c->h_counters, CNT_COUNT * sizeof(long long));
    cudaMallocHost((void**)&c->h_red, 64 * sizeof(double));
    memset(c->h_counters, 0, CNT_COUNT * sizeof(long long));
    memset(&c->times, 0, sizeof(c->times));
    *out = c;
    return IDP_OK;
}

void idp_destroy(idp_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    comm_destroy(c);
    idp_transfers_end(c); // joins the host workers of pending transfers
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_red) cudaFreeHost(c->h_red);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->evCopyFork) cudaEventDestroy(c->evCopyFork);
    if (c->evRows) cudaEventDestroy(c->evRows);
    if (c->evBlk) cudaEventDestroy(c->evBlk);
    if (c->evVal) cudaEventDestroy(c->evVal);
    if (c->h_blk) cudaFreeHost(c->h_blk);
    if (c->commStream) cudaStreamDestroy(c->commStream);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 2 * IDP_EVENT_POOL; ++i) if (c->evPool[i]) cudaEventDestroy(c->evPool[i]);
    cudaStream_t s = c->own_stream ? c->stream : nullptr;
    delete c; // frees the device buffers
    if (s) cudaStreamDestroy(s);
}

const char* idp_last_error(idp_ctx* c) { return c ? c->err.c_str() : "null context"; }

int idp_set_stream(idp_ctx* c, void* stream)
{
    if (!c) return IDP_ERR_INVALID;
    cudaStreamSynchronize(c->stream);
    if (stream) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    }
    else if (!c->own_stream) {
        IDP_CK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return IDP_OK;
}

int idp_set_mesh(idp_ctx* c, int nV, int nBN, const int* bnode, int nBE, const int* bedge2, int nBT, const int* btri3,
    const uint8_t* dbc)
{
    if (!c || nV <= 0 || nBN < 0 || nBE < 0 || nBT < 0 || (nBN && !bnode) || (nBE && !bedge2) || (nBT && !btri3))
        return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_set_mesh: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->nV = nV; c->nBN = nBN; c->nBE = nBE; c->nBT = nBT;
    c->have_x = c->have_x0 = c->have_dir = false;
    c->meanEdgeVersion = -1; // new boundary edges
    c->surfValid = false; c->nFlowElem = 0; c->haveMass = false; c->nMem = 0; c->nHinge = 0; c->nFric = 0; c->nFricActive = 0; c->have_xn = false;
    c->nRows = 0; c->nCandPT = c->nCandEE = c->nCcdPT = c->nCcdEE = 0;
    c->permValid = false;
    IDP_CK(c, c->bnode.reserve(std::max(nBN, 1)));
    IDP_CK(c, c->bedge.reserve(std::max(nBE, 1)));
    IDP_CK(c, c->btri.reserve(std::max(nBT, 1)));
    IDP_CK(c, c->dbc.reserve(nV));
    if (nBN) IDP_CK(c, cudaMemcpyAsync(c->bnode.p, bnode, nBN * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (nBE) IDP_CK(c, cudaMemcpyAsync(c->bedge.p, bedge2, nBE * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    if (nBT) {
        std::vector<int4> t4(nBT);
        for (int i = 0; i < nBT; ++i) t4[i] = make_int4(btri3[3 * i], btri3[3 * i + 1], btri3[3 * i + 2], 0);
        IDP_CK(c, cudaMemcpyAsync(c->btri.p, t4.data(), nBT * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    if (dbc) IDP_CK(c, cudaMemcpyAsync(c->dbc.p, dbc, nV, cudaMemcpyHostToDevice, c->stream));
    else IDP_CK(c, cudaMemsetAsync(c->dbc.p, 0, nV, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    return IDP_OK;
}

int idp_declare_unsupported(idp_ctx* c, int n_rod, int n_particle, int n_nn)
{
    if (!c) return IDP_ERR_INVALID;
    if (n_rod || n_particle || n_nn)
        return fail(c, IDP_ERR_UNSUPPORTED_PRIMITIVE, "%s (%s:%d)", "rod / particle / NNExclusion inputs are outside the ported path", __FILE__, __LINE__);
    return IDP_OK;
}

int idp_set_positions(idp_ctx* c, const double* x, int stride)
{
    if (!c || !x) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    return upload_positions(c, x, stride, 0);
}
int idp_set_rest_positions(idp_ctx* c, const double* x0, int stride)
{
    if (!c || !x0) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    return upload_positions(c, x0, stride, 1);
}

int idp_constraint_set(idp_ctx* c, double dhat2, double thickness, int* n_rows)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(build_constraint_set(c, dhat2, thickness));
    if (n_rows) *n_rows = (int)(c->rowsLocal ? c->nRowsGlobal : c->nRows);
    return IDP_OK;
}

// fills info2 = (weight, dHat2) for n rows whose weights live at dW (or are all one)
static int fill_info(idp_ctx* c, long n, const double* dW, bool allOne, double* info2)
{
    const double dh2 = c->cs_dhat2;
    if (allOne) { // set by idp_constraint_set (OIPC: weight 1, IPC.h:656-660): nothing to copy back
        host_parallel(n, [=](long b, long e) { for (long i = b; i < e; ++i) { info2[2 * i] = 1.0; info2[2 * i + 1] = dh2; } });
        return IDP_OK;
    }
    double* w = nullptr;
    IDP_TRY(host_stage(c, (size_t)n * sizeof(double), (void**)&w));
    IDP_CK(c, cudaMemcpyAsync(w, dW, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    host_parallel(n, [=](long b, long e) { for (long i = b; i < e; ++i) { info2[2 * i] = w[i]; info2[2 * i + 1] = dh2; } });
    return IDP_OK;
}

int idp_get_constraints(idp_ctx* c, int* rows4, double* info2)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    // sharded: the rows THIS rank holds (idp_last_count(ctx, 9) of them) -- not a collective; the global list is
    // idp_gather_constraints
    if (rows4 && c->nRows) {
        IDP_CK(c, cudaMemcpyAsync(rows4, c->rows.p, c->nRows * sizeof(Row4), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    if (info2 && c->nRows) IDP_TRY(fill_info(c, c->nRows, c->weights.p, c->weights_all_one, info2));
    return IDP_OK;
}

int idp_gather_constraints(idp_ctx* c, int* rows4, double* info2, double* dist2)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (!(c->rowsLocal && comm_on(c))) { // unsharded (or replicated rows): the local list is the global one
        IDP_TRY(idp_get_constraints(c, rows4, info2));
        if (dist2 && c->nRows) {
            if (!c->dist2Valid) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_gather_constraints: dist2 wanted but idp_min_dist2 has not run on these rows", __FILE__, __LINE__);
            IDP_CK(c, cudaMemcpyAsync(dist2, c->rowDist2.p, c->nRows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            IDP_CK(c, cudaStreamSynchronize(c->stream));
        }
        return IDP_OK;
    }
    // sharded LOCAL-ROWS mode: collective -- every rank must call it with the same non-NULL pattern; the global list (order
    // of the unsharded path) is gathered on the device and copied out. Weights are all one here (idp_constraint_set).
    const long n = c->nRowsGlobal;
    if (rows4 && n) {
        IDP_CK(c, c->rowsGlobal.reserve(n));
        IDP_TRY(comm_gather_groups(c, c->rows.p, sizeof(Row4), c->rowsGlobal.p));
        IDP_CK(c, cudaMemcpyAsync(rows4, c->rowsGlobal.p, n * sizeof(Row4), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    if (info2 && n) IDP_TRY(fill_info(c, n, nullptr, true, info2));
    if (dist2 && n) {
        int st = c->dist2Valid ? IDP_OK : IDP_ERR_INVALID;
        st = comm_agree_status(c, st);
        if (st != IDP_OK) return fail(c, st, "%s (%s:%d)", "idp_gather_constraints: dist2 wanted but idp_min_dist2 has not run on these rows", __FILE__, __LINE__);
        IDP_CK(c, c->dist2Global.reserve(n));
        IDP_TRY(comm_gather_groups(c, c->rowDist2.p, sizeof(double), c->dist2Global.p));
        IDP_CK(c, cudaMemcpyAsync(dist2, c->dist2Global.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    return IDP_OK;
}

int idp_set_constraints(idp_ctx* c, int n, const int* rows4, const double* info2)
{
    if (!c || n < 0 || (n && !rows4)) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->nRows = n;
    c->rowsLocal = false;
    c->permValid = false;
    c->dist2Valid = false;
    c->weights_all_one = false;
    IDP_CK(c, c->rows.reserve(std::max(n, 1)));
    IDP_CK(c, c->weights.reserve(std::max(n, 1)));
    if (n) {
        IDP_CK(c, cudaMemcpyAsync(c->rows.p, rows4, (size_t)n * sizeof(Row4), cudaMemcpyHostToDevice, c->stream));
        std::vector<double> w(n, 1.0);
        if (info2) {
            for (int i = 0; i < n; ++i) w[i] = info2[2 * i];
            c->cs_dhat2 = info2[1];
        }
        IDP_CK(c, cudaMemcpyAsync(c->weights.p, w.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    return IDP_OK;
}

int idp_get_candidates(idp_ctx* c, int which, long* n_pairs, int* pairs2)
{
    if (!c || which < 0 || which > 3 || !n_pairs) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    *n_pairs = which == 0 ? c->nCandPT : (which == 1 ? c->nCandEE : (which == 2 ? c->nCcdPT : c->nCcdEE));
    if (!pairs2) return IDP_OK;
    return sorted_candidates(c, which, (int2*)pairs2);
}

static int finish_energy(idp_ctx* c, double E, double* E_inout)
{
    if (comm_on(c)) {
        IDP_CK(c, cudaMemcpyAsync(c->red.p, &E, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        IDP_TRY(comm_allreduce_sum(c, c->red.p, 1));
        IDP_CK(c, cudaMemcpyAsync(&E, c->red.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    if (E_inout) *E_inout += E;
    return IDP_OK;
}
static int finish_gradient(idp_ctx* c, double* g_accum, int stride)
{
    if (comm_on(c)) IDP_TRY(comm_allreduce_sum(c, c->gbuf.p, 3L * c->nV));
    if (!g_accum) return IDP_OK;
    return idp_get_gradient(c, g_accum, stride);
}

int idp_barrier_energy(idp_ctx* c, double dhat2, double kappa, double thickness, double* E_inout)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    double E = 0;
    IDP_TRY(comm_agree_status(c, barrier_eval(c, dhat2, kappa, thickness, 1, 0, 0, 0, &E)));
    return finish_energy(c, E, E_inout);
}
int idp_barrier_gradient(idp_ctx* c, double dhat2, double kappa, double thickness, double* g_accum, int stride)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(comm_agree_status(c, barrier_eval(c, dhat2, kappa, thickness, 0, 1, 0, 0, nullptr)));
    return finish_gradient(c, g_accum, stride);
}
int idp_barrier_hessian(idp_ctx* c, double dhat2, double kappa, double thickness, int project_spd, long* nnz)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(comm_agree_status(c, barrier_eval(c, dhat2, kappa, thickness, 0, 0, 1, project_spd, nullptr)));
    IDP_TRY(assemble_csr(c));
    if (nnz) *nnz = c->nnz;
    return IDP_OK;
}
int idp_barrier_all(idp_ctx* c, double dhat2, double kappa, double thickness, int project_spd, double* E_inout, long* nnz)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    double E = 0;
    if (comm_on(c) && c->nccl_comm) {
        // Sharded over NCCL: the three reductions that follow the row kernel (status agreement, energy, 3 nV gradient) run on a
        // second stream while this rank assembles its partial CSR -- they used to be serialised after it.
        const int st = barrier_eval(c, dhat2, kappa, thickness, 1, 1, 1, project_spd, &E);
        if (!c->commStream) {
            IDP_CK(c, cudaStreamCreateWithFlags(&c->commStream, cudaStreamNonBlocking));
            IDP_CK(c, cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
            IDP_CK(c, cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
        }
        double* cell = (double*)(c->counters.p + CNT_STATUS); // [status (max), E (sum)]
        double pack[2] = {(double)st, st == IDP_OK ? E : 0.0};
        IDP_CK(c, cudaMemcpyAsync(cell, pack, sizeof(pack), cudaMemcpyHostToDevice, c->stream));
        IDP_CK(c, cudaEventRecord(c->evFork, c->stream));
        IDP_CK(c, cudaStreamWaitEvent(c->commStream, c->evFork, 0));
        cudaStream_t mainStream = c->stream;
        c->stream = c->commStream;
        int cst = comm_allreduce_max(c, cell, 1);
        if (cst == IDP_OK) cst = comm_allreduce_sum(c, cell + 1, 1);
        if (cst == IDP_OK) cst = comm_allreduce_sum(c, c->gbuf.p, 3L * c->nV);
        const cudaError_t ce = cudaEventRecord(c->evJoin, c->commStream);
        c->stream = mainStream;
        if (cst != IDP_OK) return cst;
        IDP_CK(c, ce);
        const int ast = st == IDP_OK ? assemble_csr(c) : IDP_OK;
        IDP_CK(c, cudaStreamWaitEvent(c->stream, c->evJoin, 0));
        IDP_CK(c, cudaMemcpyAsync(pack, cell, sizeof(pack), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        if (st != IDP_OK) return st;
        if ((int)pack[0] != IDP_OK) { c->err = "another rank of the sharded operator failed"; return (int)pack[0]; }
        IDP_TRY(ast);
        if (E_inout) *E_inout += pack[1];
        if (nnz) *nnz = c->nnz;
        return IDP_OK;
    }
    IDP_TRY(comm_agree_status(c, barrier_eval(c, dhat2, kappa, thickness, 1, 1, 1, project_spd, &E)));
    IDP_TRY(assemble_csr(c));
    IDP_TRY(finish_energy(c, E, E_inout));
    IDP_TRY(finish_gradient(c, nullptr, 3));
    if (nnz) *nnz = c->nnz;
    return IDP_OK;
}
// ---- the rest of the Newton system around the barrier Hessian -----------------------------------------
int idp_system_set_flow_term(idp_ctx* c, int n_elem, const int* elem3, int stride, const double* vol, double h)
{
    if (!c || n_elem < 0 || (n_elem && (!elem3 || !vol || stride < 3))) return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_flow_term: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->nFlowElem = 0;
    if (n_elem == 0) return IDP_OK;
    std::vector<int> e3(3 * (size_t)n_elem);
    for (long e = 0; e < n_elem; ++e) {
        const int a = elem3[(long)stride * e], b = elem3[(long)stride * e + 1], d = elem3[(long)stride * e + 2];
        if (a < 0 || b < 0 || d < 0 || a >= c->nV || b >= c->nV || d >= c->nV || a == b || b == d || a == d)
            return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_flow_term: element with out-of-range or repeated vertices", __FILE__, __LINE__);
        e3[3 * e] = a; e3[3 * e + 1] = b; e3[3 * e + 2] = d;
    }
    IDP_CK(c, c->flowElem.reserve(3 * (size_t)n_elem));
    IDP_CK(c, c->flowVol.reserve(n_elem));
    IDP_CK(c, cudaMemcpyAsync(c->flowElem.p, e3.data(), e3.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->flowVol.p, vol, (size_t)n_elem * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->nFlowElem = n_elem;
    c->flowH = h;
    return IDP_OK;
}
int idp_system_set_mass(idp_ctx* c, const double* m_per_vertex)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->haveMass = false;
    if (!m_per_vertex) return IDP_OK;
    IDP_CK(c, c->massDiag.reserve(c->nV));
    IDP_CK(c, cudaMemcpyAsync(c->massDiag.p, m_per_vertex, (size_t)c->nV * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->haveMass = true;
    return IDP_OK;
}
// ---- elastic terms of the shell system (elastic_kernels.cu) ------------------------------------------
int idp_system_set_membrane(idp_ctx* c, int n_elem, const int* elem3, int stride, const double* ib3, const double* vol, const double* lambda,
    const double* mu, double h)
{
    if (!c || n_elem < 0 || (n_elem && (!elem3 || !ib3 || !vol || !lambda || !mu || stride < 3)))
        return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_membrane: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->nMem = 0;
    if (n_elem == 0) return IDP_OK;
    std::vector<int> e3(3 * (size_t)n_elem);
    std::vector<double> coef((size_t)n_elem);
    for (long e = 0; e < n_elem; ++e) {
        const int a = elem3[(long)stride * e], b = elem3[(long)stride * e + 1], d = elem3[(long)stride * e + 2];
        if (a < 0 || b < 0 || d < 0 || a >= c->nV || b >= c->nV || d >= c->nV || a == b || b == d || a == d)
            return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_membrane: element with out-of-range or repeated vertices", __FILE__, __LINE__);
        e3[3 * e] = a; e3[3 * e + 1] = b; e3[3 * e + 2] = d;
        coef[e] = h * h * vol[e];
    }
    IDP_CK(c, c->memElem.reserve(3 * (size_t)n_elem)); IDP_CK(c, c->memIB.reserve(3 * (size_t)n_elem));
    IDP_CK(c, c->memCoef.reserve(n_elem)); IDP_CK(c, c->memLambda.reserve(n_elem)); IDP_CK(c, c->memMu.reserve(n_elem));
    IDP_CK(c, cudaMemcpyAsync(c->memElem.p, e3.data(), e3.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->memIB.p, ib3, 3 * (size_t)n_elem * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->memCoef.p, coef.data(), (size_t)n_elem * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->memLambda.p, lambda, (size_t)n_elem * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->memMu.p, mu, (size_t)n_elem * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->nMem = n_elem;
    return IDP_OK;
}
int idp_system_set_hinges(idp_ctx* c, int n_hinge, const int* stencil4, const double* info3, double k, double h)
{
    if (!c || n_hinge < 0 || (n_hinge && (!stencil4 || !info3)))
        return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_hinges: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->nHinge = 0;
    if (n_hinge == 0) return IDP_OK;
    for (long e = 0; e < n_hinge; ++e) {
        const int* v = stencil4 + 4 * e;
        for (int i = 0; i < 4; ++i) {
            if (v[i] < 0 || v[i] >= c->nV) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_hinges: vertex index out of range", __FILE__, __LINE__);
            for (int j = 0; j < i; ++j)
                if (v[i] == v[j]) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_hinges: repeated vertex in a hinge stencil", __FILE__, __LINE__);
        }
        if (!(info3[3 * e + 2] != 0.0)) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_system_set_hinges: zero rest height", __FILE__, __LINE__);
    }
    IDP_CK(c, c->hingeV.reserve(4 * (size_t)n_hinge)); IDP_CK(c, c->hingeInfo.reserve(3 * (size_t)n_hinge));
    IDP_CK(c, cudaMemcpyAsync(c->hingeV.p, stencil4, 4 * (size_t)n_hinge * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->hingeInfo.p, info3, 3 * (size_t)n_hinge * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->nHinge = n_hinge;
    c->hingeKh2 = h * h * k;
    return IDP_OK;
}
int idp_elastic_energy(idp_ctx* c, double* E_inout)
{
    if (!c || !E_inout) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    double E = 0;
    IDP_TRY(elastic_energy_gradient(c, 1, 0, &E));
    *E_inout += E;
    return IDP_OK;
}
int idp_elastic_gradient(idp_ctx* c, double* g_accum, int stride)
{
    if (!c || (g_accum && stride < 3)) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(elastic_energy_gradient(c, 0, 1, nullptr));
    if (!g_accum) return IDP_OK;
    std::vector<double> g(3 * (size_t)c->nV);
    IDP_CK(c, cudaMemcpyAsync(g.data(), c->elasticG.p, g.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    for (long v = 0; v < c->nV; ++v)
        for (int a = 0; a < 3; ++a) g_accum[(long)stride * v + a] += g[3 * v + a];
    return IDP_OK;
}
// ---- lagged friction (friction_kernels.cu) -----------------------------------------------------------
int idp_friction_update(idp_ctx* c, double dhat2, double kappa, double thickness, long* n_friction_rows)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(friction_update(c, dhat2, kappa, thickness));
    if (n_friction_rows) *n_friction_rows = c->nFricActive;
    return IDP_OK;
}
int idp_friction_set(idp_ctx* c, const double* xn, int stride, double epsv2_h2, double mu)
{
    if (!c || !(mu >= 0) || (mu > 0 && (!xn || !(epsv2_h2 > 0)))) return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_friction_set: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->fricMu = mu;
    c->fricEpsvh = mu > 0 ? std::sqrt(epsv2_h2) : 0.0;
    if (xn) {
        IDP_TRY(upload_positions(c, xn, stride, 3));
        c->have_xn = true;
    }
    return IDP_OK;
}
int idp_friction_set_components(idp_ctx* c, int n_comp, const int* comp_node_range, const double* mu_comp)
{
    if (!c || n_comp < 0 || (n_comp && (!comp_node_range || !mu_comp))) return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_friction_set_components: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    c->nFricComp = 0;
    if (n_comp == 0) return IDP_OK;
    IDP_CK(c, c->fricCompRange.reserve(n_comp)); IDP_CK(c, c->fricMuComp.reserve((size_t)n_comp * n_comp));
    IDP_CK(c, cudaMemcpyAsync(c->fricCompRange.p, comp_node_range, (size_t)n_comp * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaMemcpyAsync(c->fricMuComp.p, mu_comp, (size_t)n_comp * n_comp * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    c->nFricComp = n_comp;
    return IDP_OK;
}
int idp_friction_energy(idp_ctx* c, double* E_inout)
{
    if (!c || !E_inout) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    double E = 0;
    IDP_TRY(friction_energy_gradient(c, 1, 0, &E));
    *E_inout += E;
    return IDP_OK;
}
int idp_friction_gradient(idp_ctx* c, double* g_accum, int stride)
{
    if (!c || (g_accum && stride < 3)) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(friction_energy_gradient(c, 0, 1, nullptr));
    if (!g_accum) return IDP_OK;
    std::vector<double> g(3 * (size_t)c->nV);
    IDP_CK(c, cudaMemcpyAsync(g.data(), c->fricG.p, g.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    for (long v = 0; v < c->nV; ++v)
        for (int a = 0; a < 3; ++a) g_accum[(long)stride * v + a] += g[3 * v + a];
    return IDP_OK;
}
int idp_get_friction(idp_ctx* c, long* n_rows, int* rows4, double* closest2, double* basis6, double* normal_force)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (n_rows) *n_rows = c->nFricActive;
    if (!rows4 && !closest2 && !basis6 && !normal_force) return IDP_OK;
    return friction_copy_rows(c, rows4, closest2, basis6, normal_force);
}
int idp_project_dbc(idp_ctx* c)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    return project_dbc(c);
}
int idp_project_dbc_mask(idp_ctx* c, const uint8_t* mask)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    return project_dbc(c, mask);
}
int idp_solve_pcg(idp_ctx* c, const double* rhs, double* sol, double rel_tol, int max_iter, int* iters, double* rel_residual)
{
    if (!c || !rhs || max_iter < 0 || !(rel_tol >= 0)) return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_solve_pcg: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    return solve_pcg(c, rhs, sol, rel_tol, max_iter, iters, rel_residual);
}
int idp_set_mesh_from_triangles(idp_ctx* c, int nV, int nF, const int* tri, int stride, const double* x, int xstride, const uint8_t* dbc)
{
    if (!c || nV <= 0 || nF < 0 || (nF && (!tri || stride < 3)) || (x && xstride < 3))
        return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_set_mesh_from_triangles: bad arguments", __FILE__, __LINE__) : IDP_ERR_INVALID;
    for (long i = 0; i < nF; ++i)
        for (int k = 0; k < 3; ++k)
            if (tri[(long)stride * i + k] < 0 || tri[(long)stride * i + k] >= nV) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_set_mesh_from_triangles: vertex index out of range", __FILE__, __LINE__);
    IDP_CK(c, cudaSetDevice(c->device));
    c->nFlowElem = 0; c->haveMass = false; c->nMem = 0; c->nHinge = 0; c->nFric = 0; c->nFricActive = 0; c->have_xn = false; // a new mesh drops the terms of the previous one
    return extract_surface(c, nV, nF, tri, stride, x, xstride, dbc);
}
int idp_get_surface_primitives(idp_ctx* c, int* nBN, int* bnode, int* nBE, int* bedge2, int* nBT, int* btri3, double* BNArea, double* BEArea, double* BTArea)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (nBN) *nBN = c->nBN;
    if (nBE) *nBE = c->nBE;
    if (nBT) *nBT = c->nBT;
    if (bnode && c->nBN) IDP_CK(c, cudaMemcpyAsync(bnode, c->bnode.p, c->nBN * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (bedge2 && c->nBE) IDP_CK(c, cudaMemcpyAsync(bedge2, c->bedge.p, c->nBE * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
    std::vector<int4> t4;
    if (btri3 && c->nBT) {
        t4.resize(c->nBT);
        IDP_CK(c, cudaMemcpyAsync(t4.data(), c->btri.p, c->nBT * sizeof(int4), cudaMemcpyDeviceToHost, c->stream));
    }
    if ((BNArea || BEArea || BTArea) && !c->surfValid) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "areas exist only after idp_set_mesh_from_triangles", __FILE__, __LINE__);
    if (BNArea && c->nBN) IDP_CK(c, cudaMemcpyAsync(BNArea, c->surfNodeAreaC.p, c->nBN * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (BEArea && c->nBE) IDP_CK(c, cudaMemcpyAsync(BEArea, c->surfEdgeArea.p, c->nBE * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (BTArea && c->nBT) IDP_CK(c, cudaMemcpyAsync(BTArea, c->surfTriAreaH.p, c->nBT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < t4.size(); ++i) { btri3[3 * i] = t4[i].x; btri3[3 * i + 1] = t4[i].y; btri3[3 * i + 2] = t4[i].z; }
    return IDP_OK;
}
int idp_get_hessian_csr(idp_ctx* c, int* ptr, int* col, double* val)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (ptr) IDP_CK(c, cudaMemcpyAsync(ptr, c->csrPtr.p, (3 * (size_t)c->nV + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (col && c->nnz) IDP_CK(c, cudaMemcpyAsync(col, c->csrCol.p, c->nnz * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (val && c->nnz) IDP_CK(c, cudaMemcpyAsync(val, c->csrVal.p, c->nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    return IDP_OK;
}
// ---- asynchronous result transfers ---------------------------------------------------------------------------------
// The results a host-side Newton solve consumes are large (4 M triangles: 465 MB of rows, 4.8 GB of CSR). The *_begin
// calls enqueue their device-to-host copies on a copy stream that only waits for what has been computed so far and
// return; the caller goes on with the next operator (barrier evaluation while the rows travel, CCD and min-distance
// while the CSR travels) and collects everything with idp_transfers_end. The CSR crosses PCIe in compact form -- values,
// plus ONE column vertex per 3x3 block and the block-row starts (3.4 GB instead of 4.8) -- and the scalar (ptr, col) arrays
// Construct_From_CSR takes are expanded by host threads while the values are still in flight.
__global__ void __launch_bounds__(256) k_block_cols(const int* __restrict__ rowStart, const int* __restrict__ col, int nV, int* __restrict__ blkCol)
{
    const int lane = threadIdx.x & 31;
    for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < nV; v += (gridDim.x * blockDim.x) >> 5) {
        const int bs = rowStart[v], nb = rowStart[v + 1] - bs;
        for (int s = lane; s < nb; s += 32) blkCol[bs + s] = col[9L * bs + 3 * s] / 3;
    }
}
static int copy_stream_ready(idp_ctx* c)
{
    if (!c->copyStream) {
        IDP_CK(c, cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
        IDP_CK(c, cudaEventCreateWithFlags(&c->evCopyFork, cudaEventDisableTiming));
        IDP_CK(c, cudaEventCreateWithFlags(&c->evRows, cudaEventDisableTiming));
        IDP_CK(c, cudaEventCreateWithFlags(&c->evBlk, cudaEventDisableTiming));
        IDP_CK(c, cudaEventCreateWithFlags(&c->evVal, cudaEventDisableTiming));
    }
    IDP_CK(c, cudaEventRecord(c->evCopyFork, c->stream));
    IDP_CK(c, cudaStreamWaitEvent(c->copyStream, c->evCopyFork, 0));
    return IDP_OK;
}
// a large copy is cut into pieces so that the small read-backs of the operators running meanwhile (counters, scalars) are not
// queued behind gigabytes on the same copy engine
static int copy_d2h_chunked(idp_ctx* c, void* dst, const void* src, size_t bytes)
{
    const size_t chunk = 32u << 20;
    for (size_t o = 0; o < bytes; o += chunk)
        IDP_CK(c, cudaMemcpyAsync((char*)dst + o, (const char*)src + o, std::min(chunk, bytes - o), cudaMemcpyDeviceToHost, c->copyStream));
    return IDP_OK;
}
int idp_get_constraints_begin(idp_ctx* c, int* rows4, double* info2)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (c->pendRows) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_get_constraints_begin: a transfer is already pending (idp_transfers_end)", __FILE__, __LINE__);
    if (!c->weights_all_one && info2) { // caller-supplied weights: nothing to overlap, the plain getter does it
        return idp_get_constraints(c, rows4, info2);
    }
    IDP_TRY(copy_stream_ready(c));
    if (rows4 && c->nRows) IDP_TRY(copy_d2h_chunked(c, rows4, c->rows.p, c->nRows * sizeof(Row4)));
    IDP_CK(c, cudaEventRecord(c->evRows, c->copyStream));
    c->pendRows = true;
    if (info2 && c->nRows) { // weights are all one (OIPC, IPC.h:656-660): stencilInfo is filled by host threads while the rows travel
        const double dh2 = c->cs_dhat2;
        const long n = c->nRows;
        c->rowsWorker = new std::thread([=]() {
            host_parallel(n, [=](long b, long e) { for (long i = b; i < e; ++i) { info2[2 * i] = 1.0; info2[2 * i + 1] = dh2; } });
        });
    }
    return IDP_OK;
}
int idp_get_hessian_csr_begin(idp_ctx* c, int* ptr, int* col, double* val)
{
    if (!c || !ptr || !col || !val) return c ? fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_get_hessian_csr_begin: ptr, col and val are required", __FILE__, __LINE__) : IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (c->pendCsr) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "idp_get_hessian_csr_begin: a transfer is already pending (idp_transfers_end)", __FILE__, __LINE__);
    if (c->nnz <= 0) return idp_get_hessian_csr(c, ptr, col, val);
    IDP_TRY(copy_stream_ready(c));
    const long nBlk = c->nBlocksUnique;
    const size_t need = (size_t)c->nV + 1 + (size_t)nBlk;
    if (need > c->h_blk_cap) {
        if (c->h_blk) cudaFreeHost(c->h_blk);
        c->h_blk = nullptr; c->h_blk_cap = 0;
        IDP_CK(c, cudaMallocHost((void**)&c->h_blk, (need + need / 8) * sizeof(int)));
        c->h_blk_cap = need + need / 8;
    }
    IDP_CK(c, c->blkCol.reserve(nBlk));
    k_block_cols<<<std::min(blocks_for(32L * c->nV, 256), (unsigned)c->sm_count * 16), 256, 0, c->copyStream>>>(c->rowStart.p, c->csrCol.p, c->nV, c->blkCol.p);
    ++c->launches;
    IDP_CK(c, cudaMemcpyAsync(c->h_blk, c->rowStart.p, ((size_t)c->nV + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->copyStream));
    IDP_TRY(copy_d2h_chunked(c, c->h_blk + c->nV + 1, c->blkCol.p, (size_t)nBlk * sizeof(int)));
    IDP_CK(c, cudaEventRecord(c->evBlk, c->copyStream));
    IDP_TRY(copy_d2h_chunked(c, val, c->csrVal.p, c->nnz * sizeof(double)));
    IDP_CK(c, cudaEventRecord(c->evVal, c->copyStream));
    c->pendCsr = true;
    // host threads expand (block-row starts, block columns) into the scalar ptr / col arrays while the values are in flight
    const int* rs = c->h_blk;
    const int* bc = c->h_blk + c->nV + 1;
    const long nV = c->nV;
    const int device = c->device;
    cudaEvent_t evBlk = c->evBlk;
    int* status = &c->workerStatus;
    c->workerStatus = 0;
    c->csrWorker = new std::thread([=]() {
        if (cudaSetDevice(device) != cudaSuccess || cudaEventSynchronize(evBlk) != cudaSuccess) { *status = IDP_ERR_CUDA; return; }
        host_parallel(nV, [=](long b, long e) {
            for (long v = b; v < e; ++v) {
                const long bs = rs[v], nb = rs[v + 1] - bs;
                for (int a = 0; a < 3; ++a) {
                    const long p0 = 9 * bs + (long)a * 3 * nb;
                    ptr[3 * v + a] = (int)p0;
                    int* o = col + p0;
                    for (long s = 0; s < nb; ++s) { const int c3 = 3 * bc[bs + s]; o[3 * s] = c3; o[3 * s + 1] = c3 + 1; o[3 * s + 2] = c3 + 2; }
                }
            }
        });
        ptr[3 * nV] = 9 * rs[nV];
    });
    return IDP_OK;
}
int idp_transfers_end(idp_ctx* c)
{
    if (!c) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    const bool dbg = getenv("IDP_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    double t1 = t0, t2 = t0;
    auto join = [](void*& w) { if (w) { std::thread* t = (std::thread*)w; t->join(); delete t; w = nullptr; } };
    if (c->pendRows) {
        c->pendRows = false;
        join(c->rowsWorker);
        IDP_CK(c, cudaEventSynchronize(c->evRows));
    }
    t1 = now();
    if (c->pendCsr) {
        c->pendCsr = false;
        join(c->csrWorker);
        t2 = now();
        IDP_CK(c, cudaEventSynchronize(c->evVal));
        if (c->workerStatus != IDP_OK) return fail(c, IDP_ERR_CUDA, "%s (%s:%d)", "idp_transfers_end: the CSR expansion thread failed", __FILE__, __LINE__);
    }
    if (dbg) fprintf(stderr, "[idp] transfers_end: rows %.1f ms, CSR expansion %.1f, wait values %.1f\n", t1 - t0, t2 - t1, now() - t2);
    return IDP_OK;
}
int idp_hessian_csr_device(idp_ctx* c, const int** d_ptr, const int** d_col, const double** d_val, long* nnz)
{
    if (!c) return IDP_ERR_INVALID;
    if (d_ptr) *d_ptr = c->csrPtr.p;
    if (d_col) *d_col = c->csrCol.p;
    if (d_val) *d_val = c->csrVal.p;
    if (nnz) *nnz = c->nnz;
    return IDP_OK;
}
int idp_get_gradient(idp_ctx* c, double* g_accum, int stride)
{
    if (!c || !g_accum) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (stride < 3) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "stride must be >= 3", __FILE__, __LINE__);
    if (c->gbuf.cap < 3 * (size_t)c->nV) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "no gradient has been computed", __FILE__, __LINE__);
    double* g = nullptr;
    const long nV = c->nV;
    IDP_TRY(host_stage(c, 3 * (size_t)nV * sizeof(double), (void**)&g));
    IDP_CK(c, cudaMemcpyAsync(g, c->gbuf.p, 3 * (size_t)nV * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    host_parallel(nV, [=](long b, long e) {
        for (long v = b; v < e; ++v)
            for (int a = 0; a < 3; ++a) g_accum[v * stride + a] += g[3 * v + a];
    });
    return IDP_OK;
}
int idp_gradient_device(idp_ctx* c, const double** d_g)
{
    if (!c || !d_g) return IDP_ERR_INVALID;
    *d_g = c->gbuf.p;
    return IDP_OK;
}

int idp_set_search_direction(idp_ctx* c, const double* dir, int stride)
{
    if (!c || !dir) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    IDP_TRY(upload_positions(c, dir, stride, 2));
    c->have_dir = true;
    return IDP_OK;
}
int idp_ccd_step_resident(idp_ctx* c, double thickness, double* alpha_inout)
{
    if (!c || !alpha_inout) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    if (!c->have_dir) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "search direction not set", __FILE__, __LINE__);
    double a = *alpha_inout;
    IDP_TRY(ccd_step(c, thickness, &a, 1));
    *alpha_inout = a;
    return IDP_OK;
}
int idp_ccd_step(idp_ctx* c, const double* dir, int stride, double thickness, double* alpha_inout)
{
    IDP_TRY(idp_set_search_direction(c, dir, stride));
    return idp_ccd_step_resident(c, thickness, alpha_inout);
}

int idp_min_dist2(idp_ctx* c, double thickness, double* dist2, double* min_out)
{
    if (!c || !min_out) return IDP_ERR_INVALID;
    IDP_CK(c, cudaSetDevice(c->device));
    return min_dist2(c, thickness, dist2, min_out);
}

int idp_set_shard(idp_ctx* c, int rank, int nranks)
{
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return IDP_ERR_INVALID;
    if (nranks > IDP_MAX_RANKS) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "at most IDP_MAX_RANKS (8) shards", __FILE__, __LINE__);
    if (c->nccl_comm || c->local_group) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "context belongs to a communicator", __FILE__, __LINE__);
    c->rank = rank;
    c->nranks = nranks;
    c->permValid = false;
    return IDP_OK;
}

long idp_kernel_launches(idp_ctx* c) { return c ? c->launches : 0; }
long idp_library_calls(idp_ctx* c) { return c ? c->lib_launches : 0; }
void idp_reset_counters(idp_ctx* c) { if (c) { c->launches = 0; c->lib_launches = 0; } }
float idp_stage_ms(idp_ctx* c, int stage)
{
    if (!c || stage < 0 || stage >= IDP_STAGE_COUNT) return 0.f;
    idp::timers_resolve(c);
    return c->times.v[stage];
}
long idp_last_count(idp_ctx* c, int what)
{
    if (!c) return 0;
    switch (what) {
    case 0: return c->rowsLocal ? c->nRowsGlobal : c->nRows;
    case 1: return c->nCandPT;
    case 2: return c->nCandEE;
    case 3: return c->nCcdPT;
    case 4: return c->nCcdEE;
    case 5: return c->ccd_iters;
    case 6: return c->nnz;
    case 7: return c->nBlocksUnique;
    case 8: return device_alloc_counter();
    case 9: return c->nRows;
    default: return 0;
    }
}

} // extern "C"
