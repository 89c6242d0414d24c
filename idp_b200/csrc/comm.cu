// NCCL plumbing for the natural reductions of the sharded path (energy sum, gradient all-reduce, global min of the
// CCD step). libnccl is loaded at run time (dlopen) so that single-GPU use has no NCCL dependency.
#include "ctx.cuh"
#include <dlfcn.h>
#include <string.h>
#include <condition_variable>
#include <mutex>

namespace idp {

// ------------------------------------------------------------------------------------------------------------
// In-process group (idp_comm_init_local): several contexts of ONE process -- each driven by its own host thread, on
// the same or on different GPUs -- run the sharded path with the collectives done over peer memory: every rank
// publishes its device pointer, the ranks meet at a host barrier, and each rank copies / reduces what it needs straight
// out of its peers' buffers on its own stream (cudaMemcpyAsync device-to-device, or a reduction kernel that loads the
// P peer buffers in rank order -- deterministic, identical on every rank). Two host barriers per collective.
// Used by the -m gpu sharding parity tests (P shards on one GPU) and by single-process multi-GPU hosts.
// ------------------------------------------------------------------------------------------------------------
struct LocalGroup {
    int P = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    long generation = 0;
    int refs = 0;
    const void* ptr[IDP_MAX_RANKS] = {};
    long long cnt[IDP_MAX_RANKS][2 * IDP_MAX_RANKS] = {};
    int status[IDP_MAX_RANKS] = {};
    bool aborted = false; // idp_comm_abort: a rank gave up (host-side failure); every waiter returns with an error
    bool barrier()
    {
        std::unique_lock<std::mutex> lk(m);
        if (aborted) return false;
        const long gen = generation;
        if (++arrived == P) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen || aborted; });
        return !aborted;
    }
};
#define IDP_MEET(c, g) do { if (!(g)->barrier()) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "in-process group aborted by another rank", __FILE__, __LINE__); } while (0)
static inline LocalGroup* lg(idp_ctx* c) { return (LocalGroup*)c->local_group; }
// publish this rank's pointer, make its pending work visible, meet
static int local_open(idp_ctx* c, const void* p)
{
    LocalGroup* g = lg(c);
    g->ptr[c->rank] = p;
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    IDP_MEET(c, g);
    return IDP_OK;
}
// finish this rank's reads of its peers' buffers, meet (after this the buffers may be overwritten)
static int local_close(idp_ctx* c)
{
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    IDP_MEET(c, lg(c));
    return IDP_OK;
}
struct PeerPtrs { const void* p[IDP_MAX_RANKS]; int P; };
// OP 0: sum (rank order), 1: min of doubles, 2: min of uint64
template <int OP>
__global__ void k_peer_reduce(PeerPtrs pp, long n, void* __restrict__ out)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        if (OP == 2) {
            unsigned long long v = ((const unsigned long long*)pp.p[0])[i];
            for (int r = 1; r < pp.P; ++r) { const unsigned long long w = ((const unsigned long long*)pp.p[r])[i]; v = w < v ? w : v; }
            ((unsigned long long*)out)[i] = v;
        }
        else {
            double v = ((const double*)pp.p[0])[i];
            for (int r = 1; r < pp.P; ++r) { const double w = ((const double*)pp.p[r])[i]; v = OP == 0 ? v + w : fmin(v, w); }
            ((double*)out)[i] = v;
        }
    }
}
template <int OP>
static int local_allreduce(idp_ctx* c, void* dev, long n)
{
    LocalGroup* g = lg(c);
    IDP_CK(c, c->commTmp.reserve((size_t)std::max<long>(n, 1)));
    IDP_TRY(local_open(c, dev));
    PeerPtrs pp;
    pp.P = g->P;
    for (int r = 0; r < g->P; ++r) pp.p[r] = g->ptr[r];
    IDP_LAUNCH(c, k_peer_reduce<OP>, std::min(blocks_for(n, 256), (unsigned)c->sm_count * 8), 256, 0, pp, n, (void*)c->commTmp.p);
    IDP_CK(c, cudaGetLastError());
    IDP_TRY(local_close(c));
    IDP_CK(c, cudaMemcpyAsync(dev, c->commTmp.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
    return IDP_OK;
}

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void**, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_send)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_recv)(void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* h = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
    fn_allgather allgather = nullptr;
    fn_broadcast broadcast = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_send send = nullptr;
    fn_recv recv = nullptr;
};
static NcclApi g_nccl;

// accumulates the device time of the collectives of one step into IDP_STAGE_COMM (resolved lazily, no host sync)
struct CommTimer : ScopeTimer {
    CommTimer(idp_ctx* ctx) : ScopeTimer(ctx, IDP_STAGE_COMM, true) {}
};

static bool load_nccl()
{
    if (g_nccl.h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return false;
    g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.destroy = (fn_destroy)dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(g_nccl.h, "ncclGetErrorString");
    g_nccl.allgather = (fn_allgather)dlsym(g_nccl.h, "ncclAllGather");
    g_nccl.broadcast = (fn_broadcast)dlsym(g_nccl.h, "ncclBroadcast");
    g_nccl.group_start = (fn_group)dlsym(g_nccl.h, "ncclGroupStart");
    g_nccl.group_end = (fn_group)dlsym(g_nccl.h, "ncclGroupEnd");
    g_nccl.send = (fn_send)dlsym(g_nccl.h, "ncclSend");
    g_nccl.recv = (fn_recv)dlsym(g_nccl.h, "ncclRecv");
    return g_nccl.get_uid && g_nccl.init_rank && g_nccl.allreduce && g_nccl.destroy && g_nccl.allgather && g_nccl.broadcast &&
           g_nccl.group_start && g_nccl.group_end && g_nccl.send && g_nccl.recv;
}

// ncclDataType_t: ncclFloat64 = 8; ncclRedOp_t: ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4
int comm_allreduce_sum(idp_ctx* c, double* dev, long n)
{
    if (c->local_group) return local_allreduce<0>(c, dev, n);
    if (!c->nccl_comm) return IDP_OK;
    CommTimer tm(c);
    const int r = g_nccl.allreduce(dev, dev, (size_t)n, 8, 0, c->nccl_comm, c->stream);
    if (r != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r) : "?", __FILE__, __LINE__);
    return IDP_OK;
}
int comm_allreduce_min(idp_ctx* c, double* dev, long n)
{
    if (c->local_group) return local_allreduce<1>(c, dev, n);
    if (!c->nccl_comm) return IDP_OK;
    CommTimer tm(c);
    const int r = g_nccl.allreduce(dev, dev, (size_t)n, 8, 3, c->nccl_comm, c->stream);
    if (r != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r) : "?", __FILE__, __LINE__);
    return IDP_OK;
}
int comm_allreduce_max(idp_ctx* c, double* dev, long n)
{
    if (c->local_group) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "comm_allreduce_max: NCCL communicators only", __FILE__, __LINE__);
    if (!c->nccl_comm) return IDP_OK;
    CommTimer tm(c);
    const int r = g_nccl.allreduce(dev, dev, (size_t)n, 8, 2 /*ncclMax*/, c->nccl_comm, c->stream);
    if (r != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r) : "?", __FILE__, __LINE__);
    return IDP_OK;
}
// max over ranks of a status code, so that every rank leaves an operator with the same verdict (a rank-local failure
// must not strand its peers in the next collective; the reference exit(-1)s the whole process at these points)
int comm_agree_status(idp_ctx* c, int status)
{
    if (!comm_on(c)) return status;
    if (c->local_group) {
        LocalGroup* g = lg(c);
        g->status[c->rank] = status;
        if (!g->barrier()) return std::max(status, (int)IDP_ERR_INVALID);
        int mx = 0;
        for (int r = 0; r < g->P; ++r) mx = std::max(mx, g->status[r]);
        if (!g->barrier()) return std::max(status, (int)IDP_ERR_INVALID);
        if (mx != status && status == IDP_OK) c->err = "another rank of the sharded operator failed";
        return mx;
    }
    // NCCL: a failed rank may have left its stream / buffers in any state, so only the code travels (host value -> device cell)
    double v = (double)status;
    double* cell = (double*)(c->counters.p + CNT_STATUS);
    if (cudaMemcpyAsync(cell, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return std::max(status, (int)IDP_ERR_CUDA);
    const int r = g_nccl.allreduce(cell, cell, 1, 8, 2 /*ncclMax*/, c->nccl_comm, c->stream);
    if (r != 0) return std::max(status, (int)IDP_ERR_NCCL);
    if (cudaMemcpyAsync(&v, cell, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)
        return std::max(status, (int)IDP_ERR_CUDA);
    const int mx = (int)v;
    if (mx != status && status == IDP_OK) c->err = "another rank of the sharded operator failed";
    return mx;
}
#define IDP_NCCL_BEGIN(c) do { const int r__ = g_nccl.group_start(); if (r__ != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r__) : "?", __FILE__, __LINE__); } while (0)
#define IDP_NCCL_END(c) do { const int r__ = g_nccl.group_end(); if (r__ != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r__) : "?", __FILE__, __LINE__); } while (0)
// min over ranks of order-encoded 64-bit keys (ncclUint64 = 5)
int comm_allreduce_min_u64(idp_ctx* c, unsigned long long* dev, long n)
{
    if (c->local_group) return local_allreduce<2>(c, dev, n);
    if (!c->nccl_comm) return IDP_OK;
    CommTimer tm(c);
    const int r = g_nccl.allreduce(dev, dev, (size_t)n, 5, 3, c->nccl_comm, c->stream);
    if (r != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r) : "?", __FILE__, __LINE__);
    return IDP_OK;
}
int comm_allgather_i64(idp_ctx* c, long long* dev, long perRank)
{
    if (c->local_group) {
        LocalGroup* g = lg(c);
        IDP_TRY(local_open(c, dev));
        for (int r = 0; r < g->P; ++r)
            if (r != c->rank)
                IDP_CK(c, cudaMemcpyAsync(dev + (size_t)r * perRank, (const long long*)g->ptr[r] + (size_t)r * perRank, (size_t)perRank * sizeof(long long), cudaMemcpyDefault, c->stream));
        return local_close(c);
    }
    if (!c->nccl_comm) return IDP_OK;
    CommTimer tm(c);
    const int r = g_nccl.allgather(dev + (size_t)c->rank * perRank, dev, (size_t)perRank, 4 /*ncclInt64*/, c->nccl_comm, c->stream);
    if (r != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r) : "?", __FILE__, __LINE__);
    return IDP_OK;
}
// LOCAL-ROWS mode: assemble the global per-row array [all direct PT | all direct EE | all merged], each group in rank
// order (the merged group in descending rank order: its owner ranges follow the vertex slabs) = the order of the
// unsharded path, from every rank's local [A|B|U] segments. Counts are already known
// (ctx.shardCnt), so this is three grouped broadcasts without any host round trip. Collective: every rank must call it.
int comm_gather_groups(idp_ctx* c, const void* local, size_t elemSize, void* globalOut)
{
    CommTimer tm(c);
    const int P = c->nranks;
    long goff = 0;
    if (c->local_group) {
        LocalGroup* g = lg(c);
        IDP_TRY(local_open(c, local));
        for (int k = 0; k < 3; ++k)
            for (int rr = 0; rr < P; ++rr) {
                const int r = (k == 2) ? P - 1 - rr : rr;
                long loff = 0;
                for (int kk = 0; kk < k; ++kk) loff += c->shardCnt[r][kk];
                const long n = c->shardCnt[r][k];
                if (n > 0)
                    IDP_CK(c, cudaMemcpyAsync((char*)globalOut + (size_t)goff * elemSize, (const char*)g->ptr[r] + (size_t)loff * elemSize, (size_t)n * elemSize, cudaMemcpyDefault, c->stream));
                goff += n;
            }
        return local_close(c);
    }
    IDP_NCCL_BEGIN(c);
    for (int k = 0; k < 3; ++k) {
        long loff = 0;
        for (int kk = 0; kk < k; ++kk) loff += c->shardCnt[c->rank][kk];
        for (int rr = 0; rr < P; ++rr) {
            const int r = (k == 2) ? P - 1 - rr : rr; // merged group: ascending key = descending leading vertex = descending rank
            const long n = c->shardCnt[r][k];
            if (n > 0) {
                char* dst = (char*)globalOut + (size_t)goff * elemSize;
                const void* src = (r == c->rank) ? (const void*)((const char*)local + (size_t)loff * elemSize) : (const void*)dst;
                const int rc = g_nccl.broadcast(src, dst, (size_t)n * elemSize, 0 /*ncclChar*/, r, c->nccl_comm, c->stream);
                if (rc != 0) { g_nccl.group_end(); return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(rc) : "?", __FILE__, __LINE__); }
            }
            goff += n;
        }
    }
    IDP_NCCL_END(c);
    return IDP_OK;
}
// variable-size all-gather of constraint rows: every rank ends with the concatenation (in rank order) of all ranks' rows
#define IDP_NCCL(c, call)                                                                                              \
    do {                                                                                                               \
        const int r__ = (call);                                                                                        \
        if (r__ != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r__) : "?", __FILE__, __LINE__); \
    } while (0)
int comm_allgatherv(idp_ctx* c, const void* local, long nLocal, size_t elemSize, void** outPtr, size_t* outCap, long outOffset, long* nTotal)
{
    CommTimer tm(c);
    const int P = c->nranks;
    long long* dcnt = c->counters.p + 8; // scratch slots (P <= IDP_MAX_RANKS)
    long long mine = nLocal;
    long long cnt[IDP_MAX_RANKS];
    if (c->local_group) {
        LocalGroup* g = lg(c);
        g->cnt[c->rank][0] = mine;
        IDP_TRY(local_open(c, local));
        for (int r = 0; r < P; ++r) cnt[r] = g->cnt[r][0];
    }
    else {
        IDP_CK(c, cudaMemcpyAsync(dcnt + c->rank, &mine, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        IDP_NCCL(c, g_nccl.allgather(dcnt + c->rank, dcnt, 1, 4 /*ncclInt64*/, c->nccl_comm, c->stream));
        IDP_CK(c, cudaMemcpyAsync(cnt, dcnt, P * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
    }
    long total = 0;
    long off[IDP_MAX_RANKS + 1];
    for (int r = 0; r < P; ++r) { off[r] = total; total += (long)cnt[r]; }
    off[P] = total;
    const size_t need = (size_t)(outOffset + total);
    if (need > *outCap) { // grow, keeping the first outOffset elements
        const size_t ncap = need + need / 8;
        void* np = nullptr;
        IDP_CK(c, cudaMalloc(&np, ncap * elemSize));
        if (*outPtr && outOffset > 0) IDP_CK(c, cudaMemcpyAsync(np, *outPtr, (size_t)outOffset * elemSize, cudaMemcpyDeviceToDevice, c->stream));
        IDP_CK(c, cudaStreamSynchronize(c->stream));
        if (*outPtr) cudaFree(*outPtr);
        *outPtr = np;
        *outCap = ncap;
    }
    char* out = (char*)*outPtr + (size_t)outOffset * elemSize;
    if (c->local_group) {
        LocalGroup* g = lg(c);
        for (int r = 0; r < P; ++r)
            if (cnt[r] > 0) IDP_CK(c, cudaMemcpyAsync(out + (size_t)off[r] * elemSize, g->ptr[r], (size_t)cnt[r] * elemSize, cudaMemcpyDefault, c->stream));
        *nTotal = total;
        return local_close(c);
    }
    IDP_NCCL(c, g_nccl.group_start());
    for (int r = 0; r < P; ++r) {
        if (cnt[r] == 0) continue;
        char* dst = out + (size_t)off[r] * elemSize;
        const int rc = g_nccl.broadcast(r == c->rank ? local : (const void*)dst, dst, (size_t)cnt[r] * elemSize, 0 /*ncclChar*/, r, c->nccl_comm, c->stream);
        if (rc != 0) { g_nccl.group_end(); return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(rc) : "?", __FILE__, __LINE__); }
    }
    IDP_NCCL(c, g_nccl.group_end());
    *nTotal = total;
    return IDP_OK;
}

// all-to-all of 64-bit keys: this rank sends keys[sendBegin[r] .. +sendCount[r]) to rank r and receives every rank's
// slice for it, concatenated in rank order. Used to route the duplicate-merge keys to the rank that owns their key range
// (halo-free: keys, not geometry, move).
int comm_exchange_keys(idp_ctx* c, const unsigned long long* keys, const long sendBegin[IDP_MAX_RANKS], const long sendCount[IDP_MAX_RANKS], DBuf<unsigned long long>& recv, long* nRecv)
{
    CommTimer tm(c);
    const int P = c->nranks;
    if (c->local_group) {
        LocalGroup* g = lg(c);
        for (int r = 0; r < P; ++r) { g->cnt[c->rank][2 * r] = sendBegin[r]; g->cnt[c->rank][2 * r + 1] = sendCount[r]; }
        IDP_TRY(local_open(c, keys));
        long total = 0;
        for (int s2 = 0; s2 < P; ++s2) total += (long)g->cnt[s2][2 * c->rank + 1];
        IDP_CK(c, recv.reserve(std::max<long>(total, 1)));
        long off = 0;
        for (int s2 = 0; s2 < P; ++s2) {
            const long n = (long)g->cnt[s2][2 * c->rank + 1];
            if (n > 0) IDP_CK(c, cudaMemcpyAsync(recv.p + off, (const unsigned long long*)g->ptr[s2] + g->cnt[s2][2 * c->rank], (size_t)n * 8, cudaMemcpyDefault, c->stream));
            off += n;
        }
        *nRecv = total;
        return local_close(c);
    }
    IDP_CK(c, c->commCounts.reserve(IDP_MAX_RANKS * IDP_MAX_RANKS)); // P send counts of this rank, gathered into a P x P matrix
    long long mine[IDP_MAX_RANKS];
    for (int r = 0; r < P; ++r) mine[r] = sendCount[r];
    IDP_CK(c, cudaMemcpyAsync(c->commCounts.p + (size_t)c->rank * P, mine, P * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    IDP_NCCL(c, g_nccl.allgather(c->commCounts.p + (size_t)c->rank * P, c->commCounts.p, (size_t)P, 4 /*ncclInt64*/, c->nccl_comm, c->stream));
    long long all[IDP_MAX_RANKS * IDP_MAX_RANKS];
    IDP_CK(c, cudaMemcpyAsync(all, c->commCounts.p, (size_t)P * P * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    IDP_CK(c, cudaStreamSynchronize(c->stream));
    long roff[IDP_MAX_RANKS + 1];
    long total = 0;
    for (int s2 = 0; s2 < P; ++s2) { roff[s2] = total; total += (long)all[s2 * P + c->rank]; }
    roff[P] = total;
    IDP_CK(c, recv.reserve(std::max<long>(total, 1)));
    IDP_NCCL(c, g_nccl.group_start());
    for (int r = 0; r < P; ++r) {
        const long ns = sendCount[r], nr = roff[r + 1] - roff[r];
        if (ns > 0) { const int rc = g_nccl.send(keys + sendBegin[r], (size_t)ns, 5 /*ncclUint64*/, r, c->nccl_comm, c->stream); if (rc) { g_nccl.group_end(); return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(rc) : "?", __FILE__, __LINE__); } }
        if (nr > 0) { const int rc = g_nccl.recv(recv.p + roff[r], (size_t)nr, 5, r, c->nccl_comm, c->stream); if (rc) { g_nccl.group_end(); return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(rc) : "?", __FILE__, __LINE__); } }
    }
    IDP_NCCL(c, g_nccl.group_end());
    *nRecv = total;
    return IDP_OK;
}

void comm_destroy(idp_ctx* c)
{
    if (c->nccl_comm && g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
    c->nccl_comm = nullptr;
    if (LocalGroup* g = lg(c)) {
        bool last;
        { std::lock_guard<std::mutex> lk(g->m); last = --g->refs == 0; }
        if (last) delete g;
        c->local_group = nullptr;
    }
}

} // namespace idp

extern "C" int idp_comm_unique_id(void* out_id128)
{
    using namespace idp;
    if (!out_id128 || !load_nccl()) return IDP_ERR_NCCL;
    nccl_uid id;
    if (g_nccl.get_uid(&id) != 0) return IDP_ERR_NCCL;
    memcpy(out_id128, &id, 128);
    return IDP_OK;
}

extern "C" int idp_comm_init(idp_ctx* c, int rank, int nranks, const void* id128)
{
    using namespace idp;
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return IDP_ERR_INVALID;
    if (nranks > IDP_MAX_RANKS) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "at most IDP_MAX_RANKS (8) ranks: one NVSwitch box", __FILE__, __LINE__);
    if (c->local_group || c->nccl_comm) return fail(c, IDP_ERR_INVALID, "%s (%s:%d)", "context already belongs to a group", __FILE__, __LINE__);
    if (!load_nccl()) return fail(c, IDP_ERR_NCCL, "%s (%s:%d)", "libnccl.so.2 not found", __FILE__, __LINE__);
    IDP_CK(c, cudaSetDevice(c->device));
    nccl_uid id;
    memcpy(&id, id128, 128);
    const int r = g_nccl.init_rank(&c->nccl_comm, nranks, id, rank);
    if (r != 0) return fail(c, IDP_ERR_NCCL, "NCCL error: %s at %s:%d", g_nccl.errstr ? g_nccl.errstr(r) : "?", __FILE__, __LINE__);
    c->rank = rank;
    c->nranks = nranks;
    c->permValid = false;
    return IDP_OK;
}

extern "C" int idp_comm_init_local(idp_ctx** ctxs, int nranks)
{
    using namespace idp;
    if (!ctxs || nranks < 1 || nranks > IDP_MAX_RANKS) return IDP_ERR_INVALID;
    for (int r = 0; r < nranks; ++r)
        if (!ctxs[r] || ctxs[r]->local_group || ctxs[r]->nccl_comm) return IDP_ERR_INVALID;
    // kernels of one rank load its peers' buffers directly: enable peer access between distinct devices
    for (int a = 0; a < nranks; ++a)
        for (int b = 0; b < nranks; ++b) {
            if (ctxs[a]->device == ctxs[b]->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device) != cudaSuccess || !can)
                return fail(ctxs[a], IDP_ERR_CUDA, "%s (%s:%d)", "no peer access between the devices of the group", __FILE__, __LINE__);
            cudaSetDevice(ctxs[a]->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctxs[a], IDP_ERR_CUDA, "CUDA error: %s at %s:%d", cudaGetErrorString(e), __FILE__, __LINE__);
            cudaGetLastError();
        }
    LocalGroup* g = new LocalGroup();
    g->P = nranks;
    g->refs = nranks;
    for (int r = 0; r < nranks; ++r) {
        ctxs[r]->local_group = g;
        ctxs[r]->rank = r;
        ctxs[r]->nranks = nranks;
        ctxs[r]->permValid = false;
    }
    return IDP_OK;
}

extern "C" int idp_comm_abort(idp_ctx* c)
{
    using namespace idp;
    if (!c) return IDP_ERR_INVALID;
    if (LocalGroup* g = lg(c)) {
        std::lock_guard<std::mutex> lk(g->m);
        g->aborted = true;
        g->cv.notify_all();
    }
    return IDP_OK;
}
