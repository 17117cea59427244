// comm.cu -- multi-GPU layer of libcsp_b200.so: NCCL clique, tree broadcast (C1), leaf-sharded fits with
// an all-gather of the records, query shards with a pipelined gather of (leaf id, value) to one rank (C2).
//
// The path shards with no exchange step, so the collectives are exactly three: one broadcast per tree version, one all-gather per
// re-fit, one gather per query batch.  NCCL is loaded with dlopen when the clique is formed (libnccl.so.2: the copy already in
// the process -- e.g. the one torch brought -- else the system library), so a single-GPU user never needs it.
// One host thread may drive several GPUs (single-process mode): every NCCL call sequence over the local ranks is grouped.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi N;

int load_nccl() {
    if (N.h) return CSP_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already uses, if any
    const char* env = getenv("CSP_NCCL_LIB");
    if (!h && env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return csp_set_error(CSP_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) ok = false; return p; };
    N.GetVersion = (decltype(N.GetVersion))sym("ncclGetVersion");
    N.GetUniqueId = (decltype(N.GetUniqueId))sym("ncclGetUniqueId");
    N.CommInitRank = (decltype(N.CommInitRank))sym("ncclCommInitRank");
    N.CommInitAll = (decltype(N.CommInitAll))sym("ncclCommInitAll");
    N.CommDestroy = (decltype(N.CommDestroy))sym("ncclCommDestroy");
    N.GroupStart = (decltype(N.GroupStart))sym("ncclGroupStart");
    N.GroupEnd = (decltype(N.GroupEnd))sym("ncclGroupEnd");
    N.Broadcast = (decltype(N.Broadcast))sym("ncclBroadcast");
    N.Send = (decltype(N.Send))sym("ncclSend");
    N.Recv = (decltype(N.Recv))sym("ncclRecv");
    N.AllReduce = (decltype(N.AllReduce))sym("ncclAllReduce");
    N.GetErrorString = (decltype(N.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return csp_set_error(CSP_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    N.h = h;
    return CSP_OK;
}

#define CSP_NCCL(call)                                                                                              \
    do {                                                                                                            \
        ncclResult_t _r = (call);                                                                                   \
        if (_r != ncclSuccess)                                                                                      \
            return csp_set_error(CSP_ERR_NCCL, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, N.GetErrorString(_r)); \
    } while (0)

struct Rank {
    int rank = 0, device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;   // library stream (used when the caller passes no stream)
    cudaStream_t cst = nullptr;  // communication stream: gathers run here behind the next chunk's kernels
    cudaStream_t aux[2] = {nullptr, nullptr};  // two more copy streams: pushes to different peers run on different copy engines
    cudaEvent_t aux_go = nullptr, aux_done[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t chunk_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
// Gather buffers owned by the library (csp_shard_gather_alloc): plain cudaMalloc memory on the root's GPU that every other rank
// can write with its copy engines -- through the unified address space in single-process mode (peer access enabled), through a
// CUDA IPC mapping in one-process-per-GPU mode.
struct GatherBuf {
    int root = -1;
    long long cap = 0;
    bool owner = false;                 // this process allocated them (it owns the root rank)
    int32_t* leaf = nullptr;            // address usable by this process's ranks (the allocation itself, or the IPC mapping)
    double* val = nullptr;
    int* flag = nullptr;                // per local rank a 4-byte device word for the completion signal
};
struct Ctx {
    bool init = false;
    bool defer_join = false;            // csp_shard_defer_join: gathers are joined by csp_shard_join instead of by every call
    GatherBuf gb;
    std::vector<int*> sig;              // per local rank: small device scratch for the completion signals
    int world = 0, version = 0;
    std::vector<Rank> local;
    cudaEvent_t tev[6][2] = {};  // timing events on local rank 0's streams
    bool tset[6] = {};
    double wall_ms[6] = {};
};
Ctx G;

void shard_range(long long total, int world, int rank, long long* lo, long long* hi) {
    const long long base = total / world, rem = total % world;
    *lo = rank * base + std::min<long long>(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

int finish_init() {
    for (Rank& r : G.local) {
        DeviceGuard g(r.device);
        CSP_CUDA(cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking));
        // the communication stream outranks the compute streams: NCCL's send / receive CTAs are placed as soon as a kernel of
        // the pipeline retires some of its own, instead of queueing behind the next kernel's persistent grid
        int prio_lo = 0, prio_hi = 0;
        CSP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CSP_CUDA(cudaStreamCreateWithPriority(&r.cst, cudaStreamNonBlocking, prio_hi));
        CSP_CUDA(cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming));
        CSP_CUDA(cudaEventCreateWithFlags(&r.join, cudaEventDisableTiming));
        for (auto& e : r.chunk_ev) CSP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            CSP_CUDA(cudaStreamCreateWithPriority(&r.aux[k], cudaStreamNonBlocking, prio_hi));
            CSP_CUDA(cudaEventCreateWithFlags(&r.aux_done[k], cudaEventDisableTiming));
        }
        CSP_CUDA(cudaEventCreateWithFlags(&r.aux_go, cudaEventDisableTiming));
    }
    {
        DeviceGuard g(G.local[0].device);
        for (auto& pr : G.tev)
            for (auto& e : pr) CSP_CUDA(cudaEventCreate(&e));
    }
    G.init = true;
    return CSP_OK;
}

int need_init() {
    if (!G.init) return csp_set_error(CSP_ERR_INVALID, "call csp_init or csp_init_rank first");
    return CSP_OK;
}
cudaStream_t stream_of(void* const* streams, size_t i) { return streams ? (cudaStream_t)streams[i] : G.local[i].st; }
void tbegin(int k, cudaStream_t st) { DeviceGuard g(G.local[0].device); cudaEventRecord(G.tev[k][0], st); }
void tend(int k, cudaStream_t st) { DeviceGuard g(G.local[0].device); cudaEventRecord(G.tev[k][1], st); G.tset[k] = true; }
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

struct csp_shard_tree {
    std::vector<csp_tree*> trees;
};
struct csp_shard_models {
    std::vector<csp_models*> models;
    // copy-engine all-gather of the fitted records: every rank's record / status arrays mapped by every other rank
    bool ce = false;
    std::vector<std::vector<double*>> peer_data;     // [local rank][rank] -> that rank's record array, as addressable here
    std::vector<std::vector<uint8_t*>> peer_status;
    std::vector<void*> ipc_opened;                   // mappings to close
};

extern "C" {

int csp_comm_info(int64_t info[4]) {
    if (!info) return csp_set_error(CSP_ERR_INVALID, "info is NULL");
    info[0] = G.init ? G.world : 0;
    info[1] = G.init ? (int64_t)G.local.size() : 0;
    info[2] = G.init ? G.local[0].rank : 0;
    info[3] = G.init ? G.version : 0;
    return CSP_OK;
}

int csp_init(int ndev, const int* devs) {
    if (G.init) return csp_set_error(CSP_ERR_INVALID, "already initialised (csp_shutdown first)");
    CSP_CHECK(csp_require_device());
    int have = 0;
    CSP_CUDA(cudaGetDeviceCount(&have));
    if (ndev < 1 || ndev > have) return csp_set_error(CSP_ERR_INVALID, "ndev=%d, %d CUDA devices present", ndev, have);
    std::vector<int> d(ndev);
    for (int i = 0; i < ndev; ++i) {
        d[i] = devs ? devs[i] : i;
        if (d[i] < 0 || d[i] >= have) return csp_set_error(CSP_ERR_INVALID, "device %d out of range", d[i]);
    }
    G.world = ndev;
    G.local.assign(ndev, Rank());
    for (int i = 0; i < ndev; ++i) { G.local[i].rank = i; G.local[i].device = d[i]; }
    if (ndev > 1) {
        CSP_CHECK(load_nccl());
        N.GetVersion(&G.version);
        std::vector<ncclComm_t> comms(ndev);
        CSP_NCCL(N.CommInitAll(comms.data(), ndev, d.data()));
        for (int i = 0; i < ndev; ++i) G.local[i].comm = comms[i];
    }
    return finish_init();
}

int csp_comm_unique_id(void* id128) {
    if (!id128) return csp_set_error(CSP_ERR_INVALID, "id128 is NULL");
    CSP_CHECK(load_nccl());
    ncclUniqueId id;
    CSP_NCCL(N.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof id);
    return CSP_OK;
}

int csp_init_rank(int world, int rank, int device, const void* id128) {
    if (G.init) return csp_set_error(CSP_ERR_INVALID, "already initialised (csp_shutdown first)");
    CSP_CHECK(csp_require_device());
    if (world < 1 || rank < 0 || rank >= world) return csp_set_error(CSP_ERR_INVALID, "bad rank %d of %d", rank, world);
    int have = 0;
    CSP_CUDA(cudaGetDeviceCount(&have));
    if (device < 0 || device >= have) return csp_set_error(CSP_ERR_INVALID, "device %d out of range", device);
    G.world = world;
    G.local.assign(1, Rank());
    G.local[0].rank = rank; G.local[0].device = device;
    if (world > 1) {
        if (!id128) return csp_set_error(CSP_ERR_INVALID, "id128 is NULL");
        CSP_CHECK(load_nccl());
        N.GetVersion(&G.version);
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        DeviceGuard g(device);
        CSP_NCCL(N.CommInitRank(&G.local[0].comm, world, id, rank));
    }
    return finish_init();
}

int csp_shard_gather_free(void);
int csp_shutdown(void) {
    if (!G.init) return CSP_OK;
    csp_shard_gather_free();
    for (size_t i = 0; i < G.sig.size(); ++i) { DeviceGuard g(G.local[i].device); if (G.sig[i]) cudaFree(G.sig[i]); }
    for (Rank& r : G.local) {
        DeviceGuard g(r.device);
        cudaDeviceSynchronize();
        if (r.comm) N.CommDestroy(r.comm);
        if (r.st) cudaStreamDestroy(r.st);
        if (r.cst) cudaStreamDestroy(r.cst);
        if (r.fork) cudaEventDestroy(r.fork);
        if (r.join) cudaEventDestroy(r.join);
        for (auto& e : r.chunk_ev) if (e) cudaEventDestroy(e);
        for (int k = 0; k < 2; ++k) { if (r.aux[k]) cudaStreamDestroy(r.aux[k]); if (r.aux_done[k]) cudaEventDestroy(r.aux_done[k]); }
        if (r.aux_go) cudaEventDestroy(r.aux_go);
    }
    {
        DeviceGuard g(G.local[0].device);
        for (auto& pr : G.tev)
            for (auto& e : pr) if (e) cudaEventDestroy(e);
    }
    G = Ctx();
    return CSP_OK;
}

int csp_shard_synchronize(void) {
    CSP_CHECK(need_init());
    for (Rank& r : G.local) {
        DeviceGuard g(r.device);
        CSP_CUDA(cudaStreamSynchronize(r.st));
        CSP_CUDA(cudaStreamSynchronize(r.cst));
    }
    return CSP_OK;
}

int csp_shard_timing(double ms[6]) {
    if (!ms) return csp_set_error(CSP_ERR_INVALID, "ms is NULL");
    CSP_CHECK(need_init());
    for (Rank& r : G.local) { DeviceGuard g(r.device); CSP_CUDA(cudaDeviceSynchronize()); }
    DeviceGuard g(G.local[0].device);
    for (int k = 0; k < 6; ++k) {
        ms[k] = G.wall_ms[k];
        float e = 0;
        if (G.tset[k] && cudaEventElapsedTime(&e, G.tev[k][0], G.tev[k][1]) == cudaSuccess) ms[k] = e;
    }
    cudaGetLastError();
    return CSP_OK;
}

// ---- C1: tree broadcast ------------------------------------------------------------------------------------------------
int csp_shard_tree_broadcast(const void* blob, size_t bytes, int root, csp_shard_tree** out) {
    if (!out) return csp_set_error(CSP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    CSP_CHECK(need_init());
    if (root < 0 || root >= G.world) return csp_set_error(CSP_ERR_INVALID, "root %d out of range", root);
    const size_t nl = G.local.size();
    int root_local = -1;
    for (size_t i = 0; i < nl; ++i)
        if (G.local[i].rank == root) root_local = (int)i;
    if (root_local >= 0 && (!blob || bytes < 64)) return csp_set_error(CSP_ERR_INVALID, "the root rank needs a packed tree blob");
    // 1. every rank learns the size
    unsigned long long sz = root_local >= 0 ? (unsigned long long)bytes : 0ull;
    if (G.world > 1 && nl < (size_t)G.world) {  // other processes exist
        std::vector<void*> dsz(nl, nullptr);
        for (size_t i = 0; i < nl; ++i) {
            DeviceGuard g(G.local[i].device);
            CSP_CHECK(csp_dev_alloc(&dsz[i], 8));
            if ((int)i == root_local) CSP_CUDA(cudaMemcpyAsync(dsz[i], &sz, 8, cudaMemcpyHostToDevice, G.local[i].st));
        }
        CSP_NCCL(N.GroupStart());
        for (size_t i = 0; i < nl; ++i) CSP_NCCL(N.Broadcast(dsz[i], dsz[i], 8, ncclUint8, root, G.local[i].comm, G.local[i].st));
        CSP_NCCL(N.GroupEnd());
        {
            DeviceGuard g(G.local[0].device);
            CSP_CUDA(cudaMemcpyAsync(&sz, dsz[0], 8, cudaMemcpyDeviceToHost, G.local[0].st));
            CSP_CUDA(cudaStreamSynchronize(G.local[0].st));
        }
        for (size_t i = 0; i < nl; ++i) { DeviceGuard g(G.local[i].device); cudaStreamSynchronize(G.local[i].st); csp_dev_free(dsz[i]); }
    }
    if (sz < 64) return csp_set_error(CSP_ERR_INVALID, "broadcast tree blob is too small (%llu bytes)", sz);
    // 2. blob to the root's GPU, 3. one broadcast, 4. upload on every GPU from the device buffer
    std::vector<void*> dbuf(nl, nullptr);
    auto cleanup = [&]() { for (size_t i = 0; i < nl; ++i) if (dbuf[i]) { DeviceGuard g(G.local[i].device); csp_dev_free(dbuf[i]); } };
    for (size_t i = 0; i < nl; ++i) {
        DeviceGuard g(G.local[i].device);
        int rc = csp_dev_alloc(&dbuf[i], sz);
        if (rc != CSP_OK) { cleanup(); return rc; }
    }
    if (root_local >= 0) {
        Rank& r = G.local[root_local];
        DeviceGuard g(r.device);
        if (root_local == 0) tbegin(0, r.st);
        cudaError_t e = cudaMemcpyAsync(dbuf[root_local], blob, sz, cudaMemcpyHostToDevice, r.st);
        if (root_local == 0) tend(0, r.st);
        if (e != cudaSuccess) { cleanup(); return csp_set_error(CSP_ERR_CUDA, "blob copy failed: %s", cudaGetErrorString(e)); }
    }
    if (G.world > 1) {
        tbegin(1, G.local[0].st);
        ncclResult_t r1 = N.GroupStart();
        for (size_t i = 0; i < nl && r1 == ncclSuccess; ++i)
            r1 = N.Broadcast(dbuf[i], dbuf[i], sz, ncclUint8, root, G.local[i].comm, G.local[i].st);
        ncclResult_t r2 = N.GroupEnd();
        tend(1, G.local[0].st);
        if (r1 != ncclSuccess || r2 != ncclSuccess) {
            cleanup();
            return csp_set_error(CSP_ERR_NCCL, "tree broadcast failed: %s", N.GetErrorString(r1 != ncclSuccess ? r1 : r2));
        }
    }
    csp_shard_tree* st = new csp_shard_tree();
    st->trees.assign(nl, nullptr);
    int rc = CSP_OK;
    for (size_t i = 0; i < nl; ++i) {
        DeviceGuard g(G.local[i].device);
        if (cudaStreamSynchronize(G.local[i].st) != cudaSuccess) { rc = csp_set_error(CSP_ERR_CUDA, "broadcast stream failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    }
    const double t0 = now_ms();
    for (size_t i = 0; i < nl && rc == CSP_OK; ++i) {
        rc = csp_set_device(G.local[i].device);
        if (rc == CSP_OK) rc = csp_tree_upload_blob(dbuf[i], sz, &st->trees[i]);
    }
    G.wall_ms[2] = now_ms() - t0;
    cleanup();
    if (rc != CSP_OK) { csp_shard_tree_free(st); return rc; }
    *out = st;
    return CSP_OK;
}

int csp_shard_tree_get(const csp_shard_tree* st, int local_index, csp_tree** tree) {
    if (!st || !tree || local_index < 0 || local_index >= (int)st->trees.size()) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    *tree = st->trees[local_index];
    return CSP_OK;
}

int csp_shard_tree_free(csp_shard_tree* st) {
    if (!st) return CSP_OK;
    for (csp_tree* t : st->trees) csp_tree_free(t);
    delete st;
    return CSP_OK;
}

// Map every rank's record and status arrays on every other rank (peer access inside a process, CUDA IPC between processes).
// Collective; false (on every rank alike) when some rank cannot map some other.
static int bcast_bytes(void* host, size_t n, int root);
static int setup_models_ce(csp_shard_models* sm) {
    const size_t nl = G.local.size();
    const int W = G.world;
    sm->peer_data.assign(nl, std::vector<double*>(W, nullptr));
    sm->peer_status.assign(nl, std::vector<uint8_t*>(W, nullptr));
    int ok = 1;
    for (int r = 0; r < W; ++r) {
        int owner = -1;
        for (size_t i = 0; i < nl; ++i)
            if (G.local[i].rank == r) owner = (int)i;
        struct Handles { cudaIpcMemHandle_t data, status; int ok; } hd{};
        if (owner >= 0) {
            hd.ok = 1;
            if (nl < (size_t)W) {
                DeviceGuard g(G.local[owner].device);
                if (cudaIpcGetMemHandle(&hd.data, sm->models[owner]->data) != cudaSuccess ||
                    cudaIpcGetMemHandle(&hd.status, sm->models[owner]->status) != cudaSuccess) { cudaGetLastError(); hd.ok = 0; }
            }
        }
        CSP_CHECK(bcast_bytes(&hd, sizeof hd, r));
        if (!hd.ok) ok = 0;
        for (size_t i = 0; i < nl && hd.ok; ++i) {
            if ((int)i == owner) {
                sm->peer_data[i][r] = sm->models[i]->data;
                sm->peer_status[i][r] = sm->models[i]->status;
            } else if (owner >= 0) {  // same process: the unified address space, peer access enabled
                DeviceGuard g(G.local[i].device);
                cudaError_t e = cudaDeviceEnablePeerAccess(G.local[owner].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                cudaGetLastError();
                sm->peer_data[i][r] = sm->models[owner]->data;
                sm->peer_status[i][r] = sm->models[owner]->status;
            } else {  // another process
                DeviceGuard g(G.local[i].device);
                void *pd = nullptr, *ps = nullptr;
                if (cudaIpcOpenMemHandle(&pd, hd.data, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                    cudaIpcOpenMemHandle(&ps, hd.status, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = 0;
                }
                if (pd) sm->ipc_opened.push_back(pd);
                if (ps) sm->ipc_opened.push_back(ps);
                sm->peer_data[i][r] = (double*)pd;
                sm->peer_status[i][r] = (uint8_t*)ps;
            }
        }
    }
    if (W > 1 && nl < (size_t)W) {  // agree: minimum over the ranks
        int* d = G.sig[0] + G.world + 1;
        DeviceGuard g(G.local[0].device);
        CSP_CUDA(cudaMemcpy(d, &ok, 4, cudaMemcpyHostToDevice));
        CSP_NCCL(N.AllReduce(d, d, 1, ncclInt32, ncclMin, G.local[0].comm, G.local[0].st));
        CSP_CUDA(cudaStreamSynchronize(G.local[0].st));
        CSP_CUDA(cudaMemcpy(&ok, d, 4, cudaMemcpyDeviceToHost));
    }
    sm->ce = ok != 0;
    if (sm->ce)
        for (size_t i = 0; i < nl; ++i) {
            DeviceGuard g(G.local[i].device);
            CSP_CUDA(cudaEventCreateWithFlags(&sm->models[i]->ready, cudaEventDisableTiming));
        }
    return CSP_OK;
}
static int ensure_sig() {
    if (!G.sig.empty()) return CSP_OK;
    const size_t nl = G.local.size();
    G.sig.assign(nl, nullptr);
    for (size_t i = 0; i < nl; ++i) {
        DeviceGuard g(G.local[i].device);
        CSP_CUDA(cudaMalloc((void**)&G.sig[i], (size_t)(G.world + 2) * 4));
        CSP_CUDA(cudaMemset(G.sig[i], 0, (size_t)(G.world + 2) * 4));
    }
    return CSP_OK;
}
// a barrier on the communication streams: one word all-reduced (every rank's kernel is ordered behind what its stream holds)
static int comm_barrier(const std::vector<cudaStream_t>& cs) {
    ncclResult_t r1 = N.GroupStart();
    for (size_t i = 0; i < G.local.size() && r1 == ncclSuccess; ++i)
        r1 = N.AllReduce(G.sig[i], G.sig[i], 1, ncclInt32, ncclSum, G.local[i].comm, cs[i]);
    ncclResult_t r2 = N.GroupEnd();
    if (r1 != ncclSuccess || r2 != ncclSuccess) return csp_set_error(CSP_ERR_NCCL, "barrier failed: %s", N.GetErrorString(r1 != ncclSuccess ? r1 : r2));
    return CSP_OK;
}

// ---- leaf-sharded fit + all-gather of the records -------------------------------------------------------------------------
int csp_shard_fit(const csp_shard_tree* st, int32_t skip, csp_shard_models** models, void* const* streams) {
    if (!st || !models) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    CSP_CHECK(need_init());
    const size_t nl = G.local.size();
    if (st->trees.size() != nl) return csp_set_error(CSP_ERR_INVALID, "sharded tree does not belong to this clique");
    csp_shard_models* sm = *models;
    const bool fresh = sm == nullptr;
    const bool want_ce = G.world > 1 && !(getenv("CSP_ALLGATHER") && !strcmp(getenv("CSP_ALLGATHER"), "nccl"));
    if (fresh) {
        sm = new csp_shard_models();
        sm->models.assign(nl, nullptr);
        for (size_t i = 0; i < nl; ++i) {
            int rc = csp_models_alloc_ex(st->trees[i], want_ce, &sm->models[i]);
            if (rc != CSP_OK) { csp_shard_models_free(sm); return rc; }
        }
        if (want_ce) {
            int rc = ensure_sig();
            if (rc == CSP_OK) rc = setup_models_ce(sm);
            if (rc != CSP_OK) { csp_shard_models_free(sm); return rc; }
        }
    } else if (sm->models.size() != nl) {
        return csp_set_error(CSP_ERR_INVALID, "sharded models do not belong to this clique");
    }
    auto fail = [&](int rc) { if (fresh) csp_shard_models_free(sm); return rc; };
    const long long nleaves = sm->models[0]->n_models;
    std::vector<cudaStream_t> s(nl), cs(nl);
    for (size_t i = 0; i < nl; ++i) { s[i] = stream_of(streams, i); cs[i] = G.local[i].cst; }
    if (sm->ce) {
        // Copy-engine all-gather, off the callers' streams.  (1) The communication streams step behind everything queued so far
        // (the previous evaluations) and meet in a barrier: no rank's records are overwritten while another still reads them.
        for (size_t i = 0; i < nl; ++i) {
            DeviceGuard g(G.local[i].device);
            CSP_CUDA(cudaEventRecord(G.local[i].fork, s[i]));
            CSP_CUDA(cudaStreamWaitEvent(cs[i], G.local[i].fork, 0));
        }
        int rc = comm_barrier(cs);
        if (rc != CSP_OK) return fail(rc);
    }
    for (size_t i = 0; i < nl; ++i) {
        long long lo, hi;
        shard_range(nleaves, G.world, G.local[i].rank, &lo, &hi);
        if (i == 0) tbegin(3, s[i]);
        int rc = csp_models_refit_range(st->trees[i], skip, sm->models[i], lo, hi, s[i]);
        if (i == 0) tend(3, s[i]);
        if (rc != CSP_OK) return fail(rc);
    }
    if (sm->ce) {
        // (2) behind the local fit, every rank pushes its record range into every other rank's array with its copy engines
        // (device-to-device DMA over NVLink), (3) a second barrier tells everyone that all ranges have landed, (4) the
        // models' `ready` event is recorded: the descent of the next batch runs meanwhile, its evaluation waits for the event.
        for (size_t i = 0; i < nl; ++i) {
            const Rank& rk = G.local[i];
            csp_models* mo = sm->models[i];
            long long lo, hi;
            shard_range(nleaves, G.world, rk.rank, &lo, &hi);
            DeviceGuard g(rk.device);
            CSP_CUDA(cudaEventRecord(rk.join, s[i]));
            CSP_CUDA(cudaStreamWaitEvent(cs[i], rk.join, 0));
            if (i == 0) tbegin(4, cs[i]);
            // the pushes to the different peers are spread over three streams, i.e. over several copy engines and NVLink ports
            CSP_CUDA(cudaEventRecord(rk.aux_go, cs[i]));
            for (int k = 0; k < 2; ++k) CSP_CUDA(cudaStreamWaitEvent(rk.aux[k], rk.aux_go, 0));
            for (int d = 1; d < G.world && hi > lo; ++d) {
                const int r = (rk.rank + d) % G.world;  // every rank starts with a different target
                cudaStream_t ps = d % 3 == 0 ? cs[i] : rk.aux[d % 3 - 1];
                CSP_CUDA(cudaMemcpyAsync(sm->peer_data[i][r] + (size_t)lo * mo->stride, mo->data + (size_t)lo * mo->stride,
                                         (size_t)(hi - lo) * mo->stride * 8, cudaMemcpyDeviceToDevice, ps));
                CSP_CUDA(cudaMemcpyAsync(sm->peer_status[i][r] + lo, mo->status + lo, (size_t)(hi - lo), cudaMemcpyDeviceToDevice, ps));
            }
            for (int k = 0; k < 2; ++k) {
                CSP_CUDA(cudaEventRecord(rk.aux_done[k], rk.aux[k]));
                CSP_CUDA(cudaStreamWaitEvent(cs[i], rk.aux_done[k], 0));
            }
        }
        int rc = comm_barrier(cs);
        if (rc != CSP_OK) return fail(rc);
        for (size_t i = 0; i < nl; ++i) {
            DeviceGuard g(G.local[i].device);
            if (i == 0) tend(4, cs[i]);
            CSP_CUDA(cudaEventRecord(sm->models[i]->ready, cs[i]));
        }
    } else if (G.world > 1) {
        // "all-gather-v" with NCCL: every owner broadcasts its range of records (and status bytes) in place; one group
        tbegin(4, s[0]);
        ncclResult_t r1 = N.GroupStart();
        for (size_t i = 0; i < nl && r1 == ncclSuccess; ++i) {
            csp_models* mo = sm->models[i];
            for (int r = 0; r < G.world && r1 == ncclSuccess; ++r) {
                long long lo, hi;
                shard_range(nleaves, G.world, r, &lo, &hi);
                if (hi <= lo) continue;
                double* p = mo->data + (size_t)lo * mo->stride;
                r1 = N.Broadcast(p, p, (size_t)(hi - lo) * mo->stride, ncclDouble, r, G.local[i].comm, s[i]);
                if (r1 == ncclSuccess) r1 = N.Broadcast(mo->status + lo, mo->status + lo, (size_t)(hi - lo), ncclUint8, r, G.local[i].comm, s[i]);
            }
        }
        ncclResult_t r2 = N.GroupEnd();
        tend(4, s[0]);
        if (r1 != ncclSuccess || r2 != ncclSuccess)
            return fail(csp_set_error(CSP_ERR_NCCL, "record all-gather failed: %s", N.GetErrorString(r1 != ncclSuccess ? r1 : r2)));
    }
    if (fresh) {
        for (size_t i = 0; i < nl; ++i) {
            DeviceGuard g(G.local[i].device);
            cudaStreamSynchronize(s[i]);
            cudaStreamSynchronize(cs[i]);
        }
        *models = sm;
    }
    return CSP_OK;
}

int csp_shard_models_get(const csp_shard_models* sm, int local_index, csp_models** models) {
    if (!sm || !models || local_index < 0 || local_index >= (int)sm->models.size()) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    *models = sm->models[local_index];
    return CSP_OK;
}

int csp_shard_models_free(csp_shard_models* sm) {
    if (!sm) return CSP_OK;
    if (!sm->ipc_opened.empty() && !sm->models.empty() && sm->models[0]) {  // (may run after csp_shutdown: no use of the clique here)
        DeviceGuard g(sm->models[0]->device);
        cudaDeviceSynchronize();
        for (void* q : sm->ipc_opened) cudaIpcCloseMemHandle(q);
    }
    for (csp_models* m : sm->models) csp_models_free(m);
    delete sm;
    return CSP_OK;
}

// ---- C2: gather of (leaf id, value) to one rank ---------------------------------------------------------------------------
// One pipeline step: queries [off, off + cnt_r) of every rank r that still has some.  Sends / receives of all local ranks are one
// NCCL group (a single thread may drive several GPUs); the root copies its own part with a device-to-device copy.
// Copy-engine variant (library-owned gather buffers): every rank pushes its block straight into the root's arrays with
// device-to-device DMA over NVLink on its communication stream -- no SMs, so it runs beside the pipeline's kernels.
static int gather_step_ce(const int32_t* const* dleaf, const double* const* dval, const int64_t* counts, long long off, long long chunk,
                          int root, long long gbase, const std::vector<cudaStream_t>& cs) {
    const size_t nl = G.local.size();
    int root_dev = -1;
    for (size_t i = 0; i < nl; ++i)
        if (G.local[i].rank == root) root_dev = G.local[i].device;
    for (size_t i = 0; i < nl; ++i) {
        const Rank& rk = G.local[i];
        const long long cnt = std::min<long long>(chunk, counts[rk.rank] - off);
        if (cnt <= 0) continue;
        long long base = gbase;
        for (int r = 0; r < rk.rank; ++r) base += counts[r];
        int32_t* dl = G.gb.leaf + base + off;
        double* dv = G.gb.val + base + off;
        DeviceGuard g(rk.device);
        if (dl != dleaf[i] + off) {
            if (root_dev >= 0 && root_dev != rk.device) CSP_CUDA(cudaMemcpyPeerAsync(dl, root_dev, dleaf[i] + off, rk.device, (size_t)cnt * 4, cs[i]));
            else CSP_CUDA(cudaMemcpyAsync(dl, dleaf[i] + off, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, cs[i]));
        }
        if (dv != dval[i] + off) {
            if (root_dev >= 0 && root_dev != rk.device) CSP_CUDA(cudaMemcpyPeerAsync(dv, root_dev, dval[i] + off, rk.device, (size_t)cnt * 8, cs[i]));
            else CSP_CUDA(cudaMemcpyAsync(dv, dval[i] + off, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, cs[i]));
        }
    }
    return CSP_OK;
}
// ... and its completion signal, once per call: every sender's communication stream, behind its copies, sends one word to the
// root, whose communication stream receives them all: when that stream is joined the gathered block is complete on the root.
static int gather_signal_ce(int root, const std::vector<cudaStream_t>& cs) {
    if (G.world <= 1) return CSP_OK;
    const size_t nl = G.local.size();
    ncclResult_t r1 = N.GroupStart();
    for (size_t i = 0; i < nl && r1 == ncclSuccess; ++i) {
        const Rank& rk = G.local[i];
        if (rk.rank != root) {
            r1 = N.Send(G.sig[i], 1, ncclInt32, root, rk.comm, cs[i]);
        } else {
            for (int r = 0; r < G.world && r1 == ncclSuccess; ++r)
                if (r != root) r1 = N.Recv(G.sig[i] + 1 + r, 1, ncclInt32, r, rk.comm, cs[i]);
        }
    }
    ncclResult_t r2 = N.GroupEnd();
    if (r1 != ncclSuccess || r2 != ncclSuccess)
        return csp_set_error(CSP_ERR_NCCL, "gather completion signal failed: %s", N.GetErrorString(r1 != ncclSuccess ? r1 : r2));
    return CSP_OK;
}
static bool gather_uses_ce(int root) { return G.gb.leaf != nullptr && G.gb.root == root && !(getenv("CSP_GATHER") && !strcmp(getenv("CSP_GATHER"), "nccl")); }

static int gather_step(const int32_t* const* dleaf, const double* const* dval, const int64_t* counts, long long off, long long chunk,
                       int root, int32_t* dleaf_all, double* dval_all, const std::vector<cudaStream_t>& cs) {
    const size_t nl = G.local.size();
    int root_local = -1;
    for (size_t i = 0; i < nl; ++i)
        if (G.local[i].rank == root) root_local = (int)i;
    bool grouped = false;
    ncclResult_t r1 = ncclSuccess;
    auto group = [&]() { if (!grouped && G.world > 1) { r1 = N.GroupStart(); grouped = true; } };
    for (size_t i = 0; i < nl && r1 == ncclSuccess; ++i) {
        const Rank& rk = G.local[i];
        const long long cnt = std::min<long long>(chunk, counts[rk.rank] - off);
        if (cnt <= 0) continue;
        if ((int)i == root_local) continue;  // handled below
        group();
        r1 = N.Send(dleaf[i] + off, (size_t)cnt, ncclInt32, root, rk.comm, cs[i]);
        if (r1 == ncclSuccess) r1 = N.Send(dval[i] + off, (size_t)cnt, ncclFloat64, root, rk.comm, cs[i]);
    }
    if (root_local >= 0 && r1 == ncclSuccess) {
        const Rank& rk = G.local[root_local];
        long long base = 0;
        for (int r = 0; r < G.world && r1 == ncclSuccess; ++r) {
            const long long cnt = std::min<long long>(chunk, counts[r] - off);
            if (cnt > 0) {
                if (r == root) {
                    DeviceGuard g(rk.device);
                    if (dleaf_all + base + off != dleaf[root_local] + off)
                        CSP_CUDA(cudaMemcpyAsync(dleaf_all + base + off, dleaf[root_local] + off, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, cs[root_local]));
                    if (dval_all + base + off != dval[root_local] + off)
                        CSP_CUDA(cudaMemcpyAsync(dval_all + base + off, dval[root_local] + off, (size_t)cnt * 8, cudaMemcpyDeviceToDevice, cs[root_local]));
                } else {
                    group();
                    r1 = N.Recv(dleaf_all + base + off, (size_t)cnt, ncclInt32, r, rk.comm, cs[root_local]);
                    if (r1 == ncclSuccess) r1 = N.Recv(dval_all + base + off, (size_t)cnt, ncclFloat64, r, rk.comm, cs[root_local]);
                }
            }
            base += counts[r];
        }
    }
    ncclResult_t r2 = grouped ? N.GroupEnd() : ncclSuccess;
    if (r1 != ncclSuccess || r2 != ncclSuccess)
        return csp_set_error(CSP_ERR_NCCL, "result gather failed: %s", N.GetErrorString(r1 != ncclSuccess ? r1 : r2));
    return CSP_OK;
}

static int check_counts(const int64_t* counts, int gather_root, const void* a, const void* b) {
    if (!counts) return csp_set_error(CSP_ERR_INVALID, "counts is NULL");
    for (int r = 0; r < G.world; ++r)
        if (counts[r] < 0) return csp_set_error(CSP_ERR_INVALID, "counts[%d] < 0", r);
    if (gather_root >= G.world) return csp_set_error(CSP_ERR_INVALID, "gather_root %d out of range", gather_root);
    if (gather_root >= 0 && !gather_uses_ce(gather_root))
        for (const Rank& rk : G.local)
            if (rk.rank == gather_root && (!a || !b)) return csp_set_error(CSP_ERR_INVALID, "the gather root needs dleaf_all and dval_all");
    return CSP_OK;
}

static int check_gather_base(const int64_t* counts, int64_t gather_base) {
    long long tot = 0;
    for (int r = 0; r < G.world; ++r) tot += counts[r];
    if (gather_base < 0 || gather_base + tot > G.gb.cap)
        return csp_set_error(CSP_ERR_INVALID, "gather block [%lld, %lld) outside the gather buffers (capacity %lld)", (long long)gather_base,
                             (long long)(gather_base + tot), G.gb.cap);
    return CSP_OK;
}

int csp_shard_gather_dev(const int32_t* const* dleaf, const double* const* dval, const int64_t* counts, int gather_root,
                         int32_t* dleaf_all, double* dval_all, int64_t gather_base, void* const* streams) {
    CSP_CHECK(need_init());
    if (!dleaf || !dval || gather_root < 0) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    CSP_CHECK(check_counts(counts, gather_root, dleaf_all, dval_all));
    const size_t nl = G.local.size();
    std::vector<cudaStream_t> s(nl);
    long long mx = 0;
    for (size_t i = 0; i < nl; ++i) s[i] = stream_of(streams, i);
    for (int r = 0; r < G.world; ++r) mx = std::max<long long>(mx, counts[r]);
    const bool ce = gather_uses_ce(gather_root);
    if (ce) CSP_CHECK(check_gather_base(counts, gather_base));
    tbegin(5, s[0]);
    int rc = CSP_OK;
    if (mx > 0) {
        if (ce) {
            rc = gather_step_ce(dleaf, dval, counts, 0, mx, gather_root, gather_base, s);
            if (rc == CSP_OK) rc = gather_signal_ce(gather_root, s);
        } else {
            rc = gather_step(dleaf, dval, counts, 0, mx, gather_root, dleaf_all, dval_all, s);
        }
    }
    tend(5, s[0]);
    return rc;
}

int csp_shard_locate_eval_dev(const csp_shard_tree* st, const csp_shard_models* sm, const double* const* dX, const int64_t* counts,
                              int32_t* const* dleaf, double* const* dval, uint8_t* const* dstatus, int gather_root,
                              int32_t* dleaf_all, double* dval_all, int64_t gather_base, int64_t chunk, void* const* streams) {
    CSP_CHECK(need_init());
    if (!st || !sm || !dX || !dleaf || !dval) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    const size_t nl = G.local.size();
    if (st->trees.size() != nl || sm->models.size() != nl) return csp_set_error(CSP_ERR_INVALID, "handles do not belong to this clique");
    CSP_CHECK(check_counts(counts, gather_root, dleaf_all, dval_all));
    const int n = st->trees[0]->v.n;
    if (chunk <= 0) chunk = 1LL << 24;
    long long mx = 0;
    for (int r = 0; r < G.world; ++r) mx = std::max<long long>(mx, counts[r]);
    std::vector<cudaStream_t> s(nl), cs(nl);
    const bool gather = gather_root >= 0 && G.world > 1;
    const bool ce = gather_root >= 0 && gather_uses_ce(gather_root);
    if (ce) CSP_CHECK(check_gather_base(counts, gather_base));
    for (size_t i = 0; i < nl; ++i) {
        s[i] = stream_of(streams, i);
        cs[i] = gather ? G.local[i].cst : s[i];
        if (gather) {  // the communication stream starts behind everything already queued on the caller's stream
            DeviceGuard g(G.local[i].device);
            CSP_CUDA(cudaEventRecord(G.local[i].fork, s[i]));
            CSP_CUDA(cudaStreamWaitEvent(cs[i], G.local[i].fork, 0));
        }
    }
    // NCCL's send / receive kernels need room beside the persistent grids of the pipeline, or they only run between kernels
    const char* renv = getenv("CSP_GATHER_SM_RESERVE");
    struct Reserve { ~Reserve() { g_csp_sm_reserve = 0; } } reserve_guard;
    g_csp_sm_reserve = gather ? (renv ? atoi(renv) : 0) : 0;  // measured at N=2 (C3): 0 / 8 / 16 SMs reserved -> 87.4 / 91.0 / 86.9 ms per step: no gain
    // Pipeline steps of `chunk` queries.  (Tapering the last chunk so that less of the final gather is exposed was measured and
    // dropped: N = 2, C3: 78.4 -> 82.2 ms per step -- chunks of a few million queries on 1e6 leaves run the binned kernels badly.)
    std::vector<long long> sizes;
    for (long long rem = mx; rem > 0;) {
        const long long c = std::min<long long>(chunk, rem);
        sizes.push_back(c);
        rem -= c;
    }
    int step = 0;
    long long off = 0;
    for (const long long chunk : sizes) {  // (shadows the parameter: this step's size)
        for (size_t i = 0; i < nl; ++i) {
            const Rank& rk = G.local[i];
            const long long cnt = std::min<long long>(chunk, counts[rk.rank] - off);
            if (cnt <= 0) continue;
            uint8_t* stp = dstatus && dstatus[i] ? dstatus[i] + off : nullptr;
            CSP_CHECK(csp_locate_eval_dev(st->trees[i], sm->models[i], dX[i] + (size_t)off * n, cnt, dleaf[i] + off, dval[i] + off, stp, s[i]));
            if (gather) {
                DeviceGuard g(rk.device);
                cudaEvent_t ev = G.local[i].chunk_ev[step & 3];
                CSP_CUDA(cudaEventRecord(ev, s[i]));
                CSP_CUDA(cudaStreamWaitEvent(cs[i], ev, 0));
            }
        }
        if (gather_root >= 0) {
            if (ce) CSP_CHECK(gather_step_ce((const int32_t* const*)dleaf, (const double* const*)dval, counts, off, chunk, gather_root, gather_base, cs));
            else CSP_CHECK(gather_step((const int32_t* const*)dleaf, (const double* const*)dval, counts, off, chunk, gather_root, dleaf_all, dval_all, cs));
        }
        off += chunk;
        ++step;
    }
    if (ce && gather) CSP_CHECK(gather_signal_ce(gather_root, cs));
    if (gather && !G.defer_join)
        for (size_t i = 0; i < nl; ++i) {  // the caller's stream ends behind the last gather step
            DeviceGuard g(G.local[i].device);
            CSP_CUDA(cudaEventRecord(G.local[i].join, cs[i]));
            CSP_CUDA(cudaStreamWaitEvent(s[i], G.local[i].join, 0));
        }
    return CSP_OK;
}

int csp_shard_defer_join(int on) {
    CSP_CHECK(need_init());
    G.defer_join = on != 0;
    return CSP_OK;
}

int csp_shard_join(void* const* streams) {
    CSP_CHECK(need_init());
    for (size_t i = 0; i < G.local.size(); ++i) {
        DeviceGuard g(G.local[i].device);
        CSP_CUDA(cudaEventRecord(G.local[i].join, G.local[i].cst));
        CSP_CUDA(cudaStreamWaitEvent(stream_of(streams, i), G.local[i].join, 0));
    }
    return CSP_OK;
}

// ---- library-owned gather buffers: the copy-engine path -------------------------------------------------------------------------
// a few bytes from the root's process to every other process: through a device word and ncclBroadcast (set-up only)
static int bcast_bytes(void* host, size_t n, int root) {
    const size_t nl = G.local.size();
    if (G.world <= 1 || nl == (size_t)G.world) return CSP_OK;  // single process: nothing to exchange
    std::vector<void*> d(nl, nullptr);
    int root_local = -1;
    for (size_t i = 0; i < nl; ++i) {
        DeviceGuard g(G.local[i].device);
        CSP_CUDA(cudaMalloc(&d[i], n));
        if (G.local[i].rank == root) { root_local = (int)i; CSP_CUDA(cudaMemcpy(d[i], host, n, cudaMemcpyHostToDevice)); }
    }
    CSP_NCCL(N.GroupStart());
    for (size_t i = 0; i < nl; ++i) CSP_NCCL(N.Broadcast(d[i], d[i], n, ncclUint8, root, G.local[i].comm, G.local[i].st));
    CSP_NCCL(N.GroupEnd());
    for (size_t i = 0; i < nl; ++i) {
        DeviceGuard g(G.local[i].device);
        CSP_CUDA(cudaStreamSynchronize(G.local[i].st));
        if (root_local < 0 && i == 0) CSP_CUDA(cudaMemcpy(host, d[i], n, cudaMemcpyDeviceToHost));
        cudaFree(d[i]);
    }
    return CSP_OK;
}

int csp_shard_gather_free(void) {
    if (!G.init || !G.gb.leaf) return CSP_OK;
    for (Rank& r : G.local) { DeviceGuard g(r.device); cudaDeviceSynchronize(); }
    if (G.gb.owner) {
        cudaFree(G.gb.leaf);
        cudaFree(G.gb.val);
    } else {
        DeviceGuard g(G.local[0].device);
        cudaIpcCloseMemHandle(G.gb.leaf);
        cudaIpcCloseMemHandle(G.gb.val);
    }
    G.gb = GatherBuf();
    return CSP_OK;
}

int csp_shard_gather_alloc(int64_t capacity, int root, int32_t** dleaf_all, double** dval_all) {
    CSP_CHECK(need_init());
    if (capacity < 1 || root < 0 || root >= G.world) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    csp_shard_gather_free();
    const size_t nl = G.local.size();
    int root_local = -1;
    for (size_t i = 0; i < nl; ++i)
        if (G.local[i].rank == root) root_local = (int)i;
    CSP_CHECK(ensure_sig());
    struct Handles { cudaIpcMemHandle_t leaf, val; int ok; } hd{};
    if (root_local >= 0) {
        DeviceGuard g(G.local[root_local].device);
        cudaError_t e = cudaMalloc((void**)&G.gb.leaf, (size_t)capacity * 4);  // plain cudaMalloc: shareable, unlike pool memory
        if (e == cudaSuccess) e = cudaMalloc((void**)&G.gb.val, (size_t)capacity * 8);
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (G.gb.leaf) cudaFree(G.gb.leaf);
            G.gb = GatherBuf();
        } else {
            G.gb.owner = true;
            hd.ok = 1;
            if (nl < (size_t)G.world) {
                if (cudaIpcGetMemHandle(&hd.leaf, G.gb.leaf) != cudaSuccess || cudaIpcGetMemHandle(&hd.val, G.gb.val) != cudaSuccess) { cudaGetLastError(); hd.ok = 2; }
            }
        }
    }
    CSP_CHECK(bcast_bytes(&hd, sizeof hd, root));
    if (root_local < 0) {
        if (hd.ok == 1) {
            DeviceGuard g(G.local[0].device);
            if (cudaIpcOpenMemHandle((void**)&G.gb.leaf, hd.leaf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle((void**)&G.gb.val, hd.val, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                G.gb = GatherBuf();
                hd.ok = 2;
            }
        }
    } else if (hd.ok == 1) {  // single process: the other local GPUs reach the root's memory through peer access
        for (size_t i = 0; i < nl; ++i) {
            if ((int)i == root_local) continue;
            DeviceGuard g(G.local[i].device);
            cudaError_t e = cudaDeviceEnablePeerAccess(G.local[root_local].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) hd.ok = 2;
            cudaGetLastError();
        }
    }
    // every rank must agree on the path: minimum over the ranks of "my mapping works"
    int ok = (hd.ok == 1 && G.gb.leaf && G.gb.val) ? 1 : 0;
    if (G.world > 1 && nl < (size_t)G.world) {
        int* d = G.sig[0] + G.world + 1;
        DeviceGuard g(G.local[0].device);
        CSP_CUDA(cudaMemcpy(d, &ok, 4, cudaMemcpyHostToDevice));
        CSP_NCCL(N.AllReduce(d, d, 1, ncclInt32, ncclMin, G.local[0].comm, G.local[0].st));
        CSP_CUDA(cudaStreamSynchronize(G.local[0].st));
        CSP_CUDA(cudaMemcpy(&ok, d, 4, cudaMemcpyDeviceToHost));
    }
    if (!ok) {
        G.gb.root = root;  // so that csp_shard_gather_free releases whatever this rank holds
        if (G.gb.leaf) csp_shard_gather_free();
        G.gb = GatherBuf();
        return csp_set_error(CSP_ERR_UNSUPPORTED, "gather buffers cannot be shared between the GPUs (peer access / CUDA IPC unavailable); "
                                                  "pass your own dleaf_all / dval_all: the NCCL send / receive path is used then");
    }
    G.gb.root = root;
    G.gb.cap = capacity;
    if (dleaf_all) *dleaf_all = root_local >= 0 ? G.gb.leaf : nullptr;
    if (dval_all) *dval_all = root_local >= 0 ? G.gb.val : nullptr;
    return CSP_OK;
}

// ---- host buffers: one pipeline (and one host thread) per local GPU ----------------------------------------------------------
int csp_shard_locate_eval(const csp_shard_tree* st, const csp_shard_models* sm, const double* X, int64_t m, int32_t* leaf, double* val,
                          uint8_t* status) {
    CSP_CHECK(need_init());
    if (!st || !sm) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    const size_t nl = G.local.size();
    if (st->trees.size() != nl || sm->models.size() != nl) return csp_set_error(CSP_ERR_INVALID, "handles do not belong to this clique");
    if (m < 0 || (m > 0 && (!X || !val))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    if (m == 0) return CSP_OK;
    const int n = st->trees[0]->v.n;
    std::vector<int> rcs(nl, CSP_OK);
    std::vector<std::string> msgs(nl);
    auto work = [&](size_t i) {
        long long lo, hi;
        shard_range(m, (int)nl, (int)i, &lo, &hi);
        if (hi <= lo) return;
        rcs[i] = csp_locate_eval(st->trees[i], sm->models[i], X + (size_t)lo * n, hi - lo, leaf ? leaf + lo : nullptr, val + lo,
                                 status ? status + lo : nullptr);
        if (rcs[i] != CSP_OK) msgs[i] = csp_last_error();  // the message is thread-local
    };
    if (nl == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < nl; ++i) th.emplace_back(work, i);
        for (auto& t : th) t.join();
    }
    for (size_t i = 0; i < nl; ++i)
        if (rcs[i] != CSP_OK) return csp_set_error(rcs[i], "GPU %d: %s", G.local[i].device, msgs[i].c_str());
    return CSP_OK;
}

}  // extern "C"
