// tree.cu -- residency of the flattened CSp-tree: validation, derived arrays, upload, blob (wire) format.
// Replaces the Box/Split/World object graph (reference src/types.jl:7-23,155-192) by SoA arrays in HBM.
#include <cstring>

#include <chrono>
#include <mutex>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

std::atomic<long long> g_csp_launches{0};
static thread_local std::string g_last_error;
static thread_local int g_device = 0;

int csp_set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
int csp_current_device() { return g_device; }
int csp_dev_alloc(void** p, size_t bytes) {
    static std::once_flag ready[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64)
        std::call_once(ready[dev], [dev] {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long thr = ~0ULL;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        });
    *p = nullptr;
    // size classes (x1.125 steps, 64 KB aligned): a tree that grows a little every optimiser iteration keeps landing in the
    // blocks its previous version released instead of making the pool map new memory
    if (bytes > 65536) {
        size_t cls = 65536;
        while (cls < bytes) cls = (cls + cls / 8 + 65535) & ~size_t(65535);
        bytes = cls;
    }
    cudaError_t e = cudaMallocAsync(p, std::max<size_t>(bytes, 16), (cudaStream_t)0);
    if (e != cudaSuccess) {  // give cached blocks back and retry once before reporting
        cudaGetLastError();
        cudaMemPool_t pool;
        cudaDeviceSynchronize();
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocAsync(p, std::max<size_t>(bytes, 16), (cudaStream_t)0);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return csp_set_error(CSP_ERR_NOMEM, "cannot allocate %zu bytes of device memory: %s", bytes, cudaGetErrorString(e));
    }
    if (!g_csp_alloc_nosync) cudaStreamSynchronize((cudaStream_t)0);  // the block may be used from any stream right away
    return CSP_OK;
}
void csp_dev_free(void* p) {
    if (p) cudaFreeAsync(p, (cudaStream_t)0);
}
int csp_require_device() {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt <= 0) {
        cudaGetLastError();
        return csp_set_error(CSP_ERR_NO_DEVICE, "no CUDA device available (%s); libcsp_b200 has no CPU fallback",
                             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return CSP_OK;
}

bool g_csp_timing = false;
thread_local int g_csp_grid_div = 1;  // per calling thread: one host thread per GPU launches its own pipeline (comm.cu)
int g_csp_sm_reserve = 0;
thread_local bool g_csp_alloc_nosync = false;
namespace {
struct TimedLaunch { int cls; cudaEvent_t a, b; };
std::vector<TimedLaunch> g_timed;
std::mutex g_timed_mu;  // held from csp_timing_begin to csp_timing_end: instrumented launches are serialised across host threads
}
void csp_timing_begin(int cls, cudaStream_t st) {
    g_timed_mu.lock();
    TimedLaunch t{cls, nullptr, nullptr};
    cudaEventCreate(&t.a);
    cudaEventCreate(&t.b);
    cudaEventRecord(t.a, st);
    g_timed.push_back(t);
}
void csp_timing_end(cudaStream_t st) {
    cudaEventRecord(g_timed.back().b, st);
    g_timed_mu.unlock();
}

extern "C" {
int csp_timing_enable(int on) {
    g_csp_timing = on != 0;
    return CSP_OK;
}
int csp_timing_read(double* ms, int64_t* launches) {
    if (!ms || !launches) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    for (int i = 0; i < CSP_K_CLASSES; ++i) { ms[i] = 0; launches[i] = 0; }
    std::lock_guard<std::mutex> lock(g_timed_mu);
    if (!g_timed.empty()) CSP_CUDA(cudaDeviceSynchronize());
    for (auto& t : g_timed) {
        float e = 0;
        if (cudaEventElapsedTime(&e, t.a, t.b) == cudaSuccess) { ms[t.cls] += e; launches[t.cls] += 1; }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    g_timed.clear();
    return CSP_OK;
}
int csp_version(void) { return 200; }
const char* csp_last_error(void) { return g_last_error.c_str(); }
int csp_device_count(int* count) {
    if (!count) return csp_set_error(CSP_ERR_INVALID, "count is NULL");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); cnt = 0; }
    *count = cnt;
    return CSP_OK;
}
int csp_set_device(int device) {
    CSP_CHECK(csp_require_device());
    int cnt = 0;
    CSP_CUDA(cudaGetDeviceCount(&cnt));
    if (device < 0 || device >= cnt) return csp_set_error(CSP_ERR_INVALID, "device %d out of range (have %d)", device, cnt);
    g_device = device;
    return CSP_OK;
}
int64_t csp_kernel_launches(void) { return g_csp_launches.load(); }
}

// ---------------------------------------------------------------------------------------------------------------
// blob format: 64-byte header followed by the desc arrays in declaration order, each padded to 16 bytes
// ---------------------------------------------------------------------------------------------------------------
static int validate_desc(const csp_tree_desc* d) {
    if (!d) return csp_set_error(CSP_ERR_INVALID, "desc is NULL");
    if (d->n < 1) return csp_set_error(CSP_ERR_INVALID, "n must be >= 1 (got %d)", d->n);
    if (d->n > CSP_MAX_DIM) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d exceeds the supported maximum %d", d->n, CSP_MAX_DIM);
    if (d->p == 1 && d->n_internal >= (1 << 24))
        return csp_set_error(CSP_ERR_UNSUPPORTED, "CS1 trees are limited to 2^24 splits by the 16-byte descent record");
    if (d->p < 1 || d->p > 8)
        return csp_set_error(CSP_ERR_UNSUPPORTED, "degree p=%d is outside 1..8 (reference src/types.jl:194-201)", d->p);
    if (d->n_nodes < 1 || d->n_internal < 0 || d->n_leaves < 1 || d->n_internal >= d->n_nodes)
        return csp_set_error(CSP_ERR_INVALID, "inconsistent counts: nodes=%d internal=%d leaves=%d", d->n_nodes, d->n_internal, d->n_leaves);
    const void* ptrs[] = {d->lower, d->upper, d->parent, d->childindex, d->internal_id, d->leaf_id, d->val};
    for (const void* q : ptrs)
        if (!q) return csp_set_error(CSP_ERR_INVALID, "a per-node array of the tree descriptor is NULL");
    if (d->n_internal > 0 && (!d->inode || !d->dims || !d->xs || !d->child || !d->pos))
        return csp_set_error(CSP_ERR_INVALID, "a per-internal-node array of the tree descriptor is NULL");
    const int C = 1 << d->p;
    if ((long long)d->n_nodes != 1 + (long long)d->n_internal * C)
        return csp_set_error(CSP_ERR_INVALID, "n_nodes=%d is not 1 + n_internal*2^p = %lld", d->n_nodes, 1 + (long long)d->n_internal * C);
    if (d->parent[0] != 0) return csp_set_error(CSP_ERR_INVALID, "node 0 must be the root (parent[0] == 0)");
    int ni = 0, nl = 0;
    for (int v = 0; v < d->n_nodes; ++v) {
        int I = d->internal_id[v];
        if (I >= 0) {
            if (I != ni || I >= d->n_internal || d->inode[I] != v)
                return csp_set_error(CSP_ERR_INVALID, "internal ids are not preorder ranks at node %d", v);
            ++ni;
            if (d->leaf_id[v] != -1) return csp_set_error(CSP_ERR_INVALID, "internal node %d has a leaf id", v);
            if (d->child[(size_t)I * C] != v + 1)
                return csp_set_error(CSP_ERR_INVALID, "node %d: child 0 must follow its parent in preorder", v);
            for (int k = 0; k < C; ++k) {
                int c = d->child[(size_t)I * C + k];
                if (c <= v || c >= d->n_nodes || d->parent[c] != v || d->childindex[c] != k)
                    return csp_set_error(CSP_ERR_INVALID, "node %d: child %d (%d) is inconsistent with parent/childindex", v, k, c);
            }
            for (int k = 0; k < d->p; ++k) {
                int dim = d->dims[(size_t)I * d->p + k];
                int maxdim = (d->n + d->p - 1) / d->p * d->p;
                if (dim < 1 || dim > maxdim) return csp_set_error(CSP_ERR_INVALID, "node %d: split dimension %d out of range", v, dim);
            }
        } else {
            int L = d->leaf_id[v];
            if (L >= 0) {
                if (L != nl) return csp_set_error(CSP_ERR_INVALID, "leaf ids are not preorder ranks at node %d", v);
                ++nl;
            }
        }
    }
    if (ni != d->n_internal || nl != d->n_leaves)
        return csp_set_error(CSP_ERR_INVALID, "counted %d internal / %d leaves, descriptor says %d / %d", ni, nl, d->n_internal, d->n_leaves);
    return CSP_OK;
}

// tree_build.cu

template <class T>
static int to_device(csp_tree* t, const std::vector<T>& h, const T** out) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(h.size() * sizeof(T), 16);
    CSP_CHECK(csp_dev_alloc(&p, bytes));
    t->allocs.push_back(p);
    t->bytes += bytes;
    if (!h.empty()) CSP_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return CSP_OK;
}

// ancestor table for the leaf fits (fit.cu): one thread per leaf climbs to the root once per tree version
__global__ void __launch_bounds__(256) k_build_paths(const int4* __restrict__ ninfo, const int* __restrict__ lnode, const long long* __restrict__ lpoff,
                                                     int2* __restrict__ lpath, int n_leaves) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_leaves; l += gridDim.x * blockDim.x) {
        int v = lnode[l];
        int4 ci = ninfo[v];
        long long o = lpoff[l];
        while (ci.x != v) {
            v = ci.x;
            ci = ninfo[v];
            lpath[o++] = make_int2(ci.y, ci.z);
        }
    }
}

static int upload_impl(const csp_tree_desc* d, csp_tree* t) {
    const bool uprof = getenv("CSP_UPLOAD_PROF") && atoi(getenv("CSP_UPLOAD_PROF"));
    auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tp0 = tnow();
    auto lap = [&](const char* what) { if (uprof) { const double t1 = tnow(); fprintf(stderr, "[csp upload prof] %-22s %8.3f ms\n", what, t1 - tp0); tp0 = t1; } };
    const int n = d->n, p = d->p, C = 1 << p, nn = d->n_nodes, ni = d->n_internal, nl = d->n_leaves;
    // derived per-node arrays
    std::vector<int4> ninfo(nn);
    std::vector<int> icount(nn), depth(nn, 0), nleaf(d->leaf_id, d->leaf_id + nn), lnode(nl);
    for (int v = 0; v < nn; ++v) icount[v] = d->internal_id[v] >= 0 ? 1 : 0;
    for (int v = nn - 1; v >= 1; --v) icount[d->parent[v]] += icount[v];
    int before = 0, maxdepth = 0;
    for (int v = 0; v < nn; ++v) {
        bool internal = d->internal_id[v] >= 0;
        int flags = d->childindex[v];
        if (!internal) flags |= CSP_NF_LEAF;
        if (!internal && d->leaf_id[v] < 0) flags |= CSP_NF_FAKE;
        ninfo[v] = make_int4(d->parent[v], before, icount[v], flags);
        if (internal) ++before;
        if (v > 0) depth[v] = depth[d->parent[v]] + 1;
        if (d->leaf_id[v] >= 0) { lnode[d->leaf_id[v]] = v; maxdepth = std::max(maxdepth, depth[v]); }
    }
    lap("node info");
    // derived per-internal arrays (preorder ids)
    std::vector<double> ixself((size_t)ni * p), icval((size_t)ni * C);
    for (int I = 0; I < ni; ++I) {
        for (int k = 0; k < p; ++k) {
            int dim = d->dims[(size_t)I * p + k];
            ixself[(size_t)I * p + k] = dim <= n ? d->pos[(size_t)I * n + dim - 1] : 0.0;
        }
        for (int k = 0; k < C; ++k) icval[(size_t)I * C + k] = d->val[d->child[(size_t)I * C + k]];
    }
    lap("xself / child values");
    // descent records in BFS order (see common.cuh)
    std::vector<DRec1> r1(p == 1 ? ni : 0);
    std::vector<DRec2> r2(p == 2 ? ni : 0);
    std::vector<int> dleaf;
    dleaf.reserve((size_t)ni * (C - 1) + 1);
    if (ni == 0 || p > 2) dleaf.push_back(0);  // the root is the only leaf / no packed records for p > 2 (generic kernels, extras.cu)
    if (p <= 2) {
        std::vector<int> bfs;  // preorder internal ids in BFS order
        bfs.reserve(ni);
        if (ni > 0) bfs.push_back(0);
        for (size_t head = 0; head < bfs.size(); ++head) {
            const int I = bfs[head];
            const int F = (int)bfs.size(), G = (int)dleaf.size();
            unsigned mask = 0;
            int L[4] = {-1, -1, -1, -1};
            for (int k = 0; k < C; ++k) {
                const int c = d->child[(size_t)I * C + k];
                if (d->internal_id[c] >= 0) { mask |= 1u << k; bfs.push_back(d->internal_id[c]); }
                else { dleaf.push_back(d->leaf_id[c]); L[k] = d->leaf_id[c]; }
            }
            double mid[2] = {0, 0};
            unsigned dim0[2] = {0, 0}, flip[2] = {0, 0}, fict[2] = {1, 1};
            for (int k = 0; k < p; ++k) {
                const int dim = d->dims[(size_t)I * p + k];
                const double xo = d->xs[(size_t)I * p + k], xself = ixself[(size_t)I * p + k];
                if (dim <= n) {
                    mid[k] = (xself + xo) / 2;  // src/tree.jl:383, the same IEEE operations as the reference
                    dim0[k] = (unsigned)(dim - 1); flip[k] = xo < xself ? 1u : 0u; fict[k] = 0;
                }
            }
            if (p == 2) {
                DRec2 r{};
                r.mid0 = mid[0]; r.mid1 = mid[1]; r.F = F;
                csp_encode_leaf_runs(L, mask, fict[0], fict[1], G, &r.G, &r.H);
                r.meta = dim0[0] | (dim0[1] << 12) | (flip[0] << 24) | (flip[1] << 25) | (fict[0] << 26) | (fict[1] << 27) | (mask << 28);
                r2[head] = r;
            } else {
                DRec1 r{};
                r.mid = mid[0];
                r.w0 = (unsigned)F | (((unsigned)G & 0xffu) << 24);
                r.w1 = ((unsigned)G >> 8) | (dim0[0] << 16) | (flip[0] << 28) | (mask << 29);
                r1[head] = r;
            }
        }
    }
    lap("descent records");
    t->h_lower.assign(d->lower, d->lower + n);
    t->h_upper.assign(d->upper, d->upper + n);
    TreeView& v = t->v;
    v.n = n; v.p = p; v.n_nodes = nn; v.n_internal = ni; v.n_leaves = nl;
    CSP_CHECK(to_device(t, std::vector<double>(d->lower, d->lower + n), &v.lower));
    CSP_CHECK(to_device(t, std::vector<double>(d->upper, d->upper + n), &v.upper));
    if (p == 1) { const DRec1* q; CSP_CHECK(to_device(t, r1, &q)); v.drec = q; }
    else if (p == 2) { const DRec2* q; CSP_CHECK(to_device(t, r2, &q)); v.drec = q; }
    else v.drec = nullptr;
    CSP_CHECK(to_device(t, dleaf, &v.dleaf));
    CSP_CHECK(to_device(t, ninfo, &v.ninfo));
    CSP_CHECK(to_device(t, nleaf, &v.nleaf));
    CSP_CHECK(to_device(t, std::vector<int>(d->inode, d->inode + ni), &v.inode));
    CSP_CHECK(to_device(t, std::vector<int>(d->dims, d->dims + (size_t)ni * p), &v.idims));
    CSP_CHECK(to_device(t, std::vector<double>(d->xs, d->xs + (size_t)ni * p), &v.ixs));
    CSP_CHECK(to_device(t, ixself, &v.ixself));
    CSP_CHECK(to_device(t, icval, &v.icval));
    CSP_CHECK(to_device(t, std::vector<int>(d->child, d->child + (size_t)ni * C), &v.ichild));
    CSP_CHECK(to_device(t, std::vector<double>(d->pos, d->pos + (size_t)ni * n), &v.ipos));
    CSP_CHECK(to_device(t, lnode, &v.lnode));
    lap("uploads");
    {
        std::vector<long long> lpoff((size_t)nl + 1, 0);
        for (int l = 0; l < nl; ++l) lpoff[l + 1] = lpoff[l] + depth[lnode[l]];
        CSP_CHECK(to_device(t, lpoff, &v.lpoff));
        void* lp = nullptr;
        const size_t bytes = std::max<size_t>((size_t)lpoff[nl] * sizeof(int2), 16);
        CSP_CHECK(csp_dev_alloc(&lp, bytes));
        t->allocs.push_back(lp);
        t->bytes += bytes;
        v.lpath = (const int2*)lp;
        if (nl > 0) {
            k_build_paths<<<std::min(4096, (nl + 255) / 256), 256>>>(v.ninfo, v.lnode, v.lpoff, (int2*)lp, nl);
            CSP_CUDA(cudaGetLastError());
        }
    }
    lap("ancestor paths");
    {   // fit work units: the real leaves that are children of the same split share their splits() traversal
        std::vector<int4> units;
        units.reserve((size_t)ni);
        for (int I = 0; I < ni && p == 2; ++I) {
            int m[4] = {-1, -1, -1, -1}, cnt = 0;
            for (int k = 0; k < C && cnt < 4; ++k) {
                const int c = d->child[(size_t)I * C + k];
                if (d->internal_id[c] < 0 && d->leaf_id[c] >= 0) m[cnt++] = d->leaf_id[c];
            }
            if (cnt) units.push_back(make_int4(m[0], m[1], m[2], m[3]));
        }
        if (ni == 0 && nl == 1 && p == 2) units.push_back(make_int4(0, -1, -1, -1));  // the root is the only leaf
        CSP_CHECK(to_device(t, units, &t->fit_units));
        t->n_fit_units = (long long)units.size();
    }
    lap("fit units");
    t->max_depth = maxdepth;
    return CSP_OK;
}

// Host-side construction of the derived arrays (upload_impl above): kept for A/B checks of the device builder only
// (CSP_UPLOAD_HOST=1; tests/test_gpu_parity.py compares the two bit for bit).
static bool host_builder() { return getenv("CSP_UPLOAD_HOST") && atoi(getenv("CSP_UPLOAD_HOST")); }

static csp_tree* new_tree() {
    csp_tree* t = new csp_tree();
    t->device = csp_current_device();
    int sms = 0;  // (cudaGetDeviceProperties costs milliseconds per call)
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device) == cudaSuccess) t->n_sm = sms;
    return t;
}

static int check_header(int n, int p, int nn, int ni, int nl) {
    if (n < 1) return csp_set_error(CSP_ERR_INVALID, "n must be >= 1 (got %d)", n);
    if (n > CSP_MAX_DIM) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d exceeds the supported maximum %d", n, CSP_MAX_DIM);
    if (p < 1 || p > 8) return csp_set_error(CSP_ERR_UNSUPPORTED, "degree p=%d is outside 1..8 (reference src/types.jl:194-201)", p);
    if (p == 1 && ni >= (1 << 24)) return csp_set_error(CSP_ERR_UNSUPPORTED, "CS1 trees are limited to 2^24 splits by the 16-byte descent record");
    if (nn < 1 || ni < 0 || nl < 1 || ni >= nn) return csp_set_error(CSP_ERR_INVALID, "inconsistent counts: nodes=%d internal=%d leaves=%d", nn, ni, nl);
    if ((long long)nn != 1 + (long long)ni * (1 << p))
        return csp_set_error(CSP_ERR_INVALID, "n_nodes=%d is not 1 + n_internal*2^p = %lld", nn, 1 + (long long)ni * (1 << p));
    return CSP_OK;
}

// The tree owns one device buffer in the blob layout; `fill` puts the description into it; the rest is derived on the device.
template <class Fill>
static int upload_device(int n, int p, int nn, int ni, int nl, const double* h_lower, const double* h_upper, Fill&& fill, csp_tree** out) {
    CSP_CHECK(check_header(n, p, nn, ni, nl));
    csp_tree* t = new_tree();
    DeviceGuard g(t->device);
    Section s[12];
    const size_t need = blob_layout(n, p, nn, ni, s);
    void* dblob = nullptr;
    int rc = csp_dev_alloc(&dblob, need);
    if (rc == CSP_OK) {
        t->allocs.push_back(dblob);
        t->bytes += need;
        rc = fill((char*)dblob, s, need);
    }
    if (rc == CSP_OK) {
        rc = csp_tree_build_device(t, raw_tree_at((const char*)dblob, s, n, p, nn, ni, nl), h_lower, h_upper);
    }
    if (rc != CSP_OK) { csp_tree_free(t); return rc; }
    *out = t;
    return CSP_OK;
}

extern "C" {

int csp_tree_free(csp_tree* tree) {
    if (!tree) return CSP_OK;
    {
        DeviceGuard g(tree->device);
        cudaDeviceSynchronize();  // work on any stream may still read the tree
        for (void* p : tree->allocs) csp_dev_free(p);
        csp_dev_free(tree->fit_ws);
        csp_dev_free(tree->fit_coef);
        if (tree->live) csp_live_free(tree->live);
        for (int i = 0; i < 2; ++i) {
            if (tree->pipe_st[i]) cudaStreamDestroy(tree->pipe_st[i]);
            if (tree->pipe_join[i]) cudaEventDestroy(tree->pipe_join[i]);
        }
        if (tree->pipe_fork) cudaEventDestroy(tree->pipe_fork);
    }
    delete tree;
    return CSP_OK;
}

int csp_tree_upload(const csp_tree_desc* desc, csp_tree** out) {
    if (!out) return csp_set_error(CSP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!desc) return csp_set_error(CSP_ERR_INVALID, "desc is NULL");
    CSP_CHECK(csp_require_device());
    if (host_builder()) {
        CSP_CHECK(validate_desc(desc));
        csp_tree* t = new_tree();
        DeviceGuard g(t->device);
        int rc = upload_impl(desc, t);
        if (rc == CSP_OK) rc = csp_fit_build_coef(t);
        if (rc == CSP_OK && cudaDeviceSynchronize() != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc != CSP_OK) { csp_tree_free(t); return rc; }
        *out = t;
        return CSP_OK;
    }
    const csp_tree_desc* d = desc;
    const void* ptrs[] = {d->lower, d->upper, d->parent, d->childindex, d->internal_id, d->leaf_id, d->val};
    for (const void* q : ptrs)
        if (!q) return csp_set_error(CSP_ERR_INVALID, "a per-node array of the tree descriptor is NULL");
    if (d->n_internal > 0 && (!d->inode || !d->dims || !d->xs || !d->child || !d->pos))
        return csp_set_error(CSP_ERR_INVALID, "a per-internal-node array of the tree descriptor is NULL");
    return upload_device(d->n, d->p, d->n_nodes, d->n_internal, d->n_leaves, d->lower, d->upper, [&](char* dblob, const Section* s, size_t) -> int {
        const void* src[12] = {d->lower, d->upper, d->parent, d->childindex, d->internal_id, d->leaf_id, d->val,
                               d->inode, d->dims, d->xs, d->child, d->pos};
        for (int i = 0; i < 12; ++i)
            if (s[i].bytes) CSP_CUDA(cudaMemcpyAsync(dblob + s[i].off, src[i], s[i].bytes, cudaMemcpyHostToDevice, 0));
        return CSP_OK;
    }, out);
}

int csp_tree_blob_size(const csp_tree_desc* d, size_t* bytes) {
    if (!d || !bytes) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    Section s[12];
    *bytes = blob_layout(d->n, d->p, d->n_nodes, d->n_internal, s);
    return CSP_OK;
}

int csp_tree_pack(const csp_tree_desc* d, void* blob, size_t bytes) {
    if (!blob) return csp_set_error(CSP_ERR_INVALID, "blob is NULL");
    CSP_CHECK(validate_desc(d));
    Section s[12];
    size_t need = blob_layout(d->n, d->p, d->n_nodes, d->n_internal, s);
    if (bytes < need) return csp_set_error(CSP_ERR_INVALID, "blob buffer too small: %zu < %zu", bytes, need);
    memset(blob, 0, need);
    BlobHeader h{};
    memcpy(h.magic, "CSPB200", 8);
    h.version = 1; h.n = d->n; h.p = d->p; h.n_nodes = d->n_nodes; h.n_internal = d->n_internal; h.n_leaves = d->n_leaves;
    memcpy(blob, &h, sizeof h);
    const void* src[12] = {d->lower, d->upper, d->parent, d->childindex, d->internal_id, d->leaf_id, d->val,
                           d->inode, d->dims, d->xs, d->child, d->pos};
    for (int i = 0; i < 12; ++i)
        if (s[i].bytes) memcpy((char*)blob + s[i].off, src[i], s[i].bytes);
    return CSP_OK;
}

int csp_tree_upload_blob(const void* blob, size_t bytes, csp_tree** out) {
    if (!blob || !out) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (bytes < sizeof(BlobHeader)) return csp_set_error(CSP_ERR_INVALID, "blob shorter than its header");
    CSP_CHECK(csp_require_device());
    cudaPointerAttributes attr{};
    const bool on_device = cudaPointerGetAttributes(&attr, blob) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    BlobHeader h;
    std::vector<char> bounce;
    if (on_device) {  // e.g. the buffer an NCCL broadcast landed in
        DeviceGuard gsrc(attr.device);  // (pool memory is only accessible from its own device)
        if (attr.device == csp_current_device()) {
            CSP_CUDA(cudaMemcpy(&h, blob, sizeof h, cudaMemcpyDeviceToHost));
        } else {  // a blob on another GPU than the target: through the host
            bounce.resize(bytes);
            CSP_CUDA(cudaMemcpy(bounce.data(), blob, bytes, cudaMemcpyDeviceToHost));
            return csp_tree_upload_blob(bounce.data(), bytes, out);
        }
    } else {
        memcpy(&h, blob, sizeof h);
    }
    if (memcmp(h.magic, "CSPB200", 8) != 0 || h.version != 1) return csp_set_error(CSP_ERR_INVALID, "not a CSPB200 v1 blob");
    CSP_CHECK(check_header(h.n, h.p, h.n_nodes, h.n_internal, h.n_leaves));
    Section s[12];
    const size_t need = blob_layout(h.n, h.p, h.n_nodes, h.n_internal, s);
    if (bytes < need) return csp_set_error(CSP_ERR_INVALID, "blob truncated: %zu < %zu", bytes, need);
    if (host_builder()) {
        std::vector<char> host;
        const char* base = (const char*)blob;
        if (on_device) {
            host.resize(need);
            CSP_CUDA(cudaMemcpy(host.data(), blob, need, cudaMemcpyDeviceToHost));
            base = host.data();
        }
        csp_tree_desc d{};
        d.n = h.n; d.p = h.p; d.n_nodes = h.n_nodes; d.n_internal = h.n_internal; d.n_leaves = h.n_leaves;
        d.lower = (const double*)(base + s[0].off); d.upper = (const double*)(base + s[1].off);
        d.parent = (const int32_t*)(base + s[2].off); d.childindex = (const uint8_t*)(base + s[3].off);
        d.internal_id = (const int32_t*)(base + s[4].off); d.leaf_id = (const int32_t*)(base + s[5].off);
        d.val = (const double*)(base + s[6].off); d.inode = (const int32_t*)(base + s[7].off);
        d.dims = (const int32_t*)(base + s[8].off); d.xs = (const double*)(base + s[9].off);
        d.child = (const int32_t*)(base + s[10].off); d.pos = (const double*)(base + s[11].off);
        return csp_tree_upload(&d, out);
    }
    // one copy of the blob (device-to-device when it already is on a GPU: no host bounce), everything else derived on the device
    const double* hl = on_device ? nullptr : (const double*)((const char*)blob + s[0].off);
    const double* hu = on_device ? nullptr : (const double*)((const char*)blob + s[1].off);
    return upload_device(h.n, h.p, h.n_nodes, h.n_internal, h.n_leaves, hl, hu, [&](char* dblob, const Section*, size_t nbytes) -> int {
        CSP_CUDA(cudaMemcpyAsync(dblob, blob, nbytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, 0));
        return CSP_OK;
    }, out);
}

// Diagnostic: a derived device array back on the host, so that tests can compare the device builder with the host builder.
// which: 0 ninfo, 1 drec, 2 dleaf, 3 ixself, 4 icval, 5 lnode, 6 lpoff, 7 lpath, 8 fit_units.  *bytes = size of the array.
int csp_tree_export_derived(const csp_tree* t, int which, void* out, size_t cap, size_t* bytes) {
    if (!t || !bytes) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    const TreeView& v = t->v;
    const size_t C = size_t(1) << v.p, ni = v.n_internal, nl = v.n_leaves;
    long long ptotal = 0;
    DeviceGuard g(t->device);
    CSP_CUDA(cudaMemcpy(&ptotal, v.lpoff + nl, 8, cudaMemcpyDeviceToHost));
    const void* src[9] = {v.ninfo, v.drec, v.dleaf, v.ixself, v.icval, v.lnode, v.lpoff, v.lpath, t->fit_units};
    const size_t sz[9] = {(size_t)v.n_nodes * 16, v.p == 2 ? ni * sizeof(DRec2) : (v.p == 1 ? ni * sizeof(DRec1) : 0), (v.p <= 2 ? ni * (C - 1) + 1 : 1) * 4,
                          ni * v.p * 8, ni * C * 8, nl * 4, (nl + 1) * 8, (size_t)ptotal * 8, (size_t)t->n_fit_units * 16};
    if (which < 0 || which > 8) return csp_set_error(CSP_ERR_INVALID, "which out of range");
    *bytes = sz[which];
    if (!out) return CSP_OK;
    if (cap < sz[which]) return csp_set_error(CSP_ERR_INVALID, "buffer too small: %zu < %zu", cap, sz[which]);
    if (sz[which]) CSP_CUDA(cudaMemcpy(out, src[which], sz[which], cudaMemcpyDeviceToHost));
    return CSP_OK;
}

int csp_tree_info(const csp_tree* t, int64_t info[8]) {
    if (!t || !info) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    info[0] = t->v.n; info[1] = t->v.p; info[2] = t->v.n_nodes; info[3] = t->v.n_internal; info[4] = t->v.n_leaves;
    info[5] = t->device; info[6] = t->max_depth; info[7] = (int64_t)(t->bytes >> 10);
    return CSP_OK;
}

}  // extern "C"
