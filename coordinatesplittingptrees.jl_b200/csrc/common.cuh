// common.cuh -- shared definitions of libcsp_b200.so (sm_100a only; no CPU fallback, no multi-backend dispatch).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/csp_b200.h"

// ---------------------------------------------------------------------------------------------------------------
// Device-resident flattened tree (SoA).  Ids: node = preorder rank, internal = preorder rank among non-leaves,
// leaf = rank in leaves(root).  See DESIGN.md "Data layout in HBM".
// ---------------------------------------------------------------------------------------------------------------

// Descent records.  The descent kernel uses its own numbering: internal nodes in BFS order ("descent ids"), so that the
// top levels of the tree are the contiguous prefix [0, T) of the record array (staged in shared memory), and the
// internal children of a node -- as well as its leaf children, numbered by a parallel BFS "descent leaf id" -- are
// consecutive.  A record therefore needs only the first internal child F, the first leaf child G and a 4-bit mask.
// Decision rule of reference src/tree.jl:383:  bit_j = (x[sd] >= (xself+xother)/2) xor (xother < xself); fictive dims -> 0.
//
// degree 2: 32 bytes, one 256-bit load (LDG.E.256).  meta: d0 bits 0-11, d1 bits 12-23 (0-based dims), flip0 24, flip1 25,
// fictive0 26, fictive1 27, internal-children mask bits 28-31.
// G, H: the leaf ids of the node's leaf children without a table.  In preorder the leaf children of a node have consecutive
// leaf ids except across an internal sibling, i.e. at most two runs among four children: leaf child k has the id
// (an internal child lies between the first leaf child and k ? H : G) + (number of leaf children below k).  Nodes where this
// does not hold for every reachable child (fake leaves between real ones: fictive first dimension) store G = ~(first descent
// leaf id) < 0 and resolve the id through dleaf.
struct __align__(32) DRec2 {
    double mid0, mid1;
    int F, G;
    unsigned meta;
    int H;
};
// Encodes (G, H) of a degree-2 record: L[k] = leaf id of child k (< 0: fake leaf), mask = internal children, fict = fictive
// flags of the two split dimensions, gdl = the node's first descent leaf id.  Used by both tree builders.
__host__ __device__ inline void csp_encode_leaf_runs(const int L[4], unsigned mask, unsigned fict0, unsigned fict1, int gdl, int* G, int* H) {
    const unsigned lead = mask & ~(mask + 1u) & 0xfu;
    int g = -1, h = -1;
    bool haveg = false, haveh = false, ok = true;
    for (int k = 0; k < 4; ++k) {
        if ((mask >> k) & 1u) continue;
        if ((fict0 && (k & 1)) || (fict1 && (k & 2))) continue;  // never reached by a descent
        const unsigned below = (1u << k) - 1u;
        int cnt = 0;
        for (int j = 0; j < k; ++j) cnt += !((mask >> j) & 1u);
        const int base = L[k] - cnt;
        if (L[k] < 0 || base < 0) { ok = false; break; }
        if (mask & below & ~lead) {
            if (haveh && h != base) ok = false;
            h = base; haveh = true;
        } else {
            if (haveg && g != base) ok = false;
            g = base; haveg = true;
        }
    }
    if (ok) { *G = haveg ? g : 0; *H = haveh ? h : 0; }
    else { *G = ~gdl; *H = 0; }
}
// degree 1: 16 bytes, one 128-bit load.  w0 = F (24 bits) | G low 8 bits << 24;  w1 = G >> 8 (16 bits) | d << 16 (12 bits)
// | flip << 28 | internal-children mask << 29 (2 bits).  Limits CS1 trees to 2^24 internal nodes / leaves.
struct __align__(16) DRec1 {
    double mid;
    unsigned w0, w1;
};
#define CSP_MAX_DIM 4096

// per-node record: x = parent node id, y = istart (number of internal nodes before this node in preorder; the node's
// own internal id when it is internal), z = icount (internal nodes in the subtree, itself included), w = flags
#define CSP_NF_CHILDINDEX(w) ((w)&0xff)
#define CSP_NF_LEAF 0x100
#define CSP_NF_FAKE 0x200

struct TreeView {
    int n, p, n_nodes, n_internal, n_leaves;
    const double* lower;   // [n]
    const double* upper;   // [n]
    const void* drec;      // [n_internal] DRec1 / DRec2, BFS order
    const int* dleaf;      // [n_internal*(2^p-1)+1] descent leaf id -> leaf id (-1 for fake leaves)
    const int4* ninfo;     // [n_nodes]
    const int* nleaf;      // [n_nodes] leaf id or -1
    const int* inode;      // [n_internal] node id
    const int* idims;      // [n_internal*p]   1-based, > n fictive
    const double* ixs;     // [n_internal*p]   split.xs
    const double* ixself;  // [n_internal*p]   position(split.self, dims[k])
    const double* icval;   // [n_internal*2^p] value(getchild(split, k))
    const int* ichild;     // [n_internal*2^p] node ids of the children
    const double* ipos;    // [n_internal*n]   position(box)
    const int* lnode;      // [n_leaves] node id of each leaf
    const long long* lpoff;  // [n_leaves+1] offsets into lpath (depth prefix sums)
    const int2* lpath;     // per leaf, bottom-up: (first internal id, internal count) of every ancestor's subtree
};

// Per-leaf canonical models, one record of `stride` doubles per leaf.  With z = (x - b, 1) the canonical model
// c + g'dx + dx'Q dx/2 (reference src/cs2.jl:155-159) is the single quadratic form  sum_{i<=j} A[i][j] z_i z_j  with
//   A[i][j] = Q[i][j] (i < j < n),   A[i][i] = Q[i][i]/2,   A[i][n] = g[i],   A[n][n] = c,
// so every stored coefficient is used the same way (no per-element roles in the evaluation kernels):
//   degree 2: [ b[n] | A, packed upper triangle of the (n+1)x(n+1) matrix, row-major | pad ]
//   degree 1: [ b[n] | g[n] | c | pad ]                          (the last column of A only)
struct ModelsView {
    int n, degree, stride;
    long long n_models;
    const double* data;
};
static inline long long csp_model_len(int n, int degree) {
    return n + (degree == 2 ? (long long)(n + 1) * (n + 2) / 2 : (long long)n + 1);
}
static inline int csp_model_stride(int n, int degree) {
    return (int)((csp_model_len(n, degree) + 3) / 4 * 4);  // whole 32-B sectors per record
}
// offset of A[i][j] (i <= j <= n) inside the packed triangle
__host__ __device__ static inline long long csp_a_index(int n, int i, int j) {
    return (long long)i * (n + 1) - (long long)i * (i - 1) / 2 + (j - i);
}

struct LiveState;  // live.cu
struct csp_tree {
    int device = 0;
    TreeView v{};
    int max_depth = 0;
    size_t bytes = 0;
    std::vector<void*> allocs;
    std::vector<double> h_lower, h_upper;  // host copies of the world bounds (kernel parameters)
    int n_sm = 0;
    // launch geometry cache (kernel function -> {tile, stages, top, smem, resident CTAs per SM}); one caller thread per handle
    struct Geo { int tile, stages, top; size_t smem; int occ; };
    mutable std::vector<std::pair<const void*, Geo>> geo;
    // internal streams of the chunked binned pipeline: consecutive chunks alternate between them so that the
    // L1-bound descent of one chunk overlaps the DRAM-bound evaluation of the previous one
    const int4* fit_units = nullptr;  // [n_fit_units] leaf ids (-1 padded) of the real leaves under each split that has any (fit.cu)
    long long n_fit_units = 0;
    double* fit_coef = nullptr;  // [n_internal] highest-order coefficient of every split (built with the tree: csp_fit_build_coef)
    mutable void* fit_ws = nullptr;  // grow-only scratch of the fit pre-pass (chain results), one fit call at a time per handle
    mutable size_t fit_ws_bytes = 0;
    LiveState* live = nullptr;  // set for trees grown by appending to the split log (live.cu)
    mutable cudaStream_t pipe_st[2] = {nullptr, nullptr};
    mutable cudaEvent_t pipe_fork = nullptr, pipe_join[2] = {nullptr, nullptr};
};
struct csp_models {
    int device = 0;
    int n = 0, degree = 2, stride = 0;
    long long n_models = 0;
    double* data = nullptr;    // device, n_models*stride
    uint8_t* status = nullptr; // device, n_models (CSP_FIT_*)
    bool plain = false;        // data / status are plain cudaMalloc memory (shareable between GPUs / processes), not pool memory
    // Set by the multi-GPU layer when other ranks deposit parts of the records asynchronously (copy-engine all-gather, comm.cu):
    // recorded once everything has landed; every consumer of the records waits for it on its own stream first.
    cudaEvent_t ready = nullptr;
};
inline void csp_models_wait_ready(const csp_models* mo, cudaStream_t st) {
    if (mo && mo->ready) cudaStreamWaitEvent(st, mo->ready, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// wire format (csp_tree_pack / csp_tree_upload_blob): a 64-byte header, then the csp_tree_desc arrays in declaration order,
// each padded to 16 bytes.  The device copy of a tree keeps this layout; RawTree points into it.
// ---------------------------------------------------------------------------------------------------------------
struct BlobHeader {
    char magic[8];  // "CSPB200\0"
    int32_t version, n, p, n_nodes, n_internal, n_leaves;
    int32_t reserved[8];
};
static_assert(sizeof(BlobHeader) == 64, "blob header is 64 bytes");

struct Section { size_t off, bytes; };
inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
inline size_t blob_layout(int n, int p, int nn, int ni, Section s[12]) {
    size_t C = size_t(1) << p;
    size_t sz[12] = {size_t(n) * 8, size_t(n) * 8, size_t(nn) * 4, size_t(nn), size_t(nn) * 4, size_t(nn) * 4, size_t(nn) * 8,
                     size_t(ni) * 4, size_t(ni) * p * 4, size_t(ni) * p * 8, size_t(ni) * C * 4, size_t(ni) * n * 8};
    size_t off = sizeof(BlobHeader);
    for (int i = 0; i < 12; ++i) { s[i] = {off, sz[i]}; off = align16(off + sz[i]); }
    return off;
}

struct RawTree {  // device pointers to the description arrays (sections of the blob)
    int n, p, nn, ni, nl;
    const double *lower, *upper;
    const int* parent;
    const uint8_t* childindex;
    const int* internal_id;
    const int* leaf_id;
    const double* val;
    const int* inode;
    const int* dims;
    const double* xs;
    const int* child;
    const double* pos;
};
inline RawTree raw_tree_at(const char* b, const Section* s, int n, int p, int nn, int ni, int nl) {
    return RawTree{n, p, nn, ni, nl,
                   (const double*)(b + s[0].off), (const double*)(b + s[1].off), (const int*)(b + s[2].off), (const uint8_t*)(b + s[3].off),
                   (const int*)(b + s[4].off), (const int*)(b + s[5].off), (const double*)(b + s[6].off), (const int*)(b + s[7].off),
                   (const int*)(b + s[8].off), (const double*)(b + s[9].off), (const int*)(b + s[10].off), (const double*)(b + s[11].off)};
}
struct csp_tree;
int csp_tree_build_device(csp_tree* t, const RawTree& r, const double* h_lower = nullptr, const double* h_upper = nullptr);  // tree_build.cu
int csp_fit_build_coef(csp_tree* t);                       // fit.cu
struct csp_models;
extern "C" int csp_models_alloc_ex(const csp_tree* t, bool plain, csp_models** out);  // fit.cu (internal)
struct LiveState;                                          // live.cu
void csp_live_free(LiveState* ls);

// ---------------------------------------------------------------------------------------------------------------
// error plumbing / launch accounting
// ---------------------------------------------------------------------------------------------------------------
int csp_set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_csp_launches;
int csp_current_device();
int csp_require_device();  // CSP_OK or CSP_ERR_NO_DEVICE
// Device memory for handles and temporaries: stream-ordered allocations (legacy default stream) from the device's default
// pool, whose release threshold is raised once so that freed blocks are reused -- re-uploading a tree or re-fitting models
// every optimiser iteration then costs no cudaMalloc / cudaFree.  Callers synchronise before freeing memory other streams used.
int csp_dev_alloc(void** p, size_t bytes);
// While set (by a builder whose work is all on the legacy default stream and which synchronises at its end), csp_dev_alloc skips
// its per-allocation stream synchronisation.
extern thread_local bool g_csp_alloc_nosync;
struct AllocNoSync {
    bool prev;
    AllocNoSync() : prev(g_csp_alloc_nosync) { g_csp_alloc_nosync = true; }
    ~AllocNoSync() { g_csp_alloc_nosync = prev; }
};
void csp_dev_free(void* p);

#define CSP_CUDA(call)                                                                                         \
    do {                                                                                                       \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
            return csp_set_error(CSP_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)
#define CSP_CHECK(call)            \
    do {                           \
        int _rc = (call);          \
        if (_rc != CSP_OK) return _rc; \
    } while (0)
#define CSP_LAUNCHED() (g_csp_launches.fetch_add(1, std::memory_order_relaxed))

// Optional per-kernel timing with CUDA events on the launching stream (csp_timing_enable / csp_timing_read).
enum { CSP_K_LOCATE = 0, CSP_K_LOCATE_EVAL, CSP_K_SCAN, CSP_K_SCATTER, CSP_K_EVAL_BINNED, CSP_K_EVAL, CSP_K_FIT, CSP_K_CLASSES };
extern bool g_csp_timing;
extern thread_local int g_csp_grid_div;  // >1 while two chunks are in flight on two streams: persistent grids shrink so that both fit on the SMs
extern int g_csp_sm_reserve;  // SMs' worth of CTAs the persistent grids leave free while NCCL kernels run beside them (comm.cu)
inline int csp_grid_sms(int n_sm) { return std::max(1, std::max(n_sm, 1) - g_csp_sm_reserve); }
void csp_timing_begin(int cls, cudaStream_t st);
void csp_timing_end(cudaStream_t st);
struct CspTimed {
    cudaStream_t st;
    bool on;
    CspTimed(int cls, cudaStream_t s) : st(s), on(g_csp_timing) { if (on) csp_timing_begin(cls, s); }
    ~CspTimed() { if (on) csp_timing_end(st); }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

#ifdef __CUDACC__
// ---- mbarrier / bulk-copy PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// FP64 tensor-core MMA (SASS DMMA.8x8x4): D[8x8] += A[8x4] (row) * B[4x8] (col); lane = 4*row + k holds A[row][k], lane = 4*col + k
// holds B[k][col], lane = 4*row + t holds D[row][2t], D[row][2t+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- splits(box) (reference src/tree.jl:461-540) as preorder range scans, any degree p ------------------------------
// The ids splits(box) yields for one ancestor (or for the box's own subtree): up to three ascending ranges, visited in
// this order.  Scanning them as ONE virtual sequence keeps the lanes of a chunk busy (the ranges are mostly tiny).
struct RangeGroup {
    int lo1, n1, lo2, n2, lo3, n3;
    __device__ __forceinline__ int total() const { return n1 + n2 + n3; }
    __device__ __forceinline__ int id(int v) const { return v < n1 ? lo1 + v : (v < n1 + n2 ? lo2 + (v - n1) : lo3 + (v - n1 - n2)); }
};
// Drives `fn(group)` over splits(box) in iteration order until fn returns true (done).
template <class F>
__device__ __forceinline__ void for_split_groups(const TreeView& t, int box, F&& fn) {
    int cnode = box;
    int4 ci = t.ninfo[box];
    if (!(ci.w & CSP_NF_LEAF))
        if (fn(RangeGroup{ci.y, ci.z, 0, 0, 0, 0})) return;  // own split, then its subtree in preorder
    while (ci.x != cnode) {
        const int a = ci.x;
        const int4 ai = t.ninfo[a];
        // subtrees of a's children before c, those after c, then a's own split
        if (fn(RangeGroup{ai.y + 1, ci.y - (ai.y + 1), ci.y + ci.z, ai.y + ai.z - (ci.y + ci.z), ai.y, 1})) return;
        cnode = a;
        ci = ai;
    }
}

#endif
