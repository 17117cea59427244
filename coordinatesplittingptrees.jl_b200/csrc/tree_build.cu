// tree_build.cu -- everything the kernels need beyond the flattened description is DERIVED ON THE DEVICE from the blob sections
// (which stay where the upload / the NCCL broadcast put them; the TreeView points into them): subtree sizes, per-node records,
// split-box coordinates and child values, the BFS-ordered packed descent records, leaf tables, ancestor paths and fit units.
// The same kernels validate the description (what validate_desc checks on the host when a blob is packed).
// Replaces the Box/Split/World object graph of reference src/types.jl:7-23,155-192; the derived quantities are the ones the
// reference recomputes by walking to the root: position (src/tree.jl:78-93), meta/value (:316-323), preorder sizes (:393-415).
#include <cub/cub.cuh>

#include "common.cuh"

enum { E_PARENT = 1, E_INTERNAL_ID, E_LEAF_ID, E_CHILD, E_DIM, E_END };
__device__ __forceinline__ void flag_error(int* err, int code, int where) {
    if (atomicCAS(err, 0, code) == 0) err[1] = where;
}

__global__ void __launch_bounds__(256) kb_flags(const RawTree r, int* __restrict__ isint, int* __restrict__ isreal) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > r.nn) return;
    isint[v] = v < r.nn && r.internal_id[v] >= 0;
    isreal[v] = v < r.nn && r.leaf_id[v] >= 0;
}

// per node: validation, depth, end of the subtree -> internal count, node record
__global__ void __launch_bounds__(256) kb_nodes(const RawTree r, const int* __restrict__ istart, const int* __restrict__ lrank,
                                                int4* __restrict__ ninfo, int* __restrict__ depth, int* __restrict__ err) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= r.nn) return;
    const int C = 1 << r.p;
    const int par = r.parent[v];
    if (v == 0 ? par != 0 : (par < 0 || par >= v)) { flag_error(err, E_PARENT, v); return; }
    const int I = r.internal_id[v], L = r.leaf_id[v];
    const bool internal = I >= 0;
    if (internal) {
        if (I != istart[v] || I >= r.ni || r.inode[I] != v) { flag_error(err, E_INTERNAL_ID, v); return; }
        if (L != -1) { flag_error(err, E_LEAF_ID, v); return; }
    } else if (L >= 0 && L != lrank[v]) {
        flag_error(err, E_LEAF_ID, v);
        return;
    }
    int d = 0, cur = v;
    while (cur != 0) {  // parents precede their children in preorder: strictly decreasing, always terminates
        const int u = r.parent[cur];
        if (u < 0 || u >= cur) { flag_error(err, E_PARENT, cur); return; }
        cur = u;
        ++d;
    }
    depth[v] = d;
    int end = r.nn;
    cur = v;
    while (cur != 0) {  // the subtree ends where the next sibling (of the node or of an ancestor) begins
        const int u = r.parent[cur], k = r.childindex[cur], Iu = r.internal_id[u];
        if (Iu < 0 || Iu >= r.ni || k >= C) { flag_error(err, E_CHILD, cur); return; }
        if (k < C - 1) { end = r.child[(size_t)Iu * C + k + 1]; break; }
        cur = u;
    }
    if (end <= v || end > r.nn) { flag_error(err, E_END, v); return; }
    int flags = r.childindex[v];
    if (!internal) flags |= CSP_NF_LEAF;
    if (!internal && L < 0) flags |= CSP_NF_FAKE;
    ninfo[v] = make_int4(par, istart[v], istart[end] - istart[v], flags);
}

// per split: validation, own coordinates along the split dimensions, child values, depth, child counts
__global__ void __launch_bounds__(256) kb_internal(const RawTree r, const int* __restrict__ depth, double* __restrict__ ixself,
                                                   double* __restrict__ icval, int* __restrict__ idepth, int* __restrict__ iota,
                                                   int* __restrict__ uflag, int* __restrict__ err) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I > r.ni) return;
    if (I == r.ni) { uflag[I] = 0; return; }
    const int n = r.n, p = r.p, C = 1 << p;
    const int v = r.inode[I];
    iota[I] = I;
    uflag[I] = 0;
    if (v < 0 || v >= r.nn || r.internal_id[v] != I) { flag_error(err, E_INTERNAL_ID, v); idepth[I] = 0; return; }
    idepth[I] = depth[v];
    const int maxdim = (n + p - 1) / p * p;
    for (int k = 0; k < p; ++k) {
        const int dim = r.dims[(size_t)I * p + k];
        if (dim < 1 || dim > maxdim) { flag_error(err, E_DIM, v); return; }
        ixself[(size_t)I * p + k] = dim <= n ? r.pos[(size_t)I * n + dim - 1] : 0.0;
    }
    int real = 0;
    for (int k = 0; k < C; ++k) {
        const int c = r.child[(size_t)I * C + k];
        if (c <= v || c >= r.nn || r.parent[c] != v || r.childindex[c] != k || (k == 0 && c != v + 1)) { flag_error(err, E_CHILD, v); return; }
        icval[(size_t)I * C + k] = r.val[c];
        real += r.internal_id[c] < 0 && r.leaf_id[c] >= 0;
    }
    uflag[I] = real > 0;
}

__global__ void __launch_bounds__(256) kb_bfs_counts(const RawTree r, const int* __restrict__ bfs, int* __restrict__ nint, int* __restrict__ nlf) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > r.ni) return;
    if (h == r.ni) { nint[h] = 0; nlf[h] = 0; return; }
    const int C = 1 << r.p, I = bfs[h];
    int a = 0;
    for (int k = 0; k < C; ++k) a += r.internal_id[r.child[(size_t)I * C + k]] >= 0;
    nint[h] = a;
    nlf[h] = C - a;
}

// packed descent records in BFS order (common.cuh); mid = (xself + xother) / 2 with the reference's IEEE operations (src/tree.jl:383)
template <int P>
__global__ void __launch_bounds__(256) kb_records(const RawTree r, const int* __restrict__ bfs, const int* __restrict__ Fex,
                                                  const int* __restrict__ Gex, const double* __restrict__ ixself, void* __restrict__ drec,
                                                  int* __restrict__ dleaf) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= r.ni) return;
    constexpr int C = 1 << P;
    const int I = bfs[h], n = r.n;
    const int F = 1 + Fex[h], G = Gex[h];
    unsigned mask = 0;
    int g = G;
    int L[4] = {-1, -1, -1, -1};
    for (int k = 0; k < C; ++k) {
        const int c = r.child[(size_t)I * C + k];
        if (r.internal_id[c] >= 0) mask |= 1u << k;
        else { dleaf[g++] = r.leaf_id[c]; L[k] = r.leaf_id[c]; }
    }
    double mid[2] = {0, 0};
    unsigned dim0[2] = {0, 0}, flip[2] = {0, 0}, fict[2] = {1, 1};
    for (int k = 0; k < P; ++k) {
        const int dim = r.dims[(size_t)I * P + k];
        const double xo = r.xs[(size_t)I * P + k], xself = ixself[(size_t)I * P + k];
        if (dim <= n) {
            mid[k] = (xself + xo) / 2;
            dim0[k] = (unsigned)(dim - 1); flip[k] = xo < xself ? 1u : 0u; fict[k] = 0;
        }
    }
    if (P == 2) {
        DRec2 q{};
        q.mid0 = mid[0]; q.mid1 = mid[1]; q.F = F;
        csp_encode_leaf_runs(L, mask, fict[0], fict[1], G, &q.G, &q.H);
        q.meta = dim0[0] | (dim0[1] << 12) | (flip[0] << 24) | (flip[1] << 25) | (fict[0] << 26) | (fict[1] << 27) | (mask << 28);
        ((DRec2*)drec)[h] = q;
    } else {
        DRec1 q{};
        q.mid = mid[0];
        q.w0 = (unsigned)F | (((unsigned)G & 0xffu) << 24);
        q.w1 = ((unsigned)G >> 8) | (dim0[0] << 16) | (flip[0] << 28) | (mask << 29);
        ((DRec1*)drec)[h] = q;
    }
}

__global__ void __launch_bounds__(256) kb_leaves(const RawTree r, const int* __restrict__ depth, int* __restrict__ lnode, long long* __restrict__ ldepth) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v == 0) ldepth[r.nl] = 0;
    if (v >= r.nn) return;
    const int L = r.leaf_id[v];
    if (L >= 0 && L < r.nl) { lnode[L] = v; ldepth[L] = depth[v]; }
}

// fit work units (fit.cu): the real leaves that are children of the same split share their splits() traversal
__global__ void __launch_bounds__(256) kb_units(const RawTree r, const int* __restrict__ uflag, const int* __restrict__ uoff, int4* __restrict__ units) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= r.ni || !uflag[I]) return;
    const int C = 1 << r.p;
    int m[4] = {-1, -1, -1, -1}, cnt = 0;
    for (int k = 0; k < C && cnt < 4; ++k) {
        const int c = r.child[(size_t)I * C + k];
        if (r.internal_id[c] < 0 && r.leaf_id[c] >= 0) m[cnt++] = r.leaf_id[c];
    }
    units[uoff[I]] = make_int4(m[0], m[1], m[2], m[3]);
}

// ancestor table for the leaf fits (fit.cu): one thread per leaf climbs to the root once per tree version
__global__ void __launch_bounds__(256) kb_paths(const int4* __restrict__ ninfo, const int* __restrict__ lnode, const long long* __restrict__ lpoff,
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

struct BuildOut {  // small results read back once
    int err[2];
    int total_internal, total_real, n_units, pad;
    long long path_total, max_depth;
};
__global__ void kb_collect(const RawTree r, const int* __restrict__ err, const int* __restrict__ istart, const int* __restrict__ lrank,
                           const int* __restrict__ uoff, const long long* __restrict__ lpoff, const long long* __restrict__ maxd, BuildOut* out) {
    out->err[0] = err[0]; out->err[1] = err[1];
    out->total_internal = istart[r.nn]; out->total_real = lrank[r.nn];
    out->n_units = uoff[r.ni];
    out->path_total = lpoff[r.nl];
    out->max_depth = *maxd;
}

namespace {
struct Arena {  // one allocation for all temporaries
    char* base = nullptr;
    size_t cap = 0, used = 0;
    template <class T>
    T* take(size_t count) {
        used = (used + 255) & ~size_t(255);
        T* p = (T*)(base ? base + used : nullptr);
        used += count * sizeof(T);
        return p;
    }
};
template <class T>
int persist(csp_tree* t, size_t count, T** out) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    CSP_CHECK(csp_dev_alloc(&p, bytes));
    t->allocs.push_back(p);
    t->bytes += bytes;
    *out = (T*)p;
    return CSP_OK;
}
inline unsigned blocks_for(long long items) { return (unsigned)std::max<long long>(1, (items + 255) / 256); }
}  // namespace

// r: device pointers of the description (inside a buffer the tree already owns).  Fills t->v and the derived arrays.
int csp_tree_build_device(csp_tree* t, const RawTree& r, const double* h_lower, const double* h_upper) {
    AllocNoSync nosync;  // everything below runs on the legacy default stream and ends with a device synchronisation
    const int n = r.n, p = r.p, C = 1 << p, nn = r.nn, ni = r.ni, nl = r.nl;
    cudaStream_t st = 0;
    // CUB temporary storage: the largest request of the calls below
    size_t tmp_scan_i = 0, tmp_scan_l = 0, tmp_sort = 0, tmp_max = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan_i, (const int*)nullptr, (int*)nullptr, nn + 1, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan_l, (const long long*)nullptr, (long long*)nullptr, nl + 1, st);
    int depth_bits = 1;  // a depth is below the number of splits: sort only the bits that can be set
    while (depth_bits < 31 && (1LL << depth_bits) <= (long long)ni) ++depth_bits;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr, std::max(ni, 1), 0, depth_bits, st);
    cub::DeviceReduce::Max(nullptr, tmp_max, (const long long*)nullptr, (long long*)nullptr, std::max(nl, 1), st);
    const size_t tmp_bytes = std::max(std::max(tmp_scan_i, tmp_scan_l), std::max(tmp_sort, tmp_max)) + 256;

    Arena A;
    int *isint, *isreal, *istart, *lrank, *depth, *idepth, *iota, *keys_out, *bfs, *nint, *nlf, *Fex, *Gex, *uflag, *uoff, *err;
    long long *ldepth, *maxd;
    BuildOut* dout;
    char* tmp;
    auto carve = [&]() {
        A.used = 0;
        isint = A.take<int>(nn + 1); isreal = A.take<int>(nn + 1); istart = A.take<int>(nn + 1); lrank = A.take<int>(nn + 1);
        depth = A.take<int>(nn);
        idepth = A.take<int>(ni + 1); iota = A.take<int>(ni + 1); keys_out = A.take<int>(ni + 1); bfs = A.take<int>(ni + 1);
        nint = A.take<int>(ni + 1); nlf = A.take<int>(ni + 1); Fex = A.take<int>(ni + 1); Gex = A.take<int>(ni + 1);
        uflag = A.take<int>(ni + 1); uoff = A.take<int>(ni + 1);
        err = A.take<int>(2);
        ldepth = A.take<long long>(nl + 1); maxd = A.take<long long>(1);
        dout = A.take<BuildOut>(1);
        tmp = A.take<char>(tmp_bytes);
    };
    carve();
    void* arena = nullptr;
    CSP_CHECK(csp_dev_alloc(&arena, A.used + 256));
    A.base = (char*)arena;
    carve();
    struct Free { void* p; ~Free() { csp_dev_free(p); } } free_arena{arena};

    TreeView& v = t->v;
    v.n = n; v.p = p; v.n_nodes = nn; v.n_internal = ni; v.n_leaves = nl;
    v.lower = r.lower; v.upper = r.upper; v.nleaf = r.leaf_id; v.inode = r.inode; v.idims = r.dims; v.ixs = r.xs; v.ichild = r.child; v.ipos = r.pos;
    int4* ninfo; double *ixself, *icval; int *dleaf, *lnode; long long* lpoff; int4* units; void* drec = nullptr;
    CSP_CHECK(persist(t, (size_t)nn, &ninfo));
    CSP_CHECK(persist(t, (size_t)ni * p, &ixself));
    CSP_CHECK(persist(t, (size_t)ni * C, &icval));
    CSP_CHECK(persist(t, (size_t)ni * (C - 1) + 1, &dleaf));
    CSP_CHECK(persist(t, (size_t)nl, &lnode));
    CSP_CHECK(persist(t, (size_t)nl + 1, &lpoff));
    CSP_CHECK(persist(t, (size_t)std::max(ni, 1), &units));
    if (p == 1) { DRec1* q; CSP_CHECK(persist(t, (size_t)ni, &q)); drec = q; }
    else if (p == 2) { DRec2* q; CSP_CHECK(persist(t, (size_t)ni, &q)); drec = q; }

    CSP_CUDA(cudaMemsetAsync(err, 0, 8, st));
    CSP_CUDA(cudaMemsetAsync(maxd, 0, 8, st));
    CSP_CUDA(cudaMemsetAsync(dleaf, 0, 4, st));  // a tree without splits: the root is leaf 0
    kb_flags<<<blocks_for(nn + 1), 256, 0, st>>>(r, isint, isreal);
    size_t tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, tb, isint, istart, nn + 1, st);
    tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, tb, isreal, lrank, nn + 1, st);
    kb_nodes<<<blocks_for(nn), 256, 0, st>>>(r, istart, lrank, ninfo, depth, err);
    kb_internal<<<blocks_for(ni + 1), 256, 0, st>>>(r, depth, ixself, icval, idepth, iota, uflag, err);
    tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, tb, uflag, uoff, ni + 1, st);
    if (ni > 0 && p == 2) kb_units<<<blocks_for(ni), 256, 0, st>>>(r, uflag, uoff, units);
    if (ni > 0 && p <= 2) {
        // BFS order == internal nodes sorted by depth, preorder within a depth (a stable sort of the preorder sequence)
        tb = tmp_bytes;
        cub::DeviceRadixSort::SortPairs(tmp, tb, idepth, keys_out, iota, bfs, ni, 0, depth_bits, st);
        kb_bfs_counts<<<blocks_for(ni + 1), 256, 0, st>>>(r, bfs, nint, nlf);
        tb = tmp_bytes;
        cub::DeviceScan::ExclusiveSum(tmp, tb, nint, Fex, ni + 1, st);
        tb = tmp_bytes;
        cub::DeviceScan::ExclusiveSum(tmp, tb, nlf, Gex, ni + 1, st);
        if (p == 2) kb_records<2><<<blocks_for(ni), 256, 0, st>>>(r, bfs, Fex, Gex, ixself, drec, dleaf);
        else kb_records<1><<<blocks_for(ni), 256, 0, st>>>(r, bfs, Fex, Gex, ixself, drec, dleaf);
    }
    kb_leaves<<<blocks_for(nn), 256, 0, st>>>(r, depth, lnode, ldepth);
    tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, tb, ldepth, lpoff, nl + 1, st);
    tb = tmp_bytes;
    if (nl > 0) cub::DeviceReduce::Max(tmp, tb, ldepth, maxd, nl, st);
    kb_collect<<<1, 1, 0, st>>>(r, err, istart, lrank, uoff, lpoff, maxd, dout);
    g_csp_launches.fetch_add(16, std::memory_order_relaxed);
    BuildOut h{};
    CSP_CUDA(cudaMemcpy(&h, dout, sizeof h, cudaMemcpyDeviceToHost));  // synchronises
    CSP_CUDA(cudaGetLastError());
    if (h.err[0]) {
        static const char* what[] = {"", "parent ids are not a preorder tree", "internal ids are not preorder ranks", "leaf ids are not preorder ranks",
                                     "child / parent / childindex are inconsistent", "a split dimension is out of range", "subtree ranges are inconsistent"};
        return csp_set_error(CSP_ERR_INVALID, "invalid tree description at node %d: %s", h.err[1], what[std::min(std::max(h.err[0], 0), 6)]);
    }
    if (h.total_internal != ni || h.total_real != nl)
        return csp_set_error(CSP_ERR_INVALID, "counted %d internal / %d leaves, descriptor says %d / %d", h.total_internal, h.total_real, ni, nl);
    int2* lpath;
    CSP_CHECK(persist(t, (size_t)h.path_total, &lpath));
    if (nl > 0) kb_paths<<<std::min(4096, (nl + 255) / 256), 256, 0, st>>>(ninfo, lnode, lpoff, lpath, nl);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    t->h_lower.resize(n); t->h_upper.resize(n);
    if (h_lower && h_upper) {
        std::copy(h_lower, h_lower + n, t->h_lower.begin());
        std::copy(h_upper, h_upper + n, t->h_upper.begin());
    } else {
        CSP_CUDA(cudaMemcpy(t->h_lower.data(), r.lower, (size_t)n * 8, cudaMemcpyDeviceToHost));
        CSP_CUDA(cudaMemcpy(t->h_upper.data(), r.upper, (size_t)n * 8, cudaMemcpyDeviceToHost));
    }
    v.drec = drec; v.dleaf = dleaf; v.ninfo = ninfo; v.ixself = ixself; v.icval = icval; v.lnode = lnode; v.lpoff = lpoff; v.lpath = lpath;
    t->fit_units = units;
    t->n_fit_units = p == 2 ? (ni == 0 && nl == 1 ? 1 : h.n_units) : 0;
    if (p == 2 && ni == 0 && nl == 1) {  // the root is the only leaf
        const int4 u = make_int4(0, -1, -1, -1);
        CSP_CUDA(cudaMemcpy(units, &u, sizeof u, cudaMemcpyHostToDevice));
    }
    t->max_depth = (int)h.max_depth;
    CSP_CHECK(csp_fit_build_coef(t));
    CSP_CUDA(cudaDeviceSynchronize());
    return CSP_OK;
}
