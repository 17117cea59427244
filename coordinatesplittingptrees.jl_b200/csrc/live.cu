// live.cu -- incremental residency: a tree that GROWS on the device.
//
// addpoint! (reference src/CoordinateSplittingPTrees.jl:94-142) appends ceil(n/p) splits per call, and every split renumbers the
// preorder ids of everything after it.  Instead of re-flattening the Box graph on the host and re-sending ~n_nodes*(n+30) bytes
// per call, the device keeps the SPLIT LOG -- the splits in creation order, in which ids never change: split s divides box
// sp_box[s] and creates the boxes 1 + s*2^p + (0 .. 2^p-1) -- and an append sends only the new entries (~60 bytes per split).
// The preorder flattening (the csp_tree_desc arrays, in the blob layout) is then re-derived on the device: subtree sizes by every
// box adding itself to its ancestors, preorder ranks by summing the sizes of the earlier siblings along the root path,
// internal / leaf ids by prefix sums over the preorder flags, positions by walking to the nearest split that displaced each
// coordinate; tree_build.cu derives the rest, exactly as for an uploaded description.
#include <cub/cub.cuh>

#include "common.cuh"

struct LiveState {
    int n = 0, p = 0;
    long long S = 0, capS = 0;       // splits in the log / capacity
    int* sp_box = nullptr;           // [capS]
    int* sp_dims = nullptr;          // [capS*p]
    double* sp_xs = nullptr;         // [capS*p]
    double* sp_cval = nullptr;       // [capS*2^p]
    double* base = nullptr;          // [n] position(root)
    double root_val = 0.0;
    std::vector<double> lower, upper;
    std::vector<char> is_split;      // host: box already divided (a box is split at most once)
    char* blob = nullptr;            // description arrays in the blob layout, re-derived after every append
    size_t blob_cap = 0;
    int* node_of_box = nullptr;      // [cap boxes] preorder node id of every box (creation-order id)
    int* box_of_node = nullptr;
    long long cap_boxes = 0;
};

void csp_live_free(LiveState* ls) {
    if (!ls) return;
    void* ptrs[] = {ls->sp_box, ls->sp_dims, ls->sp_xs, ls->sp_cval, ls->base, ls->blob, ls->node_of_box, ls->box_of_node};
    for (void* q : ptrs) csp_dev_free(q);
    delete ls;
}

namespace {

__device__ __forceinline__ int lv_parent(const int* sp_box, int b, int C) { return sp_box[(b - 1) / C]; }

// box_split[b] = the split that divides box b, or -1
__global__ void __launch_bounds__(256) kl_box_split(const int* __restrict__ sp_box, long long S, int* __restrict__ box_split) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s < S) box_split[sp_box[s]] = (int)s;
}
// size[b] = boxes in the subtree of b: every box adds one to itself and to each of its ancestors
__global__ void __launch_bounds__(256) kl_sizes(const int* __restrict__ sp_box, long long nb, int C, int* __restrict__ size) {
    const long long b0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b0 >= nb) return;
    int cur = (int)b0;
    while (true) {
        atomicAdd(size + cur, 1);
        if (cur == 0) break;
        cur = lv_parent(sp_box, cur, C);
    }
}
// preorder rank: along the root path, one for every edge plus the subtrees of the siblings that precede the path's child
__global__ void __launch_bounds__(256) kl_preorder(const int* __restrict__ sp_box, const int* __restrict__ box_split, const int* __restrict__ sp_dims,
                                                   const double* __restrict__ sp_cval, const int* __restrict__ size, long long nb, int n, int p,
                                                   double root_val, int* __restrict__ node_of_box, int* __restrict__ box_of_node,
                                                   int* __restrict__ isint, int* __restrict__ isreal, uint8_t* __restrict__ childindex,
                                                   double* __restrict__ val) {
    const long long b0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b0 > nb) return;
    if (b0 == nb) { isint[nb] = 0; isreal[nb] = 0; return; }
    const int C = 1 << p, b = (int)b0;
    int acc = 0, cur = b;
    while (cur != 0) {
        const int s = (cur - 1) / C, ci = (cur - 1) % C, first = 1 + s * C;
        acc += 1;
        for (int j = 0; j < ci; ++j) acc += size[first + j];
        cur = sp_box[s];
    }
    node_of_box[b] = acc;
    box_of_node[acc] = b;
    const bool internal = box_split[b] >= 0;
    bool fake = false;
    int ci = 0;
    double v = root_val;
    if (b != 0) {
        const int s = (b - 1) / C;
        ci = (b - 1) % C;
        v = sp_cval[(size_t)s * C + ci];
        for (int k = 0; k < p; ++k)
            if (((ci >> k) & 1) && sp_dims[(size_t)s * p + k] > n) fake = true;  // displaced along a fictive dimension (src/tree.jl:430-442)
    }
    isint[acc] = internal;
    isreal[acc] = !internal && !fake;
    childindex[acc] = (uint8_t)ci;
    val[acc] = v;
}
// description arrays that need the preorder ranks of other boxes / the prefix sums
__global__ void __launch_bounds__(256) kl_describe(const int* __restrict__ sp_box, const int* __restrict__ box_split, const int* __restrict__ sp_dims,
                                                   const double* __restrict__ sp_xs, const int* __restrict__ node_of_box,
                                                   const int* __restrict__ box_of_node, const int* __restrict__ isint, const int* __restrict__ iscan,
                                                   const int* __restrict__ isreal, const int* __restrict__ lscan, long long nb, int p,
                                                   int* __restrict__ parent, int* __restrict__ internal_id, int* __restrict__ leaf_id,
                                                   int* __restrict__ inode, int* __restrict__ dims, double* __restrict__ xs, int* __restrict__ child) {
    const long long v0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v0 >= nb) return;
    const int C = 1 << p, v = (int)v0, b = box_of_node[v];
    parent[v] = b == 0 ? 0 : node_of_box[lv_parent(sp_box, b, C)];
    internal_id[v] = isint[v] ? iscan[v] : -1;
    leaf_id[v] = isreal[v] ? lscan[v] : -1;
    if (isint[v]) {
        const int I = iscan[v], s = box_split[b];
        inode[I] = v;
        for (int k = 0; k < p; ++k) { dims[(size_t)I * p + k] = sp_dims[(size_t)s * p + k]; xs[(size_t)I * p + k] = sp_xs[(size_t)s * p + k]; }
        for (int c = 0; c < C; ++c) child[(size_t)I * C + c] = node_of_box[1 + s * C + c];
    }
}
// position(box, d) (src/tree.jl:78-93) of every split box: the coordinate of the nearest split on the root path that displaced d
__global__ void __launch_bounds__(256) kl_positions(const int* __restrict__ sp_box, const int* __restrict__ sp_dims, const double* __restrict__ sp_xs,
                                                    const double* __restrict__ base, const int* __restrict__ inode, const int* __restrict__ box_of_node,
                                                    long long ni, int n, int p, double* __restrict__ pos) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ni * n) return;
    const int C = 1 << p, I = (int)(e / n), d = (int)(e - (long long)I * n);
    int cur = box_of_node[inode[I]];
    double x = base[d];
    while (cur != 0) {
        const int s = (cur - 1) / C, ci = (cur - 1) % C;
        bool hit = false;
        for (int k = 0; k < p; ++k)
            if (((ci >> k) & 1) && sp_dims[(size_t)s * p + k] - 1 == d) { x = sp_xs[(size_t)s * p + k]; hit = true; }
        if (hit) break;
        cur = sp_box[s];
    }
    pos[e] = x;
}

template <class T>
int grow(T** p, long long old_count, long long new_cap) {
    void* q = nullptr;
    CSP_CHECK(csp_dev_alloc(&q, (size_t)std::max<long long>(new_cap, 1) * sizeof(T)));
    if (*p && old_count > 0) CSP_CUDA(cudaMemcpy(q, *p, (size_t)old_count * sizeof(T), cudaMemcpyDeviceToDevice));
    csp_dev_free(*p);
    *p = (T*)q;
    return CSP_OK;
}
inline unsigned nblk(long long items) { return (unsigned)std::max<long long>(1, (items + 255) / 256); }

// Re-derive the description (blob layout) from the split log, then everything else (tree_build.cu).
int rebuild(csp_tree* t) {
    AllocNoSync nosync;  // all on the legacy default stream; the builder synchronises at its end
    LiveState& L = *t->live;
    const int n = L.n, p = L.p, C = 1 << p;
    const long long S = L.S, nb = 1 + S * C;
    if (nb > 0x7fffff00LL) return csp_set_error(CSP_ERR_UNSUPPORTED, "tree too large");
    cudaStream_t st = 0;
    // previous version's derived arrays go back to the pool (the description buffer and the log are kept)
    for (void* q : t->allocs) csp_dev_free(q);
    t->allocs.clear();
    t->bytes = 0;
    csp_dev_free(t->fit_coef);
    t->fit_coef = nullptr;
    t->geo.clear();
    if (nb > L.cap_boxes) {
        const long long cap = nb + nb / 2 + 1024;
        csp_dev_free(L.node_of_box); csp_dev_free(L.box_of_node);
        L.node_of_box = L.box_of_node = nullptr;
        CSP_CHECK(csp_dev_alloc((void**)&L.node_of_box, (size_t)cap * 4));
        CSP_CHECK(csp_dev_alloc((void**)&L.box_of_node, (size_t)cap * 4));
        L.cap_boxes = cap;
    }
    Section sec[12];
    const size_t need = blob_layout(n, p, (int)nb, (int)S, sec);
    if (need > L.blob_cap) {
        csp_dev_free(L.blob);
        L.blob = nullptr;
        const size_t cap = need + need / 2 + 65536;
        CSP_CHECK(csp_dev_alloc((void**)&L.blob, cap));
        L.blob_cap = cap;
    }
    char* B = L.blob;
    int* parent = (int*)(B + sec[2].off); uint8_t* childindex = (uint8_t*)(B + sec[3].off);
    int* internal_id = (int*)(B + sec[4].off); int* leaf_id = (int*)(B + sec[5].off); double* val = (double*)(B + sec[6].off);
    int* inode = (int*)(B + sec[7].off); int* dims = (int*)(B + sec[8].off); double* xs = (double*)(B + sec[9].off);
    int* child = (int*)(B + sec[10].off); double* pos = (double*)(B + sec[11].off);
    CSP_CUDA(cudaMemcpyAsync(B + sec[0].off, L.lower.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CSP_CUDA(cudaMemcpyAsync(B + sec[1].off, L.upper.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    // temporaries
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const int*)nullptr, (int*)nullptr, (int)nb + 1, st);
    tmp_bytes += 256;
    void* arena = nullptr;
    const size_t ints = (size_t)(nb + 1);
    CSP_CHECK(csp_dev_alloc(&arena, 6 * ((ints * 4 + 255) & ~size_t(255)) + tmp_bytes + 1024));
    struct Free { void* q; ~Free() { csp_dev_free(q); } } free_arena{arena};
    char* a = (char*)arena;
    auto take = [&](size_t bytes) { char* r = a; a += (bytes + 255) & ~size_t(255); return r; };
    int* box_split = (int*)take(ints * 4); int* size = (int*)take(ints * 4);
    int* isint = (int*)take(ints * 4); int* isreal = (int*)take(ints * 4); int* iscan = (int*)take(ints * 4); int* lscan = (int*)take(ints * 4);
    void* tmp = take(tmp_bytes);
    CSP_CUDA(cudaMemsetAsync(box_split, 0xff, ints * 4, st));
    CSP_CUDA(cudaMemsetAsync(size, 0, ints * 4, st));
    if (S > 0) kl_box_split<<<nblk(S), 256, 0, st>>>(L.sp_box, S, box_split);
    kl_sizes<<<nblk(nb), 256, 0, st>>>(L.sp_box, nb, C, size);
    kl_preorder<<<nblk(nb + 1), 256, 0, st>>>(L.sp_box, box_split, L.sp_dims, L.sp_cval, size, nb, n, p, L.root_val, L.node_of_box, L.box_of_node,
                                             isint, isreal, childindex, val);
    size_t tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, tb, isint, iscan, (int)nb + 1, st);
    tb = tmp_bytes;
    cub::DeviceScan::ExclusiveSum(tmp, tb, isreal, lscan, (int)nb + 1, st);
    kl_describe<<<nblk(nb), 256, 0, st>>>(L.sp_box, box_split, L.sp_dims, L.sp_xs, L.node_of_box, L.box_of_node, isint, iscan, isreal, lscan, nb, p,
                                         parent, internal_id, leaf_id, inode, dims, xs, child);
    if (S > 0) kl_positions<<<nblk(S * n), 256, 0, st>>>(L.sp_box, L.sp_dims, L.sp_xs, L.base, inode, L.box_of_node, S, n, p, pos);
    g_csp_launches.fetch_add(7, std::memory_order_relaxed);
    int nl = 0;
    CSP_CUDA(cudaMemcpy(&nl, lscan + nb, 4, cudaMemcpyDeviceToHost));  // synchronises
    CSP_CUDA(cudaGetLastError());
    return csp_tree_build_device(t, raw_tree_at(B, sec, n, p, (int)nb, (int)S, nl), L.lower.data(), L.upper.data());
}

}  // namespace

extern "C" {

int csp_tree_create_live(int32_t n, int32_t p, const double* lower, const double* upper, const double* root_position, double root_value,
                         csp_tree** out) {
    if (!out) return csp_set_error(CSP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!lower || !upper || !root_position) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    if (n < 1 || n > CSP_MAX_DIM) return csp_set_error(CSP_ERR_INVALID, "n=%d out of range", n);
    if (p < 1 || p > 8) return csp_set_error(CSP_ERR_UNSUPPORTED, "degree p=%d is outside 1..8", p);
    CSP_CHECK(csp_require_device());
    csp_tree* t = new csp_tree();
    t->device = csp_current_device();
    DeviceGuard g(t->device);
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device) == cudaSuccess) t->n_sm = sms;
    LiveState* L = new LiveState();
    t->live = L;
    L->n = n; L->p = p; L->root_val = root_value;
    L->lower.assign(lower, lower + n);
    L->upper.assign(upper, upper + n);
    int rc = csp_dev_alloc((void**)&L->base, (size_t)n * 8);
    if (rc == CSP_OK && cudaMemcpy(L->base, root_position, (size_t)n * 8, cudaMemcpyHostToDevice) != cudaSuccess)
        rc = csp_set_error(CSP_ERR_CUDA, "copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == CSP_OK) rc = rebuild(t);
    if (rc != CSP_OK) { csp_tree_free(t); return rc; }
    *out = t;
    return CSP_OK;
}

int csp_tree_append_splits(csp_tree* t, int64_t ns, const int32_t* split_box, const int32_t* dims, const double* xs, const double* child_val) {
    if (!t || !t->live) return csp_set_error(CSP_ERR_INVALID, "not a live tree (csp_tree_create_live)");
    if (ns < 0 || (ns > 0 && (!split_box || !dims || !xs || !child_val))) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    if (ns == 0) return CSP_OK;
    LiveState& L = *t->live;
    const int p = L.p, C = 1 << p, n = L.n;
    const int maxdim = (n + p - 1) / p * p;
    {   // the log must describe a tree: a box is split once, after it was created
        L.is_split.resize((size_t)(1 + (L.S + ns) * C), 0);
        int64_t k = 0;
        const char* why = nullptr;
        for (; k < ns && !why; ++k) {
            const long long limit = 1 + (L.S + k) * C;  // boxes that exist when split S+k is made
            if (split_box[k] < 0 || split_box[k] >= limit) why = "divides a box that does not exist yet";
            else if (L.is_split[split_box[k]]) why = "divides a box that is split already";
            else {
                for (int j = 0; j < p; ++j)
                    if (dims[k * p + j] < 1 || dims[k * p + j] > maxdim) why = "has a dimension out of range";
                if (!why) L.is_split[split_box[k]] = 1;
            }
        }
        if (why) {
            for (int64_t u = 0; u + 1 < k; ++u) L.is_split[split_box[u]] = 0;  // undo the marks of this call
            L.is_split.resize((size_t)(1 + L.S * C));
            return csp_set_error(CSP_ERR_INVALID, "split %lld (box %d) %s", (long long)(L.S + k - 1), split_box[k - 1], why);
        }
    }
    DeviceGuard g(t->device);
    CSP_CUDA(cudaDeviceSynchronize());  // kernels of any stream may still read the previous version
    if (L.S + ns > L.capS) {
        const long long cap = (L.S + ns) + (L.S + ns) / 2 + 256;
        CSP_CHECK(grow(&L.sp_box, L.S, cap));
        CSP_CHECK(grow(&L.sp_dims, L.S * p, cap * p));
        CSP_CHECK(grow(&L.sp_xs, L.S * p, cap * p));
        CSP_CHECK(grow(&L.sp_cval, L.S * C, cap * C));
        L.capS = cap;
    }
    CSP_CUDA(cudaMemcpyAsync(L.sp_box + L.S, split_box, (size_t)ns * 4, cudaMemcpyHostToDevice, 0));
    CSP_CUDA(cudaMemcpyAsync(L.sp_dims + L.S * p, dims, (size_t)ns * p * 4, cudaMemcpyHostToDevice, 0));
    CSP_CUDA(cudaMemcpyAsync(L.sp_xs + L.S * p, xs, (size_t)ns * p * 8, cudaMemcpyHostToDevice, 0));
    CSP_CUDA(cudaMemcpyAsync(L.sp_cval + L.S * C, child_val, (size_t)ns * C * 8, cudaMemcpyHostToDevice, 0));
    L.S += ns;
    int rc = rebuild(t);
    if (rc != CSP_OK) {  // drop the entries again so that the handle stays usable
        for (int64_t u = 0; u < ns; ++u) L.is_split[split_box[u]] = 0;
        L.S -= ns;
        L.is_split.resize((size_t)(1 + L.S * C));
        rebuild(t);
    }
    return rc;
}

int csp_tree_live_maps(const csp_tree* t, int32_t* node_of_box, int32_t* box_of_node) {
    if (!t || !t->live) return csp_set_error(CSP_ERR_INVALID, "not a live tree (csp_tree_create_live)");
    DeviceGuard g(t->device);
    const size_t nb = (size_t)t->v.n_nodes;
    if (node_of_box) CSP_CUDA(cudaMemcpy(node_of_box, t->live->node_of_box, nb * 4, cudaMemcpyDeviceToHost));
    if (box_of_node) CSP_CUDA(cudaMemcpy(box_of_node, t->live->box_of_node, nb * 4, cudaMemcpyDeviceToHost));
    return CSP_OK;
}

}  // extern "C"
