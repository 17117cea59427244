// extras.cu -- the parts of the path that are not throughput kernels (compiled with -fmad=false like fit.cu: reference order, no
// contraction):
//   * find_leaf_at from ANY box: the general bounds check of reference src/tree.jl:357-360 against boxbounds(box, i)
//     (src/tree.jl:146-171), then the descent of :362-386 from that box; also the descent for degrees p > 2 (the packed
//     32-/16-byte records of the fast kernels exist for p <= 2 only);
//   * boxbounds(box) of any box;
//   * coefficients_p for any degree p <= 8 (src/polynomials.jl:182-230; tested at p = 3 by test/runtests.jl:593-648): one thread
//     per box walks splits(box) as preorder range scans, sequentially, exactly like the reference loop;
//   * minimum / maximum / extrema over leaves(box) (src/CoordinateSplittingPTrees.jl:64-68) as a device reduction.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace {

__device__ __forceinline__ double dnan_x() { return __longlong_as_double(0x7ff8000000000000LL); }

// position(box, d) (src/tree.jl:78-93) from the stored positions: an internal box has its own row; a leaf is its parent's row
// with the coordinates along the displaced split dimensions replaced.  d is 0-based.  The root of a split-less tree has no row.
__device__ double node_position(const TreeView& t, int node, int d) {
    const int4 info = t.ninfo[node];
    if (!(info.w & CSP_NF_LEAF)) return t.ipos[(size_t)info.y * t.n + d];
    if (info.x == node) return dnan_x();
    const int Ip = t.ninfo[info.x].y, ci = CSP_NF_CHILDINDEX(info.w), p = t.p;
    double v = t.ipos[(size_t)Ip * t.n + d];
    for (int k = 0; k < p; ++k)
        if (((ci >> k) & 1) && t.idims[(size_t)Ip * p + k] - 1 == d) v = t.ixs[(size_t)Ip * p + k];
    return v;
}

// boxbounds(box, d) (src/tree.jl:146-171), one thread per dimension
__global__ void k_boxbounds(const TreeView t, int box0, double* __restrict__ lo, double* __restrict__ hi) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= t.n) return;
    double lower = t.lower[d], upper = t.upper[d];
    bool lf = false, uf = false;
    int box = box0;
    const int p = t.p;
    double x = 0.0;
    bool havex = false;
    while (t.ninfo[box].x != box && !(lf & uf)) {
        const int P = t.ninfo[box].x, I = t.ninfo[P].y;
        int i = -1;
        for (int k = 0; k < p; ++k)
            if (i < 0 && t.idims[(size_t)I * p + k] == d + 1) i = k;
        if (i >= 0) {
            const double xs = t.ipos[(size_t)I * t.n + d], xo = t.ixs[(size_t)I * p + i];
            if (!havex) { x = node_position(t, box0, d); havex = true; }
            const double xl = xs < xo ? xs : xo, xu = xs < xo ? xo : xs;
            if (xl <= x && x <= xu) {
                if (!lf && xl < x) {
                    const double a = (xl + xu) / 2, b = (xl + x) / 2;
                    lower = a > b ? a : b; lf = true;
                } else if (!uf && xu > x) {
                    const double a = (xl + xu) / 2, b = (x + xu) / 2;
                    upper = a < b ? a : b; uf = true;
                }
            }
        }
        box = P;
    }
    lo[d] = lower; hi[d] = upper;
}

// find_leaf_at from `start` (src/tree.jl:351-388), any degree: one thread per query on the preorder SoA arrays
__global__ void __launch_bounds__(256) k_locate_from(const TreeView t, int start, const double* __restrict__ lo, const double* __restrict__ hi,
                                                     const double* __restrict__ X, long long m, int* __restrict__ leaf, uint8_t* __restrict__ status) {
    const int n = t.n, p = t.p, C = 1 << p;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const double* __restrict__ x = X + (size_t)q * n;
        bool ok = true;
        for (int i = 0; i < n; ++i) {  // :357-360
            const double xi = x[i], b0 = lo[i], b1 = hi[i];
            ok &= (b0 <= xi) && (xi < b1 || (xi == b1 && b1 == t.upper[i]));
        }
        int lf = -1;
        if (ok) {
            int node = start;
            int4 info = t.ninfo[node];
            while (!(info.w & CSP_NF_LEAF)) {
                const int I = info.y;
                int idx = 0;
                for (int j = 0; j < p; ++j) {
                    const int sd = t.idims[(size_t)I * p + j];
                    if (sd > n) continue;
                    const double xo = t.ixs[(size_t)I * p + j], xself = t.ixself[(size_t)I * p + j];
                    idx |= (int)((x[sd - 1] >= (xself + xo) / 2) ^ (xo < xself)) << j;
                }
                node = t.ichild[(size_t)I * C + idx];
                info = t.ninfo[node];
            }
            lf = t.nleaf[node];
        }
        if (leaf) leaf[q] = lf;
        if (status) status[q] = ok ? CSP_Q_OK : CSP_Q_OUTSIDE;
    }
}

// coefficients_p!, any degree (src/polynomials.jl:182-230): one thread per box, the reference's sequential loop
__global__ void __launch_bounds__(64) k_coef_generic(const TreeView t, const int* __restrict__ node_ids, long long nb, int skip, long long per,
                                                     long long nnz, double* __restrict__ Cp, long long* __restrict__ visited,
                                                     long long* __restrict__ used) {
    const int n = t.n, p = t.p, C = 1 << p;
    for (long long bi = blockIdx.x * (long long)blockDim.x + threadIdx.x; bi < nb; bi += (long long)gridDim.x * blockDim.x) {
        const int box = node_ids[bi];
        double* out = Cp + bi * per;
        for (long long e = 0; e < per; ++e) out[e] = dnan_x();
        long long nrem = nnz, tp = 0, examined = 0, nset = 0;
        if (nrem > 0)
        for_split_groups(t, box, [&](const RangeGroup& grp) -> bool {
            const int total = grp.total();
            for (int v = 0; v < total; ++v, ++tp) {
                const int id = grp.id(v);
                if (tp < skip) continue;
                ++examined;
                int ds[8];
                bool fict = false;
                for (int k = 0; k < p; ++k) { ds[k] = t.idims[(size_t)id * p + k]; fict |= ds[k] > n; }
                if (fict) continue;  // maximum(dims) > n
                for (int a = 1; a < p; ++a) {  // SymmetricArray getindex sorts the index tuple (src/types.jl:451-454)
                    const int key = ds[a];
                    int b = a - 1;
                    while (b >= 0 && ds[b] > key) { ds[b + 1] = ds[b]; --b; }
                    ds[b + 1] = key;
                }
                long long off = 0, mul = 1;
                for (int k = 0; k < p; ++k) { off += (ds[k] - 1) * mul; mul *= n; }
                const double cur = out[off];
                if (cur != cur) {
                    double denom = 1.0;
                    for (int k = 0; k < p; ++k) denom = denom * (t.ixs[(size_t)id * p + k] - t.ixself[(size_t)id * p + k]);
                    const double s = (p & 1) ? -1.0 : 1.0;
                    double num = 0.0;
                    for (int ic = 0; ic < C; ++ic) {
                        const double sg = (__popc(ic) & 1) ? -1.0 : 1.0;
                        num = num + (s * sg) * t.icval[(size_t)id * C + ic];
                    }
                    out[off] = num / denom;
                    nrem -= 1;
                    nset += 1;
                }
                if (nrem <= 0) return true;
            }
            return false;
        });
        if (visited) visited[bi] = examined;
        if (used) used[bi] = nset;
    }
}

// isless(a, b) for Float64 as Julia defines it: NaN sorts last, -0.0 before 0.0
__device__ __forceinline__ bool jl_isless(double a, double b) {
    const bool an = a != a, bn = b != b;
    if (an) return false;
    if (bn) return true;
    if (a < b) return true;
    return a == b && signbit(a) && !signbit(b);
}
struct Ext { double vmin, vmax; int lmin, lmax; };
__device__ __forceinline__ Ext ext_merge(const Ext& a, const Ext& b) {  // a precedes b in leaf order
    Ext r = a;
    if (b.lmin >= 0 && (a.lmin < 0 || jl_isless(b.vmin, a.vmin))) { r.vmin = b.vmin; r.lmin = b.lmin; }
    if (b.lmax >= 0 && (a.lmax < 0 || jl_isless(a.vmax, b.vmax))) { r.vmax = b.vmax; r.lmax = b.lmax; }
    return r;
}
// leaves(box) are the real leaves among the nodes [v0, v0 + count) (a subtree is a contiguous preorder range); each block reduces
// a contiguous slice in order, the host merges the block results in order (first minimal / first maximal leaf wins ties)
__global__ void __launch_bounds__(256) k_extrema(const TreeView t, int v0, int count, Ext* __restrict__ out) {
    __shared__ Ext sh[256];
    const int per = (count + gridDim.x - 1) / gridDim.x;
    const int lo = v0 + blockIdx.x * per, hi = min(v0 + count, lo + per);
    const int span = hi > lo ? hi - lo : 0, tper = (span + 255) / 256;
    Ext e{0.0, 0.0, -1, -1};
    const int a = lo + threadIdx.x * tper, b = min(hi, a + tper);
    const int C = 1 << t.p;
    for (int v = a; v < b; ++v) {
        const int lf = t.nleaf[v];
        if (lf < 0) continue;
        const int4 info = t.ninfo[v];
        const double val = info.x == v ? dnan_x() : t.icval[(size_t)t.ninfo[info.x].y * C + CSP_NF_CHILDINDEX(info.w)];
        e = ext_merge(e, Ext{val, val, lf, lf});
    }
    sh[threadIdx.x] = e;
    __syncthreads();
    for (int s = 1; s < 256; s <<= 1) {  // ordered tree reduction: left operand always precedes the right one
        if ((threadIdx.x & (2 * s - 1)) == 0 && threadIdx.x + s < 256) sh[threadIdx.x] = ext_merge(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

int check_node(const csp_tree* t, int node_id) {
    if (!t) return csp_set_error(CSP_ERR_INVALID, "tree is NULL");
    if (node_id < 0 || node_id >= t->v.n_nodes) return csp_set_error(CSP_ERR_INVALID, "node id %d out of range", node_id);
    return CSP_OK;
}

}  // namespace

extern "C" {

int csp_boxbounds(const csp_tree* t, int32_t node_id, double* lower, double* upper) {
    CSP_CHECK(check_node(t, node_id));
    if (!lower || !upper) return csp_set_error(CSP_ERR_INVALID, "NULL output");
    CSP_CHECK(csp_require_device());
    DeviceGuard g(t->device);
    const int n = t->v.n;
    void* d = nullptr;
    CSP_CHECK(csp_dev_alloc(&d, (size_t)2 * n * 8));
    k_boxbounds<<<(n + 127) / 128, 128>>>(t->v, node_id, (double*)d, (double*)d + n);
    CSP_LAUNCHED();
    cudaError_t e = cudaMemcpy(lower, d, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(upper, (double*)d + n, (size_t)n * 8, cudaMemcpyDeviceToHost);
    csp_dev_free(d);
    if (e != cudaSuccess) return csp_set_error(CSP_ERR_CUDA, "boxbounds failed: %s", cudaGetErrorString(e));
    return CSP_OK;
}

int csp_find_leaf_from_dev(const csp_tree* t, int32_t node_id, const double* dX, int64_t m, int32_t* dleaf, uint8_t* dstatus, void* stream) {
    CSP_CHECK(check_node(t, node_id));
    if (m < 0 || (m > 0 && !dX)) return csp_set_error(CSP_ERR_INVALID, "bad query buffer");
    if (m == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    DeviceGuard g(t->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = t->v.n;
    double* d = nullptr;
    CSP_CUDA(cudaMallocAsync((void**)&d, (size_t)2 * n * 8, st));
    k_boxbounds<<<(n + 127) / 128, 128, 0, st>>>(t->v, node_id, d, d + n);
    CSP_LAUNCHED();
    const long long blocks = std::min<long long>((m + 255) / 256, (long long)std::max(t->n_sm, 1) * 8);
    {
        CspTimed tm(CSP_K_LOCATE, st);
        k_locate_from<<<(unsigned)blocks, 256, 0, st>>>(t->v, node_id, d, d + n, dX, m, dleaf, dstatus);
    }
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    CSP_CUDA(cudaFreeAsync(d, st));
    return CSP_OK;
}

int csp_find_leaf_from(const csp_tree* t, int32_t node_id, const double* X, int64_t m, int32_t* leaf, uint8_t* status) {
    CSP_CHECK(check_node(t, node_id));
    if (m < 0 || (m > 0 && (!X || !leaf))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    if (m == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    DeviceGuard g(t->device);
    const int n = t->v.n;
    const long long chunk = std::min<long long>(m, std::max<long long>(4096, (256LL << 20) / (8LL * n)));
    void *dX = nullptr, *dl = nullptr, *ds = nullptr;
    CSP_CHECK(csp_dev_alloc(&dX, (size_t)chunk * n * 8));
    int rc = csp_dev_alloc(&dl, (size_t)chunk * 4);
    if (rc == CSP_OK) rc = csp_dev_alloc(&ds, (size_t)chunk);
    for (long long off = 0; off < m && rc == CSP_OK; off += chunk) {
        const long long cnt = std::min(chunk, m - off);
        cudaError_t e = cudaMemcpy(dX, X + (size_t)off * n, (size_t)cnt * n * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            rc = csp_find_leaf_from_dev(t, node_id, (const double*)dX, cnt, (int32_t*)dl, (uint8_t*)ds, nullptr);
            if (rc != CSP_OK) break;
            e = cudaMemcpy(leaf + off, dl, (size_t)cnt * 4, cudaMemcpyDeviceToHost);
        }
        if (e == cudaSuccess && status) e = cudaMemcpy(status + off, ds, (size_t)cnt, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "find_leaf_from failed: %s", cudaGetErrorString(e));
    }
    cudaDeviceSynchronize();
    csp_dev_free(dX); csp_dev_free(dl); csp_dev_free(ds);
    return rc;
}

int csp_tree_extrema(const csp_tree* t, int32_t node_id, int32_t* leaf_min, double* val_min, int32_t* leaf_max, double* val_max) {
    CSP_CHECK(check_node(t, node_id));
    CSP_CHECK(csp_require_device());
    DeviceGuard g(t->device);
    // the subtree of a node is the preorder range [v, v + 1 + 2^p * (internal nodes below it)); icount comes from its node record
    int4 info;
    CSP_CUDA(cudaMemcpy(&info, t->v.ninfo + node_id, sizeof info, cudaMemcpyDeviceToHost));
    const long long count = 1 + ((long long)info.z << t->v.p);
    const int blocks = (int)std::min<long long>(std::max<long long>(1, (count + 4095) / 4096), 1024);
    void* d = nullptr;
    CSP_CHECK(csp_dev_alloc(&d, (size_t)blocks * sizeof(Ext)));
    k_extrema<<<blocks, 256>>>(t->v, node_id, (int)count, (Ext*)d);
    CSP_LAUNCHED();
    std::vector<Ext> h(blocks);
    cudaError_t e = cudaMemcpy(h.data(), d, (size_t)blocks * sizeof(Ext), cudaMemcpyDeviceToHost);
    csp_dev_free(d);
    if (e != cudaSuccess) return csp_set_error(CSP_ERR_CUDA, "extrema failed: %s", cudaGetErrorString(e));
    Ext r{0.0, 0.0, -1, -1};
    auto isless = [](double a, double b) {
        if (a != a) return false;
        if (b != b) return true;
        if (a < b) return true;
        return a == b && std::signbit(a) && !std::signbit(b);
    };
    for (const Ext& b : h) {
        if (b.lmin >= 0 && (r.lmin < 0 || isless(b.vmin, r.vmin))) { r.vmin = b.vmin; r.lmin = b.lmin; }
        if (b.lmax >= 0 && (r.lmax < 0 || isless(r.vmax, b.vmax))) { r.vmax = b.vmax; r.lmax = b.lmax; }
    }
    if (r.lmin < 0) return csp_set_error(CSP_ERR_INVALID, "box %d has no real leaves", node_id);
    if (leaf_min) *leaf_min = r.lmin;
    if (val_min) *val_min = r.vmin;
    if (leaf_max) *leaf_max = r.lmax;
    if (val_max) *val_max = r.vmax;
    return CSP_OK;
}

// csp_coefficients_p for p > 2 (and, for cross-checks, any p with CSP_COEF_GENERIC=1)
int csp_coefficients_generic(const csp_tree* t, const int32_t* node_ids, int64_t nb, int32_t skip, double* Cp, int64_t* visited, int64_t* used) {
    DeviceGuard g(t->device);
    const int n = t->v.n, p = t->v.p;
    long long per = 1;
    for (int k = 0; k < p; ++k) {
        per *= n;
        if (per > (1LL << 27)) return csp_set_error(CSP_ERR_UNSUPPORTED, "dense n^p output too large (n=%d, p=%d)", n, p);
    }
    double nz = n;  // nnz_offdiag_sym (src/polynomials.jl:235-242), same floating-point recipe
    for (int i = 1; i <= p - 1; ++i) nz *= (double)(n - i) / (i + 1);
    const long long nnz = llround(nz);
    const long long chunk = std::max<long long>(1, std::min<long long>(nb, (1LL << 30) / (per * 8)));
    void *dids = nullptr, *dC = nullptr, *dvis = nullptr, *dused = nullptr;
    CSP_CHECK(csp_dev_alloc(&dids, (size_t)chunk * 4));
    int rc = csp_dev_alloc(&dC, (size_t)chunk * per * 8);
    if (rc == CSP_OK) rc = csp_dev_alloc(&dvis, (size_t)chunk * 8);
    if (rc == CSP_OK) rc = csp_dev_alloc(&dused, (size_t)chunk * 8);
    std::vector<int> ids;
    for (long long off = 0; off < nb && rc == CSP_OK; off += chunk) {
        const long long cnt = std::min(chunk, nb - off);
        cudaError_t e;
        if (node_ids) {
            e = cudaMemcpy(dids, node_ids + off, (size_t)cnt * 4, cudaMemcpyHostToDevice);
        } else {
            e = cudaMemcpy(dids, t->v.lnode + off, (size_t)cnt * 4, cudaMemcpyDeviceToDevice);
        }
        if (e == cudaSuccess) {
            k_coef_generic<<<(unsigned)std::min<long long>((cnt + 63) / 64, 148 * 16), 64>>>(t->v, (const int*)dids, cnt, skip, per, nnz, (double*)dC,
                                                                                       (long long*)dvis, (long long*)dused);
            CSP_LAUNCHED();
            e = cudaMemcpy(Cp + off * per, dC, (size_t)cnt * per * 8, cudaMemcpyDeviceToHost);
        }
        if (e == cudaSuccess && visited) e = cudaMemcpy(visited + off, dvis, (size_t)cnt * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && used) e = cudaMemcpy(used + off, dused, (size_t)cnt * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "coefficients kernel failed: %s", cudaGetErrorString(e));
    }
    csp_dev_free(dids); csp_dev_free(dC); csp_dev_free(dvis); csp_dev_free(dused);
    return rc;
}

}  // extern "C"
