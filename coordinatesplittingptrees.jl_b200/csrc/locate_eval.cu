// locate_eval.cu -- K1 (batched find_leaf_at) and K3a (fused locate + canonical model evaluation) for sm_100a.
//
// K1 replaces find_leaf_at (reference src/tree.jl:351-388): the root bounds check (:357-360) and the descent
// (:362-386) with the decision rule of :383 evaluated on precomputed split midpoints.
// K3a adds modelvalue(c, g, Q, b, x) (reference src/cs2.jl:155-159) of the located leaf's model.
//
// Shape of the work: an HBM stream of queries (8n B each) against a tree that lives in L2/L1 (48 B per split).
// One CTA processes tiles of `tile` queries; a tile is ONE contiguous block of tile*n doubles in HBM and is brought
// into shared memory by a single TMA bulk copy (cp.async.bulk + mbarrier), double buffered so the next tile streams in
// while the current one is being descended.  One thread owns one query: its coordinates are read from shared memory at
// the data-dependent split dimensions, node records are fetched with 16-byte read-only loads.
#include "common.cuh"

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct LocateParams {
    TreeView t;
    ModelsView mv;
    const double* X;
    long long m;
    int* leaf;        // may be null
    double* val;      // EVAL only
    uint8_t* status;  // may be null
    int tile;         // queries per tile == blockDim.x
    int stages;       // 1 or 2
    int bulk;         // X is 16-byte aligned: TMA bulk copies allowed
};

template <int P>
__device__ __forceinline__ int descend(const TreeView& t, const double* __restrict__ xr) {
    if (t.n_internal == 0) return -2;  // the root is the only leaf (leaf id 0)
    int cur = 0;
    if (P == 2) {
        const DRec2* __restrict__ recs = (const DRec2*)t.drec;
        while (cur >= 0) {
            const double2 mid = __ldg((const double2*)&recs[cur].mid0);
            const int4 ch = __ldg((const int4*)&recs[cur].child[0]);
            const int2 dd = __ldg((const int2*)&recs[cur].d0);
            const bool f0 = dd.x == CSP_DIM_FICTIVE, f1 = dd.y == CSP_DIM_FICTIVE;
            const double x0 = xr[f0 ? 0 : (dd.x & 0x7fffffff)];
            const double x1 = xr[f1 ? 0 : (dd.y & 0x7fffffff)];
            const bool b0 = !f0 && ((x0 >= mid.x) != (dd.x < 0));
            const bool b1 = !f1 && ((x1 >= mid.y) != (dd.y < 0));
            const int lo = b0 ? ch.y : ch.x, hi = b0 ? ch.w : ch.z;
            cur = b1 ? hi : lo;
        }
    } else {
        const DRec1* __restrict__ recs = (const DRec1*)t.drec;
        while (cur >= 0) {
            const int4 a = __ldg((const int4*)&recs[cur].mid);  // mid (2 words), d, pad
            const int2 ch = __ldg((const int2*)&recs[cur].child[0]);
            const double mid = __hiloint2double(a.y, a.x);
            const bool f0 = a.z == CSP_DIM_FICTIVE;
            const double x0 = xr[f0 ? 0 : (a.z & 0x7fffffff)];
            const bool b0 = !f0 && ((x0 >= mid) != (a.z < 0));
            cur = b0 ? ch.y : ch.x;
        }
    }
    return cur;
}

template <int P, bool EVAL>
__global__ void __launch_bounds__(256) k_locate(const LocateParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = prm.t.n, tile = prm.tile, stages = prm.stages;
    const size_t tile_doubles = (size_t)tile * n;
    double* bufs = (double*)smem_raw;
    double* s_lower = bufs + stages * tile_doubles;
    double* s_upper = s_lower + n;
    uint64_t* bars = (uint64_t*)(s_upper + n);
    const int tid = threadIdx.x;
    const long long ntiles = (prm.m + tile - 1) / tile;

    for (int i = tid; i < n; i += blockDim.x) { s_lower[i] = prm.t.lower[i]; s_upper[i] = prm.t.upper[i]; }
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_rows = [&](long long tl) -> int { long long r = prm.m - tl * tile; return r < tile ? (int)r : tile; };
    auto tile_bulk = [&](long long tl) -> bool { return prm.bulk && (((size_t)tile_rows(tl) * n * 8) % 16 == 0); };
    auto issue = [&](long long tl, int s) {  // thread 0 only
        uint32_t bytes = (uint32_t)((size_t)tile_rows(tl) * n * 8);
        fence_proxy_async();
        mbar_expect_tx(&bars[s], bytes);
        bulk_g2s(bufs + s * tile_doubles, prm.X + (size_t)tl * tile_doubles, bytes, &bars[s]);
    };

    uint32_t phase[2] = {0, 0};
    long long tl = blockIdx.x;
    if (tl < ntiles && tile_bulk(tl) && tid == 0) issue(tl, 0);
    for (int it = 0; tl < ntiles; ++it, tl += gridDim.x) {
        const int s = stages == 2 ? (it & 1) : 0;
        const long long nxt = tl + gridDim.x;
        if (stages == 2 && nxt < ntiles && tile_bulk(nxt) && tid == 0) issue(nxt, s ^ 1);
        double* buf = bufs + s * tile_doubles;
        const int rows = tile_rows(tl);
        if (tile_bulk(tl)) {
            mbar_wait(&bars[s], phase[s]);
            phase[s] ^= 1;
        } else {  // unaligned base or odd-sized tail: plain coalesced copy
            const double* src = prm.X + (size_t)tl * tile_doubles;
            for (int e = tid; e < rows * n; e += blockDim.x) buf[e] = src[e];
            __syncthreads();
        }
        if (tid < rows) {
            double* xr = buf + (size_t)tid * n;
            bool ok = true;
            for (int i = 0; i < n; ++i) {  // reference src/tree.jl:357-360 at the true root: lower <= x <= upper (NaN fails)
                const double xi = xr[i];
                ok &= (xi >= s_lower[i]) & (xi <= s_upper[i]);
            }
            int lf = -1;
            if (ok) lf = -descend<P>(prm.t, xr) - 2;
            const long long q = tl * tile + tid;
            if (prm.leaf) prm.leaf[q] = lf;
            if (prm.status) prm.status[q] = ok ? CSP_Q_OK : CSP_Q_OUTSIDE;
            if (EVAL) {
                double v = __longlong_as_double(0x7ff8000000000000LL);
                if (ok) {
                    const double* __restrict__ M = prm.mv.data + (size_t)lf * prm.mv.stride;
                    double lin = 0.0;
                    for (int i = 0; i < n; ++i) {
                        const double dx = xr[i] - __ldg(M + i);
                        xr[i] = dx;
                        lin += __ldg(M + n + i) * dx;
                    }
                    double quad = 0.0;
                    int cpos = 2 * n;
                    if (prm.mv.degree == 2) {
                        const double* __restrict__ Qp = M + 2 * n;
                        int k = 0;
                        for (int i = 0; i < n; ++i) {
                            const double di = xr[i];
                            double sacc = 0.5 * __ldg(Qp + k) * di;
                            ++k;
                            for (int j = i + 1; j < n; ++j, ++k) sacc += __ldg(Qp + k) * xr[j];
                            quad += di * sacc;
                        }
                        cpos += k;
                    }
                    v = __ldg(M + cpos) + lin + quad;
                }
                prm.val[q] = v;
            }
        }
        __syncthreads();  // the buffer may be refilled by the next-but-one issue
        if (stages == 1 && nxt < ntiles && tile_bulk(nxt) && tid == 0) issue(nxt, 0);
    }
}

// eval-only kernel: leaf ids supplied (scattered models), one thread per query, coalesced-by-row reads via L1
__global__ void __launch_bounds__(256) k_eval(const ModelsView mv, const double* __restrict__ X, const int* __restrict__ leaf,
                                             long long m, double* __restrict__ val) {
    const int n = mv.n;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const int lf = leaf[q];
        double v = __longlong_as_double(0x7ff8000000000000LL);
        if (lf >= 0 && lf < mv.n_models) {
            const double* __restrict__ M = mv.data + (size_t)lf * mv.stride;
            const double* __restrict__ x = X + (size_t)q * n;
            double lin = 0.0, quad = 0.0;
            for (int i = 0; i < n; ++i) lin += __ldg(M + n + i) * (x[i] - __ldg(M + i));
            int cpos = 2 * n;
            if (mv.degree == 2) {
                const double* __restrict__ Qp = M + 2 * n;
                int k = 0;
                for (int i = 0; i < n; ++i) {
                    const double di = x[i] - __ldg(M + i);
                    double sacc = 0.5 * __ldg(Qp + k) * di;
                    ++k;
                    for (int j = i + 1; j < n; ++j, ++k) sacc += __ldg(Qp + k) * (x[j] - __ldg(M + j));
                    quad += di * sacc;
                }
                cpos += k;
            }
            v = __ldg(M + cpos) + lin + quad;
        }
        val[q] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int pick_tile(int n, int* tile, int* stages, size_t* smem) {
    int tl = 128, st = 2;
    auto need = [&](int t, int s) { return (size_t)s * t * n * 8 + (size_t)2 * n * 8 + 16; };
    while (need(tl, st) > 96 * 1024 && tl > 32) tl /= 2;
    if (need(tl, st) > 96 * 1024) st = 1;
    if (need(tl, st) > 200 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the tiled descent kernel", n);
    *tile = tl; *stages = st; *smem = need(tl, st);
    return CSP_OK;
}

template <int P, bool EVAL>
static int launch_locate_t(const csp_tree* t, LocateParams& prm, cudaStream_t st) {
    size_t smem;
    CSP_CHECK(pick_tile(prm.t.n, &prm.tile, &prm.stages, &smem));
    auto kern = k_locate<P, EVAL>;
    CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, prm.tile, smem));
    if (occ < 1) return csp_set_error(CSP_ERR_CUDA, "descent kernel does not fit on an SM (smem %zu)", smem);
    long long ntiles = (prm.m + prm.tile - 1) / prm.tile;
    long long grid = std::min<long long>(ntiles, (long long)std::max(t->n_sm, 1) * occ);
    kern<<<(unsigned)grid, prm.tile, smem, st>>>(prm);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

static int launch_locate(const csp_tree* t, const csp_models* mo, const double* dX, long long m, int* dleaf, double* dval,
                         uint8_t* dstatus, cudaStream_t st) {
    if (m == 0) return CSP_OK;
    LocateParams prm{};
    prm.t = t->v;
    prm.X = dX; prm.m = m; prm.leaf = dleaf; prm.val = dval; prm.status = dstatus;
    prm.bulk = ((uintptr_t)dX % 16 == 0) ? 1 : 0;
    if (mo) {
        prm.mv.n = mo->n; prm.mv.degree = mo->degree; prm.mv.stride = mo->stride; prm.mv.n_models = mo->n_models; prm.mv.data = mo->data;
        return t->v.p == 2 ? launch_locate_t<2, true>(t, prm, st) : launch_locate_t<1, true>(t, prm, st);
    }
    return t->v.p == 2 ? launch_locate_t<2, false>(t, prm, st) : launch_locate_t<1, false>(t, prm, st);
}

static int check_pair(const csp_tree* t, const csp_models* mo) {
    if (!t) return csp_set_error(CSP_ERR_INVALID, "tree is NULL");
    if (mo) {
        if (mo->device != t->device) return csp_set_error(CSP_ERR_INVALID, "tree and models live on different devices");
        if (mo->n != t->v.n) return csp_set_error(CSP_ERR_INVALID, "models have n=%d, tree has n=%d", mo->n, t->v.n);
        if (mo->n_models != t->v.n_leaves)
            return csp_set_error(CSP_ERR_INVALID, "models hold %lld records, tree has %d leaves", mo->n_models, t->v.n_leaves);
    }
    return CSP_OK;
}

// Host-buffer path: chunked, three streams, H2D / kernel / D2H of consecutive chunks overlap.
struct HostPipe {
    static const int NS = 3;
    cudaStream_t st[NS]{};
    double* dX[NS]{};
    int* dleaf[NS]{};
    double* dval[NS]{};
    uint8_t* dstat[NS]{};
    int created = 0;
    ~HostPipe() {
        for (int i = 0; i < created; ++i) {
            cudaFree(dX[i]); cudaFree(dleaf[i]); cudaFree(dval[i]); cudaFree(dstat[i]);
            cudaStreamDestroy(st[i]);
        }
    }
};

static int run_host(const csp_tree* t, const csp_models* mo, bool eval_only, const double* X, const int32_t* leaf_in, long long m,
                    int32_t* leaf, double* val, uint8_t* status);

extern "C" {

int csp_find_leaf_dev(const csp_tree* t, const double* dX, int64_t m, int32_t* dleaf, uint8_t* dstatus, void* stream) {
    CSP_CHECK(check_pair(t, nullptr));
    if (m < 0 || (m > 0 && !dX)) return csp_set_error(CSP_ERR_INVALID, "bad query buffer");
    DeviceGuard g(t->device);
    return launch_locate(t, nullptr, dX, m, dleaf, nullptr, dstatus, (cudaStream_t)stream);
}

int csp_locate_eval_dev(const csp_tree* t, const csp_models* mo, const double* dX, int64_t m, int32_t* dleaf, double* dval,
                        uint8_t* dstatus, void* stream) {
    if (!mo) return csp_set_error(CSP_ERR_INVALID, "models is NULL");
    CSP_CHECK(check_pair(t, mo));
    if (m < 0 || (m > 0 && (!dX || !dval))) return csp_set_error(CSP_ERR_INVALID, "bad query / value buffer");
    DeviceGuard g(t->device);
    return launch_locate(t, mo, dX, m, dleaf, dval, dstatus, (cudaStream_t)stream);
}

int csp_eval_dev(const csp_models* mo, const double* dX, const int32_t* dleaf, int64_t m, double* dval, void* stream) {
    if (!mo) return csp_set_error(CSP_ERR_INVALID, "models is NULL");
    if (m < 0 || (m > 0 && (!dX || !dleaf || !dval))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    if (m == 0) return CSP_OK;
    DeviceGuard g(mo->device);
    ModelsView mv{mo->n, mo->degree, mo->stride, mo->n_models, mo->data};
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, mo->device);
    long long blocks = std::min<long long>((m + 255) / 256, (long long)nsm * 8);
    k_eval<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mv, dX, dleaf, m, dval);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

int csp_find_leaf(const csp_tree* t, const double* X, int64_t m, int32_t* leaf, uint8_t* status) {
    CSP_CHECK(check_pair(t, nullptr));
    if (m < 0 || (m > 0 && (!X || !leaf))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    return run_host(t, nullptr, false, X, nullptr, m, leaf, nullptr, status);
}

int csp_locate_eval(const csp_tree* t, const csp_models* mo, const double* X, int64_t m, int32_t* leaf, double* val, uint8_t* status) {
    if (!mo) return csp_set_error(CSP_ERR_INVALID, "models is NULL");
    CSP_CHECK(check_pair(t, mo));
    if (m < 0 || (m > 0 && (!X || !val))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    return run_host(t, mo, false, X, nullptr, m, leaf, val, status);
}

int csp_eval(const csp_models* mo, const double* X, const int32_t* leaf, int64_t m, double* val) {
    if (!mo) return csp_set_error(CSP_ERR_INVALID, "models is NULL");
    if (m < 0 || (m > 0 && (!X || !leaf || !val))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    return run_host(nullptr, mo, true, X, leaf, m, nullptr, val, nullptr);
}

}  // extern "C"

static int run_host(const csp_tree* t, const csp_models* mo, bool eval_only, const double* X, const int32_t* leaf_in, long long m,
                    int32_t* leaf, double* val, uint8_t* status) {
    if (m == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    const int dev = eval_only ? mo->device : t->device;
    const int n = eval_only ? mo->n : t->v.n;
    DeviceGuard g(dev);
    // chunk: ~64 MB of queries, at least one tile's worth, at most m
    long long chunk = std::max<long long>(4096, (64LL << 20) / (8LL * n));
    chunk = std::min(chunk, m);
    HostPipe hp;
    for (int i = 0; i < HostPipe::NS; ++i) {
        CSP_CUDA(cudaStreamCreateWithFlags(&hp.st[i], cudaStreamNonBlocking));
        hp.created = i + 1;
        CSP_CUDA(cudaMalloc(&hp.dX[i], (size_t)chunk * n * 8));
        CSP_CUDA(cudaMalloc(&hp.dleaf[i], (size_t)chunk * 4));
        CSP_CUDA(cudaMalloc(&hp.dval[i], (size_t)chunk * 8));
        CSP_CUDA(cudaMalloc(&hp.dstat[i], (size_t)chunk));
    }
    int ci = 0;
    for (long long off = 0; off < m; off += chunk, ++ci) {
        const int s = ci % HostPipe::NS;
        const long long cnt = std::min(chunk, m - off);
        cudaStream_t st = hp.st[s];
        CSP_CUDA(cudaMemcpyAsync(hp.dX[s], X + (size_t)off * n, (size_t)cnt * n * 8, cudaMemcpyHostToDevice, st));
        if (eval_only) {
            CSP_CUDA(cudaMemcpyAsync(hp.dleaf[s], leaf_in + off, (size_t)cnt * 4, cudaMemcpyHostToDevice, st));
            CSP_CHECK(csp_eval_dev(mo, hp.dX[s], hp.dleaf[s], cnt, hp.dval[s], st));
        } else {
            CSP_CHECK(launch_locate(t, mo, hp.dX[s], cnt, hp.dleaf[s], mo ? hp.dval[s] : nullptr, hp.dstat[s], st));
        }
        if (leaf) CSP_CUDA(cudaMemcpyAsync(leaf + off, hp.dleaf[s], (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        if (val) CSP_CUDA(cudaMemcpyAsync(val + off, hp.dval[s], (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        if (status) CSP_CUDA(cudaMemcpyAsync(status + off, hp.dstat[s], (size_t)cnt, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < HostPipe::NS; ++i) CSP_CUDA(cudaStreamSynchronize(hp.st[i]));
    return CSP_OK;
}
