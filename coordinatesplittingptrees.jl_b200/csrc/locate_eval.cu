// locate_eval.cu -- K1 (batched find_leaf_at) and K3a (fused locate + canonical model evaluation) for sm_100a.
//
// K1 replaces find_leaf_at (reference src/tree.jl:351-388): the root bounds check (:357-360) and the descent
// (:362-386) with the decision rule of :383 evaluated on precomputed split midpoints.
// K3a adds modelvalue(c, g, Q, b, x) (reference src/cs2.jl:155-159) of the located leaf's model.
//
// Shape of the work: an HBM stream of queries (8n B each) against a tree that lives in L2 / L1 / shared memory
// (32 B per split).  Queries are brought into shared memory in contiguous blocks by TMA bulk copies (cp.async.bulk +
// mbarrier) while earlier blocks are being processed.  Three descent kernels, chosen by n:
//   k_locate_ring  (8 <= n <= 32)  one persistent CTA per SM: a ring of 32-row landing slots + one transposed tile per warp,
//                                  warps claim mini-tiles from a shared counter, no CTA barrier in steady state;
//   k_locate_only  (n < 8)         CTA tiles: raw tile + transposed tile per CTA, one barrier per tile;
//   k_locate_wide  (n > 32)        no staging: a warp streams its 32 rows once for the bounds check, lanes descend from L1/L2;
// and the fused k_locate (descent + per-query model fetch and evaluation) for small batches.
//   locate:   one thread owns one query; its coordinates are read from shared memory at the data-dependent split
//             dimensions; one 256-bit record load per level -- from shared memory for the BFS-first `top` records (the top
//             levels of the tree), from L1/L2 below that.
//   evaluate: warp-cooperative; the 32 lanes of a warp read each located leaf's model record with coalesced loads.
#include <atomic>
#include <cstring>
#include <mutex>

#include "common.cuh"

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
    int top;          // number of BFS-first descent records staged in shared memory
    int* hist;        // optional: per-leaf query counts (binned evaluation, binned_eval.cu)
    int* rank;        // optional, with hist: the count the query found (its rank inside its leaf's bin), -1 outside the world
};

__device__ __forceinline__ double dnan_le() { return __longlong_as_double(0x7ff8000000000000LL); }

// 256-bit read-only load of one degree-2 descent record (LDG.E.256)
__device__ __forceinline__ void ldg256(const DRec2* p, double& mid0, double& mid1, int& F, int& G, unsigned& meta, int& H) {
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    mid0 = __longlong_as_double((long long)a);
    mid1 = __longlong_as_double((long long)b);
    F = (int)(unsigned)(c & 0xffffffffu);
    G = (int)(unsigned)(c >> 32);
    meta = (unsigned)(d & 0xffffffffu);
    H = (int)(unsigned)(d >> 32);
}

// Descent (reference src/tree.jl:362-386).  Returns the leaf id.  s_top: the first `top` records, in shared memory.
// Coordinate d of the query is xr[d * xs] (xs = 1: row-major row; xs = tile: the thread's column of a transposed tile).
// topc > 0 (degree 2 only): the first topc BFS records form COMPLETE levels of the tree (every child internal), so the child
// of record i is record 4 i + 1 + idx and neither F, G nor H is needed there: those levels read the midpoints (16 B) and a
// 4-byte copy of meta from the compact array s_meta instead of the record's second 16 bytes -- a random 128-bit shared-memory
// load costs a warp ~10 wavefronts, a random 32-bit one ~3.5.
template <int P>
__device__ __forceinline__ int descend(const TreeView& t, const double* __restrict__ xr, const int xs, const void* s_top, int top,
                                       const unsigned* s_meta = nullptr, int topc = 0) {
    if (t.n_internal == 0) return 0;  // the root is the only leaf
    int cur = 0, dl;
    if (P == 2) {
        const DRec2* __restrict__ recs = (const DRec2*)t.drec;
        const DRec2* stop = (const DRec2*)s_top;
        while (cur < topc) {
            const double2 mm = *(const double2*)&stop[cur].mid0;
            const unsigned meta = s_meta[cur];
            const double x0 = xr[(meta & 0xfffu) * xs];
            const double x1 = xr[((meta >> 12) & 0xfffu) * xs];
            const unsigned b0 = (unsigned)(x0 >= mm.x) ^ ((meta >> 24) & 1u);
            const unsigned b1 = (unsigned)(x1 >= mm.y) ^ ((meta >> 25) & 1u);
            cur = 4 * cur + 1 + (int)((b0 & ~(meta >> 26) & 1u) | ((b1 & ~(meta >> 27) & 1u) << 1));
        }
        while (true) {
            double mid0, mid1;
            int F, G, H;
            unsigned meta;
            if (cur < top) {
                const double2 mm = *(const double2*)&stop[cur].mid0;
                const int4 w = *(const int4*)&stop[cur].F;
                mid0 = mm.x; mid1 = mm.y; F = w.x; G = w.y; meta = (unsigned)w.z; H = w.w;
            } else {
                ldg256(recs + cur, mid0, mid1, F, G, meta, H);
            }
            const double x0 = xr[(meta & 0xfffu) * xs];
            const double x1 = xr[((meta >> 12) & 0xfffu) * xs];
            const unsigned b0 = (unsigned)(x0 >= mid0) ^ ((meta >> 24) & 1u);
            const unsigned b1 = (unsigned)(x1 >= mid1) ^ ((meta >> 25) & 1u);
            const unsigned idx = (b0 & ~(meta >> 26) & 1u) | ((b1 & ~(meta >> 27) & 1u) << 1);
            const unsigned mask = meta >> 28, below = (1u << idx) - 1u;
            if ((mask >> idx) & 1u) {
                cur = F + __popc(mask & below);
            } else {
                // Leaf children hold consecutive leaf ids (preorder) except across an internal sibling: at most two runs among
                // four children.  G >= 0: the record carries the base of both runs (G: before the first internal child that
                // follows a leaf child, H: after it) and the leaf id needs no table; G < 0 (fake leaves in between, odd n):
                // ~G is the node's first descent leaf id and the id comes from the dleaf table.
                const unsigned cnt = __popc(~mask & below & 0xfu);
                if (G >= 0) {
                    const unsigned lead = mask & ~(mask + 1u);  // the internal children before the first leaf child
                    return ((mask & below & ~lead) ? H : G) + (int)cnt;
                }
                dl = ~G + (int)cnt;
                break;
            }
        }
    } else {
        const DRec1* __restrict__ recs = (const DRec1*)t.drec;
        const DRec1* stop = (const DRec1*)s_top;
        while (true) {
            int4 a;
            if (cur < top) a = *(const int4*)&stop[cur];
            else a = __ldg((const int4*)&recs[cur]);
            const double mid = __hiloint2double(a.y, a.x);
            const unsigned w0 = (unsigned)a.z, w1 = (unsigned)a.w;
            const int F = (int)(w0 & 0xffffffu), G = (int)((w0 >> 24) | ((w1 & 0xffffu) << 8));
            const double x0 = xr[((w1 >> 16) & 0xfffu) * xs];
            const unsigned idx = (unsigned)(x0 >= mid) ^ ((w1 >> 28) & 1u);
            const unsigned mask = (w1 >> 29) & 3u, below = (1u << idx) - 1u;
            if ((mask >> idx) & 1u) {
                cur = F + __popc(mask & below);
            } else {
                dl = G + __popc(~mask & below);
                break;
            }
        }
    }
    return __ldg(t.dleaf + dl);
}

// Index pair (i, j) of coefficient e of a model record's coefficient block (everything after b): value = sum coef*z_i*z_j.
__device__ __forceinline__ int coef_pair(int n, int degree, int e) {
    if (degree == 2) {  // packed upper triangle of the (n+1)x(n+1) matrix A, row-major
        int i = 0;
        while (e >= n + 1 - i) { e -= n + 1 - i; ++i; }
        return (i << 16) | (i + e);
    }
    return (e << 16) | n;  // [g | c]: (i, n) for i < n, (n, n) for c
}

// Warp-cooperative canonical model evaluation (reference src/cs2.jl:155-159) of the 32 queries a warp has just located.
// G lanes work on one query (32/G queries per pass): they read the leaf's record with coalesced 8-byte loads (element
// gl + G*s), the lanes holding b_k write z_k = x_k - b_k into a per-group scratch row whose entry n is the constant 1,
// every lane multiplies its coefficients with their two z factors, and a log2(G)-step shuffle tree adds the partial sums.
// NS = ceil(len / G) record slots per lane, unrolled, with the per-slot (i, j) decoded once.
template <int G, int NS>
__device__ __forceinline__ double warp_eval(const ModelsView& mv, const double* __restrict__ rows, double* __restrict__ zs,
                                            int lane, int lf_mine) {
    const unsigned FULL = 0xffffffffu;
    constexpr int NQ = 32 / G;  // queries per pass
    const int n = mv.n, stride = mv.stride;
    const int len = (int)(n + (mv.degree == 2 ? (n + 1) * (n + 2) / 2 : n + 1));
    const int grp = lane / G, gl = lane % G;
    double myval = dnan_le();
    if (__ballot_sync(FULL, lf_mine >= 0) == 0) return myval;
    double* zq = zs + (size_t)grp * (n + 1);
    if (gl == 0) zq[n] = 1.0;
    int code[NS];  // -2: b_k (k = gl + G*s < n)   -1: padding   >= 0: (i << 16) | j
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int k = gl + G * s;
        code[s] = k < n ? -2 : (k < len ? coef_pair(n, mv.degree, k - n) : -1);
    }
    auto load = [&](double (&r)[NS], int pass) {
        const int lf = __shfl_sync(FULL, lf_mine, pass * NQ + grp);
        const double* __restrict__ M = mv.data + (size_t)(lf < 0 ? 0 : lf) * stride;
#pragma unroll
        for (int s = 0; s < NS; ++s) r[s] = (gl + G * s < stride) ? __ldg(M + gl + G * s) : 0.0;
    };
    auto process = [&](const double (&r)[NS], int pass) {
        const int q = pass * NQ + grp;
        const double* xr = rows + (size_t)q * n;
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (code[s] == -2) zq[gl + G * s] = xr[gl + G * s] - r[s];
        __syncwarp();
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (code[s] >= 0) acc = fma(r[s] * zq[code[s] >> 16], zq[code[s] & 0xffff], acc);
#pragma unroll
        for (int o = G / 2; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        // lane L owns query L, which was evaluated in pass L / NQ by group L % NQ
        const double got = __shfl_sync(FULL, acc, (lane % NQ) * G);
        if (lane / NQ == pass) myval = got;
        __syncwarp();  // z scratch is rewritten by the next pass
    };
    double ra[NS], rb[NS];
    load(ra, 0);
#pragma unroll 1
    for (int pass = 0; pass < G; pass += 2) {  // software pipeline: the next pass's record is in flight during this one
        load(rb, pass + 1);
        process(ra, pass);
        if (pass + 2 < G) load(ra, pass + 2);
        process(rb, pass + 1);
    }
    return lf_mine >= 0 ? myval : dnan_le();
}

// any record length: 32 lanes per query, strided loop with an incrementally decoded (i, j)
__device__ __forceinline__ double warp_eval_generic(const ModelsView& mv, const double* __restrict__ rows, double* __restrict__ zq,
                                                    int lane, int lf_mine) {
    const unsigned FULL = 0xffffffffu;
    const int n = mv.n, stride = mv.stride;
    const int ncoef = mv.degree == 2 ? (n + 1) * (n + 2) / 2 : n + 1;
    double myval = dnan_le();
    unsigned live = __ballot_sync(FULL, lf_mine >= 0);
    if (lane == 0) zq[n] = 1.0;
    while (live) {
        const int q = __ffs(live) - 1;
        live &= live - 1;
        const int lf = __shfl_sync(FULL, lf_mine, q);
        const double* __restrict__ M = mv.data + (size_t)lf * stride;
        const double* xr = rows + (size_t)q * n;
        for (int k = lane; k < n; k += 32) zq[k] = xr[k] - __ldg(M + k);
        __syncwarp();
        double acc = 0.0;
        if (mv.degree == 2) {
            int i = 0, e = lane;  // e: offset inside row i of the packed triangle (row length n + 1 - i)
            while (i <= n && e >= n + 1 - i) { e -= n + 1 - i; ++i; }
            for (int k = lane; k < ncoef; k += 32) {
                acc = fma(__ldg(M + n + k) * zq[i], zq[i + e], acc);
                e += 32;
                while (i <= n && e >= n + 1 - i) { e -= n + 1 - i; ++i; }
            }
        } else {
            for (int k = lane; k < ncoef; k += 32) acc = fma(__ldg(M + n + k), zq[k], acc);
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == q) myval = acc;
        __syncwarp();
    }
    return myval;
}

// EG: lanes per query in the evaluation (8/16/32), 0 = generic evaluation, -1 = locate only.  ENS: record slots per lane.
template <int P, int EG, int ENS>
__global__ void __launch_bounds__(256) k_locate(const LocateParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool EVAL = EG >= 0;
    const int n = prm.t.n, tile = prm.tile, stages = prm.stages;
    const size_t tile_doubles = (size_t)tile * n;
    const size_t rec_bytes = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
    unsigned char* s_top = smem_raw;                                     // [top] descent records (32-B aligned)
    double* bufs = (double*)(smem_raw + (((size_t)prm.top * rec_bytes + 127) & ~(size_t)127));
    double* s_lower = bufs + stages * tile_doubles;
    double* s_upper = s_lower + n;
    uint64_t* bars = (uint64_t*)(s_upper + n);
    double* s_z = (double*)(bars + 2);  // evaluation scratch: per warp, 32/EG rows of n + 1 doubles
    const int tid = threadIdx.x, lane = tid & 31;
    const long long ntiles = (prm.m + tile - 1) / tile;

    for (int i = tid; i < n; i += blockDim.x) { s_lower[i] = prm.t.lower[i]; s_upper[i] = prm.t.upper[i]; }
    {   // stage the top of the tree: contiguous prefix of the BFS-ordered record array, 16 bytes per thread-step
        const int4* src = (const int4*)prm.t.drec;
        int4* dst = (int4*)s_top;
        const int n16 = (int)((size_t)prm.top * rec_bytes / 16);
        for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_rows = [&](long long tl) -> int { long long r = prm.m - tl * tile; return r < tile ? (int)r : tile; };
    auto tile_bulk = [&](long long tl) -> bool { return prm.bulk && (((size_t)tile_rows(tl) * n * 8) % 16 == 0); };
    auto issue = [&](long long tl, int s) {  // thread 0 only
        uint32_t bytes = (uint32_t)((size_t)tile_rows(tl) * n * 8);
        mbar_expect_tx(&bars[s], bytes);
        bulk_g2s(bufs + s * tile_doubles, prm.X + (size_t)tl * tile_doubles, bytes, &bars[s]);
    };

    uint32_t phase0 = 0, phase1 = 0;
    long long tl = blockIdx.x;
    if (tl < ntiles && tile_bulk(tl) && tid == 0) issue(tl, 0);
    for (int it = 0; tl < ntiles; ++it, tl += gridDim.x) {
        const int s = stages == 2 ? (it & 1) : 0;
        const long long nxt = tl + gridDim.x;
        if (stages == 2 && nxt < ntiles && tile_bulk(nxt) && tid == 0) issue(nxt, s ^ 1);
        double* buf = bufs + s * tile_doubles;
        const int rows = tile_rows(tl);
        if (tile_bulk(tl)) {
            if (s == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; }
            else { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
        } else {  // unaligned base or odd-sized tail: plain coalesced copy
            const double* src = prm.X + (size_t)tl * tile_doubles;
            for (int e = tid; e < rows * n; e += blockDim.x) buf[e] = src[e];
            __syncthreads();
        }
        int lf = -1;
        const long long q = tl * tile + tid;
        if (tid < rows) {
            const double* xr = buf + (size_t)tid * n;
            bool ok = true;
            for (int i = 0; i < n; ++i) {  // reference src/tree.jl:357-360 at the true root: lower <= x <= upper (NaN fails)
                const double xi = xr[i];
                ok &= (xi >= s_lower[i]) & (xi <= s_upper[i]);
            }
            if (ok) lf = descend<P>(prm.t, xr, 1, s_top, prm.top);
            if (prm.leaf) prm.leaf[q] = lf;
            if (prm.status) prm.status[q] = ok ? CSP_Q_OK : CSP_Q_OUTSIDE;
            if (prm.rank) prm.rank[q] = lf >= 0 ? atomicAdd(prm.hist + lf, 1) : -1;  // returning: the scatter needs no atomics
            else if (prm.hist && lf >= 0) atomicAdd(prm.hist + lf, 1);
        }
        if constexpr (EVAL) {
            __syncwarp();
            const double* wrows = buf + (size_t)(tid - lane) * n;
            double v;
            if constexpr (EG > 0) v = warp_eval<EG, ENS>(prm.mv, wrows, s_z + (size_t)(tid >> 5) * (32 / EG) * (n + 1), lane, lf);
            else v = warp_eval_generic(prm.mv, wrows, s_z + (size_t)(tid >> 5) * (n + 1), lane, lf);
            if (tid < rows) prm.val[q] = v;
        }
        __syncthreads();  // the buffer may be refilled by the next issue
        if (stages == 1 && nxt < ntiles && tile_bulk(nxt) && tid == 0) issue(nxt, 0);
    }
}

// ---- K1, locate only (also the first stage of the binned pipeline) ----------------------------------------------------
// The kernel is bound by L1/shared-memory wavefronts (random 32-byte record reads plus data-dependent coordinate reads),
// so everything around the descent is arranged to cost as few of them as possible:
//   * the TMA-staged row-major tile is transposed once into xT[d][thread]; a thread then reads coordinate d of ITS query
//     at xT[d*tile + tid]: consecutive lanes hit consecutive 8-byte banks whatever d each lane needs (conflict-free);
//   * the root bounds check (reference src/tree.jl:357-360) rides on that transpose pass, with the world bounds read from
//     the kernel parameters (constant bank) when n <= CSP_BOUNDS_PARAM;
//   * the raw tile is dead after the transpose, so a single raw stage is re-filled by the next TMA copy while the
//     descent runs on the transposed copy.
// Tried and rejected (round 2, C3, n = 20: rows 160 B apart put lanes t and t+4 of a quarter-warp into the same bank group
// for the 128-bit row reads): lanes 4-7 walking their row one column pair ahead.  Locate 6.08 -> 7.05 ms per 1e8 queries with a
// per-lane constant-bank index for the bounds, 6.75 -> 7.05 with uniform bounds selected per lane: the extra instructions
// cost more than the two-way conflict.  The same stagger in the ring kernel, whose L1 runs at 90 %: 5.53 -> 5.50 ms (nothing).
#define CSP_BOUNDS_PARAM 32
struct LocateOnlyParams {
    TreeView t;
    const double* X;
    long long m;
    int* leaf;
    uint8_t* status;
    int* hist;
    int* rank;
    int tile, bulk, top, bounds_in_params;
    double lo[CSP_BOUNDS_PARAM], hi[CSP_BOUNDS_PARAM];
};

template <int P>
__global__ void __launch_bounds__(256) k_locate_only(const __grid_constant__ LocateOnlyParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = prm.t.n, tile = prm.tile;
    const size_t tile_doubles = (size_t)tile * n;
    const size_t rec_bytes = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
    unsigned char* s_top = smem_raw;
    double* raw = (double*)(smem_raw + (((size_t)prm.top * rec_bytes + 127) & ~(size_t)127));
    double* xT = raw + tile_doubles;
    double* s_lower = xT + tile_doubles;
    double* s_upper = s_lower + n;
    uint64_t* bar = (uint64_t*)(s_upper + n);
    const int tid = threadIdx.x;
    const long long ntiles = (prm.m + tile - 1) / tile;

    if (!prm.bounds_in_params)
        for (int i = tid; i < n; i += blockDim.x) { s_lower[i] = prm.t.lower[i]; s_upper[i] = prm.t.upper[i]; }
    {
        const int4* src = (const int4*)prm.t.drec;
        int4* dst = (int4*)s_top;
        const int n16 = (int)((size_t)prm.top * rec_bytes / 16);
        for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_rows = [&](long long tl) -> int { long long r = prm.m - tl * tile; return r < tile ? (int)r : tile; };
    auto tile_bulk = [&](long long tl) -> bool { return prm.bulk && (((size_t)tile_rows(tl) * n * 8) % 16 == 0); };
    auto issue = [&](long long tl) {  // thread 0 only
        const uint32_t bytes = (uint32_t)((size_t)tile_rows(tl) * n * 8);
        fence_proxy_async();
        mbar_expect_tx(bar, bytes);
        bulk_g2s(raw, prm.X + (size_t)tl * tile_doubles, bytes, bar);
    };

    uint32_t phase = 0;
    long long tl = blockIdx.x;
    if (tl < ntiles && tile_bulk(tl) && tid == 0) issue(tl);
    for (; tl < ntiles; tl += gridDim.x) {
        const long long nxt = tl + gridDim.x;
        const int rows = tile_rows(tl);
        if (tile_bulk(tl)) {
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            const double* src = prm.X + (size_t)tl * tile_doubles;
            for (int e = tid; e < rows * n; e += blockDim.x) raw[e] = src[e];
            __syncthreads();
        }
        // transpose own row + bounds check
        bool ok = true;
        if (tid < rows) {
            const double* xr = raw + (size_t)tid * n;
            double* col = xT + tid;
            if ((n & 1) == 0) {
                for (int i = 0; i < n; i += 2) {
                    const double2 v = *(const double2*)(xr + i);
                    col[(size_t)i * tile] = v.x;
                    col[(size_t)(i + 1) * tile] = v.y;
                    if (prm.bounds_in_params)
                        ok &= (v.x >= prm.lo[i]) & (v.x <= prm.hi[i]) & (v.y >= prm.lo[i + 1]) & (v.y <= prm.hi[i + 1]);
                    else
                        ok &= (v.x >= s_lower[i]) & (v.x <= s_upper[i]) & (v.y >= s_lower[i + 1]) & (v.y <= s_upper[i + 1]);
                }
            } else {
                for (int i = 0; i < n; ++i) {
                    const double v = xr[i];
                    col[(size_t)i * tile] = v;
                    if (prm.bounds_in_params) ok &= (v >= prm.lo[i]) & (v <= prm.hi[i]);
                    else ok &= (v >= s_lower[i]) & (v <= s_upper[i]);
                }
            }
        }
        __syncthreads();  // every row of the raw tile has been consumed
        if (nxt < ntiles && tile_bulk(nxt) && tid == 0) issue(nxt);
        if (tid < rows) {
            int lf = -1;
            if (ok) lf = descend<P>(prm.t, xT + tid, tile, s_top, prm.top);
            const long long q = tl * tile + tid;
            if (prm.leaf) prm.leaf[q] = lf;
            if (prm.status) prm.status[q] = ok ? CSP_Q_OK : CSP_Q_OUTSIDE;
            if (prm.rank) prm.rank[q] = lf >= 0 ? atomicAdd(prm.hist + lf, 1) : -1;  // returning: the scatter needs no atomics
            else if (prm.hist && lf >= 0) atomicAdd(prm.hist + lf, 1);
        }
        // xT is private per thread column: no barrier needed before the next tile's transpose
    }
}

// ---- K1, warp-granular ring (default for 8 <= n <= 32) --------------------------------------------------------------------
// k_locate_only keeps a raw tile AND a transposed tile per CTA: 16 n bytes of shared memory per resident query, which at n = 20
// leaves 20 warps per SM and room for 85 top records only -- and the descent is latency bound on top of L1 bound (C3: 31 % of the
// warp slots, long_scoreboard the first stall).  Here ONE persistent CTA per SM separates the two roles of that memory:
//   * landing zone: a ring of S slots of 32 rows each, sized for the bytes in flight (S * 32 * 8n bytes, ~50 KB), filled by
//     TMA bulk copies (one per slot: 32 rows are one contiguous block of HBM), each slot with its own mbarrier;
//   * working set: one transposed 32-query tile per WARP (the warp's private xT[d][lane]), 8n bytes per resident query.
// A warp claims the next mini-tile of the CTA from a shared counter, waits for its slot, transposes it into its own xT (bounds
// check on that pass), hands the slot back by issuing the TMA copy of the mini-tile S claims ahead, and descends.  No CTA
// barrier in steady state: the warps drift apart and cover each other's latencies.  C3: 32 warps and the 341-record top.
struct LocateRingParams {
    TreeView t;
    const double* X;
    long long m;
    int* leaf;
    uint8_t* status;
    int* hist;
    int top, slots, bounds_in_params;
    int topc;  // complete-level prefix of the staged top (degree 2), see descend()
    int rotate;
    double lo[CSP_BOUNDS_PARAM], hi[CSP_BOUNDS_PARAM];
};

template <int P>
__global__ void __launch_bounds__(1024, 1) k_locate_ring(const __grid_constant__ LocateRingParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = prm.t.n, S = prm.slots;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const size_t rec_bytes = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
    const size_t mini = (size_t)32 * n;  // doubles per mini-tile
    unsigned char* s_top = smem_raw;
    double* ring = (double*)(smem_raw + (((size_t)prm.top * rec_bytes + 127) & ~(size_t)127));
    double* xT = ring + (size_t)S * mini + (size_t)warp * mini;  // this warp's transposed tile
    double* s_lower = ring + (size_t)S * mini + (size_t)nwarps * mini;
    double* s_upper = s_lower + n;
    uint64_t* bar = (uint64_t*)(s_upper + n);
    int* s_next = (int*)(bar + S);
    volatile int* s_armed = s_next + 1;  // [S] number of fills issued into each slot so far
    unsigned* s_meta = (unsigned*)(s_next + 1 + S);  // [topc]
    const long long nmini = (prm.m + 31) / 32;
    auto rows_of = [&](long long mt) -> int { const long long r = prm.m - mt * 32; return r < 32 ? (int)r : 32; };
    auto bulk_ok = [&](long long mt) -> bool { return ((size_t)rows_of(mt) * n * 8) % 16 == 0; };
    auto fill = [&](int slot, long long mt, int use) {  // one lane; use = how many times the slot was filled before
        const uint32_t bytes = (uint32_t)((size_t)rows_of(mt) * n * 8);
        fence_proxy_async();
        mbar_expect_tx(&bar[slot], bytes);
        bulk_g2s(ring + (size_t)slot * mini, prm.X + (size_t)mt * mini, bytes, &bar[slot]);
        __threadfence_block();
        s_armed[slot] = use + 1;
    };
    if (!prm.bounds_in_params)
        for (int i = tid; i < n; i += blockDim.x) { s_lower[i] = prm.t.lower[i]; s_upper[i] = prm.t.upper[i]; }
    {
        const int4* src = (const int4*)prm.t.drec;
        int4* dst = (int4*)s_top;
        const int n16 = (int)((size_t)prm.top * rec_bytes / 16);
        for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (P == 2)
        for (int i = tid; i < prm.topc; i += blockDim.x) s_meta[i] = ((const DRec2*)prm.t.drec)[i].meta;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&bar[s], 1); s_armed[s] = 0; }
        *s_next = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < S) {  // the first S mini-tiles of this CTA
        const long long mt = blockIdx.x + (long long)tid * gridDim.x;
        if (mt < nmini && bulk_ok(mt)) fill(tid, mt, 0);
    }
    int knext = 0;
    if (lane == 0) knext = atomicAdd(s_next, 1);
    while (true) {
        const int k = __shfl_sync(0xffffffffu, knext, 0);
        const long long mt = blockIdx.x + (long long)k * gridDim.x;
        if (mt >= nmini) break;
        const int slot = k % S, rows = rows_of(mt);
        bool ok = true;
        auto transpose = [&](const double* __restrict__ xr) {  // this lane's row -> its column of xT, with the root bounds check
            double* col = xT + lane;
            if (prm.rotate) {  // lanes with (lane & rotate) walk their row one column pair ahead (bank stagger, see the launcher)
                const int h = n >> 1;
                const bool ahead = (lane & prm.rotate) != 0;
                double cl0 = prm.lo[0], cl1 = prm.lo[1], cu0 = prm.hi[0], cu1 = prm.hi[1];
                for (int i = 0; i < h; ++i) {
                    const int j0 = 2 * i, j1 = i + 1 == h ? 0 : 2 * (i + 1);
                    const double nl0 = prm.lo[j1], nl1 = prm.lo[j1 + 1], nu0 = prm.hi[j1], nu1 = prm.hi[j1 + 1];
                    const int j = ahead ? j1 : j0;
                    const double2 v = *(const double2*)(xr + j);
                    col[(size_t)j * 32] = v.x;
                    col[(size_t)(j + 1) * 32] = v.y;
                    ok &= (v.x >= (ahead ? nl0 : cl0)) & (v.x <= (ahead ? nu0 : cu0)) & (v.y >= (ahead ? nl1 : cl1)) & (v.y <= (ahead ? nu1 : cu1));
                    cl0 = nl0; cl1 = nl1; cu0 = nu0; cu1 = nu1;
                }
            } else if ((n & 1) == 0) {
                for (int i = 0; i < n; i += 2) {
                    const double2 v = *(const double2*)(xr + i);
                    col[(size_t)i * 32] = v.x;
                    col[(size_t)(i + 1) * 32] = v.y;
                    if (prm.bounds_in_params)
                        ok &= (v.x >= prm.lo[i]) & (v.x <= prm.hi[i]) & (v.y >= prm.lo[i + 1]) & (v.y <= prm.hi[i + 1]);
                    else
                        ok &= (v.x >= s_lower[i]) & (v.x <= s_upper[i]) & (v.y >= s_lower[i + 1]) & (v.y <= s_upper[i + 1]);
                }
            } else {
                for (int i = 0; i < n; ++i) {
                    const double v = xr[i];
                    col[(size_t)i * 32] = v;
                    if (prm.bounds_in_params) ok &= (v >= prm.lo[i]) & (v <= prm.hi[i]);
                    else ok &= (v >= s_lower[i]) & (v <= s_upper[i]);
                }
            }
        };
        if (bulk_ok(mt)) {
            // Claims run ahead of the data (32 warps, S slots): the parity wait below is only meaningful once the fill of THIS
            // use has been issued (a parity wait on a barrier that is still one phase behind returns at once), which the
            // consumer of the slot's previous use announces in s_armed.
            const int use = k / S;
            while (s_armed[slot] <= use) __nanosleep(20);
            mbar_wait(&bar[slot], (uint32_t)(use & 1));
            if (lane < rows) transpose(ring + (size_t)slot * mini + (size_t)lane * n);  // shared-memory loads (LDS)
        } else {  // an odd number of odd-length rows (the batch's last mini-tile): not a whole number of 16-byte units.
            if (lane < rows) transpose(prm.X + (size_t)mt * mini + (size_t)lane * n);  // never armed: straight from global memory
        }
        __syncwarp();  // every row of the slot has been consumed: hand it to the mini-tile S claims ahead
        if (lane == 0) {
            const long long mt2 = blockIdx.x + (long long)(k + S) * gridDim.x;
            if (mt2 < nmini && bulk_ok(mt2)) fill(slot, mt2, k / S + 1);
            knext = atomicAdd(s_next, 1);  // the next claim: its latency hides behind the descent
        }
        if (lane < rows) {
            int lf = -1;
            if (ok) lf = descend<P>(prm.t, xT + lane, 32, s_top, prm.top, s_meta, prm.topc);
            const long long q = mt * 32 + lane;
            if (prm.leaf) prm.leaf[q] = lf;
            if (prm.status) prm.status[q] = ok ? CSP_Q_OK : CSP_Q_OUTSIDE;
            if (prm.hist && lf >= 0) atomicAdd(prm.hist + lf, 1);
        }
        __syncwarp();  // xT is rewritten by the next claim
    }
}

// K1 without staging: a warp takes 32 consecutive queries.  Their rows are one contiguous block of 32 n doubles, which the
// warp streams once, fully coalesced, for the root bounds check (src/tree.jl:357-360; one 32-bit "bad row" mask per lane,
// OR-reduced with redux.sync); each lane then descends for its own query reading the few coordinates it needs from the
// same addresses, which the first pass has just pulled into L1/L2.  No shared-memory tile, so occupancy is not tied to n.
template <int P>
__global__ void __launch_bounds__(256) k_locate_wide(const __grid_constant__ LocateOnlyParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = prm.t.n;
    const size_t rec_bytes = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
    unsigned char* s_top = smem_raw;
    double* s_lower = (double*)(smem_raw + (((size_t)prm.top * rec_bytes + 127) & ~(size_t)127));
    double* s_upper = s_lower + n;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < n; i += blockDim.x) { s_lower[i] = prm.t.lower[i]; s_upper[i] = prm.t.upper[i]; }
    {
        const int4* src = (const int4*)prm.t.drec;
        int4* dst = (int4*)s_top;
        const int n16 = (int)((size_t)prm.top * rec_bytes / 16);
        for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5, wid = (blockIdx.x * (long long)blockDim.x + tid) >> 5;
    const long long nbatch = (prm.m + 31) / 32;
    const bool vec = prm.bulk && (n & 1) == 0;  // 16-byte aligned base and even n: whole rows in double2 steps
    for (long long bt = wid; bt < nbatch; bt += nwarps) {
        const long long q0 = bt * 32;
        const int rows = (int)(prm.m - q0 < 32 ? prm.m - q0 : 32);
        const double* __restrict__ base = prm.X + (size_t)q0 * n;
        unsigned bad = 0;
        if (vec) {
            const int total = rows * n / 2;  // double2 elements; a pair never straddles two rows (n even)
            int row = 0, col = 2 * lane;
            while (col >= n) { col -= n; ++row; }
            for (int e = lane; e < total; e += 32) {
                const double2 v = __ldg((const double2*)base + e);
                const bool ok = (v.x >= s_lower[col]) & (v.x <= s_upper[col]) & (v.y >= s_lower[col + 1]) & (v.y <= s_upper[col + 1]);
                if (!ok) bad |= 1u << row;
                col += 64;
                while (col >= n) { col -= n; ++row; }
            }
        } else {
            const int total = rows * n;
            int row = 0, col = lane;
            while (col >= n) { col -= n; ++row; }
            for (int e = lane; e < total; e += 32) {
                const double v = __ldg(base + e);
                if (!((v >= s_lower[col]) & (v <= s_upper[col]))) bad |= 1u << row;
                col += 32;
                while (col >= n) { col -= n; ++row; }
            }
        }
        bad = __reduce_or_sync(0xffffffffu, bad);
        if (lane < rows) {
            const bool ok = !((bad >> lane) & 1u);
            int lf = -1;
            if (ok) lf = descend<P>(prm.t, base + (size_t)lane * n, 1, s_top, prm.top);
            const long long q = q0 + lane;
            if (prm.leaf) prm.leaf[q] = lf;
            if (prm.status) prm.status[q] = ok ? CSP_Q_OK : CSP_Q_OUTSIDE;
            if (prm.rank) prm.rank[q] = lf >= 0 ? atomicAdd(prm.hist + lf, 1) : -1;
            else if (prm.hist && lf >= 0) atomicAdd(prm.hist + lf, 1);
        }
    }
}

// eval-only kernel: leaf ids supplied (scattered models); one thread per query
__global__ void __launch_bounds__(256) k_eval(const ModelsView mv, const double* __restrict__ X, const int* __restrict__ leaf,
                                             long long m, double* __restrict__ val) {
    const int n = mv.n;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const int lf = leaf[q];
        double v = dnan_le();
        if (lf >= 0 && lf < mv.n_models) {
            const double* __restrict__ M = mv.data + (size_t)lf * mv.stride;
            const double* __restrict__ x = X + (size_t)q * n;
            const double* __restrict__ A = M + n;
            double acc = 0.0;
            if (mv.degree == 2) {
                int k = 0;
                for (int i = 0; i < n; ++i) {
                    const double zi = x[i] - __ldg(M + i);
                    double row = 0.0;
                    for (int j = i; j < n; ++j, ++k) row = fma(__ldg(A + k), x[j] - __ldg(M + j), row);
                    row += __ldg(A + k);  // A[i][n] * 1
                    ++k;
                    acc = fma(row, zi, acc);
                }
                acc += __ldg(A + k);  // A[n][n] = c
            } else {
                for (int i = 0; i < n; ++i) acc = fma(__ldg(A + i), x[i] - __ldg(M + i), acc);
                acc += __ldg(A + n);
            }
            v = acc;
        }
        val[q] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// Launch geometry: tile (== block) size, pipeline stages, number of tree-top records staged in shared memory.
static int pick_geometry(const csp_tree* t, LocateParams& prm, size_t* smem) {
    const int n = prm.t.n;
    const size_t rec = prm.t.p == 2 ? sizeof(DRec2) : sizeof(DRec1);
    const char* env = getenv("CSP_TOP_RECORDS");
    int top = env ? atoi(env) : 1024;
    top = std::max(0, std::min(top, prm.t.n_internal));
    int tl = 256, st = 2;
    auto need = [&](int tile, int stages, int tp) {  // + evaluation scratch: <= 4 rows of n + 1 doubles per warp
        return (((size_t)tp * rec + 127) & ~(size_t)127) + (size_t)stages * tile * n * 8 + (size_t)2 * n * 8 + 16 +
               (size_t)(tile / 32) * 4 * (n + 1) * 8;
    };
    const size_t budget = 100 * 1024;  // two CTAs per SM
    while (need(tl, st, top) > budget && tl > 32) tl /= 2;
    if (need(tl, st, top) > budget) st = 1;
    while (need(tl, st, top) > 200 * 1024 && top > 0) top /= 2;
    if (need(tl, st, top) > 200 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the tiled descent kernel", n);
    prm.tile = tl; prm.stages = st; prm.top = top;
    *smem = need(tl, st, top);
    return CSP_OK;
}

template <int P, int EG, int ENS>
static int launch_locate_t(const csp_tree* t, LocateParams& prm, cudaStream_t st) {
    auto kern = k_locate<P, EG, ENS>;
    const csp_tree::Geo* g = nullptr;
    for (auto& e : t->geo)
        if (e.first == (const void*)kern) g = &e.second;
    if (!g) {  // first launch of this kernel for this tree: derive and cache the geometry
        csp_tree::Geo ng{};
        CSP_CHECK(pick_geometry(t, prm, &ng.smem));
        ng.tile = prm.tile; ng.stages = prm.stages; ng.top = prm.top;
        CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ng.smem));
        CSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ng.occ, kern, ng.tile, ng.smem));
        if (ng.occ < 1) return csp_set_error(CSP_ERR_CUDA, "descent kernel does not fit on an SM (smem %zu)", ng.smem);
        t->geo.emplace_back((const void*)kern, ng);
        g = &t->geo.back().second;
    }
    prm.tile = g->tile; prm.stages = g->stages; prm.top = g->top;
    long long ntiles = (prm.m + prm.tile - 1) / prm.tile;
    long long grid = std::min<long long>(ntiles, (long long)csp_grid_sms(t->n_sm) * g->occ);
    {
        CspTimed tm(EG >= 0 ? CSP_K_LOCATE_EVAL : CSP_K_LOCATE, st);
        kern<<<(unsigned)grid, prm.tile, g->smem, st>>>(prm);
    }
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

template <int P>
static int launch_locate_wide(const csp_tree* t, const LocateParams& lp, cudaStream_t st) {
    auto kern = k_locate_wide<P>;
    const int n = t->v.n;
    const size_t rec = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
    const char* env = getenv("CSP_TOP_RECORDS");
    int top = env ? atoi(env) : 341;
    top = std::max(0, std::min(top, t->v.n_internal));
    const size_t smem = (((size_t)top * rec + 127) & ~(size_t)127) + (size_t)2 * n * 8 + 16;
    const csp_tree::Geo* g = nullptr;
    for (auto& e : t->geo)
        if (e.first == (const void*)kern) g = &e.second;
    if (!g) {
        csp_tree::Geo ng{};
        ng.tile = 256; ng.stages = 1; ng.top = top; ng.smem = smem;
        CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        CSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ng.occ, kern, 256, smem));
        if (ng.occ < 1) return csp_set_error(CSP_ERR_CUDA, "descent kernel does not fit on an SM (smem %zu)", smem);
        t->geo.emplace_back((const void*)kern, ng);
        g = &t->geo.back().second;
    }
    LocateOnlyParams prm{};
    prm.t = t->v; prm.X = lp.X; prm.m = lp.m; prm.leaf = lp.leaf; prm.status = lp.status; prm.hist = lp.hist; prm.rank = lp.rank;
    prm.tile = 256; prm.bulk = lp.bulk; prm.top = g->top;
    const long long nbatch = (prm.m + 31) / 32;
    const long long grid = std::min<long long>((nbatch + 7) / 8, (long long)csp_grid_sms(t->n_sm) * std::max(1, g->occ / g_csp_grid_div));
    {
        CspTimed tm(CSP_K_LOCATE, st);
        kern<<<(unsigned)grid, 256, g->smem, st>>>(prm);
    }
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

#define CSP_RING_FROM 8
template <int P>
static int launch_locate_ring(const csp_tree* t, const LocateParams& lp, cudaStream_t st) {
    auto kern = k_locate_ring<P>;
    const int n = t->v.n;
    // Two chunks in flight on two streams (CSP_BIN_STREAMS=2): the L1-bound descent of one chunk is meant to share every SM with
    // the DRAM-bound evaluation of the other, so the kernel keeps its full grid and takes half an SM's shared memory instead.
    const bool shared_sm = g_csp_grid_div > 1;
    const void* key = (const char*)(const void*)kern + (shared_sm ? 1 : 0);
    const csp_tree::Geo* g = nullptr;
    for (auto& e : t->geo)
        if (e.first == key) g = &e.second;
    if (!g) {
        csp_tree::Geo ng{};
        const size_t rec = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
        const size_t mini = (size_t)32 * n * 8;
        const char* wenv = getenv("CSP_RING_WARPS");
        const char* tenv = getenv("CSP_TOP_RECORDS");
        const char* senv = getenv("CSP_RING_SLOTS");
        int warps = wenv ? std::max(1, std::min(32, atoi(wenv))) : (shared_sm ? (n <= 12 ? 24 : 16) : 32);
        // sharing the SM: the tensor-core evaluation CTA takes 114 KB (+ 1 KB reserved each), the FMA evaluation kernel (n <= 12)
        // 10 KB per CTA, four of them.  C2, two chunks in flight, ms per 1e8 queries: 100 KB 9.82, 150 KB 9.36, 180 KB 9.36.
        const char* benv = getenv("CSP_RING_SHARED_KB");
        const size_t budget = shared_sm ? (size_t)(benv ? atoi(benv) : (n <= 12 ? 150 : 100)) * 1024 : 227 * 1024;
        auto fixed = [&](int w, int tp) {
            return (((size_t)tp * rec + 127) & ~(size_t)127) + (size_t)w * mini + (size_t)2 * n * 8 + 64 * 8 + 65 * 4 + (P == 2 ? (size_t)tp * 4 : 0) + 16;
        };
        // landing zone: ~48 KB in flight per SM (the HBM latency x the SM's share of the bandwidth), at least 4 slots
        int slots = senv ? atoi(senv) : (int)std::max<size_t>(4, ((shared_sm ? 24 : 48) * 1024 + mini - 1) / mini);
        // Tree top in shared memory: a level there costs ~0.4 L1 wavefronts per query against 1.0 from L1/L2, so one more level
        // (1365 records of a CS2 tree, 44 KB) is worth giving up warps for as long as 20 remain.  Measured on C3 (n = 20, ms per
        // 1e8 queries): 32 warps + 85 records 6.17, 32 + 341 5.86, 28 + 341 5.94, 24 + 1365 5.74 (the tiled kernel: 6.28).
        int top = P == 2 ? 341 : 255;
        {
            const int big = P == 2 ? 1365 : 1023;
            int w = warps;
            while (w > 4 && fixed(w, big) + (size_t)slots * mini > budget) --w;
            if (w >= (shared_sm ? 16 : 20)) top = big;
        }
        if (tenv) top = atoi(tenv);
        top = std::max(0, std::min(top, t->v.n_internal));
        while (warps > 4 && fixed(warps, top) + (size_t)slots * mini > budget) --warps;
        while (slots > 2 && fixed(warps, top) + (size_t)slots * mini > budget) --slots;
        if (fixed(warps, top) + (size_t)slots * mini > budget) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the ring descent kernel", n);
        slots = std::min(slots, 64);
        ng.tile = warps * 32; ng.stages = slots; ng.top = top; ng.smem = fixed(warps, top) + (size_t)slots * mini;
        // complete levels at the top of the tree (every child internal): read the staged records back once per tree
        ng.occ = 0;  // (re-used as topc for this kernel: one CTA per SM always)
        const char* cenv = getenv("CSP_RING_COMPACT");  // 0: plain records on every level (A/B)
        if (P == 2 && top > 0 && !(cenv && atoi(cenv) == 0)) {
            std::vector<DRec2> h((size_t)top);
            CSP_CUDA(cudaMemcpy(h.data(), t->v.drec, (size_t)top * sizeof(DRec2), cudaMemcpyDeviceToHost));
            int lvl_lo = 0, lvl_n = 1, topc = 0;
            while (lvl_lo + lvl_n <= top) {
                bool full = true;
                for (int i = lvl_lo; i < lvl_lo + lvl_n && full; ++i) full = (h[i].meta >> 28) == 0xfu && h[i].F == 4 * i + 1;
                if (!full) break;
                topc = lvl_lo + lvl_n;
                lvl_lo += lvl_n; lvl_n *= 4;
            }
            ng.occ = topc;
        }
        t->geo.emplace_back(key, ng);
        g = &t->geo.back().second;
    }
    LocateRingParams prm{};
    prm.t = t->v; prm.X = lp.X; prm.m = lp.m; prm.leaf = lp.leaf; prm.status = lp.status; prm.hist = lp.hist;
    prm.top = g->top; prm.slots = g->stages; prm.topc = g->occ;
    prm.bounds_in_params = n <= CSP_BOUNDS_PARAM;
    if (prm.bounds_in_params)
        for (int i = 0; i < n; ++i) { prm.lo[i] = t->h_lower[i]; prm.hi[i] = t->h_upper[i]; }
    {
        const char* renv = getenv("CSP_RING_ROTATE");
        // n/2 = 2 mod 4 (n = 12, 20, 28): rows are n/2 16-byte units apart, so lanes t and t+4 of a quarter-warp hit the same bank
        // group in the 128-bit row reads; lanes with bit 2 set walk their row one column pair ahead.  C3: 5.46 -> 5.37 ms per 1e8
        // queries (masks 8 and 16, i.e. other lane groupings: 5.51 / 5.52).
        prm.rotate = (prm.bounds_in_params && (n & 1) == 0 && ((n >> 1) & 3) == 2) ? (renv ? atoi(renv) : 4) : 0;
    }
    const long long nmini = (prm.m + 31) / 32;
    const long long grid = std::min<long long>(nmini, (long long)csp_grid_sms(t->n_sm));
    {
        // two geometries (whole SM / shared SM) of one kernel function: the opt-in limit must cover the one launched now
        static std::atomic<size_t> optin[2][64];
        const int dev = std::min(63, std::max(0, csp_current_device()));
        if (optin[P - 1][dev].load(std::memory_order_relaxed) < g->smem) {
            CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
            optin[P - 1][dev].store(227 * 1024, std::memory_order_relaxed);
        }
        CspTimed tm(CSP_K_LOCATE, st);
        kern<<<(unsigned)grid, g->tile, g->smem, st>>>(prm);
    }
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

template <int P>
static int launch_locate_only(const csp_tree* t, const LocateParams& lp, cudaStream_t st) {
    {
        // Measured (B200): n = 10: tiled 4.5 ms vs wide 6.0 ms per 1e8 queries; n = 20: 7.0 vs 9.5; n = 100: 6.9 vs 3.0 ms per 2e7
        // (5.3 TB/s).  The staged tile wins while whole tiles of >= 128 queries fit four CTAs to an SM; beyond n = 32 the tile
        // shrinks to a warp or two per CTA and streaming straight from global memory wins.
        const char* wenv = getenv("CSP_LOCATE_WIDE");  // 1: always, 0: never, unset: by n
        const int wide_from = 33;
        if (wenv ? atoi(wenv) != 0 : t->v.n >= wide_from) return launch_locate_wide<P>(t, lp, st);
    }
    {
        const char* renv = getenv("CSP_LOCATE_RING");  // 1: always (n <= 32), 0: never, unset: by n
        const int n = t->v.n;
        const bool ring = renv ? atoi(renv) != 0 : (n >= CSP_RING_FROM && n <= CSP_BOUNDS_PARAM);
        if (ring && lp.bulk && !lp.rank && n <= 64 && lp.m >= 32) return launch_locate_ring<P>(t, lp, st);
    }
    auto kern = k_locate_only<P>;
    const int n = t->v.n;
    const csp_tree::Geo* g = nullptr;
    for (auto& e : t->geo)
        if (e.first == (const void*)kern) g = &e.second;
    if (!g) {
        csp_tree::Geo ng{};
        const size_t rec = P == 2 ? sizeof(DRec2) : sizeof(DRec1);
        const char* env = getenv("CSP_TOP_RECORDS");
        auto need = [&](int tile, int tp) {
            return (((size_t)tp * rec + 127) & ~(size_t)127) + (size_t)2 * tile * n * 8 + (size_t)2 * n * 8 + 16;
        };
        // Staged tree top: levels 0-4 (341 records), 0-3 (85) or 0-2 (21) of a CS2 tree -- whichever keeps the most queries
        // resident per SM (the descent is bound by L1/LSU throughput and latency, not by where the top records live:
        // C3, n = 20: 341 records -> 4 CTAs, 7.0 ms per 1e8 queries; 85 -> 5 CTAs, 6.3 ms).
        int top = 0, tl = 256;
        long long best = -1;
        for (int cand : {341, 85, 21}) {
            int tp = env ? atoi(env) : cand;
            tp = std::max(0, std::min(tp, t->v.n_internal));
            int tile = 256;
            while (need(tile, tp) > 56 * 1024 && tile > 32) tile /= 2;  // four or more CTAs per SM when possible
            while (need(tile, tp) > 200 * 1024 && tp > 0) tp /= 2;
            const long long resident = std::min<long long>(2048, (long long)((227 * 1024) / std::max<size_t>(need(tile, tp), 1)) * tile);
            if (resident > best) { best = resident; top = tp; tl = tile; }
            if (env) break;
        }
        if (need(tl, top) > 200 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the tiled descent kernel", n);
        ng.tile = tl; ng.stages = 1; ng.top = top; ng.smem = need(tl, top);
        CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ng.smem));
        CSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ng.occ, kern, ng.tile, ng.smem));
        if (ng.occ < 1) return csp_set_error(CSP_ERR_CUDA, "descent kernel does not fit on an SM (smem %zu)", ng.smem);
        t->geo.emplace_back((const void*)kern, ng);
        g = &t->geo.back().second;
    }
    LocateOnlyParams prm{};
    prm.t = t->v; prm.X = lp.X; prm.m = lp.m; prm.leaf = lp.leaf; prm.status = lp.status; prm.hist = lp.hist; prm.rank = lp.rank;
    prm.tile = g->tile; prm.bulk = lp.bulk; prm.top = g->top;
    prm.bounds_in_params = n <= CSP_BOUNDS_PARAM;
    if (prm.bounds_in_params)
        for (int i = 0; i < n; ++i) { prm.lo[i] = t->h_lower[i]; prm.hi[i] = t->h_upper[i]; }
    long long ntiles = (prm.m + prm.tile - 1) / prm.tile;
    long long grid = std::min<long long>(ntiles, (long long)csp_grid_sms(t->n_sm) * std::max(1, g->occ / g_csp_grid_div));
    {
        CspTimed tm(CSP_K_LOCATE, st);
        kern<<<(unsigned)grid, prm.tile, g->smem, st>>>(prm);
    }
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

template <int P>
static int launch_locate_p(const csp_tree* t, const csp_models* mo, LocateParams& prm, cudaStream_t st) {
    if (!mo) return launch_locate_only<P>(t, prm, st);
    const long long len = csp_model_len(mo->n, mo->degree);
    if (len <= 80) return launch_locate_t<P, 8, 10>(t, prm, st);    // degree 2: n <= 10
    if (len <= 128) return launch_locate_t<P, 16, 8>(t, prm, st);   // n <= 13
    if (len <= 256) return launch_locate_t<P, 32, 8>(t, prm, st);   // n <= 20
    return launch_locate_t<P, 0, 1>(t, prm, st);
}

// binned_eval.cu
struct BinWorkspace {
    int* cnt = nullptr;
    int* off = nullptr;
    int* cursor = nullptr;
    int* partial = nullptr;
    int2* perm = nullptr;
    int* slot = nullptr;
    double* vsorted = nullptr;
    int* leaf = nullptr;
    int epoch = 0;
    bool ranked = false;
};
int csp_bin_alloc(int nl, long long m, bool need_leaf, cudaStream_t st, BinWorkspace* w);
int csp_bin_alloc_sync(int nl, long long m, BinWorkspace* w);
void csp_bin_free_sync(BinWorkspace* w);
void csp_bin_free(BinWorkspace* w, cudaStream_t st);
int csp_binned_eval(const csp_models* mo, int n_sm, const double* dX, long long m, const int* dleaf, double* dval, BinWorkspace* w,
                    cudaStream_t st);
bool csp_dense_supported(const csp_models* mo);  // dense_eval.cu
int csp_hist_launch(const int* dleaf, long long m, long long n_models, int* cnt, int* rank, int n_sm, cudaStream_t st);

// Evaluation strategy: with many queries per leaf, bin the located queries by leaf and evaluate model-stationary
// (binned_eval.cu); otherwise evaluate inside the locate kernel, fetching each query's model record.
static bool use_binned(const csp_models* mo, long long m) {
    const char* env = getenv("CSP_EVAL_MODE");
    if (env && !strcmp(env, "direct")) return false;
    if (m >= (1LL << 31)) return false;
    if (env && !strcmp(env, "binned")) return true;
    return m >= (1 << 18) && m >= 8 * mo->n_models;
}

// preset: a caller-owned workspace valid for `preset_cap` queries per chunk on stream st (host-buffer path), or null.
static int launch_locate(const csp_tree* t, const csp_models* mo, const double* dX, long long m, int* dleaf, double* dval,
                         uint8_t* dstatus, cudaStream_t st, BinWorkspace* preset = nullptr, long long preset_cap = 0) {
    if (m == 0) return CSP_OK;
    LocateParams prm{};
    prm.t = t->v;
    prm.X = dX; prm.m = m; prm.leaf = dleaf; prm.val = dval; prm.status = dstatus;
    prm.bulk = ((uintptr_t)dX % 16 == 0) ? 1 : 0;
    if (mo && use_binned(mo, m)) {
        // The pipeline runs chunk by chunk: between the kernels of a chunk its leaf ids, permutation, values and (partly)
        // coordinates are still in the 126 MB L2, so the evaluation's gathers and scattered value writes mostly avoid
        // HBM; chunks must stay large enough to amortise the four launches and to keep many queries per leaf.
        // Measured on B200 (C2, 1e8 queries): 0.5M 17.8 ms, 1M 14.5, 2M 12.8, 4M 12.6, 8M 13.2, unchunked 17.0.
        const char* cenv = getenv("CSP_BIN_CHUNK");
        // With many leaves the runs per leaf must stay long enough to amortise a record fetch per warp, while the scatter and the
        // un-permute get slower as the chunk (the range their random accesses cover) grows.  C3, 1e6 leaves, per 1e8 queries
        // (round 2 kernels): 4M chunks 18.1 ms, 8M 15.9, 16M 14.8, 25M 14.3, 40M 14.9, 50M 15.3, 100M 16.1: 25 queries per leaf per chunk.
        long long chunk = cenv ? atoll(cenv)
                               : std::max<long long>(std::max<long long>(1 << 18, (416LL << 20) / (8LL * t->v.n + 24)), 25LL * mo->n_models);
        // The tensor-core evaluation reads every x row exactly once (TMA / cp.async gathers) and expands a leaf's record per
        // run of tiles: nothing is gained from L2-sized chunks, and small chunks shorten the runs (C4, 2e7 queries: 0.5M-query
        // chunks 29.8 ms, one chunk 19.9 ms); only the workspace (24 B/query) bounds the chunk.
        if (!cenv && csp_dense_supported(mo)) chunk = std::max<long long>(chunk, 1LL << 26);
        chunk = std::min(std::max<long long>(chunk, 1024), m);
        if (preset) chunk = std::min(chunk, preset_cap);
        const long long nchunks = (m + chunk - 1) / chunk;
        const char* senv = getenv("CSP_BIN_STREAMS");
        // Two chunks in flight on two internal streams: the L1-bound descent of one chunk shares every SM with the DRAM-bound
        // evaluation of the other (the ring descent kernel then takes half an SM's shared memory, the grids of the other kernels
        // halve).  Measured on B200 per 1e8 queries: C2 (n = 10, FMA evaluation kernel) 10.44 -> 9.70 ms; C3 (n = 20, tensor-core
        // evaluation kernel with 114 KB of staged rows per CTA) 13.3 -> 16.5 ms -- both kernels slow down by more than 2x there.
        // (Round 1, with spatially split grids instead of co-resident CTAs: C2 11.1 -> 12.7 ms.)  Default: two streams where the
        // FMA evaluation kernel and the ring descent kernel are used (8 <= n <= 12), one otherwise; CSP_BIN_STREAMS=1|2 overrides.
        const bool two_default = mo->degree == 2 && t->v.n >= CSP_RING_FROM && t->v.n <= 12 && prm.bulk;
        const int ns = !preset && nchunks > 1 && (senv ? atoi(senv) == 2 : two_default) ? 2 : 1;
        cudaStream_t ss[2] = {st, st};
        if (ns == 2) {
            if (!t->pipe_st[0]) {
                for (int i = 0; i < 2; ++i) {
                    CSP_CUDA(cudaStreamCreateWithFlags(&t->pipe_st[i], cudaStreamNonBlocking));
                    CSP_CUDA(cudaEventCreateWithFlags(&t->pipe_join[i], cudaEventDisableTiming));
                }
                CSP_CUDA(cudaEventCreateWithFlags(&t->pipe_fork, cudaEventDisableTiming));
            }
            CSP_CUDA(cudaEventRecord(t->pipe_fork, st));
            for (int i = 0; i < 2; ++i) {
                ss[i] = t->pipe_st[i];
                CSP_CUDA(cudaStreamWaitEvent(ss[i], t->pipe_fork, 0));
            }
        }
        BinWorkspace w[2];
        int rc = CSP_OK;
        g_csp_grid_div = ns;
        if (preset) w[0] = *preset;
        for (int i = 0; i < ns && rc == CSP_OK && !preset; ++i) rc = csp_bin_alloc((int)mo->n_models, chunk, dleaf == nullptr, ss[i], &w[i]);
        long long ci = 0;
        for (long long off = 0; off < m && rc == CSP_OK; off += chunk, ++ci) {
            const int k = (int)(ci % ns);
            const long long cnt = std::min(chunk, m - off);
            LocateParams cp = prm;
            cp.X = dX + (size_t)off * t->v.n; cp.m = cnt;
            cp.leaf = dleaf ? dleaf + off : w[k].leaf;
            cp.status = dstatus ? dstatus + off : nullptr;
            cp.bulk = ((uintptr_t)cp.X % 16 == 0) ? 1 : 0;
            cp.hist = w[k].cnt;
            // CSP_BIN_RANKED=1: bin ranks from returning histogram atomics in K1, so that the scatter only adds the bin offset.
            // Measured on B200 per 1e8 queries: C3 locate 6.33 -> 7.31 ms and scatter+unpermute 3.31 -> 4.90 ms (the rank array and the
            // offset gather cost more than the atomics they replace); C2 locate 4.47 -> 5.37, scatter 2.37 -> 1.96.  Off by default.
            w[k].ranked = getenv("CSP_BIN_RANKED") && atoi(getenv("CSP_BIN_RANKED")) == 1;
            cp.rank = w[k].ranked ? w[k].slot : nullptr;
            rc = t->v.p == 2 ? launch_locate_p<2>(t, nullptr, cp, ss[k]) : launch_locate_p<1>(t, nullptr, cp, ss[k]);
            if (rc == CSP_OK) {
                csp_models_wait_ready(mo, ss[k]);  // the descent above did not need the records; everything below does
                rc = csp_binned_eval(mo, csp_grid_sms(t->n_sm), cp.X, cnt, cp.leaf, dval + off, &w[k], ss[k]);
            }
        }
        g_csp_grid_div = 1;
        if (preset) preset->epoch = w[0].epoch;
        for (int i = 0; i < ns; ++i) {
            if (!preset) csp_bin_free(&w[i], ss[i]);
            if (ns == 2) {
                cudaEventRecord(t->pipe_join[i], ss[i]);
                cudaStreamWaitEvent(st, t->pipe_join[i], 0);
            }
        }
        return rc;
    }
    if (mo && t->v.n > 256) {
        // Large n, small batch: the fused kernel would need a shared-memory tile of whole rows (it stops fitting near n = 700).
        // Locate with the streaming descent kernel (any n), then evaluate from the leaf ids (grouped dense / per-query kernel).
        int* lbuf = dleaf;
        if (!lbuf) CSP_CUDA(cudaMallocAsync((void**)&lbuf, (size_t)m * 4, st));
        prm.leaf = lbuf;
        int rc = t->v.p == 2 ? launch_locate_p<2>(t, nullptr, prm, st) : launch_locate_p<1>(t, nullptr, prm, st);
        if (rc == CSP_OK) rc = csp_eval_dev(mo, dX, lbuf, m, dval, st);
        if (!dleaf) cudaFreeAsync(lbuf, st);
        return rc;
    }
    if (mo) { prm.mv.n = mo->n; prm.mv.degree = mo->degree; prm.mv.stride = mo->stride; prm.mv.n_models = mo->n_models; prm.mv.data = mo->data; }
    csp_models_wait_ready(mo, st);
    return t->v.p == 2 ? launch_locate_p<2>(t, mo, prm, st) : launch_locate_p<1>(t, mo, prm, st);
}

static int check_pair(const csp_tree* t, const csp_models* mo) {
    if (!t) return csp_set_error(CSP_ERR_INVALID, "tree is NULL");
    if (mo) {
        if (t->v.p > 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "models exist for degrees p <= 2 only (the tree has p=%d)", t->v.p);
        if (mo->device != t->device) return csp_set_error(CSP_ERR_INVALID, "tree and models live on different devices");
        if (mo->n != t->v.n) return csp_set_error(CSP_ERR_INVALID, "models have n=%d, tree has n=%d", mo->n, t->v.n);
        if (mo->n_models != t->v.n_leaves)
            return csp_set_error(CSP_ERR_INVALID, "models hold %lld records, tree has %d leaves", mo->n_models, t->v.n_leaves);
    }
    return CSP_OK;
}

// Host-buffer path: chunked, three streams, H2D / kernel / D2H of consecutive chunks overlap.
// Host-buffer path: chunked, three streams, H2D / kernels / D2H of consecutive chunks overlap.  Streams, staging buffers
// and binning workspaces are created once per device and reused by later calls (grow-only).
struct HostPipe {
    static const int NS = 3;
    cudaStream_t st[NS]{};
    double* dX[NS]{};
    int* dleaf[NS]{};
    double* dval[NS]{};
    uint8_t* dstat[NS]{};
    BinWorkspace bin[NS];
    long long cap_q = 0;      // queries per chunk the staging buffers hold
    size_t cap_xbytes = 0;    // bytes of each dX buffer
    long long bin_cap = 0;    // queries per chunk the binning workspaces hold
    int bin_nl = 0;
    bool streams = false;
};
static HostPipe g_hostpipe[64];
static std::mutex g_hostpipe_mu[64];  // one pipeline per device: different GPUs run concurrently (csp_shard_locate_eval)

static int run_host(const csp_tree* t, const csp_models* mo, bool eval_only, const double* X, const int32_t* leaf_in, long long m,
                    int32_t* leaf, double* val, uint8_t* status);

extern "C" {

int csp_find_leaf_dev(const csp_tree* t, const double* dX, int64_t m, int32_t* dleaf, uint8_t* dstatus, void* stream) {
    CSP_CHECK(check_pair(t, nullptr));
    if (m < 0 || (m > 0 && !dX)) return csp_set_error(CSP_ERR_INVALID, "bad query buffer");
    if (t->v.p > 2) return csp_find_leaf_from_dev(t, 0, dX, m, dleaf, dstatus, stream);  // no packed descent records for p > 2
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
    csp_models_wait_ready(mo, (cudaStream_t)stream);
    if (csp_dense_supported(mo) && m < (1LL << 31) && !(getenv("CSP_EVAL_MODE") && !strcmp(getenv("CSP_EVAL_MODE"), "direct"))) {
        // large n: group the queries by leaf (histogram, scan, scatter) and evaluate on the FP64 tensor cores (K3b)
        cudaStream_t st = (cudaStream_t)stream;
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, mo->device);
        BinWorkspace w;
        int rc = csp_bin_alloc((int)mo->n_models, m, false, st, &w);
        w.ranked = true;
        if (rc == CSP_OK) rc = csp_hist_launch(dleaf, m, mo->n_models, w.cnt, w.slot, nsm, st);
        if (rc == CSP_OK) rc = csp_binned_eval(mo, nsm, dX, m, dleaf, dval, &w, st);
        csp_bin_free(&w, st);
        return rc;
    }
    ModelsView mv{mo->n, mo->degree, mo->stride, mo->n_models, mo->data};
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, mo->device);
    long long blocks = std::min<long long>((m + 255) / 256, (long long)nsm * 8);
    {
        CspTimed tm(CSP_K_EVAL, (cudaStream_t)stream);
        k_eval<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mv, dX, dleaf, m, dval);
    }
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

int csp_find_leaf(const csp_tree* t, const double* X, int64_t m, int32_t* leaf, uint8_t* status) {
    CSP_CHECK(check_pair(t, nullptr));
    if (m < 0 || (m > 0 && (!X || !leaf))) return csp_set_error(CSP_ERR_INVALID, "bad buffer");
    if (t->v.p > 2) return csp_find_leaf_from(t, 0, X, m, leaf, status);
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
    // chunk: ~256 MB of queries (large enough for the binned evaluation to find many queries per leaf), at most m
    long long chunk = std::max<long long>(4096, (256LL << 20) / (8LL * n));
    chunk = std::min(chunk, m);
    if (dev < 0 || dev >= 64) return csp_set_error(CSP_ERR_INVALID, "device index %d not supported by the host path", dev);
    std::lock_guard<std::mutex> lock(g_hostpipe_mu[dev]);
    HostPipe& hp = g_hostpipe[dev];
    if (!hp.streams) {
        for (int i = 0; i < HostPipe::NS; ++i) CSP_CUDA(cudaStreamCreateWithFlags(&hp.st[i], cudaStreamNonBlocking));
        hp.streams = true;
    }
    if (chunk > hp.cap_q || (size_t)chunk * n * 8 > hp.cap_xbytes) {
        for (int i = 0; i < HostPipe::NS; ++i) {
            cudaFree(hp.dX[i]); cudaFree(hp.dleaf[i]); cudaFree(hp.dval[i]); cudaFree(hp.dstat[i]);
            hp.dX[i] = nullptr; hp.dleaf[i] = nullptr; hp.dval[i] = nullptr; hp.dstat[i] = nullptr;
        }
        hp.cap_q = 0; hp.cap_xbytes = 0;
        for (int i = 0; i < HostPipe::NS; ++i) {
            CSP_CUDA(cudaMalloc(&hp.dX[i], (size_t)chunk * n * 8));
            CSP_CUDA(cudaMalloc(&hp.dleaf[i], (size_t)chunk * 4));
            CSP_CUDA(cudaMalloc(&hp.dval[i], (size_t)chunk * 8));
            CSP_CUDA(cudaMalloc(&hp.dstat[i], (size_t)chunk));
        }
        hp.cap_q = chunk; hp.cap_xbytes = (size_t)chunk * n * 8;
    }
    const bool binned = !eval_only && mo && use_binned(mo, chunk);
    if (binned && (hp.bin_cap < chunk || hp.bin_nl != (int)mo->n_models)) {
        for (int i = 0; i < HostPipe::NS; ++i) csp_bin_free_sync(&hp.bin[i]);
        hp.bin_cap = 0;
        for (int i = 0; i < HostPipe::NS; ++i) CSP_CHECK(csp_bin_alloc_sync((int)mo->n_models, chunk, &hp.bin[i]));
        hp.bin_cap = chunk; hp.bin_nl = (int)mo->n_models;
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
            CSP_CHECK(launch_locate(t, mo, hp.dX[s], cnt, hp.dleaf[s], mo ? hp.dval[s] : nullptr, hp.dstat[s], st,
                                    binned ? &hp.bin[s] : nullptr, hp.bin_cap));
        }
        if (leaf) CSP_CUDA(cudaMemcpyAsync(leaf + off, hp.dleaf[s], (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
        if (val) CSP_CUDA(cudaMemcpyAsync(val + off, hp.dval[s], (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
        if (status) CSP_CUDA(cudaMemcpyAsync(status + off, hp.dstat[s], (size_t)cnt, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < HostPipe::NS; ++i) CSP_CUDA(cudaStreamSynchronize(hp.st[i]));
    return CSP_OK;
}
