// fit.cu -- K2: batched per-box model fits for sm_100a.  COMPILED WITH -fmad=false: the reference (Julia) performs no
// FMA contraction on this path and its formulas are cancellation-heavy, so every floating-point expression below is
// evaluated in the reference's order with separate multiplies and adds; results are bit-identical to the oracle.
//
// Replaces, per box (reference paths):
//   chaintop                src/polynomials.jl:22-98
//   fit_poly1               src/cs2.jl:162-205
//   coefficients_p!         src/polynomials.jl:182-230  (dense output, p = 1 or 2)
//   fit_poly2_diag!         src/cs2.jl:208-274  (+ Qcoef_lineq :314-325)
//   canonicalize / chi      src/cs2.jl:132-140, 277-280
//
// Work decomposition: one warp per work unit -- a box, or (all-leaf fits) the real leaves under one split, whose traversals
// coincide (see k_fit).  The reference's `splits(box)` traversal (src/tree.jl:461-540) is replaced by the preorder identity
// (SURVEY.md 3.3): for each ancestor a of the box, bottom-up, coming from child c, the iterator yields the internal-node id
// ranges [I_a+1, I_c) and [I_c+S_c, I_a+S_a) in ascending order, then I_a itself.  So the traversal is a list of contiguous
// range scans over the per-split SoA arrays, flattened into one virtual sequence per window of 32 ancestors (fit_scan);
// "first split in iteration order wins" becomes an integer atomicMin on the scan position per coefficient slot (shared
// memory, or an L2-resident global table for large n), and the reference's early exit becomes "stop when every slot is
// filled".  The O(n) / O(n^2) floating-point work is then done one lane per output entry, sequentially in the reference's
// order (tiles staged through shared memory for large n).
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

#define FIT_MODE_FULL 0   // fit_quadratic (p == 2)
#define FIT_MODE_COEF 1   // coefficients_p only
#define FIT_MODE_LIN 2    // degree-1 model for CS1 trees: c = value(box), g = coefficients_p(box), b = position(box)

struct FitParams {
    TreeView t;
    const int* node_ids;  // null: every leaf in leaves(root) order
    long long nb;
    long long row_lo, row_hi;  // node_ids == null only: fit just the leaves (= output rows) in [row_lo, row_hi) (leaf-sharded fits)
    int skip, mode;
    // strided outputs (doubles); *_stride in doubles per box
    double* c; long long c_stride;
    double* g; long long g_stride;
    double* Q; long long Q_stride; int q_packed;  // 0: dense n*n .data layout; 2: the model record's packed A (common.cuh)
    double* b; long long b_stride;
    double* gnat; double* y; uint8_t* prec;       // optional native-form extras (dense, box-major)
    uint8_t* status;
    long long* visited;
    long long* used;     // optional: number of splits that supplied a coefficient (2^p callback invocations each, src/polynomials.jl:222)
    int* firstmatch;     // optional [nb*n]: fit_poly1's firstmatch (src/cs2.jl:173,191-193), 1-based partner dimension (n+1: fictive)
    // results of the chain pre-pass (k_fit_chain), box-major
    int* ch_status; int* ch_itop; double* ch_c; double* ch_y; double* ch_g; int* ch_step; int* ch_fm;
    // sparse-pattern coefficients_p (FIT_MODE_COEF, p == 2): slotmap[i*n+j] (i<j, 0-based) = output slot or -1; npat slots; the
    // output row of a box is then Q[bi*Q_stride + slot]
    const int* slotmap; int npat; const int* pat_ij;
    unsigned long long* prof;  // optional [8] per-phase clock totals (CSP_FIT_PROF=1)
    int* scratch;            // global pair tables when they do not fit in shared memory
    int pairs_in_smem;
    int q_in_smem;           // the box's Q / record is assembled in shared memory and written out once, coalesced
    int tiles;               // large n (Q stays in global memory): per-warp staging tiles for the O(n^2) loops
    const double* coef;      // [n_internal] coef_value of every split, precomputed once per tree (nullptr: compute in place)
    unsigned long long* work_counter;  // optional (zeroed before the launch): warps claim units dynamically instead of striding
    const int4* units; long long n_units;  // all-leaf fits: the real leaves (output rows, -1 padded) under each split that has any
    int warp_smem_bytes;
    // small n: the per-leaf phases of `multi` sibling leaves run side by side on lane groups of n lanes (multi * n <= 32);
    // their private arrays (b, y, g, Q / step, dkey, did per slot) start multi_off bytes into the warp's slice
    int multi, multi_off;
};

__device__ __forceinline__ double dnan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ bool newinfo(double x, double b, double y) { return (x != b) & (x != y); }

// src/cs2.jl:314-325, including the operator-precedence quirk of line 317 (case A falls through)
__device__ __forceinline__ double qcoef_lineq(bool same, double dxi, double xk, double bk, double yk, bool precik) {
    if (same) {
        if (xk == bk && dxi == yk - bk) {
        } else if (xk == yk && dxi == bk - yk) {
            return 0.0;
        }
        return dxi / 2 + xk - (bk + yk) / 2;
    }
    if (precik) {
        if (xk == bk) return 0.0;
        return 0.0 + (xk - bk);
    }
    return 0.0 + (xk - yk);
}

struct QIndex {
    int n, packed;
    __device__ __forceinline__ int operator()(int i, int j) const {  // 0-based; order-insensitive (SymmetricArray)
        const int lo = min(i, j), hi = max(i, j);  // n <= 4096: every offset fits 32 bits
        return packed ? lo * (n + 1) - lo * (lo - 1) / 2 + (hi - lo) : hi * n + lo;
    }
};

// Offset of Q[i][j] (symmetric) for j = 0, 1, 2, ... without recomputing QIndex each time: start() then step(j) moves j -> j + 1.
struct QWalk {
    int idx, i, n, packed;
    __device__ __forceinline__ QWalk(int i_, int n_, int packed_) : idx(packed_ ? i_ : i_ * n_), i(i_), n(n_), packed(packed_) {}
    __device__ __forceinline__ void step(int j) {
        if (packed) idx += j < i ? n - j : 1;  // column i of rows j < i, then row i
        else idx += j < i ? 1 : n;             // hi * n + lo
    }
};

// per-warp state shared by the helpers
struct WarpCtx {
    int lane, n, npairs;
    double *sb, *sy, *sg;
    int *sstep, *dkey, *did, *pkey, *pid;
    int* plist;  // global-scratch mode only: the slots filled so far, so that nothing has to sweep all n(n-1)/2 of them
    unsigned long long* pk64;  // global-scratch mode: (scan position << 32 | split id) per slot, all ones = empty: one atomic per hit
    int *g_lo1, *g_n1, *g_lo2, *g_lo3, *g_pre;  // one window of up to 32 range groups (g_pre has 33 entries)
    // Several sibling leaves at once (small n, k_fit's multi-member path): slot m of the diagonal scan works on
    // sb/sy = mD(m), mD(m) + n and dkey/did = mI(m) + n, mI(m) + 2n; mmask = the slots taking part.  mmask == 0: the single
    // member of the fields above.
    double* mD0;
    int* mI0;
    int mSD;
    unsigned mmask;
    __device__ __forceinline__ double* mD(int m) const { return mD0 + (size_t)m * mSD; }
    __device__ __forceinline__ int* mI(int m) const { return mI0 + (size_t)m * 3 * n; }
};
#define FIT_WINDOW_INTS (4 * 32 + 33 + 3)  // keeps the following doubles 8-byte aligned
#define FIT_TKB 16                        // k-block width of the staging tiles (rows padded to FIT_TKB + 1 doubles)
#define FIT_TILE_BYTES (2 * 32 * (FIT_TKB + 1) * 8 + 32 * 4)

// chaintop (src/polynomials.jl:22-98) on the flat arrays; executed by one lane.  Returns CSP_FIT_* and the top NODE id.
__device__ int chaintop_dev(const TreeView& t, int box, int* covered, int* top_out) {
    const int n = t.n;
    for (int i = 0; i < n; ++i) covered[i] = 0;
    int nremaining = n;
    auto process = [&](int I, bool& did) {
        const int d0 = t.idims[2 * I], d1 = t.idims[2 * I + 1];
        int nnovel = 2, ntarget = 2;
        nnovel -= (d0 > n || covered[d0 - 1]) ? 1 : 0; ntarget -= (d0 > n) ? 1 : 0;
        nnovel -= (d1 > n || covered[d1 - 1]) ? 1 : 0; ntarget -= (d1 > n) ? 1 : 0;
        if (!(nnovel == ntarget || nnovel == nremaining)) { did = false; return; }
        if (d0 <= n) covered[d0 - 1] = 1;
        if (d1 <= n) covered[d1 - 1] = 1;
        nremaining -= nnovel;
        did = true;
    };
    const int4 bi = t.ninfo[box];
    const int pb = bi.x;
    if (t.ninfo[pb].w & CSP_NF_LEAF) return CSP_FIT_CHAIN_UNDEF;  // root that is a leaf: box.parent.split undefined
    int I = t.ninfo[pb].y;
    bool did = false;
    while (nremaining > 0) {  // go down
        process(I, did);
        if (!did) break;
        const int bottom = t.ichild[4 * I + 3];
        const int4 bo = t.ninfo[bottom];
        if (bo.w & CSP_NF_LEAF) break;
        if (t.idims[2 * I] > n || t.idims[2 * I + 1] > n) break;
        I = bo.y;
    }
    int top = pb;
    if (nremaining == 0) { *top_out = top; return CSP_FIT_OK; }
    if (pb == box) return CSP_FIT_NO_CHAIN;  // isroot(box)
    int ptop = t.ninfo[top].x;
    I = t.ninfo[ptop].y;
    while (nremaining > 0) {  // go up
        process(I, did);
        if (ptop == top) return CSP_FIT_NO_CHAIN;  // isroot(top)
        top = ptop;
        ptop = t.ninfo[top].x;
        I = t.ninfo[ptop].y;
    }
    *top_out = top;
    return CSP_FIT_OK;
}

// Chain pre-pass, one THREAD per box: getleaf + chaintop (src/polynomials.jl:22-98) + the chain walk of fit_poly1
// (src/cs2.jl:183-203).  Both are short sequential integer walks; running them on one lane of a warp-per-box kernel left
// 31 lanes idle for a fifth of that kernel's instructions.
__global__ void __launch_bounds__(128) k_fit_chain(const TreeView t, const int* __restrict__ node_ids, long long row_lo, long long nb, int* __restrict__ ch_status,
                                                   int* __restrict__ ch_itop, double* __restrict__ ch_c, double* __restrict__ ch_y,
                                                   double* __restrict__ ch_g, int* __restrict__ ch_step, int* __restrict__ ch_fm) {
    const int n = t.n;
    for (long long bi = row_lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; bi < nb; bi += (long long)gridDim.x * blockDim.x) {
        const int box = node_ids ? node_ids[bi] : t.lnode[bi];
        int* step = ch_step + bi * n;
        double* y = ch_y + bi * n;
        double* g = ch_g + bi * n;
        int leaf = box;
        while (!(t.ninfo[leaf].w & CSP_NF_LEAF)) leaf = leaf + 1;  // split.self is the next node in preorder
        int top = -1;
        int status = chaintop_dev(t, leaf, step, &top);  // the step row doubles as chaintop's `covered`
        for (int i = 0; i < n; ++i) step[i] = 0;
        double cval = dnan();
        int Itop = -1;
        if (status == CSP_FIT_OK) {
            Itop = t.ninfo[top].y;
            const double* __restrict__ b = t.ipos + (size_t)Itop * n;
            cval = t.icval[4 * Itop];  // value(top) == value(top.split.self)
            int cur = top;
            const int nch = (n + 1) / 2;
            for (int k = 1; k <= nch; ++k) {
                const int4 cinfo = t.ninfo[cur];
                if (cinfo.w & CSP_NF_LEAF) { status = CSP_FIT_CHAIN_UNDEF; break; }
                const int I = cinfo.y;
                const int d1 = t.idims[2 * I], d2 = t.idims[2 * I + 1];
                if (d1 == d2 || d1 > n || step[d1 - 1] != 0 || (d2 <= n && step[d2 - 1] != 0)) {
                    status = CSP_FIT_CHAIN_ASSERT;
                    break;
                }
                const double x1 = t.ixs[2 * I], x2 = t.ixs[2 * I + 1];
                const double f0 = t.icval[4 * I], f1 = t.icval[4 * I + 1], f2 = t.icval[4 * I + 2];
                y[d1 - 1] = x1; step[d1 - 1] = k;
                if (ch_fm) ch_fm[bi * n + d1 - 1] = d2;
                if (d2 <= n) { y[d2 - 1] = x2; step[d2 - 1] = k; if (ch_fm) ch_fm[bi * n + d2 - 1] = d1; }
                g[d1 - 1] = (f1 - f0) / (x1 - b[d1 - 1]);
                if (d2 <= n) g[d2 - 1] = (f2 - f0) / (x2 - b[d2 - 1]);
                cur = t.ichild[4 * I + 3];
            }
        }
        ch_status[bi] = status;
        ch_itop[bi] = Itop;
        ch_c[bi] = cval;
    }
}

template <int P>
__device__ __forceinline__ double coef_value(const TreeView& t, int id) {
    if (P == 2) {  // src/polynomials.jl:212-224 with p = 2: s = +1, signs (+,-,-,+)
        const double2 xs = __ldg((const double2*)t.ixs + id), xf = __ldg((const double2*)t.ixself + id);
        const double2 f01 = __ldg((const double2*)t.icval + 2 * id), f23 = __ldg((const double2*)t.icval + 2 * id + 1);
        double denom = 1.0;
        denom = denom * (xs.x - xf.x);
        denom = denom * (xs.y - xf.y);
        double num = 0.0;
        num = num + 1.0 * f01.x;
        num = num + -1.0 * f01.y;
        num = num + -1.0 * f23.x;
        num = num + 1.0 * f23.y;
        return num / denom;
    } else {  // p = 1: s = -1, signs (-,+)
        double denom = 1.0;
        denom = denom * (t.ixs[id] - t.ixself[id]);
        double num = 0.0;
        num = num + -1.0 * t.icval[2 * id + 0];
        num = num + 1.0 * t.icval[2 * id + 1];
        return num / denom;
    }
}

// The coefficient a split contributes (src/polynomials.jl:212-224) depends on the split alone: evaluated once per tree, so the
// per-box work is a gather of 8 bytes instead of 64 bytes and a division per slot (same operations, same bits).
template <int P>
__global__ void __launch_bounds__(256) k_coef_table(const TreeView t, double* __restrict__ out) {
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < t.n_internal; id += gridDim.x * blockDim.x) out[id] = coef_value<P>(t, id);
}

// The splits(box) traversal, flattened: for a window of up to 32 consecutive range groups (one lane per ancestor) the
// group sizes are prefix-summed, and each 32-split chunk of the window's virtual sequence maps its lanes back to
// (group, offset) by bisection -- the many one-split groups next to a leaf share a chunk instead of taking one each.
// DIAG == false: first occurrence per off-diagonal pair slot (coefficients_p).  DIAG == true: first qualifying split per
// diagonal entry (fit_poly2_diag!).  Returns through tpos the number of positions walked and through found the slots filled.
template <int P, bool DIAG, bool BIG, bool MULTI = false>
__device__ __forceinline__ void fit_scan(const FitParams& prm, const WarpCtx& w, int box, int nslots, int& tpos, int& found, int& sw_cnt, int& sw_total) {
    const TreeView& t = prm.t;
    const int lane = w.lane, n = w.n, skip = prm.skip;
    const unsigned FULL = 0xffffffffu;
    tpos = 0; found = 0;
    if (nslots <= 0) return;
    const int4 binfo0 = t.ninfo[box];
    int cI = binfo0.y, cS = binfo0.z;                  // (first internal id, internal count) of the child we climb from
    const int lf = t.nleaf[box];
    const bool use_path = t.lpath != nullptr && lf >= 0;   // leaves: ancestor (I, S) pairs precomputed at upload
    const long long pbase = use_path ? t.lpoff[lf] : 0;
    const int pdepth = use_path ? (int)(t.lpoff[lf + 1] - pbase) : 0;
    int pk = 0;
    int wnode = box;  // lane 0's position in the fallback walk
    int4 winfo = binfo0;
    bool pseudo = !(binfo0.w & CSP_NF_LEAF);  // an internal box yields its own split and subtree first
    bool more = true, done = false;
    // sw_cnt > 0: an earlier scan of the same box found that its whole traversal is ONE window, still in shared memory
    const bool reuse = sw_cnt > 0;
    while (more && !done) {
        int cnt, Ia = 0, Sa = 0;
        if (reuse) {
            cnt = sw_cnt;
            more = false;
        } else if (pseudo) {
            cnt = 1;
        } else if (use_path) {
            cnt = min(32, pdepth - pk);
            if (lane < cnt) { const int2 a = __ldg(t.lpath + pbase + pk + lane); Ia = a.x; Sa = a.y; }
            pk += cnt;
            more = pk < pdepth;
        } else {
            if (lane == 0) {
                int c = 0;
                while (c < 32 && winfo.x != wnode) {
                    wnode = winfo.x;
                    winfo = t.ninfo[wnode];
                    w.g_lo1[c] = winfo.y; w.g_n1[c] = winfo.z;
                    ++c;
                }
                w.g_pre[0] = c; w.g_pre[1] = winfo.x != wnode;
            }
            __syncwarp();
            cnt = w.g_pre[0]; more = w.g_pre[1] != 0;
            if (lane < cnt) { Ia = w.g_lo1[lane]; Sa = w.g_n1[lane]; }
            __syncwarp();
        }
        if (cnt == 0) break;
        int total;
        if (reuse) {
            total = sw_total;
        } else {
        int lo1, n1, lo2, tot;
        if (pseudo) {
            lo1 = cI; n1 = cS; lo2 = 0; tot = lane == 0 ? cS : 0;
        } else {
            int Ic = __shfl_up_sync(FULL, Ia, 1), Sc = __shfl_up_sync(FULL, Sa, 1);
            if (lane == 0) { Ic = cI; Sc = cS; }
            lo1 = Ia + 1; n1 = Ic - lo1; lo2 = Ic + Sc;
            tot = lane < cnt ? n1 + (Ia + Sa - lo2) + 1 : 0;
            cI = __shfl_sync(FULL, Ia, cnt - 1); cS = __shfl_sync(FULL, Sa, cnt - 1);
        }
        int pre = tot;  // inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += u; }
        total = __shfl_sync(FULL, pre, 31);
        if (lane < cnt) { w.g_lo1[lane] = lo1; w.g_n1[lane] = n1; w.g_lo2[lane] = lo2; w.g_lo3[lane] = Ia; w.g_pre[lane] = pre - tot; }
        if (lane == cnt - 1) w.g_pre[cnt] = total;
        __syncwarp();
        if (!pseudo && !more && tpos == 0) { sw_cnt = cnt; sw_total = total; }  // first and only window
        }
        bool fast = false;
        if constexpr (!DIAG && BIG && P == 2) {
            if (w.plist) {  // global pair table: four chunks per round, their loads and atomics all in flight together
                fast = true;
                constexpr int KU = 4;
                for (int base = 0; base < total; base += 32 * KU) {
                    int idv[KU], pidx[KU];
                    bool newp[KU];
#pragma unroll
                    for (int u = 0; u < KU; ++u) {
                        const int v = base + 32 * u + lane;
                        idv[u] = -1; pidx[u] = -1; newp[u] = false;
                        if (v < total && tpos + 32 * u + lane >= skip) {
                            int gsel = 0;
#pragma unroll
                            for (int sft = 16; sft; sft >>= 1) { const int c = gsel + sft; if (c < cnt && w.g_pre[c] <= v) gsel = c; }
                            const int off = v - w.g_pre[gsel], gtot = w.g_pre[gsel + 1] - w.g_pre[gsel], gn1 = w.g_n1[gsel];
                            idv[u] = off < gn1 ? w.g_lo1[gsel] + off : ((pseudo || off < gtot - 1) ? w.g_lo2[gsel] + (off - gn1) : w.g_lo3[gsel]);
                        }
                    }
                    int2 ddv[KU];
#pragma unroll
                    for (int u = 0; u < KU; ++u) ddv[u] = idv[u] >= 0 ? __ldg((const int2*)t.idims + idv[u]) : make_int2(n + 1, n + 1);
#pragma unroll
                    for (int u = 0; u < KU; ++u) {
                        const int2 dd = ddv[u];
                        if (dd.x <= n && dd.y <= n && dd.x != dd.y) {
                            const int i = min(dd.x, dd.y) - 1, j = max(dd.x, dd.y) - 1;
                            pidx[u] = prm.slotmap ? __ldg(prm.slotmap + i * n + j) : i * n - i * (i + 1) / 2 + (j - i - 1);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < KU; ++u)
                        if (pidx[u] >= 0)
                            newp[u] = atomicMin(&w.pk64[pidx[u]], ((unsigned long long)(unsigned)(tpos + 32 * u + lane) << 32) | (unsigned)idv[u]) == ~0ull;
#pragma unroll
                    for (int u = 0; u < KU; ++u) {
                        const unsigned nb = __ballot_sync(FULL, newp[u]);
                        if (newp[u]) w.plist[found + __popc(nb & ((1u << lane) - 1))] = pidx[u];
                        found += __popc(nb);
                    }
                    tpos += min(32 * KU, total - base);
                    if (found >= nslots) { done = true; break; }
                }
            }
        }
        if constexpr (DIAG && BIG && P == 2) {  // large n: the diagonal candidates of four chunks per round, loads in flight together
            fast = true;
            constexpr int KU = 4;
            for (int base = 0; base < total; base += 32 * KU) {
                int idv[KU];
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    const int v = base + 32 * u + lane;
                    idv[u] = -1;
                    if (v < total && tpos + 32 * u + lane >= skip) {
                        int gsel = 0;
#pragma unroll
                        for (int sft = 16; sft; sft >>= 1) { const int c = gsel + sft; if (c < cnt && w.g_pre[c] <= v) gsel = c; }
                        const int off = v - w.g_pre[gsel], gtot = w.g_pre[gsel + 1] - w.g_pre[gsel], gn1 = w.g_n1[gsel];
                        idv[u] = off < gn1 ? w.g_lo1[gsel] + off : ((pseudo || off < gtot - 1) ? w.g_lo2[gsel] + (off - gn1) : w.g_lo3[gsel]);
                    }
                }
                int2 ddv[KU];
                double2 xsv[KU], xfv[KU];
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    ddv[u] = make_int2(n + 1, n + 1); xsv[u] = make_double2(0.0, 0.0); xfv[u] = xsv[u];
                    if (idv[u] >= 0) {
                        ddv[u] = __ldg((const int2*)t.idims + idv[u]);
                        xsv[u] = __ldg((const double2*)t.ixs + idv[u]);
                        xfv[u] = __ldg((const double2*)t.ixself + idv[u]);
                    }
                }
                int dk0[KU], dk1[KU];
                bool nd0[KU], nd1[KU];
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    const int pos = tpos + 32 * u + lane, dd0 = ddv[u].x, dd1 = ddv[u].y;
                    dk0[u] = -1; dk1[u] = -1; nd0[u] = false; nd1[u] = false;
                    if (dd0 <= n && w.dkey[dd0 - 1] > 2 * pos &&
                        (newinfo(xsv[u].x, w.sb[dd0 - 1], w.sy[dd0 - 1]) || newinfo(xfv[u].x, w.sb[dd0 - 1], w.sy[dd0 - 1]))) {
                        dk0[u] = 2 * pos;
                        nd0[u] = atomicMin(&w.dkey[dd0 - 1], dk0[u]) == INT_MAX;
                    }
                    if (dd1 <= n && w.dkey[dd1 - 1] > 2 * pos + 1 &&
                        (newinfo(xsv[u].y, w.sb[dd1 - 1], w.sy[dd1 - 1]) || newinfo(xfv[u].y, w.sb[dd1 - 1], w.sy[dd1 - 1]))) {
                        dk1[u] = 2 * pos + 1;
                        nd1[u] = atomicMin(&w.dkey[dd1 - 1], dk1[u]) == INT_MAX;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    if (dk0[u] >= 0 && w.dkey[ddv[u].x - 1] == dk0[u]) w.did[ddv[u].x - 1] = 2 * idv[u];
                    if (dk1[u] >= 0 && w.dkey[ddv[u].y - 1] == dk1[u]) w.did[ddv[u].y - 1] = 2 * idv[u] + 1;
                    found += __popc(__ballot_sync(FULL, nd0[u])) + __popc(__ballot_sync(FULL, nd1[u]));
                }
                tpos += min(32 * KU, total - base);
                if (found >= nslots) { done = true; break; }
            }
        }
        for (int base = 0; !fast && base < total; base += 32) {
            const int v = base + lane;
            int id = 0;
            if (v < total) {
                int gsel = 0;
#pragma unroll
                for (int sft = 16; sft; sft >>= 1) { const int c = gsel + sft; if (c < cnt && w.g_pre[c] <= v) gsel = c; }
                const int off = v - w.g_pre[gsel], gtot = w.g_pre[gsel + 1] - w.g_pre[gsel], gn1 = w.g_n1[gsel];
                id = off < gn1 ? w.g_lo1[gsel] + off : ((pseudo || off < gtot - 1) ? w.g_lo2[gsel] + (off - gn1) : w.g_lo3[gsel]);
            }
            const int pos = tpos + lane;
            const bool act = v < total && pos >= skip;
            if (!DIAG) {
                int pidx = -1;
                bool newp = false;
                if (act) {
                    if (P == 2) {
                        const int2 dd = __ldg((const int2*)t.idims + id);
                        if (dd.x <= n && dd.y <= n && dd.x != dd.y) {
                            const int i = min(dd.x, dd.y) - 1, j = max(dd.x, dd.y) - 1;
                            pidx = prm.slotmap ? __ldg(prm.slotmap + i * n + j) : i * n - i * (i + 1) / 2 + (j - i - 1);
                            if (pidx >= 0) {
                                if (BIG && w.plist) newp = atomicMin(&w.pk64[pidx], ((unsigned long long)(unsigned)pos << 32) | (unsigned)id) == ~0ull;
                                else newp = atomicMin(&w.pkey[pidx], pos) == INT_MAX;
                            }
                        }
                    } else {
                        const int dd0 = __ldg(t.idims + id);
                        if (dd0 <= n) {
                            pidx = dd0 - 1;
                            if (BIG && w.plist) newp = atomicMin(&w.pk64[pidx], ((unsigned long long)(unsigned)pos << 32) | (unsigned)id) == ~0ull;
                            else newp = atomicMin(&w.pkey[pidx], pos) == INT_MAX;
                        }
                    }
                }
                __syncwarp();
                if (!(BIG && w.plist) && pidx >= 0 && w.pkey[pidx] == pos) w.pid[pidx] = id;  // exactly one lane per slot holds the minimum
                const unsigned nb = __ballot_sync(FULL, newp);
                if (BIG && w.plist && newp) w.plist[found + __popc(nb & ((1u << lane) - 1))] = pidx;
                found += __popc(nb);
            } else if (MULTI) {
                // the same test against every member slot: the candidates (split, dimension) are common to the sibling leaves,
                // what counts as new information depends on each leaf's own b and y
                int dd0 = 0, dd1 = 0, newcnt = 0;
                unsigned hit0 = 0, hit1 = 0;
                if (act) {
                    const int2 dd = __ldg((const int2*)t.idims + id);
                    dd0 = dd.x; dd1 = dd.y;
                    const double2 xs = __ldg((const double2*)t.ixs + id);
                    const double2 xf = __ldg((const double2*)t.ixself + id);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        if (!((w.mmask >> m) & 1u)) continue;
                        const double *sbm = w.mD(m), *sym = w.mD(m) + n;
                        int* dkm = w.mI(m) + n;
                        if (dd0 <= n && dkm[dd0 - 1] > 2 * pos &&
                            (newinfo(xs.x, sbm[dd0 - 1], sym[dd0 - 1]) || newinfo(xf.x, sbm[dd0 - 1], sym[dd0 - 1]))) {
                            hit0 |= 1u << m;
                            newcnt += atomicMin(&dkm[dd0 - 1], 2 * pos) == INT_MAX;
                        }
                        if (dd1 <= n && dkm[dd1 - 1] > 2 * pos + 1 &&
                            (newinfo(xs.y, sbm[dd1 - 1], sym[dd1 - 1]) || newinfo(xf.y, sbm[dd1 - 1], sym[dd1 - 1]))) {
                            hit1 |= 1u << m;
                            newcnt += atomicMin(&dkm[dd1 - 1], 2 * pos + 1) == INT_MAX;
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    if ((hit0 >> m) & 1u) { int* dkm = w.mI(m) + n; if (dkm[dd0 - 1] == 2 * pos) dkm[n + dd0 - 1] = 2 * id; }
                    if ((hit1 >> m) & 1u) { int* dkm = w.mI(m) + n; if (dkm[dd1 - 1] == 2 * pos + 1) dkm[n + dd1 - 1] = 2 * id + 1; }
                }
                found += __reduce_add_sync(FULL, newcnt);
            } else {
                int dk0 = -1, dk1 = -1, dd0 = 0, dd1 = 0;
                bool newd0 = false, newd1 = false;
                if (act) {
                    const int2 dd = __ldg((const int2*)t.idims + id);
                    dd0 = dd.x; dd1 = dd.y;
                    const double2 xs = __ldg((const double2*)t.ixs + id);
                    const double2 xf = __ldg((const double2*)t.ixself + id);
                    if (dd0 <= n && w.dkey[dd0 - 1] > 2 * pos &&
                        (newinfo(xs.x, w.sb[dd0 - 1], w.sy[dd0 - 1]) || newinfo(xf.x, w.sb[dd0 - 1], w.sy[dd0 - 1]))) {
                        dk0 = 2 * pos;
                        newd0 = atomicMin(&w.dkey[dd0 - 1], dk0) == INT_MAX;
                    }
                    if (dd1 <= n && w.dkey[dd1 - 1] > 2 * pos + 1 &&
                        (newinfo(xs.y, w.sb[dd1 - 1], w.sy[dd1 - 1]) || newinfo(xf.y, w.sb[dd1 - 1], w.sy[dd1 - 1]))) {
                        dk1 = 2 * pos + 1;
                        newd1 = atomicMin(&w.dkey[dd1 - 1], dk1) == INT_MAX;
                    }
                }
                __syncwarp();
                if (dk0 >= 0 && w.dkey[dd0 - 1] == dk0) w.did[dd0 - 1] = 2 * id;
                if (dk1 >= 0 && w.dkey[dd1 - 1] == dk1) w.did[dd1 - 1] = 2 * id + 1;
                found += __popc(__ballot_sync(FULL, newd0)) + __popc(__ballot_sync(FULL, newd1));
            }
            tpos += min(32, total - base);
            if (found >= nslots) { done = true; break; }
        }
        __syncwarp();
        if (pseudo) { pseudo = false; more = true; }  // now climb from the box itself
    }
    __syncwarp();
}

// One warp per work unit.  A unit is one box, or -- for the all-leaf model fits (prm.units) -- the real leaves that are
// children of one split: their splits() traversals are identical (each leaf child has an empty subtree, so the two sibling
// ranges of the parent's group always join to the parent's whole subtree), hence so are their off-diagonal coefficients;
// the pair scan and the coefficient values are computed once and only fit_poly1 / fit_poly2_diag! / canonicalize run per leaf.
template <int P, bool BIG, bool MULTI = false>  // BIG: pair tables in global scratch and/or Q in global memory (large n); MULTI: prm.multi > 1
__global__ void __launch_bounds__(256, BIG ? 2 : (MULTI ? 3 : 4)) k_fit(const FitParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TreeView& t = prm.t;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int n = t.n;
    const int npairs = prm.slotmap ? prm.npat : (P == 2 ? n * (n - 1) / 2 : n);
    const unsigned FULL = 0xffffffffu;

    unsigned char* wbase = smem_raw + (size_t)wib * prm.warp_smem_bytes;
    WarpCtx w;
    w.lane = lane; w.n = n; w.npairs = npairs;
    w.sb = (double*)wbase; w.sy = w.sb + n; w.sg = w.sy + n;
    const long long qlen = prm.slotmap ? prm.npat : (P == 1 ? n : (prm.q_packed ? (long long)(n + 1) * (n + 2) / 2 : (long long)n * n));
    double* sQ = (double*)(w.sg + n);  // [qlen] when q_in_smem
    w.g_lo1 = (int*)(sQ + (prm.q_in_smem ? qlen : 0)); w.g_n1 = w.g_lo1 + 32; w.g_lo2 = w.g_n1 + 32; w.g_lo3 = w.g_lo2 + 32; w.g_pre = w.g_lo3 + 32;
    w.sstep = w.g_lo1 + FIT_WINDOW_INTS; w.dkey = w.sstep + n; w.did = w.dkey + n;
    if (prm.pairs_in_smem) { w.pkey = w.did + n; w.pid = w.pkey + npairs; }
    else { w.pkey = prm.scratch + ((size_t)blockIdx.x * W + wib) * 3 * (size_t)((npairs + 1) & ~1); w.pid = w.pkey + npairs; }
    w.plist = (!BIG || prm.pairs_in_smem) ? nullptr : w.pkey + 2 * (size_t)((npairs + 1) & ~1);
    w.mmask = 0; w.mD0 = nullptr; w.mI0 = nullptr; w.mSD = 0;
    w.pk64 = (unsigned long long*)w.pkey;  // (8-byte aligned: the per-warp slice is a multiple of 8 bytes)
    if (w.plist) {  // the table is cleared once; afterwards every unit resets just the slots it filled
        for (int e = lane; e < npairs; e += 32) w.pk64[e] = ~0ull;
        __syncwarp();
    }
    // staging tiles sit at the end of the warp's slice (only allocated when prm.tiles)
    double* tq = (double*)(wbase + prm.warp_smem_bytes - FIT_TILE_BYTES);
    double* tx = tq + 32 * (FIT_TKB + 1);
    int* s_id = (int*)(tx + 32 * (FIT_TKB + 1));
    const QIndex qi{n, prm.q_packed};
    const bool ranged = prm.row_lo > 0 || prm.row_hi < prm.nb;  // a leaf-range shard of an all-leaf fit
    const long long nunits = prm.units ? prm.n_units : prm.row_hi - prm.row_lo;

    long long tk = clock64();
#define FIT_TICK(k) do { if (prm.prof) { __syncwarp(); if (lane == 0) { const long long now_ = clock64(); atomicAdd(prm.prof + (k), (unsigned long long)(now_ - tk)); tk = now_; } } } while (0)
    for (long long ui = (long long)blockIdx.x * W + wib;; ui += (long long)gridDim.x * W) {
        if (prm.work_counter) {  // units differ in cost (depth, subtree sizes): claim them one at a time
            unsigned long long c = 0;
            if (lane == 0) c = atomicAdd(prm.work_counter, 1ull);
            ui = (long long)__shfl_sync(FULL, c, 0);
        }
        if (ui >= nunits) break;
        int4 members = make_int4(-1, -1, -1, -1);  // box indices (output rows) of the unit
        if (prm.units) {
            members = prm.units[ui];
            if (ranged) {  // keep the members this shard owns, in front (any member can lead: their traversals coincide)
                int mm[4] = {members.x, members.y, members.z, members.w}, k = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (mm[u] >= prm.row_lo && mm[u] < prm.row_hi) mm[k++] = mm[u];
                for (int u = k; u < 4; ++u) mm[u] = -1;
                members = make_int4(mm[0], mm[1], mm[2], mm[3]);
                if (k == 0) continue;  // a unit of another shard
            }
        }
        const long long bi0 = prm.units ? members.x : prm.row_lo + ui;
        const int box = prm.node_ids ? prm.node_ids[bi0] : t.lnode[bi0];
        double* Qo = prm.q_in_smem ? sQ : prm.Q + bi0 * prm.Q_stride;  // global-Q mode: assembled in the first member's record
        for (long long e = lane; e < qlen; e += 32) Qo[e] = dnan();
        if (!w.plist) for (int e = lane; e < npairs; e += 32) w.pkey[e] = INT_MAX;
        __syncwarp();
        FIT_TICK(0);

        // ---- coefficients_p (src/polynomials.jl:182-230): first split per off-diagonal slot, then its value
        int tpos, foundp, sw_cnt = 0, sw_total = 0;
        fit_scan<P, false, BIG>(prm, w, box, npairs, tpos, foundp, sw_cnt, sw_total);
        FIT_TICK(1);
        long long nvisited, nused = foundp;  // every filled slot was set by exactly one split
        {   // number of splits coefficients_p examined (after `skip`): up to the one that filled the last slot
            int mx = -1;
            if (w.plist) {  // four independent gathers in flight per lane
                for (int e0 = lane; e0 < foundp; e0 += 128) {
                    int ev[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) ev[u] = e0 + 32 * u < foundp ? w.plist[e0 + 32 * u] : -1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (ev[u] >= 0) mx = max(mx, (int)(w.pk64[ev[u]] >> 32));
                }
            } else for (int e = lane; e < npairs; e += 32) mx = max(mx, w.pkey[e] == INT_MAX ? -1 : w.pkey[e]);
            for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
            nvisited = (foundp >= npairs && npairs > 0) ? (long long)mx + 1 - prm.skip : max(0, tpos - prm.skip);
        }
        {
            // values (src/polynomials.jl:211-226), one lane per slot
            bool sawnan = false;
            const int nslot = w.plist ? foundp : npairs;
            if (BIG && w.plist) {
                // global table: slot -> (winning split) -> value -> output, four slots per lane per round so that the dependent
                // gathers of a round overlap; the slots are handed back empty in the same pass
                for (int e0 = lane; e0 < nslot; e0 += 128) {
                    int ev[4], sid[4];
                    double vv[4];
                    int2 ddv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) ev[u] = e0 + 32 * u < nslot ? w.plist[e0 + 32 * u] : -1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) sid[u] = ev[u] >= 0 ? (int)(w.pk64[ev[u]] & 0xffffffffu) : -1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        vv[u] = 0.0; ddv[u] = make_int2(1, 1);
                        if (sid[u] >= 0) {
                            vv[u] = prm.coef ? __ldg(prm.coef + sid[u]) : (P == 2 ? coef_value<2>(t, sid[u]) : coef_value<1>(t, sid[u]));
                            if (P == 2 && !prm.slotmap) ddv[u] = __ldg((const int2*)t.idims + sid[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (sid[u] < 0) continue;
                        sawnan |= (vv[u] != vv[u]);
                        if (P == 2 && !prm.slotmap) Qo[qi(ddv[u].x - 1, ddv[u].y - 1)] = vv[u];
                        else Qo[ev[u]] = vv[u];
                        w.pk64[ev[u]] = ~0ull;
                    }
                }
            } else if (P == 2 && prm.slotmap) {
                for (int e0 = lane; e0 < nslot; e0 += 32) {
                    const int e = w.plist ? w.plist[e0] : e0;
                    if (!w.plist && w.pkey[e] == INT_MAX) continue;
                    const int sid = w.plist ? (int)(w.pk64[e] & 0xffffffffu) : w.pid[e];
                    const double v = prm.coef ? __ldg(prm.coef + sid) : coef_value<2>(t, sid);
                    sawnan |= (v != v);
                    Qo[e] = v;
                }
            } else if (P == 2) {
                for (int e0 = lane; e0 < nslot; e0 += 32) {  // one lane per pair slot; (i, j) are the winning split's dimensions
                    const int e = w.plist ? w.plist[e0] : e0;
                    if (!w.plist && w.pkey[e] == INT_MAX) continue;
                    const int id = w.plist ? (int)(w.pk64[e] & 0xffffffffu) : w.pid[e];
                    const int2 dd = __ldg((const int2*)t.idims + id);
                    const double v = prm.coef ? __ldg(prm.coef + id) : coef_value<2>(t, id);
                    sawnan |= (v != v);
                    Qo[qi(dd.x - 1, dd.y - 1)] = v;
                }
            } else {
                for (int e0 = lane; e0 < nslot; e0 += 32) {
                    const int e = w.plist ? w.plist[e0] : e0;
                    if (!w.plist && w.pkey[e] == INT_MAX) continue;
                    const int sid = w.plist ? (int)(w.pk64[e] & 0xffffffffu) : w.pid[e];
                    const double v = prm.coef ? __ldg(prm.coef + sid) : coef_value<1>(t, sid);
                    sawnan |= (v != v);
                    Qo[e] = v;
                }
            }
            sawnan = __any_sync(FULL, sawnan) || npairs == 0;  // nothing to determine: replay the reference's loop literally
            __syncwarp();
            FIT_TICK(2);
            if (sawnan) {
                // A computed coefficient was NaN (non-finite samples or a vanishing denominator).  The reference then keeps
                // overwriting that slot from later splits while still counting it towards `nremaining`; reproduce that
                // sequential behaviour exactly with one lane (rare path).
                if (lane == 0) {
                    for (long long e = 0; e < qlen; ++e) Qo[e] = dnan();
                    long long nrem = npairs, tp = 0, examined = 0, nset = 0;
                    for_split_groups(t, box, [&](const RangeGroup& grp) -> bool {
                        const int total = grp.total();
                        for (int v = 0; v < total; ++v, ++tp) {
                            const int id = grp.id(v);
                            if (tp < prm.skip) continue;
                            ++examined;
                            long long slot;
                            if (P == 2) {
                                const int d0 = t.idims[2 * id], d1 = t.idims[2 * id + 1];
                                if (d0 > n || d1 > n) continue;
                                if (prm.slotmap) {  // a structural zero reads 0.0, not NaN: nothing is set, but the exit test below runs
                                    slot = d0 == d1 ? -1 : prm.slotmap[(min(d0, d1) - 1) * n + max(d0, d1) - 1];
                                } else {
                                    slot = qi(d0 - 1, d1 - 1);
                                }
                            } else {
                                const int d0 = t.idims[id];
                                if (d0 > n) continue;
                                slot = d0 - 1;
                            }
                            if (slot >= 0) {
                                const double cur = Qo[slot];
                                if (cur != cur) {
                                    Qo[slot] = P == 2 ? coef_value<2>(t, id) : coef_value<1>(t, id);
                                    nrem -= 1;
                                    nset += 1;
                                }
                            }
                            if (nrem <= 0) return true;
                        }
                        return false;
                    });
                    nvisited = examined;
                    nused = nset;
                }
                nvisited = __shfl_sync(FULL, nvisited, 0);
                nused = __shfl_sync(FULL, nused, 0);
                __syncwarp();
            }

        }

        if (prm.units && !prm.q_in_smem) {  // large records: hand the shared off-diagonals to the other members' records now
            __syncwarp();
#pragma unroll 1
            for (int mi = 1; mi < 4; ++mi) {
                const long long bm = mi == 1 ? members.y : mi == 2 ? members.z : members.w;
                if (bm < 0) break;
                double* Qm = prm.Q + bm * prm.Q_stride;
                for (long long e = lane; e < qlen; e += 256) {  // eight loads in flight, then eight stores
                    double tv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) tv[u] = e + 32 * u < qlen ? Qo[e + 32 * u] : 0.0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (e + 32 * u < qlen) Qm[e + 32 * u] = tv[u];
                }
            }
            __syncwarp();
        }

        // ---- small n, all-leaf model fits: the per-leaf phases of up to `multi` sibling leaves side by side.  Lane l works for
        // slot l / n on entry l % n; every slot has private b, y, g, step, dkey, did and its own copy of the record, started
        // from the shared off-diagonals in Qo.  The arithmetic per entry is the single-member code below, unchanged.
        if constexpr (MULTI) {
            {
                const int G = prm.multi;
                const int SD = 3 * n + (int)qlen;  // doubles per slot: b | y | g | record
                double* xD = (double*)(wbase + prm.multi_off);
                int* xI = (int*)(xD + (size_t)G * SD);  // ints per slot: step | dkey | did
                w.mD0 = xD; w.mI0 = xI; w.mSD = SD;
                const int slot = lane / n, e = lane - slot * n;  // this lane's slot and entry
#pragma unroll 1
                for (int m0 = 0; m0 < 4; m0 += G) {
                    const long long bfirst = m0 == 0 ? members.x : m0 == 1 ? members.y : m0 == 2 ? members.z : members.w;
                    if (bfirst < 0) break;
                    long long bis[4];
                    int gm = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int mi = m0 + u;
                        bis[u] = (u < G && mi < 4) ? (mi == 0 ? members.x : mi == 1 ? members.y : mi == 2 ? members.z : members.w) : -1;
                        if (bis[u] >= 0 && gm == u) gm = u + 1;  // members are packed in front
                    }
                    const bool mine = slot < gm && lane < G * n;
                    long long bi = -1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (slot == u) bi = bis[u];
                    double* Dm = xD + (size_t)(mine ? slot : 0) * SD;
                    int* Im = xI + (size_t)(mine ? slot : 0) * 3 * n;
                    double *sbm = Dm, *sym = Dm + n, *sgm = Dm + 2 * n, *Qm = Dm + 3 * n;
                    int *stm = Im, *dkm = Im + n, *dim_ = Im + 2 * n;
                    // every slot's record starts as the unit's shared off-diagonals (diagonal, g column and c still NaN)
                    for (int u = 0; u < gm; ++u)
                        for (long long q = lane; q < qlen; q += 32) xD[(size_t)u * SD + 3 * n + q] = Qo[q];
                    int status = CSP_FIT_OK;
                    double cval = dnan();
                    if (mine) {
                        dkm[e] = INT_MAX; stm[e] = 0;
                        status = prm.ch_status[bi];
                        if (status == CSP_FIT_OK) {
                            const int Itop = prm.ch_itop[bi];
                            cval = prm.ch_c[bi];
                            sbm[e] = t.ipos[(size_t)Itop * n + e];
                            sym[e] = prm.ch_y[bi * n + e];
                            sgm[e] = prm.ch_g[bi * n + e];
                            stm[e] = prm.ch_step[bi * n + e];
                        }
                    }
                    const bool ok = mine && status == CSP_FIT_OK;
                    unsigned okslots = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) okslots |= (__ballot_sync(FULL, ok && slot == u) != 0u ? 1u : 0u) << u;
                    __syncwarp();
                    if (okslots) {
                        w.mmask = okslots;
                        int dpos, foundd;
                        fit_scan<P, true, BIG, true>(prm, w, box, n * __popc(okslots), dpos, foundd, sw_cnt, sw_total);
                        w.mmask = 0;
                    }
                    // ---- fit_poly2_diag! (src/cs2.jl:208-274): one lane per (slot, diagonal entry), sequential over k
                    if (ok && dkm[e] != INT_MAX) {
                        const int d = e;
                        const int code = dim_[d];
                        const int id = code >> 1, dimidx = code & 1;
                        const double* xpos = t.ipos + (size_t)id * n;
                        const double bd = sbm[d], yd = sym[d];
                        const double xsd = t.ixs[2 * id + dimidx];
                        const int d2 = t.idims[2 * id + (1 - dimidx)] - 1;
                        const double xs2 = t.ixs[2 * id + (1 - dimidx)];
                        double xnew, xd, fchild, fref;
                        bool swapped;
                        if (newinfo(xsd, bd, yd)) {
                            swapped = false;
                            xnew = xsd; xd = xpos[d];
                            fchild = t.icval[4 * id + (1 + dimidx)];
                            fref = t.icval[4 * id];
                        } else {
                            swapped = true;
                            xnew = xpos[d]; xd = xsd;
                            fchild = t.icval[4 * id + (2 - dimidx)];
                            fref = t.icval[4 * id + 3];
                        }
                        const double dxi = xnew - xd;
                        const double coef = qcoef_lineq(true, dxi, xd, bd, yd, true);
                        double consts = -sgm[d];
                        const int sd = stm[d];
                        QWalk qw(d, n, prm.q_packed);
#pragma unroll 4
                        for (int k = 0; k < n; ++k) {
                            double xk = __ldg(xpos + k);
                            if (swapped && k == d2) xk = xs2;
                            const bool precdk = sd <= stm[k];
                            const double xc = qcoef_lineq(false, dxi, xk, sbm[k], sym[k], precdk);
                            const double term = Qm[qw.idx] * xc;
                            qw.step(k);
                            if (k != d) consts = consts - term;
                        }
                        Qm[qi(d, d)] = (consts + (fchild - fref) / (xnew - xd)) / coef;
                    }
                    __syncwarp();
                    // ---- canonicalize (src/cs2.jl:132-140) and outputs
                    if (mine) {
                        const int i = e;
                        double gc = dnan(), bb = dnan();
                        if (ok) {
                            gc = sgm[i];
                            const int si = stm[i];
                            QWalk qw(i, n, prm.q_packed);
                            for (int j = 0; j < n; ++j) {
                                const double chi = (i == j) ? sym[j] : (si <= stm[j] ? sbm[j] : 2 * sym[j] - sbm[j]);
                                gc = gc + (sbm[j] - chi) * Qm[qw.idx] / 2;
                                qw.step(j);
                            }
                            bb = sbm[i];
                        }
                        if (prm.q_packed) {
                            Qm[qi(i, n)] = gc;
                            if (ok) Qm[qi(i, i)] = Qm[qi(i, i)] * 0.5;
                        } else if (prm.g) {
                            prm.g[bi * prm.g_stride + i] = gc;
                        }
                        if (prm.b) prm.b[bi * prm.b_stride + i] = bb;
                        if (prm.gnat) prm.gnat[bi * n + i] = ok ? sgm[i] : dnan();
                        if (prm.y) prm.y[bi * n + i] = ok ? sym[i] : dnan();
                        if (prm.firstmatch) prm.firstmatch[bi * n + i] = ok ? prm.ch_fm[bi * n + i] : 0;
                        if (e == 0) {
                            if (prm.q_packed) Qm[qi(n, n)] = ok ? cval : dnan();
                            else if (prm.c) prm.c[bi * prm.c_stride] = ok ? cval : dnan();
                            if (prm.status) prm.status[bi] = (uint8_t)status;
                            if (prm.visited) prm.visited[bi] = nvisited;
                            if (prm.used) prm.used[bi] = nused;
                        }
                    }
                    __syncwarp();
                    // records out, coalesced, one slot after the other (a failed chain leaves an all-NaN record)
                    for (int u = 0; u < gm; ++u) {
                        const bool uok = (okslots >> u) & 1u;
                        double* Qg = prm.Q + bis[u] * prm.Q_stride;
                        const double* Qs = xD + (size_t)u * SD + 3 * n;
                        for (long long q = lane; q < qlen; q += 32) Qg[q] = uok ? Qs[q] : dnan();
                    }
                    __syncwarp();
                }
                FIT_TICK(4);
            }
        }
        // ---- per member box: fit_poly1 results, fit_poly2_diag!, canonicalize, outputs (not in the MULTI instantiation)
#pragma unroll 1
        for (int mi = 0; mi < (MULTI ? 0 : 4); ++mi) {
            const long long bi = prm.units ? (mi == 0 ? members.x : mi == 1 ? members.y : mi == 2 ? members.z : members.w) : (mi == 0 ? bi0 : -1);
            if (bi < 0) break;
            double* Qg = prm.Q + bi * prm.Q_stride;
            if (!prm.q_in_smem) Qo = Qg;
            int status = CSP_FIT_OK;
            double cval = dnan();
            if (prm.mode == FIT_MODE_FULL) {
                if (mi > 0 && prm.q_in_smem)  // entries the previous member set: diagonal, g column, c
                    for (int e = lane; e < n + (prm.q_packed ? 1 : 0); e += 32) { Qo[qi(e, e)] = dnan(); if (prm.q_packed && e < n) Qo[qi(e, n)] = dnan(); }
                for (int e = lane; e < n; e += 32) { w.dkey[e] = INT_MAX; w.sstep[e] = 0; }
                // fit_poly1 (src/cs2.jl:162-205): chain results come from the pre-pass k_fit_chain
                status = prm.ch_status[bi];
                if (status == CSP_FIT_OK) {
                    const int Itop = prm.ch_itop[bi];
                    cval = prm.ch_c[bi];
                    for (int e = lane; e < n; e += 32) {
                        w.sb[e] = t.ipos[(size_t)Itop * n + e];
                        w.sy[e] = prm.ch_y[bi * n + e];
                        w.sg[e] = prm.ch_g[bi * n + e];
                        w.sstep[e] = prm.ch_step[bi * n + e];
                    }
                }
                __syncwarp();
                if (status == CSP_FIT_OK) {
                    int dpos, foundd;
                    fit_scan<P, true, BIG>(prm, w, box, n, dpos, foundd, sw_cnt, sw_total);
                    FIT_TICK(2);
                    // ---- fit_poly2_diag! (src/cs2.jl:208-274): one lane per diagonal entry, sequential over k
                    if (BIG && prm.tiles) {
                        // Large n, Q in global memory.  Lanes own d = r0 + lane; k advances in blocks of FIT_TKB.  Left of the diagonal
                        // block Q[d][k] lives at row k, column d: adjacent lanes read adjacent doubles.  From the diagonal block on it
                        // lives in row d (a stride of ~n doubles between lanes), and the split positions x[k] are a different row
                        // per lane: both are staged through shared memory, two 128-byte row segments per load instruction.
                        for (int r0 = 0; r0 < n; r0 += 32) {
                            const int d = r0 + lane;
                            const bool have = d < n && w.dkey[d] != INT_MAX;
                            int id = -1, d2 = -1, sd = 0;
                            double xs2 = 0.0, xnew = 0.0, xd = 0.0, fchild = 0.0, fref = 0.0, dxi = 0.0, coef = 0.0, consts = 0.0;
                            bool swapped = false;
                            if (have) {
                                const int code = w.did[d];
                                id = code >> 1;
                                const int dimidx = code & 1;
                                const double bd = w.sb[d], yd = w.sy[d];
                                const double xsd = t.ixs[2 * id + dimidx], xposd = t.ipos[(size_t)id * n + d];
                                d2 = t.idims[2 * id + (1 - dimidx)] - 1;
                                xs2 = t.ixs[2 * id + (1 - dimidx)];
                                if (newinfo(xsd, bd, yd)) {
                                    xnew = xsd; xd = xposd;
                                    fchild = t.icval[4 * id + (1 + dimidx)];
                                    fref = t.icval[4 * id];
                                } else {
                                    swapped = true;
                                    xnew = xposd; xd = xsd;
                                    fchild = t.icval[4 * id + (2 - dimidx)];
                                    fref = t.icval[4 * id + 3];
                                }
                                dxi = xnew - xd;
                                coef = qcoef_lineq(true, dxi, xd, bd, yd, true);
                                consts = -w.sg[d];
                                sd = w.sstep[d];
                            }
                            __syncwarp();
                            s_id[lane] = id;
                            __syncwarp();
                            for (int k0 = 0; k0 < n; k0 += FIT_TKB) {
                                const bool stageQ = k0 + FIT_TKB > r0;  // the diagonal block and everything right of it
                                {
                                    const int kk = lane & (FIT_TKB - 1), k = k0 + kk;
                                    for (int rr = lane / FIT_TKB; rr < 32; rr += 32 / FIT_TKB) {
                                        const int rid = s_id[rr];
                                        if (rid >= 0 && k < n) {  // asynchronous copies: all of a block's loads are in flight together
                                            cp_async8(tx + rr * (FIT_TKB + 1) + kk, t.ipos + (size_t)rid * n + k);
                                            if (stageQ) cp_async8(tq + rr * (FIT_TKB + 1) + kk, Qo + qi(r0 + rr, k));
                                        }
                                    }
                                }
                                if (!stageQ && d < n) {  // left of the diagonal: row k0 + kk, column d -- coalesced as is
                                    for (int kk = 0; kk < FIT_TKB; ++kk) cp_async8(tq + lane * (FIT_TKB + 1) + kk, Qo + qi(k0 + kk, d));
                                }
                                cp_async_wait_all();
                                __syncwarp();
                                if (have) {
                                    const int kend = min(FIT_TKB, n - k0);
                                    for (int kk = 0; kk < kend; ++kk) {
                                        const int k = k0 + kk;
                                        double xk = tx[lane * (FIT_TKB + 1) + kk];
                                        if (swapped && k == d2) xk = xs2;
                                        const bool precdk = sd <= w.sstep[k];
                                        const double xc = qcoef_lineq(false, dxi, xk, w.sb[k], w.sy[k], precdk);
                                        const double q = tq[lane * (FIT_TKB + 1) + kk];
                                        const double term = q * xc;
                                        if (k != d) consts = consts - term;
                                    }
                                }
                                __syncwarp();
                            }
                            if (have) Qo[qi(d, d)] = (consts + (fchild - fref) / (xnew - xd)) / coef;
                        }
                    } else
                    for (int d = lane; d < n; d += 32) {
                        if (w.dkey[d] == INT_MAX) continue;
                        const int code = w.did[d];
                        const int id = code >> 1, dimidx = code & 1;  // 0-based dimidx
                        const double* xpos = t.ipos + (size_t)id * n;
                        const double bd = w.sb[d], yd = w.sy[d];
                        const double xsd = t.ixs[2 * id + dimidx];
                        const int d2 = t.idims[2 * id + (1 - dimidx)] - 1;  // the split's other dimension (0-based; may be fictive)
                        const double xs2 = t.ixs[2 * id + (1 - dimidx)];
                        double xnew, xd, fchild, fref;
                        bool swapped;
                        if (newinfo(xsd, bd, yd)) {
                            swapped = false;
                            xnew = xsd; xd = xpos[d];
                            fchild = t.icval[4 * id + (1 + dimidx)];  // split.others.children[dimidx]
                            fref = t.icval[4 * id];                   // split.self
                        } else {
                            swapped = true;  // evaluate from children[3]'s position instead
                            xnew = xpos[d]; xd = xsd;
                            fchild = t.icval[4 * id + (2 - dimidx)];  // split.others.children[3-dimidx]
                            fref = t.icval[4 * id + 3];
                        }
                        const double dxi = xnew - xd;
                        const double coef = qcoef_lineq(true, dxi, xd, bd, yd, true);
                        double consts = -w.sg[d];
                        const int sd = w.sstep[d];
                        QWalk qw(d, n, prm.q_packed);
#pragma unroll 4
                        for (int k = 0; k < n; ++k) {  // the loads do not depend on the running sum: unrolled, they overlap
                            double xk = __ldg(xpos + k);
                            if (swapped && k == d2) xk = xs2;
                            const bool precdk = sd <= w.sstep[k];  // prec[d,k]: d split before-or-with k
                            const double xc = qcoef_lineq(false, dxi, xk, w.sb[k], w.sy[k], precdk);
                            const double term = Qo[qw.idx] * xc;
                            qw.step(k);
                            if (k != d) consts = consts - term;
                        }
                        Qo[qi(d, d)] = (consts + (fchild - fref) / (xnew - xd)) / coef;
                    }
                    __syncwarp();
                    FIT_TICK(3);
                }
            }

            // ---- outputs
            if (prm.mode == FIT_MODE_FULL) {
                const bool ok = status == CSP_FIT_OK;
                if (!ok) for (long long e = lane; e < qlen; e += 32) Qg[e] = dnan();
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    double gc = dnan(), bb = dnan();
                    if (BIG && ok && prm.tiles) {  // canonicalize with Q staged as in the diagonal solve above (whole warp takes part)
                        const bool have = i < n;
                        const int si = have ? w.sstep[i] : 0;
                        if (have) gc = w.sg[i];
                        for (int j0 = 0; j0 < n; j0 += FIT_TKB) {
                            const bool stageQ = j0 + FIT_TKB > i0;
                            if (stageQ) {
                                const int jj = lane & (FIT_TKB - 1), j = j0 + jj;
                                for (int rr = lane / FIT_TKB; rr < 32; rr += 32 / FIT_TKB)
                                    if (i0 + rr < n && j < n) cp_async8(tq + rr * (FIT_TKB + 1) + jj, Qo + qi(i0 + rr, j));
                            } else if (have) {
                                for (int jj = 0; jj < FIT_TKB; ++jj) cp_async8(tq + lane * (FIT_TKB + 1) + jj, Qo + qi(j0 + jj, i));
                            }
                            cp_async_wait_all();
                            __syncwarp();
                            if (have) {
                                const int jend = min(FIT_TKB, n - j0);
                                for (int jj = 0; jj < jend; ++jj) {
                                    const int j = j0 + jj;
                                    const double chi = (i == j) ? w.sy[j] : (si <= w.sstep[j] ? w.sb[j] : 2 * w.sy[j] - w.sb[j]);
                                    const double q = tq[lane * (FIT_TKB + 1) + jj];
                                    gc = gc + (w.sb[j] - chi) * q / 2;
                                }
                            }
                            __syncwarp();
                        }
                        if (have) bb = w.sb[i];
                    } else if (ok && i < n) {  // canonicalize (src/cs2.jl:132-140): sequential over j
                        gc = w.sg[i];
                        const int si = w.sstep[i];
                        QWalk qw(i, n, prm.q_packed);
                        for (int j = 0; j < n; ++j) {
                            const double chi = (i == j) ? w.sy[j] : (si <= w.sstep[j] ? w.sb[j] : 2 * w.sy[j] - w.sb[j]);
                            gc = gc + (w.sb[j] - chi) * Qo[qw.idx] / 2;
                            qw.step(j);
                        }
                        bb = w.sb[i];
                    }
                    if (i >= n) continue;
                    if (prm.q_packed) {  // model record: g is A's last column, the diagonal is stored halved (exact scaling)
                        Qo[qi(i, n)] = gc;
                        if (ok) Qo[qi(i, i)] = Qo[qi(i, i)] * 0.5;
                    } else if (prm.g) {
                        prm.g[bi * prm.g_stride + i] = gc;
                    }
                    if (prm.b) prm.b[bi * prm.b_stride + i] = bb;
                    if (prm.gnat) prm.gnat[bi * n + i] = ok ? w.sg[i] : dnan();
                    if (prm.y) prm.y[bi * n + i] = ok ? w.sy[i] : dnan();
                    if (prm.firstmatch) prm.firstmatch[bi * n + i] = ok ? prm.ch_fm[bi * n + i] : 0;
                }
                if (prm.prec)
                    for (int e = lane; e < n * n; e += 32) {
                        const int j = e / n, i = e - j * n;  // column-major: prec[i,j]
                        prm.prec[bi * (long long)n * n + e] = ok && (w.sstep[i] <= w.sstep[j]) ? 1 : 0;
                    }
                if (lane == 0) {
                    if (prm.q_packed) Qo[qi(n, n)] = ok ? cval : dnan();
                    else if (prm.c) prm.c[bi * prm.c_stride] = ok ? cval : dnan();
                }
            } else if (prm.mode == FIT_MODE_LIN) {
                // degree-1 model of a CS1 box: c = value(box), b = position(box), g = Cp (already in Qo == g slot)
                const int4 binfo = t.ninfo[box];
                const bool isroot = binfo.x == box;
                const int Ip = t.ninfo[binfo.x].y, ci = CSP_NF_CHILDINDEX(binfo.w);
                const bool internal = !(binfo.w & CSP_NF_LEAF);
                for (int i = lane; i < n; i += 32) {
                    double bb;
                    if (internal) bb = t.ipos[(size_t)binfo.y * n + i];
                    else if (isroot) bb = dnan();  // a tree without splits carries no positions on the device
                    else {
                        bb = t.ipos[(size_t)Ip * n + i];
                        if (P == 1) { if (ci == 1 && t.idims[Ip] - 1 == i) bb = t.ixs[Ip]; }
                        else {
                            if ((ci & 1) && t.idims[2 * Ip] - 1 == i) bb = t.ixs[2 * Ip];
                            if ((ci & 2) && t.idims[2 * Ip + 1] - 1 == i) bb = t.ixs[2 * Ip + 1];
                        }
                    }
                    if (prm.b) prm.b[bi * prm.b_stride + i] = bb;
                }
                if (lane == 0 && prm.c) {
                    double cv;
                    if (internal) cv = t.icval[(size_t)binfo.y << P];
                    else if (isroot) cv = dnan();
                    else cv = t.icval[((size_t)Ip << P) + ci];
                    prm.c[bi * prm.c_stride] = cv;
                }
            }
            if (prm.q_in_smem && !(prm.mode == FIT_MODE_FULL && status != CSP_FIT_OK)) {
                __syncwarp();
                for (long long e = lane; e < qlen; e += 32) Qg[e] = Qo[e];
            }
            if (lane == 0) {
                if (prm.status) prm.status[bi] = (uint8_t)status;
                if (prm.visited) prm.visited[bi] = nvisited;
                if (prm.used) prm.used[bi] = nused;
            }
            __syncwarp();
            FIT_TICK(4);
        }
    }
#undef FIT_TICK
}

// ---- fit_quadratic_gdiag! (reference src/cs2.jl:332-384): chain-free gradient + diagonal from two 1-d equations per dim ----
// One warp per box; Q (dense .data layout) arrives holding coefficients_p's off-diagonals.  For every dimension d the reference
// keeps the first two splits (in splits(box) order, after `skip`) that involve d, have finite data and -- for the second --
// a different xcoef than the first.  Lanes evaluate the candidates of a 32-split chunk in parallel (each an O(n) sequential
// sum in the reference's order); one lane per dimension then replays the chunk's candidates in order.
struct GdiagParams {
    TreeView t;
    const int* node_ids;
    long long nb;
    int skip;
    double* c; double* g; double* Q; double* b;  // box-major: c[nb], g[nb*n], Q[nb*n*n] (in/out), b[nb*n]
    int warp_smem_bytes;
};

__global__ void __launch_bounds__(256) k_gdiag(const GdiagParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TreeView& t = prm.t;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int n = t.n;
    const unsigned FULL = 0xffffffffu;
    double* sb = (double*)(smem_raw + (size_t)wib * prm.warp_smem_bytes);
    double *q1 = sb + n, *q2 = q1 + n, *r1 = q2 + n, *r2 = r1 + n;
    double* cx = r2 + n;      // [64] candidate xcoef
    double* cf = cx + 64;     // [64] candidate fd
    int* cd = (int*)(cf + 64);  // [64] candidate dimension (1-based) or 0
    for (long long bi = (long long)blockIdx.x * W + wib; bi < prm.nb; bi += (long long)gridDim.x * W) {
        const int box = prm.node_ids ? prm.node_ids[bi] : t.lnode[bi];
        double* Qo = prm.Q + bi * (long long)n * n;
        // b = position(box), c = value(box)
        const int4 binfo = t.ninfo[box];
        const bool isroot = binfo.x == box, internal = !(binfo.w & CSP_NF_LEAF);
        const int Ip = t.ninfo[binfo.x].y, ci = CSP_NF_CHILDINDEX(binfo.w);
        for (int i = lane; i < n; i += 32) {
            double bb;
            if (internal) bb = t.ipos[(size_t)binfo.y * n + i];
            else if (isroot) bb = dnan();
            else {
                bb = t.ipos[(size_t)Ip * n + i];
                if ((ci & 1) && t.idims[2 * Ip] - 1 == i) bb = t.ixs[2 * Ip];
                if ((ci & 2) && t.idims[2 * Ip + 1] - 1 == i) bb = t.ixs[2 * Ip + 1];
            }
            sb[i] = bb;
            q1[i] = q2[i] = r1[i] = r2[i] = dnan();
        }
        __syncwarp();
        int tpos = 0;
        for_split_groups(t, box, [&](const RangeGroup& grp) -> bool {
            const int total = grp.total();
            for (int base = 0; base < total; base += 32) {
                const int v = base + lane;
                const int pos = tpos + lane;
                const bool act = v < total && pos >= prm.skip;
                cd[2 * lane] = cd[2 * lane + 1] = 0;
                if (act) {
                    const int id = grp.id(v);
                    const double* __restrict__ x = t.ipos + (size_t)id * n;
                    const double f0 = t.icval[4 * id];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int d = t.idims[2 * id + i];
                        if (d > n || !(q2[d - 1] != q2[d - 1])) continue;  // fictive, or both equations already known
                        const double xs = t.ixs[2 * id + i];
                        const double xd = x[d - 1];
                        const double xcoef = (xd + xs) / 2 - sb[d - 1];
                        const double dx = xs - xd;
                        double fd = (t.icval[4 * id + 1 + i] - f0) / dx;
                        for (int j = 0; j < n; ++j) {
                            if (j == d - 1) continue;
                            const int lo = min(d - 1, j), hi = max(d - 1, j);
                            fd = fd - Qo[(size_t)hi * n + lo] * (x[j] - sb[j]);
                        }
                        cd[2 * lane + i] = d; cx[2 * lane + i] = xcoef; cf[2 * lane + i] = fd;
                    }
                }
                __syncwarp();
                for (int dd = lane; dd < n; dd += 32) {
                    if (!(q2[dd] != q2[dd])) continue;
                    for (int sidx = 0; sidx < 64; ++sidx) {
                        if (cd[sidx] != dd + 1) continue;
                        const double xcoef = cx[sidx], fd = cf[sidx];
                        if (xcoef == q1[dd]) continue;  // not independent of the first equation
                        if (isfinite(xcoef) && isfinite(fd)) {
                            if (q1[dd] != q1[dd]) { q1[dd] = xcoef; r1[dd] = fd; }
                            else { q2[dd] = xcoef; r2[dd] = fd; break; }
                        }
                    }
                }
                __syncwarp();
                bool pending = false;
                for (int dd = lane; dd < n; dd += 32) pending |= (q2[dd] != q2[dd]);
                tpos += min(32, total - base);
                if (!__any_sync(FULL, pending)) return true;
            }
            return false;
        });
        __syncwarp();
        for (int d = lane; d < n; d += 32) {
            double gd = dnan();
            if (!(q2[d] != q2[d])) {  // manual inverse of the 2x2 system (a11 = a21 = 1)
                const double a12 = q1[d], a22 = q2[d], rr1 = r1[d], rr2 = r2[d];
                const double det = a22 - a12;
                gd = (a22 * rr1 - a12 * rr2) / det;
                Qo[(size_t)d * n + d] = (-rr1 + rr2) / det;
            }
            prm.g[bi * n + d] = gd;
            prm.b[bi * n + d] = sb[d];
        }
        if (lane == 0) {
            double cv;
            if (internal) cv = t.icval[4 * (size_t)binfo.y];
            else if (isroot) cv = dnan();
            else cv = t.icval[4 * (size_t)Ip + ci];
            prm.c[bi] = cv;
        }
        __syncwarp();
    }
}

// ---- possemidef (reference src/posdef.jl:41-68), in place on the resident model records ------------------------------
// One warp per model, one lane per diagonal entry k: its additions happen in the reference's order (as row index i = k for
// the columns j < k, then as column index j = k for the rows i > k).  Off-diagonals are untouched.  The record stores the
// diagonal halved (common.cuh), so Q_kk = 2 * A_kk on the way in and A_kk = Q_kk / 2 on the way out (both exact).
__device__ __forceinline__ double jl_sign(double q) { return q > 0 ? 1.0 : (q < 0 ? -1.0 : q); }
__global__ void __launch_bounds__(256) k_possemidef(double* __restrict__ data, int n, int stride, long long n_models) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long m = warp; m < n_models; m += nwarps) {
        double* A = data + m * stride + n;
        auto off = [&](int i, int j) { const int lo = min(i, j), hi = max(i, j); return lo * (n + 1) - lo * (lo - 1) / 2 + (hi - lo); };
        double newd[32];  // this lane's diagonal entries k = lane, lane + 32, ... (n <= 1024 handled in passes of 32 per lane)
        for (int base = 0; base < n; base += 32 * 32) {
            int cnt = 0;
            for (int k = base + lane; k < n && cnt < 32; k += 32, ++cnt) {
                const double qkk = fabs(2.0 * A[off(k, k)]);
                double acc = 0.0;
                for (int j = 0; j < k; ++j) {  // k plays i: Qp[i,i] += q / r with r from (qjj, qii)
                    const double q = A[off(k, j)];
                    if ((q == 0) | (q != q)) continue;
                    const double qjj = fabs(2.0 * A[off(j, j)]);
                    double r = jl_sign(q) * sqrt(qjj / qkk);
                    if (r == 0 || !isfinite(r)) r = jl_sign(q);
                    acc = acc + q / r;
                }
                for (int i = k + 1; i < n; ++i) {  // k plays j: Qp[j,j] += q * r
                    const double q = A[off(i, k)];
                    if ((q == 0) | (q != q)) continue;
                    const double qii = fabs(2.0 * A[off(i, i)]);
                    double r = jl_sign(q) * sqrt(qkk / qii);
                    if (r == 0 || !isfinite(r)) r = jl_sign(q);
                    acc = acc + q * r;
                }
                newd[cnt] = (acc != acc || qkk != qkk) ? acc + qkk : (acc > qkk ? acc : qkk);  // Julia max(): NaN-propagating
            }
            __syncwarp();  // every lane has read the old diagonal before anyone overwrites it
            cnt = 0;
            for (int k = base + lane; k < n && cnt < 32; k += 32, ++cnt) A[off(k, k)] = newd[cnt] * 0.5;
            __syncwarp();
        }
    }
}

// ---- native-form model evaluation (reference src/cs2.jl:142-153 with Qcoef_value :302-312) --------------------------------
// One thread per query point, the reference's loop order (j outer, i = j..n inner), no FMA contraction (this TU).
__global__ void __launch_bounds__(256) k_modelvalue_native(int n, double c, const double* __restrict__ g, const double* __restrict__ Q,
                                                          const double* __restrict__ b, const double* __restrict__ y,
                                                          const uint8_t* __restrict__ prec, const double* __restrict__ X, long long m,
                                                          double* __restrict__ val) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const double* __restrict__ x = X + (size_t)q * n;
        double dot = 0.0;
        for (int i = 0; i < n; ++i) dot = dot + g[i] * (x[i] - b[i]);
        double v = c + dot;
        for (int j = 0; j < n; ++j)
            for (int i = j; i < n; ++i) {
                double coef;
                if (i == j) {
                    coef = (x[i] - b[i]) * (x[i] - y[i]) / 2;
                } else {
                    const double bi = b[i], bj = b[j], yi = y[i], yj = y[j], xi = x[i], xj = x[j];
                    const bool pij = prec[(size_t)j * n + i] != 0, pji = prec[(size_t)i * n + j] != 0;  // column-major prec[i,j]
                    if (((pij & !pji) & (xi == yi)) | ((pji & !pij) & (xj == yj))) {
                        coef = 0.0 * (xi * xj / 2);  // zero(xi*xj/2)
                    } else {
                        const double chiij = pji ? bi : 2 * yi - bi;
                        const double chiji = pij ? bj : 2 * yj - bj;
                        coef = ((xi - bi) * (xj - chiji) + (xj - bj) * (xi - chiij)) / 4;
                    }
                }
                if (coef != 0) v = v + Q[(size_t)i * n + j] * coef * (double)(1 + (i != j));  // Q[i,j] reads .data[min,max] = data[j,i]
            }
        val[q] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// Highest-order coefficient of every split (src/polynomials.jl:212-224), once per tree version; on the legacy default stream,
// before the builder's final synchronisation.
int csp_fit_build_coef(csp_tree* t) {
    if (t->v.n_internal <= 0 || t->v.p > 2) return CSP_OK;
    void* pc = nullptr;
    CSP_CHECK(csp_dev_alloc(&pc, (size_t)t->v.n_internal * sizeof(double)));
    const int blocks = std::min(4096, (t->v.n_internal + 255) / 256);
    if (t->v.p == 2) k_coef_table<2><<<blocks, 256>>>(t->v, (double*)pc);
    else k_coef_table<1><<<blocks, 256>>>(t->v, (double*)pc);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    t->fit_coef = (double*)pc;
    return CSP_OK;
}

static int launch_fit(const csp_tree* t, FitParams& prm, cudaStream_t st, void** scratch_out) {
    *scratch_out = nullptr;
    if (prm.nb == 0) return CSP_OK;
    if (prm.node_ids || prm.row_hi <= 0) { prm.row_lo = 0; prm.row_hi = prm.nb; }
    if (prm.row_lo >= prm.row_hi) return CSP_OK;
    const int n = t->v.n, P = t->v.p;
    const long long npairs = prm.slotmap ? prm.npat : (P == 2 ? (long long)n * (n - 1) / 2 : n);
    const long long qlen = prm.slotmap ? prm.npat : (P == 1 ? n : (prm.q_packed ? (long long)(n + 1) * (n + 2) / 2 : (long long)n * n));
    prm.q_in_smem = qlen <= 1024 ? 1 : 0;  // up to 8 KB per warp
    size_t base = (size_t)n * 8 * 3 + (prm.q_in_smem ? (size_t)qlen * 8 : 0) + FIT_WINDOW_INTS * 4 + (size_t)n * 4 * 3;
    size_t withpairs = base + (size_t)npairs * 8;
    prm.pairs_in_smem = withpairs <= 24 * 1024;  // larger tables live in (L2-resident) global scratch, see WarpCtx::plist
    size_t per_warp = prm.pairs_in_smem ? withpairs : base;
    per_warp = (per_warp + 15) & ~size_t(15);
    prm.tiles = (!prm.q_in_smem && prm.mode == FIT_MODE_FULL) ? 1 : 0;
    if (prm.tiles) per_warp += FIT_TILE_BYTES;  // multiple of 16
    if (prm.mode == FIT_MODE_FULL && P == 2 && !prm.node_ids && prm.nb == t->v.n_leaves && t->fit_units &&
        !(getenv("CSP_FIT_SHARED") && atoi(getenv("CSP_FIT_SHARED")) == 0)) {
        prm.units = t->fit_units;  // all leaves: sibling leaves share the off-diagonal work
        prm.n_units = t->n_fit_units;
    }
    // Small n: at n = 10 only 10 of a warp's 32 lanes work in the per-leaf phases (diagonal scan, diagonal solve, canonicalize:
    // 82 % of a unit's time on C2), and a unit has about three leaves: run 32 / n of them side by side (CSP_FIT_MULTI=0: one by one).
    prm.multi = 1; prm.multi_off = 0;
    if (P == 2 && prm.units && prm.q_in_smem && prm.pairs_in_smem && !prm.tiles && !prm.prec && !prm.slotmap && n <= 16 &&
        !(getenv("CSP_FIT_MULTI") && atoi(getenv("CSP_FIT_MULTI")) == 0)) {
        const int G = std::min(4, 32 / n);
        if (G > 1) {
            prm.multi = G;
            prm.multi_off = (int)per_warp;
            per_warp += (size_t)G * (((size_t)3 * n + (size_t)qlen) * 8 + (size_t)3 * n * 4);
            per_warp = (per_warp + 15) & ~size_t(15);
        }
    }
    int W = (int)std::min<size_t>(8, std::max<size_t>(1, ((prm.tiles ? 110 : 64) * 1024) / per_warp));
    if (per_warp * W > 220 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the fit kernel", n);
    prm.warp_smem_bytes = (int)per_warp;
    size_t smem = per_warp * W;
    // the per-split coefficient table is built with the tree (csp_fit_build_coef, called by the device builder): no lazy
    // initialisation here, so concurrent fits on different streams only ever read it
    if (t->fit_coef && !(getenv("CSP_FIT_COEF_TABLE") && atoi(getenv("CSP_FIT_COEF_TABLE")) == 0)) prm.coef = t->fit_coef;
    if (prm.mode == FIT_MODE_FULL) {  // chain pre-pass: one thread per box
        const size_t nb = (size_t)prm.nb;
        const size_t need = nb * ((size_t)n * (prm.firstmatch ? 24 : 20) + 16) + 64 + 64;  // chain rows + alignment slack + the work counter
        if (need > t->fit_ws_bytes) {
            if (t->fit_ws) { CSP_CUDA(cudaDeviceSynchronize()); csp_dev_free(t->fit_ws); t->fit_ws = nullptr; t->fit_ws_bytes = 0; }
            CSP_CHECK(csp_dev_alloc(&t->fit_ws, need));
            t->fit_ws_bytes = need;
        }
        char* base_p = (char*)t->fit_ws;
        prm.ch_y = (double*)base_p; base_p += nb * n * 8;
        prm.ch_g = (double*)base_p; base_p += nb * n * 8;
        prm.ch_c = (double*)base_p; base_p += nb * 8;
        prm.ch_step = (int*)base_p; base_p += nb * n * 4;
        prm.ch_status = (int*)base_p; base_p += nb * 4;
        prm.ch_itop = (int*)base_p; base_p += nb * 4;
        if (prm.firstmatch) { prm.ch_fm = (int*)base_p; base_p += nb * n * 4; }
        prm.work_counter = (unsigned long long*)(((uintptr_t)base_p + 63) & ~(uintptr_t)63);
        if (getenv("CSP_FIT_DYNAMIC") && atoi(getenv("CSP_FIT_DYNAMIC")) == 0) prm.work_counter = nullptr;
        else CSP_CUDA(cudaMemsetAsync(prm.work_counter, 0, sizeof(unsigned long long), st));
        const long long blocks = std::min<long long>((prm.row_hi - prm.row_lo + 127) / 128, (long long)std::max(t->n_sm, 1) * 16);
        {
            CspTimed tm(CSP_K_FIT, st);
            k_fit_chain<<<(unsigned)blocks, 128, 0, st>>>(prm.t, prm.node_ids, prm.row_lo, prm.row_hi, prm.ch_status, prm.ch_itop, prm.ch_c, prm.ch_y,
                                                         prm.ch_g, prm.ch_step, prm.ch_fm);
        }
        CSP_LAUNCHED();
        CSP_CUDA(cudaGetLastError());
    }
    auto launch = [&](auto kern) -> int {
        CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        CSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem));
        if (occ < 1) return csp_set_error(CSP_ERR_CUDA, "fit kernel does not fit on an SM (smem %zu)", smem);
        // Sharing divides the number of work items by ~3: with large records (few warps resident) that only pays while there
        // are still many items per warp (C5, 3 450 splits with leaves on 1 480 warps: 16 ms per leaf vs 19 ms shared;
        // C5det, 44 850: 1 593 vs 611 ms).
        const bool force_shared = getenv("CSP_FIT_SHARED") && atoi(getenv("CSP_FIT_SHARED")) == 2;  // tests
        if (prm.units && !prm.q_in_smem && !force_shared && prm.n_units < 8LL * std::max(t->n_sm, 1) * occ * W) prm.units = nullptr;
        const long long nunits = prm.units ? prm.n_units : prm.row_hi - prm.row_lo;
        long long grid = std::min<long long>((nunits + W - 1) / W, (long long)std::max(t->n_sm, 1) * occ);
        if (!prm.pairs_in_smem) {
            CSP_CHECK(csp_dev_alloc(scratch_out, (size_t)grid * W * 3 * ((npairs + 1) & ~1LL) * sizeof(int)));
            prm.scratch = (int*)*scratch_out;
        }
        const bool prof = getenv("CSP_FIT_PROF") && atoi(getenv("CSP_FIT_PROF"));
        if (prof) {
            CSP_CUDA(cudaMalloc(&prm.prof, 8 * sizeof(unsigned long long)));
            CSP_CUDA(cudaMemset(prm.prof, 0, 8 * sizeof(unsigned long long)));
        }
        {
            CspTimed tm(CSP_K_FIT, st);
            kern<<<(unsigned)grid, W * 32, smem, st>>>(prm);
        }
        CSP_LAUNCHED();
        CSP_CUDA(cudaGetLastError());
        if (prof) {  // developer aid: warp-clock totals per phase
            unsigned long long h[8];
            CSP_CUDA(cudaMemcpy(h, prm.prof, sizeof(h), cudaMemcpyDeviceToHost));
            cudaFree(prm.prof);
            static const char* names[5] = {"init+chain", "scan", "coefficients", "diagonal", "canon+output"};
            for (int k = 0; k < 5; ++k) fprintf(stderr, "[csp fit prof] %-13s %10.0f clk/box\n", names[k], (double)h[k] / (double)prm.nb);
        }
        return CSP_OK;
    };
    const bool big = !prm.pairs_in_smem || prm.tiles;
    if (P == 2 && prm.multi > 1 && !big) return launch(k_fit<2, false, true>);
    if (P == 2) return big ? launch(k_fit<2, true>) : launch(k_fit<2, false>);
    return big ? launch(k_fit<1, true>) : launch(k_fit<1, false>);
}

// dense NaN-filled status helper for node id validation
static int check_nodes(const csp_tree* t, const int32_t* node_ids, int64_t nb) {
    if (!t) return csp_set_error(CSP_ERR_INVALID, "tree is NULL");
    if (nb < 0) return csp_set_error(CSP_ERR_INVALID, "nb < 0");
    if (!node_ids && nb != t->v.n_leaves)
        return csp_set_error(CSP_ERR_INVALID, "node_ids == NULL means all leaves: nb must be %d (got %lld)", t->v.n_leaves, (long long)nb);
    if (node_ids)
        for (int64_t i = 0; i < nb; ++i)
            if (node_ids[i] < 0 || node_ids[i] >= t->v.n_nodes)
                return csp_set_error(CSP_ERR_INVALID, "node id %d out of range at %lld", node_ids[i], (long long)i);
    return CSP_OK;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { csp_dev_free(p); }  // every user synchronises before its buffers go out of scope
    int alloc(size_t bytes) { return csp_dev_alloc(&p, bytes); }
};

extern "C" {

int csp_fit_boxes(const csp_tree* t, const int32_t* node_ids, int64_t nb, int32_t skip, double* c, double* g, double* Q, double* b,
                  double* g_native, double* y, uint8_t* prec, int32_t* firstmatch, uint8_t* status) {
    CSP_CHECK(check_nodes(t, node_ids, nb));
    if (t->v.p != 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "fit_quadratic is defined for Box{2} only (reference src/cs2.jl:89); tree has p=%d", t->v.p);
    if (skip < 0) return csp_set_error(CSP_ERR_INVALID, "skip < 0");
    if (nb == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    DeviceGuard dg(t->device);
    const int n = t->v.n;
    // chunk the boxes so that the dense Q workspace stays below ~2 GiB
    const long long per = (long long)n * n * 8;
    const long long chunk = std::max<long long>(1, std::min<long long>(nb, (2LL << 30) / per));
    DevBuf dids, dc, dg_, dQ, db, dgn, dy, dprec, dst, dfm;
    if (node_ids) CSP_CHECK(dids.alloc(chunk * 4));
    if (firstmatch) CSP_CHECK(dfm.alloc(chunk * n * 4));
    CSP_CHECK(dc.alloc(chunk * 8)); CSP_CHECK(dg_.alloc(chunk * n * 8)); CSP_CHECK(dQ.alloc(chunk * per)); CSP_CHECK(db.alloc(chunk * n * 8));
    if (g_native) CSP_CHECK(dgn.alloc(chunk * n * 8));
    if (y) CSP_CHECK(dy.alloc(chunk * n * 8));
    if (prec) CSP_CHECK(dprec.alloc(chunk * (long long)n * n));
    CSP_CHECK(dst.alloc(chunk));
    for (long long off = 0; off < nb; off += chunk) {
        const long long cnt = std::min(chunk, nb - off);
        FitParams prm{};
        prm.t = t->v;
        prm.nb = cnt; prm.skip = skip; prm.mode = FIT_MODE_FULL;
        if (node_ids) {
            CSP_CUDA(cudaMemcpy(dids.p, node_ids + off, cnt * 4, cudaMemcpyHostToDevice));
            prm.node_ids = (const int*)dids.p;
        } else {
            prm.node_ids = t->v.lnode + off;
        }
        prm.c = (double*)dc.p; prm.c_stride = 1;
        prm.g = (double*)dg_.p; prm.g_stride = n;
        prm.Q = (double*)dQ.p; prm.Q_stride = (long long)n * n; prm.q_packed = 0;
        prm.b = (double*)db.p; prm.b_stride = n;
        prm.gnat = (double*)dgn.p; prm.y = (double*)dy.p; prm.prec = (uint8_t*)dprec.p; prm.status = (uint8_t*)dst.p;
        prm.firstmatch = (int*)dfm.p;
        void* scratch = nullptr;
        int rc = launch_fit(t, prm, 0, &scratch);
        if (rc == CSP_OK && cudaDeviceSynchronize() != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "fit kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        csp_dev_free(scratch);
        CSP_CHECK(rc);
        if (c) CSP_CUDA(cudaMemcpy(c + off, dc.p, cnt * 8, cudaMemcpyDeviceToHost));
        if (g) CSP_CUDA(cudaMemcpy(g + off * n, dg_.p, cnt * n * 8, cudaMemcpyDeviceToHost));
        if (Q) CSP_CUDA(cudaMemcpy(Q + off * (long long)n * n, dQ.p, cnt * per, cudaMemcpyDeviceToHost));
        if (b) CSP_CUDA(cudaMemcpy(b + off * n, db.p, cnt * n * 8, cudaMemcpyDeviceToHost));
        if (g_native) CSP_CUDA(cudaMemcpy(g_native + off * n, dgn.p, cnt * n * 8, cudaMemcpyDeviceToHost));
        if (y) CSP_CUDA(cudaMemcpy(y + off * n, dy.p, cnt * n * 8, cudaMemcpyDeviceToHost));
        if (prec) CSP_CUDA(cudaMemcpy(prec + off * (long long)n * n, dprec.p, cnt * (long long)n * n, cudaMemcpyDeviceToHost));
        if (status) CSP_CUDA(cudaMemcpy(status + off, dst.p, cnt, cudaMemcpyDeviceToHost));
        if (firstmatch) CSP_CUDA(cudaMemcpy(firstmatch + off * n, dfm.p, cnt * n * 4, cudaMemcpyDeviceToHost));
    }
    return CSP_OK;
}

int csp_coefficients_generic(const csp_tree* t, const int32_t* node_ids, int64_t nb, int32_t skip, double* Cp, int64_t* visited, int64_t* used);  // extras.cu

int csp_coefficients_p(const csp_tree* t, const int32_t* node_ids, int64_t nb, int32_t skip, double* Cp, int64_t* visited, int64_t* used) {
    CSP_CHECK(check_nodes(t, node_ids, nb));
    if (skip < 0 || !Cp) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    if (nb == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    if (t->v.p > 2 || (getenv("CSP_COEF_GENERIC") && atoi(getenv("CSP_COEF_GENERIC")))) return csp_coefficients_generic(t, node_ids, nb, skip, Cp, visited, used);
    DeviceGuard dg(t->device);
    const int n = t->v.n;
    const long long per = (t->v.p == 2 ? (long long)n * n : n) * 8;
    const long long chunk = std::max<long long>(1, std::min<long long>(nb, (2LL << 30) / per));
    DevBuf dids, dQ, dvis, dused;
    if (node_ids) CSP_CHECK(dids.alloc(chunk * 4));
    CSP_CHECK(dQ.alloc(chunk * per));
    CSP_CHECK(dvis.alloc(chunk * 8));
    CSP_CHECK(dused.alloc(chunk * 8));
    for (long long off = 0; off < nb; off += chunk) {
        const long long cnt = std::min(chunk, nb - off);
        FitParams prm{};
        prm.t = t->v;
        prm.nb = cnt; prm.skip = skip; prm.mode = FIT_MODE_COEF;
        if (node_ids) {
            CSP_CUDA(cudaMemcpy(dids.p, node_ids + off, cnt * 4, cudaMemcpyHostToDevice));
            prm.node_ids = (const int*)dids.p;
        } else {
            prm.node_ids = t->v.lnode + off;
        }
        prm.Q = (double*)dQ.p; prm.Q_stride = per / 8; prm.q_packed = 0;
        prm.visited = (long long*)dvis.p; prm.used = (long long*)dused.p;
        void* scratch = nullptr;
        int rc = launch_fit(t, prm, 0, &scratch);
        if (rc == CSP_OK && cudaDeviceSynchronize() != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "coefficients kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        csp_dev_free(scratch);
        CSP_CHECK(rc);
        CSP_CUDA(cudaMemcpy(Cp + off * (per / 8), dQ.p, cnt * per, cudaMemcpyDeviceToHost));
        if (visited) CSP_CUDA(cudaMemcpy(visited + off, dvis.p, cnt * 8, cudaMemcpyDeviceToHost));
        if (used) CSP_CUDA(cudaMemcpy(used + off, dused.p, cnt * 8, cudaMemcpyDeviceToHost));
    }
    return CSP_OK;
}

int csp_coefficients_p_sparse(const csp_tree* t, const int32_t* node_ids, int64_t nb, int32_t skip, int32_t npat, const int32_t* pat_i,
                              const int32_t* pat_j, double* vals, int64_t* visited, int64_t* used) {
    CSP_CHECK(check_nodes(t, node_ids, nb));
    if (t->v.p != 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "sparse coefficients_p output is defined for p == 2 (symmetric matrices)");
    if (skip < 0 || npat < 0 || (npat > 0 && (!pat_i || !pat_j || !vals))) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    const int n = t->v.n;
    std::vector<int> slot((size_t)n * n, -1);
    for (int k = 0; k < npat; ++k) {
        if (!(1 <= pat_i[k] && pat_i[k] < pat_j[k] && pat_j[k] <= n))
            return csp_set_error(CSP_ERR_INVALID, "pattern entry %d = (%d, %d) must satisfy 1 <= i < j <= n", k, pat_i[k], pat_j[k]);
        slot[(size_t)(pat_i[k] - 1) * n + (pat_j[k] - 1)] = k;
    }
    if (nb == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    DeviceGuard dg(t->device);
    const long long per = std::max(npat, 1);
    const long long chunk = std::max<long long>(1, std::min<long long>(nb, (1LL << 30) / (per * 8)));
    DevBuf dids, dQ, dvis, dslot, dused;
    if (node_ids) CSP_CHECK(dids.alloc(chunk * 4));
    CSP_CHECK(dQ.alloc(chunk * per * 8));
    CSP_CHECK(dvis.alloc(chunk * 8));
    CSP_CHECK(dused.alloc(chunk * 8));
    CSP_CHECK(dslot.alloc((size_t)n * n * 4));
    CSP_CUDA(cudaMemcpy(dslot.p, slot.data(), (size_t)n * n * 4, cudaMemcpyHostToDevice));
    for (long long off = 0; off < nb; off += chunk) {
        const long long cnt = std::min(chunk, nb - off);
        FitParams prm{};
        prm.t = t->v;
        prm.nb = cnt; prm.skip = skip; prm.mode = FIT_MODE_COEF;
        if (node_ids) {
            CSP_CUDA(cudaMemcpy(dids.p, node_ids + off, cnt * 4, cudaMemcpyHostToDevice));
            prm.node_ids = (const int*)dids.p;
        } else {
            prm.node_ids = t->v.lnode + off;
        }
        prm.Q = (double*)dQ.p; prm.Q_stride = per; prm.q_packed = 0;
        prm.slotmap = (const int*)dslot.p; prm.npat = npat;
        prm.visited = (long long*)dvis.p; prm.used = (long long*)dused.p;
        void* scratch = nullptr;
        int rc = launch_fit(t, prm, 0, &scratch);
        if (rc == CSP_OK && cudaDeviceSynchronize() != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "coefficients kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        csp_dev_free(scratch);
        CSP_CHECK(rc);
        if (npat > 0) CSP_CUDA(cudaMemcpy(vals + off * npat, dQ.p, cnt * (long long)npat * 8, cudaMemcpyDeviceToHost));
        if (visited) CSP_CUDA(cudaMemcpy(visited + off, dvis.p, cnt * 8, cudaMemcpyDeviceToHost));
        if (used) CSP_CUDA(cudaMemcpy(used + off, dused.p, cnt * 8, cudaMemcpyDeviceToHost));
    }
    return CSP_OK;
}

static int models_fit_launch(const csp_tree* t, int32_t skip, csp_models* mo, cudaStream_t st, void** scratch, long long row_lo = 0,
                             long long row_hi = -1) {
    const int n = t->v.n;
    FitParams prm{};
    prm.t = t->v;
    prm.node_ids = nullptr; prm.nb = mo->n_models; prm.skip = skip;
    prm.row_lo = row_lo; prm.row_hi = row_hi < 0 ? mo->n_models : row_hi;
    prm.mode = mo->degree == 2 ? FIT_MODE_FULL : FIT_MODE_LIN;
    prm.b = mo->data; prm.b_stride = mo->stride;
    prm.status = mo->status;
    prm.Q = mo->data + n; prm.Q_stride = mo->stride;
    if (mo->degree == 2) {
        prm.q_packed = 2;  // g, c and Q all live inside the packed A (common.cuh)
    } else {
        prm.q_packed = 0;  // Cp (the gradient) lands in the g slot
        prm.c = mo->data + 2 * n; prm.c_stride = mo->stride;
    }
    return launch_fit(t, prm, st, scratch);
}

int csp_models_fit(const csp_tree* t, int32_t skip, csp_models** out) {
    if (!t || !out) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (skip < 0) return csp_set_error(CSP_ERR_INVALID, "skip < 0");
    if (t->v.p > 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "models exist for degrees p <= 2 only (the tree has p=%d)", t->v.p);
    CSP_CHECK(csp_require_device());
    DeviceGuard dg(t->device);
    const int n = t->v.n;
    csp_models* mo = new csp_models();
    mo->device = t->device; mo->n = n; mo->degree = t->v.p == 2 ? 2 : 1;
    mo->stride = csp_model_stride(n, mo->degree);
    mo->n_models = t->v.n_leaves;
    auto fail = [&](int rc) { csp_models_free(mo); return rc; };
    if (csp_dev_alloc((void**)&mo->data, (size_t)mo->n_models * mo->stride * 8) != CSP_OK || csp_dev_alloc((void**)&mo->status, (size_t)mo->n_models) != CSP_OK)
        return fail(csp_set_error(CSP_ERR_NOMEM, "cannot allocate %lld model records of %d doubles", mo->n_models, mo->stride));
    if (cudaMemset(mo->data, 0, (size_t)mo->n_models * mo->stride * 8) != cudaSuccess) return fail(csp_set_error(CSP_ERR_CUDA, "memset failed"));
    void* scratch = nullptr;
    int rc = models_fit_launch(t, skip, mo, 0, &scratch);
    if (rc == CSP_OK && cudaDeviceSynchronize() != cudaSuccess) rc = csp_set_error(CSP_ERR_CUDA, "fit kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    csp_dev_free(scratch);
    if (rc != CSP_OK) return fail(rc);
    *out = mo;
    return CSP_OK;
}

int csp_fit_gdiag(const csp_tree* t, const int32_t* node_ids, int64_t nb, int32_t skip, double* c, double* g, double* Q, double* b) {
    CSP_CHECK(check_nodes(t, node_ids, nb));
    if (t->v.p != 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "fit_quadratic_gdiag! is defined for Box{2} only (reference src/cs2.jl:332)");
    if (skip < 0 || !c || !g || !Q || !b) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    if (nb == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    DeviceGuard dg(t->device);
    const int n = t->v.n;
    const long long per = (long long)n * n * 8;
    const long long chunk = std::max<long long>(1, std::min<long long>(nb, (2LL << 30) / per));
    DevBuf dids, dc, dgv, dQ, db;
    if (node_ids) CSP_CHECK(dids.alloc(chunk * 4));
    CSP_CHECK(dc.alloc(chunk * 8)); CSP_CHECK(dgv.alloc(chunk * n * 8)); CSP_CHECK(dQ.alloc(chunk * per)); CSP_CHECK(db.alloc(chunk * n * 8));
    const size_t per_warp = (((size_t)n * 5 + 128) * 8 + 64 * 4 + 15) & ~size_t(15);
    const int W = (int)std::min<size_t>(8, std::max<size_t>(1, (64 * 1024) / per_warp));
    if (per_warp * W > 200 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the gdiag kernel", n);
    CSP_CUDA(cudaFuncSetAttribute(k_gdiag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * W)));
    for (long long off = 0; off < nb; off += chunk) {
        const long long cnt = std::min(chunk, nb - off);
        const int* ids = t->v.lnode + off;
        if (node_ids) {
            CSP_CUDA(cudaMemcpy(dids.p, node_ids + off, cnt * 4, cudaMemcpyHostToDevice));
            ids = (const int*)dids.p;
        }
        // Q = coefficients_p(box; skip)
        FitParams fp{};
        fp.t = t->v; fp.node_ids = ids; fp.nb = cnt; fp.skip = skip; fp.mode = FIT_MODE_COEF;
        fp.Q = (double*)dQ.p; fp.Q_stride = (long long)n * n; fp.q_packed = 0;
        void* scratch = nullptr;
        int rc = launch_fit(t, fp, 0, &scratch);
        // fit_quadratic_gdiag!(Q, box; skip)
        if (rc == CSP_OK) {
            GdiagParams gp{};
            gp.t = t->v; gp.node_ids = ids; gp.nb = cnt; gp.skip = skip;
            gp.c = (double*)dc.p; gp.g = (double*)dgv.p; gp.Q = (double*)dQ.p; gp.b = (double*)db.p;
            gp.warp_smem_bytes = (int)per_warp;
            const long long grid = std::min<long long>((cnt + W - 1) / W, (long long)std::max(t->n_sm, 1) * 4);
            k_gdiag<<<(unsigned)grid, W * 32, per_warp * W>>>(gp);
            CSP_LAUNCHED();
            if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
                rc = csp_set_error(CSP_ERR_CUDA, "gdiag kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        csp_dev_free(scratch);
        CSP_CHECK(rc);
        CSP_CUDA(cudaMemcpy(c + off, dc.p, cnt * 8, cudaMemcpyDeviceToHost));
        CSP_CUDA(cudaMemcpy(g + off * n, dgv.p, cnt * n * 8, cudaMemcpyDeviceToHost));
        CSP_CUDA(cudaMemcpy(Q + off * (long long)n * n, dQ.p, cnt * per, cudaMemcpyDeviceToHost));
        CSP_CUDA(cudaMemcpy(b + off * n, db.p, cnt * n * 8, cudaMemcpyDeviceToHost));
    }
    return CSP_OK;
}

int csp_models_possemidef(csp_models* mo, void* stream) {
    csp_models_wait_ready(mo, (cudaStream_t)stream);
    if (!mo) return csp_set_error(CSP_ERR_INVALID, "models is NULL");
    if (mo->degree != 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "possemidef needs quadratic (degree-2) models");
    if (mo->n > 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "possemidef kernel supports n <= 1024");
    DeviceGuard dg(mo->device);
    const long long blocks = std::min<long long>((mo->n_models + 7) / 8, 148 * 8);
    k_possemidef<<<(unsigned)std::max<long long>(blocks, 1), 256, 0, (cudaStream_t)stream>>>(mo->data, mo->n, mo->stride, mo->n_models);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}

int csp_modelvalue_native(int32_t n, double c, const double* g, const double* Q, const double* b, const double* y, const uint8_t* prec,
                          const double* X, int64_t m, double* val) {
    if (n < 1 || m < 0 || !g || !Q || !b || !y || !prec || (m > 0 && (!X || !val))) return csp_set_error(CSP_ERR_INVALID, "bad argument");
    if (m == 0) return CSP_OK;
    CSP_CHECK(csp_require_device());
    DeviceGuard dg(csp_current_device());
    DevBuf dgv, dQ, db, dy, dp, dX, dv;
    CSP_CHECK(dgv.alloc((size_t)n * 8)); CSP_CHECK(dQ.alloc((size_t)n * n * 8)); CSP_CHECK(db.alloc((size_t)n * 8));
    CSP_CHECK(dy.alloc((size_t)n * 8)); CSP_CHECK(dp.alloc((size_t)n * n)); CSP_CHECK(dX.alloc((size_t)m * n * 8)); CSP_CHECK(dv.alloc((size_t)m * 8));
    CSP_CUDA(cudaMemcpy(dgv.p, g, (size_t)n * 8, cudaMemcpyHostToDevice));
    CSP_CUDA(cudaMemcpy(dQ.p, Q, (size_t)n * n * 8, cudaMemcpyHostToDevice));
    CSP_CUDA(cudaMemcpy(db.p, b, (size_t)n * 8, cudaMemcpyHostToDevice));
    CSP_CUDA(cudaMemcpy(dy.p, y, (size_t)n * 8, cudaMemcpyHostToDevice));
    CSP_CUDA(cudaMemcpy(dp.p, prec, (size_t)n * n, cudaMemcpyHostToDevice));
    CSP_CUDA(cudaMemcpy(dX.p, X, (size_t)m * n * 8, cudaMemcpyHostToDevice));
    const long long blocks = std::min<long long>((m + 255) / 256, 148 * 8);
    k_modelvalue_native<<<(unsigned)blocks, 256>>>(n, c, (const double*)dgv.p, (const double*)dQ.p, (const double*)db.p, (const double*)dy.p,
                                                   (const uint8_t*)dp.p, (const double*)dX.p, m, (double*)dv.p);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    CSP_CUDA(cudaMemcpy(val, dv.p, (size_t)m * 8, cudaMemcpyDeviceToHost));
    return CSP_OK;
}

int csp_models_alloc(const csp_tree* t, csp_models** out) { return csp_models_alloc_ex(t, false, out); }

// plain: plain cudaMalloc memory (what cudaIpcGetMemHandle / peer copies need) instead of the stream-ordered pool
int csp_models_alloc_ex(const csp_tree* t, bool plain, csp_models** out) {
    if (!t || !out) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (t->v.p > 2) return csp_set_error(CSP_ERR_UNSUPPORTED, "models exist for degrees p <= 2 only (the tree has p=%d)", t->v.p);
    CSP_CHECK(csp_require_device());
    DeviceGuard dg(t->device);
    csp_models* mo = new csp_models();
    mo->device = t->device; mo->n = t->v.n; mo->degree = t->v.p == 2 ? 2 : 1;
    mo->stride = csp_model_stride(mo->n, mo->degree);
    mo->n_models = t->v.n_leaves;
    mo->plain = plain;
    auto alloc = [&](void** q, size_t bytes) { return plain ? (cudaMalloc(q, std::max<size_t>(bytes, 16)) == cudaSuccess ? CSP_OK : (cudaGetLastError(), CSP_ERR_NOMEM)) : csp_dev_alloc(q, bytes); };
    if (alloc((void**)&mo->data, (size_t)mo->n_models * mo->stride * 8) != CSP_OK || alloc((void**)&mo->status, (size_t)mo->n_models) != CSP_OK ||
        cudaMemset(mo->data, 0, (size_t)mo->n_models * mo->stride * 8) != cudaSuccess || cudaMemset(mo->status, 0, (size_t)mo->n_models) != cudaSuccess) {
        csp_models_free(mo);
        return csp_set_error(CSP_ERR_NOMEM, "cannot allocate %d model records", t->v.n_leaves);
    }
    *out = mo;
    return CSP_OK;
}

int csp_models_refit(const csp_tree* t, int32_t skip, csp_models* mo, void* stream) {
    return csp_models_refit_range(t, skip, mo, 0, t ? t->v.n_leaves : 0, stream);
}

int csp_models_refit_range(const csp_tree* t, int32_t skip, csp_models* mo, int64_t leaf_lo, int64_t leaf_hi, void* stream) {
    if (!t || !mo) return csp_set_error(CSP_ERR_INVALID, "NULL argument");
    if (skip < 0) return csp_set_error(CSP_ERR_INVALID, "skip < 0");
    if (mo->device != t->device || mo->n != t->v.n || mo->n_models != t->v.n_leaves || mo->degree != (t->v.p == 2 ? 2 : 1))
        return csp_set_error(CSP_ERR_INVALID, "models handle does not belong to this tree");
    if (leaf_lo < 0 || leaf_hi > t->v.n_leaves || leaf_lo > leaf_hi)
        return csp_set_error(CSP_ERR_INVALID, "leaf range [%lld, %lld) outside [0, %d)", (long long)leaf_lo, (long long)leaf_hi, t->v.n_leaves);
    if (leaf_lo == leaf_hi) return CSP_OK;
    DeviceGuard dg(t->device);
    void* scratch = nullptr;
    int rc = models_fit_launch(t, skip, mo, (cudaStream_t)stream, &scratch, leaf_lo, leaf_hi);
    if (scratch) {  // large-n path uses a global scratch table: wait before releasing it
        cudaStreamSynchronize((cudaStream_t)stream);
        csp_dev_free(scratch);
    }
    return rc;
}

}  // extern "C"
