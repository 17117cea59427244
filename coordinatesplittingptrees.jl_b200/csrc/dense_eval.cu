// dense_eval.cu -- K3b: grouped dense evaluation on the FP64 tensor cores (DMMA) for large n.
//
// Replaces modelvalue(c, g, Q, b, x) (reference src/cs2.jl:155-159) when many binned queries share a leaf's dense model:
// with z = (x - b, 1, 0...) and U the upper-triangular (n+1)x(n+1) matrix of the model record (common.cuh), the values of
// a tile of queries are  diag(Z U Z')  =  rowsum((Z U) .* Z):  one TR x NP x NP triangular GEMM per tile plus a row dot.
// Blackwell's tcgen05 has no FP64 kind; the FP64 tensor path on sm_100a is the warp-level mma.sync (SASS DMMA), here
// m8n8k4.  Queries arrive leaf-sorted from the binning pipeline (binned_eval.cu), so consecutive tiles of a CTA share a
// leaf and U is expanded into shared memory only when the leaf changes.
//
// One kernel, software-pipelined per CTA over its contiguous range of 64-row tiles:
//   * two Z buffers: while the MMAs of tile t run, the row gathers of tile t+1 are in flight (one TMA bulk copy per row,
//     or 8-byte cp.async when the rows are not 16-byte granular; both complete on the buffer's mbarrier), and the perm
//     entries of tile t+2 are already in registers;
//   * Z receives the raw query rows; z = x - b is formed in place, once per row, when the row's leaf segment starts (round 1
//     subtracted in registers at every A-fragment load: four warps per slab redid it, and every DMMA waited for a DSUB queued
//     on the same FP64 pipe); the padding columns [n] = 1, (n, LD) = 0 are written once per buffer, the gathers never touch them;
//   * U is stored packed by 8-row blocks (block kb keeps columns >= 8 kb only): half the shared memory of the square
//     layout, conflict-free fragment loads (LD = NP + 4, so every row stride is 4 mod 8 doubles);
//   * JS warps share each 16-row slab, the warp of column group jg taking the column tiles J = jg (mod JS): TR/16*JS
//     warps keep the DMMA pipe fed; the JS partial row sums are combined in a fixed order through shared memory;
//   * k-steps below a column tile's diagonal block are skipped (half the MMAs), in predicate-free unrolled stages
//     (a predicated-off DMMA still occupies the pipe).
// tools/ubench/fp64_peaks.cu measures the pipe's peak: 37.1 TFLOP/s for DMMA.8x8x4, which shares the FP64 pipe with DFMA.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

__device__ __forceinline__ double dnan_d() { return __longlong_as_double(0x7ff8000000000000LL); }


// arrive on `bar` once all of this thread's earlier cp.async have landed (the barrier's count already includes it)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct DenseParams {
    ModelsView mv;
    const double* X;
    const int2* perm;
    const int* total;
    double* val;
    int NP, LD;  // padded dimension (multiple of 8, >= n + 1) and leading dimension of the shared tiles
    int bulk;    // rows are 16-byte aligned multiples of 16 bytes: gather them with TMA bulk copies (else cp.async)
    long long* prof;  // optional [grid][8] per-phase clock counters (CSP_DMMA_PROF=1), else nullptr
};

#include "dmma_slab.cuh"

// TR rows per tile, RW rows per compute warp, JS column groups: CW = TR/RW*JS compute warps, plus PW producer warps.  The
// producers work one tile ahead of the compute warps: they issue the next tile's row gathers (a TMA bulk copy costs its issuing
// warp ~200 cycles and they serialise within a warp: with every warp issuing its share, 8 % of the tile time went there with
// the DMMA pipe idle), wait for them to land and form z = x - b in place, each row against its own leaf's b read from the
// record, so that the compute warps' loop holds nothing but fragment loads and DMMAs.  (A DADD takes ~4 cycles of the FP64
// pipe the DMMAs are queued on -- tools/ubench/fp64_peaks.cu: a 2x1-blocked DMMA loop drops from 34.1 to 27.3 TFLOP/s with one
// DADD per DMMA -- and with the subtraction in registers the JS warps of a slab each redo it: 12 % of the pipe at n = 100.)
// PW = 0: every warp computes, issues its share of the gathers, waits for the tile itself and subtracts in registers.
template <int TR, int RW, int NTJ, int JS, int PW>  // NTJ = ceil(nt / JS): every compute warp owns NTJ or NTJ - 1 column tiles
__global__ void __launch_bounds__((TR / RW * JS + PW) * 32) k_eval_dmma(const DenseParams prm) {
    constexpr int CW = TR / RW * JS;
    constexpr bool SUB = PW == 0;
    extern __shared__ __align__(16) double sm[];
    const int n = prm.mv.n, NP = prm.NP, LD = prm.LD, nt = NP / 8;
    const int usz = dmma_ubase(nt, LD);
    double* U = sm;                                  // packed upper block-triangle
    double* Zb = U + usz;                            // [2][TR][LD]
    double* sb = Zb + (size_t)2 * TR * LD;           // [NP]  (b, zero padded)
    double* s_part = sb + NP;                        // [2][JS][TR]  (alternating by tile: one barrier per tile)
    int* s_leaf = (int*)(s_part + 2 * JS * TR);      // [2][TR]
    uint64_t* bar = (uint64_t*)(s_leaf + 2 * TR);    // [2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool compute = warp < CW;
    // Row group and column group of a compute warp.  The column groups differ in cost (group jg owns the tiles J = jg mod JS, tile
    // J costs J + 1 k-blocks), and the warps of one SM sub-partition (warp mod 4) share a DMMA pipe: with 16-row warps every
    // sub-partition sees all four groups; with 32-row warps (two per sub-partition) it sees groups {0, 1} or {2, 3}, whose costs
    // add up to about the same (n = 100: 28 + 18 and 21 + 24 k-blocks).
    int rg, jg;
    if (RW == 16 || JS != 4) { rg = warp / JS; jg = (warp + rg) % JS; }
    else { rg = warp >> 2; const int s4 = warp & 3; jg = rg == 0 ? ((s4 & 1) << 1 | (s4 >> 1)) : ((((s4 & 1) << 1) | (s4 >> 1)) ^ 1); }
    const long long tot = *prm.total;
    const long long ntiles = (tot + TR - 1) / TR;
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
    if (tid == 0) {
        mbar_init(&bar[0], prm.bulk ? 1 : blockDim.x);  // one expect_tx arrival, or one cp.async arrival per thread
        mbar_init(&bar[1], prm.bulk ? 1 : blockDim.x);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * TR * (LD - n); i += blockDim.x) {  // padding columns of both Z buffers
        const int r = i / (LD - n), k = n + (i - r * (LD - n));
        Zb[(size_t)r * LD + k] = k == n ? 1.0 : 0.0;
    }
    constexpr int RPW = PW ? TR / PW : TR / CW;  // rows gathered per issuing warp
    static_assert(RPW <= 32, "one lane per gathered row");
    const bool issuer = PW ? !compute : true;
    const int iw = PW ? warp - CW : warp;
    auto perm_at = [&](long long t) -> int2 {  // this thread's perm entry of tile t (rows past the end: leaf -1)
        const long long s = t * TR + tid;
        return (tid < TR && t < t1 && s < tot) ? prm.perm[s] : make_int2(0, -1);
    };
    auto issue_at = [&](long long t) -> int2 {  // (query, leaf) of the row this lane gathers for tile t (query -1: none)
        const long long s = t * TR + iw * RPW + lane;
        return (issuer && lane < RPW && t < t1 && s < tot) ? prm.perm[s] : make_int2(-1, -1);
    };
    auto tile_rows = [&](long long t) { const long long r = tot - t * TR; return (int)(r < TR ? r : TR); };
    // Publish tile t's leaf ids, arm the barrier and start the row gathers into buffer `buf`.  No CTA barrier: the buffer and
    // s_leaf[buf] were last read before the slab barrier every thread has passed, the issuing lanes hold their rows' query
    // indices in registers (qi, prefetched two tiles ahead like ql), and the leaf ids are read only after the next slab barrier.
    auto start_tile = [&](long long t, int buf, int2 ql, int qi) {
        if (!SUB) fence_proxy_async();  // the producers' in-place z (generic-proxy writes) precede the TMA writes into this buffer
        if (tid < TR) s_leaf[buf * TR + tid] = ql.y;
        const int rows = tile_rows(t);
        if (prm.bulk) {
            if (tid == 0) mbar_expect_tx(&bar[buf], (uint32_t)((size_t)rows * n * 8));
            if (qi >= 0)
                bulk_g2s(Zb + ((size_t)buf * TR + iw * RPW + lane) * LD, prm.X + (size_t)(unsigned)qi * n, (uint32_t)(n * 8), &bar[buf]);
        } else {
            if (issuer) {
#pragma unroll 1
                for (int j = 0; j < RPW; ++j) {
                    const int q = __shfl_sync(0xffffffffu, qi, j);
                    if (q < 0) continue;
                    const double* __restrict__ src = prm.X + (size_t)(unsigned)q * n;
                    double* dst = Zb + ((size_t)buf * TR + iw * RPW + j) * LD;
                    for (int k = lane; k < n; k += 32) cp_async8(dst + k, src + k);
                }
            }
            cp_async_arrive(&bar[buf]);
        }
    };
    // producers: z = x - b in place for the rows this warp gathered into buffer `buf` (lane j holds row j's (query, leaf))
    auto fixup_rows = [&](int buf, int2 qlv) {
        double* Zt = Zb + ((size_t)buf * TR + iw * RPW) * LD;
        int prev = -2;
        double bc[4] = {0.0, 0.0, 0.0, 0.0};  // b at columns lane + 32 k  (NP <= 128 in this kernel)
#pragma unroll 4
        for (int j = 0; j < RPW; ++j) {
            const int q = __shfl_sync(0xffffffffu, qlv.x, j), lf = __shfl_sync(0xffffffffu, qlv.y, j);
            if (q < 0) continue;
            if (lf != prev) {
                const double* __restrict__ M = prm.mv.data + (size_t)lf * prm.mv.stride;
#pragma unroll
                for (int k = 0; k < 4; ++k) bc[k] = lane + 32 * k < n ? __ldg(M + lane + 32 * k) : 0.0;
                prev = lf;
            }
            double* zr = Zt + (size_t)j * LD + lane;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (lane + 32 * k < n) zr[32 * k] = zr[32 * k] - bc[k];
        }
    };
    uint32_t phase[2] = {0, 0};
    int cur_leaf = -1;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tk = clock64();
#define CSP_TICK(k) do { if (prm.prof && tid == 0) { const long long now_ = clock64(); pc[k] += now_ - tk; tk = now_; } } while (0)
    int2 nxt = make_int2(0, -1), nxt_q = make_int2(-1, -1);
    __syncthreads();
    if (t0 < t1) {
        const int2 q0 = issue_at(t0);
        start_tile(t0, 0, perm_at(t0), q0.x);
        nxt = perm_at(t0 + 1);
        nxt_q = issue_at(t0 + 1);
        if (!SUB) {  // the first tile's z, before anybody computes on it
            if (!compute) { mbar_wait(&bar[0], phase[0]); phase[0] ^= 1; fixup_rows(0, q0); }
            __syncthreads();
        }
    }
    for (long long t = t0; t < t1; ++t) {
        const int buf = (int)((t - t0) & 1);
        const int rows = tile_rows(t);
        if (t + 1 < t1) {  // the other buffer was released by the barrier that ended the previous iteration
            start_tile(t + 1, buf ^ 1, nxt, nxt_q.x);
            if (!SUB && !compute) {  // wait for the next tile's rows and form its z while the compute warps work on this tile;
                                     // the tile barrier below publishes it
                if (buf == 0) { mbar_wait(&bar[1], phase[1]); phase[1] ^= 1; }
                else { mbar_wait(&bar[0], phase[0]); phase[0] ^= 1; }
                fixup_rows(buf ^ 1, nxt_q);
            }
            nxt = perm_at(t + 2);
            nxt_q = issue_at(t + 2);
        }
        CSP_TICK(0);
        if (SUB) {
            if (buf == 0) { mbar_wait(&bar[0], phase[0]); phase[0] ^= 1; }
            else { mbar_wait(&bar[1], phase[1]); phase[1] ^= 1; }
        }
        CSP_TICK(1);
        double* Z = Zb + (size_t)buf * TR * LD;
        double* part = s_part + (size_t)buf * JS * TR;
        const int* leaf_of = s_leaf + buf * TR;
        int seg_start = 0;
        while (seg_start < rows) {
            const int lf = leaf_of[seg_start];
            int seg_end = rows;
            if (leaf_of[rows - 1] != lf) {  // sorted by leaf: first row of another leaf, by bisection (every thread, same result)
                int lo = seg_start, hi = rows - 1;  // leaf_of[lo] == lf, leaf_of[hi] != lf
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (leaf_of[mid] == lf) lo = mid; else hi = mid;
                }
                seg_end = hi;
            }
            CSP_TICK(2);
            if (lf != cur_leaf && compute) {  // expand the record: b (zero padded) and the packed U (compute warps: the
                                              // producers are busy with the next tile; barrier 1 spans the compute warps only)
                const double* __restrict__ M = prm.mv.data + (size_t)lf * prm.mv.stride;
                for (int i = tid; i < usz; i += CW * 32) U[i] = 0.0;
                for (int i = tid; i < NP; i += CW * 32) sb[i] = i < n ? __ldg(M + i) : 0.0;
                asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
                const double* __restrict__ A = M + n;
                for (int i = warp; i <= n; i += CW) {  // row i of the packed triangle: n + 1 - i coefficients, coalesced
                    const double* __restrict__ Ar = A + (i * (n + 1) - i * (i - 1) / 2);
                    const int kb = i >> 3;
                    double* Ur = U + dmma_ubase(kb, LD) + (i & 7) * (LD - 8 * kb) + (i - 8 * kb);
                    for (int e = lane; e < n + 1 - i; e += 32) Ur[e] = __ldg(Ar + e);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(CW * 32) : "memory");
                CSP_TICK(3);
            }
            cur_leaf = lf;
            const int r0 = rg * RW;
            const bool active = compute && r0 < seg_end && r0 + RW > seg_start;  // warp-uniform
            if (active) {
                if ((NTJ - 1) * JS + jg < nt) dmma_slab<NTJ, JS, RW, SUB>(Z, U, sb, part + jg * TR, r0, jg, nt, LD, lane);
                else dmma_slab<NTJ - 1, JS, RW, SUB>(Z, U, sb, part + jg * TR, r0, jg, nt, LD, lane);
            }
            CSP_TICK(5);
            __syncthreads();
            CSP_TICK(6);
            if (tid >= seg_start && tid < seg_end) {
                double v = part[tid];
#pragma unroll
                for (int g2 = 1; g2 < JS; ++g2) v += part[g2 * TR + tid];
                prm.val[t * TR + tid] = v;
            }
            seg_start = seg_end;
            // Another segment of this tile reuses `part`; the next TILE uses the other half, and nobody can reach the tile after
            // that before passing the barrier above once more, i.e. after everybody has finished this read.
            if (seg_start < rows) __syncthreads();
            CSP_TICK(4);
        }
    }
    if (prm.prof && tid == 0)
        for (int k = 0; k < 8; ++k) prm.prof[blockIdx.x * 8 + k] = pc[k];
#undef CSP_TICK
}

// ---- n > 127: U no longer fits in shared memory -------------------------------------------------------------------
// The Z tile stays resident (TR rows x LD, single buffer) and U is streamed through shared memory one 8-column panel at a
// time (double-buffered 8-byte cp.async straight from the packed record: rows k < 8 (J + 1) of columns [8J, 8J + 8); the
// panels of a leaf come from L2 for every tile after the first).  The 16 warps are TR/16 row groups x KG k-groups: a warp
// multiplies its 16 rows by its share of the panel's k range (two interleaved accumulator sets for ILP) and folds the
// result straight into its running row dot  sum_c W[r][c] z[c], which is linear in W -- so the k-groups' partial dots
// simply add up, once per tile, in a fixed order through shared memory.
constexpr int PANEL_LD = 12;  // 8 columns + 4 padding doubles: row stride 4 mod 8 (conflict-free B fragments)

template <int TR>
__global__ void __launch_bounds__(512) k_eval_dmma_panel(const DenseParams prm) {
    constexpr int RG = TR / 16, KG = 16 / RG;
    extern __shared__ __align__(16) double sm[];
    const int n = prm.mv.n, NP = prm.NP, LD = prm.LD, nt = NP / 8;
    double* Z = sm;                                   // [TR][LD]
    double* Up = Z + (size_t)TR * LD;                 // [2][NP][PANEL_LD]
    double* sb = Up + (size_t)2 * NP * PANEL_LD;      // [NP]  (b, zero padded)
    double* s_part = sb + NP;                         // [KG][TR]
    int* s_q = (int*)(s_part + KG * TR);              // [TR]
    int* s_leaf = s_q + TR;                           // [TR]
    uint64_t* bar = (uint64_t*)(s_leaf + TR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp / KG, kg = warp - rg * KG;
    const long long tot = *prm.total;
    const long long ntiles = (tot + TR - 1) / TR;
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
    if (tid == 0) {
        mbar_init(bar, prm.bulk ? 1 : blockDim.x);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < TR * (LD - n); i += blockDim.x) {  // padding columns of Z
        const int r = i / (LD - n), k = n + (i - r * (LD - n));
        Z[(size_t)r * LD + k] = k == n ? 1.0 : 0.0;
    }
    uint32_t phase = 0;
    int cur_leaf = -1;
    for (long long t = t0; t < t1; ++t) {
        const long long r_left = tot - t * TR;
        const int rows = (int)(r_left < TR ? r_left : TR);
        __syncthreads();  // previous tile fully consumed (Z, s_q, s_leaf, s_part)
        if (tid < TR) {
            const int2 ql = tid < rows ? prm.perm[t * TR + tid] : make_int2(0, -1);
            s_q[tid] = ql.x;
            s_leaf[tid] = ql.y;
        }
        if (prm.bulk) {
            fence_proxy_async();
            if (tid == 0) mbar_expect_tx(bar, (uint32_t)((size_t)rows * n * 8));
            __syncthreads();
            for (int r = tid; r < rows; r += blockDim.x)
                bulk_g2s(Z + (size_t)r * LD, prm.X + (size_t)(unsigned)s_q[r] * n, (uint32_t)(n * 8), bar);
        } else {
            __syncthreads();
            for (int r = warp; r < rows; r += 16) {
                const double* __restrict__ src = prm.X + (size_t)(unsigned)s_q[r] * n;
                double* dst = Z + (size_t)r * LD;
                for (int k = lane; k < n; k += 32) cp_async8(dst + k, src + k);
            }
            cp_async_arrive(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        int seg_start = 0;
        while (seg_start < rows) {
            const int lf = s_leaf[seg_start];
            int seg_end = rows;
            if (s_leaf[rows - 1] != lf) {  // sorted by leaf: bisection for the first row of another leaf
                int lo = seg_start, hi = rows - 1;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_leaf[mid] == lf) lo = mid; else hi = mid;
                }
                seg_end = hi;
            }
            const double* __restrict__ M = prm.mv.data + (size_t)lf * prm.mv.stride;
            const double* __restrict__ A = M + n;
            if (lf != cur_leaf) {
                for (int i = tid; i < NP; i += blockDim.x) sb[i] = i < n ? __ldg(M + i) : 0.0;
                cur_leaf = lf;
            }
            __syncthreads();
            // z = x - b in place, once per row of the segment (see k_eval_dmma); visible after the first panel barrier below
            for (int r = seg_start + warp; r < seg_end; r += 16) {
                double* zr = Z + (size_t)r * LD;
                for (int k = lane; k < n; k += 32) zr[k] = zr[k] - sb[k];
            }
            // panel J -> buffer J & 1: rows k < 8 (J + 1), columns 8J .. 8J + 7 of U (zero below the diagonal and past n)
            auto issue_panel = [&](int J) {
                double* dst = Up + (size_t)(J & 1) * NP * PANEL_LD;
                const int cnt = (8 * J + 8) * 8;
                for (int e = tid; e < cnt; e += blockDim.x) {
                    const int k = e >> 3, c = 8 * J + (e & 7);
                    if (c >= k && c <= n) cp_async8(dst + k * PANEL_LD + (e & 7), A + ((size_t)k * (n + 1) - (size_t)k * (k - 1) / 2 + (c - k)));
                    else dst[k * PANEL_LD + (e & 7)] = 0.0;
                }
            };
            const int r0 = rg * 16;
            const bool active = r0 < seg_end && r0 + 16 > seg_start;  // warp-uniform
            const double* za = Z + (size_t)(r0 + (lane >> 2)) * LD + (lane & 3);
            double sacc[2] = {0.0, 0.0};
            issue_panel(0);
            for (int J = 0; J < nt; ++J) {
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncthreads();  // panel J (and sb) visible; everybody is done with panel J - 1
                if (J + 1 < nt) issue_panel(J + 1);
                if (active) {
                    const double* ub = Up + (size_t)(J & 1) * NP * PANEL_LD + (lane & 3) * PANEL_LD + (lane >> 2);
                    const int ksteps = 2 * (J + 1);
                    const int ks0 = kg * ksteps / KG, ks1 = (kg + 1) * ksteps / KG;
                    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};  // [set][row half][2]
                    int ks = ks0;
                    for (; ks + 1 < ks1; ks += 2) {
                        const int k0 = 4 * ks;
                        const double a00 = za[k0], a01 = za[(size_t)8 * LD + k0];
                        const double a10 = za[k0 + 4], a11 = za[(size_t)8 * LD + k0 + 4];
                        const double b0 = ub[k0 * PANEL_LD], b1 = ub[(k0 + 4) * PANEL_LD];
                        dmma884(acc[0][0][0], acc[0][0][1], a00, b0);
                        dmma884(acc[0][1][0], acc[0][1][1], a01, b0);
                        dmma884(acc[1][0][0], acc[1][0][1], a10, b1);
                        dmma884(acc[1][1][0], acc[1][1][1], a11, b1);
                    }
                    if (ks < ks1) {
                        const int k0 = 4 * ks;
                        const double a00 = za[k0], a01 = za[(size_t)8 * LD + k0];
                        const double b0 = ub[k0 * PANEL_LD];
                        dmma884(acc[0][0][0], acc[0][0][1], a00, b0);
                        dmma884(acc[0][1][0], acc[0][1][1], a01, b0);
                    }
                    // fold into the running row dots: thread holds W[row][8J + 2*(lane&3) + {0,1}]
                    const int c = 8 * J + 2 * (lane & 3);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double* zr = Z + (size_t)(r0 + 8 * h + (lane >> 2)) * LD + c;
                        sacc[h] = fma(acc[0][h][0] + acc[1][h][0], zr[0], fma(acc[0][h][1] + acc[1][h][1], zr[1], sacc[h]));
                    }
                }
            }
            if (active) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double v = sacc[h];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if ((lane & 3) == 0) s_part[kg * TR + r0 + 8 * h + (lane >> 2)] = v;
                }
            }
            __syncthreads();
            if (tid >= seg_start && tid < seg_end) {
                double v = s_part[tid];
#pragma unroll
                for (int g2 = 1; g2 < KG; ++g2) v += s_part[g2 * TR + tid];
                prm.val[t * TR + tid] = v;
            }
            seg_start = seg_end;
            __syncthreads();
        }
    }
}

// per-leaf histogram for the evaluate-only entry point (leaf ids supplied by the caller)
__global__ void __launch_bounds__(256) k_hist(const int* __restrict__ leaf, long long m, long long n_models, int* __restrict__ cnt,
                                              int* __restrict__ rank) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const int lf = leaf[q];
        const bool ok = lf >= 0 && lf < n_models;
        if (rank) rank[q] = ok ? atomicAdd(cnt + lf, 1) : -1;
        else if (ok) atomicAdd(cnt + lf, 1);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static size_t panel_smem(int NP, int TR) {
    const int KG = 16 / (TR / 16);
    return ((size_t)TR * (NP + 4) + (size_t)2 * NP * PANEL_LD + NP + (size_t)KG * TR) * 8 + (size_t)(2 * TR) * 4 + 16;
}

bool csp_dense_supported(const csp_models* mo) {
    if (mo->degree != 2 || mo->n <= 32) return false;
    const int NP = (mo->n + 1 + 7) / 8 * 8;
    // NP <= 128: packed U (<= 72 KB) plus two 64-row Z tiles (<= 132 KB) fit in the 227 KB of shared memory;
    // beyond that the panel kernel needs a 16-row Z tile and two U panels (n up to ~700)
    return NP <= 128 || panel_smem(NP, 16) <= 227 * 1024;
}

int csp_dense_eval(const csp_models* mo, int n_sm, const double* dX, const int2* perm, const int* total, long long m, double* dval,
                   cudaStream_t st) {
    const int n = mo->n;
    DenseParams prm{};
    prm.mv = ModelsView{mo->n, mo->degree, mo->stride, mo->n_models, mo->data};
    prm.X = dX; prm.perm = perm; prm.total = total; prm.val = dval;
    prm.NP = (n + 1 + 7) / 8 * 8;
    prm.LD = prm.NP + 4;  // NP % 8 == 0  =>  LD % 16 in {4, 12}
    prm.bulk = ((n & 1) == 0 && (uintptr_t)dX % 16 == 0) ? 1 : 0;
    const int nt = prm.NP / 8;
    if (prm.NP > 128) {  // U streamed in panels
        auto launchp = [&](auto kern, int TR) -> int {
            const size_t smem = panel_smem(prm.NP, TR);
            const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((m + TR - 1) / TR, std::max(n_sm, 1)));
            CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            {
                CspTimed tm(CSP_K_EVAL_BINNED, st);
                kern<<<grid, 512, smem, st>>>(prm);
            }
            CSP_LAUNCHED();
            CSP_CUDA(cudaGetLastError());
            return CSP_OK;
        };
        if (panel_smem(prm.NP, 64) <= 227 * 1024) return launchp(k_eval_dmma_panel<64>, 64);
        if (panel_smem(prm.NP, 32) <= 227 * 1024) return launchp(k_eval_dmma_panel<32>, 32);
        if (panel_smem(prm.NP, 16) <= 227 * 1024) return launchp(k_eval_dmma_panel<16>, 16);
        return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the DMMA evaluation kernels", n);
    }
    {
        constexpr int JS = 4, TR = 64;
        // One CTA per SM (shared-memory bound).  Default: 16 warps of 16 rows x a column group, every warp issues its share of the
        // gathers and subtracts b in registers.  Measured alternatives (B200, n = 100, 2e7 queries; kernel ms): 10.87 default;
        // CSP_DMMA_LAYOUT=320 (8 warps of 32 rows: 31 % fewer shared-memory wavefronts per DMMA) 11.78; CSP_DMMA_LAYOUT=32 (the
        // same plus 4 producer warps that issue the gathers and form z in place one tile ahead) 13.26: the producers finish
        // after the compute warps.  tools/ubench/slab_peak.cu times the block alone: 24.5 TFLOP/s issued as used here, 30.0 with
        // z resident and no tile barrier, against 36-37 for an unstructured register-blocked DMMA loop (fp64_peaks.cu): the
        // short triangular k-loop (13 k-blocks at n = 100) with its per-slab prologue, row-dot epilogue and barrier is what
        // keeps the pipe at 0.57.  Also measured and not kept: one free-running tile pipeline per row group (the four warps that
        // share 16 rows, own Z slices / mbarriers / 128-thread named barrier, CTA barriers only on tiles with a leaf change):
        // 11.3-11.5 ms at n = 100 and 7.2-7.4 against 6.8 at n = 64, 14.7 against 15.9 at n = 126.
        const size_t smem = ((size_t)dmma_ubase(nt, prm.LD) + (size_t)2 * TR * prm.LD + prm.NP + 2 * JS * TR) * 8 + (size_t)(2 * TR) * 4 + 16;
        const char* lenv = getenv("CSP_DMMA_LAYOUT");
        const int layout = lenv ? atoi(lenv) : 16;
        const int nthreads = layout == 16 ? 512 : (layout == 320 ? 256 : 384);
        const unsigned grid2 = (unsigned)std::max<long long>(1, std::min<long long>((m + TR - 1) / TR, std::max(n_sm, 1)));
        auto launch2 = [&](auto kern) -> int {
            const bool prof = getenv("CSP_DMMA_PROF") && atoi(getenv("CSP_DMMA_PROF"));
            if (prof) CSP_CUDA(cudaMalloc(&prm.prof, (size_t)grid2 * 8 * sizeof(long long)));
            CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            {
                CspTimed tm(CSP_K_EVAL_BINNED, st);
                kern<<<grid2, nthreads, smem, st>>>(prm);
            }
            CSP_LAUNCHED();
            CSP_CUDA(cudaGetLastError());
            if (prof) {  // developer aid: per-phase clock totals, averaged over the CTAs
                std::vector<long long> h((size_t)grid2 * 8);
                CSP_CUDA(cudaMemcpy(h.data(), prm.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(prm.prof);
                static const char* names[7] = {"prefetch", "gather_wait", "segment", "expand_U", "combine+sync", "slab(warp 0)", "slab barrier"};
                for (int k = 0; k < 7; ++k) {
                    double sum = 0;
                    for (unsigned b = 0; b < grid2; ++b) sum += (double)h[(size_t)b * 8 + k];
                    fprintf(stderr, "[csp dmma prof] %-13s %12.0f clk/CTA\n", names[k], sum / grid2);
                }
            }
            return CSP_OK;
        };
        if (smem > 227 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the DMMA evaluation kernel", n);
        auto pick = [&](auto rwc, auto pwc) -> int {
            constexpr int RW = decltype(rwc)::value, PW = decltype(pwc)::value;
            if (nt <= 8) return launch2(k_eval_dmma<TR, RW, 2, JS, PW>);  // nt >= 5 (n > 32)
            if (nt <= 12) return launch2(k_eval_dmma<TR, RW, 3, JS, PW>);
            return launch2(k_eval_dmma<TR, RW, 4, JS, PW>);
        };
        if (layout == 16) return pick(std::integral_constant<int, 16>{}, std::integral_constant<int, 0>{});
        if (layout == 320) return pick(std::integral_constant<int, 32>{}, std::integral_constant<int, 0>{});  // 32-row warps, no producers
        return pick(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
    }
}

int csp_hist_launch(const int* dleaf, long long m, long long n_models, int* cnt, int* rank, int n_sm, cudaStream_t st) {
    const long long blocks = std::min<long long>((m + 255) / 256, (long long)std::max(n_sm, 1) * 8);
    k_hist<<<(unsigned)std::max<long long>(blocks, 1), 256, 0, st>>>(dleaf, m, n_models, cnt, rank);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}
