// binned_eval.cu -- K3a, large-batch form: model-stationary evaluation after binning the located queries by leaf.
//
// The fused kernel of locate_eval.cu fetches a model record (8(n+1)(n+2)/2 + 8n bytes) from L2 for EVERY query; with
// many queries per leaf that traffic (and the per-query reduction work) dominates.  Here the located queries are
// counting-sorted by leaf id, so that consecutive evaluation positions share a leaf and a record is fetched once per run.
// Pipeline (all on the caller's stream, chunk by chunk):
//   k_locate_only / k_locate_wide   leaf ids in query order + per-leaf histogram (global reductions)   [locate_eval.cu]
//   k_scan                          exclusive prefix sum over the per-leaf counts -> segment offsets (single launch)
//   k_scatter                       slot = offsets[leaf]++ ; perm[slot] = (query, leaf) ; slot_of[query] = slot
//   evaluation in sorted order      k_eval_binned_exact<N> (FMA, small n), k_eval_binned_mma<N> (FP64 tensor cores, even n in
//                                   16..30), k_eval_dmma / k_eval_dmma_panel (n > 32, dense_eval.cu), generic fallbacks;
//                                   values written at the sorted position (coalesced)
//   k_unpermute                     val[q] = sorted_val[slot_of[q]]  (the chunk's values are still in L2)
// Evaluation arithmetic: reference src/cs2.jl:155-159 (summation order not pinned by the reference).
#include <cstdlib>

#include "common.cuh"

__device__ __forceinline__ double dnan_b() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---- exclusive scan of cnt[0..nl) into off[0..nl] (off[nl] = total) -- one launch, chained across blocks ------------
// Block b publishes its sum with an epoch-stamped flag and adds up the sums of the blocks before it.  b is a LOGICAL block id
// taken from an atomic ticket when the block starts running (CUDA does not promise that blockIdx order is dispatch order): a
// block only ever waits for tickets handed out before its own, i.e. for blocks that are already resident.  The block that
// draws the last ticket resets the counter for the next launch.  cnt is zeroed on the way out, ready for the next chunk's
// histogram.
#define SCAN_T 256
#define SCAN_PER 8  // elements per thread => 2048 per block

__device__ __forceinline__ int block_sum(int v, int* ws) {  // all threads get the sum; ws: SCAN_T/32 ints
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < SCAN_T / 32; ++w) t += ws[w];
    return t;
}

__global__ void __launch_bounds__(SCAN_T) k_scan(int* __restrict__ cnt, int nl, int* __restrict__ off, int* __restrict__ cursor,
                                                volatile int* bsum, volatile int* bflag, int* ticket, int epoch) {
    __shared__ int ws[SCAN_T / 32];
    __shared__ int carry;
    __shared__ int s_ticket;
    if (threadIdx.x == 0) {
        const int tk = atomicAdd(ticket, 1);
        if (tk == (int)gridDim.x - 1) *ticket = 0;  // every ticket of this launch has been drawn
        s_ticket = tk;
    }
    __syncthreads();
    const int b = s_ticket, base = b * SCAN_T * SCAN_PER;
    int v[SCAN_PER], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {
        const int i = base + k * SCAN_T + threadIdx.x;
        v[k] = i < nl ? cnt[i] : 0;
        s += v[k];
    }
    const int mine = block_sum(s, ws);
    if (threadIdx.x == 0) {
        bsum[b] = mine;
        __threadfence();
        bflag[b] = epoch;
    }
    int before = 0;
    for (int j = threadIdx.x; j < b; j += SCAN_T) {
        while (bflag[j] != epoch) {}
        __threadfence();
        before += bsum[j];
    }
    before = block_sum(before, ws);
    if (threadIdx.x == 0) carry = before;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCAN_PER; ++k) {  // element order: k-major, then thread
        const int i = base + k * SCAN_T + threadIdx.x;
        int x = v[k];
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        int wbefore = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbefore += ws[w];
        const int ex = carry + wbefore + x - v[k];
        if (i < nl) { off[i] = ex; cursor[i] = ex; cnt[i] = 0; }
        __syncthreads();
        if (threadIdx.x == SCAN_T - 1) carry = ex + v[k];
        __syncthreads();
    }
    if (b == gridDim.x - 1 && threadIdx.x == 0) off[nl] = carry;
}

// ---- scatter: perm[slot] = (query, leaf), slot = cursor[leaf]++ ------------------------------------------------------
// Leaf ids outside [0, n_models) (caller-supplied ids of csp_eval*) are treated like "outside the world": slot -1, value NaN.
// valq != null ("direct" mode): the evaluation kernel writes each value straight to its query's place, so no slot is recorded and
// the queries without a leaf get their NaN here.
__global__ void __launch_bounds__(256) k_scatter(const int* __restrict__ leaf, long long m, int n_models, int* __restrict__ cursor,
                                                 int2* __restrict__ perm, int* __restrict__ slot_of, double* __restrict__ valq) {
    // four independent atomics in flight per thread (the returning atomic's L2 round trip is the critical path)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; q0 < m; q0 += 4 * stride) {
        int lf[4], slot[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long q = q0 + u * stride;
            lf[u] = q < m ? leaf[q] : -2;
            if (lf[u] >= n_models) lf[u] = -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) slot[u] = lf[u] >= 0 ? atomicAdd(cursor + lf[u], 1) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long q = q0 + u * stride;
            if (lf[u] >= 0) perm[slot[u]] = make_int2((int)q, lf[u]);
            if (valq) { if (lf[u] == -1) valq[q] = dnan_b(); }
            else if (lf[u] != -2) slot_of[q] = slot[u];  // -1: outside the world (the reference throws, src/tree.jl:359)
        }
    }
}

// The same without atomics: the query's rank inside its leaf's bin came back from the histogram atomic of the descent kernel
// (slot_of holds it on entry); the sorted position is the bin offset plus that rank.
__global__ void __launch_bounds__(256) k_scatter_ranked(const int* __restrict__ leaf, long long m, int n_models, const int* __restrict__ off,
                                                        int2* __restrict__ perm, int* __restrict__ slot_of) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; q0 < m; q0 += 4 * stride) {
        int lf[4], rk[4], o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long q = q0 + u * stride;
            lf[u] = q < m ? leaf[q] : -2;
            rk[u] = q < m ? slot_of[q] : -1;
            if (lf[u] >= n_models || rk[u] < 0) lf[u] = lf[u] == -2 ? -2 : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = lf[u] >= 0 ? __ldg(off + lf[u]) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long q = q0 + u * stride;
            if (lf[u] >= 0) {
                const int s = o[u] + rk[u];
                perm[s] = make_int2((int)q, lf[u]);
                slot_of[q] = s;
            } else if (lf[u] == -1) {
                slot_of[q] = -1;
            }
        }
    }
}

// values come back from leaf-sorted order to query order: a gather (reads hit L2: the chunk's sorted values were just
// written) followed by coalesced writes -- instead of scattered 8-byte writes, which cost a sector read-modify-write each
__global__ void __launch_bounds__(256) k_unpermute(const int* __restrict__ slot_of, const double* __restrict__ vsorted, long long m,
                                                   double* __restrict__ val) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const int s = slot_of[q];
        val[q] = s >= 0 ? vsorted[s] : dnan_b();
    }
}

// ---- model-stationary evaluation ---------------------------------------------------------------------------------
// One thread per binned query.  NPAD >= n is a compile-time bound so that z stays in registers (loops fully unrolled and
// guarded by i < n); record loads are warp-uniform whenever the warp's queries share a leaf.
template <int NPAD>
__global__ void __launch_bounds__(256) k_eval_binned(const ModelsView mv, const double* __restrict__ X, const int2* __restrict__ perm,
                                                    const int* __restrict__ total, double* __restrict__ val) {
    const int n = mv.n;
    const long long tot = *total;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < tot; s += (long long)gridDim.x * blockDim.x) {
        const int2 ql = perm[s];
        const double* __restrict__ M = mv.data + (size_t)ql.y * mv.stride;
        const double* __restrict__ x = X + (size_t)(unsigned)ql.x * n;
        double z[NPAD];
        if ((n & 1) == 0) {  // rows are 16-byte aligned when n is even and X is (checked by the host): 128-bit gathers
#pragma unroll
            for (int i = 0; i < NPAD; i += 2)
                if (i < n) {
                    const double2 xv = __ldg((const double2*)(x + i));
                    z[i] = xv.x - __ldg(M + i);
                    if (i + 1 < NPAD) z[i + 1] = xv.y - __ldg(M + i + 1);
                }
        } else {
#pragma unroll
            for (int i = 0; i < NPAD; ++i)
                if (i < n) z[i] = __ldg(x + i) - __ldg(M + i);
        }
        const double* __restrict__ A = M + n;
        double acc = 0.0;
        if (mv.degree == 2) {
#pragma unroll
            for (int i = 0; i < NPAD; ++i)
                if (i < n) {
                    const double* __restrict__ Ar = A + (i * (n + 1) - i * (i - 1) / 2);  // row i of the packed triangle
                    double row = __ldg(Ar + (n - i));                                     // A[i][n] * 1
#pragma unroll
                    for (int j = i; j < NPAD; ++j)
                        if (j < n) row = fma(__ldg(Ar + (j - i)), z[j], row);
                    acc = fma(row, z[i], acc);
                }
            acc += __ldg(A + ((n + 1) * (n + 2) / 2 - 1));  // A[n][n] = c
        } else {
#pragma unroll
            for (int i = 0; i < NPAD; ++i)
                if (i < n) acc = fma(__ldg(A + i), z[i], acc);
            acc += __ldg(A + n);
        }
        val[s] = acc;  // leaf-sorted position: coalesced; k_unpermute brings it back to query order
    }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// Tensor-core variant for even n in [8, 30] (degree 2).  Delivering every coefficient to every thread through shared memory
// caps the FMA kernel below at a quarter of the FP64 rate (a broadcast 128-bit shared load still costs four register-file
// wavefronts); an MMA takes each coefficient once per 8 rows instead.  A warp walks its contiguous range 32 positions at a
// time as four 8-row groups: Z (A operand, rows = queries) is loaded straight from global memory in fragment layout -- the
// four lanes of a row read 32 contiguous bytes per k-step, i.e. whole sectors of the x row -- with b subtracted in
// registers; U (B operand) fragments and the b entries a lane needs live in registers for as long as the leaf's run lasts
// and are re-read from the staged record (same double-buffered cp.async slots as below) only when the leaf changes.
// W = Z U, then the row dot with z (re-read from L1) and a quad reduction.  A group that straddles two leaves is computed
// once per leaf with the other rows masked at the store.
// Row stride (doubles) of the staged x rows: 16-byte granular (TMA destination) and (stride in words) mod 32 in {8, 24}, so that the
// 16 lanes of a half-warp -- 4 rows x 4 consecutive doubles -- read 16 distinct bank pairs when the A fragments are loaded.
__host__ __device__ constexpr int mma_row_stride(int n) {
    int rs = n;  // even
    while (((2 * rs) % 32) != 8 && ((2 * rs) % 32) != 24) rs += 2;
    return rs;
}

template <int N>
__global__ void __launch_bounds__(256) k_eval_binned_mma(const ModelsView mv, const double* __restrict__ X,
                                                        const int2* __restrict__ perm, const int* __restrict__ total,
                                                        double* __restrict__ val, const int direct) {
    static_assert((N & 1) == 0 && N >= 2 && N <= 30, "even n only");
    constexpr int NP = (N + 1 + 7) / 8 * 8, KS = NP / 4, NT = NP / 8;
    constexpr int RS = mma_row_stride(N);
    extern __shared__ __align__(16) double s_rec[];  // per warp: [2][stride] records, [2][32][RS] staged rows; then [warps][2] mbarriers
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, stride = mv.stride, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    const int r = lane >> 2, t = lane & 3;
    const size_t per_warp = (size_t)2 * stride + (size_t)2 * 32 * RS;
    double* slot = s_rec + (size_t)wib * per_warp;
    double* rows = slot + (size_t)2 * stride;
    uint64_t* bars = (uint64_t*)(s_rec + (size_t)nwb * per_warp) + 2 * wib;
    const long long tot = *total;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5, wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long per = ((tot + 31) / 32 + nwarps - 1) / nwarps * 32;
    const long long start = wid * per, end = start + per < tot ? start + per : tot;
    if (start >= end) return;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto fetch = [&](int lf, int b) {
        const double* __restrict__ M = mv.data + (size_t)lf * stride;
        double* dst = slot + (size_t)b * stride;
        for (int k = lane * 2; k < stride; k += 64) cp_async16(dst + k, M + k);
    };
    // x rows of 32 sorted positions -> stage st: one TMA bulk copy per row (each lane gathers its own), all completing on the
    // stage's mbarrier; issued one iteration ahead, so the gather latency (DRAM: these rows are scattered over the whole chunk)
    // is covered by the previous iteration's tensor-core work instead of stalling it
    auto stage_rows = [&](const int2& q, int st) {
        const bool v = q.y >= 0;
        const unsigned vm = __ballot_sync(FULL, v);
        if (lane == 0) mbar_expect_tx(&bars[st], (uint32_t)(__popc(vm) * N * 8));
        __syncwarp();
        // (no proxy fence: the stage being refilled was only READ through the generic proxy, and those reads completed -- their
        // values fed the MMAs -- before the __syncwarp that closed the previous iteration)
        if (v) bulk_g2s(rows + ((size_t)st * 32 + lane) * RS, X + (size_t)(unsigned)q.x * N, N * 8, &bars[st]);
    };
    int cur = 0, staged = -1, pre = -1;
    double uf[KS][NT], bA[KS], bE[NT][2];  // per-leaf fragments (registers)
    long long s0 = start;
    int2 ql = s0 + lane < end ? perm[s0 + lane] : make_int2(0, -1);
    int2 qn = s0 + 32 + lane < end ? perm[s0 + 32 + lane] : make_int2(0, -1);
    stage_rows(ql, 0);
    uint32_t ph0 = 0, ph1 = 0;
    int cs = 0;
    for (; s0 < end; s0 += 32, cs ^= 1) {
        const bool valid = ql.y >= 0;
        if (s0 + 32 < end) stage_rows(qn, cs ^ 1);  // the next iteration's rows start moving now
        const int2 qnn = s0 + 64 + lane < end ? perm[s0 + 64 + lane] : make_int2(0, -1);  // and the one after's perm entries
        if (cs == 0) { mbar_wait(&bars[0], ph0); ph0 ^= 1; }
        else { mbar_wait(&bars[1], ph1); ph1 ^= 1; }
        const double* crow = rows + (size_t)cs * 32 * RS;
        unsigned todo = __ballot_sync(FULL, valid);
        while (todo) {
            const int lf = __shfl_sync(FULL, ql.y, __ffs(todo) - 1);
            const unsigned minemask = __ballot_sync(FULL, valid && ql.y == lf);
            todo &= ~minemask;
            if (lf != staged) {
                if (lf == pre) {
                    cur ^= 1;
                } else {
                    cp_async_wait_all();
                    __syncwarp();
                    fetch(lf, cur);
                }
                staged = lf;
                pre = -1;
                cp_async_wait_all();
                __syncwarp();
                // fragments of the new leaf
                const double* rec = slot + (size_t)cur * stride;
                const double* A = rec + N;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int k = 4 * ks + t;
                    bA[ks] = k < N ? rec[k] : 0.0;
#pragma unroll
                    for (int J = 0; J < NT; ++J) {
                        const int col = 8 * J + r;
                        uf[ks][J] = (J >= ks / 2 && k <= col && col <= N) ? A[k * (N + 1) - k * (k - 1) / 2 + (col - k)] : 0.0;
                    }
                }
#pragma unroll
                for (int J = 0; J < NT; ++J) {
                    const int col = 8 * J + 2 * t;
                    bE[J][0] = col < N ? rec[col] : 0.0;
                    bE[J][1] = col + 1 < N ? rec[col + 1] : 0.0;
                }
            }
            if (pre < 0) {  // start fetching the next distinct leaf
                int nl = -1;
                if (todo) {
                    nl = __shfl_sync(FULL, ql.y, __ffs(todo) - 1);
                } else {
                    const unsigned other = __ballot_sync(FULL, qn.y >= 0 && qn.y != lf);
                    if (other) nl = __shfl_sync(FULL, qn.y, __ffs(other) - 1);
                }
                if (nl >= 0) {
                    fetch(nl, cur ^ 1);
                    pre = nl;
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (!((minemask >> (8 * g)) & 0xffu)) continue;  // warp-uniform: no row of this group belongs to lf
                const bool rm = (minemask >> (8 * g + r)) & 1u;
                const double* srow = crow + (size_t)(8 * g + r) * RS;
                double a[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int k = 4 * ks + t;
                    double v = 0.0;
                    if (k < N) { if (rm) v = srow[k] - bA[ks]; }
                    else if (k == N) v = 1.0;
                    a[ks] = v;
                }
                double w[NT][2];
#pragma unroll
                for (int J = 0; J < NT; ++J) { w[J][0] = 0.0; w[J][1] = 0.0; }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int J = ks / 2; J < NT; ++J) dmma884(w[J][0], w[J][1], a[ks], uf[ks][J]);
                double dot = 0.0;
#pragma unroll
                for (int J = 0; J < NT; ++J) {
                    const int col = 8 * J + 2 * t;
                    double z0 = 0.0, z1 = 0.0;
                    if (col + 1 < N) {
                        if (rm) { const double2 xv = *(const double2*)(srow + col); z0 = xv.x - bE[J][0]; z1 = xv.y - bE[J][1]; }
                    } else if (col == N) {
                        z0 = 1.0;
                    }
                    dot = fma(w[J][0], z0, fma(w[J][1], z1, dot));
                }
                dot += __shfl_xor_sync(FULL, dot, 1);
                dot += __shfl_xor_sync(FULL, dot, 2);
                const int rq = __shfl_sync(FULL, ql.x, 8 * g + r);
                if (t == 0 && rm) val[direct ? (long long)(unsigned)rq : s0 + 8 * g + r] = dot;
            }
        }
        __syncwarp();  // every lane is done with stage cs before the next iteration refills it
        ql = qn;
        qn = qnn;
    }
    cp_async_wait_all();
}

// Compile-time n.  A warp owns a CONTIGUOUS range of binned positions and walks it 32 at a time.  The x rows are gathered
// straight into registers (128-bit loads).  Leaf records go through two per-warp shared-memory slots filled by 16-byte
// cp.async copies: the record of the leaf in use stays put while the run continues over the following iterations, and the
// next distinct leaf's record (known from the sorted perm entries already in registers) is fetched while the current one
// is being evaluated, so the record latency is off the critical path.  Records are read back as broadcast 128-bit loads.
template <int N>
__global__ void __launch_bounds__(256) k_eval_binned_exact(const ModelsView mv, const double* __restrict__ X,
                                                          const int2* __restrict__ perm, const int* __restrict__ total,
                                                          double* __restrict__ val, const int direct) {
    constexpr int NC = (N + 1) * (N + 2) / 2;  // coefficients of the packed (N+1)x(N+1) triangle
    constexpr bool VEC = (N & 1) == 0;         // A block and x rows are 16-byte aligned
    extern __shared__ __align__(16) double s_rec[];  // [warps][2][stride]
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, stride = mv.stride;
    double* slot = s_rec + (size_t)(threadIdx.x >> 5) * 2 * stride;
    const long long tot = *total;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5, wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long per = ((tot + 31) / 32 + nwarps - 1) / nwarps * 32;  // positions per warp, a multiple of 32
    const long long start = wid * per, end = start + per < tot ? start + per : tot;
    if (start >= end) return;
    auto fetch = [&](int lf, int b) {  // whole record (stride % 4 == 0 doubles, 32-byte aligned) into slot b
        const double* __restrict__ M = mv.data + (size_t)lf * stride;
        double* dst = slot + (size_t)b * stride;
        for (int k = lane * 2; k < stride; k += 64) cp_async16(dst + k, M + k);
    };
    int cur = 0, staged = -1, pre = -1;  // slot in use, leaf it holds, leaf in flight into the other slot (warp-uniform)
    long long s = start + lane;
    int2 ql = s < end ? perm[s] : make_int2(0, -1);
    for (; s - lane < end; s += 32) {
        const bool valid = ql.y >= 0;
        const long long sn = s + 32;
        const int2 qn = sn < end ? perm[sn] : make_int2(0, -1);  // the next 32 positions, in flight during this iteration
        double x[N + 1];
        if (valid) {
            const double* __restrict__ xp = X + (size_t)(unsigned)ql.x * N;
            if constexpr (VEC) {
#pragma unroll
                for (int i = 0; i < N; i += 2) {
                    const double2 v = __ldg((const double2*)(xp + i));
                    x[i] = v.x;
                    x[i + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = __ldg(xp + i);
            }
        }
        unsigned todo = __ballot_sync(FULL, valid);
        double result = 0.0;
        while (todo) {
            const int lf = __shfl_sync(FULL, ql.y, __ffs(todo) - 1);
            const bool mine = valid && ql.y == lf;
            todo &= ~__ballot_sync(FULL, mine);
            // bring lf into the current slot
            if (lf != staged) {
                if (lf == pre) {
                    cur ^= 1;
                } else {  // nothing useful was prefetched (first leaf of the range)
                    cp_async_wait_all();
                    __syncwarp();
                    fetch(lf, cur);
                }
                staged = lf;
                pre = -1;
                cp_async_wait_all();
                __syncwarp();
            }
            // start fetching the next distinct leaf: among the remaining lanes of this iteration, else in the next 32 positions
            if (pre < 0) {
                int nl = -1;
                if (todo) {
                    nl = __shfl_sync(FULL, ql.y, __ffs(todo) - 1);
                } else {
                    const unsigned other = __ballot_sync(FULL, qn.y >= 0 && qn.y != lf);
                    if (other) nl = __shfl_sync(FULL, qn.y, __ffs(other) - 1);
                }
                if (nl >= 0) {
                    fetch(nl, cur ^ 1);
                    pre = nl;
                }
            }
            if (mine) {
                const double* rec = slot + (size_t)cur * stride;
                // z = x - b in place (a thread is `mine` exactly once per iteration)
                if constexpr (VEC) {
#pragma unroll
                    for (int i = 0; i < N; i += 2) {
                        const double2 b = *(const double2*)(rec + i);
                        x[i] -= b.x;
                        x[i + 1] -= b.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] -= rec[i];
                }
                x[N] = 1.0;
                const double* A = rec + N;
                // Row i of the triangle: sum_j A_ij z_j, then times z_i.  Each row's sum runs as four interleaved partial chains
                // (and the rows alternate between two outer accumulators), so the dependent-FMA latency overlaps.
                double acc0 = 0.0, acc1 = 0.0, rp[4] = {0.0, 0.0, 0.0, 0.0};
                int i = 0, j = 0;  // all static after unrolling
#pragma unroll
                for (int k = 0; k < NC; k += (VEC ? 2 : 1)) {
                    double a0, a1 = 0.0;
                    if constexpr (VEC) {
                        const double2 av = *(const double2*)(A + k);  // the element past the last coefficient is record padding
                        a0 = av.x;
                        a1 = av.y;
                    } else {
                        a0 = A[k];
                    }
#pragma unroll
                    for (int u = 0; u < (VEC ? 2 : 1); ++u) {
                        if (k + u < NC) {
                            const double a = u ? a1 : a0;
                            const int part = (j - i) & 3;
                            rp[part] = (j - i < 4) ? a * x[j] : fma(a, x[j], rp[part]);
                            if (j == N) {
                                const int len = N + 1 - i;
                                double row = rp[0];
                                if (len > 1) row += rp[1];
                                if (len > 2) row += (len > 3) ? rp[2] + rp[3] : rp[2];
                                if (i & 1) acc1 = fma(row, x[i], acc1);
                                else acc0 = fma(row, x[i], acc0);
                                ++i;
                                j = i;
                            } else {
                                ++j;
                            }
                        }
                    }
                }
                result = acc0 + acc1;
            }
            __syncwarp();
        }
        if (valid) val[direct ? (long long)(unsigned)ql.x : s] = result;  // sorted position (coalesced; k_unpermute follows) or the query's own
        ql = qn;
    }
    cp_async_wait_all();  // a prefetch may still be in flight
}

// any n: z recomputed on the fly (two extra loads per term)
__global__ void __launch_bounds__(256) k_eval_binned_generic(const ModelsView mv, const double* __restrict__ X,
                                                            const int2* __restrict__ perm, const int* __restrict__ total,
                                                            double* __restrict__ val) {
    const int n = mv.n;
    const long long tot = *total;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < tot; s += (long long)gridDim.x * blockDim.x) {
        const int2 ql = perm[s];
        const double* __restrict__ M = mv.data + (size_t)ql.y * mv.stride;
        const double* __restrict__ x = X + (size_t)(unsigned)ql.x * n;
        const double* __restrict__ A = M + n;
        double acc = 0.0;
        if (mv.degree == 2) {
            long long k = 0;
            for (int i = 0; i < n; ++i) {
                double row = 0.0;
                for (int j = i; j < n; ++j, ++k) row = fma(__ldg(A + k), __ldg(x + j) - __ldg(M + j), row);
                row += __ldg(A + k);
                ++k;
                acc = fma(row, __ldg(x + i) - __ldg(M + i), acc);
            }
            acc += __ldg(A + k);
        } else {
            for (int i = 0; i < n; ++i) acc = fma(__ldg(A + i), __ldg(x + i) - __ldg(M + i), acc);
            acc += __ldg(A + n);
        }
        val[s] = acc;  // leaf-sorted position: coalesced; k_unpermute brings it back to query order
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct BinWorkspace {
    int* cnt = nullptr;      // [nl]
    int* off = nullptr;      // [nl + 1]; off[nl] = number of located queries
    int* cursor = nullptr;   // [nl]
    int* partial = nullptr;  // [2*nblocks + 1]: block sums, epoch flags, ticket counter
    int2* perm = nullptr;    // [m]
    int* slot = nullptr;     // [m] sorted position of each query (-1: outside the world)
    double* vsorted = nullptr;  // [m] values in leaf-sorted order
    int* leaf = nullptr;     // [m] when the caller did not ask for leaf ids
    int epoch = 0;
    bool ranked = false;     // slot already holds each query's rank inside its leaf's bin (returning histogram atomics)
};

int csp_bin_alloc(int nl, long long m, bool need_leaf, cudaStream_t st, BinWorkspace* w) {
    const int nblocks = (nl + SCAN_T * SCAN_PER - 1) / (SCAN_T * SCAN_PER);
    CSP_CUDA(cudaMallocAsync(&w->cnt, (size_t)nl * 4, st));
    CSP_CUDA(cudaMallocAsync(&w->off, (size_t)(nl + 1) * 4, st));
    CSP_CUDA(cudaMallocAsync(&w->cursor, (size_t)nl * 4, st));
    CSP_CUDA(cudaMallocAsync(&w->partial, ((size_t)std::max(nblocks, 1) * 8 + 8), st));
    CSP_CUDA(cudaMemsetAsync(w->partial, 0, ((size_t)std::max(nblocks, 1) * 8 + 8), st));
    CSP_CUDA(cudaMallocAsync(&w->perm, (size_t)m * sizeof(int2), st));
    CSP_CUDA(cudaMallocAsync(&w->slot, (size_t)m * 4, st));
    CSP_CUDA(cudaMallocAsync(&w->vsorted, (size_t)m * 8, st));
    if (need_leaf) CSP_CUDA(cudaMallocAsync(&w->leaf, (size_t)m * 4, st));
    CSP_CUDA(cudaMemsetAsync(w->cnt, 0, (size_t)nl * 4, st));
    return CSP_OK;
}
// Persistent (synchronously allocated) workspace for the host-buffer path, which reuses it across chunks and calls.
int csp_bin_alloc_sync(int nl, long long m, BinWorkspace* w) {
    const int nblocks = (nl + SCAN_T * SCAN_PER - 1) / (SCAN_T * SCAN_PER);
    CSP_CUDA(cudaMalloc(&w->cnt, (size_t)nl * 4));
    CSP_CUDA(cudaMalloc(&w->off, (size_t)(nl + 1) * 4));
    CSP_CUDA(cudaMalloc(&w->cursor, (size_t)nl * 4));
    CSP_CUDA(cudaMalloc(&w->partial, ((size_t)std::max(nblocks, 1) * 8 + 8)));
    CSP_CUDA(cudaMalloc(&w->perm, (size_t)m * sizeof(int2)));
    CSP_CUDA(cudaMalloc(&w->slot, (size_t)m * 4));
    CSP_CUDA(cudaMalloc(&w->vsorted, (size_t)m * 8));
    CSP_CUDA(cudaMemset(w->cnt, 0, (size_t)nl * 4));
    CSP_CUDA(cudaMemset(w->partial, 0, ((size_t)std::max(nblocks, 1) * 8 + 8)));
    w->epoch = 0;
    return CSP_OK;
}
void csp_bin_free_sync(BinWorkspace* w) {
    void* ptrs[] = {w->cnt, w->off, w->cursor, w->partial, w->perm, w->slot, w->vsorted, w->leaf};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    *w = BinWorkspace();
}
void csp_bin_free(BinWorkspace* w, cudaStream_t st) {
    void* ptrs[] = {w->cnt, w->off, w->cursor, w->partial, w->perm, w->slot, w->vsorted, w->leaf};
    for (void* p : ptrs)
        if (p) cudaFreeAsync(p, st);
}

// dense_eval.cu
bool csp_dense_supported(const csp_models* mo);
int csp_dense_eval(const csp_models* mo, int n_sm, const double* dX, const int2* perm, const int* total, long long m, double* dval,
                   cudaStream_t st);

// After k_locate<P,-1> filled leaf[] and cnt[]: scan, scatter, evaluate.
int csp_binned_eval(const csp_models* mo, int n_sm, const double* dX, long long m, const int* dleaf, double* dval, BinWorkspace* w,
                    cudaStream_t st) {
    const int nl = (int)mo->n_models;
    const int nblocks = (nl + SCAN_T * SCAN_PER - 1) / (SCAN_T * SCAN_PER);
    // "direct": the exact-n evaluation kernels write val[query] themselves (scattered 8-byte stores that merge in L2 while the
    // chunk's value range fits there) instead of sorted values + a k_unpermute pass; see DESIGN.md for the measurements
    const int n_ = mo->n;
    const bool mma_on = !(getenv("CSP_EVAL_MMA") && atoi(getenv("CSP_EVAL_MMA")) == 0);
    const bool exact_kernel = mo->degree == 2 && !csp_dense_supported(mo) && (((uintptr_t)dX % 16 == 0) || (n_ & 1)) &&
                              (n_ == 2 || n_ == 3 || n_ == 4 || n_ == 5 || n_ == 6 || n_ == 8 || n_ == 10 || n_ == 12 || n_ == 16 || n_ == 20 ||
                               (mma_on && (n_ == 24 || n_ == 30)));
    const char* denv = getenv("CSP_EVAL_DIRECT_WRITE");
    const bool direct = exact_kernel && !w->ranked && (denv ? atoi(denv) != 0 : false);
    w->epoch += 1;
    {
        CspTimed tm(CSP_K_SCAN, st);
        k_scan<<<nblocks, SCAN_T, 0, st>>>(w->cnt, nl, w->off, w->cursor, w->partial, w->partial + nblocks, w->partial + 2 * nblocks, w->epoch);
    }
    CSP_LAUNCHED();
    const long long blocks = std::min<long long>((m + 255) / 256, (long long)std::max(n_sm, 1) * (8 / g_csp_grid_div));
    {
        CspTimed tm(CSP_K_SCATTER, st);
        if (w->ranked) k_scatter_ranked<<<(unsigned)blocks, 256, 0, st>>>(dleaf, m, nl, w->off, w->perm, w->slot);
        else k_scatter<<<(unsigned)blocks, 256, 0, st>>>(dleaf, m, nl, w->cursor, w->perm, w->slot, direct ? dval : nullptr);
    }
    CSP_LAUNCHED();
    auto unpermute = [&]() -> int {
        CspTimed tm(CSP_K_SCATTER, st);
        k_unpermute<<<(unsigned)blocks, 256, 0, st>>>(w->slot, w->vsorted, m, dval);
        CSP_LAUNCHED();
        CSP_CUDA(cudaGetLastError());
        return CSP_OK;
    };
    if (csp_dense_supported(mo)) {  // K3b: FP64 tensor cores
        CSP_CHECK(csp_dense_eval(mo, n_sm, dX, w->perm, w->off + nl, m, w->vsorted, st));
        return unpermute();
    }
    ModelsView mv{mo->n, mo->degree, mo->stride, mo->n_models, mo->data};
    const int n = mo->n;
    const bool aligned = ((uintptr_t)dX % 16 == 0);
    size_t dyn_smem = 0;  // only the exact-n kernels stage records in shared memory (8 warps x one record)
    auto launch = [&](auto kern) {
        if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
        CspTimed tm(CSP_K_EVAL_BINNED, st);
        kern<<<(unsigned)blocks, 256, dyn_smem, st>>>(mv, dX, w->perm, w->off + nl, direct ? dval : w->vsorted, direct ? 1 : 0);
        CSP_LAUNCHED();
    };
    auto launch_plain = [&](auto kern) {  // kernels without the direct-write mode
        CspTimed tm(CSP_K_EVAL_BINNED, st);
        kern<<<(unsigned)blocks, 256, 0, st>>>(mv, dX, w->perm, w->off + nl, w->vsorted);
        CSP_LAUNCHED();
    };
    const bool vec_ok = aligned || (n & 1);
    bool done = false;
    if (mo->degree == 2 && vec_ok) {
        done = true;
        // 8 warps x two record slots.  (Staging the x rows through shared memory with TMA, as k_eval_binned_mma does, was measured
        // on this kernel too: n = 10, 1e8 queries 4.06 -> 4.42 ms -- it is DRAM-bound already, the extra pass only adds work.)
        dyn_smem = (size_t)8 * 2 * mo->stride * sizeof(double);
        const bool mma = mma_on;
        // FP64 tensor cores from n = 16 up (measured, 1e8 queries: n = 20 10.4 -> 7.5 ms; n = 10 4.0 -> 7.7 ms: too little MMA work
        // per 8-row group to pay for the fragment bookkeeping, so small n keeps the FMA kernel)
        if (mma && (n == 16 || n == 20 || n == 24 || n == 30)) {
            // per warp: two record slots + two stages of 32 staged rows; as many warps per CTA as keep two CTAs on an SM
            const size_t per_warp = ((size_t)2 * mo->stride + (size_t)2 * 32 * mma_row_stride(n)) * sizeof(double);
            int warps = 8;
            while (warps > 2 && 2 * (warps * (per_warp + 16) + 1024) > 226 * 1024) warps -= 2;
            const size_t smem = warps * (per_warp + 16);
            const long long mblocks = std::min<long long>((m + 32 * warps - 1) / (32 * warps), (long long)std::max(n_sm, 1) * std::max(1, 2 / g_csp_grid_div));
            auto launch_mma = [&](auto kern) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                CspTimed tm(CSP_K_EVAL_BINNED, st);
                kern<<<(unsigned)mblocks, warps * 32, smem, st>>>(mv, dX, w->perm, w->off + nl, direct ? dval : w->vsorted, direct ? 1 : 0);
                CSP_LAUNCHED();
            };
            switch (n) {
                case 16: launch_mma(k_eval_binned_mma<16>); break;
                case 20: launch_mma(k_eval_binned_mma<20>); break;
                case 24: launch_mma(k_eval_binned_mma<24>); break;
                default: launch_mma(k_eval_binned_mma<30>); break;
            }
            CSP_CUDA(cudaGetLastError());
            return direct ? CSP_OK : unpermute();
        }
        switch (n) {
            case 2: launch(k_eval_binned_exact<2>); break;
            case 3: launch(k_eval_binned_exact<3>); break;
            case 4: launch(k_eval_binned_exact<4>); break;
            case 5: launch(k_eval_binned_exact<5>); break;
            case 6: launch(k_eval_binned_exact<6>); break;
            case 8: launch(k_eval_binned_exact<8>); break;
            case 10: launch(k_eval_binned_exact<10>); break;
            case 12: launch(k_eval_binned_exact<12>); break;
            case 16: launch(k_eval_binned_exact<16>); break;
            case 20: launch(k_eval_binned_exact<20>); break;
            default: done = false;
        }
        if (!done) dyn_smem = 0;
    }
    if (done) {}
    else if (n <= 4 && vec_ok) launch_plain(k_eval_binned<4>);
    else if (n <= 8 && vec_ok) launch_plain(k_eval_binned<8>);
    else if (n <= 12 && vec_ok) launch_plain(k_eval_binned<12>);
    else if (n <= 16 && vec_ok) launch_plain(k_eval_binned<16>);
    else if (n <= 20 && vec_ok) launch_plain(k_eval_binned<20>);
    else if (n <= 24 && vec_ok) launch_plain(k_eval_binned<24>);
    else if (n <= 32 && vec_ok) launch_plain(k_eval_binned<32>);
    else launch_plain(k_eval_binned_generic);
    CSP_CUDA(cudaGetLastError());
    return (direct && done) ? CSP_OK : unpermute();
}
