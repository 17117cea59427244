// dense_eval.cu -- K3b: grouped dense evaluation on the FP64 tensor cores (DMMA) for large n.
//
// Replaces modelvalue(c, g, Q, b, x) (reference src/cs2.jl:155-159) when many binned queries share a leaf's dense model:
// with z = (x - b, 1, 0...) and U the upper-triangular (n+1)x(n+1) matrix of the model record (common.cuh), the values of
// a tile of queries are  diag(Z U Z')  =  rowsum((Z U) .* Z):  one TR x NP x NP triangular GEMM per tile plus a row dot.
// Blackwell's tcgen05 has no FP64 kind; the FP64 tensor path on sm_100a is the warp-level mma.sync (SASS DMMA), here
// m8n8k4.  Queries arrive leaf-sorted from the binning pipeline (binned_eval.cu), so consecutive tiles of a CTA share a
// leaf and U is expanded into shared memory only when the leaf changes.
//
// Shared memory: U[NP][LD] (zero below the diagonal / in the padding), Z[TR][LD], LD = NP + 4 with LD % 16 in {4, 12} so
// that both fragment loads are conflict-free (rows 4 apart in 8-byte banks).  A warp owns 16 rows x all NP columns
// (accumulators in registers), so every U fragment it loads feeds two MMAs; k-steps below a column tile's diagonal are
// skipped (half the MMAs).
#include "common.cuh"

__device__ __forceinline__ double dnan_d() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct DenseParams {
    ModelsView mv;
    const double* X;
    const int2* perm;
    const int* total;
    double* val;
    int NP, LD;  // padded dimension (multiple of 8, >= n + 1) and leading dimension of the shared tiles
    int bulk;    // rows are 16-byte aligned multiples of 16 bytes: gather them with TMA bulk copies
};

template <int TR, int NT>  // TR rows per tile (TR/16 warps); NT >= NP/8 column tiles (compile-time bound for the accumulators)
__global__ void __launch_bounds__(TR * 2) k_eval_dmma(const DenseParams prm) {
    extern __shared__ __align__(16) double sm[];
    const int n = prm.mv.n, NP = prm.NP, LD = prm.LD, nt = NP / 8;
    double* U = sm;                        // [NP][LD]
    double* Z = U + (size_t)NP * LD;       // [TR][LD]
    double* sb = Z + (size_t)TR * LD;      // [n]
    int* s_q = (int*)(sb + ((n + 1) & ~1));  // [TR]
    int* s_leaf = s_q + TR;                // [TR]
    int* s_seg = s_leaf + TR;              // [1] (+1 pad)
    uint64_t* bar = (uint64_t*)(s_seg + 2);
    uint32_t phase = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tot = *prm.total;
    const long long ntiles = (tot + TR - 1) / TR;
    // contiguous tile range per CTA: consecutive tiles mostly share a leaf
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long t0 = (long long)blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
    int cur_leaf = -1;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (long long t = t0; t < t1; ++t) {
        const long long s0 = t * TR;
        const int rows = (int)(tot - s0 < TR ? tot - s0 : TR);
        __syncthreads();  // previous tile fully consumed
        for (int r = tid; r < TR; r += blockDim.x) {
            const int2 ql = r < rows ? prm.perm[s0 + r] : make_int2(0, -1);
            s_q[r] = ql.x;
            s_leaf[r] = ql.y;
        }
        __syncthreads();
        int seg_start = 0;
        while (seg_start < rows) {
            const int lf = s_leaf[seg_start];
            if (tid == 0) *s_seg = rows;
            __syncthreads();
            for (int r = seg_start + 1 + tid; r < rows; r += blockDim.x)  // end of the run of `lf`: first row with another leaf
                if (s_leaf[r] != lf) atomicMin(s_seg, r);
            __syncthreads();
            const int seg_end = *s_seg;
            const double* __restrict__ M = prm.mv.data + (size_t)lf * prm.mv.stride;
            if (lf != cur_leaf) {  // expand the record: b, and U (zero below the diagonal and in the padding)
                for (int i = tid; i < NP * LD; i += blockDim.x) U[i] = 0.0;
                for (int i = tid; i < n; i += blockDim.x) sb[i] = __ldg(M + i);
                __syncthreads();
                const double* __restrict__ A = M + n;
                const int nwarps = blockDim.x >> 5;
                for (int i = warp; i <= n; i += nwarps) {  // row i of the packed triangle: n + 1 - i coefficients, coalesced
                    const double* __restrict__ Ar = A + (i * (n + 1) - i * (i - 1) / 2);
                    double* Ur = U + (size_t)i * LD + i;
                    for (int e = lane; e < n + 1 - i; e += 32) Ur[e] = __ldg(Ar + e);
                }
                cur_leaf = lf;
            }
            // Z rows of the segment: z = (x - b, 1, 0 ...).  The rows are gathered by one TMA bulk copy each (all in
            // flight at once, completion on an mbarrier), then fixed up in place.
            const int nseg = seg_end - seg_start;
            if (prm.bulk) {
                fence_proxy_async();
                if (tid == 0) mbar_expect_tx(bar, (uint32_t)((size_t)nseg * n * 8));
                __syncthreads();
                for (int r = seg_start + tid; r < seg_end; r += blockDim.x)
                    bulk_g2s(Z + (size_t)r * LD, prm.X + (size_t)(unsigned)s_q[r] * n, (uint32_t)(n * 8), bar);
                mbar_wait(bar, phase);
                phase ^= 1;
                for (int rr = warp; rr < nseg; rr += (int)(blockDim.x >> 5)) {  // one warp per row, lanes over columns
                    double* zr = Z + (size_t)(seg_start + rr) * LD;
                    for (int k = lane; k < LD; k += 32) zr[k] = k < n ? zr[k] - sb[k] : (k == n ? 1.0 : 0.0);
                }
            } else {
                for (int idx = tid; idx < nseg * LD; idx += blockDim.x) {
                    const int rr = idx / LD, k = idx - rr * LD;
                    const int r = seg_start + rr;
                    double z = 0.0;
                    if (k < n) z = __ldg(prm.X + (size_t)(unsigned)s_q[r] * n + k) - sb[k];
                    else if (k == n) z = 1.0;
                    Z[(size_t)r * LD + k] = z;
                }
            }
            __syncthreads();
            // this warp's 16 rows: skip if they do not intersect the segment (warp-uniform)
            const int r0 = warp * 16;
            if (r0 < seg_end && r0 + 16 > seg_start) {
                double acc[2][NT][2];
#pragma unroll
                for (int J = 0; J < NT; ++J) { acc[0][J][0] = acc[0][J][1] = acc[1][J][0] = acc[1][J][1] = 0.0; }
                const double* za = Z + (size_t)(r0 + (lane >> 2)) * LD + (lane & 3);
                const double* ub = U + (size_t)(lane & 3) * LD + (lane >> 2);
                for (int k0 = 0; k0 < NP; k0 += 4) {
                    const double a0 = za[k0], a1 = za[(size_t)8 * LD + k0];
                    const int jmin = k0 >> 3;  // column tiles left of the diagonal block hold only zeros for these k
#pragma unroll
                    for (int J = 0; J < NT; ++J) {
                        if (J >= jmin && J < nt) {
                            const double b = ub[(size_t)k0 * LD + 8 * J];
                            dmma884(acc[0][J][0], acc[0][J][1], a0, b);
                            dmma884(acc[1][J][0], acc[1][J][1], a1, b);
                        }
                    }
                }
                // row dot with Z: thread holds W[row][8J + 2*(lane&3) + {0,1}]
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = r0 + 8 * h + (lane >> 2);
                    const double* zr = Z + (size_t)row * LD + 2 * (lane & 3);
                    double s = 0.0;
#pragma unroll
                    for (int J = 0; J < NT; ++J)
                        if (J < nt) s = fma(acc[h][J][0], zr[8 * J], fma(acc[h][J][1], zr[8 * J + 1], s));
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if ((lane & 3) == 0 && row >= seg_start && row < seg_end) prm.val[s0 + row] = s;  // leaf-sorted position
                }
            }
            seg_start = seg_end;
            __syncthreads();  // Z rows are rebuilt for the next segment
        }
    }
}

// per-leaf histogram for the evaluate-only entry point (leaf ids supplied by the caller)
__global__ void __launch_bounds__(256) k_hist(const int* __restrict__ leaf, long long m, long long n_models, int* __restrict__ cnt) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < m; q += (long long)gridDim.x * blockDim.x) {
        const int lf = leaf[q];
        if (lf >= 0 && lf < n_models) atomicAdd(cnt + lf, 1);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
bool csp_dense_supported(const csp_models* mo) {
    if (mo->degree != 2 || mo->n <= 32) return false;
    const int NP = (mo->n + 1 + 7) / 8 * 8;
    return NP <= 128;  // U (NP x (NP+4)) plus a 64-row Z tile must fit in 220 KB of shared memory
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
    auto smem_for = [&](int TR) { return ((size_t)prm.NP * prm.LD + (size_t)TR * prm.LD + ((n + 1) & ~1)) * 8 + (size_t)(2 * TR + 4) * 4 + 16; };
    const long long max_tiles = (m + 63) / 64;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(max_tiles, std::max(n_sm, 1)));
    auto launch = [&](auto kern, int TR) -> int {
        const size_t smem = smem_for(TR);
        CSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            CspTimed tm(CSP_K_EVAL_BINNED, st);
            kern<<<grid, TR * 2, smem, st>>>(prm);
        }
        CSP_LAUNCHED();
        CSP_CUDA(cudaGetLastError());
        return CSP_OK;
    };
    const int nt = prm.NP / 8;
    if (nt <= 9) return smem_for(128) <= 220 * 1024 ? launch(k_eval_dmma<128, 9>, 128) : launch(k_eval_dmma<64, 9>, 64);
    if (nt <= 13) return smem_for(128) <= 220 * 1024 ? launch(k_eval_dmma<128, 13>, 128) : launch(k_eval_dmma<64, 13>, 64);
    if (smem_for(64) > 220 * 1024) return csp_set_error(CSP_ERR_UNSUPPORTED, "n=%d is too large for the DMMA evaluation kernel", n);
    return launch(k_eval_dmma<64, 17>, 64);
}

int csp_hist_launch(const int* dleaf, long long m, long long n_models, int* cnt, int n_sm, cudaStream_t st) {
    const long long blocks = std::min<long long>((m + 255) / 256, (long long)std::max(n_sm, 1) * 8);
    k_hist<<<(unsigned)std::max<long long>(blocks, 1), 256, 0, st>>>(dleaf, m, n_models, cnt);
    CSP_LAUNCHED();
    CSP_CUDA(cudaGetLastError());
    return CSP_OK;
}
