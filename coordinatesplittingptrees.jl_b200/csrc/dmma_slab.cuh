// dmma_slab.cuh -- the warp-level triangular DMMA block of K3b (dense_eval.cu); also timed in isolation by
// tools/ubench/slab_peak.cu.
#pragma once

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b);  // mma.sync m8n8k4 f64 (common.cuh / the ubench)

__host__ __device__ __forceinline__ int dmma_ubase(int kb, int LD) { return 8 * (kb * LD - 4 * kb * (kb - 1)); }

// One warp's share of an RW-row slab (RW = 16 or 32): W = z U over its NL column tiles J = jj*JS + jg, then the partial row
// dots W . z.  Column tile J is needed for k < 8 (J + 1) only (U is upper triangular): stage s runs the k range in which
// exactly the tiles jj >= s are live, so the tile loops are unrolled and free of predicates.
// Shared-memory traffic decides this kernel (round 2 profile: at the DMMA pipe's full rate the 16-row version would need 0.87
// wavefronts per cycle per SM, and more than one in the last stage, where a warp re-reads 16 rows of z for a single live column
// tile): every z (A) fragment is loaded by the JS warps of a slab, every U (B) fragment by the TR/RW row groups.  With 32 rows per
// warp a B fragment feeds four DMMAs instead of two: 2.4 -> 1.66 wavefronts per DMMA at n = 100.
template <int NL, int JS, int RW, bool SUB>  // SUB: Z holds raw rows and z = x - b is formed in registers; else Z holds z
__device__ __forceinline__ void dmma_slab(const double* __restrict__ Z, const double* __restrict__ U, const double* __restrict__ sb,
                                          double* __restrict__ part, int r0, int jg, int nt, int LD, int lane) {
    constexpr int RH = RW / 8;  // 8-row halves per warp
    double acc[RH][NL][2];
#pragma unroll
    for (int h = 0; h < RH; ++h)
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) acc[h][jj][0] = acc[h][jj][1] = 0.0;
    const int l3 = lane & 3, l2 = lane >> 2;
    // A fragments: rows r0 + 8 h + l2, columns k0 + l3
    const double* za = Z + (size_t)(r0 + l2) * LD + l3;
    const double* sbk = sb + l3;
    // B fragments: row k0 + l3 of U's 8-row block kb (rows of LD - 8 kb doubles, columns >= 8 kb only), column 8 J + l2.  The
    // pointer advances from block to block by increments (no per-step address arithmetic in the DMMA loop):
    //   base(kb + 1) - base(kb) = 8 rowlen(kb)  [block size]  - 8 l3  [rows get 8 shorter]  - 8  [the kept columns start 8 later]
    int rowlen = LD;
    const double* ub = U + l3 * LD + l2 + 8 * jg;
    int kb = 0;
#pragma unroll
    for (int s2 = 0; s2 < NL; ++s2) {
        int kb_hi = s2 * JS + jg + 1;
        if (kb_hi > nt) kb_hi = nt;
#pragma unroll 2
        for (; kb < kb_hi; ++kb) {
            double a0[RH], a1[RH];
#pragma unroll
            for (int h = 0; h < RH; ++h) { a0[h] = za[(size_t)8 * h * LD]; a1[h] = za[(size_t)8 * h * LD + 4]; }
            if (SUB) {
                const double b0 = sbk[8 * kb], b1 = sbk[8 * kb + 4];
#pragma unroll
                for (int h = 0; h < RH; ++h) { a0[h] -= b0; a1[h] -= b1; }
            }
            const double* ub1 = ub + 4 * rowlen;
            double bf0[NL], bf1[NL];
#pragma unroll
            for (int jj = s2; jj < NL; ++jj) { bf0[jj] = ub[8 * JS * jj]; bf1[jj] = ub1[8 * JS * jj]; }
#pragma unroll
            for (int jj = s2; jj < NL; ++jj)
#pragma unroll
                for (int h = 0; h < RH; ++h) dmma884(acc[h][jj][0], acc[h][jj][1], a0[h], bf0[jj]);
#pragma unroll
            for (int jj = s2; jj < NL; ++jj)
#pragma unroll
                for (int h = 0; h < RH; ++h) dmma884(acc[h][jj][0], acc[h][jj][1], a1[h], bf1[jj]);
            za += 8;
            ub += 8 * rowlen - 8 * l3 - 8;
            rowlen -= 8;
        }
    }
    // thread holds W[row][8J + 2*(lane&3) + {0,1}] for row = r0 + 8h + lane/4
#pragma unroll
    for (int h = 0; h < RH; ++h) {
        const int row = r0 + 8 * h + l2;
        const double* zr = Z + (size_t)row * LD + 2 * l3 + 8 * jg;
        const double* br = sb + 2 * l3 + 8 * jg;
        double sacc = 0.0;
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            double z0 = zr[8 * JS * jj], z1 = zr[8 * JS * jj + 1];
            if (SUB) { z0 -= br[8 * JS * jj]; z1 -= br[8 * JS * jj + 1]; }
            sacc = fma(acc[h][jj][0], z0, fma(acc[h][jj][1], z1, sacc));
        }
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (l3 == 0) part[row] = sacc;
    }
}

