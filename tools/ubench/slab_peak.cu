// Microbenchmark: the K3b triangular DMMA block (csrc/dmma_slab.cuh) in isolation -- Z and U resident in shared memory, no
// gathers, no record expansion -- to separate what the block's own loop structure costs from what the tile pipeline around it
// costs.  One CTA per SM, the kernel's warp layouts, `iters` slabs per warp, with or without the per-tile CTA barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../coordinatesplittingptrees.jl_b200/csrc -o slab_peak slab_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#include "dmma_slab.cuh"

template <int TR, int RW, int NTJ, int JS, bool SUB, bool BARRIER>
__global__ void __launch_bounds__(TR / RW * JS * 32) k_slab(double* out, int iters, int NP, int LD) {
    extern __shared__ __align__(16) double sm[];
    const int nt = NP / 8, usz = dmma_ubase(nt, LD);
    double* U = sm;
    double* Z = U + usz;
    double* sb = Z + (size_t)TR * LD;
    double* part = sb + NP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < usz; i += blockDim.x) U[i] = 1.0 + 1e-6 * (i % 97);
    for (int i = tid; i < TR * LD; i += blockDim.x) Z[i] = 0.5 + 1e-6 * (i % 89);
    for (int i = tid; i < NP; i += blockDim.x) sb[i] = 0.25;
    __syncthreads();
    int rg, jg;
    if (RW == 16) { rg = warp / JS; jg = (warp + rg) % JS; }
    else { rg = warp >> 2; const int s4 = warp & 3; jg = rg == 0 ? ((s4 & 1) << 1 | (s4 >> 1)) : ((((s4 & 1) << 1) | (s4 >> 1)) ^ 1); }
    for (int it = 0; it < iters; ++it) {
        if ((NTJ - 1) * JS + jg < nt) dmma_slab<NTJ, JS, RW, SUB>(Z, U, sb, part + jg * TR, rg * RW, jg, nt, LD, lane);
        else dmma_slab<NTJ - 1, JS, RW, SUB>(Z, U, sb, part + jg * TR, rg * RW, jg, nt, LD, lane);
        if (BARRIER) __syncthreads();
    }
    if (tid < TR) out[blockIdx.x * TR + tid] = part[tid];
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, (size_t)sms * 64 * 8);
    const int iters = 4000;
    for (int n : {100, 64}) {
        const int NP = (n + 1 + 7) / 8 * 8, LD = NP + 4, nt = NP / 8;
        long long dm = 0;  // DMMAs of one 16-row slab: tile J costs 4 (J + 1)
        for (int J = 0; J < nt; ++J) dm += 4 * (J + 1);
        const size_t smem = ((size_t)dmma_ubase(nt, LD) + (size_t)64 * LD + NP + 4 * 64) * 8;
        auto run = [&](auto kern, int threads, const char* what) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            kern<<<sms, threads, smem>>>(out, 10, NP, LD);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            kern<<<sms, threads, smem>>>(out, iters, NP, LD);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double flop = (double)sms * iters * 4 * dm * 512.0;  // four 16-row slabs per 64-row tile
            printf("n=%3d %-46s %7.2f TFLOP/s issued  (%s)\n", n, what, flop / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
        };
        if (n == 100) {
            run(k_slab<64, 16, 4, 4, true, true>, 512, "16 warps x 16 rows, DADD in loop, tile barrier");
            run(k_slab<64, 16, 4, 4, true, false>, 512, "16 warps x 16 rows, DADD in loop, no barrier");
            run(k_slab<64, 16, 4, 4, false, true>, 512, "16 warps x 16 rows, z resident, tile barrier");
            run(k_slab<64, 16, 4, 4, false, false>, 512, "16 warps x 16 rows, z resident, no barrier");
            run(k_slab<64, 32, 4, 4, true, true>, 256, " 8 warps x 32 rows, DADD in loop, tile barrier");
            run(k_slab<64, 32, 4, 4, false, true>, 256, " 8 warps x 32 rows, z resident, tile barrier");
            run(k_slab<64, 32, 4, 4, false, false>, 256, " 8 warps x 32 rows, z resident, no barrier");
        } else {
            run(k_slab<64, 16, 3, 4, true, true>, 512, "16 warps x 16 rows, DADD in loop, tile barrier");
            run(k_slab<64, 16, 3, 4, false, true>, 512, "16 warps x 16 rows, z resident, tile barrier");
            run(k_slab<64, 32, 3, 4, false, true>, 256, " 8 warps x 32 rows, z resident, tile barrier");
            run(k_slab<64, 32, 3, 4, false, false>, 256, " 8 warps x 32 rows, z resident, no barrier");
        }
    }
    return 0;
}
