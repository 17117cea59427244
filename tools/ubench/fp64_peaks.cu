// Microbenchmark: raw FP64 throughput of one B200 for (a) mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) and (b) plain DFMA,
// to pick the roofline denominator for the large-n model evaluation kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu && ./fp64_peaks
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a, double b) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// both pipes at once: warps alternate roles
template <int NACC>
__global__ void k_mixed(double* out, int iters, double a, double b) {
    double s = 0;
    if ((threadIdx.x >> 5) & 1) {
        double acc[NACC][2];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-9;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    } else {
        double acc[2 * NACC];
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 2 * NACC; ++i) acc[i] = fma(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Register-blocked DMMA like the evaluation kernel's inner loop: RH a-fragments x NL b-fragments -> RH*NL accumulators.
// MODE 0: operands stay in registers; 1: re-loaded from shared memory every step (conflict-free LDS.64, like the A / B
// fragments); 2: as 1, plus a DADD on every a-fragment (z = x - b formed in registers).
template <int RH, int NL, int MODE>
__global__ void k_dmma_blk(double* out, int iters, double a, double b) {
    __shared__ double sa[64 * 40], sbv[64 * 40];
    for (int i = threadIdx.x; i < 64 * 40; i += blockDim.x) { sa[i] = a + i * 1e-9; sbv[i] = b - i * 1e-9; }
    __syncthreads();
    double acc[RH][NL][2];
#pragma unroll
    for (int h = 0; h < RH; ++h)
#pragma unroll
        for (int j = 0; j < NL; ++j) acc[h][j][0] = acc[h][j][1] = threadIdx.x * 1e-9;
    double av[RH], bv[NL];
#pragma unroll
    for (int h = 0; h < RH; ++h) av[h] = a + h;
#pragma unroll
    for (int j = 0; j < NL; ++j) bv[j] = b + j;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) {
            const double* pa = sa + (it & 63) * 40 + lane;   // 32 consecutive doubles per fragment: conflict-free
            const double* pb = sbv + (it & 63) * 40 + lane;
#pragma unroll
            for (int h = 0; h < RH; ++h) av[h] = MODE == 2 ? pa[h] - pb[4 + (h & 1)] : pa[h];
#pragma unroll
            for (int j = 0; j < NL; ++j) bv[j] = pb[j];
        }
#pragma unroll
        for (int j = 0; j < NL; ++j)
#pragma unroll
            for (int h = 0; h < RH; ++h) dmma884(acc[h][j][0], acc[h][j][1], av[h], bv[j]);
    }
    double s = 0;
#pragma unroll
    for (int h = 0; h < RH; ++h)
#pragma unroll
        for (int j = 0; j < NL; ++j) s += acc[h][j][0] + acc[h][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timed(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, (size_t)sms * 8 * 1024 * 8);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        const int grid = sms * (1024 / threads);
        float ms = timed([&] { k_dmma<8><<<grid, threads>>>(out, iters, 1.0000001, 0.9999999); });
        double flop = (double)grid * (threads / 32) * iters * 8 * 512.0;
        printf("DMMA.8x8x4  grid %4d x %4d thr (%d warps/SM): %7.2f TFLOP/s\n", grid, threads, 32, flop / ms * 1e-9);
        const int grid1 = sms;
        ms = timed([&] { k_dmma<8><<<grid1, threads>>>(out, iters, 1.0000001, 0.9999999); });
        flop = (double)grid1 * (threads / 32) * iters * 8 * 512.0;
        printf("DMMA.8x8x4  grid %4d x %4d thr (%d warps/SM): %7.2f TFLOP/s\n", grid1, threads, threads / 32, flop / ms * 1e-9);
        ms = timed([&] { k_dfma<16><<<grid1, threads>>>(out, iters, 1.0000001, 0.9999999); });
        flop = (double)grid1 * threads * iters * 16 * 2.0;
        printf("DFMA        grid %4d x %4d thr (%d warps/SM): %7.2f TFLOP/s\n", grid1, threads, threads / 32, flop / ms * 1e-9);
        ms = timed([&] { k_mixed<8><<<grid1, threads>>>(out, iters, 1.0000001, 0.9999999); });
        flop = (double)grid1 * (threads / 64) * iters * (8 * 512.0 + 32.0 * 16 * 2.0);
        printf("DMMA+DFMA   grid %4d x %4d thr (%d warps/SM): %7.2f TFLOP/s (half the warps each)\n", grid1, threads, threads / 32, flop / ms * 1e-9);
    }
    // the evaluation kernel's blocking: (row halves per warp) x (live column tiles)
    for (int threads : {256, 512}) {
        auto run = [&](auto kern, int RH, int NL, const char* what) {
            float ms = timed([&] { kern<<<sms, threads>>>(out, iters / 4, 1.0000001, 0.9999999); });
            const double flop = (double)sms * (threads / 32) * (iters / 4) * RH * NL * 512.0;
            printf("DMMA blocked %dx%d %-22s %d warps/SM: %7.2f TFLOP/s\n", RH, NL, what, threads / 32, flop / ms * 1e-9);
        };
        run(k_dmma_blk<4, 4, 0>, 4, 4, "registers");
        run(k_dmma_blk<4, 4, 1>, 4, 4, "LDS operands");
        run(k_dmma_blk<4, 4, 2>, 4, 4, "LDS operands + DADD");
        run(k_dmma_blk<4, 1, 1>, 4, 1, "LDS operands");
        run(k_dmma_blk<4, 1, 2>, 4, 1, "LDS operands + DADD");
        run(k_dmma_blk<2, 4, 1>, 2, 4, "LDS operands");
        run(k_dmma_blk<2, 4, 2>, 2, 4, "LDS operands + DADD");
        run(k_dmma_blk<2, 1, 1>, 2, 1, "LDS operands");
        run(k_dmma_blk<2, 1, 2>, 2, 1, "LDS operands + DADD");
    }
    return 0;
}
