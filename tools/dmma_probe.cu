// DMMA.8x8x4 issue-rate probes: which operand pattern sustains the FP64 tensor peak?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/dmma_probe.cu && /tmp/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// FM x FN accumulators, distinct A (FM) and B (FN) registers refreshed from shared memory every k4 step (LD = 1) or held (LD = 0)
template <int FM, int FN, int LD>
__global__ void __launch_bounds__(256) k_probe(double *out, int iters) {
    __shared__ double sa[64 * 20], sb[64 * 20];
    for (int i = threadIdx.x; i < 64 * 20; i += 256) { sa[i] = 1.0 + i * 1e-6; sb[i] = 1e-3 + i * 1e-9; }
    __syncthreads();
    const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    double acc[FM][FN][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[FM], b[FN];
#pragma unroll
    for (int i = 0; i < FM; ++i) a[i] = sa[(i * 8 + fr) * 20 + fk];
#pragma unroll
    for (int j = 0; j < FN; ++j) b[j] = sb[(j * 8 + fr) * 20 + fk];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
            if (LD) {
#pragma unroll
                for (int i = 0; i < FM; ++i) a[i] = sa[((i * 8 + fr) % 64) * 20 + kk + fk];
#pragma unroll
                for (int j = 0; j < FN; ++j) b[j] = sb[((j * 8 + fr) % 64) * 20 + kk + fk];
            }
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < FN; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}
template <int FM, int FN, int LD>
void run(const char *name, int ctas_per_sm, int iters) {
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    const int blocks = 148 * ctas_per_sm;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_probe<FM, FN, LD><<<blocks, 256>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double flops = 2.0 * 256 * (double)FM * FN * 4 * iters * 8.0 * blocks;
    printf("%-28s FMxFN=%dx%d ld=%d ctas/SM=%d : %.2f TFLOP/s\n", name, FM, FN, LD, ctas_per_sm, flops / (best * 1e-3) / 1e12);
    cudaFree(d);
}
// do DFMA and DMMA share the FP64 datapath?  even warps run a DFMA chain, odd warps a DMMA chain
__global__ void __launch_bounds__(256) k_mixed(double *out, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool use_dmma = mode == 1 || (mode == 2 && (warp & 1));
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, b = 1e-9;
    if (use_dmma) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(acc[i][0], acc[i][1], a, b);
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc[i][0] = fma(acc[i][0], a, b); acc[i][1] = fma(acc[i][1], a, b); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}
void run_mixed(int mode, const char *name) {
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 4;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_mixed<<<blocks, 256>>>(d, iters, mode);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // per warp per iteration: DFMA warps 32 lanes x 32 FMA; DMMA warps 16 x 256 FMA
    const double dfma = 2.0 * 32 * 32, dm = 2.0 * 16 * 256;
    const double per_iter_block = mode == 0 ? 8 * dfma : (mode == 1 ? 8 * dm : 4 * dfma + 4 * dm);
    printf("%-34s : %.2f TFLOP/s  (%.3f ms)\n", name, per_iter_block * iters * blocks / (best * 1e-3) / 1e12, best);
    cudaFree(d);
}
int main() {
    run_mixed(0, "all warps DFMA");
    run_mixed(1, "all warps DMMA");
    run_mixed(2, "half DFMA + half DMMA");
    run<8, 4, 0>("64x32 warp tile, regs held", 1, 4000);
    run<8, 4, 1>("64x32 warp tile, LDS per k4", 1, 4000);
    run<8, 4, 1>("64x32 warp tile, LDS per k4", 2, 4000);
    run<4, 4, 0>("32x32 warp tile, regs held", 2, 8000);
    run<4, 4, 1>("32x32 warp tile, LDS per k4", 2, 8000);
    run<4, 2, 1>("32x16 warp tile, LDS per k4", 2, 16000);
    run<4, 2, 1>("32x16 warp tile, LDS per k4", 4, 16000);
    run<8, 2, 1>("64x16 warp tile, LDS per k4", 2, 8000);
    run<2, 8, 1>("16x64 warp tile, LDS per k4", 2, 8000);
    return 0;
}
