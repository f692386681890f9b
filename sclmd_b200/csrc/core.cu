// Error slot, device selection and library identification for libsclmd_b200.
#include "common.cuh"

namespace sclmd {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); libsclmd_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SCLMD_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (%d visible)", device, n);
        return SCLMD_ERR_ARG;
    }
    SCLMD_CUDA(cudaSetDevice(device));
    return SCLMD_OK;
}

int sm_count(int device) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
    return n;
}

}  // namespace sclmd

extern "C" {

const char *sclmd_last_error(void) { return sclmd::g_err; }
int sclmd_version(void) { return 100; }

int sclmd_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        sclmd::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return SCLMD_ERR_CUDA;
    }
    return n;
}

int sclmd_device_info(int device, int *sms, int *cc_major, int *cc_minor, uint64_t *mem_bytes) {
    if (int e = sclmd::select_device(device)) return e;
    cudaDeviceProp p;
    SCLMD_CUDA(cudaGetDeviceProperties(&p, device));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (mem_bytes) *mem_bytes = (uint64_t)p.totalGlobalMem;
    return SCLMD_OK;
}
}

// ---- FP64 peak probes (roofline denominators; MEASURED_PEAKS.json has no FP64 figure) ----
namespace {
__global__ void __launch_bounds__(256) k_dfma_probe(double *out, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dmma_probe(double *out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
}  // namespace

extern "C" int sclmd_probe_fp64(int device, int kind, double *tflops) {
    if (int e = sclmd::select_device(device)) return e;
    SCLMD_REQUIRE(tflops && (kind == 0 || kind == 1), "sclmd_probe_fp64: kind must be 0 (DFMA) or 1 (DMMA)");
    double *d = nullptr;
    SCLMD_CUDA(cudaMalloc(&d, 8));
    const int iters = 20000, blocks = sclmd::sm_count(device) * 8;
    cudaEvent_t e0, e1;
    SCLMD_CUDA(cudaEventCreate(&e0));
    SCLMD_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        SCLMD_CUDA(cudaEventRecord(e0));
        if (kind == 0) k_dfma_probe<<<blocks, 256>>>(d, iters);
        else k_dmma_probe<<<blocks, 256>>>(d, iters);
        SCLMD_CUDA(cudaEventRecord(e1));
        SCLMD_CUDA(cudaEventSynchronize(e1));
        float ms;
        SCLMD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    // DFMA: 16 FMAs/thread/iter; DMMA: 16 mma/warp/iter, 8*8*4 FMAs each
    const double flops = kind == 0 ? 2.0 * 16 * iters * 256.0 * blocks : 2.0 * 256 * 16 * iters * 8.0 * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return SCLMD_OK;
}
