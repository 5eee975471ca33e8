// gl_bench.cu -- FP64 micro-benchmarks that measure the roofline denominators on the device the library runs on.
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the FP64 FMA and FP64 tensor-core (DMMA m8n8k4) peaks that
// bound the Hex8/Hex20/Tet10 integration kernels are measured here, with CUDA events, and reported by bench.py.
#include <cuda_runtime.h>

#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int CHAINS = 8; // independent accumulators per thread (covers the DFMA pipeline latency)

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i];
    if (s == 123.456) out[0] = s; // never true; keeps the chain alive
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
    double c[CHAINS][2];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        c[i][0] = (double)threadIdx.x;
        c[i][1] = (double)i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// DFMA and DMMA interleaved in one instruction stream: if the two share one pipe the combined rate stays at the single
// peak; if they are separate pipes it approaches the sum.
__global__ void __launch_bounds__(256) mixed_kernel(double *out, int iters, double a, double b) {
    double acc[CHAINS];
    double c[CHAINS][2];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        acc[i] = (double)(threadIdx.x + i);
        c[i][0] = (double)threadIdx.x;
        c[i][1] = (double)i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
#pragma unroll
            for (int r = 0; r < 8; r++) acc[i] = fma(acc[i], a, b); // 8 DFMA per DMMA: equal flops on both paths
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i] + c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

} // namespace

int bench_fp64(int kind, int iters, void *stream_, double *tflops) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    double *d_out = nullptr;
    if (cudaMalloc(&d_out, sizeof(double)) != cudaSuccess) return -1;
    const int grid = nsm * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int launches = 0;
    for (int rep = 0; rep < 2; rep++) { // first pass warms up
        cudaEventRecord(e0, stream);
        if (kind == 0) dfma_kernel<<<grid, block, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
        else if (kind == 1) dmma_kernel<<<grid, block, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
        else mixed_kernel<<<grid, block, 0, stream>>>(d_out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, stream);
        launches++;
        if (cudaEventSynchronize(e1) != cudaSuccess) return -1;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double threads = (double)grid * block;
    double flops;
    if (kind == 0) flops = threads * (double)iters * CHAINS * 2.0;
    else if (kind == 1) flops = (threads / 32.0) * (double)iters * CHAINS * (2.0 * 8 * 8 * 4);
    else flops = (threads / 32.0) * (double)iters * CHAINS * (2.0 * 8 * 8 * 4) + threads * (double)iters * CHAINS * 8 * 2.0;
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return launches;
}

} // namespace gl
