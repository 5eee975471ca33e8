// Standalone micro-benchmark: DFMA throughput vs. independent chains per warp and warps per SM (B200, sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i];
    if (s == 1.2345) out[0] = s;
}

template <int CHAINS>
void run(int warps_per_sm, double *d) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int iters = 20000 / CHAINS * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<CHAINS><<<nsm, warps_per_sm * 32>>>(d, iters, 0.9999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double dfma_warp_instrs_per_smsp = (double)iters * 4 * CHAINS * warps_per_sm / 4.0;
    double cycles = best * 1e-3 * 1.965e9;
    printf("chains %2d warps/SM %2d: %.3f ms  cycles per DFMA per SMSP = %.2f\n", CHAINS, warps_per_sm, best, cycles / dfma_warp_instrs_per_smsp);
}

int main() {
    double *d;
    cudaMalloc(&d, 8);
    for (int w : {4, 8, 16}) {
        run<1>(w, d);
        run<2>(w, d);
        run<4>(w, d);
        run<8>(w, d);
        run<16>(w, d);
        run<32>(w, d);
    }
    return 0;
}
