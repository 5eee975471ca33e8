// dmma_ubench.cu -- how many warps / independent accumulators does the FP64 tensor pipe of one SM sub-partition need?
// Measures DMMA m8n8k4 issue rate per SM for W warps per sub-partition and C independent accumulators per warp, and the
// progress of a warp running a dependent DADD/DMUL chain (the Jacobian loop) next to warps that stream DMMAs.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/dmma_ubench tools/dmma_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void dmma_rate(double *out, int iters, double a, double b) {
    double c[C][2];
#pragma unroll
    for (int i = 0; i < C; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < C; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// warps with (warp id / 4) == 0 run the DADD/DMUL chain (5 independent chains, like the Jacobian loop); the others stream DMMAs.
// clock64 per role is written out: cycles per chain-instruction and per DMMA.
__global__ void mixed_roles(double *out, long long *clk, int iters, double a, double b, int chain_warps) {
    const int warp = threadIdx.x >> 5;
    const bool chain = warp < chain_warps;
    long long t0 = clock64();
    if (chain) {
        double j[5] = {1, 2, 3, 4, 5};
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 5; k++) j[k] = __dadd_rn(j[k], __dmul_rn(a, j[k]));
        }
        if (j[0] + j[1] + j[2] + j[3] + j[4] == 123.456) out[0] = j[0];
    } else {
        double c[18][2];
#pragma unroll
        for (int i = 0; i < 18; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
        for (int it = 0; it < iters / 4; it++) {
#pragma unroll
            for (int i = 0; i < 18; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 18; i++) s += c[i][0] + c[i][1];
        if (s == 123.456) out[0] = s;
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) clk[warp] = t1 - t0;
}

template <int C>
void run(int warps_per_smsp, double *d) {
    const int iters = 4096;
    const int threads = warps_per_smsp * 4 * 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dmma_rate<C><<<148, threads>>>(d, 64, 1.0000001, 0.5);
    cudaEventRecord(e0);
    dmma_rate<C><<<148, threads>>>(d, iters, 1.0000001, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)iters * C * warps_per_smsp * 4 * 148;
    const double tf = n * 512 / (ms * 1e-3) / 1e12;
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * C * warps_per_smsp);
    printf("DMMA: %d warps/SMSP x %2d accumulators: %6.2f TFLOP/s, %5.1f cycles per DMMA per SMSP\n", warps_per_smsp, C, tf, cyc);
}

int main() {
    double *d; cudaMalloc(&d, 1024);
    long long *clk; cudaMalloc(&clk, 64 * sizeof(long long));
    for (int w : {1, 2, 3, 4, 8}) { run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<18>(w, d); }
    // mixed roles: 16 warps per SM (4 per SMSP); chain_warps = 4 -> one chain warp and three DMMA warps per sub-partition
    for (int cw : {0, 4, 8, 16}) {
        const int iters = 8192;
        mixed_roles<<<148, 512>>>(d, clk, iters, 1.0000001, 0.5, cw);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        printf("mixed: %2d chain warps of 16:", cw);
        if (cw > 0) printf("  chain warp: %6.1f cycles per (DMUL+DADD) pair", (double)h[0] / (iters * 5.0));
        if (cw < 16) printf("  DMMA warp: %6.1f cycles per DMMA", (double)h[15] / (iters / 4 * 18.0));
        printf("\n");
    }
    return 0;
}
