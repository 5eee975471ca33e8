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

// ---- HBM write ceilings: the stiffness kernels are bound by writing the element matrices, so the relevant roof is the
// device's WRITE bandwidth through the mechanism the kernels use.
// kind 0: every warp owns `slots` 4,608-byte staging blocks in shared memory and ships them with cp.async.bulk
//         (shared -> global), waiting only for the shared-memory read of its previous copies (the Hex8 kernel's phase 3)
// kind 1: plain 16-byte global stores from registers, grid-stride, fully coalesced
__global__ void bulk_write_kernel(double *out, unsigned long long nchunk, int chunk_bytes, int slot_stride_bytes, int slots, int mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < nwarp * slots * slot_stride_bytes / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = 1.0;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int slot = lane / (32 / slots);
    const bool issuer = lane % (32 / slots) == 0;
    const unsigned src = (unsigned)__cvta_generic_to_shared(sm + (size_t)(warp * slots + slot) * slot_stride_bytes);
    // mode bit 0: evict-first L2 policy on the copies; bit 1: blocked distribution (every warp slot writes one contiguous slab)
    unsigned long long pol;
    if (mode & 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    const unsigned long long nstream = (unsigned long long)gridDim.x * nwarp * slots;
    const unsigned long long me = ((unsigned long long)blockIdx.x * nwarp + warp) * slots + slot;
    const unsigned long long per = (nchunk + nstream - 1) / nstream;
    for (unsigned long long k = 0; k < per; k++) {
        const unsigned long long c = (mode & 2) ? me * per + k : me + k * nstream;
        if (issuer && c < nchunk) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(reinterpret_cast<unsigned char *>(out) + c * chunk_bytes),
                         "r"(src), "r"(chunk_bytes), "l"(pol)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256) stg_write_kernel(double2 *out, unsigned long long n16, int mode) {
    const double2 v = make_double2(1.0, 2.0);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * blockDim.x) {
        if (mode == 1) __stcs(out + i, v);       // st.global.cs (streaming, evict-first)
        else if (mode == 2) __stwt(out + i, v);  // st.global.wt (write-through)
        else out[i] = v;
    }
}

} // namespace

int bench_hbm_write(int kind, double *dev, unsigned long long nbytes, int warps_per_sm, void *stream_, double *gbs) {
    const int mode = kind >> 4; // upper bits: variant of the mechanism
    kind &= 15;
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev_id = 0, nsm = 148;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev_id);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int launches = 0;
    const int chunk = 4608, slot_stride = 4672, slots = 4;
    if (warps_per_sm < 1) warps_per_sm = 8;
    unsigned long long written = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, stream);
        if (kind == 0) {
            // CTAs of 4 warps; warps_per_sm / 4 CTAs per SM
            const int block = 128, ctas = (warps_per_sm + 3) / 4;
            const size_t smem = (size_t)4 * slots * slot_stride;
            cudaFuncSetAttribute(bulk_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const unsigned long long nchunk = nbytes / chunk;
            bulk_write_kernel<<<nsm * ctas, block, smem, stream>>>(dev, nchunk, chunk, slot_stride, slots, mode);
            written = nchunk * chunk;
        } else if (kind == 1) {
            stg_write_kernel<<<nsm * 8, 256, 0, stream>>>(reinterpret_cast<double2 *>(dev), nbytes / 16, mode);
            written = nbytes / 16 * 16;
        } else {
            cudaMemsetAsync(dev, 0, nbytes, stream);
            written = nbytes;
        }
        cudaEventRecord(e1, stream);
        launches++;
        if (cudaEventSynchronize(e1) != cudaSuccess) return -1;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *gbs = (double)written / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return launches;
}

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
