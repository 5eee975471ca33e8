// gl_kernels_vec_tpc.cu -- thread-per-cell kernel for the vector-dof element vectors:
//   integ::vec_02_nv   b(m,i) = sum_p N(m) v(i) c_p                    (src/integ/vec_02_nv.rs:85-140)   body forces
//   integ::vec_04_bt   d(m,i) = sum_p sum_a sigma(i,a) B(m,a) c_p [+ hoop] (src/integ/vec_04_bt.rs:93-170)  internal forces
// with c_p = det(J_p) w_p alpha [r_p if axisymmetric], tight blocks, clear = true, coefficients UNIFORM, PER_MARKER, PER_CELL
// or PER_GAUSS (the stress of a Newton step arrives per Gauss point).  These are the residual side of the same loop that
// consumes the stiffness; their blocks are tiny (6 .. 30 doubles), so -- as in gl_kernels_small2d.cu -- one thread owns one
// cell end to end and the warp's blocks are transposed through shared memory for coalesced stores.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int VT_THREADS = 128;

template <int NN, int SD, bool AXI, bool BT>
__global__ void __launch_bounds__(VT_THREADS) vec_tpc_kernel(const IntegParams p) {
    constexpr int NDOF = NN * SD;
    constexpr int TS = 33;
    constexpr int NS = 2 * SD;
    constexpr int CL = BT ? NS : SD; // doubles per coefficient record

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    double *s_w = smem;                  // [ng]
    double *s_N = s_w + ng;              // [ng][NN]
    double *s_L = s_N + ng * NN;         // [ng][NN][SD]
    double *s_T = s_L + ng * NN * SD;    // [warps][NDOF][TS]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < ng * (1 + NN + NN * SD); i += VT_THREADS) s_w[i] = p.rule[i];
    __syncthreads();
    double *Tw = s_T + (size_t)wid * NDOF * TS;
    const double hs = 0.70710678118654752440084436210484903928;

    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + VT_THREADS - 1) / VT_THREADS;
    // the point ids of the NEXT batch are fetched while this one is integrated (gl_kernels_small2d.cu): one hop of the dependent gather
    // off the critical path
    unsigned int pid_next[NN];
    auto load_ids = [&](unsigned long long c) {
#pragma unroll
        for (int m = 0; m < NN; m++) pid_next[m] = c < p.hi ? p.conn[(unsigned long long)m * p.ncell_k + c] : 0u;
    };
    load_ids(p.lo + (unsigned long long)blockIdx.x * VT_THREADS + tid);
    for (unsigned long long batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const unsigned long long wcell0 = p.lo + batch * VT_THREADS + (unsigned long long)wid * 32; // first cell of this warp
        const unsigned long long cell = wcell0 + lane;
        const bool active = cell < p.hi;
        const unsigned long long orig = active ? (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) : 0ull;

        double x[NN][SD];
#pragma unroll
        for (int m = 0; m < NN; m++) {
#pragma unroll
            for (int j = 0; j < SD; j++) x[m][j] = 0.0;
            if (active) {
                const unsigned int pid = pid_next[m];
                if constexpr (SD == 2) {
                    const double2 a = reinterpret_cast<const double2 *>(p.coords_pm)[pid];
                    x[m][0] = a.x;
                    x[m][1] = a.y;
                } else {
                    const double2 a = reinterpret_cast<const double2 *>(p.coords_pm)[2ull * pid];
                    const double2 c = reinterpret_cast<const double2 *>(p.coords_pm)[2ull * pid + 1];
                    x[m][0] = a.x;
                    x[m][1] = a.y;
                    x[m][2] = c.x;
                }
            }
        }
        load_ids(cell + (unsigned long long)gridDim.x * VT_THREADS);

        double acc[NN][SD];
#pragma unroll
        for (int m = 0; m < NN; m++)
#pragma unroll
            for (int i = 0; i < SD; i++) acc[m][i] = 0.0;

        if (active) {
            // coefficient record(s) of this cell: the uniform record, or record `rec` (+ one per Gauss point)
            const double *coef0 = nullptr;
            if (p.coef != nullptr) {
                const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell] : (orig - p.coef_begin);
                coef0 = p.coef + rec * p.coef_cell_stride;
            }
            bool bad = false;
#pragma unroll 1
            for (int gp = 0; gp < ng; gp++) {
                const double *L = s_L + gp * NN * SD;
                const double *N = s_N + gp * NN;
                double J[SD][SD];
#pragma unroll
                for (int i = 0; i < SD; i++)
#pragma unroll
                    for (int j = 0; j < SD; j++) J[i][j] = 0.0;
#pragma unroll
                for (int m = 0; m < NN; m++) geom::jacobian_add_node<SD, SD>(J, x[m], L + m * SD);
                double Ji[SD][SD];
                const double det = geom::invert(J, Ji);
                bad = bad || fabs(det) <= ZERO_DETERMINANT;
                double c = det * s_w[gp] * p.alpha;
                double r = 1.0;
                if constexpr (AXI) {
                    r = 0.0;
#pragma unroll
                    for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], x[m][0])); // radius_at_ip
                }
                double t[CL];
#pragma unroll
                for (int a = 0; a < CL; a++) t[a] = coef0 ? coef0[(unsigned long long)gp * p.coef_gp_stride + a] : p.coef_uniform[a];
                if constexpr (!BT) {
                    // vec_02_nv: b(m,i) += c [r] N(m) v(i)
                    const double cr = AXI ? c * r : c;
#pragma unroll
                    for (int m = 0; m < NN; m++)
#pragma unroll
                        for (int i = 0; i < SD; i++) acc[m][i] = fma(cr * N[m], t[i], acc[m][i]);
                } else {
                    // vec_04_bt: d(m,i) += c [r] sum_a sigma(i,a) B(m,a)  (+ c N(m) sigma_22 on the radial dof, axisymmetric)
                    double sg[SD][SD];
#pragma unroll
                    for (int i = 0; i < SD; i++)
#pragma unroll
                        for (int a = 0; a < SD; a++) {
                            if (i == a) sg[i][a] = t[i];
                            else if (SD == 2) sg[i][a] = t[3] * hs;
                            else sg[i][a] = (i + a == 1 ? t[3] : (i + a == 3 ? t[4] : t[5])) * hs;
                        }
                    const double cr = AXI ? c * r : c;
#pragma unroll
                    for (int m = 0; m < NN; m++) {
                        double bm[SD];
                        geom::gradient_row<SD>(L + m * SD, Ji, bm);
#pragma unroll
                        for (int i = 0; i < SD; i++) {
                            double sum = sg[i][0] * bm[0];
#pragma unroll
                            for (int a = 1; a < SD; a++) sum = fma(sg[i][a], bm[a], sum);
                            acc[m][i] = fma(cr, sum, acc[m][i]);
                        }
                        if constexpr (AXI) acc[m][0] = fma(c * N[m], t[2], acc[m][0]);
                    }
                }
            }
            if (bad) atomicMin(p.err_cell, orig); // calc_jacobian inverts J for solid cells too (scratchpad_calc_jacobian.rs:115-118)
        }

        __syncwarp(); // the previous batch's copy-out has finished reading Tw
#pragma unroll
        for (int m = 0; m < NN; m++)
#pragma unroll
            for (int i = 0; i < SD; i++) Tw[(SD * m + i) * TS + lane] = acc[m][i];
        __syncwarp();
        {
            unsigned long long vo = 0;
            if (active) vo = p.out_off ? (p.out_off[cell] - p.out_base) : (orig - p.cell_begin) * NDOF;
            const int nc = wcell0 < p.hi ? (int)(p.hi - wcell0 < 32 ? p.hi - wcell0 : 32) : 0;
            for (int w0 = 0; w0 < nc * NDOF; w0 += 32) {
                const int w = w0 + lane, c = w / NDOF, q = w - c * NDOF;
                const unsigned long long base = __shfl_sync(0xffffffffu, vo, c & 31);
                if (w < nc * NDOF) p.out[base + q] = Tw[q * TS + c];
            }
        }
    }
}

template <int NN, int SD>
int launch_vt_kind(const IntegParams &p, void *stream) {
    const size_t smem = sizeof(double) * ((size_t)p.ng * (1 + NN + NN * SD) + (size_t)(VT_THREADS / 32) * NN * SD * 33);
    if (smem > 200 * 1024) return 0;
    const bool bt = p.op == GL_OP_VEC_04_BT;
    typedef void (*kern_t)(const IntegParams);
    kern_t kern;
    if constexpr (SD == 2) {
        if (p.axisym) kern = bt ? vec_tpc_kernel<NN, SD, true, true> : vec_tpc_kernel<NN, SD, true, false>;
        else kern = bt ? vec_tpc_kernel<NN, SD, false, true> : vec_tpc_kernel<NN, SD, false, false>;
    } else {
        kern = bt ? vec_tpc_kernel<NN, SD, false, true> : vec_tpc_kernel<NN, SD, false, false>;
    }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((226 * 1024) / (smem + 1024));
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long nbatch = (p.hi - p.lo + VT_THREADS - 1) / VT_THREADS;
    // 8 CTAs per resident slot on large meshes, fewer on small ones so that every CTA still loops over ~8 batches (gl_kernels_small2d.cu)
    const unsigned long long slots = (unsigned long long)nsm * ctas_per_sm;
    unsigned long long mult = nbatch / (slots * 8);
    mult = mult < 1 ? 1 : (mult > 8 ? 8 : mult);
    unsigned long long grid = slots * mult;
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    kern<<<(unsigned)grid, VT_THREADS, smem, (cudaStream_t)stream>>>(p);
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("vec_tpc<nn=%d,sd=%d,op=%d>", NN, SD, p.op);
    return 1;
}

} // namespace

// vec_02_nv / vec_04_bt on solid kinds, tight blocks, overwrite; needs the point-major coordinates
int launch_vec_tpc(const IntegParams &p, int sd, void *stream) {
    if (p.op != GL_OP_VEC_02_NV && p.op != GL_OP_VEC_04_BT) return 0;
    if (!p.clear || p.coords_pm == nullptr || p.ii0 != 0) return 0;
    const unsigned ndof = (unsigned)p.nnode * (unsigned)sd;
    if (p.nrow != ndof) return 0;
    const unsigned long long cl = p.op == GL_OP_VEC_04_BT ? 2ull * sd : (unsigned long long)sd;
    if (p.coef != nullptr && p.coef_gp_stride != 0 && p.coef_gp_stride != cl) return 0;
    if (tuning_env("GL_DISABLE_VEC_TPC")) return 0; // testing knob: fall back to the generic kernel
    if (sd == 2) {
        switch (p.nnode) {
        case 3: return launch_vt_kind<3, 2>(p, stream);
        case 4: return launch_vt_kind<4, 2>(p, stream);
        case 6: return launch_vt_kind<6, 2>(p, stream);
        case 8: return launch_vt_kind<8, 2>(p, stream);
        case 9: return launch_vt_kind<9, 2>(p, stream);
        default: return 0;
        }
    }
    if (p.axisym) return 0;
    switch (p.nnode) {
    case 4: return launch_vt_kind<4, 3>(p, stream);
    case 8: return launch_vt_kind<8, 3>(p, stream);
    case 10: return launch_vt_kind<10, 3>(p, stream);
    default: return 0;
    }
}

} // namespace gl
