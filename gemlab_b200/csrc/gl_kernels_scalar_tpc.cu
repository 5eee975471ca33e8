// gl_kernels_scalar_tpc.cu -- thread-per-cell kernel for the gradient-based scalar-dof cases from one geometry pass:
//   integ::mat_03_btb  K(m,n) = sum_p B(n).T.B(m) c_p     (src/integ/mat_03_btb.rs:50-131)   e.g. heat conduction
//   integ::vec_03_bv   c(m)   = sum_p (w.B(m)) c_p         (src/integ/vec_03_bv.rs:84-141)    e.g. flux vector
// with c_p = det(J_p) w_p alpha [r_p if axisymmetric], uniform coefficients, tight blocks, clear = true  (the conductivity
// and flux half of BASELINE.json config 5).  Small blocks (9 .. 81 doubles): as in gl_kernels_small2d.cu ONE THREAD OWNS
// ONE CELL end to end -- gather from the point-major coordinates, J / inverse / B per Gauss point in the oracle's operation
// order (gl_geom.cuh), the symmetric matrix accumulated as its upper triangle in registers -- with no barrier, and the warp's
// 32 blocks are transposed through shared memory (stride 33) so that global memory sees coalesced stores.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int TPC_THREADS = 128;

template <int NN, int SD, bool AXI>
__global__ void __launch_bounds__(TPC_THREADS) scalar_tpc_kernel(const IntegParams p, const ScalarFused f) {
    constexpr int KK = NN * NN;
    constexpr int NPAIR = NN * (NN + 1) / 2;
    constexpr int TS = 33;
    const bool m03 = f.mask & 2, v03 = f.mask & 8;

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    double *s_w = smem;                  // [ng]
    double *s_N = s_w + ng;              // [ng][NN]
    double *s_L = s_N + ng * NN;         // [ng][NN][SD]
    double *s_T = s_L + ng * NN * SD;    // [warps][KK + NN][TS]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < ng * (1 + NN + NN * SD); i += TPC_THREADS) s_w[i] = p.rule[i];
    __syncthreads();
    double *Tw = s_T + (size_t)wid * (KK + NN) * TS;

    // symmetric 2nd-order tensor T from its Mandel vector (Tensor2::vector(), src/integ/mat_03_btb.rs:101-128)
    double tt[SD][SD];
    {
        const double hs = 0.70710678118654752440084436210484903928;
#pragma unroll
        for (int a = 0; a < SD; a++)
#pragma unroll
            for (int b = 0; b < SD; b++) {
                if (a == b) tt[a][b] = f.t[a];
                else if (SD == 2) tt[a][b] = f.t[3] * hs;
                else tt[a][b] = (a + b == 1 ? f.t[3] : (a + b == 3 ? f.t[4] : f.t[5])) * hs;
            }
    }

    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + TPC_THREADS - 1) / TPC_THREADS;
    // software-pipelined gather (Mesh::set_pad): the coordinates of the next batch and the point ids of the batch after it are in
    // flight while this one is integrated (see gl_kernels_small2d.cu)
    constexpr bool PFX = NN * SD <= 16; // the larger kinds have no registers left for a second set of coordinates: ids only
    unsigned int pid_next[NN];
    double xn[NN][SD];
    auto load_ids = [&](unsigned long long c) {
#pragma unroll
        for (int m = 0; m < NN; m++) pid_next[m] = c < p.hi ? p.conn[(unsigned long long)m * p.ncell_k + c] : 0u;
    };
    auto load_coords = [&](unsigned long long c) {
#pragma unroll
        for (int m = 0; m < NN; m++) {
#pragma unroll
            for (int j = 0; j < SD; j++) xn[m][j] = 0.0;
            if (c < p.hi) {
                if constexpr (SD == 2) {
                    const double2 a = reinterpret_cast<const double2 *>(p.coords_pm)[pid_next[m]];
                    xn[m][0] = a.x;
                    xn[m][1] = a.y;
                } else {
                    const double2 a = reinterpret_cast<const double2 *>(p.coords_pm)[2ull * pid_next[m]];
                    const double2 c2 = reinterpret_cast<const double2 *>(p.coords_pm)[2ull * pid_next[m] + 1];
                    xn[m][0] = a.x;
                    xn[m][1] = a.y;
                    xn[m][2] = c2.x;
                }
            }
        }
    };
    const unsigned long long cstep = (unsigned long long)gridDim.x * TPC_THREADS;
    {
        const unsigned long long c0 = p.lo + (unsigned long long)blockIdx.x * TPC_THREADS + tid;
        load_ids(c0);
        if constexpr (PFX) {
            load_coords(c0);
            load_ids(c0 + cstep);
        }
    }
    for (unsigned long long batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const unsigned long long wcell0 = p.lo + batch * TPC_THREADS + (unsigned long long)wid * 32; // first cell of this warp
        const unsigned long long cell = wcell0 + lane;
        const bool active = cell < p.hi;

        if constexpr (!PFX) load_coords(cell); // with the ids fetched during the previous batch
        double x[NN][SD];
#pragma unroll
        for (int m = 0; m < NN; m++)
#pragma unroll
            for (int j = 0; j < SD; j++) x[m][j] = xn[m][j];
        if constexpr (PFX) {
            load_coords(cell + cstep);  // next batch (its point ids arrived during the previous iteration)
            load_ids(cell + 2 * cstep); // the batch after
        } else {
            load_ids(cell + cstep);
        }

        double acc[NPAIR], av[NN];
#pragma unroll
        for (int q = 0; q < NPAIR; q++) acc[q] = 0.0;
#pragma unroll
        for (int m = 0; m < NN; m++) av[m] = 0.0;
        bool bad = false;

        if (active) {
#pragma unroll 1
            for (int gp = 0; gp < ng; gp++) {
                const double *L = s_L + gp * NN * SD;
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
                double v[NN][SD];
#pragma unroll
                for (int m = 0; m < NN; m++) geom::gradient_row<SD>(L + m * SD, Ji, v[m]);
                double c = det * s_w[gp] * p.alpha;
                if constexpr (AXI) {
                    const double *N = s_N + gp * NN;
                    double r = 0.0;
#pragma unroll
                    for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], x[m][0])); // radius_at_ip
                    c *= r;
                }
                if (v03) {
#pragma unroll
                    for (int m = 0; m < NN; m++) {
                        double dot = f.w[0] * v[m][0];
#pragma unroll
                        for (int k = 1; k < SD; k++) dot = fma(f.w[k], v[m][k], dot);
                        av[m] = fma(c, dot, av[m]);
                    }
                }
                if (m03) {
                    int pair = 0;
#pragma unroll
                    for (int n = 0; n < NN; n++) {
                        double tn[SD]; // c T.B(n)
#pragma unroll
                        for (int a = 0; a < SD; a++) {
                            double sum = tt[a][0] * v[n][0];
#pragma unroll
                            for (int b = 1; b < SD; b++) sum = fma(tt[a][b], v[n][b], sum);
                            tn[a] = sum * c;
                        }
#pragma unroll
                        for (int m = 0; m <= n; m++) {
                            double k = acc[pair];
#pragma unroll
                            for (int a = 0; a < SD; a++) k = fma(v[m][a], tn[a], k);
                            acc[pair] = k;
                            pair++;
                        }
                    }
                }
            }
            if (bad) atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
        }

        // ---- transposition buffer: entry q of this lane's block at Tw[q*TS + lane] -----------------------------------------
        __syncwarp(); // the previous batch's copy-out has finished reading Tw
        if (m03) {
            int pair = 0;
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int m = 0; m <= n; m++) {
                    Tw[(m + NN * n) * TS + lane] = acc[pair];
                    if (m != n) Tw[(n + NN * m) * TS + lane] = acc[pair]; // T is symmetric, so is K
                    pair++;
                }
        }
        if (v03) {
#pragma unroll
            for (int m = 0; m < NN; m++) Tw[(KK + m) * TS + lane] = av[m];
        }
        __syncwarp();

        // ---- coalesced copy-out: a flat sweep over the warp's entries; every lane knows the destination of its own cell and
        // hands it out by shuffle, so runs of cells that are consecutive in the output (always, on single-kind meshes; row by
        // row on mixed ones) leave as full lines
        {
            unsigned long long mo = 0, vo = 0;
            if (active) {
                const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
                mo = f.moff ? (f.moff[cell] - f.mbase) : (orig - p.cell_begin) * KK;
                vo = f.voff ? (f.voff[cell] - f.vbase) : (orig - p.cell_begin) * NN;
            }
            const int nc = wcell0 < p.hi ? (int)(p.hi - wcell0 < 32 ? p.hi - wcell0 : 32) : 0;
            // one contiguous run per output: always on single-kind meshes; on mixed ones whenever the warp's cells lie inside one stripe
            const unsigned long long mo0 = __shfl_sync(0xffffffffu, mo, 0), vo0 = __shfl_sync(0xffffffffu, vo, 0);
            const bool single = p.cell_ids == nullptr && f.moff == nullptr && f.voff == nullptr;
            const bool one_run = single || __all_sync(0xffffffffu, !active || (mo == mo0 + (unsigned long long)lane * KK && vo == vo0 + (unsigned long long)lane * NN));
            if (one_run) {
                if (m03 && nc > 0) {
                    double *dst = f.out_m03 + mo0;
                    for (int w = lane; w < nc * KK; w += 32) dst[w] = Tw[(w % KK) * TS + w / KK];
                }
                if (v03 && nc > 0) {
                    double *dst = f.out_v03 + vo0;
                    for (int w = lane; w < nc * NN; w += 32) dst[w] = Tw[(KK + w % NN) * TS + w / NN];
                }
                continue;
            }
            if (m03) {
                for (int w0 = 0; w0 < nc * KK; w0 += 32) {
                    const int w = w0 + lane, c = w / KK, q = w - c * KK;
                    const unsigned long long base = __shfl_sync(0xffffffffu, mo, c & 31);
                    if (w < nc * KK) f.out_m03[base + q] = Tw[q * TS + c];
                }
            }
            if (v03) {
                for (int w0 = 0; w0 < nc * NN; w0 += 32) {
                    const int w = w0 + lane, c = w / NN, q = w - c * NN;
                    const unsigned long long base = __shfl_sync(0xffffffffu, vo, c & 31);
                    if (w < nc * NN) f.out_v03[base + q] = Tw[(KK + q) * TS + c];
                }
            }
        }
    }
}

template <int NN, int SD>
int launch_tpc_kind(const IntegParams &p, const ScalarFused &f, void *stream) {
    const size_t smem = sizeof(double) * ((size_t)p.ng * (1 + NN + NN * SD) + (size_t)(TPC_THREADS / 32) * (NN * NN + NN) * 33);
    if (smem > 200 * 1024) return 0;
    typedef void (*kern_t)(const IntegParams, const ScalarFused);
    kern_t kern;
    if constexpr (SD == 2) kern = p.axisym ? scalar_tpc_kernel<NN, SD, true> : scalar_tpc_kernel<NN, SD, false>;
    else kern = scalar_tpc_kernel<NN, SD, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((226 * 1024) / (smem + 1024));
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long nbatch = (p.hi - p.lo + TPC_THREADS - 1) / TPC_THREADS;
    // 8 CTAs per resident slot on large meshes, fewer on small ones so that every CTA still loops over ~8 batches (gl_kernels_small2d.cu)
    const unsigned long long slots = (unsigned long long)nsm * ctas_per_sm;
    unsigned long long mult = nbatch / (slots * 8);
    mult = mult < 1 ? 1 : (mult > 8 ? 8 : mult);
    unsigned long long grid = slots * mult;
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    kern<<<(unsigned)grid, TPC_THREADS, smem, (cudaStream_t)stream>>>(p, f);
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("scalar_tpc<nn=%d,sd=%d,mask=%d>", NN, SD, f.mask);
    return 1;
}

} // namespace

// masks made of mat_03_btb (2) and vec_03_bv (8) only; needs the point-major coordinates
int launch_scalar_tpc(const IntegParams &p, const ScalarFused &f, int sd, void *stream) {
    if ((f.mask & ~10) != 0 || f.mask == 0 || p.coords_pm == nullptr) return 0;
    if (tuning_env("GL_DISABLE_TPC")) return 0; // testing knob: fall back to gl_kernels_scalar.cu
    if (sd == 2) {
        switch (p.nnode) {
        case 3: return launch_tpc_kind<3, 2>(p, f, stream);
        case 4: return launch_tpc_kind<4, 2>(p, f, stream);
        case 6: return launch_tpc_kind<6, 2>(p, f, stream);
        case 8: return launch_tpc_kind<8, 2>(p, f, stream);
        case 9: return launch_tpc_kind<9, 2>(p, f, stream);
        default: return 0;
        }
    }
    if (p.axisym) return 0;
    switch (p.nnode) {
    case 4: return launch_tpc_kind<4, 3>(p, f, stream);
    case 8: return launch_tpc_kind<8, 3>(p, f, stream);
    default: return 0;
    }
}

} // namespace gl
