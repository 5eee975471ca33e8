// gl_kernels_scalar.cu -- fused kernel for the scalar-dof integration cases over one geometry pass:
//   integ::mat_01_nsn  K(m,n) = sum_p N(m) s N(n) c_p            (src/integ/mat_01_nsn.rs:48-102)
//   integ::mat_03_btb  K(m,n) = sum_p B(n).T.B(m) c_p            (src/integ/mat_03_btb.rs:50-131)
//   integ::vec_01_ns   a(m)   = sum_p N(m) s c_p                  (src/integ/vec_01_ns.rs:83-130)
//   integ::vec_03_bv   c(m)   = sum_p (w.B(m)) c_p                (src/integ/vec_03_bv.rs:84-141)
// with c_p = det(J) w_p alpha [r_p if axisymmetric], uniform coefficients, tight blocks and clear = true.  Any subset of
// the four cases can be requested in one launch (gl_integ_fused): BASELINE.json config 4 wants mat_01_nsn + vec_01_ns of
// a Tet10 mesh from ONE Jacobian evaluation, config 5 mat_03_btb + vec_03_bv.
//
// Same skeleton as gl_kernels_bdb.cu: gather -> one thread per (cell, Gauss point) geometry -> one lane per COLUMN NODE
// accumulating the circulant half of its column in registers (both matrices are symmetric: T is a symmetric tensor) ->
// staging in shared memory -> coalesced copy-out.  N is never evaluated on the device: it is the per-(kind, rule) table.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

template <int NN, int SD, bool AXI>
__global__ void __launch_bounds__(256) scalar_kernel(const IntegParams p, const ScalarFused f, const int CPC, const int nthreads_used) {
    constexpr int ND = NN / 2 + 1;
    constexpr int XS = (NN * SD) | 1;
    constexpr int KK = NN * NN;
    const bool m01 = f.mask & 1, m03 = f.mask & 2, v01 = f.mask & 4, v03 = f.mask & 8;
    const bool need_b = m03 || v03;
    const int gs = need_b ? ((1 + NN * SD) | 1) : 1; // record per (cell, gauss point): c | B
    const int ng = p.ng;

    extern __shared__ __align__(16) double smem[];
    double *s_K1 = smem;                        // [CPC][KK] mat_01_nsn staging
    double *s_K3 = s_K1 + (m01 ? CPC * KK : 0); // [CPC][KK] mat_03_btb staging
    double *s_w = s_K3 + (m03 ? CPC * KK : 0);  // [ng]
    double *s_N = s_w + ng;                     // [ng][NN]
    double *s_L = s_N + ng * NN;                // [ng][NN][SD]
    double *s_X = s_L + ng * NN * SD;           // [CPC][XS]
    double *s_G = s_X + CPC * XS;               // [CPC][ng][gs]

    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < ng * (1 + NN + NN * SD); i += nt) s_w[i] = p.rule[i];
    const bool worker = tid < nthreads_used;
    const int cs = worker ? tid / NN : 0;
    const int n = worker ? tid % NN : 0;

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

    // software-pipelined gather (Mesh::set_pad): thread (node gm, cell slot gc) keeps the coordinates of the NEXT tile in
    // registers (loaded one tile ahead) and the point id of the tile after that (two tiles ahead), so neither the index
    // load nor the dependent coordinate loads are exposed
    const bool gather_lane = tid < CPC * NN;
    const int gm = tid / CPC, gc = tid - gm * CPC;
    const unsigned long long gstep = (unsigned long long)gridDim.x * CPC;
    double gx[SD];
    unsigned int gpid = 0;
#pragma unroll
    for (int j = 0; j < SD; j++) gx[j] = 0.0;
    if (gather_lane) {
        const unsigned long long cell = p.lo + (unsigned long long)blockIdx.x * CPC + gc;
        if (cell < p.hi) {
            const unsigned int pid = p.conn[(unsigned long long)gm * p.ncell_k + cell];
#pragma unroll
            for (int j = 0; j < SD; j++) gx[j] = p.coords[(unsigned long long)j * p.npoint + pid];
        }
        if (cell + gstep < p.hi) gpid = p.conn[(unsigned long long)gm * p.ncell_k + cell + gstep];
    }

    const unsigned long long ntile = (p.hi - p.lo + CPC - 1) / CPC;
    for (unsigned long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const unsigned long long cell0 = p.lo + tile * CPC;
        const int ncell_tile = (int)min((unsigned long long)CPC, p.hi - cell0);
        __syncthreads();

        // ---- phase A: gather (Mesh::set_pad) ---------------------------------------------------------------------------
        if (gather_lane) {
            const unsigned long long cell = cell0 + gc;
            if (gc < ncell_tile) {
#pragma unroll
                for (int j = 0; j < SD; j++) s_X[gc * XS + gm * SD + j] = gx[j];
            }
            if (cell + gstep < p.hi) {
#pragma unroll
                for (int j = 0; j < SD; j++) gx[j] = p.coords[(unsigned long long)j * p.npoint + gpid];
            }
            if (cell + 2 * gstep < p.hi) gpid = p.conn[(unsigned long long)gm * p.ncell_k + cell + 2 * gstep];
        }
        __syncthreads();

        // ---- phase B: Jacobian (and gradient when a B-based case is requested) ------------------------------------------
        for (int task = tid; task < ncell_tile * ng; task += nt) {
            const int gp = task / ncell_tile, c = task - gp * ncell_tile;
            const double *X = s_X + c * XS;
            const double *L = s_L + gp * NN * SD;
            double J[SD][SD];
#pragma unroll
            for (int i = 0; i < SD; i++)
#pragma unroll
                for (int j = 0; j < SD; j++) J[i][j] = 0.0;
#pragma unroll 4
            for (int m = 0; m < NN; m++) geom::jacobian_add_node<SD, SD>(J, X + m * SD, L + m * SD);
            double Ji[SD][SD];
            const double det = geom::invert(J, Ji);
            if (fabs(det) <= ZERO_DETERMINANT) {
                const unsigned long long cell = cell0 + c;
                atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
            }
            double *G = s_G + (c * ng + gp) * gs;
            if (need_b) {
#pragma unroll 2
                for (int m = 0; m < NN; m++) {
#pragma unroll
                    for (int j = 0; j < SD; j++) {
                        double sum = L[m * SD] * Ji[0][j];
#pragma unroll
                        for (int k = 1; k < SD; k++) sum = fma(L[m * SD + k], Ji[k][j], sum);
                        G[1 + m * SD + j] = sum;
                    }
                }
            }
            double coef = det * s_w[gp] * p.alpha;
            if constexpr (AXI) {
                const double *N = s_N + gp * NN;
                double r = 0.0;
                for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], X[m * SD])); // radius_at_ip: r = sum N.x0
                coef *= r;
            }
            G[0] = coef;
        }
        __syncthreads();

        // ---- phase C: lane = column node n ---------------------------------------------------------------------------------
        double acc1[ND], acc3[ND], a1 = 0.0, a3 = 0.0;
#pragma unroll
        for (int b = 0; b < ND; b++) acc1[b] = acc3[b] = 0.0;
        const bool active = worker && cs < ncell_tile;
        if (active) {
#pragma unroll 1
            for (int gp = 0; gp < ng; gp++) {
                const double *G = s_G + (cs * ng + gp) * gs;
                const double *N = s_N + gp * NN;
                const double c = G[0];
                if (m01 || v01) {
                    const double nn = N[n];
                    if (v01) a1 = fma(nn, f.s_vec * c, a1);
                    if (m01) {
                        const double tn = nn * (f.s_mat * c);
#pragma unroll
                        for (int b = 0; b < ND; b++) {
                            int m = n + b;
                            if (m >= NN) m -= NN;
                            acc1[b] = fma(N[m], tn, acc1[b]);
                        }
                    }
                }
                if (need_b) {
                    double bn[SD];
#pragma unroll
                    for (int k = 0; k < SD; k++) bn[k] = G[1 + n * SD + k];
                    if (v03) {
                        double dot = f.w[0] * bn[0];
#pragma unroll
                        for (int k = 1; k < SD; k++) dot = fma(f.w[k], bn[k], dot);
                        a3 = fma(c, dot, a3);
                    }
                    if (m03) {
                        double t[SD];
#pragma unroll
                        for (int a = 0; a < SD; a++) {
                            double sum = tt[a][0] * bn[0];
#pragma unroll
                            for (int b = 1; b < SD; b++) sum = fma(tt[a][b], bn[b], sum);
                            t[a] = sum * c;
                        }
#pragma unroll
                        for (int b = 0; b < ND; b++) {
                            int m = n + b;
                            if (m >= NN) m -= NN;
                            double v = acc3[b];
#pragma unroll
                            for (int k = 0; k < SD; k++) v = fma(G[1 + m * SD + k], t[k], v);
                            acc3[b] = v;
                        }
                    }
                }
            }
        }

        // ---- phase D: vectors straight from registers (contiguous across the CTA), matrices through shared memory --------
        if (active) {
            const unsigned long long cell = cell0 + cs;
            const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
            const unsigned long long voff = f.voff ? (f.voff[cell] - f.vbase) : (orig - p.cell_begin) * NN;
            if (v01) f.out_v01[voff + n] = a1;
            if (v03) f.out_v03[voff + n] = a3;
#pragma unroll
            for (int b = 0; b < ND; b++) {
                int m = n + b;
                if (m >= NN) m -= NN;
                const bool self_paired = (NN % 2 == 0) && b == NN / 2; // lane n + NN/2 holds the mirrored entry itself
                if (m01) {
                    s_K1[cs * KK + m + NN * n] = acc1[b];
                    if (b > 0 && !self_paired) s_K1[cs * KK + n + NN * m] = acc1[b];
                }
                if (m03) {
                    s_K3[cs * KK + m + NN * n] = acc3[b];
                    if (b > 0 && !self_paired) s_K3[cs * KK + n + NN * m] = acc3[b];
                }
            }
        }
        __syncthreads();
        if (m01 || m03) {
            for (int idx = tid; idx < ncell_tile * KK; idx += nt) {
                const int c = idx / KK, q = idx - c * KK;
                const unsigned long long cell = cell0 + c;
                const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
                const unsigned long long moff = f.moff ? (f.moff[cell] - f.mbase) : (orig - p.cell_begin) * KK;
                if (m01) f.out_m01[moff + q] = s_K1[idx];
                if (m03) f.out_m03[moff + q] = s_K3[idx];
            }
        }
    }
}

template <int NN, int SD>
int launch_kind(const IntegParams &p, const ScalarFused &f, void *stream) {
    const bool axi = p.axisym != 0;
    const bool need_b = (f.mask & 2) || (f.mask & 8);
    const int nmat = ((f.mask & 1) ? 1 : 0) + ((f.mask & 2) ? 1 : 0);
    const int gs = need_b ? ((1 + NN * SD) | 1) : 1;
    const int xs = (NN * SD) | 1;
    auto smem_of = [&](int cpc) {
        return sizeof(double) * ((size_t)nmat * cpc * NN * NN + (size_t)p.ng * (1 + NN + NN * SD) + (size_t)cpc * xs + (size_t)cpc * p.ng * gs);
    };
    int cpc = 256 / NN;
    while (cpc > 1 && smem_of(cpc) > 72 * 1024) cpc--;
    const size_t smem = smem_of(cpc);
    if (smem > 200 * 1024) return 0;
    const int used = cpc * NN;
    const int threads = (used + 31) / 32 * 32;
    typedef void (*kern_t)(const IntegParams, const ScalarFused, const int, const int);
    kern_t kern;
    if constexpr (SD == 2) kern = axi ? scalar_kernel<NN, SD, true> : scalar_kernel<NN, SD, false>;
    else kern = scalar_kernel<NN, SD, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    if (ctas_per_sm > 2048 / threads) ctas_per_sm = 2048 / threads;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long ntile = (p.hi - p.lo + cpc - 1) / cpc;
    unsigned long long grid = (unsigned long long)nsm * ctas_per_sm;
    if (grid > ntile) grid = ntile;
    if (grid == 0) return 0;
    kern<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(p, f, cpc, used);
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("scalar_fused<nn=%d,sd=%d,mask=%d>", NN, SD, f.mask);
    return 1;
}

} // namespace

int launch_scalar_fused(const IntegParams &p, const ScalarFused &f, int sd, void *stream) {
    if (tuning_env("GL_DISABLE_SCALAR")) return 0; // testing knob: force the generic kernel
    if (sd == 2) {
        switch (p.nnode) {
        case 3: return launch_kind<3, 2>(p, f, stream);
        case 4: return launch_kind<4, 2>(p, f, stream);
        case 6: return launch_kind<6, 2>(p, f, stream);
        case 8: return launch_kind<8, 2>(p, f, stream);
        case 9: return launch_kind<9, 2>(p, f, stream);
        default: return 0;
        }
    }
    // the 3D instantiations carry no radius: the reference multiplies by r = sum N.x0 whenever args.axisymmetric is set, whatever
    // the space dimension (mat_01_nsn.rs:78-84), so such calls go to the generic kernel, which honours the flag
    if (p.axisym) return 0;
    switch (p.nnode) {
    case 4: return launch_kind<4, 3>(p, f, stream);
    case 8: return launch_kind<8, 3>(p, f, stream);
    case 10: return launch_kind<10, 3>(p, f, stream);
    case 20: return launch_kind<20, 3>(p, f, stream);
    default: return 0;
    }
}

} // namespace gl
