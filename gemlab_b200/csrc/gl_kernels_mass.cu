// gl_kernels_mass.cu -- mass matrix and source vector from one geometry pass, contraction on the FP64 tensor cores:
//   integ::mat_01_nsn  K(m,n) = sum_p N(m) s N(n) c_p     (src/integ/mat_01_nsn.rs:48-102)
//   integ::vec_01_ns   a(m)   = sum_p N(m) s c_p           (src/integ/vec_01_ns.rs:83-130)
// with c_p = det(J_p) w_p alpha [r_p if axisymmetric], uniform s, tight blocks, clear = true  (BASELINE.json config 4: the
// Tet10 mass matrix and body-force vector).
//
// N does not depend on the cell, so  [K | a](cell, :) = C(cell, :) . W  with  C(cell, p) = c_p  and the per-(kind, rule)
// constant  W(p, m + n*nnode) = s N_p(m) N_p(n),  W(p, nnode^2 + m) = s' N_p(m):  a genuinely dense (cells x ng) . (ng x
// (nnode^2 + nnode)) product -- the one place on this path where DMMA is a real contraction (north_star).  The kernel:
//   phase 1  ONE THREAD PER CELL: gathers its nodes into registers (Mesh::set_pad, src/mesh/mesh.rs:448-456; coalesced
//            node-major connectivity) and evaluates det(J) at every Gauss point in the oracle's operation order
//            (gl_geom.cuh; no inverse is needed), L read as warp-uniform broadcasts from shared memory;
//   phase 2  the warp's 32 x ng coefficient matrix goes through shared memory into m8n8k4 A-fragments; W tiles are the
//            B-fragments; mma.sync.m8n8k4.f64 accumulates 8 cells x 8 output entries per tile;
//   phase 3  accumulator fragments are stored straight to global memory: the 8 lanes of a cell row write 64 contiguous
//            bytes of the element block -- no output staging.
// The reference calls calc_gradient in mat_01_nsn although B is unused (mat_01_nsn.rs:75-76); its only observable effect,
// the zero-determinant error, is kept.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int MASS_THREADS = 128;

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int SD>
__device__ __forceinline__ double det_of(const double (&J)[SD][SD]) {
    if constexpr (SD == 2) {
        return geom::sub(geom::mul(J[0][0], J[1][1]), geom::mul(J[0][1], J[1][0]));
    } else {
        const double c00 = geom::sub(geom::mul(J[1][1], J[2][2]), geom::mul(J[1][2], J[2][1]));
        const double c01 = geom::sub(geom::mul(J[1][0], J[2][2]), geom::mul(J[1][2], J[2][0]));
        const double c02 = geom::sub(geom::mul(J[1][0], J[2][1]), geom::mul(J[1][1], J[2][0]));
        return geom::add(geom::sub(geom::mul(J[0][0], c00), geom::mul(J[0][1], c01)), geom::mul(J[0][2], c02));
    }
}

template <int NN, int SD>
__global__ void __launch_bounds__(MASS_THREADS) mass_kernel(const IntegParams p, const ScalarFused f, const int KP, const int EP,
                                                            const int WS, const int CS) {
    constexpr int LS = NN * SD;
    constexpr int KK = NN * NN;
    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    double *s_W = smem;              // [KP][WS]   W(p, q), zero rows / columns of padding
    double *s_L = s_W + KP * WS;     // [ng][LS]
    double *s_N = s_L + ng * LS;     // [ng][NN]   (axisymmetric radius)
    double *s_w = s_N + ng * NN;     // [ng]
    double *s_C = s_w + ng;          // [warps][32][CS]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool m01 = f.mask & 1, v01 = f.mask & 4;
    // 16-byte stores of the accumulator pairs: block starts (multiples of NN^2 resp. NN) and the K / vector boundary must be even
    const bool vec2 = NN % 2 == 0 && f.moff == nullptr && f.voff == nullptr && (!m01 || (reinterpret_cast<uintptr_t>(f.out_m01) & 15) == 0) &&
                      (!v01 || (reinterpret_cast<uintptr_t>(f.out_v01) & 15) == 0) && p.cell_ids == nullptr;
    for (int i = tid; i < ng; i += MASS_THREADS) s_w[i] = p.rule[i];
    for (int i = tid; i < ng * NN; i += MASS_THREADS) s_N[i] = p.rule[ng + i];
    for (int i = tid; i < ng * LS; i += MASS_THREADS) s_L[i] = p.rule[ng * (1 + NN) + i];
    for (int i = tid; i < KP * WS; i += MASS_THREADS) {
        const int gp = i / WS, q = i - gp * WS;
        double v = 0.0;
        if (gp < ng) {
            const double *N = p.rule + ng + gp * NN;
            if (q < KK) v = m01 ? N[q % NN] * f.s_mat * N[q / NN] : 0.0; // N(m) s N(n), as the reference orders the product
            else if (q < KK + NN) v = v01 ? N[q - KK] * f.s_vec : 0.0;
        }
        s_W[i] = v;
    }
    __syncthreads();

    double *Cw = s_C + wid * 32 * CS;
    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + MASS_THREADS - 1) / MASS_THREADS;
    // the point ids of the NEXT batch are fetched while this one is integrated: one hop of the dependent gather off the critical path
    unsigned int pid_next[NN];
    auto load_ids = [&](unsigned long long c) {
#pragma unroll
        for (int m = 0; m < NN; m++) pid_next[m] = c < p.hi ? p.conn[(unsigned long long)m * p.ncell_k + c] : 0u;
    };
    load_ids(p.lo + (unsigned long long)blockIdx.x * MASS_THREADS + tid);
    for (unsigned long long batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const unsigned long long cell = p.lo + batch * MASS_THREADS + tid;
        const bool active = cell < p.hi;

        // ---- phase 1: this thread's cell -------------------------------------------------------------------------------
        double x[NN][SD];
        if (active) {
#pragma unroll
            for (int m = 0; m < NN; m++) {
                const unsigned int pid = pid_next[m];
                if (p.coords_pm) { // one 32-byte (3D) / 16-byte (2D) record per point: a single sector instead of SD
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
                } else {
#pragma unroll
                    for (int j = 0; j < SD; j++) x[m][j] = p.coords[(unsigned long long)j * p.npoint + pid];
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < NN; m++)
#pragma unroll
                for (int j = 0; j < SD; j++) x[m][j] = 0.0;
        }
        load_ids(cell + (unsigned long long)gridDim.x * MASS_THREADS);
        bool bad = false;
#pragma unroll 1
        for (int gp = 0; gp < ng; gp++) {
            const double *L = s_L + gp * LS;
            double J[SD][SD];
#pragma unroll
            for (int i = 0; i < SD; i++)
#pragma unroll
                for (int j = 0; j < SD; j++) J[i][j] = 0.0;
#pragma unroll
            for (int m = 0; m < NN; m++) geom::jacobian_add_node<SD, SD>(J, x[m], L + m * SD);
            const double det = det_of<SD>(J);
            bad = bad || fabs(det) <= ZERO_DETERMINANT;
            double coef = det * s_w[gp] * p.alpha;
            if (p.axisym) {
                const double *N = s_N + gp * NN;
                double r = 0.0;
#pragma unroll
                for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], x[m][0])); // radius_at_ip
                coef *= r;
            }
            Cw[lane * CS + gp] = active ? coef : 0.0;
        }
        for (int gp = ng; gp < KP; gp++) Cw[lane * CS + gp] = 0.0;
        if (active && bad) atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);

        // output offsets of this thread's cell, fetched by the fragment owners with shuffles
        unsigned long long moff = 0, voff = 0;
        if (active) {
            const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
            moff = f.moff ? (f.moff[cell] - f.mbase) : (orig - p.cell_begin) * KK;
            voff = f.voff ? (f.voff[cell] - f.vbase) : (orig - p.cell_begin) * NN;
        }
        __syncwarp();

        // ---- phase 2 + 3: [K | a] = C . W on the tensor cores, stored from the accumulator fragments -----------------------
        const int ar = lane >> 2, ak = lane & 3; // fragment coordinates: row (cell / output entry) and k index
        const int nks = KP / 4;
        // the A fragments (this warp's 32 x KP coefficient matrix) do not depend on the column tile: rules of up to 16 points keep them in
        // registers for the whole sweep instead of re-reading them from shared memory for each of the EP / 8 column tiles
        const bool a_in_regs = nks <= 4;
        double afr[4][4];
#pragma unroll
        for (int rt = 0; rt < 4; rt++)
#pragma unroll
            for (int ks = 0; ks < 4; ks++) afr[rt][ks] = (a_in_regs && ks < nks) ? Cw[(8 * rt + ar) * CS + 4 * ks + ak] : 0.0;
        for (int ct = 0; ct < EP / 8; ct++) {
            double acc[4][2];
#pragma unroll
            for (int rt = 0; rt < 4; rt++) acc[rt][0] = acc[rt][1] = 0.0;
            if (a_in_regs) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    if (ks < nks) {
                        const double b = s_W[(4 * ks + ak) * WS + 8 * ct + ar];
#pragma unroll
                        for (int rt = 0; rt < 4; rt++) dmma_m8n8k4(acc[rt][0], acc[rt][1], afr[rt][ks], b);
                    }
                }
            } else {
                for (int ks = 0; ks < nks; ks++) {
                    const double b = s_W[(4 * ks + ak) * WS + 8 * ct + ar];
#pragma unroll
                    for (int rt = 0; rt < 4; rt++) dmma_m8n8k4(acc[rt][0], acc[rt][1], Cw[(8 * rt + ar) * CS + 4 * ks + ak], b);
                }
            }
            const int q0 = 8 * ct + 2 * ak;
#pragma unroll
            for (int rt = 0; rt < 4; rt++) {
                const int src = 8 * rt + ar; // lane that owns this row's cell
                const unsigned long long mo = __shfl_sync(0xffffffffu, moff, src), vo = __shfl_sync(0xffffffffu, voff, src);
                const bool row_active = p.lo + batch * MASS_THREADS + (unsigned long long)(wid * 32 + src) < p.hi;
                if (!row_active) continue;
                if (vec2) {
                    // even NN on a single-kind mesh: the lane's two entries (columns q0, q0 + 1) are one aligned 16-byte store
                    if (q0 < KK) {
                        if (m01) *reinterpret_cast<double2 *>(f.out_m01 + mo + q0) = make_double2(acc[rt][0], acc[rt][1]);
                    } else if (q0 < KK + NN) {
                        if (v01) *reinterpret_cast<double2 *>(f.out_v01 + vo + (q0 - KK)) = make_double2(acc[rt][0], acc[rt][1]);
                    }
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int q = q0 + e;
                    if (q < KK) {
                        if (m01) f.out_m01[mo + q] = acc[rt][e];
                    } else if (q < KK + NN) {
                        if (v01) f.out_v01[vo + (q - KK)] = acc[rt][e];
                    }
                }
            }
        }
        __syncwarp(); // Cw is rewritten by the next batch
    }
}

template <int NN, int SD>
int launch_mass_kind(const IntegParams &p, const ScalarFused &f, void *stream) {
    const int ng = p.ng;
    const int KP = (ng + 3) / 4 * 4;
    const int NE = NN * NN + NN;
    const int EP = (NE + 7) / 8 * 8;
    const int WS = (EP + 11) / 16 * 16 + 4;   // = 4 (mod 16): the 4 x 4 address pattern of a half-warp's fragment read is conflict-free
    const int CS = (KP + 11) / 16 * 16 + 4;
    const size_t smem = sizeof(double) * ((size_t)KP * WS + (size_t)ng * NN * SD + (size_t)ng * NN + ng + (size_t)(MASS_THREADS / 32) * 32 * CS);
    if (smem > 200 * 1024) return 0;
    auto kern = mass_kernel<NN, SD>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((226 * 1024) / (smem + 1024));
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long nbatch = (p.hi - p.lo + MASS_THREADS - 1) / MASS_THREADS;
    unsigned long long grid = (unsigned long long)nsm * ctas_per_sm * 8;
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    kern<<<(unsigned)grid, MASS_THREADS, smem, (cudaStream_t)stream>>>(p, f, KP, EP, WS, CS);
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("mass_dmma<nn=%d,sd=%d,mask=%d>", NN, SD, f.mask);
    return 1;
}

} // namespace

// mask bits as in launch_scalar_fused; covers masks made of mat_01_nsn (1) and vec_01_ns (4) only
int launch_mass_dmma(const IntegParams &p, const ScalarFused &f, int sd, void *stream) {
    if ((f.mask & ~5) != 0 || f.mask == 0) return 0;
    if (tuning_env("GL_DISABLE_MASS")) return 0; // testing knob: fall back to gl_kernels_scalar.cu
    if (sd == 2) {
        switch (p.nnode) {
        case 3: return launch_mass_kind<3, 2>(p, f, stream);
        case 4: return launch_mass_kind<4, 2>(p, f, stream);
        case 6: return launch_mass_kind<6, 2>(p, f, stream);
        case 8: return launch_mass_kind<8, 2>(p, f, stream);
        case 9: return launch_mass_kind<9, 2>(p, f, stream);
        default: return 0;
        }
    }
    if (p.axisym) return 0;
    switch (p.nnode) {
    case 4: return launch_mass_kind<4, 3>(p, f, stream);
    case 8: return launch_mass_kind<8, 3>(p, f, stream);
    case 10: return launch_mass_kind<10, 3>(p, f, stream);
    default: return 0;
    }
}

} // namespace gl
