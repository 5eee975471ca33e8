// gl_kernels_generic.cu -- kind-agnostic integration kernel of libgemlab_cuda (sm_100a).
//
// One CTA integrates a tile of `cells_per_cta` cells of ONE GeoKind bucket in three phases:
//   A  gather    : Mesh::set_pad (src/mesh/mesh.rs:448-456) -- connectivity is read coalesced from the node-major SoA
//                  table, coordinates are gathered from the SoA point array into shared memory;
//   B  geometry  : one thread per (cell, Gauss point): J = Xt.L, closed-form inverse and det, B = L.inv(J)
//                  (Scratchpad::calc_jacobian / calc_gradient, src/shapes/scratchpad_calc_jacobian.rs:101-131,
//                  scratchpad_calc_gradient.rs:78-91); N and L come from the per-(kind, rule) table staged in shared
//                  memory instead of being re-evaluated per cell;
//   C  contract  : threads sweep the tile's OUTPUT in address order (so every store is fully coalesced) and each
//                  accumulates one entry over the Gauss points (the `for p in 0..npoint` loops of src/integ/*.rs).
// The kernel is generic over GeoKind (nnode, ng and the tables are run-time data), over the seven integration cases and
// over the coefficient modes; hot configurations have register-blocked specialisations (gl_kernels_hex8.cu, ...).
#include <cuda_runtime.h>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int NT = 256;
constexpr double INV_SQRT_2 = 0.70710678118654752440084436210484903928;

__host__ __device__ __forceinline__ bool op_needs_gradient(int op) {
    return op == GL_OP_MAT_10_BDB || op == GL_OP_MAT_03_BTB || op == GL_OP_VEC_03_BV || op == GL_OP_VEC_04_BT ||
           op == GL_OP_MAT_02_BVN || op == GL_OP_MAT_09_NVB || op == GL_OP_MAT_04_NSB || op == GL_OP_MAT_07_BSN;
}
__host__ __device__ __forceinline__ bool op_two_pads(int op) { return op >= GL_OP_MAT_04_NSB && op <= GL_OP_MAT_07_BSN; }
__host__ __device__ __forceinline__ bool op_boundary(int op) { return op >= GL_OP_MAT_01_NSN_BRY && op <= GL_OP_NORMAL; }

// column (m, i) of the Mandel strain-displacement matrix (calc_bb_matrix, src/integ/mat_10_bdb.rs:238-278):
// at most 3 non-zero rows; `hoop` = N[m]/r is the axisymmetric row 2 (only for i == 0 in 2D)
template <int SD>
__device__ __forceinline__ int strain_column(int i, const double *bm, bool axisym, double hoop, int *rows, double *vals) {
    if (SD == 2) {
        if (i == 0) {
            rows[0] = 0; vals[0] = bm[0];
            rows[1] = 3; vals[1] = bm[1] * INV_SQRT_2;
            if (axisym) { rows[2] = 2; vals[2] = hoop; return 3; }
            return 2;
        }
        rows[0] = 1; vals[0] = bm[1];
        rows[1] = 3; vals[1] = bm[0] * INV_SQRT_2;
        return 2;
    }
    if (i == 0) {
        rows[0] = 0; vals[0] = bm[0];
        rows[1] = 3; vals[1] = bm[1] * INV_SQRT_2;
        rows[2] = 5; vals[2] = bm[2] * INV_SQRT_2;
    } else if (i == 1) {
        rows[0] = 1; vals[0] = bm[1];
        rows[1] = 3; vals[1] = bm[0] * INV_SQRT_2;
        rows[2] = 4; vals[2] = bm[2] * INV_SQRT_2;
    } else {
        rows[0] = 2; vals[0] = bm[2];
        rows[1] = 4; vals[1] = bm[1] * INV_SQRT_2;
        rows[2] = 5; vals[2] = bm[0] * INV_SQRT_2;
    }
    return 3;
}

// component (i, a) of the symmetric 2nd-order tensor stored as Mandel vector t (Tensor2::vector())
template <int SD>
__device__ __forceinline__ double mandel_comp(const double *t, int i, int a) {
    if (i == a) return t[i];
    if (SD == 2) return t[3] * INV_SQRT_2;
    const int s = i + a; // (0,1)->1, (1,2)->3, (0,2)->2
    return (s == 1 ? t[3] : (s == 3 ? t[4] : t[5])) * INV_SQRT_2;
}

template <int SD, int GD>
__global__ void __launch_bounds__(NT) integ_generic_kernel(const IntegParams p) {
    extern __shared__ double smem[];
    const int nnode = p.nnode, ng = p.ng, C = p.cells_per_cta, op = p.op;
    const bool need_b = op_needs_gradient(op);
    const bool axisym = p.axisym != 0;
    const int tid = threadIdx.x;

    // ---- shared-memory carve-up -------------------------------------------------------------------
    double *s_w = smem;                   // [ng]
    double *s_N = s_w + ng;               // [ng][nnode]
    double *s_L = s_N + ng * nnode;       // [ng][nnode][GD]
    const int xs = (nnode * SD) | 1;      // odd cell stride: conflict-free when threads walk over cells
    double *s_X = s_L + ng * nnode * GD;  // [C][xs]
    double *s_cf = s_X + C * xs;          // [C][ng]  det(J) * w * alpha
    double *s_rr = s_cf + C * ng;         // [C][ng]  radius (axisymmetric)
    double *s_D = s_rr + C * ng;          // [36]     uniform coefficient record
    const int bs = nnode * SD;
    const int bcs = (ng * bs) | 1;        // odd cell stride of the gradient block
    double *s_B = s_D + 36;               // [C][bcs] gradients B[gp][m][j]
    // second pad (two-pad cases): Nb, Lb at the Gauss points and, for mat_05_btn, its own gradients Bb
    const int nnode_b = p.nnode_b;
    const int bsb = nnode_b * SD;
    const int bcsb = (ng * bsb) | 1;
    double *s_Nb = s_B + (need_b ? C * bcs : 0);          // [ng][nnode_b]
    double *s_Lb = s_Nb + ng * nnode_b;                    // [ng][nnode_b][GD]
    double *s_Bb = s_Lb + ng * nnode_b * GD;               // [C][bcsb]   (mat_05_btn only)
    double *s_un = s_Bb + (op == GL_OP_MAT_05_BTN ? C * bcsb : 0); // [C][ng][SD] unit normals (boundary cases)

    const int rule_len = ng * (1 + nnode + nnode * GD);
    for (int i = tid; i < rule_len; i += NT) smem[i] = p.rule[i];
    if (op_two_pads(op))
        for (int i = tid; i < ng * nnode_b * (1 + GD); i += NT) s_Nb[i] = p.rule_b[i];
    if (tid < 36) s_D[tid] = p.coef_uniform[tid];

    const unsigned int nrow = p.nrow, ncol = p.ncol;
    const unsigned int block = nrow * ncol;
    const unsigned long long ntile = (p.hi - p.lo + C - 1) / C;

    for (unsigned long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const unsigned long long cell0 = p.lo + tile * C;
        const int ncell_tile = (int)min((unsigned long long)C, p.hi - cell0);
        __syncthreads(); // previous tile fully consumed (also covers the table staging on the first pass)

        // ---- phase A: gather coordinates (Mesh::set_pad) ----------------------------------------------
        for (int idx = tid; idx < ncell_tile * nnode; idx += NT) {
            const int m = idx / ncell_tile;
            const int c = idx - m * ncell_tile;
            const unsigned int pid = p.conn[(unsigned long long)m * p.ncell_k + cell0 + c];
#pragma unroll
            for (int j = 0; j < SD; j++) s_X[c * xs + m * SD + j] = p.coords[(unsigned long long)j * p.npoint + pid];
        }
        __syncthreads();

        // ---- phase B: Jacobian, inverse, gradient at every Gauss point --------------------------------
        for (int task = tid; task < ncell_tile * ng; task += NT) {
            const int gp = task / ncell_tile;
            const int c = task - gp * ncell_tile;
            const double *X = s_X + c * xs;
            const double *L = s_L + gp * nnode * GD;
            double J[SD][GD];
#pragma unroll
            for (int i = 0; i < SD; i++)
#pragma unroll
                for (int j = 0; j < GD; j++) J[i][j] = 0.0;
            for (int m = 0; m < nnode; m++) {
#pragma unroll
                for (int i = 0; i < SD; i++)
#pragma unroll
                    for (int j = 0; j < GD; j++) J[i][j] = geom::add(J[i][j], geom::mul(X[m * SD + i], L[m * GD + j]));
            }
            double det;
            double Ji[SD][SD];
            if constexpr (GD == SD) {
                // SOLID: closed-form inverse, as russell_lab::mat_inverse does for 2x2 / 3x3 (cofactors divided by det)
                det = geom::invert(J, Ji);
                if (fabs(det) <= ZERO_DETERMINANT) {
                    const unsigned long long cell = cell0 + c;
                    atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
                }
                if (need_b) {
                    double *Bc = s_B + c * bcs + gp * bs;
                    for (int m = 0; m < nnode; m++) {
#pragma unroll
                        for (int j = 0; j < SD; j++) {
                            double sum = geom::mul(L[m * GD], Ji[0][j]);
#pragma unroll
                            for (int k = 1; k < SD; k++) sum = geom::add(sum, geom::mul(L[m * GD + k], Ji[k][j]));
                            Bc[m * SD + j] = sum;
                        }
                    }
                }
                if (op == GL_OP_MAT_05_BTN) {
                    // pad_b.calc_gradient (src/integ/mat_05_btn.rs:106): the second pad's own Jacobian from ITS nodes (the cell's
                    // first nnode_b) and its own derivatives
                    const double *Lb = s_Lb + gp * nnode_b * GD;
                    double Jb[SD][GD];
#pragma unroll
                    for (int i = 0; i < SD; i++)
#pragma unroll
                        for (int j = 0; j < GD; j++) Jb[i][j] = 0.0;
                    for (int m = 0; m < nnode_b; m++) {
#pragma unroll
                        for (int i = 0; i < SD; i++)
#pragma unroll
                            for (int j = 0; j < GD; j++) Jb[i][j] = geom::add(Jb[i][j], geom::mul(X[m * SD + i], Lb[m * GD + j]));
                    }
                    double Jbi[SD][SD];
                    const double detb = geom::invert(Jb, Jbi);
                    if (fabs(detb) <= ZERO_DETERMINANT) {
                        const unsigned long long cell = cell0 + c;
                        atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
                    }
                    double *Bbc = s_Bb + c * bcsb + gp * bsb;
                    for (int m = 0; m < nnode_b; m++) {
#pragma unroll
                        for (int j = 0; j < SD; j++) {
                            double sum = geom::mul(Lb[m * GD], Jbi[0][j]);
#pragma unroll
                            for (int k = 1; k < SD; k++) sum = geom::add(sum, geom::mul(Lb[m * GD + k], Jbi[k][j]));
                            Bbc[m * SD + j] = sum;
                        }
                    }
                }
            } else if constexpr (GD == 1) { // CABLE: norm of the Jacobian vector
                double n2 = 0.0;
#pragma unroll
                for (int i = 0; i < SD; i++) n2 = geom::add(n2, geom::mul(J[i][0], J[i][0]));
                det = sqrt(n2);
                (void)Ji;
                if (SD == 2 && op_boundary(op)) {
                    // Scratchpad::calc_normal_vector, line in 2D (scratchpad_calc_normal_vector.rs:142-151): n = e3 x g1
                    double u0 = -J[1][0], u1 = J[0][0];
                    const double mag = sqrt(geom::add(geom::mul(u0, u0), geom::mul(u1, u1)));
                    if (mag > 0.0) {
                        u0 /= mag;
                        u1 /= mag;
                    }
                    det = mag;
                    s_un[(c * ng + gp) * SD] = u0;
                    s_un[(c * ng + gp) * SD + 1] = u1;
                }
            } else { // SHELL: DET_JAC_NOT_AVAILABLE
                det = -1.0;
                (void)Ji;
                if (SD == 3 && op_boundary(op)) {
                    // surface in 3D (scratchpad_calc_normal_vector.rs:166-175): n = g1 x g2
                    double u0 = geom::sub(geom::mul(J[1][0], J[2][GD - 1]), geom::mul(J[2][0], J[1][GD - 1]));
                    double u1 = geom::sub(geom::mul(J[2][0], J[0][GD - 1]), geom::mul(J[0][0], J[2][GD - 1]));
                    double u2 = geom::sub(geom::mul(J[0][0], J[1][GD - 1]), geom::mul(J[1][0], J[0][GD - 1]));
                    const double mag = sqrt(geom::add(geom::add(geom::mul(u0, u0), geom::mul(u1, u1)), geom::mul(u2, u2)));
                    if (mag > 0.0) {
                        u0 /= mag;
                        u1 /= mag;
                        u2 /= mag;
                    }
                    det = mag;
                    s_un[(c * ng + gp) * SD] = u0;
                    s_un[(c * ng + gp) * SD + 1] = u1;
                    s_un[(c * ng + gp) * SD + (SD - 1)] = u2;
                }
            }
            double r = 0.0;
            if (axisym) {
                const double *N = s_N + gp * nnode;
                for (int m = 0; m < nnode; m++) r = geom::add(r, geom::mul(N[m], X[m * SD]));
            }
            s_cf[c * ng + gp] = (op == GL_OP_JACOBIAN || op == GL_OP_NORMAL) ? det : det * s_w[gp] * p.alpha;
            s_rr[c * ng + gp] = r;
        }
        __syncthreads();

        // ---- phase C: sweep the tile's output in address order ----------------------------------------
        const unsigned int total = (unsigned int)ncell_tile * block;
        for (unsigned int idx = tid; idx < total; idx += NT) {
            const int c = idx / block;
            const unsigned int q = idx - c * block;
            const unsigned int col = q / nrow;
            const unsigned int row = q - col * nrow;
            const unsigned long long cell = cell0 + c;
            const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
            const unsigned long long blk = p.out_off ? (p.out_off[cell] - p.out_base) : (orig - p.cell_begin) * block;
            double *dst = p.out + blk + q;
            const bool inside = row >= p.ii0 && row < p.ii0 + p.size_r && col >= p.jj0 && col < p.jj0 + p.size_c;
            if (!inside) {
                if (p.clear) *dst = 0.0;
                continue;
            }
            const int ii = row - p.ii0, jj = col - p.jj0;
            const double *cf = s_cf + c * ng;
            const double *rr = s_rr + c * ng;
            const double *Bc = s_B + c * bcs;
            const double *coef0;
            if (p.coef == nullptr) {
                coef0 = s_D;
            } else {
                const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell] : (orig - p.coef_begin);
                coef0 = p.coef + rec * p.coef_cell_stride;
            }
            const unsigned long long gstride = (p.coef == nullptr) ? 0ull : p.coef_gp_stride;
            double acc = 0.0;
            switch (op) {
            case GL_OP_MAT_10_BDB: {
                const int m = ii / SD, i = ii - m * SD, n = jj / SD, j = jj - n * SD;
                constexpr int ns = 2 * SD;
                for (int gp = 0; gp < ng; gp++) {
                    const double *d = coef0 + gp * gstride;
                    const double *bm = Bc + gp * bs + m * SD;
                    const double *bn = Bc + gp * bs + n * SD;
                    int ra[3], rb[3];
                    double va[3], vb[3];
                    double hm = 0.0, hn = 0.0, coef = cf[gp];
                    if (axisym) {
                        const double r = rr[gp];
                        hm = s_N[gp * nnode + m] / r;
                        hn = s_N[gp * nnode + n] / r;
                        coef *= r;
                    }
                    const int na = strain_column<SD>(i, bm, axisym, hm, ra, va);
                    const int nb = strain_column<SD>(j, bn, axisym, hn, rb, vb);
                    double sum = 0.0;
                    for (int a = 0; a < na; a++) {
                        double t = 0.0;
                        for (int b = 0; b < nb; b++) t += d[ra[a] + rb[b] * ns] * vb[b];
                        sum += va[a] * t;
                    }
                    acc += coef * sum;
                }
            } break;
            case GL_OP_MAT_01_NSN: {
                for (int gp = 0; gp < ng; gp++) {
                    double c1 = coef0[gp * gstride] * cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += s_N[gp * nnode + ii] * c1 * s_N[gp * nnode + jj];
                }
            } break;
            case GL_OP_MAT_03_BTB: {
                for (int gp = 0; gp < ng; gp++) {
                    const double *t = coef0 + gp * gstride;
                    const double *bm = Bc + gp * bs + ii * SD;
                    const double *bn = Bc + gp * bs + jj * SD;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    double sum = 0.0;
#pragma unroll
                    for (int a = 0; a < SD; a++) {
                        double ta = 0.0;
#pragma unroll
                        for (int b = 0; b < SD; b++) ta += mandel_comp<SD>(t, a, b) * bm[b];
                        sum += bn[a] * ta;
                    }
                    acc += c1 * sum;
                }
            } break;
            case GL_OP_MAT_02_BVN: { // K(m,n) = sum_p (B(m).v) N(n) c_p   (src/integ/mat_02_bvn.rs:96-117)
                for (int gp = 0; gp < ng; gp++) {
                    const double *v = coef0 + gp * gstride;
                    const double *bm = Bc + gp * bs + ii * SD;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    double dot = 0.0;
#pragma unroll
                    for (int a = 0; a < SD; a++) dot += bm[a] * v[a];
                    acc += c1 * dot * s_N[gp * nnode + jj];
                }
            } break;
            case GL_OP_MAT_08_NTN: { // K(m i, n j) = sum_p N(m) T(i,j) N(n) c_p   (src/integ/mat_08_ntn.rs:109-131)
                const int m = ii / SD, i = ii - m * SD, n = jj / SD, j = jj - n * SD;
                for (int gp = 0; gp < ng; gp++) {
                    const double *t = coef0 + gp * gstride;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += c1 * s_N[gp * nnode + m] * mandel_comp<SD>(t, i, j) * s_N[gp * nnode + n];
                }
            } break;
            case GL_OP_MAT_09_NVB: { // K(m i, n j) = sum_p N(m) v(i) B(n,j) c_p   (src/integ/mat_09_nvb.rs:101-122)
                const int m = ii / SD, i = ii - m * SD, n = jj / SD, j = jj - n * SD;
                for (int gp = 0; gp < ng; gp++) {
                    const double *v = coef0 + gp * gstride;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += c1 * s_N[gp * nnode + m] * v[i] * Bc[gp * bs + n * SD + j];
                }
            } break;
            case GL_OP_VEC_01_NS: {
                for (int gp = 0; gp < ng; gp++) {
                    double c1 = coef0[gp * gstride] * cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += s_N[gp * nnode + ii] * c1;
                }
            } break;
            case GL_OP_VEC_02_NV: {
                const int m = ii / SD, i = ii - m * SD;
                for (int gp = 0; gp < ng; gp++) {
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += c1 * s_N[gp * nnode + m] * coef0[gp * gstride + i];
                }
            } break;
            case GL_OP_VEC_03_BV: {
                for (int gp = 0; gp < ng; gp++) {
                    const double *w = coef0 + gp * gstride;
                    const double *bm = Bc + gp * bs + ii * SD;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    double sum = 0.0;
#pragma unroll
                    for (int a = 0; a < SD; a++) sum += w[a] * bm[a];
                    acc += c1 * sum;
                }
            } break;
            case GL_OP_VEC_04_BT: {
                const int m = ii / SD, i = ii - m * SD;
                for (int gp = 0; gp < ng; gp++) {
                    const double *t = coef0 + gp * gstride;
                    const double *bm = Bc + gp * bs + m * SD;
                    double sum = 0.0;
#pragma unroll
                    for (int a = 0; a < SD; a++) sum += mandel_comp<SD>(t, i, a) * bm[a];
                    if (axisym) {
                        acc += cf[gp] * rr[gp] * sum;
                        if (i == 0) acc += cf[gp] * s_N[gp * nnode + m] * t[2];
                    } else {
                        acc += cf[gp] * sum;
                    }
                }
            } break;
            case GL_OP_MAT_04_NSB: { // K(m, n j) = sum_p Nb(m) s c_p B(n, j)   (src/integ/mat_04_nsb.rs:115-133)
                const int n = jj / SD, j = jj - n * SD;
                for (int gp = 0; gp < ng; gp++) {
                    double c1 = coef0[gp * gstride] * cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += s_Nb[gp * nnode_b + ii] * c1 * Bc[gp * bs + n * SD + j];
                }
            } break;
            case GL_OP_MAT_05_BTN: { // K(m, n j) = sum_p (Bb(m).T)_j N(n) c_p   (src/integ/mat_05_btn.rs:128-164)
                const int n = jj / SD, j = jj - n * SD;
                const double *Bbc = s_Bb + c * bcsb;
                for (int gp = 0; gp < ng; gp++) {
                    const double *t = coef0 + gp * gstride;
                    const double *bm = Bbc + gp * bsb + ii * SD;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    double sum = 0.0;
#pragma unroll
                    for (int a = 0; a < SD; a++) sum += bm[a] * mandel_comp<SD>(t, a, j);
                    acc += c1 * sum * s_N[gp * nnode + n];
                }
            } break;
            case GL_OP_MAT_06_NVN: { // K(m i, n) = sum_p N(m) v_i Nb(n) c_p   (src/integ/mat_06_nvn.rs:119-136)
                const int m = ii / SD, i = ii - m * SD;
                for (int gp = 0; gp < ng; gp++) {
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += c1 * s_N[gp * nnode + m] * coef0[gp * gstride + i] * s_Nb[gp * nnode_b + jj];
                }
            } break;
            case GL_OP_MAT_07_BSN: { // K(m i, n) = sum_p B(m, i) s c_p Nb(n)   (src/integ/mat_07_bsn.rs:114-131)
                const int m = ii / SD, i = ii - m * SD;
                for (int gp = 0; gp < ng; gp++) {
                    double c1 = coef0[gp * gstride] * cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += Bc[gp * bs + m * SD + i] * c1 * s_Nb[gp * nnode_b + jj];
                }
            } break;
            case GL_OP_MAT_01_NSN_BRY: case GL_OP_VEC_01_NS_BRY: {
                // s = s0 + w.un (the data form of a closure that reads the unit normal); c = s mag_n w alpha [r]
                const double *un = s_un + c * ng * SD;
                for (int gp = 0; gp < ng; gp++) {
                    const double *rec = coef0 + gp * gstride;
                    double sv = rec[0];
#pragma unroll
                    for (int a = 0; a < SD; a++) sv += rec[1 + a] * un[gp * SD + a];
                    double c1 = sv * cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += (op == GL_OP_MAT_01_NSN_BRY) ? s_N[gp * nnode + ii] * c1 * s_N[gp * nnode + jj] : s_N[gp * nnode + ii] * c1;
                }
            } break;
            case GL_OP_VEC_02_NV_BRY: { // b(m i) = sum_p c_p N(m) (v_i + qn un_i)   (src/integ/vec_02_nv_bry.rs:113-127)
                const int m = ii / SD, i = ii - m * SD;
                const double *un = s_un + c * ng * SD;
                for (int gp = 0; gp < ng; gp++) {
                    const double *rec = coef0 + gp * gstride;
                    double c1 = cf[gp];
                    if (axisym) c1 *= rr[gp];
                    acc += c1 * s_N[gp * nnode + m] * (rec[i] + rec[SD] * un[gp * SD + i]);
                }
            } break;
            case GL_OP_NORMAL: { // row = gp * (SD + 1) + component: un[0..SD), mag_n
                const int gp = ii / (SD + 1), k = ii - gp * (SD + 1);
                acc = k < SD ? s_un[(c * ng + gp) * SD + k] : cf[gp];
            } break;
            default: // GL_OP_JACOBIAN
                acc = cf[0];
                break;
            }
            if (p.clear) *dst = acc;
            else *dst += acc;
        }
    }
}

template <int SD, int GD>
int launch_t(const IntegParams &p, void *stream) {
    const size_t smem = generic_smem_bytes(SD, GD, p.nnode, p.ng, p.op, p.cells_per_cta, p.nnode_b);
    // per launch: the attribute is per device, and one process may drive several (gl_multi_*)
    if (cudaFuncSetAttribute(integ_generic_kernel<SD, GD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
    const unsigned long long ntile = (p.hi - p.lo + p.cells_per_cta - 1) / p.cells_per_cta;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // persistent-style grid: a multiple of the SM count, each CTA strides over tiles
    int ctas_per_sm = (int)((200 * 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    unsigned long long grid = (unsigned long long)nsm * ctas_per_sm;
    if (grid > ntile) grid = ntile;
    if (grid == 0) return 0;
    integ_generic_kernel<SD, GD><<<(unsigned int)grid, NT, smem, (cudaStream_t)stream>>>(p);
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("generic<sd=%d,gd=%d>", SD, GD);
    return 1;
}

} // namespace

size_t generic_smem_bytes(int sd, int gd, int nnode, int ng, int op, int C, int nnode_b) {
    const bool need_b = op_needs_gradient(op);
    size_t n = (size_t)ng * (1 + nnode + nnode * gd);
    n += (size_t)C * ((nnode * sd) | 1);
    n += (size_t)2 * C * ng;
    n += 36;
    if (need_b) n += (size_t)C * ((ng * nnode * sd) | 1);
    if (op_two_pads(op)) n += (size_t)ng * nnode_b * (1 + gd);
    if (op == GL_OP_MAT_05_BTN) n += (size_t)C * ((ng * nnode_b * sd) | 1);
    if (op_boundary(op)) n += (size_t)C * ng * sd;
    return n * sizeof(double);
}

int generic_pick_cells_per_cta(int sd, int gd, int nnode, int ng, int op, int nnode_b) {
    // as many cells as fit in ~96 KB (two CTAs per SM), at most 64 and at least 1
    int best = 1;
    for (int c = 1; c <= 64; c++) {
        if (generic_smem_bytes(sd, gd, nnode, ng, op, c, nnode_b) <= 96 * 1024) best = c;
        else break;
    }
    return best;
}

int launch_generic(const IntegParams &p, int sd, int gd, void *stream) {
    if (sd == 2 && gd == 2) return launch_t<2, 2>(p, stream);
    if (sd == 3 && gd == 3) return launch_t<3, 3>(p, stream);
    if (sd == 2 && gd == 1) return launch_t<2, 1>(p, stream);
    if (sd == 3 && gd == 1) return launch_t<3, 1>(p, stream);
    if (sd == 3 && gd == 2) return launch_t<3, 2>(p, stream);
    return -1;
}

} // namespace gl
