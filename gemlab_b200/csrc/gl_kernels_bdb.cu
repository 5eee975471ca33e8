// gl_kernels_bdb.cu -- register-blocked stiffness kernel for every solid GeoKind: integ::mat_10_bdb
// (src/integ/mat_10_bdb.rs:121-233, incl. the axisymmetric branch :213-233) with a uniform modulus D, tight element
// blocks and clear = true.  It generalises the Hex8 kernel (gl_kernels_hex8.cu) to NN nodes in SD dimensions:
//
//   phase A  gather: Mesh::set_pad (src/mesh/mesh.rs:448-456), coalesced node-major connectivity;
//   phase B  one thread per (cell, Gauss point): J = Xt.L, closed-form inverse, det, B = L.inv(J), axisymmetric radius
//            (scratchpad_calc_jacobian.rs:101-131, scratchpad_calc_gradient.rs:78-91); B, c = det*w*alpha[*r] and N/r -> smem;
//   phase C  LPN lanes per COLUMN NODE n accumulate the GEOMETRIC TENSORS P(m,n)[k][l] = sum_p c_p v(m)[k] v(n)[l] of the
//            circulant half of the node pairs, m = n, n+1, .., n+NN/2 (mod NN), in registers (VD^2 DFMA per pair and Gauss
//            point, v(m) = (B(m,.) [, N(m)/r]));  P(n,m) = P(m,n)^T gives the other half;
//   phase D  D is the same at every Gauss point, so it is applied ONCE: K(m,n)[i][j] = sum_kl D_ikjl P(m,n)[k][l]
//            (add_to_kk with the Gauss sum moved inside).  Blocks are staged in smem in the final column-major layout
//            (mirrored when D is major-symmetric) and the CTA's cells leave with ONE bulk async copy (homogeneous meshes)
//            or a cooperative vector copy (mixed-kind meshes).
// M(n) is the Mandel strain-displacement column block of calc_bb_matrix (mat_10_bdb.rs:238-278) with the 1/sqrt(2)
// factors folded into D' on the host; in the axisymmetric case row 2 carries N/r and the coefficient is c*r.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

// ---- structure of the strain-displacement column (node, j): entries q = 0 .. col_n-1 -> (row, component) -----------
// component index < SD selects a gradient component, component == SD selects the hoop value N/r (2D axisymmetric)
__host__ __device__ constexpr int col_n(int sd, bool axi, int j) { return sd == 3 ? 3 : (j == 0 ? (axi ? 3 : 2) : 2); }
__host__ __device__ constexpr int col_row(int sd, int j, int q) {
    return sd == 3 ? (j == 0 ? (q == 0 ? 0 : (q == 1 ? 3 : 5)) : (j == 1 ? (q == 0 ? 1 : (q == 1 ? 3 : 4)) : (q == 0 ? 2 : (q == 1 ? 4 : 5))))
                   : (j == 0 ? (q == 0 ? 0 : (q == 1 ? 3 : 2)) : (q == 0 ? 1 : 3));
}
__host__ __device__ constexpr int col_comp(int sd, int j, int q) {
    return sd == 3 ? (j == 0 ? (q == 0 ? 0 : (q == 1 ? 1 : 2)) : (j == 1 ? (q == 0 ? 1 : (q == 1 ? 0 : 2)) : (q == 0 ? 2 : (q == 1 ? 1 : 0))))
                   : (j == 0 ? (q == 0 ? 0 : (q == 1 ? 1 : 2)) : (q == 0 ? 1 : 0));
}
// PAT 0 general, PAT 1 major-symmetric, PAT 2 symmetric with uncoupled normal / diagonal shear blocks (exact zeros)
__host__ __device__ constexpr bool d_nz(int pat, int sd, int a, int b) {
    return pat < 2 ? true : ((a < 3 && b < 3) || a == b);
}
#ifndef GL_BDB_GP_UNROLL
#define GL_BDB_GP_UNROLL 1
#endif
constexpr int GP_UNROLL = GL_BDB_GP_UNROLL; // unroll factor of the Gauss-point loop of phase C

struct BdbConsts {
    double d[36]; // d'[a][b] = f_a f_b d[a][b] (row-major, ns x ns), f = 1 for normal rows, 1/sqrt(2) for shear rows
    int gen;      // 1: general store epilogue (nrow / ncol / ii0 / jj0, clear = false, 8-byte aligned output) instead of the bulk copy
                  // 2: fused assembly -- blocks are staged transposed (row-major) and every entry is added to its slot of the global
                  //    matrix (IntegParams::scatter_pos / scatter_vals); no element block is written
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// PG = one modulus per GAUSS POINT (the reference closure's natural use, mat_10_bdb.rs:156-159: every nonlinear material).  The
// Gauss sum can no longer move inside, so phase C accumulates the element blocks themselves,
//   T(n) = c_p D'_p M(n)  (NS x SD, once per column node and Gauss point),   K(m,n) += M(m)^T T(n)  (SD col_n FMA per entry),
// exploiting the sparsity of the strain-displacement columns (3 non-zeros of 6 in 3D), which a dense tensor-core tiling of
// K = sum_p M_p^T (c_p D_p) M_p cannot (Hex8: 24 kflop per cell here against 55 kflop as 8x8x4 DMMA tiles on the same pipe).
// PAT 1 = every D_p is major-symmetric (circulant half + mirror), PAT 0 = general (all NN distances, no mirror).  The PAT 1
// variant compares D_p with its transpose while staging it and lowers *p.pg_flag when a record is not symmetric; the PAT 0
// variant is enqueued right behind it and returns at once unless that happened -- no probe pass, no host synchronisation.
template <int NN, int SD, int LPN, bool AXI, int PAT, bool PG = false>
__global__ void __launch_bounds__(256) bdb_kernel(const IntegParams p, const BdbConsts cst, const int CPC, const int nthreads_used, const int alias_g) {
    if constexpr (PG) {
        if (PAT == 0 && *p.pg_flag != 0) return; // every record was symmetric: the PAT 1 variant's result stands
    } else {
        if (p.coef_pat_dev != nullptr && *p.coef_pat_dev != PAT) return; // another pattern variant owns this call (see launcher)
    }
    constexpr bool SYM = PAT >= 1;
    constexpr int NS = 2 * SD;
    constexpr int NDOF = NN * SD;
    // circulant distances 0 .. ND-1 (NN odd: (NN+1)/2 = NN/2+1 too).  The half is enough for EVERY pattern: P(n,m) = P(m,n)^T, and for a
    // non-symmetric D the `transposed` store of phase D contracts that transpose itself.  (With ND = NN every off-diagonal block
    // was written twice, by two lanes whose values round differently: a shared-memory write race.)
    constexpr int ND = (PG && !SYM) ? NN : NN / 2 + 1;
    constexpr int NSLOT = (ND + LPN - 1) / LPN;        // blocks per lane
    constexpr int GS = (1 + NN * SD + (AXI ? NN : 0)) | 1; // record per (cell, gauss point): c | B | N/r ; odd stride
    constexpr int XS = (NN * SD) | 1;
    constexpr int LS = (NN * SD + 6) / 16 * 16 + 9;   // = 9 (mod 16): the SD lanes of up to 6 Gauss points of a half-warp read L conflict-free
    constexpr int VD = SD + (AXI ? 1 : 0);            // components of the per-node vector v(m) = (B(m,.) [, N(m)/r])

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    double *s_K = smem;                               // [CPC][NDOF*NDOF]   (first: 16-byte aligned for the bulk copy)
    double *s_w = s_K + CPC * NDOF * NDOF;            // [ng]
    double *s_N = s_w + ng;                           // [ng][NN]
    double *s_L = s_N + ng * NN;                      // [ng][LS]  L[gp][m*SD+j]; odd stride: lanes of different Gauss points hit different banks
    double *s_X = s_L + ng * LS;                      // [CPC][XS]
    // [CPC][ng][GS]; when a cell's records fit inside its element block they ALIAS the K staging (dead before K is staged,
    // see gl_kernels_tile.cu) -- roughly halves the shared memory of a CTA, i.e. doubles the resident warps
    const bool drec = !PG && p.coef != nullptr;       // one modulus record per marker / per cell instead of a uniform D
    double *s_D = s_X + CPC * XS;                     // [CPC][NS*NS] d' of every cell of the tile (record moduli only)
    double *s_G = alias_g ? smem : s_D + (drec ? CPC * NS * NS : 0);
    // PG: [CPC][ng][NS*NS] d'_p of every (cell, Gauss point), row-major, at an even (16-byte aligned) offset behind the records
    double *s_DG = smem + ((((int)(s_G - smem) + CPC * ng * GS) + 1) & ~1);

    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < ng * (1 + NN); i += nt) s_w[i] = p.rule[i];
    for (int i = tid; i < ng * NN * SD; i += nt) s_L[(i / (NN * SD)) * LS + i % (NN * SD)] = p.rule[ng * (1 + NN) + i];

    // lane roles in phase C: part h (warp-friendly: parts occupy consecutive thread ranges), cell slot cs, node n
    const int per_part = CPC * NN;
    const bool worker = tid < nthreads_used;
    const int h = worker ? tid / per_part : 0;
    const int cs = worker ? (tid - h * per_part) / NN : 0;
    const int n = worker ? (tid - h * per_part) % NN : 0;

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
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // previous bulk copy has left s_K (/ s_G)
        __syncthreads(); // previous tile consumed; tables staged

        // ---- phase A: gather ---------------------------------------------------------------------------------------
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

        if (drec) {
            // the cell's modulus record (ns x ns Mandel matrix, column-major): of its marker (PER_MARKER) or its own (PER_CELL);
            // the 1/sqrt(2) factors of the shear rows are folded in here
            const double hs = 0.70710678118654752440084436210484903928;
            for (int t = tid; t < ncell_tile * NS * NS; t += nt) {
                const int c = t / (NS * NS), idx = t - c * (NS * NS);
                const int a = idx / NS, b = idx - a * NS;
                const unsigned long long cell = cell0 + c;
                const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell]
                                                            : (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.coef_begin;
                s_D[c * NS * NS + idx] = p.coef[rec * p.coef_cell_stride + a + b * NS] * (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
            }
        }
        if constexpr (PG) {
            // the modulus records of the tile's Gauss points (ns x ns Mandel matrices, column-major), read in address order, eight
            // loads in flight per thread; the 1/sqrt(2) factors of the shear rows are folded in and the record is stored row-major
            const double hs = 0.70710678118654752440084436210484903928;
            const int nval = ncell_tile * ng * NS * NS;
            for (int t0 = tid; t0 < nval; t0 += 8 * nt) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int t = t0 + u * nt;
                    v[u] = 0.0;
                    if (t < nval) {
                        const int cg = t / (NS * NS), idx = t - cg * (NS * NS);
                        const int c = cg / ng, gp = cg - c * ng;
                        const unsigned long long cell = cell0 + c;
                        const unsigned long long rec = (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.coef_begin;
                        v[u] = __ldcs(p.coef + rec * p.coef_cell_stride + (unsigned long long)gp * p.coef_gp_stride + idx);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int t = t0 + u * nt;
                    if (t < nval) {
                        const int cg = t / (NS * NS), idx = t - cg * (NS * NS);
                        const int b = idx / NS, a = idx - b * NS;
                        s_DG[cg * NS * NS + a * NS + b] = v[u] * ((a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs));
                    }
                }
            }
        }
        // ---- phase B: geometry per (cell, gauss point), SD lanes per pair ------------------------------------------------
        // lane `sub` of a pair accumulates row `sub` of J = Xt.L, the rows are exchanged with shuffles, every lane inverts J
        // (cheap, redundant) and then fills the gradient rows of the nodes m = sub, sub+SD, ..  -- SD times more lanes busy
        // than one thread per pair, which matters for the 20-node kinds whose tiles hold few cells
        if constexpr (NN < 10) {
            // small kinds: one thread per (cell, gauss point) -- tiles hold many cells, every lane has a pair
            for (int task = tid; task < ncell_tile * ng; task += nt) {
                const int gp = task / ncell_tile, c = task - gp * ncell_tile;
                const double *X = s_X + c * XS;
                const double *L = s_L + gp * LS;
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
                double *G = s_G + (c * ng + gp) * GS;
#pragma unroll 4
                for (int m = 0; m < NN; m++) geom::gradient_row<SD>(L + m * SD, Ji, G + 1 + m * SD);
                double coef = det * s_w[gp] * p.alpha;
                if constexpr (AXI) {
                    const double *N = s_N + gp * NN;
                    double r = 0.0;
                    for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], X[m * SD])); // radius_at_ip: r = sum N.x0
                    const double rinv = 1.0 / r;
                    for (int m = 0; m < NN; m++) G[1 + NN * SD + m] = N[m] * rinv;
                    coef *= r;
                }
                G[0] = coef;
            }
        } else {
            constexpr int GPW = 32 / SD; // pairs per warp pass
            const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
            const int grp = lane / SD, sub = lane - grp * SD;
            const int npair = ncell_tile * ng;
            for (int base = wid * GPW; base < npair; base += nw * GPW) {
                const int pair = base + grp;
                const bool valid = grp < GPW && pair < npair;
                const int pr = valid ? pair : 0;
                const int gp = pr / ncell_tile, c = pr - gp * ncell_tile;
                const double *X = s_X + c * XS;
                const double *L = s_L + gp * LS;
                double Jr[SD];
#pragma unroll
                for (int j = 0; j < SD; j++) Jr[j] = 0.0;
#pragma unroll 4
                for (int m = 0; m < NN; m++) {
                    const double x = X[m * SD + sub];
#pragma unroll
                    for (int j = 0; j < SD; j++) Jr[j] = geom::add(Jr[j], geom::mul(x, L[m * SD + j]));
                }
                double J[SD][SD];
#pragma unroll
                for (int i = 0; i < SD; i++)
#pragma unroll
                    for (int j = 0; j < SD; j++) J[i][j] = __shfl_sync(0xffffffffu, Jr[j], (grp * SD + i) & 31);
                double Ji[SD][SD];
                const double det = geom::invert(J, Ji);
                if (valid) {
                    double *G = s_G + (c * ng + gp) * GS;
                    for (int m = sub; m < NN; m += SD) geom::gradient_row<SD>(L + m * SD, Ji, G + 1 + m * SD);
                    double r = 1.0;
                    if constexpr (AXI) {
                        const double *N = s_N + gp * NN;
                        r = 0.0;
                        for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], X[m * SD])); // radius_at_ip: r = sum N.x0
                        const double rinv = 1.0 / r;
                        for (int m = sub; m < NN; m += SD) G[1 + NN * SD + m] = N[m] * rinv;
                    }
                    if (sub == 0) {
                        if (fabs(det) <= ZERO_DETERMINANT) {
                            const unsigned long long cell = cell0 + c;
                            atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
                        }
                        double coef = det * s_w[gp] * p.alpha;
                        if constexpr (AXI) coef *= r;
                        G[0] = coef;
                    }
                }
            }
        }
        __syncthreads();
        if constexpr (PG && SYM) {
            // every staged record against its transpose (the scaling is symmetric, so d' is symmetric iff D is); shared memory only
            constexpr int NP = NS * (NS - 1) / 2;
            bool asym = false;
            for (int t = tid; t < ncell_tile * ng * NP; t += nt) {
                const int cg = t / NP;
                int q = t - cg * NP, a = 0;
                while (q >= NS - 1 - a) {
                    q -= NS - 1 - a;
                    a++;
                }
                const int b = a + 1 + q;
                if (s_DG[cg * NS * NS + a * NS + b] != s_DG[cg * NS * NS + b * NS + a]) asym = true;
            }
            if (asym) *p.pg_flag = 0;
        }

        // ---- phase C: geometric tensors of the node pairs ------------------------------------------------------------------
        // P(m,n)[k][l] = sum_p c_p v(m)[k] v(n)[l] with v(m) = (B(m,0..SD-1) [, N(m)/r]).  D is the same at every Gauss point,
        // so it is applied once, in phase D:  K(m,n)[i][j] = sum_qr d'[row(i,q)][row(j,r)] P(m,n)[comp(i,q)][comp(j,r)].
        // P(n,m) = P(m,n)^T, so the circulant half of the pairs is enough even for a non-symmetric D.
        double acc[NSLOT][VD * VD];
#pragma unroll
        for (int s = 0; s < NSLOT; s++)
#pragma unroll
            for (int q = 0; q < VD * VD; q++) acc[s][q] = 0.0;

        if (worker && cs < ncell_tile) {
#pragma unroll GP_UNROLL
            for (int gp = 0; gp < ng; gp++) {
                const double *G = s_G + (cs * ng + gp) * GS;
                const double c = G[0];
                double vn[VD];
#pragma unroll
                for (int k = 0; k < SD; k++) vn[k] = G[1 + n * SD + k] * c;
                if constexpr (AXI) vn[SD] = G[1 + NN * SD + n] * c;
                double T[NS][SD]; // PG: T(n) = c_p D'_p M(n); rows the strain-displacement columns never touch are dead code
                if constexpr (PG) {
                    const double *Dp = s_DG + (cs * ng + gp) * NS * NS; // the same address for every lane of the cell: broadcast
                    // term by term over ALL entries (q outermost): consecutive FMAs are independent, so a warp does not wait on its own
                    // previous result (ncu: fixed-latency "wait" stalls were a third of the samples with q innermost)
#pragma unroll
                    for (int q = 0; q < 3; q++)
#pragma unroll
                        for (int a = 0; a < NS; a++)
#pragma unroll
                            for (int j = 0; j < SD; j++) {
                                if (q >= col_n(SD, AXI, j)) continue;
                                const double dv = Dp[a * NS + col_row(SD, j, q)], bv = vn[col_comp(SD, j, q)];
                                T[a][j] = q == 0 ? dv * bv : fma(dv, bv, T[a][j]);
                            }
                }
#pragma unroll
                for (int s = 0; s < NSLOT; s++) {
                    // a slot beyond the last distance (only in the last slot of some lanes) recomputes distance 0 and is
                    // dropped in phase D: no divergent branch inside the Gauss loop
                    const int b = LPN == 1 ? s : ((s * LPN + h < ND) ? s * LPN + h : 0);
                    int m = n + b;
                    if (m >= NN) m -= NN;
                    double vm[VD];
#pragma unroll
                    for (int k = 0; k < SD; k++) vm[k] = G[1 + m * SD + k];
                    if constexpr (AXI) vm[SD] = G[1 + NN * SD + m];
                    if constexpr (PG) {
                        // K(m,n)[i][j] += sum_q M(m)[row(i,q)][i] T[row(i,q)][j]   (add_to_kk / add_to_kk_axisymmetric, mat_10_bdb.rs:179-233)
#pragma unroll
                        for (int q = 0; q < 3; q++)
#pragma unroll
                            for (int j = 0; j < SD; j++)
#pragma unroll
                                for (int i = 0; i < SD; i++) {
                                    if (q >= col_n(SD, AXI, i)) continue;
                                    acc[s][i + SD * j] = fma(vm[col_comp(SD, i, q)], T[col_row(SD, i, q)][j], acc[s][i + SD * j]);
                                }
                    } else {
#pragma unroll
                        for (int l = 0; l < VD; l++)
#pragma unroll
                            for (int k = 0; k < VD; k++) acc[s][k + VD * l] = fma(vm[k], vn[l], acc[s][k + VD * l]);
                    }
                }
            }
        }

        // ---- phase D: contract with D', stage and ship ------------------------------------------------------------------
        __syncthreads(); // every lane has finished reading G (it may alias the staging)
        const bool tr = cst.gen == 2;
        auto kidx = [&](int r, int c) -> int { return tr ? c + NDOF * r : r + NDOF * c; };
        if constexpr (PG) {
            // the accumulators are the element blocks: stage them (mirrored in the symmetric variant)
            if (worker && cs < ncell_tile) {
                double *Kc = s_K + cs * NDOF * NDOF;
#pragma unroll
                for (int s = 0; s < NSLOT; s++) {
                    const int b = LPN == 1 ? s : s * LPN + h;
                    if (b >= ND) continue;
                    int m = n + b;
                    if (m >= NN) m -= NN;
                    const bool half = (NN % 2 == 0) && b == NN / 2; // lane n + NN/2 holds the transposed pair itself
                    const bool mirror = SYM && b != 0 && !half;
#pragma unroll
                    for (int j = 0; j < SD; j++)
#pragma unroll
                        for (int i = 0; i < SD; i++) {
                            const double v = acc[s][i + SD * j];
                            Kc[kidx(SD * m + i, SD * n + j)] = v;
                            if (mirror) Kc[kidx(SD * n + j, SD * m + i)] = v;
                        }
                }
            }
        } else if (worker && cs < ncell_tile) {
            double *Kc = s_K + cs * NDOF * NDOF;
            double dl[NS * NS]; // d' of this lane's cell: the constant bank (uniform D) or the tile's record table
#pragma unroll
            for (int a = 0; a < NS; a++)
#pragma unroll
                for (int b = 0; b < NS; b++)
                    if (d_nz(PAT, SD, a, b)) dl[a * NS + b] = drec ? s_D[cs * NS * NS + a * NS + b] : cst.d[a * NS + b];
#pragma unroll
            for (int s = 0; s < NSLOT; s++) {
                const int b = LPN == 1 ? s : s * LPN + h;
                if (b >= ND) continue;
                int m = n + b;
                if (m >= NN) m -= NN;
                const bool half = (NN % 2 == 0) && b == NN / 2; // lane n + NN/2 holds the transposed pair itself
                const bool tri = SYM && (b == 0 || half);
                const bool mirror_all = SYM && !tri;
                const bool transposed = !SYM && b != 0 && !half; // general D: K(n,m) needs its own contraction
#pragma unroll
                for (int j = 0; j < SD; j++)
#pragma unroll
                    for (int i = 0; i < SD; i++) {
                        if (tri && i > j) continue;
                        double v = 0.0, u = 0.0;
                        bool have = false;
#pragma unroll
                        for (int q = 0; q < col_n(SD, AXI, i); q++)
#pragma unroll
                            for (int r = 0; r < col_n(SD, AXI, j); r++) {
                                const int ra = col_row(SD, i, q), rb = col_row(SD, j, r);
                                if (!d_nz(PAT, SD, ra, rb)) continue;
                                const double pv = acc[s][col_comp(SD, i, q) + VD * col_comp(SD, j, r)];
                                v = have ? fma(dl[ra * NS + rb], pv, v) : dl[ra * NS + rb] * pv;
                                if (!SYM) u = have ? fma(dl[rb * NS + ra], pv, u) : dl[rb * NS + ra] * pv;
                                have = true;
                            }
                        Kc[kidx(SD * m + i, SD * n + j)] = v;
                        if (mirror_all || (tri && i < j)) Kc[kidx(SD * n + j, SD * m + i)] = v;
                        if (transposed) Kc[kidx(SD * n + j, SD * m + i)] = u;
                    }
            }
        }
        if (cst.gen == 2) {
            // fused assembly: sweep the tile's (row-major) blocks and add every entry to its slot of the global matrix
            __syncthreads();
            // eight position-map loads in flight per thread before their adds (a RED must not wait on the map)
            const unsigned int total = (unsigned int)ncell_tile * (NDOF * NDOF);
            for (unsigned int idx0 = tid; idx0 < total; idx0 += 8 * nt) {
                uint32_t u[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned int idx = idx0 + k * nt;
                    u[k] = 0xFFFFFFFFu;
                    if (idx < total) {
                        const unsigned int c = idx / (NDOF * NDOF), q = idx - c * (NDOF * NDOF);
                        const unsigned long long cell = cell0 + c;
                        const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
                        const unsigned long long base = p.out_off ? (p.out_off[cell] - p.out_base) : (orig - p.cell_begin) * (unsigned long long)(NDOF * NDOF);
                        u[k] = __ldcs(p.scatter_pos + base + q);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (u[k] != 0xFFFFFFFFu) atomicAdd(p.scatter_vals + u[k], s_K[idx0 + k * nt]);
            }
        } else if (cst.gen) {
            // general blocks (CommonArgs nrow / ncol / ii0 / jj0), clear = false (accumulate into `out`), 8-byte aligned outputs:
            // the tile's blocks are swept in address order; entries outside the integrated extent are zero-filled when clearing
            __syncthreads();
            const unsigned int block = p.nrow * p.ncol;
            for (unsigned int idx = tid; idx < (unsigned int)ncell_tile * block; idx += nt) {
                const unsigned int c = idx / block, q = idx - c * block;
                const unsigned int col = q / p.nrow, row = q - col * p.nrow;
                const unsigned long long cell = cell0 + c;
                const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
                double *dst = p.out + (p.out_off ? (p.out_off[cell] - p.out_base) : (orig - p.cell_begin) * block) + q;
                const bool inside = row >= p.ii0 && row < p.ii0 + NDOF && col >= p.jj0 && col < p.jj0 + NDOF;
                if (!inside) {
                    if (p.clear) *dst = 0.0;
                    continue;
                }
                const double v = s_K[c * NDOF * NDOF + (row - p.ii0) + NDOF * (col - p.jj0)];
                *dst = p.clear ? v : *dst + v;
            }
        } else if (p.cell_ids == nullptr) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                const unsigned int nbytes = (unsigned int)ncell_tile * (unsigned int)(NDOF * NDOF * sizeof(double));
                double *dst = p.out + (cell0 - p.cell_begin) * (unsigned long long)(NDOF * NDOF);
                unsigned long long pol; // the element matrices are written once: keep them from flushing the mesh out of L2
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(s_K)),
                             "r"(nbytes), "l"(pol)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // mixed-kind mesh: every cell has its own destination (prefix of block sizes).  Block sizes of all kinds are even, so the
            // destinations are 16-byte aligned like `out`: one bulk copy per cell.  The lanes of warp 0 look the destinations up in
            // parallel, lane 0 issues the copies (a per-lane issue is serialised by ptxas into a loop that waits on the copy unit)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < 32) {
                unsigned long long pol;
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                for (int c0 = 0; c0 < ncell_tile; c0 += 32) {
                    const int c = c0 + tid;
                    unsigned long long off = 0;
                    if (c < ncell_tile) {
                        const unsigned long long cell = cell0 + c;
                        off = p.out_off ? (p.out_off[cell] - p.out_base)
                                        : ((unsigned long long)p.cell_ids[cell] - p.cell_begin) * (unsigned long long)(NDOF * NDOF);
                    }
                    const int nb = min(32, ncell_tile - c0);
                    for (int k = 0; k < nb; k++) {
                        const unsigned long long o = __shfl_sync(0xffffffffu, off, k);
                        if (tid == 0)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p.out + o),
                                         "r"(smem_u32(s_K + (c0 + k) * NDOF * NDOF)), "r"((unsigned)(NDOF * NDOF * sizeof(double))), "l"(pol)
                                         : "memory");
                    }
                }
                if (tid == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// doubles of the per-(cell, Gauss point) records of one cell: c | B | N/r, plus (PG) the ns x ns modulus of every Gauss point
template <int NN, int SD>
size_t bdb_record_doubles(int ng, bool axi, bool pg) {
    return (size_t)ng * ((1 + NN * SD + (axi ? NN : 0)) | 1) + (pg ? (size_t)ng * 4 * SD * SD : 0);
}

template <int NN, int SD, int LPN>
bool bdb_alias(int ng, int cpc, bool axi, bool pg) {
    return (size_t)cpc * bdb_record_doubles<NN, SD>(ng, axi, pg) + (pg ? 2 : 0) <= (size_t)cpc * NN * SD * NN * SD;
}

template <int NN, int SD, int LPN>
size_t bdb_smem(int ng, int cpc, bool axi, bool drec, bool pg) {
    const int xs = (NN * SD) | 1;
    const int ls = (NN * SD + 6) / 16 * 16 + 9;
    const size_t g = bdb_alias<NN, SD, LPN>(ng, cpc, axi, pg) ? 0 : (size_t)cpc * bdb_record_doubles<NN, SD>(ng, axi, pg) + (pg ? 2 : 0);
    const size_t dsz = drec ? (size_t)cpc * 4 * SD * SD : 0;
    return sizeof(double) * ((size_t)cpc * NN * SD * NN * SD + (size_t)ng * (1 + NN + ls) + (size_t)cpc * xs + dsz + g);
}

__global__ void pg_flag_reset_kernel(int *flag) { *flag = 1; }

template <int NN, int SD, int LPN>
int launch_kind(const IntegParams &p, BdbConsts cst, int pat, void *stream) {
    const bool axi = p.axisym != 0;
    const bool pg = p.coef != nullptr && p.coef_gp_stride != 0;
    // cells per CTA: fill ~256 threads, at most ~110 KB of shared memory (two CTAs per SM)
    int cpc = 256 / (NN * LPN);
    // measured (cells per CTA swept, `profiles/r01_bdb_cells_per_cta_sweep.txt`): small CTAs -- more of them resident, cheaper
    // barriers -- win where the element block is large against the arithmetic per lane
    if (NN == 10 && SD == 3) cpc = 4;  // Tet10: 266 -> 301 M el/s
    if (NN >= 8 && NN <= 9 && SD == 2) cpc = 16; // Qua8: 1.87 -> 1.99 G el/s, Qua9: 1.12 -> 1.22 G el/s
    if (NN == 4 && SD == 2 && axi) cpc = 8; // axisymmetric Qua4: 4.63 (8), 4.37 (16), 4.11 (32), 3.68 (64) G el/s
    if (const char *e = tuning_env("GL_BDB_CPC")) cpc = atoi(e); // tuning knob: cells per CTA
    if (cpc * NN * LPN > 256) cpc = 256 / (NN * LPN); // __launch_bounds__(256)
    if (cpc < 1) cpc = 1;
    const bool drec = p.coef != nullptr && !pg;
    while (cpc > 1 && bdb_smem<NN, SD, LPN>(p.ng, cpc, axi, drec, pg) > 110 * 1024) cpc--;
    const size_t smem = bdb_smem<NN, SD, LPN>(p.ng, cpc, axi, drec, pg);
    if (smem > 200 * 1024) return 0;
    const int used = cpc * NN * LPN;
    const int threads = (used + 31) / 32 * 32;
    typedef void (*kern_t)(const IntegParams, const BdbConsts, const int, const int, const int);
    auto pick = [&](int pat) -> kern_t {
        if (pg) {
            if constexpr (SD == 2) {
                if (axi) return pat >= 1 ? bdb_kernel<NN, SD, LPN, true, 1, true> : bdb_kernel<NN, SD, LPN, true, 0, true>;
            }
            return pat >= 1 ? bdb_kernel<NN, SD, LPN, false, 1, true> : bdb_kernel<NN, SD, LPN, false, 0, true>;
        }
        if constexpr (NN == 20 && LPN == 2) {
            return nullptr; // this lane split is instantiated for the per-Gauss-point variants only
        } else {
            if constexpr (SD == 2) {
                if (axi) return pat == 2 ? bdb_kernel<NN, SD, LPN, true, 2> : (pat == 1 ? bdb_kernel<NN, SD, LPN, true, 1> : bdb_kernel<NN, SD, LPN, true, 0>);
            }
            return pat == 2 ? bdb_kernel<NN, SD, LPN, false, 2> : (pat == 1 ? bdb_kernel<NN, SD, LPN, false, 1> : bdb_kernel<NN, SD, LPN, false, 0>);
        }
    };
    if (NN == 20 && LPN == 2 && !pg) return 0;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    if (ctas_per_sm > 2048 / threads) ctas_per_sm = 2048 / threads;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long ntile = (p.hi - p.lo + cpc - 1) / cpc;
    // many more CTAs than resident slots: CTAs that start and finish at different times keep the SMs out of lockstep, so
    // reads and writes of different SMs interleave instead of arriving in device-wide bursts (see gl_kernels_hex8.cu)
    unsigned long long grid = (unsigned long long)nsm * ctas_per_sm * 16;
    if (const char *e = tuning_env("GL_BDB_GRID")) grid = (unsigned long long)atoi(e); // tuning knob
    if (grid > ntile) grid = ntile;
    if (grid == 0) return 0;
    const int alias = bdb_alias<NN, SD, LPN>(p.ng, cpc, axi, pg) ? 1 : 0;
    if (pg) {
        // symmetric variant first (it checks the records it stages), the general one behind it runs only if a record was not symmetric
        pg_flag_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p.pg_flag);
        int first = 1;
        if (const char *e = tuning_env("GL_BDB_PATTERN")) first = atoi(e) >= 1 ? 1 : 0; // testing knob: 0 = general variant only
        if (first == 0) cudaMemsetAsync(p.pg_flag, 0, sizeof(int), (cudaStream_t)stream);
        for (int pt = first; pt >= 0; pt--) {
            kern_t kern = pick(pt);
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
            kern<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(p, cst, cpc, used, alias);
        }
        if (cudaGetLastError() != cudaSuccess) return -1;
        note_kernel("bdb<nn=%d,sd=%d,axi=%d,gauss%s>", NN, SD, axi ? 1 : 0, cst.gen == 2 ? ",scatter" : (cst.gen ? ",gen" : ""));
        return 2 + first;
    }
    // records whose structure only the device knows: every variant is enqueued, the probe's verdict selects the one that runs
    const bool probe = drec && p.coef_pat_dev != nullptr && tuning_env("GL_BDB_PATTERN") == nullptr;
    IntegParams q = p;
    if (!probe) q.coef_pat_dev = nullptr;
    for (int pt = probe ? 0 : pat; pt <= (probe ? 2 : pat); pt++) {
        kern_t kern = pick(pt);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(q, cst, cpc, used, alias);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("bdb<nn=%d,sd=%d,axi=%d,pat=%d%s%s>", NN, SD, axi ? 1 : 0, pat, drec ? ",rec" : "", cst.gen == 2 ? ",scatter" : (cst.gen ? ",gen" : ""));
    return probe ? 3 : 1;
}

} // namespace

// This file is compiled twice so that the 2D and the 3D instantiations build in parallel: on its own it is the 3D translation
// unit (plus the dimension dispatch); gl_kernels_bdb2d.cu defines GL_BDB_SD = 2 and includes it.
#ifndef GL_BDB_SD
#define GL_BDB_SD 3
#define GL_BDB_MAIN
#endif
#if GL_BDB_SD == 2
int launch_bdb_2d(const IntegParams &p, void *stream) {
    const int sd = 2;
#else
int launch_bdb_3d(const IntegParams &p, void *stream) {
    const int sd = 3;
#endif
    // covered: D uniform, one record per marker / per cell, or one per Gauss point; homogeneous or mixed layout; tight blocks
    // leave by bulk copy, everything else (nrow / ncol / ii0 / jj0, clear = false, 8-byte aligned output) by the general epilogue
    if (p.op != GL_OP_MAT_10_BDB) return 0;
    const bool pg = p.coef != nullptr && p.coef_gp_stride != 0;
    if (p.coef != nullptr && !pg && p.coef_cell_stride != (unsigned long long)(4 * sd * sd)) return 0;
    if (pg && (p.coef_gp_stride != (unsigned long long)(4 * sd * sd) || p.coef_index != nullptr || p.pg_flag == nullptr)) return 0;
    const unsigned ndof = (unsigned)p.nnode * (unsigned)sd;
    if (p.size_r != ndof || p.size_c != ndof) return 0;
    const bool tight = p.nrow == ndof && p.ncol == ndof && p.ii0 == 0 && p.jj0 == 0;
    if (tuning_env("GL_DISABLE_BDB")) return 0; // testing knob: force the generic kernel

    const int ns = 2 * sd;
    BdbConsts cst;
    cst.gen = !(tight && p.clear && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) ? 1 : 0;
    if (p.scatter_pos != nullptr) {
        if (!tight) return 0;
        cst.gen = 2;
    }
    const double hs = 0.70710678118654752440084436210484903928;
    bool sym = true;
    for (int a = 0; a < ns; a++)
        for (int b = 0; b < ns; b++) {
            const double fa = a < 3 ? 1.0 : hs, fb = b < 3 ? 1.0 : hs;
            cst.d[a * ns + b] = p.coef_uniform[a + b * ns] * fa * fb;
            if (p.coef_uniform[a + b * ns] != p.coef_uniform[b + a * ns]) sym = false;
        }
    int pat = 0;
    if (sym) {
        pat = 2;
        for (int a = 0; a < ns; a++)
            for (int b = 0; b < ns; b++)
                if (!d_nz(2, sd, a, b) && cst.d[a * ns + b] != 0.0) pat = 1;
    }
    if (p.coef != nullptr) pat = p.coef_pat; // structure common to all records (0 when they live on the device only)
    if (const char *e = tuning_env("GL_BDB_PATTERN")) {
        const int cap = atoi(e);
        if (cap < pat) pat = cap < 0 ? 0 : cap;
    }
#if GL_BDB_SD == 2
    switch (p.nnode) {
    case 3: return launch_kind<3, 2, 1>(p, cst, pat, stream);
    case 4: return launch_kind<4, 2, 1>(p, cst, pat, stream);
    case 6: return launch_kind<6, 2, 1>(p, cst, pat, stream);
    case 8: return launch_kind<8, 2, 1>(p, cst, pat, stream);
    case 9: return launch_kind<9, 2, 1>(p, cst, pat, stream);
    default: return 0;
    }
#else
    switch (p.nnode) {
    case 4: return launch_kind<4, 3, 1>(p, cst, pat, stream);
    case 8: return launch_kind<8, 3, 1>(p, cst, pat, stream);
    case 10: return launch_kind<10, 3, 1>(p, cst, pat, stream);
    case 20:
        // one modulus per Gauss point: every lane of a column node recomputes T(n) = D'_p M(n), so fewer, fatter lanes win
        // (2 lanes x 6 blocks: 216 DFMA per lane and Gauss point, 8,640 per cell; 4 lanes x 3 blocks: 10,800 per cell)
        if (pg && !tuning_env("GL_BDB_PG_LPN4")) return launch_kind<20, 3, 2>(p, cst, pat, stream);
        return launch_kind<20, 3, 4>(p, cst, pat, stream); // 11 circulant blocks per column node split over 4 lanes
    default: return 0;
    }
#endif
}

#if GL_BDB_SD == 2
bool bdb_covers_2d(int nnode, int ng, bool axi, bool drec, bool pg) {
    switch (nnode) {
    case 3: return bdb_smem<3, 2, 1>(ng, 1, axi, drec, pg) <= 200 * 1024;
    case 4: return bdb_smem<4, 2, 1>(ng, 1, axi, drec, pg) <= 200 * 1024;
    case 6: return bdb_smem<6, 2, 1>(ng, 1, axi, drec, pg) <= 200 * 1024;
    case 8: return bdb_smem<8, 2, 1>(ng, 1, axi, drec, pg) <= 200 * 1024;
    case 9: return bdb_smem<9, 2, 1>(ng, 1, axi, drec, pg) <= 200 * 1024;
    default: return false;
    }
}
#else
bool bdb_covers_3d(int nnode, int ng, bool drec, bool pg) {
    switch (nnode) {
    case 4: return bdb_smem<4, 3, 1>(ng, 1, false, drec, pg) <= 200 * 1024;
    case 8: return bdb_smem<8, 3, 1>(ng, 1, false, drec, pg) <= 200 * 1024;
    case 10: return bdb_smem<10, 3, 1>(ng, 1, false, drec, pg) <= 200 * 1024;
    case 20: return bdb_smem<20, 3, 4>(ng, 1, false, drec, pg) <= 200 * 1024;
    default: return false;
    }
}
#endif

#ifdef GL_BDB_MAIN
bool bdb_covers_2d(int nnode, int ng, bool axi, bool drec, bool pg);
bool bdb_covers(int nnode, int sd, int ng, bool axisym, bool record_moduli, bool per_gauss) {
    return sd == 2 ? bdb_covers_2d(nnode, ng, axisym, record_moduli, per_gauss) : (!axisym && bdb_covers_3d(nnode, ng, record_moduli, per_gauss));
}
int launch_bdb_2d(const IntegParams &p, void *stream);
int launch_bdb(const IntegParams &p, int sd, void *stream) { return sd == 2 ? launch_bdb_2d(p, stream) : launch_bdb_3d(p, stream); }
#endif

} // namespace gl
