// gl_kernels_qua8.cu -- Qua8 and Tri6 stiffness on the FP64 TENSOR CORES, plane and axisymmetric: integ::mat_10_bdb
// (src/integ/mat_10_bdb.rs:121-233, the axisymmetric branch :213-233) with a modulus D that does not vary inside a cell (uniform, or
// one record per marker / per cell).  Qua8 is the heavier half of BASELINE.json configs[4] (mixed Tri3 / Qua8, axisymmetric).
//
// Algebra (gl_kernels_bdb.cu): with the per-node vector g(m) = (B(m,0), B(m,1) [, N(m)/r]) and c' = det.w.alpha [.r],
//     P[k][l](m,n) = sum_p c'_p g_p(m)[k] g_p(n)[l],        K(2m+i, 2n+j) = sum_kl D'_(ik)(jl) P[k][l](m,n)   (once per node pair).
// The register-blocked kernel accumulates P with DFMA from per-node records in shared memory and is bound by the shared-memory
// pipe (ncu, profiles/r02_ncu_full_qua8_axisym_bdb_summary.json: LDS wavefronts 78 % of peak, FP64 pipe 35 %, a quarter of the
// stall samples at CTA barriers).  Here P = V^T diag(c') V is formed by mma.sync.m8n8k4.f64 with the columns of V ordered
// COMPONENT-MAJOR, V[p][NS k + m] = g_p(m)[k]: Qua8 has exactly eight nodes, so the 8 x 8 tile (I, J) of P is P[I][J](., .) and its
// accumulator fragment gives lane (ar, ak) the entries (m, n) = (2ak + e, ar), e = 0, 1, of EVERY component pair -- the lane owns
// the complete geometric tensors of two node pairs and contracts them with D' in registers (compile-time sparsity pattern); no
// second trip through shared memory, no pair table.  Only the upper tiles (I <= J) are computed (6 of 9 axisymmetric, 3 of 4
// plane); P[J][I](m,n) = P[I][J](n,m) sits in lane (2ak + e, ar >> 1) as element ar & 1 and is fetched with shuffles.
//
// Mapping: ONE WARP PER G CELLS (three Qua8, four Tri6), no barrier wider than a warp, 8 warps per CTA, 2 CTAs per SM.  Per group
// ("triple" below) of cells:
//   A  gather (Mesh::set_pad, src/mesh/mesh.rs:448-456): lane = (cell, node); point ids two groups ahead, coordinates one;
//   B  geometry, lane = (cell, Gauss point), in the ORACLE'S operation order (gl_geom.cuh): J = Xt.L, closed-form inverse, det,
//      r = sum N.x0; B = L.inv(J), N/r -> V (shared memory, dof-major, conflict-free strides), c';
//   C  per cell: k-steps of 6 (3) DMMA over four Gauss points each (a single leftover point by DFMA), the scaled operand c' V formed
//      in registers; transposed tiles by shuffle; contraction with D'; the lane's two 2 x 2 blocks go to the staged element block
//      as 16-byte stores (odd fragment rows store their upper half first: conflict-free for Qua8);
//   D  the group (G contiguous blocks on single-kind meshes; runs of consecutive blocks on mixed ones) leaves with bulk async copies
//      (L2 evict-first), issued by the lane that owns the first cell of a run.
// Tri6 runs on the same template with its six nodes in the first six node slots of the tiles.  gl_integ_fused can ask for the
// conductivity matrix and the flux vector of the same cells as by-products (Qua8Extra).
// Covered: uniform D or one record per marker / per cell (any pattern), rules of at most 10 (Tri6: 8) points; tight blocks that
// overwrite a 16-byte aligned buffer leave by bulk copy, everything else (nrow / ncol / ii0 / jj0, clear = false) through the general
// store epilogue; the fused assembly (scatter epilogue) with a symmetric modulus pattern.  One modulus per Gauss point and the fused
// assembly with a general D stay on gl_kernels_bdb.cu.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

// strain-displacement column (node, j) in 2D: entries q -> (Mandel row, component of g); component 2 is the hoop value N/r
// (calc_bb_matrix, src/integ/mat_10_bdb.rs:238-278)
__host__ __device__ constexpr int q8_col_n(bool axi, int j) { return j == 0 ? (axi ? 3 : 2) : 2; }
__host__ __device__ constexpr int q8_col_row(int j, int q) { return j == 0 ? (q == 0 ? 0 : (q == 1 ? 3 : 2)) : (q == 0 ? 1 : 3); }
__host__ __device__ constexpr int q8_col_comp(int j, int q) { return j == 0 ? (q == 0 ? 0 : (q == 1 ? 1 : 2)) : (q == 0 ? 1 : 0); }
// PAT 0 general, PAT 2 symmetric with uncoupled normal / diagonal shear blocks (exact zeros)
__host__ __device__ constexpr bool q8_nz(int pat, int a, int b) { return pat < 2 ? true : ((a < 3 && b < 3) || a == b); }

struct Qua8Consts {
    double d[16]; // d'[a][b] = f_a f_b d[a][b] (row-major 4 x 4), f = 1 for the normal rows, 1/sqrt(2) for the shear row
};

// by-products of the same geometric tensors (gl_integ_fused: stiffness + conductivity matrix + flux vector, BASELINE.json configs[4]):
//   mat_03_btb (src/integ/mat_03_btb.rs:50-131)  Kc(m,n) = t0 P[0][0] + t1 P[1][1] + t3/sqrt(2) (P[0][1] + P[1][0])   (same c', same alpha)
//   vec_03_bv  (src/integ/vec_03_bv.rs)          b(m)    = sum_p c'_p (w0 B_p(m,0) + w1 B_p(m,1))
struct Qua8Extra {
    int mask; // 2: mat_03_btb, 8: vec_03_bv (the bits of ScalarFused::mask)
    double t00, t11, t01, w0, w1;
    double *out_m03, *out_v03;
    const unsigned long long *moff, *voff; // mixed meshes: bucket-local cell -> absolute offset of its NN x NN block / NN-vector
    unsigned long long mbase, vbase;
};

__device__ __forceinline__ unsigned q8_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void q8_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double q8_select(int c, double a, double b) { // c ? a : b
    double r;
    asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"(c));
    return r;
}

constexpr int Q8_CTAS = 2;     // CTAs per SM
constexpr int Q8_PS = 12;      // V is stored dof-major, Vt[col][p] at col PS + p; PS = 12 (mod 16): the fragment loads of a half-warp
                               // (4 columns x 4 Gauss points) hit 16 distinct bank pairs
constexpr int Q8_XS = 18;      // coordinates of a cell (up to 16 doubles, + 2): the cells of a warp read distinct banks

// NN = 8 (Qua8) or 6 (Tri6: its six nodes occupy the first six of the eight node slots of a tile; what the fragment lanes of the two
// other slots read only reaches entries of P that nobody reads -- an MMA output depends on its own row and column alone)
template <int NN, bool AXI, bool DREC = false>
struct Qua8Dims {
    static constexpr int NDOF = 2 * NN;
    static constexpr int KB = NDOF * NDOF;    // element block (doubles)
    // cells per warp: G x ng geometry lanes (27 of 32 with the 9-point Qua8 rule, 28 with the 7-point Tri6 rule)
    static constexpr int G = NN == 6 ? 4 : 3;
    static constexpr int VD = AXI ? 3 : 2;
    static constexpr int NS = NN;             // columns of V per component: V[p][NS k + m].  The fragment lanes of the node slots a Tri6
                                              // tile does not have (ar = 6, 7) read the first columns of the next component instead
    static constexpr int NV = NS * VD;
    // cell stride = ng (mod 16) for the default rule (9 points Qua8, 7 points Tri6): the (cell, Gauss point) lanes of a half-warp
    // store one column conflict-free
    static constexpr int CS = NV * Q8_PS + ((NN == 8 ? 9 : 7) + 16 - (NV * Q8_PS) % 16) % 16;
    static constexpr int VSZ = (G * CS + 1) & ~1;
    static constexpr int WSZ = G * KB + VSZ + G * Q8_PS + G * Q8_XS + (DREC ? G * 16 : 0); // doubles per warp (even)
    // warps per CTA: eight if two CTAs of eight fit an SM (228 KB, 1 KB reserved per CTA), else seven
    static constexpr int WARPS = 2 * (8 * WSZ * 8 + 1024) <= 228 * 1024 ? 8 : 7;
    static_assert(NN == 6 || NN == 8, "node slots per tile");
    static_assert(WSZ % 2 == 0, "16-byte aligned warp regions");
};

// DREC = one modulus record per marker or per cell (multi-material meshes) instead of a uniform D: the records of the next triple are
// fetched one iteration ahead (their indices two ahead), scaled into d' and staged per cell; with device-resident records whose
// structure only the probe knows, every pattern variant is enqueued and all but the matching one return at once (gl_kernels_bdb.cu)
// SCATTER = fused assembly (gl_assemble_planned): instead of leaving as element blocks, every staged entry is added to its slot of the
// global matrix through the plan's position map (IntegParams::scatter_pos / scatter_vals, RED.ADD.F64); the element-matrix slab is
// never written.  Only with a symmetric pattern (PAT 2): the map is indexed row-major, the staged block is column-major, K = K^T.
// GEN = general store epilogue: element blocks of nrow x ncol with the integrated window at (ii0, jj0), clear = false, 8-byte aligned
// outputs (CommonArgs, src/integ/common_args.rs:5-43; clear = true overwrites the WHOLE block, mat_10_bdb.rs:141-143) -- e.g. the
// displacement block of a coupled displacement-pressure element matrix.
template <int NN, bool AXI, int PAT, bool EXTRA, bool DREC = false, bool SCATTER = false, bool GEN = false>
__global__ void __launch_bounds__(32 * Qua8Dims<NN, AXI, DREC>::WARPS, Q8_CTAS) qua8_kernel(const IntegParams p, const Qua8Consts cst, const Qua8Extra x) {
    if (DREC && p.coef_pat_dev != nullptr && (*p.coef_pat_dev == 2 ? 2 : 0) != PAT) return;
    using T = Qua8Dims<NN, AXI, DREC>;
    constexpr int VD = T::VD, CS = T::CS, PS = Q8_PS, G = T::G, XS = Q8_XS, KB = T::KB, NDOF = T::NDOF, WARPS = T::WARPS, NS = T::NS;

    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *Kst = smem + warp * T::WSZ; // [G][KB]   staged element blocks (16-byte aligned: bulk copy source)
    double *V = Kst + G * KB;           // [G][CS]   Vt[cell][NS k + m][p]
    double *cc = V + T::VSZ;            // [G][PS]   c'_p
    double *Xs = cc + G * PS;           // [G][XS]   X[2m + j]
    double *Dc = Xs + G * XS;           // [G][16]   d' of the cell, row-major (DREC only)

    const int ng = p.ng;
    // Gauss points go through the tensor cores four at a time; a single leftover point (9 = 2 x 4 + 1, the default rule) is a rank-1
    // update of the fragments with DFMA (12 + 3 instructions instead of a padded k-step of 6 DMMA = 48 DFMA slots of the same pipe)
    const bool tail1 = (ng & 3) == 1;
    const int KS = tail1 ? ng >> 2 : (ng + 3) >> 2;
    const int ar = lane >> 2, ak = lane & 3;

    // phase B role: lane = (cell of the triple, Gauss point); its rows of the reference-element table stay in registers
    const bool vl = lane < G * ng;
    const int bg = vl ? lane / ng : 0, bp = vl ? lane - bg * ng : 0;
    double l[2 * NN], nsh[NN];
#pragma unroll
    for (int i = 0; i < 2 * NN; i++) l[i] = p.rule[ng * (1 + NN) + bp * 2 * NN + i];
#pragma unroll
    for (int m = 0; m < NN; m++) nsh[m] = AXI ? p.rule[ng + bp * NN + m] : 0.0;
    const double wgt = p.rule[bp];

    // phase A role: lane = (cell of the triple, node)
    const int ag = lane / NN, am = lane - ag * NN;
    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long ntriple = (ncell + G - 1) / G;
    const unsigned long long tstride = (unsigned long long)gridDim.x * WARPS;
    auto load_pid = [&](unsigned long long t) -> unsigned int {
        const unsigned long long c = t * G + ag;
        return (lane < NN * G && t < ntriple && c < ncell) ? p.conn[(unsigned long long)am * p.ncell_k + p.lo + c] : 0u;
    };
    unsigned long long t = (unsigned long long)blockIdx.x * WARPS + warp;
    unsigned int pid = load_pid(t);
    double gx0 = 0.0, gx1 = 0.0;
    if (lane < NN * G) {
        gx0 = p.coords[pid];
        gx1 = p.coords[p.npoint + pid];
    }
    pid = load_pid(t + tstride);
    // modulus records: lane -> entry (lane & 15) of cells (lane >> 4) and 2 + (lane >> 4) of the warp's cells
    auto load_rec = [&](unsigned long long tt, int g) -> unsigned long long {
        const unsigned long long c = tt * G + g;
        if (!(g < G && tt < ntriple && c < ncell)) return 0ull;
        const unsigned long long cell = p.lo + c;
        return p.coef_index ? (unsigned long long)p.coef_index[cell] : (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.coef_begin;
    };
    unsigned long long rec0 = 0, rec1 = 0;
    double dv0 = 0.0, dv1 = 0.0;
    if constexpr (DREC) {
        rec0 = load_rec(t, lane >> 4);
        rec1 = load_rec(t, 2 + (lane >> 4));
        dv0 = p.coef[rec0 * p.coef_cell_stride + (lane & 15)];
        dv1 = p.coef[rec1 * p.coef_cell_stride + (lane & 15)];
        rec0 = load_rec(t + tstride, lane >> 4);
        rec1 = load_rec(t + tstride, 2 + (lane >> 4));
    }

    unsigned long long pol; // the element matrices are written once: keep them from flushing the mesh out of L2
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

    for (; t < ntriple; t += tstride) {
        const unsigned long long c0 = t * G; // first cell of the triple, relative to p.lo
        const int nv = (int)(ncell - c0 < (unsigned long long)G ? ncell - c0 : (unsigned long long)G);
        // ---- phase A: destinations, publish the coordinates, prefetch ------------------------------------------------------
        // destination of the lane's cell (lanes 0 .. nv-1), needed at the very end; computed BEFORE the prefetch loads are issued (the
        // address arithmetic of a conditional offset load would otherwise wait on the scoreboard it shares with them)
        unsigned long long off = 0, moffv = 0, voffv = 0;
        if (lane < nv) {
            const unsigned long long cell = p.lo + c0 + lane;
            const unsigned long long rel = (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.cell_begin;
            off = p.out_off ? p.out_off[cell] - p.out_base : rel * (GEN ? (unsigned long long)p.nrow * p.ncol : (unsigned long long)KB);
            if constexpr (EXTRA) {
                moffv = x.moff ? x.moff[cell] - x.mbase : rel * (unsigned long long)(NN * NN);
                voffv = x.voff ? x.voff[cell] - x.vbase : rel * (unsigned long long)NN;
            }
        }
        if (lane < NN * G) {
            *reinterpret_cast<double2 *>(Xs + ag * XS + 2 * am) = make_double2(gx0, gx1);
            gx0 = p.coords[pid];
            gx1 = p.coords[p.npoint + pid];
        }
        pid = load_pid(t + 2 * tstride);
        if constexpr (DREC) {
            // record entry a + 4b (column-major ns x ns Mandel matrix) -> d'[a][b], the 1/sqrt(2) of the shear row folded in
            const double hs = 0.70710678118654752440084436210484903928;
            const int a = lane & 3, b = (lane >> 2) & 3;
            const double fs = (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
            Dc[(lane >> 4) * 16 + a * 4 + b] = dv0 * fs;
            if (2 + (lane >> 4) < G) Dc[(2 + (lane >> 4)) * 16 + a * 4 + b] = dv1 * fs;
            dv0 = p.coef[rec0 * p.coef_cell_stride + (lane & 15)];
            dv1 = p.coef[rec1 * p.coef_cell_stride + (lane & 15)];
            rec0 = load_rec(t + 2 * tstride, lane >> 4);
            rec1 = load_rec(t + 2 * tstride, 2 + (lane >> 4));
        }
        __syncwarp();

        // ---- phase B: geometry, lane = (cell, Gauss point) ----------------------------------------------------------------
        {
            const double *X = Xs + bg * XS;
            double J[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            double r = 0.0;
#pragma unroll
            for (int m = 0; m < NN; m++) {
                const double2 xx = *reinterpret_cast<const double2 *>(X + 2 * m);
                J[0][0] = geom::add(J[0][0], geom::mul(xx.x, l[2 * m]));
                J[0][1] = geom::add(J[0][1], geom::mul(xx.x, l[2 * m + 1]));
                J[1][0] = geom::add(J[1][0], geom::mul(xx.y, l[2 * m]));
                J[1][1] = geom::add(J[1][1], geom::mul(xx.y, l[2 * m + 1]));
                if constexpr (AXI) r = geom::add(r, geom::mul(nsh[m], xx.x)); // radius_at_ip: r = sum N.x0
            }
            double Ji[2][2];
            const double det = geom::invert(J, Ji);
            const bool live = vl && bg < nv;
            if (live && fabs(det) <= ZERO_DETERMINANT) {
                const unsigned long long cell = p.lo + c0 + bg;
                atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
            }
            double coef = det * wgt * p.alpha;
            double rinv = 0.0;
            if constexpr (AXI) {
                rinv = 1.0 / r;
                coef *= r;
            }
            if (vl) {
                double *Vq = V + bg * CS + bp;
#pragma unroll
                for (int m = 0; m < NN; m++) {
                    Vq[m * PS] = fma(l[2 * m + 1], Ji[1][0], l[2 * m] * Ji[0][0]);
                    Vq[(NS + m) * PS] = fma(l[2 * m + 1], Ji[1][1], l[2 * m] * Ji[0][1]);
                    if constexpr (AXI) Vq[(2 * NS + m) * PS] = nsh[m] * rinv;
                }
                cc[bg * PS + bp] = coef;
            }
            if (lane < G) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the previous bulk copies have left Kst
        }
        __syncwarp();

        if constexpr (EXTRA) {
            // flux vector: lane = (cell, node) sums its column pair of V over the Gauss points
            const unsigned long long vo = __shfl_sync(0xffffffffu, voffv, ag < G ? ag : 0);
            if ((x.mask & 8) && lane < NN * G && ag < nv) {
                const double *v0 = V + ag * CS + am * PS;
                double sum = 0.0;
                for (int q = 0; q < ng; q++) sum = fma(cc[ag * PS + q], fma(x.w1, v0[NS * PS + q], x.w0 * v0[q]), sum);
                x.out_v03[vo + am] = sum;
            }
        }

        // ---- phase C: per cell, P on the tensor cores, contraction with D', staging ------------------------------------------
#pragma unroll
        for (int g = 0; g < G; g++) {
            double P[VD][VD][2];
#pragma unroll
            for (int I = 0; I < VD; I++)
#pragma unroll
                for (int Jt = 0; Jt < VD; Jt++) P[I][Jt][0] = P[I][Jt][1] = 0.0;
#pragma unroll 1
            for (int ks = 0; ks < KS; ks++) {
                const int row = 4 * ks + ak;
                const bool ok = row < ng; // the padding Gauss points of the last k-step contribute zeros
                const double c = ok ? cc[g * PS + row] : 0.0;
                const double *vr = V + g * CS + ar * PS + row;
                double v[VD], vs[VD];
#pragma unroll
                for (int k = 0; k < VD; k++) {
                    v[k] = ok ? vr[NS * k * PS] : 0.0;
                    vs[k] = v[k] * c;
                }
                // C[r][q] += sum_k A[r][k] B[k][q] with A[r][k] = V[k][8J + r], B[k][q] = c_k V[k][8I + q]: C[r][q] = P[I][J](m = q, n = r)
#pragma unroll
                for (int I = 0; I < VD; I++)
#pragma unroll
                    for (int Jt = I; Jt < VD; Jt++) q8_dmma(P[I][Jt][0], P[I][Jt][1], v[Jt], vs[I]);
            }
            if (tail1) {
                const double *vq = V + g * CS + (ng - 1);
                const double c = cc[g * PS + ng - 1];
                double gm[VD][2], gn[VD];
#pragma unroll
                for (int k = 0; k < VD; k++) {
                    gn[k] = vq[(NS * k + ar) * PS] * c;
                    gm[k][0] = vq[(NS * k + 2 * ak) * PS];
                    gm[k][1] = vq[(NS * k + 2 * ak + 1) * PS];
                }
#pragma unroll
                for (int I = 0; I < VD; I++)
#pragma unroll
                    for (int Jt = I; Jt < VD; Jt++) {
                        P[I][Jt][0] = fma(gm[I][0], gn[Jt], P[I][Jt][0]);
                        P[I][Jt][1] = fma(gm[I][1], gn[Jt], P[I][Jt][1]);
                    }
            }
            // the lane holds P[I][J](2ak + e, ar), I <= J.  P[J][I](2ak + e, ar) = P[I][J](ar, 2ak + e): lane (2ak + e, ar >> 1), element ar & 1
#pragma unroll
            for (int I = 0; I < VD; I++)
#pragma unroll
                for (int Jt = I + 1; Jt < VD; Jt++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int src = (2 * ak + e) * 4 + (ar >> 1);
                        const double t0 = __shfl_sync(0xffffffffu, P[I][Jt][0], src);
                        const double t1 = __shfl_sync(0xffffffffu, P[I][Jt][1], src);
                        P[Jt][I][e] = q8_select(ar & 1, t1, t0);
                    }
            if constexpr (EXTRA) {
                // conductivity block, column-major NN x NN: rows 2ak, 2ak + 1 of column ar.  8-byte stores: on mixed meshes the 3 x 3 blocks
                // of the Tri3 cells can leave the offset odd, and one 16-byte store where it is even measured no faster
                const unsigned long long mo = __shfl_sync(0xffffffffu, moffv, g);
                if ((x.mask & 2) && g < nv && (NN == 8 || (ar < NN && 2 * ak < NN))) {
                    double *dst = x.out_m03 + mo + NN * ar + 2 * ak;
#pragma unroll
                    for (int e = 0; e < 2; e++) dst[e] = fma(x.t11, P[1][1][e], fma(x.t01, P[0][1][e] + P[1][0][e], x.t00 * P[0][0][e]));
                }
            }
            // K(2m + i, 2n + j) = sum_qr d'[row(i,q)][row(j,r)] P[comp(i,q)][comp(j,r)](m,n)   (add_to_kk / add_to_kk_axisymmetric)
            double Kb[2][2][2]; // [e][i][j]
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        double s = 0.0;
                        bool have = false;
#pragma unroll
                        for (int q = 0; q < q8_col_n(AXI, i); q++)
#pragma unroll
                            for (int rr = 0; rr < q8_col_n(AXI, j); rr++) {
                                const int ra = q8_col_row(i, q), rb = q8_col_row(j, rr);
                                if (!q8_nz(PAT, ra, rb)) continue;
                                const double pv = P[q8_col_comp(i, q)][q8_col_comp(j, rr)][e];
                                const double dab = DREC ? Dc[g * 16 + ra * 4 + rb] : cst.d[ra * 4 + rb];
                                s = have ? fma(dab, pv, s) : dab * pv;
                                have = true;
                            }
                        Kb[e][i][j] = s;
                    }
            // rows 4ak + 2e + i of columns 2ar + j, column-major NDOF x NDOF; odd fragment rows store e = 1 first: with NDOF = 16 the eight
            // lanes of a quarter-warp then write eight distinct 16-byte slots of the 128-byte bank window.  (Tri6: the lanes of the two
            // padding node slots hold nothing.)
            if (NN == 8 || (ar < NN && 2 * ak < NN)) {
                double *Kc = Kst + g * KB + NDOF * (2 * ar) + 4 * ak;
                const int e0 = ar & 1;
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const int e = e0 ^ s;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const double a = q8_select(e, Kb[1][0][j], Kb[0][0][j]);
                        const double b = q8_select(e, Kb[1][1][j], Kb[0][1][j]);
                        *reinterpret_cast<double2 *>(Kc + NDOF * j + 2 * e) = make_double2(a, b);
                    }
                }
            }
        }

        // ---- phase D: ship ---------------------------------------------------------------------------------------------------
        if constexpr (SCATTER) {
            static_assert(!SCATTER || PAT == 2, "the scatter epilogue reads the staged block through its transpose");
            __syncwarp();
#pragma unroll
            for (int g = 0; g < G; g++) {
                const unsigned long long base = __shfl_sync(0xffffffffu, off, g);
                if (g < nv) {
                    constexpr int NIT = (KB + 31) / 32;
                    unsigned int pos[NIT];
#pragma unroll
                    for (int it = 0; it < NIT; it++) pos[it] = lane + 32 * it < KB ? p.scatter_pos[base + lane + 32 * it] : 0xffffffffu;
#pragma unroll
                    for (int it = 0; it < NIT; it++)
                        if (pos[it] != 0xffffffffu) atomicAdd(p.scatter_vals + pos[it], Kst[g * KB + lane + 32 * it]);
                }
            }
            __syncwarp();
            continue;
        }
        if constexpr (GEN) {
            __syncwarp();
            const unsigned int nrow = p.nrow, tot = p.nrow * p.ncol;
#pragma unroll
            for (int g = 0; g < G; g++) {
                const unsigned long long base = __shfl_sync(0xffffffffu, off, g);
                if (g < nv) {
                    double *dst = p.out + base;
                    for (unsigned int idx = lane; idx < tot; idx += 32) {
                        const unsigned int jj = idx / nrow, ii = idx - jj * nrow;
                        const bool in = ii >= p.ii0 && ii < p.ii0 + NDOF && jj >= p.jj0 && jj < p.jj0 + NDOF;
                        const double val = in ? Kst[g * KB + (ii - p.ii0) + NDOF * (jj - p.jj0)] : 0.0;
                        if (p.clear) dst[idx] = val;
                        else if (in) dst[idx] += val;
                    }
                }
            }
            __syncwarp();
            continue;
        }
        // runs of blocks that are consecutive in the output leave as one copy (always one, on single-kind meshes): lane k < nv owns
        // cell k, the first lane of a run issues its copy (bulk groups are per thread: that lane also waits for it)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        {
            const unsigned long long prev = __shfl_up_sync(0xffffffffu, off, 1);
            const bool start = lane < nv && (lane == 0 || off != prev + KB);
            const unsigned starts = __ballot_sync(0xffffffffu, start);
            if (start) {
                const unsigned above = starts >> (lane + 1);
                const int next = above ? lane + __ffs(above) : nv;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p.out + off),
                             "r"(q8_smem_u32(Kst + lane * KB)), "r"((unsigned)((next - lane) * KB * sizeof(double))), "l"(pol)
                             : "memory");
            }
            if (lane < G) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane < G) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NN, bool AXI>
int launch_qua8_kind(const IntegParams &p, const Qua8Consts &cst, int pat, const Qua8Extra &x, bool drec, bool scatter, bool gen, void *stream) {
    const int warps = drec ? Qua8Dims<NN, AXI, true>::WARPS : Qua8Dims<NN, AXI, false>::WARPS;
    constexpr int G = Qua8Dims<NN, AXI>::G;
    const size_t smem = sizeof(double) * (size_t)warps * (drec ? Qua8Dims<NN, AXI, true>::WSZ : Qua8Dims<NN, AXI, false>::WSZ);
    typedef void (*kern_t)(const IntegParams, const Qua8Consts, const Qua8Extra);
    auto pick = [&](int pt) -> kern_t {
        if (scatter) return drec ? qua8_kernel<NN, AXI, 2, false, true, true> : qua8_kernel<NN, AXI, 2, false, false, true>;
        if (gen) {
            if (drec) return pt == 2 ? qua8_kernel<NN, AXI, 2, false, true, false, true> : qua8_kernel<NN, AXI, 0, false, true, false, true>;
            return pt == 2 ? qua8_kernel<NN, AXI, 2, false, false, false, true> : qua8_kernel<NN, AXI, 0, false, false, false, true>;
        }
        if (drec) return pt == 2 ? qua8_kernel<NN, AXI, 2, false, true> : qua8_kernel<NN, AXI, 0, false, true>;
        if (x.mask) return pt == 2 ? qua8_kernel<NN, AXI, 2, true> : qua8_kernel<NN, AXI, 0, true>;
        return pt == 2 ? qua8_kernel<NN, AXI, 2, false> : qua8_kernel<NN, AXI, 0, false>;
    };
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long ntriple = (p.hi - p.lo + G - 1) / G;
    unsigned long long grid = (unsigned long long)nsm * Q8_CTAS; // persistent: warps stride over the triples
    if (const char *e = tuning_env("GL_QUA8_GRID")) grid = (unsigned long long)atoi(e);
    const unsigned long long need = (ntriple + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid == 0) return 0;
    // records whose structure only the device knows: both variants are enqueued, the probe's verdict selects the one that runs
    const bool probe = drec && p.coef_pat_dev != nullptr;
    IntegParams q = p;
    if (!probe) q.coef_pat_dev = nullptr;
    int nl = 0;
    for (int pt = probe ? 0 : pat; pt <= (probe ? 2 : pat); pt += 2) {
        kern_t kern = pick(pt);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, 32 * warps, smem, (cudaStream_t)stream>>>(q, cst, x);
        nl++;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    const char *family = NN == 8 ? "qua8_dmma" : "tri6_dmma";
    if (x.mask) note_kernel("%s<axi=%d,pat=%d,fused=%d>", family, AXI ? 1 : 0, pat, x.mask);
    else note_kernel("%s<axi=%d,pat=%d%s%s>", family, AXI ? 1 : 0, pat, drec ? ",rec" : "", scatter ? ",scatter" : (gen ? ",gen" : ""));
    return nl;
}

} // namespace

int launch_bdb_qua8(const IntegParams &p, int sd, void *stream, const ScalarFused *extra) {
    if (sd != 2 || p.op != GL_OP_MAT_10_BDB || (p.nnode != 8 && p.nnode != 6)) return 0;
    if (p.sigma != nullptr) return 0;
    const bool scatter = p.scatter_pos != nullptr; // fused assembly: symmetric patterns only (below)
    const bool drec = p.coef != nullptr;
    if (drec && (p.coef_gp_stride != 0 || p.coef_cell_stride != 16 || extra != nullptr)) return 0; // one modulus per Gauss point: gl_kernels_bdb.cu
    if ((p.nnode == 6 ? 4 : 3) * p.ng > 32) return 0;
    const unsigned ndof = 2u * (unsigned)p.nnode;
    if (p.size_r != ndof || p.size_c != ndof) return 0;
    const bool tight = p.nrow == ndof && p.ncol == ndof && p.ii0 == 0 && p.jj0 == 0;
    if (scatter && !tight) return 0;
    const bool gen = !scatter && !(tight && p.clear && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0); // general store epilogue
    if ((scatter || gen) && extra != nullptr) return 0;
    if (tuning_env("GL_DISABLE_QUA8")) return 0;

    Qua8Consts cst;
    const double hs = 0.70710678118654752440084436210484903928;
    bool sym = true;
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            const double fa = a < 3 ? 1.0 : hs, fb = b < 3 ? 1.0 : hs;
            cst.d[a * 4 + b] = p.coef_uniform[a + b * 4] * fa * fb;
            if (p.coef_uniform[a + b * 4] != p.coef_uniform[b + a * 4]) sym = false;
        }
    int pat = 0;
    if (sym) {
        pat = 2;
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++)
                if (!q8_nz(2, a, b) && cst.d[a * 4 + b] != 0.0) pat = 0;
    }
    if (drec) pat = p.coef_pat == 2 ? 2 : 0; // structure common to all records (0 when they live on the device only)
    if (const char *e = tuning_env("GL_BDB_PATTERN")) {
        if (atoi(e) < pat) pat = 0;
    }
    if (scatter && (pat != 2 || (drec && p.coef_pat_dev != nullptr))) return 0; // general D, or records only the device has seen: gl_kernels_bdb.cu
    Qua8Extra x;
    std::memset(&x, 0, sizeof(x));
    if (extra) { // mat_03_btb and / or vec_03_bv from the same pass (uniform coefficients, tight blocks, device outputs: checked by the caller)
        if ((extra->mask & ~10) != 0 || extra->mask == 0) return 0;
        x.mask = extra->mask;
        x.t00 = extra->t[0];
        x.t11 = extra->t[1];
        x.t01 = extra->t[3] * hs;
        x.w0 = extra->w[0];
        x.w1 = extra->w[1];
        x.out_m03 = extra->out_m03;
        x.out_v03 = extra->out_v03;
        x.moff = extra->moff;
        x.mbase = extra->mbase;
        x.voff = extra->voff;
        x.vbase = extra->vbase;
    }
    if (p.nnode == 6)
        return p.axisym ? launch_qua8_kind<6, true>(p, cst, pat, x, drec, scatter, gen, stream) : launch_qua8_kind<6, false>(p, cst, pat, x, drec, scatter, gen, stream);
    return p.axisym ? launch_qua8_kind<8, true>(p, cst, pat, x, drec, scatter, gen, stream) : launch_qua8_kind<8, false>(p, cst, pat, x, drec, scatter, gen, stream);
}

} // namespace gl
