// gl_kernels_syrk.cu -- stiffness kernel for the quadratic 3D solids (Hex20, Tet10) on the FP64 TENSOR CORES:
// integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-208) with a modulus D that does not vary inside a cell (uniform, or one
// record per marker / per cell).  BASELINE.json configs[2] ("Hex20 ... 27-point Gauss (DMMA contraction path)").
//
// Algebra (shared with gl_kernels_bdb.cu / gl_kernels_tile.cu): the Gauss sum moves inside the contraction with D,
//     P(m,n)[k][l] = sum_p c_p B_p(m,k) B_p(n,l),        K(m,n)[i][j] = sum_kl D_ikjl P(m,n)[k][l]   (once per node pair),
// and P = V^T diag(c) V with V[p][3m+k] = B_p(m,k) is a SYMMETRIC RANK-ng UPDATE of an ndof x ndof matrix: a real dense
// contraction (60 x 60 x 27 for Hex20).  The register-tiled DFMA form of gl_kernels_tile.cu needs 12 shared-memory doubles
// per 36 DFMA and is bound by the shared-memory pipe (ncu, round 1: LDS wavefronts 71 % of peak, FP64 pipe 31 %); no
// register tile that fits is large enough to change that.  mma.sync.m8n8k4.f64 shares every loaded operand between the lanes
// of a warp: with the 8 x 8 dof tiles of the upper triangle held as accumulator fragments, one k-step of a warp loads
// 8 doubles per lane for 18 DMMA (4,608 FMA) -- 9x fewer shared-memory wavefronts per FMA -- at 1.2x the flops (the 64-column
// padding and the full diagonal tiles), on the same pipe (measured: DMMA 37.1 vs DFMA 33.9 TFLOP/s).
//
// Mapping: ONE PERSISTENT CTA PER SM (Hex20) with SEVEN independent CELL SLOTS of two warps each (named barriers, no
// CTA-wide barrier in the loop); the reference-element tables are staged once per CTA and shared by the slots.  Per cell:
//   A  gather (Mesh::set_pad, src/mesh/mesh.rs:448-456), software-pipelined two cells ahead;
//   B  geometry, two lanes per Gauss point, in the ORACLE'S operation order (gl_geom.cuh): J = Xt.L, closed-form inverse,
//      det, B = L.inv(J) -> V rows in shared memory, c_p = det.w_p.alpha;
//   C  DMMA: each warp accumulates its half of the 36 upper 8 x 8 tiles of P over ceil(ng/4) k-steps; the scaled operand
//      c_p V is formed in registers;
//   D  the accumulator fragments go to shared memory in the final column-major layout (the tile is computed transposed so
//      that a lane's two values are adjacent), every lane contracts whole node pairs IN PLACE with D' (compile-time
//      sparsity pattern of the modulus, see gl_kernels_bdb.cu) and mirrors them, and the cell (28,800 contiguous bytes)
//      leaves with one bulk async copy (UBLKCP, L2 evict-first).  V, P and the staged K share one region.
// Tet10 (30 dofs -> 4 x 4 tiles, one warp per cell, 14 slots per CTA, two CTAs per SM) uses the same template.
// Non-tight blocks (nrow / ncol / ii0 / jj0) and clear = false are handled by a cooperative store epilogue (GEN).
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

// strain-displacement column (node, j) in 3D: entries q = 0..2 -> (Mandel row, gradient component)  (calc_bb_matrix,
// src/integ/mat_10_bdb.rs:238-278)
__host__ __device__ constexpr int col_row3(int j, int q) {
    return j == 0 ? (q == 0 ? 0 : (q == 1 ? 3 : 5)) : (j == 1 ? (q == 0 ? 1 : (q == 1 ? 3 : 4)) : (q == 0 ? 2 : (q == 1 ? 4 : 5)));
}
__host__ __device__ constexpr int col_comp3(int j, int q) {
    return j == 0 ? (q == 0 ? 0 : (q == 1 ? 1 : 2)) : (j == 1 ? (q == 0 ? 1 : (q == 1 ? 0 : 2)) : (q == 0 ? 2 : (q == 1 ? 1 : 0)));
}
// PAT 0 general, PAT 1 major-symmetric, PAT 2 symmetric with uncoupled normal / diagonal shear blocks (exact zeros)
__host__ __device__ constexpr bool d_nz3(int pat, int a, int b) { return pat < 2 ? true : ((a < 3 && b < 3) || a == b); }

struct SyrkConsts {
    double d[36]; // d'[a][b] = f_a f_b d[a][b] (row-major 6 x 6), f = 1 for normal rows, 1/sqrt(2) for shear rows
    int stagger_ns; // start-up delay per cell slot (see the kernel)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int SYRK_THREADS = 448;

#include "syrk_pairs.inc"

template <int NN>
struct SyrkDims {
    static constexpr int NDOF = NN * 3;
    static constexpr int NT8 = (NDOF + 7) / 8;            // 8-dof tiles per side
    static constexpr int WPC = NN == 20 ? 2 : 1;          // warps per cell
    static constexpr int LANES = 32 * WPC;                // lanes per cell slot
    static constexpr int SLOTS = SYRK_THREADS / LANES;    // cell slots per CTA
    static constexpr int MAXNG = LANES / 2;               // two lanes per Gauss point
    static constexpr int PS = MAXNG + 4;                  // V is stored DOF-MAJOR, Vt[c][p] at c PS + p; PS = 4 (mod 16): the fragment loads of a
                                                          // half-warp (4 columns x 4 Gauss points) hit 16 distinct bank pairs, and the 16 Gauss
                                                          // points of a half-warp store one dof of theirs to 16 consecutive doubles
    static constexpr int LS = NDOF | 1;                   // L row stride (odd: the Gauss points of a warp read distinct banks)
    static constexpr int RS = NDOF * NDOF;                // region per slot: Vt (8 NT8 x PS) / P / staged K
    static constexpr int NPAIR = NN * (NN + 1) / 2;       // node pairs m <= n
    static constexpr int NPASS = (NPAIR + LANES - 1) / LANES;
    static constexpr int CTAS_PER_SM = NN == 20 ? 1 : 2;
    static_assert(NDOF % 2 == 0 && NDOF <= LANES, "gather: one lane per coordinate");
    static_assert(NT8 * 8 * PS <= RS, "V must fit inside the region it shares with the staged element block");
};

// upper 8 x 8 dof tiles (I <= J) owned by warp W of a cell
__host__ __device__ constexpr int syrk_ntile(int nn, int w) { return nn == 20 ? 18 : (w == 0 ? 10 : 0); }
__host__ __device__ constexpr int syrk_tile(int nn, int w, int t, int which) {
    if (nn == 20) {
        if (w == 0) { // rows 0..2 against the columns to their right
            constexpr int I[18] = {0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2};
            constexpr int J[18] = {1, 2, 3, 4, 5, 6, 7, 2, 3, 4, 5, 6, 7, 3, 4, 5, 6, 7};
            return which == 0 ? I[t] : J[t];
        }
        constexpr int I[18] = {3, 3, 3, 3, 3, 4, 4, 4, 4, 0, 1, 2, 5, 5, 5, 6, 6, 7};
        constexpr int J[18] = {3, 4, 5, 6, 7, 4, 5, 6, 7, 0, 1, 2, 5, 6, 7, 6, 7, 7};
        return which == 0 ? I[t] : J[t];
    }
    constexpr int I[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
    constexpr int J[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};
    return which == 0 ? I[t] : J[t];
}

template <int WPC>
__device__ __forceinline__ void slot_sync(int slot) {
    if constexpr (WPC == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(32 * WPC) : "memory");
    }
}

// phase C + first half of phase D for warp W of a cell: accumulate the warp's tiles, then (after every warp of the slot has
// finished reading V) store them as P into the region
template <int NN, int W>
__device__ __forceinline__ void syrk_tiles(double *R, const double *s_c, int KS, int lane, int slot) {
    using T = SyrkDims<NN>;
    constexpr int NT8 = T::NT8, PS = T::PS, NDOF = T::NDOF;
    constexpr int NTL = syrk_ntile(NN, W);
    const int ar = lane >> 2, ak = lane & 3;
    double acc[NTL > 0 ? NTL : 1][2];
#pragma unroll
    for (int t = 0; t < NTL; t++) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll 1
    for (int ks = 0; ks < KS; ks++) {
        const int row = 4 * ks + ak;
        const double c = s_c[row];
        const double *vr = R + ar * PS + row;
        double v[NT8], vs[NT8];
#pragma unroll
        for (int t = 0; t < NT8; t++) {
            v[t] = vr[8 * t * PS];
            vs[t] = v[t] * c;
        }
        // C[r][q] += sum_k A[r][k] B[k][q] with A[r][k] = V[k][8J + r], B[k][q] = c_k V[k][8I + q]: C[r][q] = P(8I + q, 8J + r)
#pragma unroll
        for (int t = 0; t < NTL; t++) dmma_m8n8k4(acc[t][0], acc[t][1], v[syrk_tile(NN, W, t, 1)], vs[syrk_tile(NN, W, t, 0)]);
    }
    slot_sync<T::WPC>(slot); // every warp of the slot has finished reading V: the region becomes P
    // the lane holds P(8I + 2ak + e, 8J + ar), e = 0, 1.  Two 8-byte stores; odd fragment rows store e = 1 first: at
    // (8I + 2ak + e) + 60 (8J + ar) = 2ak + e + 12 ar (mod 16) the 16 lanes of a half-warp then hit 16 distinct bank pairs
    // (one 16-byte store per lane would put rows ar and ar + 1 of a quarter-warp on the same banks)
    const int e0 = ar & 1;
#pragma unroll
    for (int t = 0; t < NTL; t++) {
        const int prow = 8 * syrk_tile(NN, W, t, 0) + 2 * ak; // rows prow, prow + 1
        const int pcol = 8 * syrk_tile(NN, W, t, 1) + ar;
        if (prow < NDOF && pcol < NDOF) {
            double *dst = R + prow + NDOF * pcol;
            dst[e0] = e0 ? acc[t][1] : acc[t][0];
            dst[1 - e0] = e0 ? acc[t][0] : acc[t][1];
        }
    }
}

template <int NN, int PAT, bool DREC, bool GEN>
__global__ void __launch_bounds__(SYRK_THREADS, SyrkDims<NN>::CTAS_PER_SM) syrk_kernel(const IntegParams p, const SyrkConsts cst) {
    if (DREC && p.coef_pat_dev != nullptr && *p.coef_pat_dev != PAT) return; // another pattern variant owns this call (see launcher)
    using T = SyrkDims<NN>;
    constexpr int NDOF = T::NDOF, PS = T::PS, WPC = T::WPC, LANES = T::LANES, SLOTS = T::SLOTS, MAXNG = T::MAXNG, LS = T::LS,
                  RS = T::RS, NPASS = T::NPASS;
    constexpr bool SYM = PAT >= 1;

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    const int KS = (ng + 3) >> 2;
    double *s_R = smem;                        // [SLOTS][RS]      (16-byte aligned: bulk copy source, double2 stores)
    double *s_L = s_R + SLOTS * RS;            // [MAXNG][LS]      L[gp][3m + j]; rows >= ng are never initialised: the lanes without a
                                               // Gauss point read them (their own row: no bank shared with a live row) and discard the result
    double *s_w = s_L + MAXNG * LS;            // [ng]
    double *s_X = s_w + ng;                    // [SLOTS][NDOF]    X[3m + j]
    double *s_c = s_X + SLOTS * NDOF;          // [SLOTS][MAXNG]   c_p (0 for the padding rows of the last k-step)
    double *s_D = s_c + SLOTS * MAXNG;         // [SLOTS][36]      d' of the slot's cell (DREC only)

    const int tid = threadIdx.x;
    for (int i = tid; i < ng; i += SYRK_THREADS) s_w[i] = p.rule[i];
    for (int i = tid; i < ng * NDOF; i += SYRK_THREADS) s_L[(i / NDOF) * LS + i % NDOF] = p.rule[ng * (1 + NN) + i];
    __syncthreads();

    const int slot = tid / LANES;
    const int ls = tid - slot * LANES; // lane within the slot
    const int w = ls >> 5, lane = ls & 31;
    double *R = s_R + slot * RS;
    double *X = s_X + slot * NDOF;
    double *cc = s_c + slot * MAXNG;
    const double *Dc = s_D + slot * 36;

    // node pairs (m <= n) of this lane in the in-place contraction: a table (tools/gen_syrk_pairs.py) that gives the 16 lanes of every
    // half-warp pairs whose direct blocks (3m + 60.3n) and mirrored blocks (3n + 60.3m) start on 16 distinct bank pairs
    int pair_mn[NPASS];
#pragma unroll
    for (int k = 0; k < NPASS; k++) {
        const unsigned short v = NN == 20 ? SYRK_PAIRS_20[k * LANES + ls] : SYRK_PAIRS_10[k * LANES + ls];
        pair_mn[k] = v == 0xffff ? -1 : (int)v;
    }

    // software-pipelined gather: lane (node gm, component gj) holds the coordinate of the NEXT cell of the slot and the point id
    // of the cell after that
    const unsigned long long stride = (unsigned long long)gridDim.x * SLOTS;
    unsigned long long cell = p.lo + (unsigned long long)blockIdx.x * SLOTS + slot;
    const bool gather_lane = ls < NDOF;
    const int gm = ls / 3, gj = ls - gm * 3;
    double gx = 0.0;
    unsigned int gpid = 0;
    if (gather_lane) {
        if (cell < p.hi) gx = p.coords[(unsigned long long)gj * p.npoint + p.conn[(unsigned long long)gm * p.ncell_k + cell]];
        if (cell + stride < p.hi) gpid = p.conn[(unsigned long long)gm * p.ncell_k + cell + stride];
    }
    unsigned long long pol; // the element matrices are written once: keep them from flushing the mesh out of L2
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

    // All slots of a CTA would otherwise run in LOCKSTEP -- geometry (FP64 + shared memory), tensor cores, epilogue (shared memory
    // only) at the same time on every slot -- and the pipes would idle in turns.  A start-up delay proportional to the slot index
    // spreads the phases; every cell costs the same, so the skew persists.
    if (cst.stagger_ns > 0 && slot > 0) __nanosleep((unsigned)(cst.stagger_ns * slot));

    for (; cell < p.hi; cell += stride) {
        const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
        // ---- phase A: publish the coordinates, prefetch ----------------------------------------------------------------
        if (gather_lane) {
            X[ls] = gx;
            if (cell + stride < p.hi) gx = p.coords[(unsigned long long)gj * p.npoint + gpid];
            if (cell + 2 * stride < p.hi) gpid = p.conn[(unsigned long long)gm * p.ncell_k + cell + 2 * stride];
        }
        if constexpr (DREC) {
            // the cell's modulus record (ns x ns Mandel matrix, column-major; its marker's or its own), shear factors folded in
            const double hs = 0.70710678118654752440084436210484903928;
            const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell] : orig - p.coef_begin;
            for (int t = ls; t < 36; t += LANES) {
                const int a = t / 6, b = t - a * 6;
                s_D[slot * 36 + t] = p.coef[rec * p.coef_cell_stride + a + b * 6] * (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
            }
        }
        slot_sync<WPC>(slot);

        // ---- phase B: geometry, two lanes per Gauss point ---------------------------------------------------------------
        // The 16 Gauss points of a warp sit on lanes 0..15 (h = 0) and again on lanes 16..31 (h = 1).  Lane h of a pair
        // accumulates row 2h of J and its share of row 1 (entries (1,0) and (1,1) for h = 0, (1,2) and again (1,1) for h = 1),
        // every entry as the oracle does: m = 0 .. NN-1, product and sum rounded separately.  A half-warp reads one node of 16
        // different rows of L (odd stride: 16 distinct bank pairs).
        const int q = 16 * w + (lane & 15), h = lane >> 4;
        const bool valid = q < ng;
        const double *L = s_L + q * LS;
        double Ji[3][3];
        double det;
        {
            double Ja[3] = {0.0, 0.0, 0.0}, Jb0 = 0.0, Jb1 = 0.0;
#pragma unroll 4
            for (int m = 0; m < NN; m++) {
                const double xa = X[3 * m + 2 * h], xb = X[3 * m + 1];
                const double l0 = L[3 * m], l1 = L[3 * m + 1], l2 = L[3 * m + 2];
                Ja[0] = geom::add(Ja[0], geom::mul(xa, l0));
                Ja[1] = geom::add(Ja[1], geom::mul(xa, l1));
                Ja[2] = geom::add(Ja[2], geom::mul(xa, l2));
                Jb0 = geom::add(Jb0, geom::mul(xb, h ? l2 : l0));
                Jb1 = geom::add(Jb1, geom::mul(xb, l1));
            }
            double Jo[3], Jbo;
#pragma unroll
            for (int j = 0; j < 3; j++) Jo[j] = __shfl_xor_sync(0xffffffffu, Ja[j], 16);
            Jbo = __shfl_xor_sync(0xffffffffu, Jb0, 16);
            double J[3][3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                J[0][j] = h ? Jo[j] : Ja[j];
                J[2][j] = h ? Ja[j] : Jo[j];
            }
            J[1][0] = h ? Jbo : Jb0;
            J[1][1] = Jb1;
            J[1][2] = h ? Jb0 : Jbo;
            det = geom::invert(J, Ji);
        }
        if (valid && h == 0 && fabs(det) <= ZERO_DETERMINANT) atomicMin(p.err_cell, orig);
        if constexpr (!GEN) {
            if (ls == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the previous bulk copy has left the region
        }
        slot_sync<WPC>(slot);
        if (q < 4 * KS) {
            // gradient rows of the nodes m = h, h + 2, .. (B = L.inv(J) is well conditioned: fused multiply-adds); the padding
            // Gauss points of the last k-step are zero-filled (the region held the previous element block)
            double *Vq = R + q;
#pragma unroll 2
            for (int m = h; m < NN; m += 2) {
                const double l0 = L[3 * m], l1 = L[3 * m + 1], l2 = L[3 * m + 2];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const double b = fma(l2, Ji[2][j], fma(l1, Ji[1][j], l0 * Ji[0][j]));
                    Vq[(3 * m + j) * PS] = valid ? b : 0.0;
                }
            }
            if (h == 0) cc[q] = valid ? det * s_w[valid ? q : 0] * p.alpha : 0.0;
        }
        slot_sync<WPC>(slot);

        // ---- phase C: P = V^T diag(c) V on the tensor cores; the fragments become P in the region --------------------------
        if constexpr (WPC == 2) {
            if (w == 0) syrk_tiles<NN, 0>(R, cc, KS, lane, slot);
            else syrk_tiles<NN, 1>(R, cc, KS, lane, slot);
        } else {
            syrk_tiles<NN, 0>(R, cc, KS, lane, slot);
        }
        slot_sync<WPC>(slot);

        // ---- phase D: K(m,n) = D' : P(m,n), in place, and its mirror ------------------------------------------------------
#pragma unroll
        for (int k = 0; k < NPASS; k++) {
            if (pair_mn[k] < 0) continue;
            const int m = pair_mn[k] & 0xff, n = pair_mn[k] >> 8;
            const bool dg = m == n;
            // P(m,n)[k][l] at (3m + k) + NDOF (3n + l); only the upper triangle of P was computed, so the diagonal pair reads
            // its upper half and mirrors it
            double P[3][3];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const bool low = dg && a > b;
                    P[a][b] = R[(3 * m + (low ? b : a)) + NDOF * (3 * n + (low ? a : b))];
                }
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    // K(3m+i, 3n+j) = sum_qr d'[row(i,q)][row(j,r)] P[comp(i,q)][comp(j,r)]; for a non-symmetric D the mirrored block
                    // K(3n+j, 3m+i) = sum_qr d'[row(j,r)][row(i,q)] P[comp(i,q)][comp(j,r)]  (P(n,m) = P(m,n)^T)
                    double v = 0.0, u = 0.0;
                    bool have = false;
#pragma unroll
                    for (int qq = 0; qq < 3; qq++)
#pragma unroll
                        for (int r = 0; r < 3; r++) {
                            const int ra = col_row3(i, qq), rb = col_row3(j, r);
                            if (!d_nz3(PAT, ra, rb)) continue;
                            const double dv = DREC ? Dc[ra * 6 + rb] : cst.d[ra * 6 + rb];
                            const double pv = P[col_comp3(i, qq)][col_comp3(j, r)];
                            v = have ? fma(dv, pv, v) : dv * pv;
                            if (!SYM) {
                                const double dt = DREC ? Dc[rb * 6 + ra] : cst.d[rb * 6 + ra];
                                u = have ? fma(dt, pv, u) : dt * pv;
                            }
                            have = true;
                        }
                    R[(3 * m + i) + NDOF * (3 * n + j)] = v;
                    if (!dg) R[(3 * n + j) + NDOF * (3 * m + i)] = SYM ? v : u;
                }
        }

        // ---- ship --------------------------------------------------------------------------------------------------------
        if constexpr (!GEN) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            slot_sync<WPC>(slot);
            if (ls == 0) {
                double *dst = p.out + (p.out_off ? p.out_off[cell] - p.out_base : (orig - p.cell_begin) * (unsigned long long)RS);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(R)),
                             "r"((unsigned)(RS * sizeof(double))), "l"(pol)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // element block nrow x ncol with the integrated ndof x ndof window at (ii0, jj0): clear = true overwrites the whole
            // block (kk.fill(0.0), mat_10_bdb.rs:141-143), clear = false adds the window
            slot_sync<WPC>(slot);
            const unsigned int nrow = p.nrow, ncol = p.ncol;
            double *dst = p.out + (p.out_off ? p.out_off[cell] - p.out_base : (orig - p.cell_begin) * ((unsigned long long)nrow * ncol));
            for (unsigned int idx = ls; idx < nrow * ncol; idx += LANES) {
                const unsigned int jj = idx / nrow, ii = idx - jj * nrow;
                const bool in = ii >= p.ii0 && ii < p.ii0 + NDOF && jj >= p.jj0 && jj < p.jj0 + NDOF;
                const double val = in ? R[(ii - p.ii0) + NDOF * (jj - p.jj0)] : 0.0;
                if (p.clear) dst[idx] = val;
                else if (in) dst[idx] += val;
            }
            slot_sync<WPC>(slot);
        }
    }
    if constexpr (!GEN) {
        if (ls == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int NN>
size_t syrk_smem(int ng, bool drec) {
    using T = SyrkDims<NN>;
    return sizeof(double) * ((size_t)T::SLOTS * T::RS + (size_t)T::MAXNG * T::LS + (size_t)ng + (size_t)T::SLOTS * T::NDOF +
                             (size_t)T::SLOTS * T::MAXNG + (drec ? (size_t)T::SLOTS * 36 : 0));
}

template <int NN, bool GEN>
int launch_syrk_kind(const IntegParams &p, const SyrkConsts &cst, int pat, bool drec, void *stream) {
    using T = SyrkDims<NN>;
    if (p.ng > T::MAXNG) return 0;
    const size_t smem = syrk_smem<NN>(p.ng, drec);
    if (smem > 227 * 1024) return 0;
    typedef void (*kern_t)(const IntegParams, const SyrkConsts);
    auto pick = [&](int pt) -> kern_t {
        return drec ? (pt == 2 ? syrk_kernel<NN, 2, true, GEN> : (pt == 1 ? syrk_kernel<NN, 1, true, GEN> : syrk_kernel<NN, 0, true, GEN>))
                    : (pt == 2 ? syrk_kernel<NN, 2, false, GEN> : (pt == 1 ? syrk_kernel<NN, 1, false, GEN> : syrk_kernel<NN, 0, false, GEN>));
    };
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long ncell = p.hi - p.lo;
    unsigned long long grid = (unsigned long long)nsm * T::CTAS_PER_SM; // persistent: every slot strides over the range
    if (const char *e = tuning_env("GL_SYRK_GRID")) grid = (unsigned long long)atoi(e);
    const unsigned long long need = (ncell + T::SLOTS - 1) / T::SLOTS;
    if (grid > need) grid = need;
    if (grid == 0) return 0;
    // records whose structure only the device knows: every variant is enqueued, the probe's verdict selects the one that runs
    const bool probe = drec && p.coef_pat_dev != nullptr;
    IntegParams q = p;
    if (!probe) q.coef_pat_dev = nullptr;
    for (int pt = probe ? 0 : pat; pt <= (probe ? 2 : pat); pt++) {
        kern_t kern = pick(pt);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, SYRK_THREADS, smem, (cudaStream_t)stream>>>(q, cst);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("syrk_dmma<nn=%d,pat=%d%s%s>", NN, pat, drec ? ",rec" : "", GEN ? ",gen" : "");
    return 1;
}

} // namespace

int launch_bdb_syrk(const IntegParams &p, int sd, void *stream) {
    // covered: 3D Hex20 / Tet10, rules of at most 32 / 16 points, uniform D or one record per marker / cell; homogeneous or
    // mixed-kind layout; any element block (tight blocks that overwrite a 16-byte aligned buffer take the bulk-copy epilogue)
    if (sd != 3 || p.op != GL_OP_MAT_10_BDB || p.axisym) return 0;
    if (p.nnode != 20 && p.nnode != 10) return 0;
    const bool drec = p.coef != nullptr;
    if (drec && (p.coef_gp_stride != 0 || p.coef_cell_stride != 36)) return 0; // PER_GAUSS moduli: not a rank-ng update
    if (tuning_env("GL_DISABLE_SYRK")) return 0;
    if (p.nnode == 10 && !tuning_env("GL_SYRK_TET10")) return 0; // Tet10 stays on gl_kernels_bdb.cu until measured
    const unsigned ndof = (unsigned)p.nnode * 3u;
    const bool tight = p.nrow == ndof && p.ncol == ndof && p.ii0 == 0 && p.jj0 == 0;
    const bool gen = !(tight && p.clear && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

    SyrkConsts cst;
    cst.stagger_ns = p.nnode == 20 ? 1400 : 300;
    if (const char *e = tuning_env("GL_SYRK_STAGGER")) cst.stagger_ns = atoi(e);
    const double hs = 0.70710678118654752440084436210484903928;
    bool sym = true;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            const double fa = a < 3 ? 1.0 : hs, fb = b < 3 ? 1.0 : hs;
            cst.d[a * 6 + b] = p.coef_uniform[a + b * 6] * fa * fb;
            if (p.coef_uniform[a + b * 6] != p.coef_uniform[b + a * 6]) sym = false;
        }
    int pat = 0;
    if (sym) {
        pat = 2;
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++)
                if (!d_nz3(2, a, b) && cst.d[a * 6 + b] != 0.0) pat = 1;
    }
    if (drec) pat = p.coef_pat; // structure common to all records (0 when the records live on the device only)
    if (const char *e = tuning_env("GL_BDB_PATTERN")) {
        const int cap = atoi(e);
        if (cap < pat) pat = cap < 0 ? 0 : cap;
    }
    if (p.nnode == 20) return gen ? launch_syrk_kind<20, true>(p, cst, pat, drec, stream) : launch_syrk_kind<20, false>(p, cst, pat, drec, stream);
    return gen ? launch_syrk_kind<10, true>(p, cst, pat, drec, stream) : launch_syrk_kind<10, false>(p, cst, pat, drec, stream);
}

} // namespace gl
