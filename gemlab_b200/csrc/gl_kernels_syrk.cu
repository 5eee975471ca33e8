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
// of a warp: with the 36 upper 8 x 8 dof tiles held as accumulator fragments, one k-step loads 8 doubles per lane for 36 DMMA
// (9,216 FMA) -- 18x fewer shared-memory wavefronts per FMA -- at 1.2x the flops (the 64-column padding and the full diagonal
// tiles), on the same pipe (measured: DMMA 37.1 vs DFMA 33.9 TFLOP/s; one warp with two independent accumulators already
// saturates it, tools/dmma_ubench.cu).
//
// Mapping: ONE WARP PER CELL, seven cell slots (warps) in one persistent CTA per SM for Hex20 (the staged element block,
// 28.8 KB, bounds the cells in flight); no barrier wider than a warp; the reference-element tables are staged once per CTA.
// Slots draw cells from a global counter (the sub-partitions of an SM hold 2, 2, 2 and 1 of the warps).  Per cell:
//   A  gather (Mesh::set_pad, src/mesh/mesh.rs:448-456), point ids two cells ahead, coordinates one cell ahead;
//   B  geometry, lane = Gauss point, in the ORACLE'S operation order (gl_geom.cuh): J = Xt.L, closed-form inverse, det;
//      B = L.inv(J) -> the dof-major record Vt in shared memory, c_p = det.w_p.alpha;
//   C  DMMA: the 36 upper tiles of P over ceil(ng/4) k-steps; the scaled operand c_p V is formed in registers;
//   D  the accumulator fragments go to shared memory in the final column-major layout, every lane contracts whole node pairs
//      IN PLACE with D' (compile-time sparsity pattern of the modulus, see gl_kernels_bdb.cu) and mirrors them, and the cell
//      (28,800 contiguous bytes) leaves with one bulk async copy (UBLKCP, L2 evict-first).  Vt, P and the staged K share one
//      region.  Every shared-memory access pattern is conflict-free by construction (strides, store order, pair table).
// Tet10 (30 dofs -> 4 x 4 tiles, 14 slots per CTA) uses the same template.
// Non-tight blocks (nrow / ncol / ii0 / jj0) and clear = false are handled by a cooperative store epilogue (GEN).
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>
#include <utility>

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
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// c ? a : b on registers (written as selp so that the compiler does not turn the pair into a dynamically indexed local array)
__device__ __forceinline__ double select_f64(int c, double a, double b) {
    double r;
    asm("{\n.reg .pred p;\nsetp.ne.s32 p, %3, 0;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"(c));
    return r;
}

#include "syrk_pairs.inc"

template <int NN>
struct SyrkDims {
    static constexpr int NDOF = NN * 3;
    static constexpr int NT8 = (NDOF + 7) / 8;            // 8-dof tiles per side
    static constexpr int NTILE = NT8 * (NT8 + 1) / 2;     // upper tiles incl. the diagonal
    static constexpr int MAXNG = NN == 20 ? 32 : 16;      // one lane per Gauss point
    static constexpr int PS = MAXNG + 4;                  // V is stored DOF-MAJOR, Vt[c][p] at c PS + p; PS = 4 (mod 16): the fragment loads of a
                                                          // half-warp (4 columns x 4 Gauss points) hit 16 distinct bank pairs, and the Gauss
                                                          // points of a half-warp store one dof of theirs to consecutive doubles
    static constexpr int SLOTS = NN == 20 ? 7 : 14;       // cell slots (= warps) per CTA
    static constexpr int THREADS = 32 * SLOTS;
    static constexpr int CTAS_PER_SM = 1;
    static constexpr int LS = NDOF | 1;                   // L row stride (odd: the Gauss points of a half-warp read distinct banks)
    static constexpr int RS = NDOF * NDOF;                // region per slot: Vt (8 NT8 x PS) / P / staged K
    static constexpr int NPASS = NN == 20 ? SYRK_NPASS_20 : SYRK_NPASS_10;
    static constexpr int NCOORD = (NDOF + 31) / 32;       // coordinates gathered per lane
    static_assert(NDOF % 2 == 0, "16-byte aligned element blocks");
    static_assert(NT8 * 8 * PS <= RS, "V must fit inside the region it shares with the staged element block");
};

// upper 8 x 8 dof tiles (I <= J), row by row
__host__ __device__ constexpr int syrk_tile_i(int nt8, int t) {
    int i = 0;
    while (t >= nt8 - i) {
        t -= nt8 - i;
        i++;
    }
    return i;
}
__host__ __device__ constexpr int syrk_tile_j(int nt8, int t) {
    int i = 0;
    while (t >= nt8 - i) {
        t -= nt8 - i;
        i++;
    }
    return i + t;
}

// f(integral_constant<int, 0>{}), .., f(integral_constant<int, N-1>{}): a loop whose index is a constant EXPRESSION in the body (the
// tile -> (I, J) maps above are then evaluated by the compiler; with a merely unrolled loop they end up as run-time loops and the
// fragment arrays as dynamically indexed local memory)
template <typename F, int... Is>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

template <int NN, int PAT, bool DREC, bool GEN>
__global__ void __launch_bounds__(SyrkDims<NN>::THREADS, SyrkDims<NN>::CTAS_PER_SM) syrk_kernel(const IntegParams p, const SyrkConsts cst) {
    if (DREC && p.coef_pat_dev != nullptr && *p.coef_pat_dev != PAT) return; // another pattern variant owns this call (see launcher)
    using T = SyrkDims<NN>;
    constexpr int NDOF = T::NDOF, NT8 = T::NT8, NTILE = T::NTILE, PS = T::PS, SLOTS = T::SLOTS, MAXNG = T::MAXNG, LS = T::LS, RS = T::RS,
                  NPASS = T::NPASS, NCOORD = T::NCOORD;
    constexpr bool SYM = PAT >= 1;

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    const int KS = (ng + 3) >> 2;
    double *s_R = smem;                        // [SLOTS][RS]      (16-byte aligned: bulk copy source)
    double *s_L = s_R + SLOTS * RS;            // [MAXNG][LS]      L[gp][3m + j]; rows >= ng are zero: the lanes without a Gauss point read
                                               // their own row (no bank shared with a live row) and produce zero gradients
    double *s_w = s_L + MAXNG * LS;            // [MAXNG]
    double *s_X = s_w + MAXNG;                 // [SLOTS][NDOF]    X[3m + j]  (16-byte aligned: MAXNG (LS + 1) is even)
    double *s_c = s_X + SLOTS * NDOF;          // [SLOTS][MAXNG]   c_p (0 for the padding rows of the last k-step)
    double *s_D = s_c + SLOTS * MAXNG;         // [SLOTS][36]      d' of the slot's cell (DREC only)

    const int tid = threadIdx.x;
    for (int i = tid; i < MAXNG; i += T::THREADS) s_w[i] = i < ng ? p.rule[i] : 0.0;
    for (int i = tid; i < MAXNG * NDOF; i += T::THREADS) s_L[(i / NDOF) * LS + i % NDOF] = i < ng * NDOF ? p.rule[ng * (1 + NN) + i] : 0.0;
    __syncthreads();

    // ONE WARP PER CELL SLOT: no barrier wider than the warp below this line
    const int slot = tid >> 5, lane = tid & 31;
    double *R = s_R + slot * RS;
    double *X = s_X + slot * NDOF;
    double *cc = s_c + slot * MAXNG;
    const double *Dc = s_D + slot * 36;
    const int ar = lane >> 2, ak = lane & 3;

    // node pairs (m <= n) of this lane in the in-place contraction: a table (tools/gen_syrk_pairs.py) that gives the 16 lanes of every
    // half-warp pairs whose direct blocks (at 3m + NDOF.3n) and mirrored blocks (at 3n + NDOF.3m) start on 16 distinct bank pairs;
    // an entry holds the two block starts
    unsigned int pair_off[NPASS];
#pragma unroll
    for (int k = 0; k < NPASS; k++) pair_off[k] = NN == 20 ? SYRK_PAIRS_20[k * 32 + lane] : SYRK_PAIRS_10[k * 32 + lane];

    // Work distribution: the slots of all CTAs draw cells from one counter (equal shares would leave the sub-partition that holds
    // a single warp half idle).  A slot knows its next two cells: the point ids of the cell after next and the coordinates of the
    // next cell are in flight while the current cell is integrated.
    const unsigned long long ncell = p.hi - p.lo;
    auto draw_issue = [&]() -> unsigned long long { // lane 0 takes the next cell off the queue; the result is broadcast LATER
        // (inline PTX: atomicAdd() would be warp-aggregated by the compiler, with a shuffle that waits for the result right here)
        unsigned long long v = 0;
        if (lane == 0) asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(v) : "l"(p.work_counter) : "memory");
        return v;
    };
    auto draw_get = [&](unsigned long long raw) -> unsigned long long { return __shfl_sync(0xffffffffu, raw, 0); };
    auto load_pid = [&](unsigned long long c, unsigned int (&pid)[NCOORD]) {
#pragma unroll
        for (int k = 0; k < NCOORD; k++) {
            const int dof = lane + 32 * k;
            pid[k] = (c < ncell && dof < NDOF) ? p.conn[(unsigned long long)(dof / 3) * p.ncell_k + p.lo + c] : 0u;
        }
    };
    auto load_x = [&](unsigned long long c, const unsigned int (&pid)[NCOORD], double (&x)[NCOORD]) {
#pragma unroll
        for (int k = 0; k < NCOORD; k++) {
            const int dof = lane + 32 * k;
            x[k] = (c < ncell && dof < NDOF) ? p.coords[(unsigned long long)(dof % 3) * p.npoint + pid[k]] : 0.0;
        }
    };
    // c_cur: the cell being integrated (coordinates in gx); c_nxt: the next one (point ids in pid); raw_nn: the one after, taken off
    // the queue during the previous cell and broadcast at the top of this one, so the atomic's round trip is never waited for
    unsigned long long c_cur = draw_get(draw_issue()), c_nxt = draw_get(draw_issue()), raw_nn = draw_issue();
    unsigned int pid[NCOORD];
    double gx[NCOORD];
    load_pid(c_cur, pid);
    load_x(c_cur, pid, gx);
    load_pid(c_nxt, pid);

    unsigned long long pol; // the element matrices are written once: keep them from flushing the mesh out of L2
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));

    while (c_cur < ncell) {
        const unsigned long long cell = p.lo + c_cur;
        const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell;
        // ---- phase A: publish the coordinates, prefetch ----------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < NCOORD; k++)
            if (lane + 32 * k < NDOF) X[lane + 32 * k] = gx[k];
        {
            const unsigned long long c_nn = draw_get(raw_nn);
            load_x(c_nxt, pid, gx);  // coordinates of the next cell (its point ids arrived during the previous cell)
            load_pid(c_nn, pid);     // point ids of the cell after next
            c_cur = c_nxt;
            c_nxt = c_nn;
            raw_nn = draw_issue();
        }
        if constexpr (DREC) {
            // the cell's modulus record (ns x ns Mandel matrix, column-major; its marker's or its own), shear factors folded in
            const double hs = 0.70710678118654752440084436210484903928;
            const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell] : orig - p.coef_begin;
            for (int t = lane; t < 36; t += 32) {
                const int a = t / 6, b = t - a * 6;
                s_D[slot * 36 + t] = p.coef[rec * p.coef_cell_stride + a + b * 6] * (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
            }
        }
        __syncwarp();

        // ---- phase B: geometry, lane = Gauss point -------------------------------------------------------------------------
        // J = Xt.L exactly as the oracle evaluates it (m = 0 .. NN-1, product and sum rounded separately), closed-form inverse,
        // det; B = L.inv(J) is well conditioned and uses fused multiply-adds.  The lane's row of L (constant across cells) is read
        // once per cell and serves both loops.
        const bool valid = lane < ng;
        {
            const double *L = s_L + (lane < MAXNG ? lane : 0) * LS;
            double l[NDOF];
#pragma unroll
            for (int i = 0; i < NDOF; i++) l[i] = L[i];
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            // the cell's coordinates: the same 16-byte words for every lane (broadcast loads), four nodes (six words) at a time
            static_assert(NN % 4 == 0 || NN % 2 == 0, "coordinate words are read in pairs of nodes");
            constexpr int NB = NN % 4 == 0 ? 4 : 2; // nodes per batch
#pragma unroll
            for (int m0 = 0; m0 < NN; m0 += NB) {
                double x[3 * NB];
#pragma unroll
                for (int i = 0; i < 3 * NB / 2; i++) {
                    const double2 xx = reinterpret_cast<const double2 *>(X + 3 * m0)[i];
                    x[2 * i] = xx.x;
                    x[2 * i + 1] = xx.y;
                }
#pragma unroll
                for (int mm = 0; mm < NB; mm++) {
                    const int m = m0 + mm;
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        J[0][j] = geom::add(J[0][j], geom::mul(x[3 * mm], l[3 * m + j]));
                        J[1][j] = geom::add(J[1][j], geom::mul(x[3 * mm + 1], l[3 * m + j]));
                        J[2][j] = geom::add(J[2][j], geom::mul(x[3 * mm + 2], l[3 * m + j]));
                    }
                }
            }
            double Ji[3][3];
            const double det = geom::invert(J, Ji);
            if (!valid) { // lanes without a Gauss point: zero gradients (their row of L is zero, so J is singular)
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < 3; j++) Ji[i][j] = 0.0;
            }
            if (valid && fabs(det) <= ZERO_DETERMINANT) atomicMin(p.err_cell, orig);
            if constexpr (!GEN) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the previous bulk copy has left the region
            }
            __syncwarp();
            if (lane < 4 * KS) {
                // the padding Gauss points of the last k-step are zero-filled (the region held the previous element block)
                double *Vq = R + lane;
#pragma unroll
                for (int m = 0; m < NN; m++)
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const double b = fma(l[3 * m + 2], Ji[2][j], fma(l[3 * m + 1], Ji[1][j], l[3 * m] * Ji[0][j]));
                        Vq[(3 * m + j) * PS] = b;
                    }
                cc[lane] = valid ? det * s_w[lane] * p.alpha : 0.0;
            }
        }
        __syncwarp();

        // ---- phase C: P = V^T diag(c) V on the tensor cores -----------------------------------------------------------------
        {
            double acc[NTILE][2];
#pragma unroll
            for (int t = 0; t < NTILE; t++) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll 1
            for (int ks = 0; ks < KS; ks++) {
                const int row = 4 * ks + ak;
                const double c = cc[row];
                const double *vr = R + ar * PS + row;
                double v[NT8], vs[NT8];
#pragma unroll
                for (int t = 0; t < NT8; t++) {
                    v[t] = vr[8 * t * PS];
                    vs[t] = v[t] * c;
                }
                // C[r][q] += sum_k A[r][k] B[k][q] with A[r][k] = V[k][8J + r], B[k][q] = c_k V[k][8I + q]: C[r][q] = P(8I + q, 8J + r)
                static_for<NTILE>([&](auto tc) {
                    constexpr int t = decltype(tc)::value, I = syrk_tile_i(NT8, t), J = syrk_tile_j(NT8, t);
                    dmma_m8n8k4(acc[t][0], acc[t][1], v[J], vs[I]);
                });
            }
            __syncwarp(); // every lane has finished reading V: the region becomes P
            // the lane holds P(8I + 2ak + e, 8J + ar), e = 0, 1.  Two 8-byte stores; odd fragment rows store e = 1 first: at
            // (8I + 2ak + e) + 60 (8J + ar) = 2ak + e + 12 ar (mod 16) the 16 lanes of a half-warp then hit 16 distinct bank pairs
            // (one 16-byte store per lane would put rows ar and ar + 1 of a quarter-warp on the same banks)
            const int e0 = ar & 1;
            static_for<NTILE>([&](auto tc) {
                constexpr int t = decltype(tc)::value, I = syrk_tile_i(NT8, t), J = syrk_tile_j(NT8, t);
                const int prow = 8 * I + 2 * ak; // rows prow, prow + 1
                const int pcol = 8 * J + ar;
                if (prow < NDOF && pcol < NDOF) {
                    double *dst = R + prow + NDOF * pcol;
                    dst[e0] = select_f64(e0, acc[t][1], acc[t][0]);
                    dst[1 - e0] = select_f64(e0, acc[t][0], acc[t][1]);
                }
            });
        }
        __syncwarp();

        // ---- phase D: K(m,n) = D' : P(m,n), in place, and its mirror ------------------------------------------------------
        // Two passes at a time: the blocks of both are read before the first is written back (the pairs of different passes
        // are disjoint, which the compiler cannot know), so the loads of one pass do not queue behind the stores of the previous.
        {
            // the modulus entries the pattern uses, in registers for the whole phase (the accumulators are dead by now)
            double dl[36];
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int b = 0; b < 6; b++)
                    if (d_nz3(PAT, a, b)) dl[a * 6 + b] = DREC ? Dc[a * 6 + b] : cst.d[a * 6 + b];
#pragma unroll
            for (int k0 = 0; k0 < NPASS; k0 += 2) {
                double P[2][3][3];
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    if (k0 + kk >= NPASS || pair_off[k0 + kk] == 0xffffffffu) continue;
                    const unsigned int od = pair_off[k0 + kk] & 0xffffu, om = pair_off[k0 + kk] >> 16;
                    const bool dg = od == om;
                    // P(m,n)[a][b] at od + a + NDOF b; only the upper triangle of P was computed, so a diagonal pair reads its
                    // upper half and mirrors it
                    const double *Pb = R + od;
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int b = 0; b < 3; b++) P[kk][a][b] = a > b ? Pb[dg ? b + NDOF * a : a + NDOF * b] : Pb[a + NDOF * b];
                }
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    if (k0 + kk >= NPASS || pair_off[k0 + kk] == 0xffffffffu) continue;
                    const unsigned int od = pair_off[k0 + kk] & 0xffffu, om = pair_off[k0 + kk] >> 16;
                    const bool dg = od == om;
                    double *Kd = R + od, *Km = R + om;
#pragma unroll
                    for (int j = 0; j < 3; j++)
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            // K(3m+i, 3n+j) = sum_qr d'[row(i,q)][row(j,r)] P[comp(i,q)][comp(j,r)]; for a non-symmetric D the mirrored
                            // block K(3n+j, 3m+i) = sum_qr d'[row(j,r)][row(i,q)] P[comp(i,q)][comp(j,r)]  (P(n,m) = P(m,n)^T)
                            double v = 0.0, u = 0.0;
                            bool have = false;
#pragma unroll
                            for (int qq = 0; qq < 3; qq++)
#pragma unroll
                                for (int r = 0; r < 3; r++) {
                                    const int ra = col_row3(i, qq), rb = col_row3(j, r);
                                    if (!d_nz3(PAT, ra, rb)) continue;
                                    const double pv = P[kk][col_comp3(i, qq)][col_comp3(j, r)];
                                    v = have ? fma(dl[ra * 6 + rb], pv, v) : dl[ra * 6 + rb] * pv;
                                    if (!SYM) u = have ? fma(dl[rb * 6 + ra], pv, u) : dl[rb * 6 + ra] * pv;
                                    have = true;
                                }
                            Kd[i + NDOF * j] = v;
                            if (!dg) Km[j + NDOF * i] = SYM ? v : u;
                        }
                }
            }
        }

        // ---- ship --------------------------------------------------------------------------------------------------------
        if constexpr (!GEN) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                double *dst = p.out + (p.out_off ? p.out_off[cell] - p.out_base : (orig - p.cell_begin) * (unsigned long long)RS);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(R)),
                             "r"((unsigned)(RS * sizeof(double))), "l"(pol)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            // element block nrow x ncol with the integrated ndof x ndof window at (ii0, jj0): clear = true overwrites the whole
            // block (kk.fill(0.0), mat_10_bdb.rs:141-143), clear = false adds the window
            __syncwarp();
            const unsigned int nrow = p.nrow, ncol = p.ncol;
            double *dst = p.out + (p.out_off ? p.out_off[cell] - p.out_base : (orig - p.cell_begin) * ((unsigned long long)nrow * ncol));
            for (unsigned int idx = lane; idx < nrow * ncol; idx += 32) {
                const unsigned int jj = idx / nrow, ii = idx - jj * nrow;
                const bool in = ii >= p.ii0 && ii < p.ii0 + NDOF && jj >= p.jj0 && jj < p.jj0 + NDOF;
                const double val = in ? R[(ii - p.ii0) + NDOF * (jj - p.jj0)] : 0.0;
                if (p.clear) dst[idx] = val;
                else if (in) dst[idx] += val;
            }
            __syncwarp();
        }
    }
    if constexpr (!GEN) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int NN>
size_t syrk_smem(int ng, bool drec) {
    using T = SyrkDims<NN>;
    (void)ng;
    return sizeof(double) * ((size_t)T::SLOTS * T::RS + (size_t)T::MAXNG * T::LS + (size_t)T::MAXNG + (size_t)T::SLOTS * T::NDOF +
                             (size_t)T::SLOTS * T::MAXNG + (drec ? (size_t)T::SLOTS * 36 : 0));
}

template <int NN, bool GEN>
int launch_syrk_kind(const IntegParams &p, const SyrkConsts &cst, int pat, bool drec, void *stream) {
    using T = SyrkDims<NN>;
    if (p.ng > T::MAXNG || p.work_counter == nullptr) return 0;
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
    unsigned long long grid = (unsigned long long)nsm * T::CTAS_PER_SM; // persistent: the slots draw cells from a counter
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
        if (cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), (cudaStream_t)stream) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, T::THREADS, smem, (cudaStream_t)stream>>>(q, cst);
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
    // Tet10: measured 448 M el/s here against 263 M el/s on gl_kernels_bdb.cu (750 000 cells, profiles/r02_tet10_dmma_vs_register_blocked.txt)
    if (p.nnode == 10 && tuning_env("GL_DISABLE_SYRK_TET10")) return 0;
    const unsigned ndof = (unsigned)p.nnode * 3u;
    const bool tight = p.nrow == ndof && p.ncol == ndof && p.ii0 == 0 && p.jj0 == 0;
    const bool gen = !(tight && p.clear && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

    SyrkConsts cst;
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
