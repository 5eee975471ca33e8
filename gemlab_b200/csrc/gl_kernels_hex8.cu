// gl_kernels_hex8.cu -- register-blocked Hex8 stiffness kernel: integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-208) for
// GeoKind::Hex8 with the 8-point rule, a uniform modulus D, tight element blocks and clear = true  (BASELINE.json
// configs[1], the headline workload: 4,608 B of output and ~37 kflop per cell -- the ridge of the B200 roofline).
//
// Mapping: EIGHT LANES PER CELL, four cells per warp, no block-level barrier anywhere.
//   phase 0  lane g gathers node g of its cell (Mesh::set_pad, src/mesh/mesh.rs:448-456) -- prefetched one batch ahead;
//   phase 1  lane g owns GAUSS POINT g: J = Xt.L, closed-form inverse/det, B = L.inv(J)
//            (scratchpad_calc_jacobian.rs:101-131, scratchpad_calc_gradient.rs:78-91); B and c = det*w*alpha go to smem;
//   phase 2  lane n owns COLUMN NODE n and accumulates GEOMETRIC TENSORS  P(m,n)[k][l] = sum_p c_p B(m,k) B(n,l)
//            (9 DFMA per node pair and Gauss point) for the "circulant half" m = n..n+4 (mod 8); P(n,m) = P(m,n)^T covers
//            the other pairs.  D is uniform, so it is applied ONCE, after the Gauss loop:
//            K(m,n)[i][j] = sum_kl D_ikjl P(m,n)[k][l]  (add_to_kk, mat_10_bdb.rs:179-208, with the sum over Gauss
//            points moved inside) -- 2.5x fewer FP64 instructions than accumulating M^T.D.M per Gauss point;
//   phase 3  the contraction with D' (from the constant bank; structural zeros of an isotropic / orthotropic D' skipped at
//            compile time), blocks staged in smem in the final column-major layout (mirrored when D is major-symmetric);
//            every cell (4,608 contiguous bytes) leaves the SM as one bulk async copy (cp.async.bulk.global.shared::cta ->
//            UBLKCP, L2 evict-first), issued by one elected lane per warp, so HBM sees only full lines and no thread
//            waits on a store.
// The kernel is bound by the WRITE stream: with all arithmetic removed it runs as fast as with it (GL_HEX8_DEBUG=3), and
// the 0.7 GB of gather reads cost more than their share because they interleave with 36.9 GB of writes (GL_HEX8_DEBUG=4).
// Shared-memory strides (25 / 200 / 584 doubles) are chosen so that every 64-bit access of a half-warp (2 cells x 8
// lanes) hits 16 distinct bank pairs in all phases.
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int H8_THREADS = 128;
constexpr int H8_CELLS = H8_THREADS / 8; // cells per CTA batch
constexpr int GP_STRIDE = 25;            // doubles per (cell, gauss point): 24 gradient values + coefficient
constexpr int CELL_STRIDE = 8 * GP_STRIDE; // 200 = 8 (mod 16): second cell of a half-warp lands 16 banks away
constexpr int X_STRIDE = 25;
constexpr int K_STRIDE = 584;            // 4,608-byte element block + 64 bytes: = 8 (mod 16), so the two cells of a half-warp stage conflict-free
constexpr int L_STRIDE = 25;

struct Hex8Consts {
    int flags;    // bit 0: blocked work distribution, bit 1: evict-first stores, bit 2: evict-last coordinate loads
    double d[36]; // d'[a][b] = f_a f_b d[a][b], f = (1,1,1,1/sqrt2,1/sqrt2,1/sqrt2): the Mandel factors folded into D
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// one lane of the (converged) warp, chosen by the hardware: lets ptxas keep the guarded code on the uniform datapath
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// structure of the (Mandel-scaled) modulus d', decided on the host from the coefficient record:
//   PAT 0  general (minor-symmetric only): all 8 row blocks per lane
//   PAT 1  major-symmetric d' = d'^T      : K is symmetric -> circulant half + mirror
//   PAT 2  PAT 1 and d' = [A 0; 0 diag]   : normal/shear uncoupled with a diagonal shear block (isotropic and
//          axis-aligned orthotropic elasticity): the structural zeros of T = d'.M are skipped, which is exact
//          (the skipped terms are products with an exact 0)
__host__ __device__ constexpr bool d_nz(int pat, int a, int b) { return pat < 2 ? true : ((a < 3 && b < 3) || a == b); }
// rows of the strain-displacement column (n, j) that are non-zero, and which gradient component sits there
__host__ __device__ constexpr int col_row(int j, int t) { return j == 0 ? (t == 0 ? 0 : (t == 1 ? 3 : 5)) : (j == 1 ? (t == 0 ? 1 : (t == 1 ? 3 : 4)) : (t == 0 ? 2 : (t == 1 ? 4 : 5))); }
__host__ __device__ constexpr int col_comp(int j, int t) { return j == 0 ? (t == 0 ? 0 : (t == 1 ? 1 : 2)) : (j == 1 ? (t == 0 ? 1 : (t == 1 ? 0 : 2)) : (t == 0 ? 2 : (t == 1 ? 1 : 0))); }

template <int PAT, int PF, int DBG = 0, bool DREC = false, bool FD = false, bool SC = false>
__global__ void __launch_bounds__(H8_THREADS, 2) hex8_bdb_kernel(const IntegParams p, const Hex8Consts cst) {
    if (DREC && p.coef_pat_dev != nullptr && *p.coef_pat_dev != PAT) return; // another pattern variant owns this call
    constexpr bool SYM = PAT >= 1;
    extern __shared__ __align__(16) double smem[];
    double *s_L = smem;                         // [8][L_STRIDE]  L[gp][m*3+j]
    double *s_w = s_L + 8 * L_STRIDE;           // [8]
    double *s_X = s_w + 8;                      // [H8_CELLS][X_STRIDE]
    double *s_B = s_X + H8_CELLS * X_STRIDE;    // [H8_CELLS][CELL_STRIDE]
    double *s_K = s_B + H8_CELLS * CELL_STRIDE; // [H8_CELLS][K_STRIDE]  (16-byte aligned: offset is a multiple of 2 doubles)
    double *s_D = s_K + H8_CELLS * K_STRIDE;    // [H8_CELLS][36]  per-cell d' (DREC only: PER_MARKER / PER_CELL moduli)
    double *s_S = s_D + (DREC ? H8_CELLS * 36 : 0); // [H8_CELLS][8][6] stress tensors of the Gauss points (FD only)

    const int tid = threadIdx.x;
    const int g = tid & 7;   // lane within the cell group
    const int cs = tid >> 3; // cell slot in the batch
    const int wq = __shfl_sync(0xffffffffu, tid >> 5, 0); // warp index, known to the compiler as warp-uniform

    // reference-element table: w[8] | N[8][8] | L[8][8][3]
    for (int i = tid; i < 8 * 24; i += H8_THREADS) s_L[(i / 24) * L_STRIDE + (i % 24)] = p.rule[8 + 64 + i];
    if (tid < 8) s_w[tid] = p.rule[tid];
    __syncthreads();

    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + H8_CELLS - 1) / H8_CELLS;
    double *Xc = s_X + cs * X_STRIDE;
    double *Bc = s_B + cs * CELL_STRIDE;
    double *Kc = s_K + cs * K_STRIDE;
    double *Dc = s_D + cs * 36;
    double *Sc = s_S + cs * 48;
    // d'[a][b] of this lane's cell: the constant bank for a uniform modulus, shared memory for a modulus record per cell
    double dl[36] = {}; // DREC: register copy of the entries the pattern uses, refreshed once per batch (the staging stores in
                   // between would otherwise force a shared-memory load per term)
    auto dval = [&](int a, int b) -> double { return DREC ? dl[a * 6 + b] : cst.d[a * 6 + b]; };
    const double *Lg = s_L + g * L_STRIDE;

    // work distribution: iteration `it` of this CTA handles batch_of(it); strided (blockIdx + it*grid) or blocked
    // (every CTA streams through one contiguous slab of the output)
    const unsigned long long nit = (nbatch + gridDim.x - 1) / gridDim.x;
    const bool blocked = (cst.flags & 1) != 0;
    auto batch_of = [&](unsigned long long it) -> unsigned long long {
        if (it >= nit) return nbatch;
        return blocked ? (unsigned long long)blockIdx.x * nit + it : (unsigned long long)blockIdx.x + it * gridDim.x;
    };
    // L2 policies: the element matrices are written once and never read here (evict-first keeps them from flushing the
    // mesh out of L2); point coordinates are re-read by up to 8 cells, the furthest a whole z-layer later (evict-last)
    unsigned long long pol_w, pol_r;
    if (cst.flags & 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_w));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_w));
    if (cst.flags & 4) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_r));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_r));
    auto ldx = [&](const double *ptr) -> double {
        if (DBG == 5) return (double)(unsigned long long)ptr; // experiment: no coordinate loads
        double v;
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(ptr), "l"(pol_r));
        return v;
    };

    auto ldp = [&](unsigned int pid, double &x, double &y, double &z) {
        x = ldx(p.coords + pid);
        y = ldx(p.coords + p.npoint + pid);
        z = ldx(p.coords + 2 * p.npoint + pid);
    };

    // software-pipelined gather (Mesh::set_pad): coordinates are loaded PF batches ahead and the point id they depend
    // on PF+1 batches ahead.  Under a saturated write stream a read can take several microseconds -- longer than one
    // batch -- so one batch of lookahead is not enough to keep the loads off the critical path.
    auto cell_of = [&](unsigned long long it) -> unsigned long long { // cell of this lane in iteration it, or p.hi
        const unsigned long long bb = batch_of(it);
        if (DBG == 4 || bb >= nbatch) return p.hi;
        const unsigned long long c = p.lo + bb * H8_CELLS + cs;
        return c < p.hi ? c : p.hi;
    };
    auto ldpid = [&](unsigned long long c) -> unsigned int {
        return p.conn[(unsigned long long)g * p.ncell_k + c];
    };
    double qx[PF], qy[PF], qz[PF];
    unsigned int pid_next = 0;
#pragma unroll
    for (int d = 0; d < PF; d++) {
        qx[d] = qy[d] = qz[d] = 0.0;
        const unsigned long long c = cell_of(d);
        if (c < p.hi) ldp(ldpid(c), qx[d], qy[d], qz[d]);
    }
    {
        const unsigned long long c = cell_of(PF);
        if (c < p.hi) pid_next = ldpid(c);
    }

    for (unsigned long long it = 0; it < nit; it++) {
        const unsigned long long batch = batch_of(it);
        if (batch >= nbatch) break;
        const unsigned long long cell = p.lo + batch * H8_CELLS + cs;
        const bool active = cell < p.hi;

        // ---- phase 0: publish this batch's coordinates, prefetch later batches' ------------------------------------
        Xc[g * 3 + 0] = qx[0];
        Xc[g * 3 + 1] = qy[0];
        Xc[g * 3 + 2] = qz[0];
#pragma unroll
        for (int d = 0; d + 1 < PF; d++) {
            qx[d] = qx[d + 1];
            qy[d] = qy[d + 1];
            qz[d] = qz[d + 1];
        }
        {
            const unsigned long long c1 = cell_of(it + PF), c2 = cell_of(it + PF + 1);
            if (c1 < p.hi) ldp(pid_next, qx[PF - 1], qy[PF - 1], qz[PF - 1]); // its point id arrived during the previous iteration
            if (c2 < p.hi) pid_next = ldpid(c2);
        }
        __syncwarp();

        // ---- phase 1: geometry at Gauss point g ------------------------------------------------------------------
        if (DBG < 2 || DBG > 4) {
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int m = 0; m < 8; m++) geom::jacobian_add_node<3, 3>(J, Xc + m * 3, Lg + m * 3);
            double Ji[3][3];
            const double det = geom::invert(J, Ji);
            if (active && fabs(det) <= ZERO_DETERMINANT) atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
            double *Bg = Bc + g * GP_STRIDE;
#pragma unroll
            for (int m = 0; m < 8; m++) geom::gradient_row<3>(Lg + m * 3, Ji, Bg + m * 3);
            Bg[24] = det * s_w[g] * p.alpha;
        }
        if constexpr (FD) {
            // stress at Gauss point g of this cell (Mandel vector, vec_04_bt's PER_GAUSS record): off-diagonals become plain
            // tensor components sigma_01, sigma_12, sigma_02
            if (active) {
                const unsigned long long rec = ((p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.coef_begin) * 8ull + g;
                const double *t = p.sigma + rec * 6ull;
                const double hs = 0.70710678118654752440084436210484903928;
#pragma unroll
                for (int a = 0; a < 6; a++) Sc[g * 6 + a] = a < 3 ? t[a] : t[a] * hs;
            }
        }
        if constexpr (DREC) {
            // the cell's modulus record (ns x ns Mandel matrix, column-major): the record of its marker (PER_MARKER) or its own
            // (PER_CELL); the 8 lanes of the cell fold the 1/sqrt(2) factors in and publish d' for the epilogue
            if (active) {
                const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell]
                                                            : (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.coef_begin;
                const double *D = p.coef + rec * p.coef_cell_stride;
                const double hs = 0.70710678118654752440084436210484903928;
                for (int idx = g; idx < 36; idx += 8) {
                    const int a = idx / 6, b = idx - a * 6;
                    Dc[idx] = D[a + b * 6] * (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
                }
            }
        }
        __syncwarp();

        // ---- phase 2: lane g = column node n; geometric tensors P(m,n)[k][l] = sum_p c_p B(m,k) B(n,l) ------------------
        // D is the same at every Gauss point of the cell, so K(m,n)[i][j] = sum_kl D_ikjl P(m,n)[k][l]: the 9 FMA per node
        // pair and Gauss point below replace the 27 of the T = D'.M / M^T.T form, and D enters once, in the epilogue.
        // P(n,m) = P(m,n)^T, so the circulant half m = n .. n+4 (mod 8) covers every pair; P(n,n) is symmetric (6 values).
        double p0[6], pb[4][9];
#pragma unroll
        for (int q = 0; q < 6; q++) p0[q] = 0.0;
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int q = 0; q < 9; q++) pb[b][q] = 0.0;

        double dn[3] = {0.0, 0.0, 0.0}; // FD: internal-force entries of node n, d(n,i) = sum_p c_p sum_a sigma_ia B(n,a)
#pragma unroll 1
        for (int gp = 0; gp < (DBG >= 1 && DBG <= 4 ? 0 : 8); gp++) {
            const double *Bg = Bc + gp * GP_STRIDE;
            const double c = Bg[24];
            const double bu[3] = {Bg[g * 3 + 0], Bg[g * 3 + 1], Bg[g * 3 + 2]};
            const double vn[3] = {bu[0] * c, bu[1] * c, bu[2] * c};
            if constexpr (FD) { // integ::vec_04_bt (src/integ/vec_04_bt.rs:93-170), 3D branch
                const double *S = Sc + gp * 6;
                dn[0] = fma(S[5], vn[2], fma(S[3], vn[1], fma(S[0], vn[0], dn[0])));
                dn[1] = fma(S[4], vn[2], fma(S[1], vn[1], fma(S[3], vn[0], dn[1])));
                dn[2] = fma(S[2], vn[2], fma(S[4], vn[1], fma(S[5], vn[0], dn[2])));
            }
            p0[0] = fma(bu[0], vn[0], p0[0]);
            p0[1] = fma(bu[0], vn[1], p0[1]);
            p0[2] = fma(bu[0], vn[2], p0[2]);
            p0[3] = fma(bu[1], vn[1], p0[3]);
            p0[4] = fma(bu[1], vn[2], p0[4]);
            p0[5] = fma(bu[2], vn[2], p0[5]);
#pragma unroll
            for (int b = 1; b < 5; b++) {
                const int m = (g + b) & 7;
                const double bm[3] = {Bg[m * 3 + 0], Bg[m * 3 + 1], Bg[m * 3 + 2]};
#pragma unroll
                for (int l = 0; l < 3; l++)
#pragma unroll
                    for (int k = 0; k < 3; k++) pb[b - 1][k + 3 * l] = fma(bm[k], vn[l], pb[b - 1][k + 3 * l]);
            }
        }

        if constexpr (DREC) {
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int b = 0; b < 6; b++)
                    if (d_nz(PAT, a, b)) dl[a * 6 + b] = Dc[a * 6 + b];
        }
        if constexpr (FD) {
            if (active) { // the 8 lanes of a cell write its 24 contiguous doubles
                double *d = p.out_d + ((p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.cell_begin) * 24ull + 3 * g;
                d[0] = dn[0];
                d[1] = dn[1];
                d[2] = dn[2];
            }
        }
        // SC: position-map slots of the warp's next cell to scatter (lane l owns entries l, l + 32, .. of the row-major block)
        uint32_t sc_next[SC ? 18 : 1];
        auto sc_load = [&](unsigned long long wc) {
            if constexpr (SC) {
                if (wc < p.hi) {
                    const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[wc] : wc;
                    const uint32_t *pos = p.scatter_pos + (orig - p.cell_begin) * 576ull + (tid & 31);
#pragma unroll
                    for (int k = 0; k < 18; k++) sc_next[k] = __ldcs(pos + 32 * k);
                } else {
#pragma unroll
                    for (int k = 0; k < 18; k++) sc_next[k] = 0xFFFFFFFFu;
                }
            }
        };
        if constexpr (SC) sc_load(p.lo + batch * H8_CELLS + wq * 4);
        // ---- phase 3: contract with D', stage the blocks column-major, ship every cell with one bulk copy -------------
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the warp's previous copies have left s_K (no-op for lanes that issued none)
        __syncwarp();
        // SC (fused assembly): the block is staged TRANSPOSED (row-major), so that consecutive lanes of the scatter sweep walk along
        // one global row -- the three dofs of a node are adjacent columns of the CSR row
        auto kidx = [](int r, int c) -> int { return SC ? c + 24 * r : r + 24 * c; };
#pragma unroll
        for (int b = 0; b < (DBG >= 3 && DBG <= 4 ? 0 : 5); b++) {
            const int m = (g + b) & 7;
            // P(m,n)[k][l] of this block (compile-time indices after unrolling)
            auto P = [&](int k, int l) -> double {
                if (b == 0) return p0[k <= l ? (k == 0 ? l : (k == 1 ? 2 + l : 5)) : (l == 0 ? k : (l == 1 ? 2 + k : 5))];
                return pb[b == 0 ? 0 : b - 1][k + 3 * l];
            };
            // In the symmetric variants the diagonal block (b = 0) is itself symmetric, and the block at distance 4 is
            // held as a transposed duplicate by lane g+4: both need only their upper triangles (i <= j).
            const bool tri = SYM && (b == 0 || b == 4);
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (tri && i > j) continue;
                    // K(3m+i, 3n+j) = sum_qr d'[row(i,q)][row(j,r)] P[comp(i,q)][comp(j,r)]
                    double v = 0.0;
                    bool have = false;
#pragma unroll
                    for (int q = 0; q < 3; q++)
#pragma unroll
                        for (int r = 0; r < 3; r++)
                            if (d_nz(PAT, col_row(i, q), col_row(j, r))) {
                                const double dv = dval(col_row(i, q), col_row(j, r)), pv = P(col_comp(i, q), col_comp(j, r));
                                v = have ? fma(dv, pv, v) : dv * pv;
                                have = true;
                            }
                    Kc[kidx(3 * m + i, 3 * g + j)] = v;
                    // mirrored entry K(n,m)^T: whole blocks at distance 1..3, strict upper triangles of the b = 0 / b = 4 blocks
                    if (SYM && ((b >= 1 && b <= 3) || (tri && i < j))) Kc[kidx(3 * g + j, 3 * m + i)] = v;
                    if (!SYM && b >= 1 && b <= 3) {
                        // general D: K(3n+j, 3m+i) = sum_qr d'[row(j,r)][row(i,q)] P(n,m)[comp(j,r)][comp(i,q)], P(n,m) = P(m,n)^T
                        double u = 0.0;
                        bool hv = false;
#pragma unroll
                        for (int q = 0; q < 3; q++)
#pragma unroll
                            for (int r = 0; r < 3; r++) {
                                const double dv = dval(col_row(j, r), col_row(i, q)), pv = P(col_comp(i, q), col_comp(j, r));
                                u = hv ? fma(dv, pv, u) : dv * pv;
                                hv = true;
                            }
                        Kc[kidx(3 * g + j, 3 * m + i)] = u;
                    }
                }
        }
        if constexpr (SC) {
            // fused assembly: the warp sweeps its four staged blocks and adds every entry to its slot of the global matrix (RED.ADD.F64);
            // the element matrices never reach HBM.  The 18 slots a lane needs per cell are loaded as one batch, one cell ahead of
            // the adds (the first batch was issued before the contraction), so that no RED waits on the position map.
            __syncwarp();
            const unsigned long long wcell = p.lo + batch * H8_CELLS + wq * 4;
            const int lane = tid & 31;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t u[18];
#pragma unroll
                for (int k = 0; k < 18; k++) u[k] = sc_next[k];
                if (c + 1 < 4) sc_load(wcell + c + 1);
                if (wcell + c < p.hi) {
                    const double *src = s_K + (wq * 4 + c) * K_STRIDE + lane;
#pragma unroll
                    for (int k = 0; k < 18; k++)
                        if (u[k] != 0xFFFFFFFFu) atomicAdd(p.scatter_vals + u[k], src[32 * k]);
                }
            }
            __syncwarp();
            continue;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        // one elected lane issues the warp's four copies back to back; every operand is derived from warp-uniform values
        // (wq comes from a lane-0 shuffle) so that the copies are plain UBLKCP instructions on uniform registers
        {
            const unsigned long long wcell = p.lo + batch * H8_CELLS + wq * 4; // first cell of this warp
            if (wcell < p.hi) {
                const unsigned long long remaining = p.hi - wcell;
                const int nc = remaining < 4 ? (int)remaining : 4;
                const unsigned src0 = smem_u32(s_K + wq * 4 * K_STRIDE);
                if (elect_one()) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        if (c < nc) {
                            const unsigned long long orig = p.cell_ids ? (unsigned long long)p.cell_ids[wcell + c] : wcell + c;
                            double *dst = p.out + (orig - p.cell_begin) * 576ull;
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
                                         "r"(src0 + c * K_STRIDE * 8), "r"(4608), "l"(pol_w)
                                         : "memory");
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

constexpr size_t H8_SMEM_BASE = sizeof(double) * (8 * L_STRIDE + 8 + H8_CELLS * X_STRIDE + H8_CELLS * CELL_STRIDE + H8_CELLS * K_STRIDE);
constexpr size_t H8_SMEM_DREC = sizeof(double) * (H8_CELLS * 36); // per-cell moduli
constexpr size_t H8_SMEM_FD = sizeof(double) * (H8_CELLS * 48);   // per-Gauss-point stresses

} // namespace

int launch_hex8_bdb(const IntegParams &p, void *stream) {
    // covered: Hex8, 8-point rule, D uniform or one record per marker / per cell, tight 24x24 blocks, overwrite, homogeneous
    // layout, 16-byte aligned output
    if (p.nnode != 8 || p.ng != 8 || p.op != GL_OP_MAT_10_BDB) return 0;
    const bool drec = p.coef != nullptr;
    const bool fd = p.sigma != nullptr && p.out_d != nullptr; // fused internal-force vector (gl_integ_fused)
    if (drec && (p.coef_gp_stride != 0 || p.coef_cell_stride != 36)) return 0; // PER_GAUSS moduli: generic kernel
    const bool sc = p.scatter_pos != nullptr; // fused assembly: no element blocks are written
    if (sc && fd) return 0;
    if (p.axisym || (!p.clear && !sc) || p.out_off != nullptr) return 0;
    if (p.nrow != 24 || p.ncol != 24 || p.ii0 != 0 || p.jj0 != 0 || p.size_r != 24) return 0;
    if (!sc && (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return 0;

    Hex8Consts cst;
    static int flags = -1;
    if (flags < 0) {
        const char *e = tuning_env("GL_HEX8_FLAGS"); // tuning knob, see Hex8Consts::flags
        flags = e ? atoi(e) : 2;
    }
    cst.flags = flags;
    const double h = 0.70710678118654752440084436210484903928;
    bool sym = true;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            const double fa = a < 3 ? 1.0 : h, fb = b < 3 ? 1.0 : h;
            cst.d[a * 6 + b] = p.coef_uniform[a + b * 6] * fa * fb; // ABI record is column-major
            if (p.coef_uniform[a + b * 6] != p.coef_uniform[b + a * 6]) sym = false;
        }
    // gather lookahead in batches (tuning knob GL_HEX8_PF = 1 or 2)
    static int pf = -1;
    if (pf < 0) {
        const char *e = tuning_env("GL_HEX8_PF");
        pf = e ? atoi(e) : 1;
        if (pf < 1 || pf > 2) pf = 1;
    }
    int pat = 0;
    if (sym) {
        pat = 2;
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++)
                if (!d_nz(2, a, b) && cst.d[a * 6 + b] != 0.0) pat = 1;
    }
    if (drec) pat = p.coef_pat; // structure common to all records (0 when the records live on the device only)
    {
        const char *e = tuning_env("GL_HEX8_PATTERN"); // tuning/testing knob: cap the specialisation level
        if (e && atoi(e) < pat) pat = atoi(e) < 0 ? 0 : atoi(e);
    }
    typedef void (*kern_t)(const IntegParams, const Hex8Consts);
    auto pick = [&](auto pfc, int pat) -> kern_t {
        constexpr int P = decltype(pfc)::value;
#ifdef GL_TUNING
        if (const char *e = tuning_env("GL_HEX8_DEBUG")) { // timing experiments only (1 no accumulation, 2 no geometry, 3 no staging, 4 no gather, 5 no coordinate loads)
            switch (atoi(e)) {
            case 1: return hex8_bdb_kernel<2, P, 1>;
            case 2: return hex8_bdb_kernel<2, P, 2>;
            case 3: return hex8_bdb_kernel<2, P, 3>;
            case 4: return hex8_bdb_kernel<2, P, 4>;
            default: return hex8_bdb_kernel<2, P, 5>;
            }
        }
#endif
        if (sc && drec) return pat == 2 ? hex8_bdb_kernel<2, P, 0, true, false, true> : (pat == 1 ? hex8_bdb_kernel<1, P, 0, true, false, true> : hex8_bdb_kernel<0, P, 0, true, false, true>);
        if (sc) return pat == 2 ? hex8_bdb_kernel<2, P, 0, false, false, true> : (pat == 1 ? hex8_bdb_kernel<1, P, 0, false, false, true> : hex8_bdb_kernel<0, P, 0, false, false, true>);
        if (fd && drec) return pat == 2 ? hex8_bdb_kernel<2, P, 0, true, true> : (pat == 1 ? hex8_bdb_kernel<1, P, 0, true, true> : hex8_bdb_kernel<0, P, 0, true, true>);
        if (fd) return pat == 2 ? hex8_bdb_kernel<2, P, 0, false, true> : (pat == 1 ? hex8_bdb_kernel<1, P, 0, false, true> : hex8_bdb_kernel<0, P, 0, false, true>);
        if (drec) return pat == 2 ? hex8_bdb_kernel<2, P, 0, true> : (pat == 1 ? hex8_bdb_kernel<1, P, 0, true> : hex8_bdb_kernel<0, P, 0, true>);
        return pat == 2 ? hex8_bdb_kernel<2, P> : (pat == 1 ? hex8_bdb_kernel<1, P> : hex8_bdb_kernel<0, P>);
    };
    const size_t H8_SMEM = H8_SMEM_BASE + (drec ? H8_SMEM_DREC : 0) + (fd ? H8_SMEM_FD : 0);
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long nbatch = (p.hi - p.lo + H8_CELLS - 1) / H8_CELLS;
    // Two CTAs are resident per SM (shared memory), but the grid is 32x larger: CTAs that start and finish at different
    // times keep the SMs out of lockstep, so gather reads and matrix writes of different SMs interleave instead of
    // arriving at the memory system in device-wide bursts (measured: 6.5 ms with 296 persistent CTAs, 5.8 ms with 9472).
    unsigned long long grid = (unsigned long long)nsm * 2 * 32;
    if (const char *e = tuning_env("GL_HEX8_GRID")) grid = (unsigned long long)atoi(e); // tuning knob
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    // records whose structure only the device knows: every variant is enqueued, the probe's verdict selects the one that runs
    const bool probe = drec && p.coef_pat_dev != nullptr && tuning_env("GL_HEX8_PATTERN") == nullptr && tuning_env("GL_HEX8_DEBUG") == nullptr;
    IntegParams q = p;
    if (!probe) q.coef_pat_dev = nullptr;
    for (int pt = probe ? 0 : pat; pt <= (probe ? 2 : pat); pt++) {
        kern_t kern = pf == 1 ? pick(std::integral_constant<int, 1>{}, pt) : pick(std::integral_constant<int, 2>{}, pt);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H8_SMEM) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, H8_THREADS, H8_SMEM, (cudaStream_t)stream>>>(q, cst);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("hex8_bdb<pat=%d%s%s%s>", pat, drec ? ",rec" : "", fd ? ",fd" : "", sc ? ",scatter" : "");
    return 1;
}

} // namespace gl
