// gl_kernels_hex8pg.cu -- Hex8 stiffness with ONE MODULUS PER GAUSS POINT: integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-208) when
// the closure `fn_dd(D, p, N, B)` returns a different tangent at every integration point (mat_10_bdb.rs:156-159) -- the case of
// every nonlinear material.  Coefficient mode GL_COEF_PER_GAUSS: record (e - cell_begin) * 8 + p is the 6 x 6 Mandel matrix of
// Gauss point p of cell e (column-major), on the device.
//
// With D_p varying inside the cell the Gauss sum cannot be moved inside the contraction (the geometric-tensor form of
// gl_kernels_hex8.cu), so the element blocks themselves are accumulated:
//      T_p(n)  = c_p D'_p M_p(n)            6 x 3, 54 DFMA per column node and Gauss point (M(n) has 9 non-zeros of 18)
//      K(m,n) += M_p(m)^T T_p(n)            3 x 3, 27 DFMA per node pair and Gauss point
// = 12.1 k DFMA per cell with the circulant half of the node pairs (symmetric D_p), 17.3 k with all pairs (general D_p).  A dense
// FP64 tensor-core tiling of K = A^T A' (A = stacked M_p, k = 48) needs 108 m8n8k4 DMMA = 27.6 k DFMA-equivalents on the same
// pipe (gl_bench_fp64_peak: DMMA and DFMA share it) because it cannot skip the structural zeros of M: the sparse DFMA form is used.
//
// Mapping (as gl_kernels_hex8.cu): EIGHT LANES PER CELL, four cells per warp, no block-level barrier.
//   phase 1  lane g = Gauss point g: geometry (J, inverse, B = L.inv(J)) and the modulus record D_g of its cell -- 18 LDG.128
//            issued before the geometry, consumed after it -- scaled by the Mandel factors into shared memory.  The records
//            ALIAS the element-block staging (they are dead before the blocks are staged), so the footprint is that of the
//            uniform kernel (two CTAs per SM).  The symmetric variant compares D_g with its transpose on the way;
//   phase 2  lane n = column node n: T_p(n) from broadcast shared-memory reads of D'_p, then the blocks K(m,n), m = n .. n+4
//            (mod 8) [symmetric D_p: the mirror gives the rest] or m = n .. n+7 [general D_p], in registers;
//   phase 3  blocks staged column-major, every cell (4,608 contiguous bytes) leaves as one bulk async copy (UBLKCP, L2
//            evict-first) issued by one elected lane per warp.
// Symmetry is decided ON THE DEVICE, without a probe pass: the symmetric variant lowers *pg_flag when a record differs from its
// transpose; the general variant is enqueued right behind it and returns at once unless that happened.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int PG_THREADS = 128;
constexpr int PG_CELLS = PG_THREADS / 8;
constexpr int GP_STRIDE = 25;              // doubles per (cell, gauss point): 24 gradient values + coefficient
constexpr int CELL_STRIDE = 8 * GP_STRIDE; // 200 = 8 (mod 16)
constexpr int X_STRIDE = 25;
constexpr int K_STRIDE = 584;              // 4,608-byte element block + 64 bytes
constexpr int L_STRIDE = 25;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// rows of the strain-displacement column (n, j) that are non-zero, and which gradient component sits there (calc_bb_matrix,
// mat_10_bdb.rs:238-278, Mandel order xx yy zz xy yz zx; the 1/sqrt(2) factors live in D')
__host__ __device__ constexpr int col_row(int j, int t) { return j == 0 ? (t == 0 ? 0 : (t == 1 ? 3 : 5)) : (j == 1 ? (t == 0 ? 1 : (t == 1 ? 3 : 4)) : (t == 0 ? 2 : (t == 1 ? 4 : 5))); }
__host__ __device__ constexpr int col_comp(int j, int t) { return j == 0 ? (t == 0 ? 0 : (t == 1 ? 1 : 2)) : (j == 1 ? (t == 0 ? 1 : (t == 1 ? 0 : 2)) : (t == 0 ? 2 : (t == 1 ? 1 : 0))); }

template <bool SYM>
__global__ void __launch_bounds__(PG_THREADS, 2) hex8_gauss_kernel(const IntegParams p) {
    if (!SYM && *p.pg_flag != 0) return; // every record was symmetric: the symmetric variant's result stands
    constexpr int NB = SYM ? 5 : 8;      // circulant distances handled by a lane
    extern __shared__ __align__(16) double smem[];
    double *s_L = smem;                         // [8][L_STRIDE]  L[gp][m*3+j]
    double *s_w = s_L + 8 * L_STRIDE;           // [8]
    double *s_X = s_w + 8;                      // [PG_CELLS][X_STRIDE]
    double *s_B = s_X + PG_CELLS * X_STRIDE;    // [PG_CELLS][CELL_STRIDE]
    double *s_K = s_B + PG_CELLS * CELL_STRIDE; // [PG_CELLS][K_STRIDE]; the first 8 x 36 (+4) doubles of a cell's slot hold d'_p during phases 1-2

    const int tid = threadIdx.x;
    const int g = tid & 7;
    const int cs = tid >> 3;
    const int wq = __shfl_sync(0xffffffffu, tid >> 5, 0);

    for (int i = tid; i < 8 * 24; i += PG_THREADS) s_L[(i / 24) * L_STRIDE + (i % 24)] = p.rule[8 + 64 + i];
    if (tid < 8) s_w[tid] = p.rule[tid];
    __syncthreads();

    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + PG_CELLS - 1) / PG_CELLS;
    double *Xc = s_X + cs * X_STRIDE;
    double *Bc = s_B + cs * CELL_STRIDE;
    double *Kc = s_K + cs * K_STRIDE;
    // the four cells of a warp read their d' records at the same instant: offsets of 0 / 64 / 32 / 96 bytes (mod 128) put the
    // four 16-byte broadcasts of one LDS.128 into distinct banks
    double *Dc = Kc + ((cs >> 1) & 1) * 4;
    const double *Lg = s_L + g * L_STRIDE;

    unsigned long long pol_w;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_w));

    // software-pipelined gather (Mesh::set_pad, src/mesh/mesh.rs:448-456): coordinates one batch ahead, point ids two
    auto cell_of = [&](unsigned long long it) -> unsigned long long {
        const unsigned long long bb = (unsigned long long)blockIdx.x + it * gridDim.x;
        if (bb >= nbatch) return p.hi;
        const unsigned long long c = p.lo + bb * PG_CELLS + cs;
        return c < p.hi ? c : p.hi;
    };
    auto ldpid = [&](unsigned long long c) -> unsigned int { return p.conn[(unsigned long long)g * p.ncell_k + c]; };
    auto ldp = [&](unsigned int pid, double &x, double &y, double &z) {
        x = p.coords[pid];
        y = p.coords[p.npoint + pid];
        z = p.coords[2 * p.npoint + pid];
    };
    double qx = 0.0, qy = 0.0, qz = 0.0;
    unsigned int pid_next = 0;
    {
        const unsigned long long c0 = cell_of(0), c1 = cell_of(1);
        if (c0 < p.hi) ldp(ldpid(c0), qx, qy, qz);
        if (c1 < p.hi) pid_next = ldpid(c1);
    }
    bool asym = false;
    // the modulus record of Gauss point g of this lane's cell (18 LDG.128), loaded ONE BATCH AHEAD: issued right after the
    // accumulation of the previous batch, in flight during its staging / copy issue and this batch's gather and geometry
    double2 dr[18];
    auto load_record = [&](unsigned long long c) {
        if (c < p.hi) {
            const unsigned long long rec = ((p.cell_ids ? (unsigned long long)p.cell_ids[c] : c) - p.coef_begin) * 8ull + g;
            const double2 *D = reinterpret_cast<const double2 *>(p.coef + rec * 36ull);
#pragma unroll
            for (int i = 0; i < 18; i++) dr[i] = __ldcs(D + i); // streamed once
        } else {
#pragma unroll
            for (int i = 0; i < 18; i++) dr[i] = make_double2(0.0, 0.0);
        }
    };
    load_record(cell_of(0));

    for (unsigned long long it = 0;; it++) {
        const unsigned long long batch = (unsigned long long)blockIdx.x + it * gridDim.x;
        if (batch >= nbatch) break;
        const unsigned long long cell = p.lo + batch * PG_CELLS + cs;
        const bool active = cell < p.hi;

        // ---- phase 0: publish this batch's coordinates, prefetch the next ---------------------------------------------
        Xc[g * 3 + 0] = qx;
        Xc[g * 3 + 1] = qy;
        Xc[g * 3 + 2] = qz;
        {
            const unsigned long long c1 = cell_of(it + 1), c2 = cell_of(it + 2);
            if (c1 < p.hi) ldp(pid_next, qx, qy, qz);
            if (c2 < p.hi) pid_next = ldpid(c2);
        }
        // ---- phase 1: geometry at Gauss point g; then the modulus record (already in flight) goes to shared memory ------------
        __syncwarp();
        {
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
        // the previous batch's bulk copies must have read s_K before the records overwrite it
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        {
            // column-major record -> row-major d'[a][b] = f_a f_b D[a][b]
            const double hs = 0.70710678118654752440084436210484903928;
            auto dval = [&](int a, int b) -> double {
                const int idx = a + 6 * b;
                return (idx & 1) ? dr[idx >> 1].y : dr[idx >> 1].x;
            };
            double *Dg = Dc + g * 36;
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    const double v = dval(a, b);
                    if (SYM && b > a && v != dval(b, a)) asym = true;
                    Dg[a * 6 + b] = v * ((a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs));
                }
        }
        __syncwarp();

        // ---- phase 2: lane g = column node n ---------------------------------------------------------------------------
        double acc[NB][9];
#pragma unroll
        for (int b = 0; b < NB; b++)
#pragma unroll
            for (int q = 0; q < 9; q++) acc[b][q] = 0.0;
#pragma unroll 1
        for (int gp = 0; gp < 8; gp++) {
            const double *Bg = Bc + gp * GP_STRIDE;
            const double c = Bg[24];
            const double vn[3] = {Bg[g * 3 + 0] * c, Bg[g * 3 + 1] * c, Bg[g * 3 + 2] * c};
            const double2 *Dp = reinterpret_cast<const double2 *>(Dc + gp * 36); // same address for the 8 lanes of the cell: broadcast
            double T[6][3];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                const double2 d01 = Dp[a * 3 + 0], d23 = Dp[a * 3 + 1], d45 = Dp[a * 3 + 2];
                const double d[6] = {d01.x, d01.y, d23.x, d23.y, d45.x, d45.y};
#pragma unroll
                for (int j = 0; j < 3; j++) T[a][j] = d[col_row(j, 0)] * vn[col_comp(j, 0)];
#pragma unroll
                for (int j = 0; j < 3; j++) T[a][j] = fma(d[col_row(j, 1)], vn[col_comp(j, 1)], T[a][j]);
#pragma unroll
                for (int j = 0; j < 3; j++) T[a][j] = fma(d[col_row(j, 2)], vn[col_comp(j, 2)], T[a][j]);
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const int m = (g + b) & 7;
                const double bm[3] = {Bg[m * 3 + 0], Bg[m * 3 + 1], Bg[m * 3 + 2]};
                // K(3m+i, 3n+j) += sum_t M(m)[row(i,t)][i] T[row(i,t)][j]   (add_to_kk, mat_10_bdb.rs:179-208)
                // symmetric D_p: the diagonal block (b = 0) is symmetric and the block at distance 4 is also held, transposed, by
                // lane g+4 -- both lanes compute only i <= j and the mirror of phase 3 supplies the rest
                const bool tri = SYM && (b == 0 || b == 4);
                // (t outermost: consecutive FMAs go to different accumulators)
#pragma unroll
                for (int t = 0; t < 3; t++)
#pragma unroll
                    for (int j = 0; j < 3; j++)
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            if (tri && i > j) continue;
                            acc[b][i + 3 * j] = fma(bm[col_comp(i, t)], T[col_row(i, t)][j], acc[b][i + 3 * j]);
                        }
            }
        }
        load_record(cell_of(it + 1)); // next batch's records
        __syncwarp(); // every lane of the warp has finished reading the records that alias the staging

        // ---- phase 3: stage column-major, one bulk copy per cell ----------------------------------------------------------
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int m = (g + b) & 7;
            const bool tri = SYM && (b == 0 || b == 4);
            const bool mirror = SYM && b >= 1 && b <= 3;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (tri && i > j) continue;
                    const double v = acc[b][i + 3 * j];
                    Kc[(3 * m + i) + 24 * (3 * g + j)] = v;
                    if (mirror || (tri && i < j)) Kc[(3 * g + j) + 24 * (3 * m + i)] = v;
                }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        {
            const unsigned long long wcell = p.lo + batch * PG_CELLS + wq * 4; // first cell of this warp
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
    if (SYM && asym) *p.pg_flag = 0;
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

constexpr size_t PG_SMEM = sizeof(double) * (8 * L_STRIDE + 8 + PG_CELLS * X_STRIDE + PG_CELLS * CELL_STRIDE + PG_CELLS * K_STRIDE);

__global__ void hex8_gauss_flag_reset_kernel(int *flag) { *flag = 1; }

} // namespace

int launch_hex8_gauss(const IntegParams &p, void *stream) {
    // covered: Hex8, 8-point rule, one 6 x 6 modulus per Gauss point (contiguous records, 16-byte aligned), tight 24 x 24 blocks,
    // overwrite, homogeneous layout, 16-byte aligned output.  Everything else with PER_GAUSS moduli: gl_kernels_bdb.cu.
    if (p.nnode != 8 || p.ng != 8 || p.op != GL_OP_MAT_10_BDB) return 0;
    if (p.coef == nullptr || p.coef_gp_stride != 36 || p.coef_cell_stride != 288 || p.coef_index != nullptr || p.pg_flag == nullptr) return 0;
    if (p.sigma != nullptr || p.axisym || !p.clear || p.out_off != nullptr) return 0;
    if (p.nrow != 24 || p.ncol != 24 || p.ii0 != 0 || p.jj0 != 0 || p.size_r != 24) return 0;
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.coef) & 15) != 0) return 0;
    if (tuning_env("GL_DISABLE_HEX8_GAUSS")) return 0; // testing knob

    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long nbatch = (p.hi - p.lo + PG_CELLS - 1) / PG_CELLS;
    unsigned long long grid = (unsigned long long)nsm * 2 * 4; // measured: 579 (x4), 576 (x8), 560 (x16), 542 (x32) M el/s on 10^6 cells
    if (const char *e = tuning_env("GL_HEX8_GAUSS_GRID")) grid = (unsigned long long)atoi(e); // tuning knob
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    if (cudaFuncSetAttribute(hex8_gauss_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PG_SMEM) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(hex8_gauss_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PG_SMEM) != cudaSuccess) return -1;
    hex8_gauss_flag_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p.pg_flag);
    hex8_gauss_kernel<true><<<(unsigned)grid, PG_THREADS, PG_SMEM, (cudaStream_t)stream>>>(p);
    hex8_gauss_kernel<false><<<(unsigned)grid, PG_THREADS, PG_SMEM, (cudaStream_t)stream>>>(p);
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("hex8_gauss<sym+general>");
    return 3;
}

} // namespace gl
