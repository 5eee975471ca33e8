// gl_kernels_tile.cu -- stiffness kernel for 20-node hexahedra: integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-208) with a
// uniform modulus D or one record per marker / per cell, tight element blocks, clear = true, homogeneous mesh.
// (The template also fits Tet10, but there the register-blocked kernel of gl_kernels_bdb.cu measures faster -- 265 vs
// 163 M el/s on 750 000 cells -- so only Hex20 is instantiated.)
//
// Same algebra as gl_kernels_bdb.cu -- geometric tensors P(m,n)[k][l] = sum_p c_p B(m,k) B(n,l) accumulated over the Gauss
// points, D applied once afterwards -- but register-tiled like a symmetric rank-k update, because for 20-node cells the
// one-lane-per-column-node mapping is bound by SHARED-MEMORY bandwidth, not by the FP64 pipe (ncu, Hex20 100^3:
// l1tex data-pipe 79 % busy, 45 % of its wavefronts bank-conflict replays, FP64 pipe 23 %):
//
//   * a lane owns a 2 x 2 NODE tile (R, C), R <= C, of the symmetric node-pair matrix: 36 accumulators; per Gauss point it
//     reads the 6 gradient values of its two row nodes and of its two column nodes as six 16-byte loads and issues 36 DFMA
//     (0.17 loads per DFMA instead of 0.43); 55 tiles cover the 210 distinct pairs of a Hex20 cell;
//   * the gradient record of a Gauss point is node-major and 16-byte aligned, so that a node pair is three LDS.128;
//   * geometry: SD lanes per (cell, Gauss point) as in gl_kernels_bdb.cu (rows of J exchanged by shuffles);
//   * phase D contracts with D' (compile-time sparsity pattern), stages the cell in its final column-major layout and the
//     tile's cells leave with one bulk async copy (L2 evict-first).
//
// FP64 tensor cores (DMMA m8n8k4) were considered for the accumulation (BASELINE.json calls config 3 the "DMMA contraction
// path"): measured on B200 the DMMA peak equals the DFMA peak (37 vs 34 TFLOP/s, same pipe -- gl_bench_fp64_peak), and a
// DMMA tiling cannot skip the transposed half of the node pairs inside 8 x 8 tiles nor the 20 -> 24 node padding, so it
// would issue 1.9x the flops of this symmetric DFMA form for the same pipe; its only advantage, fewer shared-memory
// reads, is obtained here by register tiling instead.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

// strain-displacement column (node, j) in 3D: entries q = 0..2 -> (Mandel row, gradient component)
__host__ __device__ constexpr int col_row3(int j, int q) {
    return j == 0 ? (q == 0 ? 0 : (q == 1 ? 3 : 5)) : (j == 1 ? (q == 0 ? 1 : (q == 1 ? 3 : 4)) : (q == 0 ? 2 : (q == 1 ? 4 : 5)));
}
__host__ __device__ constexpr int col_comp3(int j, int q) {
    return j == 0 ? (q == 0 ? 0 : (q == 1 ? 1 : 2)) : (j == 1 ? (q == 0 ? 1 : (q == 1 ? 0 : 2)) : (q == 0 ? 2 : (q == 1 ? 1 : 0)));
}
// PAT 0 general, PAT 1 major-symmetric, PAT 2 symmetric with uncoupled normal / diagonal shear blocks (exact zeros)
__host__ __device__ constexpr bool d_nz3(int pat, int a, int b) { return pat < 2 ? true : ((a < 3 && b < 3) || a == b); }

struct TileConsts {
    double d[36]; // d'[a][b] = f_a f_b d[a][b] (row-major 6 x 6), f = 1 for normal rows, 1/sqrt(2) for shear rows
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int NN>
struct TileDims {
    static constexpr int SD = 3;
    static constexpr int NP = NN / 2;                // node pairs
    static constexpr int NT = NP * (NP + 1) / 2;     // 2 x 2 node tiles of the upper triangle (incl. the diagonal)
    static constexpr int LPC = (NT + 31) / 32 * 32;  // lanes per cell (whole warps)
    static constexpr int NDOF = NN * SD;
    static constexpr int GS = NN * SD + 2;           // record per (cell, gauss point): B[NN][3] | c | pad   (even: 16-byte aligned)
    static constexpr int XS = (NN * SD) | 1;
    static constexpr int LS = (NN * SD + 6) / 16 * 16 + 9; // = 9 (mod 16): the SD lanes of up to 6 Gauss points of a half-warp read L conflict-free
};

template <int NN, int PAT, bool DREC>
__global__ void __launch_bounds__(256) tile_kernel(const IntegParams p, const TileConsts cst, const int CPC) {
    if (DREC && p.coef_pat_dev != nullptr && *p.coef_pat_dev != PAT) return; // another pattern variant owns this call (see launcher)
    using T = TileDims<NN>;
    constexpr int SD = T::SD, NP = T::NP, NT = T::NT, LPC = T::LPC, NDOF = T::NDOF, GS = T::GS, XS = T::XS, LS = T::LS;
    constexpr bool SYM = PAT >= 1;

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    // The gradient records G are dead once the geometric tensors sit in registers, and the staged element blocks K are dead
    // once the bulk copy has read them: the two share one region (G is the smaller one), which is what lets three CTAs
    // instead of two fit on an SM.  Order per tile: wait for the previous copy's read -> G written (B) -> G read (C) ->
    // barrier -> K written over it (D) -> bulk copy.
    double *s_K = smem;                          // [CPC][NDOF*NDOF]   (16-byte aligned for the bulk copy)
    double *s_G = smem;                          // [CPC][ng][GS]       (aliases s_K; 16-byte aligned for LDS.128)
    double *s_L = s_K + CPC * NDOF * NDOF;       // [ng][LS]
    double *s_X = s_L + ng * LS;                 // [CPC][XS]
    double *s_w = s_X + CPC * XS;                // [ng]
    double *s_D = s_w + ng;                      // [CPC][36]  per-cell d' (DREC only: PER_MARKER / PER_CELL moduli)

    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < ng; i += nt) s_w[i] = p.rule[i];
    for (int i = tid; i < ng * NN * SD; i += nt) s_L[(i / (NN * SD)) * LS + i % (NN * SD)] = p.rule[ng * (1 + NN) + i];

    // phase C role: cell slot cs, tile (R, C) with R <= C.  Tiles are enumerated DIAGONAL BY DIAGONAL (d = C - R = 0, 1, ..; R
    // ascending inside a diagonal): consecutive lanes then differ in R and C together, so the staging stores of phase D --
    // at (6R + ..) + 60 (6C + ..) and mirrored -- and the LDS.128 row / column reads of phase C fall on different banks
    // (with a row-major enumeration the 16 lanes of a half-warp differ only in C, and 360 C = 8 C (mod 16) leaves two banks)
    const int cs = tid / LPC;
    const int tix = tid - cs * LPC;
    const bool has_tile = tix < NT;
    int R = 0, C = 0;
    {
        int t = has_tile ? tix : 0, d = 0;
        while (t >= NP - d) {
            t -= NP - d;
            d++;
        }
        R = t;
        C = t + d;
    }

    // software-pipelined gather (Mesh::set_pad), as in gl_kernels_bdb.cu
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
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // previous bulk copy has left s_K / s_G
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
        if constexpr (DREC) {
            // modulus record of each cell of the tile (its marker's or its own; ns x ns Mandel matrix, column-major) with the
            // 1/sqrt(2) factors of the shear rows folded in, as the launcher does for a uniform modulus
            const double hs = 0.70710678118654752440084436210484903928;
            for (int t = tid; t < ncell_tile * 36; t += nt) {
                const int c = t / 36, idx = t - c * 36;
                const int a = idx / 6, b = idx - a * 6;
                const unsigned long long cell = cell0 + c;
                const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell] : cell - p.coef_begin;
                s_D[t] = p.coef[rec * p.coef_cell_stride + a + b * 6] * (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
            }
        }
        __syncthreads();

        // ---- phase B: geometry, SD lanes per (cell, gauss point) ---------------------------------------------------
        {
            constexpr int GPW = 32 / SD;
            const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
            const int grp = lane / SD, sub = lane - grp * SD;
            const int npair = ncell_tile * ng;
            for (int base = wid * GPW; base < npair; base += nw * GPW) {
                const int pair = base + grp;
                const bool valid = grp < GPW && pair < npair;
                const int pr = valid ? pair : 0;
                const int c = pr / ng, gp = pr - c * ng; // Gauss point fastest: neighbouring groups write records 62 doubles apart
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
                    for (int m = sub; m < NN; m += SD) geom::gradient_row<SD>(L + m * SD, Ji, G + m * SD);
                    if (sub == 0) {
                        if (fabs(det) <= ZERO_DETERMINANT) atomicMin(p.err_cell, cell0 + c);
                        G[NN * SD] = det * s_w[gp] * p.alpha;
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase C: 2 x 2 node tiles of the geometric tensors -----------------------------------------------------
        // acc[a][b] = sum_p c_p v_row[a] v_col[b], a = 3*(row node - 2R) + k, b = 3*(column node - 2C) + l
        double acc[6][6];
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = 0; b < 6; b++) acc[a][b] = 0.0;
        const bool active = has_tile && cs < ncell_tile;
        if (active) {
            const double *Gc = s_G + cs * ng * GS;
#pragma unroll 3
            for (int gp = 0; gp < ng; gp++) {
                const double *G = Gc + gp * GS;
                const double2 *gr = reinterpret_cast<const double2 *>(G + 6 * R);
                const double2 *gc2 = reinterpret_cast<const double2 *>(G + 6 * C);
                const double2 r0 = gr[0], r1 = gr[1], r2 = gr[2];
                const double2 c0 = gc2[0], c1 = gc2[1], c2 = gc2[2];
                const double c = G[NN * SD];
                const double vr[6] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y};
                const double vc[6] = {c0.x * c, c0.y * c, c1.x * c, c1.y * c, c2.x * c, c2.y * c};
#pragma unroll
                for (int b = 0; b < 6; b++)
#pragma unroll
                    for (int a = 0; a < 6; a++) acc[a][b] = fma(vr[a], vc[b], acc[a][b]);
            }
        }

        // ---- phase D: contract with D', stage, ship -------------------------------------------------------------------
        __syncthreads(); // every lane has finished reading G: the region becomes the K staging
        if (active) {
            double *Kc = s_K + cs * NDOF * NDOF;
            const double *Dc = s_D + cs * 36;
#pragma unroll
            for (int j = 0; j < SD; j++)
#pragma unroll
                for (int i = 0; i < SD; i++) {
                    // K(3m+i, 3n+j) = sum_qr d'[row(i,q)][row(j,r)] P(m,n)[comp(i,q)][comp(j,r)]: the (up to) 9 modulus entries of
                    // the component pair (i, j) -- and their transposes for a non-symmetric D -- serve all four node pairs
                    double dr[3][3], dt[3][3];
#pragma unroll
                    for (int q = 0; q < 3; q++)
#pragma unroll
                        for (int r = 0; r < 3; r++) {
                            const int ra = col_row3(i, q), rb = col_row3(j, r);
                            dr[q][r] = dt[q][r] = 0.0;
                            if (!d_nz3(PAT, ra, rb)) continue;
                            dr[q][r] = DREC ? Dc[ra * 6 + rb] : cst.d[ra * 6 + rb];
                            if (!SYM) dt[q][r] = DREC ? Dc[rb * 6 + ra] : cst.d[rb * 6 + ra];
                        }
#pragma unroll
                    for (int rn = 0; rn < 2; rn++)
#pragma unroll
                        for (int cn = 0; cn < 2; cn++) {
                            const int m = 2 * R + rn, n = 2 * C + cn;
                            double v = 0.0, u = 0.0;
                            bool have = false;
#pragma unroll
                            for (int q = 0; q < 3; q++)
#pragma unroll
                                for (int r = 0; r < 3; r++) {
                                    if (!d_nz3(PAT, col_row3(i, q), col_row3(j, r))) continue;
                                    const double pv = acc[3 * rn + col_comp3(i, q)][3 * cn + col_comp3(j, r)];
                                    v = have ? fma(dr[q][r], pv, v) : dr[q][r] * pv;
                                    if (!SYM) u = have ? fma(dt[q][r], pv, u) : dt[q][r] * pv;
                                    have = true;
                                }
                            Kc[(SD * m + i) + NDOF * (SD * n + j)] = v;
                            // off-diagonal tiles also own the transposed pairs: K(n,m) = K(m,n)^T for a major-symmetric D,
                            // its own contraction (with P(n,m) = P(m,n)^T) otherwise
                            if (R != C) Kc[(SD * n + j) + NDOF * (SD * m + i)] = SYM ? v : u;
                        }
                }
        }
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
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NN>
size_t tile_smem(int ng, int cpc, bool drec) {
    using T = TileDims<NN>;
    // G aliases the K staging (ng * GS <= NDOF^2 for every instantiated kind and rule the launcher accepts)
    return sizeof(double) * ((size_t)cpc * T::NDOF * T::NDOF + (size_t)ng * T::LS + (size_t)cpc * T::XS + (size_t)ng + (drec ? (size_t)cpc * 36 : 0));
}

template <int NN>
int launch_tile_kind(const IntegParams &p, const TileConsts &cst, int pat, bool drec, void *stream) {
    using T = TileDims<NN>;
    static int cpc_env = -1;
    if (cpc_env < 0) {
        const char *e = tuning_env("GL_TILE_CPC"); // tuning knob: cells per CTA
        cpc_env = e ? atoi(e) : 0;
    }
    int cpc = cpc_env > 0 ? cpc_env : 256 / T::LPC;
    if (cpc < 1) cpc = 1;
    if (cpc * T::LPC > 256) cpc = 256 / T::LPC;
    if ((size_t)p.ng * T::GS > (size_t)T::NDOF * T::NDOF) return 0; // G must fit inside the staging region it aliases
    while (cpc > 1 && tile_smem<NN>(p.ng, cpc, drec) > 74 * 1024) cpc--;   // three CTAs per SM
    const size_t smem = tile_smem<NN>(p.ng, cpc, drec);
    if (smem > 220 * 1024) return 0;
    const int threads = cpc * T::LPC;
    typedef void (*kern_t)(const IntegParams, const TileConsts, const int);
    auto pick = [&](int pat) -> kern_t {
        return drec ? (pat == 2 ? tile_kernel<NN, 2, true> : (pat == 1 ? tile_kernel<NN, 1, true> : tile_kernel<NN, 0, true>))
                    : (pat == 2 ? tile_kernel<NN, 2, false> : (pat == 1 ? tile_kernel<NN, 1, false> : tile_kernel<NN, 0, false>));
    };
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((226 * 1024) / (smem + 1024));
    if (ctas_per_sm > 2048 / threads) ctas_per_sm = 2048 / threads;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long ntile = (p.hi - p.lo + cpc - 1) / cpc;
    unsigned long long grid = (unsigned long long)nsm * ctas_per_sm * 16; // de-synchronised CTAs, see gl_kernels_hex8.cu
    if (grid > ntile) grid = ntile;
    if (grid == 0) return 0;
    // records whose structure only the device knows: every variant is enqueued, the probe's verdict selects the one that runs
    const bool probe = drec && p.coef_pat_dev != nullptr && tuning_env("GL_BDB_PATTERN") == nullptr;
    IntegParams q = p;
    if (!probe) q.coef_pat_dev = nullptr;
    for (int pt = probe ? 0 : pat; pt <= (probe ? 2 : pat); pt++) {
        kern_t kern = pick(pt);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(q, cst, cpc);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("tile<nn=%d,pat=%d%s>", NN, pat, drec ? ",rec" : "");
    return 1;
}

} // namespace

int launch_bdb_tile(const IntegParams &p, int sd, void *stream) {
    // covered: 3D, even node count, uniform D or one record per marker / cell, tight blocks, overwrite, homogeneous layout,
    // 16-byte aligned output
    if (sd != 3 || p.op != GL_OP_MAT_10_BDB || !p.clear || p.axisym) return 0;
    const bool drec = p.coef != nullptr;
    if (drec && (p.coef_gp_stride != 0 || p.coef_cell_stride != 36)) return 0; // PER_GAUSS moduli: gl_kernels_bdb.cu
    if (p.cell_ids != nullptr || p.out_off != nullptr) return 0;
    const unsigned ndof = (unsigned)p.nnode * 3u;
    if (p.nrow != ndof || p.ncol != ndof || p.ii0 != 0 || p.jj0 != 0) return 0;
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return 0;
    if (tuning_env("GL_DISABLE_TILE")) return 0; // testing knob: fall back to gl_kernels_bdb.cu

    TileConsts cst;
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
    switch (p.nnode) {
    case 20: return launch_tile_kind<20>(p, cst, pat, drec, stream);
    default: return 0;
    }
}

} // namespace gl
