// gl_kernels_small2d.cu -- stiffness kernel for the small 2D kinds (Tri3, Qua4): integ::mat_10_bdb
// (src/integ/mat_10_bdb.rs:121-233, incl. the axisymmetric branch for Tri3) with a uniform modulus D, tight element blocks
// and clear = true  (BASELINE.json config 1: Qua4 plane strain; the Tri3 half of config 5).
//
// These elements are tiny (288 / 512 bytes of output, < 1 kflop), so the kernel is bound by the output stream and by
// per-cell overheads; the CTA-cooperative kernels spend their time in barriers.  Here ONE THREAD OWNS ONE CELL end to end:
//   gather (point-major coordinates: one 16-byte load per node) -> per Gauss point J, inverse, B in the oracle's
//   operation order (gl_geom.cuh) -> geometric tensors P(m,n) = sum_p c_p v(m) (x) v(n) of all node pairs m <= n in
//   registers (gl_kernels_bdb.cu, phase C) -> contraction with D' -> the element block,
// with no barrier at all.  The warp's 32 blocks are transposed through shared memory (stride 33: conflict-free both ways)
// so that global memory sees fully coalesced 16-byte stores of whole lines.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_geom.cuh"
#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int S2_THREADS = 128;

// strain-displacement column (node, j) in 2D: entries q -> (Mandel row, component); component 2 = hoop value N/r
__host__ __device__ constexpr int c2_n(bool axi, int j) { return j == 0 ? (axi ? 3 : 2) : 2; }
__host__ __device__ constexpr int c2_row(int j, int q) { return j == 0 ? (q == 0 ? 0 : (q == 1 ? 3 : 2)) : (q == 0 ? 1 : 3); }
__host__ __device__ constexpr int c2_comp(int j, int q) { return j == 0 ? (q == 0 ? 0 : (q == 1 ? 1 : 2)) : (q == 0 ? 1 : 0); }
__host__ __device__ constexpr bool d2_nz(int pat, int a, int b) { return pat < 2 ? true : ((a < 3 && b < 3) || a == b); }

struct Small2dConsts {
    double d[16]; // d'[a][b] = f_a f_b d[a][b] (row-major 4 x 4), f = (1, 1, 1, 1/sqrt 2)
};

template <int NN, bool AXI, int PAT, bool DREC>
__global__ void __launch_bounds__(S2_THREADS) small2d_kernel(const IntegParams p, const Small2dConsts cst) {
    if (DREC && p.coef_pat_dev != nullptr && *p.coef_pat_dev != PAT) return; // another pattern variant owns this call (see launcher)
    constexpr bool SYM = PAT >= 1;
    constexpr int SD = 2, NDOF = NN * SD, BS = NDOF * NDOF;
    constexpr int VD = SD + (AXI ? 1 : 0);
    constexpr int NPAIR = NN * (NN + 1) / 2;
    constexpr int TS = 33; // transposition stride (doubles)

    extern __shared__ __align__(16) double smem[];
    const int ng = p.ng;
    double *s_w = smem;                  // [ng]
    double *s_N = s_w + ng;              // [ng][NN]
    double *s_L = s_N + ng * NN;         // [ng][NN][2]
    double *s_T = s_L + ng * NN * SD + ((ng * (1 + NN + NN * SD)) & 1); // [warps][BS][TS], 16-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < ng * (1 + NN + NN * SD); i += S2_THREADS) s_w[i] = p.rule[i];
    __syncthreads();
    double *Tw = s_T + (size_t)wid * BS * TS;

    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + S2_THREADS - 1) / S2_THREADS;
    // software-pipelined gather: the COORDINATES of the next batch and the point ids of the batch after it are in flight while this
    // one is integrated, so neither hop of the dependent gather is on the critical path (ncu, round 1: long-scoreboard stalls 3.2 per
    // issued instruction with the coordinates loaded at the top of their own batch)
    // (not for the axisymmetric Tri3 variant: the second set of coordinates costs it a resident CTA -- 168 -> 194 registers, measured
    // 0.60 -> 0.66 ms on the Tri3 half of configs[4]; it keeps the ids one batch ahead only)
    constexpr bool PFX = !AXI;
    unsigned int pid_next[NN];
    double xn[NN][SD];
    auto load_ids = [&](unsigned long long c) {
#pragma unroll
        for (int m = 0; m < NN; m++) pid_next[m] = c < p.hi ? p.conn[(unsigned long long)m * p.ncell_k + c] : 0u;
    };
    auto load_coords = [&](unsigned long long c) {
#pragma unroll
        for (int m = 0; m < NN; m++) {
            xn[m][0] = xn[m][1] = 0.0;
            if (c < p.hi) {
                const double2 a = reinterpret_cast<const double2 *>(p.coords_pm)[pid_next[m]];
                xn[m][0] = a.x;
                xn[m][1] = a.y;
            }
        }
    };
    const unsigned long long cstep = (unsigned long long)gridDim.x * S2_THREADS;
    {
        const unsigned long long c0 = p.lo + (unsigned long long)blockIdx.x * S2_THREADS + tid;
        load_ids(c0);
        if constexpr (PFX) {
            load_coords(c0);
            load_ids(c0 + cstep);
        }
    }
    for (unsigned long long batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const unsigned long long wcell0 = p.lo + batch * S2_THREADS + (unsigned long long)wid * 32; // first cell of this warp
        const unsigned long long cell = wcell0 + lane;
        const bool active = cell < p.hi;

        if constexpr (!PFX) load_coords(cell); // with the ids fetched during the previous batch
        double x[NN][SD];
#pragma unroll
        for (int m = 0; m < NN; m++) {
            x[m][0] = xn[m][0];
            x[m][1] = xn[m][1];
        }
        if constexpr (PFX) {
            load_coords(cell + cstep);  // next batch (its point ids arrived during the previous iteration)
            load_ids(cell + 2 * cstep); // the batch after
        } else {
            load_ids(cell + cstep);
        }
        if (!active) { // a unit cell keeps the arithmetic finite; nothing of it is stored
            x[1 % NN][0] = 1.0;
            x[2 % NN][0] = NN == 3 ? 0.0 : 1.0;
            x[2 % NN][1] = 1.0;
            if (NN == 4) x[3][1] = 1.0;
            if (AXI) {
#pragma unroll
                for (int m = 0; m < NN; m++) x[m][0] += 1.0;
            }
        }

        double acc[NPAIR][VD * VD];
#pragma unroll
        for (int q = 0; q < NPAIR; q++)
#pragma unroll
            for (int k = 0; k < VD * VD; k++) acc[q][k] = 0.0;
        bool bad = false;

#pragma unroll 1
        for (int gp = 0; gp < ng; gp++) {
            const double *L = s_L + gp * NN * SD;
            double J[SD][SD] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int m = 0; m < NN; m++) geom::jacobian_add_node<SD, SD>(J, x[m], L + m * SD);
            double Ji[SD][SD];
            const double det = geom::invert(J, Ji);
            bad = bad || fabs(det) <= ZERO_DETERMINANT;
            double v[NN][VD];
#pragma unroll
            for (int m = 0; m < NN; m++) geom::gradient_row<SD>(L + m * SD, Ji, v[m]);
            double c = det * s_w[gp] * p.alpha;
            if constexpr (AXI) {
                const double *N = s_N + gp * NN;
                double r = 0.0;
#pragma unroll
                for (int m = 0; m < NN; m++) r = geom::add(r, geom::mul(N[m], x[m][0])); // radius_at_ip
                const double rinv = 1.0 / r;
#pragma unroll
                for (int m = 0; m < NN; m++) v[m][SD] = N[m] * rinv;
                c *= r;
            }
            int pair = 0;
#pragma unroll
            for (int n = 0; n < NN; n++) {
                double vn[VD];
#pragma unroll
                for (int l = 0; l < VD; l++) vn[l] = v[n][l] * c;
#pragma unroll
                for (int m = 0; m <= n; m++) {
#pragma unroll
                    for (int l = 0; l < VD; l++)
#pragma unroll
                        for (int k = 0; k < VD; k++) acc[pair][k + VD * l] = fma(v[m][k], vn[l], acc[pair][k + VD * l]);
                    pair++;
                }
            }
        }
        if (active && bad) atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);

        // one modulus per marker / per cell: the lane reads the record of its cell (ns x ns Mandel matrix, column-major) and
        // folds the 1/sqrt(2) factors of the shear row in, as the launcher does for a uniform modulus
        double dl[16] = {};
        if constexpr (DREC) {
            if (active) {
                const unsigned long long rec = p.coef_index ? (unsigned long long)p.coef_index[cell]
                                                            : (p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.coef_begin;
                const double *D = p.coef + rec * p.coef_cell_stride;
                const double hs = 0.70710678118654752440084436210484903928;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++)
                        if (d2_nz(PAT, a, b)) dl[a * 4 + b] = D[a + b * 4] * (a < 3 ? 1.0 : hs) * (b < 3 ? 1.0 : hs);
            }
        }
        auto dval = [&](int a, int b) -> double { return DREC ? dl[a * 4 + b] : cst.d[a * 4 + b]; };

        // ---- contraction with D' into the transposition buffer: entry q of this lane's block at Tw[q*TS + lane] -------------
        __syncwarp(); // the previous batch's copy-out has finished reading Tw
        {
            int pair = 0;
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int m = 0; m <= n; m++) {
#pragma unroll
                    for (int j = 0; j < SD; j++)
#pragma unroll
                        for (int i = 0; i < SD; i++) {
                            // K(2m+i, 2n+j) = sum_qr d'[row(i,q)][row(j,r)] P(m,n)[comp(i,q)][comp(j,r)]
                            double val = 0.0, tr = 0.0;
                            bool have = false;
#pragma unroll
                            for (int q = 0; q < c2_n(AXI, i); q++)
#pragma unroll
                                for (int r = 0; r < c2_n(AXI, j); r++) {
                                    const int ra = c2_row(i, q), rb = c2_row(j, r);
                                    if (!d2_nz(PAT, ra, rb)) continue;
                                    const double pv = acc[pair][c2_comp(i, q) + VD * c2_comp(j, r)];
                                    val = have ? fma(dval(ra, rb), pv, val) : dval(ra, rb) * pv;
                                    if (!SYM) tr = have ? fma(dval(rb, ra), pv, tr) : dval(rb, ra) * pv;
                                    have = true;
                                }
                            Tw[((SD * m + i) + NDOF * (SD * n + j)) * TS + lane] = val;
                            if (m != n) Tw[((SD * n + j) + NDOF * (SD * m + i)) * TS + lane] = SYM ? val : tr;
                        }
                    pair++;
                }
        }
        __syncwarp();

        // ---- coalesced copy-out of the warp's blocks: a flat sweep in 16-byte pieces; every lane knows the destination of its own
        // cell and hands it out by shuffle, so runs of cells that are consecutive in the output (always, on single-kind meshes;
        // row by row on mixed ones) leave as full lines
        {
            unsigned long long mo = 0;
            if (active) mo = p.out_off ? (p.out_off[cell] - p.out_base)
                                       : ((p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell) - p.cell_begin) * BS;
            const int nc = wcell0 < p.hi ? (int)(p.hi - wcell0 < 32 ? p.hi - wcell0 : 32) : 0;
            // one contiguous run: always on single-kind meshes; on mixed ones whenever the warp's cells lie inside one stripe
            const unsigned long long mo0 = __shfl_sync(0xffffffffu, mo, 0);
            const bool one_run = (p.cell_ids == nullptr && p.out_off == nullptr) || __all_sync(0xffffffffu, !active || mo == mo0 + (unsigned long long)lane * BS);
            if (one_run) {
                if (nc > 0) {
                    double2 *dst = reinterpret_cast<double2 *>(p.out + mo0);
                    for (int w = lane; w < nc * (BS / 2); w += 32) {
                        const int c = w / (BS / 2), q = 2 * (w - c * (BS / 2));
                        dst[w] = make_double2(Tw[q * TS + c], Tw[(q + 1) * TS + c]);
                    }
                }
                continue;
            }
            for (int w0 = 0; w0 < nc * (BS / 2); w0 += 32) {
                const int w = w0 + lane, c = w / (BS / 2), q = 2 * (w - c * (BS / 2));
                const unsigned long long base = __shfl_sync(0xffffffffu, mo, c & 31);
                if (w < nc * (BS / 2)) *reinterpret_cast<double2 *>(p.out + base + q) = make_double2(Tw[q * TS + c], Tw[(q + 1) * TS + c]);
            }
        }
    }
}

template <int NN, bool AXI>
int launch_small2d_kind(const IntegParams &p, const Small2dConsts &cst, int pat, bool drec, void *stream) {
    constexpr int BS = NN * 2 * NN * 2;
    const size_t head = (size_t)p.ng * (1 + NN + NN * 2);
    const size_t smem = sizeof(double) * (head + (head & 1) + (size_t)(S2_THREADS / 32) * BS * 33);
    auto pick = [&](int pat) {
        return drec ? (pat == 2 ? small2d_kernel<NN, AXI, 2, true> : (pat == 1 ? small2d_kernel<NN, AXI, 1, true> : small2d_kernel<NN, AXI, 0, true>))
                    : (pat == 2 ? small2d_kernel<NN, AXI, 2, false> : (pat == 1 ? small2d_kernel<NN, AXI, 1, false> : small2d_kernel<NN, AXI, 0, false>));
    };
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int ctas_per_sm = (int)((226 * 1024) / (smem + 1024));
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const unsigned long long nbatch = (p.hi - p.lo + S2_THREADS - 1) / S2_THREADS;
    // CTAs per resident slot: 8 keeps the SMs out of lockstep on large meshes (9 M cells: 11.5 G el/s at 8, 10.1 at 2), but every CTA
    // should still loop over ~8 batches for the gather pipeline to pay (10^6 cells: 10.1 G el/s at 2, 9.3 at 8)
    const unsigned long long slots = (unsigned long long)nsm * ctas_per_sm;
    unsigned long long mult = nbatch / (slots * 8);
    mult = mult < 1 ? 1 : (mult > 8 ? 8 : mult);
    unsigned long long grid = slots * mult;
    if (const char *e = tuning_env("GL_SMALL2D_GRID_MULT")) grid = (unsigned long long)nsm * ctas_per_sm * (unsigned long long)atoi(e); // tuning knob
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    // records whose structure only the device knows: every variant is enqueued, the probe's verdict selects the one that runs
    const bool probe = drec && p.coef_pat_dev != nullptr && tuning_env("GL_BDB_PATTERN") == nullptr;
    IntegParams q = p;
    if (!probe) q.coef_pat_dev = nullptr;
    for (int pt = probe ? 0 : pat; pt <= (probe ? 2 : pat); pt++) {
        auto kern = pick(pt);
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kern<<<(unsigned)grid, S2_THREADS, smem, (cudaStream_t)stream>>>(q, cst);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    note_kernel("small2d<nn=%d,axi=%d,pat=%d%s>", NN, p.axisym ? 1 : 0, pat, drec ? ",rec" : "");
    return 1;
}

} // namespace

int launch_bdb_small2d(const IntegParams &p, int sd, void *stream) {
    // covered: Tri3 (plane and axisymmetric) and Qua4 (plane), uniform D or one record per marker / cell, tight blocks,
    // overwrite, point-major coordinates
    if (sd != 2 || p.op != GL_OP_MAT_10_BDB || !p.clear || p.coords_pm == nullptr) return 0;
    const bool drec = p.coef != nullptr;
    if (drec && (p.coef_gp_stride != 0 || p.coef_cell_stride != 16)) return 0; // PER_GAUSS moduli: gl_kernels_bdb.cu
    if (p.nnode != 3 && !(p.nnode == 4 && !p.axisym)) return 0;
    const unsigned ndof = (unsigned)p.nnode * 2u;
    if (p.nrow != ndof || p.ncol != ndof || p.ii0 != 0 || p.jj0 != 0) return 0;
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return 0;
    if (tuning_env("GL_DISABLE_SMALL2D")) return 0; // testing knob: fall back to gl_kernels_bdb.cu

    Small2dConsts cst;
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
                if (!d2_nz(2, a, b) && cst.d[a * 4 + b] != 0.0) pat = 1;
    }
    if (drec) pat = p.coef_pat; // structure common to all records (0 when the records live on the device only)
    if (const char *e = tuning_env("GL_BDB_PATTERN")) {
        const int cap = atoi(e);
        if (cap < pat) pat = cap < 0 ? 0 : cap;
    }
    if (p.nnode == 3)
        return p.axisym ? launch_small2d_kind<3, true>(p, cst, pat, drec, stream) : launch_small2d_kind<3, false>(p, cst, pat, drec, stream);
    return launch_small2d_kind<4, false>(p, cst, pat, drec, stream);
}

} // namespace gl
