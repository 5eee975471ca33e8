// gl_kernels_hex8.cu -- register-blocked Hex8 stiffness kernel: integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-208) for
// GeoKind::Hex8 with the 8-point rule, a uniform modulus D, tight element blocks and clear = true  (BASELINE.json
// configs[1], the headline workload: 4,608 B of output and ~37 kflop per cell -- the ridge of the B200 roofline).
//
// Mapping: EIGHT LANES PER CELL, four cells per warp, no block-level barrier anywhere.
//   phase 0  lane g gathers node g of its cell (Mesh::set_pad, src/mesh/mesh.rs:448-456) -- prefetched one batch ahead;
//   phase 1  lane g owns GAUSS POINT g: J = Xt.L, closed-form inverse/det, B = L.inv(J)
//            (scratchpad_calc_jacobian.rs:101-131, scratchpad_calc_gradient.rs:78-91); B and c = det*w*alpha go to smem;
//   phase 2  lane n owns COLUMN NODE n: per Gauss point T = c.D'.M(n) (18 values, D' from the constant bank), then the
//            3x3 node-pair blocks K(m,n) += M(m)^T.T with 27 DFMA each, accumulated in registers over the 8 points.
//            For a major-symmetric D only the "circulant half" m = n..n+4 (mod 8) is computed and mirrored;
//   phase 3  blocks are staged in smem in the final column-major layout and the 4 cells of a warp (18,432 contiguous
//            bytes) leave the SM as ONE bulk async copy (cp.async.bulk.global.shared::cta -> UBLKCP), so HBM sees only
//            full lines and no thread waits on a store.
// Shared-memory strides (25 / 200 doubles) are chosen so that every 64-bit access of a half-warp (2 cells x 8 lanes)
// hits 16 distinct bank pairs in phases 1 and 2.
#include <cuda_runtime.h>

#include <cstdlib>

#include "gl_internal.hpp"

namespace gl {

namespace {

constexpr int H8_THREADS = 128;
constexpr int H8_CELLS = H8_THREADS / 8; // cells per CTA batch
constexpr int GP_STRIDE = 25;            // doubles per (cell, gauss point): 24 gradient values + coefficient
constexpr int CELL_STRIDE = 8 * GP_STRIDE; // 200 = 8 (mod 16): second cell of a half-warp lands 16 banks away
constexpr int X_STRIDE = 25;
constexpr int K_STRIDE = 576;            // dense: the 4 cells of a warp are one contiguous 18,432-byte run (one bulk copy)
constexpr int L_STRIDE = 25;

struct Hex8Consts {
    double d[36]; // d'[a][b] = f_a f_b d[a][b], f = (1,1,1,1/sqrt2,1/sqrt2,1/sqrt2): the Mandel factors folded into D
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

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
__host__ __device__ constexpr bool t_nz(int pat, int a, int j) { return d_nz(pat, a, col_row(j, 0)) || d_nz(pat, a, col_row(j, 1)) || d_nz(pat, a, col_row(j, 2)); }

template <int PAT, int GP_UNROLL_N>
__global__ void __launch_bounds__(H8_THREADS, 2) hex8_bdb_kernel(const IntegParams p, const Hex8Consts cst) {
    constexpr bool SYM = PAT >= 1;
    constexpr int ND = SYM ? 5 : 8; // node-pair blocks per lane
    extern __shared__ __align__(16) double smem[];
    double *s_L = smem;                         // [8][L_STRIDE]  L[gp][m*3+j]
    double *s_w = s_L + 8 * L_STRIDE;           // [8]
    double *s_X = s_w + 8;                      // [H8_CELLS][X_STRIDE]
    double *s_B = s_X + H8_CELLS * X_STRIDE;    // [H8_CELLS][CELL_STRIDE]
    double *s_K = s_B + H8_CELLS * CELL_STRIDE; // [H8_CELLS][K_STRIDE]  (16-byte aligned: offset is a multiple of 2 doubles)

    const int tid = threadIdx.x;
    const int g = tid & 7;   // lane within the cell group
    const int cs = tid >> 3; // cell slot in the batch

    // reference-element table: w[8] | N[8][8] | L[8][8][3]
    for (int i = tid; i < 8 * 24; i += H8_THREADS) s_L[(i / 24) * L_STRIDE + (i % 24)] = p.rule[8 + 64 + i];
    if (tid < 8) s_w[tid] = p.rule[tid];
    __syncthreads();

    const unsigned long long ncell = p.hi - p.lo;
    const unsigned long long nbatch = (ncell + H8_CELLS - 1) / H8_CELLS;
    double *Xc = s_X + cs * X_STRIDE;
    double *Bc = s_B + cs * CELL_STRIDE;
    double *Kc = s_K + cs * K_STRIDE;
    const double *Lg = s_L + g * L_STRIDE;

    // software-pipelined gather (Mesh::set_pad): coordinates are loaded one batch ahead and the point id they depend on
    // two batches ahead, so neither the index load nor the dependent coordinate loads ever stall a warp
    double nx = 0.0, ny = 0.0, nz = 0.0;
    unsigned int pid_next = 0;
    {
        const unsigned long long cell = p.lo + (unsigned long long)blockIdx.x * H8_CELLS + cs;
        if (cell < p.hi) {
            const unsigned int pid = p.conn[(unsigned long long)g * p.ncell_k + cell];
            nx = p.coords[pid];
            ny = p.coords[p.npoint + pid];
            nz = p.coords[2 * p.npoint + pid];
        }
        const unsigned long long cell1 = cell + (unsigned long long)gridDim.x * H8_CELLS;
        if (cell1 < p.hi) pid_next = p.conn[(unsigned long long)g * p.ncell_k + cell1];
    }

    for (unsigned long long batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const unsigned long long cell = p.lo + batch * H8_CELLS + cs;
        const bool active = cell < p.hi;

        // ---- phase 0: publish this batch's coordinates, prefetch the next batch's ---------------------------------
        Xc[g * 3 + 0] = nx;
        Xc[g * 3 + 1] = ny;
        Xc[g * 3 + 2] = nz;
        {
            const unsigned long long step = (unsigned long long)gridDim.x * H8_CELLS;
            if (cell + step < p.hi) { // next batch: its point id arrived during the previous iteration
                nx = p.coords[pid_next];
                ny = p.coords[p.npoint + pid_next];
                nz = p.coords[2 * p.npoint + pid_next];
            }
            if (cell + 2 * step < p.hi) pid_next = p.conn[(unsigned long long)g * p.ncell_k + cell + 2 * step];
        }
        __syncwarp();

        // ---- phase 1: geometry at Gauss point g ------------------------------------------------------------------
        {
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const double l0 = Lg[m * 3 + 0], l1 = Lg[m * 3 + 1], l2 = Lg[m * 3 + 2];
                const double x0 = Xc[m * 3 + 0], x1 = Xc[m * 3 + 1], x2 = Xc[m * 3 + 2];
                J[0][0] = fma(x0, l0, J[0][0]); J[0][1] = fma(x0, l1, J[0][1]); J[0][2] = fma(x0, l2, J[0][2]);
                J[1][0] = fma(x1, l0, J[1][0]); J[1][1] = fma(x1, l1, J[1][1]); J[1][2] = fma(x1, l2, J[1][2]);
                J[2][0] = fma(x2, l0, J[2][0]); J[2][1] = fma(x2, l1, J[2][1]); J[2][2] = fma(x2, l2, J[2][2]);
            }
            const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
            const double c01 = J[1][0] * J[2][2] - J[1][2] * J[2][0];
            const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            const double det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02;
            if (active && fabs(det) <= ZERO_DETERMINANT) atomicMin(p.err_cell, p.cell_ids ? (unsigned long long)p.cell_ids[cell] : cell);
            const double rdet = 1.0 / det;
            double Ji[3][3];
            Ji[0][0] = c00 * rdet;
            Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * rdet;
            Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * rdet;
            Ji[1][0] = -c01 * rdet;
            Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * rdet;
            Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * rdet;
            Ji[2][0] = c02 * rdet;
            Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * rdet;
            Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * rdet;
            double *Bg = Bc + g * GP_STRIDE;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const double l0 = Lg[m * 3 + 0], l1 = Lg[m * 3 + 1], l2 = Lg[m * 3 + 2];
                Bg[m * 3 + 0] = fma(l2, Ji[2][0], fma(l1, Ji[1][0], l0 * Ji[0][0]));
                Bg[m * 3 + 1] = fma(l2, Ji[2][1], fma(l1, Ji[1][1], l0 * Ji[0][1]));
                Bg[m * 3 + 2] = fma(l2, Ji[2][2], fma(l1, Ji[1][2], l0 * Ji[0][2]));
            }
            Bg[24] = det * s_w[g] * p.alpha;
        }
        __syncwarp();

        // ---- phase 2: lane g = column node n; accumulate its node-pair blocks over the 8 Gauss points -------------
        double acc[ND][9];
#pragma unroll
        for (int b = 0; b < ND; b++)
#pragma unroll
            for (int q = 0; q < 9; q++) acc[b][q] = 0.0;

#pragma unroll GP_UNROLL_N
        for (int gp = 0; gp < 8; gp++) {
            const double *Bg = Bc + gp * GP_STRIDE;
            const double c = Bg[24];
            const double bn[3] = {Bg[g * 3 + 0] * c, Bg[g * 3 + 1] * c, Bg[g * 3 + 2] * c};
            // T[a][j] = sum_t d'[a][row(j,t)] * bn[comp(j,t)]: column (n,j) of M' has the gradient components
            // j=0 -> rows (0,3,5) = (b0,b1,b2); j=1 -> rows (1,3,4) = (b1,b0,b2); j=2 -> rows (2,4,5) = (b2,b1,b0)
            double T[6][3];
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    double t = 0.0;
                    bool have = false;
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        if (d_nz(PAT, a, col_row(j, q))) {
                            const double dv = cst.d[a * 6 + col_row(j, q)], bv = bn[col_comp(j, q)];
                            t = have ? fma(dv, bv, t) : dv * bv;
                            have = true;
                        }
                    }
                    T[a][j] = t;
                }
#pragma unroll
            for (int b = 0; b < ND; b++) {
                const int m = (g + b) & 7;
                const double bm[3] = {Bg[m * 3 + 0], Bg[m * 3 + 1], Bg[m * 3 + 2]};
                // In the symmetric variants the diagonal block (b = 0) is itself symmetric, and the block at distance 4 is
                // held as a transposed duplicate by lane g+4: both need only their upper triangles (i <= j).
                const bool tri = SYM && (b == 0 || b == 4);
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        if (tri && i > j) continue;
                        // K(3m+i, 3n+j) += sum_t M'(m)[row(i,t)][i] * T[row(i,t)][j]
                        double v = acc[b][i + 3 * j];
#pragma unroll
                        for (int q = 2; q >= 0; q--)
                            if (t_nz(PAT, col_row(i, q), j)) v = fma(bm[col_comp(i, q)], T[col_row(i, q)][j], v);
                        acc[b][i + 3 * j] = v;
                    }
            }
        }

        // ---- phase 3: stage the blocks column-major and ship the warp's cells with one bulk copy ------------------
        if ((tid & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // previous copy has left s_K
        __syncwarp();
        // (the two cells of a half-warp are 1152 words apart: a 2-way bank conflict on these 72 stores, ~1% of the iteration)
#pragma unroll
        for (int b = 0; b < ND; b++) {
            const int m = (g + b) & 7;
            const bool tri = SYM && (b == 0 || b == 4);
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    if (tri && i > j) continue; // lower triangle comes from the mirror (b = 0) or from lane g+4 (b = 4)
                    const double v = acc[b][i + 3 * j];
                    Kc[(3 * m + i) + 24 * (3 * g + j)] = v;
                    // mirrored entry K(n,m)^T: whole blocks at distance 1..3, strict upper triangles of the b = 0 / b = 4 blocks
                    if (SYM && ((b >= 1 && b <= 3) || (tri && i < j))) Kc[(3 * g + j) + 24 * (3 * m + i)] = v;
                }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (p.cell_ids == nullptr) {
            // homogeneous mesh: the warp's 4 cells are consecutive in the output
            if ((tid & 31) == 0 && active) {
                const unsigned long long remaining = p.hi - cell;
                const unsigned int nb = (unsigned int)(remaining < 4 ? remaining : 4) * 4608u;
                double *dst = p.out + (cell - p.cell_begin) * 576ull;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(Kc)), "r"(nb)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (g == 0 && active) {
            double *dst = p.out + ((unsigned long long)p.cell_ids[cell] - p.cell_begin) * 576ull;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(Kc)), "r"(4608)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

constexpr size_t H8_SMEM = sizeof(double) * (8 * L_STRIDE + 8 + H8_CELLS * X_STRIDE + H8_CELLS * CELL_STRIDE + H8_CELLS * K_STRIDE);

} // namespace

int launch_hex8_bdb(const IntegParams &p, void *stream) {
    // covered: Hex8, 8-point rule, uniform D, tight 24x24 blocks, overwrite, homogeneous layout, 16-byte aligned output
    if (p.nnode != 8 || p.ng != 8 || p.op != GL_OP_MAT_10_BDB) return 0;
    if (p.coef != nullptr || p.axisym || !p.clear || p.out_off != nullptr) return 0;
    if (p.nrow != 24 || p.ncol != 24 || p.ii0 != 0 || p.jj0 != 0 || p.size_r != 24) return 0;
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return 0;

    Hex8Consts cst;
    const double h = 0.70710678118654752440084436210484903928;
    bool sym = true;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            const double fa = a < 3 ? 1.0 : h, fb = b < 3 ? 1.0 : h;
            cst.d[a * 6 + b] = p.coef_uniform[a + b * 6] * fa * fb; // ABI record is column-major
            if (p.coef_uniform[a + b * 6] != p.coef_uniform[b + a * 6]) sym = false;
        }
    // unroll factor of the Gauss-point loop (tuning knob; GL_HEX8_UNROLL overrides the default)
    static int unroll = -1;
    if (unroll < 0) {
        const char *e = getenv("GL_HEX8_UNROLL");
        unroll = e ? atoi(e) : 1;
        if (unroll != 1 && unroll != 2) unroll = 1;
    }
    int pat = 0;
    if (sym) {
        pat = 2;
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++)
                if (!d_nz(2, a, b) && cst.d[a * 6 + b] != 0.0) pat = 1;
    }
    {
        const char *e = getenv("GL_HEX8_PATTERN"); // tuning/testing knob: cap the specialisation level
        if (e && atoi(e) < pat) pat = atoi(e) < 0 ? 0 : atoi(e);
    }
    typedef void (*kern_t)(const IntegParams, const Hex8Consts);
    kern_t kern;
    if (unroll == 2) kern = pat == 2 ? hex8_bdb_kernel<2, 2> : (pat == 1 ? hex8_bdb_kernel<1, 2> : hex8_bdb_kernel<0, 2>);
    else kern = pat == 2 ? hex8_bdb_kernel<2, 1> : (pat == 1 ? hex8_bdb_kernel<1, 1> : hex8_bdb_kernel<0, 1>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H8_SMEM) != cudaSuccess) return -1;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long nbatch = (p.hi - p.lo + H8_CELLS - 1) / H8_CELLS;
    unsigned long long grid = (unsigned long long)nsm * 2; // persistent: two resident CTAs per SM
    if (grid > nbatch) grid = nbatch;
    if (grid == 0) return 0;
    kern<<<(unsigned)grid, H8_THREADS, H8_SMEM, (cudaStream_t)stream>>>(p, cst);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

} // namespace gl
