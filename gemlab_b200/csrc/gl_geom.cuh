// gl_geom.cuh -- per-Gauss-point geometry shared by the integration kernels, in the ORACLE'S OPERATION ORDER.
//
// J = Xt.L mixes coordinates of magnitude |x| to produce entries of magnitude h (the cell size), so its rounding error is
// amplified by |x|/h: on a 10^7-cell mesh two correct evaluations that merely sum in a different order (or fuse the
// multiply-adds) differ by ~1e-12 in the element matrices.  To keep the parity with the reference restatement independent
// of the mesh, the kernels evaluate J, det(J), inv(J), B = L.inv(J) and r = sum N.x exactly as the oracle does
// (oracle/gemlab_oracle.c: go_pad_calc_jacobian, mat_inverse_small, go_pad_calc_gradient, radius_at_ip, which follow
// src/shapes/scratchpad_calc_jacobian.rs:101-131, russell_lab::mat_inverse and scratchpad_calc_gradient.rs:78-91):
// node sums run m = 0 .. nnode-1, every product and sum is rounded separately (no FMA contraction), cofactors are
// DIVIDED by the determinant.  The well-conditioned contractions downstream use FMAs freely.
#pragma once

namespace gl {
namespace geom {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// J[i][j] += x[i] * l[j] for one node (call for m = 0, 1, ..)
template <int SD, int GD>
__device__ __forceinline__ void jacobian_add_node(double (&J)[SD][GD], const double *x, const double *l) {
#pragma unroll
    for (int i = 0; i < SD; i++)
#pragma unroll
        for (int j = 0; j < GD; j++) J[i][j] = add(J[i][j], mul(x[i], l[j]));
}

// a / b, correctly rounded, given rb = RN(1/b): q0 = RN(a.rb) is within one ulp of a/b, the FMA residual r = a - b.q0 is
// exact, and RN(q0 + r.rb) is then the correctly rounded quotient (Markstein 1990).  Three instructions instead of the ~25
// of the full division sequence; identical results for every finite, normal operand pair -- which is all a
// non-degenerate Jacobian can produce (|det| > 1e-15 is checked by the callers).
__device__ __forceinline__ double div_by(double a, double b, double rb) {
    const double q0 = mul(a, rb);
    const double r = fma(-q0, b, a);
    return fma(r, rb, q0);
}

// russell_lab::mat_inverse, closed form for 2x2 / 3x3; returns the determinant (the caller checks |det| <= 1e-15)
__device__ __forceinline__ double invert(const double (&J)[2][2], double (&Ji)[2][2]) {
    const double det = sub(mul(J[0][0], J[1][1]), mul(J[0][1], J[1][0]));
    const double rd = 1.0 / det;
    Ji[0][0] = div_by(J[1][1], det, rd);
    Ji[0][1] = div_by(-J[0][1], det, rd);
    Ji[1][0] = div_by(-J[1][0], det, rd);
    Ji[1][1] = div_by(J[0][0], det, rd);
    return det;
}

__device__ __forceinline__ double invert(const double (&J)[3][3], double (&Ji)[3][3]) {
    const double c00 = sub(mul(J[1][1], J[2][2]), mul(J[1][2], J[2][1]));
    const double c01 = sub(mul(J[1][0], J[2][2]), mul(J[1][2], J[2][0]));
    const double c02 = sub(mul(J[1][0], J[2][1]), mul(J[1][1], J[2][0]));
    const double det = add(sub(mul(J[0][0], c00), mul(J[0][1], c01)), mul(J[0][2], c02));
    const double rd = 1.0 / det;
    Ji[0][0] = div_by(c00, det, rd);
    Ji[0][1] = div_by(sub(mul(J[0][2], J[2][1]), mul(J[0][1], J[2][2])), det, rd);
    Ji[0][2] = div_by(sub(mul(J[0][1], J[1][2]), mul(J[0][2], J[1][1])), det, rd);
    Ji[1][0] = div_by(-c01, det, rd);
    Ji[1][1] = div_by(sub(mul(J[0][0], J[2][2]), mul(J[0][2], J[2][0])), det, rd);
    Ji[1][2] = div_by(sub(mul(J[0][2], J[1][0]), mul(J[0][0], J[1][2])), det, rd);
    Ji[2][0] = div_by(c02, det, rd);
    Ji[2][1] = div_by(sub(mul(J[0][1], J[2][0]), mul(J[0][0], J[2][1])), det, rd);
    Ji[2][2] = div_by(sub(mul(J[0][0], J[1][1]), mul(J[0][1], J[1][0])), det, rd);
    return det;
}

// one row of B = L.inv(J): b[j] = sum_k l[k] * Ji[k][j]
template <int SD>
__device__ __forceinline__ void gradient_row(const double *l, const double (&Ji)[SD][SD], double *b) {
#pragma unroll
    for (int j = 0; j < SD; j++) {
        double sum = mul(l[0], Ji[0][j]);
#pragma unroll
        for (int k = 1; k < SD; k++) sum = add(sum, mul(l[k], Ji[k][j]));
        b[j] = sum;
    }
}

} // namespace geom
} // namespace gl
