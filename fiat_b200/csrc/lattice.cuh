// Product-form tabulation of the nodal (Lagrange) basis of the principal lattice on the UFC simplex.
//
// For dofs that are point evaluations at the lattice points alpha/n the nodal basis that
// CiarletElement.__init__ (FIAT/finite_element.py:132-165) obtains by inverting the Vandermonde
// matrix is, in exact arithmetic,
//     phi_alpha(x) = prod_{i=0..sd} l_{alpha_i}(lambda_i(x)),   l_k(t) = prod_{j<k} (n t - j) / (j + 1),
// with lambda the barycentric coordinates.  Evaluating this product and its derivatives needs
// ~30 flops per basis function instead of the ~400 of the expansion + coefficient contraction
// (FIAT/expansions.py:140-267, FIAT/polynomial_set.py:71), which turns the FP64-bound P8
// tetrahedron case into an output-write-bound one.  The host only selects this kernel after
// checking, on the device, that it reproduces the general kernel on random points
// (fiat_b200/api.py), and the parity tests compare it with the reference like any other path.
//
// Thread per point.  The 1-D factors l_k(lambda_i) and their first/second derivatives live in a
// private shared-memory column; the loop nest a0, a1, a2 (a3 = n - a0 - a1 - a2) forms every
// basis function from the pair products P = (lambda_0, lambda_1) and Q = (lambda_2, lambda_3).
// On the UFC simplex lambda_0 = 1 - sum x, lambda_i = x_{i-1}, so d/dx_j = d_j - d_0 in
// barycentric derivatives; those differences are taken on the pair products.
#pragma once
#include "expansion.cuh"


template <int SD, int ORDER>
__global__ void __launch_bounds__(128)
k_lattice(const DevLattice L, const DevEntity E, const double* __restrict__ pts, long long npts, long long ldp,
          double* __restrict__ out, long long ostride, const __grid_constant__ DevRowMap M) {
    constexpr int NR = ORDER + 1;
    extern __shared__ double smem[];
    const int BP = blockDim.x;
    const long long p = (long long)blockIdx.x * BP + threadIdx.x;
    if (p >= npts) return;
    const int n = L.degree, n1 = n + 1;
    double* U = smem + threadIdx.x;            // U[((i * n1 + k) * NR + r) * BP]
    double x[3];
    apply_entity<SD>(E, pts + p * ldp, x);
    double lam[SD + 1];
    lam[0] = 1.0;
#pragma unroll
    for (int i = 0; i < SD; ++i) {
        lam[i + 1] = x[i];
        lam[0] -= x[i];
    }
    // 1-D factors and their derivatives with respect to the own barycentric coordinate
    const double dn = (double)n;
#pragma unroll
    for (int i = 0; i <= SD; ++i) {
        double v = 1.0, d1 = 0.0, d2 = 0.0;
        double* Ui = U + (size_t)i * n1 * NR * BP;
        Ui[0] = 1.0;
        if (ORDER >= 1) Ui[BP] = 0.0;
        if (ORDER >= 2) Ui[2 * BP] = 0.0;
        const double nt = dn * lam[i];
        for (int k = 0; k < n; ++k) {
            const double rk = __ldg(L.recip + k);
            const double f = (nt - (double)k) * rk, df = dn * rk;
            if (ORDER >= 2) d2 = fma(d2, f, 2.0 * d1 * df);
            if (ORDER >= 1) d1 = fma(d1, f, v * df);
            v *= f;
            double* Uk = Ui + (size_t)(k + 1) * NR * BP;
            Uk[0] = v;
            if (ORDER >= 1) Uk[BP] = d1;
            if (ORDER >= 2) Uk[2 * BP] = d2;
        }
    }
    const double* U0 = U;
    const double* U1 = U + (size_t)1 * n1 * NR * BP;
    const double* U2 = U + (size_t)2 * n1 * NR * BP;
    const double* U3 = U + (size_t)(SD == 3 ? 3 : 2) * n1 * NR * BP;
    const size_t nd = (size_t)M.total_rows;
    const double sg = M.sign[0];
    double* o = out + p;
    int idx = 0;
    for (int a0 = 0; a0 <= n; ++a0) {
        const double* u0 = U0 + (size_t)a0 * NR * BP;
        const double u0v = u0[0], u0d = ORDER >= 1 ? u0[BP] : 0.0, u0dd = ORDER >= 2 ? u0[2 * BP] : 0.0;
        for (int a1 = 0; a1 <= n - a0; ++a1) {
            const double* u1 = U1 + (size_t)a1 * NR * BP;
            const double u1v = u1[0], u1d = ORDER >= 1 ? u1[BP] : 0.0, u1dd = ORDER >= 2 ? u1[2 * BP] : 0.0;
            // pair (lambda_0, lambda_1): value, d0, d1 - d0, d00, d01 - d00, d11 - 2 d01 + d00
            const double P = u0v * u1v;
            const double Pa = u0d * u1v;
            const double Pd1 = fma(u0v, u1d, -Pa);
            const double Paa = u0dd * u1v;
            const double Pab = u0d * u1d;
            const double Pd2 = Pab - Paa;
            const double Pxx = fma(u0v, u1dd, fma(-2.0, Pab, Paa));
            const int m = n - a0 - a1;
            for (int a2 = (SD == 3 ? 0 : m); a2 <= m; ++a2) {
                const double* u2 = U2 + (size_t)a2 * NR * BP;
                const double u2v = u2[0], u2d = ORDER >= 1 ? u2[BP] : 0.0, u2dd = ORDER >= 2 ? u2[2 * BP] : 0.0;
                double u3v = 1.0, u3d = 0.0, u3dd = 0.0;
                if (SD == 3) {
                    const double* u3 = U3 + (size_t)(m - a2) * NR * BP;
                    u3v = u3[0];
                    if (ORDER >= 1) u3d = u3[BP];
                    if (ORDER >= 2) u3dd = u3[2 * BP];
                }
                const size_t row = (size_t)(M.dof_base + __ldg(L.rowmap + idx)) * M.nc_out + M.comp_out[0];
                ++idx;
                double* orow = o + row * ostride;
                const double Q = sg * (u2v * u3v);
                orow[0] = P * Q;
                if (ORDER >= 1) {
                    const double Qc = sg * (u2d * u3v);     // d2
                    const double Qd = sg * (u2v * u3d);     // d3 (3-D only)
                    const double PaQ = Pa * Q;
                    orow[(1 * nd) * ostride] = Pd1 * Q;                                   // d/dx
                    orow[(2 * nd) * ostride] = fma(P, Qc, -PaQ);                          // d/dy
                    if (SD == 3) orow[(3 * nd) * ostride] = fma(P, Qd, -PaQ);             // d/dz
                    if (ORDER >= 2) {
                        const double Qcc = sg * (u2dd * u3v), Qdd = sg * (u2v * u3dd), Qcd = sg * (u2d * u3d);
                        const double PaaQ = Paa * Q;
                        const double Pd2Q = Pd2 * Q;
                        if (SD == 2) {
                            orow[(3 * nd) * ostride] = Pxx * Q;                                       // xx
                            orow[(4 * nd) * ostride] = fma(Pd1, Qc, -Pd2Q);                           // xy
                            orow[(5 * nd) * ostride] = fma(P, Qcc, fma(-2.0 * Pa, Qc, PaaQ));         // yy
                        } else {
                            orow[(4 * nd) * ostride] = Pxx * Q;                                       // xx
                            orow[(5 * nd) * ostride] = fma(Pd1, Qc, -Pd2Q);                           // xy
                            orow[(6 * nd) * ostride] = fma(Pd1, Qd, -Pd2Q);                           // xz
                            orow[(7 * nd) * ostride] = fma(P, Qcc, fma(-2.0 * Pa, Qc, PaaQ));         // yy
                            orow[(8 * nd) * ostride] = fma(P, Qcd, fma(-Pa, Qc + Qd, PaaQ));          // yz
                            orow[(9 * nd) * ostride] = fma(P, Qdd, fma(-2.0 * Pa, Qd, PaaQ));         // zz
                        }
                    }
                }
            }
        }
    }
}
