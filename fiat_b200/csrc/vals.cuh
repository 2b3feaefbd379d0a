// Value-table kernel for low-degree Dubiner elements (incl. split cells): see the comment below.
#pragma once
#include "small.cuh"

// ---------------------------------------------------------------------------------------------
// Value-table kernel.  D^alpha of an expansion member of degree k is a combination of the members
// of degree <= k - |alpha| (the fact behind ExpansionSet.get_dmats, FIAT/expansions.py:577-599), so
// the host folds differentiation, chain rule, C0 fix-ups and normalisation into one coefficient
// matrix per (subcell, alpha) (plan.py: derivative_coefficients):
//     out[alpha][row] = sum_{m < C(N-|alpha|+SD, SD)} C_alpha[cell][row][m] * psi_m(x).
// Template parameter J splits the derivative levels: levels <= J come from Leibniz jets carried
// through the recurrence (FIAT/expansions.py:66-137) and share the level-0 matrix (one coefficient
// load feeds C(SD+J, J) FMAs); levels > J use the folded matrices on the member VALUES (one load
// per FMA, but only C(N-k+SD, SD) members).  Each coefficient load costs two shared-memory
// wavefronts whatever the lanes' subcells are, so J balances the shared-memory pipe against the
// FP64 pipe: HCT order 2 needs 744 wavefronts + 500 FP64 instructions per warp with J = 0,
// 456 + 685 with J = 1, 240 + 1180 with full jets (k_small).
// Coefficients sit in shared memory with the subcell index fastest (NCP = ncells padded), so the
// lanes of a warp, whose points fall in different subcells, read distinct banks.  Output rows of
// consecutive (alpha, row) are consecutive table rows: the store pointer advances by the row stride.
// ---------------------------------------------------------------------------------------------
template <int SD, int N, int K, int NCP, int NAJ, bool IDENT, bool ACC>
__device__ __forceinline__ void vals_level(const double* __restrict__& cp, const double (&T)[fb_binom(N + SD, SD)][NAJ],
                                           int nrows, const DevRowMap& M, double* __restrict__ out, double*& o,
                                           long long ostride, long long p, int& a) {
    constexpr int NM = K <= N ? fb_binom(N - K + SD, SD) : 0;
    constexpr int NAK = fb_binom(SD + K - 1, K);
    if (IDENT) {
        const int nq = NAK * nrows;
#pragma unroll 4
        for (int q = 0; q < nq; ++q) {
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < NM; ++m) s = fma(cp[m * NCP], T[m][0], s);
            *o = ACC ? (*o + s) : s;
            o += ostride;
            cp += NM * NCP;
        }
        a += NAK;
    } else {
        for (int j = 0; j < NAK; ++j, ++a) {
            for (int r = 0; r < nrows; ++r) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < NM; ++m) s = fma(cp[m * NCP], T[m][0], s);
                double sgn;
                const size_t orow = fb_map_row(M, r, sgn);
                double* w = out + ((size_t)a * M.total_rows + orow) * ostride + p;
                *w = ACC ? (*w + sgn * s) : sgn * s;
                cp += NM * NCP;
            }
        }
    }
}

// levels 0..J from the jets: one coefficient load per (row, member), NAJ FMAs
template <int SD, int N, int NCP, int NAJ, bool IDENT, bool ACC>
__device__ __forceinline__ void vals_jet_levels(const double* __restrict__ cp, const double (&T)[fb_binom(N + SD, SD)][NAJ],
                                                int nrows, const DevRowMap& M, double* __restrict__ out,
                                                long long ostride, long long p) {
    constexpr int NMEM = fb_binom(N + SD, SD);
    const long long astride = (long long)M.total_rows * ostride;
    double* o = out + p;
#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
        double s[NAJ];
#pragma unroll
        for (int a = 0; a < NAJ; ++a) s[a] = 0.0;
#pragma unroll
        for (int m = 0; m < NMEM; ++m) {
            const double c = cp[m * NCP];
#pragma unroll
            for (int a = 0; a < NAJ; ++a) s[a] = fma(c, T[m][a], s[a]);
        }
        double sgn = 1.0;
        double* w = o;
        if (IDENT) {
            o += ostride;
        } else {
            w = out + fb_map_row(M, r, sgn) * ostride + p;
        }
#pragma unroll
        for (int a = 0; a < NAJ; ++a) {
            double* wa = w + a * astride;
            const double v = IDENT ? s[a] : sgn * s[a];
            *wa = ACC ? (*wa + v) : v;
        }
        cp += NMEM * NCP;
    }
}

template <int SD, int N, int NCP, int J, bool IDENT, bool ACC>
__device__ __forceinline__ void vals_cell(const DevSimplex& P, const SmallTab& st, const double* __restrict__ geom,
                                          const double* __restrict__ cp, double inv_mult, const double (&x)[3],
                                          const DevRowMap& M, double* __restrict__ out, long long ostride, long long p) {
    constexpr int NMEM = fb_binom(N + SD, SD);
    constexpr int NAJ = Jet<SD, J>::NA;
    double xr[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < SD; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int d = 0; d < SD; ++d) acc = fma(x[d], geom[i * SD + d], acc);
        xr[i] = acc + geom[9 + i];
    }
    double fa[3], fb[3];
    recurrence_factors<SD>(xr, fa, fb);
    double T[NMEM][NAJ];
#pragma unroll
    for (int m = 0; m < NMEM; ++m)
#pragma unroll
        for (int a = 0; a < NAJ; ++a) T[m][a] = 0.0;
    T[0][0] = geom[12] * inv_mult;
    int s = 0;
    {
        double dfa[3] = {0, 0, 0}, dfb[3] = {0, 0, 0};
        if (J > 0) {
#pragma unroll
            for (int d = 0; d < SD; ++d) { dfa[d] = geom[14 + d]; dfb[d] = geom[23 + d]; }
        }
        small_chain<SD, N, J, 0>(P, st, s, 0, 0, T, fa[0], fb[0], dfa, dfb);
    }
    if (SD >= 2) {
        double dfa[3] = {0, 0, 0}, dfb[3] = {0, 0, 0};
        if (J > 0) {
#pragma unroll
            for (int d = 0; d < SD; ++d) { dfa[d] = geom[14 + 3 + d]; dfb[d] = geom[23 + 3 + d]; }
        }
#pragma unroll
        for (int pp = 0; pp < N; ++pp) small_chain<SD, N, J, 1>(P, st, s, pp, 0, T, fa[1], fb[1], dfa, dfb);
    }
    if (SD >= 3) {
        double dfa[3] = {0, 0, 0}, dfb[3] = {0, 0, 0};
        if (J > 0) {
#pragma unroll
            for (int d = 0; d < SD; ++d) { dfa[d] = geom[14 + 6 + d]; dfb[d] = geom[23 + 6 + d]; }
        }
#pragma unroll
        for (int qq = 0; qq < N; ++qq)
#pragma unroll
            for (int pp = 0; pp < N; ++pp)
                if (pp + qq < N) small_chain<SD, N, J, 2>(P, st, s, pp, qq, T, fa[2], fb[2], dfa, dfb);
    }
    vals_jet_levels<SD, N, NCP, NAJ, IDENT, ACC>(cp, T, P.nrows, M, out, ostride, p);
    // skip the coefficient blocks of the levels the jets covered
    int skip = NMEM;
#pragma unroll
    for (int k = 1; k <= J; ++k) skip += fb_binom(SD + k - 1, k) * (k <= N ? fb_binom(N - k + SD, SD) : 0);
    cp += (size_t)P.nrows * skip * NCP;
    int a = NAJ;
    double* o = out + p + (size_t)NAJ * P.nrows * ostride;      // identity maps: total_rows == nrows
    if (J < 1 && P.order >= 1) vals_level<SD, N, 1, NCP, NAJ, IDENT, ACC>(cp, T, P.nrows, M, out, o, ostride, p, a);
    if (J < 2 && P.order >= 2) vals_level<SD, N, 2, NCP, NAJ, IDENT, ACC>(cp, T, P.nrows, M, out, o, ostride, p, a);
    if (J < 3 && P.order >= 3) vals_level<SD, N, 3, NCP, NAJ, IDENT, ACC>(cp, T, P.nrows, M, out, o, ostride, p, a);
}

template <int SD, int N, int NCP, int J, bool IDENT>
__global__ void __launch_bounds__(128)
k_vals(const DevSimplex P, const __grid_constant__ SmallTab st, const DevEntity E, const double* __restrict__ pts,
       long long npts, long long ldp, int tiles_per_cta, double* __restrict__ out, long long ostride,
       const __grid_constant__ DevRowMap M) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    double* s_C = smem;                                   // cderiv_len
    double* s_g = s_C + P.cderiv_len;                     // ncells x 32
    for (int i = tid; i < P.cderiv_len; i += blockDim.x) s_C[i] = __ldg(P.cderiv + i);
    for (int i = tid; i < P.ncells * FB_GEOM_DOUBLES; i += blockDim.x) s_g[i] = __ldg(P.geom + i);
    __syncthreads();

    for (int t = 0; t < tiles_per_cta; ++t) {
        const long long p = ((long long)blockIdx.x * tiles_per_cta + t) * blockDim.x + tid;
        if (p >= npts) return;
        double x[3];
        apply_entity<SD>(E, pts + p * ldp, x);
        unsigned mask = NCP == 1 ? 1u : locate_cells<SD>(st.bary, P.ncells, P.unique, x);
        if (mask == 0) {        // in no subcell: zero column, like the reference
            fb_zero_column(M, out, ostride, p, P.na, P.nrows);
            continue;
        }
        const double inv_mult = 1.0 / (double)__popc(mask);
        int cell = NCP == 1 ? 0 : __ffs(mask) - 1;
        mask &= mask - 1;
        vals_cell<SD, N, NCP, J, IDENT, false>(P, st, s_g + cell * FB_GEOM_DOUBLES, s_C + cell, inv_mult, x, M, out, ostride, p);
        while (mask) {      // points on interior facets: add the other subcells' one-sided tables
            cell = __ffs(mask) - 1;
            mask &= mask - 1;
            vals_cell<SD, N, NCP, J, IDENT, true>(P, st, s_g + cell * FB_GEOM_DOUBLES, s_C + cell, inv_mult, x, M, out, ostride, p);
        }
    }
}
