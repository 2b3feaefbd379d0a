// Register-resident kernel for low-degree Dubiner elements (incl. split cells: HCT, Powell-Sabin,
// Alfeld/iso Lagrange): thread per point, the whole expansion table T[member][alpha] lives in
// registers because degree N, dimension SD and derivative order are template parameters and every
// loop of the recurrence (FIAT/expansions.py:202-249) is unrolled at compile time.  Members keep their
// Morton numbers (FIAT/expansions.py:16-21); the C0 fix-ups and entity reordering (:270-322) and the
// normalisation (:251-266) are folded into the per-subcell coefficient matrices on the host
// (plan.py: ccell_morton).  Per-subcell tables sit in shared memory because each thread picks its own
// subcell.
#pragma once
#include "expansion.cuh"


__host__ __device__ constexpr int fb_morton2(int p, int q) { return (p + q) * (p + q + 1) / 2 + q; }
__host__ __device__ constexpr int fb_morton3(int p, int q, int r) {
    return (p + q + r) * (p + q + r + 1) * (p + q + r + 2) / 6 + (q + r) * (q + r + 1) / 2 + r;
}
template <int SD>
__host__ __device__ constexpr int fb_member(int p, int q, int r) {
    return SD == 1 ? p : (SD == 2 ? fb_morton2(p, q) : fb_morton3(p, q, r));
}

// one chain: members (sub, 0..len) along direction `codim`; sub-index entries beyond codim are 0
template <int SD, int N, int ORDER, int CODIM>
__device__ __forceinline__ void small_chain(const DevSimplex& P, const SmallTab& st, int& s, int p, int q,
                                            double (&T)[fb_binom(N + SD, SD)][Jet<SD, ORDER>::NA],
                                            double fav, double fbv, const double (&dfa)[3], const double (&dfb)[3]) {
    typedef Jet<SD, ORDER> J;
    const int ssum = (CODIM >= 1 ? p : 0) + (CODIM >= 2 ? q : 0);
    const double fcv = fbv * fbv;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i < N - ssum) {
            const int ic = CODIM == 0 ? fb_member<SD>(i, 0, 0) : (CODIM == 1 ? fb_member<SD>(p, i, 0) : fb_member<SD>(p, q, i));
            const int in = CODIM == 0 ? fb_member<SD>(i + 1, 0, 0) : (CODIM == 1 ? fb_member<SD>(p, i + 1, 0) : fb_member<SD>(p, q, i + 1));
            const int ip = CODIM == 0 ? fb_member<SD>(i - 1, 0, 0) : (CODIM == 1 ? fb_member<SD>(p, i - 1, 0) : fb_member<SD>(p, q, i - 1));
            const double a = st.abc[s][0], b = st.abc[s][1], c = st.abc[s][2];
            ++s;
            const double F = a * fav - b * fbv;
            double dF[3], dG[3], ddG[6], g1[3];
#pragma unroll
            for (int d = 0; d < SD; ++d) dF[d] = a * dfa[d] - b * dfb[d];
            if (i == 0) {
                J::first(P, J::NA, T[in], T[ic], F, dF);
            } else {
                const double G = -c * fcv;
#pragma unroll
                for (int d = 0; d < SD; ++d) {
                    g1[d] = -2.0 * c * dfb[d];
                    dG[d] = fbv * g1[d];
                }
                int k = 0;
#pragma unroll
                for (int d1 = 0; d1 < SD; ++d1)
#pragma unroll
                    for (int d2 = d1; d2 < SD; ++d2) ddG[k++] = g1[d1] * dfb[d2];
                J::three(P, J::NA, T[in], T[ic], T[ip > 0 ? ip : 0], F, dF, G, dG, ddG);
            }
        }
    }
}

template <int SD, int N, int ORDER>
__global__ void __launch_bounds__(128, 3)
k_small(const DevSimplex P, const __grid_constant__ SmallTab st, const DevEntity E, const double* __restrict__ pts,
        long long npts, long long ldp, double* __restrict__ out, long long ostride, const __grid_constant__ DevRowMap M) {
    constexpr int NMEM = fb_binom(N + SD, SD);
    constexpr int NA = Jet<SD, ORDER>::NA;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * blockDim.x + tid;
    double* s_C = smem;                                              // ncells x nrows x NMEM
    double* s_g = s_C + (size_t)P.ncells * P.nrows * NMEM;           // ncells x 32
    for (int i = tid; i < P.ncells * P.nrows * NMEM; i += blockDim.x) s_C[i] = __ldg(P.ccell_morton + i);
    for (int i = tid; i < P.ncells * FB_GEOM_DOUBLES; i += blockDim.x) s_g[i] = __ldg(P.geom + i);
    __syncthreads();
    if (p >= npts) return;

    double x[3];
    apply_entity<SD>(E, pts + p * ldp, x);
    unsigned mask = locate_cells<SD>(st.bary, P.ncells, P.unique, x);
    if (mask == 0) {            // in no subcell: zero column, like the reference
        fb_zero_column(M, out, ostride, p, NA, P.nrows);
        return;
    }
    const double inv_mult = 1.0 / (double)__popc(mask);
    bool first = true;
    while (mask) {
        const int cell = __ffs(mask) - 1;
        mask &= mask - 1;
        const double* geom = s_g + cell * FB_GEOM_DOUBLES;
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
        double T[NMEM][NA];
#pragma unroll
        for (int m = 0; m < NMEM; ++m)
#pragma unroll
            for (int a = 0; a < NA; ++a) T[m][a] = 0.0;
        T[0][0] = geom[12] * inv_mult;
        int s = 0;
        {   // pass 0
            double dfa[3] = {0, 0, 0}, dfb[3] = {0, 0, 0};
#pragma unroll
            for (int d = 0; d < SD; ++d) { dfa[d] = geom[14 + d]; dfb[d] = geom[23 + d]; }
            small_chain<SD, N, ORDER, 0>(P, st, s, 0, 0, T, fa[0], fb[0], dfa, dfb);
        }
        if (SD >= 2) {   // pass 1: chains (p, .)
            double dfa[3] = {0, 0, 0}, dfb[3] = {0, 0, 0};
#pragma unroll
            for (int d = 0; d < SD; ++d) { dfa[d] = geom[14 + 3 + d]; dfb[d] = geom[23 + 3 + d]; }
#pragma unroll
            for (int pp = 0; pp < N; ++pp) small_chain<SD, N, ORDER, 1>(P, st, s, pp, 0, T, fa[1], fb[1], dfa, dfb);
        }
        if (SD >= 3) {   // pass 2: chains (p, q, .), q outermost
            double dfa[3] = {0, 0, 0}, dfb[3] = {0, 0, 0};
#pragma unroll
            for (int d = 0; d < SD; ++d) { dfa[d] = geom[14 + 6 + d]; dfb[d] = geom[23 + 6 + d]; }
#pragma unroll
            for (int qq = 0; qq < N; ++qq)
#pragma unroll
                for (int pp = 0; pp < N; ++pp)
                    if (pp + qq < N) small_chain<SD, N, ORDER, 2>(P, st, s, pp, qq, T, fa[2], fb[2], dfa, dfb);
        }

        // contraction with the subcell's coefficient matrix, RB rows at a time
        constexpr int RB = 4;
        const double* C = s_C + (size_t)cell * P.nrows * NMEM;
        for (int r0 = 0; r0 < P.nrows; r0 += RB) {
            double acc[RB][NA];
#pragma unroll
            for (int j = 0; j < RB; ++j)
#pragma unroll
                for (int a = 0; a < NA; ++a) acc[j][a] = 0.0;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const double* Cr = C + (size_t)min(r0 + j, P.nrows - 1) * NMEM;
#pragma unroll
                for (int m = 0; m < NMEM; ++m) {
                    const double c = Cr[m];
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[j][a] = fma(c, T[m][a], acc[j][a]);
                }
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                if (r0 + j < P.nrows) {
                    double sgn;
                    const size_t orow = fb_map_row(M, r0 + j, sgn);
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        double* o = out + ((size_t)a * M.total_rows + orow) * ostride + p;
                        *o = first ? sgn * acc[j][a] : (*o + sgn * acc[j][a]);
                    }
                }
            }
        }
        first = false;
    }
}
