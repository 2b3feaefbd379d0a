// Per-point expansion arithmetic shared by all kernels: entity transform, split-cell location,
// Dubiner / integrated-Jacobi recurrence with derivative jets, C0 fix-ups, 1-D sets.
//
// Follows the reference's arithmetic (paths relative to the FIAT source tree):
//   entity transform            FIAT/reference_element.py:570-609
//   l1 distance / binning       FIAT/reference_element.py:616-644,779-780; FIAT/expansions.py:771-811
//   recurrence factors          FIAT/expansions.py:54-63
//   three-term recurrence       FIAT/expansions.py:202-249 (Leibniz rule: :66-137)
//   C0 fix-ups                  FIAT/expansions.py:281-295
//   Legendre line set           FIAT/expansions.py:659-678, FIAT/jacobi.py:47-74
//   Lagrange line set           FIAT/barycentric_interpolation.py:22-47
// The per-pass normalisation (:251-266) and the C0 reordering (:297-322) are folded into the
// coefficient matrix on the host (fiat_b200/plan.py).
#pragma once
#include "device_plan.cuh"

// FP64 tensor-pipe MMA: D(8x8) += A(8x4) B(4x8); lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) .. +1]
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// entity transform: x_cell = x_entity * C + offset
// ---------------------------------------------------------------------------------------------
template <int SD>
__device__ __forceinline__ void apply_entity(const DevEntity& E, const double* __restrict__ pt, double (&x)[3]) {
    x[0] = x[1] = x[2] = 0.0;
    if (E.identity) {
#pragma unroll
        for (int j = 0; j < SD; ++j) x[j] = pt[j];
        return;
    }
    double e[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < 3; ++d)
        if (d < E.dim) e[d] = pt[d];
#pragma unroll
    for (int j = 0; j < SD; ++j) {
        double s = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) s = fma(e[d], E.C[d * SD + j], s);   // C rows beyond E.dim are zero
        x[j] = s + E.off[j];
    }
}

// ---------------------------------------------------------------------------------------------
// split-cell location.  No FMA contraction here: the comparison against best + 1e-12 must bin
// points exactly like the reference does.
// ---------------------------------------------------------------------------------------------
template <int SD>
__device__ __forceinline__ double l1_distance(const double* __restrict__ rows, const double (&x)[3]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i <= SD; ++i) {
        const double* r = rows + 4 * i;
        double lam = __dmul_rn(x[0], r[0]);
#pragma unroll
        for (int d = 1; d < SD; ++d) lam = __dadd_rn(lam, __dmul_rn(x[d], r[d]));
        lam = __dadd_rn(lam, r[3]);
        const double t = __dadd_rn(fabs(lam), -lam);
        s = (i == 0) ? t : __dadd_rn(s, t);
    }
    return __dmul_rn(0.5, fabs(s));
}

template <int SD>
__device__ __forceinline__ unsigned locate_cells(const double* __restrict__ bary, int ncells, int unique,
                                                 const double (&x)[3]) {
    if (ncells == 1) return 1u;
    const double best = l1_distance<SD>(bary + 16 * ncells, x);
    const double tol = __dadd_rn(best, 1e-12);
    unsigned mask = 0;
    for (int c = 0; c < ncells; ++c) {
        if (l1_distance<SD>(bary + 16 * c, x) < tol) {
            mask |= 1u << c;
            if (unique) break;
        }
    }
    return mask;
}

// ---------------------------------------------------------------------------------------------
// derivative jets.  Component order = mis(SD,0), mis(SD,1), mis(SD,2): value, gradient, then the
// upper triangle of the Hessian row by row.  ORDER = -1 selects the table-driven generic path.
// ---------------------------------------------------------------------------------------------
// Position of the multi-index (a0, a1, a2) in mis(SD,0), mis(SD,1), ... (first entry descending).
template <int SD>
__host__ __device__ constexpr int fb_alpha_index(int a0, int a1, int a2) {
    const int k = a0 + a1 + a2;
    if (SD == 1) return k;
    if (SD == 2) return k * (k + 1) / 2 + a1;
    return k * (k + 1) * (k + 2) / 6 + (a1 + a2) * (a1 + a2 + 1) / 2 + a2;
}

// Leibniz rule D^alpha(F g) for an affine F and D^alpha(G h) for a quadratic G (constant Hessian
// ddG, upper triangle row by row), for every |alpha| <= ORDER.  All loops unroll at compile time,
// so the multi-index arithmetic folds to constants and the jets stay in registers.
template <int SD, int ORDER>
struct Jet {
    static constexpr int NA = fb_binom(SD + ORDER, ORDER);
    static constexpr int CAP = NA;

    template <bool WITH_PREV>
    __device__ __forceinline__ static void apply(double* __restrict__ nx, const double* __restrict__ cu,
                                                 const double* __restrict__ pv, double F,
                                                 const double* __restrict__ dF, double G,
                                                 const double* __restrict__ dG, const double* __restrict__ ddG) {
#pragma unroll
        for (int a0 = 0; a0 <= ORDER; ++a0) {
#pragma unroll
            for (int a1 = 0; a1 <= (SD >= 2 ? ORDER - a0 : 0); ++a1) {
#pragma unroll
                for (int a2 = 0; a2 <= (SD >= 3 ? ORDER - a0 - a1 : 0); ++a2) {
                    const int al[3] = {a0, a1, a2};
                    const int j = fb_alpha_index<SD>(a0, a1, a2);
                    double v = F * cu[j];
                    if (WITH_PREV) v = fma(G, pv[j], v);
#pragma unroll
                    for (int d = 0; d < SD; ++d) {
                        if (al[d] >= 1) {
                            const int lo = fb_alpha_index<SD>(a0 - (d == 0), a1 - (d == 1), a2 - (d == 2));
                            const double m = (double)al[d];
                            v = fma(m * dF[d], cu[lo], v);
                            if (WITH_PREV) v = fma(m * dG[d], pv[lo], v);
                        }
                    }
                    if (WITH_PREV) {
                        int k = 0;
#pragma unroll
                        for (int d1 = 0; d1 < SD; ++d1) {
#pragma unroll
                            for (int d2 = d1; d2 < SD; ++d2) {
                                const int need = (d1 == d2) ? 2 : 1;
                                if (al[d1] >= need && al[d2] >= need) {
                                    const int lo = fb_alpha_index<SD>(a0 - (d1 == 0) - (d2 == 0), a1 - (d1 == 1) - (d2 == 1),
                                                                      a2 - (d1 == 2) - (d2 == 2));
                                    const double m = (d1 == d2) ? (double)(al[d1] * (al[d1] - 1) / 2)
                                                                : (double)(al[d1] * al[d2]);
                                    v = fma(m * ddG[k], pv[lo], v);
                                }
                                ++k;
                            }
                        }
                    }
                    nx[j] = v;
                }
            }
        }
    }

    __device__ __forceinline__ static void first(const DevSimplex&, int, double* __restrict__ nx,
                                                 const double* __restrict__ cu, double F,
                                                 const double* __restrict__ dF) {
        apply<false>(nx, cu, cu, F, dF, 0.0, dF, dF);
    }

    __device__ __forceinline__ static void three(const DevSimplex&, int, double* __restrict__ nx,
                                                 const double* __restrict__ cu, const double* __restrict__ pv,
                                                 double F, const double* __restrict__ dF, double G,
                                                 const double* __restrict__ dG, const double* __restrict__ ddG) {
        apply<true>(nx, cu, pv, F, dF, G, dG, ddG);
    }
};

template <int SD>
struct Jet<SD, -1> {
    static constexpr int NA = -1;
    static constexpr int CAP = FB_NA_MAX;
    static constexpr int NPAIR = SD * (SD + 1) / 2;

    __device__ static void first(const DevSimplex& P, int na, double* __restrict__ nx,
                                 const double* __restrict__ cu, double F, const double* __restrict__ dF) {
        for (int j = 0; j < na; ++j) {
            double v = F * cu[j];
            for (int d = 0; d < SD; ++d) {
                const int lo = __ldg(P.low1 + 3 * j + d);
                if (lo >= 0) v = fma(__ldg(P.mul1 + 3 * j + d) * dF[d], cu[lo], v);
            }
            nx[j] = v;
        }
    }

    __device__ static void three(const DevSimplex& P, int na, double* __restrict__ nx,
                                 const double* __restrict__ cu, const double* __restrict__ pv, double F,
                                 const double* __restrict__ dF, double G, const double* __restrict__ dG,
                                 const double* __restrict__ ddG) {
        for (int j = 0; j < na; ++j) {
            double v = fma(F, cu[j], G * pv[j]);
            for (int d = 0; d < SD; ++d) {
                const int lo = __ldg(P.low1 + 3 * j + d);
                if (lo >= 0) {
                    const double m = __ldg(P.mul1 + 3 * j + d);
                    v = fma(m * dF[d], cu[lo], v);
                    v = fma(m * dG[d], pv[lo], v);
                }
            }
            for (int k = 0; k < NPAIR; ++k) {
                const int lo = __ldg(P.low2 + 6 * j + k);
                if (lo >= 0) v = fma(__ldg(P.mul2 + 6 * j + k) * ddG[k], pv[lo], v);
            }
            nx[j] = v;
        }
    }
};

// Recurrence factors of the three collapsing passes at one point: X = (x, -1, -1),
// fb = (X[c+1] + X[c+2]) / 2, fa = X[c] + (fb + 1).
template <int SD>
__device__ __forceinline__ void recurrence_factors(const double (&x)[3], double (&fa)[3], double (&fb)[3]) {
    double X[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
#pragma unroll
    for (int i = 0; i < SD; ++i) X[i] = x[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        fb[c] = 0.5 * (X[c + 1] + X[c + 2]);
        fa[c] = X[c] + (fb[c] + 1.0);
    }
}

__device__ __forceinline__ double pick3(const double (&v)[3], int i) {
    return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]);
}

// One recurrence step for one point.  T addresses member slot s, jet component a at
// T[s * slot_stride + a * comp_stride]; `geom` is the cell's geometry record.
template <int SD, int ORDER>
__device__ __forceinline__ void run_step(const DevSimplex& P, const StepRec& r, const double* __restrict__ geom,
                                         const double (&fa)[3], const double (&fb)[3], double* __restrict__ T,
                                         int slot_stride, int comp_stride, int na) {
    typedef Jet<SD, ORDER> J;
    double cur[J::CAP], prv[J::CAP], nxt[J::CAP];
    const int codim = r.codim;
    const double fav = pick3(fa, codim), fbv = pick3(fb, codim);
    const double* src = T + (size_t)r.cur * slot_stride;
#pragma unroll
    for (int a = 0; a < J::CAP; ++a)
        if (J::NA > 0 || a < na) cur[a] = src[a * comp_stride];
    const double F = r.a * fav - r.b * fbv;
    double dfb[3], dF[3];
#pragma unroll
    for (int d = 0; d < SD; ++d) {
        dfb[d] = geom[23 + 3 * codim + d];
        dF[d] = r.a * geom[14 + 3 * codim + d] - r.b * dfb[d];
    }
    if (r.prv < 0) {
        J::first(P, na, nxt, cur, F, dF);
    } else {
        const double* srp = T + (size_t)r.prv * slot_stride;
#pragma unroll
        for (int a2 = 0; a2 < J::CAP; ++a2)
            if (J::NA > 0 || a2 < na) prv[a2] = srp[a2 * comp_stride];
        const double G = -r.c * (fbv * fbv);
        double g1[3], dG[3], ddG[6];
#pragma unroll
        for (int d = 0; d < SD; ++d) {
            g1[d] = -2.0 * r.c * dfb[d];
            dG[d] = fbv * g1[d];
        }
        int k = 0;
#pragma unroll
        for (int d1 = 0; d1 < SD; ++d1)
#pragma unroll
            for (int d2 = d1; d2 < SD; ++d2) ddG[k++] = g1[d1] * dfb[d2];
        J::three(P, na, nxt, cur, prv, F, dF, G, dG, ddG);
    }
    double* dst = T + (size_t)r.nxt * slot_stride;
#pragma unroll
    for (int a3 = 0; a3 < J::CAP; ++a3)
        if (J::NA > 0 || a3 < na) dst[a3 * comp_stride] = nxt[a3];
}

// Whole Dubiner expansion of one cell at one point, thread-private column of T.
template <int SD, int ORDER>
__device__ __forceinline__ void dubiner_point(const DevSimplex& P, const RecTab& tab, const double* __restrict__ geom,
                                              double start, const double (&xref)[3], double* __restrict__ T,
                                              int slot_stride, int comp_stride, int na) {
    double fa[3], fb[3];
    recurrence_factors<SD>(xref, fa, fb);
    double* T0 = T + (size_t)tab.start_slot * slot_stride;
    T0[0] = start;
    for (int a = 1; a < na; ++a) T0[a * comp_stride] = 0.0;
    for (int s = 0; s < tab.nsteps; ++s)
        run_step<SD, ORDER>(P, tab.steps[s], geom, fa, fb, T, slot_stride, comp_stride, na);
    for (int gi = 0; gi < tab.nfixgrp; ++gi) {
        double* t = T + (size_t)tab.fix_tgt[gi] * slot_stride;
        for (int f = tab.fix_first[gi]; f < tab.fix_first[gi] + tab.fix_cnt[gi]; ++f) {
            const double w = tab.fix_w[f];
            const double* s = T + (size_t)tab.fix_src[f] * slot_stride;
            for (int a = 0; a < na; ++a) t[a * comp_stride] = fma(-w, s[a * comp_stride], t[a * comp_stride]);
        }
    }
}

// Legendre set on a line (variant None): k-th derivative = Jacobi(k,k) times a running scale.
__device__ __forceinline__ void legendre_line_point(const DevSimplex& P, int cell, double inv_mult, double xin,
                                                    double* __restrict__ T, int slot_stride, int comp_stride) {
    const int n = P.line_n, order = P.order;
    const double* geom = P.geom + cell * FB_GEOM_DOUBLES;
    const double x = fma(xin, __ldg(geom + 0), __ldg(geom + 9));
    const double* rec = P.line_tab;
    const double* scales = P.line_tab + (size_t)(order + 1) * (n + 1) * 4 + (size_t)cell * (order + 1) * (n + 1);
    for (int k = 0; k <= order; ++k) {
        const double* rk = rec + (size_t)k * (n + 1) * 4;
        const double* sk = scales + (size_t)k * (n + 1);
        for (int p = 0; p < k && p <= n; ++p) T[(size_t)p * slot_stride + k * comp_stride] = 0.0;
        double pm2 = 0.0, pm1 = 1.0;
        for (int j = 0; j + k <= n; ++j) {
            double v;
            if (j == 0) {
                v = 1.0;
            } else if (j == 1) {
                v = __ldg(rk + 4) + __ldg(rk + 5) * x;
            } else {
                v = (__ldg(rk + 4 * j) + __ldg(rk + 4 * j + 1) * x) * pm1 - __ldg(rk + 4 * j + 2) * pm2;
            }
            pm2 = pm1;
            pm1 = v;
            T[(size_t)(j + k) * slot_stride + k * comp_stride] = v * __ldg(sk + j + k) * inv_mult;
        }
    }
}

// Lagrange set on a line through arbitrary nodes: second barycentric formula, NaN -> 1 at nodes,
// r-th derivative = dmat^r phi.
__device__ __forceinline__ void lagrange_line_point(const DevSimplex& P, int cell, double inv_mult, double x,
                                                    double* __restrict__ T, int slot_stride, int comp_stride) {
    const int nn = P.line_n, order = P.order;
    const double* nodes = P.line_tab + (size_t)cell * (2 * nn + nn * nn);
    const double* wts = nodes + nn;
    const double* dmat = wts + nn;
    double sum = 0.0;
    for (int i = 0; i < nn; ++i) {
        const double t = (1.0 / (x - __ldg(nodes + i))) * __ldg(wts + i);
        T[(size_t)i * slot_stride] = t;
        sum += t;
    }
    const double inv = 1.0 / sum;
    for (int i = 0; i < nn; ++i) {
        double v = inv * T[(size_t)i * slot_stride];
        if (v != v) v = 1.0;
        T[(size_t)i * slot_stride] = v;
    }
    for (int r = 1; r <= order; ++r) {
        for (int i = 0; i < nn; ++i) {
            double s = 0.0;
            for (int j = 0; j < nn; ++j)
                s = fma(__ldg(dmat + i * nn + j), T[(size_t)j * slot_stride + (r - 1) * comp_stride], s);
            T[(size_t)i * slot_stride + r * comp_stride] = s;
        }
    }
    if (inv_mult != 1.0) {
        for (int i = 0; i < nn; ++i)
            for (int r = 0; r <= order; ++r) T[(size_t)i * slot_stride + r * comp_stride] *= inv_mult;
    }
}

// Expansion table of one (sub)cell at one point into the thread's column of T.
template <int SD, int ORDER>
__device__ __forceinline__ void expansion_point(const DevSimplex& P, const RecTab& tab, const double* __restrict__ geom,
                                                int cell, double inv_mult, const double (&x)[3],
                                                double* __restrict__ T, int slot_stride, int comp_stride, int na) {
    if (P.expansion == 0) {
        double xr[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < SD; ++i) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < SD; ++d) s = fma(x[d], geom[i * SD + d], s);
            xr[i] = s + geom[9 + i];
        }
        dubiner_point<SD, ORDER>(P, tab, geom, geom[12] * inv_mult, xr, T, slot_stride, comp_stride, na);
    } else if (SD == 1) {
        if (P.expansion == 1) legendre_line_point(P, cell, inv_mult, x[0], T, slot_stride, comp_stride);
        else lagrange_line_point(P, cell, inv_mult, x[0], T, slot_stride, comp_stride);
    }
}
