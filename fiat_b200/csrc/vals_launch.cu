// Launcher of the value-table kernel (vals.cuh); its own translation unit (parallel compilation).
#include "host_plan.cuh"
#include "vals.cuh"

// ---- value-table kernel (derivative-folded coefficients) ----------------------------------------
bool fb_vals_applicable(const fiatb200_plan* plan) {
    const DevSimplex& P = plan->simplex;
    if (P.expansion != 0 || P.ncp == 0 || P.cderiv_len == 0 || P.order > 3 || P.degree < 1) return false;
    if (P.sd == 2 ? P.degree > 6 : (P.sd == 3 ? P.degree > 4 : true)) return false;
    if (plan->tab.nsteps > FB_SMALL_MAX_STEPS) return false;
    const size_t smem = ((size_t)P.cderiv_len + (size_t)P.ncells * FB_GEOM_DOUBLES) * sizeof(double);
    return smem <= 36 * 1024;
}

template <int SD, int N, int NCP, int J, bool IDENT>
int launch_vals_id(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                   double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevSimplex& P = plan->simplex;
    const size_t smem = ((size_t)P.cderiv_len + (size_t)P.ncells * FB_GEOM_DOUBLES) * sizeof(double);
    int rc = fb_set_smem(k_vals<SD, N, NCP, J, IDENT>, smem);
    if (rc) return rc;
    const int bp = 128;
    // a CTA stages the coefficient table once and then walks `tpc` consecutive point tiles
    int tpc = smem > 16 * 1024 ? 4 : (smem > 4 * 1024 ? 2 : 1);
    if (fb_tuning().vals_tpc >= 0) tpc = std::max(1, fb_tuning().vals_tpc);
    const long long per_cta = (long long)bp * tpc;
    const unsigned grid = (unsigned)((npts + per_cta - 1) / per_cta);
    k_vals<SD, N, NCP, J, IDENT><<<grid, bp, smem, st>>>(P, plan->small_tab, E, pts, npts, ldp, tpc, out, ostride, M);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

// Jet order J (see the kernel comment in vals.cuh).  Every coefficient load costs two shared-memory
// wavefronts per warp; a warp's share of the HBM write time is 32 points * 8 B * values/point at
// ~23.3 B per SM-cycle (6.55 TB/s over 148 SMs at 1.9 GHz).  J = 0 has the fewest FMAs and registers
// and is kept unless its loads alone would take more than 70 % of that time (measured on B200:
// PS6/PS12 order 2 run faster with J = 0, HCT / Arnold-Winther / BDM with J = 1).
int vals_jet_order(const DevSimplex& P) {
    if (fb_tuning().vals_j >= 0) return std::min(P.order, fb_tuning().vals_j > 0 ? 1 : 0);
    if (P.order < 1) return 0;
    const int sd = P.sd, n = P.degree;
    double loads = 0.0;
    for (int k = 0; k <= P.order; ++k)
        loads += (double)fb_binom(sd + k - 1, k) * P.nrows * (k <= n ? fb_binom(n - k + sd, sd) : 0);
    const double hbm_cycles = 32.0 * 8.0 * P.na * P.nrows / 23.3;
    return 2.0 * loads > 0.7 * hbm_cycles ? 1 : 0;
}

template <int SD, int N, int NCP>
int launch_vals(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    if (!M.identity) return launch_vals_id<SD, N, NCP, 0, false>(plan, E, pts, npts, ldp, out, ostride, M, st);
    if (vals_jet_order(plan->simplex) == 1)
        return launch_vals_id<SD, N, NCP, 1, true>(plan, E, pts, npts, ldp, out, ostride, M, st);
    return launch_vals_id<SD, N, NCP, 0, true>(plan, E, pts, npts, ldp, out, ostride, M, st);
}

template <int SD, int N>
int dispatch_vals_ncp(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                      double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    switch (plan->simplex.ncp) {
        case 1: return launch_vals<SD, N, 1>(plan, E, pts, npts, ldp, out, ostride, M, st);
        case 4: return launch_vals<SD, N, 4>(plan, E, pts, npts, ldp, out, ostride, M, st);
        case 16: return launch_vals<SD, N, 16>(plan, E, pts, npts, ldp, out, ostride, M, st);
    }
    return fb_fail(FIATB200_ERR_ARG, "bad subcell stride in the derivative-folded coefficient table");
}

int fb_dispatch_vals(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                  double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevSimplex& P = plan->simplex;
#define FB_VALS_CASE(SD_, N_) \
    if (P.sd == SD_ && P.degree == N_) return dispatch_vals_ncp<SD_, N_>(plan, E, pts, npts, ldp, out, ostride, M, st);
    FB_VALS_CASE(2, 1) FB_VALS_CASE(2, 2) FB_VALS_CASE(2, 3) FB_VALS_CASE(2, 4) FB_VALS_CASE(2, 5) FB_VALS_CASE(2, 6)
    FB_VALS_CASE(3, 1) FB_VALS_CASE(3, 2) FB_VALS_CASE(3, 3) FB_VALS_CASE(3, 4)
#undef FB_VALS_CASE
    return fb_fail(FIATB200_ERR_UNSUPPORTED, "value-table kernel not instantiated for this element");
}

