// Launcher of the register-resident jet kernel (small.cuh); its own translation unit so that the
// many template instantiations compile in parallel with the rest of the library.
#include "host_plan.cuh"
#include "small.cuh"

namespace {

// ---- register kernel for low-degree elements ------------------------------------------------------
template <int SD, int N, int ORDER>
int launch_small(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                 double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevSimplex& P = plan->simplex;
    const size_t smem = ((size_t)P.ncells * P.nrows * P.nslots + (size_t)P.ncells * FB_GEOM_DOUBLES) * sizeof(double);
    int rc = fb_set_smem(k_small<SD, N, ORDER>, smem);
    if (rc) return rc;
    const int bp = 128;
    const unsigned grid = (unsigned)((npts + bp - 1) / bp);
    k_small<SD, N, ORDER><<<grid, bp, smem, st>>>(P, plan->small_tab, E, pts, npts, ldp, out, ostride, M);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

}  // namespace

// (sd, degree, order) combinations whose expansion table fits in registers (members x alphas <= 64)
bool fb_small_applicable(const fiatb200_plan* plan) {
    const DevSimplex& P = plan->simplex;
    if (P.expansion != 0 || P.order > 3 || P.degree < 1 || P.sd < 2) return false;
    if (P.order == 3 && !(P.sd == 2 && P.degree <= 2)) return false;
    if ((size_t)P.nslots * P.na > 64) return false;
    if (P.sd == 2 && P.degree > 4) return false;
    if (P.sd == 3 && P.degree > 3) return false;
    const size_t smem = ((size_t)P.ncells * P.nrows * P.nslots + (size_t)P.ncells * FB_GEOM_DOUBLES) * sizeof(double);
    return smem <= 64 * 1024;
}

#define FB_SMALL_CASE(SD_, N_, O_)                                                        \
    if (P.sd == SD_ && P.degree == N_ && P.order == O_)                                   \
        return launch_small<SD_, N_, O_>(plan, E, pts, npts, ldp, out, ostride, M, st);

int fb_dispatch_small(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                   double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevSimplex& P = plan->simplex;
    FB_SMALL_CASE(2, 1, 0) FB_SMALL_CASE(2, 1, 1) FB_SMALL_CASE(2, 1, 2)
    FB_SMALL_CASE(2, 2, 0) FB_SMALL_CASE(2, 2, 1) FB_SMALL_CASE(2, 2, 2)
    FB_SMALL_CASE(2, 3, 0) FB_SMALL_CASE(2, 3, 1) FB_SMALL_CASE(2, 3, 2)
    FB_SMALL_CASE(2, 4, 0) FB_SMALL_CASE(2, 4, 1)
    FB_SMALL_CASE(2, 1, 3) FB_SMALL_CASE(2, 2, 3)
    FB_SMALL_CASE(3, 1, 0) FB_SMALL_CASE(3, 1, 1) FB_SMALL_CASE(3, 1, 2)
    FB_SMALL_CASE(3, 2, 0) FB_SMALL_CASE(3, 2, 1)
    FB_SMALL_CASE(3, 3, 0)
    return fb_fail(FIATB200_ERR_UNSUPPORTED, "register kernel not instantiated for this element");
}

