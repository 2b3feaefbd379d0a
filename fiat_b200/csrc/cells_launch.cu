// Launchers of the split-cell tile kernels (cells_reg.cuh, cells.cuh); their own translation unit (parallel compilation).
#include "host_plan.cuh"
#include "cells.cuh"
#include "cells_reg.cuh"

namespace {

bool cells_geometry(const fiatb200_plan* plan, CellsGeom* G, size_t* smem_out) {
    const DevSimplex& P = plan->simplex;
    if (P.blk_cells < 2 || P.blk_cells != P.ncells || P.expansion != 0 || P.order != 0 || P.nblk == 0 || plan->tab.nrb == 0)
        return false;
    if (P.sd < 2) return false;
    const int maxlev = std::max(1, plan->tab.nsteps);        // all step records sit in shared memory (cells.cuh)
    if (P.ncells > 32) return false;
    // widest tile first (the per-block loop overhead is amortised over the octets a subcell has in the tile):
    // 256 threads and two CTAs per SM, or 512 threads and one CTA with all of the shared memory
    const size_t budget = (size_t)plan->max_smem_optin - 1024;
    for (int pt = 128; pt >= 32; pt -= 32) {
        for (int pass = 0; pass < 2; ++pass) {
            const int threads = pass == 0 ? 256 : 512;
            const size_t limit = pass == 0 ? (size_t)110 * 1024 : budget;
            const int pts_cap = pt + 8 * P.ncells;
            int ld = pts_cap;
            while ((ld & 15) != 4 && (ld & 15) != 12) ++ld;
            const size_t bytes = ((size_t)P.kpad * ld + 6 * pts_cap) * sizeof(double) + (size_t)maxlev * sizeof(StepRec)
                                 + (size_t)pts_cap * sizeof(int) + (size_t)(threads / 32) * 8 * (pt + 2) * sizeof(double)
                                 + ((size_t)P.ncells * (plan->tab.nrb + 1) + pts_cap / 8 + 32) * sizeof(int) + 64;
            if (bytes <= limit) {
                G->PT = pt; G->PTS = pts_cap; G->ldT = ld; G->maxlev = maxlev; G->threads = threads;
                *smem_out = bytes;
                return true;
            }
        }
    }
    return false;
}

// register-operand kernel (cells_reg.cuh): 16 warps x 2 octets of columns, fixed-k block stream
bool cells_reg_geometry(const fiatb200_plan* plan, CellsRegGeom* G, size_t* smem_out) {
    const DevSimplex& P = plan->simplex;
    // Measured against the segment kernel (2^20 points, ms): Guzman-Neilan (20 members per subcell) 2.05 / 2.24,
    // Alfeld-Sorokina (10) 2.33 / 3.11 -- but HCT degree 5 (21) 0.70 / 0.66, degree 6 (28) 1.02 / 0.88, Walkington
    // (56) 2.89 / 2.71: with many members the segments of cells.cuh are long enough.  FIATB200_CELLS_REG=0 / 1 force
    // the segment / this kernel (experiments).
    if (fb_tuning().cells_reg == 0) return false;
    if (fb_tuning().cells_reg != 1 && P.kpad > 20) return false;
    if (P.cnsteps <= 0 || P.crb <= 0 || P.kpad > 64 || P.ncells < 2 || P.ncells > 24 || P.sd < 2) return false;
    if (P.expansion != 0 || P.order != 0 || plan->tab.nrb == 0 || plan->tab.nrb > P.cnsteps * P.crb) return false;
    // one CTA of 16 warps per SM; two CTAs of 8 warps (FIATB200_CELLS_THREADS=256: one CTA's binning, recurrence
    // and barriers would overlap the other's steps) measured 6-18 % slower on every element
    int threads = 512;
    if (fb_tuning().cells_threads == 256 || fb_tuning().cells_threads == 512) threads = fb_tuning().cells_threads;
    if (threads == 256 && 128 - 8 * P.ncells < 64) threads = 512;    // many subcells: the padding would eat the tile
    const int pts_cap = (threads / 32) * 16;
    const int pt = pts_cap - 8 * P.ncells;                        // subcell ranges are padded to octets
    if (pt < 32) return false;
    int ld = pts_cap;
    while ((ld & 15) != 8) ++ld;
    const int maxlev = std::max(1, plan->tab.nsteps);
    const int sp = pt + 2;
    const int astage = (P.cmaxstep + 1) & ~1;
    const int uoff = ((pts_cap + pts_cap / 8 + 1) / 2 + 1) & ~1;
    const size_t phase01 = (size_t)P.kpad * ld + 6 * (size_t)pts_cap + 4 * (size_t)maxlev;
    const size_t phase2 = 2 * (size_t)P.crb * 8 * sp + 2 * (size_t)astage;
    const size_t bytes = ((size_t)uoff + std::max(phase01, phase2)) * sizeof(double) + 64;
    if (bytes > (size_t)plan->max_smem_optin - 1024) return false;
    G->PT = pt; G->PTS = pts_cap; G->ldT = ld; G->maxlev = maxlev; G->threads = threads; G->SP = sp;
    G->astage = astage; G->uoff = uoff;
    *smem_out = bytes;
    return true;
}

template <int SD, int KB>
int launch_cells_reg(const fiatb200_plan* plan, const DevEntity& E, const CellsRegGeom& G, size_t smem, const double* pts,
                     long long npts, long long ldp, double* out, long long ostride, cudaStream_t st) {
    int rc = fb_set_smem(k_cells_reg<SD, KB>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((npts + G.PT - 1) / G.PT);
    k_cells_reg<SD, KB><<<grid, G.threads, smem, st>>>(plan->simplex, plan->tab, plan->small_tab, E, G, pts, npts, ldp,
                                                              out, ostride);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

template <int SD, int CH>
int launch_cells(const fiatb200_plan* plan, const DevEntity& E, const CellsGeom& G, size_t smem, const double* pts,
                 long long npts, long long ldp, double* out, long long ostride, cudaStream_t st) {
    int rc = fb_set_smem(k_mma_cells<SD, CH>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((npts + G.PT - 1) / G.PT);
    k_mma_cells<SD, CH><<<grid, G.threads, smem, st>>>(plan->simplex, plan->tab, plan->small_tab, E, G, pts, npts, ldp,
                                                              out, ostride);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

}  // namespace

bool fb_cells_applicable(const fiatb200_plan* plan) {
    CellsGeom G;
    CellsRegGeom R;
    size_t smem = 0;
    return cells_reg_geometry(plan, &R, &smem) || cells_geometry(plan, &G, &smem);
}

int fb_dispatch_cells(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                      double* out, long long ostride, cudaStream_t st) {
    CellsGeom G;
    size_t smem = 0;
    CellsRegGeom R;
    if (cells_reg_geometry(plan, &R, &smem)) {
        const int kb = plan->simplex.kpad / 4;
#define FB_CELLS_REG_LAUNCH(SD_)                                                                                   \
    if (kb <= 3) return launch_cells_reg<SD_, 3>(plan, E, R, smem, pts, npts, ldp, out, ostride, st);              \
    if (kb <= 6) return launch_cells_reg<SD_, 6>(plan, E, R, smem, pts, npts, ldp, out, ostride, st);              \
    if (kb <= 9) return launch_cells_reg<SD_, 9>(plan, E, R, smem, pts, npts, ldp, out, ostride, st);              \
    if (kb <= 14) return launch_cells_reg<SD_, 14>(plan, E, R, smem, pts, npts, ldp, out, ostride, st);            \
    return launch_cells_reg<SD_, 16>(plan, E, R, smem, pts, npts, ldp, out, ostride, st);
        if (plan->simplex.sd == 2) { FB_CELLS_REG_LAUNCH(2) }
        FB_CELLS_REG_LAUNCH(3)
#undef FB_CELLS_REG_LAUNCH
    }
    if (!cells_geometry(plan, &G, &smem))
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "split-cell tile kernel not applicable to this plan");
    // fragments prefetched per segment: enough for the plan's longest (row block, subcell) segment, at most 8
    const int seg = plan->max_segment;
#define FB_CELLS_LAUNCH(SD_)                                                                                   \
    if (seg <= 2) return launch_cells<SD_, 2>(plan, E, G, smem, pts, npts, ldp, out, ostride, st);             \
    if (seg <= 4) return launch_cells<SD_, 4>(plan, E, G, smem, pts, npts, ldp, out, ostride, st);             \
    return launch_cells<SD_, 8>(plan, E, G, smem, pts, npts, ldp, out, ostride, st);
    if (plan->simplex.sd == 2) { FB_CELLS_LAUNCH(2) }
    FB_CELLS_LAUNCH(3)
#undef FB_CELLS_LAUNCH
}
