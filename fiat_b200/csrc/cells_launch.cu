// Launcher of the split-cell tile kernel (cells.cuh); its own translation unit (parallel compilation).
#include "host_plan.cuh"
#include "cells.cuh"

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
    size_t smem = 0;
    return cells_geometry(plan, &G, &smem);
}

int fb_dispatch_cells(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                      double* out, long long ostride, cudaStream_t st) {
    CellsGeom G;
    size_t smem = 0;
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
