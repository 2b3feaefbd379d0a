// C ABI of the B200 tabulation library (see include/fiat_b200.h).
#include "host_plan.cuh"
#include "kernels.cuh"
#include "lattice.cuh"

namespace {
thread_local std::string g_error;
}
std::atomic<long long> fb_launches{0};

const FbTuning& fb_tuning() {
    static const FbTuning t = [] {
        auto geti = [](const char* name) {
            const char* env = getenv(name);
            return env ? atoi(env) : -1;
        };
        FbTuning v;
        v.mma_pt = geti("FIATB200_MMA_PT");
        v.mma_skip = geti("FIATB200_MMA_SKIP");
        v.mma_threads = geti("FIATB200_MMA_THREADS");
        v.tensor_bp = geti("FIATB200_TENSOR_BP");
        v.eval_bp = geti("FIATB200_EVAL_BP");
        v.eval_generic = geti("FIATB200_EVAL_GENERIC");
        v.vals_tpc = geti("FIATB200_VALS_TPC");
        v.vals_j = geti("FIATB200_VALS_J");
        v.mma_wl = geti("FIATB200_MMA_WARPLOCAL");
        v.cells_reg = geti("FIATB200_CELLS_REG");
        v.cells_threads = geti("FIATB200_CELLS_THREADS");
        return v;
    }();
    return t;
}

int fb_raise_smem_limit(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> limit;     // (device, kernel) -> bytes opted in
    int dev = 0;
    FB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> guard(mu);
    size_t& cur = limit[std::make_pair(dev, kernel)];
    if (bytes > cur) {
        FB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return FIATB200_OK;
}

int fb_fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

namespace {

// bump allocator over one host staging buffer mirrored to one device allocation
struct Arena {
    std::vector<unsigned char> host;
    size_t add(const void* src, size_t bytes) {
        size_t off = (host.size() + 255) & ~size_t(255);
        host.resize(off + bytes);
        if (bytes) memcpy(host.data() + off, src, bytes);
        return off;
    }
};

template <typename T>
const T* at(void* base, size_t off) {
    return reinterpret_cast<const T*>(static_cast<unsigned char*>(base) + off);
}

DevEntity make_entity(const fiatb200_entity_map* e, int sd) {
    DevEntity d;
    memset(&d, 0, sizeof(d));
    if (!e) {
        d.dim = sd;
        d.identity = 1;
        return d;
    }
    d.dim = e->dim;
    d.identity = e->identity;
    memcpy(d.C, e->C, sizeof(d.C));
    memcpy(d.off, e->offset, sizeof(d.off));
    return d;
}

DevRowMap make_row_map(const fiatb200_row_map* m, int ncomp, int nrows) {
    DevRowMap d;
    memset(&d, 0, sizeof(d));
    if (!m) {
        d.identity = 1;
        d.nc_in = d.nc_out = ncomp;
        d.dof_base = 0;
        d.total_rows = nrows;
        for (int k = 0; k < 9; ++k) { d.comp_out[k] = k; d.sign[k] = 1.0; }
        return d;
    }
    d.nc_in = m->nc_in; d.nc_out = m->nc_out; d.dof_base = m->dof_base; d.total_rows = m->total_rows;
    for (int k = 0; k < 9; ++k) { d.comp_out[k] = m->comp_out[k]; d.sign[k] = m->sign[k]; }
    return d;
}

// ---- thread-per-point launch -----------------------------------------------------------------
template <int SD, int ORDER>
int launch_cellwise(const fiatb200_plan* plan, const DevSimplex& P, const DevEntity& E, const double* pts, long long npts,
                    long long ldp, double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const size_t per_point = (size_t)P.nslots * P.na * sizeof(double);
    // widest block whose private expansion columns fit; very large elements end up with narrow blocks
    int bp = 128;
    while (bp > 32 && per_point * bp > 96 * 1024) bp >>= 1;
    while (bp > 8 && per_point * bp > (size_t)plan->max_smem_optin) bp >>= 1;
    size_t smem = per_point * bp;
    // small coefficient / geometry tables ride along in shared memory
    const size_t table_bytes = ((size_t)P.ncells * P.nrows * P.nslots + (size_t)P.ncells * FB_GEOM_DOUBLES) * sizeof(double);
    const int tables_in_smem = (table_bytes <= 32 * 1024 && smem + table_bytes <= (size_t)plan->max_smem_optin) ? 1 : 0;
    if (tables_in_smem) smem += table_bytes;
    if (smem > (size_t)plan->max_smem_optin)
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "expansion table of one point tile does not fit in shared memory");
    int rc = fb_set_smem(k_cellwise<SD, ORDER>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((npts + bp - 1) / bp);
    k_cellwise<SD, ORDER><<<grid, bp, smem, st>>>(P, plan->tab, E, pts, npts, ldp, out, ostride, tables_in_smem, M);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

template <int SD>
int dispatch_cellwise(const fiatb200_plan* plan, const DevSimplex& P, const DevEntity& E, const double* pts, long long npts,
                      long long ldp, double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    switch (P.order) {
        case 0: return launch_cellwise<SD, 0>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
        case 1: return launch_cellwise<SD, 1>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
        case 2: return launch_cellwise<SD, 2>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
        case 3: return launch_cellwise<SD, 3>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
        default: return launch_cellwise<SD, -1>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
    }
}

// ---- tile / DMMA launch ------------------------------------------------------------------------
bool mma_geometry(const fiatb200_plan* plan, const DevSimplex& P, int nrb, MmaGeom* G, size_t* smem_out) {
    if (P.ncells != 1 || P.expansion != 0 || P.order > 3 || P.nblk == 0 || nrb == 0) return false;
    const size_t budget = (size_t)plan->max_smem_optin - 1024;
    int pt_max = 128;
    const FbTuning& tune = fb_tuning();
    if (tune.mma_pt >= 0) pt_max = std::max(8, tune.mma_pt) & ~7;
    const int go = fb_mma_go(P.na);     // octets per contraction work item (kernels.cuh)
    int maxlev = 1;
    for (int l = 0; l < plan->tab.nlevels; ++l)
        maxlev = std::max(maxlev, (int)plan->tab.level_ptr[l + 1] - (int)plan->tab.level_ptr[l]);
    // first choice: a tile of >= 32 points small enough for two resident CTAs of 256 threads (one CTA's recurrence /
    // tail overlaps the other's contraction); otherwise the widest tile that fits one CTA of 512 threads per SM.
    for (int pass = 0; pass < 2; ++pass) {
        const size_t limit = pass == 0 ? (size_t)110 * 1024 : budget;
        const int pt_min = pass == 0 ? std::max(32, 8 * go) : 8 * go;
        if (pass == 0 && (tune.mma_pt >= 0 || tune.mma_threads >= 512)) continue;
        if (pass == 1 && tune.mma_threads >= 0 && tune.mma_threads < 512 && tune.mma_pt < 0) { /* forced 256: still allowed here */ }
        const int threads = pass == 0 ? 256 : (tune.mma_threads >= 0 && tune.mma_threads < 512 ? 256 : 512);
        for (int pt = pt_max; pt >= pt_min; pt >>= 1) {
            int ld = P.na * pt;
            while ((ld & 15) != 4 && (ld & 15) != 12) ++ld;
            const size_t table = (size_t)P.kpad * ld * sizeof(double);
            // warp-local recurrence (kernels.cuh) when the tile splits into 8 / 16 / 32 points per warp: its scratch
            // is the full list of step records; otherwise recurrence factors + two levels of records
            const int ppw_c = pt / (threads / 32);
            const bool warp_local = tune.mma_wl != 0 && pt % (threads / 32) == 0 && (ppw_c == 8 || ppw_c == 16 || ppw_c == 32);
            const size_t scratch = warp_local ? (size_t)plan->tab.nsteps * sizeof(StepRec)
                                              : (size_t)6 * pt * sizeof(double) + 2 * (size_t)maxlev * sizeof(StepRec);
            const size_t bytes = table + scratch;
            if (bytes <= limit) {
                G->PT = pt;
                G->logPT = 0;
                while ((1 << G->logPT) < pt) ++G->logPT;
                G->ldT = ld;
                G->maxlev = maxlev;
                G->ppw = warp_local ? ppw_c : 0;
                G->threads = threads;
                G->skip = 0;
                if (tune.mma_skip >= 0) G->skip = tune.mma_skip;    // profiling only
                *smem_out = bytes;
                return true;
            }
        }
    }
    return false;
}

template <int SD, int ORDER, int PW>
int launch_mma_pw(const DevSimplex& P, const RecTab& tab, const DevEntity& E, const MmaGeom& G, size_t smem, const double* pts,
                  long long npts, long long ldp, double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    int rc = fb_set_smem(k_mma<SD, ORDER, PW>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((npts + G.PT - 1) / G.PT);
    k_mma<SD, ORDER, PW><<<grid, G.threads, smem, st>>>(P, tab, E, G, pts, npts, ldp, out, ostride, M);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

template <int SD, int ORDER>
int launch_mma(const DevSimplex& P, const RecTab& tab, const DevEntity& E, const MmaGeom& G, size_t smem, const double* pts,
               long long npts, long long ldp, double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    if (G.PT >= 16) return launch_mma_pw<SD, ORDER, 16>(P, tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
    return launch_mma_pw<SD, ORDER, 8>(P, tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
}

template <int SD>
int dispatch_mma(const DevSimplex& P, const RecTab& tab, const DevEntity& E, const MmaGeom& G, size_t smem, const double* pts,
                 long long npts, long long ldp, double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    switch (P.order) {
        case 0: return launch_mma<SD, 0>(P, tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
        case 1: return launch_mma<SD, 1>(P, tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
        case 2: return launch_mma<SD, 2>(P, tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
        default: return launch_mma<SD, 3>(P, tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
    }
}

enum KernelChoice { K_CELLWISE = 1, K_MMA = 2, K_SMALL = 3, K_VALS = 4, K_LATTICE = 5, K_TENSOR = 6, K_MMA_CELLS = 7 };

// Which kernel a simplex plan runs on (flags: see include/fiat_b200.h).  Unless a flag forces one, the
// applicable kernels are ranked by a per-SM cycle estimate for 32 points:
//   HBM        32 points * 8 B * values/point at ~23.3 B per SM-cycle -- the floor of every kernel;
//   DMMA tile  16 cycles per DMMA.8x8x4 on 4 tensor pipes, 4 octets, measured ~75 % busy, plus the jets;
//   register / value-table kernels: one shared-memory wavefront per coefficient load (two when the lanes
//              of a warp sit in different subcells) against 0.5 cycles per FP64 instruction.
int choose_simplex_kernel(const fiatb200_plan* plan, uint32_t flags, MmaGeom* G, size_t* smem) {
    const DevSimplex& P = plan->simplex;
    bool use_mma = mma_geometry(plan, P, plan->tab.nrb, G, smem);
    if (flags & 1u) use_mma = false;
    if ((flags & 2u) && !use_mma) return 0;
    if (flags & 2u) return K_MMA;
    // derived order-0 elements of split-cell complexes (plan.macro_merged): points binned by subcell, DMMA
    if (!(flags & 1u) && fb_cells_applicable(plan)) return K_MMA_CELLS;
    const bool vals_ok = !(flags & 11u) && fb_vals_applicable(plan);
    const bool small_ok = !(flags & 3u) && fb_small_applicable(plan);
    const double nmem = P.nslots, rows = P.nrows, na = P.na, steps = plan->tab.nsteps;
    const double w = P.ncells > 1 ? 2.0 : 1.0;
    const double hbm = 32.0 * 8.0 * na * rows / 23.3;
    const double locate = P.ncells > 1 ? (P.ncells + 1.0) * (P.sd + 1) * (2 * P.sd + 3) : 0.0;
    double best = 1e300;
    int pick = K_CELLWISE;
    // (1-D sets with a handful of rows stay per-thread; in 2-D / 3-D the thread-per-point fallback, which keeps
    // every member's jets in shared memory, is the slowest path for any element the tile kernel can take)
    if (use_mma && ((P.sd >= 2 && P.degree >= 1) || (long long)P.nrows * P.nslots >= 256)) {
        best = std::max(hbm, 16.0 * P.nblk * na / 0.75 + 0.5 * steps * 8.0 * na);
        pick = K_MMA;
    }
    if (small_ok) {
        const double c = std::max(hbm, std::max(w * rows * nmem, 0.5 * (rows * nmem * na + steps * 8.0 * na + locate)));
        if (c <= best) { best = c; pick = K_SMALL; }
    }
    if (vals_ok) {
        double c = 1e300;
        for (int j = 0; j <= (P.order >= 1 ? 1 : 0); ++j) {
            const double naj = fb_binom(P.sd + j, j);
            double loads = rows * nmem, fma = rows * nmem * naj;
            for (int k = j + 1; k <= P.order; ++k) {
                const double work = (double)fb_binom(P.sd + k - 1, k) * rows * (k <= P.degree ? fb_binom(P.degree - k + P.sd, P.sd) : 0);
                loads += work;
                fma += work;
            }
            c = std::min(c, std::max(w * loads, 0.5 * (fma + steps * (3.0 + 5.0 * P.sd * j) + locate)));
        }
        c = std::max(c, hbm);
        if (c <= best) { best = c; pick = K_VALS; }
    }
    return pick;
}

int tabulate_simplex(const fiatb200_plan* plan, const fiatb200_entity_map* entity, const double* pts, long long npts,
                     long long ldp, double* out, long long ostride, const DevRowMap& M, uint32_t flags, cudaStream_t st) {
    const DevSimplex& P = plan->simplex;
    const DevEntity E = make_entity(entity, P.sd);
    if (E.dim < 0 || E.dim > 3) return fb_fail(FIATB200_ERR_ARG, "entity dimension out of range");
    MmaGeom G;
    size_t smem = 0;
    switch (choose_simplex_kernel(plan, flags, &G, &smem)) {
        case 0: return fb_fail(FIATB200_ERR_UNSUPPORTED, "DMMA kernel not applicable to this plan");
        case K_VALS: return fb_dispatch_vals(plan, E, pts, npts, ldp, out, ostride, M, st);
        case K_MMA_CELLS:
            if (M.identity) return fb_dispatch_cells(plan, E, pts, npts, ldp, out, ostride, st);
            break;          // placed rows: thread-per-point
        case K_SMALL: return fb_dispatch_small(plan, E, pts, npts, ldp, out, ostride, M, st);
        case K_MMA:
            switch (P.sd) {
                case 1: return dispatch_mma<1>(P, plan->tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
                case 2: return dispatch_mma<2>(P, plan->tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
                default: return dispatch_mma<3>(P, plan->tab, E, G, smem, pts, npts, ldp, out, ostride, M, st);
            }
        default: break;
    }
    switch (P.sd) {
        case 1: return dispatch_cellwise<1>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
        case 2: return dispatch_cellwise<2>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
        default: return dispatch_cellwise<3>(plan, P, E, pts, npts, ldp, out, ostride, M, st);
    }
}

int tabulate_tensor(const fiatb200_plan* plan, const double* pts, long long npts, long long ldp, double* out,
                    long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevTensor& Q = plan->tensor;
    int bp = 64;            // measured: 64-thread blocks write the many-row tables ~4 % faster than 128
    if (fb_tuning().tensor_bp >= 0) bp = std::max(32, fb_tuning().tensor_bp) & ~31;
    while (bp > 32 && (size_t)Q.total_doubles * bp * sizeof(double) > 96 * 1024) bp >>= 1;
    const size_t smem = (size_t)Q.total_doubles * bp * sizeof(double);
    if (smem > (size_t)plan->max_smem_optin)
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "factor tables of one point tile do not fit in shared memory");
    const unsigned grid = (unsigned)((npts + bp - 1) / bp);
    int rc = FIATB200_OK;
#define FB_TENSOR_LAUNCH(O_)                                                          \
    rc = fb_set_smem(k_tensor<O_>, smem);                                                \
    if (rc) return rc;                                                                \
    k_tensor<O_><<<grid, bp, smem, st>>>(Q, pts, npts, ldp, out, ostride, M);
    switch (Q.order) {
        case 0: FB_TENSOR_LAUNCH(0) break;
        case 1: FB_TENSOR_LAUNCH(1) break;
        case 2: FB_TENSOR_LAUNCH(2) break;
        case 3: FB_TENSOR_LAUNCH(3) break;
        default: FB_TENSOR_LAUNCH(-1) break;
    }
#undef FB_TENSOR_LAUNCH
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

template <int SD, int ORDER>
int launch_lattice(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                   double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevLattice& L = plan->lattice;
    const size_t per_point = (size_t)(SD + 1) * (L.degree + 1) * (ORDER + 1) * sizeof(double);
    int bp = 128;
    while (bp > 32 && per_point * bp > 56 * 1024) bp >>= 1;          // keep >= 4 CTAs per SM resident
    const size_t smem = per_point * bp;
    if (smem > (size_t)plan->max_smem_optin)
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "lattice factor tables do not fit in shared memory");
    int rc = fb_set_smem(k_lattice<SD, ORDER>, smem);
    if (rc) return rc;
    const unsigned grid = (unsigned)((npts + bp - 1) / bp);
    k_lattice<SD, ORDER><<<grid, bp, smem, st>>>(L, E, pts, npts, ldp, out, ostride, M);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

int tabulate_lattice(const fiatb200_plan* plan, const fiatb200_entity_map* entity, const double* pts, long long npts,
                     long long ldp, double* out, long long ostride, const DevRowMap& M, cudaStream_t st) {
    const DevLattice& L = plan->lattice;
    const DevEntity E = make_entity(entity, L.sd);
    if (L.sd == 2) {
        switch (L.order) {
            case 0: return launch_lattice<2, 0>(plan, E, pts, npts, ldp, out, ostride, M, st);
            case 1: return launch_lattice<2, 1>(plan, E, pts, npts, ldp, out, ostride, M, st);
            default: return launch_lattice<2, 2>(plan, E, pts, npts, ldp, out, ostride, M, st);
        }
    }
    switch (L.order) {
        case 0: return launch_lattice<3, 0>(plan, E, pts, npts, ldp, out, ostride, M, st);
        case 1: return launch_lattice<3, 1>(plan, E, pts, npts, ldp, out, ostride, M, st);
        default: return launch_lattice<3, 2>(plan, E, pts, npts, ldp, out, ostride, M, st);
    }
}

fiatb200_plan* new_plan() {
    fiatb200_plan* plan = new fiatb200_plan();
    plan->kind = 0; plan->device = 0; plan->blob = nullptr;
    memset(&plan->simplex, 0, sizeof(plan->simplex));
    memset(&plan->tab, 0, sizeof(plan->tab));
    memset(&plan->small_tab, 0, sizeof(plan->small_tab));
    memset(&plan->tensor, 0, sizeof(plan->tensor));
    memset(&plan->lattice, 0, sizeof(plan->lattice));
    plan->max_smem_optin = plan->num_sms = 0;
    plan->max_segment = 0;
    for (int i = 0; i < 2; ++i) { plan->host_stream[i] = nullptr; plan->host_pts[i] = nullptr; plan->host_out[i] = nullptr; }
    plan->host_pts_cap = plan->host_out_cap = 0;
    return plan;
}

void free_staging(fiatb200_plan* plan) {
    for (int i = 0; i < 2; ++i) {
        if (plan->host_pts[i]) cudaFree(plan->host_pts[i]);
        if (plan->host_out[i]) cudaFree(plan->host_out[i]);
        if (plan->host_stream[i]) cudaStreamDestroy(plan->host_stream[i]);
        plan->host_pts[i] = plan->host_out[i] = nullptr;
        plan->host_stream[i] = nullptr;
    }
    plan->host_pts_cap = plan->host_out_cap = 0;
}

int device_limits(fiatb200_plan* plan) {
    FB_CUDA(cudaGetDevice(&plan->device));
    FB_CUDA(cudaDeviceGetAttribute(&plan->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, plan->device));
    FB_CUDA(cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, plan->device));
    return FIATB200_OK;
}

}  // namespace

extern "C" {

int fiatb200_version(void) { return 1; }

const char* fiatb200_last_error(void) { return g_error.c_str(); }

int64_t fiatb200_launch_count(void) { return fb_launches.load(); }

int fiatb200_simplex_plan_create(const fiatb200_simplex_program* h, fiatb200_plan** out) {
    if (!h || !out) return fb_fail(FIATB200_ERR_ARG, "null argument");
    if (h->sd < 1 || h->sd > 3) return fb_fail(FIATB200_ERR_UNSUPPORTED, "spatial dimension must be 1, 2 or 3");
    if (h->ncells < 1 || h->ncells > 32) return fb_fail(FIATB200_ERR_UNSUPPORTED, "at most 32 subcells are supported");
    if (h->na < 1 || h->na > FB_NA_MAX) return fb_fail(FIATB200_ERR_UNSUPPORTED, "derivative order too high");
    if (h->expansion != 0 && h->sd != 1) return fb_fail(FIATB200_ERR_ARG, "line expansion on a non-line cell");
    fiatb200_plan* plan = new_plan();
    plan->kind = PLAN_SIMPLEX;
    int rc = device_limits(plan);
    if (rc) { delete plan; return rc; }

    if (h->nsteps > FB_MAX_STEPS || h->nlevels > FB_MAX_LEVELS || h->nfix > FB_MAX_FIX || h->nfixgrp > FB_MAX_FIX) {
        delete plan;
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "expansion degree too high for the device recurrence tables");
    }
    RecTab& R = plan->tab;
    R.nsteps = h->nsteps; R.nlevels = h->nlevels; R.nfix = h->nfix; R.nfixgrp = h->nfixgrp;
    R.start_slot = h->start_slot;
    if (h->start_slot < 0 || h->start_slot >= h->nslots) { delete plan; return fb_fail(FIATB200_ERR_ARG, "start slot out of range"); }
    for (int i = 0; i <= h->nlevels; ++i) R.level_ptr[i] = (short)h->level_ptr[i];
    for (int i = 0; i < h->nsteps; ++i) {
        StepRec& r = R.steps[i];
        r.nxt = (short)h->step_idx[4 * i + 0]; r.cur = (short)h->step_idx[4 * i + 1];
        r.prv = (short)h->step_idx[4 * i + 2]; r.codim = (short)h->step_idx[4 * i + 3];
        r.a = h->step_abc[3 * i + 0]; r.b = h->step_abc[3 * i + 1]; r.c = h->step_abc[3 * i + 2];
    }
    for (int i = 0; i < h->nfix; ++i) { R.fix_src[i] = (short)h->fix_idx[2 * i + 1]; R.fix_w[i] = h->fix_w[i]; }
    for (int g = 0; g < h->nfixgrp; ++g) {
        R.fix_first[g] = (short)h->fix_grp[2 * g]; R.fix_cnt[g] = (short)h->fix_grp[2 * g + 1];
        R.fix_tgt[g] = (short)h->fix_idx[2 * h->fix_grp[2 * g]];
    }
    for (int i = 0; i < FB_GEOM_DOUBLES; ++i) R.geom0[i] = h->geom[i];
    R.nrb = 0;
    if (h->nrb <= FB_MAX_RB) {
        R.nrb = h->nrb;
        for (int i = 0; i < h->nrb; ++i) R.rb_order[i] = (short)h->rb_order[i];
        for (int i = 0; i <= h->nrb; ++i) R.blk_ptr[i] = h->blk_ptr[i];
        for (int i = 0; i < h->nrb * 8; ++i) R.row_perm[i] = (short)(i < h->nrows ? h->row_perm[i] : -1);
    }
    memset(&plan->small_tab, 0, sizeof(plan->small_tab));
    for (int i = 0; i < h->nsteps && i < FB_SMALL_MAX_STEPS; ++i)
        for (int j = 0; j < 3; ++j) plan->small_tab.abc[i][j] = h->nat_abc[3 * i + j];
    for (int i = 0; i < 16 * (h->ncells + 1); ++i) plan->small_tab.bary[i] = h->bary[i];

    Arena A;
    const size_t o_tab = A.add(&R, sizeof(RecTab));
    const size_t o_geom = A.add(h->geom, sizeof(double) * FB_GEOM_DOUBLES * h->ncells);
    const size_t o_bary = A.add(h->bary, sizeof(double) * 16 * (h->ncells + 1));
    const size_t o_ccell = A.add(h->ccell, sizeof(double) * (size_t)h->ncells * h->nrows * h->nslots);
    const size_t o_ccellm = A.add(h->ccell_morton, sizeof(double) * (size_t)h->ncells * h->nrows * h->nslots);
    const size_t o_low1 = A.add(h->low1, sizeof(int32_t) * 3 * h->na);
    const size_t o_mul1 = A.add(h->mul1, sizeof(double) * 3 * h->na);
    const size_t o_low2 = A.add(h->low2, sizeof(int32_t) * 6 * h->na);
    const size_t o_mul2 = A.add(h->mul2, sizeof(double) * 6 * h->na);
    const size_t o_line = A.add(h->line_tab, sizeof(double) * h->line_tab_len);
    const int blk_cells = h->blk_cells > 1 ? h->blk_cells : 1;
    if (blk_cells > 1 && blk_cells != h->ncells) { delete plan; return fb_fail(FIATB200_ERR_ARG, "per-subcell block tables do not match the complex"); }
    const size_t o_blk_ptr = A.add(h->blk_ptr, sizeof(int32_t) * (size_t)blk_cells * (h->nrb + 1));
    const size_t o_blk_kb = A.add(h->blk_kb, sizeof(int32_t) * 4 * (size_t)h->nblk);      // four member slots per block
    const size_t o_blk_frag = A.add(h->blk_frag, sizeof(double) * 32 * (size_t)h->nblk);
    const size_t o_rb_order = A.add(h->rb_order, sizeof(int32_t) * h->nrb);
    const bool has_cderiv = h->ncp > 0 && h->cderiv && h->cderiv_len > 0;
    const size_t o_cderiv = A.add(h->cderiv, has_cderiv ? sizeof(double) * (size_t)h->cderiv_len : 0);
    const bool has_cstream = h->cstream && h->cstream_len > 0 && h->cstep_ptr && h->cnsteps > 0 && h->crb > 0;
    if (has_cstream && (h->cstep_ptr[h->cnsteps] != h->cstream_len || h->cnsteps * h->crb * 8 < h->nrows)) {
        delete plan;
        return fb_fail(FIATB200_ERR_ARG, "fixed-k block stream does not match its step table");
    }
    const size_t o_cstream = A.add(h->cstream, has_cstream ? sizeof(double) * (size_t)h->cstream_len : 0);
    const size_t o_cstep = A.add(h->cstep_ptr, has_cstream ? sizeof(int32_t) * (size_t)(h->cnsteps + 1) : 0);

    cudaError_t e = cudaMalloc(&plan->blob, A.host.size() + 256);
    if (e == cudaSuccess) e = cudaMemcpy(plan->blob, A.host.data(), A.host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (plan->blob) cudaFree(plan->blob);
        delete plan;
        return fb_fail(FIATB200_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
    }
    DevSimplex& P = plan->simplex;
    P.sd = h->sd; P.degree = h->degree; P.order = h->order; P.na = h->na; P.expansion = h->expansion;
    P.ncells = h->ncells; P.nslots = h->nslots; P.nrows = h->nrows; P.unique = h->unique;
    P.line_n = h->line_n;
    P.ncomp = h->ncomp > 0 ? h->ncomp : 1;
    void* b = plan->blob;
    P.tab = at<RecTab>(b, o_tab);
    P.geom = at<double>(b, o_geom);
    P.bary = at<double>(b, o_bary);
    P.ccell = at<double>(b, o_ccell);
    P.ccell_morton = at<double>(b, o_ccellm);
    P.low1 = at<int>(b, o_low1);
    P.mul1 = at<double>(b, o_mul1);
    P.low2 = at<int>(b, o_low2);
    P.mul2 = at<double>(b, o_mul2);
    P.line_tab = at<double>(b, o_line);
    P.nrb = h->nrb; P.kpad = h->kpad; P.nblk = h->nblk;
    P.blk_ptr = at<int>(b, o_blk_ptr);
    P.blk_kb = at<int>(b, o_blk_kb);
    P.blk_frag = at<double>(b, o_blk_frag);
    P.rb_order = at<int>(b, o_rb_order);
    P.cderiv = at<double>(b, o_cderiv);
    P.cderiv_len = has_cderiv ? (int)h->cderiv_len : 0;
    P.ncp = has_cderiv ? h->ncp : 0;
    P.blk_cells = blk_cells > 1 ? blk_cells : 0;
    P.cstream = at<double>(b, o_cstream);
    P.cstep_ptr = at<int>(b, o_cstep);
    P.cnsteps = has_cstream ? h->cnsteps : 0;
    P.crb = has_cstream ? h->crb : 0;
    P.cmaxstep = 0;
    for (int i = 0; i < P.cnsteps; ++i) P.cmaxstep = std::max(P.cmaxstep, h->cstep_ptr[i + 1] - h->cstep_ptr[i]);
    plan->max_segment = 0;
    for (int i = 0; i < blk_cells * (h->nrb + 1) - 1; ++i)
        if ((i + 1) % (h->nrb + 1) != 0) plan->max_segment = std::max(plan->max_segment, h->blk_ptr[i + 1] - h->blk_ptr[i]);
    *out = plan;
    return FIATB200_OK;
}

int fiatb200_tensor_plan_create(const fiatb200_tensor_leaf* leaves, int32_t nleaf, int32_t order,
                                fiatb200_plan** out) {
    if (!leaves || !out) return fb_fail(FIATB200_ERR_ARG, "null argument");
    if (nleaf < 1 || nleaf > FB_MAX_LEAVES)
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "tensor-product elements with 1..4 scalar factors are supported");
    fiatb200_plan* plan = new_plan();
    plan->kind = PLAN_TENSOR;
    int rc = device_limits(plan);
    if (rc) { delete plan; return rc; }
    DevTensor& Q = plan->tensor;
    Q.nleaf = nleaf;
    Q.order = order;
    Q.nrows = 1;
    int scratch = 0, off = 0, sd_total = 0;
    for (int l = 0; l < nleaf; ++l) {
        const fiatb200_plan* lp = leaves[l].plan;
        if (!lp || lp->kind != PLAN_SIMPLEX || lp->simplex.order != order) {
            delete plan;
            return fb_fail(FIATB200_ERR_ARG, "tensor leaves must be simplex plans of the same derivative order");
        }
        Q.leaf[l].prog = lp->simplex;
        Q.leaf[l].ent = make_entity(&leaves[l].entity, lp->simplex.sd);
        Q.leaf[l].point_offset = leaves[l].point_offset;
        scratch = std::max(scratch, lp->simplex.nslots * lp->simplex.na);
        Q.nrows *= lp->simplex.nrows;
        sd_total += lp->simplex.sd;
    }
    off = scratch;
    Q.ncomp = 1;
    int nvector = 0;
    for (int l = 0; l < nleaf; ++l) {
        Q.leaf[l].table_off = off;
        off += Q.leaf[l].prog.nrows * Q.leaf[l].prog.na;
        Q.leaf[l].ncomp = Q.leaf[l].prog.ncomp;
        Q.leaf[l].ndof = Q.leaf[l].prog.nrows / Q.leaf[l].prog.ncomp;
        if (Q.leaf[l].ncomp > 1) { Q.ncomp = Q.leaf[l].ncomp; ++nvector; }
    }
    if (nvector > 1) {
        delete plan;
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "at most one vector-valued tensor-product factor (tensor_product.py:271-272)");
    }
    for (int l = nleaf - 1, stride = 1; l >= 0; --l) {
        Q.leaf[l].dof_stride = stride;
        stride *= Q.leaf[l].ndof;
    }
    Q.scratch_doubles = scratch;
    Q.total_doubles = off;

    // product multi-indices in mis order, split into per-leaf alpha indices
    std::vector<std::vector<int>> alphas;
    for (int k = 0; k <= order; ++k) {
        // enumerate m-tuples summing to k, first entry descending (mis)
        std::vector<int> cur(sd_total, 0);
        struct Rec {
            static void go(std::vector<std::vector<int>>& outv, std::vector<int>& cur, int pos, int left) {
                const int m = (int)cur.size();
                if (pos == m - 1) { cur[pos] = left; outv.push_back(cur); return; }
                for (int v = left; v >= 0; --v) { cur[pos] = v; go(outv, cur, pos + 1, left - v); }
            }
        };
        Rec::go(alphas, cur, 0, k);
    }
    Q.nalpha = (int)alphas.size();
    auto leaf_alpha_index = [&](const int* a, int sd) {
        // position of the sd-tuple a within mis(sd,0), mis(sd,1), ..., in mis order
        int tot = 0;
        for (int i = 0; i < sd; ++i) tot += a[i];
        int idx = 0;
        for (int k = 0; k < tot; ++k) idx += fb_binom(sd + k - 1, k);
        // rank within mis(sd, tot): first entry descending
        int left = tot;
        for (int i = 0; i < sd - 1; ++i) {
            for (int v = left; v > a[i]; --v) idx += fb_binom((sd - i - 1) + (left - v) - 1, left - v);
            left -= a[i];
        }
        return idx;
    };
    std::vector<int> table((size_t)Q.nalpha * FB_MAX_LEAVES, 0);
    for (int j = 0; j < Q.nalpha; ++j) {
        int pos = 0;
        for (int l = 0; l < nleaf; ++l) {
            const int sd = Q.leaf[l].prog.sd;
            table[(size_t)j * FB_MAX_LEAVES + l] = leaf_alpha_index(alphas[j].data() + pos, sd);
            pos += sd;
        }
    }
    cudaError_t e = cudaMalloc(&plan->blob, table.size() * sizeof(int) + 256);
    if (e == cudaSuccess) e = cudaMemcpy(plan->blob, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (plan->blob) cudaFree(plan->blob);
        delete plan;
        return fb_fail(FIATB200_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
    }
    Q.alpha_leaf = static_cast<const int*>(plan->blob);
    *out = plan;
    return FIATB200_OK;
}

int fiatb200_lattice_plan_create(int32_t sd, int32_t degree, int32_t order, const int32_t* rowmap, int32_t ndofs,
                                 fiatb200_plan** out) {
    if (!rowmap || !out) return fb_fail(FIATB200_ERR_ARG, "null argument");
    if ((sd != 2 && sd != 3) || order < 0 || order > 2 || degree < 1)
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "lattice plans cover sd 2..3, order <= 2, degree >= 1");
    if (ndofs != fb_binom(degree + sd, sd)) return fb_fail(FIATB200_ERR_ARG, "ndofs does not match the lattice");
    fiatb200_plan* plan = new_plan();
    plan->kind = PLAN_LATTICE;
    int rc = device_limits(plan);
    if (rc) { delete plan; return rc; }
    std::vector<double> recip(degree);
    for (int k = 0; k < degree; ++k) recip[k] = 1.0 / (double)(k + 1);
    Arena A;
    const size_t o_map = A.add(rowmap, sizeof(int32_t) * ndofs);
    const size_t o_rec = A.add(recip.data(), sizeof(double) * degree);
    cudaError_t e = cudaMalloc(&plan->blob, A.host.size() + 256);
    if (e == cudaSuccess) e = cudaMemcpy(plan->blob, A.host.data(), A.host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (plan->blob) cudaFree(plan->blob);
        delete plan;
        return fb_fail(FIATB200_ERR_CUDA, std::string("plan upload: ") + cudaGetErrorString(e));
    }
    DevLattice& L = plan->lattice;
    L.sd = sd; L.degree = degree; L.order = order; L.na = fb_binom(sd + order, order); L.ndofs = ndofs;
    L.rowmap = at<int>(plan->blob, o_map);
    L.recip = at<double>(plan->blob, o_rec);
    *out = plan;
    return FIATB200_OK;
}

int fiatb200_plan_destroy(fiatb200_plan* plan) {
    if (!plan) return FIATB200_OK;
    free_staging(plan);
    if (plan->blob) cudaFree(plan->blob);
    delete plan;
    return FIATB200_OK;
}

int fiatb200_plan_kernel(const fiatb200_plan* plan, uint32_t flags) {
    if (!plan) return 0;
    if (plan->kind == PLAN_LATTICE) return K_LATTICE;
    if (plan->kind == PLAN_TENSOR) return K_TENSOR;
    MmaGeom G;
    size_t smem = 0;
    return choose_simplex_kernel(plan, flags, &G, &smem);
}

int fiatb200_plan_shape(const fiatb200_plan* plan, int64_t* nrows, int64_t* nalpha) {
    if (!plan) return fb_fail(FIATB200_ERR_ARG, "null plan");
    if (plan->kind == PLAN_SIMPLEX) {
        if (nrows) *nrows = plan->simplex.nrows;
        if (nalpha) *nalpha = plan->simplex.na;
    } else if (plan->kind == PLAN_LATTICE) {
        if (nrows) *nrows = plan->lattice.ndofs;
        if (nalpha) *nalpha = plan->lattice.na;
    } else {
        if (nrows) *nrows = plan->tensor.nrows;
        if (nalpha) *nalpha = plan->tensor.nalpha;
    }
    return FIATB200_OK;
}

int fiatb200_tabulate_mapped(const fiatb200_plan* plan, const fiatb200_entity_map* entity, const double* pts_dev,
                             int64_t npts, int64_t pts_ld, double* out_dev, int64_t out_row_stride,
                             const fiatb200_row_map* map, uint32_t flags, void* stream) {
    if (!plan) return fb_fail(FIATB200_ERR_ARG, "null plan");
    if (npts < 0 || out_row_stride < npts) return fb_fail(FIATB200_ERR_ARG, "bad point count / row stride");
    if (npts == 0) return FIATB200_OK;
    if (!out_dev || (!pts_dev && pts_ld != 0)) return fb_fail(FIATB200_ERR_ARG, "null device pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int ncomp = 1;
    int64_t nrows = 0;
    if (plan->kind == PLAN_SIMPLEX) { ncomp = plan->simplex.ncomp; nrows = plan->simplex.nrows; }
    else if (plan->kind == PLAN_LATTICE) { nrows = plan->lattice.ndofs; }
    else { ncomp = plan->tensor.ncomp; nrows = plan->tensor.nrows; }
    if (map) {
        if (map->nc_in != ncomp || map->nc_in < 1 || map->nc_in > 9 || map->nc_out < 1)
            return fb_fail(FIATB200_ERR_ARG, "row map does not match the plan's components");
        for (int k = 0; k < map->nc_in; ++k)
            if (map->comp_out[k] < 0 || map->comp_out[k] >= map->nc_out)
                return fb_fail(FIATB200_ERR_ARG, "row map component out of range");
        if ((int64_t)(map->dof_base + nrows / ncomp) * map->nc_out > map->total_rows)
            return fb_fail(FIATB200_ERR_ARG, "row map exceeds the output table");
    }
    const DevRowMap M = make_row_map(map, ncomp, (int)nrows);
    if (plan->kind == PLAN_SIMPLEX)
        return tabulate_simplex(plan, entity, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, flags, st);
    if (plan->kind == PLAN_LATTICE)
        return tabulate_lattice(plan, entity, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st);
    return tabulate_tensor(plan, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st);
}

int fiatb200_tabulate(const fiatb200_plan* plan, const fiatb200_entity_map* entity, const double* pts_dev,
                      int64_t npts, int64_t pts_ld, double* out_dev, int64_t out_row_stride, uint32_t flags,
                      void* stream) {
    return fiatb200_tabulate_mapped(plan, entity, pts_dev, npts, pts_ld, out_dev, out_row_stride, nullptr, flags, stream);
}

int fiatb200_evaluate_tensor(const fiatb200_plan* plan, const double* coef_dev, int32_t nfunc, const double* pts_dev,
                             int64_t npts, int64_t pts_ld, double* out_dev, int64_t out_row_stride, void* stream) {
    if (!plan || plan->kind != PLAN_TENSOR) return fb_fail(FIATB200_ERR_ARG, "a tensor-product plan is required");
    if (plan->tensor.ncomp != 1) return fb_fail(FIATB200_ERR_UNSUPPORTED, "fused evaluation covers scalar tensor-product elements");
    if (nfunc < 1 || npts < 0 || out_row_stride < npts) return fb_fail(FIATB200_ERR_ARG, "bad function / point count / row stride");
    if (npts == 0) return FIATB200_OK;
    if (!coef_dev || !out_dev || (!pts_dev && pts_ld != 0)) return fb_fail(FIATB200_ERR_ARG, "null device pointer");
    const DevTensor& Q = plan->tensor;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (measured: 128-point blocks 8.7 ms, 64-point blocks 9.2 ms per 2^22 points on the GLL Q10 hexahedron)
    int bp = 128;
    if (fb_tuning().eval_bp >= 0) bp = std::max(32, fb_tuning().eval_bp) & ~31;
    while (bp > 32 && (size_t)Q.total_doubles * bp * sizeof(double) > 96 * 1024) bp >>= 1;
    const size_t smem = (size_t)Q.total_doubles * bp * sizeof(double);
    if (smem > (size_t)plan->max_smem_optin)
        return fb_fail(FIATB200_ERR_UNSUPPORTED, "factor tables of one point tile do not fit in shared memory");
    const unsigned grid = (unsigned)((npts + bp - 1) / bp);
    int rc = FIATB200_OK;
    // three scalar line factors of moderate size (quadrilateral x interval = hexahedron): sum-factorised kernel
    bool hex = Q.nleaf == 3 && Q.order <= 2 && fb_tuning().eval_generic < 0;
    for (int l = 0; l < Q.nleaf && hex; ++l)
        hex = Q.leaf[l].prog.sd == 1 && Q.leaf[l].ncomp == 1 && Q.leaf[l].ndof <= FB_EVAL_NMAX;
    if (hex) {
#define FB_HEX_LAUNCH(O_)                                                             \
        rc = fb_set_smem(k_hex_eval<O_>, smem);                                       \
        if (rc) return rc;                                                            \
        k_hex_eval<O_><<<grid, bp, smem, st>>>(Q, coef_dev, nfunc, pts_dev, npts, pts_ld, out_dev, out_row_stride);
        switch (Q.order) {
            case 0: FB_HEX_LAUNCH(0) break;
            case 1: FB_HEX_LAUNCH(1) break;
            default: FB_HEX_LAUNCH(2) break;
        }
#undef FB_HEX_LAUNCH
        fb_launches++;
        FB_CUDA(cudaGetLastError());
        return FIATB200_OK;
    }
#define FB_EVAL_LAUNCH(O_)                                                            \
    rc = fb_set_smem(k_tensor_eval<O_>, smem);                                        \
    if (rc) return rc;                                                                \
    k_tensor_eval<O_><<<grid, bp, smem, st>>>(Q, coef_dev, nfunc, pts_dev, npts, pts_ld, out_dev, out_row_stride);
    switch (Q.order) {
        case 0: FB_EVAL_LAUNCH(0) break;
        case 1: FB_EVAL_LAUNCH(1) break;
        case 2: FB_EVAL_LAUNCH(2) break;
        case 3: FB_EVAL_LAUNCH(3) break;
        default: FB_EVAL_LAUNCH(-1) break;
    }
#undef FB_EVAL_LAUNCH
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

int fiatb200_evaluate_simplex(const fiatb200_plan* plan, int32_t nstack, int32_t ndofs, const double* coef_dev,
                              int32_t nfunc, const fiatb200_entity_map* entity, const double* pts_dev, int64_t npts,
                              int64_t pts_ld, double* out_dev, int64_t out_row_stride, void* stream) {
    if (!plan || plan->kind != PLAN_SIMPLEX) return fb_fail(FIATB200_ERR_ARG, "a simplex plan is required");
    const DevSimplex& P0 = plan->simplex;
    const int ncomp = P0.ncomp;
    if (P0.order != 0 || nstack < 1 || ndofs < 1 || nfunc < 1 || (int64_t)nstack * ndofs * ncomp != P0.nrows)
        return fb_fail(FIATB200_ERR_ARG, "evaluate needs the order-0 stacked derived plan: nrows == nstack * ndofs * ncomp");
    if (npts < 0 || out_row_stride < npts) return fb_fail(FIATB200_ERR_ARG, "bad point count / row stride");
    if (npts == 0) return FIATB200_OK;
    if (!coef_dev || !out_dev || (!pts_dev && pts_ld != 0)) return fb_fail(FIATB200_ERR_ARG, "null device pointer");
    const int64_t reval = (int64_t)nstack * nfunc * ncomp;
    if (reval > 8 * FB_MAX_RB) return fb_fail(FIATB200_ERR_UNSUPPORTED, "too many functions for one evaluation launch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const DevEntity E = make_entity(entity, P0.sd);
    if (E.dim < 0 || E.dim > 3) return fb_fail(FIATB200_ERR_ARG, "entity dimension out of range");

    // stream-ordered workspace: W, and for the tile kernel its fragments and member slots
    const int nrb = (int)((reval + 7) / 8), nkb = (P0.nslots + 3) / 4;
    const size_t w_bytes = sizeof(double) * (size_t)P0.ncells * reval * P0.nslots;
    const size_t f_bytes = sizeof(double) * 32 * (size_t)nrb * nkb, i_bytes = sizeof(int) * 4 * (size_t)nrb * nkb;
    const size_t w_off = 0, f_off = (w_bytes + 255) & ~size_t(255), i_off = (f_off + f_bytes + 255) & ~size_t(255);
    unsigned char* ws = nullptr;
    FB_CUDA(cudaMallocAsync(&ws, i_off + i_bytes + 256, st));
    double* W = reinterpret_cast<double*>(ws + w_off);
    double* frag = reinterpret_cast<double*>(ws + f_off);
    int* slots = reinterpret_cast<int*>(ws + i_off);
    {
        const long long total = (long long)P0.ncells * reval * P0.nslots;
        k_eval_weights<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(P0.ccell, P0.ncells, nstack, ndofs, ncomp, P0.nslots,
                                                                        coef_dev, nfunc, W);
        fb_launches++;
    }
    DevSimplex P = P0;
    P.nrows = (int)reval;
    P.ccell = W;
    P.ccell_morton = W;
    P.ncp = 0;
    P.cderiv_len = 0;
    P.blk_cells = 0;
    DevRowMap M = make_row_map(nullptr, ncomp, (int)reval);
    int rc = FIATB200_OK;
    MmaGeom G;
    size_t smem = 0;
    P.nrb = nrb; P.kpad = nkb * 4; P.nblk = nrb * nkb;
    P.blk_frag = frag; P.blk_kb = slots;
    if (P.sd >= 2 && P.degree >= 1 && mma_geometry(plan, P, nrb, &G, &smem)) {
        // single-cell Dubiner sets: level-parallel value recurrence + dense DMMA contraction with the weights
        k_eval_fragments<<<(unsigned)(nrb * nkb), 32, 0, st>>>(W, (int)reval, P.nslots, nkb, frag, slots);
        fb_launches++;
        RecTab tab = plan->tab;
        tab.nrb = nrb;
        for (int rb = 0; rb < nrb; ++rb) { tab.rb_order[rb] = (short)rb; tab.blk_ptr[rb] = rb * nkb; }
        tab.blk_ptr[nrb] = nrb * nkb;
        for (int i = 0; i < nrb * 8; ++i) tab.row_perm[i] = (short)(i < reval ? i : -1);
        switch (P.sd) {
            case 2: rc = dispatch_mma<2>(P, tab, E, G, smem, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st); break;
            default: rc = dispatch_mma<3>(P, tab, E, G, smem, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st); break;
        }
    } else {
        // split cells, 1-D sets: thread per point with the per-subcell weights
        switch (P.sd) {
            case 1: rc = dispatch_cellwise<1>(plan, P, E, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st); break;
            case 2: rc = dispatch_cellwise<2>(plan, P, E, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st); break;
            default: rc = dispatch_cellwise<3>(plan, P, E, pts_dev, npts, pts_ld, out_dev, out_row_stride, M, st); break;
        }
    }
    cudaError_t e = cudaFreeAsync(ws, st);
    if (rc == FIATB200_OK && e != cudaSuccess) rc = fb_fail(FIATB200_ERR_CUDA, cudaGetErrorString(e));
    return rc;
}

int fiatb200_zero_rows(double* out_dev, int64_t out_row_stride, int64_t npts, int64_t total_rows, int32_t nalpha,
                       const int32_t* rows_dev, int32_t nrows, void* stream) {
    if (npts == 0 || nrows == 0) return FIATB200_OK;
    if (!out_dev || !rows_dev || out_row_stride < npts) return fb_fail(FIATB200_ERR_ARG, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((npts + 127) / 128);
    k_zero_rows<<<grid, 128, 0, st>>>(out_dev, out_row_stride, npts, total_rows, nalpha, rows_dev, nrows);
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

int fiatb200_locate_subcells(const fiatb200_plan* plan, const fiatb200_entity_map* entity, const double* pts_dev,
                             int64_t npts, int64_t pts_ld, int32_t unique, uint32_t* mask_out_dev, void* stream) {
    if (!plan || plan->kind != PLAN_SIMPLEX) return fb_fail(FIATB200_ERR_ARG, "a simplex plan is required");
    if (npts == 0) return FIATB200_OK;
    if (!mask_out_dev || (!pts_dev && pts_ld != 0)) return fb_fail(FIATB200_ERR_ARG, "null device pointer");
    const DevSimplex& P = plan->simplex;
    const DevEntity E = make_entity(entity, P.sd);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)((npts + 127) / 128);
    switch (P.sd) {
        case 1: k_locate<1><<<grid, 128, 0, st>>>(P, E, pts_dev, npts, pts_ld, unique, mask_out_dev); break;
        case 2: k_locate<2><<<grid, 128, 0, st>>>(P, E, pts_dev, npts, pts_ld, unique, mask_out_dev); break;
        default: k_locate<3><<<grid, 128, 0, st>>>(P, E, pts_dev, npts, pts_ld, unique, mask_out_dev); break;
    }
    fb_launches++;
    FB_CUDA(cudaGetLastError());
    return FIATB200_OK;
}

}  // extern "C"

namespace {

// Host-buffer driver shared by fiatb200_tabulate_host and fiatb200_tabulate_host_list: chunks of points go
// H2D, `run` enqueues the kernels of one chunk, the chunk's rows go D2H into their final place; two streams
// with one staging buffer each (owned by `owner`, kept across calls) overlap copies and kernels.
template <typename Run>
int host_pipeline(fiatb200_plan* owner, int64_t rows, const double* pts_host, int64_t npts, int64_t pts_ld,
                  double* out_host, int64_t chunk_pts, Run run) {
    std::lock_guard<std::mutex> guard(owner->host_mutex);
    chunk_pts = std::min<int64_t>(chunk_pts, npts);
    chunk_pts = (chunk_pts + 7) & ~int64_t(7);
    const size_t need_pts = sizeof(double) * chunk_pts * std::max<int64_t>(pts_ld, 1);
    const size_t need_out = sizeof(double) * chunk_pts * rows;
    if (!owner->host_stream[0] || need_pts > owner->host_pts_cap || need_out > owner->host_out_cap) {
        free_staging(owner);
        for (int i = 0; i < 2; ++i) {
            FB_CUDA(cudaStreamCreateWithFlags(&owner->host_stream[i], cudaStreamNonBlocking));
            FB_CUDA(cudaMalloc(&owner->host_pts[i], need_pts));
            FB_CUDA(cudaMalloc(&owner->host_out[i], need_out));
        }
        owner->host_pts_cap = need_pts;
        owner->host_out_cap = need_out;
    }
    int rc = FIATB200_OK;
    int64_t done = 0;
    for (int it = 0; done < npts && rc == FIATB200_OK; ++it, done += chunk_pts) {
        const int b = it & 1;
        cudaStream_t st = owner->host_stream[b];
        const int64_t n = std::min<int64_t>(chunk_pts, npts - done);
        // (errors leave the loop but never the function: copies already queued into the caller's buffers must
        // have drained before we return)
        cudaError_t e = cudaSuccess;
        if (pts_ld > 0)
            e = cudaMemcpyAsync(owner->host_pts[b], pts_host + done * pts_ld, sizeof(double) * n * pts_ld,
                                cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            rc = run(owner->host_pts[b], n, owner->host_out[b], chunk_pts, st);
            if (rc) break;
            // rows of the chunk land at column offset `done` of the (rows x npts) host result
            e = cudaMemcpy2DAsync(out_host + done, sizeof(double) * npts, owner->host_out[b], sizeof(double) * chunk_pts,
                                  sizeof(double) * n, rows, cudaMemcpyDeviceToHost, st);
        }
        if (e != cudaSuccess) rc = fb_fail(FIATB200_ERR_CUDA, std::string("host pipeline copy: ") + cudaGetErrorString(e));
    }
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaStreamSynchronize(owner->host_stream[i]);
        if (e != cudaSuccess && rc == FIATB200_OK) rc = fb_fail(FIATB200_ERR_CUDA, cudaGetErrorString(e));
    }
    return rc;
}

}  // namespace

extern "C" {

int fiatb200_tabulate_host(const fiatb200_plan* cplan, const fiatb200_entity_map* entity, const double* pts_host,
                           int64_t npts, int64_t pts_ld, double* out_host, int64_t chunk_pts, uint32_t flags) {
    if (!cplan) return fb_fail(FIATB200_ERR_ARG, "null plan");
    if (npts == 0) return FIATB200_OK;
    if ((!pts_host && pts_ld != 0) || !out_host || chunk_pts <= 0)
        return fb_fail(FIATB200_ERR_ARG, "bad host buffers / chunk size");
    fiatb200_plan* plan = const_cast<fiatb200_plan*>(cplan);
    int64_t nrows = 0, nalpha = 0;
    fiatb200_plan_shape(plan, &nrows, &nalpha);
    return host_pipeline(plan, nrows * nalpha, pts_host, npts, pts_ld, out_host, chunk_pts,
                         [&](const double* dpts, int64_t n, double* dout, int64_t stride, cudaStream_t st) {
                             return fiatb200_tabulate(plan, entity, dpts, n, pts_ld, dout, stride, flags, st);
                         });
}

int fiatb200_tabulate_host_list(const fiatb200_launch* launches, int32_t nlaunch, int32_t nalpha, int64_t total_rows,
                                const int32_t* zero_rows_dev, int32_t nzero_rows, const double* pts_host, int64_t npts,
                                int64_t pts_ld, double* out_host, int64_t chunk_pts, uint32_t flags) {
    if (!launches || nlaunch < 1 || nalpha < 1 || total_rows < 1) return fb_fail(FIATB200_ERR_ARG, "empty launch list");
    if (npts == 0) return FIATB200_OK;
    if ((!pts_host && pts_ld != 0) || !out_host || chunk_pts <= 0)
        return fb_fail(FIATB200_ERR_ARG, "bad host buffers / chunk size");
    fiatb200_plan* owner = nullptr;
    for (int i = 0; i < nlaunch; ++i) {
        if (launches[i].alpha_offset < 0 || launches[i].alpha_offset >= nalpha)
            return fb_fail(FIATB200_ERR_ARG, "launch writes a derivative table that does not exist");
        if (launches[i].plan && !owner) owner = const_cast<fiatb200_plan*>(launches[i].plan);
    }
    if (!owner) return fb_fail(FIATB200_ERR_ARG, "launch list without a plan");
    return host_pipeline(owner, total_rows * nalpha, pts_host, npts, pts_ld, out_host, chunk_pts,
                         [&](const double* dpts, int64_t n, double* dout, int64_t stride, cudaStream_t st) {
                             int rc = FIATB200_OK;
                             if (zero_rows_dev && nzero_rows > 0)
                                 rc = fiatb200_zero_rows(dout, stride, n, total_rows, nalpha, zero_rows_dev, nzero_rows, st);
                             for (int i = 0; i < nlaunch && rc == FIATB200_OK; ++i) {
                                 const fiatb200_launch& L = launches[i];
                                 double* o = dout + (size_t)L.alpha_offset * total_rows * stride;
                                 if (!L.plan)
                                     rc = fiatb200_zero_rows(o, stride, n, total_rows, 1, L.zero_rows_dev, L.nzero_rows, st);
                                 else
                                     rc = fiatb200_tabulate_mapped(L.plan, L.entity, dpts, n, pts_ld, o, stride, L.map, flags, st);
                             }
                             return rc;
                         });
}

int fiatb200_evaluate_host(const fiatb200_plan* cplan, int32_t nstack, int32_t ndofs, const double* coef_host,
                           int32_t nfunc, const fiatb200_entity_map* entity, const double* pts_host, int64_t npts,
                           int64_t pts_ld, double* out_host, int64_t chunk_pts) {
    if (!cplan || (cplan->kind != PLAN_SIMPLEX && cplan->kind != PLAN_TENSOR))
        return fb_fail(FIATB200_ERR_ARG, "a simplex or tensor-product plan is required");
    if (npts == 0) return FIATB200_OK;
    if (!coef_host || (!pts_host && pts_ld != 0) || !out_host || chunk_pts <= 0 || nfunc < 1 || ndofs < 1 || nstack < 1)
        return fb_fail(FIATB200_ERR_ARG, "bad host buffers / sizes");
    fiatb200_plan* plan = const_cast<fiatb200_plan*>(cplan);
    const bool tensor = plan->kind == PLAN_TENSOR;
    const int ncomp = tensor ? 1 : plan->simplex.ncomp;
    if (tensor && (nstack != plan->tensor.nalpha || ndofs != plan->tensor.nrows))
        return fb_fail(FIATB200_ERR_ARG, "nstack / ndofs do not match the tensor-product plan");
    double* coef_dev = nullptr;
    FB_CUDA(cudaMalloc(&coef_dev, sizeof(double) * (size_t)nfunc * ndofs));
    cudaError_t e = cudaMemcpy(coef_dev, coef_host, sizeof(double) * (size_t)nfunc * ndofs, cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? FIATB200_OK : fb_fail(FIATB200_ERR_CUDA, cudaGetErrorString(e));
    if (rc == FIATB200_OK)
        rc = host_pipeline(plan, (int64_t)nstack * nfunc * ncomp, pts_host, npts, pts_ld, out_host, chunk_pts,
                           [&](const double* dpts, int64_t n, double* dout, int64_t stride, cudaStream_t st) {
                               if (tensor)
                                   return fiatb200_evaluate_tensor(plan, coef_dev, nfunc, dpts, n, pts_ld, dout, stride, st);
                               return fiatb200_evaluate_simplex(plan, nstack, ndofs, coef_dev, nfunc, entity, dpts, n, pts_ld,
                                                                dout, stride, st);
                           });
    cudaFree(coef_dev);
    return rc;
}

}  // extern "C"
