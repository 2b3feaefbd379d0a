// Split-cell tile kernel: FP64 tensor-pipe contraction for elements on split (macro) complexes.
//
// Input is the order-0 *derived* element of fiat_b200/plan.py: macro_merged: its rows are the original
// element's derivative tables one after the other, and for every subcell c there is one block-sparse
// coefficient matrix C_c on the un-normalised recurrence members of that subcell
// (FIAT/expansions.py:449-490 scatters per-cell tables into a zero-padded global table through
// cell_node_map; here the zeros are never formed).  A CTA owns a tile of PT points:
//   phase 0  locate every point's subcell (bit-exact binning, FIAT/expansions.py:771-811) and sort the
//            tile's columns by subcell, each subcell's range padded to a multiple of 8 columns;
//   phase 1  value recurrence (FIAT/expansions.py:202-249, no derivative jets) level by level for all
//            columns, each column in its own subcell's reference coordinates;
//   phase 2  out[:, columns of c] = C_c . T[:, columns of c] with mma.sync.m8n8k4.f64; a warp owns one 8-row
//            block for all columns of the tile, un-permutes it through shared memory and stores full rows;
//   phase 3  points on interior facets (several subcells, measure zero for random points) or in no subcell are finished
//            by their own thread: one value column per subcell, dense per-subcell rows from global
//            memory, tables divided by the multiplicity and accumulated (FIAT/expansions.py:467-489).
#pragma once
#include "expansion.cuh"

struct CellsGeom {
    int PT;        // points per tile
    int PTS;       // column capacity: PT + 8 * ncells (subcell ranges padded to octets), multiple of 8
    int ldT;       // doubles between member rows of T (>= PTS, = 4 or 12 mod 16)
    int maxlev;    // step records held in shared memory (all steps of the recurrence)
    int threads;   // 256: two CTAs per SM; 512: one CTA with the whole shared memory (wider tile)
};

#define FB_CELLS_THREADS 512   // upper bound; launched with 256 (two CTAs per SM) or 512 (one)
#define FB_CELLS_GO 8          // octets of one subcell per column chunk


// One (row block, column chunk) segment: the subcell's blocks of the row block against NOCT octets of columns, then
// the fragments go through the column permutation into the warp's staging rows.  The first CH coefficient
// fragments and their member slots (lane 4 j + t holds slot t of block j) were fetched while the previous segment
// ran (CH = 2, 4 or 8, chosen from the plan's longest segment); longer segments load the rest on demand.
template <int NOCT, int CH>
__device__ __forceinline__ void cells_segment(const double (&a)[CH], int kb, int t, const double* __restrict__ fp,
                                              const int* __restrict__ kp, int nq, const double* __restrict__ Tchunk,
                                              size_t ldT, const int* __restrict__ perm, double* __restrict__ srow) {
    double acc[NOCT][2];
#pragma unroll
    for (int o = 0; o < NOCT; ++o) acc[o][0] = acc[o][1] = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        if (j < nq) {
            const int slot = __shfl_sync(0xffffffffu, kb, 4 * j + t);
            const double* Tb = Tchunk + slot * ldT;
#pragma unroll
            for (int o = 0; o < NOCT; ++o) dmma_8x8x4(acc[o][0], acc[o][1], a[j], Tb[o * 8]);
        }
    }
#pragma unroll 1
    for (int i = CH; i < nq; ++i) {
        const double ai = __ldg(fp + (size_t)i * 32);
        const double* Tb = Tchunk + __ldg(kp + 4 * i) * ldT;
#pragma unroll
        for (int o = 0; o < NOCT; ++o) dmma_8x8x4(acc[o][0], acc[o][1], ai, Tb[o * 8]);
    }
#pragma unroll
    for (int o = 0; o < NOCT; ++o) {
        const int p0 = perm[o * 8], p1 = perm[o * 8 + 1];
        if (p0 >= 0) srow[p0] = acc[o][0];
        if (p1 >= 0) srow[p1] = acc[o][1];
    }
}

template <int SD, int CH>
__global__ void __launch_bounds__(FB_CELLS_THREADS, 1)
k_mma_cells(const DevSimplex P, const __grid_constant__ RecTab tab, const __grid_constant__ SmallTab st,
            const DevEntity E, const CellsGeom G, const double* __restrict__ pts, long long npts, long long ldp,
            double* __restrict__ out, long long ostride) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int PT = G.PT, PTS = G.PTS, ldT = G.ldT;
    double* T = smem;                                        // kpad x ldT
    double* s_fa = T + (size_t)P.kpad * ldT;                 // 3 x PTS
    double* s_fb = s_fa + 3 * PTS;                           // 3 x PTS
    StepRec* s_rec = reinterpret_cast<StepRec*>(s_fb + 3 * PTS);         // all step records
    int* s_perm = reinterpret_cast<int*>(s_rec + G.maxlev);              // PTS: column -> point of the tile (-1 padding)
    double* s_stage = reinterpret_cast<double*>(s_perm + PTS);           // warps x 8 rows x (PT + 2) (PTS is even)
    const int SP = PT + 2;      // staging row stride: rows of a fragment (same column, 8 rows) land in 8 different banks
    int* s_ptr = reinterpret_cast<int*>(s_stage + (size_t)(NT / 32) * 8 * SP);   // ncells x (nrb + 1) block offsets
    int* s_chunk = s_ptr + P.ncells * (tab.nrb + 1);                     // PTS / 8 entries: cell | oct0 << 8 | noct << 20
    __shared__ int s_cnt[32], s_off[33], s_fill[32], s_next, s_cols, s_nchunk;
    const long long base = (long long)blockIdx.x * PT;

    // ---- phase 0: subcell of every point, columns sorted by subcell -----------------------------
    if (tid < 32) { s_cnt[tid] = 0; s_fill[tid] = 0; }
    if (tid == 0) s_next = 0;
    for (int j = tid; j < PTS; j += NT) {
        s_perm[j] = -1;
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_fa[c * PTS + j] = 0.0; s_fb[c * PTS + j] = 0.0; }
    }
    for (int i = tid; i < P.kpad * ldT; i += NT) T[i] = 0.0;
    for (int i = tid; i < P.ncells * (tab.nrb + 1); i += NT) s_ptr[i] = __ldg(P.blk_ptr + i);
    __syncthreads();
    double x[3] = {0.0, 0.0, 0.0};
    unsigned mask = 0;
    const long long p = base + tid;
    const bool valid = tid < PT && p < npts;
    if (valid) {
        apply_entity<SD>(E, pts + p * ldp, x);
        mask = locate_cells<SD>(st.bary, P.ncells, P.unique, x);
    }
    const int mult = __popc(mask);
    const int cell = mult == 1 ? __ffs(mask) - 1 : -1;
    if (cell >= 0) atomicAdd(&s_cnt[cell], 1);
    __syncthreads();
    if (tid == 0) {
        int off = 0, nchunk = 0;
        for (int c = 0; c < P.ncells; ++c) {
            s_off[c] = off;
            const int noct = (s_cnt[c] + 7) >> 3;
            for (int o = 0; o < noct; o += FB_CELLS_GO)
                s_chunk[nchunk++] = c | ((off / 8 + o) << 8) | (min(FB_CELLS_GO, noct - o) << 20);
            off += noct * 8;
        }
        s_off[P.ncells] = off;
        s_cols = off;
        s_nchunk = nchunk;
    }
    __syncthreads();
    if (cell >= 0) {
        const int col = s_off[cell] + atomicAdd(&s_fill[cell], 1);
        s_perm[col] = tid;
        const double* geom = P.geom + cell * FB_GEOM_DOUBLES;
        double xr[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < SD; ++i) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < SD; ++d) s = fma(x[d], __ldg(geom + i * SD + d), s);
            xr[i] = s + __ldg(geom + 9 + i);
        }
        double fa[3], fb[3];
        recurrence_factors<SD>(xr, fa, fb);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s_fa[c * PTS + col] = fa[c];
            s_fb[c * PTS + col] = fb[c];
        }
        T[(size_t)tab.start_slot * ldT + col] = __ldg(geom + 12);
    }
    // all step records in shared memory (lanes of a warp work on different steps, which the constant cache would
    // serialise)
    for (int i = tid; i < tab.nsteps * 4; i += NT)
        reinterpret_cast<double*>(s_rec)[i] = reinterpret_cast<const double*>(tab.steps)[i];
    __syncthreads();
    const int ncols = s_cols;

    // ---- phase 1: value recurrence, warp-local -------------------------------------------------------
    // A warp owns blocks of 16 columns and runs the whole recurrence for them level by level with warp-level
    // synchronisation only (lane = (step slot, column)); the low wavefront levels have far fewer (step, column) items
    // than the CTA has threads, so CTA-wide levels would pay a block barrier and a latency-bound round each.
    for (int cb = tid >> 5; cb * 16 < ncols; cb += NT >> 5) {
        const int col = cb * 16 + (tid & 15);
        const bool active = col < ncols;
        const int cc = active ? col : 0;
        const double fa[3] = {s_fa[cc], s_fa[PTS + cc], s_fa[2 * PTS + cc]};
        const double fb[3] = {s_fb[cc], s_fb[PTS + cc], s_fb[2 * PTS + cc]};
        for (int lev = 0; lev < tab.nlevels; ++lev) {
            const int l0 = tab.level_ptr[lev], nst = tab.level_ptr[lev + 1] - l0;
            if (active)
                for (int sl = (tid & 31) >> 4; sl < nst; sl += 2)
                    run_step<SD, 0>(P, s_rec[l0 + sl], tab.geom0, fa, fb, T + col, ldT, 1, 1);
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- phase 2: block-sparse contraction on the FP64 tensor pipe -----------------------------------
    // Work item = one 8-row block, for ALL columns of the tile: the warp walks the tile's column chunks (<= GO octets
    // of one subcell each), streams that subcell's blocks of the row block through the DMMAs and scatters the
    // fragments through the column permutation into its staging buffer; the finished 8 x PT piece of the table then
    // leaves as full-width coalesced row stores.  (Storing 8-byte pieces straight through the permutation ran at
    // 0.10-0.17 of the HBM peak: every sector is then written four times, partially.)  The loops are kept rolled:
    // a fully unrolled variant was 23 000 SASS instructions and instruction-fetch bound.
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nrb = tab.nrb;
    const int nchunk = s_nchunk;
    const double* Tlane = T + g;                                     // + member slot * ldT, gathered per block
    double* stage = s_stage + (size_t)warp * 8 * SP;                 // 8 rows x PT points (+ 2 padding)
    const bool vec_ok = ((ostride & 1) == 0) && ((((size_t)out) & 15) == 0) && base + PT <= npts;
    // Segments are walked as one software-pipelined stream: while a segment's DMMAs run, the coefficient fragments and
    // member slots of the NEXT segment (next chunk of the row block, or first chunk of the warp's next row block) are
    // already in flight, so the L2 latency of a fragment is not paid once per 2-7 block segment.
    double a_cur[CH], a_nxt[CH];
    int kb_cur = 0, kb_nxt = 0;
    auto fetch_item = [&]() {
        int item = 0;
        if (lane == 0) item = atomicAdd(&s_next, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        return item < nrb ? (int)tab.rb_order[item] : -1;
    };
    int rb = nchunk > 0 ? fetch_item() : -1, ch = 0;
    if (rb >= 0) {
        const int c = s_chunk[0] & 255;
        const int q0 = s_ptr[c * (nrb + 1) + rb], q1 = s_ptr[c * (nrb + 1) + rb + 1];
#pragma unroll
        for (int j = 0; j < CH; ++j) a_cur[j] = (q0 + j < q1) ? __ldg(P.blk_frag + (size_t)(q0 + j) * 32 + lane) : 0.0;
        if (4 * q0 + lane < 4 * q1) kb_cur = __ldg(P.blk_kb + 4 * q0 + lane);
    }
    while (rb >= 0) {
        const int chunk = s_chunk[ch];
        const int c = chunk & 255, oct0 = (chunk >> 8) & 4095, noct = chunk >> 20;
        const int q0 = s_ptr[c * (nrb + 1) + rb], q1 = s_ptr[c * (nrb + 1) + rb + 1];
        // next segment and its first fragments
        const bool last = ch + 1 == nchunk;
        const int rb_n = last ? fetch_item() : rb, ch_n = last ? 0 : ch + 1;
        kb_nxt = 0;
        if (rb_n >= 0) {
            const int cn = s_chunk[ch_n] & 255;
            const int n0 = s_ptr[cn * (nrb + 1) + rb_n], n1 = s_ptr[cn * (nrb + 1) + rb_n + 1];
#pragma unroll
            for (int j = 0; j < CH; ++j) a_nxt[j] = (n0 + j < n1) ? __ldg(P.blk_frag + (size_t)(n0 + j) * 32 + lane) : 0.0;
            if (4 * n0 + lane < 4 * n1) kb_nxt = __ldg(P.blk_kb + 4 * n0 + lane);
        }
        const double* fp = P.blk_frag + (size_t)q0 * 32 + lane;
        const int* kp = P.blk_kb + 4 * q0 + t;
        const double* Tchunk = Tlane + oct0 * 8;
        const int* perm = s_perm + oct0 * 8 + 2 * t;
        double* srow = stage + g * SP;
        // one specialisation per number of octets: exactly noct DMMAs per block (an `if (o < noct)` in an unrolled
        // loop becomes predicated DMMAs that still occupy the tensor pipe) and no dispatch inside the block loop
        switch (noct) {
            case 1: cells_segment<1, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            case 2: cells_segment<2, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            case 3: cells_segment<3, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            case 4: cells_segment<4, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            case 5: cells_segment<5, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            case 6: cells_segment<6, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            case 7: cells_segment<7, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
            default: cells_segment<8, CH>(a_cur, kb_cur, t, fp, kp, q1 - q0, Tchunk, (size_t)ldT, perm, srow); break;
        }
        if (last) {
            __syncwarp();
#pragma unroll 1
            for (int r = 0; r < 8; ++r) {
                const int row = tab.row_perm[rb * 8 + r];
                if (row < 0) continue;
                double* rowp = out + (size_t)row * ostride + base;
                if (vec_ok) {
                    for (int i = lane * 2; i < PT; i += 64)
                        *reinterpret_cast<double2*>(rowp + i) = *reinterpret_cast<const double2*>(stage + r * SP + i);
                } else {
                    for (int i = lane; i < PT; i += 32)
                        if (base + i < npts) rowp[i] = stage[r * SP + i];
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) a_cur[j] = a_nxt[j];
        kb_cur = kb_nxt;
        rb = rb_n;
        ch = ch_n;
    }

    // ---- phase 3: points shared by several subcells -------------------------------------------------
    if (!__syncthreads_or(valid && mult != 1)) return;
    if (valid && mult == 0) {       // in no subcell (NaN / Inf coordinates): zero column, like the reference
        for (int r = 0; r < P.nrows; ++r) out[(size_t)r * ostride + p] = 0.0;
    }
    if (valid && mult > 1) {
        const double inv_mult = 1.0 / (double)mult;
        double* Tcol = T + tid;                             // one private column per thread (tid < PT <= ldT)
        bool first = true;
        while (mask) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1;
            expansion_point<SD, 0>(P, tab, P.geom + c * FB_GEOM_DOUBLES, c, inv_mult, x, Tcol, ldT, 1, 1);
            const double* C = P.ccell + (size_t)c * P.nrows * P.nslots;
            for (int r = 0; r < P.nrows; ++r) {
                double s = 0.0;
                for (int k = 0; k < P.nslots; ++k) s = fma(__ldg(C + (size_t)r * P.nslots + k), Tcol[(size_t)k * ldT], s);
                double* o = out + (size_t)r * ostride + p;
                *o = first ? s : (*o + s);
            }
            first = false;
        }
    }
}
