// Tabulation kernels (sm_100a).
//
//   k_cellwise   thread per point: locate subcell(s) -> expansion in a private shared-memory column
//                -> contraction with the per-subcell coefficient matrix.  Handles every simplex plan
//                (split cells, 1-D sets, any derivative order).  Low-degree elements take the
//                register-resident variant k_small (small.cuh), equispaced Lagrange elements the
//                product-form kernel k_lattice (lattice.cuh).
//   k_mma        single-cell Dubiner elements: a CTA owns a tile of PT points, runs the recurrence
//                level-parallel into a shared expansion table T[member][octet][alpha][8] and contracts
//                it with the 8x4 block-sparse coefficient matrix on the FP64 tensor pipe
//                (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), skipping all-zero blocks.
//   k_tensor     tensor-product elements (at most one vector-valued factor): factor tables per point
//                in shared memory, then the fused outer product streamed straight to global memory.
//   k_tensor_eval  fused point evaluation sum_dofs c[dof] phi_dof(x) on tensor-product elements (no table written).
//   k_locate     subcell bitmasks only.
//   k_zero_rows  zero-fill of the table rows no part of a wrapper element writes.
//
// Output: out[(alpha * total_rows + row) * ostride + point], rows placed through a DevRowMap
// (identity for plain elements); consecutive threads own consecutive points, so every warp store
// covers 256 contiguous bytes of one row.
#pragma once
#include "expansion.cuh"

// ---------------------------------------------------------------------------------------------
// thread-per-point kernel
// ---------------------------------------------------------------------------------------------
template <int SD, int ORDER>
__global__ void __launch_bounds__(128)
k_cellwise(const DevSimplex P, const __grid_constant__ RecTab tab, const DevEntity E, const double* __restrict__ pts,
           long long npts, long long ldp, double* __restrict__ out, long long ostride, int tables_in_smem,
           const __grid_constant__ DevRowMap M) {
    extern __shared__ double smem[];
    const int BP = blockDim.x;
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * BP + tid;
    const int na = (ORDER >= 0) ? Jet<SD, ORDER>::CAP : P.na;
    double* T = smem + tid;
    const int comp_stride = BP, slot_stride = na * BP;

    // per-subcell coefficient matrices and geometry: shared memory when they fit (split cells pick
    // them per thread), global memory otherwise
    const double* Call = P.ccell;
    const double* geom_all = P.geom;
    if (tables_in_smem) {
        double* s_C = smem + (size_t)P.nslots * na * BP;
        double* s_g = s_C + (size_t)P.ncells * P.nrows * P.nslots;
        for (int i = tid; i < P.ncells * P.nrows * P.nslots; i += BP) s_C[i] = __ldg(P.ccell + i);
        for (int i = tid; i < P.ncells * FB_GEOM_DOUBLES; i += BP) s_g[i] = __ldg(P.geom + i);
        __syncthreads();
        Call = s_C;
        geom_all = s_g;
    }
    if (p >= npts) return;

    double x[3];
    apply_entity<SD>(E, pts + p * ldp, x);
    unsigned mask = locate_cells<SD>(P.bary, P.ncells, P.unique, x);
    if (mask == 0) {            // in no subcell: zero column, like the reference
        fb_zero_column(M, out, ostride, p, na, P.nrows);
        return;
    }
    const double inv_mult = 1.0 / (double)__popc(mask);
    bool first = true;
    while (mask) {
        const int cell = __ffs(mask) - 1;
        mask &= mask - 1;
        expansion_point<SD, ORDER>(P, tab, geom_all + cell * FB_GEOM_DOUBLES, cell, inv_mult, x, T, slot_stride,
                                   comp_stride, na);
        const double* C = Call + (size_t)cell * P.nrows * P.nslots;
        if (ORDER >= 0) {
            // register tile: RB rows x NA derivative components; T is read once per row block
            constexpr int NA = Jet<SD, (ORDER >= 0 ? ORDER : 0)>::CAP;
            constexpr int RB = 4;
            for (int r0 = 0; r0 < P.nrows; r0 += RB) {
                double acc[RB][NA];
#pragma unroll
                for (int j = 0; j < RB; ++j)
#pragma unroll
                    for (int a = 0; a < NA; ++a) acc[j][a] = 0.0;
                const double* Cr[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) Cr[j] = C + (size_t)min(r0 + j, P.nrows - 1) * P.nslots;
                for (int k = 0; k < P.nslots; ++k) {
                    double t[NA];
                    const double* tk = T + (size_t)k * slot_stride;
#pragma unroll
                    for (int a = 0; a < NA; ++a) t[a] = tk[a * comp_stride];
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const double c = Cr[j][k];
#pragma unroll
                        for (int a = 0; a < NA; ++a) acc[j][a] = fma(c, t[a], acc[j][a]);
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
        } else {
            for (int r = 0; r < P.nrows; ++r) {
                const double* Cr = C + (size_t)r * P.nslots;
                for (int a = 0; a < na; ++a) {
                    double acc = 0.0;
                    for (int k = 0; k < P.nslots; ++k) acc = fma(Cr[k], T[(size_t)k * slot_stride + a * comp_stride], acc);
                    double sgn;
                    const size_t orow = fb_map_row(M, r, sgn);
                    double* o = out + ((size_t)a * M.total_rows + orow) * ostride + p;
                    *o = first ? sgn * acc : (*o + sgn * acc);
                }
            }
        }
        first = false;
    }
}

template <int SD>
__global__ void __launch_bounds__(128)
k_locate(const DevSimplex P, const DevEntity E, const double* __restrict__ pts, long long npts, long long ldp,
         int unique, unsigned* __restrict__ mask_out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    double x[3];
    apply_entity<SD>(E, pts + p * ldp, x);
    mask_out[p] = locate_cells<SD>(P.bary, P.ncells, unique, x);
}

// ---------------------------------------------------------------------------------------------
// tile kernel with FP64 tensor-pipe contraction
// ---------------------------------------------------------------------------------------------
struct MmaGeom {
    int PT;        // points per tile (multiple of 8)
    int ldT;       // doubles between member rows of T (>= na*PT, = 4 or 12 mod 16: conflict-free fragments)
    int logPT;     // log2(PT)
    int maxlev;    // most recurrence steps in one wavefront level
    int ppw;       // > 0: warp-local recurrence, points per warp (8, 16 or 32; tile = warps x ppw); 0: CTA-wide levels
    int threads;   // CTA size: 512 (one CTA per SM) or 256 (two)
    int skip;      // profiling only: bit 0 skips the recurrence, bit 1 the contraction, bit 3 the stores
};
#define FB_MMA_THREADS 512

// One CTA per SM, two phases that never overlap: FP64 vector work (recurrence) and FP64 tensor work
// (contraction) share one pipe on B200, and a DFMA issued between DMMAs waits for the 16-cycle
// DMMA in front of it, so interleaving the two phases of different CTAs costs more than it hides.
//
// T layout: T[slot][point group][alpha][PW points], PW = 16 (8 for the narrowest tile): 16 consecutive
// points are contiguous so that the recurrence's per-point accesses of a half warp hit 32 distinct
// banks, and every octet of a group is still 8 contiguous columns for the DMMA fragments.
// octets per contraction work item: ~16 DMMAs per coefficient fragment
__host__ __device__ constexpr int fb_mma_go(int na) {
    return na >= 8 ? 1 : (na >= 5 ? 2 : (na >= 3 ? 4 : (na == 2 ? 8 : 16)));
}

template <int SD, int ORDER, int PW>
__global__ void __launch_bounds__(FB_MMA_THREADS, 1)
k_mma(const DevSimplex P, const __grid_constant__ RecTab tab, const DevEntity E, const MmaGeom G,
      const double* __restrict__ pts, long long npts, long long ldp, double* __restrict__ out, long long ostride,
      const __grid_constant__ DevRowMap M) {
    constexpr int NA = Jet<SD, ORDER>::NA;
    extern __shared__ double smem[];
    double* T = smem;                                   // kpad x ldT
    double* s_fa = T + (size_t)P.kpad * G.ldT;          // 3 x PT
    double* s_fb = s_fa + 3 * G.PT;                     // 3 x PT
    __shared__ int s_next;
    const int tid = threadIdx.x;
    const int NT = blockDim.x;                          // 512 (one CTA per SM) or 256 (two)
    const int PT = G.PT;
    const long long base = (long long)blockIdx.x * PT;

    if (tid == 0) s_next = 0;
    if (G.ppw > 0) {
        // phases 0 + 1, warp-local: a warp owns ppw consecutive points of the tile (tile = warps x ppw points) and runs
        // the whole recurrence for them, level by level, with warp-level synchronisation only: lane = (step slot,
        // point), the lane's recurrence factors stay in registers, all step records sit in shared memory.  The
        // low wavefront levels have fewer (step, point) items than the CTA has threads, so the CTA-wide variant
        // below pays one block barrier and one latency-bound round per level with most warps idle; here all warps
        // run their own chains concurrently (P8 tet, 128-point tile: ~19k -> ~6k cycles per tile).
        StepRec* s_all = reinterpret_cast<StepRec*>(s_fa);
        for (int i = tid; i < tab.nsteps * 4; i += NT)
            reinterpret_cast<double*>(s_all)[i] = reinterpret_cast<const double*>(tab.steps)[i];
        for (int i = tid; i < (P.kpad - P.nslots) * G.ldT; i += NT) T[(size_t)P.nslots * G.ldT + i] = 0.0;
        const int ppw = G.ppw, lane1 = tid & 31;
        const int pl = (tid >> 5) * ppw + (lane1 & (ppw - 1));
        long long p = base + pl;
        if (p >= npts) p = npts - 1;                    // tail lanes repeat the last point, never stored
        double x[3], xr[3] = {0.0, 0.0, 0.0};
        apply_entity<SD>(E, pts + p * ldp, x);
#pragma unroll
        for (int i = 0; i < SD; ++i) {
            double sacc = 0.0;
#pragma unroll
            for (int d = 0; d < SD; ++d) sacc = fma(x[d], tab.geom0[i * SD + d], sacc);
            xr[i] = sacc + tab.geom0[9 + i];
        }
        double fa[3], fb[3];
        recurrence_factors<SD>(xr, fa, fb);
        double* Tcol = T + (pl / PW) * (PW * NA) + (pl % PW);
        if (lane1 < ppw) {
            double* t0 = Tcol + (size_t)tab.start_slot * G.ldT;
#pragma unroll
            for (int a = 0; a < NA; ++a) t0[a * PW] = (a == 0) ? tab.geom0[12] : 0.0;
        }
        __syncthreads();                                // step records staged
        const int slot0 = lane1 / ppw, nslot = 32 / ppw;
        for (int lev = 0; lev < ((G.skip & 1) ? 0 : tab.nlevels); ++lev) {
            const int l0 = tab.level_ptr[lev], nst = tab.level_ptr[lev + 1] - l0;
            for (int sl = slot0; sl < nst; sl += nslot) {
                const StepRec r = s_all[l0 + sl];
                run_step<SD, ORDER>(P, r, tab.geom0, fa, fb, Tcol, G.ldT, PW, NA);
            }
            __syncwarp();
        }
        __syncthreads();                                // every warp's columns are complete
    } else {
    // phase 0: points of the tile -> recurrence factors; member 0; zero padding rows
    if (tid < PT) {
        long long p = base + tid;
        if (p >= npts) p = npts - 1;                    // tail lanes repeat the last point, never stored
        double x[3], xr[3] = {0.0, 0.0, 0.0};
        apply_entity<SD>(E, pts + p * ldp, x);
#pragma unroll
        for (int i = 0; i < SD; ++i) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < SD; ++d) s = fma(x[d], tab.geom0[i * SD + d], s);
            xr[i] = s + tab.geom0[9 + i];
        }
        double fa[3], fb[3];
        recurrence_factors<SD>(xr, fa, fb);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s_fa[c * PT + tid] = fa[c];
            s_fb[c * PT + tid] = fb[c];
        }
        const double start = tab.geom0[12];
        double* t0 = T + (size_t)tab.start_slot * G.ldT + (tid / PW) * (PW * NA) + (tid % PW);
#pragma unroll
        for (int a = 0; a < NA; ++a) t0[a * PW] = (a == 0) ? start : 0.0;
    }
    for (int i = tid; i < (P.kpad - P.nslots) * G.ldT; i += NT) T[(size_t)P.nslots * G.ldT + i] = 0.0;
    __syncthreads();

    // phase 1: recurrence in wavefront order -- every member of total degree d is an independent
    // (step, point) work item once degrees d-1 and d-2 are in T.  The step records of a level are
    // staged from the constant bank into shared memory one level ahead (lanes of a warp work on
    // different steps, which the constant cache would serialise).
    StepRec* s_rec = reinterpret_cast<StepRec*>(s_fb + 3 * PT);          // 2 x G.maxlev records
    {
        const int n0 = tab.level_ptr[1] - tab.level_ptr[0];
        for (int i = tid; i < n0 * 4; i += NT)
            reinterpret_cast<double*>(s_rec)[i] = reinterpret_cast<const double*>(&tab.steps[tab.level_ptr[0]])[i];
    }
    __syncthreads();
    for (int lev = 0; lev < ((G.skip & 1) ? 0 : tab.nlevels); ++lev) {
        const int l0 = tab.level_ptr[lev];
        const int nst = tab.level_ptr[lev + 1] - l0;
        const StepRec* rec = s_rec + (lev & 1) * G.maxlev;
        if (lev + 1 < tab.nlevels) {
            const int l1 = tab.level_ptr[lev + 1];
            const int n1 = tab.level_ptr[lev + 2] - l1;
            double* dst = reinterpret_cast<double*>(s_rec + ((lev + 1) & 1) * G.maxlev);
            for (int i = tid; i < n1 * 4; i += NT) dst[i] = reinterpret_cast<const double*>(&tab.steps[l1])[i];
        }
        const int items = nst * PT;
        for (int it = tid; it < items; it += NT) {
            const int sl = it >> G.logPT, pl = it & (PT - 1);
            const double fa[3] = {s_fa[pl], s_fa[PT + pl], s_fa[2 * PT + pl]};
            const double fb[3] = {s_fb[pl], s_fb[PT + pl], s_fb[2 * PT + pl]};
            const StepRec r = rec[sl];
            run_step<SD, ORDER>(P, r, tab.geom0, fa, fb, T + (pl / PW) * (PW * NA) + (pl % PW), G.ldT, PW, NA);
        }
        __syncthreads();
    }

    }

    // phase 2: out[row, col] = sum_k C[row, k] T[k, col] on the FP64 tensor pipe (the C0 fix-ups
    // are folded into C).  Work item = (8-row block, GO point octets): one coefficient fragment feeds
    // GO * NA DMMAs.  A block multiplies ANY four member slots (gather packing, plan.py: pack_blocks): the B
    // fragment is read from the four matching rows of T, so a row block only pays for the members its rows
    // use.  Fragments are fetched CH blocks ahead so that their L2 latency hides behind
    // the DMMAs of the current chunk; row-block tables come from the constant bank.
    constexpr int CH = 8;
    constexpr int GO = fb_mma_go(NA);
    const int lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int ngrp = PT / (8 * GO);
    const int nitems = (G.skip & 2) ? 0 : tab.nrb * ngrp;
    const bool vec_ok = ((ostride & 1) == 0) && ((((size_t)out) & 15) == 0);
    const size_t astride = (size_t)M.total_rows * ostride;      // distance between derivative tables
    const double* Tlane = T + g;                        // + member slot * ldT, gathered per block
    const int ldT = G.ldT;
    // Work items are handed out dynamically (long and short row blocks alternate, plan.py: schedule_row_blocks, so that
    // the warps do not reach their store epilogues in lockstep).  The NEXT item and its first coefficient
    // fragments are fetched before the current item's stores are issued, so that the L2 latency of the fragments and
    // the shared-memory atomic overlap the store epilogue instead of delaying the next item's first DMMA.
    double a_cur[CH], a_nxt[CH];
    int kb_cur = 0, kb_nxt = 0;
    int item = 0;
    if (lane == 0) item = atomicAdd(&s_next, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item < nitems) {
        const int rb0 = tab.rb_order[item / ngrp];
        const int q0 = tab.blk_ptr[rb0], q1 = tab.blk_ptr[rb0 + 1];
#pragma unroll
        for (int j = 0; j < CH; ++j) a_cur[j] = (q0 + j < q1) ? __ldg(P.blk_frag + (size_t)(q0 + j) * 32 + lane) : 0.0;
        if (4 * q0 + lane < 4 * q1) kb_cur = __ldg(P.blk_kb + 4 * q0 + lane);
    }
    while (item < nitems) {
        const int rb = tab.rb_order[item / ngrp];
        const int oct0 = (item % ngrp) * GO;
        double acc[GO][NA][2];
#pragma unroll
        for (int o = 0; o < GO; ++o)
#pragma unroll
            for (int s = 0; s < NA; ++s) acc[o][s][0] = acc[o][s][1] = 0.0;
        const int q0 = tab.blk_ptr[rb], q1 = tab.blk_ptr[rb + 1];
        // octet o of the tile lives in point group o / (PW/8), at column offset (o % (PW/8)) * 8
        constexpr int OPG = PW / 8;                         // octets per point group
        const double* Titem = Tlane + (oct0 / OPG) * (PW * NA) + (oct0 % OPG) * 8;
        for (int q = q0; q < q1; q += CH) {
            if (q + CH < q1) {
#pragma unroll
                for (int j = 0; j < CH; ++j)
                    a_nxt[j] = (q + CH + j < q1) ? __ldg(P.blk_frag + (size_t)(q + CH + j) * 32 + lane) : 0.0;
                kb_nxt = (4 * (q + CH) + lane < 4 * q1) ? __ldg(P.blk_kb + 4 * (q + CH) + lane) : 0;
            }
            const int nj = min(CH, q1 - q);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (j < nj) {
                    // B fragment: lane (g, t) reads member slot blk_kb[4 (q + j) + t] at point g of every octet
                    const int slot = __shfl_sync(0xffffffffu, kb_cur, 4 * j + t);
                    const double* Tb = Titem + (size_t)slot * ldT;
                    double bfrag[GO * NA];
#pragma unroll
                    for (int s = 0; s < GO * NA; ++s) {
                        // column block of (octet s / NA past oct0, alpha s % NA); GO is a multiple of OPG or 1
                        const int o = s / NA, a = s % NA;
                        bfrag[s] = Tb[(o / OPG) * (PW * NA) + (o % OPG) * 8 + a * PW];
                    }
#pragma unroll
                    for (int s = 0; s < GO * NA; ++s)
                        dmma_8x8x4(acc[s / NA][s % NA][0], acc[s / NA][s % NA][1], a_cur[j], bfrag[s]);
                }
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) a_cur[j] = a_nxt[j];
            kb_cur = kb_nxt;
        }
        // next item and its first fragments (a_nxt / kb_nxt are free after the last chunk)
        int next = 0;
        if (lane == 0) next = atomicAdd(&s_next, 1);
        next = __shfl_sync(0xffffffffu, next, 0);
        kb_nxt = 0;
        if (next < nitems) {
            const int rbn = tab.rb_order[next / ngrp];
            const int n0 = tab.blk_ptr[rbn], n1 = tab.blk_ptr[rbn + 1];
#pragma unroll
            for (int j = 0; j < CH; ++j) a_nxt[j] = (n0 + j < n1) ? __ldg(P.blk_frag + (size_t)(n0 + j) * 32 + lane) : 0.0;
            if (4 * n0 + lane < 4 * n1) kb_nxt = __ldg(P.blk_kb + 4 * n0 + lane);
        }
        item = next;
        const int row = tab.row_perm[rb * 8 + g];           // table row of this lane's packed row (-1: padding)
        const long long p0 = base + oct0 * 8 + 2 * t;
        // warp-uniform: a full 8-row x GO-octet tile with aligned rows and no placement map
        const bool full_tile = GO >= 2 && vec_ok && M.identity && base + (oct0 + GO) * 8 <= npts && rb * 8 + 8 <= P.nrows;
        if (!(G.skip & 8)) {                                // (bit 3: profiling only, no stores)
        if (full_tile) {
            // trade fragments between lane groups g and g^4 so that one store instruction covers
            // 4 rows x 128 contiguous bytes (two octets) instead of 8 rows x 64 bytes
            const bool lo = g < 4;
            double* row_lo = out + (size_t)tab.row_perm[rb * 8 + (g & 3)] * ostride + base + oct0 * 8 + 2 * t;     // packed rows 0..3
            double* row_hi = out + (size_t)tab.row_perm[rb * 8 + 4 + (g & 3)] * ostride + base + oct0 * 8 + 2 * t; // packed rows 4..7
#pragma unroll
            for (int s = 0; s < NA; ++s) {
#pragma unroll
                for (int o = 0; o + 1 < GO; o += 2) {
                    // lanes g<4 send their octet o+1 piece, lanes g>=4 their octet o piece
                    const double sx = lo ? acc[o + 1][s][0] : acc[o][s][0];
                    const double sy = lo ? acc[o + 1][s][1] : acc[o][s][1];
                    const double rx = __shfl_xor_sync(0xffffffffu, sx, 16);
                    const double ry = __shfl_xor_sync(0xffffffffu, sy, 16);
                    const int ocol = (o + (lo ? 0 : 1)) * 8;
                    const double2 first = lo ? make_double2(acc[o][s][0], acc[o][s][1]) : make_double2(rx, ry);
                    *reinterpret_cast<double2*>(row_lo + (size_t)s * astride + ocol) = first;
                    const double2 second = lo ? make_double2(rx, ry) : make_double2(acc[o + 1][s][0], acc[o + 1][s][1]);
                    *reinterpret_cast<double2*>(row_hi + (size_t)s * astride + ocol) = second;
                }
            }
        } else if (row >= 0) {
            double sgn;
            const size_t orow = fb_map_row(M, row, sgn);
            if (!M.identity) {
#pragma unroll
                for (int o = 0; o < GO; ++o)
#pragma unroll
                    for (int s = 0; s < NA; ++s) {
                        acc[o][s][0] *= sgn;
                        acc[o][s][1] *= sgn;
                    }
            }
            double* rowp = out + orow * ostride + p0;
            // adjacent 64-byte pieces of a row are stored back to back (octet innermost)
#pragma unroll
            for (int s = 0; s < NA; ++s) {
                double* dsts = rowp + (size_t)s * astride;
#pragma unroll
                for (int o = 0; o < GO; ++o) {
                    const long long p = p0 + o * 8;
                    double* dst = dsts + o * 8;
                    if (vec_ok && p + 1 < npts) {
                        *reinterpret_cast<double2*>(dst) = make_double2(acc[o][s][0], acc[o][s][1]);
                    } else {
                        if (p < npts) dst[0] = acc[o][s][0];
                        if (p + 1 < npts) dst[1] = acc[o][s][1];
                    }
                }
            }
        }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) a_cur[j] = a_nxt[j];
        kb_cur = kb_nxt;
    }
}

// ---------------------------------------------------------------------------------------------
// tensor-product kernel
// ---------------------------------------------------------------------------------------------
template <int SD, int ORDER>
__device__ __forceinline__ void leaf_table(const DevTensorLeaf& L, const double* __restrict__ pt, double* scratch,
                                           double* table, int BP) {
    const DevSimplex& P = L.prog;
    double x[3];
    apply_entity<SD>(L.ent, pt + L.point_offset, x);
    const int na = P.na;
    unsigned mask = locate_cells<SD>(P.bary, P.ncells, P.unique, x);
    if (mask == 0) {            // in no subcell of a split factor: that factor's table is zero
        for (int i = 0; i < na * P.nrows; ++i) table[(size_t)i * BP] = 0.0;
        return;
    }
    const double inv_mult = 1.0 / (double)__popc(mask);
    bool first = true;
    while (mask) {
        const int cell = __ffs(mask) - 1;
        mask &= mask - 1;
        expansion_point<SD, ORDER>(P, *P.tab, P.geom + cell * FB_GEOM_DOUBLES, cell, inv_mult, x, scratch, na * BP, BP, na);
        const double* C = P.ccell + (size_t)cell * P.nrows * P.nslots;
        for (int r = 0; r < P.nrows; ++r) {
            for (int a = 0; a < na; ++a) {
                double acc = 0.0;
                for (int k = 0; k < P.nslots; ++k)
                    acc = fma(__ldg(C + (size_t)r * P.nslots + k), scratch[((size_t)k * na + a) * BP], acc);
                double* o = table + ((size_t)a * P.nrows + r) * BP;
                *o = first ? acc : (*o + acc);
            }
        }
        first = false;
    }
}

// Fused outer product over the leaf tables of one point: value = prod_l tab[l][i_l]; the product dof
// is the mixed-radix number of the leaf dofs, ((j0 * n1 + j1) * n2 + j2) * n3 + j3
// (FIAT/tensor_product.py:288-292); at most one leaf is vector valued and supplies the component
// (:293-335).
struct TensorOut {
    double* base;           // out + alpha * total_rows * ostride + point
    long long ostride;
    int nc_out, dof_base;
};

template <int L, int NL>
__device__ __forceinline__ void emit_products(double f, int dofacc, int comp, const double* const (&tab)[FB_MAX_LEAVES],
                                              const DevTensor& Q, const DevRowMap& M, int BP, const TensorOut& O) {
    const double* t = tab[L];
    const DevTensorLeaf& lf = Q.leaf[L];
    if constexpr (L == NL - 1) {
        if (lf.ncomp == 1) {
            // scalar innermost leaf: rows advance by nc_out
            const double sgn = M.sign[comp];
            double* o = O.base + ((size_t)(M.dof_base + dofacc) * M.nc_out + M.comp_out[comp]) * O.ostride;
            const long long step = (long long)M.nc_out * O.ostride;
            const double fs = sgn * f;
#pragma unroll 4
            for (int i = 0; i < lf.ndof; ++i) {
                *o = fs * t[(size_t)i * BP];
                o += step;
            }
        } else {
            for (int j = 0; j < lf.ndof; ++j)
                for (int k = 0; k < lf.ncomp; ++k) {
                    double* o = O.base + ((size_t)(M.dof_base + dofacc + j) * M.nc_out + M.comp_out[k]) * O.ostride;
                    *o = M.sign[k] * f * t[(size_t)(j * lf.ncomp + k) * BP];
                }
        }
    } else {
        if (lf.ncomp == 1) {
            for (int j = 0; j < lf.ndof; ++j)
                emit_products<L + 1, NL>(f * t[(size_t)j * BP], dofacc + j * lf.dof_stride, comp, tab, Q, M, BP, O);
        } else {
            for (int j = 0; j < lf.ndof; ++j)
                for (int k = 0; k < lf.ncomp; ++k)
                    emit_products<L + 1, NL>(f * t[(size_t)(j * lf.ncomp + k) * BP], dofacc + j * lf.dof_stride, k, tab, Q, M,
                                             BP, O);
        }
    }
}

template <int ORDER>
__global__ void __launch_bounds__(128)
k_tensor(const DevTensor Q, const double* __restrict__ pts, long long npts, long long ldp,
         double* __restrict__ out, long long ostride, const __grid_constant__ DevRowMap M) {
    extern __shared__ double smem[];
    const int BP = blockDim.x;
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * BP + tid;
    if (p >= npts) return;
    double* scratch = smem + tid;
    const double* pt = pts + p * ldp;
    for (int l = 0; l < Q.nleaf; ++l) {
        const DevTensorLeaf& L = Q.leaf[l];
        double* table = smem + (size_t)L.table_off * BP + tid;
        if (L.prog.sd == 1) leaf_table<1, ORDER>(L, pt, scratch, table, BP);
        else if (L.prog.sd == 2) leaf_table<2, ORDER>(L, pt, scratch, table, BP);
        else leaf_table<3, ORDER>(L, pt, scratch, table, BP);
    }
    for (int al = 0; al < Q.nalpha; ++al) {
        const int* aidx = Q.alpha_leaf + al * FB_MAX_LEAVES;
        const double* tab[FB_MAX_LEAVES];
#pragma unroll
        for (int l = 0; l < FB_MAX_LEAVES; ++l) {
            const int lo = l < Q.nleaf ? l : 0;
            tab[l] = smem + ((size_t)Q.leaf[lo].table_off + (size_t)__ldg(aidx + lo) * Q.leaf[lo].prog.nrows) * BP + tid;
        }
        TensorOut O;
        O.base = out + (size_t)al * M.total_rows * ostride + p;
        O.ostride = ostride;
        switch (Q.nleaf) {
            case 1: emit_products<0, 1>(1.0, 0, 0, tab, Q, M, BP, O); break;
            case 2: emit_products<0, 2>(1.0, 0, 0, tab, Q, M, BP, O); break;
            case 3: emit_products<0, 3>(1.0, 0, 0, tab, Q, M, BP, O); break;
            default: emit_products<0, 4>(1.0, 0, 0, tab, Q, M, BP, O); break;
        }
    }
}

// Fused point evaluation on tensor-product elements: u_f(x) = sum_dofs coef[f][dof] prod_l tab_l[i_l](x_l) for every
// derivative multi-index, by nested partial sums over the leaves (innermost leaf contracted first) -- the table of
// prod(n_l) values per point (42.6 kB per point for GLL Q10 on a hexahedron) is never formed, let alone written.
template <int L, int NL>
__device__ __forceinline__ double eval_products(const double* __restrict__ coef, int dofacc,
                                                const double* const (&tab)[FB_MAX_LEAVES], const DevTensor& Q, int BP) {
    const double* t = tab[L];
    const DevTensorLeaf& lf = Q.leaf[L];
    double s = 0.0;
    if constexpr (L == NL - 1) {
        const double* c = coef + dofacc;
#pragma unroll 4
        for (int i = 0; i < lf.ndof; ++i) s = fma(__ldg(c + i), t[(size_t)i * BP], s);
    } else {
        for (int j = 0; j < lf.ndof; ++j)
            s = fma(t[(size_t)j * BP], eval_products<L + 1, NL>(coef, dofacc + j * lf.dof_stride, tab, Q, BP), s);
    }
    return s;
}

template <int ORDER>
__global__ void __launch_bounds__(128)
k_tensor_eval(const DevTensor Q, const double* __restrict__ coef, int nfunc, const double* __restrict__ pts, long long npts,
              long long ldp, double* __restrict__ out, long long ostride) {
    extern __shared__ double smem[];
    const int BP = blockDim.x;
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * BP + tid;
    if (p >= npts) return;
    double* scratch = smem + tid;
    const double* pt = pts + p * ldp;
    for (int l = 0; l < Q.nleaf; ++l) {
        const DevTensorLeaf& L = Q.leaf[l];
        double* table = smem + (size_t)L.table_off * BP + tid;
        if (L.prog.sd == 1) leaf_table<1, ORDER>(L, pt, scratch, table, BP);
        else if (L.prog.sd == 2) leaf_table<2, ORDER>(L, pt, scratch, table, BP);
        else leaf_table<3, ORDER>(L, pt, scratch, table, BP);
    }
    for (int al = 0; al < Q.nalpha; ++al) {
        const int* aidx = Q.alpha_leaf + al * FB_MAX_LEAVES;
        const double* tab[FB_MAX_LEAVES];
#pragma unroll
        for (int l = 0; l < FB_MAX_LEAVES; ++l) {
            const int lo = l < Q.nleaf ? l : 0;
            tab[l] = smem + ((size_t)Q.leaf[lo].table_off + (size_t)__ldg(aidx + lo) * Q.leaf[lo].prog.nrows) * BP + tid;
        }
        for (int f = 0; f < nfunc; ++f) {
            const double* c = coef + (size_t)f * Q.nrows;
            double u;
            switch (Q.nleaf) {
                case 1: u = eval_products<0, 1>(c, 0, tab, Q, BP); break;
                case 2: u = eval_products<0, 2>(c, 0, tab, Q, BP); break;
                case 3: u = eval_products<0, 3>(c, 0, tab, Q, BP); break;
                default: u = eval_products<0, 4>(c, 0, tab, Q, BP); break;
            }
            out[((size_t)al * nfunc + f) * ostride + p] = u;
        }
    }
}

// Sum-factorised variant for three line factors (quadrilateral x interval = hexahedron): the innermost factor's
// table sits in registers, a row of coefficients is loaded once and contracted against all derivative orders of
// that factor at once, and the partial sums climb the factors:
//   inner[a2]    = sum_i2 c[i0, i1, i2] T2[a2][i2]
//   mid[a1][a2] += T1[a1][i1] inner[a2]
//   u[alpha]    += T0[a0][i0] mid[a1][a2]          alpha = (a0, a1, a2) in mis order
// 1331 coefficient loads and ~3.2 k FMAs per function for GLL Q10 at order 1 instead of 5.9 k of each.
#define FB_EVAL_NMAX 12
template <int ORDER>
__global__ void __launch_bounds__(128)
k_hex_eval(const DevTensor Q, const double* __restrict__ coef, int nfunc, const double* __restrict__ pts, long long npts,
           long long ldp, double* __restrict__ out, long long ostride) {
    constexpr int NO = ORDER + 1;
    extern __shared__ double smem[];
    const int BP = blockDim.x;
    const int tid = threadIdx.x;
    const long long p = (long long)blockIdx.x * BP + tid;
    if (p >= npts) return;
    double* scratch = smem + tid;
    const double* pt = pts + p * ldp;
    for (int l = 0; l < 3; ++l) {
        const DevTensorLeaf& L = Q.leaf[l];
        leaf_table<1, ORDER>(L, pt, scratch, smem + (size_t)L.table_off * BP + tid, BP);
    }
    const int n0 = Q.leaf[0].ndof, n1 = Q.leaf[1].ndof, n2 = Q.leaf[2].ndof;
    const double* T0 = smem + (size_t)Q.leaf[0].table_off * BP + tid;       // [a][i] at (a * n + i) * BP
    const double* T1 = smem + (size_t)Q.leaf[1].table_off * BP + tid;
    const double* T2s = smem + (size_t)Q.leaf[2].table_off * BP + tid;
    double T2[NO][FB_EVAL_NMAX];
#pragma unroll
    for (int a = 0; a < NO; ++a)
#pragma unroll
        for (int i = 0; i < FB_EVAL_NMAX; ++i) T2[a][i] = i < n2 ? T2s[(size_t)(a * n2 + i) * BP] : 0.0;
    for (int f = 0; f < nfunc; ++f) {
        const double* c = coef + (size_t)f * Q.nrows;
        double u[NO][NO][NO];
#pragma unroll
        for (int a0 = 0; a0 < NO; ++a0)
#pragma unroll
            for (int a1 = 0; a1 < NO; ++a1)
#pragma unroll
                for (int a2 = 0; a2 < NO; ++a2) u[a0][a1][a2] = 0.0;
        for (int i0 = 0; i0 < n0; ++i0) {
            double mid[NO][NO];
#pragma unroll
            for (int a1 = 0; a1 < NO; ++a1)
#pragma unroll
                for (int a2 = 0; a2 < NO; ++a2) mid[a1][a2] = 0.0;
#pragma unroll 2
            for (int i1 = 0; i1 < n1; ++i1) {
                const double* row = c + ((size_t)i0 * n1 + i1) * n2;
                // the coefficient row is zero-padded in registers; even and odd members accumulate separately
                // (two shorter dependent chains per derivative order)
                double cv[FB_EVAL_NMAX];
#pragma unroll
                for (int i2 = 0; i2 < FB_EVAL_NMAX; ++i2) cv[i2] = i2 < n2 ? __ldg(row + i2) : 0.0;
                double inner[NO][2];
#pragma unroll
                for (int a2 = 0; a2 < NO; ++a2) inner[a2][0] = inner[a2][1] = 0.0;
#pragma unroll
                for (int i2 = 0; i2 < FB_EVAL_NMAX; ++i2)
#pragma unroll
                    for (int a2 = 0; a2 < NO; ++a2) inner[a2][i2 & 1] = fma(cv[i2], T2[a2][i2], inner[a2][i2 & 1]);
#pragma unroll
                for (int a1 = 0; a1 < NO; ++a1) {
                    const double t1 = T1[(size_t)(a1 * n1 + i1) * BP];
#pragma unroll
                    for (int a2 = 0; a2 < NO; ++a2)
                        if (a1 + a2 <= ORDER) mid[a1][a2] = fma(t1, inner[a2][0] + inner[a2][1], mid[a1][a2]);
                }
            }
#pragma unroll
            for (int a0 = 0; a0 < NO; ++a0) {
                const double t0 = T0[(size_t)(a0 * n0 + i0) * BP];
#pragma unroll
                for (int a1 = 0; a1 < NO; ++a1)
#pragma unroll
                    for (int a2 = 0; a2 < NO; ++a2)
                        if (a0 + a1 + a2 <= ORDER) u[a0][a1][a2] = fma(t0, mid[a1][a2], u[a0][a1][a2]);
            }
        }
        // product multi-indices in mis order: total order ascending, first entry descending
        int al = 0;
#pragma unroll
        for (int k = 0; k <= ORDER; ++k)
#pragma unroll
            for (int a0 = k; a0 >= 0; --a0)
#pragma unroll
                for (int a1 = k - a0; a1 >= 0; --a1) {
                    out[((size_t)al * nfunc + f) * ostride + p] = u[a0][a1][k - a0 - a1];
                    ++al;
                }
    }
}

// Weights of a fused point evaluation u_f = sum_i coef[f][i] phi_i: tabulation is linear in the coefficient tensor
// (FIAT/polynomial_set.py:71), so the functions' derivative tables are the tables of W = coef . C,
//   W[cell][(j * nfunc + f) * ncomp + c][k] = sum_i coef[f][i] * C[cell][(j * ndofs + i) * ncomp + c][k],
// for the stacked derived element C (rows: table j, dof i, component c; fiat_b200/plan.py: stacked_derived).  Formed
// on the device per call, so new coefficients need no re-planning.
__global__ void k_eval_weights(const double* __restrict__ C, int ncells, int nstack, int ndofs, int ncomp, int nslots,
                               const double* __restrict__ coef, int nfunc, double* __restrict__ W) {
    const int reval = nstack * nfunc * ncomp;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)ncells * reval * nslots) return;
    const int k = (int)(idx % nslots);
    const int er = (int)((idx / nslots) % reval);
    const int cell = (int)(idx / ((long long)nslots * reval));
    const int c = er % ncomp, f = (er / ncomp) % nfunc, j = er / (ncomp * nfunc);
    const double* Cc = C + ((size_t)cell * nstack * ndofs * ncomp + (size_t)j * ndofs * ncomp + c) * nslots + k;
    const double* u = coef + (size_t)f * ndofs;
    double s = 0.0;
    for (int i = 0; i < ndofs; ++i) s = fma(__ldg(u + i), __ldg(Cc + (size_t)i * ncomp * nslots), s);
    W[idx] = s;
}

// The same weights (single-cell plans) as the A fragments of a dense grid of 8x4 blocks for the tile kernel:
// block (rb, kb) multiplies member slots 4 kb .. 4 kb + 3 (slots >= nslots are zero rows of the expansion table).
__global__ void k_eval_fragments(const double* __restrict__ W, int reval, int nslots, int nkb, double* __restrict__ frag,
                                 int* __restrict__ slots) {
    const int q = blockIdx.x;                   // block rb * nkb + kb
    const int lane = threadIdx.x;
    const int rb = q / nkb, kb = q % nkb;
    const int row = rb * 8 + (lane >> 2), slot = kb * 4 + (lane & 3);
    frag[(size_t)q * 32 + lane] = (row < reval && slot < nslots) ? W[(size_t)row * nslots + slot] : 0.0;
    if (lane < 4) slots[(size_t)q * 4 + lane] = kb * 4 + lane;
}

// zero-fill rows that no part of a wrapper element writes (all derivative tables)
__global__ void k_zero_rows(double* __restrict__ out, long long ostride, long long npts, long long total_rows,
                            int nalpha, const int* __restrict__ rows, int nrows) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    for (int a = 0; a < nalpha; ++a)
        for (int i = 0; i < nrows; ++i) out[((size_t)a * total_rows + __ldg(rows + i)) * ostride + p] = 0.0;
}
