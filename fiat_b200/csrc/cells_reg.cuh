// Split-cell tile kernel with the expansion values in REGISTERS (B operand) and the coefficient blocks streamed
// through shared memory (A operand): elements on split (macro) complexes with <= 64 members per subcell; the launcher
// (cells_launch.cu) takes it for <= 20 members, where it beats the segment kernel of cells.cuh.
//
// Same input as cells.cuh (the order-0 derived element of fiat_b200/plan.py: macro_merged, one block-sparse
// coefficient matrix per subcell; FIAT/expansions.py:449-490), other orientation.  cells.cuh gives a warp one 8-row
// block for all columns, so every (row block, subcell) pair is a segment of 1-5 blocks x <= 8 octets with its own
// fragment fetch, slot shuffle, dispatch and scatter: 21 instructions per DMMA on Walkington's element, tensor pipe
// 31 % busy.  Here
//   * a warp owns two octets of COLUMNS for the whole tile and reads their B fragments -- the member values of its 16
//     points, k-block j = member slots 4j..4j+3 -- from the expansion table into registers ONCE (KB x 2 doubles);
//   * all warps walk the row blocks in lockstep, RB at a time ("step"); a step's coefficient blocks of all subcells
//     (prefix packing of fixed k-blocks, plan.py: pack_fixed_stream) are one contiguous run of global memory that
//     the CTA copies into shared memory with cp.async one step ahead; per block a warp issues one conflict-free
//     LDS.64 and one DMMA per octet (both octets in the same subcell, the usual case for sorted columns, share it);
//   * the 8 x 8 results go through the column permutation into a CTA-wide staging buffer (double-buffered) and the
//     previous step's RB * 8 rows leave as full-width coalesced row stores while the current step's DMMAs run;
//   * one block barrier per step.
// The expansion table is dead once the B fragments are loaded, so the staging buffers reuse its shared memory.
// Phases 0 (bit-exact subcell binning, FIAT/expansions.py:771-811, columns sorted by subcell), 1 (warp-local value
// recurrence, FIAT/expansions.py:202-249) and 3 (points in several / no subcells) are those of cells.cuh.
#pragma once
#include "expansion.cuh"

struct CellsRegGeom {
    int PT;        // points per tile
    int PTS;       // column capacity = warps x 16 >= PT + 8 * ncells (subcell ranges padded to octets)
    int ldT;       // doubles between member rows of T (>= PTS, = 8 mod 16)
    int maxlev;    // step records held in shared memory
    int threads;
    int SP;        // staging row stride in doubles (PT + 2)
    int astage;    // doubles per coefficient staging buffer (>= longest step, even)
    int uoff;      // doubles before the phase union region (column permutation, octet -> subcell)
};

__device__ __forceinline__ void fb_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void fb_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// The n stored k-blocks (a prefix 0 .. n - 1, plan.py: pack_fixed_stream) of one (subcell, row block) against the
// warp's B fragments: ONE jump on n into an unrolled fall-through run of blocks, each a shared-memory load with an
// immediate offset and one DMMA per octet -- the B fragment must be a compile-time register, and a per-block test
// costs more than the block.  (Two earlier forms of this loop, measured on Walkington's element against the segment
// kernel's 2.7 ms: an unrolled `if (mask & (1 << kb))` chain is if-converted into KB *predicated* DMMAs per row
// block, and a predicated-off DMMA still takes its slot in the tensor pipe -- 148 M issued for 65 M wanted, 3.9 ms;
// a loop over the set bits with a switch on the k-block number costs 28 instructions per block -- 5.4 ms.)
#define FB_CELLS_REG_BLOCK(K)                                                    \
    if (K < KB) {                                                                \
        const double a = f[(K < KB ? K : 0) * 32];                               \
        dmma_8x8x4(a00, a01, a, B0[K < KB ? K : 0]);                             \
        if (JOINT) dmma_8x8x4(a10, a11, a, B1[K < KB ? K : 0]);                  \
    }

template <int KB, bool JOINT>
__device__ __forceinline__ void cells_reg_blocks(int n, const double* __restrict__ f, const double (&B0)[KB],
                                                 const double (&B1)[KB], double& a00, double& a01, double& a10,
                                                 double& a11) {
    // binary search for the entry point (a `switch` became a linear chain of 14 compare-and-branch groups);
    // entering at E<n> runs blocks n - 1, ..., 0
    if (n >= 9) {
        if (n >= 13) {
            if (n >= 15) { if (n >= 16) goto E16; goto E15; }
            if (n >= 14) goto E14;
            goto E13;
        }
        if (n >= 11) { if (n >= 12) goto E12; goto E11; }
        if (n >= 10) goto E10;
        goto E9;
    }
    if (n >= 5) {
        if (n >= 7) { if (n >= 8) goto E8; goto E7; }
        if (n >= 6) goto E6;
        goto E5;
    }
    if (n >= 3) { if (n >= 4) goto E4; goto E3; }
    if (n >= 2) goto E2;
    if (n >= 1) goto E1;
    goto E0;
E16: FB_CELLS_REG_BLOCK(15)
E15: FB_CELLS_REG_BLOCK(14)
E14: FB_CELLS_REG_BLOCK(13)
E13: FB_CELLS_REG_BLOCK(12)
E12: FB_CELLS_REG_BLOCK(11)
E11: FB_CELLS_REG_BLOCK(10)
E10: FB_CELLS_REG_BLOCK(9)
E9: FB_CELLS_REG_BLOCK(8)
E8: FB_CELLS_REG_BLOCK(7)
E7: FB_CELLS_REG_BLOCK(6)
E6: FB_CELLS_REG_BLOCK(5)
E5: FB_CELLS_REG_BLOCK(4)
E4: FB_CELLS_REG_BLOCK(3)
E3: FB_CELLS_REG_BLOCK(2)
E2: FB_CELLS_REG_BLOCK(1)
E1: FB_CELLS_REG_BLOCK(0)
E0:;
}
#undef FB_CELLS_REG_BLOCK

template <int SD, int KB>
__global__ void __launch_bounds__(512, 1)
k_cells_reg(const DevSimplex P, const __grid_constant__ RecTab tab, const __grid_constant__ SmallTab st,
            const DevEntity E, const CellsRegGeom G, const double* __restrict__ pts, long long npts, long long ldp,
            double* __restrict__ out, long long ostride) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    const int PT = G.PT, PTS = G.PTS, ldT = G.ldT, SP = G.SP;
    int* s_perm = reinterpret_cast<int*>(smem);              // PTS: column -> point of the tile (-1 padding)
    int* s_octcell = s_perm + PTS;                           // PTS / 8: subcell of an octet of columns (-1 unused)
    double* U = smem + G.uoff;
    double* T = U;                                           // phases 0-1: kpad x ldT
    double* s_fa = T + (size_t)P.kpad * ldT;                 // 3 x PTS
    double* s_fb = s_fa + 3 * PTS;                           // 3 x PTS
    StepRec* s_rec = reinterpret_cast<StepRec*>(s_fb + 3 * PTS);
    const int RB = P.crb, R8 = RB * 8;
    double* O = U;                                           // phase 2: 2 x (RB * 8) x SP result rows
    double* Abuf = U + 2 * (size_t)R8 * SP;                  // 2 x astage coefficient steps (SP is even)
    __shared__ int s_cnt[32], s_off[33], s_fill[32], s_cols;
    const long long base = (long long)blockIdx.x * PT;

    // ---- phase 0: subcell of every point, columns sorted by subcell -----------------------------
    if (tid < 32) { s_cnt[tid] = 0; s_fill[tid] = 0; }
    for (int j = tid; j < PTS; j += NT) {
        s_perm[j] = -1;
        if (j < PTS / 8) s_octcell[j] = -1;
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_fa[c * PTS + j] = 0.0; s_fb[c * PTS + j] = 0.0; }
    }
    for (int i = tid; i < P.kpad * ldT; i += NT) T[i] = 0.0;
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
        int off = 0;
        for (int c = 0; c < P.ncells; ++c) {
            s_off[c] = off;
            const int noct = (s_cnt[c] + 7) >> 3;
            for (int o = 0; o < noct; ++o) s_octcell[off / 8 + o] = c;
            off += noct * 8;
        }
        s_off[P.ncells] = off;
        s_cols = off;
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
    for (int i = tid; i < tab.nsteps * 4; i += NT)
        reinterpret_cast<double*>(s_rec)[i] = reinterpret_cast<const double*>(tab.steps)[i];
    __syncthreads();
    const int ncols = s_cols;

    // ---- phase 1: value recurrence, warp-local (a warp owns blocks of 16 columns) ---------------------
    for (int cb = warp; cb * 16 < ncols; cb += nwarp) {
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

    // ---- B fragments of this warp's two octets into registers ------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int noct = ncols >> 3;
    const int o0 = 2 * warp, o1 = o0 + 1;
    const int c0 = o0 < noct ? s_octcell[o0] : -1;
    const int c1 = o1 < noct ? s_octcell[o1] : -1;
    double B0[KB], B1[KB];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        const bool in = 4 * kb < P.kpad;
        B0[kb] = (in && c0 >= 0) ? T[(size_t)(4 * kb + t) * ldT + o0 * 8 + g] : 0.0;
        B1[kb] = (in && c1 >= 0) ? T[(size_t)(4 * kb + t) * ldT + o1 * 8 + g] : 0.0;
    }
    // result columns of this lane's accumulator pairs, through the column permutation (-1: padding column)
    const int p00 = c0 >= 0 ? s_perm[o0 * 8 + 2 * t] : -1, p01 = c0 >= 0 ? s_perm[o0 * 8 + 2 * t + 1] : -1;
    const int p10 = c1 >= 0 ? s_perm[o1 * 8 + 2 * t] : -1, p11 = c1 >= 0 ? s_perm[o1 * 8 + 2 * t + 1] : -1;
    __syncthreads();                                         // T is dead: its memory becomes the staging buffers

    // ---- phase 2: steps of RB row blocks ----------------------------------------------------------------
    // One loop body, every piece of it written once (no lambdas: their captured pointers lose the shared address
    // space and every access then pays a generic-to-shared conversion): iteration s waits for the coefficients of
    // step s, issues the copy of step s + 1, stores the rows of step s - 1 and computes step s.
    const int nsteps = P.cnsteps;
    const int hdr = ((P.ncells * RB + 3) >> 2) * 2;          // doubles of the step's int32 records
    const bool vec_ok = ((ostride & 1) == 0) && ((((size_t)out) & 15) == 0) && ((base & 1) == 0) && base + PT <= npts;
    const unsigned a_shared = (unsigned)__cvta_generic_to_shared(Abuf);
    const size_t stage_bytes = (size_t)G.astage * sizeof(double);
    for (int s = -1; s <= nsteps; ++s) {
        if (s >= 0 && s < nsteps) fb_cp_async_wait_all();
        __syncthreads();          // step s staged; every warp is done with step s - 1 and with the stores of step s - 2
        if (s + 1 < nsteps) {
            const int b0 = __ldg(P.cstep_ptr + s + 1), b1 = __ldg(P.cstep_ptr + s + 2);
            const char* src = reinterpret_cast<const char*>(P.cstream + b0) + tid * 16;
            unsigned dst = a_shared + (unsigned)(((s + 1) & 1) * stage_bytes) + tid * 16;
            for (int i = tid * 2; i < b1 - b0; i += NT * 2, src += NT * 16, dst += NT * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
            fb_cp_async_commit();
        }
        if (s >= 1) {             // rows of step s - 1: full-width coalesced row stores
            const double* Ob = O + (size_t)((s - 1) & 1) * R8 * SP;
            for (int r = warp; r < R8; r += nwarp) {
                const int rbi = (s - 1) * RB + (r >> 3);
                const int row = rbi < tab.nrb ? (int)tab.row_perm[rbi * 8 + (r & 7)] : -1;
                if (row < 0) continue;
                double* rowp = out + (size_t)row * ostride + base;
                const double* src = Ob + (size_t)r * SP;
                if (vec_ok) {
                    // <= 128 double2 per row (PTS <= 256): all loads first, then all stores
                    const double2* s2 = reinterpret_cast<const double2*>(src) + lane;
                    double2* d2 = reinterpret_cast<double2*>(rowp) + lane;
                    const int n2 = PT >> 1;
                    double2 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (lane + 32 * k < n2) v[k] = s2[32 * k];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (lane + 32 * k < n2) d2[32 * k] = v[k];
                } else {
                    for (int i = lane; i < PT; i += 32)
                        if (base + i < npts) rowp[i] = src[i];
                }
            }
        }
        if (s < 0 || s >= nsteps || c0 < 0) continue;
        const double* Ab = Abuf + (size_t)(s & 1) * G.astage;
        const int* meta = reinterpret_cast<const int*>(Ab);
        const double* frag = Ab + hdr + lane;
        double* Ob = O + (size_t)(s & 1) * R8 * SP + (size_t)g * SP;
#pragma unroll 1
        for (int r = 0; r < RB; ++r) {
            double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
            const int m0 = meta[c0 * RB + r];
            const double* f = frag + (size_t)(m0 >> 16) * 32;
            if (c1 == c0) {
                cells_reg_blocks<KB, true>(m0 & 0xffff, f, B0, B1, a00, a01, a10, a11);
            } else {
                cells_reg_blocks<KB, false>(m0 & 0xffff, f, B0, B0, a00, a01, a10, a11);
                if (c1 >= 0) {
                    const int m1 = meta[c1 * RB + r];
                    cells_reg_blocks<KB, false>(m1 & 0xffff, frag + (size_t)(m1 >> 16) * 32, B1, B1, a10, a11, a00, a01);
                }
            }
            double* orow = Ob + (size_t)r * 8 * SP;
            if (p00 >= 0) orow[p00] = a00;
            if (p01 >= 0) orow[p01] = a01;
            if (p10 >= 0) orow[p10] = a10;
            if (p11 >= 0) orow[p11] = a11;
        }
    }
    __syncthreads();

    // ---- phase 3: points shared by several subcells, or in none -----------------------------------------
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
