// Plan-compilation helpers (host only, no CUDA calls): combinatorial optimisation of the block-sparse packing
// the DMMA tile kernels stream (fiat_b200/plan.py: pack_gather).
//
// The contraction out = C . T (FIAT/polynomial_set.py:68-72) runs as mma.m8n8k4 blocks: 8 table rows x 4 expansion
// members.  A block may gather ANY four members (the B fragment is read from shared memory through per-block
// member indices), so a group of 8 rows costs ceil(|union of the rows' supports| / 4) blocks.  Two free choices
// remain and are optimised here:
//   * which rows share a group of 8          -> fiatb200_cluster_rows   (swap local search on the union sizes)
//   * the slot number (mod 4) of each member -> fiatb200_colour_members (the 4 rows of T a block reads are
//     bank-conflict free iff their slots differ mod 4; the kernels tolerate conflicts, this keeps them rare)
// For split-cell elements the columns come in `nseg` segments of equal width (one per subcell, same members):
// a group pays for every segment separately and all segments share the member numbering.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/fiat_b200.h"

namespace {

struct Rng {        // xorshift64*: deterministic plans on every platform
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull) {}
    uint64_t next() {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return s * 0x2545F4914F6CDD1Dull;
    }
    int below(int n) { return (int)(next() % (uint64_t)n); }
};

// bit matrix: rows x (nseg segments of W words each)
struct Bits {
    int nseg, W;
    std::vector<uint64_t> w;
    const uint64_t* row(int r) const { return w.data() + (size_t)r * nseg * W; }
};

inline int seg_blocks(const uint64_t* u, int nseg, int W, int* total_union) {
    int blocks = 0, tot = 0;
    for (int s = 0; s < nseg; ++s) {
        int n = 0;
        for (int k = 0; k < W; ++k) n += __builtin_popcountll(u[s * W + k]);
        blocks += (n + 3) >> 2;
        tot += n;
    }
    *total_union = tot;
    return blocks;
}

}  // namespace

extern "C" {

int fiatb200_cluster_rows(const uint8_t* support, int32_t nrows, int32_t ncols, int32_t nseg, int32_t greedy_init,
                          int32_t* order, int64_t iters, uint64_t seed, int32_t* blocks_out) {
    if (!support || !order || nrows < 1 || ncols < 1 || nseg < 1 || ncols % nseg) return FIATB200_ERR_ARG;
    const int width = ncols / nseg;
    const int W = (width + 63) / 64;
    const int npad = (nrows + 7) & ~7;
    const int G = npad / 8;
    const int RW = nseg * W;
    Bits B{nseg, W, std::vector<uint64_t>((size_t)npad * RW, 0)};
    for (int r = 0; r < nrows; ++r)
        for (int c = 0; c < ncols; ++c)
            if (support[(size_t)r * ncols + c]) {
                const int s = c / width, k = c % width;
                B.w[(size_t)r * RW + s * W + (k >> 6)] |= 1ull << (k & 63);
            }
    std::vector<int> grp(npad);
    if (greedy_init) {
        // seed order: start a group from the widest remaining row, keep adding the row that enlarges the group's
        // union least (ties: largest overlap)
        std::vector<int> width_of(nrows), left(nrows);
        for (int r = 0; r < nrows; ++r) {
            left[r] = r;
            int n = 0;
            for (int k = 0; k < RW; ++k) n += __builtin_popcountll(B.row(r)[k]);
            width_of[r] = n;
        }
        std::vector<uint64_t> u(RW);
        int pos = 0;
        while (!left.empty()) {
            size_t best = 0;
            for (size_t i = 1; i < left.size(); ++i)
                if (width_of[left[i]] > width_of[left[best]]) best = i;
            int seed_row = left[best];
            left.erase(left.begin() + best);
            order[pos++] = seed_row;
            for (int k = 0; k < RW; ++k) u[k] = B.row(seed_row)[k];
            for (int n = 1; n < 8 && !left.empty(); ++n) {
                size_t pick = 0;
                int best_grow = 1 << 30, best_share = -1;
                for (size_t i = 0; i < left.size(); ++i) {
                    const uint64_t* r = B.row(left[i]);
                    int grow = 0, share = 0;
                    for (int k = 0; k < RW; ++k) {
                        grow += __builtin_popcountll(r[k] & ~u[k]);
                        share += __builtin_popcountll(r[k] & u[k]);
                    }
                    if (grow < best_grow || (grow == best_grow && share > best_share)) {
                        best_grow = grow; best_share = share; pick = i;
                    }
                }
                const int r = left[pick];
                left.erase(left.begin() + pick);
                order[pos++] = r;
                for (int k = 0; k < RW; ++k) u[k] |= B.row(r)[k];
            }
        }
    }
    {
        std::vector<char> seen(npad, 0);
        for (int i = 0; i < nrows; ++i) {
            if (order[i] < 0 || order[i] >= nrows || seen[order[i]]) return FIATB200_ERR_ARG;
            seen[order[i]] = 1;
            grp[i] = order[i];
        }
        for (int i = nrows; i < npad; ++i) grp[i] = i;      // padding rows: empty support
    }
    std::vector<uint64_t> uni((size_t)G * RW), without((size_t)16 * RW), cand(RW);
    std::vector<int> cost(G), usum(G);
    auto refresh = [&](int g) {
        uint64_t* u = uni.data() + (size_t)g * RW;
        memset(u, 0, sizeof(uint64_t) * RW);
        for (int i = 0; i < 8; ++i) {
            const uint64_t* r = B.row(grp[g * 8 + i]);
            for (int k = 0; k < RW; ++k) u[k] |= r[k];
        }
        cost[g] = seg_blocks(u, nseg, W, &usum[g]);
    };
    for (int g = 0; g < G; ++g) refresh(g);
    Rng rng(seed);
    // union of a group without its i-th row, for both groups of the pair
    auto unions_without = [&](int g, uint64_t* out) {
        for (int i = 0; i < 8; ++i) {
            uint64_t* o = out + (size_t)i * RW;
            memset(o, 0, sizeof(uint64_t) * RW);
            for (int j = 0; j < 8; ++j) {
                if (j == i) continue;
                const uint64_t* r = B.row(grp[g * 8 + j]);
                for (int k = 0; k < RW; ++k) o[k] |= r[k];
            }
        }
    };
    if (G >= 2) {
        for (int64_t it = 0; it < iters; ++it) {
            const int g1 = rng.below(G);
            int g2 = rng.below(G - 1);
            if (g2 >= g1) ++g2;
            unions_without(g1, without.data());
            unions_without(g2, without.data() + (size_t)8 * RW);
            const long long old_score = (long long)(cost[g1] + cost[g2]) * 4096 + usum[g1] + usum[g2];
            long long best = old_score;
            int bi = -1, bj = -1;
            for (int i = 0; i < 8; ++i) {
                if (grp[g1 * 8 + i] >= nrows) continue;             // padding rows stay at the tail of the last group
                const uint64_t* ri = B.row(grp[g1 * 8 + i]);
                const uint64_t* w1 = without.data() + (size_t)i * RW;
                for (int j = 0; j < 8; ++j) {
                    if (grp[g2 * 8 + j] >= nrows) continue;
                    const uint64_t* rj = B.row(grp[g2 * 8 + j]);
                    const uint64_t* w2 = without.data() + (size_t)(8 + j) * RW;
                    int u1, u2;
                    for (int k = 0; k < RW; ++k) cand[k] = w1[k] | rj[k];
                    const int c1 = seg_blocks(cand.data(), nseg, W, &u1);
                    for (int k = 0; k < RW; ++k) cand[k] = w2[k] | ri[k];
                    const int c2 = seg_blocks(cand.data(), nseg, W, &u2);
                    const long long score = (long long)(c1 + c2) * 4096 + u1 + u2;
                    if (score < best) { best = score; bi = i; bj = j; }
                }
            }
            if (bi >= 0) {
                std::swap(grp[g1 * 8 + bi], grp[g2 * 8 + bj]);
                refresh(g1);
                refresh(g2);
            }
        }
    }
    // Second stage: simulated annealing over single row swaps (the best-swap search above stops in a local minimum:
    // P8 tet order 2 stacked, 1650 x 165: 2826 greedy -> 2406 local search -> ~2270 here), then the best grouping seen
    // is kept.  Cost of a swap is evaluated on bit sets: union' = (union & ~(rows held once & leaving row)) | new row.
    if (G >= 32 && iters > 0) {
        // (budget grows with the matrix; skipped when the first stage is already within 12 % of the row-wise bound
        // sum_rows ceil(nnz / 4) / 8: dense matrices -- Nedelec 2nd kind, spectral Lagrange -- are packed to 85-95 %)
        int64_t sa_iters = std::min<int64_t>(iters * 100, (int64_t)6000 * nrows);
        {
            long long quarters = 0, have = 0;
            for (int r = 0; r < nrows; ++r) {
                int tot;
                quarters += seg_blocks(B.row(r), nseg, W, &tot);
            }
            for (int g = 0; g < G; ++g) have += cost[g];
            if (have * 8 * 100 <= quarters * 112) sa_iters = 0;
        }
        std::vector<uint8_t> cnt((size_t)G * ncols, 0);
        std::vector<uint64_t> ge1((size_t)G * RW, 0), eq1((size_t)G * RW, 0);
        auto bit_of = [&](int c) { const int sg = c / width, k = c % width; return std::make_pair(sg * W + (k >> 6), 1ull << (k & 63)); };
        auto rebuild = [&](int g) {
            uint8_t* cg = cnt.data() + (size_t)g * ncols;
            memset(cg, 0, ncols);
            for (int i = 0; i < 8; ++i) {
                const int r = grp[g * 8 + i];
                if (r >= nrows) continue;
                for (int c = 0; c < ncols; ++c) cg[c] += support[(size_t)r * ncols + c] ? 1 : 0;
            }
            uint64_t* a = ge1.data() + (size_t)g * RW;
            uint64_t* b = eq1.data() + (size_t)g * RW;
            memset(a, 0, sizeof(uint64_t) * RW);
            memset(b, 0, sizeof(uint64_t) * RW);
            for (int c = 0; c < ncols; ++c) {
                const auto wb = bit_of(c);
                if (cg[c] >= 1) a[wb.first] |= wb.second;
                if (cg[c] == 1) b[wb.first] |= wb.second;
            }
        };
        for (int g = 0; g < G; ++g) rebuild(g);
        std::vector<int> ucost(G), usize(G);
        long long total = 0;
        for (int g = 0; g < G; ++g) {
            ucost[g] = seg_blocks(ge1.data() + (size_t)g * RW, nseg, W, &usize[g]);
            total += ucost[g];
        }
        long long best_total = total;
        std::vector<int> best_grp = grp;
        const double T0 = 0.3;
        for (int64_t it = 0; it < sa_iters; ++it) {
            const int g1 = rng.below(G);
            int g2 = rng.below(G - 1);
            if (g2 >= g1) ++g2;
            const int i = rng.below(8), j = rng.below(8);
            const int r1 = grp[g1 * 8 + i], r2 = grp[g2 * 8 + j];
            if (r1 >= nrows || r2 >= nrows) continue;
            const uint64_t *R1 = B.row(r1), *R2 = B.row(r2);
            const uint64_t *a1 = ge1.data() + (size_t)g1 * RW, *b1 = eq1.data() + (size_t)g1 * RW;
            const uint64_t *a2 = ge1.data() + (size_t)g2 * RW, *b2 = eq1.data() + (size_t)g2 * RW;
            for (int k = 0; k < RW; ++k) cand[k] = (a1[k] & ~(b1[k] & R1[k])) | R2[k];
            int u1, u2;
            const int c1 = seg_blocks(cand.data(), nseg, W, &u1);
            for (int k = 0; k < RW; ++k) cand[k] = (a2[k] & ~(b2[k] & R2[k])) | R1[k];
            const int c2 = seg_blocks(cand.data(), nseg, W, &u2);
            const double d = (double)(c1 + c2 - ucost[g1] - ucost[g2]) + 0.05 * (double)(u1 + u2 - usize[g1] - usize[g2]);
            bool accept = d <= 0.0;
            if (!accept) {
                const double T = T0 * (1.0 - (double)it / (double)sa_iters);
                accept = T > 1e-9 && (double)(rng.next() >> 11) * (1.0 / 9007199254740992.0) < std::exp(-d / T);
            }
            if (!accept) continue;
            std::swap(grp[g1 * 8 + i], grp[g2 * 8 + j]);
            rebuild(g1);
            rebuild(g2);
            total += c1 + c2 - ucost[g1] - ucost[g2];
            ucost[g1] = c1; ucost[g2] = c2; usize[g1] = u1; usize[g2] = u2;
            if (total < best_total) { best_total = total; best_grp = grp; }
        }
        grp = best_grp;
    }
    for (int i = 0; i < nrows; ++i) order[i] = grp[i];      // padding rows never left the tail of the last group
    if (blocks_out) {
        int total = 0;
        std::vector<uint64_t> u(RW);
        for (int g = 0; g < G; ++g) {
            std::fill(u.begin(), u.end(), 0);
            for (int i = g * 8; i < std::min(nrows, g * 8 + 8); ++i) {
                const uint64_t* r = B.row(order[i]);
                for (int k = 0; k < RW; ++k) u[k] |= r[k];
            }
            int tot;
            total += seg_blocks(u.data(), nseg, W, &tot);
        }
        *blocks_out = total;
    }
    return FIATB200_OK;
}

int fiatb200_colour_members(const uint8_t* support, int32_t nrows, int32_t ncols, int32_t nseg, const int32_t* order,
                            int64_t iters, uint64_t seed, int32_t* colour_out, int32_t* conflicts_out) {
    if (!support || !order || !colour_out || nrows < 1 || ncols < 1 || nseg < 1 || ncols % nseg) return FIATB200_ERR_ARG;
    const int K = ncols / nseg;                 // members
    const int G = (nrows + 7) / 8;
    const int NS = G * nseg;                    // (group, segment) sets
    // sup[s][m]: member m is used by set s
    std::vector<uint8_t> sup((size_t)NS * K, 0);
    for (int i = 0; i < nrows; ++i) {
        const int r = order[i];
        if (r < 0 || r >= nrows) return FIATB200_ERR_ARG;
        const int g = i / 8;
        for (int c = 0; c < ncols; ++c)
            if (support[(size_t)r * ncols + c]) sup[(size_t)(g * nseg + c / K) * K + c % K] = 1;
    }
    std::vector<int> need(NS, 0);               // blocks of the set = ceil(union / 4)
    std::vector<std::vector<int>> sets_of(K);
    for (int s = 0; s < NS; ++s) {
        int n = 0;
        for (int m = 0; m < K; ++m)
            if (sup[(size_t)s * K + m]) { ++n; sets_of[m].push_back(s); }
        need[s] = (n + 3) / 4;
    }
    // initial colours: round robin by decreasing frequency
    std::vector<int> by_freq(K);
    for (int m = 0; m < K; ++m) by_freq[m] = m;
    std::stable_sort(by_freq.begin(), by_freq.end(), [&](int a, int b) { return sets_of[a].size() > sets_of[b].size(); });
    std::vector<int> col(K);
    for (int i = 0; i < K; ++i) col[by_freq[i]] = i & 3;
    int quota[4] = {0, 0, 0, 0};                // slots of each residue among 0..K-1
    for (int i = 0; i < K; ++i) ++quota[i & 3];
    int size[4] = {0, 0, 0, 0};
    for (int m = 0; m < K; ++m) ++size[col[m]];
    std::vector<int> cnt((size_t)NS * 4, 0);
    for (int s = 0; s < NS; ++s)
        for (int m = 0; m < K; ++m)
            if (sup[(size_t)s * K + m]) ++cnt[(size_t)s * 4 + col[m]];
    // members of a colour beyond the set's block count cannot be placed conflict free
    auto excess = [&](int s) {
        int e = 0;
        for (int c = 0; c < 4; ++c) e += std::max(0, cnt[(size_t)s * 4 + c] - need[s]);
        return e;
    };
    auto delta_move = [&](int m, int to) {      // change of the total excess if member m takes colour `to`
        const int from = col[m];
        int d = 0;
        for (int s : sets_of[m]) {
            int* c = &cnt[(size_t)s * 4];
            d -= std::max(0, c[from] - need[s]) + std::max(0, c[to] - need[s]);
            d += std::max(0, c[from] - 1 - need[s]) + std::max(0, c[to] + 1 - need[s]);
        }
        return d;
    };
    auto apply_move = [&](int m, int to) {
        for (int s : sets_of[m]) { --cnt[(size_t)s * 4 + col[m]]; ++cnt[(size_t)s * 4 + to]; }
        --size[col[m]]; ++size[to];
        col[m] = to;
    };
    Rng rng(seed);
    for (int64_t it = 0; it < iters && K >= 2; ++it) {
        const int a = rng.below(K);
        if (rng.next() & 1) {
            const int to = rng.below(4);
            if (to == col[a] || size[to] >= quota[to]) continue;
            if (delta_move(a, to) <= 0) apply_move(a, to);
        } else {
            const int b = rng.below(K);
            if (col[a] == col[b]) continue;
            const int ca = col[a], cb = col[b];
            const int d1 = delta_move(a, cb);
            apply_move(a, cb);
            const int d2 = delta_move(b, ca);
            if (d1 + d2 <= 0) apply_move(b, ca);
            else apply_move(a, ca);
        }
    }
    for (int m = 0; m < K; ++m) colour_out[m] = col[m];
    if (conflicts_out) {
        int tot = 0;
        for (int s = 0; s < NS; ++s) tot += excess(s);
        *conflicts_out = tot;
    }
    return FIATB200_OK;
}

}  // extern "C"
