/*
 * aln_levels.hpp -- the host side of the level-synchronous Hirschberg recursion (aln_core.cuh): which open sub-problems
 * are leaves (edlib.cpp:1135-1157), and the two children of a split one (edlib.cpp:1321-1333).  Shared by the batch
 * runtime (csrc/b200aln.cu) and the lane emulation (tests/emu/emu_aln.cpp).
 */
#pragma once
#include <algorithm>
#include <vector>

#include "../aln_core.cuh"

namespace b200aln {

/* leaves go to `leaves`, the rest to `open`, largest first (a persistent grid drains the expensive ones first) */
inline void aln_classify(const std::vector<AlnRect>& level, std::vector<AlnRect>& open, std::vector<AlnRect>& leaves) {
    open.clear();
    for (const AlnRect& r : level) (aln_is_leaf(r.n, r.m) ? leaves : open).push_back(r);
    std::stable_sort(open.begin(), open.end(), [](const AlnRect& a, const AlnRect& b) {
        return (int64_t)a.n * a.m > (int64_t)b.n * b.m;
    });
}

/* upper-left and lower-right sub-problems of rect `r` split at query index s.r (relative, -1 .. n-1) */
inline bool aln_children(const AlnRect& r, const AlnSplit& s, AlnRect& ul, AlnRect& lr) {
    if (s.r < -1 || s.r > r.n - 1) return false;
    const int32_t lh = r.m / 2, uh = s.r + 1;
    ul = AlnRect{r.aln, r.r0, uh, r.c0, lh, 0};
    lr = AlnRect{r.aln, r.r0 + uh, r.n - uh, r.c0 + lh, r.m - lh, 0};
    return true;
}

} // namespace b200aln
