/*
 * emu_aln.cpp -- lock-step CPU emulation of the overlap aligner (TEST INFRASTRUCTURE).
 *
 * Compiles racon_gpu_b200/csrc/aln_core.cuh in its host flavour (poa_simt.cuh: POA_LANES loops over 32 lanes, shuffles
 * are plain loops) and drives it level by level exactly like the batch runtime (csrc/b200aln.cu) does on the device --
 * aln_push files the children, the leaf list is traced back at the end, aln_runs forms the runs -- so that the wavefront
 * bit-vector passes, the split rule, the leaf records, the traceback and the run formation are checked against the
 * oracle / the unmodified edlib without a GPU.  Never linked into the product library.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../racon_gpu_b200/csrc/host/aln_levels.hpp"

using namespace b200aln;

extern "C" {

/* Aligns one pair; ops_out must hold n + m bytes.  Returns the number of operations (holes removed); -1 inconsistent
 * split, -2 list overflow, -3 the runs do not spell the operations.  *score the edit distance; *levels (nullable) the
 * depth of the recursion; *n_leaves likewise; cigar_out (nullable, cigar_cap bytes) the CIGAR formed from the runs;
 * bp_out (nullable, 4 words per window of the target segment) the breaking points for window_length;
 * guess (< 0: none) a guessed bound on the edit distance: the top sub-problem runs banded and is redone without a band when
 * its optimum turns out larger (what the kernels do with the host's guess). */
int64_t emu_align(const uint8_t* q, int32_t n, const uint8_t* t, int32_t m, uint8_t* ops_out, int32_t* score,
                  int32_t* levels, int32_t* n_leaves, char* cigar_out, int64_t cigar_cap, int32_t q_first, int32_t t_begin,
                  int32_t window_length, uint32_t* bp_out, int32_t* bp_count, int32_t guess) {
    const int32_t max_len = (n > m ? n : m) + 1;
    size_t slot_bytes = 0;
    AlnSlot s;
    EqTab eq;
    aln_slot_bind(s, nullptr, max_len, &slot_bytes);
    std::vector<uint8_t> slab(slot_bytes + 512);
    aln_slot_bind(s, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(slab.data()) + 255) & ~uintptr_t(255)), max_len, nullptr);
    std::vector<uint8_t> ops((size_t)n + (size_t)m, OP_NONE);
    const int32_t cap_open = (int32_t)aln_open_capacity(n, m), cap_leaves = (int32_t)aln_leaf_capacity(n, m);
    std::vector<AlnRect> level[ALN_CLASSES], next[ALN_CLASSES], leaves((size_t)cap_leaves);
    int32_t n_level[ALN_CLASSES] = {0, 0, 0}, n_next[ALN_CLASSES] = {0, 0, 0}, n_leaf = 0, overflow = 0, depth = 0;
    AlnLists first, L;
    for (int c = 0; c < ALN_CLASSES; ++c) {
        level[c].resize((size_t)cap_open);
        next[c].resize((size_t)cap_open);
    }
    auto bind = [&](AlnLists& x, std::vector<AlnRect>* lists, int32_t* counts) {
        for (int c = 0; c < ALN_CLASSES; ++c) x.open[c] = lists[c].data();
        x.n_open = counts;
        x.cap_open = cap_open;
        x.leaves = leaves.data();
        x.n_leaves = &n_leaf;
        x.cap_leaves = cap_leaves;
        x.overflow = &overflow;
    };
    bind(first, level, n_level);
    aln_push(first, AlnRect{0, 0, n, 0, m, guess >= 0 ? (ALN_TOP | ALN_GUESS) : ALN_TOP, guess >= 0 ? guess : -1});
    while (n_level[0] + n_level[1] + n_level[2] > 0) {
        n_next[0] = n_next[1] = n_next[2] = 0;
        bind(L, next, n_next);
        for (int c = 0; c < ALN_CLASSES; ++c) { /* the emulation runs tall and huge sub-problems on one warp as well */
            for (int32_t k = 0; k < n_level[c]; ++k) {
                const AlnRect r = level[c][(size_t)k];
                AlnSplit sp;
                aln_split(s, eq, q + r.r0, t + r.c0, r.n, r.m, aln_band_of(r.best), &sp);
                if ((r.top & ALN_GUESS) && sp.best > r.best) /* the guess was too small: nothing of that pass can be trusted */
                    aln_split(s, eq, q + r.r0, t + r.c0, r.n, r.m, -1, &sp);
                if (r.top & ALN_TOP) *score = sp.best;
                AlnRect ul, lr;
                if (!aln_children(r, sp, ul, lr)) return -1;
                aln_push(L, ul);
                aln_push(L, lr);
            }
        }
        for (int c = 0; c < ALN_CLASSES; ++c) {
            level[c].swap(next[c]);
            n_level[c] = n_next[c];
        }
        ++depth;
    }
    if (overflow) return -2;
    for (int32_t k = 0; k < n_leaf; ++k) {
        const AlnRect r = leaves[(size_t)k];
        aln_leaf(s, eq, q + r.r0, t + r.c0, r.n, r.m, (r.top & ALN_GUESS) ? -1 : aln_band_of(r.best), ops.data() + r.r0 + r.c0,
                 (r.top & ALN_TOP) ? score : nullptr);
    }
    int64_t k = 0;
    for (uint8_t op : ops)
        if (op != OP_NONE) ops_out[k++] = op;
    /* the runs must spell the same operations */
    int32_t n_ops = -1;
    const int32_t n_runs = aln_runs(ops.data(), n + m, nullptr, n_ops);
    std::vector<uint32_t> runs((size_t)n_runs + 1);
    if (aln_runs(ops.data(), n + m, runs.data(), n_ops) != n_runs || n_ops != k) return -3;
    std::vector<uint8_t> spelled;
    aln_expand_runs(runs.data(), n_runs, n_ops, spelled);
    if (spelled.size() != (size_t)k || (k > 0 && std::memcmp(spelled.data(), ops_out, (size_t)k) != 0)) return -3;
    if (cigar_out) { /* the engine's own text (aln_cigar_text) must equal the host-side spelling of the runs */
        const std::string c = aln_runs_to_cigar(runs.data(), n_runs, n_ops);
        const int32_t bytes = aln_cigar_text(runs.data(), n_runs, n_ops, nullptr);
        if ((int64_t)bytes + 1 > cigar_cap || bytes != (int32_t)c.size()) return -3;
        if (aln_cigar_text(runs.data(), n_runs, n_ops, reinterpret_cast<uint8_t*>(cigar_out)) != bytes) return -3;
        cigar_out[bytes] = 0;
        if (c != cigar_out) return -3;
    }
    if (bp_out && window_length > 0) { /* breaking points from the same run starts */
        std::vector<int32_t> pre(2 * ((size_t)n_runs + 1) + 2);
        *bp_count = aln_breaking_points(runs.data(), n_runs, n_ops, q_first, t_begin, m, window_length, pre.data(), bp_out);
    }
    if (levels) *levels = depth;
    if (n_leaves) *n_leaves = n_leaf;
    return k;
}

} // extern "C"
