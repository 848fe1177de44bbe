/*
 * emu_aln.cpp -- lock-step CPU emulation of the overlap aligner (TEST INFRASTRUCTURE).
 *
 * Compiles racon_gpu_b200/csrc/aln_core.cuh in its host flavour (poa_simt.cuh: POA_LANES loops over 32 lanes, shuffles
 * are plain loops) and drives it with the product's own level logic (csrc/host/aln_levels.hpp), so that the wavefront
 * bit-vector passes, the split rule, the leaf records and the traceback are checked against the oracle / the unmodified
 * edlib without a GPU.  Never linked into the product library.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../racon_gpu_b200/csrc/host/aln_levels.hpp"

using namespace b200aln;

extern "C" {

/* Aligns one pair; ops_out must hold n + m bytes.  Returns the number of operations (holes removed), -1 on an
 * inconsistent split; *score the edit distance; *levels (nullable) the depth of the recursion; *n_leaves likewise. */
int64_t emu_align(const uint8_t* q, int32_t n, const uint8_t* t, int32_t m, uint8_t* ops_out, int32_t* score,
                  int32_t* levels, int32_t* n_leaves) {
    const int32_t max_len = (n > m ? n : m) + 1;
    size_t slot_bytes = 0;
    AlnSlot s;
    aln_slot_bind(s, nullptr, max_len, &slot_bytes);
    std::vector<uint8_t> slab(slot_bytes + 512);
    aln_slot_bind(s, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(slab.data()) + 255) & ~uintptr_t(255)), max_len, nullptr);
    std::vector<uint8_t> ops((size_t)n + (size_t)m, OP_NONE);
    std::vector<AlnRect> level{AlnRect{0, 0, n, 0, m, 1}}, open, leaves, next;
    int32_t depth = 0;
    while (!level.empty()) {
        aln_classify(level, open, leaves);
        next.clear();
        for (const AlnRect& r : open) {
            AlnSplit sp;
            aln_split(s, q + r.r0, t + r.c0, r.n, r.m, &sp);
            if (r.top) *score = sp.best;
            AlnRect ul, lr;
            if (!aln_children(r, sp, ul, lr)) return -1;
            next.push_back(ul);
            next.push_back(lr);
        }
        level.swap(next);
        if (!open.empty()) ++depth;
    }
    for (const AlnRect& r : leaves)
        aln_leaf(s, q + r.r0, t + r.c0, r.n, r.m, ops.data() + r.r0 + r.c0, r.top ? score : nullptr);
    int64_t k = 0;
    for (uint8_t op : ops)
        if (op != OP_NONE) ops_out[k++] = op;
    if (levels) *levels = depth;
    if (n_leaves) *n_leaves = (int32_t)leaves.size();
    return k;
}

} // extern "C"
