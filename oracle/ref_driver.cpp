/*
 * ref_driver.cpp -- thin C-ABI driver around the UNMODIFIED reference sources (TEST INFRASTRUCTURE).
 *
 * Compiled by oracle/Makefile together with, by path and never copied,
 *   /root/reference/src/window.cpp
 *   /root/reference/vendor/spoa/src/{alignment_engine,graph,simd_alignment_engine,sisd_alignment_engine}.cpp
 *   /root/reference/vendor/edlib/edlib/src/edlib.cpp
 * into oracle/_ref/libracon_ref.so.  It drives racon::Window exactly the way
 * racon::Polisher::polish does (src/polisher.cpp:181-185, 491-504): one spoa NW engine per
 * thread, prealloc(window_length, 5), one generate_consensus call per window.
 *
 * Used by tests (oracle validation) and by bench.py --impl reference / cpu_baseline.
 * The product never loads it.
 */
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "edlib.h"
#include "spoa/spoa.hpp"
#include "window.hpp"
#include "sequence.hpp"
#include "overlap.hpp"

namespace {

struct FlatBatch {
    int64_t n_windows;
    const int64_t* win_seq_off; /* [n_windows+1] first sequence index of each window */
    const int64_t* seq_off;     /* [n_seqs+1] byte offset of each sequence in bases/weights */
    const uint8_t* bases;
    const int8_t* weights;      /* same offsets as bases */
    const uint8_t* has_weights; /* [n_seqs] 0 => no quality (weight 1) */
    const int32_t* begins;      /* [n_seqs] */
    const int32_t* ends;        /* [n_seqs] */
};

/* Build a racon::Window for window w.  Quality strings are rebuilt as PHRED+33 chars and kept
 * alive in `store` because racon::Window keeps raw pointers (src/window.cpp:34-36,60-62). */
std::shared_ptr<racon::Window> make_window(const FlatBatch& b, int64_t w, bool tgs,
                                           std::vector<std::string>& store) {
    const int64_t s0 = b.win_seq_off[w], s1 = b.win_seq_off[w + 1];
    store.clear();
    store.reserve(static_cast<size_t>(s1 - s0));
    for (int64_t s = s0; s < s1; ++s) {
        const int64_t len = b.seq_off[s + 1] - b.seq_off[s];
        std::string q;
        if (b.has_weights[s]) {
            q.resize(static_cast<size_t>(len));
            for (int64_t k = 0; k < len; ++k)
                q[static_cast<size_t>(k)] = static_cast<char>(b.weights[b.seq_off[s] + k] + 33);
        } else if (s == s0) {
            /* backbone without quality: racon uses the dummy '!' string (polisher.cpp:171,392-395)
             * only when the caller says so; here "no weights" on the backbone means weight 1. */
            q.assign(static_cast<size_t>(len), static_cast<char>(1 + 33));
        }
        store.push_back(std::move(q));
    }
    const char* bb = reinterpret_cast<const char*>(b.bases + b.seq_off[s0]);
    const uint32_t bl = static_cast<uint32_t>(b.seq_off[s0 + 1] - b.seq_off[s0]);
    auto win = racon::createWindow(static_cast<uint64_t>(w), 0,
                                   tgs ? racon::WindowType::kTGS : racon::WindowType::kNGS, bb, bl,
                                   store[0].data(), bl);
    for (int64_t s = s0 + 1; s < s1; ++s) {
        const uint32_t len = static_cast<uint32_t>(b.seq_off[s + 1] - b.seq_off[s]);
        const std::string& q = store[static_cast<size_t>(s - s0)];
        win->add_layer(reinterpret_cast<const char*>(b.bases + b.seq_off[s]), len,
                       b.has_weights[s] ? q.data() : nullptr, b.has_weights[s] ? len : 0,
                       static_cast<uint32_t>(b.begins[s]), static_cast<uint32_t>(b.ends[s]));
    }
    return win;
}

} // namespace

extern "C" {

/* The order racon processes a window's sequences in: window.cpp:78-85 (same call as
 * cudabatch.cpp:96-104).  rank_out[0] == 0. */
void ref_layer_order(int32_t n, const int32_t* begins, int32_t* rank_out) {
    std::vector<uint32_t> rank;
    rank.reserve(static_cast<size_t>(n));
    for (int32_t i = 0; i < n; ++i) rank.emplace_back(static_cast<uint32_t>(i));
    std::sort(rank.begin() + 1, rank.end(),
              [&](uint32_t lhs, uint32_t rhs) { return begins[lhs] < begins[rhs]; });
    for (int32_t i = 0; i < n; ++i) rank_out[i] = static_cast<int32_t>(rank[static_cast<size_t>(i)]);
}

/*
 * Polish every window of a flat batch with racon::Window::generate_consensus using
 * n_threads host threads (dynamic window cursor, one engine per thread).
 * Sequences must be given in ADD order (window.cpp sorts them itself).
 * cons_out: n_windows rows of `stride` bytes; cons_len/polished: [n_windows].
 */
void ref_polish_windows(int64_t n_windows, const int64_t* win_seq_off, const int64_t* seq_off,
                        const uint8_t* bases, const int8_t* weights, const uint8_t* has_weights,
                        const int32_t* begins, const int32_t* ends, int32_t tgs, int32_t trim,
                        int32_t m, int32_t x, int32_t g, int32_t window_length, int32_t n_threads,
                        char* cons_out, int32_t stride, int32_t* cons_len, uint8_t* polished) {
    FlatBatch b{n_windows, win_seq_off, seq_off, bases, weights, has_weights, begins, ends};
    if (n_threads < 1) n_threads = 1;
    std::atomic<int64_t> cursor{0};
    auto worker = [&]() {
        std::shared_ptr<spoa::AlignmentEngine> engine = spoa::createAlignmentEngine(
            spoa::AlignmentType::kNW, static_cast<int8_t>(m), static_cast<int8_t>(x),
            static_cast<int8_t>(g));
        engine->prealloc(static_cast<uint32_t>(window_length), 5);
        std::vector<std::string> store;
        while (true) {
            const int64_t w = cursor.fetch_add(1);
            if (w >= n_windows) break;
            auto win = make_window(b, w, tgs != 0, store);
            const bool ok = win->generate_consensus(engine, trim != 0);
            const std::string& c = win->consensus();
            const int32_t len = static_cast<int32_t>(std::min<size_t>(c.size(), static_cast<size_t>(stride)));
            std::memcpy(cons_out + w * static_cast<int64_t>(stride), c.data(), static_cast<size_t>(len));
            cons_len[w] = static_cast<int32_t>(c.size());
            polished[w] = ok ? 1 : 0;
        }
    };
    if (n_threads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int32_t t = 0; t < n_threads; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
}

/*
 * Same spoa call sequence as window.cpp:73-116 for ONE full-span window, but returning the
 * untrimmed consensus AND spoa's per-base coverages (window.cpp keeps them private), plus the
 * final rank_to_node order, so coverage / topological order can be pinned too.
 * Sequences are given in PROCESSING order.  Returns consensus length.
 */
int32_t ref_spoa_window(int32_t n_seqs, const char* const* seqs, const int32_t* lens,
                        const int8_t* const* weights, int32_t m, int32_t x, int32_t g,
                        char* cons_out, uint32_t* cov_out, int32_t max_out, int32_t* rank_out,
                        int32_t max_nodes, int32_t* n_nodes_out) {
    auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, static_cast<int8_t>(m),
                                              static_cast<int8_t>(x), static_cast<int8_t>(g));
    auto graph = spoa::createGraph();
    for (int32_t i = 0; i < n_seqs; ++i) {
        spoa::Alignment aln;
        if (i > 0) aln = engine->align(seqs[i], static_cast<uint32_t>(lens[i]), graph);
        std::vector<uint32_t> w(static_cast<size_t>(lens[i]), 1u);
        if (weights[i])
            for (int32_t k = 0; k < lens[i]; ++k) w[static_cast<size_t>(k)] = static_cast<uint32_t>(weights[i][k]);
        graph->add_alignment(aln, seqs[i], static_cast<uint32_t>(lens[i]), w);
    }
    std::vector<uint32_t> cov;
    std::string cons = graph->generate_consensus(cov);
    const int32_t len = static_cast<int32_t>(cons.size());
    if (len <= max_out) {
        std::memcpy(cons_out, cons.data(), cons.size());
        if (cov_out) std::memcpy(cov_out, cov.data(), sizeof(uint32_t) * cov.size());
    }
    const auto& r2n = graph->rank_to_node_id();
    if (n_nodes_out) *n_nodes_out = static_cast<int32_t>(r2n.size());
    if (rank_out && static_cast<int32_t>(r2n.size()) <= max_nodes)
        for (size_t i = 0; i < r2n.size(); ++i) rank_out[i] = static_cast<int32_t>(r2n[i]);
    return len;
}

/*
 * spoa's multiple sequence alignment of one group of sequences, the way the reference's own MSA test produces its
 * expected value (vendor/GenomeWorks/cudapoa/tests/Test_CudapoaGenerateMSA2.cu:62-79: kNW engine, align + add_alignment
 * per sequence, Graph::generate_multiple_sequence_alignment).  Rows are written back to back into `out`
 * (n_seqs x *msa_len bytes) when they fit max_out.  Returns the number of rows.
 */
int32_t ref_spoa_window_msa(int32_t n_seqs, const char* const* seqs, const int32_t* lens,
                            const int8_t* const* weights, int32_t m, int32_t x, int32_t g, char* out,
                            int64_t max_out, int32_t* msa_len) {
    auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, static_cast<int8_t>(m),
                                              static_cast<int8_t>(x), static_cast<int8_t>(g));
    auto graph = spoa::createGraph();
    for (int32_t i = 0; i < n_seqs; ++i) {
        spoa::Alignment aln;
        if (i > 0) aln = engine->align(seqs[i], static_cast<uint32_t>(lens[i]), graph);
        std::vector<uint32_t> w(static_cast<size_t>(lens[i]), 1u);
        if (weights && weights[i])
            for (int32_t k = 0; k < lens[i]; ++k) w[static_cast<size_t>(k)] = static_cast<uint32_t>(weights[i][k]);
        graph->add_alignment(aln, seqs[i], static_cast<uint32_t>(lens[i]), w);
    }
    std::vector<std::string> msa;
    graph->generate_multiple_sequence_alignment(msa);
    const size_t L = msa.empty() ? 0 : msa[0].size();
    if (msa_len) *msa_len = static_cast<int32_t>(L);
    if (static_cast<int64_t>(L * msa.size()) <= max_out)
        for (size_t i = 0; i < msa.size(); ++i) std::memcpy(out + i * L, msa[i].data(), L);
    return static_cast<int32_t>(msa.size());
}

/*
 * The overlap alignment of racon's CPU path, verbatim: Overlap::align_overlaps (src/overlap.cpp:205-224) =
 * edlibAlign(q, q_length, t, t_length, edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, nullptr, 0)) followed by
 * edlibAlignmentToCigar(..., EDLIB_CIGAR_STANDARD).  ops_out (nullable, capacity cap) receives edlib's edit operations
 * (0 match, 1 insertion = a query base without partner, 2 deletion = a target base without partner, 3 mismatch),
 * cigar_out (nullable, capacity cigar_cap incl. NUL) the CIGAR string racon stores.  Returns the number of operations,
 * or -1 when edlib fails / a buffer is too small; *score receives the edit distance.
 */
int64_t ref_edlib_nw(const char* q, int32_t q_len, const char* t, int32_t t_len, uint8_t* ops_out, int64_t cap,
                     char* cigar_out, int64_t cigar_cap, int32_t* score) {
    EdlibAlignResult r = edlibAlign(q, q_len, t, t_len, edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, nullptr, 0));
    int64_t n = -1;
    if (r.status == EDLIB_STATUS_OK) {
        if (score) *score = r.editDistance;
        n = r.alignmentLength;
        if (ops_out) {
            if (n <= cap) std::memcpy(ops_out, r.alignment, static_cast<size_t>(n));
            else n = -1;
        }
        if (cigar_out && n >= 0) {
            char* c = edlibAlignmentToCigar(r.alignment, r.alignmentLength, EDLIB_CIGAR_STANDARD);
            const size_t L = std::strlen(c);
            if (static_cast<int64_t>(L) + 1 <= cigar_cap) std::memcpy(cigar_out, c, L + 1);
            else n = -1;
            std::free(c);
        }
    }
    edlibFreeAlignResult(r);
    return n;
}

} // extern "C"


/*
 * Breaking points of one overlap exactly as racon derives them from a CIGAR: the UNMODIFIED
 * racon::Overlap::find_breaking_points -> find_breaking_points_from_cigar (src/overlap.cpp:179-203, 226-290).
 * Overlap's constructors and fields are private; its header befriends bioparser::PafParser<Overlap> (overlap.hpp:76), so
 * this driver supplies that (otherwise unused here) specialisation to build one PAF overlap, hand it the CIGAR and mark it
 * transmuted -- no reference source is touched.  out receives (t, q) pairs, two per window; returns their number or -1.
 */
namespace bioparser {
template <>
class PafParser<racon::Overlap> {
public:
    static int64_t breaking_points(const char* cigar, uint32_t q_length, uint32_t q_begin, uint32_t q_end, int strand,
                                   uint32_t t_length, uint32_t t_begin, uint32_t t_end, uint32_t window_length,
                                   uint32_t* out, int64_t cap) {
        std::unique_ptr<racon::Overlap> o(new racon::Overlap("q", 1, q_length, q_begin, q_end, strand ? '-' : '+', "t", 1,
                                                             t_length, t_begin, t_end, 0, 0, 255));
        o->cigar_ = cigar;
        o->is_transmuted_ = true;
        std::vector<std::unique_ptr<racon::Sequence>> none;
        o->find_breaking_points(none, window_length);
        const auto& bp = o->breaking_points();
        if ((int64_t)bp.size() > cap) return -1;
        for (size_t k = 0; k < bp.size(); ++k) {
            out[2 * k] = bp[k].first;
            out[2 * k + 1] = bp[k].second;
        }
        return (int64_t)bp.size();
    }
};
} // namespace bioparser

extern "C" int64_t ref_racon_breaking_points(const char* cigar, uint32_t q_length, uint32_t q_begin, uint32_t q_end, int strand,
                                             uint32_t t_length, uint32_t t_begin, uint32_t t_end, uint32_t window_length,
                                             uint32_t* out, int64_t cap) {
    return bioparser::PafParser<racon::Overlap>::breaking_points(cigar, q_length, q_begin, q_end, strand, t_length, t_begin,
                                                                 t_end, window_length, out, cap);
}
