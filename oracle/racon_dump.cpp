/*
 * racon_dump.cpp -- dumps the windows racon builds from real data, and the consensus racon's CPU path
 * computes for each of them (TEST INFRASTRUCTURE, fixture generator only).
 *
 * Compiled by oracle/Makefile (target _ref/racon_dump) together with the UNMODIFIED reference sources,
 * by path and never copied:
 *   /root/reference/src/{polisher,overlap,sequence,window,logger}.cpp
 *   /root/reference/vendor/thread_pool/src/thread_pool.cpp, vendor/edlib/edlib/src/edlib.cpp,
 *   vendor/spoa/src/*.cpp, bioparser (header only), zlib.
 * It runs racon::Polisher::initialize (src/polisher.cpp:189-457: parse, filter overlaps, edlib breaking
 * points, createWindow/add_layer) exactly as test/racon_test.cpp:88-130 sets it up, then -- instead of
 * Polisher::polish -- calls Window::generate_consensus per window itself (same engine, same trim flag,
 * src/polisher.cpp:491-504) so that the per-window consensus and status can be recorded, stitches the
 * windows like polisher.cpp:512-533 and checks the result against the reference's own golden edit
 * distance (test/racon_test.cpp:106-107: 1312, :128-129: 1566, :194-195: 1289).
 *
 * Access to racon::Window's private layers: window.hpp declares `friend class CUDABatchProcessor` under
 * CUDA_ENABLED (src/window.hpp:57-59, the hook the reference's own GPU adapter uses); this TU alone is
 * compiled with that macro around the include and defines a class of that name as the accessor.  The
 * macro adds a friend declaration only, no member: the layout every other TU sees is unchanged.
 *
 * Output (little endian, read by tests/golden/make_lambda_golden.py):
 *   u32 magic 'RWD1', u32 n_windows, u32 window_type(0 NGS,1 TGS), u32 trim, i32 m,x,g, u32 window_length,
 *   u32 edit_distance_to_reference, u32 contig_length
 *   per window: u64 id, u32 rank, u32 n_seqs, u8 polished, u32 cons_len, cons bytes,
 *     per sequence: u32 len, u32 begin, u32 end, u8 has_quality, bases[len], (quality[len])
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#define CUDA_ENABLED
#include "window.hpp"
#undef CUDA_ENABLED
#include "overlap.hpp"
#include "polisher.hpp"
#include "sequence.hpp"

#include "bioparser/bioparser.hpp"
#include "edlib.h"
#include "spoa/spoa.hpp"
#include "thread_pool/thread_pool.hpp"

namespace racon {

class CUDABatchProcessor { /* accessor only; see the header comment */
public:
    static const std::vector<std::pair<const char*, uint32_t>>& sequences(const Window& w) { return w.sequences_; }
    static const std::vector<std::pair<const char*, uint32_t>>& qualities(const Window& w) { return w.qualities_; }
    static const std::vector<std::pair<uint32_t, uint32_t>>& positions(const Window& w) { return w.positions_; }
    static WindowType type(const Window& w) { return w.type_; }
};

class DumpPolisher : public Polisher {
public:
    DumpPolisher(const std::string& reads, const std::string& overlaps, const std::string& target, bool fastq,
                 uint32_t window_length, double quality_threshold, double error_threshold, bool trim, int8_t m,
                 int8_t x, int8_t g)
        : Polisher(fastq ? bioparser::createParser<bioparser::FastqParser, Sequence>(reads)
                         : bioparser::createParser<bioparser::FastaParser, Sequence>(reads),
                   bioparser::createParser<bioparser::PafParser, Overlap>(overlaps),
                   bioparser::createParser<bioparser::FastaParser, Sequence>(target), PolisherType::kC, window_length,
                   quality_threshold, error_threshold, trim, m, x, g, 1) {}

    int dump(const char* out_path, const std::string& reference_path, uint32_t window_length, bool trim, int m, int x,
             int g) {
        FILE* f = std::fopen(out_path, "wb");
        if (!f) return 1;
        auto u32 = [&](uint32_t v) { std::fwrite(&v, 4, 1, f); };
        auto u64 = [&](uint64_t v) { std::fwrite(&v, 8, 1, f); };
        auto u8 = [&](uint8_t v) { std::fwrite(&v, 1, 1, f); };
        /* consensus per window first (Window keeps pointers into sequences_, alive until polish()) */
        std::vector<std::string> cons(windows_.size());
        std::vector<uint8_t> ok(windows_.size());
        std::string contig;
        for (size_t i = 0; i < windows_.size(); ++i) {
            ok[i] = windows_[i]->generate_consensus(alignment_engines_[0], trim_) ? 1 : 0;
            cons[i] = windows_[i]->consensus();
            contig += cons[i]; /* one target in the sample data: polisher.cpp:512-533 concatenates per target */
        }
        /* test/racon_test.cpp:98-107: reverse complement of the polished contig vs the reference */
        std::vector<std::unique_ptr<Sequence>> seqs;
        seqs.emplace_back(createSequence("polished", contig));
        seqs[0]->create_reverse_complement();
        auto parser = bioparser::createParser<bioparser::FastaParser, Sequence>(reference_path);
        parser->parse(seqs, -1);
        const std::string& q = seqs[0]->reverse_complement();
        const std::string& t = seqs[1]->data();
        EdlibAlignResult r = edlibAlign(q.c_str(), q.size(), t.c_str(), t.size(), edlibDefaultAlignConfig());
        const uint32_t ed = static_cast<uint32_t>(r.editDistance);
        edlibFreeAlignResult(r);

        u32(0x31445752u);
        u32(static_cast<uint32_t>(windows_.size()));
        u32(windows_.empty() ? 1u : (CUDABatchProcessor::type(*windows_[0]) == WindowType::kTGS ? 1u : 0u));
        u32(trim ? 1u : 0u);
        u32(static_cast<uint32_t>(m));
        u32(static_cast<uint32_t>(x));
        u32(static_cast<uint32_t>(g));
        u32(window_length);
        u32(ed);
        u32(static_cast<uint32_t>(contig.size()));
        for (size_t i = 0; i < windows_.size(); ++i) {
            const Window& w = *windows_[i];
            const auto& s = CUDABatchProcessor::sequences(w);
            const auto& ql = CUDABatchProcessor::qualities(w);
            const auto& p = CUDABatchProcessor::positions(w);
            u64(w.id());
            u32(w.rank());
            u32(static_cast<uint32_t>(s.size()));
            u8(ok[i]);
            u32(static_cast<uint32_t>(cons[i].size()));
            std::fwrite(cons[i].data(), 1, cons[i].size(), f);
            for (size_t k = 0; k < s.size(); ++k) {
                u32(s[k].second);
                u32(p[k].first);
                u32(p[k].second);
                const bool hq = ql[k].first != nullptr && ql[k].second != 0;
                u8(hq ? 1 : 0);
                std::fwrite(s[k].first, 1, s[k].second, f);
                if (hq) std::fwrite(ql[k].first, 1, ql[k].second, f);
            }
        }
        std::fclose(f);
        std::fprintf(stderr, "[racon_dump] %zu windows, contig %zu bp, edit distance to reference %u\n", windows_.size(),
                     contig.size(), ed);
        return 0;
    }
};

} // namespace racon

int main(int argc, char** argv) {
    if (argc != 12) {
        std::fprintf(stderr, "usage: racon_dump reads overlaps.paf layout.fasta reference.fasta window_length "
                             "quality_threshold error_threshold m x g out.bin\n");
        return 2;
    }
    const std::string reads = argv[1];
    const bool fastq = reads.find(".fastq") != std::string::npos || reads.find(".fq") != std::string::npos;
    const uint32_t wl = static_cast<uint32_t>(std::atoi(argv[5]));
    const double qt = std::atof(argv[6]), et = std::atof(argv[7]);
    const int m = std::atoi(argv[8]), x = std::atoi(argv[9]), g = std::atoi(argv[10]);
    racon::DumpPolisher p(reads, argv[2], argv[3], fastq, wl, qt, et, true, static_cast<int8_t>(m), static_cast<int8_t>(x),
                          static_cast<int8_t>(g));
    p.initialize();
    return p.dump(argv[11], argv[4], wl, true, m, x, g);
}
