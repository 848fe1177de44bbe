/*
 * b200aln_aligner.hpp -- header-only C++ shim with the method set of claraparabricks::genomeworks::cudaaligner::
 * {Aligner, Alignment, create_aligner, StatusType, AlignmentType, AlignmentState}
 * (/root/reference/vendor/GenomeWorks/cudaaligner/include/claraparabricks/genomeworks/cudaaligner/
 * aligner.hpp:43-132, alignment.hpp:55-105, cudaaligner.hpp:34-58) over the C ABI (include/b200aln.h), so that the
 * body of racon's src/cuda/cudaaligner.cpp builds against it with only its includes / usings changed (INTEGRATION.md
 * section 5; tests/test_boundary.py compiles exactly that).
 *
 * Argument order: cudaaligner's (query, target) are racon's (target, query) -- racon's adapter swaps them at its call
 * site (cudaaligner.cpp:60-63) and reads the CIGAR back in racon's sense.  The shim keeps that call site intact: it
 * hands the pair to the C ABI in edlib's order (read segment = query, contig segment = target), so 'I' is a read
 * character alone and 'D' a contig character alone, exactly what Overlap::find_breaking_points_from_cigar expects from
 * the CPU path (src/overlap.cpp:226-290).  AlignmentState::insertion ("absent in query, present in target",
 * cudaaligner.hpp:56) is therefore edlib's EDLIB_EDOP_INSERT, deletion its EDLIB_EDOP_DELETE.
 */
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200aln.h"

namespace b200aln_cpp {

enum StatusType { /* cudaaligner.hpp:34-42 */
    success = 0,
    uninitialized,
    exceeded_max_alignments,
    exceeded_max_length,
    exceeded_max_alignment_difference,
    generic_error
};
enum AlignmentType { global_alignment = 0, unset };                         /* cudaaligner.hpp:45-49 */
enum AlignmentState : int8_t { match = 0, mismatch, insertion, deletion };   /* cudaaligner.hpp:52-58 */

inline StatusType to_status(int32_t st) {
    return st >= 0 && st <= static_cast<int32_t>(generic_error) ? static_cast<StatusType>(st) : generic_error;
}
inline StatusType Init() { return to_status(b200aln_init()); } /* cudaaligner.hpp:61 */

class Alignment { /* alignment.hpp:55-105 */
public:
    Alignment(const char* query, int32_t query_length, const char* target, int32_t target_length)
        : query_(query, query + query_length), target_(target, target + target_length) {}
    const std::string& get_query_sequence() const { return query_; }
    const std::string& get_target_sequence() const { return target_; }
    std::string convert_to_cigar() const { return cigar_; }
    AlignmentType get_alignment_type() const { return global_alignment; }
    bool is_optimal() const { return status_ == success; } /* no lossy band: an alignment that exists is optimal */
    StatusType get_status() const { return status_; }
    int32_t get_edit_distance() const { return edit_distance_; }
    /* one state per position; expanded from the batch's run starts on first use (racon only reads the CIGAR, so the
     * run starts stay on the device unless somebody asks), available until the Aligner is reset */
    const std::vector<AlignmentState>& get_alignment() const {
        if (!expanded_ && batch_ && status_ == success) {
            const uint32_t* runs = nullptr;
            int32_t n_runs = 0, n_ops = 0, ast = B200ALN_GENERIC_ERROR;
            if (b200aln_batch_get_alignment(batch_, index_, &runs, &n_runs, &n_ops, nullptr, &ast) == B200ALN_SUCCESS &&
                ast == B200ALN_SUCCESS) {
                static const AlignmentState state[4] = {match, insertion, deletion, mismatch}; /* b200aln_op -> state */
                alignment_.reserve(static_cast<size_t>(n_ops));
                for (int32_t r = 0; r < n_runs; ++r) {
                    const int32_t b = static_cast<int32_t>(runs[r] >> 2);
                    const int32_t e = r + 1 < n_runs ? static_cast<int32_t>(runs[r + 1] >> 2) : n_ops;
                    alignment_.insert(alignment_.end(), static_cast<size_t>(e - b), state[runs[r] & 3u]);
                }
            }
            expanded_ = true;
        }
        return alignment_;
    }

private:
    friend class Aligner;
    std::string query_, target_, cigar_;
    mutable std::vector<AlignmentState> alignment_;
    mutable bool expanded_ = false;
    StatusType status_ = uninitialized;
    int32_t edit_distance_ = -1;
    const b200aln_batch* batch_ = nullptr;
    int32_t index_ = 0;
};

class Aligner { /* aligner.hpp:43-83 */
public:
    Aligner(int32_t max_bandwidth, void* stream, int32_t device_id, int64_t max_device_memory) {
        const int32_t st = b200aln_batch_create(device_id, stream, max_device_memory, max_bandwidth, &b_);
        if (st == B200ALN_INVALID_ARGUMENT) throw std::invalid_argument("create_aligner: invalid device or memory budget");
        if (st != B200ALN_SUCCESS) throw std::runtime_error(std::string("create_aligner: ") + b200aln_status_string(st));
    }
    ~Aligner() {
        detach();
        b200aln_batch_destroy(b_);
    }
    Aligner(const Aligner&) = delete;
    Aligner& operator=(const Aligner&) = delete;

    StatusType add_alignment(const char* query, int32_t query_length, const char* target, int32_t target_length,
                             bool reverse_complement_query = false, bool reverse_complement_target = false) {
        if (reverse_complement_query || reverse_complement_target)
            throw std::invalid_argument("add_alignment: reverse complement flags are not provided (racon passes none)");
        if (query_length < 0 || target_length < 0) throw std::invalid_argument("add_alignment: negative length");
        /* edlib's order below the ABI: (read segment, contig segment) = (cudaaligner target, cudaaligner query) */
        const int32_t st = b200aln_batch_add_alignment(b_, target, target_length, query, query_length);
        if (st == B200ALN_SUCCESS)
            alignments_.push_back(std::make_shared<Alignment>(query, query_length, target, target_length));
        return to_status(st);
    }
    StatusType align_all() { return to_status(b200aln_batch_align_all(b_)); }
    StatusType sync_alignments() {
        const int32_t st = b200aln_batch_sync(b_);
        if (st != B200ALN_SUCCESS) return to_status(st);
        const char* text = nullptr;
        const int64_t* off = nullptr;
        const int32_t *len = nullptr, *ed = nullptr, *ast = nullptr;
        if (b200aln_batch_get_cigars(b_, &text, &off, &len, &ed, &ast) != B200ALN_SUCCESS) return generic_error;
        for (size_t k = 0; k < alignments_.size(); ++k) {
            Alignment& a = *alignments_[k];
            a.status_ = to_status(ast[k]);
            a.edit_distance_ = ed[k];
            a.batch_ = b_;
            a.index_ = static_cast<int32_t>(k);
            a.alignment_.clear();
            a.expanded_ = false;
            a.cigar_.assign(ast[k] == B200ALN_SUCCESS ? text + off[k] : "", ast[k] == B200ALN_SUCCESS ? static_cast<size_t>(len[k]) : 0);
        }
        return success;
    }
    const std::vector<std::shared_ptr<Alignment>>& get_alignments() const { return alignments_; }
    void reset() {
        detach();
        alignments_.clear();
        b200aln_batch_reset(b_);
    }
    b200aln_batch* handle() const { return b_; }

private:
    void detach() { /* Alignment objects a caller still holds keep what they have, but can no longer reach the batch */
        for (auto& a : alignments_) {
            if (a.use_count() > 1) a->get_alignment();
            a->batch_ = nullptr;
        }
    }
    b200aln_batch* b_ = nullptr;
    std::vector<std::shared_ptr<Alignment>> alignments_;
};

/* aligner.hpp:121-132 (the overload racon calls, src/cuda/cudaaligner.cpp:40-44) */
inline std::unique_ptr<Aligner> create_aligner(AlignmentType type, int32_t max_bandwidth, void* stream, int32_t device_id,
                                               int64_t max_device_memory = -1) {
    if (type != global_alignment) throw std::invalid_argument("create_aligner: only global alignment is provided");
    return std::unique_ptr<Aligner>(new Aligner(max_bandwidth, stream, device_id, max_device_memory));
}

} // namespace b200aln_cpp
