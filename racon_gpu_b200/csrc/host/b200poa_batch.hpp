/*
 * b200poa_batch.hpp -- header-only C++ shim with the method set of
 * claraparabricks::genomeworks::cudapoa::Batch (/root/reference/vendor/GenomeWorks/cudapoa/include/
 * claraparabricks/genomeworks/cudapoa/batch.hpp:88-160) over the C ABI (include/b200poa.h), so the
 * body of racon's src/cuda/cudabatch.cpp builds against it with only its includes/usings changed
 * (see INTEGRATION.md).  Exceptions are raised on this side of the ABI only.
 */
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200poa.h"

namespace b200poa_cpp {

enum StatusType { /* cudapoa.hpp:32-45 */
    success = 0,
    exceeded_maximum_poas,
    exceeded_maximum_sequence_size,
    exceeded_maximum_sequences_per_poa,
    node_count_exceeded_maximum_graph_size,
    edge_count_exceeded_maximum_graph_size,
    exceeded_adaptive_banded_matrix_size,
    seq_len_exceeded_maximum_nodes_per_window,
    loop_count_exceeded_upper_bound,
    output_type_unavailable,
    generic_error
};
enum BandMode { full_band = 0, static_band, adaptive_band };
enum OutputType { consensus = 0x1, msa = 0x1 << 1 };

struct Entry { /* batch.hpp:45-53 + the layer span (extension, defaults = spans the window) */
    const char* seq;
    const int8_t* weights;
    int32_t length;
    int32_t begin = -1;
    int32_t end = -1;
};
typedef std::vector<Entry> Group;

struct BatchConfig { /* batch.hpp:57-80 */
    b200poa_config c;
    BatchConfig(int32_t max_seq_sz = 1024, int32_t max_seq_per_poa = 100, int32_t band_width = 256,
                BandMode banding = BandMode::full_band) {
        if (max_seq_sz < 0 || max_seq_per_poa < 0 || band_width < 0)
            throw std::invalid_argument("BatchConfig: negative size"); /* batch.cu:64-66 */
        if (banding == BandMode::adaptive_band)
            throw std::invalid_argument("BatchConfig: adaptive_band is not provided (racon never selects it, cudabatch.cpp:59)");
        b200poa_config_default(&c, max_seq_sz, max_seq_per_poa, band_width,
                               banding == BandMode::full_band ? B200POA_FULL_BAND : B200POA_STATIC_BAND);
    }
};

inline StatusType to_status(int32_t st) {
    return st <= static_cast<int32_t>(generic_error) ? static_cast<StatusType>(st) : generic_error;
}

class Batch {
public:
    Batch(int32_t device_id, void* stream, size_t max_gpu_mem, int8_t output_mask, const BatchConfig& cfg,
          int16_t gap_score, int16_t mismatch_score, int16_t match_score) {
        const int32_t st = b200poa_batch_create(device_id, stream, max_gpu_mem, output_mask, &cfg.c,
                                                gap_score, mismatch_score, match_score, &b_);
        if (st == B200POA_INVALID_ARGUMENT) throw std::invalid_argument("create_batch: invalid configuration or memory budget");
        if (st != B200POA_SUCCESS) throw std::runtime_error(std::string("create_batch: ") + b200poa_status_string(st));
    }
    ~Batch() { b200poa_batch_destroy(b_); }
    Batch(const Batch&) = delete;
    Batch& operator=(const Batch&) = delete;

    StatusType add_poa_group(std::vector<StatusType>& per_seq_status, const Group& poa_group) {
        std::vector<b200poa_entry> e(poa_group.size());
        for (size_t i = 0; i < poa_group.size(); ++i)
            e[i] = b200poa_entry{poa_group[i].seq, poa_group[i].weights, poa_group[i].length, poa_group[i].begin, poa_group[i].end};
        std::vector<int32_t> st(poa_group.size(), 0);
        const int32_t r = b200poa_batch_add_group(b_, e.data(), static_cast<int32_t>(e.size()), st.data());
        if (r == B200POA_INVALID_ARGUMENT) throw std::invalid_argument("Base weights need to be non-negative"); /* cudapoa_batch.cuh:533-537 */
        if (r != B200POA_SUCCESS) return to_status(r);
        per_seq_status.clear(); /* batch.hpp:95: "This API clears old entries" */
        for (int32_t s : st) per_seq_status.push_back(to_status(s));
        return success;
    }
    int32_t get_total_poas() const { return b200poa_batch_total_poas(b_); }
    void generate_poa() {
        const int32_t st = b200poa_batch_generate(b_);
        if (st != B200POA_SUCCESS) throw std::runtime_error(std::string("generate_poa: ") + b200poa_status_string(st));
    }
    StatusType get_consensus(std::vector<std::string>& consensus, std::vector<std::vector<uint16_t>>& coverage,
                             std::vector<StatusType>& output_status) {
        const uint8_t* c; const uint16_t* v; const int32_t* l; const int32_t* s; const int32_t* off;
        const int32_t r = b200poa_batch_get_consensus(b_, &c, &v, &l, &s, &off, nullptr);
        if (r != B200POA_SUCCESS) return to_status(r);
        const int32_t n = get_total_poas();
        for (int32_t i = 0; i < n; ++i) {
            output_status.emplace_back(to_status(s[i]));
            if (s[i] != B200POA_SUCCESS) { /* cudapoa_batch.cuh:232-241: empty placeholders */
                consensus.emplace_back(std::string());
                coverage.emplace_back(std::vector<uint16_t>());
                continue;
            }
            const size_t o = static_cast<size_t>(off[i]); /* compact arenas: window i lives at its own offset */
            consensus.emplace_back(reinterpret_cast<const char*>(c + o), static_cast<size_t>(l[i]));
            if (v) coverage.emplace_back(v + o, v + o + l[i]);
            else coverage.emplace_back(std::vector<uint16_t>());
        }
        return success;
    }
    /* batch.hpp:141-142 / cudapoa_batch.cuh:260-313: one vector of rows per window (empty where the window failed) */
    StatusType get_msa(std::vector<std::vector<std::string>>& msa, std::vector<StatusType>& output_status) {
        const uint8_t* m; const int64_t* off; const int32_t* rows; const int32_t* cols; const int32_t* s;
        const int32_t r = b200poa_batch_get_msa(b_, &m, &off, &rows, &cols, &s);
        if (r != B200POA_SUCCESS) return to_status(r);
        const int32_t n = get_total_poas();
        for (int32_t i = 0; i < n; ++i) {
            msa.emplace_back(std::vector<std::string>());
            output_status.emplace_back(to_status(s[i]));
            if (s[i] != B200POA_SUCCESS) continue;
            for (int32_t k = 0; k < rows[i]; ++k)
                msa.back().emplace_back(reinterpret_cast<const char*>(m + off[i] + static_cast<int64_t>(k) * cols[i]),
                                        static_cast<size_t>(cols[i]));
        }
        return success;
    }
    int32_t batch_id() const { return b200poa_batch_id(b_); }
    void reset() { b200poa_batch_reset(b_); }
    b200poa_batch* raw() { return b_; }

private:
    b200poa_batch* b_ = nullptr;
};

/* batch.hpp:174-181 */
inline std::unique_ptr<Batch> create_batch(int32_t device_id, void* stream, size_t max_gpu_mem, int8_t output_mask,
                                           const BatchConfig& batch_size, int16_t gap_score,
                                           int16_t mismatch_score, int16_t match_score) {
    return std::unique_ptr<Batch>(new Batch(device_id, stream, max_gpu_mem, output_mask, batch_size,
                                            gap_score, mismatch_score, match_score));
}

inline StatusType Init() { return to_status(b200poa_init()); } /* cudapoa.hpp:56 */

} // namespace b200poa_cpp
