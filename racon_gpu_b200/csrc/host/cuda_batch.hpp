/*
 * cuda_batch.hpp -- racon's GPU batch adapter, re-hosted on the B200 engine.
 * Same class name, factory and method set as racon::CUDABatchProcessor
 * (/root/reference/src/cuda/cudabatch.hpp:26-124, src/cuda/cudabatch.cpp:41-278).
 *
 * Behavioural deltas, all deliberate: where the reference adapter deviates from
 * racon's CPU path (src/window.cpp:65-142) this one follows the CPU path, because that is the
 * parity target:
 *   - layer spans (positions_) are forwarded, so partial-span layers are recognised;
 *   - the `trim` flag and WindowType are honoured: trim only kTGS && trim, threshold (n_seqs-1)/2;
 *   - a chimeric window keeps its untrimmed consensus and reports true (window.cpp:134-137);
 *   - kNGS windows report true;
 *   - a window whose layers were dropped by the batch limits (too long / too deep,
 *     cudabatch.cpp:143-152) reports false and holds its backbone, so that the caller's CPU path
 *     polishes it exactly (the columnar polisher can instead accept the truncation like the reference:
 *     b200poa_polisher_options::accept_truncated).
 */
#pragma once
#include <atomic>
#include <memory>
#include <vector>

#include "b200poa_batch.hpp"
#include "b200_window.hpp"

namespace racon_b200 {

class CUDABatchProcessor;
std::unique_ptr<CUDABatchProcessor> createCUDABatch(uint32_t max_window_depth, uint32_t device, size_t avail_mem,
                                                    int8_t gap, int8_t mismatch, int8_t match,
                                                    bool cuda_banded_alignment, bool trim = true);

class CUDABatchProcessor {
public:
    ~CUDABatchProcessor();

    /* True if the window could be added to the batch (false = batch full, try the next round). */
    bool addWindow(std::shared_ptr<Window> window);
    bool hasWindows() const;
    /* Runs the batch; one bool per added window, in add order. */
    const std::vector<bool>& generateConsensus();
    void reset();
    uint32_t getBatchID() const { return bid_; }

    friend std::unique_ptr<CUDABatchProcessor> createCUDABatch(uint32_t, uint32_t, size_t, int8_t, int8_t, int8_t, bool, bool);

protected:
    CUDABatchProcessor(uint32_t max_window_depth, uint32_t device, size_t avail_mem, int8_t gap, int8_t mismatch,
                       int8_t match, bool cuda_banded_alignment, bool trim);
    CUDABatchProcessor(const CUDABatchProcessor&) = delete;
    const CUDABatchProcessor& operator=(const CUDABatchProcessor&) = delete;

    void generatePOA();
    void getConsensus();
    void convertPhredQualityToWeights(const char* qual, uint32_t qual_length, std::vector<int8_t>& weights);

    static std::atomic<uint32_t> batches;
    uint32_t bid_ = 0;
    uint32_t device_ = 0;
    bool trim_ = true;
    std::unique_ptr<b200poa_cpp::Batch> cudapoa_batch_;
    void* stream_ = nullptr; /* cudaStream_t */
    std::vector<std::shared_ptr<Window>> windows_;
    std::vector<int32_t> staged_index_;        /* index in the device batch, -1 = not staged (< 3 sequences) */
    std::vector<bool> window_consensus_status_;
    std::vector<uint32_t> seqs_added_per_window_;
    std::vector<bool> dropped_layers_;
    int32_t staged_ = 0;
};

} // namespace racon_b200
