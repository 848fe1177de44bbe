/*
 * cuda_polisher.hpp -- the GPU window scheduler of racon's CUDAPolisher::polish
 * (/root/reference/src/cuda/cudapolisher.cpp:216-345): `cudapoa_batches` batch processors per device,
 * one host thread each, all pulling windows from one shared cursor until none are left.
 * Everything upstream of the windows (parsing, overlap alignment, window construction) and the
 * contig stitching downstream (cudapolisher.cpp:385-411) stay in racon and are out of scope.
 */
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include "b200_window.hpp"

namespace racon_b200 {

struct PolishOptions {
    std::vector<int32_t> devices;   /* empty = every visible device (cudapolisher.cpp:45-63) */
    uint32_t cudapoa_batches = 1;   /* racon -c N : batch processors per device */
    bool cuda_banded_alignment = false; /* racon -b */
    int8_t match = 3, mismatch = -5, gap = -4;
    bool trim = true;
    uint32_t max_depth_per_window = 200; /* cudapolisher.cpp:226 */
    size_t mem_per_batch = 0;       /* 0 = 0.9 * free / cudapoa_batches (cudapolisher.cpp:233-236) */
    uint32_t max_windows_per_round = 0; /* 0 = fill the batch (reference behaviour); >0 caps a round so
                                           that several batches / devices overlap on small inputs */
};

/* Polishes every window in place (Window::consensus_) and returns the per-window status vector
 * (true = polished on the GPU; false = left for the caller's CPU path, cudapolisher.cpp:354-383). */
std::vector<bool> polish_windows(std::vector<std::shared_ptr<Window>>& windows, const PolishOptions& opt);

} // namespace racon_b200
