/*
* b200_window.hpp -- host-side window container, same data model as racon::Window
 * (/root/reference/src/window.hpp:23-75, src/window.cpp:15-63): a backbone plus layers, each a
 * borrowed (pointer, length) pair with an optional quality string and a (begin, end) span.
 * The CPU consensus method (Window::generate_consensus, src/window.cpp:65-142) is deliberately
 * absent: in this engine consensus is produced on the GPU only (CUDABatchProcessor).
 */
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace racon_b200 {

enum class WindowType {
    kNGS, // Next Generation Sequencing
    kTGS  // Third Generation Sequencing
};

class CUDABatchProcessor;

class Window {
public:
    uint64_t id() const { return id_; }
    uint32_t rank() const { return rank_; }
    WindowType type() const { return type_; }
    const std::string& consensus() const { return consensus_; }
    uint32_t num_sequences() const { return static_cast<uint32_t>(sequences_.size()); }

    /* src/window.cpp:42-63.  Returns false (instead of exit(1)) on invalid arguments. */
    bool add_layer(const char* sequence, uint32_t sequence_length, const char* quality,
                   uint32_t quality_length, uint32_t begin, uint32_t end) {
        if (sequence_length == 0 || begin == end) return true;
        if (quality != nullptr && sequence_length != quality_length) {
            std::fprintf(stderr, "[racon_b200::Window::add_layer] error: unequal quality size!\n");
            return false;
        }
        if (begin >= end || begin > sequences_.front().second || end > sequences_.front().second) {
            std::fprintf(stderr, "[racon_b200::Window::add_layer] error: layer begin and end positions are invalid!\n");
            return false;
        }
        sequences_.emplace_back(sequence, sequence_length);
        qualities_.emplace_back(quality, quality_length);
        positions_.emplace_back(begin, end);
        return true;
    }

    friend std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type,
                                                const char* backbone, uint32_t backbone_length,
                                                const char* quality, uint32_t quality_length);
    friend class CUDABatchProcessor;

private:
    Window(uint64_t id, uint32_t rank, WindowType type, const char* backbone, uint32_t backbone_length,
           const char* quality, uint32_t quality_length)
        : id_(id), rank_(rank), type_(type) {
        sequences_.emplace_back(backbone, backbone_length);
        qualities_.emplace_back(quality, quality_length);
        positions_.emplace_back(0, 0);
    }
    Window(const Window&) = delete;
    const Window& operator=(const Window&) = delete;

    uint64_t id_;
    uint32_t rank_;
    WindowType type_;
    std::string consensus_;
    std::vector<std::pair<const char*, uint32_t>> sequences_;
    std::vector<std::pair<const char*, uint32_t>> qualities_;
    std::vector<std::pair<uint32_t, uint32_t>> positions_;
};

/* src/window.cpp:15-28; nullptr instead of exit(1) on invalid arguments */
inline std::shared_ptr<Window> createWindow(uint64_t id, uint32_t rank, WindowType type,
                                            const char* backbone, uint32_t backbone_length,
                                            const char* quality, uint32_t quality_length) {
    if (backbone_length == 0 || backbone_length != quality_length) {
        std::fprintf(stderr, "[racon_b200::createWindow] error: empty backbone sequence/unequal quality length!\n");
        return nullptr;
    }
    return std::shared_ptr<Window>(new Window(id, rank, type, backbone, backbone_length, quality, quality_length));
}

} // namespace racon_b200
