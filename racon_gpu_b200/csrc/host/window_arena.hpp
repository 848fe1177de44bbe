/*
 * window_arena.hpp -- racon's windows built straight into the columnar arena the engine consumes
 * (SURVEY.md 8(f)-2): what Polisher::initialize (/root/reference/src/polisher.cpp:384-457) would fill instead
 * of one racon::Window object (three std::vectors of pairs, src/window.hpp:71-73) per window.
 *
 * Same contract as racon::createWindow / Window::add_layer (src/window.cpp:15-63): the same argument checks,
 * layers of one window keep the order in which they were added (that order is what the unstable std::sort of
 * window.cpp:78-85 starts from), sequence data is BORROWED until finalize() copies it.  Layers of different
 * windows may arrive interleaved, exactly as racon adds them (one overlap contributes a layer to every window
 * it crosses): finalize() groups them with a stable counting sort.
 *
 * finalize() produces:  win_seq_off[W+1], seq_off[S+1], bases[B], weights[B] (quality - 33, graph.cpp:138-147;
 * 1 where a sequence has no quality, :124-129), has_weights[S], begins[S], ends[S]  -- the arguments of
 * b200poa_polisher_polish / b200poa_polish_windows.
 */
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace racon_b200 {

class WindowArena {
public:
    /* src/window.cpp:15-28.  Returns the window's index, or -1 on invalid arguments (racon exits). */
    int64_t add_window(const char* backbone, uint32_t backbone_length, const char* quality, uint32_t quality_length) {
        if (finalized_) return -1;
        if (backbone_length == 0 || backbone_length != quality_length) {
            std::fprintf(stderr, "[racon_b200::WindowArena::add_window] error: empty backbone sequence/unequal quality length!\n");
            return -1;
        }
        const int64_t w = static_cast<int64_t>(backbone_len_.size());
        backbone_len_.push_back(backbone_length);
        layers_.push_back(Layer{w, backbone, quality, backbone_length, 0u, 0u}); /* window.cpp:35-37: span (0, 0) */
        return w;
    }

    /* src/window.cpp:42-63.  false on invalid arguments (racon exits); an empty layer is silently skipped. */
    bool add_layer(int64_t window, const char* sequence, uint32_t sequence_length, const char* quality,
                   uint32_t quality_length, uint32_t begin, uint32_t end) {
        if (finalized_ || window < 0 || window >= static_cast<int64_t>(backbone_len_.size())) return false;
        if (sequence_length == 0 || begin == end) return true;
        if (quality != nullptr && sequence_length != quality_length) {
            std::fprintf(stderr, "[racon_b200::WindowArena::add_layer] error: unequal quality size!\n");
            return false;
        }
        const uint32_t bb = backbone_len_[static_cast<size_t>(window)];
        if (begin >= end || begin > bb || end > bb) {
            std::fprintf(stderr, "[racon_b200::WindowArena::add_layer] error: layer begin and end positions are invalid!\n");
            return false;
        }
        layers_.push_back(Layer{window, sequence, quality, sequence_length, begin, end});
        return true;
    }

    /* group the layers by window (stable), copy bases and weights into the arena; the builder is read-only afterwards */
    void finalize() {
        if (finalized_) return;
        const size_t W = backbone_len_.size(), S = layers_.size();
        win_seq_off_.assign(W + 1, 0);
        for (const Layer& l : layers_) win_seq_off_[static_cast<size_t>(l.window) + 1] += 1;
        for (size_t w = 0; w < W; ++w) win_seq_off_[w + 1] += win_seq_off_[w];
        std::vector<int64_t> slot(win_seq_off_.begin(), win_seq_off_.end() - 1), where(S);
        for (size_t i = 0; i < S; ++i) where[i] = slot[static_cast<size_t>(layers_[i].window)]++;
        seq_off_.assign(S + 1, 0);
        for (size_t i = 0; i < S; ++i) seq_off_[static_cast<size_t>(where[i]) + 1] = layers_[i].length;
        for (size_t s = 0; s < S; ++s) seq_off_[s + 1] += seq_off_[s];
        bases_.resize(static_cast<size_t>(seq_off_[S]));
        weights_.resize(bases_.size());
        has_weights_.assign(S, 0);
        begins_.assign(S, 0);
        ends_.assign(S, 0);
        for (size_t i = 0; i < S; ++i) {
            const Layer& l = layers_[i];
            const size_t s = static_cast<size_t>(where[i]), o = static_cast<size_t>(seq_off_[s]);
            std::memcpy(bases_.data() + o, l.seq, l.length);
            if (l.quality != nullptr) {
                has_weights_[s] = 1;
                for (uint32_t k = 0; k < l.length; ++k)
                    weights_[o + k] = static_cast<int8_t>(static_cast<int32_t>(static_cast<uint8_t>(l.quality[k])) - 33);
            } else {
                std::memset(weights_.data() + o, 1, l.length);
            }
            begins_[s] = static_cast<int32_t>(l.begin);
            ends_[s] = static_cast<int32_t>(l.end);
        }
        layers_.clear();
        layers_.shrink_to_fit();
        finalized_ = true;
    }

    bool finalized() const { return finalized_; }
    int64_t n_windows() const { return static_cast<int64_t>(backbone_len_.size()); }
    int64_t n_sequences() const { return finalized_ ? static_cast<int64_t>(seq_off_.size()) - 1 : static_cast<int64_t>(layers_.size()); }
    const std::vector<int64_t>& win_seq_off() const { return win_seq_off_; }
    const std::vector<int64_t>& seq_off() const { return seq_off_; }
    const std::vector<uint8_t>& bases() const { return bases_; }
    const std::vector<int8_t>& weights() const { return weights_; }
    const std::vector<uint8_t>& has_weights() const { return has_weights_; }
    const std::vector<int32_t>& begins() const { return begins_; }
    const std::vector<int32_t>& ends() const { return ends_; }

private:
    struct Layer {
        int64_t window;
        const char* seq;
        const char* quality;
        uint32_t length, begin, end;
    };
    bool finalized_ = false;
    std::vector<uint32_t> backbone_len_;
    std::vector<Layer> layers_;
    std::vector<int64_t> win_seq_off_, seq_off_;
    std::vector<uint8_t> bases_, has_weights_;
    std::vector<int8_t> weights_;
    std::vector<int32_t> begins_, ends_;
};

} // namespace racon_b200
