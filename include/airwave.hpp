// airwave.hpp — header-only C++17 mirror of the reference's interface for the binaural path, over the C ABI
// (airwave_cuda.h).  Same names, argument meaning and error behaviour as the Swift types it stands for:
//   airwave::ConvolutionEngine        Airwave/ConvolutionEngine.swift:14-408   (factory returns nullptr like `init?`)
//   airwave::RealtimeAudioProcessor   Airwave/RealtimeAudioProcessor.swift:11-191
//   airwave::HRIRBank / BinauralEngine  the batched objects that exist only in this implementation
// Render calls never throw and never allocate; construction failures are reported (nullptr / std::runtime_error).
#pragma once

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "airwave_cuda.h"

namespace airwave {

inline void check(int status)
{
    if (status != AW_OK) throw std::runtime_error(std::string("aw_status ") + std::to_string(status) + ": " + aw_last_error());
}

class HRIRBank {
public:
    // pcm: planar [channels][frames]; speaker i uses channels left[i] / right[i]
    HRIRBank(int device, const std::vector<float> &pcm, int channels, int frames, double src_rate, double dst_rate,
             const std::vector<int> &left, const std::vector<int> &right, int block, int resample_mode = AW_RESAMPLE_REFERENCE)
    {
        check(aw_bank_create_ex(device, pcm.data(), channels, frames, src_rate, dst_rate, left.data(), right.data(), (int)left.size(), block,
                                resample_mode, &h_));
        aw_bank_info(h_, &speakers, &this->block, &partitions, &taps);
    }
    ~HRIRBank() { aw_bank_destroy(h_); }
    HRIRBank(const HRIRBank &) = delete;
    HRIRBank &operator=(const HRIRBank &) = delete;
    aw_bank *handle() const { return h_; }
    int speakers = 0, block = 0, partitions = 0, taps = 0;

private:
    aw_bank *h_ = nullptr;
};

class BinauralEngine {
public:
    BinauralEngine(int device, int n_streams, int n_speakers, int block, double sample_rate = 48000.0, int max_frames = 4096,
                   int max_partitions = 0, unsigned flags = AW_ENGINE_DEFAULT)
    {
        aw_engine_config cfg{device, n_streams, n_speakers, block, sample_rate, max_frames, max_partitions, flags};
        check(aw_engine_create(&cfg, &h_));
    }
    ~BinauralEngine() { aw_engine_destroy(h_); }
    BinauralEngine(const BinauralEngine &) = delete;
    BinauralEngine &operator=(const BinauralEngine &) = delete;
    void setBank(const HRIRBank *bank, int first, int count) { check(aw_engine_set_bank(h_, first, count, bank ? bank->handle() : nullptr)); }
    // in [stream][speaker][frames], out [stream][2][frames], host pointers; returns the aw_status (never throws)
    int process(const float *in, float *out, int frames) noexcept { return aw_engine_process(h_, in, out, frames); }
    int processStereo(const float *l, const float *r, float *ol, float *orr, int frames) noexcept { return aw_engine_process_stereo(h_, l, r, ol, orr, frames); }
    // device pointers, strides in elements; with AW_ENGINE_OVERLAP_EQ alternate two output buffers and flush() before reading the last
    int processDevice(const float *in, long long in_ss, long long in_cs, float *out, long long out_ss, long long out_cs, int frames) noexcept
    {
        return aw_engine_process_device(h_, in, in_ss, in_cs, out, out_ss, out_cs, frames);
    }
    // pipelined offline rendering (AW_ENGINE_PIPELINED): host or device input, host output valid after wait()
    int submit(const float *in, float *out, int frames) noexcept { return aw_engine_submit(h_, in, out, frames); }
    int submitDevice(const float *in, long long in_ss, long long in_cs, float *out, int frames) noexcept
    {
        return aw_engine_submit_device(h_, in, in_ss, in_cs, out, frames);
    }
    int wait() noexcept { return aw_engine_wait(h_); }
    int flush() noexcept { return aw_engine_flush(h_); }
    void reset(int first, int count, int what = AW_RESET_SPATIAL) { check(aw_engine_reset(h_, first, count, what)); }
    aw_engine *handle() const { return h_; }

private:
    aw_engine *h_ = nullptr;
};

class ConvolutionEngine {
public:
    // ConvolutionEngine.init?(hrirSamples:blockSize:): nullptr on failure (e.g. block size not a power of two)
    static std::unique_ptr<ConvolutionEngine> make(const std::vector<float> &hrirSamples, int blockSize = 512, int device = 0)
    {
        try { return std::unique_ptr<ConvolutionEngine>(new ConvolutionEngine(hrirSamples, blockSize, device)); }
        catch (const std::exception &) { return nullptr; }
    }
    // process(input:output:): input and output hold blockSize samples
    void process(const float *input, float *output) noexcept
    {
        engine_->process(input, scratch_.data(), blockSize);
        std::copy(scratch_.begin(), scratch_.begin() + blockSize, output);
    }
    // process(input:[Float], output:, frameCount:): silently returns when frameCount != blockSize (:370-372)
    bool process(const std::vector<float> &input, std::vector<float> &output, int frameCount = -1) noexcept
    {
        if ((frameCount < 0 ? blockSize : frameCount) != blockSize) return false;
        process(input.data(), output.data());
        return true;
    }
    void processAndAccumulate(const float *input, float *outputAccumulator) noexcept
    {
        engine_->process(input, scratch_.data(), blockSize);
        for (int i = 0; i < blockSize; ++i) outputAccumulator[i] += scratch_[i];
    }
    void reset() { engine_->reset(0, 1); }
    const int blockSize;
    const std::vector<float> hrirSamples;

private:
    ConvolutionEngine(const std::vector<float> &h, int block, int device)
        : blockSize(block), hrirSamples(h), scratch_(2 * (size_t)block)
    {
        std::vector<float> pcm = h.empty() ? std::vector<float>(1, 0.f) : h;
        bank_.reset(new HRIRBank(device, pcm, 1, (int)pcm.size(), 48000.0, 48000.0, {0}, {0}, block));
        engine_.reset(new BinauralEngine(device, 1, 1, block, 48000.0, block, bank_->partitions));
        engine_->setBank(bank_.get(), 0, 1);
    }
    std::unique_ptr<HRIRBank> bank_;
    std::unique_ptr<BinauralEngine> engine_;
    std::vector<float> scratch_;
};

struct VirtualSpeakerRenderer {
    std::shared_ptr<ConvolutionEngine> convolverLeftEar, convolverRightEar;
};

class RealtimeAudioProcessor {
public:
    RealtimeAudioProcessor(const std::vector<VirtualSpeakerRenderer> &renderers, int blockSize = 512, int maxFramesPerCallback = 4096,
                           int device = 0)
        : blockSize(blockSize), maxFramesPerCallback(maxFramesPerCallback)
    {
        if (blockSize <= 0 || maxFramesPerCallback <= 0) throw std::invalid_argument("precondition failed");
        engine_.reset(new BinauralEngine(device, 1, 2, blockSize, 48000.0, maxFramesPerCallback, 0, AW_ENGINE_LITERAL_STEREO));
        const int used = std::min<int>((int)renderers.size(), 2);   // RealtimeAudioProcessor.swift:145
        if (used == 0) return;
        size_t taps = 1;
        for (int i = 0; i < used; ++i)
            taps = std::max({taps, renderers[i].convolverLeftEar->hrirSamples.size(), renderers[i].convolverRightEar->hrirSamples.size()});
        std::vector<float> pcm(2 * used * taps, 0.f);
        std::vector<int> l, r;
        for (int i = 0; i < used; ++i) {
            const auto &a = renderers[i].convolverLeftEar->hrirSamples, &b = renderers[i].convolverRightEar->hrirSamples;
            std::copy(a.begin(), a.end(), pcm.begin() + (2 * i) * taps);
            std::copy(b.begin(), b.end(), pcm.begin() + (2 * i + 1) * taps);
            l.push_back(2 * i);
            r.push_back(2 * i + 1);
        }
        bank_.reset(new HRIRBank(device, pcm, 2 * used, (int)taps, 48000.0, 48000.0, l, r, blockSize));
        engine_->setBank(bank_.get(), 0, 1);
    }
    // StereoAudioProcessing.process: inputRight may be null (mono duplicated); outputs may alias.  Returns false where the
    // reference would trip precondition(frameCount <= maxFramesPerCallback).
    bool process(const float *inputLeft, const float *inputRight, float *outputLeft, float *outputRight, int frameCount) noexcept
    {
        if (frameCount <= 0) return true;
        return engine_->processStereo(inputLeft, inputRight, outputLeft, outputRight, frameCount) == AW_OK;
    }
    void reset() { engine_->reset(0, 1); }
    const int blockSize, maxFramesPerCallback;

private:
    std::unique_ptr<HRIRBank> bank_;
    std::unique_ptr<BinauralEngine> engine_;
};

}  // namespace airwave
