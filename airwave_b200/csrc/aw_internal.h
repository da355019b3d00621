// aw_internal.h — private declarations shared by the translation units of libairwave_cuda.so.
#pragma once
#include <string>
#include <vector>

struct aw_wav {              // WAVData, WAVLoader.swift:12-17
    double sample_rate = 0;
    int channels = 0;
    int frames = 0;
    std::vector<float> data; // planar [channel][frame]
};

namespace aw {
// Records the message for aw_last_error() on this thread and returns `status`.
int set_error(int status, const std::string &message);
}  // namespace aw
