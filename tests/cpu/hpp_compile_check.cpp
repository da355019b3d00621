// Compile-only check of include/airwave.hpp against include/airwave_cuda.h (no GPU needed).
#include "../../include/airwave.hpp"
int main(int argc, char **)
{
    if (argc > 100) {   // never executed: instantiates the mirror classes so the header is type-checked
        auto e = airwave::ConvolutionEngine::make({1.f}, 8);
        airwave::RealtimeAudioProcessor p({{std::move(e), nullptr}}, 8, 64);
        float x[8] = {0}, y[8];
        p.process(x, nullptr, y, y, 8);
    }
    return aw_device_count() >= 0 ? 0 : 1;
}
