// Exercises include/cpq.hpp (the C++20 host mirror of the reference interface) against libcpq.so.
// Without a GPU: prepareToPlay must fail loudly (no CPU fallback) and the host-only planner must work.
// With a GPU: a small conv -> EQ -> output run; prints a checksum the Python test compares with the ctypes path.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cpq.hpp"

static double lcg(unsigned long long& s)
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double) ((s >> 11) & ((1ULL << 53) - 1)) / (double) (1ULL << 53) - 0.5;
}

int main(int argc, char** argv)
{
    using namespace convopeq_b200;
    const bool expectGpu = argc > 1 && std::atoi(argv[1]) != 0;
    cpq_layout lay {};
    if (cpq_plan_layout(65536, 512, nullptr, 64, &lay, nullptr) != CPQ_OK || lay.num_layers != 2 || lay.layers[1].first_output_sample != 7168)
    {
        std::printf("FAIL plan\n");
        return 1;
    }
    const int T = 8192, B = 512, irLen = 20000;
    BatchEngine eng;
    const bool ok = eng.prepareToPlay(48000.0, B, 1, 2, T);
    if (!expectGpu)
    {
        if (ok || eng.lastStatus() != CPQ_ERR_CUDA) { std::printf("FAIL expected CPQ_ERR_CUDA, got %d\n", (int) eng.lastStatus()); return 1; }
        std::printf("OK nogpu: %s\n", eng.lastError().c_str());
        return 0;
    }
    if (!ok) { std::printf("FAIL prepare: %s\n", eng.lastError().c_str()); return 1; }
    unsigned long long seed = 12345;
    std::vector<double> irL(irLen), irR(irLen), l(T), r(T);
    for (int i = 0; i < irLen; ++i) { irL[i] = lcg(seed) * std::exp(-i / 3000.0) * 0.05; irR[i] = lcg(seed) * std::exp(-i / 3000.0) * 0.05; }
    for (int i = 0; i < T; ++i) { l[i] = 0.2 * lcg(seed); r[i] = 0.2 * lcg(seed); }
    FilterSpec spec = defaultFilterSpec();
    if (!eng.init(0, irL, irR, 1.0, &spec)) { std::printf("FAIL init: %s\n", eng.lastError().c_str()); return 1; }
    for (int b = 0; b < CPQ_NUM_BANDS; ++b)
    {
        eng.setBandGain(0, b, (b % 2 ? 3.0f : -2.5f));
        eng.setBandQ(0, b, 1.0f + 0.1f * b);
        eng.setBandType(0, b, b == 0 ? 0 : (b == 19 ? 2 : 1));
    }
    eng.setTotalGain(0, -1.0f);
    eng.setOutputStage(1.1, 0);
    double* planar[2] = { l.data(), r.data() };
    if (!eng.process(planar, T)) { std::printf("FAIL process: %s\n", eng.lastError().c_str()); return 1; }
    double sum = 0.0, sq = 0.0;
    for (int i = 0; i < T; ++i) { sum += l[i] - r[i]; sq += l[i] * l[i] + r[i] * r[i]; }
    std::printf("OK gpu: sum=%.17g sq=%.17g latency=%d\n", sum, sq, eng.getLatency());
    return 0;
}
