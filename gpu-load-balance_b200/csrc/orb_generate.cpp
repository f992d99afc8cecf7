// orb_generate.cpp — deterministic host-side particle generators shared by every driver
// (C++ orbit host, Python bench/tests).  No device work here.
//
// Uniform: the reference's generator (init.cu:11-25) and fill order (init.cu:47-53), one stream with the
// fixed seeds, so a rank can take its contiguous slice of the single-thread reference sequence.
// Clustered: not in the reference; defined in SURVEY.md §8(d).
#include <cmath>
#include <cstdint>
#include <limits>

#include "../../include/orb_b200.h"

namespace {

struct Xorshf96 {
    uint64_t x = 123456789ULL, y = 362436069ULL, z = 521288629ULL;   // init.cu:11
    // init.cu:13-25: period 2^96-1 xorshift; value = (float) z / ULONG_MAX - 0.5 (float division, double
    // subtraction, rounded to float on return)
    float next() {
        uint64_t t;
        x ^= x << 16;
        x ^= x >> 5;
        x ^= x << 1;
        t = x;
        x = y;
        y = z;
        z = t ^ x ^ y;
        return (float)z / std::numeric_limits<unsigned long>::max() - 0.5;
    }
};

inline float clampBox(float v) { return v < -0.5f ? -0.5f : (v > 0.5f ? 0.5f : v); }

}  // namespace

extern "C" void orb_generate_uniform(uint64_t skip, uint64_t n, float *x, float *y, float *z) {
    Xorshf96 g;
    for (uint64_t i = 0; i < skip; ++i) { g.next(); g.next(); g.next(); }
    for (uint64_t i = 0; i < n; ++i) {   // init.cu:48-52: for i: for d: particles(i,d) = xorshf96()
        x[i] = g.next();
        y[i] = g.next();
        z[i] = g.next();
    }
}

extern "C" void orb_generate_clustered(int kind, uint64_t skip, uint64_t n, float *x, float *y, float *z) {
    constexpr int K = 64;
    Xorshf96 g;
    float cx[K], cy[K], cz[K];
    for (int k = 0; k < K; ++k) {   // clump centres: first 3K draws scaled to [-0.4, 0.4]
        cx[k] = 0.8f * g.next();
        cy[k] = 0.8f * g.next();
        cz[k] = 0.8f * g.next();
    }
    const float sigma = 0.02f, a = 0.01f, twoPi = 6.28318530717958647692f;
    auto u01 = [&]() {   // uniform in (0,1]
        float u = g.next() + 0.5f;
        return u < 1e-7f ? 1e-7f : u;
    };
    for (uint64_t i = 0; i < skip + n; ++i) {
        float dx, dy, dz;
        if (kind == 0) {   // Gaussian clump: Box-Muller on four draws, three normals used
            float u1 = u01(), u2 = u01(), u3 = u01(), u4 = u01();
            float r1 = std::sqrt(-2.0f * std::log(u1)), r2 = std::sqrt(-2.0f * std::log(u3));
            dx = sigma * r1 * std::cos(twoPi * u2);
            dy = sigma * r1 * std::sin(twoPi * u2);
            dz = sigma * r2 * std::cos(twoPi * u4);
        } else {           // Plummer sphere: r = a / sqrt(u^(-2/3) - 1), isotropic direction
            float u = u01(), v = u01(), w = u01();
            if (u > 0.999f) u = 0.999f;
            float r = a / std::sqrt(std::pow(u, -2.0f / 3.0f) - 1.0f);
            float ct = 2.0f * v - 1.0f, st = std::sqrt(std::fmax(0.0f, 1.0f - ct * ct));
            dx = r * st * std::cos(twoPi * w);
            dy = r * st * std::sin(twoPi * w);
            dz = r * ct;
        }
        if (i >= skip) {
            const uint64_t o = i - skip;
            const int k = (int)(i % K);
            x[o] = clampBox(cx[k] + dx);
            y[o] = clampBox(cy[k] + dy);
            z[o] = clampBox(cz[k] + dz);
        }
    }
}
