// objectives.cpp -- synthetic objectives used to grow benchmark / test trees (part of libcsp_host.so).
// The reference ships no objective functions on this path; these are builder-supplied:
//   quadnoise : f(x) = x'Bx/2 + sigma * eta(x)   (the reference-blessed quadratic form of test/runtests.jl:827 plus a
//               deterministic hash noise eta in [-1,1) so that models are not exactly recoverable)
//   rosenbrock: sum_i 100 (x[i+1]-x[i]^2)^2 + (1-x[i])^2   (standard; config C1)
// Both are plain C functions `double f(const double* y, int n, void* ctx)` so that the same pointer can drive the
// host tree here and the CPU oracle in the tests.
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
struct QuadNoise {
    int n;
    std::vector<double> B;  // row-major, symmetric
    double sigma;
    uint64_t seed;
};
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
}  // namespace

extern "C" {

double csph_obj_quadnoise(const double* y, int n, void* ctx) {
    const QuadNoise* q = (const QuadNoise*)ctx;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        double r = 0.0;
        for (int j = 0; j < n; ++j) r += q->B[(size_t)i * n + j] * y[j];
        s += y[i] * r;
    }
    s /= 2;
    if (q->sigma != 0.0) {
        uint64_t h = q->seed;
        for (int i = 0; i < n; ++i) {
            uint64_t bits;
            memcpy(&bits, &y[i], 8);
            h = splitmix64(h ^ bits);
        }
        double eta = (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
        s += q->sigma * eta;
    }
    return s;
}
double csph_obj_rosenbrock(const double* y, int n, void*) {
    double s = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
        double a = y[i + 1] - y[i] * y[i], b = 1.0 - y[i];
        s += 100.0 * a * a + b * b;
    }
    return s;
}
void* csph_quadnoise_new(int n, const double* B, double sigma, uint64_t seed) {
    QuadNoise* q = new QuadNoise{n, std::vector<double>(B, B + (size_t)n * n), sigma, seed};
    return q;
}
void csph_quadnoise_free(void* q) { delete (QuadNoise*)q; }
void* csph_objective_ptr(int kind) {
    return kind == 0 ? (void*)&csph_obj_quadnoise : kind == 1 ? (void*)&csph_obj_rosenbrock : nullptr;
}
}
