// Host-side check of the minimizer helpers of alga_b200/csrc/common.cuh (compiled with nvcc, runs on the CPU): the sliding
// minimum must equal the from-scratch minimum for every window of random and low-complexity sequences, in both directions.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../alga_b200/csrc/common.cuh"

using namespace alga;

static uint64_t window_at(const std::vector<uint8_t> &s, size_t start, uint32_t K) {
    uint64_t w = 0;
    for (uint32_t j = 0; j < K; j++) w |= (uint64_t) s[start + j] << (2 * j);
    return w;
}

int main() {
    uint64_t rng = 12345;
    auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t) (rng >> 33); };
    long checked = 0;
    for (int trial = 0; trial < 400; trial++) {
        const uint32_t K = 12 + next() % 21;       // 12 .. 32
        const uint32_t m = 8 + next() % (K - 11);  // 8 .. K - 4
        const size_t len = K + 20 + next() % 200;
        std::vector<uint8_t> s(len);
        const int period = trial % 4 == 0 ? 1 + (int) (next() % 7) : 0;  // low-complexity reads: many equal m-mers
        for (size_t i = 0; i < len; i++) s[i] = period && i >= (size_t) period ? s[i - period] : (uint8_t) (next() & 3);
        SlidingMinimizer up, down;
        up.reset(window_at(s, 0, K), K, m);
        for (size_t st = 1; st + K <= len; st++) {
            const uint64_t w = window_at(s, st, K);
            if (up.slide_up(w, K, m) != window_minimizer_ref(w, K, m)) {
                printf("slide_up mismatch: trial %d K %u m %u start %zu\n", trial, K, m, st);
                return 1;
            }
            checked++;
        }
        down.reset(window_at(s, len - K, K), K, m);
        for (size_t st = len - K; st-- > 0;) {
            const uint64_t w = window_at(s, st, K);
            if (down.slide_down(w, K, m) != window_minimizer_ref(w, K, m)) {
                printf("slide_down mismatch: trial %d K %u m %u start %zu\n", trial, K, m, st);
                return 1;
            }
            checked++;
        }
    }
    printf("ok %ld windows\n", checked);
    return 0;
}
