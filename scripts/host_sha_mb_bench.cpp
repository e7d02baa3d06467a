// Host SHA-256 throughput: SHA-NI single stream vs AVX-512 multi-buffer (16 messages in lockstep), T threads.
//   g++ -O3 -std=c++17 -pthread -o build/host_sha_mb_bench scripts/host_sha_mb_bench.cpp rust-kzg-bn254_b200/csrc/sha256.cpp
//   build/host_sha_mb_bench [MiB per message = 16] [threads ...]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../rust-kzg-bn254_b200/csrc/sha256.hpp"
using namespace kzgb;
int main(int argc, char** argv) {
    size_t mib = argc > 1 ? atoi(argv[1]) : 16;
    size_t len = mib << 20;
    std::vector<int> ts;
    for (int i = 2; i < argc; i++) ts.push_back(atoi(argv[i]));
    if (ts.empty()) ts = {1, 2, 4, 8};
    int maxT = 0; for (int t : ts) maxT = t > maxT ? t : maxT;
    // 16 messages per thread, distinct memory (so that caches do not help)
    std::vector<std::vector<uint8_t>> bufs((size_t)maxT * 16);
    for (size_t i = 0; i < bufs.size(); i++) { bufs[i].resize(len); for (size_t j = 0; j < len; j += 4096) bufs[i][j] = (uint8_t)(i + j); memset(bufs[i].data(), (int)i & 0x2f, 64); }
    printf("shani=%d mb16=%d, %zu MiB per message\n", (int)sha256_has_shani(), (int)sha256_has_mb16(), mib);
    for (int T : ts) {
        for (int mode = 0; mode < 2; mode++) {
            if (mode == 1 && !sha256_has_mb16()) continue;
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) th.emplace_back([&, t]() {
                if (mode == 0) {
                    for (int m = 0; m < 16; m++) { Sha256 s; s.update(bufs[t * 16 + m].data(), len); volatile uint32_t x = s.h[0]; (void)x; }
                } else {
                    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
                    uint32_t st[16][8]; const uint8_t* p[16];
                    for (int m = 0; m < 16; m++) { memcpy(st[m], iv, 32); p[m] = bufs[t * 16 + m].data(); }
                    sha256_mb16_blocks(st, p, len / 64, nullptr);
                }
            });
            for (auto& x : th) x.join();
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("T=%d %s: %.3f s for %d x 16 messages -> %.2f GB/s total, %.2f GB/s per thread\n", T, mode ? "mb16 " : "shani", dt, T,
                   T * 16.0 * len / dt / 1e9, 16.0 * len / dt / 1e9);
        }
    }
}
