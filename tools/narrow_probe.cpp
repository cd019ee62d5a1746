// tools/narrow_probe.cpp -- host-only probe: how fast can T threads narrow int32 PCM to int16 when the destination is (a) a
// buffer as large as the input's half, written with streaming stores (what the feeder does), (b) a small ring per thread that
// stays in the cache, written with ordinary stores, (c) that ring with streaming stores.  g++ -O2 -mavx2 -pthread.
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static void narrow(const int32_t *src, int16_t *dst, size_t n, bool nt)
{
    for (size_t i = 0; i + 16 <= n; i += 16) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 8));
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packs_epi32(a, b), 0xD8);
        if (nt) { _mm256_stream_si256((__m256i *)(dst + i), p); } else { _mm256_store_si256((__m256i *)(dst + i), p); }
    }
    if (nt) { _mm_sfence(); }
}
int main(int argc, char **argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 16;
    const size_t N = 82ull << 20;                       // samples (328 MB of int32)
    int32_t *src = (int32_t *)aligned_alloc(64, N * 4); int16_t *big = (int16_t *)aligned_alloc(64, N * 2);
    for (size_t i = 0; i < N; i++) { src[i] = (int32_t)(i * 2654435761u) >> 17; }
    memset(big, 0, N * 2);
    const size_t ring_samples = (argc > 2 ? atoi(argv[2]) : 1024) * 1024 / 2;      // KB per thread ring
    std::vector<int16_t *> ring(T);
    for (int t = 0; t < T; t++) { ring[t] = (int16_t *)aligned_alloc(64, ring_samples * 2); memset(ring[t], 0, ring_samples * 2); }
    for (int mode = 0; mode < 3; mode++) {
        double best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            std::atomic<size_t> next{0};
            const size_t chunk = ring_samples < (256u << 10) ? ring_samples : (256u << 10);
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++) {
                th.emplace_back([&, t] {
                    size_t pos = 0;
                    for (;;) {
                        const size_t at = next.fetch_add(chunk);
                        if (at >= N) { break; }
                        const size_t n = (N - at < chunk) ? N - at : chunk;
                        if (mode == 0) { narrow(src + at, big + at, n, true); }
                        else { if (pos + n > ring_samples) { pos = 0; } narrow(src + at, ring[t] + pos, n, mode == 2); pos += n; }
                    }
                });
            }
            for (auto &x : th) { x.join(); }
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms < best) { best = ms; }
        }
        printf("threads %d mode %d (%s): %.2f ms = %.1f Gsamples/s, %.1f GB/s read\n", T, mode,
               mode == 0 ? "big buffer, streaming stores" : mode == 1 ? "per-thread ring, ordinary stores" : "per-thread ring, streaming stores",
               best, N / best / 1e6, N * 4 / best / 1e6);
    }
    return 0;
}
