// Host-side probe behind the CopyTeam settings of afb_api.cu (not a benchmark of record):
// bandwidth of a chunked pageable -> buffer copy by T threads, plain memcpy against SSE2 streaming stores.
//   g++ -O2 -pthread tools/host_copy_probe.cpp -o /tmp/host_copy_probe && /tmp/host_copy_probe
#include <emmintrin.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void copy_nt(char* dst, const char* src, size_t bytes) {
    size_t i = 0;
    while (((uintptr_t)(dst + i) & 15) && i < bytes) { dst[i] = src[i]; ++i; }
    for (; i + 64 <= bytes; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i));
        const __m128i b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32));
        const __m128i d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c);
        _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    for (; i < bytes; ++i) dst[i] = src[i];
    _mm_sfence();
}

int main() {
    const size_t total = (size_t)2 << 30, chunk = (size_t)32 << 20;
    char* src = (char*)malloc(total);
    char* dst = (char*)malloc(3 * chunk);
    memset(src, 1, total);
    memset(dst, 2, 3 * chunk);
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    for (int nt = 0; nt < 2; ++nt)
        for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
            if ((unsigned)T > std::thread::hardware_concurrency()) continue;
            double best = 0;
            for (int rep = 0; rep < 3; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                for (size_t off = 0, c = 0; off < total; off += chunk, ++c) {
                    std::vector<std::thread> th;
                    char* d = dst + (c % 3) * chunk;
                    for (int t = 0; t < T; ++t)
                        th.emplace_back([=] {
                            const size_t a = chunk * t / T & ~(size_t)63, b = (t + 1 == T) ? chunk : (chunk * (t + 1) / T & ~(size_t)63);
                            if (nt) copy_nt(d + a, src + off + a, b - a);
                            else memcpy(d + a, src + off + a, b - a);
                        });
                    for (auto& x : th) x.join();
                }
                const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                best = std::max(best, (double)total / s / 1e9);
            }
            printf("%s threads %2d : %6.1f GB/s\n", nt ? "stream" : "memcpy", T, best);
        }
    return 0;
}
