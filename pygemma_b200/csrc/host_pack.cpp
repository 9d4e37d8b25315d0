// host_pack.cpp -- host side of the pageable-genotype path of pg_scan (plain C++, compiled by the host compiler).
//
// A caller's NumPy array is pageable memory; a pageable cudaMemcpy2D runs at ~10 GB/s.  pg_scan instead packs each SNP
// block into one of two pinned bounce buffers with a few host threads while the GPU works on the previous block, and
// uploads from there at PCIe speed (reference counterpart: the second full copy of X into multiprocessing.Array,
// lmm/lmm.py:364-376).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace pg {

// One row into the pinned bounce buffer with streaming (non-temporal) stores: the destination is only read by the DMA
// engine, so it should neither be read for ownership nor displace the caller's matrix from the host caches.  A plain
// memcpy of a 25 KB row stays below glibc's non-temporal threshold and costs a third more memory traffic.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void copy_row_stream_avx2(char* dst, const char* src, size_t bytes)
{
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head > bytes) head = bytes;
    if (head) { memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
    size_t body = bytes & ~(size_t)127;
    for (size_t i = 0; i < body; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64));
        const __m256i d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a);
        _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c);
        _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    if (bytes > body) memcpy(dst + body, src + body, bytes - body);
}
#endif

static void copy_row(char* dst, const char* src, size_t bytes)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("PG_NO_STREAM_COPY");
    if (avx2 && bytes >= 1024) { copy_row_stream_avx2(dst, src, bytes); return; }
#endif
    memcpy(dst, src, bytes);
}

// Packs the strided host block (rows of `width` bytes, `rows` of them, source pitch `pitch`) into dst, contiguously.
void pack_block_threads(char* dst, const char* src, size_t pitch, size_t width, size_t rows)
{
    const unsigned hw = std::thread::hardware_concurrency();
    // one process per GPU (torchrun sets LOCAL_WORLD_SIZE): the ranks of a node share its cores -- 8 x 16 packing threads on
    // 32 cores were measured slower than the upload they feed
    static const size_t local_world = getenv("LOCAL_WORLD_SIZE") ? (size_t)std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
    static const size_t max_threads = getenv("PG_PACK_THREADS") ? (size_t)std::max(1, atoi(getenv("PG_PACK_THREADS")))
                                                                : std::max<size_t>(2, std::min<size_t>(16, (hw ? hw : 4) / local_world));
    const size_t nt = std::max<size_t>(1, std::min<size_t>({max_threads, (size_t)(hw ? hw : 4), rows}));
    auto work = [=](size_t t) {
        const size_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        for (size_t r = r0; r < r1; ++r) copy_row(dst + r * width, src + r * pitch, width);
#if defined(__x86_64__)
        _mm_sfence();   // streaming stores are weakly ordered: make them visible before the DMA is enqueued
#endif
    };
    if (nt == 1 || width * rows < (size_t(1) << 22)) { for (size_t t = 0; t < nt; ++t) work(t); return; }
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

}  // namespace pg
