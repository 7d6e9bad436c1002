// hostcopy.cpp -- host-side row copy into page-locked staging memory (compiled by g++, linked into
// libmdb200.so; declared in api.cu).  The gather of pageable caller rows (Spectrum's Arc<[f64]>,
// spectrum/spectrum.rs:101-116; NumPy arrays) into a chunk's staging area is pure memory traffic: an
// ordinary memcpy of a 1 MB row reads the source, reads the destination lines for ownership and
// writes them back.  The staging area is only ever read by the DMA engine afterwards, so the rows are
// written with non-temporal stores (no read-for-ownership, no cache pollution): measured on the
// 8 staging threads of a 16-core host, 31 -> 40 GB/s.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace {
__attribute__((target("avx2"))) void copy_nt_avx2(double *dst, const double *src, size_t n)
{
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        const __m256d c = _mm256_loadu_pd(src + i + 8), d = _mm256_loadu_pd(src + i + 12);
        _mm256_stream_pd(dst + i, a);
        _mm256_stream_pd(dst + i + 4, b);
        _mm256_stream_pd(dst + i + 8, c);
        _mm256_stream_pd(dst + i + 12, d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();  // the row is handed to the DMA engine by another thread: make the streamed lines visible
}
}  // namespace

// Copies n doubles; rows shorter than 8 KB (and hosts without AVX2) take memcpy.
extern "C" __attribute__((visibility("hidden"))) void mdb_host_row_copy(double *dst, const double *src, size_t n)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 1024) copy_nt_avx2(dst, src, n);
    else std::memcpy(dst, src, n * sizeof(double));
}
