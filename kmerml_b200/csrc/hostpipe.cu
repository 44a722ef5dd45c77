// hostpipe.cu -- the narrow device -> host wire format of kmerml_count_dense_host.
//
// The host-buffer call is bound by PCIe / the host memory system: a genome's k = 1..12 count row is 89.5 MB of
// uint32, of which the 4^10 + 4^11 + 4^12 bins of k = 10..12 are 98 % -- and for a genome of a few ten megabases
// nearly all of those counts fit one byte, those of k = 12 even four bits.  So a level k >= 10 crosses the bus as ONE
// BYTE per bin -- or as a NIBBLE per bin when the genome is small enough for a mean count <= 5 -- plus a short
// exception list (bin, count) for the bins that reached 255 (15) (narrow_levels_kernel); the small levels cross as they are.  A pool
// of host threads widens the bytes into the caller's uint32 row (non-temporal stores) while the next genomes'
// bytes and counts are on the bus.  Lossless: a genome whose exception list overflows is copied in full.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "internal.h"

namespace km {

// ---------------------------------------------------------------- device side
// One wire row per genome (blockIdx.y); a thread produces 16 bytes of the narrow block: 16 bins of a byte level
// (four 128-bit loads, bins >= 255 written as 255) or 32 bins of a nibble level (eight loads, bins >= 15 written as
// 15); those bins are listed as (row-relative bin, count).  Behind the narrow block the small levels are copied as
// they are, four uint32 per thread.
__global__ void __launch_bounds__(256)
narrow_levels_kernel(const uint32_t* __restrict__ counts_all, uint64_t counts_stride, NarrowSpec spec,
                     uint8_t* __restrict__ wire_all, uint64_t wire_stride, uint64_t narrow_off, uint64_t exc_off,
                     uint64_t small_off, uint32_t exc_cap) {
    const uint32_t* counts = counts_all + (uint64_t)blockIdx.y * counts_stride;
    uint8_t* wire = wire_all + (uint64_t)blockIdx.y * wire_stride;
    unsigned int* exc_count = reinterpret_cast<unsigned int*>(wire + exc_off);
    uint2* exc = reinterpret_cast<uint2*>(wire + exc_off + 16);
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_narrow = spec.total >> 4;
    auto note = [&](uint64_t bin, uint32_t value) {
        const unsigned int slot = atomicAdd(exc_count, 1u);
        if (slot < exc_cap) exc[slot] = make_uint2((uint32_t)bin, value);
    };
    if (v < n_narrow) {
        const uint64_t e = v << 4;                                // byte of the narrow block
        int seg = 0;
        while (seg + 1 < spec.n && e >= spec.dst_off[seg + 1]) seg++;
        uint32_t w[4];
        if (!spec.nibble[seg]) {
            const uint64_t src = spec.src_off[seg] + (e - spec.dst_off[seg]);
            const uint4* p = reinterpret_cast<const uint4*>(counts + src);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 c = __ldg(p + i);
                const uint32_t x[4] = {c.x, c.y, c.z, c.w};
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t b = x[j];
                    if (b >= 255u) { note(src + 4 * i + j, b); b = 255u; }
                    packed |= b << (8 * j);
                }
                w[i] = packed;
            }
        } else {
            const uint64_t src = spec.src_off[seg] + 2 * (e - spec.dst_off[seg]);
            const uint4* p = reinterpret_cast<const uint4*>(counts + src);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 c0 = __ldg(p + 2 * i), c1 = __ldg(p + 2 * i + 1);
                const uint32_t x[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t b = x[j];
                    if (b >= 15u) { note(src + 8 * i + j, b); b = 15u; }
                    packed |= b << (4 * j);
                }
                w[i] = packed;
            }
        }
        *reinterpret_cast<uint4*>(wire + narrow_off + e) = make_uint4(w[0], w[1], w[2], w[3]);
        return;
    }
    const uint64_t q = (v - n_narrow) << 2;                       // element of the small block
    if (q >= spec.small_total) return;
    int seg = 0;
    while (seg + 1 < spec.n_small && q >= spec.small_dst[seg + 1]) seg++;
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(counts + spec.small_src[seg] + (q - spec.small_dst[seg])));
    *reinterpret_cast<uint4*>(wire + small_off + q * 4) = c;
}

// header (magic, which levels are nibble-packed) and a zero exception count per wire row
__global__ void wire_header_kernel(uint8_t* wire_all, uint64_t wire_stride, uint64_t exc_off, int n, uint32_t magic,
                                   uint32_t mask) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    *reinterpret_cast<uint4*>(wire_all + (uint64_t)g * wire_stride) = make_uint4(magic, mask, 0, 0);
    *reinterpret_cast<uint4*>(wire_all + (uint64_t)g * wire_stride + exc_off) = make_uint4(0, 0, 0, 0);
}

int launch_narrow_levels(const uint32_t* d_counts, uint64_t counts_stride, int n_genomes, const NarrowSpec& spec,
                         uint8_t* d_wire, uint64_t wire_stride, uint64_t narrow_off, uint64_t exc_off, uint64_t small_off,
                         uint32_t exc_cap, uint32_t header_magic, uint32_t header_mask, cudaStream_t s) {
    if (n_genomes <= 0) return KMERML_OK;
    wire_header_kernel<<<(n_genomes + 63) / 64, 64, 0, s>>>(d_wire, wire_stride, exc_off, n_genomes, header_magic, header_mask);
    const uint64_t threads = (spec.total >> 4) + ((spec.small_total + 3) >> 2);
    if (threads)
        narrow_levels_kernel<<<dim3((unsigned)((threads + 255) / 256), (unsigned)n_genomes), 256, 0, s>>>(
            d_counts, counts_stride, spec, d_wire, wire_stride, narrow_off, exc_off, small_off, exc_cap);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

// ------------------------------------------------------------------ host side
void widen_u8_to_u32(const uint8_t* src, uint32_t* dst, size_t n) {
    size_t i = 0;
#if defined(__x86_64__)
    if ((((uintptr_t)dst) & 15) == 0) {
        const __m128i zero = _mm_setzero_si128();
        for (; i + 16 <= n; i += 16) {
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i lo = _mm_unpacklo_epi8(b, zero), hi = _mm_unpackhi_epi8(b, zero);
            __m128i* d = reinterpret_cast<__m128i*>(dst + i);
            _mm_stream_si128(d, _mm_unpacklo_epi16(lo, zero));          // non-temporal: the row is written once
            _mm_stream_si128(d + 1, _mm_unpackhi_epi16(lo, zero));
            _mm_stream_si128(d + 2, _mm_unpacklo_epi16(hi, zero));
            _mm_stream_si128(d + 3, _mm_unpackhi_epi16(hi, zero));
        }
        _mm_sfence();
    }
#endif
    for (; i < n; i++) dst[i] = src[i];
}

void widen_u4_to_u32(const uint8_t* src, uint32_t* dst, size_t n_bytes) {      // two bins per byte, low nibble first
    for (size_t i = 0; i < n_bytes; i++) {
        const uint32_t b = src[i];
        dst[2 * i] = b & 15u;
        dst[2 * i + 1] = b >> 4;
    }
}

class HostPool {
public:
    explicit HostPool(int n) {
        for (int i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void submit(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> g(m_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }
    int size() const { return (int)workers_.size(); }

private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

HostPool* host_pool_create(int n_threads) { return new (std::nothrow) HostPool(n_threads); }
void host_pool_destroy(HostPool* p) { delete p; }
void host_pool_submit(HostPool* p, std::function<void()> f) { p->submit(std::move(f)); }
int host_pool_size(const HostPool* p) { return p ? p->size() : 0; }

}  // namespace km
