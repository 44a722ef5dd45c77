// format.cu -- the text of a k{k}.txt file, produced on the GPU.
//
// The reference writes one line per observed k-mer, "<digits>\t<count>\n" with the digits A0 T1 C2 G3,
// in dict insertion order = order of first occurrence (kmerml/kmers/generate.py:68-91).  From a dense
// count row and the first-occurrence offsets (kmerml_first_occurrence) that is: select the non-zero
// bins, sort them by first offset, size the lines, prefix-sum, write.  The select / sort / scan are CUB
// calls from the toolkit; the sparse path (k > 14) already holds its k-mers in order and only formats.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "internal.h"

namespace km {

struct NonZeroBin {
    const uint32_t* counts;
    __device__ __forceinline__ bool operator()(uint32_t bin) const { return counts[bin] != 0u; }
};
struct NonZeroOne {
    const uint32_t* counts;
    __host__ __device__ __forceinline__ unsigned long long operator()(uint32_t bin) const { return counts[bin] != 0u ? 1ull : 0ull; }
};
using BinIter = thrust::counting_iterator<uint32_t>;
using FlagIter = thrust::transform_iterator<NonZeroOne, BinIter>;

__device__ __forceinline__ uint32_t revcomp_bin(uint32_t bin, int k) {
    uint32_t rc = 0;
    for (int i = 0; i < k; i++) {
        rc = (rc << 2) | (3u - (bin & 3u));
        bin >>= 2;
    }
    return rc;
}

__global__ void first_keys_kernel(const uint32_t* __restrict__ bins, uint64_t n, const uint32_t* __restrict__ first,
                                  int k, int canonical, uint32_t* keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = bins[i];
    uint32_t f = first[b];
    if (canonical) {                       // a canonical bin first appears where either strand's k-mer first does
        const uint32_t g = first[revcomp_bin(b, k)];
        f = g < f ? g : f;
    }
    keys[i] = f;
}

__device__ __forceinline__ int decimal_digits(uint32_t v) {
    int d = 1;
    if (v >= 10u) d = 2;
    if (v >= 100u) d = 3;
    if (v >= 1000u) d = 4;
    if (v >= 10000u) d = 5;
    if (v >= 100000u) d = 6;
    if (v >= 1000000u) d = 7;
    if (v >= 10000000u) d = 8;
    if (v >= 100000000u) d = 9;
    if (v >= 1000000000u) d = 10;
    return d;
}

// counts: indexed by line (gather == nullptr) or by the line's bin (dense rows)
template <class Code>
__global__ void line_len_kernel(const Code* __restrict__ codes, const uint32_t* __restrict__ counts, int by_code,
                                uint64_t n, int k, unsigned long long* lens) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { lens[i] = 0; return; }                  // the scan's last element = total length
    const uint32_t c = by_code ? counts[(uint64_t)codes[i]] : counts[i];
    lens[i] = (unsigned long long)(k + 2 + decimal_digits(c));
}

template <class Code>
__global__ void write_lines_kernel(const Code* __restrict__ codes, const uint32_t* __restrict__ counts, int by_code,
                                   uint64_t n, int k, const unsigned long long* __restrict__ offsets, uint8_t* text) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Code code = codes[i];
    uint32_t c = by_code ? counts[(uint64_t)code] : counts[i];
    uint8_t* p = text + offsets[i];
    // lexicographic code (A0 C1 G2 T3, first base most significant) -> the file's digits A0 T1 C2 G3
    for (int j = 0; j < k; j++) {
        const uint32_t base = (uint32_t)(code >> (2 * (k - 1 - j))) & 3u;
        p[j] = (uint8_t)((0x31333230u >> (8u * base)) & 0xFFu);        // "0231"[base]
    }
    p[k] = (uint8_t)'\t';
    const int nd = decimal_digits(c);
    for (int j = nd - 1; j >= 0; j--) {
        p[k + 1 + j] = (uint8_t)('0' + c % 10u);
        c /= 10u;
    }
    p[k + 1 + nd] = (uint8_t)'\n';
}

struct FormatWork {
    uint32_t* bins_a;
    uint32_t* bins_b;
    uint32_t* keys_a;
    uint32_t* keys_b;
    unsigned long long* lens;        // n + 1
    unsigned long long* offsets;     // n + 1
    unsigned long long* n_selected;
    void* temp;
    size_t temp_bytes;
};

static size_t format_layout(uint64_t n_bins, uint64_t max_lines, FormatWork* w, uint8_t* base) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += (bytes + 255) / 256 * 256;
        return p;
    };
    FormatWork l;
    l.bins_a = (uint32_t*)take(max_lines * 4);
    l.bins_b = (uint32_t*)take(max_lines * 4);
    l.keys_a = (uint32_t*)take(max_lines * 4);
    l.keys_b = (uint32_t*)take(max_lines * 4);
    l.lens = (unsigned long long*)take((max_lines + 1) * 8);
    l.offsets = (unsigned long long*)take((max_lines + 1) * 8);
    l.n_selected = (unsigned long long*)take(256);
    size_t t1 = 0, t2 = 0, t3 = 0;
    BinIter it(0);
    NonZeroBin pred{nullptr};
    cub::DeviceSelect::If(nullptr, t1, it, (uint32_t*)nullptr, (unsigned long long*)nullptr, (uint64_t)n_bins, pred);
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, dk, dv, (uint64_t)max_lines, 0, 32);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (uint64_t)(max_lines + 1));
    size_t t4 = 0;
    FlagIter flags(it, NonZeroOne{nullptr});
    cub::DeviceReduce::Sum(nullptr, t4, flags, (unsigned long long*)nullptr, (uint64_t)n_bins);
    l.temp_bytes = std::max(std::max(t1, t4), std::max(t2, t3));
    l.temp = take(l.temp_bytes);
    if (w) *w = l;
    return off;
}

size_t format_workspace_bytes(uint64_t n_bins, uint64_t max_lines) { return format_layout(n_bins, max_lines, nullptr, nullptr); }

template <class Code>
static int format_lines(const FormatWork& w, const Code* d_codes, const uint32_t* d_counts, int by_code, uint64_t n,
                        int k, uint8_t* d_text, uint64_t text_cap, uint64_t* h_len, cudaStream_t s) {
    *h_len = 0;
    if (!n) return KMERML_OK;
    const unsigned blocks = (unsigned)((n + 1 + 255) / 256);
    line_len_kernel<Code><<<blocks, 256, 0, s>>>(d_codes, d_counts, by_code, n, k, w.lens);
    KM_CUDA(cudaGetLastError());
    size_t tb = w.temp_bytes;
    KM_CUDA(cub::DeviceScan::ExclusiveSum(w.temp, tb, w.lens, w.offsets, (uint64_t)(n + 1), s));
    unsigned long long total = 0;
    KM_CUDA(cudaMemcpyAsync(&total, w.offsets + n, 8, cudaMemcpyDeviceToHost, s));
    KM_CUDA(cudaStreamSynchronize(s));
    *h_len = total;
    if (total > text_cap) return KMERML_OK;                // caller re-sizes and calls again
    write_lines_kernel<Code><<<blocks, 256, 0, s>>>(d_codes, d_counts, by_code, n, k, w.offsets, d_text);
    KM_CUDA(cudaGetLastError());
    KM_CUDA(cudaStreamSynchronize(s));
    return KMERML_OK;
}

int run_format_dense(void* workspace, int k, const uint32_t* d_counts, const uint32_t* d_first, bool canonical,
                     uint64_t max_lines, uint8_t* d_text, uint64_t text_cap, uint64_t* h_len, uint64_t* h_lines,
                     cudaStream_t s) {
    const uint64_t n_bins = 1ull << (2 * k);
    FormatWork w;
    format_layout(n_bins, max_lines, &w, (uint8_t*)workspace);
    *h_len = 0;
    *h_lines = 0;
    // the observed bins, ascending ...
    size_t tb = w.temp_bytes;
    BinIter it(0);
    NonZeroBin pred{d_counts};
    // (a count row with more non-zero bins than max_lines would overrun bins_a: count first)
    KM_CUDA(cudaMemsetAsync(w.n_selected, 0, 8, s));
    {
        // cub::DeviceSelect writes at most as many items as it selects; bound them by a first counting pass
        size_t tr = w.temp_bytes;
        FlagIter flags(it, NonZeroOne{d_counts});
        KM_CUDA(cub::DeviceReduce::Sum(w.temp, tr, flags, w.n_selected, (uint64_t)n_bins, s));
    }
    unsigned long long n = 0;
    KM_CUDA(cudaMemcpyAsync(&n, w.n_selected, 8, cudaMemcpyDeviceToHost, s));
    KM_CUDA(cudaStreamSynchronize(s));
    *h_lines = n;
    if (n > max_lines) return KMERML_OK;                   // caller re-sizes and calls again
    if (!n) return KMERML_OK;
    KM_CUDA(cub::DeviceSelect::If(w.temp, tb, it, w.bins_a, w.n_selected, (uint64_t)n_bins, pred, s));
    // ... ordered by the offset of their first window
    first_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w.bins_a, n, d_first, k, canonical ? 1 : 0, w.keys_a);
    KM_CUDA(cudaGetLastError());
    cub::DoubleBuffer<uint32_t> dk(w.keys_a, w.keys_b), dv(w.bins_a, w.bins_b);
    tb = w.temp_bytes;
    KM_CUDA(cub::DeviceRadixSort::SortPairs(w.temp, tb, dk, dv, (uint64_t)n, 0, 32, s));
    return format_lines<uint32_t>(w, dv.Current(), d_counts, 1, n, k, d_text, text_cap, h_len, s);
}

int run_format_lines(void* workspace, int k, const uint64_t* d_codes, const uint32_t* d_counts, uint64_t n,
                     uint8_t* d_text, uint64_t text_cap, uint64_t* h_len, cudaStream_t s) {
    FormatWork w;
    format_layout(1, n, &w, (uint8_t*)workspace);
    return format_lines<uint64_t>(w, d_codes, d_counts, 0, n, k, d_text, text_cap, h_len, s);
}

}  // namespace km
