// stats.cu -- reductions over count rows that the reference computes with pandas.
//
//   count_stats     kmerml/utils/kmer_metadata.py:59-78: total / unique / max / min / median of the OBSERVED
//                   k-mers' counts of one k (the reference reads them back from k{k}.txt; here the 4^k row is
//                   already in HBM).  The median is an exact radix select (3 passes of 11 + 11 + 10 bits over the
//                   row, a 2048-bin histogram each), not a sort.
//   column_stats    per feature column of an organisms x k-mers matrix: organisms with a non-zero entry, mean and
//                   population variance (float64) -- what filter_features(min_prevalence, min_variance) and
//                   get_top_features(n, "variance") of the aspirational API (tests/test_ml.py:9-12) need.
#include "internal.h"

namespace km {

struct SelectState {                 // device-side state of one radix select
    unsigned long long rank;         // rank still to find inside the current prefix class
    uint32_t prefix;                 // bits fixed so far (high bits)
    uint32_t pad;
};

struct CountStatsOut {               // matches the uint64[8] the C-ABI returns
    unsigned long long total, unique, max, min, med_lo, med_hi, pad0, pad1;
};

__global__ void __launch_bounds__(256)
count_basic_kernel(const uint32_t* __restrict__ c, uint64_t n, CountStatsOut* out) {
    unsigned long long total = 0, uniq = 0;
    uint32_t mx = 0, mn = 0xFFFFFFFFu;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = __ldg(c + i);
        if (v) {
            total += v;
            uniq++;
            mx = max(mx, v);
            mn = min(mn, v);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        total += __shfl_xor_sync(0xffffffffu, total, o);
        uniq += __shfl_xor_sync(0xffffffffu, uniq, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0 && uniq) {
        atomicAdd(&out->total, total);
        atomicAdd(&out->unique, uniq);
        atomicMax(&out->max, (unsigned long long)mx);
        atomicMin(&out->min, (unsigned long long)mn);
    }
}

// ranks of the two middle elements among the `unique` observed counts (equal when that number is odd)
__global__ void select_init_kernel(const CountStatsOut* out, SelectState* st) {
    const unsigned long long u = out->unique;
    st[0].rank = u ? (u - 1) / 2 : 0;
    st[1].rank = u ? u / 2 : 0;
    st[0].prefix = st[1].prefix = 0;
}

// histogram of the next digit over the observed counts whose higher bits equal the prefix, for both selects
template <int SHIFT, int BITS, int HI_SHIFT>
__global__ void __launch_bounds__(256)
select_hist_kernel(const uint32_t* __restrict__ c, uint64_t n, const SelectState* __restrict__ st, unsigned int* hist) {
    __shared__ unsigned int sh[2][1 << BITS];
    for (int i = threadIdx.x; i < 2 << BITS; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t p0 = st[0].prefix, p1 = st[1].prefix;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = __ldg(c + i);
        if (!v) continue;
        const uint32_t hi = HI_SHIFT >= 32 ? 0u : (v >> (HI_SHIFT & 31));
        const uint32_t d = (v >> SHIFT) & ((1u << BITS) - 1u);
        if (hi == (HI_SHIFT >= 32 ? 0u : (p0 >> (HI_SHIFT & 31)))) atomicAdd(&sh[0][d], 1u);
        if (hi == (HI_SHIFT >= 32 ? 0u : (p1 >> (HI_SHIFT & 31)))) atomicAdd(&sh[1][d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 << BITS; i += 256) {
        const unsigned int v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

// one block: find the digit class that holds the rank, fix its bits, clear the histogram for the next pass
template <int SHIFT, int BITS>
__global__ void __launch_bounds__(32)
select_step_kernel(SelectState* st, unsigned int* hist, CountStatsOut* out, int last) {
    const int which = threadIdx.x;               // lanes 0 and 1 walk their histogram serially (2048 entries)
    if (which < 2) {
        unsigned long long r = st[which].rank;
        const unsigned int* h = hist + (which << BITS);
        uint32_t d = 0;
        for (; d < (1u << BITS) - 1u; d++) {
            if (r < h[d]) break;
            r -= h[d];
        }
        st[which].rank = r;
        st[which].prefix |= d << SHIFT;
        if (last) {
            if (which == 0) out->med_lo = out->unique ? st[0].prefix : 0;
            else out->med_hi = out->unique ? st[1].prefix : 0;
        }
    }
    __syncwarp();
    for (int i = threadIdx.x; i < 2 << BITS; i += 32) hist[i] = 0;
}

size_t count_stats_workspace() { return 256 + 2 * 2048 * 4; }

// d_out: uint64[8] = total, unique, max, min, lower median, upper median (of the non-zero bins), 0, 0
int launch_count_stats(const uint32_t* d_counts, uint64_t n_bins, void* workspace, unsigned long long* d_out, cudaStream_t s) {
    CountStatsOut* out = reinterpret_cast<CountStatsOut*>(d_out);
    SelectState* st = reinterpret_cast<SelectState*>(workspace);
    unsigned int* hist = reinterpret_cast<unsigned int*>((uint8_t*)workspace + 256);
    CountStatsOut init;
    memset(&init, 0, sizeof(init));
    init.min = ~0ull;
    KM_CUDA(cudaMemcpyAsync(out, &init, sizeof(init), cudaMemcpyHostToDevice, s));       // (pageable: copied at once)
    KM_CUDA(cudaMemsetAsync(hist, 0, 2 * 2048 * 4, s));
    const unsigned grid = (unsigned)std::min<uint64_t>((n_bins + 255) / 256, 148u * 8u);
    if (!n_bins) return KMERML_OK;
    count_basic_kernel<<<grid, 256, 0, s>>>(d_counts, n_bins, out);
    select_init_kernel<<<1, 1, 0, s>>>(out, st);
    select_hist_kernel<21, 11, 32><<<grid, 256, 0, s>>>(d_counts, n_bins, st, hist);
    select_step_kernel<21, 11><<<1, 32, 0, s>>>(st, hist, out, 0);
    select_hist_kernel<10, 11, 21><<<grid, 256, 0, s>>>(d_counts, n_bins, st, hist);
    select_step_kernel<10, 11><<<1, 32, 0, s>>>(st, hist, out, 0);
    select_hist_kernel<0, 10, 10><<<grid, 256, 0, s>>>(d_counts, n_bins, st, hist);
    select_step_kernel<0, 10><<<1, 32, 0, s>>>(st, hist, out, 1);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

// One thread per column (consecutive threads read consecutive columns of a row: coalesced).  Two passes over
// the rows: the sum (exact for counts), then the squared deviations in float64.
template <class T>
__global__ void __launch_bounds__(256)
column_stats_kernel(const T* __restrict__ x, uint64_t stride, int n_rows, uint64_t m, uint32_t* __restrict__ nnz,
                    double* __restrict__ mean, double* __restrict__ var) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double sum = 0.0;
    uint32_t nz = 0;
    for (int r = 0; r < n_rows; r++) {
        const T v = x[(uint64_t)r * stride + j];
        sum += (double)v;
        nz += v != (T)0;
    }
    const double mu = n_rows ? sum / (double)n_rows : 0.0;
    double ss = 0.0;
    for (int r = 0; r < n_rows; r++) {
        const double d = (double)x[(uint64_t)r * stride + j] - mu;
        ss += d * d;
    }
    nnz[j] = nz;
    mean[j] = mu;
    var[j] = n_rows ? ss / (double)n_rows : 0.0;
}

int launch_column_stats(const void* d_x, int dtype, uint64_t stride, int n_rows, uint64_t m, uint32_t* d_nnz,
                        double* d_mean, double* d_var, cudaStream_t s) {
    if (!m) return KMERML_OK;
    const unsigned grid = (unsigned)((m + 255) / 256);
    if (dtype == 0) column_stats_kernel<float><<<grid, 256, 0, s>>>((const float*)d_x, stride, n_rows, m, d_nnz, d_mean, d_var);
    else if (dtype == 1) column_stats_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)d_x, stride, n_rows, m, d_nnz, d_mean, d_var);
    else column_stats_kernel<double><<<grid, 256, 0, s>>>((const double*)d_x, stride, n_rows, m, d_nnz, d_mean, d_var);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

}  // namespace km
