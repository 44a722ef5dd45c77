// featcsv.cu -- the per-k-mer feature CSV of kmerml/kmers/statistics.py:95-251 as text produced on the GPU.
//
// The reference reads k{k}.txt with pandas (the digit column becomes an integer: leading zeros, i.e. leading
// A's, are lost, statistics.py:261-271), decodes str(int) back to letters (0 A, 1 T, 2 C, 3 G, else N; :248-251),
// and computes per row gc_percent, base counts, presence flags, CpG count and ratio, Shannon entropy and a
// repeat flag (:149-240), then DataFrame.to_csv.  Every column after `count` depends only on the k-mer's
// COMPOSITION CLASS (length, A/C/G/T/N counts, CpG count, repeat flag): the host formats one suffix string per
// class that occurs (pandas' own float formatting, so the text is byte-identical) and the kernels here do the
// per-row work: parse the file's lines, classify, size, and write `<letters>,<count>,<suffix>\n`.
#include "internal.h"

namespace km {

// one thread per line [start, end): "<digits>\t<digits>", at most 18 digits each (int64 like pandas infers)
__global__ void __launch_bounds__(256)
parse_kmer_lines_kernel(const uint8_t* __restrict__ text, const long long* __restrict__ line_end, uint64_t n_lines,
                        long long* __restrict__ value, long long* __restrict__ count, unsigned int* __restrict__ bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint64_t p = i ? (uint64_t)line_end[i - 1] + 1 : 0;
    const uint64_t e = (uint64_t)line_end[i];
    long long v = 0, c = 0;
    int nd = 0, nc = 0;
    bool ok = true;
    for (; p < e && text[p] != 9; p++, nd++) {
        const uint32_t d = (uint32_t)text[p] - 48u;
        ok &= d <= 9u;
        v = v * 10 + (long long)d;
    }
    ok &= nd >= 1 && nd <= 18 && p < e;
    for (p++; p < e; p++, nc++) {
        const uint32_t d = (uint32_t)text[p] - 48u;
        ok &= d <= 9u;
        c = c * 10 + (long long)d;
    }
    ok &= nc >= 1 && nc <= 18;
    value[i] = v;
    count[i] = c;
    if (!ok) atomicOr(bad, 1u);
}

// decimal digits of v, most significant first; returns how many (str(int): "0" for 0)
__device__ __forceinline__ int decimal_digits(long long v, uint8_t* d) {
    uint8_t tmp[20];
    int n = 0;
    unsigned long long u = (unsigned long long)v;
    do {
        tmp[n++] = (uint8_t)(u % 10ull);
        u /= 10ull;
    } while (u);
    for (int i = 0; i < n; i++) d[i] = tmp[n - 1 - i];
    return n;
}

// composition class of str(v) decoded 0 A, 1 T, 2 C, 3 G, else N:
// key = n | A << 5 | C << 10 | G << 15 | T << 20 | N << 25 | cpg << 30 | repeat << 35
__global__ void __launch_bounds__(256)
feature_keys_kernel(const long long* __restrict__ value, uint64_t n_rows, long long* __restrict__ keys) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint8_t d[20];
    const int n = decimal_digits(value[i], d);
    unsigned long long na = 0, nt = 0, nc = 0, ng = 0, nn = 0, cpg = 0, rep = 0;
    for (int j = 0; j < n; j++) {
        const uint8_t x = d[j];
        na += x == 0; nt += x == 1; nc += x == 2; ng += x == 3; nn += x >= 4;
        if (j + 1 < n && x == 2 && d[j + 1] == 3) cpg++;
    }
    for (int j = 0; j + 3 < n; j++) {
        const uint8_t a0 = d[j] > 4 ? 4 : d[j], a1 = d[j + 1] > 4 ? 4 : d[j + 1];
        const uint8_t b0 = d[j + 2] > 4 ? 4 : d[j + 2], b1 = d[j + 3] > 4 ? 4 : d[j + 3];
        if (a0 == b0 && a1 == b1) rep = 1;
    }
    keys[i] = (long long)((unsigned long long)n | na << 5 | nc << 10 | ng << 15 | nt << 20 | nn << 25 | cpg << 30 | rep << 35);
}

__device__ __forceinline__ int count_digits(long long c) {
    int n = 1;
    for (unsigned long long u = (unsigned long long)c; u >= 10ull; u /= 10ull) n++;
    return n;
}

// bytes of "<letters>,<count>,<suffix>\n"
__global__ void __launch_bounds__(256)
feature_line_len_kernel(const long long* __restrict__ value, const long long* __restrict__ count,
                        const long long* __restrict__ cls, const int* __restrict__ suffix_len, uint64_t n_rows,
                        long long* __restrict__ len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    len[i] = (long long)count_digits(value[i]) + 1 + count_digits(count[i]) + 1 + suffix_len[cls[i]] + 1;
}

__global__ void __launch_bounds__(256)
feature_write_kernel(const long long* __restrict__ value, const long long* __restrict__ count,
                     const long long* __restrict__ cls, const long long* __restrict__ suffix_off,
                     const int* __restrict__ suffix_len, const uint8_t* __restrict__ suffix_text,
                     const long long* __restrict__ line_off, uint64_t n_rows, uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint8_t* o = out + line_off[i];
    uint8_t d[20];
    int n = decimal_digits(value[i], d);
    const uint32_t letters = 0x47435441u;                       // "ATCG"[digit]
    for (int j = 0; j < n; j++) *o++ = d[j] < 4 ? (uint8_t)(letters >> (8 * d[j])) : (uint8_t)'N';
    *o++ = ',';
    n = decimal_digits(count[i], d);
    for (int j = 0; j < n; j++) *o++ = (uint8_t)(48 + d[j]);
    *o++ = ',';
    const long long c = cls[i];
    const uint8_t* s = suffix_text + suffix_off[c];
    const int sl = suffix_len[c];
    for (int j = 0; j < sl; j++) *o++ = s[j];
    *o = '\n';
}

static unsigned grid_for(uint64_t n) { return (unsigned)((n + 255) / 256); }

int launch_parse_kmer_lines(const uint8_t* d_text, const long long* d_line_end, uint64_t n_lines, long long* d_value,
                            long long* d_count, unsigned int* d_bad, cudaStream_t s) {
    KM_CUDA(cudaMemsetAsync(d_bad, 0, 4, s));
    if (!n_lines) return KMERML_OK;
    parse_kmer_lines_kernel<<<grid_for(n_lines), 256, 0, s>>>(d_text, d_line_end, n_lines, d_value, d_count, d_bad);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

int launch_feature_keys(const long long* d_value, uint64_t n_rows, long long* d_keys, cudaStream_t s) {
    if (!n_rows) return KMERML_OK;
    feature_keys_kernel<<<grid_for(n_rows), 256, 0, s>>>(d_value, n_rows, d_keys);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

int launch_feature_line_len(const long long* d_value, const long long* d_count, const long long* d_cls,
                            const int* d_suffix_len, uint64_t n_rows, long long* d_len, cudaStream_t s) {
    if (!n_rows) return KMERML_OK;
    feature_line_len_kernel<<<grid_for(n_rows), 256, 0, s>>>(d_value, d_count, d_cls, d_suffix_len, n_rows, d_len);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

int launch_feature_write(const long long* d_value, const long long* d_count, const long long* d_cls,
                         const long long* d_suffix_off, const int* d_suffix_len, const uint8_t* d_suffix_text,
                         const long long* d_line_off, uint64_t n_rows, uint8_t* d_out, cudaStream_t s) {
    if (!n_rows) return KMERML_OK;
    feature_write_kernel<<<grid_for(n_rows), 256, 0, s>>>(d_value, d_count, d_cls, d_suffix_off, d_suffix_len, d_suffix_text,
                                                          d_line_off, n_rows, d_out);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

}  // namespace km
