// dense.cu -- sm_100a kernels of the dense (4^k histogram) counting path.
//
// Replaces the window loop of kmerml/kmers/generate.py:49-58 (reference tree).
// Kernels:
//   prologue_kernel   per genome: skip text before the first header line, zero stats
//   count_kernel<0>   k = 8..14: one RED.ADD per window into the L2-resident 4^k row
//   count_kernel<1>   k <= 7   : CTA-private shared-memory histogram, flushed once
//   count_kernel<2>   first-occurrence offsets (atomicMin), for the k{k}.txt writer
//   cascade_kernel    c_{j}[p] = sum_b c_{j+1}[4p+b] + tails_j[p], 5 levels per launch
//   finalize_kernel   optional canonical fold, frequency row, window totals
//
// Work decomposition: a CTA owns a slice (a few tiles) of one genome; a thread owns
// the windows that START in its 64-byte chunk of the tile (fasta_walk.cuh), so there
// is no carry between threads and no compaction pass: the FASTA bytes are read once.
#include "fasta_walk.cuh"
#include "internal.h"

namespace km {

// --------------------------------------------------------------------- sinks
struct GlobalSink {
    uint32_t* top;
    const LevelMap* lm;
    GenomeStats* st;
    uint32_t genome;
    unsigned n;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t) {
        // no return value: a REDG.ADD executed by the L2 slice that owns the bin
        asm volatile("red.global.add.u32 [%0], 1;" ::"l"(__cvta_generic_to_global(top + idx)) : "memory");
        n++;
    }
    __device__ __noinline__ void tail(int j, uint32_t idx) {
        atomicAdd(lm->ptr(genome, j) + idx, 1u);
        atomicAdd(&st->n_tail[j], 1ull);
    }
};

struct SmemSink {
    uint32_t sbase;                    // shared-window address of the CTA's histogram
    const LevelMap* lm;
    GenomeStats* st;
    uint32_t genome;
    unsigned n;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sbase + idx * 4u) : "memory");
        n++;
    }
    __device__ __noinline__ void tail(int j, uint32_t idx) {
        atomicAdd(lm->ptr(genome, j) + idx, 1u);
        atomicAdd(&st->n_tail[j], 1ull);
    }
};

struct FirstSink {
    uint32_t* first;
    uint64_t file_lo;
    unsigned n;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t pos) {
        atomicMin(first + idx, (uint32_t)(pos - file_lo));
    }
    __device__ __forceinline__ void tail(int, uint32_t) {}
};

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ bool any_byte_eq(const uint32_t (&w)[16], uint32_t pattern) {
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t x = w[i] ^ pattern;
        acc |= (x - 0x01010101u) & ~x & 0x80808080u;   // a zero byte in x
    }
    return acc != 0;
}

template <class Sink>
__device__ __forceinline__ void walk_chunk_regs(const Genome& g, uint64_t cs, const uint32_t (&w)[16],
                                                bool in_hdr, const DenseParams& P, Sink& sink) {
    WalkState s;
    s.kmer = 0; s.run = 0; s.in_hdr = in_hdr ? 1 : 0; s.pend = 0; s.rec_known = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            step_own(g, cs + (uint64_t)(4 * i + j), (w[i] >> (8 * j)) & 0xFFu, s, P, sink);
    }
    walk_overhang(g, cs + CHUNK, s, P, sink);
}

// ------------------------------------------------------------------ kernels
__global__ void prologue_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ offsets,
                                GenomeDev* gd, GenomeStats* st, int n) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint64_t lo = offsets[g], hi = offsets[g + 1];
    gd[g].file_lo = lo;
    gd[g].lo = first_header(buf, lo, hi);
    gd[g].hi = hi;
    st[g].total_top = 0;
    for (int j = 0; j < 16; j++) st[g].n_tail[j] = 0;
}

// MODE 0: global histogram, 1: shared histogram, 2: first occurrence
template <int MODE>
__global__ void __launch_bounds__(COUNT_THREADS)
count_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds,
             const Slice* __restrict__ slices, DenseParams P, LevelMap lm, GenomeStats* stats,
             uint32_t* first) {
    extern __shared__ __align__(16) uint32_t sh_hist[];
    __shared__ uint8_t flags[COUNT_THREADS];
    __shared__ unsigned long long carry[2];
    __shared__ unsigned long long sh_total;

    const int tid = threadIdx.x;
    const Slice sl = slices[blockIdx.x];
    const GenomeDev gd = gds[sl.genome];
    Genome g;
    g.b = buf;
    g.lo = gd.lo;
    g.hi = gd.hi;

    if (MODE == 1) {
        const int nb = 1 << (2 * P.k);
        for (int i = tid; i < nb; i += COUNT_THREADS) sh_hist[i] = 0;
    }
    if (tid == 0) {
        sh_total = 0;
        unsigned long long c = 0;
        uint64_t until;
        if (sl.begin > g.lo && pos_in_header(g, sl.begin, &until)) c = until;
        carry[0] = c;
        carry[1] = c;
    }

    GlobalSink gs;
    SmemSink ss;
    FirstSink fs;
    if (MODE == 0) { gs.top = lm.ptr(sl.genome, P.k); gs.lm = &lm; gs.st = stats + sl.genome; gs.genome = sl.genome; gs.n = 0; }
    if (MODE == 1) { ss.sbase = (uint32_t)__cvta_generic_to_shared(sh_hist); ss.lm = &lm; ss.st = stats + sl.genome; ss.genome = sl.genome; ss.n = 0; }
    if (MODE == 2) { fs.first = first; fs.file_lo = gd.file_lo; fs.n = 0; }

    const uint64_t end = sl.end < g.hi ? sl.end : g.hi;
    for (uint64_t tb = sl.begin; tb < end; tb += TILE_BYTES) {
        flags[tid] = 0;
        __syncthreads();                                    // flags cleared, carry[0] visible
        const uint64_t cb = tb + (uint64_t)tid * CHUNK;
        const uint64_t cs = cb > g.lo ? cb : g.lo;
        const uint64_t ce = cb + CHUNK < g.hi ? cb + CHUNK : g.hi;
        const bool has = cs < ce;
        const bool full = has && (ce - cs == CHUNK);
        uint32_t w[16];
        if (full) {
            const uint4* src = reinterpret_cast<const uint4*>(buf + cb);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 v = __ldg(src + i);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
        }
        // phase 1: header lines that start in my chunk shadow the chunks after it
        if (has && (!full || any_byte_eq(w, 0x3E3E3E3Eu))) {
            find_headers(g, cs, ce, [&](uint64_t, uint64_t until) {
                for (int j = tid + 1; j < COUNT_THREADS && tb + (uint64_t)j * CHUNK < until; j++) flags[j] = 1;
                atomicMax(&carry[1], (unsigned long long)until);
            });
        }
        __syncthreads();
        // phase 2: walk
        if (has) {
            const bool in_hdr = flags[tid] || cs < carry[0];
            if (MODE == 0) {
                if (full) walk_chunk_regs(g, cs, w, in_hdr, P, gs);
                else walk_chunk(g, cs, ce, in_hdr, P, gs, [&](uint64_t pos) -> uint32_t { return g.b[pos]; });
            } else if (MODE == 1) {
                if (full) walk_chunk_regs(g, cs, w, in_hdr, P, ss);
                else walk_chunk(g, cs, ce, in_hdr, P, ss, [&](uint64_t pos) -> uint32_t { return g.b[pos]; });
            } else {
                if (full) walk_chunk_regs(g, cs, w, in_hdr, P, fs);
                else walk_chunk(g, cs, ce, in_hdr, P, fs, [&](uint64_t pos) -> uint32_t { return g.b[pos]; });
            }
        }
        __syncthreads();
        if (tid == 0) carry[0] = carry[1];
    }

    if (MODE == 2) return;
    unsigned n = MODE == 0 ? gs.n : ss.n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((tid & 31) == 0 && n) atomicAdd(&sh_total, (unsigned long long)n);
    __syncthreads();
    if (MODE == 1) {
        uint32_t* top = lm.ptr(sl.genome, P.k);
        const int nb = 1 << (2 * P.k);
        for (int i = tid; i < nb; i += COUNT_THREADS) {
            uint32_t v = sh_hist[i];
            if (v) atomicAdd(top + i, v);
        }
    }
    if (tid == 0 && sh_total) atomicAdd(&stats[sl.genome].total_top, sh_total);
}

// Levels k_top-1 .. k_top-depth (depth <= 6) from level k_top (final) and the tails
// already accumulated in the lower levels.  Block b owns top-level bins
// [4096 b, 4096 b + 4096): each thread reads 16 of them (4 x 128-bit), produces its 4
// bins of level k_top-1 and 1 bin of level k_top-2 in registers; deeper levels go
// through a small shared-memory tree.
__global__ void __launch_bounds__(256)
cascade_kernel(LevelMap lm, int k_top, int depth, uint32_t genome0) {
    __shared__ uint32_t sh[256];
    const int tid = threadIdx.x;
    const uint32_t g = genome0 + blockIdx.y;
    const uint64_t n2 = k_top >= 2 ? 1ull << (2 * (k_top - 2)) : 0;   // bins at level k_top-2
    const uint64_t q2 = (uint64_t)blockIdx.x * 256 + tid;
    uint32_t w = 0;
    if (k_top >= 2 ? q2 < n2 : q2 == 0) {
        const uint4* hi = reinterpret_cast<const uint4*>(lm.ptr(g, k_top)) + 4 * q2;
        uint4 h0 = hi[0], h1 = hi[1], h2 = hi[2], h3 = hi[3];
        uint4* lo1 = reinterpret_cast<uint4*>(lm.ptr(g, k_top - 1)) + q2;
        uint4 t = *lo1;
        t.x += h0.x + h0.y + h0.z + h0.w;
        t.y += h1.x + h1.y + h1.z + h1.w;
        t.z += h2.x + h2.y + h2.z + h2.w;
        t.w += h3.x + h3.y + h3.z + h3.w;
        *lo1 = t;
        if (depth >= 2) {
            uint32_t* lo2 = lm.ptr(g, k_top - 2);
            w = t.x + t.y + t.z + t.w + lo2[q2];
            lo2[q2] = w;
        }
    }
    if (depth < 3) return;
    sh[tid] = w;
    __syncthreads();
    int width = 256;
    for (int d = 3; d <= depth; d++) {
        width >>= 2;
        const int level = k_top - d;
        uint32_t nv = 0;
        if (tid < width) {
            const uint64_t q = (uint64_t)blockIdx.x * width + tid;
            if (q < (1ull << (2 * level))) {
                uint32_t* lo = lm.ptr(g, level);
                nv = sh[4 * tid] + sh[4 * tid + 1] + sh[4 * tid + 2] + sh[4 * tid + 3] + lo[q];
                lo[q] = nv;
            }
        }
        __syncthreads();
        if (tid < width) sh[tid] = nv;
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t revcomp_code(uint32_t x, int k) {
    uint32_t v = __brev(~x);                                   // complement, reverse all bits
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);   // restore bit order inside each base
    return v >> (32 - 2 * k);
}

__device__ __forceinline__ void level_totals(const RowSpec& row, int k_top, const GenomeStats* stats, uint32_t g,
                                             unsigned long long* tot, uint64_t* totals, bool write) {
    const int tid = threadIdx.x;
    if (tid < row.nk) {
        const int j = row.k[tid];
        unsigned long long t = stats[g].total_top;
        for (int i = j; i < k_top; i++) t += stats[g].n_tail[i];
        tot[tid] = t;
        if (totals && write) totals[(uint64_t)g * row.nk + tid] = t;
    }
    __syncthreads();
}

// Frequency rows (count / windows) and window totals; 4 bins per thread, 128-bit I/O.
__global__ void __launch_bounds__(256)
finalize_kernel(LevelMap lm, RowSpec row, int k_top, const GenomeStats* __restrict__ stats,
                float* freq, uint64_t freq_stride, uint64_t* totals, uint32_t genome0) {
    __shared__ unsigned long long tot[16];
    const uint32_t g = genome0 + blockIdx.y;
    level_totals(row, k_top, stats, g, tot, totals, blockIdx.x == 0);
    if (!freq) return;
    const unsigned long long n4 = row.off[row.nk] >> 2;
    const unsigned long long v = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (v >= n4) return;
    const unsigned long long e = v << 2;
    int ki = 0;
    while (ki + 1 < row.nk && e >= row.off[ki + 1]) ki++;
    const uint4 c = *reinterpret_cast<const uint4*>(lm.counts + (uint64_t)g * lm.counts_stride + e);
    const double inv = tot[ki] ? 1.0 / (double)tot[ki] : 0.0;
    float4 f;
    f.x = (float)((double)c.x * inv);
    f.y = (float)((double)c.y * inv);
    f.z = (float)((double)c.z * inv);
    f.w = (float)((double)c.w * inv);
    *reinterpret_cast<float4*>(freq + (uint64_t)g * freq_stride + e) = f;
}

// Canonical mode: fold every requested level onto min(kmer, revcomp) in place, then
// normalise.  The thread of the smaller index of each {x, rc(x)} pair owns both bins.
__global__ void __launch_bounds__(256)
finalize_canonical_kernel(LevelMap lm, RowSpec row, int k_top, const GenomeStats* __restrict__ stats,
                          float* freq, uint64_t freq_stride, uint64_t* totals, uint32_t genome0) {
    __shared__ unsigned long long tot[16];
    const int tid = threadIdx.x;
    const uint32_t g = genome0 + blockIdx.y;
    level_totals(row, k_top, stats, g, tot, totals, blockIdx.x == 0);
    const unsigned long long n = row.off[row.nk];
    for (unsigned long long e = (unsigned long long)blockIdx.x * 256 + tid; e < n;
         e += (unsigned long long)gridDim.x * 256) {
        int ki = 0;
        while (ki + 1 < row.nk && e >= row.off[ki + 1]) ki++;
        const int j = row.k[ki];
        const uint32_t x = (uint32_t)(e - row.off[ki]);
        uint32_t* c = lm.ptr(g, j);
        float* f = freq ? freq + (uint64_t)g * freq_stride + row.off[ki] : nullptr;
        const double inv = tot[ki] ? 1.0 / (double)tot[ki] : 0.0;
        const uint32_t rc = revcomp_code(x, j);
        if (x < rc) {
            const uint32_t a = c[x] + c[rc];
            c[x] = a;
            c[rc] = 0;
            if (f) { f[x] = (float)((double)a * inv); f[rc] = 0.0f; }
        } else if (x == rc) {
            if (f) f[x] = (float)((double)c[x] * inv);
        }
    }
}

// ---------------------------------------------------------------- launchers
int dense_setup_attributes() {
    KM_CUDA(cudaFuncSetAttribute(count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (1 << (2 * SMEM_MAX_K)) * 4));
    return KMERML_OK;
}

int launch_prologue(const uint8_t* d_fasta, const uint64_t* d_offsets, GenomeDev* d_genomes,
                    GenomeStats* d_stats, int n_genomes, cudaStream_t s) {
    if (n_genomes <= 0) return KMERML_OK;
    prologue_kernel<<<(n_genomes + 127) / 128, 128, 0, s>>>(d_fasta, d_offsets, d_genomes, d_stats, n_genomes);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

static DenseParams make_params(int k, int min_rec, bool tails, int tail_lo) {
    DenseParams P;
    P.k = k;
    P.mask = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    P.min_rec = min_rec;
    P.tails = tails ? 1 : 0;
    P.tail_lo = tail_lo;
    return P;
}

int launch_count(const uint8_t* d_fasta, const GenomeDev* d_genomes, const Slice* d_slices, int n_slices,
                 int k, int k_bottom, int min_rec, bool use_smem, const LevelMap& lm, GenomeStats* d_stats,
                 cudaStream_t s) {
    if (n_slices <= 0) return KMERML_OK;
    DenseParams P = make_params(k, min_rec, k > k_bottom, k_bottom);
    if (use_smem) {
        size_t smem = (size_t)(1u << (2 * k)) * 4;
        count_kernel<1><<<n_slices, COUNT_THREADS, smem, s>>>(d_fasta, d_genomes, d_slices, P, lm, d_stats, nullptr);
    } else {
        count_kernel<0><<<n_slices, COUNT_THREADS, 0, s>>>(d_fasta, d_genomes, d_slices, P, lm, d_stats, nullptr);
    }
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

int launch_first_occurrence(const uint8_t* d_fasta, const GenomeDev* d_genomes, const Slice* d_slices,
                            int n_slices, int k, int min_rec, uint32_t* d_first, cudaStream_t s) {
    if (n_slices <= 0) return KMERML_OK;
    DenseParams P = make_params(k, min_rec, false, k);
    LevelMap lm = {};
    count_kernel<2><<<n_slices, COUNT_THREADS, 0, s>>>(d_fasta, d_genomes, d_slices, P, lm, nullptr, d_first);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

int launch_cascade(const LevelMap& lm, int k_top, int k_bottom, uint32_t genome0, int n_genomes,
                   cudaStream_t s) {
    int top = k_top;
    while (top > k_bottom) {
        int depth = top - k_bottom < 6 ? top - k_bottom : 6;
        uint64_t n2 = top >= 2 ? 1ull << (2 * (top - 2)) : 1;
        dim3 grid((unsigned)((n2 + 255) / 256), (unsigned)n_genomes);
        cascade_kernel<<<grid, 256, 0, s>>>(lm, top, depth, genome0);
        KM_CUDA(cudaGetLastError());
        top -= depth;
    }
    return KMERML_OK;
}

int cascade_launches(int k_top, int k_bottom) { return (k_top - k_bottom + 5) / 6; }

int launch_finalize(const LevelMap& lm, const RowSpec& row, int k_top, bool canonical,
                    const GenomeStats* d_stats, float* d_freq, uint64_t freq_stride, uint64_t* d_totals,
                    uint32_t genome0, int n_genomes, cudaStream_t s) {
    if (n_genomes <= 0) return KMERML_OK;
    unsigned long long n = row.off[row.nk];
    if (canonical) {
        unsigned gx = (unsigned)((n + 256ull * 8 - 1) / (256ull * 8));
        if (gx < 1) gx = 1;
        if (gx > 148u * 16u) gx = 148u * 16u;
        dim3 grid(gx, (unsigned)n_genomes);
        finalize_canonical_kernel<<<grid, 256, 0, s>>>(lm, row, k_top, d_stats, d_freq, freq_stride, d_totals, genome0);
    } else {
        unsigned gx = d_freq ? (unsigned)(((n >> 2) + 255) / 256) : 1u;
        dim3 grid(gx, (unsigned)n_genomes);
        finalize_kernel<<<grid, 256, 0, s>>>(lm, row, k_top, d_stats, d_freq, freq_stride, d_totals, genome0);
    }
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

}  // namespace km
