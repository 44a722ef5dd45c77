// dense.cu -- sm_100a kernels of the dense (4^k histogram) counting path.
//
// Replaces the window loop of kmerml/kmers/generate.py:49-58 (reference tree).
// Kernels:
//   prologue_kernel            per genome: skip text before the first header line, zero stats
//   slice_header_kernel        per slice: does it start inside a header line? (bounded look-back) ...
//   slice_long_scan/_resolve   ... and the same for lines too long for that (unwrapped FASTA)
//   count_kernel<1>            k <= 7   : CTA-private shared-memory histogram, flushed once
//   count8_kernel              k = 8    : packed 16-bit shared histogram, 1024 threads, swept every tile
//   partition_kernel<S>        k = 9..12: every window into its bucket's slot (shared atomics), slots written
//                                         bucket-major; run-end tails into per-genome lists
//   bucket_kernel              k = 9..12: per (bucket, genome) a 16384-bin shared histogram of its slots ->
//                                         count row, frequencies and the bucket's cascade subtree
//   overflow_kernel, tails_apply_kernel, tails_rescan_kernel   what slots / lists could not hold
//   count_kernel<0>            k = 13, 14: one RED.ADD per window into the L2-resident 4^k row
//   count_kernel<2>            first-occurrence offsets (atomicMin), for the k{k}.txt writer
//   cascade_kernel             c_{j}[p] = sum_b c_{j+1}[4p+b] + tails_j[p], 6 levels per launch
//   finalize_kernel / _low / _canonical    canonical fold, frequency rows, window totals
//
// Work decomposition: a CTA owns a slice (a run of tiles) of one genome; a thread owns the windows whose
// LAST base lies in its 32-byte chunk of the tile (fasta_walk.cuh), so there is no carry between threads
// and no compaction pass: the FASTA bytes are read once.
#include <algorithm>

#include "fasta_walk.cuh"
#include "internal.h"

namespace km {

// --------------------------------------------------------------------- sinks
// The hot sinks hold only what count() touches so that they live in registers; the
// rare run-end tails go through DevTails (passed by reference to non-inlined code).
struct DevTails {
    const LevelMap* lm;
    GenomeStats* st;
    uint32_t genome;
    __device__ __noinline__ void tail(int j, uint32_t idx) const {
        atomicAdd(lm->ptr(genome, j) + idx, 1u);
        atomicAdd(&st->n_tail[j], 1ull);
    }
};

struct NoTails {
    __device__ __forceinline__ void tail(int, uint32_t) const {}
};

// Partition path: the levels the bucket kernel writes (>= k_stop) are never zeroed and never read back,
// so their run-end tails are kept in a per-genome list (level << 32 | index) and added after the
// bucket kernel has stored the rows.  A list that runs full is dropped as a whole: the genome's run
// ends are then walked a second time (tails_rescan_kernel).
struct TailList {
    unsigned long long* list;          // all genomes of the group, see tail_list_of()
    unsigned int* counts;              // per genome (global index): entries pushed, may exceed the capacity
    unsigned int* any_full;            // set when some genome's list ran full
    uint64_t batch_lo;                 // buffer offset of the group's first genome
    uint32_t genome0;                  // ... and its index
};
__device__ __forceinline__ unsigned long long* tail_list_of(const TailList& tl, const GenomeDev& gd, uint32_t g,
                                                            uint32_t* cap) {
    *cap = (uint32_t)((gd.hi - gd.file_lo) >> 6) + 1024u;
    return tl.list + ((gd.file_lo - tl.batch_lo) >> 6) + 1024ull * (g - tl.genome0);
}
struct ListTails {
    const LevelMap* lm;
    GenomeStats* st;
    uint32_t genome;
    int k_stop;
    unsigned long long* list;
    unsigned int* count;
    unsigned int* any_full;
    uint32_t cap;
    __device__ __noinline__ void tail(int j, uint32_t idx) const {
        atomicAdd(&st->n_tail[j], 1ull);
        if (j < k_stop) {                                   // the few small levels below the bucket subtrees
            atomicAdd(lm->ptr(genome, j) + idx, 1u);
            return;
        }
        const unsigned int slot = atomicAdd(count, 1u);
        if (slot < cap) list[slot] = ((unsigned long long)j << 32) | idx;
        else *any_full = 1u;
    }
};

struct GlobalSink {
    uint32_t* top;
    unsigned n;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t) {
        // no return value: a REDG.ADD executed by the L2 slice that owns the bin
        asm volatile("red.global.add.u32 [%0], 1;" ::"l"(__cvta_generic_to_global(top + idx)) : "memory");
        n++;
    }
    __device__ __forceinline__ void count4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pa, uint64_t pb,
                                           uint64_t pc, uint64_t pd) {
        count(a, pa); count(b, pb); count(c, pc); count(d, pd);
    }
    __device__ __forceinline__ void count8(const uint32_t* w, const uint64_t* p) {
#pragma unroll
        for (int i = 0; i < 8; i++) count(w[i], p[i]);
    }
    __device__ __forceinline__ void count8_tail(const uint32_t* w, const uint64_t* p, bool last) {
#pragma unroll
        for (int i = 0; i < 7; i++) count(w[i], p[i]);
        if (last) count(w[7], p[7]);
    }
};

struct SmemSink {
    uint32_t sbase;                    // shared-window address of the CTA's histogram
    unsigned n;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sbase + idx * 4u) : "memory");
        n++;
    }
    __device__ __forceinline__ void count4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pa, uint64_t pb,
                                           uint64_t pc, uint64_t pd) {
        count(a, pa); count(b, pb); count(c, pc); count(d, pd);
    }
    __device__ __forceinline__ void count8(const uint32_t* w, const uint64_t* p) {
#pragma unroll
        for (int i = 0; i < 8; i++) count(w[i], p[i]);
    }
    __device__ __forceinline__ void count8_tail(const uint32_t* w, const uint64_t* p, bool last) {
#pragma unroll
        for (int i = 0; i < 7; i++) count(w[i], p[i]);
        if (last) count(w[7], p[7]);
    }
};

// k = 8: 65536 bins as packed 16-bit halves of 32768 shared words (128 KB), counted with non-returning
// shared atomics like the 32-bit histograms of k <= 7.  The histogram leaves room for one CTA per SM, so
// that CTA has 1024 threads (count8_kernel: 32 KB tiles).  Word i serves the bins i and i + 32768: its LOW half
// counts the windows of BOTH, its high half those of bin i + 32768 -- the addend is 1 + 0x10000 * (bit 15 of the
// window), two instructions (AND, multiply-add) instead of the select a "1 or 0x10000" addend compiles to.
// Exactness: a tile has at most 32768 windows, and after every tile the CTA sweeps the histogram and moves every
// word whose low half has reached 0x4000 to the global row (sweep_packed16), so a low half never exceeds
// 0x3FFF + 32768 = 0xBFFF and the high half never exceeds the low one: no carry for any input (tested with a
// homopolymer).
struct Packed16Sink {
    static constexpr bool raw_windows = true;       // emit_clean passes unmasked funnel words: the masks below do it
    uint32_t sbase;                    // shared-window address of the 32768 words
    unsigned n;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t) {
        uint32_t addend;
        asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(addend) : "r"(idx & 0x8000u));
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(sbase + ((idx & 0x7FFFu) << 2)), "r"(addend) : "memory");
        n++;
    }
    __device__ __forceinline__ void count4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pa, uint64_t pb,
                                           uint64_t pc, uint64_t pd) {
        count(a, pa); count(b, pb); count(c, pc); count(d, pd);
    }
    __device__ __forceinline__ void count8(const uint32_t* w, const uint64_t* p) {
#pragma unroll
        for (int i = 0; i < 8; i++) count(w[i], p[i]);
    }
    __device__ __forceinline__ void count8_tail(const uint32_t* w, const uint64_t* p, bool last) {
#pragma unroll
        for (int i = 0; i < 7; i++) count(w[i], p[i]);
        if (last) count(w[7], p[7]);
    }
};

// word -> the two bins it serves
__device__ __forceinline__ void packed16_flush_word(uint32_t v, uint32_t i, uint32_t* row) {
    const uint32_t hi = v >> 16, lo = (v & 0xFFFFu) - hi;
    if (lo) atomicAdd(row + i, lo);
    if (hi) atomicAdd(row + i + 32768u, hi);
}

// (whole CTA, between the barrier that ends a tile and the one that precedes the next tile's counting)
template <int NT>
__device__ __forceinline__ void sweep_packed16(uint32_t* hist, uint32_t* row) {
    uint4* h4 = reinterpret_cast<uint4*>(hist);
#pragma unroll
    for (int i = threadIdx.x; i < 32768 / 4; i += NT) {                     // (32768 / 4 / NT loads in flight together)
        const uint4 v = h4[i];
        if (((v.x | v.y | v.z | v.w) & 0x0000C000u) == 0u) continue;        // every low half below 0x4000
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (w[j] & 0xC000u) { packed16_flush_word(w[j], 4u * i + j, row); w[j] = 0; }
        h4[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

struct FirstSink {
    uint32_t* first;
    uint64_t file_lo;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t pos) {
        atomicMin(first + idx, (uint32_t)(pos - file_lo));
    }
    __device__ __forceinline__ void count4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pa, uint64_t pb,
                                           uint64_t pc, uint64_t pd) {
        count(a, pa); count(b, pb); count(c, pc); count(d, pd);
    }
    __device__ __forceinline__ void count8(const uint32_t* w, const uint64_t* p) {
#pragma unroll
        for (int i = 0; i < 8; i++) count(w[i], p[i]);
    }
    __device__ __forceinline__ void count8_tail(const uint32_t* w, const uint64_t* p, bool last) {
#pragma unroll
        for (int i = 0; i < 7; i++) count(w[i], p[i]);
        if (last) count(w[7], p[7]);
    }
};

// kmerml_encode: the walk with k = 1 -- every base that counts writes its 2-bit code at its byte position
struct EncodeSink {
    uint8_t* sym;
    uint64_t file_lo;
    __device__ __forceinline__ void count(uint32_t idx, uint64_t pos) { sym[pos - file_lo] = (uint8_t)idx; }
    __device__ __forceinline__ void count4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pa, uint64_t pb,
                                           uint64_t pc, uint64_t pd) {
        count(a, pa); count(b, pb); count(c, pc); count(d, pd);
    }
    __device__ __forceinline__ void count8(const uint32_t* w, const uint64_t* p) {
#pragma unroll
        for (int i = 0; i < 8; i++) count(w[i], p[i]);
    }
    __device__ __forceinline__ void count8_tail(const uint32_t* w, const uint64_t* p, bool last) {
#pragma unroll
        for (int i = 0; i < 7; i++) count(w[i], p[i]);
        if (last) count(w[7], p[7]);
    }
};

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ void level_totals(const RowSpec& row, int k_top, const GenomeStats* stats, uint32_t g,
                                             unsigned long long* tot, uint64_t* totals, bool write);

struct TileCtx {                      // shared-memory state of the tile loop
    uint8_t* flags;                   // [2][COUNT_THREADS] chunk starts inside a header line (by tile parity)
    uint8_t* clean;                   // [COUNT_THREADS] chunk is clean (bases + at most one '\n')
    uint32_t* last16;                 // [COUNT_THREADS] last 16 bases of a clean chunk
    unsigned long long* carry;        // [2] end of a header line that runs into later tiles
    uint32_t* prev_tile;              // [2] {ok, last16} of the previous tile's last chunk
};

#define KM_TILE_SMEM(prefix) KM_TILE_SMEM_N(prefix, COUNT_THREADS)
#define KM_TILE_SMEM_N(prefix, NT)                             \
    __shared__ uint8_t prefix##_flags[2 * (NT)];               \
    __shared__ uint8_t prefix##_clean[(NT)];                   \
    __shared__ uint32_t prefix##_last16[(NT)];                 \
    __shared__ unsigned long long prefix##_carry[2];           \
    __shared__ uint32_t prefix##_prev[2];                      \
    TileCtx tc;                                                \
    tc.flags = prefix##_flags; tc.clean = prefix##_clean; tc.last16 = prefix##_last16; \
    tc.carry = prefix##_carry; tc.prev_tile = prefix##_prev;

// The tile loop every counting kernel shares.  Per tile and thread: two 128-bit loads
// of the 32-byte chunk, SWAR classification, bit-compaction of clean chunks and
// header-line detection (phase 1); then either the funnel-shift emission of a clean
// chunk whose left neighbour is clean too, or the generic byte walker (phase 2).
// EARLY_LOAD: when the next tile's chunk is requested -- at the start of the current tile (a true
// prefetch; needs 8 registers across the whole tile: the histogram kernels have them) or right before
// the tile's closing barrier (the partition kernel, whose placement code needs every register: there
// the early prefetch was spilled to local memory at once, measured).
template <bool EARLY_LOAD, int NT = COUNT_THREADS, class Sink, class Tails, class PerTile>
__device__ __forceinline__ void walk_slice(const uint8_t* __restrict__ buf, const Genome& g, const Slice& sl,
                                           const DenseParams& P, Sink& sink, const Tails& tails, const TileCtx& tc,
                                           PerTile&& per_tile) {
    // Two barriers per tile.  What makes that safe:
    //  * flags[] is double-buffered by tile parity; a thread clears its entry of the OTHER buffer during
    //    phase 2, i.e. after that buffer's last readers (phase 2 of the tile before) passed a barrier and
    //    before its next writers (phase 1 of the next tile) can start;
    //  * carry[parity] only ever grows (atomicMax of header ends); every thread folds it into its own
    //    running maximum `hc` after the barrier that closes phase 1;
    //  * clean[] / last16[] are written in phase 1 and read by the right neighbour in phase 2.
    // slices are whole numbers of 16 KB tiles; a CTA with larger tiles must clip its chunks to the slice end
    constexpr bool CLIP = NT * CHUNK > TILE_BYTES;
    const int tid = threadIdx.x;
    tc.flags[tid] = 0;
    tc.flags[NT + tid] = 0;
    if (tid == 0) {
        tc.carry[0] = 0;
        tc.carry[1] = 0;
        tc.prev_tile[0] = sl.prev_ok;                       // the chunk before the slice (another CTA's)
        tc.prev_tile[1] = sl.prev16;
    }
    __syncthreads();
    unsigned long long hc = sl.hdr_until;                   // resolved per slice by slice_header_kernel
    const uint64_t end = sl.end < g.hi ? sl.end : g.hi;
    auto load_chunk = [&](uint64_t tbx, uint32_t* w) -> bool {
        const uint64_t cbx = tbx + (uint64_t)tid * CHUNK;
        if (!(cbx >= g.lo && cbx + CHUNK <= g.hi && (CLIP ? cbx : tbx) < end)) return false;
        const uint4* src = reinterpret_cast<const uint4*>(buf + cbx);
#pragma unroll
        for (int i = 0; i < CHUNK / 16; i++) {
            uint4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
        return true;
    };
    uint32_t w[CHUNK / 4], wn[CHUNK / 4];
    bool full = load_chunk(sl.begin, w);
    uint32_t tile_no = 0;
    for (uint64_t tb = sl.begin; tb < end; tb += (NT * CHUNK), tile_no++) {
        uint8_t* flags = tc.flags + (tile_no & 1u) * NT;
        bool full_next = false;
        if (EARLY_LOAD) full_next = load_chunk(tb + (NT * CHUNK), wn);
        const uint64_t cb = tb + (uint64_t)tid * CHUNK;
        const uint64_t cs = cb > g.lo ? cb : g.lo;
        const uint64_t ce = cb + CHUNK < g.hi ? cb + CHUNK : g.hi;
        const bool has = cs < ce && (!CLIP || cb < end);
        CleanChunk cc;
        cc.hi = cc.lo = 0; cc.n = 0; cc.nl = 32; cc.last16 = 0;
        bool clean = false;
        auto on_header = [&](uint64_t, uint64_t until) {
            for (int j = tid + 1; j < NT && tb + (uint64_t)j * CHUNK < until; j++) flags[j] = 1;
            atomicMax(&tc.carry[tile_no & 1u], (unsigned long long)until);
        };
        if (full) {
            clean = classify_pack(w, cc);                           // only bases and at most one '\n' ?
            // only chunks that hold a '>' can start a header line
            if (!clean && any_byte_eq_chunk(w, 0x3E3E3E3Eu)) find_headers(g, cs, ce, on_header);
        } else if (has) {
            find_headers(g, cs, ce, on_header);
        }
        tc.clean[tid] = clean ? 1 : 0;
        tc.last16[tid] = cc.last16;
        __syncthreads();
        // phase 2
        const bool in_hdr = flags[tid] || cs < hc;
        bool prev_ok;
        uint32_t carry16;
        if (tid > 0) {
            prev_ok = tc.clean[tid - 1] && !flags[tid - 1] && !(cb - CHUNK < hc);
            carry16 = tc.last16[tid - 1];
        } else {
            prev_ok = tc.prev_tile[0] != 0;
            carry16 = tc.prev_tile[1];
        }
        {
            const unsigned long long seen = tc.carry[tile_no & 1u];
            hc = seen > hc ? seen : hc;                     // header lines that run into later tiles
        }
        tc.flags[((tile_no & 1u) ^ 1u) * NT + tid] = 0;
        if (has) {
            if (clean && !in_hdr && prev_ok && P.min_rec <= P.k) {
                emit_clean(cc, carry16, cs, P, sink);
                if (ce == g.hi && P.tails) run_end_event(g, g.hi, P, tails);
            } else {
                walk_chunk(g, cs, ce, in_hdr, P, sink, tails, [&](uint64_t pos) -> uint32_t { return g.b[pos]; });
            }
        }
        if (!EARLY_LOAD) full_next = load_chunk(tb + (NT * CHUNK), wn);    // in flight during the barrier and the hook
        __syncthreads();
        if (tid == NT - 1) {
            tc.prev_tile[0] = (has && clean && !in_hdr) ? 1u : 0u;
            tc.prev_tile[1] = cc.last16;
        }
        per_tile(tile_no);
        full = full_next;
#pragma unroll
        for (int i = 0; i < CHUNK / 4; i++) w[i] = wn[i];
    }
}

__device__ __forceinline__ unsigned long long block_sum_u32(unsigned n, unsigned long long* sh_total) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(sh_total, (unsigned long long)n);
    __syncthreads();
    return *sh_total;
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

// Per slice: does its first byte lie inside a header line that started earlier?  (One
// thread per slice, so the serial backward scans overlap instead of stalling a CTA.)
__global__ void slice_header_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds,
                                    Slice* slices, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Slice sl = slices[i];
    const GenomeDev gd = gds[sl.genome];
    Genome g;
    g.b = buf;
    g.lo = gd.lo;
    g.hi = gd.hi;
    SliceHead h;
    slice_head_quick(g, sl.begin, LINE_SCAN_LIMIT, &h);
    slices[i].hdr_until = h.hdr_until;
    slices[i].prev_ok = h.prev_ok;
    slices[i].prev16 = h.prev16;
    slices[i].line_start = h.line_start;
    slices[i].scan_last = LS_UNRESOLVED;
    if (h.line_start == LS_UNRESOLVED) slices[n].genome = 1;     // scratch entry: the long-line passes have work
}

// Long lines (unwrapped FASTA): a slice whose bounded look-back found no line terminator scans the
// bytes between the previous slice's start and its own, cooperatively and backwards, for the last one.
__global__ void __launch_bounds__(256)
slice_long_scan_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds, Slice* slices, int n) {
    if (slices[n].genome == 0) return;
    const int i = blockIdx.x;
    const Slice sl = slices[i];
    if (sl.line_start != LS_UNRESOLVED) return;
    const GenomeDev gd = gds[sl.genome];
    uint64_t lo = gd.lo;
    if (i > 0 && slices[i - 1].genome == sl.genome && slices[i - 1].begin > lo) lo = slices[i - 1].begin;
    __shared__ unsigned long long s_best;                 // 1 + position of the last terminator seen, 0 = none
    if (threadIdx.x == 0) s_best = 0;
    __syncthreads();
    constexpr uint64_t WINDOW = 256 * 64;
    for (uint64_t wend = sl.begin; wend > lo; wend = wend > WINDOW ? wend - WINDOW : 0) {
        const uint64_t hi = wend - (uint64_t)threadIdx.x * 64 > wend ? 0 : wend - (uint64_t)threadIdx.x * 64;
        unsigned long long best = 0;
        if (hi >= 64 && hi > lo) {                        // the thread's 64 bytes [hi - 64, hi)
            const uint4* src = reinterpret_cast<const uint4*>(buf + hi - 64);
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = __ldg(src + j);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(v);
#pragma unroll
            for (int j = 15; j >= 0; j--) {
                const uint32_t x = w[j];
                const uint32_t a = x ^ 0x0A0A0A0Au, c = x ^ 0x0D0D0D0Du;
                const uint32_t z = (((a - 0x01010101u) & ~a) | ((c - 0x01010101u) & ~c)) & 0x80808080u;
                if (z && !best) {
                    // exact test per byte, highest first (the zero-byte trick can flag the byte above a match)
                    for (int bb = 3; bb >= 0 && !best; bb--) {
                        const uint32_t ch = (x >> (8 * bb)) & 0xFFu;
                        const uint64_t pos = hi - 64 + 4 * (uint64_t)j + (uint64_t)bb;
                        if (is_term(ch) && pos >= lo) best = pos + 1;
                    }
                }
            }
        }
        if (best) atomicMax(&s_best, best);
        if (__syncthreads_or(best != 0)) break;
    }
    if (threadIdx.x == 0) slices[i].scan_last = s_best ? (uint64_t)s_best : (lo == gd.lo ? gd.lo : LS_UNRESOLVED);
}

// ... and the line start of every open slice is the running maximum of the line starts known so far
// (the table is in buffer order).  One CTA scans the table; open slices then get their header verdict.
__global__ void __launch_bounds__(1024)
slice_long_resolve_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds, Slice* slices, int n) {
    if (slices[n].genome == 0) return;
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        unsigned long long v = 0;
        Slice sl;
        if (i < n) {
            sl = slices[i];
            const uint64_t cand = sl.line_start != LS_UNRESOLVED ? sl.line_start : sl.scan_last;
            v = cand == LS_UNRESOLVED ? 0ull : (unsigned long long)cand;
        }
        // inclusive prefix maximum over the block, then over the chunks seen so far
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o && u > v) v = u;
        }
        if (lane == 31) s_warp[wid] = v;
        __syncthreads();
        unsigned long long before = s_carry;
        for (int w = 0; w < wid; w++) before = s_warp[w] > before ? s_warp[w] : before;
        if (before > v) v = before;
        __syncthreads();
        if (tid == 1023) s_carry = v;
        if (i < n && sl.line_start == LS_UNRESOLVED) {
            const GenomeDev gd = gds[sl.genome];
            Genome g;
            g.b = buf;
            g.lo = gd.lo;
            g.hi = gd.hi;
            const uint64_t ls = v > gd.lo ? (uint64_t)v : gd.lo;
            const uint64_t until = header_until_from(g, sl.begin, ls);
            slices[i].hdr_until = until;
            if (until) slices[i].prev_ok = 0;
        }
        __syncthreads();
    }
}

// MODE 0: global histogram, 1: shared histogram, 2: first occurrence, 3: symbols (kmerml_encode)  (k = 8: count8_kernel below)
template <int MODE>
__global__ void __launch_bounds__(COUNT_THREADS)
count_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds,
             const Slice* __restrict__ slices, DenseParams P, LevelMap lm, GenomeStats* stats,
             uint32_t* first) {
    extern __shared__ __align__(16) uint32_t sh_hist[];
    KM_TILE_SMEM(ck)
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
    if (tid == 0) sh_total = 0;

    unsigned n = 0;
    if (MODE == 2) {
        FirstSink sink;
        sink.first = first; sink.file_lo = gd.file_lo;
        NoTails nt;
        walk_slice<true>(buf, g, sl, P, sink, nt, tc, [](uint32_t) {});
        return;
    }
    if (MODE == 3) {
        EncodeSink sink;
        sink.sym = reinterpret_cast<uint8_t*>(first); sink.file_lo = gd.file_lo;
        NoTails nt;
        walk_slice<true>(buf, g, sl, P, sink, nt, tc, [](uint32_t) {});
        return;
    }
    DevTails tails;
    tails.lm = &lm; tails.st = stats + sl.genome; tails.genome = sl.genome;
    if (MODE == 0) {
        GlobalSink sink;
        sink.top = lm.ptr(sl.genome, P.k); sink.n = 0;
        walk_slice<true>(buf, g, sl, P, sink, tails, tc, [](uint32_t) {});
        n = sink.n;
    } else {
        SmemSink sink;
        sink.sbase = (uint32_t)__cvta_generic_to_shared(sh_hist); sink.n = 0;
        walk_slice<true>(buf, g, sl, P, sink, tails, tc, [](uint32_t) {});
        n = sink.n;
    }
    const unsigned long long total = block_sum_u32(n, &sh_total);
    if (MODE == 1) {
        uint32_t* top = lm.ptr(sl.genome, P.k);
        const int nb = 1 << (2 * P.k);
        for (int i = tid; i < nb; i += COUNT_THREADS) {
            uint32_t v = sh_hist[i];
            if (v) atomicAdd(top + i, v);
        }
    }
    if (tid == 0 && total) atomicAdd(&stats[sl.genome].total_top, total);
}

// k = 8 (see Packed16Sink): one CTA of 1024 threads per SM, 32 KB tiles.
constexpr int COUNT8_THREADS = 1024;
__global__ void __launch_bounds__(COUNT8_THREADS, 1)
count8_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds, const Slice* __restrict__ slices,
              DenseParams P, LevelMap lm, GenomeStats* stats) {
    extern __shared__ __align__(16) uint32_t sh_hist[];
    KM_TILE_SMEM_N(c8, COUNT8_THREADS)
    __shared__ unsigned long long sh_total;
    const int tid = threadIdx.x;
    const Slice sl = slices[blockIdx.x];
    const GenomeDev gd = gds[sl.genome];
    Genome g;
    g.b = buf;
    g.lo = gd.lo;
    g.hi = gd.hi;
    for (int i = tid; i < 32768; i += COUNT8_THREADS) sh_hist[i] = 0;
    if (tid == 0) sh_total = 0;
    DevTails tails;
    tails.lm = &lm; tails.st = stats + sl.genome; tails.genome = sl.genome;
    Packed16Sink sink;
    sink.sbase = (uint32_t)__cvta_generic_to_shared(sh_hist); sink.n = 0;
    uint32_t* row8 = lm.ptr(sl.genome, 8);
    walk_slice<true, COUNT8_THREADS>(buf, g, sl, P, sink, tails, tc,
                                     [&](uint32_t) { sweep_packed16<COUNT8_THREADS>(sh_hist, row8); });
    const unsigned long long total = block_sum_u32(sink.n, &sh_total);
    for (int i = tid; i < 32768; i += COUNT8_THREADS) packed16_flush_word(sh_hist[i], (uint32_t)i, row8);
    if (tid == 0 && total) atomicAdd(&stats[sl.genome].total_top, total);
}

// ---------------------------------------------------------------- partition path
// k = 9..12: the 4^k histogram (up to 64 MB) is too big for shared memory, and L2 atomics top out
// near 190 G/s (measured) and thrash the L2 next to the streamed input, so the windows are first
// partitioned by their leading k-7 bases into nb = 4^(k-7) buckets of 16384 bins:
//   partition_kernel  one CTA per RUN of consecutive 16 KB tiles of one genome: every window is placed
//                     with ONE returning shared atomic into its bucket's slot of a 64 KB staging buffer
//                     (slot = 32768/nb uint16 14-bit payloads).  After every tile the FULL 32-byte
//                     sectors (16 payloads) of every slot are appended to the bucket's region of this
//                     run in HBM and the < 16 left-over payloads move to the front of the slot, so only
//                     live payloads travel: 2 bytes per window plus one padded sector per (bucket, run).
//                     Regions: [run][bucket][cap sectors], cap = 2 x the mean; sectors written per
//                     (run, bucket) go to a small table the bucket kernel reads.
//   bucket_kernel     one CTA per (bucket, genome): streams the regions of its bucket (coalesced 128-bit
//                     loads, one warp per region) into a 16384-bin shared histogram, then writes that
//                     64 KB slice of the count row, its frequencies and the bucket's whole cascade
//                     subtree (levels k .. k-7)
//   overflow_kernel   the rare windows that found their slot or region full (skewed genomes) are kept
//                     as plain k-mer indices and added afterwards
// Only shared-memory atomics (2.5 T/s measured on B200) sit on the hot path and every global
// access is a full-sector vector stream.
constexpr int PART_LOW = 7;
constexpr int PART_BINS = 1 << (2 * PART_LOW);        // 16384 bins per bucket
constexpr int PART_MAX_BUCKETS = 1024;                // k <= 12
constexpr int PART_SECTOR = 16;                       // payloads per 32-byte sector

// Rare for a genome of mixed sequence: the slot is full.  Tandem repeats make it common (every window of a
// tile falls into a handful of buckets), so the lanes that arrive together reserve their entries with one
// atomic on the genome's cursor instead of one each.
__device__ __noinline__ void overflow_push(uint32_t* ov, unsigned int* ov_count, uint32_t idx) {
    const unsigned act = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(act) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(ov_count, (unsigned int)__popc(act));
    base = __shfl_sync(act, base, leader);
    ov[base + __popc(act & ((1u << lane) - 1u))] = idx;
}

// Payload = the low 16 bits of the k-mer index: its 14 low-order bits (the bin inside the bucket) and, in bits
// 15:14, the two low-order bits of the bucket number.  The partition kernel stores it unmasked (one integer
// op less per window); the bucket kernel, which knows its bucket, clears the two bits with one XOR per pair of
// payloads.  A padding payload carries (bucket & 3) ^ 1 there, so that after the XOR it reads
// PART_BINS + spread: a dummy bin behind the histogram.
__device__ __forceinline__ uint32_t payload_fix(uint32_t bucket) { return ((bucket & 3u) << 14) * 0x00010001u; }
__device__ __forceinline__ uint32_t payload_pad(uint32_t bucket, uint32_t spread) {
    return (((bucket & 3u) ^ 1u) << 14) | (spread & 255u);
}

// (a whole sector whose region is full: sixteen payloads straight from the staging buffer)
__device__ __noinline__ void overflow_push_sector(uint32_t* ov, unsigned int* ov_count, uint32_t bucket,
                                                  const uint16_t* src) {
    for (int i = 0; i < PART_SECTOR; i++)
        overflow_push(ov, ov_count, (bucket << (2 * PART_LOW)) | (src[i] & (uint32_t)(PART_BINS - 1)));
}

// Slot geometry per bucket count: nb = 4^(k-7) buckets share the staging buffer.  For k <= 11 a slot is
// 32768 / nb payloads (twice the mean of a tile).  For k = 12 (1024 buckets, 16 payloads per tile on
// average) a slot also has to hold up to 15 payloads left over from the tile before, so it is 48 entries
// (96 KB of staging: two CTAs still share an SM): P(15 + Poisson(16) > 48) < 1e-4, whereas 32-entry slots
// overflowed in 7 % of all (bucket, tile) pairs (measured: the overflow path took 15 ms per C2 step).
template <int NB_SHIFT>
struct PartGeom {
    static constexpr uint32_t nb = 1u << NB_SHIFT;
    static constexpr uint32_t slot_size = NB_SHIFT == 10 ? 48u : (32768u >> NB_SHIFT);
    static constexpr uint32_t stage_entries = nb * slot_size;
    static constexpr bool owned = NB_SHIFT == 10;              // flush: one thread owns a whole slot
};

template <int NB_SHIFT>
struct SlotSink {
    static constexpr bool raw_windows = true;                  // emit_clean passes unmasked funnel words
    static constexpr uint32_t slot_size = PartGeom<NB_SHIFT>::slot_size;
    static constexpr uint32_t b4_mask = (PartGeom<NB_SHIFT>::nb - 1u) << 2;
    uint32_t cnt_s;                    // shared-window byte address of cnt[nb]
    uint32_t staged_s;                 //                            ... of staged[nb * slot_size]
    uint32_t* ov;                      // this genome's overflow list
    unsigned int* ov_count;
    uint32_t kmask;                    // 4^k - 1
    unsigned n;                        // windows seen by this thread (placed or overflowed)
    __device__ __forceinline__ uint32_t atom_inc(uint32_t b4) {
        uint32_t old;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt_s + b4) : "memory");
        return old;
    }
    __device__ __forceinline__ void store16(uint32_t b4, uint32_t pos, uint32_t x) {
        // entry (b, pos) at byte 2 * (b * slot + pos) = 2 * (b4 * (slot / 4) + pos): one IMAD, one add
        const uint32_t e = b4 * (slot_size / 4u) + pos;
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(staged_s + 2u * e), "h"((unsigned short)x) : "memory");
    }
    // (generic byte walker: idx is masked)
    __device__ __forceinline__ void count(uint32_t idx, uint64_t) {
        const uint32_t b4 = (idx >> (2 * PART_LOW - 2)) & b4_mask;
        const uint32_t pos = atom_inc(b4);
        n++;
        if (pos < slot_size) store16(b4, pos, idx);
        else overflow_push(ov, ov_count, idx);
    }
    __device__ __forceinline__ void count4(uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, uint64_t, uint64_t,
                                           uint64_t, uint64_t) {
        count(i0, 0); count(i1, 0); count(i2, 0); count(i3, 0);
    }
    __device__ __forceinline__ void count8(const uint32_t* w, const uint64_t*) { place<8>(w, true); }
    __device__ __forceinline__ void count8_tail(const uint32_t* w, const uint64_t*, bool last) { place<8>(w, last); }
    // N unmasked windows: all atomics are issued before the first returned position is needed.
    // `last_live` = false: the N-th window does not exist (a chunk of 31 bases).
    template <int N>
    __device__ __forceinline__ void place(const uint32_t* x, bool last_live = true) {
        uint32_t b4[N], pos[N];
        bool full = false;
#pragma unroll
        for (int u = 0; u < N; u++) b4[u] = (x[u] >> (2 * PART_LOW - 2)) & b4_mask;
#pragma unroll
        for (int u = 0; u < N; u++) pos[u] = (u < N - 1 || last_live) ? atom_inc(b4[u]) : 0u;
        n += last_live ? N : N - 1;
#pragma unroll
        for (int u = 0; u < N; u++) {
            const bool fits = pos[u] < slot_size;
            full |= !fits;
            if (fits && (u < N - 1 || last_live)) store16(b4[u], pos[u], x[u]);
        }
        if (full) {                                                              // rare: some slot is full
#pragma unroll
            for (int u = 0; u < N; u++)
                if (pos[u] >= slot_size) overflow_push(ov, ov_count, x[u] & kmask);
        }
    }
};

struct PartWalkSmem {                           // only live during the walk ...
    uint32_t last16[COUNT_THREADS];
    uint8_t flags[2 * COUNT_THREADS];
    uint8_t clean[COUNT_THREADS];
};

template <int NB_SHIFT>
struct PartSmem {
    uint16_t staged[PartGeom<NB_SHIFT>::stage_entries];    // 64 KB (96 KB for k = 12): nb slots
    uint32_t cnt[PART_MAX_BUCKETS];             // payloads in the slot (beyond the slot size: they overflowed)
    uint16_t written[PART_MAX_BUCKETS];         // sectors of this run already appended to the bucket's region
    PartWalkSmem walk;
    unsigned long long carry[2];
    uint32_t prev_tile[2];
    unsigned long long sh_total;
};
static_assert(sizeof(PartSmem<10>) <= 113 * 1024, "two partition CTAs must fit in one SM's shared memory");

struct GenomeRuns {                    // runs (= partition CTAs) of one genome inside the group's run list
    uint32_t run0, n_runs;
};

// One 32-byte sector with one 256-bit store (sm_100: STG.256; a lane writes a whole sector).
__device__ __forceinline__ void store_sector(uint4* dst, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// k = 12, rare: the slot ran over, or the bucket's region is about to: sector by sector, with every check.
template <int NB_SHIFT>
__device__ __noinline__ void flush_slot_slow(PartSmem<NB_SHIFT>& sm, uint32_t b, uint4* region, uint32_t cap,
                                             uint32_t* ov, unsigned int* ov_count) {
    constexpr uint32_t slot_size = PartGeom<NB_SHIFT>::slot_size;
    const uint32_t c = min(sm.cnt[b], slot_size);
    const uint32_t nfull = c / PART_SECTOR, left = c % PART_SECTOR;
    uint16_t* slot16 = sm.staged + (size_t)b * slot_size;
    uint4* slot = reinterpret_cast<uint4*>(slot16);
    uint32_t wr = sm.written[b];
    for (uint32_t q = 0; q < nfull; q++) {
        if (wr < cap) {
            region[((size_t)b * cap + wr) * 2] = slot[2 * q];
            region[((size_t)b * cap + wr) * 2 + 1] = slot[2 * q + 1];
            wr++;
        } else {
            overflow_push_sector(ov, ov_count, b, slot16 + q * PART_SECTOR);
        }
    }
    if (nfull && left) {
        const uint4 a = slot[2 * nfull], d = slot[2 * nfull + 1];
        slot[0] = a;
        slot[1] = d;
    }
    sm.written[b] = (uint16_t)wr;
    sm.cnt[b] = left;
}

template <int NB_SHIFT>
__global__ void __launch_bounds__(COUNT_THREADS, 2)
partition_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds,
                 const Slice* __restrict__ runs, DenseParams P, LevelMap lm,
                 GenomeStats* stats, uint4* __restrict__ payload, uint16_t* __restrict__ nsec, uint32_t cap,
                 uint32_t* __restrict__ overflow, unsigned int* __restrict__ ov_counts, uint64_t batch_lo,
                 TailList tl, int k_stop) {
    using G = PartGeom<NB_SHIFT>;
    constexpr uint32_t slot_size = G::slot_size;
    constexpr uint32_t nb = G::nb;
    constexpr uint32_t VPB = slot_size / 8;                                 // uint4 vectors per slot
    extern __shared__ __align__(16) unsigned char part_smem_raw[];
    PartSmem<NB_SHIFT>& sm = *reinterpret_cast<PartSmem<NB_SHIFT>*>(part_smem_raw);
    const int tid = threadIdx.x;
    const Slice sl = runs[blockIdx.x];
    const GenomeDev gd = gds[sl.genome];
    Genome g;
    g.b = buf;
    g.lo = gd.lo;
    g.hi = gd.hi;
    TileCtx tc;
    tc.flags = sm.walk.flags; tc.clean = sm.walk.clean; tc.last16 = sm.walk.last16;
    tc.carry = sm.carry; tc.prev_tile = sm.prev_tile;
    for (uint32_t i = tid; i < nb; i += COUNT_THREADS) { sm.cnt[i] = 0; sm.written[i] = 0; }
    if (tid == 0) sm.sh_total = 0;

    SlotSink<NB_SHIFT> sink;
    sink.cnt_s = (uint32_t)__cvta_generic_to_shared(sm.cnt);
    sink.staged_s = (uint32_t)__cvta_generic_to_shared(sm.staged);
    sink.ov = overflow + (gd.file_lo - batch_lo);      // one possible window per byte of the genome
    sink.ov_count = ov_counts + sl.genome;
    sink.kmask = P.mask;
    sink.n = 0;
    ListTails tails;
    tails.lm = &lm; tails.st = stats + sl.genome; tails.genome = sl.genome; tails.k_stop = k_stop;
    tails.list = tail_list_of(tl, gd, sl.genome, &tails.cap);
    tails.count = tl.counts + sl.genome;
    tails.any_full = tl.any_full;
    // this run's regions: bucket b at [(run * nb + b) * cap, ... + cap) sectors (two uint4 each);
    // nb * cap * 2 <= 2^18 vectors per run, so offsets inside a run fit 32 bits
    uint4* const region = payload + (size_t)blockIdx.x * nb * cap * 2;
    // (walk_slice starts with a __syncthreads)
    walk_slice<false>(buf, g, sl, P, sink, tails, tc, [&](uint32_t) {
        // (the walk of the tile ended with a __syncthreads; the next tile's placements start after
        // another one.)  Full sectors go to the bucket's region, the sector with the left-over payloads
        // moves to the front of the slot.
        uint4* st = reinterpret_cast<uint4*>(sm.staged);
        if (G::owned) {
            // k = 12: thread t owns slots t and t + 512 (six vectors each).  Nearly always the slot holds
            // 16..31 payloads: one sector out, the second one to the front.
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t b = (uint32_t)tid + h * COUNT_THREADS;
                const uint32_t c = sm.cnt[b];
                if (c < PART_SECTOR) continue;
                const uint32_t wr = sm.written[b];
                if (c > slot_size || wr + 3u > cap) {
                    flush_slot_slow<NB_SHIFT>(sm, b, region, cap, sink.ov, sink.ov_count);
                    continue;
                }
                // Slots are 96 bytes apart, so the lanes of a quarter warp that all read vector 0 of their slot meet
                // two by two in the same banks (6 t mod 8 is even).  Lanes 4..7 of every eight therefore take a
                // sector's two vectors in the other order: (6 t + [t & 4 ? 1 : 0]) mod 8 is a permutation.
                const uint32_t sw = ((uint32_t)tid >> 2) & 1u, sx = sw ^ 1u;
                uint4* slot = st + b * VPB;
                uint4* dst = region + (b * cap + wr) * 2u;
                {
                    const uint4 x0 = slot[sw], x1 = slot[sx];
                    dst[sw] = x0;
                    dst[sx] = x1;
                }
                uint32_t mv = 2;                                   // vectors 2, 3 hold the left-over payloads ...
                if (c >= 2 * PART_SECTOR) {
                    const uint4 x0 = slot[2 + sw], x1 = slot[2 + sx];
                    dst[2 + sw] = x0;
                    dst[2 + sx] = x1;
                    mv = 4;                                        // ... or 4, 5 ...
                    if (c == 3 * PART_SECTOR) {
                        const uint4 y0 = slot[4 + sw], y1 = slot[4 + sx];
                        dst[4 + sw] = y0;
                        dst[4 + sx] = y1;
                        mv = 0;                                    // ... or nothing is left
                    }
                }
                const uint4 m0 = slot[mv + sw], m1 = slot[mv + sx];
                slot[sw] = m0;
                slot[sx] = m1;
                sm.written[b] = (uint16_t)(wr + c / PART_SECTOR);
                sm.cnt[b] = c % PART_SECTOR;
            }
        } else {
            // k <= 11: vector v of the staging buffer belongs to bucket v / VPB; the lanes that hold one
            // bucket's vectors sit in one warp (nb = 256) or in one round of the CTA
#pragma unroll
            for (int j = 0; j < (int)(G::stage_entries / 8 / COUNT_THREADS); j++) {
                const uint32_t v = (uint32_t)tid + j * COUNT_THREADS;
                const uint32_t b = v / VPB, jv = v % VPB, q = jv >> 1;
                const uint32_t c = min(sm.cnt[b], slot_size);
                const uint32_t nfull = c / PART_SECTOR;
                const uint32_t wr = sm.written[b];
                const uint4 x = st[v];
                if (VPB <= 32) __syncwarp(); else __syncthreads();
                if (q < nfull) {
                    const uint32_t sec = wr + q;
                    if (sec < cap) region[(b * cap + sec) * 2u + (jv & 1u)] = x;
                    else if (!(jv & 1u)) overflow_push_sector(sink.ov, sink.ov_count, b, sm.staged + b * slot_size + q * PART_SECTOR);
                }
                if (VPB <= 32) __syncwarp(); else __syncthreads();     // (the overflow path reads the slot)
                if (q == nfull && nfull && (c % PART_SECTOR)) st[b * VPB + (jv & 1u)] = x;
                if (jv == 0) {
                    sm.cnt[b] = c % PART_SECTOR;
                    sm.written[b] = (uint16_t)min(wr + nfull, cap);
                }
            }
        }
    });
    __syncthreads();
    // end of the run: the left-over payloads of every slot, padded to one sector with dummy bins
    for (uint32_t b = tid; b < nb; b += COUNT_THREADS) {
        const uint32_t left = sm.cnt[b];
        uint32_t wr = sm.written[b];
        if (left) {
            uint32_t e[PART_SECTOR / 2];
            const uint16_t* src16 = sm.staged + (size_t)b * slot_size;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(src16);
            if (wr < cap) {
#pragma unroll
                for (int i = 0; i < PART_SECTOR / 2; i++) {
                    uint32_t w = src[i];
                    const uint32_t pad = payload_pad(b, b * 16u + 2u * i);
                    if (2u * i >= left) w = (w & 0xFFFF0000u) | pad;
                    if (2u * i + 1u >= left) w = (w & 0x0000FFFFu) | ((pad ^ 1u) << 16);
                    e[i] = w;
                }
                uint4* dst = region + (b * cap + wr) * 2u;
                dst[0] = make_uint4(e[0], e[1], e[2], e[3]);
                dst[1] = make_uint4(e[4], e[5], e[6], e[7]);
                wr++;
            } else {
                for (uint32_t i = 0; i < left; i++)
                    overflow_push(sink.ov, sink.ov_count, (b << (2 * PART_LOW)) | (src16[i] & (uint32_t)(PART_BINS - 1)));
            }
        }
        nsec[(size_t)blockIdx.x * nb + b] = (uint16_t)wr;
    }
    const unsigned long long total = block_sum_u32(sink.n, &sm.sh_total);
    if (tid == 0 && total) atomicAdd(&stats[sl.genome].total_top, total);
}

struct LevelInfo {                     // per level: index into the caller's k_list (-1: pass-through)
    int ki[16];
};

constexpr int BUCKET_THREADS = 512;

constexpr int BUCKET_NS_CHUNK = 1024;      // runs whose sector counts are staged at a time
#ifndef KM_BUCKET_BULK_STORE
#define KM_BUCKET_BULK_STORE 1           // 1: the level-k count slice leaves as one shared -> global bulk copy
#endif
struct BucketSmem {
    uint32_t hist[PART_BINS + 256];    // 64 KB (reused in place by the in-bucket cascade) + the padding's dummy bins
    uint16_t ns[BUCKET_NS_CHUNK];      // sectors in this bucket's region of every run of the genome
    unsigned long long tot[16];
    uint32_t* lvl_counts[16];          // per level: this bucket's slice of the count row
    float* lvl_freq[16];               //            ... of the frequency row (nullptr: not requested)
    double lvl_inv[16];                //            1 / windows of that level
};
constexpr int BUCKET_CTAS_PER_SM = 3;
static_assert(sizeof(BucketSmem) <= 75 * 1024, "three bucket CTAs must fit in one SM's shared memory");

// ---- 1-D TMA bulk copy (no tensor map)
// shared -> global, tracked by the thread's bulk group
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// 8 payload entries of one 128-bit vector: live ones are bins of the bucket, padding goes to the
// dummy bins behind them (see partition_kernel).
__device__ __forceinline__ void hist_add8(uint32_t hbase, const uint4& x, uint32_t fix) {
    const uint32_t w[4] = {x.x ^ fix, x.y ^ fix, x.z ^ fix, x.w ^ fix};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hbase + ((w[i] & 0xFFFFu) << 2)) : "memory");
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hbase + ((w[i] >> 16) << 2)) : "memory");
    }
}

__global__ void __launch_bounds__(BUCKET_THREADS, BUCKET_CTAS_PER_SM)
bucket_kernel(LevelMap lm, RowSpec row, LevelInfo li, int k, int k_stop, const GenomeRuns* __restrict__ grs,
              const uint4* __restrict__ payload, const uint16_t* __restrict__ nsec, uint32_t cap,
              const GenomeStats* __restrict__ stats, float* freq, uint64_t freq_stride, uint64_t* totals,
              uint32_t genome0) {
    extern __shared__ __align__(16) unsigned char bucket_smem_raw[];
    BucketSmem& sm = *reinterpret_cast<BucketSmem*>(bucket_smem_raw);
    const int tid = threadIdx.x;
    const uint32_t b = blockIdx.x;
    const uint32_t nb = gridDim.x;
    const uint32_t g = genome0 + blockIdx.y;
    const GenomeRuns gr = grs[g];
    if (tid < row.nk) {
        const int j = row.k[tid];
        unsigned long long t = stats[g].total_top;
        for (int i = j; i < k; i++) t += stats[g].n_tail[i];
        sm.tot[tid] = t;
        if (totals && b == 0) totals[(uint64_t)g * row.nk + tid] = t;
    }
    __syncthreads();
    if (tid >= 32 && tid < 48) {                                  // per-level pointers, once
        const int level = tid - 32;
        uint32_t* cp = nullptr;
        float* fp = nullptr;
        double inv = 0.0;
        if (level >= k_stop && level <= k) {
            const size_t n_level = (size_t)PART_BINS >> (2 * (k - level));
            cp = lm.ptr(g, level) + (size_t)b * n_level;
            const int ki = li.ki[level];
            if (ki >= 0) {
                if (freq) fp = freq + (uint64_t)g * freq_stride + row.off[ki] + (size_t)b * n_level;
                inv = sm.tot[ki] ? 1.0 / (double)sm.tot[ki] : 0.0;
            }
        }
        sm.lvl_counts[level] = cp;
        sm.lvl_freq[level] = fp;
        sm.lvl_inv[level] = inv;
    }
    constexpr int PER1 = PART_BINS / 4 / BUCKET_THREADS;          // 8 level-(k-1) bins per thread
    constexpr int PER2 = PART_BINS / 16 / BUCKET_THREADS;         // 2 level-(k-2) bins per thread
    // (run-end tails of these levels are added afterwards from the genome's tail list)
    // Stream this bucket's regions: run r of the genome holds ns[r] sectors at
    // payload[((run0 + r) * nb + b) * cap ...]; one warp per region, 128-bit coalesced loads in rounds of
    // 32 vectors, four rounds in flight
    const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(sm.hist);
    const uint32_t fix = payload_fix(b);
    {
        uint4* h4 = reinterpret_cast<uint4*>(sm.hist);
        for (int i = tid; i < (PART_BINS + 256) / 4; i += BUCKET_THREADS) h4[i] = make_uint4(0, 0, 0, 0);
    }
    constexpr int NWARP = BUCKET_THREADS / 32;
    const int lane = tid & 31, wid = tid >> 5;
    const size_t run_stride = (size_t)nb * cap * 2;                       // uint4 per run
    for (uint32_t rb = 0; rb < gr.n_runs; rb += BUCKET_NS_CHUNK) {
        const uint32_t nr = min(gr.n_runs - rb, (uint32_t)BUCKET_NS_CHUNK);
        __syncthreads();                                                  // histogram cleared / previous chunk consumed
        for (uint32_t i = tid; i < nr; i += BUCKET_THREADS)
            sm.ns[i] = __ldg(nsec + (size_t)(gr.run0 + rb + i) * nb + b);
        __syncthreads();
        const uint4* base = payload + ((size_t)(gr.run0 + rb) * nb + b) * cap * 2;
        // the warp's rounds: (run r, c) = vectors [32 c, 32 c + 32) of run r's region, r = wid, wid + 16, ...
        uint32_t r = wid, c = 0, nv = 0;
        auto settle = [&]() {                                             // skip runs that are used up
            while (r < nr && 32u * c >= (nv = 2u * sm.ns[r])) { r += NWARP; c = 0; }
        };
        auto issue = [&](uint4& x, bool& live) -> bool {                  // next round: load; false when the warp is done
            settle();
            if (r >= nr) { live = false; return false; }
            const uint32_t v = 32u * c + lane;
            live = v < nv;
            if (live) x = __ldg(base + (size_t)r * run_stride + v);
            c++;
            return true;
        };
        uint4 x0, x1, x2, x3;
        bool l0, l1, l2, l3;
        bool h0 = issue(x0, l0), h1 = issue(x1, l1), h2 = issue(x2, l2), h3 = issue(x3, l3);
        while (h0) {
            if (l0) hist_add8(hbase, x0, fix);
            h0 = issue(x0, l0);
            if (!h1) break;
            if (l1) hist_add8(hbase, x1, fix);
            h1 = issue(x1, l1);
            if (!h2) break;
            if (l2) hist_add8(hbase, x2, fix);
            h2 = issue(x2, l2);
            if (!h3) break;
            if (l3) hist_add8(hbase, x3, fix);
            h3 = issue(x3, l3);
        }
    }
    // level k: this bucket's 16384 bins leave as ONE bulk copy shared -> global (the atomics above went through
    // the generic proxy: fence, barrier, then one thread issues the copy) ...
#if KM_BUCKET_BULK_STORE
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) bulk_store(sm.lvl_counts[k], hbase, PART_BINS * 4);
#else
    __syncthreads();
#endif
    // ... while the threads turn them into frequencies and level k-1 (kept in registers)
    uint32_t v1r[PER1];
    {
        float* fk = sm.lvl_freq[k];
        const double inv = sm.lvl_inv[k];
        const bool down = k - 1 >= k_stop;
        uint32_t* c1 = down ? sm.lvl_counts[k - 1] : nullptr;
        float* f1 = down ? sm.lvl_freq[k - 1] : nullptr;
        const double inv1 = down ? sm.lvl_inv[k - 1] : 0.0;
#pragma unroll
        for (int u = 0; u < PER1; u++) {
            const int i = tid + u * BUCKET_THREADS;
            const uint4 c = reinterpret_cast<const uint4*>(sm.hist)[i];
#if !KM_BUCKET_BULK_STORE
            reinterpret_cast<uint4*>(sm.lvl_counts[k])[i] = c;
#endif
            if (fk) {
                float4 f;
                f.x = (float)((double)c.x * inv); f.y = (float)((double)c.y * inv);
                f.z = (float)((double)c.z * inv); f.w = (float)((double)c.w * inv);
                reinterpret_cast<float4*>(fk)[i] = f;
            }
            v1r[u] = c.x + c.y + c.z + c.w;
            if (down) {
                c1[i] = v1r[u];
                if (f1) f1[i] = (float)((double)v1r[u] * inv1);
            }
        }
    }
#if KM_BUCKET_BULK_STORE
    if (tid == 0) bulk_store_wait_read();                             // the bulk copy has read hist[]
#endif
    if (k - 2 < k_stop) return;
    __syncthreads();                                                  // everybody is done reading hist[]
#pragma unroll
    for (int u = 0; u < PER1; u++) sm.hist[tid + u * BUCKET_THREADS] = v1r[u];   // level k-1, in place
    __syncthreads();
    {   // level k-2: 1024 bins, whole CTA; results to hist[4096 ..)
        uint32_t* cl = sm.lvl_counts[k - 2];
        float* fl = sm.lvl_freq[k - 2];
        const double invl = sm.lvl_inv[k - 2];
#pragma unroll
        for (int u = 0; u < PER2; u++) {
            const int i = tid + u * BUCKET_THREADS;
            const uint4 c = reinterpret_cast<const uint4*>(sm.hist)[i];
            const uint32_t v = c.x + c.y + c.z + c.w;
            cl[i] = v;
            if (fl) fl[i] = (float)((double)v * invl);
            sm.hist[4096 + i] = v;
        }
    }
    __syncthreads();
    if (tid >= 32) return;
    // the small rest (256, 64, 16, 4, 1 bins) by warp 0, ping-pong inside hist[]
    uint32_t* cur = sm.hist + 4096;
    uint32_t* nxt = sm.hist + 8192;
    int n_cur = 1024;
    for (int level = k - 3; level >= k_stop; level--) {
        const int n_next = n_cur >> 2;
        uint32_t* cl = sm.lvl_counts[level];
        float* fl = sm.lvl_freq[level];
        const double invl = sm.lvl_inv[level];
        for (int i = tid; i < n_next; i += 32) {
            const uint4 c = reinterpret_cast<const uint4*>(cur)[i];
            const uint32_t v = c.x + c.y + c.z + c.w;
            cl[i] = v;
            if (fl) fl[i] = (float)((double)v * invl);
            nxt[i] = v;
        }
        __syncwarp();
        uint32_t* t = cur; cur = nxt; nxt = t;
        n_cur = n_next;
    }
}

// The windows that found their slot full: add them to level k and to every level of the bucket
// subtrees (the cascade below k_stop runs after this kernel), then refresh the touched
// frequencies from the final counts (idempotent, so concurrent duplicates are harmless).
// Slots run full where the sequence is repetitive, and then the list holds the same few k-mers over and
// over: each CTA first folds its share of the list into a small shared-memory hash table (k-mer -> how
// often), so that a tandem repeat costs a handful of global atomics per CTA instead of one per window.
constexpr int OV_TABLE = 4096;
constexpr uint32_t OV_EMPTY = 0xFFFFFFFFu;                 // not a valid k-mer index (k <= 14 here)

__global__ void __launch_bounds__(256)
overflow_kernel(LevelMap lm, RowSpec row, LevelInfo li, int k, int k_stop, const GenomeDev* __restrict__ gds,
                const uint32_t* __restrict__ overflow, const unsigned int* __restrict__ ov_counts, uint64_t batch_lo,
                const GenomeStats* __restrict__ stats, float* freq, uint64_t freq_stride, uint32_t genome0, int pass) {
    __shared__ uint32_t t_key[OV_TABLE];
    __shared__ uint32_t t_cnt[OV_TABLE];
    __shared__ unsigned int t_used;
    const uint32_t g = genome0 + blockIdx.y;
    const unsigned int n = ov_counts[g];
    if (blockIdx.x * 256u >= n) return;
    const uint32_t* ov = overflow + (gds[g].file_lo - batch_lo);
    const int tid = threadIdx.x;
    auto apply = [&](uint32_t idx, uint32_t cnt) {
        for (int level = k; level >= k_stop; level--) {
            const uint32_t x = idx >> (2 * (k - level));
            uint32_t* c = lm.ptr(g, level);
            if (pass == 0) {
                atomicAdd(c + x, cnt);
            } else if (freq && li.ki[level] >= 0) {
                const int ki = li.ki[level];
                unsigned long long t = stats[g].total_top;
                for (int q = level; q < k; q++) t += stats[g].n_tail[q];
                freq[(uint64_t)g * freq_stride + row.off[ki] + x] = t ? (float)((double)c[x] * (1.0 / (double)t)) : 0.0f;
            }
        }
    };
    auto flush = [&]() {                                   // (whole CTA, between barriers)
        for (int e = tid; e < OV_TABLE; e += 256) {
            const uint32_t key = t_key[e];
            if (key != OV_EMPTY) apply(key, t_cnt[e]);
            t_key[e] = OV_EMPTY;
            t_cnt[e] = 0;
        }
        if (tid == 0) t_used = 0;
        __syncthreads();
    };
    if (n < 4096u) {                                       // the usual handful of entries: no table
        for (unsigned int i = blockIdx.x * 256u + tid; i < n; i += gridDim.x * 256u) apply(ov[i], 1u);
        return;
    }
    for (int e = tid; e < OV_TABLE; e += 256) { t_key[e] = OV_EMPTY; t_cnt[e] = 0; }
    if (tid == 0) t_used = 0;
    __syncthreads();
    for (unsigned int base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        const unsigned int i = base + tid;
        if (i < n) {
            const uint32_t idx = ov[i];
            uint32_t h = (idx * 2654435761u) >> 20;        // 12 bits
            bool placed = false;
            for (int probe = 0; probe < 8 && !placed; probe++) {
                const uint32_t old = atomicCAS(&t_key[h], OV_EMPTY, idx);
                if (old == OV_EMPTY) atomicAdd(&t_used, 1u);
                if (old == OV_EMPTY || old == idx) {
                    atomicAdd(&t_cnt[h], 1u);
                    placed = true;
                }
                h = (h + 1) & (OV_TABLE - 1);
            }
            if (!placed) apply(idx, 1u);                   // crowded neighbourhood: straight to global memory
        }
        __syncthreads();
        const bool crowded = t_used > OV_TABLE / 2;
        __syncthreads();                                   // everybody has read it before flush() resets it
        if (crowded) flush();
    }
    flush();
}

// One run-end tail of level j (>= k_stop) counts at level j and, through the marginal sums, at every
// level below it down to k_stop (the cascade below k_stop runs later).  pass 0 adds, pass 1 refreshes
// the touched frequencies from the final counts.
struct TailApply {
    LevelMap lm;
    RowSpec row;
    LevelInfo li;
    int k, k_stop, pass;
    const GenomeStats* stats;
    float* freq;
    uint64_t freq_stride;
    __device__ __forceinline__ void apply(uint32_t g, int j, uint32_t idx) const {
        for (int level = j; level >= k_stop; level--) {
            const uint32_t x = idx >> (2 * (j - level));
            uint32_t* c = lm.ptr(g, level);
            if (pass == 0) {
                atomicAdd(c + x, 1u);
            } else if (freq && li.ki[level] >= 0) {
                const int ki = li.ki[level];
                unsigned long long t = stats[g].total_top;
                for (int q = level; q < k; q++) t += stats[g].n_tail[q];
                freq[(uint64_t)g * freq_stride + row.off[ki] + x] = t ? (float)((double)c[x] * (1.0 / (double)t)) : 0.0f;
            }
        }
    }
};

__global__ void __launch_bounds__(256)
tails_apply_kernel(TailApply ap, const GenomeDev* __restrict__ gds, TailList tl, uint32_t genome0) {
    const uint32_t g = genome0 + blockIdx.y;
    const unsigned int n = tl.counts[g];
    if (!n) return;
    uint32_t cap;
    const unsigned long long* list = tail_list_of(tl, gds[g], g, &cap);
    if (n > cap) return;                                    // dropped: tails_rescan_kernel does this genome
    for (unsigned int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const unsigned long long e = list[i];
        ap.apply(g, (int)(e >> 32), (uint32_t)e);
    }
}

struct NullSink {
    __device__ __forceinline__ void count(uint32_t, uint64_t) {}
    __device__ __forceinline__ void count4(uint32_t, uint32_t, uint32_t, uint32_t, uint64_t, uint64_t, uint64_t, uint64_t) {}
    __device__ __forceinline__ void count8(const uint32_t*, const uint64_t*) {}
    __device__ __forceinline__ void count8_tail(const uint32_t*, const uint64_t*, bool) {}
};
struct RescanTails {
    const TailApply* ap;
    uint32_t genome;
    __device__ __noinline__ void tail(int j, uint32_t idx) const {
        if (j >= ap->k_stop) ap->apply(genome, j, idx);
    }
};

// Genomes whose tail list ran full (a run end every few bytes): walk the run ends again.  A small
// persistent grid, so that the usual case (no list ran full) costs one wave of CTAs that exit at once.
__global__ void __launch_bounds__(COUNT_THREADS)
tails_rescan_kernel(const uint8_t* __restrict__ buf, const GenomeDev* __restrict__ gds, const Slice* __restrict__ slices,
                    int n_slices, DenseParams P, TailApply ap, TailList tl) {
    if (*tl.any_full == 0) return;
    KM_TILE_SMEM(rs)
    for (int i = blockIdx.x; i < n_slices; i += gridDim.x) {
        const Slice sl = slices[i];
        const GenomeDev gd = gds[sl.genome];
        uint32_t cap;
        (void)tail_list_of(tl, gd, sl.genome, &cap);
        if (tl.counts[sl.genome] <= cap) continue;
        Genome g;
        g.b = buf;
        g.lo = gd.lo;
        g.hi = gd.hi;
        NullSink sink;
        RescanTails tails;
        tails.ap = &ap; tails.genome = sl.genome;
        walk_slice<true>(buf, g, sl, P, sink, tails, tc, [](uint32_t) {});
        __syncthreads();
    }
}

// Frequencies of the requested levels below `level_limit` (the few levels the bucket
// kernel's subtrees do not reach); one block per genome.
__global__ void __launch_bounds__(256)
finalize_low_kernel(LevelMap lm, RowSpec row, int k_top, int level_limit, const GenomeStats* __restrict__ stats,
                    float* freq, uint64_t freq_stride, uint32_t genome0) {
    __shared__ unsigned long long tot[16];
    const uint32_t g = genome0 + blockIdx.x;
    level_totals(row, k_top, stats, g, tot, nullptr, false);
    if (!freq) return;
    for (int ki = 0; ki < row.nk; ki++) {
        const int j = row.k[ki];
        if (j >= level_limit) continue;
        const uint32_t* c = lm.ptr(g, j);
        float* f = freq + (uint64_t)g * freq_stride + row.off[ki];
        const double inv = tot[ki] ? 1.0 / (double)tot[ki] : 0.0;
        for (uint32_t x = threadIdx.x; x < (1u << (2 * j)); x += 256) f[x] = (float)((double)c[x] * inv);
    }
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

// Canonical mode: fold every requested level onto min(kmer, revcomp) in place, then normalise.
// Levels below CANON_TILED_MIN_K: the thread of the smaller index of each {x, rc(x)} pair owns both bins.
constexpr int CANON_TILED_MIN_K = 7;
__global__ void __launch_bounds__(256)
finalize_canonical_kernel(LevelMap lm, RowSpec row, int k_top, const GenomeStats* __restrict__ stats,
                          float* freq, uint64_t freq_stride, uint64_t* totals, uint32_t genome0) {
    __shared__ unsigned long long tot[16];
    const int tid = threadIdx.x;
    const uint32_t g = genome0 + blockIdx.y;
    level_totals(row, k_top, stats, g, tot, totals, blockIdx.x == 0);
    for (int ki = 0; ki < row.nk; ki++) {
        const int j = row.k[ki];
        if (j >= CANON_TILED_MIN_K) continue;                       // finalize_canonical_tiled_kernel
        uint32_t* c = lm.ptr(g, j);
        float* f = freq ? freq + (uint64_t)g * freq_stride + row.off[ki] : nullptr;
        const double inv = tot[ki] ? 1.0 / (double)tot[ki] : 0.0;
        for (uint32_t x = blockIdx.x * 256 + tid; x < (1u << (2 * j)); x += gridDim.x * 256) {
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
}

// Levels j >= 7: rc(x) reverses the base order, so the partner of a run of consecutive bins is a column of stride
// 4^(j-1) -- one 32-byte sector per bin for the pairwise kernel above (0.17 ms for the 64 MB row of k = 12, four times
// what its bytes cost).  Tiles make both sides contiguous: with x = A | M | B (A, B three bases, M the middle j - 6),
// rc(x) = rc(B) | rc(M) | rc(A), i.e. the 64 x 64 tile M maps onto the tile rc(M).  A CTA loads the tile pair
// {M, rc(M)} (M <= rc(M)) as 64 rows of 256 contiguous bytes into shared memory, folds, and writes both back.
__global__ void __launch_bounds__(256)
finalize_canonical_tiled_kernel(LevelMap lm, RowSpec row, int ki, int k_top, const GenomeStats* __restrict__ stats,
                                float* freq, uint64_t freq_stride, uint32_t genome0) {
    __shared__ uint32_t tile[2][64][65];
    __shared__ double s_inv;
    const int j = row.k[ki];
    const int mid = j - 6;
    const uint32_t M = blockIdx.x;
    const uint32_t Mr = revcomp_code(M, mid);
    if (M > Mr) return;
    const int tid = threadIdx.x;
    const uint32_t g = genome0 + blockIdx.y;
    const int n_tiles = M == Mr ? 1 : 2;
    if (tid == 0) {
        unsigned long long t = stats[g].total_top;
        for (int i = j; i < k_top; i++) t += stats[g].n_tail[i];
        s_inv = t ? 1.0 / (double)t : 0.0;
    }
    uint32_t* c = lm.ptr(g, j);
    float* f = freq ? freq + (uint64_t)g * freq_stride + row.off[ki] : nullptr;
    const uint32_t mids[2] = {M, Mr};
    // a thread moves 4 consecutive bins (128 bits) of 4 rows per tile
    const int q = tid & 15, r0 = tid >> 4;
#pragma unroll
    for (int t = 0; t < 2; t++) {
        if (t >= n_tiles) break;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int A = r0 + 16 * i;
            const uint32_t x = ((uint32_t)A << (2 * (j - 3))) | (mids[t] << 6) | (uint32_t)(4 * q);
            const uint4 v = *reinterpret_cast<const uint4*>(c + x);
            tile[t][A][4 * q] = v.x; tile[t][A][4 * q + 1] = v.y; tile[t][A][4 * q + 2] = v.z; tile[t][A][4 * q + 3] = v.w;
        }
    }
    __syncthreads();
    const double inv = s_inv;
#pragma unroll
    for (int t = 0; t < 2; t++) {
        if (t >= n_tiles) break;
        const int p = n_tiles == 1 ? 0 : (t ^ 1);                   // the tile that holds the partners
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int A = r0 + 16 * i;
            const uint32_t Ar = revcomp_code((uint32_t)A, 3);
            const uint32_t x0 = ((uint32_t)A << (2 * (j - 3))) | (mids[t] << 6) | (uint32_t)(4 * q);
            uint32_t out[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t B = (uint32_t)(4 * q + b);
                const uint32_t Br = revcomp_code(B, 3);
                const uint32_t x = x0 + b;
                const uint32_t rc = (Br << (2 * (j - 3))) | (mids[p] << 6) | Ar;
                const uint32_t own = tile[t][A][B];
                out[b] = x < rc ? own + tile[p][Br][Ar] : (x == rc ? own : 0u);
            }
            *reinterpret_cast<uint4*>(c + x0) = make_uint4(out[0], out[1], out[2], out[3]);
            if (f)
                *reinterpret_cast<float4*>(f + x0) = make_float4((float)((double)out[0] * inv), (float)((double)out[1] * inv),
                                                                 (float)((double)out[2] * inv), (float)((double)out[3] * inv));
        }
    }
}

// ---------------------------------------------------------------- launchers
int dense_setup_attributes() {
    KM_CUDA(cudaFuncSetAttribute(count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (1 << (2 * SMEM_MAX_K)) * 4));
    KM_CUDA(cudaFuncSetAttribute(count8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
    KM_CUDA(cudaFuncSetAttribute(partition_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem<10>)));
    KM_CUDA(cudaFuncSetAttribute(partition_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem<8>)));
    KM_CUDA(cudaFuncSetAttribute(partition_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem<6>)));
    KM_CUDA(cudaFuncSetAttribute(partition_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PartSmem<4>)));
    KM_CUDA(cudaFuncSetAttribute(bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BucketSmem)));
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

int launch_slice_headers(const uint8_t* d_fasta, const GenomeDev* d_genomes, Slice* d_slices, int n_slices,
                         cudaStream_t s) {
    if (n_slices <= 0) return KMERML_OK;
    slice_header_kernel<<<(n_slices + 127) / 128, 128, 0, s>>>(d_fasta, d_genomes, d_slices, n_slices);
    slice_long_scan_kernel<<<n_slices, 256, 0, s>>>(d_fasta, d_genomes, d_slices, n_slices);
    slice_long_resolve_kernel<<<1, 1024, 0, s>>>(d_fasta, d_genomes, d_slices, n_slices);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

int launch_count(const uint8_t* d_fasta, const GenomeDev* d_genomes, const Slice* d_slices, int n_slices,
                 int k, int k_bottom, int min_rec, bool use_smem, const LevelMap& lm, GenomeStats* d_stats,
                 cudaStream_t s) {
    if (n_slices <= 0) return KMERML_OK;
    DenseParams P = make_params(k, min_rec, k > k_bottom, k_bottom);
    if (use_smem && k == 8) {
        count8_kernel<<<n_slices, COUNT8_THREADS, 32768 * 4, s>>>(d_fasta, d_genomes, d_slices, P, lm, d_stats);
    } else if (use_smem) {
        size_t smem = (size_t)(1u << (2 * k)) * 4;
        count_kernel<1><<<n_slices, COUNT_THREADS, smem, s>>>(d_fasta, d_genomes, d_slices, P, lm, d_stats, nullptr);
    } else {
        count_kernel<0><<<n_slices, COUNT_THREADS, 0, s>>>(d_fasta, d_genomes, d_slices, P, lm, d_stats, nullptr);
    }
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

// Sectors one (run, bucket) region can hold: twice the mean of a run of `tiles_per_run` tiles.
uint32_t part_region_cap(int k, int tiles_per_run) {
    const uint32_t nb = 1u << (2 * (k - PART_LOW));
    return (uint32_t)tiles_per_run * (2u * TILE_BYTES / PART_SECTOR) / nb;
}

int launch_partition(const uint8_t* d_fasta, const GenomeDev* d_genomes, const Slice* d_runs, int n_runs,
                     int tiles_per_run, int k, int k_bottom, int min_rec, const LevelMap& lm,
                     GenomeStats* d_stats, void* d_payload, uint16_t* d_nsec, uint32_t* d_overflow,
                     unsigned int* d_ov_counts, uint64_t batch_lo, unsigned long long* d_tail_list,
                     unsigned int* d_tail_counts, unsigned int* d_tail_any, uint32_t genome0, cudaStream_t s) {
    if (n_runs <= 0) return KMERML_OK;
    DenseParams P = make_params(k, min_rec, k > k_bottom, k_bottom);
    const int k_stop = std::max(k - PART_LOW, k_bottom);
    const uint32_t cap = part_region_cap(k, tiles_per_run);
    TailList tl;
    tl.list = d_tail_list; tl.counts = d_tail_counts; tl.batch_lo = batch_lo; tl.genome0 = genome0;
    tl.any_full = d_tail_any;
#define KM_LAUNCH_PART(NBS)                                                                           \
    partition_kernel<NBS><<<n_runs, COUNT_THREADS, sizeof(PartSmem<NBS>), s>>>(                       \
        d_fasta, d_genomes, d_runs, P, lm, d_stats, (uint4*)d_payload, d_nsec, cap, d_overflow,       \
        d_ov_counts, batch_lo, tl, k_stop)
    switch (2 * (k - PART_LOW)) {                // log2(buckets)
        case 10: KM_LAUNCH_PART(10); break;      // k = 12
        case 8: KM_LAUNCH_PART(8); break;        // k = 11
        case 6: KM_LAUNCH_PART(6); break;        // k = 10
        default: KM_LAUNCH_PART(4); break;       // k = 9
    }
#undef KM_LAUNCH_PART
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

static LevelInfo make_level_info(const RowSpec& row) {
    LevelInfo li;
    for (int j = 0; j < 16; j++) {
        li.ki[j] = -1;
        for (int i = 0; i < row.nk; i++)
            if (row.k[i] == j) li.ki[j] = i;
    }
    return li;
}

int launch_bucket(const LevelMap& lm, const RowSpec& row, int k, int k_bottom, const void* d_genome_runs,
                  int tiles_per_run, const void* d_payload, const uint16_t* d_nsec, const GenomeStats* d_stats,
                  float* d_freq, uint64_t freq_stride, uint64_t* d_totals, uint32_t genome0, int n_genomes,
                  cudaStream_t s) {
    if (n_genomes <= 0) return KMERML_OK;
    const int nb = 1 << (2 * (k - PART_LOW));
    const LevelInfo li = make_level_info(row);
    const int k_stop = std::max(k - PART_LOW, k_bottom);
    const uint32_t cap = part_region_cap(k, tiles_per_run);
    for (int h0 = 0; h0 < n_genomes; h0 += 32768) {           // gridDim.y <= 65535
        dim3 grid((unsigned)nb, (unsigned)std::min(32768, n_genomes - h0));
        bucket_kernel<<<grid, BUCKET_THREADS, sizeof(BucketSmem), s>>>(lm, row, li, k, k_stop,
            (const GenomeRuns*)d_genome_runs, (const uint4*)d_payload, d_nsec, cap, d_stats, d_freq, freq_stride,
            d_totals, genome0 + (uint32_t)h0);
        KM_CUDA(cudaGetLastError());
    }
    return KMERML_OK;
}

int launch_overflow(const LevelMap& lm, const RowSpec& row, int k, int k_bottom, const GenomeDev* d_genomes,
                    const uint32_t* d_overflow, const unsigned int* d_ov_counts, uint64_t batch_lo,
                    const GenomeStats* d_stats, float* d_freq, uint64_t freq_stride, uint32_t genome0, int n_genomes,
                    cudaStream_t s) {
    if (n_genomes <= 0) return KMERML_OK;
    const LevelInfo li = make_level_info(row);
    const int k_stop = std::max(k - PART_LOW, k_bottom);
    dim3 grid(64, (unsigned)n_genomes);
    for (int pass = 0; pass < (d_freq ? 2 : 1); pass++) {
        overflow_kernel<<<grid, 256, 0, s>>>(lm, row, li, k, k_stop, d_genomes, d_overflow, d_ov_counts, batch_lo,
                                             d_stats, d_freq, freq_stride, genome0, pass);
        KM_CUDA(cudaGetLastError());
    }
    return KMERML_OK;
}

// Adds the run-end tails the partition kernel listed (levels >= k_stop) to the rows the bucket kernel
// stored; `pass` as in launch_overflow.  Two launches per pass: the lists, and the rescan of genomes
// whose list ran full (its CTAs exit at once otherwise).
int launch_tails(const uint8_t* d_fasta, const LevelMap& lm, const RowSpec& row, int k, int k_bottom, int min_rec,
                 const GenomeDev* d_genomes, const Slice* d_tiles, int n_tiles, unsigned long long* d_tail_list,
                 unsigned int* d_tail_counts, unsigned int* d_tail_any, uint64_t batch_lo, const GenomeStats* d_stats,
                 float* d_freq, uint64_t freq_stride, uint32_t genome0, int n_genomes, cudaStream_t s) {
    if (n_genomes <= 0 || k <= k_bottom) return KMERML_OK;
    TailApply ap;
    ap.lm = lm; ap.row = row; ap.li = make_level_info(row);
    ap.k = k; ap.k_stop = std::max(k - PART_LOW, k_bottom);
    ap.stats = d_stats; ap.freq = d_freq; ap.freq_stride = freq_stride;
    TailList tl;
    tl.list = d_tail_list; tl.counts = d_tail_counts; tl.batch_lo = batch_lo; tl.genome0 = genome0;
    tl.any_full = d_tail_any;
    DenseParams P = make_params(k, min_rec, true, ap.k_stop);
    for (int pass = 0; pass < (d_freq ? 2 : 1); pass++) {
        ap.pass = pass;
        for (int h0 = 0; h0 < n_genomes; h0 += 32768) {
            dim3 grid(8, (unsigned)std::min(32768, n_genomes - h0));
            tails_apply_kernel<<<grid, 256, 0, s>>>(ap, d_genomes, tl, genome0 + (uint32_t)h0);
        }
        if (n_tiles > 0)
            tails_rescan_kernel<<<std::min(n_tiles, 4 * 148), COUNT_THREADS, 0, s>>>(d_fasta, d_genomes, d_tiles, n_tiles, P, ap, tl);
        KM_CUDA(cudaGetLastError());
    }
    return KMERML_OK;
}

int launch_finalize_low(const LevelMap& lm, const RowSpec& row, int k_top, int level_limit,
                        const GenomeStats* d_stats, float* d_freq, uint64_t freq_stride, uint32_t genome0,
                        int n_genomes, cudaStream_t s) {
    if (n_genomes <= 0 || !d_freq) return KMERML_OK;
    finalize_low_kernel<<<n_genomes, 256, 0, s>>>(lm, row, k_top, level_limit, d_stats, d_freq, freq_stride, genome0);
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

int launch_encode(const uint8_t* d_fasta, const GenomeDev* d_genomes, const Slice* d_slices, int n_slices,
                  uint8_t* d_symbols, cudaStream_t s) {
    if (n_slices <= 0) return KMERML_OK;
    DenseParams P = make_params(1, 1, false, 1);
    LevelMap lm = {};
    count_kernel<3><<<n_slices, COUNT_THREADS, 0, s>>>(d_fasta, d_genomes, d_slices, P, lm, nullptr,
                                                        reinterpret_cast<uint32_t*>(d_symbols));
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

int finalize_launches(const RowSpec& row, bool canonical) {
    int n = 1;
    if (canonical)
        for (int i = 0; i < row.nk; i++) n += row.k[i] >= CANON_TILED_MIN_K ? 1 : 0;
    return n;
}

int launch_finalize(const LevelMap& lm, const RowSpec& row, int k_top, bool canonical,
                    const GenomeStats* d_stats, float* d_freq, uint64_t freq_stride, uint64_t* d_totals,
                    uint32_t genome0, int n_genomes, cudaStream_t s) {
    if (n_genomes <= 0) return KMERML_OK;
    unsigned long long n = row.off[row.nk];
    if (canonical) {
        unsigned long long n_small = 0;
        for (int i = 0; i < row.nk; i++)
            if (row.k[i] < CANON_TILED_MIN_K) n_small = std::max(n_small, 1ull << (2 * row.k[i]));
        unsigned gx = (unsigned)((n_small + 256ull * 4 - 1) / (256ull * 4));
        if (gx < 1) gx = 1;                                          // (always launched: it writes the totals)
        dim3 grid(gx, (unsigned)n_genomes);
        finalize_canonical_kernel<<<grid, 256, 0, s>>>(lm, row, k_top, d_stats, d_freq, freq_stride, d_totals, genome0);
        for (int i = 0; i < row.nk; i++) {
            if (row.k[i] < CANON_TILED_MIN_K) continue;
            dim3 tg(1u << (2 * (row.k[i] - 6)), (unsigned)n_genomes);
            finalize_canonical_tiled_kernel<<<tg, 256, 0, s>>>(lm, row, i, k_top, d_stats, d_freq, freq_stride, genome0);
        }
    } else {
        unsigned gx = d_freq ? (unsigned)(((n >> 2) + 255) / 256) : 1u;
        dim3 grid(gx, (unsigned)n_genomes);
        finalize_kernel<<<grid, 256, 0, s>>>(lm, row, k_top, d_stats, d_freq, freq_stride, d_totals, genome0);
    }
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

}  // namespace km
