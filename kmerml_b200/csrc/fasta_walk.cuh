// fasta_walk.cuh -- exact FASTA byte-stream semantics for carry-free k-mer walking.
//
// Shared between the sm_100a kernels (dense.cu, sparse.cu, features.cu) and the CPU thread
// emulator in tests/emu (same source compiled with g++), so the per-thread logic
// that runs on the GPU is the logic the CPU tests exercise.
//
// Semantics restated (paths relative to the reference tree):
//   * record / line structure of Bio.SeqIO.parse(..., "fasta") as used at
//     kmerml/kmers/generate.py:39 -- text mode (\n, \r\n, \r end a line), a line
//     whose first byte is '>' starts a record, everything before the first such
//     line is ignored, sequence lines are rstrip()ed and ' ' removed;
//   * upper-casing (generate.py:41) is folded into base_code (c & 0xDF);
//   * a window is counted iff all k symbols are in ACGT (generate.py:55-56), never
//     across records (:39), and only in records with len >= max(k_values) (:44-46).
//
// Design: every thread owns the windows whose LAST base lies in its 32-byte chunk.
//   * clean chunk (only bases and at most one '\n') right after a clean chunk: the
//     chunk is bit-compacted into one 64-bit register, the k-1 bases before it come
//     from the neighbour thread's register, and every window is one funnel shift --
//     no per-byte state machine at all (pack_clean / emit_clean);
//   * anything else (header lines, N runs, IUPAC codes, '\r', blanks, the first
//     chunk of a slice ...): generic byte walker whose start state is recovered by
//     walking BACKWARDS over global memory (lookback), so no state is ever carried
//     between threads, warps, CTAs or GPUs.
#pragma once
#include <stdint.h>

#include <type_traits>

#if defined(__CUDACC__)
#define KM_HD __host__ __device__ __forceinline__
#define KM_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define KM_HD inline
#define KM_HD_NOINLINE inline
#endif

namespace km {

constexpr int CHUNK = 32;            // bytes owned by one thread per tile
constexpr int MAX_DENSE_K = 15;

// One genome = one FASTA file's bytes [lo, hi) inside a batch buffer.  `lo` has
// already been advanced to the first header line (or to hi when there is none).
struct Genome {
    const uint8_t* b;
    uint64_t lo, hi;
};

enum : int { SYM_INV = 4, SYM_SKIP = 5, SYM_HDR = 6 };

KM_HD bool is_term(uint32_t c) { return c == 10u || c == 13u; }
KM_HD bool is_softws(uint32_t c) {
    // what str.rstrip() strips besides ' ', '\n', '\r' (ASCII part of str.isspace())
    return c == 9u || c == 11u || c == 12u || (c >= 0x1cu && c <= 0x1fu);
}

// 0..3 (A C G T, lexicographic) for a base of either case, -1 otherwise.
KM_HD int base_code(uint32_t c) {
    uint32_t u = c & 0xDFu;
    uint32_t code = ((u >> 1) ^ (u >> 2)) & 3u;
    uint32_t expect = (0x54474341u >> (8u * code)) & 0xFFu;   // "ACGT"[code]
    return (u == expect) ? (int)code : -1;
}

// Kind of a non-base byte that lies in a sequence line.
KM_HD_NOINLINE int classify_nonbase(const Genome& g, uint64_t pos, uint32_t c) {
    if (c == 10u || c == 13u || c == 32u) return SYM_SKIP;
    if (c == (uint32_t)'>') return (pos == g.lo || is_term(g.b[pos - 1])) ? SYM_HDR : SYM_INV;
    if (is_softws(c)) {
        // stripped only when nothing but whitespace follows on the line (rstrip)
        uint64_t q = pos + 1;
        while (q < g.hi && (is_softws(g.b[q]) || g.b[q] == 32u)) q++;
        return (q == g.hi || is_term(g.b[q])) ? SYM_SKIP : SYM_INV;
    }
    return SYM_INV;
}

// Next symbol at or after r (r lies in a sequence line).  Returns 0..3, SYM_INV
// (r advanced past it) or SYM_HDR (r stays on the '>' / at g.hi).
KM_HD_NOINLINE int next_symbol(const Genome& g, uint64_t& r) {
    while (r < g.hi) {
        uint32_t c = g.b[r];
        int code = base_code(c);
        if (code >= 0) { r++; return code; }
        int kind = classify_nonbase(g, r, c);
        if (kind == SYM_HDR) return SYM_HDR;
        r++;
        if (kind == SYM_INV) return SYM_INV;
    }
    return SYM_HDR;
}

// Previous symbol strictly before q (q lies in a sequence line or on a line
// start).  Returns 0..3 or SYM_INV with q moved onto that symbol, or SYM_HDR when
// the record start (a header line, or g.lo) is reached.
KM_HD_NOINLINE int prev_symbol(const Genome& g, uint64_t& q) {
    while (q > g.lo) {
        uint64_t p = q - 1;
        uint32_t c = g.b[p];
        if (is_term(c)) {
            // the line that ends at p: is it a header line?
            uint64_t ls = p;
            while (ls > g.lo && !is_term(g.b[ls - 1])) ls--;
            if (ls < p && g.b[ls] == (uint8_t)'>') { q = ls; return SYM_HDR; }
            q = p;
            continue;
        }
        q = p;
        int code = base_code(c);
        if (code >= 0) return code;
        int kind = classify_nonbase(g, p, c);
        if (kind == SYM_SKIP) continue;
        if (kind == SYM_HDR) return SYM_HDR;       // only if the caller started inside a header line
        return SYM_INV;
    }
    return SYM_HDR;
}

// Does the record that contains byte position p (in a sequence line) hold at
// least `need` symbols?  (generate.py:44 -- len(sequence) < max(k_values))
KM_HD_NOINLINE bool record_len_at_least(const Genome& g, uint64_t p, int need) {
    int total = 0;
    uint64_t q = p;
    while (total < need) {
        if (prev_symbol(g, q) == SYM_HDR) break;
        total++;
    }
    uint64_t r = p;
    while (total < need) {
        if (next_symbol(g, r) == SYM_HDR) break;
        total++;
    }
    return total >= need;
}

// ---- slice starts: is the first byte of a slice inside a header line? ----------------------------
// Looking back for the start of the line costs the line's length, and an unwrapped FASTA file holds a
// whole chromosome on one line.  So the look-back is bounded; slices it leaves open are settled by a
// cooperative scan of the preceding slice plus a prefix maximum over the slice table (dense.cu).
constexpr uint64_t LS_UNRESOLVED = ~0ull;
constexpr uint32_t LINE_SCAN_LIMIT = 1024;

struct SliceHead {
    uint64_t hdr_until;      // p lies in a header line that ends here (0: it does not / not known yet)
    uint64_t line_start;     // start of p's line, LS_UNRESOLVED when the bounded look-back gave up
    uint32_t prev_ok;        // the 32-byte chunk before p is a clean sequence chunk ...
    uint32_t prev16;         // ... and these are its last 16 bases
};

// Start of the line that holds p (g.lo < p < g.hi), looking back at most `limit` bytes.
KM_HD_NOINLINE bool line_start_bounded(const Genome& g, uint64_t p, uint64_t limit, uint64_t* out) {
    const uint64_t stop = (p - g.lo > limit) ? p - limit : g.lo;
    uint64_t ls = p;
    while (ls > stop && !is_term(g.b[ls - 1])) ls--;
    if (ls > g.lo && !is_term(g.b[ls - 1])) return false;
    *out = ls;
    return true;
}

// p's line starts at ls: 0 when that is a sequence line (or p is its first byte), else the header's end.
KM_HD_NOINLINE uint64_t header_until_from(const Genome& g, uint64_t p, uint64_t ls) {
    if (ls >= p || g.b[ls] != (uint8_t)'>') return 0;
    uint64_t e = p;
    while (e < g.hi && !is_term(g.b[e])) e++;
    return e + 1;
}

// First header line at or after `from` (used to skip text before the first record).
KM_HD_NOINLINE uint64_t first_header(const uint8_t* b, uint64_t from, uint64_t hi) {
    uint64_t p = from;
    while (p < hi) {
        if (b[p] == (uint8_t)'>' && (p == from || is_term(b[p - 1]))) return p;
        p++;
    }
    return hi;
}

// ---------------------------------------------------------------------------
// Dense walking
// ---------------------------------------------------------------------------
struct DenseParams {
    int k;            // the counted level (largest dense k of the call)
    uint32_t mask;    // 4^k - 1
    int min_rec;      // records shorter than this are dropped (>= k)
    int tails;        // emit run-end tails for levels tail_lo..k-1 (needed by the cascade)
    int tail_lo;      // lowest level the cascade descends to
};

struct WalkState {
    uint32_t kmer;
    int run;          // valid bases right before the current position (capped lookback + own bases)
    int in_hdr;
    int rec_known;    // 0 unknown, 1 record is long enough, 2 record is too short
};

// A run of valid bases ended just before `pos` (pos = the invalid symbol, the '>'
// of the next record, or g.hi).  Emits the last j-mer of the run for every
// j < k with run >= j: the j-mers that have no (j+1)-mer extension, which the
// marginalisation cascade cannot see.  `tails` is kept apart from the hot sink so
// that the sink's counters stay in registers (this function is never inlined).
template <class Tails>
KM_HD_NOINLINE void run_end_event(const Genome& g, uint64_t pos, const DenseParams& P, const Tails& tails) {
    uint64_t q = pos;
    uint32_t code = 0;
    int cnt = 0;
    while (cnt < P.k - 1) {
        int kind = prev_symbol(g, q);
        if (kind > 3) break;
        code |= (uint32_t)kind << (2 * cnt);
        cnt++;
    }
    if (cnt == 0 || cnt < P.tail_lo) return;
    uint64_t inside = pos;
    (void)prev_symbol(g, inside);                       // byte position of the run's last base
    if (!record_len_at_least(g, inside, P.min_rec)) return;
    for (int j = P.tail_lo > 1 ? P.tail_lo : 1; j <= cnt; j++) tails.tail(j, code & ((1u << (2 * j)) - 1u));
}

template <class Sink>
KM_HD void on_base(const Genome& g, uint64_t pos, int code, WalkState& s, const DenseParams& P, Sink& sink) {
    s.kmer = (s.kmer << 2) | (uint32_t)code;
    s.run++;
    if (s.run >= P.k) {
        if (P.min_rec > P.k) {                           // rare mode: k list mixes dense and larger k
            if (s.rec_known == 0) s.rec_known = (s.run >= P.min_rec || record_len_at_least(g, pos, P.min_rec)) ? 1 : 2;
            if (s.rec_known == 2) return;
        }
        sink.count(s.kmer & P.mask, pos);
    }
}

// One byte of the thread's own chunk (generic path).  run > 0 <=> the previous symbol
// is a valid base, so a non-base symbol ends a run exactly when run > 0.
template <class Sink, class Tails>
KM_HD void step_own(const Genome& g, uint64_t pos, uint32_t c, WalkState& s, const DenseParams& P, Sink& sink,
                    const Tails& tails) {
    int code = base_code(c);
    if (code >= 0 && !s.in_hdr) {
        on_base(g, pos, code, s, P, sink);
        return;
    }
    if (s.in_hdr) {
        if (is_term(c)) s.in_hdr = 0;
        return;
    }
    int kind = classify_nonbase(g, pos, c);
    if (kind == SYM_SKIP) return;
    if (s.run > 0 && P.tails) run_end_event(g, pos, P, tails);
    s.run = 0;
    if (kind == SYM_HDR) { s.in_hdr = 1; s.rec_known = 0; }
}

// Start state of a chunk that begins at cs (not inside a header line): the up to k-1
// valid bases that immediately precede it (run is capped at k-1, which is all the
// window test `run >= k` and the run-end test `run > 0` need).
KM_HD_NOINLINE void lookback(const Genome& g, uint64_t cs, const DenseParams& P, WalkState& s) {
    uint64_t q = cs;
    uint32_t code = 0;
    int cnt = 0;
    while (cnt < P.k - 1) {
        int kind = prev_symbol(g, q);
        if (kind > 3) break;
        code |= (uint32_t)kind << (2 * cnt);
        cnt++;
    }
    s.kmer = code;
    s.run = cnt;
}

// Generic walk of one chunk [cs, ce): the windows that END in it.  Bytes are read
// through `at(pos)`.
template <class Sink, class Tails, class ByteAt>
KM_HD void walk_chunk(const Genome& g, uint64_t cs, uint64_t ce, bool starts_in_header,
                      const DenseParams& P, Sink& sink, const Tails& tails, ByteAt&& at) {
    WalkState s;
    s.kmer = 0; s.run = 0; s.in_hdr = starts_in_header ? 1 : 0; s.rec_known = 0;
    if (!starts_in_header && cs > g.lo) lookback(g, cs, P, s);
    for (uint64_t pos = cs; pos < ce; pos++) step_own(g, pos, at(pos), s, P, sink, tails);
    // the genome ends with this chunk: a run that is still open ends here
    if (ce == g.hi && s.run > 0 && !s.in_hdr && P.tails) run_end_event(g, g.hi, P, tails);
}

// 4-entry byte LUT: byte i of the result = byte (sel nibble i) of `lut` (PRMT on the GPU).
KM_HD uint32_t prmt4(uint32_t lut, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(lut, 0u, sel);
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= ((lut >> (8 * ((sel >> (4 * i)) & 3u))) & 0xFFu) << (8 * i);
    return r;
#endif
}

// Per 32-bit word of FASTA bytes: 2-bit codes in bits [1:0] of every byte (y), a 0x80
// flag in every byte that is not one of AaCcGgTt (bad), and a non-zero `weird` when
// one of those is anything but '\n'.  ~18 integer ops per 4 bases, no table lookups.
KM_HD void classify_word(uint32_t x, uint32_t& y, uint32_t& bad, uint32_t& weird) {
    const uint32_t u = x & 0xDFDFDFDFu;                       // upper-case (generate.py:41)
    y = ((u >> 1) ^ (u >> 2)) & 0x03030303u;                   // A0 C1 G2 T3
    const uint32_t t = y | (y >> 4);
    const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);   // the 4 codes as PRMT selector nibbles
    const uint32_t d = prmt4(0x54474341u, sel) ^ u;            // "ACGT"[code] == byte ?
    bad = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    const uint32_t n = x ^ 0x0A0A0A0Au;
    const uint32_t not_nl = (((n & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | n) & 0x80808080u;
    weird = bad & not_nl;
}

// Classify the words of a chunk; returns non-zero when the chunk holds a byte that
// is neither a base nor '\n' (N runs, IUPAC codes, '>', '\r', blanks ...).
KM_HD uint32_t classify_chunk(const uint32_t* w, uint32_t* y, uint32_t* bad) {
    uint32_t weird = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < CHUNK / 4; i++) {
        uint32_t wd;
        classify_word(w[i], y[i], bad[i], wd);
        weird |= wd;
    }
    return weird;
}

// A clean chunk: its 31 or 32 bases as one top-aligned 64-bit string (base 0 in bits
// 63:62 of hi:lo), the '\n' (if any) squeezed out.
struct CleanChunk {
    uint32_t hi, lo;      // top-aligned 2-bit codes
    int n;                // 31 or 32 bases
    int nl;               // byte index of the removed '\n', 32 if none
    uint32_t last16;      // the last 16 bases, right-aligned: the carry of the next chunk
};

// Compact a classified chunk (classify_chunk returned 0) -- false when it holds more
// than one '\n' (lines shorter than 32 columns): the caller takes the generic path.
KM_HD bool pack_clean(const uint32_t* y, const uint32_t* bad, CleanChunk& c) {
    static_assert(CHUNK == 32, "pack_clean packs 32 bases into 64 bits");
    uint32_t pk[8], m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) {
        pk[i] = (y[i] * 0x40100401u) >> 24;                               // 4 codes -> 8 bits, first base on top
        m |= ((((bad[i] >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * i);   // bit b <=> byte b is the '\n'
    }
    uint64_t B = ((uint64_t)((pk[0] << 24) | (pk[1] << 16) | (pk[2] << 8) | pk[3]) << 32) |
                 (uint64_t)((pk[4] << 24) | (pk[5] << 16) | (pk[6] << 8) | pk[7]);
    if (m & (m - 1)) return false;                                        // two or more line feeds
    if (m == 0) {
        c.n = 32;
        c.nl = 32;
        c.last16 = (uint32_t)B;
    } else {
        int p = 0;
#if defined(__CUDA_ARCH__)
        p = __ffs((int)m) - 1;
#else
        while (!((m >> p) & 1u)) p++;
#endif
        const uint64_t upper = p ? (B >> (64 - 2 * p)) : 0ull;                            // bases 0 .. p-1
        const uint64_t lower = p == 31 ? 0ull : (B & ((1ull << (62 - 2 * p)) - 1ull));     // bases p+1 .. 31
        const uint64_t V = p == 31 ? upper : ((upper << (62 - 2 * p)) | lower);           // 31 bases, right-aligned
        c.n = 31;
        c.nl = p;
        c.last16 = (uint32_t)V;
        B = V << 2;
    }
    c.hi = (uint32_t)(B >> 32);
    c.lo = (uint32_t)B;
    return true;
}

// classify_chunk + pack_clean in one pass, for the counting kernels' hot loop: true when the chunk is clean (only
// bases and at most one '\n'), with the chunk packed into `c`.  Cheaper than the pair above: the flags of the bytes
// that are no bases are gathered into ONE word (bit 8 b + i <=> byte b of word i), so "how many, and where" is a
// test on that word, and only a chunk with exactly one such byte looks at it again to see that it is the '\n'
// (the pair above compares every byte with '\n' and builds a per-word nibble mask: ~9 more integer ops per word).
KM_HD bool classify_pack(const uint32_t* w, CleanChunk& c) {
    static_assert(CHUNK == 32, "classify_pack packs 32 bases into 64 bits");
    uint32_t pk[8], acc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) {
        const uint32_t u = w[i] & 0xDFDFDFDFu;                     // upper-case (generate.py:41)
        const uint32_t y = ((u >> 1) ^ (u >> 2)) & 0x03030303u;    // A0 C1 G2 T3
        const uint32_t t = y | (y >> 4);
#if defined(__CUDA_ARCH__)
        const uint32_t sel = __byte_perm(t, 0u, 0x4420u);          // the 4 codes as PRMT selector nibbles
#else
        const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);
#endif
        const uint32_t d = prmt4(0x54474341u, sel) ^ u;            // "ACGT"[code] == byte ?
        const uint32_t bad = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
        acc |= bad >> (7 - i);
        pk[i] = (y * 0x40100401u) >> 24;                           // 4 codes -> 8 bits, first base on top
    }
    uint64_t B = ((uint64_t)((pk[0] << 24) | (pk[1] << 16) | (pk[2] << 8) | pk[3]) << 32) |
                 (uint64_t)((pk[4] << 24) | (pk[5] << 16) | (pk[6] << 8) | pk[7]);
    if (acc == 0) {
        c.n = 32;
        c.nl = 32;
        c.last16 = (uint32_t)B;
    } else {
        if (acc & (acc - 1)) return false;                         // two or more bytes that are no bases
        int bit = 0;
#if defined(__CUDA_ARCH__)
        bit = __ffs((int)acc) - 1;
#else
        while (!((acc >> bit) & 1u)) bit++;
#endif
        const int i = bit & 7, b = bit >> 3;
        const uint32_t a0 = (i & 1) ? w[1] : w[0], a1 = (i & 1) ? w[3] : w[2];
        const uint32_t a2 = (i & 1) ? w[5] : w[4], a3 = (i & 1) ? w[7] : w[6];
        const uint32_t b0 = (i & 2) ? a1 : a0, b1 = (i & 2) ? a3 : a2;
        const uint32_t word = (i & 4) ? b1 : b0;
        if (((word >> (8 * b)) & 0xFFu) != 0x0Au) return false;    // ... and that one must be the line feed
        const int p = 4 * i + b;
        const uint64_t upper = p ? (B >> (64 - 2 * p)) : 0ull;                            // bases 0 .. p-1
        const uint64_t lower = p == 31 ? 0ull : (B & ((1ull << (62 - 2 * p)) - 1ull));     // bases p+1 .. 31
        const uint64_t V = p == 31 ? upper : ((upper << (62 - 2 * p)) | lower);           // 31 bases, right-aligned
        c.n = 31;
        c.nl = p;
        c.last16 = (uint32_t)V;
        B = V << 2;
    }
    c.hi = (uint32_t)(B >> 32);
    c.lo = (uint32_t)B;
    return true;
}

KM_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int s) {                 // bits [s, s+32) of hi:lo, 0 <= s < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, (unsigned)s);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> s);
#endif
}

// A sink that declares `static constexpr bool raw_windows = true` receives the windows of a clean chunk
// UNMASKED (the k-mer in the low 2k bits, earlier bases above it) and masks what it needs itself.
template <class S, class = void>
struct sink_raw_windows : std::false_type {};
template <class S>
struct sink_raw_windows<S, std::void_t<decltype(S::raw_windows)>> : std::integral_constant<bool, S::raw_windows> {};

// Every window that ends in a clean chunk whose predecessor chunk is clean too:
// carry16 = the predecessor's last 16 bases (k <= 16).  Two integer ops per window.
template <class Sink>
KM_HD void emit_clean(const CleanChunk& c, uint32_t carry16, uint64_t cs, const DenseParams& P, Sink& sink) {
    const uint32_t mask = sink_raw_windows<Sink>::value ? 0xFFFFFFFFu : P.mask;
    auto window = [&](int j) -> uint32_t {
        const int sft = 62 - 2 * j;                                        // bit position of base j in hi:lo
        return (sft >= 32 ? funnel_r(c.hi, carry16, sft - 32) : funnel_r(c.lo, c.hi, sft)) & mask;
    };
    auto at = [&](int j) -> uint64_t { return cs + (uint64_t)(j + (j >= c.nl ? 1 : 0)); };   // byte of base j
    // a clean chunk has >= 31 bases; windows go to the sink eight at a time so that a sink with
    // returning atomics (the partition path) can keep eight of them in flight
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 24; j += 8) {
        const uint32_t wv[8] = {window(j), window(j + 1), window(j + 2), window(j + 3),
                                window(j + 4), window(j + 5), window(j + 6), window(j + 7)};
        const uint64_t pv[8] = {at(j), at(j + 1), at(j + 2), at(j + 3), at(j + 4), at(j + 5), at(j + 6), at(j + 7)};
        sink.count8(wv, pv);
    }
    {   // windows 24..30 and, when the chunk holds 32 bases, the 32nd: one more batch of eight
        const uint32_t wv[8] = {window(24), window(25), window(26), window(27), window(28), window(29), window(30),
                                c.lo & mask};
        const uint64_t pv[8] = {at(24), at(25), at(26), at(27), at(28), at(29), at(30), cs + 31};
        sink.count8_tail(wv, pv, c.n == 32);
    }
}

// Does any byte of the chunk's words equal the byte replicated in `pattern`?
KM_HD bool any_byte_eq_chunk(const uint32_t* w, uint32_t pattern) {
    uint32_t acc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < CHUNK / 4; i++) {
        const uint32_t x = w[i] ^ pattern;
        acc |= (x - 0x01010101u) & ~x & 0x80808080u;           // a zero byte in x
    }
    return acc != 0;
}

// Header lines that START in [cs, ce): callback(h, until) with the header
// occupying [h, until) (until = one past its terminator, or > g.hi at EOF).
template <class F>
KM_HD void find_headers(const Genome& g, uint64_t cs, uint64_t ce, F&& on_header) {
    for (uint64_t pos = cs; pos < ce; pos++) {
        if (g.b[pos] != (uint8_t)'>') continue;
        if (!(pos == g.lo || is_term(g.b[pos - 1]))) continue;
        uint64_t e = pos + 1;
        while (e < g.hi && !is_term(g.b[e])) e++;
        on_header(pos, e + 1);
        pos = e;                                         // nothing inside a header line starts a record
    }
}

// First pass over one slice start p (one thread per slice).
KM_HD_NOINLINE void slice_head_quick(const Genome& g, uint64_t p, uint64_t limit, SliceHead* h) {
    h->hdr_until = 0;
    h->prev_ok = 0;
    h->prev16 = 0;
    h->line_start = g.lo;                 // nothing of the genome lies before p: by convention
    if (p <= g.lo || p >= g.hi) return;
    uint64_t ls = 0;
    const bool known = line_start_bounded(g, p, limit, &ls);
    h->line_start = known ? ls : LS_UNRESOLVED;
    if (known) h->hdr_until = header_until_from(g, p, ls);
    if (p < g.lo + CHUNK) return;
    // the chunk right before the slice: packed once so that the slice's first chunk can take the clean
    // path like every other chunk (it must be sequence, not header text made of base letters)
    uint32_t w[CHUNK / 4], y[CHUNK / 4], bad[CHUNK / 4];
    const uint64_t pp = p - CHUNK;
    for (int j = 0; j < CHUNK / 4; j++) {
        const uint8_t* q = g.b + pp + 4 * j;
        w[j] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    }
    CleanChunk pc;
    if (!classify_pack(w, pc)) return;
    bool in_hdr;
    if (!known) {
        in_hdr = false;                   // same line as p (limit >= CHUNK): the table pass clears prev_ok if it is a header
    } else if (ls <= pp) {
        in_hdr = ls < pp && g.b[ls] == (uint8_t)'>';
    } else {                              // a line ends inside the chunk: pp belongs to the line before
        uint64_t ls2 = 0;
        if (pp <= g.lo) in_hdr = false;
        else if (!line_start_bounded(g, pp, limit, &ls2)) in_hdr = true;        // give up: generic path
        else in_hdr = ls2 < pp && g.b[ls2] == (uint8_t)'>';
    }
    if (!in_hdr) {
        h->prev_ok = 1;
        h->prev16 = pc.last16;
    }
}

}  // namespace km
