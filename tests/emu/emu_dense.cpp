// CPU emulation of the GPU thread decomposition of the dense counting path.
// TEST INFRASTRUCTURE: compiles kmerml_b200/csrc/fasta_walk.cuh with g++ and
// runs every "thread" of every tile sequentially, with the same slice / tile /
// chunk / header-flag structure the kernels in dense_kernels.cu use.  It lets
// the carry-free walking logic be fuzzed against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kmerml_b200/csrc/fasta_walk.cuh"

using namespace km;

namespace {
struct EmuSink {
    std::vector<uint64_t>* top;
    std::vector<std::vector<uint64_t>>* tails;
    std::vector<uint64_t>* first;     // optional first-occurrence (min end offset)
    uint64_t n_count = 0;
    void count(uint32_t idx, uint64_t pos) {
        (*top)[idx]++;
        n_count++;
        if (first && pos < (*first)[idx]) (*first)[idx] = pos;
    }
    void count4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t pa, uint64_t pb, uint64_t pc, uint64_t pd) {
        count(a, pa); count(b, pb); count(c, pc); count(d, pd);
    }
    void count8(const uint32_t* w, const uint64_t* p) { for (int i = 0; i < 8; i++) count(w[i], p[i]); }
    void count8_tail(const uint32_t* w, const uint64_t* p, bool last) { for (int i = 0; i < (last ? 8 : 7); i++) count(w[i], p[i]); }
    void tail(int j, uint32_t idx) const { (*tails)[j][idx]++; }
};
}  // namespace

static uint64_t g_line_limit = km::LINE_SCAN_LIMIT;
static uint64_t g_slice_bytes = 0;
extern "C" void emu_set_slice_bytes(uint64_t v) { g_slice_bytes = v; }
extern "C" void emu_set_line_limit(uint64_t v) { g_line_limit = v ? v : km::LINE_SCAN_LIMIT; }
extern "C" {

// bytes[0..n): one FASTA file placed at `base_off` inside a zero-padded buffer
// (exercises unaligned genome starts).  Counts level kmax with the tile/slice
// geometry given, then cascades to every level 1..kmax.  out_levels receives
// levels 1..kmax concatenated (uint64).  Returns counted windows at kmax.
static int64_t emu_core(const uint8_t* bytes, uint64_t n, uint64_t base_off, int kmax, int min_rec,
                        int threads_per_tile, int tiles_per_slice, uint64_t* out_levels,
                        uint64_t* out_first /* 4^kmax or NULL */, uint64_t range_begin, uint64_t range_end) {
    std::vector<uint8_t> buf(base_off + n + 256, 0);
    if (n) memcpy(buf.data() + base_off, bytes, n);
    Genome g;
    g.b = buf.data();
    g.hi = base_off + n;
    g.lo = first_header(g.b, base_off, g.hi);

    DenseParams P;
    P.k = kmax;
    P.mask = (kmax >= 16) ? 0xFFFFFFFFu : ((1u << (2 * kmax)) - 1u);
    P.min_rec = min_rec;
    P.tails = kmax > 1;
    P.tail_lo = 1;

    std::vector<uint64_t> top(1ull << (2 * kmax), 0);
    std::vector<std::vector<uint64_t>> tails(kmax + 1);
    for (int j = 1; j < kmax; j++) tails[j].assign(1ull << (2 * j), 0);
    std::vector<uint64_t> first;
    if (out_first) first.assign(1ull << (2 * kmax), UINT64_MAX);
    EmuSink sink;
    sink.top = &top;
    sink.tails = &tails;
    sink.first = out_first ? &first : nullptr;

    const uint64_t tile_bytes = (uint64_t)threads_per_tile * CHUNK;
    // g_slice_bytes != 0: slices that are NOT a whole number of tiles (count8_kernel: 32 KB tiles over slices
    // of 16 KB units); the chunks of the last tile are then clipped to the slice end (walk_slice's CLIP)
    const uint64_t slice_bytes = g_slice_bytes ? g_slice_bytes : tile_bytes * (uint64_t)tiles_per_slice;
    // slices are aligned to absolute multiples of slice_bytes (as on the GPU)
    uint64_t first_slice = g.lo / slice_bytes;
    // slice table passes (slice_header_kernel, slice_long_scan_kernel, slice_long_resolve_kernel)
    std::vector<uint64_t> starts;
    for (uint64_t sb = first_slice * slice_bytes; sb < g.hi; sb += slice_bytes) {
        // a byte range (relative to the file start) selects whole slices: the multi-GPU unit
        if (range_end && (sb + slice_bytes <= base_off + range_begin || sb >= base_off + range_end)) continue;
        starts.push_back(sb);
    }
    std::vector<SliceHead> heads(starts.size());
    std::vector<uint64_t> scan_last(starts.size(), LS_UNRESOLVED);
    for (size_t i = 0; i < starts.size(); i++) slice_head_quick(g, starts[i], g_line_limit, &heads[i]);
    for (size_t i = 0; i < starts.size(); i++) {
        if (heads[i].line_start != LS_UNRESOLVED) continue;
        uint64_t lo = g.lo;
        if (i > 0 && starts[i - 1] > lo) lo = starts[i - 1];
        uint64_t best = 0;
        for (uint64_t pos = starts[i]; pos > lo; pos--)
            if (is_term(g.b[pos - 1])) { best = pos; break; }
        scan_last[i] = best ? best : (lo == g.lo ? g.lo : LS_UNRESOLVED);
    }
    {
        uint64_t run_max = 0;
        for (size_t i = 0; i < starts.size(); i++) {
            uint64_t cand = heads[i].line_start != LS_UNRESOLVED ? heads[i].line_start : scan_last[i];
            if (cand != LS_UNRESOLVED && cand > run_max) run_max = cand;
            if (heads[i].line_start == LS_UNRESOLVED) {
                uint64_t ls = std::max(run_max, g.lo);
                heads[i].hdr_until = header_until_from(g, starts[i], ls);
                if (heads[i].hdr_until) heads[i].prev_ok = 0;
            }
        }
    }
    for (size_t si = 0; si < starts.size(); si++) {
        const uint64_t sb = starts[si];
        uint64_t hdr_carry = heads[si].hdr_until;
        std::vector<uint8_t> flags(threads_per_tile), clean(threads_per_tile);
        std::vector<CleanChunk> cc(threads_per_tile);
        // last chunk before the current tile: clean and in sequence?  For the slice's first tile
        // the chunk belongs to another CTA (slice table).
        bool prev_tile_ok = heads[si].prev_ok != 0;
        uint32_t prev_tile_last16 = heads[si].prev16;
        for (uint64_t tb = sb; tb < std::min(sb + slice_bytes, g.hi); tb += tile_bytes) {
            std::fill(flags.begin(), flags.end(), 0);
            std::fill(clean.begin(), clean.end(), 0);
            uint64_t next_carry = hdr_carry;
            auto clip = [&](int t, uint64_t& cs, uint64_t& ce) {
                cs = std::max<uint64_t>(tb + (uint64_t)t * CHUNK, g.lo);
                ce = std::min<uint64_t>(tb + (uint64_t)(t + 1) * CHUNK, g.hi);
                return cs < ce && tb + (uint64_t)t * CHUNK < sb + slice_bytes;
            };
            for (int t = 0; t < threads_per_tile; t++) {            // phase 1: classify, pack, find header lines
                uint64_t cs, ce;
                if (!clip(t, cs, ce)) continue;
                auto on_header = [&](uint64_t h, uint64_t until) {
                    (void)h;
                    for (int j = t + 1; j < threads_per_tile && tb + (uint64_t)j * CHUNK < until; j++) flags[j] = 1;
                    next_carry = std::max(next_carry, until);
                };
                if (ce - cs == CHUNK) {
                    uint32_t w[CHUNK / 4];
                    memcpy(w, g.b + cs, CHUNK);
                    clean[t] = classify_pack(w, cc[t]);
                    if (!clean[t] && any_byte_eq_chunk(w, 0x3E3E3E3Eu)) find_headers(g, cs, ce, on_header);
                } else {
                    find_headers(g, cs, ce, on_header);
                }
            }
            bool last_ok = false;
            uint32_t last_l16 = 0;
            for (int t = 0; t < threads_per_tile; t++) {            // phase 2: walk
                uint64_t cs, ce;
                const bool has = clip(t, cs, ce);
                const bool in_hdr = flags[t] || cs < hdr_carry;
                bool prev_ok;
                uint32_t carry;
                if (t > 0) {
                    prev_ok = clean[t - 1] && !flags[t - 1] && !(tb + (uint64_t)(t - 1) * CHUNK < hdr_carry);
                    carry = cc[t - 1].last16;
                } else {
                    prev_ok = prev_tile_ok;
                    carry = prev_tile_last16;
                }
                if (has) {
                    if (clean[t] && !in_hdr && prev_ok && P.min_rec <= P.k) {
                        emit_clean(cc[t], carry, cs, P, sink);       // the path nearly every GPU thread takes
                        if (ce == g.hi && P.tails) run_end_event(g, g.hi, P, sink);
                    } else {
                        walk_chunk(g, cs, ce, in_hdr, P, sink, sink, [&](uint64_t pos) -> uint32_t { return g.b[pos]; });
                    }
                }
                if (t == threads_per_tile - 1) {
                    last_ok = has && clean[t] && !in_hdr;
                    last_l16 = cc[t].last16;
                }
            }
            prev_tile_ok = last_ok;
            prev_tile_last16 = last_l16;
            hdr_carry = next_carry;
        }
    }
    // cascade: c_j[p] = sum_b c_{j+1}[4p+b] + tails_j[p]
    std::vector<std::vector<uint64_t>> lv(kmax + 1);
    lv[kmax] = top;
    for (int j = kmax - 1; j >= 1; j--) {
        lv[j].assign(1ull << (2 * j), 0);
        for (uint64_t p = 0; p < lv[j].size(); p++)
            lv[j][p] = lv[j + 1][4 * p] + lv[j + 1][4 * p + 1] + lv[j + 1][4 * p + 2] + lv[j + 1][4 * p + 3] + tails[j][p];
    }
    uint64_t off = 0;
    for (int j = 1; j <= kmax; j++) {
        memcpy(out_levels + off, lv[j].data(), lv[j].size() * sizeof(uint64_t));
        off += lv[j].size();
    }
    if (out_first) memcpy(out_first, first.data(), first.size() * sizeof(uint64_t));
    return (int64_t)sink.n_count;
}

int64_t emu_count_dense(const uint8_t* bytes, uint64_t n, uint64_t base_off, int kmax, int min_rec,
                        int threads_per_tile, int tiles_per_slice, uint64_t* out_levels, uint64_t* out_first) {
    return emu_core(bytes, n, base_off, kmax, min_rec, threads_per_tile, tiles_per_slice, out_levels, out_first, 0, 0);
}

// Only the slices inside [range_begin, range_end) (multiples of the slice size, base_off = 0).
int64_t emu_count_dense_range(const uint8_t* bytes, uint64_t n, int kmax, int min_rec, int threads_per_tile,
                              int tiles_per_slice, uint64_t range_begin, uint64_t range_end, uint64_t* out_levels) {
    return emu_core(bytes, n, 0, kmax, min_rec, threads_per_tile, tiles_per_slice, out_levels, nullptr, range_begin,
                    range_end);
}
}
