/*
 * kmer_oracle.c -- CPU restatement of kmer-ml's k-mer extraction path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (kmerml_b200/) never links, imports or falls back to it.
 *
 * Parity status: PINNED against the unmodified reference run under the
 * Bio.SeqIO shim (tests/golden/make_golden.py -> tests/golden/ JSON files, checked
 * by tests/test_oracle_golden.py).  The reference's own tests hold no
 * assertions or golden vectors, and FASTA corner cases go
 * through biopython, which is absent from /root/reference: the parser below
 * restates biopython==1.85's published plain-FASTA text iterator
 * (requirements.txt:1) -- that part is "parity unpinned".
 *
 * What each function follows (paths relative to /root/reference):
 *   kmo_parse          Bio.SeqIO.parse(..., "fasta") as called at
 *                      kmerml/kmers/generate.py:39 (see tests/_ref/Bio/SeqIO.py)
 *   kmo_count_dense    kmerml/kmers/generate.py:36-58 (upper-case :41, short
 *                      record skip :44-46, window loop :51-52, ACGT filter
 *                      :55-56, count :58), dict insertion order :88
 *   kmo_count_sparse   same, for k up to 32 (sorted instead of a dense array)
 *
 * k-mer index convention used everywhere in this repo: lexicographic ACGT,
 * A=0 C=1 G=2 T=3, first base most significant.  The file digit code
 * A0 T1 C2 G3 (generate.py:71) is applied only by writers.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KMO_NONE UINT64_MAX

static int is_py_space(uint8_t c) {
    /* ASCII subset of str.isspace(): what str.rstrip() removes */
    return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x1F) || c == 0x20;
}

/*
 * Text-mode line iteration (universal newlines): a line ends at "\n", "\r\n" or a
 * lone "\r".  Returns the index one past the line's terminator; *content_end is
 * the index of the terminator (== n when the last line has none).
 */
static uint64_t next_line(const uint8_t *b, uint64_t n, uint64_t start, uint64_t *content_end) {
    uint64_t i = start;
    while (i < n && b[i] != '\n' && b[i] != '\r') i++;
    *content_end = i;
    if (i >= n) return n;
    if (b[i] == '\r' && i + 1 < n && b[i + 1] == '\n') return i + 2;
    return i + 1;
}

/*
 * Parse FASTA bytes into records.  seq receives the concatenated record
 * sequences (case preserved), rec_seq_off/rec_len locate record r inside seq,
 * rec_hdr_off is the byte offset of its '>' in the input.  Arrays may be NULL
 * (then only the count is returned).  Returns the number of records, or -1 if
 * max_rec is too small.
 */
int64_t kmo_parse(const uint8_t *b, uint64_t n, uint8_t *seq, uint64_t *rec_seq_off,
                  uint64_t *rec_len, uint64_t *rec_hdr_off, uint64_t max_rec) {
    int64_t nrec = 0;
    uint64_t out = 0, pos = 0, cur_start = 0;
    int in_record = 0;
    while (pos < n) {
        uint64_t ce, nx = next_line(b, n, pos, &ce);
        if (b[pos] == '>') {                       /* line[:1] == ">" */
            if (in_record && rec_len) rec_len[nrec - 1] = out - cur_start;
            if (rec_hdr_off && (uint64_t)nrec >= max_rec) return -1;
            if (rec_hdr_off) rec_hdr_off[nrec] = pos;
            if (rec_seq_off) rec_seq_off[nrec] = out;
            cur_start = out;
            nrec++;
            in_record = 1;
        } else if (in_record) {
            uint64_t e = ce;                       /* line.rstrip() */
            while (e > pos && is_py_space(b[e - 1])) e--;
            for (uint64_t i = pos; i < e; i++) {   /* .replace(" ", "").replace("\r", "") */
                if (b[i] == ' ' || b[i] == '\r') continue;
                if (seq) seq[out] = b[i];
                out++;
            }
        }
        pos = nx;
    }
    if (in_record && rec_len) rec_len[nrec - 1] = out - cur_start;
    return nrec;
}

static inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

static inline int code_of(uint8_t c) {             /* after upper-casing */
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default:  return -1;
    }
}

typedef struct {
    uint8_t *seq;
    uint64_t *off, *len;
    int64_t nrec;
} parsed_t;

static int parse_all(const uint8_t *b, uint64_t n, parsed_t *p) {
    int64_t nrec = kmo_parse(b, n, NULL, NULL, NULL, NULL, 0);
    p->nrec = nrec;
    p->seq = (uint8_t *)malloc(n ? n : 1);
    p->off = (uint64_t *)malloc(sizeof(uint64_t) * (nrec ? nrec : 1));
    p->len = (uint64_t *)malloc(sizeof(uint64_t) * (nrec ? nrec : 1));
    uint64_t *hdr = (uint64_t *)malloc(sizeof(uint64_t) * (nrec ? nrec : 1));
    if (!p->seq || !p->off || !p->len || !hdr) return -1;
    kmo_parse(b, n, p->seq, p->off, p->len, hdr, (uint64_t)nrec);
    free(hdr);
    return 0;
}

static void parsed_free(parsed_t *p) { free(p->seq); free(p->off); free(p->len); }

/*
 * Dense forward-strand counts of one k (k <= 15), generate.py:36-58.
 * counts[4^k] (zeroed here).  first_ord[4^k] (optional) receives the ordinal of
 * the window that first inserted each k-mer into the dict (KMO_NONE if never),
 * i.e. dict insertion order of generate.py:88.  min_len is max(k_values).
 * Returns the number of counted windows, or -1 on error.
 */
int64_t kmo_count_dense(const uint8_t *b, uint64_t n, int k, int min_len, uint64_t *counts,
                        uint64_t *first_ord) {
    if (k < 1 || k > 15) return -1;
    uint64_t nb = 1ull << (2 * k);
    memset(counts, 0, nb * sizeof(uint64_t));
    if (first_ord) for (uint64_t i = 0; i < nb; i++) first_ord[i] = KMO_NONE;
    parsed_t p;
    if (parse_all(b, n, &p)) return -1;
    int64_t windows = 0;
    uint64_t ordinal = 0;
    for (int64_t r = 0; r < p.nrec; r++) {
        uint8_t *s = p.seq + p.off[r];
        uint64_t L = p.len[r];
        for (uint64_t i = 0; i < L; i++) s[i] = up(s[i]);        /* :41 */
        if (L < (uint64_t)min_len) continue;                      /* :44-46 */
        if (L < (uint64_t)k) continue;
        for (uint64_t i = 0; i + k <= L; i++, ordinal++) {        /* :51 */
            uint64_t idx = 0;
            int ok = 1;
            for (int j = 0; j < k; j++) {                         /* :55 */
                int c = code_of(s[i + j]);
                if (c < 0) { ok = 0; break; }
                idx = (idx << 2) | (uint64_t)c;
            }
            if (!ok) continue;
            if (first_ord && counts[idx] == 0) first_ord[idx] = ordinal;
            counts[idx]++;                                        /* :58 */
            windows++;
        }
    }
    parsed_free(&p);
    return windows;
}

/* Same, all k of a list into one concatenated uint64 row (no first_ord). */
int64_t kmo_count_dense_multi(const uint8_t *b, uint64_t n, const int *ks, int nk, int min_len,
                              uint64_t *row) {
    int64_t total = 0;
    uint64_t off = 0;
    for (int i = 0; i < nk; i++) {
        int64_t w = kmo_count_dense(b, n, ks[i], min_len, row + off, NULL);
        if (w < 0) return -1;
        total += w;
        off += 1ull << (2 * ks[i]);
    }
    return total;
}

typedef struct { uint64_t code, ord; } pair_t;

static int cmp_code_ord(const void *a, const void *b) {
    const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    return x->ord < y->ord ? -1 : (x->ord > y->ord);
}

typedef struct { uint64_t code, count, ord; } trip_t;

static int cmp_ord(const void *a, const void *b) {
    const trip_t *x = (const trip_t *)a, *y = (const trip_t *)b;
    return x->ord < y->ord ? -1 : (x->ord > y->ord);
}

/*
 * Any k <= 32: distinct k-mers (codes, counts) in dict insertion order.
 * Returns the number of distinct k-mers (may exceed cap: then only cap are
 * written), or -1 on error.  *windows_out gets the counted windows.
 */
int64_t kmo_count_sparse(const uint8_t *b, uint64_t n, int k, int min_len, uint64_t *codes,
                         uint64_t *counts, uint64_t cap, uint64_t *windows_out) {
    if (k < 1 || k > 32) return -1;
    parsed_t p;
    if (parse_all(b, n, &p)) return -1;
    uint64_t total = 0;
    for (int64_t r = 0; r < p.nrec; r++) total += p.len[r];
    pair_t *v = (pair_t *)malloc(sizeof(pair_t) * (total ? total : 1));
    if (!v) { parsed_free(&p); return -1; }
    uint64_t m = 0, ordinal = 0;
    for (int64_t r = 0; r < p.nrec; r++) {
        uint8_t *s = p.seq + p.off[r];
        uint64_t L = p.len[r];
        for (uint64_t i = 0; i < L; i++) s[i] = up(s[i]);
        if (L < (uint64_t)min_len || L < (uint64_t)k) continue;
        for (uint64_t i = 0; i + k <= L; i++, ordinal++) {
            uint64_t idx = 0;
            int ok = 1;
            for (int j = 0; j < k; j++) {
                int c = code_of(s[i + j]);
                if (c < 0) { ok = 0; break; }
                idx = (idx << 2) | (uint64_t)c;
            }
            if (!ok) continue;
            v[m].code = idx;
            v[m].ord = ordinal;
            m++;
        }
    }
    parsed_free(&p);
    if (windows_out) *windows_out = m;
    qsort(v, m, sizeof(pair_t), cmp_code_ord);
    trip_t *t = (trip_t *)malloc(sizeof(trip_t) * (m ? m : 1));
    if (!t) { free(v); return -1; }
    uint64_t d = 0;
    for (uint64_t i = 0; i < m;) {
        uint64_t j = i;
        while (j < m && v[j].code == v[i].code) j++;
        t[d].code = v[i].code;
        t[d].count = j - i;
        t[d].ord = v[i].ord;
        d++;
        i = j;
    }
    free(v);
    qsort(t, d, sizeof(trip_t), cmp_ord);
    for (uint64_t i = 0; i < d && i < cap; i++) {
        codes[i] = t[i].code;
        counts[i] = t[i].count;
    }
    free(t);
    return (int64_t)d;
}

/* Base tallies as kmerml/utils/genome_metadata.py:55-85 computes them. */
int kmo_genome_stats(const uint8_t *b, uint64_t n, uint64_t *contigs, uint64_t *total_size,
                     uint64_t *gc, uint64_t *n_count) {
    parsed_t p;
    if (parse_all(b, n, &p)) return -1;
    *contigs = (uint64_t)p.nrec;
    *total_size = *gc = *n_count = 0;
    for (int64_t r = 0; r < p.nrec; r++) {
        uint8_t *s = p.seq + p.off[r];
        for (uint64_t i = 0; i < p.len[r]; i++) {
            uint8_t c = up(s[i]);
            if (c == 'G' || c == 'C') (*gc)++;
            if (c == 'N') (*n_count)++;
        }
        *total_size += p.len[r];
    }
    parsed_free(&p);
    return 0;
}

int kmo_version(void) { return 1; }
