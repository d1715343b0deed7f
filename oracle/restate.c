/*
 * TEST INFRASTRUCTURE (oracle) -- plain-C restatement of the leaf algorithms of the
 * stage-1 assembly hot path.  Not part of the product; only tests/, smoke() and the
 * cpu_baseline leg of bench.py may load it.
 *
 * The full path is checked against the REAL reference (oracle/_ref/libt4ref.so, built
 * from /root/reference by oracle/Makefile).  This file restates the self-contained
 * pieces independently of both the reference build and the CUDA engine, so that the
 * GPU box can cross-check them even if oracle/_ref is absent, and so that the
 * restatement itself is pinned against the reference by tests/test_oracle.py:
 *   - rolling 2-bit k-mer with N tracking            KmerCode.hpp:94-109
 *   - which k-mers of a contig get indexed           KmerIndex.hpp:118-141
 *   - IsBaseEqual on a posWeight column              AlignAlgo.hpp:49-55
 *   - GlobalAlignment_PosWeight (full matrix)        AlignAlgo.hpp:57-216
 *   - ComputeNomatchGapLimit                         SeqSet.hpp:2476-2482
 *   - union length of k-mer hits                     SeqSet.hpp:3330-3350
 *   - HasMotif                                       SeqSet.hpp:5029-5073
 * Pinned: yes (tests/test_oracle.py compares every function with the compiled reference
 * on seeded inputs and with tests/golden/dp_cases.json.gz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int nuc(char c) /* nucToNum[c - 'A'] & 3, main.cpp:39-42 */
{
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    case 'N': return 0;
    default: return 3;
    }
}

/* KmerCode.hpp:94-109 applied to s[0..len): codes[i], valid[i] for the k-mer ending at i (i >= k-1). */
void rs_kmer_codes(const char *s, int len, int k, uint64_t *codes, unsigned char *valid)
{
    uint64_t mask = k < 32 ? ((1ull << (2 * k)) - 1ull) : ~0ull;
    uint64_t code = 0;
    int invalid_pos = -1;
    for (int i = 0; i < len; ++i) {
        if (invalid_pos != -1)
            ++invalid_pos;
        code = ((code << 2) & mask) | (uint64_t)nuc(s[i]);
        if (s[i] == 'N')
            invalid_pos = 0;
        if (invalid_pos >= k)
            invalid_pos = -1;
        codes[i] = code;
        valid[i] = (invalid_pos == -1);
    }
}

/* KmerIndex.hpp:118-141: offsets (and codes) BuildIndexFromRead inserts for s.  Returns the count. */
int rs_index_offsets(const char *s, int len, int k, int shift, int32_t *offsets, uint64_t *codes_out)
{
    if (len < k)
        return 0;
    uint64_t *codes = malloc(sizeof(uint64_t) * len);
    unsigned char *valid = malloc(len);
    rs_kmer_codes(s, len, k, codes, valid);
    uint64_t prev = 0;
    int n = 0;
    for (int i = k - 1; i < len; ++i) {
        if (valid[i] && (i == k || codes[i] != prev)) {
            offsets[n] = i - k + 1 + shift;
            codes_out[n] = codes[i];
            ++n;
        }
        prev = codes[i];
    }
    free(codes);
    free(valid);
    return n;
}

/* AlignAlgo.hpp:49-55 */
int rs_base_equal(const int32_t *w, char c)
{
    int sum = w[0] + w[1] + w[2] + w[3];
    return sum == 0 || c == 'N' || sum < 3 * w[nuc(c)];
}

/* AlignAlgo.hpp:57-216, full (lenp+1) x (lent+1) matrix like the reference.  align needs lent+lenp+2 entries. */
int rs_dp_pos_weight(const int32_t *tw, int lent, const char *p, int lenp, signed char *align)
{
    enum { M = 0, X = 1, INS = 2, DEL = 3 };
    if (lent == 0 || lenp == 0) {
        align[0] = -1;
        return 0;
    }
    if (lent == 1 && lenp == 1) {
        int eq = rs_base_equal(tw, p[0]);
        align[0] = eq ? M : X;
        align[1] = -1;
        return eq ? 2 : -2;
    }
    if (lent == lenp) {
        int score = 0, i;
        for (i = 0; i < lent; ++i) {
            int eq = rs_base_equal(tw + 4 * i, p[i]);
            align[i] = eq ? M : X;
            score += eq ? 2 : -2;
        }
        align[i] = -1;
        if (score >= 2 * lent - 8)
            return score;
    }
    int left = 5, right = 5;
    if (lent > lenp)
        right += lent - lenp;
    else if (lent < lenp)
        left += lenp - lent;
    int neg_inf = (lent + 1) * (lenp + 1) * -4;
    int bmax = lent + 1;
    int *m = calloc((size_t)(lenp + 1) * (lent + 1), sizeof(int));
    m[0] = 0;
    for (int i = 1; i <= lenp; ++i)
        m[i * bmax] = -4 + i * -4;
    for (int j = 1; j <= lent; ++j)
        m[j] = -4 + j * -4;
    for (int i = 1; i <= lenp; ++i) {
        int start = i - left < 1 ? 1 : i - left;
        int end = i + right > lent ? lent : i + right;
        if (start > 1)
            m[i * bmax + start - 1] = neg_inf;
        if (end < lent)
            m[i * bmax + end + 1] = neg_inf;
        for (int j = start; j <= end; ++j) {
            int s = m[(i - 1) * bmax + j - 1] + (rs_base_equal(tw + 4 * (j - 1), p[i - 1]) ? 2 : -2);
            int l = m[i * bmax + j - 1] - 4;
            int u = m[(i - 1) * bmax + j] - 4;
            if (l > s) s = l;
            if (u > s) s = u;
            m[i * bmax + j] = s;
        }
    }
    int ret = m[lenp * bmax + lent];
    int ti = lenp, tj = lent, tag = 0;
    while (ti > 0 || tj > 0) {
        int max = m[ti * bmax + tj];
        int a = 0;
        if (tj > 0 && m[ti * bmax + tj - 1] - 4 == max)
            a = DEL;
        if (ti > 0 && m[(ti - 1) * bmax + tj] - 4 == max)
            a = INS;
        if (tj > 0 && ti > 0) {
            int diff = rs_base_equal(tw + 4 * (tj - 1), p[ti - 1]) ? 2 : -2;
            if (m[(ti - 1) * bmax + tj - 1] + diff == max)
                a = diff == 2 ? M : X;
        }
        align[tag++] = (signed char)a;
        if (a == DEL)
            --tj;
        else if (a == INS)
            --ti;
        else {
            --ti;
            --tj;
        }
    }
    align[tag] = -1;
    for (int i = 0, j = tag - 1; i < j; ++i, --j) {
        signed char t = align[i];
        align[i] = align[j];
        align[j] = t;
    }
    free(m);
    return ret;
}

/* SeqSet.hpp:2476-2482 */
int rs_nomatch_gap_limit(int k)
{
    double p = pow(0.8, k);
    return (int)(k * (log(0.01) / log(1 - p))) + 1;
}

/* SeqSet.hpp:3330-3350: union length of k-mers starting at offs[0..n) (ascending). */
int rs_total_hit_length(const int32_t *offs, int n, int k)
{
    int ret = 0;
    for (int i = 0; i < n;) {
        int j;
        for (j = i + 1; j < n; ++j)
            if (offs[j] > offs[j - 1] + k - 1)
                break;
        ret += offs[j - 1] - offs[i] + k;
        i = j;
    }
    return ret;
}

/* SeqSet.hpp:638-750 (DnaToAa) + 5029-5073 (HasMotif; translates `read` itself for either strand). */
static char aa_of(char a, char b, char c)
{
    static const char *tab = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV_Y_YSSSS_CWCLFLF";
    if (a == 'N' || b == 'N' || c == 'N')
        return '-';
    return tab[16 * nuc(a) + 4 * nuc(b) + nuc(c)];
}

int rs_has_motif(const char *read, int strand)
{
    if (strand == 0)
        return 0;
    int len = (int)strlen(read), ret = 0;
    char *aa = malloc(len + 1);
    for (int k = 0; k <= 2; ++k) {
        int i, j;
        for (i = k, j = 0; i + 2 < len; i += 3, ++j)
            aa[j] = aa_of(read[i], read[i + 1], read[i + 2]);
        for (i = 0; i + 2 < j; ++i)
            if (aa[i] == 'Y' && aa[i + 1] == 'Y' && aa[i + 2] == 'C') {
                ret |= 2;
                break;
            }
        for (i = 0; i + 3 < j; ++i)
            if ((aa[i] == 'F' || aa[i] == 'W') && aa[i + 1] == 'G' && aa[i + 3] == 'G') {
                ret |= 1;
                break;
            }
    }
    free(aa);
    return ret;
}

/* ---------------------------------------------------------------------------------------------------------
 * SeqSet::GetHitsFromRead (SeqSet.hpp:1341-1501) + SortHits (SeqSet.hpp:1306), restated over a freshly built
 * index of the given contigs (KmerIndex::BuildIndexFromRead per contig, idx = position in the array).
 * Includes the equal-to-previous rule with its stale prevKmerCode (the `continue`s skip the update), the
 * >= 100 postings skip rule (skipLimit = k/2, not for the first / last k-mer) and allowTotalSkip.
 * Output: int32[5] per hit {idx, offset, readOffset, strand, repeats}, ordered by (strand, idx, readOffset, offset).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct { uint64_t code; int32_t idx, off; } rs_posting;
typedef struct { int32_t idx, off, roff, strand, rep; } rs_hit;

static int cmp_posting(const void *a, const void *b)
{
    const rs_posting *x = a, *y = b;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return x->off < y->off ? -1 : (x->off > y->off);
}

static int cmp_hit(const void *a, const void *b)
{
    const rs_hit *x = a, *y = b;
    if (x->strand != y->strand) return x->strand < y->strand ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    if (x->roff != y->roff) return x->roff < y->roff ? -1 : 1;
    return x->off < y->off ? -1 : (x->off > y->off);
}

static void revcomp(char *rc, const char *s, int len) /* SeqSet.hpp:2616 */
{
    for (int i = 0; i < len; ++i) {
        char c = s[len - 1 - i];
        rc[i] = c == 'N' ? 'N' : "ACGT"[3 - nuc(c)];
    }
    rc[len] = 0;
}

int rs_get_hits(const char *const *contigs, int n_contigs, const char *read, int strand, int k, int allow_total_skip,
                int32_t *out, int cap)
{
    /* index */
    size_t total = 0;
    for (int c = 0; c < n_contigs; ++c)
        total += strlen(contigs[c]);
    rs_posting *post = malloc(sizeof(rs_posting) * (total + 1));
    int32_t *offs = malloc(sizeof(int32_t) * (total + 1));
    uint64_t *codes = malloc(sizeof(uint64_t) * (total + 1));
    size_t np = 0;
    for (int c = 0; c < n_contigs; ++c) {
        int len = (int)strlen(contigs[c]);
        int m = rs_index_offsets(contigs[c], len, k, 0, offs, codes);
        for (int i = 0; i < m; ++i) {
            post[np].code = codes[i];
            post[np].idx = c;
            post[np].off = offs[i];
            ++np;
        }
    }
    qsort(post, np, sizeof(rs_posting), cmp_posting);

    int len = (int)strlen(read);
    char *rc = malloc(len + 1);
    revcomp(rc, read, len);
    uint64_t *rcodes = malloc(sizeof(uint64_t) * (len + 1));
    unsigned char *rvalid = malloc(len + 1);
    rs_hit *hits = NULL;
    size_t nh = 0, hcap = 0;
    uint64_t prev = 0;              /* KmerCode prevKmerCode(k): code 0; carried over from the forward to the reverse pass */
    int skip_limit = k / 2;
    for (int pass = 0; pass < 2; ++pass) {
        if ((pass == 0 && strand == -1) || (pass == 1 && strand == 1))
            continue;
        const char *r = pass ? rc : read;
        rs_kmer_codes(r, len, k, rcodes, rvalid);
        int skip_cnt = 0;
        for (int i = k - 1; i < len; ++i) {
            if (i == k - 1 || rcodes[i] != prev) {
                /* KmerIndex::Search: postings of this code (none when the k-mer contains an N) */
                size_t lo = 0, hi = np;
                if (rvalid[i]) {
                    while (lo < hi) { size_t mid = (lo + hi) / 2; if (post[mid].code < rcodes[i]) lo = mid + 1; else hi = mid; }
                    hi = lo;
                    while (hi < np && post[hi].code == rcodes[i]) ++hi;
                } else
                    lo = hi = 0;
                int size = (int)(hi - lo);
                if (size >= 100 && i != k - 1 && i != len - 1 && skip_cnt < skip_limit) {
                    ++skip_cnt;
                    continue;           /* NB: prev is NOT updated */
                }
                if (size >= 100 && allow_total_skip)
                    continue;
                skip_cnt = 0;
                for (size_t j = lo; j < hi; ++j) {
                    if (nh == hcap) { hcap = hcap ? 2 * hcap : 1024; hits = realloc(hits, hcap * sizeof(rs_hit)); }
                    hits[nh].idx = post[j].idx;
                    hits[nh].off = post[j].off;
                    hits[nh].roff = i - k + 1;
                    hits[nh].strand = pass ? -1 : 1;
                    hits[nh].rep = size;
                    ++nh;
                }
            }
            prev = rcodes[i];
        }
    }
    qsort(hits, nh, sizeof(rs_hit), cmp_hit);
    for (size_t i = 0; i < nh && (int)i < cap; ++i) {
        out[5 * i] = hits[i].idx;
        out[5 * i + 1] = hits[i].off;
        out[5 * i + 2] = hits[i].roff;
        out[5 * i + 3] = hits[i].strand;
        out[5 * i + 4] = hits[i].rep;
    }
    free(hits); free(post); free(offs); free(codes); free(rc); free(rcodes); free(rvalid);
    return (int)nh;
}

/* ---------------------------------------------------------------------------------------------------------
 * SeqSet::GetOverlapsFromHits (SeqSet.hpp:763-1063) for a set of novel contigs (isRef false everywhere, so
 * adjustRadius = 0 and LongestIncreasingSubsequence is the identity on a run of one diagonal), restated over the
 * hit array produced by rs_get_hits (ordered by strand, idx, readOffset, offset = SortHits order).
 * filter == 1 runs the candidate-count pre-pass including its `i = j ; ++i` group skipping (SeqSet.hpp:784-810).
 * contig_lens is not needed: the novel-contig test `2 * hitLen < seqSpan` only uses the chain itself.
 * Output: int32[8] per chain {seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt, nHits}, in emission
 * order (group order, then diagonal ascending).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct { int a, b, c; } rs_triple;

static int cmp_triple(const void *x, const void *y)
{
    const rs_triple *p = x, *q = y;
    if (p->c != q->c) return p->c < q->c ? -1 : 1;
    if (p->b != q->b) return p->b < q->b ? -1 : 1;
    return p->a < q->a ? -1 : (p->a > q->a);
}

int rs_get_chains(const int32_t *hits, int n_hits, int k, int hit_len_required, int filter, int32_t *out, int cap)
{
#define H_IDX(i) hits[5 * (i)]
#define H_OFF(i) hits[5 * (i) + 1]
#define H_ROFF(i) hits[5 * (i) + 2]
#define H_STRAND(i) hits[5 * (i) + 3]
#define H_REP(i) hits[5 * (i) + 4]
    int novel_min[2] = {3, 3};
    int remove_only_repeats[2] = {0, 0};
    int i, j, x;
    if (filter == 1) {
        int possible[2] = {0, 0}, longest[2] = {0, 0};
        for (i = 0; i < n_hits; ++i) {
            int plus = (1 + H_STRAND(i)) / 2;
            for (j = i + 1; j < n_hits; ++j)
                if (H_STRAND(j) != H_STRAND(i) || H_IDX(j) != H_IDX(i))
                    break;
            if (j - i > novel_min[plus])
                ++possible[plus];
            if (j - i > longest[plus])
                longest[plus] = j - i;
            if (!remove_only_repeats[plus]) {
                int cnt = 0;
                for (x = i; x < j; ++x)
                    if (H_REP(x) <= 10000)
                        ++cnt;
                if (cnt >= novel_min[plus])
                    remove_only_repeats[plus] = 1;
            }
            i = j; /* and the for-loop adds one more */
        }
        for (i = 0; i <= 1; ++i) {
            if (possible[i] > 100000) novel_min[i] = (int)(longest[i] * 0.75);
            else if (possible[i] > 10000) novel_min[i] = longest[i] / 2;
            else if (possible[i] > 1000) novel_min[i] = longest[i] / 3;
            else if (possible[i] > 100) novel_min[i] = longest[i] / 4;
        }
    }
    int n_out = 0;
    rs_triple *t = malloc(sizeof(rs_triple) * (n_hits + 1));
    for (i = 0; i < n_hits;) {
        for (j = i + 1; j < n_hits; ++j)
            if (H_STRAND(j) != H_STRAND(i) || H_IDX(j) != H_IDX(i))
                break;
        int plus = (1 + H_STRAND(i)) / 2;
        int min_hit = novel_min[plus];
        if (j - i < min_hit) { i = j; continue; }
        if (remove_only_repeats[plus]) {
            int has_unique = 0;
            for (x = i; x < j; ++x)
                if (H_REP(x) <= 10000) { has_unique = 1; break; }
            if (!has_unique) { i = j; continue; }
        }
        for (x = i; x < j; ++x) {
            t[x - i].a = H_ROFF(x);
            t[x - i].b = H_OFF(x);
            t[x - i].c = H_ROFF(x) - H_OFF(x);
        }
        qsort(t, j - i, sizeof(rs_triple), cmp_triple);
        int s, e;
        for (s = 0; s < j - i;) {
            for (e = s + 1; e < j - i; ++e)
                if (t[e].c != t[e - 1].c) /* adjustRadius 0 */
                    break;
            if (e - s < min_hit || (e - s) * k < hit_len_required) { s = e; continue; }
            if (remove_only_repeats[plus]) {
                int has_unique = 0;
                for (x = s; x < e; ++x) /* the reference indexes hits[] with the run-local x here (SeqSet.hpp:934-941) */
                    if (H_REP(x) <= 10000) { has_unique = 1; break; }
                if (!has_unique) { s = e; continue; }
            }
            /* chain = the run; union length of its k-mers (identical on read and contig) */
            int hit_len = 0;
            for (x = s; x < e;) {
                int y;
                for (y = x + 1; y < e; ++y)
                    if (t[y].a > t[y - 1].a + k - 1)
                        break;
                hit_len += t[y - 1].a - t[x].a + k;
                x = y;
            }
            if (hit_len < hit_len_required) { s = e; continue; }
            int seq_start = t[s].b, seq_end = t[e - 1].b + k - 1;
            if (hit_len * 2 < seq_end - seq_start + 1) { s = e; continue; }
            if (n_out < cap) {
                int32_t *o = out + 8 * n_out;
                o[0] = H_IDX(i);
                o[1] = t[s].a;
                o[2] = t[e - 1].a + k - 1;
                o[3] = seq_start;
                o[4] = seq_end;
                o[5] = H_STRAND(i);
                o[6] = 2 * hit_len;
                o[7] = e - s;
            }
            ++n_out;
            s = e;
        }
        i = j;
    }
    free(t);
    return n_out;
}
