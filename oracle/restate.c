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
