/* TEST INFRASTRUCTURE ONLY -- see oracle.h.
 *
 * Restatement of the two steps that produce the read set the graph creators see:
 *
 *  oracle_read_input   InputReader::readInput (src/IO/InputReader.cpp:44-139) in its --threads=1 order:
 *                      readOneRead1 (:142-180: plain tokens / 2-line FASTA / 4-line FASTQ records, reading stops at the
 *                      first record whose sequence line is empty), the space stripping and the 3+3 end trimming of
 *                      readParallelJob (:286-303), the character check (:316-334, the reference exits there; here
 *                      ORACLE_E_BADCHAR), reads with an N dropped (:341-347), reads whose minimal period is <= 20 dropped
 *                      (:343-353, MyUtils::MinPeriod, MyUtils.h:160-170), the reverse complement twin (:362-377,
 *                      getComplimentaryString :23-33), Read::createSequence (Read.cpp:40-68), the interleaving of the
 *                      two mate files (:54-72), the swap that puts the reverse complement first (:78-80) and ids = index
 *                      (:83-85).  With several threads and a malformed file (an empty sequence line before the end) the
 *                      reference reads a thread-count-dependent subset; --threads=1 is the contract.
 *
 *  oracle_remap        the compaction and renumbering of main.cpp:150-232 (after Global::removeRead of the reads
 *                      ReadPreprocess marked, main.cpp:133-140), literally, with Global::pairedReadOffset.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* std::getline on the buffer: line = [*pos, end of line), *pos moves behind the '\n'; at the end of the buffer the line
 * is empty (and stays empty), as with a failed stream. */
static void get_line(const uint8_t *t, uint64_t n, uint64_t *pos, uint64_t *b, uint64_t *e) {
    uint64_t p = *pos;
    *b = p;
    while (p < n && t[p] != '\n') p++;
    *e = p;
    *pos = p < n ? p + 1 : n;
}

static int is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

/* readOneRead1, InputReader.cpp:142-180 (PFASTA with paired reads reads like FASTA) */
static void read_one(const uint8_t *t, uint64_t n, uint64_t *pos, int file_type, uint64_t *b, uint64_t *e) {
    uint64_t x, y;
    if (file_type == ORACLE_INPUT_PLAIN) { /* str >> s */
        uint64_t p = *pos;
        while (p < n && is_space(t[p])) p++;
        *b = p;
        while (p < n && !is_space(t[p])) p++;
        *e = p;
        *pos = p;
        return;
    }
    get_line(t, n, pos, &x, &y);
    get_line(t, n, pos, b, e);
    if (file_type == ORACLE_INPUT_FASTQ) {
        get_line(t, n, pos, &x, &y);
        get_line(t, n, pos, &x, &y);
    }
}

/* MyUtils::MinPeriod, MyUtils.h:160-170 */
static int min_period(const uint8_t *s, int len) {
    int *pre = (int *) malloc((size_t) (len + 2) * sizeof(int));
    int k = 0;
    pre[0] = 0;
    pre[1] = 0;
    for (int q = 1; q < len; q++) {
        while (k > 0 && s[k] != s[q]) k = pre[k];
        if (s[k] == s[q]) k++;
        pre[q + 1] = k;
    }
    const int r = len - pre[len];
    free(pre);
    return r;
}

typedef struct {
    uint8_t *seq; /* trimmed sequence, NULL = removed read */
    uint32_t len;
} rec_t;

/* Read::createSequence, Read.cpp:40-68 */
static void pack(const uint8_t *s, uint32_t len, uint32_t *w) {
    for (uint32_t i = 0; i < len; i++) {
        uint32_t c = 0;
        switch (s[i]) {
            case 'C': c = 1; break;
            case 'G': c = 2; break;
            case 'T': c = 3; break;
            default: c = 0;
        }
        w[i >> 4] |= c << ((i & 15u) * 2u);
    }
}

static int read_file(const uint8_t *t, uint64_t n, const oracle_input_params *p, rec_t **out, uint32_t *n_out,
                     uint64_t *n_with_n, uint64_t *n_str) {
    uint64_t pos = 0, b, e;
    uint32_t cap = 1024, cnt = 0;
    rec_t *v = (rec_t *) malloc(cap * sizeof(rec_t));
    const int thr = p->str_threshold > 0 ? p->str_threshold : 20;
    int rc = 0;
    while (1) {
        read_one(t, n, &pos, p->file_type, &b, &e);
        if (b == e) break; /* s == "" */
        /* :286-291 strip leading spaces, cut at the next space */
        while (b < e && t[b] == ' ') b++;
        uint64_t q = b;
        while (q < e && t[q] != ' ') q++;
        e = q;
        /* :298-303 */
        const uint64_t size = e - b;
        if (size >= (uint64_t) (p->trim_left + p->trim_right + 10)) {
            b += (uint64_t) p->trim_left;
            e -= (uint64_t) p->trim_right;
        }
        const uint32_t len = (uint32_t) (e - b);
        uint8_t *s = (uint8_t *) malloc(len + 1u);
        memcpy(s, t + b, len);
        s[len] = 0;
        int has_n = 0;
        for (uint32_t i = 0; i < len; i++) { /* :316-334 */
            const uint8_t c = s[i];
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N' && c != 'U') rc = ORACLE_E_BADCHAR;
            if (c == 'N') has_n = 1;
            else if (p->rna && c == 'U') s[i] = 'T';
        }
        if (rc) {
            free(s);
            break;
        }
        if (cnt + 2 > cap) {
            cap *= 2;
            v = (rec_t *) realloc(v, cap * sizeof(rec_t));
        }
        if (has_n) { /* :341-347, REMOVE_READS_WITH_N = 1 */
            free(s);
            v[cnt].seq = NULL, v[cnt].len = 0;
            v[cnt + 1] = v[cnt];
            (*n_with_n)++;
        } else if (len == 0 || min_period(s, (int) len) <= thr) { /* :348-353; an all-space line gives the empty string, whose
                                                                    MinPeriod is undefined in the reference: dropped here */
            free(s);
            v[cnt].seq = NULL, v[cnt].len = 0;
            v[cnt + 1] = v[cnt];
            (*n_str)++;
        } else {
            uint8_t *rcs = (uint8_t *) malloc(len + 1u);
            for (uint32_t i = 0; i < len; i++) { /* :363-365, getComplimentaryString :23-33 */
                const uint8_t c = s[len - 1u - i];
                rcs[i] = c == 'A' ? 'T' : (c == 'C' ? 'G' : (c == 'G' ? 'C' : (c == 'T' ? 'A' : c)));
            }
            rcs[len] = 0;
            v[cnt].seq = s, v[cnt].len = len;         /* forward */
            v[cnt + 1].seq = rcs, v[cnt + 1].len = len; /* reverse complement (its MinPeriod equals the forward one's) */
        }
        cnt += 2;
    }
    if (rc) {
        for (uint32_t i = 0; i < cnt; i++) free(v[i].seq);
        free(v);
        return rc;
    }
    *out = v;
    *n_out = cnt;
    return 0;
}

int oracle_read_input(const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2, const oracle_input_params *p,
                      uint32_t *n_reads, uint32_t **len_nt, uint64_t **word_off, uint32_t **words, uint64_t *n_with_n,
                      uint64_t *n_str) {
    rec_t *a = NULL, *b = NULL;
    uint32_t na = 0, nb = 0;
    uint64_t cn = 0, cs = 0;
    int rc = read_file(text1, n1, p, &a, &na, &cn, &cs);
    if (rc) return rc;
    if (text2) {
        rc = read_file(text2, n2, p, &b, &nb, &cn, &cs);
        if (rc == 0 && na != nb) rc = ORACLE_E_PAIRING; /* InputReader.cpp:58-63 indexes past the shorter file */
        if (rc) {
            for (uint32_t i = 0; i < na; i++) free(a[i].seq);
            free(a);
            if (b) {
                for (uint32_t i = 0; i < nb; i++) free(b[i].seq);
                free(b);
            }
            return rc;
        }
    }
    const uint32_t n = na + nb;
    rec_t *all = (rec_t *) malloc((size_t) (n ? n : 1) * sizeof(rec_t));
    if (text2) { /* :58-63 */
        for (uint32_t i = 0; 4ull * i < n; i++) {
            all[4 * i] = a[2 * i];
            all[4 * i + 1] = a[2 * i + 1];
            all[4 * i + 2] = b[2 * i];
            all[4 * i + 3] = b[2 * i + 1];
        }
    } else {
        memcpy(all, a, (size_t) n * sizeof(rec_t));
    }
    for (uint32_t i = 0; i + 1 < n; i += 2) { /* :78-80 */
        rec_t x = all[i];
        all[i] = all[i + 1];
        all[i + 1] = x;
    }
    *n_reads = n;
    *len_nt = (uint32_t *) calloc(n ? n : 1, sizeof(uint32_t));
    *word_off = (uint64_t *) calloc((size_t) n + 1, sizeof(uint64_t));
    for (uint32_t i = 0; i < n; i++) {
        (*len_nt)[i] = all[i].seq ? all[i].len : 0;
        (*word_off)[i + 1] = (*word_off)[i] + ((*len_nt)[i] + 15u) / 16u;
    }
    *words = (uint32_t *) calloc((size_t) ((*word_off)[n] ? (*word_off)[n] : 1), sizeof(uint32_t));
    for (uint32_t i = 0; i < n; i++)
        if (all[i].seq) pack(all[i].seq, all[i].len, *words + (*word_off)[i]);
    for (uint32_t i = 0; i < n; i++) free(all[i].seq);
    free(all);
    free(a);
    free(b);
    if (n_with_n) *n_with_n = cn;
    if (n_str) *n_str = cs;
    return 0;
}

/* main.cpp:133-140 (Global::removeRead of the marked reads) and :150-232, literally: READS[] holds the old id of the read
 * in each slot or -1 for nullptr.  old_id / paired_offset need room for n entries. */
int oracle_remap(const uint32_t *len_nt, const uint8_t *mask, uint32_t n, uint32_t *old_id, uint8_t *paired_offset,
                 uint32_t *n_out) {
    if (n & 1u) return ORACLE_E_PAIRING;
    int64_t *reads = (int64_t *) malloc((size_t) (n ? n : 1) * sizeof(int64_t));
    for (uint32_t i = 0; i < n; i++) reads[i] = (len_nt[i] != 0 && !(mask && mask[i])) ? (int64_t) i : -1;
    uint32_t back = 0, np = 0;
    int rc = 0;
    for (uint32_t i = 0; i < n; i += 2) {
        if (reads[i] < 0) continue;
        if (reads[i + 1] < 0) { /* assert, main.cpp:173-174 */
            rc = ORACLE_E_PAIRING;
            break;
        }
        if ((i & 3u) == 0) {
            if (i + 2 < n && reads[i + 2] >= 0) {
                paired_offset[np++] = 1, paired_offset[np++] = 1;
                reads[back] = reads[i], reads[back + 1] = reads[i + 1], back += 2;
                paired_offset[np++] = 2, paired_offset[np++] = 2;
                reads[back] = reads[i + 2], reads[back + 1] = reads[i + 3], back += 2;
            } else {
                paired_offset[np++] = 0, paired_offset[np++] = 0;
                reads[back] = reads[i], reads[back + 1] = reads[i + 1], back += 2;
            }
        } else if (reads[i - 2] < 0) {
            paired_offset[np++] = 0, paired_offset[np++] = 0;
            reads[back] = reads[i], reads[back + 1] = reads[i + 1], back += 2;
        }
    }
    if (rc == 0) {
        for (uint32_t k = 0; k < back; k++) {
            if (reads[k] < 0) { /* assert, main.cpp:222 */
                rc = ORACLE_E_PAIRING;
                break;
            }
            old_id[k] = (uint32_t) reads[k];
        }
        *n_out = back;
    }
    free(reads);
    return rc;
}
