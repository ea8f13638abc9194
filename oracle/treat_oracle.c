/*
 * oracle/treat_oracle.c -- CPU restatement of ubccr/treat's alignment path.
 * TEST INFRASTRUCTURE ONLY (see treat_oracle.h).  Plain C11, no dependencies.
 *
 * Sources restated (reference repo paths): const.go:20-23, fragment.go:55,75-147,
 * template.go:33-34,51-196, align.go:88-279,316-520, cmd/treat/align.go:39-122, and the
 * un-vendored modules
 *   github.com/aebruno/nwalgo  v0.0.0-20160817130739-4a232086e3ad (go.mod:7)
 *   github.com/willf/bitset    v1.1.10                            (go.mod:16)
 *   github.com/aebruno/gofasta v0.0.0-20150407023551-e776ef625791 (go.mod:6)
 * whose published behaviour is restated here and anchored on the reference's call
 * sites and golden vectors (tests/golden/).
 *
 * ASCII assumption: Go's strings.ToUpper / unicode.ToUpper are applied to runes; every
 * FASTA the reference ships is ASCII, and bytes >= 0x80 are passed through unchanged here.
 */
#define _GNU_SOURCE
#include "treat_oracle.h"

#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int g_tie_order = ORC_TIE_ULD;
void orc_set_tie_order(int order) { g_tie_order = order; }
int orc_get_tie_order(void) { return g_tie_order; }
void orc_free(void *p) { free(p); }

static void *xmalloc(size_t n)
{
    void *p = malloc(n ? n : 1);
    if (!p) {
        fprintf(stderr, "oracle: out of memory\n");
        abort();
    }
    return p;
}
static void *xcalloc(size_t n, size_t s)
{
    void *p = calloc(n ? n : 1, s ? s : 1);
    if (!p) {
        fprintf(stderr, "oracle: out of memory\n");
        abort();
    }
    return p;
}
static char *xstrndup(const char *s, size_t n)
{
    char *p = xmalloc(n + 1);
    memcpy(p, s, n);
    p[n] = 0;
    return p;
}
static char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
/* Go's unicode.IsSpace restricted to ASCII */
static int go_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

/* growable byte buffer (bytes.Buffer) */
typedef struct {
    char *p;
    size_t len, cap;
} buf_t;
static void buf_reserve(buf_t *b, size_t extra)
{
    if (b->len + extra + 1 > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 64;
        while (nc < b->len + extra + 1) nc *= 2;
        b->p = realloc(b->p, nc);
        if (!b->p) abort();
        b->cap = nc;
    }
}
static void buf_putc(buf_t *b, char c)
{
    buf_reserve(b, 1);
    b->p[b->len++] = c;
    b->p[b->len] = 0;
}
static void buf_rep(buf_t *b, char c, size_t n)
{
    buf_reserve(b, n);
    memset(b->p + b->len, c, n);
    b->len += n;
    b->p[b->len] = 0;
}
static void buf_put(buf_t *b, const char *s, size_t n)
{
    buf_reserve(b, n);
    memcpy(b->p + b->len, s, n);
    b->len += n;
    b->p[b->len] = 0;
}
static void buf_puts(buf_t *b, const char *s) { buf_put(b, s, strlen(s)); }

/* ---------------------------------------------------------------- fragment.go */

/* fragment.go:75-89 with fastxPattern `.+[_\-](\d+)$` (fragment.go:55) */
uint32_t orc_parse_merge_count(const char *id, size_t len)
{
    size_t b = 0, e = len;
    while (b < e && go_space(id[b])) b++; /* strings.TrimSpace */
    while (e > b && go_space(id[e - 1])) e--;
    size_t t = b;
    while (t < e && id[t] != ' ') t++; /* SplitN(..., " ", 2)[0] */
    /* greedy `.+` then [_-] then digits to the end: the separator is the last '_'/'-' */
    size_t sep = t;
    for (size_t i = t; i > b; i--) {
        if (id[i - 1] == '_' || id[i - 1] == '-') {
            sep = i - 1;
            break;
        }
    }
    if (sep == t) return 1;          /* no separator */
    if (sep == b) return 1;          /* `.+` needs one char before it */
    if (sep + 1 >= t) return 1;      /* `\d+` needs one digit */
    for (size_t i = b; i < t; i++)
        if (id[i] == '\n') return 1; /* `.` does not match newline (cannot occur after TrimSpace/Split on FASTA ids) */
    unsigned long long v = 0;
    for (size_t i = sep + 1; i < t; i++) {
        if (id[i] < '0' || id[i] > '9') return 1;
        unsigned d = (unsigned)(id[i] - '0');
        if (v > (0x7fffffffffffffffULL - d) / 10) return 1; /* strconv.Atoi range error -> 1 */
        v = v * 10 + d;
    }
    return (uint32_t)v; /* uint32(count) truncation */
}

/* fragment.go:91-121 */
orc_fragment *orc_new_fragment(const char *name, const char *seq, size_t seqlen, int orientation,
                               char base)
{
    orc_fragment *f = xcalloc(1, sizeof *f);
    base = up(base);
    char *s = xmalloc(seqlen + 1);
    for (size_t i = 0; i < seqlen; i++) {
        size_t src = (orientation == ORC_REVERSE) ? seqlen - 1 - i : i; /* plain reversal, no complement */
        s[i] = up(seq[src]);
    }
    size_t n = 0;
    for (size_t i = 0; i < seqlen; i++)
        if (s[i] != base) n++;
    f->bases = xmalloc(n + 1);
    f->edit_site = xcalloc(n + 1, sizeof(uint32_t));
    uint32_t run = 0;
    size_t idx = 0;
    for (size_t i = 0; i < seqlen; i++) {
        if (s[i] != base) {
            f->edit_site[idx] = run;
            run = 0;
            f->bases[idx++] = s[i];
        } else {
            run++;
        }
    }
    f->edit_site[idx] = run;
    f->bases[n] = 0;
    f->n = n;
    f->edit_base = base;
    f->name = xstrndup(name, strlen(name));
    f->read_count = orc_parse_merge_count(name, strlen(name));
    f->norm = 0.0;
    free(s);
    return f;
}

/* fragment.go:136-147 */
char *orc_fragment_string(const orc_fragment *f)
{
    buf_t b = {0};
    buf_reserve(&b, 0);
    for (size_t i = 0; i <= f->n; i++) {
        buf_rep(&b, f->edit_base, f->edit_site[i]);
        if (i < f->n) buf_putc(&b, f->bases[i]);
    }
    return b.p;
}

void orc_fragment_free(orc_fragment *f)
{
    if (!f) return;
    free(f->name);
    free(f->bases);
    free(f->edit_site);
    free(f);
}

/* ---------------------------------------------------------------- gofasta */
/*
 * gofasta.SimpleParser: records start at '>', Id = the whole header line after '>'
 * (template.go:68-75 and align_test.go:85 regex attributes out of rec.Id), sequence lines
 * are concatenated (template_test.go:27-34 expects 5 records from the wrapped
 * examples/templates.fa).  ASSUMPTIONS (source not available): Id is TrimSpace'd; a
 * trailing '\r' on a line is dropped; a last line without '\n' is kept.
 */
orc_fasta *orc_fasta_read(const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *data = xmalloc((size_t)sz + 1);
    if (fread(data, 1, (size_t)sz, fp) != (size_t)sz) {
        fclose(fp);
        free(data);
        return NULL;
    }
    fclose(fp);
    data[sz] = 0;
    orc_fasta *fa = xcalloc(1, sizeof *fa);
    size_t cap = 0;
    buf_t seq = {0};
    int have = 0;
    size_t pos = 0;
    /* skip everything before the first '>' */
    while (pos < (size_t)sz && data[pos] != '>') pos++;
    while (pos < (size_t)sz) {
        size_t eol = pos;
        while (eol < (size_t)sz && data[eol] != '\n') eol++;
        size_t le = eol;
        if (le > pos && data[le - 1] == '\r') le--;
        if (data[pos] == '>') {
            if (have) {
                fa->seqs[fa->count - 1] = xstrndup(seq.p ? seq.p : "", seq.len);
                seq.len = 0;
            }
            if (fa->count == cap) {
                cap = cap ? cap * 2 : 16;
                fa->ids = realloc(fa->ids, cap * sizeof(char *));
                fa->seqs = realloc(fa->seqs, cap * sizeof(char *));
            }
            size_t b = pos + 1, e = le;
            while (b < e && go_space(data[b])) b++;
            while (e > b && go_space(data[e - 1])) e--;
            fa->ids[fa->count] = xstrndup(data + b, e - b);
            fa->seqs[fa->count] = NULL;
            fa->count++;
            have = 1;
        } else if (have) {
            buf_put(&seq, data + pos, le - pos);
        }
        pos = eol + 1;
    }
    if (have) fa->seqs[fa->count - 1] = xstrndup(seq.p ? seq.p : "", seq.len);
    free(seq.p);
    free(data);
    return fa;
}

void orc_fasta_free(orc_fasta *f)
{
    if (!f) return;
    for (size_t i = 0; i < f->count; i++) {
        free(f->ids[i]);
        free(f->seqs[i]);
    }
    free(f->ids);
    free(f->seqs);
    free(f);
}

/* ---------------------------------------------------------------- template.go */

/* template.go:96-151 */
orc_template *orc_new_template(const orc_fragment *full, const orc_fragment *pre,
                               const orc_fragment *const *alt, size_t n_alt, const int *alt_start,
                               const int *alt_end, size_t n_region, char *err, size_t errlen)
{
    if (full->n != pre->n || memcmp(full->bases, pre->bases, full->n) != 0 ||
        full->edit_base != pre->edit_base) {
        snprintf(err, errlen, "Invalid template sequences. Full and Pre templates must have the same non-edit bases");
        return NULL;
    }
    for (size_t i = 0; i < n_alt; i++) {
        if (full->n != alt[i]->n || memcmp(full->bases, alt[i]->bases, full->n) != 0 ||
            full->edit_base != alt[i]->edit_base) {
            snprintf(err, errlen, "Invalid alt template sequence. All templates must have the same non-edit bases");
            return NULL;
        }
    }
    if (n_alt != n_region) {
        snprintf(err, errlen, "Invalid alt templates. Please specify the alt regions");
        return NULL;
    }
    orc_template *t = xcalloc(1, sizeof *t);
    size_t len = full->n + 1;
    t->m = full->n;
    t->bases = xstrndup(full->bases, full->n);
    t->edit_base = full->edit_base;
    t->K = n_alt + 2;
    t->edit_site = xcalloc(t->K, sizeof(uint32_t *));
    for (size_t k = 0; k < t->K; k++) {
        const orc_fragment *src = k == 0 ? full : k == 1 ? pre : alt[k - 2];
        t->edit_site[k] = xmalloc(len * sizeof(uint32_t));
        memcpy(t->edit_site[k], src->edit_site, len * sizeof(uint32_t));
    }
    t->n_alt = n_alt;
    t->alt_start = xcalloc(n_alt, sizeof(int));
    t->alt_end = xcalloc(n_alt, sizeof(int));
    for (size_t i = 0; i < n_alt; i++) {
        t->alt_start[i] = alt_start[i];
        t->alt_end[i] = alt_end[i];
    }
    /* BaseIndex, template.go:122-135 */
    t->base_index = xcalloc(len, sizeof(uint32_t));
    uint32_t index = 0;
    for (size_t i = 0; i < len; i++) {
        uint32_t mx = t->edit_site[0][i];
        if (t->edit_site[1][i] > mx) mx = t->edit_site[1][i];
        index += mx;
        if (i > 0 && i != len - 1) index++;
        t->base_index[i] = index;
    }
    /* template EditStop, template.go:139-148 */
    t->edit_stop = (int)len - 1;
    for (int j = t->edit_stop; j >= 0; j--) {
        if (t->edit_site[0][j] != t->edit_site[1][j]) {
            t->edit_stop = ((int)len - 1) - j;
            break;
        }
    }
    t->edit_stop--;
    return t;
}

/* first regex match of `<key>(\d+)`; returns -1 if none; Atoi range error -> 0 (template.go:68-82) */
static int find_alt_attr(const char *id, const char *key)
{
    size_t kl = strlen(key);
    const char *p = id;
    while ((p = strstr(p, key)) != NULL) {
        const char *d = p + kl;
        if (*d >= '0' && *d <= '9') {
            unsigned long long v = 0;
            int overflow = 0;
            while (*d >= '0' && *d <= '9') {
                unsigned dd = (unsigned)(*d - '0');
                if (v > (0x7fffffffffffffffULL - dd) / 10) overflow = 1;
                if (!overflow) v = v * 10 + dd;
                d++;
            }
            return overflow ? 0 : (int)v;
        }
        p++;
    }
    return -1;
}

/* template.go:51-94 */
orc_template *orc_template_from_fasta(const char *path, int orientation, char base, char *err,
                                      size_t errlen)
{
    orc_fasta *fa = orc_fasta_read(path);
    if (!fa) {
        snprintf(err, errlen, "Invalid FASTA file: open %s", path);
        return NULL;
    }
    size_t nf = fa->count;
    orc_fragment **t = xcalloc(nf, sizeof *t);
    int *as = xcalloc(nf, sizeof(int)), *ae = xcalloc(nf, sizeof(int));
    size_t nreg = 0;
    for (size_t i = 0; i < nf; i++) {
        t[i] = orc_new_fragment(fa->ids[i], fa->seqs[i], strlen(fa->seqs[i]), orientation, base);
        if (i + 1 > 2) {
            int start = find_alt_attr(fa->ids[i], "alt_start=");
            int end = find_alt_attr(fa->ids[i], "alt_stop=");
            if (start > -1 && end > -1) {
                as[nreg] = start;
                ae[nreg] = end;
                nreg++;
            }
        }
    }
    orc_template *tm = NULL;
    if (nf < 2) {
        snprintf(err, errlen, "Must provide at least 2 templates. Full and Pre edited");
    } else {
        tm = orc_new_template(t[0], t[1], (const orc_fragment *const *)(t + 2), nf - 2, as, ae, nreg,
                              err, errlen);
    }
    for (size_t i = 0; i < nf; i++) orc_fragment_free(t[i]);
    free(t);
    free(as);
    free(ae);
    orc_fasta_free(fa);
    return tm;
}

/* template.go:153-161 (mutates alt regions in place; not idempotent) */
void orc_template_set_offset(orc_template *t, int offset)
{
    t->edit_offset = (uint32_t)offset;
    t->edit_stop += offset;
    for (size_t i = 0; i < t->n_alt; i++) {
        t->alt_start[i] -= offset;
        t->alt_end[i] -= offset;
    }
}

/* template.go:186-194 */
uint32_t orc_template_max(const orc_template *t, size_t i)
{
    uint32_t mx = 0;
    for (size_t k = 0; k < t->K; k++)
        if (t->edit_site[k][i] > mx) mx = t->edit_site[k][i];
    return mx;
}

void orc_template_free(orc_template *t)
{
    if (!t) return;
    free(t->bases);
    for (size_t k = 0; k < t->K; k++) free(t->edit_site[k]);
    free(t->edit_site);
    free(t->base_index);
    free(t->alt_start);
    free(t->alt_end);
    free(t);
}

/* ---------------------------------------------------------------- nwalgo */
/*
 * nwalgo.Align(a, b, match, mismatch, gap) -- restated from the published algorithm of
 * github.com/aebruno/nwalgo@4a232086e3ad (not in the reference tree; call sites
 * align.go:129,338,393).  Full (m+1)(n+1) score matrix of Go `int` and a byte pointer
 * matrix; boundary pointers Up on column 0 and Left on row 0; per cell
 *   max = diag; if hgap(=f[i-1][j]+gap) > max ...; if vgap(=f[i][j-1]+gap) > max ...
 *   pointer = Up if max == hgap, else Left if max == vgap, else Diagonal
 * ("ULD"; g_tie_order selects the other five orders for the SURVEY 8c probe).
 * Up consumes a[i-1] against '-', Left consumes b[j-1] against '-'.
 */
enum { P_UP = 1, P_LEFT = 2, P_NW = 3 };

static uint8_t pick_pointer(int64_t mx, int64_t diag, int64_t hgap, int64_t vgap)
{
    int isU = mx == hgap, isL = mx == vgap, isD = mx == diag;
    switch (g_tie_order) {
    default:
    case ORC_TIE_ULD: return isU ? P_UP : isL ? P_LEFT : P_NW;
    case ORC_TIE_UDL: return isU ? P_UP : isD ? P_NW : P_LEFT;
    case ORC_TIE_LUD: return isL ? P_LEFT : isU ? P_UP : P_NW;
    case ORC_TIE_LDU: return isL ? P_LEFT : isD ? P_NW : P_UP;
    case ORC_TIE_DUL: return isD ? P_NW : isU ? P_UP : P_LEFT;
    case ORC_TIE_DLU: return isD ? P_NW : isL ? P_LEFT : P_UP;
    }
}

/* returns ops in REVERSE order (traceback order); ops: 0 diag, 1 up, 2 left */
static size_t nw_core(const char *a, size_t m, const char *b, size_t n, int match, int mismatch,
                      int gap, uint8_t *ops_rev, int *score)
{
    size_t al = m + 1, bl = n + 1;
    int64_t *f = xcalloc(al * bl, sizeof(int64_t));
    uint8_t *ptr = xcalloc(al * bl, 1);
    for (size_t i = 1; i < al; i++) {
        f[i * bl] = (int64_t)gap * (int64_t)i;
        ptr[i * bl] = P_UP;
    }
    for (size_t j = 1; j < bl; j++) {
        f[j] = (int64_t)gap * (int64_t)j;
        ptr[j] = P_LEFT;
    }
    for (size_t i = 1; i < al; i++) {
        for (size_t j = 1; j < bl; j++) {
            int64_t mm = (a[i - 1] == b[j - 1]) ? match : mismatch;
            int64_t diag = f[(i - 1) * bl + (j - 1)] + mm;
            int64_t hgap = f[(i - 1) * bl + j] + gap;
            int64_t vgap = f[i * bl + (j - 1)] + gap;
            int64_t mx = diag;
            if (hgap > mx) mx = hgap;
            if (vgap > mx) mx = vgap;
            ptr[i * bl + j] = pick_pointer(mx, diag, hgap, vgap);
            f[i * bl + j] = mx;
        }
    }
    size_t len = 0;
    size_t i = m, j = n;
    while (i > 0 || j > 0) {
        uint8_t p = ptr[i * bl + j];
        if (p == P_NW) {
            ops_rev[len++] = 0;
            i--;
            j--;
        } else if (p == P_UP) {
            ops_rev[len++] = 1;
            i--;
        } else {
            ops_rev[len++] = 2;
            j--;
        }
    }
    if (score) *score = (int)f[al * bl - 1];
    free(f);
    free(ptr);
    return len;
}

size_t orc_nw_align(const char *a, size_t m, const char *b, size_t n, int match, int mismatch,
                    int gap, char **aln1, char **aln2, int *score)
{
    uint8_t *ops = xmalloc(m + n + 1);
    size_t len = nw_core(a, m, b, n, match, mismatch, gap, ops, score);
    char *x = xmalloc(len + 1), *y = xmalloc(len + 1);
    size_t i = 0, j = 0;
    for (size_t c = 0; c < len; c++) {
        uint8_t op = ops[len - 1 - c];
        x[c] = (op == 2) ? '-' : a[i++];
        y[c] = (op == 1) ? '-' : b[j++];
    }
    x[len] = y[len] = 0;
    free(ops);
    *aln1 = x;
    *aln2 = y;
    return len;
}

/* ---------------------------------------------------------------- bitset (willf/bitset v1.1.10 semantics) */
typedef struct {
    size_t len;
    uint8_t *bit; /* one byte per bit: clarity over speed */
} bs_t;
static bs_t bs_new(size_t len)
{
    bs_t b = {len, xcalloc(len, 1)};
    return b;
}
static void bs_set(bs_t *b, size_t i) { if (i < b->len) b->bit[i] = 1; }
static int bs_test(const bs_t *b, unsigned long long i) { return i < b->len ? b->bit[i] : 0; } /* Test(i >= Len) == false */
static int bs_all(const bs_t *b)
{
    for (size_t i = 0; i < b->len; i++)
        if (!b->bit[i]) return 0;
    return 1;
}
static int bs_none(const bs_t *b)
{
    for (size_t i = 0; i < b->len; i++)
        if (b->bit[i]) return 0;
    return 1;
}

/* align.go:95-108: SymmetricDifference with the all-ones set = the unset positions */
static int find_jes(const bs_t *b)
{
    if (bs_all(b)) return -1;
    if (bs_none(b)) return (int)b->len - 1;
    size_t mx = 0;
    for (size_t i = 0; i < b->len; i++)
        if (!b->bit[i]) mx = i;
    return (int)mx;
}
/* align.go:110-120 */
static int find_jss(const bs_t *b)
{
    if (bs_all(b)) return (int)b->len - 1;
    if (bs_none(b)) return -1;
    for (size_t i = 0; i < b->len; i++)
        if (!b->bit[i]) return (int)i;
    return 0; /* unreachable */
}

/* ---------------------------------------------------------------- align.go */

/* align.go:122-181; ops_fwd in alignment order */
static bs_t *compute_t(orc_alignment *a, const orc_fragment *frag, const orc_template *tmpl,
                       int exclude_snps, const uint8_t *ops_fwd, size_t alen)
{
    size_t size = tmpl->m + 1;
    bs_t *T = xcalloc(tmpl->K, sizeof *T);
    for (size_t k = 0; k < tmpl->K; k++) T[k] = bs_new(size);
    size_t fi = 0, ti = 0;
    for (size_t ai = 0; ai < alen; ai++) {
        if (ops_fwd[ai] == 2) { /* aln1[ai] == '-' : insertion */
            fi++;
            a->has_mutation = 1;
            a->indel = 1;
            continue;
        }
        uint32_t count = 0;
        if (ops_fwd[ai] != 1) {
            count = frag->edit_site[fi];
            if (frag->bases[fi] != tmpl->bases[ti]) {
                a->mismatches++; /* uint8 wraps */
                if (exclude_snps || a->mismatches > 2) a->has_mutation = 1;
            }
        } else { /* deletion */
            a->has_mutation = 1;
            a->indel = 1;
        }
        for (size_t k = 0; k < tmpl->K; k++)
            if (tmpl->edit_site[k][ti] == count) bs_set(&T[k], (size - 1) - ti);
        if (ops_fwd[ai] != 1) fi++;
        ti++;
    }
    for (size_t k = 0; k < tmpl->K; k++)
        if (tmpl->edit_site[k][ti] == frag->edit_site[fi]) bs_set(&T[k], (size - 1) - ti);
    return T;
}

/* align.go:183-235 */
static void compute_alt_editing(orc_alignment *a, const orc_template *tmpl, const bs_t *T)
{
    int shift = a->junc_start;
    int alt = 0, has_alt = 0;
    int len = (int)tmpl->m + 1;
    for (size_t i = 2; i < tmpl->K; i++) {
        if (bs_test(&T[i], (unsigned long long)(long long)shift)) { /* uint(shift): -1 wraps, Test false */
            alt = (int)i - 2;
            has_alt = 1;
            break;
        }
    }
    if (!has_alt) return;
    if (shift != tmpl->alt_start[alt]) return;
    int full_match = 1;
    for (int x = shift; x < len; x++) {
        if (bs_test(&T[alt + 2], (unsigned long long)x)) continue;
        full_match = 0;
        shift = x;
        break;
    }
    if (full_match || shift >= tmpl->alt_end[alt]) {
        a->alt_editing = (uint8_t)(alt + 1);
        if (full_match) {
            a->junc_start = len - 1;
        } else {
            for (int j = shift; j < len; j++) {
                if (!bs_test(&T[0], (unsigned long long)j)) {
                    a->junc_start = j;
                    break;
                }
            }
        }
    }
}

static orc_alignment *new_alignment_ops(const orc_fragment *frag, const orc_template *tmpl,
                                        int exclude_snps, uint8_t *ops_fwd_out, size_t *alen_out)
{
    orc_alignment *a = xcalloc(1, sizeof *a);
    uint8_t *rev = xmalloc(tmpl->m + frag->n + 1);
    size_t alen = nw_core(tmpl->bases, tmpl->m, frag->bases, frag->n, 1, -1, -1, rev, &a->score);
    uint8_t *fwd = xmalloc(alen + 1);
    for (size_t c = 0; c < alen; c++) fwd[c] = rev[alen - 1 - c];
    free(rev);

    bs_t *T = compute_t(a, frag, tmpl, exclude_snps, fwd, alen);
    int len = (int)tmpl->m + 1;
    buf_t js = {0};
    buf_reserve(&js, 0);

    a->junc_start = find_jss(&T[0]);
    compute_alt_editing(a, tmpl, T);
    a->junc_end = find_jes(&T[1]);
    a->edit_stop = a->junc_start - 1;

    if (a->junc_end > a->edit_stop) {
        a->junc_len = a->junc_end - a->edit_stop;
        if (a->has_mutation == 0) {
            int from = (len - 1) - a->junc_end;
            int to = (len - 1) - a->edit_stop;
            for (int i = from; i < to; i++) {
                if ((size_t)i > frag->n) { /* Go: index out of range on frag.EditSite -> panic */
                    a->would_panic = 1;
                    break;
                }
                buf_rep(&js, frag->edit_base, frag->edit_site[i]);
                if ((size_t)i < frag->n) buf_putc(&js, frag->bases[i]);
            }
        }
    } else if (a->junc_end < a->edit_stop) {
        a->junc_end = a->edit_stop;
        a->junc_len = 0;
    }
    if (bs_all(&T[0])) {
        a->junc_end = a->junc_start;
        a->edit_stop = a->junc_start;
        a->junc_len = 0;
        js.len = 0;
        js.p[0] = 0;
    }
    if (a->would_panic) { /* the reference produces no value; keep the fields, drop the partial string */
        js.len = 0;
        js.p[0] = 0;
    }
    a->junc_seq = js.p;
    a->junc_seq_len = js.len;
    a->read_count = frag->read_count;
    a->norm = frag->norm;
    a->edit_stop += (int)tmpl->edit_offset;
    a->junc_start += (int)tmpl->edit_offset;
    a->junc_end += (int)tmpl->edit_offset;

    for (size_t k = 0; k < tmpl->K; k++) free(T[k].bit);
    free(T);
    if (ops_fwd_out) memcpy(ops_fwd_out, fwd, alen);
    if (alen_out) *alen_out = alen;
    free(fwd);
    return a;
}

/* align.go:237-279 */
orc_alignment *orc_new_alignment(const orc_fragment *frag, const orc_template *tmpl, int exclude_snps)
{
    return new_alignment_ops(frag, tmpl, exclude_snps, NULL, NULL);
}

void orc_alignment_free(orc_alignment *a)
{
    if (!a) return;
    free(a->junc_seq);
    free(a);
}

/* align.go:88-93 */
static void write_base(buf_t *b, char base, uint32_t count, uint32_t max)
{
    buf_rep(b, '-', max - count);
    if (count > 0) buf_rep(b, base, count);
}

/* align.go:337-386 */
void orc_simple_align(const orc_fragment *f1, const orc_fragment *f2, char **s1, char **s2)
{
    char *aln1, *aln2;
    size_t n = orc_nw_align(f1->bases, f1->n, f2->bases, f2->n, 1, -1, -1, &aln1, &aln2, NULL);
    buf_t b0 = {0}, b1 = {0};
    buf_reserve(&b0, 0);
    buf_reserve(&b1, 0);
    size_t fi = 0, ti = 0;
    for (size_t ai = 0; ai < n; ai++) {
        if (aln1[ai] == '-') {
            write_base(&b0, f1->edit_base, 0, f2->edit_site[fi]);
            buf_putc(&b0, '-');
            write_base(&b1, f2->edit_base, f2->edit_site[fi], f2->edit_site[fi]);
            buf_putc(&b1, f2->bases[fi]);
            fi++;
        } else if (aln2[ai] == '-') {
            write_base(&b0, f1->edit_base, f1->edit_site[ti], f1->edit_site[ti]);
            buf_putc(&b0, f1->bases[ti]);
            write_base(&b1, '-', 0, f1->edit_site[ti]);
            buf_putc(&b1, '-');
            ti++;
        } else {
            uint32_t mx = f1->edit_site[ti];
            if (f2->edit_site[fi] > mx) mx = f2->edit_site[fi];
            write_base(&b0, f1->edit_base, f1->edit_site[ti], mx);
            buf_putc(&b0, f1->bases[ti]);
            write_base(&b1, f2->edit_base, f2->edit_site[fi], mx);
            buf_putc(&b1, f2->bases[fi]);
            fi++;
            ti++;
        }
    }
    uint32_t mx = f1->edit_site[ti];
    if (f2->edit_site[fi] > mx) mx = f2->edit_site[fi];
    write_base(&b0, f1->edit_base, f1->edit_site[ti], mx);
    write_base(&b1, f2->edit_base, f2->edit_site[fi], mx);
    free(aln1);
    free(aln2);
    *s1 = b0.p;
    *s2 = b1.p;
}

/* align.go:388-520 */
char *orc_write_to(const orc_alignment *a, const orc_fragment *frag, const orc_template *tmpl, int tw,
                   size_t *outlen)
{
    if (tw <= 0) tw = 80;
    char *aln1, *aln2;
    size_t n = orc_nw_align(tmpl->bases, tmpl->m, frag->bases, frag->n, 1, -1, -1, &aln1, &aln2, NULL);
    size_t fc = tmpl->K + 1;
    buf_t *buf = xcalloc(fc, sizeof *buf);
    for (size_t i = 0; i < fc; i++) buf_reserve(&buf[i], 0);
    size_t fi = 0, ti = 0;
    for (size_t ai = 0; ai < n; ai++) {
        if (aln1[ai] == '-') {
            for (size_t k = 0; k < tmpl->K; k++) {
                write_base(&buf[k], tmpl->edit_base, 0, frag->edit_site[fi]);
                buf_putc(&buf[k], '-');
            }
            write_base(&buf[fc - 1], frag->edit_base, frag->edit_site[fi], frag->edit_site[fi]);
            buf_putc(&buf[fc - 1], frag->bases[fi]);
            fi++;
        } else if (aln2[ai] == '-') {
            uint32_t mx = orc_template_max(tmpl, ti);
            for (size_t k = 0; k < tmpl->K; k++) {
                write_base(&buf[k], tmpl->edit_base, tmpl->edit_site[k][ti], mx);
                buf_putc(&buf[k], tmpl->bases[ti]);
            }
            write_base(&buf[fc - 1], '-', 0, mx);
            buf_putc(&buf[fc - 1], '-');
            ti++;
        } else {
            uint32_t mx = orc_template_max(tmpl, ti);
            if (frag->edit_site[fi] > mx) mx = frag->edit_site[fi];
            for (size_t k = 0; k < tmpl->K; k++) {
                write_base(&buf[k], tmpl->edit_base, tmpl->edit_site[k][ti], mx);
                buf_putc(&buf[k], tmpl->bases[ti]);
            }
            write_base(&buf[fc - 1], frag->edit_base, frag->edit_site[fi], mx);
            buf_putc(&buf[fc - 1], frag->bases[fi]);
            fi++;
            ti++;
        }
    }
    uint32_t mx = orc_template_max(tmpl, ti);
    if (frag->edit_site[fi] > mx) mx = frag->edit_site[fi];
    for (size_t k = 0; k < tmpl->K; k++) write_base(&buf[k], tmpl->edit_base, tmpl->edit_site[k][ti], mx);
    write_base(&buf[fc - 1], frag->edit_base, frag->edit_site[fi], mx);

    size_t cols = (size_t)(tw - 4);
    size_t rows = buf[0].len / cols;
    if (buf[0].len % cols > 0) rows++;

    buf_t out = {0};
    buf_reserve(&out, 0);
    char line[160];
    buf_rep(&out, '=', (size_t)tw);
    buf_putc(&out, '\n');
    buf_puts(&out, frag->name);
    buf_putc(&out, '\n');
    buf_rep(&out, '=', (size_t)tw);
    buf_putc(&out, '\n');
    snprintf(line, sizeof line, "JSS: %d\nESS: %d\nJES: %d\nJunc Len: %d\n", a->junc_start, a->edit_stop,
             a->junc_end, a->junc_len);
    buf_puts(&out, line);
    if (a->alt_editing > 0) {
        snprintf(line, sizeof line, "Alt: %d\n", a->alt_editing);
        buf_puts(&out, line);
    }
    buf_rep(&out, '=', (size_t)tw);
    buf_puts(&out, "\n\n");
    for (size_t r = 0; r < rows; r++) {
        for (size_t i = 0; i < fc; i++) {
            /* labels: FE, PE, A1..A<len(AltRegion)>, RD (align.go:450-454) */
            if (i == 0) buf_puts(&out, "FE");
            else if (i == 1) buf_puts(&out, "PE");
            else if (i == fc - 1) buf_puts(&out, "RD");
            else {
                snprintf(line, sizeof line, "A%zu", i - 1);
                buf_puts(&out, line);
            }
            buf_puts(&out, ": ");
            size_t end = r * cols + cols;
            if (end > buf[i].len) end = buf[i].len;
            if (end > r * cols) buf_put(&out, buf[i].p + r * cols, end - r * cols);
            buf_putc(&out, '\n');
        }
        buf_putc(&out, '\n');
    }
    for (size_t i = 0; i < fc; i++) free(buf[i].p);
    free(buf);
    free(aln1);
    free(aln2);
    if (outlen) *outlen = out.len;
    return out.p;
}

/* align.go:316-335 (writeInt64 writes the low 32 bits big-endian, align.go:288-293) */
static void be32(uint8_t *b, uint32_t v)
{
    b[0] = (uint8_t)(v >> 24);
    b[1] = (uint8_t)(v >> 16);
    b[2] = (uint8_t)(v >> 8);
    b[3] = (uint8_t)v;
}
uint8_t *orc_marshal_binary(const orc_alignment *a, size_t *outlen)
{
    uint8_t *buf = xcalloc(36 + a->junc_seq_len, 1);
    be32(buf + 0, (uint32_t)a->edit_stop);
    be32(buf + 4, (uint32_t)a->junc_start);
    be32(buf + 8, (uint32_t)a->junc_end);
    be32(buf + 12, (uint32_t)a->junc_len);
    be32(buf + 16, a->read_count);
    buf[20] = a->has_mutation;
    buf[21] = a->alt_editing;
    buf[22] = a->mismatches;
    buf[23] = a->indel;
    uint64_t bits;
    memcpy(&bits, &a->norm, 8);
    be32(buf + 24, (uint32_t)(bits >> 32));
    be32(buf + 28, (uint32_t)bits);
    be32(buf + 32, (uint32_t)a->junc_seq_len);
    memcpy(buf + 36, a->junc_seq, a->junc_seq_len);
    *outlen = 36 + a->junc_seq_len;
    return buf;
}

/* ---------------------------------------------------------------- cmd/treat/align.go */

/* cmd/treat/align.go:39-57 */
static void print_alignment(buf_t *out, const char *a1, const char *a2, int tw)
{
    size_t n = strlen(a1), cols = (size_t)tw;
    size_t rows = n / cols;
    if (n % cols > 0) rows++;
    for (size_t r = 0; r < rows; r++) {
        size_t end = r * cols + cols;
        if (end > n) end = n;
        buf_put(out, a1 + r * cols, end - r * cols);
        buf_putc(out, '\n');
        buf_put(out, a2 + r * cols, end - r * cols);
        buf_putc(out, '\n');
        buf_putc(out, '\n');
    }
}

/* cmd/treat/align.go:59-122 */
char *orc_cli_align(const char *template_path, const char *fragment_path, const char *s1, const char *s2,
                    char base, int offset, size_t *outlen, char *err, size_t errlen)
{
    buf_t out = {0};
    buf_reserve(&out, 0);
    int has1 = s1 && *s1, has2 = s2 && *s2;
    if (has1 != has2) {
        snprintf(err, errlen, "Please provide 2 fragments to align");
        free(out.p);
        return NULL;
    }
    if (has1 && has2) {
        orc_fragment *f1 = orc_new_fragment("1-1", s1, strlen(s1), ORC_FORWARD, base);
        orc_fragment *f2 = orc_new_fragment("2-1", s2, strlen(s2), ORC_FORWARD, base);
        char *a1, *a2;
        orc_simple_align(f1, f2, &a1, &a2);
        print_alignment(&out, a1, a2, 80);
        free(a1);
        free(a2);
        orc_fragment_free(f1);
        orc_fragment_free(f2);
        *outlen = out.len;
        return out.p;
    }
    if (!fragment_path || !*fragment_path) {
        snprintf(err, errlen, "Please provide either 2 sequences to align or a path to fragment FASTA file");
        free(out.p);
        return NULL;
    }
    orc_template *tmpl = NULL;
    if (template_path && *template_path) {
        tmpl = orc_template_from_fasta(template_path, ORC_FORWARD, base, err, errlen);
        if (!tmpl) {
            free(out.p);
            return NULL;
        }
        orc_template_set_offset(tmpl, offset);
    }
    orc_fasta *fa = orc_fasta_read(fragment_path);
    if (!fa) {
        snprintf(err, errlen, "open %s: no such file or directory", fragment_path);
        orc_template_free(tmpl);
        free(out.p);
        return NULL;
    }
    if (!tmpl) {
        if (fa->count < 2) {
            snprintf(err, errlen, "Need 2 fragments to align");
            orc_fasta_free(fa);
            free(out.p);
            return NULL;
        }
        orc_fragment *f1 = orc_new_fragment(fa->ids[0], fa->seqs[0], strlen(fa->seqs[0]), ORC_FORWARD, base);
        orc_fragment *f2 = orc_new_fragment(fa->ids[1], fa->seqs[1], strlen(fa->seqs[1]), ORC_FORWARD, base);
        char *a1, *a2;
        orc_simple_align(f1, f2, &a1, &a2);
        print_alignment(&out, a1, a2, 80);
        free(a1);
        free(a2);
        orc_fragment_free(f1);
        orc_fragment_free(f2);
    } else {
        for (size_t i = 0; i < fa->count; i++) {
            orc_fragment *fr = orc_new_fragment(fa->ids[i], fa->seqs[i], strlen(fa->seqs[i]), ORC_FORWARD, base);
            orc_alignment *al = orc_new_alignment(fr, tmpl, 0);
            size_t tl;
            char *txt = orc_write_to(al, fr, tmpl, 80, &tl);
            buf_put(&out, txt, tl);
            free(txt);
            orc_alignment_free(al);
            orc_fragment_free(fr);
        }
    }
    orc_fasta_free(fa);
    orc_template_free(tmpl);
    *outlen = out.len;
    return out.p;
}

/* ---------------------------------------------------------------- batch driver */

typedef struct {
    const orc_template *tmpl;
    const uint8_t *seqs;
    const uint64_t *seq_off;
    const uint32_t *read_count;
    uint64_t lo, hi;
    int exclude_snps, mode;
    orc_record *rec;
    uint8_t *ops;
    size_t ops_stride;
    uint64_t text_bytes;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *jb = arg;
    const orc_template *tmpl = jb->tmpl;
    uint8_t *opsbuf = NULL;
    size_t opscap = 0;
    for (uint64_t r = jb->lo; r < jb->hi; r++) {
        const char *seq = (const char *)jb->seqs + jb->seq_off[r];
        size_t L = (size_t)(jb->seq_off[r + 1] - jb->seq_off[r]);
        orc_fragment *fr = orc_new_fragment("", seq, L, ORC_FORWARD, tmpl->edit_base);
        if (jb->read_count) fr->read_count = jb->read_count[r];
        if (tmpl->m + fr->n + 1 > opscap) {
            opscap = (tmpl->m + fr->n + 1) * 2;
            opsbuf = realloc(opsbuf, opscap);
        }
        size_t alen = 0;
        orc_alignment *a = new_alignment_ops(fr, tmpl, jb->exclude_snps, opsbuf, &alen);
        orc_record *o = &jb->rec[r];
        memset(o, 0, sizeof *o);
        o->edit_stop = a->edit_stop;
        o->junc_start = a->junc_start;
        o->junc_end = a->junc_end;
        o->junc_len = a->junc_len;
        o->read_count = a->read_count;
        o->has_mutation = a->has_mutation;
        o->alt_editing = a->alt_editing;
        o->mismatches = a->mismatches;
        o->indel = a->indel;
        o->score = a->score;
        o->aln_len = (uint16_t)alen;
        o->n_bases = (uint16_t)fr->n;
        o->status = a->would_panic ? 1 : 0;
        o->junc_seq_len = (uint32_t)a->junc_seq_len;
        o->junc_seq_off = 0;
        if (a->junc_seq_len) {
            /* JuncSeq is the upper-cased raw read from the start of the edit-base run in
             * front of base `from` (align.go:247-258) */
            int len = (int)tmpl->m + 1;
            int from = (len - 1) - (a->junc_end - (int)tmpl->edit_offset);
            uint64_t p = 0;
            for (int j = 0; j < from; j++) p += fr->edit_site[j] + 1;
            o->junc_seq_off = (uint32_t)p;
        }
        if (jb->ops) {
            uint8_t *dst = jb->ops + r * jb->ops_stride;
            memset(dst, 0, jb->ops_stride);
            for (size_t c = 0; c < alen && c / 4 < jb->ops_stride; c++)
                dst[c / 4] |= (uint8_t)(opsbuf[c] << (2 * (c % 4)));
        }
        if (jb->mode == 1) {
            size_t tl;
            char *txt = orc_write_to(a, fr, tmpl, 80, &tl);
            jb->text_bytes += tl;
            free(txt);
        }
        orc_alignment_free(a);
        orc_fragment_free(fr);
    }
    free(opsbuf);
    return NULL;
}

uint64_t orc_align_batch(const orc_template *tmpl, const uint8_t *seqs, const uint64_t *seq_off,
                         const uint32_t *read_count, uint64_t N, int exclude_snps, int mode, int nthreads,
                         orc_record *rec, uint8_t *ops, size_t ops_stride)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > N && N > 0) nthreads = (int)N;
    batch_job *jobs = xcalloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = xcalloc((size_t)nthreads, sizeof *th);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (batch_job){tmpl, seqs, seq_off, read_count, N * (uint64_t)t / (uint64_t)nthreads,
                              N * (uint64_t)(t + 1) / (uint64_t)nthreads, exclude_snps, mode, rec, ops,
                              ops_stride, 0};
        if (nthreads == 1) batch_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    uint64_t total = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        total += jobs[t].text_bytes;
    }
    free(jobs);
    free(th);
    return total;
}

/* ---------------------------------------------------------------- text of a batch (`treat align -t -f` stdout)
 * The per-read loop of cmd/treat/align.go:113-121 -- NewFragment, NewAlignment, WriteTo(tw = 80) -- for reads
 * [0, N) with names "<prefix><index>-<read_count>", threaded; the texts are concatenated in read order.  Returns
 * a malloc'ed buffer (release with orc_free) and its length; text_off (N + 1, may be NULL) gets the offsets. */
typedef struct {
    const orc_template *tmpl;
    const uint8_t *seqs;
    const uint64_t *seq_off;
    const uint32_t *read_count;
    const char *prefix;
    uint64_t lo, hi;
    char **txt;
    size_t *len;
} text_job;

static void *text_worker(void *arg)
{
    text_job *jb = arg;
    char name[128];
    for (uint64_t r = jb->lo; r < jb->hi; r++) {
        const char *seq = (const char *)jb->seqs + jb->seq_off[r];
        size_t L = (size_t)(jb->seq_off[r + 1] - jb->seq_off[r]);
        snprintf(name, sizeof name, "%s%llu-%u", jb->prefix, (unsigned long long)r, jb->read_count ? jb->read_count[r] : 1u);
        orc_fragment *fr = orc_new_fragment(name, seq, L, ORC_FORWARD, jb->tmpl->edit_base);
        orc_alignment *a = orc_new_alignment(fr, jb->tmpl, 0);
        jb->txt[r] = orc_write_to(a, fr, jb->tmpl, 80, &jb->len[r]);
        orc_alignment_free(a);
        orc_fragment_free(fr);
    }
    return NULL;
}

char *orc_text_batch(const orc_template *tmpl, const uint8_t *seqs, const uint64_t *seq_off, const uint32_t *read_count,
                     uint64_t N, const char *prefix, int nthreads, uint64_t *text_off, uint64_t *total_out)
{
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > N && N > 0) nthreads = (int)N;
    char **txt = xcalloc(N ? N : 1, sizeof *txt);
    size_t *len = xcalloc(N ? N : 1, sizeof *len);
    text_job *jobs = xcalloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = xcalloc((size_t)nthreads, sizeof *th);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (text_job){tmpl, seqs, seq_off, read_count, prefix ? prefix : "", N * (uint64_t)t / (uint64_t)nthreads,
                             N * (uint64_t)(t + 1) / (uint64_t)nthreads, txt, len};
        if (nthreads == 1) text_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, text_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++)
        if (nthreads > 1) pthread_join(th[t], NULL);
    uint64_t total = 0;
    for (uint64_t r = 0; r < N; r++) total += len[r];
    char *out = malloc(total ? total : 1);
    if (!out) abort();
    uint64_t at = 0;
    for (uint64_t r = 0; r < N; r++) {
        if (text_off) text_off[r] = at;
        memcpy(out + at, txt[r], len[r]);
        at += len[r];
        free(txt[r]);
    }
    if (text_off) text_off[N] = at;
    free(txt);
    free(len);
    free(jobs);
    free(th);
    if (total_out) *total_out = total;
    return out;
}
