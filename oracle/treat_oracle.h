/*
 * oracle/treat_oracle.h -- CPU restatement of ubccr/treat's read-vs-template
 * alignment path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  The product (treat_b200/,
 * libtreat_b200.so) never links, imports or calls it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).  The Needleman-Wunsch itself lives in the
 * un-vendored dependency github.com/aebruno/nwalgo
 * @ v0.0.0-20160817130739-4a232086e3ad (go.mod:7); orc_nw_align restates its
 * published algorithm.  PARITY STATUS: summaries pinned by the reference's own
 * goldens (align_test.go:68-137, README.rst:70-82,176-179); the Left-vs-Diagonal
 * traceback tie order is NOT pinned by any reference test, so
 * gapped strings of reads with insertions are "parity unpinned".
 */
#ifndef TREAT_ORACLE_H
#define TREAT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FORWARD 1  /* const.go:22 */
#define ORC_REVERSE (-1) /* const.go:23 */

/* traceback tie orders (first listed wins).  nwalgo: Up, then Left, then Diag. */
#define ORC_TIE_ULD 0
#define ORC_TIE_UDL 1
#define ORC_TIE_LUD 2
#define ORC_TIE_LDU 3
#define ORC_TIE_DUL 4
#define ORC_TIE_DLU 5

/* fragment.go:57-64 */
typedef struct {
    char *name;
    uint32_t read_count;
    double norm;
    char *bases;         /* n bytes + NUL */
    size_t n;
    char edit_base;
    uint32_t *edit_site; /* n + 1 */
} orc_fragment;

/* template.go:36-49 */
typedef struct {
    char *bases;
    size_t m;
    uint32_t edit_offset;
    int edit_stop;
    char edit_base;
    size_t K;             /* Size(): FE, PE, alts */
    uint32_t **edit_site; /* [K][m+1] */
    uint32_t *base_index; /* [m+1] */
    size_t n_alt;
    int *alt_start;
    int *alt_end;
} orc_template;

/* align.go:41-55 */
typedef struct {
    int edit_stop, junc_start, junc_end, junc_len;
    uint32_t read_count;
    double norm;
    uint8_t has_mutation, mismatches, indel, alt_editing;
    char *junc_seq; /* NUL terminated, may be empty */
    size_t junc_seq_len;
    int would_panic; /* reference indexes frag.EditSite out of range (align.go:250-257) */
    int score;       /* nwalgo's third return value */
} orc_alignment;

/* flat record used by the batch entry point (host endianness) */
typedef struct {
    int32_t edit_stop, junc_start, junc_end, junc_len;
    uint32_t read_count;
    uint8_t has_mutation, alt_editing, mismatches, indel;
    int32_t score;
    uint32_t junc_seq_off; /* offset of JuncSeq inside the upper-cased raw read */
    uint32_t junc_seq_len;
    uint16_t aln_len;      /* number of alignment columns */
    uint16_t n_bases;      /* len(frag.Bases) */
    uint8_t status;        /* bit0: reference would panic */
    uint8_t pad[3];
} orc_record;

void orc_set_tie_order(int order);
int orc_get_tie_order(void);

uint32_t orc_parse_merge_count(const char *id, size_t len);                        /* fragment.go:75-89 */
orc_fragment *orc_new_fragment(const char *name, const char *seq, size_t seqlen,
                               int orientation, char base);                        /* fragment.go:91-121 */
char *orc_fragment_string(const orc_fragment *f);                                  /* fragment.go:136-147 */
void orc_fragment_free(orc_fragment *f);

orc_template *orc_new_template(const orc_fragment *full, const orc_fragment *pre,
                               const orc_fragment *const *alt, size_t n_alt,
                               const int *alt_start, const int *alt_end, size_t n_region,
                               char *err, size_t errlen);                          /* template.go:96-151 */
orc_template *orc_template_from_fasta(const char *path, int orientation, char base,
                                      char *err, size_t errlen);                   /* template.go:51-94 */
void orc_template_set_offset(orc_template *t, int offset);                         /* template.go:153-161 */
uint32_t orc_template_max(const orc_template *t, size_t i);                        /* template.go:186-194 */
void orc_template_free(orc_template *t);

/* nwalgo.Align(a, b, match, mismatch, gap): returns alignment length; aln1/aln2 malloc'ed */
size_t orc_nw_align(const char *a, size_t m, const char *b, size_t n, int match, int mismatch,
                    int gap, char **aln1, char **aln2, int *score);

orc_alignment *orc_new_alignment(const orc_fragment *frag, const orc_template *tmpl,
                                 int exclude_snps);                                /* align.go:237-279 */
void orc_alignment_free(orc_alignment *a);
/* align.go:388-520; returns malloc'ed text, *outlen bytes */
char *orc_write_to(const orc_alignment *a, const orc_fragment *frag, const orc_template *tmpl,
                   int tw, size_t *outlen);
/* align.go:337-386 */
void orc_simple_align(const orc_fragment *f1, const orc_fragment *f2, char **s1, char **s2);
/* align.go:316-335; returns malloc'ed bytes */
uint8_t *orc_marshal_binary(const orc_alignment *a, size_t *outlen);

/* FASTA reader restating github.com/aebruno/gofasta SimpleParser (assumptions in treat_oracle.c) */
typedef struct {
    size_t count;
    char **ids;
    char **seqs;
} orc_fasta;
orc_fasta *orc_fasta_read(const char *path);
void orc_fasta_free(orc_fasta *f);

/* cmd/treat/align.go:59-122 `treat align`; returns malloc'ed stdout text or NULL + err */
char *orc_cli_align(const char *template_path, const char *fragment_path, const char *s1,
                    const char *s2, char base, int offset, size_t *outlen, char *err,
                    size_t errlen);

/*
 * Batch driver used for parity at scale and as the CPU baseline.
 * seqs/seq_off: concatenated raw reads, read i = seqs[seq_off[i] .. seq_off[i+1]).
 * ops (optional): ops_stride bytes per read, 2 bits per alignment column, 4 columns per
 * byte, low bits first: 0 = both consumed, 1 = template base vs '-', 2 = '-' vs read base.
 * mode 0 ("lean"): NewFragment + NewAlignment (one DP).  mode 1 ("faithful"): additionally
 * WriteTo into a scratch buffer (second DP + text), as `treat align` does.
 * Returns total bytes of text produced (mode 1) or 0.
 */
uint64_t orc_align_batch(const orc_template *tmpl, const uint8_t *seqs, const uint64_t *seq_off,
                         const uint32_t *read_count, uint64_t N, int exclude_snps, int mode,
                         int nthreads, orc_record *rec, uint8_t *ops, size_t ops_stride);

void orc_free(void *p);

/* cmd/treat/align.go:113-121 for reads [0, N) named "<prefix><index>-<read_count>": concatenated WriteTo(tw = 80) texts */
char *orc_text_batch(const orc_template *tmpl, const uint8_t *seqs, const uint64_t *seq_off, const uint32_t *read_count,
                     uint64_t N, const char *prefix, int nthreads, uint64_t *text_off, uint64_t *total_out);

#ifdef __cplusplus
}
#endif
#endif
