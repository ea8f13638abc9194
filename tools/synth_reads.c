/*
 * tools/synth_reads.c -- deterministic synthetic read generator.
 * Bench/test tooling, not part of the product path.
 *
 * Every read is generated from a counter-based splitmix64 stream keyed by (seed, read index),
 * so any shard [first, first+N) can be produced independently on any rank.
 *
 * Model (site s = 0 is the 3'-most edit site; array index ti = Len-1-s):
 *   ess ~ U{-1..Len-1}; sites <= ess copy the fully edited template (FE);
 *   jl  ~ min(Geom(1/8), Len-1-ess) junction sites with counts U{0..8}, whose two end sites are
 *         redrawn until they differ from FE (at ess+1) and from PE (at ess+jl);
 *   sites above the junction copy the pre-edited template (PE);
 *   p=0.02: the sites of one alternative template's region copy that template;
 *   p=0.05 one substitution, p=0.01 three substitutions, p=0.02 one single-base insertion or
 *   deletion, p=0.002 one 'N'; p_trunc: cut U{0..100} bases off the 5' or 3' end;
 *   ReadCount c ~ 1 + Geom(1/50).
 * synth_reads_noisy adds, from a second stream keyed the same way (so the base model's draws stay as they are):
 *   every base substituted with probability p_sub_base, and with probability p_indel_read one more single-base
 *   insertion or deletion -- the "noisy" point of bench.py (sequencing errors on top of the editing pattern).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint32_t rnd(uint64_t *s, uint32_t n) { return (uint32_t)(splitmix64(s) % n); } /* U{0..n-1} */
static inline double uni(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }
static uint32_t geom(uint64_t *s, double p) /* failures before the first success */
{
    uint32_t k = 0;
    while (uni(s) >= p && k < 100000) k++;
    return k;
}

/* upper bound of the raw length of any generated read */
uint64_t synth_max_len(int m, const uint32_t *edit_site, int K)
{
    uint64_t tot = (uint64_t)m + 2;
    for (int i = 0; i <= m; i++) {
        uint32_t mx = 8;
        for (int k = 0; k < K; k++)
            if (edit_site[(size_t)k * (m + 1) + i] > mx) mx = edit_site[(size_t)k * (m + 1) + i];
        tot += mx;
    }
    return tot;
}

static const char alpha[3] = {'A', 'C', 'G'};

/* one single-base insertion or deletion at a random base of (bs, cnt, n) */
static int one_indel(uint64_t *st, uint8_t *bs, uint32_t *cnt, int n)
{
    const int p = (int)rnd(st, (uint32_t)n);
    if (rnd(st, 2)) { /* deletion of base p: its edit-base run merges into the next site */
        cnt[p + 1] += cnt[p];
        memmove(bs + p, bs + p + 1, (size_t)(n - p - 1));
        memmove(cnt + p, cnt + p + 1, sizeof(uint32_t) * (size_t)(n - p));
        return n - 1;
    }
    /* insertion in front of base p, carrying no edit bases */
    memmove(bs + p + 1, bs + p, (size_t)(n - p));
    memmove(cnt + p + 1, cnt + p, sizeof(uint32_t) * (size_t)(n + 1 - p));
    bs[p] = (uint8_t)alpha[rnd(st, 3)];
    cnt[p] = 0;
    return n + 1;
}

/* returns the number of sequence bytes written, or 0 if cap is too small */
uint64_t synth_reads_noisy(const uint8_t *bases, int m, const uint32_t *edit_site, int K, const int *alt_start,
                           const int *alt_end, uint8_t edit_base, uint64_t seed, uint64_t first, uint64_t N,
                           double p_trunc, double p_sub_base, double p_indel_read, uint8_t *seqs, uint64_t cap,
                           uint64_t *off, uint32_t *read_count)
{
    const int len = m + 1;
    const uint32_t *FE = edit_site, *PE = edit_site + len;
    uint32_t *cnt = malloc(sizeof(uint32_t) * (size_t)(len + 4));
    uint8_t *bs = malloc((size_t)m + 8);
    uint64_t pos = 0;
    off[0] = 0;
    for (uint64_t r = 0; r < N; r++) {
        uint64_t st = seed ^ ((first + r + 1) * 0xD1342543DE82EF95ULL);
        splitmix64(&st);
        int ess = (int)rnd(&st, (uint32_t)len + 1) - 1; /* -1 .. len-1 */
        int jl = (int)geom(&st, 1.0 / 8.0);
        if (jl > len - 1 - ess) jl = len - 1 - ess;
        for (int s = 0; s < len; s++) {
            const int ti = len - 1 - s;
            if (s <= ess) cnt[ti] = FE[ti];
            else if (s <= ess + jl) cnt[ti] = rnd(&st, 9);
            else cnt[ti] = PE[ti];
        }
        if (jl > 0) {
            int ti = len - 1 - (ess + 1);
            while (cnt[ti] == FE[ti]) cnt[ti] = rnd(&st, 9);
            ti = len - 1 - (ess + jl);
            if (jl > 1 || cnt[ti] == PE[ti]) {
                int guard = 0;
                while ((cnt[ti] == PE[ti] || (jl == 1 && cnt[ti] == FE[ti])) && guard++ < 64) cnt[ti] = rnd(&st, 9);
            }
        }
        if (K > 2 && uni(&st) < 0.02) {
            const int a = (int)rnd(&st, (uint32_t)(K - 2));
            const uint32_t *AT = edit_site + (size_t)(a + 2) * len;
            for (int s = alt_start[a]; s <= alt_end[a] && s < len; s++)
                if (s >= 0) cnt[len - 1 - s] = AT[len - 1 - s];
        }
        memcpy(bs, bases, (size_t)m);
        int n = m;
        double u = uni(&st);
        int nsub = u < 0.05 ? 1 : (u < 0.06 ? 3 : 0);
        for (int q = 0; q < nsub && n > 0; q++) {
            const int p = (int)rnd(&st, (uint32_t)n);
            char c;
            do c = alpha[rnd(&st, 3)]; while (c == (char)bs[p]);
            bs[p] = (uint8_t)c;
        }
        if (uni(&st) < 0.02 && n > 1) n = one_indel(&st, bs, cnt, n);
        if (uni(&st) < 0.002 && n > 0) bs[rnd(&st, (uint32_t)n)] = 'N';
        if (p_sub_base > 0 || p_indel_read > 0) { /* sequencing noise, from its own stream */
            uint64_t s2 = seed ^ ((first + r + 1) * 0x9FB21C651E98DF25ULL);
            splitmix64(&s2);
            for (int i = 0; i < n; i++)
                if (uni(&s2) < p_sub_base) {
                    char c;
                    do c = alpha[rnd(&s2, 3)]; while (c == (char)bs[i]);
                    bs[i] = (uint8_t)c;
                }
            if (uni(&s2) < p_indel_read && n > 1) n = one_indel(&s2, bs, cnt, n);
        }
        int lo = 0, hi = n; /* bases kept: [lo, hi) */
        if (p_trunc > 0 && uni(&st) < p_trunc) {
            int cut = (int)rnd(&st, 101);
            if (cut > n - 1) cut = n - 1;
            if (rnd(&st, 2)) lo = cut; else hi = n - cut;
        }
        uint64_t need = 0;
        for (int i = lo; i < hi; i++) need += cnt[i] + 1;
        if (hi == n) need += cnt[n];
        if (pos + need > cap) {
            free(cnt);
            free(bs);
            return 0;
        }
        for (int i = lo; i < hi; i++) {
            memset(seqs + pos, edit_base, cnt[i]);
            pos += cnt[i];
            seqs[pos++] = bs[i];
        }
        if (hi == n) {
            memset(seqs + pos, edit_base, cnt[n]);
            pos += cnt[n];
        }
        off[r + 1] = pos;
        if (read_count) read_count[r] = 1 + geom(&st, 1.0 / 50.0);
    }
    free(cnt);
    free(bs);
    return pos;
}

uint64_t synth_reads(const uint8_t *bases, int m, const uint32_t *edit_site, int K, const int *alt_start,
                     const int *alt_end, uint8_t edit_base, uint64_t seed, uint64_t first, uint64_t N,
                     double p_trunc, uint8_t *seqs, uint64_t cap, uint64_t *off, uint32_t *read_count)
{
    return synth_reads_noisy(bases, m, edit_site, K, alt_start, alt_end, edit_base, seed, first, N, p_trunc, 0.0, 0.0, seqs,
                             cap, off, read_count);
}

/* synthetic long-gene template set (config 5): m bases U{A,C,G}; FE counts with an RPS12-like
 * histogram (mean ~1, max 8); PE sparse (~15% non-zero); n_alt alternatives = FE with an 8-site
 * window redrawn.  edit_site is [(2+n_alt)][m+1]; alt regions are in site numbering. */
void synth_long_template(int m, int n_alt, uint64_t seed, uint8_t *bases, uint32_t *edit_site, int *alt_start,
                         int *alt_end)
{
    static const char alpha[3] = {'A', 'C', 'G'};
    static const double cdf[9] = {0.45, 0.72, 0.86, 0.92, 0.95, 0.97, 0.985, 0.995, 1.0};
    const int len = m + 1;
    uint64_t st = seed;
    for (int i = 0; i < m; i++) bases[i] = (uint8_t)alpha[rnd(&st, 3)];
    for (int i = 0; i < len; i++) {
        const double u = uni(&st);
        uint32_t c = 0;
        while (c < 8 && u >= cdf[c]) c++;
        edit_site[i] = c;
        edit_site[len + i] = uni(&st) < 0.15 ? 1 + rnd(&st, 4) : 0;
    }
    /* keep the 5'-most sites identical so that the template edit stop is interior */
    for (int i = 0; i < 6 && i < len; i++) edit_site[len + i] = edit_site[i];
    for (int a = 0; a < n_alt; a++) {
        uint32_t *row = edit_site + (size_t)(2 + a) * len;
        memcpy(row, edit_site, sizeof(uint32_t) * (size_t)len);
        const int s0 = len > 32 ? 8 + (int)rnd(&st, (uint32_t)(len - 24)) : (int)rnd(&st, (uint32_t)(len > 8 ? len - 8 : 1));
        alt_start[a] = s0;
        alt_end[a] = s0 + 7;
        for (int s = s0; s <= s0 + 7 && s < len; s++) {
            const int ti = len - 1 - s;
            uint32_t c;
            do c = rnd(&st, 9); while (c == edit_site[ti]);
            row[ti] = c;
        }
    }
}
