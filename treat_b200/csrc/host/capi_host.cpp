// capi_host.cpp -- C exports of the host-side mirror (treat_host.hpp) declared in
// include/treat_b200.h: fragments, templates, FASTA, text formatting.  No GPU work except
// treat_b200_simple_align / treat_b200_cli_align, which call the batch entry point.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <string>

#include "treat_host.hpp"

using treat::Fragment;
using treat::Template;

namespace {
void set_err(char *err, size_t errlen, const std::string &msg)
{
    if (err && errlen) {
        const size_t n = std::min(errlen - 1, msg.size());
        memcpy(err, msg.data(), n);
        err[n] = 0;
    }
}
const Template *T(const treat_b200_template *t) { return reinterpret_cast<const Template *>(t); }
int64_t emit(const std::string &s, char *out, size_t cap)
{
    if (out && cap) memcpy(out, s.data(), std::min(cap, s.size()));
    return (int64_t)s.size();
}
}  // namespace

struct treat_b200_fasta {
    treat::FastaBatch b;
};

extern "C" {

uint32_t treat_b200_parse_merge_count(const char *id, size_t len) { return treat::parseMergeCount(std::string(id, len)); }

int treat_b200_new_fragment(const uint8_t *seq, size_t len, int orientation, uint8_t base, uint8_t *bases_out,
                            uint32_t *edit_site_out, size_t *n_out)
{
    if ((!seq && len) || !bases_out || !edit_site_out || !n_out) return TREAT_B200_ERR_INVALID;
    Fragment f = treat::NewFragment("", std::string((const char *)seq, len),
                                    orientation == TREAT_B200_REVERSE ? treat::REVERSE : treat::FORWARD, (char)base);
    memcpy(bases_out, f.Bases.data(), f.Bases.size());
    memcpy(edit_site_out, f.EditSite.data(), f.EditSite.size() * sizeof(uint32_t));
    *n_out = f.Bases.size();
    return TREAT_B200_OK;
}

int treat_b200_template_from_fasta(const char *path, int orientation, uint8_t base, int32_t offset,
                                   treat_b200_template **out, char *err, size_t errlen)
{
    if (!path || !out) return TREAT_B200_ERR_INVALID;
    *out = nullptr;
    Template *t = new Template();
    std::string e;
    if (!treat::NewTemplateFromFasta(path, orientation == TREAT_B200_REVERSE ? treat::REVERSE : treat::FORWARD,
                                     (char)base, t, &e)) {
        delete t;
        set_err(err, errlen, e);
        return e.rfind("Invalid FASTA file", 0) == 0 ? TREAT_B200_ERR_IO : TREAT_B200_ERR_TEMPLATE;
    }
    if (offset) t->SetOffset(offset);
    *out = reinterpret_cast<treat_b200_template *>(t);
    return TREAT_B200_OK;
}

int treat_b200_template_new(const char *const *seqs, int32_t n_seqs, const int32_t *alt_start, const int32_t *alt_end,
                            int32_t n_regions, int orientation, uint8_t base, treat_b200_template **out, char *err,
                            size_t errlen)
{
    if (!seqs || !out || n_seqs < 0 || n_regions < 0) return TREAT_B200_ERR_INVALID;
    *out = nullptr;
    if (n_seqs < 2) {
        set_err(err, errlen, "Must provide at least 2 templates. Full and Pre edited");
        return TREAT_B200_ERR_TEMPLATE;
    }
    const treat::OrientationType o = orientation == TREAT_B200_REVERSE ? treat::REVERSE : treat::FORWARD;
    std::vector<Fragment> fr;
    for (int i = 0; i < n_seqs; i++) fr.push_back(treat::NewFragment("", seqs[i], o, (char)base));
    std::vector<treat::AltRegion> regions;
    for (int i = 0; i < n_regions; i++) regions.push_back({alt_start ? alt_start[i] : 0, alt_end ? alt_end[i] : 0});
    Template *t = new Template();
    std::string e;
    if (!treat::NewTemplate(fr[0], fr[1], std::vector<Fragment>(fr.begin() + 2, fr.end()), regions, t, &e)) {
        delete t;
        set_err(err, errlen, e);
        return TREAT_B200_ERR_TEMPLATE;
    }
    *out = reinterpret_cast<treat_b200_template *>(t);
    return TREAT_B200_OK;
}

void treat_b200_template_set_offset(treat_b200_template *t, int32_t offset)
{
    if (t) reinterpret_cast<Template *>(t)->SetOffset(offset);
}
void treat_b200_template_free(treat_b200_template *t) { delete reinterpret_cast<Template *>(t); }
int32_t treat_b200_template_size(const treat_b200_template *t) { return T(t)->Size(); }
int32_t treat_b200_template_len(const treat_b200_template *t) { return T(t)->Len(); }
int32_t treat_b200_template_edit_stop(const treat_b200_template *t) { return T(t)->EditStop; }
uint32_t treat_b200_template_edit_offset(const treat_b200_template *t) { return T(t)->EditOffset; }
uint8_t treat_b200_template_edit_base(const treat_b200_template *t) { return (uint8_t)T(t)->EditBase; }
const uint8_t *treat_b200_template_bases(const treat_b200_template *t) { return (const uint8_t *)T(t)->Bases.data(); }
const uint32_t *treat_b200_template_edit_site(const treat_b200_template *t, int32_t k)
{
    return (k < 0 || k >= T(t)->Size()) ? nullptr : T(t)->EditSite[k].data();
}
const uint32_t *treat_b200_template_base_index(const treat_b200_template *t) { return T(t)->BaseIndex.data(); }
int32_t treat_b200_template_n_alt(const treat_b200_template *t) { return (int32_t)T(t)->AltRegion.size(); }
int32_t treat_b200_template_alt_start(const treat_b200_template *t, int32_t i) { return T(t)->AltRegion[i].Start; }
int32_t treat_b200_template_alt_end(const treat_b200_template *t, int32_t i) { return T(t)->AltRegion[i].End; }
uint32_t treat_b200_template_max(const treat_b200_template *t, int32_t i) { return T(t)->Max(i); }

int treat_b200_fasta_read(const char *path, treat_b200_fasta **out)
{
    if (!path || !out) return TREAT_B200_ERR_INVALID;
    treat_b200_fasta *f = new treat_b200_fasta();
    std::string e;
    if (!treat::ReadFasta(path, &f->b, &e)) {
        delete f;
        *out = nullptr;
        return TREAT_B200_ERR_IO;
    }
    *out = f;
    return TREAT_B200_OK;
}
int treat_b200_fasta_parse(const uint8_t *bytes, uint64_t len, treat_b200_fasta **out)
{
    if ((!bytes && len) || !out) return TREAT_B200_ERR_INVALID;
    treat_b200_fasta *f = new treat_b200_fasta();
    treat::ParseFasta((const char *)bytes, (size_t)len, &f->b);
    *out = f;
    return TREAT_B200_OK;
}
int treat_b200_fasta_split(const uint8_t *fasta, uint64_t len, uint32_t parts, uint64_t *offsets)
{
    if ((!fasta && len) || !parts || !offsets) return TREAT_B200_ERR_INVALID;
    treat::FastaSplit(fasta, len, parts, offsets);
    return TREAT_B200_OK;
}
void treat_b200_fasta_free(treat_b200_fasta *f) { delete f; }
uint64_t treat_b200_fasta_count(const treat_b200_fasta *f) { return f->b.size(); }
const char *treat_b200_fasta_id(const treat_b200_fasta *f, uint64_t i) { return f->b.ids[i].c_str(); }
const uint8_t *treat_b200_fasta_seqs(const treat_b200_fasta *f) { return f->b.seqs.data(); }
const uint64_t *treat_b200_fasta_offsets(const treat_b200_fasta *f) { return f->b.offsets.data(); }
const uint32_t *treat_b200_fasta_read_counts(const treat_b200_fasta *f) { return f->b.read_counts.data(); }

int64_t treat_b200_format_alignment(const treat_b200_template *t, const treat_b200_record *rec, const uint8_t *ops,
                                    const char *name, const uint8_t *seq, size_t seq_len, int32_t tw, char *out,
                                    size_t cap)
{
    if (!t || !rec || !ops || (!seq && seq_len)) return TREAT_B200_ERR_INVALID;
    if (tw > 0 && tw <= 4) return TREAT_B200_ERR_INVALID;  // the reference divides by tw-4 (align.go:450)
    const Template &tm = *T(t);
    Fragment frag = treat::NewFragment(name ? name : "", std::string((const char *)seq, seq_len), treat::FORWARD, tm.EditBase);
    treat::Alignment a = treat::Alignment::FromRecord(*rec, ops, seq, seq_len);
    const std::string text = a.WriteTo(frag, tm, tw);
    if (text.empty()) return TREAT_B200_ERR_INVALID;  // record/ops do not belong to this read and template
    return emit(text, out, cap);
}

int64_t treat_b200_simple_align(treat_b200_ctx *ctx, const uint8_t *s1, size_t l1, const uint8_t *s2, size_t l2,
                                uint8_t base, char *out1, char *out2, size_t cap)
{
    if (!ctx || (!s1 && l1) || (!s2 && l2)) return TREAT_B200_ERR_INVALID;
    Fragment f1 = treat::NewFragment("1-1", std::string((const char *)s1, l1), treat::FORWARD, (char)base);
    Fragment f2 = treat::NewFragment("2-1", std::string((const char *)s2, l2), treat::FORWARD, (char)base);
    treat::Aligner eng(ctx);
    std::string a1, a2, err;
    if (!eng.SimpleAlign(f1, f2, &a1, &a2, &err)) return TREAT_B200_ERR_CUDA;
    emit(a1, out1, cap);
    return emit(a2, out2, cap);
}

int64_t treat_b200_marshal_record(const treat_b200_record *rec, const uint8_t *seq, size_t seq_len, uint8_t *out,
                                  size_t cap)
{
    if (!rec) return TREAT_B200_ERR_INVALID;
    treat::Alignment a = treat::Alignment::FromRecord(*rec, nullptr, seq, seq_len);
    std::vector<uint8_t> b = a.MarshalBinary();
    if (out && cap) memcpy(out, b.data(), std::min(cap, b.size()));
    return (int64_t)b.size();
}

static bool search_match(const treat_b200_search &f, const treat_b200_record &a);

// highChartHist (cmd/treat/handlers.go:162-299): samples[key.Sample][f(a)] += a.Norm over the matching alignments
int treat_b200_hist_norm(const treat_b200_search *f, const treat_b200_record *rec, const double *norm, uint64_t N,
                         int32_t tmpl_edit_stop, int32_t which, int32_t lo, int32_t hi, double *out)
{
    if (!f || (N && (!rec || !norm)) || !out || hi < lo || which < 0 || which > 2) return TREAT_B200_ERR_INVALID;
    for (int64_t v = lo; v <= hi; v++) out[v - lo] = 0.0;
    for (uint64_t i = 0; i < N; i++) {
        const treat_b200_record &a = rec[i];
        if (!search_match(*f, a)) continue;                                  // Storage.Search with Limit = Offset = 0
        if (a.edit_stop == tmpl_edit_stop && a.junc_len == 0) continue;       // handlers.go:203-205
        const int32_t v = which == 0 ? a.edit_stop : which == 1 ? a.junc_len : a.junc_end;
        if (v >= lo && v <= hi) out[v - lo] += norm[i];
    }
    return TREAT_B200_OK;
}

// `treat search` rows (cmd/treat/search.go:30-89) through Go's encoding/csv writer
static void csv_field(std::string &out, const std::string &f, char comma)
{
    // encoding/csv Writer.fieldNeedsQuotes
    bool quote = false;
    if (!f.empty()) {
        if (f == "\\.") quote = true;
        for (char c : f)
            if (c == '\n' || c == '\r' || c == '"' || c == comma) quote = true;
        const unsigned char c0 = (unsigned char)f[0];
        if (c0 == ' ' || (c0 >= '\t' && c0 <= '\r')) quote = true;                                                  // unicode.IsSpace, ASCII
        if (c0 == 0xC2 && f.size() > 1 && ((unsigned char)f[1] == 0x85 || (unsigned char)f[1] == 0xA0)) quote = true;  // U+0085, U+00A0
    }
    if (!quote) {
        out += f;
        return;
    }
    out += '"';
    for (char c : f) {
        if (c == '"') out += "\"\"";
        else out += c;
    }
    out += '"';
}

int treat_b200_search_text(const char *gene, const char *sample, const treat_b200_record *rec, const double *norm, const uint8_t *seqs,
                           const uint64_t *seq_off, const uint64_t *idx, uint64_t n, int32_t csv, int32_t header, char **out,
                           size_t *out_len)
{
    if (!gene || !sample || !out || !out_len || (n && (!rec || !idx || !seqs || !seq_off))) return TREAT_B200_ERR_INVALID;
    const char comma = csv ? ',' : '\t';  // search.go:38-40
    std::string text;
    auto row = [&](const std::vector<std::string> &fields) {
        for (size_t k = 0; k < fields.size(); k++) {
            if (k) text += comma;
            csv_field(text, fields[k], comma);
        }
        text += '\n';
    };
    if (header) row({"gene", "sample", "norm", "read_count", "alt_editing", "has_mutation", "edit_stop", "junc_end", "junc_len", "junc_seq"});
    char buf[64];
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t i = idx[k];
        const treat_b200_record &a = rec[i];
        std::vector<std::string> f;
        f.emplace_back(gene);
        f.emplace_back(sample);
        const double v = norm ? norm[i] : 0.0;
        snprintf(buf, sizeof buf, "%.4f", floor(v * 10000.0 + .5) / 10000.0);  // RoundPlus(a.Norm, 4), storage.go:190-198
        f.emplace_back(buf);
        f.emplace_back(std::to_string(a.read_count));
        f.emplace_back(a.alt_editing ? "A" + std::to_string((int)a.alt_editing) : std::string("0"));  // search.go:58-61
        f.emplace_back(std::to_string((int)a.has_mutation));
        f.emplace_back(std::to_string(a.edit_stop));
        f.emplace_back(std::to_string(a.junc_end));
        f.emplace_back(std::to_string(a.junc_len));
        std::string js((const char *)seqs + seq_off[i] + a.junc_seq_off, a.junc_seq_len);
        for (char &c : js)
            if (c >= 'a' && c <= 'z') c = (char)(c - 32);
        f.emplace_back(js);
        row(f);
    }
    *out = (char *)malloc(text.size() + 1);
    if (!*out) return TREAT_B200_ERR_NOMEM;
    memcpy(*out, text.data(), text.size());
    (*out)[text.size()] = 0;
    *out_len = text.size();
    return TREAT_B200_OK;
}

// SearchFields.HasMatch (cmd/treat/storage.go:162-188) + the alt-editing rule of Storage.Search (:269-272)
static bool search_match(const treat_b200_search &f, const treat_b200_record &a)
{
    if (!f.all) {
        if (f.has_mutation && a.has_mutation == 0) return false;
        if (!f.has_mutation && a.has_mutation == 1) return false;
    }
    if (f.edit_stop >= 0 && f.edit_stop != a.edit_stop) return false;
    if (f.junc_len >= 0 && f.junc_len != a.junc_len) return false;
    if (f.junc_end >= 0 && f.junc_end != a.junc_end) return false;
    if (f.has_alt && a.alt_editing == 0) return false;
    if (f.alt_region > 0 && (uint8_t)f.alt_region != a.alt_editing) return false;
    if (!f.has_alt && a.alt_editing > 0) return false;  // by default, don't include alt editing
    return true;
}

int treat_b200_search_records(const treat_b200_search *f, const treat_b200_record *rec, uint64_t N, uint64_t *idx, uint64_t cap,
                              uint64_t *count)
{
    if (!f || (N && !rec) || !count) return TREAT_B200_ERR_INVALID;
    uint64_t n = 0, offset = 0;  // storage.go:241-242
    for (uint64_t i = 0; i < N; i++) {
        if (!search_match(*f, rec[i])) continue;
        if (f->offset > 0 && offset < (uint64_t)f->offset) {  // :274-277
            offset++;
            continue;
        }
        if (f->limit > 0 && n >= (uint64_t)f->limit) break;  // :279-281
        if (idx && n < cap) idx[n] = i;
        n++;
        offset++;
    }
    *count = n;
    return TREAT_B200_OK;
}

// `treat norm` (cmd/treat/norm.go:25-73, cmd/treat/storage.go:801-878)
int treat_b200_norm_default(const uint64_t *std_reads_per_sample, int32_t n_samples, double *norm)
{
    if (!std_reads_per_sample || n_samples < 1 || !norm) return TREAT_B200_ERR_INVALID;
    uint64_t total = 0;  // norm.go:52-56: total += int(a.ReadCount) over the standard reads of every sample
    for (int32_t i = 0; i < n_samples; i++) total += std_reads_per_sample[i];
    *norm = (double)total / (double)n_samples;  // norm.go:61
    return TREAT_B200_OK;
}

int treat_b200_norm_scale(uint64_t std_reads_of_sample, double norm, double *scale)
{
    if (!scale) return TREAT_B200_ERR_INVALID;
    *scale = std_reads_of_sample > 0 ? norm / (double)std_reads_of_sample : 1.0;  // storage.go:840-843
    return TREAT_B200_OK;
}

int treat_b200_normalize_records(const treat_b200_record *rec, uint64_t N, double scale, double *norm_out, uint8_t *blob,
                                 const uint64_t *blob_off)
{
    if ((N && !rec) || (blob && !blob_off)) return TREAT_B200_ERR_INVALID;
    for (uint64_t i = 0; i < N; i++) {
        const bool standard = rec[i].has_mutation == 0;
        const double v = standard ? scale * (double)rec[i].read_count : 0.0;  // storage.go:858-860
        if (norm_out) norm_out[i] = v;
        if (blob && standard) {
            uint64_t bits;
            memcpy(&bits, &v, 8);
            uint8_t *dst = blob + blob_off[i] + 24;  // align.go:328: binary.BigEndian.PutUint64(buf[24:32], Float64bits(a.Norm))
            for (int k = 0; k < 8; k++) dst[k] = (uint8_t)(bits >> (56 - 8 * k));
        }
    }
    return TREAT_B200_OK;
}

int64_t treat_b200_marshal_batch(const treat_b200_record *rec, const uint8_t *seqs, const uint64_t *seq_off, uint64_t N,
                                 uint8_t *out, uint64_t out_cap, uint64_t *out_off, int32_t threads)
{
    if (!rec && N) return TREAT_B200_ERR_INVALID;
    if ((!seqs || !seq_off) && N) return TREAT_B200_ERR_INVALID;
    const uint64_t need = treat::MarshalBatch(rec, seqs, seq_off, N, nullptr, nullptr, 1);
    if (!out || !out_off || out_cap < need) return (int64_t)need;
    for (uint64_t i = 0; i < N; i++)
        if ((uint64_t)rec[i].junc_seq_off + rec[i].junc_seq_len > seq_off[i + 1] - seq_off[i]) return TREAT_B200_ERR_INVALID;
    return (int64_t)treat::MarshalBatch(rec, seqs, seq_off, N, out, out_off, threads);
}

int treat_b200_mutant_report(const treat_b200_template *t, const treat_b200_record *rec, const uint8_t *ops,
                             uint32_t ops_stride, const uint8_t *seqs, const uint64_t *seq_off, uint64_t N, int32_t n,
                             char **out, size_t *out_len)
{
    if (!t || !out || !out_len || (N && (!rec || !ops || !seqs || !seq_off))) return TREAT_B200_ERR_INVALID;
    const std::string text = treat::MutantFromOps(*T(t), rec, ops, ops_stride, seqs, seq_off, N).Report(n);
    *out = (char *)malloc(text.size() + 1);
    if (!*out) return TREAT_B200_ERR_NOMEM;
    memcpy(*out, text.data(), text.size());
    (*out)[text.size()] = 0;
    *out_len = text.size();
    return TREAT_B200_OK;
}

int treat_b200_unmarshal_record(const uint8_t *blob, size_t len, treat_b200_record *rec, double *norm, const uint8_t **junc_seq)
{
    if (!blob || !rec) return TREAT_B200_ERR_INVALID;
    treat::Alignment a;
    if (!a.UnmarshalBinary(blob, len)) return TREAT_B200_ERR_INVALID;
    memset(rec, 0, sizeof *rec);
    rec->edit_stop = a.EditStop;
    rec->junc_start = a.JuncStart;
    rec->junc_end = a.JuncEnd;
    rec->junc_len = a.JuncLen;
    rec->read_count = a.ReadCount;
    rec->has_mutation = a.HasMutation;
    rec->alt_editing = a.AltEditing;
    rec->mismatches = a.Mismatches;
    rec->indel = a.Indel;
    rec->junc_seq_len = (uint32_t)a.JuncSeq.size();
    if (norm) *norm = a.Norm;
    if (junc_seq) *junc_seq = blob + 36;
    return TREAT_B200_OK;
}

int treat_b200_stats_text(const treat_b200_template *t, const char *gene, const char *const *samples,
                          const uint64_t *const *summaries, int32_t n_samples, int32_t unique, int32_t norm, char **out,
                          size_t *out_len)
{
    if (!t || !gene || !out || !out_len || n_samples < 0 || (n_samples && (!samples || !summaries))) return TREAT_B200_ERR_INVALID;
    const treat::Template &tm = *T(t);
    const int base = unique ? 7 : 0;  // COUNT_UNIQUE / COUNT_FRAG halves of the summary header
    // Stats{Total, Std, NonStd, Indels, Snps, SingleMismatch, DoubleMismatch} (stats.go:33-41) in summary order
    long long g[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_samples; i++)
        for (int k = 0; k < 7; k++) g[k] += (long long)summaries[i][base + k];
    auto pct = [](long long x, long long y) { return y == 0 ? 0.0 : (double)x / (double)y * 100.0; };  // stats.go:54-60
    std::string text;
    char line[256];
    auto put = [&](const char *fmt, auto... a) {
        snprintf(line, sizeof line, fmt, a...);
        text += line;
    };
    const std::string rule(80, '='), dash(80, '-');
    text += rule + "\n" + gene + "\n" + rule + "\n";
    put("%20s%11lld\n", "Total Alignments:", g[0]);
    if (!norm) {
        put("%20s%11lld\n", "Standard:", g[1]);
        put("%20s%11lld\n", "Non-Standard:", g[2]);
        put("%20s%11lld\n", "1-Mismatch:", g[5]);
        put("%20s%11lld\n", "2-Mismatch:", g[6]);
        put("%20s%11lld\n", ">3-Mismatch:", g[4]);
        put("%20s%11lld\n", "Indels:", g[3]);
    }
    put("%20s%11d\n", "Template Edit Stop:", tm.EditStop);
    put("%20s%11s\n", "Edit Base:", std::string(1, tm.EditBase).c_str());
    put("%20s%11d\n", "Alt Templates:", (int)tm.AltRegion.size());
    text += dash + "\n";
    if (!norm) put("%-15s%9s%9s%5s%9s%5s%9s%5s%9s%5s\n", "Sample", "Total", "Std", "%", "Non-Std", "%", "1MM", "%", "2MM", "%");
    else put("%-15s%9s\n", "Sample", "Total");
    text += dash + "\n";
    for (int i = 0; i < n_samples; i++) {
        std::string name = samples[i] ? samples[i] : "";
        if (name.size() > 12) name = name.substr(0, 12) + "..";
        const uint64_t *c = summaries[i] + base;
        if (!norm)
            put("%-15s%9lld%9lld%5.1f%9lld%5.1f%9lld%5.1f%9lld%5.1f\n", name.c_str(), (long long)c[0], (long long)c[1],
                pct((long long)c[1], (long long)c[0]), (long long)c[2], pct((long long)c[2], (long long)c[0]), (long long)c[5],
                pct((long long)c[5], (long long)c[0]), (long long)c[6], pct((long long)c[6], (long long)c[0]));
        else
            put("%-15s%9lld\n", name.c_str(), (long long)c[0]);
    }
    text += "\n";
    *out = (char *)malloc(text.size() + 1);
    if (!*out) return TREAT_B200_ERR_NOMEM;
    memcpy(*out, text.data(), text.size());
    (*out)[text.size()] = 0;
    *out_len = text.size();
    return TREAT_B200_OK;
}

int treat_b200_cli_align(treat_b200_ctx *ctx, const char *template_path, const char *fragment_path, const char *s1,
                         const char *s2, uint8_t base, int32_t offset, char **out, size_t *out_len, char *err,
                         size_t errlen)
{
    if (!ctx || !out || !out_len) return TREAT_B200_ERR_INVALID;
    std::string text, e;
    const bool ok = treat::CliAlign(ctx, template_path ? template_path : "", fragment_path ? fragment_path : "",
                                    s1 ? s1 : "", s2 ? s2 : "", std::string(1, (char)base), offset, &text, &e);
    if (!ok) {
        set_err(err, errlen, e);
        *out = nullptr;
        *out_len = 0;
        return TREAT_B200_ERR_INVALID;
    }
    *out = (char *)malloc(text.size() + 1);
    if (!*out) return TREAT_B200_ERR_NOMEM;
    memcpy(*out, text.data(), text.size());
    (*out)[text.size()] = 0;
    *out_len = text.size();
    return TREAT_B200_OK;
}

void treat_b200_free(void *p) { free(p); }

}  // extern "C"
