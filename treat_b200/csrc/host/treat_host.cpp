// treat_host.cpp -- host-side mirror of package treat (see treat_host.hpp).  No DP here.
#include "treat_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iterator>
#include <map>
#include <thread>

namespace treat {

namespace {

inline char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
inline bool isSpace(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }  // unicode.IsSpace, ASCII subset

std::string trimSpace(const std::string &s)
{
    size_t b = 0, e = s.size();
    while (b < e && isSpace(s[b])) b++;
    while (e > b && isSpace(s[e - 1])) e--;
    return s.substr(b, e - b);
}

// strconv.Atoi on a string of ASCII digits: false on int64 range error
bool atoiDigits(const std::string &d, int64_t *out)
{
    uint64_t v = 0;
    for (char c : d) {
        const uint64_t x = (uint64_t)(c - '0');
        if (v > (0x7fffffffffffffffULL - x) / 10) return false;
        v = v * 10 + x;
    }
    *out = (int64_t)v;
    return true;
}

// first match of `<key>(\d+)` in id (startPattern / endPattern of template.go:33-34)
int findAttr(const std::string &id, const std::string &key)
{
    size_t p = 0;
    while ((p = id.find(key, p)) != std::string::npos) {
        size_t d = p + key.size(), e = d;
        while (e < id.size() && id[e] >= '0' && id[e] <= '9') e++;
        if (e > d) {
            int64_t v;
            if (!atoiDigits(id.substr(d, e - d), &v)) return 0;  // Atoi error -> 0 (template.go:70-72)
            return (int)v;
        }
        p++;
    }
    return -1;
}

// writeBase, align.go:88-93
inline void writeBase(std::string &buf, char base, uint32_t count, uint32_t max)
{
    buf.append(max - count, '-');
    if (count > 0) buf.append(count, base);
}

}  // namespace

// ------------------------------------------------------------------ fragment.go

uint32_t parseMergeCount(const std::string &recId)
{
    const std::string t = trimSpace(recId);
    const std::string head = t.substr(0, t.find(' '));  // SplitN(..., " ", 2)[0]
    // `.+[_\-](\d+)$`: greedy prefix, so the separator is the last '_' or '-'
    const size_t sep = head.find_last_of("_-");
    if (sep == std::string::npos || sep == 0 || sep + 1 >= head.size()) return 1;
    const std::string digits = head.substr(sep + 1);
    if (!std::all_of(digits.begin(), digits.end(), [](char c) { return c >= '0' && c <= '9'; })) return 1;
    int64_t v;
    if (!atoiDigits(digits, &v)) return 1;
    return (uint32_t)v;
}

Fragment NewFragment(const std::string &name, const std::string &seqIn, OrientationType orientation, char base)
{
    Fragment f;
    base = upper(base);
    std::string seq(seqIn);
    if (orientation == REVERSE) std::reverse(seq.begin(), seq.end());  // reversal only, no complement
    std::transform(seq.begin(), seq.end(), seq.begin(), upper);
    uint32_t run = 0;
    for (char c : seq) {
        if (c != base) {
            f.EditSite.push_back(run);
            run = 0;
            f.Bases.push_back(c);
        } else {
            run++;
        }
    }
    f.EditSite.push_back(run);
    f.Name = name;
    f.ReadCount = parseMergeCount(name);
    f.EditBase = base;
    return f;
}

std::string Fragment::String() const
{
    std::string s;
    for (size_t i = 0; i < EditSite.size(); i++) {
        s.append(EditSite[i], EditBase);
        if (i < Bases.size()) s.push_back(Bases[i]);
    }
    return s;
}

std::string Fragment::ToFasta() const { return ">" + Name + "\n" + String(); }

// ------------------------------------------------------------------ gofasta

void ParseFasta(const char *bytes, size_t size, FastaBatch *out)
{
    const std::string_view data(bytes, size);
    out->ids.clear();
    out->seqs.clear();
    out->offsets.assign(1, 0);
    out->read_counts.clear();
    out->seqs.reserve(data.size());
    size_t pos = data.find('>');  // bytes before the first record are skipped
    bool open = false;
    while (pos != std::string_view::npos && pos < data.size()) {
        size_t eol = data.find('\n', pos);
        if (eol == std::string_view::npos) eol = data.size();
        size_t le = eol;
        if (le > pos && data[le - 1] == '\r') le--;
        if (data[pos] == '>') {
            if (open) out->offsets.push_back(out->seqs.size());
            out->ids.push_back(trimSpace(std::string(data.substr(pos + 1, le - pos - 1))));
            open = true;
        } else if (open) {
            out->seqs.insert(out->seqs.end(), data.begin() + pos, data.begin() + le);
        }
        pos = eol + 1;
    }
    if (open) out->offsets.push_back(out->seqs.size());
    out->read_counts.reserve(out->ids.size());
    for (const auto &id : out->ids) out->read_counts.push_back(parseMergeCount(id));
}

bool ReadFasta(const std::string &path, FastaBatch *out, std::string *err)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        if (err) *err = "open " + path + ": no such file or directory";
        return false;
    }
    std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    ParseFasta(data.data(), data.size(), out);
    return true;
}

uint64_t FastaNextRecordStart(const uint8_t *fasta, uint64_t len, uint64_t from)
{
    if (from == 0) return 0;  // the head of the file belongs to the first slice whatever it holds
    const uint8_t *p = fasta + from, *end = fasta + len;
    while (p < end) {
        const uint8_t *g = (const uint8_t *)memchr(p, '>', (size_t)(end - p));
        if (!g) break;
        if (g[-1] == '\n') return (uint64_t)(g - fasta);
        p = g + 1;
    }
    return len;
}

void FastaSplit(const uint8_t *fasta, uint64_t len, uint32_t parts, uint64_t *offsets)
{
    offsets[0] = 0;
    for (uint32_t k = 1; k < parts; k++) {
        const uint64_t nominal = (uint64_t)((unsigned __int128)len * k / parts);
        offsets[k] = FastaNextRecordStart(fasta, len, std::max(nominal, offsets[k - 1]));
    }
    offsets[parts] = len;
}

// ------------------------------------------------------------------ template.go

bool NewTemplate(const Fragment &full, const Fragment &pre, const std::vector<Fragment> &alt,
                 const std::vector<AltRegion> &altRegion, Template *out, std::string *err)
{
    if (full.Bases != pre.Bases || full.EditBase != pre.EditBase) {
        if (err) *err = "Invalid template sequences. Full and Pre templates must have the same non-edit bases";
        return false;
    }
    for (const auto &a : alt) {
        if (full.Bases != a.Bases || full.EditBase != a.EditBase) {
            if (err) *err = "Invalid alt template sequence. All templates must have the same non-edit bases";
            return false;
        }
    }
    if (alt.size() != altRegion.size()) {
        if (err) *err = "Invalid alt templates. Please specify the alt regions";
        return false;
    }
    Template t;
    t.Bases = full.Bases;
    t.EditBase = full.EditBase;
    t.EditSite.push_back(full.EditSite);
    t.EditSite.push_back(pre.EditSite);
    for (const auto &a : alt) t.EditSite.push_back(a.EditSite);
    t.AltRegion = altRegion;
    const int n = (int)t.EditSite[0].size();
    uint32_t index = 0;
    for (int i = 0; i < n; i++) {
        index += std::max(t.EditSite[0][i], t.EditSite[1][i]);
        if (i > 0 && i != n - 1) index++;
        t.BaseIndex.push_back(index);
    }
    t.EditStop = n - 1;
    for (int j = n - 1; j >= 0; j--) {
        if (t.EditSite[0][j] != t.EditSite[1][j]) {
            t.EditStop = (n - 1) - j;
            break;
        }
    }
    t.EditStop--;
    *out = std::move(t);
    return true;
}

bool NewTemplateFromFasta(const std::string &path, OrientationType orientation, char base, Template *out,
                          std::string *err)
{
    FastaBatch fa;
    std::string e;
    if (!ReadFasta(path, &fa, &e)) {
        if (err) *err = "Invalid FASTA file: " + e;
        return false;
    }
    std::vector<Fragment> t;
    std::vector<AltRegion> alt;
    for (size_t i = 0; i < fa.size(); i++) {
        t.push_back(NewFragment(fa.ids[i], fa.seq(i), orientation, base));
        if (t.size() > 2) {
            const int start = findAttr(fa.ids[i], "alt_start="), end = findAttr(fa.ids[i], "alt_stop=");
            if (start > -1 && end > -1) alt.push_back(AltRegion{start, end});
        }
    }
    if (t.size() < 2) {
        if (err) *err = "Must provide at least 2 templates. Full and Pre edited";
        return false;
    }
    return NewTemplate(t[0], t[1], std::vector<Fragment>(t.begin() + 2, t.end()), alt, out, err);
}

void Template::SetOffset(int offset)
{
    EditOffset = (uint32_t)offset;
    EditStop += offset;
    for (auto &r : AltRegion) {
        r.Start -= offset;
        r.End -= offset;
    }
}

uint32_t Template::Max(int i) const
{
    uint32_t mx = 0;
    for (const auto &row : EditSite) mx = std::max(mx, row[i]);
    return mx;
}

std::string Template::String() const
{
    std::string s;
    for (size_t i = 0; i < EditSite[0].size(); i++) {
        s.append(EditSite[0][i], EditBase);
        if (i < Bases.size()) s.push_back(Bases[i]);
    }
    return s;
}

// ------------------------------------------------------------------ align.go (text side)

Alignment Alignment::FromRecord(const treat_b200_record &r, const uint8_t *packed_ops, const uint8_t *seq,
                                size_t seq_len)
{
    Alignment a;
    a.EditStop = r.edit_stop;
    a.JuncStart = r.junc_start;
    a.JuncEnd = r.junc_end;
    a.JuncLen = r.junc_len;
    a.ReadCount = r.read_count;
    a.HasMutation = r.has_mutation;
    a.AltEditing = r.alt_editing;
    a.Mismatches = r.mismatches;
    a.Indel = r.indel;
    a.Score = r.score;
    a.Status = r.status;
    if (seq && r.junc_seq_len && (size_t)r.junc_seq_off + r.junc_seq_len <= seq_len) {
        a.JuncSeq.assign((const char *)seq + r.junc_seq_off, r.junc_seq_len);
        std::transform(a.JuncSeq.begin(), a.JuncSeq.end(), a.JuncSeq.begin(), upper);
    }
    if (packed_ops) {
        // a read without a gap aligns base over base: its row is all zero and, with TREAT_B200_OPS_GAPPED_ONLY, not even written
        a.Ops.assign(r.aln_len, (uint8_t)TREAT_B200_OP_MATCH);
        if (r.indel)
            for (uint32_t c = 0; c < r.aln_len; c++) a.Ops[c] = (packed_ops[c >> 2] >> (2 * (c & 3))) & 3u;
    }
    return a;
}

// ops must consume exactly the template's and the fragment's bases
static bool opsConsistent(const std::vector<uint8_t> &ops, size_t m, size_t n)
{
    size_t ti = 0, fi = 0;
    for (uint8_t op : ops) {
        if (op > TREAT_B200_OP_INS) return false;
        if (op != TREAT_B200_OP_INS) ti++;
        if (op != TREAT_B200_OP_DEL) fi++;
    }
    return ti == m && fi == n;
}

std::string Alignment::WriteTo(const Fragment &frag, const Template &tmpl, int tw) const
{
    if (tw <= 0) tw = 80;
    if (!opsConsistent(Ops, tmpl.Bases.size(), frag.Bases.size())) return std::string();  // corrupt input: no text
    const int K = tmpl.Size();
    std::vector<std::string> buf(K + 1);
    std::string &rd = buf[K];
    int fi = 0, ti = 0;
    for (uint8_t op : Ops) {
        if (op == TREAT_B200_OP_INS) {  // aln1[ai] == '-'
            for (int k = 0; k < K; k++) {
                writeBase(buf[k], tmpl.EditBase, 0, frag.EditSite[fi]);
                buf[k].push_back('-');
            }
            writeBase(rd, frag.EditBase, frag.EditSite[fi], frag.EditSite[fi]);
            rd.push_back(frag.Bases[fi]);
            fi++;
        } else if (op == TREAT_B200_OP_DEL) {  // aln2[ai] == '-'
            const uint32_t mx = tmpl.Max(ti);
            for (int k = 0; k < K; k++) {
                writeBase(buf[k], tmpl.EditBase, tmpl.EditSite[k][ti], mx);
                buf[k].push_back(tmpl.Bases[ti]);
            }
            writeBase(rd, '-', 0, mx);
            rd.push_back('-');
            ti++;
        } else {
            const uint32_t mx = std::max(tmpl.Max(ti), frag.EditSite[fi]);
            for (int k = 0; k < K; k++) {
                writeBase(buf[k], tmpl.EditBase, tmpl.EditSite[k][ti], mx);
                buf[k].push_back(tmpl.Bases[ti]);
            }
            writeBase(rd, frag.EditBase, frag.EditSite[fi], mx);
            rd.push_back(frag.Bases[fi]);
            fi++;
            ti++;
        }
    }
    const uint32_t mx = std::max(tmpl.Max(ti), frag.EditSite[fi]);  // last edit site: edit bases only
    for (int k = 0; k < K; k++) writeBase(buf[k], tmpl.EditBase, tmpl.EditSite[k][ti], mx);
    writeBase(rd, frag.EditBase, frag.EditSite[fi], mx);

    const size_t cols = (size_t)(tw - 4);
    size_t rows = buf[0].size() / cols;
    if (buf[0].size() % cols > 0) rows++;
    std::vector<std::string> labels = {"FE", "PE"};
    for (size_t i = 0; i < tmpl.AltRegion.size(); i++) labels.push_back("A" + std::to_string(i + 1));
    labels.push_back("RD");

    std::string out;
    const std::string bar((size_t)tw, '=');
    out += bar + "\n" + frag.Name + "\n" + bar + "\n";
    char line[128];
    snprintf(line, sizeof line, "JSS: %d\nESS: %d\nJES: %d\nJunc Len: %d\n", JuncStart, EditStop, JuncEnd, JuncLen);
    out += line;
    if (AltEditing > 0) {
        snprintf(line, sizeof line, "Alt: %d\n", (int)AltEditing);
        out += line;
    }
    out += bar + "\n\n";
    for (size_t r = 0; r < rows; r++) {
        for (size_t i = 0; i < buf.size(); i++) {
            out += labels[i] + ": ";
            const size_t lo = r * cols, hi = std::min(lo + cols, buf[i].size());
            if (hi > lo) out.append(buf[i], lo, hi - lo);
            out.push_back('\n');
        }
        out.push_back('\n');
    }
    return out;
}

bool Alignment::UnmarshalBinary(const uint8_t *buf, size_t len)
{
    if (!buf || len < 36) return false;
    auto be32 = [&](size_t o) { return ((uint32_t)buf[o] << 24) | ((uint32_t)buf[o + 1] << 16) | ((uint32_t)buf[o + 2] << 8) | (uint32_t)buf[o + 3]; };
    const uint32_t seq_len = be32(32);
    if ((size_t)seq_len > len - 36) return false;
    EditStop = (int)(int32_t)be32(0);  // readInt64 sign-extends 32 bits (align.go:281-287)
    JuncStart = (int)(int32_t)be32(4);
    JuncEnd = (int)(int32_t)be32(8);
    JuncLen = (int)(int32_t)be32(12);
    ReadCount = be32(16);
    HasMutation = buf[20];
    AltEditing = buf[21];
    Mismatches = buf[22];
    Indel = buf[23];
    const uint64_t bits = ((uint64_t)be32(24) << 32) | be32(28);
    memcpy(&Norm, &bits, 8);
    JuncSeq.assign(reinterpret_cast<const char *>(buf) + 36, seq_len);
    return true;
}

std::vector<uint8_t> Alignment::MarshalBinary() const
{
    std::vector<uint8_t> b(36 + JuncSeq.size());
    auto be32 = [&](size_t o, uint32_t v) {
        b[o] = (uint8_t)(v >> 24);
        b[o + 1] = (uint8_t)(v >> 16);
        b[o + 2] = (uint8_t)(v >> 8);
        b[o + 3] = (uint8_t)v;
    };
    be32(0, (uint32_t)EditStop);  // writeInt64 keeps the low 32 bits (align.go:288-293)
    be32(4, (uint32_t)JuncStart);
    be32(8, (uint32_t)JuncEnd);
    be32(12, (uint32_t)JuncLen);
    be32(16, ReadCount);
    b[20] = HasMutation;
    b[21] = AltEditing;
    b[22] = Mismatches;
    b[23] = Indel;
    uint64_t bits;
    memcpy(&bits, &Norm, 8);
    be32(24, (uint32_t)(bits >> 32));
    be32(28, (uint32_t)bits);
    be32(32, (uint32_t)JuncSeq.size());
    std::copy(JuncSeq.begin(), JuncSeq.end(), b.begin() + 36);
    return b;
}

void SimpleAlignStrings(const Fragment &f1, const Fragment &f2, const std::vector<uint8_t> &ops, std::string *s1,
                        std::string *s2)
{
    std::string &b0 = *s1, &b1 = *s2;
    b0.clear();
    b1.clear();
    if (!opsConsistent(ops, f1.Bases.size(), f2.Bases.size())) return;
    int fi = 0, ti = 0;
    for (uint8_t op : ops) {
        if (op == TREAT_B200_OP_INS) {
            writeBase(b0, f1.EditBase, 0, f2.EditSite[fi]);
            b0.push_back('-');
            writeBase(b1, f2.EditBase, f2.EditSite[fi], f2.EditSite[fi]);
            b1.push_back(f2.Bases[fi]);
            fi++;
        } else if (op == TREAT_B200_OP_DEL) {
            writeBase(b0, f1.EditBase, f1.EditSite[ti], f1.EditSite[ti]);
            b0.push_back(f1.Bases[ti]);
            writeBase(b1, '-', 0, f1.EditSite[ti]);
            b1.push_back('-');
            ti++;
        } else {
            const uint32_t mx = std::max(f1.EditSite[ti], f2.EditSite[fi]);
            writeBase(b0, f1.EditBase, f1.EditSite[ti], mx);
            b0.push_back(f1.Bases[ti]);
            writeBase(b1, f2.EditBase, f2.EditSite[fi], mx);
            b1.push_back(f2.Bases[fi]);
            fi++;
            ti++;
        }
    }
    const uint32_t mx = std::max(f1.EditSite[ti], f2.EditSite[fi]);
    writeBase(b0, f1.EditBase, f1.EditSite[ti], mx);
    writeBase(b1, f2.EditBase, f2.EditSite[fi], mx);
}

std::string PrintAlignment(const std::string &a1, const std::string &a2, int tw)
{
    std::string out;
    const size_t n = a1.size(), cols = (size_t)tw;
    for (size_t lo = 0; lo < n; lo += cols) {
        const size_t hi = std::min(lo + cols, n);
        out.append(a1, lo, hi - lo);
        out.push_back('\n');
        out.append(a2, lo, hi - lo);
        out += "\n\n";
    }
    return out;
}

// ------------------------------------------------------------------ treat load / treat mutant helpers

uint64_t MarshalBatch(const treat_b200_record *rec, const uint8_t *seqs, const uint64_t *seq_off, uint64_t N,
                      uint8_t *out, uint64_t *out_off, int threads)
{
    uint64_t total = 0;
    for (uint64_t i = 0; i < N; i++) {
        if (out_off) out_off[i] = total;
        total += 36 + rec[i].junc_seq_len;
    }
    if (out_off) out_off[N] = total;
    if (!out || !out_off) return total;
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const treat_b200_record &r = rec[i];
            uint8_t *b = out + out_off[i];
            auto be32 = [&](size_t o, uint32_t v) {
                b[o] = (uint8_t)(v >> 24);
                b[o + 1] = (uint8_t)(v >> 16);
                b[o + 2] = (uint8_t)(v >> 8);
                b[o + 3] = (uint8_t)v;
            };
            be32(0, (uint32_t)r.edit_stop);
            be32(4, (uint32_t)r.junc_start);
            be32(8, (uint32_t)r.junc_end);
            be32(12, (uint32_t)r.junc_len);
            be32(16, r.read_count);
            b[20] = r.has_mutation;
            b[21] = r.alt_editing;
            b[22] = r.mismatches;
            b[23] = r.indel;
            memset(b + 24, 0, 8);  // Norm = 0.0 on this path
            be32(32, r.junc_seq_len);
            const uint8_t *src = seqs + seq_off[i] + r.junc_seq_off;
            for (uint32_t k = 0; k < r.junc_seq_len; k++) b[36 + k] = (uint8_t)upper((char)src[k]);
        }
    };
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > N) threads = N ? (int)N : 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work, N * t / threads, N * (t + 1) / threads);
    work(0, N / threads);
    for (auto &th : pool) th.join();
    return total;
}

MutantCensus MutantFromOps(const Template &tmpl, const treat_b200_record *rec, const uint8_t *ops, uint32_t ops_stride,
                           const uint8_t *seqs, const uint64_t *seq_off, uint64_t N)
{
    std::map<std::string, int> tm, fm;
    const char eb = tmpl.EditBase;
    std::string aln1, aln2, bases;
    for (uint64_t i = 0; i < N; i++) {
        const treat_b200_record &r = rec[i];
        if (!r.indel) continue;  // no '-' in either gapped string
        // Fragment.Bases of this read (upper-cased, edit base stripped)
        bases.clear();
        for (uint64_t p = seq_off[i]; p < seq_off[i + 1]; p++) {
            const char c = upper((char)seqs[p]);
            if (c != eb) bases.push_back(c);
        }
        const uint8_t *o = ops + i * (uint64_t)ops_stride;
        aln1.clear();
        aln2.clear();
        size_t ti = 0, fi = 0;
        bool gap1 = false, gap2 = false, ok = true;
        for (uint32_t c = 0; c < r.aln_len && ok; c++) {
            const uint8_t op = (o[c >> 2] >> (2 * (c & 3))) & 3u;
            if (op == TREAT_B200_OP_INS) {
                ok = fi < bases.size();
                aln1.push_back('-');
                if (ok) aln2.push_back(bases[fi++]);
                gap1 = true;
            } else if (op == TREAT_B200_OP_DEL) {
                ok = ti < tmpl.Bases.size();
                if (ok) aln1.push_back(tmpl.Bases[ti++]);
                aln2.push_back('-');
                gap2 = true;
            } else {
                ok = ti < tmpl.Bases.size() && fi < bases.size();
                if (ok) {
                    aln1.push_back(tmpl.Bases[ti++]);
                    aln2.push_back(bases[fi++]);
                }
            }
        }
        if (!ok) continue;  // ops do not belong to this read: skip rather than read out of bounds
        if (gap1) tm[aln1]++;
        if (gap2) fm[aln2]++;
    }
    MutantCensus out;
    auto rank = [](const std::map<std::string, int> &m, std::vector<std::pair<std::string, int>> *v) {
        v->assign(m.begin(), m.end());
        std::stable_sort(v->begin(), v->end(), [](const auto &a, const auto &b) { return a.second > b.second; });
    };
    rank(tm, &out.tmpl);
    rank(fm, &out.frag);
    return out;
}

std::string MutantCensus::Report(int n) const
{
    // cmd/treat/mutant.go:91-109: the loops print before they test count > n, so n+1 lines at most
    std::string s = "Template Indels\n";
    int count = 0;
    for (const auto &p : tmpl) {
        s += std::to_string(p.second) + ": " + p.first + "\n";
        if (++count > n) break;
    }
    s += "\nFragment Indels\n";
    count = 0;
    for (const auto &p : frag) {
        s += std::to_string(p.second) + ": " + p.first + "\n";
        if (++count > n) break;
    }
    return s;
}

// ------------------------------------------------------------------ batch engine

bool Aligner::SetTemplate(const Template &t, std::string *err)
{
    int rc = treat_b200_use_template(ctx_, reinterpret_cast<const treat_b200_template *>(&t));
    if (rc != TREAT_B200_OK) {
        if (err) *err = std::string(treat_b200_strerror(rc)) + ": " + treat_b200_last_error(ctx_);
        return false;
    }
    return true;
}

bool Aligner::AlignBatch(const FastaBatch &reads, bool excludeSnps, bool wantOps, std::vector<Alignment> *out,
                         std::string *err)
{
    const uint64_t N = reads.size();
    out->clear();
    if (N == 0) return true;
    uint32_t max_len = 0;
    for (uint64_t i = 0; i < N; i++) max_len = std::max<uint32_t>(max_len, (uint32_t)(reads.offsets[i + 1] - reads.offsets[i]));
    const uint32_t stride = treat_b200_ops_stride(ctx_, max_len);
    std::vector<treat_b200_record> rec(N);
    std::vector<uint8_t> ops(wantOps ? (size_t)N * stride : 0);
    static const uint8_t kEmpty = 0;
    // FromRecord never reads the (all-zero) ops row of a read without a gap, so only gapped reads' rows need to come back
    const uint32_t flags = (excludeSnps ? TREAT_B200_EXCLUDE_SNPS : 0) | (wantOps ? TREAT_B200_WANT_OPS | TREAT_B200_OPS_GAPPED_ONLY : 0);
    int rc = treat_b200_align_batch(ctx_, reads.seqs.empty() ? &kEmpty : reads.seqs.data(), reads.offsets.data(),
                                    reads.read_counts.data(), N, flags, rec.data(), wantOps ? ops.data() : nullptr, stride);
    if (rc != TREAT_B200_OK) {
        if (err) *err = std::string(treat_b200_strerror(rc)) + ": " + treat_b200_last_error(ctx_);
        return false;
    }
    out->reserve(N);
    for (uint64_t i = 0; i < N; i++)
        out->push_back(Alignment::FromRecord(rec[i], wantOps ? ops.data() + i * stride : nullptr,
                                             reads.seqs.data() + reads.offsets[i], reads.offsets[i + 1] - reads.offsets[i]));
    return true;
}

bool Aligner::SimpleAlign(const Fragment &f1, const Fragment &f2, std::string *s1, std::string *s2, std::string *err)
{
    // nwalgo.Align(f1.Bases, f2.Bases, 1, -1, -1) on the GPU: f1 plays the template (FE = PE = f1)
    if (f1.Bases.empty()) {  // nothing to align against: every read base is an insertion
        SimpleAlignStrings(f1, f2, std::vector<uint8_t>(f2.Bases.size(), TREAT_B200_OP_INS), s1, s2);
        return true;
    }
    Template t;
    std::string e;
    if (!NewTemplate(f1, f1, {}, {}, &t, &e)) {
        if (err) *err = e;
        return false;
    }
    if (!SetTemplate(t, err)) return false;
    FastaBatch b;
    b.ids.push_back(f2.Name);
    const std::string raw = f2.String();
    b.seqs.assign(raw.begin(), raw.end());
    b.offsets = {0, (uint64_t)raw.size()};
    b.read_counts.push_back(f2.ReadCount);
    const uint32_t stride = treat_b200_ops_stride(ctx_, (uint32_t)raw.size());
    treat_b200_record rec;
    std::vector<uint8_t> ops(stride);
    static const uint8_t kEmpty = 0;
    int rc = treat_b200_align_batch(ctx_, b.seqs.empty() ? &kEmpty : b.seqs.data(), b.offsets.data(), nullptr, 1,
                                    TREAT_B200_WANT_OPS | TREAT_B200_NO_SUMMARY, &rec, ops.data(), stride);
    if (rc != TREAT_B200_OK) {
        if (err) *err = std::string(treat_b200_strerror(rc)) + ": " + treat_b200_last_error(ctx_);
        return false;
    }
    Alignment a = Alignment::FromRecord(rec, ops.data(), nullptr, 0);
    // the edit base of f2 may differ from f1's only through the caller; the device stripped
    // f2.String() with f1's edit base, which is the same rune in every reference call site
    SimpleAlignStrings(f1, f2, a.Ops, s1, s2);
    return true;
}

static int cli_text_sink(void *user, const uint8_t *text, uint64_t len, uint64_t, uint64_t)
{
    const CliEmit &emit = *static_cast<const CliEmit *>(user);
    return emit(reinterpret_cast<const char *>(text), (size_t)len) ? 0 : 1;
}

bool CliAlignStream(treat_b200_ctx *ctx, const std::string &templatePath, const std::string &fragmentPath,
                    const std::string &s1, const std::string &s2, const std::string &editBase, int editOffset,
                    const CliEmit &emit, std::string *err)
{
    auto put = [&](const std::string &t) { return emit(t.data(), t.size()); };
    if (editBase.size() != 1) {
        *err = "Please provide the edit base";
        return false;
    }
    const char base = editBase[0];
    if (s1.empty() != s2.empty()) {
        *err = "Please provide 2 fragments to align";
        return false;
    }
    Aligner eng(ctx);
    if (!s1.empty() && !s2.empty()) {
        Fragment f1 = NewFragment("1-1", s1, FORWARD, base), f2 = NewFragment("2-1", s2, FORWARD, base);
        std::string a1, a2;
        if (!eng.SimpleAlign(f1, f2, &a1, &a2, err)) return false;
        return put(PrintAlignment(a1, a2, 80));
    }
    if (fragmentPath.empty()) {
        *err = "Please provide either 2 sequences to align or a path to fragment FASTA file";
        return false;
    }
    if (templatePath.empty()) {  // the first two records of the file against each other (cmd/treat/align.go:95-111)
        FastaBatch reads;
        if (!ReadFasta(fragmentPath, &reads, err)) return false;
        if (reads.size() < 2) {
            *err = "Need 2 fragments to align";
            return false;
        }
        Fragment f1 = NewFragment(reads.ids[0], reads.seq(0), FORWARD, base);
        Fragment f2 = NewFragment(reads.ids[1], reads.seq(1), FORWARD, base);
        std::string a1, a2;
        if (!eng.SimpleAlign(f1, f2, &a1, &a2, err)) return false;
        return put(PrintAlignment(a1, a2, 80));
    }
    // every record of the file against the template (cmd/treat/align.go:113-121): FASTA parse, NewFragment, NewAlignment
    // and WriteTo all run on the device; the text arrives piece by piece in file order
    Template tmpl;
    if (!NewTemplateFromFasta(templatePath, FORWARD, base, &tmpl, err)) return false;
    tmpl.SetOffset(editOffset);
    // the file's bytes go straight into pinned memory (the H2D copies then run by DMA beside the text coming back)
    FILE *fp = fopen(fragmentPath.c_str(), "rb");
    if (!fp) {
        *err = "open " + fragmentPath + ": no such file or directory";
        return false;
    }
    fseek(fp, 0, SEEK_END);
    const long fsize = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    void *pinned = nullptr;
    std::string fallback;
    const uint8_t *data = nullptr;
    size_t size = 0;
    if (fsize > 0 && treat_b200_host_alloc(&pinned, (size_t)fsize) == TREAT_B200_OK) {
        size = fread(pinned, 1, (size_t)fsize, fp);
        data = static_cast<const uint8_t *>(pinned);
    } else {  // not a regular file (a pipe), or no pinned memory: read it the plain way
        char buf[1 << 16];
        size_t k;
        while ((k = fread(buf, 1, sizeof buf, fp)) > 0) fallback.append(buf, k);
        data = reinterpret_cast<const uint8_t *>(fallback.data());
        size = fallback.size();
    }
    fclose(fp);
    int rc = TREAT_B200_OK;
    if (!eng.SetTemplate(tmpl, err)) rc = TREAT_B200_ERR_TEMPLATE;
    else {
        rc = treat_b200_fasta_text(ctx, data, size, TREAT_B200_NO_SUMMARY, 80, cli_text_sink, const_cast<CliEmit *>(&emit), nullptr);
        if (rc != TREAT_B200_OK) *err = std::string(treat_b200_strerror(rc)) + ": " + treat_b200_last_error(ctx);
    }
    if (pinned) treat_b200_host_free(pinned);
    return rc == TREAT_B200_OK;
}


bool CliAlign(treat_b200_ctx *ctx, const std::string &templatePath, const std::string &fragmentPath,
              const std::string &s1, const std::string &s2, const std::string &editBase, int editOffset,
              std::string *out, std::string *err)
{
    out->clear();
    return CliAlignStream(ctx, templatePath, fragmentPath, s1, s2, editBase, editOffset,
                          [out](const char *p, size_t n) {
                              out->append(p, n);
                              return true;
                          },
                          err);
}

}  // namespace treat
