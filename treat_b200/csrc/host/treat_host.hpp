// treat_host.hpp -- host-side mirror of the reference's Go package API (package treat).
// Same names, argument meaning and error behaviour as fragment.go / template.go / align.go;
// the DP itself never runs here: Alignment values come from the GPU (treat_b200_align_batch)
// and WriteTo / SimpleAlign only expand the device's traceback ops into text.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../../include/treat_b200.h"

namespace treat {

enum OrientationType : int8_t { FORWARD = 1, REVERSE = -1 };  // const.go:20-23

// fragment.go:57-64
struct Fragment {
    std::string Name;
    uint32_t ReadCount = 0;
    double Norm = 0.0;
    std::string Bases;
    char EditBase = 'T';
    std::vector<uint32_t> EditSite;

    int Len() const { return (int)EditSite.size(); }  // fragment.go:132-134
    std::string String() const;                       // fragment.go:136-147
    std::string ToFasta() const;                      // fragment.go:123-130
};

uint32_t parseMergeCount(const std::string &recId);  // fragment.go:75-89
Fragment NewFragment(const std::string &name, const std::string &seq, OrientationType orientation, char base);  // fragment.go:91-121

struct AltRegion {  // template.go:36-39
    int Start = 0, End = 0;
};

// template.go:41-49
struct Template {
    std::string Bases;
    uint32_t EditOffset = 0;
    int EditStop = 0;
    char EditBase = 'T';
    std::vector<std::vector<uint32_t>> EditSite;
    std::vector<uint32_t> BaseIndex;
    std::vector<treat::AltRegion> AltRegion;

    void SetOffset(int offset);                        // template.go:153-161
    int Size() const { return (int)EditSite.size(); }  // template.go:163-165
    int Len() const { return (int)EditSite[0].size(); }  // template.go:167-169
    uint32_t Max(int i) const;                         // template.go:186-194
    int IndexLabel(int i) const { return i + (int)EditOffset; }  // template.go:196
    std::string String() const;                        // template.go:171-184
};

// Both return false and fill err on the reference's error paths (template.go:54,90,102,107,112).
bool NewTemplate(const Fragment &full, const Fragment &pre, const std::vector<Fragment> &alt,
                 const std::vector<AltRegion> &altRegion, Template *out, std::string *err);  // template.go:96-151
bool NewTemplateFromFasta(const std::string &path, OrientationType orientation, char base, Template *out,
                          std::string *err);  // template.go:51-94

// gofasta.SimpleParser record semantics (cmd/treat/align.go:114); batch-friendly layout.
struct FastaBatch {
    std::vector<std::string> ids;
    std::vector<uint8_t> seqs;         // concatenated
    std::vector<uint64_t> offsets;     // ids.size() + 1
    std::vector<uint32_t> read_counts; // parseMergeCount(id)
    size_t size() const { return ids.size(); }
    std::string seq(size_t i) const { return std::string(seqs.begin() + offsets[i], seqs.begin() + offsets[i + 1]); }
};
bool ReadFasta(const std::string &path, FastaBatch *out, std::string *err);
void ParseFasta(const char *bytes, size_t size, FastaBatch *out);  // the same on bytes in memory
// First '>' at or behind `from` that is the first byte of a line (every record but the first of a file starts at one);
// len if there is none; from == 0 -> 0.
uint64_t FastaNextRecordStart(const uint8_t *fasta, uint64_t len, uint64_t from);
// `parts` slices of about equal size that start at record starts: offsets[parts + 1], offsets[0] = 0, offsets[parts] = len
void FastaSplit(const uint8_t *fasta, uint64_t len, uint32_t parts, uint64_t *offsets);

// align.go:41-55; filled from a treat_b200_record
struct Alignment {
    int EditStop = 0, JuncStart = 0, JuncEnd = 0, JuncLen = 0;
    uint32_t ReadCount = 0;
    double Norm = 0.0;
    uint8_t HasMutation = 0, Mismatches = 0, Indel = 0, AltEditing = 0;
    std::string JuncSeq;
    int Score = 0;
    uint8_t Status = 0;
    std::vector<uint8_t> Ops;  // one op per alignment column (TREAT_B200_OP_*)

    static Alignment FromRecord(const treat_b200_record &rec, const uint8_t *packed_ops, const uint8_t *seq,
                                size_t seq_len);
    // align.go:388-520
    std::string WriteTo(const Fragment &frag, const Template &tmpl, int tw) const;
    // align.go:316-335
    std::vector<uint8_t> MarshalBinary() const;
    // align.go:295-314: the inverse (36-byte big-endian header + JuncSeq); false when the blob is shorter than it says
    // (the reference panics with a slice-bounds error there)
    bool UnmarshalBinary(const uint8_t *buf, size_t len);
};

// align.go:337-386 string assembly from ops (ops computed by the GPU with f1 as the template)
void SimpleAlignStrings(const Fragment &f1, const Fragment &f2, const std::vector<uint8_t> &ops, std::string *s1,
                        std::string *s2);
// cmd/treat/align.go:39-57
std::string PrintAlignment(const std::string &a1, const std::string &a2, int tw);

// Batch engine over the C ABI: the reference's per-read loop (cmd/treat/align.go:114-120) as one call.
class Aligner {
public:
    explicit Aligner(treat_b200_ctx *ctx) : ctx_(ctx) {}
    bool SetTemplate(const Template &t, std::string *err);
    // NewFragment + NewAlignment for every read of the batch
    bool AlignBatch(const FastaBatch &reads, bool excludeSnps, bool wantOps, std::vector<Alignment> *out,
                    std::string *err);
    // Alignment.SimpleAlign(f1, f2)
    bool SimpleAlign(const Fragment &f1, const Fragment &f2, std::string *s1, std::string *s2, std::string *err);

private:
    treat_b200_ctx *ctx_;
};

// `treat load` record path (cmd/treat/storage.go:758-770): Alignment.MarshalBinary for a whole batch.
// Blob i is out[out_off[i] .. out_off[i+1]); returns the total size (call with out == nullptr to size).
uint64_t MarshalBatch(const treat_b200_record *rec, const uint8_t *seqs, const uint64_t *seq_off, uint64_t N,
                      uint8_t *out, uint64_t *out_off, int threads);

// `treat mutant` (cmd/treat/mutant.go:80-109): census of the distinct gapped template strings (reads with an
// insertion) and gapped read strings (reads with a deletion), built from the device's traceback ops.
struct MutantCensus {
    std::vector<std::pair<std::string, int>> tmpl, frag;  // sorted by count (descending), then by string
    std::string Report(int n) const;                      // the text `treat mutant -n n` prints
};
MutantCensus MutantFromOps(const Template &tmpl, const treat_b200_record *rec, const uint8_t *ops, uint32_t ops_stride,
                           const uint8_t *seqs, const uint64_t *seq_off, uint64_t N);

// `treat align` (cmd/treat/align.go:59-122): returns false + err where the reference calls logrus.Fatal
// The same with the text handed to `emit` as it is produced (file mode: one call per piece of the reads file, in file
// order); emit returns false to stop.
using CliEmit = std::function<bool(const char *, size_t)>;
bool CliAlignStream(treat_b200_ctx *ctx, const std::string &templatePath, const std::string &fragmentPath,
                    const std::string &s1, const std::string &s2, const std::string &editBase, int editOffset,
                    const CliEmit &emit, std::string *err);
bool CliAlign(treat_b200_ctx *ctx, const std::string &templatePath, const std::string &fragmentPath,
              const std::string &s1, const std::string &s2, const std::string &editBase, int editOffset,
              std::string *out, std::string *err);

}  // namespace treat
