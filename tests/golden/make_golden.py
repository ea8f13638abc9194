#!/usr/bin/env python
"""Generate tests/golden/fixtures.json from the reference checkout (run in the build container).

    python tests/golden/make_golden.py [/root/reference]

/root/reference does not exist on the GPU box, so everything a test needs from it is
frozen here:
  * the reference's example FASTA inputs (examples/*.fa) -- input data only;
  * the reference's OWN expectations: the per-read attributes in the headers of
    examples/test-sample.fa (asserted by align_test.go:68-137), the tables of
    fragment_test.go:25-76 / template_test.go:26-70, and the documentation goldens
    README.rst:42-44, 70-82, 176-179;
  * DERIVED regression values (not produced by the Go binary -- there is no Go toolchain
    here): oracle/pyref.py outputs under the ULD tie order for every example read.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyref  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
EX = os.path.join(REF, "examples")

fx = {"_generated_by": "tests/golden/make_golden.py", "fasta": {}}
for name in sorted(os.listdir(EX)):
    if name.endswith(".fa"):
        fx["fasta"][name] = open(os.path.join(EX, name), "rb").read().decode("latin-1")

# --- the reference's own expectations -------------------------------------------------
exp = []
for rid, seq in pyref.read_fasta(os.path.join(EX, "test-sample.fa")):
    attr = {k: int(v) for k, v in (kv.split("=") for kv in re.findall(r"[^=\s]+=\-?\d+", rid))}  # align_test.go:54-66
    exp.append({"id": rid, "seq": seq, "attr": attr})
fx["test_sample"] = {"templates": "test-templates.fa", "base": "t", "reads": exp, "size": 3,
                     "cite": "align_test.go:68-137"}

readme = open(os.path.join(REF, "README.rst")).read().split("\n")
i0 = next(i for i, l in enumerate(readme) if "treat align -t simple-templates.fa" in l)
block = []
for l in readme[i0 + 1:]:
    if l and not l.startswith("  "):
        break
    block.append(l[2:] if l.startswith("  ") else l)
while block and block[-1] == "":
    block.pop()
fx["readme_block"] = {"templates": "simple-templates.fa", "reads": "simple-sequences.fa", "base": "T",
                      "stdout": "\n".join(block) + "\n\n", "cite": "README.rst:70-82"}
fx["readme_pairwise"] = {"s1": "ATCTGTATGT", "s2": "ATTCGATTG", "base": "T",
                         "stdout": "A-TCTGTA-TGT\nATTC-G-ATTG-\n\n", "cite": "README.rst:42-44"}
fx["readme_csv"] = {"templates": "templates.fa", "reads": "sample-1.fa", "offset": 0, "cite": "README.rst:176-179",
                    "rows": [
                        {"id": "1-10", "read_count": 10, "alt_editing": 0, "has_mutation": 0, "edit_stop": 137,
                         "junc_end": 143, "junc_len": 6, "junc_seq": "ATATAATATTTTTG"},
                        {"id": "2-9", "read_count": 9, "alt_editing": 0, "has_mutation": 0, "edit_stop": 95,
                         "junc_end": 123, "junc_len": 28,
                         "junc_seq": "TTCGGTATTTGTTTTATGTTATTATATGAGTCCGCGATTGCCCAGCTCTG"}]}
fx["fragment_test"] = {"cite": "fragment_test.go:25-50", "base": "t",
                       "seqs": ["CTGGTTTCA", "CTAATACACTTTTGAttttt", "TtTCTGGCTATT", "ttCGAGTATTT"]}
fx["parse_read_counts"] = {"cite": "fragment_test.go:52-76", "table": {
    " 132-2082": 2082, "SAMPLE1_GENE_123432_2082": 2082, "132-2082": 2082, "GENE_88772-2082": 2082,
    "791-2082": 2082, "791-2082            ": 2082, "791-2082 attr1=x attr2=y attr3=z": 2082,
    "7912082 attr1=x attr2=y attr3=z": 1, "badtest2082": 1, "2082": 1, "-2082": 1, "_2082": 1, "2082_": 1}}
fx["template_test"] = {"cite": "template_test.go:26-70", "rps12_size": 5,
                       "bad": {"full": "ttCCAATTGCAATTT", "pre": "ttCAATT"},
                       "good": {"full": "ttCCAATTGCAATTT", "pre": "ttttCCAATTTTGCAATTTTT"}}
fx["align_smoke"] = {"cite": "align_test.go:32-52", "fe": "TTTCTGAGTTTAGTAT", "pe": "TTTTTTCTTTTGAGTTTTTTAGTATT",
                     "read": "TTTTCTTTGAGTTTAGTATTT", "base": "t"}

# --- derived regression values (pyref, ULD) --------------------------------------------
def run_set(tmpl_file, reads_file, base, offset=0, exclude_snps=False):
    t = pyref.template_from_fasta(os.path.join(EX, tmpl_file), pyref.FORWARD, base)
    t.SetOffset(offset)
    out = []
    for rid, seq in pyref.read_fasta(os.path.join(EX, reads_file)):
        f = pyref.Fragment(rid, seq, pyref.FORWARD, base)
        a = pyref.NewAlignment(f, t, exclude_snps)
        aln1, aln2, score = pyref.nw_align(t.Bases, f.Bases)
        out.append({"id": rid, "fields": a.fields(), "aln1": aln1, "aln2": aln2,
                    "text": a.WriteTo(f, t, 80), "marshal_hex": a.MarshalBinary().hex()})
    return {"templates": tmpl_file, "reads": reads_file, "base": base, "offset": offset,
            "exclude_snps": exclude_snps, "results": out}

fx["derived"] = {
    "_note": "oracle/pyref.py under the ULD tie order; regression values, NOT Go output",
    "sets": [
        run_set("test-templates.fa", "test-sample.fa", "t"),
        run_set("test-templates.fa", "test-sample.fa", "t", exclude_snps=True),
        run_set("simple-templates.fa", "simple-sequences.fa", "T"),
        run_set("templates.fa", "sample-1.fa", "T"),
        run_set("templates.fa", "sample-1.fa", "T", offset=10),
        run_set("templates.fa", "clones.fa", "T"),
        run_set("templates.fa", "fastx-out.fa", "T"),
    ],
}
json.dump(fx, open(os.path.join(HERE, "fixtures.json"), "w"), indent=1, sort_keys=True)
print("wrote", os.path.join(HERE, "fixtures.json"), os.path.getsize(os.path.join(HERE, "fixtures.json")), "bytes")
