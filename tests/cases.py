"""Parity cases: the reference's own test inputs (tests/golden/simple, tests/golden/full) plus small synthetic ones.

Each case: reads (comma list, relative to tests/golden or generated), ref, k, extra `find` flags.
Expected outputs of the unmodified reference binary are committed under tests/golden/ref_outputs/
(see tests/golden/make_reference_fixtures.py for provenance).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
SYN_DIR = os.path.join(GOLD, "_synth")  # generated on demand (git-ignored), deterministic

S = "simple/"
R = "simple/references/"
CASES = {
    # /root/reference/test/simple_test.sh:65-112
    "clean_insert": dict(reads=S + "master.fasta", ref=R + "deleted.fasta", k=31, flags=["-insert-only"]),
    "inserts_ref10k": dict(reads=S + "readref10K.fasta", ref=R + "g10K_del.fasta", k=31, flags=["-insert-only"]),
    "one_snp": dict(reads=S + "master.fasta", ref=R + "sSNP.fasta", k=31, flags=["-snp-only"]),
    "multi_snp": dict(reads=S + "master.fasta", ref=R + "multiSNP.fasta", k=31, flags=["-snp-only"]),
    "snp_before_insert": dict(reads=S + "master.fasta", ref=R + "deleted_before_SNP.fasta", k=31, flags=["-no-deletion", "-homo-only"]),
    "hetero_insert": dict(reads=S + "deleted.fasta," + S + "master.fasta", ref=R + "deleted.fasta", k=31, flags=["-hete-only", "-max-rep", "2"]),
    "deletion": dict(reads=S + "deleted.fasta", ref=R + "master.fasta", k=31, flags=["-deletion-only"]),
    "fuzzy_deletion": dict(reads=S + "deletionfuzzy.fasta", ref=R + "deletionfuzzy.fasta", k=31, flags=["-deletion-only"]),
    "n_in_stretch": dict(reads=S + "master.fasta", ref=R + "n_in_stretch.fasta", k=31, flags=["-insert-only"]),
    "n_before_gap": dict(reads=S + "master.fasta", ref=R + "n_before_gap.fasta", k=31, flags=["-insert-only"]),
    "n_after_gap": dict(reads=S + "master.fasta", ref=R + "n_after_gap.fasta", k=31, flags=["-insert-only"]),
    # /root/reference/test/simple_full_test.sh:36 (bundled example; gold files are the reference's own)
    "full": dict(reads="full/reads_r1.fastq,full/reads_r2.fastq", ref="full/reference.fasta", k=31, flags=[]),
    "full_k63": dict(reads="full/reads_r1.fastq,full/reads_r2.fastq", ref="full/reference.fasta", k=63, flags=[]),
    "full_backup": dict(reads="full/reads_r1.fastq,full/reads_r2.fastq", ref="full/reference.fasta", k=31, flags=["-backup", "-branching-filter", "-1"]),
    "full_k21_amin3": dict(reads="full/reads_r1.fastq,full/reads_r2.fastq", ref="full/reference.fasta", k=21, flags=["-abundance-min", "3", "-max-rep", "8"]),
    # deterministic synthetic (tools/synth.py), all finders enabled
    "syn_tiny_k31": dict(synth="tiny", seed=20241, k=31, flags=[]),
    "syn_tiny_k63": dict(synth="tiny", seed=20241, k=63, flags=[]),
    "syn_tiny_k32": dict(synth="tiny", seed=20242, k=32, flags=[]),
    "syn_small_k31": dict(synth="small", seed=20243, k=31, flags=[]),
    "syn_small_k47_homo": dict(synth="small", seed=20244, k=47, flags=["-homo-only"]),
}


def case_paths(case, make=True):
    """Return (reads_uri, ref_path) with absolute paths; synthetic inputs are generated on first use."""
    if "synth" in case:
        d = os.path.join(SYN_DIR, "%s_%d" % (case["synth"], case["seed"]))
        reads = os.path.join(d, "r1.fq") + "," + os.path.join(d, "r2.fq")
        ref = os.path.join(d, "ref.fa")
        if make and not os.path.exists(os.path.join(d, "truth.json")):
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import synth
            synth.make_dataset(d, synth.CONFIGS[case["synth"]], case["seed"])
        return reads, ref
    reads = ",".join(os.path.join(GOLD, p) for p in case["reads"].split(","))
    return reads, os.path.join(GOLD, case["ref"])


def expected(name):
    """(breakpoints text, vcf record text, info text) produced by the reference binary."""
    base = os.path.join(GOLD, "ref_outputs", name)
    return (open(base + ".breakpoints").read(), open(base + ".vcf").read(), open(base + ".info").read())
