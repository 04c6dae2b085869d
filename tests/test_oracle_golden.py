"""Pins the ORACLE (oracle/) against the reference's own golden vectors and against outputs of the unmodified
reference binary (tests/golden/ref_outputs, provenance in tests/golden/make_reference_fixtures.py)."""
import os
import re
import subprocess

import numpy as np
import pytest

from tests import oracle_py
from tests.cases import CASES, GOLD, case_paths, expected


def run_oracle(oracle_bin, name, tmp_path, extra=()):
    case = CASES[name]
    reads, ref = case_paths(case)
    out = str(tmp_path / name)
    r = subprocess.run([oracle_bin, "find", "-in", reads, "-ref", ref, "-kmer-size", str(case["k"]), "-out", out]
                       + case["flags"] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    info = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())
    bk = open(out + ".breakpoints").read()
    vcf = "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#"))
    return bk, vcf, info


def test_bundled_example_equals_reference_gold_files(oracle_bin, tmp_path):
    """/root/reference/test/simple_full_test.sh:36-76 with the reference's own gold files (byte-exact, stricter
    than the script, which ignores header lines and VCF INFO columns)."""
    bk, vcf, info = run_oracle(oracle_bin, "full", tmp_path)
    assert bk == open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    gold_vcf = "".join(l for l in open(os.path.join(GOLD, "full", "gold.othervariants.vcf")) if not l.startswith("#"))
    assert vcf == gold_vcf
    gold_out = open(os.path.join(GOLD, "full", "gold_find.output")).read()
    assert int(re.search(r"abundance_min \(auto inferred\)\s*:\s*(\d+)", gold_out).group(1)) == int(info["cutoff_auto"]) == 7
    assert int(re.search(r"nb_solid_kmers\s*:\s*(\d+)", gold_out).group(1)) == int(info["nb_solid"]) == 7419
    for key, pat in [("homo_clean", r"clean\s*:\s*(\d+)"), ("snps", r"SNPs\s*:\s*(\d+)"), ("deletions", r"deletions\s*:\s*(\d+)")]:
        assert int(re.search(pat, gold_out).group(1)) == int(info[key])


def test_bundled_example_branching_nodes_equal_reference_gold_output(oracle):
    """`nb_branching_nodes : 36` of the reference's own gold_find.output (BranchingAlgorithm on the bundled example)."""
    case = CASES["full"]
    reads, _ = case_paths(case)
    stream = b"\n".join(s for _, s in oracle_py.read_sequences(reads)) + b"\n"
    o = oracle_py.count_stream(stream, 31, abundance_min=-1, nthreads=2)
    g = oracle_py.Graph(o["lo"], o["hi"], 31)
    nb, topo, lo, hi = g.branching()
    gold_out = open(os.path.join(GOLD, "full", "gold_find.output")).read()
    assert nb == int(re.search(r"nb_branching_nodes\s*:\s*(\d+)", gold_out).group(1)) == 36
    assert int(topo.sum()) == nb and topo[1, 1] == 0 and (np.diff(lo.astype(np.int64)) > 0).all()
    g.close()


def test_bundled_example_with_bed_equals_reference_gold_files(oracle_bin, tmp_path):
    """/root/reference/test/simple_full_test.sh:79-118: `find -bed gold.bed` against the reference's own gold_bed files."""
    bed = os.path.join(GOLD, "full_bed", "gold.bed")
    bk, vcf, info = run_oracle(oracle_bin, "full", tmp_path, extra=["-bed", bed])
    assert bk == open(os.path.join(GOLD, "full_bed", "gold_bed.breakpoints")).read()
    gold_vcf = "".join(l for l in open(os.path.join(GOLD, "full_bed", "gold_bed.othervariants.vcf")) if not l.startswith("#"))
    assert vcf == gold_vcf


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_reference_binary_outputs(oracle_bin, tmp_path, name):
    bk, vcf, info = run_oracle(oracle_bin, name, tmp_path)
    ebk, evcf, einfo = expected(name)
    assert bk == ebk
    assert vcf == evcf
    m = re.search(r"nb_solid_kmers\s*:\s*(\d+)", einfo)
    assert int(m.group(1)) == int(info["nb_solid"])
    m = re.search(r"abundance_min \(used\)\s*:\s*(\d+)", einfo)
    assert int(m.group(1)) == int(info["abundance_min_used"])


def test_kmer_values_TestKmer_kat(oracle):
    """gatb-core test/unit/src/kmer/TestKmer.cpp:145-152: "CATTGATAGTGG", k=3, direct k-mers."""
    f_lo, _, c_lo, _, valid = oracle_py.kmers(b"CATTGATAGTGG", 3)
    assert f_lo.tolist() == [18, 10, 43, 44, 50, 8, 35, 14, 59, 47]
    assert valid.all()
    # canonical = min(fwd, revcomp) and revcomp is an involution
    for v in f_lo.tolist():
        import ctypes as C
        lo, hi = C.c_uint64(), C.c_uint64()
        oracle.mtgo_revcomp(v, 0, 3, C.byref(lo), C.byref(hi))
        lo2, hi2 = C.c_uint64(), C.c_uint64()
        oracle.mtgo_revcomp(lo.value, 0, 3, C.byref(lo2), C.byref(hi2))
        assert lo2.value == v


def test_kmer_badchar(oracle):
    """TestKmer.cpp kmer_badchar (:542): an N invalidates the k windows that contain it, nothing else."""
    seq = b"ACGTACGTNACGTACGTACGT"
    k = 5
    *_, valid = oracle_py.kmers(seq, k)
    exp = [int(b"N" not in seq[i:i + k]) for i in range(len(seq) - k + 1)]
    assert valid.tolist() == exp


def test_minimizer_TestKmer_kat(oracle):
    """TestKmer.cpp:390-440 (kmer_minimizer2, ModelDirect LUT): k=15, m=7."""
    table = [("ATGTCTGAAGTGACC", "AAGTGAC"), ("TGTCTGAAGTGACCT", "AAGTGAC"), ("GTCTGAAGTGACCTA", "AAGTGAC"),
             ("TCTGAAGTGACCTAA", "AAGTGAC"), ("CTGAAGTGACCTAAC", "AAGTGAC"), ("TGAAGTGACCTAACA", "AAGTGAC"),
             ("GAAGTGACCTAACAT", "AAGTGAC"), ("AAGTGACCTAACATT", "AAGTGAC"), ("AGTGACCTAACATTG", "AACATTG"),
             ("GTGACCTAACATTGC", "AACATTG"), ("TGACCTAACATTGCA", "AACATTG")]
    code = {"A": 0, "C": 1, "T": 2, "G": 3}

    def val(s):
        v = 0
        for ch in s:
            v = v * 4 + code[ch]
        return v
    for kmer, mini in table:
        assert oracle.mtgo_minimizer(val(kmer), 0, 15, 7, 0) == val(mini)


SEQS4 = None


def _dsk_seqs():
    return open(os.path.join(GOLD, "dsk_check1_seqs.txt")).read().split()


@pytest.mark.parametrize("k,nks,expect", [(9, 1, 2540), (9, 2, 151), (9, 3, 18), (9, 4, 3), (9, 5, 2), (9, 6, 0),
                                          (11, 1, 2667), (11, 2, 41), (11, 3, 0), (13, 1, 2690), (13, 2, 12), (13, 3, 0),
                                          (15, 1, 2691), (15, 2, 5), (15, 3, 0)])
def test_dsk_check1_solid_counts(oracle, k, nks, expect):
    """gatb-core test/unit/src/kmer/TestDSK.cpp:147-243 (numbers "computed with the original minia")."""
    stream = "\n".join(_dsk_seqs()).encode()
    r = oracle_py.count_stream(stream, k, abundance_min=nks)
    assert len(r["lo"]) == expect


@pytest.mark.parametrize("n,k,nks,expect", [(1, 27, 1, 1), (1, 26, 1, 2), (1, 27, 2, 0), (2, 27, 2, 1), (2, 26, 2, 2),
                                            (2, 27, 3, 0), (3, 27, 3, 1), (3, 26, 3, 2), (3, 26, 4, 0)])
def test_dsk_check1_small(oracle, n, k, nks, expect):
    s1 = b"GATCCTCCCCAGGCCCCTACACCCAAT"
    r = oracle_py.count_stream(b"\n".join([s1] * n), k, abundance_min=nks)
    assert len(r["lo"]) == expect


def test_dsk_check2_solid_values(oracle):
    """TestDSK.cpp:245-305: k=31 canonical values and checksum 0x8b0c176c3b43d207."""
    r = oracle_py.count_stream(b"GATCGATTCTTAGCACGTCCCCCCCTACACCCAAT", 31, abundance_min=1)
    ok = {0x1CA68D1E55561150, 0x09CA68D1E5556115, 0x2729A34795558454, 0x32729A3479555845, 0x0AFEE3FFF1ED8309}
    assert set(r["lo"].tolist()) == ok
    assert sum(r["lo"].tolist()) & (2**64 - 1) == 0x8b0c176c3b43d207


def test_count_stream_threads_agree(oracle):
    rng = np.random.default_rng(5)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=80, p=[.245, .245, .245, .245, .02])) for _ in range(3000)]
    reads += reads[:1500]
    stream = b"\n".join(reads)
    for k in (21, 41):
        a = oracle_py.count_stream(stream, k, abundance_min=2, nthreads=1)
        b = oracle_py.count_stream(stream, k, abundance_min=2, nthreads=4)
        assert (a["lo"] == b["lo"]).all() and (a["hi"] == b["hi"]).all() and (a["abundance"] == b["abundance"]).all()
        assert (a["histogram"] == b["histogram"]).all()


@pytest.mark.parametrize("name", ["full", "full_k63", "syn_small_k31"])
def test_graph_build_threads_agree(oracle_bin, tmp_path, name):
    """The threaded graph build of the CPU baseline (oracle/graph_mt.hpp) gives the same Bloom / cascade / BooPHF bits, sizes and
    cFP set as the single-threaded restatement, and `oracle_find -nb-cores 4` the same files as `-nb-cores 1`."""
    case = CASES[name]
    reads, _ = case_paths(case)
    stream = b"\n".join(s for _, s in oracle_py.read_sequences(reads)) + b"\n"
    amin = -1
    for i, fl in enumerate(case["flags"]):
        if fl == "-abundance-min":
            amin = int(case["flags"][i + 1])
    o = oracle_py.count_stream(stream, case["k"], abundance_min=amin, nthreads=2)
    a = oracle_py.Graph(o["lo"], o["hi"], case["k"])
    b = oracle_py.Graph(o["lo"], o["hi"], case["k"], nthreads=4)
    assert a.info() == b.info()
    for which in (0, 1, 2, 3, 5):
        assert (a.bits(which) == b.bits(which)).all(), which
    rng = np.random.default_rng(2)
    qlo = np.concatenate([o["lo"][:3000], rng.integers(0, 1 << 62, 3000, dtype=np.uint64)])
    qhi = np.concatenate([o["hi"][:3000], np.zeros(3000, dtype=np.uint64)])
    assert (a.query(qlo, qhi) == b.query(qlo, qhi)).all()
    a.close(); b.close()
    bk1, vcf1, info1 = run_oracle(oracle_bin, name, tmp_path, extra=["-nb-cores", "1"])
    bk4, vcf4, info4 = run_oracle(oracle_bin, name, tmp_path, extra=["-nb-cores", "4"])
    assert bk1 == bk4 and vcf1 == vcf4 and info1["nb_solid"] == info4["nb_solid"]


H5_CASES = ["full", "full_k63", "full_k21_amin3", "inserts_ref10k", "hetero_insert", "syn_tiny_k31", "syn_tiny_k32", "syn_tiny_k63",
            "syn_small_k31", "syn_small_k47_homo"]


@pytest.mark.parametrize("name", H5_CASES)
def test_oracle_bloom_bytes_equal_reference_h5_datasets(oracle, name):
    """The "golden bits" of SURVEY.md 8c: /bloom/bloom and /debloom/bloom2,3,4 of the .h5 written by the unmodified reference
    binary (tests/golden/ref_outputs/<case>.h5bits.json, dumped with gatb-h5dump) equal the oracle's arrays byte for byte, and
    /debloom/cfp holds as many k-mers as the oracle's cFP set."""
    import base64
    import hashlib
    import json
    case = CASES[name]
    fx = json.load(open(os.path.join(GOLD, "ref_outputs", name + ".h5bits.json")))
    reads, _ = case_paths(case)
    stream = b"\n".join(s for _, s in oracle_py.read_sequences(reads)) + b"\n"
    amin = -1
    for i, fl in enumerate(case["flags"]):
        if fl == "-abundance-min":
            amin = int(case["flags"][i + 1])
    o = oracle_py.count_stream(stream, case["k"], abundance_min=amin, nthreads=2)
    g = oracle_py.Graph(o["lo"], o["hi"], case["k"])
    for which, ds in enumerate(["/bloom/bloom", "/debloom/bloom2", "/debloom/bloom3", "/debloom/bloom4"]):
        if ds not in fx:
            assert which > 0 and g.info()["nb_critical"] == 0, ds   # no critical FP -> the reference writes no cascade
            continue
        bits = g.bits(which).tobytes()
        assert len(bits) == fx[ds]["bytes"], (ds, len(bits), fx[ds]["bytes"])
        assert hashlib.sha256(bits).hexdigest() == fx[ds]["sha256"], ds
        if "base64" in fx[ds]:
            assert bits == base64.b64decode(fx[ds]["base64"]), ds
    if "/debloom/cfp" in fx:
        assert g.info()["cfp_set"] * (8 if case["k"] <= 31 else 16) == fx["/debloom/cfp"]["bytes"]
    g.close()
