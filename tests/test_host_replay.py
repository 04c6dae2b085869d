"""Host logic of the product (no GPU): mindthegap_b200/csrc/replay.hpp -- gap state machine, observers, skip-ahead over
uninteresting runs and the collect-pass / probe-log prefetch -- driven by ORACLE-computed features and probe answers
(tests/host/replay_check.cpp) and compared with the outputs of the unmodified reference binary."""
import os
import subprocess

import pytest

from tests.cases import CASES, ROOT, case_paths, expected

EXE = os.path.join(ROOT, "tests", "host", "_build", "replay_check")


@pytest.fixture(scope="module")
def replay_check():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "host", "replay_check.cpp")
    deps = [src, os.path.join(ROOT, "mindthegap_b200", "csrc", "replay.hpp"), os.path.join(ROOT, "mindthegap_b200", "csrc", "seqio.hpp")] + [
        os.path.join(ROOT, "oracle", f) for f in ("scan_oracle.hpp", "graph_oracle.hpp", "kmer_oracle.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", EXE, src, "-lz"], check=True)
    return EXE


def run(exe, name, tmp_path, extra):
    import mindthegap_b200 as m
    case = CASES[name]
    reads, ref = case_paths(case)
    p = m.FindParams.from_cli(["-kmer-size", str(case["k"])] + list(case["flags"]))
    out = str(tmp_path / name)
    amin = "auto" if p.abundance_min < 0 else str(p.abundance_min)
    r = subprocess.run([exe, "-in", reads, "-ref", ref, "-kmer-size", str(p.kmer_size), "-abundance-min", amin, "-max-rep", str(p.max_repeat),
                        "-het-max-occ", str(p.het_max_occ), "-snp-min-val", str(p.snp_min_val), "-branching-filter", str(p.branching_filter),
                        "-flags", str(p.flags), "-out", out] + extra, stdout=subprocess.PIPE, text=True, check=True)
    info = dict(l.split() for l in r.stdout.strip().splitlines())
    return open(out + ".breakpoints").read(), open(out + ".vcf").read(), {k: int(v) for k, v in info.items()}


MODES = {"default": [], "tiny_segments": ["-seg", "777", "-skip-min", "1"], "walk_everything": ["-no-interest"],
         # chunked replay: cut at every steady point at least 64 positions after the previous cut, 4 host threads
         "parallel_chunks": ["-chunk", "64", "-threads", "4"], "parallel_chunks_1thread": ["-chunk", "1000", "-threads", "1", "-skip-min", "8"],
         # features arriving in stages: the replay of a stage starts before the next has arrived, batch probes run beside the pool
         # one chunk without steady points (no bitmap): its probe log is bounded per segment
         "walk_everything_small_segments": ["-no-interest", "-seg", "777", "-threads", "2"],
         "staged": ["-chunk", "64", "-threads", "4", "-stages", "3"], "staged_small_batches": ["-chunk", "64", "-threads", "3", "-stages", "5", "-seg", "500"]}


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("name", sorted(CASES))
def test_replay_equals_reference(replay_check, tmp_path, name, mode):
    bk, vcf, info = run(replay_check, name, tmp_path, MODES[mode])
    ebk, evcf, _ = expected(name)
    assert bk == ebk
    assert vcf == evcf


def test_collect_pass_foresees_most_queries(replay_check, tmp_path):
    """The collect pass + probe log must leave only a small share of the observer queries to immediate round trips."""
    for name in ("full", "syn_small_k31"):
        _, _, info = run(replay_check, name, tmp_path, [])
        assert info["observer_queries"] > 0
        assert info["unforeseen_queries"] <= 0.2 * info["observer_queries"], info


def test_chunked_replay_really_cuts(replay_check, tmp_path):
    """The parallel modes above must exercise more than one chunk per sequence (cuts at steady points of the gap machine)."""
    _, _, one = run(replay_check, "syn_small_k31", tmp_path, [])
    _, _, many = run(replay_check, "syn_small_k31", tmp_path, MODES["parallel_chunks"])
    assert one["chunks"] <= 4 and many["chunks"] > 50, (one, many)
    assert one["observer_queries"] == many["observer_queries"]


def test_bed_replay_equals_reference_gold_files(replay_check, tmp_path):
    """-bed (src/FindBreakpoints.hpp:459-553): product bed parser + restricted replay against the reference's own gold_bed
    files (/root/reference/test/simple_full_test.sh:79-118)."""
    from tests.cases import GOLD
    bed = os.path.join(GOLD, "full_bed", "gold.bed")
    bk, vcf, _ = run(replay_check, "full", tmp_path, ["-bed", bed])
    assert bk == open(os.path.join(GOLD, "full_bed", "gold_bed.breakpoints")).read()
    assert vcf == "".join(l for l in open(os.path.join(GOLD, "full_bed", "gold_bed.othervariants.vcf")) if not l.startswith("#"))


BED_CASES = {
    # interval starting at 0 (no reset), overlapping / unsorted intervals (stale ones dropped one per position), an interval
    # past the end, a malformed line with end < begin (kept by the unsigned test), comment lines, other chromosomes
    "edge": "#c\n@c\n\nSeq0\t0\t200\nSeq0\t100\t160\nSeq0\t50\t90\nSeq0\t300\t700 x\nother\t1\t1000\nSeq1\t400\t100\nSeq1\t500\t100000\nSeq2\t1\t40\nSeq2\t2\t20\n",
    "whole": "Seq0\t0\t100000\nSeq1\t0\t100000\nSeq2\t0\t100000\nSeq3\t1\t100000\n",
}


@pytest.mark.parametrize("bed_name", sorted(BED_CASES))
def test_bed_replay_equals_oracle(replay_check, oracle_bin, tmp_path, bed_name):
    """Edge cases of the interval walk: the product replay against the oracle's restatement of the reference loop."""
    bedf = tmp_path / "x.bed"
    bedf.write_text(BED_CASES[bed_name])
    bk, vcf, _ = run(replay_check, "full", tmp_path, ["-bed", str(bedf)])
    case = CASES["full"]
    reads, ref = case_paths(case)
    out = str(tmp_path / "o")
    subprocess.run([oracle_bin, "find", "-in", reads, "-ref", ref, "-kmer-size", "31", "-out", out, "-bed", str(bedf)],
                   stdout=subprocess.PIPE, check=True)
    assert bk == open(out + ".breakpoints").read()
    assert vcf == "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#"))
    assert len(bk) > 0


def test_host_reader_plain_and_gzip(tmp_path):
    """The host reader behind `mtg_find -ref` / MTG_F_HOST_PARSE (seqio.hpp): FASTA + FASTQ, multi-line, gzip or plain, comma list --
    the same records as the tests' own reader (gatb's BankFasta reads gzip through zlib, BankFasta.cpp:52-60)."""
    import gzip
    from tests import oracle_py
    exe = os.path.join(os.path.dirname(EXE), "seqio_check")
    src = os.path.join(ROOT, "tests", "host", "seqio_check.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src, "-lz"], check=True)
    case = CASES["full"]
    reads, ref = case_paths(case)
    r1, r2 = reads.split(",")
    gz = str(tmp_path / "r2.fastq.gz")
    with gzip.open(gz, "wb") as g:
        g.write(open(r2, "rb").read())
    refgz = str(tmp_path / "ref.fa.gz")
    with gzip.open(refgz, "wb") as g:
        g.write(open(ref, "rb").read())
    for uri, plain in ((",".join([r1, gz, refgz]), ",".join([r1, r2, ref])), (ref, ref)):
        out = subprocess.run([exe, uri], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        want = ["%s\t%s" % (n, s.decode()) for n, s in oracle_py.read_sequences(plain)]
        assert out == want
    r = subprocess.run([exe, str(tmp_path / "missing.fa")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Cannot open file" in r.stderr


def test_host_reader_file_of_files_and_headerless_input(tmp_path):
    """A list entry may be a "file of files" (README.md:166, gatb BankAlbum.cpp:48-94: one path per line, bare names relative to the
    album's directory); a non-empty file without any record is an error instead of a silently empty bank."""
    import shutil
    from tests import oracle_py
    exe = os.path.join(os.path.dirname(EXE), "seqio_check")
    src = os.path.join(ROOT, "tests", "host", "seqio_check.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src, "-lz"], check=True)
    reads, ref = case_paths(CASES["full"])
    r1, r2 = reads.split(",")
    shutil.copy(r1, tmp_path / "a.fastq")
    album = tmp_path / "reads.fof"
    album.write_text("a.fastq\n\n%s\n" % r2)          # bare name (album's directory) + absolute path + blank line
    nested = tmp_path / "all.fof"
    nested.write_text("reads.fof\n")
    want = ["%s\t%s" % (n, s.decode()) for n, s in oracle_py.read_sequences(reads)]
    for uri in (str(album), str(nested)):
        out = subprocess.run([exe, uri], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        assert out == want
    out = subprocess.run([exe, str(album) + "," + ref], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
    assert out == want + ["%s\t%s" % (n, s.decode()) for n, s in oracle_py.read_sequences(ref)]
    junk = tmp_path / "junk.txt"
    junk.write_text("this is not a sequence file\nnor a list of existing files\n")
    r = subprocess.run([exe, str(junk)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no FASTA/FASTQ record" in r.stderr
    empty = tmp_path / "empty.fa"
    empty.write_text("")
    r = subprocess.run([exe, str(empty)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout == ""


def test_observer_query_counts_are_pinned(replay_check, tmp_path):
    """The k-mers the observers ask about are part of the contract with the GPU probe (and a cheap detector of a silently
    changed observer: equal outputs can hide a dropped query whose answer happens to be `contained` on the test data).
    tests/golden/replay_query_counts.json: `observer_queries` of every case, from the replay whose outputs were compared with
    the reference binary at full size on the GPU (r02)."""
    import json
    pinned = json.load(open(os.path.join(ROOT, "tests", "golden", "replay_query_counts.json")))
    got = {}
    for name in sorted(CASES):
        _, _, info = run(replay_check, name, tmp_path, [])
        got[name] = info["observer_queries"]
    assert got == pinned
