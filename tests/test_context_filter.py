"""Host logic of the connectivity post-filter (mindthegap_b200/context_filter.py), the drop-in for the reference's
scripts/python3/Context_genome_WG.py. The reference ships no golden output for this script, so the check is a literal, k-mer by
k-mer restatement of its loop (analyze_genomic_context_direct, lines 80-113) on degrees from the ORACLE graph, against the
product's batched implementation fed by the same degrees; the GPU test (test_gpu_parity.py) repeats it with the engine's degrees."""
import os

import numpy as np
import pytest

from mindthegap_b200.context_filter import context_filter, encode_kmers, parse_breakpoints
from tests import oracle_py
from tests.cases import CASES, GOLD, case_paths


def script_restatement(graph, k, bk_text, ref_records, threshold):
    """Context_genome_WG.py:80-113 line by line (Bio.SeqIO replaced by parse_breakpoints, graph[kmer] by the oracle's degrees)."""
    dico_first, dico_second, total, count = {}, {}, 0, 0
    for desc, _ in parse_breakpoints(bk_text):
        if count % 2 == 0:
            dico_first.setdefault(desc.split("_")[1], []).append(int(desc.split("_")[3]))
            total += 1
        count += 1
    for id_chrom, str_chromosome in ref_records:
        if id_chrom in dico_first:
            for value in dico_first[id_chrom]:
                if value - 49 - k < 0:
                    continue                      # documented deviation: the script's negative slice start
                sum_degree = []
                for i in range(50):
                    kmer = str_chromosome[value - i - k:value - i]
                    lo, hi = encode_kmers([kmer], k)
                    d = int(graph.degrees(lo, hi)[0])
                    sum_degree.append(d >> 4)     # node.out_degree
                    sum_degree.append(d & 15)     # node.in_degree
                if (sum_degree.count(1) + sum_degree.count(2)) / len(sum_degree) > threshold:
                    dico_second.setdefault(id_chrom, []).append(int(value))
    rows = []
    for desc, seq in parse_breakpoints(bk_text):
        if int(desc.split("_")[3]) in dico_second.get(desc.split("_")[1], []):
            rows.append(">" + desc + "\n" + seq + "\r\n")
    return "".join(rows), sum(len(v) for v in dico_second.values()), total


def bundled_graph():
    case = CASES["full"]
    reads, ref = case_paths(case)
    stream = b"\n".join(s for _, s in oracle_py.read_sequences(reads)) + b"\n"
    o = oracle_py.count_stream(stream, 31, abundance_min=-1, nthreads=2)
    refs = [(n, s.decode()) for n, s in oracle_py.read_sequences(ref)]
    return oracle_py.Graph(o["lo"], o["hi"], 31), refs


def test_encode_kmers_matches_oracle_kmers():
    seq = b"ACGTTGCANACGTACGTTTGACCAGTACGATCGATCGGGATATCGCGCTAGCTAGCTAGGATCGAC"
    for k in (5, 31, 32, 63):
        kms = oracle_py.kmers(seq.replace(b"N", b"G"), k)
        lo, hi = encode_kmers([seq[i:i + k].decode() for i in range(len(seq) - k + 1)], k)
        assert (lo == kms[0]).all() and (hi == kms[1]).all()   # forward k-mer, low and high words


@pytest.mark.parametrize("threshold", [0.80, 0.95, 0.5])
def test_filter_equals_script_restatement_on_the_reference_gold_breakpoints(threshold):
    g, refs = bundled_graph()
    bk = open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    want = script_restatement(g, 31, bk, refs, threshold)
    got = context_filter(g.degrees, 31, bk, refs, threshold)
    assert got == want
    assert got[2] == 8 and 0 <= got[1] <= 8
    if threshold == 0.5:
        assert got[1] > 0 and got[0].count(">") == 2 * got[1] and got[0].endswith("\r\n")
    g.close()


@pytest.mark.parametrize("threshold,kept", [(0.80, 6), (0.5, 8), (0.95, 6)])
def test_filter_equals_the_unmodified_reference_script(threshold, kept):
    """Pinned to the reference: tests/golden/context/ holds what the UNMODIFIED scripts/python3/Context_genome_WG.py printed and
    wrote on the bundled example (tests/golden/make_context_fixture.py: the script imported as it is, pygatb / Biopython replaced by
    stand-ins whose degrees come from gatb-core's own Graph::indegree / outdegree on the reference's .h5). Kept / total counts at
    every threshold, and the output file byte for byte where the script gets as far as writing it (at 0.80 and 0.95 it raises
    KeyError for the chromosome that keeps no breakpoint -- our documented deviation returns the 6 kept records instead)."""
    import json
    g, refs = bundled_graph()
    bk = open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    text, nkept, total = context_filter(g.degrees, 31, bk, refs, threshold)
    fx = json.load(open(os.path.join(GOLD, "context", "summary.json")))[str(threshold)]
    assert fx["stdout"] == "total breakpoints kept :  %d  on  %d" % (nkept, total) and nkept == kept
    if fx["status"] == "ok":
        assert text.encode() == open(os.path.join(GOLD, "context", "threshold_%s.bkpt" % threshold), "rb").read()
    else:
        assert fx["status"].startswith("KeyError") and text.count(">") == 2 * kept
    g.close()
