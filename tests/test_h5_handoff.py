"""The .h5 hand-off, CPU part: a solid set in the product's dump format -> `mtg_h5 write` (the host tool that
links the reference's gatb-core, mindthegap_b200/csrc/h5_handoff.cpp) -> the UNMODIFIED reference binary consumes it:
`MindTheGap find -graph x.h5` must print the reference's own gold outputs and `MindTheGap fill -graph x.h5` must assemble the same
insertions as from the reference's own graph (test/simple_full_test.sh:123-163). The solid set comes from the oracle here; the
-m gpu test (tests/test_gpu_parity.py) feeds the GPU's export through the same path. Needs oracle/_ref (skipped otherwise)."""
import os
import subprocess

import numpy as np
import pytest

from tests import oracle_py
from tests.cases import CASES, GOLD, ROOT, case_paths

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "MindTheGap")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built (oracle/build_ref.sh)")


def dsk_layout(lo, hi, ab, k, m, nparts):
    """Any partition function consistent with its table is valid for gatb (DebloomMinimizerAlgorithm only needs the k-mers of
    partition p to be those whose minimizer maps to p): minimizer (gatb's rule, from the oracle) -> repart[minimizer] = minimizer % P."""
    L = oracle_py.load()
    mini = np.array([L.mtgo_minimizer(int(a), int(b), k, m, 1) for a, b in zip(lo, hi)], dtype=np.int64)
    repart = (np.arange(4 ** m) % nparts).astype(np.uint16)
    part = repart[mini].astype(np.int64)
    order = np.lexsort((lo, hi, part))
    offs = np.searchsorted(part[order], np.arange(nparts + 1)).astype(np.uint64)
    return repart, offs, lo[order], hi[order], ab[order]


def make_h5(tmp_path, name, nparts=4, m=10, complete=True):
    import mindthegap_b200.api as api
    if not os.path.exists(api.h5_tool_path()):
        pytest.skip("mtg_h5 not built")
    case = CASES[name]
    reads, ref = case_paths(case)
    stream = b"\n".join(s for _, s in oracle_py.read_sequences(reads)) + b"\n"
    o = oracle_py.count_stream(stream, case["k"], nthreads=2)
    repart, offs, lo, hi, ab = dsk_layout(o["lo"], o["hi"], o["abundance"], case["k"], m, nparts)
    binp, h5 = str(tmp_path / "s.bin"), str(tmp_path / "gpu.h5")
    api.write_solid_bin(binp, case["k"], m, repart, offs, lo, hi, ab, o["histogram"], o["threshold"], o["cutoff_auto"], o["nb_kmers_valid"], o["nb_distinct"])
    api.run_h5_tool("write", h5, binp)
    if complete:
        api.run_h5_tool("complete", h5, "2")   # gatb-core's own Graph::create finishes the file in place (Bloom, debloom, MPHF, branching)
    return h5, ref, o


@needs_ref
@pytest.mark.parametrize("name", ["full", "full_k63"])
def test_reference_find_consumes_the_handoff_h5(tmp_path, name):
    h5, ref, o = make_h5(tmp_path, name)
    out = str(tmp_path / "o")
    r = subprocess.run([REF_BIN, "find", "-graph", h5, "-ref", ref, "-out", out, "-nb-cores", "2"], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    want = os.path.join(GOLD, "ref_outputs", name)
    assert open(out + ".breakpoints").read() == open(want + ".breakpoints").read()
    assert "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#")) == open(want + ".vcf").read()
    assert ("nb_solid_kmers                           : %d" % len(o["lo"])) in r.stdout


@needs_ref
def test_reference_find_completes_a_counting_only_h5_itself(tmp_path):
    """`MindTheGap find -in x.h5`: the reference's Graph::create takes the counting-only file (state = k-mer counting done) and builds
    the rest itself (Graph.cpp:859-902) -- same gold outputs, same info lines."""
    h5, ref, o = make_h5(tmp_path, "full", complete=False)
    out = str(tmp_path / "o")
    r = subprocess.run([REF_BIN, "find", "-in", h5, "-ref", ref, "-out", out, "-nb-cores", "2"], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    assert open(out + ".breakpoints").read() == open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    for line in ("abundance_min (auto inferred)            : 7", "nb_solid_kmers                           : 7419", "nb_branching_nodes                       : 36"):
        assert line in r.stdout, line


@needs_ref
def test_reference_fill_consumes_the_handoff_h5(tmp_path):
    """simple_full_test.sh:123-163: fill from the graph + breakpoints; the assembled insertions must be those the reference obtains
    from the .h5 it built itself."""
    h5, ref, _ = make_h5(tmp_path, "full")
    reads, _ = case_paths(CASES["full"])
    own = str(tmp_path / "own")
    r = subprocess.run([REF_BIN, "find", "-in", reads, "-ref", ref, "-out", own, "-nb-cores", "2"], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0
    res = {}
    for tag, graph in (("ref", own + ".h5"), ("gpu", h5)):
        out = str(tmp_path / ("fill_" + tag))
        r = subprocess.run([REF_BIN, "fill", "-graph", graph, "-bkpt", own + ".breakpoints", "-out", out, "-nb-cores", "2"], cwd=tmp_path,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
        res[tag] = (open(out + ".insertions.fasta").read(), open(out + ".insertions.vcf").read() if os.path.exists(out + ".insertions.vcf") else "")
    assert res["gpu"][0] == res["ref"][0] and len(res["ref"][0]) > 0
    body = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("##"))
    assert body(res["gpu"][1]) == body(res["ref"][1])


def test_solid_bin_round_trip(tmp_path):
    import mindthegap_b200.api as api
    rng = np.random.default_rng(3)
    n = 1000
    lo = rng.integers(0, 1 << 62, n, dtype=np.uint64); hi = rng.integers(0, 1 << 60, n, dtype=np.uint64); ab = rng.integers(1, 99, n).astype(np.uint32)
    for k in (31, 47):
        p = str(tmp_path / ("x%d.bin" % k))
        api.write_solid_bin(p, k, 4, np.zeros(256, dtype=np.uint16), np.array([0, 400, n], dtype=np.uint64), lo, hi, ab, np.zeros(10001, dtype=np.uint64), 3, 3)
        k2, lo2, hi2, ab2 = api.read_solid_bin(p)
        assert k2 == k and (lo2 == lo).all() and (ab2 == ab).all() and ((hi2 == hi).all() if k > 31 else not hi2.any())
