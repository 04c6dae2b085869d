"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of mindthegap_b200.dist.DistFind over a fake (numpy + oracle) engine
must give the outputs of the unmodified reference binary; plus unit checks of the merge helpers."""
import json
import os
import re
import socket
import subprocess
import sys

import pytest

from tests.cases import ROOT, expected


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("case,world,mode,build", [("syn_tiny_k31", 2, "auto", "sharded"), ("full", 2, "segments", "sharded"),
                                                   ("full", 2, "chromosomes", "replicated"), ("syn_tiny_k31", 3, "segments", "replicated"),
                                                   ("full", 3, "auto", "sharded")])
def test_two_rank_find_equals_reference(tmp_path, oracle, case, world, mode, build):
    out = str(tmp_path / "out.json")
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), case, out, mode, build], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = json.load(open(out))
    ebk, evcf, einfo = expected(case)
    assert res["bk"] == ebk
    assert res["vcf"] == evcf
    assert res["nb_solid"] == int(re.search(r"nb_solid_kmers\s*:\s*(\d+)", einfo).group(1))
    assert res["threshold"] == int(re.search(r"abundance_min \(used\)\s*:\s*(\d+)", einfo).group(1))
    assert res["scan_mode"] == (mode if mode != "auto" else res["scan_mode"])


def test_renumber_and_segments():
    from mindthegap_b200.dist import assign_chromosomes, renumber, segment_bounds
    owner, load = assign_chromosomes([10, 0, 7, 7, 3], 2)
    assert sorted(load) == [13, 14] and owner[0] != owner[2] and sum(load) == 27
    bk = ">bkpt1_chr2_pos_10_fuzzy_0_HOM  left_kmer\nACGT\n>bkpt1_chr2_pos_10_fuzzy_0_HOM  right_kmer\nACGT\n" \
         ">bkpt3_chr2_bkpt9_pos_99_fuzzy_1_HET REPEATED left_kmer\nAC\n"
    vcf = "chr2\t5\tbkpt2\tA\tC\t.\tPASS\tTYPE=SNP;LEN=1;FUZZY=0\tGT\t1/1\n"
    b2, v2, used = renumber(bk, vcf, 10)
    assert used == 3
    assert b2.count(">bkpt11_chr2_pos_10") == 2 and ">bkpt13_chr2_bkpt9_pos_99" in b2   # only the id at the line start moves
    assert v2.split("\t")[2] == "bkpt12"
    assert renumber("", "", 5) == ("", "", 0)
    for npos, w in [(1000, 3), (31, 4), (64, 2), (4546844, 8)]:
        b = segment_bounds(npos, w)
        assert b[0] == 0 and b[-1] == npos and all(x <= y for x, y in zip(b, b[1:])) and all(x % 32 == 0 or x == npos for x in b[:-1])
