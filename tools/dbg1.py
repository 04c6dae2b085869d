import sys, numpy as np
sys.path.insert(0, '.')
from tests import oracle_py
from tests.cases import CASES, case_paths
import mindthegap_b200 as m
from tests.test_gpu_parity import _finder, _stream
case = CASES["full_k63"]
reads, ref = case_paths(case)
stream, _ = _stream(reads)
f = _finder(case); f.push_reads(stream); f.finish_count()
lo, hi, ab = f.export_solid()
g = oracle_py.Graph(lo, hi, case["k"])
print("stats", {k: v for k, v in f.stats().items() if k.startswith("graph")})
print("oracle info", g.info())
got = f.contains(lo[:2000], hi[:2000])
exp = g.query(lo[:2000], hi[:2000])
print("solid queries: gpu bits", np.bincount(got, minlength=32)[:32].tolist())
print("solid queries: ora bits", np.bincount(exp, minlength=32)[:32].tolist())
a, b = f.copy_bits(0), g.bits(0)
# build a second oracle graph from the same keys but swapped halves to see if positions match
g2 = oracle_py.Graph(hi, lo, case["k"])
b2 = g2.bits(0)
print("swap test ndiff", int((a != b2).sum()), "orig ndiff", int((a != b).sum()))
