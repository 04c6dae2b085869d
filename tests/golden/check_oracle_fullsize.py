#!/usr/bin/env python
"""Manual full-size check of the CPU side (not part of the default suites: ~15 min and ~8 GB of RAM for cfg3).

The -m gpu tests compare the GPU with the reference binary's fixtures at full size; the default CPU suite pins the oracle and the
host replay on the 20 small cases. This script closes the remaining corner on a machine without a GPU, for one full-size case:
  1. `oracle/_ref/bin/h5solid dump` reads the solid set back from the .h5 the reference wrote (make_fullsize_fixtures.py keeps it
     under $MTG_FULLSIZE_TMP/<config>_<seed>/ref_k<k>.h5),
  2. the ORACLE graph built from that set: Bloom / bloom2-4 bytes and cFP count against the fixture's .h5 hashes,
  3. the ORACLE scan (oracle/scan_oracle.hpp) of the reference sequences: `.breakpoints` and VCF records against the fixture texts,
  4. the PRODUCT's host replay (csrc/replay.hpp through tests/host/replay_check.cpp, oracle features and probe answers, staged +
     chunked on 4 threads): the same texts, and its observer query count.
r02 result for cfg3_k31: all equal (63 947 508 solid k-mers; 3 846 breakpoint records, 6 671 VCF records; 3 966 691 observer queries, the number the GPU run reports).

  python tests/golden/check_oracle_fullsize.py [case]        (default cfg3_k31)
"""
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import oracle_py as o  # noqa: E402
from tests.fullsize import FULLSIZE, fixture  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_k31"
    case = FULLSIZE[name]
    fx, bk_ref, vcf_ref = fixture(name)
    k = case["k"]
    d = os.path.join(os.environ.get("MTG_FULLSIZE_TMP", "/tmp/mtg_fullsize"), "%s_%d" % (case["config"], case["seed"]))
    h5 = os.path.join(d, "ref_k%d.h5" % k)
    if not os.path.exists(h5):
        raise SystemExit("run tests/golden/make_fullsize_fixtures.py %s first (it leaves %s)" % (name, h5))
    solid_bin = os.path.join(d, "solid_k%d.bin" % k)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "bin", "h5solid"), "dump", h5, solid_bin], check=True, stdout=subprocess.DEVNULL)
    rec = np.fromfile(solid_bin, dtype=[("lo", "<u8"), ("hi", "<u8"), ("ab", "<u4"), ("part", "<u4")])
    assert len(rec) == fx["solid"]["n"], (len(rec), fx["solid"]["n"])
    ok = True
    t0 = time.time()
    g = o.Graph(rec["lo"].copy(), rec["hi"].copy(), k, nthreads=os.cpu_count() or 1)
    print("oracle graph of %d k-mers built in %.0f s" % (len(rec), time.time() - t0), flush=True)
    del rec
    for which, ds in enumerate(["/bloom/bloom", "/debloom/bloom2", "/debloom/bloom3", "/debloom/bloom4"]):
        ref_bits = fx["h5_bits"].get(ds)
        if ref_bits is None:
            continue
        same = hashlib.sha256(g.bits(which).tobytes()).hexdigest() == ref_bits["sha256"]
        print("%-16s %s" % (ds, "equal" if same else "DIFFERENT"), flush=True)
        ok &= same
    if fx["h5_bits"].get("/debloom/cfp"):
        same = g.info()["cfp_set"] * (8 if k <= 31 else 16) == fx["h5_bits"]["/debloom/cfp"]["bytes"]
        print("%-16s %s" % ("/debloom/cfp", "equal size" if same else "DIFFERENT size"), flush=True)
        ok &= same
    seqs = o.read_sequences(os.path.join(d, "ref.fa"))
    g.set_reference(b"\n".join(s for _, s in seqs), 1)
    import mindthegap_b200 as m
    p = m.FindParams.from_cli(["-kmer-size", str(k)] + list(case["flags"]))
    t0 = time.time()
    bk, vcf = "", ""
    if len(seqs) == 1:   # Graph.scan starts every sequence with fresh ids
        bk, vcf = g.scan(seqs[0][0], seqs[0][1], p.max_repeat, p.het_max_occ, p.snp_min_val, p.branching_filter, p.flags)
        same = bk.encode() == bk_ref and vcf.encode() == vcf_ref
        print("oracle scan (%.0f s): %s" % (time.time() - t0, "outputs equal the reference binary's" if same else "outputs DIFFER"), flush=True)
        ok &= same
    g.close()
    exe = os.path.join(ROOT, "tests", "host", "_build", "replay_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "host", "replay_check.cpp"), "-lz"], check=True)
    out = os.path.join(d, "replay_check_k%d" % k)
    t0 = time.time()
    r = subprocess.run([exe, "-solid-bin", solid_bin, "-ref", os.path.join(d, "ref.fa"), "-kmer-size", str(k), "-max-rep", str(p.max_repeat),
                        "-het-max-occ", str(p.het_max_occ), "-snp-min-val", str(p.snp_min_val), "-branching-filter", str(p.branching_filter),
                        "-flags", str(p.flags), "-threads", "4", "-stages", "3", "-out", out], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True)
    info = dict(l.split() for l in r.stdout.strip().splitlines())
    same = open(out + ".breakpoints", "rb").read() == bk_ref and open(out + ".vcf", "rb").read() == vcf_ref
    print("host replay (%.0f s, %s observer queries): %s" % (time.time() - t0, info["observer_queries"],
                                                              "outputs equal the reference binary's" if same else "outputs DIFFER"), flush=True)
    ok &= same
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
