#!/usr/bin/env python
"""Full-size golden fixtures from the UNMODIFIED reference binary (oracle/_ref/bin/MindTheGap, built by oracle/build_ref.sh).

For every case of tests/fullsize.py (BASELINE.json configs[1], [2] and [4]: cfg2 and cfg3 at k = 31 and k = 63, at their
stated size) this script
  1. writes the deterministic synthetic dataset (tools/synth.py make_dataset, the same generator and seed bench.py uses),
  2. runs `MindTheGap find -in r1.fq,r2.fq -ref ref.fa -kmer-size K -out o -nb-cores C` (stock code path),
  3. stores under tests/golden/fullsize/<case>.json: sha256 + sizes of `.breakpoints` and of the non-header VCF records, the
     `abundance_min` / `nb_solid_kmers` info lines, an order-independent checksum of the solid set read back from `dsk/solid` of
     the reference's .h5 (oracle/_ref/bin/h5solid), sha256 of the Bloom / debloom byte arrays of that .h5 (gatb-h5dump), the
     wall-clock of the reference run; and the two texts themselves, gzip'ed, so that a failing GPU test can show a diff.
The -m gpu tests (tests/test_gpu_fullsize.py) regenerate the same inputs on the GPU box and compare the engine's outputs with
these files; nothing there reads /root/reference or needs oracle/_ref.

  python tests/golden/make_fullsize_fixtures.py [case ...]        (default: all cases; needs ~10 GB under $MTG_FULLSIZE_TMP or /tmp)
"""
import gzip
import hashlib
import json
import os
import re
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from tests.fullsize import FULLSIZE, FULLSIZE_DIR  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
H5_SETS = ["/bloom/bloom", "/debloom/bloom2", "/debloom/bloom3", "/debloom/bloom4", "/debloom/cfp"]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def h5_bytes(h5, dataset, tmp):
    out = os.path.join(tmp, "ds.bin")
    if os.path.exists(out):
        os.remove(out)
    r = subprocess.run([os.path.join(REF_BIN, "gatb-h5dump"), "-d", dataset, "-b", "LE", "-o", out, h5], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if r.returncode != 0 or not os.path.exists(out):
        return None
    return open(out, "rb").read()


def run_case(name, case, tmp_root, cores):
    import synth
    d = os.path.join(tmp_root, "%s_%d" % (case["config"], case["seed"]))
    if not os.path.exists(os.path.join(d, "truth.json")):
        t0 = time.time()
        synth.make_dataset(d, synth.CONFIGS[case["config"]], case["seed"])
        print("  dataset written in %.0f s" % (time.time() - t0), flush=True)
    out = os.path.join(d, "ref_k%d" % case["k"])
    cmd = [os.path.join(REF_BIN, "MindTheGap"), "find", "-in", os.path.join(d, "r1.fq") + "," + os.path.join(d, "r2.fq"), "-ref", os.path.join(d, "ref.fa"),
           "-kmer-size", str(case["k"]), "-out", out, "-nb-cores", str(cores)] + case["flags"]
    t0 = time.time()
    r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    if r.returncode != 0:
        raise SystemExit("%s: reference failed: %s" % (name, r.stderr[-800:]))
    bk = open(out + ".breakpoints", "rb").read()
    vcf = b"".join(l for l in open(out + ".othervariants.vcf", "rb") if not l.startswith(b"#"))
    info = [l.strip() for l in r.stdout.splitlines() if re.search(r"abundance_min|nb_solid_kmers|nb_branching_nodes", l)]
    times = [l.strip() for l in r.stdout.splitlines() if re.search(r"^\s*(time|Time|graph construction|find)", l)]
    s = subprocess.run([os.path.join(REF_BIN, "h5solid"), "sum", out + ".h5"], stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    solid = dict(zip(s[0::2], s[1::2]))
    bits = {}
    for ds in H5_SETS:
        b = h5_bytes(out + ".h5", ds, d)
        bits[ds] = None if b is None else {"bytes": len(b), "sha256": sha(b)}
    fx = {"case": name, "config": case["config"], "seed": case["seed"], "k": case["k"], "flags": case["flags"],
          "command": " ".join(",".join(os.path.basename(x) for x in c.split(",")) if c.startswith("/") else c for c in cmd),
          "reference_wall_s": round(wall, 2), "reference_cores": cores, "reference_times": times, "info": info,
          "breakpoints": {"bytes": len(bk), "sha256": sha(bk), "records": bk.count(b">") // 2},
          "vcf": {"bytes": len(vcf), "sha256": sha(vcf), "records": vcf.count(b"\n")},
          "solid": {"n": int(solid["n"]), "xor_lo": solid["xor_lo"], "xor_hi": solid["xor_hi"], "mixsum": solid["mixsum"],
                    "abundance_sum": int(solid["abundance_sum"])},
          "h5_bits": bits}
    os.makedirs(FULLSIZE_DIR, exist_ok=True)
    json.dump(fx, open(os.path.join(FULLSIZE_DIR, name + ".json"), "w"), indent=1)
    with gzip.GzipFile(os.path.join(FULLSIZE_DIR, name + ".breakpoints.gz"), "wb", mtime=0) as f:
        f.write(bk)
    with gzip.GzipFile(os.path.join(FULLSIZE_DIR, name + ".vcf.gz"), "wb", mtime=0) as f:
        f.write(vcf)
    print("%s: %.0f s, %d breakpoint records, %d vcf records, %s solid k-mers, %s" % (
        name, wall, fx["breakpoints"]["records"], fx["vcf"]["records"], solid["n"], info), flush=True)


def main():
    names = sys.argv[1:] or list(FULLSIZE)
    tmp_root = os.environ.get("MTG_FULLSIZE_TMP", "/tmp/mtg_fullsize")
    os.makedirs(tmp_root, exist_ok=True)
    cores = os.cpu_count() or 1
    for n in names:
        print("== %s" % n, flush=True)
        run_case(n, FULLSIZE[n], tmp_root, cores)


if __name__ == "__main__":
    main()
