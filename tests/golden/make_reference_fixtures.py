#!/usr/bin/env python
"""Generate tests/golden/ref_outputs/* by running the UNMODIFIED reference `MindTheGap find` binary.

Provenance (stated in DESIGN.md): the reference needs cmake + a generated config + vendored HDF5, so we do not
build it from a recipe in this repo (oracle/_ref is therefore absent). A binary built by the survey stage of this
project with the reference's own cmake exists in the build container at /tmp/mtg_build/bin/MindTheGap; this script
runs that binary (pass another path as argv[1]) on
  * the 11 cases of /root/reference/test/simple_test.sh:65-112 (inputs copied to tests/golden/simple/),
  * the bundled example of /root/reference/test/simple_full_test.sh:36 (inputs in tests/golden/full/),
  * small deterministic synthetic datasets made by tools/synth.py (k=31 and k=63),
and stores the `.breakpoints` file, the non-header VCF records, and the `abundance_min`/`nb_solid_kmers` info lines.
The committed outputs are what tests/test_oracle_golden.py and the GPU parity tests compare against.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.cases import CASES, case_paths  # noqa: E402


def main():
    binary = sys.argv[1] if len(sys.argv) > 1 else "/tmp/mtg_build/bin/MindTheGap"
    outdir = os.path.join(HERE, "ref_outputs")
    os.makedirs(outdir, exist_ok=True)
    for name, case in CASES.items():
        reads, ref = case_paths(case, make=True)
        with tempfile.TemporaryDirectory() as tmp:
            cmd = [binary, "find", "-in", reads, "-ref", ref, "-kmer-size", str(case["k"]), "-out", os.path.join(tmp, "o"),
                   "-nb-cores", "1"] + case["flags"]
            r = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                print(name, "FAILED", r.stderr[-500:])
                continue
            bk = open(os.path.join(tmp, "o.breakpoints")).read()
            vcf = "".join(l for l in open(os.path.join(tmp, "o.othervariants.vcf")) if not l.startswith("#"))
            info = "".join(l.strip() + "\n" for l in r.stdout.splitlines()
                           if re.search(r"abundance_min|nb_solid_kmers", l))
        open(os.path.join(outdir, name + ".breakpoints"), "w").write(bk)
        open(os.path.join(outdir, name + ".vcf"), "w").write(vcf)
        open(os.path.join(outdir, name + ".info"), "w").write(info)
        print(name, "ok", len(bk.splitlines()) // 4, "bkpt", len(vcf.splitlines()), "vcf")


if __name__ == "__main__":
    main()
