#!/usr/bin/env python
"""Generate tests/golden/ref_outputs/* by running the UNMODIFIED reference `MindTheGap find` binary.

Provenance: oracle/build_ref.sh compiles the reference from /root/reference (cmake recipe of SURVEY.md 8c) into
oracle/_ref/bin/MindTheGap; this script runs that binary (pass another path as argv[1]) on
  * the 11 cases of /root/reference/test/simple_test.sh:65-112 (inputs copied to tests/golden/simple/),
  * the bundled example of /root/reference/test/simple_full_test.sh:36 (inputs in tests/golden/full/),
  * small deterministic synthetic datasets made by tools/synth.py (k=31 and k=63),
and stores the `.breakpoints` file, the non-header VCF records, the `abundance_min`/`nb_solid_kmers` info lines, and
`<case>.h5bits.json`: size + sha256 of the Bloom / cascading-Bloom / cFP datasets of the .h5 the reference wrote (gatb-h5dump -b
LE), with the raw bytes (base64) when they are small -- the "golden bits" of SURVEY.md 8c.
The committed outputs are what tests/test_oracle_golden.py and the GPU parity tests compare against.
"""
import base64
import hashlib
import json
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
    binary = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref", "bin", "MindTheGap")
    h5dump = os.path.join(os.path.dirname(binary), "gatb-h5dump")
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
            bits = {}
            for ds in ("/bloom/bloom", "/debloom/bloom2", "/debloom/bloom3", "/debloom/bloom4", "/debloom/cfp"):
                raw = os.path.join(tmp, "ds.bin")
                if os.path.exists(raw):
                    os.remove(raw)
                rr = subprocess.run([h5dump, "-d", ds, "-b", "LE", "-o", raw, os.path.join(tmp, "o.h5")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                if rr.returncode == 0 and os.path.exists(raw):
                    b = open(raw, "rb").read()
                    bits[ds] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
                    if len(b) <= 16384:
                        bits[ds]["base64"] = base64.b64encode(b).decode()
        json.dump(bits, open(os.path.join(outdir, name + ".h5bits.json"), "w"), indent=1)
        open(os.path.join(outdir, name + ".breakpoints"), "w").write(bk)
        open(os.path.join(outdir, name + ".vcf"), "w").write(vcf)
        open(os.path.join(outdir, name + ".info"), "w").write(info)
        print(name, "ok", len(bk.splitlines()) // 4, "bkpt", len(vcf.splitlines()), "vcf")


if __name__ == "__main__":
    main()
