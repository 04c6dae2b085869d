#!/usr/bin/env python
"""Golden output of the reference's connectivity post-filter, produced by the UNMODIFIED script
/root/reference/scripts/python3/Context_genome_WG.py (analyze_genomic_context_direct): the script is imported as it is; `gatb` and
`Bio`, which it imports and which are not installable here, are the stand-ins of tests/golden/context_stubs/ -- the graph degrees
behind `graph[kmer]` come from gatb-core itself (oracle/_ref/bin/refdegrees: Graph::load, buildNode, indegree/outdegree on the
.h5 the unmodified `MindTheGap find` wrote). Inputs: the bundled example (tests/golden/full) with the reference's own gold
.breakpoints. Output: tests/golden/context/<threshold>.bkpt (the script's output file, byte for byte) + kept/total counts."""
import importlib.util
import io
import json
import os
import subprocess
import sys
import tempfile
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SCRIPT = "/root/reference/scripts/python3/Context_genome_WG.py"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def main():
    os.environ["MTG_REFDEGREES"] = os.path.join(REF_BIN, "refdegrees")
    sys.path.insert(0, os.path.join(HERE, "context_stubs"))
    spec = importlib.util.spec_from_file_location("Context_genome_WG", SCRIPT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    reads = ",".join(os.path.join(HERE, "full", f) for f in ("reads_r1.fastq", "reads_r2.fastq"))
    ref = os.path.join(HERE, "full", "reference.fasta")
    bk = os.path.join(HERE, "full", "gold.breakpoints")
    outdir = os.path.join(HERE, "context")
    os.makedirs(outdir, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        g = os.path.join(tmp, "g")
        subprocess.run([os.path.join(REF_BIN, "MindTheGap"), "find", "-in", reads, "-ref", ref, "-out", g], cwd=tmp, stdout=subprocess.PIPE, check=True)
        assert open(g + ".breakpoints").read() == open(bk).read()
        summary = {}
        for thr in (0.80, 0.5, 0.95):
            out = os.path.join(outdir, "threshold_%s.bkpt" % thr)
            buf = io.StringIO()
            try:
                with redirect_stdout(buf):
                    mod.analyze_genomic_context_direct(bk, g + ".h5", ref, out, thr)
                status = "ok"
            except KeyError as e:   # the script raises when a chromosome keeps no breakpoint (Context_genome_WG.py:112)
                status = "KeyError %s" % e
            line = [l for l in buf.getvalue().splitlines() if "total breakpoints kept" in l]
            summary[str(thr)] = {"status": status, "stdout": line[0] if line else "", "bytes": os.path.getsize(out)}
            print(thr, summary[str(thr)])
        json.dump(summary, open(os.path.join(outdir, "summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
