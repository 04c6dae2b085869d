#!/usr/bin/env python
"""Builds the patched copy of the reference's src/ for the reference-side binding (integration/finder_shim.hpp):
    python integration/make_shim.py /root/reference/src <out_dir>
copies every source file and, in Finder.cpp only, includes the shim header and replaces exactly three statements (each must be
found exactly once; anything else is an error, so a changed reference cannot be patched silently)."""
import os
import shutil
import sys

EDITS = [
    ("#include <FindSmallInsertion.hpp>", "#include <FindSmallInsertion.hpp>\n#include \"finder_shim.hpp\""),
    ("_graph = Graph::create (getInput());", "mtg_shim_create_graph(this);   /* was: _graph = Graph::create (getInput()); */"),
    ("_graph = Graph::load (getInput()->getStr(STR_URI_GRAPH));\n        _kmerSize = _graph.getKmerSize();",
     "mtg_shim_load_graph(this);   /* was: _graph = Graph::load(...); _kmerSize = _graph.getKmerSize(); */"),
    ("Integer::apply<runFindBreakpoints,Finder*> (_kmerSize, this);", "mtg_shim_scan(this);   /* was: Integer::apply<runFindBreakpoints,Finder*> (_kmerSize, this); */"),
]


def main():
    src, out = sys.argv[1], sys.argv[2]
    if os.path.isdir(out):
        shutil.rmtree(out)
    shutil.copytree(src, out)
    p = os.path.join(out, "Finder.cpp")
    s = open(p).read()
    for old, new in EDITS:
        if s.count(old) != 1:
            raise SystemExit("make_shim: expected exactly one %r in Finder.cpp, found %d" % (old, s.count(old)))
        s = s.replace(old, new)
    open(p, "w").write(s)
    shutil.copy(os.path.join(os.path.dirname(os.path.abspath(__file__)), "finder_shim.hpp"), os.path.join(out, "finder_shim.hpp"))
    print("make_shim: %s patched (%d edits)" % (p, len(EDITS)))


if __name__ == "__main__":
    main()
