"""Breakpoint post-filter on k-mer connectivity, the GPU-backed drop-in for the reference's
scripts/python3/Context_genome_WG.py (analyze_genomic_context_direct, lines 65-114; usage in scripts/python3/README.txt).

For every breakpoint (every other record of the .breakpoints file, position = 4th '_'-field of the header) the script takes the 50
k-mers ending at reference offsets value, value-1, ... value-49 (`chrom[value-i-31 : value-i]`), asks the graph for the out- and
in-degree of each (pygatb `graph[kmer].out_degree / in_degree`, i.e. Graph::outdegree / indegree of the node in the strand of the
string, computed with `contains` like `find` does), and keeps the breakpoint when more than `threshold` (0.80) of the 100 degrees
are 1 or 2. The reference script needs the .h5 graph and pygatb; here the degrees come from the engine that just ran `find`
(`Finder.degrees`, one batched probe for all breakpoints), or from any `degrees_fn(lo, hi) -> uint8 (in | out << 4)`.

Kept as in the script: k-mer windows are plain string slices (an `N` is encoded like GATB encodes it, (c >> 1) & 3 = G); the
chromosome of a breakpoint is the 2nd '_'-field of its header and must equal the reference record's name; output rows are
`>description\\nsequence\\r\\n` (csv.writer with delimiter '\\n'). Deviations, on purpose: a breakpoint closer than 49 + k bases to the
start of its chromosome is dropped (the script's negative slice start gives a meaningless k-mer), and a chromosome without any kept
breakpoint does not raise KeyError.
"""
import numpy as np

WINDOW = 50   # k-mers looked at before each breakpoint (Context_genome_WG.py:96)


def parse_breakpoints(text):
    """[(description, sequence)] of a .breakpoints file, like Bio.SeqIO.parse(..., "fasta") reads it (description = header line
    without '>' and trailing white space, sequence = following lines joined)."""
    recs = []
    for block in text.split(">")[1:]:
        lines = block.split("\n")
        recs.append((lines[0].rstrip(), "".join(l.strip() for l in lines[1:])))
    return recs


def encode_kmers(seqs, k):
    """2-bit GATB codes ((c >> 1) & 3, first base most significant) of equal-length strings -> (lo, hi) uint64 arrays."""
    n = len(seqs)
    a = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(n, k) if n else np.zeros((0, k), dtype=np.uint8)
    codes = ((a >> 1) & 3).astype(np.uint64)
    lo = np.zeros(n, dtype=np.uint64); hi = np.zeros(n, dtype=np.uint64)
    for i in range(k):
        sh = 2 * (k - 1 - i)
        if sh >= 64:
            hi |= codes[:, i] << np.uint64(sh - 64)
        else:
            lo |= codes[:, i] << np.uint64(sh)
    return lo, hi


def context_filter(degrees_fn, k, breakpoints_text, ref_records, threshold=0.80):
    """Returns (filtered .breakpoints text, number kept, number of breakpoints). ref_records: [(name, sequence str or bytes)]."""
    recs = parse_breakpoints(breakpoints_text)
    first = {}
    for i, (desc, _) in enumerate(recs):
        if i % 2 == 0:
            f = desc.split("_")
            first.setdefault(f[1], []).append(int(f[3]))
    total = sum(len(v) for v in first.values())
    # all windows of all breakpoints in one batch
    owners, seqs = [], []
    for name, seq in ref_records:
        if name not in first:
            continue
        s = seq.decode() if isinstance(seq, (bytes, bytearray)) else (seq.tobytes().decode() if hasattr(seq, "tobytes") else seq)
        for value in first[name]:
            if value - (WINDOW - 1) - k < 0 or value > len(s):
                continue
            owners.append((name, value))
            seqs.extend(s[value - i - k: value - i] for i in range(WINDOW))
    kept = {}
    if seqs:
        lo, hi = encode_kmers(seqs, k)
        d = np.asarray(degrees_fn(lo, hi if k > 31 else None)).reshape(len(owners), WINDOW)
        din, dout = d & 15, d >> 4
        good = ((din == 1) | (din == 2)).sum(axis=1) + ((dout == 1) | (dout == 2)).sum(axis=1)
        for (name, value), g in zip(owners, good.tolist()):
            if g / (2.0 * WINDOW) > threshold:
                kept.setdefault(name, []).append(value)
    out = []
    for desc, seq in recs:
        f = desc.split("_")
        if int(f[3]) in kept.get(f[1], ()):
            out.append(">" + desc + "\n" + seq + "\r\n")
    return "".join(out), sum(len(v) for v in kept.values()), total


def main(argv=None):
    """python -m mindthegap_b200.context_filter -in reads[,reads] -p reference.fa -b x.breakpoints -o filtered.breakpoints [-m 0.8] [-k 31]
    (-in replaces the script's -g graph.h5: the graph is rebuilt on the GPU from the reads)."""
    import argparse

    from .api import Finder, FindParams
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("-in", dest="reads", required=True)
    ap.add_argument("-p", dest="genome", required=True)
    ap.add_argument("-b", dest="bkpt", required=True)
    ap.add_argument("-o", dest="out", required=True)
    ap.add_argument("-m", dest="threshold", type=float, default=0.80)
    ap.add_argument("-k", dest="k", type=int, default=31)
    a = ap.parse_args(argv)
    f = Finder(FindParams(kmer_size=a.k))
    f.count_files(a.reads)
    f.finish_count()
    refs, name, parts = [], None, []
    for line in open(a.genome):
        if line.startswith(">"):
            if name is not None:
                refs.append((name, "".join(parts)))
            name, parts = line[1:].rstrip(), []
        else:
            parts.append(line.strip())
    if name is not None:
        refs.append((name, "".join(parts)))
    text, kept, total = context_filter(f.degrees, a.k, open(a.bkpt).read(), refs, a.threshold)
    with open(a.out, "w", newline="") as o:
        o.write(text)
    print("total breakpoints kept : ", kept, " on ", total)
    f.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
