#!/usr/bin/env python
"""Deterministic synthetic datasets for `find`: random genome, planted variants, simulated reads.

genome  : i.i.d. uniform ACGT, numpy PCG64 seeded with `seed`
variants: planted at sorted random positions at least `spacing` apart:
          HOM insertion  (segment present in both donor haplotypes, deleted from the reference), length U[50,500]
          HET insertion  (segment present in haplotype 1 only, deleted from the reference)
          SNP            (donor base differs from the reference base, both haplotypes)
          DEL            (segment present in the reference, absent from both haplotypes), length U[10,200]
reads   : paired 2 x L, fragments sampled uniformly from both haplotypes and both strands, fixed fragment length
          3L, substitution errors at rate `err`, fixed quality; written as two FASTQ files or kept in memory.
Everything is vectorised with numpy so that the 4.6 Mbp / 50x configuration is generated in a few seconds.
"""
import argparse
import json
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgtN", b"TGCAtgcaN"):
    COMP[a] = b


def random_genome(n, rng):
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def plant(base, rng, n_hom=0, n_het=0, n_snp=0, n_del=0, spacing=1000, ins_len=(50, 500), del_len=(10, 200)):
    """Return (reference, hap1, hap2, truth list). `base` is the ancestral sequence (uint8 ASCII)."""
    n = len(base)
    total = n_hom + n_het + n_snp + n_del
    if total == 0:
        return base.copy(), base, base, []
    margin = 2000 if n > 20000 else 200
    slots = (n - 2 * margin) // spacing
    assert slots >= total, "genome too small for that many variants"
    pos = margin + np.sort(rng.choice(slots, size=total, replace=False)) * spacing
    kinds = np.array(["HOM"] * n_hom + ["HET"] * n_het + ["SNP"] * n_snp + ["DEL"] * n_del)
    rng.shuffle(kinds)
    ref_parts, h1_parts, h2_parts, truth = [], [], [], []
    cur = 0
    for p, kind in zip(pos.tolist(), kinds.tolist()):
        seg = base[cur:p]
        ref_parts.append(seg); h1_parts.append(seg); h2_parts.append(seg)
        if kind in ("HOM", "HET"):
            L = int(rng.integers(ins_len[0], min(ins_len[1], spacing // 2) + 1))
            ins = base[p:p + L]
            h1_parts.append(ins)
            if kind == "HOM":
                h2_parts.append(ins)
            truth.append((kind, p, L))
            cur = p + L
        elif kind == "DEL":
            L = int(rng.integers(del_len[0], min(del_len[1], spacing // 2) + 1))
            ref_parts.append(base[p:p + L])
            truth.append((kind, p, L))
            cur = p + L
        else:  # SNP: donor = base, reference gets another base
            b = base[p]
            alt = ACGT[(np.searchsorted(ACGT, b) + int(rng.integers(1, 4))) % 4]
            ref_parts.append(np.array([alt], dtype=np.uint8))
            h1_parts.append(base[p:p + 1]); h2_parts.append(base[p:p + 1])
            truth.append((kind, p, 1))
            cur = p + 1
    tail = base[cur:]
    ref_parts.append(tail); h1_parts.append(tail); h2_parts.append(tail)
    return np.concatenate(ref_parts), np.concatenate(h1_parts), np.concatenate(h2_parts), truth


def simulate_reads(haps, n_pairs, L, rng, err=0.005, chunk=1 << 20):
    """Yield (r1, r2) uint8 matrices of shape (m, L) in chunks."""
    frag = 3 * L
    ar = np.arange(L, dtype=np.int64)
    done = 0
    while done < n_pairs:
        m = min(chunk, n_pairs - done)
        done += m
        r1 = np.empty((m, L), dtype=np.uint8)
        r2 = np.empty((m, L), dtype=np.uint8)
        which = rng.integers(0, len(haps), size=m)
        strand = rng.integers(0, 2, size=m).astype(bool)
        for h, hap in enumerate(haps):
            sel = np.nonzero(which == h)[0]
            if len(sel) == 0:
                continue
            st = rng.integers(0, len(hap) - frag + 1, size=len(sel))
            left = hap[st[:, None] + ar]                       # forward, fragment start
            right = COMP[hap[st[:, None] + (frag - 1 - ar)]]   # revcomp of fragment end
            sw = strand[sel]
            r1[sel] = np.where(sw[:, None], right, left)
            r2[sel] = np.where(sw[:, None], left, right)
        for r in (r1, r2):
            ne = rng.binomial(r.size, err)
            if ne:
                idx = rng.integers(0, r.size, size=ne)
                flat = r.reshape(-1)
                flat[idx] = ACGT[(np.searchsorted(ACGT, flat[idx]) + rng.integers(1, 4, size=ne)) % 4]
        yield r1, r2


def write_fasta(path, name_seqs, width=80):
    with open(path, "wb") as f:
        for name, seq in name_seqs:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((n // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def fastq_block(reads, first_id, tag):
    """(m, L) matrix -> bytes of m FASTQ records with fixed-width names."""
    m, L = reads.shape
    ids = np.char.zfill((first_id + np.arange(m)).astype(str), 10).astype("S10")
    idm = np.frombuffer(ids.tobytes(), dtype=np.uint8).reshape(m, 10)
    hdr = np.frombuffer(("@" + tag).encode(), dtype=np.uint8)
    w = len(hdr) + 10 + 1 + L + 1 + 2 + L + 1
    out = np.empty((m, w), dtype=np.uint8)
    c = 0
    out[:, c:c + len(hdr)] = hdr; c += len(hdr)
    out[:, c:c + 10] = idm; c += 10
    out[:, c] = 10; c += 1
    out[:, c:c + L] = reads; c += L
    out[:, c] = 10; c += 1
    out[:, c] = ord("+"); out[:, c + 1] = 10; c += 2
    out[:, c:c + L] = ord("I"); c += L
    out[:, c] = 10
    return out.tobytes()


CONFIGS = {
    # name: genome_len, chroms, coverage, read_len, n_hom, n_het, n_snp, n_del
    "tiny": dict(genome_len=60_000, chroms=2, coverage=30, read_len=100, n_hom=6, n_het=4, n_snp=8, n_del=4, spacing=1000),
    "small": dict(genome_len=400_000, chroms=1, coverage=30, read_len=100, n_hom=20, n_het=10, n_snp=40, n_del=10, spacing=1000),
    "cfg2": dict(genome_len=4_600_000, chroms=1, coverage=50, read_len=150, n_hom=200, n_het=0, n_snp=0, n_del=0, spacing=2000),
    # one GPU's eighth of BASELINE configs[3] (3.1 Gbp human-scale genome, 24 chromosomes, 30x 2x100 bp): 3 chromosomes, 387.5 Mbp
    "cfg4s": dict(genome_len=387_500_000, chroms=3, coverage=30, read_len=100, n_hom=12000, n_het=12000, n_snp=38750, n_del=6000, spacing=2000),
    "cfg3": dict(genome_len=64_000_000, chroms=1, coverage=30, read_len=150, n_hom=2000, n_het=2000, n_snp=6400, n_del=1000, spacing=2000),
}


def build(cfg, seed):
    """Return (ref_records [(name, uint8 array)], haps_per_chrom [(h1,h2)], truth)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    chroms = cfg.get("chroms", 1)
    per = cfg["genome_len"] // chroms
    refs, haps, truth = [], [], []

    def share(n, c):
        return n // chroms + (1 if c < n % chroms else 0)

    for c in range(chroms):
        base = random_genome(per, rng)
        r, h1, h2, t = plant(base, rng, share(cfg["n_hom"], c), share(cfg["n_het"], c), share(cfg["n_snp"], c),
                             share(cfg["n_del"], c), spacing=cfg.get("spacing", 1000))
        refs.append(("chr%d" % (c + 1), r)); haps.append((h1, h2)); truth.append(t)
    return refs, haps, truth, rng


def reads_in_memory(cfg, seed):
    """Return (ref_records, list of read matrices) without touching the disk (bench / tests)."""
    refs, haps, truth, rng = build(cfg, seed)
    L = cfg["read_len"]
    mats = []
    for (h1, h2) in haps:
        n_pairs = int(cfg["coverage"] * len(h1) / (2 * L))
        for r1, r2 in simulate_reads((h1, h2), n_pairs, L, rng, err=cfg.get("err", 0.005)):
            mats.append(r1); mats.append(r2)
    return refs, mats, truth


def make_dataset(outdir, cfg, seed, fmt="fastq"):
    os.makedirs(outdir, exist_ok=True)
    refs, haps, truth, rng = build(cfg, seed)
    write_fasta(os.path.join(outdir, "ref.fa"), refs)
    L = cfg["read_len"]
    nid = 0
    with open(os.path.join(outdir, "r1.fq"), "wb") as f1, open(os.path.join(outdir, "r2.fq"), "wb") as f2:
        for (h1, h2) in haps:
            n_pairs = int(cfg["coverage"] * len(h1) / (2 * L))
            for r1, r2 in simulate_reads((h1, h2), n_pairs, L, rng, err=cfg.get("err", 0.005)):
                f1.write(fastq_block(r1, nid, "p")); f2.write(fastq_block(r2, nid, "p"))
                nid += len(r1)
    with open(os.path.join(outdir, "truth.json"), "w") as f:
        json.dump({"seed": seed, "cfg": cfg, "truth": truth}, f)
    return os.path.join(outdir, "r1.fq") + "," + os.path.join(outdir, "r2.fq"), os.path.join(outdir, "ref.fa")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("outdir")
    ap.add_argument("--seed", type=int, default=20240)
    a = ap.parse_args()
    print(make_dataset(a.outdir, CONFIGS[a.config], a.seed))


def truth_recall(bk_text, vcf_text, truth, chrom_names, tol=12):
    """Size-independent check of `find` outputs against the planted variants: which planted insertions have a breakpoint,
    which planted SNPs / deletions have a VCF record, within `tol` bases. truth: per chromosome [(kind, ancestral position,
    length)] in position order (build()); reference coordinate = ancestral position minus the inserted bases before it
    (insertions are absent from the reference; deletions and SNPs keep its length). Returns a dict of counts and ratios."""
    import bisect
    import re
    found = {"INS": {}, "SNP": {}, "DEL": {}}
    for m in re.finditer(r"^>bkpt\d+_(.+)_pos_(\d+)_fuzzy_\d+_(HOM|HET)\s+(?:REPEATED )?\s*left_kmer", bk_text, re.M):
        found["INS"].setdefault(m.group(1), []).append(int(m.group(2)))
    for line in vcf_text.splitlines():
        f = line.split("\t")
        if len(f) > 7 and not line.startswith("#"):
            t = "SNP" if "TYPE=SNP" in f[7] else "DEL" if "TYPE=DEL" in f[7] else None
            if t and not (t == "DEL" and len(f[3]) - len(f[4]) < 10):   # planted deletions are >= 10 bp
                found[t].setdefault(f[0], []).append(int(f[1]))
    for d in found.values():
        for v in d.values():
            v.sort()
    planted = {"INS": 0, "SNP": 0, "DEL": 0}
    hit = {"INS": 0, "SNP": 0, "DEL": 0}
    expected = {"INS": {}, "SNP": {}, "DEL": {}}
    for name, tr in zip(chrom_names, truth):
        shift = 0
        for kind, p, L in tr:
            t = "INS" if kind in ("HOM", "HET") else kind
            x = p - shift + (1 if t == "SNP" else 0)
            expected[t].setdefault(name, []).append(x)
            if t == "INS":
                shift += L
    def near(sorted_list, x):
        i = bisect.bisect_left(sorted_list, x - tol)
        return i < len(sorted_list) and sorted_list[i] <= x + tol
    out = {}
    for t in ("INS", "SNP", "DEL"):
        nfound = sum(len(v) for v in found[t].values())
        good = 0
        for name, xs in expected[t].items():
            planted[t] += len(xs)
            hit[t] += sum(near(found[t].get(name, []), x) for x in xs)
            xs_sorted = sorted(xs)
            good += sum(near(xs_sorted, y) for y in found[t].get(name, []))
        out[t] = {"planted": planted[t], "recovered": hit[t], "reported": nfound, "reported_at_planted_site": good,
                  "recall": hit[t] / planted[t] if planted[t] else None, "precision": good / nfound if nfound else None}
    return out
