"""ctypes access to the ORACLE (oracle/_build/liboracle.so). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    L = C.CDLL(path)
    L.mtgo_kmers.restype = C.c_uint64
    L.mtgo_kmers.argtypes = [C.c_char_p, C.c_uint64, C.c_int, u64p, u64p, u64p, u64p, u8p]
    L.mtgo_minimizer.restype = C.c_uint32
    L.mtgo_minimizer.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.mtgo_superkmers.restype = C.c_uint64
    L.mtgo_superkmers.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p, u32p, u32p]
    L.mtgo_hash1.restype = C.c_uint64
    L.mtgo_hash1.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64]
    L.mtgo_simplehash16.restype = C.c_uint64
    L.mtgo_simplehash16.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.mtgo_revcomp.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.mtgo_compute_threshold.restype = C.c_int
    L.mtgo_compute_threshold.argtypes = [u64p, C.c_uint64, C.c_int]
    L.mtgo_count_stream.restype = C.c_void_p
    L.mtgo_count_stream.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int64, C.c_int]
    L.mtgo_count_free.argtypes = [C.c_void_p]
    L.mtgo_count_nb_solid.restype = C.c_uint64
    L.mtgo_count_nb_solid.argtypes = [C.c_void_p]
    L.mtgo_count_threshold.restype = C.c_int
    L.mtgo_count_threshold.argtypes = [C.c_void_p]
    L.mtgo_count_cutoff_auto.restype = C.c_int
    L.mtgo_count_cutoff_auto.argtypes = [C.c_void_p]
    L.mtgo_count_stats.argtypes = [C.c_void_p, u64p]
    L.mtgo_count_histogram.argtypes = [C.c_void_p, u64p]
    L.mtgo_count_solid.argtypes = [C.c_void_p, u64p, u64p, u32p]
    L.mtgo_graph_new.restype = C.c_void_p
    L.mtgo_graph_new.argtypes = [u64p, u64p, C.c_uint64, C.c_int]
    L.mtgo_graph_new_threads.restype = C.c_void_p
    L.mtgo_graph_new_threads.argtypes = [u64p, u64p, C.c_uint64, C.c_int, C.c_int]
    L.mtgo_graph_free.argtypes = [C.c_void_p]
    L.mtgo_graph_query.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u8p]
    L.mtgo_graph_bits.restype = C.c_uint64
    L.mtgo_graph_bits.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mtgo_graph_info.argtypes = [C.c_void_p, u64p]
    L.mtgo_graph_degrees.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u8p]
    L.mtgo_graph_branching.restype = C.c_uint64
    L.mtgo_graph_branching.argtypes = [C.c_void_p, u64p, C.c_void_p, C.c_void_p]
    L.mtgo_graph_set_reference.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int]
    L.mtgo_graph_features.restype = C.c_uint64
    L.mtgo_graph_features.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, u8p, u8p]
    L.mtgo_graph_scan.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_char_p,
                                  C.c_char_p, u64p]
    _lib = L
    return L


def kmers(seq: bytes, k: int):
    L = load()
    n = max(0, len(seq) - k + 1)
    a = [np.zeros(max(n, 1), dtype=np.uint64) for _ in range(4)]
    v = np.zeros(max(n, 1), dtype=np.uint8)
    got = L.mtgo_kmers(seq, len(seq), k, a[0], a[1], a[2], a[3], v)
    assert got == n
    return [x[:n] for x in a] + [v[:n]]


def count_stream(stream: bytes, k: int, abundance_min=-1, abundance_max=2147483647, nthreads=1):
    """Returns dict(lo, hi, abundance, histogram, threshold, cutoff_auto, stats)."""
    L = load()
    h = L.mtgo_count_stream(stream, len(stream), k, abundance_min, abundance_max, nthreads)
    try:
        n = L.mtgo_count_nb_solid(h)
        lo = np.zeros(max(n, 1), dtype=np.uint64); hi = np.zeros(max(n, 1), dtype=np.uint64)
        ab = np.zeros(max(n, 1), dtype=np.uint32)
        L.mtgo_count_solid(h, lo, hi, ab)
        hist = np.zeros(10001, dtype=np.uint64)
        L.mtgo_count_histogram(h, hist)
        st = np.zeros(4, dtype=np.uint64)
        L.mtgo_count_stats(h, st)
        return dict(lo=lo[:n], hi=hi[:n], abundance=ab[:n], histogram=hist, threshold=L.mtgo_count_threshold(h),
                    cutoff_auto=L.mtgo_count_cutoff_auto(h), nb_kmers_total=int(st[0]), nb_kmers_valid=int(st[1]),
                    nb_distinct=int(st[2]))
    finally:
        L.mtgo_count_free(h)


class Graph:
    def __init__(self, lo, hi, k, nthreads=1):
        self.L = load()
        self.k = k
        lo = np.ascontiguousarray(lo, dtype=np.uint64); hi = np.ascontiguousarray(hi, dtype=np.uint64)
        self.h = self.L.mtgo_graph_new(lo, hi, len(lo), k) if nthreads <= 1 else self.L.mtgo_graph_new_threads(lo, hi, len(lo), k, nthreads)

    def close(self):
        if self.h:
            self.L.mtgo_graph_free(self.h); self.h = None

    def query(self, lo, hi):
        lo = np.ascontiguousarray(lo, dtype=np.uint64); hi = np.ascontiguousarray(hi, dtype=np.uint64)
        out = np.zeros(max(len(lo), 1), dtype=np.uint8)
        self.L.mtgo_graph_query(self.h, lo, hi, len(lo), out)
        return out[:len(lo)]

    def bits(self, which):
        n = self.L.mtgo_graph_bits(self.h, which, None)
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        self.L.mtgo_graph_bits(self.h, which, buf.ctypes.data_as(C.c_void_p))
        return buf[:n]

    def info(self):
        o = np.zeros(8, dtype=np.uint64)
        self.L.mtgo_graph_info(self.h, o)
        return dict(bloom=int(o[0]), nb_critical=int(o[1]), bloom2=int(o[2]), bloom3=int(o[3]), bloom4=int(o[4]),
                    cfp_set=int(o[5]), ref_repeated=int(o[6]), refbloom=int(o[7]))

    def degrees(self, lo, hi=None):
        """indegree | outdegree << 4 of forward k-mers (same contract as Finder.degrees)."""
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        hi = np.zeros(len(lo), dtype=np.uint64) if hi is None else np.ascontiguousarray(hi, dtype=np.uint64)
        out = np.zeros(max(len(lo), 1), dtype=np.uint8)
        self.L.mtgo_graph_degrees(self.h, lo, hi, len(lo), out)
        return out[:len(lo)]

    def branching(self):
        """(nb_branching, topology[in][out], lo, hi) -- BranchingAlgorithm restated (oracle/graph_oracle.hpp)."""
        topo = np.zeros(25, dtype=np.uint64)
        n = int(self.L.mtgo_graph_branching(self.h, topo, None, None))
        lo = np.zeros(max(n, 1), dtype=np.uint64); hi = np.zeros(max(n, 1), dtype=np.uint64)
        self.L.mtgo_graph_branching(self.h, topo, lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p))
        return n, topo.reshape(5, 5), lo[:n], hi[:n]

    def set_reference(self, stream: bytes, het_max_occ=1):
        self.L.mtgo_graph_set_reference(self.h, stream, len(stream), het_max_occ)

    def scan(self, name: str, seq: bytes, max_repeat=5, het_max_occ=1, snp_min_val=5, branching=15, flags=0x6E):
        """One sequence through the scan oracle with fresh state (ids from 1): (breakpoints text, vcf records)."""
        sizes = np.zeros(2, dtype=np.uint64)
        self.L.mtgo_graph_scan(self.h, name.encode(), seq, len(seq), max_repeat, het_max_occ, snp_min_val, branching, flags, None, None, sizes)
        bk = C.create_string_buffer(int(sizes[0]) + 1); vcf = C.create_string_buffer(int(sizes[1]) + 1)
        self.L.mtgo_graph_scan(self.h, name.encode(), seq, len(seq), max_repeat, het_max_occ, snp_min_val, branching, flags, bk, vcf, sizes)
        return bk.raw[:int(sizes[0])].decode(), vcf.raw[:int(sizes[1])].decode()

    def features(self, seq: bytes):
        n = max(0, len(seq) - self.k + 1)
        f = np.zeros(max(n, 1), dtype=np.uint8); r = np.zeros(max(n, 1), dtype=np.uint8)
        got = self.L.mtgo_graph_features(self.h, seq, len(seq), f, r)
        assert got == n
        return f[:n], r[:n]


def read_sequences(uri):
    """Tiny FASTA/FASTQ reader for tests (plain text). Returns list of (name, bytes)."""
    out = []
    for path in uri.split(","):
        with open(path, "rb") as f:
            lines = f.read().split(b"\n")
        i = 0
        while i < len(lines):
            ln = lines[i]
            if ln.startswith(b">"):
                name = ln[1:].split()[0] if ln[1:].split() else b""
                i += 1
                seq = []
                while i < len(lines) and not lines[i].startswith(b">"):
                    seq.append(lines[i].strip()); i += 1
                out.append((name.decode(), b"".join(seq)))
            elif ln.startswith(b"@"):
                name = ln[1:].split()[0]
                out.append((name.decode(), lines[i + 1].strip()))
                i += 4
            else:
                i += 1
    return out
