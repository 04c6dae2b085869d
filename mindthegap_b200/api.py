"""ctypes binding of include/mtg_b200.h and a `Finder` class mirroring MindTheGap's Finder (src/Finder.cpp).

There is NO fallback: if the CUDA library is missing or no CUDA device is present, construction fails loudly.
"""
import ctypes as C
import os
import re
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmtg_b200.so")

F_HOMO_ONLY, F_HOMO_INSERT, F_HETE_INSERT, F_SNP, F_BACKUP, F_DELETION, F_SMALL_HOMO = 1, 2, 4, 8, 16, 32, 64
F_HOST_PARSE = 128   # mtg_count_files: kseq-style host reader instead of the GPU parser
F_DEFAULT = F_HOMO_INSERT | F_HETE_INSERT | F_SNP | F_DELETION | F_SMALL_HOMO
ABUNDANCE_AUTO = -1


class MtgError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("abundance_min", C.c_int32), ("abundance_max", C.c_int64),
                ("minimizer_size", C.c_int32), ("max_repeat", C.c_int32), ("het_max_occ", C.c_int32),
                ("snp_min_val", C.c_int32), ("branching_filter", C.c_int32), ("flags", C.c_uint32), ("device", C.c_int32),
                ("stream", C.c_uint64)]


u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

EXPORTS = [
    "mtg_default_params", "mtg_create", "mtg_destroy", "mtg_last_error", "mtg_version", "mtg_count_reserve", "mtg_push_reads",
    "mtg_push_reads_device", "mtg_push_reads_text", "mtg_push_reads_text_device", "mtg_text_record_cut", "mtg_count_files", "mtg_count_finish", "mtg_get_threshold", "mtg_get_cutoff_auto", "mtg_get_nb_solid",
    "mtg_get_histogram", "mtg_get_stats", "mtg_stat_name", "mtg_export_solid", "mtg_load_solid", "mtg_set_reference",
    "mtg_contains_batch", "mtg_degree_batch", "mtg_ref_repeat_batch", "mtg_sequence_features", "mtg_sequence_features_device",
    "mtg_scan_reference", "mtg_scan_reference_bed", "mtg_scan_reference_device", "mtg_set_reference_device", "mtg_breakpoints_text", "mtg_vcf_text", "mtg_reset_outputs", "mtg_get_find_counters", "mtg_copy_bits",
    "mtg_bench_random_gather", "mtg_count_local_info", "mtg_count_copy_packed", "mtg_count_partition_records", "mtg_count_import",
    "mtg_count_run", "mtg_count_filter", "mtg_solid_copy", "mtg_graph_build_device", "mtg_sequence_features_device2", "mtg_replay_sequence",
    "mtg_graph_build_begin", "mtg_graph_critical", "mtg_graph_critical_copy", "mtg_graph_build_end", "mtg_set_host_threads", "mtg_set_minimizer_size", "mtg_get_minimizer_size",
    "mtg_solid_partition", "mtg_partition_keys", "mtg_graph_shard_begin", "mtg_graph_shard_critical", "mtg_graph_adj_pack", "mtg_graph_adj_unpack",
    "mtg_graph_branching", "mtg_export_dsk_partitions", "mtg_graph_critical_set_share", "mtg_graph_shard_cascade", "mtg_graph_set_cfp", "mtg_graph_shard_mphf_level", "mtg_graph_shard_mphf_begin", "mtg_graph_shard_mphf_plan", "mtg_graph_shard_mphf_step", "mtg_graph_shard_mphf_tail", "mtg_get_stream", "mtg_set_mode_flags", "mtg_renumber_text", "mtg_get_ids_used", "mtg_set_reference_sharded", "mtg_ref_repeats_copy", "mtg_set_ref_repeats_device", "mtg_graph_shard_finish", "mtg_graph_buffer", "mtg_or_chunks",
]

_lib = None


def text_record_cut(text: bytes, fmt: int, final=False) -> int:
    """mtg_text_record_cut: largest prefix of `text` that ends at a FASTA (fmt 1) / FASTQ (fmt 2) record start."""
    return int(load_library().mtg_text_record_cut(text, len(text), fmt, 1 if final else 0))


def build_library():
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "csrc")], check=True)
    return LIB_PATH


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MtgError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mtg_default_params.argtypes = [C.POINTER(_Params)]
    L.mtg_create.restype = vp
    L.mtg_create.argtypes = [C.POINTER(_Params)]
    L.mtg_destroy.argtypes = [vp]
    L.mtg_last_error.restype = C.c_char_p
    L.mtg_version.restype = C.c_char_p
    L.mtg_count_reserve.argtypes = [vp, C.c_uint64]
    L.mtg_push_reads.argtypes = [vp, vp, C.c_uint64]
    L.mtg_push_reads_device.argtypes = [vp, vp, C.c_uint64]
    L.mtg_count_files.argtypes = [vp, C.c_char_p]
    L.mtg_push_reads_text.argtypes = [vp, vp, C.c_uint64, C.c_int32]
    L.mtg_push_reads_text_device.argtypes = [vp, vp, C.c_uint64, C.c_int32]
    L.mtg_text_record_cut.restype = C.c_uint64
    L.mtg_text_record_cut.argtypes = [C.c_char_p, C.c_uint64, C.c_int32, C.c_int32]
    L.mtg_count_finish.argtypes = [vp]
    L.mtg_get_threshold.argtypes = [vp]
    L.mtg_get_cutoff_auto.argtypes = [vp]
    L.mtg_get_nb_solid.restype = C.c_uint64
    L.mtg_get_nb_solid.argtypes = [vp]
    L.mtg_get_histogram.argtypes = [vp, u64p]
    L.mtg_get_stats.argtypes = [vp, f64p, C.c_int]
    L.mtg_stat_name.restype = C.c_char_p
    L.mtg_stat_name.argtypes = [C.c_int]
    L.mtg_export_solid.argtypes = [vp, u64p, vp, vp, C.c_uint64]
    L.mtg_load_solid.argtypes = [vp, u64p, vp, C.c_uint64]
    L.mtg_export_dsk_partitions.argtypes = [vp, C.c_uint32, C.c_uint32, vp, u64p, u64p, vp, vp, C.c_uint64]
    L.mtg_graph_branching.argtypes = [vp, u64p, vp, vp, vp, vp, C.c_uint64]
    L.mtg_set_reference.argtypes = [vp, vp, C.c_uint64]
    for fn in (L.mtg_contains_batch, L.mtg_degree_batch, L.mtg_ref_repeat_batch):
        fn.argtypes = [vp, u64p, vp, C.c_uint64, u8p]
    L.mtg_sequence_features.argtypes = [vp, vp, C.c_uint64, u8p, u8p, u64p]
    L.mtg_sequence_features_device.argtypes = [vp, vp, C.c_uint64, vp, vp, u64p]
    L.mtg_scan_reference.argtypes = [vp, C.c_char_p, vp, C.c_uint64]
    L.mtg_scan_reference_device.argtypes = [vp, C.c_char_p, vp, vp, C.c_uint64]
    L.mtg_scan_reference_bed.argtypes = [vp, C.c_char_p, vp, C.c_uint64, vp, C.c_uint64]
    L.mtg_set_reference_device.argtypes = [vp, vp, C.c_uint64]
    L.mtg_breakpoints_text.restype = vp
    L.mtg_breakpoints_text.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mtg_vcf_text.restype = vp
    L.mtg_vcf_text.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mtg_reset_outputs.argtypes = [vp]
    L.mtg_set_host_threads.argtypes = [vp, C.c_int32]
    L.mtg_set_minimizer_size.argtypes = [vp, C.c_int32]
    L.mtg_get_minimizer_size.argtypes = [vp]
    L.mtg_get_minimizer_size.restype = C.c_int32
    L.mtg_get_find_counters.argtypes = [vp, u64p]
    L.mtg_copy_bits.restype = C.c_int64
    L.mtg_copy_bits.argtypes = [vp, C.c_int, vp, C.c_uint64]
    L.mtg_count_local_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.mtg_count_copy_packed.argtypes = [vp, vp, vp, C.c_uint64]
    L.mtg_count_partition_records.argtypes = [vp, C.c_int, C.c_uint64, vp, u64p]
    L.mtg_count_import.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint64]
    L.mtg_count_run.argtypes = [vp]
    L.mtg_count_filter.argtypes = [vp, vp]
    L.mtg_solid_copy.argtypes = [vp, vp, vp, C.c_uint64]
    L.mtg_graph_build_device.argtypes = [vp, vp, C.c_uint64]
    L.mtg_graph_build_begin.argtypes = [vp, vp, C.c_uint64]
    L.mtg_graph_critical.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.mtg_graph_critical_copy.argtypes = [vp, vp, C.c_uint64]
    L.mtg_graph_build_end.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64]
    L.mtg_solid_partition.argtypes = [vp, C.c_uint32, vp, u64p]
    L.mtg_partition_keys.argtypes = [vp, vp, C.c_uint64, C.c_uint32, vp, u64p]
    L.mtg_graph_shard_begin.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    L.mtg_graph_shard_critical.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mtg_graph_adj_pack.argtypes = [vp]
    L.mtg_graph_adj_unpack.argtypes = [vp]
    L.mtg_graph_critical_set_share.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.mtg_graph_shard_cascade.argtypes = [vp, C.c_int, C.c_uint64, C.POINTER(C.c_uint64)]
    L.mtg_graph_set_cfp.argtypes = [vp, vp, C.c_uint64]
    L.mtg_graph_shard_finish.argtypes = [vp]
    L.mtg_graph_shard_mphf_begin.argtypes = [vp]
    L.mtg_graph_shard_mphf_level.argtypes = [vp, C.c_int32]
    L.mtg_graph_shard_mphf_plan.argtypes = [vp, u64p, C.c_int32, C.POINTER(C.c_int32)]
    L.mtg_graph_shard_mphf_step.argtypes = [vp, C.c_int32, C.c_int32]
    L.mtg_graph_shard_mphf_tail.argtypes = [vp, vp]
    L.mtg_set_reference_sharded.argtypes = [vp, vp, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
    L.mtg_ref_repeats_copy.argtypes = [vp, vp, C.c_uint64]
    L.mtg_set_ref_repeats_device.argtypes = [vp, vp, C.c_uint64]
    L.mtg_set_mode_flags.argtypes = [vp, C.c_uint32]
    L.mtg_get_ids_used.restype = C.c_uint64
    L.mtg_get_ids_used.argtypes = [vp]
    L.mtg_renumber_text.restype = C.c_int64
    L.mtg_renumber_text.argtypes = [C.c_char_p, C.c_uint64, C.c_int32, C.c_uint64, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.mtg_get_stream.restype = C.c_void_p
    L.mtg_get_stream.argtypes = [vp]
    L.mtg_graph_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.mtg_or_chunks.argtypes = [vp, vp, C.c_uint32, C.c_uint64, vp]
    L.mtg_sequence_features_device2.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, u64p]
    L.mtg_replay_sequence.argtypes = [vp, C.c_char_p, vp, C.c_uint64, u8p, u8p, vp]
    L.mtg_bench_random_gather.restype = C.c_double
    L.mtg_bench_random_gather.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int]
    _lib = L
    return L


@dataclass
class FindParams:
    """Options of `MindTheGap find` (src/Finder.cpp:97-171); defaults are the reference's."""
    kmer_size: int = 31
    abundance_min: int = ABUNDANCE_AUTO  # "auto"
    abundance_max: int = 2147483647
    max_repeat: int = 5
    het_max_occ: int = 1
    snp_min_val: int = 5
    branching_filter: int = 15
    flags: int = F_DEFAULT
    device: int = 0
    minimizer_size: int = 10
    stream: int = 0   # cudaStream_t of the caller (0 = own stream); see include/mtg_b200.h

    @staticmethod
    def from_cli(args):
        """Parse `find`-style flags (subset used by the tests), same fixed evaluation order as src/Finder.cpp:321-398."""
        p = FindParams()
        it = iter(args)
        seen = set()
        for a in it:
            if a == "-kmer-size": p.kmer_size = int(next(it))
            elif a == "-abundance-min":
                v = next(it); p.abundance_min = ABUNDANCE_AUTO if v == "auto" else int(v)
            elif a == "-abundance-max": p.abundance_max = int(next(it))
            elif a == "-max-rep": p.max_repeat = int(next(it))
            elif a == "-het-max-occ": p.het_max_occ = max(1, int(next(it)))
            elif a == "-snp-min-val": p.snp_min_val = int(next(it))
            elif a == "-branching-filter": p.branching_filter = int(next(it))
            else: seen.add(a)
        homo_only, homo_insert, hete_insert, snp, backup, deletion, small = False, True, True, True, False, True, True
        if "-homo-only" in seen: homo_only, homo_insert, hete_insert, snp, backup, deletion = True, True, False, True, False, True
        if "-insert-only" in seen: homo_only, homo_insert, hete_insert, snp, backup, deletion = False, True, True, False, False, False
        if "-snp-only" in seen: homo_only, homo_insert, hete_insert, snp, backup, deletion = True, False, False, True, False, False
        if "-deletion-only" in seen: homo_only, homo_insert, hete_insert, snp, backup, deletion = True, False, False, False, False, True
        if "-hete-only" in seen: homo_only, homo_insert, hete_insert, snp, backup, deletion = False, False, True, False, False, False
        if "-backup" in seen: backup = True
        if "-no-snp" in seen: snp = False
        if "-no-insert" in seen: homo_insert = False
        if "-no-deletion" in seen: deletion = False
        if "-no-hetero" in seen: hete_insert = False
        p.flags = (F_HOMO_ONLY * homo_only | F_HOMO_INSERT * homo_insert | F_HETE_INSERT * hete_insert | F_SNP * snp |
                   F_BACKUP * backup | F_DELETION * deletion | F_SMALL_HOMO * small)
        return p


_STOI = re.compile(r"[ \t\n\v\f\r]*([+-]?[0-9]+)")


def parse_bed(text, chrom, k):
    """Intervals of one chromosome as the reference reads them (src/FindBreakpoints.hpp:462-495): skip empty lines and lines
    starting with '#' or '@'; tab-separated; field 0 == chromosome short name; begin/end through std::stoi semantics (leading
    blanks, sign, digits, the rest ignored); kept when (end - begin) > k in unsigned 64-bit arithmetic."""
    if isinstance(text, bytes):
        text = text.decode()
    if isinstance(chrom, bytes):
        chrom = chrom.decode()
    out = []
    for line in text.split("\n"):
        if not line or line[0] in "#@":
            continue
        v = line.split("\t")
        if v[0] != chrom:
            continue
        if len(v) < 3:
            raise MtgError("bed: fewer than 3 tab-separated fields in line: " + line)
        vals = []
        for f in v[1:3]:
            m = _STOI.match(f)
            if not m or not -2**31 <= int(m.group(1)) < 2**31:
                raise MtgError("bed: not a number (or out of int range) in line: " + line)
            vals.append(int(m.group(1)))
        if ((vals[1] - vals[0]) & 0xFFFFFFFFFFFFFFFF) > k:
            out.append((vals[0], vals[1]))
    return out


SOLID_HDR = np.dtype([("magic", "S8"), ("version", "<u4"), ("kmer_size", "<u4"), ("nb_partitions", "<u4"), ("minimizer_size", "<u4"),
                      ("n", "<u8"), ("nb_kmers_valid", "<u8"), ("nb_distinct", "<u8"), ("threshold", "<i4"), ("cutoff_auto", "<i4")])


def write_solid_bin(path, k, m, repart, offs, lo, hi, ab, histogram, threshold, cutoff_auto, nb_valid=0, nb_distinct=0):
    """The file mtg_h5 reads (csrc/h5_handoff.cpp: MtgSolidHeader, repart[4^m], part_offsets[P+1], histogram[10001], lo, (hi), abundance)."""
    h = np.zeros(1, dtype=SOLID_HDR)
    h["magic"] = b"MTGSOLID"; h["version"] = 1; h["kmer_size"] = k; h["nb_partitions"] = len(offs) - 1; h["minimizer_size"] = m
    h["n"] = len(lo); h["nb_kmers_valid"] = nb_valid; h["nb_distinct"] = nb_distinct; h["threshold"] = threshold
    h["cutoff_auto"] = cutoff_auto if cutoff_auto is not None else -1
    with open(path, "wb") as f:
        f.write(h.tobytes())
        f.write(np.ascontiguousarray(repart, dtype="<u2").tobytes())
        f.write(np.ascontiguousarray(offs, dtype="<u8").tobytes())
        f.write(np.ascontiguousarray(histogram, dtype="<u8").tobytes())
        f.write(np.ascontiguousarray(lo, dtype="<u8").tobytes())
        if k > 31:
            f.write(np.ascontiguousarray(hi, dtype="<u8").tobytes())
        f.write(np.ascontiguousarray(ab, dtype="<u4").tobytes())


def read_solid_bin(path):
    """(k, lo, hi, abundance) of a file written by `mtg_h5 dump` / write_solid_bin."""
    raw = np.fromfile(path, dtype=np.uint8)
    h = raw[:SOLID_HDR.itemsize].view(SOLID_HDR)[0]
    assert h["magic"] == b"MTGSOLID"
    k, n, P, m = int(h["kmer_size"]), int(h["n"]), int(h["nb_partitions"]), int(h["minimizer_size"])
    o = SOLID_HDR.itemsize + 2 * 4 ** m + 8 * (P + 1) + 8 * 10001
    lo = raw[o:o + 8 * n].view("<u8").copy(); o += 8 * n
    hi = np.zeros(n, dtype=np.uint64)
    if k > 31:
        hi = raw[o:o + 8 * n].view("<u8").copy(); o += 8 * n
    ab = raw[o:o + 4 * n].view("<u4").copy()
    return k, lo, hi, ab


def h5_tool_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "mtg_h5")


def run_h5_tool(verb, h5, binp):
    exe = h5_tool_path()
    if not os.path.exists(exe):
        raise MtgError("mtg_h5 is not built (it links the reference's gatb-core: run oracle/build_ref.sh where /root/reference exists)")
    r = subprocess.run([exe, verb, h5, binp], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=os.path.dirname(os.path.abspath(h5)) or ".")
    if r.returncode != 0:
        raise MtgError("mtg_h5 %s failed: %s" % (verb, r.stderr.strip()[-500:]))


def renumber_text(text, kind, offset):
    """(text with every bkpt id shifted by offset, largest id in the result); kind 0 = .breakpoints, 1 = VCF records. C++ (capi.cu)."""
    L = load_library()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    mx = C.c_uint64()
    cap = len(raw) + 24 * (raw.count(b"\n") + 1)
    buf = C.create_string_buffer(cap)
    n = L.mtg_renumber_text(raw, len(raw), kind, int(offset), buf, cap, C.byref(mx))
    if n < 0:
        raise MtgError("mtg_renumber_text: buffer too small")
    return buf.raw[:n].decode(), int(mx.value)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Finder:
    """Host mirror of MindTheGap's Finder for the GPU path: count reads -> graph -> scan reference -> outputs."""

    def __init__(self, params: FindParams = None):
        self.L = load_library()
        self.params = params or FindParams()
        p = _Params()
        self.L.mtg_default_params(C.byref(p))
        for f, _ in _Params._fields_:
            setattr(p, f, getattr(self.params, f))
        self.ctx = self.L.mtg_create(C.byref(p))
        if not self.ctx:
            raise MtgError(self.L.mtg_last_error().decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.L.mtg_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise MtgError(self.L.mtg_last_error().decode())

    # ---- stage 1
    def reserve(self, nb_bases):
        self._check(self.L.mtg_count_reserve(self.ctx, nb_bases))

    def push_reads(self, stream):
        """stream: bytes / uint8 array of ASCII bases, reads separated by a non-ACGT byte."""
        a = np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else np.ascontiguousarray(stream, dtype=np.uint8)
        self._check(self.L.mtg_push_reads(self.ctx, _ptr(a), a.size))

    def push_reads_device(self, dev_ptr, nbytes):
        self._check(self.L.mtg_push_reads_device(self.ctx, C.c_void_p(dev_ptr), nbytes))

    def push_reads_text(self, text, fmt=0):
        """Raw FASTA/FASTQ text (bytes or uint8 array; starts at a header, ends at a record boundary), parsed on the GPU.
        fmt: 0 by the first byte, 1 FASTA, 2 FASTQ."""
        a = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, dtype=np.uint8)
        self._check(self.L.mtg_push_reads_text(self.ctx, _ptr(a), a.size, int(fmt)))

    def push_reads_text_device(self, dev_ptr, nbytes, fmt=0):
        self._check(self.L.mtg_push_reads_text_device(self.ctx, C.c_void_p(dev_ptr), nbytes, int(fmt)))

    def count_files(self, uri):
        self._check(self.L.mtg_count_files(self.ctx, uri.encode()))

    def finish_count(self):
        self._check(self.L.mtg_count_finish(self.ctx))

    @property
    def threshold(self):
        return self.L.mtg_get_threshold(self.ctx)

    @property
    def cutoff_auto(self):
        return self.L.mtg_get_cutoff_auto(self.ctx)

    @property
    def nb_solid(self):
        return int(self.L.mtg_get_nb_solid(self.ctx))

    def histogram(self):
        h = np.zeros(10001, dtype=np.uint64)
        self._check(self.L.mtg_get_histogram(self.ctx, h))
        return h

    def stats(self):
        v = np.zeros(128, dtype=np.float64)
        n = self.L.mtg_get_stats(self.ctx, v, 128)
        return {self.L.mtg_stat_name(i).decode(): float(v[i]) for i in range(n)}

    def export_solid(self):
        n = self.nb_solid
        lo = np.zeros(max(n, 1), dtype=np.uint64); hi = np.zeros(max(n, 1), dtype=np.uint64)
        ab = np.zeros(max(n, 1), dtype=np.uint32)
        self._check(self.L.mtg_export_solid(self.ctx, lo, _ptr(hi), _ptr(ab), max(n, 1)))
        return lo[:n], hi[:n], ab[:n]

    def export_dsk_partitions(self, nb_partitions=4, minimizer_size=10):
        """Solid k-mers ordered by (DSK partition, k-mer) + the minimRepart table: (repart uint16[4^m], offsets[nparts+1], lo, hi, ab)."""
        n = self.nb_solid
        repart = np.zeros(4 ** minimizer_size, dtype=np.uint16)
        offs = np.zeros(nb_partitions + 1, dtype=np.uint64)
        lo = np.zeros(max(n, 1), dtype=np.uint64); hi = np.zeros(max(n, 1), dtype=np.uint64)
        ab = np.zeros(max(n, 1), dtype=np.uint32)
        self._check(self.L.mtg_export_dsk_partitions(self.ctx, nb_partitions, minimizer_size, _ptr(repart), offs, lo, _ptr(hi), _ptr(ab), max(n, 1)))
        return repart, offs, lo[:n], hi[:n], ab[:n]

    def branching(self, nodes=True):
        """BranchingAlgorithm (gatb-core debruijn/impl/BranchingAlgorithm.cpp:150-310): (nb_branching, topology[in][out] 5x5,
        lo, hi, abundance) with the collection sorted by k-mer; nodes=False only counts."""
        nb = np.zeros(1, dtype=np.uint64)
        topo = np.zeros(25, dtype=np.uint64)
        self._check(self.L.mtg_graph_branching(self.ctx, nb, _ptr(topo), None, None, None, 0))
        n = int(nb[0])
        if not nodes:
            return n, topo.reshape(5, 5)
        lo = np.zeros(max(n, 1), dtype=np.uint64); hi = np.zeros(max(n, 1), dtype=np.uint64)
        ab = np.zeros(max(n, 1), dtype=np.uint32)
        self._check(self.L.mtg_graph_branching(self.ctx, nb, _ptr(topo), _ptr(lo), _ptr(hi), _ptr(ab), max(n, 1)))
        return n, topo.reshape(5, 5), lo[:n], hi[:n], ab[:n]

    # ---- .h5 hand-off: the host tool mtg_h5 (csrc/h5_handoff.cpp, links the reference's gatb-core) does the HDF5 I/O
    def write_graph_h5(self, path, nb_partitions=4, minimizer_size=10, complete=True, nb_cores=0):
        """Write `path` (.h5) holding the GPU-counted solid k-mers in DSK's layout; complete=True then lets gatb-core finish the file
        (its own CPU Bloom / debloom / MPHF / branching, `mtg_h5 complete`) so that the unchanged `MindTheGap fill -graph` /
        `find -graph` load it like a graph `find` wrote itself. Off the timed path."""
        st = self.stats()
        repart, offs, lo, hi, ab = self.export_dsk_partitions(nb_partitions, minimizer_size)
        binp = path + ".solid.bin"
        write_solid_bin(binp, self.params.kmer_size, minimizer_size, repart, offs, lo, hi, ab, self.histogram(), self.threshold,
                        self.cutoff_auto, int(st["count.nb_valid_kmers"]), int(self.histogram().sum()))
        try:
            run_h5_tool("write", path, binp)
        finally:
            os.remove(binp)
        if complete:
            run_h5_tool("complete", path, str(nb_cores))

    def load_graph_h5(self, path):
        """`-graph x.h5`: dsk/solid of a gatb .h5 -> the device graph (replaces Graph::load, src/Finder.cpp:277)."""
        binp = path + ".solid.bin"
        try:
            run_h5_tool("dump", path, binp)
            k, lo, hi, _ = read_solid_bin(binp)
        finally:
            if os.path.exists(binp):
                os.remove(binp)
        if k != self.params.kmer_size:
            raise MtgError("graph %s was built with k=%d, this context uses k=%d" % (path, k, self.params.kmer_size))
        self.load_solid(lo, hi if k > 31 else None)

    def load_solid(self, lo, hi=None):
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        hi = None if hi is None else np.ascontiguousarray(hi, dtype=np.uint64)
        self._check(self.L.mtg_load_solid(self.ctx, lo, _ptr(hi), len(lo)))

    # ---- stage 2
    def set_reference(self, stream):
        a = np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else np.ascontiguousarray(stream, dtype=np.uint8)
        self._check(self.L.mtg_set_reference(self.ctx, _ptr(a), a.size))

    def set_reference_device(self, dev_ptr, nbytes):
        self._check(self.L.mtg_set_reference_device(self.ctx, C.c_void_p(dev_ptr), nbytes))

    def set_reference_sharded(self, dev_ptr, nbytes, nparts, part):
        n = C.c_uint64()
        self._check(self.L.mtg_set_reference_sharded(self.ctx, C.c_void_p(dev_ptr), int(nbytes), int(nparts), int(part), C.byref(n)))
        return int(n.value)

    def ref_repeats_copy(self, out_t, capacity):
        self._check(self.L.mtg_ref_repeats_copy(self.ctx, C.c_void_p(out_t.data_ptr()), int(capacity)))

    def set_ref_repeats_device(self, keys_t, n):
        self._check(self.L.mtg_set_ref_repeats_device(self.ctx, C.c_void_p(keys_t.data_ptr()), int(n)))

    def scan_reference_device(self, name, seq, dev_ptr):
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        self._check(self.L.mtg_scan_reference_device(self.ctx, name.encode(), _ptr(a), C.c_void_p(dev_ptr), a.size))

    def _batch(self, fn, lo, hi):
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        hi = None if hi is None else np.ascontiguousarray(hi, dtype=np.uint64)
        out = np.zeros(max(len(lo), 1), dtype=np.uint8)
        self._check(fn(self.ctx, lo, _ptr(hi), len(lo), out))
        return out[:len(lo)]

    def contains(self, lo, hi=None):
        return self._batch(self.L.mtg_contains_batch, lo, hi)

    def degrees(self, lo, hi=None):
        return self._batch(self.L.mtg_degree_batch, lo, hi)

    def context_filter(self, breakpoints_text, ref_records, threshold=0.80):
        """Connectivity post-filter of the reference's scripts/python3/Context_genome_WG.py on this engine's graph:
        (filtered .breakpoints text, kept, total). See mindthegap_b200/context_filter.py."""
        from .context_filter import context_filter
        return context_filter(self.degrees, self.params.kmer_size, breakpoints_text, ref_records, threshold)

    def ref_repeat(self, lo, hi=None):
        return self._batch(self.L.mtg_ref_repeat_batch, lo, hi)

    def features(self, seq):
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        n = max(0, a.size - self.params.kmer_size + 1)
        f = np.zeros(max(n, 1), dtype=np.uint8); r = np.zeros(max(n, 1), dtype=np.uint8)
        c4 = np.zeros(4, dtype=np.uint64)
        self._check(self.L.mtg_sequence_features(self.ctx, _ptr(a), a.size, f, r, c4))
        return f[:n], r[:n], c4

    def features_device(self, dev_seq, length, dev_feat, dev_rep):
        c4 = np.zeros(4, dtype=np.uint64)
        self._check(self.L.mtg_sequence_features_device(self.ctx, C.c_void_p(dev_seq), length, C.c_void_p(dev_feat), C.c_void_p(dev_rep), c4))
        return c4

    def scan_reference(self, name, seq):
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        self._check(self.L.mtg_scan_reference(self.ctx, name.encode(), _ptr(a), a.size))

    def scan_reference_bed(self, name, seq, intervals):
        """-bed scan of one sequence; intervals = [(begin, end), ...] of this chromosome (parse_bed)."""
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        iv = np.array([x & 0xFFFFFFFFFFFFFFFF for pair in intervals for x in pair], dtype=np.uint64)
        self._check(self.L.mtg_scan_reference_bed(self.ctx, name.encode(), _ptr(a), a.size, _ptr(iv) if iv.size else None, iv.size // 2))

    def breakpoints_text(self):
        n = C.c_uint64()
        p = self.L.mtg_breakpoints_text(self.ctx, C.byref(n))
        return C.string_at(p, n.value).decode()

    def vcf_text(self):
        n = C.c_uint64()
        p = self.L.mtg_vcf_text(self.ctx, C.byref(n))
        return C.string_at(p, n.value).decode()

    def ids_used(self):
        return int(self.L.mtg_get_ids_used(self.ctx))

    def reset_outputs(self):
        self._check(self.L.mtg_reset_outputs(self.ctx))

    def set_minimizer_size(self, m):
        """Force the partitioning minimizer length (all GPUs of one find must agree); before the first push."""
        self._check(self.L.mtg_set_minimizer_size(self.ctx, int(m)))

    def set_host_threads(self, n):
        """-nb-cores: host threads of the chunked event replay (0 = all cores)."""
        self._check(self.L.mtg_set_host_threads(self.ctx, int(n)))

    def find_counters(self):
        o = np.zeros(12, dtype=np.uint64)
        self._check(self.L.mtg_get_find_counters(self.ctx, o))
        names = ["homo_clean", "homo_fuzzy", "hetero_clean", "hetero_fuzzy", "clean_deletion", "fuzzy_deletion", "solo_snp", "multi_snp",
                 "backup", "homo_indel", "hetero_indel", "observer_queries"]
        return dict(zip(names, o.tolist()))

    def copy_bits(self, which):
        n = self.L.mtg_copy_bits(self.ctx, which, None, 0)
        if n < 0:
            raise MtgError(self.L.mtg_last_error().decode())
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        self.L.mtg_copy_bits(self.ctx, which, _ptr(buf), n)
        return buf[:n]

    # ---- multi-GPU building blocks (device buffers are torch tensors owned by the caller; see dist.py)
    def count_local_info(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.L.mtg_count_local_info(self.ctx, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def count_copy_packed(self, packed_t, inv_t):
        assert packed_t.numel() == inv_t.numel()
        self._check(self.L.mtg_count_copy_packed(self.ctx, C.c_void_p(packed_t.data_ptr()), C.c_void_p(inv_t.data_ptr()), packed_t.numel()))

    def count_partition_records(self, nparts, pos_offset_bases, out_t):
        counts = np.zeros(nparts, dtype=np.uint64)
        self._check(self.L.mtg_count_partition_records(self.ctx, nparts, pos_offset_bases, C.c_void_p(out_t.data_ptr()), counts))
        return [int(x) for x in counts]

    def count_import(self, packed_t, inv_t, records_t):
        self._check(self.L.mtg_count_import(self.ctx, C.c_void_p(packed_t.data_ptr()), C.c_void_p(inv_t.data_ptr()), packed_t.numel(),
                                            C.c_void_p(records_t.data_ptr()), records_t.numel()))

    def count_run(self):
        self._check(self.L.mtg_count_run(self.ctx))

    def count_filter(self, histogram=None):
        h = None if histogram is None else np.ascontiguousarray(histogram, dtype=np.uint64)
        self._check(self.L.mtg_count_filter(self.ctx, _ptr(h)))

    def nb_solid_local(self):
        """This rank's share of the solid set (between count_filter and graph_build_device)."""
        return int(self.L.mtg_get_nb_solid(self.ctx))

    @property
    def key_words(self):
        """64-bit words per k-mer key: 1 (k <= 31) or 2 ({lo, hi})."""
        return 1 if self.params.kmer_size <= 31 else 2

    solid_share_is_table_range = True   # the rank that counted a minimizer bin owns the table range of its k-mers (common.cuh mini_owner)

    def solid_copy(self, keys_t, counts_t):
        self._check(self.L.mtg_solid_copy(self.ctx, C.c_void_p(keys_t.data_ptr()), C.c_void_p(counts_t.data_ptr()) if counts_t is not None else None,
                                          keys_t.numel() // self.key_words))

    def graph_build_device(self, keys_t, n):
        self._check(self.L.mtg_graph_build_device(self.ctx, C.c_void_p(keys_t.data_ptr()), n))

    def graph_build_begin(self, keys_t, n):
        self._check(self.L.mtg_graph_build_begin(self.ctx, C.c_void_p(keys_t.data_ptr()), n))

    def graph_critical(self, keys_t, n):
        out = C.c_uint64()
        self._check(self.L.mtg_graph_critical(self.ctx, C.c_void_p(keys_t.data_ptr()), n, C.byref(out)))
        return out.value

    def graph_critical_copy(self, out_t):
        self._check(self.L.mtg_graph_critical_copy(self.ctx, C.c_void_p(out_t.data_ptr()), out_t.numel() // self.key_words))

    def graph_build_end(self, keys_t, n, cand_t, ncand):
        self._check(self.L.mtg_graph_build_end(self.ctx, C.c_void_p(keys_t.data_ptr()), n, C.c_void_p(cand_t.data_ptr()), ncand))

    # ---- graph build sharded over N GPUs (include/mtg_b200.h "SHARDED over N GPUs"; choreography in dist.py)
    def solid_partition(self, nshards, out_t):
        counts = np.zeros(nshards, dtype=np.uint64)
        self._check(self.L.mtg_solid_partition(self.ctx, nshards, C.c_void_p(out_t.data_ptr()), counts))
        return [int(x) for x in counts]

    def partition_keys(self, keys_t, n, nshards, out_t):
        counts = np.zeros(nshards, dtype=np.uint64)
        self._check(self.L.mtg_partition_keys(self.ctx, C.c_void_p(keys_t.data_ptr()), n, nshards, C.c_void_p(out_t.data_ptr()), counts))
        return [int(x) for x in counts]

    def graph_shard_begin(self, keys_t, n_share, n_total, max_share, nshards, shard):
        self._check(self.L.mtg_graph_shard_begin(self.ctx, C.c_void_p(keys_t.data_ptr()), n_share, n_total, max_share, nshards, shard))

    def graph_shard_critical(self):
        out = C.c_uint64()
        self._check(self.L.mtg_graph_shard_critical(self.ctx, C.byref(out)))
        return out.value

    def graph_adj_pack(self):
        self._check(self.L.mtg_graph_adj_pack(self.ctx))

    def graph_adj_unpack(self):
        self._check(self.L.mtg_graph_adj_unpack(self.ctx))

    def graph_critical_set_share(self, cand_t, n):
        out = C.c_uint64()
        self._check(self.L.mtg_graph_critical_set_share(self.ctx, C.c_void_p(cand_t.data_ptr()), n, C.byref(out)))
        return out.value

    def graph_shard_cascade(self, step, ncrit_total):
        out = C.c_uint64()
        self._check(self.L.mtg_graph_shard_cascade(self.ctx, step, ncrit_total, C.byref(out)))
        return out.value

    def graph_set_cfp(self, cfp_t, n):
        self._check(self.L.mtg_graph_set_cfp(self.ctx, C.c_void_p(cfp_t.data_ptr()), n))

    def graph_shard_mphf_level(self, level):
        self._check(self.L.mtg_graph_shard_mphf_level(self.ctx, int(level)))

    def graph_shard_mphf_plan(self, max_levels=24):
        caps = np.zeros(32, dtype=np.uint64)
        n = C.c_int32()
        self._check(self.L.mtg_graph_shard_mphf_plan(self.ctx, caps, int(max_levels), C.byref(n)))
        return [int(c) for c in caps[:n.value]]

    def graph_shard_mphf_step(self, level, phase):
        self._check(self.L.mtg_graph_shard_mphf_step(self.ctx, int(level), int(phase)))

    def graph_shard_mphf_tail(self, gathered_t):
        self._check(self.L.mtg_graph_shard_mphf_tail(self.ctx, C.c_void_p(gathered_t.data_ptr())))

    def stream_ptr(self):
        return self.L.mtg_get_stream(self.ctx)

    def graph_shard_mphf_begin(self):
        self._check(self.L.mtg_graph_shard_mphf_begin(self.ctx))

    def graph_shard_finish(self):
        self._check(self.L.mtg_graph_shard_finish(self.ctx))

    def graph_buffer(self, which):
        """Zero-copy uint8 torch view of a library-owned device buffer (valid until the next build call)."""
        import torch
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.mtg_graph_buffer(self.ctx, which, C.byref(p), C.byref(n)))
        if not n.value:
            return torch.empty(0, dtype=torch.uint8, device="cuda:%d" % self.params.device)

        class _View:
            __cuda_array_interface__ = {"shape": (n.value,), "typestr": "|u1", "data": (p.value, False), "version": 2}
        return torch.as_tensor(_View(), device="cuda:%d" % self.params.device)

    def or_chunks(self, in_t, nchunks, nwords, out_t):
        self._check(self.L.mtg_or_chunks(self.ctx, C.c_void_p(in_t.data_ptr()), nchunks, nwords, C.c_void_p(out_t.data_ptr())))

    def features_segment(self, seq_t):
        """seq_t: uint8 device tensor holding bases [a, b+k-1); returns (feat, rep, interest) device tensors for b-a positions."""
        import torch
        n = max(0, seq_t.numel() - self.params.kmer_size + 1)
        feat = torch.empty(max(n, 1), dtype=torch.uint8, device=seq_t.device)
        rep = torch.empty(max(n, 1), dtype=torch.uint8, device=seq_t.device)
        interest = torch.zeros((n + 31) // 32 + 1, dtype=torch.int32, device=seq_t.device)
        c4 = np.zeros(4, dtype=np.uint64)
        if n:
            self._check(self.L.mtg_sequence_features_device2(self.ctx, C.c_void_p(seq_t.data_ptr()), seq_t.numel(), C.c_void_p(feat.data_ptr()),
                                                             C.c_void_p(rep.data_ptr()), C.c_void_p(interest.data_ptr()), c4))
        return feat[:n], rep[:n], interest[:(n + 31) // 32]

    def replay_sequence(self, name, seq, feat, rep, interest=None):
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        feat = np.ascontiguousarray(feat, dtype=np.uint8); rep = np.ascontiguousarray(rep, dtype=np.uint8)
        it = None if interest is None else np.ascontiguousarray(interest).view(np.uint32)
        self._check(self.L.mtg_replay_sequence(self.ctx, name.encode(), _ptr(a), a.size, feat, rep, _ptr(it)))

    # ---- whole `find` on in-memory inputs (what bench.py and the parity tests drive)
    def find(self, read_stream, ref_records, bed_text=None):
        """read_stream: '\\n'-separated bases; ref_records: list of (name, bytes); bed_text: contents of a -bed file.
        Returns (breakpoints, vcf records)."""
        self.push_reads(read_stream)
        self.finish_count()
        self.set_reference(b"\n".join(s for _, s in ref_records))
        for name, seq in ref_records:
            if bed_text is None:
                self.scan_reference(name, seq)
            else:
                self.scan_reference_bed(name, seq, parse_bed(bed_text, name, self.params.kmer_size))
        return self.breakpoints_text(), self.vcf_text()
