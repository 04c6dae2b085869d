"""CPU checks of the boundary: the C-ABI library builds, loads and exports every symbol include/mtg_b200.h declares;
without a CUDA device the product fails loudly (no CPU fallback); the host-side CLI flag logic mirrors Finder.cpp."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.cases import ROOT

HEADER = os.path.join(ROOT, "include", "mtg_b200.h")
LIB = os.path.join(ROOT, "mindthegap_b200", "_build", "libmtg_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "mindthegap_b200", "csrc")], check=True)
    return ctypes.CDLL(LIB)


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libmtg_b200.so does not export %s" % s


def test_python_mirror_lists_every_symbol():
    import mindthegap_b200.api as api
    assert sorted(api.EXPORTS) == declared_symbols()


def test_no_cpu_fallback(lib):
    """Without a GPU mtg_create must fail with a message (this container has no CUDA device)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import mindthegap_b200 as m
    with pytest.raises(m.MtgError) as e:
        m.Finder(m.FindParams(kmer_size=31))
    assert "CUDA" in str(e.value)
    exe = os.path.join(ROOT, "mindthegap_b200", "_build", "mtg_find")
    r = subprocess.run([exe, "find", "-in", HEADER, "-ref", HEADER], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and r.stdout.startswith("EXCEPTION:")


def test_mode_flags_follow_finder_order():
    """src/Finder.cpp:321-398 applies the -x-only / -no-x flags in a fixed order, not in command-line order."""
    import mindthegap_b200.api as api
    p = api.FindParams.from_cli(["-no-snp", "-homo-only"])
    assert p.flags & api.F_HOMO_ONLY and not p.flags & api.F_SNP and not p.flags & api.F_HETE_INSERT and p.flags & api.F_DELETION
    p = api.FindParams.from_cli(["-hete-only", "-max-rep", "2"])
    assert p.max_repeat == 2 and p.flags == (api.F_HETE_INSERT | api.F_SMALL_HOMO)
    p = api.FindParams.from_cli(["-abundance-min", "auto", "-het-max-occ", "0"])
    assert p.abundance_min == api.ABUNDANCE_AUTO and p.het_max_occ == 1


def test_parse_bed_follows_reference_rules():
    """src/FindBreakpoints.hpp:462-495: comment lines, tab fields, std::stoi numbers, unsigned (end - begin) > k filter."""
    import mindthegap_b200.api as api
    from tests.cases import GOLD
    text = open(os.path.join(GOLD, "full_bed", "gold.bed")).read()
    assert api.parse_bed(text, "Seq0", 31) == [(60, 140), (90, 150), (200, 450)]
    assert api.parse_bed(text, "Seq1", 31) == [(300, 400), (700, 847)]
    assert api.parse_bed(text, "Seq3", 31) == []
    assert api.parse_bed("#x\n@y\n\nc\t10\t41\nc\t10\t42 note\nc\t400\t100\nd\t0\t99\n", "c", 31) == [(10, 42), (400, 100)]
    with pytest.raises(api.MtgError):
        api.parse_bed("c\tx\t10\n", "c", 31)
    with pytest.raises(api.MtgError):
        api.parse_bed("c\t10\n", "c", 31)


def test_text_record_cut(lib):
    """mtg_text_record_cut (host helper of the GPU text ingest): chunks end where a record starts; a quality line that
    begins with '@' is not mistaken for a header (gatb-core bank/impl/BankFasta.cpp:485-574 reads records sequentially,
    a chunked reader has to find the boundary from the text alone)."""
    lib.mtg_text_record_cut.restype = ctypes.c_uint64
    lib.mtg_text_record_cut.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32]

    def cut(text, fmt, final=False):
        return int(lib.mtg_text_record_cut(text, len(text), fmt, 1 if final else 0))

    rec = [b"@r0 x\nACGT\n+\nIIII\n", b"@r1\nGGGG\n+r1\n@III\n", b"@r2\nTTTT\n+\n@@@@\n", b"@r3\nCCCC\n+\nII"]
    fq = b"".join(rec)
    starts = [0, len(rec[0]), len(rec[0]) + len(rec[1]), len(rec[0]) + len(rec[1]) + len(rec[2])]
    assert cut(fq, 2, final=True) == len(fq)
    # every prefix: the cut is the last record start whose '+' line is inside the prefix, never a quality line
    for n in range(1, len(fq) + 1):
        c = cut(fq[:n], 2)
        assert c in starts and c < n
        complete = [s for s in starts[1:] if fq[:n].count(b"\n", s) >= 2 and n > fq.index(b"\n+", s) + 1]
        assert c == (max(complete) if complete else 0), (n, c)
    fa = b">a\nACGT\nAC\n>b c\nGG\n\n>c\nT"
    assert cut(fa, 1) == fa.rindex(b">c") and cut(fa[:11], 1) == 0 and cut(fa[:12], 1) == 11 and cut(fa[:14], 1) == 11
    assert cut(fa, 1, final=True) == len(fa)
    assert cut(b"", 2) == 0 and cut(fq, 0) == 0


def test_renumber_text_host_helper_equals_python_merge():
    """mtg_renumber_text (host-only entry point used by the N-GPU merge) shifts the shared bkpt ids exactly like dist.renumber."""
    from mindthegap_b200.api import renumber_text
    from mindthegap_b200.dist import renumber
    from tests.cases import expected
    for name in ("syn_small_k31", "full", "full_k63", "one_snp", "n_in_stretch"):
        bk, vcf, _ = expected(name)
        for off in (0, 7, 123456):
            b2, v2, used = renumber(bk, vcf, off)
            nb, m1 = renumber_text(bk, 0, off)
            nv, m2 = renumber_text(vcf, 1, off)
            assert nb == b2 and nv == v2
            assert max(m1, m2) == (used + off if used else 0)
