"""GPU parity at the STATED sizes of BASELINE.json configs[1], [2], [4] (cfg2 and cfg3, k = 31 and k = 63) against outputs of
the unmodified reference binary (tests/golden/fullsize/*, made by tests/golden/make_fullsize_fixtures.py with oracle/_ref):
`.breakpoints`, VCF records, cut-off, the solid set (order-independent checksum of dsk/solid) and every byte of the Bloom /
cascading-Bloom arrays of the reference's .h5. Inputs are regenerated here by tools/synth.py (same generator and seed)."""
import hashlib
import os
import sys

import numpy as np
import pytest

from tests.fullsize import FULLSIZE, fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

_WORKLOADS = {}


def _workload(config, seed):
    """(read stream uint8 with '\\n' separators, [(name, uint8 array)]) -- kept for the k=31 and k=63 case of one config."""
    key = (config, seed)
    if key not in _WORKLOADS:
        _WORKLOADS.clear()   # one config resident at a time (cfg3 is 2 GB of reads)
        import synth
        cfg = synth.CONFIGS[config]
        refs, mats, _ = synth.reads_in_memory(cfg, seed)
        L = cfg["read_len"]
        tot = sum(m.shape[0] for m in mats)
        buf = np.empty((tot, L + 1), dtype=np.uint8)
        buf[:, L] = 10
        o = 0
        for m in mats:
            buf[o:o + len(m), :L] = m
            o += len(m)
        _WORKLOADS[key] = (buf.reshape(-1), refs)
    return _WORKLOADS[key]


def _mix64(x):
    x = x.copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd); x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53); x ^= x >> np.uint64(33)
    return x


def solid_checksum(lo, hi, ab):
    """The order-independent checksum oracle/ref_tools/h5solid.cpp prints for dsk/solid."""
    with np.errstate(over="ignore"):
        h = _mix64(lo ^ _mix64(hi + np.uint64(0x9E3779B97F4A7C15)))
        ms = int(np.sum(h * (ab.astype(np.uint64) + np.uint64(1)), dtype=np.uint64))
    return {"n": int(len(lo)), "xor_lo": "%016x" % int(np.bitwise_xor.reduce(lo)) if len(lo) else "0" * 16,
            "xor_hi": "%016x" % int(np.bitwise_xor.reduce(hi)) if len(hi) else "0" * 16, "mixsum": "%016x" % ms,
            "abundance_sum": int(ab.astype(np.uint64).sum())}


def _first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return "line %d: ours %r / reference %r" % (i + 1, x[:200], y[:200])
    return "length differs: %d vs %d lines" % (len(la), len(lb))


@pytest.mark.parametrize("name", list(FULLSIZE))
def test_fullsize_outputs_equal_reference_binary(name):
    fx_all = fixture(name)
    if fx_all is None:
        pytest.skip("no committed fixture for %s" % name)
    fx, bk_ref, vcf_ref = fx_all
    case = FULLSIZE[name]
    import mindthegap_b200 as m
    stream, refs = _workload(case["config"], case["seed"])
    f = m.Finder(m.FindParams.from_cli(["-kmer-size", str(case["k"])] + list(case["flags"])))
    bk, vcf = f.find(stream, refs)
    bk, vcf = bk.encode(), vcf.encode()
    info = "\n".join(fx["info"])
    assert ("abundance_min (auto inferred)            : %d" % f.cutoff_auto) in info, (f.cutoff_auto, info)
    assert f.nb_solid == fx["solid"]["n"]
    assert solid_checksum(*f.export_solid()) == fx["solid"]
    # every byte of the reference's .h5 Bloom datasets ("the golden bits are the .h5 datasets")
    for which, ds in enumerate(["/bloom/bloom", "/debloom/bloom2", "/debloom/bloom3", "/debloom/bloom4"]):
        ref_bits = fx["h5_bits"].get(ds)
        if ref_bits is None:
            continue
        bits = f.copy_bits(which).tobytes()
        assert len(bits) == ref_bits["bytes"], (ds, len(bits), ref_bits["bytes"])
        assert hashlib.sha256(bits).hexdigest() == ref_bits["sha256"], ds
    if fx["h5_bits"].get("/debloom/cfp"):
        ksz = 8 if case["k"] <= 31 else 16
        assert int(f.stats()["graph.cfp_set"]) * ksz == fx["h5_bits"]["/debloom/cfp"]["bytes"]
    f.close()
    assert hashlib.sha256(bk).hexdigest() == fx["breakpoints"]["sha256"], _first_diff(bk, bk_ref)
    assert hashlib.sha256(vcf).hexdigest() == fx["vcf"]["sha256"], _first_diff(vcf, vcf_ref)
