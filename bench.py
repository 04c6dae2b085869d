#!/usr/bin/env python
"""bench.py -- `MindTheGap find` hot path on B200 (contract: see DESIGN.md "Measurement").

One "step" = one whole `find` over the workload: count the read k-mers (solid set, auto threshold), build the
membership structures, scan every reference k-mer and replay the gap finders -> .breakpoints + .vcf text.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our CUDA engine through the C ABI (libmtg_b200.so)
  python bench.py --impl reference ...                        the UNMODIFIED reference binary (oracle/_ref/bin/MindTheGap find,
                                                              stock code path, -nb-cores = all host threads) on a bounded sample

Workload (default --config cfg3 = BASELINE.json configs[2], the largest single-GPU configuration: its 1 GB k-mer table leaves
the 126 MB L2): synthetic 64 Mbp genome, 30x 2x150 bp reads, 2000 HOM + 2000 HET insertions, 6400 SNPs, 1000 deletions, k=31,
generated deterministically in memory by tools/synth.py (seed 20241). --config cfg2 = configs[1] (4.6 Mbp, 50x). Under
torchrun (N>1) every rank brings the read set of its own genome (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

K = 31
SEED = 20241
METRIC = "read+reference k-mers processed per second by one whole find (count + graph + scan)"
UNIT = "kmers/s"


# ---------------------------------------------------------------------------------------------------- workload
_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, on the real stdout (fd saved by main before libraries could print to it)."""
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def scaled_config(config, scale=1.0, genome_mult=1):
    import synth
    cfg = dict(synth.CONFIGS[config])
    cfg["genome_len"] = int(cfg["genome_len"] * scale * genome_mult)
    for key in ("n_hom", "n_het", "n_snp", "n_del"):
        if cfg.get(key):
            cfg[key] = max(1, int(cfg[key] * scale * genome_mult))
    return cfg


def workload_name(config, cfg):
    return "%s: synthetic %.2f Mbp genome, %dx 2x%dbp reads, %d planted homozygous insertions%s, k=%d" % (
        config, cfg["genome_len"] / 1e6, cfg["coverage"], cfg["read_len"], cfg["n_hom"],
        (", %d heterozygous insertions, %d SNPs, %d deletions" % (cfg["n_het"], cfg["n_snp"], cfg["n_del"])) if cfg.get("n_het") else "", K)


def bench_config(args, world):
    """The `config` object of the JSON line: a function of the command line only, so that both arms print the same one."""
    return {"workload": workload_name(args.config, scaled_config(args.config, args.scale)), "kmer_size": K, "abundance_min": "auto",
            "scale": args.scale, "seed": SEED,
            "l2_policy": "inputs (the reads of one GPU) are larger than the 126 MB L2; every step starts from a fresh context",
            "parallelism": ("1 process per GPU (%d), weak scaling: every rank brings the reads of its own genome; records all-to-all by "
                            "minimizer owner, table ranges all-gathered, Bloom arrays OR-reduced (replica per GPU), reference scanned per rank" % world)
            if world > 1 else "single GPU"}


def make_workload(scale=1.0, seed=SEED, genome_mult=1, config="cfg2"):
    """cfg2 / cfg3 (optionally scaled): returns dict(refs=[(name, uint8 array)], stream=uint8 array of '\\n'-separated reads,
    mats, n_reads, read_len, read_kmers, ref_kmers)."""
    import synth
    cfg = scaled_config(config, scale, genome_mult)
    refs, mats, truth = synth.reads_in_memory(cfg, seed)
    L = cfg["read_len"]
    tot = sum(m.shape[0] for m in mats)
    buf = np.empty((tot, L + 1), dtype=np.uint8)
    buf[:, L] = 10
    o = 0
    for m in mats:
        buf[o:o + len(m), :L] = m
        o += len(m)
    return dict(cfg=cfg, refs=refs, stream=buf.reshape(-1), n_reads=tot, read_len=L, read_kmers=tot * max(0, L - K + 1),
                ref_kmers=sum(max(0, len(s) - K + 1) for _, s in refs), truth=truth,
                name=workload_name(config, cfg))


def read_fasta(path):
    """Host parse of a (multi-line) FASTA file: [(name, bytes)]."""
    out = []
    for rec in open(path, "rb").read().split(b">")[1:]:
        head, _, body = rec.partition(b"\n")
        out.append((head.split()[0].decode() if head.split() else "", body.replace(b"\n", b"")))
    return out


def write_inputs(wl, d):
    """Write the workload as FASTA files for the CPU implementation; returns (reads_uri, ref_path)."""
    import synth
    os.makedirs(d, exist_ok=True)
    L = wl["read_len"]
    rows = wl["stream"].reshape(-1, L + 1)
    out = np.empty((rows.shape[0], L + 4), dtype=np.uint8)
    out[:, 0] = ord(">"); out[:, 1] = ord("r"); out[:, 2] = 10
    out[:, 3:] = rows
    reads = os.path.join(d, "reads.fa")
    with open(reads, "wb") as f:
        f.write(out.tobytes())
    ref = os.path.join(d, "ref.fa")
    synth.write_fasta(ref, wl["refs"])
    return reads, ref


# ---------------------------------------------------------------------------------------------------- CPU implementation
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "MindTheGap")


def _find_outputs(out):
    bk = open(out + ".breakpoints").read()
    vcf = "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#"))
    return bk, vcf


def run_reference_find(wl, cores, workdir, tag="ref", split=False):
    """The UNMODIFIED reference (`oracle/_ref/bin/MindTheGap find`, built by oracle/build_ref.sh, stock code path) on `wl` with
    -nb-cores `cores`: wall clock of the whole process (file parsing, graph build, scan, output files). split=True adds a second
    run `find -graph <out>.h5` (scan only) to separate graph construction from the reference scan."""
    reads, ref = write_inputs(wl, workdir)
    out = os.path.join(workdir, tag)
    cmd = [REF_BIN, "find", "-in", reads, "-ref", ref, "-kmer-size", str(K), "-out", out, "-nb-cores", str(cores)]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    secs = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit("bench.py: reference binary failed: %s" % r.stderr[-500:])
    info = {}
    for l in r.stdout.replace("\r", "\n").splitlines():
        if ":" in l and not l.startswith("["):
            k, v = l.split(":", 1)
            info[k.strip()] = v.strip()
    bk, vcf = _find_outputs(out)
    res = dict(seconds=secs, value=(wl["read_kmers"] + wl["ref_kmers"]) / secs, breakpoints=bk, vcf=vcf, kind="reference",
               nb_solid=int(info.get("nb_solid_kmers", -1)), cutoff=int(info.get("abundance_min (auto inferred)", -1)))
    if split:
        t0 = time.perf_counter()
        r2 = subprocess.run([REF_BIN, "find", "-graph", out + ".h5", "-ref", ref, "-out", out + "_scan", "-nb-cores", str(cores)], cwd=workdir,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        scan = time.perf_counter() - t0
        if r2.returncode == 0:
            res["seconds_scan"] = scan
            res["seconds_graph"] = max(secs - scan, 1e-9)
    return res


def run_port_find(wl, cores, workdir, tag="cpu"):
    """The oracle port (oracle_find, a restatement of the reference's algorithm) -- only used when oracle/_ref is absent."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    exe = os.path.join(ROOT, "oracle", "_build", "oracle_find")
    reads, ref = write_inputs(wl, workdir)
    out = os.path.join(workdir, tag)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "find", "-in", reads, "-ref", ref, "-kmer-size", str(K), "-out", out, "-nb-cores", str(cores)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    secs = time.perf_counter() - t0
    info = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())
    bk, vcf = _find_outputs(out)
    return dict(seconds=secs, value=(wl["read_kmers"] + wl["ref_kmers"]) / secs, breakpoints=bk, vcf=vcf, kind="port",
                nb_solid=int(info["nb_solid"]), cutoff=int(info.get("cutoff_auto", -1)),
                seconds_graph=float(info["time_count"]) + float(info["time_graph"]), seconds_scan=float(info["time_refbloom"]) + float(info["time_scan"]))


def run_cpu_find(wl, cores, workdir, tag="cpu", split=False):
    if os.path.exists(REF_BIN) and not os.environ.get("MTG_BENCH_PORT"):
        return run_reference_find(wl, cores, workdir, tag, split)
    return run_port_find(wl, cores, workdir, tag)


def cpu_sample_scale(args):
    """Genome scale of the bounded CPU sample: --cpu-scale of cfg2's 4.6 Mbp, i.e. the same number of bases for every config."""
    import synth
    return min(args.scale, args.cpu_scale * 4.6e6 / synth.CONFIGS[args.config]["genome_len"])


def cpu_sample_text(swl, args, r):
    return ("%s (a %.3g-scale sample of the workload's generator, same coverage / read length / variant densities); one whole `find` "
            "process: FASTA parsing + graph construction + reference scan + output files, wall clock %.2f s" % (swl["name"], cpu_sample_scale(args), r["seconds"]))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    wl = make_workload(scale=cpu_sample_scale(args), config=args.config)
    times = []
    with tempfile.TemporaryDirectory() as tmp:
        for i in range(args.warmup + args.steps):
            r = run_cpu_find(wl, cores, tmp, split=(i == args.warmup + args.steps - 1))
            if i >= args.warmup:
                times.append(r["seconds"])
    t = sum(times) / len(times)
    value = (wl["read_kmers"] + wl["ref_kmers"]) / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64" if K <= 31 else "u128", "data": "synthetic", "config": bench_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": cpu_sample_text(wl, args, r)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if "seconds_graph" in r:
        line["read_kmers_counted_per_s"] = wl["read_kmers"] / r["seconds_graph"]
        line["ref_kmers_queried_per_s"] = wl["ref_kmers"] / r["seconds_scan"]
        line["split"] = {"graph_construction_s": r["seconds_graph"], "reference_scan_s": r["seconds_scan"],
                         "how": "second run `find -graph <out>.h5` = scan only (single-threaded by construction); graph = total - scan"}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / power / throttle reasons of one GPU during the timed region (NVML in a thread, 50 ms period)."""

    def __init__(self, index=0, period=0.2):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.stop_flag = threading.Event()
        self.th = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = "nvml unavailable: %s" % e
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                r = int(get_reasons(self.h))
                self.samples.append((sm, pw))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                break
            self.stop_flag.wait(self.period)

    def stop(self):
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not sampled"]}
        self.stop_flag.set()
        self.th.join(timeout=2)
        sm = [x[0] for x in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "power_w_max": max(x[1] for x in self.samples) if self.samples else None, "samples": len(sm),
                "reasons": sorted(self.reasons), "how": "NVML, %d ms period, during the timed resident steps" % int(self.period * 1e3)}


# ---------------------------------------------------------------------------------------------------- our arm
def own_arm(args):
    import torch
    import torch.distributed as dist

    import mindthegap_b200 as m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = m.load_library()

    # every rank brings the read set of its own genome segment (weak scaling); N=1 is exactly cfg2
    wl = make_workload(scale=args.scale, seed=SEED + 1000 * rank, config=args.config)
    params = m.FindParams(kmer_size=K, device=local_rank)
    if world > 1:   # every find of this process runs on one persistent stream shared with torch / NCCL (dist.shared_stream)
        from mindthegap_b200.dist import shared_stream
        params.stream = shared_stream(torch.device("cuda", local_rank)).cuda_stream
    stream_host = torch.from_numpy(wl["stream"]).pin_memory()
    stream_np = stream_host.numpy()
    ref_stream = np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in wl["refs"]])
    ref_host = torch.from_numpy(ref_stream).pin_memory()
    stream_dev = stream_host.cuda()
    ref_dev = ref_host.cuda()
    ref_offsets = np.cumsum([0] + [len(s) + 1 for _, s in wl["refs"]])
    nbytes = int(stream_np.size)

    all_refs = None
    if world > 1:   # every rank needs every chromosome (features are sharded by position, not by chromosome)
        import synth
        from mindthegap_b200.dist import DistFind
        all_refs = []
        for r in range(world):
            refs_r = wl["refs"] if r == rank else synth.build(wl["cfg"], SEED + 1000 * r)[0]
            all_refs += [("g%d_%s" % (r, n), s) for n, s in refs_r]
        wl["ref_kmers"] = sum(max(0, len(s) - K + 1) for _, s in wl["refs"])
        all_ref_stream = np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in all_refs])
        all_ref_host = torch.from_numpy(all_ref_stream).pin_memory()   # e2e: pinned host buffer, copied inside the timed region
        all_ref_dev = all_ref_host.cuda()                              # resident: already in HBM

    def one_find(resident):
        f = m.Finder(params)
        f.reserve(nbytes)
        if world > 1:
            d = DistFind(f, torch.device("cuda", local_rank))
            if resident:
                d.push_reads(dev_ptr=stream_dev.data_ptr(), nbytes=nbytes)
            else:
                d.push_reads(stream_np)
        elif resident:
            f.push_reads_device(stream_dev.data_ptr(), nbytes)
        else:
            f.push_reads(stream_np)
        if world > 1:
            bk, vcf = d.find(all_refs, all_ref_dev if resident else all_ref_host)
            if d.trace is not None and rank == 0:
                print("[dist trace] " + json.dumps({k: [v[0], round(v[1], 3)] for k, v in sorted(d.trace.items(), key=lambda kv: -kv[1][1])}), file=sys.stderr)
            st = f.stats()
            st["nb_solid"] = d.nb_solid
            st["threshold"] = f.threshold
            st.update({"find." + k: v for k, v in f.find_counters().items()})
            st.update({"exchange." + k: float(v) for k, v in d.exchange_bytes.items()})
            st.update({"dist.ms_" + k: float(v) for k, v in d.timing.items()})
            f.close()
            return bk, vcf, st
        f.finish_count()
        if resident:
            f.set_reference_device(ref_dev.data_ptr(), int(ref_stream.size))
        else:
            f.set_reference(ref_host.numpy())
        for i, (name, seq) in enumerate(wl["refs"]):
            if resident:
                f.scan_reference_device(name, seq, ref_dev.data_ptr() + int(ref_offsets[i]))
            else:
                f.scan_reference(name, seq)
        bk, vcf = f.breakpoints_text(), f.vcf_text()
        st = f.stats()
        st["nb_solid"] = f.nb_solid
        st["threshold"] = f.threshold
        st.update({"find." + k: v for k, v in f.find_counters().items()})
        f.close()
        return bk, vcf, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, steps, warmup):
        for _ in range(warmup):
            out = one_find(resident)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        stats = []
        for _ in range(steps):
            out = one_find(resident)
            stats.append(out[2])
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, wall / steps, out, stats

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, wall_res, out_res, stats = timed(True, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, wall_e2e, out_e2e, _ = timed(False, args.steps, max(1, args.warmup // 2))
    assert out_res[0] == out_e2e[0] and out_res[1] == out_e2e[1], "resident and host-buffer runs disagree"
    if world > 1 and args.verify:   # N-rank outputs == one-GPU outputs on the union of the inputs (rank 0 does the single-GPU run)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(wl["stream"], gathered, dst=0)
        if rank == 0:
            f = m.Finder(params)
            for part in gathered:
                f.push_reads(part)
            bk1, vcf1 = None, None
            f.finish_count()
            f.set_reference(np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in all_refs]))
            for name, seq in all_refs:
                f.scan_reference(name, seq)
            bk1, vcf1 = f.breakpoints_text(), f.vcf_text()
            f.close()
            assert bk1 == out_res[0] and vcf1 == out_res[1], "N-GPU outputs differ from the single-GPU outputs on the same inputs"
            print("[bench verify] %d-GPU outputs identical to the single-GPU run (%d breakpoint lines)" % (world, len(bk1.splitlines())), file=sys.stderr)
    if args.trace and world == 1:  # one extra resident find with the library's stage trace on stderr
        os.environ["MTG_TRACE"] = "1"
        t0 = time.perf_counter()
        one_find(True)
        print("[bench trace] resident find %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
        os.environ["MTG_TRACE"] = "0"

    tot_kmers = wl["read_kmers"] + wl["ref_kmers"]
    if world > 1:
        t = torch.tensor([wl["read_kmers"], wl["ref_kmers"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        read_kmers, ref_kmers = float(t[0].item()), float(t[1].item())
        tot_kmers = read_kmers + ref_kmers
    else:
        read_kmers, ref_kmers = float(wl["read_kmers"]), float(wl["ref_kmers"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- per-kernel roofline: CUDA-event times measured inside the library on its own stream (averaged over the timed steps),
    # each kernel against ITS OWN algorithmic bytes (DESIGN.md section 3 states every figure):
    #   streaming kernels against the measured HBM copy peak (MEASURED_PEAKS.json), random-access kernels against the random
    #   128-byte gather peak measured live (mtg_bench_random_gather over an 8 GB table)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    gather_peak = lib.mtg_bench_random_gather(local_rank, 8 << 30, 1 << 26, 3)
    avg = {k: float(np.mean([s[k] for s in stats])) for k in stats[0]}
    nk = wl["read_kmers"]
    ksz = 8 if K <= 31 else 16
    bpk = 2 * ksz + 0.25
    nsolid = avg["nb_solid"]
    nhash = int(0.7 * avg["graph.bloom_bits"] / max(nsolid, 1) + 1e-6) if nsolid else 4
    count_stage_ms = avg["count.ms_pack"] + avg["count.ms_extract"] + avg["count.ms_group"] + avg["count.ms_scatter"] + avg["count.ms_count"] + avg["count.ms_filter"]
    probes = avg["scan.table_probes"]
    G, H = "gather", "hbm"
    kernels = [
        {"kernel": "count stage (pack+superkmer+group+scatter+count+filter)", "ms": count_stage_ms, "bytes": bpk * nk, "bound": H,
         "note": "SURVEY 8d stage figure: 2*sizeof(kmer)+0.25 B per read k-mer instance (the stage, not one kernel)"},
        {"kernel": "count_kernel_dd" if K <= 31 else "count_kernel", "ms": avg["count.ms_count"], "bytes": ksz * nk + 8.0 * avg["count.nb_records"], "bound": H,
         "note": "sizeof(kmer) per k-mer instance read from the partition + the 8-byte super-k-mer records (SURVEY 8d's per-unit figure; "
                 "k <= 31: the de-duplicating kernel streams 32-byte records that carry their bases, identical super-k-mers are folded "
                 "before expansion)"},
        {"kernel": "superkmer_kernel", "ms": avg["count.ms_extract"], "bytes": 0.375 * nbytes + 8.0 * avg["count.nb_records"], "bound": H,
         "note": "packed bases + invalid mask in, 8-byte records out"},
        {"kernel": "scatter_kernel", "ms": avg["count.ms_scatter"], "bytes": (40.0 if K <= 31 else 16.0) * avg["count.nb_records"], "bound": H,
         "note": "8-byte records read; written once with their bases (one 32-byte sector per record) for k <= 31, as 8-byte records else; "
                 "the kernel is bound by one L2 atomic per record, not by these bytes"},
        {"kernel": "pack_kernel", "ms": avg["count.ms_pack"], "bytes": 1.375 * nbytes, "bound": H, "note": "ASCII in, 2-bit words + mask out"},
        {"kernel": "filter_kernel", "ms": avg["count.ms_filter"], "bytes": (ksz + 4.0) * (avg["count.nb_candidates"] + nsolid), "bound": H,
         "note": "candidates read, solid set written"},
        {"kernel": "table_build_kernel", "ms": avg["graph.ms_table"], "bytes": (ksz + 128.0) * nsolid, "bound": G,
         "note": "key read + one 128-B bucket per solid k-mer"},
        {"kernel": "bloom_neighbor_insert_kernel", "ms": avg["graph.ms_bloom"], "bytes": (ksz + 32.0 * nhash) * nsolid, "bound": G,
         "note": "key read + nhash 32-B sectors set inside one 512-B window"},
        {"kernel": "critical_kernel", "ms": avg["graph.ms_critical"], "bytes": (ksz + 2 * 32.0 * nhash + 128.0 * 3.3) * nsolid, "bound": G,
         "note": "per solid k-mer: 2 Bloom windows x nhash 32-B sectors + ~2.3 neighbour buckets + own bucket (adjacency byte)"},
        {"kernel": "cascade (bloom_cache_insert/bloom_cascade/cfp_set kernels)", "ms": avg["graph.ms_cascade"],
         "bytes": (2 * ksz + 3 * 32.0) * nsolid, "bound": G, "note": "solid set read twice, ~3 Bloom sectors per k-mer"},
        {"kernel": "mphf_level/clear/compact kernels", "ms": avg["graph.ms_mphf"], "bytes": 1.39 * (2 * ksz + 64.0) * nsolid, "bound": G,
         "note": "surviving keys (sum over levels 1.39 N) read twice, one 32-B sector RMW + one sector read each; on one GPU these kernels "
                 "run on a side stream under critical_kernel: ms = first launch to last completion THERE, stretched by sharing the SMs "
                 "(alone, MTG_MPHF_SERIAL=1: 2.1 ms = 0.98 of the gather peak on cfg3); exposed_ms = what the step still waits for"},
        {"kernel": "features_kernel (probe)", "ms": avg["scan.ms_features"], "bytes": 128.0 * probes, "bound": G,
         "note": "128 B x exact-table probes actually issued (one per valid position thanks to the adjacency byte; SURVEY 8d counts R + 8 R_solid)"},
    ]
    for kq in kernels:
        pk = hbm if kq["bound"] == H else gather_peak
        kq["achieved_gbs"] = kq["bytes"] / (kq["ms"] * 1e-3) / 1e9 if kq["ms"] > 0 else None
        kq["peak_gbs"] = pk
        kq["frac"] = kq["achieved_gbs"] / pk if kq["achieved_gbs"] and pk > 0 else None
    # dominant single kernel = the longest one of the step; its ncu DRAM traffic per launch comes from the committed `ncu --set full`
    # capture of the same workload (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); null when none exists
    # BooPHF levels run on a side stream under critical_kernel on one GPU: their wall duration there is not a kernel time
    mq = [q for q in kernels if q["kernel"].startswith("mphf_level")][0]
    if "graph.ms_mphf_exposed" in avg and avg["graph.ms_mphf_exposed"] < 0.5 * avg["graph.ms_mphf"]:
        mq["overlapped_with"] = "critical_kernel (side stream)"
        mq["exposed_ms"] = avg["graph.ms_mphf_exposed"]
    dom = max([q for q in kernels[1:] if "overlapped_with" not in q], key=lambda q: q["ms"])
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tj.get("%s_k%d" % (args.config, K), {}) if args.scale == 1.0 else {}
        for kq in kernels[1:]:
            kq["ncu_dram_bytes_per_launch"] = ent.get(kq["kernel"].split(" ")[0])
        traffic = dom.get("ncu_dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    fk = kernels[-1]
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": dom["peak_gbs"], "unit": "GB/s",
                "frac": dom["frac"], "traffic": traffic,
                "peak_source": peak_src if dom["bound"] == H else "random 128-B gather over an 8 GB table, measured live (mtg_bench_random_gather)",
                "ms_per_launch": dom["ms"], "algorithmic_bytes_per_launch": dom["bytes"],
                "note": "integer hashing / probing: the dominant kernel is named with its own algorithmic bytes; kernels[] holds every kernel of "
                        "the step (streaming ones against the HBM copy peak, random-access ones against the measured gather peak)",
                "frac_above_one": ("the algorithmic probe bytes of this kernel are mostly served from L1/L2 (k-mers are placed by minimizer "
                                   "bin, the Bloom fits L2): `traffic` holds the DRAM bytes ncu measured for the same launch"
                                   if dom["frac"] and dom["frac"] > 1.0 else None),
                "hbm_copy_peak_gbs": hbm, "random_128B_gather_peak_gbs": gather_peak,
                "count_stage": {"ms": kernels[0]["ms"], "achieved": kernels[0]["achieved_gbs"], "frac": kernels[0]["frac"]},
                "probe_frac_of_gather_peak": fk["frac"]}

    # ---- size-independent property at FULL size (the oracle only covers the bounded sample below): the planted variants are found
    truth_check = None
    if world == 1:
        import synth
        truth_check = synth.truth_recall(out_res[0], out_res[1], wl["truth"], [n for n, _ in wl["refs"]])
        if K <= 31 and truth_check["INS"]["planted"] and truth_check["INS"]["recall"] < 0.8:
            raise SystemExit("bench.py: only %d of %d planted insertions have a breakpoint" % (truth_check["INS"]["recovered"], truth_check["INS"]["planted"]))

    # ---- text ingest: the same reads as 4-line FASTQ text, parsed on the GPU (csrc/ingest.cu). Reported beside
    # the find step, not inside it: the bench line's host buffers are already-parsed bases, like the CPU arm whose parse time is excluded.
    ingest = None
    if world == 1:
        L = wl["read_len"]
        rows = wl["stream"].reshape(-1, L + 1)[:1600000]      # bounded: at most 1.6 M reads (cfg2's whole read set)
        fq = np.empty((rows.shape[0], 2 * L + 7), dtype=np.uint8)
        fq[:, 0] = ord("@"); fq[:, 1] = ord("r"); fq[:, 2] = 10
        fq[:, 3:3 + L + 1] = rows
        fq[:, L + 4] = ord("+"); fq[:, L + 5] = 10
        fq[:, L + 6:2 * L + 6] = ord("I"); fq[:, 2 * L + 6] = 10
        fq_host = torch.from_numpy(fq.reshape(-1)).pin_memory()
        fq_dev = fq_host.cuda()
        ms_dev, ms_host, ing = [], [], None
        for rep in range(4):
            for resident in (True, False):
                f = m.Finder(params)
                f.reserve(int(rows.size))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if resident:
                    f.push_reads_text_device(fq_dev.data_ptr(), fq_dev.numel(), 2)
                else:
                    f.push_reads_text(fq_host.numpy(), 2)
                dt = (time.perf_counter() - t0) * 1e3
                st_i = f.stats()
                if rep:   # first repetition warms the arena
                    (ms_dev if resident else ms_host).append(dt)
                    if resident:
                        ing = st_i
                f.close()
        alg = float(ing["ingest.bytes_in"] + ing["ingest.bytes_out"])
        ingest = {"workload": "%d of the step's reads as 4-line FASTQ text (%d bytes)" % (rows.shape[0], fq_dev.numel()), "parse_ms": ing["ingest.ms"],
                  "parse_gbs_text": ing["ingest.bytes_in"] / (ing["ingest.ms"] * 1e-3) / 1e9,
                  "algorithmic_bytes": alg, "achieved_gbs": alg / (ing["ingest.ms"] * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ing["ingest.ms"] * 1e-3) / 1e9 / hbm,
                  "sequences": int(ing["ingest.nb_sequences"]), "push_text_resident_ms": float(np.median(ms_dev)),
                  "push_text_from_pinned_host_ms": float(np.median(ms_host)),
                  "note": "parse = 3 streaming passes + 2 scans (text read 3x, bases written once); push_text = parse + pack + super-k-mers; "
                          "from host adds the H2D copy of the text"}
        del fq_dev, fq_host, fq
        kernels.append({"kernel": "ingest (FASTQ text -> base stream, 7 launches)", "ms": ing["ingest.ms"], "bytes": alg,
                        "note": "text read once + bases written once (algorithmic); the passes re-read the text 3x",
                        "achieved_gbs": ingest["achieved_gbs"], "frac": ingest["frac_of_hbm_peak"]})

    # ---- parity at FULL size: the step's outputs against the committed outputs of the unmodified reference binary on the same
    # inputs (tests/golden/fullsize/<config>_k<K>.json, made by tests/golden/make_fullsize_fixtures.py with oracle/_ref)
    parity_full = None
    if world == 1 and args.scale == 1.0:
        import hashlib
        from tests.fullsize import fixture
        fx_all = fixture("%s_k%d" % (args.config, K))
        if fx_all is not None:
            fx = fx_all[0]
            info = "\n".join(fx["info"])
            parity_full = {"fixture": "tests/golden/fullsize/%s_k%d.json (reference binary, %s)" % (args.config, K, fx["command"]),
                           "breakpoints_equal": hashlib.sha256(out_res[0].encode()).hexdigest() == fx["breakpoints"]["sha256"],
                           "vcf_equal": hashlib.sha256(out_res[1].encode()).hexdigest() == fx["vcf"]["sha256"],
                           "nb_solid_equal": int(avg["nb_solid"]) == fx["solid"]["n"],
                           "cutoff_equal": ("abundance_min (auto inferred)            : %d" % int(avg["threshold"])) in info,
                           "breakpoint_records": fx["breakpoints"]["records"], "vcf_records": fx["vcf"]["records"]}
            if not all(v for k, v in parity_full.items() if k.endswith("_equal")):
                raise SystemExit("bench.py: outputs differ from the reference binary's at full size: %s" % parity_full)

    # ---- end to end FROM FILES (what the `find` CLI does): FASTQ text files -> count_files (gzread + GPU parse) -> graph ->
    # reference FASTA parsed on the host -> scan -> texts. Files sit in the page cache (written just before), like the CPU arm's.
    e2e_files = None
    if world == 1 and args.files_steps > 0 and nbytes < (3 << 30):
        import synth
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
            L = wl["read_len"]
            rows = wl["stream"].reshape(-1, L + 1)
            half = rows.shape[0] // 2
            paths = []
            for i, part in enumerate((rows[:half], rows[half:])):
                pth = os.path.join(tmp, "r%d.fq" % (i + 1))
                with open(pth, "wb") as fh:
                    for o in range(0, part.shape[0], 1 << 20):
                        fh.write(synth.fastq_block(part[o:o + (1 << 20), :L], o, "p"))
                paths.append(pth)
            refp = os.path.join(tmp, "ref.fa")
            synth.write_fasta(refp, wl["refs"])
            fsz = sum(os.path.getsize(x) for x in paths) + os.path.getsize(refp)
            tms = []
            for it in range(args.files_steps + 1):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                f = m.Finder(params)
                f.count_files(",".join(paths))
                f.finish_count()
                rrecs = read_fasta(refp)
                f.set_reference(b"\n".join(sq for _, sq in rrecs))
                for name, sq in rrecs:
                    f.scan_reference(name, sq)
                bkf, vcff = f.breakpoints_text(), f.vcf_text()
                f.close()
                if it:
                    tms.append(time.perf_counter() - t0)
            assert bkf == out_res[0] and vcff == out_res[1], "from-files run disagrees with the in-memory run"
            e2e_files = {"value": tot_kmers / float(np.median(tms)), "unit": UNIT, "seconds": float(np.median(tms)), "file_bytes": int(fsz),
                         "steps": args.files_steps,
                         "how": "2 FASTQ files + reference FASTA in the page cache -> mtg_count_files (zlib read into pinned chunks, parsed and "
                                "packed on the GPU) -> graph -> scan -> texts; host wall clock of the whole find"}

    # ---- CPU baseline on a bounded sample (the unmodified reference binary when oracle/_ref exists) + parity on that sample
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        cpu_scale = cpu_sample_scale(args)
        swl = wl if cpu_scale >= args.scale else make_workload(scale=cpu_scale, config=args.config)
        with tempfile.TemporaryDirectory() as tmp:
            r = run_cpu_find(swl, cores, tmp, split=True)
        cpu = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": cpu_sample_text(swl, args, r)}
        if "seconds_graph" in r:
            cpu["read_kmers_counted_per_s"] = swl["read_kmers"] / r["seconds_graph"]
            cpu["ref_kmers_queried_per_s"] = swl["ref_kmers"] / r["seconds_scan"]
        # parity on the same sample: GPU outputs vs the CPU implementation's
        f = m.Finder(params)
        bk, vcf = f.find(swl["stream"], [(n, s) for n, s in swl["refs"]])
        nsolid_s, cut_s = f.nb_solid, f.cutoff_auto
        f.close()
        parity = {"against": r["kind"], "sample_breakpoints_equal": bk == r["breakpoints"], "sample_vcf_equal": vcf == r["vcf"],
                  "sample_nb_solid_equal": nsolid_s == r["nb_solid"], "sample_cutoff_equal": cut_s == r["cutoff"],
                  "breakpoint_records": len(bk.splitlines()) // 4, "vcf_records": len(vcf.splitlines())}
        if not all(v for k, v in parity.items() if k.endswith("_equal")):
            raise SystemExit("bench.py: GPU outputs differ from the CPU implementation (%s) on the sample: %s" % (r["kind"], parity))

    launches = int(sum(s["count.launches"] + s["graph.launches"] for s in stats))
    line = {"metric": METRIC, "value": tot_kmers / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64" if K <= 31 else "u128", "data": "synthetic",
            "config": bench_config(args, world), "abundance_min_inferred": int(avg["threshold"]), "per_gpu_read_bytes": nbytes,
            "e2e": {"value": tot_kmers / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(nbytes + (all_ref_stream.size if world > 1 else ref_stream.size)),
                    "d2h_bytes_per_step": int(2 * wl["ref_kmers"] + len(out_e2e[0]) + len(out_e2e[1])), "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "parity": parity,
            "parity_full": parity_full, "e2e_from_files": e2e_files,
            "read_kmers_counted_per_s": read_kmers / (1e-3 * count_stage_ms),
            "ref_kmers_queried_per_s": ref_kmers / (1e-3 * (avg["scan.ms_features"] + avg["scan.ms_replay"])),
            "find_wall_s": wall_e2e,
            "stage_ms": {k: avg[k] for k in sorted(avg) if ".ms_" in k},
            "counts": {"read_kmers": read_kmers, "ref_kmers": ref_kmers, "nb_solid": avg["nb_solid"], "table_probes": probes,
                       "bloom_emulations": avg["scan.bloom_emulations"], "observer_queries": avg["scan.observer_queries"],
                       "nb_records": avg["count.nb_records"], "nb_groups": avg["count.nb_groups"], "nb_items": avg["count.nb_items"],
                       "nb_multipass_groups": avg["count.nb_multipass_groups"], "nb_candidates": avg["count.nb_candidates"],
                       "prefetched_queries": avg["scan.prefetched_queries"], "unforeseen_queries": avg["scan.unforeseen_queries"],
                       "probe_batches": avg["scan.probe_batches"],
                       "breakpoint_records": len(out_res[0].splitlines()) // 4, "vcf_records": len(out_res[1].splitlines())}}
    if ingest is not None:
        line["ingest"] = ingest
    if truth_check is not None:
        line["truth_check"] = truth_check
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale of the GPU workload (1.0 = the named config)")
    ap.add_argument("--config", default="cfg3", choices=["cfg2", "cfg3", "cfg4s"],
                    help="BASELINE.json configs[2] (default, the bench line: largest single-GPU config), configs[1], or cfg4s = one GPU's eighth of configs[3]")
    ap.add_argument("--files-steps", type=int, default=2, help="timed repetitions of the from-files end-to-end figure (0 = skip)")
    ap.add_argument("--kmer-size", type=int, default=31, help="k (31 = the bench line; 63 exercises the 128-bit kernels, configs[4])")
    ap.add_argument("--cpu-scale", type=float, default=0.5, help="genome scale of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--verify", action="store_true", help="N>1: also run the union of the inputs on one GPU and compare the outputs")
    ap.add_argument("--trace", action="store_true", help="print the library's per-stage wall clock for one extra find (stderr)")
    args = ap.parse_args()
    global K
    K = args.kmer_size
    # stdout carries exactly ONE JSON line: everything libraries print (NCCL's version banner, ...) goes to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)
    return own_arm(args)


if __name__ == "__main__":
    sys.exit(main())
