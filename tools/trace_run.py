"""Debug helper (GPU box): per-find wall clock and memory-pool state over back-to-back finds."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import bench, mindthegap_b200 as m
wl = bench.make_workload()
host = torch.from_numpy(wl["stream"]).pin_memory(); dev = host.cuda()
ref_stream = np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in wl["refs"]])
n = int(host.numel())
p = m.FindParams(kmer_size=31)
for mode in ("resident", "host", "resident"):
    for it in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        f = m.Finder(p); f.reserve(n)
        if mode == "resident": f.push_reads_device(dev.data_ptr(), n)
        else: f.push_reads(host.numpy())
        f.finish_count()
        f.set_reference(ref_stream)
        for name, seq in wl["refs"]: f.scan_reference(name, seq)
        st = f.stats()
        f.close()
        t1 = time.perf_counter()
        print("%s find %7.3f ms  push %.2f finish %.2f setref %.2f scan %.2f  pool reserved %.0f MB used %.0f MB" % (
            mode, (t1 - t0) * 1e3, st["api.ms_push_reads"], st["api.ms_count_finish"], st["api.ms_set_reference"], st["api.ms_scan_reference"],
            st["mem.arena_cached_mb"], st["mem.arena_live_mb"]), file=sys.stderr)
