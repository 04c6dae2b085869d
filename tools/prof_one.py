"""Profiling driver (GPU box, run under ncu): N resident finds on cfg2 (scaled by argv[2]), nothing else.
usage: python tools/prof_one.py [n_finds=2] [scale=1.0] [k=31] [config=cfg2]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import bench
import mindthegap_b200 as m

n_finds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
k = int(sys.argv[3]) if len(sys.argv) > 3 else 31
config = sys.argv[4] if len(sys.argv) > 4 else "cfg2"
bench.K = k
wl = bench.make_workload(scale=scale, config=config)
dev = torch.from_numpy(wl["stream"]).cuda()
ref_stream = np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in wl["refs"]])
n = int(dev.numel())
p = m.FindParams(kmer_size=k)
for it in range(n_finds):
    f = m.Finder(p)
    f.reserve(n)
    f.push_reads_device(dev.data_ptr(), n)
    f.finish_count()
    f.set_reference(ref_stream)
    for name, seq in wl["refs"]:
        f.scan_reference(name, seq)
    st = f.stats()
    f.close()
print("prof_one: %d finds, nb_solid %d, launches/find %d" % (n_finds, st["count.nb_candidates"], st["count.launches"] + st["graph.launches"]))
