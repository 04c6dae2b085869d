"""Profiling driver for the GPU text ingest (csrc/ingest.cu): the cfg2 reads as 4-line FASTQ text, pushed twice.
Run under `ncu --metrics gpu__time_duration.sum -k regex:ig_` for the per-kernel launch list."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import mindthegap_b200 as m  # noqa: E402

wl = bench.make_workload(scale=float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
L = wl["read_len"]
rows = wl["stream"].reshape(-1, L + 1)
fq = np.empty((rows.shape[0], 2 * L + 7), dtype=np.uint8)
fq[:, 0] = ord("@"); fq[:, 1] = ord("r"); fq[:, 2] = 10
fq[:, 3:3 + L + 1] = rows
fq[:, L + 4] = ord("+"); fq[:, L + 5] = 10
fq[:, L + 6:2 * L + 6] = ord("I"); fq[:, 2 * L + 6] = 10
dev = torch.from_numpy(fq.reshape(-1)).cuda()
for rep in range(2):
    f = m.Finder(m.FindParams(kmer_size=31))
    f.push_reads_text_device(dev.data_ptr(), dev.numel(), 2)
    st = f.stats()
    print("ingest: %.3f ms for %d bytes -> %d bytes, %d sequences" % (st["ingest.ms"], st["ingest.bytes_in"], st["ingest.bytes_out"], st["ingest.nb_sequences"]))
    f.close()
