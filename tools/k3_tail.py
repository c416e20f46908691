"""K3 tail probe: distribution of per-component clocks on a segmented workload, for several
schedule-parameter variants.  usage: python tools/k3_tail.py c4 "seg=16384" "seg=16896,gri=32" ...
keys: gri (gr_interval_min), grl (gr_levels_pct), grr (gr_relabel_pct), seg (seg_len)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

pkg = load_package()
wname = sys.argv[1]
variants = sys.argv[2:] or [""]
wl = bench.WORKLOADS[wname]
st, en, _, fx = bench.generate(wl, [0], pinned=False)
dev = torch.device("cuda", 0)
d_s, d_e = st.to(dev), en.to(dev)
n = d_s.numel()
off = np.array([0, n], dtype=np.uint64)
bm = torch.zeros(n // 32 + 4, dtype=torch.int32, device=dev)
solver = pkg.Solver(0)
dump = "/tmp/k3_tail_dump.txt"
os.environ["GDS_DUMP_COMP"] = dump
for var in variants:
    kv = dict(x.split("=") for x in var.split(",") if x)
    prm = (int(kv.get("gri", 64)), int(kv.get("grl", 150)), int(kv.get("grr", 1)), 0, int(kv.get("seg", 0)))
    best = None
    for _ in range(3):
        r = solver.solve_device(d_s.data_ptr(), d_e.data_ptr(), n, [wl["L"]], wl["M"], bm.data_ptr(),
                                read_off=off, len_hint=(wl["R"], wl["R"]), params=prm)
        if best is None or r.ms_maxflow < best.ms_maxflow:
            best = r
    rows = np.loadtxt(dump, skiprows=1, ndmin=2)
    cyc = rows[:, 8]
    q = np.percentile(cyc, [0, 25, 50, 75, 90, 99, 100]) / 1e3
    print("%-28s K3 %6.3f ms comps %4d kept %9d seg %6d | kcyc min %6.0f q25 %6.0f med %6.0f q75 %6.0f q90 %6.0f q99 %6.0f max %6.0f"
          % (var or "default", best.ms_maxflow, best.n_components, best.n_kept, best.seg_len, *q), flush=True)
    order = np.argsort(-cyc)[:4]
    for i in order:
        c = rows[i]
        print("      comp %4d rounds %5d pushes %7d relabels %6d grs %2d levels %5d maxfront %4d kcyc %7.0f"
              % (c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[8] / 1e3), flush=True)
