"""Sweep of the deterministic-schedule knobs (gds_params) on config 4 (50M reads / 5 Mb / M=500)
and config 2-like data: K3 milliseconds, rounds of the slowest component, kept reads."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

pkg = load_package()
wname = sys.argv[1] if len(sys.argv) > 1 else "c4"
SHAPES = ("uniform1000", "low_sides", "hole", "zero_sides")  # coverage_tester.cpp:120-175
if wname == "c3":  # config 3: 100 k reads over 30 kb, M=100 (same law as config 1)
    wl = dict(bench.WORKLOADS["c1"], pairs=50_000)
elif wname in SHAPES:  # the reference's own random test cases: 2 M reads / 30 kb, M = 8000 (1000)
    from __graft_entry__ import load_oracle
    wl = dict(bench.WORKLOADS["c1"], pairs=1_000_000, M=1000 if wname == "uniform1000" else 8000)
else:
    wl = bench.WORKLOADS[wname]
if wname in SHAPES:
    s_, e_, _, _ = load_oracle().gen_reads(12345, 1_000_000, 30_000, 150,
                                           "uniform" if wname == "uniform1000" else wname)
    st, en, fx = torch.from_numpy(s_.view(np.int32)), torch.from_numpy(e_.view(np.int32)), None
else:
    st, en, _, fx = bench.generate(wl, [0], pinned=False)
dev = torch.device("cuda", 0)
d_s, d_e = st.to(dev), en.to(dev)
n = d_s.numel()
bm = torch.zeros(n // 32 + 4, dtype=torch.int32, device=dev)
solver = pkg.Solver(0)


def run(prm):
    best = None
    for _ in range(3):
        r = solver.solve_device(d_s.data_ptr(), d_e.data_ptr(), n, [wl["L"]], wl["M"], bm.data_ptr(),
                                params=prm, len_hint=(wl["R"], wl["R"]) if fx is None else None,
                                verify=True)
        assert r.verify_violations == 0 and r.flow_value == r.fstar
        if best is None or r.ms_maxflow < best.ms_maxflow:
            best = r
    return best


base = None
segs = (32768, 16384, 8192, 65536) if wl["L"] > 65536 else (0,)
for seg in segs:
    for imin, lpct, rpct in ((64, 150, 1), (32, 100, 1), (16, 50, 1), (64, 50, 1), (128, 300, 1),
                             (64, 150, 0), (64, 150, 5), (32, 50, 0), (256, 400, 1), (16, 25, 0),
                             (8, 25, 0), (1000000, 150, 1)):
        r = run((imin, lpct, rpct, 0, seg))
        print("seg %6d interval_min %4d levels_pct %4d relabel_pct %d: K3 %7.3f ms total %7.3f ms "
              "comps %4d rounds_max %5d grs %4d bfs_levels %6d kept %d" %
              (seg, imin, lpct, rpct, r.ms_maxflow, r.ms_total, r.n_components, r.rounds_max,
               r.global_relabels, r.bfs_levels, r.n_kept), flush=True)
