"""K3 probe: per-component clock breakdown (set-up / first relabel / phase A / phase B) of the
max-flow kernel on the bench workloads, for a list of environment variants.
usage: python tools/k3_probe.py <workload> [samples] [VAR=val,VAR=val ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_oracle, load_package  # noqa: E402
import bench  # noqa: E402

pkg = load_package()
wname = sys.argv[1] if len(sys.argv) > 1 else "c1"
nsamp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
variants = sys.argv[3:] or [""]
SHAPES = ("uniform1000", "low_sides", "hole", "zero_sides")
if wname in SHAPES:
    O = load_oracle()
    wl = dict(bench.WORKLOADS["c1"], pairs=1_000_000, M=1000 if wname == "uniform1000" else 8000)
    s_, e_, _, _ = O.gen_reads(12345, 1_000_000, 30_000, 150, "uniform" if wname == "uniform1000" else wname)
    st, en, fx = torch.from_numpy(s_.view(np.int32)), torch.from_numpy(e_.view(np.int32)), None
    ids = [0]
else:
    wl = bench.WORKLOADS["c5" if wname == "c5" else wname]
    ids = list(range(nsamp))
    st, en, _, fx = bench.generate(wl, ids, pinned=False)
dev = torch.device("cuda", 0)
d_s, d_e = st.to(dev), en.to(dev)
n = d_s.numel()
per = n // len(ids)
off = np.arange(len(ids) + 1, dtype=np.uint64) * per
bm = torch.zeros(n // 32 + 4, dtype=torch.int32, device=dev)
solver = pkg.Solver(0)
dump = "/tmp/k3_dump.txt"
for var in variants:
    env = dict(kv.split("=") for kv in var.split(",") if kv)
    env["GDS_DUMP_COMP"] = dump
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    best = None
    for _ in range(4):
        r = solver.solve_device(d_s.data_ptr(), d_e.data_ptr(), n, [wl["L"]] * len(ids), wl["M"],
                                bm.data_ptr(), read_off=off,
                                len_hint=(wl["R"], wl["R"]) if fx is None else None)
        if best is None or r.ms_maxflow < best.ms_maxflow:
            best = r
    rows = np.loadtxt(dump, skiprows=1, ndmin=2)
    for k, v in old.items():
        if v is None:
            del os.environ[k]
        else:
            os.environ[k] = v
    cyc = rows[:, 8]
    print("%-40s K3 %7.3f ms comps %4d rounds_max %5d levels %6d | per comp (kcyc) total %8.1f (max %8.1f) "
          "setup %6.1f bfs1 %7.1f later_grs %7.1f (%d) phaseA %7.1f phaseB %7.1f | /round A %6.0f B %6.0f /level(all grs) %5.0f" %
          (var or "default", best.ms_maxflow, best.n_components, best.rounds_max, best.bfs_levels,
           cyc.mean() / 1e3, cyc.max() / 1e3, rows[:, 9].mean() / 1e3, rows[:, 10].mean() / 1e3,
           rows[:, 13].mean() / 1e3, rows[:, 4].max(), rows[:, 11].mean() / 1e3, rows[:, 12].mean() / 1e3,
           rows[:, 11].sum() / max(1, rows[:, 1].sum()), rows[:, 12].sum() / max(1, rows[:, 1].sum()),
           (rows[:, 10].sum() + rows[:, 13].sum()) / max(1, rows[:, 5].sum())), flush=True)
