import sys, numpy as np
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_oracle, load_package
pkg = load_package(); O = load_oracle()
solver = pkg.Solver(0)
for shape in ("uniform", "low_sides", "hole", "zero_sides"):
    s, e, q, l = O.gen_reads(12345, 1_000_000, 30_000, 150, shape)
    M = 1000 if shape == "uniform" else 8000
    for it in range(2):
        r = solver.solve(s.astype(np.uint16), None, 30_000, M, verify=True, len_hint=(150, 150))
        print(shape, it, "total %.2f h2d %.2f filter %.2f graph %.2f maxflow %.2f select %.2f verify %.2f d2h %.2f | rounds %d grs %d bfs %d partial %d cand %d path %d" % (
            r.ms_total, r.ms_h2d, r.ms_filter, r.ms_graph, r.ms_maxflow, r.ms_select, r.ms_verify, r.ms_d2h, r.rounds_total, r.global_relabels, r.bfs_levels, r.partial_bundles, r.partial_candidates, r.bundle_path), flush=True)
