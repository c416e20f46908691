"""Parity campaign for the express schedule at realistic sizes: long references cut into segments
(default rule and explicit lengths), one or several read lengths, uniform and shaped coverage,
supplies from 50 to 3000 (frontiers below and above the CTA width, queue spills, heavy cut nodes).
Every case is solved with gds_params.schedule 0 and compared with the oracle's replay bit for bit.
Run from the repo root on a GPU box:  python tools/gpu_fuzz_express.py [N_CASES] [SEED]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_oracle, load_package  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    pkg, O = load_package(), load_oracle()
    solver = pkg.Solver(0)
    bad = n_express = n_comp = 0
    for it in range(n_cases):
        L = int(rng.choice([40_000, 90_000, 150_000, 400_000]))
        R = int(rng.choice([75, 100, 150, 250]))
        cov = int(rng.choice([200, 600, 1500, 3000]))
        M = int(rng.choice([50, 200, 500, 1000, 2500]))
        shape = str(rng.choice(["uniform", "uniform", "hole", "low_sides", "zero_sides"]))
        pairs = max(1, L * cov // R // 2)
        s, e, _, _ = O.gen_reads(int(rng.integers(1, 1 << 30)), pairs, L, R, shape)
        if rng.integers(0, 4) == 0:  # several read lengths
            e = np.minimum(e - rng.integers(0, 12, size=len(e)).astype(np.uint32), L - 1).astype(np.uint32)
            e = np.maximum(e, s)
        seg = int(rng.choice([0, 0, 4096, 8192, 20000]))
        sched = int(rng.choice([0, 0, 0, 2]))
        prm = (int(rng.choice([16, 64])), int(rng.choice([50, 150])), 1, 0, seg)
        ns = 1
        Ls = [L]
        off = np.array([0, len(s)], np.uint64)
        if rng.integers(0, 3) == 0:  # a whole 30 kb sample in the same call
            s2, e2, _, _ = O.gen_reads(int(rng.integers(1, 1 << 30)), 50_000, 30_000, 150)
            s = np.concatenate([s, s2]); e = np.concatenate([e, e2])
            Ls = [L, 30_000]
            off = np.array([0, off[1], len(s)], np.uint64)
        try:
            r = solver.solve(s, e, Ls, M, read_off=off, params=prm + (0, 0, sched), verify=True,
                             want_vectors=True)
            bm, st, dem, cov_v = O.sync_solve(s, e, Ls, off, M, params=prm + (sched,), want_vectors=True)
            ok = (r.fstar == st.fstar == r.flow_value == st.flow_value and
                  np.array_equal(r.demand, dem) and r.n_kept == st.n_kept and
                  np.array_equal(r.kept_bitmap, bm) and r.rounds_total == st.rounds_total and
                  r.pushes == st.pushes and r.relabels == st.relabels and
                  r.bfs_levels == st.bfs_levels and r.n_components == st.n_components and
                  r.verify_violations == 0)
            n_express += st.n_express
            n_comp += st.n_components
        except Exception as ex:  # noqa: BLE001
            ok = False
            print("case %d raised %r" % (it, ex))
        print("case %3d L=%6d R=%3d cov=%4d M=%4d %-10s seg=%5d sched=%d reads=%8d comps=%3d express=%3d "
              "rounds=%6d %s" % (it, L, R, cov, M, shape, seg, sched, len(s), st.n_components, st.n_express,
                                 st.rounds_total, "ok" if ok else "MISMATCH"), flush=True)
        if not ok:
            bad += 1
            if bad > 3:
                break
    print("express fuzz: %d cases, %d mismatches, %d of %d components on the express schedule"
          % (it + 1, bad, n_express, n_comp))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
