"""Ad-hoc GPU probe used during bring-up: parity against the oracle on a ladder of cases, with
timings.  Not collected by pytest (no test_ prefix); the real parity tests are test_gpu_*.py."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from __graft_entry__ import load_oracle, load_package  # noqa: E402

PRM = (64, 150, 1, 0)


def check(solver, O, name, s, e, ref_lens, read_off, M, log):
    for _ in range(2 if len(s) > 50000 else 1):  # second call = warm arenas
        t0 = time.time()
        r = solver.solve(s, e, ref_lens, M, read_off=read_off, params=PRM, verify=True,
                         want_vectors=True)
        t1 = time.time()
    bm, st, dem, cov = O.sync_solve(s, e, ref_lens, read_off, M, params=PRM, want_vectors=True)
    t2 = time.time()
    ok = dict(
        fstar=r.fstar == st.fstar, flow=r.flow_value == st.flow_value,
        demand=bool(np.array_equal(r.demand, dem)),
        cov=bool(np.array_equal(r.cov_capped, np.minimum(cov, M))),
        bitmap=bool(np.array_equal(r.kept_bitmap, bm)), verify=r.verify_violations == 0,
        rounds=r.rounds_total == st.rounds_total, bundles=r.n_bundles == st.n_bundles,
        comps=r.n_components == st.n_components)
    rec = dict(name=name, ok=all(ok.values()), detail=ok, n=len(s), fstar=int(r.fstar),
               kept=int(r.n_kept), rounds=int(r.rounds_total), bfs=int(r.bfs_levels),
               oracle_rounds=int(st.rounds_total), passes=int(r.sort_passes),
               ms=dict(h2d=r.ms_h2d, filt=r.ms_filter, graph=r.ms_graph, mf=r.ms_maxflow,
                       sel=r.ms_select, ver=r.ms_verify, d2h=r.ms_d2h, total=r.ms_total),
               wall_gpu_s=t1 - t0, wall_oracle_s=t2 - t1)
    print(json.dumps(rec), flush=True)
    log.append(rec)
    return rec["ok"]


def main():
    pkg = load_package()
    O = load_oracle()
    solver = pkg.Solver(0)
    log = []
    ex = O.SMALL_EXAMPLE
    small_only = "--big-only" in sys.argv
    allok = check(solver, O, "small16", ex["start"], ex["end"], [ex["L"]], [0, 16], ex["M"], log)
    rng = np.random.default_rng(7)
    for it in range(0 if small_only else 40):
        ns = int(rng.integers(1, 5))
        Ls = rng.integers(1, 400, size=ns)
        ss, ee, off = [], [], [0]
        for L in Ls:
            n = 2 * int(rng.integers(0, 300))
            mode = int(rng.integers(0, 3))
            if mode == 0:
                s = rng.integers(0, L, size=n); ln = rng.integers(1, max(2, L // 2 + 1), size=n)
            elif mode == 1:
                s = rng.integers(0, max(1, L // 3), size=n); ln = rng.integers(1, 6, size=n)
            else:
                s = rng.integers(0, L, size=n); ln = np.full(n, rng.integers(1, 30))
            e = np.minimum(s + ln - 1, L - 1)
            ss.append(s); ee.append(e); off.append(off[-1] + n)
        s = np.concatenate(ss).astype(np.uint32); e = np.concatenate(ee).astype(np.uint32)
        allok &= check(solver, O, "fuzz%d" % it, s, e, Ls.astype(np.uint32),
                       np.array(off, np.uint64), int(rng.integers(1, 12)), log)
    cases = [("c3", 50_000, 30_000, 100, "uniform"), ("c1", 500_000, 30_000, 100, "uniform"),
             ("t_uni", 1_000_000, 30_000, 1000, "uniform"),
             ("t_low", 1_000_000, 30_000, 8000, "low_sides"),
             ("t_hole", 1_000_000, 30_000, 8000, "hole"),
             ("t_zero", 1_000_000, 30_000, 8000, "zero_sides")]
    for name, pairs, L, M, shape in cases:
        s, e, q, l = O.gen_reads(12345, pairs, L, 150, shape)
        allok &= check(solver, O, name, s, e, [L], [0, len(s)], M, log)
    # batch of 8 samples
    ss, ee, off = [], [], [0]
    for k in range(8):
        s, e, q, l = O.gen_reads(12345 + k, 250_000, 30_000, 150)
        ss.append(s); ee.append(e); off.append(off[-1] + len(s))
    allok &= check(solver, O, "batch8", np.concatenate(ss), np.concatenate(ee), [30_000] * 8,
                   np.array(off, np.uint64), 100, log)
    if "--c4" in sys.argv:
        s, e, q, l = O.gen_reads(12345, 25_000_000, 5_000_000, 150)
        allok &= check(solver, O, "c4", s, e, [5_000_000], [0, len(s)], 500, log)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(log, open("gpurun_out/probe.json", "w"), indent=1)
    print("ALL OK" if allok else "FAILURES")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
