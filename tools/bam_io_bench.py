"""Host-side throughput of the BAM front and back ends (no GPU needed unless --solve is given):
writes a synthetic coordinate-sorted BAM, then times BamApi's read (inflate + field extraction +
QNAME pairing) and the selective copy for 1..T threads.
    python tools/bam_io_bench.py [--pairs 1000000] [--genome 30000] [--threads 1,4,8] [--solve M]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from genome_downsampler_b200 import hostlib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--genome", type=int, default=30_000)
    ap.add_argument("--threads", default="1,4,8")
    ap.add_argument("--solve", type=int, default=0, help="also run quasi-mcp-b200 with this MAX_COVERAGE (GPU)")
    ap.add_argument("--dir", default=None)
    a = ap.parse_args()
    n = 2 * a.pairs
    s = np.empty(n, np.uint32); e = np.empty(n, np.uint32); q = np.empty(n, np.uint8); l = np.empty(n, np.uint32)
    hostlib.gen_reads_into(12345, a.pairs, a.genome, 150, s, e, q, l)
    with tempfile.TemporaryDirectory(dir=a.dir) as d:
        path = os.path.join(d, "in.bam")
        t0 = time.perf_counter()
        hostlib.write_synthetic_bam(path, a.genome, s, e, q, l, coordinate_sorted=True, threads=os.cpu_count() or 8)
        t_gen = time.perf_counter() - t0
        size = os.path.getsize(path)
        out = {"reads": n, "bam_bytes": size, "host_cores": os.cpu_count(), "write_synthetic_s": round(t_gen, 3), "runs": []}
        for t in [int(x) for x in a.threads.split(",")]:
            b = hostlib.BamFile(path, threads=t)
            t0 = time.perf_counter()
            cnt = b.record_count  # triggers read_bam
            wall = time.perf_counter() - t0
            run = {"threads": t, "read_bam_s": round(b.read_seconds, 4), "reads_per_s": round(cnt / b.read_seconds),
                   "compressed_MBps": round(size / b.read_seconds / 1e6, 1), "wall_s": round(wall, 4)}
            if a.solve:
                t0 = time.perf_counter()
                kept = b.solve("quasi-mcp-b200", a.solve)
                run["solve_s"] = round(time.perf_counter() - t0, 4)
                run["kept"] = int(len(kept))
            else:
                b.reads()
                kept = np.arange(0, cnt, 50, dtype=np.uint64)  # 2 % of the reads, like config 1
            t0 = time.perf_counter()
            w = b.write_solution(os.path.join(d, "out.bam"), kept, with_pairs=True)
            run["write_s"] = round(time.perf_counter() - t0, 4)
            run["written"] = int(w)
            out["runs"].append(run)
            b.close()
        print(json.dumps(out))


if __name__ == "__main__":
    main()
