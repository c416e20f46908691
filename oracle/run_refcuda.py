"""Runs the REFERENCE'S OWN CUDA solver (quasi-mcp-cuda, oracle/_ref/libgds_refcuda.so built by
`make -C oracle refcuda`) on one input, in its own process: its error handling is std::terminate
and its driver loop has no iteration bound, so callers run this under a timeout.
TEST INFRASTRUCTURE (tests/ and bench.py's baseline legs only).

    python oracle/run_refcuda.py gen SEED PAIRS L R M [REPEATS]   -> JSON on stdout
    python oracle/run_refcuda.py npy IN.npz OUT.npy M             -> kept ids to OUT.npy
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libgds_refcuda.so")


def available():
    return os.path.exists(LIB)


def solve(lib, s, e, L, M):
    out = np.zeros(len(s), np.uint64)
    k = lib.ref_cuda_solve(len(s), s.ctypes.data, e.ctypes.data, L, M, out.ctypes.data, len(out))
    if k < 0:
        raise RuntimeError("ref_cuda_solve failed")
    return out[:k]


def main():
    lib = C.CDLL(LIB)
    lib.ref_cuda_solve.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                   C.c_void_p, C.c_uint64]
    lib.ref_cuda_solve.restype = C.c_int64
    if sys.argv[1] == "gen":
        sys.path.insert(0, HERE)
        import pyoracle as O
        seed, pairs, L, R, M = (int(x) for x in sys.argv[2:7])
        reps = int(sys.argv[7]) if len(sys.argv) > 7 else 1
        s, e, _, _ = O.gen_reads(seed, pairs, L, R)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            ids = solve(lib, s, e, L, M)
            times.append(time.perf_counter() - t0)
        mask = np.zeros(len(s), np.uint8)
        mask[ids] = 1
        cin = O.coverage_fast(s, e, L)
        cout = O.coverage_fast(s, e, L, mask)
        print(json.dumps({"reads": len(s), "L": L, "M": M, "n_kept": int(len(ids)),
                          "seconds": times, "best_s": min(times),
                          "invariant_ok": bool(np.all(np.minimum(cin, M) <= cout)),
                          "capped_equal": bool(np.array_equal(np.minimum(cin, M),
                                                              np.minimum(cout, M)))}))
    else:
        d = np.load(sys.argv[2])
        ids = solve(lib, np.ascontiguousarray(d["s"], np.uint32),
                    np.ascontiguousarray(d["e"], np.uint32), int(d["L"]), int(sys.argv[4]))
        np.save(sys.argv[3], ids)
    return 0


if __name__ == "__main__":
    sys.exit(main())
