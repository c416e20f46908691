"""ctypes view of libgds_host.so: the C++ host mirror (reads-gen, BamApi, SolverManager and the
qmcp::Solver plugin `quasi-mcp-b200`) for Python callers."""
import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
SHAPES = {"uniform": 0, "low_sides": 1, "hole": 2, "zero_sides": 3}


def load():
    global _LIB
    if _LIB is None:
        p = os.path.join(_HERE, "libgds_host.so")
        if not os.path.exists(p):
            raise RuntimeError("libgds_host.so is not built (run python __graft_entry__.py)")
        L = C.CDLL(p)
        L.gdsh_gen_reads.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gdsh_gen_reads.restype = C.c_int
        L.gdsh_gen_reads_amplicon.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_void_p, C.c_double, C.c_uint32,
                                              C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.gdsh_gen_reads_amplicon.restype = C.c_int
        L.gdsh_plugin_solve.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_void_p, C.c_uint64]
        L.gdsh_plugin_solve.restype = C.c_int64
        _LIB = L
    return _LIB


def gen_reads_into(seed, pairs, genome_len, read_len, start, end, mapq=None, seq_len=None,
                   shape="uniform"):
    """Fill caller-provided uint32 (uint8 for mapq) numpy views with one synthetic sample."""
    rc = load().gdsh_gen_reads(seed, pairs, genome_len, read_len, SHAPES[shape],
                               start.ctypes.data, end.ctypes.data,
                               mapq.ctypes.data if mapq is not None else None,
                               seq_len.ctypes.data if seq_len is not None else None)
    if rc != 0:
        raise ValueError("gdsh_gen_reads rc=%d" % rc)


def artic_amplicons(genome_len=30_000, n_amplicons=98, amp_len=400, overlap=98):
    """Synthetic ARTIC-style tiling: amplicon k = [k*(amp_len-overlap), +amp_len-1]."""
    k = np.arange(n_amplicons, dtype=np.uint32)
    a0 = k * np.uint32(amp_len - overlap)
    a1 = a0 + np.uint32(amp_len - 1)
    assert int(a1[-1]) < genome_len
    return a0, a1


def gen_reads_amplicon_into(seed, pairs, genome_len, amp_start, amp_end, start, end, mapq, seq_len,
                            p_inside=0.9, min_len=60, max_len=150):
    a0 = np.ascontiguousarray(amp_start, np.uint32)
    a1 = np.ascontiguousarray(amp_end, np.uint32)
    rc = load().gdsh_gen_reads_amplicon(seed, pairs, genome_len, len(a0), a0.ctypes.data,
                                        a1.ctypes.data, p_inside, min_len, max_len,
                                        start.ctypes.data, end.ctypes.data, mapq.ctypes.data,
                                        seq_len.ctypes.data)
    if rc != 0:
        raise ValueError("gdsh_gen_reads_amplicon rc=%d" % rc)


def gen_batch(seeds, pairs, genome_len, read_len, start, end, threads=8):
    """Sample k (seed seeds[k]) goes to start/end[2*pairs*k : 2*pairs*(k+1)].  The C call releases
    the GIL, so samples are generated on `threads` host threads."""
    n = 2 * pairs

    def one(k):
        gen_reads_into(int(seeds[k]), pairs, genome_len, read_len, start[k * n:(k + 1) * n],
                       end[k * n:(k + 1) * n])
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(len(seeds))))


def plugin_solve(algorithm, start, end, genome_len, max_coverage):
    """SolverManager.get(algorithm).solve(max_coverage, BamApi(reads)) -> ascending kept indices."""
    start = np.ascontiguousarray(start, np.uint32)
    end = np.ascontiguousarray(end, np.uint32)
    out = np.zeros(len(start), np.uint64)
    k = load().gdsh_plugin_solve(algorithm.encode(), len(start), genome_len, start.ctypes.data,
                                 end.ctypes.data, max_coverage, out.ctypes.data, len(out))
    if k < 0:
        raise KeyError("unknown algorithm %r" % algorithm)
    return out[:k]
