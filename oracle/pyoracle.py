"""ctypes view of the CPU oracle (oracle/libgds_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/gds_oracle.h).  The product path never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgds_oracle.so")
_lib = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class RefStats(C.Structure):
    _fields_ = [("flow_value", C.c_int64), ("n_kept", C.c_uint64), ("n_arcs", C.c_uint64),
                ("pushes", C.c_uint64), ("relabels", C.c_uint64), ("global_relabels", C.c_uint64),
                ("t_coverage_s", C.c_double), ("t_graph_s", C.c_double),
                ("t_maxflow_s", C.c_double), ("t_select_s", C.c_double), ("t_total_s", C.c_double)]


class SyncParams(C.Structure):
    _fields_ = [("gr_interval_min", C.c_uint32), ("gr_levels_pct", C.c_uint32),
                ("gr_relabel_pct", C.c_uint32),
                ("max_rounds", C.c_uint32), ("seg_len", C.c_uint32), ("schedule", C.c_uint32)]


class SyncStats(C.Structure):
    _fields_ = [("flow_value", C.c_int64), ("fstar", C.c_int64), ("n_kept", C.c_uint64),
                ("n_bundles", C.c_uint64), ("n_components", C.c_uint32),
                ("rounds_total", C.c_uint64), ("rounds_max", C.c_uint64), ("pushes", C.c_uint64),
                ("relabels", C.c_uint64), ("global_relabels", C.c_uint64),
                ("bfs_levels", C.c_uint64), ("max_frontier", C.c_uint64), ("n_express", C.c_uint32),
                ("t_build_s", C.c_double), ("t_solve_s", C.c_double), ("t_select_s", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "gds_oracle.cpp"))):
        subprocess.check_call(["make", "-C", _HERE, "libgds_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.orc_gen_reads.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                C.c_int32, u32p, u32p, u32p, u32p]
    L.orc_gen_reads.restype = C.c_int
    L.orc_gen_artic_scheme.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64]
    L.orc_gen_artic_scheme.restype = C.c_int
    L.orc_gen_reads_amplicon.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, u32p,
                                         u32p, C.c_double, C.c_uint32, C.c_uint32, C.c_int32,
                                         u32p, u32p, u32p, u32p]
    L.orc_gen_reads_amplicon.restype = C.c_int
    L.orc_parse_amplicons.argtypes = [C.c_char_p, C.c_char_p, u32p, u32p, C.c_uint64]
    L.orc_parse_amplicons.restype = C.c_int64
    L.orc_filter_pairs.argtypes = [C.c_uint64, u32p, u32p, u32p, u32p, C.c_uint32, C.c_uint32,
                                   C.c_int, C.c_uint32, u32p, u32p, u8p]
    L.orc_filter_pairs.restype = C.c_uint64
    L.orc_coverage_ref.argtypes = [C.c_uint64, u32p, u32p, C.c_uint32, u32p]
    L.orc_coverage_ref.restype = C.c_int
    L.orc_coverage_subset.argtypes = [C.c_uint64, u32p, u32p, u8p, C.c_uint32, u32p]
    L.orc_coverage_subset.restype = C.c_int
    L.orc_demand.argtypes = [u32p, C.c_uint32, C.c_uint32, i32p]
    L.orc_demand.restype = C.c_int64
    L.orc_ref_solve.argtypes = [C.c_uint64, u32p, u32p, C.c_uint32, C.c_uint32, u8p,
                                C.POINTER(RefStats)]
    L.orc_ref_solve.restype = C.c_int
    L.orc_sync_solve.argtypes = [C.c_uint32, u64p, u32p, u32p, u32p, C.c_uint32,
                                 C.POINTER(SyncParams), u32p, C.c_void_p, C.c_void_p,
                                 C.POINTER(SyncStats)]
    L.orc_sync_solve.restype = C.c_int
    L.orc_sweep_solve.argtypes = L.orc_sync_solve.argtypes
    L.orc_sweep_solve.restype = C.c_int
    L.orc_greedy_multicover.argtypes = [C.c_uint64, u32p, u32p, C.c_uint32, C.c_uint32, u8p]
    L.orc_greedy_multicover.restype = C.c_uint64
    L.orc_find_pairs_bitmap.argtypes = [C.c_uint64, u32p]
    L.orc_find_pairs_bitmap.restype = None
    _lib = L
    return L


SHAPES = {"uniform": 0, "low_sides": 1, "hole": 2, "zero_sides": 3}


def gen_reads(seed, pairs, genome_len, read_len, shape="uniform", max_quality=100):
    n = 2 * pairs
    s = np.empty(n, np.uint32); e = np.empty(n, np.uint32)
    q = np.empty(n, np.uint32); l = np.empty(n, np.uint32)
    rc = lib().orc_gen_reads(seed, pairs, genome_len, read_len, SHAPES[shape], max_quality,
                             s, e, q, l)
    if rc != 0:
        raise ValueError("orc_gen_reads rc=%d" % rc)
    return s, e, q, l


def artic_scheme(genome_len=30000, n_amplicons=98, amp_len=400, overlap=98, primer_len=22):
    bed = C.create_string_buffer(1 << 16); tsv = C.create_string_buffer(1 << 16)
    rc = lib().orc_gen_artic_scheme(genome_len, n_amplicons, amp_len, overlap, primer_len,
                                    bed, len(bed), tsv, len(tsv))
    if rc != 0:
        raise ValueError("orc_gen_artic_scheme rc=%d" % rc)
    return bed.value.decode(), tsv.value.decode()


def parse_amplicons(bed_text, tsv_text=None, cap=65536):
    a0 = np.zeros(cap, np.uint32); a1 = np.zeros(cap, np.uint32)
    n = lib().orc_parse_amplicons(bed_text.encode(), tsv_text.encode() if tsv_text else None,
                                  a0, a1, cap)
    if n < 0:
        raise ValueError("orc_parse_amplicons rc=%d" % n)
    return a0[:n].copy(), a1[:n].copy()


def gen_reads_amplicon(seed, pairs, genome_len, amp_start, amp_end, p_inside=0.9, min_len=60,
                       max_len=150, max_quality=100):
    n = 2 * pairs
    s = np.empty(n, np.uint32); e = np.empty(n, np.uint32)
    q = np.empty(n, np.uint32); l = np.empty(n, np.uint32)
    rc = lib().orc_gen_reads_amplicon(seed, pairs, genome_len, len(amp_start),
                                      np.ascontiguousarray(amp_start, np.uint32),
                                      np.ascontiguousarray(amp_end, np.uint32), p_inside, min_len,
                                      max_len, max_quality, s, e, q, l)
    if rc != 0:
        raise ValueError("orc_gen_reads_amplicon rc=%d" % rc)
    return s, e, q, l


def filter_pairs(start, end, quality, seq_len, min_len, min_mapq, amp_start=None, amp_end=None):
    n = len(start)
    pp = np.zeros(n // 2, np.uint8)
    use = amp_start is not None
    a0 = np.ascontiguousarray(amp_start if use else [0], np.uint32)
    a1 = np.ascontiguousarray(amp_end if use else [0], np.uint32)
    kept = lib().orc_filter_pairs(n, start, end, quality, seq_len, min_len, min_mapq, int(use),
                                  len(a0) if use else 0, a0, a1, pp)
    return pp, int(kept)


def coverage(start, end, L, kept=None):
    cov = np.zeros(L, np.uint32)
    if kept is None:
        rc = lib().orc_coverage_ref(len(start), start, end, L, cov)
    else:
        rc = lib().orc_coverage_subset(len(start), start, end,
                                       np.ascontiguousarray(kept, np.uint8), L, cov)
    if rc != 0:
        raise ValueError("bad read coordinates")
    return cov


def coverage_fast(start, end, L, kept=None):
    """numpy difference-array coverage (for sizes where the per-base loop is too slow)."""
    if kept is not None:
        m = np.asarray(kept, bool)
        start, end = start[m], end[m]
    d = np.zeros(L + 1, np.int64)
    np.add.at(d, start, 1)
    # end + 1 in 32-bit arithmetic: a read of length 0 (end == start - 1) cancels itself, also at 0
    np.add.at(d, (np.asarray(end, np.uint32) + np.uint32(1)).astype(np.int64), -1)
    return np.cumsum(d[:L]).astype(np.uint32)


def demand(cov, M):
    L = len(cov)
    d = np.zeros(L + 1, np.int32)
    f = lib().orc_demand(np.ascontiguousarray(cov, np.uint32), L, M, d)
    return d, int(f)


def ref_solve(start, end, L, M):
    kept = np.zeros(len(start), np.uint8)
    st = RefStats()
    rc = lib().orc_ref_solve(len(start), start, end, L, M, kept, C.byref(st))
    if rc != 0:
        raise ValueError("orc_ref_solve rc=%d" % rc)
    return kept, st


def sweep_solve(start, end, ref_lens, read_off, M, params=None, want_vectors=False):
    """Minimum-cardinality solve (mcp-cpu's objective) as the device's deterministic sweep."""
    return sync_solve(start, end, ref_lens, read_off, M, params, want_vectors, _fn="orc_sweep_solve")


def sync_solve(start, end, ref_lens, read_off, M, params=None, want_vectors=False, _fn="orc_sync_solve"):
    """Batch-aware deterministic schedule.  Returns (bitmap words, stats[, demand, covR])."""
    ref_lens = np.ascontiguousarray(ref_lens, np.uint32)
    read_off = np.ascontiguousarray(read_off, np.uint64)
    n = int(read_off[-1])
    bm = np.zeros((n + 31) // 32, np.uint32)
    st = SyncStats()
    prm = SyncParams(*params) if params is not None else None
    nn = int(ref_lens.astype(np.int64).sum() + len(ref_lens))
    dem = np.zeros(nn, np.int32) if want_vectors else None
    cov = np.zeros(nn, np.uint32) if want_vectors else None
    rc = getattr(lib(), _fn)(len(ref_lens), read_off, ref_lens, start, end, M,
                              C.byref(prm) if prm is not None else None, bm,
                              dem.ctypes.data if want_vectors else None,
                              cov.ctypes.data if want_vectors else None, C.byref(st))
    if rc != 0:
        raise ValueError("%s rc=%d" % (_fn, rc))
    if want_vectors:
        return bm, st, dem, cov
    return bm, st


def greedy_multicover(start, end, L, M):
    kept = np.zeros(len(start), np.uint8)
    n = lib().orc_greedy_multicover(len(start), start, end, L, M, kept)
    return kept, int(n)


def find_pairs_bitmap(bitmap, n):
    bm = np.ascontiguousarray(bitmap, np.uint32).copy()
    lib().orc_find_pairs_bitmap(n, bm)
    return bm


def bitmap_to_mask(bm, n):
    return np.unpackbits(bm.view(np.uint8), bitorder="little")[:n].astype(np.uint8)


SMALL_EXAMPLE = dict(  # src/tests/coverage_tester.cpp:72-93
    start=np.array([0, 6, 2, 6, 1, 7, 3, 9, 0, 7, 4, 9, 1, 6, 0, 4], np.uint32),
    end=np.array([2, 9, 4, 8, 3, 10, 6, 10, 4, 9, 6, 10, 4, 8, 2, 6], np.uint32),
    L=11, M=4)
