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
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.gdsh_plugin_solve_batch.argtypes = [u32, vp, u32, vp, vp, u32, C.c_int, vp, vp, vp, u64]
        L.gdsh_plugin_solve_batch.restype = C.c_int64
        L.gdsh_encode_compact.argtypes = [vp, vp, u64, vp, vp, vp, u32]
        L.gdsh_encode_compact.restype = C.c_int
        L.gdsh_write_synthetic_bam.argtypes = [C.c_char_p, u64, u32, vp, vp, vp, vp, C.c_int, u32, u32]
        L.gdsh_write_synthetic_bam.restype = C.c_int64
        L.gdsh_bam_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, u32, u32, C.c_int, u32]
        L.gdsh_bam_open.restype = vp
        L.gdsh_bam_close.argtypes = [vp]
        L.gdsh_bam_close.restype = None
        for f in (L.gdsh_bam_unfiltered, L.gdsh_bam_reads):
            f.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp]
            f.restype = u64
        L.gdsh_bam_filtered_out.argtypes = [vp, u64, vp]
        L.gdsh_bam_filtered_out.restype = u64
        for f in (L.gdsh_bam_ref_length, L.gdsh_bam_record_count):
            f.argtypes = [vp]
            f.restype = u64
        L.gdsh_bam_read_seconds.argtypes = [vp]
        L.gdsh_bam_read_seconds.restype = C.c_double
        L.gdsh_bam_solve.argtypes = [vp, C.c_char_p, u32, vp, u64]
        L.gdsh_bam_solve.restype = C.c_int64
        L.gdsh_bam_write_solution.argtypes = [vp, C.c_char_p, vp, u64, C.c_int]
        L.gdsh_bam_write_solution.restype = C.c_int64
        L.gdsh_bam_write_filtered_out.argtypes = [vp, C.c_char_p]
        L.gdsh_bam_write_filtered_out.restype = C.c_int64
        L.gdsh_write_bam.argtypes = [C.c_char_p, C.c_char_p, vp, u64, u32]
        L.gdsh_write_bam.restype = C.c_int64
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


def plugin_solve_batch(start, end, read_off, genome_len, max_coverage, repeats=1, want_indices=True):
    """QuasiMcpB200MaxFlowSolver::solve_batch on BamApi objects built from the arrays.  Returns
    (list of per-sample ascending index arrays or None, per-sample kept counts, best seconds of the
    solve_batch call itself: narrowing + device + index expansion)."""
    start = np.ascontiguousarray(start, np.uint32)
    end = np.ascontiguousarray(end, np.uint32)
    read_off = np.ascontiguousarray(read_off, np.uint64)
    ns = len(read_off) - 1
    counts = np.zeros(ns, np.uint64)
    secs = C.c_double(0)
    cap = len(start) if want_indices else 0
    out = np.zeros(max(cap, 1), np.uint64)
    tot = load().gdsh_plugin_solve_batch(ns, read_off.ctypes.data, genome_len, start.ctypes.data,
                                         end.ctypes.data, max_coverage, repeats, C.addressof(secs),
                                         counts.ctypes.data, out.ctypes.data if want_indices else None,
                                         cap)
    per = None
    if want_indices:
        cuts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        per = [out[cuts[k]:cuts[k + 1]] for k in range(ns)]
    assert tot == int(counts.sum())
    return per, counts, float(secs.value)


def encode_compact(start_ptr, end_ptr, n, start16_ptr, threads=16):
    """u32 start/end columns -> 16-bit starts + exact (len_min, len_max); raw pointers so it can
    work on slices of pinned buffers.  Returns (fits16, len_min, len_max).  Releases the GIL."""
    lo, hi = C.c_uint32(0), C.c_uint32(0)
    ok = load().gdsh_encode_compact(start_ptr, end_ptr, n, start16_ptr, C.addressof(lo),
                                    C.addressof(hi), threads)
    return bool(ok), int(lo.value), int(hi.value)


# ---- BAM files: the file-backed BamApi (bam_api.cpp:30-42, :359-656) ----

AMPLICON_BEHAVIOUR = {"ignore": 0, "filter": 1, "grade": 2}


def write_synthetic_bam(path, genome_len, start, end, mapq=None, seq_len=None, coordinate_sorted=False,
                        threads=4, seed=1):
    start = np.ascontiguousarray(start, np.uint32)
    end = np.ascontiguousarray(end, np.uint32)
    mapq = None if mapq is None else np.ascontiguousarray(mapq, np.uint8)
    seq_len = None if seq_len is None else np.ascontiguousarray(seq_len, np.uint32)
    rc = load().gdsh_write_synthetic_bam(str(path).encode(), len(start), genome_len, start.ctypes.data,
                                         end.ctypes.data, mapq.ctypes.data if mapq is not None else None,
                                         seq_len.ctypes.data if seq_len is not None else None,
                                         int(coordinate_sorted), threads, seed)
    if rc < 0:
        raise IOError("cannot write %s" % path)
    return rc


class BamFile:
    """BamApi(input_filepath, config).  Errors in the file end the process (the reference's
    log-and-exit), so tests of broken inputs run in a subprocess."""

    def __init__(self, path, min_len=0, min_mapq=0, bed=None, tsv=None, amplicon_behaviour="filter",
                 threads=1):
        self._L = load()
        self._h = self._L.gdsh_bam_open(str(path).encode(), str(bed).encode() if bed else None,
                                        str(tsv).encode() if tsv else None, min_len, min_mapq,
                                        AMPLICON_BEHAVIOUR[amplicon_behaviour], threads)

    def close(self):
        if self._h:
            self._L.gdsh_bam_close(self._h)
            self._h = None

    __del__ = close

    def _columns(self, fn):
        n = fn(self._h, 0, None, None, None, None, None, None)
        cols = {"bam_id": np.empty(n, np.uint64), "start": np.empty(n, np.uint64),
                "end": np.empty(n, np.uint64), "quality": np.empty(n, np.uint32),
                "seq_length": np.empty(n, np.uint32), "is_first": np.empty(n, np.uint8)}
        fn(self._h, n, *[c.ctypes.data for c in cols.values()])
        return cols

    def unfiltered(self):
        """Pair-ordered reads before the filter ({} once the filter has been applied)."""
        return self._columns(self._L.gdsh_bam_unfiltered)

    def reads(self):
        """get_paired_reads_soa(): post-filter arrays."""
        return self._columns(self._L.gdsh_bam_reads)

    def filtered_out(self):
        n = self._L.gdsh_bam_filtered_out(self._h, 0, None)
        out = np.empty(n, np.uint64)
        self._L.gdsh_bam_filtered_out(self._h, n, out.ctypes.data)
        return out

    @property
    def ref_length(self):
        return self._L.gdsh_bam_ref_length(self._h)

    @property
    def record_count(self):
        return self._L.gdsh_bam_record_count(self._h)

    @property
    def read_seconds(self):
        return self._L.gdsh_bam_read_seconds(self._h)

    def solve(self, algorithm, max_coverage):
        cap = max(int(self.record_count), 1)
        out = np.empty(cap, np.uint64)
        k = self._L.gdsh_bam_solve(self._h, algorithm.encode(), max_coverage, out.ctypes.data, cap)
        if k < 0:
            raise KeyError("unknown algorithm %r" % algorithm)
        return out[:k]

    def write_solution(self, out_path, kept, with_pairs=True):
        kept = np.ascontiguousarray(kept, np.uint64)
        return self._L.gdsh_bam_write_solution(self._h, str(out_path).encode(), kept.ctypes.data,
                                               len(kept), int(with_pairs))

    def write_filtered_out(self, out_path):
        return self._L.gdsh_bam_write_filtered_out(self._h, str(out_path).encode())


def write_bam(in_path, out_path, bam_ids, threads=1):
    ids = np.ascontiguousarray(bam_ids, np.uint64)
    return load().gdsh_write_bam(str(in_path).encode(), str(out_path).encode(), ids.ctypes.data,
                                 len(ids), threads)
