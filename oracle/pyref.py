"""ctypes view of oracle/_ref/libgds_ref.so — the REFERENCE'S OWN sources (reads-gen, bam-api
containers, BamApi in-memory paths + read_bam over a fake in-memory BAM, CoverageTester) compiled
unmodified from /root/reference by oracle/Makefile.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box; the prebuilt .so travels with the snapshot."""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgds_ref.so")
_lib = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")

SOLVE_CB = C.CFUNCTYPE(C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32,
                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64))


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_gen_reads.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                    u32p, u32p, u32p, u32p]
        L.ref_input_cover.argtypes = [C.c_uint64, u32p, u32p, C.c_uint32, u32p]
        L.ref_filtered_cover.argtypes = [C.c_uint64, u32p, u32p, C.c_uint32, u64p, C.c_uint64, u32p]
        L.ref_find_pairs.argtypes = [C.c_uint64, u32p, u32p, C.c_uint32, u64p, C.c_uint64, u64p]
        L.ref_find_pairs.restype = C.c_uint64
        L.ref_read_bam.argtypes = [C.c_uint64, u32p, u32p, u32p, u32p, C.c_uint32, C.c_char_p,
                                   C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, u32p, u32p, u32p,
                                   u32p, u64p, u64p, C.POINTER(C.c_uint64)]
        L.ref_read_bam.restype = C.c_int64
        L.ref_run_coverage_tests.argtypes = [SOLVE_CB, C.c_void_p]
        L.ref_run_coverage_tests.restype = C.c_int
        _lib = L
    return _lib


def gen_reads(seed, pairs, L, R, shape=0):
    n = 2 * pairs
    s = np.empty(n, np.uint32); e = np.empty(n, np.uint32)
    q = np.empty(n, np.uint32); l = np.empty(n, np.uint32)
    lib().ref_gen_reads(seed, pairs, L, R, shape, s, e, q, l)
    return s, e, q, l


def input_cover(s, e, L):
    cov = np.zeros(L, np.uint32)
    lib().ref_input_cover(len(s), s, e, L, cov)
    return cov


def filtered_cover(s, e, L, ids):
    cov = np.zeros(L, np.uint32)
    ids = np.ascontiguousarray(ids, np.uint64)
    lib().ref_filtered_cover(len(s), s, e, L, ids, len(ids), cov)
    return cov


def find_pairs(s, e, L, ids):
    ids = np.ascontiguousarray(ids, np.uint64)
    out = np.zeros(len(s), np.uint64)
    k = lib().ref_find_pairs(len(s), s, e, L, ids, len(ids), out)
    return out[:k]


def read_bam(s, e, q, l, L, min_len, min_mapq, bed_text=None, tsv_text=None, grade=False):
    """Runs the reference's BamApi::read_bam on a fake in-memory BAM of these reads."""
    n = len(s)
    os_ = np.zeros(n, np.uint32); oe = np.zeros(n, np.uint32)
    oq = np.zeros(n, np.uint32); ol = np.zeros(n, np.uint32)
    oid = np.zeros(n, np.uint64); fo = np.zeros(n, np.uint64)
    nfo = C.c_uint64(0)
    with tempfile.TemporaryDirectory() as d:
        bed = tsv = b""
        mode = 0
        if bed_text:
            bp = os.path.join(d, "scheme.bed")
            open(bp, "w").write(bed_text)
            bed = bp.encode()
            mode = 2 if grade else 1
            if tsv_text:
                tp = os.path.join(d, "pairs.tsv")
                open(tp, "w").write(tsv_text)
                tsv = tp.encode()
        m = lib().ref_read_bam(n, s, e, q, l, L, bed, tsv if tsv else None, min_len, min_mapq,
                               mode, os_, oe, oq, ol, oid, fo, C.byref(nfo))
    return dict(start=os_[:m], end=oe[:m], quality=oq[:m], seq_len=ol[:m], bam_id=oid[:m],
                filtered_out=fo[:nfo.value])


def run_coverage_tests(solve_fn):
    """solve_fn(M, L, start, end) -> ascending kept indices.  Runs the reference's
    CoverageTester::test (5 cases, live asserts) against it; returns number of solve calls."""
    def cb(user, M, n, L, sp, ep, outp):
        s = np.ctypeslib.as_array(sp, shape=(n,)).copy()
        e = np.ctypeslib.as_array(ep, shape=(n,)).copy()
        ids = np.asarray(solve_fn(int(M), int(L), s, e), np.uint64)
        C.memmove(outp, ids.ctypes.data, ids.nbytes)
        return len(ids)
    fn = SOLVE_CB(cb)
    return lib().ref_run_coverage_tests(fn, None)
