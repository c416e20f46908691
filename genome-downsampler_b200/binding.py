"""ctypes binding of include/gds.h.  No torch types, no oracle imports: this is product code and
it fails loudly when libgds_b200.so or a CUDA device is missing (there is no CPU fallback)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GDS_FLAGS = dict(INPUT_ON_DEVICE=1, OUTPUT_ON_DEVICE=2, VERIFY=4, FIND_PAIRS=8, NO_SOLVE=16,
                 PROFILE_KERNELS=32)
_ERR = {1: "GDS_ERR_ARG", 2: "GDS_ERR_RANGE", 3: "GDS_ERR_CUDA", 4: "GDS_ERR_NOMEM",
        5: "GDS_ERR_NOCONVERGE"}

ENTRY_POINTS = ["gds_abi_version", "gds_create", "gds_destroy", "gds_last_error", "gds_set_stream",
                "gds_solve", "gds_kernel_profile", "gds_kernel_profile_reset", "gds_bitmap_to_indices",
                "gds_host_alloc", "gds_host_free"]


class GdsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (_ERR.get(code, code), msg))
        self.code = code


class _Reads(C.Structure):
    _fields_ = [("n_samples", C.c_uint32), ("read_off", C.c_void_p), ("ref_len", C.c_void_p),
                ("start", C.c_void_p), ("end", C.c_void_p), ("mapq", C.c_void_p),
                ("seq_len", C.c_void_p), ("len_min", C.c_uint32), ("len_max", C.c_uint32),
                ("start16", C.c_void_p)]


class _Filter(C.Structure):
    _fields_ = [("min_seq_length", C.c_uint32), ("min_mapq", C.c_uint32),
                ("n_amplicons", C.c_uint32), ("amp_start", C.c_void_p), ("amp_end", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("gr_interval_min", C.c_uint32), ("gr_levels_pct", C.c_uint32),
                ("gr_relabel_pct", C.c_uint32), ("max_rounds", C.c_uint32),
                ("seg_len", C.c_uint32), ("bundle_mode", C.c_uint32), ("algorithm", C.c_uint32),
                ("schedule", C.c_uint32)]


class _KStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_float), ("launches", C.c_uint32),
                ("bytes", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("kept_bitmap", C.c_void_p), ("pair_pass", C.c_void_p), ("filt_off", C.c_void_p),
                ("cov_capped", C.c_void_p), ("demand", C.c_void_p),
                ("n_reads_in", C.c_uint64), ("n_filtered", C.c_uint64), ("n_kept", C.c_uint64),
                ("n_bundles", C.c_uint64), ("n_arc_items", C.c_uint64),
                ("n_nodes", C.c_uint32), ("n_components", C.c_uint32),
                ("fstar", C.c_int64), ("flow_value", C.c_int64),
                ("rounds_total", C.c_uint64), ("rounds_max", C.c_uint64), ("pushes", C.c_uint64),
                ("relabels", C.c_uint64), ("global_relabels", C.c_uint64),
                ("bfs_levels", C.c_uint64), ("max_frontier", C.c_uint64),
                ("verify_violations", C.c_uint64), ("key_bits", C.c_uint32),
                ("sort_passes", C.c_uint32), ("kernel_launches", C.c_uint64),
                ("partial_bundles", C.c_uint64), ("partial_candidates", C.c_uint64),
                ("bundle_path", C.c_uint32), ("seg_len", C.c_uint32),
                ("ms_h2d", C.c_float), ("ms_filter", C.c_float), ("ms_graph", C.c_float),
                ("ms_maxflow", C.c_float), ("ms_select", C.c_float), ("ms_verify", C.c_float),
                ("ms_d2h", C.c_float), ("ms_total", C.c_float)]


_SCALARS = [n for n, _ in _Result._fields_[5:]]


def lib_path():
    # GDS_LIB_PATH: an instrumented build of the same library (tools/, diagnostics only)
    return os.environ.get("GDS_LIB_PATH") or os.path.join(_HERE, "libgds_b200.so")


def exported_symbols():
    return list(ENTRY_POINTS)


def load_library():
    """dlopen libgds_b200.so.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise GdsError(3, "libgds_b200.so is not built (run python __graft_entry__.py); "
                          "there is no CPU fallback")
    L = C.CDLL(p)
    L.gds_abi_version.restype = C.c_int
    L.gds_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.gds_create.restype = C.c_int
    L.gds_destroy.argtypes = [C.c_void_p]
    L.gds_destroy.restype = None
    L.gds_last_error.argtypes = [C.c_void_p]
    L.gds_last_error.restype = C.c_char_p
    L.gds_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.gds_set_stream.restype = C.c_int
    L.gds_solve.argtypes = [C.c_void_p, C.POINTER(_Reads), C.POINTER(_Filter), C.c_uint32,
                            C.POINTER(_Params), C.c_uint32, C.POINTER(_Result)]
    L.gds_solve.restype = C.c_int
    L.gds_kernel_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.gds_kernel_profile.restype = C.c_uint32
    L.gds_kernel_profile_reset.argtypes = [C.c_void_p]
    L.gds_kernel_profile_reset.restype = None
    L.gds_bitmap_to_indices.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    L.gds_bitmap_to_indices.restype = C.c_uint64
    _LIB = L
    return L


class Result(dict):
    """Scalars of gds_result plus whichever output arrays were requested."""
    __getattr__ = dict.__getitem__


def _ptr(a):
    return None if a is None else a.ctypes.data


class Solver:
    """One device context (gds_ctx).  Reusable across calls like the reference's solver
    instances (src/tests/coverage_tester.cpp:28-43)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.gds_create(device, C.byref(h))
        if rc != 0:
            raise GdsError(rc, "gds_create(device=%d) failed: no usable CUDA device; "
                               "this library has no CPU fallback" % device)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gds_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._lib.gds_set_stream(self._h, C.c_void_p(cuda_stream_ptr))

    def kernel_profile_reset(self):
        self._lib.gds_kernel_profile_reset(self._h)

    def kernel_profile(self):
        """[{name, ms, launches, bytes}] over the profile=True calls since the last reset."""
        buf = (_KStat * 64)()
        n = self._lib.gds_kernel_profile(self._h, buf, 64)
        return [dict(name=buf[i].name.decode(), ms=float(buf[i].ms), launches=int(buf[i].launches),
                     bytes=int(buf[i].bytes)) for i in range(min(n, 64))]

    def _check(self, rc):
        if rc != 0:
            raise GdsError(rc, self._lib.gds_last_error(self._h).decode())

    # ------------------------------------------------------------------ host-buffer path
    def solve(self, start, end, ref_len, max_coverage, read_off=None, mapq=None, seq_len=None,
              filt=None, params=None, verify=False, find_pairs=False, no_solve=False,
              want_vectors=False, profile=False, len_hint=None):
        """Host numpy arrays in, host numpy arrays out (copies happen inside the C call).

        filt = dict(min_len=, min_mapq=, amp_start=None, amp_end=None) or None.
        """
        # compact transport (gds_reads.start16 / end == NULL): a uint16 start array is passed as
        # start16, end=None means fixed-length reads of len_hint[0] == len_hint[1]
        start16 = None
        if isinstance(start, np.ndarray) and start.dtype == np.uint16:
            start16 = np.ascontiguousarray(start)
            start = None
        else:
            start = np.ascontiguousarray(start, np.uint32)
        if end is not None:
            end = np.ascontiguousarray(end, np.uint32)
        n = len(start16 if start16 is not None else start)
        ref_len = np.atleast_1d(np.ascontiguousarray(ref_len, np.uint32))
        ns = len(ref_len)
        if read_off is None:
            read_off = np.array([0, n], np.uint64)
        read_off = np.ascontiguousarray(read_off, np.uint64)
        lh = len_hint or (0, 0)  # exact (min, max) of end-start+1, see gds_reads.len_min
        rd = _Reads(ns, _ptr(read_off), _ptr(ref_len), _ptr(start), _ptr(end), None, None,
                    int(lh[0]), int(lh[1]), _ptr(start16))
        keep = [start, start16, end, ref_len, read_off]
        fl = None
        if filt is not None:
            mapq = np.ascontiguousarray(mapq, np.uint8)
            seq_len = np.ascontiguousarray(seq_len, np.uint32)
            rd.mapq, rd.seq_len = _ptr(mapq), _ptr(seq_len)
            a0 = filt.get("amp_start")
            a1 = filt.get("amp_end")
            na = 0 if a0 is None else len(a0)
            a0 = np.ascontiguousarray(a0 if na else [0], np.uint32)
            a1 = np.ascontiguousarray(a1 if na else [0], np.uint32)
            fl = _Filter(filt.get("min_len", 0), filt.get("min_mapq", 0), na, _ptr(a0), _ptr(a1))
            keep += [mapq, seq_len, a0, a1]
        nn = int(ref_len.astype(np.int64).sum() + ns)
        res = _Result()
        bitmap = np.zeros((n + 31) // 32 + 1, np.uint32)
        pair_pass = np.zeros(max(n // 2, 1), np.uint8) if filt is not None else None
        filt_off = np.zeros(ns + 1, np.uint64)
        cov = np.zeros(nn, np.uint32) if want_vectors else None
        dem = np.zeros(nn, np.int32) if want_vectors else None
        res.kept_bitmap, res.pair_pass = _ptr(bitmap), _ptr(pair_pass)
        res.filt_off, res.cov_capped, res.demand = _ptr(filt_off), _ptr(cov), _ptr(dem)
        flags = (4 if verify else 0) | (8 if find_pairs else 0) | (16 if no_solve else 0) | \
                (32 if profile else 0)
        prm = _Params(*params) if params is not None else None
        rc = self._lib.gds_solve(self._h, C.byref(rd), C.byref(fl) if fl is not None else None,
                                 int(max_coverage), C.byref(prm) if prm is not None else None,
                                 flags, C.byref(res))
        self._check(rc)
        out = Result({k: getattr(res, k) for k in _SCALARS})
        nf = int(res.n_filtered)
        out["kept_bitmap"] = bitmap[:(nf + 31) // 32]
        out["pair_pass"] = pair_pass[:n // 2] if pair_pass is not None else None
        out["filt_off"] = filt_off
        out["cov_capped"] = cov
        out["demand"] = dem
        del keep
        return out

    # ------------------------------------------------------------------ device-pointer path
    def solve_device(self, start_ptr, end_ptr, n_reads, ref_len, max_coverage, bitmap_ptr,
                     read_off=None, mapq_ptr=None, seq_len_ptr=None, filt=None, params=None,
                     verify=False, find_pairs=False, pair_pass_ptr=None, profile=False,
                     input_on_device=True, len_hint=None, start16_ptr=None):
        """Raw-pointer path.  The bitmap (and pair_pass) are device pointers; the reads are device
        pointers too (tensor.data_ptr()) unless input_on_device=False, in which case they are
        HOST pointers (ideally pinned) and the library does the host->device copies itself."""
        ref_len = np.atleast_1d(np.ascontiguousarray(ref_len, np.uint32))
        ns = len(ref_len)
        if read_off is None:
            read_off = np.array([0, n_reads], np.uint64)
        read_off = np.ascontiguousarray(read_off, np.uint64)
        lh = len_hint or (0, 0)
        # start16_ptr: 16-bit starts instead of start_ptr; end_ptr None/0: fixed-length reads
        rd = _Reads(ns, _ptr(read_off), _ptr(ref_len), start_ptr or None, end_ptr or None, mapq_ptr,
                    seq_len_ptr, int(lh[0]), int(lh[1]), start16_ptr or None)
        fl = None
        keep = []
        if filt is not None:
            a0 = filt.get("amp_start")
            a1 = filt.get("amp_end")
            na = 0 if a0 is None else len(a0)
            a0 = np.ascontiguousarray(a0 if na else [0], np.uint32)
            a1 = np.ascontiguousarray(a1 if na else [0], np.uint32)
            fl = _Filter(filt.get("min_len", 0), filt.get("min_mapq", 0), na, _ptr(a0), _ptr(a1))
            keep += [a0, a1]
        res = _Result()
        filt_off = np.zeros(ns + 1, np.uint64)
        res.kept_bitmap = bitmap_ptr
        res.pair_pass = pair_pass_ptr
        res.filt_off = _ptr(filt_off)
        flags = (1 if input_on_device else 0) | 2 | (4 if verify else 0) | \
                (8 if find_pairs else 0) | (32 if profile else 0)
        prm = _Params(*params) if params is not None else None
        rc = self._lib.gds_solve(self._h, C.byref(rd), C.byref(fl) if fl is not None else None,
                                 int(max_coverage), C.byref(prm) if prm is not None else None,
                                 flags, C.byref(res))
        self._check(rc)
        out = Result({k: getattr(res, k) for k in _SCALARS})
        out["filt_off"] = filt_off
        del keep
        return out

    @staticmethod
    def bitmap_to_indices(bitmap, n_bits):
        """Ascending kept indices (qmcp::Solution order, solver.hpp:13)."""
        L = load_library()
        bitmap = np.ascontiguousarray(bitmap, np.uint32)
        cnt = L.gds_bitmap_to_indices(bitmap.ctypes.data, n_bits, None, 0)
        idx = np.zeros(cnt, np.uint64)
        L.gds_bitmap_to_indices(bitmap.ctypes.data, n_bits, idx.ctypes.data, cnt)
        return idx
