"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/gds.h
declares, host-only helpers work, and there is NO CPU fallback (create fails without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "gds.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gds_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    fns = declared_functions()
    assert {"gds_create", "gds_destroy", "gds_solve", "gds_last_error", "gds_set_stream",
            "gds_abi_version", "gds_bitmap_to_indices"} <= set(fns)


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.lib_path())
    for fn in declared_functions():
        assert hasattr(lib, fn), "libgds_b200.so does not export %s" % fn
    assert lib.gds_abi_version() == 7
    assert sorted(pkg.exported_symbols()) == declared_functions()


def test_product_does_not_link_or_import_the_oracle(pkg):
    # the oracle is test infrastructure: nothing under the package may reference it
    pk = os.path.join(ROOT, "genome-downsampler_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"import\s+pyoracle|from\s+pyoracle|#include\s*[\"<][^\n]*oracle"
                                     r"|\borc_[a-z_]+\s*\(|load_oracle", src), f
    out = os.popen("ldd %s" % pkg.lib_path()).read()
    assert "oracle" not in out


def test_bitmap_to_indices_host_helper(pkg):
    bm = np.array([0b1011, 0x80000000, 0x1], np.uint32)
    idx = pkg.Solver.bitmap_to_indices(bm, 65)
    assert idx.tolist() == [0, 1, 3, 63, 64]
    assert pkg.Solver.bitmap_to_indices(bm, 64).tolist() == [0, 1, 3, 63]


def test_no_cpu_fallback_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.GdsError) as ei:
        pkg.Solver(0)
    assert ei.value.code == 3  # GDS_ERR_CUDA


def test_ctypes_structs_match_the_c_header(pkg, tmp_path):
    # binding.py mirrors include/gds.h by hand: sizes and the offsets the binding relies on must be
    # what a C compiler sees (a drifted struct would corrupt results silently)
    import subprocess
    from genome_downsampler_b200 import binding as B
    probes = [("gds_reads", None), ("gds_filter", None), ("gds_params", None), ("gds_result", None),
              ("gds_kernel_stat", None), ("gds_reads", "start16"), ("gds_reads", "len_min"),
              ("gds_params", "bundle_mode"), ("gds_params", "algorithm"), ("gds_params", "schedule"), ("gds_result", "n_reads_in"), ("gds_result", "fstar"),
              ("gds_result", "kernel_launches"), ("gds_result", "bundle_path"), ("gds_result", "ms_h2d"),
              ("gds_result", "ms_total")]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gds.h"', 'int main(void) {']
    for st, fld in probes:
        expr = "sizeof(%s)" % st if fld is None else "offsetof(%s, %s)" % (st, fld)
        lines.append('  printf("%%zu\\n", (size_t)%s);' % expr)
    lines += ['  return 0;', '}']
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirror = {"gds_reads": B._Reads, "gds_filter": B._Filter, "gds_params": B._Params,
              "gds_result": B._Result, "gds_kernel_stat": B._KStat}
    want = [ctypes.sizeof(mirror[st]) if fld is None else getattr(mirror[st], fld).offset
            for st, fld in probes]
    assert got == want, list(zip(probes, got, want))

