"""genome-downsampler_b200 — B200-native quasi-MCP downsampler (one hot path of
migoox/genome-downsampler).  Python here is a thin ctypes view of the C ABI in include/gds.h,
used by tests/ and bench.py; the product is libgds_b200.so and the C++ host mirror in host/.

The directory name has a hyphen (it is the name the task fixes), so load it with
``__graft_entry__.load_package()`` which registers it as ``genome_downsampler_b200``.
"""
from .binding import (  # noqa: F401
    GdsError, Solver, Result, lib_path, load_library, exported_symbols, GDS_FLAGS,
)
from .pipeline import ChunkedSolver  # noqa: F401
