"""kit4b_b200 - B200-native (sm_100a) replacement for the `ngskit4b hammings` hot path.

The product is the C-ABI shared library `libk4bhamm.so` (CUDA kernels + extern "C" boundary,
sources in kit4b_b200/csrc, interface in include/k4b_hamm.h) and the C++ host front end
(`kit4b_b200/bin/k4b_hammings`).  This package is a thin ctypes mirror of that boundary for
tests and bench.py; it has no CPU implementation and raises if the library is missing.
"""
from .hamm import (  # noqa: F401
    K4BError,
    Packed,
    allpairs_min_device,
    exhaustive,
    set_engine,
    get_engine,
    exhaustive_shard,
    gpu_count,
    gpu_init,
    gpu_shutdown,
    lib_path,
    load_lib,
    microbench_intpipe,
    targeted,
)
