"""B200-native (sm_100a) implementation of the 3D acoustic FDTD hot path of
ycnliu/Accelerated-3D-Acoustic-FDTD-Kernel behind the reference's own operator ABI.

The product is ``libfdtd_b200.so`` (csrc/, C ABI in include/fdtd_b200.h).  This package is the
thin Python host mirror used by the tests and bench: same names and argument meaning as the
reference's ``Kernel_*`` entry points.  There is no CPU fallback: importing works without a GPU,
but every compute call goes to the CUDA library and raises if it cannot run.
"""
from .host import (  # noqa: F401
    HALO,
    WARMUP_STEPS,
    Checksum,
    Dataobj,
    FdtdError,
    Geometry,
    Plan,
    Profiler,
    FDTD_SetRuntimeConfig,
    Kernel_B200,
    Kernel_CUDA_Optimized,
    build,
    exported_symbols,
    fill_ricker,
    fill_source_coords,
    lib,
    lib_path,
    run_slabs,
    slab_source_cells,
    slab_source_cells2,
    source_table,
    write_benchmark_csv,
)
from . import slab  # noqa: E402,F401
from .slab import LocalSlabs, SlabRun, partition  # noqa: E402,F401
