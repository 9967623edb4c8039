"""nairn_mpm_fea_b200: NairnMPM's explicit MPM time step on B200 (sm_100a).

The product is the C-ABI library `libmpmgpu.so` (include/mpmgpu.h, csrc/).  This package is the
Python host side above it: a ctypes binding (`capi`), the host-side set-up the reference driver does
before the time loop (`problem`, `materials`), the multi-GPU slab driver (`slab`) and a writer of the reference's
binary particle archives (`archive`).
"""
from . import capi, materials, problem  # noqa: F401
from .capi import MpmGpu, MpmGpuError, load_library  # noqa: F401
from .problem import Problem  # noqa: F401
