"""B200-native MAP super-resolution gradient engine (drop-in for the hot path of
rteammco/super-resolution: ObjectiveFunction::ComputeAllTerms and the interfaces under it).

The package directory is named after the project (`super-resolution_b200`), which is not a valid
Python identifier: import it with `importlib.import_module("super-resolution_b200")` or through
the `srb200` alias module at the repository root.

  engine      ctypes binding of the C-ABI library libsrb200.so (include/srb200.h)
  build       in-tree nvcc build of that library (sm_100a)
"""
from . import build, engine  # noqa: F401
from .engine import (Engine, MultiEngine, SrbError, SpectralPCA, EnviHeader, envi_read_header, envi_write, REG_NONE, REG_TV, REG_TV3D, REG_BTV, PATH_AUTO,  # noqa: F401
                     PATH_REFERENCE_ORDER, PATH_FUSED, PARTITION_FRAMES, PARTITION_ROWS, device_count, load_library, pin_host,
                     unpin_host, plan, quantize_shift, sample_is_special, dev_alloc, dev_free, ipc_export, ipc_open, ipc_close)
