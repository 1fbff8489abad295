"""gdpathtracing_b200 -- B200-native backend for the GDPathTracing hot path.

Layers (top to bottom):
  nodes.py / scenes.py   Python faces of the host twins + deterministic scene descriptions
  libgdpt_host.so        C++ host layer: GeometryGroup3D / PathTracingCamera twins, BLAS/TLAS builder
  libgdpt_cuda.so        C-ABI boundary (include/gdpt.h): hand-written sm_100a kernels + engine

Importing the package loads both shared libraries; it raises if they are not built.  There is no
CPU implementation of the render path anywhere in this package.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is missing)
from .nodes import GeometryGroup3D, PathTracingCamera, make_camera_block  # noqa: F401
from . import scenes  # noqa: F401

__all__ = ["GeometryGroup3D", "PathTracingCamera", "make_camera_block", "scenes"]
