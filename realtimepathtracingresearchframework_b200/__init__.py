"""realtimepathtracingresearchframework_b200 -- a B200 (sm_100a) wavefront path tracer behind the plugin surface of
intel/RealTimePathTracingResearchFramework (`rptr --backend cuda`).

  csrc/        hand-written CUDA kernels + the C ABI (include/rptr_cuda.h) -> librptr_cuda.so
  backend.py   RenderCuda: host mirror of RenderBackend / RaytraceBackend over that ABI
  scenes.py    procedural scenes of BASELINE.json's configs in the reference's in-memory scene model
  types.py     ctypes mirrors of the boundary PODs (include/rptr_types.h)
"""
from . import types  # noqa: F401
from .backend import (RenderConfiguration, RenderCuda, RptrError, create_cuda_backend, load_library, load_pointset_tables, load_sky_fit,  # noqa: F401
                      read_pfm, write_pfm)
