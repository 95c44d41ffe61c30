"""amgpu — B200-native bulk change-replay engine behind automerge-classic's Backend API.

    from automerge_classic_b200 import Backend          # init / applyChanges / getPatch / ... (backend/index.js:1-8)

The module object `Backend` mirrors the module a caller hands to `Automerge.setDefaultBackend()`.
"""
from .backend import Backend as _Facade
from .engine import GpuBackendDoc, AmgError, Unsupported

Backend = _Facade(GpuBackendDoc)
__all__ = ['Backend', 'GpuBackendDoc', 'AmgError', 'Unsupported']
