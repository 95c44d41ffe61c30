"""amgpu — B200-native bulk change-replay engine behind automerge-classic's Backend API.

    from automerge_classic_b200 import Backend          # init / applyChanges / getPatch / ... (backend/index.js:1-8)

The module object `Backend` mirrors the module a caller hands to `Automerge.setDefaultBackend()`.
"""
from .backend import Backend as _Facade
from .engine import GpuBackendDoc, AmgError, Unsupported

from . import sync as _sync


def bind_sync(facade):
    """Adds the sync functions of backend/index.js:2, 7 to a Backend facade (the reference's sync.js is tied to its own backend)."""
    s = _sync.Sync(facade)
    facade.generateSyncMessage, facade.receiveSyncMessage = s.generateSyncMessage, s.receiveSyncMessage
    for name in ('encodeSyncMessage', 'decodeSyncMessage', 'encodeSyncState', 'decodeSyncState', 'initSyncState'):
        setattr(facade, name, getattr(_sync, name))
    return facade


Backend = bind_sync(_Facade(GpuBackendDoc))
__all__ = ['Backend', 'GpuBackendDoc', 'AmgError', 'Unsupported', 'bind_sync']
