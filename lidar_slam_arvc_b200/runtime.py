"""Process-wide engine handle used by the drop-in KeyFrame / KeyFrameManager classes.

One process drives one GPU: the device is LOCAL_RANK (torchrun) or ARVC_DEVICE, default 0.
`set_engine` lets a caller share an engine it created itself (several managers on one context) or inject a test
double; the default is always the CUDA engine and construction fails loudly without a GPU — no CPU fallback.
"""
import itertools
import os

_engine = None
_ids = itertools.count(1)


def get_engine():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(int(os.environ.get("ARVC_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _engine


def set_engine(engine):
    global _engine, _loader
    if _loader is not None:
        _loader.close()
        _loader = None
    _engine = engine


_loader = None


def get_loader():
    """The process-wide PCD loader (pinned staging + read-ahead) bound to the current engine."""
    global _loader
    if _loader is None:
        from .loader import ScanLoader
        _loader = ScanLoader(get_engine())
    return _loader


def new_scan_id():
    return next(_ids)
