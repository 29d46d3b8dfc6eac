"""Per-thread evaluation state shared by the blocks and the graph optimiser."""
import contextlib
import threading

_local = threading.local()


def keep_on_device():
    """True while a graph is being computed by ``core.compute``: rasters that
    flow between tasks stay in HBM (``DeviceArray``) instead of numpy arrays."""
    return getattr(_local, "keep", False)


@contextlib.contextmanager
def device_resident(flag=True):
    previous = keep_on_device()
    _local.keep = flag
    try:
        yield
    finally:
        _local.keep = previous
