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


class RowWindow(object):
    """Target of stencil calls that each compute a row window of ONE output raster: the first call
    allocates the full (bands, full_rows, width) output, every call writes its rows at ``r0``
    (raster/spatial.py `_call_stencil`; single-band rasters only -- arrays are contiguous)."""

    def __init__(self, full_rows):
        self.full_rows = int(full_rows)
        self.r0 = 0
        self.out = None


def row_window():
    return getattr(_local, "row_window", None)


@contextlib.contextmanager
def into_row_window(target):
    previous = row_window()
    _local.row_window = target
    try:
        yield target
    finally:
        _local.row_window = previous
