from .base import RasterBlock  # NOQA
from .elemwise import *  # NOQA
from .misc import *  # NOQA
from .sources import *  # NOQA
from .spatial import *  # NOQA
from .temporal import *  # NOQA
from .parallelize import *  # NOQA
