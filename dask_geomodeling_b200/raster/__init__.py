from .base import RasterBlock  # NOQA
from .elemwise import *  # NOQA
from .misc import *  # NOQA
from .sources import *  # NOQA
