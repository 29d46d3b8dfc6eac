from .base import *  # NOQA
from .sources import *  # NOQA
from .aggregate import *  # NOQA
