from .graphs import *  # NOQA
