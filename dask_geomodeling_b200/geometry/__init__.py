from .base import *  # NOQA
