from .pb import *  # noqa
