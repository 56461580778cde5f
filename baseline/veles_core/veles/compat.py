import lzma  # noqa: F401
from enum import IntEnum  # noqa: F401


def from_none(exc):
    exc.__cause__ = None
    exc.__suppress_context__ = True
    return exc
