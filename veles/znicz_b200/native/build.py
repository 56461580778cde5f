"""Build of the native forward-only runtime (placeholder until native/ sources land)."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=True):
    src = os.path.join(HERE, "src")
    if not os.path.isdir(src):
        return None
    from . import _build_impl
    return _build_impl.build(verbose=verbose)
