"""Forward / gradient-descent units. Importing the package registers every layer type
of the ``layers`` DSL."""
from . import nn_units  # noqa
from . import (all2all, gd, conv, gd_conv, pooling, gd_pooling, depooling,  # noqa
               activation, dropout, normalization, cutter, multiplier, summator,
               weights_zerofilling)
import importlib as _il
for _m in ("deconv", "gd_deconv", "lstm", "kohonen", "rbm_units", "rprop_gd",
           "resizable_all2all"):
    try:
        _il.import_module("." + _m, __name__)
    except ModuleNotFoundError as _e:   # module not written yet in this round
        if _m not in str(_e):
            raise
from . import lstm_seq  # noqa: F401
