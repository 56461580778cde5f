"""Reference-style module aliases (``veles.*`` core names, ``veles.znicz.*``)."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import sys

_B = "veles.znicz_b200."
# reference module path -> module in this package
ALIASES = {
    # core (absent in the reference tree)
    "veles.config": _B + "core.config",
    "veles.mutable": _B + "core.mutable",
    "veles.memory": _B + "core.memory",
    "veles.prng": _B + "core.prng",
    "veles.units": _B + "core.units",
    "veles.workflow": _B + "core.workflow",
    "veles.plumbing": _B + "core.workflow",
    "veles.dummy": _B + "core.workflow",
    "veles.avatar": _B + "core.avatar",
    "veles.accelerated_units": _B + "core.accelerated_units",
    "veles.backends": _B + "core.backends",
    "veles.distributable": _B + "core.distributable",
    "veles.result_provider": _B + "core.result_provider",
    "veles.normalization": _B + "core.normalization",
    "veles.snapshotter": _B + "core.snapshotter",
    "veles.loader": _B + "loader",
    "veles.loader.base": _B + "loader.base",
    "veles.loader.fullbatch": _B + "loader.fullbatch",
    "veles.loader.image": _B + "loader.image",
    "veles.loader.file_image": _B + "loader.image",
    "veles.loader.fullbatch_image": _B + "loader.image",
    "veles.loader.saver": _B + "loader.saver",
    "veles.genetics": _B + "core.genetics",
    "veles.launcher": _B + "launcher",
    "veles.plotting_units": _B + "utils.plotting_units",
    "veles.downloader": _B + "utils.downloader",
    "veles.publishing": _B + "utils.publishing",
    "veles.interaction": _B + "utils.interaction",
    "veles.mean_disp_normalizer": _B + "utils.mean_disp_normalizer",
    "veles.input_joiner": _B + "core.input_joiner",
    # znicz units
    "veles.znicz": _B[:-1],
    "veles.znicz.nn_units": _B + "ops.nn_units",
    "veles.znicz.all2all": _B + "ops.all2all",
    "veles.znicz.gd": _B + "ops.gd",
    "veles.znicz.conv": _B + "ops.conv",
    "veles.znicz.gd_conv": _B + "ops.gd_conv",
    "veles.znicz.deconv": _B + "ops.deconv",
    "veles.znicz.gd_deconv": _B + "ops.gd_deconv",
    "veles.znicz.pooling": _B + "ops.pooling",
    "veles.znicz.gd_pooling": _B + "ops.gd_pooling",
    "veles.znicz.depooling": _B + "ops.depooling",
    "veles.znicz.activation": _B + "ops.activation",
    "veles.znicz.dropout": _B + "ops.dropout",
    "veles.znicz.normalization": _B + "ops.normalization",
    "veles.znicz.cutter": _B + "ops.cutter",
    "veles.znicz.multiplier": _B + "ops.multiplier",
    "veles.znicz.summator": _B + "ops.summator",
    "veles.znicz.lstm": _B + "ops.lstm",
    "veles.znicz.kohonen": _B + "ops.kohonen",
    "veles.znicz.rbm_units": _B + "ops.rbm_units",
    "veles.znicz.rprop_gd": _B + "ops.rprop_gd",
    "veles.znicz.resizable_all2all": _B + "ops.resizable_all2all",
    "veles.znicz.weights_zerofilling": _B + "ops.weights_zerofilling",
    "veles.znicz.evaluator": _B + "workflow.evaluator",
    "veles.znicz.decision": _B + "workflow.decision",
    "veles.znicz.lr_adjust": _B + "workflow.lr_adjust",
    "veles.znicz.nn_rollback": _B + "workflow.nn_rollback",
    "veles.znicz.standard_workflow": _B + "workflow.standard_workflow",
    "veles.znicz.standard_workflow_base": _B + "workflow.standard_workflow_base",
    "veles.znicz.diff_stats": _B + "utils.diff_stats",
    "veles.znicz.accumulator": _B + "utils.accumulator",
    "veles.znicz.image_saver": _B + "utils.image_saver",
    "veles.znicz.labels_printer": _B + "utils.labels_printer",
    "veles.znicz.nn_plotting_units": _B + "utils.nn_plotting_units",
    "veles.znicz.diversity": _B + "utils.diversity",
    "veles.znicz.loader": _B + "loader",
    "veles.znicz.loader.loader_lmdb": _B + "loader.loader_lmdb",
    "veles.znicz.loader.loader_stl": _B + "loader.loader_stl",
    "veles.znicz.loader.imagenet_loader": _B + "loader.imagenet_loader",
    "veles.znicz.loader.caffe": _B + "loader.caffe",
    "veles.znicz.loader.caffe.protobuf2": _B + "loader.caffe.protobuf2",
}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in ALIASES:
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        mod = importlib.import_module(ALIASES[spec.name])
        return mod

    def exec_module(self, module):
        pass


_installed = False


def install():
    global _installed
    if _installed:
        return
    sys.meta_path.insert(0, _AliasFinder())
    _installed = True
