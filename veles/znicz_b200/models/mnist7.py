"""Mnist7: digits → 7-segment display code (MSE loss, nearest-target accuracy).
Parity: /root/reference/tests/research/Mnist7/mnist7.py:64-210, mnist7_config.py:43-56."""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..loader.base import UserLoaderRegistry
from ..loader.fullbatch import FullBatchLoaderMSE
from .fc_mse import FullyConnectedMSEWorkflow
from ..loader.synthetic import SyntheticMnistLoader
from .mnist import MnistLoader

root.mnist7.update({
    "decision": {"fail_iterations": 25, "max_epochs": 1000000},
    "snapshotter": {"prefix": "mnist7", "time_interval": 0, "interval": 1},
    "loader_name": "mnist7_loader",
    "loader": {"minibatch_size": 60, "force_numpy": False, "normalization_type": "linear",
               "data_path": os.path.join(str(root.common.dirs.datasets), "MNIST"),
               "target_normalization_type": "none", "target_normalization_parameters": {}},
    "weights_plotter": {"limit": 25},
    "learning_rate": 0.0001,
    "weights_decay": 0.00005,
    "layers": [100, 100, 7]})

SEGMENTS = numpy.array(
    [[1, 1, 1, -1, 1, 1, 1], [-1, -1, 1, -1, -1, 1, -1], [1, -1, 1, 1, 1, -1, 1],
     [1, -1, 1, 1, -1, 1, 1], [-1, 1, 1, 1, -1, 1, -1], [1, 1, -1, 1, -1, 1, 1],
     [1, 1, -1, 1, 1, 1, 1], [1, 1, 1, -1, -1, 1, -1], [1, 1, 1, 1, 1, 1, 1],
     [1, 1, 1, 1, -1, 1, 1]], dtype=numpy.float32)


class SegmentTargetsMixin(object):
    """Turns a labelled digits loader into an MSE loader with 7-segment targets."""

    def load_data(self):
        super().load_data()
        self.class_targets.reset(SEGMENTS.astype(self.dtype))
        labels = numpy.asarray(self.original_labels, dtype=numpy.int64)
        self.original_targets.reset(self.class_targets.mem[labels])


class Mnist7Loader(SegmentTargetsMixin, MnistLoader, FullBatchLoaderMSE):
    MAPPING = "mnist7_loader"


class SyntheticMnist7Loader(SegmentTargetsMixin, SyntheticMnistLoader, FullBatchLoaderMSE):
    """Offline stand-in (no MNIST files in the sandbox): synthetic digits, same targets."""
    MAPPING = "synthetic_mnist7"


class Mnist7Workflow(FullyConnectedMSEWorkflow):
    def __init__(self, workflow, **kwargs):
        name = kwargs.pop("loader_name", root.mnist7.loader_name)
        cfg = dict(root.mnist7.loader.to_dict())
        cfg.update(kwargs.pop("loader_config", {}))
        factory = UserLoaderRegistry.get_factory(name, **cfg)
        kwargs.setdefault("use_class_targets", True)
        super().__init__(workflow, root.mnist7, factory, **kwargs)


def build(launcher=None, **kwargs):
    from ..core.workflow import DummyLauncher
    return Mnist7Workflow(launcher or DummyLauncher(), **kwargs)


def run(load, main):
    load(Mnist7Workflow, layers=root.mnist7.layers)
    main(learning_rate=root.mnist7.learning_rate, weights_decay=root.mnist7.weights_decay)
