"""Approximator: FC regression between two MATLAB arrays (decimated → original
apertures). Parity: /root/reference/tests/research/Approximator/approximator.py:67-300
(``.mat`` variables via scipy.io, ``mean_disp`` on inputs and targets, layers [810, 9]).
``.npy`` / ``.npz`` files are accepted too."""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..loader.base import UserLoaderRegistry, TEST, VALID, TRAIN, LoaderError
from ..loader.fullbatch import FullBatchLoaderMSE
from .fc_mse import FullyConnectedMSEWorkflow

_d = os.path.join(str(root.common.dirs.datasets), "approximator")
root.approximator.update({
    "decision": {"fail_iterations": 1000, "max_epochs": 1000000000},
    "snapshotter": {"prefix": "approximator"},
    "loader_name": "approximator_loader",
    "loader": {"minibatch_size": 100,
               "train_paths": [os.path.join(_d, "all_dec_apertures.mat")],
               "target_paths": [os.path.join(_d, "all_org_apertures.mat")],
               "normalization_type": "mean_disp", "target_normalization_type": "mean_disp",
               "validation_ratio": 0.15},
    "learning_rate": 0.0001,
    "weights_decay": 0.00005,
    "layers": [810, 9]})


def load_matrix(path):
    if path.endswith(".npy"):
        return numpy.load(path)
    if path.endswith(".npz"):
        z = numpy.load(path)
        return z[z.files[0]]
    import scipy.io
    mat = scipy.io.loadmat(path)
    for key, val in mat.items():
        if not key.startswith("_"):
            return numpy.asarray(val)
    raise LoaderError("Could not find a variable to import in %s" % path)


class ApproximatorLoader(FullBatchLoaderMSE):
    MAPPING = "approximator_loader"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.test_paths = kwargs.get("test_paths", [])
        self.validation_paths = kwargs.get("validation_paths", [])
        self.train_paths = kwargs.get("train_paths", [])
        self.target_paths = kwargs["target_paths"]

    def load_data(self):
        chunks = []
        for cls, paths in ((TEST, self.test_paths), (VALID, self.validation_paths),
                           (TRAIN, self.train_paths)):
            n = 0
            for p in paths or ():
                m = load_matrix(p).astype(self.dtype)
                chunks.append(m.reshape(m.shape[0], -1))
                n += m.shape[0]
            self.class_lengths[cls] = n
        if not chunks:
            raise LoaderError("no input matrices given")
        data = numpy.concatenate(chunks)
        targets = numpy.concatenate([
            (lambda m: m.reshape(m.shape[0], -1))(load_matrix(p).astype(self.dtype))
            for p in self.target_paths])
        if targets.shape[0] != data.shape[0]:
            raise LoaderError("targets (%d rows) do not match inputs (%d rows)" % (
                targets.shape[0], data.shape[0]))
        self.original_data.reset(data)
        self.original_targets.reset(targets)


class ApproximatorWorkflow(FullyConnectedMSEWorkflow):
    def __init__(self, workflow, **kwargs):
        cfg = dict(root.approximator.loader.to_dict())
        cfg.update(kwargs.pop("loader_config", {}))
        factory = UserLoaderRegistry.get_factory(
            kwargs.pop("loader_name", root.approximator.loader_name), **cfg)
        super().__init__(workflow, root.approximator, factory, **kwargs)


def build(launcher=None, **kwargs):
    from ..core.workflow import DummyLauncher
    return ApproximatorWorkflow(launcher or DummyLauncher(), **kwargs)


def run(load, main):
    load(ApproximatorWorkflow, layers=root.approximator.layers)
    main(learning_rate=root.approximator.learning_rate,
         weights_decay=root.approximator.weights_decay)
