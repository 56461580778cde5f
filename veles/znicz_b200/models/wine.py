"""Wine sample: FC-tanh(8) → softmax(3), hand-wired NNWorkflow.

Parity: /root/reference/samples/Wine/wine.py:66-181 and wine_config.py:43-58
(minibatch 10, lr 0.3, weights_decay 0, pointwise normalisation). The dataset file
of the reference is downloaded from the VelesForge; offline we use the identical UCI
Wine table that ships inside scikit-learn (178×13, 3 classes) or a CSV given in
``root.wine.loader.dataset_file``.
"""
from __future__ import annotations

import numpy

from ..core.config import root
from ..loader.base import TEST, VALID, TRAIN
from ..loader.fullbatch import FullBatchLoader
from ..ops import all2all, gd
from ..ops.nn_units import NNWorkflow, NNSnapshotterToFile
from ..workflow import decision, evaluator

root.wine.update({
    "decision": {"fail_iterations": 200, "max_epochs": 100},
    "snapshotter": {"prefix": "wine", "interval": 10, "time_interval": 0},
    "loader": {"minibatch_size": 10, "dataset_file": None, "force_numpy": False},
    "learning_rate": 0.3,
    "weights_decay": 0.0,
    "layers": [8, 3]})


class WineLoader(FullBatchLoader):
    """Loads the Wine dataset (/root/reference/loader/loader_wine.py:48-66):
    CSV with the class in column 0 (1-based), pointwise normalisation."""
    MAPPING = "wine_loader"

    def __init__(self, workflow, **kwargs):
        kwargs["normalization_type"] = "pointwise"
        super().__init__(workflow, **kwargs)
        self.dataset_file = kwargs.get("dataset_file")

    def load_data(self):
        if self.dataset_file:
            arr = numpy.loadtxt(self.dataset_file, delimiter=",", dtype=numpy.float32)
            data = arr[:, 1:]
            labels = arr[:, 0].ravel().astype(numpy.int32) - 1
        else:
            from sklearn.datasets import load_wine
            ds = load_wine()
            data = ds.data.astype(numpy.float32)
            labels = ds.target.astype(numpy.int32)
        self.original_data.reset(numpy.ascontiguousarray(data, dtype=self.dtype))
        self.original_labels = labels.tolist()
        if not self.testing:
            self.class_lengths[TEST] = self.class_lengths[VALID] = 0
            self.class_lengths[TRAIN] = data.shape[0]
        else:
            self.class_lengths[TEST] = data.shape[0]
            self.class_lengths[VALID] = self.class_lengths[TRAIN] = 0


class WineWorkflow(NNWorkflow):
    """Fully connected NN with softmax loss for the Wine dataset."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        layers = kwargs.get("layers", root.wine.layers)
        self.repeater.link_from(self.start_point)
        self.loader = WineLoader(
            self, minibatch_size=root.wine.loader.minibatch_size,
            force_numpy=root.wine.loader.force_numpy,
            dataset_file=root.wine.loader.get("dataset_file"))
        self.loader.link_from(self.repeater)

        del self.forwards[:]
        for i, layer in enumerate(layers):
            cls = all2all.All2AllTanh if i < len(layers) - 1 else all2all.All2AllSoftmax
            aa = cls(self, output_sample_shape=(layer,), weights_stddev=0.05,
                     bias_stddev=0.05)
            self.forwards.append(aa)
            if i:
                aa.link_from(self.forwards[-2])
                aa.link_attrs(self.forwards[-2], ("input", "output"))
            else:
                aa.link_from(self.loader)
                aa.link_attrs(self.loader, ("input", "minibatch_data"))

        self.evaluator = evaluator.EvaluatorSoftmax(self)
        self.evaluator.link_from(self.forwards[-1])
        self.evaluator.link_attrs(self.forwards[-1], "output", "max_idx")
        self.evaluator.link_attrs(self.loader,
                                  ("batch_size", "minibatch_size"),
                                  ("max_samples_per_epoch", "total_samples"),
                                  ("labels", "minibatch_labels"),
                                  ("offset", "minibatch_offset"),
                                  "class_lengths")

        self.decision = decision.DecisionGD(
            self, fail_iterations=root.wine.decision.fail_iterations,
            max_epochs=root.wine.decision.max_epochs)
        self.decision.link_from(self.evaluator)
        self.decision.link_attrs(self.loader, "minibatch_class", "minibatch_size",
                                 "last_minibatch", "class_lengths", "epoch_ended",
                                 "epoch_number")
        self.decision.link_attrs(
            self.evaluator, ("minibatch_n_err", "n_err"),
            ("minibatch_confusion_matrix", "confusion_matrix"),
            ("minibatch_max_err_y_sum", "max_err_output_sum"))

        self.snapshotter = NNSnapshotterToFile(
            self, prefix=root.wine.snapshotter.prefix,
            directory=root.common.dirs.snapshots, compression="",
            interval=root.wine.snapshotter.interval,
            time_interval=root.wine.snapshotter.time_interval)
        self.snapshotter.link_from(self.decision)
        self.snapshotter.link_attrs(self.decision, ("suffix", "snapshot_suffix"))
        self.snapshotter.gate_skip = ~self.loader.epoch_ended
        self.snapshotter.skip = ~self.decision.improved

        self.end_point.link_from(self.snapshotter)
        self.end_point.gate_block = ~self.decision.complete

        self.gds[:] = (None,) * len(self.forwards)
        self.gds[-1] = gd.GDSoftmax(self) \
            .link_from(self.snapshotter) \
            .link_attrs(self.evaluator, "err_output") \
            .link_attrs(self.forwards[-1], "output", "input", "weights", "bias") \
            .link_attrs(self.loader, ("batch_size", "minibatch_size"))
        self.gds[-1].gate_skip = self.decision.gd_skip
        self.gds[-1].gate_block = self.decision.complete
        for i in range(len(self.forwards) - 2, -1, -1):
            self.gds[i] = gd.GDTanh(self) \
                .link_from(self.gds[i + 1]) \
                .link_attrs(self.gds[i + 1], ("err_output", "err_input")) \
                .link_attrs(self.forwards[i], "output", "input", "weights", "bias") \
                .link_attrs(self.loader, ("batch_size", "minibatch_size"))
            self.gds[i].gate_skip = self.decision.gd_skip
        for g, f in zip(self.gds, self.forwards):
            g.forward_unit = f
        self.gds[0].need_err_input = False
        self.repeater.link_from(self.gds[0])
        self.loader.gate_block = self.decision.complete

    def initialize(self, learning_rate=None, weights_decay=None, device=None, **kwargs):
        if learning_rate is None:
            learning_rate = root.wine.learning_rate
        if weights_decay is None:
            weights_decay = root.wine.weights_decay
        return super().initialize(learning_rate=learning_rate,
                                  weights_decay=weights_decay, device=device, **kwargs)


def run(load, main):
    load(WineWorkflow, layers=root.wine.layers)
    main(learning_rate=root.wine.learning_rate, weights_decay=root.wine.weights_decay)
