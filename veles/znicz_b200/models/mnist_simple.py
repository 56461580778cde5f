"""MnistSimple: hand-wired FC-tanh(100) → softmax(10) MNIST workflow with gradient
statistics (``DiffStats``), a Shell unit and error plotters in the loop.
Parity: /root/reference/tests/research/MnistSimple/mnist.py:56-215."""
from __future__ import annotations

import os

from ..core.config import root
from ..loader.base import UserLoaderRegistry
from ..ops import all2all, gd
from ..ops.nn_units import NNWorkflow, NNSnapshotterToFile
from ..utils import plotting_units
from ..utils.diff_stats import DiffStats
from ..utils.interaction import Shell
from ..workflow import decision, evaluator
from . import mnist  # noqa: F401  (registers mnist_loader)

root.mnist_simple.update({
    "all2all": {"weights_stddev": 0.05},
    "decision": {"fail_iterations": 100, "max_epochs": 1000000000},
    "snapshotter": {"prefix": "mnist", "time_interval": 15},
    "loader_name": "mnist_loader",
    "loader": {"minibatch_size": 60, "force_numpy": False, "normalization_type": "linear"},
    "learning_rate": 0.03,
    "weights_decay": 0.0005,
    "factor_ortho": 0.0,
    "diff_stats_file": "stats.pickle",
    "layers": [100, 10]})


class MnistSimpleWorkflow(NNWorkflow):
    def __init__(self, workflow, **kwargs):
        cfg = root.mnist_simple
        layers = kwargs.get("layers") or cfg.layers
        super().__init__(workflow, **kwargs)
        self.repeater.link_from(self.start_point)
        lcfg = dict(cfg.loader.to_dict())
        lcfg.update(kwargs.get("loader_config", {}))
        self.loader = UserLoaderRegistry.get_factory(
            kwargs.get("loader_name", cfg.loader_name), **lcfg)(self)
        self.loader.link_from(self.repeater)
        del self.forwards[:]
        for i, layer in enumerate(layers):
            cls = all2all.All2AllTanh if i < len(layers) - 1 else all2all.All2AllSoftmax
            aa = cls(self, output_sample_shape=[layer],
                     weights_stddev=cfg.all2all.weights_stddev)
            self.forwards.append(aa)
            src = self.forwards[-2] if i else self.loader
            aa.link_from(src)
            aa.link_attrs(src, ("input", "output" if i else "minibatch_data"))
        self.evaluator = evaluator.EvaluatorSoftmax(self)
        self.evaluator.link_from(self.forwards[-1])
        self.evaluator.link_attrs(self.forwards[-1], "output", "max_idx")
        self.evaluator.link_attrs(self.loader, ("labels", "minibatch_labels"),
                                  ("batch_size", "minibatch_size"),
                                  ("max_samples_per_epoch", "total_samples"),
                                  ("offset", "minibatch_offset"), "class_lengths")
        self.decision = decision.DecisionGD(
            self, fail_iterations=cfg.decision.fail_iterations,
            max_epochs=cfg.decision.max_epochs)
        self.decision.link_from(self.evaluator)
        self.decision.link_attrs(self.loader, "minibatch_class", "minibatch_size",
                                 "last_minibatch", "class_lengths", "epoch_ended",
                                 "epoch_number")
        self.decision.link_attrs(
            self.evaluator, ("minibatch_n_err", "n_err"),
            ("minibatch_confusion_matrix", "confusion_matrix"),
            ("minibatch_max_err_y_sum", "max_err_output_sum"))
        self.snapshotter = NNSnapshotterToFile(
            self, prefix=cfg.snapshotter.prefix, directory=root.common.dirs.snapshots,
            time_interval=cfg.snapshotter.time_interval)
        self.snapshotter.link_from(self.decision)
        self.snapshotter.link_attrs(self.decision, ("suffix", "snapshot_suffix"))
        self.snapshotter.gate_skip = ~self.loader.epoch_ended
        self.snapshotter.skip = ~self.decision.improved
        self.ipython = Shell(self, enabled=kwargs.get("shell", False))
        self.ipython.link_from(self.snapshotter)
        self.ipython.gate_skip = ~self.decision.epoch_ended

        self.gds[:] = (None,) * len(self.forwards)
        self.gds[-1] = gd.GDSoftmax(self, learning_rate=cfg.learning_rate,
                                    weights_decay=cfg.weights_decay) \
            .link_from(self.ipython) \
            .link_attrs(self.evaluator, "err_output") \
            .link_attrs(self.forwards[-1], "output", "input", "weights", "bias") \
            .link_attrs(self.loader, ("batch_size", "minibatch_size"))
        self.gds[-1].gate_skip = self.decision.gd_skip
        for i in range(len(self.forwards) - 2, -1, -1):
            self.gds[i] = gd.GDTanh(self, learning_rate=cfg.learning_rate,
                                    weights_decay=cfg.weights_decay,
                                    factor_ortho=cfg.factor_ortho) \
                .link_from(self.gds[i + 1]) \
                .link_attrs(self.gds[i + 1], ("err_output", "err_input")) \
                .link_attrs(self.forwards[i], "output", "input", "weights", "bias") \
                .link_attrs(self.loader, ("batch_size", "minibatch_size"))
            self.gds[i].gate_skip = self.decision.gd_skip
        for g, f in zip(self.gds, self.forwards):
            g.forward_unit = f
        self.gds[0].need_err_input = False
        # gradient statistics (sum |delta| of every layer's gradient per step)
        self.diff_stats = DiffStats(
            self, arrays={u: ("gradient_weights",) for u in self.gds},
            file_name=os.path.join(str(root.common.dirs.cache), cfg.diff_stats_file))
        self.diff_stats.link_from(self.gds[0])
        self.diff_stats.gate_skip = self.decision.gd_skip
        self.repeater.link_from(self.diff_stats)
        self.repeater.gate_block = self.decision.complete
        self.end_point.link_from(self.gds[0])
        self.end_point.gate_block = ~self.decision.complete
        self.loader.gate_block = self.decision.complete
        self.slaves_plotter = plotting_units.SlaveStats(self)
        self.slaves_plotter.link_from(self.decision).gate_block = self.decision.complete
        self.plt = []
        for i, style in enumerate(("g-", "r-", "k-")):
            p = plotting_units.AccumulatingPlotter(self, name="Errors %d" % i,
                                                   plot_style=style)
            p.link_attrs(self.decision, ("input", "epoch_n_err_pt"))
            p.input_field = i
            p.link_from(self.plt[-1] if self.plt else self.decision)
            p.gate_block = ~self.decision.epoch_ended | self.decision.complete
            self.plt.append(p)


def build(launcher=None, **kwargs):
    from ..core.workflow import DummyLauncher
    return MnistSimpleWorkflow(launcher or DummyLauncher(), **kwargs)


def run(load, main):
    load(MnistSimpleWorkflow, layers=root.mnist_simple.layers)
    main()
