"""MnistAE: one-layer convolutional autoencoder with tied weights.

Parity: /root/reference/tests/research/MnistAE/mnist_ae.py:69-236, mnist_ae_config.py:43-62:
Conv(5 kernels 5x5, no bias) → StochasticAbsPooling(3x3 / 2) → depooling done by a
``GDMaxAbsPooling`` whose err_output is the pooled output → Deconv sharing the conv weights
→ EvaluatorMSE against the input image → DecisionMSE → GDDeconv (trains the shared weights).
"""
from __future__ import annotations

import os

from ..core.config import root
from ..core.normalization import NoneNormalizer
from ..loader.base import UserLoaderRegistry
from ..ops import conv, deconv, gd_deconv, gd_pooling, pooling
from ..ops.nn_units import NNWorkflow, NNSnapshotterToFile
from ..utils import nn_plotting_units, plotting_units
from ..workflow import decision, evaluator
from . import mnist  # noqa: F401

root.mnist_ae.update({
    "all2all": {"weights_stddev": 0.05},
    "decision": {"fail_iterations": 20, "max_epochs": 1000000000},
    "snapshotter": {"prefix": "mnist_ae", "time_interval": 0, "compression": "",
                    "interval": 1},
    "loader_name": "mnist_loader",
    "loader": {"minibatch_size": 100, "force_numpy": False, "normalization_type": "linear",
               "data_path": os.path.join(str(root.common.dirs.datasets), "MNIST")},
    "learning_rate": 0.000001,
    "weights_decay": 0.00005,
    "gradient_moment": 0.00001,
    "weights_plotter": {"limit": 16},
    "pooling": {"kx": 3, "ky": 3, "sliding": (2, 2)},
    "include_bias": False,
    "unsafe_padding": True,
    "n_kernels": 5,
    "kx": 5,
    "ky": 5})


class MnistAEWorkflow(NNWorkflow):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        cfg = root.mnist_ae
        self.repeater.link_from(self.start_point)
        lcfg = dict(cfg.loader.to_dict())
        lcfg.update(kwargs.get("loader_config", {}))
        self.loader = UserLoaderRegistry.get_factory(
            kwargs.get("loader_name", cfg.loader_name), **lcfg)(self)
        self.loader.link_from(self.repeater)
        self.target_normalizer = NoneNormalizer()

        self.conv = conv.Conv(self, n_kernels=cfg.n_kernels, kx=cfg.kx, ky=cfg.ky,
                              weights_filling="uniform", include_bias=cfg.include_bias)
        self.conv.link_from(self.loader)
        self.conv.link_attrs(self.loader, ("input", "minibatch_data"))

        self.pool = pooling.StochasticAbsPooling(
            self, kx=cfg.pooling.kx, ky=cfg.pooling.ky, sliding=tuple(cfg.pooling.sliding))
        self.pool.link_from(self.conv)
        self.pool.link_attrs(self.conv, ("input", "output"))
        self.pool.link_attrs(self.loader, "minibatch_class")

        self.depool = gd_pooling.GDMaxAbsPooling(
            self, kx=cfg.pooling.kx, ky=cfg.pooling.ky, sliding=tuple(cfg.pooling.sliding))
        self.depool.link_from(self.pool)
        self.depool.link_attrs(self.pool, "input", "input_offset", ("err_output", "output"))

        self.deconv = deconv.Deconv(self, unsafe_padding=cfg.unsafe_padding)
        self.deconv.link_from(self.depool)
        self.deconv.link_attrs(self.conv, "weights")
        self.deconv.link_conv_attrs(self.conv)
        self.deconv.link_attrs(self.depool, ("input", "err_input"))
        self.deconv.link_attrs(self.conv, ("output_shape_source", "input"))
        del self.forwards[:]
        self.forwards.extend([self.conv, self.pool, self.depool, self.deconv])

        self.evaluator = evaluator.EvaluatorMSE(self)
        self.evaluator.link_from(self.deconv)
        self.evaluator.link_attrs(self.deconv, "output")
        self.evaluator.link_attrs(self.loader, ("batch_size", "minibatch_size"),
                                  ("target", "minibatch_data"))
        self.evaluator.link_attrs(self, ("normalizer", "target_normalizer"))

        self.decision = decision.DecisionMSE(
            self, fail_iterations=cfg.decision.fail_iterations,
            max_epochs=kwargs.get("max_epochs", cfg.decision.max_epochs))
        self.decision.link_from(self.evaluator)
        self.decision.link_attrs(self.loader, "minibatch_class", "minibatch_size",
                                 "last_minibatch", "class_lengths", "epoch_ended",
                                 "epoch_number")
        self.decision.link_attrs(self.evaluator, ("minibatch_metrics", "metrics"))

        snap = cfg.snapshotter
        self.snapshotter = NNSnapshotterToFile(
            self, prefix=snap.prefix, compression=snap.compression,
            directory=root.common.dirs.snapshots, time_interval=snap.time_interval,
            interval=snap.interval)
        self.snapshotter.link_from(self.decision)
        self.snapshotter.link_attrs(self.decision, ("suffix", "snapshot_suffix"))
        self.snapshotter.gate_skip = ~self.loader.epoch_ended | ~self.decision.improved
        self.end_point.link_from(self.snapshotter)
        self.end_point.gate_block = ~self.decision.complete

        self.gd_deconv = gd_deconv.GDDeconv(
            self, learning_rate=kwargs.get("learning_rate", cfg.learning_rate),
            weights_decay=cfg.weights_decay, gradient_moment=cfg.gradient_moment)
        self.gd_deconv.link_attrs(self.evaluator, "err_output")
        self.gd_deconv.link_attrs(self.deconv, "weights", "input", "hits", "n_kernels",
                                  "kx", "ky", "sliding", "padding", "unpack_size")
        self.gd_deconv.forward_unit = self.conv
        self.gd_deconv.gate_skip = self.decision.gd_skip
        self.gd_deconv.need_err_input = False
        self.gd_deconv.gate_block = self.decision.complete
        del self.gds[:]
        self.gds.append(self.gd_deconv)
        self.repeater.link_from(self.gd_deconv)
        self.loader.gate_block = self.decision.complete

        prev = self.snapshotter
        self.plt = []
        for i, style in ((1, "b-"), (2, "k-")):
            p = plotting_units.AccumulatingPlotter(self, name="mse %d" % i, plot_style=style)
            p.input = self.decision.epoch_metrics
            p.input_field = i
            p.input_offset = 0
            p.link_from(prev)
            p.gate_skip = ~self.decision.epoch_ended
            p.gate_block = self.decision.complete
            self.plt.append(p)
            prev = p
        side = 28
        for name, unit, attr, shape in (
                ("Weights", self.conv, "weights", [cfg.kx, cfg.ky, 1]),
                ("First Layer Input", self.conv, "input", [side, side, 1]),
                ("First Layer Output", self.conv, "output",
                 [side - cfg.kx + 1, side - cfg.ky + 1, cfg.n_kernels]),
                ("Deconv result", self.deconv, "output", [side, side, 1])):
            w = nn_plotting_units.Weights2D(self, name=name, limit=cfg.weights_plotter.limit)
            w.link_attrs(unit, ("input", attr))
            w.get_shape_from = shape
            w.link_from(prev)
            w.gate_skip = ~self.decision.epoch_ended
            w.gate_block = self.decision.complete
            self.plt.append(w)
            prev = w
        self.gd_deconv.link_from(prev)


def build(launcher=None, **kwargs):
    from ..core.workflow import DummyLauncher
    return MnistAEWorkflow(launcher or DummyLauncher(), **kwargs)


def run(load, main):
    load(MnistAEWorkflow)
    main()
