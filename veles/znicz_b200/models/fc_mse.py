"""Shared hand-wired fully connected MSE workflow used by the research models
(Mnist7 /root/reference/tests/research/Mnist7/mnist7.py:94-210, Approximator
/root/reference/tests/research/Approximator/approximator.py:181-300, VideoAE
/root/reference/tests/research/VideoAE/video_ae.py): FC-tanh stack → EvaluatorMSE →
DecisionMSE → snapshotter → GDTanh chain, optional plotters."""
from __future__ import annotations

from ..core.config import root
from ..ops import all2all, gd
from ..ops.nn_units import NNWorkflow, NNSnapshotterToFile
from ..utils import nn_plotting_units, plotting_units
from ..workflow import decision, evaluator


class FullyConnectedMSEWorkflow(NNWorkflow):
    """``config`` is the sample's config node (``root.mnist7`` …) providing ``layers``,
    ``decision``, ``snapshotter``, ``learning_rate``, ``weights_decay``; ``loader`` is a
    factory ``(workflow) -> loader`` whose loader serves ``minibatch_targets``."""

    def __init__(self, workflow, config, loader_factory, **kwargs):
        super().__init__(workflow, **kwargs)
        self.config_ = config
        layers = kwargs.get("layers") or config.layers
        self.repeater.link_from(self.start_point)
        self.loader = loader_factory(self)
        self.loader.link_from(self.repeater)
        del self.forwards[:]
        for i, layer in enumerate(layers):
            aa = all2all.All2AllTanh(
                self, output_sample_shape=layer if isinstance(layer, (tuple, list))
                else (layer,), **kwargs.get("forward_kwargs", {}))
            self.forwards.append(aa)
            src = self.forwards[-2] if i else self.loader
            aa.link_from(src)
            aa.link_attrs(src, ("input", "output" if i else "minibatch_data"))

        self.evaluator = evaluator.EvaluatorMSE(self)
        self.evaluator.link_from(self.forwards[-1])
        self.evaluator.link_attrs(self.forwards[-1], "output")
        self.evaluator.link_attrs(
            self.loader, ("batch_size", "minibatch_size"), ("target", "minibatch_targets"),
            ("normalizer", "target_normalizer"))
        if kwargs.get("use_class_targets", False):
            self.evaluator.link_attrs(self.loader, ("labels", "minibatch_labels"),
                                      "class_targets")

        self.decision = decision.DecisionMSE(
            self, fail_iterations=config.decision.fail_iterations,
            max_epochs=config.decision.max_epochs)
        self.decision.link_from(self.evaluator)
        self.decision.link_attrs(self.loader, "minibatch_class", "minibatch_size",
                                 "last_minibatch", "class_lengths", "epoch_ended",
                                 "epoch_number")
        self.decision.link_attrs(self.evaluator, ("minibatch_metrics", "metrics"),
                                 ("minibatch_mse", "mse"))
        if kwargs.get("use_class_targets", False):
            self.decision.link_attrs(self.evaluator, ("minibatch_n_err", "n_err"))

        snap = config.snapshotter
        self.snapshotter = NNSnapshotterToFile(
            self, prefix=snap.prefix, directory=root.common.dirs.snapshots,
            compression=snap.get("compression", ""), interval=snap.get("interval", 1),
            time_interval=snap.get("time_interval", 0))
        self.snapshotter.link_from(self.decision)
        self.snapshotter.link_attrs(self.decision, ("suffix", "snapshot_suffix"))
        self.snapshotter.gate_skip = ~self.loader.epoch_ended
        self.snapshotter.skip = ~self.decision.improved
        self.end_point.link_from(self.snapshotter)
        self.end_point.gate_block = ~self.decision.complete

        last = self.snapshotter
        self.plotters = []
        if kwargs.get("add_plotters", False):
            for i, style in ((1, "b-"), (2, "k-")):
                p = plotting_units.AccumulatingPlotter(self, name="mse %d" % i,
                                                       plot_style=style)
                p.input = self.decision.epoch_metrics
                p.input_field = i
                p.input_offset = 0
                p.link_from(last)
                p.gate_skip = ~self.decision.epoch_ended
                self.plotters.append(p)
                last = p
            wp = config.get("weights_plotter", None)
            w = nn_plotting_units.Weights2D(
                self, name="First Layer Weights",
                limit=wp.get("limit", 16) if wp is not None else 16)
            w.link_attrs(self.forwards[0], ("input", "weights"))
            w.link_from(last)
            w.gate_skip = ~self.decision.epoch_ended
            self.plotters.append(w)
            hist = nn_plotting_units.MSEHistogram(self, name="MSE Histogram")
            hist.link_attrs(self.evaluator, "mse")
            hist.link_from(w)
            hist.gate_skip = ~self.decision.epoch_ended
            self.plotters.append(hist)
            last = hist

        self.gds[:] = (None,) * len(self.forwards)
        for i in range(len(self.forwards) - 1, -1, -1):
            g = gd.GDTanh(self, **kwargs.get("gd_kwargs", {}))
            if i == len(self.forwards) - 1:
                g.link_from(last)
                g.link_attrs(self.evaluator, "err_output")
            else:
                g.link_from(self.gds[i + 1])
                g.link_attrs(self.gds[i + 1], ("err_output", "err_input"))
            g.link_attrs(self.forwards[i], "output", "input", "weights", "bias")
            g.link_attrs(self.loader, ("batch_size", "minibatch_size"))
            g.gate_skip = self.decision.gd_skip
            g.forward_unit = self.forwards[i]
            self.gds[i] = g
        self.gds[-1].gate_block = self.decision.complete
        self.gds[0].need_err_input = False
        self.repeater.link_from(self.gds[0])
        self.loader.gate_block = self.decision.complete

    def initialize(self, learning_rate=None, weights_decay=None, device=None, **kwargs):
        cfg = self.config_
        return super().initialize(
            learning_rate=cfg.learning_rate if learning_rate is None else learning_rate,
            weights_decay=cfg.weights_decay if weights_decay is None else weights_decay,
            device=device, **kwargs)
