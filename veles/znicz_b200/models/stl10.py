"""STL-10 sample: the CIFAR caffe-style conv net on STL-10 scaled to 32x32.
Parity: /root/reference/tests/research/Stl10/stl10.py, stl10_config.py:44-130."""
from __future__ import annotations

import os

from ..core.config import root
from ..workflow.standard_workflow import StandardWorkflow
from . import cifar

root.stl.update({
    "loader_name": "full_batch_stl_10",
    "loss_function": "softmax",
    "downloader": {"url": None, "directory": root.common.dirs.datasets,
                   "files": ["stl10_binary"]},
    "snapshotter": {"prefix": "stl10", "interval": 1, "time_interval": 0},
    "decision": {"fail_iterations": 200, "max_epochs": 10000},
    "loader": {"directory": os.path.join(str(root.common.dirs.datasets), "stl10_binary"),
               "minibatch_size": 50, "scale": (32, 32),
               "normalization_type": "internal_mean"},
    "weights_plotter": {"limit": 256, "split_channels": False},
    "layers": cifar.caffe_layers()})


class Stl10Workflow(StandardWorkflow):
    def create_workflow(self):
        self.link_downloader(self.start_point)
        self.link_repeater(self.downloader)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [link(self.decision) for link in (
            self.link_snapshotter, self.link_error_plotter, self.link_conf_matrix_plotter)]
        end_units.append(self.link_weights_plotter("weights", self.decision))
        self.link_loop(self.link_gds(*end_units))
        self.link_end_point(self.gds[0])


def kwargs_from_config():
    c = root.stl
    return dict(decision_config=c.decision, snapshotter_config=c.snapshotter,
                loader_name=c.loader_name, loader_config=c.loader, layers=c.layers,
                downloader_config=c.downloader, loss_function=c.loss_function,
                weights_plotter_config=c.weights_plotter)


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = kwargs_from_config()
    kw.update(overrides)
    return Stl10Workflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    load(Stl10Workflow, **kwargs_from_config())
    main()
