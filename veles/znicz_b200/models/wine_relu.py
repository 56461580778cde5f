"""WineRelu: the Wine task as a StandardWorkflow with a softplus ("relu") hidden layer.
Parity: /root/reference/tests/research/WineRelu/wine_relu.py, wine_relu_config.py:44-63."""
from __future__ import annotations

from ..core.config import root
from ..workflow.standard_workflow import StandardWorkflow
from . import wine  # noqa: F401  (registers wine_loader)

root.wine_relu.update({
    "decision": {"fail_iterations": 250, "max_epochs": 100000},
    "downloader": {"url": None, "directory": root.common.dirs.datasets, "files": []},
    "snapshotter": {"prefix": "wine_relu", "interval": 1, "time_interval": 0},
    "loader_name": "wine_loader",
    "loader": {"minibatch_size": 10, "force_numpy": False, "dataset_file": None},
    "layers": [{"name": "fc_relu1", "type": "all2all_relu",
                "->": {"output_sample_shape": 10},
                "<-": {"learning_rate": 0.03, "weights_decay": 0.0}},
               {"name": "fc_softmax2", "type": "softmax",
                "<-": {"learning_rate": 0.03, "weights_decay": 0.0}}]})


class WineReluWorkflow(StandardWorkflow):
    def create_workflow(self):
        self.link_downloader(self.start_point)
        self.link_repeater(self.downloader)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        self.link_snapshotter(self.decision)
        self.link_loop(self.link_gds(self.snapshotter))
        self.link_end_point(self.gds[0])


def kwargs_from_config():
    c = root.wine_relu
    return dict(decision_config=c.decision, snapshotter_config=c.snapshotter,
                loader_name=c.loader_name, loader_config=c.loader, layers=c.layers,
                downloader_config=c.downloader, loss_function="softmax")


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = kwargs_from_config()
    kw.update(overrides)
    return WineReluWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    load(WineReluWorkflow, **kwargs_from_config())
    main()
