"""Lines sample: geometric figure recognition; topology given as an MCDNNIC string.

Parity: /root/reference/samples/Lines/lines.py:47-86, lines_config.py:45-71
("12x256x256-32C4-MP2-64C4-MP3-32N-4N", softmax, ``mean_disp`` normalisation, batch 12,
table / weights / multi-histogram plotters inside the loop) and draw_lines.py (generator,
re-implemented in ``generate_dataset``: horizontal / vertical / two diagonal line classes).
"""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..workflow.standard_workflow import StandardWorkflow

_train = os.path.join(str(root.common.dirs.datasets), "lines_min", "learn")
_valid = os.path.join(str(root.common.dirs.datasets), "lines_min", "test")

root.lines.mcdnnic_parameters = {"<-": {"learning_rate": 0.01}}
root.lines.update({
    "loader_name": "full_batch_auto_label_file_image",
    "loss_function": "softmax",
    "downloader": {"url": None, "directory": root.common.dirs.datasets,
                   "files": ["lines_min"]},
    "mcdnnic_topology": "12x256x256-32C4-MP2-64C4-MP3-32N-4N",
    "decision": {"fail_iterations": 100, "max_epochs": numpy.iinfo(numpy.uint32).max},
    "snapshotter": {"prefix": "lines", "interval": 1, "time_interval": 0},
    "image_saver": {"out_dirs": [
        os.path.join(str(root.common.dirs.cache), "tmp", d)
        for d in ("test", "validation", "train")]},
    "loader": {"minibatch_size": 12, "force_numpy": False, "color_space": "RGB",
               "file_subtypes": ["jpeg", "png"], "normalization_type": "mean_disp",
               "train_paths": [_train], "validation_paths": [_valid]},
    "weights_plotter": {"limit": 32, "split_channels": False}})


def generate_dataset(directory, size=256, per_class=(30, 10), seed=11):
    """``learn/<class>/*.png`` and ``test/<class>/*.png`` with 4 line orientations."""
    import cv2
    rs = numpy.random.RandomState(seed)
    classes = ("horizontal", "vertical", "diagonal_up", "diagonal_down")
    for split, n in zip(("learn", "test"), per_class):
        for ci, cname in enumerate(classes):
            d = os.path.join(directory, split, cname)
            os.makedirs(d, exist_ok=True)
            for i in range(n):
                img = numpy.full((size, size, 3), 255, numpy.uint8)
                for _ in range(rs.randint(1, 4)):
                    c = rs.randint(size // 8, size - size // 8, 2)
                    ln = rs.randint(size // 4, size // 2)
                    dx, dy = ((ln, 0), (0, ln), (ln, -ln), (ln, ln))[ci]
                    col = tuple(int(v) for v in rs.randint(0, 120, 3))
                    cv2.line(img, (int(c[0] - dx // 2), int(c[1] - dy // 2)),
                             (int(c[0] + dx // 2), int(c[1] + dy // 2)), col,
                             int(rs.randint(2, 6)))
                cv2.imwrite(os.path.join(d, "%03d.png" % i), img)
    return directory


class LinesWorkflow(StandardWorkflow):
    def create_workflow(self):
        self.link_downloader(self.start_point)
        self.link_repeater(self.downloader)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [link(self.decision) for link in (self.link_snapshotter,
                                                      self.link_error_plotter)]
        self.link_image_saver(*end_units)
        gd = self.link_gds(self.image_saver)
        self.link_table_plotter(gd).gate_block = self.decision.complete
        last_weights = self.link_weights_plotter("gradient_weights", self.table_plotter)
        self.link_multi_hist_plotter("gradient_weights", last_weights)
        self.repeater.link_from(self.multi_hist_plotter[-1])
        self.link_end_point(gd)


def kwargs_from_config():
    return dict(
        decision_config=root.lines.decision, snapshotter_config=root.lines.snapshotter,
        image_saver_config=root.lines.image_saver, loader_config=root.lines.loader,
        loader_name=root.lines.loader_name, loss_function=root.lines.loss_function,
        downloader_config=root.lines.downloader,
        weights_plotter_config=root.lines.weights_plotter,
        mcdnnic_topology=root.lines.mcdnnic_topology,
        mcdnnic_parameters=root.lines.mcdnnic_parameters)


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = kwargs_from_config()
    kw.update(overrides)
    return LinesWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    load(LinesWorkflow, **kwargs_from_config())
    main()
