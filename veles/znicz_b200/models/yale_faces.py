"""YaleFaces sample: FC-tanh(100) → softmax face recognition, plus the preprocessing
variant that only records minibatches.

Parity: /root/reference/samples/YaleFaces/yale_faces.py:43-76, yale_faces_config.py:43-92
(batch 40, ``mean_disp``, validation_ratio 0.15, ``.*Ambient.*`` ignored, Publisher at the
end) and yale_faces_preprocessing.py:40-65 (loader → ``MinibatchesSaver`` loop that stops
when the train set has been served once).
"""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..workflow.standard_workflow import StandardWorkflow

root.yalefaces.update({
    "downloader": {"url": None, "directory": root.common.dirs.datasets,
                   "files": ["CroppedYale"]},
    "name_workflow": "FullyConnected_YaleFaces",
    "decision": {"fail_iterations": 50, "max_epochs": 1000},
    "loss_function": "softmax",
    "loader_name": "full_batch_auto_label_file_image",
    "snapshotter": {"prefix": "yalefaces", "interval": 1, "time_interval": 0},
    "publisher": {"backends": {"json": {}, "markdown": {}}},
    "preprocessing": False,
    "datasaver": {"file_name": os.path.join(str(root.common.dirs.datasets),
                                            "yale_faces_minibatches.dat")},
    "loader": {"minibatch_size": 40, "force_numpy": False, "validation_ratio": 0.15,
               "file_subtypes": ["x-portable-graymap"], "ignored_files": [".*Ambient.*"],
               "shuffle_limit": numpy.iinfo(numpy.uint32).max, "add_sobel": False,
               "mirror": False, "color_space": "GRAY", "background_color": (0,),
               "normalization_type": "mean_disp",
               "train_paths": [os.path.join(str(root.common.dirs.datasets), "CroppedYale")]},
    "layers": [
        {"name": "fc_tanh1", "type": "all2all_tanh",
         "->": {"output_sample_shape": 100, "weights_filling": "uniform",
                "weights_stddev": 0.05, "bias_filling": "uniform", "bias_stddev": 0.05},
         "<-": {"learning_rate": 0.01, "weights_decay": 0.00005}},
        {"name": "fc_softmax2", "type": "softmax",
         "->": {"output_sample_shape": 39, "weights_filling": "uniform",
                "weights_stddev": 0.05, "bias_filling": "uniform", "bias_stddev": 0.05},
         "<-": {"learning_rate": 0.01, "weights_decay": 0.00005}}]})


def generate_dataset(directory, subjects=6, per_subject=12, size=(42, 48), seed=13):
    """Synthetic ``CroppedYale``-shaped tree: ``yaleBNN/*.pgm`` with per-subject "faces"
    (smooth random fields) under varying illumination, plus one ``*Ambient*`` file each."""
    import cv2
    rs = numpy.random.RandomState(seed)
    w, h = size
    yy, xx = numpy.mgrid[0:h, 0:w].astype(numpy.float32)
    for s in range(subjects):
        d = os.path.join(directory, "yaleB%02d" % (s + 1))
        os.makedirs(d, exist_ok=True)
        face = cv2.GaussianBlur(rs.rand(h, w).astype(numpy.float32), (0, 0), 4)
        face = (face - face.min()) / (face.max() - face.min())
        for i in range(per_subject):
            ang = rs.uniform(0, 2 * numpy.pi)
            light = 0.6 + 0.4 * (numpy.cos(ang) * (xx / w - 0.5) + numpy.sin(ang) * (yy / h - 0.5))
            img = numpy.clip(face * light * 255 + rs.randn(h, w) * 4, 0, 255)
            cv2.imwrite(os.path.join(d, "yaleB%02d_P00A%+04dE%+03d.pgm" % (
                s + 1, int(numpy.degrees(ang)) - 180, i)), img.astype(numpy.uint8))
        cv2.imwrite(os.path.join(d, "yaleB%02d_P00_Ambient.pgm" % (s + 1)),
                    numpy.zeros((h, w), numpy.uint8))
    return directory


class YaleFacesWorkflow(StandardWorkflow):
    def create_workflow(self):
        self.link_downloader(self.start_point)
        self.link_repeater(self.downloader)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [link(self.decision) for link in (
            self.link_snapshotter, self.link_error_plotter, self.link_err_y_plotter)]
        self.link_loop(self.link_gds(*end_units))
        self.link_publisher(self.gds[0])
        self.link_end_point(self.publisher)


class YaleFacesPreprocessingWorkflow(StandardWorkflow):
    """loader → MinibatchesSaver → repeat until every class has been recorded once."""

    def link_end_point(self, *parents):
        self.end_point.link_from(*parents).gate_block = ~self.loader.train_ended
        self.loader.gate_block = self.loader.train_ended
        return self.end_point

    def create_workflow(self):
        self.link_repeater(self.start_point)
        self.link_loader(self.repeater)
        self.link_data_saver(self.loader)
        self.repeater.link_from(self.data_saver)
        self.link_end_point(self.data_saver)


def kwargs_from_config():
    return dict(
        decision_config=root.yalefaces.decision, snapshotter_config=root.yalefaces.snapshotter,
        loader_config=root.yalefaces.loader, layers=root.yalefaces.layers,
        loss_function=root.yalefaces.loss_function, loader_name=root.yalefaces.loader_name,
        name=root.yalefaces.name_workflow, publisher_config=root.yalefaces.publisher,
        downloader_config=root.yalefaces.downloader)


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = kwargs_from_config()
    kw.update(overrides)
    return YaleFacesWorkflow(launcher or DummyLauncher(), **kw)


def build_preprocessing(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = dict(loader_name=root.yalefaces.loader_name,
              data_saver_config=root.yalefaces.datasaver, preprocessing=True,
              loader_config=root.yalefaces.loader)
    kw.update(overrides)
    return YaleFacesPreprocessingWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    if root.yalefaces.get("preprocessing", False):
        load(YaleFacesPreprocessingWorkflow, loader_name=root.yalefaces.loader_name,
             data_saver_config=root.yalefaces.datasaver, preprocessing=True,
             loader_config=root.yalefaces.loader)
    else:
        load(YaleFacesWorkflow, **kwargs_from_config())
    main()
