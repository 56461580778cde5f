"""Two small image-classification research workflows that differ only in their configs:

* **Hands** (/root/reference/tests/research/Hands/hands.py, hands_config.py:44-70): raw ``.raw``
  grey images of hands / not-hands, FC-tanh(30) → softmax(2), linear normalisation.
* **TvChannels** (/root/reference/tests/research/TvChannels/channels.py, channels_config.py:43-90):
  TV-channel logo recognition on 224x224 HSV frames with a Sobel channel, aspect-preserving
  scaling on a transparent background, FC-tanh → softmax.

Both are ``StandardWorkflow`` loops: downloader → loader → forwards → evaluator → decision →
snapshotter (+ image saver) → GDs.
"""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..loader.base import LoaderError
from ..loader.image import FullBatchAutoLabelFileImageLoader
from ..workflow.standard_workflow import StandardWorkflow

_ds = str(root.common.dirs.datasets)
_cache = str(root.common.dirs.cache)

root.hands.update({
    "decision": {"fail_iterations": 100, "max_epochs": 10000},
    "loss_function": "softmax",
    "downloader": {"url": None, "directory": root.common.dirs.datasets, "files": ["hands"]},
    "image_saver": {"do": True, "out_dirs": [os.path.join(_cache, "tmp", d)
                                             for d in ("test", "validation", "train")]},
    "loader_name": "hands_loader",
    "snapshotter": {"prefix": "hands", "interval": 1, "time_interval": 0},
    "loader": {"minibatch_size": 40, "train_paths": [os.path.join(_ds, "hands", "Training")],
               "force_numpy": False, "color_space": "GRAY", "background_color": (0,),
               "normalization_type": "linear", "raw_shape": (32, 32),
               "validation_paths": [os.path.join(_ds, "hands", "Testing")]},
    "layers": [{"name": "fc_tanh1", "type": "all2all_tanh",
                "->": {"output_sample_shape": 30},
                "<-": {"learning_rate": 0.008, "weights_decay": 0.0}},
               {"name": "fc_softmax2", "type": "softmax",
                "<-": {"learning_rate": 0.008, "weights_decay": 0.0}}]})

root.channels.update({
    "decision": {"fail_iterations": 50, "max_epochs": numpy.iinfo(numpy.uint32).max},
    "downloader": {"url": None, "directory": root.common.dirs.datasets,
                   "files": ["channels_train"]},
    "snapshotter": {"prefix": "channels", "interval": 1, "time_interval": 0},
    "image_saver": {"out_dirs": [os.path.join(_cache, "tmp", d)
                                 for d in ("test", "validation", "train")]},
    "loss_function": "softmax",
    "loader_name": "full_batch_auto_label_file_image",
    "loader": {"minibatch_size": 30, "force_numpy": False, "validation_ratio": 0.15,
               "shuffle_limit": numpy.iinfo(numpy.uint32).max, "normalization_type": "mean_disp",
               "add_sobel": True, "file_subtypes": ["png", "jpeg"], "mirror": False,
               "color_space": "HSV", "scale": (224, 224), "background_color": (0, 0, 0),
               "scale_maintain_aspect_ratio": True,
               "train_paths": [os.path.join(_ds, "channels_train")]},
    "layers": [{"name": "fc_tanh1", "type": "all2all_tanh",
                "->": {"output_sample_shape": 54},
                "<-": {"learning_rate": 0.001, "weights_decay": 0.00005}},
               {"name": "fc_softmax2", "type": "softmax",
                "<-": {"learning_rate": 0.001, "weights_decay": 0.00005}}]})


class HandsLoader(FullBatchAutoLabelFileImageLoader):
    """The Hands dataset stores headerless 8-bit ``.raw`` files next to ordinary images."""
    MAPPING = "hands_loader"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.raw_shape = tuple(kwargs.get("raw_shape", (32, 32)))

    def is_valid_filename(self, path):
        if path.lower().endswith(".raw"):
            name = os.path.basename(path)
            return not any(p.match(name) for p in self.ignored_files)
        return super().is_valid_filename(path)

    def decode(self, key):
        if not key.lower().endswith(".raw"):
            return super().decode(key)
        raw = numpy.fromfile(key, dtype=numpy.uint8)
        h, w = self.raw_shape
        if raw.size != h * w:
            raise LoaderError("%s holds %d bytes, expected %dx%d" % (key, raw.size, h, w))
        return raw.reshape(h, w, 1)


class ImageClassifierWorkflow(StandardWorkflow):
    def create_workflow(self):
        self.link_downloader(self.start_point)
        self.link_repeater(self.downloader)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [self.link_snapshotter(self.decision)]
        if self.config.image_saver.get("out_dirs"):
            end_units.append(self.link_image_saver(self.decision))
        end_units.append(self.link_error_plotter(self.decision))
        self.link_loop(self.link_gds(*end_units))
        self.link_end_point(self.gds[0])


HandsWorkflow = ImageClassifierWorkflow
ChannelsWorkflow = ImageClassifierWorkflow


def _kwargs(cfg):
    return dict(decision_config=cfg.decision, snapshotter_config=cfg.snapshotter,
                loader_name=cfg.loader_name, loader_config=cfg.loader, layers=cfg.layers,
                downloader_config=cfg.downloader, loss_function=cfg.loss_function,
                image_saver_config={k: v for k, v in cfg.image_saver.to_dict().items()
                                    if k != "do"})


def build_hands(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = _kwargs(root.hands)
    kw.update(overrides)
    return ImageClassifierWorkflow(launcher or DummyLauncher(), **kw)


def build_channels(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = _kwargs(root.channels)
    kw.update(overrides)
    return ImageClassifierWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    cfg = root.channels if root.common.get("sample") == "channels" else root.hands
    load(ImageClassifierWorkflow, **_kwargs(cfg))
    main()
