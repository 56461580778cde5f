"""VideoAE: fully connected autoencoder on video frames (input == target, MSE).
Parity: /root/reference/tests/research/VideoAE/video_ae.py (frames from image files,
``layers`` [9, full-frame], lr 0.000004/weights_decay 0.00005, GRAY, scale)."""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..loader.base import UserLoaderRegistry
from ..loader.fullbatch import FullBatchLoaderMSE
from ..loader.image import FullBatchFileImageLoader, ImageOptionsMixin
from .fc_mse import FullyConnectedMSEWorkflow

root.video_ae.update({
    "decision": {"fail_iterations": 100, "max_epochs": 100000},
    "snapshotter": {"prefix": "video_ae"},
    "loader_name": "video_ae_loader",
    "loader": {"minibatch_size": 50, "force_numpy": False,
               "train_paths": [os.path.join(str(root.common.dirs.datasets), "video_ae", "img")],
               "color_space": "GRAY", "background_color": (0x80,), "normalization_type": "linear",
               "target_normalization_type": "range_linear", "validation_ratio": 0.1,
               "label_regexp": r"^(\\D*)", "file_subtypes": ["png", "jpeg"]},
    "weights_plotter": {"limit": 16},
    "learning_rate": 0.000004,
    "weights_decay": 0.00005,
    "layers": [9, [90, 160]]})


class VideoAELoader(FullBatchLoaderMSE, ImageOptionsMixin):
    """Frames decoded once; the target of every frame is the frame itself."""
    MAPPING = "video_ae_loader"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self._init_image_options(kwargs)
        self.class_keys = [[], [], []]

    _labelled = lambda self: False          # noqa: E731
    load_images = FullBatchFileImageLoader.load_data

    def load_data(self):
        self.load_images()
        self.original_labels = []
        frames = self.original_data.mem
        if frames.shape[-1] == 1:
            frames = frames[..., 0]
            self.original_data.reset(frames)
        self.original_targets.reset(frames.astype(numpy.float32).copy())


class VideoAEWorkflow(FullyConnectedMSEWorkflow):
    def __init__(self, workflow, **kwargs):
        cfg = dict(root.video_ae.loader.to_dict())
        cfg.update(kwargs.pop("loader_config", {}))
        factory = UserLoaderRegistry.get_factory(
            kwargs.pop("loader_name", root.video_ae.loader_name), **cfg)
        super().__init__(workflow, root.video_ae, factory, **kwargs)


def build(launcher=None, **kwargs):
    from ..core.workflow import DummyLauncher
    return VideoAEWorkflow(launcher or DummyLauncher(), **kwargs)


def run(load, main):
    load(VideoAEWorkflow, layers=root.video_ae.layers)
    main(learning_rate=root.video_ae.learning_rate, weights_decay=root.video_ae.weights_decay)
