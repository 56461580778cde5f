"""Kanji sample: FC-tanh 250-250-(24x24) regression of glyph images (MSE loss).

Parity: /root/reference/samples/Kanji/kanji.py:46-111, kanji_config.py:45-98 (loader
``full_batch_auto_label_file_image_mse``, ``range_linear`` target normalisation, batch 50,
weight injection in ``initialize``) and generate_kanji.py (dataset generator; here glyphs
are rendered with OpenCV's Hershey fonts + random affine jitter because freetype/CJK fonts
are not in the image — ``generate_dataset`` produces the same directory layout:
``<dir>/train/<label>/*.png`` and ``<dir>/target/<label>.png``).
"""
from __future__ import annotations

import logging
import os

import numpy

from ..core.config import root
from ..workflow.standard_workflow import StandardWorkflow

_base = os.path.join(str(root.common.dirs.datasets), "kanji")

root.kanji.update({
    "decision": {"fail_iterations": 1000, "max_epochs": 10000},
    "downloader": {"url": None, "directory": root.common.dirs.datasets, "files": ["kanji"]},
    "loss_function": "mse",
    "loader_name": "full_batch_auto_label_file_image_mse",
    "add_plotters": True,
    "image_saver": {"out_dirs": [
        os.path.join(str(root.common.dirs.cache), "tmp", d)
        for d in ("test", "validation", "train")]},
    "loader": {"minibatch_size": 50, "force_numpy": False, "file_subtypes": ["png"],
               "train_paths": [os.path.join(_base, "train")],
               "target_paths": [os.path.join(_base, "target")],
               "color_space": "GRAY", "normalization_type": "linear",
               "target_normalization_type": "range_linear",
               "target_normalization_parameters": {},
               "targets_shape": (24, 24), "background_color": (0,),
               "validation_ratio": 0.15},
    "snapshotter": {"prefix": "kanji"},
    "weights_plotter": {"limit": 16},
    "layers": [
        {"name": "fc_tanh1", "type": "all2all_tanh",
         "->": {"output_sample_shape": 250, "weights_filling": "uniform",
                "weights_stddev": 0.03125, "bias_filling": "uniform",
                "bias_stddev": 0.03125},
         "<-": {"learning_rate": 0.0001, "weights_decay": 0.00005}},
        {"name": "fc_tanh2", "type": "all2all_tanh",
         "->": {"output_sample_shape": 250, "weights_filling": "uniform",
                "weights_stddev": 0.036858530918682665, "bias_filling": "uniform",
                "bias_stddev": 0.036858530918682665},
         "<-": {"learning_rate": 0.0001, "weights_decay": 0.00005}},
        {"name": "fc_tanh3", "type": "all2all_tanh",
         "->": {"output_sample_shape": (24, 24), "weights_filling": "uniform",
                "weights_stddev": 0.036858530918682665, "bias_filling": "uniform",
                "bias_stddev": 0.036858530918682665},
         "<-": {"learning_rate": 0.0001, "weights_decay": 0.00005}}]})


def generate_dataset(directory, glyphs="ABCDEFGHJK", per_glyph=20, size=32, target=24,
                     seed=7):
    """Write ``train/<glyph>/NNN.png`` (jittered renders) and ``target/<glyph>.png``."""
    import cv2
    rs = numpy.random.RandomState(seed)

    def render(ch, sz, jitter):
        img = numpy.zeros((sz, sz), numpy.uint8)
        scale = sz / 32.0
        (tw, th), _ = cv2.getTextSize(ch, cv2.FONT_HERSHEY_SIMPLEX, scale, 2)
        org = ((sz - tw) // 2, (sz + th) // 2)
        cv2.putText(img, ch, org, cv2.FONT_HERSHEY_SIMPLEX, scale, 255, 2, cv2.LINE_AA)
        if jitter:
            ang = rs.uniform(-12, 12)
            m = cv2.getRotationMatrix2D((sz / 2, sz / 2), ang, rs.uniform(0.85, 1.1))
            m[:, 2] += rs.uniform(-2, 2, 2)
            img = cv2.warpAffine(img, m, (sz, sz))
            noise = rs.randn(sz, sz) * 8
            img = numpy.clip(img.astype(numpy.float32) + noise, 0, 255).astype(numpy.uint8)
        return img
    os.makedirs(os.path.join(directory, "target"), exist_ok=True)
    for ch in glyphs:
        d = os.path.join(directory, "train", ch)
        os.makedirs(d, exist_ok=True)
        cv2.imwrite(os.path.join(directory, "target", ch + ".png"), render(ch, target, False))
        for i in range(per_glyph):
            cv2.imwrite(os.path.join(d, "%03d.png" % i), render(ch, size, True))
    return directory


class KanjiWorkflow(StandardWorkflow):
    """Fully connected network with MSE loss reproducing the target glyph of the class."""

    def create_workflow(self):
        self.link_downloader(self.start_point)
        self.link_repeater(self.downloader)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [link(self.decision) for link in (self.link_snapshotter,
                                                      self.link_image_saver)]
        if root.kanji.add_plotters:
            end_units.extend((
                self.link_error_plotter(self.decision),
                self.link_min_max_plotter(False, self.decision),
                self.link_min_max_plotter(True, self.max_plotter[-1]),
                self.link_mse_plotter(self.decision)))
        last_gd = self.link_gds(*end_units)
        self.link_loop(last_gd)
        self.link_end_point(last_gd)

    def initialize(self, device=None, weights=None, bias=None, **kwargs):
        res = super().initialize(device=device, **kwargs)
        for arrs, name in ((weights, "weights"), (bias, "bias")):
            if arrs is None:
                continue
            for i, fwd in enumerate(self.forwards):
                a = getattr(fwd, name)
                a.map_invalidate()
                a.mem[:] = arrs[i][:]
                a.unmap()
            for fwd in self.forwards:
                if getattr(fwd, "on_cuda", False):
                    fwd.refresh_shadows()
        return res


def kwargs_from_config():
    return dict(
        decision_config=root.kanji.decision, loader_config=root.kanji.loader,
        loader_name=root.kanji.loader_name, snapshotter_config=root.kanji.snapshotter,
        layers=root.kanji.layers, downloader_config=root.kanji.downloader,
        weights_plotter_config=root.kanji.weights_plotter,
        image_saver_config=root.kanji.image_saver, loss_function=root.kanji.loss_function)


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = kwargs_from_config()
    kw.update(overrides)
    return KanjiWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    weights = bias = None
    w, snapshot = load(KanjiWorkflow, **kwargs_from_config())
    if snapshot:
        if isinstance(w, tuple):
            logging.info("Will load weights")
            weights, bias = w[0], w[1]
        else:
            logging.info("Will load workflow")
            w.decision.improved <<= True
    main(weights=weights, bias=bias)
