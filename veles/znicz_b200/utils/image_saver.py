"""ImageSaver: dumps misclassified samples (softmax) or input/output/target triples
(MSE) as PNG files. Parity: /root/reference/image_saver.py:50-273."""
from __future__ import annotations

import glob
import os

import numpy

from ..core.config import root
from ..core.memory import Array
from ..core.units import Unit


def _host(v):
    if isinstance(v, Array):
        v.map_read()
        return v.mem
    return v


def normalize_image(image, colorspace=None):
    """Stretch to [0, 255] uint8; converts ``colorspace`` → RGB when it is not RGB."""
    img = numpy.array(image, dtype=numpy.float32)
    img -= img.min()
    mx = img.max()
    if mx:
        img *= 255.0 / mx
    else:
        img[...] = 127.5
    img = img.astype(numpy.uint8)
    if colorspace not in (None, "RGB") and img.ndim == 3 and img.shape[2] == 3:
        import cv2
        img = cv2.cvtColor(img, getattr(cv2, "COLOR_" + colorspace + "2RGB"))
    return img


class ImageSaver(Unit):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        cache = str(root.common.dirs.get("cache", "."))
        self.out_dirs = kwargs.get("out_dirs", [
            os.path.join(cache, "tmpimg", d) for d in ("test", "validation", "train")])
        self.limit = kwargs.get("limit", 100)
        self.output = None
        self.target = None
        self.max_idx = None
        self._last_save_time = 0
        self.save_time = 0
        self._n_saved = [0, 0, 0]
        self.color_space = kwargs.get("color_space", "RGB")
        self.reversed_labels_mapping = None
        self.demand("input", "indices", "labels", "minibatch_class", "minibatch_size")

    @staticmethod
    def as_image(inp):
        if inp.ndim == 1:
            return None
        if inp.ndim == 2:
            return None if 1 in inp.shape else inp
        if inp.ndim == 3:
            if inp.shape[2] == 3:
                return inp
            if inp.shape[0] == 3:
                return inp.transpose(1, 2, 0)
            if inp.shape[2] == 4:
                return inp[:, :, :3]
            if inp.shape[2] == 1:
                return inp[:, :, 0]
            return None
        raise ValueError("unsupported sample shape %s" % (inp.shape,))

    def initialize(self, **kwargs):
        pass

    def _purge_if_new_snapshot(self):
        if self._last_save_time >= self.save_time:
            return
        self._last_save_time = self.save_time
        self._n_saved = [0, 0, 0]
        for d in self.out_dirs:
            for f in glob.glob(os.path.join(d, "**", "*.png"), recursive=True):
                try:
                    os.unlink(f)
                except OSError:
                    pass

    def _save(self, image, path):
        from PIL import Image
        try:
            Image.fromarray(image).save(path)
        except (ValueError, OSError, TypeError) as e:
            self.warning("Could not save image to %s: %s", path, e)

    def run(self):
        inp, idx, labels = _host(self.input), _host(self.indices), _host(self.labels)
        out, tgt, mx = _host(self.output), _host(self.target), _host(self.max_idx)
        cls = int(self.minibatch_class)
        for d in self.out_dirs:
            os.makedirs(d, exist_ok=True)
        self._purge_if_new_snapshot()
        rlm = self.reversed_labels_mapping
        for i in range(int(self.minibatch_size)):
            if self._n_saved[cls] >= self.limit:
                return
            true_label = labels[i] if labels is not None and len(labels) > i else -1
            if mx is not None and mx[i] == true_label:
                continue            # softmax: only failures are interesting
            img = self.as_image(inp[i])
            if img is None:
                continue
            tl = rlm[true_label] if rlm else true_label
            if mx is not None:
                pl = rlm[mx[i]] if rlm else mx[i]
                conf = float(out[i].ravel()[int(mx[i])]) if out is not None else 0.0
                tail = "%s_as_%s.%.0fpt.%d.png" % (tl, pl, conf * 100, idx[i])
                odir = self.out_dirs[cls]
                oimg = timg = None
            else:
                odir = os.path.join(self.out_dirs[cls], "%d" % idx[i])
                oimg = timg = None
                mse = None
                if out is not None and tgt is not None:
                    timg = self.as_image(tgt[i])
                    oimg = None if timg is None else out[i].reshape(timg.shape)
                    mse = float(numpy.linalg.norm(out[i].ravel() - tgt[i].ravel()) /
                                max(out[i].size, 1))
                tail = ("%.6f_%s_%d.png" % (mse, tl, idx[i]) if mse is not None
                        else "%s_%d.png" % (tl, idx[i]))
            os.makedirs(odir, exist_ok=True)
            self._save(normalize_image(img, self.color_space),
                       os.path.join(odir, "input_image_" + tail))
            if oimg is not None:
                self._save(normalize_image(oimg, self.color_space),
                           os.path.join(odir, "output_image_" + tail))
            if timg is not None:
                self._save(normalize_image(timg, self.color_space),
                           os.path.join(odir, "target_" + tail))
            self._n_saved[cls] += 1
