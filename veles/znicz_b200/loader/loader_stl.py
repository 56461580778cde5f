"""STL-10 binary loader. Parity: /root/reference/loader/loader_stl.py:45-116.

``directory`` holds ``class_names.txt``, ``train_X.bin`` / ``train_y.bin`` (TRAIN) and
``test_X.bin`` / ``test_y.bin`` (VALID). Images are 3x96x96 uint8 planes stored
column-major (the dataset's MATLAB heritage), labels are 1-based uint8.
"""
from __future__ import annotations

import os

import numpy

from .base import TEST, VALID, TRAIN
from .image import FullBatchImageLoaderBase


class STL10FullBatchLoader(FullBatchImageLoaderBase):
    MAPPING = "full_batch_stl_10"
    SIZE = (96, 96)

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.directory = kwargs["directory"]
        self.size = tuple(kwargs.get("size", self.SIZE))
        self.scale = kwargs.get("scale")        # (w, h) served size, None = native

    def init_unpickled(self):
        super().init_unpickled()
        self._maps_ = {}

    @property
    def square(self):
        return self.size[0] * self.size[1] * 3

    def _open(self, cls):
        if cls in self._maps_:
            return self._maps_[cls]
        if not os.path.isdir(self.directory):
            raise ValueError("%r must be a directory" % self.directory)
        stem = {TRAIN: "train", VALID: "test"}[cls]
        x = numpy.memmap(os.path.join(self.directory, stem + "_X.bin"), dtype=numpy.uint8,
                         mode="r")
        y = numpy.fromfile(os.path.join(self.directory, stem + "_y.bin"), dtype=numpy.uint8)
        if x.size != y.size * self.square:
            raise ValueError("%s_X.bin holds %d bytes, expected %d images of %d bytes" % (
                stem, x.size, y.size, self.square))
        with open(os.path.join(self.directory, "class_names.txt")) as f:
            self._class_names = f.read().split()
        self._maps_[cls] = (x.reshape(y.size, 3, self.size[0], self.size[1]), y)
        return self._maps_[cls]

    def get_keys(self, index):
        if index == TEST:
            return []
        return [(index, i) for i in range(len(self._open(index)[1]))]

    def get_image_label(self, key):
        return self._class_names[int(self._open(key[0])[1][key[1]]) - 1]

    def get_image_data(self, key):
        planes = self._open(key[0])[0][key[1]]        # [c][col][row]
        img = numpy.ascontiguousarray(planes.transpose(2, 1, 0))
        if self.scale is not None and tuple(self.scale) != (img.shape[1], img.shape[0]):
            from .image import fit_image
            img = fit_image(img, tuple(self.scale))
        return img
