"""``veles.loader.pickles_image``: datasets stored as pickled batches (CIFAR-10 python format:
dict with ``data`` uint8 [N, 3072] and ``labels``)."""
import pickle

import numpy

from veles.loader.base import TRAIN, VALID, TEST
from veles.loader.fullbatch import FullBatchLoader
from veles.loader.image import ImageLoader
from veles import memory


class PicklesImageFullBatchLoader(FullBatchLoader, ImageLoader):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super(PicklesImageFullBatchLoader, self).__init__(workflow, **kwargs)
        self.color_space = kwargs.get("color_space", "RGB")
        self.test_pickles = list(kwargs.get("test_pickles", []))
        self.validation_pickles = list(kwargs.get("validation_pickles", []))
        self.train_pickles = list(kwargs.get("train_pickles", []))
        self.add_sobel = kwargs.get("add_sobel", False)

    def reshape(self, shape):
        return shape

    def transform_data(self, data):
        """[N, C, H, W] -> [N, H, W, C] (interleaved), as the units expect."""
        return memory.interleave(data)

    def _load_pickle(self, path):
        with open(path, "rb") as f:
            d = pickle.load(f, encoding="latin1")
        data = numpy.asarray(d["data"])
        labels = list(d.get("labels", d.get("fine_labels", [])))
        return data, labels

    def load_data(self):
        parts, labels = [], []
        for cls, files in ((TEST, self.test_pickles), (VALID, self.validation_pickles),
                           (TRAIN, self.train_pickles)):
            n = 0
            for p in files:
                data, lbl = self._load_pickle(p)
                self.reshape(data.shape[1:])
                data = self.transform_data(data)
                parts.append(data)
                labels.extend(lbl)
                n += data.shape[0]
            self.class_lengths[cls] = n
        self.original_data.reset(numpy.concatenate(parts).astype(self.dtype))
        self.original_labels = labels
