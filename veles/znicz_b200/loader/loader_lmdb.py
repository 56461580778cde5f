"""LMDB (Caffe dataset) loader. Parity: /root/reference/loader/loader_lmdb.py:50-169.

Every record is a Caffe ``Datum`` (CHW uint8 planes + int label), decoded by the wire
codec in ``loader/caffe/protobuf2.py``. The environment is opened with the ``lmdb``
binding when it is importable and with the pure-python reader (``lmdb_mini``) otherwise.
kwargs: ``train_path`` / ``validation_path`` / ``test_path`` (LMDB directories),
``db_color_space`` (colour space stored in the DB, default RGB), ``color_space`` (served),
``db_splitted_channels`` (True = CHW planes, the Caffe default), ``use_cache``.
"""
from __future__ import annotations

import numpy

from .base import TEST, VALID, TRAIN, LoaderError
from .caffe.protobuf2 import Datum
from .image import ImageLoaderBase, FullBatchImageLoaderBase


def _open_env(path):
    try:
        import lmdb
        return lmdb.open(path, readonly=True, lock=False)
    except ImportError:
        from . import lmdb_mini
        return lmdb_mini.open(path)


class _LMDBMixin(object):
    def _init_lmdb(self, kwargs):
        self._files = (kwargs.get("test_path"), kwargs.get("validation_path"),
                       kwargs.get("train_path"))
        self.db_color_space = kwargs.get("db_color_space", "RGB")
        self.color_space = kwargs.get("color_space", self.db_color_space)
        self.db_splitted_channels = bool(kwargs.get("db_splitted_channels", True))
        self.use_cache = kwargs.get("use_cache", True)
        self.cache_hits = self.cache_misses = 0

    def _init_lmdb_transient(self):
        self._envs_ = [None] * 3
        self._cursors_ = [None] * 3
        self._cache_ = (None, None)

    files = property(lambda self: self._files)

    def _cursor(self, index):
        if self._files == (None, None, None):
            raise OSError("No LMDB path was given (train_path / validation_path / test_path)")
        if self._cursors_[index] is None and self._files[index]:
            env = _open_env(self._files[index])
            self._envs_[index] = env
            self._cursors_[index] = env.begin().cursor()
        return self._cursors_[index]

    def get_keys(self, index):
        cur = self._cursor(index)
        if cur is None:
            return []
        keys = []
        ok = cur.first()
        while ok:
            keys.append((index, cur.key()))
            ok = cur.next()
        return keys

    def get_datum(self, key):
        if self.use_cache and key == self._cache_[0]:
            self.cache_hits += 1
            return self._cache_[1]
        self.cache_misses += 1
        index, dkey = key
        raw = self._cursor(index).get(dkey)
        if raw is None:
            raise LoaderError("LMDB key %r vanished" % (dkey,))
        datum = Datum.FromString(raw)
        self._cache_ = (key, datum)
        return datum

    def get_image_label(self, key):
        return self.get_datum(key).label

    def get_image_data(self, key):
        img = self.get_datum(key).to_hwc(self.db_splitted_channels)
        if self.color_space != self.db_color_space and img.shape[2] == 3:
            import cv2
            img = cv2.cvtColor(numpy.ascontiguousarray(img), getattr(
                cv2, "COLOR_%s2%s" % (self.db_color_space, self.color_space)))
            if img.ndim == 2:
                img = img[:, :, None]
        return img

    def stop(self):
        super().stop()
        total = self.cache_hits + self.cache_misses
        if total:
            self.info("Datum cache hits/misses: %d/%d (%d%%)", self.cache_hits,
                      self.cache_misses, self.cache_hits * 100 // total)


class LMDBLoader(_LMDBMixin, ImageLoaderBase):
    MAPPING = "lmdb"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self._init_lmdb(kwargs)

    def init_unpickled(self):
        super().init_unpickled()
        self._init_lmdb_transient()


class FullBatchLMDBLoader(_LMDBMixin, FullBatchImageLoaderBase):
    """Whole LMDB decoded into memory once (then optionally resident in HBM)."""
    MAPPING = "full_batch_lmdb"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self._init_lmdb(kwargs)

    def init_unpickled(self):
        super().init_unpickled()
        self._init_lmdb_transient()
