"""MinibatchesSaver / MinibatchesLoader (``veles.loader.saver``; linked by
StandardWorkflow.link_data_saver, /root/reference/standard_workflow.py:1121-1150).

The saver records the first pass over the dataset (every served minibatch, already
normalised) into one file; the loader replays that file as a dataset without re-running
the original preprocessing.

File layout (own design, seekable so the loader can mmap-free random access):
    magic  b"ZNMB1\\n"
    header pickle {class_lengths, max_minibatch_size, shape, dtype, has_labels,
                   labels_mapping, compression}
    records: for every minibatch  <u32 class> <u32 size> <u64 nbytes> <payload>
             payload = compress(data[:size].tobytes() + labels[:size].tobytes())
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import os
import pickle
import struct
import zlib

import numpy

from ..core.config import root
from ..core.units import Unit
from .base import TEST, VALID, TRAIN
from .fullbatch import FullBatchLoader

MAGIC = b"ZNMB1\n"
_CODECS = {
    None: (lambda b: b, lambda b: b), "": (lambda b: b, lambda b: b),
    "gz": (lambda b: gzip.compress(b, 4), gzip.decompress),
    "zlib": (lambda b: zlib.compress(b, 4), zlib.decompress),
    "bz2": (bz2.compress, bz2.decompress),
    "xz": (lzma.compress, lzma.decompress),
}
_CODECS["snappy"] = _CODECS["zlib"]     # python-snappy is not in the image


class MinibatchesSaver(Unit):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "SERVICE")
        super().__init__(workflow, **kwargs)
        self.file_name = os.path.abspath(kwargs.get("file_name", os.path.join(
            str(root.common.dirs.get("cache", ".")), "minibatches.dat")))
        self.compression = kwargs.get("compression", "zlib")
        if self.compression not in _CODECS:
            raise ValueError("unknown compression %r" % (self.compression,))
        self.class_chunk_sizes = [0, 0, 0]
        self.offset_table = []
        self.demand("minibatch_data", "minibatch_labels", "minibatch_class",
                    "class_lengths", "max_minibatch_size", "minibatch_size",
                    "shuffle_limit", "has_labels", "labels_mapping")

    def init_unpickled(self):
        super().init_unpickled()
        self._file_ = None
        self._saved_ = [0, 0, 0]
        self._closed_ = False

    def initialize(self, **kwargs):
        if self.shuffle_limit != 0:
            raise ValueError("You must disable shuffling in your loader (set "
                             "shuffle_limit to 0) to record minibatches")
        os.makedirs(os.path.dirname(self.file_name) or ".", exist_ok=True)
        self._file_ = open(self.file_name, "wb")
        self._file_.write(MAGIC)
        pickle.dump({
            "class_lengths": list(self.class_lengths),
            "max_minibatch_size": int(self.max_minibatch_size),
            "shape": tuple(self.minibatch_data.shape[1:]),
            "dtype": self.minibatch_data.dtype.str,
            "has_labels": bool(self.has_labels),
            "labels_mapping": dict(self.labels_mapping or {}),
            "compression": self.compression,
        }, self._file_, protocol=4)

    def run(self):
        if self._closed_:
            return
        cls, size = int(self.minibatch_class), int(self.minibatch_size)
        if self._saved_[cls] >= self.class_lengths[cls]:
            # a class came around the second time: the first pass is complete
            if all(s >= n for s, n in zip(self._saved_, self.class_lengths)):
                self.stop()
            return
        self.minibatch_data.map_read()
        payload = self.minibatch_data.mem[:size].tobytes()
        if self.has_labels:
            self.minibatch_labels.map_read()
            payload += self.minibatch_labels.mem[:size].astype(numpy.int32).tobytes()
        blob = _CODECS[self.compression][0](payload)
        self.offset_table.append(self._file_.tell())
        self._file_.write(struct.pack("<IIQ", cls, size, len(blob)))
        self._file_.write(blob)
        self._saved_[cls] += size
        self.class_chunk_sizes[cls] += 1

    def stop(self):
        if self._file_ is not None and not self._closed_:
            self._file_.close()
            self._closed_ = True
            self.info("Wrote %d minibatches to %s", len(self.offset_table), self.file_name)


def read_minibatches(file_name):
    """Generator of (class, data[size, ...], labels or None); first item is the header."""
    with open(file_name, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError("%s is not a minibatches file" % file_name)
        header = pickle.load(f)
        yield header
        decomp = _CODECS[header["compression"]][1]
        dtype = numpy.dtype(header["dtype"])
        sample = int(numpy.prod(header["shape"])) if header["shape"] else 1
        while True:
            rec = f.read(16)
            if len(rec) < 16:
                return
            cls, size, nbytes = struct.unpack("<IIQ", rec)
            raw = decomp(f.read(nbytes))
            nd = size * sample * dtype.itemsize
            data = numpy.frombuffer(raw[:nd], dtype=dtype).reshape(
                (size,) + tuple(header["shape"]))
            labels = (numpy.frombuffer(raw[nd:], dtype=numpy.int32)
                      if header["has_labels"] else None)
            yield cls, data, labels


class MinibatchesLoader(FullBatchLoader):
    """Replays a MinibatchesSaver file as a full-batch dataset."""
    MAPPING = "minibatches_loader"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.file_name = kwargs["file_name"]

    def load_data(self):
        it = read_minibatches(self.file_name)
        header = next(it)
        parts = {TEST: [], VALID: [], TRAIN: []}
        labs = {TEST: [], VALID: [], TRAIN: []}
        for cls, data, labels in it:
            parts[cls].append(data)
            if labels is not None:
                labs[cls].append(labels)
        order = (TEST, VALID, TRAIN)
        chunks = [c for k in order for c in parts[k]]
        self.original_data.reset(numpy.concatenate(chunks) if chunks else
                                 numpy.zeros((0,) + tuple(header["shape"])))
        self.class_lengths[:] = [sum(len(c) for c in parts[k]) for k in order]
        if header["has_labels"]:
            self.original_labels = [int(v) for k in order for c in labs[k] for v in c]
            # labels in the file are already mapped 0..n-1
            n = max(self.original_labels) + 1 if self.original_labels else 0
            self.labels_mapping = {i: i for i in range(n)}
            self.reversed_labels_mapping = list(range(n))
        self.saved_labels_mapping = header["labels_mapping"]
