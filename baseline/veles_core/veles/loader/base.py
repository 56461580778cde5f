"""``veles.loader.base``: minibatch serving protocol (TEST -> VALID -> TRAIN per epoch, short
last minibatch, flags as mutable Bools, labels mapping, normalizers)."""
import numpy
from zope.interface import Interface

from veles.accelerated_units import AcceleratedUnit
from veles.memory import Array
from veles.mutable import Bool
from veles import normalization
from veles import prng
from veles.unit_registry import MappedUnitRegistry
from veles.units import Unit

TEST, VALID, TRAIN = 0, 1, 2
TRIAGE = {"train": TRAIN, "validation": VALID, "valid": VALID, "test": TEST}
CLASS_NAME = ["test", "validation", "train"]


class LoaderError(Exception):
    pass


class ILoader(Interface):
    pass


class UserLoaderRegistry(MappedUnitRegistry):
    mapping = "loaders"
    base = Unit

    @staticmethod
    def get_factory(name, **kwargs):
        cls = UserLoaderRegistry.loaders[name]

        def factory(workflow, **more):
            kw = dict(kwargs)
            kw.update(more)
            return cls(workflow, **kw)
        return factory


class Loader(AcceleratedUnit, metaclass=UserLoaderRegistry):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "LOADER")
        super(Loader, self).__init__(workflow, **kwargs)
        self.prng = kwargs.get("prng", prng.get())
        self.max_minibatch_size = kwargs.get("minibatch_size", 100)
        self.shuffle_limit = kwargs.get("shuffle_limit", numpy.iinfo(numpy.uint32).max)
        self.normalization_type = kwargs.get("normalization_type", "none")
        self.normalization_parameters = kwargs.get("normalization_parameters", {})
        self.train_ratio = kwargs.get("train_ratio", 1.0)
        self.testing_mode = kwargs.get("testing", False)
        self.class_lengths = [0, 0, 0]
        self.class_end_offsets = [0, 0, 0]
        self.total_samples = 0
        self.epoch_number = 0
        self.epoch_ended = Bool(False)
        self.train_ended = Bool(False)
        self.last_minibatch = Bool(False)
        self.complete = Bool(False)
        self.minibatch_class = 0
        self.minibatch_size = 0
        self.minibatch_offset = 0
        self.global_offset = 0
        self.samples_served = 0
        self.minibatch_data = Array(shallow_pickle=True)
        self.minibatch_indices = Array(shallow_pickle=True)
        self.minibatch_labels = Array(shallow_pickle=True)
        self.shuffled_indices = Array()
        self.labels_mapping = {}
        self.reversed_labels_mapping = []
        self.class_keys = [[], [], []]
        self.on_initialized = None
        self.exports = ["minibatch_data", "minibatch_labels", "minibatch_indices",
                        "minibatch_class", "minibatch_size", "minibatch_offset",
                        "last_minibatch", "epoch_ended", "epoch_number", "class_lengths"]
        self._normalizer = None

    # -- properties --------------------------------------------------------------------------
    @property
    def has_labels(self):
        return True

    @property
    def unique_labels_count(self):
        return len(self.labels_mapping)

    @property
    def normalizer(self):
        if self._normalizer is None:
            self._normalizer = normalization.factory(
                self.normalization_type, **dict(self.normalization_parameters))
        return self._normalizer

    @property
    def class_ended(self):
        return self.global_offset == self.class_end_offsets[self.minibatch_class]

    def derive_from(self, other):
        self.normalization_type = other.normalization_type
        self._normalizer = other.normalizer
        self.labels_mapping = other.labels_mapping
        self.reversed_labels_mapping = other.reversed_labels_mapping

    # -- to override -------------------------------------------------------------------------
    def load_data(self):
        raise NotImplementedError

    def create_minibatch_data(self):
        raise NotImplementedError

    def fill_minibatch(self):
        raise NotImplementedError

    def fill_indices(self, start, count):
        self.minibatch_indices.map_invalidate()
        self.shuffled_indices.map_read()
        self.minibatch_indices.mem[:count] = self.shuffled_indices.mem[start:start + count]
        self.minibatch_indices.mem[count:] = -1
        return False

    # -- life cycle --------------------------------------------------------------------------
    def initialize(self, device=None, **kwargs):
        super(Loader, self).initialize(device=device, **kwargs)
        self.load_data()
        self._update_total_samples()
        self.info("Samples number: test %d, validation %d, train %d", *self.class_lengths)
        self.max_minibatch_size = int(min(self.max_minibatch_size, max(self.class_lengths)))
        self.minibatch_labels.reset(numpy.zeros(self.max_minibatch_size, numpy.int32)
                                    if self.has_labels else None)
        self.minibatch_indices.reset(numpy.zeros(self.max_minibatch_size, numpy.int32))
        self.create_minibatch_data()
        if not self.shuffled_indices:
            self.shuffled_indices.reset(numpy.arange(self.total_samples, dtype=numpy.int32))
        self.analyze_dataset()
        if self.on_initialized is not None:
            self.on_initialized()
        self.global_offset = 0
        self.shuffle()

    def analyze_dataset(self):
        pass

    def _update_total_samples(self):
        acc = 0
        for i, n in enumerate(self.class_lengths):
            acc += n
            self.class_end_offsets[i] = acc
        self.total_samples = acc
        if acc == 0:
            raise LoaderError("There is no data to serve")

    def shuffle(self):
        if self.shuffle_limit <= 0 or self.class_lengths[TRAIN] == 0:
            return
        self.shuffle_limit -= 1
        self.shuffled_indices.map_write()
        self.prng.shuffle(self.shuffled_indices.mem[self.class_end_offsets[VALID]:])

    def class_index_by_offset(self, offset):
        for i, e in enumerate(self.class_end_offsets):
            if offset < e:
                return i
        raise LoaderError("offset %d is out of range" % offset)

    def run(self):
        if self.global_offset >= self.total_samples:
            self.global_offset = 0
            self.shuffle()
        off = self.global_offset
        cls = self.class_index_by_offset(off)
        count = min(self.max_minibatch_size, self.class_end_offsets[cls] - off)
        self.minibatch_class = cls
        self.minibatch_size = count
        self.global_offset = off + count
        self.minibatch_offset = self.global_offset
        self.samples_served += count
        if not self.fill_indices(off, count):
            self.fill_minibatch()
        self._update_flags()

    def _update_flags(self):
        cls = self.minibatch_class
        last_mb = self.global_offset == self.class_end_offsets[cls]
        self.last_minibatch <<= last_mb
        no_valid = self.class_lengths[VALID] == 0
        self.epoch_ended <<= last_mb and (
            cls == VALID or (cls == TRAIN and no_valid) or
            (cls == TEST and no_valid and self.class_lengths[TRAIN] == 0))
        self.train_ended <<= last_mb and cls == TRAIN
        if last_mb and self.global_offset >= self.total_samples:
            self.epoch_number += 1
            if self.testing_mode:
                self.complete <<= True

    numpy_run = cuda_run = ocl_run = run

    # -- trivially distributable -------------------------------------------------------------
    def generate_data_for_slave(self, slave):
        return None

    def generate_data_for_master(self):
        return None

    def apply_data_from_master(self, data):
        pass

    def apply_data_from_slave(self, data, slave):
        pass

    def drop_slave(self, slave):
        pass


class LoaderMSEMixin(object):
    pass


class LoaderMSE(LoaderMSEMixin, Loader):
    hide_from_registry = True
