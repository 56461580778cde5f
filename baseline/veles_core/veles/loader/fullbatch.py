"""``veles.loader.fullbatch``: the whole dataset in one array; minibatches are gathered by
index — on the device when the dataset fits there (stock behaviour), on the host with a
per-minibatch upload when ``force_numpy`` is set."""
import numpy
from zope.interface import Interface

from veles.config import root
from veles.loader.base import Loader, LoaderMSEMixin, TRAIN
from veles.memory import Array
import veles.opencl_types as opencl_types


class IFullBatchLoader(Interface):
    pass


class FullBatchLoader(Loader):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super(FullBatchLoader, self).__init__(workflow, **kwargs)
        self.original_data = Array()
        self.original_labels = []
        self.validation_ratio = kwargs.get("validation_ratio", None)
        self.dtype = opencl_types.dtypes[root.common.engine.precision_type]

    def init_unpickled(self):
        super(FullBatchLoader, self).init_unpickled()
        self.sources_["fullbatch_loader"] = {}
        self._mapped_labels_ = Array()

    @property
    def on_device(self):
        return self.device is not None and self.device.exists and not self.force_numpy

    def create_minibatch_data(self):
        self.minibatch_data.reset(numpy.zeros(
            (self.max_minibatch_size,) + self.original_data.shape[1:], self.dtype))

    def analyze_dataset(self):
        # labels -> dense ints
        if self.original_labels is not None and len(self.original_labels):
            uniq = sorted(set(self.original_labels))
            self.labels_mapping = {l: i for i, l in enumerate(uniq)}
            self.reversed_labels_mapping = uniq
            self._mapped_labels_.reset(numpy.array(
                [self.labels_mapping[l] for l in self.original_labels], numpy.int32))
        # normalisation: analysed on TRAIN, applied to everything
        if self.normalization_type != "none":
            data = self.original_data.mem
            train = data[self.class_end_offsets[TRAIN - 1]:]
            self.normalizer.analyze(train)
            self.normalizer.normalize(data)

    def numpy_init(self):
        pass

    def cuda_init(self):
        if not self.on_device:
            return
        self.init_vectors(self.original_data, self._mapped_labels_, self.minibatch_data,
                          self.minibatch_labels, self.minibatch_indices)
        self.build_program(
            {"SAMPLE_SIZE": self.original_data.sample_size,
             "MAX_MINIBATCH_SIZE": self.max_minibatch_size,
             "original_data_dtype": opencl_types.numpy_dtype_to_opencl(self.original_data.dtype),
             "minibatch_data_dtype": opencl_types.numpy_dtype_to_opencl(self.minibatch_data.dtype)},
            "fullbatch_loader", dtype=self.minibatch_data.dtype)
        self.assign_kernel("fill_minibatch_data_labels")
        self.set_args(self.original_data, self.minibatch_data, self.device.skip(2),
                      self._mapped_labels_, self.minibatch_labels, self.minibatch_indices)
        self._krn_const = numpy.zeros(2, numpy.int32)

    def fill_indices(self, start, count):
        if not self.on_device:
            return super(FullBatchLoader, self).fill_indices(start, count)
        # device path: indices go up (one tiny H2D), data + labels are gathered by a kernel
        super(FullBatchLoader, self).fill_indices(start, count)
        self.unmap_vectors(self.original_data, self.minibatch_data, self._mapped_labels_,
                           self.minibatch_labels, self.minibatch_indices)
        self._krn_const[0] = count
        self._krn_const[1] = self.original_data.sample_size
        self.set_arg(2, self._krn_const[0:1])
        self.set_arg(3, self._krn_const[1:2])
        total = self.max_minibatch_size * self.original_data.sample_size
        block = 256
        self.execute_kernel(((total + block - 1) // block, 1, 1), (block, 1, 1))
        return True

    def fill_minibatch(self):
        idx = self.minibatch_indices.mem[:self.minibatch_size]
        self.minibatch_data.map_invalidate()
        numpy.take(self.original_data.mem, idx, axis=0,
                   out=self.minibatch_data.mem[:self.minibatch_size])
        self.minibatch_data.mem[self.minibatch_size:] = 0
        if self.has_labels and self._mapped_labels_:
            self.minibatch_labels.map_invalidate()
            self.minibatch_labels.mem[:self.minibatch_size] = self._mapped_labels_.mem[idx]
            self.minibatch_labels.mem[self.minibatch_size:] = -1


class FullBatchLoaderMSEMixin(LoaderMSEMixin):
    pass


class FullBatchLoaderMSE(FullBatchLoaderMSEMixin, FullBatchLoader):
    hide_from_registry = True
