"""Loaders (fresh design for ``veles.loader`` + the reference's ``loader/`` package)."""
from .base import (Loader, LoaderMSEMixin, UserLoaderRegistry, TEST, VALID, TRAIN,  # noqa
                   CLASS_NAME, LoaderError)
from .fullbatch import FullBatchLoader, FullBatchLoaderMSE  # noqa
from . import synthetic  # noqa  (registers the synthetic_* loaders)
from . import image  # noqa  (registers the *file_image loaders)
from . import saver  # noqa  (registers minibatches_loader)
from . import loader_lmdb, loader_stl, imagenet_loader  # noqa
