from veles.loader.base import (  # noqa: F401
    Loader, ILoader, LoaderMSEMixin, LoaderMSE, UserLoaderRegistry, CLASS_NAME, TRAIN, VALID,
    TEST, TRIAGE, LoaderError)
from veles.loader.fullbatch import (  # noqa: F401
    FullBatchLoader, IFullBatchLoader, FullBatchLoaderMSE, FullBatchLoaderMSEMixin)
from veles.loader.pickles import PicklesImageFullBatchLoader  # noqa: F401
from veles.loader.image import ImageLoader, IImageLoader  # noqa: F401
