"""Data-parallel execution on one 8×B200 node (one process per GPU)."""
from .data_parallel import DataParallel  # noqa
