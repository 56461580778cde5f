"""``veles.snapshotter``: whole-workflow pickles (gz). Disabled in the benchmark."""
import gzip
import os
import pickle
import time

from veles.config import root
from veles.mutable import Bool
from veles.unit_registry import MappedUnitRegistry
from veles.units import Unit


class SnapshotterRegistry(MappedUnitRegistry):
    mapping = "snapshotters"
    base = Unit


class SnapshotterBase(Unit, metaclass=SnapshotterRegistry):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "SERVICE")
        super(SnapshotterBase, self).__init__(workflow, **kwargs)
        self.prefix = kwargs.get("prefix", "")
        self.directory = kwargs.get("directory", root.common.dirs.snapshots)
        self.compression = kwargs.get("compression", "gz")
        self.interval = kwargs.get("interval", 1)
        self.time_interval = kwargs.get("time_interval", 15)
        self.time = 0
        self._skipped_counter = 0
        self.skip = Bool(False)
        self.suffix = None
        self.destination = None

    def initialize(self, **kwargs):
        self.time = time.time()

    def run(self):
        if root.common.disable.get("snapshotting", False) or self.is_slave:
            return False
        self._skipped_counter += 1
        if bool(self.skip) or self._skipped_counter < self.interval:
            return False
        self._skipped_counter = 0
        self.export()
        self.time = time.time()
        return True

    def export(self):
        raise NotImplementedError


class SnapshotterToFile(SnapshotterBase):
    MAPPING = "file"

    def export(self):
        os.makedirs(self.directory, exist_ok=True)
        path = os.path.join(self.directory, "%s_%s.%d.pickle.gz" % (
            self.prefix, self.suffix or "snapshot", pickle.HIGHEST_PROTOCOL))
        with gzip.open(path, "wb") as f:
            pickle.dump(self.workflow, f, protocol=pickle.HIGHEST_PROTOCOL)
        self.destination = path

    @staticmethod
    def import_(path):
        with gzip.open(path, "rb") as f:
            return pickle.load(f)


class SnapshotterToDB(SnapshotterBase):
    MAPPING = "odbc"

    def export(self):
        raise NotImplementedError("no ODBC in the shim")
