"""Whole-workflow pickle snapshots (``veles.snapshotter``).

Parity: ``Snapshotter.run`` pickles the **entire workflow object**
when the decision reports an improvement at epoch end; file name
``<prefix>_<suffix>.<pickle-proto>.pickle[.gz|.bz2|.xz]``; resume =
``SnapshotterToFile.import_(file)`` → ``initialize(device, snapshot=True)`` → ``run()``
(/root/reference/tests/functional/test_cifar_caffe.py:210-221). Triggers ``interval`` /
``time_interval`` (/root/reference/samples/Wine/wine_config.py:51); gates
``gate_skip=~epoch_ended`` and ``skip=~improved`` are wired by StandardWorkflow.

Class-path aliases ``veles.znicz.* -> veles.znicz_b200.*`` are installed by the
top-level ``veles`` package so pickles keep reference-style module paths loadable.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import os
import pickle
import time

from .config import root
from .mutable import Bool
from .registry import make_registry
from .units import Unit

SnapshotterRegistry = make_registry("snapshotters")
# reference spelling: SnapshotterRegistry.snapshotters[name]
SnapshotterRegistry.snapshotters = SnapshotterRegistry.registry

_OPENERS = {
    None: (open, ""), "": (open, ""),
    "gz": (gzip.open, ".gz"), "bz2": (bz2.open, ".bz2"), "xz": (lzma.open, ".xz"),
}
try:  # snappy is optional in the reference too
    import snappy  # noqa: F401
    _HAS_SNAPPY = True
except Exception:
    _HAS_SNAPPY = False


class SnapshotterBase(Unit, metaclass=SnapshotterRegistry):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "SERVICE")
        super().__init__(workflow, **kwargs)
        self.prefix = kwargs.get("prefix", "")
        self.directory = kwargs.get("directory", root.common.dirs.snapshots)
        self.compression = kwargs.get("compression", "gz")
        self.compression_level = kwargs.get("compression_level", 6)
        self.interval = kwargs.get("interval", 1)
        self.time_interval = kwargs.get("time_interval", 15)
        self.time = 0
        self._skipped_counter = 0
        self.skip = Bool(False)
        self.suffix = None
        self.destination = None

    def initialize(self, **kwargs):
        self.time = time.time()
        self._skipped_counter = 0

    def run(self):
        """Returns True when a snapshot was actually taken."""
        if root.common.disable.get("snapshotting", False) or self.is_slave:
            return False
        if self._dp_rank() != 0:
            # data parallel: replicas are bit-identical, `improved` is all-reduced - only rank 0
            # writes (N ranks opening the same path with "wb" corrupt each other's pickle)
            return False
        self._skipped_counter += 1
        if bool(self.skip):
            return False
        if self._skipped_counter < self.interval:
            return False
        delta = time.time() - self.time
        if delta < self.time_interval and self._taken_once:
            return False
        self._skipped_counter = 0
        self.export()
        self.time = time.time()
        self._taken_once_ = True
        return True

    def init_unpickled(self):
        super().init_unpickled()
        self._taken_once_ = False

    def _dp_rank(self):
        dp = getattr(self.workflow, "dp_", None)
        if dp is not None:
            return int(dp.rank)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            return int(os.environ.get("RANK", "0"))
        return 0

    @property
    def _taken_once(self):
        return self._taken_once_

    def export(self):
        raise NotImplementedError

    @staticmethod
    def import_(file_name):
        return SnapshotterToFile.import_file(file_name)


class SnapshotterToFile(SnapshotterBase):
    MAPPING = "file"
    WRITE_CODECS = _OPENERS

    def export(self):
        ext = self.compression or ""
        if ext not in _OPENERS:
            if ext == "snappy" and not _HAS_SNAPPY:
                self.warning("snappy is unavailable, falling back to gz")
                ext = "gz"
            else:
                raise ValueError("Unknown compression %r" % ext)
        opener, fext = _OPENERS[ext]
        os.makedirs(self.directory, exist_ok=True)
        rel = "%s_%s.%d.pickle%s" % (
            self.prefix, self.suffix if self.suffix else "snapshot",
            pickle.HIGHEST_PROTOCOL, fext)
        path = os.path.join(self.directory, rel)
        t0 = time.time()
        wf = self.workflow
        # write-then-rename: a reader (or a crash) never sees a half-written snapshot
        tmp = "%s.tmp.%d" % (path, os.getpid())
        with opener(tmp, "wb") as fout:
            pickle.dump(wf, fout, protocol=pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, path)
        self.destination = path
        self.info("Snapshotted to %s in %.2f sec (%d bytes)", path,
                  time.time() - t0, os.path.getsize(path))
        link = os.path.join(self.directory, "%s_current.lnk" % self.prefix)
        try:
            if os.path.lexists(link):
                os.remove(link)
            os.symlink(rel, link)
        except OSError:  # pragma: no cover
            pass

    @staticmethod
    def import_file(file_name):
        file_name = os.path.realpath(file_name)   # "<prefix>_current.lnk" symlinks
        if file_name.endswith(".gz"):
            opener = gzip.open
        elif file_name.endswith(".bz2"):
            opener = bz2.open
        elif file_name.endswith(".xz"):
            opener = lzma.open
        else:
            opener = open
        with opener(file_name, "rb") as fin:
            return pickle.load(fin)


class SnapshotterToDB(SnapshotterBase):
    """ODBC variant of the reference (``nnodbc``). Without an ODBC driver in the
    image the blob is stored in a local sqlite database with the same columns."""
    MAPPING = "odbc"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.odbc = kwargs.get("odbc", os.path.join(
            str(root.common.dirs.snapshots), "snapshots.sqlite"))
        self.table = kwargs.get("table", "veles")

    def export(self):
        import sqlite3
        os.makedirs(os.path.dirname(self.odbc) or ".", exist_ok=True)
        blob = pickle.dumps(self.workflow, protocol=pickle.HIGHEST_PROTOCOL)
        if self.compression == "gz":
            blob = gzip.compress(blob, self.compression_level)
        con = sqlite3.connect(self.odbc)
        with con:
            con.execute("CREATE TABLE IF NOT EXISTS %s (timestamp REAL, id TEXT, "
                        "log_id TEXT, workflow TEXT, name TEXT, codec TEXT, data BLOB)"
                        % self.table)
            con.execute("INSERT INTO %s VALUES (?,?,?,?,?,?,?)" % self.table,
                        (time.time(), self.id, "", type(self.workflow).__name__,
                         "%s_%s" % (self.prefix, self.suffix), self.compression or "",
                         blob))
        con.close()
        self.destination = "%s#%s_%s" % (self.odbc, self.prefix, self.suffix)

    @staticmethod
    def import_db(odbc, table="veles", name=None):
        import sqlite3
        con = sqlite3.connect(odbc)
        q = "SELECT codec, data FROM %s %s ORDER BY timestamp DESC LIMIT 1" % (
            table, "WHERE name=?" if name else "")
        row = con.execute(q, (name,) if name else ()).fetchone()
        con.close()
        if row is None:
            raise KeyError("no snapshot in %s" % odbc)
        data = gzip.decompress(row[1]) if row[0] == "gz" else row[1]
        return pickle.loads(data)
