"""Downloader unit (``veles.downloader.Downloader``): makes sure dataset files exist.
There is no network in this environment: when the files are already present the unit is
a no-op; otherwise it fetches ``url`` with urllib and unpacks tar/zip archives."""
from __future__ import annotations

import os
import tarfile
import zipfile

from ..core.config import root
from ..core.units import Unit


class Downloader(Unit):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "SERVICE")
        super().__init__(workflow, **kwargs)
        self.url = kwargs.get("url")
        self.directory = kwargs.get("directory", root.common.dirs.datasets)
        self.files = list(kwargs.get("files", []))

    @property
    def missing(self):
        return [f for f in self.files
                if not os.path.exists(os.path.join(str(self.directory), f))]

    def initialize(self, **kwargs):
        if not self.missing or not self.url:
            return
        import urllib.request
        os.makedirs(str(self.directory), exist_ok=True)
        dst = os.path.join(str(self.directory), os.path.basename(self.url))
        self.info("Downloading %s -> %s", self.url, dst)
        urllib.request.urlretrieve(self.url, dst)
        if tarfile.is_tarfile(dst):
            with tarfile.open(dst) as tar:
                tar.extractall(str(self.directory))
        elif zipfile.is_zipfile(dst):
            with zipfile.ZipFile(dst) as z:
                z.extractall(str(self.directory))
        if self.missing:
            raise RuntimeError("Downloader: still missing %s" % self.missing)

    def run(self):
        pass
