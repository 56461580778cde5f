"""Publisher (``veles.publishing.Publisher``): collects metrics of the registered
``IResultProvider`` units into a report (json / markdown) when training completes."""
from __future__ import annotations

import json
import os
import time

from ..core.config import root
from ..core.units import Unit
from ..core.workflow import _json_default


class Publisher(Unit):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "SERVICE")
        super().__init__(workflow, **kwargs)
        self.backends = kwargs.get("backends", {"json": {}})
        self.directory = kwargs.get("directory", os.path.join(
            str(root.common.dirs.get("cache", ".")), "reports"))
        self.result_providers = set()
        self.loader_unit = None
        self.report = None

    def initialize(self, **kwargs):
        pass

    def gather(self):
        info = {"workflow": type(self.workflow).__name__, "time": time.time(), "results": {}}
        for p in self.result_providers:
            names = p.get_metric_names()
            vals = p.get_metric_values()
            for n in names:
                if n in vals:
                    info["results"][n] = vals[n]
        lu = self.loader_unit
        if lu is not None:
            info["dataset"] = {"class_lengths": list(lu.class_lengths),
                               "total_samples": lu.total_samples,
                               "normalization": getattr(lu, "normalization_type", None)}
        return info

    def run(self):
        if root.common.disable.get("publishing", False):
            self.report = self.gather()
            return
        self.report = self.gather()
        os.makedirs(self.directory, exist_ok=True)
        base = os.path.join(self.directory, "%s_%d" % (self.report["workflow"],
                                                      int(self.report["time"])))
        if "json" in self.backends:
            with open(base + ".json", "w") as f:
                json.dump(self.report, f, indent=1, default=_json_default)
        if "markdown" in self.backends:
            with open(base + ".md", "w") as f:
                f.write("# %s\n\n" % self.report["workflow"])
                for k, v in self.report["results"].items():
                    f.write("* **%s**: %s\n" % (k, v))
