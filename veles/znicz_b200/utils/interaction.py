"""Shell unit (``veles.interaction.Shell``): drops into an interactive console with the
workflow in scope. Without a tty (batch runs) it is a no-op."""
from __future__ import annotations

import code
import sys

from ..core.units import Unit


class Shell(Unit):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "SERVICE")
        super().__init__(workflow, **kwargs)
        self.enabled = kwargs.get("enabled", True)

    def initialize(self, **kwargs):
        pass

    def run(self):
        if not self.enabled or not sys.stdin.isatty():
            return
        try:
            from IPython import embed
            embed(user_ns={"workflow": self.workflow, "unit": self})
        except ImportError:
            code.interact(local={"workflow": self.workflow, "unit": self})
