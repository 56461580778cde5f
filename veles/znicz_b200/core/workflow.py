"""Workflow: container + scheduler of units.

Fresh design for the absent ``veles.workflow`` (``Workflow``, ``Repeater``,
``NoMoreJobs``) and ``veles.plumbing.FireStarter`` / ``veles.dummy`` launchers.
Contract per SURVEY §8 and the reference usages
(/root/reference/standard_workflow.py:173-208,518-529,
/root/reference/tests/unit/test_gd_workflow.py:289-291,
/root/reference/samples/MNIST/mnist.py:123-126).

The run loop is a FIFO over (src, dst) control edges: a unit runs once all its
``links_from`` fired (``Repeater`` ignores that), unless ``gate_block``;
``gate_skip`` propagates without running. ``EndPoint.run`` finishes the loop.
"""
from __future__ import annotations

import collections
import hashlib
import io
import json
import os
import tarfile
import time
import zipfile

import numpy

from .config import root
from .mutable import Bool
from .units import Unit, TrivialUnit


class NoMoreJobs(Exception):
    """Raised by ``generate_data_for_slave`` when training is complete."""


class StartPoint(TrivialUnit):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Start")
        super().__init__(workflow, **kwargs)


class EndPoint(TrivialUnit):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "End")
        super().__init__(workflow, **kwargs)

    def run(self):
        self.workflow.on_workflow_finished()


class Repeater(TrivialUnit):
    """Loop head: fires on *any* incoming edge (ignores the all-links gate)."""
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Repeater")
        kwargs["ignores_gate"] = True
        super().__init__(workflow, **kwargs)


class FireStarter(Unit):
    """Resets ``stopped``/gates of a list of units (``veles.plumbing.FireStarter``)."""
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.units = set(kwargs.get("units", ()))

    def run(self):
        for u in self.units:
            u.stopped = False


class Workflow(Unit):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        self._units = []
        self._finished = Bool(False)
        self.result_file = kwargs.get("result_file")
        super().__init__(workflow, **kwargs)
        self.start_point = StartPoint(self)
        self.end_point = EndPoint(self)
        self.negotiates_on_connect = False
        self._run_started = None
        self._sync_event = None

    def init_unpickled(self):
        super().init_unpickled()
        self._queue_ = None
        self._restored_from_snapshot_ = None
        self.step_hooks_ = []

    # -- container ------------------------------------------------------------
    def add_ref(self, unit):
        if unit is self:
            raise ValueError("Workflow cannot contain itself")
        if unit not in self._units:
            self._units.append(unit)

    def del_ref(self, unit):
        if unit in self._units:
            self._units.remove(unit)

    @property
    def units(self):
        return list(self._units)

    def __iter__(self):
        return iter(self._units)

    def __len__(self):
        return len(self._units)

    def __getitem__(self, key):
        """Find a unit by name (str) or index."""
        if isinstance(key, str):
            found = [u for u in self._units if u.name == key]
            if not found:
                raise KeyError(key)
            return found[0] if len(found) == 1 else found
        return self._units[key]

    @property
    def units_in_dependency_order(self):
        order = self.start_point.dependent_units()
        seen = set(order)
        for u in self._units:  # units not reachable by control edges
            if u not in seen:
                order.append(u)
        return order

    # -- distributed role (delegated to the launcher) -------------------------
    @property
    def is_standalone(self):
        wf = self._workflow
        return True if wf is None else wf.is_standalone

    @property
    def is_master(self):
        wf = self._workflow
        return False if wf is None else wf.is_master

    @property
    def is_slave(self):
        wf = self._workflow
        return False if wf is None else wf.is_slave

    @property
    def is_running(self):
        return self._queue_ is not None

    @property
    def restored_from_snapshot(self):
        return bool(self._restored_from_snapshot_)

    # -- life cycle -----------------------------------------------------------
    def initialize(self, **kwargs):
        """Initialise units in dependency order; units that return True (or whose
        demands are not met yet) are re-queued until no progress is made."""
        snapshot = kwargs.get("snapshot", False)
        self._restored_from_snapshot_ = snapshot
        if isinstance(kwargs.get("device"), str):      # "numpy" / "cuda" / "auto"
            from .backends import get_device
            kwargs["device"] = get_device(kwargs["device"])
        if "device" in kwargs:
            self.__dict__["device"] = kwargs["device"]
        pending = [u for u in self.units_in_dependency_order if u is not self]
        max_rounds = len(pending) + 2
        for _ in range(max_rounds):
            if not pending:
                break
            nxt = []
            for u in pending:
                res = u._initialize_checked(**kwargs)
                if res:
                    nxt.append((u, res))
            if len(nxt) == len(pending):
                # no progress: report
                msgs = []
                for u, res in nxt:
                    if isinstance(res, list):
                        msgs.append("%s: unsatisfied demands %s" % (u, res))
                    else:
                        msgs.append("%s: initialize() keeps returning True" % u)
                raise RuntimeError(
                    "Workflow %s failed to initialize:\n  %s" %
                    (self.name, "\n  ".join(msgs)))
            pending = [u for u, _ in nxt]
        self._finished <<= False
        self._is_initialized = True
        return None

    def run(self, iterations=None):
        """Run until the EndPoint fires (or ``stop()``). With ``iterations=K`` the loop
        is paused after exactly K passes of the Repeater cycle (K minibatches); calling
        ``run`` again continues with the next minibatch. ``step_hooks_`` (callables taking
        the workflow) are invoked after every completed pass."""
        self._finished <<= False
        self._stopped = False
        self._run_started = time.time()
        Unit.refresh_trace_flags()
        q = collections.deque()
        self._queue_ = q
        hooks = self.step_hooks_
        count = 0
        try:
            sp = self.start_point
            sp._is_initialized = True
            sp._run_timed()
            sp.run_dependent()
            fin = self._finished
            pop = q.popleft
            while q and not fin._value:
                src, dst = pop()
                if src is not sp and type(dst) is Repeater:
                    count += 1
                    for h in hooks:
                        h(self)
                    if iterations is not None and count >= iterations:
                        break
                dst._check_gate_and_run(src)
        finally:
            self._queue_ = None
        self._run_time += time.time() - self._run_started
        return count

    def _check_gate_and_run(self, src):
        # nested workflow used as a unit
        if not self.open_gate(src):
            return
        if bool(self._gate_block):
            return
        if not bool(self._gate_skip):
            outer_q = self._queue_
            self.run()
            self._queue_ = outer_q
        self.run_dependent()

    def stop(self):
        self._stopped = True
        self._finished <<= True
        for u in self._units:
            try:
                u.stop()
            except Exception:  # pragma: no cover
                self.exception("stop() failed in %s", u)

    def on_workflow_finished(self):
        self._finished <<= True
        for u in self._units:
            u.stop() if u is not self.end_point and hasattr(u, "stop") else None
        if self.result_file:
            self.write_results(self.result_file)

    @property
    def finished(self):
        return self._finished

    # -- distributed protocol fan-out (IDistributable over all units) ---------
    def generate_data_for_slave(self, slave=None):
        data = []
        for u in self.units_in_dependency_order:
            fn = getattr(u, "generate_data_for_slave", None)
            data.append(fn(slave) if fn is not None and u is not self else None)
        return data

    def apply_data_from_master(self, data):
        for u, d in zip(self.units_in_dependency_order, data):
            fn = getattr(u, "apply_data_from_master", None)
            if fn is not None and d is not None and u is not self:
                fn(d)

    def generate_data_for_master(self):
        data = []
        for u in self.units_in_dependency_order:
            fn = getattr(u, "generate_data_for_master", None)
            data.append(fn() if fn is not None and u is not self else None)
        return data

    def apply_data_from_slave(self, data, slave=None):
        for u, d in zip(self.units_in_dependency_order, data):
            fn = getattr(u, "apply_data_from_slave", None)
            if fn is not None and d is not None and u is not self:
                fn(d, slave)

    def drop_slave(self, slave=None):
        for u in self._units:
            fn = getattr(u, "drop_slave", None)
            if fn is not None:
                fn(slave)

    # -- results / metrics ------------------------------------------------------
    def gather_results(self):
        results = {}
        for u in self._units:
            names = getattr(u, "get_metric_names", None)
            if names is None:
                continue
            vals = u.get_metric_values()
            for n in names():
                if n in vals:
                    results[n] = vals[n]
        return results

    def write_results(self, file=None):
        results = self.gather_results()
        if file is None:
            return results
        if isinstance(file, str):
            with open(file, "w") as fout:
                json.dump(results, fout, default=_json_default, indent=2)
        else:
            json.dump(results, file, default=_json_default, indent=2)
        return results

    def print_stats(self, top=10):
        rows = sorted(((u.total_run_time, u.name) for u in self._units), reverse=True)
        total = sum(r[0] for r in rows) or 1.0
        for t, n in rows[:top]:
            self.info("%-32s %8.3f s  %5.1f%%", n, t, 100 * t / total)

    # -- export for the native runtime ----------------------------------------
    def package_export(self, file_name, archive_format="zip", precision=32):
        """Write ``contents.json`` + ``NNNN_shape.npy`` (format spec:
        /root/reference/libZnicz/tests/workflow_files/mnist.zip,
        /root/reference/tests/functional/test_package_export.py:110-136)."""
        from ..export.package import package_export
        return package_export(self, file_name, archive_format, precision)

    def checksum(self):
        h = hashlib.sha1()
        h.update(type(self).__name__.encode())
        for u in self.units_in_dependency_order:
            h.update(type(u).__name__.encode())
        return h.hexdigest()


def _json_default(o):
    if isinstance(o, numpy.ndarray):
        return o.tolist()
    if isinstance(o, (numpy.integer,)):
        return int(o)
    if isinstance(o, (numpy.floating,)):
        return float(o)
    return str(o)


# ---------------------------------------------------------------------------
# Launchers (``veles.dummy`` / core launcher equivalents)
# ---------------------------------------------------------------------------
class DummyLauncher(object):
    """Standalone launcher: owns the device and the distributed role."""
    _is_launcher = True

    def __init__(self, mode="standalone", **kwargs):
        self.mode = mode
        self.workflow = None
        self.testing = kwargs.get("testing", False)
        self._children = []
        self.device = kwargs.get("device")
        self.stopped = False

    is_standalone = property(lambda self: self.mode == "standalone")
    is_master = property(lambda self: self.mode == "master")
    is_slave = property(lambda self: self.mode == "slave")

    def add_ref(self, wf):
        self.workflow = wf
        self._children.append(wf)

    def del_ref(self, wf):
        if wf in self._children:
            self._children.remove(wf)

    def on_workflow_finished(self):
        self.stopped = True

    def stop(self):
        self.stopped = True


class DummyWorkflow(Workflow):
    """A workflow with its own launcher; the parent of units in unit tests."""
    hide_from_registry = True

    def __init__(self, **kwargs):
        super().__init__(DummyLauncher(**kwargs), name=kwargs.get("name", "Dummy"))


class DummyUnit(TrivialUnit):
    hide_from_registry = True

    def __init__(self, workflow=None, **kwargs):
        super().__init__(workflow, **{k: v for k, v in kwargs.items()
                                      if k in ("name", "view_group")})
        self.__dict__.update({k: v for k, v in kwargs.items()
                              if k not in ("name", "view_group")})
        self._is_initialized = True
