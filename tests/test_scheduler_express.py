"""Unit scheduler: the express path through a linear CUDA-graph segment (core/units.py,
core/graphs.py::GraphSegment.express) must run exactly the units the plain walk runs, in an order
that keeps every dependency, and fall back when the members do not form a plain chain."""
import os

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.graphs import GraphSegment
from veles.znicz_b200.core.units import Unit
from veles.znicz_b200.core.workflow import DummyWorkflow, Repeater


class _Rec(Unit):
    hide_from_registry = True
    log = []

    def initialize(self, **kwargs):
        pass

    def run(self):
        _Rec.log.append(self.name)


class _Seg(GraphSegment):
    """Segment stand-in: 'executes' all members when the first one runs (what replay does)."""

    def execute(self):
        for u in self.units:
            _Rec.log.append("seg:" + u.name)


def _build(branch=False):
    wf = DummyWorkflow()
    rep = Repeater(wf)
    rep.link_from(wf.start_point)
    loader = _Rec(wf, name="loader")
    loader.link_from(rep)
    chain = [_Rec(wf, name="f%d" % i) for i in range(4)]
    prev = loader
    for u in chain:
        u.link_from(prev)
        prev = u
    decision = _Rec(wf, name="decision")
    decision.link_from(chain[-1])
    extra = None
    if branch:                                 # a plotter hanging off the middle of the chain
        extra = _Rec(wf, name="plotter")
        extra.link_from(chain[1])
    gds = [_Rec(wf, name="g%d" % i) for i in range(3)]
    prev = decision
    for u in gds:
        u.link_from(prev)
        prev = u
    rep.link_from(gds[-1])
    for u in wf.units:
        u._is_initialized = True
    seg_f = _Seg("forward", chain, enabled=False)
    seg_b = _Seg("backward", gds, enabled=False)
    return wf, seg_f, seg_b


def _run(wf, iterations=2):
    _Rec.log = []
    wf.run(iterations=iterations)
    return list(_Rec.log)


def test_express_path_matches_plain_walk():
    wf, seg_f, seg_b = _build()
    assert seg_f.express and seg_b.express
    fast = _run(wf)
    os.environ["ZNICZ_EXPRESS"] = "0"
    try:
        wf2, _, _ = _build()
        slow = _run(wf2)
    finally:
        del os.environ["ZNICZ_EXPRESS"]
    assert fast == slow
    one = ["loader"] + ["seg:f%d" % i for i in range(4)] + ["decision"] + ["seg:g%d" % i for i in range(3)]
    assert fast == one * 2
    # the skipped members neither ran nor were queued, but their successors ran exactly once
    wf3, seg_f3, _ = _build()
    _run(wf3, iterations=3)
    assert [u._run_calls for u in seg_f3.units] == [3, 0, 0, 0]


def test_branching_segment_keeps_the_plain_walk():
    wf, seg_f, seg_b = _build(branch=True)
    assert not seg_f.express and seg_b.express
    log = _run(wf, iterations=1)
    assert log.count("plotter") == 1 and log.index("plotter") > log.index("seg:f1")
    assert log.count("decision") == 1


def test_tracing_disables_the_express_path():
    wf, seg_f, _ = _build()
    root.common.trace.run = True
    try:
        log = _run(wf, iterations=1)
        assert [u._run_calls for u in seg_f.units] == [1, 1, 1, 1]     # every member walked
        assert log.count("decision") == 1
    finally:
        root.common.trace.run = False
