"""String -> class registries.

Parity: ``veles.unit_registry.MappedUnitRegistry`` / ``veles.mapped_object_registry``
as used by the reference for evaluators (/root/reference/evaluator.py:58-68),
decisions (/root/reference/decision.py:71-80), LR policies, snapshotters, loaders
and normalizers, plus the forward/backward ``MatchingObject`` table
(/root/reference/nn_units.py:64-107) that forms the ``layers`` DSL.
"""
from __future__ import annotations

from collections import defaultdict


class UnitRegistry(type):
    """Metaclass that records every unit class (for introspection / CLI)."""
    units = set()

    def __init__(cls, name, bases, clsdict):
        super().__init__(name, bases, clsdict)
        if not clsdict.get("hide_from_registry", False):
            UnitRegistry.units.add(cls)


_tables = defaultdict(dict)


def make_registry(name, loss_key=None):
    """Create a metaclass with its own ``registry`` dict: classes that define
    ``MAPPING = "some_name"`` become reachable as ``Meta.registry["some_name"]``.
    With ``loss_key`` (e.g. "LOSS") a second table ``loss_mapping`` maps the loss
    function name to the MAPPING string."""
    table = _tables[name]
    losses = {}

    class _Registry(UnitRegistry):
        registry = table
        loss_mapping = losses

        def __init__(cls, cname, bases, clsdict):
            super().__init__(cname, bases, clsdict)
            mapping = clsdict.get("MAPPING")
            if isinstance(mapping, str) and mapping:
                table[mapping] = cls
                if loss_key and loss_key in clsdict:
                    losses[clsdict[loss_key]] = mapping

    _Registry.__name__ = "%sRegistry" % name.title().replace("_", "")
    return _Registry


class Match(list):
    """One ``layers`` DSL entry: exactly one Forward class + N backward classes."""

    forward_base = None  # set by ops.nn_units

    @property
    def forward(self):
        for item in self:
            if issubclass(item, Match.forward_base):
                return item
        raise IndexError("no forward unit registered")

    @property
    def has_forward(self):
        return any(issubclass(i, Match.forward_base) for i in self)

    @property
    def backwards(self):
        for item in self:
            if not issubclass(item, Match.forward_base):
                yield item


class MatchingObject(UnitRegistry):
    """Metaclass of all forward / GD units: ``MAPPING = {"conv_relu", ...}`` (a set)
    registers the class under each string (/root/reference/nn_units.py:86-107)."""
    mapping = defaultdict(Match)

    def __init__(cls, name, bases, clsdict):
        super().__init__(name, bases, clsdict)
        mapping = clsdict.get("MAPPING", None)
        if mapping is None:
            return
        if not isinstance(mapping, (set, frozenset)):
            raise TypeError("%s: MAPPING must be of type 'set'" % cls)
        fb = Match.forward_base
        for val in mapping:
            match = MatchingObject.mapping[val]
            if fb is not None and issubclass(cls, fb) and match.has_forward \
                    and cls is not match.forward:
                raise ValueError(
                    "%s: attempted to add a second Forward %s to %s" %
                    (val, cls, match.forward))
            match.append(cls)
