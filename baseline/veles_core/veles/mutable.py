"""Mutable boolean expressions used as unit gates.

Parity target: ``veles.mutable.Bool`` as used by the reference
(/root/reference/standard_workflow.py:488-489,514-515,598-599): gates are built
once from ``~decision.complete``, ``a | b`` … and re-evaluated lazily every time
a unit checks ``gate_block`` / ``gate_skip``; ``b <<= value`` assigns in place so
every expression holding a reference sees the change.
"""



class Bool(object):
    """A boolean cell or a lazily evaluated expression over other cells."""

    __slots__ = ("_value", "_expr", "_ops", "on_true", "on_false", "__weakref__")

    def __init__(self, value=False):
        self._expr = None
        self._ops = ()
        self.on_true = None
        self.on_false = None
        self._value = bool(value)

    # -- construction of derived expressions -------------------------------
    @classmethod
    def _derived(cls, expr, *ops):
        b = cls()
        b._expr = expr
        b._ops = ops
        return b

    def __invert__(self):
        return Bool._derived("not", self)

    def __or__(self, other):
        return Bool._derived("or", self, _lift(other))

    __ror__ = __or__

    def __and__(self, other):
        return Bool._derived("and", self, _lift(other))

    __rand__ = __and__

    def __xor__(self, other):
        return Bool._derived("xor", self, _lift(other))

    __rxor__ = __xor__

    # -- evaluation ---------------------------------------------------------
    def __bool__(self):
        e = self._expr
        if e is None:
            return self._value
        ops = self._ops
        if e == "not":
            return not bool(ops[0])
        if e == "or":
            return bool(ops[0]) or bool(ops[1])
        if e == "and":
            return bool(ops[0]) and bool(ops[1])
        return bool(ops[0]) != bool(ops[1])

    __nonzero__ = __bool__

    # -- in-place assignment ------------------------------------------------
    def __ilshift__(self, value):
        """``b <<= x`` stores ``bool(x)`` (snapshot, not a live link)."""
        if self._expr is not None:
            raise RuntimeError("Derived Bool expressions are read-only")
        new = bool(value)
        old = self._value
        self._value = new
        if new != old:
            cb = self.on_true if new else self.on_false
            if cb is not None:
                cb(self)
        return self

    def set(self, value=True):
        self <<= value

    def unset(self):
        self <<= False

    @property
    def is_expression(self):
        return self._expr is not None

    def __repr__(self):
        if self._expr is None:
            return "<Bool %s>" % self._value
        return "<Bool %s%s = %s>" % (self._expr, self._ops, bool(self))

    # -- pickling (callbacks are transient) ----------------------------------
    def __getstate__(self):
        return {"value": self._value, "expr": self._expr, "ops": self._ops}

    def __setstate__(self, state):
        self._value = state["value"]
        self._expr = state["expr"]
        self._ops = state["ops"]
        self.on_true = None
        self.on_false = None


def _lift(x):
    return x if isinstance(x, Bool) else Bool(x)


class LinkableAttribute(object):
    """Data descriptor that forwards ``obj.<name>`` to another object's attribute.

    Installed on the *class* of the linking unit the first time any instance
    links that name (same trick as the Veles core); instances without a link for
    the name fall back to their own ``__dict__``.
    """

    def __init__(self, name):
        self.name = name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        links = obj.__dict__.get("_linked_attrs")
        if links:
            tgt = links.get(self.name)
            if tgt is not None:
                return getattr(tgt[0], tgt[1])
        try:
            return obj.__dict__[self.name]
        except KeyError:
            raise AttributeError(
                "%s has no attribute %r" % (type(obj).__name__, self.name))

    def __set__(self, obj, value):
        links = obj.__dict__.get("_linked_attrs")
        if links:
            tgt = links.get(self.name)
            if tgt is not None:
                if tgt[2]:  # two-way link: write through
                    setattr(tgt[0], tgt[1], value)
                    return
                # one-way: assignment breaks the link (Veles raises; we detach)
                del links[self.name]
        obj.__dict__[self.name] = value

    def __delete__(self, obj):
        links = obj.__dict__.get("_linked_attrs")
        if links and self.name in links:
            del links[self.name]
            return
        obj.__dict__.pop(self.name, None)

    @staticmethod
    def install(obj, name, other, other_name, two_way=False):
        cls = type(obj)
        cur = cls.__dict__.get(name)
        if not isinstance(cur, LinkableAttribute):
            # class-level plain attributes/properties must not be shadowed
            for klass in cls.__mro__:
                if name in klass.__dict__ and not isinstance(
                        klass.__dict__[name], LinkableAttribute):
                    member = klass.__dict__[name]
                    if isinstance(member, property) or callable(member):
                        raise AttributeError(
                            "Cannot link %s.%s: it is a class member" %
                            (cls.__name__, name))
            setattr(cls, name, LinkableAttribute(name))
        links = obj.__dict__.setdefault("_linked_attrs", {})
        obj.__dict__.pop(name, None)
        links[name] = (other, other_name, two_way)
