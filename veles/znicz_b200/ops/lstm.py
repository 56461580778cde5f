"""LSTM cell as a nested workflow + its BPTT workflow.

Parity: /root/reference/lstm.py (LSTM :52-143 — InputJoiner, 3× All2AllSigmoid gates,
All2AllTanh memory maker, ForwardTanh output activation, 3× Multiplier, Summator;
``simple`` = no peephole from the memory cell to the output gate; ``link_weights`` :137;
GDLSTM :146-308 — the hand-wired chain of GD units and Cutter1D slices).
Sequences are built by chaining cells and sharing weights with ``link_weights``; BPTT by
chaining ``GDLSTM.err_prev_output / err_prev_memory`` (SURVEY §5 "long-context").

One fix w.r.t. the reference: ``err_memory`` (the gradient arriving from the next time
step's forget path) is actually added to the memory-cell gradient (the reference demands
it but never wires it).
"""
from __future__ import annotations

import weakref

from ..core.accelerated_units import AcceleratedWorkflow
from ..core.input_joiner import InputJoiner
from ..core.mutable import Bool
from .activation import ForwardTanh, BackwardTanh
from .all2all import All2AllSigmoid, All2AllTanh
from .cutter import Cutter1D
from .gd import GDSigmoid, GDTanh
from .multiplier import Multiplier, GDMultiplier
from .nn_units import FullyConnectedOutput
from .summator import Summator

_GATE_KW = ("output_sample_shape", "weights_stddev", "bias_stddev", "weights_filling",
            "bias_filling", "weights_transposed", "include_bias", "rand")
_GD_KW = ("learning_rate", "learning_rate_bias", "weights_decay", "weights_decay_bias",
          "gradient_moment", "gradient_moment_bias", "l1_vs_l2", "factor_ortho",
          "weights_transposed", "include_bias", "apply_gradient", "accumulate_gradient",
          "acc_alpha", "acc_beta", "gd_alpha", "gd_beta")


class LSTM(FullyConnectedOutput, AcceleratedWorkflow):
    """One LSTM step.

    Must be assigned before initialize(): ``input``, ``prev_output``, ``prev_memory``.
    Updates after run(): ``output`` (hidden state), ``memory`` (cell state).
    """
    MAPPING = {"LSTM"}

    # The cell is a small dataflow graph; it is declared as data and wired by one loop.
    #   c_t = i . g + f . c_{t-1};   h_t = o . tanh(c_t)
    #   i, f, o = sigmoid(W [x, h_{t-1}] (+ peephole to c_t for o when not ``simple``))
    # node -> (factory, constructor kwargs taken from the gate kwargs?)
    _NODES = (
        ("ij", InputJoiner, False), ("input_gate", All2AllSigmoid, True),
        ("forget_gate", All2AllSigmoid, True), ("memory_maker", All2AllTanh, True),
        ("output_gate", All2AllSigmoid, True), ("output_activation", ForwardTanh, False),
        ("input_mul", Multiplier, False), ("forget_mul", Multiplier, False),
        ("summator", Summator, False), ("output_mul", Multiplier, False),
    )
    # consumer -> ((attribute of the consumer, producer, attribute of the producer), ...);
    # "self" is the cell itself, "@gate_src" the joiner feeding the output gate
    _EDGES = {
        "input_gate": (("input", "ij", "output"),),
        "forget_gate": (("input", "ij", "output"),),
        "memory_maker": (("input", "ij", "output"),),
        "input_mul": (("x", "input_gate", "output"), ("y", "memory_maker", "output")),
        "forget_mul": (("x", "forget_gate", "output"), ("y", "self", "prev_memory")),
        "summator": (("x", "input_mul", "output"), ("y", "forget_mul", "output")),
        "output_activation": (("input", "summator", "output"),),
        "output_gate": (("input", "@gate_src", "output"),),
        "output_mul": (("x", "output_gate", "output"), ("y", "output_activation", "output")),
    }

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.simple = kwargs.pop("simple", True)
        gate_kw = {k: kwargs[k] for k in _GATE_KW if k in kwargs}
        names = {"summator": "memory_cell"}
        for attr, factory, takes_gate_kw in self._NODES:
            kw = dict(gate_kw) if takes_gate_kw else {}
            if factory is not InputJoiner:
                kw["name"] = names.get(attr, attr)
            setattr(self, attr, factory(self, **kw))
        gate_src = "ij"
        if not self.simple:               # peephole: the output gate also sees the new cell state
            self.ij_output = InputJoiner(self)
            gate_src = "ij_output"

        def node(key):
            return self if key == "self" else getattr(self, gate_src if key == "@gate_src" else key)

        # data aliases, and the control edges they imply (a unit runs after its producers)
        self.ij.link_from(self.start_point)
        self.ij.link_inputs(self, "input", "prev_output")
        if not self.simple:
            self.ij_output.link_inputs(self.ij, "output")
            self.ij_output.link_inputs(self.summator, "output")
            self.ij_output.link_from(self.summator, self.ij)
        for consumer, edges in self._EDGES.items():
            unit = getattr(self, consumer)
            producers = []
            for mine, src, theirs in edges:
                unit.link_attrs(node(src), (mine, theirs))
                if src != "self":
                    producers.append(node(src))
            unit.link_from(*producers)
        self.end_point.link_from(self.output_mul)
        self.link_attrs(self.output_mul, "output")
        self.link_attrs(self.summator, ("memory", "output"))
        self.demand("input", "prev_output", "prev_memory")

    def link_weights(self, src):
        """Share all gate weights with another cell (unrolling over time)."""
        for attr in ("input_gate", "forget_gate", "memory_maker", "output_gate"):
            getattr(self, attr).link_attrs(getattr(src, attr), "weights", "bias")
        return self

    @property
    def gates(self):
        return (self.input_gate, self.forget_gate, self.memory_maker, self.output_gate)


class GDLSTM(AcceleratedWorkflow):
    """Backward pass of one LSTM step.

    Must be assigned before initialize(): ``err_output``, ``err_memory`` (may be None
    for the last step). Updates after run(): ``err_input``, ``err_prev_output``,
    ``err_prev_memory``.
    """
    MAPPING = {"LSTM"}

    def __init__(self, workflow, forward, **kwargs):
        if forward is None:
            raise ValueError("forward must be provided")
        super().__init__(workflow, **kwargs)
        gkw = {k: kwargs[k] for k in _GD_KW if k in kwargs}
        self.err_memory = kwargs.get("err_memory")
        self.gd_output_mul = GDMultiplier(self, name="gd_output_mul")
        self.gd_output_activation = BackwardTanh(self, name="gd_output_activation")
        self.gd_output_gate = GDSigmoid(self, name="gd_output_gate", **gkw)
        self.add_err_memory = Cutter1D(self, name="add_err_memory", alpha=1, beta=1)
        if not forward.simple:
            self.og_to_summator = Cutter1D(self, name="og_to_summator", alpha=1, beta=1)
            self.og_to_ij = Cutter1D(self, name="og_to_ij", alpha=1, beta=0)
        self.gd_forget_mul = GDMultiplier(self, name="gd_forget_mul")
        self.gd_input_mul = GDMultiplier(self, name="gd_input_mul")
        acc = dict(gkw, err_input_alpha=1, err_input_beta=1)
        self.gd_memory_maker = GDTanh(self, name="gd_memory_maker", **acc)
        self.gd_forget_gate = GDSigmoid(self, name="gd_forget_gate", **acc)
        self.gd_input_gate = GDSigmoid(self, name="gd_input_gate", **acc)
        self.ij_to_input = Cutter1D(self, name="ij_to_input", alpha=1, beta=0)
        self.ij_to_prev_output = Cutter1D(self, name="ij_to_prev_output", alpha=1, beta=0)

        # control order = reverse topological order of the forward cell (a linear chain)
        order = ["gd_output_mul", "gd_output_activation", "add_err_memory", "gd_output_gate"]
        if not forward.simple:
            order += ["og_to_summator", "og_to_ij"]
        order += ["gd_forget_mul", "gd_input_mul", "gd_forget_gate", "gd_memory_maker",
                  "gd_input_gate", "ij_to_input", "ij_to_prev_output"]
        tail = self.start_point
        for unit_name in order:
            tail = getattr(self, unit_name).link_from(tail)
        self.end_point.link_from(tail)

        self.gd_output_mul.link_attrs(self, "err_output")
        self.gd_output_mul.link_attrs(forward.output_mul, "x", "y")
        self.gd_output_gate.link_attrs(self.gd_output_mul, ("err_output", "err_x"))
        self.gd_output_gate.link_attrs(forward.output_gate, "weights", "bias", "input",
                                       "output")
        self.gd_output_gate.forward_unit = forward.output_gate
        self.gd_output_activation.link_attrs(self.gd_output_mul, ("err_output", "err_y"))
        self.gd_output_activation.link_attrs(forward.output_activation, "input", "output")
        # memory gradient from the next time step
        self.add_err_memory.link_attrs(self, ("input", "err_memory"))
        self.add_err_memory.link_attrs(self.gd_output_activation, ("output", "err_input"))
        self.add_err_memory.gate_skip = _NoneAttr(self, "err_memory")
        if not forward.simple:
            self.og_to_summator.link_attrs(self.gd_output_gate, ("input", "err_input"))
            self.og_to_summator.link_attrs(forward.ij_output, ("input_offset", "offset_1"),
                                           ("length", "length_1"))
            self.og_to_summator.link_attrs(self.gd_output_activation,
                                           ("output", "err_input"))
            self.og_to_ij.link_attrs(self.gd_output_gate, ("input", "err_input"))
            self.og_to_ij.link_attrs(forward.ij_output, ("input_offset", "offset_0"),
                                     ("length", "length_0"))
            first, first_attr = self.og_to_ij, "output"
        else:
            first, first_attr = self.gd_output_gate, "err_input"
        self.gd_forget_mul.link_attrs(self.gd_output_activation, ("err_output", "err_input"))
        self.gd_forget_mul.link_attrs(forward.forget_mul, "x", "y")
        self.link_attrs(self.gd_forget_mul, ("err_prev_memory", "err_y"))
        self.gd_forget_gate.link_attrs(self.gd_forget_mul, ("err_output", "err_x"))
        self.gd_forget_gate.link_attrs(forward.forget_gate, "weights", "bias", "input",
                                       "output")
        self.gd_forget_gate.forward_unit = forward.forget_gate
        self.gd_forget_gate.link_attrs(first, ("err_input", first_attr))
        self.gd_input_mul.link_attrs(self.gd_output_activation, ("err_output", "err_input"))
        self.gd_input_mul.link_attrs(forward.input_mul, "x", "y")
        self.gd_input_gate.link_attrs(self.gd_input_mul, ("err_output", "err_x"))
        self.gd_input_gate.link_attrs(forward.input_gate, "weights", "bias", "input",
                                      "output")
        self.gd_input_gate.forward_unit = forward.input_gate
        self.gd_input_gate.link_attrs(first, ("err_input", first_attr))
        self.gd_memory_maker.link_attrs(self.gd_input_mul, ("err_output", "err_y"))
        self.gd_memory_maker.link_attrs(forward.memory_maker, "weights", "bias", "input",
                                        "output")
        self.gd_memory_maker.forward_unit = forward.memory_maker
        self.gd_memory_maker.link_attrs(first, ("err_input", first_attr))
        self.ij_to_input.link_attrs(first, ("input", first_attr))
        self.ij_to_input.link_attrs(forward.ij, ("input_offset", "offset_0"),
                                    ("length", "length_0"))
        self.link_attrs(self.ij_to_input, ("err_input", "output"))
        self.ij_to_prev_output.link_attrs(first, ("input", first_attr))
        self.ij_to_prev_output.link_attrs(forward.ij, ("input_offset", "offset_1"),
                                          ("length", "length_1"))
        self.link_attrs(self.ij_to_prev_output, ("err_prev_output", "output"))
        self.demand("err_output")
        self.forward_ref = forward

    @property
    def gd_gates(self):
        return (self.gd_input_gate, self.gd_forget_gate, self.gd_memory_maker,
                self.gd_output_gate)


class _NoneAttr(Bool):
    """Gate that is true while ``getattr(obj, name)`` is None/empty."""
    __slots__ = ("obj", "name")

    def __init__(self, obj=None, name=None):
        super().__init__(False)
        self.obj = obj
        self.name = name

    def __bool__(self):
        v = getattr(self.obj, self.name, None)
        return v is None or not v

    def __getstate__(self):
        return {"obj": self.obj, "name": self.name}

    def __setstate__(self, state):
        Bool.__setstate__(self, {"value": False, "expr": None, "ops": ()})
        self.obj = state["obj"]
        self.name = state["name"]
