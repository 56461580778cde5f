"""StandardWorkflowBase: builds the forward chain from a ``layers`` list.

Parity: /root/reference/standard_workflow_base.py (StandardWorkflowBase :59,
``link_forwards`` :272, ``_get_layer_type_kwargs`` :406, ``_add_forward_unit`` :424,
MCDNNIC topology strings like ``"12x256x256-32C4-MP2-64C4-MP3-32N-4N"`` :72-82,229-270,
auto-sizing of the last layer from the loader :310-335).
"""
from __future__ import annotations

import re
from collections import namedtuple

import numpy

from ..core.config import Config
from ..core.workflow import FireStarter
from ..loader.base import UserLoaderRegistry, LoaderMSEMixin
from ..ops import nn_units
from ..ops.all2all import All2AllSoftmax
from ..ops.dropout import DropoutForward
from ..ops.weights_zerofilling import ZeroFiller
from ..core.registry import MatchingObject

BaseWorkflowConfig = namedtuple("BaseWorkflowConfig", ("loader",))


def _reset_unit(fn):
    def wrapped(self, *args, **kwargs):
        self.unlink_unit(fn.__name__[5:])
        return fn(self, *args, **kwargs)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


def _check_forward_units(fn):
    def wrapped(self, *args, **kwargs):
        self._check_forwards()
        return fn(self, *args, **kwargs)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


def _check_backward_units(fn):
    def wrapped(self, *args, **kwargs):
        self._check_gds()
        return fn(self, *args, **kwargs)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


class _LastLayerSizer(object):
    """Picklable ``loader.on_initialized`` callback: sizes the last layer from the
    loader (labels count or target shape), /root/reference/standard_workflow_base.py:
    310-335."""

    def __init__(self, workflow, last_fwd):
        self.workflow = workflow
        self.last_fwd = last_fwd

    def __call__(self):
        loader = self.workflow.real_loader
        last_fwd = self.last_fwd
        if last_fwd.is_initialized:
            return
        if isinstance(loader, LoaderMSEMixin):
            last_fwd.output_sample_shape = tuple(loader.targets_shape)
        elif isinstance(last_fwd, All2AllSoftmax):
            ulc = loader.unique_labels_count
            oss = last_fwd.output_sample_shape
            if oss != tuple() and numpy.prod(oss) != ulc:
                self.workflow.warning(
                    "Overriding %s.output_sample_shape %s with (%s,)", last_fwd, oss, ulc)
            last_fwd.output_sample_shape = ulc


class StandardWorkflowBase(nn_units.NNWorkflow):
    """
    Arguments:
        layers: list of layer dicts ``{"type", "->": fwd kwargs, "<-": gd kwargs, ...}``
        loader_name / loader_factory: which Loader to create
        loader_config: loader kwargs
        mcdnnic_topology: alternative compact topology string
    """
    WorkflowConfig = BaseWorkflowConfig
    mcdnnic_topology_regexp = re.compile(
        r"(\d+)x(\d+)x(\d+)(-(?:(\d+C\d+)|(MP\d+)|(\d+N)))*$")
    mcdnnic_layer_patern = re.compile(r"(?P<C>\d+C\d+)|(?P<MP>MP\d+)|(?P<N>\d+N)")

    reset_unit = staticmethod(_reset_unit)
    check_forward_units = staticmethod(_check_forward_units)
    check_backward_units = staticmethod(_check_backward_units)

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.layer_map = MatchingObject.mapping
        self._preprocessing = kwargs.get("preprocessing", False)
        self._mcdnnic_topology = None
        self.mcdnnic_topology = kwargs.get("mcdnnic_topology", None)
        self.mcdnnic_parameters = kwargs.get("mcdnnic_parameters", None)
        self._layers = [{}]
        self.layers = kwargs.get("layers", [{}])
        self._loader_name = None
        self._loader_factory = None
        self.real_loader = None
        self.apply_config(**kwargs)
        if "loader_name" in kwargs:
            self.loader_name = kwargs["loader_name"]
        elif "loader_factory" in kwargs:
            self.loader_factory = kwargs["loader_factory"]
        else:
            raise KeyError("loader_name or loader_factory must be given")

    # -- configuration ----------------------------------------------------------------------
    @property
    def loader_name(self):
        return self._loader_name

    @loader_name.setter
    def loader_name(self, value):
        if value is None:
            self._loader_name = None
            return
        loader_kwargs = dict(self.dictify(self.config.loader))
        if self.mcdnnic_topology is not None:
            loader_kwargs = self._update_loader_kwargs_from_mcdnnic(
                loader_kwargs, self.mcdnnic_topology)
        self._loader_factory = UserLoaderRegistry.get_factory(value, **loader_kwargs)
        self._loader_name = value

    @property
    def loader_factory(self):
        return self._loader_factory

    @loader_factory.setter
    def loader_factory(self, value):
        if not callable(value):
            raise TypeError("loader_factory must be callable")
        self._loader_name = None
        self._loader_factory = value

    def unlink_unit(self, remove_unit_name):
        unit = self.__dict__.get(remove_unit_name)
        if unit is not None and hasattr(unit, "unlink_all"):
            self.warning("Instance %s exists. It will be removed and unlinked",
                         remove_unit_name)
            unit.unlink_all()
            self.del_ref(unit)

    def apply_config(self, **kwargs):
        old_config = getattr(self, "config", None)
        self.config = self.WorkflowConfig(**{
            f: self.config2kwargs(kwargs.pop("%s_config" % f,
                                             getattr(old_config, f, {})))
            for f in self.WorkflowConfig._fields})

    @staticmethod
    def dictify(obj):
        if isinstance(obj, Config):
            return obj.to_dict()
        return obj

    def config2kwargs(self, unit_config):
        return {} if unit_config is None else dict(self.dictify(unit_config))

    @property
    def mcdnnic_topology(self):
        return self._mcdnnic_topology

    @mcdnnic_topology.setter
    def mcdnnic_topology(self, value):
        if value is not None:
            if not isinstance(value, str):
                raise TypeError("mcdnnic_topology must be a string")
            if not self.mcdnnic_topology_regexp.match(value):
                raise ValueError(
                    "mcdnnic_topology value must match the following regular "
                    "expression: %s (got %s)" %
                    (self.mcdnnic_topology_regexp.pattern, value))
        self._mcdnnic_topology = value

    @property
    def layers(self):
        if self.mcdnnic_topology is not None:
            return self._get_layers_from_mcdnnic(self.mcdnnic_topology)
        return self._layers

    @layers.setter
    def layers(self, value):
        if self.mcdnnic_topology is not None and value != [{}]:
            raise ValueError(
                "Please do not set mcdnnic_topology and layers at the same time.")
        if not isinstance(value, list):
            raise ValueError("layers should be a list of dicts")
        if value == [{}] and self.mcdnnic_topology is None and not self.preprocessing:
            raise ValueError(
                "layers is empty and mcdnnic_topology is not defined: set layers "
                "(list of dicts) or an MCDNNIC topology string")
        for layer in value:
            if not isinstance(layer, dict):
                raise ValueError("layers should be a list of dicts")
        self._layers = value

    @property
    def preprocessing(self):
        return self._preprocessing

    @preprocessing.setter
    def preprocessing(self, value):
        self._preprocessing = value

    # -- MCDNNIC ------------------------------------------------------------------------------
    def _get_mcdnnic_parameters(self, arrow):
        if self.mcdnnic_parameters is not None and arrow in self.mcdnnic_parameters:
            return dict(self.mcdnnic_parameters[arrow])
        return {}

    @staticmethod
    def _parse_mcdnnic_c(last, value):
        kernels, kx = value.split("C")
        return {"type": "conv",
                "->": {"n_kernels": int(kernels), "kx": int(kx), "ky": int(kx)}}

    @staticmethod
    def _parse_mcdnnic_mp(last, value):
        _, kx = value.split("MP")
        return {"type": "max_pooling", "->": {"kx": int(kx), "ky": int(kx)}}

    @staticmethod
    def _parse_mcdnnic_n(last, value):
        neurons, _ = value.split("N")
        return {"type": "softmax" if last else "all2all",
                "->": {"output_sample_shape": int(neurons)}}

    def _get_layers_from_mcdnnic(self, description):
        parse = {"C": self._parse_mcdnnic_c, "N": self._parse_mcdnnic_n,
                 "MP": self._parse_mcdnnic_mp}
        layers = []
        matches = tuple(re.finditer(self.mcdnnic_layer_patern, description))
        for index, match in enumerate(matches):
            name = next(n for n, v in match.groupdict().items() if v)
            layer_config = parse[name](index == len(matches) - 1, match.group(name))
            layer_config["->"].update(self._get_mcdnnic_parameters("->"))
            layer_config["<-"] = self._get_mcdnnic_parameters("<-")
            layers.append(layer_config)
        return layers

    @staticmethod
    def _update_loader_kwargs_from_mcdnnic(kwargs, description):
        inp = description.split("-")[0]
        minibatch_size, y_size, x_size = inp.split("x")
        kwargs["minibatch_size"] = int(minibatch_size)
        kwargs["scale"] = (int(y_size), int(x_size))
        return kwargs

    # -- graph construction -------------------------------------------------------------------
    def _check_forwards(self):
        if not self.forwards:
            raise ValueError("forwards is empty: call link_forwards() first")

    def _check_gds(self):
        if not self.gds:
            raise ValueError("gds is empty: call link_gds() first")

    def link_forwards(self, init_attrs, *parents):
        """Create forward units from ``layers`` and chain them after ``parents``;
        the first unit's ``init_attrs`` = (mine, theirs) is linked to parents[0]."""
        del self.forwards[:]
        for layer in self.layers:
            tpe, kwargs, _ = self._get_layer_type_kwargs(layer)
            try:
                unit = self.layer_map[tpe].forward(self, **kwargs)
            except IndexError:
                raise ValueError("Failed to find a Forward in %s" % tpe) from None
            self._add_forward_unit(unit, init_attrs, *parents)
        # ZeroFiller masks the *next* layer's weights
        for prev_forward, forward in zip(self.forwards, self.forwards[1:]):
            if isinstance(prev_forward, ZeroFiller):
                prev_forward.link_attrs(forward, "weights")
        last_fwd = self.forwards[-1]
        if not isinstance(last_fwd, All2AllSoftmax) and \
                not isinstance(self.real_loader, LoaderMSEMixin):
            return last_fwd

        self.real_loader.on_initialized = _LastLayerSizer(self, last_fwd)
        return last_fwd

    def link_repeater(self, *parents):
        self.repeater.link_from(*parents)
        return self.repeater

    def link_fire_starter(self, *parents):
        self.fire_starter = FireStarter(self)
        self.fire_starter.link_from(*parents)
        return self.fire_starter

    def link_loader(self, *parents):
        self.loader = self.loader_factory(self)
        self.loader.link_from(*parents)
        self.real_loader = self.loader
        return self.loader

    def link_end_point(self, *parents):
        for sink in (self.repeater, self.end_point):      # close the loop and offer the exit
            sink.link_from(*parents)
        return self.end_point

    def create_workflow(self):
        """Forward-only skeleton: start -> repeater -> loader -> forwards; the exit stays shut
        until the loader reports completion."""
        loader = self.link_loader(self.link_repeater(self.start_point))
        self.link_forwards(("input", "minibatch_data"), loader)
        self.end_point.gate_block = ~loader.complete

    _LAYER_META_KEYS = frozenset(("type", "->", "<-", "name"))

    def _get_layer_type_kwargs(self, layer):
        """One ``layers`` entry -> (registry type, forward kwargs, GD kwargs): "->" feeds the
        forward unit, "<-" the GD unit, every other key both; ``name`` gets a per-direction
        suffix (DSL semantics: /root/reference/standard_workflow_base.py:406-422)."""
        kind = str(layer.get("type", "")).strip()
        if kind == "":
            raise ValueError("layer type must not be an empty string")
        if kind not in self.layer_map:
            raise ValueError("Unknown layer type %s" % kind)
        shared = {k: v for k, v in layer.items() if k not in self._LAYER_META_KEYS}
        per_direction = []
        for arrow, suffix in (("->", "_forward"), ("<-", "_backward")):
            kw = dict(layer.get(arrow, {}), **shared)
            if "name" in layer:
                kw["name"] = layer["name"] + suffix
            per_direction.append(kw)
        return kind, per_direction[0], per_direction[1]

    def _add_forward_unit(self, new_unit, init_attrs=None, *parents):
        if self.forwards:
            prev = (self.forwards[-1],)
        else:
            if not parents:
                raise ValueError("No parent units were specified for the first forward!")
            prev = parents
        new_unit.link_from(*prev)
        if isinstance(new_unit, DropoutForward):
            new_unit.link_attrs(self.loader, "minibatch_class")
        self.forwards.append(new_unit)
        if "input" not in new_unit.demanded and not hasattr(new_unit, "input"):
            return
        for fwd in reversed(self.forwards[:-1]):
            if hasattr(fwd, "output"):
                new_unit.link_attrs(fwd, ("input", "output"))
                break
        else:
            new_unit.link_attrs(parents[0], init_attrs)
