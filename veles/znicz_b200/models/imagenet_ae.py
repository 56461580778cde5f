"""ImagenetAE: greedy layer-wise convolutional auto-encoder pre-training followed by
supervised fine-tuning, with graph surgery on a live (or unpickled) workflow.

Capability parity with /root/reference/tests/research/ImagenetAE/imagenet_ae.py:124-752 and
imagenet_ae_config.py:101-170:

* the ``layers`` list contains pseudo layers ``{"type": "ae_begin"}`` / ``{"type": "ae_end"}``
  that delimit auto-encoder blocks (conv [+ activation] ... + one stochastic pooling);
  everything after the last ``ae_end`` is the classifier tail (FC / dropout / softmax);
* stage *k* trains block *k* only: encoder units of the earlier blocks are frozen, the
  block's pooling runs as a pooling+depooling unit, the decoder is a chain of ``Deconv``
  units with weights tied to the block's convolutions, the loss is MSE against the block's
  own input (``EvaluatorMSE`` + ``DecisionMSE``);
* ``add_ae_layer()`` — called by ``initialize`` when the workflow was restored from a
  snapshot with ``from_snapshot_add_layer`` — removes decoder / evaluator / decision / GD
  units, swaps ``Stochastic(Abs)PoolingDepooling`` for the plain pooling and stacks the next
  block (/root/reference/.../imagenet_ae.py:205-381);
* ``switch_to_fine_tuning()`` converts the stack into a softmax classifier: the tail layers
  are appended, the evaluator/decision become softmax/``DecisionGD`` and every layer gets a
  GD unit with its ``learning_rate_ft`` (:383-455).

The design differs from the reference in that stages are explicit (``stage`` counter, one
``_build_stage``) instead of being re-derived from the unit list.
"""
from __future__ import annotations

from ..core.config import root
from ..core.normalization import NoneNormalizer
from ..ops import conv, deconv, gd_deconv, pooling
from ..ops.conv import ConvolutionalBase
from ..ops.gd_pooling import GDPooling
from ..workflow import decision, evaluator
from ..workflow.standard_workflow import StandardWorkflow

TO_DEPOOL = {pooling.StochasticAbsPooling: pooling.StochasticAbsPoolingDepooling,
             pooling.StochasticPooling: pooling.StochasticPoolingDepooling}
FROM_DEPOOL = {v: k for k, v in TO_DEPOOL.items()}

root.imagenet_ae.update({
    "loader_name": "imagenet_loader_base",
    "loader": {"minibatch_size": 32, "sx": 216, "sy": 216, "channels": 3,
               "normalization_type": "none"},
    "decision_mse": {"fail_iterations": 50, "max_epochs": 50},
    "decision_gd": {"fail_iterations": 70, "max_epochs": 300},
    "snapshotter": {"prefix": "imagenet_ae", "interval": 1, "time_interval": 0},
    "from_snapshot_add_layer": True,
    "add_epochs": 0,
    "fine_tuning_noise": 1.0e-6,
    "layers": [
        {"type": "ae_begin"},
        {"name": "conv1", "type": "conv",
         "->": {"n_kernels": 108, "kx": 9, "ky": 9, "sliding": (3, 3),
                "weights_filling": "gaussian", "weights_stddev": 0.01, "include_bias": False},
         "<-": {"learning_rate": 2.0e-6, "learning_rate_ft": 2.0e-4, "weights_decay": 0.0005,
                "gradient_moment": 0.9}},
        {"name": "pool1", "type": "stochastic_abs_pooling",
         "->": {"kx": 3, "ky": 3, "sliding": (3, 3)}},
        {"type": "ae_end"},
        {"name": "mul1", "type": "activation_mul"},
        {"type": "ae_begin"},
        {"name": "conv2", "type": "conv",
         "->": {"n_kernels": 192, "kx": 5, "ky": 5, "sliding": (1, 1),
                "weights_filling": "gaussian", "weights_stddev": 0.01, "include_bias": False},
         "<-": {"learning_rate": 2.0e-6, "learning_rate_ft": 2.0e-4, "weights_decay": 0.0005,
                "gradient_moment": 0.9}},
        {"name": "pool2", "type": "stochastic_abs_pooling",
         "->": {"kx": 2, "ky": 2, "sliding": (2, 2)}},
        {"type": "ae_end"},
        {"name": "fc3", "type": "all2all_tanh",
         "->": {"output_sample_shape": 512, "weights_stddev": 0.01},
         "<-": {"learning_rate": 1.0e-3, "learning_rate_ft": 1.0e-3, "weights_decay": 0.0005,
                "gradient_moment": 0.9}},
        {"name": "drop3", "type": "dropout", "dropout_ratio": 0.5},
        {"name": "softmax4", "type": "softmax",
         "->": {"weights_stddev": 0.01},
         "<-": {"learning_rate": 1.0e-3, "learning_rate_ft": 1.0e-3, "weights_decay": 0.0005,
                "gradient_moment": 0.9}}]})


def split_layers(layers):
    """→ (AE blocks, inter, tail): ``blocks[k]`` = layer dicts between the k-th
    ae_begin/ae_end pair, ``inter[k]`` = plain layers that precede block k (k >= 1, e.g. the
    ``activation_mul`` between blocks), ``tail`` = everything after the last ``ae_end``."""
    blocks, inter, cur, pending = [], [], None, []
    for l in layers:
        t = l.get("type")
        if t == "ae_begin":
            if cur is not None:
                raise ValueError("nested ae_begin")
            if not blocks and pending:
                raise ValueError("layers before the first ae_begin are not supported")
            inter.append(pending)
            pending, cur = [], []
        elif t == "ae_end":
            if cur is None:
                raise ValueError("ae_end without ae_begin")
            blocks.append(cur)
            cur = None
        elif cur is not None:
            cur.append(l)
        else:
            pending.append(l)
    if cur is not None:
        raise ValueError("ae_begin without ae_end")
    if not blocks:
        raise ValueError("no ae_begin/ae_end block in layers")
    return blocks, inter, pending


class Destroyer(object):
    """Removes units from a workflow (the reference's ``Destroyer`` end unit + the
    ``del_ref``/``unlink_all`` calls scattered through ``adjust_workflow``)."""

    @staticmethod
    def destroy(workflow, *units):
        for u in units:
            if u is None:
                continue
            u.unlink_all()
            workflow.del_ref(u)


class ImagenetAEWorkflow(StandardWorkflow):
    def __init__(self, workflow, **kwargs):
        self.from_snapshot_add_layer = kwargs.get("from_snapshot_add_layer", True)
        self.add_epochs = kwargs.get("add_epochs", 0)
        self.fine_tuning_noise = kwargs.get("fine_tuning_noise", 1.0e-6)
        self.decision_mse_config = dict(kwargs.get("decision_mse_config", {}))
        self.decision_gd_config = dict(kwargs.get("decision_gd_config", {}))
        self.ae_blocks, self.inter_layers, self.tail_layers = split_layers(kwargs["layers"])
        self.stage = 0
        self.fine_tuning = False
        self.encoder_layers = []       # layer dicts of self.forwards (encoder part)
        self.decoder = []
        self.ae_gds = []
        self.target_normalizer = NoneNormalizer()
        kwargs = dict(kwargs)
        kwargs["layers"] = [l for i, b in enumerate(self.ae_blocks)
                            for l in self.inter_layers[i] + b] + self.tail_layers
        kwargs.setdefault("loss_function", "mse")
        super().__init__(workflow, **kwargs)

    # -- construction ----------------------------------------------------------------------------
    def create_workflow(self):
        self.link_repeater(self.start_point)
        self.link_loader(self.repeater)
        # raw uint8 pixels are normalised on the device (x - mean) * rdisp like the reference
        # (imagenet_ae.py:187 ``link_meandispnorm``) when the loader carries mean/rdisp
        self.data_source = (self.loader, "minibatch_data")
        if hasattr(self.real_loader, "rdisp"):
            self.link_meandispnorm(self.loader)
            self.data_source = (self.meandispnorm, "output")
        del self.forwards[:]
        self._build_stage()

    def _new_forward(self, layer, depool=False):
        tpe, kwargs, _ = self._get_layer_type_kwargs(layer)
        cls = self.layer_map[tpe].forward
        if depool:
            cls = TO_DEPOOL.get(cls, cls)
        unit = cls(self, **kwargs)
        if self.forwards:
            prev = self.forwards[-1]
            unit.link_from(prev)
            unit.link_attrs(prev, ("input", "output"))
        else:
            unit.link_from(self.data_source[0])
            unit.link_attrs(self.data_source[0], ("input", self.data_source[1]))
        if isinstance(unit, pooling.StochasticPoolingBase) or hasattr(unit, "minibatch_class"):
            try:
                unit.link_attrs(self.loader, "minibatch_class")
            except Exception:     # units without that attribute
                pass
        self.forwards.append(unit)
        self.encoder_layers.append(layer)
        return unit

    def _build_stage(self):
        """Append block ``self.stage`` + its decoder + MSE loop + GDs of that block only."""
        block = self.ae_blocks[self.stage]
        for layer in self.inter_layers[self.stage]:      # frozen glue between the blocks
            self._new_forward(layer)
        first = len(self.forwards)
        block_input_unit = self.forwards[-1] if self.forwards else self.data_source[0]
        block_input_attr = "output" if self.forwards else self.data_source[1]
        for i, layer in enumerate(block):
            self._new_forward(layer, depool=(i == len(block) - 1))
        block_units = self.forwards[first:]
        if not any(isinstance(u, conv.Conv) for u in block_units):
            raise ValueError("an auto-encoder block needs at least one conv layer")
        # decoder: deconvolutions tied to the block's convolutions, in reverse order
        self.decoder = []
        prev = self.forwards[-1]
        for u in reversed(block_units):
            if not isinstance(u, conv.Conv):
                continue
            d = deconv.Deconv(self, unsafe_padding=True, name="de_" + u.name)
            d.link_from(prev)
            d.link_attrs(u, "weights")
            d.link_conv_attrs(u)
            d.link_attrs(prev, ("input", "output"))
            d.link_attrs(u, ("output_shape_source", "input"))
            self.decoder.append((d, u))
            prev = d
        last = prev
        # loss against the block's own input
        self.loss_function = "mse"
        self.unlink_unit("evaluator")
        self.evaluator = evaluator.EvaluatorMSE(self)
        self.evaluator.link_from(last)
        self.evaluator.link_attrs(last, "output")
        self.evaluator.link_attrs(self.loader, ("batch_size", "minibatch_size"))
        self.evaluator.link_attrs(block_input_unit, ("target", block_input_attr))
        self.evaluator.link_attrs(self, ("normalizer", "target_normalizer"))
        self.unlink_unit("decision")
        cfg = dict(self.decision_mse_config)
        self.decision = decision.DecisionMSE(self, **cfg)
        self.decision.link_from(self.evaluator)
        self.decision.link_attrs(self.loader, "minibatch_class", "last_minibatch",
                                 "minibatch_size", "class_lengths", "epoch_ended",
                                 "epoch_number")
        self.decision.link_attrs(self.evaluator, ("minibatch_metrics", "metrics"))
        self.decision.autoencoder = True
        self.repeater.gate_block = self.decision.complete
        self.real_loader.gate_block = self.decision.complete
        self.link_snapshotter(self.decision)
        # GD chain over the decoder (last deconvolution first)
        self.ae_gds = []
        prev_gd = None
        for d, u in reversed(self.decoder):
            layer = self.encoder_layers[self.forwards.index(u)]
            _, _, gkw = self._get_layer_type_kwargs(layer)
            gkw = {k: v for k, v in gkw.items() if k not in ("name", "learning_rate_ft")}
            g = gd_deconv.GDDeconv(self, name="gd_" + d.name, **gkw)
            self._chain_gd(g, prev_gd)
            g.link_attrs(d, "weights", "input", "hits", "n_kernels", "kx", "ky", "sliding",
                         "padding", "unpack_size")
            g.forward_unit = u
            prev_gd = g
        # like the reference (imagenet_ae.py:587-606: ``assert len(self.gds) == 1``) only the
        # decoder is differentiated in an auto-encoder stage: the tied weights learn through
        # the deconvolution, nothing is back-propagated into the encoder half
        self.ae_gds[-1].need_err_input = False
        del self.gds[:]
        self.gds.extend(self.ae_gds)
        self.repeater.link_from(self.ae_gds[-1])
        self.link_end_point(self.ae_gds[-1])

    def _chain_gd(self, g, prev_gd):
        if prev_gd is None:
            g.link_from(self.snapshotter)
            g.link_attrs(self.evaluator, "err_output")
        else:
            g.link_from(prev_gd)
            g.link_attrs(prev_gd, ("err_output", "err_input"))
        g.gate_skip = self.decision.gd_skip
        self.ae_gds.append(g)

    def _new_gd(self, layer, fwd, prev_gd, fine_tuning=False):
        tpe, _, gkw = self._get_layer_type_kwargs(layer)
        try:
            cls = next(self.layer_map[tpe].backwards)
        except StopIteration:
            return None
        gkw = dict(gkw)
        ft = gkw.pop("learning_rate_ft", None)
        if fine_tuning and ft is not None:
            gkw["learning_rate"] = ft
            gkw["learning_rate_bias"] = gkw.get("learning_rate_ft_bias", ft)
        gkw.pop("learning_rate_ft_bias", None)
        if "name" in gkw:
            gkw["name"] = "gd_" + gkw["name"]
        g = cls(self, **gkw)
        self._chain_gd(g, prev_gd)
        attrs = {"input", "weights", "bias", "input_offset", "mask", "output"}
        if isinstance(g, ConvolutionalBase):
            attrs.update(ConvolutionalBase.CONV_ATTRS)
        if isinstance(g, GDPooling):
            attrs.update(GDPooling.POOL_ATTRS)
        attrs.update(getattr(fwd, "GD_LINK_ATTRS", ()))
        g.link_attrs(fwd, *[a for a in sorted(attrs) if hasattr(fwd, a)])
        g.forward_unit = fwd
        return g

    # -- graph surgery ---------------------------------------------------------------------------
    def _strip_training_units(self):
        Destroyer.destroy(self, *[d for d, _ in self.decoder])
        Destroyer.destroy(self, *self.ae_gds)
        self.decoder, self.ae_gds = [], []
        del self.gds[:]
        self.unlink_unit("evaluator")
        self.unlink_unit("decision")
        self.unlink_unit("snapshotter")
        self.end_point.unlink_before()
        self.segments_ = []
        self.fused_step_ = None

    def _undepool_last(self):
        """Stochastic(Abs)PoolingDepooling → plain pooling with the same geometry."""
        old = self.forwards[-1]
        cls = FROM_DEPOOL.get(type(old))
        if cls is None:
            return
        new = cls(self, kx=old.kx, ky=old.ky, sliding=tuple(old.sliding), name=old.name)
        src = self.forwards[-2] if len(self.forwards) > 1 else self.data_source[0]
        new.link_from(src)
        new.link_attrs(src, ("input", "output" if len(self.forwards) > 1
                             else self.data_source[1]))
        new.link_attrs(self.loader, "minibatch_class")
        Destroyer.destroy(self, old)
        self.forwards[-1] = new

    @property
    def has_more_blocks(self):
        return self.stage + 1 < len(self.ae_blocks)

    def add_ae_layer(self):
        """Freeze the trained block and stack the next auto-encoder on top of it."""
        if not self.has_more_blocks:
            raise ValueError("all auto-encoder blocks are already trained")
        self._strip_training_units()
        self._undepool_last()
        self.stage += 1
        self._build_stage()
        self._is_initialized = False

    def switch_to_fine_tuning(self):
        """Auto-encoder stack → softmax classifier trained end to end."""
        self._strip_training_units()
        self._undepool_last()
        for layer in self.tail_layers:
            self._new_forward(layer)
        self.fine_tuning = True
        self.loss_function = "softmax"
        self.evaluator = evaluator.EvaluatorSoftmax(self)
        last = self.forwards[-1]
        self.evaluator.link_from(last)
        self.evaluator.link_attrs(last, "output", "max_idx")
        self.evaluator.link_attrs(self.loader, ("batch_size", "minibatch_size"),
                                  ("labels", "minibatch_labels"),
                                  ("max_samples_per_epoch", "total_samples"), "class_lengths",
                                  ("offset", "minibatch_offset"))
        self.decision = decision.DecisionGD(self, **self.decision_gd_config)
        self.decision.link_from(self.evaluator)
        self.decision.link_attrs(self.loader, "minibatch_class", "last_minibatch",
                                 "minibatch_size", "class_lengths", "epoch_ended",
                                 "epoch_number")
        self.decision.link_attrs(
            self.evaluator, ("minibatch_n_err", "n_err"),
            ("minibatch_confusion_matrix", "confusion_matrix"),
            ("minibatch_max_err_y_sum", "max_err_output_sum"))
        self.repeater.gate_block = self.decision.complete
        self.real_loader.gate_block = self.decision.complete
        self.link_snapshotter(self.decision)
        from ..workflow.standard_workflow_base import _LastLayerSizer
        self.real_loader.on_initialized = _LastLayerSizer(self, last)
        self.ae_gds = []
        prev_gd = None
        for fwd, layer in reversed(list(zip(self.forwards, self.encoder_layers))):
            g = self._new_gd(layer, fwd, prev_gd, fine_tuning=True)
            if g is not None:
                prev_gd = g
        self.ae_gds[-1].need_err_input = False
        self.gds.extend(self.ae_gds)
        self.repeater.link_from(self.ae_gds[-1])
        self.link_end_point(self.ae_gds[-1])
        self._is_initialized = False

    # -- life cycle ------------------------------------------------------------------------------
    def initialize(self, device=None, **kwargs):
        restored = kwargs.get("snapshot", False)
        if restored and self.from_snapshot_add_layer and not self.fine_tuning:
            if self.has_more_blocks:
                self.info("Restored from a snapshot: stacking auto-encoder block %d",
                          self.stage + 2)
                self.add_ae_layer()
            else:
                self.info("Restored from a snapshot: switching to fine tuning")
                self.switch_to_fine_tuning()
            kwargs = dict(kwargs, snapshot=False)
            # trained weights stay; the loader must re-run its first epoch bookkeeping
        elif restored:
            self.decision.max_epochs += self.add_epochs
            self.decision.complete <<= False
        res = super().initialize(device=device, **kwargs)
        if self.fine_tuning and self.fine_tuning_noise:
            self._perturb_tail()
        return res

    def _perturb_tail(self):
        """Tiny noise on the pre-trained weights breaks exact symmetries before fine tuning
        (``fine_tuning_noise`` in the reference config)."""
        from ..core import prng
        for f in self.forwards:
            w = getattr(f, "weights", None)
            if w is not None and w and not getattr(f, "_ft_noised", False):
                w.map_write()
                w.mem += prng.get().normal(0, self.fine_tuning_noise, w.mem.shape) \
                    .astype(w.mem.dtype)
                w.unmap()
                f._ft_noised = True
                if getattr(f, "on_cuda", False):
                    f.refresh_shadows()


def DummyLauncherFactory():
    """A stand-alone launcher for workflows restored from a snapshot in scripts/tests."""
    from ..core.workflow import DummyLauncher
    return DummyLauncher()


def kwargs_from_config():
    c = root.imagenet_ae
    return dict(loader_name=c.loader_name, loader_config=c.loader,
                decision_mse_config=c.decision_mse, decision_gd_config=c.decision_gd,
                snapshotter_config=c.snapshotter, layers=c.layers,
                from_snapshot_add_layer=c.from_snapshot_add_layer, add_epochs=c.add_epochs,
                fine_tuning_noise=c.fine_tuning_noise)


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = kwargs_from_config()
    kw.update(overrides)
    for k in ("decision_mse_config", "decision_gd_config"):
        if hasattr(kw[k], "to_dict"):
            kw[k] = kw[k].to_dict()
    return ImagenetAEWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    load(ImagenetAEWorkflow, **kwargs_from_config())
    main()
