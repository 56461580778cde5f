"""CIFAR-10 sample workflows (north-star config 3).

Parity: /root/reference/samples/CIFAR10/cifar.py:47-118 and its three configs
(``cifar_caffe_config.py:52-145`` — the benchmark net: conv32-5p2 / maxpool3s2 / relu /
LRN / conv32-5p2 / relu / avgpool3s2 / LRN / conv64-5p2 / relu / avgpool3s2 / softmax,
batch 100, SGD momentum 0.9, L2 5e-4, factor_ortho 1e-3, ``arbitrary_step`` LR;
``cifar_config.py:48-86`` FC-sincos; ``cifar_nin_config.py:60-248`` NiN).
"""
from __future__ import annotations

import os
import pickle

import numpy

from ..core.config import root
from ..loader.base import TEST, VALID, TRAIN
from ..loader.fullbatch import FullBatchLoader
from ..workflow.standard_workflow import StandardWorkflow


def caffe_layers():
    gd = {"learning_rate": 0.001, "learning_rate_bias": 0.002,
          "weights_decay": 0.0005, "weights_decay_bias": 0.0005,
          "factor_ortho": 0.001, "gradient_moment": 0.9, "gradient_moment_bias": 0.9}

    def conv(name, n, std):
        return {"name": name, "type": "conv",
                "->": {"n_kernels": n, "kx": 5, "ky": 5, "padding": (2, 2, 2, 2),
                       "sliding": (1, 1), "weights_filling": "gaussian",
                       "weights_stddev": std, "bias_filling": "constant",
                       "bias_stddev": 0},
                "<-": dict(gd)}
    c3 = conv("conv3", 64, 0.01)
    c3["<-"]["learning_rate_bias"] = 0.001
    return [
        conv("conv1", 32, 0.0001),
        {"name": "pool1", "type": "max_pooling",
         "->": {"kx": 3, "ky": 3, "sliding": (2, 2)}},
        {"name": "relu1", "type": "activation_str"},
        {"name": "norm1", "type": "norm", "alpha": 0.00005, "beta": 0.75, "n": 3, "k": 1},
        conv("conv2", 32, 0.01),
        {"name": "relu2", "type": "activation_str"},
        {"name": "pool2", "type": "avg_pooling",
         "->": {"kx": 3, "ky": 3, "sliding": (2, 2)}},
        {"name": "norm2", "type": "norm", "alpha": 0.00005, "beta": 0.75, "n": 3, "k": 1},
        c3,
        {"name": "relu3", "type": "activation_str"},
        {"name": "pool3", "type": "avg_pooling",
         "->": {"kx": 3, "ky": 3, "sliding": (2, 2)}},
        {"name": "fc_softmax4", "type": "softmax",
         "->": {"output_sample_shape": 10, "weights_filling": "gaussian",
                "weights_stddev": 0.01, "bias_filling": "constant", "bias_stddev": 0},
         "<-": {"learning_rate": 0.001, "learning_rate_bias": 0.002,
                "weights_decay": 1.0, "weights_decay_bias": 0,
                "gradient_moment": 0.9, "gradient_moment_bias": 0.9}}]


def fc_layers():
    """cifar_config.py: FC486-sincos ×2 → softmax (batch 81)."""
    fc = {"->": {"output_sample_shape": 486, "weights_filling": "uniform",
                 "weights_stddev": 0.05, "bias_filling": "uniform", "bias_stddev": 0.05},
          "<-": {"learning_rate": 0.0005, "weights_decay": 0.0, "gradient_moment": 0.9,
                 "factor_ortho": 0.001}}
    return [dict(fc, name="fc_linear1", type="all2all"),
            {"name": "sincos1", "type": "activation_sincos"},
            dict(fc, name="fc_linear2", type="all2all"),
            {"name": "sincos2", "type": "activation_sincos"},
            {"name": "fc_softmax3", "type": "softmax",
             "->": {"output_sample_shape": 10, "weights_filling": "uniform",
                    "weights_stddev": 0.05, "bias_filling": "uniform",
                    "bias_stddev": 0.05},
             "<-": {"learning_rate": 0.0005, "weights_decay": 0.0,
                    "gradient_moment": 0.9}}]


def nin_layers():
    """cifar_nin_config.py: Network-in-Network (9 convs, dropout, avg-pool head)."""
    def conv(name, n, k, pad, std=0.05, lr=0.01):
        return {"name": name, "type": "conv_str",
                "->": {"n_kernels": n, "kx": k, "ky": k, "padding": (pad,) * 4,
                       "sliding": (1, 1), "weights_filling": "gaussian",
                       "weights_stddev": std, "bias_filling": "constant",
                       "bias_stddev": 0},
                "<-": {"learning_rate": lr, "learning_rate_bias": 2 * lr,
                       "weights_decay": 0.0001, "weights_decay_bias": 0,
                       "gradient_moment": 0.9, "gradient_moment_bias": 0.9}}
    return [
        conv("conv1", 192, 5, 2), conv("cccp1", 160, 1, 0), conv("cccp2", 96, 1, 0),
        {"name": "pool1", "type": "max_pooling",
         "->": {"kx": 3, "ky": 3, "sliding": (2, 2)}},
        {"name": "drop1", "type": "dropout", "dropout_ratio": 0.5},
        conv("conv2", 192, 5, 2), conv("cccp3", 192, 1, 0), conv("cccp4", 192, 1, 0),
        {"name": "pool2", "type": "avg_pooling",
         "->": {"kx": 3, "ky": 3, "sliding": (2, 2)}},
        {"name": "drop2", "type": "dropout", "dropout_ratio": 0.5},
        conv("conv3", 192, 3, 1), conv("cccp5", 192, 1, 0), conv("cccp6", 10, 1, 0),
        {"name": "pool3", "type": "avg_pooling",
         "->": {"kx": 8, "ky": 8, "sliding": (1, 1)}},
        {"name": "fc_softmax", "type": "softmax",
         "->": {"output_sample_shape": 10, "weights_filling": "gaussian",
                "weights_stddev": 0.05, "bias_filling": "constant", "bias_stddev": 0},
         "<-": {"learning_rate": 0.01, "learning_rate_bias": 0.02,
                "weights_decay": 0.0001, "gradient_moment": 0.9,
                "gradient_moment_bias": 0.9}}]


root.cifar.update({
    "loader_name": "synthetic_cifar",
    "decision": {"fail_iterations": 250, "max_epochs": 1000000000},
    "lr_adjuster": {
        "do": True, "lr_policy_name": "arbitrary_step",
        "bias_lr_policy_name": "arbitrary_step",
        "lr_parameters": {
            "lrs_with_lengths": [(1, 60000), (0.1, 5000), (0.01, 100000000)]},
        "bias_lr_parameters": {
            "lrs_with_lengths": [(1, 60000), (0.1, 5000), (0.01, 100000000)]}},
    "snapshotter": {"prefix": "cifar_caffe", "interval": 1},
    "loss_function": "softmax",
    "add_plotters": False,
    "image_saver": {"do": False},
    "loader": {"minibatch_size": 100, "normalization_type": "internal_mean",
               "shuffle_limit": 2000000000},
    "weights_plotter": {"limit": 256},
    "similar_weights_plotter": {"form_threshold": 1.1, "peak_threshold": 0.5,
                                "magnitude_threshold": 0.65},
    "layers": caffe_layers()})


class CifarLoader(FullBatchLoader):
    """Loads the python-pickle CIFAR-10 batches: 3×32×32 planes → HWC
    (/root/reference/samples/CIFAR10/cifar.py:47-66). ``data_path`` must hold
    ``data_batch_1..5`` and ``test_batch``."""
    MAPPING = "cifar_loader"

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.data_path = kwargs.get("data_path", os.path.join(
            str(root.common.dirs.datasets), "cifar-10-batches-py"))

    def _read(self, name):
        with open(os.path.join(self.data_path, name), "rb") as fin:
            d = pickle.load(fin, encoding="latin1")
        data = numpy.asarray(d["data"], dtype=numpy.uint8).reshape(-1, 3, 32, 32)
        return data.transpose(0, 2, 3, 1), list(d["labels"])

    def load_data(self):
        vd, vl = self._read("test_batch")
        td, tl = [], []
        for i in range(1, 6):
            d, l = self._read("data_batch_%d" % i)
            td.append(d)
            tl.extend(l)
        td = numpy.concatenate(td)
        self.class_lengths[TEST] = 0
        self.class_lengths[VALID] = len(vl)
        self.class_lengths[TRAIN] = len(tl)
        self.original_data.reset(
            numpy.concatenate([vd, td]).astype(self.dtype))
        self.original_labels = list(vl) + list(tl)


class CifarWorkflow(StandardWorkflow):
    """Self-constructing CIFAR-10 model (layers come from the config)."""

    def create_workflow(self):
        self.link_repeater(self.start_point)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [self.link_snapshotter(self.decision)]
        if root.cifar.image_saver.get("do", False):
            end_units.append(self.link_image_saver(self.decision))
        if root.cifar.get("add_plotters", False):
            end_units.extend(link(self.decision) for link in (
                self.link_error_plotter, self.link_conf_matrix_plotter,
                self.link_err_y_plotter))
            end_units.append(self.link_weights_plotter("weights", self.decision))
            self.link_gds(*end_units)
            last = self.gds[0]
        else:
            self.link_gds(*end_units)
            last = self.gds[0]
        if self.config.lr_adjuster.get("do", False):
            last = self.link_lr_adjuster(last)
        self.repeater.link_from(last)
        self.link_end_point(last)


def build(launcher=None, **overrides):
    """Public factory: ``build(loader_config={...}, layers=...) -> CifarWorkflow``."""
    from ..core.workflow import DummyLauncher
    kw = dict(
        decision_config=root.cifar.decision, snapshotter_config=root.cifar.snapshotter,
        loader_config=root.cifar.loader, layers=root.cifar.layers,
        loader_name=root.cifar.loader_name, loss_function=root.cifar.loss_function,
        lr_adjuster_config=root.cifar.lr_adjuster,
        weights_plotter_config=root.cifar.weights_plotter,
        similar_weights_plotter_config=root.cifar.similar_weights_plotter)
    kw.update(overrides)
    return CifarWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    load(CifarWorkflow,
         decision_config=root.cifar.decision, snapshotter_config=root.cifar.snapshotter,
         loader_config=root.cifar.loader, layers=root.cifar.layers,
         loader_name=root.cifar.loader_name, loss_function=root.cifar.loss_function,
         lr_adjuster_config=root.cifar.lr_adjuster)
    main()
