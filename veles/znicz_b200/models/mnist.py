"""MNIST sample workflows (north-star config 2 = ``mnist_conv_config``).

Parity: /root/reference/samples/MNIST/mnist.py:53-140 and configs
``mnist_config.py:44-86`` (FC-tanh 100 → softmax), ``mnist_caffe_config.py:48-101``
(LeNet, batch 64, ``inv`` LR), ``mnist_conv_config.py:48-117`` (GA-tuned conv64-5 /
mp2 / conv87-5 / mp2 / FC791-softplus / softmax, **batch 6**, momentum + L2 + ortho).
"""
from __future__ import annotations

import os
import struct

import numpy

from ..core.config import root
from ..loader.base import TEST, VALID, TRAIN
from ..loader.fullbatch import FullBatchLoader
from ..workflow.standard_workflow import StandardWorkflow


def conv_layers():
    return [
        {"name": "conv1", "type": "conv",
         "->": {"n_kernels": 64, "kx": 5, "ky": 5, "sliding": (1, 1),
                "weights_filling": "uniform", "weights_stddev": 0.0944569801138958,
                "bias_filling": "constant", "bias_stddev": 0.048},
         "<-": {"learning_rate": 0.03, "learning_rate_bias": 0.358,
                "gradient_moment": 0.36508255921752014, "gradient_moment_bias": 0.385,
                "weights_decay": 0.0005, "weights_decay_bias": 0.1980997902551238,
                "factor_ortho": 0.001}},
        {"name": "max_pool1", "type": "max_pooling",
         "->": {"kx": 2, "ky": 2, "sliding": (2, 2)}},
        {"name": "conv2", "type": "conv",
         "->": {"n_kernels": 87, "kx": 5, "ky": 5, "sliding": (1, 1),
                "weights_filling": "uniform", "weights_stddev": 0.067,
                "bias_filling": "constant", "bias_stddev": 0.444},
         "<-": {"learning_rate": 0.03, "learning_rate_bias": 0.381,
                "gradient_moment": 0.115, "gradient_moment_bias": 0.741,
                "weights_decay": 0.0005, "factor_ortho": 0.001,
                "weights_decay_bias": 0.039}},
        {"name": "max_pool2", "type": "max_pooling",
         "->": {"kx": 2, "ky": 2, "sliding": (2, 2)}},
        {"name": "fc_relu3", "type": "all2all_relu",
         "->": {"output_sample_shape": 791, "weights_stddev": 0.039,
                "bias_filling": "constant", "weights_filling": "uniform",
                "bias_stddev": 1.0},
         "<-": {"learning_rate": 0.03, "learning_rate_bias": 0.196,
                "gradient_moment": 0.81, "gradient_moment_bias": 0.619,
                "weights_decay": 0.0005, "factor_ortho": 0.001,
                "weights_decay_bias": 0.11487830567238211}},
        {"name": "fc_softmax4", "type": "softmax",
         "->": {"output_sample_shape": 10, "weights_filling": "uniform",
                "weights_stddev": 0.024, "bias_filling": "constant",
                "bias_stddev": 0.255},
         "<-": {"learning_rate": 0.03, "learning_rate_bias": 0.488,
                "gradient_moment": 0.133, "gradient_moment_bias": 0.8422143625658985,
                "weights_decay": 0.0005, "weights_decay_bias": 0.476}}]


def caffe_layers():
    """LeNet (mnist_caffe_config.py)."""
    gd = {"learning_rate": 0.01, "learning_rate_bias": 0.02, "gradient_moment": 0.9,
          "gradient_moment_bias": 0, "weights_decay": 0.0005, "weights_decay_bias": 0}
    return [
        {"name": "conv1", "type": "conv",
         "->": {"n_kernels": 20, "kx": 5, "ky": 5, "sliding": (1, 1),
                "weights_filling": "uniform", "bias_filling": "constant",
                "bias_stddev": 0}, "<-": dict(gd)},
        {"name": "pool1", "type": "max_pooling",
         "->": {"kx": 2, "ky": 2, "sliding": (2, 2)}},
        {"name": "conv2", "type": "conv",
         "->": {"n_kernels": 50, "kx": 5, "ky": 5, "sliding": (1, 1),
                "weights_filling": "uniform", "bias_filling": "constant",
                "bias_stddev": 0}, "<-": dict(gd)},
        {"name": "pool2", "type": "max_pooling",
         "->": {"kx": 2, "ky": 2, "sliding": (2, 2)}},
        {"name": "fc_relu3", "type": "all2all_relu",
         "->": {"output_sample_shape": 500, "weights_filling": "uniform",
                "bias_filling": "constant", "bias_stddev": 0}, "<-": dict(gd)},
        {"name": "fc_softmax4", "type": "softmax",
         "->": {"output_sample_shape": 10, "weights_filling": "uniform",
                "bias_filling": "constant", "bias_stddev": 0}, "<-": dict(gd)}]


def fc_layers():
    """mnist_config.py: FC-tanh 100 → softmax."""
    return [
        {"name": "fc_tanh1", "type": "all2all_tanh",
         "->": {"output_sample_shape": 100, "weights_filling": "uniform",
                "weights_stddev": 0.05, "bias_filling": "uniform", "bias_stddev": 0.05},
         "<-": {"learning_rate": 0.03, "weights_decay": 0.0, "gradient_moment": 0.0}},
        {"name": "fc_softmax2", "type": "softmax",
         "->": {"output_sample_shape": 10, "weights_filling": "uniform",
                "weights_stddev": 0.05, "bias_filling": "uniform", "bias_stddev": 0.05},
         "<-": {"learning_rate": 0.03, "weights_decay": 0.0, "gradient_moment": 0.0}}]


root.mnistr.update({
    "loss_function": "softmax",
    "loader_name": "synthetic_mnist",
    "lr_adjuster": {"do": True, "lr_policy_name": "inv", "bias_lr_policy_name": "inv",
                    "lr_parameters": {"gamma": 0.0001, "pow_ratio": 0.75},
                    "bias_lr_parameters": {"gamma": 0.0001, "pow_ratio": 0.75}},
    "decision": {"max_epochs": 10000000, "fail_iterations": 100},
    "snapshotter": {"prefix": "mnist_conv", "time_interval": 0, "compression": ""},
    "loader": {"minibatch_size": 6, "normalization_type": "linear"},
    "weights_plotter": {"limit": 64},
    "export_wf": False,
    "layers": conv_layers()})


class MnistLoader(FullBatchLoader):
    """IDX parser with header checks 2049/2051 (/root/reference/loader/loader_mnist.py:
    113-186): VALID = t10k (10000), TRAIN = 60000. ``data_path`` holds the four raw
    IDX files (no download here — there is no network)."""
    MAPPING = "mnist_loader"
    FILES = {"test_labels": "t10k-labels.idx1-ubyte", "test_images": "t10k-images.idx3-ubyte",
             "train_labels": "train-labels.idx1-ubyte",
             "train_images": "train-images.idx3-ubyte"}

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.data_path = kwargs.get("data_path", os.path.join(
            str(root.common.dirs.datasets), "MNIST"))

    @staticmethod
    def read_idx_labels(path, expected=None):
        with open(path, "rb") as fin:
            header, = struct.unpack(">i", fin.read(4))
            if header != 2049:
                raise ValueError("Wrong header in file with labels")
            n, = struct.unpack(">i", fin.read(4))
            if expected is not None and n != expected:
                raise ValueError("Wrong number of labels in %s" % path)
            arr = numpy.frombuffer(fin.read(n), dtype=numpy.uint8)
        if arr.size != n:
            raise ValueError("Truncated labels file %s" % path)
        if arr.min() < 0 or arr.max() > 9:
            raise ValueError("Wrong labels range in train dataset.")
        return arr.astype(numpy.int32)

    @staticmethod
    def read_idx_images(path, expected=None):
        with open(path, "rb") as fin:
            header, = struct.unpack(">i", fin.read(4))
            if header != 2051:
                raise ValueError("Wrong header in file with images")
            n, rows, cols = struct.unpack(">iii", fin.read(12))
            if expected is not None and n != expected:
                raise ValueError("Wrong number of images in %s" % path)
            if rows != 28 or cols != 28:
                raise ValueError("Wrong images size in %s, should be 28*28" % path)
            arr = numpy.frombuffer(fin.read(n * rows * cols), dtype=numpy.uint8)
        if arr.size != n * rows * cols:
            raise ValueError("Truncated images file %s" % path)
        return arr.reshape(n, rows, cols, 1)

    def load_data(self):
        p = self.data_path
        vl = self.read_idx_labels(os.path.join(p, self.FILES["test_labels"]))
        vi = self.read_idx_images(os.path.join(p, self.FILES["test_images"]), len(vl))
        tl = self.read_idx_labels(os.path.join(p, self.FILES["train_labels"]))
        ti = self.read_idx_images(os.path.join(p, self.FILES["train_images"]), len(tl))
        self.class_lengths[TEST] = 0
        self.class_lengths[VALID] = len(vl)
        self.class_lengths[TRAIN] = len(tl)
        self.original_data.reset(numpy.concatenate([vi, ti]).astype(self.dtype))
        self.original_labels = vl.tolist() + tl.tolist()


class MnistWorkflow(StandardWorkflow):
    """Self-constructing MNIST model."""

    def __init__(self, workflow, **kwargs):
        self.export_wf = kwargs.get("export_wf", False)
        self.package_name = kwargs.get("package_name", os.path.join(
            str(root.common.dirs.user), "mnist.zip"))
        super().__init__(workflow, **kwargs)

    def create_workflow(self):
        self.link_repeater(self.start_point)
        self.link_loader(self.repeater)
        self.link_forwards(("input", "minibatch_data"), self.loader)
        self.link_evaluator(self.forwards[-1])
        self.link_decision(self.evaluator)
        end_units = [link(self.decision) for link in (
            self.link_snapshotter, self.link_error_plotter,
            self.link_conf_matrix_plotter, self.link_err_y_plotter)]
        self.link_gds(*end_units)
        last = self.gds[0]
        if self.config.lr_adjuster.get("do", False):
            last = self.link_lr_adjuster(last)
        self.link_end_point(last)
        self.repeater.link_from(last)

    def on_workflow_finished(self):
        super().on_workflow_finished()
        if self.export_wf:
            self.package_export(self.package_name, precision=16)


def build(launcher=None, **overrides):
    from ..core.workflow import DummyLauncher
    kw = dict(
        decision_config=root.mnistr.decision, snapshotter_config=root.mnistr.snapshotter,
        loader_name=root.mnistr.loader_name, loader_config=root.mnistr.loader,
        layers=root.mnistr.layers, loss_function=root.mnistr.loss_function,
        lr_adjuster_config=root.mnistr.lr_adjuster)
    kw.update(overrides)
    return MnistWorkflow(launcher or DummyLauncher(), **kw)


def run(load, main):
    load(MnistWorkflow,
         decision_config=root.mnistr.decision, snapshotter_config=root.mnistr.snapshotter,
         loader_name=root.mnistr.loader_name, loader_config=root.mnistr.loader,
         layers=root.mnistr.layers, loss_function=root.mnistr.loss_function,
         lr_adjuster_config=root.mnistr.lr_adjuster,
         export_wf=root.mnistr.export_wf)
    main()
