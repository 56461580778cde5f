"""Structure of the CUDA-only graph fusion passes (workflow/fusion.py), checked on the CPU with a
stand-in device object: which units get which marks for the CIFAR caffe and AlexNet layer lists."""
from veles.znicz_b200.core.config import root
from veles.znicz_b200.workflow import fusion


class FakeCudaDevice(object):
    is_cuda = True


def _marks(wf):
    fwd = {f.name: f for f in wf.forwards}
    gds = {g.forward_unit.name: g for g in wf.gds if g is not None and g.forward_unit is not None}
    return fwd, gds


def test_fusion_passes_on_cifar_caffe():
    from veles.znicz_b200.models import cifar
    root.common.disable.snapshotting = True
    try:
        wf = cifar.build(loader_config={"minibatch_size": 10, "n_train": 40, "n_valid": 0,
                                        "n_test": 0})
        dev = FakeCudaDevice()
        n_act = fusion.fuse_activations(wf, dev)
        n_der = fusion.fuse_backward_derivatives(wf, dev)
        n_ev = fusion.fuse_evaluator(wf, dev)
        assert (n_act, n_der, n_ev) == (3, 2, 1)
        fwd, gds = _marks(wf)
        names = [f.name for f in wf.forwards]
        convs = [n for n in names if n.startswith("conv")]
        pools = [n for n in names if n.startswith("pool")]
        # relu1 folds into pool1 (conv1 stays linear), relu2 / relu3 fold into conv2 / conv3
        assert fwd[pools[0]].__dict__.get("fused_act_") == 3
        assert fwd[convs[0]].__dict__.get("fused_act_", 0) == 0
        assert fwd[convs[1]].__dict__.get("fused_act_") == 3
        assert fwd[convs[2]].__dict__.get("fused_act_") == 3
        # the derivative of conv2 / conv3 moves into the backward of pool2 / pool3
        assert gds[convs[1]].__dict__.get("deriv_upstream_") is True
        assert gds[convs[2]].__dict__.get("deriv_upstream_") is True
        assert not gds[convs[0]].__dict__.get("deriv_upstream_")
        assert gds[pools[1]].__dict__.get("in_deriv_act_") == 3
        assert gds[pools[2]].__dict__.get("in_deriv_act_") == 3
        assert not gds[pools[0]].__dict__.get("in_deriv_act_")
        # softmax layer carries the evaluator
        assert wf.forwards[-1].__dict__.get("fused_eval_") is wf.evaluator
        # the numpy device (or the switch) clears every mark
        fusion.fuse_activations(wf, None)
        assert fusion.fuse_backward_derivatives(wf, None) == 0
        assert fusion.fuse_evaluator(wf, None) == 0
        assert not any(u.__dict__.get("fused_act_") or u.__dict__.get("in_deriv_act_") or
                       u.__dict__.get("deriv_upstream_") or u.__dict__.get("fused_eval_")
                       for u in list(wf.forwards) + [g for g in wf.gds if g is not None])
    finally:
        root.common.disable.snapshotting = False


def test_derivative_fusion_on_alexnet():
    """conv1 / conv2 / conv5 feed max-pooling layers (their derivative moves there); conv3 / conv4
    feed the next convolution, whose dgrad epilogue applies it."""
    from veles.znicz_b200.models import alexnet
    root.common.disable.snapshotting = True
    try:
        wf = alexnet.build(
            loader_name="synthetic_imagenet", layers=alexnet.alexnet_layers(n_classes=10),
            loader_config={"minibatch_size": 2, "shape": (67, 67, 3), "n_classes": 10,
                           "n_train": 4, "n_valid": 0})
        dev = FakeCudaDevice()
        fusion.fuse_activations(wf, dev)
        n = fusion.fuse_backward_derivatives(wf, dev)
        marked = [type(g).__name__ for g in wf.gds
                  if g is not None and g.__dict__.get("in_deriv_act_")]
        relieved = [g.forward_unit.name for g in wf.gds
                    if g is not None and g.__dict__.get("deriv_upstream_")]
        assert n == len(marked) == len(relieved) == 5
        assert sum(m in ("GDMaxPooling", "LRNormalizerBackward") for m in marked) == 3
        assert sum("Conv" in m for m in marked) == 2        # conv4 / conv5 dgrad epilogues
    finally:
        root.common.disable.snapshotting = False
