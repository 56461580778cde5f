"""StandardWorkflow end-to-end on the numpy backend with synthetic data
(mirrors /root/reference/tests/functional/test_mnist_conv.py / test_cifar_caffe.py:
train N epochs, check improvement, snapshot, resume)."""
import glob
import os

import numpy
import pytest

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.snapshotter import SnapshotterToFile
from veles.znicz_b200.core.workflow import DummyLauncher
from veles.znicz_b200.models import cifar, mnist
from veles.znicz_b200.workflow.standard_workflow_base import StandardWorkflowBase


def _fast_layers():
    layers = cifar.caffe_layers()
    for l in layers:
        if "<-" in l:
            l["<-"].update(learning_rate=0.02, learning_rate_bias=0.02,
                           weights_decay=0.0005)
        if l["type"] == "conv":
            l["->"]["weights_stddev"] = 0.05
    return layers


def _small_cifar(max_epochs=3):
    return cifar.build(
        layers=_fast_layers(),
        loader_config={"minibatch_size": 20, "n_train": 200, "n_valid": 60,
                       "normalization_type": "internal_mean", "noise": 0.3},
        decision_config={"max_epochs": max_epochs, "fail_iterations": 50},
        snapshotter_config={"prefix": "cifar_t", "interval": 1, "time_interval": 0,
                            "compression": "gz"})


def test_cifar_caffe_numpy_trains_and_resumes():
    wf = _small_cifar(4)
    wf.initialize(device="numpy")
    assert len(wf.forwards) == 12 and len(wf.gds) == 12
    assert wf.forwards[0].output.shape == (20, 32, 32, 32)
    assert wf.forwards[-1].input.shape == (20, 4, 4, 64)
    wf.run()
    dec = wf.decision
    assert bool(dec.complete)
    assert wf.loader.epoch_number == 4
    # synthetic classes are separable: validation error must drop well below chance
    assert dec.best_n_err_pt[1] < 60.0, dec.best_n_err_pt
    snaps = sorted(glob.glob(os.path.join(root.common.dirs.snapshots, "cifar_t_*")),
                   key=os.path.getmtime)
    assert snaps
    wf2 = SnapshotterToFile.import_file(snaps[-1])
    wf2.workflow = DummyLauncher()
    ep = wf2.loader.epoch_number
    wf2.decision.max_epochs = ep + 1
    wf2.decision.complete <<= False
    wf2.initialize(device="numpy", snapshot=True)
    wf2.run()
    assert wf2.loader.epoch_number == ep + 1


def test_mnist_conv_numpy_one_epoch():
    wf = mnist.build(
        loader_config={"minibatch_size": 6, "n_train": 60, "n_valid": 24,
                       "normalization_type": "linear"},
        decision_config={"max_epochs": 2, "fail_iterations": 10},
        snapshotter_config={"prefix": "mnist_t", "interval": 1, "time_interval": 0,
                            "compression": ""})
    wf.initialize(device="numpy")
    assert wf.forwards[2].weights.shape == (87, 1600)
    assert wf.forwards[4].weights.shape == (791, 1392)
    wf.run()
    assert bool(wf.decision.complete)
    assert numpy.isfinite(wf.forwards[0].weights.mem).all()


def test_mcdnnic_topology_parsing():
    wf = StandardWorkflowBase(
        DummyLauncher(), loader_name="synthetic_image",
        loader_config={"shape": (16, 16, 1), "n_classes": 4, "n_train": 24,
                       "n_valid": 12},
        mcdnnic_topology="12x16x16-8C4-MP2-6C4-MP3-32N-4N")
    layers = wf.layers
    assert [l["type"] for l in layers] == [
        "conv", "max_pooling", "conv", "max_pooling", "all2all", "softmax"]
    assert layers[0]["->"] == {"n_kernels": 8, "kx": 4, "ky": 4}
    assert layers[-1]["->"]["output_sample_shape"] == 4
    with pytest.raises(ValueError):
        StandardWorkflowBase(DummyLauncher(), loader_name="synthetic_image",
                             mcdnnic_topology="bad-topology")
    with pytest.raises(ValueError):
        StandardWorkflowBase(DummyLauncher(), loader_name="synthetic_image",
                             layers=[{"type": "conv"}],
                             mcdnnic_topology="12x16x16-8C4-4N")


def test_avatar_mirrors_loader_exports():
    """`link_avatar` (/root/reference/standard_workflow.py:386-404): downstream units read the
    avatar's stable copy of the loader's exports while the loader may run ahead."""
    import numpy
    from veles.znicz_b200.core.avatar import Avatar
    from veles.znicz_b200.core.memory import Array
    from veles.znicz_b200.core.mutable import Bool
    from veles.znicz_b200.core.units import TrivialUnit
    from veles.znicz_b200.core.workflow import DummyWorkflow
    wf = DummyWorkflow()
    real = TrivialUnit(wf, name="loader")
    real.minibatch_data = Array(numpy.arange(6, dtype=numpy.float32).reshape(2, 3))
    real.last_minibatch = Bool(False)
    real.minibatch_size = 2
    av = Avatar(wf)
    av.reals[real] = ("minibatch_data", "last_minibatch", "minibatch_size")
    av.initialize()
    assert numpy.array_equal(av.minibatch_data.mem, real.minibatch_data.mem)
    assert av.minibatch_data is not real.minibatch_data
    real.minibatch_data.map_write()
    real.minibatch_data.mem[...] = -1.0              # the real loader already serves the next batch
    real.last_minibatch <<= True
    real.minibatch_size = 1
    assert av.minibatch_data.mem[0, 1] == 1.0 and not bool(av.last_minibatch) and av.minibatch_size == 2
    av.run()
    assert (av.minibatch_data.mem == -1.0).all() and bool(av.last_minibatch) and av.minibatch_size == 1


def test_extract_forward_workflow_carries_trained_weights():
    """`extract_forward_workflow` (/root/reference/standard_workflow.py:210-286): a forward-only
    workflow over another loader with the trained parameters."""
    import numpy
    from veles.znicz_b200.models import mnist
    wf = mnist.build(
        layers=mnist.fc_layers(), loader_name="synthetic_mnist",
        loader_config={"minibatch_size": 10, "n_train": 40, "n_valid": 20, "noise": 0.3,
                       "normalization_type": "linear"},
        decision_config={"max_epochs": 2, "fail_iterations": 5})
    wf.initialize(device="numpy")
    wf.run()
    fwd = wf.extract_forward_workflow(
        loader_name="synthetic_mnist",
        loader_config={"minibatch_size": 10, "n_train": 0, "n_valid": 0, "n_test": 20,
                       "noise": 0.3, "normalization_type": "linear"}, cyclic=False)
    assert len(fwd.forwards) == len(wf.forwards)
    fwd.initialize(device="numpy")
    for a, b in zip(wf.forwards, fwd.forwards):
        assert numpy.array_equal(a.weights.mem, b.weights.mem) and a.weights is not b.weights
        assert numpy.array_equal(a.bias.mem, b.bias.mem)
        assert b.forward_mode
    fwd.run()
    out = fwd.forwards[-1].output.mem
    assert out.shape == (10, 10) and numpy.allclose(out.sum(axis=1), 1.0, atol=1e-4)
