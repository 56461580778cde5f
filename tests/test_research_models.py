"""Research workflows on synthetic / generated data, numpy backend."""
import os

import numpy
import pytest

from veles.znicz_b200.core.config import root


@pytest.fixture(autouse=True)
def _dirs(tmp_path):
    old = (root.common.dirs.cache, root.common.dirs.snapshots)
    root.common.dirs.cache = str(tmp_path / "cache")
    root.common.dirs.snapshots = str(tmp_path / "snap")
    os.makedirs(root.common.dirs.cache, exist_ok=True)
    os.makedirs(root.common.dirs.snapshots, exist_ok=True)
    yield
    root.common.dirs.cache, root.common.dirs.snapshots = old


def test_mnist7_segment_targets():
    from veles.znicz_b200.models import mnist7
    root.mnist7.decision.max_epochs = 5
    wf = mnist7.build(loader_name="synthetic_mnist7", layers=[32, 7],
                      loader_config={"minibatch_size": 20, "n_train": 200, "n_valid": 60,
                                     "noise": 0.3},
                      add_plotters=True)
    wf.initialize(device="numpy", learning_rate=0.01)
    assert wf.loader.class_targets.shape == (10, 7)
    assert wf.loader.original_targets.shape == (260, 7)
    wf.run()
    assert bool(wf.decision.complete)
    # nearest-target accuracy is tracked like a classification error
    assert wf.decision.best_n_err_pt[1] is not None and wf.decision.best_n_err_pt[1] < 60.0
    assert wf.plotters[0].values and wf.plotters[-1].val_mse.sum() > 0


def test_approximator(tmp_path):
    import scipy.io
    from veles.znicz_b200.models import approximator
    rs = numpy.random.RandomState(4)
    x = rs.randn(300, 12).astype(numpy.float32)
    w = rs.randn(12, 3).astype(numpy.float32)
    y = numpy.tanh(x.dot(w) * 0.3)
    scipy.io.savemat(str(tmp_path / "dec.mat"), {"dec": x})
    numpy.save(str(tmp_path / "org.npy"), y)
    root.approximator.decision.max_epochs = 12
    wf = approximator.build(layers=[16, 3], loader_config={
        "minibatch_size": 25, "train_paths": [str(tmp_path / "dec.mat")],
        "target_paths": [str(tmp_path / "org.npy")]})
    wf.initialize(device="numpy", learning_rate=0.02)
    assert list(wf.loader.class_lengths) == [0, 45, 255]
    wf.run()
    hist = wf.decision.best_mse
    assert hist[1] is not None and hist[1] < 1.0


def test_video_ae(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from veles.znicz_b200.models import video_ae
    d = tmp_path / "img"
    os.makedirs(d)
    rs = numpy.random.RandomState(2)
    base = cv2.GaussianBlur(rs.rand(18, 32).astype(numpy.float32), (0, 0), 3)
    for i in range(40):
        frame = numpy.roll(base, i, axis=1)
        frame = (frame - frame.min()) / (frame.max() - frame.min()) * 255
        cv2.imwrite(str(d / ("frame%03d.png" % i)), frame.astype(numpy.uint8))
    root.video_ae.decision.max_epochs = 4
    wf = video_ae.build(layers=[9, [18, 32]], loader_config={
        "minibatch_size": 10, "train_paths": [str(d)]})
    wf.initialize(device="numpy", learning_rate=0.002)
    assert wf.loader.original_data.shape == (40, 18, 32)
    assert wf.forwards[-1].output.shape[1:] == (18, 32)
    wf.run()
    assert bool(wf.decision.complete) and wf.decision.best_mse[2] is not None
