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


def test_wine_relu():
    from veles.znicz_b200.models import wine_relu
    wf = wine_relu.build(decision_config={"max_epochs": 25, "fail_iterations": 50},
                         snapshotter_config={"prefix": "wr", "interval": 1000,
                                             "time_interval": 1e9})
    assert type(wf.forwards[0]).__name__ == "All2AllRELU"
    wf.initialize(device="numpy")
    wf.run()
    assert wf.decision.best_n_err_pt[2] < 15.0


def test_mnist_simple_with_diff_stats():
    import pickle
    from veles.znicz_b200.models import mnist_simple
    root.mnist_simple.decision.max_epochs = 3
    wf = mnist_simple.build(loader_name="synthetic_mnist", layers=[30, 10],
                            loader_config={"minibatch_size": 20, "n_train": 200,
                                           "n_valid": 60, "noise": 0.3})
    wf.initialize(device="numpy")
    wf.run()
    assert wf.decision.best_n_err_pt[1] < 30.0
    assert wf.diff_stats.size > 0 and wf.slaves_plotter.records
    with open(wf.diff_stats.file_name, "rb") as f:
        stats = pickle.load(f)
    assert any("gradient_weights" in v for v in stats.values())
    assert len(wf.plt[1].values) >= 3


def test_stl10_workflow(tmp_path):
    pytest.importorskip("cv2")
    from veles.znicz_b200.models import stl10
    rs = numpy.random.RandomState(7)
    (tmp_path / "class_names.txt").write_text("a b c\n")
    protos = rs.randint(0, 255, (3, 3, 96, 96))
    for stem, n in (("train", 24), ("test", 9)):
        y = rs.randint(1, 4, n).astype(numpy.uint8)
        x = numpy.clip(protos[y - 1] + rs.randn(n, 3, 96, 96) * 20, 0, 255).astype(numpy.uint8)
        x.tofile(str(tmp_path / (stem + "_X.bin")))
        y.tofile(str(tmp_path / (stem + "_y.bin")))
    wf = stl10.build(loader_config={"directory": str(tmp_path), "minibatch_size": 8,
                                    "scale": (32, 32), "normalization_type": "internal_mean"},
                     decision_config={"max_epochs": 1, "fail_iterations": 5},
                     snapshotter_config={"prefix": "stl_t", "interval": 1000,
                                         "time_interval": 1e9})
    wf.initialize(device="numpy")
    assert wf.loader.original_data.shape == (33, 32, 32, 3)
    assert wf.forwards[-1].output.shape == (8, 3)
    wf.run()
    assert bool(wf.decision.complete)


def test_mnist_rbm_workflow():
    from veles.znicz_b200.models import mnist_rbm
    wf = mnist_rbm.build(minibatch_size=32, n_samples=96, h_size=40, max_epochs=3,
                         learning_rate=0.05)
    wf.initialize(device="numpy")
    w0 = wf.forwards[1].weights.mem.copy()
    wf.run()
    assert bool(wf.decision.complete)
    assert numpy.abs(wf.forwards[1].weights.mem - w0).max() > 1e-5
    assert numpy.isfinite(wf.forwards[1].weights.mem).all()


def test_mnist_ae_workflow():
    from veles.znicz_b200.models import mnist_ae
    wf = mnist_ae.build(loader_name="synthetic_mnist", max_epochs=3, learning_rate=0.0005,
                        loader_config={"minibatch_size": 10, "n_train": 60, "n_valid": 20,
                                       "noise": 0.3, "normalization_type": "linear"})
    wf.initialize(device="numpy")
    assert wf.deconv.output.shape == wf.loader.minibatch_data.shape
    assert wf.deconv.weights is wf.conv.weights or \
        wf.deconv.weights.mem is wf.conv.weights.mem
    w0 = wf.conv.weights.mem.copy()
    wf.run()
    assert bool(wf.decision.complete)
    assert numpy.abs(wf.conv.weights.mem - w0).max() > 0
    assert wf.plt[2].pics and wf.plt[-1].pics


def _tiny_imagenet(tmp_path, n_val=8, n_train=24, side=40, n_classes=4):
    import json
    import pickle
    rs = numpy.random.RandomState(3)
    n = n_val + n_train
    labels = rs.randint(0, n_classes, n)
    protos = rs.randint(0, 255, (n_classes, side, side, 3))
    samples = numpy.clip(protos[labels] + rs.randn(n, side, side, 3) * 25, 0, 255) \
        .astype(numpy.uint8)
    samples.tofile(str(tmp_path / "samples.dat"))
    with open(tmp_path / "labels.pickle", "wb") as f:
        pickle.dump([("n%03d" % l, int(l)) for l in labels], f)
    with open(tmp_path / "count.json", "w") as f:
        json.dump({"test": 0, "val": n_val, "train": n_train}, f)
    mean = samples[n_val:].mean(axis=0)
    with open(tmp_path / "matrixes.pickle", "wb") as f:
        pickle.dump([mean, numpy.ones_like(mean, dtype=numpy.float32) / 64], f)
    return dict(sx=side, sy=side, crop_size_sx=32, crop_size_sy=32, mirror=True, channels=3,
                minibatch_size=8, normalization_type="none",
                original_labels_filename=str(tmp_path / "labels.pickle"),
                count_samples_filename=str(tmp_path / "count.json"),
                samples_filename=str(tmp_path / "samples.dat"),
                matrixes_filename=str(tmp_path / "matrixes.pickle"))


def _shrink(layers, n_classes, div=16, fc=32):
    """Same topology, far fewer kernels/neurons so the CPU test stays quick."""
    out = []
    for l in layers:
        l = dict(l)
        if "->" in l:
            fwd = dict(l["->"])
            if "n_kernels" in fwd:
                fwd["n_kernels"] = max(4, fwd["n_kernels"] // div)
                if fwd["kx"] == 11:
                    fwd.update(kx=5, ky=5, sliding=(2, 2))
            if "output_sample_shape" in fwd:
                fwd["output_sample_shape"] = n_classes if l["type"] == "softmax" else fc
            l["->"] = fwd
        out.append(l)
    return out


def test_alexnet_family_topologies():
    from veles.znicz_b200.models import alexnet
    a, n, v = alexnet.alexnet_layers(), alexnet.nin_layers(), alexnet.vgga_layers()
    assert [l["type"] for l in a].count("conv_str") == 5
    assert [l["type"] for l in a].count("zero_filter") == 4
    assert [l["type"] for l in n].count("conv") == 12 and n[-2]["type"] == "avg_pooling"
    assert [l["type"] for l in v].count("conv_str") == 13
    assert [l["type"] for l in v].count("max_pooling") == 5
    assert a[0]["->"]["sliding"] == (4, 4) and a[-1]["->"]["output_sample_shape"] == 1000


def test_alexnet_workflow_tiny(tmp_path):
    from veles.znicz_b200.models import alexnet
    loader = _tiny_imagenet(tmp_path)
    layers = _shrink(alexnet.alexnet_layers(), 4, div=16)
    # the 3/2 pools of the real net need >= 3x3 maps: use padding-friendly small maps
    wf = alexnet.build(
        loader_config=loader, layers=layers,
        decision_config={"max_epochs": 2, "fail_iterations": 10},
        snapshotter_config={"prefix": "alex_t", "interval": 1000, "time_interval": 1e9},
        lr_adjuster_config={"lr_policy_name": "arbitrary_step",
                            "bias_lr_policy_name": "arbitrary_step",
                            "lr_parameters": {"lrs_with_lengths": [(1, 3), (0.1, 100)]},
                            "bias_lr_parameters": {"lrs_with_lengths": [(1, 3), (0.1, 100)]}})
    wf.initialize(device="numpy")
    names = [type(f).__name__ for f in wf.forwards]
    assert names.count("ZeroFiller") == 4 and names.count("DropoutForward") == 2
    lr0 = wf.gds[0].learning_rate
    wf.run()
    assert bool(wf.decision.complete)
    assert wf.gds[0].learning_rate < lr0                 # arbitrary_step kicked in
    # grouping: conv2 only sees its own half of the input channels
    conv2 = [f for f in wf.forwards if f.name == "conv_str2_forward"][0]
    w = conv2.weights.mem.reshape(conv2.n_kernels, conv2.ky, conv2.kx, -1)
    # the mask zeroes kernel % g == channel % g (/root/reference/weights_zerofilling.py:95-98)
    k_idx = numpy.arange(conv2.n_kernels)[:, None] % 2
    c_idx = numpy.arange(w.shape[-1])[None, :] % 2
    off = (k_idx == c_idx)[:, None, None, :]
    assert numpy.all(w[numpy.broadcast_to(off, w.shape)] == 0)
    assert numpy.any(w[numpy.broadcast_to(~off, w.shape)] != 0)


def test_imagenet_ae_stages_and_fine_tuning(tmp_path):
    """Greedy AE pre-training: stage 1 → (snapshot) → stack stage 2 → (snapshot) →
    fine tuning as a softmax classifier; encoder weights carry over between stages."""
    import glob
    from veles.znicz_b200.core.snapshotter import SnapshotterToFile
    from veles.znicz_b200.models import imagenet_ae
    loader = _tiny_imagenet(tmp_path, side=24)
    for k in ("crop_size_sx", "crop_size_sy", "mirror"):
        loader.pop(k)
    layers = [
        {"type": "ae_begin"},
        {"name": "conv1", "type": "conv",
         "->": {"n_kernels": 6, "kx": 5, "ky": 5, "sliding": (1, 1), "include_bias": False,
                "weights_filling": "gaussian", "weights_stddev": 0.05},
         "<-": {"learning_rate": 1e-5, "learning_rate_ft": 1e-3, "weights_decay": 0.0}},
        {"name": "pool1", "type": "stochastic_abs_pooling",
         "->": {"kx": 2, "ky": 2, "sliding": (2, 2)}},
        {"type": "ae_end"},
        {"name": "mul1", "type": "activation_mul"},
        {"type": "ae_begin"},
        {"name": "conv2", "type": "conv",
         "->": {"n_kernels": 8, "kx": 3, "ky": 3, "sliding": (1, 1), "include_bias": False,
                "weights_filling": "gaussian", "weights_stddev": 0.05},
         "<-": {"learning_rate": 1e-5, "learning_rate_ft": 1e-3, "weights_decay": 0.0}},
        {"type": "ae_end"},
        {"name": "fc3", "type": "all2all_tanh", "->": {"output_sample_shape": 16},
         "<-": {"learning_rate": 1e-2, "learning_rate_ft": 1e-2}},
        {"name": "softmax4", "type": "softmax", "<-": {"learning_rate": 1e-2}}]
    blocks, inter, tail = imagenet_ae.split_layers(layers)
    assert [len(b) for b in blocks] == [2, 1] and len(inter[1]) == 1 and len(tail) == 2
    snap_dir = str(tmp_path / "snaps")
    wf = imagenet_ae.build(
        loader_name="imagenet_loader_base", loader_config=loader, layers=layers,
        decision_mse_config={"max_epochs": 2, "fail_iterations": 10},
        decision_gd_config={"max_epochs": 2, "fail_iterations": 10},
        snapshotter_config={"prefix": "iae", "interval": 1, "time_interval": 0,
                            "directory": snap_dir, "compression": ""})
    wf.initialize(device="numpy")
    assert [type(f).__name__ for f in wf.forwards] == ["Conv", "StochasticAbsPoolingDepooling"]
    assert len(wf.decoder) == 1 and wf.decoder[0][0].weights.mem is wf.forwards[0].weights.mem
    w_before = wf.forwards[0].weights.mem.copy()
    wf.run()
    assert bool(wf.decision.complete)
    w1 = wf.forwards[0].weights.mem.copy()
    assert numpy.abs(w1 - w_before).max() > 0
    snaps = sorted(f for f in glob.glob(os.path.join(snap_dir, "iae*.pickle"))
                   if not os.path.islink(f))
    assert snaps
    # -- restore → graph surgery stacks block 2 ------------------------------------------
    wf2 = SnapshotterToFile.import_file(snaps[-1])
    wf2.workflow = imagenet_ae.DummyLauncherFactory()
    w1 = wf2.forwards[0].weights.mem.copy()        # the snapshot holds the best epoch
    wf2.initialize(device="numpy", snapshot=True)
    assert wf2.stage == 1
    assert [type(f).__name__ for f in wf2.forwards] == [
        "Conv", "StochasticAbsPooling", "ForwardMul", "Conv"]
    numpy.testing.assert_array_equal(wf2.forwards[0].weights.mem, w1)
    assert all(g.forward_unit is not wf2.forwards[0] for g in wf2.gds)   # block 1 is frozen
    wf2.run()
    numpy.testing.assert_array_equal(wf2.forwards[0].weights.mem, w1)
    # -- fine tuning ---------------------------------------------------------------------
    wf2.switch_to_fine_tuning()
    wf2.initialize(device="numpy")
    assert type(wf2.forwards[-1]).__name__ == "All2AllSoftmax"
    assert wf2.forwards[-1].output.shape[1] == 4
    assert len(wf2.gds) == 6 and wf2.gds[0].learning_rate == 1e-2
    conv_gd = [g for g in wf2.gds if g.forward_unit is wf2.forwards[0]][0]
    assert conv_gd.learning_rate == 1e-3                  # learning_rate_ft applied
    wf2.run()
    assert bool(wf2.decision.complete)
    assert numpy.abs(wf2.forwards[0].weights.mem - w1).max() > 0


def test_hands_and_channels(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from veles.znicz_b200.models import image_classifiers as ic
    rs = numpy.random.RandomState(8)
    # Hands: headerless .raw grey images in Training/<class>/ and Testing/<class>/
    for split, n in (("Training", 12), ("Testing", 4)):
        for ci, cname in enumerate(("Positive", "Negative")):
            d = tmp_path / "hands" / split / cname
            os.makedirs(d)
            for k in range(n):
                img = (rs.rand(16, 16) * 60).astype(numpy.uint8)
                img[:, 8 * ci:8 * ci + 8] += 150          # bright left / right half
                img.tofile(str(d / ("h%02d.raw" % k)))
    wf = ic.build_hands(
        loader_config=dict(root.hands.loader.to_dict(), minibatch_size=8, raw_shape=(16, 16),
                           train_paths=[str(tmp_path / "hands" / "Training")],
                           validation_paths=[str(tmp_path / "hands" / "Testing")]),
        decision_config={"max_epochs": 12, "fail_iterations": 20},
        image_saver_config={"out_dirs": [str(tmp_path / ("hs%d" % i)) for i in range(3)]},
        snapshotter_config={"prefix": "hands_t", "interval": 1000, "time_interval": 1e9})
    wf.initialize(device="numpy")
    assert wf.loader.original_data.shape == (32, 16, 16, 1)
    wf.run()
    assert wf.decision.best_n_err_pt[1] < 30.0
    # TvChannels: HSV + Sobel channel, aspect-preserving scale on a background
    for ci, cname in enumerate(("first", "second", "third")):
        d = tmp_path / "channels" / cname
        os.makedirs(d)
        for k in range(8):
            img = numpy.full((30, 50, 3), 30, numpy.uint8)
            cv2.circle(img, (10 + 12 * ci, 15), 6, (40 + 90 * ci, 200 - 60 * ci, 90), -1)
            img = numpy.clip(img + rs.randn(30, 50, 3) * 6, 0, 255).astype(numpy.uint8)
            cv2.imwrite(str(d / ("c%02d.png" % k)), img)
    wf = ic.build_channels(
        loader_config=dict(root.channels.loader.to_dict(), minibatch_size=6, scale=(32, 32),
                           train_paths=[str(tmp_path / "channels")]),
        layers=[{"name": "fc_tanh1", "type": "all2all_tanh", "->": {"output_sample_shape": 16},
                 "<-": {"learning_rate": 0.01, "weights_decay": 0.0}},
                {"name": "fc_softmax2", "type": "softmax",
                 "<-": {"learning_rate": 0.01, "weights_decay": 0.0}}],
        decision_config={"max_epochs": 6, "fail_iterations": 10},
        image_saver_config={"out_dirs": [str(tmp_path / ("cs%d" % i)) for i in range(3)]},
        snapshotter_config={"prefix": "chan_t", "interval": 1000, "time_interval": 1e9})
    wf.initialize(device="numpy")
    assert wf.loader.original_data.shape[1:] == (32, 32, 4)       # HSV + sobel
    wf.run()
    assert bool(wf.decision.complete)


def test_spam_kohonen(tmp_path):
    from veles.znicz_b200.models import spam_kohonen as sk
    path = sk.generate_dataset(str(tmp_path / "spam.txt.xz"), n=160, n_lemmas=24)
    out = str(tmp_path / "classified.txt")
    root.spam_kohonen.forward.shape = (4, 4)
    wf = sk.build(file=path, ids=True, classes=True, minibatch_size=40, epochs=5,
                  export_file=out)
    wf.initialize(device="numpy")
    assert wf.loader.original_data.shape[0] == 160 and len(wf.loader.ids) == 160
    wf.run()
    assert bool(wf.decision.complete)
    lines = open(out).read().split("\n")
    assert len([l for l in lines if l]) == 160 and lines[0].startswith("msg")
    # the two synthetic topics occupy different regions of the map
    assert wf.validator.fitness > 0.6
